#!/usr/bin/env python
"""Plan-knob sweep for the down-2 form (FIR + stride-2 3x3 as a space-to-depth 'same' conv).  One line per shape: median microseconds per knob setting."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pasta_gan_b200.torch_utils.ops import conv_igemm, upfirdn2d
from conv_tune import timeit

dev = torch.device('cuda:0')
f = upfirdn2d.setup_filter([1, 3, 3, 1]).to(dev)
shapes = [(64, 128, 512), (64, 64, 512), (128, 256, 256), (64, 128, 256), (256, 256, 128), (64, 64, 256)]
knobs = [dict(NACC=0, PAIR=1), dict(NACC=1, PAIR=1), dict(NACC=2, PAIR=1), dict(NACC=4, PAIR=0), dict(NACC=2, PAIR=0)]
with torch.no_grad():
    for cin, cout, res in shapes:
        x = torch.randn(16, cin, res, res, device=dev)
        w = torch.randn(cout, cin, 3, 3, device=dev) / (cin * 9) ** 0.5
        out = []
        for kn in knobs:
            for key, v in kn.items():
                os.environ['PASTA_B200_CONV_' + key] = str(v)
            try:
                t = timeit(lambda: conv_igemm.conv2d_igemm(x, w, f=f, down=2))
            except Exception as e:
                t = float('nan')
            out.append(f"n{kn['NACC']}c{kn['PAIR']}={t:.0f}")
        print(f'{cin}->{cout} @{res} down2: ' + '  '.join(out), flush=True)
