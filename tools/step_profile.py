#!/usr/bin/env python
"""One eager generator step between cudaProfilerStart/Stop, for `ncu --profile-from-start off` launch lists:
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/step_profile.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import procedural
from pasta_gan_b200 import networks as N

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device('cuda:0')
torch.backends.cudnn.benchmark = True
G = N.build_generator_full().eval().requires_grad_(False)
procedural.fill_(G)
G.to(dev)
inp = procedural.synth_inputs(batch, device=dev)
with torch.no_grad():
    for _ in range(3):
        G(**inp, noise_mode='const')
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    out = G(**inp, noise_mode='const')
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print('ok', float(out[0].abs().mean()))
