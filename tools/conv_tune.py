#!/usr/bin/env python
"""Tuning sweep of the tcgen05 conv's plan knobs (environment variables read by the C library at each call).
usage: conv_tune.py  ->  one line per (shape, knob setting): median microseconds"""
import itertools, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pasta_gan_b200.torch_utils.ops import conv_igemm, upfirdn2d

dev = torch.device('cuda:0')
f = upfirdn2d.setup_filter([1, 3, 3, 1]).to(dev)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2] * 1e3


def main():
    shapes = [(256, 128, 128, 3, 1), (128, 128, 128, 3, 1), (128, 256, 128, 3, 1), (64, 64, 256, 3, 1), (128, 64, 256, 3, 1), (256, 256, 64, 3, 1),
              (512, 512, 32, 3, 1), (512, 512, 16, 3, 1), (192, 128, 128, 1, 1), (128, 64, 256, 1, 1), (256, 128, 64, 3, 2), (128, 64, 128, 3, 2), (512, 256, 32, 3, 2)]
    knobs = [dict(PIPE=p, NACC=n, PAIR=pr) for p, n, pr in
             [(0, 0, 1), (1, 0, 1), (0, 1, 1), (0, 2, 1), (0, 4, 1), (0, 1, 0), (0, 2, 0), (0, 4, 0), (1, 4, 0)]]
    with torch.no_grad():
        for cin, cout, res, k, up in shapes:
            x = torch.randn(16, cin, res, res, device=dev)
            w = torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5
            out = []
            for kn in knobs:
                for key, v in kn.items():
                    os.environ['PASTA_B200_CONV_' + key] = str(v)
                conv_igemm._pack_cache.clear() if hasattr(conv_igemm, '_pack_cache') else None
                try:
                    t = timeit(lambda: conv_igemm.conv2d_igemm(x, w, f=f if up == 2 else None, up=up, flip_weight=(up == 1)))
                except Exception as e:
                    t = float('nan')
                out.append(f"p{kn['PIPE']}n{kn['NACC']}c{kn['PAIR']}={t:.0f}")
            print(f'{cin}->{cout} @{res} k{k} up{up}: ' + '  '.join(out), flush=True)


if __name__ == '__main__':
    main()
