#!/usr/bin/env python
"""Residual-epilogue variants of the TMA conv at one shape (timing with CUDA events, or a few launches for ncu with --ncu)."""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pasta_gan_b200.torch_utils.ops import conv_igemm as K

ap = argparse.ArgumentParser()
ap.add_argument('--c', type=int, default=128); ap.add_argument('--res', type=int, default=128); ap.add_argument('--n', type=int, default=16)
ap.add_argument('--ncu', type=int, default=0)
a = ap.parse_args()
dev = torch.device('cuda:0')
x = K.to_c8(torch.randn(a.n, a.c, a.res, a.res, device=dev).half())
w = torch.randn(a.c, a.c, 3, 3, device=dev) / (a.c * 9) ** 0.5
r32 = torch.randn(a.n, a.c, a.res, a.res, device=dev)
r8 = K.to_c8(r32.half())
variants = {
    'c8 in, f32 out': dict(),
    'c8 in, f32 out, c8 residual': dict(residual=r8),
    'c8 in, f32 out, f32 residual': dict(residual=r32),
    'c8 in, c8 out': dict(out_c8=True),
    'c8 in, c8 out, c8 residual': dict(residual=r8, out_c8=True),
}
with torch.no_grad():
    for name, kw in variants.items():
        for _ in range(1 if a.ncu else 3):
            y = K.conv2d_igemm(x, w, cache_weights=True, **kw)
        if a.ncu:
            continue
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            y = K.conv2d_igemm(x, w, cache_weights=True, **kw)
        e1.record(); torch.cuda.synchronize()
        print(f'{e0.elapsed_time(e1) / 20 * 1e3:8.1f} us  {name}')
torch.cuda.synchronize()
