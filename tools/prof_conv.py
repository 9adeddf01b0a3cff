#!/usr/bin/env python
"""Tiny driver for ncu: a few launches of the tcgen05 conv at one shape (default: the SPADE-block shape, 128->128 @128^2, N=16)."""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pasta_gan_b200.torch_utils.ops import conv_igemm, upfirdn2d

ap = argparse.ArgumentParser()
ap.add_argument('--cin', type=int, default=128); ap.add_argument('--cout', type=int, default=128)
ap.add_argument('--res', type=int, default=128); ap.add_argument('--k', type=int, default=3)
ap.add_argument('--up', type=int, default=1); ap.add_argument('--mod', type=int, default=0); ap.add_argument('--n', type=int, default=16); ap.add_argument('--iters', type=int, default=4)
a = ap.parse_args()
dev = torch.device('cuda:0')
x = torch.randn(a.n, a.cin, a.res, a.res, device=dev)
w = torch.randn(a.cout, a.cin, a.k, a.k, device=dev) / (a.cin * a.k * a.k) ** 0.5
f = upfirdn2d.setup_filter([1, 3, 3, 1]).to(dev)
b = torch.randn(a.cout, device=dev)
st = (1 + 0.3 * torch.randn(a.n, a.cin, device=dev)) if a.mod else None
dc = torch.rand(a.n, a.cout, device=dev) + 0.5 if a.mod else None
nz = torch.randn(a.res * a.up, a.res * a.up, device=dev) * 0.1 if a.mod else None
with torch.no_grad():
    for _ in range(a.iters):
        y = conv_igemm.conv2d_igemm(x, w, f=f if a.up == 2 else None, up=a.up, styles=st, dcoefs=dc, noise=nz, bias=b, act='lrelu', gain=2 ** 0.5, clamp=256)
torch.cuda.synchronize()
print('ok', float(y.abs().mean()))
