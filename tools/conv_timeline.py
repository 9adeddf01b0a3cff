#!/usr/bin/env python
"""Per-CTA phase timeline of the tcgen05 conv kernel (clock64 stamps via pg_debug_set_buffer): prologue, pipeline fill, main loop,
converter finish, epilogue.  usage: conv_timeline.py [--cin 128 --cout 128 --res 128 --k 3 --up 1 --n 16]"""
import argparse, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# the phase stamps exist only in the -DPG_DEBUG build of the library (pasta-gan_b200/build.py --debug); load that one in this process
os.environ.setdefault('PASTA_B200_LIB', os.path.join(ROOT, 'pasta-gan_b200', 'lib', 'libpasta_b200_dbg.so'))
if not os.path.exists(os.environ['PASTA_B200_LIB']):
    import subprocess
    subprocess.run([sys.executable, os.path.join(ROOT, 'pasta-gan_b200', 'build.py'), '--debug'], check=True)
import pasta_gan_b200
from pasta_gan_b200.torch_utils.ops import conv_igemm, upfirdn2d
ap = argparse.ArgumentParser()
for k, d in dict(cin=128, cout=128, res=128, k=3, up=1, n=16, tma=0).items():
    ap.add_argument('--' + k, type=int, default=d)
a = ap.parse_args()
dev = torch.device('cuda:0')
x = torch.randn(a.n, a.cin, a.res, a.res, device=dev)
w = torch.randn(a.cout, a.cin, a.k, a.k, device=dev) / (a.cin * a.k * a.k) ** 0.5
f = upfirdn2d.setup_filter([1, 3, 3, 1]).to(dev)
lib = pasta_gan_b200.capi.load()
xin = conv_igemm.to_c8(x) if a.tma else x
run = lambda: conv_igemm.conv2d_igemm(xin, w, f=f if a.up == 2 else None, up=a.up)
with torch.no_grad():
    for _ in range(3):
        run()
    buf = torch.zeros(16 * 8192, dtype=torch.int64, device=dev)
    lib.pg_debug_set_buffer(buf.data_ptr())
    run(); torch.cuda.synchronize()
    lib.pg_debug_set_buffer(None)
t = buf.view(-1, 16).cpu()
t = t[t[:, 0] != 0].double()
names = ['prologue (start -> setup sync)', 'fill (setup -> first A stage ready at the MMA thread)', 'main loop (first A ready -> last MMA issued)',
         'converters done (setup -> last stage stored)', 'accumulator ready seen by epilogue (setup -> acc_full)', 'epilogue (acc_full -> stores done)', 'total',
         'MMA thread: cycles waiting on A stages', 'MMA thread: cycles waiting on B slots', 'converter warp 0: cycles waiting on free A stages',
         'converter warp 0: cycles issuing loads (non-pipelined mode)', 'converter warp 0: cycles converting + storing incl. waiting for the loaded data and free stages']
vals = [t[:, 1] - t[:, 0], t[:, 2] - t[:, 1], t[:, 3] - t[:, 2], t[:, 6] - t[:, 1], t[:, 4] - t[:, 1], t[:, 5] - t[:, 4], t[:, 5] - t[:, 0], t[:, 8], t[:, 9], t[:, 10], t[:, 14], t[:, 15]]
print(f'{t.shape[0]} CTAs; cycles (median / p10 / p90)')
for nme, v in zip(names, vals):
    q = torch.quantile(v, torch.tensor([0.5, 0.1, 0.9], dtype=torch.float64))
    print(f'  {nme:62s} {q[0]:9.0f} {q[1]:9.0f} {q[2]:9.0f}')

# wall-clock view: globaltimer (ns) per CTA, SM ids -> kernel span, SM clock under load, idle gaps between consecutive CTAs of one SM
gt0, gt1, sm = t[:, 12], t[:, 13], t[:, 11].long()
span_ns = float(gt1.max() - gt0.min())
cyc_per_ns = float(((t[:, 5] - t[:, 0]) / (gt1 - gt0).clamp(min=1)).median())
busy = torch.zeros(int(sm.max()) + 1, dtype=torch.float64).index_add_(0, sm, gt1 - gt0)
print(f'  kernel span {span_ns / 1e3:.1f} us; SM clock under load ~{cyc_per_ns * 1e3:.0f} MHz; CTA-time per SM / span: median {float((busy / span_ns).median()):.2f} '
      f'(2.0 = two resident CTAs busy all the time); CTAs per SM min/max {int(torch.bincount(sm).min())}/{int(torch.bincount(sm).max())}')
