#!/usr/bin/env python
"""Per-launch breakdown of one eager generator step: every C-ABI launch with its shape tag, CUDA-event time, algorithmic TFLOP/s / GB/s.
usage: step_breakdown.py [--workload gen256|gen512] [--batch 16]"""
import argparse, collections, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import procedural
import pasta_gan_b200
from pasta_gan_b200 import networks as N

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='gen256'); ap.add_argument('--batch', type=int, default=16)
a = ap.parse_args()
dev = torch.device('cuda:0')
capi = pasta_gan_b200.capi
if a.workload == 'gen512':
    G = N.build_generator_512().eval().requires_grad_(False); inp = procedural.synth_inputs_512(a.batch, seed=4321, device=dev)
else:
    G = N.build_generator_full().eval().requires_grad_(False); inp = procedural.synth_inputs(a.batch, seed=1234, device=dev)
procedural.fill_(G)
G.to(dev)
inp = {k: v.to(dev) for k, v in inp.items()}
with torch.no_grad():
    for _ in range(3):
        G(**inp, noise_mode='const')
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); G(**inp, noise_mode='const'); e1.record(); torch.cuda.synchronize()
    print(f'eager step (no profiler): {e0.elapsed_time(e1):.2f} ms')
    torch.cuda._sleep(int(3e7))          # ~15 ms of spin: the host queues every launch before the first executes, so the events bracket kernels, not host gaps
    with capi.LaunchProfiler() as prof:
        G(**inp, noise_mode='const')
    torch.cuda.synchronize()
agg = collections.OrderedDict()
tot = 0.0
for name, nbytes, flops, s0, s1, tag in prof.records:
    ms = s0.elapsed_time(s1)
    tot += ms
    r = agg.setdefault((name, tag), [0, 0.0, 0, 0])
    r[0] += 1; r[1] += ms; r[2] += nbytes; r[3] += flops
print(f'{len(prof.records)} launches of our kernels, {tot:.2f} ms inside spans')
for (name, tag), (n, ms, nb, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{ms:8.3f} ms  n={n:3d}  avg {1e3 * ms / n:7.1f} us  {fl / ms / 1e9 if ms else 0:7.1f} TFLOP/s  {nb / ms / 1e6 if ms else 0:7.0f} GB/s  {name:12s} {tag}')
