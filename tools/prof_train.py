#!/usr/bin/env python
"""Where one training iteration spends its time: wall clock per phase, GPU-busy time and the top CUDA kernels (torch.profiler)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import procedural
from pasta_gan_b200 import networks as N
from pasta_gan_b200.training import TryOnTrainer, synth_training_batch

dev = torch.device('cuda:0')
torch.backends.cudnn.benchmark = True
G = N.build_generator_full(); D = N.build_discriminator(num_fp16_res=3)
procedural.fill_(G); procedural.fill_(D)
G.to(dev).train().requires_grad_(True); D.to(dev).train().requires_grad_(True)
tr = TryOnTrainer(G, D)
batch = synth_training_batch(int(sys.argv[1]) if len(sys.argv) > 1 else 4, device=dev)
for _ in range(3):
    tr.step(batch)
torch.cuda.synchronize()
for name, fn in (('g_main', lambda: tr.g_main(batch)), ('d_main', lambda: tr.d_phase(batch, True, False)), ('d_reg', lambda: tr.d_phase(batch, False, True))):
    t0 = time.perf_counter(); fn(); t_cpu = time.perf_counter() - t0; torch.cuda.synchronize(); t_all = time.perf_counter() - t0
    print(f'{name}: host issue {1e3 * t_cpu:.1f} ms, until GPU done {1e3 * t_all:.1f} ms')
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    tr.g_main(batch); tr.d_phase(batch, True, False)
    torch.cuda.synchronize()
ka = prof.key_averages()
tot_cuda = sum(k.self_device_time_total for k in ka) / 1e3
print(f'GPU busy (sum of kernel times) {tot_cuda:.1f} ms for g_main + d_main')
rows = sorted(ka, key=lambda k: -k.self_device_time_total)[:30]
for k in rows:
    print(f'{k.self_device_time_total / 1e3:8.2f} ms  n={k.count:5d}  {k.key[:100]}')
print('--- top CPU ops')
for k in sorted(ka, key=lambda k: -k.self_cpu_time_total)[:15]:
    print(f'{k.self_cpu_time_total / 1e3:8.2f} ms  n={k.count:5d}  {k.key[:100]}')
print('--- library convolutions by shape (device time of the op, ms; these are the calls conv2d_gradfix leaves on cuDNN)')
conv_ops = ('aten::cudnn_convolution', 'aten::cudnn_convolution_transpose', 'aten::convolution_backward')
by = [k for k in prof.key_averages(group_by_input_shape=True) if k.key in conv_ops]
print(f'total {sum(k.device_time_total for k in by) / 1e3:.2f} ms in {sum(k.count for k in by)} calls')
for k in sorted(by, key=lambda k: -k.device_time_total)[:24]:
    print(f'{k.device_time_total / 1e3:8.2f} ms  n={k.count:4d}  {k.key[6:]:28s} {[s_ for s_ in k.input_shapes[:3]]}')
