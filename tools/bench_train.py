#!/usr/bin/env python
"""Training-step benchmark (BASELINE.json configs[3]): GeneratorFull + Discriminator, batch 4 per GPU, Gmain + Dmain every iteration and
Dreg (R1, gamma 10) every 16th, flat NCCL gradient all-reduce per phase.  Run alone or under torchrun:
    python tools/bench_train.py --steps 8
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/bench_train.py
Prints one JSON line on rank 0 (img/s counts real images consumed per second across all ranks)."""
import argparse, json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import procedural
from pasta_gan_b200 import networks as N, data_parallel as dp
from pasta_gan_b200.training import TryOnTrainer, synth_training_batch

ap = argparse.ArgumentParser()
ap.add_argument('--steps', type=int, default=16); ap.add_argument('--warmup', type=int, default=17); ap.add_argument('--batch-gpu', type=int, default=4); ap.add_argument('--graphs', type=int, default=1); ap.add_argument('--out', default=None); ap.add_argument('--allow-tf32', type=int, default=0, help='library convolutions / GEMMs on TF32 (the reference trains with it off)')
a = ap.parse_args()
world, rank, local = int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0))
dev = torch.device('cuda', local); torch.cuda.set_device(dev)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
torch.backends.cudnn.benchmark = True
G = N.build_generator_full(); D = N.build_discriminator(num_fp16_res=3)
procedural.fill_(G); procedural.fill_(D)
G.to(dev).train().requires_grad_(True); D.to(dev).train().requires_grad_(True)
dp.broadcast_parameters(G); dp.broadcast_parameters(D)
tr = TryOnTrainer(G, D, capturable=bool(a.graphs), allow_tf32=bool(a.allow_tf32))
batch = synth_training_batch(a.batch_gpu, seed=1234 + rank, device=dev)
if a.graphs:
    tr.capture(batch)
for _ in range(a.warmup):
    tr.step(batch)
tr.it = 0
torch.cuda.synchronize()
if world > 1: dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    stats = tr.step(batch)
e1.record(); torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    sec = float(t.item())
    line = (json.dumps(dict(metric='training images/sec (G+D step, R1 every 16)', value=world * a.batch_gpu * a.steps / sec, unit='img/s', n_gpus=world,
                          steps=a.steps, ms_per_step=1e3 * sec / a.steps, batch_per_gpu=a.batch_gpu, global_batch=world * a.batch_gpu,
                          allreduce_bytes_per_step=dict(G=tr.g_bucket.nbytes, D=tr.d_bucket.nbytes),
                          losses={k: float(v) for k, v in stats.items()}, max_mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30, cuda_graphs=bool(a.graphs), library_tf32=bool(a.allow_tf32))))
    print(line, flush=True)
    if a.out:
        open(a.out, 'w').write(line + '\n')
if world > 1:
    dist.destroy_process_group()
