#!/usr/bin/env python
"""bench.py — try-on images/sec of the PASTA-GAN operator hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload gen256|gen512]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A step = one GeneratorFull forward (style/const encoders + mapping + synthesis, eval, noise_mode='const') over one
batch of 16 synthetic 256x192 (padded to 256x256) person / garment / pose / parsing tensors per GPU
(BASELINE.json configs[1]; batch-sharded replicas at N > 1, weak scaling, no data-path collective).

Prints ONE JSON line on rank 0:
  value      whole-job img/s, inputs resident in HBM when the timed region starts (CUDA events, max over ranks)
  e2e        same metric through the public API (TryOnSession.step_from_host): pinned-host H2D of the batch and D2H of the images
             inside the timed region; e2e_u8: the uint8-in / uint8-out form of the same call (device-side normalise and photo conversion)
  roofline   dominant hand-written kernel: algorithmic bytes (or flops) per launch / CUDA-event time per launch vs
             MEASURED_PEAKS.json; measured in one instrumented eager step right after the timed region (per-launch events cannot
             be recorded inside a replayed CUDA graph)
  cpu_baseline   the reference's own impl='ref' CPU path on this box's host cores, bounded sample (rank 0, N = 1): the UNMODIFIED reference tree
             (baseline/_ref, shipped copy) through baseline/run_reference.py when it is present (kind "reference"), else the oracle port (kind "port")
  --impl reference   times that CPU arm alone
  extra      dropin: the unmodified reference networks.py over OUR ops on this GPU, eager, img/s;
             reference_gpu: the unmodified reference with its OWN CUDA plugins (JIT) + cuDNN fp32 on this GPU — the GPU code to beat;
             gen512: the same bench line for BASELINE configs[2] (Generator_512, 512x320)
  gpu_library_baseline   our module tree with every convolution routed to cuDNN fp32 (the reference's library path), eager, same batch
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))

import torch  # noqa: E402

METRIC = 'try-on images/sec (generator inference, 256x192 padded to 256x256)'
UNIT = 'img/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=16, help='images per GPU per step (test.sh: 16)')
    ap.add_argument('--workload', default='gen256', choices=['gen256', 'gen512'], help='gen256 = BASELINE configs[1] (default); gen512 = configs[2]')
    ap.add_argument('--no-graph', action='store_true', help='eager launches instead of CUDA-graph replay')
    ap.add_argument('--skip-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip the drop-in / reference-GPU / gen512 / library-baseline legs')
    return ap.parse_args()


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        return dict(hbm=p['hbm_gbs'], tf=p['bf16_tflops'], tf_sustained=p.get('bf16_tflops_sustained', p['bf16_tflops']), src='measured')
    except Exception:
        return dict(hbm=6650.0, tf=1590.0, tf_sustained=1400.0, src='fallback')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region (B200_PROFILING.md recipe)."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def _nvml_loop(self):
        import pynvml as nv
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        bits = {'hw_slowdown': nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, 'nvmlClocksEventReasonHwSlowdown') else nv.nvmlClocksThrottleReasonHwSlowdown,
                'hw_thermal_slowdown': getattr(nv, 'nvmlClocksEventReasonHwThermalSlowdown', None) or nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                'sw_thermal_slowdown': getattr(nv, 'nvmlClocksEventReasonSwThermalSlowdown', None) or nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                'sw_power_cap': getattr(nv, 'nvmlClocksEventReasonSwPowerCap', None) or nv.nvmlClocksThrottleReasonSwPowerCap}
        get_reasons = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        self.nvml_max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        while not self._stop.is_set():
            self.nvml_sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            r = int(get_reasons(h))
            for nm, b in bits.items():
                if r & int(b):
                    self.nvml_reasons.add(nm)
            self._stop.wait(0.01)

    def start(self):
        # NVML in-process (a sample every ~10 ms: a 0.2 s timed region gets ~20 samples); nvidia-smi -lms as the fall-back
        try:
            import pynvml as nv
            nv.nvmlInit()
            # physical index of this process's device (CUDA_VISIBLE_DEVICES may remap): NVML enumerates all GPUs of the box
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            if vis:
                ids = [v.strip() for v in vis.split(',') if v.strip()]
                if self.index < len(ids) and ids[self.index].isdigit():
                    self.index = int(ids[self.index])
            self.nvml_sm, self.nvml_reasons, self.nvml_max, self._stop = [], set(), None, threading.Event()
            self.thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.thread.start()
            self.proc = 'nvml'
            return
        except Exception:
            self.proc = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '200', '-i', str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        if self.proc == 'nvml':
            self._stop.set()
            self.thread.join(timeout=2)
            sm = sorted(self.nvml_sm)
            return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=self.nvml_max, reasons=sorted(self.nvml_reasons), samples=len(sm), source='nvml')
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [t.strip() for t in ln.split(',')]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        sm.sort()
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


def dist_setup(args):
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and args.impl != 'reference':      # the reference arm runs on rank 0 alone: no rendezvous needed
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29511')
        dist.init_process_group(backend='nccl', rank=rank, world_size=world, device_id=torch.device('cuda', local))
    return world, rank, local


def max_over_ranks(value, world, device):
    if world == 1:
        return value
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


# ------------------------------------------------------------------------------------------------ reference tree (baseline/_ref) legs

HARNESS = os.path.join(ROOT, 'baseline', 'run_reference.py')
HAVE_REF = os.path.isdir(os.path.join(ROOT, 'baseline', '_ref', 'torch_utils'))
DTYPE = 'f16 tensor-core operands, f32 accumulate (TMEM), f32 I/O at the API boundary'


def run_harness(mode, batch, steps, warmup, timeout, check=False, env=None, extra_args=()):
    """One arrangement of the unmodified reference in its own process -> its JSON dict (or {'unavailable': why})."""
    if not HAVE_REF:
        return dict(unavailable='baseline/_ref not present')
    cmd = [sys.executable, HARNESS, '--mode', mode, '--batch', str(batch), '--steps', str(steps), '--warmup', str(warmup)] + (['--check'] if check else []) + list(extra_args)
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=dict(os.environ, **(env or {})))
        lines = [ln for ln in r.stdout.splitlines() if ln.startswith('{')]
        if r.returncode != 0 or not lines:
            return dict(unavailable=f'rc {r.returncode}: {(r.stderr or r.stdout)[-300:]}')
        return json.loads(lines[-1])
    except subprocess.TimeoutExpired:
        return dict(unavailable=f'timed out after {timeout} s')
    except Exception as e:  # noqa: BLE001
        return dict(unavailable=repr(e)[:300])


def time_cpu_reference(steps, warmup, batch=1):
    """The reference's own CPU implementation (impl='ref' ops, unmodified networks.py) on the host cores; falls back to the oracle port only where the
    shipped copy of the reference tree is absent."""
    if HAVE_REF:
        # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1 to its workers: the CPU arm would otherwise run on one thread)
        ncpu = str(os.cpu_count() or 1)
        r = run_harness('cpu', batch, steps, warmup, timeout=1200, env=dict(OMP_NUM_THREADS=ncpu, MKL_NUM_THREADS=ncpu), extra_args=['--threads', ncpu])
        if 'unavailable' not in r:
            return dict(value=r['img_s'], unit=UNIT, cores=r['cores'], threads=r['threads'], kind='reference',
                        sample=f'{steps} forward(s) of batch {batch}: UNMODIFIED reference GeneratorFull 256x256 (training/networks.py:5844) with its impl=ref ops '
                               f'on torch-CPU, {r["threads"]} threads (baseline/_ref via baseline/run_reference.py --mode cpu)',
                        seconds=steps * r['ms_per_step'] * 1e-3, ms_per_step=r['ms_per_step'])
    return time_cpu_oracle(steps, warmup, batch)


# ------------------------------------------------------------------------------------------------ CPU oracle arm

def cpu_generator():
    import procedural
    from oracle import ops_oracle as O
    from pasta_gan_b200 import networks as N
    G = N.build_generator_full().eval().requires_grad_(False)
    procedural.fill_(G)
    return N.use_ops(G, O.operator_table(fast=True))


def time_cpu_oracle(steps, warmup, batch=1):
    """The oracle port of the reference impl='ref' path on the host cores: `steps` forwards of `batch` image(s)."""
    import procedural
    try:    # keep freed activation buffers in the heap instead of re-faulting fresh mmap pages on every op (glibc mallopt)
        import ctypes
        libc = ctypes.CDLL('libc.so.6')
        libc.mallopt(-3, 32 * 1024 * 1024)      # M_MMAP_THRESHOLD (max)
        libc.mallopt(-1, 2 ** 31 - 1)           # M_TRIM_THRESHOLD
    except Exception:
        pass
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    G = cpu_generator()
    inp = procedural.synth_inputs(batch)
    with torch.no_grad():
        for _ in range(warmup):
            G(**inp, noise_mode='const')
        t0 = time.perf_counter()
        for _ in range(steps):
            G(**inp, noise_mode='const')
        dt = time.perf_counter() - t0
    return dict(value=batch * steps / dt, unit=UNIT, cores=cores, threads=torch.get_num_threads(), kind='port',
                sample=f'{steps} forward(s) of batch {batch} (GeneratorFull 256x256, fp32, oracle/ops_oracle.py lowered port on torch-CPU/oneDNN)',
                seconds=dt, ms_per_step=1e3 * dt / steps)


def time_patch_routing(batch=16, reps=5):
    """the data loader's patch routing (training/dataset.py:838-927: 56 cv2.warpPerspective calls per sample) as two launches per batch
    on the GPU, beside the UNMODIFIED reference normalize() with the real cv2 on the host cores (baseline/_ref; the oracle's numpy restatement on one core
    only where cv2 or the tree is absent).  Wall clock around whole normalize() calls (host geometry + job table upload + kernels + sync), uint8 images
    resident on the device.  The GPU result is compared byte for byte with what the reference wrote."""
    import tempfile
    import numpy as np
    from pasta_gan_b200 import patch_routing as PR, synthetic
    d = synthetic.synth_patch_routing_inputs(batch, seed=3)
    dev = {k: torch.from_numpy(v).cuda() for k, v in d.items() if k != 'keypoints'}
    router = PR.PatchRouter()
    run = lambda: router.normalize(dev['upper_img'], dev['lower_img'], dev['upper_clothes_mask'], dev['lower_clothes_mask'], d['keypoints'], 2)
    run(); torch.cuda.synchronize()
    ts = []
    for _ in range(max(reps, 9)):                                         # median of per-call wall times: the host side (geometry, job table) is part of the op
        t0 = time.perf_counter()
        out = run()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    t_gpu = sorted(ts)[len(ts) // 2]
    res = dict(value=batch / t_gpu, unit='samples/s', ms_per_batch=1e3 * t_gpu, batch=batch, launches_per_batch=3,
               what='PatchRouter.normalize: 28 rectifying warps + the back-warp composite per sample, uint8, OpenCV fixed-point bilinear')
    with tempfile.TemporaryDirectory(prefix='pasta_pr_') as tmp:
        np.savez(os.path.join(tmp, 'in.npz'), **d)
        r = run_harness('patch_routing', batch, 3, 2, timeout=180, extra_args=['--io-in', os.path.join(tmp, 'in.npz'), '--io-out', os.path.join(tmp, 'out.npz')])
        if 'unavailable' not in r:
            ref = np.load(os.path.join(tmp, 'out.npz'))
            keep = ref['img'].shape[0]
            names = ('img', 'img_lower', 'denorm_upper_img', 'denorm_lower_img', 'M_invs', 'hand_masks', 'clothes_masks', 'clothes_masks_lower')
            exact = all(np.array_equal(out[i][:keep].cpu().numpy() if torch.is_tensor(out[i]) else np.asarray(out[i][:keep]), ref[names[i]].astype(np.float64) if i == 4 else ref[names[i]])
                        for i in range(8))
            res.update(bit_exact_vs_reference=bool(exact),
                       cpu_baseline=dict(value=r['samples_s'], unit='samples/s', cores=r['cv2_threads'], kind='reference',
                                         sample=f"{r['samples']} samples: UNMODIFIED reference UvitonDatasetFull.normalize (training/dataset.py:838-927) with cv2 {r['opencv']}, "
                                                f"{r['cv2_threads']} cv2 threads of {r['cores']} cores (baseline/_ref via baseline/run_reference.py --mode patch_routing)"))
            return res
    from oracle import warp_oracle as WO                                  # checker + CPU baseline only (no cv2 / no reference tree on this box)
    t0 = time.perf_counter()
    ncpu = 2
    for b in range(ncpu):
        ref = WO.normalize(d['upper_img'][b], d['lower_img'][b], d['upper_clothes_mask'][b], d['lower_clothes_mask'][b], d['keypoints'][b], 2)
    t_cpu = (time.perf_counter() - t0) / ncpu
    exact = all(np.array_equal(out[i][ncpu - 1].cpu().numpy(), ref[i]) for i in (0, 1, 2, 3, 6, 7))
    res.update(bit_exact_vs_oracle=bool(exact), reference_unavailable=r['unavailable'],
               cpu_baseline=dict(value=1.0 / t_cpu, unit='samples/s', cores=1, kind='port',
                                 sample=f'{ncpu} samples, numpy restatement of cv2.warpPerspective (oracle/warp_oracle.py, pinned to OpenCV 4.13.0 by tests/golden/warp.npz)'))
    return res


WORKLOADS = {
    'gen256': 'GeneratorFull 256x192 (256x256 padded) full-body try-on inference, batch 16 per GPU (BASELINE configs[1])',
    'gen512': 'Generator_512 512x320 (512x512 padded) try-on inference, batch 16 per GPU (BASELINE configs[2]; the only 512-px generator in the reference tree)',
}


def workload_config(args, world):
    """The `config` object of the JSON line: names the workload, identical in both arms (the reference arm runs a bounded sample of it, see its
    `cpu_baseline.sample`); what is specific to how an arm executes it is in the line's `arm` object."""
    return {'workload': WORKLOADS[args.workload], 'batch_per_gpu': args.batch, 'global_batch': world * args.batch,
            'parallelism': f'replicas x{world} (batch-sharded, no collective)',
            'l2': 'no explicit flush: one step streams > 10 GB of activations through the operators (L2 = 126 MB)',
            'weights': 'procedural (name-keyed, pasta-gan_b200/synthetic.py)', 'noise_mode': 'const'}


def time_training_step(world, rank, timeout=240, cmd=None):
    """BASELINE configs[3] -- the one path with a collective: the G + D training step (batch 4 per GPU, CUDA-graph phases, one flat NCCL all-reduce per
    phase) measured by tools/bench_train.py in a CHILD process per rank, with its own rendezvous on MASTER_PORT + 17.  Isolation on purpose: whatever
    happens in there (an exception on one rank, a stuck collective) ends with the child's timeout and cannot take the headline line down.  Every rank
    calls this at the same point; rank 0 returns the child's record, the others None."""
    env = dict(os.environ)
    if world > 1:
        env['MASTER_PORT'] = str(int(os.environ.get('MASTER_PORT', '29511')) + 17)
    for k in ('TORCHELASTIC_RUN_ID', 'TORCHELASTIC_RESTART_COUNT', 'TORCHELASTIC_MAX_RESTARTS', 'TORCHELASTIC_USE_AGENT_STORE'):
        env.pop(k, None)                                  # the child is a plain env:// rendezvous between the N children, not a worker of torchrun's agent
    cmd = cmd or [sys.executable, os.path.join(ROOT, 'tools', 'bench_train.py'), '--steps', '16', '--warmup', '17']
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=env)
        if rank != 0:
            return None
        lines = [ln for ln in r.stdout.splitlines() if ln.startswith('{')]
        if r.returncode != 0 or not lines:
            return dict(unavailable=f'rc {r.returncode}: {(r.stderr or r.stdout)[-300:]}')
        d = json.loads(lines[-1])
        return dict(value=d['value'], unit=d['unit'], ms_per_step=d['ms_per_step'], n_gpus=d['n_gpus'], batch_per_gpu=d.get('batch_per_gpu'),
                    global_batch=d.get('global_batch'), allreduce_bytes_per_step=d.get('allreduce_bytes_per_step'), cuda_graphs=d.get('cuda_graphs'),
                    library_tf32=d.get('library_tf32'), scaling='weak',
                    what='G + D training step of loss_wo_flow_fullbody (Gmain + Dmain every iteration, R1 every 16th), 16 timed steps, device-timed, max over ranks '
                         '(tools/bench_train.py in a child process per rank)')
    except subprocess.TimeoutExpired:
        return dict(unavailable=f'timed out after {timeout} s') if rank == 0 else None
    except Exception as e:  # noqa: BLE001
        return dict(unavailable=repr(e)[:300]) if rank == 0 else None


def run_reference(args, world, rank):
    if rank != 0:
        return
    r = time_cpu_reference(args.steps, args.warmup, batch=1)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': r['value'], 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': r['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args, max(1, args.gpus)),
        'arm': {'what': ('the unmodified reference (impl=ref ops)' if r['kind'] == 'reference' else 'CPU oracle port of impl=ref') + ' on the host cores of rank 0',
                'sample': 'each step = ONE image of the workload\'s batch (batch 1 is the CPU path\'s fastest shape per image: the reference\'s grouped-conv '
                          'form makes batch 2 about 5x slower per image, SURVEY 6); value = images / time',
                'batch_per_step': 1, 'l2': 'n/a (CPU)'},
        'cpu_baseline': {k: r[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
        'e2e': {'value': r['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ B200 arm

def ncu_traffic(kernel):
    """dram__bytes_read + dram__bytes_write of one `ncu --set full` launch of this kernel family (profiles/r2_kernels_ncu_full.csv, captured
    with tools/prof_ops.py at the kernel's largest generator shape); None if the capture is not in the tree."""
    import csv
    match = {'conv_igemm': 'conv_igemm_tma_persistent_kernel', 'bias_act': 'bias_act_vec_kernel', 'upfirdn2d': 'upfirdn2d_band_kernel<float, 1, 0>',
             'upfirdn2d_bias_act': 'upfirdn2d_band_kernel<float, 1, 1>', 'torgb_skip': 'torgb_skip_kernel'}.get(kernel)
    shape = {'conv_igemm': '3x3 256->128 @128^2, N=16, channel-blocked fp16 in / out, persistent TMA kernel (algorithmic 134 + 67 MB + weights)', 'bias_act': '[16,64,256,256] lrelu (algorithmic 537 MB)',
             'upfirdn2d': '[16,64,257,257]->256^2 (algorithmic 539 MB)', 'upfirdn2d_bias_act': '[16,64,257,257]->256^2 (algorithmic 539 MB)',
             'torgb_skip': '[16,64,256,256]->3 ch (algorithmic 284 MB)'}.get(kernel)
    try:
        for r in csv.DictReader(open(os.path.join(ROOT, 'profiles', 'r2_kernels_ncu_full.csv'))):
            if match and match in r['Kernel Name']:
                tot = (float(r['dram__bytes_read.sum [Mbyte]']) + float(r['dram__bytes_write.sum [Mbyte]'])) * 1e6
                return dict(traffic=tot, traffic_note=f'ncu --set full, one launch, {shape}; writes still resident in L2 at kernel end are not counted by ncu')
    except Exception:
        pass
    return dict(traffic=None)


def roofline_from(profile, pk):
    """Dominant hand-written kernel of one instrumented step -> the roofline object."""
    mine = {k: v for k, v in profile.items() if not k.startswith('library')}
    if not mine:
        return None, profile
    name = max(mine, key=lambda k: mine[k]['ms'])
    a = mine[name]
    sec = a['ms'] * 1e-3 / a['launches']
    if name.startswith('conv_igemm'):
        ach = a['flops'] / a['launches'] / sec / 1e12
        roof = dict(kernel=name, bound='tensor', achieved=ach, peak=pk['tf_sustained'], unit='TFLOP/s', frac=ach / pk['tf_sustained'],
                    peak_source=pk['src'] + ' (sustained bf16: the kernel is timed inside a long step)')
    else:
        ach = a['bytes'] / a['launches'] / sec / 1e9
        roof = dict(kernel=name, bound='hbm', achieved=ach, peak=pk['hbm'], unit='GB/s', frac=ach / pk['hbm'], peak_source=pk['src'])
    roof.update(ncu_traffic(name))
    roof['launches_per_step'] = a['launches']
    roof['avg_launch_us'] = sec * 1e6
    roof['measured'] = 'CUDA events around each launch, one instrumented eager step after the timed region'
    return roof, profile


def run_b200(args, world, rank, local):
    assert torch.cuda.is_available(), 'bench.py (impl=b200) needs a CUDA device; there is no CPU fallback'
    import procedural
    import pasta_gan_b200
    from pasta_gan_b200 import networks as N
    from pasta_gan_b200.inference import TryOnSession

    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    capi = pasta_gan_b200.capi
    pk = peaks()
    torch.backends.cudnn.benchmark = True
    if args.workload == 'gen512':
        G = N.build_generator_512().eval().requires_grad_(False)
        inp = procedural.synth_inputs_512(args.batch, seed=4321 + 100 * rank, device=dev)
    else:
        G = N.build_generator_full().eval().requires_grad_(False)
        inp = procedural.synth_inputs(args.batch, seed=1234 + 100 * rank, device=dev)
    procedural.fill_(G)
    sess = TryOnSession(G, inp, dev, use_graph=not args.no_graph, warmup=max(3, args.warmup))
    sess.synchronize()

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        sess.synchronize()
        barrier(world)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(sess.stream):
            e0.record()
        for _ in range(args.steps):
            fn()
        sess.join_streams()                                      # copies on the side streams are inside the timed region
        with torch.cuda.stream(sess.stream):
            e1.record()
        sess.synchronize()
        torch.cuda.synchronize(dev)
        barrier(world)
        return max_over_ranks(e0.elapsed_time(e1) * 1e-3, world, dev)

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    t_dev = timed(sess.step)
    clk = clocks.stop() if rank == 0 else None
    t_e2e = timed(sess.step_from_host)
    # the same call with the loader's uint8 tensors in and the uint8 BGR photo out (device-side normalise / concat / crop / BGR kernels, test.py:105-135)
    if args.workload == 'gen512':
        u8 = procedural.synth_inputs_u8(args.batch, res=512, parts_ch=48, parts_res=128, seed=4321 + 100 * rank, full_body=False)
        sess.enable_u8_io(u8, crop=(96, 416))
    else:
        u8 = procedural.synth_inputs_u8(args.batch, seed=1234 + 100 * rank)
        sess.enable_u8_io(u8)
    t_u8 = timed(sess.step_from_host_u8)

    # one instrumented eager step (per-launch CUDA events) for the roofline of the dominant hand-written kernel
    with torch.cuda.stream(sess.stream), torch.no_grad():
        l0 = capi.launch_count()
        sess.G(**sess.static_in, noise_mode='const')            # eager warm-up on the session stream (allocator pools, cuDNN plans)
        per_fwd = capi.launch_count() - l0                      # launches of OUR kernels in one forward (what each graph replay re-issues)
        # per-launch CUDA events only measure the kernels if the GPU never waits for the host: park ~15 ms of spin in front of the step so that the host
        # has queued every launch before the first one executes (otherwise a 5 us kernel shows up as the ~50 us it takes Python to issue the next call)
        torch.cuda._sleep(int(3e7))
        with capi.LaunchProfiler() as prof:
            sess.G(**sess.static_in, noise_mode='const')
    roof, profile = roofline_from(prof.summary(), pk)
    if roof is not None:
        # the dominant kernel family by layer shape: algorithmic TFLOP/s and GB/s per shape (the family mixes tensor-bound 128^2 layers with
        # HBM-bound 64-channel 256^2 layers and latency-bound 4^2..16^2 layers, so the family-wide fraction understates the big shapes)
        by = {}
        for name, nbytes, flops, e0, e1, tag in prof.records:
            if name == roof['kernel']:
                a = by.setdefault(tag, [0, 0.0, 0, 0])
                a[0] += 1; a[1] += e0.elapsed_time(e1); a[2] += nbytes; a[3] += flops
        roof['by_shape'] = [dict(shape=t, launches=n, ms=round(ms, 3), tflops=round(fl / ms / 1e9, 1), gbs=round(nb / ms / 1e6),
                                 frac_tensor=round(fl / ms / 1e9 / pk['tf_sustained'], 3), frac_hbm=round(nb / ms / 1e6 / pk['hbm'], 3))
                            for t, (n, ms, nb, fl) in sorted(by.items(), key=lambda kv: -kv[1][1])[:8] if ms > 0]

    # library baseline: the same module tree with every convolution on cuDNN fp32 (what the reference's conv2d_gradfix calls; allow_tf32 = False as in
    # training_loop_wo_flow_fullbody.py:243,253), FIR / bias_act on our kernels; eager, same batch, CUDA events
    lib_base = None
    extras = world == 1 and not args.no_extras
    if extras:
        from pasta_gan_b200.torch_utils.ops import conv_igemm as K
        old = (K.enabled, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        K.enabled, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = False, False, False
        try:
            with torch.cuda.stream(sess.stream), torch.no_grad():
                for _ in range(2):
                    sess.G(**sess.static_in, noise_mode='const')
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(3):
                    sess.G(**sess.static_in, noise_mode='const')
                e1.record()
            sess.synchronize()
            ms = e0.elapsed_time(e1) / 3
            lib_base = dict(value=args.batch / (ms * 1e-3), unit=UNIT, ms_per_step=ms,
                            what='same module tree, convolutions on cuDNN fp32 (conv_igemm disabled, allow_tf32=False), eager, 3 steps')
        except Exception as e:  # noqa: BLE001 -- a side leg must never take the headline measurement down with it
            lib_base = dict(unavailable=repr(e)[:300])
        finally:
            K.enabled, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old

    # configs[3] (training step, the path with the NCCL exchange) at this N, in child processes: every rank takes part
    train_leg = None
    if not args.no_extras and args.workload == 'gen256':
        barrier(world)
        train_leg = time_training_step(world, rank)
        barrier(world)
    if rank != 0:
        return
    imgs = world * args.batch * args.steps
    cpu = None
    if world == 1 and not args.skip_cpu_baseline:
        cpu = time_cpu_reference(steps=20, warmup=2, batch=1)
        cpu = {k: cpu[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
    extra = {}
    if train_leg is not None:
        extra['training'] = train_leg
    io_bytes = dict(h2d=sess.h2d_bytes, d2h=sess.d2h_bytes, h2d_u8=sess.h2d_bytes_u8, d2h_u8=sess.d2h_bytes_u8)
    if extras:
        try:
            t_x = time.perf_counter()
            del sess                                                 # free the session's buffers for the reference-tree processes below
            torch.cuda.empty_cache()
            # the real drop-in: unmodified reference networks.py over OUR ops on this GPU (eager; every modulated conv arrives as groups = N)
            d = run_harness('overlay', args.batch, 5, 2, timeout=300)
            extra['dropin'] = d if 'unavailable' in d else dict(value=d['img_s'], unit=UNIT, ms_per_step=d['ms_per_step'], our_kernel_launches=d['our_kernel_launches'],
                                                                  **{k: d[k] for k in ('graph_img_s', 'graph_ms_per_step', 'graph_unavailable') if k in d},
                                                                  what='UNMODIFIED reference training/networks.py GeneratorFull over our torch_utils/ops overlay, eager, batch %d' % args.batch)
            if time.perf_counter() - t_x < 120:
                # the same after a legacy.load_network_pkl round trip with the persistence.import_hook recipe of INTEGRATION.md 2 (the reference's own
                # hook re-routes modulated_conv2d of the pickled source to our one-launch layer; everything else stays the reference's code)
                d = run_harness('overlay_hook', args.batch, 5, 2, timeout=300)
                extra['dropin_hook'] = d if 'unavailable' in d else dict(value=d['img_s'], unit=UNIT, ms_per_step=d['ms_per_step'], our_kernel_launches=d['our_kernel_launches'],
                                                                           **{k: d[k] for k in ('graph_img_s', 'graph_ms_per_step', 'graph_unavailable') if k in d},
                                                                           what='UNMODIFIED reference networks, pickled and re-loaded through legacy.load_network_pkl with the import_hook that swaps modulated_conv2d, eager, batch %d' % args.batch)
            if time.perf_counter() - t_x < 240:
                d = run_harness('gpu', args.batch, 5, 2, timeout=420)
                extra['reference_gpu'] = d if 'unavailable' in d else dict(value=d['img_s'], unit=UNIT, ms_per_step=d['ms_per_step'], plugins=d.get('plugins'),
                                                                             what='UNMODIFIED reference with its own CUDA plugins (JIT, sm_100) + cuDNN fp32, eager, batch %d' % args.batch)
            if args.workload == 'gen256' and time.perf_counter() - t_x < 480:
                try:
                    r = subprocess.run([sys.executable, os.path.abspath(__file__), '--workload', 'gen512', '--steps', str(args.steps), '--warmup', str(args.warmup),
                                        '--batch', str(args.batch), '--skip-cpu-baseline', '--no-extras'], capture_output=True, text=True, timeout=300, cwd=ROOT)
                    g = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith('{')][-1])
                    extra['gen512'] = {k: g[k] for k in ('metric', 'value', 'unit', 'ms_per_step', 'e2e', 'e2e_u8', 'roofline', 'gpu_launches_per_step', 'config')}
                except Exception as e:  # noqa: BLE001
                    extra['gen512'] = dict(unavailable=repr(e)[:300])
            try:
                extra['patch_routing'] = time_patch_routing(args.batch)
            except Exception as e:  # noqa: BLE001
                extra['patch_routing'] = dict(unavailable=repr(e)[:300])
        except Exception as e:  # noqa: BLE001 -- side legs never take the headline line down
            extra['error'] = repr(e)[:300]
    act_bytes = sum(v['bytes'] for v in profile.values())
    line = {
        'metric': METRIC if args.workload == 'gen256' else METRIC.replace('256x192 padded to 256x256', '512x320 padded to 512x512'), 'value': imgs / t_dev, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * t_dev / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': DTYPE, 'data': 'synthetic',
        'config': workload_config(args, world),
        'arm': {'cuda_graph': not args.no_graph, 'e2e_pipeline': 'H2D / compute / D2H on three streams, 2 buffer sets',
                'activation_bytes_per_step': act_bytes, 'l2': f'one step streams ~{act_bytes / 1e9:.1f} GB of activations through the operators (> 126 MB L2)'},
        'e2e': {'value': imgs / t_e2e, 'unit': UNIT, 'ms_per_step': 1e3 * t_e2e / args.steps,
                'h2d_bytes_per_step': io_bytes['h2d'], 'd2h_bytes_per_step': io_bytes['d2h']},
        'e2e_u8': {'value': imgs / t_u8, 'unit': UNIT, 'ms_per_step': 1e3 * t_u8 / args.steps, 'h2d_bytes_per_step': io_bytes['h2d_u8'],
                   'd2h_bytes_per_step': io_bytes['d2h_u8'], 'note': 'uint8 loader tensors in, uint8 BGR photo out (TryOnSession.step_from_host_u8)'},
        'gpu_launches': per_fwd * args.steps,
        'gpu_launches_per_step': per_fwd,
        'clocks': clk,
        'roofline': roof,
        'cpu_baseline': cpu,
        'gpu_library_baseline': lib_base,
        'extra': extra,
        'kernel_breakdown_ms_per_step': {k: round(v['ms'], 3) for k, v in sorted(profile.items(), key=lambda kv: -kv[1]['ms'])},
        'kernel_launches_per_step': {k: v['launches'] for k, v in profile.items()},
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    world, rank, local = dist_setup(args)
    try:
        if args.impl == 'reference':
            run_reference(args, world, rank)
        else:
            run_b200(args, world, rank, local)
    finally:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()


if __name__ == '__main__':
    main()
