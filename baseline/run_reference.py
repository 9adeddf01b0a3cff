#!/usr/bin/env python
"""Runs the UNMODIFIED reference tree in this process, in one of several arrangements, and prints ONE JSON line.

The tree is ``baseline/_ref`` — a plain file copy of the reference checkout made by ``__graft_entry__.build()`` in the authoring container
(git-ignored, so never part of the history; it travels to the GPU box with the gpurun snapshot).  Nothing of it is modified: the harness only
supplies what the reference expects from its environment: a stub ``matplotlib``, cwd = tree root for ``./human_colormap.mat``,
and ``torch.version.cuda == '11.0'`` while ``training.networks`` is imported (its lines 1206-1222 otherwise try to JIT a second upfirdn2d op from
files that do not exist).

    --mode cpu       reference networks + reference ops, CPU tensors => every op takes impl='ref' (upfirdn2d.py:162-164, bias_act.py:87-89).
                     This is the reference's own CPU implementation of the hot path: bench.py's --impl reference arm and cpu_baseline.
    --mode gpu       reference networks + reference ops on the GPU: its own CUDA plugins (upfirdn2d.cu / bias_act.cu JIT-built for sm_100 through
                     torch_utils/custom_ops.py) and cuDNN fp32 convolutions (allow_tf32 = False, training_loop...py:243,253).  The GPU code to beat.
    --mode overlay   reference networks (unmodified training/networks.py: GeneratorFull :5844) over OUR torch_utils/ops/*.py — the drop-in of
                     INTEGRATION.md §2, built as an overlay copy in a temp dir.  Eval mode, so modulated convs arrive as groups = N.
    --mode overlay_hook   the same after a pickle round trip through legacy.load_network_pkl with the persistence.import_hook recipe of
                     INTEGRATION.md §2 that re-routes the pickled modulated_conv2d to the one-launch kernel.
    --mode ops       per-op timings of the reference's CUDA plugins and cuDNN fp32 on the op-microbenchmark grid (BASELINE configs[4]).
    --mode ops_cpu   the same grid through the reference's impl='ref' path on CPU tensors (host cores): the "vs impl='ref'" column of BASELINE configs[4].
    --mode patch_routing   the reference's own data-loader patch routing on the host: UvitonDatasetFull.normalize (training/dataset.py:838-927, 56
                     cv2.warpPerspective calls per sample) on the inputs in --io-in (.npz), outputs of the first samples written to --io-out.

Parity (``--check``): outputs at batch 2 against tests/golden/generator_full.npz (written from this same reference on CPU).
"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('PASTA_REFERENCE_TREE', os.path.join(ROOT, 'baseline', '_ref'))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--mode', required=True, choices=['cpu', 'gpu', 'overlay', 'overlay_hook', 'ops', 'ops_cpu', 'patch_routing'])
    ap.add_argument('--profile', action='store_true', help='print the top CUDA kernels of one forward (torch.profiler) to stderr')
    ap.add_argument('--batch', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=1)
    ap.add_argument('--check', action='store_true', help='also run batch 2 and compare with tests/golden/generator_full.npz')
    ap.add_argument('--threads', type=int, default=0)
    ap.add_argument('--io-in', default='', help='patch_routing: .npz with upper_img / lower_img / upper_clothes_mask / lower_clothes_mask / keypoints')
    ap.add_argument('--io-out', default='', help='patch_routing: where to write the reference outputs (.npz)')
    return ap.parse_args()


def make_overlay_tree():
    """Copy of the reference tree with our torch_utils/ops/*.py laid over its own (exactly the `cp` of INTEGRATION.md §2)."""
    tree = os.path.join(tempfile.mkdtemp(prefix='pasta_overlay_'), 'ref')
    shutil.copytree(REF, tree, ignore=shutil.ignore_patterns('__pycache__', '*.pyc', '.git'))
    ops_src = os.path.join(ROOT, 'pasta-gan_b200', 'torch_utils', 'ops')
    for fn in os.listdir(ops_src):
        if fn.endswith('.py') and fn != '__init__.py':
            shutil.copy(os.path.join(ops_src, fn), os.path.join(tree, 'torch_utils', 'ops', fn))
    os.environ['PASTA_B200_HOME'] = os.path.join(ROOT, 'pasta-gan_b200')
    return tree


def import_reference(tree):
    os.environ.setdefault('PYTHONDONTWRITEBYTECODE', '1')
    sys.dont_write_bytecode = True
    sys.path.insert(0, tree)
    sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
    for m in ('matplotlib', 'matplotlib.pyplot'):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    import torch
    os.chdir(tree)
    real = torch.version.cuda
    torch.version.cuda = '11.0'
    try:
        import training.networks as R_net
        import legacy
    finally:
        torch.version.cuda = real
    return R_net, legacy


def build_generator(R_net):
    return R_net.GeneratorFull(z_dim=0, c_dim=512, w_dim=512, img_resolution=256, img_channels=3, mapping_kwargs=dict(num_layers=1),
                               synthesis_kwargs=dict(channel_base=16384, channel_max=512, num_fp16_res=3, conv_clamp=256, use_noise=True)).eval().requires_grad_(False)


def golden_check(G, device):
    """batch-2 outputs against the committed fixture: max-abs relative error of the coarse image and parsing logits, relative L2 of all three."""
    import numpy as np
    import torch
    import procedural
    z = np.load(os.path.join(ROOT, 'tests', 'golden', 'generator_full.npz'))
    inp = procedural.synth_inputs(2, device=device)
    with torch.no_grad():
        img, fimg, parsing = G(**inp, noise_mode='const')
    out = dict(img=img, finetune_img=fimg, pred_parsing=parsing)
    res = {}
    for k, v in out.items():
        ref = torch.from_numpy(z[k].astype(np.float32)).double()
        got = v.detach().cpu().double()
        res[k] = dict(max_rel=float((got - ref).abs().max() / ref.abs().max()), l2_rel=float((got - ref).norm() / ref.norm()))
    return res


def time_generator(G, batch, steps, warmup, device):
    import torch
    import procedural
    inp = procedural.synth_inputs(batch, seed=1234, device=device)
    cuda = torch.device(device).type == 'cuda'
    with torch.no_grad():
        for _ in range(warmup):
            G(**inp, noise_mode='const')
        if cuda:
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        t0 = time.perf_counter()
        for _ in range(steps):
            G(**inp, noise_mode='const')
        if cuda:
            e1.record()
            torch.cuda.synchronize()
            dt = e0.elapsed_time(e1) * 1e-3
        else:
            dt = time.perf_counter() - t0
    return dict(img_s=batch * steps / dt, ms_per_step=1e3 * dt / steps, batch=batch, steps=steps, warmup=warmup)


def time_generator_graph(G, batch, steps, device):
    """The same forward captured once into a CUDA graph and replayed: what is left of the drop-in's time when the host's per-op overhead
    (1 960 launches per step from the unmodified reference code) is taken out.  Returns {} when the reference code cannot be captured."""
    import torch
    import procedural
    inp = procedural.synth_inputs(batch, seed=1234, device=device)
    try:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        g = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.stream(s):
            G(**inp, noise_mode='const')
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s):
                out = G(**inp, noise_mode='const')
        torch.cuda.synchronize()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        dt = e0.elapsed_time(e1) * 1e-3
        return dict(graph_img_s=batch * steps / dt, graph_ms_per_step=1e3 * dt / steps)
    except Exception as e:  # noqa: BLE001
        torch.cuda.synchronize()
        return dict(graph_unavailable=repr(e)[:200])


def run_ops(out):
    """Reference CUDA plugins + cuDNN fp32 on the op grid of BASELINE configs[4] (N = 16): microseconds per call, CUDA events, inputs cycled through a
    pool larger than L2."""
    import torch
    from torch_utils.ops import upfirdn2d, bias_act
    dev = 'cuda'
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    f = upfirdn2d.setup_filter([1, 3, 3, 1], device=dev)
    rows = []

    def bench(fn, make, nbytes, reps=10):
        pool = [make() for _ in range(max(2, int(400e6 // max(nbytes, 1)) + 1))][:8]
        for p in pool[:2]:
            fn(*p)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            fn(*pool[i % len(pool)])
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / reps

    n = 16
    for res, c in [(512, 32), (256, 64), (128, 128), (64, 256), (32, 512)]:
        numel = n * c * res * res
        with torch.no_grad():
            us = bench(lambda x, b: bias_act.bias_act(x, b, act='lrelu', clamp=256), lambda: (torch.randn(n, c, res, res, device=dev), torch.randn(c, device=dev)), 8 * numel)
            rows.append(dict(op='bias_act lrelu+clamp', res=res, c=c, us=us, gbs=8 * numel / us / 1e3))
            us = bench(lambda x: upfirdn2d.upfirdn2d(x, f, padding=[1, 1, 1, 1], gain=4), lambda: (torch.randn(n, c, res + 1, res + 1, device=dev),), 8 * numel)
            rows.append(dict(op='upfirdn2d filter pad1 gain4', res=res, c=c, us=us, gbs=4 * (numel + n * c * (res + 1) ** 2) / us / 1e3))
            us = bench(lambda x: upfirdn2d.upfirdn2d(x, f, down=2, padding=[1, 1, 1, 1]), lambda: (torch.randn(n, c, res, res, device=dev),), 5 * numel)
            rows.append(dict(op='upfirdn2d down2', res=res, c=c, us=us, gbs=5 * numel / us / 1e3))
            if res <= 256:
                us = bench(lambda x: upfirdn2d.upsample2d(x, f), lambda: (torch.randn(n, c, res // 2, res // 2, device=dev),), 5 * numel)
                rows.append(dict(op='upfirdn2d up2 (upsample2d)', res=res, c=c, us=us, gbs=5 * numel / us / 1e3))
            us = bench(lambda x: upfirdn2d.upsample2d(x, f), lambda: (torch.randn(n, 3, res // 2, res // 2, device=dev),), 5 * n * 3 * res * res)
            rows.append(dict(op='upfirdn2d up2 rgb', res=res, c=3, us=us, gbs=5 * n * 3 * res * res / us / 1e3))
            w = torch.randn(c, c, 3, 3, device=dev) / (c * 9) ** 0.5
            us = bench(lambda x: torch.nn.functional.conv2d(x, w, padding=1), lambda: (torch.randn(n, c, res, res, device=dev),), 8 * numel, reps=6)
            rows.append(dict(op='cudnn fp32 conv3x3', res=res, c=c, us=us, tflops=2 * numel * c * 9 / us / 1e6))
    out['ops'] = rows
    out['plugins'] = dict(upfirdn2d=upfirdn2d._plugin is not None, bias_act=bias_act._plugin is not None)


def run_ops_cpu(out, R_net, budget_s=150.0):
    """The op grid of run_ops through the reference's impl='ref' code on CPU tensors (upfirdn2d.py:169-208, bias_act.py:94-123, networks.py:37-94 with
    fused_modconv=True), all host threads; one untimed call then the best of two timed calls per cell; cells are skipped once the time budget is spent."""
    import torch
    from torch_utils.ops import upfirdn2d, bias_act
    f = upfirdn2d.setup_filter([1, 3, 3, 1])
    rows, t_start = [], time.perf_counter()

    def bench(fn, *a):
        if time.perf_counter() - t_start > budget_s:
            return None
        with torch.no_grad():
            fn(*a)
            ts = []
            for _ in range(2):
                t0 = time.perf_counter()
                fn(*a)
                ts.append(time.perf_counter() - t0)
        return min(ts) * 1e6

    n = 16
    for res, c in [(32, 512), (64, 256), (128, 128), (256, 64), (512, 32)]:
        numel = n * c * res * res
        cells = [
            ('bias_act lrelu+clamp', lambda x, b: bias_act.bias_act(x, b, act='lrelu', clamp=256), (torch.randn(n, c, res, res), torch.randn(c)), 8 * numel, None),
            ('upfirdn2d filter pad1 gain4', lambda x: upfirdn2d.upfirdn2d(x, f, padding=[1, 1, 1, 1], gain=4), (torch.randn(n, c, res + 1, res + 1),),
             4 * (numel + n * c * (res + 1) ** 2), None),
            ('upfirdn2d down2', lambda x: upfirdn2d.upfirdn2d(x, f, down=2, padding=[1, 1, 1, 1]), (torch.randn(n, c, res, res),), 5 * numel, None),
            ('upfirdn2d up2 rgb', lambda x: upfirdn2d.upsample2d(x, f), (torch.randn(n, 3, res // 2, res // 2),), 5 * n * 3 * res * res, None),
        ]
        if res <= 256:
            cells.append(('upfirdn2d up2 (upsample2d)', lambda x: upfirdn2d.upsample2d(x, f), (torch.randn(n, c, res // 2, res // 2),), 5 * numel, None))
            w = torch.randn(c, c, 3, 3)
            st = torch.randn(n, c)
            cells.append(('modulated_conv2d 3x3 fused', lambda x: R_net.modulated_conv2d(x=x, weight=w, styles=st, padding=1, fused_modconv=True),
                          (torch.randn(n, c, res, res),), 8 * numel, 2 * numel * c * 9))
        for name, fn, args, nbytes, flops in cells:
            us = bench(fn, *args)
            if us is None:
                rows.append(dict(op=name, res=res, c=c, skipped='time budget'))
                continue
            r = dict(op=name, res=res, c=3 if name.endswith('rgb') else c, us=us, gbs=nbytes / us / 1e3)
            if flops:
                r['tflops'] = flops / us / 1e6
            rows.append(r)
    out['ops'] = rows
    out['cores'] = os.cpu_count() or 1
    out['threads'] = torch.get_num_threads()
    out['what'] = "UNMODIFIED reference ops, impl='ref' (CPU tensors), N = 16, best of 2 after one warm-up call"


NORM_NAMES = ('img', 'img_lower', 'denorm_upper_img', 'denorm_lower_img', 'M_invs', 'hand_masks', 'clothes_masks', 'clothes_masks_lower')


def run_patch_routing(args, out):
    """The unmodified reference methods on the host cores (what a DataLoader worker executes per sample, dataset.py:553-565): needs cv2.  The modules the
    data-set file imports at its top and that this image lacks (skimage.draw, pycocotools.mask, matplotlib) are irrelevant to the three methods and are
    satisfied by empty stand-ins; the methods run unbound on an object carrying the sample's ``keypoints`` (set by __getitem__ at dataset.py:744)."""
    try:
        import cv2
    except ImportError as e:
        out['unavailable'] = f'cv2 cannot be imported here ({e})'
        return
    import numpy as np
    os.environ.setdefault('PYTHONDONTWRITEBYTECODE', '1')
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    for m in ('matplotlib', 'matplotlib.pyplot', 'skimage', 'skimage.draw', 'pycocotools', 'pycocotools.mask'):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    sys.modules['skimage'].draw = sys.modules['skimage.draw']
    if not hasattr(sys.modules['skimage.draw'], 'circle'):
        sys.modules['skimage.draw'].circle = sys.modules['skimage.draw'].line_aa = None
    sys.modules['pycocotools'].mask = sys.modules['pycocotools.mask']
    import training.dataset as R_ds

    class Sample:
        valid_joints = R_ds.UvitonDatasetFull.valid_joints
        get_crop = R_ds.UvitonDatasetFull.get_crop
        normalize = R_ds.UvitonDatasetFull.normalize

        def __init__(self, keypoints):
            self.keypoints = keypoints

    d = dict(np.load(args.io_in))
    B = d['keypoints'].shape[0]
    if args.threads:
        cv2.setNumThreads(args.threads)
    run = lambda b: Sample(d['keypoints'][b]).normalize(d['upper_img'][b], d['lower_img'][b], d['upper_clothes_mask'][b], d['lower_clothes_mask'][b], 2)
    for i in range(args.warmup):
        run(i % B)
    t0 = time.perf_counter()
    n = 0
    for _ in range(args.steps):
        for b in range(B):
            res = run(b)
            n += 1
    dt = time.perf_counter() - t0
    out.update(samples_s=n / dt, ms_per_sample=1e3 * dt / n, samples=n, batch=B, opencv=cv2.__version__, cv2_threads=cv2.getNumThreads(), cores=os.cpu_count() or 1)
    if args.io_out:
        keep = min(B, 4)
        cols = {k: [] for k in NORM_NAMES}
        for b in range(keep):
            for k, v in zip(NORM_NAMES, run(b)):
                cols[k].append(np.stack(v) if k == 'hand_masks' else np.asarray(v))
        np.savez(args.io_out, **{k: np.stack(v) for k, v in cols.items()})


def main():
    args = parse()
    out = dict(mode=args.mode, ref_tree=os.path.relpath(REF, ROOT))
    if not os.path.isdir(os.path.join(REF, 'torch_utils')):
        out['unavailable'] = f'{REF} is missing (it is copied from the reference checkout by __graft_entry__.build())'
        print(json.dumps(out), flush=True)
        return
    if args.mode == 'patch_routing':
        run_patch_routing(args, out)
        print(json.dumps(out), flush=True)
        return
    import torch
    if args.threads:
        torch.set_num_threads(args.threads)
    overlay = args.mode.startswith('overlay')
    tree = make_overlay_tree() if overlay else REF
    if args.mode in ('gpu', 'ops'):
        os.environ.setdefault('TORCH_CUDA_ARCH_LIST', '10.0')
        os.environ.setdefault('TORCH_EXTENSIONS_DIR', os.path.join(tempfile.gettempdir(), 'pasta_ref_plugins'))
        # custom_ops.get_plugin builds with torch.utils.cpp_extension.load and then does importlib.import_module(name) (custom_ops.py:107-111); current
        # torch no longer leaves the built module importable by name, so put the build directories on sys.path (environment, not a source change)
        import torch.utils.cpp_extension as _ce
        for _name in ('bias_act_plugin', 'upfirdn2d_plugin'):
            sys.path.insert(0, _ce._get_build_directory(_name, verbose=False))
    R_net, legacy = import_reference(tree)
    import procedural
    device = 'cpu' if args.mode in ('cpu', 'ops_cpu') else 'cuda'
    if device == 'cuda':
        torch.backends.cudnn.benchmark = True                  # training_loop_wo_flow_fullbody.py:242
        torch.backends.cudnn.allow_tf32 = False                # :243, :253 — the reference's own GPU convolutions are full fp32
        torch.backends.cuda.matmul.allow_tf32 = False
    if args.mode == 'ops':
        run_ops(out)
        print(json.dumps(out), flush=True)
        return
    if args.mode == 'ops_cpu':
        run_ops_cpu(out, R_net)
        print(json.dumps(out), flush=True)
        return
    if args.mode == 'cpu':
        try:    # keep freed activation buffers in the heap instead of re-faulting fresh mmap pages on every op (glibc mallopt)
            import ctypes
            libc = ctypes.CDLL('libc.so.6')
            libc.mallopt(-3, 32 * 1024 * 1024)
            libc.mallopt(-1, 2 ** 31 - 1)
        except Exception:
            pass
        out['cores'] = os.cpu_count() or 1
        out['threads'] = torch.get_num_threads()
    G = build_generator(R_net)
    procedural.fill_(G)
    if args.mode == 'overlay_hook':
        # pickle with the reference's persistence, re-load through legacy.load_network_pkl with an import hook that swaps the
        # pickled module's modulated_conv2d for the one-launch kernel entry point
        import io
        import pickle
        from torch_utils import persistence
        sys.path.insert(0, ROOT)

        @persistence.import_hook
        def _use_b200_modconv(meta):
            if meta.type == 'class' and 'def modulated_conv2d(' in meta.module_src:
                meta.module_src = meta.module_src.replace('def modulated_conv2d(', 'def _modulated_conv2d_ref(') + \
                    '\nfrom pasta_gan_b200.networks import modulated_conv2d\n'
            return meta
        buf = io.BytesIO()
        D_small = R_net.Discriminator(c_dim=512, img_resolution=256, img_channels=3, channel_base=1024, channel_max=32, num_fp16_res=3, conv_clamp=256)
        pickle.dump(dict(G=G, D=D_small, G_ema=G), buf)       # the snapshot layout of training_loop_wo_flow_fullbody.py:587-602
        buf.seek(0)
        real = torch.version.cuda
        torch.version.cuda = '11.0'                      # the pickled module source is exec'd again on load (persistence.py:216-227)
        try:
            G = legacy.load_network_pkl(buf)['G_ema'].eval().requires_grad_(False)
        finally:
            torch.version.cuda = real
        out['hooked'] = 'pasta_gan_b200' in sys.modules
    G = G.to(device)
    if overlay:
        from torch_utils.ops import upfirdn2d
        out['overlay_in_effect'] = 'pg_upfirdn2d' in open(upfirdn2d.__file__).read()
    if args.check:
        out['golden'] = golden_check(G, device)
    if device == 'cuda':
        capi = sys.modules.get('pasta_b200_capi')              # our C-ABI binding, if the overlay loaded it into this process
        l0 = capi.launch_count() if capi is not None else 0
    out.update(time_generator(G, args.batch, args.steps, args.warmup, device))
    if args.profile and device == 'cuda':
        import procedural
        from torch.profiler import profile, ProfilerActivity
        inp = procedural.synth_inputs(args.batch, seed=1234, device=device)
        with torch.no_grad(), profile(activities=[ProfilerActivity.CUDA]) as prof:
            G(**inp, noise_mode='const')
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=30, max_name_column_width=90), file=sys.stderr)
    if device == 'cuda':
        out['our_kernel_launches'] = (capi.launch_count() - l0) if capi is not None else 0
        if overlay and not args.check:
            out.update(time_generator_graph(G, args.batch, args.steps, device))
        if args.mode == 'gpu':
            from torch_utils.ops import upfirdn2d, bias_act
            out['plugins'] = dict(upfirdn2d=upfirdn2d._plugin is not None, bias_act=bias_act._plugin is not None)
    print(json.dumps(out), flush=True)


if __name__ == '__main__':
    main()
