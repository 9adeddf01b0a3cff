#!/usr/bin/env python
"""Op microbenchmark sweep (BASELINE.json configs[4]): upfirdn2d forms + bias_act + fused, at (res, C) of the
generator, N in {1, 16}.  Prints one JSON line per case: algorithmic bytes / CUDA-event time vs the measured HBM peak.
Inputs are rotated through a pool larger than L2 (126 MB) so every timed launch reads from HBM."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pasta_gan_b200.torch_utils.ops import upfirdn2d, bias_act  # noqa: E402


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'], 'measured'
    except Exception:
        return 6650.0, 'fallback'


def timeit(fn, inputs, iters):
    """inputs: list of argument tuples rotated per call (pool > L2).  The `iters` launches are captured into ONE CUDA graph and the
    replay is timed, so that short kernels are measured on the device and not by the host's per-call overhead (~40 us through Python)."""
    for k in range(3):
        fn(*inputs[k % len(inputs)])
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for k in range(iters):
                fn(*inputs[k % len(inputs)])
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / iters)
    ts.sort()
    return ts[len(ts) // 2] * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--batches', type=int, nargs='+', default=[16, 1])
    ap.add_argument('--out', default=None)
    a = ap.parse_args()
    dev = torch.device('cuda:0')
    hbm, src = peaks()
    f = upfirdn2d.setup_filter([1, 3, 3, 1]).to(dev)
    sweep = [(8, 512), (16, 512), (32, 512), (64, 256), (128, 128), (256, 64), (512, 32)]
    lines = []

    def pool(shape, min_bytes=400e6):
        n = int(max(2, min(8, min_bytes // (4 * torch.Size(shape).numel()) + 1)))
        return [torch.randn(*shape, device=dev) for _ in range(n)]

    for N in a.batches:
        for res, C in sweep:
            cases = []
            xs = pool([N, C, res + 1, res + 1])
            cases.append(('filter_pad1_g4', lambda x: upfirdn2d.upfirdn2d(x, f, padding=[1, 1, 1, 1], gain=4), [(x,) for x in xs],
                          4 * N * C * ((res + 1) ** 2 + res ** 2)))
            b = torch.randn(C, device=dev)
            cases.append(('filter_pad1_g4+bias_lrelu(fused)', lambda x: upfirdn2d.upfirdn2d_bias_act(x, f, b, padding=[1, 1, 1, 1], gain=4, act='lrelu', clamp=256),
                          [(x,) for x in xs], 4 * N * C * ((res + 1) ** 2 + res ** 2)))
            xs2 = pool([N, C, res, res])
            cases.append(('filter_pad2', lambda x: upfirdn2d.upfirdn2d(x, f, padding=[2, 2, 2, 2]), [(x,) for x in xs2],
                          4 * N * C * (res ** 2 + (res + 1) ** 2)))
            cases.append(('down2', lambda x: upfirdn2d.downsample2d(x, f), [(x,) for x in xs2], 4 * N * C * (res ** 2 + (res // 2) ** 2)))
            cases.append(('bias_act_lrelu_clamp', lambda x: bias_act.bias_act(x, b, act='lrelu', clamp=256), [(x,) for x in xs2], 8 * N * C * res * res))
            xs4 = pool([N, C, res // 2, res // 2])                  # full-channel up-2: the backward of every down-2 (upfirdn2d.py:251-261)
            cases.append(('up2_full_channels', lambda x: upfirdn2d.upsample2d(x, f), [(x,) for x in xs4], 4 * N * C * ((res // 2) ** 2 + res ** 2)))
            xs3 = pool([N, 3, res // 2, res // 2])
            cases.append(('up2_rgb', lambda x: upfirdn2d.upsample2d(x, f), [(x,) for x in xs3], 4 * N * 3 * ((res // 2) ** 2 + res ** 2)))
            for name, fn, inputs, nbytes in cases:
                with torch.no_grad():
                    t = timeit(fn, inputs, a.iters)
                line = dict(op=name, N=N, C=C, res=res, us=round(t * 1e6, 2), GBps=round(nbytes / t / 1e9, 1),
                            frac_of_hbm=round(nbytes / t / 1e9 / hbm, 3), peak=hbm, peak_src=src, bytes=nbytes)
                lines.append(line)
                print(json.dumps(line), flush=True)
            del xs, xs2, xs3, xs4
            torch.cuda.empty_cache()
    if a.out:
        with open(a.out, 'w') as fh:
            for ln in lines:
                fh.write(json.dumps(ln) + '\n')


if __name__ == '__main__':
    main()
