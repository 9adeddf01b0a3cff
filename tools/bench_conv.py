#!/usr/bin/env python
"""Convolution microbenchmark: tcgen05 implicit-GEMM (ours) vs the library convolution the reference calls (cuDNN, TF32 on
and off), at the generator's layer shapes, N = 16.  FLOPs per SURVEY.md §8(d): 2*N*Cout*Cin*k*k*Hout*Wout (stride 1) and
2*N*Cout*Cin*9*Hin*Win for up-2 (zero taps excluded)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pasta_gan_b200.torch_utils.ops import conv_igemm, conv2d_resample, upfirdn2d  # noqa: E402


def timeit(fn, iters=10):
    """Median of 3 runs of `iters` back-to-back launches between two CUDA events.  A spin kernel is parked in front so that the host has queued all
    launches before the first executes: short kernels are then timed on the device, not by the ~50 us Python needs to issue each call."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        torch.cuda._sleep(int(4e6))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / iters)
    ts.sort()
    return ts[1] * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=None)
    ap.add_argument('--n', type=int, default=16)
    a = ap.parse_args()
    dev = torch.device('cuda:0')
    N = a.n
    peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['bf16_tflops'] if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 1590.0
    f = upfirdn2d.setup_filter([1, 3, 3, 1]).to(dev)
    shapes = [  # (Cin, Cout, res_in, k, up)
        (512, 512, 4, 3, 1), (512, 512, 8, 3, 1), (512, 512, 16, 3, 1), (512, 512, 32, 3, 1), (256, 256, 64, 3, 1), (128, 128, 128, 3, 1),
        (256, 128, 128, 3, 1), (64, 64, 256, 3, 1), (128, 64, 256, 3, 1), (192, 128, 128, 1, 1), (128, 64, 256, 1, 1), (64, 3, 256, 1, 1),
        (512, 512, 16, 3, 2), (512, 256, 32, 3, 2), (256, 128, 64, 3, 2), (128, 64, 128, 3, 2),
        (3, 64, 256, 7, 1), (3, 64, 256, 3, 1),          # first layers: row-folded kernel
    ]
    lines = []
    with torch.no_grad():
        for cin, cout, res, k, up in shapes:
            x = torch.randn(N, cin, res, res, device=dev)
            w = torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5
            flops = 2 * N * cout * cin * k * k * res * res
            exec_flops = flops * (4 if up == 2 else 1)
            ours = lambda: conv_igemm.conv2d_igemm(x, w, f=f if up == 2 else None, up=up, flip_weight=(up == 1))
            t_tma = t_tma_o = t_psw = None
            if conv_igemm.c8_input_ok(cin, res, res, k, up):
                # channel-blocked fp16 input loaded by TMA (no converter warps); and with a channel-blocked output as well
                xc = conv_igemm.to_c8(x)
                t_tma = timeit(lambda: conv_igemm.conv2d_igemm(xc, w, f=f if up == 2 else None, up=up, flip_weight=(up == 1)))
                if up == 1 and cout % 16 == 0:
                    t_tma_o = timeit(lambda: conv_igemm.conv2d_igemm(xc, w, out_c8=True))
                del xc
            if cin * cout * k * k * N * 2 <= 64e6:
                # groups = N form (per-sample weights, packed per call): the reference's fused modulated conv through conv2d_resample
                wn = w.unsqueeze(0).repeat(N, 1, 1, 1, 1)
                t_psw = timeit(lambda: conv_igemm.conv2d_igemm(x, wn, f=f if up == 2 else None, up=up, flip_weight=(up == 1), per_sample_weights=True))
                del wn
            t_wg = t_dg = t_lib_wg = t_lib_dg = None
            if up == 1 and conv_igemm.grad_supported(x.shape, w.shape, x.dtype, x.device, (1, 1), (k // 2, k // 2)):
                # training kernels: weight gradient (pg_conv2d_wgrad) and input gradient (forward kernel on dy, transposed weights) vs cuDNN (TF32 allowed)
                dy = torch.randn(N, cout, res, res, device=dev) * 1e-3
                t_wg = timeit(lambda: conv_igemm.conv2d_wgrad(x, dy, k), iters=6)
                t_dg = timeit(lambda: conv_igemm.conv2d_dgrad(dy, w), iters=6)
                cb = lambda mask: torch.ops.aten.convolution_backward(dy, x, w, None, [1, 1], [k // 2, k // 2], [1, 1], False, [0, 0], 1, mask)
                torch.backends.cudnn.allow_tf32 = True
                t_lib_wg = timeit(lambda: cb([False, True, False]), iters=6)
                t_lib_dg = timeit(lambda: cb([True, False, False]), iters=6)
                del dy
            conv_igemm.enabled = False
            lib = lambda: conv2d_resample.conv2d_resample(x, w, f=f, up=up, padding=k // 2, flip_weight=(up == 1))
            t_ours = timeit(ours)
            t_ours_tf32 = timeit(lambda: conv_igemm.conv2d_igemm(x, w, f=f if up == 2 else None, up=up, flip_weight=(up == 1), fmt='tf32'))   # kind::tf32 operands
            torch.backends.cudnn.allow_tf32 = True
            t_tf32 = timeit(lib)
            torch.backends.cudnn.allow_tf32 = False
            t_fp32 = timeit(lib)
            torch.backends.cudnn.allow_tf32 = True
            conv_igemm.enabled = True
            line = dict(cin=cin, cout=cout, res=res, k=k, up=up, N=N, gflop=round(flops / 1e9, 2),
                        ours_us=round(t_ours * 1e6, 1), ours_tf32_us=round(t_ours_tf32 * 1e6, 1), cudnn_tf32_us=round(t_tf32 * 1e6, 1), cudnn_fp32_us=round(t_fp32 * 1e6, 1),
                        tma_us=None if t_tma is None else round(t_tma * 1e6, 1), tma_c8out_us=None if t_tma_o is None else round(t_tma_o * 1e6, 1),
                        tma_tflops=None if t_tma is None else round(exec_flops / t_tma / 1e12, 1), groupsN_us=None if t_psw is None else round(t_psw * 1e6, 1),
                        wgrad_us=None if t_wg is None else round(t_wg * 1e6, 1), cudnn_tf32_wgrad_us=None if t_lib_wg is None else round(t_lib_wg * 1e6, 1),
                        dgrad_us=None if t_dg is None else round(t_dg * 1e6, 1), cudnn_tf32_dgrad_us=None if t_lib_dg is None else round(t_lib_dg * 1e6, 1),
                        ours_tflops=round(flops / t_ours / 1e12, 1), ours_executed_tflops=round(exec_flops / t_ours / 1e12, 1),
                        frac_of_bf16_peak=round(exec_flops / t_ours / 1e12 / peak, 3), speedup_vs_tf32=round(t_tf32 / t_ours, 2))
            lines.append(line)
            print(json.dumps(line), flush=True)
    if a.out:
        with open(a.out, 'w') as fh:
            for ln in lines:
                fh.write(json.dumps(ln) + '\n')


if __name__ == '__main__':
    main()
