#!/usr/bin/env python
"""Driver for `ncu --set full`: two launches each of the hand-written kernels at their largest generator shapes (batch 16)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pasta_gan_b200.torch_utils.ops import upfirdn2d, bias_act, conv_igemm, torgb
dev = torch.device('cuda:0')
f = upfirdn2d.setup_filter([1, 3, 3, 1]).to(dev)
with torch.no_grad():
    x = torch.randn(16, 64, 256, 256, device=dev); b = torch.randn(64, device=dev)
    x257 = torch.randn(16, 64, 257, 257, device=dev)
    for _ in range(2):
        bias_act.bias_act(x, b, act='lrelu', clamp=256)                                             # bias_act_vec_kernel
        upfirdn2d.upfirdn2d(x257, f, padding=[1, 1, 1, 1], gain=4)                                  # upfirdn2d_band_kernel<1,0>
        upfirdn2d.upfirdn2d_bias_act(x257, f, b, padding=[1, 1, 1, 1], gain=4, act='lrelu', clamp=256)   # <1,1> fused epilogue
        upfirdn2d.downsample2d(x, f)                                                                # <2,0>
        upfirdn2d.upsample2d(x[:, :, :128, :128].contiguous(), f)                                   # upfirdn2d_up2_band_kernel (polyphase up-2)
    w3 = torch.randn(3, 64, 1, 1, device=dev); s = torch.randn(16, 64, device=dev) / 8; img = torch.randn(16, 3, 128, 128, device=dev)
    for _ in range(2):
        torgb.torgb_skip(x, w3, styles=s, bias=torch.zeros(3, device=dev), clamp=256, img=img, f=f)  # torgb_skip_kernel<3>
    xa = torch.randn(16, 256, 128, 128, device=dev); wa = torch.randn(128, 256, 3, 3, device=dev) / 48
    xb = torch.randn(16, 128, 128, 128, device=dev); wb = torch.randn(128, 128, 3, 3, device=dev) / 34
    xac, xbc = conv_igemm.to_c8(xa), conv_igemm.to_c8(xb)
    xs = torch.randn(16, 128, 128, 128, device=dev); wg = torch.randn(128, 128, 3, 3, device=dev) / 34; wbeta = torch.randn(128, 128, 3, 3, device=dev) / 34
    dyb = torch.randn(16, 128, 128, 128, device=dev) * 1e-3
    for _ in range(2):
        conv_igemm.conv2d_igemm(xa, wa, bias=torch.zeros(128, device=dev), act='lrelu', gain=2 ** 0.5, clamp=256)   # 256->128 @128^2, converter path (conv_igemm_kernel)
        conv_igemm.conv2d_igemm(xb, wb)                                                                            # 128->128 @128^2, converter path
        conv_igemm.conv2d_igemm(xac, wa, act='relu', out_c8=True)                                                  # 256->128 @128^2, TMA persistent kernel, channel-blocked in / out
        conv_igemm.conv2d_igemm(xbc, wb)                                                                           # 128->128 @128^2, TMA persistent kernel, fp32 out
        conv_igemm.spade_conv_norm(xs, xbc, wg, wbeta, act='relu', gain=1.0, out_c8=True)                          # SPADE gamma|beta 128->256, TMA persistent
        conv_igemm.conv2d_wgrad(xb, dyb, 3)                                                                        # conv_wgrad_kernel + reduce
    x3 = torch.randn(32, 3, 256, 256, device=dev); w7 = torch.randn(64, 3, 7, 7, device=dev) / 12
    for _ in range(2):
        conv_igemm.conv2d_igemm(x3, w7, bias=torch.zeros(64, device=dev), act='relu', gain=2 ** 0.5, out_c8=True)  # conv_rowfold_kernel: 3->64 7x7 @256^2, batch 32 (garment encoder stem)
        conv_igemm.conv2d_igemm(xb, wb, fmt='tf32')                                                                # 128->128 @128^2, kind::tf32 operands (conv_igemm_kernel)
    x32 = torch.randn(16, 512, 32, 32, device=dev); w512 = torch.randn(3, 512, 1, 1, device=dev); s512 = torch.randn(16, 512, device=dev) / 16
    xd = conv_igemm.to_c8(torch.randn(16, 64, 256, 256, device=dev)); wd = torch.randn(128, 64, 3, 3, device=dev) / 24
    for _ in range(2):
        torgb.torgb_skip(x32, w512, styles=s512, bias=torch.zeros(3, device=dev), clamp=256, img=torch.randn(16, 3, 16, 16, device=dev), f=f)   # torgb_skip_splitc_kernel<3>
        conv_igemm.conv2d_igemm(xd, wd, f=f, down=2, act='lrelu', gain=2 ** 0.5, out_c8=True)                      # down-2 64->128 @256^2 from C8: strided 4-D TMA box
    # patch routing: all rectifying warps of a batch in one launch + the back-warp composite
    from pasta_gan_b200 import patch_routing, synthetic
    d = synthetic.synth_patch_routing_inputs(16, seed=3)
    t = {k: torch.from_numpy(v).to(dev) for k, v in d.items() if k != 'keypoints'}
    for _ in range(2):
        patch_routing.PatchRouter().normalize(t['upper_img'], t['lower_img'], t['upper_clothes_mask'], t['lower_clothes_mask'], d['keypoints'], 2)
torch.cuda.synchronize()
print('ok')
