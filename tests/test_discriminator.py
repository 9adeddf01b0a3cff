"""Discriminator mirror: parameter parity + forward on CPU (oracle table), and on the GPU the full R1 path — logits, gradient of the
logits wrt the image, the R1 penalty and parameter gradients of the penalty (double backward through conv / upfirdn2d / bias_act-lrelu /
minibatch-std / FC) — against the golden vectors produced by the unmodified reference (loss_wo_flow_fullbody.py:231-254 shape of use)."""
import numpy as np
import pytest
import torch

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
@pytest.mark.parametrize('tensor_cores', [False, True, 'r1'], ids=['library_fp32', 'tcgen05_bf16', 'tcgen05_bf16_incl_r1_inner_grad'])
def test_discriminator_r1_double_backward_gpu(golden, tensor_cores):
    """The complete R1 chain (logits -> image gradient under no_weight_gradients -> penalty -> second-order parameter gradients) against the
    reference's CPU run.  library_fp32: every convolution on the fp32 library path (tight bounds).  tcgen05_bf16: the stride-1 'same' fp32
    convolutions run forward / dgrad / wgrad on the tcgen05 kernels with bf16 operands -- the north_star's 1e-2 class for tensor-core convolutions."""
    from pasta_gan_b200.torch_utils.ops import conv2d_gradfix
    g = golden('discriminator')
    meta = g.meta[0]
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, conv2d_gradfix.enabled, conv2d_gradfix.tensor_core_training,
           conv2d_gradfix.tensor_core_min_flops, conv2d_gradfix.tensor_core_r1)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    conv2d_gradfix.enabled = True
    conv2d_gradfix.tensor_core_training = bool(tensor_cores)
    conv2d_gradfix.tensor_core_min_flops = 0                     # every eligible layer, whatever its size
    conv2d_gradfix.tensor_core_r1 = tensor_cores == 'r1'
    # bf16 operands (8-bit mantissa): forward / first-order passes within the tensor-core class of the north_star; with the R1 inner gradient on bf16 as
    # well (not the default) the ~1e-7-sized image gradient is only good to 2e-1 pointwise, which is why that pass defaults to the fp32 library path
    tol = dict(logits=1e-4, grad=1e-2, pen=1e-3, params=1e-2)
    if tensor_cores is True:
        tol = dict(logits=1e-2, grad=5e-2, pen=3e-2, params=5e-2)
    if tensor_cores == 'r1':
        tol = dict(logits=1e-2, grad=4e-1, pen=5e-2, params=2e-1)
    try:
        D = N.build_discriminator(num_fp16_res=0)
        procedural.fill_(D)
        D.to('cuda').requires_grad_(True)
        img, c = _inputs('cuda')
        img.requires_grad_(True)
        logits = D(img, c)
        assert rel_err(logits, g.t('logits')) < tol['logits']
        with conv2d_gradfix.no_weight_gradients():
            r1_grad, = torch.autograd.grad(logits.sum(), img, create_graph=True)
        e_grad = rel_err(r1_grad[:, :, ::4, ::4], g.t('r1_grad'))
        penalty = r1_grad.square().sum([1, 2, 3])
        e_pen = rel_err(penalty, g.t('penalty'))
        print('r1 grad rel err', e_grad, 'penalty rel err', e_pen, 'penalty', penalty.tolist(), 'golden', g.t('penalty').tolist())
        # pointwise image gradient: ~20 fp32 convolutions deep in two different libraries (oneDNN vs cuDNN): 1e-2 max-abs; its norm (the R1
        # penalty) is held to 1e-3
        assert e_grad < tol['grad']
        assert e_pen < tol['pen']
        (penalty.mean() * 5).backward()
        params = dict(D.named_parameters())
        errs = {}
        for name in meta['grad_names']:
            gr = params[name].grad
            ref = g.t('grad/' + name)
            got = gr if gr.numel() <= 70000 else gr.flatten()[::37]
            errs[name] = rel_err(got, ref)
        print('second-order parameter gradient rel errs', errs)
        assert all(v < tol['params'] for v in errs.values()), errs
    finally:
        (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, conv2d_gradfix.enabled, conv2d_gradfix.tensor_core_training,
         conv2d_gradfix.tensor_core_min_flops, conv2d_gradfix.tensor_core_r1) = old


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
