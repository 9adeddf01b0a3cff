"""conv2d / conv_transpose2d with the reference's call surface (torch_utils/ops/conv2d_gradfix.py:22-43):
``enabled``, ``weight_gradients_disabled``, ``no_weight_gradients()``, ``conv2d``, ``conv_transpose2d``.

Unlike the reference on torch >= 1.10 (where :47-56 silently degrades to plain F.conv2d and
``no_weight_gradients`` stops having any effect), the custom autograd path here is always live for CUDA
tensors when ``enabled`` is set: forward is the dense convolution, the input gradient is the
opposite-transpose convolution (itself differentiable, which is what R1's double backward needs), and the
weight gradient is skipped while ``no_weight_gradients()`` is active (loss_wo_flow_fullbody.py:246).

Dense convolutions in this module are the library call the reference also makes (cuDNN through ATen);
the hand-written tcgen05 implicit-GEMM kernels take over from ``conv2d_resample`` / ``modulated_conv2d``
for the shapes they cover.
"""
import contextlib

import os

import torch

enabled = False                     # the reference's training loop sets this to True (training_loop...py:255)
weight_gradients_disabled = False


@contextlib.contextmanager
def no_weight_gradients():
    global weight_gradients_disabled
    old = weight_gradients_disabled
    weight_gradients_disabled = True
    try:
        yield
    finally:
        weight_gradients_disabled = old


def _pair(v):
    v = tuple(v) if isinstance(v, (tuple, list)) else (v, v)
    assert len(v) == 2 and all(isinstance(t, int) for t in v)
    return v


def _use_custom(input):
    assert isinstance(input, torch.Tensor)
    return enabled and input.device.type == 'cuda' and torch.backends.cudnn.enabled


def conv2d(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
    if _use_custom(input):
        return _Conv.apply(input, weight, bias, False, _pair(stride), _pair(padding), (0, 0), _pair(dilation), groups)
    return torch.nn.functional.conv2d(input=input, weight=weight, bias=bias, stride=stride, padding=padding, dilation=dilation, groups=groups)


def conv_transpose2d(input, weight, bias=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1):
    if _use_custom(input):
        return _Conv.apply(input, weight, bias, True, _pair(stride), _pair(padding), _pair(output_padding), _pair(dilation), groups)
    return torch.nn.functional.conv_transpose2d(input=input, weight=weight, bias=bias, stride=stride, padding=padding,
                                                output_padding=output_padding, groups=groups, dilation=dilation)


# Training on the tensor cores: the stride-1 'same' 1x1 / 3x3 fp32 convolutions of the custom op run their forward, their input gradient (the same
# tcgen05 kernel on dy with transposed, mirrored weights -- itself differentiable, which R1's double backward needs) and their weight gradient
# (pg_conv2d_wgrad) on tcgen05 with fp16 (forward) / bf16 (gradient) operands and fp32 accumulation.  Everything else (strided / transposed-strided forms, fp16 blocks of the
# discriminator, grouped convs) stays on the library path.  PASTA_B200_TC_TRAIN=0 switches it off (pure library convolutions).
tensor_core_training = os.environ.get('PASTA_B200_TC_TRAIN', '1') != '0'
# operand formats: the forward convolution multiplies activations and weights (O(1), clamped at 256 in the synthesis layers) -> fp16, 11-bit mantissa,
# as in inference; gradients are tiny (R1's are ~1e-7) -> bf16 for the exponent range.  (Forward in bf16 measured 2e-1 pointwise on R1's image gradient.)
tensor_core_format = os.environ.get('PASTA_B200_TC_TRAIN_FMT', 'fp16')
tensor_core_grad_format = os.environ.get('PASTA_B200_TC_TRAIN_GRAD_FMT', 'bf16')
# Layers below this many FLOPs per call stay on the library path: at the training batch (4 per GPU) a 512-channel layer at 4^2..16^2 is a few
# microseconds of math behind a weight tensor that has to be re-packed every step; the tensor cores pay where pixels, not weights, dominate (SPADE
# blocks, >= 32^2 layers).  Measured with the library in fp32 (the trainer's default): 8 GFLOP 72.2 ms / step, 3 GFLOP 70.6, 1 GFLOP 68.1.  (Against
# cuDNN TF32 the break-even was higher: 65.8 ms with every layer on tcgen05 vs 57.8 with an 8 GFLOP floor.)
tensor_core_min_flops = float(os.environ.get('PASTA_B200_TC_TRAIN_MIN_GFLOP', '1')) * 1e9
# R1 (loss_wo_flow_fullbody.py:231-254) differentiates the image gradient a second time and its values are ~1e-7: with 10-bit-mantissa products in the
# forward pass the pointwise image gradient measured 9e-2 and some second-order parameter gradients 3e-1 against the reference (2e-1 / 4e-1 with bf16),
# while the fp32 library path holds 1e-2.  So the tensor cores serve the first-order phases (Gmain, Dmain) and the trainer runs the Dreg phase -- one
# iteration in 16 -- under tensor_cores(False); the inner gradient taken under no_weight_gradients() never uses them unless PASTA_B200_TC_TRAIN_R1=1.
tensor_core_r1 = os.environ.get('PASTA_B200_TC_TRAIN_R1', '0') == '1'

@contextlib.contextmanager
def tensor_cores(on):
    """Switch the tcgen05 training path on / off for a region (the trainer runs the R1 phase, whose second-order gradients need fp32 products, with
    it off)."""
    global tensor_core_training
    old = tensor_core_training
    tensor_core_training = bool(on) and old
    try:
        yield
    finally:
        tensor_core_training = old


def _tc_ok(input, weight_shape, transpose, stride, padding, output_padding, dilation, groups):
    if not (tensor_core_training and output_padding == (0, 0) and weight_shape[2] in (1, 3)):
        return False
    if weight_gradients_disabled and not tensor_core_r1:
        return False
    from . import conv_igemm
    if not conv_igemm.grad_supported(input.shape, weight_shape, input.dtype, input.device, stride, padding, dilation, groups):
        return False
    flops = 2.0 * input.shape[0] * input.shape[2] * input.shape[3] * weight_shape[0] * weight_shape[1] * weight_shape[2] * weight_shape[3]
    return flops >= tensor_core_min_flops


def _forward(input, weight, bias, transpose, stride, padding, output_padding, dilation, groups):
    """Dense convolution of the custom autograd op.  Called with grad mode off (inside Function.forward), so the tensor-core kernel may be
    used for the shapes it covers: a stride-1 'same' conv, and the stride-1 transposed conv that is its input gradient
    (conv_transpose2d(x, w) == conv2d(x, w^T mirrored))."""
    if bias is None and _tc_ok(input, weight.shape, transpose, stride, padding, output_padding, dilation, groups):
        from . import conv_igemm
        w = weight.transpose(0, 1) if transpose else weight
        return conv_igemm.conv2d_igemm(input, w, flip_weight=not transpose, fmt=tensor_core_grad_format if transpose else tensor_core_format)
    if not transpose:
        return torch.nn.functional.conv2d(input, weight, bias, stride, padding, dilation, groups)
    return torch.nn.functional.conv_transpose2d(input, weight, bias, stride, padding, output_padding, groups, dilation)


class _Conv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, weight, bias, transpose, stride, padding, output_padding, dilation, groups):
        ctx.save_for_backward(input, weight)
        ctx.cfg = (transpose, stride, padding, output_padding, dilation, groups)
        ctx.has_bias = bias is not None
        return _forward(input, weight, bias, transpose, stride, padding, output_padding, dilation, groups)

    @staticmethod
    def backward(ctx, grad_output):
        input, weight = ctx.saved_tensors
        transpose, stride, padding, output_padding, dilation, groups = ctx.cfg
        grad_input = grad_weight = grad_bias = None
        if ctx.needs_input_grad[0]:
            # input gradient = the opposite-transpose convolution with the same weights
            if transpose:
                op = (0, 0)
            else:
                op = tuple(input.shape[i + 2] - (grad_output.shape[i + 2] - 1) * stride[i] - (1 - 2 * padding[i])
                           - dilation[i] * (weight.shape[i + 2] - 1) for i in range(2))
            grad_input = _Conv.apply(grad_output, weight, None, not transpose, stride, padding, op, dilation, groups)
            assert grad_input.shape == input.shape
        if ctx.needs_input_grad[1] and not weight_gradients_disabled:
            grad_weight = _ConvGradWeight.apply(grad_output, input, weight.shape, ctx.cfg)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            grad_bias = grad_output.sum([0, 2, 3])
        return grad_input, grad_weight, grad_bias, None, None, None, None, None, None


class _ConvGradWeight(torch.autograd.Function):
    @staticmethod
    def forward(ctx, grad_output, input, weight_shape, cfg):
        transpose, stride, padding, output_padding, dilation, groups = cfg
        ctx.save_for_backward(grad_output, input)
        ctx.cfg, ctx.weight_shape = cfg, weight_shape
        if not transpose and grad_output.dtype == torch.float32 and _tc_ok(input, weight_shape, transpose, stride, padding, output_padding, dilation, groups):
            from . import conv_igemm
            return conv_igemm.conv2d_wgrad(input, grad_output, int(weight_shape[2]))
        dummy_w = input.new_empty(weight_shape)
        _, gw, _ = torch.ops.aten.convolution_backward(grad_output, input, dummy_w, None, list(stride), list(padding), list(dilation),
                                                       transpose, list(output_padding), groups, [False, True, False])
        return gw

    @staticmethod
    def backward(ctx, g2_weight):
        grad_output, input = ctx.saved_tensors
        transpose, stride, padding, output_padding, dilation, groups = ctx.cfg
        g2_grad_output = g2_input = None
        if ctx.needs_input_grad[0]:
            g2_grad_output = _Conv.apply(input, g2_weight, None, transpose, stride, padding, output_padding, dilation, groups)
        if ctx.needs_input_grad[1]:
            if transpose:
                op = (0, 0)
            else:
                op = tuple(input.shape[i + 2] - (grad_output.shape[i + 2] - 1) * stride[i] - (1 - 2 * padding[i])
                           - dilation[i] * (ctx.weight_shape[i + 2] - 1) for i in range(2))
            g2_input = _Conv.apply(grad_output, g2_weight, None, not transpose, stride, padding, op, dilation, groups)
        return g2_grad_output, g2_input, None, None
