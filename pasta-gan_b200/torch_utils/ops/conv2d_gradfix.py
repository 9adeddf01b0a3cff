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


# Opt-in: run the stride-1 'same' 1x1 / 3x3 convolutions of the custom op (forward AND the input-gradient conv) on the tcgen05 kernel.
# Off by default: at the training batch (4 per GPU) it measured 2x slower than cuDNN (weights are re-packed every step, tiles under-fill the
# GPU) and fp16 operands cost ~1.5e-4 on D's logits; the training path keeps fp32 library convolutions until dgrad/wgrad kernels exist.
tensor_core_forward = os.environ.get('PASTA_B200_TC_TRAIN', '0') == '1'


def _forward(input, weight, bias, transpose, stride, padding, output_padding, dilation, groups):
    """Dense convolution of the custom autograd op.  Called with grad mode off (inside Function.forward), so the tensor-core kernel may be
    used for the shapes it covers: a stride-1 'same' conv, and the stride-1 transposed conv that is its input gradient
    (conv_transpose2d(x, w) == conv2d(x, w^T mirrored))."""
    if tensor_core_forward and bias is None and groups == 1 and stride == (1, 1) and dilation == (1, 1) and output_padding == (0, 0) \
            and input.dtype == torch.float32 and weight.shape[2] == weight.shape[3] and padding == (weight.shape[2] // 2,) * 2:
        from . import conv_igemm
        w = weight.transpose(0, 1) if transpose else weight
        if conv_igemm.supported(input, w, padding=(padding[0],) * 4):
            return conv_igemm.conv2d_igemm(input, w, flip_weight=not transpose)
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
