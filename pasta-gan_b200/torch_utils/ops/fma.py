"""``fma(a, b, c) = a * b + c`` with broadcast-aware gradients (reference: torch_utils/ops/fma.py:15-60).
Used by the non-fused modulated conv to apply demodulation and noise in one elementwise pass
(training/networks.py:77)."""
import torch


def fma(a, b, c):
    return _Fma.apply(a, b, c)


def _reduce_to(g, shape):
    """Sum a broadcast gradient back down to ``shape``."""
    lead = g.ndim - len(shape)
    assert lead >= 0
    dims = [i for i in range(g.ndim) if g.shape[i] > 1 and (i < lead or shape[i - lead] == 1)]
    if dims:
        g = g.sum(dim=dims, keepdim=True)
    if lead:
        g = g.reshape(-1, *g.shape[lead + 1:])
    assert g.shape == shape
    return g


class _Fma(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, c):
        ctx.save_for_backward(a, b)
        ctx.c_shape = c.shape
        return torch.addcmul(c, a, b)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        da = _reduce_to(g * b, a.shape) if ctx.needs_input_grad[0] else None
        db = _reduce_to(g * a, b.shape) if ctx.needs_input_grad[1] else None
        dc = _reduce_to(g, ctx.c_shape) if ctx.needs_input_grad[2] else None
        return da, db, dc
