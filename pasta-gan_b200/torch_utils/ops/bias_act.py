"""bias_act on sm_100a with the reference's call surface.

Mirrors torch_utils/ops/bias_act.py of the reference (``activation_funcs`` :23-33, ``bias_act`` :55;
autograd structure of ``BiasActCuda`` :145-175 and ``BiasActCudaGrad`` :178-205) — same names, defaults,
saved-tensor policy and second-order behaviour — with every evaluation done by ``pg_bias_act`` of
libpasta_b200.so (grad = 0 / 1 / 2 modes).  CPU tensors and ``impl='ref'`` raise: there is no
``_bias_act_ref`` here.
"""
import numpy as np
import torch

from . import _backend

try:                                    # the reference builds the table from dnnlib.EasyDict (:23)
    from dnnlib import EasyDict as _EasyDict
except Exception:                       # standalone: attribute-dict with the same behaviour
    class _EasyDict(dict):
        def __getattr__(self, name):
            try:
                return self[name]
            except KeyError:
                raise AttributeError(name)

        def __setattr__(self, name, value):
            self[name] = value

        def __delattr__(self, name):
            del self[name]


def _spec(func, def_alpha, def_gain, cuda_idx, ref, has_2nd_grad):
    return _EasyDict(func=func, def_alpha=def_alpha, def_gain=def_gain, cuda_idx=cuda_idx, ref=ref, has_2nd_grad=has_2nd_grad)


_F = torch.nn.functional
# ``func`` is kept for API compatibility (networks read .def_gain; nothing here evaluates .func).
activation_funcs = {
    'linear':   _spec(lambda x, **_: x,                              0,   1,          1, '',  False),
    'relu':     _spec(lambda x, **_: _F.relu(x),                     0,   np.sqrt(2), 2, 'y', False),
    'lrelu':    _spec(lambda x, alpha, **_: _F.leaky_relu(x, alpha), 0.2, np.sqrt(2), 3, 'y', False),
    'tanh':     _spec(lambda x, **_: torch.tanh(x),                  0,   1,          4, 'y', True),
    'sigmoid':  _spec(lambda x, **_: torch.sigmoid(x),               0,   1,          5, 'y', True),
    'elu':      _spec(lambda x, **_: _F.elu(x),                      0,   1,          6, 'y', True),
    'selu':     _spec(lambda x, **_: _F.selu(x),                     0,   1,          7, 'y', True),
    'softplus': _spec(lambda x, **_: _F.softplus(x),                 0,   1,          8, 'y', True),
    'swish':    _spec(lambda x, **_: torch.sigmoid(x) * x,           0,   np.sqrt(2), 9, 'x', True),
}


def _memory_format(t):
    return torch.channels_last if t.ndim == 4 and t.stride(1) == 1 and t.shape[1] > 1 else torch.contiguous_format


def _dense(t, memory_format):
    return t.contiguous(memory_format=memory_format) if t.ndim == 4 else t.contiguous()


def _kernel(x, b, xref, yref, dy, grad, dim, cfg):
    """One pg_bias_act launch.  x / xref / yref / dy share x's dense layout; None = absent."""
    capi = _backend.capi()
    act_idx, alpha, gain, clamp = cfg
    for t in (xref, yref, dy):
        if t is not None and t.numel() and (t.shape != x.shape or t.stride() != x.stride() or t.dtype != x.dtype):
            raise RuntimeError('xref / yref / dy must have the same shape, dtype and layout as x')
    if b is not None and b.numel():
        if b.ndim != 1:
            raise RuntimeError('b must have rank 1')
        if not (0 <= dim < x.ndim):
            raise RuntimeError('dim is out of bounds')
        if b.shape[0] != x.shape[dim]:
            raise RuntimeError('b has wrong number of elements')
        if b.dtype != x.dtype or b.device != x.device:
            raise RuntimeError('b must have the same dtype and device as x')
        step_b = x.stride(dim) if x.shape[dim] > 1 else (1 << 20)     # size-1 dim: any step selects b[0]
        size_b = b.shape[0]
    else:
        b, step_b, size_b = None, 1, 0
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        capi.require_device()
        nops = 2 + (xref is not None) + (yref is not None) + (dy is not None)
        sp = capi.span('bias_act' if grad == 0 else f'bias_act_grad{grad}', nops * x.numel() * x.element_size())
        rc = capi.load().pg_bias_act(capi.ptr(x), capi.ptr(b), capi.ptr(xref), capi.ptr(yref), capi.ptr(dy), capi.ptr(y),
                                     x.numel(), size_b, step_b, grad, act_idx, alpha, gain, clamp,
                                     capi.dtype_code(x.dtype), capi.current_stream(x.device))
        capi.check(rc, 'pg_bias_act')
        if sp:
            sp.close()
    return y


def _sum_to_bias(t, dim):
    dims = [i for i in range(t.ndim) if i != dim]
    return t.sum(dims) if dims else t          # 1-D input: nothing to reduce (sum([]) would reduce everything)


class _BiasAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, b, dim, act, cfg):
        spec = activation_funcs[act]
        _, alpha, gain, clamp = cfg
        ctx.memory_format = _memory_format(x)
        x = _dense(x, ctx.memory_format)
        b = b.contiguous() if b is not None else None
        trivial = act == 'linear' and gain == 1 and clamp < 0
        y = x if (trivial and b is None) else _kernel(x, b, None, None, None, 0, dim, cfg)
        keep_x = 'x' in spec.ref or spec.has_2nd_grad
        # y is also kept for 'linear' when clamping: the gradient must vanish where the output saturated, as in the
        # reference's impl='ref' path (autograd through x.clamp, bias_act.py:121-122).  The reference's CUDA plugin
        # passes no yref for 'linear' (ref='' at :24) and therefore never masks — we follow the ref path.
        keep_y = 'y' in spec.ref or (act == 'linear' and clamp >= 0)
        ctx.save_for_backward(x if keep_x else None, b if keep_x else None, y if keep_y else None)
        ctx.dim, ctx.act, ctx.cfg, ctx.trivial = dim, act, cfg, trivial
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = _dense(dy, ctx.memory_format)
        x, b, y = ctx.saved_tensors
        dx = db = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            dx = dy if ctx.trivial else _BiasActGrad.apply(dy, x, b, y, ctx.dim, ctx.act, ctx.cfg)
        if ctx.needs_input_grad[1]:
            db = _sum_to_bias(dx, ctx.dim)
        return dx, db, None, None, None


class _BiasActGrad(torch.autograd.Function):
    """dx = dy * gain * act'(.), itself differentiable: wrt dy it is the same op, wrt x it is the grad=2 kernel."""

    @staticmethod
    def forward(ctx, dy, x, b, y, dim, act, cfg):
        spec = activation_funcs[act]
        ctx.memory_format = _memory_format(dy)
        dx = _kernel(dy, b, x, y, None, 1, dim, cfg)
        ctx.save_for_backward(dy if spec.has_2nd_grad else None, x, b, y)
        ctx.dim, ctx.act, ctx.cfg = dim, act, cfg
        return dx

    @staticmethod
    def backward(ctx, d_dx):
        d_dx = _dense(d_dx, ctx.memory_format)
        dy, x, b, y = ctx.saved_tensors
        spec = activation_funcs[ctx.act]
        d_dy = d_x = d_b = None
        if ctx.needs_input_grad[0]:
            d_dy = _BiasActGrad.apply(d_dx, x, b, y, ctx.dim, ctx.act, ctx.cfg)
        if spec.has_2nd_grad and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2]):
            d_x = _kernel(d_dx, b, x, y, dy, 2, ctx.dim, ctx.cfg)
        if spec.has_2nd_grad and ctx.needs_input_grad[2]:
            d_b = _sum_to_bias(d_x, ctx.dim)
        return d_dy, d_x, d_b, None, None, None, None


def _resolve(act, alpha, gain, clamp):
    assert clamp is None or clamp >= 0
    spec = activation_funcs[act]
    alpha = float(alpha if alpha is not None else spec.def_alpha)
    gain = float(gain if gain is not None else spec.def_gain)
    clamp = float(clamp if clamp is not None else -1)
    return (int(spec.cuda_idx), alpha, gain, clamp)


def _grad_from_y(dy, y, act, alpha, gain, clamp):
    """dz for the fused FIR+bias_act backward (relu / lrelu / linear: needs y only)."""
    cfg = (int(activation_funcs[act].cuda_idx), float(alpha), float(gain), float(clamp))
    dy = _dense(dy, _memory_format(y))
    return _BiasActGrad.apply(dy, None, None, y, 1, act, cfg)


def bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None, impl='cuda'):
    r"""``clamp(act(x + b) * gain, -clamp, clamp)`` in one pass over ``x`` (any shape, float32/16/64).

    ``b``: 1-D bias matching ``x.shape[dim]`` or ``None``.  ``act``: a key of ``activation_funcs``.
    ``alpha`` / ``gain``: ``None`` selects the activation's default.  ``clamp``: ``None`` disables.
    First- and second-order gradients are supported (third order is not), exactly as in the reference.
    """
    assert isinstance(x, torch.Tensor)
    _backend.refuse_ref(impl, 'bias_act')
    _backend.require_cuda(x, 'bias_act')
    cfg = _resolve(act, alpha, gain, clamp)
    return _BiasAct.apply(x, b, dim, act, cfg)
