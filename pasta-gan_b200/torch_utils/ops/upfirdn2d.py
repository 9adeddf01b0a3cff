"""upfirdn2d / filter2d / upsample2d / downsample2d on sm_100a, with the reference's call surface.

Mirrors torch_utils/ops/upfirdn2d.py of the reference (setup_filter :72, upfirdn2d :120, filter2d :272,
upsample2d :308, downsample2d :347, helpers _parse_scaling :37, _parse_padding :46, _get_filter_size :57)
— same names, argument meaning and defaults — but every call lands in ``pg_upfirdn2d`` /
``pg_upfirdn2d_bias_act`` of libpasta_b200.so.  There is no ``_upfirdn2d_ref``: CPU tensors and
``impl='ref'`` raise.

The op is linear in ``x``, so its gradient of any order is the same kernel with up<->down swapped,
mirrored padding and the filter flip toggled (reference :246-263); that is how first- and second-order
(R1) gradients are served here too.
"""
import numpy as np
import torch

from .. import misc
from . import _backend

# ---------------------------------------------------------------------------- argument helpers


def _parse_scaling(scaling):
    if isinstance(scaling, int):
        scaling = [scaling, scaling]
    assert isinstance(scaling, (list, tuple)) and len(scaling) == 2
    assert all(isinstance(v, int) for v in scaling)
    sx, sy = scaling
    assert sx >= 1 and sy >= 1
    return sx, sy


def _parse_padding(padding):
    if isinstance(padding, int):
        padding = [padding, padding]
    assert isinstance(padding, (list, tuple))
    assert all(isinstance(v, int) for v in padding)
    if len(padding) == 2:
        px, py = padding
        padding = [px, px, py, py]
    px0, px1, py0, py1 = padding
    return px0, px1, py0, py1


def _get_filter_size(f):
    """-> (fw, fh); (1, 1) for ``None``."""
    if f is None:
        return 1, 1
    assert isinstance(f, torch.Tensor) and f.ndim in [1, 2]
    with misc.suppress_tracer_warnings():
        fw, fh = int(f.shape[-1]), int(f.shape[0])
    assert fw >= 1 and fh >= 1
    return fw, fh


def setup_filter(f, device=torch.device('cpu'), normalize=True, flip_filter=False, gain=1, separable=None):
    """Build the float32 FIR tensor ``upfirdn2d`` expects.  1-D input with fewer than 8 taps becomes its
    2-D outer product (so PASTA-GAN's [1,3,3,1] arrives as a 4x4 filter); ``gain`` is split as
    gain**(ndim/2) so that separable passes multiply back to ``gain``."""
    if f is None:
        f = 1
    f = torch.as_tensor(f, dtype=torch.float32)
    assert f.ndim in [0, 1, 2] and f.numel() > 0
    if f.ndim == 0:
        f = f[np.newaxis]
    if separable is None:
        separable = (f.ndim == 1 and f.numel() >= 8)
    if f.ndim == 1 and not separable:
        f = f.ger(f)
    assert f.ndim == (1 if separable else 2)
    if normalize:
        f = f / f.sum()
    if flip_filter:
        f = f.flip(list(range(f.ndim)))
    f = f * (gain ** (f.ndim / 2))
    return f.to(device=device)


# ---------------------------------------------------------------------------- launch


def _out_hw(h, w, fh, fw, cfg):
    upx, upy, downx, downy, px0, px1, py0, py1 = cfg[:8]
    return (h * upy + py0 + py1 - fh + downy) // downy, (w * upx + px0 + px1 - fw + downx) // downx


def _launch(x, f2d, cfg, epilogue=None):
    """One pg_upfirdn2d[_bias_act] launch on x's device and torch's current stream.  ``f2d`` is rank-2."""
    capi = _backend.capi()
    upx, upy, downx, downy, px0, px1, py0, py1, flip, gain = cfg
    assert x.ndim == 4, 'x must be rank 4'
    assert f2d.ndim == 2 and f2d.dtype == torch.float32, 'f must be float32'
    if f2d.device != x.device:
        raise RuntimeError('f must reside on the same device as x')
    n, c, h, w = x.shape
    fh, fw = int(f2d.shape[0]), int(f2d.shape[1])
    oh, ow = _out_hw(h, w, fh, fw, cfg)
    if oh < 1 or ow < 1:
        raise RuntimeError('output must be at least 1x1')
    channels_last = x.ndim == 4 and x.stride(1) == 1 and c > 1 and x.is_contiguous(memory_format=torch.channels_last)
    y = torch.empty([n, c, oh, ow], dtype=x.dtype, device=x.device,
                    memory_format=torch.channels_last if channels_last else torch.contiguous_format)
    with torch.cuda.device(x.device):
        capi.require_device()
        stream = capi.current_stream(x.device)
        common = (capi.I32x4(n, c, h, w), capi.I64x4(*x.stride()), capi.I32x4(n, c, oh, ow), capi.I64x4(*y.stride()),
                  fh, fw, f2d.stride(0), f2d.stride(1), upx, upy, downx, downy, px0, px1, py0, py1, int(bool(flip)), float(gain))
        esz = x.element_size()
        sp = capi.span('upfirdn2d' if epilogue is None else 'upfirdn2d_bias_act', (x.numel() + y.numel()) * esz)
        if epilogue is None:
            rc = capi.load().pg_upfirdn2d(capi.ptr(x), capi.ptr(f2d), capi.ptr(y), *common, capi.dtype_code(x.dtype), stream)
            capi.check(rc, 'pg_upfirdn2d')
        else:
            b, act_idx, alpha, act_gain, clamp = epilogue
            rc = capi.load().pg_upfirdn2d_bias_act(capi.ptr(x), capi.ptr(f2d), capi.ptr(b), capi.ptr(y), *common,
                                                   act_idx, float(alpha), float(act_gain), float(clamp),
                                                   capi.dtype_code(x.dtype), stream)
            capi.check(rc, 'pg_upfirdn2d_bias_act')
        if sp:
            sp.close()
    return y


def _resample(x, f, cfg):
    """2-D filter: one launch.  Separable 1-D filter: a horizontal and a vertical launch, sqrt(gain) each."""
    upx, upy, downx, downy, px0, px1, py0, py1, flip, gain = cfg
    if f.ndim == 2:
        return _launch(x, f, cfg)
    g = float(np.sqrt(gain))
    y = _launch(x, f.unsqueeze(0), (upx, 1, downx, 1, px0, px1, 0, 0, flip, g))
    return _launch(y, f.unsqueeze(1), (1, upy, 1, downy, 0, 0, py0, py1, flip, g))


class _Upfirdn2d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, f, cfg):
        ctx.save_for_backward(f)
        ctx.cfg = cfg
        ctx.in_hw = (int(x.shape[2]), int(x.shape[3]))
        return _resample(x, f, cfg)

    @staticmethod
    def backward(ctx, dy):
        f, = ctx.saved_tensors
        dx = None
        if ctx.needs_input_grad[0]:
            upx, upy, downx, downy, px0, px1, py0, py1, flip, gain = ctx.cfg
            ih, iw = ctx.in_hw
            oh, ow = int(dy.shape[2]), int(dy.shape[3])
            fw, fh = _get_filter_size(f)
            tcfg = (downx, downy, upx, upy,
                    fw - px0 - 1, iw * upx - ow * downx + px0 - upx + 1,
                    fh - py0 - 1, ih * upy - oh * downy + py0 - upy + 1,
                    not flip, gain)
            dx = _Upfirdn2d.apply(dy, f, tcfg)
        assert not ctx.needs_input_grad[1], 'the FIR filter is a constant'
        return dx, None, None


# ---------------------------------------------------------------------------- public API


def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1, impl='cuda'):
    r"""Pad, zero-insert upsample by ``up``, convolve with FIR ``f`` and keep every ``down``-th pixel, per
    channel of ``x`` ``[N, C, H, W]`` (float32 / float16 / float64, NCHW or channels_last).

    ``f``: float32 ``[fh, fw]``, ``[taps]`` (separable) or ``None`` (identity).  ``up`` / ``down``: int or
    ``[x, y]``.  ``padding``: int, ``[x, y]`` or ``[x0, x1, y0, y1]`` relative to the upsampled image,
    negative = crop.  ``flip_filter=False`` is true convolution.  Differentiable to any order in ``x``.
    """
    assert isinstance(x, torch.Tensor)
    _backend.refuse_ref(impl, 'upfirdn2d')
    _backend.require_cuda(x, 'upfirdn2d')
    assert x.ndim == 4
    if f is None:
        f = torch.ones([1, 1], dtype=torch.float32, device=x.device)
    assert isinstance(f, torch.Tensor) and f.ndim in [1, 2]
    assert f.dtype == torch.float32 and not f.requires_grad
    upx, upy = _parse_scaling(up)
    downx, downy = _parse_scaling(down)
    px0, px1, py0, py1 = _parse_padding(padding)
    cfg = (upx, upy, downx, downy, px0, px1, py0, py1, bool(flip_filter), gain)
    return _Upfirdn2d.apply(x, f, cfg)


def filter2d(x, f, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """FIR-filter without resampling; default padding keeps the input size."""
    px0, px1, py0, py1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [px0 + fw // 2, px1 + (fw - 1) // 2, py0 + fh // 2, py1 + (fh - 1) // 2]
    return upfirdn2d(x, f, padding=p, flip_filter=flip_filter, gain=gain, impl=impl)


def upsample2d(x, f, up=2, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """Upsample by ``up``; default padding makes the output exactly ``up`` x the input; gain is multiplied
    by upx*upy so a DC signal keeps its magnitude."""
    upx, upy = _parse_scaling(up)
    px0, px1, py0, py1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [px0 + (fw + upx - 1) // 2, px1 + (fw - upx) // 2, py0 + (fh + upy - 1) // 2, py1 + (fh - upy) // 2]
    return upfirdn2d(x, f, up=up, padding=p, flip_filter=flip_filter, gain=gain * upx * upy, impl=impl)


def downsample2d(x, f, down=2, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """Downsample by ``down``; default padding makes the output exactly 1/``down`` of the input."""
    downx, downy = _parse_scaling(down)
    px0, px1, py0, py1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [px0 + (fw - downx + 1) // 2, px1 + (fw - downx) // 2, py0 + (fh - downy + 1) // 2, py1 + (fh - downy) // 2]
    return upfirdn2d(x, f, down=down, padding=p, flip_filter=flip_filter, gain=gain, impl=impl)


# ---------------------------------------------------------------------------- fused extension


class _Upfirdn2dBiasAct(torch.autograd.Function):
    """y = clamp(act(upfirdn2d(x) + b[c]) * act_gain) in one pass (pg_upfirdn2d_bias_act).
    Backward: dz = bias_act'(dy; y) (the grad=1 kernel), then the transposed resampling."""

    @staticmethod
    def forward(ctx, x, f, b, cfg, act, alpha, act_gain, clamp):
        from . import bias_act as ba
        spec = ba.activation_funcs[act]
        y = _launch(x, f, cfg, epilogue=(b, spec.cuda_idx, alpha, act_gain, clamp))
        ctx.save_for_backward(f, y, b if b is not None else torch.empty(0, device=x.device))
        ctx.cfg, ctx.act_cfg, ctx.in_hw, ctx.has_b = cfg, (act, alpha, act_gain, clamp), (int(x.shape[2]), int(x.shape[3])), b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        from . import bias_act as ba
        f, y, b = ctx.saved_tensors
        act, alpha, act_gain, clamp = ctx.act_cfg
        dz = ba._grad_from_y(dy, y, act=act, alpha=alpha, gain=act_gain, clamp=clamp)
        dx = db = None
        if ctx.needs_input_grad[0]:
            upx, upy, downx, downy, px0, px1, py0, py1, flip, gain = ctx.cfg
            ih, iw = ctx.in_hw
            oh, ow = int(dy.shape[2]), int(dy.shape[3])
            fw, fh = _get_filter_size(f)
            tcfg = (downx, downy, upx, upy, fw - px0 - 1, iw * upx - ow * downx + px0 - upx + 1,
                    fh - py0 - 1, ih * upy - oh * downy + py0 - upy + 1, not flip, gain)
            dx = _Upfirdn2d.apply(dz, f, tcfg)
        if ctx.has_b and ctx.needs_input_grad[2]:
            db = dz.sum([0, 2, 3])
        return dx, None, db, None, None, None, None, None


def upfirdn2d_bias_act(x, f, b=None, up=1, down=1, padding=0, flip_filter=False, gain=1,
                       act='linear', alpha=None, act_gain=None, clamp=None):
    """Extension (not in the reference): ``bias_act(upfirdn2d(x, f, ...), b, act=..., gain=act_gain, clamp=...)``
    as ONE kernel — the FIR output never round-trips through HBM.  2-D filters and linear/relu/lrelu only."""
    from . import bias_act as ba
    _backend.require_cuda(x, 'upfirdn2d_bias_act')
    assert f is not None and f.ndim == 2 and f.dtype == torch.float32
    spec = ba.activation_funcs[act]
    assert act in ('linear', 'relu', 'lrelu')
    alpha = float(alpha if alpha is not None else spec.def_alpha)
    act_gain = float(act_gain if act_gain is not None else spec.def_gain)
    clamp = float(clamp if clamp is not None else -1)
    upx, upy = _parse_scaling(up)
    downx, downy = _parse_scaling(down)
    px0, px1, py0, py1 = _parse_padding(padding)
    cfg = (upx, upy, downx, downy, px0, px1, py0, py1, bool(flip_filter), gain)
    if b is not None:
        b = b.contiguous()
    return _Upfirdn2dBiasAct.apply(x, f, b, cfg, act, alpha, act_gain, clamp)
