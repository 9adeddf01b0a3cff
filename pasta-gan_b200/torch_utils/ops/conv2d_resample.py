"""conv2d_resample: convolution fused with FIR up/down-sampling, padding applied once up front.

Same contract as torch_utils/ops/conv2d_resample.py:59-154 of the reference (argument names, the padding
algebra of :94-104, and the lowering chosen for each (kernel, up, down) combination, which fixes the
rounding behaviour): 1x1+down -> FIR-decimate then conv; 1x1+up -> conv then FIR-upsample; kxk+down ->
FIR (padded) then strided conv; up -> transposed strided conv then FIR with gain up^2; plain -> conv.
All FIR stages run on the sm_100a upfirdn2d kernels.
"""
import torch

from .. import misc
from . import _backend
from . import conv2d_gradfix
from . import conv_igemm
from . import upfirdn2d
from .upfirdn2d import _get_filter_size, _parse_padding


def _get_weight_shape(w):
    with misc.suppress_tracer_warnings():
        return [int(sz) for sz in w.shape]


def _conv2d_wrapper(x, w, stride=1, padding=0, groups=1, transpose=False, flip_weight=True):
    """``flip_weight=True`` is cross-correlation (what conv2d computes); ``False`` flips the taps first."""
    if not flip_weight:
        w = w.flip([2, 3])
    py, px = (padding, padding) if isinstance(padding, int) else tuple(padding)
    if not transpose and stride == 1 and conv_igemm.supported(x, w, groups=groups, padding=(px, px, py, py)):
        return _igemm(x, w, groups, flip_weight=True)                # plain stride-1 'same' conv inside a lowering; w already carries the flip
    op = conv2d_gradfix.conv_transpose2d if transpose else conv2d_gradfix.conv2d
    sp = _backend.capi().span('library_conv(cudnn)') if x.is_cuda else None
    y = op(x, w, stride=stride, padding=padding, groups=groups)
    if sp:
        kh, kw = int(w.shape[2]), int(w.shape[3])
        taps = (x.shape[2] * x.shape[3]) if transpose else (y.shape[2] * y.shape[3])
        cin_g, cout = (int(w.shape[0]) // groups, int(w.shape[1]) * groups) if transpose else (int(w.shape[1]), int(w.shape[0]))
        sp.flops = 2 * int(x.shape[0]) * cout * cin_g * kh * kw * int(taps)
        sp.nbytes = (x.numel() + y.numel() + w.numel()) * x.element_size()
        sp.close()
    return y


def _igemm(x, w, groups, **kw):
    """tcgen05 kernel call; ``groups = N`` is the reference's fused modulated convolution (x [1, N*I, H, W], w [N*O, I, k, k],
    training/networks.py:88-90): the same GEMM with one weight set per sample, x viewed as [N, I, H, W]."""
    if groups == 1:
        return conv_igemm.conv2d_igemm(x, w, **kw)
    cout, cin_g, kh, kwid = (int(v) for v in w.shape)
    y = conv_igemm.conv2d_igemm(x.reshape(groups, cin_g, *x.shape[2:]), w.reshape(groups, cout // groups, cin_g, kh, kwid), per_sample_weights=True, **kw)
    return y.reshape(1, cout, *y.shape[2:])


@misc.profiled_function
def conv2d_resample(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False):
    r"""x ``[N, Cin, H, W]``, w ``[Cout, Cin // groups, kh, kw]``, f from ``upfirdn2d.setup_filter`` (or None),
    integer ``up`` / ``down``, ``padding`` relative to the upsampled image."""
    assert isinstance(x, torch.Tensor) and x.ndim == 4
    assert isinstance(w, torch.Tensor) and w.ndim == 4 and w.dtype == x.dtype
    assert f is None or (isinstance(f, torch.Tensor) and f.ndim in [1, 2] and f.dtype == torch.float32)
    assert isinstance(up, int) and up >= 1
    assert isinstance(down, int) and down >= 1
    assert isinstance(groups, int) and groups >= 1
    cout, cin_g, kh, kw = _get_weight_shape(w)
    taps = _get_filter_size(f)                                   # (fw, fh)
    pad = list(_parse_padding(padding))                          # [x0, x1, y0, y1], relative to the up-sampled image

    # Padding owed to the resampling filters, so that sizes come out as H * up / down (reference :94-104): per axis, an up-sampling FIR of
    # t taps needs ((t + up - 1) // 2, (t - up) // 2) more, a down-sampling one ((t - down + 1) // 2, (t - down) // 2).
    for axis, t in enumerate(taps):
        lo, hi = 2 * axis, 2 * axis + 1
        if up > 1:
            pad[lo] += (t + up - 1) // 2
            pad[hi] += (t - up) // 2
        if down > 1:
            pad[lo] += (t - down + 1) // 2
            pad[hi] += (t - down) // 2
    # tcgen05 implicit-GEMM path (inference, dense fp32 NCHW): plain 'same' 1x1 / 3x3 convolutions, and the up-2 3x3 form
    # evaluated polyphase on the low-resolution input (no (2H+1)^2 intermediate, no separate FIR pass).
    if conv_igemm.supported(x, w, up=up, down=down, groups=groups, f=f, padding=_parse_padding(padding), flip_filter=flip_filter):
        return _igemm(x, w, groups, f=f, up=up, down=down, flip_weight=flip_weight)

    pointwise = (kw == 1 and kh == 1)
    fir = dict(f=f, flip_filter=flip_filter)
    conv = dict(groups=groups, flip_weight=flip_weight)

    if pointwise and down > 1 and up == 1:          # decimate first: 4x fewer pixels through the GEMM
        return _conv2d_wrapper(x=upfirdn2d.upfirdn2d(x=x, down=down, padding=pad, **fir), w=w, **conv)

    if pointwise and up > 1 and down == 1:          # conv at low resolution, then upsample
        return upfirdn2d.upfirdn2d(x=_conv2d_wrapper(x=x, w=w, **conv), up=up, padding=pad, gain=up ** 2, **fir)

    if down > 1 and up == 1:                        # low-pass at full resolution, strided conv
        return _conv2d_wrapper(x=upfirdn2d.upfirdn2d(x=x, padding=pad, **fir), w=w, stride=down, **conv)

    if up > 1:                                      # transposed strided conv, then low-pass (gain up^2), then the optional decimation
        if groups == 1:
            wt = w.transpose(0, 1)
        else:                                       # swap (out, in) inside every group
            wt = w.reshape(groups, cout // groups, cin_g, kh, kw).transpose(1, 2).reshape(groups * cin_g, cout // groups, kh, kw)
        # the transposed conv grows each axis by k - 1 on the left and k - up on the right; whatever part of that the FIR padding does not
        # want back is cropped by the transposed conv's own (symmetric) padding, the rest by the FIR stage
        rest = [pad[0] - (kw - 1), pad[1] - (kw - up), pad[2] - (kh - 1), pad[3] - (kh - up)]
        crop = [max(min(-rest[0], -rest[1]), 0), max(min(-rest[2], -rest[3]), 0)]          # (x, y)
        x = _conv2d_wrapper(x=x, w=wt, stride=up, padding=[crop[1], crop[0]], groups=groups, transpose=True, flip_weight=(not flip_weight))
        x = upfirdn2d.upfirdn2d(x=x, padding=[rest[0] + crop[0], rest[1] + crop[0], rest[2] + crop[1], rest[3] + crop[1]], gain=up ** 2, **fir)
        return upfirdn2d.upfirdn2d(x=x, down=down, **fir) if down > 1 else x

    if pad[0] == pad[1] and pad[2] == pad[3] and pad[0] >= 0 and pad[2] >= 0:               # plain convolution, symmetric padding
        return _conv2d_wrapper(x=x, w=w, padding=[pad[2], pad[0]], **conv)

    # Anything else (asymmetric / negative padding without resampling): explicit pad-or-crop, then conv.
    return _conv2d_wrapper(x=upfirdn2d.upfirdn2d(x=x, f=None, padding=pad, flip_filter=flip_filter), w=w, **conv)
