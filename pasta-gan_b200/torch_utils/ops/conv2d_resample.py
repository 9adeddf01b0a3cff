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
        return conv_igemm.conv2d_igemm(x, w, flip_weight=True)       # plain stride-1 'same' conv inside a lowering; w already carries the flip
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
    fw, fh = _get_filter_size(f)
    px0, px1, py0, py1 = _parse_padding(padding)

    # Padding owed to the resampling filters, so that sizes come out as H*up/down.
    if up > 1:
        px0 += (fw + up - 1) // 2
        px1 += (fw - up) // 2
        py0 += (fh + up - 1) // 2
        py1 += (fh - up) // 2
    if down > 1:
        px0 += (fw - down + 1) // 2
        px1 += (fw - down) // 2
        py0 += (fh - down + 1) // 2
        py1 += (fh - down) // 2
    # tcgen05 implicit-GEMM path (inference, dense fp32 NCHW): plain 'same' 1x1 / 3x3 convolutions, and the up-2 3x3 form
    # evaluated polyphase on the low-resolution input (no (2H+1)^2 intermediate, no separate FIR pass).
    if conv_igemm.supported(x, w, up=up, down=down, groups=groups, f=f, padding=_parse_padding(padding), flip_filter=flip_filter):
        return conv_igemm.conv2d_igemm(x, w, f=f, up=up, down=down, flip_weight=flip_weight)

    pad = [px0, px1, py0, py1]
    pointwise = (kw == 1 and kh == 1)

    if pointwise and down > 1 and up == 1:          # decimate first: 4x fewer pixels through the GEMM
        x = upfirdn2d.upfirdn2d(x=x, f=f, down=down, padding=pad, flip_filter=flip_filter)
        return _conv2d_wrapper(x=x, w=w, groups=groups, flip_weight=flip_weight)

    if pointwise and up > 1 and down == 1:          # conv at low resolution, then upsample
        x = _conv2d_wrapper(x=x, w=w, groups=groups, flip_weight=flip_weight)
        return upfirdn2d.upfirdn2d(x=x, f=f, up=up, padding=pad, gain=up ** 2, flip_filter=flip_filter)

    if down > 1 and up == 1:                        # low-pass at full resolution, strided conv
        x = upfirdn2d.upfirdn2d(x=x, f=f, padding=pad, flip_filter=flip_filter)
        return _conv2d_wrapper(x=x, w=w, stride=down, groups=groups, flip_weight=flip_weight)

    if up > 1:                                      # transposed strided conv, then low-pass (gain up^2)
        if groups == 1:
            wt = w.transpose(0, 1)
        else:
            wt = w.reshape(groups, cout // groups, cin_g, kh, kw).transpose(1, 2).reshape(groups * cin_g, cout // groups, kh, kw)
        px0 -= kw - 1
        px1 -= kw - up
        py0 -= kh - 1
        py1 -= kh - up
        pxt = max(min(-px0, -px1), 0)
        pyt = max(min(-py0, -py1), 0)
        x = _conv2d_wrapper(x=x, w=wt, stride=up, padding=[pyt, pxt], groups=groups, transpose=True, flip_weight=(not flip_weight))
        x = upfirdn2d.upfirdn2d(x=x, f=f, padding=[px0 + pxt, px1 + pxt, py0 + pyt, py1 + pyt], gain=up ** 2, flip_filter=flip_filter)
        if down > 1:
            x = upfirdn2d.upfirdn2d(x=x, f=f, down=down, flip_filter=flip_filter)
        return x

    if px0 == px1 and py0 == py1 and px0 >= 0 and py0 >= 0:      # plain convolution
        return _conv2d_wrapper(x=x, w=w, padding=[py0, px0], groups=groups, flip_weight=flip_weight)

    # Anything else (asymmetric / negative padding without resampling): explicit pad-or-crop, then conv.
    x = upfirdn2d.upfirdn2d(x=x, f=None, padding=pad, flip_filter=flip_filter)
    return _conv2d_wrapper(x=x, w=w, groups=groups, flip_weight=flip_weight)
