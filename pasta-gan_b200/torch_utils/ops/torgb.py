"""ToRGB skip path in one kernel (``pg_torgb_skip``): ``upsample2d(img) + clamp(modulated 1x1 conv(x) + b)``.

Extension of the reference surface (it composes upfirdn2d.upsample2d, modulated_conv2d, bias_act and add_ for this,
training/networks.py:5601-5611, :5709-5715).  Forward only: under autograd the callers keep the composed path."""
import torch

from . import _backend


def supported(x, weight, img=None, f=None):
    c8 = x.ndim == 5 and x.dtype == torch.float16 and x.shape[4] == 8           # channel-blocked fp16 [N, C/8, H, W, 8]
    if not (x.is_cuda and ((x.dtype == torch.float32 and x.ndim == 4) or c8) and weight.ndim == 4):
        return False
    if torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad):
        return False
    o, c, kh, kw = weight.shape
    h, w = (int(x.shape[2]), int(x.shape[3]))
    if kh != 1 or kw != 1 or o > 8 or w % 4 != 0 or x.numel() == 0 or (c8 and int(x.shape[1]) * 8 != c):
        return False
    if img is not None and (f is None or tuple(f.shape) != (4, 4) or h % 2 or img.shape[1] != o or
                            img.shape[2] * 2 != h or img.shape[3] * 2 != w or img.dtype != torch.float32):
        return False
    return True


def torgb_skip(x, weight, styles=None, bias=None, clamp=None, img=None, f=None):
    """x [N,C,H,W]; weight [O,C,1,1]; styles [N,C] (weight_gain already applied); img [N,O,H/2,W/2] or None."""
    capi = _backend.capi()
    _backend.require_cuda(x, 'torgb_skip')
    c8 = x.ndim == 5
    if c8:
        n, cb, h, w, _ = (int(v) for v in x.shape)
        c = cb * 8
    else:
        n, c, h, w = (int(v) for v in x.shape)
    o = int(weight.shape[0])
    x = x.contiguous()
    wt = weight.detach().reshape(o, c).to(torch.float32).contiguous()
    opt = lambda t: None if t is None else t.to(torch.float32).contiguous()
    styles, bias, img, f = opt(styles), opt(bias), opt(img), opt(f)
    out = torch.empty([n, o, h, w], dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        capi.require_device()
        sp = capi.span('torgb_skip', nbytes=x.element_size() * x.numel() + 4 * (out.numel() + (img.numel() if img is not None else 0)), tag='c8' if c8 else '')
        fn = capi.load().pg_torgb_skip_c8 if c8 else capi.load().pg_torgb_skip
        rc = fn(capi.ptr(x), capi.ptr(wt), capi.ptr(styles), capi.ptr(bias), capi.ptr(img), capi.ptr(f), capi.ptr(out),
                                       n, c, o, h, w, float(-1 if clamp is None else clamp), capi.current_stream(x.device))
        capi.check(rc, 'pg_torgb_skip')
        if sp:
            sp.close()
    return out
