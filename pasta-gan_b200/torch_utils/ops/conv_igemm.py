"""tcgen05 / TMEM implicit-GEMM convolution (``pg_conv2d_igemm_fwd``) — the B200 replacement for the cuDNN calls behind
``conv2d_gradfix`` plus the modulation / demodulation / noise / bias_act passes the reference runs around them
(training/networks.py:37-94, :170-179, :296-315, :4342-4354).

Forward only: it is used when no gradient is required (inference); under autograd the callers keep the
``conv2d_gradfix`` route.  fp32 NCHW in / out, fp16 (default) or bf16 operands, fp32 accumulation in TMEM.
"""
import os

import torch

from . import _backend

_ACT = {'linear': 1, 'relu': 2, 'lrelu': 3}
_FMT = {'fp16': 0, 'bf16': 1}

enabled = os.environ.get('PASTA_B200_CONV', '1') != '0'
operand_format = os.environ.get('PASTA_B200_CONV_FMT', 'fp16')


def supported(x, w, up=1, down=1, groups=1, f=None, padding=None, flip_filter=False, x2=None, residual=None, allow_half=False):
    """Shapes the kernel covers: dense fp32 NCHW on CUDA, 1x1 / 3x3, stride 1, 'same' padding, optional polyphase up-2."""
    if not (enabled and x.is_cuda and x.dtype in (torch.float32, torch.float16) and w.dtype == torch.float32 and x.ndim == 4):
        return False
    if x.dtype == torch.float16 and not (allow_half and half_input_ok(x, w, up, down, x2)):
        return False                                     # fp16 activations of the reference's own fp16 blocks (discriminator) keep their path
    if torch.is_grad_enabled() and (x.requires_grad or w.requires_grad):
        return False
    k = int(w.shape[2])
    folded = up == 1 and down == 1 and k in (3, 5, 7) and int(w.shape[1]) * k * k <= 160     # small-Cin layers: taps folded into the GEMM K dimension
    if groups != 1 or down not in (1, 2) or up not in (1, 2) or w.shape[2] != w.shape[3] or (k not in (1, 3) and not folded):
        return False
    if down == 2 and (up != 1 or k != 3 or f is None or f.ndim != 2 or tuple(f.shape) != (4, 4) or flip_filter or
                      x.shape[1] % 16 != 0 or x.shape[2] % 2 or x.shape[3] % 2):
        return False
    if padding is not None and tuple(padding) != (k // 2,) * 4:
        return False
    if up == 2 and (k != 3 or f is None or f.ndim != 2 or tuple(f.shape) != (4, 4) or flip_filter or w.shape[0] % 16 != 0):
        return False
    if x.numel() == 0 or x.numel() > 2 ** 31 - 1:
        return False
    if x2 is not None and (down != 1 or x2.dtype != torch.float32 or x.shape[1] % 8 or x2.shape[0] != x.shape[0] or x2.shape[2:] != x.shape[2:] or
                           (torch.is_grad_enabled() and x2.requires_grad)):
        return False
    if residual is not None and (residual.dtype != torch.float32 or (torch.is_grad_enabled() and residual.requires_grad)):
        return False
    if down == 2 and x.data_ptr() % 8:
        return False
    return True


def half_input_ok(x, w, up=1, down=1, x2=None):
    """fp16 NCHW activations are accepted as operand bits by plain stride-1 1x1 / 3x3 layers with even W (include/pasta_b200.h); W <= 256 keeps the
    converter's task count inside the register-batched loader."""
    k = int(w.shape[2])
    return (operand_format == 'fp16' and up == 1 and down == 1 and x2 is None and k in (1, 3) and (k == 1 or int(w.shape[1]) * k * k > 160) and
            x.shape[3] % 2 == 0 and x.shape[3] <= 256 and x.data_ptr() % 4 == 0)


_pack_cache = {}          # (id(param), version, ...) -> (param, packed weights): inference packs each parameter once
_PACK_CACHE_MAX = 512


def _packed_weights(capi, w, f, w_scale, up, flip_weight, fmt_code, cache):
    """fp16/bf16 GEMM tiles of ``w * w_scale`` (pg_conv2d_igemm_prepack).  With ``cache=True`` the caller promises that ``w`` is a
    long-lived tensor (a Parameter / buffer): the packed copy is reused until the tensor's version counter changes."""
    cout, cin, k, _ = (int(v) for v in w.shape)
    key = None
    if cache:
        key = (id(w), w._version, w.data_ptr(), float(w_scale), up, bool(flip_weight), fmt_code)
        hit = _pack_cache.get(key)
        if hit is not None and hit[0] is w:
            return hit[1]
    lib = capi.load()
    ws_bytes = int(lib.pg_conv2d_igemm_workspace_bytes(cin, cout, k, up))
    if ws_bytes < 0:
        capi.check(2, 'pg_conv2d_igemm_workspace_bytes')
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=w.device)
    wc = w.detach().contiguous()
    rc = lib.pg_conv2d_igemm_prepack(capi.ptr(wc), capi.ptr(f) if up != 1 else None, float(w_scale), cin, cout, k, up,
                                     int(bool(flip_weight)), fmt_code, capi.ptr(ws), ws_bytes, capi.current_stream(w.device))
    capi.check(rc, 'pg_conv2d_igemm_prepack')
    if key is not None:
        if len(_pack_cache) >= _PACK_CACHE_MAX:
            _pack_cache.pop(next(iter(_pack_cache)))
        _pack_cache[key] = (w, ws)
    return ws


def conv2d_igemm(x, w, f=None, up=1, down=1, flip_weight=True, styles=None, dcoefs=None, noise=None, bias=None,
                 in_act='linear', in_alpha=0.2, in_gain=1.0, act='linear', alpha=0.2, gain=1.0, clamp=None, fmt=None,
                 w_scale=1.0, cache_weights=False, x2=None, residual=None, out_dtype=torch.float32):
    """y = clamp(act(dcoefs * conv(styles * in_gain * in_act([x ; x2]), w * w_scale) + noise + bias) * gain) + residual; see include/pasta_b200.h.
    ``x2``: second part of the input along channels (fused torch.cat); ``residual``: tensor of the output's shape added last."""
    capi = _backend.capi()
    _backend.require_cuda(x, 'conv2d_igemm')
    n, cin1, h, wd = (int(v) for v in x.shape)
    cin = cin1 + (int(x2.shape[1]) if x2 is not None else 0)
    cout, cin_w, k, _ = (int(v) for v in w.shape)
    assert cin_w == cin, 'weight / input channel mismatch'
    x = x.contiguous()
    if x2 is not None:
        assert x2.dtype == torch.float32 and tuple(x2.shape[2:]) == (h, wd) and int(x2.shape[0]) == n and down == 1 and cin1 % 8 == 0
        x2 = x2.contiguous()
    assert not (up == 2 and down == 2)
    mode = -2 if down == 2 else up                       # PG_CONV_DOWN2 / 2 / 1
    oh, ow = (h // 2, wd // 2) if down == 2 else (h * up, wd * up)
    assert out_dtype in (torch.float32, torch.float16) and x.dtype in (torch.float32, torch.float16)
    if x.dtype == torch.float16:
        assert styles is None and in_act == 'linear' and in_gain == 1.0 and half_input_ok(x, w, up, down, x2), 'float16 input: plain stride-1 layer only'
    y = torch.empty([n, cout, oh, ow], dtype=out_dtype, device=x.device)
    nb_stride = 0
    if noise is not None:
        noise = noise.to(torch.float32).contiguous()
        assert noise.shape[-2:] == (oh, ow), 'noise must match the output resolution'
        if noise.ndim == 4:
            assert noise.shape[1] == 1
            nb_stride = 0 if noise.shape[0] == 1 else oh * ow
        else:
            assert noise.ndim == 2
    opt = lambda t: None if t is None else t.to(torch.float32).contiguous()
    styles, dcoefs, bias = opt(styles), opt(dcoefs), opt(bias)
    if styles is not None:
        assert tuple(styles.shape) == (n, cin)
    if dcoefs is not None:
        assert tuple(dcoefs.shape) == (n, cout)
    if bias is not None:
        assert tuple(bias.shape) == (cout,)
    if mode != 1:
        f = f.to(torch.float32).contiguous()
    if residual is not None:
        assert residual.dtype == torch.float32 and residual.shape == y.shape
        residual = residual.contiguous()
    fmt_code = _FMT[fmt or operand_format]
    with torch.cuda.device(x.device):
        capi.require_device()
        wpack = _packed_weights(capi, w, f, w_scale, mode, flip_weight, fmt_code, cache_weights)
        # algorithmic FLOPs: output pixels for stride-1 / down-2, INPUT pixels for up-2 (zero-inserted taps excluded)
        sp = capi.span('conv_igemm', flops=2 * n * cout * cin * k * k * (oh * ow if up == 1 else h * wd),
                       nbytes=x.element_size() * x.numel() + 4 * ((x2.numel() if x2 is not None else 0) + (y.numel() if residual is not None else 0) + w.numel()) + y.element_size() * y.numel(),
                       tag=f'{cin}->{cout} @{h}x{wd} k{k} mode{mode}' + (' mod' if styles is not None else '') + (' cat' if x2 is not None else '') + (' res' if residual is not None else '') + (f' in_{in_act}' if in_act != 'linear' else ''))
        rc = capi.load().pg_conv2d_igemm_run2(capi.ptr(x), capi.ptr(x2), cin1, capi.ptr(wpack), capi.ptr(styles), capi.ptr(dcoefs),
                                              capi.ptr(noise), nb_stride, capi.ptr(bias), capi.ptr(residual), capi.ptr(y),
                                              n, cin, cout, h, wd, k, mode,
                                              _ACT[in_act], float(in_alpha), float(in_gain),
                                              _ACT[act], float(alpha), float(gain), float(-1 if clamp is None else clamp),
                                              fmt_code, capi.dtype_code(x.dtype), capi.dtype_code(out_dtype), capi.current_stream(x.device))
        capi.check(rc, 'pg_conv2d_igemm_run2')
        if sp:
            sp.close()
    return y


_cat_cache = {}


def _gamma_beta_weights(w_gamma, w_beta):
    """[w_gamma ; w_beta] as one long-lived tensor (so the packed-weight cache can key on it), rebuilt when either parameter changes."""
    key = (id(w_gamma), id(w_beta), w_gamma._version, w_beta._version, w_gamma.data_ptr(), w_beta.data_ptr())
    hit = _cat_cache.get(key)
    if hit is None or hit[0] is not w_gamma or hit[1] is not w_beta:
        if len(_cat_cache) > 64:
            _cat_cache.clear()
        hit = (w_gamma, w_beta, torch.cat([w_gamma.detach(), w_beta.detach()], dim=0).contiguous())
        _cat_cache[key] = hit
    return hit[2]


def spade_supported(x, feat, w_gamma, w_beta):
    if not (enabled and x.is_cuda and x.dtype == torch.float32 and feat.dtype in (torch.float32, torch.float16) and x.ndim == 4):
        return False
    if torch.is_grad_enabled() and (x.requires_grad or feat.requires_grad or w_gamma.requires_grad):
        return False
    c = int(x.shape[1])
    k = int(w_gamma.shape[2])
    if feat.dtype == torch.float16 and not half_input_ok(feat, w_gamma):
        return False
    return (w_gamma.shape == w_beta.shape and w_gamma.shape[0] == c and 2 * c <= 256 and c % 16 == 0 and k in (1, 3) and
            w_gamma.shape[2] == w_gamma.shape[3] and feat.shape[2:] == x.shape[2:] and feat.shape[1] == w_gamma.shape[1] and x.numel() > 0)


def instance_stats(x, eps=1e-5):
    """(mean, rstd) [N, C] of nn.InstanceNorm2d(affine=False) for x [N, C, H, W] fp32, one streaming pass (pg_instance_norm_stats)."""
    capi = _backend.capi()
    _backend.require_cuda(x, 'instance_stats')
    x = x.contiguous()
    n, c, h, wd = (int(v) for v in x.shape)
    mean = torch.empty([n, c], dtype=torch.float32, device=x.device)
    rstd = torch.empty_like(mean)
    with torch.cuda.device(x.device):
        capi.require_device()
        sp = capi.span('instance_stats', nbytes=4 * x.numel(), tag=f'{n * c} x {h * wd}')
        rc = capi.load().pg_instance_norm_stats(capi.ptr(x), capi.ptr(mean), capi.ptr(rstd), n * c, h * wd, float(eps), capi.current_stream(x.device))
        capi.check(rc, 'pg_instance_norm_stats')
        if sp:
            sp.close()
    return mean, rstd


def spade_conv_norm(x, feat, w_gamma, w_beta, w_scale=1.0, act='linear', alpha=0.2, gain=1.0, eps=1e-5, fmt=None, stats=None, out_dtype=torch.float32):
    """act(instance_norm(x) * (1 + conv(feat, w_gamma)) + conv(feat, w_beta)) * gain  in one tcgen05 launch (gamma / beta stay in TMEM)."""
    capi = _backend.capi()
    n, c, h, wd = (int(v) for v in x.shape)
    cin, k = int(feat.shape[1]), int(w_gamma.shape[2])
    x = x.contiguous()
    feat = feat.contiguous()
    mean, rstd = stats if stats is not None else instance_stats(x, eps)      # `stats`: reuse when several norm blocks share x
    wcat = _gamma_beta_weights(w_gamma, w_beta)
    fmt_code = _FMT[fmt or operand_format]
    y = torch.empty_like(x, dtype=out_dtype)
    with torch.cuda.device(x.device):
        capi.require_device()
        wpack = _packed_weights(capi, wcat, None, w_scale, 1, True, fmt_code, True)
        sp = capi.span('conv_igemm', flops=2 * n * 2 * c * cin * k * k * h * wd, nbytes=4 * (x.numel() + wcat.numel()) + feat.element_size() * feat.numel() + y.element_size() * y.numel(),
                       tag=f'{cin}->{2 * c} @{h}x{wd} k{k} spade' + (' in16' if feat.dtype == torch.float16 else '') + (' out16' if out_dtype == torch.float16 else ''))
        rc = capi.load().pg_conv2d_igemm_spade_run(capi.ptr(feat), capi.ptr(wpack), capi.ptr(x), capi.ptr(mean), capi.ptr(rstd), capi.ptr(y),
                                                   n, cin, c, h, wd, k, _ACT[act], float(alpha), float(gain), fmt_code,
                                                   capi.dtype_code(feat.dtype), capi.dtype_code(out_dtype), capi.current_stream(x.device))
        capi.check(rc, 'pg_conv2d_igemm_spade_run')
        if sp:
            sp.close()
    return y
