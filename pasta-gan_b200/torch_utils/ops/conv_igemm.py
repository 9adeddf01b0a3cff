"""tcgen05 / TMEM implicit-GEMM convolution (``pg_conv2d_igemm_launch``) — the B200 replacement for the cuDNN calls behind
``conv2d_gradfix`` plus the modulation / demodulation / noise / bias_act passes the reference runs around them
(training/networks.py:37-94, :170-179, :296-315, :4342-4354).

Forward only: it is used when no gradient is required (inference); under autograd the callers keep the
``conv2d_gradfix`` route.  Activations: fp32 NCHW at the API boundary; between two of our own layers they may travel as fp16 NCHW or as
channel-blocked fp16 ("C8", a 5-D tensor ``[N, C/8, H, W, 8]``) which the kernel loads by TMA.  fp16 (default) or bf16 operands, fp32
accumulation in TMEM.
"""
import os
import weakref

import torch

from . import _backend

_ACT = {'linear': 1, 'relu': 2, 'lrelu': 3}
_FMT = {'fp16': 0, 'bf16': 1, 'tf32': 2}

enabled = os.environ.get('PASTA_B200_CONV', '1') != '0'
operand_format = os.environ.get('PASTA_B200_CONV_FMT', 'fp16')


def is_c8(x):
    """Channel-blocked fp16 activation: [N, C/8, H, W, 8]."""
    return x.ndim == 5 and x.dtype == torch.float16 and x.shape[4] == 8


def supported(x, w, up=1, down=1, groups=1, f=None, padding=None, flip_filter=False, x2=None, residual=None, allow_half=False):
    """Shapes the kernel covers: dense fp32 NCHW on CUDA, 1x1 / 3x3, stride 1, 'same' padding, optional polyphase up-2 / fused down-2; with
    ``groups = N`` the reference's fused modulated convolution x [1, N*I, H, W] * w [N*O, I, k, k] (training/networks.py:88-90)."""
    if not (enabled and x.is_cuda and x.dtype in (torch.float32, torch.float16) and w.dtype == torch.float32 and x.ndim == 4):
        return False
    if x.dtype == torch.float16 and not (allow_half and half_input_ok(x, w, up, down, x2)):
        return False                                     # fp16 activations of the reference's own fp16 blocks (discriminator) keep their path
    if torch.is_grad_enabled() and (x.requires_grad or w.requires_grad):
        return False
    k = int(w.shape[2])
    cin_g, cout = int(w.shape[1]), int(w.shape[0])
    if groups != 1:
        # per-sample weights: one "sample" per group
        if not (x.shape[0] == 1 and x.shape[1] == groups * cin_g and cout % groups == 0 and x2 is None and residual is None and down == 1 and
                groups <= 65535 and x.dtype == torch.float32):
            return False
        cout //= groups
    folded = up == 1 and down == 1 and k in (3, 5, 7) and cin_g * k * k <= 160     # small-Cin layers: taps folded into the GEMM K dimension
    if down not in (1, 2) or up not in (1, 2) or w.shape[2] != w.shape[3] or (k not in (1, 3) and not folded):
        return False
    if down == 2 and (up != 1 or k != 3 or f is None or f.ndim != 2 or tuple(f.shape) != (4, 4) or flip_filter or
                      x.shape[1] % 16 != 0 or x.shape[2] % 2 or x.shape[3] % 2):
        return False
    if padding is not None and tuple(padding) != (k // 2,) * 4:
        return False
    if up == 2 and (k != 3 or f is None or f.ndim != 2 or tuple(f.shape) != (4, 4) or flip_filter or cout % 16 != 0):
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


_DOWN2_TMA_MIN = 256        # smallest input height for which channel-blocked down-2 layers are routed through the strided TMA box (tests lower it)


def c8_input_ok(c, h, wd, k, up=1, down=1):
    """A channel-blocked input is loaded by TMA: plain stride-1 (or up-2) 1x1 / 3x3 layer, fp16 operands, whole 16-channel chunks; 3x3 layers wider than
    127 columns run in 64-column bands (even W); 1x1 layers wider than 128 columns are viewed as rows of 128 pixels (W % 128 == 0)."""
    if down == 2:
        # down-2 3x3: the space-to-depth planes come through a strided TMA box; the GEMM runs at the output resolution (bands above 127 columns)
        # (measured: wins from 256-px inputs up -- 64->128 @512^2 584 -> 480 us; at 128 px and below the converter path's per-sample CTAs are faster)
        return (enabled and operand_format == 'fp16' and up == 1 and k in (1, 3) and c % 16 == 0 and h % 2 == 0 and wd % 2 == 0 and h >= _DOWN2_TMA_MIN and
                (wd // 2 <= 127 or (wd // 2) % 2 == 0) and os.environ.get('PASTA_B200_CONV_TMA', '1') != '0')
    return (enabled and operand_format == 'fp16' and down == 1 and up in (1, 2) and k in (1, 3) and c % 16 == 0 and (k == 1 or c * k * k > 160) and
            ((wd <= 128 or wd % 128 == 0) if k == 1 else (wd <= 127 or wd % 2 == 0)) and
            os.environ.get('PASTA_B200_CONV_TMA', '1') != '0')


def to_c8(x):
    """dense NCHW fp32 / fp16 -> channel-blocked fp16 [N, ceil(C/8), H, W, 8] (pg_nchw_to_c8)."""
    capi = _backend.capi()
    _backend.require_cuda(x, 'to_c8')
    x = x.contiguous()
    n, c, h, wd = (int(v) for v in x.shape)
    y = torch.empty([n, (c + 7) // 8, h, wd, 8], dtype=torch.float16, device=x.device)
    with torch.cuda.device(x.device):
        capi.require_device()
        sp = capi.span('layout', nbytes=x.element_size() * x.numel() + 2 * y.numel(), tag='nchw->c8')
        capi.check(capi.load().pg_nchw_to_c8(capi.ptr(x), capi.ptr(y), n, c, h * wd, capi.dtype_code(x.dtype), capi.current_stream(x.device)), 'pg_nchw_to_c8')
        if sp:
            sp.close()
    return y


def from_c8(x, channels=None, dtype=torch.float32):
    """channel-blocked fp16 -> dense NCHW (pg_c8_to_nchw); ``channels``: the logical channel count (default: all stored)."""
    capi = _backend.capi()
    assert is_c8(x)
    x = x.contiguous()
    n, cb, h, wd, _ = (int(v) for v in x.shape)
    c = cb * 8 if channels is None else int(channels)
    assert (c + 7) // 8 == cb
    y = torch.empty([n, c, h, wd], dtype=dtype, device=x.device)
    with torch.cuda.device(x.device):
        capi.require_device()
        sp = capi.span('layout', nbytes=2 * x.numel() + y.element_size() * y.numel(), tag='c8->nchw')
        capi.check(capi.load().pg_c8_to_nchw(capi.ptr(x), capi.ptr(y), n, c, h * wd, capi.dtype_code(dtype), capi.current_stream(x.device)), 'pg_c8_to_nchw')
        if sp:
            sp.close()
    return y


# ---------------------------------------------------------------------------------------------------------------- packed-weight store
# One entry per (parameter, packing configuration).  An entry owns ONE workspace buffer for its whole life: when the parameter's version counter
# moves (optimizer step, EMA update, load_state_dict) the new values are packed INTO THE SAME STORAGE, so
#   * stale versions never accumulate (the round-1 cache kept every version until a FIFO eviction),
#   * a CUDA graph that baked the buffer's address keeps reading live weights after `refresh()` (TryOnSession.refresh_weights),
#   * nothing a captured graph points to is ever freed while the parameter lives: entries die with their parameter (weakref finalizer).
# The key carries the FIR tensor's identity for the resampling composites, and consumers on another stream wait on the pack's event.


class _PackEntry:
    __slots__ = ('wref', 'version', 'ptr', 'f', 'f_version', 'ws', 'event', 'stream', 'cfg', 'synced')


_pack_store = {}          # (id(w), cfg) -> _PackEntry


def _pack_cfg(w_scale, mode, flip_weight, fmt_code, n_tile=0):
    return (float(w_scale), int(mode), bool(flip_weight), int(fmt_code), int(n_tile))


def _drop_entries(wid):
    for key in [k for k in _pack_store if k[0] == wid]:
        _pack_store.pop(key, None)


_center_cache = {}        # id(w1x1) -> dict(ref, version, ptr, w3): 1x1 weights embedded as the centre tap of a 3x3 kernel


def _center_3x3(w):
    """A 1x1 down-2 layer (FIR + stride-2 1x1, the skip branch of a down-sampling res-block) is the 3x3 down-2 form with only the centre tap set:
    conv2d_resample(down=2) of a 3x3 kernel filters with pad 2 and strides from offset 0, the 1x1 form filters with pad 1, and the centre tap sits one
    sample in.  The embedded tensor is kept per parameter and refreshed in place when the parameter moves, so the packed-weight store and captured
    CUDA graphs follow weight updates (refresh_packed_weights)."""
    hit = _center_cache.get(id(w))
    if hit is not None and hit['ref']() is not w:
        _center_cache.pop(id(w), None)
        hit = None
    if hit is None:
        wid = id(w)
        hit = dict(ref=weakref.ref(w, lambda _r, wid=wid: _center_cache.pop(wid, None)), version=None, ptr=None,
                   w3=torch.zeros([int(w.shape[0]), int(w.shape[1]), 3, 3], dtype=torch.float32, device=w.device))
        _center_cache[id(w)] = hit
    if hit['version'] != w._version or hit['ptr'] != w.data_ptr():
        with torch.no_grad():
            hit['w3'][:, :, 1, 1].copy_(w.detach()[:, :, 0, 0])
        hit['version'], hit['ptr'] = w._version, w.data_ptr()
    return hit['w3']


def _run_prepack(capi, w, f, w_scale, mode, flip_weight, fmt_code, ws, batch=1, w_batch_stride=0, styles=None, n_tile=0):
    cout, cin, k = int(w.shape[-4]), int(w.shape[-3]), int(w.shape[-1])
    wc = w.detach().contiguous()
    rc = capi.load().pg_conv2d_igemm_prepack_batched(capi.ptr(wc), int(w_batch_stride), capi.ptr(styles), int(batch),
                                                     capi.ptr(f) if mode != 1 else None, float(w_scale), cin, cout, k, mode,
                                                     int(bool(flip_weight)), fmt_code, int(n_tile), capi.ptr(ws), int(ws.numel()), capi.current_stream(w.device))
    capi.check(rc, 'pg_conv2d_igemm_prepack')


def _workspace_bytes(capi, cin, cout, k, mode, fmt_code=0):
    n = int(capi.load().pg_conv2d_igemm_workspace_bytes_fmt(cin, cout, k, mode, fmt_code))
    if n < 0:
        capi.check(2, 'pg_conv2d_igemm_workspace_bytes')
    return n


def _packed_weights(capi, w, f, w_scale, mode, flip_weight, fmt_code, cache, n_tile=0):
    """fp16/bf16 GEMM tiles of ``w * w_scale`` (pg_conv2d_igemm_prepack).  With ``cache=True`` the caller promises that ``w`` is a
    long-lived tensor (a Parameter / buffer): the packed copy lives in the store above and is refreshed in place when ``w`` changes."""
    cout, cin, k, _ = (int(v) for v in w.shape)
    if not cache:
        ws = torch.empty(_workspace_bytes(capi, cin, cout, k, mode, fmt_code), dtype=torch.uint8, device=w.device)
        _run_prepack(capi, w, f, w_scale, mode, flip_weight, fmt_code, ws, n_tile=n_tile)
        return ws
    key = (id(w), _pack_cfg(w_scale, mode, flip_weight, fmt_code, n_tile))
    e = _pack_store.get(key)
    if e is not None and e.wref() is not w:               # id() reuse after the old tensor died without its finalizer having run yet
        _pack_store.pop(key, None)
        e = None
    cur = torch.cuda.current_stream(w.device)
    fresh = e is None
    if fresh:
        e = _PackEntry()
        wid = id(w)
        e.wref = weakref.ref(w, lambda _r, wid=wid: _drop_entries(wid))
        e.ws = torch.empty(_workspace_bytes(capi, cin, cout, k, mode, fmt_code), dtype=torch.uint8, device=w.device)
        e.cfg = (mode, flip_weight, fmt_code, w_scale, n_tile)
        e.version = None
        _pack_store[key] = e
    f_ver = None if mode == 1 else (f.data_ptr(), f._version)
    if fresh or e.version != w._version or e.ptr != w.data_ptr() or (mode != 1 and (e.f is not f or e.f_version != f_ver)):
        _run_prepack(capi, w, f, w_scale, mode, flip_weight, fmt_code, e.ws, n_tile=n_tile)
        e.version, e.ptr = w._version, w.data_ptr()
        e.f, e.f_version = (None, None) if mode == 1 else (f, f_ver)
        e.stream, e.synced = cur, {cur.cuda_stream}
        if not torch.cuda.is_current_stream_capturing():
            e.event = torch.cuda.Event()
            e.event.record(cur)
        else:
            e.event = None
    elif cur.cuda_stream not in e.synced and e.event is not None and not torch.cuda.is_current_stream_capturing():
        # packed on another stream: order this consumer (and everything later on its stream) after the pack -- once per stream.  Not while capturing:
        # a graph is always preceded by an eager warm-up on the same stream, which has done the wait
        cur.wait_event(e.event)
        e.synced.add(cur.cuda_stream)
    return e.ws


def refresh_packed_weights(device=None):
    """Re-pack, in place, every stored parameter whose version moved (after load_state_dict / an optimizer step).  Buffers keep their addresses, so
    CUDA graphs captured earlier read the new weights on their next replay.  Returns the number of re-packed entries."""
    capi = _backend.capi()
    done = 0
    if device is not None:
        device = torch.device(device)
        if device.type == 'cuda' and device.index is None:
            device = torch.device('cuda', torch.cuda.current_device())
    for hit in _cat_cache.values():                       # [gamma ; beta] concatenations first: rebuilt in place, their packed copies follow below
        done += _refresh_cat(hit)
    for hit in list(_center_cache.values()):              # centre-tap embeddings of 1x1 down-2 weights likewise
        w1 = hit['ref']()
        if w1 is not None:
            _center_3x3(w1)
    for (wid, _cfg), e in list(_pack_store.items()):
        w = e.wref()
        if w is None or (device is not None and w.device != device):
            continue
        f_moved = e.f is not None and e.f_version != (e.f.data_ptr(), e.f._version)
        if e.version != w._version or e.ptr != w.data_ptr() or f_moved:
            mode, flip_weight, fmt_code, w_scale, n_tile = e.cfg
            with torch.cuda.device(w.device):
                _run_prepack(capi, w, e.f, w_scale, mode, flip_weight, fmt_code, e.ws, n_tile=n_tile)
            e.version, e.ptr = w._version, w.data_ptr()
            if e.f is not None:
                e.f_version = (e.f.data_ptr(), e.f._version)
            e.stream = torch.cuda.current_stream(w.device)
            e.synced = {e.stream.cuda_stream}
            e.event = torch.cuda.Event()
            e.event.record(e.stream)
            done += 1
    return done


def packed_weight_buffers():
    """Every live packed-weight buffer (for owners of captured CUDA graphs that want to hold strong references)."""
    return [e.ws for e in _pack_store.values()] + [h['cat'] for h in _cat_cache.values()]


def _auto_n_tile(n, cout, oh, ow, k, up):
    """Latency-bound layers (4^2 .. 16^2, 512 channels): with the default 256-column N tile only N * tiles * ceil(Cout / 256) CTAs exist (32 at 4^2) and
    each streams ~2.4 MB of weights through one SM.  A narrower N tile spreads the weight stream over all SMs.  0 = the library default."""
    if up != 1 or cout < 128 or _NT_AUTO == '0':
        return 0
    tiles = (oh * (ow + 1 if k == 3 else ow) + 127) // 128
    if n * tiles * ((cout + 255) // 256) >= 148:
        return 0
    for bn in (128, 64, 32):
        if cout % bn == 0 and n * tiles * (cout // bn) >= 148:
            return bn
    return 32 if cout % 32 == 0 else 0


_NT_AUTO = os.environ.get('PASTA_B200_CONV_AUTO_NTILE', '1')


def conv2d_igemm(x, w, f=None, up=1, down=1, flip_weight=True, styles=None, dcoefs=None, noise=None, bias=None,
                 in_act='linear', in_alpha=0.2, in_gain=1.0, act='linear', alpha=0.2, gain=1.0, clamp=None, fmt=None,
                 w_scale=1.0, cache_weights=False, x2=None, residual=None, out_dtype=torch.float32, out_c8=False,
                 per_sample_weights=False, fold_styles=False, styles_normalized=False):
    """y = clamp(act(dcoefs * conv(styles * in_gain * in_act([x ; x2]), w * w_scale) + noise + bias) * gain) + residual; see include/pasta_b200.h.
    ``x2``: second part of the input along channels (fused torch.cat); ``residual``: tensor of the output's shape added last.
    ``x`` may be channel-blocked fp16 (``is_c8``): TMA operand path; ``out_c8``: write the result channel-blocked.
    ``per_sample_weights``: ``w`` is [N, O, I, k, k] — the groups = N form of the reference's fused modulated conv.
    ``fold_styles``: multiply the styles into per-sample packed weights instead of the activations (needed when x is channel-blocked)."""
    capi = _backend.capi()
    _backend.require_cuda(x, 'conv2d_igemm')
    x_c8 = is_c8(x)
    if x_c8:
        n, cb_in, h, wd, _ = (int(v) for v in x.shape)
        cin1 = cb_in * 8
        assert x2 is None or (is_c8(x2) and cin1 % 16 == 0 and tuple(x2.shape[2:4]) == (h, wd) and int(x2.shape[0]) == n), 'a channel-blocked x needs a channel-blocked x2'
        cin = cin1 + (int(x2.shape[1]) * 8 if x2 is not None else 0)
    else:
        n, cin1, h, wd = (int(v) for v in x.shape)
        cin = cin1 + (int(x2.shape[1]) if x2 is not None else 0)
    if per_sample_weights:
        assert w.ndim == 5 and int(w.shape[0]) == n and not cache_weights and styles is None
        cout, cin_w, k = int(w.shape[1]), int(w.shape[2]), int(w.shape[4])
    else:
        cout, cin_w, k, _ = (int(v) for v in w.shape)
        if x_c8 and down == 2 and k == 1:                 # skip branch of a down-sampling res-block on the strided-TMA path
            w, k = _center_3x3(w), 3
    assert cin_w == cin, 'weight / input channel mismatch'
    x = x.contiguous()
    if x2 is not None and not x_c8:
        assert x2.dtype == torch.float32 and tuple(x2.shape[2:]) == (h, wd) and int(x2.shape[0]) == n and down == 1 and cin1 % 8 == 0
    if x2 is not None:
        x2 = x2.contiguous()
    assert not (up == 2 and down == 2)
    mode = (-3 if x_c8 else -2) if down == 2 else up     # PG_CONV_DOWN2_C8 / PG_CONV_DOWN2 / 2 / 1
    oh, ow = (h // 2, wd // 2) if down == 2 else (h * up, wd * up)
    assert out_dtype in (torch.float32, torch.float16) and x.dtype in (torch.float32, torch.float16)
    if x_c8:
        assert in_act == 'linear' and in_gain == 1.0 and (styles is None or fold_styles) and c8_input_ok(cin, h, wd, k, up, down), \
            'channel-blocked input: plain (or style-folded) stride-1 / up-2 layer only'
    elif x.dtype == torch.float16:
        assert styles is None and in_act == 'linear' and in_gain == 1.0 and half_input_ok(x, w, up, down, x2), 'float16 input: plain stride-1 layer only'
    if out_c8:
        assert cout % 16 == 0 and (up == 1 or cout <= 128) and (residual is None or is_c8(residual))
        y = torch.empty([n, cout // 8, oh, ow, 8], dtype=torch.float16, device=x.device)
    else:
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
        if not styles_normalized:
            # fp16 operands: keep x * s inside the fp16 range whatever the style magnitude (the reference pre-normalises its own fp16 path the same
            # way, training/networks.py:57-59): scale each sample's styles to unit inf-norm and hand the factor to the epilogue's per-sample coefficient
            smax = styles.abs().amax(dim=1, keepdim=True).clamp_min(1e-20)
            styles = styles / smax
            dcoefs = smax.expand(n, cout).contiguous() if dcoefs is None else dcoefs * smax
    if dcoefs is not None:
        assert tuple(dcoefs.shape) == (n, cout)
    if bias is not None:
        assert tuple(bias.shape) == (cout,)
    if mode != 1:
        f = f.to(torch.float32).contiguous()
    res_c8 = residual is not None and is_c8(residual)
    if residual is not None:
        if res_c8:
            assert tuple(residual.shape) == (n, cout // 8, oh, ow, 8) and cout % 16 == 0 and up == 1
        else:
            assert residual.dtype == torch.float32 and residual.shape == y.shape
        residual = residual.contiguous()
    fmt_code = _FMT[fmt or operand_format]
    if fmt_code == 2:
        assert not x_c8 and x.dtype == torch.float32, 'tf32 operands are converted from dense float32 activations (no channel-blocked / float16 input)'
    n_tile = _auto_n_tile(n, cout, oh, ow, k, up)
    with torch.cuda.device(x.device):
        capi.require_device()
        sample_stride = 0
        if per_sample_weights or (fold_styles and styles is not None):
            # one packed weight set per sample, built per call (the weights are a function of this batch's styles)
            per = _workspace_bytes(capi, cin, cout, k, mode, fmt_code)
            wpack = torch.empty(per * n, dtype=torch.uint8, device=x.device)
            if per_sample_weights:
                _run_prepack(capi, w, f, w_scale, mode, flip_weight, fmt_code, wpack, batch=n, w_batch_stride=cout * cin * k * k, n_tile=n_tile)
            else:
                _run_prepack(capi, w, f, w_scale, mode, flip_weight, fmt_code, wpack, batch=n, w_batch_stride=0, styles=styles, n_tile=n_tile)
                styles = None
            sample_stride = per
        else:
            wpack = _packed_weights(capi, w, f, w_scale, mode, flip_weight, fmt_code, cache_weights, n_tile=n_tile)
        # algorithmic FLOPs: output pixels for stride-1 / down-2, INPUT pixels for up-2 (zero-inserted taps excluded)
        sp = capi.span('conv_igemm', flops=2 * n * cout * cin * k * k * (oh * ow if up == 1 else h * wd),
                       nbytes=x.element_size() * x.numel() + (x2.element_size() * x2.numel() if x2 is not None else 0) + (residual.element_size() * residual.numel() if residual is not None else 0) + 4 * w.numel() + y.element_size() * y.numel(),
                       tag=f'{cin}->{cout} @{h}x{wd} k{k} mode{mode}' + (' mod' if styles is not None else '') + (' cat' if x2 is not None else '') + (' res' if residual is not None else '') +
                           (f' in_{in_act}' if in_act != 'linear' else '') + (' tma' if x_c8 else '') + (' psw' if sample_stride else '') + (' oc8' if out_c8 else ''))
        a = capi.ConvArgs()
        a.struct_bytes = _ARGS_BYTES
        a.N, a.Cin, a.Cout, a.H, a.W, a.ksize, a.up = n, cin, cout, h, wd, k, mode
        a.x, a.x_dtype, a.x_layout = capi.ptr(x), capi.dtype_code(x.dtype), capi.LAYOUT_C8 if x_c8 else capi.LAYOUT_NCHW
        a.x2, a.cin1, a.residual_layout = capi.ptr(x2), cin1, (capi.LAYOUT_C8 if res_c8 else capi.LAYOUT_NCHW)
        a.wpack, a.wpack_sample_stride = capi.ptr(wpack), sample_stride
        a.styles, a.dcoefs, a.noise, a.noise_batch_stride, a.bias, a.residual = capi.ptr(styles), capi.ptr(dcoefs), capi.ptr(noise), nb_stride, capi.ptr(bias), capi.ptr(residual)
        a.y, a.y_dtype, a.y_layout = capi.ptr(y), capi.dtype_code(y.dtype), capi.LAYOUT_C8 if out_c8 else capi.LAYOUT_NCHW
        a.in_act, a.in_alpha, a.in_gain = _ACT[in_act], float(in_alpha), float(in_gain)
        a.act, a.alpha, a.gain, a.clamp = _ACT[act], float(alpha), float(gain), float(-1 if clamp is None else clamp)
        a.operand_format, a.n_tile = fmt_code, n_tile
        a.stream = capi.current_stream(x.device)
        rc = capi.load().pg_conv2d_igemm_launch(_byref(a))
        capi.check(rc, 'pg_conv2d_igemm_launch')
        if sp:
            sp.close()
    return y


import ctypes as _ctypes  # noqa: E402

_byref = _ctypes.byref
_ARGS_BYTES = _ctypes.sizeof(_backend.capi().ConvArgs)

_cat_cache = {}           # (id(w_gamma), id(w_beta)) -> dict(refs, versions, cat)


def _refresh_cat(hit):
    wg, wb = hit['g'](), hit['b']()
    if wg is None or wb is None:
        return 0
    ver = (wg._version, wb._version, wg.data_ptr(), wb.data_ptr())
    if ver == hit['ver']:
        return 0
    with torch.no_grad():
        c, ct = int(wg.shape[0]), hit['ct']
        cat = hit['cat'].view(c // ct, 2, ct, *wg.shape[1:])            # tile t: [gamma[t*ct:(t+1)*ct] ; beta[t*ct:(t+1)*ct]]
        cat[:, 0].copy_(wg.detach().view(c // ct, ct, *wg.shape[1:]))
        cat[:, 1].copy_(wb.detach().view(c // ct, ct, *wb.shape[1:]))   # in place: the tensor (and its packed copy's storage) keep their addresses
    hit['ver'] = ver
    return 1


def _gamma_beta_weights(w_gamma, w_beta, ct=None):
    """[w_gamma ; w_beta] as one long-lived tensor (so the packed-weight store can key on it), rebuilt IN PLACE when either parameter changes.
    ``ct``: channels per N tile -- rows are ordered tile by tile, [gamma_t ; beta_t] (default: one tile)."""
    ct = int(w_gamma.shape[0]) if ct is None else int(ct)
    key = (id(w_gamma), id(w_beta), ct)
    hit = _cat_cache.get(key)
    if hit is not None and (hit['g']() is not w_gamma or hit['b']() is not w_beta):
        _cat_cache.pop(key, None)
        hit = None
    if hit is None:
        drop = lambda _r, key=key: _cat_cache.pop(key, None)
        hit = dict(g=weakref.ref(w_gamma, drop), b=weakref.ref(w_beta, drop), ver=None, ct=ct,
                   cat=torch.empty([2 * int(w_gamma.shape[0])] + list(w_gamma.shape[1:]), dtype=torch.float32, device=w_gamma.device))
        _cat_cache[key] = hit
    _refresh_cat(hit)
    return hit['cat']


def spade_supported(x, feat, w_gamma, w_beta):
    if not (enabled and x.is_cuda and x.dtype == torch.float32 and feat.dtype in (torch.float32, torch.float16) and x.ndim == 4):
        return False
    if torch.is_grad_enabled() and (x.requires_grad or feat.requires_grad or w_gamma.requires_grad):
        return False
    c = int(x.shape[1])
    k = int(w_gamma.shape[2])
    if is_c8(feat):
        if not (c8_input_ok(int(feat.shape[1]) * 8, int(feat.shape[2]), int(feat.shape[3]), k) and tuple(feat.shape[2:4]) == tuple(x.shape[2:]) and
                int(feat.shape[1]) * 8 == int(w_gamma.shape[1])):
            return False
    else:
        if feat.ndim != 4 or (feat.dtype == torch.float16 and not half_input_ok(feat, w_gamma)):
            return False
        if not (feat.shape[2:] == x.shape[2:] and feat.shape[1] == w_gamma.shape[1]):
            return False
    return (w_gamma.shape == w_beta.shape and w_gamma.shape[0] == c and 2 * c <= 256 and c % 16 == 0 and k in (1, 3) and
            w_gamma.shape[2] == w_gamma.shape[3] and x.numel() > 0)


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


def instance_norm_act(x, act='lrelu', alpha=0.01, gain=1.0, eps=1e-5):
    """act(InstanceNorm2d(affine=False)(x)) * gain in one pass (pg_instance_norm_act); planes of at most 24576 elements."""
    capi = _backend.capi()
    _backend.require_cuda(x, 'instance_norm_act')
    x = x.contiguous()
    n, c, h, wd = (int(v) for v in x.shape)
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        capi.require_device()
        sp = capi.span('instance_norm_act', nbytes=8 * x.numel(), tag=f'{n * c} x {h * wd}')
        rc = capi.load().pg_instance_norm_act(capi.ptr(x), capi.ptr(y), n * c, h * wd, float(eps), _ACT[act], float(alpha), float(gain), capi.current_stream(x.device))
        capi.check(rc, 'pg_instance_norm_act')
        if sp:
            sp.close()
    return y


def spade_conv_norm(x, feat, w_gamma, w_beta, w_scale=1.0, act='linear', alpha=0.2, gain=1.0, eps=1e-5, fmt=None, stats=None, out_dtype=torch.float32,
                    out_c8=False):
    """act(instance_norm(x) * (1 + conv(feat, w_gamma)) + conv(feat, w_beta)) * gain  in one tcgen05 launch (gamma / beta stay in TMEM).
    ``feat`` may be channel-blocked fp16 (TMA operand path); ``out_c8``: channel-blocked fp16 result."""
    capi = _backend.capi()
    n, c, h, wd = (int(v) for v in x.shape)
    f_c8 = is_c8(feat)
    cin, k = (int(feat.shape[1]) * 8 if f_c8 else int(feat.shape[1])), int(w_gamma.shape[2])
    x = x.contiguous()
    feat = feat.contiguous()
    mean, rstd = stats if stats is not None else instance_stats(x, eps)      # `stats`: reuse when several norm blocks share x
    # 2C = 256 output columns with a TMA-fed A operand: two N tiles [gamma_t | beta_t] of 128 columns, so that two CTAs share an SM and one's epilogue
    # overlaps the other's main loop (the A boxes are re-read from L2 by TMA, no converter work is duplicated)
    n_tile = 128 if (f_c8 and 2 * c == 256 and os.environ.get('PASTA_B200_SPADE_TILE', '128') == '128') else 0
    wcat = _gamma_beta_weights(w_gamma, w_beta, ct=(n_tile // 2 if n_tile else None))
    fmt_code = _FMT[fmt or operand_format]
    y = torch.empty([n, c // 8, h, wd, 8], dtype=torch.float16, device=x.device) if out_c8 else torch.empty_like(x, dtype=out_dtype)
    with torch.cuda.device(x.device):
        capi.require_device()
        wpack = _packed_weights(capi, wcat, None, w_scale, 1, True, fmt_code, True, n_tile=n_tile)
        sp = capi.span('conv_igemm', flops=2 * n * 2 * c * cin * k * k * h * wd, nbytes=4 * (x.numel() + wcat.numel()) + feat.element_size() * feat.numel() + y.element_size() * y.numel(),
                       tag=f'{cin}->{2 * c} @{h}x{wd} k{k} spade' + (' tma' if f_c8 else (' in16' if feat.dtype == torch.float16 else '')) +
                           (' oc8' if out_c8 else (' out16' if out_dtype == torch.float16 else '')))
        a = capi.ConvArgs()
        a.struct_bytes = _ARGS_BYTES
        a.N, a.Cin, a.Cout, a.H, a.W, a.ksize, a.up = n, cin, 2 * c, h, wd, k, 1
        a.x, a.x_dtype, a.x_layout = capi.ptr(feat), capi.dtype_code(feat.dtype), capi.LAYOUT_C8 if f_c8 else capi.LAYOUT_NCHW
        a.wpack = capi.ptr(wpack)
        a.y, a.y_dtype, a.y_layout = capi.ptr(y), capi.dtype_code(y.dtype), capi.LAYOUT_C8 if out_c8 else capi.LAYOUT_NCHW
        a.in_act, a.in_alpha, a.in_gain = _ACT['linear'], 0.0, 1.0
        a.act, a.alpha, a.gain, a.clamp = _ACT[act], float(alpha), float(gain), -1.0
        a.operand_format, a.n_tile = fmt_code, n_tile
        a.spade_x, a.spade_mean, a.spade_rstd = capi.ptr(x), capi.ptr(mean), capi.ptr(rstd)
        a.stream = capi.current_stream(x.device)
        rc = capi.load().pg_conv2d_igemm_launch(_byref(a))
        capi.check(rc, 'pg_conv2d_igemm_launch(spade)')
        if sp:
            sp.close()
    return y


# ---------------------------------------------------------------------------------------------------------------- training: dgrad / wgrad

def grad_supported(x_shape, w_shape, dtype, device, stride=(1, 1), padding=(0, 0), dilation=(1, 1), groups=1):
    """Convolutions whose gradients the tcgen05 kernels cover: dense fp32 NCHW on CUDA, stride 1, 'same' padding, 1x1 / 3x3, groups 1."""
    k = int(w_shape[2])
    return (enabled and device.type == 'cuda' and dtype == torch.float32 and groups == 1 and tuple(stride) == (1, 1) and tuple(dilation) == (1, 1) and
            int(w_shape[2]) == int(w_shape[3]) and k in (1, 3) and tuple(padding) == (k // 2, k // 2) and (k == 1 or int(x_shape[3]) <= 256) and
            int(x_shape[0]) * int(x_shape[1]) * int(x_shape[2]) * int(x_shape[3]) < 2 ** 31 and int(x_shape[0]) > 0)


def conv2d_dgrad(dy, w, fmt='bf16'):
    """Input gradient of y = conv2d(x, w, padding=k//2): dx = conv2d(dy, w^T mirrored) -- the forward kernel on dy with Cin <-> Cout swapped and
    flip_weight=False (true convolution).  bf16 operands by default: gradients need the exponent range."""
    return conv2d_igemm(dy, w.transpose(0, 1), flip_weight=False, fmt=fmt)


def conv2d_wgrad(x, dy, ksize, scale=1.0, out=None):
    """Weight gradient of y = conv2d(x, w, padding=k//2) (pg_conv2d_wgrad): dw [Cout, Cin, k, k]; ``out``: accumulate into this tensor."""
    capi = _backend.capi()
    _backend.require_cuda(x, 'conv2d_wgrad')
    x, dy = x.contiguous(), dy.contiguous()
    n, cin, h, wd = (int(v) for v in x.shape)
    cout = int(dy.shape[1])
    assert x.dtype == torch.float32 and dy.dtype == torch.float32 and tuple(dy.shape) == (n, cout, h, wd)
    lib = capi.load()
    nbytes = int(lib.pg_conv2d_wgrad_workspace_bytes(n, cin, cout, h, wd, ksize))
    if nbytes < 0:
        capi.check(2, 'pg_conv2d_wgrad_workspace_bytes')
    dw = out if out is not None else torch.empty([cout, cin, ksize, ksize], dtype=torch.float32, device=x.device)
    assert dw.is_contiguous() and tuple(dw.shape) == (cout, cin, ksize, ksize)
    with torch.cuda.device(x.device):
        capi.require_device()
        ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=x.device)
        sp = capi.span('conv_wgrad', flops=2 * n * cout * cin * ksize * ksize * h * wd, nbytes=4 * (x.numel() + dy.numel() + dw.numel()), tag=f'{cin}->{cout} @{h}x{wd} k{ksize}')
        rc = lib.pg_conv2d_wgrad(capi.ptr(x), capi.ptr(dy), capi.ptr(dw), n, cin, cout, h, wd, ksize, float(scale), int(out is not None),
                                 capi.ptr(ws), nbytes, capi.current_stream(x.device))
        capi.check(rc, 'pg_conv2d_wgrad')
        if sp:
            sp.close()
    return dw
