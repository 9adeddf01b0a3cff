"""Garment-feature completion of ``SynthesisNetworkFull.get_spade_feat`` (reference training/networks.py:5777-5800) on two streaming
kernels (``pg_masked_plane_sum``, ``pg_masked_fill``): pixels the predicted mask covers but the source garment does not are filled with the
garment's mean feature, and the result is written straight into a channel slice of the concatenated upper|lower tensor."""
import torch

from . import _backend


def supported(feat, valid, rest, out):
    if isinstance(out, tuple):                            # (channel-blocked fp16 tensor [N, CB, H, W, 8], first channel block)
        buf, cb0 = out
        return (feat.is_cuda and feat.dtype == torch.float32 and feat.ndim == 4 and valid.dtype == torch.float32 and rest.dtype == torch.float32 and
                buf.dtype == torch.float16 and buf.ndim == 5 and buf.is_contiguous() and feat.shape[1] % 8 == 0 and
                not (torch.is_grad_enabled() and feat.requires_grad) and feat.shape[0] * feat.shape[1] // 8 <= 65535)
    return (feat.is_cuda and feat.dtype == torch.float32 and feat.ndim == 4 and valid.dtype == torch.float32 and rest.dtype == torch.float32 and
            out.dtype in (torch.float32, torch.float16) and not (torch.is_grad_enabled() and feat.requires_grad) and feat.shape[0] * feat.shape[1] <= 65535)


def masked_mean_fill(feat, valid, rest, out, min_count=10):
    """feat [N,C,H,W]; valid, rest [N,1,H,W] in {0,1}; out: [N,C,H,W] view whose channels are contiguous planes (e.g. ``buf[:, c0:c0+C]``).
    out = feat * (1 - rest) + mean * rest,  mean[n,c] = sum(feat * valid) / count,  count = #valid if > min_count else H*W."""
    capi = _backend.capi()
    _backend.require_cuda(feat, 'masked_mean_fill')
    n, c, h, w = (int(v) for v in feat.shape)
    feat = feat.contiguous()
    valid = valid.reshape(n, h * w).contiguous()
    rest = rest.reshape(n, h * w).contiguous()
    c8 = isinstance(out, tuple)
    if c8:
        buf, cb0 = out
        assert tuple(buf.shape[2:]) == (h, w, 8) and int(buf.shape[0]) == n and cb0 + c // 8 <= int(buf.shape[1])
    else:
        assert tuple(out.shape) == (n, c, h, w) and out.stride(3) == 1 and out.stride(2) == w and out.stride(1) == h * w
    sums = torch.empty([n, c], dtype=torch.float32, device=feat.device)
    lib = capi.load()
    with torch.cuda.device(feat.device):
        capi.require_device()
        stream = capi.current_stream(feat.device)
        sp = capi.span('spade_feat', nbytes=4 * feat.numel(), tag='masked sum')
        capi.check(lib.pg_masked_plane_sum(capi.ptr(feat), capi.ptr(valid), capi.ptr(sums), n, c, h * w, stream), 'pg_masked_plane_sum')
        if sp:
            sp.close()
        count = valid.sum(dim=1, keepdim=True)
        enough = (count > min_count).to(feat.dtype)
        count = count * enough + (h * w) * (1 - enough)
        fill = (sums / count).contiguous()
        if c8:
            sp = capi.span('spade_feat', nbytes=(4 + 2) * feat.numel(), tag='masked fill c8')
            capi.check(lib.pg_masked_fill_c8(capi.ptr(feat), capi.ptr(rest), capi.ptr(fill), capi.ptr(buf), n, c, h * w, int(buf.shape[1]), int(cb0), stream),
                       'pg_masked_fill_c8')
            if sp:
                sp.close()
            return buf
        sp = capi.span('spade_feat', nbytes=(4 + out.element_size()) * feat.numel(), tag='masked fill')
        capi.check(lib.pg_masked_fill(capi.ptr(feat), capi.ptr(rest), capi.ptr(fill), capi.ptr(out), n, c, h * w, int(out.stride(0)), capi.dtype_code(out.dtype), stream),
                   'pg_masked_fill')
        if sp:
            sp.close()
    return out
