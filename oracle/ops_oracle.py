"""CPU oracle for the PASTA-GAN operator hot path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch CPU restatement (plain PyTorch CPU tensor ops, any
float dtype, differentiable to any order through autograd) of the algorithm the
reference implements in its ``impl='ref'`` branch.  It is NOT part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  The product package
(``pasta-gan_b200/``) never imports anything from ``oracle/``.

Parity status: PINNED.  ``tests/golden/gen_golden.py`` imports the unmodified
reference from ``/root/reference`` (CPU, ``impl='ref'`` semantics) and writes
seeded input/output/gradient vectors to ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function below against them.
(The reference itself ships no tests or golden vectors — SURVEY.md §4.)

Reference lines each function follows (paths relative to the reference root):

* ``upfirdn2d``        torch_utils/ops/upfirdn2d.py:169-208  (_upfirdn2d_ref)
* ``setup_filter``     torch_utils/ops/upfirdn2d.py:72-116
* ``filter2d/upsample2d/downsample2d``  torch_utils/ops/upfirdn2d.py:272-384
* ``bias_act``         torch_utils/ops/bias_act.py:94-123    (_bias_act_ref)
* ``bias_act_grad``    torch_utils/ops/bias_act.cu:23-147    (grad = 1, 2 branches)
* ``conv2d_resample``  torch_utils/ops/conv2d_resample.py:59-154
* ``fma``              torch_utils/ops/fma.py:15-38
* ``modulated_conv2d`` training/networks.py:37-94
* ``instance_norm_stats``  training/networks.py:4363,4377 (nn.InstanceNorm2d(affine=False) inside Spade_Norm_Block)
* ``spade_norm``       training/networks.py:4371-4379 (Spade_Norm_Block.forward) + the consumer's pre-activation :4345-4349
* ``masked_mean_fill`` training/networks.py:5791-5800 (tail of SynthesisNetworkFull.get_spade_feat)
* ``u8_normalize`` / ``image_to_u8_bgr``  test.py:105-115 / :131-135 (device-side input normalisation, CPU-side photo conversion)

The restatements deliberately use a different decomposition from the reference
(tap-loop FIR instead of a depthwise conv2d; the *definitional* zero-insert →
FIR → conv → FIR → decimate pipeline instead of the transposed-conv fast paths;
scale-activations modulated conv instead of the grouped conv), so agreement with
the goldens is evidence about the algorithm, not about shared code.
"""

import math

import numpy as np
import torch

# -----------------------------------------------------------------------------
# argument parsing (same accepted forms as the reference: upfirdn2d.py:37-68)


def _pair(v):
    if isinstance(v, int):
        return v, v
    v = list(v)
    assert len(v) == 2 and all(isinstance(t, int) for t in v)
    return v[0], v[1]


def _pad4(p):
    if isinstance(p, int):
        return p, p, p, p
    p = list(p)
    assert all(isinstance(t, int) for t in p)
    if len(p) == 2:
        return p[0], p[0], p[1], p[1]
    assert len(p) == 4
    return tuple(p)


def _fsize(f):
    """-> (fw, fh), reference order."""
    if f is None:
        return 1, 1
    assert f.ndim in (1, 2)
    return int(f.shape[-1]), int(f.shape[0])


# -----------------------------------------------------------------------------


def setup_filter(f, normalize=True, flip_filter=False, gain=1, separable=None):
    """upfirdn2d.py:72-116. 1-D filters with < 8 taps become their outer product."""
    if f is None:
        f = 1
    f = torch.as_tensor(f, dtype=torch.float32).clone()
    if f.ndim == 0:
        f = f[None]
    if separable is None:
        separable = f.ndim == 1 and f.numel() >= 8
    if f.ndim == 1 and not separable:
        f = torch.outer(f, f)
    if normalize:
        f = f / f.sum()
    if flip_filter:
        f = f.flip(list(range(f.ndim)))
    return f * (gain ** (f.ndim / 2))


def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1):
    """Zero-insert upsample, pad/crop, FIR (true convolution unless flip_filter),
    decimate.  Tap-loop formulation: y = gain * sum_ij k[i,j] * xpad[i::, j::]."""
    assert x.ndim == 4
    n, c, h, w = x.shape
    upx, upy = _pair(up)
    downx, downy = _pair(down)
    px0, px1, py0, py1 = _pad4(padding)
    if f is None:
        f = torch.ones([1, 1], dtype=torch.float32)
    f = f.to(device=x.device)
    fw, fh = _fsize(f)

    # zero insertion: sample (iy, ix) lands on (iy*upy, ix*upx)
    z = x.new_zeros([n, c, h * upy, w * upx])
    z[:, :, ::upy, ::upx] = x
    # pad (>0) or crop (<0)
    z = torch.nn.functional.pad(z, [max(px0, 0), max(px1, 0), max(py0, 0), max(py1, 0)])
    z = z[:, :, max(-py0, 0): z.shape[2] - max(-py1, 0), max(-px0, 0): z.shape[3] - max(-px1, 0)]
    zh, zw = z.shape[2], z.shape[3]
    oh_full, ow_full = zh - fh + 1, zw - fw + 1
    assert oh_full >= 1 and ow_full >= 1

    # kernel actually correlated with the signal: flipped f == true convolution
    if f.ndim == 1:
        k2 = None
        kx = f.to(x.dtype) * (gain ** 0.5)
        ky = f.to(x.dtype) * (gain ** 0.5)
        if not flip_filter:
            kx, ky = kx.flip(0), ky.flip(0)
        # horizontal then vertical pass (upfirdn2d.py:201-204)
        t = 0
        for j in range(fw):
            t = t + kx[j] * z[:, :, :, j: j + ow_full]
        y = 0
        for i in range(fh):
            y = y + ky[i] * t[:, :, i: i + oh_full, :]
    else:
        k2 = f.to(x.dtype) * gain
        if not flip_filter:
            k2 = k2.flip([0, 1])
        y = 0
        for i in range(fh):
            for j in range(fw):
                y = y + k2[i, j] * z[:, :, i: i + oh_full, j: j + ow_full]
    return y[:, :, ::downy, ::downx]


def filter2d(x, f, padding=0, flip_filter=False, gain=1):
    px0, px1, py0, py1 = _pad4(padding)
    fw, fh = _fsize(f)
    p = [px0 + fw // 2, px1 + (fw - 1) // 2, py0 + fh // 2, py1 + (fh - 1) // 2]
    return upfirdn2d(x, f, padding=p, flip_filter=flip_filter, gain=gain)


def upsample2d(x, f, up=2, padding=0, flip_filter=False, gain=1):
    upx, upy = _pair(up)
    px0, px1, py0, py1 = _pad4(padding)
    fw, fh = _fsize(f)
    p = [px0 + (fw + upx - 1) // 2, px1 + (fw - upx) // 2, py0 + (fh + upy - 1) // 2, py1 + (fh - upy) // 2]
    return upfirdn2d(x, f, up=up, padding=p, flip_filter=flip_filter, gain=gain * upx * upy)


def downsample2d(x, f, down=2, padding=0, flip_filter=False, gain=1):
    dx, dy = _pair(down)
    px0, px1, py0, py1 = _pad4(padding)
    fw, fh = _fsize(f)
    p = [px0 + (fw - dx + 1) // 2, px1 + (fw - dx) // 2, py0 + (fh - dy + 1) // 2, py1 + (fh - dy) // 2]
    return upfirdn2d(x, f, down=down, padding=p, flip_filter=flip_filter, gain=gain)


# -----------------------------------------------------------------------------
# bias_act  (bias_act.py:23-33 table; defaults alpha / gain; cuda_idx order)

_SELU_SCALE = 1.0507009873554804934193349852946
_SELU_ALPHA = 1.6732632423543772848170429916717

ACT_TABLE = {
    #  name      (idx, def_alpha, def_gain,   ref,  has_2nd_grad)
    'linear':   (1, 0.0, 1.0, '', False),
    'relu':     (2, 0.0, math.sqrt(2), 'y', False),
    'lrelu':    (3, 0.2, math.sqrt(2), 'y', False),
    'tanh':     (4, 0.0, 1.0, 'y', True),
    'sigmoid':  (5, 0.0, 1.0, 'y', True),
    'elu':      (6, 0.0, 1.0, 'y', True),
    'selu':     (7, 0.0, 1.0, 'y', True),
    'softplus': (8, 0.0, 1.0, 'y', True),
    'swish':    (9, 0.0, math.sqrt(2), 'x', True),
}


def _act(x, act, alpha):
    if act == 'linear':
        return x
    if act == 'relu':
        return torch.clamp_min(x, 0)
    if act == 'lrelu':
        return torch.where(x > 0, x, x * alpha)
    if act == 'tanh':
        return torch.tanh(x)
    if act == 'sigmoid':
        return torch.sigmoid(x)
    if act == 'elu':
        return torch.where(x >= 0, x, torch.expm1(x))
    if act == 'selu':
        return _SELU_SCALE * torch.where(x >= 0, x, _SELU_ALPHA * torch.expm1(x))
    if act == 'softplus':
        return torch.nn.functional.softplus(x)
    if act == 'swish':
        return torch.sigmoid(x) * x
    raise KeyError(act)


def _resolve(act, alpha, gain, clamp):
    _, da, dg, _, _ = ACT_TABLE[act]
    alpha = float(da if alpha is None else alpha)
    gain = float(dg if gain is None else gain)
    clamp = float(-1 if clamp is None else clamp)
    return alpha, gain, clamp


def bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None):
    """y = clamp(act(x + b) * gain)   (bias_act.py:94-123)."""
    alpha, gain, clamp = _resolve(act, alpha, gain, clamp)
    if b is not None:
        assert b.ndim == 1 and b.shape[0] == x.shape[dim]
        x = x + b.reshape([-1 if i == dim else 1 for i in range(x.ndim)])
    y = _act(x, act, alpha)
    if gain != 1:
        y = y * gain
    if clamp >= 0:
        y = y.clamp(-clamp, clamp)
    return y


def bias_act_grad(grad, dy_in, x=None, b=None, y=None, dy=None, dim=1, act='linear', alpha=None, gain=None, clamp=None):
    """Explicit first/second-order gradient kernels (the plugin's grad=1 / grad=2
    modes, bias_act.cu:54-142), written from the closed forms:

    grad=1:  out = dy_in * gain * act'(.)      masked to 0 where |y| >= clamp
    grad=2:  out = dy_in * dy * gain * act''(.) masked likewise
    where act' / act'' are expressed through yy = y/gain (or x+b for swish).
    """
    alpha, gain, clamp = _resolve(act, alpha, gain, clamp)
    assert grad in (1, 2)
    t = dy_in
    yy = (y / gain) if (y is not None and gain != 0) else None
    xb = None
    if x is not None:
        xb = x if b is None else x + b.reshape([-1 if i == dim else 1 for i in range(x.ndim)])
    zero = torch.zeros_like(t)
    if act == 'linear':
        d1, d2 = torch.ones_like(t), zero
    elif act == 'relu':
        d1, d2 = (yy > 0).to(t.dtype), zero
    elif act == 'lrelu':
        d1, d2 = torch.where(yy > 0, torch.ones_like(t), torch.full_like(t, alpha)), zero
    elif act == 'tanh':
        d1 = 1 - yy * yy
        d2 = d1 * (-2 * yy)
    elif act == 'sigmoid':
        d1 = yy * (1 - yy)
        d2 = d1 * (1 - 2 * yy)
    elif act == 'elu':
        d1 = torch.where(yy >= 0, torch.ones_like(t), yy + 1)
        d2 = torch.where(yy >= 0, zero, yy + 1)
    elif act == 'selu':
        sa = _SELU_SCALE * _SELU_ALPHA
        d1 = torch.where(yy >= 0, torch.full_like(t, _SELU_SCALE), yy + sa)
        d2 = torch.where(yy >= 0, zero, yy + sa)
    elif act == 'softplus':
        c = torch.exp(-yy)
        d1 = 1 - c
        d2 = c * (1 - c)
    elif act == 'swish':
        s = torch.sigmoid(xb)
        d1 = s * (1 + xb * (1 - s))
        d2 = s * (1 - s) * (2 + xb * (1 - 2 * s))
        y = xb * s * gain  # the clamp mask is recomputed from x (bias_act.cu:128)
    out = t * (d1 if grad == 1 else d2) * gain
    if grad == 2 and dy is not None:
        out = out * dy
    if clamp >= 0:
        out = torch.where((y > -clamp) & (y < clamp), out, zero)
    return out


# -----------------------------------------------------------------------------


def _conv(x, w, stride=1, padding=0, groups=1, flip_weight=True):
    """flip_weight=True is correlation (== F.conv2d); False is true convolution
    (conv2d_resample.py:29-54)."""
    if not flip_weight:
        w = w.flip([2, 3])
    return torch.nn.functional.conv2d(x, w, stride=stride, padding=padding, groups=groups)


def conv2d_resample(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False):
    """Definitional pipeline: pad once -> zero-insert+FIR (gain up^2) -> conv -> FIR+decimate.
    This is the reference's generic fallback (conv2d_resample.py:150-154); its five
    fast paths (:107-147) are algebraically identical and are what the goldens ran."""
    assert x.ndim == 4 and w.ndim == 4
    fw, fh = _fsize(f)
    px0, px1, py0, py1 = _pad4(padding)
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
    x = upfirdn2d(x, f if up > 1 else None, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    x = _conv(x, w, groups=groups, flip_weight=flip_weight)
    if down > 1:
        x = upfirdn2d(x, f, down=down, flip_filter=flip_filter)
    return x


def fma(a, b, c):
    return a * b + c


def modulated_conv2d(x, weight, styles, noise=None, up=1, down=1, padding=0, resample_filter=None,
                     demodulate=True, flip_weight=True, fused_modconv=True):
    """networks.py:37-94, restated in the scale-activations form for BOTH values of
    ``fused_modconv`` (identities I7/I8):
        y = conv(x * s[n,i], W) * d[n,o] + noise,
        d[n,o] = rsqrt(sum_i s[n,i]^2 * sum_k W[o,i,k]^2 + 1e-8).
    (The fp16 pre-normalisation branch, :57-59, is part of the contract too.)"""
    n = x.shape[0]
    o, i, kh, kw = weight.shape
    assert x.shape[1] == i and styles.shape == (n, i)
    if x.dtype == torch.float16 and demodulate:
        weight = weight * (1 / np.sqrt(i * kh * kw) / weight.norm(float('inf'), dim=[1, 2, 3], keepdim=True))
        styles = styles / styles.norm(float('inf'), dim=1, keepdim=True)
    d = None
    if demodulate:
        wsq = weight.square().sum(dim=[2, 3])            # [O, I]
        d = (styles.square() @ wsq.t() + 1e-8).rsqrt()   # [N, O]
    y = x * styles.to(x.dtype).reshape(n, i, 1, 1)
    y = conv2d_resample(y, weight.to(x.dtype), f=resample_filter, up=up, down=down, padding=padding, flip_weight=flip_weight)
    if d is not None:
        y = y * d.to(x.dtype).reshape(n, o, 1, 1)
    if noise is not None:
        y = y + noise.to(x.dtype)
    return y


# -----------------------------------------------------------------------------
# "Fast port": the same algorithm in the formulation a CPU implementation would actually ship — the FIR as a
# depthwise conv2d and conv2d_resample lowered the way the reference lowers it (strided / transposed-strided conv,
# conv2d_resample.py:107-147).  Used for the CPU-baseline timing so the baseline is not handicapped by the
# definitional tap loops above; pinned to the same golden vectors (tests/test_oracle_golden.py).


def upfirdn2d_fast(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1):
    n, c, h, w = x.shape
    upx, upy = _pair(up)
    downx, downy = _pair(down)
    px0, px1, py0, py1 = _pad4(padding)
    if f is None:
        f = torch.ones([1, 1], dtype=torch.float32)
    if upx > 1 or upy > 1:
        z = x.new_zeros([n, c, h * upy, w * upx])
        z[:, :, ::upy, ::upx] = x
        x = z
    x = torch.nn.functional.pad(x, [max(px0, 0), max(px1, 0), max(py0, 0), max(py1, 0)])
    x = x[:, :, max(-py0, 0): x.shape[2] - max(-py1, 0), max(-px0, 0): x.shape[3] - max(-px1, 0)]
    k = (f * (gain ** (f.ndim / 2))).to(x.dtype)
    if not flip_filter:
        k = k.flip(list(range(k.ndim)))
    if k.ndim == 2:
        x = torch.nn.functional.conv2d(x, k[None, None].repeat(c, 1, 1, 1), groups=c)
    else:
        x = torch.nn.functional.conv2d(x, k[None, None, None, :].repeat(c, 1, 1, 1), groups=c)
        x = torch.nn.functional.conv2d(x, k[None, None, :, None].repeat(c, 1, 1, 1), groups=c)
    return x[:, :, ::downy, ::downx]


def conv2d_resample_fast(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False):
    cout, cin_g, kh, kw = w.shape
    fw, fh = _fsize(f)
    px0, px1, py0, py1 = _pad4(padding)
    if up > 1:
        px0 += (fw + up - 1) // 2; px1 += (fw - up) // 2; py0 += (fh + up - 1) // 2; py1 += (fh - up) // 2
    if down > 1:
        px0 += (fw - down + 1) // 2; px1 += (fw - down) // 2; py0 += (fh - down + 1) // 2; py1 += (fh - down) // 2
    if kw == 1 and kh == 1 and down > 1 and up == 1:
        return _conv(upfirdn2d_fast(x, f, down=down, padding=[px0, px1, py0, py1], flip_filter=flip_filter), w, groups=groups, flip_weight=flip_weight)
    if kw == 1 and kh == 1 and up > 1 and down == 1:
        return upfirdn2d_fast(_conv(x, w, groups=groups, flip_weight=flip_weight), f, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    if down > 1 and up == 1:
        return _conv(upfirdn2d_fast(x, f, padding=[px0, px1, py0, py1], flip_filter=flip_filter), w, stride=down, groups=groups, flip_weight=flip_weight)
    if up > 1:
        wt = w.transpose(0, 1) if groups == 1 else \
            w.reshape(groups, cout // groups, cin_g, kh, kw).transpose(1, 2).reshape(groups * cin_g, cout // groups, kh, kw)
        px0 -= kw - 1; px1 -= kw - up; py0 -= kh - 1; py1 -= kh - up
        pxt, pyt = max(min(-px0, -px1), 0), max(min(-py0, -py1), 0)
        wt = wt if not flip_weight else wt.flip([2, 3])          # transposed conv flips once more
        x = torch.nn.functional.conv_transpose2d(x, wt, stride=up, padding=[pyt, pxt], groups=groups)
        x = upfirdn2d_fast(x, f, padding=[px0 + pxt, px1 + pxt, py0 + pyt, py1 + pyt], gain=up ** 2, flip_filter=flip_filter)
        return upfirdn2d_fast(x, f, down=down, flip_filter=flip_filter) if down > 1 else x
    if px0 == px1 and py0 == py1 and px0 >= 0 and py0 >= 0:
        return _conv(x, w, padding=[py0, px0], groups=groups, flip_weight=flip_weight)
    return _conv(upfirdn2d_fast(x, None, padding=[px0, px1, py0, py1]), w, groups=groups, flip_weight=flip_weight)


def modulated_conv2d_fast(x, weight, styles, noise=None, up=1, down=1, padding=0, resample_filter=None,
                          demodulate=True, flip_weight=True, fused_modconv=True):
    n = x.shape[0]
    o, i, kh, kw = weight.shape
    d = None
    if demodulate:
        d = (styles.square() @ weight.square().sum(dim=[2, 3]).t() + 1e-8).rsqrt()
    y = conv2d_resample_fast(x * styles.reshape(n, i, 1, 1), weight, f=resample_filter, up=up, down=down, padding=padding, flip_weight=flip_weight)
    if d is not None:
        y = y * d.reshape(n, o, 1, 1)
    return y if noise is None else y + noise


# The operator table handed to the host-side network mirror when tests / the CPU
# baseline want the whole generator evaluated by the oracle.
# ----------------------------------------------------------------------------- helpers of the SPADE / IO rows

def instance_norm_stats(x, eps=1e-5):
    """(mean, rstd) per (n, c) plane: biased variance, as nn.InstanceNorm2d(affine=False) (networks.py:4363)."""
    mean = x.mean(dim=(2, 3))
    var = (x - mean[:, :, None, None]).square().mean(dim=(2, 3))
    return mean, (var + eps).rsqrt()


def spade_norm(x, gamma, beta, act=None, gain=1.0, eps=1e-5):
    """normalized * (1 + gamma) + beta (networks.py:4377-4379), optionally followed by the consuming Spade conv's pre-activation
    (relu / lrelu with gain, :4345-4349)."""
    mean, rstd = instance_norm_stats(x, eps)
    y = (x - mean[:, :, None, None]) * rstd[:, :, None, None] * (1 + gamma) + beta
    return y if act is None else bias_act(y, None, act=act, gain=gain)


def masked_mean_fill(feat, valid, rest, min_count=10):
    """feat * (1 - rest) + mean * rest with mean = sum(feat * valid) / count and count falling back to H * W when at most ``min_count``
    pixels are valid (networks.py:5791-5800)."""
    h, w = feat.shape[2:]
    total = (feat * valid).sum(dim=(2, 3), keepdim=True)
    count = valid.sum(dim=(2, 3), keepdim=True)
    enough = (count > min_count).to(feat.dtype)
    count = count * enough + (h * w) * (1 - enough)
    return feat * (1 - rest) + (total / count) * rest


def u8_normalize(x_u8, normalize=True):
    """x.to(float32) / 127.5 - 1 (test.py:105-111); masks are plain casts (:109, :112).  float32 arithmetic, true division."""
    x = x_u8.to(torch.float32)
    return x / 127.5 - 1 if normalize else x


def image_to_u8_bgr(img, crop=None):
    """uint8(clip((img.transpose(1, 2, 0) + 1) * 127.5, 0, 255)) with the columns cropped to the photo and RGB -> BGR (test.py:131-135);
    img [N,3,H,W] float32 -> [N,H,x1-x0,3] uint8."""
    import numpy as np
    g = img.detach().cpu().numpy().astype(np.float32)
    x0, x1 = crop if crop is not None else (g.shape[3] // 8, g.shape[3] - g.shape[3] // 8)
    out = [np.clip(((g[i].transpose(1, 2, 0) + np.float32(1.0)) * np.float32(127.5))[:, x0:x1, [2, 1, 0]], 0, 255).astype(np.uint8) for i in range(g.shape[0])]
    return torch.from_numpy(np.stack(out))


class _Namespace:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def operator_table(fast=False):
    """fast=False: the definitional restatement (the parity checker).  fast=True: the lowered CPU port (the timing baseline)."""
    if fast:
        def up2(x, f, up=2, padding=0, flip_filter=False, gain=1):
            upx, upy = _pair(up)
            px0, px1, py0, py1 = _pad4(padding)
            fw, fh = _fsize(f)
            p = [px0 + (fw + upx - 1) // 2, px1 + (fw - upx) // 2, py0 + (fh + upy - 1) // 2, py1 + (fh - upy) // 2]
            return upfirdn2d_fast(x, f, up=up, padding=p, flip_filter=flip_filter, gain=gain * upx * upy)
        return _Namespace(name='oracle-cpu-fast', setup_filter=lambda f, device=None, **kw: setup_filter(f, **kw),
                          upfirdn2d=upfirdn2d_fast, upsample2d=up2, bias_act=bias_act, conv2d_resample=conv2d_resample_fast,
                          fma=fma, modulated_conv2d=modulated_conv2d_fast, act_def_gain={k: v[2] for k, v in ACT_TABLE.items()})
    return _Namespace(
        name='oracle-cpu',
        setup_filter=lambda f, device=None, **kw: setup_filter(f, **kw),
        upfirdn2d=upfirdn2d, filter2d=filter2d, upsample2d=upsample2d, downsample2d=downsample2d,
        bias_act=bias_act, conv2d_resample=conv2d_resample, fma=fma, modulated_conv2d=modulated_conv2d,
        act_def_gain={k: v[2] for k, v in ACT_TABLE.items()},
    )
