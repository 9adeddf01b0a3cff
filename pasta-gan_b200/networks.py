"""Host-side mirror of the PASTA-GAN modules that CALL the operator hot path.

These classes reproduce the module tree, parameter/buffer names and forward arithmetic of the reference's
full-body generator so that (a) a reference ``state_dict`` loads into them key-for-key and (b) the try-on
workload of BASELINE.json (256x192 full-body generator inference) can be driven end to end through the
sm_100a operators.  They are a re-expression, not a copy: every layer is a thin composition over an
*operator table* (``ops``) whose entries have the reference's ``torch_utils.ops`` semantics.

    reference class (training/networks.py)          here
    ------------------------------------------------------------------
    modulated_conv2d               :37-94            modulated_conv2d
    FullyConnectedLayer            :99-130           FullyConnectedLayer
    Conv2dLayer                    :133-180          Conv2dLayer
    MappingNetwork                 :184-262          MappingNetwork
    SynthesisLayer                 :264-316          SynthesisLayer
    ResBlock                       :529-558          ResBlock
    ConstEncoderNetwork            :561-578          ConstEncoderNetwork
    Dense                          :594-611          Dense
    Spade_Conv2dLayer              :4305-4354        SpadeConv2dLayer
    Spade_Norm_Block               :4358-4379        SpadeNormBlock
    StyleEncoderNetworkV16         :4837-4883        StyleEncoderNetworkV16   (patch-routed style encoder)
    Spade_ResBlockV2               :5230-5274        SpadeResBlockV2
    ToRGBLayerFull                 :5583-5611        ToRGBLayerFull
    SynthesisBlockFull             :5615-5719        SynthesisBlockFull
    SynthesisNetworkFull           :5723-5840        SynthesisNetworkFull
    GeneratorFull                  :5844-5880        GeneratorFull
    ToRGBLayer                     :320-336          ToRGBLayer
    SynthesisBlockV_512            :3578-3677        SynthesisBlock512
    SynthesisNetwork_512           :3680-3729        SynthesisNetwork512
    StyleEncoderNetwork_512        :3732-3779        StyleEncoderNetwork512
    Generator_512                  :3782-3815        Generator512
    DiscriminatorBlock             :917-997          DiscriminatorBlock
    MinibatchStdLayer              :1001-1023        MinibatchStdLayer
    DiscriminatorEpilogue          :1027-1081        DiscriminatorEpilogue
    Discriminator                  :1085-1143        Discriminator

The default operator table is the CUDA product (``cuda_ops()``); tests and the CPU baseline inject the oracle's
table with ``use_ops(module, table)``.  The product never imports the oracle.
"""
import types
import weakref

import os

import numpy as np
import torch
import torch.nn as nn

from .torch_utils import misc

# ----------------------------------------------------------------------------- operator tables

_CUDA_OPS = None


def cuda_ops():
    """The product operator table: every entry lands in libpasta_b200.so (dense convs: see conv2d_gradfix)."""
    global _CUDA_OPS
    if _CUDA_OPS is None:
        from .torch_utils.ops import upfirdn2d as U, bias_act as B, conv2d_resample as C, fma as F, conv_igemm as K
        _CUDA_OPS = types.SimpleNamespace(
            name='sm100a',
            setup_filter=U.setup_filter, upfirdn2d=U.upfirdn2d, filter2d=U.filter2d, upsample2d=U.upsample2d,
            downsample2d=U.downsample2d, bias_act=B.bias_act, conv2d_resample=C.conv2d_resample, fma=F.fma,
            modulated_conv2d=modulated_conv2d, conv_layer=conv_layer, modconv_layer=modconv_layer, torgb_skip=_torgb_skip, spade_conv_norm=_spade_conv_norm, masked_mean_fill=_masked_mean_fill, half_intermediates=_half_intermediates, c8_ok=_c8_ok,
            instance_stats=lambda x: K.instance_stats(x) if (K.enabled and x.is_cuda and x.dtype == torch.float32) else None,
            act_def_gain={k: float(v.def_gain) for k, v in B.activation_funcs.items()},
        )
    return _CUDA_OPS


class OpsModule(nn.Module):
    """nn.Module whose numerical work goes through an operator table (default: the CUDA product)."""
    _ops = None

    @property
    def ops(self):
        return self._ops if self._ops is not None else cuda_ops()


def use_ops(module, table):
    """Route every layer under ``module`` through ``table`` (None restores the CUDA product)."""
    for m in module.modules():
        if isinstance(m, OpsModule):
            m._ops = table
    return module


_ACT_GAIN = dict(linear=1.0, relu=float(np.sqrt(2)), lrelu=float(np.sqrt(2)), tanh=1.0, sigmoid=1.0, elu=1.0, selu=1.0,
                 softplus=1.0, swish=float(np.sqrt(2)))
_FIR = [1, 3, 3, 1]


def _fir_buffer(taps):
    """[1,3,3,1] -> normalised 4x4 outer product (upfirdn2d.setup_filter), built without touching the op table."""
    f = torch.as_tensor(taps, dtype=torch.float32)
    if f.ndim == 1 and f.numel() < 8:
        f = torch.outer(f, f)
    return f / f.sum()


# ----------------------------------------------------------------------------- functional pieces


def normalize_2nd_moment(x, dim=1, eps=1e-8):
    return x * (x.square().mean(dim=dim, keepdim=True) + eps).rsqrt()


@misc.profiled_function
def modulated_conv2d(x, weight, styles, noise=None, up=1, down=1, padding=0, resample_filter=None, demodulate=True,
                     flip_weight=True, fused_modconv=True):
    """Style-modulated convolution with the reference's signature (networks.py:37-49).

    Both values of ``fused_modconv`` evaluate  y = conv(x * s[n,i], W) * d[n,o] + noise  with
    d[n,o] = rsqrt(sum_i s[n,i]^2 * sum_k W[o,i,k]^2 + 1e-8): the per-sample weight tensor [N,O,I,k,k] and the
    groups=N convolution of the reference's fused branch (:84-94) are never materialised (SURVEY.md appendix A,
    I7/I8).  The fp16 pre-normalisation of :57-59 is kept."""
    from .torch_utils.ops import conv2d_resample as C, fma as F, conv_igemm as K
    n = int(x.shape[0])
    cout, cin, kh, kw = weight.shape
    misc.assert_shape(weight, [cout, cin, kh, kw])
    misc.assert_shape(x, [n, cin, None, None])
    misc.assert_shape(styles, [n, cin])
    if kh == kw and padding == kh // 2 and not styles.requires_grad and \
            K.supported(x, weight, up=up, down=down, f=resample_filter, padding=(padding,) * 4):
        # inference on the tcgen05 kernel: style in the operand prologue, demodulation and noise in the epilogue, one launch
        dcoefs = torch.addmm(_EPS.get(x.device), styles.square(), _weight_sq_sums(weight).t()).rsqrt() if demodulate else None
        return K.conv2d_igemm(x, weight, f=resample_filter, up=up, flip_weight=flip_weight, styles=styles, dcoefs=dcoefs, noise=noise,
                              cache_weights=isinstance(weight, nn.Parameter))
    if x.dtype == torch.float16 and demodulate:
        weight = weight * (1 / np.sqrt(cin * kh * kw) / weight.norm(float('inf'), dim=[1, 2, 3], keepdim=True))
        styles = styles / styles.norm(float('inf'), dim=1, keepdim=True)
    dcoefs = None
    if demodulate:
        wsq = weight.float().square().sum(dim=[2, 3])                        # [O, I]
        dcoefs = torch.addmm(torch.full([1, 1], 1e-8, device=x.device), styles.float().square(), wsq.t()).rsqrt()
    x = x * styles.to(x.dtype).reshape(n, cin, 1, 1)
    x = C.conv2d_resample(x=x, w=weight.to(x.dtype), f=resample_filter, up=up, down=down, padding=padding, flip_weight=flip_weight)
    if dcoefs is not None and noise is not None:
        x = F.fma(x, dcoefs.to(x.dtype).reshape(n, cout, 1, 1), noise.to(x.dtype))
    elif dcoefs is not None:
        x = x * dcoefs.to(x.dtype).reshape(n, cout, 1, 1)
    elif noise is not None:
        x = x.add_(noise.to(x.dtype))
    return x


class _Eps:
    """[1,1] tensors holding 1e-8 per device (the addmm bias of the demodulation GEMV), created once."""
    def __init__(self):
        self.t = {}

    def get(self, device):
        if device not in self.t:
            self.t[device] = torch.full([1, 1], 1e-8, device=device)
        return self.t[device]


_EPS = _Eps()
_WSQ_CACHE = {}


def _weight_sq_sums(weight):
    """sum_k W[o,i,k]^2  ([O, I]); cached per Parameter version (it only changes when the optimizer steps)."""
    if not isinstance(weight, nn.Parameter) or torch.is_grad_enabled():
        return weight.square().sum(dim=[2, 3])
    key = id(weight)
    hit = _WSQ_CACHE.get(key)
    if hit is not None and hit['w']() is not weight:
        _WSQ_CACHE.pop(key, None)
        hit = None
    ver = (weight._version, weight.data_ptr())
    if hit is None:
        hit = dict(w=weakref.ref(weight, lambda _r, key=key: _WSQ_CACHE.pop(key, None)), ver=ver, sq=weight.detach().square().sum(dim=[2, 3]))
        _WSQ_CACHE[key] = hit
    elif hit['ver'] != ver:
        hit['sq'].copy_(weight.detach().square().sum(dim=[2, 3]))      # in place: one buffer per parameter, valid for captured graphs
        hit['ver'] = ver
    return hit['sq']


def refresh_weight_caches(module=None):
    """After load_state_dict / an optimizer step under a captured session: bring every derived buffer (per-parameter squared sums, StyleBanks)
    up to date in place.  The packed GEMM tiles are handled by conv_igemm.refresh_packed_weights()."""
    for hit in list(_WSQ_CACHE.values()):
        w = hit['w']()
        if w is not None and hit['ver'] != (w._version, w.data_ptr()):
            hit['sq'].copy_(w.detach().square().sum(dim=[2, 3]))
            hit['ver'] = (w._version, w.data_ptr())
    if module is not None:
        for m in module.modules():
            bank = getattr(m, '_bank', None)
            if bank is not None:
                with torch.no_grad():
                    bank.refresh()


def conv_layer(x, w, b=None, f=None, up=1, down=1, padding=0, flip_weight=True, act='linear', act_gain=1.0, clamp=None,
               in_act=None, in_gain=1.0, w_scale=1.0, cache_weights=False, x2=None, residual=None, out_dtype=None, half_ok=False, out_c8=False):
    """Product form of one plain conv layer: [in_act([x ; x2]) * in_gain ->] conv2d_resample -> bias_act [-> + residual].  When the tcgen05
    kernel covers the shape the whole layer is ONE launch (bias, activation, gain, clamp and the residual add live in the GEMM epilogue; the
    SPADE pre-activation and the channel concatenation live in the operand prologue); otherwise it is composed from the same operators
    the reference calls."""
    from .torch_utils.ops import conv_igemm as K, conv2d_resample as C, bias_act as B
    pad4 = (padding,) * 4 if isinstance(padding, int) else None
    if K.is_c8(x):
        # channel-blocked fp16 from one of our own epilogues: TMA operand path.  The producer only emits this layout for consumers it has checked
        # (K.c8_input_ok), so there is nothing to fall back to here.
        assert in_act is None and act in ('linear', 'relu', 'lrelu') and padding == int(w.shape[2]) // 2 and (down == 1 or (down == 2 and up == 1 and x2 is None))
        return K.conv2d_igemm(x, w, f=f, up=up, down=down, flip_weight=flip_weight, bias=b, act=act, gain=act_gain, clamp=clamp, w_scale=w_scale,
                              cache_weights=cache_weights, x2=x2, residual=residual, out_dtype=out_dtype or torch.float32, out_c8=out_c8)
    if act in ('linear', 'relu', 'lrelu') and in_act in (None, 'relu', 'lrelu') and \
            K.supported(x, w, up=up, down=down, f=f, padding=pad4, x2=x2, residual=residual, allow_half=half_ok):
        return K.conv2d_igemm(x, w, f=f, up=up, down=down, flip_weight=flip_weight, bias=b, in_act=in_act or 'linear', in_gain=in_gain,
                              act=act, gain=act_gain, clamp=clamp, w_scale=w_scale, cache_weights=cache_weights, x2=x2, residual=residual,
                              out_dtype=out_dtype or torch.float32, out_c8=out_c8)
    assert not out_c8, 'a channel-blocked output was requested from a layer the tcgen05 kernel does not cover'
    if half_ok and x.dtype == torch.float16:
        x = x.float()                                     # an fp16 intermediate of ours reached a layer the tcgen05 kernel does not cover
    if x2 is not None:
        x = torch.cat([x, x2.to(x.dtype)], dim=1)
    if in_act is not None:
        x = B.bias_act(x, None, act=in_act, gain=in_gain)
    if w_scale != 1.0:
        w = w * w_scale
    x = C.conv2d_resample(x=x, w=w.to(x.dtype), f=f, up=up, down=down, padding=padding, flip_weight=flip_weight)
    if not (b is None and act == 'linear' and act_gain == 1 and clamp is None):
        x = B.bias_act(x, b, act=act, gain=act_gain, clamp=clamp)
    x = x if residual is None else residual.add_(x)
    return x if out_dtype is None else x.to(out_dtype)


def modconv_layer(x, weight, styles, noise=None, up=1, padding=0, resample_filter=None, demodulate=True, flip_weight=True,
                  fused_modconv=True, bias=None, act='linear', act_gain=1.0, clamp=None, dcoefs=None, styles_normalized=False, out_c8=False):
    """Product form of modulated_conv2d + bias_act (SynthesisLayer / ToRGB): one tcgen05 launch with the style folded into the
    activation operand, demodulation / noise / bias / activation / clamp in the epilogue.  A channel-blocked fp16 ``x`` is loaded by TMA and cannot be
    scaled on the way in: the styles are then folded into per-sample packed weights (the reference's own formulation, networks.py:64-66)."""
    from .torch_utils.ops import conv_igemm as K, bias_act as B
    k = int(weight.shape[2])
    if K.is_c8(x):
        assert act in ('linear', 'relu', 'lrelu') and padding == k // 2
        if not demodulate:
            dcoefs = None
        elif dcoefs is None:
            dcoefs = torch.addmm(_EPS.get(x.device), styles.square(), _weight_sq_sums(weight).t()).rsqrt()
        return K.conv2d_igemm(x, weight, f=resample_filter, up=up, flip_weight=flip_weight, styles=styles, dcoefs=dcoefs, noise=noise,
                              bias=bias, act=act, gain=act_gain, clamp=clamp, fold_styles=True, styles_normalized=True, out_c8=out_c8)
    if act in ('linear', 'relu', 'lrelu') and padding == k // 2 and \
            K.supported(x, weight, up=up, f=resample_filter, padding=(padding,) * 4) and not styles.requires_grad:
        if not demodulate:
            dcoefs = None
        elif dcoefs is None:                              # (a StyleBank hands precomputed coefficients in)
            dcoefs = torch.addmm(_EPS.get(x.device), styles.square(), _weight_sq_sums(weight).t()).rsqrt()
        return K.conv2d_igemm(x, weight, f=resample_filter, up=up, flip_weight=flip_weight, styles=styles, dcoefs=dcoefs, noise=noise,
                              bias=bias, act=act, gain=act_gain, clamp=clamp, cache_weights=isinstance(weight, nn.Parameter),
                              styles_normalized=styles_normalized, out_c8=out_c8)
    # (unit-inf-norm styles from the StyleBank are fine here: a demodulated layer is invariant to the scale of its styles, and only those get normalised)
    assert not out_c8 and (demodulate or not styles_normalized), 'channel-blocked outputs are only produced on the tcgen05 path'
    x = modulated_conv2d(x=x, weight=weight, styles=styles, noise=noise, up=up, padding=padding, resample_filter=resample_filter,
                         demodulate=demodulate, flip_weight=flip_weight, fused_modconv=fused_modconv)
    return B.bias_act(x, bias, act=act, gain=act_gain, clamp=clamp)


def _spade_conv_norm(x, actv, w_gamma, w_beta, w_scale, post_act, stats=None, out_dtype=None, out_c8=False):
    """Product SPADE normalisation: instance-norm + (1 + gamma) * . + beta (+ the consumer's pre-activation) inside the epilogue of the
    merged gamma|beta convolution; None when the shape is not covered."""
    from .torch_utils.ops import conv_igemm as K
    if not K.spade_supported(x, actv, w_gamma, w_beta):
        return None
    act, gain = post_act if post_act is not None else ('linear', 1.0)
    if act not in ('linear', 'relu', 'lrelu'):
        return None
    return K.spade_conv_norm(x, actv, w_gamma, w_beta, w_scale=w_scale, act=act, gain=gain, stats=stats, out_dtype=out_dtype or torch.float32, out_c8=out_c8)


def _half_intermediates(x):
    """fp16 for tensors that only travel between two tcgen05 convolutions (SPADE `actv`, the normalised maps, the garment features): the
    consumer rounds to fp16 operands anyway, so the values it multiplies are the same bits at half the traffic."""
    from .torch_utils.ops import conv_igemm as K
    return (K.enabled and K.operand_format == 'fp16' and os.environ.get('PASTA_B200_HALF_INTERMEDIATES', '1') != '0' and x.is_cuda and
            not torch.is_grad_enabled() and x.shape[3] % 2 == 0 and x.shape[3] <= 256)


def _c8_ok(channels, h, w, k, up=1, down=1):
    """May a tensor [N, channels, h, w] consumed only by a plain k x k tcgen05 convolution travel as channel-blocked fp16 (TMA operand path)?"""
    from .torch_utils.ops import conv_igemm as K
    return (not torch.is_grad_enabled()) and os.environ.get('PASTA_B200_HALF_INTERMEDIATES', '1') != '0' and \
        K.c8_input_ok(int(channels), int(h), int(w), int(k), up, down)


def _masked_mean_fill(feat, valid, rest, out):
    """Product get_spade_feat tail (two streaming kernels); None when the tensors are not covered."""
    from .torch_utils.ops import spade_feat as S
    if not S.supported(feat, valid, rest, out):
        return None
    return S.masked_mean_fill(feat, valid, rest, out)


def _torgb_skip(x, weight, styles, bias, clamp, img, f):
    """Product ToRGB skip: one streaming kernel when the shape allows it, else None (the caller composes the reference's four calls)."""
    from .torch_utils.ops import torgb as T
    if not T.supported(x, weight, img, f) or styles.requires_grad:
        return None
    return T.torgb_skip(x, weight, styles=styles, bias=bias, clamp=clamp, img=img, f=f)


class StyleBank:
    """All per-layer styles (affine(w), reference :296-299 / :5602) and demodulation coefficients (:65-68, in the GEMV form of SURVEY appendix A, I8)
    of a synthesis network in TWO batched GEMMs instead of ~6 tiny kernels per layer (weight * gain, addmm + split-K reduce, square, addmm,
    rsqrt): they depend only on ``ws``, which is known before the first block runs.  Inference only (no autograd); the packed affine
    weights are cached and rebuilt when any source parameter changes.  Layers pick their rows up through ``layer._pre``."""

    def __init__(self):
        self.key = None

    def _pack(self, layers):
        key = tuple((id(l), l.affine.weight._version, l.affine.bias._version, l.weight._version, l.weight.data_ptr()) for l in layers)
        if key == self.key:
            return
        dev = layers[0].weight.device
        L, wd = len(layers), int(layers[0].affine.weight.shape[1])
        cmax = max(int(l.affine.weight.shape[0]) for l in layers)
        self.demod = [i for i, l in enumerate(layers) if isinstance(l, SynthesisLayer)]
        omax = max([int(layers[i].weight.shape[0]) for i in self.demod] or [1])
        same = self.key is not None and len(self.key) == len(key) and all(a[0] == b[0] for a, b in zip(self.key, key)) and self.W.device == dev
        if same:
            # same layers, new parameter values: rebuild IN PLACE (captured CUDA graphs hold the addresses of W / B / Q)
            W, B, Q = self.W.zero_(), self.B.zero_(), self.Q.zero_()
        else:
            W = torch.zeros([L, wd, cmax], device=dev); B = torch.zeros([L, 1, cmax], device=dev)
            Q = torch.zeros([max(len(self.demod), 1), cmax, omax], device=dev)
        with torch.no_grad():
            for i, l in enumerate(layers):
                c = int(l.affine.weight.shape[0])
                g = 1.0 if isinstance(l, SynthesisLayer) else float(l.weight_gain)          # ToRGB: styles * weight_gain (:5602)
                W[i, :, :c] = (l.affine.weight * (l.affine.weight_gain * g)).t()
                B[i, 0, :c] = l.affine.bias * (l.affine.bias_gain * g)
            for j, i in enumerate(self.demod):
                w = layers[i].weight
                Q[j, :w.shape[1], :w.shape[0]] = w.square().sum(dim=[2, 3]).t()
        self.W, self.B, self.Q, self.key = W, B, Q, key
        if not same:
            self.demod_idx = torch.tensor(self.demod, device=dev, dtype=torch.long)
        self._layers = [weakref.ref(l) for l in layers]

    def refresh(self):
        """Rebuild the packed affine / demodulation matrices in place if any source parameter changed (TryOnSession.refresh_weights)."""
        layers = [r() for r in getattr(self, '_layers', [])]
        if layers and all(l is not None for l in layers):
            self._pack(layers)

    def fill(self, entries):
        """entries: [(layer, w [N, w_dim])] in any order; sets layer._pre = (styles [N, Cin], dcoefs [N, Cout] or None)."""
        layers = [l for l, _ in entries]
        self._pack(layers)
        X = torch.stack([w for _, w in entries], dim=0).to(torch.float32)                    # [L, N, w_dim]
        S = torch.baddbmm(self.B, X, self.W)                                                 # [L, N, Cmax]
        D = Sn = None
        if self.demod:
            Sd = S.index_select(0, self.demod_idx)
            D = torch.baddbmm(_EPS.get(X.device).reshape(1, 1, 1), Sd.square(), self.Q).rsqrt()
            # fp16 operand range: each sample's styles go to the kernel with unit inf-norm, the factor rides in the demodulation coefficient
            # (what conv2d_igemm does per call; batched here).  ToRGB layers keep raw styles: their fused kernel is fp32 arithmetic.
            smax = Sd.abs().amax(dim=2, keepdim=True).clamp_min(1e-20)
            Sn, D = Sd / smax, D * smax
        pos = {i: j for j, i in enumerate(self.demod)}
        for i, l in enumerate(layers):
            c = int(l.affine.weight.shape[0])
            if i in pos:
                l._pre = (Sn[pos[i], :, :c], D[pos[i], :, :int(l.weight.shape[0])], True)
            else:
                l._pre = (S[i, :, :c], None, False)

    @staticmethod
    def clear(entries):
        for l, _ in entries:
            l._pre = None


def _block_style_entries(blk, cur):
    """(layer, w) pairs of one synthesis block in the order its forward consumes ``cur`` (conv0 if present, conv1, torgb)."""
    out, it = [], 0
    if hasattr(blk, 'conv0'):
        out.append((blk.conv0, cur[:, it])); it += 1
    out.append((blk.conv1, cur[:, it])); it += 1
    if hasattr(blk, 'torgb'):
        out.append((blk.torgb, cur[:, it]))
    return out


def _style_bank_usable(module, ws):
    return (not torch.is_grad_enabled()) and ws.is_cuda and getattr(module.ops, 'modconv_layer', None) is not None and \
        os.environ.get('PASTA_B200_STYLE_BANK', '1') != '0'


# ----------------------------------------------------------------------------- layers


class FullyConnectedLayer(OpsModule):
    def __init__(self, in_features, out_features, bias=True, activation='linear', lr_multiplier=1, bias_init=0):
        super().__init__()
        self.activation = activation
        self.weight = nn.Parameter(torch.randn([out_features, in_features]) / lr_multiplier)
        self.bias = nn.Parameter(torch.full([out_features], np.float32(bias_init))) if bias else None
        self.weight_gain = lr_multiplier / np.sqrt(in_features)
        self.bias_gain = lr_multiplier

    def forward(self, x):
        w = self.weight.to(x.dtype) * self.weight_gain
        b = self.bias
        if b is not None:
            b = b.to(x.dtype)
            if self.bias_gain != 1:
                b = b * self.bias_gain
        if self.activation == 'linear' and b is not None:
            return torch.addmm(b.unsqueeze(0), x, w.t())
        return self.ops.bias_act(x.matmul(w.t()), b, act=self.activation)


class Conv2dLayer(OpsModule):
    """conv2d_resample + bias_act.  ``pre_act`` selects the SPADE variant (activation BEFORE the convolution,
    reference Spade_Conv2dLayer.forward :4342-4354); otherwise it is the plain Conv2dLayer (:170-179)."""

    def __init__(self, in_channels, out_channels, kernel_size, bias=True, activation='linear', up=1, down=1,
                 resample_filter=_FIR, conv_clamp=None, channels_last=False, trainable=True):
        super().__init__()
        self.activation, self.up, self.down, self.conv_clamp = activation, up, down, conv_clamp
        self.register_buffer('resample_filter', _fir_buffer(resample_filter))
        self.padding = kernel_size // 2
        self.weight_gain = 1 / np.sqrt(in_channels * (kernel_size ** 2))
        self.act_gain = _ACT_GAIN[activation]
        w = torch.randn([out_channels, in_channels, kernel_size, kernel_size])
        b = torch.zeros([out_channels]) if bias else None
        if trainable:
            self.weight = nn.Parameter(w)
            self.bias = nn.Parameter(b) if b is not None else None
        else:
            self.register_buffer('weight', w)
            if b is not None:
                self.register_buffer('bias', b)
            else:
                self.bias = None

    def _conv(self, x):
        w = self.weight * self.weight_gain
        return self.ops.conv2d_resample(x=x, w=w.to(x.dtype), f=self.resample_filter, up=self.up, down=self.down,
                                        padding=self.padding, flip_weight=(self.up == 1))

    def _act(self, x, gain):
        b = self.bias.to(x.dtype) if self.bias is not None else None
        clamp = self.conv_clamp * gain if self.conv_clamp is not None else None
        return self.ops.bias_act(x, b, act=self.activation, gain=self.act_gain * gain, clamp=clamp)

    def _fused(self, x, gain, pre_act, x2=None, residual=None, out_c8=False):
        """One-launch layer when the operator table offers it (the CUDA product does; the oracle table does not)."""
        layer = getattr(self.ops, 'conv_layer', None)
        if layer is None:
            return None
        w = self.weight                                   # raw parameter: weight_gain is applied when the GEMM tiles are packed
        b = self.bias.to(torch.float32 if x.ndim == 5 else x.dtype) if self.bias is not None else None
        clamp = self.conv_clamp * gain if self.conv_clamp is not None else None
        kw = dict(f=self.resample_filter, up=self.up, down=self.down, padding=self.padding, flip_weight=(self.up == 1),
                  w_scale=float(self.weight_gain), cache_weights=True, x2=x2, residual=residual)
        if out_c8:
            kw['out_c8'] = True
        if isinstance(self, SpadeConv2dLayer):
            kw['half_ok'] = True                          # SPADE convs may be fed the fp16 intermediates of the norm blocks
        if pre_act is None:
            return layer(x, w, b, act=self.activation, act_gain=self.act_gain * gain, clamp=clamp, **kw)
        if not pre_act:                                   # SPADE layer called with no_act=True: bare convolution
            return layer(x, w, None, **kw)
        if b is not None or clamp is not None or self.activation not in ('relu', 'lrelu'):
            return None
        return layer(x, w, None, in_act=self.activation, in_gain=self.act_gain * gain, **kw)

    def forward(self, x, gain=1, x2=None, residual=None, out_c8=False):
        """``x2``: second input concatenated along channels (reference: torch.cat before the call, :5705); ``residual``: added to the
        result (reference: ``y.add_(x)`` after the call, :990); ``out_c8``: channel-blocked fp16 result for a consumer that loads it by TMA."""
        y = self._fused(x, gain, None, x2, residual, out_c8)
        if y is not None:
            return y
        assert not out_c8 and x.ndim == 4
        if x2 is not None:
            x = torch.cat([x, x2.to(x.dtype)], dim=1)
        y = self._act(self._conv(x), gain)
        return y if residual is None else residual.add_(y)


class SpadeConv2dLayer(Conv2dLayer):
    def __init__(self, in_channels, out_channels, kernel_size, bias=True, activation='relu', **kw):
        super().__init__(in_channels, out_channels, kernel_size, bias=bias, activation=activation, **kw)

    def forward(self, x, gain=1, no_act=False, residual=None, out_c8=False):
        y = self._fused(x, gain, not no_act, None, residual, out_c8)
        if y is not None:
            return y
        assert not out_c8
        if not no_act:
            x = self._act(x, gain)
        y = self._conv(x)
        return y if residual is None else residual.add_(y)


class MappingNetwork(OpsModule):
    def __init__(self, z_dim, c_dim, w_dim, num_ws, num_layers=8, embed_features=None, layer_features=None,
                 activation='lrelu', lr_multiplier=0.01, w_avg_beta=0.995):
        super().__init__()
        self.z_dim, self.c_dim, self.w_dim, self.num_ws, self.num_layers, self.w_avg_beta = z_dim, c_dim, w_dim, num_ws, num_layers, w_avg_beta
        embed_features = w_dim if embed_features is None else embed_features
        if c_dim == 0:
            embed_features = 0
        layer_features = w_dim if layer_features is None else layer_features
        feats = [z_dim + embed_features] + [layer_features] * (num_layers - 1) + [w_dim]
        if c_dim > 0:
            self.embed = FullyConnectedLayer(c_dim, embed_features)
        for i in range(num_layers):
            setattr(self, f'fc{i}', FullyConnectedLayer(feats[i], feats[i + 1], activation=activation, lr_multiplier=lr_multiplier))
        if num_ws is not None and w_avg_beta is not None:
            self.register_buffer('w_avg', torch.zeros([w_dim]))

    def forward(self, z, c, truncation_psi=1, truncation_cutoff=None, skip_w_avg_update=False):
        x = None
        if self.z_dim > 0:
            misc.assert_shape(z, [None, self.z_dim])
            x = normalize_2nd_moment(z.to(torch.float32))
        if self.c_dim > 0:
            misc.assert_shape(c, [None, self.c_dim])
            y = normalize_2nd_moment(self.embed(c.to(torch.float32)))
            x = torch.cat([x, y], dim=1) if x is not None else y
        for i in range(self.num_layers):
            x = getattr(self, f'fc{i}')(x)
        if self.w_avg_beta is not None and self.training and not skip_w_avg_update:
            self.w_avg.copy_(x.detach().mean(dim=0).lerp(self.w_avg, self.w_avg_beta))
        if self.num_ws is not None:
            x = x.unsqueeze(1).repeat([1, self.num_ws, 1])
        if truncation_psi != 1:
            assert self.w_avg_beta is not None
            if self.num_ws is None or truncation_cutoff is None:
                x = self.w_avg.lerp(x, truncation_psi)
            else:
                x[:, :truncation_cutoff] = self.w_avg.lerp(x[:, :truncation_cutoff], truncation_psi)
        return x


class SynthesisLayer(OpsModule):
    def __init__(self, in_channels, out_channels, w_dim, resolution, kernel_size=3, up=1, use_noise=True, activation='lrelu',
                 resample_filter=_FIR, conv_clamp=None, channels_last=False):
        super().__init__()
        self.resolution, self.up, self.use_noise, self.activation, self.conv_clamp = resolution, up, use_noise, activation, conv_clamp
        self.register_buffer('resample_filter', _fir_buffer(resample_filter))
        self.padding = kernel_size // 2
        self.act_gain = _ACT_GAIN[activation]
        self.affine = FullyConnectedLayer(w_dim, in_channels, bias_init=1)
        self.weight = nn.Parameter(torch.randn([out_channels, in_channels, kernel_size, kernel_size]))
        if use_noise:
            self.register_buffer('noise_const', torch.randn([resolution, resolution]))
            self.noise_strength = nn.Parameter(torch.zeros([]))
        self.bias = nn.Parameter(torch.zeros([out_channels]))

    def forward(self, x, w, noise_mode='random', fused_modconv=True, gain=1, out_c8=False):
        assert noise_mode in ['random', 'const', 'none']
        if x.ndim == 5:                                   # channel-blocked fp16 [N, C/8, H, W, 8] from one of our own epilogues
            misc.assert_shape(x, [None, self.weight.shape[1] // 8, self.resolution // self.up, self.resolution // self.up, 8])
        else:
            misc.assert_shape(x, [None, self.weight.shape[1], self.resolution // self.up, self.resolution // self.up])
        pre = getattr(self, '_pre', None)
        styles, dcoefs, normalized = pre if pre is not None else (self.affine(w), None, False)
        noise = None
        if self.use_noise and noise_mode == 'random':
            noise = torch.randn([x.shape[0], 1, self.resolution, self.resolution], device=x.device) * self.noise_strength
        if self.use_noise and noise_mode == 'const':
            noise = self.noise_const * self.noise_strength
        clamp = self.conv_clamp * gain if self.conv_clamp is not None else None
        layer = getattr(self.ops, 'modconv_layer', None)
        if layer is not None:
            return layer(x, self.weight, styles, noise=noise, up=self.up, padding=self.padding, resample_filter=self.resample_filter,
                         flip_weight=(self.up == 1), fused_modconv=fused_modconv, bias=self.bias.to(torch.float32 if x.ndim == 5 else x.dtype), act=self.activation,
                         act_gain=self.act_gain * gain, clamp=clamp, dcoefs=dcoefs, styles_normalized=normalized, **(dict(out_c8=True) if out_c8 else {}))
        x = self.ops.modulated_conv2d(x=x, weight=self.weight, styles=styles, noise=noise, up=self.up, padding=self.padding,
                                      resample_filter=self.resample_filter, flip_weight=(self.up == 1), fused_modconv=fused_modconv)
        return self.ops.bias_act(x, self.bias.to(x.dtype), act=self.activation, gain=self.act_gain * gain, clamp=clamp)


class ToRGBLayerFull(OpsModule):
    """1x1 modulated conv (no demodulation) to RGB; the last block of the style branch also predicts the 6-class
    parsing map from the same styles."""

    def __init__(self, in_channels, out_channels, w_dim, kernel_size=1, conv_clamp=None, channels_last=False, is_last=False, is_style=False):
        super().__init__()
        self.conv_clamp = conv_clamp
        self.affine = FullyConnectedLayer(w_dim, in_channels, bias_init=1)
        self.weight = nn.Parameter(torch.randn([out_channels, in_channels, kernel_size, kernel_size]))
        self.bias = nn.Parameter(torch.zeros([out_channels]))
        self.weight_gain = 1 / np.sqrt(in_channels * (kernel_size ** 2))
        self.predicts_parsing = bool(is_last and is_style)
        if self.predicts_parsing:
            self.m_weight1 = nn.Parameter(torch.randn([6, in_channels, kernel_size, kernel_size]))
            self.m_bias1 = nn.Parameter(torch.zeros([6]))

    def forward_skip(self, x, w, img, resample_filter):
        """Fused skip path: returns (upsample2d(img) + rgb, parsing) or None when the operator table has no fused kernel for this shape."""
        fused = getattr(self.ops, 'torgb_skip', None)
        if fused is None:
            return None
        pre = getattr(self, '_pre', None)
        styles = pre[0] if pre is not None else self.affine(w) * self.weight_gain
        rgb = fused(x, self.weight, styles, self.bias, self.conv_clamp, img, resample_filter)
        if rgb is None:
            return None
        parsing = fused(x, self.m_weight1, styles, self.m_bias1, self.conv_clamp, None, None) if self.predicts_parsing else None
        return rgb, parsing

    def forward(self, x, w, fused_modconv=True):
        pre = getattr(self, '_pre', None)
        styles = pre[0] if pre is not None else self.affine(w) * self.weight_gain
        layer = getattr(self.ops, 'modconv_layer', None)

        def head(weight, bias):
            if layer is not None:
                return layer(x, weight, styles, demodulate=False, fused_modconv=fused_modconv, bias=bias.to(x.dtype), clamp=self.conv_clamp)
            y = self.ops.modulated_conv2d(x=x, weight=weight, styles=styles, demodulate=False, fused_modconv=fused_modconv)
            return self.ops.bias_act(y, bias.to(x.dtype), clamp=self.conv_clamp)

        parsing = head(self.m_weight1, self.m_bias1) if self.predicts_parsing else None
        return head(self.weight, self.bias), parsing


def _supported_c8_out(layer, x):
    """True when ``layer`` (a plain stride-1 Conv2dLayer) applied to the fp32 NCHW tensor ``x`` runs on the tcgen05 kernel, i.e. can write a
    channel-blocked result."""
    from .torch_utils.ops import conv_igemm as K
    pad = int(layer.padding)
    return x.ndim == 4 and layer.up == 1 and layer.down == 1 and int(layer.weight.shape[0]) % 16 == 0 and \
        K.supported(x, layer.weight, up=1, down=1, f=layer.resample_filter, padding=(pad,) * 4)


class ResBlock(OpsModule):
    def __init__(self, in_channels, out_channels, kernel_size, bias=True, activation='linear', up=1, down=1,
                 resample_filter=_FIR, conv_clamp=None, channels_last=False, trainable=True):
        super().__init__()
        self.register_buffer('resample_filter', _fir_buffer(resample_filter))
        self.conv0 = Conv2dLayer(in_channels, out_channels, kernel_size=3, activation=activation, up=up, down=down, bias=bias,
                                 resample_filter=resample_filter, conv_clamp=conv_clamp)
        self.conv1 = Conv2dLayer(out_channels, out_channels, kernel_size=3, activation=activation, bias=bias,
                                 resample_filter=resample_filter, conv_clamp=conv_clamp)
        self.skip = Conv2dLayer(in_channels, out_channels, kernel_size=1, bias=False, up=up, down=down,
                                resample_filter=resample_filter, conv_clamp=conv_clamp)

    def forward(self, x, out_c8=False):
        """Inference on the CUDA table: the tensors that only travel inside the block (conv0's result, and the skip branch when the input is already
        channel-blocked) are channel-blocked fp16 and feed the next convolution by TMA; ``out_c8`` asks for a channel-blocked result as well."""
        c8_ok = getattr(self.ops, 'c8_ok', None)
        cout = int(self.conv1.weight.shape[0])
        inner = False
        if c8_ok is not None and x.is_cuda and not torch.is_grad_enabled() and self.conv0.up == 1 and cout % 16 == 0:
            oh, ow = (int(x.shape[2]) // self.conv0.down, int(x.shape[3]) // self.conv0.down)
            inner = bool(c8_ok(cout, oh, ow, 3))
        if x.ndim == 5:
            # channel-blocked input: every convolution of the block loads by TMA -- stride 1, or (down-sampling block) the 3x3 and the 1x1 skip both
            # through the strided space-to-depth box, the 1x1 as a centre-tap 3x3
            assert inner and (self.conv0.down == 1 or c8_ok(int(x.shape[1]) * 8, int(x.shape[2]), int(x.shape[3]), 3, 1, 2)), \
                'a channel-blocked input needs the channel-blocked chain'
            y = self.skip(x, gain=np.sqrt(0.5), out_c8=True)
            return self.conv1(self.conv0(x, out_c8=True), gain=np.sqrt(0.5), residual=y, out_c8=out_c8)
        if inner and self.conv0.down == 1 and _supported_c8_out(self.skip, x):
            # the skip branch is only ever read back as conv1's residual: channel-blocked fp16 (two 16-byte loads per thread and chunk in the
            # epilogue instead of sixteen strided 4-byte loads)
            return self.conv1(self.conv0(x, out_c8=True), gain=np.sqrt(0.5), residual=self.skip(x, gain=np.sqrt(0.5), out_c8=True))
        y = self.skip(x, gain=np.sqrt(0.5))
        if inner:
            return self.conv1(self.conv0(x, out_c8=True), gain=np.sqrt(0.5), residual=y)
        return self.conv1(self.conv0(x), gain=np.sqrt(0.5), residual=y)          # y + conv1(...): the add rides in conv1's epilogue


class ConstEncoderNetwork(OpsModule):
    """Pose/retain encoder: 1x1 stem, then ``n_downsampling`` stride-2 3x3 convs (256^2 -> 4^2 x 512)."""

    def __init__(self, input_nc, output_nc, ngf=64, n_downsampling=4):
        super().__init__()
        mult_in, mult_out = [1, 2, 4, 4, 4, 8], [2, 4, 4, 4, 8, 8]
        layers = [Conv2dLayer(input_nc, ngf, kernel_size=1)]
        layers += [Conv2dLayer(ngf * mult_in[i], ngf * mult_out[i], kernel_size=3, down=2) for i in range(n_downsampling)]
        self.model = nn.Sequential(*layers)

    def forward(self, x):
        """Inference on the CUDA table: the tensors between the stem and the stride-2 convolutions travel channel-blocked fp16; each down-2 layer reads
        its space-to-depth planes through a strided TMA box (half the bytes of the fp32 NCHW tensor, no converter warps).  The last layer's
        result is dense fp32 as everywhere else."""
        c8_ok = getattr(self.ops, 'c8_ok', None)
        layers = list(self.model)
        if c8_ok is None or not x.is_cuda or torch.is_grad_enabled() or os.environ.get('PASTA_B200_C8_CHAIN', '1') == '0':
            return self.model(x)
        h, w = int(x.shape[2]), int(x.shape[3])
        for i, layer in enumerate(layers):
            nxt = layers[i + 1] if i + 1 < len(layers) else None
            oh, ow = h // layer.down, w // layer.down
            out_c8 = bool(nxt is not None and nxt.down == 2 and int(layer.weight.shape[0]) % 16 == 0 and
                          c8_ok(int(nxt.weight.shape[1]), oh, ow, int(nxt.weight.shape[2]), 1, 2) and
                          (x.ndim == 5 or _supported_c8_out(layer, x) or (layer.down == 2 and _down2_supported(layer, x))))
            x = layer(x, out_c8=out_c8)
            h, w = oh, ow
        return x


def _down2_supported(layer, x):
    from .torch_utils.ops import conv_igemm as K
    pad = int(layer.padding)
    return x.ndim == 4 and K.supported(x, layer.weight, up=1, down=2, f=layer.resample_filter, padding=(pad,) * 4)


class Dense(nn.Module):
    """Per-pixel Linear -> InstanceNorm -> LeakyReLU(0.01) (plain torch in the reference as well)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.bn = nn.InstanceNorm2d(out_channels)
        self.activation = nn.LeakyReLU()
        self.linear = nn.Linear(in_channels, out_channels)

    def forward(self, x):
        if x.is_cuda and x.dtype == torch.float32 and not torch.is_grad_enabled() and x.shape[2] * x.shape[3] <= 24576 and not self.bn.affine:
            # inference on the sm_100a kernels: the per-pixel Linear is a 1x1 convolution (tcgen05), InstanceNorm + LeakyReLU one streaming kernel
            from .torch_utils.ops import conv_igemm as K
            if K.enabled:
                w4 = getattr(self, '_w4', None)
                if w4 is None or w4.data_ptr() != self.linear.weight.data_ptr():
                    # a detached 4-D view of the Linear weight: shares storage and version counter, so the packed-weight store tracks updates
                    w4 = self.linear.weight.detach().view(self.out_channels, self.in_channels, 1, 1)
                    object.__setattr__(self, '_w4', w4)
                y = K.conv2d_igemm(x, w4, bias=self.linear.bias, cache_weights=True)
                return K.instance_norm_act(y, act='lrelu', alpha=float(self.activation.negative_slope), eps=float(self.bn.eps))
        y = self.linear(x.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)
        return self.activation(self.bn(y))


class StyleEncoderNetworkV16(OpsModule):
    """Patch-routed style encoder: the normalised garment patches (10 upper + 4 lower, RGB) -> 512-d style code; a second
    small encoder turns the retained-person image into 64-channel features at 256/128/64/32 px for the merge convs."""

    def __init__(self, input_nc, output_nc, ngf=64, n_downsampling=4):
        super().__init__()
        enc = [Conv2dLayer(input_nc, ngf, kernel_size=1)]
        for m_in, m_out in zip([1, 2, 4], [2, 4, 8]):
            enc += [Dense(ngf * m_in, ngf * m_in), Conv2dLayer(ngf * m_in, ngf * m_out, kernel_size=3, down=2)]
        for _ in range(3):
            enc += [Dense(ngf * 8, ngf * 8), Conv2dLayer(ngf * 8, ngf * 8, kernel_size=3)]
        enc += [nn.AdaptiveAvgPool2d(1)]
        self.model = nn.Sequential(*enc)
        self.fc = FullyConnectedLayer(output_nc, output_nc)
        feat = [Conv2dLayer(3, ngf, kernel_size=3)] + [Conv2dLayer(ngf, ngf, kernel_size=3, down=2) for _ in range(3)]
        self.feat_enc = nn.Sequential(*feat)

    def _retain_features(self, x):
        """feat_enc.  Inference on the CUDA table: the >= 128 px feature maps are only read by the merge convolutions of the channel-blocked synthesis
        blocks and by the next stride-2 layer, so they are produced channel-blocked fp16 straight away (no fp32 copy, no nchw -> c8 pass); the
        stride-2 layers read them through the strided TMA box.  Below 128 px the maps are dense fp32 as before."""
        c8_ok = getattr(self.ops, 'c8_ok', None)
        chain = c8_ok is not None and x.is_cuda and not torch.is_grad_enabled() and os.environ.get('PASTA_B200_C8_CHAIN', '1') != '0'
        feats = []
        h, w = int(x.shape[2]), int(x.shape[3])
        for layer in self.feat_enc:
            if not chain:
                x = layer(x)
                feats.append(x)
                continue
            cin, cout, k = int(layer.weight.shape[1]), int(layer.weight.shape[0]), int(layer.weight.shape[2])
            oh, ow = h // layer.down, w // layer.down
            if x.ndim == 5 and not (layer.down == 2 and c8_ok(cin, h, w, k, 1, 2)):
                x = _spade_to_nchw(x).float()
            want = oh >= 128 and cout % 16 == 0 and c8_ok(cout, oh, ow, 1)
            out_c8 = bool(want and (x.ndim == 5 or _supported_c8_out(layer, x) or (layer.down == 2 and _down2_supported(layer, x))))
            x = layer(x, out_c8=out_c8)
            feats.append(x)
            h, w = oh, ow
        return feats

    def forward(self, x, const_input):
        feats = self._retain_features(const_input)
        x = self.model(x)
        return self.fc(x.view(x.size(0), -1)), feats


class SpadeNormBlock(OpsModule):
    def __init__(self, in_channels, norm_channels):
        super().__init__()
        self.conv_mlp = SpadeConv2dLayer(in_channels, norm_channels, kernel_size=3, bias=False)
        self.conv_mlp_act = nn.ReLU()
        self.conv_gamma = SpadeConv2dLayer(norm_channels, norm_channels, kernel_size=3, bias=False)
        self.conv_beta = SpadeConv2dLayer(norm_channels, norm_channels, kernel_size=3, bias=False)
        self.param_free_norm = nn.InstanceNorm2d(norm_channels, affine=False)

    def forward(self, x, denorm_feats, post_act=None, stats=None, out_half=False, out_k=None):
        """``post_act = (name, gain)``: apply the pre-activation of the Spade conv that consumes the result here (then call it with
        ``no_act=True``).  ``stats``: (mean, rstd) of ``x`` when the caller already has them (two norm blocks of a SPADE res-block
        normalise the same tensor); only used on the fused path.  ``out_half``: the caller feeds the result straight into a bare Spade
        conv, so it may be an fp16 tensor when the operator table keeps fp16 intermediates; ``out_k``: that conv's kernel size -- when the table's
        ``c8_ok`` agrees, the result (and ``actv`` in between) travel channel-blocked and are loaded by TMA."""
        fused = getattr(self.ops, 'spade_conv_norm', None)
        if fused is not None and torch.is_grad_enabled() and (x.requires_grad or denorm_feats.requires_grad or self.conv_gamma.weight.requires_grad):
            fused = None                                  # training: the fused epilogue has no autograd -- decide before conv_mlp runs, not after (it ran twice)
        if fused is not None:
            half = getattr(self.ops, 'half_intermediates', None)
            half = torch.float16 if (half is not None and half(x)) else None
            c8_ok = getattr(self.ops, 'c8_ok', None)
            h, w = int(x.shape[2]), int(x.shape[3])
            actv_c8 = c8_ok is not None and c8_ok(self.conv_gamma.weight.shape[1], h, w, self.conv_gamma.weight.shape[2])
            out_c8 = bool(c8_ok is not None and post_act is not None and out_half and out_k is not None and c8_ok(x.shape[1], h, w, out_k))
            actv = self.ops.conv_layer(denorm_feats, self.conv_mlp.weight, None, padding=1, act='relu', act_gain=1.0,
                                       w_scale=float(self.conv_mlp.weight_gain), cache_weights=True, out_dtype=half, half_ok=True, out_c8=actv_c8)   # consumed by the next launch only
            y = fused(x, actv, self.conv_gamma.weight, self.conv_beta.weight, float(self.conv_gamma.weight_gain), post_act, stats,
                      half if (post_act is not None and out_half) else None, out_c8)
            if y is not None:
                return y
            actv = self._to_nchw(actv).float()
            denorm_feats = self._to_nchw(denorm_feats)
        actv = self.conv_mlp_act(self.conv_mlp(self._to_nchw(denorm_feats).to(x.dtype), no_act=True))
        gamma = self.conv_gamma(actv, no_act=True)
        beta = self.conv_beta(actv, no_act=True)
        y = self.param_free_norm(x) * (1 + gamma) + beta
        if post_act is not None:
            y = self.ops.bias_act(y, None, act=post_act[0], gain=post_act[1])
        return y


def _spade_to_nchw(t):
    if t.ndim == 5:
        from .torch_utils.ops import conv_igemm as K
        return K.from_c8(t, dtype=torch.float16)
    return t


SpadeNormBlock._to_nchw = staticmethod(_spade_to_nchw)


class SpadeResBlockV2(OpsModule):
    def __init__(self, in_channels, out_channels, resample_filter=_FIR, conv_clamp=None, resolution=128):
        super().__init__()
        self.register_buffer('resample_filter', _fir_buffer(resample_filter))
        kw = dict(bias=False, resample_filter=resample_filter, conv_clamp=conv_clamp)
        self.conv = SpadeConv2dLayer(in_channels, in_channels, kernel_size=3, **kw)
        self.conv0 = SpadeConv2dLayer(in_channels, out_channels, kernel_size=3, **kw)
        self.conv1 = SpadeConv2dLayer(out_channels, out_channels, kernel_size=3, **kw)
        self.skip = SpadeConv2dLayer(in_channels, out_channels, kernel_size=1, **kw)
        feat_channels = 128 * 2 if resolution == 128 else 64 * 2
        self.spade_skip = SpadeNormBlock(feat_channels, in_channels)
        self.spade0 = SpadeNormBlock(feat_channels, in_channels)
        self.spade1 = SpadeNormBlock(feat_channels, out_channels)

    def forward(self, x, denorm_feat, out_c8=False):
        """``out_c8``: the caller's consumer loads a channel-blocked tensor (the next SPADE block's first convolution, or the up-2 convolution of the
        texture block): the block result then leaves conv1's epilogue as channel-blocked fp16 -- half the bytes written, and the consumer runs on
        the TMA operand path instead of the fp32 converter path."""
        x = self.conv(x, no_act=True)
        # the pre-activation (relu * act_gain * gain) of each consuming Spade conv is handed to the norm block, which applies it in the
        # same pass (fused: in the GEMM epilogue that produces gamma / beta); the convs then run bare
        pre = lambda conv, gain: (conv.activation, float(conv.act_gain * gain))
        stats_fn = getattr(self.ops, 'instance_stats', None)
        stats = stats_fn(x) if stats_fn is not None and not (torch.is_grad_enabled() and x.requires_grad) else None   # shared by spade_skip / spade0
        s_in = self.spade_skip(x, denorm_feat, post_act=pre(self.skip, np.sqrt(0.5)), stats=stats, out_half=True, out_k=1)
        # inside a channel-blocked chain the skip branch (read back only as conv1's residual) stays channel-blocked fp16 too
        y = self.skip(s_in, no_act=True, out_c8=(s_in.ndim == 5 and int(self.skip.weight.shape[0]) % 16 == 0))
        x = self.conv0(self.spade0(x, denorm_feat, post_act=pre(self.conv0, 1), stats=stats, out_half=True, out_k=3), no_act=True)
        t = self.spade1(x, denorm_feat, post_act=pre(self.conv1, np.sqrt(0.5)), out_half=True, out_k=3)
        if y.ndim == 5 and t.ndim != 5:
            y = _spade_to_nchw(y).float()                 # conv1 is not on the TMA path: its residual must be plain fp32
        return self.conv1(t, no_act=True, residual=y, out_c8=bool(out_c8 and t.ndim == 5 and y.ndim == 5))


def _c8_chain_ok(block, x, cat_feat):
    """Channel-blocked chain through a synthesis block: only under inference on the CUDA operator table, at >= 128 px with <= 128 output channels, when
    the caller provided the channel-blocked retain-person features and every layer's shape is TMA-loadable."""
    c8_ok = getattr(block.ops, 'c8_ok', None)
    res, cout = block.resolution, int(block.conv1.weight.shape[0])
    if c8_ok is None or torch.is_grad_enabled() or not x.is_cuda or res < 128 or cout > 128 or cout % 16 or f'{res}_c8' not in cat_feat:
        return False
    if os.environ.get('PASTA_B200_C8_CHAIN', '1') == '0' or getattr(block.ops, 'torgb_skip', None) is None:
        return False
    return bool(c8_ok(cout, res, res, 3) and c8_ok(cout, res, res, 1) and (x.ndim == 4 or c8_ok(int(x.shape[1]) * 8, res // 2, res // 2, 3, 2)))


def _with_c8_cat_feats(module, cat_feat):
    """Channel-blocked copies ('<res>_c8') of the retain-person feature maps at >= 128 px, read by the merge convs of the channel-blocked blocks."""
    c8_ok = getattr(module.ops, 'c8_ok', None)
    if c8_ok is None or torch.is_grad_enabled() or os.environ.get('PASTA_B200_C8_CHAIN', '1') == '0':
        return cat_feat
    from .torch_utils.ops import conv_igemm as K
    cat_feat = dict(cat_feat)
    for key in [k for k in cat_feat if k.isdigit() and int(k) >= 128]:
        t = cat_feat[key]
        if t.ndim == 5:                                   # already channel-blocked (StyleEncoderNetworkV16._retain_features)
            cat_feat[key + '_c8'] = t
            cat_feat[key] = _LazyDense(t)
        elif t.is_cuda and t.ndim == 4 and t.shape[1] % 16 == 0 and c8_ok(t.shape[1], t.shape[2], t.shape[3], 1):
            cat_feat[key + '_c8'] = K.to_c8(t)
    return cat_feat


class _LazyDense:
    """A channel-blocked feature map standing in for its dense form: converted only if a consumer outside the channel-blocked chain asks for it."""

    def __init__(self, t):
        self.t = t

    def to(self, dtype):
        from .torch_utils.ops import conv_igemm as K
        return K.from_c8(self.t, dtype=torch.float16).to(dtype)


class SynthesisBlockFull(OpsModule):
    def __init__(self, in_channels, out_channels, w_dim, resolution, img_channels, is_last, is_style=False, architecture='skip',
                 resample_filter=_FIR, conv_clamp=None, use_fp16=False, fp16_channels_last=False, **layer_kwargs):
        assert architecture in ['orig', 'skip', 'resnet']
        super().__init__()
        self.in_channels, self.w_dim, self.resolution, self.img_channels = in_channels, w_dim, resolution, img_channels
        self.is_last, self.architecture = is_last, architecture
        self.register_buffer('resample_filter', _fir_buffer(resample_filter))
        self.num_conv = self.num_torgb = 0
        if in_channels == 0:
            self.const = nn.Parameter(torch.randn([out_channels, resolution, resolution]))   # kept for state_dict parity; unused
        else:
            self.conv0 = SynthesisLayer(in_channels, out_channels, w_dim=w_dim, resolution=resolution, up=2,
                                        resample_filter=resample_filter, conv_clamp=conv_clamp, **layer_kwargs)
            self.num_conv += 1
        self.conv1 = SynthesisLayer(out_channels, out_channels, w_dim=w_dim, resolution=resolution, conv_clamp=conv_clamp, **layer_kwargs)
        self.num_conv += 1
        if is_last or architecture == 'skip':
            self.torgb = ToRGBLayerFull(out_channels, img_channels, w_dim=w_dim, conv_clamp=conv_clamp, is_last=is_last, is_style=is_style)
            self.num_torgb += 1
        if in_channels != 0 and architecture == 'resnet':
            self.skip = Conv2dLayer(in_channels, out_channels, kernel_size=1, bias=False, up=2, resample_filter=resample_filter)
        if resolution > 16:
            self.merge_conv = Conv2dLayer(out_channels + 64, out_channels, kernel_size=1, resample_filter=resample_filter)

    def forward(self, x, img, ws, pose_feature, cat_feat, force_fp32=False, fused_modconv=None, **layer_kwargs):
        misc.assert_shape(ws, [None, self.num_conv + self.num_torgb, self.w_dim])
        w_iter = iter(ws.unbind(dim=1))
        if fused_modconv is None:
            fused_modconv = not self.training            # the generator always runs fp32 (reference :5747-5748, :5818)
        if self.in_channels == 0:
            x = pose_feature.to(torch.float32)
            x = self.conv1(x, next(w_iter), fused_modconv=fused_modconv, **layer_kwargs)
        elif self.architecture == 'resnet':
            y = self.skip(x, gain=np.sqrt(0.5))
            x = self.conv0(x, next(w_iter), fused_modconv=fused_modconv, **layer_kwargs)
            x = self.conv1(x, next(w_iter), fused_modconv=fused_modconv, gain=np.sqrt(0.5), **layer_kwargs)
            x = y.add_(x)
        else:
            # high-resolution blocks (>= 128 px, <= 128 channels) under inference: activations travel between conv0 / conv1 / merge_conv / ToRGB as
            # channel-blocked fp16 -- half the HBM bytes, TMA operand loads, styles folded into per-sample weights (a few MB at these widths)
            chain = _c8_chain_ok(self, x, cat_feat)
            if x.ndim == 4:
                misc.assert_shape(x, [None, self.in_channels, self.resolution // 2, self.resolution // 2])
                x = x.to(torch.float32)
            c8kw = dict(out_c8=True) if chain else {}
            x = self.conv0(x, next(w_iter), fused_modconv=fused_modconv, **layer_kwargs, **c8kw)
            x = self.conv1(x, next(w_iter), fused_modconv=fused_modconv, **layer_kwargs, **c8kw)
            if x.shape[2] > 16:                          # merge the warped retain-person features
                if chain:
                    x = self.merge_conv(x, x2=cat_feat[f'{x.shape[2]}_c8'], out_c8=True)
                else:
                    x = self.merge_conv(x, x2=cat_feat[str(x.shape[2])].to(torch.float32))     # concat fused into the 1x1 conv's operand loader
        if img is not None:
            misc.assert_shape(img, [None, self.img_channels, self.resolution // 2, self.resolution // 2])
        parsing = None
        if self.is_last or self.architecture == 'skip':
            w_rgb = next(w_iter)
            fused = self.torgb.forward_skip(x, w_rgb, img, self.resample_filter)
            if fused is not None:                           # one streaming kernel: upsample2d(img) + clamp(1x1 modconv + b)
                img, parsing = fused
            else:
                assert x.ndim == 4, 'channel-blocked activations need the fused ToRGB kernel'
                if img is not None:
                    img = self.ops.upsample2d(img, self.resample_filter)
                y, parsing = self.torgb(x, w_rgb, fused_modconv=fused_modconv)
                y = y.to(dtype=torch.float32)
                img = img.add_(y) if img is not None else y
        elif img is not None:
            img = self.ops.upsample2d(img, self.resample_filter)
        return x, img, parsing


class SynthesisNetworkFull(OpsModule):
    def __init__(self, w_dim, img_resolution, img_channels, channel_base=32768, channel_max=512, num_fp16_res=0, **block_kwargs):
        assert img_resolution >= 4 and img_resolution & (img_resolution - 1) == 0
        super().__init__()
        self.w_dim, self.img_resolution, self.img_channels = w_dim, img_resolution, img_channels
        self.img_resolution_log2 = int(np.log2(img_resolution))
        self.block_resolutions = [2 ** i for i in range(2, self.img_resolution_log2 + 1)]
        ch = {res: min(channel_base // res, channel_max) for res in self.block_resolutions}
        self.num_ws = 0
        for res in self.block_resolutions:
            block = SynthesisBlockFull(ch[res // 2] if res > 4 else 0, ch[res], w_dim=w_dim, resolution=res, img_channels=img_channels,
                                       is_last=(res == img_resolution), is_style=True, **block_kwargs)
            self.num_ws += block.num_conv + (block.num_torgb if res == img_resolution else 0)
            setattr(self, f'b{res}', block)
        mid, top = self.block_resolutions[-2], self.block_resolutions[-1]
        for k in (1, 2, 3):
            setattr(self, f'spade_b128_{k}', SpadeResBlockV2(ch[mid], ch[mid]))
        self.texture_b256 = SynthesisBlockFull(ch[top // 2], ch[top], w_dim=w_dim, resolution=top, img_channels=img_channels,
                                               is_last=True, is_style=False, **block_kwargs)
        ngf = 64
        self.spade_encoder = nn.Sequential(Conv2dLayer(3, ngf, kernel_size=7, activation='relu'),
                                           ResBlock(ngf, ngf, kernel_size=4, activation='relu'),
                                           ResBlock(ngf, ngf * 2, kernel_size=4, activation='relu', down=2))

    def encode_garment(self, x):
        """spade_encoder (7x7 conv + two ResBlocks, reference :5768-5771).  Inference on the CUDA table: the 7x7 layer hands the first ResBlock a
        channel-blocked fp16 tensor, so that block's three convolutions load their operand by TMA and move half the bytes."""
        enc = self.spade_encoder
        c8_ok = getattr(self.ops, 'c8_ok', None)
        c0 = int(enc[0].weight.shape[0])
        if c8_ok is not None and x.is_cuda and not torch.is_grad_enabled() and os.environ.get('PASTA_B200_C8_CHAIN', '1') != '0' and \
                c0 % 16 == 0 and c8_ok(c0, x.shape[2], x.shape[3], 3) and c8_ok(c0, x.shape[2], x.shape[3], 1):
            down_c8 = bool(c8_ok(c0, x.shape[2], x.shape[3], 3, 1, 2))     # the down-sampling block reads its input through the strided TMA box
            return enc[2](enc[1](enc[0](x, out_c8=True), out_c8=down_c8))
        return enc(x)

    def get_spade_feat(self, mask_256, denorm_mask, denorm_input, out=None, feat=None):
        """Garment features at 128 px; pixels the predicted mask covers but the source garment does not are filled with the
        garment's mean feature (reference :5777-5800).  ``out``: channel slice of the concatenated upper|lower tensor to write into
        (fused path: two streaming kernels instead of five elementwise passes and the torch.cat)."""
        half = lambda t: torch.nn.functional.interpolate(t, scale_factor=0.5)
        binar = lambda t: (t > 0.9).to(mask_256.dtype)
        mask_256 = binar(mask_256)
        mask_128 = binar(half(mask_256))
        denorm_mask_128 = binar(half(denorm_mask))
        valid = ((mask_128 + denorm_mask_128) == 2.0).to(mask_256.dtype)
        rest = mask_128 - valid
        feat = self.encode_garment(denorm_input * mask_256 - (1 - mask_256)) if feat is None else feat
        fused = getattr(self.ops, 'masked_mean_fill', None)
        if fused is not None and out is not None:
            y = fused(feat, valid, rest, out)
            if y is not None:
                return y
        assert not isinstance(out, tuple), 'channel-blocked garment features need the fused fill'

        feat_sum = (feat * valid).sum(dim=(2, 3), keepdim=True)
        count = valid.sum(dim=(2, 3), keepdim=True)
        enough = (count > 10).to(mask_256.dtype)
        count = count * enough + (128 * 128) * (1 - enough)
        y = feat * (1 - rest) + (feat_sum / count) * rest
        if out is not None:
            out.copy_(y)
            return out
        return y

    def forward(self, ws, pose_feat, cat_feat, denorm_upper_input, denorm_lower_input, denorm_upper_mask, denorm_lower_mask, label_override=None,
                **block_kwargs):
        """``label_override`` ([N, H, W] integer class map): use these labels instead of argmax(pred_parsing) for the garment masks of the fine stage
        (a test hook: the fine image is a discontinuous function of the logits, reference :5823-5826, so parity tests pin the labels)."""
        misc.assert_shape(ws, [None, self.num_ws, self.w_dim])
        ws = ws.to(torch.float32)
        block_ws, idx = [], 0
        for res in self.block_resolutions:
            block = getattr(self, f'b{res}')
            block_ws.append(ws.narrow(1, idx, block.num_conv + block.num_torgb))
            idx += block.num_conv
        entries = []
        if _style_bank_usable(self, ws):
            for blk, cur in list(zip([getattr(self, f'b{res}') for res in self.block_resolutions], block_ws)) + [(self.texture_b256, block_ws[-1])]:
                entries += _block_style_entries(blk, cur)
            if not hasattr(self, '_bank'):
                object.__setattr__(self, '_bank', StyleBank())
            self._bank.fill(entries)
        try:
            return self._forward_blocks(block_ws, pose_feat, cat_feat, denorm_upper_input, denorm_lower_input, denorm_upper_mask, denorm_lower_mask,
                                        label_override=label_override, **block_kwargs)
        finally:
            StyleBank.clear(entries)

    def _forward_blocks(self, block_ws, pose_feat, cat_feat, denorm_upper_input, denorm_lower_input, denorm_upper_mask, denorm_lower_mask,
                        label_override=None, **block_kwargs):
        x = img = parsing = None
        cat_feat = _with_c8_cat_feats(self, cat_feat)
        for res, cur in zip(self.block_resolutions, block_ws):
            x, img, parsing = getattr(self, f'b{res}')(x, img, cur, pose_feat, cat_feat, force_fp32=True, **block_kwargs)
            if res == 128:
                # reference :5820 clones both; none of the layers of this mirror writes into its input (every conv / ToRGB result is a fresh tensor),
                # so inference keeps the aliases and saves a 134 MB copy
                x_128, img_128 = (x, img) if not torch.is_grad_enabled() else (x.clone(), img.clone())
        label = torch.argmax(torch.softmax(parsing.detach(), dim=1), dim=1)[:, None].float()
        if label_override is not None:
            label = label_override.to(parsing.device)[:, None].float()
        cf = self.spade_encoder[-1].conv1.weight.shape[0]                 # channels of one garment's feature map (128)
        half = getattr(self.ops, 'half_intermediates', None)
        feat_dtype = torch.float16 if (half is not None and label.is_cuda and half(label[:, :, ::2, ::2])) else torch.float32   # read by conv_mlp only
        fh, fw = label.shape[2] // 2, label.shape[3] // 2
        c8_ok = getattr(self.ops, 'c8_ok', None)
        m_up, m_lo = (label == 1).float(), (label == 2).float()
        f_up = f_lo = None
        if label.is_cuda and not torch.is_grad_enabled():
            # one encoder pass over both garments (batch 2N): same weights, half the launches, fuller waves
            binar = lambda t: (t > 0.9).to(t.dtype)
            both = torch.cat([denorm_upper_input * binar(m_up) - (1 - binar(m_up)), denorm_lower_input * binar(m_lo) - (1 - binar(m_lo))], dim=0)
            f_up, f_lo = self.encode_garment(both).chunk(2, dim=0)
        if c8_ok is not None and label.is_cuda and feat_dtype == torch.float16 and cf % 8 == 0 and c8_ok(2 * cf, fh, fw, 3):
            # channel-blocked fp16: the nine conv_mlp convolutions that read this tensor load it by TMA
            spade_feat = torch.empty([label.shape[0], 2 * cf // 8, fh, fw, 8], dtype=torch.float16, device=label.device)
            self.get_spade_feat(m_up, denorm_upper_mask, denorm_upper_input, out=(spade_feat, 0), feat=f_up)
            self.get_spade_feat(m_lo, denorm_lower_mask, denorm_lower_input, out=(spade_feat, cf // 8), feat=f_lo)
        else:
            spade_feat = torch.empty([label.shape[0], 2 * cf, fh, fw], dtype=feat_dtype, device=label.device)
            self.get_spade_feat(m_up, denorm_upper_mask, denorm_upper_input, out=spade_feat[:, :cf], feat=f_up)      # upper | lower (:5831)
            self.get_spade_feat(m_lo, denorm_lower_mask, denorm_lower_input, out=spade_feat[:, cf:], feat=f_lo)
        x = x_128
        # the trunk between the SPADE blocks (and into the texture block) stays channel-blocked when the consumers load it by TMA
        n_, c_ = int(x.shape[0]), int(self.spade_b128_3.conv1.weight.shape[0])
        probe = types.SimpleNamespace(is_cuda=x.is_cuda, ndim=5, shape=(n_, c_ // 8, int(x.shape[2]), int(x.shape[3]), 8))
        trunk_c8 = c8_ok is not None and x.is_cuda and not torch.is_grad_enabled() and c_ % 16 == 0 and os.environ.get('PASTA_B200_C8_CHAIN', '1') != '0' and \
            bool(c8_ok(c_, int(x.shape[2]), int(x.shape[3]), 3))
        tex_c8 = trunk_c8 and _c8_chain_ok(self.texture_b256, probe, cat_feat)
        for k in (1, 2, 3):
            x = getattr(self, f'spade_b128_{k}')(x, spade_feat, out_c8=(trunk_c8 if k < 3 else tex_c8))
        _, finetune_img, _ = self.texture_b256(x, img_128, block_ws[-1], pose_feat, cat_feat, force_fp32=True, **block_kwargs)
        return img, finetune_img, parsing


class GeneratorFull(OpsModule):
    """The 256x192 full-body try-on generator (train_wo_flow_fullbody.py:190-201 builds exactly this)."""

    def __init__(self, z_dim, c_dim, w_dim, img_resolution, img_channels, mapping_kwargs={}, synthesis_kwargs={}):
        super().__init__()
        self.z_dim, self.c_dim, self.w_dim, self.img_resolution, self.img_channels = z_dim, c_dim, w_dim, img_resolution, img_channels
        self.synthesis = SynthesisNetworkFull(w_dim=w_dim, img_resolution=img_resolution, img_channels=img_channels, **synthesis_kwargs)
        self.num_ws = self.synthesis.num_ws
        self.mapping = MappingNetwork(z_dim=z_dim, c_dim=c_dim, w_dim=w_dim, num_ws=self.num_ws, **mapping_kwargs)
        self.const_encoding = ConstEncoderNetwork(input_nc=3 + 3, output_nc=512, ngf=64, n_downsampling=6)
        self.style_encoding = StyleEncoderNetworkV16(input_nc=(10 * 3 + 4 * 3), output_nc=512, ngf=64, n_downsampling=6)

    def forward(self, z, c, retain, pose, denorm_upper_input, denorm_lower_input, denorm_upper_mask, denorm_lower_mask,
                truncation_psi=1, truncation_cutoff=None, **synthesis_kwargs):
        pose_feat = self.const_encoding(pose)
        stylecode, feats = self.style_encoding(c, retain)
        ws = self.mapping(z, stylecode, truncation_psi=truncation_psi, truncation_cutoff=truncation_cutoff)
        cat_feats = {str(f.shape[2]): f for f in feats}
        return self.synthesis(ws, pose_feat, cat_feats, denorm_upper_input, denorm_lower_input, denorm_upper_mask, denorm_lower_mask,
                              **synthesis_kwargs)


# ----------------------------------------------------------------------------- 512 x 512 generator (the only 512-px network in the reference tree)


class ToRGBLayer(OpsModule):
    """Plain ToRGB (reference :320-336): modulated 1x1 conv without demodulation + bias + clamp."""

    def __init__(self, in_channels, out_channels, w_dim, kernel_size=1, conv_clamp=None, channels_last=False):
        super().__init__()
        self.conv_clamp = conv_clamp
        self.affine = FullyConnectedLayer(w_dim, in_channels, bias_init=1)
        self.weight = nn.Parameter(torch.randn([out_channels, in_channels, kernel_size, kernel_size]))
        self.bias = nn.Parameter(torch.zeros([out_channels]))
        self.weight_gain = 1 / np.sqrt(in_channels * (kernel_size ** 2))
        self.predicts_parsing = False

    forward_skip = ToRGBLayerFull.forward_skip

    def forward(self, x, w, fused_modconv=True):
        pre = getattr(self, '_pre', None)
        styles = pre[0] if pre is not None else self.affine(w) * self.weight_gain
        layer = getattr(self.ops, 'modconv_layer', None)
        if layer is not None:
            return layer(x, self.weight, styles, demodulate=False, fused_modconv=fused_modconv, bias=self.bias.to(x.dtype), clamp=self.conv_clamp)
        y = self.ops.modulated_conv2d(x=x, weight=self.weight, styles=styles, demodulate=False, fused_modconv=fused_modconv)
        return self.ops.bias_act(y, self.bias.to(x.dtype), clamp=self.conv_clamp)


class SynthesisBlock512(OpsModule):
    """reference SynthesisBlockV_512 :3578-3677 (skip architecture; retain-person features merged above 32 px)."""

    def __init__(self, in_channels, out_channels, w_dim, resolution, img_channels, is_last, architecture='skip', resample_filter=_FIR,
                 conv_clamp=None, use_fp16=False, fp16_channels_last=False, **layer_kwargs):
        assert architecture == 'skip'
        super().__init__()
        self.in_channels, self.w_dim, self.resolution, self.img_channels, self.is_last = in_channels, w_dim, resolution, img_channels, is_last
        self.register_buffer('resample_filter', _fir_buffer(resample_filter))
        self.num_conv = self.num_torgb = 0
        if in_channels == 0:
            self.const = nn.Parameter(torch.randn([out_channels, resolution, resolution]))      # state_dict parity; unused
        else:
            self.conv0 = SynthesisLayer(in_channels, out_channels, w_dim=w_dim, resolution=resolution, up=2, resample_filter=resample_filter,
                                        conv_clamp=conv_clamp, **layer_kwargs)
            self.num_conv += 1
        self.conv1 = SynthesisLayer(out_channels, out_channels, w_dim=w_dim, resolution=resolution, conv_clamp=conv_clamp, **layer_kwargs)
        self.num_conv += 1
        self.torgb = ToRGBLayer(out_channels, img_channels, w_dim=w_dim, conv_clamp=conv_clamp)
        self.num_torgb += 1
        self.merge_conv = Conv2dLayer(out_channels + 64, out_channels, kernel_size=1, resample_filter=resample_filter)

    def forward(self, x, img, ws, pose_feature, cat_feat, fused_modconv=None, **layer_kwargs):
        misc.assert_shape(ws, [None, self.num_conv + self.num_torgb, self.w_dim])
        w_iter = iter(ws.unbind(dim=1))
        if fused_modconv is None:
            fused_modconv = not self.training
        if self.in_channels == 0:
            x = self.conv1(pose_feature.to(torch.float32), next(w_iter), fused_modconv=fused_modconv, **layer_kwargs)
        else:
            chain = _c8_chain_ok(self, x, cat_feat)            # channel-blocked fp16 chain, as in SynthesisBlockFull
            if x.ndim == 4:
                x = x.to(torch.float32)
            c8kw = dict(out_c8=True) if chain else {}
            x = self.conv0(x, next(w_iter), fused_modconv=fused_modconv, **layer_kwargs, **c8kw)
            x = self.conv1(x, next(w_iter), fused_modconv=fused_modconv, **layer_kwargs, **c8kw)
            if x.shape[2] > 32:
                if chain:
                    x = self.merge_conv(x, x2=cat_feat[f'{x.shape[2]}_c8'], out_c8=True)
                else:
                    x = self.merge_conv(x, x2=cat_feat[str(x.shape[2])].to(torch.float32))     # concat fused into the 1x1 conv's operand loader
        w_rgb = next(w_iter)
        fused = self.torgb.forward_skip(x, w_rgb, img, self.resample_filter)
        if fused is not None:
            img = fused[0]
        else:
            assert x.ndim == 4, 'channel-blocked activations need the fused ToRGB kernel'
            if img is not None:
                img = self.ops.upsample2d(img, self.resample_filter)
            y = self.torgb(x, w_rgb, fused_modconv=fused_modconv).to(torch.float32)
            img = img.add_(y) if img is not None else y
        return x, img


class SynthesisNetwork512(OpsModule):
    def __init__(self, w_dim, img_resolution, img_channels, channel_base=32768, channel_max=512, num_fp16_res=0, **block_kwargs):
        assert img_resolution >= 8 and img_resolution & (img_resolution - 1) == 0
        super().__init__()
        self.w_dim, self.img_resolution, self.img_channels = w_dim, img_resolution, img_channels
        self.block_resolutions = [2 ** i for i in range(3, int(np.log2(img_resolution)) + 1)]
        ch = {res: min(channel_base // res, channel_max) for res in self.block_resolutions}
        self.num_ws = 0
        for res in self.block_resolutions:
            block = SynthesisBlock512(ch[res // 2] if res > 8 else 0, ch[res], w_dim=w_dim, resolution=res, img_channels=img_channels,
                                      is_last=(res == img_resolution), **block_kwargs)
            self.num_ws += block.num_conv + (block.num_torgb if res == img_resolution else 0)
            setattr(self, f'b{res}', block)

    def forward(self, ws, pose_feat, cat_feat, **block_kwargs):
        misc.assert_shape(ws, [None, self.num_ws, self.w_dim])
        ws = ws.to(torch.float32)
        x = img = None
        idx, plan = 0, []
        for res in self.block_resolutions:
            block = getattr(self, f'b{res}')
            plan.append((block, ws.narrow(1, idx, block.num_conv + block.num_torgb)))
            idx += block.num_conv
        entries = []
        if _style_bank_usable(self, ws):
            for block, cur in plan:
                entries += _block_style_entries(block, cur)
            if not hasattr(self, '_bank'):
                object.__setattr__(self, '_bank', StyleBank())
            self._bank.fill(entries)
        cat_feat = _with_c8_cat_feats(self, cat_feat)
        try:
            for block, cur in plan:
                x, img = block(x, img, cur, pose_feat, cat_feat, **block_kwargs)
        finally:
            StyleBank.clear(entries)
        return img


class StyleEncoderNetwork512(OpsModule):
    def __init__(self, input_nc, output_nc, ngf=64, n_downsampling=4):
        super().__init__()
        enc = [Conv2dLayer(input_nc, ngf, kernel_size=1)]
        for m_in, m_out in zip([1, 2, 4], [2, 4, 8]):
            enc += [Dense(ngf * m_in, ngf * m_in), Conv2dLayer(ngf * m_in, ngf * m_out, kernel_size=3, down=2)]
        enc += [nn.AdaptiveAvgPool2d(1)]
        self.model = nn.Sequential(*enc)
        self.fc = FullyConnectedLayer(output_nc, output_nc)
        self.feat_enc = nn.Sequential(Conv2dLayer(3, ngf, kernel_size=3), *[Conv2dLayer(ngf, ngf, kernel_size=3, down=2) for _ in range(3)])

    _retain_features = StyleEncoderNetworkV16._retain_features

    def forward(self, x, const_input):
        feats = self._retain_features(const_input)
        x = self.model(x)
        return self.fc(x.view(x.size(0), -1)), feats


class Generator512(OpsModule):
    """reference Generator_512 :3782-3815 — the only 512-px generator in the source tree (the released 512x320 pickle is not available;
    BASELINE configs[2] therefore uses this network, as SURVEY.md §8(d) states)."""

    def __init__(self, z_dim, c_dim, w_dim, img_resolution, img_channels, mapping_kwargs={}, synthesis_kwargs={}):
        super().__init__()
        self.z_dim, self.c_dim, self.w_dim, self.img_resolution, self.img_channels = z_dim, c_dim, w_dim, img_resolution, img_channels
        self.synthesis = SynthesisNetwork512(w_dim=w_dim, img_resolution=img_resolution, img_channels=img_channels, **synthesis_kwargs)
        self.num_ws = self.synthesis.num_ws
        self.mapping = MappingNetwork(z_dim=z_dim, c_dim=c_dim, w_dim=w_dim, num_ws=self.num_ws, **mapping_kwargs)
        self.const_encoding = ConstEncoderNetwork(input_nc=3 + 3, output_nc=512, ngf=64, n_downsampling=6)
        self.style_encoding = StyleEncoderNetwork512(input_nc=24 * 2, output_nc=512, ngf=64, n_downsampling=6)

    def forward(self, z, c, retain, pose, truncation_psi=1, truncation_cutoff=None, **synthesis_kwargs):
        pose_feat = self.const_encoding(pose)
        stylecode, feats = self.style_encoding(c, retain)
        ws = self.mapping(z, stylecode, truncation_psi=truncation_psi, truncation_cutoff=truncation_cutoff)
        return self.synthesis(ws, pose_feat, {str(f.shape[2]): f for f in feats}, **synthesis_kwargs)


def build_generator_512(channel_base=16384, channel_max=512):
    return Generator512(z_dim=0, c_dim=512, w_dim=512, img_resolution=512, img_channels=3, mapping_kwargs=dict(num_layers=1),
                        synthesis_kwargs=dict(channel_base=channel_base, channel_max=channel_max, num_fp16_res=0, conv_clamp=256, use_noise=True))


# ----------------------------------------------------------------------------- discriminator (training path, R1 double backward)


class DiscriminatorBlock(OpsModule):
    """reference :917-997 — fromrgb (first block / skip arch), 3x3 conv, 3x3 down-2 conv, optional 1x1 down-2 residual skip."""

    def __init__(self, in_channels, tmp_channels, out_channels, resolution, img_channels, first_layer_idx, architecture='resnet',
                 activation='lrelu', resample_filter=_FIR, conv_clamp=None, use_fp16=False, fp16_channels_last=False, freeze_layers=0):
        assert in_channels in [0, tmp_channels]
        assert architecture in ['orig', 'skip', 'resnet']
        super().__init__()
        self.in_channels, self.resolution, self.img_channels = in_channels, resolution, img_channels
        self.first_layer_idx, self.architecture, self.use_fp16 = first_layer_idx, architecture, use_fp16
        self.channels_last = use_fp16 and fp16_channels_last
        self.register_buffer('resample_filter', _fir_buffer(resample_filter))
        self.num_layers = 0

        def trainable():
            t = (self.first_layer_idx + self.num_layers) >= freeze_layers
            self.num_layers += 1
            return t

        if in_channels == 0 or architecture == 'skip':
            self.fromrgb = Conv2dLayer(img_channels, tmp_channels, kernel_size=1, activation=activation, trainable=trainable(), conv_clamp=conv_clamp)
        self.conv0 = Conv2dLayer(tmp_channels, tmp_channels, kernel_size=3, activation=activation, trainable=trainable(), conv_clamp=conv_clamp)
        self.conv1 = Conv2dLayer(tmp_channels, out_channels, kernel_size=3, activation=activation, down=2, trainable=trainable(),
                                 resample_filter=resample_filter, conv_clamp=conv_clamp)
        if architecture == 'resnet':
            self.skip = Conv2dLayer(tmp_channels, out_channels, kernel_size=1, bias=False, down=2, trainable=trainable(), resample_filter=resample_filter)

    def forward(self, x, img, force_fp32=False):
        dtype = torch.float16 if self.use_fp16 and not force_fp32 else torch.float32
        mf = torch.channels_last if self.channels_last and not force_fp32 else torch.contiguous_format
        if x is not None:
            misc.assert_shape(x, [None, self.in_channels, self.resolution, self.resolution])
            x = x.to(dtype=dtype, memory_format=mf)
        if self.in_channels == 0 or self.architecture == 'skip':
            misc.assert_shape(img, [None, self.img_channels, self.resolution, self.resolution])
            img = img.to(dtype=dtype, memory_format=mf)
            y = self.fromrgb(img)
            x = x + y if x is not None else y
            img = self.ops.downsample2d(img, self.resample_filter) if self.architecture == 'skip' else None
        if self.architecture == 'resnet':
            y = self.skip(x, gain=np.sqrt(0.5))
            x = self.conv1(self.conv0(x), gain=np.sqrt(0.5), residual=y)
        else:
            x = self.conv1(self.conv0(x))
        assert x.dtype == dtype
        return x, img


class MinibatchStdLayer(nn.Module):
    """Appends the per-group feature standard deviation as an extra channel (reference :1001-1023)."""

    def __init__(self, group_size, num_channels=1):
        super().__init__()
        self.group_size, self.num_channels = group_size, num_channels

    def forward(self, x):
        n, c, h, w = x.shape
        g = min(int(self.group_size), int(n)) if self.group_size is not None else int(n)
        f = self.num_channels
        y = x.reshape(g, -1, f, c // f, h, w)
        y = y - y.mean(dim=0)
        y = (y.square().mean(dim=0) + 1e-8).sqrt()
        y = y.mean(dim=[2, 3, 4]).reshape(-1, f, 1, 1).repeat(g, 1, h, w)
        return torch.cat([x, y], dim=1)


class DiscriminatorEpilogue(OpsModule):
    def __init__(self, in_channels, cmap_dim, resolution, img_channels, architecture='resnet', mbstd_group_size=4, mbstd_num_channels=1,
                 activation='lrelu', conv_clamp=None):
        assert architecture in ['orig', 'skip', 'resnet']
        super().__init__()
        self.in_channels, self.cmap_dim, self.resolution, self.img_channels, self.architecture = in_channels, cmap_dim, resolution, img_channels, architecture
        if architecture == 'skip':
            self.fromrgb = Conv2dLayer(img_channels, in_channels, kernel_size=1, activation=activation)
        self.mbstd = MinibatchStdLayer(group_size=mbstd_group_size, num_channels=mbstd_num_channels) if mbstd_num_channels > 0 else None
        self.conv = Conv2dLayer(in_channels + mbstd_num_channels, in_channels, kernel_size=3, activation=activation, conv_clamp=conv_clamp)
        self.fc = FullyConnectedLayer(in_channels * (resolution ** 2), in_channels, activation=activation)
        self.out = FullyConnectedLayer(in_channels, 1 if cmap_dim == 0 else cmap_dim)

    def forward(self, x, img, cmap, force_fp32=False):
        misc.assert_shape(x, [None, self.in_channels, self.resolution, self.resolution])
        x = x.to(dtype=torch.float32, memory_format=torch.contiguous_format)
        if self.architecture == 'skip':
            x = x + self.fromrgb(img.to(torch.float32))
        if self.mbstd is not None:
            x = self.mbstd(x)
        x = self.out(self.fc(self.conv(x).flatten(1)))
        if self.cmap_dim > 0:
            misc.assert_shape(cmap, [None, self.cmap_dim])
            x = (x * cmap).sum(dim=1, keepdim=True) * (1 / np.sqrt(self.cmap_dim))
        return x


class Discriminator(OpsModule):
    """reference :1085-1143.  Residual architecture, fp16 storage for the `num_fp16_res` highest resolutions, projection on c."""

    def __init__(self, c_dim, img_resolution, img_channels, architecture='resnet', channel_base=32768, channel_max=512, num_fp16_res=0,
                 conv_clamp=None, cmap_dim=None, block_kwargs={}, mapping_kwargs={}, epilogue_kwargs={}):
        super().__init__()
        self.c_dim, self.img_resolution, self.img_channels = c_dim, img_resolution, img_channels
        self.img_resolution_log2 = int(np.log2(img_resolution))
        self.block_resolutions = [2 ** i for i in range(self.img_resolution_log2, 2, -1)]
        ch = {res: min(channel_base // res, channel_max) for res in self.block_resolutions + [4]}
        fp16_resolution = max(2 ** (self.img_resolution_log2 + 1 - num_fp16_res), 8)
        if cmap_dim is None:
            cmap_dim = ch[4]
        if c_dim == 0:
            cmap_dim = 0
        common = dict(img_channels=img_channels, architecture=architecture, conv_clamp=conv_clamp)
        idx = 0
        for res in self.block_resolutions:
            block = DiscriminatorBlock(ch[res] if res < img_resolution else 0, ch[res], ch[res // 2], resolution=res, first_layer_idx=idx,
                                       use_fp16=(res >= fp16_resolution), **block_kwargs, **common)
            setattr(self, f'b{res}', block)
            idx += block.num_layers
        if c_dim > 0:
            self.mapping = MappingNetwork(z_dim=0, c_dim=c_dim, w_dim=cmap_dim, num_ws=None, w_avg_beta=None, **mapping_kwargs)
        self.b4 = DiscriminatorEpilogue(ch[4], cmap_dim=cmap_dim, resolution=4, **epilogue_kwargs, **common)

    def forward(self, img, c, **block_kwargs):
        x = None
        for res in self.block_resolutions:
            x, img = getattr(self, f'b{res}')(x, img, **block_kwargs)
        cmap = self.mapping(None, c) if self.c_dim > 0 else None
        return self.b4(x, img, cmap)


def build_discriminator(img_resolution=256, channel_base=16384, channel_max=512, num_fp16_res=3):
    """Discriminator at the BASELINE training config (train_wo_flow_fullbody.py:191-197): c_dim 512, conv_clamp 256, mbstd group 4."""
    return Discriminator(c_dim=512, img_resolution=img_resolution, img_channels=3, channel_base=channel_base, channel_max=channel_max,
                         num_fp16_res=num_fp16_res, conv_clamp=256, epilogue_kwargs=dict(mbstd_group_size=4))


def build_generator_full(img_resolution=256, channel_base=16384, channel_max=512):
    """GeneratorFull at the BASELINE config: z_dim 0, c_dim = w_dim = 512, 1 mapping layer, conv_clamp 256, noise on."""
    return GeneratorFull(z_dim=0, c_dim=512, w_dim=512, img_resolution=img_resolution, img_channels=3,
                         mapping_kwargs=dict(num_layers=1),
                         synthesis_kwargs=dict(channel_base=channel_base, channel_max=channel_max, num_fp16_res=3, conv_clamp=256, use_noise=True))
