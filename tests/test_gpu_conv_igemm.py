"""GPU parity of the tcgen05 implicit-GEMM convolution against the CPU oracle (fp64 reference on the same inputs).

Tolerance: operands are rounded to fp16 (10-bit mantissa, the TF32 mantissa) or bf16 (7-bit) before the tensor-core
product, accumulation is fp32: 2e-3 relative (max-abs / max-abs) for fp16 operands, 1.5e-2 for bf16, per convolution.
The north_star bound for this kernel is 1e-2 relative on generator outputs (checked in test_gpu_network.py)."""
import pytest
import torch

from conftest import rel_err
from oracle import ops_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda'
TOL = {'fp16': 2e-3, 'bf16': 1.5e-2, 'tf32': 2e-3}


@pytest.fixture(scope='module')
def cv():
    from pasta_gan_b200.torch_utils.ops import conv_igemm
    return conv_igemm


import pasta_gan_b200  # noqa: E402

capi = pasta_gan_b200.capi


PLAIN = [
    # (N, Cin, Cout, H, W, k)
    (1, 16, 16, 8, 8, 3), (2, 32, 16, 16, 16, 3), (1, 64, 64, 32, 32, 3), (2, 128, 128, 32, 32, 3),
    (1, 64, 64, 64, 64, 1), (2, 48, 24, 20, 28, 3), (1, 3, 64, 33, 31, 3), (2, 42, 64, 16, 16, 1),
    (1, 512, 512, 4, 4, 3), (1, 512, 512, 16, 16, 3), (1, 256, 128, 64, 64, 3), (2, 64, 3, 32, 32, 1),
    (1, 192, 128, 40, 40, 1), (1, 64, 64, 128, 128, 3), (3, 20, 300, 12, 12, 3),
    # W = 128 with N tiles of >= 128 columns: column bands at a staged / useful ratio of 2
    (1, 128, 128, 24, 128, 3), (1, 64, 256, 16, 128, 3), (2, 32, 144, 130, 128, 3),
    # small Cin: taps folded into the GEMM K dimension (7x7 / 5x5 / 3x3)
    (2, 3, 64, 40, 36, 7), (1, 3, 64, 256, 256, 7), (2, 6, 32, 17, 19, 5), (1, 2, 16, 8, 8, 7), (2, 17, 48, 24, 24, 3),
]


@pytest.mark.parametrize('shape', PLAIN, ids=[str(s) for s in PLAIN])
@pytest.mark.parametrize('fmt', ['fp16', 'bf16', 'tf32'])
def test_plain_conv(cv, shape, fmt):
    n, cin, cout, h, w, k = shape
    torch.manual_seed(sum(shape))
    x = torch.randn(n, cin, h, w)
    wt = torch.randn(cout, cin, k, k) / (cin * k * k) ** 0.5
    for flip_weight in (True, False):
        ref = O._conv(x.double(), wt.double(), padding=k // 2, flip_weight=flip_weight)
        y = cv.conv2d_igemm(x.to(DEV), wt.to(DEV), flip_weight=flip_weight, fmt=fmt)
        assert y.shape == ref.shape
        assert rel_err(y, ref) < TOL[fmt], (shape, flip_weight)


ROWFOLD = [
    # (N, Cin, Cout, H, W, k): W % 128 == 0 and Cin * k <= 32 -> the row-folded first-layer kernel
    (2, 3, 64, 40, 128, 7), (1, 3, 64, 256, 256, 7), (3, 3, 64, 17, 256, 3), (1, 4, 32, 9, 128, 5), (2, 6, 128, 12, 384, 3), (1, 1, 16, 5, 128, 7), (1, 3, 48, 300, 128, 7), (2, 6, 64, 20, 256, 1), (1, 4, 32, 8, 128, 1), (3, 3, 64, 7, 384, 7),
]


@pytest.mark.parametrize('shape', ROWFOLD, ids=[str(s) for s in ROWFOLD])
def test_rowfold_first_layer(cv, shape):
    """Small-Cin layers on wide images run conv_rowfold_kernel (kernel rows folded into K, horizontal taps as descriptor offsets): against the fp64
    oracle, against the folded-tap path it replaces (same fp16 operand roundings, different summation order), and with a channel-blocked output."""
    n, cin, cout, h, w, k = shape
    torch.manual_seed(sum(shape))
    x = torch.randn(n, cin, h, w)
    wt = torch.randn(cout, cin, k, k) / (cin * k * k) ** 0.5
    b = torch.randn(cout) * 0.2
    xd, wd, bd = x.to(DEV), wt.to(DEV), b.to(DEV)
    for flip_weight, act, gain, clamp in ((True, 'relu', 2 ** 0.5, None), (False, 'lrelu', 1.3, 1.5), (True, 'linear', 1.0, None)):
        ref = O.bias_act(O._conv(x.double(), wt.double(), padding=k // 2, flip_weight=flip_weight), b.double(), act=act, gain=gain, clamp=clamp)
        n0 = capi.launch_count()
        y = cv.conv2d_igemm(xd, wd, flip_weight=flip_weight, bias=bd, act=act, gain=gain, clamp=clamp)
        assert rel_err(y, ref) < TOL['fp16'], (shape, act)
        capi.set_tuning('conv_rowfold', 0)
        try:
            y_old = cv.conv2d_igemm(xd, wd, flip_weight=flip_weight, bias=bd, act=act, gain=gain, clamp=clamp)
        finally:
            capi.set_tuning('conv_rowfold', 1)
        assert rel_err(y, y_old) < 5e-5                                    # same operand roundings, different summation order
        if cout % 16 == 0:
            yc = cv.conv2d_igemm(xd, wd, flip_weight=flip_weight, bias=bd, act=act, gain=gain, clamp=clamp, out_c8=True)
            assert torch.equal(cv.from_c8(yc, cout, dtype=torch.float16), y.half())
        assert capi.launch_count() > n0


def test_tf32_operand_kind(cv):
    """kind::tf32 (north_star: "TF32/BF16 tensor cores, fp32 accumulate"): fp32 exponent range, 10-bit mantissa.  Modulated layer with noise / bias /
    lrelu / clamp, up-2 and down-2 forms, the fused concat, a SPADE-style input activation -- all against the fp64 oracle -- and activations far
    outside the fp16 range (1e6), which the fp16 operand format cannot represent."""
    torch.manual_seed(7)
    n, cin, cout, h, w = 2, 48, 64, 20, 28
    x = torch.randn(n, cin, h, w)
    wt = torch.randn(cout, cin, 3, 3) / (cin * 9) ** 0.5
    st = 1 + 0.3 * torch.randn(n, cin)
    dc = (torch.einsum('oikl,ni->no', wt.square(), st.square()) + 1e-8).rsqrt()
    b = torch.randn(cout) * 0.2
    nz = torch.randn(h, w) * 0.1
    ref = O.bias_act(O._conv(x.double() * st.double()[:, :, None, None], wt.double(), padding=1) * dc.double()[:, :, None, None] + nz.double(), b.double(),
                     act='lrelu', gain=2 ** 0.5, clamp=3.0)
    y = cv.conv2d_igemm(x.to(DEV), wt.to(DEV), styles=st.to(DEV), dcoefs=dc.to(DEV), noise=nz.to(DEV), bias=b.to(DEV), act='lrelu', gain=2 ** 0.5, clamp=3.0, fmt='tf32')
    assert rel_err(y, ref) < TOL['tf32']
    f = O.setup_filter([1, 3, 3, 1])
    for up, down in ((2, 1), (1, 2)):
        ref = O.conv2d_resample(x.double(), wt.double(), f.double(), up=up, down=down, padding=1, flip_weight=(up == 1))
        y = cv.conv2d_igemm(x.to(DEV), wt.to(DEV), f=f.to(DEV), up=up, down=down, flip_weight=(up == 1), fmt='tf32')
        assert rel_err(y, ref) < TOL['tf32'], (up, down)
    x2 = torch.randn(n, 16, h, w)
    w2 = torch.randn(cout, cin + 16, 1, 1) / (cin + 16) ** 0.5
    ref = O._conv(torch.relu(torch.cat([x, x2], 1).double()) * 1.3, w2.double(), padding=0)
    y = cv.conv2d_igemm(x.to(DEV), w2.to(DEV), x2=x2.to(DEV), in_act='relu', in_gain=1.3, fmt='tf32')
    assert rel_err(y, ref) < TOL['tf32']
    big = x * 1e6                                                         # fp16 saturates at 65504: tf32 keeps the fp32 exponent
    ref = O._conv(big.double(), wt.double(), padding=1)
    assert rel_err(cv.conv2d_igemm(big.to(DEV), wt.to(DEV), fmt='tf32'), ref) < TOL['tf32']
    assert rel_err(cv.conv2d_igemm(big.to(DEV), wt.to(DEV), fmt='fp16'), ref) > 0.5


def test_fused_epilogue_and_modulation(cv):
    """SynthesisLayer in one launch: styles, demodulation, noise, bias, lrelu, gain, clamp."""
    torch.manual_seed(1)
    n, cin, cout, h = 3, 64, 48, 24
    x = torch.randn(n, cin, h, h)
    wt = torch.randn(cout, cin, 3, 3)
    s = 1 + 0.5 * torch.randn(n, cin)
    noise = torch.randn(h, h) * 0.3
    b = torch.randn(cout) * 0.2
    ref = O.bias_act(O.modulated_conv2d(x.double(), wt.double(), s.double(), noise=noise.double(), padding=1), b.double(), act='lrelu', gain=1.2, clamp=1.5)
    d = (s.square() @ wt.square().sum(dim=[2, 3]).t() + 1e-8).rsqrt()
    y = cv.conv2d_igemm(x.to(DEV), wt.to(DEV), styles=s.to(DEV), dcoefs=d.to(DEV), noise=noise.to(DEV), bias=b.to(DEV),
                        act='lrelu', gain=1.2, clamp=1.5)
    assert rel_err(y, ref) < 3e-3
    # per-sample (random-mode) noise and a SPADE-style pre-activation
    nz = torch.randn(n, 1, h, h) * 0.3
    ref2 = O._conv(O.bias_act(x.double(), act='relu', gain=2 ** 0.5), wt.double() / 24, padding=1) + nz.double()
    y2 = cv.conv2d_igemm(x.to(DEV), (wt / 24).to(DEV), noise=nz.to(DEV), in_act='relu', in_gain=2 ** 0.5)
    assert rel_err(y2, ref2) < 3e-3


UP2 = [(1, 16, 16, 8, 8), (2, 64, 32, 16, 16), (1, 128, 64, 32, 32), (2, 32, 128, 12, 20), (1, 512, 512, 4, 4), (1, 128, 64, 128, 128)]


@pytest.mark.parametrize('shape', UP2, ids=[str(s) for s in UP2])
def test_up2_polyphase(cv, shape):
    """conv2d_resample(up=2, k=3, padding=1) == zero-insert + FIR(gain 4) + conv, evaluated on the low-res input."""
    n, cin, cout, h, w = shape
    torch.manual_seed(sum(shape))
    f = O.setup_filter([1, 3, 3, 1])
    x = torch.randn(n, cin, h, w)
    wt = torch.randn(cout, cin, 3, 3) / (cin * 9) ** 0.5
    for flip_weight in (False, True):
        ref = O.conv2d_resample(x.double(), wt.double(), f=f.double(), up=2, padding=1, flip_weight=flip_weight)
        y = cv.conv2d_igemm(x.to(DEV), wt.to(DEV), f=f.to(DEV), up=2, flip_weight=flip_weight)
        assert y.shape == ref.shape
        assert rel_err(y, ref) < 3e-3, (shape, flip_weight)


def test_linearity_at_baseline_size(cv):
    """[16,128,128,128] -> 128 channels (the SPADE-block shape that carries 71 % of the generator's FLOPs): linear in x, and
    one sample's first output channels match the oracle."""
    torch.manual_seed(3)
    x1 = torch.randn(16, 128, 128, 128, device=DEV)
    x2 = torch.randn(16, 128, 128, 128, device=DEV)
    wt = torch.randn(128, 128, 3, 3, device=DEV) / (128 * 9) ** 0.5
    y1, y2 = cv.conv2d_igemm(x1, wt), cv.conv2d_igemm(x2, wt)
    y12 = cv.conv2d_igemm(x1 + x2, wt)
    assert rel_err(y12, y1 + y2) < 3e-3
    ref = O._conv(x1[5:6].cpu().double(), wt[:8].cpu().double(), padding=1)
    assert rel_err(y1[5:6, :8], ref) < 2e-3


@pytest.mark.parametrize('shape', [(2, 64, 3, 32, 32, True), (1, 512, 3, 4, 4, False), (2, 128, 6, 16, 16, False), (1, 64, 3, 256, 256, True), (3, 20, 3, 8, 12, True)],
                         ids=str)
def test_torgb_skip_kernel(shape):
    """img_out = upsample2d(img_in) + clamp(modulated 1x1 conv + b): fp32 SIMT kernel, 1e-5 relative vs the oracle composition."""
    from pasta_gan_b200.torch_utils.ops import torgb
    n, c, o, h, w, with_img = shape
    torch.manual_seed(sum(shape[:5]))
    f = O.setup_filter([1, 3, 3, 1])
    x = torch.randn(n, c, h, w)
    wt = torch.randn(o, c, 1, 1)
    s = (1 + 0.5 * torch.randn(n, c)) / c ** 0.5
    b = torch.randn(o) * 0.1
    img = torch.randn(n, o, h // 2, w // 2) if with_img else None
    y = O.bias_act(O.modulated_conv2d(x.double(), wt.double(), s.double(), demodulate=False), b.double(), clamp=1.0)
    ref = y + O.upsample2d(img.double(), f.double()) if with_img else y
    out = torgb.torgb_skip(x.to(DEV), wt.to(DEV), styles=s.to(DEV), bias=b.to(DEV), clamp=1.0, img=(img.to(DEV) if with_img else None), f=f.to(DEV))
    assert rel_err(out, ref) < 1e-5


@pytest.mark.parametrize('shape', [(2, 32, 48, 24, 24, 3), (1, 128, 128, 64, 64, 3), (2, 64, 16, 20, 28, 1)], ids=str)
def test_spade_fused_epilogue(cv, shape):
    """act(instance_norm(x) * (1 + conv(feat, Wg)) + conv(feat, Wb)) * gain with gamma / beta kept in TMEM (reference Spade_Norm_Block)."""
    n, c, cin, h, w, k = shape
    torch.manual_seed(sum(shape))
    x = torch.randn(n, c, h, w) * 2 + 0.5
    feat = torch.randn(n, cin, h, w)
    wg = torch.randn(c, cin, k, k) / (cin * k * k) ** 0.5
    wb = torch.randn(c, cin, k, k) / (cin * k * k) ** 0.5
    xd = x.double()
    norm = (xd - xd.mean(dim=(2, 3), keepdim=True)) / (xd.var(dim=(2, 3), unbiased=False, keepdim=True) + 1e-5).sqrt()
    ref = norm * (1 + O._conv(feat.double(), wg.double(), padding=k // 2)) + O._conv(feat.double(), wb.double(), padding=k // 2)
    ref = O.bias_act(ref, None, act='relu', gain=1.3)
    assert cv.spade_supported(x.to(DEV), feat.to(DEV), wg.to(DEV), wb.to(DEV))
    y = cv.spade_conv_norm(x.to(DEV), feat.to(DEV), wg.to(DEV), wb.to(DEV), act='relu', gain=1.3)
    assert rel_err(y, ref) < 3e-3


DOWN2 = [(1, 16, 16, 8, 8), (2, 64, 128, 32, 32), (1, 128, 256, 64, 64), (2, 32, 48, 12, 20), (1, 64, 64, 256, 256)]


@pytest.mark.parametrize('shape', DOWN2, ids=[str(s) for s in DOWN2])
def test_down2_space_to_depth(cv, shape):
    """conv2d_resample(down=2, k=3, padding=1) == FIR(pad 2) + 3x3 stride-2 conv, evaluated as one 'same' conv over the s2d planes of x."""
    n, cin, cout, h, w = shape
    torch.manual_seed(sum(shape))
    f = O.setup_filter([1, 3, 3, 1])
    x = torch.randn(n, cin, h, w)
    wt = torch.randn(cout, cin, 3, 3) / (cin * 9) ** 0.5
    b = torch.randn(cout) * 0.1
    for flip_weight in (True, False):
        ref = O.conv2d_resample(x.double(), wt.double(), f=f.double(), down=2, padding=1, flip_weight=flip_weight)
        y = cv.conv2d_igemm(x.to(DEV), wt.to(DEV), f=f.to(DEV), down=2, flip_weight=flip_weight)
        assert y.shape == ref.shape
        assert rel_err(y, ref) < 3e-3, (shape, flip_weight)
    y2 = cv.conv2d_igemm(x.to(DEV), wt.to(DEV), f=f.to(DEV), down=2, bias=b.to(DEV), act='lrelu', gain=2 ** 0.5, clamp=256)
    ref2 = O.bias_act(O.conv2d_resample(x.double(), wt.double(), f=f.double(), down=2, padding=1), b.double(), act='lrelu', clamp=256)
    assert rel_err(y2, ref2) < 3e-3


SPLIT = [(2, 64, 64, 32, 20, 28, 1), (1, 128, 64, 128, 33, 31, 1), (2, 16, 24, 48, 16, 16, 3), (1, 512, 64, 512, 32, 32, 1), (2, 8, 42, 64, 10, 12, 3)]


@pytest.mark.parametrize('shape', SPLIT, ids=[str(s) for s in SPLIT])
def test_split_input_and_residual(cv, shape):
    """conv([x ; x2]) without the concatenation (merge_conv, reference :5705-5706) and the residual add in the epilogue (`y.add_(x)`, :990)."""
    n, c1, c2, cout, h, w, k = shape
    torch.manual_seed(sum(shape))
    x, x2 = torch.randn(n, c1, h, w), torch.randn(n, c2, h, w)
    wt = torch.randn(cout, c1 + c2, k, k) / ((c1 + c2) * k * k) ** 0.5
    b = torch.randn(cout) * 0.2
    res = torch.randn(n, cout, h, w)
    ref = O.bias_act(O._conv(torch.cat([x, x2], 1).double(), wt.double(), padding=k // 2), b.double(), act='lrelu', gain=0.7, clamp=0.9) + res.double()
    y = cv.conv2d_igemm(x.to(DEV), wt.to(DEV), x2=x2.to(DEV), bias=b.to(DEV), act='lrelu', gain=0.7, clamp=0.9, residual=res.to(DEV))
    assert rel_err(y, ref) < 2e-3
    # residual with the polyphase up-2 output mapping
    if k == 3 and cout % 16 == 0:
        f = O.setup_filter([1, 3, 3, 1])
        res2 = torch.randn(n, cout, 2 * h, 2 * w)
        xx = torch.cat([x, x2], 1)
        ref2 = O.conv2d_resample(xx.double(), wt.double(), f.double(), up=2, padding=1, flip_weight=False) + res2.double()
        y2 = cv.conv2d_igemm(xx.to(DEV), wt.to(DEV), f=f.to(DEV), up=2, flip_weight=False, residual=res2.to(DEV))
        assert rel_err(y2, ref2) < 2e-3


def test_residual_blocks_match_unfused(cv):
    """ResBlock / SpadeResBlockV2 with the fused residual equal the same modules with the tcgen05 path disabled (library convs + add_)."""
    from pasta_gan_b200 import networks as N
    torch.manual_seed(5)
    blk = N.ResBlock(32, 64, kernel_size=3, activation='relu', down=2).to(DEV).eval().requires_grad_(False)
    x = torch.randn(2, 32, 32, 32, device=DEV)
    with torch.no_grad():
        y = blk(x)
        cv.enabled = False
        try:
            ref = blk(x)
        finally:
            cv.enabled = True
    assert rel_err(y, ref) < 3e-3


@pytest.mark.parametrize('shape', [(2, 16, 128, 128), (3, 5, 33, 31), (1, 128, 16, 16), (2, 3, 4, 4), (1, 2, 256, 256)])
def test_instance_stats(cv, shape):
    """pg_instance_norm_stats == nn.InstanceNorm2d's statistics (biased variance, eps 1e-5), also with a large common offset."""
    torch.manual_seed(sum(shape))
    for offset in (0.0, 100.0):
        x = (torch.randn(*shape) * 1.7 + offset)
        mean, rstd = cv.instance_stats(x.to(DEV))
        m_ref, r_ref = O.instance_norm_stats(x.double())
        assert float((mean.double().cpu() - m_ref).abs().max()) < 1e-6 * max(1.7, offset)      # mean error relative to the data scale
        assert rel_err(rstd, r_ref) < (1e-5 if offset == 0 else 2e-4)


@pytest.mark.parametrize('shape', [(2, 16, 32, 32), (3, 5, 9, 7), (1, 128, 128, 128)])
def test_masked_mean_fill(shape):
    """pg_masked_plane_sum + pg_masked_fill == the tail of get_spade_feat (reference :5791-5800), written into a channel slice."""
    from pasta_gan_b200.torch_utils.ops import spade_feat as S
    n, c, h, w = shape
    torch.manual_seed(sum(shape))
    feat = torch.randn(n, c, h, w)
    m1 = (torch.rand(n, 1, h, w) > 0.5).float(); m2 = (torch.rand(n, 1, h, w) > 0.4).float()
    if n > 1:
        m2[0] = 0                                                   # a sample without enough valid pixels: count falls back to H*W
    valid = ((m1 + m2) == 2.0).float(); rest = m1 - valid
    ref = O.masked_mean_fill(feat.double(), valid.double(), rest.double())
    buf = torch.full([n, 2 * c + 3, h, w], float('nan'), device=DEV)
    out = S.masked_mean_fill(feat.to(DEV), valid.to(DEV), rest.to(DEV), buf[:, 2:2 + c])
    assert out.data_ptr() == buf[:, 2:2 + c].data_ptr()
    assert rel_err(buf[:, 2:2 + c], ref) < 1e-5
    assert torch.isnan(buf[:, :2]).all() and torch.isnan(buf[:, 2 + c:]).all()          # neighbours of the slice untouched


@pytest.mark.parametrize('shape', [(2, 128, 64, 32, 32, 3), (1, 256, 128, 128, 128, 3), (2, 64, 48, 20, 28, 1), (1, 24, 32, 16, 16, 3), (1, 128, 128, 64, 256, 3)])
def test_half_intermediates_are_bit_identical(cv, shape):
    """An fp16 NCHW input is taken as the operand bits and an fp16 output is the fp32 result rounded like the consumer's loader would round it:
    conv(x.half()) == conv(x.half().float()) exactly, and conv(..., out_dtype=half) == conv(...).half() (same kernel, same accumulation order)."""
    n, cin, cout, h, w, k = shape
    torch.manual_seed(sum(shape))
    x = (torch.randn(n, cin, h, w) * 3).to(DEV)
    wt = (torch.randn(cout, cin, k, k) / (cin * k * k) ** 0.5).to(DEV)
    b = torch.randn(cout, device=DEV)
    xh = x.half()
    y32 = cv.conv2d_igemm(xh.float(), wt, bias=b, act='relu', gain=1.3)
    y_in16 = cv.conv2d_igemm(xh, wt, bias=b, act='relu', gain=1.3)
    assert torch.equal(y_in16, y32)
    y_out16 = cv.conv2d_igemm(xh, wt, bias=b, act='relu', gain=1.3, out_dtype=torch.float16)
    assert y_out16.dtype == torch.float16 and torch.equal(y_out16, y32.half())
    # SPADE epilogue with fp16 feature input and fp16 output
    if k == 3 and cout % 16 == 0 and 2 * cout <= 256 and cin * 9 > 160:
        xs = torch.randn(n, cout, h, w, device=DEV)
        wg = torch.randn(cout, cin, 3, 3, device=DEV) / (cin * 9) ** 0.5; wb = torch.randn(cout, cin, 3, 3, device=DEV) / (cin * 9) ** 0.5
        r32 = cv.spade_conv_norm(xs, xh.float(), wg, wb, act='relu', gain=1.1)
        r16 = cv.spade_conv_norm(xs, xh, wg, wb, act='relu', gain=1.1, out_dtype=torch.float16)
        assert torch.equal(r16, r32.half())


def test_persistent_variant_matches(cv, monkeypatch):
    """The opt-in persistent kernel (one CTA per SM, double-buffered TMEM, dedicated epilogue warps; pg_set_tuning('conv_persist', 1)) computes the same
    tiles with the same accumulation order as the default one-tile kernel: identical bits."""
    torch.manual_seed(11)
    for (n, cin, cout, h, w, k, mod) in [(16, 128, 128, 64, 64, 3, False), (16, 64, 64, 128, 128, 3, True), (16, 96, 64, 64, 96, 1, False)]:
        x = torch.randn(n, cin, h, w, device=DEV)
        wt = torch.randn(cout, cin, k, k, device=DEV) / (cin * k * k) ** 0.5
        b = torch.randn(cout, device=DEV)
        st = (1 + 0.3 * torch.randn(n, cin, device=DEV)) if mod else None
        dc = (torch.rand(n, cout, device=DEV) + 0.5) if mod else None
        res = torch.randn(n, cout, h, w, device=DEV)
        capi.set_tuning('conv_persist', 0)
        y0 = cv.conv2d_igemm(x, wt, styles=st, dcoefs=dc, bias=b, act='lrelu', gain=1.2, clamp=3.0, residual=res)
        capi.set_tuning('conv_persist', 1)
        y1 = cv.conv2d_igemm(x, wt, styles=st, dcoefs=dc, bias=b, act='lrelu', gain=1.2, clamp=3.0, residual=res)
        capi.set_tuning('conv_persist', 0)
        assert torch.equal(y0, y1), (n, cin, cout, h, w, k, mod)


BANDS = [(1, 64, 64, 256, 256), (1, 32, 48, 40, 320), (2, 16, 16, 24, 258), (1, 24, 32, 130, 512), (2, 64, 32, 9, 256)]


@pytest.mark.parametrize('shape', BANDS, ids=[str(s) for s in BANDS])
def test_column_band_mode(cv, shape, monkeypatch):
    """W >= 256: the image is processed in 64-column bands with real halo columns (strip pitch 68) instead of one full-width strip.  Same GEMM,
    different tiling: parity against the oracle with styles / demodulation / noise / bias / lrelu / clamp / residual, and bit-equality with
    the full-width tiling (pg_set_tuning('conv_bands', 0)) -- every output is the same sum of the same fp16 products in the same chunk order."""
    n, cin, cout, h, w = shape
    torch.manual_seed(sum(shape))
    x = torch.randn(n, cin, h, w)
    wt = torch.randn(cout, cin, 3, 3)
    s = 1 + 0.5 * torch.randn(n, cin)
    noise = torch.randn(h, w) * 0.3
    b = torch.randn(cout) * 0.2
    res = torch.randn(n, cout, h, w)
    ref = O.bias_act(O.modulated_conv2d(x.double(), wt.double(), s.double(), noise=noise.double(), padding=1), b.double(), act='lrelu', gain=1.2, clamp=1.5) + res.double()
    d = (s.square() @ wt.square().sum(dim=[2, 3]).t() + 1e-8).rsqrt()
    args = dict(styles=s.to(DEV), dcoefs=d.to(DEV), noise=noise.to(DEV), bias=b.to(DEV), act='lrelu', gain=1.2, clamp=1.5, residual=res.to(DEV))
    y = cv.conv2d_igemm(x.to(DEV), wt.to(DEV), **args)
    assert rel_err(y, ref) < 3e-3
    capi.set_tuning('conv_bands', 0)
    y0 = cv.conv2d_igemm(x.to(DEV), wt.to(DEV), **args)
    capi.set_tuning('conv_bands', 1)
    assert torch.equal(y, y0)
    # plain layer, fp16 input and output
    wp = (wt / (cin * 9) ** 0.5).to(DEV)
    xh = x.to(DEV).half()
    if w <= 256 and cin * 9 > 160:
        assert torch.equal(cv.conv2d_igemm(xh, wp, out_dtype=torch.float16), cv.conv2d_igemm(xh.float(), wp).half())


@pytest.mark.parametrize('shape', [(1, 32, 32, 16, 512, 'down'), (2, 16, 48, 40, 520, 'down'), (1, 32, 16, 24, 256, 'up'), (2, 16, 32, 10, 322, 'up')])
def test_column_band_mode_resampling(cv, shape, monkeypatch):
    """Band tiling under the fused resampling forms: down-2 (bands over the OUTPUT columns of the space-to-depth GEMM) and polyphase up-2 (bands over
    the INPUT columns).  Parity against conv2d_resample's definition and bit-equality with the full-width tiling."""
    n, cin, cout, h, w, mode = shape
    torch.manual_seed(sum(shape[:5]))
    f = O.setup_filter([1, 3, 3, 1])
    x = torch.randn(n, cin, h, w)
    wt = torch.randn(cout, cin, 3, 3) / (cin * 9) ** 0.5
    b = torch.randn(cout) * 0.2
    if mode == 'down':
        ref = O.bias_act(O.conv2d_resample(x.double(), wt.double(), f.double(), down=2, padding=1), b.double(), act='lrelu')
        run = lambda: cv.conv2d_igemm(x.to(DEV), wt.to(DEV), f=f.to(DEV), down=2, bias=b.to(DEV), act='lrelu', gain=2 ** 0.5)
    else:
        ref = O.bias_act(O.conv2d_resample(x.double(), wt.double(), f.double(), up=2, padding=1, flip_weight=False), b.double(), act='lrelu')
        run = lambda: cv.conv2d_igemm(x.to(DEV), wt.to(DEV), f=f.to(DEV), up=2, flip_weight=False, bias=b.to(DEV), act='lrelu', gain=2 ** 0.5)
    y = run()
    assert y.shape == ref.shape and rel_err(y, ref) < 3e-3
    capi.set_tuning('conv_bands', 0)
    y0 = run()
    capi.set_tuning('conv_bands', 1)
    assert torch.equal(y, y0)


def test_band_mode_with_split_input_spade_and_modulated_down2(cv, monkeypatch):
    """Remaining combinations of the band tiling and the lean down-2 loader: 3x3 over a split input, the SPADE epilogue (bands vs full width,
    bit-equal), and a style-modulated down-2 layer against the oracle."""
    torch.manual_seed(3)
    # split input, 3x3, W = 128, 128 output channels -> banded
    x, x2 = torch.randn(1, 64, 20, 128), torch.randn(1, 64, 20, 128)
    wt = torch.randn(128, 128, 3, 3) / (128 * 9) ** 0.5
    ref = O._conv(torch.cat([x, x2], 1).double(), wt.double(), padding=1)
    y = cv.conv2d_igemm(x.to(DEV), wt.to(DEV), x2=x2.to(DEV))
    assert rel_err(y, ref) < 2e-3
    # SPADE epilogue: bands on / off
    xs = torch.randn(2, 128, 24, 128, device=DEV); feat = torch.randn(2, 64, 24, 128, device=DEV)
    wg = torch.randn(128, 64, 3, 3, device=DEV) / 24; wb = torch.randn(128, 64, 3, 3, device=DEV) / 24
    r1 = cv.spade_conv_norm(xs, feat, wg, wb, act='relu', gain=1.1)
    capi.set_tuning('conv_bands', 0)
    r0 = cv.spade_conv_norm(xs, feat, wg, wb, act='relu', gain=1.1)
    capi.set_tuning('conv_bands', 1)
    assert torch.equal(r0, r1)
    # modulated down-2 (styles in the lean down-2 loader's prologue)
    f = O.setup_filter([1, 3, 3, 1])
    xd = torch.randn(2, 32, 24, 40); wd = torch.randn(48, 32, 3, 3); st = 1 + 0.4 * torch.randn(2, 32)
    refd = O.modulated_conv2d(xd.double(), wd.double(), st.double(), down=2, padding=1, resample_filter=f.double())
    dco = (st.square() @ wd.square().sum(dim=[2, 3]).t() + 1e-8).rsqrt()
    yd = cv.conv2d_igemm(xd.to(DEV), wd.to(DEV), f=f.to(DEV), down=2, styles=st.to(DEV), dcoefs=dco.to(DEV))
    assert yd.shape == refd.shape and rel_err(yd, refd) < 3e-3


# ------------------------------------------------------------------------------------------------ groups = N (the reference's fused modulated conv)

def test_groups_n_matches_modulated_goldens(golden):
    """The fused cases of tests/golden/modulated_conv2d.npz through the reference's own arithmetic (networks.py:64-94): per-sample weights
    w * s * d built with torch, then conv2d_resample(x [1, N*I, H, W], w [N*O, I, k, k], groups = N) -- which must land on the tcgen05 kernel
    (no library convolution) for every case the kernel covers, including the up-2 form."""
    from pasta_gan_b200.torch_utils.ops import conv2d_resample as C
    g = golden('modulated_conv2d')
    f4 = g.t('f4', device=DEV)
    ran = 0
    for idx, c in enumerate(g.meta):
        if not c['fused']:
            continue
        x, w, s = g.t(f'{idx}/x', device=DEV), g.t(f'{idx}/w', device=DEV), g.t(f'{idx}/s', device=DEV)
        n, o, i, k = c['n'], c['o'], c['i'], c['k']
        wn = w.unsqueeze(0) * s.reshape(n, 1, -1, 1, 1)
        if c['demod']:
            wn = wn * (wn.square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt().reshape(n, -1, 1, 1, 1)
        l0 = capi.launch_count()
        with torch.no_grad():
            y = C.conv2d_resample(x=x.reshape(1, -1, *x.shape[2:]), w=wn.reshape(-1, i, k, k), f=f4, up=c['up'], padding=k // 2, groups=n,
                                  flip_weight=c['flip_weight'])
        y = y.reshape(n, -1, *y.shape[2:])
        if g.has(f'{idx}/noise'):
            y = y + g.t(f'{idx}/noise', device=DEV)
        assert rel_err(y, g.t(f'{idx}/y')) < 3e-3, c
        if c['up'] == 1 or o % 16 == 0:
            assert capi.launch_count() - l0 == 2, 'expected exactly: batched weight pack + one tcgen05 launch'
            ran += 1
    assert ran >= 2


@pytest.mark.parametrize('shape', [(16, 64, 64, 32, 32, 3, 1), (4, 128, 64, 16, 16, 3, 2), (3, 32, 3, 24, 24, 1, 1), (2, 512, 512, 8, 8, 3, 1), (5, 48, 32, 20, 36, 3, 2)])
def test_groups_n_vs_oracle(cv, shape):
    """groups = N at generator-like shapes against the oracle's grouped conv2d_resample (fp64), stride 1 and up-2."""
    from pasta_gan_b200.torch_utils.ops import conv2d_resample as C
    n, cin, cout, h, w, k, up = shape
    torch.manual_seed(sum(shape))
    f = O.setup_filter([1, 3, 3, 1])
    x = torch.randn(1, n * cin, h, w)
    wt = torch.randn(n * cout, cin, k, k) / (cin * k * k) ** 0.5
    ref = O.conv2d_resample(x.double(), wt.double(), f.double(), up=up, padding=k // 2, groups=n, flip_weight=(up == 1))
    with torch.no_grad():
        y = C.conv2d_resample(x.to(DEV), wt.to(DEV), f.to(DEV), up=up, padding=k // 2, groups=n, flip_weight=(up == 1))
    assert y.shape == ref.shape and rel_err(y, ref) < TOL['fp16']


# ------------------------------------------------------------------------------------------------ channel-blocked fp16 (TMA operand path)

def test_c8_layout_round_trip(cv):
    torch.manual_seed(5)
    for shape in [(2, 16, 8, 8), (1, 20, 5, 7), (3, 64, 33, 31)]:
        x = torch.randn(*shape, device=DEV)
        y = cv.to_c8(x)
        assert y.shape == (shape[0], (shape[1] + 7) // 8, shape[2], shape[3], 8)
        ref = torch.zeros(shape[0], y.shape[1] * 8, shape[2], shape[3], device=DEV, dtype=torch.float16)
        ref[:, :shape[1]] = x.half()
        assert torch.equal(y, ref.reshape(shape[0], -1, 8, shape[2], shape[3]).permute(0, 1, 3, 4, 2))
        assert torch.equal(cv.from_c8(y, shape[1], dtype=torch.float16), x.half())
        assert torch.equal(cv.to_c8(x.half()), y)


C8 = [
    # (N, Cin, Cout, H, W, k): full-width strips (W <= 127), 64-column bands (W >= 128), 1x1, ragged heights / band counts
    (2, 32, 32, 16, 16, 3), (1, 64, 48, 33, 31, 3), (2, 128, 128, 64, 64, 3), (1, 256, 128, 24, 128, 3), (2, 64, 64, 40, 256, 3),
    (1, 32, 16, 9, 130, 3), (2, 128, 128, 32, 128, 1), (1, 64, 32, 50, 50, 1), (1, 32, 16, 4, 4, 3), (3, 48, 272, 20, 20, 3), (1, 32, 32, 130, 126, 3),
]


@pytest.mark.parametrize('shape', C8, ids=[str(s) for s in C8])
def test_c8_tma_path_bit_identical(cv, shape):
    """A channel-blocked fp16 input is the same operand bits the fp16-NCHW loader stages, multiplied in the same chunk order: results must be
    bit-identical to the converter path; a channel-blocked output holds the same fp16 values as the fp16-NCHW output."""
    n, cin, cout, h, w, k = shape
    torch.manual_seed(sum(shape))
    xh = torch.randn(n, cin, h, w, device=DEV).half()
    wt = torch.randn(cout, cin, k, k, device=DEV) / (cin * k * k) ** 0.5
    b = torch.randn(cout, device=DEV) * 0.2
    ref = cv.conv2d_igemm(xh.float(), wt, bias=b, act='lrelu', gain=1.3, clamp=2.0)
    assert rel_err(ref, O.bias_act(O._conv(xh.double().cpu(), wt.double().cpu(), padding=k // 2), b.double().cpu(), act='lrelu', gain=1.3, clamp=2.0)) < TOL['fp16']
    xc = cv.to_c8(xh)
    y = cv.conv2d_igemm(xc, wt, bias=b, act='lrelu', gain=1.3, clamp=2.0)
    assert torch.equal(y, ref)
    res = torch.randn_like(ref)
    assert torch.equal(cv.conv2d_igemm(xc, wt, bias=b, residual=res), cv.conv2d_igemm(xh.float(), wt, bias=b, residual=res))
    if cout % 16 == 0:
        yc = cv.conv2d_igemm(xc, wt, bias=b, act='lrelu', gain=1.3, clamp=2.0, out_c8=True)
        assert torch.equal(cv.from_c8(yc, cout, dtype=torch.float16), ref.half())
        y2 = cv.conv2d_igemm(xh.float(), wt, bias=b, act='lrelu', gain=1.3, clamp=2.0, out_c8=True)      # converter path in, blocked out
        assert torch.equal(y2, yc)


@pytest.mark.parametrize('shape', [(2, 64, 64, 24, 24, 3), (1, 128, 128, 40, 128, 3), (2, 32, 48, 16, 256, 1)], ids=str)
def test_channel_blocked_residual_with_dense_output(cv, shape):
    """A skip branch that is only read back as a residual travels channel-blocked fp16 while the block's result stays fp32 NCHW: the same numbers
    as the dense fp32 residual holding the fp16-rounded values, for the TMA and the converter operand paths."""
    n, cin, cout, h, w, k = shape
    torch.manual_seed(sum(shape))
    xh = torch.randn(n, cin, h, w, device=DEV).half()
    wt = torch.randn(cout, cin, k, k, device=DEV) / (cin * k * k) ** 0.5
    b = torch.randn(cout, device=DEV) * 0.2
    res = torch.randn(n, cout, h, w, device=DEV).half()
    for x in (cv.to_c8(xh), xh.float()):
        ref = cv.conv2d_igemm(x, wt, bias=b, gain=0.7, residual=res.float())
        got = cv.conv2d_igemm(x, wt, bias=b, gain=0.7, residual=cv.to_c8(res))
        assert got.dtype == torch.float32 and torch.equal(got, ref)


@pytest.mark.parametrize('shape', [(2, 64, 128, 256, 256), (1, 128, 256, 128, 128), (3, 32, 48, 36, 20), (2, 256, 256, 32, 32), (1, 16, 16, 8, 512), (1, 64, 64, 6, 260)], ids=str)
def test_c8_tma_down2(cv, shape, monkeypatch):
    """Down-2 3x3 from a channel-blocked input: the space-to-depth planes come through a strided 4-D TMA box (PG_CONV_DOWN2_C8).  Against the fp64
    oracle of conv2d_resample(down=2) and against the converter path on the same fp16-rounded input (same products, different K order)."""
    n, cin, cout, h, w = shape
    monkeypatch.setattr(cv, '_DOWN2_TMA_MIN', 4)                            # (the network only routes >= 256-px inputs this way)
    torch.manual_seed(sum(shape))
    xh = torch.randn(n, cin, h, w, device=DEV).half()
    wt = torch.randn(cout, cin, 3, 3, device=DEV) / (cin * 9) ** 0.5
    b = torch.randn(cout, device=DEV) * 0.2
    f = O.setup_filter([1, 3, 3, 1]).to(DEV)
    assert cv.c8_input_ok(cin, h, w, 3, 1, 2)
    ref = O.bias_act(O.conv2d_resample(xh.double().cpu(), wt.double().cpu(), f.double().cpu(), down=2, padding=1), b.double().cpu(), act='lrelu', gain=1.2)
    y = cv.conv2d_igemm(cv.to_c8(xh), wt, f=f, down=2, bias=b, act='lrelu', gain=1.2)
    assert y.shape == ref.shape and rel_err(y, ref) < TOL['fp16']
    y_conv = cv.conv2d_igemm(xh.float(), wt, f=f, down=2, bias=b, act='lrelu', gain=1.2)
    assert rel_err(y, y_conv) < 5e-5
    if cout % 16 == 0:
        yc = cv.conv2d_igemm(cv.to_c8(xh), wt, f=f, down=2, bias=b, act='lrelu', gain=1.2, out_c8=True)
        assert torch.equal(cv.from_c8(yc, cout, dtype=torch.float16), y.half())


def test_c8_tma_down2_1x1_skip(cv, monkeypatch):
    """The 1x1 down-2 skip of a down-sampling res-block on the strided-TMA path: a centre-tap 3x3 (conv2d_resample(down=2) of a 1x1 kernel filters with
    pad 1, of a 3x3 kernel with pad 2 -- the centre tap sits one sample in).  Against the oracle's 1x1 form, and following an in-place weight update."""
    monkeypatch.setattr(cv, '_DOWN2_TMA_MIN', 4)
    torch.manual_seed(3)
    n, cin, cout, h, w = 2, 64, 128, 64, 96
    xh = torch.randn(n, cin, h, w, device=DEV).half()
    wt = torch.nn.Parameter(torch.randn(cout, cin, 1, 1, device=DEV) / cin ** 0.5)
    f = O.setup_filter([1, 3, 3, 1]).to(DEV)
    for _ in range(2):
        ref = O.conv2d_resample(xh.double().cpu(), wt.detach().double().cpu(), f.double().cpu(), down=2, padding=0)
        y = cv.conv2d_igemm(cv.to_c8(xh), wt, f=f, down=2, gain=0.7, cache_weights=True)
        assert y.shape == ref.shape and rel_err(y, ref * 0.7) < TOL['fp16']
        with torch.no_grad():
            wt.mul_(-1.5)                                                    # optimizer-style in-place update: the embedded copy and its pack must follow


def test_c8_tma_up2_spade_and_folded_styles(cv):
    """TMA operand path under the other epilogues: polyphase up-2, the SPADE epilogue (blocked in, blocked out), and a modulated layer whose styles
    are folded into per-sample packed weights (the activations cannot be scaled on the way in)."""
    torch.manual_seed(17)
    f = O.setup_filter([1, 3, 3, 1]).to(DEV)
    xh = torch.randn(2, 64, 24, 24, device=DEV).half()
    wt = torch.randn(32, 64, 3, 3, device=DEV) / 24
    up_ref = cv.conv2d_igemm(xh.float(), wt, f=f, up=2, flip_weight=False, act='lrelu', gain=2 ** 0.5)
    assert torch.equal(cv.conv2d_igemm(cv.to_c8(xh), wt, f=f, up=2, flip_weight=False, act='lrelu', gain=2 ** 0.5), up_ref)
    # SPADE
    xs = torch.randn(2, 64, 24, 128, device=DEV); feat = torch.randn(2, 32, 24, 128, device=DEV).half()
    wg = torch.randn(64, 32, 3, 3, device=DEV) / 17; wb = torch.randn(64, 32, 3, 3, device=DEV) / 17
    r = cv.spade_conv_norm(xs, feat, wg, wb, act='relu', gain=1.1, out_dtype=torch.float16)
    rc = cv.spade_conv_norm(xs, cv.to_c8(feat), wg, wb, act='relu', gain=1.1, out_c8=True)
    assert torch.equal(cv.from_c8(rc, 64, dtype=torch.float16), r)
    # SPADE at C = 128 (the generator's shape): two [gamma_t | beta_t] N tiles of 128 columns on the TMA path vs one 256-column tile on the converter path
    xs = torch.randn(2, 128, 20, 128, device=DEV); feat = torch.randn(2, 128, 20, 128, device=DEV).half()
    wg = torch.randn(128, 128, 3, 3, device=DEV) / 34; wb = torch.randn(128, 128, 3, 3, device=DEV) / 34
    r = cv.spade_conv_norm(xs, feat, wg, wb, act='relu', gain=1.1, out_dtype=torch.float16)
    rc = cv.spade_conv_norm(xs, cv.to_c8(feat), wg, wb, act='relu', gain=1.1, out_c8=True)
    assert torch.equal(cv.from_c8(rc, 128, dtype=torch.float16), r)
    assert torch.equal(cv.spade_conv_norm(xs, cv.to_c8(feat), wg, wb, act='relu', gain=1.1), cv.spade_conv_norm(xs, feat, wg, wb, act='relu', gain=1.1))
    # styles folded into the weights vs styles on the activations: same math, different fp16 roundings -> oracle tolerance
    x = torch.randn(3, 64, 40, 40); wm = torch.randn(48, 64, 3, 3); st = 1 + 0.5 * torch.randn(3, 64); nz = torch.randn(40, 40) * 0.3; b = torch.randn(48) * 0.1
    ref = O.bias_act(O.modulated_conv2d(x.double(), wm.double(), st.double(), noise=nz.double(), padding=1), b.double(), act='lrelu', clamp=256)
    dco = (st.square() @ wm.square().sum(dim=[2, 3]).t() + 1e-8).rsqrt()
    y = cv.conv2d_igemm(cv.to_c8(x.to(DEV)), wm.to(DEV), styles=st.to(DEV), dcoefs=dco.to(DEV), noise=nz.to(DEV), bias=b.to(DEV), act='lrelu',
                        gain=2 ** 0.5, clamp=256, fold_styles=True)
    assert rel_err(y, ref) < 3e-3


# ------------------------------------------------------------------------------------------------ dynamic range / weight store

@pytest.mark.parametrize('fmt', ['fp16', 'bf16'])
def test_trained_like_dynamic_range(cv, fmt):
    """Unit-variance inputs never reach the edges of the fp16 range.  Here: activations at the conv_clamp (+-256), styles in the tens to hundreds
    (x * s would overflow fp16's 65504 un-normalised), weights x 100, plus a block of tiny (1e-6) activations.  The per-sample style
    normalisation (reference networks.py:57-59, folded into the demodulation coefficient) keeps the operands finite; the result must match the
    fp64 oracle to the operand format's tolerance with no inf / nan."""
    torch.manual_seed(23)
    n, cin, cout, h = 4, 64, 48, 24
    x = (torch.randn(n, cin, h, h) * 200).clamp(-256, 256)
    x[:, :8] = torch.randn(n, 8, h, h) * 1e-6
    wt = torch.randn(cout, cin, 3, 3) * 100
    st = torch.randn(n, cin) * 150 + 40 * torch.sign(torch.randn(n, cin))
    dco = (st.double().square() @ wt.double().square().sum(dim=[2, 3]).t() + 1e-8).rsqrt().float()
    ref = O.bias_act(O.modulated_conv2d(x.double(), wt.double(), st.double(), padding=1), None, act='lrelu', clamp=256)
    y = cv.conv2d_igemm(x.to(DEV), wt.to(DEV), styles=st.to(DEV), dcoefs=dco.to(DEV), act='lrelu', gain=2 ** 0.5, clamp=256, fmt=fmt)
    assert torch.isfinite(y).all()
    assert rel_err(y, ref) < TOL[fmt]
    # without demodulation (ToRGB form): outputs in the 1e7 range, still finite and accurate
    ref2 = O.modulated_conv2d(x.double(), wt.double()[:3, :, :1, :1], st.double(), demodulate=False)
    y2 = cv.conv2d_igemm(x.to(DEV), wt[:3, :, :1, :1].contiguous().to(DEV), styles=st.to(DEV), fmt=fmt)
    assert torch.isfinite(y2).all() and rel_err(y2, ref2) < TOL[fmt]
    # plain layer, activations +-1e4 (no clamp in the encoders): inside fp16 range, must not saturate
    xe = torch.randn(2, 32, 20, 20) * 1e4
    we = torch.randn(32, 32, 3, 3) / 17
    assert rel_err(cv.conv2d_igemm(xe.to(DEV), we.to(DEV), fmt=fmt), O._conv(xe.double(), we.double(), padding=1)) < TOL[fmt]


def test_packed_weight_store_refreshes_in_place(cv):
    """The packed copy of a parameter lives in ONE buffer: an in-place parameter update re-packs into the same storage (no stale versions pile up,
    addresses baked into CUDA graphs stay valid), and entries disappear with their parameter."""
    torch.manual_seed(29)
    w = torch.nn.Parameter(torch.randn(32, 32, 3, 3, device=DEV) / 17)
    x = torch.randn(2, 32, 16, 16, device=DEV)
    n0 = len(cv._pack_store)
    with torch.no_grad():
        y0 = cv.conv2d_igemm(x, w, cache_weights=True)
        key = [k for k in cv._pack_store if k[0] == id(w)]
        assert len(key) == 1 and len(cv._pack_store) == n0 + 1
        ptr = cv._pack_store[key[0]].ws.data_ptr()
        l0 = capi.launch_count()
        assert torch.equal(cv.conv2d_igemm(x, w, cache_weights=True), y0) and capi.launch_count() - l0 == 1       # cached: no second pack
        w.mul_(2.0)
        y1 = cv.conv2d_igemm(x, w, cache_weights=True)
        assert len(cv._pack_store) == n0 + 1 and cv._pack_store[key[0]].ws.data_ptr() == ptr
        assert rel_err(y1, 2 * y0) < 1e-6
        w.mul_(0.5)
        assert cv.refresh_packed_weights() >= 1 and cv._pack_store[key[0]].ws.data_ptr() == ptr
        assert torch.equal(cv.conv2d_igemm(x, w, cache_weights=True), y0)
    del w
    import gc
    gc.collect()
    assert len(cv._pack_store) == n0


# ------------------------------------------------------------------------------------------------ training: tcgen05 dgrad / wgrad

WGRAD = [(2, 32, 32, 16, 16, 3), (3, 48, 64, 20, 28, 3), (1, 128, 128, 64, 64, 3), (2, 64, 200, 33, 31, 3), (4, 20, 24, 12, 12, 3), (2, 3, 64, 40, 40, 1),
         (2, 192, 128, 24, 24, 1), (1, 512, 512, 4, 4, 3), (2, 64, 64, 128, 128, 3), (1, 40, 16, 9, 256, 3)]


@pytest.mark.parametrize('shape', WGRAD, ids=[str(s) for s in WGRAD])
def test_wgrad_and_dgrad_vs_autograd(cv, shape):
    """pg_conv2d_wgrad and the dgrad form of the forward kernel against fp64 autograd of F.conv2d on the same inputs.  bf16 operands, fp32
    accumulation over up to N*H*W products: 1e-2 relative (max-abs / max-abs), the tolerance the north_star gives tensor-core convolutions."""
    n, cin, cout, h, w, k = shape
    torch.manual_seed(sum(shape))
    x = torch.randn(n, cin, h, w)
    wt = torch.randn(cout, cin, k, k) / (cin * k * k) ** 0.5
    dy = torch.randn(n, cout, h, w) * 1e-3                       # gradient-like magnitudes
    xr, wr = x.double().requires_grad_(True), wt.double().requires_grad_(True)
    yr = torch.nn.functional.conv2d(xr, wr, padding=k // 2)
    dx_ref, dw_ref = torch.autograd.grad(yr, [xr, wr], dy.double())
    dw = cv.conv2d_wgrad(x.to(DEV), dy.to(DEV), k)
    assert dw.shape == dw_ref.shape and rel_err(dw, dw_ref) < 1e-2
    dx = cv.conv2d_dgrad(dy.to(DEV), wt.to(DEV))
    assert dx.shape == dx_ref.shape and rel_err(dx, dx_ref) < 1e-2
    acc = torch.ones_like(dw)
    cv.conv2d_wgrad(x.to(DEV), dy.to(DEV), k, scale=2.0, out=acc)
    assert rel_err(acc - 1, 2 * dw_ref) < 1e-2


def test_conv2d_gradfix_trains_on_tensor_cores():
    """conv2d_gradfix.conv2d under autograd: forward, input gradient, weight gradient and the double backward of R1 (grad of grad_input w.r.t. the
    weights, under no_weight_gradients for the inner grad) run on the tcgen05 kernels for a stride-1 'same' fp32 convolution and match fp64
    autograd within 1e-2; the launch counter proves the kernels ran."""
    from pasta_gan_b200.torch_utils.ops import conv2d_gradfix as G
    torch.manual_seed(3)
    old = (G.enabled, G.tensor_core_training, G.tensor_core_min_flops, G.tensor_core_r1)
    G.enabled, G.tensor_core_training, G.tensor_core_min_flops, G.tensor_core_r1 = True, True, 0, True
    try:
        x = torch.randn(2, 32, 24, 24); w = torch.randn(48, 32, 3, 3) / 17
        xg, wg = x.to(DEV).requires_grad_(True), w.to(DEV).requires_grad_(True)
        xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
        l0 = capi.launch_count()
        y = G.conv2d(xg, wg, padding=1)
        yr = torch.nn.functional.conv2d(xr, wr, padding=1)
        assert rel_err(y, yr) < 1e-2
        dy = torch.randn_like(yr)
        dx, dw = torch.autograd.grad(y, [xg, wg], dy.float().to(DEV), create_graph=True)
        dxr, dwr = torch.autograd.grad(yr, [xr, wr], dy, create_graph=True)
        assert rel_err(dx, dxr) < 1e-2 and rel_err(dw, dwr) < 1e-2
        assert capi.launch_count() - l0 >= 6                      # 3 weight packs (fwd, dgrad) + fwd + dgrad + wgrad (+ reduce)
        # R1-style second order: penalty = |dL/dx|^2, gradient w.r.t. the weights
        with G.no_weight_gradients():
            gx, = torch.autograd.grad(G.conv2d(xg, wg, padding=1).sum(), xg, create_graph=True)
        gxr, = torch.autograd.grad(torch.nn.functional.conv2d(xr, wr, padding=1).sum(), xr, create_graph=True)
        pw, = torch.autograd.grad(gx.square().sum(), wg)
        pwr, = torch.autograd.grad(gxr.square().sum(), wr)
        assert rel_err(pw, pwr) < 2e-2
    finally:
        G.enabled, G.tensor_core_training, G.tensor_core_min_flops, G.tensor_core_r1 = old
