"""GPU parity of the sm_100a operator kernels (through the C ABI) against
  (1) the golden vectors written by the unmodified reference's impl='ref' path, and
  (2) the CPU oracle on seeded inputs at sizes it finishes in seconds, and
  (3) size-independent properties at the BASELINE sizes (linearity, adjointness, DC gain, idempotent clamp).

Tolerance (north_star): 1e-5 relative (max|a-b| / max|b|) for upfirdn2d and bias_act in fp32."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import ops_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope='module')
def ops():
    from pasta_gan_b200.torch_utils.ops import upfirdn2d, bias_act, conv2d_resample, conv2d_gradfix, fma
    import types
    return types.SimpleNamespace(up=upfirdn2d, ba=bias_act, cr=conv2d_resample, cg=conv2d_gradfix, fma=fma)


DEV = 'cuda'


# ----------------------------------------------------------------------------- upfirdn2d vs goldens
@pytest.mark.parametrize('i', range(17))
def test_upfirdn2d_golden(golden, ops, i):
    g = golden('upfirdn2d')
    m = g.meta[i]
    x = g.t(f'{i}/x', device=DEV).requires_grad_(True)
    f = g.t(f'{i}/f', device=DEV) if m.get('f') else None
    y = ops.up.upfirdn2d(x, f, up=m['up'], down=m['down'], padding=m['padding'], flip_filter=m['flip'], gain=m['gain'])
    assert rel_err(y, g.t(f'{i}/y')) < TOL
    dy = g.t(f'{i}/dy', device=DEV).requires_grad_(True)
    dx, = torch.autograd.grad(y, x, dy, create_graph=True)
    assert rel_err(dx, g.t(f'{i}/dx')) < TOL
    ddy, = torch.autograd.grad(dx, dy, g.t(f'{i}/ddx', device=DEV))       # second order (R1 path)
    assert rel_err(ddy, g.t(f'{i}/ddy')) < TOL


def test_upfirdn2d_wrappers_golden(golden, ops):
    g = golden('upfirdn2d')
    x, f4, f12 = g.t('w/x', device=DEV), g.t('w/f4', device=DEV), g.t('w/f12', device=DEV)
    assert rel_err(ops.up.filter2d(x, f4, padding=1, gain=2), g.t('w/filter2d')) < TOL
    assert rel_err(ops.up.filter2d(x, f12, flip_filter=True), g.t('w/filter2d_sep')) < TOL
    assert rel_err(ops.up.upsample2d(x, f4), g.t('w/upsample2d')) < TOL
    assert rel_err(ops.up.upsample2d(x, f12, up=[2, 1], padding=[1, 0]), g.t('w/upsample2d_sep')) < TOL
    assert rel_err(ops.up.downsample2d(x, f4), g.t('w/downsample2d')) < TOL
    assert rel_err(ops.up.downsample2d(x, f12, down=[1, 2], gain=3), g.t('w/downsample2d_sep')) < TOL


# the four call forms of SURVEY.md §8 A1 at PASTA-GAN shapes the oracle still finishes quickly
FORMS = [
    dict(name='filter_after_convT', shape=[2, 16, 129, 129], kw=dict(padding=[1, 1, 1, 1], gain=4)),
    dict(name='filter_before_s2conv', shape=[2, 16, 128, 128], kw=dict(padding=[2, 2, 2, 2], gain=1)),
    dict(name='down2_skip', shape=[2, 16, 128, 128], kw=dict(down=2, padding=[1, 1, 1, 1])),
    dict(name='up2_rgb', shape=[4, 3, 64, 64], kw=dict(up=2, padding=[2, 1, 2, 1], gain=4)),
    dict(name='filter_after_convT_257', shape=[1, 4, 257, 257], kw=dict(padding=[1, 1, 1, 1], gain=4)),
    dict(name='filter_before_s2conv_256', shape=[1, 4, 256, 256], kw=dict(padding=[2, 2, 2, 2])),
    dict(name='filter_513', shape=[1, 2, 513, 513], kw=dict(padding=[1, 1, 1, 1], gain=4)),
    dict(name='down2_512', shape=[1, 2, 512, 512], kw=dict(down=2, padding=[1, 1, 1, 1])),
    dict(name='bwd_of_down2', shape=[2, 8, 64, 64], kw=dict(up=2, padding=[2, 1, 2, 1], gain=1, flip_filter=True)),
    # polyphase up-2 band kernel: odd pads (both tap parities), ragged widths (scalar stores), many bands, a general (non rank-1) filter below
    dict(name='up2_rgb_128', shape=[2, 3, 128, 128], kw=dict(up=2, padding=[2, 1, 2, 1], gain=4)),
    dict(name='up2_oddpad', shape=[2, 5, 37, 45], kw=dict(up=2, padding=[1, 2, 3, 0], gain=4)),
    dict(name='up2_ragged', shape=[1, 4, 33, 21], kw=dict(up=2, padding=[2, 2, 2, 2], gain=2, flip_filter=True)),
    dict(name='up2_bigpad', shape=[1, 3, 20, 20], kw=dict(up=2, padding=[5, 4, 4, 6], gain=4)),
    dict(name='small_33', shape=[3, 5, 33, 33], kw=dict(padding=[1, 1, 1, 1], gain=4)),
    dict(name='tiny_9', shape=[2, 7, 9, 9], kw=dict(padding=[1, 1, 1, 1], gain=4)),
    dict(name='nonsquare', shape=[2, 3, 40, 70], kw=dict(padding=[2, 2, 2, 2])),
    dict(name='nonsquare_down', shape=[2, 3, 50, 96], kw=dict(down=2, padding=[1, 1, 1, 1])),
]


@pytest.mark.parametrize('form', FORMS, ids=[f['name'] for f in FORMS])
@pytest.mark.parametrize('dtype', [torch.float32, torch.float16, torch.float64])
def test_upfirdn2d_forms_vs_oracle(ops, form, dtype):
    import zlib; torch.manual_seed(zlib.crc32(form["name"].encode()) % 1000)
    f = O.setup_filter([1, 3, 3, 1])
    x = torch.randn(*form['shape'])
    xr = x.to(dtype).double().requires_grad_(True)                 # oracle in fp64 on the same (rounded) inputs
    y_ref = O.upfirdn2d(xr, f.double().float(), **form['kw'])
    xg = x.to(DEV, dtype).requires_grad_(True)
    y = ops.up.upfirdn2d(xg, f.to(DEV), **form['kw'])
    assert y.dtype == dtype and y.shape == y_ref.shape
    tol = {torch.float32: TOL, torch.float16: 2e-3, torch.float64: 1e-12}[dtype]
    assert rel_err(y, y_ref) < tol
    dy = torch.randn_like(y_ref)
    gx_ref, = torch.autograd.grad(y_ref, xr, dy)
    gx, = torch.autograd.grad(y, xg, dy.to(DEV, dtype))
    assert rel_err(gx, gx_ref) < tol


def test_upfirdn2d_channels_last_and_strided(ops):
    torch.manual_seed(3)
    f = O.setup_filter([1, 3, 3, 1])
    x = torch.randn(2, 6, 40, 40)
    ref = O.upfirdn2d(x, f, padding=[2, 2, 2, 2])
    y = ops.up.upfirdn2d(x.to(DEV).to(memory_format=torch.channels_last), f.to(DEV), padding=[2, 2, 2, 2])
    assert y.is_contiguous(memory_format=torch.channels_last)
    assert rel_err(y, ref) < TOL
    xs = torch.randn(2, 6, 40, 80).to(DEV)[:, :, :, ::2]            # non-dense input view
    assert rel_err(ops.up.upfirdn2d(xs, f.to(DEV), down=2, padding=[1, 1, 1, 1]), O.upfirdn2d(xs.cpu(), f, down=2, padding=[1, 1, 1, 1])) < TOL


def test_upfirdn2d_errors(ops):
    f = O.setup_filter([1, 3, 3, 1]).to(DEV)
    x = torch.randn(1, 1, 2, 2, device=DEV)
    with pytest.raises(RuntimeError, match='at least 1x1'):
        ops.up.upfirdn2d(x, f)                                          # 2x2 input, 4x4 filter, no padding
    with pytest.raises(RuntimeError, match='same device'):
        ops.up.upfirdn2d(torch.randn(1, 1, 8, 8, device=DEV), f.cpu())
    with pytest.raises(AssertionError):
        ops.up.upfirdn2d(torch.randn(8, 8, device=DEV), f)


def test_upfirdn2d_properties_at_baseline_size(ops):
    """[16, 64, 257, 257] -> 256^2 (the largest call of the 256x192 generator, 270 MB in): linearity, DC gain and
    adjointness <A x, y> == <x, A^T y> — properties that do not need the oracle to finish at this size."""
    f = O.setup_filter([1, 3, 3, 1]).to(DEV)
    torch.manual_seed(5)
    x1 = torch.randn(16, 64, 257, 257, device=DEV)
    x2 = torch.randn(16, 64, 257, 257, device=DEV)
    A = lambda t: ops.up.upfirdn2d(t, f, padding=[1, 1, 1, 1], gain=4)
    y1, y2 = A(x1), A(x2)
    assert y1.shape == (16, 64, 256, 256)
    assert rel_err(A(x1 + 2 * x2), y1 + 2 * y2) < TOL
    ones = torch.ones(1, 1, 257, 257, device=DEV)
    assert torch.allclose(A(ones)[..., 2:-2, 2:-2], torch.full((1, 1, 252, 252), 4.0, device=DEV), atol=1e-5)
    x1.requires_grad_(True)
    y = A(x1)
    w = torch.randn_like(y)
    gx, = torch.autograd.grad(y, x1, w)
    lhs = (y.double() * w.double()).sum()
    rhs = (x1.double() * gx.double()).sum()
    assert abs(float(lhs - rhs)) / abs(float(lhs)) < 1e-6
    # spot-check one plane against the oracle
    assert rel_err(y1[3, 17:18].unsqueeze(0), O.upfirdn2d(x1.detach()[3, 17:18].unsqueeze(0).cpu(), f.cpu(), padding=[1, 1, 1, 1], gain=4)) < TOL


# ----------------------------------------------------------------------------- bias_act
@pytest.mark.parametrize('i', range(28))
def test_bias_act_golden(golden, ops, i):
    g = golden('bias_act')
    m = g.meta[i]
    x = g.t(f'{i}/x', device=DEV).requires_grad_(True)
    b = g.t(f'{i}/b', device=DEV).requires_grad_(True) if m['bias'] else None
    y = ops.ba.bias_act(x, b, dim=m['dim'], act=m['act'], alpha=m['alpha'], gain=m['gain'], clamp=m['clamp'])
    assert rel_err(y, g.t(f'{i}/y')) < TOL
    dy = g.t(f'{i}/dy', device=DEV).requires_grad_(True)
    grads = torch.autograd.grad(y, [x] + ([b] if b is not None else []), dy, create_graph=True)
    assert rel_err(grads[0], g.t(f'{i}/dx')) < TOL
    if b is not None:
        assert rel_err(grads[1], g.t(f'{i}/db')) < 2e-5            # a reduction over N*H*W in a different order
    if g.has(f'{i}/d_dy') and grads[0].requires_grad:
        gg = torch.autograd.grad(grads[0], [dy, x], g.t(f'{i}/ddx', device=DEV), allow_unused=True)
        assert rel_err(gg[0], g.t(f'{i}/d_dy')) < TOL
        ref = g.t(f'{i}/d_x')
        if float(ref.abs().max()) == 0:
            assert gg[1] is None or float(gg[1].abs().max()) == 0
        else:
            assert rel_err(gg[1], ref) < 2e-5


BA_CONFIGS = [  # the configurations PASTA-GAN actually issues
    dict(act='lrelu', gain=2 ** 0.5, clamp=256, bias=True),
    dict(act='lrelu', gain=1.0, clamp=256 * 0.5 ** 0.5, bias=True),
    dict(act='linear', gain=1, clamp=256, bias=True),
    dict(act='linear', gain=0.5 ** 0.5, clamp=None, bias=False),
    dict(act='relu', gain=None, clamp=None, bias=False),
    dict(act='sigmoid', gain=None, clamp=None, bias=True),
]


@pytest.mark.parametrize('cfg', BA_CONFIGS, ids=lambda c: f"{c['act']}-{c['gain']}")
@pytest.mark.parametrize('shape', [[4, 64, 64, 64], [2, 3, 33, 31], [5, 512], [1, 7, 1, 1], [3, 16, 5, 5]])
@pytest.mark.parametrize('dtype', [torch.float32, torch.float16])
def test_bias_act_vs_oracle(ops, cfg, shape, dtype):
    torch.manual_seed(11)
    x = (torch.randn(*shape) * 200 if cfg['clamp'] else torch.randn(*shape) * 2).to(dtype)
    b = torch.randn(shape[1]).to(dtype) if cfg['bias'] else None
    xr = x.double().requires_grad_(True)
    br = b.double() if b is not None else None
    y_ref = O.bias_act(xr, br, act=cfg['act'], gain=cfg['gain'], clamp=cfg['clamp'])
    xg = x.to(DEV).requires_grad_(True)
    y = ops.ba.bias_act(xg, b.to(DEV) if b is not None else None, act=cfg['act'], gain=cfg['gain'], clamp=cfg['clamp'])
    tol = TOL if dtype == torch.float32 else 2e-3
    assert y.dtype == dtype and rel_err(y, y_ref) < tol
    dy = torch.randn_like(y_ref)
    gr, = torch.autograd.grad(y_ref, xr, dy)
    gg, = torch.autograd.grad(y, xg, dy.to(DEV, dtype))
    if dtype == torch.float16 and cfg['clamp']:
        # the clamp mask is evaluated on the STORED y (fp16, as in the reference plugin, bias_act.cu:136-142): outputs that round onto
        # the clamp value get a zero gradient, so elements within fp16 rounding distance of the clamp are excluded from the comparison
        keep = ((y_ref.detach().abs() - cfg['clamp']).abs() > 0.01 * cfg['clamp']).to(gr.dtype)
        gr, gg = gr * keep, gg.cpu().double() * keep
    assert rel_err(gg, gr) < (TOL if dtype == torch.float32 else 5e-3)


def test_bias_act_channels_last_dims_and_empty(ops):
    torch.manual_seed(2)
    x = torch.randn(2, 8, 6, 6)
    b = torch.randn(8)
    ref = O.bias_act(x, b, act='lrelu', clamp=1.0)
    y = ops.ba.bias_act(x.to(DEV).to(memory_format=torch.channels_last), b.to(DEV), act='lrelu', clamp=1.0)
    assert y.is_contiguous(memory_format=torch.channels_last) and rel_err(y, ref) < TOL
    b3 = torch.randn(6)
    assert rel_err(ops.ba.bias_act(x.to(DEV), b3.to(DEV), dim=3, act='relu'), O.bias_act(x, b3, dim=3, act='relu')) < TOL
    e = ops.ba.bias_act(torch.empty(0, 8, 4, 4, device=DEV), b.to(DEV), act='lrelu')
    assert e.shape == (0, 8, 4, 4)
    with pytest.raises(RuntimeError, match='wrong number of elements'):
        ops.ba.bias_act(x.to(DEV), torch.randn(5, device=DEV), act='lrelu')
    # unaligned base pointer (odd offset view, flattened) takes the scalar kernel
    flat = torch.randn(1001, device=DEV)[1:]
    assert rel_err(ops.ba.bias_act(flat, act='lrelu'), O.bias_act(flat.cpu(), act='lrelu')) < TOL


def test_bias_act_properties_at_baseline_size(ops):
    """[16, 64, 256, 256] (268 MB, the biggest bias_act of the generator): clamp is idempotent, lrelu is positively
    homogeneous, and a CPU-checked slice matches."""
    torch.manual_seed(7)
    x = torch.randn(16, 64, 256, 256, device=DEV) * 100
    b = torch.randn(64, device=DEV)
    y = ops.ba.bias_act(x, b, act='lrelu', clamp=256)
    assert float(y.abs().max()) <= 256
    assert torch.equal(ops.ba.bias_act(y, None, act='linear', gain=1, clamp=256), y)
    y2 = ops.ba.bias_act(2 * x, 2 * b, act='lrelu', clamp=None)
    y1 = ops.ba.bias_act(x, b, act='lrelu', clamp=None)
    assert rel_err(y2, 2 * y1) < 1e-6
    assert rel_err(y[5, 9:11], O.bias_act(x[5:6, 9:11].cpu(), b[9:11].cpu(), act='lrelu', clamp=256)[0]) < TOL


# ----------------------------------------------------------------------------- fused FIR + bias_act
@pytest.mark.parametrize('form', FORMS[:3] + FORMS[9:11], ids=[f['name'] for f in FORMS[:3] + FORMS[9:11]])
@pytest.mark.parametrize('act', ['lrelu', 'linear', 'relu'])
def test_fused_upfirdn2d_bias_act(ops, form, act):
    torch.manual_seed(21)
    f = O.setup_filter([1, 3, 3, 1])
    x = torch.randn(*form['shape']) * 3
    b = torch.randn(form['shape'][1])
    xr, br = x.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y_ref = O.bias_act(O.upfirdn2d(xr, f, **form['kw']), br, act=act, clamp=4.0)
    xg, bg = x.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
    y = ops.up.upfirdn2d_bias_act(xg, f.to(DEV), bg, act=act, clamp=4.0, **form['kw'])
    assert rel_err(y, y_ref) < TOL
    dy = torch.randn_like(y_ref)
    gx_ref, gb_ref = torch.autograd.grad(y_ref, [xr, br], dy)
    gx, gb = torch.autograd.grad(y, [xg, bg], dy.to(DEV))
    assert rel_err(gx, gx_ref) < TOL and rel_err(gb, gb_ref) < 5e-5


# ----------------------------------------------------------------------------- conv2d_resample / gradfix
@pytest.mark.parametrize('i', range(12))
def test_conv2d_resample_golden(golden, ops, i):
    """Dense convolutions here are fp32 library calls (TF32 disabled), so the tolerance is fp32 round-off."""
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        g = golden('conv2d_resample')
        m = g.meta[i]
        x = g.t(f'{i}/x', device=DEV).requires_grad_(True)
        w = g.t(f'{i}/w', device=DEV).requires_grad_(True)
        f = None if m.get('nofilter') else g.t('f4', device=DEV)
        y = ops.cr.conv2d_resample(x, w, f=f, up=m['up'], down=m['down'], padding=m['padding'], groups=m['groups'], flip_weight=m['flip_weight'])
        assert rel_err(y, g.t(f'{i}/y')) < 5e-5
        dx, dw = torch.autograd.grad(y, [x, w], g.t(f'{i}/dy', device=DEV))
        assert rel_err(dx, g.t(f'{i}/dx')) < 5e-5
        assert rel_err(dw, g.t(f'{i}/dw')) < 5e-5
    finally:
        torch.backends.cudnn.allow_tf32 = old


def test_conv2d_gradfix_double_backward_and_no_weight_gradients(ops):
    """R1 shape of use: grad of sum(D(x)) wrt x under no_weight_gradients, then backward through that gradient."""
    torch.manual_seed(9)
    old_tf32, old_en = torch.backends.cudnn.allow_tf32, ops.cg.enabled
    torch.backends.cudnn.allow_tf32 = False
    ops.cg.enabled = True
    try:
        x = torch.randn(2, 3, 12, 12, dtype=torch.float64)
        w = torch.randn(4, 3, 3, 3, dtype=torch.float64)
        xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
        yr = torch.nn.functional.conv2d(xr, wr, padding=1, stride=2)
        gr, = torch.autograd.grad(yr.square().sum(), xr, create_graph=True)
        gr.square().sum().backward()
        xg, wg = x.to(DEV).requires_grad_(True), w.to(DEV).requires_grad_(True)
        yg = ops.cg.conv2d(xg, wg, padding=1, stride=2)
        with ops.cg.no_weight_gradients():
            gg, = torch.autograd.grad(yg.square().sum(), xg, create_graph=True)
        assert rel_err(gg, gr) < 1e-10
        gg.square().sum().backward()
        assert rel_err(wg.grad, wr.grad) < 1e-10 and rel_err(xg.grad, xr.grad) < 1e-10
        # transposed form
        wt = torch.randn(3, 4, 3, 3, dtype=torch.float64)
        ytr = torch.nn.functional.conv_transpose2d(x, wt, stride=2, padding=1)
        ytg = ops.cg.conv_transpose2d(x.to(DEV), wt.to(DEV), stride=2, padding=1)
        assert rel_err(ytg, ytr) < 1e-10
    finally:
        torch.backends.cudnn.allow_tf32, ops.cg.enabled = old_tf32, old_en


# ----------------------------------------------------------------------------- edge cases
def test_empty_and_degenerate_inputs(ops):
    """Empty batches, 1x1 planes, single-channel tensors and the INT32_MAX guard behave like the reference launchers (no crash, right shapes)."""
    f = O.setup_filter([1, 3, 3, 1]).to(DEV)
    e = ops.up.upfirdn2d(torch.empty(0, 3, 8, 8, device=DEV), f, padding=[2, 1, 2, 1], up=2)
    assert e.shape == (0, 3, 16, 16)
    one = torch.randn(1, 1, 1, 1, device=DEV)
    assert rel_err(ops.up.upfirdn2d(one, f, up=2, padding=[2, 1, 2, 1], gain=4), O.upfirdn2d(one.cpu(), f.cpu(), up=2, padding=[2, 1, 2, 1], gain=4)) < TOL
    assert rel_err(ops.ba.bias_act(one, torch.ones(1, device=DEV), act='tanh'), O.bias_act(one.cpu(), torch.ones(1), act='tanh')) < TOL
    x = torch.randn(2, 1, 37, 41, device=DEV)                       # single channel, odd sizes: band kernel with ragged quads
    assert rel_err(ops.up.upfirdn2d(x, f, padding=[2, 2, 2, 2]), O.upfirdn2d(x.cpu(), f.cpu(), padding=[2, 2, 2, 2])) < TOL
    assert rel_err(ops.up.downsample2d(x[:, :, :36, :40].contiguous(), f), O.downsample2d(x.cpu()[:, :, :36, :40], f.cpu())) < TOL
    g = torch.randn(4, 4, device=DEV)                               # a general (rank-4) 4x4 filter takes the non-separable path of the band kernel
    assert rel_err(ops.up.upfirdn2d(x, g, padding=[1, 2, 2, 1], flip_filter=True), O.upfirdn2d(x.cpu(), g.cpu(), padding=[1, 2, 2, 1], flip_filter=True)) < TOL
    xu = torch.randn(2, 2, 24, 30, device=DEV)                      # ... and the polyphase up-2 band kernel has no rank-1 assumption at all
    assert rel_err(ops.up.upfirdn2d(xu, g, up=2, padding=[2, 1, 1, 2], gain=3), O.upfirdn2d(xu.cpu(), g.cpu(), up=2, padding=[2, 1, 1, 2], gain=3)) < TOL
    b = torch.randn(2, device=DEV)                                  # fused bias_act epilogue on the up-2 kernel
    assert rel_err(ops.up.upfirdn2d_bias_act(xu, f, b, up=2, padding=[2, 1, 2, 1], gain=4, act='lrelu', clamp=1.5),
                   O.bias_act(O.upfirdn2d(xu.cpu(), f.cpu(), up=2, padding=[2, 1, 2, 1], gain=4), b.cpu(), act='lrelu', clamp=1.5)) < TOL
    from pasta_gan_b200.torch_utils.ops import conv_igemm
    y = conv_igemm.conv2d_igemm(torch.randn(1, 16, 1, 1, device=DEV), torch.randn(8, 16, 3, 3, device=DEV))    # 1x1 image, 3x3 kernel
    assert y.shape == (1, 8, 1, 1)
    assert not conv_igemm.supported(torch.empty(0, 16, 8, 8, device=DEV), torch.randn(8, 16, 3, 3, device=DEV))  # empty batch -> library path
    with torch.no_grad():
        z = ops.cr.conv2d_resample(torch.empty(0, 16, 8, 8, device=DEV), torch.randn(8, 16, 3, 3, device=DEV), padding=1)
    assert z.shape == (0, 8, 8, 8)


def test_conv_igemm_1x1_image_value(ops):
    from pasta_gan_b200.torch_utils.ops import conv_igemm
    torch.manual_seed(4)
    x = torch.randn(3, 32, 1, 1)
    w = torch.randn(16, 32, 3, 3) / 10
    ref = O._conv(x.double(), w.double(), padding=1)
    assert rel_err(conv_igemm.conv2d_igemm(x.to(DEV), w.to(DEV)), ref) < 2e-3
