"""Pins the CPU oracle (oracle/ops_oracle.py) to the golden vectors produced by the unmodified
reference's impl='ref' path (tests/golden/gen_golden.py).  Runs without a GPU.

Tolerance: the oracle is a different decomposition of the same fp32 arithmetic, so agreement is to
fp32 round-off: 2e-6 relative (max-abs over max-abs) for single ops, 2e-5 after a convolution."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import ops_oracle as O

TOL = 2e-6
TOL_CONV = 2e-5


def _flt(g, i, meta):
    return g.t(f'{i}/f') if meta.get('f') else None


def test_setup_filter(golden):
    g = golden('upfirdn2d')
    assert rel_err(O.setup_filter([1, 3, 3, 1]), g.t('sf/1331')) < TOL
    assert rel_err(O.setup_filter([1, 2, 3, 1], gain=4, flip_filter=True), g.t('sf/1331_g4_flip')) < TOL
    assert rel_err(O.setup_filter([1, 2, 3, 4, 4, 3, 2, 1], gain=4), g.t('sf/sep8')) < TOL
    assert rel_err(O.setup_filter(None), g.t('sf/none')) < TOL
    assert rel_err(O.setup_filter([[1, 2], [3, 4]], normalize=False), g.t('sf/nonorm2d')) < TOL
    assert O.setup_filter([1, 3, 3, 1]).shape == (4, 4)          # < 8 taps => 2-D outer product
    assert O.setup_filter([1] * 8).shape == (8,)


@pytest.mark.parametrize('i', range(17))
def test_upfirdn2d_fwd_bwd_2nd(golden, i):
    g = golden('upfirdn2d')
    m = g.meta[i]
    x = g.t(f'{i}/x').requires_grad_(True)
    f = _flt(g, i, m)
    kw = dict(up=m['up'], down=m['down'], padding=m['padding'], flip_filter=m['flip'], gain=m['gain'])
    y = O.upfirdn2d(x, f, **kw)
    assert rel_err(y, g.t(f'{i}/y')) < TOL
    dy = g.t(f'{i}/dy').requires_grad_(True)
    dx, = torch.autograd.grad(y, x, dy, create_graph=True)
    assert rel_err(dx, g.t(f'{i}/dx')) < TOL
    ddy, = torch.autograd.grad(dx, dy, g.t(f'{i}/ddx'))
    assert rel_err(ddy, g.t(f'{i}/ddy')) < TOL


def test_upfirdn2d_wrappers(golden):
    g = golden('upfirdn2d')
    x, f4, f12 = g.t('w/x'), g.t('w/f4'), g.t('w/f12')
    assert rel_err(O.filter2d(x, f4, padding=1, gain=2), g.t('w/filter2d')) < TOL
    assert rel_err(O.filter2d(x, f12, flip_filter=True), g.t('w/filter2d_sep')) < TOL
    assert rel_err(O.upsample2d(x, f4), g.t('w/upsample2d')) < TOL
    assert rel_err(O.upsample2d(x, f12, up=[2, 1], padding=[1, 0]), g.t('w/upsample2d_sep')) < TOL
    assert rel_err(O.downsample2d(x, f4), g.t('w/downsample2d')) < TOL
    assert rel_err(O.downsample2d(x, f12, down=[1, 2], gain=3), g.t('w/downsample2d_sep')) < TOL


@pytest.mark.parametrize('i', range(28))
def test_bias_act_all_orders(golden, i):
    g = golden('bias_act')
    m = g.meta[i]
    x = g.t(f'{i}/x').requires_grad_(True)
    b = g.t(f'{i}/b').requires_grad_(True) if m['bias'] else None
    kw = dict(dim=m['dim'], act=m['act'], alpha=m['alpha'], gain=m['gain'], clamp=m['clamp'])
    y = O.bias_act(x, b, **kw)
    assert rel_err(y, g.t(f'{i}/y')) < TOL
    dy = g.t(f'{i}/dy').requires_grad_(True)
    grads = torch.autograd.grad(y, [x] + ([b] if b is not None else []), dy, create_graph=True)
    assert rel_err(grads[0], g.t(f'{i}/dx')) < 1e-5
    if b is not None:
        assert rel_err(grads[1], g.t(f'{i}/db')) < 1e-5
    # closed-form gradient kernels (what the CUDA grad=1 / grad=2 modes compute)
    dx_cf = O.bias_act_grad(1, dy.detach(), x=x.detach(), b=(b.detach() if b is not None else None), y=y.detach(), **kw)
    assert rel_err(dx_cf, g.t(f'{i}/dx')) < 1e-5
    if g.has(f'{i}/d_dy'):
        ddx = g.t(f'{i}/ddx')
        d_dy = O.bias_act_grad(1, ddx, x=x.detach(), b=(b.detach() if b is not None else None), y=y.detach(), **kw)
        assert rel_err(d_dy, g.t(f'{i}/d_dy')) < 1e-5
        d_x = O.bias_act_grad(2, ddx, x=x.detach(), b=(b.detach() if b is not None else None), y=y.detach(), dy=dy.detach(), **kw)
        ref = g.t(f'{i}/d_x')
        if float(ref.abs().max()) == 0:
            assert float(d_x.abs().max()) == 0
        else:
            assert rel_err(d_x, ref) < 2e-5


@pytest.mark.parametrize('i', range(12))
def test_conv2d_resample(golden, i):
    g = golden('conv2d_resample')
    m = g.meta[i]
    x = g.t(f'{i}/x').requires_grad_(True)
    w = g.t(f'{i}/w').requires_grad_(True)
    f = None if m.get('nofilter') else g.t('f4')
    y = O.conv2d_resample(x, w, f=f, up=m['up'], down=m['down'], padding=m['padding'], groups=m['groups'], flip_weight=m['flip_weight'])
    assert rel_err(y, g.t(f'{i}/y')) < TOL_CONV
    dx, dw = torch.autograd.grad(y, [x, w], g.t(f'{i}/dy'))
    assert rel_err(dx, g.t(f'{i}/dx')) < TOL_CONV
    assert rel_err(dw, g.t(f'{i}/dw')) < TOL_CONV


@pytest.mark.parametrize('i', range(10))
def test_modulated_conv2d(golden, i):
    g = golden('modulated_conv2d')
    m = g.meta[i]
    x = g.t(f'{i}/x').requires_grad_(True)
    w = g.t(f'{i}/w').requires_grad_(True)
    s = g.t(f'{i}/s').requires_grad_(True)
    noise = g.t(f'{i}/noise') if g.has(f'{i}/noise') else None
    y = O.modulated_conv2d(x, w, s, noise=noise, up=m['up'], padding=m['k'] // 2, resample_filter=g.t('f4'),
                           demodulate=m['demod'], flip_weight=m['flip_weight'], fused_modconv=m['fused'])
    assert rel_err(y, g.t(f'{i}/y')) < TOL_CONV
    dx, dw, ds = torch.autograd.grad(y, [x, w, s], g.t(f'{i}/dy'))
    assert rel_err(dx, g.t(f'{i}/dx')) < 5e-5
    assert rel_err(dw, g.t(f'{i}/dw')) < 5e-5
    assert rel_err(ds, g.t(f'{i}/ds')) < 5e-5


def test_float64_gradcheck_oracle():
    """The oracle is differentiable to 2nd order in fp64 (what the R1 path needs)."""
    torch.manual_seed(0)
    f = O.setup_filter([1, 3, 3, 1])
    x = torch.randn(1, 2, 6, 6, dtype=torch.float64, requires_grad=True)
    assert torch.autograd.gradcheck(lambda t: O.upsample2d(t, f), (x,))
    assert torch.autograd.gradgradcheck(lambda t: O.downsample2d(t, f), (x,))
    b = torch.randn(2, dtype=torch.float64, requires_grad=True)
    assert torch.autograd.gradcheck(lambda t, bb: O.bias_act(t, bb, act='lrelu', clamp=0.7), (x, b))
    assert torch.autograd.gradgradcheck(lambda t, bb: O.bias_act(t, bb, act='swish'), (x, b))


# ---- the lowered "fast port" used for CPU-baseline timing is pinned to the same vectors
@pytest.mark.parametrize('i', range(17))
def test_fast_port_upfirdn2d(golden, i):
    g = golden('upfirdn2d')
    m = g.meta[i]
    y = O.upfirdn2d_fast(g.t(f'{i}/x'), _flt(g, i, m), up=m['up'], down=m['down'], padding=m['padding'], flip_filter=m['flip'], gain=m['gain'])
    assert rel_err(y, g.t(f'{i}/y')) < TOL


@pytest.mark.parametrize('i', range(12))
def test_fast_port_conv2d_resample(golden, i):
    g = golden('conv2d_resample')
    m = g.meta[i]
    f = None if m.get('nofilter') else g.t('f4')
    y = O.conv2d_resample_fast(g.t(f'{i}/x'), g.t(f'{i}/w'), f=f, up=m['up'], down=m['down'], padding=m['padding'], groups=m['groups'],
                               flip_weight=m['flip_weight'])
    assert rel_err(y, g.t(f'{i}/y')) < TOL_CONV


@pytest.mark.parametrize('i', range(10))
def test_fast_port_modulated_conv2d(golden, i):
    g = golden('modulated_conv2d')
    m = g.meta[i]
    noise = g.t(f'{i}/noise') if g.has(f'{i}/noise') else None
    y = O.modulated_conv2d_fast(g.t(f'{i}/x'), g.t(f'{i}/w'), g.t(f'{i}/s'), noise=noise, up=m['up'], padding=m['k'] // 2,
                                resample_filter=g.t('f4'), demodulate=m['demod'], flip_weight=m['flip_weight'])
    assert rel_err(y, g.t(f'{i}/y')) < TOL_CONV


def test_oracle_helpers_of_the_spade_and_io_rows():
    """The oracle's restatements of the SPADE / IO helpers against the library modules and literal expressions the reference uses
    (nn.InstanceNorm2d: networks.py:4363; get_spade_feat tail :5791-5800; test.py:105-115, :131-135)."""
    import numpy as np
    from oracle import ops_oracle as O
    torch.manual_seed(0)
    x = torch.randn(2, 6, 9, 11, dtype=torch.float64) * 3 + 2
    mean, rstd = O.instance_norm_stats(x)
    ref = torch.nn.InstanceNorm2d(6, affine=False)(x)
    assert torch.allclose((x - mean[:, :, None, None]) * rstd[:, :, None, None], ref, atol=1e-12)
    gamma, beta = torch.randn_like(x), torch.randn_like(x)
    assert torch.allclose(O.spade_norm(x, gamma, beta), ref * (1 + gamma) + beta, atol=1e-12)
    assert torch.allclose(O.spade_norm(x, gamma, beta, act='relu', gain=2 ** 0.5), torch.relu(ref * (1 + gamma) + beta) * 2 ** 0.5, atol=1e-12)
    # get_spade_feat tail, including the "not enough valid pixels" fallback
    feat = torch.randn(3, 4, 8, 8, dtype=torch.float64)
    m1 = (torch.rand(3, 1, 8, 8) > 0.5).double(); m2 = (torch.rand(3, 1, 8, 8) > 0.3).double(); m2[0] = 0
    valid = ((m1 + m2) == 2.0).double(); rest = m1 - valid
    fs = (feat * valid).sum(dim=(2, 3), keepdim=True); cnt = valid.sum(dim=(2, 3), keepdim=True)
    en = (cnt > 10).double(); cnt = cnt * en + (8 * 8) * (1 - en)
    assert torch.equal(O.masked_mean_fill(feat, valid, rest), feat * (1 - rest) + (fs / cnt) * rest)
    # IO expressions
    u8 = torch.arange(256, dtype=torch.uint8).reshape(1, 1, 16, 16)
    assert torch.equal(O.u8_normalize(u8), u8.to(torch.float32) / 127.5 - 1) and torch.equal(O.u8_normalize(u8, False), u8.float())
    img = torch.linspace(-1.3, 1.3, 3 * 8 * 16).reshape(1, 3, 8, 16)
    out = O.image_to_u8_bgr(img, crop=(2, 14))
    g = img.numpy()
    exp = np.clip(((g[0].transpose(1, 2, 0) + 1.0) * 127.5)[:, 2:14, [2, 1, 0]], 0, 255).astype(np.uint8)
    assert out.shape == (1, 8, 12, 3) and np.array_equal(out[0].numpy(), exp)
