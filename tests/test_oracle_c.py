"""The plain-C oracle (oracle/oracle.c, built by `make -C oracle` / __graft_entry__.build()) against the golden vectors of the unmodified
reference, and against the torch-CPU oracle: two independent restatements of the algorithm agree with the reference's impl='ref' output."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT, rel_err
from oracle import ops_oracle as O

F = ctypes.POINTER(ctypes.c_float)


@pytest.fixture(scope='module')
def lib():
    subprocess.run(['make', '-s', '-C', os.path.join(ROOT, 'oracle')], check=True)
    return ctypes.CDLL(os.path.join(ROOT, 'oracle', '_build', 'liboracle.so'))


def ptr(t):
    return ctypes.cast(t.data_ptr(), F) if t is not None else None


def pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


@pytest.mark.parametrize('i', range(17))
def test_c_upfirdn2d_vs_golden(lib, golden, i):
    g = golden('upfirdn2d')
    m = g.meta[i]
    x = g.t(f'{i}/x').contiguous()
    f = g.t(f'{i}/f') if m.get('f') else torch.ones(1, 1)
    if f.ndim == 1:                                      # separable filter: the 2-D outer product is the same operator
        f = torch.outer(f, f)
    f = f.contiguous()
    ref = g.t(f'{i}/y')
    y = torch.empty_like(ref)
    upx, upy = pair(m['up']); dx, dy = pair(m['down'])
    p = m['padding']
    px0, px1, py0, py1 = (p, p, p, p) if isinstance(p, int) else ((p[0], p[0], p[1], p[1]) if len(p) == 2 else p)
    n, c, h, w = x.shape
    rc = lib.orc_upfirdn2d(ptr(x), ptr(f), ptr(y), n, c, h, w, f.shape[0], f.shape[1], upx, upy, dx, dy, px0, px1, py0, py1,
                           int(m['flip']), ctypes.c_float(m['gain']))
    assert rc == 0 and rel_err(y, ref) < 2e-6


@pytest.mark.parametrize('i', range(28))
def test_c_bias_act_vs_golden(lib, golden, i):
    g = golden('bias_act')
    m = g.meta[i]
    if m['dim'] != 1 and len(m['shape']) > 1 and m['bias']:
        step = int(np.prod(m['shape'][m['dim'] + 1:]))
    else:
        step = int(np.prod(m['shape'][2:])) if len(m['shape']) > 1 else 1
    if len(m['shape']) == 1:
        step = 1
    x = g.t(f'{i}/x').contiguous()
    b = g.t(f'{i}/b').contiguous() if m['bias'] else None
    idx, da, dg, _, _ = O.ACT_TABLE[m['act']]
    alpha = da if m['alpha'] is None else m['alpha']
    gain = dg if m['gain'] is None else m['gain']
    clamp = -1 if m['clamp'] is None else m['clamp']
    y = torch.empty_like(x)
    rc = lib.orc_bias_act(ptr(x), ptr(b), None, ptr(y), ctypes.c_int64(x.numel()), (b.numel() if b is not None else 1), ctypes.c_int64(step),
                          0, idx, ctypes.c_float(alpha), ctypes.c_float(gain), ctypes.c_float(clamp))
    assert rc == 0 and rel_err(y, g.t(f'{i}/y')) < 2e-6
    if m['act'] in ('linear', 'relu', 'lrelu'):
        dy = g.t(f'{i}/dy').contiguous()
        dx = torch.empty_like(x)
        rc = lib.orc_bias_act(ptr(dy), None, ptr(y), ptr(dx), ctypes.c_int64(x.numel()), 1, ctypes.c_int64(1), 1, idx,
                              ctypes.c_float(alpha), ctypes.c_float(gain), ctypes.c_float(clamp))
        assert rc == 0 and rel_err(dx, g.t(f'{i}/dx')) < 1e-5


@pytest.mark.parametrize('i', [0, 1, 8, 11])
def test_c_conv2d_vs_golden(lib, golden, i):
    """The plain (no resampling, groups = 1) conv2d_resample cases."""
    g = golden('conv2d_resample')
    m = g.meta[i]
    x, w, ref = g.t(f'{i}/x').contiguous(), g.t(f'{i}/w').contiguous(), g.t(f'{i}/y')
    y = torch.empty_like(ref)
    n, cin, h, wd = x.shape
    cout, _, kh, kw = w.shape
    pad = m['padding']
    rc = lib.orc_conv2d(ptr(x), ptr(w), ptr(y), n, cin, h, wd, cout, kh, kw, 1, pad, pad, int(m['flip_weight']))
    assert rc == 0 and rel_err(y, ref) < 2e-5


def test_c_and_torch_oracles_agree_on_a_down2_layer(lib):
    """FIR(pad 2) -> 3x3 stride-2 conv -> bias_act, composed from the C primitives, equals the torch oracle's conv2d_resample + bias_act."""
    torch.manual_seed(0)
    f = O.setup_filter([1, 3, 3, 1]).contiguous()
    x = torch.randn(1, 4, 10, 12)
    w = torch.randn(5, 4, 3, 3) / 6
    b = torch.randn(5) * 0.1
    ref = O.bias_act(O.conv2d_resample(x, w, f=f, down=2, padding=1), b, act='lrelu', clamp=1.0)
    xf = torch.empty(1, 4, 11, 13)
    assert lib.orc_upfirdn2d(ptr(x), ptr(f), ptr(xf), 1, 4, 10, 12, 4, 4, 1, 1, 1, 1, 2, 2, 2, 2, 0, ctypes.c_float(1.0)) == 0
    yc = torch.empty(1, 5, 5, 6)
    assert lib.orc_conv2d(ptr(xf), ptr(w.contiguous()), ptr(yc), 1, 4, 11, 13, 5, 3, 3, 2, 0, 0, 1) == 0
    y = torch.empty_like(yc)
    assert lib.orc_bias_act(ptr(yc), ptr(b), None, ptr(y), ctypes.c_int64(yc.numel()), 5, ctypes.c_int64(30), 0, 3,
                            ctypes.c_float(0.2), ctypes.c_float(2 ** 0.5), ctypes.c_float(1.0)) == 0
    assert rel_err(y, ref) < 2e-6


# ----------------------------------------------------------------------------- patch routing: the C restatement of the OpenCV warp

def test_c_warp_matches_opencv_golden(lib):
    """orc_get_perspective_transform / orc_warp_perspective_u8 against tests/golden/warp.npz: matrices as bit-equal doubles, images as bytes written by
    cv2.warpPerspective 4.13.0 (both borders, 1 / 3 / 4 channels), and the first warp of the reference's normalize() chain; and against the numpy
    restatement (oracle/warp_oracle.py) on fresh cases."""
    import json
    from oracle import warp_oracle as WO
    g = dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'warp.npz')))
    meta = json.loads(bytes(g['meta']).decode())
    D, U8 = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_ubyte)
    lib.orc_get_perspective_transform.argtypes = [F, F, D]
    lib.orc_warp_perspective_u8.argtypes = [U8, ctypes.c_int, ctypes.c_int, ctypes.c_int, D, U8, ctypes.c_int, ctypes.c_int, ctypes.c_int]

    def persp(src, dst):
        src, dst, m = np.ascontiguousarray(src, np.float32), np.ascontiguousarray(dst, np.float32), np.empty(9, np.float64)
        assert lib.orc_get_perspective_transform(src.ctypes.data_as(F), dst.ctypes.data_as(F), m.ctypes.data_as(D)) == 0
        return m.reshape(3, 3)

    def warp(img, M, w, h, border):
        img = np.ascontiguousarray(img)
        H, W, C = img.shape
        M = np.ascontiguousarray(M, np.float64)
        out = np.empty((h, w, C), np.uint8)
        assert lib.orc_warp_perspective_u8(img.ctypes.data_as(U8), H, W, C, M.ctypes.data_as(D), out.ctypes.data_as(U8), h, w, border) == 0
        return out

    for t in range(meta['raw_trials']):
        M = persp(g[f'raw_{t}_src'], g[f'raw_{t}_dst'])
        assert np.array_equal(M, g[f'raw_{t}_M']) and np.array_equal(persp(g[f'raw_{t}_dst'], g[f'raw_{t}_src']), g[f'raw_{t}_Minv'])
        for name, border in (('constant', 0), ('replicate', 1)):
            want = g[f'raw_{t}_{name}']
            got = warp(g[f'raw_{t}_img'], M, want.shape[1], want.shape[0], border)
            assert np.array_equal(got, want), (t, name, int((got != want).sum()))
    # the reference's own chain: part 0 of sample 0 (torso quadrilateral -> 64 x 64 patch, BORDER_REPLICATE) are channels 0..2 of normalize()'s first output
    assert g['norm_out_valid'][0, 0]
    patch = warp(g['norm_upper_img'][0], g['norm_out_M'][0, 0], 64, 64, 1)
    assert np.array_equal(patch, g['norm_out_img'][0][:, :, 0:3])
    rng = np.random.default_rng(5)
    for trial in range(6):
        img = rng.integers(0, 256, (90, 70, 3), dtype=np.uint8)
        src = np.float32([[5, 8], [3, 80], [60, 85], [66, 4]] + rng.normal(0, 6, (4, 2)))
        dst = np.float32([[0, 0], [0, 48], [40, 48], [40, 0]])
        M = persp(src, dst)
        assert np.array_equal(M, WO.get_perspective_transform(src, dst))
        for border in (0, 1):
            assert np.array_equal(warp(img, M, 40, 48, border), WO.warp_perspective_u8(img, M, (40, 48), border))
