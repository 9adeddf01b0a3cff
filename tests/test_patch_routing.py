"""Patch-routing perspective warp (reference training/dataset.py:751-927).

CPU: the oracle's restatement of OpenCV's fixed-point warp against golden vectors made by the real ``cv2`` and by the UNMODIFIED reference
``normalize`` / ``get_crop`` (tests/golden/warp.npz, written by tests/golden/gen_warp_golden.py with OpenCV 4.13.0), against ``cv2`` itself where it can be
imported, and against hand-derivable properties; the product's batched host geometry against the oracle's per-call version and the goldens.
GPU: the two kernels, through the C ABI, bit-exact against the goldens and against the oracle."""
import ctypes
import json
import os

import numpy as np
import pytest
import torch

from oracle import warp_oracle as WO
from pasta_gan_b200 import _capi, patch_routing as PR, synthetic


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'warp.npz')
NORM_NAMES = ('img', 'img_lower', 'denorm_upper_img', 'denorm_lower_img', 'M_invs', 'hand_masks', 'clothes_masks', 'clothes_masks_lower')
BORDERS = (('constant', WO.BORDER_CONSTANT), ('replicate', WO.BORDER_REPLICATE))


@pytest.fixture(scope='module')
def golden():
    g = dict(np.load(GOLDEN))
    g['meta'] = json.loads(bytes(g['meta']).decode())
    return g


def _rand_img(rng, h, w, c=3):
    return rng.integers(0, 256, (h, w, c), dtype=np.uint8)


# ------------------------------------------------------------------------------------------------------------------ oracle vs OpenCV goldens (CPU)

def test_oracle_pinned_to_opencv_raw_warps(golden):
    """getPerspectiveTransform (bit-equal doubles) and warpPerspective (bit-equal bytes, both borders, 1 / 3 / 4 channels) as cv2 computed them."""
    assert golden['meta']['opencv'].startswith('4.')
    for t in range(golden['meta']['raw_trials']):
        src, dst, img = golden[f'raw_{t}_src'], golden[f'raw_{t}_dst'], golden[f'raw_{t}_img']
        M = WO.get_perspective_transform(src, dst)
        assert np.array_equal(M, golden[f'raw_{t}_M']), t
        assert np.array_equal(WO.get_perspective_transform(dst, src), golden[f'raw_{t}_Minv']), t
        for name, border in BORDERS:
            want = golden[f'raw_{t}_{name}']
            got = WO.warp_perspective_u8(img, M, (want.shape[1], want.shape[0]), border).reshape(want.shape)
            assert np.array_equal(got, want), (t, name, int((got != want).sum()))


def test_oracle_pinned_to_reference_normalize(golden):
    """The oracle's normalize / get_crop against the outputs of the unmodified reference methods (training/dataset.py:751-927) on the same inputs."""
    kp = golden['norm_keypoints']
    B = kp.shape[0]
    valid = golden['norm_out_valid']
    assert valid.any() and not valid.all()
    wh = np.expand_dims(np.array([64, 64]), 0)
    for b in range(B):
        got = WO.normalize(golden['norm_upper_img'][b], golden['norm_lower_img'][b], golden['norm_upper_clothes_mask'][b],
                           golden['norm_lower_clothes_mask'][b], kp[b], 2)
        for n, v in zip(NORM_NAMES, got):
            v = np.stack(v) if n == 'hand_masks' else np.asarray(v)
            want = golden['norm_out_' + n][b]
            assert v.shape == want.shape and np.array_equal(v, want), (b, n)
        for p in range(10):
            m, _ = WO.get_crop(kp[b], list(WO.BPARTS[p]), wh, 256, 256, 0.5)
            assert (m is not None) == bool(valid[b, p])
            if m is not None:
                assert np.array_equal(m, golden['norm_out_M'][b, p])


def test_golden_inputs_are_the_synthetic_set(golden):
    """The fixture's inputs are what synthetic.synth_patch_routing_inputs(4, seed=9) makes today (so the GPU tests may regenerate them)."""
    d = synthetic.synth_patch_routing_inputs(4, seed=9)
    for k, v in d.items():
        assert np.array_equal(v, golden['norm_' + k]), k


def test_oracle_against_live_cv2():
    """Where cv2 can be imported (the authoring container; not required on the GPU box), fresh random cases beyond the committed ones."""
    cv2 = pytest.importorskip('cv2')
    rng = np.random.default_rng(99)
    modes = {WO.BORDER_CONSTANT: cv2.BORDER_CONSTANT, WO.BORDER_REPLICATE: cv2.BORDER_REPLICATE}
    for trial in range(16):
        H, W, C = [(256, 256, 3), (64, 64, 3), (77, 201, 1), (130, 33, 4)][trial % 4]
        h, w = [(64, 64), (256, 256), (90, 50), (70, 200)][trial % 4]
        img = _rand_img(rng, H, W, C)
        src = np.float32([[0.1 * W, 0.1 * H], [0.05 * W, 0.9 * H], [0.95 * W, 0.85 * H], [0.9 * W, 0.05 * H]] + rng.normal(0, 0.1 * min(H, W), (4, 2)))
        src += np.float32(rng.uniform(-0.5, 0.5, 2) * (W, H)) * (trial % 2)
        dst = np.float32([[0, 0], [0, h], [w, h], [w, 0]])
        M = cv2.getPerspectiveTransform(src, dst)
        assert np.array_equal(WO.get_perspective_transform(src, dst), M)
        for border, mode in modes.items():
            want = cv2.warpPerspective(img, M, (w, h), borderMode=mode).reshape(h, w, C)
            got = WO.warp_perspective_u8(img, M, (w, h), border).reshape(h, w, C)
            assert np.array_equal(got, want), (trial, border, int((got != want).sum()))


# ------------------------------------------------------------------------------------------------------------------ oracle properties (CPU)

def test_oracle_interpolation_table():
    t = WO.bilinear_tab_i().astype(np.int64)
    assert t.shape == (1024, 4)
    assert (t.sum(1) == 1 << 15).all()
    assert t[0].tolist() == [32767, 0, 0, 1]                                  # OpenCV's saturation + sum correction at the integer position
    fy, fx = np.divmod(np.arange(1024), 32)
    closed = np.stack([(32 - fy) * (32 - fx), (32 - fy) * fx, fy * (32 - fx), fy * fx], 1) * 32
    assert (t[1:] == closed[1:]).all()                                       # what csrc/pg_patch_route.cu evaluates instead of a table


def test_oracle_identity_and_translation_are_copies():
    rng = np.random.default_rng(0)
    img = _rand_img(rng, 96, 160)
    for border in (WO.BORDER_CONSTANT, WO.BORDER_REPLICATE):
        assert (WO.warp_perspective_u8(img, np.eye(3), (160, 96), border) == img).all()
    T = np.array([[1, 0, 7], [0, 1, -4], [0, 0, 1.0]])
    out = WO.warp_perspective_u8(img, T, (160, 96), WO.BORDER_CONSTANT)
    assert (out[:-4, 7:] == img[4:, :-7]).all() and out[:, :7].max() == 0 and out[-4:].max() == 0
    rep = WO.warp_perspective_u8(img, T, (160, 96), WO.BORDER_REPLICATE)
    assert (rep[:-4, :7] == img[4:, :1]).all() and (rep[-4:, 7:] == img[-1:, :-7]).all()


def test_oracle_subpixel_values():
    """A shift by k/32 px interpolates with weights (32-k)/32, k/32 exactly: (a*(32-k)*1024 + b*k*1024 + 2^14) >> 15."""
    rng = np.random.default_rng(1)
    img = _rand_img(rng, 8, 64, 1)
    for k in (1, 5, 16, 31):
        T = np.array([[1, 0, -k / 32.0], [0, 1, 0], [0, 0, 1.0]])
        out = WO.warp_perspective_u8(img, T, (64, 8), WO.BORDER_REPLICATE)[:, :-1, 0].astype(np.int64)
        a, b = img[:, :-1, 0].astype(np.int64), img[:, 1:, 0].astype(np.int64)
        assert (out == (a * (32 - k) * 1024 + b * k * 1024 + (1 << 14)) >> 15).all()


def test_oracle_perspective_transform_maps_its_points():
    rng = np.random.default_rng(2)
    for _ in range(20):
        src = np.float32([[30, 40], [20, 200], [180, 220], [200, 30]] + rng.normal(0, 8, (4, 2)))
        dst = np.float32([[0, 0], [0, 64], [64, 64], [64, 0]])
        M = WO.get_perspective_transform(src, dst)
        p = np.c_[src, np.ones(4)] @ M.T
        assert np.abs(p[:, :2] / p[:, 2:] - dst).max() < 1e-9
        assert np.abs(WO.invert3x3(M) @ M - np.eye(3)).max() < 1e-9


# ------------------------------------------------------------------------------------------------------------------ host geometry (CPU)

def test_host_geometry_matches_oracle():
    d = synthetic.synth_patch_routing_inputs(8, seed=5)
    M, M_inv, valid = PR.crop_transforms(d['keypoints'], 64, 64, 256)
    assert valid.any() and not valid.all()                                   # the synthetic set exercises valid, fall-back and invalid parts
    wh = np.expand_dims(np.array([64, 64]), 0)
    for b in range(8):
        for p in range(10):
            m, mi = WO.get_crop(d['keypoints'][b], list(WO.BPARTS[p]), wh, 256, 256, 0.5)
            assert (m is not None) == bool(valid[b, p])
            if m is not None:
                assert np.array_equal(M[b, p], m) and np.array_equal(M_inv[b, p], mi)          # same operations in the same order: bit-equal
                assert np.array_equal(PR.invert3x3(M[b, p]), WO.invert3x3(m))
            else:
                assert not M[b, p].any() and not M_inv[b, p].any()


def test_host_geometry_matches_reference_golden(golden):
    """The product's batched crop geometry against the matrices the unmodified reference get_crop returned (cv2.getPerspectiveTransform inside)."""
    M, M_inv, valid = PR.crop_transforms(golden['norm_keypoints'], 64, 64, 256)
    assert np.array_equal(valid, golden['norm_out_valid'])
    assert np.array_equal(M, golden['norm_out_M'])
    assert np.array_equal(M_inv, golden['norm_out_M_invs'].astype(np.float64))


def test_native_geometry_bit_equal(golden):
    """pg_patch_crop_transforms (host code of the C-ABI library) against the numpy statement of the same arithmetic, the oracle and the reference's
    matrices: valid, fall-back and invalid parts, many random drop-outs, jittered and extreme keypoints."""
    M, M_inv, valid, to_patch, to_image = PR.crop_transforms_native(golden['norm_keypoints'], 64, 64, 256)
    assert np.array_equal(valid, golden['norm_out_valid']) and np.array_equal(M, golden['norm_out_M'])
    assert np.array_equal(M_inv, golden['norm_out_M_invs'].astype(np.float64))
    rng = np.random.default_rng(21)
    for trial, (res, hw) in enumerate([(256, 64), (256, 64), (512, 128), (256, 32)]):
        kp = synthetic.synth_patch_routing_inputs(24, seed=40 + trial, drop_joints=False)['keypoints']
        kp[:, :, :2] *= res / 256.0
        kp[:, :, :2] += rng.normal(0, 12 if trial else 0.0, kp[:, :, :2].shape)
        kp[:, :, 2] = np.where(rng.random((24, 18)) < 0.25, rng.choice([0.0, 0.05, 0.0999]), kp[:, :, 2])
        if trial == 1:
            kp[0, :, :2] = 100.0                                             # every joint on one point: singular systems -> zero matrices, still 'valid'
            kp[1, :, 2] = 0.1                                                # confidence exactly at the threshold counts as confident
        want = PR.crop_transforms(kp, hw, hw, res)
        got = PR.crop_transforms_native(kp, hw, hw, res)
        assert np.array_equal(got[2], want[2])
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]), trial
        assert np.array_equal(got[3], PR.invert3x3(want[0])) and np.array_equal(got[4], PR.invert3x3(want[1])), trial
    with pytest.raises(_capi.PastaB200Error):
        PR.crop_transforms_native(np.zeros((2, 17, 3)), 64, 64, 256)
    # the C entry point itself: bad arguments come back as an error code with a message, an empty batch is a no-op
    lib = _capi.load()
    assert lib.pg_patch_crop_transforms(None, 1, 64, 64, 256, 0.5, None, None, None, None, None) != 0 and b'null pointer' in lib.pg_last_error()
    kp = np.zeros((1, 18, 3)); out = np.zeros((2, 90)); v = np.zeros(10, np.uint8)
    assert lib.pg_patch_crop_transforms(kp.ctypes.data, 1, 0, 64, 256, 0.5, out[0].ctypes.data, out[1].ctypes.data, None, None, v.ctypes.data) != 0
    assert lib.pg_patch_crop_transforms(None, 0, 64, 64, 256, 0.5, None, None, None, None, None) == 0


def test_native_geometry_against_live_cv2():
    """pg_patch_crop_transforms against cv2.getPerspectiveTransform itself on a few hundred random poses (where cv2 can be imported)."""
    cv2 = pytest.importorskip('cv2')
    rng = np.random.default_rng(123)
    kp = synthetic.synth_patch_routing_inputs(40, seed=77, drop_joints=False)['keypoints']
    kp[:, :, :2] += rng.normal(0, 15, kp[:, :, :2].shape)
    kp[:, :, 2] = np.where(rng.random((40, 18)) < 0.2, 0.0, kp[:, :, 2])
    M, M_inv, valid, to_patch, to_image = PR.crop_transforms_native(kp, 64, 64, 256)
    dst = np.float32([[0, 0], [0, 64], [64, 64], [64, 0]])
    checked = 0
    for p in range(PR.NUM_PARTS):
        quads, ok = PR.part_quadrilaterals(kp, p, 256)
        assert np.array_equal(ok, valid[:, p])
        for b in np.flatnonzero(ok):
            q = np.ascontiguousarray(quads[b], np.float32)
            assert np.array_equal(M[b, p], cv2.getPerspectiveTransform(q, dst)), (b, p)
            assert np.array_equal(M_inv[b, p], cv2.getPerspectiveTransform(dst, q)), (b, p)
            assert np.array_equal(to_patch[b, p], cv2.invert(M[b, p])[1]) and np.array_equal(to_image[b, p], cv2.invert(M_inv[b, p])[1])
            checked += 1
    assert checked > 250


def test_host_fallback_parts():
    kp = synthetic.synth_patch_routing_inputs(1, drop_joints=False)['keypoints'][0]
    for joint, part, expect in (('lknee', 6, True), ('cnose', 1, True), ('lelbow', 2, False), ('lhip', 6, False)):
        k = kp.copy()
        k[WO.ORDER.index(joint), 2] = 0.0
        assert (PR.part_quadrilateral(k, part, 256) is not None) == expect


def test_warp_job_struct_is_128_bytes():
    assert ctypes.sizeof(_capi.WarpJob) == 128
    # the numpy record the product builds its job tables with has the layout of the C struct (include/pasta_b200.h: pg_warp_job)
    assert PR.WARP_JOB_DTYPE.itemsize == 128
    for name, _ in _capi.WarpJob._fields_:
        assert PR.WARP_JOB_DTYPE.fields[name][1] == getattr(_capi.WarpJob, name).offset, name


def test_routing_job_table():
    """The one-pass job table against a record-by-record construction (one record per cv2.warpPerspective call of dataset.py:879-890)."""
    kp = synthetic.synth_patch_routing_inputs(5, seed=4)['keypoints']
    M, M_inv, valid, to_patch, _ = PR.crop_transforms_native(kp, 64, 64, 256)
    H = W = 256; h = w = 64
    groups = ((0x10000000, 0x20000000, 30, 0), (0x30000000, 0x40000000, 30, 0), (0x50000000, 0x60000000, 12, 6), (0x70000000, 0x80000000, 12, 6))
    jobs = PR._routing_jobs(valid, to_patch, H, W, h, w, groups)
    want = []
    for src, dst, nch, first in groups:
        for b in range(5):
            for p in range(first, 10):
                if valid[b, p]:
                    want.append((to_patch[b, p].reshape(9), src + b * H * W * 3, dst + b * h * w * nch + 3 * (p - first), H, W, W * 3, 3, h, w, w * nch, nch, 3, 1))
    assert jobs.shape[0] == len(want) > 0
    for j, wnt in zip(jobs, want):
        assert np.array_equal(j['m'], wnt[0])
        assert tuple(int(j[f]) for f in PR.WARP_JOB_DTYPE.names[1:]) == wnt[1:]
    assert PR._routing_jobs(np.zeros((5, 10), bool), to_patch, H, W, h, w, groups).shape == (0,)


def test_host_geometry_batched_equals_per_sample():
    rng = np.random.default_rng(11)
    kp = synthetic.synth_patch_routing_inputs(12, seed=2, drop_joints=False)['keypoints']
    kp[:, :, 2] = np.where(rng.random((12, 18)) < 0.3, 0.0, kp[:, :, 2])       # many missing joints: fall-backs and invalid parts in every sample
    for p in range(PR.NUM_PARTS):
        quads, ok = PR.part_quadrilaterals(kp, p, 256)
        for b in range(12):
            q = PR.part_quadrilateral(kp[b], p, 256)
            assert (q is not None) == bool(ok[b])
            if q is not None:
                assert np.array_equal(q, quads[b])


def test_no_cpu_path():
    with pytest.raises(_capi.PastaB200Error):
        PR.warp_perspective(torch.zeros(8, 8, 3, dtype=torch.uint8), np.eye(3), (8, 8))


# ------------------------------------------------------------------------------------------------------------------ kernels (GPU)

@pytest.mark.gpu
@pytest.mark.parametrize('border', [WO.BORDER_CONSTANT, WO.BORDER_REPLICATE])
def test_warp_perspective_bit_exact(border):
    rng = np.random.default_rng(10 + border)
    for trial in range(12):
        H, W, C = [(256, 256, 3), (64, 64, 3), (100, 37, 1), (33, 130, 4)][trial % 4]
        h, w = [(64, 64), (256, 256), (50, 90), (200, 70)][trial % 4]
        img = _rand_img(rng, H, W, C)
        src = np.float32([[0.1 * W, 0.1 * H], [0.05 * W, 0.9 * H], [0.95 * W, 0.85 * H], [0.9 * W, 0.05 * H]] + rng.normal(0, 0.08 * min(H, W), (4, 2)))
        if trial % 3 == 0:
            src -= np.float32([0.4 * W, 0.3 * H])                              # quadrilateral partly outside the image: border handling
        dst = np.float32([[0, 0], [0, h], [w, h], [w, 0]])
        M = WO.get_perspective_transform(src, dst)
        want = WO.warp_perspective_u8(img, M, (w, h), border)
        got = PR.warp_perspective(torch.from_numpy(img).cuda(), M, (w, h), border).cpu().numpy()
        assert np.array_equal(got, want), (trial, int((got != want).sum()))


@pytest.mark.gpu
def test_warp_perspective_matches_opencv_golden(golden):
    """The warp kernel against bytes written by cv2.warpPerspective itself (no oracle in between)."""
    for t in range(golden['meta']['raw_trials']):
        img, M = golden[f'raw_{t}_img'], golden[f'raw_{t}_M']
        for name, border in BORDERS:
            want = golden[f'raw_{t}_{name}']
            got = PR.warp_perspective(torch.from_numpy(img).cuda(), M, (want.shape[1], want.shape[0]), border).cpu().numpy().reshape(want.shape)
            assert np.array_equal(got, want), (t, name, int((got != want).sum()))


@pytest.mark.gpu
def test_normalize_matches_reference_golden(golden):
    """PatchRouter.normalize against the outputs of the unmodified reference normalize (56 cv2.warpPerspective calls per sample) on the same inputs."""
    B = golden['norm_keypoints'].shape[0]
    dev = {k: torch.from_numpy(golden['norm_' + k]).cuda() for k in ('upper_img', 'lower_img', 'upper_clothes_mask', 'lower_clothes_mask')}
    got = PR.PatchRouter().normalize(dev['upper_img'], dev['lower_img'], dev['upper_clothes_mask'], dev['lower_clothes_mask'], golden['norm_keypoints'], 2)
    for i in (0, 1, 2, 3, 6, 7):
        assert np.array_equal(got[i].cpu().numpy(), golden['norm_out_' + NORM_NAMES[i]]), NORM_NAMES[i]
    assert np.array_equal(got[4], golden['norm_out_M_invs'].astype(np.float64))
    assert np.array_equal(got[5].cpu().numpy(), golden['norm_out_hand_masks'])
    assert int((golden['norm_out_denorm_upper_img'] > 0).sum()) > 1000


@pytest.mark.gpu
def test_warp_perspective_exact_copies():
    rng = np.random.default_rng(3)
    img = _rand_img(rng, 256, 256)
    t = torch.from_numpy(img).cuda()
    assert torch.equal(PR.warp_perspective(t, np.eye(3), (256, 256)), t)
    out = PR.warp_perspective(t, np.array([[1, 0, 70], [0, 1, 0], [0, 0, 1.0]]), (256, 256)).cpu().numpy()   # crosses the 64-column block boundary
    assert (out[:, 70:] == img[:, :-70]).all() and out[:, :70].max() == 0


@pytest.mark.gpu
def test_normalize_matches_oracle():
    B = 6
    d = synthetic.synth_patch_routing_inputs(B, seed=9)
    dev = {k: torch.from_numpy(v).cuda() for k, v in d.items() if k != 'keypoints'}
    got = PR.PatchRouter().normalize(dev['upper_img'], dev['lower_img'], dev['upper_clothes_mask'], dev['lower_clothes_mask'], d['keypoints'], 2)
    claimed = 0
    for b in range(B):
        want = WO.normalize(d['upper_img'][b], d['lower_img'][b], d['upper_clothes_mask'][b], d['lower_clothes_mask'][b], d['keypoints'][b], 2)
        for i in (0, 1, 2, 3, 6, 7):
            assert np.array_equal(got[i][b].cpu().numpy(), want[i]), (b, i)
        assert np.array_equal(got[4][b], np.asarray(want[4], np.float64))
        for k in range(4):
            assert np.array_equal(got[5][b, k].cpu().numpy(), want[5][k]), (b, 'hand mask', k)
        claimed += int((want[2] > 0).sum())
    assert claimed > 1000                                                     # the composite is not trivially empty
