"""Patch routing on the GPU: batched replacement for the perspective warps of the reference's data loader.

The reference rectifies ten body-part quadrilaterals of the garment images into 64 x 64 patches and warps them back, one
``cv2.warpPerspective`` call at a time on the CPU (``UvitonDatasetFull*.normalize``, training/dataset.py:838-927; crop geometry
``get_crop``, :751-836).  Here the geometry (a handful of float32 operations and 8 x 8 linear systems per part) stays on the host --
one native call per batch (``pg_patch_crop_transforms``, csrc/pg_patch_geometry.cu; the numpy functions below are the same arithmetic, kept as the
readable statement the tests compare it with) -- and every pixel operation runs in two kernel launches per batch (csrc/pg_patch_route.cu):
``pg_warp_perspective_u8`` (all rectifying warps of all samples, written straight into the channel-concatenated tensors) and
``pg_patch_denorm_u8`` (the back-warp + mask == 255 composite).  The arithmetic is OpenCV's fixed-point bilinear path, restated
bit for bit; there is no CPU fallback.

``PatchRouter.normalize`` keeps the argument order and the 8-tuple of the reference method, with a leading batch axis on every array.
``warp_perspective`` is the single-call analogue of ``cv2.warpPerspective`` for uint8 images.
"""
import numpy as np
import torch

from . import _capi

BORDER_CONSTANT, BORDER_REPLICATE = 0, 1

# (joints of the part, joints of its fall-back or None)  -- training/dataset.py:847-857 and the fall-backs of :756-776
_ORDER = ['cnose', 'cneck', 'rshoulder', 'relbow', 'rwrist', 'lshoulder', 'lelbow', 'lwrist', 'rhip', 'rknee', 'rankle', 'lhip', 'lknee',
          'lankle', 'reye', 'leye', 'rear', 'lear']
_J = {n: i for i, n in enumerate(_ORDER)}
_PARTS = [
    (('lshoulder', 'lhip', 'rhip', 'rshoulder'), None),
    (('lshoulder', 'rshoulder', 'cnose'), ('lshoulder', 'rshoulder', 'rshoulder')),
    (('lshoulder', 'lelbow'), None),
    (('lelbow', 'lwrist'), None),
    (('rshoulder', 'relbow'), None),
    (('relbow', 'rwrist'), None),
    (('lhip', 'lknee'), ('lhip',)),
    (('lknee', 'lankle'), None),
    (('rhip', 'rknee'), ('rhip',)),
    (('rknee', 'rankle'), None),
]
NUM_PARTS = len(_PARTS)
FIRST_LOWER_PART = 6                      # parts 6..9 (the legs) are also cut from the lower-garment image (dataset.py:888)


def _segment_box(p0, p1, half_width):
    """Quadrilateral around the segment p0 -> p1, half_width x |segment| to either side (dataset.py:822-829).  float32 throughout."""
    seg = p1 - p0
    nrm = np.array([-seg[1], seg[0]])
    off = half_width * nrm
    return np.float32([p0 + off, p0 - off, p1 - off, p1 + off])


def part_quadrilateral(keypoints, part, o_h, ar=0.5):
    """Source quadrilateral (4 x 2 float32, in the padded frame) of body part ``part`` for one sample, or None if its joints are not all
    confident (>= 0.1) even after the reference's fall-back.  keypoints: [18, 3] (x, y, confidence), un-padded 192-wide frame."""
    names, fallback = _PARTS[part]
    idx = [_J[n] for n in names]
    if not (keypoints[idx, 2] >= 0.1).all():
        if fallback is None:
            return None
        names = fallback
        idx = [_J[n] for n in names]
        if not (keypoints[idx, 2] >= 0.1).all():
            return None
    pts = np.float32(keypoints[idx, :2])
    pts[:, 0] += 32                                                        # the 192 -> 256 padding of the frame (dataset.py:780)
    if len(names) == 4:
        return pts
    if len(names) == 1:                                                    # hip without knee: drop a vertical to the bottom edge
        return _segment_box(pts[0], np.float32([pts[0][0], o_h - 1]), ar / 2.0)
    if len(names) == 2:
        return _segment_box(pts[0], pts[1], ar / 2.0)
    if names[2] == 'rshoulder':                                            # shoulders without nose: a square above the shoulder line
        seg = pts[1] - pts[0]
        nrm = np.array([-seg[1], seg[0]])
        if nrm[1] > 0.0:
            nrm = -nrm
        return np.float32([pts[0] + nrm, pts[0], pts[1], pts[1] + nrm])
    neck = 0.5 * (pts[0] + pts[1])                                         # head box: from twice the neck-to-nose vector down to the neck
    top = np.float32(neck + 2 * (pts[2] - neck))
    a, b, c, d = _segment_box(top, np.float32(neck), 0.5)
    return np.float32([b, c, d, a])


def perspective_transforms(src, dst):
    """Batched ``cv2.getPerspectiveTransform``: src, dst [J, 4, 2] float32 -> [J, 3, 3] float64.  Same elimination order as OpenCV's LU solver
    (row pivoting on the first largest magnitude, updates ``a += alpha * pivot_row``), vectorised over the J systems; singular systems give zeros
    with M[2,2] = 1."""
    src = np.asarray(src, np.float32).reshape(-1, 4, 2)
    dst = np.asarray(dst, np.float32).reshape(-1, 4, 2)
    J = src.shape[0]
    A = np.zeros((J, 8, 8), np.float64)
    b = np.zeros((J, 8), np.float64)
    A[:, :4, 0] = A[:, 4:, 3] = src[:, :, 0]
    A[:, :4, 1] = A[:, 4:, 4] = src[:, :, 1]
    A[:, :4, 2] = A[:, 4:, 5] = 1
    A[:, :4, 6] = -src[:, :, 0] * dst[:, :, 0]                             # float32 products, widened on assignment
    A[:, :4, 7] = -src[:, :, 1] * dst[:, :, 0]
    A[:, 4:, 6] = -src[:, :, 0] * dst[:, :, 1]
    A[:, 4:, 7] = -src[:, :, 1] * dst[:, :, 1]
    b[:, :4] = dst[:, :, 0]
    b[:, 4:] = dst[:, :, 1]
    ok = np.ones(J, bool)
    rows = np.arange(J)
    eps = np.finfo(np.float64).eps * 100
    for i in range(8):
        k = i + np.argmax(np.abs(A[:, i:, i]), axis=1)
        ok &= np.abs(A[rows, k, i]) >= eps
        Ai, Ak = A[rows, i].copy(), A[rows, k].copy()
        A[rows, i], A[rows, k] = Ak, Ai
        bi, bk = b[rows, i].copy(), b[rows, k].copy()
        b[rows, i], b[rows, k] = bk, bi
        with np.errstate(divide='ignore', invalid='ignore'):
            d = -1.0 / A[:, i, i]
            alpha = A[:, i + 1:, i] * d[:, None]                           # [J, rows below]
            A[:, i + 1:, i + 1:] += alpha[:, :, None] * A[:, i, None, i + 1:]
            b[:, i + 1:] += alpha * b[:, i, None]
    with np.errstate(divide='ignore', invalid='ignore'):
        for i in range(7, -1, -1):
            s = b[:, i].copy()
            for c in range(i + 1, 8):
                s -= A[:, i, c] * b[:, c]
            b[:, i] = s / A[:, i, i]
    b[~ok] = 0.0
    return np.concatenate([b, np.ones((J, 1))], axis=1).reshape(J, 3, 3)


def invert3x3(M):
    """Batched closed-form inverse as ``cv::invert`` evaluates it for 3 x 3 doubles (adjugate times 1/det, zeros when det == 0).  M [..., 3, 3]."""
    S = np.asarray(M, np.float64)
    c00 = S[..., 1, 1] * S[..., 2, 2] - S[..., 1, 2] * S[..., 2, 1]
    c01 = S[..., 1, 0] * S[..., 2, 2] - S[..., 1, 2] * S[..., 2, 0]
    c02 = S[..., 1, 0] * S[..., 2, 1] - S[..., 1, 1] * S[..., 2, 0]
    det = S[..., 0, 0] * c00 - S[..., 0, 1] * c01 + S[..., 0, 2] * c02
    with np.errstate(divide='ignore'):
        d = np.where(det != 0, 1.0 / det, 0.0)
    t = np.empty(S.shape, np.float64)
    t[..., 0, 0] = c00 * d
    t[..., 0, 1] = (S[..., 0, 2] * S[..., 2, 1] - S[..., 0, 1] * S[..., 2, 2]) * d
    t[..., 0, 2] = (S[..., 0, 1] * S[..., 1, 2] - S[..., 0, 2] * S[..., 1, 1]) * d
    t[..., 1, 0] = (S[..., 1, 2] * S[..., 2, 0] - S[..., 1, 0] * S[..., 2, 2]) * d
    t[..., 1, 1] = (S[..., 0, 0] * S[..., 2, 2] - S[..., 0, 2] * S[..., 2, 0]) * d
    t[..., 1, 2] = (S[..., 0, 2] * S[..., 1, 0] - S[..., 0, 0] * S[..., 1, 2]) * d
    t[..., 2, 0] = c02 * d
    t[..., 2, 1] = (S[..., 0, 1] * S[..., 2, 0] - S[..., 0, 0] * S[..., 2, 1]) * d
    t[..., 2, 2] = (S[..., 0, 0] * S[..., 1, 1] - S[..., 0, 1] * S[..., 1, 0]) * d
    return t


def _segment_boxes(p0, p1, half_width):
    """``_segment_box`` for arrays of segments: p0, p1 [B, 2] float32 -> [B, 4, 2] float32 (same float32 operations, element for element)."""
    seg = p1 - p0
    nrm = np.stack([-seg[:, 1], seg[:, 0]], axis=1)
    off = half_width * nrm
    return np.stack([p0 + off, p0 - off, p1 - off, p1 + off], axis=1).astype(np.float32)


def part_quadrilaterals(keypoints, part, o_h, ar=0.5):
    """``part_quadrilateral`` over the batch: keypoints [B, 18, 3] -> (quads [B, 4, 2] float32, valid [B] bool).  Rows of invalid samples are
    unspecified.  Every value is produced by the same float32 operations, in the same order, as the per-sample function (tests compare them)."""
    kp = np.asarray(keypoints)
    B = kp.shape[0]
    names, fallback = _PARTS[part]
    idx = [_J[n] for n in names]
    ok = (kp[:, idx, 2] >= 0.1).all(axis=1)
    pts = np.float32(kp[:, idx, :2])
    pts[:, :, 0] += 32
    if len(names) == 4:
        return pts, ok
    if len(names) == 2:
        quads = _segment_boxes(pts[:, 0], pts[:, 1], ar / 2.0)
        if fallback is not None:                                           # hip without knee: a vertical from the hip to the bottom edge
            fidx = _J[fallback[0]]
            fok = ~ok & (kp[:, fidx, 2] >= 0.1)
            hip = np.float32(kp[:, fidx, :2])
            hip[:, 0] += 32
            foot = np.stack([hip[:, 0], np.full(B, o_h - 1, np.float32)], axis=1).astype(np.float32)
            fq = _segment_boxes(hip, foot, ar / 2.0)
            quads = np.where(fok[:, None, None], fq, quads)
            ok = ok | fok
        return quads, ok
    # head box: from twice the neck-to-nose vector down to the neck; without the nose, a square above the shoulder line
    neck = 0.5 * (pts[:, 0] + pts[:, 1])
    top = np.float32(neck + 2 * (pts[:, 2] - neck))
    box = _segment_boxes(top, np.float32(neck), 0.5)
    quads = box[:, [1, 2, 3, 0]]
    fok = ~ok & (kp[:, [_J['lshoulder'], _J['rshoulder']], 2] >= 0.1).all(axis=1)
    seg = pts[:, 1] - pts[:, 0]
    nrm = np.stack([-seg[:, 1], seg[:, 0]], axis=1)
    nrm = np.where((nrm[:, 1] > 0.0)[:, None], -nrm, nrm)
    fq = np.stack([pts[:, 0] + nrm, pts[:, 0], pts[:, 1], pts[:, 1] + nrm], axis=1).astype(np.float32)
    quads = np.where(fok[:, None, None], fq, quads)
    return np.ascontiguousarray(quads, np.float32), ok | fok


def crop_transforms(keypoints, h, w, o_h, ar=0.5):
    """``get_crop`` for every (sample, part): keypoints [B, 18, 3] -> (M [B,10,3,3], M_inv [B,10,3,3], valid [B,10] bool); invalid parts are zero.
    Vectorised over the batch: ten array-valued quadrilateral constructions and two batched 8 x 8 solves."""
    keypoints = np.asarray(keypoints)
    B = keypoints.shape[0]
    dst = np.float32(np.array([[w, h]]) * np.float32([[0.0, 0.0], [0.0, 1.0], [1.0, 1.0], [1.0, 0.0]]))
    quads = np.empty((B, NUM_PARTS, 4, 2), np.float32)
    valid = np.empty((B, NUM_PARTS), bool)
    for p in range(NUM_PARTS):
        quads[:, p], valid[:, p] = part_quadrilaterals(keypoints, p, o_h, ar)
    M = np.zeros((B, NUM_PARTS, 3, 3), np.float64)
    M_inv = np.zeros((B, NUM_PARTS, 3, 3), np.float64)
    if valid.any():
        q = quads[valid]
        dsts = np.broadcast_to(dst, q.shape)
        M[valid] = perspective_transforms(q, dsts)
        M_inv[valid] = perspective_transforms(dsts, q)
    return M, M_inv, valid


def crop_transforms_native(keypoints, h, w, o_h, ar=0.5):
    """``crop_transforms`` plus the two inversions ``cv2.warpPerspective`` applies, in one native host call (``pg_patch_crop_transforms``):
    keypoints [B, 18, 3] -> (M, M_inv, valid, to_patch = inv(M), to_image = inv(M_inv)); matrices [B,10,3,3] float64, bit-equal to the numpy path."""
    kp = np.ascontiguousarray(keypoints, np.float64)
    if kp.ndim != 3 or kp.shape[1:] != (18, 3):
        raise _capi.PastaB200Error(f'keypoints: expected [B, 18, 3], got {kp.shape}')
    B = kp.shape[0]
    out = np.empty((4, B, NUM_PARTS, 3, 3), np.float64)
    valid = np.empty((B, NUM_PARTS), np.uint8)
    ptr = lambda a: a.ctypes.data
    _capi.check(_capi.load().pg_patch_crop_transforms(ptr(kp), B, int(h), int(w), int(o_h), float(ar), ptr(out[0]), ptr(out[1]), ptr(out[2]), ptr(out[3]),
                                                      ptr(valid)), 'pg_patch_crop_transforms')
    return out[0], out[1], valid.astype(bool), out[2], out[3]


WARP_JOB_DTYPE = np.dtype([('m', '<f8', (9,)), ('src', '<u8'), ('dst', '<u8'), ('src_h', '<i4'), ('src_w', '<i4'), ('src_row_stride', '<i4'),
                           ('src_pix_stride', '<i4'), ('dst_h', '<i4'), ('dst_w', '<i4'), ('dst_row_stride', '<i4'), ('dst_pix_stride', '<i4'),
                           ('channels', '<i4'), ('border', '<i4')])          # pg_warp_job of include/pasta_b200.h (128 bytes)


def _check_u8(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.uint8 and t.is_contiguous()):
        raise _capi.PastaB200Error(f'{name}: expected a contiguous uint8 CUDA tensor (the patch-routing path has no CPU implementation)')


def _upload(arrays, device):
    """Host arrays -> ONE device byte buffer (one H2D copy); returns (buffer, device address of each array).  Every array starts 16-byte aligned."""
    offs, total = [], 0
    for a in arrays:
        offs.append(total)
        total += (a.nbytes + 15) // 16 * 16
    host = np.empty(max(total, 16), np.uint8)
    for a, o in zip(arrays, offs):
        host[o:o + a.nbytes] = np.ascontiguousarray(a).reshape(-1).view(np.uint8)
    buf = torch.from_numpy(host).to(device, non_blocking=False)
    return buf, [buf.data_ptr() + o for o in offs]


def _launch_warps(jobs, jobs_dev, device):
    """jobs: structured array of WARP_JOB_DTYPE records (host), jobs_dev: the device address of its copy.  One launch per 65 535 jobs (gridDim.y)."""
    assert jobs.dtype == WARP_JOB_DTYPE and jobs.dtype.itemsize == 128
    max_pix = int((jobs['dst_h'].astype(np.int64) * jobs['dst_w']).max())
    _capi.require_device()
    for first in range(0, int(jobs.shape[0]), _MAX_JOBS_PER_LAUNCH):
        n = min(_MAX_JOBS_PER_LAUNCH, int(jobs.shape[0]) - first)
        _capi.check(_capi.load().pg_warp_perspective_u8(jobs_dev + first * WARP_JOB_DTYPE.itemsize, n, max_pix, _capi.current_stream(device)),
                    'pg_warp_perspective_u8')


_MAX_JOBS_PER_LAUNCH = 65535


def _jobs(n):
    return np.zeros(n, WARP_JOB_DTYPE)


def _routing_jobs(valid, to_patch, H, W, h, w, groups):
    """Job table of the rectifying warps of a batch.  groups: (source base address, destination base address, destination channels, first part) per
    (image, patch tensor) pair; sources are [B, H, W, 3], destinations [B, h, w, channels] with part p at channels 3 (p - first) ..  One record per
    valid (sample, part >= first) and group, built in one pass over all groups."""
    bi, pi = np.nonzero(valid)
    sels = [np.flatnonzero(pi >= first) for _, _, _, first in groups]
    n = sum(int(x.shape[0]) for x in sels)
    j = np.zeros(n, WARP_JOB_DTYPE)
    if n == 0:
        return j
    idx = np.concatenate(sels)
    b_, p_ = bi[idx].astype(np.uint64), pi[idx]
    per = lambda vals, dt: np.repeat(np.asarray(vals, dt), [int(x.shape[0]) for x in sels])
    src0, dst0 = per([g[0] for g in groups], np.uint64), per([g[1] for g in groups], np.uint64)
    nch, first = per([g[2] for g in groups], np.int64), per([g[3] for g in groups], np.int64)
    j['m'] = to_patch[bi[idx], p_].reshape(-1, 9)
    j['src'] = src0 + b_ * np.uint64(H * W * 3)
    j['dst'] = dst0 + b_ * (np.uint64(h * w) * nch.astype(np.uint64)) + (3 * (p_ - first)).astype(np.uint64)
    j['src_h'], j['src_w'], j['src_row_stride'], j['src_pix_stride'] = H, W, W * 3, 3
    j['dst_h'], j['dst_w'], j['dst_row_stride'], j['dst_pix_stride'] = h, w, w * nch, nch
    j['channels'], j['border'] = 3, BORDER_REPLICATE
    return j


def warp_perspective(src, M, dsize, border_mode=BORDER_CONSTANT):
    """``cv2.warpPerspective(src, M, dsize, flags=INTER_LINEAR, borderMode=border_mode, borderValue=0)`` for a uint8 CUDA image [H, W, C] (C <= 4)
    or [H, W].  dsize = (width, height) as in OpenCV."""
    squeeze = src.dim() == 2
    img = src.unsqueeze(-1) if squeeze else src
    _check_u8(img, 'warp_perspective')
    H, W, C = img.shape
    if C > 4:
        raise _capi.PastaB200Error('warp_perspective: at most 4 channels')
    w, h = int(dsize[0]), int(dsize[1])
    out = torch.empty((h, w, C), dtype=torch.uint8, device=img.device)
    coeffs = invert3x3(np.asarray(M, np.float64).reshape(3, 3))
    j = _jobs(1)
    j['m'][0] = coeffs.reshape(9)
    j['src'], j['dst'] = img.data_ptr(), out.data_ptr()
    j['src_h'], j['src_w'], j['src_row_stride'], j['src_pix_stride'] = H, W, W * C, C
    j['dst_h'], j['dst_w'], j['dst_row_stride'], j['dst_pix_stride'] = h, w, w * C, C
    j['channels'], j['border'] = C, int(border_mode)
    keep, (jobs_dev,) = _upload([j], img.device)
    _launch_warps(j, jobs_dev, img.device)
    del keep                                                               # stream-ordered free: the launch above is already queued on this stream
    return out[..., 0] if squeeze else out


class PatchRouter:
    """Batched ``normalize`` of the reference's datasets (training/dataset.py:838-927)."""

    def __init__(self, box_factor=2, ar=0.5):
        self.box_factor, self.ar = box_factor, ar

    def normalize(self, upper_img, lower_img, upper_clothes_mask, lower_clothes_mask, keypoints, box_factor=None):
        """upper_img, lower_img, upper_clothes_mask, lower_clothes_mask: uint8 CUDA tensors [B, H, W, 3] (masks are 0 / 255 triples);
        keypoints: array [B, 18, 3] (x, y, confidence) in the un-padded frame.  Returns the reference's tuple with a leading batch axis:
        (img [B,h,w,30], img_lower [B,h,w,12], denorm_upper_img [B,H,W,3], denorm_lower_img [B,H,W,3], M_invs [B,10,3,3] float64 (numpy),
        denorm_hand_masks [B,4,H,W,1], clothes_masks [B,h,w,30], clothes_masks_lower [B,h,w,12])."""
        for name, t in (('upper_img', upper_img), ('lower_img', lower_img), ('upper_clothes_mask', upper_clothes_mask), ('lower_clothes_mask', lower_clothes_mask)):
            _check_u8(t, name)
            if t.shape != upper_img.shape or t.dim() != 4 or t.shape[-1] != 3:
                raise _capi.PastaB200Error(f'{name}: expected [B, H, W, 3] like upper_img, got {tuple(t.shape)}')
        box_factor = self.box_factor if box_factor is None else box_factor
        B, H, W, _ = upper_img.shape
        h, w = H // 2 ** box_factor, W // 2 ** box_factor
        dev = upper_img.device
        # get_crop for every (sample, part); cv2.warpPerspective(img, M, ...) walks the patch and samples img at to_patch = inv(M), the back-warp
        # samples the patch at to_image = inv(M_inv)
        M, M_inv, valid, to_patch, to_image = crop_transforms_native(keypoints, h, w, H, self.ar)
        P, PL = NUM_PARTS, NUM_PARTS - FIRST_LOWER_PART
        patch_buf = torch.zeros((B * h * w * 6 * (P + PL),), dtype=torch.uint8, device=dev)      # the four patch tensors in one allocation / one fill
        cut = np.cumsum([0, 3 * P, 3 * P, 3 * PL, 3 * PL]) * (B * h * w)
        img, masks, img_lower, masks_lower = (patch_buf[cut[i]:cut[i + 1]].view(B, h, w, -1) for i in range(4))
        # one record per cv2.warpPerspective call of the reference: (upper image, upper mask) for every valid part, plus (lower image, lower mask) for the legs
        jobs = _routing_jobs(valid, to_patch, H, W, h, w,
                             ((upper_img.data_ptr(), img.data_ptr(), 3 * P, 0), (upper_clothes_mask.data_ptr(), masks.data_ptr(), 3 * P, 0),
                              (lower_img.data_ptr(), img_lower.data_ptr(), 3 * PL, FIRST_LOWER_PART),
                              (lower_clothes_mask.data_ptr(), masks_lower.data_ptr(), 3 * PL, FIRST_LOWER_PART)))
        # the job table, the back-warp matrices and the validity flags of both garments travel in one H2D copy
        v8 = valid.astype(np.uint8)
        keep, (jobs_dev, m_up, m_lo, v_up, v_lo) = _upload([jobs, to_image, to_image[:, FIRST_LOWER_PART:], v8, v8[:, FIRST_LOWER_PART:]], dev)
        if jobs.shape[0]:
            _launch_warps(jobs, jobs_dev, dev)
        lib, stream = _capi.load(), _capi.current_stream(dev)
        denorm_upper = torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev)
        denorm_lower = torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev)
        part_masks = torch.empty((B, P, H, W), dtype=torch.uint8, device=dev)
        _capi.check(lib.pg_patch_denorm_u8(img.data_ptr(), masks.data_ptr(), m_up, v_up, denorm_upper.data_ptr(),
                                           part_masks.data_ptr(), B, P, h, w, H, W, stream), 'pg_patch_denorm_u8')
        _capi.check(lib.pg_patch_denorm_u8(img_lower.data_ptr(), masks_lower.data_ptr(), m_lo, v_lo, denorm_lower.data_ptr(),
                                           None, B, PL, h, w, H, W, stream), 'pg_patch_denorm_u8')
        del keep
        hand_masks = part_masks[:, 2:6].unsqueeze(-1)                      # the four arm parts (dataset.py:903-907)
        return img, img_lower, denorm_upper, denorm_lower, M_inv, hand_masks, masks, masks_lower
