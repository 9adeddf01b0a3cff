"""CPU oracle for the patch-routing perspective warp.  TEST INFRASTRUCTURE ONLY.

Restates what the reference's data loader computes in ``UvitonDatasetFull*.normalize`` (training/dataset.py:838-927) and
``get_crop`` (training/dataset.py:751-836): ten body-part quadrilaterals are rectified into 64 x 64 patches with
``cv2.warpPerspective(..., borderMode=BORDER_REPLICATE)``, and warped back with ``BORDER_CONSTANT`` to composite the
"denormalised" garment images and the hand masks.

Parity status: **PINNED** to OpenCV 4.13.0 and to the unmodified reference methods.  The arithmetic lives in a third-party dependency that is
absent from the reference tree: ``opencv-python`` (the reference pins no version: ``Dockerfile:14`` / ``README.md:15`` say
``pip install opencv-python``).  ``tests/golden/gen_warp_golden.py`` runs, in the authoring container, (a) ``cv2.getPerspectiveTransform`` /
``cv2.warpPerspective`` directly and (b) the reference's own ``UvitonDatasetFull.normalize`` / ``get_crop`` (imported unmodified from
/root/reference) on seeded inputs and commits the results as ``tests/golden/warp.npz``; ``tests/test_patch_routing.py`` holds this restatement
bit-equal to those vectors (matrices as doubles, images as bytes) and, where ``cv2`` is importable, to fresh live calls.
What is restated is OpenCV's published algorithm (modules/imgproc/src/imgwarp.cpp: ``getPerspectiveTransform``, ``warpPerspective`` /
``WarpPerspectiveInvoker``, ``remapBilinear`` with the fixed-point ``BilinearTab_i`` of ``initInterTab2D``; modules/core/src/lapack.cpp /
matrix_decomp.cpp: ``invert`` for 3 x 3, ``LUImpl``), anchored on the reference's own call sites.  Hand-derivable properties are tested as well:
identity / integer-translation warps are exact copies, the interpolation table sums to 2^15, exact bilinear values at 1/32-pixel offsets,
both border modes, and ``getPerspectiveTransform`` mapping its four points.

Details that matter for bit-level agreement with OpenCV and are reproduced here:
  * ``warpPerspective`` inverts the matrix first (closed-form 3 x 3 adjugate in double), then walks the destination in blocks
    of 64 x 16 pixels: ``X0 = M0*xb + M1*y + M2`` at the block's first column ``xb``, then ``(X0 + M0*x1) * (32 / (W0 + M6*x1))``
    for the column offset ``x1`` -- the double rounding sequence depends on that split;
  * coordinates are rounded half-to-even to 1/32 pixel (``INTER_BITS = 5``), the integer part saturates to int16;
  * interpolation weights are the 32 x 32 table of int16 quadruples (scale 2^15, entry (0,0) = {32767, 0, 0, 1} after OpenCV's
    sum correction), the result is ``(sum w*p + 2^14) >> 15``;
  * ``BORDER_REPLICATE`` clamps each of the four taps, ``BORDER_CONSTANT`` substitutes 0 per tap and short-cuts fully outside
    pixels.
A plain-C twin of the two OpenCV pieces lives in oracle/oracle.c (``orc_get_perspective_transform``, ``orc_warp_perspective_u8``), pinned to the same
fixture by tests/test_oracle_c.py.
Only ``tests/`` and bench / smoke checkers may import this file; the product (``pasta-gan_b200/patch_routing.py``) has its own
batched host code and CUDA kernels.
"""

import numpy as np

INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS
INTER_REMAP_COEF_BITS = 15
INTER_REMAP_COEF_SCALE = 1 << INTER_REMAP_COEF_BITS
BORDER_CONSTANT, BORDER_REPLICATE = 0, 1

BPARTS = [
    ["lshoulder", "lhip", "rhip", "rshoulder"],
    ["lshoulder", "rshoulder", "cnose"],
    ["lshoulder", "lelbow"],
    ["lelbow", "lwrist"],
    ["rshoulder", "relbow"],
    ["relbow", "rwrist"],
    ["lhip", "lknee"],
    ["lknee", "lankle"],
    ["rhip", "rknee"],
    ["rknee", "rankle"]]                                                # training/dataset.py:847-857
ORDER = ['cnose', 'cneck', 'rshoulder', 'relbow', 'rwrist', 'lshoulder', 'lelbow', 'lwrist', 'rhip', 'rknee', 'rankle', 'lhip', 'lknee',
         'lankle', 'reye', 'leye', 'rear', 'lear']                      # training/dataset.py:859-861


# ---------------------------------------------------------------------------------------------------------------------
# OpenCV pieces

def bilinear_tab_i():
    """``BilinearTab_i`` of initInterTab2D(INTER_LINEAR, fixpt=true): [32*32][4] int16 weights {tl, tr, bl, br}, index = fy*32 + fx."""
    tab1 = np.empty((INTER_TAB_SIZE, 2), np.float32)
    scale = np.float32(1.0) / np.float32(INTER_TAB_SIZE)
    for i in range(INTER_TAB_SIZE):
        x = np.float32(i) * scale
        tab1[i] = (np.float32(1.0) - x, x)
    itab = np.zeros((INTER_TAB_SIZE * INTER_TAB_SIZE + 2, 4), np.int64)       # two spare entries: the correction loop below peeks past a 2x2 block
    flat = itab.reshape(-1)
    for i in range(INTER_TAB_SIZE):
        for j in range(INTER_TAB_SIZE):
            base = (i * INTER_TAB_SIZE + j) * 4
            isum = 0
            for k1 in range(2):
                for k2 in range(2):
                    v = np.float32(tab1[i, k1] * tab1[j, k2])
                    iv = int(np.clip(np.rint(np.float32(v * np.float32(INTER_REMAP_COEF_SCALE))), -32768, 32767))   # saturate_cast<short>
                    flat[base + k1 * 2 + k2] = iv
                    isum += iv
            if isum != INTER_REMAP_COEF_SCALE:
                diff = isum - INTER_REMAP_COEF_SCALE
                ks2 = 1
                Mk1 = Mk2 = mk1 = mk2 = ks2
                for k1 in range(ks2, ks2 + 2):
                    for k2 in range(ks2, ks2 + 2):
                        if flat[base + k1 * 2 + k2] < flat[base + mk1 * 2 + mk2]:
                            mk1, mk2 = k1, k2
                        elif flat[base + k1 * 2 + k2] > flat[base + Mk1 * 2 + Mk2]:
                            Mk1, Mk2 = k1, k2
                if diff < 0:
                    flat[base + Mk1 * 2 + Mk2] -= diff
                else:
                    flat[base + mk1 * 2 + mk2] -= diff
    return itab[:INTER_TAB_SIZE * INTER_TAB_SIZE].astype(np.int16)


_TAB = None


def _tab():
    global _TAB
    if _TAB is None:
        _TAB = bilinear_tab_i()
    return _TAB


def invert3x3(m):
    """cv::invert on a 3 x 3 CV_64F matrix (closed-form adjugate times 1/det); singular input gives zeros."""
    S = np.asarray(m, np.float64).reshape(3, 3)
    d = (S[0, 0] * (S[1, 1] * S[2, 2] - S[1, 2] * S[2, 1]) - S[0, 1] * (S[1, 0] * S[2, 2] - S[1, 2] * S[2, 0])
         + S[0, 2] * (S[1, 0] * S[2, 1] - S[1, 1] * S[2, 0]))
    t = np.zeros(9, np.float64)
    if d != 0.0:
        d = 1.0 / d
        t[0] = (S[1, 1] * S[2, 2] - S[1, 2] * S[2, 1]) * d
        t[1] = (S[0, 2] * S[2, 1] - S[0, 1] * S[2, 2]) * d
        t[2] = (S[0, 1] * S[1, 2] - S[0, 2] * S[1, 1]) * d
        t[3] = (S[1, 2] * S[2, 0] - S[1, 0] * S[2, 2]) * d
        t[4] = (S[0, 0] * S[2, 2] - S[0, 2] * S[2, 0]) * d
        t[5] = (S[0, 2] * S[1, 0] - S[0, 0] * S[1, 2]) * d
        t[6] = (S[1, 0] * S[2, 1] - S[1, 1] * S[2, 0]) * d
        t[7] = (S[0, 1] * S[2, 0] - S[0, 0] * S[2, 1]) * d
        t[8] = (S[0, 0] * S[1, 1] - S[0, 1] * S[1, 0]) * d
    return t.reshape(3, 3)


def _lu_solve(A, b):
    """cv::solve(DECOMP_LU) on a small CV_64F system: in-place Gaussian elimination with row pivoting, then back substitution (LUImpl)."""
    A = np.array(A, np.float64)
    b = np.array(b, np.float64)
    m = A.shape[0]
    eps = np.finfo(np.float64).eps * 100
    for i in range(m):
        k = i
        for j in range(i + 1, m):
            if abs(A[j, i]) > abs(A[k, i]):
                k = j
        if abs(A[k, i]) < eps:
            return None
        if k != i:
            A[[i, k], i:] = A[[k, i], i:]
            b[[i, k]] = b[[k, i]]
        d = -1.0 / A[i, i]
        for j in range(i + 1, m):
            alpha = A[j, i] * d
            for c in range(i + 1, m):
                A[j, c] += alpha * A[i, c]
            b[j] += alpha * b[i]
    for i in range(m - 1, -1, -1):
        s = b[i]
        for c in range(i + 1, m):
            s -= A[i, c] * b[c]
        b[i] = s / A[i, i]
    return b


def get_perspective_transform(src, dst):
    """cv2.getPerspectiveTransform(src, dst) for 4 float32 point pairs -> 3 x 3 float64 (M[2,2] = 1)."""
    src = np.asarray(src, np.float32).reshape(4, 2)
    dst = np.asarray(dst, np.float32).reshape(4, 2)
    a = np.zeros((8, 8), np.float64)
    b = np.zeros(8, np.float64)
    for i in range(4):
        a[i, 0] = a[i + 4, 3] = src[i, 0]
        a[i, 1] = a[i + 4, 4] = src[i, 1]
        a[i, 2] = a[i + 4, 5] = 1
        a[i, 6] = np.float32(-src[i, 0] * dst[i, 0])          # float product, then widened
        a[i, 7] = np.float32(-src[i, 1] * dst[i, 0])
        a[i + 4, 6] = np.float32(-src[i, 0] * dst[i, 1])
        a[i + 4, 7] = np.float32(-src[i, 1] * dst[i, 1])
        b[i] = dst[i, 0]
        b[i + 4] = dst[i, 1]
    x = _lu_solve(a, b)
    if x is None:
        x = np.zeros(8, np.float64)                            # cv::solve leaves the (zero-initialised by Mat) result on a singular system
    return np.concatenate([x, [1.0]]).reshape(3, 3)


def warp_coords(coeffs, dst_h, dst_w):
    """Fixed-point source coordinates for every destination pixel, as WarpPerspectiveInvoker computes them.
    ``coeffs`` maps destination to source (i.e. already inverted).  Returns (sx, sy) int16-saturated integer parts and the table index."""
    M = np.asarray(coeffs, np.float64).reshape(9)
    bh0 = min(16, dst_h)
    bw0 = min(1024 // bh0, dst_w)
    ys = np.arange(dst_h, dtype=np.float64)[:, None]
    xs = np.arange(dst_w)
    xb = ((xs // bw0) * bw0).astype(np.float64)[None, :]
    x1 = (xs % bw0).astype(np.float64)[None, :]
    X0 = (M[0] * xb + M[1] * ys) + M[2]
    Y0 = (M[3] * xb + M[4] * ys) + M[5]
    W0 = (M[6] * xb + M[7] * ys) + M[8]
    W = W0 + M[6] * x1
    with np.errstate(divide='ignore', invalid='ignore'):
        W = np.where(W != 0, INTER_TAB_SIZE / W, 0.0)
    lo, hi = float(np.iinfo(np.int32).min), float(np.iinfo(np.int32).max)
    fX = np.maximum(lo, np.minimum(hi, (X0 + M[0] * x1) * W))
    fY = np.maximum(lo, np.minimum(hi, (Y0 + M[3] * x1) * W))
    X = np.rint(fX).astype(np.int64)                           # saturate_cast<int>(double) = round half to even; already inside int32
    Y = np.rint(fY).astype(np.int64)
    sx = np.clip(X >> INTER_BITS, -32768, 32767)
    sy = np.clip(Y >> INTER_BITS, -32768, 32767)
    alpha = (Y & (INTER_TAB_SIZE - 1)) * INTER_TAB_SIZE + (X & (INTER_TAB_SIZE - 1))
    return sx, sy, alpha


def remap_bilinear_u8(src, sx, sy, alpha, border):
    """remapBilinear<FixedPtCast<int, uchar, 15>>: src [H, W, C] uint8, integer coordinates + table index per destination pixel."""
    src = np.asarray(src, np.uint8)
    if src.ndim == 2:
        src = src[:, :, None]
    H, W, C = src.shape
    w = _tab()[alpha].astype(np.int64)                         # [h, w, 4]
    s = src.astype(np.int64)

    def tap(yy, xx):
        if border == BORDER_REPLICATE:
            return s[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)]
        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        return np.where(ok[..., None], s[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)], 0)

    acc = (tap(sy, sx) * w[..., 0:1] + tap(sy, sx + 1) * w[..., 1:2] + tap(sy + 1, sx) * w[..., 2:3] + tap(sy + 1, sx + 1) * w[..., 3:4])
    out = (acc + (1 << (INTER_REMAP_COEF_BITS - 1))) >> INTER_REMAP_COEF_BITS
    if border == BORDER_CONSTANT:
        outside = (sx >= W) | (sx + 1 < 0) | (sy >= H) | (sy + 1 < 0)
        out = np.where(outside[..., None], 0, out)
    return np.clip(out, 0, 255).astype(np.uint8)


def warp_perspective_u8(src, M, dsize, border=BORDER_CONSTANT):
    """cv2.warpPerspective(src, M, dsize=(w, h), flags=INTER_LINEAR, borderMode=border, borderValue=0) for uint8 images."""
    w, h = dsize
    sx, sy, alpha = warp_coords(invert3x3(M), h, w)
    out = remap_bilinear_u8(src, sx, sy, alpha, border)
    return out if np.asarray(src).ndim == 3 else out[:, :, 0]


# ---------------------------------------------------------------------------------------------------------------------
# The reference's routing logic (training/dataset.py)

def valid_joints(conf):
    return bool((conf >= 0.1).all())                           # training/dataset.py:748-749


def get_crop(keypoints, bpart, wh, o_w, o_h, ar=1.0):
    """training/dataset.py:751-836.  keypoints [18, 3] (x, y, confidence) in the un-padded 256 x 192 frame.  Returns (M, M_inv) or (None, None)."""
    joints = np.asarray(keypoints)
    idx = [ORDER.index(b) for b in bpart]
    part_src = np.float32(joints[idx][:, :2])
    if not valid_joints(joints[idx][:, 2]):
        if bpart[0] == "lhip" and bpart[1] == "lknee":
            bpart = ["lhip"]
        elif bpart[0] == "rhip" and bpart[1] == "rknee":
            bpart = ["rhip"]
        elif bpart[0] == "lshoulder" and bpart[1] == "rshoulder" and bpart[2] == "cnose":
            bpart = ["lshoulder", "rshoulder", "rshoulder"]
        idx = [ORDER.index(b) for b in bpart]
        part_src = np.float32(joints[idx][:, :2])
    if not valid_joints(joints[idx][:, 2]):
        return None, None
    part_src[:, 0] = part_src[:, 0] + 32
    if part_src.shape[0] == 1:
        a = part_src[0]
        b = np.float32([a[0], o_h - 1])
        part_src = np.float32([a, b])
    if part_src.shape[0] == 4:
        pass
    elif part_src.shape[0] == 3:
        if bpart == ["lshoulder", "rshoulder", "rshoulder"]:
            segment = part_src[1] - part_src[0]
            normal = np.array([-segment[1], segment[0]])
            if normal[1] > 0.0:
                normal = -normal
            a = part_src[0] + normal
            b = part_src[0]
            c = part_src[1]
            d = part_src[1] + normal
            part_src = np.float32([a, b, c, d])
        else:
            neck = 0.5 * (part_src[0] + part_src[1])
            neck_to_nose = part_src[2] - neck
            part_src = np.float32([neck + 2 * neck_to_nose, neck])
            segment = part_src[1] - part_src[0]
            normal = np.array([-segment[1], segment[0]])
            alpha = 1.0 / 2.0
            a = part_src[0] + alpha * normal
            b = part_src[0] - alpha * normal
            c = part_src[1] - alpha * normal
            d = part_src[1] + alpha * normal
            part_src = np.float32([b, c, d, a])
    else:
        segment = part_src[1] - part_src[0]
        normal = np.array([-segment[1], segment[0]])
        alpha = ar / 2.0
        a = part_src[0] + alpha * normal
        b = part_src[0] - alpha * normal
        c = part_src[1] - alpha * normal
        d = part_src[1] + alpha * normal
        part_src = np.float32([a, b, c, d])
    dst = np.float32([[0.0, 0.0], [0.0, 1.0], [1.0, 1.0], [1.0, 0.0]])
    part_dst = np.float32(wh * dst)
    return get_perspective_transform(part_src, part_dst), get_perspective_transform(part_dst, part_src)


def normalize(upper_img, lower_img, upper_clothes_mask, lower_clothes_mask, keypoints, box_factor=2):
    """training/dataset.py:838-927 for one sample.  Images [H, W, 3] uint8 (masks are 0 / 255 RGB triples).  Same return tuple as the reference:
    (img [h,w,30], img_lower [h,w,12], denorm_upper_img, denorm_lower_img, M_invs [10,3,3], denorm_hand_masks (4 x [H,W,1]), clothes_masks [h,w,30],
    clothes_masks_lower [h,w,12])."""
    o_h, o_w = upper_img.shape[:2]
    h, w = o_h // 2 ** box_factor, o_w // 2 ** box_factor
    wh = np.expand_dims(np.array([w, h]), 0)
    ar = 0.5
    part_imgs, part_imgs_lower, part_masks, part_masks_lower, M_invs, hand_masks = [], [], [], [], [], []
    denorm_upper = np.zeros_like(upper_img)
    denorm_lower = np.zeros_like(upper_img)
    for ii, bpart in enumerate(BPARTS):
        part_img = np.zeros((h, w, 3), np.uint8)
        part_img_lower = np.zeros((h, w, 3), np.uint8)
        part_mask = np.zeros((h, w, 3), np.uint8)
        part_mask_lower = np.zeros((h, w, 3), np.uint8)
        M, M_inv = get_crop(keypoints, list(bpart), wh, o_w, o_h, ar)
        patch_mask = None
        if M is not None:
            part_img = warp_perspective_u8(upper_img, M, (w, h), BORDER_REPLICATE)
            part_mask = warp_perspective_u8(upper_clothes_mask, M, (w, h), BORDER_REPLICATE)
            denorm_patch = warp_perspective_u8(part_img, M_inv, (o_w, o_h), BORDER_CONSTANT)
            patch_mask = warp_perspective_u8(part_mask, M_inv, (o_w, o_h), BORDER_CONSTANT)[..., 0:1]
            patch_mask = (patch_mask == 255).astype(np.uint8)
            denorm_upper = denorm_patch * patch_mask + denorm_upper * (1 - patch_mask)
            if ii >= 6:
                part_img_lower = warp_perspective_u8(lower_img, M, (w, h), BORDER_REPLICATE)
                part_mask_lower = warp_perspective_u8(lower_clothes_mask, M, (w, h), BORDER_REPLICATE)
                denorm_patch_lower = warp_perspective_u8(part_img_lower, M_inv, (o_w, o_h), BORDER_CONSTANT)
                patch_mask_lower = warp_perspective_u8(part_mask_lower, M_inv, (o_w, o_h), BORDER_CONSTANT)[..., 0:1]
                patch_mask_lower = (patch_mask_lower == 255).astype(np.uint8)
                denorm_lower = denorm_patch_lower * patch_mask_lower + denorm_lower * (1 - patch_mask_lower)
            M_invs.append(M_inv[np.newaxis, ...])
        else:
            M_invs.append(np.zeros((1, 3, 3), np.float32))
        if 2 <= ii <= 5:
            hand_masks.append(patch_mask if M is not None else np.zeros_like(upper_img)[..., 0:1])
        part_imgs.append(part_img)
        part_masks.append(part_mask)
        if ii >= 6:
            part_imgs_lower.append(part_img_lower)
            part_masks_lower.append(part_mask_lower)
    return (np.concatenate(part_imgs, axis=2), np.concatenate(part_imgs_lower, axis=2), denorm_upper, denorm_lower,
            np.concatenate(M_invs, axis=0), hand_masks, np.concatenate(part_masks, axis=2), np.concatenate(part_masks_lower, axis=2))
