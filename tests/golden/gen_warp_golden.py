#!/usr/bin/env python
"""Generate ``tests/golden/warp.npz``: golden vectors of the patch-routing path from the real thing.

Run (in the authoring container, where /root/reference is mounted read-only and ``cv2`` is importable):

    python tests/golden/gen_warp_golden.py

Two groups of vectors:

  * ``raw_*``   -- ``cv2.getPerspectiveTransform`` / ``cv2.warpPerspective`` (INTER_LINEAR, BORDER_CONSTANT and BORDER_REPLICATE, uint8,
    1 / 3 / 4 channels) called directly on seeded random images and quadrilaterals, some partly outside the image;
  * ``norm_*``  -- the UNMODIFIED reference methods ``UvitonDatasetFull.normalize`` / ``get_crop`` / ``valid_joints``
    (training/dataset.py:748-927), executed as they are on synthetic garment images, masks and stick-figure keypoints
    (pasta-gan_b200/synthetic.py:synth_patch_routing_inputs), including samples whose missing joints take the fall-back and the
    invalid-part branches.

Harness shims only (the reference tree is untouched): ``skimage.draw``, ``pycocotools.mask``, ``matplotlib.pyplot`` are absent from this image
and irrelevant to the three methods, so empty stand-in modules satisfy the module-level imports of training/dataset.py and training/utils.py;
the methods are called unbound on a minimal object that carries ``keypoints`` (which ``__getitem__`` sets from the pose file, dataset.py:744).

The OpenCV build that produced the committed file is recorded in the fixture's ``meta`` (version, algorithm hint).
This script is the only reader of /root/reference and the only user of cv2; the tests read the committed .npz.
"""

import json
import os
import sys
import types

os.environ.setdefault('PYTHONDONTWRITEBYTECODE', '1')
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('PASTA_REFERENCE', '/root/reference')
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

import cv2                                             # noqa: E402
import numpy as np                                     # noqa: E402

for _m in ('matplotlib', 'matplotlib.pyplot', 'skimage', 'skimage.draw', 'pycocotools', 'pycocotools.mask'):
    sys.modules.setdefault(_m, types.ModuleType(_m))
sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
sys.modules['skimage'].draw = sys.modules['skimage.draw']
sys.modules['skimage.draw'].circle = sys.modules['skimage.draw'].line_aa = None
sys.modules['pycocotools'].mask = sys.modules['pycocotools.mask']
import training.dataset as R_ds                        # noqa: E402

from pasta_gan_b200 import synthetic                   # noqa: E402


RAW_SHAPES = [((256, 256, 3), (64, 64)), ((64, 64, 3), (256, 256)), ((100, 37, 1), (50, 90)), ((33, 130, 4), (200, 70))]   # (source H, W, C), (dst h, w)
RAW_TRIALS = 8


def raw_cases():
    """Seeded images and quadrilaterals -> the same calls the reference makes (dataset.py:834-835, :883-897)."""
    out = {}
    rng = np.random.default_rng(2024)
    for t in range(RAW_TRIALS):
        (H, W, C), (h, w) = RAW_SHAPES[t % 4]
        img = rng.integers(0, 256, (H, W, C), dtype=np.uint8)
        src = np.float32([[0.1 * W, 0.1 * H], [0.05 * W, 0.9 * H], [0.95 * W, 0.85 * H], [0.9 * W, 0.05 * H]] + rng.normal(0, 0.08 * min(H, W), (4, 2)))
        if t % 3 == 0:
            src -= np.float32([0.4 * W, 0.3 * H])                              # partly outside the image: border handling
        dst = np.float32([[0, 0], [0, h], [w, h], [w, 0]])
        M = cv2.getPerspectiveTransform(src, dst)
        out[f'raw_{t}_img'] = img
        out[f'raw_{t}_src'] = src
        out[f'raw_{t}_dst'] = dst
        out[f'raw_{t}_M'] = M
        out[f'raw_{t}_Minv'] = cv2.getPerspectiveTransform(dst, src)
        for name, mode in (('constant', cv2.BORDER_CONSTANT), ('replicate', cv2.BORDER_REPLICATE)):
            o = cv2.warpPerspective(img, M, (w, h), borderMode=mode)
            out[f'raw_{t}_{name}'] = o.reshape(h, w, C)
    return out


class _Sample:
    """What the three methods touch of ``self``: the keypoints of the current sample and each other."""
    valid_joints = R_ds.UvitonDatasetFull.valid_joints
    get_crop = R_ds.UvitonDatasetFull.get_crop
    normalize = R_ds.UvitonDatasetFull.normalize

    def __init__(self, keypoints):
        self.keypoints = keypoints


def normalize_cases(batch=4, seed=9):
    d = synthetic.synth_patch_routing_inputs(batch, seed=seed)
    out = {'norm_' + k: v for k, v in d.items()}
    names = ('img', 'img_lower', 'denorm_upper_img', 'denorm_lower_img', 'M_invs', 'hand_masks', 'clothes_masks', 'clothes_masks_lower')
    res = {n: [] for n in names}
    for b in range(batch):
        r = _Sample(d['keypoints'][b]).normalize(d['upper_img'][b], d['lower_img'][b], d['upper_clothes_mask'][b], d['lower_clothes_mask'][b], 2)
        for n, v in zip(names, r):
            res[n].append(np.stack(v) if n == 'hand_masks' else np.asarray(v))
    for n in names:
        out['norm_out_' + n] = np.stack(res[n])
    # the forward matrices too (normalize() returns only the inverse ones)
    wh = np.expand_dims(np.array([64, 64]), 0)
    order = ['cnose', 'cneck', 'rshoulder', 'relbow', 'rwrist', 'lshoulder', 'lelbow', 'lwrist', 'rhip', 'rknee', 'rankle', 'lhip', 'lknee',
             'lankle', 'reye', 'leye', 'rear', 'lear']
    bparts = [["lshoulder", "lhip", "rhip", "rshoulder"], ["lshoulder", "rshoulder", "cnose"], ["lshoulder", "lelbow"], ["lelbow", "lwrist"],
              ["rshoulder", "relbow"], ["relbow", "rwrist"], ["lhip", "lknee"], ["lknee", "lankle"], ["rhip", "rknee"], ["rknee", "rankle"]]
    Ms = np.zeros((batch, 10, 3, 3), np.float64)
    valid = np.zeros((batch, 10), bool)
    for b in range(batch):
        s = _Sample(d['keypoints'][b])
        for p, bp in enumerate(bparts):
            M, _ = s.get_crop(list(bp), order, wh, 256, 256, 0.5)
            if M is not None:
                Ms[b, p], valid[b, p] = M, True
    out['norm_out_M'], out['norm_out_valid'] = Ms, valid
    return out


def main():
    arrays = {}
    arrays.update(raw_cases())
    arrays.update(normalize_cases())
    info = cv2.getBuildInformation()
    hint = [ln.split(':', 1)[1].strip() for ln in info.splitlines() if 'Algorithm Hint' in ln]
    meta = dict(opencv=cv2.__version__, algorithm_hint=hint[0] if hint else None, numpy=np.__version__, raw_trials=RAW_TRIALS,
                reference='training/dataset.py:748-927 (UvitonDatasetFull.valid_joints / get_crop / normalize), unmodified')
    arrays['meta'] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(os.environ.get('PASTA_GOLDEN_OUT', HERE), 'warp.npz')
    np.savez_compressed(path, **arrays)
    print(path, os.path.getsize(path), 'bytes', meta)


if __name__ == '__main__':
    main()
