"""Name-keyed procedural weights and synthetic try-on inputs.

Shared by ``tests/golden/gen_golden.py`` (which fills the *reference* networks)
and by the tests / bench (which fill the host-side mirror), so that both sides see
bit-identical parameters without shipping a 180 MB state_dict: every tensor is a
pure function of its dotted name and shape.  Depends on torch only.
"""

import zlib

import torch


def _gen(name):
    g = torch.Generator(device='cpu')
    g.manual_seed(zlib.crc32(name.encode()) & 0x7FFFFFFF)
    return g


def tensor_for(name, shape, dtype=torch.float32):
    """Deterministic value for parameter/buffer ``name``; None = leave untouched."""
    leaf = name.rsplit('.', 1)[-1]
    shape = tuple(shape)
    if leaf in ('resample_filter', 'w_avg') or leaf.startswith('num_batches'):
        return None
    r = torch.randn(shape, generator=_gen(name), dtype=torch.float32)
    if leaf == 'noise_strength':
        v = torch.full(shape, 0.1)
    elif leaf == 'noise_const' or leaf == 'const':
        v = r
    elif name.endswith('affine.bias'):
        v = 1 + 0.1 * r
    elif 'bias' in leaf:
        v = 0.1 * r
    elif name.endswith('linear.weight'):          # torch.nn.Linear inside Dense
        v = r / (shape[1] ** 0.5)
    elif 'weight' in leaf:
        v = r
    else:
        return None
    return v.to(dtype)


@torch.no_grad()
def fill_(module):
    """In-place procedural init of every parameter and buffer of ``module``."""
    for name, t in list(module.named_parameters()) + list(module.named_buffers()):
        v = tensor_for(name, t.shape, t.dtype)
        if v is not None:
            t.copy_(v)
    return module


def fingerprint(module):
    """{name: (sum, abs-sum)} in float64 — lets a test prove two networks hold the same numbers."""
    out = {}
    for name, t in list(module.named_parameters()) + list(module.named_buffers()):
        d = t.detach().double()
        out[name] = (float(d.sum()), float(d.abs().sum()))
    return out


def synth_inputs(batch, res=256, parts_ch=42, parts_res=64, seed=1234, content_w=None, device='cpu'):
    """Synthetic person / garment / pose / parsing tensors of the shapes the full-body
    generator consumes (reference test.py:104-118,
    training_loop_wo_flow_fullbody.py:289-297).  Images are U{0..255}/127.5-1 with white
    (+1) side bands (the 256x192 photo padded to 256x256), pose/patch background -1,
    masks Bernoulli(0.5)."""
    if content_w is None:
        content_w = res * 3 // 4            # 192 of 256, 384 of 512 (reference uses 320 at 512)
    band = (res - content_w) // 2

    def img(idx, ch, r, bands):
        g = torch.Generator(device='cpu')
        g.manual_seed(seed + idx)
        t = torch.randint(0, 256, (batch, ch, r, r), generator=g).float() / 127.5 - 1
        if bands:
            t[..., :band] = 1.0
            t[..., r - band:] = 1.0
        return t

    def mask(idx):
        g = torch.Generator(device='cpu')
        g.manual_seed(seed + idx)
        return (torch.rand((batch, 1, res, res), generator=g) < 0.5).float()

    retain = img(1, 3, res, True)
    skeleton = img(2, 3, res, False)
    skeleton = torch.where(skeleton > 0.8, skeleton, torch.full_like(skeleton, -1.0))   # sparse skeleton on -1
    parts = img(0, parts_ch, parts_res, False)
    d = dict(
        z=torch.zeros(batch, 0),
        c=parts,
        retain=retain,
        pose=torch.cat([skeleton, retain], dim=1),
        denorm_upper_input=img(3, 3, res, True),
        denorm_lower_input=img(4, 3, res, True),
        denorm_upper_mask=mask(5),
        denorm_lower_mask=mask(6),
    )
    return {k: v.to(device) for k, v in d.items()}


def synth_inputs_512(batch, seed=4321, device='cpu'):
    """Inputs of the 512 x 512 generator (reference test_512.py:104-118 shapes): 48-channel 128 px garment patches, retain, pose."""
    d = synth_inputs(batch, res=512, parts_ch=48, parts_res=128, seed=seed, content_w=320, device=device)
    return {k: d[k] for k in ('z', 'c', 'retain', 'pose')}


def synth_inputs_u8(batch, res=256, parts_ch=42, parts_res=64, seed=1234, full_body=True, device='cpu'):
    """uint8 loader tensors as test.py:103 receives them (image, pose skeleton, garment patches, de-normalised clothes and masks)."""
    def u8(idx, ch, r, hi=256):
        g = torch.Generator(device='cpu')
        g.manual_seed(seed + idx)
        return torch.randint(0, hi, (batch, ch, r, r), generator=g, dtype=torch.uint8)
    d = dict(image=u8(1, 3, res), pose=u8(2, 3, res), norm_img=u8(0, parts_ch, parts_res))
    if full_body:
        d.update(denorm_upper_clothes=u8(3, 3, res), denorm_lower_clothes=u8(4, 3, res), denorm_upper_mask=u8(5, 1, res, 2), denorm_lower_mask=u8(6, 1, res, 2))
    return {k: v.to(device) for k, v in d.items()}


_STICK_FIGURE = {   # (x, y) in the un-padded 192 x 256 frame, OpenPose-18 names as the reference orders them (training/dataset.py:859-861)
    'cnose': (96, 30), 'cneck': (96, 50), 'rshoulder': (70, 55), 'relbow': (60, 95), 'rwrist': (55, 130), 'lshoulder': (122, 55), 'lelbow': (132, 95),
    'lwrist': (138, 130), 'rhip': (80, 135), 'rknee': (78, 185), 'rankle': (77, 235), 'lhip': (112, 135), 'lknee': (114, 185), 'lankle': (115, 235),
    'reye': (90, 25), 'leye': (102, 25), 'rear': (85, 28), 'lear': (107, 28)}
_STICK_ORDER = ['cnose', 'cneck', 'rshoulder', 'relbow', 'rwrist', 'lshoulder', 'lelbow', 'lwrist', 'rhip', 'rknee', 'rankle', 'lhip', 'lknee',
                'lankle', 'reye', 'leye', 'rear', 'lear']


def synth_patch_routing_inputs(batch, res=256, seed=77, drop_joints=True):
    """Synthetic inputs of the patch-routing step (training/dataset.py:553-565): garment images and 0 / 255 garment masks as uint8 [B, res, res, 3]
    (numpy), and jittered stick-figure keypoints [B, 18, 3] (x, y, confidence).  Masks are unions of discs so that back-warped masks reach 255;
    with ``drop_joints`` a few joints per sample get confidence 0 to exercise the reference's fall-backs and the invalid-part path."""
    import numpy as np
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:res, 0:res]
    kps = np.zeros((batch, 18, 3), np.float64)
    imgs = rng.integers(0, 256, (2, batch, res, res, 3), dtype=np.uint8)
    masks = np.zeros((2, batch, res, res, 1), np.uint8)
    for b in range(batch):
        for i, n in enumerate(_STICK_ORDER):
            x, y = _STICK_FIGURE[n]
            kps[b, i] = (x + rng.normal(0, 4), y + rng.normal(0, 4), 0.9)
        if drop_joints and b % 4:
            kps[b, rng.choice(18, size=b % 4, replace=False), 2] = 0.0
        for which, (cy, ry) in enumerate(((95, 60), (185, 70))):              # upper garment around the torso and arms, lower around the legs
            for _ in range(6):
                cx, cyy, r = 128 + rng.normal(0, 30), cy + rng.normal(0, ry / 2), rng.uniform(20, 45)
                masks[which, b, :, :, 0] |= ((xx - cx) ** 2 + (yy - cyy) ** 2 < r * r).astype(np.uint8)
    upper_mask, lower_mask = masks[0], masks[1] * (1 - masks[0])
    rgb = lambda m: np.repeat(m, 3, axis=-1) * np.uint8(255)
    return dict(upper_img=imgs[0] * upper_mask, lower_img=imgs[1] * lower_mask, upper_clothes_mask=rgb(upper_mask), lower_clothes_mask=rgb(lower_mask),
                keypoints=kps)
