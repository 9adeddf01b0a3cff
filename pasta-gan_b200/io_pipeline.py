"""Device-side input / output pipeline of the try-on generator.

The reference moves the loader's uint8 tensors to the GPU and normalises them there with one elementwise op per tensor plus a
``torch.cat`` (test.py:105-115), and converts the generated images on the CPU (test.py:131-135: transpose, +1, *127.5, crop to the
256x192 photo, BGR, clip, uint8).  Here both sides are one kernel each (``pg_u8_normalize``, ``pg_image_to_u8_bgr``), bit-identical to
those expressions, so that a batch costs 17 MB of PCIe traffic in and 2.4 MB out instead of 82 MB and 25 MB."""
import ctypes

import torch

from . import capi

# loader tensor -> (generator input, normalised?); ``pose`` is the 3-channel skeleton, concatenated with the person image below
U8_KEYS = ('image', 'pose', 'norm_img', 'denorm_upper_clothes', 'denorm_lower_clothes', 'denorm_upper_mask', 'denorm_lower_mask')


def normalize_u8_batch(u8, out=None):
    """u8: dict of uint8 CUDA tensors named as in test.py:103 (``image`` [N,3,H,W], ``pose`` [N,3,H,W], ``norm_img`` [N,P,h,w] and, for the full-body
    generator, ``denorm_*_clothes`` [N,3,H,W] / ``denorm_*_mask`` [N,1,H,W]).  Returns the generator's float32 keyword inputs
    (``retain``, ``pose`` = skeleton || retain, ``c``, ``denorm_*_input``, ``denorm_*_mask``), written into ``out`` when given."""
    img = u8['image']
    n, _, h, w = (int(v) for v in img.shape)
    dev = img.device
    full = 'denorm_upper_clothes' in u8
    if out is None:
        out = dict(retain=torch.empty([n, 3, h, w], device=dev), pose=torch.empty([n, 6, h, w], device=dev),
                   c=torch.empty(u8['norm_img'].shape, device=dev))
        if full:
            for k in ('denorm_upper_input', 'denorm_lower_input'):
                out[k] = torch.empty([n, 3, h, w], device=dev)
            for k in ('denorm_upper_mask', 'denorm_lower_mask'):
                out[k] = torch.empty([n, 1, h, w], device=dev)
    jobs = []      # (src, dst, rows, row_len, src_stride, dst_stride, normalize)
    whole = lambda s, d, nz: jobs.append((s, d, 1, s.numel(), s.numel(), s.numel(), nz))
    whole(img, out['retain'], 1)
    whole(u8['norm_img'], out['c'], 1)
    hw3 = 3 * h * w
    jobs.append((u8['pose'], out['pose'], n, hw3, hw3, 2 * hw3, 1))                       # pose[:, 0:3] = skeleton
    jobs.append((img, out['pose'][:, 3:], n, hw3, hw3, 2 * hw3, 1))                       # pose[:, 3:6] = retain   (torch.cat, test.py:115)
    if full:
        whole(u8['denorm_upper_clothes'], out['denorm_upper_input'], 1)
        whole(u8['denorm_lower_clothes'], out['denorm_lower_input'], 1)
        whole(u8['denorm_upper_mask'], out['denorm_upper_mask'], 0)
        whole(u8['denorm_lower_mask'], out['denorm_lower_mask'], 0)
    for s, d, *_ in jobs:
        assert s.dtype == torch.uint8 and d.dtype == torch.float32 and s.is_cuda and d.is_cuda and s.is_contiguous()
    k = len(jobs)
    P, I64, I32 = ctypes.c_void_p * k, ctypes.c_int64 * k, ctypes.c_int32 * k
    with torch.cuda.device(dev):
        capi.require_device()
        sp = capi.span('io_pipeline', nbytes=sum(5 * j[2] * j[3] for j in jobs), tag='u8 -> float inputs')
        rc = capi.load().pg_u8_normalize(P(*[j[0].data_ptr() for j in jobs]), P(*[j[1].data_ptr() for j in jobs]), I64(*[j[2] for j in jobs]),
                                         I64(*[j[3] for j in jobs]), I64(*[j[4] for j in jobs]), I64(*[j[5] for j in jobs]), I32(*[j[6] for j in jobs]),
                                         k, capi.current_stream(dev))
        capi.check(rc, 'pg_u8_normalize')
        if sp:
            sp.close()
    return out


def images_to_u8(img, crop=None, out=None):
    """img [N,3,H,W] float32 in [-1, 1] -> [N,H,x1-x0,3] uint8 BGR (test.py:131-135); ``crop`` = (x0, x1) columns, default the 3:4 photo
    (32:224 of 256)."""
    n, c, h, w = (int(v) for v in img.shape)
    assert c == 3 and img.dtype == torch.float32 and img.is_cuda
    x0, x1 = crop if crop is not None else (w // 8, w - w // 8)
    img = img.contiguous()
    if out is None:
        out = torch.empty([n, h, x1 - x0, 3], dtype=torch.uint8, device=img.device)
    with torch.cuda.device(img.device):
        capi.require_device()
        sp = capi.span('io_pipeline', nbytes=15 * n * h * (x1 - x0), tag='float image -> u8 BGR')
        rc = capi.load().pg_image_to_u8_bgr(capi.ptr(img), capi.ptr(out), n, h, w, x0, x1, capi.current_stream(img.device))
        capi.check(rc, 'pg_image_to_u8_bgr')
        if sp:
            sp.close()
    return out
