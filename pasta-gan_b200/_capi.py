"""ctypes binding of libpasta_b200.so (include/pasta_b200.h) — the only way the Python layer reaches
the GPU kernels.  Replaces the reference's pybind11 plugin modules (torch_utils/custom_ops.py:46-124,
upfirdn2d.cpp:98-101, bias_act.cpp:94-97).  There is NO fallback: a missing library, a missing symbol
or a non-sm_100 device raises instead of routing anywhere else."""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get('PASTA_B200_LIB') or os.path.join(_HERE, 'lib', 'libpasta_b200.so')   # override: the -DPG_DEBUG build of tools/

_lib = None
_lock = threading.Lock()
_device_checked = False

c_int = ctypes.c_int
c_i32 = ctypes.c_int32
c_i64 = ctypes.c_int64
c_f32 = ctypes.c_float
c_ptr = ctypes.c_void_p
I32x4 = ctypes.c_int32 * 4
I64x4 = ctypes.c_int64 * 4

# name -> argtypes; must list every symbol include/pasta_b200.h declares (tests check this)
SIGNATURES = {
    'pg_abi_version': [],
    'pg_last_error': [],
    'pg_check_device': [],
    'pg_launch_count': [],
    'pg_bias_act': [c_ptr] * 6 + [c_i64, c_i32, c_i64, c_i32, c_i32, c_f32, c_f32, c_f32, c_i32, c_ptr],
    'pg_upfirdn2d': [c_ptr, c_ptr, c_ptr, I32x4, I64x4, I32x4, I64x4, c_i32, c_i32, c_i64, c_i64] + [c_i32] * 8 +
                    [c_i32, c_f32, c_i32, c_ptr],
    'pg_conv2d_igemm_workspace_bytes': [c_i32, c_i32, c_i32, c_i32],
    'pg_conv2d_igemm_workspace_bytes_fmt': [c_i32, c_i32, c_i32, c_i32, c_i32],
    'pg_conv2d_igemm_prepack': [c_ptr, c_ptr, c_f32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_ptr, c_i64, c_ptr],
    'pg_conv2d_igemm_run': [c_ptr] * 5 + [c_i64, c_ptr, c_ptr] + [c_i32] * 7 + [c_i32, c_f32, c_f32, c_i32, c_f32, c_f32, c_f32, c_i32, c_ptr],
    'pg_conv2d_igemm_run2': [c_ptr, c_ptr, c_i32] + [c_ptr] * 4 + [c_i64, c_ptr, c_ptr, c_ptr] + [c_i32] * 7 + [c_i32, c_f32, c_f32, c_i32, c_f32, c_f32, c_f32, c_i32, c_i32, c_i32, c_ptr],
    'pg_u8_normalize': [c_ptr] * 7 + [c_i32, c_ptr],
    'pg_image_to_u8_bgr': [c_ptr, c_ptr, c_i32, c_i32, c_i32, c_i32, c_i32, c_ptr],
    'pg_masked_plane_sum': [c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_i64, c_ptr],
    'pg_masked_fill': [c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_i64, c_i64, c_i32, c_ptr],
    'pg_instance_norm_stats': [c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_f32, c_ptr],
    'pg_instance_norm_act': [c_ptr, c_ptr, c_i64, c_i64, c_f32, c_i32, c_f32, c_f32, c_ptr],
    'pg_conv2d_igemm_spade_run': [c_ptr] * 6 + [c_i32] * 6 + [c_i32, c_f32, c_f32, c_i32, c_i32, c_i32, c_ptr],
    'pg_conv2d_igemm_fwd': [c_ptr] * 6 + [c_i64, c_ptr, c_ptr] + [c_i32] * 7 + [c_i32, c_i32, c_f32, c_f32, c_i32, c_f32, c_f32, c_f32, c_i32,
                            c_ptr, c_i64, c_ptr],
    'pg_torgb_skip': [c_ptr] * 7 + [c_i32] * 5 + [c_f32, c_ptr],
    'pg_torgb_skip_c8': [c_ptr] * 7 + [c_i32] * 5 + [c_f32, c_ptr],
    'pg_conv2d_igemm_prepack_batched': [c_ptr, c_i64, c_ptr, c_i32, c_ptr, c_f32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_ptr, c_i64, c_ptr],
    'pg_conv2d_igemm_launch': [c_ptr],
    'pg_conv2d_wgrad_workspace_bytes': [c_i32] * 6,
    'pg_conv2d_wgrad': [c_ptr, c_ptr, c_ptr] + [c_i32] * 6 + [c_f32, c_i32, c_ptr, c_i64, c_ptr],
    'pg_set_tuning': [ctypes.c_char_p, c_i32],
    'pg_masked_fill_c8': [c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_i64, c_i64, c_i64, c_ptr],
    'pg_nchw_to_c8': [c_ptr, c_ptr, c_i64, c_i64, c_i64, c_i32, c_ptr],
    'pg_c8_to_nchw': [c_ptr, c_ptr, c_i64, c_i64, c_i64, c_i32, c_ptr],
    'pg_upfirdn2d_bias_act': [c_ptr, c_ptr, c_ptr, c_ptr, I32x4, I64x4, I32x4, I64x4, c_i32, c_i32, c_i64, c_i64] + [c_i32] * 8 +
                             [c_i32, c_f32, c_i32, c_f32, c_f32, c_f32, c_i32, c_ptr],
    'pg_warp_perspective_u8': [c_ptr, c_i32, c_i32, c_ptr],
    'pg_patch_denorm_u8': [c_ptr] * 6 + [c_i32] * 6 + [c_ptr],
    'pg_patch_crop_transforms': [c_ptr, c_i32, c_i32, c_i32, c_i32, ctypes.c_double] + [c_ptr] * 5,
}


class WarpJob(ctypes.Structure):
    """pg_warp_job of include/pasta_b200.h (128 bytes; one cv2.warpPerspective call)."""
    _fields_ = [
        ('m', ctypes.c_double * 9),
        ('src', c_ptr), ('dst', c_ptr),
        ('src_h', c_i32), ('src_w', c_i32), ('src_row_stride', c_i32), ('src_pix_stride', c_i32),
        ('dst_h', c_i32), ('dst_w', c_i32), ('dst_row_stride', c_i32), ('dst_pix_stride', c_i32),
        ('channels', c_i32), ('border', c_i32),
    ]


class ConvArgs(ctypes.Structure):
    """pg_conv_args of include/pasta_b200.h (field order and types must match the header)."""
    _fields_ = [
        ('struct_bytes', ctypes.c_uint32),
        ('N', c_i32), ('Cin', c_i32), ('Cout', c_i32), ('H', c_i32), ('W', c_i32), ('ksize', c_i32), ('up', c_i32),
        ('x', c_ptr), ('x_dtype', c_i32), ('x_layout', c_i32),
        ('x2', c_ptr), ('cin1', c_i32), ('residual_layout', c_i32),
        ('wpack', c_ptr), ('wpack_sample_stride', c_i64),
        ('styles', c_ptr), ('dcoefs', c_ptr), ('noise', c_ptr), ('noise_batch_stride', c_i64), ('bias', c_ptr), ('residual', c_ptr),
        ('y', c_ptr), ('y_dtype', c_i32), ('y_layout', c_i32),
        ('in_act', c_i32), ('in_alpha', c_f32), ('in_gain', c_f32),
        ('act', c_i32), ('alpha', c_f32), ('gain', c_f32), ('clamp', c_f32),
        ('operand_format', c_i32), ('n_tile', c_i32),
        ('spade_x', c_ptr), ('spade_mean', c_ptr), ('spade_rstd', c_ptr),
        ('stream', c_ptr),
    ]


LAYOUT_NCHW, LAYOUT_C8 = 0, 1


class PastaB200Error(RuntimeError):
    """Raised for every non-zero return of the C ABI (the reference raises RuntimeError via TORCH_CHECK)."""


def lib_path():
    return _LIB_PATH


def load():
    """dlopen libpasta_b200.so once per process; raise loudly if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(_LIB_PATH):
            raise PastaB200Error(
                f'{_LIB_PATH} is missing: build it with `python pasta-gan_b200/build.py` (or __graft_entry__.build()). '
                'pasta-b200 has no CPU or PyTorch fallback for the operator hot path.')
        lib = ctypes.CDLL(_LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)                     # AttributeError if the symbol is not exported
            fn.argtypes = argtypes
            fn.restype = {'pg_last_error': ctypes.c_char_p, 'pg_launch_count': c_i64, 'pg_conv2d_igemm_workspace_bytes': c_i64, 'pg_conv2d_igemm_workspace_bytes_fmt': c_i64, 'pg_conv2d_wgrad_workspace_bytes': c_i64}.get(name, c_int)
        if lib.pg_abi_version() != 2:
            raise PastaB200Error(f'ABI mismatch: library reports {lib.pg_abi_version()}, binding expects 2')
        if hasattr(lib, 'pg_debug_set_buffer'):                      # -DPG_DEBUG builds only (tools/conv_timeline.py)
            lib.pg_debug_set_buffer.argtypes, lib.pg_debug_set_buffer.restype = [c_ptr], None
        _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        msg = load().pg_last_error()
        raise PastaB200Error(f'{what}: {msg.decode() if msg else "error"} (code {rc})')


def set_tuning(key, value):
    """pg_set_tuning: override one of the conv plan / loader choices for the rest of the process (tests and tools compare variants)."""
    check(load().pg_set_tuning(key.encode(), int(value)), 'pg_set_tuning')


def require_device():
    """Call once before the first launch: the current CUDA device must be able to run the sm_100a images."""
    global _device_checked
    if not _device_checked:
        check(load().pg_check_device(), 'pg_check_device')
        _device_checked = True


DTYPE_CODES = {'torch.float32': 0, 'torch.float16': 1, 'torch.float64': 2}


def dtype_code(dtype):
    try:
        return DTYPE_CODES[str(dtype)]
    except KeyError:
        raise PastaB200Error(f'unsupported dtype {dtype}: the operator path handles float32, float16 and float64') from None


def ptr(t):
    """Device pointer of a tensor, or NULL for None / empty (the reference passes empty tensors for 'absent')."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def launch_count():
    """Kernels launched by libpasta_b200.so in this process so far."""
    return int(load().pg_launch_count())


def current_stream(device):
    import torch
    return torch.cuda.current_stream(device).cuda_stream


# ---------------------------------------------------------------------------------------------------------------------
# Optional per-launch timing (bench.py's roofline leg): CUDA events on the launching stream around each C-ABI call.
# Disabled (zero work) unless a LaunchProfiler is installed; never active under CUDA-graph capture.

_profiler = None


class LaunchProfiler:
    """Collects (kernel name, algorithmic bytes, algorithmic flops, start event, end event) per launch."""

    def __init__(self):
        self.records = []

    def __enter__(self):
        global _profiler
        _profiler = self
        return self

    def __exit__(self, *exc):
        global _profiler
        _profiler = None

    def summary(self):
        import torch
        torch.cuda.synchronize()
        agg = {}
        for name, nbytes, flops, e0, e1, _tag in self.records:
            a = agg.setdefault(name, dict(launches=0, ms=0.0, bytes=0, flops=0))
            a['launches'] += 1
            a['ms'] += e0.elapsed_time(e1)
            a['bytes'] += nbytes
            a['flops'] += flops
        return agg


class _Span:
    __slots__ = ('name', 'nbytes', 'flops', 'e0', 'tag')

    def __init__(self, name, nbytes, flops, tag=''):
        import torch
        self.name, self.nbytes, self.flops, self.tag = name, nbytes, flops, tag
        self.e0 = torch.cuda.Event(enable_timing=True)
        self.e0.record()

    def close(self):
        import torch
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        if _profiler is not None:
            _profiler.records.append((self.name, self.nbytes, self.flops, self.e0, e1, self.tag))


def span(name, nbytes=0, flops=0, tag=''):
    """Open a timing span if a profiler is installed; returns None otherwise (callers do `if s: s.close()`)."""
    if _profiler is None:
        return None
    return _Span(name, nbytes, flops, tag)
