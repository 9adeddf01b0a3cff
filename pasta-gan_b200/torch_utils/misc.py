"""The three helpers of the reference's torch_utils/misc.py that the operator layer itself calls
(assert_shape :86-99, profiled_function :104-109, suppress_tracer_warnings :72-80), written fresh."""
import contextlib
import functools
import warnings

import torch


@contextlib.contextmanager
def suppress_tracer_warnings():
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', category=torch.jit.TracerWarning)
        yield


def assert_shape(tensor, ref_shape):
    """ref_shape entries: int = must match, None = any size."""
    if tensor.ndim != len(ref_shape):
        raise AssertionError(f'Wrong number of dimensions: got {tensor.ndim}, expected {len(ref_shape)}')
    for idx, (size, ref) in enumerate(zip(tensor.shape, ref_shape)):
        if ref is not None and int(size) != int(ref):
            raise AssertionError(f'Wrong size for dimension {idx}: got {int(size)}, expected {int(ref)}')


def profiled_function(fn):
    """Same profiler range names as the reference, so traces line up."""
    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        with torch.autograd.profiler.record_function(fn.__name__):
            return fn(*args, **kwargs)
    return wrapped
