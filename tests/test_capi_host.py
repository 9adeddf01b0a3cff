"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/pasta_b200.h declares, the binding lists each of them, and the product path refuses to run
without a GPU (no CPU fallback, no impl='ref').  No kernel is launched here."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT, rel_err

import pasta_gan_b200
from pasta_gan_b200.torch_utils.ops import upfirdn2d, bias_act, conv2d_resample, conv2d_gradfix, fma, grid_sample_gradfix


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'pasta_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(pg_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    capi = pasta_gan_b200.capi
    lib = ctypes.CDLL(capi.lib_path())
    names = header_symbols()
    assert len(names) >= 6
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/pasta_b200.h but not exported'
        assert n in capi.SIGNATURES, f'{n} has no ctypes signature in _capi.py'
    assert sorted(capi.SIGNATURES) == names
    assert capi.load().pg_abi_version() == 2


def test_argument_errors_are_reported_not_crashed():
    """Validation happens before any launch, so it can be exercised without a device."""
    capi = pasta_gan_b200.capi
    lib = capi.load()
    rc = lib.pg_bias_act(None, None, None, None, None, None, 2 ** 31, 0, 1, 0, 3, 0.2, 1.0, -1.0, 0, None)
    assert rc == 1 and b'too large' in lib.pg_last_error()
    rc = lib.pg_bias_act(None, None, None, None, None, None, 16, 0, 1, 5, 3, 0.2, 1.0, -1.0, 0, None)
    assert rc == 1 and b'grad' in lib.pg_last_error()
    sz, st = capi.I32x4(1, 1, 4, 4), capi.I64x4(16, 16, 4, 1)
    rc = lib.pg_upfirdn2d(None, None, None, sz, st, sz, st, 4, 4, 4, 1, 0, 1, 1, 1, 0, 0, 0, 0, 0, 1.0, 0, None)
    assert rc == 1 and b'upsampling factor' in lib.pg_last_error()
    rc = lib.pg_upfirdn2d(None, None, None, sz, st, sz, st, 4, 4, 4, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 1.0, 0, None)
    assert rc == 1 and b'out_size' in lib.pg_last_error()            # 4x4 filter, no padding -> 1x1 output
    with pytest.raises(capi.PastaB200Error):
        capi.check(rc, 'pg_upfirdn2d')


def test_conv_args_struct_matches_header():
    """pg_conv2d_igemm_launch checks pg_conv_args.struct_bytes against its own sizeof before touching the device: an N = 0 call with the binding's
    ctypes Structure must pass that check (rc 0), a wrong size must be refused; pg_set_tuning knows its keys."""
    import ctypes
    capi = pasta_gan_b200.capi
    lib = capi.load()
    a = capi.ConvArgs()
    a.struct_bytes = ctypes.sizeof(capi.ConvArgs)
    a.N, a.Cin, a.Cout, a.H, a.W, a.ksize, a.up = 0, 16, 16, 8, 8, 3, 1
    a.in_act, a.act, a.gain, a.in_gain, a.clamp = 1, 1, 1.0, 1.0, -1.0
    assert lib.pg_conv2d_igemm_launch(ctypes.byref(a)) == 0, lib.pg_last_error()
    a.struct_bytes -= 8
    assert lib.pg_conv2d_igemm_launch(ctypes.byref(a)) == 1 and b'struct_bytes' in lib.pg_last_error()
    assert lib.pg_set_tuning(b'conv_bands', 1) == 0
    assert lib.pg_set_tuning(b'no_such_key', 1) == 1 and b'unknown key' in lib.pg_last_error()
    assert not hasattr(ctypes.CDLL(capi.lib_path()), 'pg_debug_set_buffer'), 'the release library must not export the debug hook'


def test_no_cpu_fallback():
    x = torch.randn(1, 2, 8, 8)
    f = upfirdn2d.setup_filter([1, 3, 3, 1])
    with pytest.raises(RuntimeError, match='sm_100a only'):
        upfirdn2d.upfirdn2d(x, f)
    with pytest.raises(RuntimeError, match='sm_100a only'):
        bias_act.bias_act(x, act='lrelu')
    with pytest.raises(RuntimeError, match="impl='ref'"):
        bias_act.bias_act(x, act='lrelu', impl='ref')
    with pytest.raises(RuntimeError, match='sm_100a only'):
        conv2d_resample.conv2d_resample(x, torch.randn(2, 2, 3, 3), f=f, up=2, padding=1)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'pasta-gan_b200')
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M), f'{fn} imports oracle/'


def test_setup_filter_matches_reference(golden):
    g = golden('upfirdn2d')
    assert rel_err(upfirdn2d.setup_filter([1, 3, 3, 1]), g.t('sf/1331')) < 1e-7
    assert rel_err(upfirdn2d.setup_filter([1, 2, 3, 1], gain=4, flip_filter=True), g.t('sf/1331_g4_flip')) < 1e-7
    assert rel_err(upfirdn2d.setup_filter([1, 2, 3, 4, 4, 3, 2, 1], gain=4), g.t('sf/sep8')) < 1e-7
    assert rel_err(upfirdn2d.setup_filter(None), g.t('sf/none')) < 1e-7
    assert rel_err(upfirdn2d.setup_filter([[1, 2], [3, 4]], normalize=False), g.t('sf/nonorm2d')) < 1e-7


def test_call_surface():
    """Names, defaults and helper re-exports the reference's callers rely on."""
    import inspect
    sig = lambda fn: list(inspect.signature(fn).parameters)
    assert sig(upfirdn2d.upfirdn2d) == ['x', 'f', 'up', 'down', 'padding', 'flip_filter', 'gain', 'impl']
    assert sig(upfirdn2d.filter2d) == ['x', 'f', 'padding', 'flip_filter', 'gain', 'impl']
    assert sig(upfirdn2d.upsample2d) == ['x', 'f', 'up', 'padding', 'flip_filter', 'gain', 'impl']
    assert sig(upfirdn2d.downsample2d) == ['x', 'f', 'down', 'padding', 'flip_filter', 'gain', 'impl']
    assert sig(upfirdn2d.setup_filter) == ['f', 'device', 'normalize', 'flip_filter', 'gain', 'separable']
    assert sig(bias_act.bias_act) == ['x', 'b', 'dim', 'act', 'alpha', 'gain', 'clamp', 'impl']
    assert sig(conv2d_resample.conv2d_resample) == ['x', 'w', 'f', 'up', 'down', 'padding', 'groups', 'flip_weight', 'flip_filter']
    assert sig(conv2d_gradfix.conv2d) == ['input', 'weight', 'bias', 'stride', 'padding', 'dilation', 'groups']
    assert sig(conv2d_gradfix.conv_transpose2d) == ['input', 'weight', 'bias', 'stride', 'padding', 'output_padding', 'groups', 'dilation']
    assert sig(fma.fma) == ['a', 'b', 'c']
    assert sig(grid_sample_gradfix.grid_sample) == ['input', 'grid']
    assert conv2d_gradfix.enabled is False and conv2d_gradfix.weight_gradients_disabled is False
    with conv2d_gradfix.no_weight_gradients():
        assert conv2d_gradfix.weight_gradients_disabled is True
    assert conv2d_gradfix.weight_gradients_disabled is False
    spec = bias_act.activation_funcs
    assert list(spec) == ['linear', 'relu', 'lrelu', 'tanh', 'sigmoid', 'elu', 'selu', 'softplus', 'swish']
    assert [spec[k].cuda_idx for k in spec] == list(range(1, 10))
    assert abs(spec['lrelu'].def_gain - 2 ** 0.5) < 1e-12 and spec['lrelu'].def_alpha == 0.2
    assert spec['swish'].ref == 'x' and spec['lrelu'].has_2nd_grad is False and spec['tanh'].has_2nd_grad is True
    assert upfirdn2d._parse_padding(3) == (3, 3, 3, 3) and upfirdn2d._parse_padding([1, 2]) == (1, 1, 2, 2)
    assert upfirdn2d._parse_scaling(2) == (2, 2) and upfirdn2d._get_filter_size(None) == (1, 1)
    assert conv2d_resample._parse_padding is upfirdn2d._parse_padding


def test_fma_cpu_matches_autograd():
    """fma is plain torch (elementwise), so its broadcast-aware gradients can be checked on CPU."""
    torch.manual_seed(0)
    a = torch.randn(2, 3, 4, 4, dtype=torch.float64, requires_grad=True)
    b = torch.randn(2, 3, 1, 1, dtype=torch.float64, requires_grad=True)
    c = torch.randn(4, 4, dtype=torch.float64, requires_grad=True)
    assert torch.autograd.gradcheck(fma.fma, (a, b, c))
    assert torch.allclose(fma.fma(a, b, c), a * b + c)
