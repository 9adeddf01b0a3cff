"""Locates the C-ABI binding (``_capi.py`` next to ``csrc/``) no matter how this ``ops`` directory was
imported: as ``pasta_gan_b200.torch_utils.ops``, as a top-level ``torch_utils.ops`` overlay, or copied
into another tree (then set PASTA_B200_HOME to the ``pasta-gan_b200`` directory).  The binding is loaded
once per process under the canonical module name ``pasta_b200_capi``."""
import importlib.util
import os
import sys

_NAME = 'pasta_b200_capi'


def capi():
    mod = sys.modules.get(_NAME)
    if mod is not None:
        return mod
    home = os.environ.get('PASTA_B200_HOME') or os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    path = os.path.join(home, '_capi.py')
    if not os.path.exists(path):
        raise ImportError(f'pasta-b200: cannot find {path}; set PASTA_B200_HOME to the pasta-gan_b200 directory. '
                          'There is no fallback implementation of the operator hot path.')
    spec = importlib.util.spec_from_file_location(_NAME, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod


def require_cuda(t, opname):
    """The product path is CUDA-only: refuse CPU tensors instead of silently computing them elsewhere."""
    if t.device.type != 'cuda':
        raise RuntimeError(f'{opname}: got a {t.device.type} tensor. pasta-b200 runs this operator on sm_100a only and ships '
                           "no CPU / impl='ref' path (the CPU restatement lives in oracle/ as test infrastructure).")


def refuse_ref(impl, opname):
    assert impl in ['ref', 'cuda']
    if impl == 'ref':
        raise RuntimeError(f"{opname}: impl='ref' is not shipped by pasta-b200; use oracle/ops_oracle.py in tests.")
