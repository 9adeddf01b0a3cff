"""pasta-b200: the PASTA-GAN generator/discriminator operator hot path, hand-written for B200 (sm_100a).

    from pasta_gan_b200 import ops            # upfirdn2d, bias_act, conv2d_resample, conv2d_gradfix, fma
    from pasta_gan_b200 import networks       # host-side mirror of the modules that call the ops

Everything numerical goes through the C ABI in include/pasta_b200.h (libpasta_b200.so, built by build.py);
there is no CPU fallback and no alternative backend."""
from .torch_utils.ops import _backend as _backend

capi = _backend.capi()


def build_library(force=False, verbose=False):
    """Compile the CUDA sources in csrc/ for sm_100a (in-tree) and return the path of libpasta_b200.so."""
    import importlib
    return importlib.import_module(__name__ + '.build').build(force=force, verbose=verbose)


def __getattr__(name):
    import importlib
    if name == 'ops':
        return importlib.import_module(__name__ + '.torch_utils.ops')
    if name in ('networks', 'torch_utils', 'data_parallel'):
        return importlib.import_module(__name__ + '.' + name)
    raise AttributeError(name)
