"""Import shim: the product directory is ``pasta-gan_b200/`` (not a valid Python identifier), so
``import pasta_gan_b200`` is mapped onto it here."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), 'pasta-gan_b200')
__path__ = [_real]
with open(_os.path.join(_real, '__init__.py')) as _fh:
    exec(compile(_fh.read(), _os.path.join(_real, '__init__.py'), 'exec'))
del _fh
