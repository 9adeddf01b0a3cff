"""Shim: the name-keyed procedural weights / synthetic inputs live in the product package (``pasta-gan_b200/synthetic.py``, torch-only) because the
training step and bench use them too; tests and ``gen_golden.py`` keep importing ``procedural``.  Loaded by file path so that ``gen_golden.py`` can
use it next to the REFERENCE's ``torch_utils`` without importing our package."""
import importlib.util as _u
import os as _os

_path = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))), 'pasta-gan_b200', 'synthetic.py')
_spec = _u.spec_from_file_location('pasta_b200_synthetic', _path)
_mod = _u.module_from_spec(_spec)
_spec.loader.exec_module(_mod)
globals().update({k: v for k, v in vars(_mod).items() if not k.startswith('__')})
