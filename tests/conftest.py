"""pytest configuration: registers the ``gpu`` marker and puts the repo root on sys.path.

``-m "not gpu"`` : oracle vs golden vectors, host logic, C-ABI symbol checks, gloo world_size-2 (no GPU needed).
``-m gpu``       : parity of the CUDA path (through the C-ABI) against the oracle and the golden vectors.
"""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests', 'golden')):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


class Golden:
    """A committed fixture written by tests/golden/gen_golden.py."""

    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, name + '.npz'))
        self.arrays = {k: z[k] for k in z.files if k != 'meta'}
        self.meta = json.loads(bytes(z['meta']).decode())

    def t(self, key, dtype=None, device='cpu'):
        import torch
        a = torch.from_numpy(self.arrays[key].copy())
        if dtype is not None:
            a = a.to(dtype)
        return a.to(device)

    def has(self, key):
        return key in self.arrays


@pytest.fixture(scope='session')
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = Golden(name)
        return cache[name]
    return get


def rel_err(a, b):
    """max |a-b| / max(|b|) — the relative-error measure every tolerance in tests/ refers to."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    if b.numel() == 0:
        return 0.0
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
