"""The real drop-in on the GPU: the UNMODIFIED reference ``training/networks.py`` (``GeneratorFull`` :5844 and the four sub-module calls of
test.py:121-128) running over OUR ``torch_utils/ops`` — the file overlay of INTEGRATION.md §2 — against the golden outputs of the same reference
on CPU, and the ``persistence.import_hook`` recipe after a ``legacy.load_network_pkl`` round trip.

The reference tree is ``baseline/_ref`` (a plain copy made by ``__graft_entry__.build()`` in the authoring container; it ships with the gpurun
snapshot, ``/root/reference`` does not exist on the GPU box).  Each arrangement runs in its own process (``baseline/run_reference.py``), because
the overlay must own the top-level name ``torch_utils``."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
HARNESS = os.path.join(ROOT, 'baseline', 'run_reference.py')
HAVE_REF = os.path.isdir(os.path.join(ROOT, 'baseline', '_ref', 'torch_utils'))


def run(mode, *extra):
    r = subprocess.run([sys.executable, HARNESS, '--mode', mode, *extra], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith('{')][-1]
    return json.loads(line)


@pytest.mark.skipif(not HAVE_REF, reason='baseline/_ref not present (run __graft_entry__.build() where the reference checkout exists)')
@pytest.mark.parametrize('mode', ['overlay', 'overlay_hook'])
def test_unmodified_reference_generator_over_our_ops(mode):
    out = run(mode, '--batch', '2', '--steps', '1', '--warmup', '1', '--check')
    print(out)
    assert out['overlay_in_effect'], 'the reference ops were imported instead of the overlay'
    assert out['our_kernel_launches'] > 100, 'the forward did not run on libpasta_b200.so'
    g = out['golden']
    # coarse image and parsing logits: max-abs relative within the north_star's 1e-2; the fine image depends on argmax(parsing) (networks.py:5823-5826)
    # and is held in relative L2, as in test_gpu_network.py
    assert g['img']['max_rel'] < 1e-2 and g['pred_parsing']['max_rel'] < 1e-2, g
    assert all(v['l2_rel'] < 1e-2 for v in g.values()), g
    if mode == 'overlay_hook':
        assert out['hooked'], 'the import hook did not pull in pasta_gan_b200.networks.modulated_conv2d'
