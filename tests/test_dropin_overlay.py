"""Drop-in check of INTEGRATION.md §2 (CPU, needs the reference checkout, skipped where it is absent): a copy of the reference's Python
tree with OUR torch_utils/ops/*.py laid over its own imports cleanly, builds GeneratorFull / Discriminator (constructors call
upfirdn2d.setup_filter and read bias_act.activation_funcs), round-trips through the reference's persistence pickling, and the overlaid
ops refuse CPU tensors (i.e. they really are ours, and there is no hidden fallback).  The overlaid tree lives in a temp dir; nothing
is copied into the repository."""
import os
import shutil
import subprocess
import sys
import textwrap

import pytest

from conftest import ROOT

REF = os.environ.get('PASTA_REFERENCE', '/root/reference')


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'torch_utils')), reason='reference checkout not available')
def test_overlay_imports_constructs_and_pickles(tmp_path):
    tree = tmp_path / 'ref'
    shutil.copytree(REF, tree, ignore=shutil.ignore_patterns('__pycache__', '*.pyc', '.git'))
    ops_src = os.path.join(ROOT, 'pasta-gan_b200', 'torch_utils', 'ops')
    for fn in os.listdir(ops_src):
        if fn.endswith('.py') and fn != '__init__.py':
            shutil.copy(os.path.join(ops_src, fn), tree / 'torch_utils' / 'ops' / fn)
    script = textwrap.dedent('''
        import io, os, pickle, sys, types
        for m in ('matplotlib', 'matplotlib.pyplot'):
            sys.modules.setdefault(m, types.ModuleType(m))
        sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
        import torch
        torch.version.cuda = '11.0'
        from torch_utils.ops import upfirdn2d, bias_act, conv2d_resample, conv2d_gradfix, fma
        assert 'pg_upfirdn2d' in open(upfirdn2d.__file__).read(), 'overlay not in effect'
        import training.networks as N
        import legacy
        G = N.GeneratorFull(z_dim=0, c_dim=512, w_dim=512, img_resolution=256, img_channels=3, mapping_kwargs=dict(num_layers=1),
                            synthesis_kwargs=dict(channel_base=1024, channel_max=32, num_fp16_res=3, conv_clamp=256, use_noise=True))
        D = N.Discriminator(c_dim=512, img_resolution=256, img_channels=3, channel_base=1024, channel_max=32, num_fp16_res=3, conv_clamp=256)
        buf = io.BytesIO(); pickle.dump(dict(G_ema=G, G=G, D=D), buf); buf.seek(0)
        G2 = legacy.load_network_pkl(buf)['G_ema']
        assert sorted(dict(G2.named_parameters())) == sorted(dict(G.named_parameters()))
        try:
            G2.synthesis.b4.conv1(torch.zeros(1, 32, 4, 4), torch.zeros(1, 512))
        except RuntimeError as e:
            assert 'sm_100a only' in str(e), e
            print('OVERLAY-OK')
        else:
            raise SystemExit('a CPU tensor was accepted: the overlay has a fallback')
    ''')
    env = dict(os.environ, PYTHONPATH=str(tree), PASTA_B200_HOME=os.path.join(ROOT, 'pasta-gan_b200'), PYTHONDONTWRITEBYTECODE='1')
    r = subprocess.run([sys.executable, '-c', script], cwd=tree, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and 'OVERLAY-OK' in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
