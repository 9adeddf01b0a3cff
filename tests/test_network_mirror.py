"""The host-side mirror of the generator (pasta-gan_b200/networks.py) evaluated by the ORACLE operator table on CPU must
reproduce the unmodified reference GeneratorFull: same parameter names/shapes (fingerprints of procedurally-filled
weights agree) and same outputs on the synthetic try-on inputs.  Golden outputs are stored in fp16 (5e-4 relative),
so the network-level tolerance here is 2e-3; per-layer fixtures are fp32 and held to 1e-4."""
import numpy as np
import pytest
import torch

import procedural
from conftest import rel_err
from oracle import ops_oracle as O
from pasta_gan_b200 import networks as N


@pytest.fixture(scope='module')
def oracle_generator():
    G = N.build_generator_full().eval()
    procedural.fill_(G)
    return N.use_ops(G, O.operator_table())


def test_parameter_names_and_values_match_reference(golden, oracle_generator):
    g = golden('generator_full')
    meta = g.meta[0]
    fp = procedural.fingerprint(oracle_generator)
    assert sorted(fp) == meta['names'], set(fp) ^ set(meta['names'])
    assert sum(p.numel() for p in oracle_generator.parameters()) == meta['n_params']
    assert oracle_generator.num_ws == meta['num_ws']
    s = np.array([fp[n][0] for n in meta['names']])
    a = np.array([fp[n][1] for n in meta['names']])
    np.testing.assert_allclose(s, g.arrays['fp_sum'], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(a, g.arrays['fp_abs'], rtol=1e-9, atol=1e-9)


def test_generator_forward_matches_reference(golden, oracle_generator):
    g = golden('generator_full')
    G = oracle_generator
    inp = procedural.synth_inputs(2)
    with torch.no_grad():
        pose_feat = G.const_encoding(inp['pose'])
        stylecode, feats = G.style_encoding(inp['c'], inp['retain'])
        ws = G.mapping(inp['z'], stylecode)
        assert rel_err(pose_feat, g.t('pose_feat')) < 1e-4
        assert rel_err(stylecode, g.t('stylecode')) < 1e-4
        assert rel_err(ws[:, 0], g.t('ws0')) < 1e-4
        assert rel_err(feats[2][:, :, ::4, ::4], g.t('feat64')) < 1e-4
        assert rel_err(feats[0][:, ::8, ::16, ::16], g.t('feat256')) < 1e-4
        img, fimg, parsing = G(**inp, noise_mode='const')
    assert rel_err(img, g.t('img', dtype=torch.float32)) < 2e-3
    assert rel_err(fimg, g.t('finetune_img', dtype=torch.float32)) < 2e-3
    assert rel_err(parsing, g.t('pred_parsing', dtype=torch.float32)) < 2e-3


LAYER_BUILDERS = {
    'synth_s1': lambda: N.SynthesisLayer(8, 6, w_dim=512, resolution=16, conv_clamp=256),
    'synth_up': lambda: N.SynthesisLayer(8, 6, w_dim=512, resolution=16, up=2, conv_clamp=256),
    'torgb': lambda: N.ToRGBLayerFull(8, 3, w_dim=512, conv_clamp=256, is_last=True, is_style=True),
    'conv_plain': lambda: N.Conv2dLayer(6, 7, kernel_size=3, activation='lrelu', conv_clamp=256),
    'conv_down': lambda: N.Conv2dLayer(6, 7, kernel_size=3, down=2),
    'conv_up': lambda: N.Conv2dLayer(6, 7, kernel_size=1, bias=False, up=2),
    'conv_7x7': lambda: N.Conv2dLayer(3, 8, kernel_size=7, activation='relu'),
    'resblock_down': lambda: N.ResBlock(6, 8, kernel_size=4, activation='relu', down=2),
    'fc_lrelu': lambda: N.FullyConnectedLayer(12, 9, activation='lrelu', lr_multiplier=0.01),
    'fc_linear': lambda: N.FullyConnectedLayer(12, 9, bias_init=1),
    'dense': lambda: N.Dense(8, 8),
    'spade_norm': lambda: N.SpadeNormBlock(10, 6),
}
LAYER_KWARGS = {
    'synth_s1_eval': dict(noise_mode='const', fused_modconv=True), 'synth_s1_train': dict(noise_mode='const', fused_modconv=False),
    'synth_up_eval': dict(noise_mode='const', fused_modconv=True, gain=0.5 ** 0.5),
    'synth_up_train': dict(noise_mode='const', fused_modconv=False, gain=0.5 ** 0.5),
    'torgb_eval': dict(fused_modconv=True), 'torgb_train': dict(fused_modconv=False), 'conv_up': dict(gain=0.5 ** 0.5),
}


def run_layer_case(golden, tag, table, device='cpu', tol=1e-4):
    g = golden('layers')
    base = tag.replace('_eval', '').replace('_train', '')
    mod = LAYER_BUILDERS[base]()
    procedural.fill_(mod)
    mod.train(tag.endswith('_train'))
    N.use_ops(mod, table)
    mod.to(device)
    args = [g.t(f'{tag}/in{j}', device=device) for j in range(4) if g.has(f'{tag}/in{j}')]
    with torch.no_grad():
        out = mod(*args, **LAYER_KWARGS.get(tag, {}))
    out = out if isinstance(out, (tuple, list)) else [out]
    for j, o in enumerate(out):
        if o is not None:
            assert rel_err(o, g.t(f'{tag}/out{j}')) < tol, (tag, j)


@pytest.mark.parametrize('tag', ['synth_s1_eval', 'synth_s1_train', 'synth_up_eval', 'synth_up_train', 'torgb_eval', 'torgb_train',
                                 'conv_plain', 'conv_down', 'conv_up', 'conv_7x7', 'resblock_down', 'fc_lrelu', 'fc_linear', 'dense', 'spade_norm'])
def test_layers_oracle_vs_reference(golden, tag):
    run_layer_case(golden, tag, O.operator_table())


def test_generator_512_mirror_matches_reference(golden):
    """BASELINE configs[2] network (reference Generator_512, the only 512-px generator in the tree) on CPU through the oracle table."""
    g = golden('generator_512')
    meta = g.meta[0]
    G = N.build_generator_512().eval()
    procedural.fill_(G)
    fp = procedural.fingerprint(G)
    assert sorted(fp) == meta['names'] and G.num_ws == meta['num_ws']
    np.testing.assert_allclose(np.array([fp[n][0] for n in meta['names']]), g.arrays['fp_sum'], rtol=1e-9, atol=1e-9)
    N.use_ops(G, O.operator_table(fast=True))
    with torch.no_grad():
        img = G(**procedural.synth_inputs_512(1), noise_mode='const')
    assert rel_err(img, g.t('img', dtype=torch.float32)) < 2e-3
