#!/usr/bin/env python
"""Generate the golden vectors in this directory from the UNMODIFIED reference.

Run (in the authoring container, where /root/reference is mounted read-only):

    python tests/golden/gen_golden.py [case ...]  # writes tests/golden/*.npz (into $PASTA_GOLDEN_OUT if set); the patch-routing vectors: gen_warp_golden.py

The reference is imported as-is from /root/reference on CPU, so every op takes its
``impl='ref'`` branch (torch_utils/ops/upfirdn2d.py:162-164, bias_act.py:87-89).
Harness shims only (the reference tree is untouched): stub ``matplotlib``, chdir to
the reference root (``util_functions.py:11`` opens ./human_colormap.mat), pretend
``torch.version.cuda == '11.0'`` while importing ``training.networks``
(networks.py:1206-1222), PYTHONDONTWRITEBYTECODE.

This script is the only thing that reads /root/reference; tests, smoke() and
bench.py read the committed .npz fixtures instead.
"""

import json
import os
import sys
import types

os.environ.setdefault('PYTHONDONTWRITEBYTECODE', '1')
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('PASTA_REFERENCE', '/root/reference')
sys.path.insert(0, REF)
sys.path.insert(0, HERE)

import numpy as np
import torch

for _m in ('matplotlib', 'matplotlib.pyplot'):
    sys.modules.setdefault(_m, types.ModuleType(_m))
sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
os.chdir(REF)
_cuda = torch.version.cuda
torch.version.cuda = '11.0'
from torch_utils.ops import upfirdn2d as R_up          # noqa: E402
from torch_utils.ops import bias_act as R_ba           # noqa: E402
from torch_utils.ops import conv2d_resample as R_cr    # noqa: E402
import training.networks as R_net                      # noqa: E402
torch.version.cuda = _cuda
os.chdir(HERE)

import procedural                                      # noqa: E402

torch.set_num_threads(max(1, os.cpu_count() or 1))


def rnd(seed, *shape, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=dtype)


def save(name, arrays, meta):
    arrays = {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in arrays.items()}
    arrays['meta'] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(os.environ.get('PASTA_GOLDEN_OUT', HERE), name + '.npz')      # PASTA_GOLDEN_OUT=/tmp/x: regenerate beside the committed files to compare
    np.savez_compressed(path, **arrays)
    print(f'{name}.npz  {os.path.getsize(path) / 1e3:.1f} kB  ({len(meta)} cases)')


# ----------------------------------------------------------------------------- upfirdn2d

def gen_upfirdn2d():
    f4 = R_up.setup_filter([1, 3, 3, 1])
    f12 = torch.tensor([0.0154, -0.0349, -0.1180, 0.0483, 0.4911, 0.7873, 0.3379, -0.0726, -0.0211, 0.0447, 0.0018, -0.0078])
    f12 = R_up.setup_filter(f12.tolist())                       # 1-D separable (>= 8 taps), like augment.py's sym6
    f35 = rnd(7, 3, 5)
    cases = [
        # the closed family of SURVEY.md appendix A (forward forms and their backward forms)
        dict(shape=[2, 3, 17, 17], f='f4', up=1, down=1, padding=[1, 1, 1, 1], gain=4, flip=False),
        dict(shape=[2, 3, 16, 16], f='f4', up=1, down=1, padding=[2, 2, 2, 2], gain=1, flip=False),
        dict(shape=[2, 3, 16, 16], f='f4', up=1, down=2, padding=[1, 1, 1, 1], gain=1, flip=False),
        dict(shape=[2, 3, 8, 8], f='f4', up=2, down=1, padding=[2, 1, 2, 1], gain=4, flip=False),
        dict(shape=[1, 5, 33, 33], f='f4', up=1, down=1, padding=[1, 1, 1, 1], gain=4, flip=True),
        dict(shape=[1, 2, 65, 65], f='f4', up=1, down=1, padding=[2, 2, 2, 2], gain=1, flip=False),
        dict(shape=[1, 4, 64, 64], f='f4', up=1, down=2, padding=[1, 1, 1, 1], gain=4, flip=False),
        dict(shape=[3, 3, 32, 32], f='f4', up=2, down=1, padding=[2, 1, 2, 1], gain=1, flip=False),
        # generic: separable 12-tap, flips, negative padding, anisotropic factors, non-square
        dict(shape=[2, 2, 20, 24], f='f12', up=1, down=1, padding=[5, 6, 5, 6], gain=1, flip=True),
        dict(shape=[1, 3, 20, 24], f='f12', up=2, down=1, padding=[6, 5, 6, 5], gain=4, flip=False),
        dict(shape=[1, 3, 40, 36], f='f12', up=1, down=2, padding=[-3, -2, 4, -1], gain=1, flip=True),
        dict(shape=[2, 2, 9, 13], f='f35', up=[2, 1], down=[1, 3], padding=[3, 1, 0, 2], gain=1.5, flip=False),
        dict(shape=[2, 2, 9, 13], f='f35', up=[3, 2], down=[2, 2], padding=[2, 2, 1, 3], gain=0.5, flip=True),
        dict(shape=[1, 1, 7, 7], f=None, up=2, down=1, padding=0, gain=1, flip=False),
        dict(shape=[1, 2, 12, 12], f=None, up=1, down=3, padding=[1, 0, 2, 0], gain=2, flip=False),
        dict(shape=[1, 1, 4, 4], f='f4', up=1, down=1, padding=[2, 2, 2, 2], gain=1, flip=False),
        dict(shape=[1, 1, 5, 5], f='f4', up=1, down=1, padding=[1, 1, 1, 1], gain=4, flip=False),
    ]
    filters = dict(f4=f4, f12=f12, f35=f35)
    arrays, meta = {}, []
    for i, c in enumerate(cases):
        x = rnd(100 + i, *c['shape']).requires_grad_(True)
        f = filters[c['f']] if c['f'] else None
        y = R_up.upfirdn2d(x, f, up=c['up'], down=c['down'], padding=c['padding'], flip_filter=c['flip'], gain=c['gain'], impl='ref')
        dy = rnd(200 + i, *y.shape).requires_grad_(True)
        dx, = torch.autograd.grad(y, x, dy, create_graph=True)
        ddx = rnd(300 + i, *dx.shape)
        ddy, = torch.autograd.grad(dx, dy, ddx)                 # second order: linear op => forward of ddx
        arrays.update({f'{i}/x': x, f'{i}/y': y, f'{i}/dy': dy, f'{i}/dx': dx, f'{i}/ddx': ddx, f'{i}/ddy': ddy})
        if f is not None:
            arrays[f'{i}/f'] = f
        meta.append(c)
    # the wrappers
    x = rnd(400, 2, 3, 12, 12)
    arrays['w/x'] = x
    arrays['w/f4'] = f4
    arrays['w/f12'] = f12
    arrays['w/filter2d'] = R_up.filter2d(x, f4, padding=1, gain=2, impl='ref')
    arrays['w/filter2d_sep'] = R_up.filter2d(x, f12, flip_filter=True, impl='ref')
    arrays['w/upsample2d'] = R_up.upsample2d(x, f4, impl='ref')
    arrays['w/upsample2d_sep'] = R_up.upsample2d(x, f12, up=[2, 1], padding=[1, 0], impl='ref')
    arrays['w/downsample2d'] = R_up.downsample2d(x, f4, impl='ref')
    arrays['w/downsample2d_sep'] = R_up.downsample2d(x, f12, down=[1, 2], gain=3, impl='ref')
    # setup_filter table
    arrays['sf/1331'] = R_up.setup_filter([1, 3, 3, 1])
    arrays['sf/1331_g4_flip'] = R_up.setup_filter([1, 2, 3, 1], gain=4, flip_filter=True)
    arrays['sf/sep8'] = R_up.setup_filter([1, 2, 3, 4, 4, 3, 2, 1], gain=4)
    arrays['sf/none'] = R_up.setup_filter(None)
    arrays['sf/nonorm2d'] = R_up.setup_filter([[1, 2], [3, 4]], normalize=False)
    save('upfirdn2d', arrays, meta)


# ----------------------------------------------------------------------------- bias_act

def gen_bias_act():
    cases = []
    for act in R_ba.activation_funcs:
        cases.append(dict(act=act, shape=[2, 5, 6, 7], dim=1, bias=True, gain=None, clamp=None, alpha=None))
        cases.append(dict(act=act, shape=[2, 5, 6, 7], dim=1, bias=True, gain=1.7, clamp=0.9, alpha=(0.1 if act == 'lrelu' else None)))
    cases += [
        dict(act='lrelu', shape=[3, 8, 4, 4], dim=1, bias=True, gain=2 ** 0.5, clamp=256, alpha=None),
        dict(act='lrelu', shape=[3, 8, 4, 4], dim=1, bias=True, gain=1.0, clamp=256 * 0.5 ** 0.5, alpha=None),
        dict(act='lrelu', shape=[4, 16], dim=1, bias=True, gain=None, clamp=None, alpha=None),
        dict(act='linear', shape=[2, 3, 9, 9], dim=1, bias=True, gain=1, clamp=256, alpha=None),
        dict(act='linear', shape=[2, 6, 5, 5], dim=1, bias=False, gain=0.5 ** 0.5, clamp=None, alpha=None),
        dict(act='relu', shape=[2, 6, 5, 5], dim=1, bias=False, gain=None, clamp=None, alpha=None),
        dict(act='sigmoid', shape=[2, 1, 8, 8], dim=1, bias=True, gain=None, clamp=None, alpha=None),
        dict(act='lrelu', shape=[2, 3, 4, 5], dim=3, bias=True, gain=None, clamp=0.5, alpha=None),
        dict(act='swish', shape=[2, 3, 4, 5], dim=0, bias=True, gain=None, clamp=1.0, alpha=None),
        dict(act='linear', shape=[7], dim=0, bias=True, gain=3.0, clamp=None, alpha=None),
    ]
    arrays, meta = {}, []
    for i, c in enumerate(cases):
        x = (rnd(500 + i, *c['shape']) * 1.5).requires_grad_(True)
        b = (rnd(600 + i, c['shape'][c['dim']]) * 0.5).requires_grad_(True) if c['bias'] else None
        y = R_ba.bias_act(x, b, dim=c['dim'], act=c['act'], alpha=c['alpha'], gain=c['gain'], clamp=c['clamp'], impl='ref')
        dy = rnd(700 + i, *y.shape).requires_grad_(True)
        ins = [x] + ([b] if b is not None else [])
        g = torch.autograd.grad(y, ins, dy, create_graph=True)
        dx = g[0]
        arrays.update({f'{i}/x': x, f'{i}/y': y, f'{i}/dy': dy, f'{i}/dx': dx})
        if b is not None:
            arrays[f'{i}/b'] = b
            arrays[f'{i}/db'] = g[1]
        # second order through dx: wrt dy (always) and x (only where act'' != 0)
        ddx = rnd(800 + i, *dx.shape)
        arrays[f'{i}/ddx'] = ddx
        if dx.requires_grad:
            gg = torch.autograd.grad(dx, [dy, x], ddx, allow_unused=True)
            arrays[f'{i}/d_dy'] = gg[0] if gg[0] is not None else torch.zeros_like(dy)
            arrays[f'{i}/d_x'] = gg[1] if gg[1] is not None else torch.zeros_like(x)
        meta.append(c)
    save('bias_act', arrays, meta)


# ----------------------------------------------------------------------------- conv2d_resample

def gen_conv2d_resample():
    f4 = R_up.setup_filter([1, 3, 3, 1])
    cases = [
        dict(x=[2, 6, 12, 12], w=[5, 6, 3, 3], up=1, down=1, padding=1, groups=1, flip_weight=True),
        dict(x=[2, 6, 12, 12], w=[5, 6, 1, 1], up=1, down=1, padding=0, groups=1, flip_weight=True),
        dict(x=[2, 6, 8, 8], w=[5, 6, 3, 3], up=2, down=1, padding=1, groups=1, flip_weight=False),     # SynthesisLayer.conv0
        dict(x=[2, 6, 16, 16], w=[5, 6, 3, 3], up=1, down=2, padding=1, groups=1, flip_weight=True),    # encoders / D conv1
        dict(x=[2, 6, 16, 16], w=[5, 6, 1, 1], up=1, down=2, padding=0, groups=1, flip_weight=True),    # ResBlock / D skip
        dict(x=[2, 6, 8, 8], w=[5, 6, 1, 1], up=2, down=1, padding=0, groups=1, flip_weight=False),     # resnet-arch skip
        dict(x=[1, 8, 8, 8], w=[6, 4, 3, 3], up=2, down=1, padding=1, groups=2, flip_weight=False),     # fused modconv up (groups = N)
        dict(x=[1, 8, 10, 10], w=[6, 4, 3, 3], up=1, down=1, padding=1, groups=2, flip_weight=True),    # fused modconv
        dict(x=[1, 3, 20, 20], w=[4, 3, 7, 7], up=1, down=1, padding=3, groups=1, flip_weight=True),    # spade_encoder 7x7
        dict(x=[1, 3, 9, 11], w=[4, 3, 3, 3], up=1, down=1, padding=[1, 2, 0, 1], groups=1, flip_weight=False),  # generic fallback
        dict(x=[1, 3, 8, 8], w=[4, 3, 3, 3], up=2, down=2, padding=1, groups=1, flip_weight=True),      # up then down
        dict(x=[1, 4, 33, 33], w=[4, 4, 3, 3], up=1, down=1, padding=1, groups=1, flip_weight=True, nofilter=True),
    ]
    arrays, meta = {}, []
    for i, c in enumerate(cases):
        x = rnd(900 + i, *c['x']).requires_grad_(True)
        w = (rnd(1000 + i, *c['w']) / np.sqrt(np.prod(c['w'][1:]))).requires_grad_(True)
        f = None if c.get('nofilter') else f4
        y = R_cr.conv2d_resample(x=x, w=w, f=f, up=c['up'], down=c['down'], padding=c['padding'], groups=c['groups'], flip_weight=c['flip_weight'])
        dy = rnd(1100 + i, *y.shape)
        dx, dw = torch.autograd.grad(y, [x, w], dy)
        arrays.update({f'{i}/x': x, f'{i}/w': w, f'{i}/y': y, f'{i}/dy': dy, f'{i}/dx': dx, f'{i}/dw': dw})
        meta.append(c)
    arrays['f4'] = f4
    save('conv2d_resample', arrays, meta)


# ----------------------------------------------------------------------------- modulated_conv2d + layers

def gen_modulated_conv2d():
    f4 = R_up.setup_filter([1, 3, 3, 1])
    cases = []
    for fused in (True, False):
        cases += [
            dict(n=3, i=8, o=6, k=3, res=8, up=1, demod=True, noise='const', flip_weight=True, fused=fused),
            dict(n=3, i=8, o=6, k=3, res=8, up=2, demod=True, noise='const', flip_weight=False, fused=fused),
            dict(n=2, i=8, o=3, k=1, res=8, up=1, demod=False, noise=None, flip_weight=True, fused=fused),      # ToRGB
            dict(n=2, i=4, o=5, k=3, res=16, up=2, demod=True, noise='random', flip_weight=False, fused=fused),
            dict(n=1, i=4, o=5, k=3, res=5, up=1, demod=True, noise=None, flip_weight=True, fused=fused),
        ]
    arrays, meta = {}, []
    for idx, c in enumerate(cases):
        s = 1200 + 10 * (idx % 5)                      # fused / non-fused twins share inputs
        x = rnd(s, c['n'], c['i'], c['res'], c['res']).requires_grad_(True)
        w = rnd(s + 1, c['o'], c['i'], c['k'], c['k']).requires_grad_(True)
        st = (1 + 0.5 * rnd(s + 2, c['n'], c['i'])).requires_grad_(True)
        ores = c['res'] * c['up']
        noise = None
        if c['noise'] == 'const':
            noise = rnd(s + 3, ores, ores) * 0.3
        elif c['noise'] == 'random':
            noise = rnd(s + 3, c['n'], 1, ores, ores) * 0.3
        y = R_net.modulated_conv2d(x=x, weight=w, styles=st, noise=(noise.clone() if noise is not None else None), up=c['up'],
                                   padding=c['k'] // 2, resample_filter=f4, demodulate=c['demod'], flip_weight=c['flip_weight'],
                                   fused_modconv=c['fused'])
        dy = rnd(s + 4, *y.shape)
        dx, dw, ds = torch.autograd.grad(y, [x, w, st], dy)
        arrays.update({f'{idx}/x': x, f'{idx}/w': w, f'{idx}/s': st, f'{idx}/y': y, f'{idx}/dy': dy,
                       f'{idx}/dx': dx, f'{idx}/dw': dw, f'{idx}/ds': ds})
        if noise is not None:
            arrays[f'{idx}/noise'] = noise
        meta.append(c)
    arrays['f4'] = f4
    save('modulated_conv2d', arrays, meta)


# ----------------------------------------------------------------------------- hot-path modules (A9) and networks

def gen_layers():
    """SynthesisLayer / ToRGBLayerFull / Conv2dLayer / SynthesisBlockFull / Spade blocks with procedural
    weights, eval (fused) and train (non-fused) mode."""
    arrays, meta = {}, []

    def run(tag, mod, args, kwargs=None, train=False):
        procedural.fill_(mod)
        mod.train(train)
        with torch.no_grad():
            out = mod(*args, **(kwargs or {}))
        out = out if isinstance(out, (tuple, list)) else [out]
        for j, a in enumerate(args):
            if isinstance(a, torch.Tensor):
                arrays[f'{tag}/in{j}'] = a
        for j, o in enumerate(out):
            if isinstance(o, torch.Tensor):
                arrays[f'{tag}/out{j}'] = o
        meta.append(dict(tag=tag, train=train))

    w512 = rnd(1301, 2, 512)
    for train in (False, True):
        t = 'train' if train else 'eval'
        run(f'synth_s1_{t}', R_net.SynthesisLayer(8, 6, w_dim=512, resolution=16, conv_clamp=256),
            [rnd(1302, 2, 8, 16, 16), w512], dict(noise_mode='const', fused_modconv=not train), train)
        run(f'synth_up_{t}', R_net.SynthesisLayer(8, 6, w_dim=512, resolution=16, up=2, conv_clamp=256),
            [rnd(1303, 2, 8, 8, 8), w512], dict(noise_mode='const', fused_modconv=not train, gain=0.5 ** 0.5), train)
        run(f'torgb_{t}', R_net.ToRGBLayerFull(8, 3, w_dim=512, conv_clamp=256, is_last=True, is_style=True),
            [rnd(1304, 2, 8, 16, 16), w512], dict(fused_modconv=not train), train)
    run('conv_plain', R_net.Conv2dLayer(6, 7, kernel_size=3, activation='lrelu', conv_clamp=256), [rnd(1305, 2, 6, 12, 12)])
    run('conv_down', R_net.Conv2dLayer(6, 7, kernel_size=3, down=2), [rnd(1306, 2, 6, 12, 12)])
    run('conv_up', R_net.Conv2dLayer(6, 7, kernel_size=1, bias=False, up=2), [rnd(1307, 2, 6, 6, 6)], dict(gain=0.5 ** 0.5))
    run('conv_7x7', R_net.Conv2dLayer(3, 8, kernel_size=7, activation='relu'), [rnd(1308, 1, 3, 16, 16)])
    run('resblock_down', R_net.ResBlock(6, 8, kernel_size=4, activation='relu', down=2), [rnd(1309, 2, 6, 16, 16)])
    run('fc_lrelu', R_net.FullyConnectedLayer(12, 9, activation='lrelu', lr_multiplier=0.01), [rnd(1310, 4, 12)])
    run('fc_linear', R_net.FullyConnectedLayer(12, 9, bias_init=1), [rnd(1311, 4, 12)])
    run('dense', R_net.Dense(8, 8), [rnd(1312, 2, 8, 6, 6)])
    run('spade_norm', R_net.Spade_Norm_Block(10, 6), [rnd(1313, 2, 6, 8, 8), rnd(1314, 2, 10, 8, 8)])
    save('layers', arrays, meta)


def gen_generator():
    """GeneratorFull at the BASELINE config (channel_base 16384, channel_max 512), N = 2, eval,
    noise_mode='const', procedural weights and synthetic inputs.  Outputs stored as float16
    (network-level tolerance is 1e-2) to keep the fixture small; two intermediate feature maps kept
    in float32 subsampled form for localisation."""
    G = R_net.GeneratorFull(z_dim=0, c_dim=512, w_dim=512, img_resolution=256, img_channels=3,
                            mapping_kwargs=dict(num_layers=1),
                            synthesis_kwargs=dict(channel_base=16384, channel_max=512, num_fp16_res=3, conv_clamp=256, use_noise=True)).eval()
    procedural.fill_(G)
    inp = procedural.synth_inputs(2)
    with torch.no_grad():
        pose_feat = G.const_encoding(inp['pose'])
        stylecode, feats = G.style_encoding(inp['c'], inp['retain'])
        ws = G.mapping(inp['z'], stylecode)
        img, fimg, parsing = G(**inp, noise_mode='const')
    fp = procedural.fingerprint(G)
    names = sorted(fp)
    arrays = {
        'img': img.half(), 'finetune_img': fimg.half(), 'pred_parsing': parsing.half(),
        'pose_feat': pose_feat, 'stylecode': stylecode, 'ws0': ws[:, 0],
        'feat64': feats[2][:, :, ::4, ::4], 'feat256': feats[0][:, ::8, ::16, ::16],
        'fp_sum': np.array([fp[n][0] for n in names]), 'fp_abs': np.array([fp[n][1] for n in names]),
    }
    meta = dict(names=names, n_params=sum(p.numel() for p in G.parameters()), num_ws=int(G.num_ws),
                stats=dict(img_absmax=float(img.abs().max()), img_std=float(img.std()),
                           fimg_absmax=float(fimg.abs().max()), parsing_absmax=float(parsing.abs().max())))
    save('generator_full', arrays, [meta])


def gen_generator_n16():
    """GeneratorFull at the BASELINE config AND the BASELINE batch (N = 16, configs[1]): tile counts, strip lengths and band choices of the CUDA
    kernels depend on N * H * W, so parity is pinned at the measured batch as well.  To keep the fixture small the three outputs are stored
    float16 and spatially subsampled (images stride 2, parsing logits stride 4); the reference's argmax label map is stored in full (uint8) so
    that a test can feed BOTH sides the same labels and hold the fine-tuned image to a max-abs bound (networks.py:5823-5826 makes it a
    discontinuous function of the logits)."""
    G = R_net.GeneratorFull(z_dim=0, c_dim=512, w_dim=512, img_resolution=256, img_channels=3,
                            mapping_kwargs=dict(num_layers=1),
                            synthesis_kwargs=dict(channel_base=16384, channel_max=512, num_fp16_res=3, conv_clamp=256, use_noise=True)).eval()
    procedural.fill_(G)
    inp = procedural.synth_inputs(16, seed=2468)
    with torch.no_grad():
        img, fimg, parsing = G(**inp, noise_mode='const')
    label = torch.argmax(torch.softmax(parsing, dim=1), dim=1).to(torch.uint8)
    arrays = {'img': img[:, :, ::2, ::2].half(), 'finetune_img': fimg[:, :, ::2, ::2].half(), 'pred_parsing': parsing[:, :, ::4, ::4].half(),
              'label': label}
    meta = dict(seed=2468, batch=16, stats=dict(img_absmax=float(img.abs().max()), fimg_absmax=float(fimg.abs().max()),
                                                parsing_absmax=float(parsing.abs().max())))
    save('generator_full_n16', arrays, [meta])


def gen_generator_labels():
    """The reference's argmax label map for the N = 2 case of generator_full.npz (same weights, same inputs), so the fine-tuned image of that
    fixture can be compared with the labels pinned (see gen_generator_n16)."""
    G = R_net.GeneratorFull(z_dim=0, c_dim=512, w_dim=512, img_resolution=256, img_channels=3,
                            mapping_kwargs=dict(num_layers=1),
                            synthesis_kwargs=dict(channel_base=16384, channel_max=512, num_fp16_res=3, conv_clamp=256, use_noise=True)).eval()
    procedural.fill_(G)
    inp = procedural.synth_inputs(2)
    with torch.no_grad():
        img, fimg, parsing = G(**inp, noise_mode='const')
    label = torch.argmax(torch.softmax(parsing, dim=1), dim=1).to(torch.uint8)
    save('generator_full_labels', {'label': label, 'finetune_img': fimg.half()}, [dict(batch=2)])


def gen_generator_512():
    """Generator_512 (the only 512-px network in the tree) at channel_base 16384, N = 1, eval, const noise; fp16-stored output."""
    G = R_net.Generator_512(z_dim=0, c_dim=512, w_dim=512, img_resolution=512, img_channels=3, mapping_kwargs=dict(num_layers=1),
                            synthesis_kwargs=dict(channel_base=16384, channel_max=512, num_fp16_res=0, conv_clamp=256, use_noise=True)).eval()
    procedural.fill_(G)
    inp = procedural.synth_inputs_512(1)
    with torch.no_grad():
        img = G(**inp, noise_mode='const')
    fp = procedural.fingerprint(G)
    names = sorted(fp)
    save('generator_512', {'img': img.half(), 'fp_sum': np.array([fp[n][0] for n in names]), 'fp_abs': np.array([fp[n][1] for n in names])},
         [dict(names=names, n_params=sum(p.numel() for p in G.parameters()), num_ws=int(G.num_ws), img_absmax=float(img.abs().max()))])


def gen_generator_512_n16():
    """Generator_512 at the BASELINE batch of configs[2] (N = 16): the band / tile choices of the CUDA kernels depend on N * H * W.  Output stored
    float16, every 4th pixel."""
    G = R_net.Generator_512(z_dim=0, c_dim=512, w_dim=512, img_resolution=512, img_channels=3, mapping_kwargs=dict(num_layers=1),
                            synthesis_kwargs=dict(channel_base=16384, channel_max=512, num_fp16_res=0, conv_clamp=256, use_noise=True)).eval()
    procedural.fill_(G)
    inp = procedural.synth_inputs_512(16, seed=8642)
    with torch.no_grad():
        img = G(**inp, noise_mode='const')
    save('generator_512_n16', {'img': img[:, :, ::4, ::4].half()}, [dict(seed=8642, batch=16, img_absmax=float(img.abs().max()), img_std=float(img.std()))])


def gen_discriminator():
    """Discriminator (fp32 blocks) at the BASELINE widths on a 4-image batch: logits, the R1 gradient wrt the image, the R1 penalty and a
    few parameter gradients of the penalty (a full double backward through conv / upfirdn2d / bias_act-lrelu / mbstd / FC)."""
    D = R_net.Discriminator(c_dim=512, img_resolution=256, img_channels=3, channel_base=16384, channel_max=512, num_fp16_res=0,
                            conv_clamp=256, epilogue_kwargs=dict(mbstd_group_size=4))
    procedural.fill_(D)
    g = torch.Generator().manual_seed(4321)
    img = (torch.rand(4, 3, 256, 256, generator=g) * 2 - 1).requires_grad_(True)
    c = torch.randn(4, 512, generator=g)
    logits = D(img, c)
    r1_grad, = torch.autograd.grad(logits.sum(), img, create_graph=True)
    penalty = r1_grad.square().sum([1, 2, 3])
    (penalty.mean() * 5).backward()                       # r1_gamma / 2 = 5
    names = ['b256.fromrgb.weight', 'b256.conv0.bias', 'b64.conv1.weight', 'b16.skip.weight', 'b4.conv.weight', 'b4.fc.bias', 'b4.out.weight', 'mapping.fc3.weight']
    params = dict(D.named_parameters())
    arrays = {'logits': logits, 'r1_grad': r1_grad[:, :, ::4, ::4], 'penalty': penalty}
    for n in names:
        gr = params[n].grad
        arrays['grad/' + n] = gr if gr.numel() <= 70000 else gr.flatten()[::37]
    fp = procedural.fingerprint(D)
    pn = sorted(fp)
    arrays['fp_sum'] = np.array([fp[n][0] for n in pn])
    arrays['fp_abs'] = np.array([fp[n][1] for n in pn])
    save('discriminator', arrays, [dict(names=pn, grad_names=names, n_params=sum(p.numel() for p in D.parameters()))])


def _grad_summary(module, seed_tag):
    """Per-parameter gradient fingerprint: L2 norm and the projections on four fixed +-1 vectors (drawn from a generator seeded by the parameter's
    name), float64.  Parameters whose grad is None (unused in the phase) get NaN norms."""
    import zlib
    names, norms, projs = [], [], []
    for n, p_ in module.named_parameters():
        names.append(n)
        if p_.grad is None:
            norms.append(float('nan')); projs.append([float('nan')] * 4)
            continue
        g = p_.grad.detach().double().flatten()
        gen = torch.Generator().manual_seed(zlib.crc32((seed_tag + n).encode()))
        signs = torch.randint(0, 2, (4, g.numel()), generator=gen, dtype=torch.int8).double() * 2 - 1
        norms.append(float(g.norm())); projs.append([float(v) for v in signs @ g])
    return names, np.array(norms), np.array(projs)


def gen_training_step(num_fp16_res=3, name='training_step'):
    """The reference's own loss object (training/loss_wo_flow_fullbody.py:32-254, StyleGAN2Loss.accumulate_gradients) on BASELINE configs[3] at batch 2:
    phases Gmain, Dmain and Dreg with the weights of train.sh (l1 40, mask 20, r1_gamma 10, pl 0, contextual 0, style mixing 0 as
    train_wo_flow_fullbody.py:219 sets it) and vgg_weight 0 (the VGG checkpoint is not available).  G and D are the unmodified reference modules in
    train mode with procedural weights; noise_strength is zeroed (train mode draws fresh noise per call, which no second implementation can
    reproduce).  Stored: the loss terms the reference reports (training_stats.report is tapped, a harness shim), and per-parameter gradient
    fingerprints (norm + four +-1 projections) of every phase, plus a few gradient tensors (every 37th element of the large ones).
    With num_fp16_res = 3 the R1 phase is degenerate on these weights (the reference's fp16 blocks flush the ~1e-8 gradients to exact zeros: penalty
    0.0); its parity is pinned by discriminator.npz (fp32 blocks) and by the fp32 variant of this fixture."""
    import training.loss_wo_flow_fullbody as R_loss
    from torch_utils import training_stats
    torch.manual_seed(0)
    G = R_net.GeneratorFull(z_dim=0, c_dim=512, w_dim=512, img_resolution=256, img_channels=3, mapping_kwargs=dict(num_layers=1),
                            synthesis_kwargs=dict(channel_base=16384, channel_max=512, num_fp16_res=num_fp16_res, conv_clamp=256, use_noise=True))
    D = R_net.Discriminator(c_dim=512, img_resolution=256, img_channels=3, channel_base=16384, channel_max=512, num_fp16_res=num_fp16_res,
                            conv_clamp=256, epilogue_kwargs=dict(mbstd_group_size=4))
    procedural.fill_(G)
    procedural.fill_(D)
    with torch.no_grad():
        for n_, p_ in G.named_parameters():
            if n_.endswith('noise_strength'):
                p_.zero_()
    G.train().requires_grad_(True)
    D.train().requires_grad_(True)
    loss = R_loss.StyleGAN2Loss(device=torch.device('cpu'), G_mapping=G.mapping, G_synthesis=G.synthesis, G_const_encoding=G.const_encoding,
                                G_style_encoding=G.style_encoding, D=D, augment_pipe=None, style_mixing_prob=0, r1_gamma=10, pl_weight=0,
                                l1_weight=40, vgg_weight=0, contextual_weight=0, mask_weight=20)
    b = procedural.synth_inputs(2, seed=1234)
    g = torch.Generator().manual_seed(1234 + 99)
    real_img = torch.randint(0, 256, (2, 3, 256, 256), generator=g).float() / 127.5 - 1
    gt_parsing = torch.randint(0, 6, (2, 1, 256, 256), generator=g).float()
    reported = {}
    real_report = training_stats.report
    tap = lambda name, value: reported.__setitem__(name, torch.as_tensor(value).detach().double().mean().item())
    training_stats.report = tap
    R_loss.training_stats.report = tap
    arrays, meta = {}, {}
    full = {'Gmain': ['synthesis.b256.conv1.weight', 'synthesis.b256.torgb.weight', 'synthesis.b16.conv1.affine.weight', 'mapping.fc0.weight'],
            'Dmain': ['b256.fromrgb.weight', 'b32.conv0.weight', 'b4.out.weight'], 'Dreg': ['b256.fromrgb.weight', 'b32.conv0.weight', 'b4.out.weight']}
    try:
        for phase, gain, module, tag in (('Gmain', 1, G, 'G'), ('Dmain', 1, D, 'D'), ('Dreg', 16, D, 'D')):
            G.zero_grad(set_to_none=True)
            D.zero_grad(set_to_none=True)
            G.requires_grad_(tag == 'G')                 # training_loop_wo_flow_fullbody.py:489-491: only the phase's module takes gradients
            D.requires_grad_(tag == 'D')
            reported.clear()
            loss.accumulate_gradients(phase=phase, real_img=real_img, gen_z=b['z'], style_input=b['c'], retain=b['retain'], pose=b['pose'],
                                      denorm_upper_input=b['denorm_upper_input'], denorm_lower_input=b['denorm_lower_input'],
                                      denorm_upper_mask=b['denorm_upper_mask'], denorm_lower_mask=b['denorm_lower_mask'], gt_parsing=gt_parsing,
                                      sync=True, gain=gain)
            names, norms, projs = _grad_summary(module, phase + '/')
            arrays[phase + '/norm'], arrays[phase + '/proj'] = norms, projs
            meta[phase] = dict(names=names, reported=dict(reported), gain=gain)
            params = dict(module.named_parameters())
            for n in full[phase]:
                gr = params[n].grad.detach()
                arrays[f'{phase}/grad/{n}'] = gr.clone() if gr.numel() <= 70000 else gr.flatten()[::37].clone()
            print(phase, {k: float('%.6g' % v) for k, v in reported.items()}, 'params with grad:', int(np.isfinite(norms).sum()), '/', len(names), flush=True)
    finally:
        training_stats.report = real_report
    meta['full'] = full
    meta['batch'] = 2
    meta['num_fp16_res'] = num_fp16_res
    save(name, arrays, [meta])


def gen_training_step_fp32():
    """The same with every block of G and D in fp32 (num_fp16_res = 0): the fixture that separates the rounding of OUR tensor-core training path from
    the rounding of the reference's own fp16 blocks (whose CPU and GPU convolutions already differ by a few per cent in the deepest gradients)."""
    gen_training_step(num_fp16_res=0, name='training_step_fp32')


if __name__ == '__main__':
    which = sys.argv[1:] or ['upfirdn2d', 'bias_act', 'conv2d_resample', 'modulated_conv2d', 'layers', 'generator', 'discriminator', 'generator_512',
                             'generator_labels', 'generator_n16', 'generator_512_n16', 'training_step', 'training_step_fp32']      # (the last two: ~5 + ~3 min of CPU)
    for w in which:
        globals()['gen_' + w]()
