import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import procedural
from pasta_gan_b200 import networks as N
from pasta_gan_b200.training import TryOnTrainer, synth_training_batch
from pasta_gan_b200.torch_utils.ops import conv2d_gradfix as G_
dev = 'cuda'
torch.manual_seed(0)
G = N.build_generator_full(); D = N.build_discriminator(num_fp16_res=3)
procedural.fill_(G); procedural.fill_(D)
G.to(dev).train().requires_grad_(True); D.to(dev).train().requires_grad_(True)
tr = TryOnTrainer(G, D)
batch = synth_training_batch(2, device=dev)
def report(tag, bucket):
    f = bucket.flat
    print(tag, 'nan', int(torch.isnan(f).sum()), 'inf', int(torch.isinf(f).sum()), 'absmax', float(torch.nan_to_num(f).abs().max()), 'nonzero', int((f != 0).sum()))
    off = 0
    bad = []
    for (name, p) in [(n, p) for n, p in (list(G.named_parameters()) + list(D.named_parameters())) if any(p is q for q in bucket.params)]:
        pass
for fmt in ('fp16', 'bf16'):
    G_.tensor_core_format = fmt
    tr.d_bucket.zero()
    import types
    # run d_main backward only (no optimizer step)
    tr.d_phase(batch, True, False, finish=False)
    report(f'd_main fwd={fmt}', tr.d_bucket)
    names = dict(D.named_parameters())
    for n_, p_ in names.items():
        if p_.grad is not None and not torch.isfinite(p_.grad).all():
            print('   non-finite grad:', n_, tuple(p_.shape)); 
G_.tensor_core_training = False
tr.d_phase(batch, True, False, finish=False)
report('d_main library', tr.d_bucket)
