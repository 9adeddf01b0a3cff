"""One real training iteration (Gmain + Dmain + Dreg/R1) of GeneratorFull + Discriminator on the GPU through the sm_100a operators:
losses are finite, every phase moves its network's parameters, gradients live in the flat all-reduce bucket, and the R1 phase (double
backward) runs under no_weight_gradients.  Numerical parity of the pieces is covered by test_gpu_ops / test_discriminator."""
import pytest
import torch

import procedural
from pasta_gan_b200 import networks as N
from pasta_gan_b200.training import TryOnTrainer, synth_training_batch

pytestmark = pytest.mark.gpu


def test_training_iteration_runs_and_updates():
    dev = 'cuda'
    torch.manual_seed(0)
    G = N.build_generator_full()
    D = N.build_discriminator(num_fp16_res=3)
    procedural.fill_(G)
    procedural.fill_(D)
    G.to(dev).train().requires_grad_(True)
    D.to(dev).train().requires_grad_(True)
    tr = TryOnTrainer(G, D)
    g0 = G.synthesis.b64.conv1.weight.detach().clone()
    d0 = D.b64.conv0.weight.detach().clone()
    batch = synth_training_batch(2, device=dev)
    stats = tr.step(batch)                                     # it = 0: includes the R1 phase
    assert {'G_adv', 'G_l1', 'G_mask', 'D_gen', 'D_real', 'r1_penalty'} <= set(stats)
    assert all(torch.isfinite(v).all() for v in stats.values()), stats
    assert not torch.equal(G.synthesis.b64.conv1.weight.detach(), g0)
    assert not torch.equal(D.b64.conv0.weight.detach(), d0)
    assert G.synthesis.b64.conv1.weight.grad.data_ptr() >= tr.g_bucket.flat.data_ptr()
    stats2 = tr.step(batch)                                    # it = 1: no R1
    assert 'r1_penalty' not in stats2 and all(torch.isfinite(v).all() for v in stats2.values())
