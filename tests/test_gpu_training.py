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
    d0 = D.b16.conv0.weight.detach().clone()          # an fp32 block (the fp16 blocks' tiny gradients can underflow to exact zeros at this batch size)
    batch = synth_training_batch(2, device=dev)
    stats = tr.step(batch)                                     # it = 0: includes the R1 phase
    assert {'G_adv', 'G_l1', 'G_mask', 'D_gen', 'D_real', 'r1_penalty'} <= set(stats)
    assert all(torch.isfinite(v).all() for v in stats.values()), stats
    assert not torch.equal(G.synthesis.b64.conv1.weight.detach(), g0)
    assert not torch.equal(D.b16.conv0.weight.detach(), d0)
    assert G.synthesis.b64.conv1.weight.grad.data_ptr() >= tr.g_bucket.flat.data_ptr()
    stats2 = tr.step(batch)                                    # it = 1: no R1
    assert 'r1_penalty' not in stats2 and all(torch.isfinite(v).all() for v in stats2.values())


def test_captured_training_phases_match_eager():
    """TryOnTrainer.capture(): each phase replayed from CUDA graphs (zero-grad + forward + backward | NaN guard + Adam, all-reduce between them) leaves
    the networks where the op-by-op step leaves them.  Random noise is switched off (noise_strength = 0) so both runs see the same arithmetic."""
    import copy
    dev = 'cuda'
    torch.manual_seed(0)
    G = N.build_generator_full()
    D = N.build_discriminator(num_fp16_res=3)
    procedural.fill_(G)
    procedural.fill_(D)
    with torch.no_grad():
        for n_, p_ in G.named_parameters():
            if n_.endswith('noise_strength'):
                p_.zero_()
    G.to(dev).train().requires_grad_(True)
    D.to(dev).train().requires_grad_(True)
    G2, D2 = copy.deepcopy(G), copy.deepcopy(D)
    batch = synth_training_batch(2, device=dev)
    eager = TryOnTrainer(G, D)
    graphed = TryOnTrainer(G2, D2, capturable=True)
    g_init = {k: v.detach().clone() for k, v in G2.state_dict().items()}
    d_init = {k: v.detach().clone() for k, v in D2.state_dict().items()}
    graphed.capture(batch, warmup=2)
    # capture warm-up ran real optimizer steps: rewind parameters, buffers and optimizer state so both trainers start from the same point
    G2.load_state_dict(g_init); D2.load_state_dict(d_init)
    for opt in (graphed.g_opt, graphed.d_opt):
        for st in opt.state.values():
            for v in st.values():
                if torch.is_tensor(v):
                    v.zero_()
    for _ in range(2):
        s_e = eager.step(batch)
        s_g = graphed.step(batch)
    torch.cuda.synchronize()
    for k in s_e:
        assert abs(float(s_e[k]) - float(s_g[k])) <= 2e-3 * max(1.0, abs(float(s_e[k]))), (k, float(s_e[k]), float(s_g[k]))
    w_e, w_g = G.synthesis.b64.conv1.weight.detach(), G2.synthesis.b64.conv1.weight.detach()
    # Adam's first steps are sign-like (m / sqrt(v)): an element whose gradient is at rounding level may move by +-lr in either run, so the bound is a few lr
    assert float((w_e - w_g).abs().max()) < 5e-3 * float(w_e.abs().max()) and float((w_e - w_g).abs().mean()) < 2e-4 * float(w_e.abs().max())
    d_e, d_g = D.b64.conv0.weight.detach(), D2.b64.conv0.weight.detach()
    assert float((d_e - d_g).abs().max()) < 5e-3 * float(d_e.abs().max()) and float((d_e - d_g).abs().mean()) < 2e-4 * float(d_e.abs().max())
