"""One PASTA-GAN training iteration on the sm_100a operator path (BASELINE.json configs[3]).

Mirrors the arithmetic of the reference's loss (training/loss_wo_flow_fullbody.py:106-254: non-saturating logistic GAN loss on the
coarse and the fine-tuned image, L1 x 40, parsing cross-entropy x 20, R1 with gamma 10 evaluated by a double backward under
conv2d_gradfix.no_weight_gradients()) and of the loop's phase schedule (training_loop_wo_flow_fullbody.py:332-343, :484-518:
Gmain and Dmain every iteration, Dreg every 16th with lazy-regularisation Adam hyper-parameters, NaN guard, Adam step).  VGG /
contextual losses need downloaded weights and are left out (vgg_weight = 0), the augmentation pipe is off (aug = noaug), as stated in
BASELINE.md.  Gradients are averaged across ranks with one flat all-reduce per phase (data_parallel.FlatGradBucket).
"""
import numpy as np
import torch

from . import data_parallel as dp
from .torch_utils.ops import conv2d_gradfix


class TryOnTrainer:
    def __init__(self, G, D, lr=0.002, r1_gamma=10.0, l1_weight=40.0, mask_weight=20.0, d_reg_interval=16, group=None):
        self.G, self.D, self.group = G, D, group
        self.r1_gamma, self.l1_weight, self.mask_weight, self.d_reg_interval = r1_gamma, l1_weight, mask_weight, d_reg_interval
        conv2d_gradfix.enabled = True                          # training_loop_wo_flow_fullbody.py:255
        self.g_bucket = dp.FlatGradBucket(G.parameters())
        self.d_bucket = dp.FlatGradBucket(D.parameters())
        mb = d_reg_interval / (d_reg_interval + 1)             # lazy regularisation (training_loop...py:336-343)
        self.g_opt = torch.optim.Adam(self.g_bucket.params, lr=lr, betas=(0.0, 0.99), eps=1e-8)
        self.d_opt = torch.optim.Adam(self.d_bucket.params, lr=lr * mb, betas=(0.0 ** mb, 0.99 ** mb), eps=1e-8)
        self.ce = torch.nn.CrossEntropyLoss(reduction='none')
        self.it = 0

    # --- forward helpers -----------------------------------------------------------------------------------------
    def _run_G(self, b, stylecode, feats):
        G = self.G
        pose_feat = G.const_encoding(b['pose'])
        ws = G.mapping(b['z'], stylecode)
        cat = {str(f.shape[2]): f for f in feats}
        return G.synthesis(ws, pose_feat, cat, b['denorm_upper_input'], b['denorm_lower_input'], b['denorm_upper_mask'], b['denorm_lower_mask'])

    def _finish(self, bucket, opt):
        bucket.allreduce(self.group)
        bucket.sanitize()
        opt.step()

    # --- phases ----------------------------------------------------------------------------------------------------
    def g_main(self, b):
        self.g_bucket.zero()
        self.D.requires_grad_(False)
        stylecode, feats = self.G.style_encoding(b['c'], b['retain'])
        img, fimg, parsing = self._run_G(b, stylecode, feats)
        sp = torch.nn.functional.softplus
        loss_adv = (sp(-self.D(img, stylecode)).mean() + sp(-self.D(fimg, stylecode)).mean()) / 2
        loss_l1 = (torch.nn.functional.l1_loss(img, b['real_img']) + torch.nn.functional.l1_loss(fimg, b['real_img'])) / 2 * self.l1_weight
        loss_mask = self.ce(parsing, b['gt_parsing'].long()[:, 0]).mean() * self.mask_weight
        loss = loss_adv + loss_l1 + loss_mask
        loss.backward()
        self.D.requires_grad_(True)
        self._finish(self.g_bucket, self.g_opt)
        return dict(G_adv=loss_adv.detach(), G_l1=loss_l1.detach(), G_mask=loss_mask.detach())

    def d_phase(self, b, do_main, do_r1):
        self.d_bucket.zero()
        sp = torch.nn.functional.softplus
        out = {}
        with torch.no_grad():
            stylecode, feats = self.G.style_encoding(b['c'], b['retain'])
        gain = self.d_reg_interval if (do_r1 and not do_main) else 1
        if do_main:
            with torch.no_grad():
                img, fimg, _ = self._run_G(b, stylecode, feats)
            loss_gen = (sp(self.D(img, stylecode)).mean() + sp(self.D(fimg, stylecode)).mean()) / 2
            loss_gen.backward()
            out['D_gen'] = loss_gen.detach()
        real = b['real_img'].detach().requires_grad_(do_r1)
        logits = self.D(real, stylecode)
        loss_real = sp(-logits) if do_main else 0
        loss_r1 = 0
        if do_r1:
            with conv2d_gradfix.no_weight_gradients():
                grads, = torch.autograd.grad(outputs=[logits.sum()], inputs=[real], create_graph=True, only_inputs=True)
            penalty = grads.square().sum([1, 2, 3])
            loss_r1 = penalty * (self.r1_gamma / 2)
            out['r1_penalty'] = penalty.mean().detach()
        (logits * 0 + loss_real + loss_r1).mean().mul(gain).backward()
        if do_main:
            out['D_real'] = loss_real.mean().detach()
        self._finish(self.d_bucket, self.d_opt)
        return out

    def step(self, batch):
        """Gmain + Dmain, plus Dreg (R1) on every ``d_reg_interval``-th iteration.  ``batch`` is this rank's shard."""
        stats = {}
        stats.update(self.g_main(batch))
        stats.update(self.d_phase(batch, do_main=True, do_r1=False))
        if self.it % self.d_reg_interval == 0:
            stats.update(self.d_phase(batch, do_main=False, do_r1=True))
        self.it += 1
        return stats


def synth_training_batch(batch, seed=1234, device='cpu'):
    """Generator inputs (synthetic.synth_inputs layout) + a real image and a 6-class parsing map."""
    from . import synthetic
    b = synthetic.synth_inputs(batch, seed=seed, device=device)
    g = torch.Generator().manual_seed(seed + 99)
    b['real_img'] = (torch.randint(0, 256, (batch, 3, 256, 256), generator=g).float() / 127.5 - 1).to(device)
    b['gt_parsing'] = torch.randint(0, 6, (batch, 1, 256, 256), generator=g).float().to(device)
    return b
