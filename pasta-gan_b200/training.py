"""One PASTA-GAN training iteration on the sm_100a operator path (BASELINE.json configs[3]).

Mirrors the arithmetic of the reference's loss (training/loss_wo_flow_fullbody.py:106-254: non-saturating logistic GAN loss on the
coarse and the fine-tuned image, L1 x 40, parsing cross-entropy x 20, R1 with gamma 10 evaluated by a double backward under
conv2d_gradfix.no_weight_gradients()) and of the loop's phase schedule (training_loop_wo_flow_fullbody.py:332-343, :484-518:
Gmain and Dmain every iteration, Dreg every 16th with lazy-regularisation Adam hyper-parameters for BOTH networks -- G_reg_interval = 4,
D_reg_interval = 16, so lr and betas are scaled by 4/5 and 16/17 even though the G regulariser itself is a no-op (pl_weight = 0) -- NaN guard, Adam
step).  VGG / contextual losses need downloaded weights and are left out (vgg_weight = 0), the augmentation pipe is off (aug = noaug), as stated
in BASELINE.md.  Gradients are averaged across ranks with one flat all-reduce per phase (data_parallel.FlatGradBucket), launched asynchronously on
a communication stream; parameters that get no gradient in a phase see zeros where the reference has grad = None (an Adam update with zero first
and second moments leaves the parameter unchanged, so the trajectories agree).

``capture()`` records each phase as two CUDA graphs (zero-grad + forward + backward | NaN guard + Adam step) with the all-reduce between them: the
training step is host-issue bound when launched op by op (~5 100 launches per iteration), a graph replay is not.
"""
import numpy as np
import torch

from . import data_parallel as dp
from .torch_utils.ops import conv2d_gradfix


class TryOnTrainer:
    def __init__(self, G, D, lr=0.0025, r1_gamma=10.0, l1_weight=40.0, mask_weight=20.0, d_reg_interval=16, g_reg_interval=4, group=None, capturable=False, allow_tf32=False):
        self.G, self.D, self.group = G, D, group
        self.r1_gamma, self.l1_weight, self.mask_weight, self.d_reg_interval = r1_gamma, l1_weight, mask_weight, d_reg_interval
        conv2d_gradfix.enabled = True                          # training_loop_wo_flow_fullbody.py:255
        # the convolutions and GEMMs left on the library (strided / transposed forms, small layers) run in full fp32 as in the reference's loop
        # (:243, :253).  allow_tf32=True is ~30 % faster (75 vs 59 img/s on one B200) but leaves the style encoder's gradients off by up to 13 % of their
        # norm against 2-4 % (tests/test_training_parity.py; it is the LARGE strided convolutions that cause it, not the small ones)
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = bool(allow_tf32)
        self.g_bucket = dp.FlatGradBucket(G.parameters())
        self.d_bucket = dp.FlatGradBucket(D.parameters())
        mb = d_reg_interval / (d_reg_interval + 1)             # lazy regularisation (training_loop...py:336-343), applied to G as well (G_reg_interval = 4)
        mg = g_reg_interval / (g_reg_interval + 1) if g_reg_interval else 1.0
        self.g_opt = torch.optim.Adam(self.g_bucket.params, lr=lr * mg, betas=(0.0 ** mg, 0.99 ** mg), eps=1e-8, capturable=capturable)
        self.d_opt = torch.optim.Adam(self.d_bucket.params, lr=lr * mb, betas=(0.0 ** mb, 0.99 ** mb), eps=1e-8, capturable=capturable)
        # class-weighted parsing loss, weighted mean over the pixels (loss_wo_flow_fullbody.py:56-57)
        self.ce = torch.nn.CrossEntropyLoss(ignore_index=255, weight=torch.tensor([1.0, 2.0, 2.0, 3.0, 3.0, 3.0], device=self.g_bucket.flat.device))
        self.it = 0
        self.graphs = None
        self.comm_stream = torch.cuda.Stream(self.g_bucket.flat.device) if self.g_bucket.flat.is_cuda else None

    # --- forward helpers -----------------------------------------------------------------------------------------
    def _run_G(self, b, stylecode, feats):
        G = self.G
        pose_feat = G.const_encoding(b['pose'])
        ws = G.mapping(b['z'], stylecode)
        cat = {str(f.shape[2]): f for f in feats}
        return G.synthesis(ws, pose_feat, cat, b['denorm_upper_input'], b['denorm_lower_input'], b['denorm_upper_mask'], b['denorm_lower_mask'])

    def _exchange(self, bucket):
        """Average the phase's gradients across ranks: one asynchronous all-reduce of the flat buffer on the communication stream, so the host goes
        straight on to issuing the next phase's forward while NCCL runs; the optimizer step (same stream order) waits for it."""
        if self.comm_stream is None or not (torch.distributed.is_available() and torch.distributed.is_initialized()):
            bucket.allreduce(self.group)
            return
        cur = torch.cuda.current_stream()
        self.comm_stream.wait_stream(cur)
        with torch.cuda.stream(self.comm_stream):
            bucket.allreduce(self.group)
        cur.wait_stream(self.comm_stream)

    def _apply(self, bucket, opt):
        bucket.sanitize()
        opt.step()

    def _finish(self, bucket, opt):
        self._exchange(bucket)
        self._apply(bucket, opt)

    # --- phases ----------------------------------------------------------------------------------------------------
    def g_main_losses(self, b):
        """The three terms of the generator loss (loss_wo_flow_fullbody.py:118-172 with vgg_weight = 0): adversarial (mean of the coarse and the
        fine-tuned image), L1 x l1_weight (likewise), class-weighted parsing cross-entropy x mask_weight."""
        stylecode, feats = self.G.style_encoding(b['c'], b['retain'])
        img, fimg, parsing = self._run_G(b, stylecode, feats)
        sp = torch.nn.functional.softplus
        loss_adv = (sp(-self.D(img, stylecode)).mean() + sp(-self.D(fimg, stylecode)).mean()) / 2
        loss_l1 = (torch.nn.functional.l1_loss(img, b['real_img']) + torch.nn.functional.l1_loss(fimg, b['real_img'])) / 2 * self.l1_weight
        loss_mask = self.ce(parsing, b['gt_parsing'].long()[:, 0]) * self.mask_weight
        return loss_adv, loss_l1, loss_mask

    def g_main(self, b, finish=True):
        self.g_bucket.zero()
        self.D.requires_grad_(False)
        loss_adv, loss_l1, loss_mask = self.g_main_losses(b)
        loss = loss_adv + loss_l1 + loss_mask
        loss.backward()
        self.D.requires_grad_(True)
        if finish:
            self._finish(self.g_bucket, self.g_opt)
        return dict(G_adv=loss_adv.detach(), G_l1=loss_l1.detach(), G_mask=loss_mask.detach())

    def d_phase(self, b, do_main, do_r1, finish=True):
        # the R1 phase needs fp32 products (second-order gradients of ~1e-7-sized values, see conv2d_gradfix): tensor cores off for it
        with conv2d_gradfix.tensor_cores(not do_r1):
            return self._d_phase(b, do_main, do_r1, finish)

    def _d_phase(self, b, do_main, do_r1, finish=True):
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
        if finish:
            self._finish(self.d_bucket, self.d_opt)
        return out

    # --- CUDA graphs ---------------------------------------------------------------------------------------------------
    def capture(self, batch, warmup=3):
        """Record the three phases as CUDA graphs over static copies of ``batch`` (later batches are copied into them by ``step``).  Needs
        ``capturable=True`` optimizers.  Each phase = graph A (zero-grad, forward, backward) -> eager flat all-reduce -> graph B (NaN guard, Adam)."""
        assert self.g_opt.defaults.get('capturable') and self.d_opt.defaults.get('capturable'), 'construct the trainer with capturable=True'
        self.static = {k: v.clone() for k, v in batch.items()}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):                       # allocator pools, cuDNN plans, optimizer state, NCCL communicators
                self.g_main(self.static)
                self.d_phase(self.static, True, False)
                self.d_phase(self.static, False, True)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        phases = dict(g_main=(lambda: self.g_main(self.static, finish=False), self.g_bucket, self.g_opt),
                      d_main=(lambda: self.d_phase(self.static, True, False, finish=False), self.d_bucket, self.d_opt),
                      d_reg=(lambda: self.d_phase(self.static, False, True, finish=False), self.d_bucket, self.d_opt))
        graphs = {}
        pool = None
        for name, (fn, bucket, opt) in phases.items():
            ga, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(ga, pool=pool):
                stats = fn()
            pool = ga.pool()
            with torch.cuda.graph(gb, pool=pool):
                self._apply(bucket, opt)
            graphs[name] = (ga, gb, bucket, stats)
        self.graphs = graphs
        return self

    def _replay(self, name):
        ga, gb, bucket, stats = self.graphs[name]
        ga.replay()
        self._exchange(bucket)
        gb.replay()
        return stats

    def step(self, batch):
        """Gmain + Dmain, plus Dreg (R1) on every ``d_reg_interval``-th iteration.  ``batch`` is this rank's shard."""
        if self.graphs is not None:
            if batch is not self.static:
                for k, v in batch.items():
                    self.static[k].copy_(v, non_blocking=True)
            stats = dict(self._replay('g_main'))
            stats.update(self._replay('d_main'))
            if self.it % self.d_reg_interval == 0:
                stats.update(self._replay('d_reg'))
            self.it += 1
            return stats
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
