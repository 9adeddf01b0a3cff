"""Try-on inference session: the public call a user makes to run the generator on one GPU.

Replaces the hot loop of the reference's test.py:103-131 (pinned host batch -> H2D -> style/const encoders -> mapping ->
synthesis -> D2H) with a static-shape session: inputs live in fixed device buffers, the whole forward (~300 kernel
launches) is captured once into a CUDA graph on a side stream and replayed per batch, and H2D / D2H use
pinned host buffers.  noise_mode='const' (what test.py uses) makes the forward deterministic and capturable."""
import torch

INPUT_KEYS = ('c', 'retain', 'pose', 'denorm_upper_input', 'denorm_lower_input', 'denorm_upper_mask', 'denorm_lower_mask')


class TryOnSession:
    def __init__(self, generator, example_inputs, device, use_graph=True, warmup=3):
        self.G = generator.to(device).eval().requires_grad_(False)
        self.device = torch.device(device)
        self.batch = int(example_inputs['retain'].shape[0])
        self.keys = tuple(k for k in example_inputs if k != 'z')      # GeneratorFull: INPUT_KEYS; Generator512: c, retain, pose
        self.static_in = {k: torch.empty_like(example_inputs[k], device=self.device) for k in self.keys}
        self.static_in['z'] = torch.zeros(self.batch, self.G.z_dim, device=self.device)
        self.host_in = {k: torch.empty_like(example_inputs[k], device='cpu').pin_memory() for k in self.keys}
        self.graph = None
        self.out = None
        self.stream = torch.cuda.Stream(self.device)
        self.load(example_inputs)
        with torch.cuda.stream(self.stream), torch.no_grad():
            for _ in range(warmup):
                self.out = self._forward()
        self.stream.synchronize()
        if use_graph:
            self.graph = torch.cuda.CUDAGraph()
            with torch.no_grad(), torch.cuda.graph(self.graph, stream=self.stream):
                self.out = self._forward()
        self.host_out = [torch.empty_like(o, device='cpu').pin_memory() for o in self.out[:2]]      # the image(s); parsing logits stay on device
        self.h2d_bytes = sum(self.host_in[k].numel() * self.host_in[k].element_size() for k in self.keys)
        self.d2h_bytes = sum(o.numel() * o.element_size() for o in self.host_out)

    def _forward(self):
        out = self.G(**self.static_in, noise_mode='const')
        return out if isinstance(out, (tuple, list)) else (out,)

    def load(self, inputs):
        """Device-resident inputs -> static buffers (no host traffic)."""
        with torch.cuda.stream(self.stream):
            for k in self.keys:
                self.static_in[k].copy_(inputs[k], non_blocking=True)

    def step(self):
        """One generator forward over the static buffers, on the session stream.  Returns (img, finetune_img, parsing)."""
        with torch.cuda.stream(self.stream), torch.no_grad():
            if self.graph is not None:
                self.graph.replay()
            else:
                self.out = self._forward()
        return self.out

    def step_from_host(self, host_inputs=None):
        """End-to-end step: pinned host batch -> device, forward, images -> pinned host.  Asynchronous on the session stream;
        call ``synchronize()`` before reading ``host_out``."""
        src = self.host_in if host_inputs is None else host_inputs
        with torch.cuda.stream(self.stream):
            for k in self.keys:
                self.static_in[k].copy_(src[k], non_blocking=True)
        out = self.step()
        with torch.cuda.stream(self.stream):
            for h, o in zip(self.host_out, out[:2]):
                h.copy_(o, non_blocking=True)
        return self.host_out

    def synchronize(self):
        self.stream.synchronize()
