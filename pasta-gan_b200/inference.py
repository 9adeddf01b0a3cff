"""Try-on inference session: the public call a user makes to run the generator on one GPU.

Replaces the hot loop of the reference's test.py:103-131 (pinned host batch -> H2D -> style/const encoders -> mapping ->
synthesis -> D2H) with a static-shape session: inputs live in fixed device buffers, the whole forward (~300 kernel
launches) is captured once into a CUDA graph on a side stream and replayed per batch, and H2D / D2H use
pinned host buffers.  noise_mode='const' (what test.py uses) makes the forward deterministic and capturable."""
import torch

INPUT_KEYS = ('c', 'retain', 'pose', 'denorm_upper_input', 'denorm_lower_input', 'denorm_upper_mask', 'denorm_lower_mask')


class TryOnSession:
    """``depth`` static input/output buffer sets (default 2), each with its own captured graph: ``step_from_host`` copies batch i+1 to the
    device on a copy stream while batch i computes and batch i-1's images travel back on a third stream, so the end-to-end rate is
    max(compute, PCIe) instead of their sum."""

    def __init__(self, generator, example_inputs, device, use_graph=True, warmup=3, depth=2):
        self.G = generator.to(device).eval().requires_grad_(False)
        self.device = torch.device(device)
        self.batch = int(example_inputs['retain'].shape[0])
        self.keys = tuple(k for k in example_inputs if k != 'z')      # GeneratorFull: INPUT_KEYS; Generator512: c, retain, pose
        self.depth = max(1, int(depth))
        self.slots_in = []
        for _ in range(self.depth):
            d = {k: torch.empty_like(example_inputs[k], device=self.device) for k in self.keys}
            d['z'] = torch.zeros(self.batch, self.G.z_dim, device=self.device)
            self.slots_in.append(d)
        self.static_in = self.slots_in[0]
        self.host_in = {k: torch.empty_like(example_inputs[k], device='cpu').pin_memory() for k in self.keys}
        self.stream = torch.cuda.Stream(self.device)               # compute
        self.h2d_stream = torch.cuda.Stream(self.device)
        self.d2h_stream = torch.cuda.Stream(self.device)
        with torch.cuda.stream(self.stream):
            for d in self.slots_in:
                for k in self.keys:
                    d[k].copy_(example_inputs[k], non_blocking=True)
        with torch.cuda.stream(self.stream), torch.no_grad():
            for _ in range(warmup):
                self.out = self._forward()
        self.stream.synchronize()
        self.graphs, self.slots_out = [], []
        for i in range(self.depth):
            self.static_in = self.slots_in[i]
            if use_graph:
                g = torch.cuda.CUDAGraph()
                with torch.no_grad(), torch.cuda.graph(g, stream=self.stream):
                    out = self._forward()
                self.graphs.append(g)
            else:
                with torch.cuda.stream(self.stream), torch.no_grad():
                    out = self._forward()
                self.graphs.append(None)
            self.slots_out.append(out)
        self.static_in = self.slots_in[0]
        self.graph = self.graphs[0]
        self.out = self.slots_out[0]
        self.host_outs = [[torch.empty_like(o, device='cpu').pin_memory() for o in out[:2]] for out in self.slots_out]
        self.host_out = self.host_outs[0]                          # the image(s); parsing logits stay on device
        self.ev_h2d = [torch.cuda.Event() for _ in range(self.depth)]
        self.ev_comp = [torch.cuda.Event() for _ in range(self.depth)]
        self.ev_d2h = [torch.cuda.Event() for _ in range(self.depth)]
        self._turn = 0
        self.h2d_bytes = sum(self.host_in[k].numel() * self.host_in[k].element_size() for k in self.keys)
        self.d2h_bytes = sum(o.numel() * o.element_size() for o in self.host_out)

    def refresh_weights(self):
        """Call after changing the generator's parameters (load_state_dict, EMA update): every buffer derived from them -- packed tcgen05 weight tiles,
        gamma|beta concatenations, per-parameter squared sums, the StyleBank matrices -- is rebuilt IN PLACE on the session stream, so the captured
        graphs (which hold those addresses) replay with the new weights.  No re-capture."""
        from . import networks
        from .torch_utils.ops import conv_igemm
        self.synchronize()
        with torch.cuda.stream(self.stream), torch.no_grad():
            networks.refresh_weight_caches(self.G)
            n = conv_igemm.refresh_packed_weights(self.device)
        self.stream.synchronize()
        return n

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
        """End-to-end step: pinned host batch -> device (copy stream), forward (compute stream), images -> pinned host (third stream).
        Fully asynchronous and software-pipelined over ``depth`` buffer sets; returns the pinned output tensors of THIS step, valid after
        ``synchronize()`` (or after ``depth`` further calls have been synchronised)."""
        src = self.host_in if host_inputs is None else host_inputs
        b = self._turn % self.depth
        self._turn += 1
        ins, outs, hosts = self.slots_in[b], self.slots_out[b], self.host_outs[b]
        with torch.cuda.stream(self.h2d_stream):
            self.h2d_stream.wait_event(self.ev_comp[b])            # the forward that last read this input slot has finished
            for k in self.keys:
                ins[k].copy_(src[k], non_blocking=True)
            self.ev_h2d[b].record(self.h2d_stream)
        with torch.cuda.stream(self.stream), torch.no_grad():
            self.stream.wait_event(self.ev_h2d[b])
            self.stream.wait_event(self.ev_d2h[b])                 # the previous images of this slot have left the device
            if self.graphs[b] is not None:
                self.graphs[b].replay()
            else:
                self.static_in = ins
                outs = self._forward()
            self.ev_comp[b].record(self.stream)
        with torch.cuda.stream(self.d2h_stream):
            self.d2h_stream.wait_event(self.ev_comp[b])
            for h, o in zip(hosts, outs[:2]):
                h.copy_(o, non_blocking=True)
            self.ev_d2h[b].record(self.d2h_stream)
        self.host_out = hosts
        return hosts

    # ---- uint8 in / uint8 out (the loader's tensors in, the images test.py writes out) ---------------------------------------------------
    def enable_u8_io(self, example_u8, crop=None):
        """Allocate the uint8 side of the pipeline: pinned host + device buffers shaped like ``example_u8`` (io_pipeline.U8_KEYS), and uint8
        BGR output buffers for the final image.  After this, ``step_from_host_u8`` is the end-to-end call."""
        from . import io_pipeline
        self._io = io_pipeline
        self.u8_keys = tuple(k for k in io_pipeline.U8_KEYS if k in example_u8)
        self.u8_host = {k: example_u8[k].clone().contiguous().pin_memory() for k in self.u8_keys}
        self.u8_dev = [{k: torch.empty_like(example_u8[k], device=self.device) for k in self.u8_keys} for _ in range(self.depth)]
        self.u8_crop = crop
        with torch.cuda.stream(self.stream):
            self.u8_out_dev = [io_pipeline.images_to_u8(out[1] if len(out) > 1 else out[0], crop) for out in self.slots_out]
        self.stream.synchronize()
        self.u8_out_host = [torch.empty_like(o, device='cpu').pin_memory() for o in self.u8_out_dev]
        self.h2d_bytes_u8 = sum(t.numel() for t in self.u8_host.values())
        self.d2h_bytes_u8 = self.u8_out_host[0].numel()

    def step_from_host_u8(self, host_u8=None):
        """uint8 loader tensors (pinned host) -> device -> normalise / concatenate (one kernel) -> generator -> uint8 BGR photo (one kernel) -> pinned
        host.  Same three-stream software pipeline as ``step_from_host``; returns this step's pinned uint8 images [N, H, Wcrop, 3]."""
        src = self.u8_host if host_u8 is None else host_u8
        b = self._turn % self.depth
        self._turn += 1
        ins, outs = self.slots_in[b], self.slots_out[b]
        with torch.cuda.stream(self.h2d_stream):
            self.h2d_stream.wait_event(self.ev_comp[b])
            for k in self.u8_keys:
                self.u8_dev[b][k].copy_(src[k], non_blocking=True)
            self.ev_h2d[b].record(self.h2d_stream)
        with torch.cuda.stream(self.stream), torch.no_grad():
            self.stream.wait_event(self.ev_h2d[b])
            self.stream.wait_event(self.ev_d2h[b])
            self._io.normalize_u8_batch(self.u8_dev[b], out=ins)
            if self.graphs[b] is not None:
                self.graphs[b].replay()
            else:
                self.static_in = ins
                outs = self._forward()
            self._io.images_to_u8(outs[1] if len(outs) > 1 else outs[0], self.u8_crop, out=self.u8_out_dev[b])
            self.ev_comp[b].record(self.stream)
        with torch.cuda.stream(self.d2h_stream):
            self.d2h_stream.wait_event(self.ev_comp[b])
            self.u8_out_host[b].copy_(self.u8_out_dev[b], non_blocking=True)
            self.ev_d2h[b].record(self.d2h_stream)
        return self.u8_out_host[b]

    def join_streams(self):
        """Make the compute stream wait for every outstanding copy, so an event recorded on it afterwards closes the whole pipeline."""
        for b in range(self.depth):
            self.stream.wait_event(self.ev_h2d[b])
            self.stream.wait_event(self.ev_d2h[b])

    def synchronize(self):
        self.stream.synchronize()
        self.h2d_stream.synchronize()
        self.d2h_stream.synchronize()
