"""Data-parallel plumbing for the operator path on one 8 x B200 node: one process per GPU, torch.distributed (NCCL over
NVLink 5 / NVSwitch on the GPU box, gloo in the CPU tests).

* Inference is batch-sharded replicas: every rank holds the full generator and takes samples ``rank::world`` — no collective on
  the data path.
* Training is data parallel like the reference (training_loop_wo_flow_fullbody.py:315-324), but instead of five
  DistributedDataParallel wrappers with their own 25 MB bucket sets, every trainable parameter's ``.grad`` is a VIEW into one flat
  fp32 buffer per network, and a phase ends with ONE all-reduce of that buffer (183 MB for G, 107 MB for D: ~0.3-0.5 ms at the
  ~700 GB/s bus bandwidth measured on this pod) — NVSwitch makes collective cost latency-bound, so fewer, larger messages win.
  Parameters that receive no gradient in a phase (the reference needs find_unused_parameters=True for them) simply contribute zeros.
"""
import torch
import torch.distributed as dist


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def shard_indices(global_batch, world, rank):
    """Samples owned by ``rank``: rank, rank + world, ...  (a partition of range(global_batch))."""
    assert 0 <= rank < world
    return list(range(rank, global_batch, world))


def shard_batch(batch, world, rank):
    """Slice every tensor of a dict-batch along dim 0 for this rank."""
    if world == 1:
        return batch
    n = next(iter(batch.values())).shape[0]
    idx = torch.as_tensor(shard_indices(n, world, rank), dtype=torch.long)
    return {k: v.index_select(0, idx.to(v.device)) for k, v in batch.items()}


class FlatGradBucket:
    """All gradients of a parameter list live in one contiguous buffer; ``allreduce()`` averages it across ranks in one collective."""

    def __init__(self, params, dtype=torch.float32):
        self.params = [p for p in params if p.requires_grad]
        assert self.params, 'no trainable parameters'
        device = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=dtype, device=device)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n
        self.nbytes = self.flat.numel() * self.flat.element_size()

    def zero(self):
        self.flat.zero_()
        for p, v in zip(self.params, self._views()):      # re-attach in case something replaced .grad
            if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                p.grad = v

    def _views(self):
        off = 0
        for p in self.params:
            n = p.numel()
            yield self.flat[off:off + n].view_as(p)
            off += n

    def allreduce(self, group=None, async_op=False):
        """Average over the process group.  Returns the work handle when ``async_op``."""
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        if world == 1:
            return None
        self.flat.div_(world)
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)

    def sanitize(self):
        """The reference's NaN guard before every optimizer step (training_loop_wo_flow_fullbody.py:513-515)."""
        torch.nan_to_num(self.flat, nan=0.0, posinf=1e5, neginf=-1e5, out=self.flat)


def broadcast_parameters(module, src=0, group=None):
    """Make every replica start from rank ``src``'s weights (DDP's initial broadcast)."""
    world, _ = world_info()
    if world == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)
