"""Multi-GPU plumbing of the hot path: one process per GPU, independent 4-frame windows sharded over
the ranks (each window owns a fresh lattice -- train_ln.py:236-239 -- so inference needs no data-path
collective) and, for training only, ONE all-reduce of the flattened gradients per optimizer step.

The reference is single-GPU (no torch.distributed call site anywhere); this is
the data-parallel shape SURVEY.md section 8(e) defines.  Backend-agnostic: NCCL on the GPU box, gloo in the
CPU tests (tests/test_distributed_cpu.py, world_size 2).
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_windows(nr_windows, rank=None, world_size=None):
    """window w -> rank w mod world_size (round robin keeps the load even when window sizes drift)"""
    if rank is None or world_size is None:
        rank, world_size = world()
    return list(range(rank, nr_windows, world_size))


def broadcast_parameters(module, src=0):
    """Parameters are created lazily by the first window (train_ln.py:177-191), so ranks first run one
    window each and then take rank `src`'s values, tensor by tensor in state-dict order."""
    rank, ws = world()
    if ws == 1:
        return
    for _, t in sorted(module.state_dict().items()):
        dist.broadcast(t, src)


class FlatGradAllReduce:
    """Flattens every .grad into one contiguous buffer, all-reduces it once (SUM) and scatters the
    result back divided by `divide_by` (world size for the data-parallel mean).  ~4.5 M parameters =
    18 MB for the gru-gru-aflow-gru cfg: one latency-bound NVLink collective per step."""

    def __init__(self, params, divide_by=None):
        self.params = [p for p in params if p.requires_grad]
        self.divide_by = divide_by
        self._flat = None

    def __call__(self):
        rank, ws = world()
        grads = [p.grad for p in self.params if p.grad is not None]
        if not grads:
            return 0
        n = sum(g.numel() for g in grads)
        if self._flat is None or self._flat.numel() != n or self._flat.device != grads[0].device:
            self._flat = torch.empty(n, dtype=grads[0].dtype, device=grads[0].device)
        torch.cat([g.reshape(-1) for g in grads], out=self._flat)
        if ws > 1:
            dist.all_reduce(self._flat, op=dist.ReduceOp.SUM)
        div = self.divide_by if self.divide_by is not None else ws
        if div != 1:
            self._flat.div_(div)
        o = 0
        for g in grads:
            k = g.numel()
            g.copy_(self._flat[o:o + k].view_as(g))
            o += k
        return n
