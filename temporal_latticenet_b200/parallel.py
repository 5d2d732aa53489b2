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
    """The data-parallel gradient exchange of the training step: every .grad is a VIEW of one flat fp32
    buffer (autograd accumulates into it in place), so the exchange is an all-reduce of that buffer -- no concatenation,
    no copy back.  ~4.5 M parameters = 18 MB for the gru-gru-aflow-gru cfg.

    Two buckets.  `early` (optional) names the parameters that only the LAST frame of a window uses (the slice head and the
    up-path residual blocks, ~40 % of the bytes): with back-propagation through time their gradients are final as soon as
    the last frame's backward has run, so their all-reduce is issued from a post-accumulate hook on a side stream and runs
    under the backward of the earlier frames.  The rest goes after backward.  __call__() issues what is left and makes the
    current stream wait for both; the result is divided by `divide_by` (world size for the data-parallel mean)."""

    def __init__(self, params, divide_by=None, early=None):
        self.params = [p for p in params if p.requires_grad]
        self.divide_by = divide_by
        early_ids = {id(p) for p in (early or [])}
        # early bucket first, so that each bucket is one contiguous slice of the flat buffer
        self.params.sort(key=lambda p: 0 if id(p) in early_ids else 1)
        self.n_early = sum(p.numel() for p in self.params if id(p) in early_ids)
        self.n = sum(p.numel() for p in self.params)
        self._flat = None
        self._early_left = 0
        self._early_params = [p for p in self.params if id(p) in early_ids]
        self._early_work = None
        self._side = None
        self._hooks = []
        self.last_ms = None

    def _bind(self):
        """(re)creates the flat buffer and points every .grad at its slice; existing gradient values are kept"""
        dev = self.params[0].device
        flat = torch.zeros(self.n, dtype=torch.float32, device=dev)
        o = 0
        for p in self.params:
            k = p.numel()
            view = flat[o:o + k].view_as(p)
            if p.grad is not None:
                view.copy_(p.grad)
            p.grad = view
            o += k
        self._flat = flat
        if self._early_params and not self._hooks and hasattr(torch.Tensor, "register_post_accumulate_grad_hook"):
            for p in self._early_params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_early_grad))
            if dev.type == "cuda":
                self._side = torch.cuda.Stream(device=dev)

    def _bound(self):
        if self._flat is None:
            return False
        o = 0
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() != self._flat.data_ptr() + 4 * o:
                return False
            o += p.numel()
        return True

    def prepare(self):
        """call before backward: binds the gradient views (after set_to_none or a new optimizer) and arms the early bucket"""
        if not self._bound():
            self._bind()
        self._early_left = len(self._early_params) if self._hooks else 0
        self._early_work = None
        return self

    def _on_early_grad(self, _param):
        if self._early_left <= 0:
            return
        self._early_left -= 1
        if self._early_left == 0:
            self._early_work = self._reduce(self._flat[: self.n_early], side=True)

    def _reduce(self, buf, side):
        rank, ws = world()
        div = self.divide_by if self.divide_by is not None else ws
        if buf.is_cuda and side and self._side is not None:
            self._side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._side):
                if ws > 1:
                    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
                if div != 1:
                    buf.div_(div)
            return self._side
        if ws > 1:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        if div != 1:
            buf.div_(div)
        return None

    def __call__(self):
        if not self._bound():          # first step / gradients that were not produced through prepare(): gather them once
            self._bind()
            self._early_work = None
        start = 0
        if self._hooks and self._early_left == 0 and self.n_early > 0:
            start = self.n_early       # the early bucket is on its way (or done)
        elif self._hooks:
            self._early_left = 0       # some early gradient never fired (unused parameter): everything goes now
        timed = self._flat.is_cuda
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        self._reduce(self._flat[start:], side=False)
        if isinstance(self._early_work, torch.cuda.Stream):
            torch.cuda.current_stream().wait_stream(self._early_work)
        self._early_work = None
        if timed:
            e1.record()
            self._events = (e0, e1)
        return self.n

    def elapsed_ms(self):
        """device time of the exposed (post-backward) part of the last exchange"""
        ev = getattr(self, "_events", None)
        if ev is None:
            return None
        ev[1].synchronize()
        return ev[0].elapsed_time(ev[1])
