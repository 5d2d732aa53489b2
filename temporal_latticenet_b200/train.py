"""Training step over one window: the loop body of /root/reference/train_ln.py:160-239 -- 4 frames
forward with BPTT (hidden states are not detached, lattice_modules.py:56-63), loss on the last frame
= 0.5 Lovasz-softmax + 0.5 NLL (train_ln.py:212-216, ignore_index 0), AdamW(amsgrad) (train_ln.py:181)
-- plus the data-parallel gradient all-reduce of SURVEY.md section 8(e).
"""
import torch

from .config import ConfigParser
from .lattice import Lattice, ModelParams
from .lovasz import LovaszSoftmax
from .model import LatticeNetSeq
from .parallel import FlatGradAllReduce, broadcast_parameters
from .scores import Scores


class WindowTrainer:
    def __init__(self, cfg_path, nr_classes=26, device=None, lr=1e-3, weight_decay=1e-3, ignore_index=0):
        self.cfg_path = cfg_path
        self.device = torch.device(device if device is not None else ("cuda:%d" % torch.cuda.current_device()))
        self.model = LatticeNetSeq(nr_classes, ModelParams.create(cfg_path), ConfigParser(cfg_path)).to(self.device)
        self.model.train(True)
        self.lovasz = LovaszSoftmax(ignore_index=ignore_index)
        self.nll = torch.nn.NLLLoss(ignore_index=ignore_index)
        self.lr, self.weight_decay = lr, weight_decay
        self.optimizer = None
        self.allreduce = None
        self.ignore_index = ignore_index
        self.scores = Scores()       # phase.scores of the reference (callbacks/state_callback.py:11-16), accumulated on the device
        self.last_output = None

    def forward_window(self, frames, target):
        self.model.reset_sequence()
        ls = Lattice.create(self.cfg_path, "lattice")
        out = None
        last = len(frames) - 1
        for i, (p, v) in enumerate(frames):
            out, _, ls = self.model(ls, p, v, i != last, True)
        self.last_output = out.detach()
        return 0.5 * self.lovasz(out, target) + 0.5 * self.nll(out, target)

    def materialise(self, frames, target, state_dict_fn=None):
        """first window creates the lazy parameters (train_ln.py:177-191); then the optimizer exists"""
        with torch.no_grad():
            self.forward_window(frames, target)
        if state_dict_fn is not None:
            shapes = {k: tuple(v.shape) for k, v in self.model.state_dict().items()}
            self.model.load_state_dict(state_dict_fn(shapes))
        broadcast_parameters(self.model)
        self.optimizer = torch.optim.AdamW(self.model.parameters(), lr=self.lr, weight_decay=self.weight_decay, amsgrad=True)
        # the parameters only the last frame uses (slice head, up-path residual blocks -- models.py:435-437,465): their
        # gradients are final once the last frame's backward has run, so their exchange overlaps the earlier frames' backward
        early = list(self.model.slice_fast_cuda.parameters()) + list(self.model.resnet_blocks_per_up_lvl_list.parameters())
        self.allreduce = FlatGradAllReduce(self.model.parameters(), early=early)
        return self

    def step(self, frames, target):
        loss = self.forward_window(frames, target)
        self.optimizer.zero_grad(set_to_none=False)
        self.allreduce.prepare()      # gradients are views of one flat buffer; arms the early bucket
        loss.backward()
        self.allreduce()
        self.optimizer.step()
        # cb.after_forward_pass(pred_softmax=..., target=...) of train_ln.py:219: IoU bookkeeping, no host round trip
        self.scores.accumulate_scores(self.last_output, target, self.ignore_index)
        return loss.detach()

    @property
    def allreduce_ms(self):
        return self.allreduce.elapsed_ms() if self.allreduce is not None else None
