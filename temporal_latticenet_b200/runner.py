"""Window runner: the loop the reference's drivers run around the model (test_ln.py:149-231,
train_ln.py:160-239) reduced to what the hot path needs -- a fresh lattice per window, frames fed in
order with early_return on all but the last, arg-max labels of the last frame back on the host.

This is the public entry point `bench.py` times end to end (host buffers in, host labels out).
"""
import torch

from . import ops
from .config import ConfigParser
from .lattice import Lattice, ModelParams
from .model import LatticeNetSeq


class WindowRunner:
    def __init__(self, cfg_path, nr_classes=26, device=None, operands="f16"):
        self.cfg_path = cfg_path
        self.device = torch.device(device if device is not None else ("cuda:%d" % torch.cuda.current_device()))
        self.model = LatticeNetSeq(nr_classes, ModelParams.create(cfg_path), ConfigParser(cfg_path)).to(self.device)
        self.model.train(False)
        self.lattice = None
        self._labels_host = None
        # operand type of the fp32-parity tensor-core convolutions: fp16 hi/lo by default (ops.tc_operands), with the
        # device-side range flag checked after every window and a tf32 re-run when it was raised
        self.operands = operands
        self.range_flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.range_fallbacks = 0

    def new_lattice(self):
        """a fresh lattice per window (train_ln.py:236-239, test_ln.py:264)"""
        self.model.reset_sequence()
        self.lattice = Lattice.create(self.cfg_path, "lattice")
        return self.lattice

    def materialise_parameters(self, frames_dev, state_dict_fn=None):
        """Parameters are created lazily by the first full window (test_ln.py:165-185); then an
        optional name->tensor function provides the weights (the pretrained checkpoint is missing
        from the reference mount, so tests and the bench seed them by name)."""
        with torch.no_grad():
            self.infer_window_device(frames_dev)
        if state_dict_fn is not None:
            shapes = {k: tuple(v.shape) for k, v in self.model.state_dict().items()}
            self.model.load_state_dict(state_dict_fn(shapes))
        return self

    def infer_window_device(self, frames_dev):
        """frames_dev: list of (positions [N,3], values [N,1]) CUDA tensors.  Returns the final
        frame's log-softmax [N, classes] (device)."""
        for operands in ((self.operands, "tf32") if self.operands != "tf32" else ("tf32",)):
            ls = self.new_lattice()
            out = None
            last = len(frames_dev) - 1
            with torch.no_grad(), ops.tc_operands(operands, self.range_flag):
                for i, (p, v) in enumerate(frames_dev):
                    out, _, ls = self.model(ls, p, v, i != last, False)
            self.lattice = ls
            if operands == "tf32" or not self.range_raised():
                break
            self.range_fallbacks += 1     # an activation left the fp16 range: the tf32 operands redo the window
        return out

    def range_raised(self):
        """reads (and clears) the fp16 range flag; synchronises the current stream"""
        raised = bool(int(self.range_flag.item()))
        if raised:
            self.range_flag.zero_()
        return raised

    def evaluate_window(self, frames_dev, target, scores, unlabeled_idx=0):
        """test_ln.py:165-234 for one window: predict the last frame and fold it into `scores` (scores.Scores, the
        reference's phase.scores) on the device; returns the log-softmax"""
        out = self.infer_window_device(frames_dev)
        scores.accumulate_scores(out, target, unlabeled_idx)
        return out

    def infer_window(self, frames_host):
        """frames_host: list of (positions, values) PINNED host tensors.  Returns predicted labels of
        the last frame as a pinned host int64 tensor (test_ln.py:219-222), synchronised."""
        dev = self.device
        frames = [(p.to(dev, non_blocking=True), v.to(dev, non_blocking=True)) for p, v in frames_host]
        out = self.infer_window_device(frames)
        labels = out.argmax(1)
        if self._labels_host is None or self._labels_host.shape[0] < labels.shape[0]:
            self._labels_host = torch.empty(labels.shape[0], dtype=torch.int64).pin_memory()
        host = self._labels_host[: labels.shape[0]]
        host.copy_(labels, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return host
