"""Per-class IoU accumulation on the device -- the `Scores` object of the reference's training / evaluation
loop (/root/reference/callbacks/scores.py:7-95, fed by StateCallback.after_forward_pass,
callbacks/state_callback.py:11-16, once per window from train_ln.py:219 / test_ln.py:234).

Same method names and semantics: a cloud contributes to class l only when l occurs in ITS ground truth
(scores.py:18,24), the unlabeled class is skipped, union = |gt == l| + |pred == l| - intersection,
mean IoU over the classes with a non-empty union (scores.py:32-45).

What differs is where it runs: the reference does one `.item()` round trip per present class and cloud
(scores.py:27-30, ~50 stream synchronisations per window); here arg-max + confusion counting is one kernel
(csrc/ltn_reduce.cu::k_confusion) plus a 64-thread fold, nothing synchronises until the statistics are read.
"""
import csv

import torch

from . import _lib


class Scores:
    def __init__(self):
        self.clear()

    def _alloc(self, nr_classes, device):
        if nr_classes > 64:
            raise RuntimeError("Scores supports up to 64 classes")
        self.nr_classes = int(nr_classes)
        self._conf = torch.zeros(nr_classes, nr_classes, dtype=torch.int64, device=device)
        self._inter = torch.zeros(nr_classes, dtype=torch.int64, device=device)
        self._union = torch.zeros(nr_classes, dtype=torch.int64, device=device)

    def accumulate_scores(self, pred_softmax, gt, unlabeled_idx, rows_dev=None):
        """pred_softmax [N,K] float32 (any monotone transform of the class probabilities: the reference passes
        log-softmax, train_ln.py:219), gt [N] int64 -- both CUDA tensors."""
        _lib.require_cuda()
        if pred_softmax.dim() != 2 or gt.dim() != 1 or gt.shape[0] != pred_softmax.shape[0]:
            raise RuntimeError("expected pred [N,K] and gt [N]")
        if self._inter is None or self.nr_classes != pred_softmax.shape[1] or self._inter.device != pred_softmax.device:
            self._alloc(pred_softmax.shape[1], pred_softmax.device)
        pred = pred_softmax.detach().contiguous().float()
        gt = gt.detach().contiguous().to(torch.int64)
        p = _lib.ptr
        _lib.check(_lib.load().ltn_scores_accumulate(p(pred), p(gt), pred.shape[0], rows_dev, self.nr_classes, int(unlabeled_idx),
                                                     p(self._conf), p(self._inter), p(self._union), _lib.stream()),
                   "ltn_scores_accumulate")

    # the reference exposes the two lists as attributes (scores.py:20-22)
    @property
    def intersection_per_class(self):
        return None if self._inter is None else self._inter.tolist()

    @property
    def union_per_class(self):
        return None if self._union is None else self._union.tolist()

    def compute_stats(self, print_per_class_iou=False):
        """one device -> host read of the 2 x K counters"""
        inter, union = self.intersection_per_class, self.union_per_class
        valid, iou_sum, iou_dict = 0, 0.0, {}
        for i in range(self.nr_classes):
            if union[i] > 0:
                valid += 1
                iou = inter[i] / union[i]
                iou_sum += iou
                if print_per_class_iou:
                    print("class iou for idx", i, " is ", iou)
                iou_dict[i] = iou
        return iou_sum / valid, iou_dict   # ZeroDivisionError without any labelled point, like the reference

    def avg_class_iou(self, print_per_class_iou=False):
        return self.compute_stats(print_per_class_iou)[0]

    def iou_per_class(self, print_per_class_iou=False):
        return self.compute_stats(print_per_class_iou)[1]

    def update_best(self):
        avg_iou, iou_dict = self.compute_stats(False)
        if avg_iou > self.best_iou:
            self.best_iou, self.best_iou_dict = avg_iou, iou_dict

    def show(self, epoch_nr):
        self.avg_class_iou(print_per_class_iou=True)

    def clear(self):
        self.start_fresh_eval()
        self.best_iou = -99999999
        self.best_iou_dict = {}

    def start_fresh_eval(self):
        self._conf = self._inter = self._union = None
        self.labels = None
        self.nr_classes = None

    def write_iou_to_csv(self, filename):
        avg_iou, iou_dict = self.compute_stats(False)
        with open(filename, "w") as f:
            w = csv.writer(f)
            for key, val in iou_dict.items():
                w.writerow([key, val])
            w.writerow(["mean_iou", avg_iou])

    def write_best_iou_to_csv(self, filename):
        with open(filename, "w") as f:
            w = csv.writer(f)
            for key, val in self.best_iou_dict.items():
                w.writerow([key, val])
            w.writerow(["best_iou", self.best_iou])
