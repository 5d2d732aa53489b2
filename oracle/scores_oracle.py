"""CPU oracle of the IoU accumulation (callbacks/scores.py:13-47 of the reference), numpy.

TEST INFRASTRUCTURE ONLY (part of oracle/).  PARITY: pinned -- tests/golden/scores.npz holds the outputs of the
reference's OWN `Scores` class (imported unmodified from /root/reference by tests/golden/make_scores_golden.py,
`torchnet` stubbed) on seeded random clouds; tests/test_oracle_cpu.py checks this restatement against it.
"""
import numpy as np


class ScoresOracle:
    def __init__(self):
        self.inter = None
        self.union = None
        self.nr_classes = None

    def accumulate_scores(self, pred_softmax, gt, unlabeled_idx):
        pred_softmax, gt = np.asarray(pred_softmax), np.asarray(gt)
        self.nr_classes = pred_softmax.shape[1]                       # scores.py:14
        pred = pred_softmax.argmax(1)                                 # :15
        if self.inter is None:                                        # :20-22
            self.inter = [0] * self.nr_classes
            self.union = [0] * self.nr_classes
        for l in np.unique(gt):                                       # :18,24  classes present in THIS cloud's gt
            if l == unlabeled_idx:                                    # :26
                continue
            cur = int(((pred == gt) & (gt == l)).sum())               # :27
            self.inter[int(l)] += cur                                 # :28
            self.union[int(l)] += int((gt == l).sum()) + int((pred == l).sum()) - cur   # :29

    def compute_stats(self):
        valid, total, per = 0, 0.0, {}
        for i in range(self.nr_classes):                              # :35-43
            if self.union[i] > 0:
                valid += 1
                iou = self.inter[i] / self.union[i]
                total += iou
                per[i] = iou
        return total / valid, per                                     # :44
