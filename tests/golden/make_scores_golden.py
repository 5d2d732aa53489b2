"""Generates tests/golden/scores.npz with the REFERENCE's own callbacks/scores.py::Scores (imported unmodified from
/root/reference; its `import torchnet` is satisfied by an empty stub module -- the class never uses it).
Run in the build container only:  python tests/golden/make_scores_golden.py"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.modules.setdefault("torchnet", types.ModuleType("torchnet"))
sys.path.insert(0, "/root/reference")
from callbacks.scores import Scores  # noqa: E402  (reference file, unmodified)


def clouds(seed, nr_clouds, nr_classes, n):
    rng = np.random.default_rng(seed)
    out = []
    for c in range(nr_clouds):
        present = rng.choice(nr_classes, size=rng.integers(3, nr_classes), replace=False)   # not every class in every cloud
        gt = present[rng.integers(0, len(present), n)]
        logits = rng.standard_normal((n, nr_classes)).astype(np.float32)
        logits[np.arange(n), gt] += 1.5            # a classifier that is right about half of the time
        out.append((logits, gt.astype(np.int64)))
    return out


if __name__ == "__main__":
    arrays = {}
    for case, (seed, nr_clouds, K, n, unl) in enumerate([(0, 4, 26, 900, 0), (1, 3, 20, 700, 0), (2, 4, 7, 300, 3)]):
        s = Scores()
        for i, (logits, gt) in enumerate(clouds(seed, nr_clouds, K, n)):
            s.accumulate_scores(torch.from_numpy(logits), torch.from_numpy(gt), unl)
            arrays["c%d_logits%d" % (case, i)], arrays["c%d_gt%d" % (case, i)] = logits, gt
        avg, per = s.compute_stats()
        arrays["c%d_meta" % case] = np.array([nr_clouds, K, unl], np.int64)
        arrays["c%d_inter" % case] = np.array([int(x) for x in s.intersection_per_class], np.int64)
        arrays["c%d_union" % case] = np.array([int(x) for x in s.union_per_class], np.int64)
        arrays["c%d_avg" % case] = np.float64(avg)
        arrays["c%d_per" % case] = np.array([per.get(i, -1.0) for i in range(K)], np.float64)
    np.savez_compressed(os.path.join(HERE, "scores.npz"), **arrays)
    print("wrote scores.npz", {k: v.shape for k, v in arrays.items() if "meta" in k})
