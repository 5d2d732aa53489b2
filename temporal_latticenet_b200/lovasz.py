"""Lovasz-softmax loss on log-probabilities (train_ln.py:119,214: `LovaszSoftmax(ignore_index)`
applied to `pred_logsoftmax`).  Published algorithm: Berman, Triki, Blaschko, CVPR 2018."""
import torch


class LovaszSoftmax(torch.nn.Module):
    def __init__(self, ignore_index=None):
        super().__init__()
        self.ignore_index = ignore_index

    def forward(self, logsoftmax, target):
        p = logsoftmax.exp()
        if self.ignore_index is not None:
            keep = target != self.ignore_index
            p, target = p[keep], target[keep]
        n, k = p.shape
        if n == 0:
            return p.sum() * 0.0
        fg = torch.nn.functional.one_hot(target, k).to(p.dtype)          # [n,k]
        present = fg.sum(0) > 0
        err = (fg - p).abs()
        err_sorted, perm = torch.sort(err, 0, descending=True)           # all classes at once
        fg_sorted = fg.gather(0, perm)
        total = fg_sorted.sum(0, keepdim=True)
        inter = total - fg_sorted.cumsum(0)
        union = total + (1.0 - fg_sorted).cumsum(0)
        jac = 1.0 - inter / union
        jac = torch.cat([jac[:1], jac[1:] - jac[:-1]], 0)
        per_class = (err_sorted * jac).sum(0)
        if not bool(present.any()):
            return p.sum() * 0.0
        return per_class[present].mean()
