"""CPU oracle of `latticenet_py.lattice.lovasz_loss.LovaszSoftmax` (train_ln.py:17,119,214).

TEST INFRASTRUCTURE ONLY (part of oracle/).  Restates the published Lovasz-softmax loss (Berman,
Triki, Blaschko, CVPR 2018): per present class, sort the absolute errors, weight them with the
discrete gradient of the Jaccard index, average over present classes.  Input is log-softmax
(train_ln.py:214 passes `pred_logsoftmax`).
"""
import torch


def jaccard_gradient(gt_sorted):
    n = gt_sorted.numel()
    total = gt_sorted.sum()
    inter = total - gt_sorted.cumsum(0)
    union = total + (1.0 - gt_sorted).cumsum(0)
    jac = 1.0 - inter / union
    if n > 1:
        jac = torch.cat([jac[:1], jac[1:] - jac[:-1]])
    return jac


class LovaszSoftmax(torch.nn.Module):
    def __init__(self, ignore_index=None):
        super().__init__()
        self.ignore_index = ignore_index

    def forward(self, logsoftmax, target):
        probs = logsoftmax.exp()
        if self.ignore_index is not None:
            keep = target != self.ignore_index
            probs, target = probs[keep], target[keep]
        if probs.numel() == 0:
            return probs.sum() * 0.0
        losses = []
        for c in range(probs.shape[1]):
            fg = (target == c).to(probs.dtype)
            if fg.sum() == 0:
                continue
            err = (fg - probs[:, c]).abs()
            err_sorted, perm = torch.sort(err, 0, descending=True)
            losses.append(torch.dot(err_sorted, jaccard_gradient(fg[perm])))
        if not losses:
            return probs.sum() * 0.0
        return torch.stack(losses).mean()
