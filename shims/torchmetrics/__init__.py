"""torchmetrics.AUROC / Accuracy as recsys/dlrm_main.py:304-305,325-330 uses them: binary scores in [0, 1], integer
labels, accumulate over an epoch (compute_on_step=False), compute() over all ranks."""
import torch
import torch.distributed as dist
import torch.nn as nn


class _Accumulating(nn.Module):
    def __init__(self, compute_on_step: bool = False, **_):
        super().__init__()
        self._preds, self._target = [], []

    def update(self, preds: torch.Tensor, target: torch.Tensor) -> None:
        self._preds.append(preds.detach().reshape(-1).float())
        self._target.append(target.detach().reshape(-1).long())

    def forward(self, preds, target):
        self.update(preds, target)

    def reset(self) -> None:
        self._preds, self._target = [], []

    def _gathered(self):
        preds, target = torch.cat(self._preds), torch.cat(self._target)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            sizes = [torch.zeros(1, dtype=torch.long, device=preds.device) for _ in range(dist.get_world_size())]
            dist.all_gather(sizes, torch.tensor([preds.numel()], device=preds.device))
            longest = int(max(s.item() for s in sizes))
            pad = lambda t: torch.cat([t, t.new_zeros(longest - t.numel())])
            ps = [torch.empty(longest, dtype=preds.dtype, device=preds.device) for _ in sizes]
            ts = [torch.empty(longest, dtype=target.dtype, device=preds.device) for _ in sizes]
            dist.all_gather(ps, pad(preds))
            dist.all_gather(ts, pad(target))
            preds = torch.cat([p[:int(s.item())] for p, s in zip(ps, sizes)])
            target = torch.cat([t[:int(s.item())] for t, s in zip(ts, sizes)])
        return preds, target


class Accuracy(_Accumulating):
    def __init__(self, threshold: float = 0.5, **kwargs):
        super().__init__(**kwargs)
        self.threshold = threshold

    def compute(self) -> torch.Tensor:
        preds, target = self._gathered()
        return ((preds >= self.threshold).long() == target).float().mean()


class AUROC(_Accumulating):
    def compute(self) -> torch.Tensor:
        """Area under the ROC curve = P(score of a positive > score of a negative), ties counted half (rank statistic)."""
        preds, target = self._gathered()
        pos = target == 1
        n_pos, n_neg = int(pos.sum()), int((~pos).sum())
        if n_pos == 0 or n_neg == 0:
            return torch.tensor(0.5, device=preds.device)
        vals, inverse, counts = torch.unique(preds, return_inverse=True, return_counts=True)
        ends = torch.cumsum(counts, 0).double()
        avg_rank = ends - (counts.double() - 1) / 2          # average 1-based rank of each distinct score
        ranks = avg_rank[inverse]
        u = ranks[pos].sum() - n_pos * (n_pos + 1) / 2
        return (u / (n_pos * n_neg)).float()
