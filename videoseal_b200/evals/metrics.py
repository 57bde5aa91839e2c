"""Quality / robustness metrics the reference's reports are built from (videoseal/evals/metrics.py), for tensors on any
device.  Only the metrics of the embed/detect slice are here: PSNR (:22-36), L-inf (:56-64), bit accuracy (:150-178),
p-value (:104-121), capacity (:123-148).  SSIM / VMAF need packages that are not part of this path."""
from __future__ import annotations

import math

import torch


def psnr(x: torch.Tensor, y: torch.Tensor, is_video: bool = False) -> torch.Tensor:
    """x, y in ~[0,1], [..., C, H, W]; per image, or over the whole clip when is_video"""
    delta = (255 * (x - y)).reshape(-1, x.shape[-3], x.shape[-2], x.shape[-1])
    dims = (0, 1, 2, 3) if is_video else (1, 2, 3)
    return 20 * math.log10(255.0) - 10 * torch.log10(torch.mean(delta ** 2, dim=dims))


def linf(x: torch.Tensor, y: torch.Tensor, data_range: float = 1.0) -> torch.Tensor:
    """largest absolute pixel difference, on the 0..255 scale"""
    mult = 255.0 / data_range
    return torch.max(torch.abs(mult * (x - y)))


def bit_accuracy(preds: torch.Tensor, targets: torch.Tensor, mask: torch.Tensor = None, threshold: float = 0.0) -> torch.Tensor:
    """preds: bit logits [B, K] (or pixel-wise [B, K, H, W], reduced by majority over the unmasked pixels); targets [B, K] in {0,1}.
    NOTE: `detect()['preds']` is [B, 1+K]; column 0 is the detection logit, pass preds[:, 1:]."""
    p = preds > threshold
    if p.dim() == 4:
        bsz, nbits = p.shape[:2]
        if mask is not None:
            m = mask.expand_as(p).bool()
            p = p.masked_select(m).view(bsz, nbits, -1).mean(dim=-1, dtype=float)
        else:
            p = p.mean(dim=(-2, -1), dtype=float)
        p = p > 0.5
    return (p == (targets > 0.5)).float().mean(dim=-1)


def pvalue(preds: torch.Tensor, targets: torch.Tensor, mask: torch.Tensor = None, threshold: float = 0.0) -> torch.Tensor:
    """one-sided binomial test of the number of matching bits against chance (p = 0.5)"""
    from scipy import stats
    nbits = targets.shape[-1]
    accs = bit_accuracy(preds, targets, mask, threshold)
    return torch.tensor([stats.binomtest(int(a * nbits), nbits, 0.5, alternative="greater").pvalue for a in accs])


def _plogp(p: torch.Tensor) -> torch.Tensor:
    out = p * torch.log2(p)
    out[p == 0] = 0
    return out


def capacity(preds: torch.Tensor, targets: torch.Tensor, mask: torch.Tensor = None, threshold: float = 0.0) -> torch.Tensor:
    """bits through a binary symmetric channel whose error rate is 1 - bit accuracy"""
    nbits = targets.shape[-1]
    a = bit_accuracy(preds, targets, mask, threshold)
    return nbits * (1 + _plogp(a) + _plogp(1 - a))
