"""Synthetic Bosphorus-shaped face clouds (benchmark / test inputs).

Modelled on AddData/face0424.txt of the reference: a 2.5-D height field (nose bump, face dome,
eye sockets) sampled on a jittered grid inside an ellipse, random row permutation
(readbnt.py:21-25), centred and scaled to unit max norm (pointnet/bosphorus_dataset.py:74-76).
Deterministic per (seed, sample index); tie-free by construction (continuous jitter).
"""
import math

import torch


def face_cloud(n, seed):
    g = torch.Generator().manual_seed(int(seed))
    side = int(math.ceil(math.sqrt(n * 4.0 / math.pi * 1.15))) + 2     # grid that over-fills the ellipse
    pitch_x, pitch_y = 1.4 / side, 1.6 / side
    ix, iy = torch.meshgrid(torch.arange(side), torch.arange(side), indexing="ij")
    x = (ix.reshape(-1).double() + 0.5) * pitch_x - 0.7 + (torch.rand(side * side, generator=g, dtype=torch.float64) * 0.6 - 0.3) * pitch_x
    y = (iy.reshape(-1).double() + 0.5) * pitch_y - 0.8 + (torch.rand(side * side, generator=g, dtype=torch.float64) * 0.6 - 0.3) * pitch_y
    r2 = (x / 0.7) ** 2 + (y / 0.8) ** 2
    keep = torch.argsort(r2)[:n]                                        # the n points closest to the centre
    if keep.numel() < n:
        raise ValueError("grid too small")
    x, y = x[keep], y[keep]
    z = (0.35 * torch.exp(-(x ** 2 + (y + 0.05) ** 2) / 0.02)
         + 0.25 * torch.sqrt(torch.clamp(1 - (x / 0.8) ** 2 - (y / 0.9) ** 2, min=0))
         - 0.05 * torch.exp(-((x.abs() - 0.25) ** 2 + (y - 0.2) ** 2) / 0.005)
         + 0.002 * torch.randn(n, generator=g, dtype=torch.float64))
    pts = torch.stack([x, y, z], 1)
    pts = pts[torch.randperm(n, generator=g)]
    pts = pts - pts.mean(0, keepdim=True)
    pts = pts / pts.norm(dim=1).max()
    return pts.float()


def face_clouds(B, n, seed=1234, first_sample=0):
    """[B, n, 3] fp32 (CPU); sample b uses seed + first_sample + b, so shards of a batch are
    independent of how the batch is split over GPUs."""
    return torch.stack([face_cloud(n, seed + first_sample + b) for b in range(B)])


def perturb(ori, sigma, seed=0, budget=0.18, first_sample=0):
    """adv = ori + N(0, sigma^2), clipped per point to L2 <= budget (Eval_CW.py:89, ClipPointsLinf)."""
    out = []
    for b in range(ori.shape[0]):
        g = torch.Generator().manual_seed(int(seed) * 1000003 + first_sample + b)
        d = torch.randn(ori.shape[1:], generator=g) * sigma
        nrm = d.norm(dim=-1, keepdim=True)
        d = d * torch.clamp(budget / (nrm + 1e-9), max=1.0)
        out.append(ori[b] + d)
    return torch.stack(out)
