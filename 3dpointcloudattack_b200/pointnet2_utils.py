"""Drop-in for model/pointnet2_utils.py:19-38, 84-104 (copies: pointnet/pointnet2_utils.py,
model/curvenet_util.py:38-113, attack/SIadv/baselines/defense/DUP_Net/pu_utils.py:7-96)."""
import torch

from . import functional as F


def square_distance(src, dst):
    """:19-38 -- materialising compatibility API ([B,N,M] matrix, plain torch, same op order).
    query_ball_point below does not call it."""
    B, N, _ = src.shape
    _, M, _ = dst.shape
    dist = -2 * torch.matmul(src, dst.permute(0, 2, 1))
    dist += torch.sum(src ** 2, -1).view(B, N, 1)
    dist += torch.sum(dst ** 2, -1).view(B, 1, M)
    return dist


def query_ball_point(radius, nsample, xyz, new_xyz):
    """:84-104 -> group_idx [B,S,nsample] int64: the first nsample in-radius indices in ascending
    index order, padded with the first hit (N when a row has none).  The reference builds a
    [B,S,N] int64 tensor and fully sorts every row; here one warp per query sweeps the
    columns in order and stops after nsample hits."""
    return F.ball_query(radius, nsample, xyz, new_xyz).long()
