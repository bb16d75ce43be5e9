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


def farthest_point_sample(xyz, npoint):
    """:59-81 -> centroids [B,npoint] int64.  The start index is drawn exactly as the reference
    draws it (torch.randint on the CPU generator, :71), so a seeded run picks the same points;
    the npoint-iteration Python loop (6 launches per iteration) is one persistent kernel."""
    B, N, _ = xyz.shape
    start = torch.randint(0, N, (B,), dtype=torch.long)
    return F.farthest_point_sample(xyz, npoint, start).long()


def index_points(points, idx):
    """:41-57: points[B,N,C], idx[B,S] or [B,S,K] -> points[b, idx[b,...], :] (differentiable
    gather; plain torch indexing, one kernel either way)."""
    B = points.shape[0]
    view_shape = [B] + [1] * (idx.dim() - 1)
    batch_indices = torch.arange(B, dtype=torch.long, device=points.device).view(view_shape)
    return points[batch_indices, idx.long(), :]


def sample_and_group(npoint, radius, nsample, xyz, points, returnfps=False):
    """:107-135 with the FPS and ball-query kernels."""
    B, N, C = xyz.shape
    S = npoint
    fps_idx = farthest_point_sample(xyz, npoint)
    new_xyz = index_points(xyz, fps_idx)
    idx = query_ball_point(radius, nsample, xyz, new_xyz)
    grouped_xyz = index_points(xyz, idx)
    grouped_xyz_norm = grouped_xyz - new_xyz.view(B, S, 1, C)
    if points is not None:
        new_points = torch.cat([grouped_xyz_norm, index_points(points, idx)], dim=-1)
    else:
        new_points = grouped_xyz_norm
    if returnfps:
        return new_xyz, new_points, grouped_xyz, fps_idx
    return new_xyz, new_points
