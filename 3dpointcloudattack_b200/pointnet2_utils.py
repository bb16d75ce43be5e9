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


def three_nn_interpolate(xyz1, xyz2, points2):
    """The 3-NN inverse-distance interpolation inside PointNetFeaturePropagation.forward (:289-300):
    xyz1 [B,N,3] (targets), xyz2 [B,S,3] (sources), points2 [B,S,D] -> [B,N,D].  The reference builds
    the [B,N,S] matrix with square_distance and fully sorts every row to take three columns; here the
    k-NN select kernel returns the three smallest (same arithmetic: FORM_ROW_COL, ascending,
    lowest index on ties) and their gradient w.r.t. both clouds."""
    B, N, _ = xyz1.shape
    S = xyz2.shape[1]
    if S == 1:
        return points2.repeat(1, N, 1)
    dists, idx = F.knn(xyz1, xyz2, 3, form=F.FORM_ROW_COL, norm=F.NORM_MULSUM)
    dist_recip = 1.0 / (dists + 1e-8)
    norm = torch.sum(dist_recip, dim=2, keepdim=True)
    weight = dist_recip / norm
    return torch.sum(index_points(points2, idx) * weight.view(B, N, 3, 1), dim=2)


def feature_propagation_forward(self, xyz1, xyz2, points1, points2):
    """Replacement body for PointNetFeaturePropagation.forward (:273-311); `self` is the reference's
    module (mlp_convs / mlp_bns are read from it)."""
    import torch.nn.functional as nnF
    xyz1 = xyz1.permute(0, 2, 1)
    xyz2 = xyz2.permute(0, 2, 1)
    points2 = points2.permute(0, 2, 1)
    interpolated = three_nn_interpolate(xyz1, xyz2, points2)
    if points1 is not None:
        new_points = torch.cat([points1.permute(0, 2, 1), interpolated], dim=-1)
    else:
        new_points = interpolated
    new_points = new_points.permute(0, 2, 1)
    for i, conv in enumerate(self.mlp_convs):
        new_points = nnF.relu(self.mlp_bns[i](conv(new_points)))
    return new_points
