"""Drop-in for the local-geometry helpers of attack/GeoA3/utility.py (clouds are [b,3,n]).

    _normalize                       utility.py:33-34
    estimate_normal                  utility.py:43-92
    estimate_normal_via_ori_normal   utility.py:94-111
    get_perpendicular_jitter         utility.py:113-117
    estimate_perpendicular           utility.py:119-152
    jitter_input                     utility.py:36-41

The reference's `estimate_normal` / `estimate_perpendicular` call `torch.symeig`, which current
torch no longer has; here the k-NN select and the per-point covariance eigen-frame are two kernels
(functional.knn + functional.local_frames), batch-parallel (the reference loops over the batch in
Python) and without the [b,3,n,k] neighbour gather.  Sign of the normal: the reference multiplies
by -sign(<n, sum of the centred neighbours>), and that sum is rounding noise around zero -- the
sign is arbitrary there and here; every consumer (kappa = |<v,n>|, offset_proj) is sign-free.
The hard-coded `cuda:6` device (utility.py:29) becomes the input's device.
"""
import torch

from . import functional as F
from .knn_utils import knn_gather, knn_points


def _normalize(input, p=2, dim=1, eps=1e-12):
    return input / input.norm(p, dim, keepdim=True).clamp(min=eps).expand_as(input)


def jitter_input(data, sigma=0.01, clip=0.05):
    assert data.size(1) == 3
    assert clip > 0
    B, _, N = data.size()
    return torch.clamp(sigma * torch.randn(B, 3, N), -1 * clip, clip).to(data.device)


def _self_knn_idx(pc, K):
    pts = pc.detach().permute(0, 2, 1)
    _, idx = F.knn(pts, pts, K, form=F.FORM_COL_ROW, norm=F.NORM_MULSUM, swap_norms=True)   # knn_points arithmetic
    return pts, idx


def estimate_normal(pc, k):
    """pc [b,3,n] -> unit normals [b,3,n] (no gradient): eigenvector of the smallest eigenvalue of the
    covariance of the k nearest neighbours (self excluded)."""
    with torch.no_grad():
        pts, idx = _self_knn_idx(pc, k + 1)
        normal, _, _ = F.local_frames(pts, idx, skip_first=True, normals=True, frames=False)
    return normal.permute(0, 2, 1).contiguous().float()


def estimate_normal_via_ori_normal(pc_adv, pc_ori, normal_ori, k):
    """utility.py:94-111 -- mean of the normals of the k nearest original points (the nearest one's normal
    where the point has not moved)."""
    intra_KNN = knn_points(pc_adv.permute(0, 2, 1), pc_ori.permute(0, 2, 1), K=k)
    inter_value = intra_KNN.dists[:, :, 0].contiguous()
    normal_pts = knn_gather(normal_ori.permute(0, 2, 1), intra_KNN.idx).permute(0, 3, 1, 2).contiguous()
    normal_pts_avg = normal_pts.mean(dim=-1)
    normal_pts_avg = normal_pts_avg / (normal_pts_avg.norm(dim=1) + 1e-12)
    normal_ori_select = normal_pts[:, :, :, 0]
    condition = (inter_value < 1e-6).unsqueeze(1).expand_as(normal_ori_select)
    return torch.where(condition, normal_ori_select, normal_pts_avg)


def get_perpendicular_jitter(vector, sigma=0.01, clip=0.05):
    b, _, n = vector.size()
    aux_vector1 = sigma * torch.randn(b, 3, n).to(vector.device)
    aux_vector2 = sigma * torch.randn(b, 3, n).to(vector.device)
    return torch.clamp(torch.cross(vector, aux_vector1, dim=1), -1 * clip, clip) + \
        torch.clamp(torch.cross(vector, aux_vector2, dim=1), -1 * clip, clip)


def estimate_perpendicular(pc, k, sigma=0.01, clip=0.05):
    """utility.py:119-152 -- random jitter inside the local tangent plane: the two eigenvectors of the larger
    eigenvalues, each scaled by its own N(0, sigma^2) draw and clipped."""
    with torch.no_grad():
        b, _, n = pc.size()
        pts, idx = _self_knn_idx(pc, k + 1)
        _, evecs, _ = F.local_frames(pts, idx, skip_first=True, normals=False, frames=True)
        perpendi_vector_1 = evecs[:, :, 2, :].permute(0, 2, 1)        # largest eigenvalue
        perpendi_vector_2 = evecs[:, :, 1, :].permute(0, 2, 1)
        aux_vector1 = sigma * torch.randn(b, n).unsqueeze(1).to(pc.device)
        aux_vector2 = sigma * torch.randn(b, n).unsqueeze(1).to(pc.device)
    return torch.clamp(perpendi_vector_1 * aux_vector1, -1 * clip, clip) + torch.clamp(perpendi_vector_2 * aux_vector2, -1 * clip, clip)
