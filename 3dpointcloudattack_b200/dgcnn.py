"""Drop-in for the kNN-graph helpers of model/dgcnn.py:194-227 (dups: pointnet/model.py:236-269,
attack/AOF/TAOF_attack.py:13-28, attack/AOF/Eval_AOF.py:46-62)."""
import torch

from . import functional as F


def knn(x, k):
    """x[B,C,N] (C = 3 / 64 / 128 in DGCNN) -> idx[B,N,k] int64, nearest first, self included.
    The reference's pairwise = (-xx - inner) - xx^T is the exact negation of FORM_COL_ROW, so its
    top-k largest are the k smallest here.  No gradient (index output)."""
    pts = x.detach().transpose(2, 1)                 # [B,N,C] view
    _, idx = F.knn(pts, pts, k, form=F.FORM_COL_ROW, norm=F.NORM_MULSUM)
    return idx.long()


def get_graph_feature(x, k=20, idx=None):
    """model/dgcnn.py:203-227 -> [B, 2C, N, k] = cat(feature - x, x); the hard-coded `cuda:0`
    (:209) becomes x.device."""
    batch_size = x.size(0)
    num_points = x.size(2)
    x = x.view(batch_size, -1, num_points)
    if idx is None:
        idx = knn(x, k=k)
    idx_base = torch.arange(0, batch_size, device=x.device).view(-1, 1, 1) * num_points
    idx = (idx + idx_base).view(-1)
    _, num_dims, _ = x.size()
    x = x.transpose(2, 1).contiguous()
    feature = x.view(batch_size * num_points, -1)[idx, :]
    feature = feature.view(batch_size, num_points, k, num_dims)
    x = x.view(batch_size, num_points, 1, num_dims).repeat(1, 1, k, 1)
    return torch.cat((feature - x, x), dim=3).permute(0, 3, 1, 2).contiguous()
