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
    """model/dgcnn.py:203-227 -> [B, 2C, N, k] = cat(feature - x, x) permuted channel-first.
    The reference's flat gather + repeat + cat + permute().contiguous() chain (four
    [B,N,k,2C]-sized intermediates) is one gather kernel writing the output once; its
    backward is one kernel with shared-memory accumulators.  The hard-coded `cuda:0` (:209)
    becomes x.device.  Like the reference, `k` must equal idx.shape[2] when idx is given."""
    batch_size = x.size(0)
    num_points = x.size(2)
    x = x.view(batch_size, -1, num_points)
    if idx is None:
        pts = x.detach().transpose(2, 1)
        _, idx = F.knn(pts, pts, k, form=F.FORM_COL_ROW, norm=F.NORM_MULSUM)      # int32, no .long() round trip
    elif idx.shape[2] != k:
        raise RuntimeError(f"shape mismatch: idx has {idx.shape[2]} neighbours, k={k}")   # the reference's view() raises here
    return F.edge_feature(x, idx, (F.EDGE_DIFF, F.EDGE_CENTER))
