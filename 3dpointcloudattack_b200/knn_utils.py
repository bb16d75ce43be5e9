"""Drop-in for attack/GeoA3/knn_utils.py (the reference's pure-torch stand-in for
pytorch3d.ops.knn_points / knn_gather).

`knn_points` reproduces the reference arithmetic INCLUDING its broadcast quirk
(knn_utils.py:13-15): dist[i,j] = |p1_j|^2 - 2 p1_i.p2_j + |p2_i|^2, which requires
P1 == P2 and equals the true squared distance only for p1 is p2.  lengths1/lengths2/version/
return_sorted are accepted and ignored, as in the reference (:44-50).  idx is int64.
"""
from collections import namedtuple
from typing import Union

import torch

from . import functional as F

_KNN = namedtuple("KNN", "dists idx knn")


def knn_points(p1: torch.Tensor, p2: torch.Tensor,
               lengths1: Union[torch.Tensor, None] = None, lengths2: Union[torch.Tensor, None] = None,
               K: int = 1, version: int = -1, return_nn: bool = False, return_sorted: bool = True) -> _KNN:
    if p1.shape[0] != p2.shape[0]:
        raise ValueError("pts1 and pts2 must have the same batch dimension.")
    if p1.shape[2] != p2.shape[2]:
        raise ValueError("pts1 and pts2 must have the same point dimension.")
    if p1.shape[1] != p2.shape[1]:
        # the reference's `p1_2 + inner + p2_2^T` broadcast fails for P1 != P2
        raise RuntimeError(f"The size of tensor a ({p1.shape[1]}) must match the size of tensor b "
                           f"({p2.shape[1]}) at non-singleton dimension 2")
    if K == 1 and p1.shape[2] == 3:
        r = F.nn1(p1, p2, F.FORM_COL_ROW, F.NORM_MULSUM, swap_norms=True)
        dists, idx = r.row_min.unsqueeze(-1), r.row_arg.long().unsqueeze(-1)
    else:
        dists, idx32 = F.knn(p1, p2, K, form=F.FORM_COL_ROW, norm=F.NORM_MULSUM, swap_norms=True)
        idx = idx32.long()
    p2_nn = knn_gather(p2, idx, lengths2) if return_nn else None
    return _KNN(dists=dists, idx=idx, knn=p2_nn)


def knn_gather(x: torch.Tensor, idx: torch.Tensor, lengths: Union[torch.Tensor, None] = None):
    """x[B,M,U], idx[B,L,K] -> [B,L,K,U] (knn_utils.py:58-86); differentiable w.r.t. x."""
    N, M, U = x.shape
    _N, L, K = idx.shape
    if N != _N:
        raise ValueError("x and idx must have same batch dimension.")
    idx_expanded = idx[:, :, :, None].expand(-1, -1, -1, U)
    return x[:, :, None].expand(-1, -1, K, -1).gather(1, idx_expanded)
