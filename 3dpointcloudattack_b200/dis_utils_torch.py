"""Drop-in for utils/dis_utils_torch.py (and its copy attack/CTA/utils/dis_utils_torch.py).

Same names, argument order, layouts ([B,3,N] channel-first) and quirks as the reference:
  * `chamfer` divides by a.shape[1] / b.shape[1] (= 3 for [B,3,N] input, not N) and returns
    sample 0 only (utils/dis_utils_torch.py:14-16);
  * `sgd_hausdorff_dis` / `bid_hausdorff_dis` look at sample 0 only (:19-28).
The distances are torch.cdist's mm-path arithmetic (x1_=[-2x,|x|^2,1] @ x2_=[y,1,|y|^2]^T,
clamp_min(0).sqrt()) evaluated by the sm_100a NN-1 sweep -- the [B,N,M] matrix is never built.

Gradient note: the reference's full-reduce torch.max splits the gradient evenly between
exactly tied maxima; this implementation routes it to the first one (lowest index).
"""
import torch

from . import functional as F


def euclidean_distances(a: torch.Tensor, b: torch.Tensor, p=2):
    """utils/dis_utils_torch.py:4-5 -- materialising compatibility path (plain torch; only
    referenced from commented-out code in the reference, attack/CTA/CTA.py:169)."""
    return torch.sum(torch.diagonal(torch.cdist(a, b, p=2)))


def pairwise_distances(a: torch.Tensor, b: torch.Tensor, p=2):
    """utils/dis_utils_torch.py:8-11 -- materialising compatibility API ([B,N,M] matrix). The
    loss functions below do not call it."""
    a = a.permute(0, 2, 1)
    b = b.permute(0, 2, 1)
    return torch.cdist(a, b, p=2)


def _sweep(a, b):
    # rows = points of a, cols = points of b; sample 0 is all the reference ever reads
    return F.nn1(a[:1].permute(0, 2, 1), b[:1].permute(0, 2, 1), F.FORM_ROW_COL, F.NORM_MULSUM,
                 swap_norms=False, transform=F.VALUE_SQRT_CLAMP,
                 row_sum_scale=1.0 / b.shape[1], col_sum_scale=1.0 / a.shape[1])


def chamfer(a, b):
    """utils/dis_utils_torch.py:14-16: (M.min(1)[0].sum(1))/a.shape[1] + (M.min(2)[0].sum(1))/b.shape[1], [0]."""
    r = _sweep(a, b)          # the divisors a.shape[1], b.shape[1] are folded into the kernel's sums
    return (r.col_sum + r.row_sum)[0]


def sgd_hausdorff_dis(a, b):
    """utils/dis_utils_torch.py:19-22: max_i min_j M[0]."""
    return _sweep(a, b).row_max[0]


def bid_hausdorff_dis(a, b):
    """utils/dis_utils_torch.py:25-28: max(d_ab, d_ba); both directions come from ONE sweep
    (d_ba = max_j min_i M[0] is the column side)."""
    r = _sweep(a, b)
    return torch.max(r.row_max[0], r.col_max[0])
