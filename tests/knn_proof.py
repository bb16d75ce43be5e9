"""Proof-style comparison of two k-NN index tensors computed in fp32 with DIFFERENT summation orders
(the reference's library GEMM vs. the sequential chain of the oracle / the kernels).

For C >= 64 the reference's `matmul` is blocked by the BLAS library, so its fp32 distances differ from
a sequential evaluation by rounding, and the indices may differ where two candidates are closer than
that rounding.  Instead of a bare agreement fraction the tests assert what is provable:

  Let D be the exact (float64) squared distances of the fp32 inputs and, for one row i,
      e = max_j  (2C + 6) u (|x_i|^2 + |x_j|^2),    u = 2^-24,
  an upper bound of |fl(d_ij) - D_ij| for ANY fp32 evaluation order of (-|x_i|^2 + 2 x_i.x_j) - |x_j|^2
  (gamma_C for the dot product and each norm, two more roundings for the additions).  Then
    1. at every rank r, |D[ours_r] - D[ref_r]| <= 4e            (both picks lie within 2e of the true r-th value)
    2. if the true gap D_(k+1) - D_(k) > 2e the two k-sets are identical
    3. if every gap among the first k+1 true values exceeds 2e the two index rows are identical.
Returned: statistics for the log (fraction of differing entries, of rows covered by 2. and 3.).
"""
import numpy as np


def exact_sqdist(x_cf):
    x = np.asarray(x_cf, np.float64).transpose(0, 2, 1)               # [B, N, C]
    xx = (x * x).sum(-1)
    return xx[:, :, None] + xx[:, None, :] - 2.0 * (x @ x.transpose(0, 2, 1)), xx


def assert_knn_near_tie_proof(idx, ref_idx, x_cf, what=""):
    idx = np.asarray(idx, np.int64); ref_idx = np.asarray(ref_idx, np.int64)
    B, C, N = x_cf.shape
    k = idx.shape[2]
    assert idx.shape == ref_idx.shape == (B, N, k)
    D, xx = exact_sqdist(x_cf)
    u = 2.0 ** -24
    e = ((2 * C + 6) * u * (xx[:, :, None] + xx[:, None, :])).max(-1)            # [B, N]
    d_ours = np.take_along_axis(D, idx, 2); d_ref = np.take_along_axis(D, ref_idx, 2)
    # 1. rank-wise near ties
    worst = np.abs(d_ours - d_ref) / (4 * e[:, :, None])
    assert worst.max() <= 1.0, f"{what}: picks differ by {worst.max():.2f} x the rounding bound"
    # 2. / 3. exact agreement wherever the true gaps exceed the bound
    Ds = np.sort(D, -1)[:, :, :k + 1]
    gaps = np.diff(Ds, axis=-1)                                                  # [B, N, k]
    set_clear = gaps[:, :, k - 1] > 2 * e if N > k else np.ones((B, N), bool)
    rank_clear = (gaps > 2 * e[:, :, None]).all(-1) if N > k else (gaps[:, :, :k - 1] > 2 * e[:, :, None]).all(-1)
    same_set = (np.sort(idx, -1) == np.sort(ref_idx, -1)).all(-1)
    same_row = (idx == ref_idx).all(-1)
    assert same_set[set_clear].all(), f"{what}: k-sets differ although the k/(k+1) gap exceeds the rounding bound"
    assert same_row[rank_clear].all(), f"{what}: ranks differ although every gap exceeds the rounding bound"
    return {"differing_entries": float((idx != ref_idx).mean()), "rows_set_provable": float(set_clear.mean()),
            "rows_rank_provable": float(rank_clear.mean()), "rows_identical": float(same_row.mean()),
            "worst_over_bound": float(worst.max())}
