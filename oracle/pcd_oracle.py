"""CPU oracle for the point-set distance path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product (3dpointcloudattack_b200/) never does and
fails loudly when its CUDA library is missing.

Two layers:

* thin ctypes bindings over oracle/pcd_oracle.c (the fp32, op-order-faithful
  arithmetic: see the header of that file for the three addition orders), and
* numpy restatements of every reference symbol on the path, each citing the
  reference file:line it follows (paths relative to /root/reference).

Parity status: PINNED.  oracle/make_golden.py imports the unmodified reference from
/root/reference, runs it on seeded inputs on CPU (torch 2.11.0) and stores inputs and
outputs under tests/golden/; tests/test_oracle_golden.py checks this module against
those vectors (distance matrices and indices bit-exact, reductions to 1e-6).
"""
from __future__ import annotations

import ctypes
import os
from collections import namedtuple

import numpy as np

FORM_ROW_COL, FORM_COL_ROW, FORM_SUM_FIRST = 0, 1, 2
NORM_MULSUM, NORM_FMA = 0, 1

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libpcd_oracle.so")
        if not os.path.exists(path):
            import importlib.util
            spec = importlib.util.spec_from_file_location("_orc_build", os.path.join(_HERE, "build.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mod.build()
        _LIB = ctypes.CDLL(path)
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# --------------------------------------------------------------------------- C bindings
def norms(kind, pts):
    """pts [B,N,C] point-major -> [B,N] squared norms in the given rounding order."""
    pts = _f32(pts)
    B, N, C = pts.shape
    out = np.empty((B, N), np.float32)
    lib().orc_norms(ctypes.c_int(kind), _p(pts), ctypes.c_int64(B * N), ctypes.c_int(C), _p(out))
    return out


def pairwise(form, rows, cols, nrow, ncol):
    rows, cols, nrow, ncol = _f32(rows), _f32(cols), _f32(nrow), _f32(ncol)
    B, N, C = rows.shape
    M = cols.shape[1]
    out = np.empty((B, N, M), np.float32)
    lib().orc_pairwise(ctypes.c_int(form), _p(rows), _p(cols), _p(nrow), _p(ncol),
                       ctypes.c_int(B), ctypes.c_int(N), ctypes.c_int(M), ctypes.c_int(C), _p(out))
    return out


NN1 = namedtuple("NN1", "row_min row_arg col_min col_arg")


def nn1(form, rows, cols, nrow, ncol):
    """Row/column minima with lowest-index argmin, no [B,N,M] matrix."""
    rows, cols, nrow, ncol = _f32(rows), _f32(cols), _f32(nrow), _f32(ncol)
    B, N, C = rows.shape
    M = cols.shape[1]
    rmin = np.empty((B, N), np.float32); rarg = np.empty((B, N), np.int32)
    cmin = np.empty((B, M), np.float32); carg = np.empty((B, M), np.int32)
    lib().orc_nn1(ctypes.c_int(form), _p(rows), _p(cols), _p(nrow), _p(ncol),
                  ctypes.c_int(B), ctypes.c_int(N), ctypes.c_int(M), ctypes.c_int(C),
                  _p(rmin), _p(rarg), _p(cmin), _p(carg))
    return NN1(rmin, rarg, cmin, carg)


def knn(form, rows, cols, nrow, ncol, K):
    """K smallest per row, ascending by (distance, index)."""
    rows, cols, nrow, ncol = _f32(rows), _f32(cols), _f32(nrow), _f32(ncol)
    B, N, C = rows.shape
    M = cols.shape[1]
    assert 1 <= K <= M
    d = np.empty((B, N, K), np.float32); i = np.empty((B, N, K), np.int32)
    lib().orc_knn(ctypes.c_int(form), _p(rows), _p(cols), _p(nrow), _p(ncol),
                  ctypes.c_int(B), ctypes.c_int(N), ctypes.c_int(M), ctypes.c_int(C),
                  ctypes.c_int(K), _p(d), _p(i))
    return d, i


def ball_query(radius, nsample, xyz, new_xyz):
    """model/pointnet2_utils.py:84-104 -> idx [B,S,nsample] int64."""
    xyz, new_xyz = _f32(xyz), _f32(new_xyz)
    B, N, C = xyz.shape
    S = new_xyz.shape[1]
    nrow = norms(NORM_MULSUM, new_xyz); ncol = norms(NORM_MULSUM, xyz)
    r2 = np.float32(float(radius) ** 2)       # python float radius**2, compared in fp32 (line 99)
    out = np.empty((B, S, nsample), np.int32)
    lib().orc_ball_query(_p(new_xyz), _p(xyz), _p(nrow), _p(ncol), ctypes.c_int(B), ctypes.c_int(S),
                         ctypes.c_int(N), ctypes.c_int(C), ctypes.c_float(r2), ctypes.c_int(nsample), _p(out))
    return out.astype(np.int64)


# ------------------------------------------------------------------ a1: utils/dis_utils_torch.py
def _cf_to_pm(a):
    """[B,C,N] channel-first -> [B,N,C] point-major contiguous."""
    return _f32(np.transpose(np.asarray(a), (0, 2, 1)))


def dis_pairwise_raw(a, b):
    """Pre-clamp, pre-sqrt matrix of torch.cdist's mm path (utils/dis_utils_torch.py:8-11;
    ATen _euclidean_dist: x1_=[-2*x1, |x1|^2, 1], x2_=[x2, 1, |x2|^2], x1_ @ x2_^T)."""
    A, Bm = _cf_to_pm(a), _cf_to_pm(b)
    return pairwise(FORM_ROW_COL, A, Bm, norms(NORM_MULSUM, A), norms(NORM_MULSUM, Bm))


def dis_pairwise_distances(a, b):
    """utils/dis_utils_torch.py:8-11 -> [B,N,M] L2 (clamp_min(0).sqrt())."""
    return np.sqrt(np.maximum(dis_pairwise_raw(a, b), np.float32(0)))


def dis_chamfer(a, b):
    """utils/dis_utils_torch.py:14-16.  Divisors are a.shape[1], b.shape[1] (= 3 for
    [B,3,N] input -- reference quirk) and only sample 0 is returned."""
    Mx = dis_pairwise_distances(a, b)
    return (Mx.min(1).sum(1, dtype=np.float32) / np.float32(a.shape[1])
            + Mx.min(2).sum(1, dtype=np.float32) / np.float32(b.shape[1]))[0]


def dis_sgd_hausdorff(a, b):
    """utils/dis_utils_torch.py:19-22: max_i min_j M[0]."""
    return dis_pairwise_distances(a, b)[0].min(1).max()


def dis_bid_hausdorff(a, b):
    """utils/dis_utils_torch.py:25-28."""
    return np.maximum(dis_sgd_hausdorff(a, b), dis_sgd_hausdorff(b, a))


def dis_grads(a, b, w_col_sum=0.0, w_row_sum=0.0, w_row_max=0.0, w_col_max=0.0):
    """float64 closed-form gradient (through the fp32 argmins) of
         w_col_sum * sum_j min_i M + w_row_sum * sum_i min_j M + w_row_max * max_i min_j M
         + w_col_max * max_j min_i M            for sample 0 of utils/dis_utils_torch.py's M,
    i.e. chamfer = (1/3, 1/3, 0, 0), sgd_hausdorff = (0, 0, 1, 0).  d sqrt(d2)/dp = (p - q)/M,
    0 where M == 0 (cdist backward).  Returns (grad_a, grad_b) shaped like the inputs."""
    A, Bm = _cf_to_pm(a)[:1], _cf_to_pm(b)[:1]
    r = nn1(FORM_ROW_COL, A, Bm, norms(NORM_MULSUM, A), norms(NORM_MULSUM, Bm))
    A64, B64 = A[0].astype(np.float64), Bm[0].astype(np.float64)
    ga = np.zeros_like(A64); gb = np.zeros_like(B64)
    rowv = np.sqrt(np.maximum(r.row_min[0], 0)).astype(np.float64)
    colv = np.sqrt(np.maximum(r.col_min[0], 0)).astype(np.float64)
    gr = np.full(A64.shape[0], float(w_row_sum)); gr[int(np.argmax(rowv))] += float(w_row_max)
    gc = np.full(B64.shape[0], float(w_col_sum)); gc[int(np.argmax(colv))] += float(w_col_max)
    with np.errstate(divide="ignore", invalid="ignore"):
        fr = np.where(rowv > 0, gr / rowv, 0.0); fc = np.where(colv > 0, gc / colv, 0.0)
    diff = A64 - B64[r.row_arg[0]]
    ga += fr[:, None] * diff; np.add.at(gb, r.row_arg[0], -fr[:, None] * diff)
    diff = B64 - A64[r.col_arg[0]]
    gb += fc[:, None] * diff; np.add.at(ga, r.col_arg[0], -fc[:, None] * diff)
    out_a = np.zeros(np.asarray(a).shape, np.float64); out_b = np.zeros(np.asarray(b).shape, np.float64)
    out_a[0] = ga.T; out_b[0] = gb.T
    return out_a, out_b


# ------------------------------------------------------- a2: attack/CW/CW_utils/distance.py
def batch_pairwise_dist(x, y):
    """attack/CW/CW_utils/distance.py:15-32: P = rx^T + ry - 2*zz, norms = diag of bmm."""
    x, y = _f32(x), _f32(y)
    return pairwise(FORM_SUM_FIRST, x, y, norms(NORM_FMA, x), norms(NORM_FMA, y))


def _nn1_cw(preds, gts):
    gts, preds = _f32(gts), _f32(preds)
    return nn1(FORM_SUM_FIRST, gts, preds, norms(NORM_FMA, gts), norms(NORM_FMA, preds))


def chamfer_distance(preds, gts):
    """attack/CW/CW_utils/distance.py:40-50 -> (loss1[B], loss2[B]).
    P = batch_pairwise_dist(gts, preds): rows = gts, cols = preds;
    loss1 = mean over preds of min over gts (column minima), loss2 = mean of row minima."""
    r = _nn1_cw(preds, gts)
    return r.col_min.mean(1, dtype=np.float32), r.row_min.mean(1, dtype=np.float32)


def hausdorff_distance(preds, gts):
    """attack/CW/CW_utils/distance.py:58-70."""
    r = _nn1_cw(preds, gts)
    return r.col_min.max(1), r.row_min.max(1)


def chamfer_distance_grads(preds, gts, g1, g2):
    """d(sum_b g1[b]*loss1[b] + g2[b]*loss2[b]) / d(preds, gts) in float64, closed form
    through the argmin indices (what autograd does through torch.min(dim))."""
    preds64, gts64 = np.asarray(preds, np.float64), np.asarray(gts, np.float64)
    r = _nn1_cw(preds, gts)
    B, N1, _ = preds64.shape
    N2 = gts64.shape[1]
    gp = np.zeros_like(preds64); gg = np.zeros_like(gts64)
    for b in range(B):
        # loss1: each pred j -> nearest gt col_arg[j]
        w = g1[b] / N1
        diff = preds64[b] - gts64[b][r.col_arg[b]]
        gp[b] += 2 * w * diff
        np.add.at(gg[b], r.col_arg[b], -2 * w * diff)
        # loss2: each gt i -> nearest pred row_arg[i]
        w = g2[b] / N2
        diff = gts64[b] - preds64[b][r.row_arg[b]]
        gg[b] += 2 * w * diff
        np.add.at(gp[b], r.row_arg[b], -2 * w * diff)
    return gp, gg


def hausdorff_distance_grads(preds, gts, g1, g2):
    """Same for attack/CW/CW_utils/distance.py:58-70: torch.max(mins, dim=1) routes the
    gradient to the first maximal entry."""
    preds64, gts64 = np.asarray(preds, np.float64), np.asarray(gts, np.float64)
    r = _nn1_cw(preds, gts)
    gp = np.zeros_like(preds64); gg = np.zeros_like(gts64)
    for b in range(preds64.shape[0]):
        j = int(np.argmax(r.col_min[b])); i = int(r.col_arg[b, j])
        diff = preds64[b, j] - gts64[b, i]
        gp[b, j] += 2 * g1[b] * diff; gg[b, i] -= 2 * g1[b] * diff
        i = int(np.argmax(r.row_min[b])); j = int(r.row_arg[b, i])
        diff = gts64[b, i] - preds64[b, j]
        gg[b, i] += 2 * g2[b] * diff; gp[b, j] -= 2 * g2[b] * diff
    return gp, gg


# ------------------------------------------------------------- a3: attack/GeoA3/knn_utils.py
def knn_points(p1, p2, K=1):
    """attack/GeoA3/knn_utils.py:10-55.  dist[i,j] = (|p1_j|^2 + inner_ij) + |p2_i|^2 --
    the reference's broadcast lays p1's norms along the COLUMN axis and p2's along the ROW
    axis (requires P1 == P2).  Returns (dists[B,P1,K] ascending, idx[B,P1,K] int64)."""
    p1, p2 = _f32(p1), _f32(p2)
    if p1.shape[1] != p2.shape[1]:
        raise RuntimeError("reference broadcast requires P1 == P2")
    ncol = norms(NORM_MULSUM, p1)    # indexed by j
    nrow = norms(NORM_MULSUM, p2)    # indexed by i
    d, i = knn(FORM_COL_ROW, p1, p2, nrow, ncol, K)
    return d, i.astype(np.int64)


def knn_points_grads(p1, p2, idx, gw):
    """float64 closed-form gradient of sum(dists * gw) for attack/GeoA3/knn_utils.py's
    dist[i,j] = |p1_j|^2 - 2 p1_i.p2_j + |p2_i|^2 through GIVEN indices idx[B,P,K]:
    d/dp1_i = -2g p2_j ; d/dp2_j = -2g p1_i ; d/dp1_j = 2g p1_j ; d/dp2_i = 2g p2_i."""
    a, b = np.asarray(p1, np.float64), np.asarray(p2, np.float64)
    g1, g2 = np.zeros_like(a), np.zeros_like(b)
    B, P, K = idx.shape
    for bb in range(B):
        for k in range(K):
            j = idx[bb, :, k]; g = np.asarray(gw, np.float64)[bb, :, k][:, None]
            g1[bb] += -2 * g * b[bb][j]
            np.add.at(g2[bb], j, -2 * g * a[bb])
            np.add.at(g1[bb], j, 2 * g * a[bb][j])
            g2[bb] += 2 * g * b[bb]
    return g1, g2


def knn_points_matrix(p1, p2):
    p1, p2 = _f32(p1), _f32(p2)
    return pairwise(FORM_COL_ROW, p1, p2, norms(NORM_MULSUM, p2), norms(NORM_MULSUM, p1))


def knn_gather(x, idx):
    """attack/GeoA3/knn_utils.py:58-86: x[B,M,U], idx[B,L,K] -> [B,L,K,U]."""
    x = np.asarray(x)
    B = x.shape[0]
    return x[np.arange(B)[:, None, None], idx]


# ------------------------------------- a5: attack/CW/CW_utils/dist_utils.py:125-160 (KNNDist)
def knn_dist_matrix(pc):
    pc = _f32(pc)
    n = norms(NORM_MULSUM, pc)
    return pairwise(FORM_COL_ROW, pc, pc, n, n)


def knn_dist_loss(pc, k=5, alpha=1.05):
    """attack/CW/CW_utils/dist_utils.py:125-160 with weights=None -> loss[B]."""
    pc = _f32(pc)
    n = norms(NORM_MULSUM, pc)
    d, _ = knn(FORM_COL_ROW, pc, pc, n, n, k + 1)
    value = d[..., 1:].mean(-1, dtype=np.float32)                    # [B,K]
    mean = value.mean(-1, dtype=np.float32)
    std = value.std(-1, ddof=1, dtype=np.float64).astype(np.float32)  # torch.std: unbiased
    thr = mean + np.float32(alpha) * std
    mask = (value > thr[:, None]).astype(np.float32)
    return (value * mask).mean(1, dtype=np.float32), value, mask


# ---------------------------------------------- a6: model/dgcnn.py:194-200, curvenet_util.py
def dgcnn_knn(x, k):
    """model/dgcnn.py:194-200: x[B,C,N] -> idx[B,N,k] int64, nearest first, self included.
    pairwise = (-xx - inner) - xx^T is the exact negation of FORM_COL_ROW, so top-k largest
    of it = k smallest here."""
    pm = _cf_to_pm(x)
    n = norms(NORM_MULSUM, pm)
    _, i = knn(FORM_COL_ROW, pm, pm, n, n, k)
    return i.astype(np.int64)


def dgcnn_neg_matrix(x):
    pm = _cf_to_pm(x)
    n = norms(NORM_MULSUM, pm)
    return -pairwise(FORM_COL_ROW, pm, pm, n, n)


# --------------------------------------------------- a7: model/pointnet2_utils.py:19-38
def square_distance(src, dst):
    src, dst = _f32(src), _f32(dst)
    return pairwise(FORM_ROW_COL, src, dst, norms(NORM_MULSUM, src), norms(NORM_MULSUM, dst))


def query_ball_point(radius, nsample, xyz, new_xyz):
    return ball_query(radius, nsample, xyz, new_xyz)


# ------------------------------ f-2: model/dgcnn.py:203-227, model/curvenet_util.py:206-236
EDGE_CENTER, EDGE_NEIGHBOR, EDGE_DIFF = 0, 1, 2


def edge_feature(x, idx, ops):
    """out[b, q*C + c, n, j] = ops[q](centre x[b,c,n], neighbour x[b,c,idx[b,n,j]]):
    CENTER -> centre, NEIGHBOR -> neighbour, DIFF -> neighbour - centre (one fp32 subtraction,
    exact restatement of `feature - x`).  x[B,C,N], idx[B,N,k] -> [B, len(ops)*C, N, k]."""
    x = _f32(x); idx = np.asarray(idx)
    B, C, N = x.shape
    k = idx.shape[2]
    nb = x[np.arange(B)[:, None, None, None], np.arange(C)[None, :, None, None], idx[:, None].astype(np.int64)]
    ctr = np.broadcast_to(x[:, :, :, None], (B, C, N, k))
    blocks = []
    for op in ops:
        blocks.append(ctr if op == EDGE_CENTER else nb if op == EDGE_NEIGHBOR else nb - ctr)
    return np.ascontiguousarray(np.concatenate(blocks, axis=1), dtype=np.float32)


def edge_feature_grad(g, idx, ops, C):
    """float64 closed form of d(sum(g*out))/dx for edge_feature: own terms plus the scatter
    through idx (what autograd's index / cat / sub backward computes)."""
    g = np.asarray(g, np.float64); idx = np.asarray(idx).astype(np.int64)
    B, _, N, k = g.shape
    gx = np.zeros((B, C, N), np.float64)
    for q, op in enumerate(ops):
        gq = g[:, q * C:(q + 1) * C]                       # [B,C,N,k]
        if op in (EDGE_CENTER, EDGE_DIFF):
            gx += gq.sum(3) * (1.0 if op == EDGE_CENTER else -1.0)
        if op in (EDGE_NEIGHBOR, EDGE_DIFF):
            for b in range(B):
                for c in range(C):
                    np.add.at(gx[b, c], idx[b].reshape(-1), gq[b, c].reshape(-1))
    return gx


def get_graph_feature(x, k=20, idx=None):
    """model/dgcnn.py:203-227: cat(feature - x, x) permuted to [B,2C,N,k]."""
    if idx is None:
        idx = dgcnn_knn(x, k)
    return edge_feature(x, idx, (EDGE_DIFF, EDGE_CENTER))


def lpfa_point_feature(xyz, idx):
    """model/curvenet_util.py:219-227: cat(points, point_feature, point_feature - points) -> [B,9,N,k]."""
    return edge_feature(xyz, idx, (EDGE_CENTER, EDGE_NEIGHBOR, EDGE_DIFF))


# ---------------- f-3: model/pointnet2_utils.py:41-81, model/curvenet_util.py:69-90 (start 0)
def farthest_point_sample(xyz, npoint, start=None):
    """xyz[B,N,3] -> centroids[B,npoint] int64; start[B] = the reference's torch.randint draw
    (pointnet2_utils.py:71) or None for CurveNet's fixed start 0."""
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    out = np.empty((B, npoint), np.int32)
    st = None if start is None else np.ascontiguousarray(start, dtype=np.int32)
    lib().orc_fps(_p(xyz), ctypes.c_int(B), ctypes.c_int(N), ctypes.c_int(npoint),
                  None if st is None else _p(st), _p(out))
    return out.astype(np.int64)


def index_points(points, idx):
    """model/pointnet2_utils.py:41-57: points[B,N,C], idx[B,S] or [B,S,K] -> gathered."""
    points = np.asarray(points); idx = np.asarray(idx)
    B = points.shape[0]
    bi = np.arange(B).reshape((B,) + (1,) * (idx.ndim - 1))
    return points[bi, idx]


def three_nn_interpolate(xyz1, xyz2, points2):
    """model/pointnet2_utils.py:289-300: xyz1[B,N,3], xyz2[B,S,3], points2[B,S,D] -> [B,N,D]
    (square_distance = FORM_ROW_COL, three smallest ascending, 1/(d+1e-8) weights)."""
    xyz1, xyz2, points2 = _f32(xyz1), _f32(xyz2), _f32(points2)
    d, i = knn(FORM_ROW_COL, xyz1, xyz2, norms(NORM_MULSUM, xyz1), norms(NORM_MULSUM, xyz2), 3)
    recip = np.float32(1.0) / (d + np.float32(1e-8))
    w = recip / recip.sum(2, keepdims=True, dtype=np.float32)
    return (index_points(points2, i.astype(np.int64)) * w[..., None]).sum(2, dtype=np.float32), d, i


# ------------------------ f-4: attack/GeoA3/utility.py:43-92, loss_utils.py:60-141, AOF TAOF_attack.py:31-52
def self_knn_idx(pc_cf, K):
    """Self k-NN of a [b,3,n] cloud in knn_points arithmetic (exact for p1 is p2) -> idx [b,n,K] int64."""
    pm = _cf_to_pm(pc_cf)
    _, i = knn_points(pm, pm, K)
    return i


def patch_covariance(pc_pm, idx, skip_first=True):
    """utility.py:56-62: fp32 covariance of each point's k neighbours, formed as the reference forms it --
    mean = sum / k, centred set, bmm (k ascending), times fp32(1/(k-1)).  Returns (cov [b,n,3,3] fp32,
    nbr_sum [b,n,3] fp32 = sum of the centred neighbours, :68)."""
    pc = _f32(pc_pm)
    nb = idx[:, :, 1:] if skip_first else idx
    B, N, k = nb.shape
    pts = pc[np.arange(B)[:, None, None], nb]                       # [b,n,k,3]
    s = np.zeros((B, N, 3), np.float32)
    for j in range(k):
        s = (s + pts[:, :, j]).astype(np.float32)
    mean = (s / np.float32(k)).astype(np.float32)
    c = (pts - mean[:, :, None, :]).astype(np.float32)
    cov = np.zeros((B, N, 3, 3), np.float32)
    nsum = np.zeros((B, N, 3), np.float32)
    for j in range(k):
        # fma(c_a, c_b, acc): evaluate in float64 and round once (the product of two fp32 is exact in fp64)
        cov = (cov.astype(np.float64) + c[:, :, j, :, None].astype(np.float64) * c[:, :, j, None, :].astype(np.float64)).astype(np.float32)
        nsum = (nsum + c[:, :, j]).astype(np.float32)
    fact = np.float32(1.0) / np.float32(k - 1)
    return (fact * cov).astype(np.float32), nsum


def local_frames(pc_pm, idx, skip_first=True):
    """Eigen-frame of patch_covariance in float64 (numpy eigh): (evals [b,n,3] ascending, evecs [b,n,3,3] with
    ROWS = eigenvectors in that order, nbr_sum).  The normal of estimate_normal is +-evecs[..., 0, :]."""
    cov, nsum = patch_covariance(pc_pm, idx, skip_first)
    w, v = np.linalg.eigh(cov.astype(np.float64))
    return w, np.swapaxes(v, -1, -2), nsum


def kappa(pc_pm, normal_pm, idx, nidx=None, skip_first=True, dtype=np.float32):
    """loss_utils.py:60-90: mean_j |<unit(q_j - p), n>|, unit(d) = d / max(|d|, 1e-12); n = normal[nidx or self].
    dtype float32 restates the reference's op order; float64 is the closed form used to pin tolerances."""
    pc = np.asarray(pc_pm, dtype); nrm = np.asarray(normal_pm, dtype)
    nb = idx[:, :, 1:] if skip_first else idx
    B, N, k = nb.shape
    ar = np.arange(B)[:, None]
    n_ = nrm[ar, nidx] if nidx is not None else nrm
    q = pc[np.arange(B)[:, None, None], nb]
    d = (q - pc[:, :, None, :]).astype(dtype)
    r = np.sqrt(((d[..., 0] * d[..., 0]).astype(dtype) + (d[..., 1] * d[..., 1]).astype(dtype)).astype(dtype) + (d[..., 2] * d[..., 2]).astype(dtype)).astype(dtype)
    u = (d / np.maximum(r, dtype(1e-12))[..., None]).astype(dtype)
    dot = ((u[..., 0] * n_[:, :, None, 0]).astype(dtype) + (u[..., 1] * n_[:, :, None, 1]).astype(dtype)).astype(dtype) + (u[..., 2] * n_[:, :, None, 2]).astype(dtype)
    a = np.abs(dot.astype(dtype))
    acc = np.zeros((B, N), dtype)
    for j in range(k):
        acc = (acc + a[:, :, j]).astype(dtype)
    return (acc / dtype(k)).astype(dtype)


def kappa_grad64(pc_pm, normal_pm, idx, g, nidx=None, skip_first=True):
    """float64 closed-form gradient of sum_{b,i} g[b,i] kappa[b,i] w.r.t. the cloud (normals constant)."""
    pc = np.asarray(pc_pm, np.float64); nrm = np.asarray(normal_pm, np.float64); g = np.asarray(g, np.float64)
    nb = idx[:, :, 1:] if skip_first else idx
    B, N, k = nb.shape
    n_ = nrm[np.arange(B)[:, None], nidx] if nidx is not None else nrm
    q = pc[np.arange(B)[:, None, None], nb]
    d = q - pc[:, :, None, :]
    r = np.maximum(np.sqrt((d * d).sum(-1)), 1e-300)
    u = d / r[..., None]
    s = (u * n_[:, :, None, :]).sum(-1)
    term = (np.sign(s) * g[:, :, None] / k)[..., None] * (n_[:, :, None, :] - s[..., None] * u) / r[..., None]     # d/d(d_ij)
    grad = np.zeros_like(pc)
    grad -= term.sum(2)
    for b in range(B):
        np.add.at(grad[b], nb[b].reshape(-1), term[b].reshape(-1, 3))
    return grad


def graph_laplacian(pc_pm, idx):
    """attack/AOF/TAOF_attack.py:36-50: L = D - A, A_ij = exp(-((dx^2 + dy^2) + dz^2)) on the symmetrised k-NN graph."""
    pc = _f32(pc_pm)
    B, N, _ = pc.shape
    L = np.zeros((B, N, N), np.float32)
    for b in range(B):
        d = (pc[b][:, None, :] - pc[b][None, :, :]).astype(np.float32)
        d2 = ((d[..., 0] * d[..., 0]).astype(np.float32) + (d[..., 1] * d[..., 1]).astype(np.float32)).astype(np.float32) + (d[..., 2] * d[..., 2]).astype(np.float32)
        A = np.exp(-d2.astype(np.float32)).astype(np.float32)
        mask = np.zeros((N, N), bool)
        mask[np.arange(N)[:, None], idx[b]] = True
        mask |= mask.T
        A = np.where(mask, A, np.float32(0))
        L[b] = np.diag(A.sum(1, dtype=np.float64).astype(np.float32)) - A
    return L


# ------------------------------------- f-1 epilogues: clip_utils.py:5-136, GeoA3_attack.py:62-101
# numpy float32 arithmetic rounds every operation on its own (no contraction), which is the contract of the kernels.
_F = np.float32


def _sumsq3(d):                      # torch.sum(d ** 2, dim=1) on [B,3,K]: (x*x + y*y) + z*z
    return (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]


def _clip_scale(norm, budget):       # clamp(budget / (norm + 1e-9), max=1); scalar / tensor = reciprocal(tensor) * scalar
    s = (_F(1.0) / (norm + _F(1e-9))) * _F(budget)
    return np.where(s > _F(1.0), _F(1.0), s).astype(np.float32)


def clip_points_linf(pc, ori, budget):
    """attack/CW/CW_utils/clip_utils.py:32-56, pc/ori [B,3,K]."""
    pc, ori = _f32(pc), _f32(ori)
    d = pc - ori
    with np.errstate(all="ignore"):
        s = _clip_scale(np.sqrt(_sumsq3(d)), budget)
    return ori + d * s[:, None, :]


def clip_points_l2(pc, ori, budget):
    """clip_utils.py:5-29 (the 3K-term sum in float32 pairwise order: agrees with torch / the kernel to rounding only)."""
    pc, ori = _f32(pc), _f32(ori)
    d = pc - ori
    norm = np.sqrt((d * d).reshape(d.shape[0], -1).sum(1, dtype=np.float32))
    return ori + d * _clip_scale(norm, budget)[:, None, None]


def _cross3(a, b):                   # torch.cross(a, b, dim=1)
    return np.stack([a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1], a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2],
                     a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]], 1)


def project_inner_clip_linf(pc, ori, normal, budget):
    """clip_utils.py:59-136 (ProjectInnerPoints, then ClipPointsLinf)."""
    pc, ori, normal = _f32(pc), _f32(ori), _f32(normal)
    d = pc - ori
    inner = ((d[:, 0] * normal[:, 0] + d[:, 1] * normal[:, 1]) + d[:, 2] * normal[:, 2]) < 0
    vng = _cross3(normal, d)
    vng_norm = np.sqrt(_sumsq3(vng))
    vref = _cross3(vng, normal)
    proj = d * vref / (np.sqrt(_sumsq3(vref)) + _F(1e-9))[:, None, :]
    proj = np.where((inner & (vng_norm < _F(1e-6)))[:, None, :], _F(0), proj)
    d = np.where(inner[:, None, :], proj, d).astype(np.float32)
    return clip_points_linf(ori + d, ori, budget)


def lp_clip(offset, cc_linf):
    """attack/GeoA3/GeoA3_attack.py:92-101."""
    o = _f32(offset)
    ln = np.sqrt(_sumsq3(o))[:, None, :]
    with np.errstate(all="ignore"):
        scaled = np.where(ln > _F(1e-6), o / ln * _F(cc_linf), _F(0))
    return np.where(ln < _F(cc_linf), o, scaled).astype(np.float32)


def _nearest_ori(q_cf, ori_cf):
    """the K=1 knn_points of GeoA3_attack.py:68 / :84 on channel-first clouds -> idx [B,K]"""
    return knn_points(np.ascontiguousarray(q_cf.transpose(0, 2, 1)), np.ascontiguousarray(ori_cf.transpose(0, 2, 1)), K=1)[1][:, :, 0]


def offset_proj(offset, ori_pc, ori_normal, idx=None):
    """GeoA3_attack.py:62-81 (channel-first [B,3,K]); idx = the K=1 neighbour of every offset among ori_pc."""
    o, nrm = _f32(offset), _f32(ori_normal)
    idx = _nearest_ori(o, _f32(ori_pc)) if idx is None else np.asarray(idx)
    g = np.take_along_axis(nrm, idx[:, None, :].repeat(3, 1), 2)
    t = (np.sqrt(_sumsq3(g)) + _F(1e-6))[:, None, :]
    a = o * g / t
    dot = ((a[:, 0] + a[:, 1]) + a[:, 2])[:, None, :]
    return (dot * g / t).astype(np.float32)


def find_offset(ori_pc, adv_pc, idx=None):
    """GeoA3_attack.py:83-89."""
    ori, adv = _f32(ori_pc), _f32(adv_pc)
    idx = _nearest_ori(adv, ori) if idx is None else np.asarray(idx)
    return adv - np.take_along_axis(ori, idx[:, None, :].repeat(3, 1), 2)
