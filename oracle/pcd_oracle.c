/*
 * pcd_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * Plain-C, fp32, op-order-faithful restatement of the point-set distance
 * arithmetic of LI-Yiquan/3DPointCloudAttack.  Only tests/, the smoke check in
 * __graft_entry__.py and the cpu_baseline / --impl reference legs of bench.py
 * may load this library; the product path (3dpointcloudattack_b200/) never does.
 *
 * Every reference formulation of the path evaluates the expansion
 *     d(i,j) = |row_i|^2 + |col_j|^2 - 2 <row_i, col_j>
 * with a GEMM for the inner product and broadcast adds for the norms.  What
 * differs between the five copies is the ORDER of the fp32 additions, which
 * matters because near neighbours are cancellation dominated.  The inner
 * product of every reference GEMM (torch.bmm / matmul, K = C) is a sequential
 * FMA chain over k = 0..C-1 starting from the plain product (probed bit-exact
 * against torch 2.11 CPU, see oracle/make_golden.py).  The factor -2 is a power
 * of two and therefore commutes with every rounding; it is folded into t.
 *
 *   t(i,j) = -2 * fma(row[C-1], col[C-1], ... fma(row[1], col[1], row[0]*col[0]))
 *
 *   FORM_ROW_COL   d = (t + nrow[i]) + ncol[j]
 *       utils/dis_utils_torch.py:8-11 (torch.cdist mm path, _euclidean_dist:
 *       x1_=[-2x,|x|^2,1] @ x2_=[y,1,|y|^2]^T is the same 5-term chain) and
 *       model/pointnet2_utils.py:35-37 (square_distance: dist=-2mm; +=src; +=dst)
 *   FORM_COL_ROW   d = (t + ncol[j]) + nrow[i]
 *       attack/GeoA3/knn_utils.py:12-15 (apply_knn: p1_2 + inner + p2_2^T),
 *       attack/CW/CW_utils/dist_utils.py:135-137 (KNNDist: xx + inner + xx^T),
 *       model/dgcnn.py:195-197 and model/curvenet_util.py:12-14 (negated)
 *   FORM_SUM_FIRST d = (nrow[i] + ncol[j]) + t
 *       attack/CW/CW_utils/distance.py:18-31 (rx^T + ry - 2*zz)
 *
 * Norm kinds:
 *   NORM_MULSUM  ((x0*x0 + x1*x1) + x2*x2) + ...   torch.sum(x**2, dim) / x.pow(2).sum(-1)
 *   NORM_FMA     fma(x2,x2, fma(x1,x1, x0*x0))      diagonal of torch.bmm(x, x^T)
 *
 * Tie rule everywhere: lowest index wins (torch.min(dim) semantics; for top-k
 * this is the "stable" restatement SURVEY.md section 7 hard part 4 asks for).
 *
 * Build: gcc -O2 -ffp-contract=off -mfma -fopenmp -shared -fPIC (oracle/build.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { FORM_ROW_COL = 0, FORM_COL_ROW = 1, FORM_SUM_FIRST = 2 };
enum { NORM_MULSUM = 0, NORM_FMA = 1 };

/* -2 * sequential FMA chain, k ascending (torch.bmm/matmul with K=C on CPU). */
static inline float inner_m2(const float *a, const float *b, int C) {
    float acc = a[0] * b[0];
    for (int k = 1; k < C; ++k) acc = fmaf(a[k], b[k], acc);
    return -2.0f * acc;
}

static inline float pair_dist(int form, const float *row, const float *col, int C,
                              float nrow, float ncol) {
    float t = inner_m2(row, col, C);
    switch (form) {
    case FORM_ROW_COL: { float u = t + nrow; return u + ncol; }
    case FORM_COL_ROW: { float u = t + ncol; return u + nrow; }
    default:           { float s = nrow + ncol; return s + t; }
    }
}

/* pts[B*N, C] point-major -> out[B*N] */
void orc_norms(int kind, const float *pts, int64_t n_points, int C, float *out) {
    #pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < n_points; ++p) {
        const float *x = pts + p * C;
        float acc = x[0] * x[0];
        if (kind == NORM_FMA) {
            for (int k = 1; k < C; ++k) acc = fmaf(x[k], x[k], acc);
        } else {
            for (int k = 1; k < C; ++k) { float sq = x[k] * x[k]; acc = acc + sq; }
        }
        out[p] = acc;
    }
}

/* Full matrix out[B,N,M] (materialising reference APIs; small sizes only). */
void orc_pairwise(int form, const float *row, const float *col, const float *nrow,
                  const float *ncol, int B, int N, int M, int C, float *out) {
    #pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < N; ++i) {
            const float *ri = row + ((int64_t)b * N + i) * C;
            float nr = nrow[(int64_t)b * N + i];
            float *o = out + ((int64_t)b * N + i) * M;
            for (int j = 0; j < M; ++j)
                o[j] = pair_dist(form, ri, col + ((int64_t)b * M + j) * C, C, nr,
                                 ncol[(int64_t)b * M + j]);
        }
}

/* Row minima (over j) and column minima (over i) with lowest-index argmin,
 * without materialising the matrix.  torch.min(P, dim) semantics. */
void orc_nn1(int form, const float *row, const float *col, const float *nrow,
             const float *ncol, int B, int N, int M, int C,
             float *row_min, int32_t *row_arg, float *col_min, int32_t *col_arg) {
    #pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        float *cm = col_min + (int64_t)b * M;
        int32_t *ca = col_arg + (int64_t)b * M;
        for (int j = 0; j < M; ++j) { cm[j] = INFINITY; ca[j] = 0; }
        for (int i = 0; i < N; ++i) {
            const float *ri = row + ((int64_t)b * N + i) * C;
            float nr = nrow[(int64_t)b * N + i];
            float best = INFINITY; int32_t arg = 0;
            for (int j = 0; j < M; ++j) {
                float d = pair_dist(form, ri, col + ((int64_t)b * M + j) * C, C, nr,
                                    ncol[(int64_t)b * M + j]);
                if (d < best) { best = d; arg = j; }
                if (d < cm[j]) { cm[j] = d; ca[j] = i; }
            }
            row_min[(int64_t)b * N + i] = best;
            row_arg[(int64_t)b * N + i] = arg;
        }
    }
}

/* K smallest per row, ascending by (distance, index): stable lowest-index
 * restatement of (-dist).topk(K) (attack/GeoA3/knn_utils.py:16,
 * attack/CW/CW_utils/dist_utils.py:141, model/dgcnn.py:199). */
void orc_knn(int form, const float *row, const float *col, const float *nrow,
             const float *ncol, int B, int N, int M, int C, int K,
             float *dists, int32_t *idx) {
    #pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < N; ++i) {
            const float *ri = row + ((int64_t)b * N + i) * C;
            float nr = nrow[(int64_t)b * N + i];
            float *bd = dists + ((int64_t)b * N + i) * K;
            int32_t *bi = idx + ((int64_t)b * N + i) * K;
            int cnt = 0;
            for (int j = 0; j < M; ++j) {
                float d = pair_dist(form, ri, col + ((int64_t)b * M + j) * C, C, nr,
                                    ncol[(int64_t)b * M + j]);
                if (cnt == K && !(d < bd[K - 1])) continue;   /* ties keep the earlier index */
                int p = (cnt < K) ? cnt : K - 1;
                while (p > 0 && d < bd[p - 1]) { bd[p] = bd[p - 1]; bi[p] = bi[p - 1]; --p; }
                bd[p] = d; bi[p] = j;
                if (cnt < K) ++cnt;
            }
        }
}

/* query_ball_point (model/pointnet2_utils.py:84-104): first nsample indices j
 * (ascending) with NOT(sqrdist > r2), padded with the first hit; a row without
 * any hit is filled with N (what the reference's sort leaves behind).
 * rows = new_xyz[B,S,3] (src), cols = xyz[B,N,3] (dst), FORM_ROW_COL. */
void orc_ball_query(const float *row, const float *col, const float *nrow,
                    const float *ncol, int B, int S, int N, int C, float r2,
                    int nsample, int32_t *idx) {
    #pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < S; ++i) {
            const float *ri = row + ((int64_t)b * S + i) * C;
            float nr = nrow[(int64_t)b * S + i];
            int32_t *o = idx + ((int64_t)b * S + i) * nsample;
            int cnt = 0;
            for (int j = 0; j < N && cnt < nsample; ++j) {
                float d = pair_dist(FORM_ROW_COL, ri, col + ((int64_t)b * N + j) * C, C, nr,
                                    ncol[(int64_t)b * N + j]);
                if (!(d > r2)) o[cnt++] = j;
            }
            int32_t first = cnt ? o[0] : N;
            for (int s = cnt; s < nsample; ++s) o[s] = first;
        }
}

/* farthest_point_sample (model/pointnet2_utils.py:59-81; model/curvenet_util.py:69-90 with
 * start index 0): distance = 1e10; per iteration: record farthest, dist = sum((xyz-c)^2,-1)
 * = ((dx*dx + dy*dy) + dz*dz) (torch.sum over the contiguous last dim of 3, sequential),
 * distance = where(dist < distance, dist, distance), farthest = FIRST index of max(distance).
 * xyz[B,N,3] point-major, start[B] (NULL = 0), out[B,npoint]. */
void orc_fps(const float *xyz, int B, int N, int npoint, const int32_t *start, int32_t *out) {
    #pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b) {
        const float *p = xyz + (int64_t)b * N * 3;
        float *distance = (float *)malloc(sizeof(float) * (size_t)N);
        for (int i = 0; i < N; ++i) distance[i] = 1e10f;
        int far = start ? start[b] : 0;
        for (int it = 0; it < npoint; ++it) {
            out[(int64_t)b * npoint + it] = far;
            float cx = p[far * 3], cy = p[far * 3 + 1], cz = p[far * 3 + 2];
            float best = -1.0f; int bi = 0;
            for (int i = 0; i < N; ++i) {
                float dx = p[i * 3] - cx, dy = p[i * 3 + 1] - cy, dz = p[i * 3 + 2] - cz;
                float s = dx * dx; s = s + dy * dy; s = s + dz * dz;
                if (s < distance[i]) distance[i] = s;
                if (distance[i] > best) { best = distance[i]; bi = i; }
            }
            far = bi;
        }
        free(distance);
    }
}
