// pcd_geom.cu -- local geometry on a k-NN graph, sm_100a.
//
// The consumers of the k-NN select in GeoA3 / AOF, each fused into ONE pass over the index tensor so
// that the [B,N,K,3] neighbour gather the reference materialises never reaches HBM:
//   local_frames     attack/GeoA3/utility.py:43-92 (estimate_normal), :119-152 (estimate_perpendicular):
//                    per point the 3x3 covariance of its k neighbours and its eigen-frame.  The
//                    reference calls torch.symeig, which no longer exists; here the covariance is
//                    formed in fp32 exactly as the reference forms it (mean, centring, bmm, 1/(k-1))
//                    and diagonalised by a cyclic Jacobi iteration in fp64 registers.
//   kappa            attack/GeoA3/loss_utils.py:60-90, :116-125: mean_j |<unit(q_j - p_i), n_i>| with
//                    its backward (own term + scatter through the neighbour indices).
//   graph_laplacian  attack/AOF/TAOF_attack.py:31-52: L = D - A, A_ij = exp(-|p_i - p_j|^2) on the
//                    symmetrised k-NN graph, dense [B,N,N] as the reference's eigensolver wants it.
// All of it is HBM / latency bound index work: one thread (frames, kappa) or one warp (Laplacian
// rows) per point, coalesced index reads, no shared memory.
#include "pcd_common.cuh"

namespace pcd {

struct Cloud {
    const float *p;
    long long sb, sp, sc;
};
__device__ __forceinline__ float3 ld_point(const Cloud &c, int b, int i) {
    const float *s = c.p + (size_t)b * c.sb + (size_t)i * c.sp;
    return make_float3(__ldg(s), __ldg(s + c.sc), __ldg(s + 2 * c.sc));
}

// ------------------------------------------------------------------ symmetric 3x3 eigen-frame
// Cyclic Jacobi in fp64: a (symmetric, destroyed) -> eigenvalues w, eigenvectors as the COLUMNS of v.
// Six sweeps bring a 3x3 matrix to machine precision (quadratic convergence); fixed count, no data
// dependent exit, so every thread of a warp runs the same instruction stream.
__device__ __forceinline__ void jacobi_rotate(double a[3][3], double v[3][3], int p, int q) {
    const double apq = a[p][q];
    if (fabs(apq) < 1e-300) return;
    const double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
    const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
    const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
    const int r = 3 - p - q;
    const double arp = a[r][p], arq = a[r][q];
    a[p][p] -= t * apq;
    a[q][q] += t * apq;
    a[p][q] = a[q][p] = 0.0;
    a[r][p] = a[p][r] = c * arp - s * arq;
    a[r][q] = a[q][r] = s * arp + c * arq;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double vkp = v[k][p], vkq = v[k][q];
        v[k][p] = c * vkp - s * vkq;
        v[k][q] = s * vkp + c * vkq;
    }
}
__device__ __forceinline__ void eigen_sym3(double a[3][3], double v[3][3], double w[3]) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) v[i][j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 6; ++sweep) {
        jacobi_rotate(a, v, 0, 1);
        jacobi_rotate(a, v, 0, 2);
        jacobi_rotate(a, v, 1, 2);
    }
    w[0] = a[0][0]; w[1] = a[1][1]; w[2] = a[2][2];
}

// One thread per point.  idx[b,i,first..K1) are the neighbours (first = 1 drops the self column the
// reference slices away, utility.py:50).  Outputs (each optional):
//   normal[b,i,:]   eigenvector of the smallest eigenvalue with the reference's sign rule
//                   -sign(<n, sum of the centred neighbours>) (utility.py:67-69; the sum is rounding
//                   noise around 0, so the sign is as arbitrary as the reference's own)
//   evecs[b,i,r,:]  r = 0,1,2: eigenvectors by ascending eigenvalue; evals[b,i,r] the eigenvalues
__global__ void __launch_bounds__(128) local_frames_kernel(Cloud pc, const int32_t *__restrict__ idx, int B, int N, int K1, int first,
                                                           float *__restrict__ normal, long long n_sb, long long n_sp, long long n_sc,
                                                           float *__restrict__ evecs, float *__restrict__ evals) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= (long long)B * N) return;
    const int b = (int)(t / N), i = (int)(t - (long long)b * N);
    const int32_t *nb = idx + (size_t)t * K1;
    const int k = K1 - first;
    // mean of the neighbours (torch.mean: sequential sum, then the division)
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (int j = first; j < K1; ++j) {
        const float3 q = ld_point(pc, b, min(max(nb[j], 0), N - 1));
        sx = __fadd_rn(sx, q.x); sy = __fadd_rn(sy, q.y); sz = __fadd_rn(sz, q.z);
    }
    const float mx = __fdiv_rn(sx, (float)k), my = __fdiv_rn(sy, (float)k), mz = __fdiv_rn(sz, (float)k);
    // covariance = fact * C C^T in fp32 (bmm: k ascending FMA chain), neighbour sum of the centred set
    float cxx = 0.f, cxy = 0.f, cxz = 0.f, cyy = 0.f, cyz = 0.f, czz = 0.f, nsx = 0.f, nsy = 0.f, nsz = 0.f;
    for (int j = first; j < K1; ++j) {
        const float3 q = ld_point(pc, b, min(max(nb[j], 0), N - 1));
        const float dx = __fsub_rn(q.x, mx), dy = __fsub_rn(q.y, my), dz = __fsub_rn(q.z, mz);
        cxx = __fmaf_rn(dx, dx, cxx); cxy = __fmaf_rn(dx, dy, cxy); cxz = __fmaf_rn(dx, dz, cxz);
        cyy = __fmaf_rn(dy, dy, cyy); cyz = __fmaf_rn(dy, dz, cyz); czz = __fmaf_rn(dz, dz, czz);
        nsx = __fadd_rn(nsx, dx); nsy = __fadd_rn(nsy, dy); nsz = __fadd_rn(nsz, dz);
    }
    const float fact = 1.0f / (float)(k - 1);
    double a[3][3], v[3][3], w[3];
    a[0][0] = (double)__fmul_rn(fact, cxx); a[0][1] = a[1][0] = (double)__fmul_rn(fact, cxy); a[0][2] = a[2][0] = (double)__fmul_rn(fact, cxz);
    a[1][1] = (double)__fmul_rn(fact, cyy); a[1][2] = a[2][1] = (double)__fmul_rn(fact, cyz); a[2][2] = (double)__fmul_rn(fact, czz);
    eigen_sym3(a, v, w);
    // ascending order of the eigenvalues (first index wins ties, like torch.argmin)
    int o0 = 0, o1 = 1, o2 = 2;
    if (w[o1] < w[o0]) { const int s = o0; o0 = o1; o1 = s; }
    if (w[o2] < w[o0]) { const int s = o0; o0 = o2; o2 = s; }
    if (w[o2] < w[o1]) { const int s = o1; o1 = o2; o2 = s; }
    if (normal) {
        float nx = (float)v[0][o0], ny = (float)v[1][o0], nz = (float)v[2][o0];
        const float d = __fmaf_rn(nz, nsz, __fmaf_rn(ny, nsy, __fmul_rn(nx, nsx)));
        const float sgn = d > 0.f ? -1.f : (d < 0.f ? 1.f : 0.f);          // -torch.sign(.)
        float *o = normal + (size_t)b * n_sb + (size_t)i * n_sp;
        o[0] = sgn * nx; o[n_sc] = sgn * ny; o[2 * n_sc] = sgn * nz;
    }
    if (evecs) {
        float *o = evecs + (size_t)t * 9;
        const int ord[3] = {o0, o1, o2};
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            o[3 * r] = (float)v[0][ord[r]]; o[3 * r + 1] = (float)v[1][ord[r]]; o[3 * r + 2] = (float)v[2][ord[r]];
        }
    }
    if (evals) {
        float *o = evals + (size_t)t * 3;
        o[0] = (float)w[o0]; o[1] = (float)w[o1]; o[2] = (float)w[o2];
    }
}

// ---------------------------------------------------------------------------------- kappa
// kappa[b,i] = mean_j |<unit(q_j - p_i), n>|, unit(d) = d / max(|d|, eps) (utility.py:_normalize, eps 1e-12),
// n = normal[b, nidx ? nidx[b,i] : i].  Arithmetic as the reference's torch chain: d = q - p,
// |d| = sqrt((dx^2 + dy^2) + dz^2), u = d / max(|d|, eps), dot = (ux nx + uy ny) + uz nz, mean = sum / k.
struct KappaArgs {
    Cloud pc, normal;
    const int32_t *idx, *nidx;
    int B, N, K1, first;
};
__device__ __forceinline__ float3 kappa_normal(const KappaArgs &a, int b, int i, size_t t) {
    int ni = a.nidx ? a.nidx[t] : i;
    ni = min(max(ni, 0), a.N - 1);
    return ld_point(a.normal, b, ni);
}

__global__ void __launch_bounds__(128) kappa_fwd_kernel(KappaArgs a, float *__restrict__ kappa) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= (long long)a.B * a.N) return;
    const int b = (int)(t / a.N), i = (int)(t - (long long)b * a.N);
    const float3 p = ld_point(a.pc, b, i), n = kappa_normal(a, b, i, (size_t)t);
    const int32_t *nb = a.idx + (size_t)t * a.K1;
    float acc = 0.f;
    for (int j = a.first; j < a.K1; ++j) {
        const float3 q = ld_point(a.pc, b, min(max(nb[j], 0), a.N - 1));
        const float dx = __fsub_rn(q.x, p.x), dy = __fsub_rn(q.y, p.y), dz = __fsub_rn(q.z, p.z);
        const float r = fmaxf(__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz))), 1e-12f);
        const float ux = __fdiv_rn(dx, r), uy = __fdiv_rn(dy, r), uz = __fdiv_rn(dz, r);
        acc = __fadd_rn(acc, fabsf(__fadd_rn(__fadd_rn(__fmul_rn(ux, n.x), __fmul_rn(uy, n.y)), __fmul_rn(uz, n.z))));
    }
    kappa[t] = __fdiv_rn(acc, (float)(a.K1 - a.first));
}

// d kappa_i / d(d_ij) = sign(s) (n - s u) / (k r) with u = d / r, s = <u, n>   (r >= eps; the clamp passes no gradient
// below eps, where u = d / eps and the term is sign(s) n / (k eps)); p_i receives the negative sum, q_j the term itself.
__global__ void __launch_bounds__(128) kappa_bwd_kernel(KappaArgs a, const float *__restrict__ g, float *__restrict__ grad) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= (long long)a.B * a.N) return;
    const int b = (int)(t / a.N), i = (int)(t - (long long)b * a.N);
    const float gi = g[t] / (float)(a.K1 - a.first);
    if (gi == 0.f) return;
    const float3 p = ld_point(a.pc, b, i), n = kappa_normal(a, b, i, (size_t)t);
    const int32_t *nb = a.idx + (size_t)t * a.K1;
    float ox = 0.f, oy = 0.f, oz = 0.f;
    for (int j = a.first; j < a.K1; ++j) {
        const int qi = min(max(nb[j], 0), a.N - 1);
        const float3 q = ld_point(a.pc, b, qi);
        const float dx = q.x - p.x, dy = q.y - p.y, dz = q.z - p.z;
        const float r2 = dx * dx + dy * dy + dz * dz;
        const float r = sqrtf(r2);
        float tx, ty, tz;
        if (r >= 1e-12f) {
            const float inv = 1.f / r, ux = dx * inv, uy = dy * inv, uz = dz * inv;
            const float s = ux * n.x + uy * n.y + uz * n.z;
            const float sg = s > 0.f ? gi : (s < 0.f ? -gi : 0.f);
            tx = sg * (n.x - s * ux) * inv; ty = sg * (n.y - s * uy) * inv; tz = sg * (n.z - s * uz) * inv;
        } else {
            const float s = (dx * n.x + dy * n.y + dz * n.z) * 1e12f;
            const float sg = (s > 0.f ? gi : (s < 0.f ? -gi : 0.f)) * 1e12f;
            tx = sg * n.x; ty = sg * n.y; tz = sg * n.z;
        }
        ox -= tx; oy -= ty; oz -= tz;
        float *gq = grad + ((size_t)b * a.N + qi) * 3;
        atomicAdd(gq, tx); atomicAdd(gq + 1, ty); atomicAdd(gq + 2, tz);
    }
    float *gp = grad + (size_t)t * 3;
    atomicAdd(gp, ox); atomicAdd(gp + 1, oy); atomicAdd(gp + 2, oz);
}

// ------------------------------------------------------------------------- graph Laplacian
// L[b] = D - A on the symmetrised k-NN graph: A_ij = exp(-|p_i - p_j|^2) if j in kNN(i) or i in kNN(j), else 0
// (the self loop of the k-NN list gives A_ii = 1, which cancels in D - A exactly as in the reference).
// Pass 1 (after a memset): every edge (i, j) writes -a to (i,j) and (j,i) -- both directions compute
// bit-identical values, so duplicate writes are benign.  Pass 2: one warp per row sums the row in a
// fixed order and sets the diagonal to -(sum of the off-diagonal entries) ... i.e. D_ii - A_ii.
__global__ void __launch_bounds__(256) laplacian_edges_kernel(Cloud pc, const int32_t *__restrict__ idx, int B, int N, int K,
                                                              float *__restrict__ L) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= (long long)B * N * K) return;
    const long long bi = t / K;
    const int b = (int)(bi / N), i = (int)(bi - (long long)b * N);
    const int j = min(max(idx[t], 0), N - 1);
    const float3 p = ld_point(pc, b, i), q = ld_point(pc, b, j);
    const float dx = __fsub_rn(p.x, q.x), dy = __fsub_rn(p.y, q.y), dz = __fsub_rn(p.z, q.z);
    const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    const float a = -expf(-d2);
    float *Lb = L + (size_t)b * N * N;
    Lb[(size_t)i * N + j] = a;
    Lb[(size_t)j * N + i] = a;
}
__global__ void __launch_bounds__(256) laplacian_diag_kernel(int B, int N, float *__restrict__ L) {
    const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= (long long)B * N) return;
    const int i = (int)(row % N);
    float *Lr = L + (size_t)row * N;
    float s = 0.f;
    for (int j = lane; j < N; j += 32) s += (j == i) ? 0.f : Lr[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) Lr[i] = -s;            // D_ii - A_ii = sum_{j != i} A_ij
}

}  // namespace pcd

using namespace pcd;

static bool bad_knn_args(const void *pc, const void *idx, int B, int N, int K1, int first) {
    return !pc || !idx || B <= 0 || N <= 0 || K1 <= 0 || first < 0 || first >= K1;
}

extern "C" int pcd_local_frames(const float *pc, int64_t sb, int64_t sp, int64_t sc, const int32_t *idx, int B, int N, int K1,
                                int skip_first, float *normal, int64_t n_sb, int64_t n_sp, int64_t n_sc, float *evecs, float *evals,
                                void *stream) {
    if (bad_knn_args(pc, idx, B, N, K1, skip_first ? 1 : 0) || (!normal && !evecs && !evals)) {
        set_error("pcd_local_frames: bad argument");
        return PCD_ERR_ARG;
    }
    const long long total = (long long)B * N;
    local_frames_kernel<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        Cloud{pc, sb, sp, sc}, idx, B, N, K1, skip_first ? 1 : 0, normal, n_sb, n_sp, n_sc, evecs, evals);
    PCD_CUDA_CHECK(cudaGetLastError());
    return PCD_OK;
}

extern "C" int pcd_kappa_forward(const float *pc, int64_t sb, int64_t sp, int64_t sc, const float *normal, int64_t n_sb, int64_t n_sp,
                                 int64_t n_sc, const int32_t *nidx, const int32_t *idx, int B, int N, int K1, int skip_first,
                                 float *kappa, void *stream) {
    if (bad_knn_args(pc, idx, B, N, K1, skip_first ? 1 : 0) || !normal || !kappa) {
        set_error("pcd_kappa_forward: bad argument");
        return PCD_ERR_ARG;
    }
    KappaArgs a{Cloud{pc, sb, sp, sc}, Cloud{normal, n_sb, n_sp, n_sc}, idx, nidx, B, N, K1, skip_first ? 1 : 0};
    const long long total = (long long)B * N;
    kappa_fwd_kernel<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a, kappa);
    PCD_CUDA_CHECK(cudaGetLastError());
    return PCD_OK;
}

extern "C" int pcd_kappa_backward(const float *pc, int64_t sb, int64_t sp, int64_t sc, const float *normal, int64_t n_sb, int64_t n_sp,
                                  int64_t n_sc, const int32_t *nidx, const int32_t *idx, int B, int N, int K1, int skip_first,
                                  const float *g_kappa, float *grad_pc, void *stream) {
    if (bad_knn_args(pc, idx, B, N, K1, skip_first ? 1 : 0) || !normal || !g_kappa || !grad_pc) {
        set_error("pcd_kappa_backward: bad argument");
        return PCD_ERR_ARG;
    }
    KappaArgs a{Cloud{pc, sb, sp, sc}, Cloud{normal, n_sb, n_sp, n_sc}, idx, nidx, B, N, K1, skip_first ? 1 : 0};
    const long long total = (long long)B * N;
    cudaStream_t st = (cudaStream_t)stream;
    PCD_CUDA_CHECK(cudaMemsetAsync(grad_pc, 0, (size_t)total * 3 * sizeof(float), st));
    kappa_bwd_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(a, g_kappa, grad_pc);
    PCD_CUDA_CHECK(cudaGetLastError());
    return PCD_OK;
}

extern "C" int pcd_graph_laplacian(const float *pc, int64_t sb, int64_t sp, int64_t sc, const int32_t *idx, int B, int N, int K,
                                   float *L, void *stream) {
    if (!pc || !idx || !L || B <= 0 || N <= 0 || K <= 0) {
        set_error("pcd_graph_laplacian: bad argument");
        return PCD_ERR_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    PCD_CUDA_CHECK(cudaMemsetAsync(L, 0, (size_t)B * N * N * sizeof(float), st));
    const long long edges = (long long)B * N * K;
    laplacian_edges_kernel<<<(unsigned)((edges + 255) / 256), 256, 0, st>>>(Cloud{pc, sb, sp, sc}, idx, B, N, K, L);
    PCD_CUDA_CHECK(cudaGetLastError());
    const long long rows = (long long)B * N;
    laplacian_diag_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, st>>>(B, N, L);
    PCD_CUDA_CHECK(cudaGetLastError());
    return PCD_OK;
}
