// pcd_nn1.cu -- NN-1 sweep (Chamfer / Hausdorff / knn_points K=1) for sm_100a.
//
// Pipeline of pcd_nn1_forward (4 launches on the caller's stream, no host sync):
//   1. nn1_prep    pack both clouds into the sweep layout, compute |p|^2 in the reference's
//                  rounding order, reset the (value,tag) keys.
//   2. nn1_sweep   THE hot kernel: every (row i, col j) distance exactly once; row minima and
//                  column minima are both taken from the same tile sweep.  Packed fp32x2
//                  math (FMUL2/FFMA2/FADD2), FMNMX3 running minima, CREDUX warp minima,
//                  column tiles streamed by 1-D TMA bulk copies behind an mbarrier, stream-K
//                  partition of (sample, row tile, col tile) units over a persistent grid.
//                  Emits per point a 64-bit key = (ordered min value, tag of the 32-wide /
//                  32R-wide chunk that produced it) with atomicMin.
//   3. nn1_fixup   eight lanes per point re-evaluate its winning chunk (bit-identical
//                  arithmetic) for the lowest index with d == min.
//   4. nn1_reduce  per-sample scaled sum / max / first-argmax of both minima arrays (fixed order).
//
// Reference semantics served: utils/dis_utils_torch.py:8-28, attack/CW/CW_utils/distance.py:15-70,
// attack/GeoA3/knn_utils.py:10-55 (K=1); see include/pcdist.h.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "pcd_common.cuh"

namespace pcd {

// ------------------------------------------------------------------------------ error state
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int cuda_fail(cudaError_t e, const char *what) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return PCD_ERR_CUDA;
}

// ---------------------------------------------------------------------------------- layout
constexpr int kSweepWarps = 4;
constexpr int kSweepThreads = kSweepWarps * 32;
constexpr int kColChunk = 32;     // columns per row-direction tag
constexpr int kMaxColTile = 256;  // columns per TMA stage (16 B each)
constexpr int kRowPadUnit = 2048; // rows are padded to a multiple of 128*R, R <= 16

struct Nn1Layout {
    int Npad, Mpad;
    size_t rowpk, rowpp, colpk, rowkey, colkey, total;
};
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static Nn1Layout nn1_layout(int B, int N, int M) {
    Nn1Layout L;
    L.Npad = (int)align_up((size_t)N, kRowPadUnit);
    L.Mpad = (int)align_up((size_t)M, kMaxColTile);
    size_t off = 0;
    L.rowpk = off; off = align_up(off + (size_t)B * L.Npad * 16, 256);
    L.rowpp = off; off = align_up(off + (size_t)B * L.Npad * 16, 256);   // the same records in sweep (slot) order
    L.colpk = off; off = align_up(off + (size_t)B * L.Mpad * 16 + 64, 256);   // +64: the sweep prefetches one record past a tile
    L.rowkey = off; off = align_up(off + (size_t)B * L.Npad * 8, 256);
    L.colkey = off; off = align_up(off + (size_t)B * L.Mpad * 8, 256);
    L.total = off;
    return L;
}

// ------------------------------------------------------------------------------------ prep
// Row key slots.  Lane l of a sweep warp owns the R consecutive rows i = blk*32R + l*R + r; the
// key of row i lives at slot blk*32R + r*32 + l, so that the warp's atomicMin flushes are
// coalesced (one 256-B run per r instead of 32 scattered lines: the scattered flush cost 1.8 us
// per row tile, 8 us of 131 at BASELINE config 2).  The records exist twice: rowpp in slot order
// for the sweep's coalesced loads (another 4 us), rowpk in row order for the fix-up.
__host__ __device__ __forceinline__ int row_slot(int i, int R) {      // R is 2, 4, 8 or 16: shifts, no divisions
    const int rs = R == 16 ? 4 : (R == 8 ? 3 : (R == 4 ? 2 : 1));
    const int w = i & ((32 << rs) - 1);
    return (i - w) + ((w & (R - 1)) << 5) + (w >> rs);
}

// rowpk[b][i] = float4(-2x, -2y, -2z, nrow)           (AoS, one LDG.128 per query)
// colpk[b][j/2] = {x0,x1,y0,y1,z0,z1,n0,n1}           (pair records: two LDS.128 feed 2 columns
//                                                      as ready-made fp32x2 operands)
// Padded points are inert: coordinates 0, norm +inf  => every distance through them is +inf.
__global__ void nn1_prep_kernel(const float *__restrict__ rows, int64_t r_sb, int64_t r_sp, int64_t r_sc,
                                const float *__restrict__ cols, int64_t c_sb, int64_t c_sp, int64_t c_sc,
                                int B, int N, int M, int Npad, int Mpad, int R, int norm_kind, int swap_norms,
                                float4 *__restrict__ rowpk, float4 *__restrict__ rowpp, float *__restrict__ colpk,
                                unsigned long long *__restrict__ rowkey, unsigned long long *__restrict__ colkey) {
    const long long per_b = (long long)Npad + Mpad;
    const long long total = per_b * B;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(t / per_b);
        const int p = (int)(t - (long long)b * per_b);
        if (p < Npad) {
            const int i = p;
            float x = 0.f, y = 0.f, z = 0.f, n = __int_as_float(0x7f800000);
            if (i < N) {
                const float *s = rows + b * r_sb + i * r_sp;
                x = s[0]; y = s[r_sc]; z = s[2 * r_sc];
                if (swap_norms) {
                    const float *o = cols + b * c_sb + i * c_sp;
                    n = sq_norm3(norm_kind, o[0], o[c_sc], o[2 * c_sc]);
                } else {
                    n = sq_norm3(norm_kind, x, y, z);
                }
            }
            const float4 rec = make_float4(-2.f * x, -2.f * y, -2.f * z, n);
            rowpk[(size_t)b * Npad + i] = rec;
            rowpp[(size_t)b * Npad + row_slot(i, R)] = rec;
            rowkey[(size_t)b * Npad + i] = ~0ull;
        } else {
            const int j = p - Npad;
            float x = 0.f, y = 0.f, z = 0.f, n = __int_as_float(0x7f800000);
            if (j < M) {
                const float *s = cols + b * c_sb + j * c_sp;
                x = s[0]; y = s[c_sc]; z = s[2 * c_sc];
                if (swap_norms) {
                    const float *o = rows + b * r_sb + j * r_sp;
                    n = sq_norm3(norm_kind, o[0], o[r_sc], o[2 * r_sc]);
                } else {
                    n = sq_norm3(norm_kind, x, y, z);
                }
            }
            float *rec = colpk + ((size_t)b * Mpad + (j & ~1)) * 4 + (j & 1);
            rec[0] = x; rec[2] = y; rec[4] = z; rec[6] = n;
            colkey[(size_t)b * Mpad + j] = ~0ull;
        }
    }
}

// ----------------------------------------------------------------------------------- sweep
// Lane l of warp w of the CTA working on row tile qt owns the R consecutive rows
//     i = qt*QT + w*32R + l*R + r,  r < R
// Row direction: running minimum per row over 32-column chunks, tag = chunk index.
// Column direction: per column the lane folds its R rows (FMNMX3), CREDUX gives the warp
// minimum, FSETP+VOTE the ballot of the lanes that hold it; (value, ballot) goes to shared
// memory and the per-tile flush turns the lowest set lane into the tag
//     tag = (qt*4 + w)*32 + lane      ->  the R rows of that lane.
struct SweepSmem {
    float4 tile[2][kMaxColTile + 2];                 // column pair-records (TMA destination) + prefetch pad
    uint2 colpart[2][kSweepWarps][kMaxColTile];      // per-warp (column minimum, lane ballot)
    uint64_t full[2];                                // mbarriers: tile[s] has landed
};

constexpr int kQuad = 4;   // columns per scheduling unit (two packed steps)

#ifdef PCD_SWEEP_TRACE
// development build only (tools/trace_sweep.py): per-CTA timestamps of the sweep
__device__ unsigned long long *g_sweep_trace = nullptr;
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#endif

template <int FORM, int R>
__global__ void __launch_bounds__(kSweepThreads, (R >= 16) ? 2 : ((R >= 8) ? 4 : ((R >= 4) ? 5 : 6)))
nn1_sweep_kernel(const float4 *__restrict__ rowpp /* slot order */, const float4 *__restrict__ colpk,
                 unsigned long long *__restrict__ rowkey, unsigned long long *__restrict__ colkey,
                 int Npad, int Mpad, int qpt /* quads per TMA tile */, int nqt, int nq /* quads per row tile */,
                 int units) {
    constexpr int QW = 32 * R;            // rows per warp
    constexpr int QT = kSweepWarps * QW;  // rows per CTA tile
    constexpr int kQuadsPerChunk = kColChunk / kQuad;
    __shared__ __align__(128) SweepSmem sm;

    // Stream-K at 4-column granularity: the (sample, row tile, column quad) space is cut into
    // gridDim.x equal contiguous ranges, so every CTA sweeps the same number of pairs (+-1 quad).
    // Ranges may start or end inside a 32-column chunk: a row tag only says "the minimum is in
    // this chunk", and the fix-up rescans the whole chunk, so partial chunks stay exact.
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifdef PCD_SWEEP_TRACE
    unsigned long long *trace = g_sweep_trace;
    if (trace && tid == 0) {
        unsigned int smid;
        asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        trace[blockIdx.x * 8 + 0] = globaltimer_ns();
        trace[blockIdx.x * 8 + 3] = smid;
    }
#endif
    const int u0 = (int)((long long)units * blockIdx.x / gridDim.x);
    const int u1 = (int)((long long)units * (blockIdx.x + 1) / gridDim.x);
    if (u0 >= u1) return;

    // a segment = the quads [u, u+n) that share one row tile and one qpt-aligned column tile
    auto seg_len = [&](int u) -> int {
        const int q = u % nq;
        int e = (q / qpt + 1) * qpt;
        if (e > nq) e = nq;
        const int n = e - q;
        return n < u1 - u ? n : u1 - u;
    };
    int pu = u0;   // producer cursor (thread 0)
    auto issue = [&](int buf) {
        if (pu >= u1) return;
        const int bq = pu / nq, q = pu - bq * nq, b = bq / nqt;
        const int n = seg_len(pu);
        const uint32_t bytes = (uint32_t)n * kQuad * 16u;
        mbar_expect_tx(&sm.full[buf], bytes);
        tma_load_1d(sm.tile[buf], colpk + (size_t)b * Mpad + (size_t)q * kQuad, bytes, &sm.full[buf]);
        pu += n;
    };
    if (tid == 0) {
        mbar_init(&sm.full[0], 1);
        mbar_init(&sm.full[1], 1);
        fence_mbar_init();
        fence_proxy_async();
        issue(0);
        issue(1);
    }
    __syncthreads();

    float qx[R], qy[R], qz[R], qn[R], best[R];
    uint32_t btag[R];
    int cur_bq = -1;
    size_t row_base = 0;

    int it = 0;
    for (int u = u0; u < u1; ++it) {
        const int buf = it & 1;
        const uint32_t parity = (it >> 1) & 1;
        const int bq = u / nq, q0 = u - bq * nq;
        const int b = bq / nqt, qt = bq - b * nqt;
        const int nseg = seg_len(u);

        if (bq != cur_bq) {
            if (cur_bq >= 0) {
#pragma unroll
                for (int r = 0; r < R; ++r) atomicMin(&rowkey[row_base + r * 32], make_key(best[r], btag[r]));
            }
            cur_bq = bq;
            row_base = (size_t)b * Npad + (size_t)qt * QT + warp * QW + lane;    // key slot of row r: + r*32
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float4 q = __ldg(&rowpp[row_base + r * 32]);
                qx[r] = q.x; qy[r] = q.y; qz[r] = q.z; qn[r] = q.w;
                best[r] = __int_as_float(0x7f800000);
                btag[r] = 0;
            }
        }

        mbar_wait(&sm.full[buf], parity);
#ifdef PCD_SWEEP_TRACE
        if (trace && tid == 0 && it == 0) trace[blockIdx.x * 8 + 4] = globaltimer_ns();
#endif

        const float4 *t4 = sm.tile[buf];
        uint2 *cp = sm.colpart[buf][warp];
        float4 A = t4[0], Bv = t4[1];                       // operands of the step about to run
        uint2 pend_lo = make_uint2(0x7f800000u, 0u), pend_hi = pend_lo;   // column results of the previous step
        int pend_at = -1;
        int qd = 0;                                          // quad cursor inside the segment
        while (qd < nseg) {
            // piece = the quads of this segment that fall into one 32-column chunk
            const int chunk = (q0 + qd) / kQuadsPerChunk;
            int pe = (chunk + 1) * kQuadsPerChunk - q0;
            if (pe > nseg) pe = nseg;
            float m[R];
#pragma unroll
            for (int r = 0; r < R; ++r) m[r] = __int_as_float(0x7f800000);
            for (; qd < pe; ++qd) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int step = qd * 2 + h;
                    const float4 An = t4[step * 2 + 2], Bn = t4[step * 2 + 3];   // prefetch (pad keeps it in bounds)
                    const f32x2 X = pack2(A.x, A.y), Y = pack2(A.z, A.w);
                    const f32x2 Z = pack2(Bv.x, Bv.y), Nn = pack2(Bv.z, Bv.w);
                    float lo[R], hi[R];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const f32x2 d = pair_dist_x2<FORM>(qx[r], qy[r], qz[r], qn[r], X, Y, Z, Nn);
                        unpack2(d, lo[r], hi[r]);
                        m[r] = min3(m[r], lo[r], hi[r]);
                    }
                    float clo = lo[0], chi = hi[0];
#pragma unroll
                    for (int r = 1; r + 1 < R; r += 2) {
                        clo = min3(clo, lo[r], lo[r + 1]);
                        chi = min3(chi, hi[r], hi[r + 1]);
                    }
                    if ((R & 1) == 0) {
                        clo = fminf(clo, lo[R - 1]);
                        chi = fminf(chi, hi[R - 1]);
                    }
                    // retire the previous step's column results (their CREDUX latency is long gone)
                    if (pend_at >= 0) *reinterpret_cast<uint4 *>(&cp[pend_at]) = make_uint4(pend_lo.x, pend_lo.y, pend_hi.x, pend_hi.y);
                    const float vlo = warp_min_f32(clo), vhi = warp_min_f32(chi);
                    pend_lo = make_uint2(__float_as_uint(vlo), __ballot_sync(0xffffffffu, clo == vlo));
                    pend_hi = make_uint2(__float_as_uint(vhi), __ballot_sync(0xffffffffu, chi == vhi));
                    pend_at = 2 * step;
                    A = An; Bv = Bn;
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (m[r] < best[r]) {
                    best[r] = m[r];
                    btag[r] = (uint32_t)chunk;
                }
            }
        }
        *reinterpret_cast<uint4 *>(&cp[pend_at]) = make_uint4(pend_lo.x, pend_lo.y, pend_hi.x, pend_hi.y);
        __syncthreads();  // tile[buf] fully read, colpart[buf] fully written

        if (tid == 0) issue(buf);
        // column flush: min over the CTA's warps (lowest warp on ties), lowest lane of its ballot
        const int ncols = nseg * kQuad;
        for (int col = tid; col < ncols; col += kSweepThreads) {
            uint2 e = sm.colpart[buf][0][col];
            float v = __uint_as_float(e.x);
            uint32_t w = 0, msk = e.y;
#pragma unroll
            for (int k = 1; k < kSweepWarps; ++k) {
                const uint2 o = sm.colpart[buf][k][col];
                if (__uint_as_float(o.x) < v) { v = __uint_as_float(o.x); w = k; msk = o.y; }
            }
            if (v < __int_as_float(0x7f800000))
                atomicMin(&colkey[(size_t)b * Mpad + (size_t)q0 * kQuad + col],
                          make_key(v, (((uint32_t)qt * kSweepWarps + w) << 5) + (uint32_t)(__ffs(msk) - 1)));
        }
        u += nseg;
    }
#ifdef PCD_SWEEP_TRACE
    if (trace && tid == 0) trace[blockIdx.x * 8 + 1] = globaltimer_ns();
#endif
#pragma unroll
    for (int r = 0; r < R; ++r) atomicMin(&rowkey[row_base + r * 32], make_key(best[r], btag[r]));
#ifdef PCD_SWEEP_TRACE
    __syncthreads();
    if (trace && tid == 0) trace[blockIdx.x * 8 + 2] = globaltimer_ns();
#endif
}

// --------------------------------------------------------------------- fix-up + reduction
__device__ __forceinline__ float apply_transform(int transform, float v) {
    return transform == PCD_VALUE_SQRT_CLAMP ? sqrtf(fmaxf(v, 0.0f)) : v;
}

// Four lanes per point (eight independent 16-byte loads per lane; 2x the points in flight of the
// eight-lane version, which was bound by the key -> chunk load round trips).  A row point
// re-evaluates the 32 columns of its winning chunk, a column point the R rows of its winning
// lane -- with the sweep's exact arithmetic -- and takes the lowest index whose distance equals
// the minimum.  The per-sample statistics (sum, max, first argmax) are reduced by
// nn1_reduce_kernel in a fixed order, so they are run-to-run deterministic.
__device__ __forceinline__ void reduce_smf(float &s, float &mx, int &am) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        const float om = __shfl_xor_sync(0xffffffffu, mx, o);
        const int oa = __shfl_xor_sync(0xffffffffu, am, o);
        if (om > mx || (om == mx && oa < am)) { mx = om; am = oa; }
    }
}

template <int FORM>
__global__ void __launch_bounds__(256, 8)        // 64 warps per SM: the key -> chunk load round trips are the bound
nn1_fixup_kernel(const float4 *__restrict__ rowpk, const float4 *__restrict__ colpk,
                 const unsigned long long *__restrict__ rowkey, const unsigned long long *__restrict__ colkey,
                 int N, int M, int Npad, int Mpad, int R, int transform,
                 float *__restrict__ row_min, int32_t *__restrict__ row_arg,
                 float *__restrict__ col_min, int32_t *__restrict__ col_arg) {
    const int b = blockIdx.y;
    const int l4 = threadIdx.x & 3;
    int p = blockIdx.x * 64 + (threadIdx.x >> 2);           // 64 points per block, rows first then columns
    const bool is_col = p >= N;
    if (is_col) p -= N;
    const bool live = p < (is_col ? M : N);
    int arg = 0x7fffffff;
    float v = 0.f;
    if (live) {
        if (!is_col) {
            const unsigned long long key = rowkey[(size_t)b * Npad + row_slot(p, R)];
            v = ordered_to_f32((uint32_t)(key >> 32));
            const int j0 = (int)(uint32_t)key * kColChunk;
            const float4 q = __ldg(&rowpk[(size_t)b * Npad + p]);
            const float4 *rec = colpk + (size_t)b * Mpad + j0;
            // lane l4 takes the pair records l4, l4+4, l4+8, l4+12 (columns 2 rec, 2 rec + 1): eight
            // independent 16-byte loads in flight per lane, highest column first so the lowest match wins
            float4 a[4], c[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) { a[t] = __ldg(&rec[2 * (l4 + 4 * t)]); c[t] = __ldg(&rec[2 * (l4 + 4 * t) + 1]); }
#pragma unroll
            for (int t = 3; t >= 0; --t) {
                const int j = j0 + 2 * (l4 + 4 * t);
                if (pair_dist_scalar<FORM>(q.x, q.y, q.z, q.w, a[t].y, a[t].w, c[t].y, c[t].w) == v) arg = j + 1;
                if (pair_dist_scalar<FORM>(q.x, q.y, q.z, q.w, a[t].x, a[t].z, c[t].x, c[t].z) == v) arg = j;
            }
        } else {
            const unsigned long long key = colkey[(size_t)b * Mpad + p];
            v = ordered_to_f32((uint32_t)(key >> 32));
            const uint32_t tag = (uint32_t)key;
            const int i0 = (int)(tag >> 5) * (32 * R) + (int)(tag & 31u) * R;
            const float *rec = reinterpret_cast<const float *>(colpk) + ((size_t)b * Mpad + (p & ~1)) * 4 + (p & 1);
            const float cx = rec[0], cy = rec[2], cz = rec[4], cn = rec[6];
            const float4 *rq = rowpk + (size_t)b * Npad + i0;
            // lane l4 takes rows l4, l4+4, l4+8, l4+12 of the winning lane's R rows (R = 2, 4, 8 or 16)
            float4 q[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) q[t] = (l4 + 4 * t < R) ? __ldg(&rq[l4 + 4 * t]) : make_float4(0.f, 0.f, 0.f, __int_as_float(0x7fc00000));
#pragma unroll
            for (int t = 3; t >= 0; --t)
                if (l4 + 4 * t < R && pair_dist_scalar<FORM>(q[t].x, q[t].y, q[t].z, q[t].w, cx, cy, cz, cn) == v) arg = i0 + l4 + 4 * t;
        }
    }
#pragma unroll
    for (int o = 2; o > 0; o >>= 1) arg = min(arg, __shfl_xor_sync(0xffffffffu, arg, o, 4));   // all lanes take part
    if (live && l4 == 0) {
        if (arg == 0x7fffffff) arg = 0;        // cannot happen: the tagged chunk holds the minimum
        const float val = apply_transform(transform, v);
        if (!is_col) { row_min[(size_t)b * N + p] = val; row_arg[(size_t)b * N + p] = arg; }
        else { col_min[(size_t)b * M + p] = val; col_arg[(size_t)b * M + p] = arg; }
    }
}

// Per-sample scaled sum / max / first argmax of the minima: grid (B, 2), one block per (sample,
// side), fixed summation order (per-thread strided partials, xor-shuffle tree, warps in order).
__global__ void __launch_bounds__(1024)
nn1_reduce_kernel(const float *__restrict__ row_min, const float *__restrict__ col_min, int N, int M, int B,
                  float row_scale, float col_scale, float *__restrict__ stats_f, int32_t *__restrict__ stats_i) {
    const int b = blockIdx.x, side = blockIdx.y;
    const int n = side ? M : N;
    const float *v = (side ? col_min : row_min) + (size_t)b * n;
    float s = 0.f, mx = -__int_as_float(0x7f800000);
    int am = 0x7fffffff;
    for (int i = threadIdx.x; i < n; i += 1024) {
        const float x = v[i];
        s += x;
        if (x > mx) { mx = x; am = i; }
    }
    reduce_smf(s, mx, am);
    __shared__ float ws[32], wmx[32];
    __shared__ int wam[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { ws[warp] = s; wmx[warp] = mx; wam[warp] = am; }
    __syncthreads();
    if (warp == 0) {
        s = ws[lane]; mx = wmx[lane]; am = wam[lane];
        reduce_smf(s, mx, am);
        if (lane == 0) {
            stats_f[(side * 2 + 0) * B + b] = s * (side ? col_scale : row_scale);
            stats_f[(side * 2 + 1) * B + b] = mx;
            stats_i[side * B + b] = am == 0x7fffffff ? 0 : am;
        }
    }
}

// -------------------------------------------------------------------------------- backward
struct BwdArgs {
    const float *rows; int64_t r_sb, r_sp, r_sc;
    const float *cols; int64_t c_sb, c_sp, c_sc;
    int B, N, M, swap_norms, transform;
    const int32_t *row_arg, *col_arg;
    const float *row_min, *col_min;
    const float *g_row, *g_col;
    const float *w_row_all, *w_row_max; const int32_t *row_argmax;
    const float *w_col_all, *w_col_max; const int32_t *col_argmax;
    int64_t ws0, ws1, ws2, ws3;     // element strides of the four w arrays (0 = broadcast scalar)
    float row_scale, col_scale;
    float *grad_rows; int64_t gr_sb, gr_sp, gr_sc;
    float *grad_cols; int64_t gc_sb, gc_sp, gc_sc;
};

__device__ __forceinline__ float3 ld3(const float *base, int64_t sc) {
    return make_float3(base[0], base[sc], base[2 * sc]);
}
// upstream gradient of minimum (b,p) on one side
__device__ __forceinline__ float upstream(const float *g, const float *w_all, int64_t s_all, const float *w_max,
                                          int64_t s_max, const int32_t *argmax, int b, int p, int n, float scale) {
    float r = 0.f;
    if (g) r += g[(size_t)b * n + p];
    if (w_all) r += w_all[b * s_all] * scale;
    if (w_max && argmax[b] == p) r += w_max[b * s_max];
    return r;
}
// d(value)/d(d2) factor: squared -> 2 * g (applied to (p - q)); sqrt -> g / v, 0 at v == 0
__device__ __forceinline__ float chain_factor(int transform, float g, const float *vals, size_t off) {
    if (transform == PCD_VALUE_SQRT_CLAMP) {
        const float v = vals[off];
        return v > 0.f ? g / v : 0.f;
    }
    return 2.f * g;
}

// MODE 0 (plain stores, writes every gradient element): the terms indexed by the thread's own
// point.  MODE 1 (atomics): the terms that land on the argmin partner.  MODE 2 = both in one
// launch with atomics only, for gradients the host has zeroed (dense outputs: one memset +
// one kernel instead of two dependent kernels).
template <int MODE>
__device__ __forceinline__ void emit3(float *gp, int64_t sc, float x, float y, float z) {
    if (MODE == 0) {
        gp[0] = x; gp[sc] = y; gp[2 * sc] = z;
    } else {
        atomicAdd(gp, x); atomicAdd(gp + sc, y); atomicAdd(gp + 2 * sc, z);
    }
}

template <int MODE>
__global__ void __launch_bounds__(256) nn1_bwd_kernel(BwdArgs a) {
    constexpr bool OWN = MODE != 1, SCAT = MODE != 0;
    const long long total = (long long)a.B * (a.N + a.M);
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(t / (a.N + a.M));
        const int p = (int)(t - (long long)b * (a.N + a.M));
        const float *rb = a.rows + b * a.r_sb, *cb = a.cols + b * a.c_sb;
        if (p < a.N) {
            const int i = p;
            const float g = upstream(a.g_row, a.w_row_all, a.ws0, a.w_row_max, a.ws1, a.row_argmax, b, i, a.N, a.row_scale);
            const int j = a.row_arg[(size_t)b * a.N + i];
            const float3 r = ld3(rb + i * a.r_sp, a.r_sc), c = ld3(cb + j * a.c_sp, a.c_sc);
            const float f = chain_factor(a.transform, g, a.row_min, (size_t)b * a.N + i);
            if (OWN && a.grad_rows) {
                float3 o;
                if (a.swap_norms) {
                    // entry (i,j): -2 g c_j ; column-direction entry (i*, j=i): +2 g' r_i
                    const float g2 = upstream(a.g_col, a.w_col_all, a.ws2, a.w_col_max, a.ws3, a.col_argmax, b, i, a.M, a.col_scale);
                    o = make_float3(-f * c.x + 2.f * g2 * r.x, -f * c.y + 2.f * g2 * r.y, -f * c.z + 2.f * g2 * r.z);
                } else {
                    o = make_float3(f * (r.x - c.x), f * (r.y - c.y), f * (r.z - c.z));
                }
                emit3<MODE>(a.grad_rows + b * a.gr_sb + i * a.gr_sp, a.gr_sc, o.x, o.y, o.z);
            }
            if (SCAT && g != 0.f) {
                if (a.grad_cols) {
                    float *gp = a.grad_cols + b * a.gc_sb + j * a.gc_sp;
                    if (a.swap_norms) emit3<1>(gp, a.gc_sc, -f * r.x, -f * r.y, -f * r.z);
                    else emit3<1>(gp, a.gc_sc, -f * (r.x - c.x), -f * (r.y - c.y), -f * (r.z - c.z));
                }
                if (a.swap_norms && a.grad_rows) {   // |rows_j|^2 term of entry (i,j)
                    const float3 rj = ld3(rb + j * a.r_sp, a.r_sc);
                    emit3<1>(a.grad_rows + b * a.gr_sb + j * a.gr_sp, a.gr_sc, f * rj.x, f * rj.y, f * rj.z);
                }
            }
        } else {
            const int j = p - a.N;
            const float g = upstream(a.g_col, a.w_col_all, a.ws2, a.w_col_max, a.ws3, a.col_argmax, b, j, a.M, a.col_scale);
            const int i = a.col_arg[(size_t)b * a.M + j];
            const float3 r = ld3(rb + i * a.r_sp, a.r_sc), c = ld3(cb + j * a.c_sp, a.c_sc);
            const float f = chain_factor(a.transform, g, a.col_min, (size_t)b * a.M + j);
            if (OWN && a.grad_cols) {
                float3 o;
                if (a.swap_norms) {
                    // entry (i*,j): -2 g' r_i* ; row-direction entry (i=j, j*): +2 g c_j
                    const float g1 = upstream(a.g_row, a.w_row_all, a.ws0, a.w_row_max, a.ws1, a.row_argmax, b, j, a.N, a.row_scale);
                    o = make_float3(-f * r.x + 2.f * g1 * c.x, -f * r.y + 2.f * g1 * c.y, -f * r.z + 2.f * g1 * c.z);
                } else {
                    o = make_float3(f * (c.x - r.x), f * (c.y - r.y), f * (c.z - r.z));
                }
                emit3<MODE>(a.grad_cols + b * a.gc_sb + j * a.gc_sp, a.gc_sc, o.x, o.y, o.z);
            }
            if (SCAT && g != 0.f) {
                if (a.grad_rows) {
                    float *gp = a.grad_rows + b * a.gr_sb + i * a.gr_sp;
                    if (a.swap_norms) emit3<1>(gp, a.gr_sc, -f * c.x, -f * c.y, -f * c.z);
                    else emit3<1>(gp, a.gr_sc, -f * (c.x - r.x), -f * (c.y - r.y), -f * (c.z - r.z));
                }
                if (a.swap_norms && a.grad_cols) {   // |cols_i*|^2 term of entry (i*,j)
                    const float3 ci = ld3(cb + i * a.c_sp, a.c_sc);
                    emit3<1>(a.grad_cols + b * a.gc_sb + i * a.gc_sp, a.gc_sc, f * ci.x, f * ci.y, f * ci.z);
                }
            }
        }
    }
}

// ----------------------------------------------------------------------------- host helpers
static int g_num_sms = 0;
static int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 0;
        if (cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) g_num_sms = 0;
    }
    return g_num_sms;
}

template <int FORM, int R>
static cudaError_t launch_sweep(const float4 *rowpk, const float4 *colpk, unsigned long long *rowkey,
                                unsigned long long *colkey, int B, int N, int M, int Npad, int Mpad, int mt,
                                int sms, cudaStream_t st) {
    const int QT = kSweepWarps * 32 * R;
    const int nqt = (N + QT - 1) / QT;                       // fully inert row tiles are skipped
    const int nq = (M + kQuad - 1) / kQuad;                  // ... and fully inert column quads
    const long long units = (long long)B * nqt * nq;
    if (units >= (1LL << 31)) return cudaErrorInvalidValue;
    static int occ = 0;                                      // per instantiation; queried once (the query costs microseconds)
    if (occ == 0) {
        int o = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, nn1_sweep_kernel<FORM, R>, kSweepThreads, 0);
        if (e != cudaSuccess) return e;
        occ = o < 1 ? 1 : o;
    }
    long long grid = (long long)sms * occ;
    if (grid > units) grid = units;
    nn1_sweep_kernel<FORM, R><<<(unsigned)grid, kSweepThreads, 0, st>>>(rowpk, colpk, rowkey, colkey, Npad, Mpad,
                                                                         mt / kQuad, nqt, nq, (int)units);
    return cudaGetLastError();
}

template <int FORM>
static cudaError_t launch_sweep_r(int R, const float4 *rowpk, const float4 *colpk, unsigned long long *rowkey,
                                  unsigned long long *colkey, int B, int N, int M, int Npad, int Mpad, int mt,
                                  int sms, cudaStream_t st) {
    switch (R) {
    case 16: return launch_sweep<FORM, 16>(rowpk, colpk, rowkey, colkey, B, N, M, Npad, Mpad, mt, sms, st);
    case 8: return launch_sweep<FORM, 8>(rowpk, colpk, rowkey, colkey, B, N, M, Npad, Mpad, mt, sms, st);
    case 4: return launch_sweep<FORM, 4>(rowpk, colpk, rowkey, colkey, B, N, M, Npad, Mpad, mt, sms, st);
    default: return launch_sweep<FORM, 2>(rowpk, colpk, rowkey, colkey, B, N, M, Npad, Mpad, mt, sms, st);
    }
}

// Tile-shape heuristic.  R rows per lane (register blocking: the per-step overhead -- operand
// LDS, CREDUX, ballots, stores -- is amortised over 2R pairs) against padding waste and the
// number of chunks each CTA of the persistent grid gets.  PCD_SWEEP_R / PCD_SWEEP_MT override.
static void choose_tiling(int B, int N, int M, int sms, int *R_out, int *mt_out) {
    static const int occ_of[5] = {6, 5, 4, 2, 0};       // CTAs/SM for R = 2, 4, 8, 16
    const long long nch = (M + kColChunk - 1) / kColChunk;
    int R = 16, oi = 3;
    while (R > 2) {
        const long long qtiles = (long long)B * ((N + 128 * R - 1) / (128 * R));
        const long long padded = qtiles * 128 * R;
        const bool waste = padded * 8 > (long long)B * N * 9;                 // > 12.5 % inert rows
        const bool starved = qtiles * nch < 6LL * sms * occ_of[oi];            // < 6 chunks per CTA
        if (!waste && !starved) break;
        R >>= 1; --oi;
    }
    int mt = kMaxColTile;
    if (const char *e = getenv("PCD_SWEEP_R")) { int v = atoi(e); if (v == 2 || v == 4 || v == 8 || v == 16) R = v; }
    if (const char *e = getenv("PCD_SWEEP_MT")) { int v = atoi(e); if (v >= 32 && v <= 256 && (v & (v - 1)) == 0) mt = v; }
    *R_out = R; *mt_out = mt;
}

}  // namespace pcd

// =================================================================================== C ABI
using namespace pcd;

extern "C" int pcd_version(void) { return PCD_VERSION; }
extern "C" const char *pcd_last_error(void) { return g_err; }

// optional profiling hook: events recorded around the sweep launch of the next forward calls
static thread_local cudaEvent_t g_sweep_ev0 = nullptr, g_sweep_ev1 = nullptr;
// process-wide, not thread-local: autograd runs backward functions on its own worker thread
static cudaEvent_t g_bwd_ev0 = nullptr, g_bwd_ev1 = nullptr;
extern "C" int pcd_nn1_set_backward_events(void *start_event, void *stop_event) {
    g_bwd_ev0 = (cudaEvent_t)start_event;
    g_bwd_ev1 = (cudaEvent_t)stop_event;
    return PCD_OK;
}
extern "C" int pcd_nn1_set_sweep_events(void *start_event, void *stop_event) {
    g_sweep_ev0 = (cudaEvent_t)start_event;
    g_sweep_ev1 = (cudaEvent_t)stop_event;
    return PCD_OK;
}

#ifdef PCD_SWEEP_TRACE
extern "C" int pcd_debug_set_sweep_trace(void *buf) {
    unsigned long long *p = (unsigned long long *)buf;
    PCD_CUDA_CHECK(cudaMemcpyToSymbol(g_sweep_trace, &p, sizeof(p)));
    return PCD_OK;
}
#endif

extern "C" size_t pcd_nn1_workspace_bytes(int B, int N, int M) {
    if (B <= 0 || N <= 0 || M <= 0) return 0;
    return nn1_layout(B, N, M).total;
}

extern "C" int pcd_nn1_forward(const float *rows, int64_t r_sb, int64_t r_sp, int64_t r_sc,
                               const float *cols, int64_t c_sb, int64_t c_sp, int64_t c_sc,
                               int B, int N, int M, int form, int norm_kind, int swap_norms, int transform,
                               float row_sum_scale, float col_sum_scale,
                               float *row_min, int32_t *row_arg, float *col_min, int32_t *col_arg,
                               float *stats_f, int32_t *stats_i,
                               void *workspace, size_t workspace_bytes, void *stream) {
    if (!rows || !cols || !row_min || !row_arg || !col_min || !col_arg || !stats_f || !stats_i || !workspace) {
        set_error("pcd_nn1_forward: NULL pointer argument");
        return PCD_ERR_ARG;
    }
    if (B <= 0 || N <= 0 || M <= 0 || form < 0 || form > 2 || norm_kind < 0 || norm_kind > 1 || transform < 0 ||
        transform > 1 || B > 65535) {
        set_error("pcd_nn1_forward: bad argument B=%d N=%d M=%d form=%d norm=%d transform=%d", B, N, M, form,
                  norm_kind, transform);
        return PCD_ERR_ARG;
    }
    if (swap_norms && N != M) {
        set_error("pcd_nn1_forward: swap_norms requires N == M (got %d, %d), as the reference's broadcast does", N, M);
        return PCD_ERR_ARG;
    }
    const Nn1Layout L = nn1_layout(B, N, M);
    if (workspace_bytes < L.total) {
        set_error("pcd_nn1_forward: workspace %zu < required %zu bytes", workspace_bytes, L.total);
        return PCD_ERR_WORKSPACE;
    }
    if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) {
        set_error("pcd_nn1_forward: workspace must be 256-byte aligned");
        return PCD_ERR_ARG;
    }
    const int sms = num_sms();
    if (sms <= 0) return cuda_fail(cudaGetLastError(), "no CUDA device");
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = (char *)workspace;
    float4 *rowpk = (float4 *)(ws + L.rowpk);
    float4 *rowpp = (float4 *)(ws + L.rowpp);
    float *colpk = (float *)(ws + L.colpk);
    unsigned long long *rowkey = (unsigned long long *)(ws + L.rowkey);
    unsigned long long *colkey = (unsigned long long *)(ws + L.colkey);

    int R, mt;
    choose_tiling(B, N, M, sms, &R, &mt);

    {
        const long long total = (long long)B * (L.Npad + L.Mpad);
        const int grid = (int)((total + 255) / 256 < (long long)sms * 8 ? (total + 255) / 256 : (long long)sms * 8);
        nn1_prep_kernel<<<grid, 256, 0, st>>>(rows, r_sb, r_sp, r_sc, cols, c_sb, c_sp, c_sc, B, N, M, L.Npad,
                                              L.Mpad, R, norm_kind, swap_norms, rowpk, rowpp, colpk, rowkey, colkey);
        PCD_CUDA_CHECK(cudaGetLastError());
    }
    {
        cudaError_t e;
        if (g_sweep_ev0) PCD_CUDA_CHECK(cudaEventRecord(g_sweep_ev0, st));
        if (form == PCD_FORM_ROW_COL)
            e = launch_sweep_r<PCD_FORM_ROW_COL>(R, rowpp, (const float4 *)colpk, rowkey, colkey, B, N, M, L.Npad, L.Mpad, mt, sms, st);
        else if (form == PCD_FORM_COL_ROW)
            e = launch_sweep_r<PCD_FORM_COL_ROW>(R, rowpp, (const float4 *)colpk, rowkey, colkey, B, N, M, L.Npad, L.Mpad, mt, sms, st);
        else
            e = launch_sweep_r<PCD_FORM_SUM_FIRST>(R, rowpp, (const float4 *)colpk, rowkey, colkey, B, N, M, L.Npad, L.Mpad, mt, sms, st);
        PCD_CUDA_CHECK(e);
        if (g_sweep_ev1) PCD_CUDA_CHECK(cudaEventRecord(g_sweep_ev1, st));
    }
    {
        const dim3 grid((N + M + 63) / 64 + 1, B);
        const float4 *colpk4 = (const float4 *)colpk;
#define PCD_LAUNCH_FIXUP(F)                                                                                    \
    nn1_fixup_kernel<F><<<grid, 256, 0, st>>>(rowpk, colpk4, rowkey, colkey, N, M, L.Npad, L.Mpad, R, transform, \
                                              row_min, row_arg, col_min, col_arg)
        if (form == PCD_FORM_ROW_COL) PCD_LAUNCH_FIXUP(PCD_FORM_ROW_COL);
        else if (form == PCD_FORM_COL_ROW) PCD_LAUNCH_FIXUP(PCD_FORM_COL_ROW);
        else PCD_LAUNCH_FIXUP(PCD_FORM_SUM_FIRST);
#undef PCD_LAUNCH_FIXUP
        PCD_CUDA_CHECK(cudaGetLastError());
        nn1_reduce_kernel<<<dim3(B, 2), 1024, 0, st>>>(row_min, col_min, N, M, B, row_sum_scale, col_sum_scale,
                                                       stats_f, stats_i);
        PCD_CUDA_CHECK(cudaGetLastError());
    }
    return PCD_OK;
}

extern "C" int pcd_nn1_backward(const float *rows, int64_t r_sb, int64_t r_sp, int64_t r_sc,
                                const float *cols, int64_t c_sb, int64_t c_sp, int64_t c_sc,
                                int B, int N, int M, int swap_norms, int transform,
                                const int32_t *row_arg, const int32_t *col_arg,
                                const float *row_min, const float *col_min,
                                const float *g_row, const float *g_col,
                                const float *w_row_all, const float *w_row_max, const int32_t *row_argmax,
                                const float *w_col_all, const float *w_col_max, const int32_t *col_argmax,
                                const int64_t *w_strides, float row_sum_scale, float col_sum_scale,
                                float *grad_rows, int64_t gr_sb, int64_t gr_sp, int64_t gr_sc,
                                float *grad_cols, int64_t gc_sb, int64_t gc_sp, int64_t gc_sc, void *stream) {
    if (!rows || !cols || !row_arg || !col_arg || B <= 0 || N <= 0 || M <= 0) {
        set_error("pcd_nn1_backward: bad argument");
        return PCD_ERR_ARG;
    }
    if ((w_row_max && !row_argmax) || (w_col_max && !col_argmax)) {
        set_error("pcd_nn1_backward: w_*_max given without *_argmax");
        return PCD_ERR_ARG;
    }
    if (transform == PCD_VALUE_SQRT_CLAMP && (!row_min || !col_min)) {
        set_error("pcd_nn1_backward: SQRT_CLAMP needs the stored minima");
        return PCD_ERR_ARG;
    }
    if (swap_norms && N != M) {
        set_error("pcd_nn1_backward: swap_norms requires N == M");
        return PCD_ERR_ARG;
    }
    if (!grad_rows && !grad_cols) return PCD_OK;
    const int sms = num_sms();
    if (sms <= 0) return cuda_fail(cudaGetLastError(), "no CUDA device");
    BwdArgs a{rows, r_sb, r_sp, r_sc, cols, c_sb, c_sp, c_sc, B, N, M, swap_norms, transform,
              row_arg, col_arg, row_min, col_min, g_row, g_col,
              w_row_all, w_row_max, row_argmax, w_col_all, w_col_max, col_argmax,
              w_strides ? w_strides[0] : 1, w_strides ? w_strides[1] : 1, w_strides ? w_strides[2] : 1,
              w_strides ? w_strides[3] : 1, row_sum_scale, col_sum_scale,
              grad_rows, gr_sb, gr_sp, gr_sc, grad_cols, gc_sb, gc_sp, gc_sc};
    const long long total = (long long)B * (N + M);
    const long long want = (total + 255) / 256;
    const int grid = (int)(want < (long long)sms * 16 ? want : (long long)sms * 16);
    cudaStream_t st = (cudaStream_t)stream;
    const bool dense_r = !grad_rows || (gr_sc == 1 && gr_sp == 3 && gr_sb == (int64_t)N * 3);
    const bool dense_c = !grad_cols || (gc_sc == 1 && gc_sp == 3 && gc_sb == (int64_t)M * 3);
    if (g_bwd_ev0) PCD_CUDA_CHECK(cudaEventRecord(g_bwd_ev0, st));
    if (dense_r && dense_c) {
        if (grad_rows) PCD_CUDA_CHECK(cudaMemsetAsync(grad_rows, 0, (size_t)B * N * 3 * sizeof(float), st));
        if (grad_cols) PCD_CUDA_CHECK(cudaMemsetAsync(grad_cols, 0, (size_t)B * M * 3 * sizeof(float), st));
        nn1_bwd_kernel<2><<<grid, 256, 0, st>>>(a);
        PCD_CUDA_CHECK(cudaGetLastError());
    } else {
        nn1_bwd_kernel<0><<<grid, 256, 0, st>>>(a);
        PCD_CUDA_CHECK(cudaGetLastError());
        nn1_bwd_kernel<1><<<grid, 256, 0, st>>>(a);
        PCD_CUDA_CHECK(cudaGetLastError());
    }
    if (g_bwd_ev1) PCD_CUDA_CHECK(cudaEventRecord(g_bwd_ev1, st));
    return PCD_OK;
}
