// Inner-loop laboratory for the NN-1 sweep (development tool, not product).
// Replays the sweep's step (2 columns x R rows per lane, packed math, row + column minima)
// on a shared-memory resident tile with different min strategies and reports cycles per pair
// per SM sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o tools/sweep_lab tools/sweep_lab.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../3dpointcloudattack_b200/csrc/pcd_common.cuh"
using namespace pcd;
namespace pcd { void set_error(const char *, ...) {} int cuda_fail(cudaError_t, const char *) { return 1; } }

constexpr int kCols = 256;

__device__ __forceinline__ float min2(float a, float b) {
    float d;
    asm("min.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
    return d;
}

// NaN-propagating twin: alternating it with min2 keeps ptxas from fusing two 2-input mins into FMNMX3
__device__ __forceinline__ float min2n(float a, float b) {
    float d;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
    return d;
}

// ROWM: 0 none, 2 two-input, 3 three-input.   COLM: 0 none, 2 two-input chain, 3 three-input tree,
// 4 = two-input tree.  COLRED: 0 no warp reduction, 1 CREDUX + ballot + STS.
// MATH: 0 = the reference's five-instruction form; 1 = four instructions: v = fma(qz,cz, fma(qy,cy, fma(qx,cx, ncol))) feeds the
// row minima, d = v + nrow the column minima (approximate sweep + exact fix-up); the column ballot then takes every lane within a window
template <int R, int ROWM, int COLM, int COLRED, int OCC, int MATH = 0>
__global__ void __launch_bounds__(128, OCC) lab(float *out, int iters) {
    __shared__ __align__(128) float4 tile[kCols + 2];
    __shared__ uint2 colpart[4][kCols];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < kCols + 2; i += 128) tile[i] = make_float4(0.001f * i, 0.002f * i + 1, 0.5f - 0.001f * i, 0.25f + i);
    __syncthreads();
    float qx[R], qy[R], qz[R], qn[R], best[R];
    uint32_t btag[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        qx[r] = -2.f * (0.01f * (tid * R + r)); qy[r] = 0.3f + r; qz[r] = 0.1f * lane + r; qn[r] = 1.f + r;
        best[r] = __int_as_float(0x7f800000); btag[r] = 0;
    }
    uint2 *cp = colpart[warp];
    for (int it = 0; it < iters; ++it) {
        float4 A = tile[0], Bv = tile[1];
        uint2 pend_lo = make_uint2(0x7f800000u, 0u), pend_hi = pend_lo;
        int pend_at = -1;
        for (int chunk = 0; chunk < kCols / 32; ++chunk) {
            float m[R];
#pragma unroll
            for (int r = 0; r < R; ++r) m[r] = __int_as_float(0x7f800000);
#pragma unroll 2
            for (int s = 0; s < 16; ++s) {
                const int step = chunk * 16 + s;
                const float4 An = tile[step * 2 + 2], Bn = tile[step * 2 + 3];
                const f32x2 X = pack2(A.x, A.y), Y = pack2(A.z, A.w);
                const f32x2 Z = pack2(Bv.x, Bv.y), Nn = pack2(Bv.z, Bv.w);
                float lo[R], hi[R];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (MATH == 1) {
                        const f32x2 v = fma2_s(qz[r], Z, fma2_s(qy[r], Y, fma2_s(qx[r], X, Nn)));
                        float vl, vh;
                        unpack2(v, vl, vh);
                        m[r] = min3(m[r], vl, vh);
                        unpack2(add2_s(qn[r], v), lo[r], hi[r]);
                        continue;
                    }
                    const f32x2 d = pair_dist_x2<PCD_FORM_SUM_FIRST>(qx[r], qy[r], qz[r], qn[r], X, Y, Z, Nn);
                    unpack2(d, lo[r], hi[r]);
                    if (ROWM == 3) m[r] = min3(m[r], lo[r], hi[r]);
                    if (ROWM == 2) { m[r] = min2(m[r], lo[r]); m[r] = min2n(m[r], hi[r]); }
                    if (ROWM == 0) m[r] += lo[r] + hi[r];
                }
                float clo = lo[0], chi = hi[0];
                if (COLM == 3) {
#pragma unroll
                    for (int r = 1; r + 1 < R; r += 2) { clo = min3(clo, lo[r], lo[r + 1]); chi = min3(chi, hi[r], hi[r + 1]); }
                    clo = min2(clo, lo[R - 1]); chi = min2(chi, hi[R - 1]);
                } else if (COLM == 5) {        // balanced three-input tree (depth 3 for R = 16 instead of a chain of 8)
                    static_assert(COLM != 5 || R == 16, "hand-unrolled for R = 16");
                    const float a0 = min3(lo[0], lo[1], lo[2]), a1 = min3(lo[3], lo[4], lo[5]), a2 = min3(lo[6], lo[7], lo[8]);
                    const float a3 = min3(lo[9], lo[10], lo[11]), a4 = min3(lo[12], lo[13], lo[14]);
                    const float b0 = min3(a0, a1, a2), b1 = min3(a3, a4, lo[15]);
                    clo = min2(b0, b1);
                    const float c0 = min3(hi[0], hi[1], hi[2]), c1 = min3(hi[3], hi[4], hi[5]), c2 = min3(hi[6], hi[7], hi[8]);
                    const float c3 = min3(hi[9], hi[10], hi[11]), c4 = min3(hi[12], hi[13], hi[14]);
                    const float d0 = min3(c0, c1, c2), d1 = min3(c3, c4, hi[15]);
                    chi = min2(d0, d1);
                } else if (COLM == 2) {
#pragma unroll
                    for (int r = 1; r < R; ++r) { clo = (r & 1) ? min2n(clo, lo[r]) : min2(clo, lo[r]); chi = (r & 1) ? min2n(chi, hi[r]) : min2(chi, hi[r]); }
                } else if (COLM == 4) {
                    float tl[R], th[R];
#pragma unroll
                    for (int r = 0; r < R; ++r) { tl[r] = lo[r]; th[r] = hi[r]; }
#pragma unroll
                    for (int w = R / 2; w >= 1; w >>= 1)
#pragma unroll
                        for (int r = 0; r < w; ++r) { tl[r] = (w & 5) ? min2n(tl[r], tl[r + w]) : min2(tl[r], tl[r + w]); th[r] = (w & 5) ? min2n(th[r], th[r + w]) : min2(th[r], th[r + w]); }
                    clo = tl[0]; chi = th[0];
                }
                if (COLRED) {
                    if (pend_at >= 0) *reinterpret_cast<uint4 *>(&cp[pend_at]) = make_uint4(pend_lo.x, pend_lo.y, pend_hi.x, pend_hi.y);
                    const float vlo = warp_min_f32(clo), vhi = warp_min_f32(chi);
                    if (MATH == 1) {
                        pend_lo = make_uint2(__float_as_uint(vlo), __ballot_sync(0xffffffffu, clo <= vlo + qn[0]));
                        pend_hi = make_uint2(__float_as_uint(vhi), __ballot_sync(0xffffffffu, chi <= vhi + qn[0]));
                    } else {
                    pend_lo = make_uint2(__float_as_uint(vlo), __ballot_sync(0xffffffffu, clo == vlo));
                    pend_hi = make_uint2(__float_as_uint(vhi), __ballot_sync(0xffffffffu, chi == vhi));
                    }
                    pend_at = 2 * step;
                } else if (COLM) {
                    m[0] = min2(m[0], clo); m[1] = min2(m[1], chi);
                }
                A = An; Bv = Bn;
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (m[r] < best[r]) { best[r] = m[r]; btag[r] = (uint32_t)(chunk + it); }
            }
        }
        if (COLRED) *reinterpret_cast<uint4 *>(&cp[pend_at]) = make_uint4(pend_lo.x, pend_lo.y, pend_hi.x, pend_hi.y);
    }
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) s += best[r] + btag[r];
    __syncthreads();
    if (COLRED) s += __uint_as_float(colpart[warp][lane].x);
    if (iters == -12345) out[tid] = s;
}

template <int R, int ROWM, int COLM, int COLRED, int OCC, int MATH = 0>
void run(const char *name, int sms, float *d) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 64, grid = sms * OCC;
    float best = 1e9;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); lab<R, ROWM, COLM, COLRED, OCC, MATH><<<grid, 128>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
    }
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, lab<R, ROWM, COLM, COLRED, OCC, MATH>);
    // pairs per SMSP = OCC CTAs * 1 warp each (4 warps per CTA over 4 SMSPs) * 32 lanes * R rows * kCols * iters
    const double pairs_per_smsp_lane = (double)OCC * R * kCols * iters;
    const double cyc = best * 1e-3 * 1.965e9;
    printf("%-44s R=%2d occ=%d regs=%3d  %8.1f us  %6.3f cycles/pair  -> %5.1f %% of FFMA peak\n", name, R, OCC, fa.numRegs,
           best * 1e3, cyc / pairs_per_smsp_lane, 100.0 * 4.0 / (cyc / pairs_per_smsp_lane));
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *d; cudaMalloc(&d, 256);
    run<16, 0, 0, 0, 2>("math only (sum sink)", sms, d);
    run<16, 3, 0, 0, 2>("row min3", sms, d);
    run<16, 2, 0, 0, 2>("row min2", sms, d);
    run<16, 3, 3, 0, 2>("row min3 + col min3 tree (no warp red)", sms, d);
    run<16, 3, 3, 1, 2>("row min3 + col min3 tree + CREDUX (current)", sms, d);
    run<16, 3, 5, 1, 2>("row min3 + col BALANCED min3 tree + CREDUX", sms, d);
    run<16, 3, 5, 1, 2, 1>("four-instruction math, balanced tree + window ballot", sms, d);
    run<16, 3, 3, 1, 2, 1>("FOUR-instruction math + window ballot", sms, d);
    run<16, 3, 3, 0, 2, 1>("FOUR-instruction math, no warp red", sms, d);
    run<16, 3, 3, 1, 1, 1>("FOUR-instruction math, one CTA per SM", sms, d);
    run<16, 3, 3, 1, 1>("current, ONE CTA per SM (1 warp per scheduler)", sms, d);
    run<16, 2, 2, 1, 2>("row min2 + col min2 chain + CREDUX", sms, d);
    run<16, 2, 4, 1, 2>("row min2 + col min2 tree + CREDUX", sms, d);
    run<16, 3, 2, 1, 2>("row min3 + col min2 chain + CREDUX", sms, d);
    run<16, 3, 4, 1, 2>("row min3 + col min2 tree + CREDUX", sms, d);
    run<16, 2, 3, 1, 2>("row min2 + col min3 tree + CREDUX", sms, d);
    run<8, 3, 3, 1, 4>("R=8 current", sms, d);
    run<8, 2, 4, 1, 4>("R=8 row min2 + col min2 tree + CREDUX", sms, d);
    run<8, 2, 2, 1, 4>("R=8 row min2 + col min2 chain + CREDUX", sms, d);
    run<8, 3, 3, 1, 3>("R=8 current occ3", sms, d);
    run<12, 3, 3, 1, 2>("R=12 current", sms, d);
    run<12, 2, 2, 1, 3>("R=12 min2 chain occ3", sms, d);
    return 0;
}
