// Issue-slot microbenchmarks, round 2 (development tool): the sweep's real per-step mix -- 16 rows x
// (FMUL2, FFMA2, FFMA2, FADD2, FADD2) on fresh operands -- with different ways of folding the 32 results
// into row minima and column minima.  Reports cycles per pair per SMSP lane (ideal math only = 5.0).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o tools/ubench2 tools/ubench2.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 mul2_s(float s, f32x2 b) { f32x2 d; asm volatile("{\n.reg .b64 q;\nmov.b64 q, {%1, %1};\nmul.rn.f32x2 %0, q, %2;\n}" : "=l"(d) : "f"(s), "l"(b)); return d; }
__device__ __forceinline__ f32x2 fma2_s(float s, f32x2 b, f32x2 c) { f32x2 d; asm volatile("{\n.reg .b64 q;\nmov.b64 q, {%1, %1};\nfma.rn.f32x2 %0, q, %2, %3;\n}" : "=l"(d) : "f"(s), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 add2_s(float s, f32x2 b) { f32x2 d; asm volatile("{\n.reg .b64 q;\nmov.b64 q, {%1, %1};\nadd.rn.f32x2 %0, q, %2;\n}" : "=l"(d) : "f"(s), "l"(b)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float min3(float a, float b, float c) { float d; asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float min2(float a, float b) { float d; asm volatile("min.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float min2n(float a, float b) { float d; asm volatile("min.NaN.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ int imin2(int a, int b) { int d; asm volatile("min.s32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ unsigned umin2(unsigned a, unsigned b) { unsigned d; asm volatile("min.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }

constexpr int R = 16;
// ROW: 0 none, 1 min3(m,lo,hi), 2 two accumulators min2, 3 single accumulator min2/min2n alternating, 4 integer min.s32 x2 (two acc),
//      5 min2(lo,hi) then min2(m, .) (tree form, single acc)
// COL: 0 none, 1 min3 tree (current), 2 min2 with 4 accumulators per half, 3 min.s32 with 4 accumulators, 4 min2 chain alternating NaN
template <int ROW, int COL>
__global__ void __launch_bounds__(128, 2) k(float *out, int iters, float a, float b) {
    float qx[R], qy[R], qz[R], qn[R], m[R], m2[R];
    for (int r = 0; r < R; ++r) { qx[r] = a * (threadIdx.x + r); qy[r] = b + r; qz[r] = a - r; qn[r] = 1.f + r; m[r] = 1e30f; m2[r] = 1e30f; }
    f32x2 X = pack2(a, b), Y = pack2(b, a), Z = pack2(a + 1, b + 1), Nn = pack2(0.5f, 0.25f);
    float csink = 1e30f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 2
        for (int s = 0; s < 16; ++s) {
            float lo[R], hi[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                f32x2 t = mul2_s(qx[r], X);
                t = fma2_s(qy[r], Y, t);
                t = fma2_s(qz[r], Z, t);
                const f32x2 d = add2(add2_s(qn[r], Nn), t);
                unpack2(d, lo[r], hi[r]);
                if (ROW == 1) m[r] = min3(m[r], lo[r], hi[r]);
                if (ROW == 2) { m[r] = min2(m[r], lo[r]); m2[r] = min2(m2[r], hi[r]); }
                if (ROW == 3) { m[r] = min2(m[r], lo[r]); m[r] = min2n(m[r], hi[r]); }
                if (ROW == 4) { m[r] = __int_as_float(imin2(__float_as_int(m[r]), __float_as_int(lo[r]))); m2[r] = __int_as_float(imin2(__float_as_int(m2[r]), __float_as_int(hi[r]))); }
                if (ROW == 5) { m[r] = min2(m[r], min2(lo[r], hi[r])); }
            }
            if (COL == 1) {
                float clo = lo[0], chi = hi[0];
#pragma unroll
                for (int r = 1; r + 1 < R; r += 2) { clo = min3(clo, lo[r], lo[r + 1]); chi = min3(chi, hi[r], hi[r + 1]); }
                clo = min2(clo, lo[R - 1]); chi = min2(chi, hi[R - 1]);
                csink = min3(csink, clo, chi);
            } else if (COL == 2 || COL == 3) {
                float cl[4], ch[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) { cl[q] = lo[q]; ch[q] = hi[q]; }
#pragma unroll
                for (int r = 4; r < R; ++r) {
                    if (COL == 2) { cl[r & 3] = min2(cl[r & 3], lo[r]); ch[r & 3] = min2(ch[r & 3], hi[r]); }
                    else { cl[r & 3] = __int_as_float(imin2(__float_as_int(cl[r & 3]), __float_as_int(lo[r]))); ch[r & 3] = __int_as_float(imin2(__float_as_int(ch[r & 3]), __float_as_int(hi[r]))); }
                }
                const float clo = min2(min2(cl[0], cl[1]), min2(cl[2], cl[3])), chi = min2(min2(ch[0], ch[1]), min2(ch[2], ch[3]));
                csink = min2(min2(csink, clo), chi);
            } else if (COL == 4) {
                float clo = lo[0], chi = hi[0];
#pragma unroll
                for (int r = 1; r < R; ++r) { clo = (r & 1) ? min2n(clo, lo[r]) : min2(clo, lo[r]); chi = (r & 1) ? min2n(chi, hi[r]) : min2(chi, hi[r]); }
                csink = min2(min2(csink, clo), chi);
            }
            // rotate the broadcast operands a little so nothing is loop invariant
            X = add2(X, Nn); 
        }
    }
    float sres = csink;
    for (int r = 0; r < R; ++r) sres += m[r] + m2[r];
    if (sres == 123.456f) out[0] = sres;
}
template <int ROW, int COL> void run(const char *name, int sms, float *d) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 512, grid = sms * 2;
    float best = 1e9;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0); k<ROW, COL><<<grid, 128>>>(d, iters, 1.0001f, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
    }
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k<ROW, COL>);
    // per SMSP: 2 warps (2 CTAs x 4 warps / 4 SMSPs); pairs per lane = iters * 16 steps * 2 cols * R rows
    const double pairs = 2.0 * iters * 16 * 2 * R;
    printf("%-58s regs=%3d %8.3f ms  %6.3f cycles/pair\n", name, fa.numRegs, best, best * 1e-3 * 1.965e9 / pairs);
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *d; cudaMalloc(&d, 256);
    run<0, 0>("math only", sms, d);
    run<1, 0>("row min3", sms, d);
    run<2, 0>("row min2 x2 accumulators", sms, d);
    run<3, 0>("row min2/min2.NaN single accumulator", sms, d);
    run<4, 0>("row min.s32 x2 accumulators", sms, d);
    run<5, 0>("row min2(m, min2(lo,hi))", sms, d);
    run<0, 1>("col min3 tree", sms, d);
    run<0, 2>("col min2 x4 accumulators", sms, d);
    run<0, 3>("col min.s32 x4 accumulators", sms, d);
    run<0, 4>("col min2/min2.NaN chain", sms, d);
    run<1, 1>("row min3 + col min3 tree (current)", sms, d);
    run<2, 2>("row min2 x2 + col min2 x4", sms, d);
    run<4, 3>("row min.s32 x2 + col min.s32 x4", sms, d);
    run<1, 2>("row min3 + col min2 x4", sms, d);
    run<2, 1>("row min2 x2 + col min3 tree", sms, d);
    return 0;
}
