// Issue-slot microbenchmarks for the sweep's instruction mix (development tool, not product).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 fma2_s(float s, f32x2 b, f32x2 c) { f32x2 d; asm volatile("{\n.reg .b64 q;\nmov.b64 q, {%1, %1};\nfma.rn.f32x2 %0, q, %2, %3;\n}" : "=l"(d) : "f"(s), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float min3(float a, float b, float c) { float d; asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float fmaf_v(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float min2_v(float a, float b) { float d; asm volatile("min.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }

// MODE 0: 8 FFMA2 (packed, 3 packed operands)   1: 8 FFMA2 scalar-broadcast form
// MODE 2: 8 FFMA2 + 4 FMNMX3                    3: 8 FFMA2 + 8 FMNMX3
// MODE 4: 16 FFMA scalar                         5: 16 FFMA + 4 FMNMX3     6: 16 FFMA + 8 FMNMX3
// MODE 7: 8 FFMA2 + 8 FMNMX (2-input)            8: 16 FFMA + 8 FMNMX
template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int iters, float a, float b) {
    f32x2 acc[8]; float sc[16]; float m[8];
    for (int i = 0; i < 8; ++i) { acc[i] = pack2(threadIdx.x + i, i); m[i] = 1e30f + i; }
    for (int i = 0; i < 16; ++i) sc[i] = threadIdx.x + i;
    const f32x2 av = pack2(a, a * 0.5f), bv = pack2(b, b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep) {
            if (MODE <= 3 || MODE == 7) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    acc[i] = (MODE == 1) ? fma2_s(a, acc[i], bv) : fma2(acc[i], av, bv);
                    if (MODE == 2 && (i & 1)) { m[i] = min3(m[i], sc[i], sc[i + 8]); }
                    if (MODE == 3) { m[i] = min3(m[i], sc[i], sc[i + 8]); }
                    if (MODE == 7) { m[i] = min2_v(m[i], sc[i]); }
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    sc[i] = fmaf_v(sc[i], a, b);
                    if (MODE == 5 && (i & 3) == 3) m[i >> 2] = min3(m[i >> 2], sc[i], sc[i - 1]);
                    if (MODE == 6 && (i & 1)) m[i >> 1] = min3(m[i >> 1], sc[i], sc[i - 1]);
                    if (MODE == 8 && (i & 1)) m[i >> 1] = min2_v(m[i >> 1], sc[i]);
                }
            }
        }
    }
    float s = 0.f;
    for (int i = 0; i < 8; ++i) { float lo, hi; unpack2(acc[i], lo, hi); s += lo + hi + m[i]; }
    for (int i = 0; i < 16; ++i) s += sc[i];
    if (s == 123.456f) out[0] = s;
}
template <int MODE> void run(const char *name, int sms, float *d) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096, grid = sms * 8;
    float best = 1e9;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0); k<MODE><<<grid, 256>>>(d, iters, 1.0001f, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
    }
    const double fma = 64.0 * iters * 256.0 * grid;   // 64 FMAs per thread per iteration in every mode
    printf("%-34s %8.3f ms  %6.1f TFLOP/s (FMA only)  cycles/iter/SMSP @1.965GHz: %.1f\n", name, best, 2 * fma / best / 1e9,
           best * 1e-3 * 1.965e9 / iters / 16.0 /* warps per SMSP = 8 blk*8 warps/4 */);
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *d; cudaMalloc(&d, 256);
    run<0>("32 FFMA2 (packed operands)", sms, d);
    run<1>("32 FFMA2 (scalar-bcast operand)", sms, d);
    run<2>("32 FFMA2 + 16 FMNMX3", sms, d);
    run<3>("32 FFMA2 + 32 FMNMX3", sms, d);
    run<7>("32 FFMA2 + 32 FMNMX", sms, d);
    run<4>("64 FFMA", sms, d);
    run<5>("64 FFMA + 16 FMNMX3", sms, d);
    run<6>("64 FFMA + 32 FMNMX3", sms, d);
    run<8>("64 FFMA + 32 FMNMX", sms, d);
    return 0;
}
