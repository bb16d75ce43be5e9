// pcd_peak.cu -- FP32 FMA roofline probe (measurement helper for bench.py).
//
// MEASURED_PEAKS.json carries HBM and bf16 tensor peaks only; the NN-1 / k-NN sweeps are bound
// by the CUDA-core fp32 pipe, so bench.py measures that denominator in the same run: an
// FFMA-only kernel (16 independent accumulator chains per thread, scalar FFMA and packed
// FFMA2 variants; the larger of the two is reported).
#include "pcd_common.cuh"

namespace pcd {

template <bool PACKED>
__global__ void __launch_bounds__(256) fma_peak_kernel(float *out, int iters, float a, float b) {
    if (PACKED) {
        f32x2 acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = pack2((float)threadIdx.x + i, (float)i);
        const f32x2 av = pack2(a, a * 0.5f), bv = pack2(b, b);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int rep = 0; rep < 8; ++rep) {
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = fma2(acc[i], av, bv);
            }
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { float lo, hi; unpack2(acc[i], lo, hi); s += lo + hi; }
        if (s == 123.456f) out[0] = s;
    } else {
        float acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = (float)threadIdx.x + i;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = __fmaf_rn(acc[i], a, b);
            }
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) s += acc[i];
        if (s == 123.456f) out[0] = s;
    }
}

}  // namespace pcd

using namespace pcd;

extern "C" int pcd_measure_fp32_peak(int iters, double *flops_per_s, void *stream) {
    if (!flops_per_s || iters <= 0) {
        set_error("pcd_measure_fp32_peak: bad argument");
        return PCD_ERR_ARG;
    }
    int dev = 0, sms = 0;
    PCD_CUDA_CHECK(cudaGetDevice(&dev));
    PCD_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cudaStream_t st = (cudaStream_t)stream;
    float *scratch = nullptr;
    PCD_CUDA_CHECK(cudaMalloc(&scratch, 256));
    cudaEvent_t e0, e1;
    PCD_CUDA_CHECK(cudaEventCreate(&e0));
    PCD_CUDA_CHECK(cudaEventCreate(&e1));
    const int grid = sms * 8;
    double best = 0.0;
    for (int variant = 0; variant < 2; ++variant) {
        for (int rep = 0; rep < 4; ++rep) {   // first rep = warm-up
            PCD_CUDA_CHECK(cudaEventRecord(e0, st));
            if (variant == 0) fma_peak_kernel<false><<<grid, 256, 0, st>>>(scratch, iters, 1.0001f, 0.5f);
            else fma_peak_kernel<true><<<grid, 256, 0, st>>>(scratch, iters, 1.0001f, 0.5f);
            PCD_CUDA_CHECK(cudaGetLastError());
            PCD_CUDA_CHECK(cudaEventRecord(e1, st));
            PCD_CUDA_CHECK(cudaEventSynchronize(e1));
            float ms = 0.f;
            PCD_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
            // 64 FMA per thread per iteration in both variants (16*4 scalar, 8*8 packed x 2 / 2 .. see below)
            const double fma_per_thread = (variant == 0) ? 64.0 * iters : 128.0 * iters;
            const double flops = 2.0 * fma_per_thread * 256.0 * grid / (ms * 1e-3);
            if (rep > 0 && flops > best) best = flops;
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(scratch);
    *flops_per_s = best;
    return PCD_OK;
}
