// pcd_clip.cu -- the projection / clipping epilogues of the attack loops as ONE launch each, sm_100a.
//
// Every attack iteration ends with a clip of the perturbation that the reference writes as a chain of 8-25
// elementwise torch ops on [B,3,K] tensors:
//   attack/CW/CW_utils/clip_utils.py:5-29     ClipPointsL2          (one scale per sample)
//   attack/CW/CW_utils/clip_utils.py:32-56    ClipPointsLinf        (one scale per point)
//   attack/CW/CW_utils/clip_utils.py:59-136   ProjectInnerPoints (+ ClipPointsLinf = ProjectInnerClipLinf)
//   attack/GeoA3/GeoA3_attack.py:92-101       lp_clip
//   attack/GeoA3/GeoA3_attack.py:62-81        offset_proj           (after its knn_points, K = 1)
//   attack/GeoA3/GeoA3_attack.py:83-89        find_offset           (after its knn_points, K = 1)
// The clouds are a few hundred KB, so each torch op is pure launch latency (~2 us in a CUDA graph, ~8 us eager);
// here a thread owns a point and runs the chain in registers.  The fp32 operations are the reference's, in the
// reference's order, each rounded on its own (the file is compiled with --fmad=false):  x ** 2 is x * x,
// ** 0.5 is sqrt, sum(dim=1) is (x + y) + z, `budget / t` with a Python scalar on the left is
// reciprocal(t) * budget (torch.Tensor.__rtruediv__), tensor / tensor is an IEEE division.
// All tensors are channel-first contiguous [B,3,K] (as the attack loops hold them).
#include "pcd_common.cuh"

namespace pcd {

constexpr int kClipThreads = 256;

struct P3 { float x, y, z; };

__device__ __forceinline__ P3 load3(const float *p, size_t base, size_t K) { return {p[base], p[base + K], p[base + 2 * K]}; }
__device__ __forceinline__ void store3(float *p, size_t base, size_t K, P3 v) { p[base] = v.x; p[base + K] = v.y; p[base + 2 * K] = v.z; }
__device__ __forceinline__ P3 sub3(P3 a, P3 b) { return {__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)}; }
__device__ __forceinline__ float sumsq3(P3 a) {                 // torch.sum(a ** 2, dim=1)
    return __fadd_rn(__fadd_rn(__fmul_rn(a.x, a.x), __fmul_rn(a.y, a.y)), __fmul_rn(a.z, a.z));
}
__device__ __forceinline__ float dot3(P3 a, P3 b) {             // torch.sum(a * b, dim=1)
    return __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z));
}
__device__ __forceinline__ P3 cross3(P3 a, P3 b) {              // torch.cross(a, b, dim=1), products rounded before the subtraction
    return {__fsub_rn(__fmul_rn(a.y, b.z), __fmul_rn(a.z, b.y)), __fsub_rn(__fmul_rn(a.z, b.x), __fmul_rn(a.x, b.z)),
            __fsub_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x))};
}
// scale = clamp(budget / (norm + 1e-9), max=1): budget / tensor = reciprocal(tensor) * budget; NaN propagates as in torch.clamp
__device__ __forceinline__ float clip_scale(float norm, float budget) {
    const float s = __fmul_rn(__frcp_rn(__fadd_rn(norm, 1e-9f)), budget);
    return s > 1.f ? 1.f : s;
}
// pc = ori + (pc - ori) * min(budget / (|pc - ori| + 1e-9), 1)         clip_utils.py:50-56
__device__ __forceinline__ P3 clip_linf_point(P3 pc, P3 ori, float budget) {
    const P3 d = sub3(pc, ori);
    const float s = clip_scale(__fsqrt_rn(sumsq3(d)), budget);
    return {__fadd_rn(ori.x, __fmul_rn(d.x, s)), __fadd_rn(ori.y, __fmul_rn(d.y, s)), __fadd_rn(ori.z, __fmul_rn(d.z, s))};
}
// clip_utils.py:78-109: a point pushed inside the surface (offset . normal < 0) gets the offset diff * vref / (|vref| + 1e-9),
// vref = (normal x diff) x normal; 0 where diff is opposite to the normal (|normal x diff| < 1e-6)
__device__ __forceinline__ P3 project_inner_point(P3 pc, P3 ori, P3 n) {
    P3 d = sub3(pc, ori);
    const bool inner = dot3(d, n) < 0.f;
    const P3 vng = cross3(n, d);
    const float vng_norm = __fsqrt_rn(sumsq3(vng));
    const P3 vref = cross3(vng, n);
    const float t = __fadd_rn(__fsqrt_rn(sumsq3(vref)), 1e-9f);
    P3 proj = {__fdiv_rn(__fmul_rn(d.x, vref.x), t), __fdiv_rn(__fmul_rn(d.y, vref.y), t), __fdiv_rn(__fmul_rn(d.z, vref.z), t)};
    if (inner && vng_norm < 1e-6f) proj = {0.f, 0.f, 0.f};
    if (inner) d = proj;
    return {__fadd_rn(ori.x, d.x), __fadd_rn(ori.y, d.y), __fadd_rn(ori.z, d.z)};
}

template <int MODE>
__global__ void __launch_bounds__(kClipThreads)
clip_points_kernel(float *__restrict__ pc, const float *__restrict__ ori, const float *__restrict__ normal, int B, int K, float budget) {
    const long long total = (long long)B * K;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(t / K);
        const size_t base = (size_t)b * 3 * K + (size_t)(t - (long long)b * K);
        P3 p = load3(pc, base, K);
        const P3 o = load3(ori, base, K);
        if (MODE == PCD_CLIP_PROJECT_LINF) p = project_inner_point(p, o, load3(normal, base, K));
        store3(pc, base, K, clip_linf_point(p, o, budget));
    }
}

// ClipPointsL2 (clip_utils.py:22-29): ONE scale per sample from the norm of the whole perturbation.  One CTA per
// sample, two passes over its 3K floats (the second one hits L1/L2); the 3K-term sum runs in a fixed tree order
// (torch's reduction order is its own: values agree to rounding, ~1e-7, not bit for bit).
__global__ void __launch_bounds__(1024) clip_l2_kernel(float *__restrict__ pc, const float *__restrict__ ori, int K, float budget) {
    __shared__ float sh[32];
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *p = pc + (size_t)b * 3 * K;
    const float *o = ori + (size_t)b * 3 * K;
    float s = 0.f;
    for (int i = threadIdx.x; i < 3 * K; i += 1024) {
        const float d = __fsub_rn(p[i], o[i]);
        s = __fadd_rn(s, __fmul_rn(d, d));
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) sh[warp] = s;
    __syncthreads();
    float tot = sh[lane];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, off);
    const float scale = clip_scale(__fsqrt_rn(tot), budget);
    for (int i = threadIdx.x; i < 3 * K; i += 1024) p[i] = __fadd_rn(o[i], __fmul_rn(__fsub_rn(p[i], o[i]), scale));
}

// GeoA3_attack.py:92-101
__global__ void __launch_bounds__(kClipThreads)
lp_clip_kernel(const float *__restrict__ offset, int B, int K, float cc, float *__restrict__ out) {
    const long long total = (long long)B * K;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(t / K);
        const size_t base = (size_t)b * 3 * K + (size_t)(t - (long long)b * K);
        const P3 o = load3(offset, base, K);
        const float len = __fsqrt_rn(sumsq3(o));
        P3 r = {0.f, 0.f, 0.f};
        if (len > 1e-6f) r = {__fmul_rn(__fdiv_rn(o.x, len), cc), __fmul_rn(__fdiv_rn(o.y, len), cc), __fmul_rn(__fdiv_rn(o.z, len), cc)};
        if (len < cc) r = o;
        store3(out, base, K, r);
    }
}

// GeoA3_attack.py:62-81 (PROJ) and :83-89 (FIND) after their knn_points(K=1): idx[b,k] = the original point nearest to
// point k.  PROJ: the offset's component along the (normalised) normal of that point; FIND: adv - ori[idx].
template <bool PROJ>
__global__ void __launch_bounds__(kClipThreads)
offset_gather_kernel(const float *__restrict__ a, const float *__restrict__ table, const int32_t *__restrict__ idx, int B, int K, int M,
                     float *__restrict__ out) {
    const long long total = (long long)B * K;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(t / K);
        const size_t base = (size_t)b * 3 * K + (size_t)(t - (long long)b * K);
        int j = idx[t];
        j = j < 0 ? 0 : (j >= M ? M - 1 : j);
        const P3 v = load3(a, base, K);
        const P3 g = load3(table, (size_t)b * 3 * M + j, M);
        P3 r;
        if (PROJ) {
            const float tl = __fadd_rn(__fsqrt_rn(sumsq3(g)), 1e-6f);
            const float dot = __fadd_rn(__fadd_rn(__fdiv_rn(__fmul_rn(v.x, g.x), tl), __fdiv_rn(__fmul_rn(v.y, g.y), tl)),
                                        __fdiv_rn(__fmul_rn(v.z, g.z), tl));
            r = {__fdiv_rn(__fmul_rn(dot, g.x), tl), __fdiv_rn(__fmul_rn(dot, g.y), tl), __fdiv_rn(__fmul_rn(dot, g.z), tl)};
        } else {
            r = sub3(v, g);
        }
        store3(out, base, K, r);
    }
}

static inline int clip_grid(long long total) {
    long long g = (total + kClipThreads - 1) / kClipThreads;
    const long long cap = (long long)num_sms() * 8;
    if (g > cap) g = cap;
    return g < 1 ? 1 : (int)g;
}

}  // namespace pcd

using namespace pcd;

extern "C" int pcd_clip_points(float *pc, const float *ori, const float *normal, int B, int K, int mode, float budget, void *stream) {
    if (!pc || !ori || B <= 0 || K <= 0 || mode < PCD_CLIP_LINF || mode > PCD_CLIP_L2 || (mode == PCD_CLIP_PROJECT_LINF && !normal)) {
        set_error("pcd_clip_points: bad argument B=%d K=%d mode=%d", B, K, mode);
        return PCD_ERR_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = clip_grid((long long)B * K);
    if (mode == PCD_CLIP_LINF) clip_points_kernel<PCD_CLIP_LINF><<<grid, kClipThreads, 0, st>>>(pc, ori, nullptr, B, K, budget);
    else if (mode == PCD_CLIP_PROJECT_LINF) clip_points_kernel<PCD_CLIP_PROJECT_LINF><<<grid, kClipThreads, 0, st>>>(pc, ori, normal, B, K, budget);
    else clip_l2_kernel<<<B, 1024, 0, st>>>(pc, ori, K, budget);
    PCD_CUDA_CHECK(cudaGetLastError());
    return PCD_OK;
}

extern "C" int pcd_lp_clip(const float *offset, int B, int K, float cc_linf, float *out, void *stream) {
    if (!offset || !out || B <= 0 || K <= 0) {
        set_error("pcd_lp_clip: bad argument B=%d K=%d", B, K);
        return PCD_ERR_ARG;
    }
    lp_clip_kernel<<<clip_grid((long long)B * K), kClipThreads, 0, (cudaStream_t)stream>>>(offset, B, K, cc_linf, out);
    PCD_CUDA_CHECK(cudaGetLastError());
    return PCD_OK;
}

extern "C" int pcd_offset_proj(const float *offset, const float *ori_normal, const int32_t *idx, int B, int K, int M, float *out,
                               void *stream) {
    if (!offset || !ori_normal || !idx || !out || B <= 0 || K <= 0 || M <= 0) {
        set_error("pcd_offset_proj: bad argument B=%d K=%d M=%d", B, K, M);
        return PCD_ERR_ARG;
    }
    offset_gather_kernel<true><<<clip_grid((long long)B * K), kClipThreads, 0, (cudaStream_t)stream>>>(offset, ori_normal, idx, B, K, M, out);
    PCD_CUDA_CHECK(cudaGetLastError());
    return PCD_OK;
}

extern "C" int pcd_find_offset(const float *adv, const float *ori, const int32_t *idx, int B, int K, int M, float *out, void *stream) {
    if (!adv || !ori || !idx || !out || B <= 0 || K <= 0 || M <= 0) {
        set_error("pcd_find_offset: bad argument B=%d K=%d M=%d", B, K, M);
        return PCD_ERR_ARG;
    }
    offset_gather_kernel<false><<<clip_grid((long long)B * K), kClipThreads, 0, (cudaStream_t)stream>>>(adv, ori, idx, B, K, M, out);
    PCD_CUDA_CHECK(cudaGetLastError());
    return PCD_OK;
}
