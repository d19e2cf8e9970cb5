// R1-R4: pixel grid -> camera directions -> world rays -> unit-sphere far bound -> packed [o,d,near,far].
// Reference: util/ray.py:8-12,25-31,46-54,81-99 and dataset/base.py:211-219.
// Rounding follows the reference's CPU evaluation so o and d are bit-identical: integer pixel centres,
// (i-cx)/fx as sub+div, the 3x3 rotation as an FMA chain in k order, norm as FMA chain + sqrt + div,
// the three dot products of rays_intersect_sphere as ((p0+p1)+p2).  `far` can differ by 1 ulp on ~0.4 %
// of rays because the reference's vectorised CPU sqrt is not correctly rounded while sqrt.rn is.
#include "launchers.h"

namespace clift {
namespace {

struct RayGenParams {
    float fx, fy, cx, cy;
    float R[9];
    float o[3];
    float near, r2;
    int H, W;
    float* rays;
    int32_t* bad;
};

__device__ __forceinline__ float dot3_seq(const float* a, const float* b) {
    return __fadd_rn(__fadd_rn(__fmul_rn(a[0], b[0]), __fmul_rn(a[1], b[1])), __fmul_rn(a[2], b[2]));
}

__global__ void __launch_bounds__(256) gen_rays_kernel(const __grid_constant__ RayGenParams P) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)P.H * P.W) return;
    const int row = (int)(idx / P.W), col = (int)(idx - (int64_t)row * P.W);
    const float dir[3] = {__fdiv_rn(__fsub_rn((float)col, P.cx), P.fx), __fdiv_rn(__fsub_rn((float)row, P.cy), P.fy), 1.0f};
    float d[3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
        d[k] = __fmaf_rn(dir[2], P.R[k * 3 + 2], __fmaf_rn(dir[1], P.R[k * 3 + 1], __fmul_rn(dir[0], P.R[k * 3 + 0])));
    const float nrm = __fsqrt_rn(__fmaf_rn(d[2], d[2], __fmaf_rn(d[1], d[1], __fmul_rn(d[0], d[0]))));
#pragma unroll
    for (int k = 0; k < 3; ++k) d[k] = __fdiv_rn(d[k], nrm);
    const float od = dot3_seq(P.o, d), dd = dot3_seq(d, d), oo = dot3_seq(P.o, P.o);
    const float det = __fadd_rn(__fmul_rn(od, od), __fmul_rn(__fsub_rn(P.r2, oo), dd));
    if (!(det >= 0.0f)) atomicAdd(P.bad, 1);
    const float far = __fdiv_rn(__fsub_rn(__fsqrt_rn(det), od), dd);
    float4* out = reinterpret_cast<float4*>(P.rays + idx * 8);
    out[0] = make_float4(P.o[0], P.o[1], P.o[2], d[0]);
    out[1] = make_float4(d[1], d[2], P.near, far);
}

}  // namespace
}  // namespace clift

using namespace clift;

extern "C" int32_t clift_gen_rays(const float* K, const float* c2w, int32_t height, int32_t width, float near, float radius,
                                  float* rays, int32_t* bad_rays, void* stream) {
    CLIFT_CHECK_ARG(K && c2w && rays && bad_rays && height > 0 && width > 0, "null pointer or non-positive size");
    CLIFT_CHECK_ARG((reinterpret_cast<uintptr_t>(rays) & 15) == 0, "rays must be 16-byte aligned");
    RayGenParams P;
    P.fx = K[0];
    P.fy = K[4];
    P.cx = K[2];
    P.cy = K[5];
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) P.R[r * 3 + c] = c2w[r * 4 + c];
        P.o[r] = c2w[r * 4 + 3];
    }
    P.near = near;
    P.r2 = (float)((double)radius * (double)radius);
    P.H = height;
    P.W = width;
    P.rays = rays;
    P.bad = bad_rays;
    CLIFT_CUDA(cudaMemsetAsync(bad_rays, 0, sizeof(int32_t), (cudaStream_t)stream));
    gen_rays_kernel<<<(unsigned)ceil_div((int64_t)height * width, 256), 256, 0, (cudaStream_t)stream>>>(P);
    CLIFT_AFTER_LAUNCH("gen_rays_kernel");
    return CLIFT_OK;
}
