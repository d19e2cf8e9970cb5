// Rows of SURVEY.md section 8(f) that sit beside the per-step render path:
//   rank 2  fused Adam step over a table of tensors          trainer/__init__.py:134-139, trainer:98-103,199,221
//   rank 3  dense-alpha sweep -> 3x3x3 max-pool -> bounding box   renderer:668-729
//           bilinear (align_corners) factor upsampling            tensoRF.py:179-197
//   rank 4  nearest-centroid assignment of rendered embeddings    inference/render_panopli.py:371-419
// All are HBM-bound streaming kernels: one coalesced pass over their input, grid sized from the SM count.
#include "launchers.h"

namespace clift {
namespace {

// ---------------------------------------------------------------------------------------------------------
// Adam (torch.optim.Adam, amsgrad=False, maximize=False): the arithmetic of torch's _single_tensor_adam
//   g = grad * grad_scale (+ wd * p);  m += (g - m) * (1 - b1);  v = v * b2 + (1 - b2) * g * g
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// One launch covers every tensor of a parameter group: blockIdx.y walks the table.
// ---------------------------------------------------------------------------------------------------------
struct AdamScalars {
    float one_minus_b1, b2, one_minus_b2, step_size, inv_bc2_sqrt, eps, wd, grad_scale;
};

// Up to CLIFT_MAX_ADAM_GROUPS param groups in one launch: tensor blockIdx.y belongs to the group whose [first, first + count)
// range holds it and is updated with that group's scalars.
struct AdamGroups {
    AdamScalars a[CLIFT_MAX_ADAM_GROUPS];
    int first[CLIFT_MAX_ADAM_GROUPS + 1];      // first[n_groups] = n_tensors
    int n_groups;
};

__device__ __forceinline__ void adam_update(const clift_adam_tensor& t, const AdamScalars& a);

__global__ void __launch_bounds__(256) adam_kernel(const clift_adam_tensor* __restrict__ table, AdamScalars a) {
    adam_update(table[blockIdx.y], a);
}

__global__ void __launch_bounds__(256) adam_groups_kernel(const clift_adam_tensor* __restrict__ table,
                                                          const __grid_constant__ AdamGroups G) {
    int g = 0;
    while (g + 1 < G.n_groups && (int)blockIdx.y >= G.first[g + 1]) ++g;
    adam_update(table[blockIdx.y], G.a[g]);
}

__device__ __forceinline__ void adam_update(const clift_adam_tensor& t, const AdamScalars& a) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n4 = ((reinterpret_cast<uintptr_t>(t.param) | reinterpret_cast<uintptr_t>(t.grad) |
                         reinterpret_cast<uintptr_t>(t.exp_avg) | reinterpret_cast<uintptr_t>(t.exp_avg_sq)) & 15) == 0
                           ? t.n / 4 : 0;
    auto one = [&](float& p, float g, float& m, float& v) {
        g *= a.grad_scale;
        if (a.wd != 0.0f) g = fmaf(a.wd, p, g);
        m = fmaf(g - m, a.one_minus_b1, m);
        v = fmaf(a.one_minus_b2 * g, g, v * a.b2);
        const float denom = sqrtf(v) * a.inv_bc2_sqrt + a.eps;
        p -= a.step_size * (m / denom);
    };
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 p = reinterpret_cast<float4*>(t.param)[i];
        const float4 g = __ldg(reinterpret_cast<const float4*>(t.grad) + i);
        float4 m = reinterpret_cast<float4*>(t.exp_avg)[i];
        float4 v = reinterpret_cast<float4*>(t.exp_avg_sq)[i];
        one(p.x, g.x, m.x, v.x);
        one(p.y, g.y, m.y, v.y);
        one(p.z, g.z, m.z, v.z);
        one(p.w, g.w, m.w, v.w);
        reinterpret_cast<float4*>(t.param)[i] = p;
        reinterpret_cast<float4*>(t.exp_avg)[i] = m;
        reinterpret_cast<float4*>(t.exp_avg_sq)[i] = v;
    }
    for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < t.n; i += stride)
        one(t.param[i], t.grad[i], t.exp_avg[i], t.exp_avg_sq[i]);
}

// ---------------------------------------------------------------------------------------------------------
// dense alpha (renderer:717-729, 744-748): voxel (i,j,k) of the grid_dim lattice sits at
//   p = aabb0 * (1 - s) + aabb1 * s   (s = the caller's torch.linspace(0,1,G) per axis; separately rounded ops)
// alpha = 1 - exp(-softplus(VM(p) + shift) * step_size); stored [G0][G1][G2] (k fastest).  Quad per voxel.
// ---------------------------------------------------------------------------------------------------------
struct AlphaParams {
    GeomParams g;
    FactorParams f;
    float shift;
    const float* s[3];
    int n[3];
    float* alpha;
};

__device__ __forceinline__ float lattice_coord(float a0, float a1, float s) {
    return __fadd_rn(__fmul_rn(a0, __fsub_rn(1.0f, s)), __fmul_rn(a1, s));
}

template <int NV>
__global__ void __launch_bounds__(256) dense_alpha_kernel(const __grid_constant__ AlphaParams P) {
    const int64_t total = (int64_t)P.n[0] * P.n[1] * P.n[2];
    const int64_t vox = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const int q = threadIdx.x & 3;
    float acc = 0.0f;
    if (vox < total) {
        const int k = (int)(vox % P.n[2]), j = (int)((vox / P.n[2]) % P.n[1]), i = (int)(vox / ((int64_t)P.n[2] * P.n[1]));
        const int idx[3] = {i, j, k};
        float x[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float p = lattice_coord(P.g.amin[c], P.g.amax[c], __ldg(P.s[c] + idx[c]));
            x[c] = __fsub_rn(__fmul_rn(__fsub_rn(p, P.g.amin[c]), P.g.inv[c]), 1.0f);      // renderer:633-634
        }
        acc = vm_dot_partial<NV>(P.f, nullptr, x[0], x[1], x[2], q);
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (vox < total && q == 0) P.alpha[vox] = 1.0f - expf(-softplus_f(acc + P.shift) * P.g.step);
}

// ---------------------------------------------------------------------------------------------------------
// renderer:671-681: clamp(alpha,0,1) -> max_pool3d(3, stride 1, padding 1) -> >= threshold -> bounding box of the
// surviving lattice positions and their count.  Floats are reduced as order-preserving unsigned keys.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned key_of(float f) {
    const unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float float_of(unsigned k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

struct BboxParams {
    const float* alpha;
    const float* s[3];
    int n[3];
    float amin[3], amax[3];
    float threshold;
    unsigned* keys;     // [0..2] min keys, [3..5] max keys, [6] count
};

__global__ void bbox_init_kernel(unsigned* keys) {
    if (threadIdx.x < 3) keys[threadIdx.x] = 0xffffffffu;
    else if (threadIdx.x < 7) keys[threadIdx.x] = 0u;
}

__global__ void __launch_bounds__(256) alpha_bbox_kernel(const __grid_constant__ BboxParams P) {
    const int64_t total = (int64_t)P.n[0] * P.n[1] * P.n[2];
    unsigned kmin[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, kmax[3] = {0u, 0u, 0u}, cnt = 0;
    for (int64_t vox = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; vox < total; vox += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(vox % P.n[2]), j = (int)((vox / P.n[2]) % P.n[1]), i = (int)(vox / ((int64_t)P.n[2] * P.n[1]));
        float m = -INFINITY;
        for (int di = max(i - 1, 0); di <= min(i + 1, P.n[0] - 1); ++di)
            for (int dj = max(j - 1, 0); dj <= min(j + 1, P.n[1] - 1); ++dj)
                for (int dk = max(k - 1, 0); dk <= min(k + 1, P.n[2] - 1); ++dk)
                    m = fmaxf(m, fminf(fmaxf(__ldg(P.alpha + ((int64_t)di * P.n[1] + dj) * P.n[2] + dk), 0.0f), 1.0f));
        if (m >= P.threshold) {
            const int idx[3] = {i, j, k};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const unsigned key = key_of(lattice_coord(P.amin[c], P.amax[c], __ldg(P.s[c] + idx[c])));
                kmin[c] = min(kmin[c], key);
                kmax[c] = max(kmax[c], key);
            }
            ++cnt;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            kmin[c] = min(kmin[c], __shfl_xor_sync(0xffffffffu, kmin[c], o));
            kmax[c] = max(kmax[c], __shfl_xor_sync(0xffffffffu, kmax[c], o));
        }
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if ((threadIdx.x & 31) == 0 && cnt) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            atomicMin(P.keys + c, kmin[c]);
            atomicMax(P.keys + 3 + c, kmax[c]);
        }
        atomicAdd(P.keys + 6, cnt);
    }
}

__global__ void bbox_finish_kernel(const unsigned* keys, float* bbox6, int32_t* count) {
    if (threadIdx.x < 6) bbox6[threadIdx.x] = keys[6] ? float_of(keys[threadIdx.x]) : 0.0f;
    if (threadIdx.x == 6) *count = (int32_t)keys[6];
}

// ---------------------------------------------------------------------------------------------------------
// F.interpolate(mode="bilinear", align_corners=True) on (1,C,H,W) (tensoRF.py:190-197); lines are the W=1 case.
// ATen: scale = (in-1)/(out-1) (0 if out == 1); src = scale*dst; i0 = (int)src; l1 = src - i0; l0 = 1 - l1;
//       out = l0h*(l0w*v00 + l1w*v01) + l1h*(l0w*v10 + l1w*v11)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) upsample_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int H, int W,
                                                       int H2, int W2, float sh, float sw) {
    const int64_t total = (int64_t)C * H2 * W2;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(o % W2), y = (int)((o / W2) % H2), c = (int)(o / ((int64_t)W2 * H2));
        const float fy = __fmul_rn(sh, (float)y), fx = __fmul_rn(sw, (float)x);
        const int y0 = (int)fy, x0 = (int)fx;
        const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
        const float ly1 = __fsub_rn(fy, (float)y0), lx1 = __fsub_rn(fx, (float)x0);
        const float ly0 = __fsub_rn(1.0f, ly1), lx0 = __fsub_rn(1.0f, lx1);
        const float* p = src + (int64_t)c * H * W;
        const float v00 = __ldg(p + (int64_t)y0 * W + x0), v01 = __ldg(p + (int64_t)y0 * W + x1);
        const float v10 = __ldg(p + (int64_t)y1 * W + x0), v11 = __ldg(p + (int64_t)y1 * W + x1);
        const float top = __fadd_rn(__fmul_rn(lx0, v00), __fmul_rn(lx1, v01));
        const float bot = __fadd_rn(__fmul_rn(lx0, v10), __fmul_rn(lx1, v11));
        dst[o] = __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
    }
}

// ---------------------------------------------------------------------------------------------------------
// nearest centroid (render_panopli.py:389-396: torch.cdist p=2 then argmin; ties -> lowest index).  Centroids in
// shared memory, one thread per point, squared distances (same argmin as the Euclidean distance).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) centroid_kernel(const float* __restrict__ feats, int64_t n, int d, int feat_stride,
                                                       const float* __restrict__ centroids, int k, int32_t* __restrict__ labels,
                                                       float* __restrict__ dist) {
    extern __shared__ float cs[];
    for (int i = threadIdx.x; i < k * d; i += blockDim.x) cs[i] = centroids[i];
    __syncthreads();
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        float f[16];
        for (int c = 0; c < d; ++c) f[c] = feats[p * feat_stride + c];
        float best = INFINITY;
        int arg = 0;
        for (int j = 0; j < k; ++j) {
            float acc = 0.0f;
            for (int c = 0; c < d; ++c) {
                const float diff = f[c] - cs[j * d + c];
                acc = fmaf(diff, diff, acc);
            }
            if (acc < best) {
                best = acc;
                arg = j;
            }
        }
        labels[p] = arg;
        if (dist) dist[p] = sqrtf(best);
    }
}

// ---------------------------------------------------------------------------------------------------------
// panoptic instance labels from rendered embeddings (the device form of render_panopli.py:371-419): every "thing" point
// takes the nearest centroid of ITS semantic class; classes present in the data get consecutive label ranges in ascending
// class order, each range as long as the highest centroid index any point of the class chose, + 1; stuff points get label 0.
//   pass 1: class = argmax of the semantic scores (first maximum), nearest centroid inside the class's table slice,
//           code = class << 20 | local index (-1 = stuff), atomicMax of local + 1 per class
//   pass 2: one thread walks the classes: first label of each present class (exclusive running sum), total label count
//   pass 3: label = first[class] + local + 1; optional one-hot rows (fp64, as the reference's numpy one-hot)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cluster_local_kernel(const float* __restrict__ feats, int64_t n, int d, int feat_stride,
                                                            const float* __restrict__ scores, int n_cls,
                                                            const float* __restrict__ centroids,
                                                            const int32_t* __restrict__ cls_first,
                                                            const int32_t* __restrict__ cls_count, int32_t* __restrict__ code,
                                                            uint32_t* __restrict__ cls_top, int32_t* __restrict__ missing) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const float* row = feats + p * feat_stride;
        if (!(row[0] == -INFINITY)) {        // column 0 marks thing points with -inf (create_instances_from_semantics)
            code[p] = -1;
            continue;
        }
        int c = 0;
        float top = scores[p * n_cls];
        for (int j = 1; j < n_cls; ++j) {
            const float s = scores[p * n_cls + j];
            if (s > top) {
                top = s;
                c = j;
            }
        }
        const int k = cls_count[c];
        if (k <= 0) {                         // the reference raises KeyError here; reported through `missing`
            atomicMax(missing, c + 1);
            code[p] = -1;
            continue;
        }
        float f[16];
        for (int i = 0; i < d; ++i) f[i] = row[1 + i];
        const float* cs = centroids + (int64_t)cls_first[c] * d;
        float best = INFINITY;
        int arg = 0;
        for (int j = 0; j < k; ++j) {
            float acc = 0.0f;
            for (int i = 0; i < d; ++i) {
                const float diff = f[i] - __ldg(cs + j * d + i);
                acc = fmaf(diff, diff, acc);
            }
            if (acc < best) {
                best = acc;
                arg = j;
            }
        }
        code[p] = (c << 20) | arg;
        atomicMax(cls_top + c, (uint32_t)arg + 1u);
    }
}

__global__ void cluster_ranges_kernel(uint32_t* __restrict__ cls_top, int n_cls, int32_t* __restrict__ n_labels) {
    uint32_t run = 0;
    for (int c = 0; c < n_cls; ++c) {         // cls_top[c]: range length in, first label of the class out
        const uint32_t len = cls_top[c];
        cls_top[c] = run;
        run += len;
    }
    *n_labels = (int32_t)run + 1;             // + the stuff label 0
}

__global__ void __launch_bounds__(256) cluster_label_kernel(const int32_t* __restrict__ code, int64_t n,
                                                            const uint32_t* __restrict__ cls_first_label,
                                                            int32_t* __restrict__ labels) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const int32_t v = code[p];
        labels[p] = v < 0 ? 0 : (int32_t)cls_first_label[v >> 20] + (v & 0xfffff) + 1;
    }
}

__global__ void __launch_bounds__(256) onehot_kernel(const int32_t* __restrict__ labels, int64_t n, int width,
                                                     double* __restrict__ out) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const int32_t l = labels[p];
        if (l >= 0 && l < width) out[p * width + l] = 1.0;
    }
}

}  // namespace
}  // namespace clift

using namespace clift;

extern "C" int32_t clift_assign_clusters(const float* features, int64_t n, int32_t dim, int32_t feature_stride,
                                         const float* scores, int32_t n_classes, const float* centroids,
                                         const int32_t* class_first, const int32_t* class_count, int32_t* labels,
                                         uint32_t* class_scratch, int32_t* stats2, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CLIFT_CHECK_ARG(n >= 0 && dim > 0 && feature_stride >= dim + 1 && n_classes > 0, "bad size");
    CLIFT_CHECK_ARG(class_first && class_count && class_scratch && stats2, "null pointer");
    CLIFT_CHECK_SUPPORTED(dim <= 16 && n_classes <= 2048, "embedding wider than 16 or more than 2048 classes");
    CLIFT_CUDA(cudaMemsetAsync(class_scratch, 0, (size_t)n_classes * sizeof(uint32_t), stream));
    CLIFT_CUDA(cudaMemsetAsync(stats2, 0, 2 * sizeof(int32_t), stream));
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, 256), 8 * sm_count()));
    if (n > 0) {
        CLIFT_CHECK_ARG(features && scores && centroids && labels, "null pointer");
        cluster_local_kernel<<<grid, 256, 0, stream>>>(features, n, dim, feature_stride, scores, n_classes, centroids, class_first,
                                                       class_count, labels, class_scratch, stats2 + 1);
        CLIFT_AFTER_LAUNCH("cluster_local_kernel");
    }
    cluster_ranges_kernel<<<1, 1, 0, stream>>>(class_scratch, n_classes, stats2);
    CLIFT_AFTER_LAUNCH("cluster_ranges_kernel");
    if (n > 0) {
        cluster_label_kernel<<<grid, 256, 0, stream>>>(labels, n, class_scratch, labels);
        CLIFT_AFTER_LAUNCH("cluster_label_kernel");
    }
    return CLIFT_OK;
}

extern "C" int32_t clift_labels_onehot(const int32_t* labels, int64_t n, int32_t width, double* onehot, void* stream) {
    CLIFT_CHECK_ARG(n >= 0 && width > 0, "bad size");
    if (n == 0) return CLIFT_OK;
    CLIFT_CHECK_ARG(labels && onehot, "null pointer");
    CLIFT_CUDA(cudaMemsetAsync(onehot, 0, (size_t)n * width * sizeof(double), (cudaStream_t)stream));
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, 256), 8 * sm_count()));
    onehot_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(labels, n, width, onehot);
    CLIFT_AFTER_LAUNCH("onehot_kernel");
    return CLIFT_OK;
}

extern "C" int32_t clift_adam_step(const clift_adam_tensor* table_dev, int32_t n_tensors, int64_t max_n, float lr, float beta1,
                                   float beta2, float eps, float weight_decay, int64_t step, float grad_scale, void* stream) {
    CLIFT_CHECK_ARG(table_dev && n_tensors >= 0 && max_n >= 0 && step >= 1, "null table, negative size or step < 1");
    if (n_tensors == 0 || max_n == 0) return CLIFT_OK;
    CLIFT_CHECK_SUPPORTED(n_tensors <= 65535, "more than 65535 tensors in one call");
    // scalars as torch computes them (Python doubles, then one rounding to fp32)
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    AdamScalars a;
    a.one_minus_b1 = (float)(1.0 - (double)beta1);
    a.b2 = beta2;
    a.one_minus_b2 = (float)(1.0 - (double)beta2);
    a.step_size = (float)((double)lr / bc1);
    a.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
    a.eps = eps;
    a.wd = weight_decay;
    a.grad_scale = grad_scale;
    const unsigned gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(max_n, 256 * 4), 4 * sm_count()));
    adam_kernel<<<dim3(gx, (unsigned)n_tensors), 256, 0, (cudaStream_t)stream>>>(table_dev, a);
    CLIFT_AFTER_LAUNCH("adam_kernel");
    return CLIFT_OK;
}

extern "C" int32_t clift_adam_step_groups(const clift_adam_tensor* table_dev, int32_t n_tensors, int64_t max_n,
                                          const clift_adam_group* groups, int32_t n_groups, float grad_scale, void* stream) {
    CLIFT_CHECK_ARG(n_tensors >= 0 && max_n >= 0 && n_groups >= 0, "negative size");
    if (n_tensors == 0 || max_n == 0 || n_groups == 0) return CLIFT_OK;
    CLIFT_CHECK_ARG(table_dev && groups, "null table");
    CLIFT_CHECK_SUPPORTED(n_tensors <= 65535 && n_groups <= CLIFT_MAX_ADAM_GROUPS, "more than 65535 tensors or too many groups in one call");
    AdamGroups G;
    memset(&G, 0, sizeof(G));
    G.n_groups = n_groups;
    int next = 0;
    for (int g = 0; g < n_groups; ++g) {
        const clift_adam_group& h = groups[g];
        CLIFT_CHECK_ARG(h.step >= 1 && h.count >= 0 && h.first == next, "groups must tile the table in order, step >= 1");
        next += h.count;
        // scalars as torch computes them (Python doubles, then one rounding to fp32)
        const double bc1 = 1.0 - pow((double)h.beta1, (double)h.step), bc2 = 1.0 - pow((double)h.beta2, (double)h.step);
        AdamScalars& a = G.a[g];
        a.one_minus_b1 = (float)(1.0 - (double)h.beta1);
        a.b2 = h.beta2;
        a.one_minus_b2 = (float)(1.0 - (double)h.beta2);
        a.step_size = (float)((double)h.lr / bc1);
        a.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
        a.eps = h.eps;
        a.wd = h.weight_decay;
        a.grad_scale = grad_scale;
        G.first[g] = h.first;
    }
    CLIFT_CHECK_ARG(next == n_tensors, "groups do not cover the table");
    G.first[n_groups] = n_tensors;
    const unsigned gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(max_n, 256 * 4), 4 * sm_count()));
    adam_groups_kernel<<<dim3(gx, (unsigned)n_tensors), 256, 0, (cudaStream_t)stream>>>(table_dev, G);
    CLIFT_AFTER_LAUNCH("adam_groups_kernel");
    return CLIFT_OK;
}

extern "C" int32_t clift_dense_alpha(const clift_render_cfg* cfg, const clift_field* field, const float* sx, const float* sy,
                                     const float* sz, float* alpha, void* stream) {
    CLIFT_CHECK_ARG(cfg && field && sx && sy && sz && alpha, "null pointer");
    AlphaParams P;
    P.g = make_geom(cfg);
    P.f = make_factors(field, false);
    P.shift = field->density_shift;
    P.s[0] = sx;
    P.s[1] = sy;
    P.s[2] = sz;
    for (int c = 0; c < 3; ++c) P.n[c] = field->grid[c];
    P.alpha = alpha;
    const int64_t total = (int64_t)P.n[0] * P.n[1] * P.n[2];
    CLIFT_CHECK_ARG(total > 0 && total * 4 < (1ll << 40), "empty or oversized grid");
    const unsigned grid = (unsigned)ceil_div(total * 4, 256);
    switch (P.f.comps / 16) {
        case 1: dense_alpha_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(P); break;
        case 2: dense_alpha_kernel<2><<<grid, 256, 0, (cudaStream_t)stream>>>(P); break;
        case 3: dense_alpha_kernel<3><<<grid, 256, 0, (cudaStream_t)stream>>>(P); break;
        default:
            set_error("clift_dense_alpha: density_comps %d not in {16,32,48}", P.f.comps);
            return CLIFT_ERR_UNSUPPORTED;
    }
    CLIFT_AFTER_LAUNCH("dense_alpha_kernel");
    return CLIFT_OK;
}

extern "C" int32_t clift_alpha_bbox(const float* alpha, const int32_t* grid3, const float* sx, const float* sy, const float* sz,
                                    const float* aabb_min3, const float* aabb_max3, float threshold, float* bbox6, int32_t* count,
                                    uint32_t* scratch8, void* stream) {
    CLIFT_CHECK_ARG(alpha && grid3 && sx && sy && sz && aabb_min3 && aabb_max3 && bbox6 && count && scratch8, "null pointer");
    BboxParams P;
    P.alpha = alpha;
    P.s[0] = sx;
    P.s[1] = sy;
    P.s[2] = sz;
    for (int c = 0; c < 3; ++c) {
        P.n[c] = grid3[c];
        P.amin[c] = aabb_min3[c];
        P.amax[c] = aabb_max3[c];
        CLIFT_CHECK_ARG(grid3[c] > 0, "non-positive grid size");
    }
    P.threshold = threshold;
    P.keys = scratch8;
    cudaStream_t st = (cudaStream_t)stream;
    bbox_init_kernel<<<1, 32, 0, st>>>(scratch8);
    CLIFT_AFTER_LAUNCH("bbox_init_kernel");
    const int64_t total = (int64_t)P.n[0] * P.n[1] * P.n[2];
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(total, 256), 8 * sm_count()));
    alpha_bbox_kernel<<<grid, 256, 0, st>>>(P);
    CLIFT_AFTER_LAUNCH("alpha_bbox_kernel");
    bbox_finish_kernel<<<1, 32, 0, st>>>(scratch8, bbox6, count);
    CLIFT_AFTER_LAUNCH("bbox_finish_kernel");
    return CLIFT_OK;
}

extern "C" int32_t clift_upsample_bilinear(const float* src, float* dst, int32_t channels, int32_t h, int32_t w, int32_t h2,
                                           int32_t w2, void* stream) {
    CLIFT_CHECK_ARG(src && dst && channels > 0 && h > 0 && w > 0 && h2 > 0 && w2 > 0, "null pointer or non-positive size");
    const float sh = h2 > 1 ? (float)(h - 1) / (float)(h2 - 1) : 0.0f, sw = w2 > 1 ? (float)(w - 1) / (float)(w2 - 1) : 0.0f;
    const int64_t total = (int64_t)channels * h2 * w2;
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(total, 256), 8 * sm_count()));
    upsample_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, dst, channels, h, w, h2, w2, sh, sw);
    CLIFT_AFTER_LAUNCH("upsample_kernel");
    return CLIFT_OK;
}

extern "C" int32_t clift_assign_centroids(const float* features, int64_t n, int32_t dim, int32_t feature_stride,
                                          const float* centroids, int32_t k, int32_t* labels, float* distances, void* stream) {
    CLIFT_CHECK_ARG(n >= 0 && k > 0 && dim > 0 && feature_stride >= dim, "bad size");
    if (n == 0) return CLIFT_OK;
    CLIFT_CHECK_ARG(features && centroids && labels, "null pointer");
    CLIFT_CHECK_SUPPORTED(dim <= 16 && (int64_t)k * dim * 4 <= 48 * 1024, "embedding wider than 16 or centroid table above 48 KB");
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, 256), 8 * sm_count()));
    centroid_kernel<<<grid, 256, (size_t)k * dim * 4, (cudaStream_t)stream>>>(features, n, dim, feature_stride, centroids, k, labels,
                                                                             distances);
    CLIFT_AFTER_LAUNCH("centroid_kernel");
    return CLIFT_OK;
}
