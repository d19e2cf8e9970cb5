// Contrastive-fusion losses on rendered embeddings, each as ONE fused forward+gradient kernel.
//   slow-fast loss  : trainer/train_panopli_tensorf.py:256-310   (N <= ~1024 rays of one image)
//   EMA of slow net : trainer/train_panopli_tensorf.py:325-329
//   vanilla loss    : model/loss/loss.py:62-82
//   TV regulariser  : model/loss/loss.py:9-26 (tensoRF.py:248-290 supplies the 1e-2*lambda scale)
// The reference needs torch.unique, per-label Python loops, boolean-mask indexing (host syncs) and an
// (N/2)^2 cdist; here every CTA (the slow-fast loss: of one thread-block cluster) keeps features/labels in shared memory,
// resolves label groups by first-occurrence scans, and all reductions use a fixed tree so the result is run-to-run identical.
#include "launchers.h"

namespace clift {
namespace {

constexpr int kLossThreads = 1024;

// thread-block cluster helpers (PTX: no cooperative_groups dependency)
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned cluster_nctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_dsmem_f32(const float* local_smem_ptr, unsigned rank) {
    const unsigned local = (unsigned)__cvta_generic_to_shared(local_smem_ptr);
    unsigned remote;
    float v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(rank));
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
    return v;
}

__device__ __forceinline__ float block_sum(float v, float* s_red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    float t = (threadIdx.x < (blockDim.x >> 5)) ? s_red[threadIdx.x] : 0.0f;
    if (warp == 0) {
        t = warp_sum(t);
        if (lane == 0) s_red[32] = t;
    }
    __syncthreads();
    return s_red[32];
}

// ----------------------------------------------------------------------------------------------
// slow-fast
// ----------------------------------------------------------------------------------------------
// One thread-block cluster (1, 8 or 16 CTAs by problem size); every CTA holds all features / labels in shared memory and
// owns the fast samples i = rank*32 + warp, + 32*cluster size, ...: ONE WARP per fast sample, its lanes split the slow
// samples j (lane, lane + 32, ...) and combine with a fixed xor tree, so the N/2 x N/2 pair loop that a single CTA walked in
// 512-long dependent chains is 16 x 32 = 512 warps wide.  The four loss sums are combined in rank order through distributed
// shared memory (fixed order: run-to-run identical), after which every CTA knows the totals the gradient needs.
__device__ __forceinline__ float warp_sum_xor(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(kLossThreads) slowfast_kernel(const float* __restrict__ feats, const long long* __restrict__ labels,
                                                                const float* __restrict__ conf, int n, int d,
                                                                float* __restrict__ loss_out, float* __restrict__ grad) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float s_red[33];
    __shared__ float s_part[4];
    const int nf = n / 2, ns = n - nf, w2 = 2 * d;
    const unsigned rank = cluster_ctarank(), csize = cluster_nctarank();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    long long* s_lab = reinterpret_cast<long long*>(smem_raw);      // [n]
    float* s_feat = reinterpret_cast<float*>(s_lab + n);              // [n][2d]
    float* s_cent = s_feat + (size_t)n * w2;                          // [ns][d] centroid of slow sample j's label group
    float* s_num = s_cent + (size_t)ns * d;                           // [nf]
    float* s_den = s_num + nf;                                        // [nf]
    float* s_cw = s_den + nf;                                         // [nf] concentration weight conf*e/(n_l), 0 if label unmatched
    int* s_rep = reinterpret_cast<int*>(s_cw + nf);                   // [nf] slow representative (absolute index) or -1

    for (int i = threadIdx.x; i < n; i += blockDim.x) s_lab[i] = labels[i];
    for (int i = threadIdx.x; i < n * w2; i += blockDim.x) s_feat[i] = feats[i];
    if (grad)      // slow rows carry no gradient (trainer:267: the slow half is detached); fast rows are written whole below
        for (int i = nf * w2 + (int)rank * (int)blockDim.x + threadIdx.x; i < n * w2; i += (int)csize * blockDim.x) grad[i] = 0.0f;
    __syncthreads();
    if (nf == 0 || ns == 0) {   // trainer:285-288
        if (rank == 0 && threadIdx.x == 0) loss_out[0] = 0.0f;
        if (grad)
            for (int i = (int)rank * (int)blockDim.x + threadIdx.x; i < nf * w2; i += (int)csize * blockDim.x) grad[i] = 0.0f;
        return;
    }
    // slow centroids: every slow sample gets the mean of its label group (members summed in index order)
    for (int j = threadIdx.x; j < ns; j += blockDim.x) {
        const long long l = s_lab[nf + j];
        int cnt = 0;
        for (int k = 0; k < d; ++k) s_cent[j * d + k] = 0.0f;
        for (int jj = 0; jj < ns; ++jj)
            if (s_lab[nf + jj] == l) {
                ++cnt;
                for (int k = 0; k < d; ++k) s_cent[j * d + k] += s_feat[(nf + jj) * w2 + d + k];
            }
        for (int k = 0; k < d; ++k) s_cent[j * d + k] /= (float)cnt;
    }
    __syncthreads();
    float my_conc = 0.0f, my_logp = 0.0f, my_valid = 0.0f, my_first = 0.0f;      // lane 0 of each warp accumulates
    for (int i = (int)rank * n_warps + warp; i < nf; i += (int)csize * n_warps) {
        const long long l = s_lab[i];
        int rep = 0x7fffffff;
        for (int j = lane; j < ns; j += 32)
            if (s_lab[nf + j] == l) {
                rep = j;
                break;
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rep = min(rep, __shfl_xor_sync(0xffffffffu, rep, o));
        if (rep == 0x7fffffff) rep = -1;
        int n_l = 0, earlier = 0;
        for (int ii = lane; ii < nf; ii += 32)
            if (s_lab[ii] == l) {
                ++n_l;
                if (ii < i) earlier = 1;
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            n_l += __shfl_xor_sync(0xffffffffu, n_l, o);
            earlier |= __shfl_xor_sync(0xffffffffu, earlier, o);
        }
        float cw = 0.0f;
        if (rep >= 0) {
            float dsq = 0.0f;
            for (int k = 0; k < d; ++k) {
                const float df = s_feat[i * w2 + k] - s_cent[rep * d + k];
                dsq += df * df;
            }
            cw = expf(-dsq) * conf[i] / (float)n_l;
            if (lane == 0) {
                my_conc += cw;
                if (!earlier) my_first += 1.0f;
            }
        }
        float num = 0.0f, den = 0.0f;
        for (int j = lane; j < ns; j += 32) {
            float dsq = 0.0f;
            for (int k = 0; k < d; ++k) {
                const float df = s_feat[i * w2 + k] - s_feat[(nf + j) * w2 + d + k];
                dsq += df * df;
            }
            const float e = expf(expf(-sqrtf(dsq)));
            den += e;
            if (s_lab[nf + j] == l) num += e;
        }
        num = warp_sum_xor(num);
        den = warp_sum_xor(den);
        if (lane == 0) {
            s_cw[i] = cw;
            s_rep[i] = rep;
            s_num[i] = num;
            s_den[i] = den;
            const float prob = num / den;
            if (prob != 0.0f) {
                my_logp += logf(prob);
                my_valid += 1.0f;
            }
        }
    }
    const float b_conc = block_sum(my_conc, s_red);
    const float b_first = block_sum(my_first, s_red);
    const float b_logp = block_sum(my_logp, s_red);
    const float b_valid = block_sum(my_valid, s_red);
    if (threadIdx.x == 0) {
        s_part[0] = b_conc;
        s_part[1] = b_first;
        s_part[2] = b_logp;
        s_part[3] = b_valid;
    }
    cluster_sync_all();
    float conc = 0.0f, n_int = 0.0f, logp = 0.0f, n_valid = 0.0f;
    for (unsigned r = 0; r < csize; ++r) {        // rank order: the same totals, bit for bit, in every CTA
        conc += ld_dsmem_f32(&s_part[0], r);
        n_int += ld_dsmem_f32(&s_part[1], r);
        logp += ld_dsmem_f32(&s_part[2], r);
        n_valid += ld_dsmem_f32(&s_part[3], r);
    }
    cluster_sync_all();                            // peers' partials stay alive until everybody has read them
    if (rank == 0 && threadIdx.x == 0) {
        float l = 0.0f;
        if (n_int > 0.0f) l = -conc / n_int;
        l += -(logp / n_valid);      // 0/0 -> NaN like mean() of an empty selection
        loss_out[0] = l;
    }
    if (!grad) return;
    for (int i = (int)rank * n_warps + warp; i < nf; i += (int)csize * n_warps) {
        const long long l = s_lab[i];
        const float num = s_num[i], den = s_den[i];
        const bool valid = (num / den) != 0.0f;
        for (int k = 0; k < d; ++k) {
            float g = 0.0f;
            if (valid)
                for (int j = lane; j < ns; j += 32) {
                    float dsq = 0.0f;
                    for (int kk = 0; kk < d; ++kk) {
                        const float df = s_feat[i * w2 + kk] - s_feat[(nf + j) * w2 + d + kk];
                        dsq += df * df;
                    }
                    const float dist = sqrtf(dsq);
                    if (dist == 0.0f) continue;   // cdist backward yields 0 at coincident points
                    const float sx = expf(-dist), e = expf(sx);
                    const float c = ((s_lab[nf + j] == l ? 1.0f / num : 0.0f) - 1.0f / den) * e * sx / (dist * n_valid);
                    g += c * (s_feat[i * w2 + k] - s_feat[(nf + j) * w2 + d + k]);
                }
            g = warp_sum_xor(g);
            if (lane == 0) {
                if (s_rep[i] >= 0) g += 2.0f * s_cw[i] / n_int * (s_feat[i * w2 + k] - s_cent[s_rep[i] * d + k]);
                grad[i * w2 + k] = g;
                grad[i * w2 + d + k] = 0.0f;
            }
        }
    }
}

// ----------------------------------------------------------------------------------------------
// vanilla contrastive (loss.py:62-82)
// ----------------------------------------------------------------------------------------------
// Same organisation as the slow-fast kernel: one thread-block cluster, every CTA holds all features / labels, one warp per
// sample i with its lanes splitting the partners j.  The gradient of sample i needs 1/p_j and 1/Z_j of every partner, so after
// the first pass the CTAs all-gather those two vectors through distributed shared memory (each reads the entries its peers own).
__global__ void __launch_bounds__(kLossThreads) contrastive_kernel(const float* __restrict__ feats, const long long* __restrict__ labels,
                                                                   int n, int d, float temperature, float* __restrict__ loss_out,
                                                                   float* __restrict__ grad) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float s_red[33];
    __shared__ float s_part[1];
    const unsigned rank = cluster_ctarank(), csize = cluster_nctarank();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    long long* s_lab = reinterpret_cast<long long*>(smem_raw);   // [n]
    float* s_feat = reinterpret_cast<float*>(s_lab + n);           // [n][d]
    float* s_a = s_feat + (size_t)n * d;                           // [n] 1/p_i if i contributes else 0
    float* s_b = s_a + n;                                          // [n] 1/Z_i if i contributes else 0
    for (int i = threadIdx.x; i < n; i += blockDim.x) s_lab[i] = labels[i];
    for (int i = threadIdx.x; i < n * d; i += blockDim.x) s_feat[i] = feats[i];
    __syncthreads();
    float my_log = 0.0f;                                           // lane 0 of each warp accumulates
    for (int i = (int)rank * n_warps + warp; i < n; i += (int)csize * n_warps) {
        const long long l = s_lab[i];
        float p = 0.0f, z = 0.0f;
        for (int j = lane; j < n; j += 32) {
            float dsq = 0.0f;
            for (int k = 0; k < d; ++k) {
                const float df = s_feat[i * d + k] - s_feat[j * d + k];
                dsq += df * df;
            }
            const bool pos = (j != i) && (s_lab[j] == l);
            const float e = expf(expf(-dsq / (pos ? temperature : 1.0f)));
            z += e;
            if (pos) p += e;
        }
        p = warp_sum_xor(p);
        z = warp_sum_xor(z);
        if (lane == 0) {
            const float prob = p / z;
            const bool valid = prob != 0.0f;
            if (valid) my_log += logf(prob);
            s_a[i] = valid ? 1.0f / p : 0.0f;
            s_b[i] = valid ? 1.0f / z : 0.0f;
        }
    }
    const float b_log = block_sum(my_log, s_red);
    if (threadIdx.x == 0) s_part[0] = b_log;
    cluster_sync_all();
    // all-gather 1/p and 1/Z: sample i belongs to CTA (i / n_warps) % csize
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const unsigned owner = (unsigned)(i / n_warps) % csize;
        if (owner != rank) {
            s_a[i] = ld_dsmem_f32(&s_a[i], owner);
            s_b[i] = ld_dsmem_f32(&s_b[i], owner);
        }
    }
    float tot = 0.0f;
    for (unsigned r = 0; r < csize; ++r) tot += ld_dsmem_f32(&s_part[0], r);      // rank order: fixed
    cluster_sync_all();                                            // peers' vectors stay alive until everybody has read them
    if (rank == 0 && threadIdx.x == 0) loss_out[0] = -tot / (float)n;
    if (!grad) return;
    for (int i = (int)rank * n_warps + warp; i < n; i += (int)csize * n_warps) {
        const long long l = s_lab[i];
        for (int k = 0; k < d; ++k) {
            float g = 0.0f;
            for (int j = lane; j < n; j += 32) {
                if (j == i) continue;
                float dsq = 0.0f;
                for (int kk = 0; kk < d; ++kk) {
                    const float df = s_feat[i * d + kk] - s_feat[j * d + kk];
                    dsq += df * df;
                }
                const bool pos = s_lab[j] == l;
                const float temp = pos ? temperature : 1.0f;
                const float sx = expf(-dsq / temp), e = expf(sx);
                const float m = pos ? 1.0f : 0.0f;
                const float c = e * sx * (2.0f / temp) * ((m * s_a[i] - s_b[i]) + (m * s_a[j] - s_b[j])) / (float)n;
                g += c * (s_feat[i * d + k] - s_feat[j * d + k]);
            }
            g = warp_sum_xor(g);
            if (lane == 0) grad[i * d + k] = g;
        }
    }
}

// ----------------------------------------------------------------------------------------------
// EMA, TV
// ----------------------------------------------------------------------------------------------
__global__ void ema_kernel(float* __restrict__ slow, const float* __restrict__ fast, int64_t n, float m, float om) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) slow[i] = __fadd_rn(__fmul_rn(slow[i], m), __fmul_rn(om, fast[i]));   // mul_ then add_ of a scaled copy
}

// every (slow, fast) pair of a network in one launch: blockIdx.y = pair, same arithmetic as ema_kernel
__global__ void ema_batch_kernel(const clift_ema_pair* __restrict__ table, float m, float om) {
    const clift_ema_pair t = table[blockIdx.y];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < t.n; i += (int64_t)gridDim.x * blockDim.x)
        t.slow[i] = __fadd_rn(__fmul_rn(t.slow[i], m), __fmul_rn(om, t.fast[i]));
}

// One thread-block cluster of kTvCluster CTAs (fixed-order sum, no scratch memory, no atomics): every CTA reduces a strided
// share of the plane, parks its two partial sums in its own shared memory, and rank 0 adds the partials in rank order through
// distributed shared memory - run-to-run identical, 16x the bandwidth of a single CTA (one 128^2 x 48 plane: ~0.37 ms -> ~25 us).
constexpr int kTvCluster = 16;

__device__ __forceinline__ void tv_value_body(const float* __restrict__ x, int C, int H, int W, float* __restrict__ out);

__global__ void __launch_bounds__(1024) tv_value_kernel(const float* __restrict__ x, int C, int H, int W, float* __restrict__ out) {
    tv_value_body(x, C, H, W, out);
}

// clift_tv_loss_batch: cluster c of the grid reduces plane c of the table (same fixed-order sum per plane)
__global__ void __launch_bounds__(1024) tv_value_batch_kernel(const clift_tv_job* __restrict__ jobs) {
    const clift_tv_job j = jobs[blockIdx.x / kTvCluster];
    if (j.loss) tv_value_body(j.plane_hwc, j.comps, j.h, j.w, j.loss);
}

__device__ __forceinline__ void tv_value_body(const float* __restrict__ x, int C, int H, int W, float* __restrict__ out) {
    __shared__ float s_red[33];
    __shared__ float s_part[2];
    const int64_t n = (int64_t)H * W * C;
    const unsigned rank = cluster_ctarank(), csize = cluster_nctarank();
    float th = 0.0f, tw = 0.0f;
    for (int64_t i = (int64_t)rank * blockDim.x + threadIdx.x; i < n; i += (int64_t)csize * blockDim.x) {
        const int64_t pix = i / C;
        const int h = (int)(pix / W), w = (int)(pix - (int64_t)h * W);
        const float v = x[i];
        if (h + 1 < H) {
            const float dd = x[i + (int64_t)W * C] - v;
            th += dd * dd;
        }
        if (w + 1 < W) {
            const float dd = x[i + C] - v;
            tw += dd * dd;
        }
    }
    const float sh = block_sum(th, s_red);
    const float sw = block_sum(tw, s_red);
    if (threadIdx.x == 0) {
        s_part[0] = sh;
        s_part[1] = sw;
    }
    cluster_sync_all();
    if (rank == 0 && threadIdx.x == 0) {
        float tot_h = 0.0f, tot_w = 0.0f;
        for (unsigned r = 0; r < csize; ++r) {
            tot_h += ld_dsmem_f32(&s_part[0], r);
            tot_w += ld_dsmem_f32(&s_part[1], r);
        }
        const float cnt_h = (float)((double)C * (H - 1) * W + 1e-4), cnt_w = (float)((double)C * H * (W - 1) + 1e-4);
        out[0] = 2.0f * (tot_h / cnt_h + tot_w / cnt_w);
    }
    cluster_sync_all();      // peers' shared memory stays alive until rank 0 has read it
}

__device__ __forceinline__ void tv_grad_body(const float* __restrict__ x, int C, int H, int W, float* __restrict__ g, float scale,
                                             int64_t i);

__global__ void __launch_bounds__(256) tv_grad_kernel(const float* __restrict__ x, int C, int H, int W, float* __restrict__ g,
                                                      float scale) {
    tv_grad_body(x, C, H, W, g, scale, (int64_t)blockIdx.x * blockDim.x + threadIdx.x);
}

// clift_tv_loss_batch: blockIdx.y = plane of the table, grid-stride over its elements
__global__ void __launch_bounds__(256) tv_grad_batch_kernel(const clift_tv_job* __restrict__ jobs) {
    const clift_tv_job j = jobs[blockIdx.y];
    if (!j.grad_hwc) return;
    const int64_t n = (int64_t)j.comps * j.h * j.w;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        tv_grad_body(j.plane_hwc, j.comps, j.h, j.w, j.grad_hwc, j.grad_scale, i);
}

__device__ __forceinline__ void tv_grad_body(const float* __restrict__ x, int C, int H, int W, float* __restrict__ g, float scale,
                                             int64_t i) {
    if (i >= (int64_t)H * W * C) return;
    const int64_t pix = i / C;
    const int h = (int)(pix / W), w = (int)(pix - (int64_t)h * W);
    const float cnt_h = (float)((double)C * (H - 1) * W + 1e-4), cnt_w = (float)((double)C * H * (W - 1) + 1e-4);
    const float v = x[i];
    float gh = 0.0f, gw = 0.0f;
    if (h > 0) gh += v - x[i - (int64_t)W * C];
    if (h + 1 < H) gh -= x[i + (int64_t)W * C] - v;
    if (w > 0) gw += v - x[i - C];
    if (w + 1 < W) gw -= x[i + C] - v;
    g[i] += scale * 4.0f * (gh / cnt_h + gw / cnt_w);
}

}  // namespace
}  // namespace clift

using namespace clift;

extern "C" int32_t clift_slowfast_loss(const float* features, const int64_t* labels, const float* confidences, int32_t n,
                                       int32_t d, float* loss, float* grad_features, void* stream) {
    CLIFT_CHECK_ARG(features && labels && confidences && loss && n >= 0 && d > 0, "null pointer or bad size");
    if (n == 0) {
        CLIFT_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), (cudaStream_t)stream));
        return CLIFT_OK;
    }
    const int nf = n / 2, ns = n - nf;
    const size_t smem = (size_t)n * 8 + (size_t)n * 2 * d * 4 + (size_t)ns * d * 4 + (size_t)nf * 4 * 4 + 64;
    CLIFT_CHECK_SUPPORTED(smem <= 200 * 1024, "slow-fast loss: N*d too large for one CTA's shared memory");
    CLIFT_CUDA(cudaFuncSetAttribute(slowfast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CLIFT_CUDA(cudaFuncSetAttribute(slowfast_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    const int cluster = nf >= 512 ? 16 : (nf >= 64 ? 8 : 1);      // one warp per fast sample: 32 warps per CTA
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)cluster);
    lc.blockDim = dim3(kLossThreads);
    lc.dynamicSmemBytes = smem;
    lc.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    lc.attrs = attr;
    lc.numAttrs = 1;
    CLIFT_CUDA(cudaLaunchKernelEx(&lc, slowfast_kernel, features, (const long long*)labels, confidences, (int)n, (int)d, loss,
                                  grad_features));
    CLIFT_AFTER_LAUNCH("slowfast_kernel");
    return CLIFT_OK;
}

extern "C" int32_t clift_contrastive_loss(const float* features, const int64_t* labels, int32_t n, int32_t dim,
                                          float temperature, float* loss, float* grad_features, void* stream) {
    CLIFT_CHECK_ARG(features && labels && loss && n > 0 && dim > 0, "null pointer or bad size");
    const size_t smem = (size_t)n * 8 + (size_t)n * dim * 4 + (size_t)n * 8 + 64;
    CLIFT_CHECK_SUPPORTED(smem <= 200 * 1024, "contrastive loss: N*D too large for one CTA's shared memory");
    CLIFT_CUDA(cudaFuncSetAttribute(contrastive_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CLIFT_CUDA(cudaFuncSetAttribute(contrastive_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    const int cluster = n >= 512 ? 16 : (n >= 64 ? 8 : 1);        // one warp per sample: 32 warps per CTA
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)cluster);
    lc.blockDim = dim3(kLossThreads);
    lc.dynamicSmemBytes = smem;
    lc.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    lc.attrs = attr;
    lc.numAttrs = 1;
    CLIFT_CUDA(cudaLaunchKernelEx(&lc, contrastive_kernel, features, (const long long*)labels, (int)n, (int)dim, temperature, loss,
                                  grad_features));
    CLIFT_AFTER_LAUNCH("contrastive_kernel");
    return CLIFT_OK;
}

extern "C" int32_t clift_ema_update(float* slow, const float* fast, int64_t n, double momentum, void* stream) {
    CLIFT_CHECK_ARG(slow && fast && n >= 0, "null pointer or bad size");
    if (n == 0) return CLIFT_OK;
    // the reference scales by the Python double (1 - momentum) rounded to fp32 when it meets the tensor
    const float om = (float)(1.0 - momentum);
    ema_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(slow, fast, n, (float)momentum, om);
    CLIFT_AFTER_LAUNCH("ema_kernel");
    return CLIFT_OK;
}

extern "C" int32_t clift_ema_update_batch(const clift_ema_pair* table_dev, int32_t n_pairs, int64_t max_n, double momentum,
                                          void* stream) {
    CLIFT_CHECK_ARG(n_pairs >= 0 && max_n >= 0, "negative size");
    if (n_pairs == 0 || max_n == 0) return CLIFT_OK;
    CLIFT_CHECK_ARG(table_dev != nullptr, "null table");
    CLIFT_CHECK_SUPPORTED(n_pairs <= 65535, "more than 65535 pairs in one call");
    const float om = (float)(1.0 - momentum);       // as clift_ema_update
    const unsigned gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(max_n, 256), 4 * sm_count()));
    ema_batch_kernel<<<dim3(gx, (unsigned)n_pairs), 256, 0, (cudaStream_t)stream>>>(table_dev, (float)momentum, om);
    CLIFT_AFTER_LAUNCH("ema_batch_kernel");
    return CLIFT_OK;
}

extern "C" int32_t clift_tv_loss_batch(const clift_tv_job* jobs_dev, int32_t n_jobs, int64_t max_n, void* stream) {
    CLIFT_CHECK_ARG(n_jobs >= 0 && max_n >= 0, "negative size");
    if (n_jobs == 0 || max_n == 0) return CLIFT_OK;
    CLIFT_CHECK_ARG(jobs_dev != nullptr, "null table");
    CLIFT_CHECK_SUPPORTED(n_jobs <= 4096, "more than 4096 planes in one call");
    {
        CLIFT_CUDA(cudaFuncSetAttribute(tv_value_batch_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3((unsigned)(kTvCluster * n_jobs));
        lc.blockDim = dim3(1024);
        lc.dynamicSmemBytes = 0;
        lc.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = kTvCluster;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        lc.attrs = attr;
        lc.numAttrs = 1;
        CLIFT_CUDA(cudaLaunchKernelEx(&lc, tv_value_batch_kernel, jobs_dev));
        CLIFT_AFTER_LAUNCH("tv_value_batch_kernel");
    }
    const unsigned gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(max_n, 256), 4 * sm_count()));
    tv_grad_batch_kernel<<<dim3(gx, (unsigned)n_jobs), 256, 0, (cudaStream_t)stream>>>(jobs_dev);
    CLIFT_AFTER_LAUNCH("tv_grad_batch_kernel");
    return CLIFT_OK;
}

extern "C" int32_t clift_tv_loss(const float* plane_hwc, int32_t comps, int32_t h, int32_t w, float* loss, float* grad_hwc,
                                 float grad_scale, void* stream) {
    CLIFT_CHECK_ARG(plane_hwc && comps > 0 && h > 0 && w > 0, "null pointer or bad size");
    if (loss) {
        // 16 CTAs per cluster is above the portable 8: opt in (per device, so on every call - it is a host-side table write)
        CLIFT_CUDA(cudaFuncSetAttribute(tv_value_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3(kTvCluster);
        lc.blockDim = dim3(1024);
        lc.dynamicSmemBytes = 0;
        lc.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = kTvCluster;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        lc.attrs = attr;
        lc.numAttrs = 1;
        CLIFT_CUDA(cudaLaunchKernelEx(&lc, tv_value_kernel, plane_hwc, (int)comps, (int)h, (int)w, loss));
        CLIFT_AFTER_LAUNCH("tv_value_kernel");
    }
    if (grad_hwc) {
        const int64_t n = (int64_t)comps * h * w;
        tv_grad_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(plane_hwc, comps, h, w, grad_hwc, grad_scale);
        CLIFT_AFTER_LAUNCH("tv_grad_kernel");
    }
    return CLIFT_OK;
}
