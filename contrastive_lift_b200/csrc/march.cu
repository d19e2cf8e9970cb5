// Ray march: sampling (bit-exact), density lookup, transmittance scan, per-ray reductions,
// active-sample compaction.  One warp owns one ray; samples are processed 32 at a time.
//
// Reference: model/renderer/panopli_tensoRF_renderer.py:80-103,137,174,626-631,800-817 and
// model/radiance_field/tensoRF.py:108-125.  HBM/L2-gather bound: per in-box sample 18 taps x
// 16 channels x 4 B = 1152 algorithmic bytes.  Layout choices:
//   * planes are channel-last, so one bilinear tap of one quad lane is a single 16 B load and a
//     quad reads the 64 B texel contiguously (8 samples per warp instruction);
//   * the three density line factors (<= 37 KB at 192^3) are staged once per CTA into shared
//     memory with 1-D bulk TMA (cp.async.bulk + mbarrier) - 6 of the 18 taps never touch L1/L2;
//   * transmittance is a warp-level inclusive product scan with a running carry (C1), the
//     distortion loss (C2) two warp prefix sums; nothing [B,S]-shaped is written except the
//     compositing weights the compaction pass re-reads.
#include "launchers.h"
#include "march.cuh"

namespace clift {

namespace {

constexpr int kWarpsPerCta = 8;
constexpr int kFullLaneMin = 12;   // in-box samples per 32-sample chunk from which the lane-per-sample lookup wins

template <int NV, bool kFused>
__global__ void __launch_bounds__(kWarpsPerCta * 32) march_kernel(const __grid_constant__ MarchParams P) {
    extern __shared__ __align__(16) float s_lines[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ long long s_group;                 // fused: the claimed group / its exclusive record offset
    __shared__ int s_cnt[kWarpsPerCta + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* lines_s = nullptr;
    if (P.lines_in_smem) {
        if (threadIdx.x == 0) {
            mbar_init(&s_bar, 1);
            mbar_fence_init();
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t total = 0;
            for (int m = 0; m < 3; ++m) total += (uint32_t)P.f.ll[m] * P.f.comps * 4u;
            mbar_expect_tx(&s_bar, total);
            uint32_t off = 0;
            for (int m = 0; m < 3; ++m) {
                const uint32_t bytes = (uint32_t)P.f.ll[m] * P.f.comps * 4u;
                for (uint32_t done = 0; done < bytes; done += 16384u) {
                    const uint32_t n = min(16384u, bytes - done);
                    bulk_g2s((char*)s_lines + off + done, (const char*)P.f.line[m] + done, n, &s_bar);
                }
                off += bytes;
            }
        }
        mbar_wait(&s_bar, 0);
        lines_s = s_lines;
    }

    const GeomParams& G = P.g;
    const int S = G.S;
    const int n_chunks = (S + 31) >> 5;
    const int q = lane & 3;
    unsigned long long inbox_total = 0;
    // fused: this warp's weights of the running ray, [round32(S)] floats behind the staged line factors
    size_t line_floats = 0;
    if (P.lines_in_smem)
        for (int m = 0; m < 3; ++m) line_floats += (size_t)P.f.ll[m] * P.f.comps;
    float* wst = s_lines + line_floats + (size_t)warp * (size_t)(n_chunks * 32);
    const long long n_groups = (P.n_rays + kWarpsPerCta - 1) / kWarpsPerCta;

    for (int64_t it = (int64_t)blockIdx.x;; it += (int64_t)gridDim.x) {
        int64_t ray;
        long long group = it;
        if (kFused) {       // groups in ticket order: every predecessor of a claimed group is running or done
            __syncthreads();
            if (threadIdx.x == 0) s_group = (long long)atomicAdd(P.stats + 4, 1ull);
            __syncthreads();
            group = s_group;
        }
        if (group >= n_groups) break;
        ray = group * kWarpsPerCta + warp;
        const bool live = ray < P.n_rays;             // fused: the last group may be ragged; its idle warps still take part
        if (!kFused && !live) continue;
        if (!live) ray = P.n_rays - 1;
        const float rv = lane < 8 ? __ldg(P.rays + ray * 8 + lane) : 0.0f;
        RayGeom g;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            g.o[k] = __shfl_sync(0xffffffffu, rv, k);
            g.d[k] = __shfl_sync(0xffffffffu, rv, 3 + k);
        }
        const float near = __shfl_sync(0xffffffffu, rv, 6), far = __shfl_sync(0xffffffffu, rv, 7);
        g.t_min = ray_t_min(g.o, g.d, near, far, G.amin, G.amax);
        g.has_jit = P.jitter != nullptr;
        g.jit = g.has_jit ? __ldg(P.jitter + ray) : 0.0f;

        float T_run = 1.0f, W_run = 0.0f, WM_run = 0.0f;
        float opa = 0.0f, dep = 0.0f, uni = 0.0f, bi = 0.0f;
        int n_act = 0, n_in = 0;
        float* wrow = kFused ? nullptr : P.w_dense + ray * S;

        for (int c = 0; c < (live ? n_chunks : 0); ++c) {
            const int i = c * 32 + lane;
            const bool valid = i < S;
            const float t = sample_t(g, G.step, i);
            float x[3];
            const bool in = sample_point(g, t, G.amin, G.amax, G.inv, x) && valid;
            const unsigned inm = __ballot_sync(0xffffffffu, in);
            float w = 0.0f, sigma_out = 0.0f, trans_out = T_run;
            if (inm != 0u) {
                n_in += __popc(inm);
                float feat = 0.0f;
                if (__popc(inm) >= kFullLaneMin) {
                    // mostly in-box chunk: every lane evaluates its own sample
                    if (in) feat = vm_dot_full<NV>(P.f, lines_s, x[0], x[1], x[2]);
                } else {
                    // sparse chunk: quads (4 lanes x 4 channels) walk only the 8-sample groups that have work
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (((inm >> (8 * j)) & 0xffu) == 0u) continue;   // warp-uniform
                        const int src = j * 8 + (lane >> 2);
                        const float sx = __shfl_sync(0xffffffffu, x[0], src);
                        const float sy = __shfl_sync(0xffffffffu, x[1], src);
                        const float sz = __shfl_sync(0xffffffffu, x[2], src);
                        float acc = 0.0f;
                        if ((inm >> src) & 1u) acc = vm_dot_partial<NV>(P.f, lines_s, sx, sy, sz, q);
                        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                        const float v = __shfl_sync(0xffffffffu, acc, (lane & 7) * 4);
                        if ((lane >> 3) == j) feat = v;
                    }
                }
                const float sigma = in ? softplus_f(feat + P.shift) : 0.0f;
                const float tn = sample_t(g, G.step, i + 1);
                const float delta = (i < S - 1) ? __fsub_rn(tn, t) : 0.0f;
                const float alpha = __fsub_rn(1.0f, expf(__fmul_rn(-sigma, __fmul_rn(delta, G.scale))));
                const float fac = __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f);
                // inclusive product scan -> exclusive transmittance
                float pr = fac;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const float v = __shfl_up_sync(0xffffffffu, pr, o);
                    if (lane >= o) pr *= v;
                }
                float ex = __shfl_up_sync(0xffffffffu, pr, 1);
                if (lane == 0) ex = 1.0f;
                trans_out = T_run * ex;
                sigma_out = sigma;
                w = alpha * trans_out;
                T_run *= __shfl_sync(0xffffffffu, pr, 31);
                const float mid = (i < S - 1) ? __fmul_rn(__fadd_rn(tn, t), 0.5f) : sample_t(g, G.step, S - 2);
                // exclusive prefix sums of w and w*mid for the distortion loss
                const float wm = w * mid;
                float sw = w, swm = wm;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const float a = __shfl_up_sync(0xffffffffu, sw, o);
                    const float b = __shfl_up_sync(0xffffffffu, swm, o);
                    if (lane >= o) {
                        sw += a;
                        swm += b;
                    }
                }
                const float w_ex = W_run + (sw - w), wm_ex = WM_run + (swm - wm);
                bi += wm * w_ex - w * wm_ex;
                uni += w * w * delta;
                W_run += __shfl_sync(0xffffffffu, sw, 31);
                WM_run += __shfl_sync(0xffffffffu, swm, 31);
                opa += w;
                dep += w * t;
                n_act += __popc(__ballot_sync(0xffffffffu, w > G.thres));
            }
            if (kFused) {
                wst[i] = w;
                if (valid && P.sigma_dense) {       // training: the march backward reads sigma and T of every in-box sample
                    P.sigma_dense[ray * S + i] = sigma_out;
                    P.trans_dense[ray * S + i] = trans_out;
                }
            } else if (valid) {
                wrow[i] = w;
                if (P.sigma_dense) {
                    P.sigma_dense[ray * S + i] = sigma_out;
                    P.trans_dense[ray * S + i] = trans_out;
                }
            }
        }
        opa = warp_sum(opa);
        dep = warp_sum(dep);
        if (P.dist_ray) {
            uni = warp_sum(uni);
            bi = warp_sum(bi);
        }
        if (lane == 0 && live) {
            if (!kFused) P.count[ray] = n_act;
            P.opacity[ray] = opa;
            P.depth[ray] = dep;
            if (P.dist_ray) P.dist_ray[ray] = uni * (1.0f / 3.0f) + 2.0f * bi;
        }
        if (P.points && lane < 3 && live) P.points[ray * 3 + lane] = __fadd_rn(g.o[lane], __fmul_rn(dep, g.d[lane]));
        inbox_total += (unsigned long long)n_in;
        if (kFused) {
            // ---- record offsets: intra-group exclusive scan + decoupled look-back over the groups before this one ----
            if (lane == 0) s_cnt[warp] = live ? n_act : 0;
            __syncthreads();
            if (warp == 0) {
                const int c = lane < kWarpsPerCta ? s_cnt[lane] : 0;
                int inc = c;
#pragma unroll
                for (int o = 1; o < kWarpsPerCta; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += t;
                }
                const unsigned long long total = (unsigned long long)__shfl_sync(0xffffffffu, inc, kWarpsPerCta - 1);
                constexpr unsigned long long kAgg = 1ull << 62, kInc = 2ull << 62, kVal = (1ull << 62) - 1ull;
                volatile unsigned long long* st = P.scan_state;
                if (lane == 0) atomicExch(P.scan_state + group, (group == 0 ? kInc : kAgg) | total);
                unsigned long long excl = 0;
                long long j = group - 1;
                while (j >= 0) {        // 32 predecessors per step, nearest first in lane 0
                    const long long jj = j - lane;
                    unsigned long long v = 0;
                    do {
                        v = jj >= 0 ? st[jj] : kInc;                       // before the first group: an inclusive prefix of 0
                    } while (__any_sync(0xffffffffu, (v >> 62) == 0ull));   // wait until every read state is published
                    const unsigned inc_mask = __ballot_sync(0xffffffffu, (v >> 62) == 2ull);
                    const int stop = inc_mask ? __ffs(inc_mask) - 1 : 32;   // nearest lane holding an inclusive prefix
                    unsigned long long part = lane <= stop ? (v & kVal) : 0ull;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                    excl += part;
                    if (inc_mask) break;
                    j -= 32;
                }
                if (lane == 0) {
                    if (group > 0) atomicExch(P.scan_state + group, kInc | (excl + total));
                    s_group = (long long)excl;
                    if (group == n_groups - 1) {     // the last group knows the total
                        const unsigned long long all = excl + total;
                        P.stats[0] = all;
                        P.stats[2] = (long long)all > P.cap ? 1ull : 0ull;
                        P.stats[3] = (unsigned long long)((min((long long)all, P.cap) + CLIFT_TILE - 1) / CLIFT_TILE);
                    }
                }
                if (lane < kWarpsPerCta) s_cnt[lane] = inc - c;            // exclusive offset of each ray inside the group
            }
            __syncthreads();
            // ---- emit this ray's records (same arithmetic as fill_kernel: positions are recomputed, bit-exact) ----
            if (live && n_act > 0) {
                const long long pos = s_group + s_cnt[warp];
                int found = 0;
                for (int c = 0; c < n_chunks && found < n_act; ++c) {
                    const int i = c * 32 + lane;
                    const float w = wst[i];
                    const bool act = i < S && w > G.thres;
                    const unsigned m = __ballot_sync(0xffffffffu, act);
                    if (act) {
                        const long long dst = pos + found + __popc(m & ((1u << lane) - 1u));
                        if (dst < P.cap) {
                            float x[3];
                            sample_point(g, sample_t(g, G.step, i), G.amin, G.amax, G.inv, x);
                            P.rec_pos[dst] = make_float4(x[0], x[1], x[2], w);
                            P.rec_ray[dst] = (int32_t)ray;
                            P.rec_idx[dst] = i;
                        }
                    }
                    found += __popc(m);
                }
            }
        }
    }
    if (lane == 0 && inbox_total) atomicAdd(P.stats + 1, inbox_total);
}

// ---------------------------------------------------------------------------------------
// exclusive scan of per-ray active counts (three tiny kernels, deterministic)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int block_exclusive_scan_256(int v, int* s_warp, int& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int ws = lane < 8 ? s_warp[lane] : 0;
        int winc = ws;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < 8) s_warp[lane] = winc - ws;
        if (lane == 7) s_warp[8] = winc;
    }
    __syncthreads();
    total = s_warp[8];
    const int r = s_warp[warp] + inc - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(256) scan_reduce_kernel(const int32_t* __restrict__ count, int32_t* __restrict__ bsum,
                                                          int64_t n) {
    __shared__ int s_warp[9];
    const int64_t base = (int64_t)blockIdx.x * kScanBlock;
    int v = 0;
    for (int k = 0; k < kScanBlock / 256; ++k) {
        const int64_t i = base + k * 256 + threadIdx.x;
        if (i < n) v += count[i];
    }
    int total;
    block_exclusive_scan_256(v, s_warp, total);
    if (threadIdx.x == 0) bsum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(256) scan_blocks_kernel(int32_t* __restrict__ bsum, int nblk, int32_t* __restrict__ offset,
                                                          int64_t n, unsigned long long* stats, long long cap) {
    __shared__ int s_warp[9];
    int carry = 0;
    for (int b0 = 0; b0 < nblk; b0 += 256) {
        const int b = b0 + threadIdx.x;
        const int v = b < nblk ? bsum[b] : 0;
        int total;
        const int ex = block_exclusive_scan_256(v, s_warp, total);
        if (b < nblk) bsum[b] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) {
        offset[n] = carry;
        stats[0] = (unsigned long long)carry;
        stats[2] = (long long)carry > cap ? 1ull : 0ull;
        stats[3] = (unsigned long long)((min((long long)carry, cap) + CLIFT_TILE - 1) / CLIFT_TILE);
    }
}

__global__ void __launch_bounds__(256) scan_apply_kernel(const int32_t* __restrict__ count, const int32_t* __restrict__ bsum,
                                                         int32_t* __restrict__ offset, int64_t n) {
    __shared__ int s_warp[9];
    const int64_t base = (int64_t)blockIdx.x * kScanBlock + (int64_t)threadIdx.x * (kScanBlock / 256);
    int v[kScanBlock / 256];
    int sum = 0;
#pragma unroll
    for (int k = 0; k < kScanBlock / 256; ++k) {
        v[k] = (base + k < n) ? count[base + k] : 0;
        sum += v[k];
    }
    int total;
    int ex = block_exclusive_scan_256(sum, s_warp, total) + bsum[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanBlock / 256; ++k) {
        if (base + k < n) offset[base + k] = ex;
        ex += v[k];
    }
}

// ---------------------------------------------------------------------------------------
// compaction: (x,y,z,w), ray, sample index of every active sample, ray-major, sample-minor
// ---------------------------------------------------------------------------------------
struct FillParams {
    GeomParams g;
    const float* rays;
    const float* jitter;
    int64_t n_rays;
    const float* w_dense;
    const int32_t* count;
    const int32_t* offset;
    float4* rec_pos;
    int32_t* rec_ray;
    int32_t* rec_idx;
    long long cap;
};

__global__ void __launch_bounds__(256) fill_kernel(const __grid_constant__ FillParams P) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const GeomParams& G = P.g;
    const int S = G.S;
    for (int64_t ray = (int64_t)blockIdx.x * 8 + warp; ray < P.n_rays; ray += (int64_t)gridDim.x * 8) {
        const int cnt = P.count[ray];
        if (cnt == 0) continue;
        long long pos = P.offset[ray];
        const float rv = lane < 8 ? __ldg(P.rays + ray * 8 + lane) : 0.0f;
        RayGeom g;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            g.o[k] = __shfl_sync(0xffffffffu, rv, k);
            g.d[k] = __shfl_sync(0xffffffffu, rv, 3 + k);
        }
        g.t_min = ray_t_min(g.o, g.d, __shfl_sync(0xffffffffu, rv, 6), __shfl_sync(0xffffffffu, rv, 7), G.amin, G.amax);
        g.has_jit = P.jitter != nullptr;
        g.jit = g.has_jit ? __ldg(P.jitter + ray) : 0.0f;
        const float* wrow = P.w_dense + ray * S;
        int found = 0;
        for (int c = 0; c * 32 < S && found < cnt; ++c) {
            const int i = c * 32 + lane;
            const float w = i < S ? wrow[i] : 0.0f;
            const bool act = w > G.thres;
            const unsigned m = __ballot_sync(0xffffffffu, act);
            if (act) {
                const long long dst = pos + found + __popc(m & ((1u << lane) - 1u));
                if (dst < P.cap) {
                    float x[3];
                    sample_point(g, sample_t(g, G.step, i), G.amin, G.amax, G.inv, x);
                    P.rec_pos[dst] = make_float4(x[0], x[1], x[2], w);
                    P.rec_ray[dst] = (int32_t)ray;
                    P.rec_idx[dst] = i;
                }
            }
            found += __popc(m);
        }
    }
}

// ---------------------------------------------------------------------------------------
// per-ray epilogue (renderer:160-167) and the ordered distortion-loss mean
// ---------------------------------------------------------------------------------------
struct FinishParams {
    int64_t n_rays;
    int n_cls;
    int softmax;
    int add_bg;
    const float* opacity;
    const float* rgb_raw;
    const float* sem_raw;
    float* rgb;
    float* sem;
};

__global__ void __launch_bounds__(256) finish_kernel(const __grid_constant__ FinishParams P) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P.n_rays) return;
    if (P.rgb) {
        const float bg = P.add_bg ? (1.0f - P.opacity[r]) : 0.0f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float v = P.rgb_raw[r * 3 + k];
            if (P.add_bg) v = v + bg;
            P.rgb[r * 3 + k] = fminf(fmaxf(v, 0.0f), 1.0f);
        }
    }
    if (P.sem) {
        const float* s = P.sem_raw + r * P.n_cls;
        float* o = P.sem + r * P.n_cls;
        if (P.softmax) {
            float tot = 0.0f;
            for (int c = 0; c < P.n_cls; ++c) tot += s[c];
            tot += 1e-8f;
            for (int c = 0; c < P.n_cls; ++c) o[c] = logf(s[c] / tot + 1e-8f);
        } else {
            for (int c = 0; c < P.n_cls; ++c) o[c] = s[c];
        }
    }
}

__global__ void __launch_bounds__(1024) mean_kernel(const float* __restrict__ v, int64_t n, float* __restrict__ out) {
    __shared__ float s[32];
    float acc = 0.0f;
    for (int64_t i = threadIdx.x; i < n; i += 1024) acc += v[i];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = s[threadIdx.x];
        t = warp_sum(t);
        if (threadIdx.x == 0) out[0] = t / (float)n;
    }
}

// ---------------------------------------------------------------------------------------
// parity/debug entries: materialised sampling (S1,S3) and point-wise density (F1)
// ---------------------------------------------------------------------------------------
struct SampleParams {
    GeomParams g;
    const float* rays;
    const float* jitter;
    int64_t n_rays;
    float* z;
    float* xyz;
    uint8_t* inbox;
};

__global__ void __launch_bounds__(256) sample_points_kernel(const __grid_constant__ SampleParams P) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int S = P.g.S;
    if (idx >= P.n_rays * S) return;
    const int64_t ray = idx / S;
    const int i = (int)(idx - ray * S);
    RayGeom g;
    for (int k = 0; k < 3; ++k) {
        g.o[k] = P.rays[ray * 8 + k];
        g.d[k] = P.rays[ray * 8 + 3 + k];
    }
    g.t_min = ray_t_min(g.o, g.d, P.rays[ray * 8 + 6], P.rays[ray * 8 + 7], P.g.amin, P.g.amax);
    g.has_jit = P.jitter != nullptr;
    g.jit = g.has_jit ? P.jitter[ray] : 0.0f;
    const float t = sample_t(g, P.g.step, i);
    float x[3];
    const bool in = sample_point(g, t, P.g.amin, P.g.amax, P.g.inv, x);
    if (P.z) P.z[idx] = t;
    if (P.xyz) {
        P.xyz[idx * 3 + 0] = x[0];
        P.xyz[idx * 3 + 1] = x[1];
        P.xyz[idx * 3 + 2] = x[2];
    }
    if (P.inbox) P.inbox[idx] = in ? 1 : 0;
}

template <int NV>
__global__ void __launch_bounds__(256) density_kernel(const __grid_constant__ FactorParams F, float shift,
                                                      const float* __restrict__ xyz, int64_t n, float* __restrict__ sigma) {
    const int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const int q = threadIdx.x & 3;
    float acc = 0.0f;
    if (s < n) acc = vm_dot_partial<NV>(F, nullptr, xyz[s * 3], xyz[s * 3 + 1], xyz[s * 3 + 2], q);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (s < n && q == 0) sigma[s] = softplus_f(acc + shift);
}


// ---------------------------------------------------------------------------------------
// Backward of the march (renderer:97-101,137,164 through raw_to_alpha, softplus and the VM lookup).
// One warp per ray, samples walked from the far end in chunks of 32 so that the suffix sums
//   S_i = sum_{j>i} w_j g_j   (compositing)      W_{>i}, WM_{>i}   (distortion loss)
// are running carries.  sigma_i and T_i come from the forward's dense buffers (bit-identical alpha, w).
//   g_i      = dL/dw_i = g_opacity + g_w_rgb[i] + gD/B * (2/3 w_i d_i + 2 (m_i W_<i - WM_<i + WM_>i - m_i W_>i))
//   dL/da_i  = T_i g_i - S_i / (1 - a_i + 1e-10)
//   dL/dsig  = dL/da_i * d_i * scale * (1 - a_i),   dL/dfeat = dL/dsig * (1 - exp(-sigma_i))   [softplus' = sigmoid]
// and dL/dfeat is scattered to the 18 taps x comps of the density factors with 16-byte vector reductions.
// ---------------------------------------------------------------------------------------
struct MarchBwdParams {
    GeomParams g;
    FactorParams f;
    const float* rays;
    const float* jitter;
    int64_t n_rays;
    const float* sigma_dense;
    const float* trans_dense;
    const float* g_w;        // [B,S] or null
    const float* g_ray;      // [B][stride], g_opacity at [stride-1]
    int ray_stride;
    const float* g_dist;     // device scalar or null
    float* g_plane[3];
    float* g_line[3];
};

__device__ __forceinline__ void red_add4f(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ float suffix_inclusive(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_down_sync(0xffffffffu, v, o);
        if (lane + o < 32) v += t;
    }
    return v;
}

template <int NV>
__global__ void __launch_bounds__(kWarpsPerCta * 32) march_backward_kernel(const __grid_constant__ MarchBwdParams P) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const GeomParams& G = P.g;
    const FactorParams& F = P.f;
    const int S = G.S;
    const int n_chunks = (S + 31) >> 5;
    const int q = lane & 3;
    const float gD = P.g_dist ? __ldg(P.g_dist) / (float)P.n_rays : 0.0f;

    for (int64_t ray = (int64_t)blockIdx.x * kWarpsPerCta + warp; ray < P.n_rays; ray += (int64_t)gridDim.x * kWarpsPerCta) {
        const float rv = lane < 8 ? __ldg(P.rays + ray * 8 + lane) : 0.0f;
        RayGeom g;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            g.o[k] = __shfl_sync(0xffffffffu, rv, k);
            g.d[k] = __shfl_sync(0xffffffffu, rv, 3 + k);
        }
        g.t_min = ray_t_min(g.o, g.d, __shfl_sync(0xffffffffu, rv, 6), __shfl_sync(0xffffffffu, rv, 7), G.amin, G.amax);
        g.has_jit = P.jitter != nullptr;
        g.jit = g.has_jit ? __ldg(P.jitter + ray) : 0.0f;
        const float g_opa = __ldg(P.g_ray + ray * P.ray_stride + P.ray_stride - 1);
        const float* srow = P.sigma_dense + ray * S;
        const float* trow = P.trans_dense + ray * S;
        const float* gwrow = P.g_w ? P.g_w + ray * S : nullptr;

        float W_tot = 0.0f, WM_tot = 0.0f;
        if (gD != 0.0f) {
            for (int c = 0; c < n_chunks; ++c) {
                const int i = c * 32 + lane;
                if (i < S) {
                    const float t = sample_t(g, G.step, i), tn = sample_t(g, G.step, i + 1);
                    const float delta = (i < S - 1) ? __fsub_rn(tn, t) : 0.0f;
                    const float mid = (i < S - 1) ? __fmul_rn(__fadd_rn(tn, t), 0.5f) : sample_t(g, G.step, S - 2);
                    const float alpha = __fsub_rn(1.0f, expf(__fmul_rn(-srow[i], __fmul_rn(delta, G.scale))));
                    const float w = alpha * trow[i];
                    W_tot += w;
                    WM_tot += w * mid;
                }
            }
            W_tot = warp_sum(W_tot);
            WM_tot = warp_sum(WM_tot);
        }

        float S_run = 0.0f, Wgt_run = 0.0f, WMgt_run = 0.0f;
        for (int c = n_chunks - 1; c >= 0; --c) {
            const int i = c * 32 + lane;
            const bool valid = i < S;
            const float t = sample_t(g, G.step, i), tn = sample_t(g, G.step, i + 1);
            float x[3];
            const bool in = sample_point(g, t, G.amin, G.amax, G.inv, x) && valid;
            const float delta = (valid && i < S - 1) ? __fsub_rn(tn, t) : 0.0f;
            const float mid = (i < S - 1) ? __fmul_rn(__fadd_rn(tn, t), 0.5f) : sample_t(g, G.step, S - 2);
            const float sigma = valid ? srow[i] : 0.0f;
            const float T = valid ? trow[i] : 0.0f;
            const float alpha = __fsub_rn(1.0f, expf(__fmul_rn(-sigma, __fmul_rn(delta, G.scale))));
            const float w = alpha * T;
            const float wm = w * mid;
            // distortion-loss part of dL/dw
            const float sw = suffix_inclusive(w, lane), swm = suffix_inclusive(wm, lane);
            const float W_gt = Wgt_run + (sw - w), WM_gt = WMgt_run + (swm - wm);
            Wgt_run += __shfl_sync(0xffffffffu, sw, 0);
            WMgt_run += __shfl_sync(0xffffffffu, swm, 0);
            float gi = g_opa + ((gwrow && valid) ? gwrow[i] : 0.0f);
            if (gD != 0.0f) {
                const float W_lt = W_tot - W_gt - w, WM_lt = WM_tot - WM_gt - wm;
                gi += gD * ((2.0f / 3.0f) * w * delta + 2.0f * (mid * W_lt - WM_lt + WM_gt - mid * W_gt));
            }
            if (!valid) gi = 0.0f;
            const float wg = w * gi;
            const float swg = suffix_inclusive(wg, lane);
            const float S_i = S_run + (swg - wg);
            S_run += __shfl_sync(0xffffffffu, swg, 0);
            const float fac = __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f);
            const float d_alpha = T * gi - S_i / fac;
            const float d_sigma = d_alpha * __fmul_rn(delta, G.scale) * (1.0f - alpha);
            const float dsp = sigma > 20.0f ? 1.0f : -expm1f(-sigma);
            const float dfeat = in ? d_sigma * dsp : 0.0f;
            const unsigned nz = __ballot_sync(0xffffffffu, dfeat != 0.0f);
            if (nz == 0u) continue;
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
                if (((nz >> (8 * j)) & 0xffu) == 0u) continue;   // warp-uniform
                const int src = j * 8 + (lane >> 2);
                const float sx = __shfl_sync(0xffffffffu, x[0], src);
                const float sy = __shfl_sync(0xffffffffu, x[1], src);
                const float sz = __shfl_sync(0xffffffffu, x[2], src);
                const float df = __shfl_sync(0xffffffffu, dfeat, src);
                if (df == 0.0f) continue;
                const float xs[3] = {sx, sy, sz};
#pragma unroll
                for (int m = 0; m < 3; ++m) {
                    const int W = F.pw[m];
                    const Tap2 t2 = make_tap2(xs[mode_a(m)], xs[mode_b(m)], W, F.ph[m]);
                    const Tap1 t1 = make_tap1(xs[mode_v(m)], F.ll[m]);
                    const int64_t row0 = (int64_t)t2.y0 * W, row1 = row0 + W;
#pragma unroll
                    for (int v = 0; v < NV; ++v) {
                        const int ch = v * 16 + q * 4;
                        const float4 pv = plane_tap(F.plane[m], t2, W, F.comps, ch);
                        const float4 lv = line_tap(F.line[m], t1, F.comps, ch);
                        const float a0 = df * lv.x, a1 = df * lv.y, a2 = df * lv.z, a3 = df * lv.w;
                        float* gp = P.g_plane[m];
                        if (t2.w00 != 0.0f) red_add4f(gp + (row0 + t2.x0) * F.comps + ch, a0 * t2.w00, a1 * t2.w00, a2 * t2.w00, a3 * t2.w00);
                        if (t2.w10 != 0.0f) red_add4f(gp + (row0 + t2.x0 + 1) * F.comps + ch, a0 * t2.w10, a1 * t2.w10, a2 * t2.w10, a3 * t2.w10);
                        if (t2.w01 != 0.0f) red_add4f(gp + (row1 + t2.x0) * F.comps + ch, a0 * t2.w01, a1 * t2.w01, a2 * t2.w01, a3 * t2.w01);
                        if (t2.w11 != 0.0f) red_add4f(gp + (row1 + t2.x0 + 1) * F.comps + ch, a0 * t2.w11, a1 * t2.w11, a2 * t2.w11, a3 * t2.w11);
                        const float b0 = df * pv.x, b1 = df * pv.y, b2 = df * pv.z, b3 = df * pv.w;
                        float* gl = P.g_line[m];
                        if (t1.w0 != 0.0f) red_add4f(gl + (int64_t)t1.z0 * F.comps + ch, b0 * t1.w0, b1 * t1.w0, b2 * t1.w0, b3 * t1.w0);
                        if (t1.w1 != 0.0f) red_add4f(gl + (int64_t)(t1.z0 + 1) * F.comps + ch, b0 * t1.w1, b1 * t1.w1, b2 * t1.w1, b3 * t1.w1);
                    }
                }
            }
        }
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------
int launch_march(const MarchParams& P_in, cudaStream_t stream) {
    MarchParams P = P_in;
    const int comps = P.f.comps;
    size_t line_bytes = 0;
    for (int m = 0; m < 3; ++m) line_bytes += (size_t)P.f.ll[m] * comps * 4;
    P.lines_in_smem = line_bytes <= 96 * 1024 ? 1 : 0;
    const bool fused = P.scan_state != nullptr;
    // fused compaction: + one row of round32(S) weights per warp behind the line factors
    const size_t stage_bytes = fused ? (size_t)kWarpsPerCta * (size_t)round_up(P.g.S, 32) * 4 : 0;
    if (fused && (P.lines_in_smem ? line_bytes : 0) + stage_bytes > 200 * 1024) {
        set_error("launch_march: fused compaction needs %zu bytes of shared memory", line_bytes + stage_bytes);
        return CLIFT_ERR_UNSUPPORTED;
    }
    const size_t smem = (P.lines_in_smem ? line_bytes : 0) + stage_bytes;
    const int64_t want = ceil_div(P.n_rays, kWarpsPerCta);
    const int grid = (int)std::min<int64_t>(want, (int64_t)sm_count() * 6);
    if (grid <= 0) return CLIFT_OK;
#define CLIFT_MARCH_CASE(NV)                                                                               \
    case NV: {                                                                                             \
        if (fused) {                                                                                       \
            CLIFT_CUDA(cudaFuncSetAttribute(march_kernel<NV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            march_kernel<NV, true><<<grid, kWarpsPerCta * 32, smem, stream>>>(P);                          \
        } else {                                                                                           \
            CLIFT_CUDA(cudaFuncSetAttribute(march_kernel<NV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            march_kernel<NV, false><<<grid, kWarpsPerCta * 32, smem, stream>>>(P);                         \
        }                                                                                                  \
        break;                                                                                             \
    }
    switch (comps / 16) {
        CLIFT_MARCH_CASE(1)
        CLIFT_MARCH_CASE(2)
        CLIFT_MARCH_CASE(3)
        default:
            set_error("launch_march: density_comps %d not in {16,32,48}", comps);
            return CLIFT_ERR_UNSUPPORTED;
    }
#undef CLIFT_MARCH_CASE
    CLIFT_AFTER_LAUNCH("march_kernel");
    return CLIFT_OK;
}

int launch_scan(const int32_t* count, int32_t* offset, int32_t* bsum, unsigned long long* stats, int64_t n, int64_t cap,
                cudaStream_t stream) {
    const int nblk = (int)ceil_div(n, kScanBlock);
    if (nblk <= 0) return CLIFT_OK;
    scan_reduce_kernel<<<nblk, 256, 0, stream>>>(count, bsum, n);
    CLIFT_AFTER_LAUNCH("scan_reduce_kernel");
    scan_blocks_kernel<<<1, 256, 0, stream>>>(bsum, nblk, offset, n, stats, (long long)cap);
    CLIFT_AFTER_LAUNCH("scan_blocks_kernel");
    scan_apply_kernel<<<nblk, 256, 0, stream>>>(count, bsum, offset, n);
    CLIFT_AFTER_LAUNCH("scan_apply_kernel");
    return CLIFT_OK;
}

int launch_fill(const clift_render_cfg* cfg, const float* rays, const float* jitter, int64_t n_rays, const Workspace& ws,
                int64_t cap, cudaStream_t stream) {
    FillParams P;
    P.g = make_geom(cfg);
    P.rays = rays;
    P.jitter = jitter;
    P.n_rays = n_rays;
    P.w_dense = ws.w_dense;
    P.count = ws.count;
    P.offset = ws.offset;
    P.rec_pos = ws.rec_pos;
    P.rec_ray = ws.rec_ray;
    P.rec_idx = ws.rec_idx;
    P.cap = cap;
    const int grid = (int)std::min<int64_t>(ceil_div(n_rays, 8), (int64_t)sm_count() * 8);
    if (grid <= 0) return CLIFT_OK;
    fill_kernel<<<grid, 256, 0, stream>>>(P);
    CLIFT_AFTER_LAUNCH("fill_kernel");
    return CLIFT_OK;
}

int launch_finish(int64_t n_rays, int n_cls, int softmax, int add_bg, const float* opacity, const float* rgb_raw,
                  const float* sem_raw, float* rgb, float* sem, const float* dist_ray, float* dist_reg, cudaStream_t stream) {
    if (n_rays <= 0) return CLIFT_OK;
    if (rgb || sem) {
        FinishParams P{n_rays, n_cls, softmax, add_bg, opacity, rgb_raw, sem_raw, rgb, sem};
        finish_kernel<<<(unsigned)ceil_div(n_rays, 256), 256, 0, stream>>>(P);
        CLIFT_AFTER_LAUNCH("finish_kernel");
    }
    if (dist_ray && dist_reg) {
        mean_kernel<<<1, 1024, 0, stream>>>(dist_ray, n_rays, dist_reg);
        CLIFT_AFTER_LAUNCH("mean_kernel");
    }
    return CLIFT_OK;
}

int launch_sample_points(const clift_render_cfg* cfg, const float* rays, const float* jitter, int64_t n_rays, float* z,
                         float* xyz, uint8_t* inbox, cudaStream_t stream) {
    SampleParams P{make_geom(cfg), rays, jitter, n_rays, z, xyz, inbox};
    const int64_t total = n_rays * cfg->n_samples;
    if (total <= 0) return CLIFT_OK;
    sample_points_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, stream>>>(P);
    CLIFT_AFTER_LAUNCH("sample_points_kernel");
    return CLIFT_OK;
}

int launch_density(const clift_field* field, const float* xyz, int64_t n, float* sigma, cudaStream_t stream) {
    if (n <= 0) return CLIFT_OK;
    FactorParams F = make_factors(field, false);
    const unsigned grid = (unsigned)ceil_div(n * 4, 256);
    switch (F.comps / 16) {
        case 1: density_kernel<1><<<grid, 256, 0, stream>>>(F, field->density_shift, xyz, n, sigma); break;
        case 2: density_kernel<2><<<grid, 256, 0, stream>>>(F, field->density_shift, xyz, n, sigma); break;
        case 3: density_kernel<3><<<grid, 256, 0, stream>>>(F, field->density_shift, xyz, n, sigma); break;
        default:
            set_error("launch_density: density_comps %d not in {16,32,48}", F.comps);
            return CLIFT_ERR_UNSUPPORTED;
    }
    CLIFT_AFTER_LAUNCH("density_kernel");
    return CLIFT_OK;
}

}  // namespace clift

namespace clift {

int launch_march_backward(const clift_render_cfg* cfg, const clift_field* field, const float* rays, const float* jitter,
                          int64_t n_rays, const Workspace& ws, int ray_stride, const float* g_dist, bool have_g_w,
                          const clift_field_grad* grad, cudaStream_t stream) {
    MarchBwdParams P;
    P.g = make_geom(cfg);
    P.f = make_factors(field, false);
    P.rays = rays;
    P.jitter = jitter;
    P.n_rays = n_rays;
    P.sigma_dense = ws.sigma_dense;
    P.trans_dense = ws.trans_dense;
    P.g_w = have_g_w ? ws.g_w : nullptr;
    P.g_ray = ws.g_ray;
    P.ray_stride = ray_stride;
    P.g_dist = g_dist;
    for (int m = 0; m < 3; ++m) {
        P.g_plane[m] = grad->density_plane[m];
        P.g_line[m] = grad->density_line[m];
    }
    const int grid = (int)std::min<int64_t>(ceil_div(n_rays, kWarpsPerCta), (int64_t)sm_count() * 6);
    if (grid <= 0) return CLIFT_OK;
    switch (P.f.comps / 16) {
        case 1: march_backward_kernel<1><<<grid, kWarpsPerCta * 32, 0, stream>>>(P); break;
        case 2: march_backward_kernel<2><<<grid, kWarpsPerCta * 32, 0, stream>>>(P); break;
        case 3: march_backward_kernel<3><<<grid, kWarpsPerCta * 32, 0, stream>>>(P); break;
        default:
            set_error("launch_march_backward: density_comps %d not in {16,32,48}", P.f.comps);
            return CLIFT_ERR_UNSUPPORTED;
    }
    CLIFT_AFTER_LAUNCH("march_backward_kernel");
    return CLIFT_OK;
}

}  // namespace clift
