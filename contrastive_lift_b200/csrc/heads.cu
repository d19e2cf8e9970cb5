// MLP heads on the compacted active samples (tensoRF.py:127-137, 383-418, 462-511, 565-594) and the
// weighted per-ray compositing of their outputs (renderer:103-131, 137-156).
//
// Work unit: a tile of CLIFT_TILE=128 consecutive active-sample records (ray-major, so a ray's samples are a
// contiguous run).  One CTA keeps the tile's activations resident in shared memory, K-major ([k][128]),
// and streams every layer's packed W^T through a double-buffered 16-row cp.async slab, so weights are
// read once per 128 samples and activations never leave the SM between layers:
//     semantic : xyz(+PE) -> 256 -> 256 -> 256 -> 256 -> C   (+softmax)
//     instance : xyz(+PE) -> 256 -> 256 -> 256 -> d           (fast, then slow)
//     rgb      : 18 taps x 48 ch appearance gather -> 144 -> basis 27 -> [feat,dir,PE] 150 -> 128 -> 128 -> 3
// Head outputs are multiplied by the compositing weight, summed per ray run inside the tile (fixed order)
// and added to the zero-initialised per-ray maps: a ray inside one tile is a plain 0+x, a ray split over
// two tiles is a commutative two-term sum, so results are run-to-run reproducible for <= 256 active
// samples per ray.
//
// This file is the fp32 CUDA-core (FFMA) implementation: exact fp32 products/accumulation like the
// reference's sgemm path (TF32 is off in the reference, trainer:33-35).
#include <cuda_fp16.h>

#include "launchers.h"
#include "tcgen05.cuh"

namespace clift {

namespace {

constexpr int kTile = CLIFT_TILE;
constexpr int kThreads = 256;
constexpr int kActRows = CLIFT_MAX_WIDTH;
constexpr int kSlabRows = 16;

// ---- tensor-core data-gradient engine of the backward kernel (tcgen05 kind::f16, 3-product fp16 split) ----------------
// dA[n][m] = sum_k W[k][n] dZ[k][m] for one 128-record tile: the M = 128 records are the MMA rows, the dZ rows of `act` are
// converted k-step by k-step (16 K rows) into a double-buffered fp16 (hi, lo) operand chunk, the weights are the
// clift_pack_linear_tc16() slabs of W^T streamed by bulk TMA through a small ring, the accumulator lives in tensor memory.
constexpr int kDgStages = 3;                 // weight ring depth (16 KB slabs: one k-step at N = 256)
constexpr int kDgStageBytes = 16384;
constexpr int kDgABytes = 8192;              // one k-step of the operand: [hi | lo][2 k-chunks][128 rows][8 halves]
constexpr int kDgGroup = 2;                  // k-steps converted, published and issued per barrier (operand buffer = a group)
constexpr int kDgBufBytes = kDgGroup * kDgABytes;
constexpr uint32_t kDgTmemCols = 256;
constexpr int kDgHeaderFloats = 16;          // header of a clift_pack_linear_tc16() operand

struct HeadsParams {
    const float4* rec_pos;
    const int32_t* rec_ray;
    const unsigned long long* stats;
    long long cap;
    const float* rays;
    FactorParams app;
    const float* basis_wt;
    int dim_app, pe_view, pe_feat, pe_sem, pe_ins;
    FactorParams semg, insg;          // grid-mode semantic / instance heads (comps == 0: MLP mode, input = xyz)
    const float* semg_basis;
    const float* insg_basis;
    int semg_dim, insg_dim;
    clift_mlp rgb, sem, insf, inss;
    int n_cls, d_ins, slow_fast, softmax, heads;
    float* rgb_raw;
    float* sem_raw;
    float* ins;
    float* rec_rgb;
    float* stash_a;     // training: per-tile layer inputs (StashLayout), else null
    StashLayout lay;
};

struct Smem {
    float* act;      // [kActRows][kTile]
    float* wslab;    // [2][kSlabRows][256]
    float4* pos;     // [kTile]
    int* ray;        // [kTile]
    int* runs;       // [kTile + 1]
    float* dir;      // [3][kTile]
};

// state of the tensor-core dgrad engine; the counters are identical in every thread (uniform control flow)
struct DgEngine {
    bool mask_pass;          // ReLU masks in a pass of their own instead of inside the next emit pass
    unsigned char* a_op;     // [2][kDgBufBytes]
    unsigned char* w_ring;   // [kDgStages][kDgStageBytes]
    uint64_t* w_full;        // [kDgStages]
    uint64_t* w_empty;       // [kDgStages]
    uint64_t* a_free;        // [2]
    uint64_t* d_done;
    float* red;              // [8]
    uint32_t tmem;
    uint32_t a_fills0, a_fills1;   // fills of each operand buffer so far (two scalars: a runtime-indexed array would put the
                                   // whole struct in local memory)
    uint32_t w_loads;        // weight slabs loaded so far (stage = w_loads % kDgStages)
    uint32_t w_used;         // weight slabs consumed so far
    uint32_t d_count;        // GEMMs completed so far
    bool on;
};
// the weight ring aliases the FMA path's two 16 KB cp.async slabs (the two engines never run at the same time) + one more
constexpr size_t kDgExtraBytes = 1024 + (size_t)(kDgStages - 2) * kDgStageBytes + 2 * kDgBufBytes + (2 * kDgStages + 3) * 8 + 64;
static_assert(kDgStageBytes == kSlabRows * 256 * 4, "ring stage = one FMA weight slab");

constexpr size_t kSmemBytes = (size_t)kActRows * kTile * 4 + 2 * kSlabRows * 256 * 4 + kTile * 16 + kTile * 4 +
                              (kTile + 4) * 4 + 3 * kTile * 4;

static_assert(kSmemBytes + kDgExtraBytes <= 232448, "backward kernel: shared memory budget of one sm_100a CTA");

// act[k][m] (all kpad rows valid) x W^T[k][n] -> act[n][m], bias, optional ReLU.  N = 64*NJ.
template <int NJ>
__device__ __forceinline__ void mlp_layer(const Smem& sm, const float* __restrict__ wt, const float* __restrict__ bias,
                                          int kpad, bool relu) {
    constexpr int N = 64 * NJ;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    float acc[8][4 * NJ];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4 * NJ; ++j) acc[i][j] = 0.0f;

    const int nslab = kpad / kSlabRows;
    auto issue = [&](int s) {
        float* dst = sm.wslab + (s & 1) * kSlabRows * 256;
        const float* src = wt + (size_t)s * kSlabRows * N;
        for (int i = tid; i < kSlabRows * N / 4; i += kThreads) cp_async16(dst + i * 4, src + i * 4);
        cp_async_commit();
    };
    issue(0);
    for (int s = 0; s < nslab; ++s) {
        if (s + 1 < nslab) {
            issue(s + 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float* ws = sm.wslab + (s & 1) * kSlabRows * 256;
        const float* arow = sm.act + (size_t)s * kSlabRows * kTile;
#pragma unroll 4
        for (int kk = 0; kk < kSlabRows; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(arow + kk * kTile + ty * 4);
            const float4 a1 = *reinterpret_cast<const float4*>(arow + kk * kTile + 64 + ty * 4);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const float4 b = *reinterpret_cast<const float4*>(ws + kk * N + j * 64 + tx * 4);
                const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int mi = 0; mi < 8; ++mi)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[mi][j * 4 + c] = fmaf(av[mi], bv[c], acc[mi][j * 4 + c]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int n = j * 64 + tx * 4 + c;
            const float b = bias ? __ldg(bias + n) : 0.0f;
            float v[8];
#pragma unroll
            for (int mi = 0; mi < 8; ++mi) {
                v[mi] = acc[mi][j * 4 + c] + b;
                if (relu) v[mi] = fmaxf(v[mi], 0.0f);
            }
            *reinterpret_cast<float4*>(sm.act + (size_t)n * kTile + ty * 4) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(sm.act + (size_t)n * kTile + 64 + ty * 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
    __syncthreads();
}

__device__ __forceinline__ void run_layer(const Smem& sm, const float* wt, const float* bias, int n_in, int n_out, bool relu) {
    const int kp = (n_in + 15) & ~15;
    const int np = (n_out + 63) & ~63;
    if (np == 64)
        mlp_layer<1>(sm, wt, bias, kp, relu);
    else if (np == 128)
        mlp_layer<2>(sm, wt, bias, kp, relu);
    else
        mlp_layer<4>(sm, wt, bias, kp, relu);
}

// rows [0,rows) of act -> a stash block of `rows` rows (stash_idx layout: a warp writes one row as 8 pieces of 64 bytes)
__device__ __forceinline__ void store_rows(const Smem& sm, float* __restrict__ dst, int rows) {
    float4* d4 = reinterpret_cast<float4*>(dst);
    const float4* s4 = reinterpret_cast<const float4*>(sm.act);
    for (int i = threadIdx.x; i < rows * (kTile / 4); i += kThreads) d4[stash_idx4(rows, i >> 5, i & 31)] = s4[i];
}

// `stash` (may be null) = this tile's A-stash base; a_off = the MLP's row offsets: the INPUT of every layer is saved.
__device__ __forceinline__ void run_mlp(const Smem& sm, const clift_mlp& mlp, float* stash = nullptr, const int* a_off = nullptr) {
    for (int l = 0; l < mlp.n_layers; ++l) {
        if (stash) store_rows(sm, stash + (size_t)a_off[l] * kTile, (mlp.dims[l] + 15) & ~15);
        run_layer(sm, mlp.wt[l], mlp.bias[l], mlp.dims[l], mlp.dims[l + 1], l + 1 < mlp.n_layers);
    }
}

// rows [0,3) = xyz, then sin/cos positional encoding (dimension-major, frequency-minor), zero pad to 16.
__device__ __forceinline__ void build_xyz_input(const Smem& sm, int pe) {
    const int n_in = 3 + 6 * pe;
    const int kp = (n_in + 15) & ~15;
    for (int idx = threadIdx.x; idx < kp * kTile; idx += kThreads) {
        const int r = idx / kTile, m = idx - r * kTile;
        const float4 p = sm.pos[m];
        const float xyz[3] = {p.x, p.y, p.z};
        float v = 0.0f;
        if (r < 3) {
            v = xyz[r];
        } else if (r < n_in) {
            const int j = (r - 3) % (3 * pe);
            const float arg = xyz[j / pe] * (float)(1 << (j % pe));
            v = (r - 3 < 3 * pe) ? sinf(arg) : cosf(arg);
        }
        sm.act[idx] = v;
    }
    __syncthreads();
}

// Sum rows [0,nch) of act over each ray run (fixed order) and add into dst[ray*stride + col0 + c].
__device__ __forceinline__ void reduce_runs(const Smem& sm, int n_runs, int nch, float* __restrict__ dst, int stride, int col0) {
    for (int idx = threadIdx.x; idx < n_runs * nch; idx += kThreads) {
        const int r = idx / nch, c = idx - r * nch;
        const int m0 = sm.runs[r], m1 = sm.runs[r + 1];
        float s = 0.0f;
        for (int m = m0; m < m1; ++m) s += sm.act[(size_t)c * kTile + m];
        atomicAdd(dst + (int64_t)sm.ray[m0] * stride + col0 + c, s);
    }
    __syncthreads();
}

template <int NV>
__device__ __forceinline__ void gather_appearance(const Smem& sm, const FactorParams& f) {
    const int q = threadIdx.x & 3;
#pragma unroll
    for (int pass = 0; pass < kTile / 64; ++pass) {
        const int m = pass * 64 + (threadIdx.x >> 2);
        const float4 p = sm.pos[m];
        const float xs[3] = {p.x, p.y, p.z};
#pragma unroll
        for (int mode = 0; mode < 3; ++mode) {
            const Tap2 t2 = make_tap2(xs[mode_a(mode)], xs[mode_b(mode)], f.pw[mode], f.ph[mode]);
            const Tap1 t1 = make_tap1(xs[mode_v(mode)], f.ll[mode]);
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int ch = v * 16 + q * 4;
                const float4 pv = plane_tap(f.plane[mode], t2, f.pw[mode], f.comps, ch);
                const float4 lv = line_tap(f.line[mode], t1, f.comps, ch);
                float* dst = sm.act + (size_t)(mode * f.comps + ch) * kTile + m;
                dst[0] = pv.x * lv.x;
                dst[kTile] = pv.y * lv.y;
                dst[2 * kTile] = pv.z * lv.z;
                dst[3 * kTile] = pv.w * lv.w;
            }
        }
    }
    __syncthreads();
}

// Grid-mode head input (tensoRF.py:127-134, 142-156): plane*line products of the head's own VM factor set (rows
// [0, 3*comps), saved to `stash_rows` for the basis weight gradient) -> bias-free basis Linear -> act rows [0, 64)
// (rows >= dim are exact zeros: the packed weight is zero padded).
__device__ __forceinline__ void grid_head_input(const Smem& sm, const FactorParams& f, const float* basis_wt, int dim,
                                                float* stash_rows) {
    switch (f.comps >> 4) {
        case 1: gather_appearance<1>(sm, f); break;
        case 2: gather_appearance<2>(sm, f); break;
        case 3: gather_appearance<3>(sm, f); break;
        default: gather_appearance<4>(sm, f); break;
    }
    if (stash_rows) store_rows(sm, stash_rows, 3 * f.comps);
    run_layer(sm, basis_wt, nullptr, 3 * f.comps, dim, false);
}

// rows [A, in) of the RGB MLP input: viewdir, sin/cos PE of the 27 features, sin/cos PE of the viewdir
__device__ __forceinline__ void build_rgb_input(const Smem& sm, int A, int pf, int pv) {
    const int o_sf = A + 3, o_cf = o_sf + A * pf, o_sd = o_cf + A * pf, o_cd = o_sd + 3 * pv, n_in = o_cd + 3 * pv;
    const int kp = (n_in + 15) & ~15;
    for (int idx = threadIdx.x; idx < (kp - A) * kTile; idx += kThreads) {
        const int r = A + idx / kTile, m = idx % kTile;
        float v = 0.0f;
        if (r < o_sf) {
            v = sm.dir[(r - A) * kTile + m];
        } else if (r < o_sd) {
            const int j = (r < o_cf) ? r - o_sf : r - o_cf;
            const float arg = sm.act[(size_t)(j / pf) * kTile + m] * (float)(1 << (j % pf));
            v = (r < o_cf) ? sinf(arg) : cosf(arg);
        } else if (r < n_in) {
            const int j = (r < o_cd) ? r - o_sd : r - o_cd;
            const float arg = sm.dir[(j / pv) * kTile + m] * (float)(1 << (j % pv));
            v = (r < o_cd) ? sinf(arg) : cosf(arg);
        }
        sm.act[(size_t)r * kTile + m] = v;
    }
    __syncthreads();
}

template <int NV>
__global__ void __launch_bounds__(kThreads, 1) heads_forward_kernel(const __grid_constant__ HeadsParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_nruns;
    Smem sm;
    sm.act = reinterpret_cast<float*>(smem_raw);
    sm.wslab = sm.act + kActRows * kTile;
    sm.pos = reinterpret_cast<float4*>(sm.wslab + 2 * kSlabRows * 256);
    sm.ray = reinterpret_cast<int*>(sm.pos + kTile);
    sm.runs = sm.ray + kTile;
    sm.dir = reinterpret_cast<float*>(sm.runs + kTile + 4);

    const long long n_act = min((long long)P.stats[0], P.cap);
    const long long n_tiles = (n_act + kTile - 1) / kTile;
    const int tid = threadIdx.x;

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long base = tile * kTile;
        const int nv = (int)min((long long)kTile, n_act - base);
        if (tid < kTile) {
            float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
            int ray = -1;
            float d0 = 0.f, d1 = 0.f, d2 = 1.f;
            if (tid < nv) {
                p = P.rec_pos[base + tid];
                ray = P.rec_ray[base + tid];
                d0 = __ldg(P.rays + (int64_t)ray * 8 + 3);
                d1 = __ldg(P.rays + (int64_t)ray * 8 + 4);
                d2 = __ldg(P.rays + (int64_t)ray * 8 + 5);
            }
            sm.pos[tid] = p;
            sm.ray[tid] = ray;
            sm.dir[tid] = d0;
            sm.dir[kTile + tid] = d1;
            sm.dir[2 * kTile + tid] = d2;
        }
        __syncthreads();
        if (tid < 32) {   // run starts, in order
            int n = 0;
            for (int w = 0; w < kTile / 32; ++w) {
                const int m = w * 32 + tid;
                const bool start = m < nv && (m == 0 || sm.ray[m] != sm.ray[m - 1]);
                const unsigned bits = __ballot_sync(0xffffffffu, start);
                if (start) sm.runs[n + __popc(bits & ((1u << tid) - 1u))] = m;
                n += __popc(bits);
            }
            if (tid == 0) {
                sm.runs[n] = nv;
                s_nruns = n;
            }
        }
        __syncthreads();
        const int n_runs = s_nruns;

        float* stash = P.stash_a ? P.stash_a + (size_t)tile * P.lay.a_rows * kTile : nullptr;
        if (P.heads & CLIFT_HEAD_SEMANTIC) {
            if (P.semg.comps)
                grid_head_input(sm, P.semg, P.semg_basis, P.semg_dim, stash ? stash + (size_t)P.lay.a_off[5][0] * kTile : nullptr);
            else
                build_xyz_input(sm, P.pe_sem);
            run_mlp(sm, P.sem, stash, P.lay.a_off[0]);
            if (P.softmax) {
                if (tid < kTile) {
                    float mx = -INFINITY;
                    for (int c = 0; c < P.n_cls; ++c) mx = fmaxf(mx, sm.act[(size_t)c * kTile + tid]);
                    float tot = 0.0f;
                    for (int c = 0; c < P.n_cls; ++c) {
                        const float e = expf(sm.act[(size_t)c * kTile + tid] - mx);
                        sm.act[(size_t)c * kTile + tid] = e;
                        tot += e;
                    }
                    for (int c = 0; c < P.n_cls; ++c) sm.act[(size_t)c * kTile + tid] /= tot;
                }
                __syncthreads();
                if (stash) {
                    store_rows(sm, stash + (size_t)P.lay.prob_off * kTile, P.n_cls);
                    __syncthreads();
                }
            }
            for (int idx = tid; idx < P.n_cls * kTile; idx += kThreads) sm.act[idx] *= sm.pos[idx % kTile].w;
            __syncthreads();
            reduce_runs(sm, n_runs, P.n_cls, P.sem_raw, P.n_cls, 0);
        }
        if (P.heads & CLIFT_HEAD_INSTANCE) {
            const int width = P.d_ins * (P.slow_fast ? 2 : 1);
            for (int net = 0; net < (P.slow_fast ? 2 : 1); ++net) {
                if (P.insg.comps)     // both nets read the same basis feature (tensoRF.py:497-511); re-gathered per net
                    grid_head_input(sm, P.insg, P.insg_basis, P.insg_dim,
                                    stash && net == 0 ? stash + (size_t)P.lay.a_off[6][0] * kTile : nullptr);
                else
                    build_xyz_input(sm, P.pe_ins);
                run_mlp(sm, net == 0 ? P.insf : P.inss, stash, P.lay.a_off[1 + net]);
                for (int idx = tid; idx < P.d_ins * kTile; idx += kThreads) sm.act[idx] *= sm.pos[idx % kTile].w;
                __syncthreads();
                reduce_runs(sm, n_runs, P.d_ins, P.ins, width, net * P.d_ins);
            }
        }
        if (P.heads & CLIFT_HEAD_RGB) {
            gather_appearance<NV>(sm, P.app);
            if (stash) store_rows(sm, stash + (size_t)P.lay.a_off[4][0] * kTile, 3 * P.app.comps);
            run_layer(sm, P.basis_wt, nullptr, 3 * P.app.comps, P.dim_app, false);
            build_rgb_input(sm, P.dim_app, P.pe_feat, P.pe_view);
            run_mlp(sm, P.rgb, stash, P.lay.a_off[3]);
            for (int idx = tid; idx < 3 * kTile; idx += kThreads) {
                const int m = idx % kTile;
                const float c = 1.0f / (1.0f + expf(-sm.act[idx]));
                if (P.rec_rgb && m < nv) P.rec_rgb[(base + m) * 4 + idx / kTile] = c;
                sm.act[idx] = c * sm.pos[m].w;
            }
            __syncthreads();
            reduce_runs(sm, n_runs, 3, P.rgb_raw, 3, 0);
        }
    }
}


// =================================================================================================
// Backward of the heads (training).  Two kernel families:
//   heads_backward_kernel : per tile, fused across layers - output gradient, softmax/sigmoid backward, data gradients
//                           layer by layer with the dgrad-packed weights (same FFMA tile GEMM as the forward), ReLU
//                           masks from the A-stash, bias gradients, dL/dw through rgb, positional-encoding backward and
//                           the scatter of the appearance-factor gradients.  Every layer's dZ goes to the Z-stash.
//   wgrad_kernel          : per layer, dW^T[k][n] += sum over records A[k][m] * dZ[n][m]  (split over tiles).
// =================================================================================================
struct HeadsBwdParams {
    const float4* rec_pos;
    const int32_t* rec_ray;
    const int32_t* rec_idx;
    const float* rec_rgb;
    const unsigned long long* stats;
    long long cap;
    int S;
    FactorParams app;
    float* g_app_plane[3];
    float* g_app_line[3];
    const float* basis_dgrad;
    const void* basis_dg16;           // tensor-core operand of the basis data gradient (null: FP32 FMA)
    int use_tc;                       // 1: data gradients with a w_dg16 operand run on tcgen05
    int mask_pass;                    // 1: ReLU masks in a pass of their own (development switch CLIFT_BWD_MASK_PASS=1)
    int dim_app, pe_view, pe_feat;
    FactorParams semg, insg;          // grid-mode semantic / instance heads (comps == 0: MLP mode)
    float* g_semg_plane[3];
    float* g_semg_line[3];
    float* g_insg_plane[3];
    float* g_insg_line[3];
    const float* semg_basis_dgrad;
    const float* insg_basis_dgrad;
    int semg_dim, insg_dim;
    clift_mlp rgb, sem, insf, inss;
    clift_mlp_grad g_rgb_mlp, g_sem_mlp, g_insf_mlp, g_inss_mlp;
    int n_cls, d_ins, slow_fast, softmax;
    int do_rgb, do_sem, do_ins;
    const float* g_ray;      // [B][ray_stride]: g_rgb_raw(3) | g_sem_raw(C) | g_ins(DI) | g_opacity
    int ray_stride, off_sem, off_ins;
    float* g_w;              // [B*S]
    const float* stash_a;
    float* stash_z;
    StashLayout lay;
};

__device__ __forceinline__ void red_add4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// rows [0, rows_pad) of act -> Z-stash; rows [0, n_out) summed over the tile into the bias gradient.
__device__ __forceinline__ void emit_dz(const Smem& sm, float* __restrict__ zdst, int rows_pad, int n_out, float* __restrict__ g_bias) {
    // one pass: a warp iteration covers exactly one row (32 lanes x float4 = 128 records), so the row's stash store and its
    // bias-gradient sum (fixed-order lane tree, one atomic per row and tile) share the shared-memory read
    float4* d4 = reinterpret_cast<float4*>(zdst);
    const float4* s4 = reinterpret_cast<const float4*>(sm.act);
    const int lane = threadIdx.x & 31;
    const int total = rows_pad * (kTile / 4);            // a multiple of 8 * kThreads for rows_pad in {64, 128, 256}
    for (int i0 = threadIdx.x; i0 < total; i0 += 8 * kThreads) {
        float part[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * kThreads;
            const float4 v = s4[i];
            d4[stash_idx4(rows_pad, i >> 5, i & 31)] = v;
            part[u] = (v.x + v.y) + (v.z + v.w);
        }
        if (g_bias) {      // the eight rows' lane trees side by side (same fixed order per row as warp_sum)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int u = 0; u < 8; ++u) part[u] += __shfl_xor_sync(0xffffffffu, part[u], o);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int row = (i0 + u * kThreads) >> 5;
                if (lane == 0 && row < n_out) atomicAdd(g_bias + row, part[u]);
            }
        }
    }
}

// the `rows` x 128 floats of a stash block -> L2 (one 128-byte line per request), issued before a data-gradient GEMM for the
// block the pass after it reads: its loads then cost an L2 hit instead of a DRAM round trip
__device__ __forceinline__ void prefetch_block(const float* __restrict__ blk, int rows) {
    const char* p = reinterpret_cast<const char*>(blk);
    const int lines = rows * kTile * 4 / 128;
    for (int i = threadIdx.x; i < lines; i += kThreads) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + (size_t)i * 128));
}

// emit_dz with the ReLU mask of the layer above folded in: act rows [0, rows_pad) hold dL/d(post-ReLU output) of this layer;
// `a_src` = the A-stash block of that output (rows_pad rows, saved as the next layer's input): where it is <= 0 the gradient
// is zeroed - in shared memory (the data-gradient GEMM reads it next), in the Z-stash and in the bias sums - in one pass
// whose stash loads run four deep under the stores of the previous group.  Ends with the barrier the mask pass had.
__device__ __forceinline__ void emit_dz_masked(const Smem& sm, float* __restrict__ zdst, int rows_pad, int n_out,
                                               float* __restrict__ g_bias, const float* __restrict__ a_src) {
    float4* d4 = reinterpret_cast<float4*>(zdst);
    float4* s4 = reinterpret_cast<float4*>(sm.act);
    const float4* a4 = reinterpret_cast<const float4*>(a_src);
    const int lane = threadIdx.x & 31;
    const int total = rows_pad * (kTile / 4);            // a multiple of 8 * kThreads for rows_pad in {64, 128, 256}
    for (int i0 = threadIdx.x; i0 < total; i0 += 8 * kThreads) {
        float4 a[8];                                     // eight stash loads in flight per thread (L2 hits: prefetch_block)
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * kThreads;
            a[u] = __ldg(a4 + stash_idx4(rows_pad, i >> 5, i & 31));
        }
        float part[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * kThreads, row = i >> 5;
            float4 v = s4[i];
            v.x = a[u].x > 0.0f ? v.x : 0.0f;
            v.y = a[u].y > 0.0f ? v.y : 0.0f;
            v.z = a[u].z > 0.0f ? v.z : 0.0f;
            v.w = a[u].w > 0.0f ? v.w : 0.0f;
            s4[i] = v;
            d4[stash_idx4(rows_pad, row, i & 31)] = v;
            part[u] = (v.x + v.y) + (v.z + v.w);
        }
        if (g_bias) {      // the eight rows' lane trees side by side (same fixed order per row as warp_sum)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int u = 0; u < 8; ++u) part[u] += __shfl_xor_sync(0xffffffffu, part[u], o);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int row = (i0 + u * kThreads) >> 5;
                if (lane == 0 && row < n_out) atomicAdd(g_bias + row, part[u]);
            }
        }
    }
    __syncthreads();
}

// dA[k][m] = sum_n W[n][k] dZ[n][m] with the clift_pack_linear_dgrad() operand ([up16(out)][dgrad_pad(in)])
__device__ __forceinline__ void run_dgrad(const Smem& sm, const float* w_dgrad, int n_out, int n_in) {
    const int np = n_in <= 64 ? 64 : (n_in <= 128 ? 128 : 256);
    const int kp = (n_out + 15) & ~15;
    if (np == 64)
        mlp_layer<1>(sm, w_dgrad, nullptr, kp, false);
    else if (np == 128)
        mlp_layer<2>(sm, w_dgrad, nullptr, kp, false);
    else
        mlp_layer<4>(sm, w_dgrad, nullptr, kp, false);
}

__device__ __forceinline__ uint32_t dg_bits(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

// thread 0: slab `idx` of the running GEMM -> next ring stage (waits until the MMAs that read that stage have completed)
__device__ __forceinline__ void dg_load(const DgEngine& E, const unsigned char* slabs, uint32_t slab_bytes, int idx) {
    const uint32_t st = E.w_loads % kDgStages, round = E.w_loads / kDgStages;
    if (round > 0) tc::mbar_wait(&E.w_empty[st], (round - 1) & 1u);
    tc::mbar_arrive_expect_tx(&E.w_full[st], slab_bytes);
    tc::bulk_load(E.w_ring + (size_t)st * kDgStageBytes, slabs + (size_t)idx * slab_bytes, slab_bytes, &E.w_full[st]);
}

// Tensor-core form of run_dgrad: act rows [0, round16(K)) = dZ[k][m] (rows >= K zero) -> act rows [0, out_rows) = dA[n][m]
// (rows >= n_in zero).  `w16` = clift_pack_linear_tc16() operand of W^T (K = the layer's out width, N = its in width).
// The operand scale is dynamic: s = 2^floor(log2(2^14 / max|dZ|)) over the tile (exact power of two), so every scaled
// dZ is below 2^14 and splits into fp16 (hi, lo) with 22 significant bits; the weight scale is the pack-time one.
// `mask` (may be null): the A-stash block of this layer's input (mask_rows rows, stash_idx layout) - the ReLU mask
// (input > 0) is applied while the accumulator is read back, its loads issued one 16-column chunk ahead (the first before
// the wait for the MMAs), so the separate mask pass over the activation tile and its exposed HBM latency disappear.
__device__ __forceinline__ void run_dgrad_tc(const Smem& sm, DgEngine& E, const void* w16, int K, int n_in, int out_rows,
                                             const float* __restrict__ mask = nullptr, int mask_rows = 0) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k_steps = (K + 15) >> 4, kp = k_steps * 16;
    const int n_pad = (n_in + 31) & ~31;
    float mx = 0.0f;
    {
        const float4* a4 = reinterpret_cast<const float4*>(sm.act);
        for (int i = tid; i < kp * (kTile / 4); i += kThreads) {
            const float4 v = a4[i];
            mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) E.red[warp] = mx;
        __syncthreads();
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) mx = fmaxf(mx, E.red[w]);
    }
    if (mx == 0.0f) {       // nothing flows back through this layer for this tile (e.g. the detached slow net): dA = 0
        float4* d4 = reinterpret_cast<float4*>(sm.act);
        for (int i = tid; i < out_rows * (kTile / 4); i += kThreads) d4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
        return;
    }
    int ex = 0;
    frexpf(mx, &ex);                                   // mx < 2^ex
    const float sa = ldexpf(1.0f, max(-60, min(60, 14 - ex)));
    const float* meta = reinterpret_cast<const float*>(w16);
    const float inv = 1.0f / (sa * meta[1]);
    const unsigned char* slabs = reinterpret_cast<const unsigned char*>(meta + kDgHeaderFloats);
    const uint32_t slab_bytes = 64u * (uint32_t)n_pad;
    const bool stacked = n_pad <= 128;
    const int n_pre = min(kDgStages, k_steps);
    for (int i = 0; i < n_pre; ++i) {
        if (tid == 0) dg_load(E, slabs, slab_bytes, i);
        ++E.w_loads;
    }
    const int m = tid & (kTile - 1), c = tid >> 7;
    // kDgGroup k-steps per round: every thread converts its share of the group, one barrier publishes it, thread 0 issues the
    // group's MMAs (its fixed issue cost - waits, descriptors, commits - and the barrier are paid per group, not per k-step)
    for (int ks0 = 0; ks0 < k_steps; ks0 += kDgGroup) {
        const int b = (ks0 / kDgGroup) & 1, ks1 = min(k_steps, ks0 + kDgGroup);
        const uint32_t fills = b ? E.a_fills1 : E.a_fills0;
        if (fills > 0) tc::mbar_wait(&E.a_free[b], (fills - 1) & 1u);
        if (b)
            ++E.a_fills1;
        else
            ++E.a_fills0;
        for (int ks = ks0; ks < ks1; ++ks) {
            // records m, K rows [16 ks + 8 c, + 8) -> one 16-byte row of the hi chunk and one of the lo chunk
            const float* src = sm.act + (size_t)(ks * 16 + c * 8) * kTile + m;
            uint32_t h[4], l[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float x = src[(2 * i) * kTile] * sa, y = src[(2 * i + 1) * kTile] * sa;
                const __half2 hh = __floats2half2_rn(x, y);
                const float2 back = __half22float2(hh);
                h[i] = dg_bits(hh);
                l[i] = dg_bits(__floats2half2_rn(x - back.x, y - back.y));
            }
            unsigned char* dst = E.a_op + (size_t)b * kDgBufBytes + (size_t)(ks - ks0) * kDgABytes + ((size_t)c * kTile + m) * 16;
            *reinterpret_cast<uint4*>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(dst + kDgABytes / 2) = make_uint4(l[0], l[1], l[2], l[3]);
        }
        tc::fence_proxy_async_smem();
        __syncthreads();
        for (int ks = ks0; ks < ks1; ++ks) {        // every thread keeps the ring counters; thread 0 does the work
            if (tid == 0) {
                const uint32_t st = E.w_used % kDgStages;
                tc::mbar_wait(&E.w_full[st], (E.w_used / kDgStages) & 1u);
                tc::fence_after_sync();
                constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);                 // SBO 128 B, descriptor version 1
                auto desc = [](uint32_t lo) { return ((uint64_t)kDescHi << 32) | lo; };
                const uint32_t a_lbo = (kTile * 16u >> 4) << 16;
                const unsigned char* a_base = E.a_op + (size_t)b * kDgBufBytes + (size_t)(ks - ks0) * kDgABytes;
                const uint32_t ah = (tc::smem_addr(a_base) >> 4) | a_lbo;
                const uint32_t al = (tc::smem_addr(a_base + kDgABytes / 2) >> 4) | a_lbo;
                const uint32_t w_lo = tc::smem_addr(E.w_ring + (size_t)st * kDgStageBytes) >> 4;
                const uint32_t rows1 = stacked ? 2u * n_pad : (uint32_t)n_pad;
                const uint32_t idesc = tc::make_idesc_f16(kTile, n_pad);
                const uint64_t b1 = desc(w_lo | (rows1 << 16));
                const uint32_t acc = ks > 0 ? 1u : 0u;
                if (stacked) {      // slab = [2 k-chunks][hi | lo][n_pad][8]: A_hi*[W_hi ; W_lo] in one MMA, then A_lo*W_hi
                    tc::mma_ss_f16(E.tmem, desc(ah), b1, tc::make_idesc_f16(kTile, 2 * n_pad), acc);
                    tc::mma_ss_f16(E.tmem, desc(al), b1, idesc, 1u);
                } else {            // slab = [hi | lo][2 k-chunks][n_pad][8]
                    tc::mma_ss_f16(E.tmem, desc(ah), b1, idesc, acc);
                    tc::mma_ss_f16(E.tmem, desc(ah), desc((w_lo + 2u * rows1) | (rows1 << 16)), idesc, 1u);
                    tc::mma_ss_f16(E.tmem, desc(al), b1, idesc, 1u);
                }
                tc::mma_commit(&E.w_empty[st]);
                if (ks == ks1 - 1) tc::mma_commit(&E.a_free[b]);
                if (ks == k_steps - 1) tc::mma_commit(E.d_done);
                // refill one k-step late: the stage of slab ks - 1 is free (or about to be) while this k-step's MMAs run
                if (ks >= 1 && ks - 1 + n_pre < k_steps) dg_load(E, slabs, slab_bytes, ks - 1 + n_pre);
            }
            ++E.w_used;
            if (ks >= 1 && ks - 1 + n_pre < k_steps) ++E.w_loads;
        }
    }
    const int q = warp & 3, mrow = q * 32 + lane;
    float mk_next[16];
    auto load_mask = [&](int c0, float* mk) {
#pragma unroll
        for (int i = 0; i < 16; ++i) mk[i] = (c0 + i < mask_rows) ? __ldg(mask + stash_idx(mask_rows, c0 + i, mrow)) : 1.0f;
    };
    if (mask && (warp >> 2) * 16 < n_pad) load_mask((warp >> 2) * 16, mk_next);
    tc::mbar_wait(E.d_done, E.d_count & 1u);
    ++E.d_count;
    tc::fence_after_sync();
    {   // accumulator (lane = record, column = n) -> act[n][m]; warps w and w + 4 share a lane quarter and split the columns
        const uint32_t taddr = E.tmem + ((uint32_t)(q * 32) << 16);
        for (int c0 = (warp >> 2) * 16; c0 < n_pad; c0 += 32) {
            float mk[16];
            if (mask) {
#pragma unroll
                for (int i = 0; i < 16; ++i) mk[i] = mk_next[i];
                if (c0 + 32 < n_pad) load_mask(c0 + 32, mk_next);
            }
            float v[16];
            tc::tmem_ld16(taddr + (uint32_t)c0, v);
            if (stacked) {
                float u[16];
                tc::tmem_ld16(taddr + (uint32_t)(n_pad + c0), u);
                tc::tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] += u[i];
            } else {
                tc::tmem_wait_ld();
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) sm.act[(size_t)(c0 + i) * kTile + mrow] = (!mask || mk[i] > 0.0f) ? v[i] * inv : 0.0f;
        }
        float4* d4 = reinterpret_cast<float4*>(sm.act + (size_t)n_pad * kTile);
        for (int i = tid; i < (out_rows - n_pad) * (kTile / 4); i += kThreads) d4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    tc::fence_before_sync();
    __syncthreads();
}

// act rows [0, n_pad(out)) hold dZ of the LAST layer (pad rows zero).  Walks the stack backwards.
// On return (need_input_grad) act rows [0, dgrad_pad(dims[0])) hold dL/d(input of layer 0).
__device__ __forceinline__ void mlp_backward(const Smem& sm, DgEngine& E, const clift_mlp& mlp, const clift_mlp_grad& g,
                                             const float* stash_a, float* stash_z, const int* a_off, const int* z_off,
                                             bool need_input_grad) {
    const float* mask_src = nullptr;      // A-stash block whose sign masks act; applied by the next layer's emit pass
    for (int l = mlp.n_layers - 1; l >= 0; --l) {
        const int n_out = mlp.dims[l + 1], n_in = mlp.dims[l];
        if (mask_src)
            emit_dz_masked(sm, stash_z + (size_t)z_off[l] * kTile, (n_out + 63) & ~63, n_out, g.bias[l], mask_src);
        else
            emit_dz(sm, stash_z + (size_t)z_off[l] * kTile, (n_out + 63) & ~63, n_out, g.bias[l]);
        mask_src = nullptr;
        if (l == 0 && !need_input_grad) break;
        if (l > 0) prefetch_block(stash_a + (size_t)a_off[l] * kTile, (n_in + 15) & ~15);      // the mask this GEMM's output meets
        // (folding the ReLU mask into run_dgrad_tc's accumulator read-back - its `mask` argument - was measured SLOWER:
        // 16 scalar stash loads per chunk and 32 more live registers cost more than the float4 mask pass below)
        if (E.on && mlp.w_dg16[l])
            run_dgrad_tc(sm, E, mlp.w_dg16[l], n_out, n_in, n_in <= 64 ? 64 : (n_in <= 128 ? 128 : 256));
        else
            run_dgrad(sm, mlp.w_dgrad[l], n_out, n_in);
        // ReLU mask from the saved input of layer l (= post-ReLU output of layer l-1).  Hidden widths that are multiples of
        // 64 (every shipped stack: 128 / 256): the stash block and the Z-stash block of layer l-1 have the same row count, so
        // the mask rides in that layer's emit pass instead of a pass of its own
        if (l > 0 && (n_in & 63) == 0 && !E.mask_pass) {
            mask_src = stash_a + (size_t)a_off[l] * kTile;       // (both data-gradient forms end with a barrier)
        } else if (l > 0) {
            const float4* a4 = reinterpret_cast<const float4*>(stash_a + (size_t)a_off[l] * kTile);
            float4* d4 = reinterpret_cast<float4*>(sm.act);
            const int rows = (n_in + 15) & ~15;
            for (int i = threadIdx.x; i < rows * (kTile / 4); i += kThreads) {
                const float4 a = a4[stash_idx4(rows, i >> 5, i & 31)];
                float4 d = d4[i];
                d.x = a.x > 0.0f ? d.x : 0.0f;
                d.y = a.y > 0.0f ? d.y : 0.0f;
                d.z = a.z > 0.0f ? d.z : 0.0f;
                d.w = a.w > 0.0f ? d.w : 0.0f;
                d4[i] = d;
            }
            // rows [rows, n_pad(n_in)) are zero already (dgrad weights are zero padded)
            __syncthreads();
        }
    }
    __syncthreads();
}

// act rows [0, 3*comps) = dL/d(plane*line product) of the tile's records -> factor gradients of the set `f`
template <int NV>
__device__ __forceinline__ void scatter_factors(const Smem& sm, const FactorParams& f, float* const* g_plane, float* const* g_line,
                                                int nv) {
    const int q = threadIdx.x & 3;
#pragma unroll 1
    for (int pass = 0; pass < kTile / 64; ++pass) {
        const int m = pass * 64 + (threadIdx.x >> 2);
        if (m >= nv) continue;
        const float4 p = sm.pos[m];
        const float xs[3] = {p.x, p.y, p.z};
#pragma unroll
        for (int mode = 0; mode < 3; ++mode) {
            const int W = f.pw[mode];
            const Tap2 t2 = make_tap2(xs[mode_a(mode)], xs[mode_b(mode)], W, f.ph[mode]);
            const Tap1 t1 = make_tap1(xs[mode_v(mode)], f.ll[mode]);
            const int64_t row0 = (int64_t)t2.y0 * W, row1 = row0 + W;
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int ch = v * 16 + q * 4;
                const float4 pv = plane_tap(f.plane[mode], t2, W, f.comps, ch);
                const float4 lv = line_tap(f.line[mode], t1, f.comps, ch);
                const float* g = sm.act + (size_t)(mode * f.comps + ch) * kTile + m;
                const float g0 = g[0], g1 = g[kTile], g2 = g[2 * kTile], g3 = g[3 * kTile];
                // d plane = g * L * bilinear weight ; d line = g * P * linear weight
                const float a0 = g0 * lv.x, a1 = g1 * lv.y, a2 = g2 * lv.z, a3 = g3 * lv.w;
                float* gp = g_plane[mode];
                if (t2.w00 != 0.0f) red_add4(gp + (row0 + t2.x0) * f.comps + ch, a0 * t2.w00, a1 * t2.w00, a2 * t2.w00, a3 * t2.w00);
                if (t2.w10 != 0.0f) red_add4(gp + (row0 + t2.x0 + 1) * f.comps + ch, a0 * t2.w10, a1 * t2.w10, a2 * t2.w10, a3 * t2.w10);
                if (t2.w01 != 0.0f) red_add4(gp + (row1 + t2.x0) * f.comps + ch, a0 * t2.w01, a1 * t2.w01, a2 * t2.w01, a3 * t2.w01);
                if (t2.w11 != 0.0f) red_add4(gp + (row1 + t2.x0 + 1) * f.comps + ch, a0 * t2.w11, a1 * t2.w11, a2 * t2.w11, a3 * t2.w11);
                const float b0 = g0 * pv.x, b1 = g1 * pv.y, b2 = g2 * pv.z, b3 = g3 * pv.w;
                float* gl = g_line[mode];
                if (t1.w0 != 0.0f) red_add4(gl + (int64_t)t1.z0 * f.comps + ch, b0 * t1.w0, b1 * t1.w0, b2 * t1.w0, b3 * t1.w0);
                if (t1.w1 != 0.0f) red_add4(gl + (int64_t)(t1.z0 + 1) * f.comps + ch, b0 * t1.w1, b1 * t1.w1, b2 * t1.w1, b3 * t1.w1);
            }
        }
    }
    __syncthreads();
}

// Grid-mode head input backward: act rows [0, 64) = dL/d(basis feature) -> Z-stash (basis weight gradient) -> products
// gradient through basis^T -> scatter into the head's own factor gradients.
__device__ __forceinline__ void grid_head_backward(const Smem& sm, const FactorParams& f, const float* basis_dgrad, int dim,
                                                   float* z_rows, float* const* g_plane, float* const* g_line, int nv) {
    store_rows(sm, z_rows, 64);
    run_dgrad(sm, basis_dgrad, dim, 3 * f.comps);
    switch (f.comps >> 4) {
        case 1: scatter_factors<1>(sm, f, g_plane, g_line, nv); break;
        case 2: scatter_factors<2>(sm, f, g_plane, g_line, nv); break;
        case 3: scatter_factors<3>(sm, f, g_plane, g_line, nv); break;
        default: scatter_factors<4>(sm, f, g_plane, g_line, nv); break;
    }
}

template <int NV>
__global__ void __launch_bounds__(kThreads, 1) heads_backward_kernel(const __grid_constant__ HeadsBwdParams P) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Smem sm;
    sm.act = reinterpret_cast<float*>(smem_raw);
    sm.wslab = sm.act + kActRows * kTile;
    sm.pos = reinterpret_cast<float4*>(sm.wslab + kDgStages * kSlabRows * 256);   // 3 slabs here: the dgrad weight ring
    sm.ray = reinterpret_cast<int*>(sm.pos + kTile);
    sm.runs = sm.ray + kTile;
    sm.dir = reinterpret_cast<float*>(sm.runs + kTile + 4);   // here: [3][kTile] saved rgb of the records
    const int tid = threadIdx.x;

    // tensor-core dgrad engine: weight ring = the FMA path's cp.async slabs + one more 16 KB stage (the two engines never
    // run at the same time), operand chunks, barriers, tensor memory
    DgEngine E;
    E.on = P.use_tc != 0;
    E.mask_pass = P.mask_pass != 0;
    E.w_ring = reinterpret_cast<unsigned char*>(sm.wslab);
    {
        // layout after the FMA regions: [pos | ray | runs | dir] stay where they are; the engine's extra stage, operand
        // chunks and barriers follow them, 1 KB aligned
        uintptr_t p = reinterpret_cast<uintptr_t>(sm.dir + 3 * kTile);
        p = (p + 1023) & ~(uintptr_t)1023;
        unsigned char* q = reinterpret_cast<unsigned char*>(p);
        E.a_op = q;
        q += 2 * kDgBufBytes;
        E.w_full = reinterpret_cast<uint64_t*>(q);
        E.w_empty = E.w_full + kDgStages;
        E.a_free = E.w_empty + kDgStages;
        E.d_done = E.a_free + 2;
        E.red = reinterpret_cast<float*>(E.d_done + 1);
    }
    E.a_fills0 = E.a_fills1 = 0;
    E.w_loads = E.w_used = E.d_count = 0;
    E.tmem = 0;
    if (E.on) {
        uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(E.red + 8);
        if (tid == 0) {
            for (int i = 0; i < kDgStages; ++i) {
                tc::mbar_init(&E.w_full[i], 1);
                tc::mbar_init(&E.w_empty[i], 1);
            }
            tc::mbar_init(&E.a_free[0], 1);
            tc::mbar_init(&E.a_free[1], 1);
            tc::mbar_init(E.d_done, 1);
            tc::fence_barrier_init();
        }
        if ((tid >> 5) == 0) tc::tmem_alloc(tmem_slot, kDgTmemCols);
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
        E.tmem = *tmem_slot;
    }

    const long long n_act = min((long long)P.stats[0], P.cap);
    const long long n_tiles = (n_act + kTile - 1) / kTile;

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long base = tile * kTile;
        const int nv = (int)min((long long)kTile, n_act - base);
        __syncthreads();
        if (tid < kTile) {
            float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
            int ray = -1;
            float c0 = 0.f, c1 = 0.f, c2 = 0.f;
            if (tid < nv) {
                p = P.rec_pos[base + tid];
                ray = P.rec_ray[base + tid];
                if (P.do_rgb) {
                    const float4 c = *reinterpret_cast<const float4*>(P.rec_rgb + (base + tid) * 4);
                    c0 = c.x, c1 = c.y, c2 = c.z;
                }
            }
            sm.pos[tid] = p;
            sm.ray[tid] = ray;
            sm.dir[tid] = c0;
            sm.dir[kTile + tid] = c1;
            sm.dir[2 * kTile + tid] = c2;
        }
        __syncthreads();
        const float* sa = P.stash_a + (size_t)tile * P.lay.a_rows * kTile;
        float* sz = P.stash_z + (size_t)tile * P.lay.z_rows * kTile;

        if (P.do_sem) {
            // dOut_c = w * g_sem_raw[ray][c]; softmax backward with the saved probabilities
            const int rows = (P.n_cls + 63) & ~63;
            if (tid < kTile) {
                const int ray = sm.ray[tid];
                const float w = sm.pos[tid].w;
                const float* g = P.g_ray + (int64_t)max(ray, 0) * P.ray_stride + P.off_sem;
                if (P.softmax) {
                    const float* pr = sa + (size_t)P.lay.prob_off * kTile;      // block of n_cls rows (store_rows in the forward)
                    float dot = 0.0f;
                    for (int c = 0; c < P.n_cls; ++c) dot += (ray >= 0 ? w * g[c] : 0.0f) * pr[stash_idx(P.n_cls, c, tid)];
                    for (int c = 0; c < P.n_cls; ++c) {
                        const float d = ray >= 0 ? w * g[c] : 0.0f;
                        sm.act[(size_t)c * kTile + tid] = pr[stash_idx(P.n_cls, c, tid)] * (d - dot);
                    }
                } else {
                    for (int c = 0; c < P.n_cls; ++c) sm.act[(size_t)c * kTile + tid] = ray >= 0 ? w * g[c] : 0.0f;
                }
                for (int c = P.n_cls; c < rows; ++c) sm.act[(size_t)c * kTile + tid] = 0.0f;
            }
            __syncthreads();
            mlp_backward(sm, E, P.sem, P.g_sem_mlp, sa, sz, P.lay.a_off[0], P.lay.z_off[0], P.semg.comps != 0);
            if (P.semg.comps)
                grid_head_backward(sm, P.semg, P.semg_basis_dgrad, P.semg_dim, sz + (size_t)P.lay.z_off[5][0] * kTile,
                                   P.g_semg_plane, P.g_semg_line, nv);
        }
        if (P.do_ins) {
            for (int net = 0; net < (P.slow_fast ? 2 : 1); ++net) {
                const int rows = (P.d_ins + 63) & ~63;
                if (tid < kTile) {
                    const int ray = sm.ray[tid];
                    const float w = sm.pos[tid].w;
                    const float* g = P.g_ray + (int64_t)max(ray, 0) * P.ray_stride + P.off_ins + net * P.d_ins;
                    for (int c = 0; c < P.d_ins; ++c) sm.act[(size_t)c * kTile + tid] = ray >= 0 ? w * g[c] : 0.0f;
                    for (int c = P.d_ins; c < rows; ++c) sm.act[(size_t)c * kTile + tid] = 0.0f;
                }
                __syncthreads();
                mlp_backward(sm, E, net == 0 ? P.insf : P.inss, net == 0 ? P.g_insf_mlp : P.g_inss_mlp, sa, sz,
                             P.lay.a_off[1 + net], P.lay.z_off[1 + net], P.insg.comps != 0);
                if (P.insg.comps) {
                    // the fast and the slow net read the same basis feature: its gradient is their sum.  The fast net's
                    // share is parked in the basis Z-stash rows (same thread <-> element mapping as store_rows) and added
                    // back after the slow net's backward.
                    float* zrow = sz + (size_t)P.lay.z_off[6][0] * kTile;
                    if (P.slow_fast && net == 0) {
                        store_rows(sm, zrow, 64);
                        __syncthreads();
                        continue;
                    }
                    if (P.slow_fast) {
                        const float4* z4 = reinterpret_cast<const float4*>(zrow);
                        float4* d4 = reinterpret_cast<float4*>(sm.act);
                        for (int i = tid; i < 64 * (kTile / 4); i += kThreads) {
                            const float4 z = z4[stash_idx4(64, i >> 5, i & 31)];
                            float4 d = d4[i];
                            d.x += z.x, d.y += z.y, d.z += z.z, d.w += z.w;
                            d4[i] = d;
                        }
                        __syncthreads();
                    }
                    grid_head_backward(sm, P.insg, P.insg_basis_dgrad, P.insg_dim, zrow, P.g_insg_plane, P.g_insg_line, nv);
                }
            }
        }
        if (P.do_rgb) {
            if (tid < kTile) {
                const int ray = sm.ray[tid];
                const float w = sm.pos[tid].w;
                const float* g = P.g_ray + (int64_t)max(ray, 0) * P.ray_stride;
                float gw = 0.0f;
                for (int k = 0; k < 3; ++k) {
                    const float c = sm.dir[k * kTile + tid];
                    const float gk = ray >= 0 ? g[k] : 0.0f;
                    gw += gk * c;
                    sm.act[(size_t)k * kTile + tid] = w * gk * c * (1.0f - c);   // through the sigmoid
                }
                for (int c = 3; c < 64; ++c) sm.act[(size_t)c * kTile + tid] = 0.0f;
                if (ray >= 0) P.g_w[(int64_t)ray * P.S + P.rec_idx[base + tid]] = gw;
            }
            __syncthreads();
            mlp_backward(sm, E, P.rgb, P.g_rgb_mlp, sa, sz, P.lay.a_off[3], P.lay.z_off[3], true);
            // positional-encoding backward: dfeat_a = dIn[a] + sum_j 2^j (cos_aj * dIn[sin_aj] - sin_aj * dIn[cos_aj])
            const int A = P.dim_app, pf = P.pe_feat;
            const int o_sf = A + 3, o_cf = o_sf + A * pf;
            const float* in = sa + (size_t)P.lay.a_off[3][0] * kTile;
            const int in_rows = (P.rgb.dims[0] + 15) & ~15;      // rows of that stash block
            // dfeat goes to the spare rows [192, 192+A) first (the MLP input occupies rows < 192), then down to [0, A)
            float* spare = sm.act + (size_t)192 * kTile;
            for (int idx = tid; idx < A * kTile; idx += kThreads) {
                const int a = idx / kTile, m = idx - a * kTile;
                float v = sm.act[(size_t)a * kTile + m];
                for (int j = 0; j < pf; ++j) {
                    const int rs = o_sf + a * pf + j, rc = o_cf + a * pf + j;
                    const float sv = in[stash_idx(in_rows, rs, m)], cv = in[stash_idx(in_rows, rc, m)];
                    v += (float)(1 << j) * (cv * sm.act[(size_t)rs * kTile + m] - sv * sm.act[(size_t)rc * kTile + m]);
                }
                spare[idx] = v;
            }
            __syncthreads();
            for (int idx = tid; idx < 64 * kTile; idx += kThreads) sm.act[idx] = idx < A * kTile ? spare[idx] : 0.0f;
            __syncthreads();
            // basis layer: dZ = dfeat (64 rows, zero padded) -> Z-stash, then dprod = basis^T dfeat
            store_rows(sm, sz + (size_t)P.lay.z_off[4][0] * kTile, 64);
            if (E.on && P.basis_dg16) {
                const int n_in = 3 * P.app.comps;
                run_dgrad_tc(sm, E, P.basis_dg16, A, n_in, n_in <= 64 ? 64 : (n_in <= 128 ? 128 : 256));
            } else {
                run_dgrad(sm, P.basis_dgrad, A, 3 * P.app.comps);
            }
            scatter_factors<NV>(sm, P.app, P.g_app_plane, P.g_app_line, nv);
        }
    }
    if (E.on) {
        tc::fence_before_sync();
        __syncthreads();
        if ((tid >> 5) == 0) tc::tmem_dealloc(E.tmem, kDgTmemCols);
    }
}

// -------------------------------------------------------------------------------------------------
// wgrad: dWt[k][n] += sum_tiles sum_m A[t][k][m] * Z[t][n][m].   64x64 output block per CTA, 4x4 per thread,
// record dimension streamed in 32-wide slabs through a cp.async double buffer (row stride 36: 16 B aligned, conflict-free).
// -------------------------------------------------------------------------------------------------
struct WgradParams {
    const float* a;      // stash_a + a_off*kTile
    const float* z;      // stash_z + z_off*kTile
    long long a_tile_stride, z_tile_stride;   // floats
    int K, N;            // valid rows (k_pad(in), n_pad(out)); output pitch = N
    float* out;          // packed dWt [K][N]
    const unsigned long long* stats;
    long long cap;
    int splits;
};

constexpr int kWgSlab = 32;
constexpr int kWgPitch = kWgSlab + 4;

__global__ void __launch_bounds__(256) wgrad_kernel(const __grid_constant__ WgradParams P) {
    __shared__ __align__(16) float sA[2][64 * kWgPitch];
    __shared__ __align__(16) float sZ[2][64 * kWgPitch];
    const long long n_act = min((long long)P.stats[0], P.cap);
    const long long n_tiles = (n_act + kTile - 1) / kTile;
    const int kb = blockIdx.x * 64, nb = blockIdx.y * 64, split = blockIdx.z;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

    const long long n_slabs = ((n_tiles - split + P.splits - 1) / P.splits) * (kTile / kWgSlab);   // slabs this CTA walks
    auto issue = [&](long long s) {
        const long long tile = split + (s / (kTile / kWgSlab)) * P.splits;
        const int m0 = (int)(s % (kTile / kWgSlab)) * kWgSlab;
        const int buf = (int)(s & 1);
        const float* ga = P.a + tile * P.a_tile_stride;      // stash blocks of K / N rows (stash_idx layout)
        const float* gz = P.z + tile * P.z_tile_stride;
        // 64 rows x 8 chunks of 16 B per operand = 512 chunks; 256 threads -> 2 each per operand
        for (int c = tid; c < 64 * (kWgSlab / 4); c += 256) {
            const int r = c / (kWgSlab / 4), q = c % (kWgSlab / 4);
            if (kb + r < P.K) cp_async16(&sA[buf][r * kWgPitch + q * 4], ga + stash_idx(P.K, kb + r, m0 + q * 4));
            if (nb + r < P.N) cp_async16(&sZ[buf][r * kWgPitch + q * 4], gz + stash_idx(P.N, nb + r, m0 + q * 4));
        }
        cp_async_commit();
    };
    // rows beyond K / N are never loaded: zero them once so they contribute nothing
    for (int i = tid; i < 2 * 64 * kWgPitch; i += 256) {
        (&sA[0][0])[i] = 0.0f;
        (&sZ[0][0])[i] = 0.0f;
    }
    __syncthreads();
    if (n_slabs > 0) issue(0);
    for (long long s = 0; s < n_slabs; ++s) {
        if (s + 1 < n_slabs) {
            issue(s + 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float* a = sA[s & 1];
        const float* z = sZ[s & 1];
#pragma unroll
        for (int mm = 0; mm < kWgSlab; mm += 4) {
            float4 av[4], zv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = *reinterpret_cast<const float4*>(a + (ty + 16 * i) * kWgPitch + mm);
#pragma unroll
            for (int j = 0; j < 4; ++j) zv[j] = *reinterpret_cast<const float4*>(z + (tx + 16 * j) * kWgPitch + mm);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[i][j] = fmaf(av[i].x, zv[j].x, acc[i][j]);
                    acc[i][j] = fmaf(av[i].y, zv[j].y, acc[i][j]);
                    acc[i][j] = fmaf(av[i].z, zv[j].z, acc[i][j]);
                    acc[i][j] = fmaf(av[i].w, zv[j].w, acc[i][j]);
                }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = kb + ty + 16 * i, n = nb + tx + 16 * j;
            if (k < P.K && n < P.N && acc[i][j] != 0.0f) atomicAdd(P.out + (size_t)k * P.N + n, acc[i][j]);
        }
}

// -------------------------------------------------------------------------------------------------
// per-ray epilogue backward (renderer:160-167): g_rgb -> g_rgb_raw (+ g_opacity), g_sem -> g_sem_raw, g_ins copy
// -------------------------------------------------------------------------------------------------
struct RayBwdParams {
    int64_t n_rays;
    int n_cls, d_all, softmax, add_bg, stride, off_sem, off_ins;
    const float* opacity;
    const float* rgb_raw;
    const float* sem_raw;
    const float* g_rgb;
    const float* g_sem;
    const float* g_ins;
    float* g_ray;
};

__global__ void __launch_bounds__(256) ray_backward_kernel(const __grid_constant__ RayBwdParams P) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P.n_rays) return;
    float* o = P.g_ray + r * P.stride;
    float g_opa = 0.0f;
    for (int k = 0; k < 3; ++k) {
        float g = 0.0f;
        if (P.g_rgb) {
            float v = P.rgb_raw[r * 3 + k];
            if (P.add_bg) v = v + (1.0f - P.opacity[r]);
            g = (v >= 0.0f && v <= 1.0f) ? P.g_rgb[r * 3 + k] : 0.0f;   // clamp passes gradient on the closed interval
            if (P.add_bg) g_opa -= g;
        }
        o[k] = g;
    }
    o[P.stride - 1] = g_opa;
    if (P.g_sem) {
        const float* s = P.sem_raw + r * P.n_cls;
        const float* g = P.g_sem + r * P.n_cls;
        if (P.softmax) {
            float tot = 0.0f;
            for (int c = 0; c < P.n_cls; ++c) tot += s[c];
            tot += 1e-8f;
            float dot = 0.0f;   // sum_j (dL/dp_j) * s_j
            for (int c = 0; c < P.n_cls; ++c) dot += g[c] / (s[c] / tot + 1e-8f) * s[c];
            for (int c = 0; c < P.n_cls; ++c) o[P.off_sem + c] = g[c] / (s[c] / tot + 1e-8f) / tot - dot / (tot * tot);
        } else {
            for (int c = 0; c < P.n_cls; ++c) o[P.off_sem + c] = g[c];
        }
    } else {
        for (int c = 0; c < P.n_cls; ++c) o[P.off_sem + c] = 0.0f;
    }
    for (int c = 0; c < P.d_all; ++c) o[P.off_ins + c] = P.g_ins ? P.g_ins[r * P.d_all + c] : 0.0f;
}

}  // namespace

int launch_heads_forward(const clift_render_cfg* cfg, const clift_field* field, const float* rays, const Workspace& ws,
                         int64_t cap, int64_t n_rays, float* rgb_raw, float* sem_raw, float* ins, const StashLayout* lay,
                         cudaStream_t stream) {
    HeadsParams P;
    P.rec_pos = ws.rec_pos;
    P.rec_ray = ws.rec_ray;
    P.stats = reinterpret_cast<const unsigned long long*>(ws.stats);
    P.cap = cap;
    P.rays = rays;
    P.app = make_factors(field, true);
    P.basis_wt = field->basis;
    P.dim_app = field->dim_appearance;
    P.pe_view = field->pe_view;
    P.pe_feat = field->pe_feat;
    P.pe_sem = field->pe_sem;
    P.pe_ins = field->pe_ins;
    P.semg = make_grid_factors(field, field->semantic_grid);
    P.insg = make_grid_factors(field, field->instance_grid);
    P.semg_basis = field->semantic_grid.basis;
    P.insg_basis = field->instance_grid.basis;
    P.semg_dim = field->semantic_grid.dim;
    P.insg_dim = field->instance_grid.dim;
    P.rgb = field->rgb;
    P.sem = field->semantic;
    P.insf = field->instance_fast;
    P.inss = field->instance_slow;
    P.n_cls = field->num_classes;
    P.d_ins = field->dim_instance;
    P.slow_fast = field->slow_fast;
    P.softmax = cfg->semantic_softmax;
    P.heads = cfg->heads;
    if (!rgb_raw) P.heads &= ~CLIFT_HEAD_RGB;
    if (!sem_raw) P.heads &= ~CLIFT_HEAD_SEMANTIC;
    if (!ins) P.heads &= ~CLIFT_HEAD_INSTANCE;
    P.rgb_raw = rgb_raw;
    P.sem_raw = sem_raw;
    P.ins = ins;
    P.rec_rgb = lay ? ws.rec_rgb : nullptr;
    P.stash_a = lay ? ws.stash_a : nullptr;
    if (lay) P.lay = *lay;
    if (P.heads == 0 || n_rays <= 0) return CLIFT_OK;
    const int grid = sm_count();
#define CLIFT_HEADS_CASE(NV)                                                                                          \
    case NV: {                                                                                                        \
        CLIFT_CUDA(cudaFuncSetAttribute(heads_forward_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                        (int)kSmemBytes));                                                            \
        heads_forward_kernel<NV><<<grid, kThreads, kSmemBytes, stream>>>(P);                                          \
        break;                                                                                                        \
    }
    switch (P.app.comps / 16) {
        CLIFT_HEADS_CASE(1)
        CLIFT_HEADS_CASE(2)
        CLIFT_HEADS_CASE(3)
        CLIFT_HEADS_CASE(4)
        default:
            set_error("launch_heads_forward: appearance_comps %d not in {16,32,48,64}", P.app.comps);
            return CLIFT_ERR_UNSUPPORTED;
    }
#undef CLIFT_HEADS_CASE
    CLIFT_AFTER_LAUNCH("heads_forward_kernel");
    return CLIFT_OK;
}

}  // namespace clift

namespace clift {

// weight gradient of layer l of stack `id`: queued for the tensor-core kernel (wgrad_tc.cu) when its shape qualifies, else
// one FP32-FMA split-K launch (the K = 16 xyz input layers)
struct WgradQueue {
    WgradTcItem items[kWgradTcMaxLayers];
    int n = 0;
    bool use_tc = true;
};

static int launch_wgrad(const Workspace& ws, const StashLayout& lay, int id, int l, const clift_mlp* m, float* out, int64_t cap,
                        cudaStream_t stream, WgradQueue* q = nullptr) {
    if (!out) return CLIFT_OK;
    const int K = k_pad(m->dims[l]), N = n_pad(m->dims[l + 1]);
    if (q && q->use_tc && q->n < kWgradTcMaxLayers && wgrad_tc_eligible(K, N)) {
        WgradTcItem& it = q->items[q->n++];
        it.a_row = lay.a_off[id][l];
        it.z_row = lay.z_off[id][l];
        it.K = K;
        it.N = N;
        it.out = out;
        return CLIFT_OK;
    }
    WgradParams W;
    W.a = ws.stash_a + (size_t)lay.a_off[id][l] * kTile;
    W.z = ws.stash_z + (size_t)lay.z_off[id][l] * kTile;
    W.a_tile_stride = (long long)lay.a_rows * kTile;
    W.z_tile_stride = (long long)lay.z_rows * kTile;
    W.K = K;
    W.N = N;
    W.out = out;
    W.stats = reinterpret_cast<const unsigned long long*>(ws.stats);
    W.cap = cap;
    const int bk = (int)ceil_div(W.K, 64), bn = (int)ceil_div(W.N, 64);
    W.splits = std::max(1, (2 * sm_count()) / (bk * bn));
    dim3 grid(bk, bn, W.splits);
    wgrad_kernel<<<grid, 256, 0, stream>>>(W);
    CLIFT_AFTER_LAUNCH("wgrad_kernel");
    return CLIFT_OK;
}

// Per-ray epilogue backward -> heads backward (dgrad, stash dZ, bias grads, dL/dw, appearance scatter) -> wgrad per layer.
int launch_heads_backward(const clift_render_cfg* cfg, const clift_field* field, const Workspace& ws, const StashLayout& lay,
                          int64_t cap, int64_t n_rays, int add_bg, const clift_render_out* saved, const float* g_rgb,
                          const float* g_sem, const float* g_ins, const clift_field_grad* grad, int* ray_stride_out,
                          cudaStream_t stream) {
    const int C = field->num_classes, DI = field->dim_instance * (field->slow_fast ? 2 : 1);
    const int stride = 3 + C + DI + 1;
    *ray_stride_out = stride;
    const int heads = cfg->heads;
    const bool do_rgb = (heads & CLIFT_HEAD_RGB) && g_rgb;
    const bool do_sem = (heads & CLIFT_HEAD_SEMANTIC) && g_sem && grad->semantic.wt[0];
    const bool do_ins = (heads & CLIFT_HEAD_INSTANCE) && g_ins && grad->instance_fast.wt[0];
    {
        RayBwdParams R;
        R.n_rays = n_rays;
        R.n_cls = C;
        R.d_all = DI;
        R.softmax = cfg->semantic_softmax;
        R.add_bg = add_bg;
        R.stride = stride;
        R.off_sem = 3;
        R.off_ins = 3 + C;
        R.opacity = saved->opacity;
        R.rgb_raw = saved->rgb_raw;
        R.sem_raw = saved->semantic_raw;
        R.g_rgb = do_rgb ? g_rgb : nullptr;
        R.g_sem = do_sem ? g_sem : nullptr;
        R.g_ins = do_ins ? g_ins : nullptr;
        R.g_ray = ws.g_ray;
        ray_backward_kernel<<<(unsigned)ceil_div(n_rays, 256), 256, 0, stream>>>(R);
        CLIFT_AFTER_LAUNCH("ray_backward_kernel");
    }
    if (!(do_rgb || do_sem || do_ins)) return CLIFT_OK;
    HeadsBwdParams P;
    memset(&P, 0, sizeof(P));
    P.rec_pos = ws.rec_pos;
    P.rec_ray = ws.rec_ray;
    P.rec_idx = ws.rec_idx;
    P.rec_rgb = ws.rec_rgb;
    P.stats = reinterpret_cast<const unsigned long long*>(ws.stats);
    P.cap = cap;
    P.S = cfg->n_samples;
    P.app = make_factors(field, true);
    for (int m = 0; m < 3; ++m) {
        P.g_app_plane[m] = grad->appearance_plane[m];
        P.g_app_line[m] = grad->appearance_line[m];
    }
    P.basis_dgrad = field->basis_dgrad;
    P.basis_dg16 = field->basis_dg16;
    {   // development switch: CLIFT_DGRAD_FMA=1 keeps every data gradient on the FP32-FMA tile GEMM
        const char* e = getenv("CLIFT_DGRAD_FMA");
        P.use_tc = !(e && atoi(e) != 0);
        const char* m = getenv("CLIFT_BWD_MASK_PASS");     // A/B of the mask fused into the emit pass (default) vs its own pass
        P.mask_pass = (m && atoi(m) != 0) ? 1 : 0;
    }
    P.dim_app = field->dim_appearance;
    P.pe_view = field->pe_view;
    P.pe_feat = field->pe_feat;
    P.semg = make_grid_factors(field, field->semantic_grid);
    P.insg = make_grid_factors(field, field->instance_grid);
    if (!do_sem) P.semg.comps = 0;
    if (!do_ins) P.insg.comps = 0;
    for (int m = 0; m < 3; ++m) {
        P.g_semg_plane[m] = grad->semantic_grid.plane[m];
        P.g_semg_line[m] = grad->semantic_grid.line[m];
        P.g_insg_plane[m] = grad->instance_grid.plane[m];
        P.g_insg_line[m] = grad->instance_grid.line[m];
    }
    P.semg_basis_dgrad = field->semantic_grid.basis_dgrad;
    P.insg_basis_dgrad = field->instance_grid.basis_dgrad;
    P.semg_dim = field->semantic_grid.dim;
    P.insg_dim = field->instance_grid.dim;
    for (int h = 0; h < 2; ++h) {
        const FactorParams& gf = h == 0 ? P.semg : P.insg;
        if (!gf.comps) continue;
        const clift_grid_head_grad& gg = h == 0 ? grad->semantic_grid : grad->instance_grid;
        const clift_grid_head& gh = h == 0 ? field->semantic_grid : field->instance_grid;
        for (int m = 0; m < 3; ++m)
            if (!gg.plane[m] || !gg.line[m]) {
                set_error("clift_render_backward: grid-mode %s head without factor gradient buffers", h == 0 ? "semantic" : "instance");
                return CLIFT_ERR_ARG;
            }
        if (!gg.basis || !gh.basis_dgrad) {
            set_error("clift_render_backward: grid-mode %s head without basis dgrad operand / gradient buffer",
                      h == 0 ? "semantic" : "instance");
            return CLIFT_ERR_ARG;
        }
    }
    P.rgb = field->rgb;
    P.sem = field->semantic;
    P.insf = field->instance_fast;
    P.inss = field->instance_slow;
    P.g_rgb_mlp = grad->rgb;
    P.g_sem_mlp = grad->semantic;
    P.g_insf_mlp = grad->instance_fast;
    P.g_inss_mlp = grad->instance_slow;
    P.n_cls = C;
    P.d_ins = field->dim_instance;
    P.slow_fast = field->slow_fast;
    P.softmax = cfg->semantic_softmax;
    P.do_rgb = do_rgb;
    P.do_sem = do_sem;
    P.do_ins = do_ins;
    P.g_ray = ws.g_ray;
    P.ray_stride = stride;
    P.off_sem = 3;
    P.off_ins = 3 + C;
    P.g_w = ws.g_w;
    P.stash_a = ws.stash_a;
    P.stash_z = ws.stash_z;
    P.lay = lay;
    if (do_rgb) {
        for (int m = 0; m < 3; ++m)
            if (!P.g_app_plane[m] || !P.g_app_line[m]) {
                set_error("clift_render_backward: rgb gradient requested without appearance factor gradient buffers");
                return CLIFT_ERR_ARG;
            }
        if (!field->basis_dgrad || !grad->basis) {
            set_error("clift_render_backward: rgb gradient requested without basis dgrad operand / gradient buffer");
            return CLIFT_ERR_ARG;
        }
        if (k_pad(field->rgb.dims[0]) > 192 || field->dim_appearance > 64) {
            set_error("clift_render_backward: rgb head input wider than 192 is not supported in training");
            return CLIFT_ERR_UNSUPPORTED;
        }
        CLIFT_CUDA(cudaMemsetAsync(ws.g_w, 0, n_rays * (int64_t)cfg->n_samples * sizeof(float), stream));
    }
    auto need_dgrad = [&](const clift_mlp& m, bool from0) {
        for (int l = from0 ? 0 : 1; l < m.n_layers; ++l)
            if (!m.w_dgrad[l]) return false;
        return true;
    };
    const bool sg = P.semg.comps != 0, ig = P.insg.comps != 0;
    if ((do_rgb && !need_dgrad(field->rgb, true)) || (do_sem && !need_dgrad(field->semantic, sg)) ||
        (do_ins && (!need_dgrad(field->instance_fast, ig) || (field->slow_fast && !need_dgrad(field->instance_slow, ig))))) {
        set_error("clift_render_backward: field lacks the w_dgrad operands (pack with clift_pack_linear_dgrad)");
        return CLIFT_ERR_ARG;
    }
    const int grid = sm_count();
#define CLIFT_HEADS_BWD_CASE(NV)                                                                                       \
    case NV: {                                                                                                         \
        CLIFT_CUDA(cudaFuncSetAttribute(heads_backward_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                        (int)(kSmemBytes + kDgExtraBytes)));                                           \
        heads_backward_kernel<NV><<<grid, kThreads, kSmemBytes + kDgExtraBytes, stream>>>(P);                          \
        break;                                                                                                         \
    }
    switch (P.app.comps / 16) {
        CLIFT_HEADS_BWD_CASE(1)
        CLIFT_HEADS_BWD_CASE(2)
        CLIFT_HEADS_BWD_CASE(3)
        CLIFT_HEADS_BWD_CASE(4)
        default:
            set_error("launch_heads_backward: appearance_comps %d not in {16,32,48,64}", P.app.comps);
            return CLIFT_ERR_UNSUPPORTED;
    }
#undef CLIFT_HEADS_BWD_CASE
    CLIFT_AFTER_LAUNCH("heads_backward_kernel");
    // weight gradients: one tensor-core launch for every layer with K >= 64 (CLIFT_WGRAD_FMA=1: the FP32-FMA kernel for all)
    int rc;
    clift_mlp tmp;
    WgradQueue queue;
    {
        const char* e = getenv("CLIFT_WGRAD_FMA");
        queue.use_tc = !(e && atoi(e) != 0);
    }
    WgradQueue* q = &queue;
    if (do_sem) {
        for (int l = 0; l < field->semantic.n_layers; ++l)
            if ((rc = launch_wgrad(ws, lay, 0, l, &field->semantic, grad->semantic.wt[l], cap, stream, q))) return rc;
        if (sg && (rc = launch_wgrad(ws, lay, 5, 0, field_mlp(field, 5, &tmp), grad->semantic_grid.basis, cap, stream, q))) return rc;
    }
    if (do_ins && ig)
        if ((rc = launch_wgrad(ws, lay, 6, 0, field_mlp(field, 6, &tmp), grad->instance_grid.basis, cap, stream, q))) return rc;
    if (do_ins) {
        for (int l = 0; l < field->instance_fast.n_layers; ++l)
            if ((rc = launch_wgrad(ws, lay, 1, l, &field->instance_fast, grad->instance_fast.wt[l], cap, stream, q))) return rc;
        if (field->slow_fast)
            for (int l = 0; l < field->instance_slow.n_layers; ++l)
                if ((rc = launch_wgrad(ws, lay, 2, l, &field->instance_slow, grad->instance_slow.wt[l], cap, stream, q))) return rc;
    }
    if (do_rgb) {
        for (int l = 0; l < field->rgb.n_layers; ++l)
            if ((rc = launch_wgrad(ws, lay, 3, l, &field->rgb, grad->rgb.wt[l], cap, stream, q))) return rc;
        if ((rc = launch_wgrad(ws, lay, 4, 0, field_mlp(field, 4, &tmp), grad->basis, cap, stream, q))) return rc;
    }
    return launch_wgrad_tc(ws, lay, queue.items, queue.n, cap, stream);
}

}  // namespace clift
