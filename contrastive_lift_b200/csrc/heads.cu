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
#include "launchers.h"

namespace clift {

namespace {

constexpr int kTile = CLIFT_TILE;
constexpr int kThreads = 256;
constexpr int kActRows = CLIFT_MAX_WIDTH;
constexpr int kSlabRows = 16;

struct HeadsParams {
    const float4* rec_pos;
    const int32_t* rec_ray;
    const unsigned long long* stats;
    long long cap;
    const float* rays;
    FactorParams app;
    const float* basis_wt;
    int dim_app, pe_view, pe_feat, pe_sem, pe_ins;
    clift_mlp rgb, sem, insf, inss;
    int n_cls, d_ins, slow_fast, softmax, heads;
    float* rgb_raw;
    float* sem_raw;
    float* ins;
    float* rec_rgb;
};

struct Smem {
    float* act;      // [kActRows][kTile]
    float* wslab;    // [2][kSlabRows][256]
    float4* pos;     // [kTile]
    int* ray;        // [kTile]
    int* runs;       // [kTile + 1]
    float* dir;      // [3][kTile]
};

constexpr size_t kSmemBytes = (size_t)kActRows * kTile * 4 + 2 * kSlabRows * 256 * 4 + kTile * 16 + kTile * 4 +
                              (kTile + 4) * 4 + 3 * kTile * 4;

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

__device__ __forceinline__ void run_mlp(const Smem& sm, const clift_mlp& mlp) {
    for (int l = 0; l < mlp.n_layers; ++l)
        run_layer(sm, mlp.wt[l], mlp.bias[l], mlp.dims[l], mlp.dims[l + 1], l + 1 < mlp.n_layers);
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

        if (P.heads & CLIFT_HEAD_SEMANTIC) {
            build_xyz_input(sm, P.pe_sem);
            run_mlp(sm, P.sem);
            if (tid < kTile) {
                const float w = sm.pos[tid].w;
                if (P.softmax) {
                    float mx = -INFINITY;
                    for (int c = 0; c < P.n_cls; ++c) mx = fmaxf(mx, sm.act[(size_t)c * kTile + tid]);
                    float tot = 0.0f;
                    for (int c = 0; c < P.n_cls; ++c) {
                        const float e = expf(sm.act[(size_t)c * kTile + tid] - mx);
                        sm.act[(size_t)c * kTile + tid] = e;
                        tot += e;
                    }
                    const float sc = w / tot;
                    for (int c = 0; c < P.n_cls; ++c) sm.act[(size_t)c * kTile + tid] *= sc;
                } else {
                    for (int c = 0; c < P.n_cls; ++c) sm.act[(size_t)c * kTile + tid] *= w;
                }
            }
            __syncthreads();
            reduce_runs(sm, n_runs, P.n_cls, P.sem_raw, P.n_cls, 0);
        }
        if (P.heads & CLIFT_HEAD_INSTANCE) {
            const int width = P.d_ins * (P.slow_fast ? 2 : 1);
            for (int net = 0; net < (P.slow_fast ? 2 : 1); ++net) {
                build_xyz_input(sm, P.pe_ins);
                run_mlp(sm, net == 0 ? P.insf : P.inss);
                for (int idx = tid; idx < P.d_ins * kTile; idx += kThreads) sm.act[idx] *= sm.pos[idx % kTile].w;
                __syncthreads();
                reduce_runs(sm, n_runs, P.d_ins, P.ins, width, net * P.d_ins);
            }
        }
        if (P.heads & CLIFT_HEAD_RGB) {
            gather_appearance<NV>(sm, P.app);
            run_layer(sm, P.basis_wt, nullptr, 3 * P.app.comps, P.dim_app, false);
            build_rgb_input(sm, P.dim_app, P.pe_feat, P.pe_view);
            run_mlp(sm, P.rgb);
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

}  // namespace

int launch_heads_forward(const clift_render_cfg* cfg, const clift_field* field, const float* rays, const Workspace& ws,
                         int64_t cap, int64_t n_rays, float* rgb_raw, float* sem_raw, float* ins, bool save_rgb,
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
    P.rec_rgb = save_rgb ? ws.rec_rgb : nullptr;
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
