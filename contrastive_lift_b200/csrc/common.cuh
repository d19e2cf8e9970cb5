// Internal helpers shared by the clift_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/clift_b200.h"

namespace clift {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int sm_count();

#define CLIFT_CHECK_ARG(cond, msg)                                    \
    do {                                                              \
        if (!(cond)) {                                                \
            ::clift::set_error("%s: %s", __func__, msg);              \
            return CLIFT_ERR_ARG;                                     \
        }                                                             \
    } while (0)

#define CLIFT_CHECK_SUPPORTED(cond, msg)                              \
    do {                                                              \
        if (!(cond)) {                                                \
            ::clift::set_error("%s: unsupported: %s", __func__, msg); \
            return CLIFT_ERR_UNSUPPORTED;                             \
        }                                                             \
    } while (0)

#define CLIFT_CUDA(expr)                                                                       \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            ::clift::set_error("%s: %s -> %s", __func__, #expr, cudaGetErrorString(_e));       \
            return CLIFT_ERR_CUDA;                                                             \
        }                                                                                      \
    } while (0)

#define CLIFT_AFTER_LAUNCH(name)                                                               \
    do {                                                                                       \
        ::clift::count_launch();                                                               \
        cudaError_t _e = cudaGetLastError();                                                   \
        if (_e != cudaSuccess) {                                                               \
            ::clift::set_error("%s: launch %s -> %s", __func__, name, cudaGetErrorString(_e)); \
            return CLIFT_ERR_CUDA;                                                             \
        }                                                                                      \
    } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }
static inline int k_pad(int k) { return (int)round_up(k, 16); }
static inline int n_pad(int n) { return (int)round_up(n, 64); }
static inline int dgrad_pad(int n) { return n <= 64 ? 64 : (n <= 128 ? 128 : 256); }

// tensoRF.py:61-62
__host__ __device__ __forceinline__ int mode_a(int m) { return m == 2 ? 1 : 0; }
__host__ __device__ __forceinline__ int mode_b(int m) { return m == 0 ? 1 : 2; }
__host__ __device__ __forceinline__ int mode_v(int m) { return 2 - m; }

// Per-tile activation stash of the training path: blocks of rows x CLIFT_TILE floats (one value per record of the tile),
// a tile's blocks back to back.  MLP ids below.
//   a_off[id][l] : first row of the block holding the INPUT of layer l (k_pad(dims[l]) rows, zero padded)
//   z_off[id][l] : first row of the block holding dL/d(pre-activation OUTPUT of layer l) (n_pad(dims[l+1]) rows, zero padded)
//   prob_off     : softmax probabilities of the semantic head (n_pad(C) rows reserved, C written)
// Block layout: R rows x 128 records are stored as 8 groups of 16 records, each group row-major [R][16 floats], so the
// tensor-core weight-gradient kernel (wgrad_tc.cu) streams one 16-record group of ALL rows as one contiguous R*64-byte
// piece (sequential DRAM reads instead of 64-byte pieces at a 512-byte stride).
// stash_idx = float offset of element (row, record m) inside a block; stash_idx4 = float4 offset of the q-th float4 of a row.
__host__ __device__ __forceinline__ size_t stash_idx(int R, int row, int m) {
    return ((size_t)(m >> 4) * R + row) * 16 + (m & 15);
}
__host__ __device__ __forceinline__ size_t stash_idx4(int R, int row, int q) {
    return ((size_t)(q >> 2) * R + row) * 4 + (q & 3);
}

// stack ids: 0 semantic mlp, 1 instance fast, 2 instance slow, 3 rgb mlp, 4 appearance basis,
//            5 semantic-grid basis, 6 instance-grid basis (grid-mode heads only)
constexpr int kStashIds = 7;
struct StashLayout {
    int a_off[kStashIds][CLIFT_MAX_LAYERS];
    int z_off[kStashIds][CLIFT_MAX_LAYERS];
    int prob_off;
    int a_rows, z_rows;   // rows per tile in the A / Z stash
};

inline const clift_mlp* field_mlp(const clift_field* f, int id, clift_mlp* basis_tmp) {
    switch (id) {
        case 0: return &f->semantic;
        case 1: return &f->instance_fast;
        case 2: return &f->instance_slow;
        case 3: return &f->rgb;
        case 4:
            memset(basis_tmp, 0, sizeof(*basis_tmp));
            basis_tmp->n_layers = 1;
            basis_tmp->dims[0] = 3 * f->appearance_comps;
            basis_tmp->dims[1] = f->dim_appearance;
            basis_tmp->wt[0] = f->basis;
            basis_tmp->w_dgrad[0] = f->basis_dgrad;
            return basis_tmp;
        default: {
            const clift_grid_head& g = id == 5 ? f->semantic_grid : f->instance_grid;
            memset(basis_tmp, 0, sizeof(*basis_tmp));
            basis_tmp->n_layers = 1;
            basis_tmp->dims[0] = 3 * g.comps;
            basis_tmp->dims[1] = g.dim;
            basis_tmp->wt[0] = g.basis;
            basis_tmp->w_dgrad[0] = g.basis_dgrad;
            return basis_tmp;
        }
    }
}

inline StashLayout make_stash_layout(const clift_field* f, int heads) {
    StashLayout L;
    memset(&L, 0, sizeof(L));
    int a = 0, z = 0;
    for (int id = 0; id < kStashIds; ++id) {
        const bool on = (id == 0 && (heads & CLIFT_HEAD_SEMANTIC)) || (id == 1 && (heads & CLIFT_HEAD_INSTANCE)) ||
                        (id == 2 && (heads & CLIFT_HEAD_INSTANCE) && f->slow_fast) ||
                        ((id == 3 || id == 4) && (heads & CLIFT_HEAD_RGB)) ||
                        (id == 5 && (heads & CLIFT_HEAD_SEMANTIC) && f->semantic_grid.comps > 0) ||
                        (id == 6 && (heads & CLIFT_HEAD_INSTANCE) && f->instance_grid.comps > 0);
        if (!on) continue;
        clift_mlp tmp;
        const clift_mlp* m = field_mlp(f, id, &tmp);
        for (int l = 0; l < m->n_layers; ++l) {
            L.a_off[id][l] = a;
            a += k_pad(m->dims[l]);
            L.z_off[id][l] = z;
            z += n_pad(m->dims[l + 1]);
        }
        if (id == 0) {
            L.prob_off = a;
            a += n_pad(f->num_classes);
        }
    }
    L.a_rows = a;
    L.z_rows = z;
    return L;
}

// ---------------------------------------------------------------------------------------
// Workspace carve-up (host side).  All regions 256-byte aligned.
// ---------------------------------------------------------------------------------------
struct Workspace {
    int32_t* stats;     // [16]: u64 view: 0 n_active, 1 n_inbox, 2 overflow flag, 3 n_tiles
    float* w_dense;     // [B*S] compositing weights (march -> fill, backward)
    int32_t* count;     // [B] active samples per ray
    int32_t* offset;    // [B+1] exclusive scan of count
    int32_t* bsum;      // [scan blocks + 1]
    float4* rec_pos;    // [cap] (x,y,z normalised, w)
    int32_t* rec_ray;   // [cap] ray index
    int32_t* rec_idx;   // [cap] sample index within the ray
    // ---- training only (save_for_backward) ----
    float* rec_rgb;     // [cap*4] per-record rgb (the weight gradient needs it)
    float* sigma_dense; // [B*S]
    float* trans_dense; // [B*S] transmittance T_i
    float* g_w;         // [B*S] dL/dw_i through the rgb map (heads backward -> march backward)
    float* g_ray;       // [B*(out_width+2)] per-ray upstream after the epilogue backward
    float* stash_a;     // [tiles * a_rows * CLIFT_TILE]
    float* stash_z;     // [tiles * z_rows * CLIFT_TILE]
    int64_t bytes;
};

static const int kScanBlock = 2048;

// `with_z`: the dL/dZ stash lives inside the workspace (save_for_backward = 1); false: the caller supplies it at backward
// time (save_for_backward = 2), so forwards that wait for their backward hold only the activation stash
inline Workspace carve_workspace(void* base, int64_t n_rays, int n_samples, int64_t cap, int out_width, bool save,
                                 const StashLayout* layout, bool with_z = true) {
    Workspace w;
    memset(&w, 0, sizeof(w));
    char* p = (char*)base;
    int64_t off = 0;
    auto take = [&](int64_t bytes) {
        char* r = p ? p + off : nullptr;
        off += round_up(bytes, 256);
        return r;
    };
    const int64_t tiles = ceil_div(cap, CLIFT_TILE);
    w.stats = (int32_t*)take(16 * sizeof(int32_t));
    w.w_dense = (float*)take(n_rays * n_samples * sizeof(float));
    w.count = (int32_t*)take(n_rays * sizeof(int32_t));
    w.offset = (int32_t*)take((n_rays + 1) * sizeof(int32_t));
    w.bsum = (int32_t*)take((ceil_div(n_rays, kScanBlock) + 2) * sizeof(int32_t));
    w.rec_pos = (float4*)take(tiles * CLIFT_TILE * sizeof(float4));
    w.rec_ray = (int32_t*)take(tiles * CLIFT_TILE * sizeof(int32_t));
    w.rec_idx = (int32_t*)take(tiles * CLIFT_TILE * sizeof(int32_t));
    if (save) {
        w.rec_rgb = (float*)take(tiles * CLIFT_TILE * 4 * sizeof(float));
        w.sigma_dense = (float*)take(n_rays * n_samples * sizeof(float));
        w.trans_dense = (float*)take(n_rays * n_samples * sizeof(float));
        w.g_w = (float*)take(n_rays * n_samples * sizeof(float));
        w.g_ray = (float*)take(n_rays * (int64_t)(out_width + 2) * sizeof(float));
        w.stash_a = (float*)take(tiles * (int64_t)layout->a_rows * CLIFT_TILE * sizeof(float));
        if (with_z) w.stash_z = (float*)take(tiles * (int64_t)layout->z_rows * CLIFT_TILE * sizeof(float));
    }
    w.bytes = off;
    return w;
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------
// Bit-exact sampling arithmetic (renderer:800-817, 633-634).  PyTorch evaluates every
// elementwise op with its own rounding, so nothing here may be contracted into an FMA:
// only __f*_rn intrinsics are used.
// ---------------------------------------------------------------------------------------
struct RayGeom {
    float o[3], d[3];
    float t_min;
    float jit;
    int has_jit;
};

__device__ __forceinline__ float ray_t_min(const float* o, const float* d, float near, float far,
                                           const float* amin, const float* amax) {
    float t = -INFINITY;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float vec = (d[k] == 0.0f) ? 1e-6f : d[k];
        float ra = __fdiv_rn(__fsub_rn(amax[k], o[k]), vec);
        float rb = __fdiv_rn(__fsub_rn(amin[k], o[k]), vec);
        t = fmaxf(t, fminf(ra, rb));
    }
    return fminf(fmaxf(t, near), far);
}

__device__ __forceinline__ float sample_t(const RayGeom& g, float step, int i) {
    float r = (float)i;
    if (g.has_jit) r = __fadd_rn(r, g.jit);
    return __fadd_rn(g.t_min, __fmul_rn(step, r));
}

// world point, in-box test, normalised coordinate
__device__ __forceinline__ bool sample_point(const RayGeom& g, float t, const float* amin, const float* amax,
                                             const float* inv, float* x) {
    bool in = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float p = __fadd_rn(g.o[k], __fmul_rn(g.d[k], t));
        in = in && !(amin[k] > p) && !(p > amax[k]);
        x[k] = __fsub_rn(__fmul_rn(__fsub_rn(p, amin[k]), inv[k]), 1.0f);
    }
    return in;
}

// ---------------------------------------------------------------------------------------
// VM lookup taps (tensoRF.py:108-134 via F.grid_sample bilinear / zeros / align_corners=True).
// Channel-last planes: one tap of one quad lane = one 16-byte load.
// ---------------------------------------------------------------------------------------
struct Tap2 {   // bilinear footprint on a (H,W) plane
    int x0, y0;
    float w00, w10, w01, w11;   // (x0,y0) (x0+1,y0) (x0,y0+1) (x0+1,y0+1), zero if out of bounds
};
struct Tap1 {
    int z0;
    float w0, w1;
};

__device__ __forceinline__ float unnormalize(float c, int size) {
    return __fmul_rn(__fmul_rn(__fadd_rn(c, 1.0f), 0.5f), (float)(size - 1));
}

__device__ __forceinline__ Tap2 make_tap2(float ca, float cb, int W, int H) {
    float fx = unnormalize(ca, W), fy = unnormalize(cb, H);
    float x0f = floorf(fx), y0f = floorf(fy);
    float wx1 = fx - x0f, wx0 = (x0f + 1.0f) - fx;
    float wy1 = fy - y0f, wy0 = (y0f + 1.0f) - fy;
    Tap2 t;
    t.x0 = (int)x0f;
    t.y0 = (int)y0f;
    bool xa = t.x0 >= 0 && t.x0 < W, xb = t.x0 + 1 >= 0 && t.x0 + 1 < W;
    bool ya = t.y0 >= 0 && t.y0 < H, yb = t.y0 + 1 >= 0 && t.y0 + 1 < H;
    t.w00 = (xa && ya) ? wx0 * wy0 : 0.0f;
    t.w10 = (xb && ya) ? wx1 * wy0 : 0.0f;
    t.w01 = (xa && yb) ? wx0 * wy1 : 0.0f;
    t.w11 = (xb && yb) ? wx1 * wy1 : 0.0f;
    return t;
}

__device__ __forceinline__ Tap1 make_tap1(float cv, int L) {
    float fz = unnormalize(cv, L);
    float z0f = floorf(fz);
    Tap1 t;
    t.z0 = (int)z0f;
    bool a = t.z0 >= 0 && t.z0 < L, b = t.z0 + 1 >= 0 && t.z0 + 1 < L;
    t.w0 = a ? (z0f + 1.0f) - fz : 0.0f;
    t.w1 = b ? fz - z0f : 0.0f;
    return t;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ void fma4(float4& acc, const float4& v, float w) {
    acc.x = fmaf(v.x, w, acc.x);
    acc.y = fmaf(v.y, w, acc.y);
    acc.z = fmaf(v.z, w, acc.z);
    acc.w = fmaf(v.w, w, acc.w);
}

// Bilinear plane read of 4 consecutive channels starting at `ch`.  Out-of-range taps carry weight 0
// (zeros padding); their address is never formed.
__device__ __forceinline__ float4 plane_tap(const float* __restrict__ plane, const Tap2& t, int W, int comps, int ch) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int64_t row0 = (int64_t)t.y0 * W, row1 = row0 + W;
    if (t.w00 != 0.0f) fma4(acc, ldg4(plane + (row0 + t.x0) * comps + ch), t.w00);
    if (t.w10 != 0.0f) fma4(acc, ldg4(plane + (row0 + t.x0 + 1) * comps + ch), t.w10);
    if (t.w01 != 0.0f) fma4(acc, ldg4(plane + (row1 + t.x0) * comps + ch), t.w01);
    if (t.w11 != 0.0f) fma4(acc, ldg4(plane + (row1 + t.x0 + 1) * comps + ch), t.w11);
    return acc;
}

// Linear line read (global or shared memory) of 4 consecutive channels.
__device__ __forceinline__ float4 line_tap(const float* line, const Tap1& t, int comps, int ch) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t.w0 != 0.0f) fma4(acc, *reinterpret_cast<const float4*>(line + (int64_t)t.z0 * comps + ch), t.w0);
    if (t.w1 != 0.0f) fma4(acc, *reinterpret_cast<const float4*>(line + (int64_t)(t.z0 + 1) * comps + ch), t.w1);
    return acc;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier + 1-D bulk TMA (cp.async.bulk -> SASS UBLKCP) ------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
#endif  // __CUDACC__

}  // namespace clift
