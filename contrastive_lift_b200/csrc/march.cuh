// Launch-parameter blocks shared between march.cu, heads.cu and api.cu.
#pragma once
#include "common.cuh"

namespace clift {

struct GeomParams {   // TensoRFRenderer constants, by value
    float amin[3], amax[3], inv[3];
    float step, scale, thres;
    int S;
};

struct FactorParams {   // one VM factor set (density / appearance), packed layout
    const float* plane[3];
    const float* line[3];
    int pw[3], ph[3], ll[3];   // plane width/height, line length per mode
    int comps;
};

struct MarchParams {
    GeomParams g;
    FactorParams f;
    float shift;
    int lines_in_smem;
    const float* rays;
    const float* jitter;
    int64_t n_rays;
    float* w_dense;     // [B,S]
    float* sigma_dense; // [B,S] or null (training)
    float* trans_dense; // [B,S] or null (training)
    int32_t* count;     // [B]
    float* opacity;     // [B]
    float* depth;       // [B]
    float* dist_ray;    // [B] or null
    float* points;      // [B,3] or null
    unsigned long long* stats;   // [8] u64: 0 n_active 1 n_inbox 2 overflow 3 n_tiles 4 group ticket (fused compaction)
    // fused compaction (inference; null scan_state = off): the march itself emits the ray-major active-sample records -
    // groups of 8 consecutive rays are claimed in ticket order and their record offsets come from a decoupled look-back
    // over `scan_state` (one u64 per group: flag << 62 | count), so no dense [B,S] weight array is written or re-read
    unsigned long long* scan_state;   // [ceil(B / 8)], zeroed by the caller
    float4* rec_pos;
    int32_t* rec_ray;
    int32_t* rec_idx;
    long long cap;
};

inline GeomParams make_geom(const clift_render_cfg* c) {
    GeomParams g;
    for (int k = 0; k < 3; ++k) {
        g.amin[k] = c->aabb_min[k];
        g.amax[k] = c->aabb_max[k];
        g.inv[k] = c->inv_extent[k];
    }
    g.step = c->step_size;
    g.scale = c->distance_scale;
    g.thres = c->weight_thres;
    g.S = c->n_samples;
    return g;
}

inline FactorParams make_factors(const clift_field* f, bool appearance) {
    FactorParams p;
    for (int m = 0; m < 3; ++m) {
        p.plane[m] = appearance ? f->appearance_plane[m] : f->density_plane[m];
        p.line[m] = appearance ? f->appearance_line[m] : f->density_line[m];
        p.pw[m] = f->grid[mode_a(m)];
        p.ph[m] = f->grid[mode_b(m)];
        p.ll[m] = f->grid[mode_v(m)];
    }
    p.comps = appearance ? f->appearance_comps : f->density_comps;
    return p;
}

// factor set of a grid-mode semantic / instance head (comps == 0: the head is in MLP mode)
inline FactorParams make_grid_factors(const clift_field* f, const clift_grid_head& g) {
    FactorParams p;
    for (int m = 0; m < 3; ++m) {
        p.plane[m] = g.plane[m];
        p.line[m] = g.line[m];
        p.pw[m] = f->grid[mode_a(m)];
        p.ph[m] = f->grid[mode_b(m)];
        p.ll[m] = f->grid[mode_v(m)];
    }
    p.comps = g.comps;
    return p;
}

#ifdef __CUDACC__
// One quad lane's share (channels ch, ch+16, ...) of  sum_modes sum_c P_c(x_a,x_b) * L_c(x_v).
// `lines_s` != null: density line factors staged in shared memory, modes back to back.
template <int NV>
__device__ __forceinline__ float vm_dot_partial(const FactorParams& f, const float* lines_s, float x0, float x1, float x2,
                                                int q) {
    const float xs[3] = {x0, x1, x2};
    float acc = 0.0f;
    int soff = 0;
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        const Tap2 t2 = make_tap2(xs[mode_a(m)], xs[mode_b(m)], f.pw[m], f.ph[m]);
        const Tap1 t1 = make_tap1(xs[mode_v(m)], f.ll[m]);
        const float* line = lines_s ? lines_s + soff : f.line[m];
        soff += f.ll[m] * f.comps;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const int ch = v * 16 + q * 4;
            const float4 p = plane_tap(f.plane[m], t2, f.pw[m], f.comps, ch);
            const float4 l = line_tap(line, t1, f.comps, ch);
            acc = fmaf(p.x, l.x, acc);
            acc = fmaf(p.y, l.y, acc);
            acc = fmaf(p.z, l.z, acc);
            acc = fmaf(p.w, l.w, acc);
        }
    }
    return acc;
}

// One lane's FULL sum_modes sum_c P_c(x_a,x_b) * L_c(x_v): the lane reads each tap's whole texel (comps x 4 B
// contiguous) as float4 pieces.  Half the instructions of the quad form above (taps are set up once per sample, no
// shuffles), at the price of idle lanes for out-of-box samples of a partially in-box chunk.
template <int NV>
__device__ __forceinline__ float vm_dot_full(const FactorParams& f, const float* lines_s, float x0, float x1, float x2) {
    const float xs[3] = {x0, x1, x2};
    float acc = 0.0f;
    int soff = 0;
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        const Tap2 t2 = make_tap2(xs[mode_a(m)], xs[mode_b(m)], f.pw[m], f.ph[m]);
        const Tap1 t1 = make_tap1(xs[mode_v(m)], f.ll[m]);
        const float* line = lines_s ? lines_s + soff : f.line[m];
        soff += f.ll[m] * f.comps;
#pragma unroll
        for (int v = 0; v < NV * 4; ++v) {
            const float4 p = plane_tap(f.plane[m], t2, f.pw[m], f.comps, v * 4);
            const float4 l = line_tap(line, t1, f.comps, v * 4);
            acc = fmaf(p.x, l.x, acc);
            acc = fmaf(p.y, l.y, acc);
            acc = fmaf(p.z, l.z, acc);
            acc = fmaf(p.w, l.w, acc);
        }
    }
    return acc;
}

__device__ __forceinline__ float softplus_f(float x) { return x > 20.0f ? x : log1pf(expf(x)); }
#endif

}  // namespace clift
