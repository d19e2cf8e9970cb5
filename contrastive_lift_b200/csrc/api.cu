// extern "C" surface of libclift_b200.so: error state, configuration checks and the render pipeline
// (march -> scan -> fill -> heads -> finish) on the caller's stream.  See include/clift_b200.h.
#include <atomic>
#include <cstdarg>
#include <cstring>

#include "launchers.h"

namespace clift {

static thread_local char g_error[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count() {     // of the CURRENT device (callers run under a device guard); cached per device ordinal
    static std::atomic<int> cached[64];
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    n = cached[dev].load(std::memory_order_relaxed);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

static bool g_profile = false;
static cudaEvent_t g_ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // [5]: between the two head kernels
static bool g_ev_valid = false;
static bool g_ev_split = false;
static int g_last_head_path = 0;     // clift_debug_last_head_path()

static void profile_mark(int i, cudaStream_t stream) {
    if (g_profile && g_ev[i]) cudaEventRecord(g_ev[i], stream);
}

static int check_mlp(const clift_mlp& m, const char* name, int expect_out) {
    if (m.n_layers < 1 || m.n_layers > CLIFT_MAX_LAYERS) {
        set_error("%s: n_layers %d outside [1,%d]", name, m.n_layers, CLIFT_MAX_LAYERS);
        return CLIFT_ERR_UNSUPPORTED;
    }
    for (int l = 0; l <= m.n_layers; ++l)
        if (m.dims[l] < 1 || m.dims[l] > CLIFT_MAX_WIDTH) {
            set_error("%s: width %d of layer %d outside [1,%d]", name, m.dims[l], l, CLIFT_MAX_WIDTH);
            return CLIFT_ERR_UNSUPPORTED;
        }
    for (int l = 0; l < m.n_layers; ++l)
        if (!m.wt[l] || !m.bias[l]) {
            set_error("%s: null weight/bias pointer at layer %d", name, l);
            return CLIFT_ERR_ARG;
        }
    if (expect_out > 0 && m.dims[m.n_layers] != expect_out) {
        set_error("%s: output width %d != %d", name, m.dims[m.n_layers], expect_out);
        return CLIFT_ERR_ARG;
    }
    return CLIFT_OK;
}

static int check_grid_head(const clift_grid_head& g, const char* name) {
    if (g.comps % 16 != 0 || g.comps < 16 || g.comps > 64) {
        set_error("%s grid: comps %d not in {16,32,48,64}", name, g.comps);
        return CLIFT_ERR_UNSUPPORTED;
    }
    if (g.dim < 1 || g.dim > 64) {
        set_error("%s grid: basis width %d outside [1,64]", name, g.dim);
        return CLIFT_ERR_UNSUPPORTED;
    }
    for (int m = 0; m < 3; ++m)
        if (!g.plane[m] || !g.line[m]) {
            set_error("%s grid: null factor pointer", name);
            return CLIFT_ERR_ARG;
        }
    if (!g.basis) {
        set_error("%s grid: null basis", name);
        return CLIFT_ERR_ARG;
    }
    return CLIFT_OK;
}

static int check_field(const clift_field* f, int heads) {
    CLIFT_CHECK_ARG(f != nullptr, "null field");
    for (int k = 0; k < 3; ++k) CLIFT_CHECK_ARG(f->grid[k] >= 2, "grid dimension < 2");
    CLIFT_CHECK_SUPPORTED(f->density_comps % 16 == 0 && f->density_comps >= 16 && f->density_comps <= 48,
                          "density_comps must be 16, 32 or 48");
    for (int m = 0; m < 3; ++m) CLIFT_CHECK_ARG(f->density_plane[m] && f->density_line[m], "null density factor");
    if (heads & CLIFT_HEAD_RGB) {
        CLIFT_CHECK_SUPPORTED(f->appearance_comps % 16 == 0 && f->appearance_comps >= 16 && f->appearance_comps <= 64,
                              "appearance_comps must be 16..64 in steps of 16");
        for (int m = 0; m < 3; ++m) CLIFT_CHECK_ARG(f->appearance_plane[m] && f->appearance_line[m], "null appearance factor");
        CLIFT_CHECK_ARG(f->basis != nullptr, "null basis");
        CLIFT_CHECK_SUPPORTED(f->pe_view >= 1 && f->pe_feat >= 1, "view-independent rgb head (pe_view=pe_feat=0)");
        const int n_in = f->dim_appearance * (1 + 2 * f->pe_feat) + 3 * (1 + 2 * f->pe_view);
        CLIFT_CHECK_SUPPORTED(n_in <= CLIFT_MAX_WIDTH && 3 * f->appearance_comps <= CLIFT_MAX_WIDTH, "rgb head input too wide");
        int rc = check_mlp(f->rgb, "rgb mlp", 3);
        if (rc) return rc;
        CLIFT_CHECK_ARG(f->rgb.dims[0] == n_in, "rgb mlp input width does not match dim_appearance/pe");
    }
    if (heads & CLIFT_HEAD_SEMANTIC) {
        CLIFT_CHECK_SUPPORTED(f->num_classes >= 1 && f->num_classes <= CLIFT_MAX_HEAD_OUT, "num_classes outside [1,64]");
        int rc = check_mlp(f->semantic, "semantic mlp", f->num_classes);
        if (rc) return rc;
        if (f->semantic_grid.comps) {
            rc = check_grid_head(f->semantic_grid, "semantic");
            if (rc) return rc;
            CLIFT_CHECK_ARG(f->semantic.dims[0] == f->semantic_grid.dim, "semantic mlp input width != semantic_grid.dim");
        } else {
            CLIFT_CHECK_ARG(f->semantic.dims[0] == 3 + 6 * f->pe_sem, "semantic mlp input width != 3+6*pe_sem");
        }
    }
    if (heads & CLIFT_HEAD_INSTANCE) {
        CLIFT_CHECK_SUPPORTED(f->dim_instance >= 1 && f->dim_instance <= CLIFT_MAX_HEAD_OUT, "dim_instance outside [1,64]");
        int rc = check_mlp(f->instance_fast, "instance mlp", f->dim_instance);
        if (rc) return rc;
        if (f->instance_grid.comps) {
            rc = check_grid_head(f->instance_grid, "instance");
            if (rc) return rc;
            CLIFT_CHECK_ARG(f->instance_fast.dims[0] == f->instance_grid.dim, "instance mlp input width != instance_grid.dim");
        } else {
            CLIFT_CHECK_ARG(f->instance_fast.dims[0] == 3 + 6 * f->pe_ins, "instance mlp input width != 3+6*pe_ins");
        }
        if (f->slow_fast) {
            rc = check_mlp(f->instance_slow, "instance slow mlp", f->dim_instance);
            if (rc) return rc;
        }
    }
    return CLIFT_OK;
}

static int check_cfg(const clift_render_cfg* c) {
    CLIFT_CHECK_ARG(c != nullptr, "null cfg");
    CLIFT_CHECK_ARG(c->n_samples >= 2, "n_samples < 2");
    CLIFT_CHECK_ARG(c->step_size > 0.0f, "step_size <= 0");
    return CLIFT_OK;
}

static int out_width(const clift_field* f) { return 3 + f->num_classes + f->dim_instance * (f->slow_fast ? 2 : 1); }

}  // namespace clift

using namespace clift;

extern "C" int32_t clift_abi_version(void) { return CLIFT_ABI_VERSION; }
extern "C" const char* clift_last_error(void) { return g_error; }
extern "C" int64_t clift_launch_count(void) { return (int64_t)g_launches.load(); }

extern "C" int32_t clift_sample_points(const clift_render_cfg* cfg, const float* rays, const float* jitter, int64_t n_rays,
                                       float* z, float* xyz, uint8_t* inbox, void* stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    CLIFT_CHECK_ARG(rays && n_rays >= 0, "null rays");
    return launch_sample_points(cfg, rays, jitter, n_rays, z, xyz, inbox, (cudaStream_t)stream);
}

extern "C" int32_t clift_density(const clift_field* field, const float* xyz, int64_t n, float* sigma, void* stream) {
    int rc = check_field(field, 0);
    if (rc) return rc;
    CLIFT_CHECK_ARG(xyz && sigma && n >= 0, "null pointer");
    return launch_density(field, xyz, n, sigma, (cudaStream_t)stream);
}

extern "C" int64_t clift_render_workspace_bytes(const clift_render_cfg* cfg, const clift_field* field, int64_t n_rays,
                                                int64_t max_active, int32_t save_for_backward) {
    if (!cfg || !field || n_rays < 0) return CLIFT_ERR_ARG;
    if (max_active <= 0) max_active = n_rays * cfg->n_samples;
    const StashLayout lay = make_stash_layout(field, cfg->heads);
    return carve_workspace(nullptr, n_rays, cfg->n_samples, max_active, out_width(field), save_for_backward != 0, &lay,
                           save_for_backward != 2).bytes;
}

extern "C" int64_t clift_render_stash_z_bytes(const clift_render_cfg* cfg, const clift_field* field, int64_t n_rays,
                                              int64_t max_active) {
    if (!cfg || !field || n_rays < 0) return CLIFT_ERR_ARG;
    if (max_active <= 0) max_active = n_rays * cfg->n_samples;
    const StashLayout lay = make_stash_layout(field, cfg->heads);
    return round_up(ceil_div(max_active, CLIFT_TILE) * (int64_t)lay.z_rows * CLIFT_TILE * (int64_t)sizeof(float), 256);
}

extern "C" int32_t clift_render_forward(const clift_render_cfg* cfg, const clift_field* field, const float* rays,
                                        const float* jitter, int64_t n_rays, int32_t add_background, void* workspace,
                                        int64_t workspace_bytes, int64_t max_active, const clift_render_out* out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = check_cfg(cfg);
    if (rc) return rc;
    rc = check_field(field, cfg->heads);
    if (rc) return rc;
    CLIFT_CHECK_ARG(n_rays >= 0 && out, "negative n_rays or null out");
    if (n_rays == 0) {   // empty chunk: nothing to launch except the (empty) mean
        if (out->dist_reg) CLIFT_CUDA(cudaMemsetAsync(out->dist_reg, 0, sizeof(float), stream));
        return CLIFT_OK;
    }
    CLIFT_CHECK_ARG(rays && workspace, "null pointer");
    CLIFT_CHECK_ARG(n_rays * (int64_t)cfg->n_samples < (1ll << 31), "n_rays*n_samples must be < 2^31 per call");
    if (max_active <= 0) max_active = n_rays * cfg->n_samples;
    const int C = field->num_classes, DI = field->dim_instance * (field->slow_fast ? 2 : 1);
    const bool save = out->save_for_backward != 0;
    const StashLayout lay = make_stash_layout(field, cfg->heads);
    Workspace ws = carve_workspace(workspace, n_rays, cfg->n_samples, max_active, out_width(field), save, &lay,
                                   out->save_for_backward != 2);
    if (ws.bytes > workspace_bytes) {
        set_error("clift_render_forward: workspace %lld bytes < required %lld", (long long)workspace_bytes, (long long)ws.bytes);
        return CLIFT_ERR_WORKSPACE;
    }
    const int heads = cfg->heads;
    CLIFT_CHECK_ARG(out->opacity && out->depth, "opacity and depth outputs are required");
    if (heads & CLIFT_HEAD_RGB) CLIFT_CHECK_ARG(out->rgb && out->rgb_raw, "rgb and rgb_raw outputs required for the rgb head");
    if (heads & CLIFT_HEAD_SEMANTIC)
        CLIFT_CHECK_ARG(out->semantic && out->semantic_raw, "semantic and semantic_raw outputs required for the semantic head");
    if (heads & CLIFT_HEAD_INSTANCE) CLIFT_CHECK_ARG(out->instance, "instance output required for the instance head");
    if (out->dist_reg) CLIFT_CHECK_ARG(out->dist_ray, "dist_ray is required when dist_reg is requested");

    CLIFT_CUDA(cudaMemsetAsync(ws.stats, 0, 16 * sizeof(int32_t), stream));
    g_ev_split = false;
    MarchParams M;
    M.g = make_geom(cfg);
    M.f = make_factors(field, false);
    M.shift = field->density_shift;
    M.lines_in_smem = 0;
    M.rays = rays;
    M.jitter = jitter;
    M.n_rays = n_rays;
    M.w_dense = out->weights ? out->weights : ws.w_dense;
    M.sigma_dense = save ? ws.sigma_dense : nullptr;
    M.trans_dense = save ? ws.trans_dense : nullptr;
    M.count = ws.count;
    M.opacity = out->opacity;
    M.depth = out->depth;
    M.dist_ray = out->dist_ray;
    M.points = out->points;
    M.stats = reinterpret_cast<unsigned long long*>(ws.stats);
    M.scan_state = nullptr;
    M.rec_pos = ws.rec_pos;
    M.rec_ray = ws.rec_ray;
    M.rec_idx = ws.rec_idx;
    M.cap = max_active;
    // the march emits the active-sample records itself (no dense [B,S] weights, no scan / fill launches); the per-group scan
    // state lives in the (then unused) offset array.  Training forwards additionally write sigma and T of the in-box samples
    // for the march backward.  Callers that ask for the dense weights keep the two-pass form.  CLIFT_MARCH_FUSED=0:
    // development switch back to the two-pass form.
    bool fused = heads && !out->weights && (int64_t)round_up(cfg->n_samples, 32) * 8 * 4 <= 96 * 1024;
    {
        const char* e = getenv("CLIFT_MARCH_FUSED");
        if (e && atoi(e) == 0) fused = false;
    }
    if (fused) {
        M.scan_state = reinterpret_cast<unsigned long long*>(ws.offset);
        CLIFT_CUDA(cudaMemsetAsync(ws.offset, 0, (size_t)ceil_div(n_rays, 8) * 8, stream));
    }
    profile_mark(0, stream);
    rc = launch_march(M, stream);
    if (rc) return rc;
    profile_mark(1, stream);
    if (heads) {
        if (!fused) {
            rc = launch_scan(ws.count, ws.offset, ws.bsum, M.stats, n_rays, max_active, stream);
            if (rc) return rc;
            Workspace wf = ws;
            wf.w_dense = M.w_dense;
            rc = launch_fill(cfg, rays, jitter, n_rays, wf, max_active, stream);
            if (rc) return rc;
        }
        if (heads & CLIFT_HEAD_RGB) CLIFT_CUDA(cudaMemsetAsync(out->rgb_raw, 0, n_rays * 3 * sizeof(float), stream));
        if (heads & CLIFT_HEAD_SEMANTIC) CLIFT_CUDA(cudaMemsetAsync(out->semantic_raw, 0, n_rays * C * sizeof(float), stream));
        if (heads & CLIFT_HEAD_INSTANCE) CLIFT_CUDA(cudaMemsetAsync(out->instance, 0, n_rays * DI * sizeof(float), stream));
        profile_mark(2, stream);
        float* o_rgb = (heads & CLIFT_HEAD_RGB) ? out->rgb_raw : nullptr;
        float* o_sem = (heads & CLIFT_HEAD_SEMANTIC) ? out->semantic_raw : nullptr;
        float* o_ins = (heads & CLIFT_HEAD_INSTANCE) ? out->instance : nullptr;
        int path = CLIFT_HEADS_FMA;
        const bool grid_heads = ((heads & CLIFT_HEAD_SEMANTIC) && field->semantic_grid.comps) ||
                                ((heads & CLIFT_HEAD_INSTANCE) && field->instance_grid.comps);
        const bool tc16_train = save && heads_tc16_available(field, heads) && heads_tc16_stash_ok(field, heads);
        if (cfg->head_path == CLIFT_HEADS_TENSOR || cfg->head_path == CLIFT_HEADS_TENSOR16) {
            CLIFT_CHECK_SUPPORTED(!save || (cfg->head_path == CLIFT_HEADS_TENSOR16 && tc16_train),
                                  "this tensor-core head path cannot record the training stash for this field (use CLIFT_HEADS_AUTO/FMA)");
            // grid-mode heads exist on the fp16-split kernel only: _TENSOR means that kernel for them
            path = grid_heads ? CLIFT_HEADS_TENSOR16 : cfg->head_path;
        } else if (cfg->head_path == CLIFT_HEADS_AUTO && save) {
            if (tc16_train) path = CLIFT_HEADS_TENSOR16;
        } else if (cfg->head_path == CLIFT_HEADS_AUTO) {
            if (heads_tc16_available(field, heads))
                path = CLIFT_HEADS_TENSOR16;
            else if (!grid_heads && heads_tc_available(field, heads))
                path = CLIFT_HEADS_TENSOR;
        }
        const int xyz_heads = grid_heads ? 0 : heads & (CLIFT_HEAD_SEMANTIC | CLIFT_HEAD_INSTANCE);
        g_last_head_path = path;
        if (path == CLIFT_HEADS_TENSOR16 && xyz_heads && heads_x16_available(field, xyz_heads)) {
            g_last_head_path = path | 16;
            // the xyz stacks on the pipelined kernel, the rgb stack (gather, basis, encoding) on the serial one; training
            // forwards record the stash from both
            rc = launch_heads_forward_x16(cfg, field, ws, max_active, n_rays, o_sem, o_ins, stream, save ? &lay : nullptr);
            profile_mark(5, stream);
            g_ev_split = g_profile;
            if (!rc && o_rgb)
                rc = launch_heads_forward_tc16(cfg, field, rays, ws, max_active, n_rays, o_rgb, nullptr, nullptr, stream,
                                               save ? &lay : nullptr);
        } else if (path == CLIFT_HEADS_TENSOR16)
            rc = launch_heads_forward_tc16(cfg, field, rays, ws, max_active, n_rays, o_rgb, o_sem, o_ins, stream, save ? &lay : nullptr);
        else if (path == CLIFT_HEADS_TENSOR)
            rc = launch_heads_forward_tc(cfg, field, rays, ws, max_active, n_rays, o_rgb, o_sem, o_ins, stream);
        else
            rc = launch_heads_forward(cfg, field, rays, ws, max_active, n_rays, o_rgb, o_sem, o_ins, save ? &lay : nullptr, stream);
        if (rc) return rc;
    } else {
        profile_mark(2, stream);
    }
    profile_mark(3, stream);
    rc = launch_finish(n_rays, C, cfg->semantic_softmax, add_background, out->opacity,
                         (heads & CLIFT_HEAD_RGB) ? out->rgb_raw : nullptr, (heads & CLIFT_HEAD_SEMANTIC) ? out->semantic_raw : nullptr,
                         (heads & CLIFT_HEAD_RGB) ? out->rgb : nullptr, (heads & CLIFT_HEAD_SEMANTIC) ? out->semantic : nullptr,
                         out->dist_ray, out->dist_reg, stream);
    profile_mark(4, stream);
    g_ev_valid = g_profile;
    return rc;
}

extern "C" int32_t clift_profile_enable(int32_t on) {
    if (on && !g_ev[0])
        for (int i = 0; i < 6; ++i) CLIFT_CUDA(cudaEventCreate(&g_ev[i]));
    g_profile = on != 0;
    g_ev_valid = false;
    return CLIFT_OK;
}

extern "C" int32_t clift_profile_stage_ms(float* ms4) {
    CLIFT_CHECK_ARG(ms4 != nullptr, "null pointer");
    CLIFT_CHECK_ARG(g_ev_valid, "no profiled clift_render_forward since clift_profile_enable(1)");
    CLIFT_CUDA(cudaEventSynchronize(g_ev[4]));
    for (int i = 0; i < 4; ++i) CLIFT_CUDA(cudaEventElapsedTime(&ms4[i], g_ev[i], g_ev[i + 1]));
    return CLIFT_OK;
}

extern "C" int32_t clift_debug_last_head_path(void) { return g_last_head_path; }

extern "C" int32_t clift_profile_heads_split_ms(float* ms2) {
    CLIFT_CHECK_ARG(ms2 != nullptr, "null pointer");
    CLIFT_CHECK_ARG(g_ev_valid, "no profiled clift_render_forward since clift_profile_enable(1)");
    ms2[0] = ms2[1] = 0.0f;
    if (!g_ev_split) return CLIFT_OK;          // one head kernel ran: no split
    CLIFT_CUDA(cudaEventSynchronize(g_ev[4]));
    CLIFT_CUDA(cudaEventElapsedTime(&ms2[0], g_ev[2], g_ev[5]));
    CLIFT_CUDA(cudaEventElapsedTime(&ms2[1], g_ev[5], g_ev[3]));
    return CLIFT_OK;
}

extern "C" int32_t clift_render_stats(const void* workspace, int64_t* stats4, void* stream) {
    CLIFT_CHECK_ARG(workspace && stats4, "null pointer");
    CLIFT_CUDA(cudaMemcpyAsync(stats4, workspace, 4 * sizeof(int64_t), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return CLIFT_OK;
}
