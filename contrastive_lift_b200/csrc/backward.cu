// clift_render_backward: orchestration of the training backward on the caller's stream.
//   ray epilogue backward -> heads backward (+ per-layer wgrad) -> march backward (density factors)
// State comes from the workspace the forward (save_for_backward = 1) left behind; see include/clift_b200.h.
#include "launchers.h"

using namespace clift;

extern "C" int32_t clift_render_backward(const clift_render_cfg* cfg, const clift_field* field, const float* rays,
                                         const float* jitter, int64_t n_rays, int32_t add_background, void* workspace,
                                         int64_t workspace_bytes, int64_t max_active, const clift_render_out* saved,
                                         const float* g_rgb, const float* g_semantic, const float* g_instance,
                                         const float* g_dist_reg, const clift_field_grad* grad, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CLIFT_CHECK_ARG(cfg && field && saved && grad && n_rays >= 0, "null pointer");
    if (n_rays == 0) return CLIFT_OK;
    CLIFT_CHECK_ARG(rays && workspace, "null pointer");
    CLIFT_CHECK_ARG(saved->opacity != nullptr, "saved->opacity is required");
    const int heads = cfg->heads;
    if ((heads & CLIFT_HEAD_RGB) && g_rgb) CLIFT_CHECK_ARG(saved->rgb_raw != nullptr, "saved->rgb_raw is required for g_rgb");
    if ((heads & CLIFT_HEAD_SEMANTIC) && g_semantic)
        CLIFT_CHECK_ARG(saved->semantic_raw != nullptr, "saved->semantic_raw is required for g_semantic");
    if (max_active <= 0) max_active = n_rays * cfg->n_samples;
    const int C = field->num_classes, DI = field->dim_instance * (field->slow_fast ? 2 : 1);
    const StashLayout lay = make_stash_layout(field, heads);
    const bool z_external = saved->save_for_backward == 2;
    if (z_external) CLIFT_CHECK_ARG(saved->stash_z != nullptr, "saved->stash_z is required after a save_for_backward = 2 forward");
    Workspace ws = carve_workspace(workspace, n_rays, cfg->n_samples, max_active, 3 + C + DI, true, &lay, !z_external);
    if (z_external) ws.stash_z = saved->stash_z;
    if (ws.bytes > workspace_bytes) {
        set_error("clift_render_backward: workspace %lld bytes < required %lld", (long long)workspace_bytes, (long long)ws.bytes);
        return CLIFT_ERR_WORKSPACE;
    }
    int stride = 0;
    int rc = launch_heads_backward(cfg, field, ws, lay, max_active, n_rays, add_background, saved, g_rgb, g_semantic, g_instance,
                                   grad, &stride, stream);
    if (rc) return rc;
    const bool want_density = grad->density_plane[0] != nullptr;
    if (want_density && (heads & CLIFT_HEAD_RGB) && (g_rgb || g_dist_reg)) {
        for (int m = 0; m < 3; ++m)
            CLIFT_CHECK_ARG(grad->density_plane[m] && grad->density_line[m], "density gradient buffers must all be set or all null");
        rc = launch_march_backward(cfg, field, rays, jitter, n_rays, ws, stride, g_dist_reg, g_rgb != nullptr, grad, stream);
        if (rc) return rc;
    }
    return CLIFT_OK;
}
