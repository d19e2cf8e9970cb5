// Internal host-side launch functions (one per kernel family).  Return clift_status.
#pragma once
#include <algorithm>

#include "common.cuh"
#include "march.cuh"

namespace clift {

int launch_march(const MarchParams& P, cudaStream_t stream);
int launch_scan(const int32_t* count, int32_t* offset, int32_t* bsum, unsigned long long* stats, int64_t n, int64_t cap,
                cudaStream_t stream);
int launch_fill(const clift_render_cfg* cfg, const float* rays, const float* jitter, int64_t n_rays, const Workspace& ws,
                int64_t cap, cudaStream_t stream);
int launch_finish(int64_t n_rays, int n_cls, int softmax, int add_bg, const float* opacity, const float* rgb_raw,
                  const float* sem_raw, float* rgb, float* sem, const float* dist_ray, float* dist_reg, cudaStream_t stream);
int launch_sample_points(const clift_render_cfg* cfg, const float* rays, const float* jitter, int64_t n_rays, float* z,
                         float* xyz, uint8_t* inbox, cudaStream_t stream);
int launch_density(const clift_field* field, const float* xyz, int64_t n, float* sigma, cudaStream_t stream);

// heads.cu
int launch_heads_forward(const clift_render_cfg* cfg, const clift_field* field, const float* rays, const Workspace& ws,
                         int64_t cap, int64_t n_rays, float* rgb_raw, float* sem_raw, float* ins, const StashLayout* lay,
                         cudaStream_t stream);

int launch_heads_backward(const clift_render_cfg* cfg, const clift_field* field, const Workspace& ws, const StashLayout& lay,
                          int64_t cap, int64_t n_rays, int add_bg, const clift_render_out* saved, const float* g_rgb,
                          const float* g_sem, const float* g_ins, const clift_field_grad* grad, int* ray_stride_out,
                          cudaStream_t stream);
int launch_march_backward(const clift_render_cfg* cfg, const clift_field* field, const float* rays, const float* jitter,
                          int64_t n_rays, const Workspace& ws, int ray_stride, const float* g_dist, bool have_g_w,
                          const clift_field_grad* grad, cudaStream_t stream);

// heads_tc.cu
bool heads_tc_available(const clift_field* f, int heads);
int launch_heads_forward_tc(const clift_render_cfg* cfg, const clift_field* field, const float* rays, const Workspace& ws,
                            int64_t cap, int64_t n_rays, float* rgb_raw, float* sem_raw, float* ins, cudaStream_t stream);

// heads_tc16.cu
bool heads_tc16_available(const clift_field* f, int heads);
bool heads_tc16_stash_ok(const clift_field* f, int heads);     // the tensor-core forward can record the training stash
int launch_heads_forward_tc16(const clift_render_cfg* cfg, const clift_field* field, const float* rays, const Workspace& ws,
                              int64_t cap, int64_t n_rays, float* rgb_raw, float* sem_raw, float* ins, cudaStream_t stream,
                              const StashLayout* lay = nullptr);

// heads_x16.cu: pipelined tensor-core kernel of the xyz stacks (semantic / instance), inference
bool heads_x16_available(const clift_field* f, int heads);
int launch_heads_forward_x16(const clift_render_cfg* cfg, const clift_field* field, const Workspace& ws, int64_t cap,
                             int64_t n_rays, float* sem_raw, float* ins, cudaStream_t stream, const StashLayout* lay = nullptr);

// wgrad_tc.cu: tensor-core (tcgen05 3xTF32) weight gradients over the training stashes, all layers in one launch
constexpr int kWgradTcMaxLayers = 24;
struct WgradTcItem {
    int a_row, z_row;   // first row of the layer's input / dZ inside a tile's A- / Z-stash
    int K, N;           // k_pad(in), n_pad(out)
    float* out;         // packed dW^T [K][N]
};
bool wgrad_tc_eligible(int K, int N);
int launch_wgrad_tc(const Workspace& ws, const StashLayout& lay, const WgradTcItem* items, int n_items, int64_t cap,
                    cudaStream_t stream);

// pack.cu
int launch_transpose(const float* src, float* dst, int rows, int cols, int dst_rows_pad, int dst_cols_pad, cudaStream_t stream);

}  // namespace clift
