/*
 * clift_b200.h - C ABI of libclift_b200.so: the B200 (sm_100a) implementation of the
 * volumetric-render + contrastive-fusion hot path of yashbhalgat/Contrastive-Lift.
 *
 * The reference has no FFI of its own (it is pure Python over ATen); each entry point below
 * replaces one group of reference functions, cited as file:line relative to the reference tree.
 * INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *  - plain C, no torch types; every array is a caller-owned DEVICE pointer to contiguous fp32
 *    (int32/int64 where stated); shapes are explicit integers; `stream` is a cudaStream_t passed
 *    as void* (the caller's current stream).
 *  - calls are stream-ordered, never synchronise the device, never allocate device memory;
 *    scratch comes from a caller-provided workspace sized by clift_render_workspace_bytes().
 *  - every function returns 0 on success or a negative clift_status; clift_last_error() gives the
 *    message for the calling thread.  Nothing throws across the boundary.
 *  - there is no CPU fallback: without a CUDA device every compute call returns CLIFT_ERR_CUDA.
 */
#ifndef CLIFT_B200_H
#define CLIFT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLIFT_ABI_VERSION 13
#define CLIFT_MAX_LAYERS 8
#define CLIFT_MAX_WIDTH 256      /* widest MLP layer / widest head input */
#define CLIFT_MAX_HEAD_OUT 64    /* semantic classes, and instance-embedding width per net */
#define CLIFT_TILE 128           /* active samples per head tile */

typedef enum {
    CLIFT_OK = 0,
    CLIFT_ERR_ARG = -1,          /* null pointer, negative size, misaligned buffer */
    CLIFT_ERR_UNSUPPORTED = -2,  /* configuration outside the compiled envelope */
    CLIFT_ERR_CUDA = -3,         /* CUDA runtime / launch failure (message has the cudaError) */
    CLIFT_ERR_WORKSPACE = -4     /* workspace too small */
} clift_status;

/* One fully-connected stack (nn.Sequential of Linear/ReLU: tensoRF.py:393-397, 473-490, 577-582).
 * Weights are the PACKED form written by clift_pack_linear(): W^T, row-major [k_pad][n_pad] with
 * k_pad = round_up(in,16), n_pad = round_up(out,64), zero filled; bias is [n_pad] zero filled. */
typedef struct {
    int32_t n_layers;
    int32_t dims[CLIFT_MAX_LAYERS + 1];       /* logical widths: in, hidden..., out */
    const float* wt[CLIFT_MAX_LAYERS];
    const float* bias[CLIFT_MAX_LAYERS];
    /* training only (may be null for inference): the same weights in the data-gradient layout written
     * by clift_pack_linear_dgrad(): W row-major [round_up(out,16)][dgrad_pad(in)], zero filled, where
     * dgrad_pad(in) = 64, 128 or 256 (smallest that fits). */
    const float* w_dgrad[CLIFT_MAX_LAYERS];
    /* tensor-core operand written by clift_pack_linear_tc() (null = the FP32-FMA head kernel is used) */
    const float* w_tc[CLIFT_MAX_LAYERS];
    /* fp16-split tensor-core operand written by clift_pack_linear_tc16() (null = not available) */
    const void* w_tc16[CLIFT_MAX_LAYERS];
    /* data-gradient operand for the tensor-core training backward: clift_pack_linear_tc16() of W^T ([in][out] row-major,
     * no bias), i.e. the B operand of dA[m][i] = sum_o dZ[m][o] W[o][i] (null = that layer's data gradient runs on FP32 FMA
     * from w_dgrad) */
    const void* w_dg16[CLIFT_MAX_LAYERS];
    /* clift_pack_linear_x16() re-layout of w_tc16 for the pipelined inference kernel of the xyz stacks (256-wide layers: one
     * 8 KB stage per k-step and N-half; null = the layer is not 256 wide or the stack runs on the serial kernel) */
    const void* w_x16[CLIFT_MAX_LAYERS];
} clift_mlp;

/* Gradient mirror of clift_mlp (same packed shapes); null pointers = do not accumulate. */
typedef struct {
    float* wt[CLIFT_MAX_LAYERS];
    float* bias[CLIFT_MAX_LAYERS];
} clift_mlp_grad;

/* Grid-mode semantic / instance head input (use_mlp_for_semantics / use_mlp_for_instances = False: allgrid.yaml,
 * instGRIDsemMLP.yaml, onlyRGBsegGRID.yaml; tensoRF.py:72-85, 127-156): a VM factor set in the same packed layout as the
 * appearance factors, reduced by a bias-free basis Linear (3*comps -> dim) whose output is the head MLP's input
 * (dims[0] of that clift_mlp == dim).  comps == 0: MLP mode, the head reads xyz (+ positional encoding). */
typedef struct {
    int32_t comps;               /* 0 = off, else 16, 32, 48 or 64 per mode (the reference builds 32) */
    int32_t dim;                 /* dim_semantics / dim_instances = 27 (<= 64) */
    const float* plane[3];
    const float* line[3];
    const float* basis;          /* clift_pack_linear() of the basis weight */
    const float* basis_dgrad;    /* clift_pack_linear_dgrad() of it (training only, else null) */
    const void* basis_tc16;      /* clift_pack_linear_tc16() of it (tensor-core heads, else null) */
} clift_grid_head;

typedef struct {
    float* plane[3];
    float* line[3];
    float* basis;
} clift_grid_head_grad;

/* TensorVMSplit parameters (tensoRF.py:32-106) in the kernels' HBM layout.
 * plane i: channel-last [H=grid[b]][W=grid[a]][comps] with (a,b)=matrix_mode[i] in {(0,1),(0,2),(1,2)}
 * line  i: [grid[v]][comps] with v=vector_mode[i] in {2,1,0};  written by clift_pack_plane/line(). */
typedef struct {
    int32_t grid[3];                 /* samples per axis x,y,z (renderer.grid_dim) */
    int32_t density_comps;           /* 16 (supported: 16, 32, 48) */
    int32_t appearance_comps;        /* 48 (supported: 16, 32, 48, 64) */
    int32_t dim_appearance;          /* 27 */
    int32_t pe_view, pe_feat;        /* 2, 2 (tensoRF.py:400-418) */
    int32_t pe_sem, pe_ins;          /* 0, 0 */
    int32_t num_classes;             /* semantic head width C */
    int32_t dim_instance;            /* per-net embedding width d (max_instances) */
    int32_t slow_fast;               /* 1: instance output is [fast | slow], 2d wide */
    float density_shift;             /* -10 */
    const float* density_plane[3];
    const float* density_line[3];
    const float* appearance_plane[3];
    const float* appearance_line[3];
    const float* basis;              /* appearance_basis_mat packed as a 1-layer clift_mlp weight */
    const float* basis_dgrad;        /* ... and in clift_pack_linear_dgrad() layout (training only, else null) */
    const float* basis_tc;           /* ... and in clift_pack_linear_tc() layout (tensor-core heads, else null) */
    const void* basis_tc16;          /* ... and in clift_pack_linear_tc16() layout (fp16-split tensor-core heads, else null) */
    const void* basis_dg16;          /* ... and its transpose in clift_pack_linear_tc16() layout (tensor-core data gradient) */
    clift_mlp rgb;                   /* render_appearance_mlp.mlp       (H1) */
    clift_mlp semantic;              /* render_semantic_mlp.mlp         (H2) */
    clift_mlp instance_fast;         /* render_instance_mlp.mlp         (H3) */
    clift_mlp instance_slow;         /* render_instance_mlp.slow_mlp    (H3) */
    clift_grid_head semantic_grid;   /* semantic_plane/line/basis_mat   (F3; comps 0 in MLP mode) */
    clift_grid_head instance_grid;   /* instance_plane/line/basis_mat   (F3; fast and slow nets share the feature) */
} clift_field;

typedef struct {
    float* density_plane[3];
    float* density_line[3];
    float* appearance_plane[3];
    float* appearance_line[3];
    float* basis;
    clift_mlp_grad rgb, semantic, instance_fast, instance_slow;
    clift_grid_head_grad semantic_grid, instance_grid;   /* required when the matching head is in grid mode */
} clift_field_grad;

/* TensoRFRenderer constants (renderer:39-78). */
typedef struct {
    float aabb_min[3], aabb_max[3];  /* renderer.bbox_aabb */
    float inv_extent[3];             /* renderer.inv_box_extent = 2/extent (fp32, as stored) */
    float step_size;                 /* renderer.step_size */
    int32_t n_samples;               /* renderer.n_samples */
    float distance_scale;            /* 25 */
    float weight_thres;              /* raymarch_weight_thres 1e-4 */
    int32_t semantic_softmax;        /* semantic_weight_mode == "softmax" */
    int32_t heads;                   /* bit mask of CLIFT_HEAD_* to evaluate */
    int32_t head_path;               /* CLIFT_HEADS_AUTO / _FMA / _TENSOR / _TENSOR16 */
} clift_render_cfg;

/* MLP-head implementation.  AUTO = tcgen05 tensor cores for inference - the fp16-split path (3 kind::f16 MMAs per
 * product, scaled operands, fp32-faithful) when the field carries w_tc16 operands, else the 3xTF32 path (w_tc) -
 * and FP32 FMA otherwise.  save_for_backward forwards (they record the training stash) take the fp16-split path when every
 * stash block has a writer there, else FP32 FMA.  Grid-mode semantic / instance heads (clift_grid_head.comps > 0): inference
 * runs on the fp16-split tensor-core kernel (gather of the head's factor set -> basis GEMM -> MLP stack) when the head carries
 * basis_tc16 (inference and training forwards); for such heads _TENSOR means the same kernel (there is no 3xTF32 form). */
#define CLIFT_HEADS_AUTO 0
#define CLIFT_HEADS_FMA 1
#define CLIFT_HEADS_TENSOR 2
#define CLIFT_HEADS_TENSOR16 3

#define CLIFT_HEAD_RGB 1
#define CLIFT_HEAD_SEMANTIC 2
#define CLIFT_HEAD_INSTANCE 4
#define CLIFT_HEAD_ALL 7

/* Per-call outputs of the march + heads (all [n_rays, ...], fp32). Null = not wanted. */
typedef struct {
    float* rgb;          /* [B,3]   final: clamp(raw + bg*(1-opacity), 0, 1)       renderer:164-167 */
    float* semantic;     /* [B,C]   final: log(p/(sum+1e-8)+1e-8) in softmax mode  renderer:160-162 */
    float* instance;     /* [B,2d]  (or [B,d] without slow net)                    renderer:149    */
    float* depth;        /* [B]     sum w*t                                        renderer:174    */
    float* opacity;      /* [B]     sum w                                          renderer:137    */
    float* dist_reg;     /* [1]     mean over rays of the distortion loss          renderer:101    */
    float* rgb_raw;      /* [B,3]   sum w*rgb before bg/clamp   (saved for backward) */
    float* semantic_raw; /* [B,C]   sum w*sem before normalise  (saved for backward) */
    float* dist_ray;     /* [B]     per-ray distortion loss */
    float* points;       /* [B,3]   o + depth*d (forward_instance_feature, renderer:211-213) */
    float* weights;      /* [B,S]   dense compositing weights w_i (parity tests); null = keep in workspace */
    int32_t save_for_backward; /* 1: keep what clift_render_backward needs in the workspace; 2: the same, except that the
                                  dL/dZ stash of the backward is NOT part of the workspace - the caller passes it to
                                  clift_render_backward in `stash_z` (clift_render_stash_z_bytes), so a forward that waits
                                  for its backward (several chunks per step) holds only its activation stash */
    float* stash_z;            /* clift_render_backward with save_for_backward = 2: scratch of clift_render_stash_z_bytes() */
} clift_render_out;

/* ---- library ---------------------------------------------------------------------------- */
int32_t clift_abi_version(void);
const char* clift_last_error(void);
/* Number of kernel launches issued by this library since load (all threads); bench.py's gpu_launches. */
int64_t clift_launch_count(void);

/* Per-stage device timing of clift_render_forward for bench.py's roofline figures.  When enabled, CUDA events
 * are recorded on the caller's stream between the stages of every forward; clift_profile_stage_ms() waits for
 * the last profiled forward and returns its stage durations in ms:
 *   ms[0] march (sampling+density+scan+compositing)  ms[1] active-sample scan+compaction
 *   ms[2] MLP heads                                   ms[3] per-ray epilogue
 * Off by default (no events, no synchronisation).  Not thread-safe: one profiled stream at a time. */
int32_t clift_profile_enable(int32_t on);
int32_t clift_profile_stage_ms(float* ms4);
/* When the last profiled forward ran the heads as two kernels (pipelined xyz-stack kernel + rgb-stack kernel, the inference
 * default): their device times {xyz, rgb} in ms; {0, 0} when one kernel evaluated all heads. */
int32_t clift_profile_heads_split_ms(float* ms2);
/* Which head kernels the last clift_render_forward of this process chose: CLIFT_HEADS_FMA / _TENSOR / _TENSOR16, + 16 when
 * the xyz stacks ran on the pipelined kernel (heads_x16) next to the fp16-split rgb kernel; 0 = none yet.  Tests use it to
 * prove a configuration did not fall back to another path. */
int32_t clift_debug_last_head_path(void);

/* ---- layout packing (one transpose kernel; used for parameters and, inverted, for gradients) --- */
/* (1,C,H,W) -> [H][W][C]  and back.   tensoRF.py:99-106 layouts. */
int32_t clift_pack_plane(const float* nchw, float* hwc, int32_t comps, int32_t h, int32_t w, void* stream);
int32_t clift_unpack_plane(const float* hwc, float* nchw, int32_t comps, int32_t h, int32_t w, void* stream);
/* nn.Linear weight [out][in] (+ bias[out] or null) -> packed W^T [k_pad][n_pad], bias [n_pad]; and back. */
int32_t clift_pack_linear(const float* w, const float* b, float* wt, float* bias_pad,
                          int32_t n_out, int32_t n_in, void* stream);
int32_t clift_unpack_linear(const float* wt, const float* bias_pad, float* w, float* b,
                            int32_t n_out, int32_t n_in, void* stream);
/* nn.Linear weight [out][in] -> zero padded copy [round_up(out,16)][dgrad_pad(in)] (data-gradient operand). */
int32_t clift_pack_linear_dgrad(const float* w, float* w_dgrad, int32_t n_out, int32_t n_in, void* stream);

/* All of the above for a whole model in ONE launch (a training step repacks every parameter before each render and unpacks
 * every gradient after each backward: ~150 tiny launches otherwise).  `jobs` is a DEVICE array; job i covers the 32x32
 * destination tiles [first_tile, first_tile + ceil(d_rows/32)*ceil(d_cols/32)) of the launch, first_tile being the running
 * sum over the jobs before it; total_tiles = the sum over all jobs.
 *   kind 0: dst[i][j] = src[j*s_pitch + i]  (transpose)     kind 1: dst[i][j] = src[i*s_pitch + j]  (copy)
 * for i < d_rows, j < d_cols, zero where the source index is outside s_rows x s_cols or src is null.
 *   clift_pack_plane   = kind 0, src (comps x h*w, pitch h*w) -> dst (h*w x comps);  clift_unpack_plane the inverse
 *   clift_pack_linear  = kind 0, src (n_out x n_in, pitch n_in) -> dst (k_pad x n_pad), + kind 1 for the bias (1 x n_pad)
 *   clift_pack_linear_dgrad = kind 1, src (n_out x n_in) -> dst (round_up(out,16) x dgrad_pad(in)) */
typedef struct {
    const float* src;
    float* dst;
    int32_t s_pitch, s_rows, s_cols;
    int32_t d_rows, d_cols;
    int32_t kind;
    int32_t first_tile;
    int32_t reserved;
} clift_pack_job;
int32_t clift_pack_batch(const clift_pack_job* jobs, int32_t n_jobs, int32_t total_tiles, void* stream);

/* Tensor-core operand of one nn.Linear for the tcgen05 head kernels: W [out][in] (+ bias[out] or null) -> tf32-exact
 * (hi, lo) pairs in 8-row K slabs (one tcgen05.mma k-step each), n_pad = round_up(out, 32); with a bias one more slab
 * follows whose first K row is the bias (the kernels multiply it with a constant-one operand column).
 * clift_tc_weight_floats() sizes the buffer. */
int64_t clift_tc_weight_floats(int32_t n_out, int32_t n_in, int32_t has_bias);
int32_t clift_pack_linear_tc(const float* w, const float* bias, float* dst, int32_t n_out, int32_t n_in, void* stream);
/* Pipelined xyz-stack kernel (heads_x16.cu): the fp16-split operand of a 256-wide layer re-ordered from
 * clift_pack_linear_tc16()'s k-step slabs into one 8 KB stage per (N-half, k-step): [W_hi: 2 k-chunks x 128 rows x 16 B | W_lo],
 * half 0's stages first, the bias step last in each half.  clift_x16_weight_bytes() sizes dst (< 0: layer not eligible). */
int64_t clift_x16_weight_bytes(int32_t n_out, int32_t n_in, int32_t has_bias);
int32_t clift_pack_linear_x16(const void* w_tc16, void* dst, int32_t n_out, int32_t n_in, int32_t has_bias, void* stream);
/* The same for a device-resident table of layers in ONE launch (job i covers blocks [first_block, first_block +
 * ceil(steps * 1024 / 256)) of the launch; steps = ceil(n_in / 16) + has_bias; total_blocks = their sum). */
typedef struct {
    const void* w_tc16;
    void* dst;
    int32_t steps;
    int32_t first_block;
} clift_x16_job;
int32_t clift_pack_linear_x16_batch(const clift_x16_job* jobs, int32_t n_jobs, int32_t total_blocks, void* stream);
/* Bring-up / parity entry for the tensor-core GEMM core: out[128][round_up(n_out,32)] = a[128][k] * W^T (+ bias) with
 * the 3xTF32 split (one CTA).  Used by tests only. */
int32_t clift_debug_tc_gemm(const float* a, const float* w_tc, float* out, int32_t k, int32_t n_out, int32_t has_bias,
                            void* stream);

/* fp16-split tensor-core operand of one nn.Linear (heads_tc16): a 64-byte header of fp32 scalars
 *   [0] ca = power-of-two scale of the layer's input   [1] cw = scale of the weights   [2] 1/(ca*cw)
 *   [3] bound on |output| = in_bound * max_n sum_k |W_nk| + max|b|   [4] in_bound   [5..7] the three norms
 * followed by fp16 (hi, lo) pairs of cw*W in 16-row K slabs (one kind::f16 MMA k-step each), n_pad = round_up(out, 32);
 * with a bias one more slab follows whose first K row is ca*cw*b.  The input bound is max(in_bound_floor, *in_bound)
 * (`in_bound` = device scalar or null): pass the previous layer's header + 3 to chain a stack, 1.0 as the floor for inputs
 * that contain coordinates / sin / cos / unit directions.  Everything is computed on the device, stream-ordered.
 * clift_tc16_weight_bytes() sizes the buffer (16-byte aligned). */
int64_t clift_tc16_weight_bytes(int32_t n_out, int32_t n_in, int32_t has_bias);
int32_t clift_pack_linear_tc16(const float* w, const float* bias, void* dst, int32_t n_out, int32_t n_in,
                               const float* in_bound, float in_bound_floor, void* stream);
/* The same for a whole model in two launches.  `jobs` is a DEVICE array in dependency order; jobs with the same `chain`
 * (0 .. n_chains-1) are one stack of layers planned one after the other by one CTA, so a job's `in_bound` may point at
 * header word 3 of the previous job of its chain.  Job i owns the pack blocks [first_block, first_block + ceil(slabs*16*n_pad
 * / 256)) of the second launch (running sum, n_pad = round_up(n_out,32), slabs = ceil(n_in/16) + (bias != null));
 * total_blocks = the sum.  Same argument rules as clift_pack_linear_tc16 (not re-checked per job). */
typedef struct {
    const float* w;
    const float* bias;
    void* dst;
    const float* in_bound;
    float in_bound_floor;
    int32_t n_out, n_in;
    int32_t chain;
    int32_t first_block;
    int32_t reserved;
} clift_tc16_job;
int32_t clift_pack_linear_tc16_batch(const clift_tc16_job* jobs, int32_t n_jobs, int32_t n_chains, int32_t total_blocks,
                                     void* stream);
/* Bound on |plane * line| products of a VM factor set for the chain above: scratch8[6] = max_mode max|plane| * max|line|
 * (scratch8 = 8 device floats; packed or unpacked factors, only the element counts matter). */
int32_t clift_tc16_factor_bound(const float* const* planes3, const float* const* lines3, const int64_t* plane_elems3,
                                const int64_t* line_elems3, float* scratch8, void* stream);
/* Bring-up / parity entry: out[128][round_up(n_out,32)] = a[128][k] * W^T (+ bias) through the fp16-split core (one CTA);
 * |a| must respect the input bound the operand was packed with.  Used by tests only. */
int32_t clift_debug_tc16_gemm(const float* a, const void* w_tc16, float* out, int32_t k, int32_t n_out, int32_t has_bias,
                              void* stream);

/* Debug: device buffer of 4*24*10 int64 that CTA 0 of the tensor-core head kernel fills with clock64() stamps per
 * (tile, GEMM): 0 accumulator seen, 1 next operand written, 2 operand published, 3 MMA thread saw operand,
 * 4 MMAs issued, 5 first weight slab seen, 6-7 phase marks inside the operand build.  null disables. */
int32_t clift_debug_tc_trace(long long* device_buf);

/* ---- R1-R4: util/ray.py:8-12,25-31,46-54,81-99 + dataset/base.py:211-219 ------------------------
 * rays[row*W+col] = [o(3), d(3), near, far]; intrinsics (3x3) and cam2world (4x4) are HOST row-major.
 * *bad_rays (device int32) counts rays whose sphere determinant is negative (the reference asserts). */
int32_t clift_gen_rays(const float* intrinsics9, const float* cam2world16, int32_t height, int32_t width,
                       float near, float radius, float* rays, int32_t* bad_rays, void* stream);

/* ---- S1,S3: renderer:800-817, 633-634.  Debug/parity entry: materialises what the march computes
 * on the fly.  z [B,S], xyz [B,S,3] (normalised), inbox [B,S] uint8; jitter [B] or null. */
int32_t clift_sample_points(const clift_render_cfg* cfg, const float* rays, const float* jitter, int64_t n_rays,
                            float* z, float* xyz, uint8_t* inbox, void* stream);

/* ---- F1: tensoRF.py:108-125  sigma[n] = softplus(sum_modes sum_c P*L + shift) at normalised xyz [n,3]. */
int32_t clift_density(const clift_field* field, const float* xyz, int64_t n, float* sigma, void* stream);

/* ---- S1-C3: TensoRFRenderer.forward (renderer:80-176), forward_instance_feature (:178-217),
 * forward_segment_feature (:259-300) - selected by cfg->heads.
 *  rays [B,8]; jitter [B] = perturb*U[0,1) or null (inference); add_background = outcome of
 *  `white_bg or (is_train and rand<0.5)` (renderer:164) - RNG stays with the caller for parity.
 *  max_active bounds the compacted active-sample list (<= 0: worst case n_rays*n_samples); one call
 *  handles n_rays*n_samples < 2^31 - callers split larger frames.                                */
int64_t clift_render_workspace_bytes(const clift_render_cfg* cfg, const clift_field* field, int64_t n_rays,
                                     int64_t max_active, int32_t save_for_backward);
int64_t clift_render_stash_z_bytes(const clift_render_cfg* cfg, const clift_field* field, int64_t n_rays, int64_t max_active);
int32_t clift_render_forward(const clift_render_cfg* cfg, const clift_field* field, const float* rays,
                             const float* jitter, int64_t n_rays, int32_t add_background,
                             void* workspace, int64_t workspace_bytes, int64_t max_active,
                             const clift_render_out* out, void* stream);
/* Counters of the last forward on this workspace, copied device-to-device into int64 stats4[4] =
 * {n_active (weight > thres), n_inbox, overflow flag (n_active > max_active), n_tiles}. */
int32_t clift_render_stats(const void* workspace, int64_t* stats4, void* stream);

/* Backward of clift_render_forward given dL/d(rgb, semantic, instance, dist_reg) (null = zero).
 * Must follow a forward with out->save_for_backward = 1 on the SAME workspace (which then holds the
 * per-sample sigma / transmittance / weights, the compacted active-sample records and the per-layer
 * head activations of those records).  Accumulates (+=) into `grad` in packed layout; null members are
 * skipped.  Semantic/instance gradients stop at the heads (stop_semantic_grad=True, renderer:144-149):
 * density factors receive gradient only through rgb and dist_reg.  The workspace is consumed. */
int32_t clift_render_backward(const clift_render_cfg* cfg, const clift_field* field, const float* rays,
                              const float* jitter, int64_t n_rays, int32_t add_background,
                              void* workspace, int64_t workspace_bytes, int64_t max_active,
                              const clift_render_out* saved, const float* g_rgb, const float* g_semantic,
                              const float* g_instance, const float* g_dist_reg,
                              const clift_field_grad* grad, void* stream);

/* ---- L1: slow-fast contrastive loss, trainer/train_panopli_tensorf.py:256-310 (forward + gradient).
 * features [N,2d] = [fast|slow], labels int64 [N], confidences [N]; loss [1]; grad_features [N,2d]
 * (d loss / d features for upstream 1; only fast half / fast columns are non-zero).            */
int32_t clift_slowfast_loss(const float* features, const int64_t* labels, const float* confidences,
                            int32_t n, int32_t d, float* loss, float* grad_features, void* stream);
/* trainer:325-329: slow = slow*momentum + (1-momentum)*fast over a flat fp32 arena.  `momentum` is the Python
 * double: the reference rounds momentum and (1 - momentum) to fp32 separately, which needs the double here. */
int32_t clift_ema_update(float* slow, const float* fast, int64_t n, double momentum, void* stream);
/* Every (slow, fast) parameter pair of a network in ONE launch; `table` is a DEVICE array, max_n the largest n. */
typedef struct {
    float* slow;
    const float* fast;
    int64_t n;
} clift_ema_pair;
int32_t clift_ema_update_batch(const clift_ema_pair* table, int32_t n_pairs, int64_t max_n, double momentum, void* stream);
/* ---- L2: model/loss/loss.py:62-82. features [N,D]. */
int32_t clift_contrastive_loss(const float* features, const int64_t* labels, int32_t n, int32_t dim,
                               float temperature, float* loss, float* grad_features, void* stream);
/* ---- TV: model/loss/loss.py:9-26 on a packed plane [H][W][C]; adds scale*dTV/dplane into grad (may be null). */
int32_t clift_tv_loss(const float* plane_hwc, int32_t comps, int32_t h, int32_t w, float* loss,
                      float* grad_hwc, float grad_scale, void* stream);
/* The same for a list of planes in two launches (all values, then all gradients); `jobs` is a DEVICE array, max_n the
 * largest comps*h*w; a null loss / grad_hwc skips that half for the plane. */
typedef struct {
    const float* plane_hwc;
    float* loss;
    float* grad_hwc;
    int32_t comps, h, w;
    float grad_scale;
} clift_tv_job;
int32_t clift_tv_loss_batch(const clift_tv_job* jobs, int32_t n_jobs, int64_t max_n, void* stream);

/* ---- torch.optim.Adam step (trainer/__init__.py:134-139; trainer:98-103,199,221), fused over a table
 * of tensors: ONE launch updates every parameter of a param group (amsgrad=False, maximize=False - the reference's use).
 * `table` is a DEVICE array of n_tensors entries; max_n = the largest entry's n (sizes the grid); step = the 1-based step
 * count AFTER this update; grad_scale multiplies every gradient first (1.0, or 1/world_size after a SUM all-reduce). */
typedef struct {
    float* param;
    const float* grad;
    float* exp_avg;
    float* exp_avg_sq;
    int64_t n;
} clift_adam_tensor;
int32_t clift_adam_step(const clift_adam_tensor* table, int32_t n_tensors, int64_t max_n, float lr, float beta1, float beta2,
                        float eps, float weight_decay, int64_t step, float grad_scale, void* stream);
/* The same for several param groups (an optimizer's whole step) in ONE launch: `groups` is a HOST array of n_groups <=
 * CLIFT_MAX_ADAM_GROUPS entries that tile the table in order (group g owns tensors [first, first + count)), each with its own
 * hyper-parameters and step count. */
#define CLIFT_MAX_ADAM_GROUPS 8
typedef struct {
    float lr, beta1, beta2, eps, weight_decay;
    int32_t first, count;
    int32_t reserved;
    int64_t step;
} clift_adam_group;
int32_t clift_adam_step_groups(const clift_adam_tensor* table, int32_t n_tensors, int64_t max_n, const clift_adam_group* groups,
                               int32_t n_groups, float grad_scale, void* stream);

/* ---- epoch-boundary volume operations.
 * clift_dense_alpha  (renderer:717-729,744-748): alpha[i][j][k] = 1 - exp(-sigma(p_ijk) * cfg->step_size) at the lattice
 *   p = aabb_min*(1-s) + aabb_max*s, s = the caller's torch.linspace(0,1,grid[c]) per axis (device arrays sx, sy, sz).
 * clift_alpha_bbox   (renderer:671-681): clamp -> 3x3x3 max-pool (stride 1, padding 1) -> >= threshold; bbox6 = min xyz,
 *   max xyz of the surviving lattice positions, *count their number (bbox6 = 0 when none).  scratch8: 8 device uint32;
 *   grid3, aabb_min3, aabb_max3 are HOST arrays (the renderer's grid_dim / bbox_aabb).
 * clift_upsample_bilinear (tensoRF.py:179-197): F.interpolate(bilinear, align_corners=True) of (1,C,H,W) -> (1,C,H2,W2). */
int32_t clift_dense_alpha(const clift_render_cfg* cfg, const clift_field* field, const float* sx, const float* sy,
                          const float* sz, float* alpha, void* stream);
int32_t clift_alpha_bbox(const float* alpha, const int32_t* grid3, const float* sx, const float* sy, const float* sz,
                         const float* aabb_min3, const float* aabb_max3, float threshold, float* bbox6, int32_t* count,
                         uint32_t* scratch8, void* stream);
int32_t clift_upsample_bilinear(const float* src, float* dst, int32_t channels, int32_t h, int32_t w, int32_t h2, int32_t w2,
                                void* stream);

/* ---- nearest-centroid assignment of rendered embeddings (inference/render_panopli.py:389-396:
 * torch.cdist(p=2) + argmin, ties to the lowest index).  features [n][feature_stride] (first `dim` columns used, dim <= 16),
 * centroids [k][dim]; labels int32 [n]; distances [n] (Euclidean, to the chosen centroid) or null. */
int32_t clift_assign_centroids(const float* features, int64_t n, int32_t dim, int32_t feature_stride, const float* centroids,
                               int32_t k, int32_t* labels, float* distances, void* stream);

/* Panoptic instance labels on the device (the whole of inference/render_panopli.py:371-419 assign_clusters): point p is a
 * "thing" when features[p][0] == -inf (create_instances_from_semantics, render_panopli.py:422-427); its embedding is
 * features[p][1 .. dim]; its class is the first argmax of scores[p][0 .. n_classes); it takes the nearest centroid among
 * centroids[class_first[c] .. + class_count[c]) ([K][dim] table of all classes).  Classes present in the data receive
 * consecutive label ranges in ascending class order, each as long as the highest chosen centroid index + 1 (the reference's
 * `max_label = labels.max() + 1`); labels[p] = 0 for stuff, first_label(class) + centroid + 1 for things.
 * class_scratch: n_classes device uint32 (out: first label of each class); stats2: device int32[2] = {label count including
 * label 0, 1 + highest class id that had points but no centroids (0 = none; the reference raises KeyError there)}.
 * clift_labels_onehot writes the fp64 one-hot rows [n][width] the reference returns. */
int32_t clift_assign_clusters(const float* features, int64_t n, int32_t dim, int32_t feature_stride, const float* scores,
                              int32_t n_classes, const float* centroids, const int32_t* class_first, const int32_t* class_count,
                              int32_t* labels, uint32_t* class_scratch, int32_t* stats2, void* stream);
int32_t clift_labels_onehot(const int32_t* labels, int64_t n, int32_t width, double* onehot, void* stream);

/* ---- the path's one collective.  Sum-all-reduce (in place) of `count` fp32 values - the flat gradient arena of
 * one optimizer's parameters - over the caller's NCCL communicator (an ncclComm_t, passed as void*) on `stream`; what
 * Lightning's DDP does after every manual_backward (trainer/__init__.py:95-108; trainer:198,220).  The mean's 1/N is left to
 * clift_adam_step (grad_scale).  NCCL is not linked: ncclAllReduce is resolved at first call from the libnccl.so.2 already
 * loaded in the process (CLIFT_ERR_UNSUPPORTED when there is none). */
int32_t clift_allreduce_grads(void* nccl_comm, float* arena, int64_t count, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CLIFT_B200_H */
