// Layout packing: every parameter (and, inverted, every gradient) moves between the reference's
// checkpoint layout and the kernels' HBM layout through one tiled transpose.
//   plane (1,C,H,W)  <-> [H*W][C] channel-last            (tensoRF.py:99-106)
//   Linear [out][in] <-> W^T [k_pad][n_pad], zero padded  (k_pad = up16(in), n_pad = up64(out))
#include "launchers.h"

namespace clift {
namespace {

// dst[i][j] (dense d_rows x d_cols) = (j < s_rows && i < s_cols) ? src[j*s_pitch + i] : 0
__global__ void __launch_bounds__(256) transpose_pad_kernel(const float* __restrict__ src, int s_pitch, int s_rows,
                                                            int s_cols, float* __restrict__ dst, int d_rows, int d_cols) {
    __shared__ float tile[32][33];
    const int bi = blockIdx.y * 32, bj = blockIdx.x * 32;   // dst tile origin (row, col)
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int sj = bj + r, si = bi + tx;   // src row = dst col, src col = dst row
        tile[r][tx] = (sj < s_rows && si < s_cols) ? src[(int64_t)sj * s_pitch + si] : 0.0f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int di = bi + r, dj = bj + tx;
        if (di < d_rows && dj < d_cols) dst[(int64_t)di * d_cols + dj] = tile[tx][r];
    }
}

__global__ void copy_pad_kernel(const float* __restrict__ src, int n, float* __restrict__ dst, int n_dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_dst) dst[i] = (src && i < n) ? src[i] : 0.0f;
}

__global__ void copy2d_pad_kernel(const float* __restrict__ src, int s_rows, int s_cols, float* __restrict__ dst, int d_rows,
                                  int d_cols) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d_rows * d_cols) return;
    const int r = i / d_cols, c = i - r * d_cols;
    dst[i] = (r < s_rows && c < s_cols) ? src[(int64_t)r * s_cols + c] : 0.0f;
}

// One launch for a table of transpose / copy jobs (clift_pack_batch): CTA b finds its job by binary search over the jobs'
// first-tile prefix and handles one 32 x 32 destination tile of it.
__global__ void __launch_bounds__(256) pack_batch_kernel(const clift_pack_job* __restrict__ jobs, int n_jobs) {
    __shared__ float tile[32][33];
    int lo = 0, hi = n_jobs - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (jobs[mid].first_tile <= (int)blockIdx.x)
            lo = mid;
        else
            hi = mid - 1;
    }
    const clift_pack_job J = jobs[lo];
    const int t = (int)blockIdx.x - J.first_tile;
    const int tiles_x = (J.d_cols + 31) >> 5;
    const int bi = (t / tiles_x) * 32, bj = (t % tiles_x) * 32;   // dst tile origin (row, col)
    if (bi >= J.d_rows) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    if (J.kind == 0) {
        for (int r = ty; r < 32; r += 8) {
            const int sj = bj + r, si = bi + tx;   // src row = dst col, src col = dst row
            tile[r][tx] = (J.src && sj < J.s_rows && si < J.s_cols) ? J.src[(int64_t)sj * J.s_pitch + si] : 0.0f;
        }
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {
            const int di = bi + r, dj = bj + tx;
            if (di < J.d_rows && dj < J.d_cols) J.dst[(int64_t)di * J.d_cols + dj] = tile[tx][r];
        }
    } else {
        for (int r = ty; r < 32; r += 8) {
            const int di = bi + r, dj = bj + tx;
            if (di < J.d_rows && dj < J.d_cols)
                J.dst[(int64_t)di * J.d_cols + dj] =
                    (J.src && di < J.s_rows && dj < J.s_cols) ? J.src[(int64_t)di * J.s_pitch + dj] : 0.0f;
        }
    }
}

}  // namespace

static int transpose_pad(const float* src, int s_pitch, int s_rows, int s_cols, float* dst, int d_rows, int d_cols,
                         cudaStream_t stream) {
    if (d_rows <= 0 || d_cols <= 0) return CLIFT_OK;
    dim3 grid((unsigned)ceil_div(d_cols, 32), (unsigned)ceil_div(d_rows, 32));
    transpose_pad_kernel<<<grid, 256, 0, stream>>>(src, s_pitch, s_rows, s_cols, dst, d_rows, d_cols);
    CLIFT_AFTER_LAUNCH("transpose_pad_kernel");
    return CLIFT_OK;
}

}  // namespace clift

using namespace clift;

extern "C" int32_t clift_pack_plane(const float* nchw, float* hwc, int32_t comps, int32_t h, int32_t w, void* stream) {
    CLIFT_CHECK_ARG(nchw && hwc && comps > 0 && h > 0 && w > 0, "null pointer or non-positive size");
    return transpose_pad(nchw, h * w, comps, h * w, hwc, h * w, comps, (cudaStream_t)stream);
}

extern "C" int32_t clift_unpack_plane(const float* hwc, float* nchw, int32_t comps, int32_t h, int32_t w, void* stream) {
    CLIFT_CHECK_ARG(nchw && hwc && comps > 0 && h > 0 && w > 0, "null pointer or non-positive size");
    return transpose_pad(hwc, comps, h * w, comps, nchw, comps, h * w, (cudaStream_t)stream);
}

extern "C" int32_t clift_pack_linear(const float* w, const float* b, float* wt, float* bias_pad, int32_t n_out, int32_t n_in,
                                     void* stream) {
    CLIFT_CHECK_ARG(w && wt && n_out > 0 && n_in > 0, "null pointer or non-positive size");
    const int kp = k_pad(n_in), np = n_pad(n_out);
    int rc = transpose_pad(w, n_in, n_out, n_in, wt, kp, np, (cudaStream_t)stream);
    if (rc != CLIFT_OK) return rc;
    if (bias_pad) {
        copy_pad_kernel<<<(unsigned)ceil_div(np, 256), 256, 0, (cudaStream_t)stream>>>(b, n_out, bias_pad, np);
        CLIFT_AFTER_LAUNCH("copy_pad_kernel");
    }
    return CLIFT_OK;
}

extern "C" int32_t clift_unpack_linear(const float* wt, const float* bias_pad, float* w, float* b, int32_t n_out,
                                       int32_t n_in, void* stream) {
    CLIFT_CHECK_ARG(w && wt && n_out > 0 && n_in > 0, "null pointer or non-positive size");
    const int np = n_pad(n_out);
    int rc = transpose_pad(wt, np, n_in, n_out, w, n_out, n_in, (cudaStream_t)stream);
    if (rc != CLIFT_OK) return rc;
    if (b && bias_pad) {
        copy_pad_kernel<<<(unsigned)ceil_div(n_out, 256), 256, 0, (cudaStream_t)stream>>>(bias_pad, n_out, b, n_out);
        CLIFT_AFTER_LAUNCH("copy_pad_kernel");
    }
    return CLIFT_OK;
}

extern "C" int32_t clift_pack_linear_dgrad(const float* w, float* w_dgrad, int32_t n_out, int32_t n_in, void* stream) {
    CLIFT_CHECK_ARG(w && w_dgrad && n_out > 0 && n_in > 0, "null pointer or non-positive size");
    CLIFT_CHECK_SUPPORTED(n_in <= CLIFT_MAX_WIDTH && n_out <= CLIFT_MAX_WIDTH, "layer wider than CLIFT_MAX_WIDTH");
    const int rows = k_pad(n_out), cols = dgrad_pad(n_in);
    copy2d_pad_kernel<<<(unsigned)ceil_div((int64_t)rows * cols, 256), 256, 0, (cudaStream_t)stream>>>(w, n_out, n_in, w_dgrad, rows, cols);
    CLIFT_AFTER_LAUNCH("copy2d_pad_kernel");
    return CLIFT_OK;
}

extern "C" int32_t clift_pack_batch(const clift_pack_job* jobs, int32_t n_jobs, int32_t total_tiles, void* stream) {
    CLIFT_CHECK_ARG(n_jobs >= 0 && total_tiles >= 0 && (n_jobs == 0 || jobs), "null table or negative size");
    if (n_jobs == 0 || total_tiles == 0) return CLIFT_OK;
    pack_batch_kernel<<<(unsigned)total_tiles, 256, 0, (cudaStream_t)stream>>>(jobs, n_jobs);
    CLIFT_AFTER_LAUNCH("pack_batch_kernel");
    return CLIFT_OK;
}
