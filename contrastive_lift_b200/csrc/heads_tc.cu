// Tensor-core (tcgen05 / TMEM) implementation of the MLP-head GEMMs (tensoRF.py:383-418, 462-511, 565-594).
//
// fp32-faithful arithmetic on the tf32 tensor pipe: every operand x is split into tf32-exact hi + lo and
//     A*W ~= A_hi*W_hi + A_hi*W_lo + A_lo*W_hi        (error ~2^-21 per product, fp32 accumulation in TMEM)
// A_hi lives in shared memory ([K/4][128][4] floats, canonical K-major/no-swizzle UMMA layout), A_lo in tensor memory
// (128 lanes x K columns, read by the .ts form of tcgen05.mma), the accumulator D in tensor memory (128 lanes x N
// columns), and the pre-split weights stream through a bulk-TMA + mbarrier ring in 16-row K slabs.
//
// Roles (192 threads): warp 0 weight producer, warp 1 MMA issuer (one elected lane), warps 2-5 = 128 "row" threads
// (thread <-> record <-> TMEM lane) that build layer inputs, run epilogues (bias, ReLU, split, write next A) and reduce.
#include "launchers.h"
#include "tcgen05.cuh"

namespace clift {
namespace {

constexpr int kTcThreads = 192;
constexpr int kTcRows = 128;            // records per tile = UMMA M
constexpr int kTcMaxK = 256;
constexpr int kTcStages = 2;
constexpr int kTcSlabK = 16;            // K rows per weight stage
constexpr int kTcStageFloats = 2 * kTcSlabK * 256;   // hi + lo, N up to 256
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kTmemALo = 256;      // column offset of A_lo

struct TcSmem {
    float* a_hi;        // [kTcMaxK/4][128][4]
    float* w;           // [kTcStages][kTcStageFloats]
    uint64_t* full;     // [kTcStages]
    uint64_t* empty;    // [kTcStages]
    uint64_t* bar_a;    // A operand ready (row threads -> MMA)
    uint64_t* bar_d;    // accumulator ready (MMA -> row threads)
    uint32_t* tmem_base;
};

constexpr size_t kTcSmemBytes = (size_t)kTcMaxK * kTcRows * 4 + (size_t)kTcStages * kTcStageFloats * 4 + 256;

__device__ __forceinline__ TcSmem carve_tc_smem(unsigned char* raw) {
    TcSmem s;
    s.a_hi = reinterpret_cast<float*>(raw);
    s.w = s.a_hi + kTcMaxK * kTcRows;
    s.full = reinterpret_cast<uint64_t*>(s.w + kTcStages * kTcStageFloats);
    s.empty = s.full + kTcStages;
    s.bar_a = s.empty + kTcStages;
    s.bar_d = s.bar_a + 1;
    s.tmem_base = reinterpret_cast<uint32_t*>(s.bar_d + 1);
    return s;
}

// One GEMM of the schedule: D[128 x n_pad] = A[128 x k_pad] * W^T, weights packed by clift_pack_linear_tc.
struct TcGemm {
    const float* w;     // [slabs][2][4][n_pad][4]
    int k_steps;        // ceil(K / 8)
    int n_pad;          // multiple of 32, <= 256
};

struct PipeState {      // ring position shared by construction between producer and MMA warps
    uint32_t slab = 0;  // running slab counter over the whole kernel
    __device__ __forceinline__ int stage() const { return slab % kTcStages; }
    __device__ __forceinline__ uint32_t phase() const { return (slab / kTcStages) & 1; }
};

// warp 0, one lane: stream the GEMM's weight slabs
__device__ __forceinline__ void tc_produce(const TcSmem& s, const TcGemm& g, PipeState& ps) {
    const int slabs = (g.k_steps + 1) / 2;
    const uint32_t bytes = 2u * kTcSlabK * g.n_pad * 4u;
    for (int i = 0; i < slabs; ++i, ++ps.slab) {
        tc::mbar_wait(&s.empty[ps.stage()], ps.phase() ^ 1);
        tc::mbar_arrive_expect_tx(&s.full[ps.stage()], bytes);
        tc::bulk_load(s.w + (size_t)ps.stage() * kTcStageFloats, g.w + (size_t)i * (2 * kTcSlabK * g.n_pad), bytes, &s.full[ps.stage()]);
    }
}

// warp 1, one lane: issue the 3xTF32 MMAs of one GEMM
__device__ __forceinline__ void tc_issue(const TcSmem& s, const TcGemm& g, PipeState& ps, uint32_t tmem, uint32_t a_parity) {
    const uint32_t idesc = tc::make_idesc_tf32(kTcRows, g.n_pad);
    const uint32_t a_base = tc::smem_addr(s.a_hi);
    const uint32_t w_lbo = (uint32_t)g.n_pad * 16u;
    tc::mbar_wait(s.bar_a, a_parity);
    tc::fence_after_sync();
    const int slabs = (g.k_steps + 1) / 2;
    for (int i = 0; i < slabs; ++i, ++ps.slab) {
        tc::mbar_wait(&s.full[ps.stage()], ps.phase());
        tc::fence_after_sync();
        const uint32_t w_hi = tc::smem_addr(s.w + (size_t)ps.stage() * kTcStageFloats);
        const uint32_t w_lo = w_hi + (uint32_t)kTcSlabK * g.n_pad * 4u;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
            const int ks = i * 2 + kk;
            if (ks >= g.k_steps) break;
            const uint64_t a_desc = tc::make_smem_desc(a_base + (uint32_t)ks * 2u * (kTcRows * 16u), kTcRows * 16u, 128u);
            const uint64_t bh = tc::make_smem_desc(w_hi + (uint32_t)kk * 2u * w_lbo, w_lbo, 128u);
            const uint64_t bl = tc::make_smem_desc(w_lo + (uint32_t)kk * 2u * w_lbo, w_lbo, 128u);
            tc::mma_ss(tmem, a_desc, bh, idesc, ks > 0 ? 1u : 0u);
            tc::mma_ss(tmem, a_desc, bl, idesc, 1u);
            tc::mma_ts(tmem, tmem + kTmemALo + (uint32_t)ks * 8u, bh, idesc, 1u);
        }
        tc::mma_commit(&s.empty[ps.stage()]);
    }
    tc::mma_commit(s.bar_d);
}

// row thread: write 8 consecutive K values (k0 multiple of 8) of its record into the A operand (hi -> smem, lo -> TMEM)
__device__ __forceinline__ void tc_put8(const TcSmem& s, uint32_t tmem_lane_base, int row, int k0, const float* v) {
    float hi[8], lo[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) tc::split_tf32(v[i], hi[i], lo[i]);
    float4* dst = reinterpret_cast<float4*>(s.a_hi + ((size_t)(k0 >> 2) * kTcRows + row) * 4);
    dst[0] = make_float4(hi[0], hi[1], hi[2], hi[3]);
    dst[kTcRows] = make_float4(hi[4], hi[5], hi[6], hi[7]);
    tc::tmem_st8(tmem_lane_base + kTmemALo + (uint32_t)k0, lo);
}

// row threads: publish the A operand they just wrote
__device__ __forceinline__ void tc_publish_a(const TcSmem& s) {
    tc::tmem_wait_st();
    tc::fence_proxy_async_smem();
    tc::fence_before_sync();
    tc::named_bar_sync(1, kTcRows);
    if ((threadIdx.x & 127) == 64) tc::mbar_arrive(s.bar_a);   // any single row thread
}

// ---------------------------------------------------------------------------------------------------------
// parity / bring-up kernel: out[128][n_pad] = a[128][K] * W^T (no bias), one CTA
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTcThreads, 1) tc_gemm_test_kernel(const float* __restrict__ a, int K, TcGemm g, float* __restrict__ out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    TcSmem s = carve_tc_smem(smem_raw);
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kTcStages; ++i) {
            tc::mbar_init(&s.full[i], 1);
            tc::mbar_init(&s.empty[i], 1);
        }
        tc::mbar_init(s.bar_a, 1);
        tc::mbar_init(s.bar_d, 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(s.tmem_base, kTmemCols);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *s.tmem_base;

    if (warp == 0) {
        if (tc::elect_one()) {
            PipeState ps;
            tc_produce(s, g, ps);
        }
    } else if (warp == 1) {
        if (tc::elect_one()) {
            PipeState ps;
            tc_issue(s, g, ps, tmem, 0);
        }
    } else {
        const int quarter = warp & 3;                        // TMEM lane quarter this warp may touch
        const int row = quarter * 32 + (threadIdx.x & 31);
        const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16);
        for (int k0 = 0; k0 < g.k_steps * 8; k0 += 8) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = (k0 + i < K) ? a[(size_t)row * K + k0 + i] : 0.0f;
            tc_put8(s, lane_base, row, k0, v);
        }
        tc_publish_a(s);
        tc::mbar_wait(s.bar_d, 0);
        tc::fence_after_sync();
        for (int c0 = 0; c0 < g.n_pad; c0 += 16) {
            float v[16];
            tc::tmem_ld16(lane_base + (uint32_t)c0, v);
            tc::tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) out[(size_t)row * g.n_pad + c0 + i] = v[i];
        }
        tc::fence_before_sync();
    }
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem, kTmemCols);
}

// W [out][in] -> [slab][hi|lo][4 k-chunks][n_pad][4]  (zero padded; hi/lo tf32-exact)
__global__ void pack_linear_tc_kernel(const float* __restrict__ w, int n_out, int n_in, float* __restrict__ dst, int n_pad, int slabs) {
    const int64_t total = (int64_t)slabs * kTcSlabK * n_pad;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int k = (int)(idx / n_pad), n = (int)(idx % n_pad);
    const float x = (k < n_in && n < n_out) ? w[(size_t)n * n_in + k] : 0.0f;
    float hi, lo;
    tc::split_tf32(x, hi, lo);
    const int slab = k / kTcSlabK, kc = (k % kTcSlabK) / 4, ki = k & 3;
    float* base = dst + (size_t)slab * (2 * kTcSlabK * n_pad);
    const size_t off = ((size_t)kc * n_pad + n) * 4 + ki;
    base[off] = hi;
    base[(size_t)kTcSlabK * n_pad + off] = lo;
}

}  // namespace
}  // namespace clift

using namespace clift;

extern "C" int64_t clift_tc_weight_floats(int32_t n_out, int32_t n_in) {
    if (n_out <= 0 || n_in <= 0 || n_out > 256 || n_in > kTcMaxK) return CLIFT_ERR_UNSUPPORTED;
    const int n_pad = (int)round_up(n_out, 32), slabs = (int)ceil_div(n_in, kTcSlabK);
    return (int64_t)slabs * 2 * kTcSlabK * n_pad;
}

extern "C" int32_t clift_pack_linear_tc(const float* w, float* dst, int32_t n_out, int32_t n_in, void* stream) {
    CLIFT_CHECK_ARG(w && dst, "null pointer");
    CLIFT_CHECK_SUPPORTED(n_out > 0 && n_in > 0 && n_out <= 256 && n_in <= kTcMaxK, "layer wider than 256");
    const int n_pad = (int)round_up(n_out, 32), slabs = (int)ceil_div(n_in, kTcSlabK);
    const int64_t total = (int64_t)slabs * kTcSlabK * n_pad;
    pack_linear_tc_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(w, n_out, n_in, dst, n_pad, slabs);
    CLIFT_AFTER_LAUNCH("pack_linear_tc_kernel");
    return CLIFT_OK;
}

extern "C" int32_t clift_debug_tc_gemm(const float* a, const float* w_tc, float* out, int32_t k, int32_t n_out, void* stream) {
    CLIFT_CHECK_ARG(a && w_tc && out, "null pointer");
    CLIFT_CHECK_SUPPORTED(n_out > 0 && k > 0 && n_out <= 256 && k <= kTcMaxK, "layer wider than 256");
    TcGemm g;
    g.w = w_tc;
    g.k_steps = (int)ceil_div(k, 8);
    g.n_pad = (int)round_up(n_out, 32);
    CLIFT_CUDA(cudaFuncSetAttribute(tc_gemm_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes));
    tc_gemm_test_kernel<<<1, kTcThreads, kTcSmemBytes, (cudaStream_t)stream>>>(a, k, g, out);
    CLIFT_AFTER_LAUNCH("tc_gemm_test_kernel");
    return CLIFT_OK;
}
