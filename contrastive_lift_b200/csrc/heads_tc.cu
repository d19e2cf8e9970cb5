// Tensor-core (tcgen05 / TMEM) implementation of the MLP heads on the compacted active samples
// (tensoRF.py:127-137, 383-418, 462-511, 565-594; renderer:103-131, 137-156) - the inference path.
//
// fp32-faithful arithmetic on the tf32 tensor pipe: every operand x is split into a tf32-exact hi part and the
// remainder lo = x - hi, and
//     A*W ~= A_hi*W_hi + A_hi*W_lo + A_lo*W_hi        (error ~2^-21 per product, fp32 accumulation in tensor memory)
// measured 1e-6..3e-6 relative against fp64 (tests/test_gpu_tc.py), i.e. inside the 1e-4 budget with two orders to spare.
//
// Per CTA (one per SM, persistent over 128-record tiles):
//   shared memory : A_hi [K/4][128][4] fp32 (128 KB, canonical K-major / no-swizzle UMMA layout: 8-row x 16-byte core
//                   matrices, SBO 128 B, LBO 2 KB), a 5-stage x 16 KB weight ring fed by 1-D bulk TMA + mbarriers, a
//                   constant "ones" operand chunk, record positions / ray ids / run table
//   tensor memory : D = 128 lanes x 256 columns (accumulator), A_lo = 128 lanes x 256 columns (read by the .ts MMA form)
//   warp 0        : weight producer (one lane)       warp 1 : MMA issuer (one lane, running descriptors)
//   warps 2..13   : 384 "row" threads, three per record (thread <-> TMEM lane quarter of its warp): build layer inputs,
//                   run epilogues (TMEM -> ReLU -> hi to smem / lo to TMEM), softmax / sigmoid, per-ray run sums
// One tile runs 17 dependent GEMMs (semantic 5, instance fast 4 + slow 4, basis 1, rgb 3).  Choices that the
// clock64 traces under profiles/ drove:
//   * the bias is one more MMA k-step against the ones chunk instead of an add per accumulator element;
//   * for N <= 128 the hi and lo weight blocks are one stacked B operand (one MMA, epilogue adds the two column blocks):
//     small-N MMAs are latency-bound (~76 cycles each), not N-bound;
//   * ring stages carry several k-steps for small N, so short GEMMs are not bulk-copy round-trip bound;
//   * the appearance gather runs in quad layout (64-byte coalesced texel segments), writes raw products into the A rows,
//     and the row threads split them in place; the rgb MLP input is built the same way from one SFU sincos per base value.
// MMA and epilogue of a tile are serial by data dependence (one A buffer fills the SM), so the tensor pipe is busy
// ~55-60 % of the kernel; DESIGN.md has the cycle budget.
#include "launchers.h"
#include "tcgen05.cuh"

namespace clift {
namespace {

constexpr int kTcRows = 128;            // records per tile = UMMA M
constexpr int kTcParts = 3;             // threads per record (each takes every kTcParts-th column chunk)
constexpr int kTcRowThreads = kTcRows * kTcParts;
constexpr int kTcThreads = 64 + kTcRowThreads;
constexpr int kTcMaxK = 256;
constexpr int kTcStages = 5;
constexpr int kTcSlabK = 8;             // K rows per weight stage = one tcgen05.mma k-step
constexpr int kTcStageFloats = 2 * kTcSlabK * 256;   // hi + lo, N up to 256
constexpr int kTcMaxGemms = 24;
constexpr int kTcOnesFloats = 2 * kTcRows * 4;   // A chunk [2 k-chunks][128][4] with column 0 = 1 (bias k-step)
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kTmemALo = 256;      // column offset of A_lo

struct TcSmem {
    float* a_hi;        // [kTcMaxK/4][128][4]   (also the [c][128] scratch of the final epilogues)
    float* w;           // [kTcStages][kTcStageFloats]
    float* ones;        // [kTcOnesFloats] constant A operand of the bias k-step: A[m][0] = 1, A[m][1..7] = 0
    float4* pos;        // [128] record positions (x, y, z normalised, w) of the tile
    int* ray;           // [128]
    int* runs;          // [129]
    int* n_runs;        // [1]
    uint64_t* full;     // [kTcStages]
    uint64_t* empty;    // [kTcStages]
    uint64_t* bar_a;    // A operand ready (row threads -> MMA)
    uint64_t* bar_d;    // accumulator ready (MMA -> row threads)
    uint32_t* tmem_base;
};

constexpr size_t kTcSmemBytes = (size_t)kTcMaxK * kTcRows * 4 + (size_t)kTcStages * kTcStageFloats * 4 +
                                (size_t)kTcOnesFloats * 4 + (size_t)kTcRows * 16 + (2 * kTcRows + 8) * 4 + 256;
static_assert(kTcSmemBytes <= 232448, "shared memory budget of one sm_100a CTA");

__device__ __forceinline__ TcSmem carve_tc_smem(unsigned char* raw) {
    TcSmem s;
    s.a_hi = reinterpret_cast<float*>(raw);
    s.w = s.a_hi + kTcMaxK * kTcRows;
    s.ones = s.w + kTcStages * kTcStageFloats;
    s.pos = reinterpret_cast<float4*>(s.ones + kTcOnesFloats);
    s.ray = reinterpret_cast<int*>(s.pos + kTcRows);
    s.runs = s.ray + kTcRows;
    s.n_runs = s.runs + kTcRows + 1;
    s.full = reinterpret_cast<uint64_t*>(s.n_runs + 7);
    s.empty = s.full + kTcStages;
    s.bar_a = s.empty + kTcStages;
    s.bar_d = s.bar_a + 1;
    s.tmem_base = reinterpret_cast<uint32_t*>(s.bar_d + 1);
    return s;
}

// One GEMM of the schedule: D[128 x n_pad] = A[128 x 8*k_steps] * W^T, weights packed by clift_pack_linear_tc.
struct TcGemm {
    const float* w;     // [k_steps] slabs, see tc_issue for the two slab layouts
    int k_steps;        // ceil(K / 8)
    int n_pad;          // multiple of 32, <= 256
    int has_bias;       // 1: one more slab follows whose k row 0 is the bias; its A operand is the constant ones chunk,
                        //    so the bias add is one extra MMA k-step instead of an add per accumulator element
};

struct PipeState {      // ring position; producer and MMA warps advance it identically
    int stage = 0;
    uint32_t phase = 0;
    __device__ __forceinline__ void advance() {
        if (++stage == kTcStages) {
            stage = 0;
            phase ^= 1u;
        }
    }
};

// k-steps carried by one ring stage: small-N GEMMs pack several (their per-k-step slab is tiny, and one bulk copy +
// barrier round trip per 2 KB would leave them latency-bound)
__device__ __forceinline__ int tc_ksteps_per_stage(int n_pad) { return n_pad >= 256 ? 1 : 256 / n_pad; }

// warp 0, one lane: stream the GEMM's weight slabs
__device__ __forceinline__ void tc_produce(const TcSmem& s, const TcGemm& g, PipeState& ps) {
    const int per = tc_ksteps_per_stage(g.n_pad);
    const uint32_t kstep_floats = 2u * kTcSlabK * g.n_pad;
    const int steps = g.k_steps + g.has_bias;
    for (int k0 = 0; k0 < steps; k0 += per, ps.advance()) {
        const uint32_t bytes = (uint32_t)min(per, steps - k0) * kstep_floats * 4u;
        tc::mbar_wait(&s.empty[ps.stage], ps.phase ^ 1);
        tc::mbar_arrive_expect_tx(&s.full[ps.stage], bytes);
        tc::bulk_load(s.w + (size_t)ps.stage * kTcStageFloats, g.w + (size_t)k0 * kstep_floats, bytes, &s.full[ps.stage]);
    }
}

// warp 1, one lane: issue the 3xTF32 MMAs of one GEMM.  The issuing thread has ~384 cycles per k-step before it, not
// the tensor pipe, becomes the limiter, so descriptors are running values updated by adds (low word = address >> 4).
// n_pad > 128 : slab = [hi | lo][2 k-chunks][n_pad][4], three MMAs per k-step (A_hi*W_hi, A_hi*W_lo, A_lo*W_hi)
// n_pad <= 128: slab = [2 k-chunks][hi | lo][n_pad][4]: one descriptor with 2*n_pad rows covers [W_hi ; W_lo], so
//               A_hi*W_hi and A_hi*W_lo are ONE MMA writing D[:, 0:n_pad) and D[:, n_pad:2n_pad) (the epilogue adds
//               the two column blocks) - small-N MMAs are latency-bound, not N-bound.
// bias        : the slab after the last K slab carries the bias in its k row 0; its A operand is the constant ones
//               chunk (exact in tf32, so no A_lo term).
__device__ __forceinline__ void tc_issue(const TcSmem& s, const TcGemm& g, PipeState& ps, uint32_t tmem, uint32_t a_parity,
                                         long long* trace = nullptr) {
    const uint32_t idesc = tc::make_idesc_tf32(kTcRows, g.n_pad);
    const uint32_t rows16 = (uint32_t)g.n_pad;                          // n_pad*16 bytes >> 4
    const bool stacked = g.n_pad <= 128;
    const uint32_t idesc_ss = stacked ? tc::make_idesc_tf32(kTcRows, 2 * g.n_pad) : idesc;
    const int per = tc_ksteps_per_stage(g.n_pad);
    const uint32_t kstep16 = 2u * kTcSlabK * g.n_pad * 4u >> 4;         // bytes of one k-step slab >> 4
    constexpr uint32_t kStage16 = kTcStageFloats * 4u >> 4;
    // descriptor words: high = SBO (128 B) | version 1 ; low = address>>4 | LBO<<16
    constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);
    const uint32_t a_lbo = (kTcRows * 16u >> 4) << 16;
    const uint32_t b_lbo = (stacked ? 2u * rows16 : rows16) << 16;
    uint32_t a_lo = (tc::smem_addr(s.a_hi) >> 4) | a_lbo;
    const uint32_t ones_lo = (tc::smem_addr(s.ones) >> 4) | a_lbo;
    const uint32_t w_lo0 = (tc::smem_addr(s.w) >> 4) | b_lbo;
    const uint32_t lo_off = stacked ? 0u : 2u * rows16;                 // W_lo block inside a non-stacked slab
    uint32_t a_tmem = tmem + kTmemALo;
    auto desc = [](uint32_t lo) { return ((uint64_t)kDescHi << 32) | lo; };

    tc::mbar_wait(s.bar_a, a_parity);
    tc::fence_after_sync();
    if (trace) trace[3] = clock64();
    const int steps = g.k_steps + g.has_bias;
    uint32_t acc = 0u;
    for (int k0 = 0; k0 < steps; k0 += per, ps.advance()) {
        tc::mbar_wait(&s.full[ps.stage], ps.phase);
        tc::fence_after_sync();
        if (trace && k0 == 0) trace[5] = clock64();
        uint32_t w_lo = w_lo0 + (uint32_t)ps.stage * kStage16;
        const int n_here = min(per, steps - k0);
        for (int j = 0; j < n_here; ++j, w_lo += kstep16) {
            const bool bias_step = k0 + j >= g.k_steps;
            const uint64_t a_desc = desc(bias_step ? ones_lo : a_lo);
            const uint64_t bh = desc(w_lo);
            tc::mma_ss(tmem, a_desc, bh, idesc_ss, acc);
            acc = 1u;
            if (!stacked) tc::mma_ss(tmem, a_desc, desc(w_lo + lo_off), idesc, 1u);
            if (!bias_step) tc::mma_ts(tmem, a_tmem, bh, idesc, 1u);
            a_lo += 2u * (kTcRows * 16u >> 4);
            a_tmem += 8u;
        }
        tc::mma_commit(&s.empty[ps.stage]);
    }
    tc::mma_commit(s.bar_d);
    if (trace) trace[4] = clock64();
}

// row thread: write 8 consecutive K values (k0 multiple of 8) of its record into the A operand (hi -> smem, lo -> TMEM)
__device__ __forceinline__ void tc_put8(const TcSmem& s, uint32_t lane_base, int row, int k0, const float* v) {
    float hi[8], lo[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {   // hi is tf32-exact; lo = v - hi is exact in fp32 and is read as tf32 by the tensor core
        hi[i] = __uint_as_float(__float_as_uint(v[i]) & 0xffffe000u);
        lo[i] = v[i] - hi[i];
    }
    float4* dst = reinterpret_cast<float4*>(s.a_hi + ((size_t)(k0 >> 2) * kTcRows + row) * 4);
    dst[0] = make_float4(hi[0], hi[1], hi[2], hi[3]);
    dst[kTcRows] = make_float4(hi[4], hi[5], hi[6], hi[7]);
    tc::tmem_st8(lane_base + kTmemALo + (uint32_t)k0, lo);
}

// row threads: the A operand region holds RAW fp32 values for K rows [0, k_rows) (written by any thread); convert them in
// place to the tf32-exact hi part and move the lo part to tensor memory.  A record's threads alternate over 8-row chunks.
__device__ __forceinline__ void tc_split_rows(const TcSmem& s, uint32_t lane_base, int row, int part, int k_rows) {
    for (int k0 = part * 8; k0 < k_rows; k0 += 8 * kTcParts) {
        const float4* src = reinterpret_cast<const float4*>(s.a_hi + ((size_t)(k0 >> 2) * kTcRows + row) * 4);
        const float4 a = src[0], b = src[kTcRows];
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        tc_put8(s, lane_base, row, k0, v);
    }
}

// row threads: publish the A operand they just wrote
__device__ __forceinline__ void tc_publish_a(const TcSmem& s) {
    tc::tmem_wait_st();
    tc::fence_proxy_async_smem();
    tc::fence_before_sync();
    tc::named_bar_sync(1, kTcRowThreads);
    if (threadIdx.x == 64) tc::mbar_arrive(s.bar_a);
}

struct RowId {
    int row, half, rt;          // record row 0..127, part 0..kTcParts-1 of the record's threads, index among the row threads
    uint32_t lane_base;         // TMEM address of this warp's lane quarter
};

__device__ __forceinline__ RowId make_row_id(uint32_t tmem) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    RowId r;
    const int quarter = warp & 3;           // the TMEM lane quarter a warp may touch is warp_id % 4
    r.row = quarter * 32 + lane;
    r.half = (warp - 2) >> 2;               // warps 2..5 part 0, 6..9 part 1, ...
    r.rt = threadIdx.x - 64;
    r.lane_base = tmem + ((uint32_t)(quarter * 32) << 16);
    return r;
}

// 16 accumulator columns of this thread's record; stacked GEMMs (n_pad <= 128) keep A_hi*W_lo in columns [n_pad, 2 n_pad)
__device__ __forceinline__ void tc_ld_acc16(const RowId& r, int c0, int n_pad, float* v) {
    tc::tmem_ld16(r.lane_base + (uint32_t)c0, v);
    if (n_pad <= 128) {
        float u[16];
        tc::tmem_ld16(r.lane_base + (uint32_t)(n_pad + c0), u);
        tc::tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += u[i];
    } else {
        tc::tmem_wait_ld();
    }
}

__device__ __forceinline__ void tc_init(const TcSmem& s) {
    for (int i = threadIdx.x; i < kTcOnesFloats; i += kTcThreads) s.ones[i] = (i < kTcRows * 4 && (i & 3) == 0) ? 1.0f : 0.0f;
    tc::fence_proxy_async_smem();
    if (threadIdx.x == 0) {
        for (int i = 0; i < kTcStages; ++i) {
            tc::mbar_init(&s.full[i], 1);
            tc::mbar_init(&s.empty[i], 1);
        }
        tc::mbar_init(s.bar_a, 1);
        tc::mbar_init(s.bar_d, 1);
        tc::fence_barrier_init();
    }
    if ((threadIdx.x >> 5) == 1) tc::tmem_alloc(s.tmem_base, kTmemCols);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
}

// ---------------------------------------------------------------------------------------------------------
// parity / bring-up kernel: out[128][n_pad] = a[128][K] * W^T (no bias), one CTA
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTcThreads, 1) tc_gemm_test_kernel(const float* __restrict__ a, int K, TcGemm g, float* __restrict__ out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    TcSmem s = carve_tc_smem(smem_raw);
    const int warp = threadIdx.x >> 5;
    tc_init(s);
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
        const RowId r = make_row_id(tmem);
        for (int k0 = r.half * 8; k0 < g.k_steps * 8; k0 += 8 * kTcParts) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = (k0 + i < K) ? a[(size_t)r.row * K + k0 + i] : 0.0f;
            tc_put8(s, r.lane_base, r.row, k0, v);
        }
        tc_publish_a(s);
        tc::mbar_wait(s.bar_d, 0);
        tc::fence_after_sync();
        for (int c0 = r.half * 16; c0 < g.n_pad; c0 += 16 * kTcParts) {
            float v[16];
            tc_ld_acc16(r, c0, g.n_pad, v);
#pragma unroll
            for (int i = 0; i < 16; ++i) out[(size_t)r.row * g.n_pad + c0 + i] = v[i];
        }
        tc::fence_before_sync();
    }
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem, kTmemCols);
}

// ---------------------------------------------------------------------------------------------------------
// production kernel: all heads of one 128-record tile per iteration, persistent over tiles
// ---------------------------------------------------------------------------------------------------------
struct TcHeadsParams {
    const float4* rec_pos;
    const int32_t* rec_ray;
    const unsigned long long* stats;
    long long cap;
    const float* rays;
    FactorParams app;
    int dim_app, pe_view, pe_feat, pe_sem, pe_ins;
    int n_cls, d_ins, slow_fast, softmax, heads;
    // GEMM schedule of one tile, in issue order: semantic | instance fast | instance slow | basis | rgb
    int n_gemms;
    TcGemm g[kTcMaxGemms];
    int n_sem, n_ins, n_rgb;          // layers per stack (0 = head off)
    float* rgb_raw;
    float* sem_raw;
    float* ins;
    long long* trace;                 // debug: [4 tiles][kTcMaxGemms][10] clock64 stamps of CTA 0, or null
};

__device__ __forceinline__ void tc_stamp(const TcHeadsParams& P, long long tile_local, int gi, int slot) {
    if (P.trace && blockIdx.x == 0 && tile_local < 4) P.trace[(tile_local * kTcMaxGemms + gi) * 10 + slot] = clock64();
}

// hidden-layer epilogue: D (bias already accumulated by the bias k-step) -> ReLU -> next layer's A operand.  A record's
// kTcParts threads take the 16-column chunks round-robin; the next TMEM load is in flight while a chunk is processed.
__device__ __forceinline__ void tc_relu_put16(const TcSmem& s, const RowId& r, int c0, float* v) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.0f);
    tc_put8(s, r.lane_base, r.row, c0, v);
    tc_put8(s, r.lane_base, r.row, c0 + 8, v + 8);
}

__device__ __forceinline__ void tc_epilogue_hidden(const TcSmem& s, const RowId& r, int n_pad) {
    constexpr int kStride = 16 * kTcParts;
    const int c_begin = r.half * 16;
    if (c_begin >= n_pad) return;
    if (n_pad <= 128) {   // stacked accumulator: columns [n_pad, 2 n_pad) hold A_hi * W_lo
        for (int c0 = c_begin; c0 < n_pad; c0 += kStride) {
            float v[16];
            tc_ld_acc16(r, c0, n_pad, v);
            tc_relu_put16(s, r, c0, v);
        }
        return;
    }
    float a[16], b[16];
    tc::tmem_ld16(r.lane_base + (uint32_t)c_begin, a);
    for (int c0 = c_begin; c0 < n_pad; c0 += 2 * kStride) {
        tc::tmem_wait_ld();
        if (c0 + kStride < n_pad) tc::tmem_ld16(r.lane_base + (uint32_t)(c0 + kStride), b);
        tc_relu_put16(s, r, c0, a);
        if (c0 + kStride < n_pad) {
            tc::tmem_wait_ld();
            if (c0 + 2 * kStride < n_pad) tc::tmem_ld16(r.lane_base + (uint32_t)(c0 + 2 * kStride), a);
            tc_relu_put16(s, r, c0 + kStride, b);
        }
    }
}

// final-layer epilogue (part 0 threads): D -> scratch[c][row] for c < n_out (scratch = a_hi region)
__device__ __forceinline__ void tc_epilogue_final(const TcSmem& s, const RowId& r, int n_out, int n_pad) {
    if (r.half != 0) return;
    for (int c0 = 0; c0 < n_out; c0 += 16) {
        float v[16];
        tc_ld_acc16(r, c0, n_pad, v);
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (c0 + i < n_out) s.a_hi[(size_t)(c0 + i) * kTcRows + r.row] = v[i];
    }
}

// semantic final layer for n_cls <= 32: logits stay in registers -> softmax -> * w -> scratch[c][row]
__device__ __forceinline__ void tc_epilogue_semantic32(const TcSmem& s, const RowId& r, int n_cls, int n_pad, int softmax, float w) {
    if (r.half != 0) return;
    float v[32];
    tc_ld_acc16(r, 0, n_pad, v);
    if (n_cls > 16) {
        tc_ld_acc16(r, 16, n_pad, v + 16);
    } else {
#pragma unroll
        for (int i = 16; i < 32; ++i) v[i] = 0.0f;
    }
    if (softmax) {
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (i < n_cls) mx = fmaxf(mx, v[i]);
        float tot = 0.0f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            v[i] = i < n_cls ? expf(v[i] - mx) : 0.0f;
            tot += v[i];
        }
        const float sc = w / tot;            // one division per record (the FMA kernel divides per class: 1 ulp apart)
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= sc;
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= w;
    }
#pragma unroll
    for (int i = 0; i < 32; ++i)
        if (i < n_cls) s.a_hi[(size_t)i * kTcRows + r.row] = v[i];
}

// xyz (+ sin/cos PE, dimension-major frequency-minor) -> A operand, zero padded to a multiple of 8
__device__ __forceinline__ void tc_build_xyz(const TcSmem& s, const RowId& r, const float4& p, int pe) {
    if (pe == 0) {   // the shipped configuration (pe_sem = pe_ins = 0): one chunk, no decode
        if (r.half == 0) {
            const float v[8] = {p.x, p.y, p.z, 0.f, 0.f, 0.f, 0.f, 0.f};
            tc_put8(s, r.lane_base, r.row, 0, v);
        }
        return;
    }
    const int n_in = 3 + 6 * pe;
    const float xyz[3] = {p.x, p.y, p.z};
    for (int k0 = r.half * 8; k0 < n_in; k0 += 8 * kTcParts) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int q = k0 + i;
            float x = 0.0f;
            if (q < 3) {
                x = q == 0 ? xyz[0] : (q == 1 ? xyz[1] : xyz[2]);
            } else if (q < n_in) {
                const int j = (q - 3) % (3 * pe);
                const int d = j / pe;
                const float arg = (d == 0 ? xyz[0] : (d == 1 ? xyz[1] : xyz[2])) * (float)(1 << (j % pe));
                x = (q - 3 < 3 * pe) ? sinf(arg) : cosf(arg);
            }
            v[i] = x;
        }
        tc_put8(s, r.lane_base, r.row, k0, v);
    }
}

// row threads: sum rows [0,nch) of scratch over each ray run and add into dst[ray*stride + col0 + c].  Four adjacent
// lanes share one (run, channel) item: each sums a contiguous quarter of the run, then a fixed-order shuffle tree
// combines them (bit-reproducible), so the dependent-add chain is a quarter of the run length.
__device__ __forceinline__ void tc_reduce_runs(const TcSmem& s, int rt, int nch, float* __restrict__ dst, int stride, int col0) {
    tc::named_bar_sync(1, kTcRowThreads);
    const int n_runs = *s.n_runs;
    const int items = n_runs * nch, sub = rt & 3;
    for (int base = 0; base < items; base += kTcRowThreads / 4) {
        const int idx = base + (rt >> 2);
        float acc = 0.0f;
        int ray = 0, c = 0;
        if (idx < items) {
            const int rr = idx / nch;
            c = idx - rr * nch;
            const int m0 = s.runs[rr], m1 = s.runs[rr + 1];
            const int len = m1 - m0, q = (len + 3) >> 2;
            const int a = m0 + min(sub * q, len), b = m0 + min((sub + 1) * q, len);
            for (int m = a; m < b; ++m) acc += s.a_hi[(size_t)c * kTcRows + m];
            ray = s.ray[m0];
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        if (idx < items && sub == 0) atomicAdd(dst + (int64_t)ray * stride + col0 + c, acc);
    }
    tc::named_bar_sync(1, kTcRowThreads);
}

template <int NV>
__global__ void __launch_bounds__(kTcThreads, 1) heads_tc_forward_kernel(const __grid_constant__ TcHeadsParams P) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    TcSmem s = carve_tc_smem(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    tc_init(s);
    const uint32_t tmem = *s.tmem_base;
    const long long n_act = min((long long)P.stats[0], P.cap);
    const long long n_tiles = (n_act + kTcRows - 1) / kTcRows;
    // (Measured: offsetting CTAs in time so their appearance gathers do not coincide is SLOWER - a gather then competes
    // with the other CTAs' weight streams for L2; in lockstep nobody streams weights while everybody gathers.)

    if (warp == 0) {
        if (tc::elect_one()) {
            PipeState ps;
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
                for (int gi = 0; gi < P.n_gemms; ++gi) tc_produce(s, P.g[gi], ps);
        }
    } else if (warp == 1) {
        if (tc::elect_one()) {
            PipeState ps;
            uint32_t count = 0;
            long long tl = 0;
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl)
                for (int gi = 0; gi < P.n_gemms; ++gi, ++count) {
                    tc_issue(s, P.g[gi], ps, tmem, count & 1, P.trace && blockIdx.x == 0 && tl < 4 ? P.trace + (tl * kTcMaxGemms + gi) * 10 : nullptr);
                }
        }
    } else {
        const RowId r = make_row_id(tmem);
        const int row = r.row;
        uint32_t count = 0;                                    // GEMMs consumed so far (bar_d parity)
        long long tl = -1;
        int gi = 0;
        auto wait_d = [&]() {
            tc::mbar_wait(s.bar_d, count & 1);
            ++count;
            tc::fence_after_sync();
            if (threadIdx.x == 64) tc_stamp(P, tl, gi, 0);
        };
        auto publish = [&](int g_next) {
            if (threadIdx.x == 64) tc_stamp(P, tl, g_next, 1);
            tc_publish_a(s);
            if (threadIdx.x == 64) tc_stamp(P, tl, g_next, 2);
        };
        float4 p_next = make_float4(0.f, 0.f, 0.f, 0.f);
        int ray_next = -1;
        auto fetch = [&](long long tile) {   // record of this thread's row in `tile` (issued early: ~1 us of latency)
            p_next = make_float4(0.f, 0.f, 0.f, 0.f);
            ray_next = -1;
            if (tile < n_tiles && tile * kTcRows + row < n_act) {
                p_next = P.rec_pos[tile * kTcRows + row];
                ray_next = P.rec_ray[tile * kTcRows + row];
            }
        };
        fetch(blockIdx.x);
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            ++tl;
            const long long base = tile * kTcRows;
            const int nv = (int)min((long long)kTcRows, n_act - base);
            const float4 p = p_next;
            const int ray = ray_next;
            fetch(tile + gridDim.x);
            if (r.half == 0) {
                s.ray[row] = ray;
                s.pos[row] = p;
            }
            tc::named_bar_sync(1, kTcRowThreads);
            if (warp == 2) {   // run starts, in record order
                int n = 0;
                for (int w4 = 0; w4 < kTcRows / 32; ++w4) {
                    const int m = w4 * 32 + lane;
                    const bool start = m < nv && (m == 0 || s.ray[m] != s.ray[m - 1]);
                    const unsigned bits = __ballot_sync(0xffffffffu, start);
                    if (start) s.runs[n + __popc(bits & ((1u << lane) - 1u))] = m;
                    n += __popc(bits);
                }
                if (lane == 0) {
                    s.runs[n] = nv;
                    *s.n_runs = n;
                }
            }
            gi = 0;
            if (P.n_sem > 0) {
                tc_build_xyz(s, r, p, P.pe_sem);
                publish(gi);
                for (int l = 0; l < P.n_sem; ++l, ++gi) {
                    wait_d();
                    if (l + 1 < P.n_sem) {
                        tc_epilogue_hidden(s, r, P.g[gi].n_pad);
                        publish(gi + 1);
                    } else if (P.n_cls <= 32) {
                        tc_epilogue_semantic32(s, r, P.n_cls, P.g[gi].n_pad, P.softmax, p.w);
                    } else {
                        tc_epilogue_final(s, r, P.n_cls, P.g[gi].n_pad);
                    }
                }
                if (r.half == 0 && P.n_cls > 32) {   // wide heads: softmax over the thread's own column of the scratch
                    if (P.softmax) {
                        float mx = -INFINITY;
                        for (int c = 0; c < P.n_cls; ++c) mx = fmaxf(mx, s.a_hi[(size_t)c * kTcRows + row]);
                        float tot = 0.0f;
                        for (int c = 0; c < P.n_cls; ++c) {
                            const float e = expf(s.a_hi[(size_t)c * kTcRows + row] - mx);
                            s.a_hi[(size_t)c * kTcRows + row] = e;
                            tot += e;
                        }
                        for (int c = 0; c < P.n_cls; ++c)
                            s.a_hi[(size_t)c * kTcRows + row] = (s.a_hi[(size_t)c * kTcRows + row] / tot) * p.w;
                    } else {
                        for (int c = 0; c < P.n_cls; ++c) s.a_hi[(size_t)c * kTcRows + row] *= p.w;
                    }
                }
                tc_reduce_runs(s, r.rt, P.n_cls, P.sem_raw, P.n_cls, 0);
            }
            if (P.n_ins > 0) {
                const int width = P.d_ins * (P.slow_fast ? 2 : 1);
                for (int net = 0; net < (P.slow_fast ? 2 : 1); ++net) {
                    tc_build_xyz(s, r, p, P.pe_ins);
                    publish(gi);
                    for (int l = 0; l < P.n_ins; ++l, ++gi) {
                        wait_d();
                        if (l + 1 < P.n_ins) {
                            tc_epilogue_hidden(s, r, P.g[gi].n_pad);
                            publish(gi + 1);
                        } else {
                            tc_epilogue_final(s, r, P.d_ins, P.g[gi].n_pad);
                        }
                    }
                    if (r.half == 0)
                        for (int c = 0; c < P.d_ins; ++c) s.a_hi[(size_t)c * kTcRows + row] *= p.w;
                    if (threadIdx.x == 64) tc_stamp(P, tl, gi, 6);
                    tc_reduce_runs(s, r.rt, P.d_ins, P.ins, width, net * P.d_ins);
                    if (threadIdx.x == 64) tc_stamp(P, tl, gi, 7);
                }
            }
            if (P.n_rgb > 0) {
                const FactorParams& f = P.app;
                // appearance gather in quad layout (4 lanes x float4 = one 64-byte texel segment per tap, fully
                // coalesced): plane*line products go RAW into the A operand rows, then the row-mapped threads split them
                {
                    const int q = r.rt & 3;
                    // one (record, mode) item per quad and step: 128 x 3 items over 96 quads, perfectly balanced
                    for (int item = r.rt >> 2; item < 3 * kTcRows; item += kTcRowThreads / 4) {
                        const int m = item / 3, mode = item - m * 3;
                        const float4 pm = s.pos[m];
                        const float ca = mode == 2 ? pm.y : pm.x, cb = mode == 0 ? pm.y : pm.z;
                        const float cv = mode == 0 ? pm.z : (mode == 1 ? pm.y : pm.x);
                        const int W = f.pw[mode];
                        const Tap2 t2 = make_tap2(ca, cb, W, f.ph[mode]);
                        const Tap1 t1 = make_tap1(cv, f.ll[mode]);
                        const float* plane = f.plane[mode];
                        const float* line = f.line[mode];
#pragma unroll
                        for (int v = 0; v < NV; ++v) {
                            const int ch = v * 16 + q * 4;
                            const float4 pv4 = plane_tap(plane, t2, W, f.comps, ch);
                            const float4 lv4 = line_tap(line, t1, f.comps, ch);
                            const int k = mode * f.comps + ch;
                            *reinterpret_cast<float4*>(s.a_hi + ((size_t)(k >> 2) * kTcRows + m) * 4) =
                                make_float4(pv4.x * lv4.x, pv4.y * lv4.y, pv4.z * lv4.z, pv4.w * lv4.w);
                        }
                    }
                }
                tc::named_bar_sync(1, kTcRowThreads);
                tc_split_rows(s, r.lane_base, row, r.half, 3 * f.comps);
                publish(gi);
                wait_d();   // basis GEMM: features in D columns [0, dim_app)
                const int A = P.dim_app, pf = P.pe_feat, pv = P.pe_view;
                const int n_base = A + 3;
                // staging behind the K rows of the first rgb GEMM: base values x_b (A features, 3 direction components)
                const int k_rows = P.g[gi + 1].k_steps * 8;
                float* xb = s.a_hi + (size_t)k_rows * kTcRows;
                if (r.half == 0) {
                    for (int c0 = 0; c0 < A; c0 += 16) {
                        float v[16];
                        tc_ld_acc16(r, c0, P.g[gi].n_pad, v);
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (c0 + i < A) xb[(size_t)(c0 + i) * kTcRows + row] = v[i];
                    }
                } else if (r.half == 1 && ray >= 0) {
                    for (int k = 0; k < 3; ++k) xb[(size_t)(A + k) * kTcRows + row] = __ldg(P.rays + (int64_t)ray * 8 + 3 + k);
                } else if (r.half == 1) {
                    for (int k = 0; k < 3; ++k) xb[(size_t)(A + k) * kTcRows + row] = k == 2 ? 1.0f : 0.0f;
                }
                ++gi;
                tc::named_bar_sync(1, kTcRowThreads);
                if (threadIdx.x == 64) tc_stamp(P, tl, gi, 6);
                // MLP input [feat, dir, sin(feat 2^j), cos(feat 2^j), sin(dir 2^j), cos(dir 2^j)] (tensoRF.py:400-418):
                // one (record, base value) item per thread and step - SFU sincos (|error| < 4e-7 on these O(1) arguments,
                // far inside the 1e-4 budget; the FP32-FMA kernel keeps libm), higher frequencies by angle doubling -
                // written RAW to their K rows, then split like the gather.
                {
                    const int o_sf = A + 3, o_cf = o_sf + A * pf, o_sd = o_cf + A * pf, o_cd = o_sd + 3 * pv, n_in = o_cd + 3 * pv;
                    auto put = [&](int k, int m, float x) { s.a_hi[((size_t)(k >> 2) * kTcRows + m) * 4 + (k & 3)] = x; };
                    for (int item = r.rt; item < n_base * kTcRows; item += kTcRowThreads) {
                        const int b = item / kTcRows, m = item - b * kTcRows;     // b is warp-uniform
                        const float x = xb[(size_t)b * kTcRows + m];
                        const bool is_feat = b < A;
                        const int nf = is_feat ? pf : pv;
                        const int ks = is_feat ? o_sf + b * pf : o_sd + (b - A) * pv;
                        const int kc = is_feat ? o_cf + b * pf : o_cd + (b - A) * pv;
                        put(b, m, x);                                               // feat rows [0,A), dir rows [A,A+3)
                        float sv, cv;
                        __sincosf(x, &sv, &cv);
                        for (int j = 0; j < nf; ++j) {
                            put(ks + j, m, sv);
                            put(kc + j, m, cv);
                            const float s2 = 2.0f * sv * cv, c2 = 1.0f - 2.0f * sv * sv;   // (sin, cos)(2t)
                            sv = s2;
                            cv = c2;
                        }
                    }
                    for (int item = r.rt; item < (k_rows - n_in) * kTcRows; item += kTcRowThreads)
                        put(n_in + item / kTcRows, item % kTcRows, 0.0f);           // zero pad rows
                }
                tc::named_bar_sync(1, kTcRowThreads);
                if (threadIdx.x == 64) tc_stamp(P, tl, gi, 7);
                tc_split_rows(s, r.lane_base, row, r.half, k_rows);
                publish(gi);
                for (int l = 0; l < P.n_rgb; ++l, ++gi) {
                    wait_d();
                    if (l + 1 < P.n_rgb) {
                        tc_epilogue_hidden(s, r, P.g[gi].n_pad);
                        publish(gi + 1);
                    } else {
                        tc_epilogue_final(s, r, 3, P.g[gi].n_pad);
                    }
                }
                if (r.half == 0)
                    for (int c = 0; c < 3; ++c) {
                        const float x = s.a_hi[(size_t)c * kTcRows + row];
                        s.a_hi[(size_t)c * kTcRows + row] = (1.0f / (1.0f + expf(-x))) * p.w;
                    }
                tc_reduce_runs(s, r.rt, 3, P.rgb_raw, 3, 0);
            }
        }
        tc::fence_before_sync();
    }
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem, kTmemCols);
}

// W [out][in] -> per k-step slab, [hi|lo][2 k-chunks][n_pad][4] for n_pad > 128 else [2 k-chunks][hi|lo][n_pad][4]
// (zero padded; hi/lo tf32-exact)
__global__ void pack_linear_tc_kernel(const float* __restrict__ w, const float* __restrict__ bias, int n_out, int n_in,
                                      float* __restrict__ dst, int n_pad, int slabs) {
    const int64_t total = (int64_t)slabs * kTcSlabK * n_pad;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int k = (int)(idx / n_pad), n = (int)(idx % n_pad);
    const int k_bias = (n_in + kTcSlabK - 1) / kTcSlabK * kTcSlabK;      // row 0 of the extra slab
    float x = (k < n_in && n < n_out) ? w[(size_t)n * n_in + k] : 0.0f;
    if (bias && k == k_bias && n < n_out) x = bias[n];
    float hi, lo;
    tc::split_tf32(x, hi, lo);
    const int slab = k / kTcSlabK, kc = (k % kTcSlabK) / 4, ki = k & 3;   // kc in {0,1}
    float* base = dst + (size_t)slab * (2 * kTcSlabK * n_pad);
    if (n_pad <= 128) {   // [k chunk][hi | lo][n_pad][4]
        const size_t off = ((size_t)(kc * 2) * n_pad + n) * 4 + ki;
        base[off] = hi;
        base[(size_t)n_pad * 4 + off] = lo;
    } else {              // [hi | lo][k chunk][n_pad][4]
        const size_t off = ((size_t)kc * n_pad + n) * 4 + ki;
        base[off] = hi;
        base[(size_t)kTcSlabK * n_pad + off] = lo;
    }
}

}  // namespace
}  // namespace clift


namespace clift {

static bool tc_add_stack(TcHeadsParams& P, const clift_mlp& m) {
    for (int l = 0; l < m.n_layers; ++l) {
        if (!m.w_tc[l] || P.n_gemms >= kTcMaxGemms) return false;
        TcGemm& g = P.g[P.n_gemms];
        g.w = m.w_tc[l];
        g.k_steps = (int)ceil_div(m.dims[l], 8);
        g.n_pad = (int)round_up(m.dims[l + 1], 32);
        g.has_bias = 1;        // clift_pack_linear_tc was given the layer's bias (all MLP layers have one)
        if (m.dims[l] > kTcMaxK || g.n_pad > 256) return false;
        ++P.n_gemms;
    }
    return true;
}

bool heads_tc_available(const clift_field* f, int heads) {
    auto ok = [](const clift_mlp& m) {
        for (int l = 0; l < m.n_layers; ++l)
            if (!m.w_tc[l] || m.dims[l] > kTcMaxK || m.dims[l + 1] > 256) return false;
        return m.n_layers >= 1;
    };
    if ((heads & CLIFT_HEAD_SEMANTIC) && !ok(f->semantic)) return false;
    if ((heads & CLIFT_HEAD_INSTANCE) && (!ok(f->instance_fast) || (f->slow_fast && !ok(f->instance_slow)))) return false;
    if (heads & CLIFT_HEAD_RGB) {
        if (!ok(f->rgb) || !f->basis_tc || f->dim_appearance > 64 || f->appearance_comps % 8) return false;
        // the base/sin/cos staging of the rgb input lives behind the K rows of the first rgb GEMM
        if ((int)round_up(f->rgb.dims[0], 8) + (f->dim_appearance + 3) > kTcMaxK) return false;
    }
    return true;
}

static long long* g_tc_trace = nullptr;
void set_tc_trace(long long* p) { g_tc_trace = p; }
long long* get_tc_trace() { return g_tc_trace; }

int launch_heads_forward_tc(const clift_render_cfg* cfg, const clift_field* field, const float* rays, const Workspace& ws,
                            int64_t cap, int64_t n_rays, float* rgb_raw, float* sem_raw, float* ins, cudaStream_t stream) {
    TcHeadsParams P;
    memset(&P, 0, sizeof(P));
    P.rec_pos = ws.rec_pos;
    P.rec_ray = ws.rec_ray;
    P.stats = reinterpret_cast<const unsigned long long*>(ws.stats);
    P.cap = cap;
    P.rays = rays;
    P.app = make_factors(field, true);
    P.dim_app = field->dim_appearance;
    P.pe_view = field->pe_view;
    P.pe_feat = field->pe_feat;
    P.pe_sem = field->pe_sem;
    P.pe_ins = field->pe_ins;
    P.n_cls = field->num_classes;
    P.d_ins = field->dim_instance;
    P.slow_fast = field->slow_fast;
    P.softmax = cfg->semantic_softmax;
    P.rgb_raw = rgb_raw;
    P.sem_raw = sem_raw;
    P.ins = ins;
    P.trace = g_tc_trace;
    bool ok = true;
    if (sem_raw) {
        P.n_sem = field->semantic.n_layers;
        ok = ok && tc_add_stack(P, field->semantic);
    }
    if (ins) {
        P.n_ins = field->instance_fast.n_layers;
        ok = ok && tc_add_stack(P, field->instance_fast);
        if (field->slow_fast) ok = ok && tc_add_stack(P, field->instance_slow);
    }
    if (rgb_raw) {
        P.n_rgb = field->rgb.n_layers;
        if (P.n_gemms < kTcMaxGemms && field->basis_tc) {
            TcGemm& g = P.g[P.n_gemms];
            g.w = field->basis_tc;
            g.k_steps = (int)ceil_div(3 * field->appearance_comps, 8);
            g.n_pad = (int)round_up(field->dim_appearance, 32);
            g.has_bias = 0;    // appearance_basis_mat has bias=False (tensoRF.py:65)
            ++P.n_gemms;
        } else {
            ok = false;
        }
        ok = ok && tc_add_stack(P, field->rgb);
    }
    if (!ok) {
        set_error("launch_heads_forward_tc: field lacks tensor-core operands (clift_pack_linear_tc) or exceeds the envelope");
        return CLIFT_ERR_UNSUPPORTED;
    }
    if (P.n_gemms == 0 || n_rays <= 0) return CLIFT_OK;
    const int grid = sm_count();
#define CLIFT_TC_CASE(NV)                                                                                              \
    case NV: {                                                                                                         \
        CLIFT_CUDA(cudaFuncSetAttribute(heads_tc_forward_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                        (int)kTcSmemBytes));                                                           \
        heads_tc_forward_kernel<NV><<<grid, kTcThreads, kTcSmemBytes, stream>>>(P);                                    \
        break;                                                                                                         \
    }
    switch (P.app.comps / 16) {
        CLIFT_TC_CASE(1)
        CLIFT_TC_CASE(2)
        CLIFT_TC_CASE(3)
        CLIFT_TC_CASE(4)
        default:
            set_error("launch_heads_forward_tc: appearance_comps %d not in {16,32,48,64}", P.app.comps);
            return CLIFT_ERR_UNSUPPORTED;
    }
#undef CLIFT_TC_CASE
    CLIFT_AFTER_LAUNCH("heads_tc_forward_kernel");
    return CLIFT_OK;
}

}  // namespace clift

using namespace clift;

extern "C" int32_t clift_debug_tc_trace(long long* device_buf) {
    clift::set_tc_trace(device_buf);
    return CLIFT_OK;
}

extern "C" int64_t clift_tc_weight_floats(int32_t n_out, int32_t n_in, int32_t has_bias) {
    if (n_out <= 0 || n_in <= 0 || n_out > 256 || n_in > kTcMaxK) return CLIFT_ERR_UNSUPPORTED;
    const int n_pad = (int)round_up(n_out, 32), slabs = (int)ceil_div(n_in, kTcSlabK) + (has_bias ? 1 : 0);
    return (int64_t)slabs * 2 * kTcSlabK * n_pad;
}

extern "C" int32_t clift_pack_linear_tc(const float* w, const float* bias, float* dst, int32_t n_out, int32_t n_in, void* stream) {
    CLIFT_CHECK_ARG(w && dst, "null pointer");
    CLIFT_CHECK_SUPPORTED(n_out > 0 && n_in > 0 && n_out <= 256 && n_in <= kTcMaxK, "layer wider than 256");
    const int n_pad = (int)round_up(n_out, 32), slabs = (int)ceil_div(n_in, kTcSlabK) + (bias ? 1 : 0);
    const int64_t total = (int64_t)slabs * kTcSlabK * n_pad;
    pack_linear_tc_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(w, bias, n_out, n_in, dst, n_pad, slabs);
    CLIFT_AFTER_LAUNCH("pack_linear_tc_kernel");
    return CLIFT_OK;
}

extern "C" int32_t clift_debug_tc_gemm(const float* a, const float* w_tc, float* out, int32_t k, int32_t n_out, int32_t has_bias,
                                       void* stream) {
    CLIFT_CHECK_ARG(a && w_tc && out, "null pointer");
    CLIFT_CHECK_SUPPORTED(n_out > 0 && k > 0 && n_out <= 256 && k <= kTcMaxK, "layer wider than 256");
    TcGemm g;
    g.w = w_tc;
    g.k_steps = (int)ceil_div(k, 8);
    g.n_pad = (int)round_up(n_out, 32);
    g.has_bias = has_bias ? 1 : 0;
    CLIFT_CUDA(cudaFuncSetAttribute(tc_gemm_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes));
    tc_gemm_test_kernel<<<1, kTcThreads, kTcSmemBytes, (cudaStream_t)stream>>>(a, k, g, out);
    CLIFT_AFTER_LAUNCH("tc_gemm_test_kernel");
    return CLIFT_OK;
}
