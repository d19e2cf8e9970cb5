// Tensor-core (tcgen05 / TMEM) MLP heads on the compacted active samples, fp16-split variant - the default inference path
// (tensoRF.py:127-137, 383-418, 462-511, 565-594; renderer:103-131, 137-156).
//
// fp32-faithful arithmetic at the kind::f16 rate (K = 16 per MMA, twice the tf32 rate): every operand x is scaled by a
// power of two s (exact) and split into two fp16 numbers, hi = fp16(s x), lo = fp16(s x - hi), and
//     (s_a A)(s_w W) ~= A_hi*W_hi + A_hi*W_lo + A_lo*W_hi        (22 significant bits per operand, fp32 accumulation in TMEM)
// The scales come from a pack-time bound chain (clift_pack_linear_tc16): with |input| <= B_in, every activation of a layer is
// bounded by B_out = B_in * max_n sum_k |W_nk| + max|b|; s_a = 2^floor(log2(2^14 / B_in)) keeps every scaled activation below
// 2^14 (fp16 overflows at 65504), s_w does the same for the weights and the bias row.  The bound is loose (it ignores
// cancellation) but fp16 keeps 22 bits for any value above 2^-3 of... the scaled range [2^-3, 2^14], i.e. the bound may be
// 2^17 times the actual activation before precision starts to degrade (gracefully: absolute error 2^-25 scaled).  The
// un-scaling (one FMUL per accumulator element) rides in the epilogues.  Measured against fp64: same 1e-6 class as 3xTF32
// (tests/test_gpu_tc.py).
//
// Per CTA (one per SM, persistent over 128-record tiles):
//   shared memory : A_hi, A_lo  [K/8][128][8] fp16 each (64 KB + 64 KB, canonical K-major / no-swizzle UMMA layout: 8-row x
//                   16-byte core matrices, SBO 128 B, LBO 2 KB), a 5-stage x 16 KB weight ring fed by 1-D bulk TMA + mbarriers,
//                   a constant "ones" operand chunk (bias k-step), record positions / ray ids / run table / scale table
//   tensor memory : D = 128 lanes x 256 columns (fp32 accumulator)
//   warp 0        : weight producer (one lane)       warp 1 : MMA issuer (one lane, running descriptors)
//   warps 2..13   : 384 "row" threads, three per record (thread <-> TMEM lane quarter of its warp): build layer inputs,
//                   run epilogues (TMEM -> ReLU -> rescale -> fp16 split -> smem), softmax / sigmoid, per-ray run sums
// Both operand halves live in shared memory, so any thread may write any record's operand row: the appearance gather
// (quad layout) and the rgb-input builder write the final fp16 pairs directly - no staging pass.
#include <cuda_fp16.h>

#include "launchers.h"
#include "tcgen05.cuh"

namespace clift {

long long* get_tc_trace();   // heads_tc.cu (clift_debug_tc_trace)

namespace {

constexpr int kRows = 128;              // records per tile = UMMA M
constexpr int kParts = 3;               // threads per record
constexpr int kRowThreads = kRows * kParts;
constexpr int kThreads = 64 + kRowThreads;
constexpr int kMaxK = 256;
constexpr int kStages = 5;
constexpr int kStepK = 16;              // K per tcgen05.mma kind::f16
constexpr int kStageBytes = 2 * kStepK * 256 * 2;    // hi + lo, N up to 256
constexpr int kABytes = kMaxK * kRows * 2;           // one operand half
constexpr int kOnesBytes = 2 * kRows * 16;
constexpr int kMaxGemms = 24;
constexpr int kHeaderFloats = 16;       // per packed layer: ca, cw, 1/(ca cw), out bound, in bound, L, max|W|, max|b|
constexpr uint32_t kTmemCols = 512;     // two 256-column accumulators: GEMM g accumulates in buffer g & 1
constexpr int kRoundCols = 16 * kParts; // accumulator columns (= next layer's K rows) finished per epilogue round
constexpr int kMaxRounds = 8;
constexpr int kXbStride = kRows + 8;    // row stride of the rgb base-value staging (bank-conflict-free for 4 bases x 8 rows)

struct Smem {
    unsigned char* a_hi;   // [kMaxK/8][128][8] fp16  (also the [c][128] float scratch of the final epilogues)
    unsigned char* a_lo;
    unsigned char* w;      // [kStages][kStageBytes]
    unsigned char* ones;   // A chunk pair [2][128][8] fp16 with element 0 of chunk 0 = 1
    float4* pos;           // [128]
    int* ray;              // [128]
    int* runs;             // [129]
    int* n_runs;
    float2* sc;            // [kMaxGemms] (ca, 1/(ca cw)) per GEMM
    uint64_t* full;
    uint64_t* empty;
    uint64_t* bar_a;       // [kMaxRounds] operand rounds ready (every row thread arrives), one phase per GEMM and round
    uint64_t* bar_d;       // accumulator ready (MMA -> row threads)
    uint64_t* full_peer;   // [kStages] CTA pairs: the peer's half of the weight stage has landed (leader's copy is used)
    uint32_t* tmem_base;
};

constexpr size_t kSmemBytes = 2 * (size_t)kABytes + (size_t)kStages * kStageBytes + kOnesBytes + (size_t)kRows * 16 +
                              (2 * kRows + 8) * 4 + kMaxGemms * 8 + 256;
static_assert(kSmemBytes <= 232448, "shared memory budget of one sm_100a CTA");

__device__ __forceinline__ Smem carve_smem(unsigned char* raw) {
    Smem s;
    s.a_hi = raw;
    s.a_lo = s.a_hi + kABytes;
    s.w = s.a_lo + kABytes;
    s.ones = s.w + (size_t)kStages * kStageBytes;
    s.pos = reinterpret_cast<float4*>(s.ones + kOnesBytes);
    s.ray = reinterpret_cast<int*>(s.pos + kRows);
    s.runs = s.ray + kRows;
    s.n_runs = s.runs + kRows + 1;
    s.sc = reinterpret_cast<float2*>(s.n_runs + 7);
    s.full = reinterpret_cast<uint64_t*>(s.sc + kMaxGemms);
    s.empty = s.full + kStages;
    s.bar_a = s.empty + kStages;
    s.bar_d = s.bar_a + kMaxRounds;
    s.full_peer = s.bar_d + 1;
    s.tmem_base = reinterpret_cast<uint32_t*>(s.full_peer + kStages);
    return s;
}

// One GEMM of the schedule: D[128 x n_pad] = A[128 x 16*k_steps] * W^T (+ bias row), weights packed by clift_pack_linear_tc16
struct Gemm {
    const __half* w;    // slabs (after the header)
    const __half* w_pair;   // the same weights in CTA-pair layout: [rank][k-step][per-CTA slab], see pack_linear_tc16_kernel
    const float* meta;  // header of the packed layer
    int k_steps;        // ceil(K / 16)
    int n_pad;          // multiple of 32, <= 256
    int has_bias;
    int n_sets;         // stacked small-N GEMMs (n_pad <= 64) spread their k-steps round-robin over n_sets accumulator column sets
                        //    (2 n_pad columns each; the epilogue adds them): the dependent chain of latency-bound MMAs shortens
    int a_rounds;       // 1: the A operand is published whole; else ceil(previous n_pad / 48): it streams out of the previous
                        //    layer's epilogue 48 K rows (3 k-steps) per round
};

struct PipeState {
    int stage = 0;
    uint32_t phase = 0;
    __device__ __forceinline__ void advance() {
        if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
        }
    }
};

__device__ __forceinline__ int ksteps_per_stage(int n_pad) { return n_pad >= 256 ? 1 : 256 / n_pad; }
// CTA pairs: bytes of one k-step's slab in ONE CTA, and k-steps per 16 KB ring stage
__device__ __host__ __forceinline__ uint32_t pair_slab_bytes(int n_pad) { return (n_pad > 128 ? 32u : 48u) * (uint32_t)n_pad; }
__device__ __forceinline__ int pair_ksteps_per_stage(int n_pad) { return min(8, (int)(kStageBytes / pair_slab_bytes(n_pad))); }

// warp 0, one lane: stream the GEMM's weight slabs (64 * n_pad bytes per k-step; CTA pairs: this CTA's half)
template <bool kPair>
__device__ __forceinline__ void produce(const Smem& s, const Gemm& g, PipeState& ps, uint32_t rank) {
    const int per = kPair ? pair_ksteps_per_stage(g.n_pad) : ksteps_per_stage(g.n_pad);
    const uint32_t kstep_bytes = kPair ? pair_slab_bytes(g.n_pad) : 64u * g.n_pad;
    const int steps = g.k_steps + g.has_bias;
    const unsigned char* src = kPair ? reinterpret_cast<const unsigned char*>(g.w_pair) + (size_t)rank * steps * kstep_bytes
                                     : reinterpret_cast<const unsigned char*>(g.w);
    for (int k0 = 0; k0 < steps; k0 += per, ps.advance()) {
        const uint32_t bytes = (uint32_t)min(per, steps - k0) * kstep_bytes;
        tc::mbar_wait(&s.empty[ps.stage], ps.phase ^ 1);
        tc::mbar_arrive_expect_tx(&s.full[ps.stage], bytes);
        tc::bulk_load(s.w + (size_t)ps.stage * kStageBytes, src + (size_t)k0 * kstep_bytes, bytes, &s.full[ps.stage]);
    }
}

// CTA pairs, peer CTA's warp 1, one lane: tell the leader's MMA thread when this CTA's half of a weight stage has landed
__device__ __forceinline__ void relay(const Smem& s, const Gemm& g, PipeState& ps, uint32_t leader_full_peer) {
    const int per = pair_ksteps_per_stage(g.n_pad);
    const int steps = g.k_steps + g.has_bias;
    for (int k0 = 0; k0 < steps; k0 += per, ps.advance()) {
        tc::mbar_wait(&s.full[ps.stage], ps.phase);
        tc::mbar_arrive_cluster_relaxed(leader_full_peer + 8u * (uint32_t)ps.stage);
    }
}

// warp 1, one lane: issue the three fp16 MMAs per k-step.
// n_pad > 128 : slab = [hi | lo][2 k-chunks][n_pad][8]: A_hi*W_hi, A_hi*W_lo, A_lo*W_hi
// n_pad <= 128: slab = [2 k-chunks][hi | lo][n_pad][8]: A_hi*[W_hi ; W_lo] is ONE MMA writing D[:, 0:n_pad) and
//               D[:, n_pad:2n_pad) (the epilogue adds the two column blocks), then A_lo*W_hi
// bias        : the slab after the last K slab carries the scaled bias in its k row 0; its A operand is the ones chunk.
// The issuing thread runs ~1 dependent instruction per 4-6 cycles, so the per-k-step body must stay a few dozen instructions
// (budget 384 cycles per N=256 k-step): descriptors are running values, GEMM fields are copied to registers, and the operand
// round logic exists only in the kStream instantiation.
// CTA pairs (kPair): M = 256 over both CTAs, each CTA's shared memory holds half of B's rows at the descriptor address:
//   n_pad > 128 : per-CTA slab = [hi | lo][2 k-chunks][n_pad/2][8]; the same three MMAs with N = n_pad split over the pair
//   n_pad <= 128: per-CTA slab = [Y: 2 k-chunks x n_pad rows][X: 2 k-chunks x n_pad/2 rows]; Y = W_hi in the leader and W_lo
//                 in the peer, so A_hi*[W_hi ; W_lo] (N = 2 n_pad) is one MMA; X = the CTA's half of W_hi for A_lo*W_hi
template <bool kStream, bool kPair>
__device__ __forceinline__ void issue(const Smem& s, const Gemm& gm, PipeState& ps, uint32_t d_tmem, uint32_t& a_parity,
                                      long long* trace = nullptr) {
    const int n_pad = gm.n_pad, k_steps = gm.k_steps, n_sets = gm.n_sets, a_rounds = gm.a_rounds;
    const int steps = k_steps + gm.has_bias;
    const int m_rows = kPair ? 2 * kRows : kRows;
    const uint32_t idesc = tc::make_idesc_f16(m_rows, n_pad);
    const bool stacked = n_pad <= 128;
    const uint32_t idesc_ss = stacked ? tc::make_idesc_f16(m_rows, 2 * n_pad) : idesc;
    const int per = kPair ? pair_ksteps_per_stage(n_pad) : ksteps_per_stage(n_pad);
    const uint32_t kstep16 = (kPair ? pair_slab_bytes(n_pad) : 64u * n_pad) >> 4;
    constexpr uint32_t kStage16 = kStageBytes >> 4;
    constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);     // SBO 128 B, descriptor version 1
    constexpr uint32_t kAStep = 2u * (kRows * 16u >> 4);       // two k-chunks per MMA
    const uint32_t a_lbo = (kRows * 16u >> 4) << 16;
    // first B operand (W_hi, or the stacked [W_hi ; W_lo]): rows per k-chunk in this CTA's shared memory
    const uint32_t rows1 = kPair ? (stacked ? (uint32_t)n_pad : (uint32_t)n_pad / 2) : (stacked ? 2u * n_pad : (uint32_t)n_pad);
    // second B operand: W_lo (not stacked) / the W_hi copy for A_lo*W_hi (stacked): offset from the first, rows per k-chunk
    const uint32_t off2 = 2u * rows1;
    const uint32_t rows2 = kPair ? (uint32_t)n_pad / 2 : rows1;
    uint32_t ah = (tc::smem_addr(s.a_hi) >> 4) | a_lbo;
    uint32_t al = (tc::smem_addr(s.a_lo) >> 4) | a_lbo;
    const uint32_t ones_lo = (tc::smem_addr(s.ones) >> 4) | a_lbo;
    const uint32_t w_base = tc::smem_addr(s.w) >> 4;
    const uint32_t set_stride = n_sets > 1 ? 2u * n_pad : 0u;
    auto desc = [](uint32_t lo) { return ((uint64_t)kDescHi << 32) | lo; };
    auto mma = [](uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
        if (kPair)
            tc::mma_ss_f16_pair(d, a, b, id, acc);
        else
            tc::mma_ss_f16(d, a, b, id, acc);
    };

    int rounds_seen = 0;
    auto need_round = [&](int r) {     // operand rounds 0..r written and published by every row thread
        while (rounds_seen <= r) {
            // (CTA pairs: CTA-scope wait + relaxed remote arrives, as in heads_x16.cu - each SM's tensor core reads its own
            // CTA's operand rows, published there with fence.proxy.async; the cluster-scope forms cost ~1 K cycles per hand-off)
            tc::mbar_wait(&s.bar_a[rounds_seen], (a_parity >> rounds_seen) & 1u);
            a_parity ^= 1u << rounds_seen;
            ++rounds_seen;
            tc::fence_after_sync();
        }
    };
    if (!kStream) need_round(a_rounds - 1);
    if (trace) trace[3] = clock64();
    int kk = 0, set = 0;
    uint32_t set_off = 0u;
    for (int k0 = 0; k0 < steps; k0 += per, ps.advance()) {
        tc::mbar_wait(&s.full[ps.stage], ps.phase);
        if (kPair) tc::mbar_wait(&s.full_peer[ps.stage], ps.phase);
        tc::fence_after_sync();
        if (trace && k0 == 0) trace[5] = clock64();
        uint32_t w_lo = w_base + (uint32_t)ps.stage * kStage16;
        const int n_here = min(per, steps - k0);
        for (int j = 0; j < n_here; ++j, ++kk, w_lo += kstep16, ah += kAStep, al += kAStep) {
            if (kStream) need_round(min(kk * kStepK / kRoundCols, a_rounds - 1));
            const bool bias_step = kk >= k_steps;
            const uint64_t a_desc = desc(bias_step ? ones_lo : ah);
            const uint64_t b1 = desc(w_lo | (rows1 << 16));
            const uint32_t dt = d_tmem + set_off;
            mma(dt, a_desc, b1, idesc_ss, kk >= n_sets ? 1u : 0u);   // the first MMA into a column set overwrites
            if (stacked) {
                if (!bias_step) mma(dt, desc(al), kPair ? desc((w_lo + off2) | (rows2 << 16)) : desc(w_lo | (rows1 << 16)), idesc, 1u);
            } else {
                mma(dt, a_desc, desc((w_lo + off2) | (rows2 << 16)), idesc, 1u);
                if (!bias_step) mma(dt, desc(al), b1, idesc, 1u);
            }
            set_off += set_stride;
            if (++set == n_sets) {
                set = 0;
                set_off = 0u;
            }
        }
        if (kPair)
            tc::mma_commit_pair(&s.empty[ps.stage], 3);
        else
            tc::mma_commit(&s.empty[ps.stage]);
    }
    if (kStream) need_round(a_rounds - 1);   // keep the round barriers' parities in step when K ends before the last round
    if (kPair)
        tc::mma_commit_pair(s.bar_d, 3);
    else
        tc::mma_commit(s.bar_d);
    if (trace) trace[4] = clock64();
}

__device__ __forceinline__ uint32_t h2_bits(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

// (x, y) -> fp16 pair of the values and of their remainders
__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(x, y);
    const float2 b = __half22float2(h);
    hi = h2_bits(h);
    lo = h2_bits(__floats2half2_rn(x - b.x, y - b.y));
}

// 16 consecutive K values (k0 multiple of 16, already scaled) of record `row` -> A_hi / A_lo
__device__ __forceinline__ void put16(const Smem& s, int row, int k0, const float* v) {
    uint32_t h[8], l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) split2(v[2 * i], v[2 * i + 1], h[i], l[i]);
    const size_t off = ((size_t)(k0 >> 3) * kRows + row) * 16;
    uint4* dh = reinterpret_cast<uint4*>(s.a_hi + off);
    uint4* dl = reinterpret_cast<uint4*>(s.a_lo + off);
    dh[0] = make_uint4(h[0], h[1], h[2], h[3]);
    dh[kRows] = make_uint4(h[4], h[5], h[6], h[7]);
    dl[0] = make_uint4(l[0], l[1], l[2], l[3]);
    dl[kRows] = make_uint4(l[4], l[5], l[6], l[7]);
}

// one K value of one record (any thread)
__device__ __forceinline__ void put1(const Smem& s, int row, int k, float x) {
    const __half h = __float2half_rn(x);
    const __half l = __float2half_rn(x - __half2float(h));
    const size_t off = ((size_t)(k >> 3) * kRows + row) * 16 + (size_t)(k & 7) * 2;
    *reinterpret_cast<__half*>(s.a_hi + off) = h;
    *reinterpret_cast<__half*>(s.a_lo + off) = l;
}

// two adjacent K values (k even) of one record
__device__ __forceinline__ void put2(const Smem& s, int row, int k, float x, float y) {
    uint32_t hi, lo;
    split2(x, y, hi, lo);
    const size_t off = ((size_t)(k >> 3) * kRows + row) * 16 + (size_t)(k & 7) * 2;
    *reinterpret_cast<uint32_t*>(s.a_hi + off) = hi;
    *reinterpret_cast<uint32_t*>(s.a_lo + off) = lo;
}

// every row thread: the operand rows it wrote for round `r` of the next GEMM are visible to the tensor core
// (CTA pairs: the barrier lives in the leader CTA; `leader_bar_a` is its shared::cluster address, 0 = single-CTA mode)
__device__ __forceinline__ void arrive_round(const Smem& s, int r, uint32_t leader_bar_a = 0u) {
    tc::fence_proxy_async_smem();
    tc::fence_before_sync();
    if (leader_bar_a)
        tc::mbar_arrive_cluster_relaxed(leader_bar_a + 8u * (uint32_t)r);
    else
        tc::mbar_arrive(&s.bar_a[r]);
}

struct RowId {
    int row, part, rt;
    uint32_t lane_base;
};

__device__ __forceinline__ RowId make_row_id(uint32_t tmem) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    RowId r;
    const int quarter = warp & 3;           // the TMEM lane quarter a warp may touch is warp_id % 4
    r.row = quarter * 32 + lane;
    r.part = (warp - 2) >> 2;
    r.rt = threadIdx.x - 64;
    r.lane_base = tmem + ((uint32_t)(quarter * 32) << 16);
    return r;
}

// 16 accumulator columns of this thread's record; stacked GEMMs keep A_hi*W_lo in columns [n_pad, 2 n_pad) and may spread
// their k-steps over n_sets such column sets
__device__ __forceinline__ void ld_acc16(uint32_t d, int c0, int n_pad, int n_sets, float* v) {
    tc::tmem_ld16(d + (uint32_t)c0, v);
    if (n_pad <= 128) {
        for (int b = 1; b < 2 * n_sets; ++b) {
            float u[16];
            tc::tmem_ld16(d + (uint32_t)(b * n_pad + c0), u);
            tc::tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += u[i];
        }
    } else {
        tc::tmem_wait_ld();
    }
}

template <bool kPair = false>
__device__ __forceinline__ void tc16_init(const Smem& s) {
    for (int i = threadIdx.x; i < kOnesBytes / 2; i += kThreads)
        reinterpret_cast<__half*>(s.ones)[i] = __float2half_rn((i < kRows * 8 && (i & 7) == 0) ? 1.0f : 0.0f);
    tc::fence_proxy_async_smem();
    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) {
            tc::mbar_init(&s.full[i], 1);
            tc::mbar_init(&s.empty[i], 1);
        }
        for (int i = 0; i < kMaxRounds; ++i) tc::mbar_init(&s.bar_a[i], kPair ? 2 * kRowThreads : kRowThreads);
        for (int i = 0; i < kStages; ++i) tc::mbar_init(&s.full_peer[i], 1);
        tc::mbar_init(s.bar_d, 1);
        tc::fence_barrier_init();
    }
    if ((threadIdx.x >> 5) == 1) {
        if (kPair)
            tc::tmem_alloc_pair(s.tmem_base, kTmemCols);
        else
            tc::tmem_alloc(s.tmem_base, kTmemCols);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (kPair) tc::cluster_sync();     // the peer's barriers exist before anyone arrives on them remotely
    tc::fence_after_sync();
}

// ---------------------------------------------------------------------------------------------------------
// parity / bring-up kernel: out[128][n_pad] = a[128][K] * W^T (+ bias), one CTA
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1) tc16_gemm_test_kernel(const float* __restrict__ a, int K, Gemm g, float* __restrict__ out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Smem s = carve_smem(smem_raw);
    const int warp = threadIdx.x >> 5;
    tc16_init(s);
    const uint32_t tmem = *s.tmem_base;
    if (warp == 0) {
        if (tc::elect_one()) {
            PipeState ps;
            produce<false>(s, g, ps, 0);
        }
    } else if (warp == 1) {
        if (tc::elect_one()) {
            PipeState ps;
            uint32_t a_parity = 0;
            issue<false, false>(s, g, ps, tmem, a_parity);
        }
    } else {
        const RowId r = make_row_id(tmem);
        const float ca = g.meta[0], inv = g.meta[2];
        for (int k0 = r.part * 16; k0 < g.k_steps * 16; k0 += 16 * kParts) {
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = (k0 + i < K) ? a[(size_t)r.row * K + k0 + i] * ca : 0.0f;
            put16(s, r.row, k0, v);
        }
        arrive_round(s, 0);
        tc::mbar_wait(s.bar_d, 0);
        tc::fence_after_sync();
        for (int c0 = r.part * 16; c0 < g.n_pad; c0 += 16 * kParts) {
            float v[16];
            ld_acc16(r.lane_base, c0, g.n_pad, g.n_sets, v);
#pragma unroll
            for (int i = 0; i < 16; ++i) out[(size_t)r.row * g.n_pad + c0 + i] = v[i] * inv;
        }
        tc::fence_before_sync();
    }
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem, kTmemCols);
}

// ---------------------------------------------------------------------------------------------------------
// production kernel
// ---------------------------------------------------------------------------------------------------------
struct HeadsParams {
    const float4* rec_pos;
    const int32_t* rec_ray;
    const unsigned long long* stats;
    long long cap;
    const float* rays;
    FactorParams app;
    FactorParams semg, insg;    // grid-mode semantic / instance heads (comps == 0: MLP mode, the stack reads xyz): the stack's
                                // first GEMM is then the bias-free basis Linear over the set's plane*line products
    int dim_app, pe_view, pe_feat, pe_sem, pe_ins;
    int n_cls, d_ins, slow_fast, softmax, use_sets;
    int n_gemms;
    Gemm g[kMaxGemms];          // semantic | instance fast | instance slow | basis | rgb
    int n_sem, n_ins, n_rgb;
    float* rgb_raw;
    float* sem_raw;
    float* ins;
    long long* trace;
    int stream;                 // 1: hidden-layer operands stream out of the epilogues in rounds, accumulators ping-pong
    int pair;                   // 1: CTA pairs (clusters of 2, cta_group::2 MMAs with M = 256): each SM stages half of every weight slab
    int park;                   // 1: the appearance gather runs row-mapped inside the MMA phases of the xyz stacks and parks
                                //    its fp16 pairs in spare tensor-memory columns (requires !stream: one accumulator buffer)
    // training forwards (save_for_backward): every layer input is also written, unscaled fp32, to the A-stash the backward
    // kernels read (heads.cu: StashLayout / stash_idx), the per-record rgb to rec_rgb; null = inference
    float* stash_a;
    float* rec_rgb;
    StashLayout lay;
};

constexpr uint32_t kParkCol = 256;      // first tensor-memory column of the parked appearance products

__device__ __forceinline__ void stamp(const HeadsParams& P, int tile_local, int gi, int slot) {
    if (P.trace && blockIdx.x == 0 && tile_local < 4) P.trace[(tile_local * kMaxGemms + gi) * 10 + slot] = clock64();
}

// hidden-layer epilogue: D (bias included) -> ReLU -> * e (rescale to the next layer's operand scale) -> fp16 pairs
// `st_blk` (training): the stash block of the NEXT layer's input (st_rows rows); it gets the unscaled activation D * inv
__device__ __forceinline__ void relu_put16(const Smem& s, const RowId& r, int c0, float* v, float e, float* st_blk, int st_rows,
                                           float inv, bool relu = true) {
    if (relu) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.0f);
    }
    if (st_blk) {
#pragma unroll
        for (int i = 0; i < 16; ++i) st_blk[stash_idx(st_rows, c0 + i, r.row)] = v[i] * inv;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] *= e;
    put16(s, r.row, c0, v);
}

// Streams the next layer's operand: round j = this thread's chunk of accumulator columns [48 j, 48 j + 48), i.e. K rows
// (3 k-steps) of the next GEMM, which the MMA thread issues as soon as all 384 row threads have arrived on bar_a[j] - the
// next layer's MMAs (into the other accumulator buffer) overlap the rest of this epilogue.
// `relu` false: a basis Linear (grid-mode heads) - its output is the next layer's input as it is.
__device__ __forceinline__ void epilogue_hidden(const Smem& s, const RowId& r, uint32_t d, const Gemm& g, float e, bool stream,
                                                uint32_t leader_bar_a, float* st_blk = nullptr, int st_rows = 0, float inv = 0.0f,
                                                bool relu = true) {
    const int n_pad = g.n_pad;
    const int rounds = (n_pad + kRoundCols - 1) / kRoundCols;
    const int c_begin = r.part * 16;
    if (n_pad <= 128) {
        for (int j = 0, c0 = c_begin; j < rounds; ++j, c0 += kRoundCols) {
            if (c0 < n_pad) {
                float v[16];
                ld_acc16(d, c0, n_pad, g.n_sets, v);
                relu_put16(s, r, c0, v, e, st_blk, st_rows, inv, relu);
            }
            if (stream) arrive_round(s, j, leader_bar_a);
        }
        if (!stream) arrive_round(s, 0, leader_bar_a);
        return;
    }
    for (int j = 0, c0 = c_begin; j < rounds; ++j, c0 += kRoundCols) {
        if (c0 < n_pad) {
            float v[16];
            tc::tmem_ld16(d + (uint32_t)c0, v);
            tc::tmem_wait_ld();
            relu_put16(s, r, c0, v, e, st_blk, st_rows, inv, relu);
        }
        if (stream) arrive_round(s, j, leader_bar_a);
    }
    if (!stream) arrive_round(s, 0, leader_bar_a);
}

// final-layer epilogue (part 0 threads): D * inv -> scratch[c][row] for c < n_out (scratch = A_hi region as floats)
__device__ __forceinline__ void epilogue_final(const Smem& s, const RowId& r, uint32_t d, int n_out, const Gemm& g, float inv) {
    if (r.part != 0) return;
    float* scratch = reinterpret_cast<float*>(s.a_hi);
    for (int c0 = 0; c0 < n_out; c0 += 16) {
        float v[16];
        ld_acc16(d, c0, g.n_pad, g.n_sets, v);
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (c0 + i < n_out) scratch[(size_t)(c0 + i) * kRows + r.row] = v[i] * inv;
    }
}

// semantic final layer for n_cls <= 32: logits in registers -> softmax -> * w -> scratch[c][row]
__device__ __forceinline__ void epilogue_semantic32(const Smem& s, const RowId& r, uint32_t d, int n_cls, const Gemm& g, int softmax,
                                                    float w, float inv, float* st_prob = nullptr) {
    if (r.part != 0) return;
    float* scratch = reinterpret_cast<float*>(s.a_hi);
    float v[32];
    ld_acc16(d, 0, g.n_pad, g.n_sets, v);
    if (n_cls > 16) {
        ld_acc16(d, 16, g.n_pad, g.n_sets, v + 16);
    } else {
#pragma unroll
        for (int i = 16; i < 32; ++i) v[i] = 0.0f;
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] *= inv;
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
        if (st_prob) {      // training: the softmax backward needs the probabilities
            const float it = 1.0f / tot;
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (i < n_cls) st_prob[stash_idx(n_cls, i, r.row)] = v[i] * it;
        }
        const float sc = w / tot;
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= sc;
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= w;
    }
#pragma unroll
    for (int i = 0; i < 32; ++i)
        if (i < n_cls) scratch[(size_t)i * kRows + r.row] = v[i];
}

// xyz (+ sin/cos PE, dimension-major frequency-minor) * ca -> A operand, zero padded to a multiple of 16
__device__ __forceinline__ void build_xyz(const Smem& s, const RowId& r, const float4& p, int pe, float ca) {
    if (pe == 0) {   // the shipped configuration (pe_sem = pe_ins = 0): one k-step, no decode
        if (r.part == 0) {
            const float v[16] = {p.x * ca, p.y * ca, p.z * ca, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            put16(s, r.row, 0, v);
        }
        return;
    }
    const int n_in = 3 + 6 * pe;
    const float xyz[3] = {p.x, p.y, p.z};
    for (int k0 = r.part * 16; k0 < n_in; k0 += 16 * kParts) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
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
            v[i] = x * ca;
        }
        put16(s, r.row, k0, v);
    }
}

// row threads: sum rows [0,nch) of the float scratch over each ray run and add into dst[ray*stride + col0 + c].  Four adjacent
// lanes share one (run, channel) item: each sums a contiguous quarter of the run, then a fixed-order shuffle tree combines
// them (bit-reproducible).
__device__ __forceinline__ void reduce_runs(const Smem& s, int rt, int nch, float* __restrict__ dst, int stride, int col0) {
    const float* scratch = reinterpret_cast<const float*>(s.a_hi);
    tc::named_bar_sync(1, kRowThreads);
    const int n_runs = *s.n_runs;
    const int items = n_runs * nch, sub = rt & 3;
    for (int base = 0; base < items; base += kRowThreads / 4) {
        const int idx = base + (rt >> 2);
        float acc = 0.0f;
        int ray = 0, c = 0;
        if (idx < items) {
            const int rr = idx / nch;
            c = idx - rr * nch;
            const int m0 = s.runs[rr], m1 = s.runs[rr + 1];
            const int len = m1 - m0, q = (len + 3) >> 2;
            const int a = m0 + min(sub * q, len), b = m0 + min((sub + 1) * q, len);
            for (int m = a; m < b; ++m) acc += scratch[(size_t)c * kRows + m];
            ray = s.ray[m0];
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        if (idx < items && sub == 0) atomicAdd(dst + (int64_t)ray * stride + col0 + c, acc);
    }
    tc::named_bar_sync(1, kRowThreads);
}

// all tap loads of one (record, mode) gather item in flight at once: NV x 4 plane taps (L2, the long latency) first, then
// the line taps (L1-resident); out-of-range taps carry weight 0 and read a clamped (valid) texel
// `c_off`: first channel of the NV x 16 handled by this call (sets wider than the register budget go in two calls)
// `st` (training): stash block of 3*C rows that receives the unscaled products (the basis weight gradient reads them)
template <int NV>
__device__ __forceinline__ void gather_item(const FactorParams& f, int mode, const float4& pm, int q, float ca, const Smem& s,
                                            int m, int c_off = 0, float* st = nullptr) {
    const float c_a = mode == 2 ? pm.y : pm.x, c_b = mode == 0 ? pm.y : pm.z;
    const float c_v = mode == 0 ? pm.z : (mode == 1 ? pm.y : pm.x);
    const int W = f.pw[mode], H = f.ph[mode], Ln = f.ll[mode], C = f.comps;
    const Tap2 t2 = make_tap2(c_a, c_b, W, H);
    const Tap1 t1 = make_tap1(c_v, Ln);
    const int x0 = min(max(t2.x0, 0), W - 1), x1 = min(max(t2.x0 + 1, 0), W - 1);
    const int y0 = min(max(t2.y0, 0), H - 1), y1 = min(max(t2.y0 + 1, 0), H - 1);
    const int z0 = min(max(t1.z0, 0), Ln - 1), z1 = min(max(t1.z0 + 1, 0), Ln - 1);
    const float* p00 = f.plane[mode] + ((int64_t)y0 * W + x0) * C + q * 4 + c_off;
    const float* p10 = f.plane[mode] + ((int64_t)y0 * W + x1) * C + q * 4 + c_off;
    const float* p01 = f.plane[mode] + ((int64_t)y1 * W + x0) * C + q * 4 + c_off;
    const float* p11 = f.plane[mode] + ((int64_t)y1 * W + x1) * C + q * 4 + c_off;
    const float* l0 = f.line[mode] + (int64_t)z0 * C + q * 4 + c_off;
    const float* l1 = f.line[mode] + (int64_t)z1 * C + q * 4 + c_off;
    float4 a[NV], b[NV], c[NV], d[NV], u[NV], w[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        a[v] = ldg4(p00 + v * 16);
        b[v] = ldg4(p10 + v * 16);
        c[v] = ldg4(p01 + v * 16);
        d[v] = ldg4(p11 + v * 16);
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        u[v] = ldg4(l0 + v * 16);
        w[v] = ldg4(l1 + v * 16);
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        float4 pv = make_float4(0.f, 0.f, 0.f, 0.f), lv = make_float4(0.f, 0.f, 0.f, 0.f);
        fma4(pv, a[v], t2.w00);
        fma4(pv, b[v], t2.w10);
        fma4(pv, c[v], t2.w01);
        fma4(pv, d[v], t2.w11);
        fma4(lv, u[v], t1.w0);
        fma4(lv, w[v], t1.w1);
        const int k = mode * C + c_off + v * 16 + q * 4;      // multiple of 4
        if (st) {
            st[stash_idx(3 * C, k, m)] = pv.x * lv.x;
            st[stash_idx(3 * C, k + 1, m)] = pv.y * lv.y;
            st[stash_idx(3 * C, k + 2, m)] = pv.z * lv.z;
            st[stash_idx(3 * C, k + 3, m)] = pv.w * lv.w;
        }
        uint32_t h0, lo0, h1, lo1;
        split2(pv.x * lv.x * ca, pv.y * lv.y * ca, h0, lo0);
        split2(pv.z * lv.z * ca, pv.w * lv.w * ca, h1, lo1);
        const size_t off = ((size_t)(k >> 3) * kRows + m) * 16 + (size_t)(k & 4) * 2;
        *reinterpret_cast<uint2*>(s.a_hi + off) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(s.a_lo + off) = make_uint2(lo0, lo1);
    }
}

// one 32-byte sector, not allocated in L1 (the L1 array is the shared-memory array the MMAs are reading their operands from)
__device__ __forceinline__ void ldg8_na(const float* p, float4& lo, float4& hi) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w), "=f"(hi.x), "=f"(hi.y), "=f"(hi.z), "=f"(hi.w)
                 : "l"(p));
}

// Row-mapped gather slice: channels [8 slice, 8 slice + 8) of mode `mode` at this thread's own record, 12 16-byte loads in
// flight (one 32-byte sector per tap), products scaled and split to fp16 pairs, parked in 8 tensor-memory columns of the
// record's lane as [4 words hi | 4 words lo] - the image of one 16-byte A_hi chunk and one A_lo chunk.
__device__ __forceinline__ void gather_slice(const FactorParams& f, int mode, const float4& pm, int slice, float ca, uint32_t park,
                                             float* st_app = nullptr, int row = 0) {
    const float c_a = mode == 2 ? pm.y : pm.x, c_b = mode == 0 ? pm.y : pm.z;
    const float c_v = mode == 0 ? pm.z : (mode == 1 ? pm.y : pm.x);
    const int W = f.pw[mode], H = f.ph[mode], Ln = f.ll[mode], C = f.comps;
    const Tap2 t2 = make_tap2(c_a, c_b, W, H);
    const Tap1 t1 = make_tap1(c_v, Ln);
    const int x0 = min(max(t2.x0, 0), W - 1), x1 = min(max(t2.x0 + 1, 0), W - 1);
    const int y0 = min(max(t2.y0, 0), H - 1), y1 = min(max(t2.y0 + 1, 0), H - 1);
    const int z0 = min(max(t1.z0, 0), Ln - 1), z1 = min(max(t1.z0 + 1, 0), Ln - 1);
    const int ch = slice * 8;
    const float* p00 = f.plane[mode] + ((int64_t)y0 * W + x0) * C + ch;
    const float* p10 = f.plane[mode] + ((int64_t)y0 * W + x1) * C + ch;
    const float* p01 = f.plane[mode] + ((int64_t)y1 * W + x0) * C + ch;
    const float* p11 = f.plane[mode] + ((int64_t)y1 * W + x1) * C + ch;
    const float* l0 = f.line[mode] + (int64_t)z0 * C + ch;
    const float* l1 = f.line[mode] + (int64_t)z1 * C + ch;
    float4 a[2], b[2], c[2], d[2], u[2], w[2];
    ldg8_na(p00, a[0], a[1]);
    ldg8_na(p10, b[0], b[1]);
    ldg8_na(p01, c[0], c[1]);
    ldg8_na(p11, d[0], d[1]);
    ldg8_na(l0, u[0], u[1]);
    ldg8_na(l1, w[0], w[1]);
    uint32_t words[8];
#pragma unroll
    for (int v = 0; v < 2; ++v) {
        float4 pv = make_float4(0.f, 0.f, 0.f, 0.f), lv = make_float4(0.f, 0.f, 0.f, 0.f);
        fma4(pv, a[v], t2.w00);
        fma4(pv, b[v], t2.w10);
        fma4(pv, c[v], t2.w01);
        fma4(pv, d[v], t2.w11);
        fma4(lv, u[v], t1.w0);
        fma4(lv, w[v], t1.w1);
        split2(pv.x * lv.x * ca, pv.y * lv.y * ca, words[2 * v], words[4 + 2 * v]);
        split2(pv.z * lv.z * ca, pv.w * lv.w * ca, words[2 * v + 1], words[4 + 2 * v + 1]);
        if (st_app) {       // training: the basis weight gradient needs the plane*line products (block of 3*C rows)
            const int k = mode * C + ch + 4 * v, R = 3 * C;
            st_app[stash_idx(R, k, row)] = pv.x * lv.x;
            st_app[stash_idx(R, k + 1, row)] = pv.y * lv.y;
            st_app[stash_idx(R, k + 2, row)] = pv.z * lv.z;
            st_app[stash_idx(R, k + 3, row)] = pv.w * lv.w;
        }
    }
    tc::tmem_st8u(park + (uint32_t)(slice * 8), words);
}

// kPair: launched as clusters of two CTAs; the pair works on two tiles at once with M = 256 MMAs issued by the leader
// (cluster rank 0) - see issue().  Every CTA of a pair runs the same number of iterations (a CTA whose tile index is past the
// end carries an empty tile through the same schedule).
// kStash: training forward - every layer input is also written to the A-stash (compiled out of the inference instantiation:
// the extra predicates and address math in the epilogues cost it 6 %).
template <int NV, bool kPair, bool kStash>
__global__ void __launch_bounds__(kThreads, 1) heads_tc16_forward_kernel(const __grid_constant__ HeadsParams P) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Smem s = carve_smem(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < P.n_gemms) s.sc[threadIdx.x] = make_float2(P.g[threadIdx.x].meta[0], P.g[threadIdx.x].meta[2]);
    tc16_init<kPair>(s);
    const uint32_t tmem = *s.tmem_base;
    const long long n_act = min((long long)P.stats[0], P.cap);
    const int n_tiles = (int)((n_act + kRows - 1) / kRows);    // < 2^24: one call handles n_rays * n_samples < 2^31
    const uint32_t rank = kPair ? tc::cluster_ctarank() : 0u;
    // tiles of this CTA: tile_first, tile_first + tile_step, ... while the PAIR's (even) tile index is in range
    const int tile_step = kPair ? (int)gridDim.x : (int)gridDim.x;
    const int tile_first = kPair ? (int)(blockIdx.x & ~1u) : (int)blockIdx.x;     // iteration key (pairs: the leader's tile)
    const int tile_mine = kPair ? (int)rank : 0;                                   // + this CTA's offset inside the pair
    const uint32_t leader_bar_a = kPair ? tc::map_to_rank(&s.bar_a[0], 0) : 0u;

    if (warp == 0) {
        if (tc::elect_one()) {
            PipeState ps;
            for (int tile = tile_first; tile < n_tiles; tile += tile_step)
                for (int gi = 0; gi < P.n_gemms; ++gi) produce<kPair>(s, P.g[gi], ps, rank);
        }
    } else if (warp == 1) {
        if (tc::elect_one()) {
            PipeState ps;
            if (kPair && rank != 0) {
                const uint32_t leader_full_peer = tc::map_to_rank(&s.full_peer[0], 0);
                for (int tile = tile_first; tile < n_tiles; tile += tile_step)
                    for (int gi = 0; gi < P.n_gemms; ++gi) relay(s, P.g[gi], ps, leader_full_peer);
            } else {
                uint32_t count = 0, a_parity = 0;
                int tl = 0;
                for (int tile = tile_first; tile < n_tiles; tile += tile_step, ++tl)
                    for (int gi = 0; gi < P.n_gemms; ++gi, ++count) {
                        long long* tr = P.trace && blockIdx.x == 0 && tl < 4 ? P.trace + (tl * kMaxGemms + gi) * 10 : nullptr;
                        if (P.stream)
                            issue<true, kPair>(s, P.g[gi], ps, tmem + ((count & 1u) ? 256u : 0u), a_parity, tr);
                        else
                            issue<false, kPair>(s, P.g[gi], ps, tmem, a_parity, tr);
                    }
            }
        }
    } else {
        const RowId r = make_row_id(tmem);
        const int row = r.row;
        float* scratch = reinterpret_cast<float*>(s.a_hi);
        uint32_t count = 0;                                    // GEMMs consumed so far (bar_d parity, accumulator buffer)
        uint32_t d = r.lane_base;                              // accumulator of the GEMM waited for last
        int tl = -1;
        int gi = 0;
        auto wait_d = [&]() {
            tc::mbar_wait(s.bar_d, count & 1);
            d = r.lane_base + ((P.stream && (count & 1u)) ? 256u : 0u);
            ++count;
            tc::fence_after_sync();
            if (threadIdx.x == 64) stamp(P, tl, gi, 0);
        };
        auto publish = [&](int g_next) {                       // whole operand written by this thread: round 0
            if (threadIdx.x == 64) stamp(P, tl, g_next, 1);
            arrive_round(s, 0, leader_bar_a);
        };
        // one MLP stack whose first operand is already published: hidden epilogues stream the next layer's operand in place
        // parked appearance gather: this thread's record, mode = its part; one 8-channel slice per long MMA phase
        const int g_basis = P.n_gemms - P.n_rgb - 1;
        const uint32_t park = r.lane_base + kParkCol + (uint32_t)(r.part * 16 * NV);
        int park_next = 2 * NV;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 p_next = make_float4(0.f, 0.f, 0.f, 0.f);
        int ray_next = -1;
        float* st = nullptr;                                   // training: this tile's A-stash
        auto st_block = [&](int id, int l) -> float* {
            if (!kStash) return nullptr;
            return st + (size_t)P.lay.a_off[id][l] * kRows;
        };
        auto park_slice = [&](const float4& pp) {
            gather_slice(P.app, r.part, pp, park_next, s.sc[g_basis].x, park, st_block(4, 0), row);
            ++park_next;
        };
        // rgb stack alone (the xyz stacks run on heads_x16.cu): there is no long MMA phase of this tile to hide the gather
        // under, so the slices of the NEXT tile's records are gathered under this tile's four GEMMs and wait in the parked
        // columns across the tile boundary (the first tile of a CTA gathers in one burst)
        const bool prefetch = !kStash && P.park && P.n_sem == 0 && P.n_ins == 0 && P.n_rgb > 0;
        auto stash_xyz = [&](int id) {                         // layer-0 input of an xyz stack (pe = 0): 16 rows, 3 valid
            if (kStash && r.part == 0) {
                float* b = st_block(id, 0);
                b[stash_idx(16, 0, row)] = p.x;
                b[stash_idx(16, 1, row)] = p.y;
                b[stash_idx(16, 2, row)] = p.z;
                for (int k = 3; k < 16; ++k) b[stash_idx(16, k, row)] = 0.0f;
            }
        };
        // grid-mode head (tensoRF.py:72-85, 142-156): the stack's first operand = plane*line products of the head's own factor
        // set, in quad layout like the burst form of the appearance gather (comps 64: two passes of 32 channels)
        auto gather_set = [&](const FactorParams& f, float ca, float* st_prod) {
            const int q = r.rt & 3;
            for (int item = r.rt >> 2; item < 3 * kRows; item += kRowThreads / 4) {
                const int m = item / 3, mode = item - m * 3;
                switch (f.comps) {
                    case 16: gather_item<1>(f, mode, s.pos[m], q, ca, s, m, 0, st_prod); break;
                    case 32: gather_item<2>(f, mode, s.pos[m], q, ca, s, m, 0, st_prod); break;
                    case 48: gather_item<3>(f, mode, s.pos[m], q, ca, s, m, 0, st_prod); break;
                    default:
                        gather_item<2>(f, mode, s.pos[m], q, ca, s, m, 0, st_prod);
                        gather_item<2>(f, mode, s.pos[m], q, ca, s, m, 32, st_prod);
                        break;
                }
            }
        };
        // `basis_first`: GEMM 0 of the stack is a grid-mode head's basis Linear (no bias, no ReLU)
        auto run_hidden = [&](int n_layers, int id, bool basis_first = false) {
            for (int l = 0; l + 1 < n_layers; ++l, ++gi) {
                wait_d();
                // training: the epilogue of GEMM l writes the input of the NEXT Linear of stack `id` (the basis is not one)
                epilogue_hidden(s, r, d, P.g[gi], s.sc[gi + 1].x * s.sc[gi].y, P.stream != 0, leader_bar_a,
                                st_block(id, l + (basis_first ? 0 : 1)), P.g[gi].n_pad, s.sc[gi].y, !(basis_first && l == 0));
                if (threadIdx.x == 64) stamp(P, tl, gi + 1, 1);
                if (!prefetch && park_next < 2 * NV && P.g[gi + 1].k_steps >= 8 && P.g[gi + 1].n_pad > 128) park_slice(p);   // hides under that GEMM
                if (prefetch && park_next < 2 * NV) park_slice(p_next);
            }
            wait_d();   // final layer: the caller reads D, then advances gi
        };
        auto fetch = [&](int tile) {
            p_next = make_float4(0.f, 0.f, 0.f, 0.f);
            ray_next = -1;
            const long long rec = (long long)tile * kRows + row;
            if (tile < n_tiles && rec < n_act) {
                p_next = P.rec_pos[rec];
                ray_next = P.rec_ray[rec];
            }
        };
        fetch(tile_first + tile_mine);
        for (int tile_key = tile_first; tile_key < n_tiles; tile_key += tile_step) {
            ++tl;
            const int tile = tile_key + tile_mine;
            const int nv = (int)max(0ll, min((long long)kRows, n_act - (long long)tile * kRows));
            p = p_next;
            const int ray = ray_next;
            fetch(tile + tile_step);
            if (kStash) st = P.stash_a + (size_t)tile * P.lay.a_rows * kRows;
            if (!(prefetch && tl > 0)) park_next = (P.park && P.n_rgb > 0) ? 0 : 2 * NV;   // prefetch: slices parked so far stay
            if (r.part == 0) {
                s.ray[row] = ray;
                s.pos[row] = p;
            }
            tc::named_bar_sync(1, kRowThreads);
            if (warp == 2) {   // run starts, in record order
                int n = 0;
                for (int w4 = 0; w4 < kRows / 32; ++w4) {
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
                if (P.semg.comps) {
                    gather_set(P.semg, s.sc[gi].x, st_block(5, 0));
                } else {
                    build_xyz(s, r, p, P.pe_sem, s.sc[gi].x);
                    stash_xyz(0);
                }
                publish(gi);
                run_hidden(P.n_sem, 0, P.semg.comps != 0);
                if (P.n_cls <= 32) {
                    epilogue_semantic32(s, r, d, P.n_cls, P.g[gi], P.softmax, p.w, s.sc[gi].y,
                                        kStash ? st + (size_t)P.lay.prob_off * kRows : nullptr);
                } else {
                    epilogue_final(s, r, d, P.n_cls, P.g[gi], s.sc[gi].y);
                    if (r.part == 0) {   // wide heads: softmax over the thread's own column of the scratch
                        if (P.softmax) {
                            float mx = -INFINITY;
                            for (int c = 0; c < P.n_cls; ++c) mx = fmaxf(mx, scratch[(size_t)c * kRows + row]);
                            float tot = 0.0f;
                            for (int c = 0; c < P.n_cls; ++c) {
                                const float e = expf(scratch[(size_t)c * kRows + row] - mx);
                                scratch[(size_t)c * kRows + row] = e;
                                tot += e;
                            }
                            for (int c = 0; c < P.n_cls; ++c)
                                scratch[(size_t)c * kRows + row] = (scratch[(size_t)c * kRows + row] / tot) * p.w;
                        } else {
                            for (int c = 0; c < P.n_cls; ++c) scratch[(size_t)c * kRows + row] *= p.w;
                        }
                    }
                }
                ++gi;
                reduce_runs(s, r.rt, P.n_cls, P.sem_raw, P.n_cls, 0);
            }
            if (P.n_ins > 0) {
                const int width = P.d_ins * (P.slow_fast ? 2 : 1);
                for (int net = 0; net < (P.slow_fast ? 2 : 1); ++net) {
                    if (P.insg.comps) {        // fast and slow nets read the same basis feature (tensoRF.py:497-511)
                        gather_set(P.insg, s.sc[gi].x, net == 0 ? st_block(6, 0) : nullptr);
                    } else {
                        build_xyz(s, r, p, P.pe_ins, s.sc[gi].x);
                        stash_xyz(1 + net);
                    }
                    publish(gi);
                    run_hidden(P.n_ins, 1 + net, P.insg.comps != 0);
                    epilogue_final(s, r, d, P.d_ins, P.g[gi], s.sc[gi].y * p.w);
                    ++gi;
                    if (threadIdx.x == 64) stamp(P, tl, gi, 6);
                    reduce_runs(s, r.rt, P.d_ins, P.ins, width, net * P.d_ins);
                    if (threadIdx.x == 64) stamp(P, tl, gi, 7);
                }
            }
            if (P.n_rgb > 0) {
                const FactorParams& f = P.app;
                if (P.park) {
                    // the products were gathered slice by slice under the xyz stacks' MMAs (leftover slices now) and wait in
                    // tensor memory as chunk images: copy them into the operand rows
                    while (park_next < 2 * NV) park_slice(p);
                    tc::tmem_wait_st();
#pragma unroll
                    for (int j = 0; j < NV; ++j) {
                        uint32_t v[16];
                        tc::tmem_ld16u(park + (uint32_t)(16 * j), v);
                        tc::tmem_wait_ld();
                        const size_t off = ((size_t)(r.part * 2 * NV + 2 * j) * kRows + row) * 16;
                        *reinterpret_cast<uint4*>(s.a_hi + off) = make_uint4(v[0], v[1], v[2], v[3]);
                        *reinterpret_cast<uint4*>(s.a_lo + off) = make_uint4(v[4], v[5], v[6], v[7]);
                        *reinterpret_cast<uint4*>(s.a_hi + off + kRows * 16) = make_uint4(v[8], v[9], v[10], v[11]);
                        *reinterpret_cast<uint4*>(s.a_lo + off + kRows * 16) = make_uint4(v[12], v[13], v[14], v[15]);
                    }
                    if (prefetch) park_next = tile_key + tile_step < n_tiles ? 0 : 2 * NV;   // the parked columns are free again
                } else {
                    // appearance gather in quad layout (4 lanes x float4 = one 64-byte texel segment per tap, fully
                    // coalesced): scaled plane*line products go straight into the operand rows as fp16 pairs.  One
                    // (record, mode) item per quad and step: 128 x 3 items over 96 quads.
                    const int q = r.rt & 3;
                    const float ca = s.sc[gi].x;
                    for (int item = r.rt >> 2; item < 3 * kRows; item += kRowThreads / 4) {
                        const int m = item / 3, mode = item - m * 3;
                        gather_item<NV>(f, mode, s.pos[m], q, ca, s, m);
                    }
                }
                publish(gi);
                if (prefetch) {
                    if (park_next < 2 * NV) park_slice(p_next);
                    if (park_next < 2 * NV) park_slice(p_next);
                }
                wait_d();   // basis GEMM: features in D columns [0, dim_app)
                const int A = P.dim_app, pf = P.pe_feat, pv = P.pe_view;
                const int n_base = A + 3;
                // staging behind the K rows of the first rgb GEMM (inside the A_hi region): base values x_b [b][kXbStride]
                const int k_rows = P.g[gi + 1].k_steps * kStepK;
                float* xb = reinterpret_cast<float*>(s.a_hi + (size_t)(k_rows >> 3) * kRows * 16);
                if (r.part == 0) {
                    const float inv = s.sc[gi].y;
                    for (int c0 = 0; c0 < A; c0 += 16) {
                        float v[16];
                        ld_acc16(d, c0, P.g[gi].n_pad, P.g[gi].n_sets, v);
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (c0 + i < A) xb[(size_t)(c0 + i) * kXbStride + row] = v[i] * inv;
                    }
                } else if (r.part == 1 && ray >= 0) {
                    for (int k = 0; k < 3; ++k) xb[(size_t)(A + k) * kXbStride + row] = __ldg(P.rays + (int64_t)ray * 8 + 3 + k);
                } else if (r.part == 1) {
                    for (int k = 0; k < 3; ++k) xb[(size_t)(A + k) * kXbStride + row] = k == 2 ? 1.0f : 0.0f;
                }
                ++gi;
                tc::fence_before_sync();
                tc::named_bar_sync(1, kRowThreads);
                if (threadIdx.x == 64) stamp(P, tl, gi, 6);
                // MLP input [feat, dir, sin(feat 2^j), cos(feat 2^j), sin(dir 2^j), cos(dir 2^j)] (tensoRF.py:400-418): one
                // (record, base value) item per thread and step - SFU sincos (|error| < 4e-7 on these O(1) arguments),
                // higher frequencies by angle doubling - written as fp16 pairs at the first rgb layer's operand scale.
                // A warp covers 8 records x 4 consecutive base values, which spreads its 2-byte operand stores over all banks.
                {
                    const float ca = s.sc[gi].x;
                    float* st_in = st_block(3, 0);             // training: the rgb MLP input (k_rows rows)
                    const int o_sf = A + 3, o_cf = o_sf + A * pf, o_sd = o_cf + A * pf, o_cd = o_sd + 3 * pv, n_in = o_cd + 3 * pv;
                    const int n_items = ((n_base + 3) >> 2) * 4 * kRows;
                    for (int item = r.rt; item < n_items; item += kRowThreads) {
                        const int b = ((item >> 9) << 2) | (item & 3), m = (item >> 2) & (kRows - 1);
                        if (b >= n_base) continue;
                        const float x = xb[(size_t)b * kXbStride + m];
                        const bool is_feat = b < A;
                        const int nf = is_feat ? pf : pv;
                        const int ks = is_feat ? o_sf + b * pf : o_sd + (b - A) * pv;
                        const int kc = is_feat ? o_cf + b * pf : o_cd + (b - A) * pv;
                        put1(s, m, b, x * ca);
                        float sv, cv;
                        __sincosf(x, &sv, &cv);
                        if (st_in) st_in[stash_idx(k_rows, b, m)] = x;
                        if (nf == 2 && !((ks | kc) & 1)) {      // the shipped pe_feat = pe_view = 2: (f, 2f) pairs are aligned
                            put2(s, m, ks, sv * ca, 2.0f * sv * cv * ca);
                            put2(s, m, kc, cv * ca, (1.0f - 2.0f * sv * sv) * ca);
                            if (st_in) {
                                st_in[stash_idx(k_rows, ks, m)] = sv;
                                st_in[stash_idx(k_rows, ks + 1, m)] = 2.0f * sv * cv;
                                st_in[stash_idx(k_rows, kc, m)] = cv;
                                st_in[stash_idx(k_rows, kc + 1, m)] = 1.0f - 2.0f * sv * sv;
                            }
                            continue;
                        }
                        for (int j = 0; j < nf; ++j) {
                            put1(s, m, ks + j, sv * ca);
                            put1(s, m, kc + j, cv * ca);
                            if (st_in) {
                                st_in[stash_idx(k_rows, ks + j, m)] = sv;
                                st_in[stash_idx(k_rows, kc + j, m)] = cv;
                            }
                            const float s2 = 2.0f * sv * cv, c2 = 1.0f - 2.0f * sv * sv;
                            sv = s2;
                            cv = c2;
                        }
                    }
                    for (int item = r.rt; item < (k_rows - n_in) * kRows; item += kRowThreads) {
                        put1(s, item % kRows, n_in + item / kRows, 0.0f);
                        if (st_in) st_in[stash_idx(k_rows, n_in + item / kRows, item % kRows)] = 0.0f;
                    }
                }
                if (threadIdx.x == 64) stamp(P, tl, gi, 7);
                publish(gi);
                run_hidden(P.n_rgb, 3);
                epilogue_final(s, r, d, 3, P.g[gi], s.sc[gi].y);
                ++gi;
                if (r.part == 0)
                    for (int c = 0; c < 3; ++c) {
                        const float x = scratch[(size_t)c * kRows + row];
                        const float sig = 1.0f / (1.0f + expf(-x));
                        if (kStash && row < nv) P.rec_rgb[((long long)tile * kRows + row) * 4 + c] = sig;
                        scratch[(size_t)c * kRows + row] = sig * p.w;
                    }
                reduce_runs(s, r.rt, 3, P.rgb_raw, 3, 0);
            }
        }
        tc::fence_before_sync();
    }
    __syncthreads();
    if (kPair) tc::cluster_sync();     // nobody leaves while the peer may still signal its barriers / read its operands
    if (warp == 1) {
        if (kPair)
            tc::tmem_dealloc_pair(tmem, kTmemCols);
        else
            tc::tmem_dealloc(tmem, kTmemCols);
    }
}

// ---------------------------------------------------------------------------------------------------------
// pack-time kernels
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float pow2_floor_ratio(float num_log2, float x) {   // 2^floor(num_log2 - log2(x)), exponent clamped
    int e = 0;
    frexpf(x, &e);                        // x = m * 2^e, m in [0.5, 1)  =>  2^(e-1) <= x < 2^e
    int k = (int)num_log2 - e;            // 2^num / x > 2^(num - e)
    k = max(-60, min(60, k));
    return ldexpf(1.0f, k);
}

// header[0..7] = ca, cw, 1/(ca cw), out bound, in bound, L = max_n sum_k |W_nk|, max|W|, max|b|
// body of the plan kernels (whole CTA of 1024 threads; `red` = 96 floats of shared memory)
__device__ __forceinline__ void tc16_plan_body(const float* __restrict__ w, const float* __restrict__ bias, int n_out, int n_in,
                                               const float* in_bound_dev, float in_floor, float* __restrict__ header,
                                               float (*red)[32]) {
    // one warp per weight row (coalesced reads, fixed-order lane tree): it runs before every training forward
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float l1 = 0.0f, wm = 0.0f, bm = 0.0f;
    for (int n = warp; n < n_out; n += 32) {
        float sum = 0.0f;
        for (int k = lane; k < n_in; k += 32) {
            const float a = fabsf(w[(size_t)n * n_in + k]);
            sum += a;
            wm = fmaxf(wm, a);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        l1 = fmaxf(l1, sum);
        if (bias) bm = fmaxf(bm, fabsf(bias[n]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, o));
    if (lane == 0) {
        red[0][warp] = l1;
        red[1][warp] = wm;
        red[2][warp] = bm;
    }
    __syncthreads();
    if (threadIdx.x == 0)
        for (int i = 1; i < 32; ++i)
            for (int j = 0; j < 3; ++j) red[j][0] = fmaxf(red[j][0], red[j][i]);
    if (threadIdx.x == 0) {
        l1 = red[0][0] * 1.0001f;    // rounding slack of the fp32 sums
        wm = red[1][0];
        bm = red[2][0];
        float in_bound = in_floor;
        if (in_bound_dev) in_bound = fmaxf(in_bound, *in_bound_dev);
        in_bound = fminf(fmaxf(in_bound, 1e-30f), 1e30f);
        const float ca = pow2_floor_ratio(14.0f, in_bound);
        float cw = pow2_floor_ratio(14.0f, fmaxf(wm, 1e-30f));
        if (bm > 0.0f) cw = fminf(cw, pow2_floor_ratio(14.0f, fminf(bm * ca, 1e30f)));
        header[0] = ca;
        header[1] = cw;
        header[2] = 1.0f / (ca * cw);
        header[3] = in_bound * l1 + bm;
        header[4] = in_bound;
        header[5] = l1;
        header[6] = wm;
        header[7] = bm;
    }
}

__global__ void __launch_bounds__(1024) tc16_plan_kernel(const float* __restrict__ w, const float* __restrict__ bias, int n_out,
                                                         int n_in, const float* __restrict__ in_bound_dev, float in_floor,
                                                         float* __restrict__ header) {
    __shared__ float red[3][32];
    tc16_plan_body(w, bias, n_out, n_in, in_bound_dev, in_floor, header, red);
}

// clift_pack_linear_tc16_batch, launch 1: CTA c plans the jobs of chain c in table order - a chain is a stack of layers whose
// input bound is the previous layer's output bound (read back from the header that layer's plan has just written)
__global__ void __launch_bounds__(1024) tc16_plan_batch_kernel(const clift_tc16_job* __restrict__ jobs, int n_jobs) {
    __shared__ float red[3][32];
    for (int j = 0; j < n_jobs; ++j) {
        if (jobs[j].chain != (int)blockIdx.x) continue;
        const clift_tc16_job J = jobs[j];
        tc16_plan_body(J.w, J.bias, J.n_out, J.n_in, J.in_bound, J.in_bound_floor, reinterpret_cast<float*>(J.dst), red);
        __threadfence();
        __syncthreads();
    }
}

// W [out][in] * cw (+ bias * ca * cw in k row 0 of one more slab) -> per k-step slab of fp16 (hi, lo) pairs:
// [hi|lo][2 k-chunks][n_pad][8] for n_pad > 128 else [2 k-chunks][hi|lo][n_pad][8]; zero padded.
// `pair` (the CTA-pair copy, [rank][k-step][per-CTA slab], h = n_pad / 2):
//   n_pad > 128 : CTA r holds rows [r h, r h + h): [hi | lo][2 k-chunks][h][8]
//   n_pad <= 128: [Y: 2 k-chunks x n_pad rows][X: 2 k-chunks x h rows]; Y = hi (rank 0) / lo (rank 1) of all rows,
//                 X = hi of rows [r h, r h + h)
__device__ __forceinline__ void pack_linear_tc16_body(const float* __restrict__ w, const float* __restrict__ bias, int n_out,
                                                      int n_in, const float* __restrict__ header, __half* __restrict__ dst,
                                                      __half* __restrict__ pair, int n_pad, int slabs, int64_t idx) {
    const int64_t total = (int64_t)slabs * kStepK * n_pad;
    if (idx >= total) return;
    const float ca = header[0], cw = header[1];
    const int k = (int)(idx / n_pad), n = (int)(idx % n_pad);
    const int k_bias = (n_in + kStepK - 1) / kStepK * kStepK;
    float x = (k < n_in && n < n_out) ? w[(size_t)n * n_in + k] * cw : 0.0f;
    if (bias && k == k_bias && n < n_out) x = bias[n] * (ca * cw);
    const __half hi = __float2half_rn(x);
    const __half lo = __float2half_rn(x - __half2float(hi));
    const int slab = k / kStepK, kc = (k % kStepK) / 8, ki = k & 7;
    __half* base = dst + (size_t)slab * (2 * kStepK * n_pad);
    if (n_pad <= 128) {
        const size_t off = ((size_t)(kc * 2) * n_pad + n) * 8 + ki;
        base[off] = hi;
        base[(size_t)n_pad * 8 + off] = lo;
    } else {
        const size_t off = ((size_t)kc * n_pad + n) * 8 + ki;
        base[off] = hi;
        base[(size_t)kStepK * n_pad + off] = lo;
    }
    // CTA-pair copy
    const int h = n_pad / 2;
    const size_t pb = pair_slab_bytes(n_pad) / 2;                       // halves per k-step slab of one CTA
    __half* r0 = pair + (size_t)slab * pb;
    __half* r1 = pair + ((size_t)slabs + slab) * pb;
    if (n_pad > 128) {
        __half* b = (n < h ? r0 : r1);
        const size_t off = ((size_t)kc * h + (n % h)) * 8 + ki;
        b[off] = hi;
        b[(size_t)2 * h * 8 + off] = lo;
    } else {
        const size_t off_y = ((size_t)kc * n_pad + n) * 8 + ki;
        r0[off_y] = hi;
        r1[off_y] = lo;
        __half* b = (n < h ? r0 : r1) + (size_t)2 * n_pad * 8;
        b[((size_t)kc * h + (n % h)) * 8 + ki] = hi;
    }
}

__global__ void pack_linear_tc16_kernel(const float* __restrict__ w, const float* __restrict__ bias, int n_out, int n_in,
                                        const float* __restrict__ header, __half* __restrict__ dst, __half* __restrict__ pair,
                                        int n_pad, int slabs) {
    pack_linear_tc16_body(w, bias, n_out, n_in, header, dst, pair, n_pad, slabs, (int64_t)blockIdx.x * blockDim.x + threadIdx.x);
}

// clift_pack_linear_tc16_batch, launch 2: block b finds its job by binary search over the jobs' first-block prefix
__global__ void __launch_bounds__(256) pack_linear_tc16_batch_kernel(const clift_tc16_job* __restrict__ jobs, int n_jobs) {
    int lo = 0, hi = n_jobs - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (jobs[mid].first_block <= (int)blockIdx.x)
            lo = mid;
        else
            hi = mid - 1;
    }
    const clift_tc16_job J = jobs[lo];
    const int n_pad = (J.n_out + 31) & ~31, slabs = (J.n_in + kStepK - 1) / kStepK + (J.bias ? 1 : 0);
    const float* header = reinterpret_cast<const float*>(J.dst);
    __half* single = reinterpret_cast<__half*>(reinterpret_cast<float*>(J.dst) + kHeaderFloats);
    pack_linear_tc16_body(J.w, J.bias, J.n_out, J.n_in, header, single, single + (size_t)slabs * 2 * kStepK * n_pad, n_pad, slabs,
                          (int64_t)((int)blockIdx.x - J.first_block) * 256 + threadIdx.x);
}

// dst[slot] = max |x| of the six factor tensors of a VM set in one launch: blockIdx.y = tensor = slot (dst zeroed by the
// caller; |x| >= 0 so the uint order of the bits is the float order)
struct AbsmaxSix {
    const float* x[6];
    long long n[6];
};
__global__ void absmax6_kernel(const __grid_constant__ AbsmaxSix A, float* __restrict__ dst) {
    const float* __restrict__ x = A.x[blockIdx.y];
    const int64_t n = A.n[blockIdx.y];
    float m = 0.0f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(x[i]));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
    if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(reinterpret_cast<unsigned int*>(dst + blockIdx.y), __float_as_uint(m));
}

__global__ void factor_bound_kernel(float* __restrict__ scratch8) {   // [0..2] plane maxima, [3..5] line maxima -> [6]
    float b = 0.0f;
    for (int m = 0; m < 3; ++m) b = fmaxf(b, scratch8[m] * scratch8[3 + m]);
    scratch8[6] = b * 1.0001f;
}

}  // namespace
}  // namespace clift


namespace clift {

static bool tc16_add_stack(HeadsParams& P, const clift_mlp& m) {
    for (int l = 0; l < m.n_layers; ++l) {
        if (!m.w_tc16[l] || P.n_gemms >= kMaxGemms) return false;
        Gemm& g = P.g[P.n_gemms];
        g.meta = reinterpret_cast<const float*>(m.w_tc16[l]);
        g.w = reinterpret_cast<const __half*>(g.meta + kHeaderFloats);
        g.k_steps = (int)ceil_div(m.dims[l], kStepK);
        g.n_pad = (int)round_up(m.dims[l + 1], 32);
        g.has_bias = 1;
        g.w_pair = g.w + (size_t)(g.k_steps + 1) * 2 * kStepK * g.n_pad;
        g.a_rounds = (l == 0 || !P.stream) ? 1 : (int)ceil_div(P.g[P.n_gemms - 1].n_pad, kRoundCols);
        g.n_sets = P.use_sets ? std::max(1, std::min(std::min(4, 128 / g.n_pad), g.k_steps + g.has_bias)) : 1;
        if (m.dims[l] > kMaxK || g.n_pad > 256) return false;
        ++P.n_gemms;
    }
    return true;
}

// the bias-free basis Linear of a factor set (appearance, or a grid-mode head's own set) as the next GEMM of P
static bool tc16_add_basis(HeadsParams& P, const void* basis_tc16, int comps, int dim) {
    if (!basis_tc16 || P.n_gemms >= kMaxGemms) return false;
    Gemm& g = P.g[P.n_gemms];
    g.meta = reinterpret_cast<const float*>(basis_tc16);
    g.w = reinterpret_cast<const __half*>(g.meta + kHeaderFloats);
    g.k_steps = (int)ceil_div(3 * comps, kStepK);
    g.n_pad = (int)round_up(dim, 32);
    g.has_bias = 0;    // appearance_basis_mat / semantic_basis_mat / instance_basis_mat have bias=False (tensoRF.py:65,74,83)
    g.w_pair = g.w + (size_t)g.k_steps * 2 * kStepK * g.n_pad;
    g.a_rounds = 1;
    g.n_sets = P.use_sets ? std::max(1, std::min(std::min(4, 128 / g.n_pad), g.k_steps)) : 1;
    ++P.n_gemms;
    return true;
}

bool heads_tc16_available(const clift_field* f, int heads) {
    auto ok = [](const clift_mlp& m) {
        for (int l = 0; l < m.n_layers; ++l)
            if (!m.w_tc16[l] || m.dims[l] > kMaxK || m.dims[l + 1] > 256) return false;
        return m.n_layers >= 1;
    };
    // grid-mode head: fp16-split operand of its basis, factor set inside the gather's envelope
    auto grid_ok = [](const clift_grid_head& g) {
        return g.comps == 0 || (g.basis_tc16 && g.comps % 16 == 0 && 3 * g.comps <= kMaxK && g.dim <= 64);
    };
    if (f->num_classes > CLIFT_MAX_HEAD_OUT || f->dim_instance > CLIFT_MAX_HEAD_OUT) return false;
    if ((heads & CLIFT_HEAD_SEMANTIC) && (!ok(f->semantic) || !grid_ok(f->semantic_grid))) return false;
    if ((heads & CLIFT_HEAD_INSTANCE) &&
        (!ok(f->instance_fast) || (f->slow_fast && !ok(f->instance_slow)) || !grid_ok(f->instance_grid)))
        return false;
    if (heads & CLIFT_HEAD_RGB) {
        if (!ok(f->rgb) || !f->basis_tc16 || f->dim_appearance > 64 || f->appearance_comps % 16 || 3 * f->appearance_comps > kMaxK)
            return false;
        // the base-value staging of the rgb input lives behind the K rows of the first rgb GEMM, inside the A_hi region
        if (round_up(f->rgb.dims[0], kStepK) * kRows * 2 + (int64_t)(f->dim_appearance + 3) * kXbStride * 4 > kABytes) return false;
    }
    return true;
}

// Training forwards through the tensor-core kernel: possible when every stash block has a writer there - xyz stacks without
// positional encoding, hidden widths that are multiples of 32 (accumulator columns = stash rows), <= 32 classes.
bool heads_tc16_stash_ok(const clift_field* f, int heads) {
    auto hidden_ok = [](const clift_mlp& m) {
        for (int l = 1; l < m.n_layers; ++l)
            if (m.dims[l] % 32) return false;
        return true;
    };
    // grid-mode heads: the basis output (accumulator columns, padded to 32) is the stash block of the MLP input (16-row pad)
    auto grid_ok = [](const clift_grid_head& g) { return g.comps == 0 || round_up(g.dim, 16) == round_up(g.dim, 32); };
    if ((heads & CLIFT_HEAD_SEMANTIC) &&
        ((f->pe_sem != 0 && !f->semantic_grid.comps) || f->num_classes > 32 || !hidden_ok(f->semantic) || !grid_ok(f->semantic_grid)))
        return false;
    if ((heads & CLIFT_HEAD_INSTANCE) &&
        ((f->pe_ins != 0 && !f->instance_grid.comps) || !hidden_ok(f->instance_fast) ||
         (f->slow_fast && !hidden_ok(f->instance_slow)) || !grid_ok(f->instance_grid)))
        return false;
    if ((heads & CLIFT_HEAD_RGB) && !hidden_ok(f->rgb)) return false;
    const char* e = getenv("CLIFT_TRAIN_FWD_FMA");       // development switch: training forwards on the FP32-FMA kernel
    return !(e && atoi(e) != 0);
}

int launch_heads_forward_tc16(const clift_render_cfg* cfg, const clift_field* field, const float* rays, const Workspace& ws,
                              int64_t cap, int64_t n_rays, float* rgb_raw, float* sem_raw, float* ins, cudaStream_t stream,
                              const StashLayout* lay) {
    HeadsParams P;
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
    P.trace = get_tc_trace();
    {   // development switches; the defaults are the measured best on B200 (profiles/r01_tc16_ab.md):
        //   CLIFT_TC16_STREAM=1  stream hidden-layer operands out of the epilogues + ping-pong accumulators (slower: the MMA
        //                        phase is shared-memory-bandwidth bound, the epilogue's operand stores stretch it)
        //   CLIFT_TC16_PARK=0    gather the appearance products in one burst before the basis GEMM instead of under the MMAs
        //   CLIFT_TC16_SETS=1    spread small-N GEMMs over several accumulator column sets (no gain: they are issue-bound)
        const char* e_stream = getenv("CLIFT_TC16_STREAM");
        const char* e_park = getenv("CLIFT_TC16_PARK");
        const char* e_sets = getenv("CLIFT_TC16_SETS");
        const char* e_pair = getenv("CLIFT_TC16_PAIR");
        P.pair = e_pair && atoi(e_pair) != 0;
        P.stream = e_stream && atoi(e_stream) != 0;
        P.park = !P.stream && !(e_park && atoi(e_park) == 0);
        P.use_sets = (e_sets && atoi(e_sets) != 0) ? 1 : 0;
    }
    if (lay) {      // training forward: record the stash; only the default schedule has the writers
        P.pair = 0;
        P.stream = 0;
        P.park = 1;
        P.use_sets = 0;
        P.stash_a = ws.stash_a;
        P.rec_rgb = ws.rec_rgb;
        P.lay = *lay;
    }
    int heads = (sem_raw ? CLIFT_HEAD_SEMANTIC : 0) | (ins ? CLIFT_HEAD_INSTANCE : 0) | (rgb_raw ? CLIFT_HEAD_RGB : 0);
    bool ok = heads_tc16_available(field, heads) && (!lay || heads_tc16_stash_ok(field, heads));
    // a grid-mode head's stack = [basis of its factor set, MLP layers]; n_sem / n_ins count the GEMMs of one stack
    if (ok && sem_raw) {
        const clift_grid_head& gh = field->semantic_grid;
        P.n_sem = field->semantic.n_layers + (gh.comps ? 1 : 0);
        if (gh.comps) {
            P.semg = make_grid_factors(field, gh);
            ok = ok && tc16_add_basis(P, gh.basis_tc16, gh.comps, gh.dim);
        }
        ok = ok && tc16_add_stack(P, field->semantic);
    }
    if (ok && ins) {
        const clift_grid_head& gh = field->instance_grid;
        P.n_ins = field->instance_fast.n_layers + (gh.comps ? 1 : 0);
        if (gh.comps) P.insg = make_grid_factors(field, gh);
        for (int net = 0; net < (field->slow_fast ? 2 : 1); ++net) {
            if (gh.comps) ok = ok && tc16_add_basis(P, gh.basis_tc16, gh.comps, gh.dim);
            ok = ok && tc16_add_stack(P, net == 0 ? field->instance_fast : field->instance_slow);
        }
    }
    if (ok && rgb_raw) {
        P.n_rgb = field->rgb.n_layers;
        ok = ok && tc16_add_basis(P, field->basis_tc16, field->appearance_comps, field->dim_appearance);
        ok = ok && tc16_add_stack(P, field->rgb);
    }
    if (!ok) {
        set_error("launch_heads_forward_tc16: field lacks fp16 tensor-core operands (clift_pack_linear_tc16) or exceeds the envelope");
        return CLIFT_ERR_UNSUPPORTED;
    }
    if (P.n_gemms == 0 || n_rays <= 0) return CLIFT_OK;
    const int grid = P.pair ? (sm_count() & ~1) : sm_count();
    auto launch = [&](auto kernel) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
        if (e != cudaSuccess) return e;
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3((unsigned)grid);
        lc.blockDim = dim3(kThreads);
        lc.dynamicSmemBytes = kSmemBytes;
        lc.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = P.pair ? 2 : 1;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        lc.attrs = attr;
        lc.numAttrs = 1;
        return cudaLaunchKernelEx(&lc, kernel, P);
    };
#define CLIFT_TC16_CASE(NV)                                                                                            \
    case NV: {                                                                                                         \
        CLIFT_CUDA(lay ? launch(heads_tc16_forward_kernel<NV, false, true>)                                            \
                       : (P.pair ? launch(heads_tc16_forward_kernel<NV, true, false>)                                  \
                                 : launch(heads_tc16_forward_kernel<NV, false, false>)));                              \
        break;                                                                                                         \
    }
    switch (P.app.comps / 16) {
        CLIFT_TC16_CASE(1)
        CLIFT_TC16_CASE(2)
        CLIFT_TC16_CASE(3)
        CLIFT_TC16_CASE(4)
        default:
            set_error("launch_heads_forward_tc16: appearance_comps %d not in {16,32,48,64}", P.app.comps);
            return CLIFT_ERR_UNSUPPORTED;
    }
#undef CLIFT_TC16_CASE
    CLIFT_AFTER_LAUNCH("heads_tc16_forward_kernel");
    return CLIFT_OK;
}

}  // namespace clift

using namespace clift;

extern "C" int64_t clift_tc16_weight_bytes(int32_t n_out, int32_t n_in, int32_t has_bias) {
    if (n_out <= 0 || n_in <= 0 || n_out > 256 || n_in > kMaxK) return CLIFT_ERR_UNSUPPORTED;
    const int n_pad = (int)round_up(n_out, 32), slabs = (int)ceil_div(n_in, kStepK) + (has_bias ? 1 : 0);
    return (int64_t)kHeaderFloats * 4 + (int64_t)slabs * 2 * kStepK * n_pad * 2 + (int64_t)2 * slabs * pair_slab_bytes(n_pad);
}

extern "C" int32_t clift_pack_linear_tc16(const float* w, const float* bias, void* dst, int32_t n_out, int32_t n_in,
                                          const float* in_bound, float in_bound_floor, void* stream) {
    CLIFT_CHECK_ARG(w && dst, "null pointer");
    CLIFT_CHECK_ARG(((uintptr_t)dst & 15) == 0, "dst must be 16-byte aligned");
    CLIFT_CHECK_ARG(in_bound || in_bound_floor > 0.0f, "an input bound is required (device scalar and/or a positive floor)");
    CLIFT_CHECK_SUPPORTED(n_out > 0 && n_in > 0 && n_out <= 256 && n_in <= kMaxK, "layer wider than 256");
    const int n_pad = (int)round_up(n_out, 32), slabs = (int)ceil_div(n_in, kStepK) + (bias ? 1 : 0);
    float* header = reinterpret_cast<float*>(dst);
    tc16_plan_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(w, bias, n_out, n_in, in_bound, in_bound_floor, header);
    CLIFT_AFTER_LAUNCH("tc16_plan_kernel");
    const int64_t total = (int64_t)slabs * kStepK * n_pad;
    __half* single = reinterpret_cast<__half*>(header + kHeaderFloats);
    pack_linear_tc16_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(
        w, bias, n_out, n_in, header, single, single + (size_t)slabs * 2 * kStepK * n_pad, n_pad, slabs);
    CLIFT_AFTER_LAUNCH("pack_linear_tc16_kernel");
    return CLIFT_OK;
}

extern "C" int32_t clift_pack_linear_tc16_batch(const clift_tc16_job* jobs, int32_t n_jobs, int32_t n_chains, int32_t total_blocks,
                                                void* stream) {
    CLIFT_CHECK_ARG(n_jobs >= 0 && n_chains >= 0 && total_blocks >= 0 && (n_jobs == 0 || jobs), "null table or negative size");
    if (n_jobs == 0 || n_chains == 0 || total_blocks == 0) return CLIFT_OK;
    tc16_plan_batch_kernel<<<(unsigned)n_chains, 1024, 0, (cudaStream_t)stream>>>(jobs, n_jobs);
    CLIFT_AFTER_LAUNCH("tc16_plan_batch_kernel");
    pack_linear_tc16_batch_kernel<<<(unsigned)total_blocks, 256, 0, (cudaStream_t)stream>>>(jobs, n_jobs);
    CLIFT_AFTER_LAUNCH("pack_linear_tc16_batch_kernel");
    return CLIFT_OK;
}

extern "C" int32_t clift_tc16_factor_bound(const float* const* planes3, const float* const* lines3, const int64_t* plane_elems3,
                                           const int64_t* line_elems3, float* scratch8, void* stream) {
    CLIFT_CHECK_ARG(planes3 && lines3 && plane_elems3 && line_elems3 && scratch8, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    CLIFT_CUDA(cudaMemsetAsync(scratch8, 0, 8 * sizeof(float), st));
    AbsmaxSix A;
    int64_t largest = 1;
    for (int m = 0; m < 3; ++m) {
        CLIFT_CHECK_ARG(planes3[m] && lines3[m] && plane_elems3[m] > 0 && line_elems3[m] > 0, "null factor or empty factor");
        A.x[m] = planes3[m];
        A.n[m] = plane_elems3[m];
        A.x[3 + m] = lines3[m];
        A.n[3 + m] = line_elems3[m];
        largest = std::max<int64_t>(largest, std::max(plane_elems3[m], line_elems3[m]));
    }
    const int gx = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(largest, 256 * 8), 2 * sm_count()));
    absmax6_kernel<<<dim3((unsigned)gx, 6), 256, 0, st>>>(A, scratch8);      // slots 0-2 planes, 3-5 lines
    CLIFT_AFTER_LAUNCH("absmax6_kernel");
    factor_bound_kernel<<<1, 1, 0, st>>>(scratch8);
    CLIFT_AFTER_LAUNCH("factor_bound_kernel");
    return CLIFT_OK;
}

extern "C" int32_t clift_debug_tc16_gemm(const float* a, const void* w_tc16, float* out, int32_t k, int32_t n_out,
                                         int32_t has_bias, void* stream) {
    CLIFT_CHECK_ARG(a && w_tc16 && out, "null pointer");
    CLIFT_CHECK_SUPPORTED(n_out > 0 && k > 0 && n_out <= 256 && k <= kMaxK, "layer wider than 256");
    Gemm g;
    g.meta = reinterpret_cast<const float*>(w_tc16);
    g.w = reinterpret_cast<const __half*>(g.meta + kHeaderFloats);
    g.k_steps = (int)ceil_div(k, kStepK);
    g.n_pad = (int)round_up(n_out, 32);
    g.has_bias = has_bias ? 1 : 0;
    g.w_pair = nullptr;
    g.a_rounds = 1;
    g.n_sets = std::max(1, std::min(std::min(4, 128 / g.n_pad), g.k_steps + g.has_bias));
    CLIFT_CUDA(cudaFuncSetAttribute(tc16_gemm_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    tc16_gemm_test_kernel<<<1, kThreads, kSmemBytes, (cudaStream_t)stream>>>(a, k, g, out);
    CLIFT_AFTER_LAUNCH("tc16_gemm_test_kernel");
    return CLIFT_OK;
}
