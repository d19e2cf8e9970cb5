// Pipelined tensor-core evaluation of the xyz-input MLP stacks (semantic: 3 -> 256^4 -> C, instance fast / slow: 3 -> 256^3 -> d;
// tensoRF.py:462-511, 565-594; renderer:119-131, 137-156) on the compacted active samples - the inference default.
//
// Same arithmetic as heads_tc16.cu (tcgen05 kind::f16, 3-product fp16 split: A_hi*W_hi + A_hi*W_lo + A_lo*W_hi, fp32
// accumulation in tensor memory, power-of-two scale chain from clift_pack_linear_tc16), different schedule: there every GEMM
// of a tile ran strictly after the previous layer's epilogue (tensor pipe idle through every epilogue, row warps idle through
// every MMA phase).  Here a 256-wide layer is issued as two N = 128 accumulator units that ping-pong between two 128-column
// accumulators, the operand of the NEXT layer is double buffered, and its hi half lives in tensor memory:
//
//   tensor memory (512 columns): D0 | D1 (2 x 128 fp32 accumulator columns) | AH0 | AH1 (2 x 128 columns: A_hi of a K = 256
//                 operand, two fp16 per column - read by the .ts MMA form, written by the epilogues with tcgen05.st)
//   shared memory: AL0 | AL1 (2 x 64 KB: A_lo, canonical K-major no-swizzle UMMA layout), an 8-stage x 8 KB weight ring
//                 (one stage = one k-step of one N-half: W_hi 4 KB | W_lo 4 KB, 1-D bulk TMA + mbarriers), the xyz operand of
//                 the tile (all three stacks read it), a constant "ones" chunk for the bias step, the final-layer scratch
//
// so while the tensor pipe runs unit (L, h = 1) the row warps drain unit (L, 0) into the k rows [0, 128) of layer L + 1's
// operand, the first eight k-steps of (L + 1, 0) run under the epilogue of (L, 1), and the final-layer softmax / per-ray sums
// of one stack run under the first hidden layer of the next.  MMA operand traffic through shared memory drops as well: two
// of the three products read A from tensor memory.
//
//   warp 0 : weight producer (one lane)      warp 1 : MMA issuer (one lane)
//   warps 2..13 : 384 row threads, three per record (thread <-> TMEM lane of its warp's quarter)
//
// CTA pairs (default; CLIFT_X16_PAIR=0 for single CTAs): the kernel is launched as clusters of two, each CTA works on its own
// tile, and the leader's issuer drives both tensor cores with cta_group::2 MMAs (M = 256): every SM then stages only HALF
// of each weight stage (64 of the unit's 128 B rows).  This matters because the weight stream, not the tensor pipe, bounds
// the single-CTA schedule: a pipelined 256-wide layer needs 16 KB of weights per 384 tensor cycles = 43 B/clk per SM, and
// an SM ingests ~34 B/clk from L2 when all 148 stream at once (measured: profiles/r02_x16_trace_single.txt - every unit
// takes 17 stages x 241 cycles whatever the schedule; replicating the weights 8x in L2 changes nothing, so it is the SM's
// ingest rate, not L2 slice contention).
//
// Envelope (else the serial kernel in heads_tc16.cu runs): xyz stacks without positional encoding, hidden width 256,
// >= 3 layers, <= 32 outputs.  The rgb stack (appearance gather, basis, positional encoding) stays on heads_tc16.cu.
#include <cuda_fp16.h>

#include "launchers.h"
#include "tcgen05.cuh"

namespace clift {

long long* get_tc_trace();   // heads_tc.cu (clift_debug_tc_trace)

namespace {

constexpr int kTraceUnits = 32, kTraceSlots = 8;   // clock64 trace of CTA 0: [tile < 4][unit][slot]
constexpr int kRows = 128;
constexpr int kParts = 3;
constexpr int kRowThreads = kRows * kParts;
constexpr int kThreads = 64 + kRowThreads;
constexpr int kRingBytes = 65536;
template <bool kPair>
struct Cfg {
    static constexpr int kBlockBytes = kPair ? 4096 : 8192;       // one k-step of one accumulator unit (this CTA's share)
    static constexpr int kStageBytes = 2 * kBlockBytes;           // a ring stage = two k-steps: the per-stage issue overhead
                                                                  // (barrier wait, commit, descriptors) needs ~6 MMAs to hide
    static constexpr int kStages = kRingBytes / kStageBytes;
    static constexpr uint32_t kBRows = kPair ? 64u : 128u;         // B rows of a unit staged by this CTA
    static constexpr int kM = kPair ? 256 : 128;
    static constexpr uint32_t kArrivals = kPair ? 2u * 384u : 384u;
};
constexpr int kMaxStagesAny = kRingBytes / 8192;
constexpr int kChunkBytes = kRows * 16;             // one k-chunk (8 k values) of all 128 rows
constexpr int kALoBytes = 256 / 8 * kChunkBytes;    // 64 KB
constexpr int kMaxStacks = 3;
constexpr int kMaxGemms = 16;
constexpr int kHeaderFloats = 16;
constexpr int kScratchRows = 32;
constexpr int kScrStride = kScratchRows + 1;   // scratch is [record][channel], odd stride: the per-ray run sums read (run, channel)
                                              // items channel-fastest, which at [channel][record] put 8 items of a warp on ONE bank
                                              // and slowed the MMAs running beside them (A_lo / B operand reads share the banks)
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColD = 0, kColAH = 256;

enum { kL0 = 0, kHid = 1, kFin = 2 };

struct XGemm {
    const unsigned char* w;   // weight stream: x16 stages (kL0 / kHid) or clift_pack_linear_tc16 stacked slabs (kFin)
    const float* meta;        // clift_pack_linear_tc16 header (ca, cw, 1/(ca cw), ...)
    int kind, n_pad, k_steps;
    const unsigned char* w_pair;   // the same weights in CTA-pair layout (rank-major), see pack_x16_kernel / heads_tc16.cu
};

struct XStack {
    int n_layers, g_first, n_out, is_sem, out_col0;
    int stash_id;      // training: StashLayout id of the stack (0 semantic, 1 instance fast, 2 instance slow)
};

struct XParams {
    const float4* rec_pos;
    const int32_t* rec_ray;
    const unsigned long long* stats;
    long long cap;
    int n_stacks, n_gemms;
    XStack st[kMaxStacks];
    XGemm g[kMaxGemms];
    int n_cls, softmax, ins_width;
    float* sem_raw;
    float* ins;
    long long* trace;
    // training forwards (save_for_backward): every layer input is also written, unscaled fp32, to the A-stash the backward
    // kernels read (StashLayout / stash_idx in common.cuh), the softmax probabilities to its prob block; null = inference
    float* stash_a;
    StashLayout lay;
};

struct Smem {
    unsigned char* a_lo;    // [2][kALoBytes]
    unsigned char* w;       // [kStages][kStageBytes]
    unsigned char* xyz;     // hi [2 chunks] | lo [2 chunks]
    unsigned char* ones;    // [2 chunks], element 0 of every row = 1
    float* scratch;         // [kRows][kScrStride]
    int* ray;               // [kRows]
    int* runs;              // [kRows + 1]
    int* n_runs;
    float2* sc;             // [kMaxGemms]
    uint64_t* full;         // [kMaxStagesAny]
    uint64_t* empty;        // [kMaxStagesAny]
    uint64_t* full_peer;    // [kMaxStagesAny] CTA pairs, leader: the peer's share of the stage has landed
    uint64_t* a_ready;      // [2 buffers][2 halves]   (CTA pairs: the leader's barriers collect both CTAs' row threads)
    uint64_t* d_full;       // [2]
    uint64_t* d_free;       // [2]
    uint64_t* xyz_ready;
    uint32_t* tmem_base;
};

constexpr size_t kSmemBytes = 2 * (size_t)kALoBytes + (size_t)kRingBytes + 4 * kChunkBytes + 2 * kChunkBytes +
                              (size_t)kScrStride * kRows * 4 + (2 * kRows + 8) * 4 + kMaxGemms * 8 + (3 * kMaxStagesAny + 9) * 8 + 64;
static_assert(kSmemBytes <= 232448, "shared memory budget of one sm_100a CTA");

__device__ __forceinline__ Smem carve_smem(unsigned char* raw) {
    Smem s;
    s.a_lo = raw;
    s.w = s.a_lo + 2 * kALoBytes;
    s.xyz = s.w + (size_t)kRingBytes;
    s.ones = s.xyz + 4 * kChunkBytes;
    s.scratch = reinterpret_cast<float*>(s.ones + 2 * kChunkBytes);
    s.ray = reinterpret_cast<int*>(s.scratch + kScrStride * kRows);
    s.runs = s.ray + kRows;
    s.n_runs = s.runs + kRows + 1;
    s.sc = reinterpret_cast<float2*>(s.n_runs + 7);
    s.full = reinterpret_cast<uint64_t*>(s.sc + kMaxGemms);
    s.empty = s.full + kMaxStagesAny;
    s.full_peer = s.empty + kMaxStagesAny;
    s.a_ready = s.full_peer + kMaxStagesAny;
    s.d_full = s.a_ready + 4;
    s.d_free = s.d_full + 2;
    s.xyz_ready = s.d_free + 2;
    s.tmem_base = reinterpret_cast<uint32_t*>(s.xyz_ready + 1);
    return s;
}

template <bool kPair>
struct Ring {
    int stage = 0;
    uint32_t phase = 0;
    __device__ __forceinline__ void advance() {
        if (++stage == Cfg<kPair>::kStages) {
            stage = 0;
            phase ^= 1u;
        }
    }
};

// ---- MMA / barrier wrappers, single CTA or CTA pair -----------------------------------------------------------
template <bool kPair>
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (kPair)
        tc::mma_ss_f16_pair(d, a, b, idesc, acc);
    else
        tc::mma_ss_f16(d, a, b, idesc, acc);
}
// D[tmem] (+)= A[tmem] * B[smem]^T, fp16 operands (A: 128 lanes x 8 columns per k-step, two fp16 per column)
template <bool kPair>
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    if (kPair)
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n"
            "}\n" ::"r"(d_tmem),
            "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
    else
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
            "}\n" ::"r"(d_tmem),
            "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
}
template <bool kPair>
__device__ __forceinline__ void commit(uint64_t* bar) {   // CTA pairs: the same barrier offset in both CTAs
    if (kPair)
        tc::mma_commit_pair(bar, 3);
    else
        tc::mma_commit(bar);
}
// issuer-side wait on a barrier the row threads of (both) CTA(s) arrive on
// (CTA pairs: plain CTA-scope waits and relaxed remote arrives on purpose.  What the barriers order is tensor-core / TMA
// traffic on each CTA's OWN shared and tensor memory - covered by fence.proxy.async and the tcgen05 fences on both sides; no
// generic-proxy data crosses the CTAs.  The cluster-scope acquire / release forms cost ~700 cycles per weight stage and
// ~1000 per accumulator hand-off here (profiles/r02_x16_ab.md), which made the pair schedule 2.5x slower than single CTAs.)
template <bool kPair>
__device__ __forceinline__ void wait_rows(uint64_t* bar, uint32_t parity) {
    tc::mbar_wait(bar, parity);
    tc::fence_after_sync();
}
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// row-thread arrive on the barrier the issuer waits on (CTA pairs: it lives in the leader CTA)
template <bool kPair>
__device__ __forceinline__ void arrive_rows(uint64_t* local_bar, uint32_t leader_addr) {
    if (kPair)
        mbar_arrive_remote_relaxed(leader_addr);
    else
        tc::mbar_arrive(local_bar);
}

__device__ __forceinline__ uint32_t h2_bits(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(x, y);
    const float2 b = __half22float2(h);
    hi = h2_bits(h);
    lo = h2_bits(__floats2half2_rn(x - b.x, y - b.y));
}

// final layer (n_pad <= 128), bytes of one k-step's slab staged by ONE CTA: clift_pack_linear_tc16's stacked slab
// ([2 k-chunks][hi | lo][n_pad][8] = 64 n_pad) or its CTA-pair form ([Y: 2 x n_pad rows][X: 2 x n_pad/2 rows] = 48 n_pad)
template <bool kPair>
__device__ __forceinline__ uint32_t fin_kstep_bytes(int n_pad) { return (kPair ? 48u : 64u) * (uint32_t)n_pad; }
template <bool kPair>
__device__ __forceinline__ int fin_ksteps_per_stage(int n_pad) {
    // four k-steps per stage whenever they fit (always for n_pad <= 64): the final layer's issue loop is then four fixed
    // groups of 8 MMAs + the bias step, with no per-k-step stage bookkeeping
    return min(4, Cfg<kPair>::kStageBytes / (int)fin_kstep_bytes<kPair>(n_pad));
}
template <bool kPair>
__device__ __forceinline__ int gemm_stages(const XGemm& g) {
    if (g.kind != kFin) return 2 * ((g.k_steps + 2) / 2);
    const int per = fin_ksteps_per_stage<kPair>(g.n_pad);
    return (g.k_steps + 1 + per - 1) / per;
}

// ---- warp 0, one lane ------------------------------------------------------------------------------------
template <bool kPair>
__device__ __forceinline__ void produce(const Smem& s, const XGemm& g, Ring<kPair>& r, uint32_t rank) {
    constexpr int kSB = Cfg<kPair>::kStageBytes;
    if (g.kind != kFin) {
        // per N-half: k_steps + 1 blocks (the bias block last), two blocks per stage, the last stage possibly a single block
        constexpr int kBB = Cfg<kPair>::kBlockBytes;
        const int blocks = g.k_steps + 1;
        const unsigned char* src = kPair ? g.w_pair + (size_t)rank * 2 * blocks * kBB : g.w;
        for (int h = 0; h < 2; ++h)
            for (int b0 = 0; b0 < blocks; b0 += 2, r.advance()) {
                const uint32_t bytes = (uint32_t)min(2, blocks - b0) * kBB;
                tc::mbar_wait(&s.empty[r.stage], r.phase ^ 1);
                tc::mbar_arrive_expect_tx(&s.full[r.stage], bytes);
                tc::bulk_load(s.w + (size_t)r.stage * kSB, src + (size_t)(h * blocks + b0) * kBB, bytes, &s.full[r.stage]);
            }
        return;
    }
    const int per = fin_ksteps_per_stage<kPair>(g.n_pad), steps = g.k_steps + 1;
    const uint32_t kstep_bytes = fin_kstep_bytes<kPair>(g.n_pad);
    const unsigned char* src = kPair ? g.w_pair + (size_t)rank * steps * kstep_bytes : g.w;
    for (int k0 = 0; k0 < steps; k0 += per, r.advance()) {
        const uint32_t bytes = (uint32_t)min(per, steps - k0) * kstep_bytes;
        tc::mbar_wait(&s.empty[r.stage], r.phase ^ 1);
        tc::mbar_arrive_expect_tx(&s.full[r.stage], bytes);
        tc::bulk_load(s.w + (size_t)r.stage * kSB, src + (size_t)k0 * kstep_bytes, bytes, &s.full[r.stage]);
    }
}

// CTA pairs, peer CTA's warp 1, one lane: tell the leader's issuer when this CTA's share of a weight stage has landed
__device__ __forceinline__ void relay(const Smem& s, int stages, Ring<true>& r, uint32_t leader_full_peer) {
    for (int i = 0; i < stages; ++i, r.advance()) {
        tc::mbar_wait(&s.full[r.stage], r.phase);
        mbar_arrive_remote_relaxed(leader_full_peer + 8u * (uint32_t)r.stage);
    }
}

// ---- warp 1, one lane (CTA pairs: the leader's) ---------------------------------------------------------------
// Issue-rate notes (profiles/r02_x16_ab.md): the issuing lane retires ~50 dependent instructions per ring stage (barrier
// wait, descriptor arithmetic, R2UR moves of every MMA operand, commit) = ~240 cycles whatever the MMA shapes are.  With one
// k-step of an N = 128 unit per stage (3 MMAs = 192 tensor cycles) the issuer, not the tensor pipe, set the pace; two k-steps
// per stage (6 MMAs = 384 tensor cycles) put it back under the tensor time.  Running the loops on the whole warp with only
// the MMAs predicated was slower still (412 cycles per stage: 32 lanes polling the mbarriers).
template <bool kPair>
struct Issuer {
    static constexpr uint32_t kMask = Cfg<kPair>::kStages - 1;
    static constexpr uint32_t kShift = Cfg<kPair>::kStages == 8 ? 3 : 2;
    static_assert((Cfg<kPair>::kStages & (Cfg<kPair>::kStages - 1)) == 0, "ring stages: power of two");
    uint32_t cnt = 0;         // weight stages consumed so far
    uint32_t unit = 0;        // accumulator units issued so far (unit u writes D[u & 1])
    uint32_t par_a = 0;       // parity bit per a_ready barrier
    int cur = 0;              // operand buffer written by the most recent hidden epilogue
    uint32_t w16 = 0;         // shared-memory address of the ring >> 4
    long long* tr = nullptr;  // trace row of the running tile (null: off)
    int tu = 0;               // unit index inside the tile
    __device__ __forceinline__ void stamp(int slot) {
        if (tr && tu < kTraceUnits) tr[tu * kTraceSlots + slot] = clock64();
    }
};

__device__ __forceinline__ uint64_t mk_desc(uint32_t lo) {
    constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);     // SBO 128 B, descriptor version 1
    return ((uint64_t)kDescHi << 32) | lo;
}

template <bool kPair>
__device__ __forceinline__ void wait_d_free(const Smem& s, uint32_t unit) {
    if (unit >= 2) wait_rows<kPair>(&s.d_free[unit & 1], ((unit >> 1) - 1) & 1u);
}

template <bool kPair>
__device__ __forceinline__ void wait_a_ready(const Smem& s, Issuer<kPair>& I, int buf, int half) {
    const int b = buf * 2 + half;
    wait_rows<kPair>(&s.a_ready[b], (I.par_a >> b) & 1u);
    I.par_a ^= 1u << b;
}

// the weight stage at the ring head has landed (in both CTAs) -> its shared-memory address >> 4
template <bool kPair>
__device__ __forceinline__ uint32_t wait_stage(const Smem& s, const Issuer<kPair>& I, uint32_t& st) {
    st = I.cnt & Issuer<kPair>::kMask;
    const uint32_t ph = (I.cnt >> Issuer<kPair>::kShift) & 1u;
    tc::mbar_wait(&s.full[st], ph);
    if (kPair) tc::mbar_wait(&s.full_peer[st], ph);
    tc::fence_after_sync();
    return I.w16 + st * (uint32_t)(Cfg<kPair>::kStageBytes >> 4);
}

// one x16 block (W_hi | W_lo of one k-step of one unit) at `w`: the three products
template <bool kPair, bool kATmem>
__device__ __forceinline__ void block_mmas(uint32_t d, uint32_t w, uint32_t a_hi, uint64_t a_hi_desc, uint64_t a_lo_desc,
                                           uint32_t idesc, uint32_t acc) {
    constexpr uint32_t kB = Cfg<kPair>::kBRows;
    const uint64_t bh = mk_desc(w | (kB << 16)), bl = mk_desc((w + 2u * kB) | (kB << 16));
    if (kATmem) {
        mma_ts<kPair>(d, a_hi, bh, idesc, acc);
        mma_ts<kPair>(d, a_hi, bl, idesc, 1u);
    } else {
        mma_ss<kPair>(d, a_hi_desc, bh, idesc, acc);
        mma_ss<kPair>(d, a_hi_desc, bl, idesc, 1u);
    }
    mma_ss<kPair>(d, a_lo_desc, bh, idesc, 1u);
}
// the bias block: (ones chunk) x (b_hi | b_lo)
template <bool kPair>
__device__ __forceinline__ void bias_mmas(uint32_t d, uint32_t w, uint64_t ones_desc, uint32_t idesc) {
    constexpr uint32_t kB = Cfg<kPair>::kBRows;
    mma_ss<kPair>(d, ones_desc, mk_desc(w | (kB << 16)), idesc, 1u);
    mma_ss<kPair>(d, ones_desc, mk_desc((w + 2u * kB) | (kB << 16)), idesc, 1u);
}

// first layer (K = 16: one stage per N-half = the k-step block + the bias block), A = the tile's xyz operand in shared memory
template <bool kPair>
__device__ __forceinline__ void issue_l0(const Smem& s, Issuer<kPair>& I, uint32_t tmem) {
    constexpr uint32_t kBlk16 = Cfg<kPair>::kBlockBytes >> 4;
    const uint32_t idesc = tc::make_idesc_f16(Cfg<kPair>::kM, 128);
    const uint32_t a_lbo = (kChunkBytes >> 4) << 16;
    const uint64_t xh = mk_desc((tc::smem_addr(s.xyz) >> 4) | a_lbo), xl = mk_desc((tc::smem_addr(s.xyz + 2 * kChunkBytes) >> 4) | a_lbo);
    const uint64_t on = mk_desc((tc::smem_addr(s.ones) >> 4) | a_lbo);
    for (int h = 0; h < 2; ++h) {
        wait_d_free<kPair>(s, I.unit);
        I.stamp(0);
        const uint32_t d = tmem + kColD + (I.unit & 1u) * 128u;
        uint32_t st;
        const uint32_t w = wait_stage<kPair>(s, I, st);
        block_mmas<kPair, false>(d, w, 0u, xh, xl, idesc, 0u);
        bias_mmas<kPair>(d, w + kBlk16, on, idesc);
        commit<kPair>(&s.empty[st]);
        ++I.cnt;
        commit<kPair>(&s.d_full[I.unit & 1u]);
        I.stamp(2);
        ++I.unit;
        ++I.tu;
    }
    I.cur ^= 1;
}

// hidden layer (K = 256, N = 256 as two units): A_hi from tensor memory, A_lo from shared memory; k_steps is even
template <bool kPair>
__device__ __forceinline__ void issue_hidden(const Smem& s, Issuer<kPair>& I, uint32_t tmem, int k_steps) {
    constexpr uint32_t kBlk16 = Cfg<kPair>::kBlockBytes >> 4;
    constexpr uint32_t kAl = 2u * (kChunkBytes >> 4);          // A_lo descriptor advance per k-step (two k-chunks)
    const uint32_t idesc = tc::make_idesc_f16(Cfg<kPair>::kM, 128);
    const uint32_t a_lbo = (kChunkBytes >> 4) << 16;
    const int in = I.cur;
    const uint32_t ah0 = tmem + kColAH + (uint32_t)in * 128u;
    const uint32_t al0 = (tc::smem_addr(s.a_lo + (size_t)in * kALoBytes) >> 4) | a_lbo;
    const uint64_t on = mk_desc((tc::smem_addr(s.ones) >> 4) | a_lbo);
    const int pairs = k_steps >> 1, mid = min(4, pairs);
    for (int h = 0; h < 2; ++h) {
        wait_d_free<kPair>(s, I.unit);
        I.stamp(0);
        const uint32_t d = tmem + kColD + (I.unit & 1u) * 128u;
        uint32_t ah = ah0, al = al0, st;
        if (h == 0) wait_a_ready<kPair>(s, I, in, 0);
        I.stamp(1);
        for (int i = 0; i < mid; ++i, ah += 16u, al += 2u * kAl, ++I.cnt) {
            const uint32_t w = wait_stage<kPair>(s, I, st);
            block_mmas<kPair, true>(d, w, ah, 0ull, mk_desc(al), idesc, i > 0 ? 1u : 0u);
            block_mmas<kPair, true>(d, w + kBlk16, ah + 8u, 0ull, mk_desc(al + kAl), idesc, 1u);
            commit<kPair>(&s.empty[st]);
        }
        if (h == 0) {
            I.stamp(6);
            wait_a_ready<kPair>(s, I, in, 1);
            I.stamp(7);
        }
        for (int i = mid; i < pairs; ++i, ah += 16u, al += 2u * kAl, ++I.cnt) {
            const uint32_t w = wait_stage<kPair>(s, I, st);
            block_mmas<kPair, true>(d, w, ah, 0ull, mk_desc(al), idesc, 1u);
            block_mmas<kPair, true>(d, w + kBlk16, ah + 8u, 0ull, mk_desc(al + kAl), idesc, 1u);
            commit<kPair>(&s.empty[st]);
        }
        {
            const uint32_t w = wait_stage<kPair>(s, I, st);
            bias_mmas<kPair>(d, w, on, idesc);
            commit<kPair>(&s.empty[st]);
            ++I.cnt;
        }
        commit<kPair>(&s.d_full[I.unit & 1u]);
        I.stamp(2);
        ++I.unit;
        ++I.tu;
    }
    I.cur ^= 1;
}

// final layer (K = 256, n_pad <= 128).  Single CTA: clift_pack_linear_tc16 "stacked" slabs [2 k-chunks][hi | lo][n_pad][8]:
// A_hi*[W_hi ; W_lo] is one MMA into columns [0, 2 n_pad), A_lo*W_hi adds into [0, n_pad); the epilogue sums the two blocks.
// CTA pairs: per-CTA slab [Y: 2 k-chunks x n_pad rows][X: 2 k-chunks x n_pad/2 rows], Y = W_hi in the leader and W_lo in the
// peer (so the pair's B of the first MMA is [W_hi ; W_lo] again), X = this CTA's half of W_hi for A_lo*W_hi.
template <bool kPair>
__device__ __forceinline__ void issue_final(const Smem& s, Issuer<kPair>& I, uint32_t tmem, const XGemm& g) {
    const int n_pad = g.n_pad, k_steps = g.k_steps, steps = k_steps + 1, per = fin_ksteps_per_stage<kPair>(n_pad);
    const uint32_t idesc = tc::make_idesc_f16(Cfg<kPair>::kM, n_pad), idesc2 = tc::make_idesc_f16(Cfg<kPair>::kM, 2 * n_pad);
    const uint32_t a_lbo = (kChunkBytes >> 4) << 16;
    const int in = I.cur;
    uint32_t ah = tmem + kColAH + (uint32_t)in * 128u;
    uint32_t al = (tc::smem_addr(s.a_lo + (size_t)in * kALoBytes) >> 4) | a_lbo;
    const uint64_t on = mk_desc((tc::smem_addr(s.ones) >> 4) | a_lbo);
    const uint32_t rows1 = kPair ? (uint32_t)n_pad : 2u * (uint32_t)n_pad;     // rows per k-chunk of the first B operand
    const uint32_t off2 = kPair ? 2u * rows1 : 0u, rows2 = kPair ? (uint32_t)n_pad / 2 : rows1;
    const uint32_t kstep16 = fin_kstep_bytes<kPair>(n_pad) >> 4;
    wait_d_free<kPair>(s, I.unit);
    I.stamp(0);
    const uint32_t d = tmem + kColD + (I.unit & 1u) * 128u;
    // k-steps [0, 8) need the first half of the operand, [8, k_steps) the second; the bias step closes the unit
    uint32_t st = 0;
    auto kstep = [&](uint32_t w, uint32_t acc) {
        mma_ts<kPair>(d, ah, mk_desc(w | (rows1 << 16)), idesc2, acc);
        mma_ss<kPair>(d, mk_desc(al), mk_desc((w + off2) | (rows2 << 16)), idesc, 1u);
        ah += 8u;
        al += 2u * (kChunkBytes >> 4);
    };
    wait_a_ready<kPair>(s, I, in, 0);
    I.stamp(1);
    if (per == 4 && k_steps == 16) {            // the shipped shape: 4 stages of 4 k-steps + the bias stage
#pragma unroll 1
        for (int g4 = 0; g4 < 4; ++g4) {
            if (g4 == 2) wait_a_ready<kPair>(s, I, in, 1);
            const uint32_t w = wait_stage<kPair>(s, I, st);
            kstep(w, g4 > 0 ? 1u : 0u);
            kstep(w + kstep16, 1u);
            kstep(w + 2u * kstep16, 1u);
            kstep(w + 3u * kstep16, 1u);
            commit<kPair>(&s.empty[st]);
            ++I.cnt;
        }
        const uint32_t w = wait_stage<kPair>(s, I, st);
        mma_ss<kPair>(d, on, mk_desc(w | (rows1 << 16)), idesc2, 1u);
        commit<kPair>(&s.empty[st]);
        ++I.cnt;
    } else {                                    // general shape: running position inside the stage
        uint32_t w = 0;
        int left = 0, done = 0;
        auto next_block = [&]() {
            if (left == 0) {
                if (done) {
                    commit<kPair>(&s.empty[st]);
                    ++I.cnt;
                }
                w = wait_stage<kPair>(s, I, st);
                left = per;
                ++done;
            } else {
                w += kstep16;
            }
            --left;
        };
        const int mid = min(8, k_steps);
        for (int kk = 0; kk < mid; ++kk) {
            next_block();
            kstep(w, kk > 0 ? 1u : 0u);
        }
        wait_a_ready<kPair>(s, I, in, 1);
        for (int kk = mid; kk < k_steps; ++kk) {
            next_block();
            kstep(w, 1u);
        }
        next_block();
        mma_ss<kPair>(d, on, mk_desc(w | (rows1 << 16)), idesc2, 1u);
        commit<kPair>(&s.empty[st]);
        ++I.cnt;
    }
    (void)steps;
    commit<kPair>(&s.d_full[I.unit & 1u]);
    I.stamp(2);
    ++I.unit;
    ++I.tu;
}

// ---- row threads -----------------------------------------------------------------------------------------
struct Row {
    int row, part, rt;
    uint32_t lane_base;
    uint32_t unit = 0;      // accumulator units consumed so far
    int cur = 0;            // mirrors Issuer::cur
    uint32_t l_a_ready = 0, l_d_free = 0, l_xyz = 0;   // CTA pairs: shared::cluster addresses of the leader's barriers
    long long* tr = nullptr;
    int tu = 0;
    __device__ __forceinline__ void stamp(int slot) const {
        if (tr && tu < kTraceUnits && threadIdx.x == 64) tr[tu * kTraceSlots + slot] = clock64();
    }
};

__device__ __forceinline__ uint32_t wait_d_full(const Smem& s, Row& r) {
    tc::mbar_wait(&s.d_full[r.unit & 1u], (r.unit >> 1) & 1u);
    tc::fence_after_sync();
    r.stamp(3);
    return r.lane_base + kColD + (r.unit & 1u) * 128u;
}

template <bool kPair>
__device__ __forceinline__ void release_d(const Smem& s, Row& r) {
    tc::tmem_wait_ld();
    tc::fence_before_sync();
    arrive_rows<kPair>(&s.d_free[r.unit & 1u], r.l_d_free + 8u * (r.unit & 1u));
    r.stamp(4);
    ++r.unit;
    ++r.tu;
}

// one accumulator unit (128 columns = k rows [128 h, 128 h + 128) of the next layer's operand): ReLU, rescale to the next
// layer's operand scale, fp16 split; hi -> tensor memory (buffer `out`), lo -> shared memory
template <bool kPair>
__device__ __forceinline__ void epilogue_half(const Smem& s, Row& r, int h, int out, float e, float* st_blk = nullptr,
                                              float inv = 0.0f) {
    const uint32_t d = wait_d_full(s, r);
    float v[3][16];
    const int n_chunks = r.part < 2 ? 3 : 2;              // chunks part, part + 3, part + 6 of the unit's eight
#pragma unroll
    for (int j = 0; j < 3; ++j)
        if (j < n_chunks) tc::tmem_ld16(d + (uint32_t)(16 * (r.part + 3 * j)), v[j]);
    release_d<kPair>(s, r);                                // the accumulator is free again as soon as it sits in registers
    const uint32_t ah = r.lane_base + kColAH + (uint32_t)out * 128u + (uint32_t)h * 64u;
    unsigned char* al = s.a_lo + (size_t)out * kALoBytes + (size_t)(16 * h) * kChunkBytes + (size_t)r.row * 16;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        if (j >= n_chunks) break;
        const int c = r.part + 3 * j;
        uint32_t hi[8], lo[8];
        if (st_blk) {      // training: the next layer's input block (256 rows) gets the unscaled activation
#pragma unroll
            for (int i = 0; i < 16; ++i) st_blk[stash_idx(256, 128 * h + 16 * c + i, r.row)] = fmaxf(v[j][i], 0.0f) * inv;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
            split2(fmaxf(v[j][2 * i], 0.0f) * e, fmaxf(v[j][2 * i + 1], 0.0f) * e, hi[i], lo[i]);
        tc::tmem_st8u(ah + (uint32_t)(8 * c), hi);
        *reinterpret_cast<uint4*>(al + (size_t)(2 * c) * kChunkBytes) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        *reinterpret_cast<uint4*>(al + (size_t)(2 * c + 1) * kChunkBytes) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
    }
    tc::tmem_wait_st();
    tc::fence_proxy_async_smem();
    tc::fence_before_sync();
    arrive_rows<kPair>(&s.a_ready[out * 2 + h], r.l_a_ready + 8u * (uint32_t)(out * 2 + h));
    --r.tu;
    r.stamp(5);
    ++r.tu;
}

// final layer, phase A: the stacked accumulator's two column blocks summed, un-scaled (x `mul`) -> scratch[row][c]
template <bool kPair>
__device__ __forceinline__ void final_to_scratch(const Smem& s, Row& r, int n_out, int n_pad, float mul) {
    const uint32_t d = wait_d_full(s, r);
    const int c0 = 16 * r.part;
    if (c0 < n_out) {
        float v[16], u[16];
        tc::tmem_ld16(d + (uint32_t)c0, v);
        tc::tmem_ld16(d + (uint32_t)(n_pad + c0), u);
        tc::tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (c0 + i < n_out) s.scratch[(size_t)r.row * kScrStride + c0 + i] = (v[i] + u[i]) * mul;
    }
    release_d<kPair>(s, r);
}

// rows [0, nch) of the scratch summed over each ray run and added into dst[ray * stride + col0 + c]: four adjacent lanes
// share one (run, channel) item, each sums a contiguous quarter of the run, a fixed-order shuffle tree combines them
__device__ __forceinline__ void reduce_runs(const Smem& s, int rt, int nch, float* __restrict__ dst, int stride, int col0) {
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
            for (int m = a; m < b; ++m) acc += s.scratch[(size_t)m * kScrStride + c];
            ray = s.ray[m0];
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        if (idx < items && sub == 0) atomicAdd(dst + (int64_t)ray * stride + col0 + c, acc);
    }
    tc::named_bar_sync(1, kRowThreads);
}

// final layer, phase B: (softmax, x compositing weight,) per-ray run sums
__device__ __forceinline__ void final_reduce(const Smem& s, const Row& r, const XParams& P, const XStack& st, float w,
                                             float* st_prob = nullptr) {
    if (st.is_sem) {
        tc::named_bar_sync(1, kRowThreads);            // scratch complete (two parts wrote it)
        if (r.part == 0) {
            const int n = P.n_cls;
            float v[kScratchRows];
#pragma unroll
            for (int i = 0; i < kScratchRows; ++i) v[i] = i < n ? s.scratch[(size_t)r.row * kScrStride + i] : 0.0f;
            if (P.softmax) {
                float mx = -INFINITY;
#pragma unroll
                for (int i = 0; i < kScratchRows; ++i)
                    if (i < n) mx = fmaxf(mx, v[i]);
                float tot = 0.0f;
#pragma unroll
                for (int i = 0; i < kScratchRows; ++i) {
                    v[i] = i < n ? expf(v[i] - mx) : 0.0f;
                    tot += v[i];
                }
                if (st_prob) {      // training: the softmax backward needs the probabilities
                    const float it = 1.0f / tot;
#pragma unroll
                    for (int i = 0; i < kScratchRows; ++i)
                        if (i < n) st_prob[stash_idx(n, i, r.row)] = v[i] * it;
                }
                const float sc = w / tot;
#pragma unroll
                for (int i = 0; i < kScratchRows; ++i) v[i] *= sc;
            } else {
#pragma unroll
                for (int i = 0; i < kScratchRows; ++i) v[i] *= w;
            }
#pragma unroll
            for (int i = 0; i < kScratchRows; ++i)
                if (i < n) s.scratch[(size_t)r.row * kScrStride + i] = v[i];
        }
        reduce_runs(s, r.rt, P.n_cls, P.sem_raw, P.n_cls, 0);
    } else {
        reduce_runs(s, r.rt, st.n_out, P.ins, P.ins_width, st.out_col0);
    }
}

// the tile's first-layer operand: (x, y, z) * ca in k rows 0..2 of chunk 0 (chunk 1 and the rest of chunk 0 stay zero)
template <bool kPair>
__device__ __forceinline__ void build_xyz(const Smem& s, const Row& r, const float4& p, float ca) {
    if (r.part == 0) {
        uint32_t h0, l0, h1, l1;
        split2(p.x * ca, p.y * ca, h0, l0);
        split2(p.z * ca, 0.0f, h1, l1);
        *reinterpret_cast<uint4*>(s.xyz + (size_t)r.row * 16) = make_uint4(h0, h1, 0u, 0u);
        *reinterpret_cast<uint4*>(s.xyz + 2 * kChunkBytes + (size_t)r.row * 16) = make_uint4(l0, l1, 0u, 0u);
    }
    tc::fence_proxy_async_smem();
    tc::fence_before_sync();
    arrive_rows<kPair>(s.xyz_ready, r.l_xyz);
}

// kPair: clusters of two CTAs; both run the same number of iterations (a CTA whose tile index is past the end carries an
// empty tile through the same schedule), the leader (cluster rank 0) issues every MMA for the pair
template <bool kPair, bool kStash>
__global__ void __launch_bounds__(kThreads, 1) heads_x16_kernel(const __grid_constant__ XParams P) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Smem s = carve_smem(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < P.n_gemms) s.sc[threadIdx.x] = make_float2(P.g[threadIdx.x].meta[0], P.g[threadIdx.x].meta[2]);
    for (int i = threadIdx.x; i < 2 * kChunkBytes / 2; i += kThreads)
        reinterpret_cast<__half*>(s.ones)[i] = __float2half_rn((i < kRows * 8 && (i & 7) == 0) ? 1.0f : 0.0f);
    for (int i = threadIdx.x; i < 4 * kChunkBytes / 4; i += kThreads) reinterpret_cast<uint32_t*>(s.xyz)[i] = 0u;
    tc::fence_proxy_async_smem();
    if (threadIdx.x == 0) {
        for (int i = 0; i < Cfg<kPair>::kStages; ++i) {
            tc::mbar_init(&s.full[i], 1);
            tc::mbar_init(&s.empty[i], 1);
            tc::mbar_init(&s.full_peer[i], 1);
        }
        for (int i = 0; i < 4; ++i) tc::mbar_init(&s.a_ready[i], Cfg<kPair>::kArrivals);
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&s.d_full[i], 1);
            tc::mbar_init(&s.d_free[i], Cfg<kPair>::kArrivals);
        }
        tc::mbar_init(s.xyz_ready, Cfg<kPair>::kArrivals);
        tc::fence_barrier_init();
    }
    if (warp == 1) {
        if (kPair)
            tc::tmem_alloc_pair(s.tmem_base, kTmemCols);
        else
            tc::tmem_alloc(s.tmem_base, kTmemCols);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (kPair) tc::cluster_sync();       // the leader's barriers exist before anyone arrives on them remotely
    tc::fence_after_sync();
    const uint32_t tmem = *s.tmem_base;
    const long long n_act = min((long long)P.stats[0], P.cap);
    const int n_tiles = (int)((n_act + kRows - 1) / kRows);
    const uint32_t rank = kPair ? tc::cluster_ctarank() : 0u;
    const int tile_first = kPair ? (int)(blockIdx.x & ~1u) : (int)blockIdx.x;   // iteration key (pairs: the leader's tile)
    const int tile_step = (int)gridDim.x;

    if (warp == 0) {
        if (tc::elect_one()) {
            Ring<kPair> ring;
            for (int key = tile_first; key < n_tiles; key += tile_step)
                for (int gi = 0; gi < P.n_gemms; ++gi) produce<kPair>(s, P.g[gi], ring, rank);
        }
    } else if (warp == 1) {
        if (tc::elect_one()) {
            if (kPair && rank != 0) {
                Ring<true> ring;
                const uint32_t leader_full_peer = tc::map_to_rank(&s.full_peer[0], 0);
                for (int key = tile_first; key < n_tiles; key += tile_step)
                    for (int gi = 0; gi < P.n_gemms; ++gi) relay(s, gemm_stages<true>(P.g[gi]), ring, leader_full_peer);
            } else {
                Issuer<kPair> I;
                I.w16 = tc::smem_addr(s.w) >> 4;
                uint32_t tl = 0;
                for (int key = tile_first; key < n_tiles; key += tile_step, ++tl) {
                    I.tr = P.trace && blockIdx.x == 0 && tl < 4 ? P.trace + (size_t)tl * kTraceUnits * kTraceSlots : nullptr;
                    I.tu = 0;
                    wait_rows<kPair>(s.xyz_ready, tl & 1u);
                    for (int gi = 0; gi < P.n_gemms; ++gi) {
                        const XGemm& g = P.g[gi];
                        if (g.kind == kL0)
                            issue_l0<kPair>(s, I, tmem);
                        else if (g.kind == kHid)
                            issue_hidden<kPair>(s, I, tmem, g.k_steps);
                        else
                            issue_final<kPair>(s, I, tmem, g);
                    }
                }
            }
        }
    } else {
        Row r;
        {
            const int quarter = warp & 3;                  // the TMEM lane quarter a warp may touch is warp_id % 4
            r.row = quarter * 32 + lane;
            r.part = (warp - 2) >> 2;
            r.rt = threadIdx.x - 64;
            r.lane_base = tmem + ((uint32_t)(quarter * 32) << 16);
            if (kPair) {
                r.l_a_ready = tc::map_to_rank(&s.a_ready[0], 0);
                r.l_d_free = tc::map_to_rank(&s.d_free[0], 0);
                r.l_xyz = tc::map_to_rank(s.xyz_ready, 0);
            }
        }
        const float ca0 = s.sc[0].x;                       // operand scale of the shared xyz operand (first stack's L0)
        auto fetch = [&](int tile, bool valid, float4& p, int& ray) {
            p = make_float4(0.f, 0.f, 0.f, 0.f);
            ray = -1;
            const long long rec = (long long)tile * kRows + r.row;
            if (valid && tile < n_tiles && rec < n_act) {
                p = P.rec_pos[rec];
                ray = P.rec_ray[rec];
            }
        };
        float4 p, p_next;
        int ray, ray_next;
        fetch(tile_first + (int)rank, tile_first < n_tiles, p_next, ray_next);
        if (tile_first < n_tiles) build_xyz<kPair>(s, r, p_next, ca0);
        uint32_t tl = 0;
        for (int key = tile_first; key < n_tiles; key += tile_step, ++tl) {
            const int tile = key + (int)rank;
            r.tr = P.trace && blockIdx.x == 0 && tl < 4 ? P.trace + (size_t)tl * kTraceUnits * kTraceSlots : nullptr;
            r.tu = 0;
            p = p_next;
            ray = ray_next;
            const bool more = key + tile_step < n_tiles;
            fetch(tile + tile_step, more, p_next, ray_next);
            const int nv = (int)max(0ll, min((long long)kRows, n_act - (long long)tile * kRows));
            if (r.part == 0) s.ray[r.row] = ray;
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
            int pending = -1;                               // stack whose final layer waits in the scratch for its reduction
            float* stash = nullptr;                         // training: this tile's A-stash
            if (kStash) stash = P.stash_a + (size_t)tile * P.lay.a_rows * kRows;
            float* st_prob = kStash ? stash + (size_t)P.lay.prob_off * kRows : nullptr;
            for (int si = 0; si < P.n_stacks; ++si) {
                const XStack& st = P.st[si];
                if (kStash && r.part == 0) {                // layer-0 input of the stack (pe = 0): 16 rows, 3 valid
                    float* b = stash + (size_t)P.lay.a_off[st.stash_id][0] * kRows;
                    b[stash_idx(16, 0, r.row)] = p.x;
                    b[stash_idx(16, 1, r.row)] = p.y;
                    b[stash_idx(16, 2, r.row)] = p.z;
                    for (int k = 3; k < 16; ++k) b[stash_idx(16, k, r.row)] = 0.0f;
                }
                for (int l = 0; l + 1 < st.n_layers; ++l) {
                    const int gi = st.g_first + l;
                    // un-scale of this accumulator, re-scale to the next layer's operand (all powers of two)
                    const float inv = l == 0 ? s.sc[gi].y * (s.sc[gi].x / ca0) : s.sc[gi].y;
                    const float e = s.sc[gi + 1].x * inv;
                    const int out = r.cur ^ 1;
                    float* blk = kStash ? stash + (size_t)P.lay.a_off[st.stash_id][l + 1] * kRows : nullptr;
                    epilogue_half<kPair>(s, r, 0, out, e, blk, inv);
                    epilogue_half<kPair>(s, r, 1, out, e, blk, inv);
                    r.cur = out;
                    if (l == 0 && pending >= 0) {           // under the next stack's first hidden layer
                        final_reduce(s, r, P, P.st[pending], p.w, st_prob);
                        pending = -1;
                    }
                }
                const int gf = st.g_first + st.n_layers - 1;
                final_to_scratch<kPair>(s, r, st.n_out, P.g[gf].n_pad, st.is_sem ? s.sc[gf].y : s.sc[gf].y * p.w);
                pending = si;
                if (si + 1 == P.n_stacks) {
                    if (more) build_xyz<kPair>(s, r, p_next, ca0);   // the next tile's first layers start now
                    final_reduce(s, r, P, st, p.w, st_prob);
                    pending = -1;
                }
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

// w_tc16 slabs of an N = 256 layer ([hi | lo][2 k-chunks][256][8] per k-step) -> x16 stages.
//   single CTA: for each N-half h, for each k-step (the bias step last): [hi: 2 k-chunks x 128 rows x 16 B | lo: same]
//   CTA pairs : for each cluster rank r, N-half h, k-step: [hi: 2 k-chunks x 64 rows | lo]: rows [64 r, 64 r + 64) of the half
__global__ void __launch_bounds__(256) pack_x16_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int steps) {
    const int64_t total = (int64_t)steps * 1024;          // 16-byte units per layer: steps x 16 KB
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int ks = (int)(i >> 10), rem = (int)(i & 1023);
        const int part = rem >> 9, chunk = (rem >> 8) & 1, row = rem & 255;      // source order
        const int h = row >> 7, r = row & 127;
        const uint4 v = src[i];
        dst[((int64_t)(h * steps + ks) << 9) + (part << 8) + (chunk << 7) + r] = v;
        const int rank = r >> 6, rr = r & 63;
        dst[total + (int64_t)rank * (2 * steps * 256) + ((int64_t)(h * steps + ks) << 8) + (part << 7) + (chunk << 6) + rr] = v;
    }
}

// clift_pack_linear_x16_batch: block b finds its job by binary search over the jobs' first-block prefix
__global__ void __launch_bounds__(256) pack_x16_batch_kernel(const clift_x16_job* __restrict__ jobs, int n_jobs) {
    int lo = 0, hi = n_jobs - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (jobs[mid].first_block <= (int)blockIdx.x)
            lo = mid;
        else
            hi = mid - 1;
    }
    const clift_x16_job J = jobs[lo];
    const int64_t total = (int64_t)J.steps * 1024;
    const int64_t i = (int64_t)((int)blockIdx.x - J.first_block) * 256 + threadIdx.x;
    if (i >= total) return;
    const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(J.w_tc16) + kHeaderFloats);
    uint4* dst = reinterpret_cast<uint4*>(J.dst);
    const int ks = (int)(i >> 10), rem = (int)(i & 1023);
    const int part = rem >> 9, chunk = (rem >> 8) & 1, row = rem & 255;
    const int h = row >> 7, r = row & 127;
    const uint4 v = src[i];
    dst[((int64_t)(h * J.steps + ks) << 9) + (part << 8) + (chunk << 7) + r] = v;
    const int rank = r >> 6, rr = r & 63;
    dst[total + (int64_t)rank * (2 * J.steps * 256) + ((int64_t)(h * J.steps + ks) << 8) + (part << 7) + (chunk << 6) + rr] = v;
}

}  // namespace

static bool x16_stack_ok(const clift_mlp& m, int n_out) {
    if (m.n_layers < 3 || m.n_layers > CLIFT_MAX_LAYERS || m.dims[0] != 3 || n_out > kScratchRows) return false;
    for (int l = 0; l < m.n_layers; ++l)
        if (!m.w_tc16[l]) return false;
    for (int l = 0; l + 1 < m.n_layers; ++l)
        if (m.dims[l + 1] != 256 || !m.w_x16[l]) return false;
    return m.dims[m.n_layers] == n_out;
}

bool heads_x16_available(const clift_field* f, int heads) {
    if (heads & CLIFT_HEAD_RGB) return false;             // the caller splits the rgb stack off first
    if (!(heads & (CLIFT_HEAD_SEMANTIC | CLIFT_HEAD_INSTANCE))) return false;
    if ((heads & CLIFT_HEAD_SEMANTIC) && (f->pe_sem != 0 || f->semantic_grid.comps || !x16_stack_ok(f->semantic, f->num_classes)))
        return false;
    if (heads & CLIFT_HEAD_INSTANCE) {
        if (f->pe_ins != 0 || f->instance_grid.comps || !x16_stack_ok(f->instance_fast, f->dim_instance)) return false;
        if (f->slow_fast && !x16_stack_ok(f->instance_slow, f->dim_instance)) return false;
    }
    const char* e = getenv("CLIFT_X16");                 // development switch: CLIFT_X16=0 keeps the serial kernel
    return !(e && atoi(e) == 0);
}

int launch_heads_forward_x16(const clift_render_cfg* cfg, const clift_field* field, const Workspace& ws, int64_t cap,
                             int64_t n_rays, float* sem_raw, float* ins, cudaStream_t stream, const StashLayout* lay) {
    XParams P;
    memset(&P, 0, sizeof(P));
    P.rec_pos = ws.rec_pos;
    P.rec_ray = ws.rec_ray;
    P.stats = reinterpret_cast<const unsigned long long*>(ws.stats);
    P.cap = cap;
    P.n_cls = field->num_classes;
    P.softmax = cfg->semantic_softmax;
    P.ins_width = field->dim_instance * (field->slow_fast ? 2 : 1);
    P.sem_raw = sem_raw;
    P.ins = ins;
    P.trace = get_tc_trace();
    if (lay) {      // training forward: record the stash
        P.stash_a = ws.stash_a;
        P.lay = *lay;
    }
    auto add = [&](const clift_mlp& m, int is_sem, int col0, int stash_id) {
        XStack& st = P.st[P.n_stacks++];
        st.stash_id = stash_id;
        st.n_layers = m.n_layers;
        st.g_first = P.n_gemms;
        st.n_out = m.dims[m.n_layers];
        st.is_sem = is_sem;
        st.out_col0 = col0;
        for (int l = 0; l < m.n_layers; ++l) {
            XGemm& g = P.g[P.n_gemms++];
            const bool fin = l + 1 == m.n_layers;
            g.meta = reinterpret_cast<const float*>(m.w_tc16[l]);
            g.kind = fin ? kFin : (l == 0 ? kL0 : kHid);
            g.n_pad = (int)round_up(m.dims[l + 1], 32);
            g.k_steps = (int)ceil_div(m.dims[l], 16);
            const size_t steps = (size_t)g.k_steps + 1;
            g.w = fin ? reinterpret_cast<const unsigned char*>(g.meta + kHeaderFloats) : reinterpret_cast<const unsigned char*>(m.w_x16[l]);
            g.w_pair = g.w + (fin ? steps * 64 * g.n_pad : steps * 16384);
        }
    };
    if (sem_raw) add(field->semantic, 1, 0, 0);
    if (ins) {
        add(field->instance_fast, 0, 0, 1);
        if (field->slow_fast) add(field->instance_slow, 0, field->dim_instance, 2);
    }
    if (P.n_stacks == 0 || n_rays <= 0) return CLIFT_OK;
    // CTA pairs by default (each SM stages half of every weight stage); CLIFT_X16_PAIR=0: single CTAs
    const char* e_pair = getenv("CLIFT_X16_PAIR");
    const bool pair = !(e_pair && atoi(e_pair) == 0) && sm_count() >= 2;
    auto launch = [&](auto kernel, int grid, int cluster) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
        if (e != cudaSuccess) return e;
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3((unsigned)grid);
        lc.blockDim = dim3(kThreads);
        lc.dynamicSmemBytes = kSmemBytes;
        lc.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)cluster;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        lc.attrs = attr;
        lc.numAttrs = 1;
        return cudaLaunchKernelEx(&lc, kernel, P);
    };
    if (pair && lay)
        CLIFT_CUDA(launch(heads_x16_kernel<true, true>, sm_count() & ~1, 2));
    else if (pair)
        CLIFT_CUDA(launch(heads_x16_kernel<true, false>, sm_count() & ~1, 2));
    else if (lay)
        CLIFT_CUDA(launch(heads_x16_kernel<false, true>, sm_count(), 1));
    else
        CLIFT_CUDA(launch(heads_x16_kernel<false, false>, sm_count(), 1));
    CLIFT_AFTER_LAUNCH("heads_x16_kernel");
    return CLIFT_OK;
}

}  // namespace clift

using namespace clift;

extern "C" int64_t clift_x16_weight_bytes(int32_t n_out, int32_t n_in, int32_t has_bias) {
    if (n_out != 256 || n_in <= 0 || n_in > 256) return CLIFT_ERR_UNSUPPORTED;
    return ((int64_t)ceil_div(n_in, 16) + (has_bias ? 1 : 0)) * 16384 * 2;     // single-CTA stages + CTA-pair stages
}

extern "C" int32_t clift_pack_linear_x16_batch(const clift_x16_job* jobs, int32_t n_jobs, int32_t total_blocks, void* stream) {
    CLIFT_CHECK_ARG(n_jobs >= 0 && total_blocks >= 0 && (n_jobs == 0 || jobs), "null table or negative size");
    if (n_jobs == 0 || total_blocks == 0) return CLIFT_OK;
    pack_x16_batch_kernel<<<(unsigned)total_blocks, 256, 0, (cudaStream_t)stream>>>(jobs, n_jobs);
    CLIFT_AFTER_LAUNCH("pack_x16_batch_kernel");
    return CLIFT_OK;
}

extern "C" int32_t clift_pack_linear_x16(const void* w_tc16, void* dst, int32_t n_out, int32_t n_in, int32_t has_bias, void* stream) {
    CLIFT_CHECK_ARG(w_tc16 && dst, "null pointer");
    CLIFT_CHECK_SUPPORTED(n_out == 256 && n_in > 0 && n_in <= 256, "x16 stages exist for 256-wide layers only");
    CLIFT_CHECK_ARG((((uintptr_t)w_tc16 | (uintptr_t)dst) & 15) == 0, "operands must be 16-byte aligned");
    const int steps = (int)ceil_div(n_in, 16) + (has_bias ? 1 : 0);
    const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(w_tc16) + kHeaderFloats);
    pack_x16_kernel<<<(unsigned)std::min<int64_t>(ceil_div((int64_t)steps * 1024, 256), 4 * sm_count()), 256, 0, (cudaStream_t)stream>>>(
        src, reinterpret_cast<uint4*>(dst), steps);
    CLIFT_AFTER_LAUNCH("pack_x16_kernel");
    return CLIFT_OK;
}
