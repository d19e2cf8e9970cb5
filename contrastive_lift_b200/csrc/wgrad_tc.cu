// Tensor-core weight gradients of the MLP heads (training backward): for every Linear of the evaluated heads
//     dW^T[k][n] += sum over active records  A[k][m] * dZ[n][m]
// with A = the layer's input (A-stash) and dZ = dL/d(pre-activation) (Z-stash), both written per 128-record tile by the
// head forward / backward kernels (heads.cu) in 16-record groups of [rows][16 floats] (stash_idx in common.cuh).  The
// record index m is the REDUCTION dimension, so both stashes already are "K-major" UMMA operands: 16 bytes = 4 consecutive
// records of one row = one core-matrix row, and one 16-record group of all rows is one contiguous piece of memory.
//
// One launch covers all layers.  A CTA owns (layer, split): it walks tiles split, split + n_splits, ... and accumulates the
// whole K x N gradient block in tensor memory (ceil(K/128) accumulators of 128 lanes x N fp32 columns, <= 512 columns),
// then adds it to the packed gradient with vector reductions - one flush per CTA instead of one per tile.
//   warps 1..8 (256 loader threads): per ring stage (one 16-record group, two for small layers), 16-byte cp.async copies of the A and dZ row pieces straight into
//       the canonical no-swizzle K-major layout [record chunk of 4][row][4 floats], then the 3xTF32 split in place
//       (hi = x with the low 13 mantissa bits cleared, lo = x - hi); a ring of 3 (256 x 256 layer) to 8 (small layers)
//       stages with full/empty mbarriers keeps ring-depth - 1 stages of loads in flight
//   warp 0, one lane: per 16-record group 2 k-steps x ceil(K/128) row blocks x 3 products (A_hi*Z_hi, A_hi*Z_lo, A_lo*Z_hi) of
//       tcgen05.mma kind::tf32 (M = 128, N = n_pad, K = 8), fp32 accumulation; tcgen05.commit releases the stage
// fp32-faithful like the 3xTF32 forward heads: products carry 21+ significant bits per operand.
#include "launchers.h"
#include "tcgen05.cuh"

namespace clift {

long long* get_tc_trace();   // heads_tc.cu (clift_debug_tc_trace): here 4 int64 per CTA = layer, stages, cycles to the last MMA, total

namespace {

constexpr int kTile = CLIFT_TILE;
constexpr int kWtChunkM = 16;                    // records per stash group = two kind::tf32 k-steps
constexpr int kWtLoaders = 256;
constexpr int kWtThreads = 32 + kWtLoaders;
constexpr int kWtMaxStages = 8;
constexpr int kWtRingBytes = 224 * 1024;         // stages of 2 * (K + N) * 64 bytes: 3 for a 256 x 256 layer, up to 8 for small ones
constexpr size_t kWtSmemBytes = (size_t)kWtRingBytes + 256;
constexpr uint32_t kWtTmemCols = 512;
constexpr int kWtMaxF4 = (256 + 256) * (kWtChunkM / 4) / kWtLoaders;   // 16-byte pieces per loader thread and stage (8)

struct WtLayer {
    const float* a;      // A-stash rows of this layer in tile 0
    const float* z;      // Z-stash rows of this layer in tile 0
    float* out;          // packed dW^T [K][N]
    int K, N;            // k_pad(in) (multiple of 16, <= 256), n_pad(out) (64, 128 or 256)
    int first_cta, n_splits;
    int groups;          // 16-record groups per ring stage (1 or 2): small layers move twice the records per handoff
};

struct WgradTcParams {
    WtLayer layer[kWgradTcMaxLayers];
    int n_layers;
    long long a_tile_stride, z_tile_stride;   // floats
    const unsigned long long* stats;
    long long cap;
    long long* trace;
};

__device__ __forceinline__ void red_add4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// cp.async.wait_group with a run-time count (the ring depth depends on the layer)
__device__ __forceinline__ void cp_async_wait_dyn(int n) {
    switch (n) {
        case 0: cp_async_wait<0>(); break;
        case 1: cp_async_wait<1>(); break;
        case 2: cp_async_wait<2>(); break;
        case 3: cp_async_wait<3>(); break;
        case 4: cp_async_wait<4>(); break;
        case 5: cp_async_wait<5>(); break;
        default: cp_async_wait<6>(); break;
    }
}

// Stage layout (part = groups * (K + N) * 64 bytes): [A_hi: 4*groups chunks x K rows x 16 B][Z_hi: 4*groups x N x 16 B][A_lo][Z_lo].
// The loaders cp.async the RAW fp32 pieces into the hi half (no register staging: register-destination loads of two stages
// end up sharing hardware scoreboards and serialise), then every thread splits exactly the pieces it copied itself, in
// place (hi) and into the lo half, and publishes the stage to the tensor core.
__global__ void __launch_bounds__(kWtThreads, 1) wgrad_tc_kernel(const __grid_constant__ WgradTcParams P) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + kWtRingBytes);
    uint64_t* empty = full + kWtMaxStages;
    uint64_t* done = empty + kWtMaxStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long t_start = clock64();
    long long t_mma = t_start;

    int li = 0;
    while (li + 1 < P.n_layers && (int)blockIdx.x >= P.layer[li + 1].first_cta) ++li;
    const WtLayer& Lr = P.layer[li];
    const int K = Lr.K, N = Lr.N;
    const int split = (int)blockIdx.x - Lr.first_cta, n_splits = Lr.n_splits;
    const long long n_act = min((long long)P.stats[0], P.cap);
    const long long n_tiles = (n_act + kTile - 1) / kTile;
    const long long my_tiles = split < n_tiles ? (n_tiles - split + n_splits - 1) / n_splits : 0;
    const int groups = Lr.groups;
    const int iters_per_tile = kTile / (kWtChunkM * groups);
    const long long n_iter = my_tiles * iters_per_tile;
    const int n_blk = (K + 127) >> 7;
    const uint32_t part_bytes = (uint32_t)(K + N) * 64u * (uint32_t)groups;
    const uint32_t stage_bytes = 2u * part_bytes;
    const int n_stages = min(kWtMaxStages, (int)(kWtRingBytes / stage_bytes));

    if (threadIdx.x == 0) {
        for (int i = 0; i < kWtMaxStages; ++i) {
            tc::mbar_init(&full[i], kWtLoaders / 32);     // one arrival per loader warp
            tc::mbar_init(&empty[i], 1);
        }
        tc::mbar_init(done, 1);
        tc::fence_barrier_init();
    }
    if (warp == 0) tc::tmem_alloc(tmem_slot, kWtTmemCols);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (n_iter > 0 && tc::elect_one()) {
            const uint32_t idesc = tc::make_idesc_tf32(128, N);
            const uint32_t lbo_a = (uint32_t)K * 16u, lbo_z = (uint32_t)N * 16u;
            const uint32_t z_base = (uint32_t)K * 64u * (uint32_t)groups;
            const int k_steps = 2 * groups;
            int stage = 0;
            uint32_t phase = 0;
            for (long long it = 0; it < n_iter; ++it) {
                tc::mbar_wait(&full[stage], phase);
                tc::fence_after_sync();
                const uint32_t base = tc::smem_addr(smem_raw) + (uint32_t)stage * stage_bytes;
                for (int ks = 0; ks < k_steps; ++ks) {
                    const uint64_t z_hi = tc::make_smem_desc(base + z_base + (uint32_t)ks * 2u * lbo_z, lbo_z, 128u);
                    const uint64_t z_lo = tc::make_smem_desc(base + part_bytes + z_base + (uint32_t)ks * 2u * lbo_z, lbo_z, 128u);
                    for (int b = 0; b < n_blk; ++b) {
                        const uint32_t a_off = (uint32_t)ks * 2u * lbo_a + (uint32_t)b * 128u * 16u;
                        const uint64_t a_hi = tc::make_smem_desc(base + a_off, lbo_a, 128u);
                        const uint64_t a_lo = tc::make_smem_desc(base + part_bytes + a_off, lbo_a, 128u);
                        const uint32_t d = tmem + (uint32_t)(b * N);
                        tc::mma_ss(d, a_hi, z_hi, idesc, (it == 0 && ks == 0) ? 0u : 1u);
                        tc::mma_ss(d, a_hi, z_lo, idesc, 1u);
                        tc::mma_ss(d, a_lo, z_hi, idesc, 1u);
                    }
                }
                tc::mma_commit(&empty[stage]);
                if (++stage == n_stages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
            tc::mma_commit(done);
            tc::mbar_wait(done, 0);
            t_mma = clock64();
            if (P.trace && blockIdx.x < 240) P.trace[blockIdx.x * 4 + 2] = t_mma - t_start;
        }
    } else if (n_iter > 0) {
        const int lt = threadIdx.x - 32;
        const int n_f4 = (K + N) * (kWtChunkM / 4) * groups;
        const int depth = n_stages - 1;            // stages of loads in flight (the remaining one is under the tensor core)
        // Pieces of a stage: A part first (groups x K rows x 4 chunks), then the Z part.  Within a 16-record group a warp-wide
        // access covers 16 rows x 2 chunks: lanes 0-15 -> chunk c of rows r..r+15, lanes 16-31 -> chunk c+1: every half-warp
        // of a 16-byte shared access touches 256 contiguous bytes (two core matrices, conflict-free) and the warp as a whole
        // reads one 32-byte sector of each of 16 consecutive rows of the group's [rows][16] block.  off = byte offset inside the stage's hi half, src_off = float
        // offset from the stage's first group in the A block (or, tagged by bit 31, the Z block).
        uint32_t off[kWtMaxF4];
        uint32_t src_off[kWtMaxF4];
#pragma unroll
        for (int j = 0; j < kWtMaxF4; ++j) {
            int idx = lt + j * kWtLoaders;
            const bool is_z = idx >= K * 4 * groups;
            const int R = is_z ? N : K;
            if (is_z) idx -= K * 4 * groups;
            const int g = idx / (R * 4), rem = idx - g * (R * 4);
            const int wi = rem >> 5;                                    // warp-wide access: 16 rows x 2 chunks
            const int row = (wi >> 1) * 16 + (rem & 15), c = (wi & 1) * 2 + ((rem >> 4) & 1);
            off[j] = (is_z ? (uint32_t)K * 64u * (uint32_t)groups : 0u) + (uint32_t)((g * 4 + c) * R + row) * 16u;
            src_off[j] = (is_z ? 0x80000000u : 0u) | (uint32_t)(g * R * 16 + row * 16 + c * 4);
        }
        auto issue = [&](long long it) {           // cp.async the raw pieces of iteration `it` (nothing when past the end)
            if (it < n_iter) {
                const int stage = (int)(it % n_stages);
                const uint32_t phase = (uint32_t)((it / n_stages) & 1);
                tc::mbar_wait(&empty[stage], phase ^ 1u);
                const long long tile = split + (it / iters_per_tile) * n_splits;
                const int grp = (int)(it % iters_per_tile) * groups;      // first 16-record group of the stage
                const float* ga = Lr.a + tile * P.a_tile_stride + (size_t)grp * K * kWtChunkM;
                const float* gz = Lr.z + tile * P.z_tile_stride + (size_t)grp * N * kWtChunkM;
                unsigned char* base = smem_raw + (size_t)stage * stage_bytes;
#pragma unroll
                for (int j = 0; j < kWtMaxF4; ++j)
                    if (lt + j * kWtLoaders < n_f4)
                        cp_async16(base + off[j], (src_off[j] & 0x80000000u) ? gz + (src_off[j] & 0x7fffffffu) : ga + src_off[j]);
            }
            cp_async_commit();                       // one group per iteration, empty past the end: wait counts stay uniform
        };
        for (int d = 0; d < depth; ++d) issue(d);
        for (long long it = 0; it < n_iter; ++it) {
            cp_async_wait_dyn(depth - 1);            // this thread's pieces of iteration `it` have landed
            const int stage = (int)(it % n_stages);
            unsigned char* base = smem_raw + (size_t)stage * stage_bytes;
#pragma unroll
            for (int j = 0; j < kWtMaxF4; ++j) {
                if (lt + j * kWtLoaders < n_f4) {
                    float4* ph = reinterpret_cast<float4*>(base + off[j]);
                    const float4 x = *ph;
                    float4 hi, lo;
                    hi.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
                    hi.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
                    hi.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
                    hi.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
                    lo.x = x.x - hi.x;
                    lo.y = x.y - hi.y;
                    lo.z = x.z - hi.z;
                    lo.w = x.w - hi.w;
                    *ph = hi;
                    *reinterpret_cast<float4*>(base + part_bytes + off[j]) = lo;
                }
            }
            tc::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&full[stage]);
            issue(it + depth);
        }
        // flush: this warp's TMEM lane quarter, half of the columns
        tc::mbar_wait(done, 0);
        tc::fence_after_sync();
        const int quarter = warp & 3, half = (warp - 1) >> 2;
        const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16);
        for (int b = 0; b < n_blk; ++b) {
            const int k = b * 128 + quarter * 32 + lane;
            for (int c0 = half * (N / 2); c0 < (half + 1) * (N / 2); c0 += 16) {
                float v[16];
                tc::tmem_ld16(lane_base + (uint32_t)(b * N + c0), v);
                tc::tmem_wait_ld();
                if (k < K) {
                    float* dst = Lr.out + (size_t)k * N + c0;
#pragma unroll
                    for (int i = 0; i < 16; i += 4) red_add4(dst + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
                }
            }
        }
        tc::fence_before_sync();
    }
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, kWtTmemCols);
    if (P.trace && threadIdx.x == 32 && blockIdx.x < 240) {
        P.trace[blockIdx.x * 4 + 0] = li;
        P.trace[blockIdx.x * 4 + 1] = n_iter;
        P.trace[blockIdx.x * 4 + 3] = clock64() - t_start;
    }
}

}  // namespace

bool wgrad_tc_eligible(int K, int N) { return K >= 64 && K <= 256 && (N == 64 || N == 128 || N == 256); }

// `items`: the layers to cover (K, N, stash offsets in rows, output); one launch.
int launch_wgrad_tc(const Workspace& ws, const StashLayout& lay, const WgradTcItem* items, int n_items, int64_t cap,
                    cudaStream_t stream) {
    if (n_items <= 0) return CLIFT_OK;
    if (n_items > kWgradTcMaxLayers) {
        set_error("launch_wgrad_tc: %d layers > %d", n_items, kWgradTcMaxLayers);
        return CLIFT_ERR_UNSUPPORTED;
    }
    WgradTcParams P;
    memset(&P, 0, sizeof(P));
    P.n_layers = n_items;
    P.a_tile_stride = (long long)lay.a_rows * kTile;
    P.z_tile_stride = (long long)lay.z_rows * kTile;
    P.stats = reinterpret_cast<const unsigned long long*>(ws.stats);
    P.cap = cap;
    P.trace = get_tc_trace();
    // CTAs in proportion to the layer's time per 16 records; one wave of the SMs in total
    auto groups_of = [](const WgradTcItem& it) { return it.K + it.N <= 256 ? 2 : 1; };
    // measured (clock64 trace per CTA, scripts/wgrad_trace.py): a ring stage costs ~1200 cycles of handoff latency plus
    // ~3.8 cycles per 64-byte row piece it streams, for every layer shape - the loads, not the MMAs, set the pace
    auto weight = [&](const WgradTcItem& it) { return 1200ll / groups_of(it) + (long long)(3.8 * (it.K + it.N)); };
    long long total_w = 0;
    for (int i = 0; i < n_items; ++i) total_w += weight(items[i]);
    const int budget = std::max(sm_count(), n_items);
    int next = 0;
    for (int i = 0; i < n_items; ++i) {
        const long long w = weight(items[i]);
        int splits = (int)std::max(1ll, (long long)budget * w / total_w);
        WtLayer& L = P.layer[i];
        L.a = ws.stash_a + (size_t)items[i].a_row * kTile;
        L.z = ws.stash_z + (size_t)items[i].z_row * kTile;
        L.out = items[i].out;
        L.K = items[i].K;
        L.N = items[i].N;
        L.first_cta = next;
        L.n_splits = splits;
        L.groups = groups_of(items[i]);
        next += splits;
    }
    CLIFT_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWtSmemBytes));
    wgrad_tc_kernel<<<next, kWtThreads, kWtSmemBytes, stream>>>(P);
    CLIFT_AFTER_LAUNCH("wgrad_tc_kernel");
    return CLIFT_OK;
}

}  // namespace clift
