// Persistent, warp-specialised fp16-operand / fp32-accumulate GEMM main loop for sm_100a:
//   D[m][n] = sum_k A[m][k] * B[n][k]       (A: [M,K] row-major, B: [N,K] row-major = nn.Linear weight)
// One CTA per SM, 320 threads:
//   warp 0      TMA producer   (one lane issues cp.async.bulk.tensor for A and B k-blocks)
//   warp 1      TMEM allocator + MMA issuer (one lane issues tcgen05.mma, commits to mbarriers)
//   warps 2..5  epilogue group 0, warps 6..9 epilogue group 1
//               (tcgen05.ld accumulator -> registers -> Epi functor; a warp's TMEM lane quarter is warp%4)
// Pipelines: smem ring full/empty (TMA <-> MMA), two TMEM accumulator stages full/empty
// (MMA <-> epilogue), static persistent tile schedule (tile += gridDim.x).
//
// A tile = 128 rows x (SUBTILES sub-blocks of BLOCK_N accumulator columns); sub-blocks of one tile are
// processed back to back by the same CTA.  The two epilogue groups share the work in one of two ways:
//   Epi::kSplitColumns = true   both groups drain every accumulator stage, each its half of the columns
//                               (plain epilogues: halves the per-tile drain latency)
//   Epi::kSplitColumns = false  group g owns accumulator stage g, i.e. every other (tile, sub) unit, and
//                               processes all its columns (stateful epilogues: the MDN head's logsumexp);
//                               with SUBTILES == 2 the two groups hold the two halves of one tile's
//                               state and Epi::merge() combines them through shared memory.
#pragma once
#include "ptx.cuh"

namespace vitad {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 fp16 = 128 bytes = one swizzle row
constexpr int kGemmThreads = 320;
constexpr int kSmemBudget = 227 * 1024;

template <int BLOCK_N>
struct GemmSmem {
    static constexpr int kABytes = kBlockM * kBlockK * 2;
    static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kBarrierBytes = 512;
    static constexpr int kScratchBytes = 2 * kBlockM * 8;  // Epi::merge scratch: float2 per row, two parities
    static constexpr int kTailBytes = kBarrierBytes + kScratchBytes;
    static constexpr int kStages = (kSmemBudget - 1024 - kTailBytes) / kStageBytes > 8
                                       ? 8
                                       : (kSmemBudget - 1024 - kTailBytes) / kStageBytes;
    static constexpr int kTotalBytes = kStages * kStageBytes + kTailBytes + 1024;
    static_assert(kStages >= 2, "tile too large for shared memory");
    static_assert(kBBytes % 1024 == 0, "B stage must keep 1024-byte alignment (BLOCK_N % 8 == 0)");
};

// Optional in-kernel timeline (diagnostic builds only: `make TL=1` -> lib/libvitad_tl.so): clock64 stamps of the
// role loops of every CTA into a [grid][64] buffer, read back by tools/gpu_timeline.py.
#ifdef VITAD_TIMELINE
static __device__ unsigned long long* g_timeline = nullptr;
__device__ __forceinline__ void tl_stamp(int slot, bool global = false) {
    if (g_timeline != nullptr && slot < 64) {
        unsigned long long t;
        if (global)
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        else
            t = static_cast<unsigned long long>(clock64());
        g_timeline[(static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * 64 + slot] = t;
    }
}
#define VITAD_TL(slot) ::vitad::tl_stamp(slot)
#define VITAD_TLG(slot) ::vitad::tl_stamp(slot, true)
#else
#define VITAD_TL(slot)
#define VITAD_TLG(slot)
#endif

__host__ __device__ constexpr uint32_t tmem_cols_pow2(int n) {
    return n <= 32 ? 32u : n <= 64 ? 64u : n <= 128 ? 128u : n <= 256 ? 256u : 512u;
}

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// Epi interface (all __device__):
//   static constexpr bool kSplitColumns
//   void tile_begin(int m_blk, int n_tile, int row)                       once per tile per thread
//   void sub(int sub, int m_blk, int n_tile, int row, uint32_t taddr, int c0, int c1)
//        drain accumulator columns [c0, c1) of this stage (taddr has the lane base folded in); must have
//        completed every tcgen05.ld before returning
//   void merge(int group, float2* scratch_row)   only for !kSplitColumns && SUBTILES == 2 (see above)
//   void tile_end(int m_blk, int n_tile, int row)
template <int BLOCK_N, int SUBTILES, class Epi>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, int M,
               int num_n_tiles, int K, Epi epi) {
    using S = GemmSmem<BLOCK_N>;
    constexpr int kStages = S::kStages;
    constexpr bool kSplit = Epi::kSplitColumns;
    static_assert(BLOCK_N % 16 == 0 && BLOCK_N >= 16 && BLOCK_N <= 256, "invalid UMMA N");
    static_assert(2 * BLOCK_N <= 512, "two accumulator stages must fit TMEM");
    static_assert(SUBTILES == 1 || SUBTILES == 2, "one or two sub-blocks per tile");
    static_assert(!kSplit || BLOCK_N % 32 == 0, "column split needs two halves of whole 16-column chunks");
    constexpr uint32_t kTmemCols = tmem_cols_pow2(2 * BLOCK_N);

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kStages * S::kABytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * S::kStageBytes);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full = empty_bar + kStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float2* scratch = reinterpret_cast<float2*>(smem + kStages * S::kStageBytes + S::kBarrierBytes);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_m_blks = (M + kBlockM - 1) / kBlockM;
    const int num_tiles = num_m_blks * num_n_tiles;
    const int num_k16 = K / 16;
    const int num_kb = (num_k16 + 3) / 4;

    griddep_launch_dependents();
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_b);
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], kSplit ? 8 : 4);  // one arrive per draining epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_base_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;
    griddep_wait();

    if (warp == 0) {
        // TMA producer: warp-uniform loop, one elected lane issues.
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m_blk = tile % num_m_blks;
            const int n_tile = tile / num_m_blks;
            for (int sub = 0; sub < SUBTILES; ++sub) {
                const int n_row0 = (n_tile * SUBTILES + sub) * BLOCK_N;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&full_bar[stage], S::kStageBytes);
                        tma_load_2d(smem_a + stage * S::kABytes, &tma_a, &full_bar[stage], kb * kBlockK,
                                    m_blk * kBlockM);
                        tma_load_2d(smem_b + stage * S::kBBytes, &tma_b, &full_bar[stage], kb * kBlockK, n_row0);
                    }
                    __syncwarp();
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // MMA issuer: warp-uniform loop (descriptors live in uniform registers), one elected lane issues.
        constexpr uint32_t idesc = make_idesc_f16(kBlockM, BLOCK_N);
        const uint32_t a_lo0 = smem_desc_lo(smem_u32(smem_a));
        const uint32_t b_lo0 = smem_desc_lo(smem_u32(smem_b));
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            for (int sub = 0; sub < SUBTILES; ++sub) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_lo = a_lo0 + stage * (S::kABytes >> 4);
                    const uint32_t b_lo = b_lo0 + stage * (S::kBBytes >> 4);
                    const int nk = num_k16 - kb * 4;
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (k < nk)
                                umma_f16_ss(d_tmem, smem_desc_join(a_lo + 2 * k), smem_desc_join(b_lo + 2 * k), idesc,
                                            (kb | k) != 0 ? 1u : 0u);
                        }
                        umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
                    }
                    __syncwarp();
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (elect_one()) umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
                __syncwarp();
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
    } else {
        const int quarter = warp & 3;         // TMEM lane quarter this warp may access
        const int group = (warp - 2) >> 2;    // epilogue group 0 / 1
        const int row_in_tile = quarter * 32 + lane;
        const int c0 = kSplit ? group * (BLOCK_N / 2) : 0;
        const int c1 = kSplit ? c0 + BLOCK_N / 2 : BLOCK_N;
        int unit = 0;  // running (tile, sub) index of this CTA: stage = unit & 1, use count = unit >> 1
        int tile_iter = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_iter) {
            const int m_blk = tile % num_m_blks;
            const int n_tile = tile / num_m_blks;
            const int row = m_blk * kBlockM + row_in_tile;
            epi.tile_begin(m_blk, n_tile, row);
#pragma unroll
            for (int sub = 0; sub < SUBTILES; ++sub, ++unit) {
                const int acc = unit & 1;
                if (kSplit || acc == group) {
                    mbar_wait(&tmem_full[acc], (unit >> 1) & 1);
                    __syncwarp();
                    tc_fence_after();
                    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BLOCK_N;
                    epi.sub(sub, m_blk, n_tile, row, taddr, c0, c1);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                }
            }
            if constexpr (kSplit) {
                epi.tile_end(m_blk, n_tile, row);
            } else if constexpr (SUBTILES == 1) {
                if (((unit - 1) & 1) == group) epi.tile_end(m_blk, n_tile, row);
            } else {
                // group g drained sub-block g; fold group 1's state into group 0 through shared memory
                float2* slot = scratch + (tile_iter & 1) * kBlockM + row_in_tile;
                if (group == 1) epi.merge(1, slot);
                named_bar_sync(1 + quarter, 64);
                if (group == 0) {
                    epi.merge(0, slot);
                    epi.tile_end(m_blk, n_tile, row);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace vitad
