// CTA-pair variant of the GEMM main loop (tcgen05 cta_group::2): two CTAs on the two SMs of a TPC
// cooperate on a 256-row x BLOCK_N tile.  Each CTA stages ITS 128 rows of A and HALF of the B tile
// (BLOCK_N/2 weight rows) per k-block, the leader CTA issues one M=256 MMA that reads both CTAs'
// shared memory, and each CTA's TMEM receives the accumulator rows of its own 128 tokens.  Compared with
// the single-CTA kernel this cuts L2->SM operand traffic per output by the B half (the single-CTA MDN kernel
// measured at the L2 delivery limit: ~12.6 TB/s at 54% tensor-pipe activity) and frees shared memory for a
// deeper ring.  Same warp roles, epilogue interface and epilogue-group modes as gemm_core.cuh.
//
// Cross-CTA protocol (rank 0 = leader):
//   full[s]        leader only; both producers' TMA transactions complete on it (expect_tx = 2 x stage bytes)
//   empty[s]       both CTAs; the leader's tcgen05.commit multicasts the arrive to both
//   tmem_full[a]   both CTAs; multicast commit after the last k-block of a unit
//   tmem_empty[a]  leader only; epilogue warps of BOTH CTAs arrive on it (remote arrive from rank 1)
#pragma once
#include "gemm_core.cuh"

namespace vitad {

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> rank 0

__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
        "%4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
        // default semantics: what is ordered here are this warp's TMEM reads (tcgen05.wait::ld +
        // tcgen05.fence::before_thread_sync precede the arrive), not its global stores; a
        // .release.cluster arrive costs MEMBAR + ERRBAR = a full drain of the epilogue's stores (15-20% of
        // the epilogue warps' samples in ncu)
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(smem_u32(bar))
        : "memory");
}
// Wait with cluster-scope acquire: the arrivals come from the peer CTA as well.
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    long long t0 = clock64();
    uint32_t spins = 0;
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return;
        if (((++spins) & 0x3ff) == 0 && clock64() - t0 > VITAD_MBAR_TIMEOUT_CYCLES) {
            printf("vitad: cluster mbarrier timeout block %d thread %d\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                  uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
        : "memory");
}

template <int BLOCK_N>
struct PairSmem {
    static constexpr int kABytes = kBlockM * kBlockK * 2;          // this CTA's 128 rows
    static constexpr int kBBytes = (BLOCK_N / 2) * kBlockK * 2;    // this CTA's half of the weight rows
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kBarrierBytes = 512;
    static constexpr int kScratchBytes = 2 * kBlockM * 8;
    static constexpr int kTailBytes = kBarrierBytes + kScratchBytes;
    static constexpr int kStages = (kSmemBudget - 1024 - kTailBytes) / kStageBytes > 10
                                       ? 10
                                       : (kSmemBudget - 1024 - kTailBytes) / kStageBytes;
    static constexpr int kTotalBytes = kStages * kStageBytes + kTailBytes + 1024;
    static_assert(kBBytes % 1024 == 0, "half B stage must keep 1024-byte alignment (BLOCK_N % 16 == 0)");
    static_assert(2 * kStages * 8 + 48 <= kBarrierBytes, "barrier area");
};

// kIndependent: the grid does not consume the preceding kernel's results and is launched (programmatic stream
// serialization) to run BESIDE it: no wait before the main loop, but one before exit, so that this grid's completion still
// implies the completion of everything before it in the stream (the next kernel only waits for its direct predecessor).
template <int BLOCK_N, int SUBTILES, class Epi, bool kIndependent = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm2_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, int M,
                int num_n_tiles, int K, Epi epi) {
    using S = PairSmem<BLOCK_N>;
    constexpr int kStages = S::kStages;
    constexpr bool kSplit = Epi::kSplitColumns;
    static_assert(BLOCK_N % 16 == 0 && BLOCK_N >= 32 && BLOCK_N <= 256, "invalid UMMA N for M=256");
    static_assert(SUBTILES == 1 || SUBTILES == 2, "one or two sub-blocks per tile");
    static_assert(!kSplit || BLOCK_N % 32 == 0, "column split needs two halves of whole 16-column chunks");
    constexpr uint32_t kTmemCols = tmem_cols_pow2(2 * BLOCK_N);
    constexpr int kPairM = 2 * kBlockM;

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
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;
    const int num_m_blks = (M + kPairM - 1) / kPairM;
    const int num_tiles = num_m_blks * num_n_tiles;
    const int num_k16 = K / 16;
    const int num_kb = (num_k16 + 3) / 4;

    griddep_launch_dependents();
    if (threadIdx.x == 0) {
        VITAD_TL(0);
        VITAD_TLG(1);
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_b);
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], (kSplit ? 8 : 4) * 2);  // draining warps of both CTAs
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_pair(tmem_base_slot, kTmemCols);
    tc_fence_before();
    __syncwarp();
    cluster_sync_all();  // barriers of both CTAs initialised before any remote signal
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;
    if constexpr (!kIndependent) griddep_wait();
    if (threadIdx.x == 0) VITAD_TL(2);

    if (warp == 0) {
        // TMA producer (both CTAs): warp-uniform loop, one elected lane issues.
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
            const int m_blk = tile % num_m_blks;
            const int n_tile = tile / num_m_blks;
            const int row0 = m_blk * kPairM + static_cast<int>(rank) * kBlockM;
            for (int sub = 0; sub < SUBTILES; ++sub) {
                const int n_row0 = (n_tile * SUBTILES + sub) * BLOCK_N + static_cast<int>(rank) * (BLOCK_N / 2);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (elect_one()) {
                        if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * S::kStageBytes);
                        tma_load_2d_pair(smem_a + stage * S::kABytes, &tma_a, &full_bar[stage], kb * kBlockK, row0);
                        tma_load_2d_pair(smem_b + stage * S::kBBytes, &tma_b, &full_bar[stage], kb * kBlockK, n_row0);
                        if (tile == cluster_id && sub == 0 && kb == 0) VITAD_TL(3);
                        VITAD_TL(4);
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
        if (rank == 0) {
            // MMA issuer (leader CTA only): warp-uniform loop, one elected lane issues.
            constexpr uint32_t idesc = make_idesc_f16(kPairM, BLOCK_N);
            const uint32_t a_lo0 = smem_desc_lo(smem_u32(smem_a));
            const uint32_t b_lo0 = smem_desc_lo(smem_u32(smem_b));
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            int tl_i = 0;
            (void)tl_i;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                for (int sub = 0; sub < SUBTILES; ++sub) {
                    mbar_wait_cluster(&tmem_empty[acc], acc_phase ^ 1);
                    tc_fence_after();
                    if (elect_one() && tl_i < 12) VITAD_TL(8 + 3 * tl_i);
                    const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                    for (int kb = 0; kb < num_kb; ++kb) {
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        if (kb == 0 && elect_one() && tl_i < 12) VITAD_TL(9 + 3 * tl_i);
                        const uint32_t a_lo = a_lo0 + stage * (S::kABytes >> 4);
                        const uint32_t b_lo = b_lo0 + stage * (S::kBBytes >> 4);
                        const int nk = num_k16 - kb * 4;
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if (k < nk)
                                    umma_f16_ss_pair(d_tmem, smem_desc_join(a_lo + 2 * k), smem_desc_join(b_lo + 2 * k),
                                                     idesc, (kb | k) != 0 ? 1u : 0u);
                            }
                            umma_commit_pair(&empty_bar[stage]);
                        }
                        __syncwarp();
                        if (++stage == kStages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    if (elect_one()) {
                        umma_commit_pair(&tmem_full[acc]);
                        if (tl_i < 12) VITAD_TL(10 + 3 * tl_i);
                    }
                    ++tl_i;
                    __syncwarp();
                    if (++acc == 2) {
                        acc = 0;
                        acc_phase ^= 1;
                    }
                }
            }
        }
    } else {
        const int quarter = warp & 3;
        const int group = (warp - 2) >> 2;
        const int row_in_tile = quarter * 32 + lane;
        const int c0 = kSplit ? group * (BLOCK_N / 2) : 0;
        const int c1 = kSplit ? c0 + BLOCK_N / 2 : BLOCK_N;
        int unit = 0;
        int tile_iter = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++tile_iter) {
            const int m_blk = tile % num_m_blks;
            const int n_tile = tile / num_m_blks;
            const int row = m_blk * kPairM + static_cast<int>(rank) * kBlockM + row_in_tile;
            epi.tile_begin(m_blk, n_tile, row);
#pragma unroll
            for (int sub = 0; sub < SUBTILES; ++sub, ++unit) {
                const int acc = unit & 1;
                if (kSplit || acc == group) {
                    mbar_wait(&tmem_full[acc], (unit >> 1) & 1);
                    __syncwarp();
                    tc_fence_after();
                    if (threadIdx.x == 64 && unit < 8) VITAD_TL(44 + 2 * unit);
                    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BLOCK_N;
                    epi.sub(sub, m_blk, n_tile, row, taddr, c0, c1);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);
                    if (threadIdx.x == 64 && unit < 8) VITAD_TL(45 + 2 * unit);
                }
            }
            if constexpr (kSplit) {
                epi.tile_end(m_blk, n_tile, row);
            } else if constexpr (SUBTILES == 1) {
                if (((unit - 1) & 1) == group) epi.tile_end(m_blk, n_tile, row);
            } else {
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
    __syncwarp();
    cluster_sync_all();  // neither CTA may exit (or free TMEM) while its peer can still signal or read it
    if constexpr (kIndependent) griddep_wait();
    if (threadIdx.x == 0) {
        VITAD_TL(60);
        VITAD_TLG(61);
    }
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, kTmemCols);
    }
}

}  // namespace vitad
