// Four-CTA cluster variant of the CTA-pair GEMM (gemm_pair.cuh): two MMA pairs (CTA ranks {0,1} and {2,3}) work on the
// SAME 256-row block of A and on two adjacent n-tiles.  Each CTA fetches only 64 of the 128 A rows its pair slot needs and
// TMA-multicasts them to the CTA with the same in-pair rank of the other pair, so A crosses the L2->SM fabric once per
// cluster instead of once per pair: a 256 x 208 pair tile drops from 58 KB to 42 KB of operand traffic per k-block.
// The fused GMM kernel runs at 85% of what the L2 can deliver; only 33 clusters of 4 fit the 148 SMs
// (132 SMs busy), which is the price of the scheme.
//
// Protocol changes against the pair kernel:
//   full[s]        per pair leader (ranks 0 and 2); both pairs' TMA traffic into a pair's two CTAs completes on it
//                  (cta_group::2 + multicast: the signal goes to the leader of each DESTINATION pair)
//   empty[s]       count 2: a slot is written by its own CTA and by the partner CTA of the other pair, so it is free
//                  only when BOTH pairs' MMAs have consumed it; each leader's tcgen05.commit multicasts to all four CTAs
//   tmem_full[a]   multicast to the two CTAs of the issuing pair only
//   tmem_empty[a]  in the pair leader; epilogue warps arrive on rank (own rank & ~1)
#pragma once
#include "gemm_pair.cuh"

namespace vitad {

__device__ __forceinline__ void tma_load_2d_pair_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                    uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
        "[%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1),
          "h"(cta_mask)
        : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(smem_u32(bar)), "h"(cta_mask)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_rank(uint64_t* bar, uint32_t target_rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(smem_u32(bar)), "r"(target_rank)
        : "memory");
}

template <int BLOCK_N, int SUBTILES, class Epi>
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm4_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, int M,
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
    const uint32_t rank4 = cluster_ctarank();  // 0..3
    const uint32_t rank = rank4 & 1;           // rank inside the MMA pair
    const uint32_t pair = rank4 >> 1;          // which pair of the cluster
    const int cluster_id = blockIdx.x >> 2;
    const int num_clusters = gridDim.x >> 2;
    const int num_m_blks = (M + kPairM - 1) / kPairM;
    const int num_tiles = num_m_blks * (num_n_tiles >> 1);  // cluster tiles: one m-block x two adjacent n-tiles
    const int num_k16 = K / 16;
    const int num_kb = (num_k16 + 3) / 4;

    if (threadIdx.x == 0) {
        VITAD_TL(0);
        VITAD_TLG(1);
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_b);
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 2);  // both pairs' MMAs must have consumed a slot: either pair's TMA writes into it
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
    griddep_wait();
    // The dependents are released only now, after this grid's own prerequisites have completed: the kernel launched right
    // behind this one may be an INDEPENDENT grid (the CTA-pair kernel that takes the last features on the SMs the 33 clusters
    // of four leave idle, mdn.cu).  It starts once every CTA of this grid is resident and past this point, so it can neither
    // take SMs away from the clusters nor read operands that are not written yet.
    griddep_launch_dependents();
    if (threadIdx.x == 0) VITAD_TL(2);

    if (warp == 0) {
        // TMA producer (both CTAs): warp-uniform loop, one elected lane issues.
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
            const int m_blk = tile % num_m_blks;
            const int n_tile = 2 * (tile / num_m_blks) + static_cast<int>(pair);
            // this CTA fetches 64 of the 128 A rows its MMA slot needs and multicasts them to the CTA with the same
            // in-pair rank of the other pair, which fetches the other 64
            const int row0 = m_blk * kPairM + static_cast<int>(rank) * kBlockM + static_cast<int>(pair) * (kBlockM / 2);
            const uint16_t a_mask = static_cast<uint16_t>((1u << rank) | (1u << (rank + 2)));
            for (int sub = 0; sub < SUBTILES; ++sub) {
                const int n_row0 = (n_tile * SUBTILES + sub) * BLOCK_N + static_cast<int>(rank) * (BLOCK_N / 2);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (elect_one()) {
                        if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * S::kStageBytes);
                        tma_load_2d_pair_mc(smem_a + stage * S::kABytes + pair * (S::kABytes / 2), &tma_a, &full_bar[stage],
                                            kb * kBlockK, row0, a_mask);
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
                            umma_commit_mc(&empty_bar[stage], 0xF);  // frees the slot in all four CTAs
                        }
                        __syncwarp();
                        if (++stage == kStages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    if (elect_one()) {
                        umma_commit_mc(&tmem_full[acc], static_cast<uint16_t>(3u << (2 * pair)));
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
            const int n_tile = 2 * (tile / num_m_blks) + static_cast<int>(pair);
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
                    if (lane == 0) mbar_arrive_rank(&tmem_empty[acc], rank4 & ~1u);
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
