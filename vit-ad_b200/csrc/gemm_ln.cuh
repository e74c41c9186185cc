// Residual GEMM + LayerNorm in one kernel (timm Block: x = x + proj(attn(...)); h = norm2(x)  and  x = x + fc2(...);
// h = norm1_next(x); call sites TransformerEncoder.py:150-165):
//     x[r][:]  = x[r][:] + A[r][:] . W^T + bias            fp32 residual stream, in place
//     h[r][:]  = LayerNorm(x[r][:]; gamma, beta, eps)      fp16 operand of the next GEMM
// for the encoder's N = 768.  The two GEMMs that end in a residual (N = 768: 75 CTA-pair tiles on 74 SM pairs) were the
// encoder kernels furthest from their roofline, and the LayerNorm that followed each re-read the 19.5 MB stream they had
// just written (26 launches, 9 % of the step).  Here one cluster of FOUR CTAs owns a 256-row block over all 768 columns:
//   * two MMA pairs (CTA ranks {0,1} and {2,3}), pair p computes columns [384 p, 384 p + 384) as two M=256 x N=192
//     tcgen05.mma accumulators (384 fp32 TMEM columns per CTA); the pairs share the A block by TMA multicast exactly as
//     the fused GMM kernel does (gemm_quad.cuh): each CTA fetches 64 of its 128 rows and multicasts them to the CTA with
//     the same in-pair rank of the other pair;
//   * one tile per cluster (25 clusters at batch 32: a single wave, no tail), so when the accumulator is complete the
//     operand ring is dead and its 200 KB become the epilogue's staging memory: the residual tile arrives by TMA
//     (12 boxes of 128 rows x 32 fp32, 128-byte swizzle), every thread owns one row (the tcgen05.ld layout), adds
//     accumulator + bias + residual, writes x back in place and a TMA store sends it out — no register-staged
//     transposes, every global access a full line;
//   * LayerNorm statistics per thread = per row: (mean, M2) per 32-column unit, merged with Chan's parallel update; the
//     four partials of a row (2 epilogue groups x 2 pairs) meet through shared memory / DSMEM (st.shared::cluster +
//     mbarrier, release/acquire at cluster scope) and are merged in a fixed order, so a row's result does not depend on
//     the batch it is in; the second pass re-reads x from shared memory, normalises and TMA-stores fp16 h.
#pragma once
#include "gemm_quad.cuh"
#include "ln_tree.cuh"

namespace vitad {

constexpr int kLnN = 768;          // output width (the whole LayerNorm row)
constexpr int kLnPairN = 384;      // columns per MMA pair
constexpr int kLnHalfN = 192;      // one tcgen05.mma accumulator
constexpr int kLnUnits = kLnPairN / 32;

struct LnSmem {
    static constexpr int kABytes = kBlockM * kBlockK * 2;                // 128 rows (64 own + 64 multicast)
    static constexpr int kBBytes = 2 * (kLnHalfN / 2) * kBlockK * 2;     // two halves x 96 weight rows
    static constexpr int kStageBytes = kABytes + kBBytes;                // 40 KB
    static constexpr int kStages = 5;
    static constexpr int kUnitBytes = kBlockM * 32 * 4;                  // residual box: 128 rows x 32 fp32
    static constexpr int kBarrierBytes = 512;
    static constexpr int kStatsBytes = 2 * 4 * kBlockM * 8;              // [pair][group <= 4][row] float2
    static constexpr int kVecBytes = 3 * kLnPairN * 4;                   // bias | gamma | beta of this pair's columns
    static constexpr int kTotalBytes = kStages * kStageBytes + kBarrierBytes + kStatsBytes + kVecBytes + 1024;
    static_assert(kLnUnits * kUnitBytes <= kStages * kStageBytes, "the residual tile must fit the retired operand ring");
    static_assert(kTotalBytes <= kSmemBudget, "shared memory");
    static_assert((kBBytes / 2) % 1024 == 0 && kABytes % 1024 == 0, "1024-byte aligned operand tiles");
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void st_cluster_f2(const void* local, uint32_t target_rank, float a, float b) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "st.shared::cluster.v2.f32 [ra], {%2, %3};\n\t}"
        ::"r"(smem_u32(local)), "r"(target_rank), "f"(a), "f"(b)
        : "memory");
}
// release at cluster scope: the arriving warp's shared-memory / DSMEM writes are visible to whoever acquires the phase
__device__ __forceinline__ void mbar_arrive_rank_release(uint64_t* bar, uint32_t target_rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(smem_u32(bar)), "r"(target_rank)
        : "memory");
}

// 16-byte chunk `chunk` (0..7) of row `row` inside a box of 128-byte rows written by TMA with the 128-byte swizzle
// (box base 1024-byte aligned): chunk index XOR (row mod 8).
__device__ __forceinline__ uint32_t sw128(uint32_t base, int row, int chunk) {
    return base + static_cast<uint32_t>(row) * 128u + (static_cast<uint32_t>(chunk ^ (row & 7)) << 4);
}


struct GemmLnParams {
    const float* bias;   // [768]
    const float* gamma;  // [768]
    const float* beta;   // [768]
    float eps;
    int M, K;
};

// EPI_WARPS = 8 or 16: the EPI_WARPS / 4 warps of a TMEM lane quarter (group g = 0..) take the 32-column units g, g + G, ...
template <int EPI_WARPS>
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(64 + 32 * EPI_WARPS, 1)
gemm_ln_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_w,
               const __grid_constant__ CUtensorMap tma_x_ld, const __grid_constant__ CUtensorMap tma_x_st,
               const __grid_constant__ CUtensorMap tma_h_st, const __grid_constant__ CUtensorMap tma_h_st32,
               const GemmLnParams p) {
    using S = LnSmem;
    constexpr int kStages = S::kStages;
    constexpr int G = EPI_WARPS / 4;            // epilogue groups
    constexpr int kUnitsPerWarp = kLnUnits / G;  // 6 or 3
    static_assert(EPI_WARPS == 8 || EPI_WARPS == 16, "8 or 16 epilogue warps");
    constexpr uint32_t kTmemCols = 512;  // 384 used
    constexpr int kPairM = 2 * kBlockM;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kStages * S::kABytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * S::kStageBytes);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full = empty_bar + kStages;
    uint64_t* stats_bar = tmem_full + 1;
    uint64_t* unit_bar = stats_bar + 1;  // [kLnUnits] residual boxes
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(unit_bar + kLnUnits);
    float2* stats = reinterpret_cast<float2*>(smem + kStages * S::kStageBytes + S::kBarrierBytes);  // [2][4][128]
    float* vec = reinterpret_cast<float*>(smem + kStages * S::kStageBytes + S::kBarrierBytes + S::kStatsBytes);
    static_assert((2 * kStages + 2 + kLnUnits) * 8 + 8 <= S::kBarrierBytes, "barrier area");

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank4 = cluster_ctarank();  // 0..3
    const uint32_t rank = rank4 & 1;           // rank inside the MMA pair
    const uint32_t pair = rank4 >> 1;          // which pair of the cluster = which 384-column half
    const int m_blk = blockIdx.x >> 2;         // one 256-row block per cluster
    const int num_k16 = p.K / 16;
    const int num_kb = (num_k16 + 3) / 4;
    const int col0 = static_cast<int>(pair) * kLnPairN;

    griddep_launch_dependents();
    if (threadIdx.x == 0) {
        VITAD_TL(0);
        VITAD_TLG(1);
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_w);
        tma_prefetch_desc(&tma_x_ld);
        tma_prefetch_desc(&tma_x_st);
        tma_prefetch_desc(&tma_h_st);
        tma_prefetch_desc(&tma_h_st32);
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 2);  // both pairs' MMAs must have consumed a slot: either pair's TMA writes into it
        }
        mbar_init(tmem_full, 1);
        // every epilogue thread of this CTA and of the same-rank CTA of the other pair arrives itself: a thread's
        // release covers its own shared-memory / DSMEM stores without relying on warp-level cumulativity
        mbar_init(stats_bar, 2 * 32 * EPI_WARPS);
        for (int i = 0; i < kLnUnits; ++i) mbar_init(&unit_bar[i], 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_pair(tmem_base_slot, kTmemCols);
    // bias / gamma / beta of this pair's columns: parameters, not produced by the preceding kernels
    for (int i = threadIdx.x; i < 3 * kLnPairN; i += blockDim.x) {
        const int which = i / kLnPairN, c = i - which * kLnPairN;
        const float* src = which == 0 ? p.bias : (which == 1 ? p.gamma : p.beta);
        vec[i] = __ldg(src + col0 + c);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // barriers of all four CTAs initialised before any remote signal
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;
    griddep_wait();  // operands and the residual stream belong to the preceding kernels up to here
    if (threadIdx.x == 0) VITAD_TL(2);

    if (warp == 0) {
        // TMA producer: 64 of this CTA's 128 A rows (multicast to the same-rank CTA of the other pair) + its 96 + 96 weight rows
        int stage = 0;
        uint32_t phase = 0;
        const int row0 = m_blk * kPairM + static_cast<int>(rank) * kBlockM + static_cast<int>(pair) * (kBlockM / 2);
        const uint16_t a_mask = static_cast<uint16_t>((1u << rank) | (1u << (rank + 2)));
        const int n_row0 = col0 + static_cast<int>(rank) * (kLnHalfN / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (elect_one()) {
                if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * S::kStageBytes);
                tma_load_2d_pair_mc(smem_a + stage * S::kABytes + pair * (S::kABytes / 2), &tma_a, &full_bar[stage], kb * kBlockK,
                                    row0, a_mask);
                tma_load_2d_pair(smem_b + stage * S::kBBytes, &tma_w, &full_bar[stage], kb * kBlockK, n_row0);
                tma_load_2d_pair(smem_b + stage * S::kBBytes + S::kBBytes / 2, &tma_w, &full_bar[stage], kb * kBlockK,
                                 n_row0 + kLnHalfN);
                if (kb == 0) VITAD_TL(3);
                VITAD_TL(4);
            }
            __syncwarp();
            if (++stage == kStages) {
                stage = 0;
                phase ^= 1;
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            // MMA issuer (pair leader): two M=256 x N=192 accumulators per k-slice
            constexpr uint32_t idesc = make_idesc_f16(kPairM, kLnHalfN);
            const uint32_t a_lo0 = smem_desc_lo(smem_u32(smem_a));
            const uint32_t b_lo0 = smem_desc_lo(smem_u32(smem_b));
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (kb == 0 && elect_one()) VITAD_TL(8);
                const uint32_t a_lo = a_lo0 + stage * (S::kABytes >> 4);
                const uint32_t b_lo = b_lo0 + stage * (S::kBBytes >> 4);
                const int nk = num_k16 - kb * 4;
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (k < nk) {
                            umma_f16_ss_pair(tmem_base, smem_desc_join(a_lo + 2 * k), smem_desc_join(b_lo + 2 * k), idesc,
                                             (kb | k) != 0 ? 1u : 0u);
                            umma_f16_ss_pair(tmem_base + kLnHalfN, smem_desc_join(a_lo + 2 * k),
                                             smem_desc_join(b_lo + (S::kBBytes >> 5) + 2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
                        }
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
                umma_commit_mc(tmem_full, static_cast<uint16_t>(3u << (2 * pair)));
                VITAD_TL(10);
            }
            __syncwarp();
        }
    } else {
        const int quarter = warp & 3;         // TMEM lane quarter this warp may access
        const int group = (warp - 2) >> 2;    // 0 .. G-1
        const int row_in_cta = quarter * 32 + lane;
        const int grow0 = m_blk * kPairM + static_cast<int>(rank) * kBlockM;  // first global row of this CTA
        uint8_t* tile = smem;  // the retired operand ring: unit j at j * kUnitBytes
        const float* vbias = vec;
        const float* vgamma = vec + kLnPairN;
        const float* vbeta = vec + 2 * kLnPairN;

        mbar_wait(tmem_full, 0);
        __syncwarp();
        tc_fence_after();
        if (threadIdx.x == 64) VITAD_TL(20);
        // every MMA of this pair (and with it every TMA write into this CTA's ring) has completed: fetch the residual tile
        if (warp == 2 && elect_one()) {
#pragma unroll 1
            for (int j = 0; j < kLnUnits; ++j) {
                mbar_arrive_expect_tx(&unit_bar[j], S::kUnitBytes);
                tma_load_2d(tile + j * S::kUnitBytes, &tma_x_ld, &unit_bar[j], col0 + j * 32, grow0);
            }
            VITAD_TL(21);
        }
        __syncwarp();

        // ---- pass 1: x = acc + bias + resid (in place in shared memory), TMA store, per-row statistics
        float mean_run = 0.f, m2_run = 0.f;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
        for (int u = 0; u < kUnitsPerWarp; ++u) {
            const int j = group * kUnitsPerWarp + u;  // a group owns a contiguous column range (pass 2 pairs adjacent units)
            uint32_t acc[32];
            tmem_ld_x32(taddr + j * 32, acc);
            mbar_wait(&unit_bar[j], 0);
            if (threadIdx.x == 64) VITAD_TL(22 + u);
            const uint32_t ubase = smem_u32(tile + j * S::kUnitBytes);
            float x[32];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 r = ld_shared_f4(sw128(ubase, row_in_cta, c));
                x[4 * c + 0] = r.x, x[4 * c + 1] = r.y, x[4 * c + 2] = r.z, x[4 * c + 3] = r.w;
            }
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 32; ++c)  // the residual epilogue's association (gemm_staged.cuh SEpiResidualF32): r + (a + b)
                x[c] = __fadd_rn(x[c], __fadd_rn(__uint_as_float(acc[c]), vbias[j * 32 + c]));
            float mu, q;
            tree_unit_stats(x, mu, q);  // ln_tree.cuh: the arithmetic the standalone kernel reproduces bit for bit
            if (u == 0) {
                mean_run = mu;
                m2_run = q;
            } else {
                tree_merge(mean_run, m2_run, 32.0f * u, mu, q, 32.0f);
            }
#pragma unroll
            for (int c = 0; c < 8; ++c)
                st_shared_v4(sw128(ubase, row_in_cta, c), __float_as_uint(x[4 * c + 0]), __float_as_uint(x[4 * c + 1]),
                             __float_as_uint(x[4 * c + 2]), __float_as_uint(x[4 * c + 3]));
            fence_proxy_async();
            __syncwarp();
            if (elect_one()) {
                tma_store_2d(&tma_x_st, tile + j * S::kUnitBytes + quarter * (32 * 128), col0 + j * 32, grow0 + quarter * 32);
                tma_store_commit();
            }
            __syncwarp();
        }
        if (threadIdx.x == 64) VITAD_TL(30);
        // ---- statistics of the whole row: 2 pairs x G groups partials, merged in a fixed order in every CTA
        {
            float2* mine = stats + (pair * 4 + group) * kBlockM + row_in_cta;
            *mine = make_float2(mean_run, m2_run);
            st_cluster_f2(mine, rank4 ^ 2u, mean_run, m2_run);
            mbar_arrive_rank_release(stats_bar, rank4);
            mbar_arrive_rank_release(stats_bar, rank4 ^ 2u);
            mbar_wait_cluster(stats_bar, 0);
        }
        if (threadIdx.x == 64) VITAD_TL(31);
        float mean = 0.f, m2 = 0.f;
        {
            constexpr float kPart = static_cast<float>(kLnPairN / G);
#pragma unroll
            for (int pp = 0; pp < 2; ++pp) {
                float mp = 0.f, qp = 0.f;
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const float2 t = stats[(pp * 4 + g) * kBlockM + row_in_cta];
                    if (g == 0) {
                        mp = t.x;
                        qp = t.y;
                    } else {
                        tree_merge(mp, qp, kPart * g, t.x, t.y, kPart);
                    }
                }
                if (pp == 0) {
                    mean = mp;
                    m2 = qp;
                } else {
                    tree_merge(mean, m2, static_cast<float>(kLnPairN), mp, qp, static_cast<float>(kLnPairN));
                }
            }
        }
        const float rstd = tree_rstd(m2, p.eps);

        // ---- pass 2: h = (x - mean) * rstd * gamma + beta  -> fp16, staged over the (dead) x slice, TMA store
        if (elect_one()) tma_store_wait_read();  // this warp's x stores have finished reading shared memory
        __syncwarp();
        if (threadIdx.x == 64) VITAD_TL(32);
        // Two adjacent units per step: 64 fp16 columns = 128-byte rows, one TMA store of 32 rows (the TMA unit moves whole
        // lines: 64-byte rows cost as much per row as 128-byte ones).  The fp16 box is staged, 128-byte swizzled, over the
        // first unit's (dead) x slice of this warp.
        static_assert(kUnitsPerWarp % 2 == 0 || kUnitsPerWarp == 3, "units per warp");
#pragma unroll 1
        for (int u = 0; u < kUnitsPerWarp; u += 2) {
            const int j = group * kUnitsPerWarp + u;
            const bool two = u + 1 < kUnitsPerWarp;  // compile-time per iteration for even counts; the odd tail is a half box
            const uint32_t ubase = smem_u32(tile + j * S::kUnitBytes);
            const uint32_t ubase2 = ubase + S::kUnitBytes;
            uint32_t hp[32];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                if (half == 1 && !two) break;
                const uint32_t ub = half ? ubase2 : ubase;
                const int cbase = (j + half) * 32;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 r = ld_shared_f4(sw128(ub, row_in_cta, c));
                    const float y0 = tree_normalize(r.x, mean, rstd, vgamma[cbase + 4 * c + 0], vbeta[cbase + 4 * c + 0]);
                    const float y1 = tree_normalize(r.y, mean, rstd, vgamma[cbase + 4 * c + 1], vbeta[cbase + 4 * c + 1]);
                    const float y2 = tree_normalize(r.z, mean, rstd, vgamma[cbase + 4 * c + 2], vbeta[cbase + 4 * c + 2]);
                    const float y3 = tree_normalize(r.w, mean, rstd, vgamma[cbase + 4 * c + 3], vbeta[cbase + 4 * c + 3]);
                    hp[half * 16 + 2 * c] = pack_h2(y0, y1);
                    hp[half * 16 + 2 * c + 1] = pack_h2(y2, y3);
                }
            }
            __syncwarp();  // every lane has read its x rows before the slice is overwritten with h
            const uint32_t hslice = ubase + quarter * (32 * 128);  // 1024-byte aligned: swizzle pattern = row & 7 of the slice
            if (two) {
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    st_shared_v4(sw128(hslice, lane, c), hp[4 * c], hp[4 * c + 1], hp[4 * c + 2], hp[4 * c + 3]);
            } else {
#pragma unroll
                for (int c = 0; c < 4; ++c) st_shared_v4(hslice + lane * 64 + c * 16, hp[4 * c], hp[4 * c + 1], hp[4 * c + 2], hp[4 * c + 3]);
            }
            fence_proxy_async();
            __syncwarp();
            if (elect_one()) {
                tma_store_2d(two ? &tma_h_st : &tma_h_st32, tile + j * S::kUnitBytes + quarter * (32 * 128), col0 + j * 32,
                             grow0 + quarter * 32);
                tma_store_commit();
            }
            __syncwarp();
        }
        if (threadIdx.x == 64) VITAD_TL(40);
        if (elect_one()) tma_store_wait_read();  // shared memory must outlive the bulk stores' reads
        __syncwarp();
        if (threadIdx.x == 64) VITAD_TL(41);
    }

    tc_fence_before();
    __syncwarp();
    cluster_sync_all();  // no CTA may exit (or free TMEM) while a peer can still signal it or write its shared memory
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
