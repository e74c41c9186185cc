// Encoder GEMM main loop, third generation: CTA pairs (cta_group::2, 256-row tiles) with
//   (1) wide tiles + a cut tail.  The encoder's GEMMs are small (M = 6336 at batch 32): with 256x96 tiles the
//       operand traffic (69.8 FLOP per L2 byte) saturates the ~12.6 TB/s the L2 delivers to the SMs at 37% tensor-pipe
//       activity, with 256x256 tiles (128 FLOP/B) the persistent schedule ends in an almost empty wave
//       (N = 768: 75 tiles on 74 SM pairs).  Here full waves run 256 x BLOCK_N tiles and the tiles that would not
//       fill a wave are cut into `tail_split` narrower tiles (runtime UMMA N, second TMA map for B) so the last wave is
//       short AND full;
//   (2) a staged, coalesced epilogue.  tcgen05.ld hands each thread one accumulator ROW; storing from that
//       layout makes every warp-level load/store touch 32 cache lines (the fp32 residual epilogue measured 200 cycles
//       per column and was slower than the MMAs it follows).  Each epilogue warp now transposes 32x32 accumulator
//       blocks through a private 4 KB swizzled shared-memory slab and finishes in a row-contiguous layout (8 lanes
//       = 128 bytes of one row), where residual/pos-embed loads are prefetched before the TMEM wait and every
//       global access is a full-line transaction.
// Warp roles, barriers and the cross-CTA protocol are those of gemm_pair.cuh.
#pragma once
#include "gemm_pair.cuh"

namespace vitad {

struct TileSched {
    int num_m;       // 256-row blocks
    int big_tiles;   // full-width tiles; tile t -> (m_blk = t % num_m, n_tile = t / num_m)
    int tail_tiles;  // = (tiles cut) * tail_split
    int tail_split;  // pieces per cut tile (1, 2, 4, 8)
    int tail_w;      // BLOCK_N / tail_split (multiple of 32)
    // implicit 3x3 convolution (conv_pitch > 0): A is a zero-bordered NHWC activation [B*(g+2)*(g+2), C]; the K loop walks
    // 9 taps x kpt k-blocks, tap (ty, tx) reads the A rows shifted by (ty-1)*conv_pitch + (tx-1), conv_pitch =
    // g+2 (rows outside the tensor are zero-filled by TMA and only reach border rows, which the epilogue drops)
    int conv_pitch;
    // Tapped / split A operand (kpt > 0): the K loop walks taps x kpt k-blocks; tap t reads A columns from t * tap_cols
    // (explicit im2col layouts; 0 for the row-shifted taps of the implicit convolution above, which sets conv_pitch too).
    // split_kb > 0 — split-fp16 operands: A rows hold [hi (C) | lo (C)] per tap, W rows [w_hi | w_hi | w_lo] per tap
    // (C = 64 * split_kb), so kpt = 3 * split_kb and the three segments of a tap read A's hi, lo, hi columns:
    // hi*w_hi + lo*w_hi + hi*w_lo, every fp16 x fp16 product exact in the fp32 accumulator (fp32-grade products on the fp16
    // tensor cores; the dropped lo*w_lo term is 2^-22 relative).
    int kpt, tap_cols, split_kb;
};

template <int BLOCK_N, int EPI_WARPS>
struct StagedSmem {
    static constexpr int kABytes = kBlockM * kBlockK * 2;
    static constexpr int kBBytes = (BLOCK_N / 2) * kBlockK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kBarrierBytes = 512;
    static constexpr int kSlabBytes = 32 * 128;          // one epilogue warp: 32 rows x 32 fp32
    static constexpr int kEpiBytes = EPI_WARPS * kSlabBytes;
    static constexpr int kAvail = kSmemBudget - 1024 - kBarrierBytes - kEpiBytes;
    static constexpr int kStages = kAvail / kStageBytes > 10 ? 10 : kAvail / kStageBytes;
    static constexpr int kTotalBytes = kStages * kStageBytes + kBarrierBytes + kEpiBytes + 1024;
    static_assert(kStages >= 3, "tile too large for shared memory");
    static_assert(kBBytes % 1024 == 0, "half B stage must keep 1024-byte alignment");
    static_assert(2 * kStages * 8 + 48 <= kBarrierBytes, "barrier area");
};

struct TileCoord {
    int m_blk, n0, w;
    bool tail;
};
template <int BLOCK_N>
__device__ __forceinline__ TileCoord decode_tile(const TileSched& s, int tile) {
    TileCoord t;
    if (tile < s.big_tiles) {
        t.m_blk = tile % s.num_m;
        t.n0 = (tile / s.num_m) * BLOCK_N;
        t.w = BLOCK_N;
        t.tail = false;
    } else {
        const int u = tile - s.big_tiles;
        const int big = s.big_tiles + u / s.tail_split;
        const int part = u % s.tail_split;
        t.m_blk = big % s.num_m;
        t.n0 = (big / s.num_m) * BLOCK_N + part * s.tail_w;
        t.w = s.tail_w;
        t.tail = true;
    }
    return t;
}

struct RowCtx {
    int a, b;
};

// Epilogue functor interface (all __device__, called by the 8 epilogue warps):
//   static constexpr bool kRowCtx                 rows need a decoded context (shuffled from the lane that owns the row)
//   RowCtx row_ctx(int row)                       once per tile, lane l decodes row_base + l
//   bool   direct(int col0)                       warp-uniform: this 32-column unit bypasses the staging slab
//   void   direct_unit(int row, RowCtx, int col0, const uint32_t (&acc)[32])      thread = row layout
//   Pre    prefetch(int row, RowCtx, int col)     row-contiguous layout: issue loads the store will need
//   ColC   col_const(int col)                     per-lane column constants (bias) of the 4 columns at `col`
//   void   store(int row, RowCtx, int col, float4 acc, Pre, ColC)
template <class Epi>
__device__ __forceinline__ void epilogue_unit(const Epi& epi, int row_base, int lane, RowCtx my_ctx, int col0,
                                              uint32_t taddr, uint32_t slab) {
    uint32_t r[32];
    if (epi.direct(col0)) {
        tmem_ld_x32(taddr, r);
        tmem_ld_wait();
        epi.direct_unit(row_base + lane, my_ctx, col0, r);
        return;
    }
    const int sub = lane >> 3, ch = lane & 7;
    const int col = col0 + ch * 4;
    RowCtx ctx[8];
    typename Epi::Pre pre[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if constexpr (Epi::kRowCtx) {
            ctx[i].a = __shfl_sync(0xffffffffu, my_ctx.a, 4 * i + sub);
            ctx[i].b = __shfl_sync(0xffffffffu, my_ctx.b, 4 * i + sub);
        } else {
            ctx[i] = RowCtx{0, 0};
        }
        pre[i] = epi.prefetch(row_base + 4 * i + sub, ctx[i], col);
    }
    const typename Epi::ColC cc = epi.col_const(col);
    tmem_ld_x32(taddr, r);
    tmem_ld_wait();
    const uint32_t wrow = slab + lane * 128;
#pragma unroll
    for (int j = 0; j < 8; ++j)
        st_shared_v4(wrow + ((j ^ (lane & 7)) << 4), r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int rr = 4 * i + sub;
        const float4 a = ld_shared_f4(slab + rr * 128 + ((ch ^ (rr & 7)) << 4));
        epi.store(row_base + rr, ctx[i], col, a, pre[i], cc);
    }
    __syncwarp();
}

// EPI_WARPS = 8 or 16 epilogue warps (2 or 4 per scheduler).  A warp may only touch TMEM lane quarter warp % 4,
// so the EPI_WARPS / 4 warps of a quarter take the tile's 32-column units round-robin.  16 warps hide the
// TMEM -> slab -> global latency chain of math-heavy epilogues (GELU: 2 MUFU + 12 FP32 ops per element) that 8
// warps leave exposed; their register budget is 112 per thread.
template <int BLOCK_N, int EPI_WARPS, class Epi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 32 * EPI_WARPS, 1)
gemm3_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                const __grid_constant__ CUtensorMap tma_b_tail, const TileSched sched, int K, Epi epi) {
    using S = StagedSmem<BLOCK_N, EPI_WARPS>;
    constexpr int kStages = S::kStages;
    static_assert(EPI_WARPS == 8 || EPI_WARPS == 16, "8 or 16 epilogue warps");
    static_assert(BLOCK_N == 128 || BLOCK_N == 192 || BLOCK_N == 256, "BLOCK_N 128, 192 or 256");
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
    uint8_t* epi_slabs = smem + kStages * S::kStageBytes + S::kBarrierBytes;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;
    const int num_tiles = sched.big_tiles + sched.tail_tiles;
    const int num_k16 = K / 16;
    const int num_kb = (num_k16 + 3) / 4;

    griddep_launch_dependents();
    if (threadIdx.x == 0) {
        VITAD_TL(0);
        VITAD_TLG(1);
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_b);
        tma_prefetch_desc(&tma_b_tail);
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 2 * EPI_WARPS);  // epilogue warps of both CTAs
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_pair(tmem_base_slot, kTmemCols);
    tc_fence_before();
    __syncwarp();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;
    griddep_wait();  // operands, residuals and outputs belong to the preceding kernels up to here
    if (threadIdx.x == 0) VITAD_TL(2);

    if (warp == 0) {
        // TMA producer (both CTAs): this CTA's 128 rows of A and its half of the tile's weight rows
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
            const TileCoord t = decode_tile<BLOCK_N>(sched, tile);
            const int row0 = t.m_blk * kPairM + static_cast<int>(rank) * kBlockM;
            const int n_row0 = t.n0 + static_cast<int>(rank) * (t.w >> 1);
            const CUtensorMap* bmap = t.tail ? &tma_b_tail : &tma_b;
            const uint32_t tx = 2u * static_cast<uint32_t>(S::kABytes + (t.w >> 1) * (kBlockK * 2));
            for (int kb = 0; kb < num_kb; ++kb) {
                int a_k = kb * kBlockK, a_row = row0;
                if (sched.kpt > 0) {
                    const int tap = kb / sched.kpt;
                    int r = kb - tap * sched.kpt;
                    if (sched.split_kb > 0) {  // segments hi, lo, hi of this tap
                        const int seg = r / sched.split_kb;
                        r = r - seg * sched.split_kb + (seg == 1 ? sched.split_kb : 0);
                    }
                    a_k = tap * sched.tap_cols + r * kBlockK;
                    if (sched.conv_pitch > 0) a_row = row0 + (tap / 3 - 1) * sched.conv_pitch + (tap % 3 - 1);
                }
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (elect_one()) {
                    if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], tx);
                    tma_load_2d_pair(smem_a + stage * S::kABytes, &tma_a, &full_bar[stage], a_k, a_row);
                    tma_load_2d_pair(smem_b + stage * S::kBBytes, bmap, &full_bar[stage], kb * kBlockK, n_row0);
                    if (tile == cluster_id && kb == 0) VITAD_TL(3);
                    VITAD_TL(4);
                }
                __syncwarp();
                if (++stage == kStages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            // MMA issuer (leader CTA): one M=256 x N=w MMA per 16-element k-slice
            const uint32_t a_lo0 = smem_desc_lo(smem_u32(smem_a));
            const uint32_t b_lo0 = smem_desc_lo(smem_u32(smem_b));
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
                const TileCoord t = decode_tile<BLOCK_N>(sched, tile);
                const uint32_t idesc = make_idesc_f16(kPairM, t.w);
                const int acc = it & 1;
                mbar_wait_cluster(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                if (elect_one() && it < 12) VITAD_TL(8 + 3 * it);
                const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (kb == 0 && elect_one() && it < 12) VITAD_TL(9 + 3 * it);
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
                    if (it < 12) VITAD_TL(10 + 3 * it);
                }
                __syncwarp();
            }
        }
    } else {
        const int quarter = warp & 3;       // TMEM lane quarter this warp may access
        const int group = (warp - 2) >> 2;  // the warps of a quarter alternate over the tile's 32-column units
        const uint32_t slab = smem_u32(epi_slabs + (warp - 2) * S::kSlabBytes);
        int it = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
            const TileCoord t = decode_tile<BLOCK_N>(sched, tile);
            const int row_base = t.m_blk * kPairM + static_cast<int>(rank) * kBlockM + quarter * 32;
            RowCtx my_ctx{0, 0};
            if constexpr (Epi::kRowCtx) my_ctx = epi.row_ctx(row_base + lane);
            const int acc = it & 1;
            mbar_wait(&tmem_full[acc], (it >> 1) & 1);
            __syncwarp();
            tc_fence_after();
            if (threadIdx.x == 64 && it < 8) VITAD_TL(44 + 2 * it);
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BLOCK_N;
#pragma unroll 1
            for (int c = group * 32; c < t.w; c += 8 * EPI_WARPS)
                epilogue_unit(epi, row_base, lane, my_ctx, t.n0 + c, taddr + c, slab);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);
            if (threadIdx.x == 64 && it < 8) VITAD_TL(45 + 2 * it);
        }
    }

    tc_fence_before();
    __syncwarp();
    cluster_sync_all();
    if (threadIdx.x == 0) {
        VITAD_TL(60);
        VITAD_TLG(61);
    }
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------------ epilogues
// GELU(x) = x * Phi(x), Phi(x) = sigmoid(x * p(x^2)) with p the degree-4 minimax fit of logit(Phi(x)) / x on [-9, 9]
// (max |error| of GELU 3.0e-6, i.e. below half an fp16 ulp of every output >= 6e-3 in magnitude; erff() costs ~35
// instructions per element and made the fc1 epilogue issue-bound).  Coefficients are pre-multiplied by -log2(e).
// Outside [-9, 9] the odd polynomial keeps growing monotonically, so the sigmoid saturates to exactly 0 / 1 (and an
// overflowing x^8 gives exp2(-+inf)): GELU(x) = x or 0 there, no clamp needed.
__device__ __forceinline__ float gelu_sigmoid(float x) {
    const float xc = x;
    const float x2 = xc * xc;
    float p = fmaf(x2, -3.2289875e-06f, 8.8238030e-05f);
    p = fmaf(x2, p, 3.6027565e-04f);
    p = fmaf(x2, p, -1.0522669e-01f);
    p = fmaf(x2, p, -2.3020454e+00f);
    const float e = ex2f(p * xc);  // exp(-x p(x^2))
    return __fdividef(x, 1.0f + e);
}

struct NoPre {};

// out_f16[row][col] = act(acc + bias[col]);  ACT: 0 identity, 1 GELU (timm Mlp: nn.GELU()), 2 ReLU
template <int ACT>
struct SEpiBiasH {
    static constexpr bool kRowCtx = false;
    static constexpr int kDefaultEpiWarps = ACT != 0 ? 16 : 8;  // GELU: latency hiding; ReLU: the small-K subnet/decoder GEMMs are output-bound
    using Pre = NoPre;
    using ColC = float4;
    const float* bias;
    __half* out;
    int ldo, M, N;
    int split;  // 1: out rows are [hi (N) | lo (N)] (split-fp16 activations, see split_h4)
    __device__ __forceinline__ RowCtx row_ctx(int) const { return RowCtx{0, 0}; }
    __device__ __forceinline__ bool direct(int) const { return false; }
    __device__ __forceinline__ void direct_unit(int, RowCtx, int, const uint32_t (&)[32]) const {}
    __device__ __forceinline__ Pre prefetch(int, RowCtx, int) const { return NoPre{}; }
    __device__ __forceinline__ ColC col_const(int col) const {
        return col < N ? __ldg(reinterpret_cast<const float4*>(bias + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __device__ __forceinline__ static float act(float x) {
        return ACT == 1 ? gelu_sigmoid(x) : (ACT == 2 ? fmaxf(x, 0.f) : x);
    }
    __device__ __forceinline__ void store(int row, RowCtx, int col, float4 a, Pre, ColC b) const {
        if (row < M && col < N) {
            __half* dst = out + static_cast<size_t>(row) * ldo + col;
            uint2 u, lo;
            if (split) {
                split_h4(act(a.x + b.x), act(a.y + b.y), act(a.z + b.z), act(a.w + b.w), u, lo);
                *reinterpret_cast<uint2*>(dst + N) = lo;
            } else {
                u.x = pack_h2(act(a.x + b.x), act(a.y + b.y));
                u.y = pack_h2(act(a.z + b.z), act(a.w + b.w));
            }
            *reinterpret_cast<uint2*>(dst) = u;
        }
    }
};

// Pixel-row maps between the plain NHWC layout [B, g, g] and the zero-bordered one [B, g+2, g+2] that the implicit 3x3
// convolution reads (TileSched::conv_pitch).  -1: the row has no image (out of range, or a border row).
__device__ __forceinline__ int row_plain_to_padded(int row, int g) {
    const int x = row % g;
    const int t = row / g;
    const int y = t % g;
    const int b = t / g;
    const int P = g + 2;
    return (b * P + y + 1) * P + x + 1;
}
__device__ __forceinline__ int row_padded_to_plain(int row, int g) {
    const int P = g + 2;
    const int xp = row % P;
    const int t = row / P;
    const int yp = t % P;
    const int b = t / P;
    if (xp < 1 || xp > g || yp < 1 || yp > g) return -1;
    return (b * g + yp - 1) * g + xp - 1;
}

// SEpiBiasH with a pixel-row map: mode 1 writes plain GEMM rows into the zero-bordered layout (the producer of an
// implicit 3x3 convolution), mode 2 writes the interior rows of a zero-bordered GEMM (the convolution itself) as plain rows.
template <int ACT>
struct SEpiBiasHMap {
    static constexpr bool kRowCtx = true;
    static constexpr int kDefaultEpiWarps = 16;  // output-bound (small K): measured 0.995 -> 0.886 ms on the reverse-ResNet decoder
    using Pre = NoPre;
    using ColC = float4;
    const float* bias;
    __half* out;
    int ldo, M, N, mode, g;
    int split;  // 1: out rows are [hi (N) | lo (N)]
    __device__ __forceinline__ RowCtx row_ctx(int row) const {
        if (row >= M) return RowCtx{0, -1};
        const int r = mode == 1 ? row_plain_to_padded(row, g) : row_padded_to_plain(row, g);
        return RowCtx{r, r < 0 ? -1 : 0};
    }
    __device__ __forceinline__ bool direct(int) const { return false; }
    __device__ __forceinline__ void direct_unit(int, RowCtx, int, const uint32_t (&)[32]) const {}
    __device__ __forceinline__ Pre prefetch(int, RowCtx, int) const { return NoPre{}; }
    __device__ __forceinline__ ColC col_const(int col) const {
        return col < N ? __ldg(reinterpret_cast<const float4*>(bias + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __device__ __forceinline__ void store(int, RowCtx ctx, int col, float4 a, Pre, ColC b) const {
        if (ctx.b < 0 || col >= N) return;
        __half* dst = out + static_cast<size_t>(ctx.a) * ldo + col;
        uint2 u, lo;
        if (split) {
            split_h4(SEpiBiasH<ACT>::act(a.x + b.x), SEpiBiasH<ACT>::act(a.y + b.y), SEpiBiasH<ACT>::act(a.z + b.z),
                     SEpiBiasH<ACT>::act(a.w + b.w), u, lo);
            *reinterpret_cast<uint2*>(dst + N) = lo;
        } else {
            u.x = pack_h2(SEpiBiasH<ACT>::act(a.x + b.x), SEpiBiasH<ACT>::act(a.y + b.y));
            u.y = pack_h2(SEpiBiasH<ACT>::act(a.z + b.z), SEpiBiasH<ACT>::act(a.w + b.w));
        }
        *reinterpret_cast<uint2*>(dst) = u;
    }
};

// out_f32[row][col] = resid_f32[row][col] + acc + bias[col]   (residual stream stays fp32; out may alias resid:
// every element is read and written by the same thread)
struct SEpiResidualF32 {
    static constexpr bool kRowCtx = false;
    static constexpr int kDefaultEpiWarps = 8;
    using Pre = float4;
    using ColC = float4;
    const float* bias;
    const float* resid;
    float* out;
    int ld, M, N;
    __device__ __forceinline__ RowCtx row_ctx(int) const { return RowCtx{0, 0}; }
    __device__ __forceinline__ bool direct(int) const { return false; }
    __device__ __forceinline__ void direct_unit(int, RowCtx, int, const uint32_t (&)[32]) const {}
    __device__ __forceinline__ Pre prefetch(int row, RowCtx, int col) const {
        return (row < M && col < N) ? *reinterpret_cast<const float4*>(resid + static_cast<size_t>(row) * ld + col)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __device__ __forceinline__ ColC col_const(int col) const {
        return col < N ? __ldg(reinterpret_cast<const float4*>(bias + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __device__ __forceinline__ void store(int row, RowCtx, int col, float4 a, Pre r, ColC b) const {
        if (row < M && col < N) {
            float4 o;
            o.x = r.x + (a.x + b.x);
            o.y = r.y + (a.y + b.y);
            o.z = r.z + (a.z + b.z);
            o.w = r.w + (a.w + b.w);
            *reinterpret_cast<float4*>(out + static_cast<size_t>(row) * ld + col) = o;
        }
    }
};

// Fused QKV projection epilogue (timm Attention.qkv / Swin WindowAttention.qkv + reshape/permute; for Swin the
// cyclic shift + window_partition of SwinTransformerModule.py:360-384 is folded into the store address):
//   row = b*L + t, col = which*C + h*hd + e;  (window, pos) = tok2win[t] (identity when tok2win is null)
//   q [bw][h][pos][e] = (acc+bias) * scale      k [bw][h][pos][e] = acc+bias
//   vt[bw][h][e][pos] =  acc+bias   (V transposed, position contiguous, padded to Tpad: K-major B operand of P@V)
struct SEpiQkv {
    static constexpr bool kRowCtx = true;
    static constexpr int kDefaultEpiWarps = 8;
    using Pre = NoPre;
    struct ColC {
        float4 bias;
        int which;
        int off;  // h*T*hd + e for q/k
    };
    const float* bias;
    __half* q;
    __half* k;
    __half* vt;
    const int* tok2win;
    int M, L, T, Tpad, H, hd, nW;
    float scale;
    int v_nat;  // 1: V is stored like K ([bw][h][pos][e], into `vt`) for the MN-major P.V operand of the attention kernel
    // ctx.a = bw*H*T + pos  (q/k row index before the head offset);  ctx.b = bw (-1: row out of range)
    __device__ __forceinline__ RowCtx row_ctx(int row) const {
        if (row >= M) return RowCtx{0, -1};
        const int b = row / L;
        const int t = row - b * L;
        int widx = 0, pos = t;
        if (tok2win != nullptr) {
            const int wp = __ldg(tok2win + t);
            widx = wp / T;
            pos = wp - widx * T;
        }
        const int bw = b * nW + widx;
        return RowCtx{bw * H * T + pos, bw};
    }
    __device__ __forceinline__ bool direct(int col0) const { return col0 >= 2 * H * hd; }
    __device__ __forceinline__ void direct_unit(int, RowCtx ctx, int col0, const uint32_t (&acc)[32]) const {
        const int C = H * hd;
        if (ctx.b < 0 || col0 >= 3 * C) return;
        const int cc = col0 - 2 * C;  // = h*hd + e0; a 32-column unit never straddles a head (hd = 32 or 64)
        if (v_nat) {
            // natural layout [bw][h][pos][e]: this thread's 32 columns are 64 contiguous bytes of its own row (measured:
            // 4 x 16-byte stores per thread beat both the staged path and the 32 two-byte stores of the transposed form)
            const int hh = cc / hd;
            __half* dst = vt + (static_cast<size_t>(ctx.a) + static_cast<size_t>(hh) * T) * hd + (cc - hh * hd);
            const float4* bp = reinterpret_cast<const float4*>(bias + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 b0 = __ldg(bp + 2 * j), b1 = __ldg(bp + 2 * j + 1);
                uint4 u;
                u.x = pack_h2(__uint_as_float(acc[8 * j + 0]) + b0.x, __uint_as_float(acc[8 * j + 1]) + b0.y);
                u.y = pack_h2(__uint_as_float(acc[8 * j + 2]) + b0.z, __uint_as_float(acc[8 * j + 3]) + b0.w);
                u.z = pack_h2(__uint_as_float(acc[8 * j + 4]) + b1.x, __uint_as_float(acc[8 * j + 5]) + b1.y);
                u.w = pack_h2(__uint_as_float(acc[8 * j + 6]) + b1.z, __uint_as_float(acc[8 * j + 7]) + b1.w);
                *reinterpret_cast<uint4*>(dst + 8 * j) = u;
            }
            return;
        }
        const int pos = ctx.a - ctx.b * H * T;
        __half* dst = vt + (static_cast<size_t>(ctx.b) * C + cc) * Tpad + pos;
        const float4* bp = reinterpret_cast<const float4*>(bias + col0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 bb = __ldg(bp + j);
            dst[static_cast<size_t>(4 * j + 0) * Tpad] = to_h(__uint_as_float(acc[4 * j + 0]) + bb.x);
            dst[static_cast<size_t>(4 * j + 1) * Tpad] = to_h(__uint_as_float(acc[4 * j + 1]) + bb.y);
            dst[static_cast<size_t>(4 * j + 2) * Tpad] = to_h(__uint_as_float(acc[4 * j + 2]) + bb.z);
            dst[static_cast<size_t>(4 * j + 3) * Tpad] = to_h(__uint_as_float(acc[4 * j + 3]) + bb.w);
        }
    }
    __device__ __forceinline__ Pre prefetch(int, RowCtx, int) const { return NoPre{}; }
    __device__ __forceinline__ ColC col_const(int col) const {
        ColC c;
        const int C = H * hd;
        if (col >= 2 * C) {
            c.bias = make_float4(0.f, 0.f, 0.f, 0.f);
            c.which = 2;  // V leaves through direct_unit
            c.off = 0;
            return c;
        }
        c.bias = __ldg(reinterpret_cast<const float4*>(bias + col));
        c.which = col >= C ? 1 : 0;
        const int cc = col - c.which * C;
        const int h = cc / hd;
        c.off = h * T * hd + (cc - h * hd);
        return c;
    }
    __device__ __forceinline__ void store(int, RowCtx ctx, int, float4 a, Pre, ColC c) const {
        if (ctx.b < 0 || c.which == 2) return;
        const float s = c.which == 0 ? scale : 1.0f;
        uint2 u;
        u.x = pack_h2((a.x + c.bias.x) * s, (a.y + c.bias.y) * s);
        u.y = pack_h2((a.z + c.bias.z) * s, (a.w + c.bias.w) * s);
        __half* base = c.which == 0 ? q : k;
        // (ctx.a - pos) * hd + h*T*hd + pos*hd + e  ==  (bw*H*T)*hd + ... : ctx.a already holds bw*H*T + pos
        *reinterpret_cast<uint2*>(base + static_cast<size_t>(ctx.a) * hd + c.off) = u;
    }
};

// Patch-embed epilogue (timm PatchEmbed conv-as-GEMM + pos_embed add, prefix tokens skipped):
//   row = b*P + p  ->  x[b][prefix + p][col] = acc + bias[col] + pos[prefix + p][col]
struct SEpiPatchEmbed {
    static constexpr bool kRowCtx = true;
    static constexpr int kDefaultEpiWarps = 8;
    using Pre = float4;
    using ColC = float4;
    const float* bias;
    const float* pos;
    float* out;
    int M, P, prefix, C;
    // ctx.a = output row (b*(prefix+P) + prefix + p), ctx.b = pos row (prefix + p), -1 when out of range
    __device__ __forceinline__ RowCtx row_ctx(int row) const {
        if (row >= M) return RowCtx{0, -1};
        const int b = row / P;
        const int p = row - b * P;
        return RowCtx{b * (prefix + P) + prefix + p, prefix + p};
    }
    __device__ __forceinline__ bool direct(int) const { return false; }
    __device__ __forceinline__ void direct_unit(int, RowCtx, int, const uint32_t (&)[32]) const {}
    __device__ __forceinline__ Pre prefetch(int, RowCtx ctx, int col) const {
        return (ctx.b >= 0 && col < C) ? __ldg(reinterpret_cast<const float4*>(pos + static_cast<size_t>(ctx.b) * C + col))
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __device__ __forceinline__ ColC col_const(int col) const {
        return col < C ? __ldg(reinterpret_cast<const float4*>(bias + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __device__ __forceinline__ void store(int, RowCtx ctx, int col, float4 a, Pre r, ColC b) const {
        if (ctx.b >= 0 && col < C) {
            float4 o;
            o.x = r.x + (a.x + b.x);
            o.y = r.y + (a.y + b.y);
            o.z = r.z + (a.z + b.z);
            o.w = r.w + (a.w + b.w);
            *reinterpret_cast<float4*>(out + static_cast<size_t>(ctx.a) * C + col) = o;
        }
    }
};

// ConvTranspose2d(kernel 3, stride 2, padding 1, output_padding 1) + folded BatchNorm + ReLU as ONE GEMM
// (CnnDecoder.py:47-117): GEMM row r = input pixel (b, i, j) of a Wg x Wg grid, K = the 2x2 input neighbourhood x C_in
// (im2col2x2), GEMM column = a*(2*Cp) + c*Cp + co for output phase (a, c) — the output pixel (2i+a, 2j+c).  The two
// c-phases of a row are adjacent NHWC pixels, so with output "rows" (b, y, j) of 2*Cp channels:
//   out_row = 2r - (r mod Wg) + a*Wg,   out_col = col mod (2*Cp).
struct SEpiConvT {
    static constexpr bool kRowCtx = true;
    static constexpr int kDefaultEpiWarps = 8;
    using Pre = NoPre;
    using ColC = float4;
    const float* bias;  // [N] = per (phase, channel), BN folded
    __half* out;        // NHWC fp16 [B, 2Wg, 2Wg, Cp]
    int M, N, Wg;       // N = 4*Cp
    int split;          // 1: every output pixel holds [hi (Cp) | lo (Cp)] channels
    __device__ __forceinline__ RowCtx row_ctx(int row) const {
        if (row >= M) return RowCtx{0, -1};
        return RowCtx{2 * row - (row % Wg), 0};
    }
    __device__ __forceinline__ bool direct(int) const { return false; }
    __device__ __forceinline__ void direct_unit(int, RowCtx, int, const uint32_t (&)[32]) const {}
    __device__ __forceinline__ Pre prefetch(int, RowCtx, int) const { return NoPre{}; }
    __device__ __forceinline__ ColC col_const(int col) const {
        return col < N ? __ldg(reinterpret_cast<const float4*>(bias + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __device__ __forceinline__ void store(int, RowCtx ctx, int col, float4 a, Pre, ColC b) const {
        if (ctx.b < 0 || col >= N) return;
        const int half_n = N >> 1;  // 2*Cp
        const int pa = col >= half_n ? 1 : 0;
        const int cc = col - pa * half_n;
        if (split) {
            // output "row" (b, y, j) = two adjacent pixels of 2*Cp channels each ([hi | lo]); cc = c*Cp + co
            const int Cp = N >> 2;
            const int pc = cc >= Cp ? 1 : 0;
            __half* dst = out + (static_cast<size_t>(ctx.a) + pa * Wg) * (2 * half_n) + pc * (2 * Cp) + (cc - pc * Cp);
            uint2 u, lo;
            split_h4(fmaxf(a.x + b.x, 0.f), fmaxf(a.y + b.y, 0.f), fmaxf(a.z + b.z, 0.f), fmaxf(a.w + b.w, 0.f), u, lo);
            *reinterpret_cast<uint2*>(dst) = u;
            *reinterpret_cast<uint2*>(dst + Cp) = lo;
            return;
        }
        uint2 u;
        u.x = pack_h2(fmaxf(a.x + b.x, 0.f), fmaxf(a.y + b.y, 0.f));
        u.y = pack_h2(fmaxf(a.z + b.z, 0.f), fmaxf(a.w + b.w, 0.f));
        *reinterpret_cast<uint2*>(out + (static_cast<size_t>(ctx.a) + pa * Wg) * half_n + cc) = u;
    }
};

// Tail of a reverse-ResNet Bottleneck (ReverseResNet.py:86-103): out_f16 = relu(acc + bias + identity).
// rg == 0: identity row = output row.  rg = g > 0: the identity path was computed on the g x g input grid (stride-2 1x1
// transposed convolution, :190-195) and lands on the even pixels of the 2g x 2g output grid; the other pixels only see
// its BatchNorm shift, which the host folds into `bias`.
struct SEpiResReluH {
    static constexpr bool kRowCtx = true;
    static constexpr int kDefaultEpiWarps = 16;  // output-bound (small K): measured 0.995 -> 0.886 ms on the reverse-ResNet decoder
    using Pre = uint4;  // residual hi quad (.x, .y) and, in split mode, lo quad (.z, .w)
    using ColC = float4;
    const float* bias;
    const __half* resid;
    __half* out;
    int ldo, ldr, M, N, rg;
    int pad_g;  // > 0: output rows go to the zero-bordered layout [B, pad_g+2, pad_g+2] (feeds an implicit 3x3 convolution)
    int split;  // 1: residual and output rows are [hi (N) | lo (N)]: the identity stream keeps ~22 mantissa bits
    // ctx.a = residual row (-1: none), ctx.b = output row (-1 when out of range)
    __device__ __forceinline__ RowCtx row_ctx(int row) const {
        if (row >= M) return RowCtx{-1, -1};
        const int orow = pad_g > 0 ? row_plain_to_padded(row, pad_g) : row;
        if (rg == 0) return RowCtx{row, orow};
        const int G = 2 * rg;
        const int x = row % G;
        const int t = row / G;
        const int y = t % G;
        const int b = t / G;
        if ((x | y) & 1) return RowCtx{-1, orow};
        return RowCtx{(b * rg + (y >> 1)) * rg + (x >> 1), orow};
    }
    __device__ __forceinline__ bool direct(int) const { return false; }
    __device__ __forceinline__ void direct_unit(int, RowCtx, int, const uint32_t (&)[32]) const {}
    __device__ __forceinline__ Pre prefetch(int, RowCtx ctx, int col) const {
        uint4 r = make_uint4(0u, 0u, 0u, 0u);
        if (ctx.a >= 0 && col < N) {
            const __half* src = resid + static_cast<size_t>(ctx.a) * ldr + col;
            const uint2 hi = *reinterpret_cast<const uint2*>(src);
            r.x = hi.x, r.y = hi.y;
            if (split) {
                const uint2 lo = *reinterpret_cast<const uint2*>(src + N);
                r.z = lo.x, r.w = lo.y;
            }
        }
        return r;
    }
    __device__ __forceinline__ ColC col_const(int col) const {
        return col < N ? __ldg(reinterpret_cast<const float4*>(bias + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __device__ __forceinline__ void store(int, RowCtx ctx, int col, float4 a, Pre r, ColC b) const {
        if (ctx.b < 0 || col >= N) return;
        float2 r0 = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
        float2 r1 = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
        __half* dst = out + static_cast<size_t>(ctx.b) * ldo + col;
        uint2 u, lo;
        if (split) {
            const float2 l0 = __half22float2(*reinterpret_cast<const __half2*>(&r.z));
            const float2 l1 = __half22float2(*reinterpret_cast<const __half2*>(&r.w));
            r0.x += l0.x, r0.y += l0.y, r1.x += l1.x, r1.y += l1.y;
            split_h4(fmaxf(a.x + b.x + r0.x, 0.f), fmaxf(a.y + b.y + r0.y, 0.f), fmaxf(a.z + b.z + r1.x, 0.f),
                     fmaxf(a.w + b.w + r1.y, 0.f), u, lo);
            *reinterpret_cast<uint2*>(dst + N) = lo;
        } else {
            u.x = pack_h2(fmaxf(a.x + b.x + r0.x, 0.f), fmaxf(a.y + b.y + r0.y, 0.f));
            u.y = pack_h2(fmaxf(a.z + b.z + r1.x, 0.f), fmaxf(a.w + b.w + r1.y, 0.f));
        }
        *reinterpret_cast<uint2*>(dst) = u;
    }
};

// Image head of the reverse-ResNet decoder (CnnDecoder.py:189-194: nearest upsample 56 -> 112, ConvTranspose2d(64 -> 3,
// k7, s2, p3, op1), BatchNorm2d, Tanh) after the host collapsed it to a 3x3 convolution on the Wg x Wg grid: GEMM row =
// pixel (b, J, I), column = c*16 + py*4 + px; out[b][c][4J+py][4I+px] = tanh(acc + bias).  A thread's 4 columns are the
// 4 px of one (c, py): one 16-byte store.
struct SEpiTanhPix4 {
    static constexpr bool kRowCtx = true;
    static constexpr int kDefaultEpiWarps = 16;  // output-bound (small K): measured 0.995 -> 0.886 ms on the reverse-ResNet decoder
    using Pre = NoPre;
    using ColC = float4;
    const float* bias;  // [64], 48 live
    float* out;         // fp32 NCHW [B, 3, 4Wg, 4Wg]
    int M, Wg;
    int padded;         // GEMM rows are the zero-bordered pixels of an implicit convolution
    // ctx.a = image b, ctx.b = J*Wg + I (-1: out of range or a border row)
    __device__ __forceinline__ RowCtx row_ctx(int row) const {
        if (row >= M) return RowCtx{0, -1};
        if (padded) {
            row = row_padded_to_plain(row, Wg);
            if (row < 0) return RowCtx{0, -1};
        }
        const int pix = Wg * Wg;
        const int b = row / pix;
        return RowCtx{b, row - b * pix};
    }
    __device__ __forceinline__ bool direct(int) const { return false; }
    __device__ __forceinline__ void direct_unit(int, RowCtx, int, const uint32_t (&)[32]) const {}
    __device__ __forceinline__ Pre prefetch(int, RowCtx, int) const { return NoPre{}; }
    __device__ __forceinline__ ColC col_const(int col) const {
        return col < 48 ? __ldg(reinterpret_cast<const float4*>(bias + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __device__ __forceinline__ void store(int, RowCtx ctx, int col, float4 a, Pre, ColC b) const {
        if (ctx.b < 0 || col >= 48) return;
        const int c = col >> 4, py = (col >> 2) & 3;
        const int J = ctx.b / Wg, I = ctx.b - J * Wg;
        const int S = 4 * Wg;
        float4 o;
        o.x = tanhf(a.x + b.x);
        o.y = tanhf(a.y + b.y);
        o.z = tanhf(a.z + b.z);
        o.w = tanhf(a.w + b.w);
        *reinterpret_cast<float4*>(out + ((static_cast<size_t>(ctx.a) * 3 + c) * S + 4 * J + py) * S + 4 * I) = o;
    }
};

// Plain fp32 output (+ optional bias), any N and pitch: used by tests and small projections.
struct SEpiBiasF32 {
    static constexpr bool kRowCtx = false;
    static constexpr int kDefaultEpiWarps = 8;
    using Pre = NoPre;
    using ColC = float4;
    const float* bias;  // may be null
    float* out;
    int ldo, M, N;
    __device__ __forceinline__ RowCtx row_ctx(int) const { return RowCtx{0, 0}; }
    __device__ __forceinline__ bool direct(int) const { return false; }
    __device__ __forceinline__ void direct_unit(int, RowCtx, int, const uint32_t (&)[32]) const {}
    __device__ __forceinline__ Pre prefetch(int, RowCtx, int) const { return NoPre{}; }
    __device__ __forceinline__ ColC col_const(int col) const {
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias != nullptr) {
            if (col + 0 < N) b.x = __ldg(bias + col + 0);
            if (col + 1 < N) b.y = __ldg(bias + col + 1);
            if (col + 2 < N) b.z = __ldg(bias + col + 2);
            if (col + 3 < N) b.w = __ldg(bias + col + 3);
        }
        return b;
    }
    __device__ __forceinline__ void store(int row, RowCtx, int col, float4 a, Pre, ColC b) const {
        if (row < M) {
            float* o = out + static_cast<size_t>(row) * ldo + col;
            if (col + 0 < N) o[0] = a.x + b.x;
            if (col + 1 < N) o[1] = a.y + b.y;
            if (col + 2 < N) o[2] = a.z + b.z;
            if (col + 3 < N) o[3] = a.w + b.w;
        }
    }
};

}  // namespace vitad
