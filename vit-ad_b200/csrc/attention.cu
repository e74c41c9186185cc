// Fused multi-head self-attention for short sequences:
//   O = softmax(Q K^T + bias + mask) V      (Q is pre-scaled by hd^-0.5 in the QKV projection epilogue)
// DeiT (198 tokens, head dim 64): timm Attention.forward, called through
//   src/classes/transformer/TransformerEncoder.py:150-165;
// Swin / EsViT (windows of 196 or 49 tokens, head dim 32): WindowAttention.forward
//   (src/classes/transformer/SwinTransformerModule.py:144-193) with the relative-position bias, the shifted-window
//   mask of create_attn_mask (:316-347, value -100 between different regions) and window_reverse + the
//   reverse cyclic shift (:392-408) folded into the output row address.
// The reference materialises [B,H,T,T] scores in HBM (and rebuilds the shift mask on the CPU every forward).
//
// One CTA = one (window, head, 128-query tile); both GEMMs run on tcgen05 with accumulators in TMEM:
//   S[128 x TKP] = Q[128 x 64] . K[TKP x 64]^T      (TMA-loaded, 128B-swizzled operands; head dim 32 is
//                                                    zero-extended to 64 by the TMA box)
//   softmax over the T valid keys: each thread owns one query row (one TMEM lane), two passes over
//   TMEM (max, then exp/sum); P is written to shared memory as fp16 in the K-major swizzled layout
//   tcgen05 expects, overlaying the (dead) Q/K tiles
//   O[128 x 64] = P[128 x TKP] . Vt[64 x TKP]^T     (V is stored transposed by the QKV epilogue)
// Scores never touch HBM.  ~97 KB smem and 256 TMEM columns per CTA -> two CTAs per SM.
#include <atomic>

#include "gemm_core.cuh"
#include "host_util.cuh"
#include "ptx.cuh"

namespace vitad {
extern std::atomic<uint64_t> g_launches;

constexpr float kLog2e = 1.4426950408889634f;

template <int TKP>
struct AttnSmem {
    static constexpr int kKeyBlocks = (TKP + 63) / 64;
    static constexpr int kQBytes = 128 * 128;
    static constexpr int kKBytes = TKP * 128;
    static constexpr int kVtBlockBytes = 64 * 128;  // 64 rows x 64 keys
    static constexpr int kVtBytes = kKeyBlocks * kVtBlockBytes;
    static constexpr int kRegion0 = ((kQBytes + kKBytes + 1023) / 1024) * 1024;  // Q | K, later the output slabs
    static constexpr int kTotal = kRegion0 + kVtBytes + 64 + 256 + 3072 + 1024;  // + barriers + key region ids + max/ref/sum exchange
    static_assert(kRegion0 >= 8 * 2048, "output staging slabs overlay Q|K");
};

// TMEM columns of one CTA (256 allocated): S = Q.K^T in [0, TKP); P (fp16 pairs, 8 columns per 16 keys) is written
// in place over the S columns its own thread has already consumed for the first key half ([0, 56)) and into the unused
// tail [208, 256) for the second; O accumulates in [64, 128) (even key chunks) and [128, 192) (odd) once S is dead.
constexpr uint32_t kPCol1 = 208, kOCol = 64;

struct AttnParams {
    __half* out;
    const float* bias;          // [H][T key][T query] or null
    const signed char* region;  // [nW][T] or null
    const int* win2tok;         // [nW*T] or null
    int T, H, hd, nW, L;        // window tokens, heads, head dim, windows per image, tokens per image
    // Read only by the MMA-issuing warp, straight from the constant bank, so that its loop counters and operand
    // addresses stay in uniform registers (a value shared with per-thread code lives in a vector register and costs
    // four R2UR round trips, ~190 cycles, per issued MMA).
    int u_nch, u_h0;            // 16-key chunks with valid keys; chunks of the first key half
    uint32_t u_idesc_s;         // instruction descriptor of S = Q.K^T (M 128, N 16*u_nch)
};

constexpr int kAttnThreads = 256;

// 256 threads: warp w owns TMEM lane quarter (w & 3) = query rows 32*(w&3)..+31 of the tile and key half (w >> 2);
// two threads share a query row and exchange its running max / partial sum through shared memory.  Warps whose
// 32 rows are all beyond T (the second tile of 198 tokens has 70 valid rows, a 49-token window 49) skip the softmax.
// kBias / kMask specialise the relative-position bias and shifted-window mask away for DeiT: predicated-off
// instructions still issue, and the generic loop spent 20 instructions per score where 5 are needed.
// kVNat: V arrives in its natural layout [BW, H, T, hd] (hd = 64: rows of 128 bytes, the same tile K uses) and enters
// O = P.V as an MN-major B operand (instruction-descriptor bit 16; one 16-key slice = 16 rows = 2048 bytes), so the QKV
// projection writes V exactly like K — no transposed 2-byte scatter in its epilogue, no zero-padded vT buffer to clear.
template <int TKP, bool kBias, bool kMask, bool kVNat = false>
__global__ void __launch_bounds__(kAttnThreads, 2)
attention_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                 const __grid_constant__ CUtensorMap tma_vt, const AttnParams p) {
    using S = AttnSmem<TKP>;
    static_assert(TKP % 16 == 0 && TKP <= 256, "key padding must be a legal UMMA N");
    constexpr uint32_t kTmemCols = 256;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t* sQ = smem;
    uint8_t* sK = smem + S::kQBytes;
    uint8_t* sVt = smem + S::kRegion0;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sVt + S::kVtBytes);  // qk, v, s, o
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
    signed char* s_region = reinterpret_cast<signed char*>(bars + 8);  // key region ids of this window
    float* s_xch = reinterpret_cast<float*>(s_region + 256);           // 3 x [2 halves][128 rows]: first-group max, reference, sum

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int quarter = warp & 3, half = warp >> 2;
    const int bh = blockIdx.y;  // (batch*window, head)
    const int m0 = blockIdx.x * 128;
    const int T = p.T;
    const int bw = bh / p.H, h = bh - bw * p.H;
    const int b = bw / p.nW, widx = bw - b * p.nW;
    const int nch = (T + 15) >> 4;  // 16-key chunks that hold valid keys

    griddep_launch_dependents();
    griddep_wait();  // q/k/vt come from the preceding QKV projection; everything below may touch global memory
    // Warp 0 drives TMA and the tensor core warp-uniformly (one elected lane issues; a loop under `if (lane == 0)`
    // costs ~115 cycles per MMA in vector->uniform register moves, 1500 cycles for the 13 P.V MMAs).  The loads are
    // issued before the TMEM allocation so their L2 latency overlaps it.
    if (warp == 0) {
        if (elect_one()) {
            VITAD_TL(0);
            VITAD_TLG(1);
            tma_prefetch_desc(&tma_q);
            tma_prefetch_desc(&tma_k);
            tma_prefetch_desc(&tma_vt);
            for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
            fence_barrier_init();
            mbar_arrive_expect_tx(&bars[0], S::kQBytes + S::kKBytes);
            tma_load_3d(sQ, &tma_q, &bars[0], 0, m0, bh);  // rows >= T and columns >= hd: zero-filled
            tma_load_3d(sK, &tma_k, &bars[0], 0, 0, bh);
            if constexpr (kVNat) {
                mbar_arrive_expect_tx(&bars[1], S::kKBytes);
                tma_load_3d(sVt, &tma_vt, &bars[1], 0, 0, bh);  // [TKP keys][64]: keys >= T zero-filled
            } else {
                mbar_arrive_expect_tx(&bars[1], S::kVtBytes);
                for (int kb = 0; kb < S::kKeyBlocks; ++kb)
                    tma_load_2d(sVt + kb * S::kVtBlockBytes, &tma_vt, &bars[1], kb * 64, bh * p.hd);
            }
        }
        __syncwarp();
        tmem_alloc(tmem_slot, kTmemCols);
    }
    if constexpr (kMask)
        for (int i = threadIdx.x; i < T; i += kAttnThreads) s_region[i] = p.region[widx * T + i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) VITAD_TL(2);
        mbar_wait(&bars[0], 0);
        tc_fence_after();
        if (lane == 0) VITAD_TL(3);
        const uint32_t idesc_s = p.u_idesc_s;
        const uint32_t q_lo = smem_desc_lo(smem_u32(sQ)), k_lo = smem_desc_lo(smem_u32(sK));
        if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                umma_f16_ss(tmem, smem_desc_join(q_lo + 2 * k), smem_desc_join(k_lo + 2 * k), idesc_s, k != 0);
            umma_commit(&bars[2]);
        }
        __syncwarp();
    }

    // ---- softmax: threads (quarter, lane) of both halves own query row m0 + r = TMEM lane r
    const int r = quarter * 32 + lane;
    const int tq = m0 + r;
    const bool qvalid = tq < T;
    const bool warp_active = m0 + quarter * 32 < T;  // warp-uniform
    const int c_begin = half == 0 ? 0 : ((nch + 1) >> 1) * 16;
    const int c_end = half == 0 ? ((nch + 1) >> 1) * 16 : nch * 16;
    // bias is stored key-major ([H][key][query]) so that the 32 query lanes of a warp read contiguous memory
    const float* brow = (kBias && qvalid) ? p.bias + static_cast<size_t>(h) * T * T + tq : nullptr;
    const uint32_t trow = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
    float sum = 0.f;

    mbar_wait(&bars[2], 0);
    __syncwarp();
    tc_fence_after();
    if (threadIdx.x == 32) VITAD_TL(5);
    const int qreg = (kMask && qvalid) ? s_region[tq] : -1;

    // scores of two 16-key chunks (the second only when c + 16 < c_end): accumulator + bias (independent loads
    // issued before the single TMEM wait so their latency overlaps) + region mask.  Key indices are clamped so
    // the loads stay in bounds; keys >= T are masked by the caller.
    auto load_pair = [&](int c, bool two, float (&sc)[32]) {
        uint32_t v[32];
        tmem_ld_x16(trow + c, v);
        if (two) tmem_ld_x16(trow + c + 16, v + 16);
        float bv[32];
        if constexpr (kBias) {
            if (brow != nullptr) {
#pragma unroll
                for (int j = 0; j < 32; ++j) bv[j] = __ldg(brow + static_cast<size_t>(min(c + j, T - 1)) * T);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) bv[j] = 0.f;
            }
        }
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            float sv = __uint_as_float(v[j]);
            if constexpr (kBias) sv += bv[j];
            if constexpr (kMask) {
                if (qreg >= 0 && s_region[min(c + j, T - 1)] != qreg) sv -= 100.0f;
            }
            sc[j] = sv;
        }
    };
    // max over the valid keys of a pair; `full` (warp-uniform): both chunks present and all 32 keys < T
    auto pair_max = [&](int c, bool two, const float (&sc)[32]) {
        float gm = -INFINITY;
        if (two && c + 32 <= T) {
#pragma unroll
            for (int j = 0; j < 32; ++j) gm = fmaxf(gm, sc[j]);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if ((j < 16 || two) && c + j < T) gm = fmaxf(gm, sc[j]);
        }
        return gm;
    };

    // Single pass over the scores (TMEM reads, 64 B/clk per SM, bound this kernel: a max pass + an exp pass cost
    // 2 x 106 KB per tile).  The two threads of a row agree on a reference m = max over their first 32 keys each
    // (one exchange), then write P = 2^(s' - m') with the reference only raised when a later group exceeds it by
    // more than 2^8 (P <= 256 stays far inside fp16): the rare raise, and a final mismatch between the two halves,
    // rescale the thread's own P entries in shared memory.  Normalisation by the fp32 row sum happens on O.
    constexpr float kTau = 8.0f;
    float m2 = -INFINITY;  // reference in the log2 domain (m * log2 e)
    float sc0[32];
    const bool have0 = warp_active && c_begin < c_end;
    const bool two0 = c_begin + 16 < c_end;
    float g0 = -INFINITY;
    if (have0) {
        load_pair(c_begin, two0, sc0);
        g0 = pair_max(c_begin, two0, sc0);
    }
    if (warp_active) s_xch[half * 128 + r] = g0;
    __syncthreads();
    if (threadIdx.x == 32) VITAD_TL(6);

    // TMEM address of this thread's P entries for the 16-key chunk starting at key cc
    auto p_addr = [&](int cc) { return trow + (half == 0 ? (cc >> 1) : kPCol1 + ((cc - c_begin) >> 1)); };
    // multiply the warp's P entries of key chunks [c_begin, c_done) by the per-row factor f (warp-collective)
    auto rescale_p = [&](int c_done, float f) {
        const __half2 f2 = __float2half2_rn(f);
        for (int cc = c_begin; cc < c_done; cc += 16) {
            uint32_t u[8];
            tmem_ld_x8(p_addr(cc), u);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                __half2 hv = *reinterpret_cast<__half2*>(&u[e]);
                hv = __hmul2(hv, f2);
                u[e] = *reinterpret_cast<uint32_t*>(&hv);
            }
            tmem_st_x8(p_addr(cc), u);
        }
    };
    auto emit_group = [&](int c, bool two, float gmax, const float (&sc)[32]) {
        const float g2 = gmax * kLog2e;
        const bool raise = g2 > m2 + kTau;
        if (__any_sync(0xffffffffu, raise)) {  // rare: raise the reference, rescale what the warp has written so far
            const float f = raise ? ex2f(m2 - g2) : 1.0f;
            rescale_p(c, f);
            sum *= f;
            if (raise) m2 = g2;
        }
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            if (g == 1 && !two) break;
            const int cc = c + 16 * g;
            float pv[16];
            if (cc + 16 <= T) {  // warp-uniform: no per-key validity test
#pragma unroll
                for (int j = 0; j < 16; ++j) pv[j] = ex2f(fmaf(sc[16 * g + j], kLog2e, -m2));
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) pv[j] = (cc + j < T) ? ex2f(fmaf(sc[16 * g + j], kLog2e, -m2)) : 0.f;
            }
            float ps = 0.f;
#pragma unroll
            for (int j = 0; j < 16; j += 4) ps += (pv[j] + pv[j + 1]) + (pv[j + 2] + pv[j + 3]);
            sum += ps;
            uint32_t u[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) u[j] = pack_h2(pv[2 * j], pv[2 * j + 1]);
            tmem_st_x8(p_addr(cc), u);
        }
    };

    float* s_m2 = s_xch + 256;   // final references of both halves
    float* s_sum = s_xch + 512;  // partial row sums of both halves
    if (warp_active) {
        m2 = fmaxf(g0, s_xch[(half ^ 1) * 128 + r]) * kLog2e;  // key 0 is always valid: finite for valid rows
        if (!qvalid) m2 = 0.f;
        if (have0) emit_group(c_begin, two0, g0, sc0);
#pragma unroll 1
        for (int c = c_begin + 32; c < c_end; c += 32) {
            const bool two = c + 16 < c_end;
            float sc[32];
            load_pair(c, two, sc);
            emit_group(c, two, pair_max(c, two, sc), sc);
        }
        s_m2[half * 128 + r] = m2;
        s_sum[half * 128 + r] = sum;
    }
    __syncthreads();
    if (warp_active) {
        const float mp = s_m2[(half ^ 1) * 128 + r];
        const float sp = s_sum[(half ^ 1) * 128 + r];
        const float mm = fmaxf(m2, mp);
        if (__any_sync(0xffffffffu, mp > m2))  // rare: the other half raised its reference
            rescale_p(c_end, mp > m2 ? ex2f(m2 - mp) : 1.0f);
        sum = sum * ex2f(m2 - mm) + sp * ex2f(mp - mm);
        tmem_st_wait();
    }
    // P (tcgen05.st) complete and ordered before the MMA that reads it; all S reads done.
    if (threadIdx.x == 32) VITAD_TL(7);
    tc_fence_before();
    __syncthreads();

    if (warp == 0) {
        if (lane == 0) VITAD_TL(8);
        tc_fence_after();
        mbar_wait(&bars[1], 0);
        tc_fence_after();
        constexpr uint32_t idesc_o = make_idesc_f16(128, 64) | (kVNat ? (1u << 16) : 0u);  // bit 16: B is MN-major
        const uint32_t v_lo = smem_desc_lo(smem_u32(sVt));
        // fully unrolled inside one elected region: every operand is (uniform base + constant), so the
        // vector->uniform moves are hoisted and pipelined instead of serialising each MMA of a rolled loop
        const int h0 = p.u_h0, nch_u = p.u_nch;
        const uint32_t p1_base = tmem + kPCol1 - 8 * h0;  // second key half: chunk kk at p1_base + 8*kk
        if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < TKP / 16; ++kk) {
                if (kk < nch_u) {
                    const uint32_t a_tmem = kk < h0 ? tmem + kk * 8 : p1_base + kk * 8;  // A = P from TMEM
                    const uint32_t b_lo = kVNat ? v_lo + kk * (2048 >> 4)  // 16 key rows of 128 bytes
                                                : v_lo + (kk >> 2) * (S::kVtBlockBytes >> 4) + 2 * (kk & 3);
                    // two accumulators (even / odd key chunks): consecutive MMAs do not depend on each other
                    umma_f16_ts(tmem + kOCol + (kk & 1) * 64, a_tmem, smem_desc_join(b_lo), idesc_o, kk >= 2);
                }
            }
        }
        __syncwarp();
        if (elect_one()) umma_commit(&bars[3]);
        __syncwarp();
        if (lane == 0) VITAD_TL(13);
    }

    mbar_wait(&bars[3], 0);
    __syncwarp();
    tc_fence_after();
    if (threadIdx.x == 32) VITAD_TL(9);
    // warp (quarter, half) owns output columns [32*half, 32*half + 32) of its 32 rows.  Each thread holds one row
    // (64 bytes as fp16); the warp transposes through a private 2 KB slab of the dead Q|K region so that a store
    // instruction covers 8 rows x 64 contiguous bytes (full sectors) instead of 32 rows x 16 bytes.
    if (warp_active && half * 32 < p.hd) {
        const float inv = 1.0f / sum;
        const int tok = qvalid ? (kBias && p.win2tok != nullptr ? __ldg(p.win2tok + widx * T + tq) : tq) : -1;
        uint32_t v[32];
        if (threadIdx.x == 32) VITAD_TL(11);
        tmem_ld_x32(trow + kOCol + half * 32, v);
        if (nch >= 2) {  // second accumulator (odd key chunks)
            uint32_t v2[32];
            tmem_ld_x32(trow + kOCol + 64 + half * 32, v2);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
        }
        tmem_ld_wait();
        if (threadIdx.x == 32) VITAD_TL(12);
        const uint32_t slab = smem_u32(sQ) + warp * 2048;  // Q|K are dead since S was produced
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t x0 = pack_h2(__uint_as_float(v[8 * j + 0]) * inv, __uint_as_float(v[8 * j + 1]) * inv);
            const uint32_t x1 = pack_h2(__uint_as_float(v[8 * j + 2]) * inv, __uint_as_float(v[8 * j + 3]) * inv);
            const uint32_t x2 = pack_h2(__uint_as_float(v[8 * j + 4]) * inv, __uint_as_float(v[8 * j + 5]) * inv);
            const uint32_t x3 = pack_h2(__uint_as_float(v[8 * j + 6]) * inv, __uint_as_float(v[8 * j + 7]) * inv);
            // row = lane (64 bytes), 16-byte chunk j rotated by the row so the 8 lanes of a phase hit distinct banks
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(slab + lane * 64 + (((j + (lane >> 1)) & 3) << 4)),
                         "r"(x0), "r"(x1), "r"(x2), "r"(x3)
                         : "memory");
        }
        __syncwarp();
        const int sub = lane >> 2, ch = lane & 3;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int rr = 8 * i + sub;
            const int rtok = __shfl_sync(0xffffffffu, tok, rr);
            uint4 u;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
                         : "r"(slab + rr * 64 + (((ch + (rr >> 1)) & 3) << 4))
                         : "memory");
            if (rtok >= 0)
                *reinterpret_cast<uint4*>(p.out + (static_cast<size_t>(b) * p.L + rtok) * (p.H * p.hd) + h * p.hd +
                                          half * 32 + ch * 8) = u;
        }
    }
    if (threadIdx.x == 32) VITAD_TL(10);
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
        VITAD_TL(60);
        VITAD_TLG(61);
    }
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, kTmemCols);
    }
}

}  // namespace vitad

using namespace vitad;

#ifdef VITAD_TIMELINE
// Diagnostic builds only: timeline stamps of the attention kernel into a [grid.y * grid.x][64] uint64 device buffer.
extern "C" int vitad_debug_timeline_attention(void* device_buffer) {
    unsigned long long* p = static_cast<unsigned long long*>(device_buffer);
    VITAD_CUDA_OK(cudaMemcpyToSymbol(vitad::g_timeline, &p, sizeof(p)));
    return VITAD_OK;
}
#endif

extern "C" int vitad_attention_f16(const vitad_attention_args* args, void* stream) {
    VITAD_NVTX("vitad_attention_f16");
    int rc = check_device_arch();
    if (rc) return rc;
    VITAD_REQUIRE(args, VITAD_ERR_ARG, "null args");
    const vitad_attention_args& a = *args;
    VITAD_REQUIRE(a.q && a.k && (a.vt || a.v) && !(a.vt && a.v) && a.out, VITAD_ERR_ARG, "null pointer (exactly one of vt / v)");
    VITAD_REQUIRE(!a.v || (a.head_dim == 64 && !a.bias && !a.region), VITAD_ERR_SHAPE,
                  "natural-layout V: head_dim 64 without bias / region mask");
    VITAD_REQUIRE(a.head_dim == 64 || a.head_dim == 32, VITAD_ERR_SHAPE, "head_dim %d unsupported (32, 64)", a.head_dim);
    VITAD_REQUIRE(a.tokens > 0 && a.tokens <= 208, VITAD_ERR_SHAPE, "tokens=%d unsupported (1..208)", a.tokens);
    VITAD_REQUIRE(a.v || (a.tokens_pad >= 256 && a.tokens_pad % 8 == 0), VITAD_ERR_SHAPE,
                  "tokens_pad=%d must be >= 256 and a multiple of 8", a.tokens_pad);
    const int nW = a.windows > 0 ? a.windows : 1;
    VITAD_REQUIRE(a.batch_windows > 0 && a.batch_windows % nW == 0 && a.heads > 0, VITAD_ERR_SHAPE, "batch/windows");
    VITAD_REQUIRE(nW == 1 || a.win2tok, VITAD_ERR_ARG, "window attention needs the win2tok map");
    VITAD_REQUIRE(aligned16(a.out), VITAD_ERR_ALIGN, "output alignment");
    constexpr int TKP = 208;
    using S = AttnSmem<TKP>;
    const int BH = a.batch_windows * a.heads;
    const uint64_t hd = a.head_dim;
    CUtensorMap tq, tk, tv;
    rc = make_tmap_f16_3d(&tq, a.q, BH, a.tokens, hd, hd, static_cast<uint64_t>(a.tokens) * hd, 128);
    if (rc) return rc;
    rc = make_tmap_f16_3d(&tk, a.k, BH, a.tokens, hd, hd, static_cast<uint64_t>(a.tokens) * hd, TKP);
    if (rc) return rc;
    if (a.v)
        rc = make_tmap_f16_3d(&tv, a.v, BH, a.tokens, hd, hd, static_cast<uint64_t>(a.tokens) * hd, TKP);
    else
        rc = make_tmap_f16_2d(&tv, a.vt, static_cast<uint64_t>(BH) * hd, a.tokens_pad, a.tokens_pad, 64);
    if (rc) return rc;
    VITAD_REQUIRE(!a.region || a.bias, VITAD_ERR_ARG, "a region mask is only supported together with a bias");
    auto kern = a.region ? attention_kernel<TKP, true, true>
                         : (a.bias ? attention_kernel<TKP, true, false>
                                   : (a.v ? attention_kernel<TKP, false, false, true> : attention_kernel<TKP, false, false>));
    static bool attr_set = false;
    if (!attr_set) {
        VITAD_CUDA_OK(cudaFuncSetAttribute(attention_kernel<TKP, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
        VITAD_CUDA_OK(cudaFuncSetAttribute(attention_kernel<TKP, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
        VITAD_CUDA_OK(cudaFuncSetAttribute(attention_kernel<TKP, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
        VITAD_CUDA_OK(cudaFuncSetAttribute(attention_kernel<TKP, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
        attr_set = true;
    }
    const int nch = (a.tokens + 15) / 16;
    AttnParams p{static_cast<__half*>(a.out), a.bias, a.region, a.win2tok, a.tokens, a.heads, a.head_dim, nW,
                 nW * a.tokens, nch, (nch + 1) / 2, make_idesc_f16(128, nch * 16)};
    dim3 grid((a.tokens + 127) / 128, BH);
    char pname[64];
    snprintf(pname, sizeof(pname), "attention_t%d_h%d_hd%d_bw%d%s", a.tokens, a.heads, a.head_dim, a.batch_windows,
             a.region ? "_shift" : "");
    ProfScope prof(pname, static_cast<cudaStream_t>(stream));
    VITAD_CUDA_OK(launch_pdl(kern, grid, dim3(kAttnThreads), S::kTotal, static_cast<cudaStream_t>(stream), tq, tk, tv, p));
    VITAD_CUDA_OK(cudaGetLastError());
    g_launches.fetch_add(1);
    return VITAD_OK;
}
