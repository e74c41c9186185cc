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

#include "host_util.cuh"
#include "ptx.cuh"

namespace vitad {
extern std::atomic<uint64_t> g_launches;

constexpr float kLog2e = 1.4426950408889634f;

template <int TKP>
struct AttnSmem {
    static constexpr int kKeyBlocks = (TKP + 63) / 64;
    static constexpr int kPBytes = kKeyBlocks * 128 * 128;  // P: 128 rows x 64 keys x fp16 per block
    static constexpr int kQBytes = 128 * 128;
    static constexpr int kKBytes = TKP * 128;
    static constexpr int kVtBlockBytes = 64 * 128;  // 64 rows x 64 keys
    static constexpr int kVtBytes = kKeyBlocks * kVtBlockBytes;
    static constexpr int kRegion0 = kPBytes > kQBytes + kKBytes ? kPBytes : kQBytes + kKBytes;
    static constexpr int kTotal = kRegion0 + kVtBytes + 64 + 256 + 1024;  // + barriers + key region ids
};

struct AttnParams {
    __half* out;
    const float* bias;          // [H][T key][T query] or null
    const signed char* region;  // [nW][T] or null
    const int* win2tok;         // [nW*T] or null
    int T, H, hd, nW, L;        // window tokens, heads, head dim, windows per image, tokens per image
};

template <int TKP>
__global__ void __launch_bounds__(128, 2)
attention_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                 const __grid_constant__ CUtensorMap tma_vt, const AttnParams p) {
    using S = AttnSmem<TKP>;
    static_assert(TKP % 16 == 0 && TKP <= 256, "key padding must be a legal UMMA N");
    constexpr uint32_t kTmemCols = 256;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t* sP = smem;  // overlays sQ|sK once S has been produced
    uint8_t* sQ = smem;
    uint8_t* sK = smem + S::kQBytes;
    uint8_t* sVt = smem + S::kRegion0;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sVt + S::kVtBytes);  // qk, v, s, o
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
    signed char* s_region = reinterpret_cast<signed char*>(bars + 8);  // key region ids of this window

    const int warp = threadIdx.x >> 5;
    const int bh = blockIdx.y;  // (batch*window, head)
    const int m0 = blockIdx.x * 128;
    const int T = p.T;
    const int bw = bh / p.H, h = bh - bw * p.H;
    const int b = bw / p.nW, widx = bw - b * p.nW;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tma_q);
        tma_prefetch_desc(&tma_k);
        tma_prefetch_desc(&tma_vt);
        for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
        fence_barrier_init();
    }
    if (p.region != nullptr)
        for (int i = threadIdx.x; i < T; i += 128) s_region[i] = p.region[widx * T + i];
    if (warp == 0) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bars[0], S::kQBytes + S::kKBytes);
        tma_load_3d(sQ, &tma_q, &bars[0], 0, m0, bh);  // rows >= T and columns >= hd: zero-filled
        tma_load_3d(sK, &tma_k, &bars[0], 0, 0, bh);
        mbar_arrive_expect_tx(&bars[1], S::kVtBytes);
        for (int kb = 0; kb < S::kKeyBlocks; ++kb)
            tma_load_2d(sVt + kb * S::kVtBlockBytes, &tma_vt, &bars[1], kb * 64, bh * p.hd);
        mbar_wait(&bars[0], 0);
        tc_fence_after();
        constexpr uint32_t idesc_s = make_idesc_f16(128, TKP);
        const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            umma_f16_ss(tmem, make_smem_desc_sw128(qa + k * 32), make_smem_desc_sw128(ka + k * 32), idesc_s, k != 0);
        umma_commit(&bars[2]);
    }

    // ---- softmax: thread r owns query row m0 + r = TMEM lane r
    mbar_wait(&bars[2], 0);
    __syncwarp();
    tc_fence_after();
    const int r = threadIdx.x;
    const int tq = m0 + r;
    const bool qvalid = tq < T;
    // bias is stored key-major ([H][key][query]) so that the 32 query lanes of a warp read contiguous memory
    const float* brow = (p.bias != nullptr && qvalid) ? p.bias + static_cast<size_t>(h) * T * T + tq : nullptr;
    const int qreg = (p.region != nullptr && qvalid) ? s_region[tq] : -1;
    const uint32_t trow = tmem + (static_cast<uint32_t>(warp * 32) << 16);

    // scores of one 16-key chunk: accumulator + bias (16 independent loads issued before the TMEM wait so their
    // latency overlaps; a load per element inside the compare chain serialised ~400 L2 round trips per thread)
    // + region mask.  Key indices are clamped so the loads stay in bounds; invalid keys are masked by the caller.
    auto load_chunk = [&](int c, float (&sc)[16]) {
        uint32_t v[16];
        tmem_ld_x16(trow + c, v);
        float bv[16];
        if (brow != nullptr) {
#pragma unroll
            for (int j = 0; j < 16; ++j) bv[j] = __ldg(brow + static_cast<size_t>(min(c + j, T - 1)) * T);
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) bv[j] = 0.f;
        }
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            float s = __uint_as_float(v[j]) + bv[j];
            if (qreg >= 0 && s_region[c + j] != qreg) s -= 100.0f;
            sc[j] = s;
        }
    };

    float mx = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < TKP; c += 16) {
        float sc[16];
        load_chunk(c, sc);
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (c + j < T) mx = fmaxf(mx, sc[j]);
    }
    const float mxl = mx * kLog2e;
    float sum = 0.f;
#pragma unroll 1
    for (int c = 0; c < TKP; c += 16) {
        float sc[16];
        load_chunk(c, sc);
        float pv[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            pv[j] = (c + j < T) ? ex2f(fmaf(sc[j], kLog2e, -mxl)) : 0.f;
            sum += pv[j];
        }
        // two 16-byte chunks (8 keys each) of row r in key block c/64, 128B-swizzled
        const int blk = c >> 6;
        const int chunk0 = (c & 63) >> 3;
        uint8_t* rowp = sP + blk * (128 * 128) + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
            uint4 u;
            u.x = pack_h2(pv[8 * h2 + 0], pv[8 * h2 + 1]);
            u.y = pack_h2(pv[8 * h2 + 2], pv[8 * h2 + 3]);
            u.z = pack_h2(pv[8 * h2 + 4], pv[8 * h2 + 5]);
            u.w = pack_h2(pv[8 * h2 + 6], pv[8 * h2 + 7]);
            *reinterpret_cast<uint4*>(rowp + (((chunk0 + h2) ^ (r & 7)) << 4)) = u;
        }
    }
    // P (generic-proxy stores) must be visible to the tensor core (async proxy); S reads must be done.
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();

    if (threadIdx.x == 0) {
        tc_fence_after();
        mbar_wait(&bars[1], 0);
        tc_fence_after();
        constexpr uint32_t idesc_o = make_idesc_f16(128, 64);
        const uint32_t pa = smem_u32(sP), va = smem_u32(sVt);
#pragma unroll 1
        for (int kk = 0; kk < TKP / 16; ++kk) {
            const int blk = kk >> 2, w = kk & 3;
            umma_f16_ss(tmem, make_smem_desc_sw128(pa + blk * (128 * 128) + w * 32),
                        make_smem_desc_sw128(va + blk * S::kVtBlockBytes + w * 32), idesc_o, kk != 0);
        }
        umma_commit(&bars[3]);
    }

    mbar_wait(&bars[3], 0);
    __syncwarp();
    tc_fence_after();
    const float inv = 1.0f / sum;
    const int tok = qvalid ? (p.win2tok != nullptr ? __ldg(p.win2tok + widx * T + tq) : tq) : 0;
    __half* orow = p.out + (static_cast<size_t>(b) * p.L + tok) * (p.H * p.hd) + h * p.hd;
#pragma unroll
    for (int c = 0; c < 64; c += 32) {
        uint32_t v[32];
        tmem_ld_x32(trow + c, v);
        tmem_ld_wait();
        if (qvalid && c < p.hd) {
            uint4* o = reinterpret_cast<uint4*>(orow + c);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint4 u;
                u.x = pack_h2(__uint_as_float(v[8 * j + 0]) * inv, __uint_as_float(v[8 * j + 1]) * inv);
                u.y = pack_h2(__uint_as_float(v[8 * j + 2]) * inv, __uint_as_float(v[8 * j + 3]) * inv);
                u.z = pack_h2(__uint_as_float(v[8 * j + 4]) * inv, __uint_as_float(v[8 * j + 5]) * inv);
                u.w = pack_h2(__uint_as_float(v[8 * j + 6]) * inv, __uint_as_float(v[8 * j + 7]) * inv);
                o[j] = u;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, kTmemCols);
    }
}

}  // namespace vitad

using namespace vitad;

extern "C" int vitad_attention_f16(const vitad_attention_args* args, void* stream) {
    int rc = check_device_arch();
    if (rc) return rc;
    VITAD_REQUIRE(args, VITAD_ERR_ARG, "null args");
    const vitad_attention_args& a = *args;
    VITAD_REQUIRE(a.q && a.k && a.vt && a.out, VITAD_ERR_ARG, "null pointer");
    VITAD_REQUIRE(a.head_dim == 64 || a.head_dim == 32, VITAD_ERR_SHAPE, "head_dim %d unsupported (32, 64)", a.head_dim);
    VITAD_REQUIRE(a.tokens > 0 && a.tokens <= 208, VITAD_ERR_SHAPE, "tokens=%d unsupported (1..208)", a.tokens);
    VITAD_REQUIRE(a.tokens_pad >= 256 && a.tokens_pad % 8 == 0, VITAD_ERR_SHAPE,
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
    rc = make_tmap_f16_2d(&tv, a.vt, static_cast<uint64_t>(BH) * hd, a.tokens_pad, a.tokens_pad, 64);
    if (rc) return rc;
    auto kern = attention_kernel<TKP>;
    static bool attr_set = false;
    if (!attr_set) {
        VITAD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
        attr_set = true;
    }
    AttnParams p{static_cast<__half*>(a.out), a.bias, a.region, a.win2tok, a.tokens, a.heads, a.head_dim, nW,
                 nW * a.tokens};
    dim3 grid((a.tokens + 127) / 128, BH);
    char pname[64];
    snprintf(pname, sizeof(pname), "attention_t%d_h%d_hd%d_bw%d%s", a.tokens, a.heads, a.head_dim, a.batch_windows,
             a.region ? "_shift" : "");
    ProfScope prof(pname, static_cast<cudaStream_t>(stream));
    kern<<<grid, 128, S::kTotal, static_cast<cudaStream_t>(stream)>>>(tq, tk, tv, p);
    VITAD_CUDA_OK(cudaGetLastError());
    g_launches.fetch_add(1);
    return VITAD_OK;
}
