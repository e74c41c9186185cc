// Fused multi-head self-attention for short sequences (DeiT: 198 tokens, head dim 64):
//   O = softmax(Q K^T) V      (Q is pre-scaled by hd^-0.5 in the QKV projection epilogue)
// replaces timm Attention.forward's two bmm + softmax, which materialise [B,H,T,T] in HBM
// (reference call site: src/classes/transformer/TransformerEncoder.py:150-165 via timm Block).
//
// One CTA = one (image, head, 128-query tile); both GEMMs run on tcgen05 with accumulators in TMEM:
//   S[128 x TKP] = Q[128 x 64] . K[TKP x 64]^T      (TMA-loaded, 128B-swizzled operands)
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
    static constexpr int kVtBlockBytes = 64 * 128;  // 64 (head dim) rows x 64 keys
    static constexpr int kVtBytes = kKeyBlocks * kVtBlockBytes;
    static constexpr int kRegion0 = kPBytes > kQBytes + kKBytes ? kPBytes : kQBytes + kKBytes;
    static constexpr int kTotal = kRegion0 + kVtBytes + 64 + 1024;
};

template <int TKP>
__global__ void __launch_bounds__(128, 2)
attention_hd64_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                      const __grid_constant__ CUtensorMap tma_vt, __half* __restrict__ out, int T, int H) {
    using S = AttnSmem<TKP>;
    static_assert(TKP % 16 == 0 && TKP <= 256, "key padding must be a legal UMMA N");
    constexpr int HD = 64;
    constexpr uint32_t kTmemCols = 256;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t* sP = smem;                 // overlays sQ|sK once S has been produced
    uint8_t* sQ = smem;
    uint8_t* sK = smem + S::kQBytes;
    uint8_t* sVt = smem + S::kRegion0;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sVt + S::kVtBytes);  // qk, v, s, o
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

    const int warp = threadIdx.x >> 5;
    const int bh = blockIdx.y;
    const int m0 = blockIdx.x * 128;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tma_q);
        tma_prefetch_desc(&tma_k);
        tma_prefetch_desc(&tma_vt);
        for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bars[0], S::kQBytes + S::kKBytes);
        tma_load_3d(sQ, &tma_q, &bars[0], 0, m0, bh);  // rows >= T of this head: zero-filled
        tma_load_3d(sK, &tma_k, &bars[0], 0, 0, bh);
        mbar_arrive_expect_tx(&bars[1], S::kVtBytes);
        for (int kb = 0; kb < S::kKeyBlocks; ++kb)
            tma_load_2d(sVt + kb * S::kVtBlockBytes, &tma_vt, &bars[1], kb * 64, bh * HD);
        // S = Q K^T
        mbar_wait(&bars[0], 0);
        tc_fence_after();
        constexpr uint32_t idesc_s = make_idesc_f16(128, TKP);
        const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
            umma_f16_ss(tmem, make_smem_desc_sw128(qa + k * 32), make_smem_desc_sw128(ka + k * 32), idesc_s, k != 0);
        umma_commit(&bars[2]);
    }

    // ---- softmax: thread r owns query row m0 + r = TMEM lane r
    mbar_wait(&bars[2], 0);
    __syncwarp();
    tc_fence_after();
    const int r = threadIdx.x;
    const uint32_t trow = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    float mx = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < TKP; c += 16) {
        uint32_t v[16];
        tmem_ld_x16(trow + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (c + j < T) mx = fmaxf(mx, __uint_as_float(v[j]));
    }
    const float mxl = mx * kLog2e;
    float sum = 0.f;
#pragma unroll 1
    for (int c = 0; c < TKP; c += 16) {
        uint32_t v[16];
        tmem_ld_x16(trow + c, v);
        tmem_ld_wait();
        float p[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            p[j] = (c + j < T) ? ex2f(fmaf(__uint_as_float(v[j]), kLog2e, -mxl)) : 0.f;
            sum += p[j];
        }
        // two 16-byte chunks (8 keys each) of row r in key block c/64, 128B-swizzled
        const int blk = c >> 6;
        const int chunk0 = (c & 63) >> 3;
        uint8_t* rowp = sP + blk * (128 * 128) + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
            uint4 u;
            u.x = pack_h2(p[8 * h2 + 0], p[8 * h2 + 1]);
            u.y = pack_h2(p[8 * h2 + 2], p[8 * h2 + 3]);
            u.z = pack_h2(p[8 * h2 + 4], p[8 * h2 + 5]);
            u.w = pack_h2(p[8 * h2 + 6], p[8 * h2 + 7]);
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
        constexpr uint32_t idesc_o = make_idesc_f16(128, HD);
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
    const int t = m0 + r;
    const int b = bh / H, h = bh - b * H;
    __half* orow = out + (static_cast<size_t>(b) * T + t) * (H * HD) + h * HD;
#pragma unroll
    for (int c = 0; c < HD; c += 32) {
        uint32_t v[32];
        tmem_ld_x32(trow + c, v);
        tmem_ld_wait();
        if (t < T) {
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

// q, k: fp16 [B,H,T,64]; vt: fp16 [B,H,64,Tpad] (zero beyond T); out: fp16 [B*T, H*64].
extern "C" int vitad_attention_f16(const void* q, const void* k, const void* vt, void* out, int batch, int heads,
                                   int tokens, int tokens_pad, int head_dim, void* stream) {
    int rc = check_device_arch();
    if (rc) return rc;
    VITAD_REQUIRE(q && k && vt && out, VITAD_ERR_ARG, "null pointer");
    VITAD_REQUIRE(head_dim == 64, VITAD_ERR_SHAPE, "head_dim %d unsupported (64)", head_dim);
    VITAD_REQUIRE(tokens > 0 && tokens <= 208, VITAD_ERR_SHAPE, "tokens=%d unsupported (1..208)", tokens);
    VITAD_REQUIRE(tokens_pad >= 256 && tokens_pad % 8 == 0, VITAD_ERR_SHAPE,
                  "tokens_pad=%d must be >= 256 and a multiple of 8", tokens_pad);
    VITAD_REQUIRE(aligned16(out), VITAD_ERR_ALIGN, "output alignment");
    constexpr int TKP = 208;
    using S = AttnSmem<TKP>;
    const int BH = batch * heads;
    CUtensorMap tq, tk, tv;
    rc = make_tmap_f16_3d(&tq, q, BH, tokens, 64, 64, static_cast<uint64_t>(tokens) * 64, 128);
    if (rc) return rc;
    rc = make_tmap_f16_3d(&tk, k, BH, tokens, 64, 64, static_cast<uint64_t>(tokens) * 64, TKP);
    if (rc) return rc;
    rc = make_tmap_f16_2d(&tv, vt, static_cast<uint64_t>(BH) * 64, tokens_pad, tokens_pad, 64);
    if (rc) return rc;
    auto kern = attention_hd64_kernel<TKP>;
    static bool attr_set = false;
    if (!attr_set) {
        VITAD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
        attr_set = true;
    }
    dim3 grid((tokens + 127) / 128, BH);
    ProfScope prof("attention", static_cast<cudaStream_t>(stream));
    kern<<<grid, 128, S::kTotal, static_cast<cudaStream_t>(stream)>>>(tq, tk, tv, static_cast<__half*>(out), tokens,
                                                                     heads);
    VITAD_CUDA_OK(cudaGetLastError());
    g_launches.fetch_add(1);
    return VITAD_OK;
}
