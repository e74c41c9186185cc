// Mixture-density ("GMM") head, fused:  src/classes/MixtureDensityNetwork.py:35-97,151-171 and the
// score tail of src/pipeline/ValidatorMDN.py:133-172.
//
// The reference materialises sigma and mu as two [B,P,768,K] fp32 tensors (1.9 GB each at B=32, K=100)
// and then makes ~10 elementwise passes over them.  Here the two 768 -> 768*K projections run as ONE
// tcgen05 GEMM per (128-token tile, feature d): the packed weight tile holds the K sigma rows and the K
// mu rows of feature d side by side, so the accumulator tile in TMEM is [128 tokens x (sigma_k | mu_k)],
// and the epilogue turns it straight into  LL[t,d] = logsumexp_k(log pi[t,k] + log N(x[t,d]; mu, sigma))
// without sigma/mu ever leaving the SM.  Biases ride in two extra K columns (hi/lo fp16 split) of the
// operands, so the epilogue starts from complete pre-activations.
//
// Packed layouts (mixtures are split into n_kc chunks of KCV = ceil(K/n_kc) valid + padding to KC):
//   Wpk   fp16 [768][n_kc][2 (sigma,mu)][KC][KA=784]   cols 0..767 weight row (d*K+k), 768/769 bias hi/lo
//   xaug  fp16 [M][KA]                                  cols 0..767 LayerNorm output, 768/769 = 1, rest 0
//   lp2   fp32 [M][n_kc*KC]                             log2(softmax(pi+g)+1e-15); padding = -1e30
#include <atomic>

#include "gemm_pair.cuh"
#include "gemm_quad.cuh"
#include "host_util.cuh"

namespace vitad {
extern std::atomic<uint64_t> g_launches;
extern std::atomic<int> g_use_pair;
std::atomic<int> g_gmm_quad{1};  // fused GMM kernel on 4-CTA clusters with A multicast (0: CTA pairs)
// Features the CTA-pair kernel takes on the SMs the 4-CTA clusters cannot cover (148 SMs hold 33 clusters of four: 16 SMs
// stay idle), launched right behind the 4-CTA grid as an independent programmatic dependent.  -1: balanced automatically,
// 0: off, > 0: that many.  Needs programmatic dependent launch (vitad_set_pdl(1), the default).
std::atomic<int> g_gmm_split{-1};

constexpr float kLog2eF = 1.4426950408889634f;
constexpr float kLn2F = 0.6931471805599453f;
constexpr float kHalfLog2Pi = 0.9189385332046727f;  // 0.5*log(2*pi)
constexpr float kLog2SigmaFloor = -49.82892142331043f;  // log2(1e-15): sigma = ELU+1+1e-15 never goes below
constexpr float kPadLogPi = -1e30f;
constexpr int kMdnKA = 784;  // 768 + 16: K extent of the packed operands

// Epilogue of the fused projection: per thread one token row, per tile one feature d.
template <int KC>
struct EpiMdn {
    const float* lp2;
    const float* x;
    float* ll;  // [768][ldl]
    int ldp, ldx, ldl, M;
    float m_run, s_run, xv;

    __device__ __forceinline__ void tile_begin(int, int d, int row) {
        m_run = -INFINITY;
        s_run = 0.f;
        xv = row < M ? __ldg(x + static_cast<size_t>(row) * ldx + d) : 0.f;
    }

    template <int W>
    __device__ __forceinline__ void chunk(const uint32_t* as, const uint32_t* am, const float* lp) {
        float v[W];
        float cm = -INFINITY;
#pragma unroll
        for (int j = 0; j < W; ++j) {
            const float a = __uint_as_float(as[j]);
            // log2(sigma), sigma = ELU(a)+1:  a > 0 -> log2(1+a);  a <= 0 -> a*log2(e)   (one MUFU)
            float t2 = lg2f(fmaxf(a, 0.f) + 1.f) + fminf(a, 0.f) * kLog2eF;
            t2 = fmaxf(t2, kLog2SigmaFloor);
            const float inv_s = ex2f(-t2);
            const float z = (xv - __uint_as_float(am[j])) * inv_s;
            v[j] = (lp[j] - t2) - (0.5f * kLog2eF) * z * z;
            cm = fmaxf(cm, v[j]);
        }
        const float m_new = fmaxf(m_run, cm);
        float acc = s_run * ex2f(m_run - m_new);
#pragma unroll
        for (int j = 0; j < W; ++j) acc += ex2f(v[j] - m_new);
        s_run = acc;
        m_run = m_new;
    }

    static constexpr bool kSplitColumns = false;  // group g owns accumulator stage g (see gemm_core.cuh)

    // Two-chunk mixtures (K > 112): each epilogue group holds the running (max, sum) of one chunk.
    __device__ __forceinline__ void merge(int group, float2* slot) {
        if (group == 1) {
            *slot = make_float2(m_run, s_run);
        } else {
            const float2 o = *slot;
            const float m_new = fmaxf(m_run, o.x);
            s_run = s_run * ex2f(m_run - m_new) + o.y * ex2f(o.x - m_new);
            m_run = m_new;
        }
    }

    __device__ __forceinline__ void sub(int kc, int, int, int row, uint32_t taddr, int, int) {
        const bool valid = row < M;
        const float* lprow = lp2 + static_cast<size_t>(valid ? row : 0) * ldp + kc * KC;
        constexpr int kFull = KC / 16;
#pragma unroll 1
        for (int c = 0; c < kFull; ++c) {
            uint32_t as[16], am[16];
            tmem_ld_x16(taddr + c * 16, as);
            tmem_ld_x16(taddr + KC + c * 16, am);
            float lp[16];
            const float4* l4 = reinterpret_cast<const float4*>(lprow + c * 16);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 t = __ldg(l4 + j);
                lp[4 * j + 0] = t.x;
                lp[4 * j + 1] = t.y;
                lp[4 * j + 2] = t.z;
                lp[4 * j + 3] = t.w;
            }
            tmem_ld_wait();
            chunk<16>(as, am, lp);
        }
        if constexpr (KC % 16 == 8) {
            uint32_t as[8], am[8];
            tmem_ld_x8(taddr + kFull * 16, as);
            tmem_ld_x8(taddr + KC + kFull * 16, am);
            float lp[8];
            const float4* l4 = reinterpret_cast<const float4*>(lprow + kFull * 16);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float4 t = __ldg(l4 + j);
                lp[4 * j + 0] = t.x;
                lp[4 * j + 1] = t.y;
                lp[4 * j + 2] = t.z;
                lp[4 * j + 3] = t.w;
            }
            tmem_ld_wait();
            chunk<8>(as, am, lp);
        }
    }

    __device__ __forceinline__ void tile_end(int, int d, int row) {
        if (row < M) ll[static_cast<size_t>(d) * ldl + row] = (m_run + lg2f(s_run)) * kLn2F - kHalfLog2Pi;
    }
};

// ------------------------------------------------------------------------------ weight packing
__global__ void __launch_bounds__(256) gmm_pack_kernel(const float* __restrict__ ws, const float* __restrict__ bs,
                                                       const float* __restrict__ wm, const float* __restrict__ bm,
                                                       __half* __restrict__ wpk, int D, int K, int n_kc, int KC,
                                                       int KCV, size_t total) {
    const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int col = static_cast<int>(idx % kMdnKA);
    size_t r = idx / kMdnKA;
    const int j = static_cast<int>(r % KC);
    r /= KC;
    const int which = static_cast<int>(r & 1);
    r >>= 1;
    const int kc = static_cast<int>(r % n_kc);
    const int d = static_cast<int>(r / n_kc);
    const int k = kc * KCV + j;
    float val = 0.f;
    if (j < KCV && k < K) {
        const size_t src_row = static_cast<size_t>(d) * K + k;  // view(B,P,D,K): k fastest
        const float* w = which ? wm : ws;
        const float* b = which ? bm : bs;
        if (col < D) {
            val = w[src_row * D + col];
        } else if (col == D) {
            val = __half2float(__float2half_rn(b[src_row]));
        } else if (col == D + 1) {
            const float hi = __half2float(__float2half_rn(b[src_row]));
            val = b[src_row] - hi;
        }
    }
    wpk[idx] = __float2half_rn(val);
}

// xaug[t][:] = (fp16(x[t][0..D)), 1, 1, 0...) for callers that hand the head plain fp32 features
// (the encoder's final LayerNorm emits this operand directly on the fast path).
__global__ void __launch_bounds__(256) gmm_operand_kernel(const float* __restrict__ x, int ldx,
                                                          __half* __restrict__ xaug, int M, int D) {
    const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;  // 4 columns per thread
    const int per_row = kMdnKA / 4;
    if (idx >= static_cast<size_t>(M) * per_row) return;
    const size_t t = idx / per_row;
    const int c = static_cast<int>(idx - t * per_row) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < D)
        v = *reinterpret_cast<const float4*>(x + t * ldx + c);
    else if (c == D)
        v.x = 1.f, v.y = 1.f;
    uint2 u;
    u.x = pack_h2(v.x, v.y);
    u.y = pack_h2(v.z, v.w);
    *reinterpret_cast<uint2*>(xaug + t * kMdnKA + c) = u;
}


// ------------------------------------------------------------------------ counter-based Gumbel noise
// The reference draws fresh Gumbel noise in every log_likelihood call (gumbel_softmax, MixtureDensityNetwork.py:62) from
// torch's global generator, so a score depends on how many draws preceded it.  Here the noise of element (t, k) of
// global batch `batch_index` is a pure function of (seed, batch_index, t, k): Philox4x32-10 with key = seed and counter =
// (t, k % 32, (k / 32) / 4, batch_index), word (k / 32) % 4, mapped to g = -log(-log(u)), u = ((word >> 9) + 0.5) * 2^-23 in
// (0, 1) — 23 bits, so that the + 0.5 is exact in fp32 and u never rounds to 1 (24 bits did, once in 2^24 draws: g = inf).  A sharded run (rank r scores batches r, r + W, ...) therefore reproduces the unsharded scores bit for bit.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 key) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ key.x, lo1, hi0 ^ c.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return c;
}
__device__ __forceinline__ float gumbel_from_word(uint32_t w) {
    const float u = (static_cast<float>(w >> 9) + 0.5f) * 1.1920928955078125e-07f;  // 2^-23: exact, in [2^-24, 1 - 2^-24]
    return -logf(-logf(u));
}
struct GumbelKey {
    uint32_t seed_lo, seed_hi, batch_index;
};
// the four noise values of lane tx (mixtures tx + 32 (4 q + w), w = 0..3) of token t
__device__ __forceinline__ void gumbel4(const GumbelKey& gk, int t, int tx, int q, float (&g)[4]) {
    const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(t), static_cast<uint32_t>(tx), static_cast<uint32_t>(q),
                                             gk.batch_index),
                                  make_uint2(gk.seed_lo, gk.seed_hi));
    g[0] = gumbel_from_word(r.x);
    g[1] = gumbel_from_word(r.y);
    g[2] = gumbel_from_word(r.z);
    g[3] = gumbel_from_word(r.w);
}
// out[t][k] fp32 = the noise the seeded log_pi kernels add (tests, and callers that want to see it)
__global__ void __launch_bounds__(256) gumbel_fill_kernel(float* __restrict__ out, int M, int K, GumbelKey gk) {
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int tx = threadIdx.x & 31;
    if (t >= M) return;
    for (int q = 0; q * 128 < K; ++q) {
        float g[4];
        gumbel4(gk, t, tx, q, g);
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const int k = tx + 32 * (4 * q + w);
            if (k < K) out[static_cast<size_t>(t) * K + k] = g[w];
        }
    }
}

// ---------------------------------------------------------------------------- mixing weights
// lp2[t][kmap(k)] = log2( softmax_k(x[t].Wpi[k] + bpi[k] + g[t][k]) + 1e-15 )   — fp32 on CUDA cores: the
// logits enter every feature's logsumexp with the same sign, so they need better than fp16-GEMM accuracy.
// CTA = 4 warps x TPW tokens x up to 160 mixtures: lane tx of warp ty accumulates tokens (ty*TPW .. +TPW-1) x mixtures
// tx+32*j in registers, so a token's logits stay inside its warp and the softmax needs shuffles only.  TPW is chosen
// by the host so that the tokens fill the 148 SMs in ONE balanced wave (M = 6272 -> TPW 11, 143 CTAs): with a fixed
// 32-token CTA the 196 CTAs left 48 SMs with two CTAs and the rest with one (88 us; this form: see DESIGN.md 4.5).
constexpr int kPiBK = 32, kPiMaxK = 224;
// NJ = mixture slots per lane (K <= 32 NJ): 4 for K <= 128 (K = 100: 20 % fewer FMAs and weight loads than 5), 5 / 6 / 7 up to
// 160 / 192 / 224.  gumbel == nullptr: the noise is generated in place from `gk` (see above).
template <int TPW, int WARPS, int NJ>
__global__ void __launch_bounds__(32 * WARPS) gmm_logpi_kernel(const float* __restrict__ x, int ldx,
                                                               const float* __restrict__ wpi, const float* __restrict__ bpi,
                                                               const float* __restrict__ gumbel, GumbelKey gk,
                                                               float* __restrict__ lp2, int M, int D, int K, int n_kc, int KC,
                                                               int KCV) {
    griddep_launch_dependents();
    griddep_wait();
    constexpr int BM = WARPS * TPW;
    constexpr int kPiThreads = 32 * WARPS;
    // k is the contiguous index of both staging arrays and is read four at a time (LDS.128; the x reads are warp-wide
    // broadcasts).  Row pitch 36 floats: 16-byte aligned, the 8 lanes of a quarter-warp land on 8 distinct bank groups.
    constexpr int kPitch = kPiBK + 4;
    __shared__ __align__(16) float xs[BM][kPitch];
    __shared__ __align__(16) float wsh[32 * NJ][kPitch];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int m0 = blockIdx.x * BM;
    float acc[TPW][NJ];
#pragma unroll
    for (int i = 0; i < TPW; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j] = 0.f;
    // global -> register prefetch of the next k-block overlaps the FMAs of the current one
    constexpr int kQ = kPiBK / 4;                  // float4 per staged row
    constexpr int kRowsPerPass = kPiThreads / kQ;  // 16 rows staged per pass
    constexpr int kXPasses = (BM + kRowsPerPass - 1) / kRowsPerPass;
    const int xr = threadIdx.x / kQ, xc = (threadIdx.x % kQ) * 4;
    float4 px[kXPasses], pw[32 * NJ / kRowsPerPass];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int j = 0; j < kXPasses; ++j) {
            const int r = xr + kRowsPerPass * j;
            px[j] = (r < BM && m0 + r < M) ? *reinterpret_cast<const float4*>(x + static_cast<size_t>(m0 + r) * ldx + k0 + xc)
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < 32 * NJ / kRowsPerPass; ++j) {
            const int r = xr + kRowsPerPass * j;
            pw[j] = (r < K) ? __ldg(reinterpret_cast<const float4*>(wpi + static_cast<size_t>(r) * D + k0 + xc))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    fetch(0);
    for (int k0 = 0; k0 < D; k0 += kPiBK) {
#pragma unroll
        for (int j = 0; j < kXPasses; ++j)
            if (xr + kRowsPerPass * j < BM) *reinterpret_cast<float4*>(&xs[xr + kRowsPerPass * j][xc]) = px[j];
#pragma unroll
        for (int j = 0; j < 32 * NJ / kRowsPerPass; ++j) *reinterpret_cast<float4*>(&wsh[xr + kRowsPerPass * j][xc]) = pw[j];
        __syncthreads();
        if (k0 + kPiBK < D) fetch(k0 + kPiBK);
#pragma unroll
        for (int k = 0; k < kPiBK; k += 4) {
            float4 wv[NJ];
#pragma unroll
            for (int j = 0; j < NJ; ++j) wv[j] = *reinterpret_cast<const float4*>(&wsh[tx + 32 * j][k]);
#pragma unroll
            for (int i = 0; i < TPW; ++i) {
                const float4 xv = *reinterpret_cast<const float4*>(&xs[ty * TPW + i][k]);
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    acc[i][j] = fmaf(xv.x, wv[j].x, acc[i][j]);
                    acc[i][j] = fmaf(xv.y, wv[j].y, acc[i][j]);
                    acc[i][j] = fmaf(xv.z, wv[j].z, acc[i][j]);
                    acc[i][j] = fmaf(xv.w, wv[j].w, acc[i][j]);
                }
            }
        }
        __syncthreads();
    }
    // softmax + log per token, straight from the accumulators: lane tx holds mixtures tx + 32 j of each of its warp's tokens
    const int ldp = n_kc * KC;
#pragma unroll
    for (int i = 0; i < TPW; ++i) {
        const int t = m0 + ty * TPW + i;
        if (t >= M) break;  // warp-uniform
        float z[NJ];
        float mx = -INFINITY;
        float gn[(NJ + 3) / 4 * 4];
        if (gumbel == nullptr) {  // warp-uniform
#pragma unroll
            for (int q = 0; q < (NJ + 3) / 4; ++q) {
                float g4[4];
                gumbel4(gk, t, tx, q, g4);
#pragma unroll
                for (int w = 0; w < 4; ++w) gn[4 * q + w] = g4[w];
            }
        }
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int k = tx + 32 * j;
            const float g = gumbel != nullptr ? (k < K ? gumbel[static_cast<size_t>(t) * K + k] : 0.f) : gn[j];
            z[j] = (k < K) ? acc[i][j] + bpi[k] + g : -INFINITY;
            mx = fmaxf(mx, z[j]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float e[NJ], sum = 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            e[j] = (tx + 32 * j < K) ? expf(z[j] - mx) : 0.f;
            sum += e[j];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        float* dst = lp2 + static_cast<size_t>(t) * ldp;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int k = tx + 32 * j;
            if (k < K) {
                const int kc = k / KCV;
                dst[kc * KC + (k - kc * KCV)] = log2f(e[j] / sum + 1e-15f);
            }
        }
        // padding slots of this row
        for (int s = tx; s < ldp; s += 32) {
            const int kc = s / KC, j = s - kc * KC;
            if (j >= KCV || kc * KCV + j >= K) dst[s] = kPadLogPi;
        }
    }
}

template <int TPW, int WARPS>
static cudaError_t launch_logpi(cudaStream_t s, const float* x, int ldx, const float* pi_w, const float* pi_b,
                                const float* gumbel, GumbelKey gk, float* lp2, int tokens, int dim, int K, int n_kc, int kc,
                                int kcv) {
    const dim3 grid((tokens + WARPS * TPW - 1) / (WARPS * TPW)), block(32 * WARPS);
#define VITAD_LOGPI_NJ(NJ)                                                                                                    \
    return launch_pdl(gmm_logpi_kernel<TPW, WARPS, NJ>, grid, block, 0, s, x, ldx, pi_w, pi_b, gumbel, gk, lp2, tokens, dim, K, \
                      n_kc, kc, kcv)
    if (K <= 128) VITAD_LOGPI_NJ(4);
    if (K <= 160) VITAD_LOGPI_NJ(5);
    if (K <= 192) VITAD_LOGPI_NJ(6);
    VITAD_LOGPI_NJ(7);
#undef VITAD_LOGPI_NJ
}

// ---------------------------------------------------------- mixing weights on the tensor cores
// The same projection as one fp16 GEMM with fp32-grade accuracy: x = xh + xl, W = Wh + Wl (each an fp16 value, xl and Wl
// the rounding remainders), and  x.W ~= xh.Wh + xl.Wh + xh.Wl  — every fp16 x fp16 product is exact in the fp32
// accumulator, the dropped xl.Wl term is 2^-22 relative.  Laid out as K' = 3 * 768:  A' = [xh | xl | xh],
// B' = [Wh | Wh | Wl].  (A plain fp16 GEMM is not enough here: the logits enter every feature's logsumexp with the same
// sign.)  The fp32 CUDA-core kernel above needs 88 us at batch 32 and 67 us at batch 1 (7 CTAs, serial k loop); this
// path is an operand kernel + a 2304-deep GEMM + a per-token softmax kernel.
__global__ void __launch_bounds__(256) gmm_pi_operand_kernel(const float* __restrict__ x, int ldx, __half* __restrict__ a3,
                                                             int M, int D) {
    griddep_launch_dependents();
    griddep_wait();
    const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;  // one float4 of x
    const int per_row = D >> 2;
    if (idx >= static_cast<size_t>(M) * per_row) return;
    const size_t t = idx / per_row;
    const int c = static_cast<int>(idx - t * per_row) << 2;
    const float4 v = *reinterpret_cast<const float4*>(x + t * ldx + c);
    const float f[4] = {v.x, v.y, v.z, v.w};
    __half hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        hi[j] = to_h(f[j]);
        lo[j] = to_h(f[j] - __half2float(hi[j]));
    }
    __half* row = a3 + t * (3 * static_cast<size_t>(D));
    *reinterpret_cast<uint2*>(row + c) = *reinterpret_cast<const uint2*>(hi);
    *reinterpret_cast<uint2*>(row + D + c) = *reinterpret_cast<const uint2*>(lo);
    *reinterpret_cast<uint2*>(row + 2 * D + c) = *reinterpret_cast<const uint2*>(hi);
}

// B' = [Wh | Wh | Wl] from pi.weight fp32 [K, D]
__global__ void __launch_bounds__(256) gmm_pi_pack_kernel(const float* __restrict__ w, __half* __restrict__ b3, int K,
                                                          int D) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= K * D) return;
    const int k = idx / D, c = idx - k * D;
    const float f = w[idx];
    const __half hi = to_h(f);
    const __half lo = to_h(f - __half2float(hi));
    __half* row = b3 + static_cast<size_t>(k) * 3 * D;
    row[c] = hi;
    row[D + c] = hi;
    row[2 * D + c] = lo;
}

// lp2 row of one token from its logits: + bias + gumbel, softmax, log2(. + 1e-15), chunked slot layout.  One warp per token.
__global__ void __launch_bounds__(256) gmm_pi_softmax_kernel(const float* __restrict__ logits, int ldl,
                                                             const float* __restrict__ bpi,
                                                             const float* __restrict__ gumbel, float* __restrict__ lp2,
                                                             int M, int K, int n_kc, int KC, int KCV) {
    griddep_launch_dependents();
    griddep_wait();
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int tx = threadIdx.x & 31;
    if (t >= M) return;
    const int ldp = n_kc * KC;
    float z[5];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const int k = tx + 32 * j;
        z[j] = (k < K) ? logits[static_cast<size_t>(t) * ldl + k] + bpi[k] + gumbel[static_cast<size_t>(t) * K + k] : -INFINITY;
        mx = fmaxf(mx, z[j]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float e[5], sum = 0.f;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        e[j] = (tx + 32 * j < K) ? expf(z[j] - mx) : 0.f;
        sum += e[j];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    float* dst = lp2 + static_cast<size_t>(t) * ldp;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const int k = tx + 32 * j;
        if (k < K) {
            const int kc = k / KCV;
            dst[kc * KC + (k - kc * KCV)] = log2f(e[j] / sum + 1e-15f);
        }
    }
    for (int s = tx; s < ldp; s += 32) {
        const int kc = s / KC, j = s - kc * KC;
        if (j >= KCV || kc * KCV + j >= K) dst[s] = kPadLogPi;
    }
}

// ------------------------------------------------------------------------------- score tail
// L[t] = mean_d LL[d][t]   (torch.mean over features, MixtureDensityNetwork.py:86-88); fixed summation order.
// CTA = 32 tokens x 32 feature groups (1024 threads): a warp reads 128 contiguous bytes of one feature row, group g
// sums features g, g+32, ... with 8 loads in flight, the 32 partial sums are added in group order -> bit-reproducible.
constexpr int kMeanGroups = 32;
__global__ void __launch_bounds__(32 * kMeanGroups) gmm_mean_kernel(const float* __restrict__ ll, int ldl,
                                                                    float* __restrict__ L, int M, int D) {
    griddep_launch_dependents();
    griddep_wait();
    __shared__ float part[kMeanGroups][33];
    const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int t = blockIdx.x * 32 + lane;
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (t < M) {
        int d = g;
        for (; d + 7 * kMeanGroups < D; d += 8 * kMeanGroups) {
#pragma unroll
            for (int i = 0; i < 8; ++i) s[i] += ll[static_cast<size_t>(d + i * kMeanGroups) * ldl + t];
        }
        for (; d < D; d += kMeanGroups) s[0] += ll[static_cast<size_t>(d) * ldl + t];
    }
    part[g][lane] = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
    __syncthreads();
    if (g == 0 && t < M) {
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < kMeanGroups; ++i) a += part[i][lane];
        L[t] = a / D;
    }
}

// One CTA: batch-global max (MixtureDensityNetwork.py:90-92), prob = exp(L - max) (:93-95),
// image score = 1 - min_p prob (ValidatorMDN.py:133,170).
__global__ void __launch_bounds__(1024) gmm_finish_kernel(const float* __restrict__ L, float* __restrict__ prob,
                                                          float* __restrict__ scores, int B, int P) {
    griddep_launch_dependents();
    griddep_wait();
    __shared__ float red[32];
    __shared__ float gmax;
    const int M = B * P;
    float mx = -INFINITY;
    // a NaN log-likelihood poisons the batch maximum, hence every probability and score, as torch.max does in the
    // reference (:90-92); fmaxf / fminf alone would drop it
    bool saw_nan = false;
    for (int i = threadIdx.x; i < M; i += blockDim.x) {
        mx = fmaxf(mx, L[i]);
        saw_nan |= L[i] != L[i];
    }
    saw_nan = __syncthreads_or(saw_nan);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : -INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
        if (threadIdx.x == 0) gmax = saw_nan ? __int_as_float(0x7FC00000) : v;
    }
    __syncthreads();
    const float g = gmax;
    for (int i = threadIdx.x; i < M; i += blockDim.x) prob[i] = expf(L[i] - g);
    // per-image min over patches: one warp per image
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int b = warp; b < B; b += (blockDim.x >> 5)) {
        float mn = INFINITY;
        for (int p = lane; p < P; p += 32) mn = fminf(mn, expf(L[b * P + p] - g));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        if (lane == 0) scores[b] = saw_nan ? g : 1.0f - mn;
    }
}

template <int KC, int NKC>
static int launch_mdn(const void* xaug, const void* wpk, const float* lp2, const float* x, int ldx, float* ll, int ldl,
                      int M, int D, cudaStream_t stream) {
    constexpr int BN = 2 * KC;
    using S = GemmSmem<BN>;
    using Epi = EpiMdn<KC>;
    CUtensorMap ta, tb;
    int rc = make_tmap_f16_2d(&ta, xaug, M, kMdnKA, kMdnKA, kBlockM);
    if (rc) return rc;
    rc = make_tmap_f16_2d(&tb, wpk, static_cast<uint64_t>(D) * NKC * BN, kMdnKA, kMdnKA, BN);
    if (rc) return rc;
    Epi epi{lp2, x, ll, NKC * KC, ldx, ldl, M, 0.f, 0.f, 0.f};
    if (g_use_pair.load() && g_gmm_quad.load() && M > 2 * kBlockM && D % 2 == 0) {
        // clusters of four CTAs: two pairs share the A block of one 256-token tile by TMA multicast (gemm_quad.cuh)
        using SP = PairSmem<BN>;
        CUtensorMap ta64;
        rc = make_tmap_f16_2d(&ta64, xaug, M, kMdnKA, kMdnKA, kBlockM / 2);
        if (rc) return rc;
        rc = make_tmap_f16_2d(&tb, wpk, static_cast<uint64_t>(D) * NKC * BN, kMdnKA, kMdnKA, BN / 2);
        if (rc) return rc;
        auto kern4 = gemm4_tc_kernel<BN, NKC, Epi>;
        static int max_clusters4 = -1;
        if (max_clusters4 < 0) {
            VITAD_CUDA_OK(cudaFuncSetAttribute(kern4, cudaFuncAttributeMaxDynamicSharedMemorySize, SP::kTotalBytes));
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(device_sm_count() / 4 * 4);
            cfg.blockDim = dim3(kGemmThreads);
            cfg.dynamicSmemBytes = SP::kTotalBytes;
            int n = 0;
            VITAD_CUDA_OK(cudaOccupancyMaxActiveClusters(&n, kern4, &cfg));  // 33 on a 148-SM B200
            VITAD_REQUIRE(n > 0, VITAD_ERR_CUDA, "no 4-CTA cluster fits this device");
            max_clusters4 = n;
        }
        const int num_m4 = (M + 2 * kBlockM - 1) / (2 * kBlockM);
        // Side launch: the last Db features go to the CTA-pair kernel on the leftover SMs.  A quad tile (two features) costs
        // 0.843 of two pair tiles (measured: 1.22 ms on 33 quads vs 1.29 ms on 74 pairs), which balances the two kernels.
        const int spare_pairs = (device_sm_count() - 4 * max_clusters4) / 2;
        int Db = g_gmm_split.load();
        if (Db < 0) {
            const double ratio = 0.843 * spare_pairs / (2.0 * max_clusters4);
            Db = static_cast<int>(D * ratio / (1.0 + ratio) + 0.5);
        }
        Db &= ~1;
        if (spare_pairs < 1 || num_m4 < 4 || Db >= D || !pdl_enabled()) Db = 0;
        const int Dq = D - Db;
        const int tiles4 = num_m4 * (Dq / 2);
        const int clusters = tiles4 < max_clusters4 ? tiles4 : max_clusters4;
        VITAD_CUDA_OK(launch_pdl(kern4, dim3(4 * clusters), dim3(kGemmThreads), SP::kTotalBytes, stream, ta64, tb, M, Dq, kMdnKA, epi));
        VITAD_CUDA_OK(cudaGetLastError());
        g_launches.fetch_add(1);
        if (Db > 0) {
            CUtensorMap tb2;
            const __half* w2 = static_cast<const __half*>(wpk) + static_cast<size_t>(Dq) * NKC * BN * kMdnKA;
            rc = make_tmap_f16_2d(&tb2, w2, static_cast<uint64_t>(Db) * NKC * BN, kMdnKA, kMdnKA, BN / 2);
            if (rc) return rc;
            Epi epi2 = epi;
            epi2.x = x + Dq;
            epi2.ll = ll + static_cast<size_t>(Dq) * ldl;
            // same stream, programmatic launch: starts when every CTA of the 4-CTA grid is resident (gemm_quad.cuh), reads only
            // what that grid's own prerequisites produced, and waits for that grid before it exits (kIndependent)
            auto kern2 = gemm2_tc_kernel<BN, NKC, Epi, true>;
            static bool attr2i_set = false;
            if (!attr2i_set) {
                VITAD_CUDA_OK(cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, SP::kTotalBytes));
                attr2i_set = true;
            }
            VITAD_CUDA_OK(launch_pdl(kern2, dim3(2 * spare_pairs), dim3(kGemmThreads), SP::kTotalBytes, stream, ta, tb2, M, Db, kMdnKA,
                                     epi2));
            VITAD_CUDA_OK(cudaGetLastError());
            g_launches.fetch_add(1);
        }
        return VITAD_OK;
    }
    if (g_use_pair.load() && M > kBlockM) {
        // CTA pairs: rank 0 stages the sigma rows of feature d, rank 1 the mu rows (the two halves of the tile)
        using SP = PairSmem<BN>;
        rc = make_tmap_f16_2d(&tb, wpk, static_cast<uint64_t>(D) * NKC * BN, kMdnKA, kMdnKA, BN / 2);
        if (rc) return rc;
        auto kern2 = gemm2_tc_kernel<BN, NKC, Epi>;
        static bool attr2_set = false;
        if (!attr2_set) {
            VITAD_CUDA_OK(cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, SP::kTotalBytes));
            attr2_set = true;
        }
        const int num_m2 = (M + 2 * kBlockM - 1) / (2 * kBlockM);
        const int tiles2 = num_m2 * D;
        const int max_clusters = device_sm_count() / 2;
        const int clusters = tiles2 < max_clusters ? tiles2 : max_clusters;
        VITAD_CUDA_OK(launch_pdl(kern2, dim3(2 * clusters), dim3(kGemmThreads), SP::kTotalBytes, stream, ta, tb, M, D, kMdnKA, epi));
        VITAD_CUDA_OK(cudaGetLastError());
        g_launches.fetch_add(1);
        return VITAD_OK;
    }
    auto kern = gemm_tc_kernel<BN, NKC, Epi>;
    static bool attr_set = false;
    if (!attr_set) {
        VITAD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotalBytes));
        attr_set = true;
    }
    const int num_m = (M + kBlockM - 1) / kBlockM;
    const int tiles = num_m * D;
    const int grid = tiles < device_sm_count() ? tiles : device_sm_count();
    VITAD_CUDA_OK(launch_pdl(kern, dim3(grid), dim3(kGemmThreads), S::kTotalBytes, stream, ta, tb, M, D, kMdnKA, epi));
    VITAD_CUDA_OK(cudaGetLastError());
    g_launches.fetch_add(1);
    return VITAD_OK;
}

}  // namespace vitad

using namespace vitad;

extern "C" void vitad_set_gmm_cluster4(int enable) { vitad::g_gmm_quad.store(enable ? 1 : 0); }
extern "C" void vitad_set_gmm_split(int features) { vitad::g_gmm_split.store(features); }

extern "C" int vitad_gmm_plan(int num_gaussians, int* n_kc, int* kc, int* kcv) {
    VITAD_REQUIRE(n_kc && kc && kcv, VITAD_ERR_ARG, "null pointer");
    VITAD_REQUIRE(num_gaussians >= 1 && num_gaussians <= 208, VITAD_ERR_SHAPE,
                  "num_gaussians=%d unsupported (1..208)", num_gaussians);
    if (num_gaussians <= 104) {  // K = 100: 104 slots (4% padding); the tile is 256 tokens x (104 sigma | 104 mu)
        *n_kc = 1, *kc = 104, *kcv = num_gaussians;
    } else if (num_gaussians <= 112) {
        *n_kc = 1, *kc = 112, *kcv = num_gaussians;
    } else {
        // two chunks of ceil(K/2) mixtures: 72 slots up to K = 144 (K = 130: 65 valid each), 80 up to 160 (the reference's
        // default K = 150, startTraining_mdn.py:37: 75 valid), 88 up to 176 (K = 170 of csv_results_gmm), 104 up to 208
        *n_kc = 2, *kcv = (num_gaussians + 1) / 2;
        *kc = num_gaussians <= 144 ? 72 : num_gaussians <= 160 ? 80 : num_gaussians <= 176 ? 88 : 104;
    }
    return VITAD_OK;
}

extern "C" size_t vitad_gmm_packed_weight_bytes(int dim, int num_gaussians) {
    int n_kc, kc, kcv;
    if (vitad_gmm_plan(num_gaussians, &n_kc, &kc, &kcv)) return 0;
    return static_cast<size_t>(dim) * n_kc * 2 * kc * kMdnKA * sizeof(__half);
}

extern "C" int vitad_gmm_pack_weights(const float* sigma_w, const float* sigma_b, const float* mu_w, const float* mu_b,
                                      int dim, int num_gaussians, void* packed, void* stream) {
    VITAD_NVTX("vitad_gmm_pack_weights");
    int rc = check_device_arch();
    if (rc) return rc;
    VITAD_REQUIRE(sigma_w && sigma_b && mu_w && mu_b && packed, VITAD_ERR_ARG, "null pointer");
    VITAD_REQUIRE(dim == 768, VITAD_ERR_SHAPE, "dim=%d unsupported (768)", dim);
    int n_kc, kc, kcv;
    rc = vitad_gmm_plan(num_gaussians, &n_kc, &kc, &kcv);
    if (rc) return rc;
    const size_t total = static_cast<size_t>(dim) * n_kc * 2 * kc * kMdnKA;
    gmm_pack_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        sigma_w, sigma_b, mu_w, mu_b, static_cast<__half*>(packed), dim, num_gaussians, n_kc, kc, kcv, total);
    VITAD_CUDA_OK(cudaGetLastError());
    g_launches.fetch_add(1);
    return VITAD_OK;
}

extern "C" int vitad_gmm_make_operand(const float* x, int ldx, void* xaug, int tokens, int dim, void* stream) {
    int rc = check_device_arch();
    if (rc) return rc;
    VITAD_REQUIRE(x && xaug && tokens > 0, VITAD_ERR_ARG, "null pointer");
    VITAD_REQUIRE(dim == 768 && ldx % 4 == 0 && aligned16(x) && aligned16(xaug), VITAD_ERR_ALIGN,
                  "dim must be 768 and x 16-byte aligned with pitch %% 4 == 0");
    const size_t total = static_cast<size_t>(tokens) * (kMdnKA / 4);
    gmm_operand_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, ldx, static_cast<__half*>(xaug), tokens, dim);
    VITAD_CUDA_OK(cudaGetLastError());
    g_launches.fetch_add(1);
    return VITAD_OK;
}

static int gmm_log_pi_impl(const float* x, int ldx, const float* pi_w, const float* pi_b, const float* gumbel, GumbelKey gk,
                           float* lp2, int tokens, int dim, int num_gaussians, void* stream) {
    VITAD_NVTX("vitad_gmm_log_pi");
    int rc = check_device_arch();
    if (rc) return rc;
    VITAD_REQUIRE(x && pi_w && pi_b && lp2, VITAD_ERR_ARG, "null pointer");
    VITAD_REQUIRE(dim % kPiBK == 0 && tokens > 0 && num_gaussians <= kPiMaxK, VITAD_ERR_SHAPE, "dim %% 32 != 0, no tokens or K > 224");
    VITAD_REQUIRE(ldx % 4 == 0 && aligned16(x) && aligned16(pi_w), VITAD_ERR_ALIGN,
                  "log_pi: x / pi_w must be 16-byte aligned with ldx %% 4 == 0");
    int n_kc, kc, kcv;
    rc = vitad_gmm_plan(num_gaussians, &n_kc, &kc, &kcv);
    if (rc) return rc;
    ProfScope prof("gmm_logpi", static_cast<cudaStream_t>(stream));
    // tokens per warp so that one wave of CTAs covers the device: 8 warps per CTA (two per scheduler hide the
    // shared-memory and FMA latencies), 4 for small inputs so that enough CTAs exist
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int sms = device_sm_count();
    cudaError_t e;
#define VITAD_LOGPI(T, W) e = launch_logpi<T, W>(s, x, ldx, pi_w, pi_b, gumbel, gk, lp2, tokens, dim, num_gaussians, n_kc, kc, kcv)
    if (tokens <= 8 * sms) {
        const int per_warp = (tokens + 4 * sms - 1) / (4 * sms);
        if (per_warp <= 1) VITAD_LOGPI(1, 4);
        else VITAD_LOGPI(2, 4);
    } else {
        const int per_warp = (tokens + 8 * sms - 1) / (8 * sms);
        if (per_warp <= 2) VITAD_LOGPI(2, 8);
        else if (per_warp <= 3) VITAD_LOGPI(3, 8);
        else if (per_warp <= 4) VITAD_LOGPI(4, 8);
        else if (per_warp <= 6) VITAD_LOGPI(6, 8);
        else VITAD_LOGPI(8, 8);
    }
#undef VITAD_LOGPI
    VITAD_CUDA_OK(e);
    VITAD_CUDA_OK(cudaGetLastError());
    g_launches.fetch_add(1);
    return VITAD_OK;
}

extern "C" int vitad_gmm_log_pi(const float* x, int ldx, const float* pi_w, const float* pi_b, const float* gumbel,
                                float* lp2, int tokens, int dim, int num_gaussians, void* stream) {
    VITAD_REQUIRE(gumbel, VITAD_ERR_ARG, "null gumbel (use vitad_gmm_log_pi_seeded for in-kernel noise)");
    return gmm_log_pi_impl(x, ldx, pi_w, pi_b, gumbel, GumbelKey{0u, 0u, 0u}, lp2, tokens, dim, num_gaussians, stream);
}

// Same, with the Gumbel noise generated in the kernel from (seed, batch_index, token, mixture): see gumbel4().
extern "C" int vitad_gmm_log_pi_seeded(const float* x, int ldx, const float* pi_w, const float* pi_b, uint64_t seed,
                                       uint32_t batch_index, float* lp2, int tokens, int dim, int num_gaussians,
                                       void* stream) {
    const GumbelKey gk{static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), batch_index};
    return gmm_log_pi_impl(x, ldx, pi_w, pi_b, nullptr, gk, lp2, tokens, dim, num_gaussians, stream);
}

// out fp32 [tokens, K]: the noise vitad_gmm_log_pi_seeded adds for (seed, batch_index).
extern "C" int vitad_gumbel_noise(uint64_t seed, uint32_t batch_index, float* out, int tokens, int num_gaussians,
                                  void* stream) {
    int rc = check_device_arch();
    if (rc) return rc;
    VITAD_REQUIRE(out && tokens > 0 && num_gaussians > 0 && num_gaussians <= kPiMaxK, VITAD_ERR_ARG, "gumbel_noise args");
    const GumbelKey gk{static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), batch_index};
    gumbel_fill_kernel<<<(tokens + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(out, tokens, num_gaussians, gk);
    VITAD_CUDA_OK(cudaGetLastError());
    g_launches.fetch_add(1);
    return VITAD_OK;
}

extern "C" size_t vitad_gmm_pi_packed_bytes(int dim, int num_gaussians) {
    return static_cast<size_t>(num_gaussians) * 3 * dim * sizeof(__half);
}

extern "C" int vitad_gmm_pack_pi(const float* pi_w, int dim, int num_gaussians, void* packed, void* stream) {
    int rc = check_device_arch();
    if (rc) return rc;
    VITAD_REQUIRE(pi_w && packed && dim % 16 == 0 && num_gaussians > 0, VITAD_ERR_ARG, "pack_pi arguments");
    const int total = num_gaussians * dim;
    gmm_pi_pack_kernel<<<(total + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        pi_w, static_cast<__half*>(packed), num_gaussians, dim);
    VITAD_CUDA_OK(cudaGetLastError());
    g_launches.fetch_add(1);
    return VITAD_OK;
}

extern "C" size_t vitad_gmm_log_pi_workspace_bytes(int tokens, int dim, int num_gaussians) {
    const size_t a3 = (static_cast<size_t>(tokens) * 3 * dim * sizeof(__half) + 255) & ~static_cast<size_t>(255);
    const size_t lg = static_cast<size_t>(tokens) * ((num_gaussians + 3) / 4 * 4) * sizeof(float);
    return a3 + lg;
}

extern "C" int vitad_linear_f16(const vitad_linear_args* args, void* stream);

// Tensor-core form of vitad_gmm_log_pi: pi_packed from vitad_gmm_pack_pi, workspace of vitad_gmm_log_pi_workspace_bytes.
extern "C" int vitad_gmm_log_pi_tc(const float* x, int ldx, const void* pi_packed, const float* pi_b, const float* gumbel,
                                   float* lp2, int tokens, int dim, int num_gaussians, void* workspace,
                                   size_t workspace_bytes, void* stream) {
    VITAD_NVTX("vitad_gmm_log_pi_tc");
    int rc = check_device_arch();
    if (rc) return rc;
    VITAD_REQUIRE(x && pi_packed && pi_b && gumbel && lp2 && workspace, VITAD_ERR_ARG, "null pointer");
    VITAD_REQUIRE(dim % 16 == 0 && tokens > 0 && num_gaussians <= 160, VITAD_ERR_SHAPE, "dim %% 16 != 0, no tokens or K > 160");
    VITAD_REQUIRE(ldx % 4 == 0 && aligned16(x) && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, VITAD_ERR_ALIGN,
                  "log_pi_tc: x 16-byte aligned with ldx %% 4 == 0, workspace 256-byte aligned");
    VITAD_REQUIRE(workspace_bytes >= vitad_gmm_log_pi_workspace_bytes(tokens, dim, num_gaussians), VITAD_ERR_WORKSPACE,
                  "log_pi_tc workspace too small");
    int n_kc, kc, kcv;
    rc = vitad_gmm_plan(num_gaussians, &n_kc, &kc, &kcv);
    if (rc) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    __half* a3 = static_cast<__half*>(workspace);
    const size_t a3_bytes = (static_cast<size_t>(tokens) * 3 * dim * sizeof(__half) + 255) & ~static_cast<size_t>(255);
    float* logits = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + a3_bytes);
    const int ldl = (num_gaussians + 3) / 4 * 4;
    ProfScope prof("gmm_logpi", s);
    const size_t n4 = static_cast<size_t>(tokens) * (dim / 4);
    VITAD_CUDA_OK(launch_pdl(gmm_pi_operand_kernel, dim3(static_cast<unsigned>((n4 + 255) / 256)), dim3(256), 0, s, x, ldx, a3,
                             tokens, dim));
    g_launches.fetch_add(1);
    vitad_linear_args la;
    memset(&la, 0, sizeof(la));
    la.a = a3, la.w = pi_packed, la.m = tokens, la.n = num_gaussians, la.k = 3 * dim, la.lda = 3 * dim, la.ldw = 3 * dim;
    la.epilogue = VITAD_EPI_F32, la.out = logits, la.ldo = ldl;
    if ((rc = vitad_linear_f16(&la, s))) return rc;
    VITAD_CUDA_OK(launch_pdl(gmm_pi_softmax_kernel, dim3((tokens + 7) / 8), dim3(256), 0, s, static_cast<const float*>(logits), ldl,
                             pi_b, gumbel, lp2, tokens, num_gaussians, n_kc, kc, kcv));
    g_launches.fetch_add(1);
    return VITAD_OK;
}

// xaug fp16 [tokens,784]; packed from vitad_gmm_pack_weights; lp2 from vitad_gmm_log_pi; x fp32 [tokens,ldx];
// ll_ws fp32 workspace [768][ld_ws] with ld_ws >= tokens; L fp32 [tokens] = mean_d logsumexp_k(...).
extern "C" int vitad_gmm_patch_loglik(const void* xaug, const void* packed, const float* lp2, const float* x, int ldx,
                                      float* ll_ws, int ld_ws, float* L, int tokens, int dim, int num_gaussians,
                                      void* stream) {
    VITAD_NVTX("vitad_gmm_patch_loglik");
    int rc = check_device_arch();
    if (rc) return rc;
    VITAD_REQUIRE(xaug && packed && lp2 && x && ll_ws && L, VITAD_ERR_ARG, "null pointer");
    VITAD_REQUIRE(dim == 768 && tokens > 0 && ld_ws >= tokens, VITAD_ERR_SHAPE, "dim=%d tokens=%d ld_ws=%d", dim,
                  tokens, ld_ws);
    VITAD_REQUIRE(aligned16(lp2), VITAD_ERR_ALIGN, "lp2 alignment");
    int n_kc, kc, kcv;
    rc = vitad_gmm_plan(num_gaussians, &n_kc, &kc, &kcv);
    if (rc) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    {
        ProfScope prof("gmm_fused", s);
        if (n_kc == 1 && kc == 104)
            rc = launch_mdn<104, 1>(xaug, packed, lp2, x, ldx, ll_ws, ld_ws, tokens, dim, s);
        else if (n_kc == 1)
            rc = launch_mdn<112, 1>(xaug, packed, lp2, x, ldx, ll_ws, ld_ws, tokens, dim, s);
        else if (kc == 72)
            rc = launch_mdn<72, 2>(xaug, packed, lp2, x, ldx, ll_ws, ld_ws, tokens, dim, s);
        else if (kc == 80)
            rc = launch_mdn<80, 2>(xaug, packed, lp2, x, ldx, ll_ws, ld_ws, tokens, dim, s);
        else if (kc == 88)
            rc = launch_mdn<88, 2>(xaug, packed, lp2, x, ldx, ll_ws, ld_ws, tokens, dim, s);
        else
            rc = launch_mdn<104, 2>(xaug, packed, lp2, x, ldx, ll_ws, ld_ws, tokens, dim, s);
    }
    if (rc) return rc;
    ProfScope prof2("gmm_mean", s);
    VITAD_CUDA_OK(launch_pdl(gmm_mean_kernel, dim3((tokens + 31) / 32), dim3(32 * kMeanGroups), 0, s, ll_ws, ld_ws, L, tokens, dim));
    VITAD_CUDA_OK(cudaGetLastError());
    g_launches.fetch_add(1);
    return VITAD_OK;
}

// L fp32 [batch*patches] -> prob fp32 [batch*patches] (exp(L - max over the whole batch)), scores fp32 [batch].
extern "C" int vitad_gmm_finish(const float* L, float* prob, float* scores, int batch, int patches, void* stream) {
    VITAD_NVTX("vitad_gmm_finish");
    int rc = check_device_arch();
    if (rc) return rc;
    VITAD_REQUIRE(L && prob && scores && batch > 0 && patches > 0, VITAD_ERR_ARG, "gmm_finish args");
    VITAD_CUDA_OK(launch_pdl(gmm_finish_kernel, dim3(1), dim3(1024), 0, static_cast<cudaStream_t>(stream), L, prob, scores, batch,
                             patches));
    VITAD_CUDA_OK(cudaGetLastError());
    g_launches.fetch_add(1);
    return VITAD_OK;
}
