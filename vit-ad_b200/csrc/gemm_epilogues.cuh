// Epilogue functors for the GEMM main loops: each thread owns one accumulator row (one token) and
// receives W consecutive fp32 columns at a time (W = 32, or 16 when a group's column share is not a
// multiple of 32, e.g. BLOCK_N = 96).
#pragma once
#include "gemm_core.cuh"

namespace vitad {

template <int BLOCK_N>
struct EpiChunk {
    static constexpr int kW = (BLOCK_N / 2) % 32 == 0 ? 32 : 16;
    static_assert((BLOCK_N / 2) % kW == 0, "BLOCK_N/2 must be a multiple of 16");
};

// Walk accumulator columns [c0, c1) in chunks of W and hand each chunk to f(col_in_block, v[W]).
// All lanes execute the tcgen05.ld.
template <int W, class F>
__device__ __forceinline__ void for_each_chunk(uint32_t taddr, int c0, int c1, F&& f) {
#pragma unroll 1
    for (int c = c0; c < c1; c += W) {
        uint32_t r[W];
        if constexpr (W == 32)
            tmem_ld_x32(taddr + c, r);
        else
            tmem_ld_x16(taddr + c, r);
        tmem_ld_wait();
        float v[W];
#pragma unroll
        for (int j = 0; j < W; ++j) v[j] = __uint_as_float(r[j]);
        f(c, v);
    }
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

template <int W>
__device__ __forceinline__ void load_bias(const float* __restrict__ bias, int col, float (&b)[W]) {
    const float4* p = reinterpret_cast<const float4*>(bias + col);
#pragma unroll
    for (int j = 0; j < W / 4; ++j) {
        float4 t = __ldg(p + j);
        b[4 * j + 0] = t.x;
        b[4 * j + 1] = t.y;
        b[4 * j + 2] = t.z;
        b[4 * j + 3] = t.w;
    }
}

template <int W>
__device__ __forceinline__ void store_h(__half* dst, const float (&v)[W]) {
    uint4* p = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int j = 0; j < W / 8; ++j) {
        uint4 u;
        u.x = pack_h2(v[8 * j + 0], v[8 * j + 1]);
        u.y = pack_h2(v[8 * j + 2], v[8 * j + 3]);
        u.z = pack_h2(v[8 * j + 4], v[8 * j + 5]);
        u.w = pack_h2(v[8 * j + 6], v[8 * j + 7]);
        p[j] = u;
    }
}

// out_f16[row][col] = act(acc + bias[col]);  ACT: 0 identity, 1 exact-erf GELU (timm Mlp: nn.GELU()), 2 ReLU.
template <int BLOCK_N, int ACT>
struct EpiBiasH {
    static constexpr int W = EpiChunk<BLOCK_N>::kW;
    static constexpr bool kSplitColumns = true;
    const float* bias;
    __half* out;
    int ldo, M, N;
    __device__ __forceinline__ void tile_begin(int, int, int) const {}
    __device__ __forceinline__ void tile_end(int, int, int) const {}
    __device__ __forceinline__ void sub(int, int, int n_tile, int row, uint32_t taddr, int c0, int c1) const {
        const int n0 = n_tile * BLOCK_N;
        for_each_chunk<W>(taddr, c0, c1, [&](int c, float (&v)[W]) {
            const int col = n0 + c;
            if (row < M && col < N) {
                float b[W];
                load_bias<W>(bias, col, b);
#pragma unroll
                for (int j = 0; j < W; ++j) {
                    float x = v[j] + b[j];
                    v[j] = ACT == 1 ? gelu_erf(x) : (ACT == 2 ? fmaxf(x, 0.f) : x);
                }
                store_h<W>(out + static_cast<size_t>(row) * ldo + col, v);
            }
        });
    }
};

// out_f32[row][col] = resid_f32[row][col] + acc + bias[col]   (residual stream stays fp32; out may alias resid)
template <int BLOCK_N>
struct EpiResidualF32 {
    static constexpr int W = EpiChunk<BLOCK_N>::kW;
    static constexpr bool kSplitColumns = true;
    const float* bias;
    const float* resid;
    float* out;
    int ld, M, N;
    __device__ __forceinline__ void tile_begin(int, int, int) const {}
    __device__ __forceinline__ void tile_end(int, int, int) const {}
    __device__ __forceinline__ void sub(int, int, int n_tile, int row, uint32_t taddr, int c0, int c1) const {
        const int n0 = n_tile * BLOCK_N;
        for_each_chunk<W>(taddr, c0, c1, [&](int c, float (&v)[W]) {
            const int col = n0 + c;
            if (row < M && col < N) {
                float b[W];
                load_bias<W>(bias, col, b);
                const float4* r = reinterpret_cast<const float4*>(resid + static_cast<size_t>(row) * ld + col);
                float4* o = reinterpret_cast<float4*>(out + static_cast<size_t>(row) * ld + col);
                float4 t[W / 4];
#pragma unroll
                for (int j = 0; j < W / 4; ++j) t[j] = r[j];
#pragma unroll
                for (int j = 0; j < W / 4; ++j) {
                    t[j].x += v[4 * j + 0] + b[4 * j + 0];
                    t[j].y += v[4 * j + 1] + b[4 * j + 1];
                    t[j].z += v[4 * j + 2] + b[4 * j + 2];
                    t[j].w += v[4 * j + 3] + b[4 * j + 3];
                    o[j] = t[j];
                }
            }
        });
    }
};

// Fused QKV projection epilogue (timm Attention.qkv / Swin WindowAttention.qkv + reshape/permute, and for Swin
// the cyclic shift + window_partition of SwinTransformerModule.py:360-384 folded into the store address):
//   row = b*L + t, col = which*C + h*hd + e;  (window, pos) = tok2win[t] (identity when tok2win is null)
//   q [bw][h][pos][e] = (acc+bias) * scale   (scale = hd^-0.5 folded in)
//   k [bw][h][pos][e] =  acc+bias
//   vt[bw][h][e][pos] =  acc+bias            (V stored transposed, position index contiguous, padded to Tpad,
//                                             so P@V reads a K-major B operand);   bw = b*nW + window
template <int BLOCK_N>
struct EpiQkv {
    static constexpr int W = EpiChunk<BLOCK_N>::kW;
    static constexpr bool kSplitColumns = true;
    const float* bias;
    __half* q;
    __half* k;
    __half* vt;
    const int* tok2win;       // [L] token -> window*T + pos, or null
    int M, L, T, Tpad, H, hd, nW;  // rows, tokens per image, tokens per window, padded T, heads, head dim, windows
    float scale;
    __device__ __forceinline__ void tile_begin(int, int, int) const {}
    __device__ __forceinline__ void tile_end(int, int, int) const {}
    __device__ __forceinline__ void sub(int, int, int n_tile, int row, uint32_t taddr, int c0, int c1) const {
        const int C = H * hd;
        const int n0 = n_tile * BLOCK_N;
        const int b = row / L;
        const int t = row - b * L;
        int widx = 0, pos = t;
        if (tok2win != nullptr && row < M) {
            const int wp = __ldg(tok2win + t);
            widx = wp / T;
            pos = wp - widx * T;
        }
        const size_t bw = static_cast<size_t>(b) * nW + widx;
        for_each_chunk<W>(taddr, c0, c1, [&](int c, float (&v)[W]) {
            const int col = n0 + c;
            if (row < M && col < 3 * C) {
                float bb[W];
                load_bias<W>(bias, col, bb);
                const int which = col / C;
                const int cc = col - which * C;
                const int h = cc / hd;
                const int e0 = cc - h * hd;
                const size_t bh = bw * H + h;
                if (which == 0) {
#pragma unroll
                    for (int j = 0; j < W; ++j) v[j] = (v[j] + bb[j]) * scale;
                    store_h<W>(q + (bh * T + pos) * hd + e0, v);
                } else if (which == 1) {
#pragma unroll
                    for (int j = 0; j < W; ++j) v[j] = v[j] + bb[j];
                    store_h<W>(k + (bh * T + pos) * hd + e0, v);
                } else {
                    __half* dst = vt + (bh * hd + e0) * Tpad + pos;
#pragma unroll
                    for (int j = 0; j < W; ++j) dst[static_cast<size_t>(j) * Tpad] = to_h(v[j] + bb[j]);
                }
            }
        });
    }
};

// Patch-embed epilogue (timm PatchEmbed conv-as-GEMM + pos_embed add, prefix tokens skipped):
//   row = b*P + p  ->  x[b][prefix + p][col] = acc + bias[col] + pos[prefix + p][col]
template <int BLOCK_N>
struct EpiPatchEmbed {
    static constexpr int W = EpiChunk<BLOCK_N>::kW;
    static constexpr bool kSplitColumns = true;
    const float* bias;
    const float* pos;  // [prefix+P, C]
    float* out;        // [B, prefix+P, C]
    int M, P, prefix, C;
    __device__ __forceinline__ void tile_begin(int, int, int) const {}
    __device__ __forceinline__ void tile_end(int, int, int) const {}
    __device__ __forceinline__ void sub(int, int, int n_tile, int row, uint32_t taddr, int c0, int c1) const {
        const int n0 = n_tile * BLOCK_N;
        const int b = row / P;
        const int p = row - b * P;
        for_each_chunk<W>(taddr, c0, c1, [&](int c, float (&v)[W]) {
            const int col = n0 + c;
            if (row < M && col < C) {
                float bb[W];
                load_bias<W>(bias, col, bb);
                const float4* ps = reinterpret_cast<const float4*>(pos + static_cast<size_t>(prefix + p) * C + col);
                float4* o = reinterpret_cast<float4*>(out + (static_cast<size_t>(b) * (prefix + P) + prefix + p) * C + col);
#pragma unroll
                for (int j = 0; j < W / 4; ++j) {
                    float4 t = __ldg(ps + j);
                    t.x += v[4 * j + 0] + bb[4 * j + 0];
                    t.y += v[4 * j + 1] + bb[4 * j + 1];
                    t.z += v[4 * j + 2] + bb[4 * j + 2];
                    t.w += v[4 * j + 3] + bb[4 * j + 3];
                    o[j] = t;
                }
            }
        });
    }
};

// Plain fp32 output (+ optional bias): used by tests and small projections.
template <int BLOCK_N>
struct EpiBiasF32 {
    static constexpr int W = EpiChunk<BLOCK_N>::kW;
    static constexpr bool kSplitColumns = true;
    const float* bias;  // may be null
    float* out;
    int ldo, M, N;
    __device__ __forceinline__ void tile_begin(int, int, int) const {}
    __device__ __forceinline__ void tile_end(int, int, int) const {}
    __device__ __forceinline__ void sub(int, int, int n_tile, int row, uint32_t taddr, int c0, int c1) const {
        const int n0 = n_tile * BLOCK_N;
        for_each_chunk<W>(taddr, c0, c1, [&](int c, float (&v)[W]) {
            const int col = n0 + c;
            if (row < M && col < N) {
                float* o = out + static_cast<size_t>(row) * ldo + col;
#pragma unroll
                for (int j = 0; j < W; ++j) {
                    if (col + j < N) o[j] = v[j] + (bias ? __ldg(bias + col + j) : 0.0f);
                }
            }
        });
    }
};

}  // namespace vitad
