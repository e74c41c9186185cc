// LayerNorm arithmetic shared by the fused residual GEMM + LayerNorm kernel (gemm_ln.cuh: one THREAD per row) and the
// standalone kernel the encoder uses for the same LayerNorms when the fused kernel does not pay (few rows) or has no
// producer GEMM (the first block's norm1): layernorm768_tree_kernel, one WARP per row.  Both evaluate the SAME expression
// tree with explicitly rounded operations (no contraction the compiler could apply differently), so a row's result is
// bit-identical whichever kernel — and therefore whichever batch size — produced it:
//   unit (32 columns):  s = xor-butterfly sum of the 32 values (lane c holds column c; a thread emulates the butterfly),
//                       mu = s / 32,  q = butterfly sum of (x - mu)^2
//   row  (768 columns): four partials of 6 units each, merged left to right with Chan's update, then (P0 + P1) + (P2 + P3)
//   y = ((x - mean) * rstd) * gamma + beta,  rstd = rsqrt(M2 / 768 + eps)
#pragma once
#include "ptx.cuh"

namespace vitad {

constexpr int kTreeC = 768;
constexpr int kTreeUnits = kTreeC / 32;        // 24
constexpr int kTreePartUnits = 6;              // units per partial (= per epilogue warp of the fused kernel's default config)

// Chan et al.: merge (n_b, mean_b, M2_b) into (n_a, mean_a, M2_a); every operation explicitly rounded
__device__ __forceinline__ void tree_merge(float& mean_a, float& m2_a, float n_a, float mean_b, float m2_b, float n_b) {
    const float delta = __fsub_rn(mean_b, mean_a);
    const float n = n_a + n_b;                 // exact small integers
    mean_a = __fmaf_rn(delta, __fdiv_rn(n_b, n), mean_a);
    const float w = __fdiv_rn(__fmul_rn(n_a, n_b), n);
    m2_a = __fadd_rn(__fadd_rn(m2_a, m2_b), __fmul_rn(__fmul_rn(delta, delta), w));
}

// butterfly sum of 32 values held by one thread: the association a warp's xor-shuffle reduction produces
__device__ __forceinline__ float tree_sum32(float (&t)[32]) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int c = 0; c < o; ++c) t[c] = __fadd_rn(t[c], t[c + o]);
    return t[0];
}
__device__ __forceinline__ float tree_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// (mean, M2) of one 32-column unit held by one thread
__device__ __forceinline__ void tree_unit_stats(const float (&x)[32], float& mu, float& q) {
    float t[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) t[c] = x[c];
    mu = __fmul_rn(tree_sum32(t), 1.0f / 32.0f);
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        const float d = __fsub_rn(x[c], mu);
        t[c] = __fmul_rn(d, d);
    }
    q = tree_sum32(t);
}
// the same for a unit spread over a warp (lane c holds column c); every lane gets the result
__device__ __forceinline__ void tree_unit_stats_warp(float xv, float& mu, float& q) {
    mu = __fmul_rn(tree_warp_sum(xv), 1.0f / 32.0f);
    const float d = __fsub_rn(xv, mu);
    q = tree_warp_sum(__fmul_rn(d, d));
}

__device__ __forceinline__ float tree_rstd(float m2, float eps) { return rsqrtf(__fadd_rn(__fmul_rn(m2, 1.0f / kTreeC), eps)); }
__device__ __forceinline__ float tree_normalize(float x, float mean, float rstd, float gamma, float beta) {
    return __fmaf_rn(__fmul_rn(__fsub_rn(x, mean), rstd), gamma, beta);
}

}  // namespace vitad
