// Bandwidth-bound row-wise kernels of the encoder path: LayerNorm (fp32 residual stream -> fp16 GEMM
// operand and/or fp32), patch gathering (im2col for the 16x16/s16 patch-embed conv), prefix tokens.
// One warp per row, 128-bit loads, warp-shuffle reductions; grids sized in whole waves of 148 SMs
// where the row count allows.
#include <atomic>

#include "host_util.cuh"
#include "ln_tree.cuh"
#include "ptx.cuh"

namespace vitad {
extern std::atomic<uint64_t> g_launches;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// y = (x - mean) / sqrt(var + eps) * w + b over the last dim C (biased variance, as nn.LayerNorm).
// Output row r reads input row  (r / out_tokens) * in_tokens + skip + (r % out_tokens)  so the final
// norm can drop the prefix tokens (x[:, 2:, :], TransformerEncoder.py:168) while normalising.
// out_h: fp16 [rows, ldh]; columns [C, C+aug_ones) are set to 1 and [C+aug_ones, ldh) to 0 when
// aug_ones > 0 (the MDN GEMM folds its biases into two extra K columns).  out_f: fp32 [rows, ldf].
// One warp normalises kRows rows at once: all their loads are issued before the first reduction (the kernel is
// latency-bound: one row per warp left each SM with ~3 KB per warp in flight and needed 1.07 waves of blocks),
// and the affine parameters are fetched once per warp.
constexpr int kLnRows = 2;
template <int MAXV>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ b, __half* __restrict__ out_h,
                                                        float* __restrict__ out_f, int rows, int C, int ldx, int ldh,
                                                        int ldf, int in_tokens, int out_tokens, int skip, float eps,
                                                        int aug_ones) {
    griddep_launch_dependents();
    griddep_wait();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int row0 = warp * kLnRows;
    if (row0 >= rows) return;
    const int nv = C >> 2;
    float4 v[kLnRows][MAXV];
    float s[kLnRows];
#pragma unroll
    for (int r = 0; r < kLnRows; ++r) {
        const int row = min(row0 + r, rows - 1);  // a clamped duplicate row is computed and not stored
        const int bi = row / out_tokens;
        const int in_row = bi * in_tokens + skip + (row - bi * out_tokens);
        const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(in_row) * ldx);
        s[r] = 0.f;
#pragma unroll
        for (int j = 0; j < MAXV; ++j) {
            const int i = lane + 32 * j;
            if (i < nv) {
                v[r][j] = xr[i];
                s[r] += (v[r][j].x + v[r][j].y) + (v[r][j].z + v[r][j].w);
            }
        }
    }
    float mean[kLnRows], rstd[kLnRows];
#pragma unroll
    for (int r = 0; r < kLnRows; ++r) mean[r] = warp_sum(s[r]) / C;
#pragma unroll
    for (int r = 0; r < kLnRows; ++r) {
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < MAXV; ++j) {
            const int i = lane + 32 * j;
            if (i < nv) {
                const float a = v[r][j].x - mean[r], bb = v[r][j].y - mean[r], c = v[r][j].z - mean[r],
                            d = v[r][j].w - mean[r];
                q += (a * a + bb * bb) + (c * c + d * d);
            }
        }
        rstd[r] = rsqrtf(warp_sum(q) / C + eps);
    }
    const float4* wr = reinterpret_cast<const float4*>(w);
    const float4* br = reinterpret_cast<const float4*>(b);
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
        const int i = lane + 32 * j;
        if (i < nv) {
            const float4 ww = __ldg(wr + i), bb = __ldg(br + i);
#pragma unroll
            for (int r = 0; r < kLnRows; ++r) {
                const int row = row0 + r;
                if (row >= rows) break;
                float4 y;
                y.x = (v[r][j].x - mean[r]) * rstd[r] * ww.x + bb.x;
                y.y = (v[r][j].y - mean[r]) * rstd[r] * ww.y + bb.y;
                y.z = (v[r][j].z - mean[r]) * rstd[r] * ww.z + bb.z;
                y.w = (v[r][j].w - mean[r]) * rstd[r] * ww.w + bb.w;
                if (out_f) reinterpret_cast<float4*>(out_f + static_cast<size_t>(row) * ldf)[i] = y;
                if (out_h) {
                    uint2 u;
                    u.x = pack_h2(y.x, y.y);
                    u.y = pack_h2(y.z, y.w);
                    reinterpret_cast<uint2*>(out_h + static_cast<size_t>(row) * ldh)[i] = u;
                }
            }
        }
    }
    if (out_h && aug_ones > 0) {
#pragma unroll
        for (int r = 0; r < kLnRows; ++r) {
            const int row = row0 + r;
            if (row >= rows) break;
            for (int c = C + lane; c < ldh; c += 32)
                out_h[static_cast<size_t>(row) * ldh + c] = to_h(c < C + aug_ones ? 1.0f : 0.0f);
        }
    }
}

// Narrow rows (Swin stages 0/1: C = 96, 192 over 100 352 / 25 088 rows): LPR lanes per row, 32/LPR rows per warp pass and
// kLnRows passes per warp, every lane holds VPL float4 of its row (C = 4*LPR*VPL exactly).  The one-warp-per-two-rows
// kernel above leaves 8 of 32 lanes idle at C = 96 and needs 6272 blocks (8 waves) for stage 0.
template <int LPR, int VPL>
__global__ void __launch_bounds__(256) layernorm_narrow_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                               const float* __restrict__ b, __half* __restrict__ out_h,
                                                               float* __restrict__ out_f, int rows, int ldx, int ldh,
                                                               int ldf, float eps) {
    griddep_launch_dependents();
    griddep_wait();
    constexpr int C = 4 * LPR * VPL;
    constexpr int kRowsPerPass = 32 / LPR;
    constexpr int kPasses = 4;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int sub = lane / LPR, l = lane % LPR;
    const int row0 = warp * kRowsPerPass * kPasses + sub;
    float4 v[kPasses][VPL];
    float s[kPasses];
#pragma unroll
    for (int p = 0; p < kPasses; ++p) {
        const int row = min(row0 + p * kRowsPerPass, rows - 1);
        const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * ldx);
        s[p] = 0.f;
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
            v[p][j] = xr[l + LPR * j];
            s[p] += (v[p][j].x + v[p][j].y) + (v[p][j].z + v[p][j].w);
        }
    }
    float mean[kPasses], rstd[kPasses];
#pragma unroll
    for (int p = 0; p < kPasses; ++p) {
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) s[p] += __shfl_xor_sync(0xffffffffu, s[p], o);
        mean[p] = s[p] / C;
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
            const float a = v[p][j].x - mean[p], bb = v[p][j].y - mean[p], c = v[p][j].z - mean[p], d = v[p][j].w - mean[p];
            q += (a * a + bb * bb) + (c * c + d * d);
        }
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        rstd[p] = rsqrtf(q / C + eps);
    }
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
        const int i = l + LPR * j;
        const float4 ww = __ldg(reinterpret_cast<const float4*>(w) + i), bb = __ldg(reinterpret_cast<const float4*>(b) + i);
#pragma unroll
        for (int p = 0; p < kPasses; ++p) {
            const int row = row0 + p * kRowsPerPass;
            if (row >= rows) continue;
            float4 y;
            y.x = (v[p][j].x - mean[p]) * rstd[p] * ww.x + bb.x;
            y.y = (v[p][j].y - mean[p]) * rstd[p] * ww.y + bb.y;
            y.z = (v[p][j].z - mean[p]) * rstd[p] * ww.z + bb.z;
            y.w = (v[p][j].w - mean[p]) * rstd[p] * ww.w + bb.w;
            if (out_f) reinterpret_cast<float4*>(out_f + static_cast<size_t>(row) * ldf)[i] = y;
            if (out_h) {
                uint2 u;
                u.x = pack_h2(y.x, y.y);
                u.y = pack_h2(y.z, y.w);
                reinterpret_cast<uint2*>(out_h + static_cast<size_t>(row) * ldh)[i] = u;
            }
        }
    }
}

// Patch gathering for the non-overlapping PxP/sP conv (timm PatchEmbed.proj): images fp32 [B,Cin,S,S]
// -> A fp16 [B*g*g, Cin*P*P], column order (c, i, j) = the conv weight's flattened order.
// One thread moves 8 consecutive j (two float4 loads, one 16-byte store).
__global__ void __launch_bounds__(256) patchify_kernel(const float* __restrict__ img, __half* __restrict__ out, int B,
                                                       int Cin, int S, int P, size_t total8) {
    griddep_launch_dependents();
    griddep_wait();
    const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total8) return;
    const int g = S / P;
    const int kcols = Cin * P * P;
    const int per_row = kcols >> 3;
    const size_t row = idx / per_row;
    const int col = static_cast<int>(idx - row * per_row) << 3;
    const int c = col / (P * P);
    const int rem = col - c * P * P;
    const int i = rem / P, j = rem - i * P;
    const int bimg = static_cast<int>(row / (g * g));
    const int p = static_cast<int>(row - static_cast<size_t>(bimg) * g * g);
    const int py = p / g, px = p - py * g;
    const float* src = img + ((static_cast<size_t>(bimg) * Cin + c) * S + (py * P + i)) * S + px * P + j;
    const float4 a = __ldg(reinterpret_cast<const float4*>(src));
    const float4 b2 = __ldg(reinterpret_cast<const float4*>(src) + 1);
    uint4 u;
    u.x = pack_h2(a.x, a.y);
    u.y = pack_h2(a.z, a.w);
    u.z = pack_h2(b2.x, b2.y);
    u.w = pack_h2(b2.z, b2.w);
    *reinterpret_cast<uint4*>(out + row * kcols + col) = u;
}

// Same gather from uint8 images (the dataset's native pixels): value = u / 255.0f, exactly torchvision's ToTensor
// (GeneralDataset.py:46-53), then the fp16 rounding of the operand; a batch crosses PCIe as 4.8 MB instead of 19.3 MB.
__global__ void __launch_bounds__(256) patchify_u8_kernel(const uint8_t* __restrict__ img, __half* __restrict__ out,
                                                          int B, int Cin, int S, int P, size_t total8) {
    griddep_launch_dependents();
    griddep_wait();
    const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total8) return;
    const int g = S / P;
    const int kcols = Cin * P * P;
    const int per_row = kcols >> 3;
    const size_t row = idx / per_row;
    const int col = static_cast<int>(idx - row * per_row) << 3;
    const int c = col / (P * P);
    const int rem = col - c * P * P;
    const int i = rem / P, j = rem - i * P;
    const int bimg = static_cast<int>(row / (g * g));
    const int p = static_cast<int>(row - static_cast<size_t>(bimg) * g * g);
    const int py = p / g, px = p - py * g;
    const uint8_t* src = img + ((static_cast<size_t>(bimg) * Cin + c) * S + (py * P + i)) * S + px * P + j;
    const uint2 raw = __ldg(reinterpret_cast<const uint2*>(src));  // 8 pixels (px*P + j is a multiple of 8)
    float v[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        v[k] = static_cast<float>((raw.x >> (8 * k)) & 0xffu) / 255.0f;
        v[4 + k] = static_cast<float>((raw.y >> (8 * k)) & 0xffu) / 255.0f;
    }
    uint4 u;
    u.x = pack_h2(v[0], v[1]);
    u.y = pack_h2(v[2], v[3]);
    u.z = pack_h2(v[4], v[5]);
    u.w = pack_h2(v[6], v[7]);
    *reinterpret_cast<uint4*>(out + row * kcols + col) = u;
}

// x[b][t][:] = tok[t][:] + pos[t][:] for the `prefix` leading tokens (cls, dist) of every image.
__global__ void prefix_tokens_kernel(const float* __restrict__ tok, const float* __restrict__ pos,
                                     float* __restrict__ x, int B, int prefix, int T, int C) {
    griddep_launch_dependents();
    griddep_wait();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = B * prefix * C;
    if (idx >= total) return;
    const int c = idx % C;
    const int t = (idx / C) % prefix;
    const int b = idx / (C * prefix);
    x[(static_cast<size_t>(b) * T + t) * C + c] = tok[t * C + c] + pos[t * C + c];
}

}  // namespace vitad

using namespace vitad;

// LayerNorm over C = 768 with the fused residual-GEMM kernel's arithmetic (ln_tree.cuh), one warp per row: lane c holds
// column 32 u + c of unit u, so a unit's butterfly sum is the warp's xor-shuffle reduction.  Bit-identical to what
// gemm_ln_kernel writes for the same row, which lets the encoder choose between the two by row count without changing a bit.
__global__ void __launch_bounds__(256) layernorm768_tree_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                const float* __restrict__ b, __half* __restrict__ out_h,
                                                                int rows, int ldx, int ldh, float eps) {
    griddep_launch_dependents();
    griddep_wait();
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + static_cast<size_t>(row) * ldx + lane;
    float v[kTreeUnits];
#pragma unroll
    for (int u = 0; u < kTreeUnits; ++u) v[u] = xr[32 * u];
    float pm[4], pq[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
#pragma unroll
        for (int i = 0; i < kTreePartUnits; ++i) {
            float mu, q;
            tree_unit_stats_warp(v[p * kTreePartUnits + i], mu, q);
            if (i == 0) {
                pm[p] = mu;
                pq[p] = q;
            } else {
                tree_merge(pm[p], pq[p], 32.0f * i, mu, q, 32.0f);
            }
        }
    }
    constexpr float kPart = 32.0f * kTreePartUnits;
    tree_merge(pm[0], pq[0], kPart, pm[1], pq[1], kPart);
    tree_merge(pm[2], pq[2], kPart, pm[3], pq[3], kPart);
    tree_merge(pm[0], pq[0], 2 * kPart, pm[2], pq[2], 2 * kPart);
    const float mean = pm[0], rstd = tree_rstd(pq[0], eps);
    __half* hr = out_h + static_cast<size_t>(row) * ldh + lane;
#pragma unroll
    for (int u = 0; u < kTreeUnits; ++u)
        hr[32 * u] = to_h(tree_normalize(v[u], mean, rstd, __ldg(w + 32 * u + lane), __ldg(b + 32 * u + lane)));
}

// x fp32 [rows, ldx] -> out fp16 [rows, ldh], C = 768, the arithmetic of vitad_linear_resid_ln_f16's LayerNorm half
extern "C" int vitad_layernorm768_tree(const float* x, const float* weight, const float* bias, void* out_f16, int rows,
                                       int ldx, int ldh, float eps, void* stream) {
    int rc = check_device_arch();
    if (rc) return rc;
    VITAD_REQUIRE(x && weight && bias && out_f16 && rows > 0 && ldx >= kTreeC && ldh >= kTreeC, VITAD_ERR_ARG,
                  "layernorm768_tree arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ProfScope prof("layernorm_tree", s);
    VITAD_CUDA_OK(launch_pdl(layernorm768_tree_kernel, dim3((rows + 7) / 8), dim3(256), 0, s, x, weight, bias,
                             static_cast<__half*>(out_f16), rows, ldx, ldh, eps));
    VITAD_CUDA_OK(cudaGetLastError());
    g_launches.fetch_add(1);
    return VITAD_OK;
}

extern "C" int vitad_layernorm(const float* x, const float* weight, const float* bias, void* out_f16, float* out_f32,
                               int rows, int c, int ldx, int ld_f16, int ld_f32, int in_tokens, int out_tokens,
                               int skip, float eps, int aug_ones, void* stream) {
    int rc = check_device_arch();
    if (rc) return rc;
    VITAD_REQUIRE(x && weight && bias && (out_f16 || out_f32), VITAD_ERR_ARG, "null pointer");
    VITAD_REQUIRE(rows > 0 && c > 0 && c % 4 == 0 && c <= 1536, VITAD_ERR_SHAPE, "C=%d must be %%4 and <= 1536", c);
    VITAD_REQUIRE(ldx % 4 == 0 && aligned16(x) && aligned16(weight) && aligned16(bias), VITAD_ERR_ALIGN,
                  "layernorm input alignment");
    VITAD_REQUIRE(!out_f16 || (ld_f16 % 4 == 0 && ld_f16 >= c + aug_ones && (reinterpret_cast<uintptr_t>(out_f16) & 7) == 0),
                  VITAD_ERR_ALIGN, "fp16 output pitch/alignment");
    VITAD_REQUIRE(!out_f32 || (ld_f32 % 4 == 0 && ld_f32 >= c && aligned16(out_f32)), VITAD_ERR_ALIGN,
                  "fp32 output pitch/alignment");
    VITAD_REQUIRE(in_tokens > 0 && out_tokens > 0 && skip >= 0 && skip + out_tokens <= in_tokens, VITAD_ERR_SHAPE,
                  "token remap");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ProfScope prof("layernorm", s);
    // narrow rows without token remap / operand augmentation: sub-warp lane groups per row
    const bool plain = in_tokens == out_tokens && skip == 0 && aug_ones == 0;
    if (plain && (c == 96 || c == 192)) {
        const int rows_per_warp = (c == 96 ? 4 : 2) * 4;
        const int nblocks = (rows + 8 * rows_per_warp - 1) / (8 * rows_per_warp);
        if (c == 96)
            VITAD_CUDA_OK(launch_pdl(layernorm_narrow_kernel<8, 3>, dim3(nblocks), dim3(256), 0, s, x, weight, bias,
                                     static_cast<__half*>(out_f16), out_f32, rows, ldx, ld_f16, ld_f32, eps));
        else
            VITAD_CUDA_OK(launch_pdl(layernorm_narrow_kernel<16, 3>, dim3(nblocks), dim3(256), 0, s, x, weight, bias,
                                     static_cast<__half*>(out_f16), out_f32, rows, ldx, ld_f16, ld_f32, eps));
        VITAD_CUDA_OK(cudaGetLastError());
        g_launches.fetch_add(1);
        return VITAD_OK;
    }
    const int blocks = (rows + 8 * kLnRows - 1) / (8 * kLnRows);
    if (c <= 768)
        VITAD_CUDA_OK(launch_pdl(layernorm_kernel<6>, dim3(blocks), dim3(256), 0, s, x, weight, bias,
                                 static_cast<__half*>(out_f16), out_f32, rows, c, ldx, ld_f16, ld_f32, in_tokens, out_tokens,
                                 skip, eps, aug_ones));
    else
        VITAD_CUDA_OK(launch_pdl(layernorm_kernel<12>, dim3(blocks), dim3(256), 0, s, x, weight, bias,
                                 static_cast<__half*>(out_f16), out_f32, rows, c, ldx, ld_f16, ld_f32, in_tokens, out_tokens,
                                 skip, eps, aug_ones));
    VITAD_CUDA_OK(cudaGetLastError());
    g_launches.fetch_add(1);
    return VITAD_OK;
}

extern "C" int vitad_patchify(const float* images, void* out_f16, int batch, int channels, int size, int patch,
                              void* stream) {
    int rc = check_device_arch();
    if (rc) return rc;
    VITAD_REQUIRE(images && out_f16, VITAD_ERR_ARG, "null pointer");
    VITAD_REQUIRE(batch > 0 && channels > 0 && patch % 8 == 0 && size % patch == 0, VITAD_ERR_SHAPE,
                  "patchify needs patch %% 8 == 0 and size %% patch == 0");
    VITAD_REQUIRE(aligned16(images) && aligned16(out_f16), VITAD_ERR_ALIGN, "patchify alignment");
    const int g = size / patch;
    const size_t total8 = static_cast<size_t>(batch) * g * g * channels * patch * patch / 8;
    const unsigned blocks = static_cast<unsigned>((total8 + 255) / 256);
    VITAD_CUDA_OK(launch_pdl(patchify_kernel, dim3(blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), images,
                             static_cast<__half*>(out_f16), batch, channels, size, patch, total8));
    VITAD_CUDA_OK(cudaGetLastError());
    g_launches.fetch_add(1);
    return VITAD_OK;
}

extern "C" int vitad_patchify_u8(const uint8_t* images, void* out_f16, int batch, int channels, int size, int patch,
                                 void* stream) {
    int rc = check_device_arch();
    if (rc) return rc;
    VITAD_REQUIRE(images && out_f16, VITAD_ERR_ARG, "null pointer");
    VITAD_REQUIRE(batch > 0 && channels > 0 && patch % 8 == 0 && size % patch == 0, VITAD_ERR_SHAPE,
                  "patchify needs patch %% 8 == 0 and size %% patch == 0");
    VITAD_REQUIRE((reinterpret_cast<uintptr_t>(images) & 7) == 0 && aligned16(out_f16), VITAD_ERR_ALIGN, "patchify_u8 alignment");
    const int g = size / patch;
    const size_t total8 = static_cast<size_t>(batch) * g * g * channels * patch * patch / 8;
    const unsigned blocks = static_cast<unsigned>((total8 + 255) / 256);
    VITAD_CUDA_OK(launch_pdl(patchify_u8_kernel, dim3(blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), images,
                             static_cast<__half*>(out_f16), batch, channels, size, patch, total8));
    VITAD_CUDA_OK(cudaGetLastError());
    g_launches.fetch_add(1);
    return VITAD_OK;
}

extern "C" int vitad_prefix_tokens(const float* tokens, const float* pos, float* x, int batch, int prefix, int t, int c,
                                   void* stream) {
    int rc = check_device_arch();
    if (rc) return rc;
    VITAD_REQUIRE(tokens && pos && x && batch > 0 && prefix > 0 && prefix <= t, VITAD_ERR_ARG, "prefix tokens args");
    const int total = batch * prefix * c;
    VITAD_CUDA_OK(launch_pdl(prefix_tokens_kernel, dim3((total + 255) / 256), dim3(256), 0, static_cast<cudaStream_t>(stream),
                             tokens, pos, x, batch, prefix, t, c));
    VITAD_CUDA_OK(cudaGetLastError());
    g_launches.fetch_add(1);
    return VITAD_OK;
}
