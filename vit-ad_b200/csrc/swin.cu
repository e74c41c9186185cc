// vitad_swin_forward: the EsViT Swin-T (window 14) encoder forward as one C-ABI call
// (EncoderEsVit.forward, src/classes/transformer/TransformerEncoder.py:269-273 = SwinTransformer.forward_features,
// src/classes/transformer/SwinTransformerModule.py:821-837; blocks :349-416, PatchMerging :478-505,
// PatchEmbed :645-655).  Evaluated in inference mode (no DropPath).
//
// Window partition, cyclic shift, window reverse and the shift mask cost no HBM pass: the QKV GEMM epilogue
// scatters rows into (shifted) windows through a token->window map, the attention kernel adds the dense
// relative-position bias and the region mask on the fly and writes its output back in token order.
#include <atomic>

#include "host_util.cuh"
#include "ptx.cuh"

extern "C" int vitad_layernorm(const float*, const float*, const float*, void*, float*, int, int, int, int, int, int,
                               int, int, float, int, void*);

namespace vitad {
extern std::atomic<uint64_t> g_launches;

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// 4x4/s4 patch gathering (PatchEmbed.proj as a GEMM operand): images fp32 [B,3,S,S] -> fp16 [B*g*g, 48],
// column (c, i, j).  One thread = one (patch, c, i): 4 consecutive j.
__global__ void __launch_bounds__(256) patchify4_kernel(const float* __restrict__ img, __half* __restrict__ out, int S,
                                                        size_t total) {
    griddep_launch_dependents();
    griddep_wait();
    const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int g = S / 4;
    const int ci = static_cast<int>(idx % 12);  // c*4 + i
    const size_t row = idx / 12;
    const int c = ci >> 2, i = ci & 3;
    const int bimg = static_cast<int>(row / (g * g));
    const int p = static_cast<int>(row - static_cast<size_t>(bimg) * g * g);
    const int py = p / g, px = p - py * g;
    const float4 a = __ldg(reinterpret_cast<const float4*>(
        img + ((static_cast<size_t>(bimg) * 3 + c) * S + (py * 4 + i)) * S + px * 4));
    uint2 u;
    u.x = pack_h2(a.x, a.y);
    u.y = pack_h2(a.z, a.w);
    *reinterpret_cast<uint2*>(out + row * 48 + ci * 4) = u;
}

// PatchMerging gather + LayerNorm(4C, eps): x fp32 [B, H*H, C] -> out fp16 [B*(H/2)^2, 4C]; channel blocks in the
// reference's order x0=(0,0) x1=(1,0) x2=(0,1) x3=(1,1) (row offset, column offset).  One warp per output row.
__global__ void __launch_bounds__(256) merge_layernorm_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                              const float* __restrict__ b, __half* __restrict__ out,
                                                              int rows_out, int H, int C, float eps) {
    griddep_launch_dependents();
    griddep_wait();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= rows_out) return;
    const int H2 = H / 2;
    const int bimg = warp / (H2 * H2);
    const int p = warp - bimg * H2 * H2;
    const int y2 = p / H2, x2 = p - y2 * H2;
    const int nvq = C >> 2;      // float4 per source row
    const int nv = 4 * nvq;      // float4 per merged row (<= 384)
    float4 v[12];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 12; ++j) {
        const int i = lane + 32 * j;
        if (i < nv) {
            const int q = i / nvq, ii = i - q * nvq;
            const int dy = q & 1, dx = q >> 1;
            const float4* src = reinterpret_cast<const float4*>(
                x + (static_cast<size_t>(bimg) * H * H + (2 * y2 + dy) * H + (2 * x2 + dx)) * C);
            v[j] = src[ii];
            s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
        }
    }
    const int C4 = 4 * C;
    const float mean = warp_sum_f(s) / C4;
    float qv = 0.f;
#pragma unroll
    for (int j = 0; j < 12; ++j) {
        const int i = lane + 32 * j;
        if (i < nv) {
            const float a0 = v[j].x - mean, a1 = v[j].y - mean, a2 = v[j].z - mean, a3 = v[j].w - mean;
            qv += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
        }
    }
    const float rstd = rsqrtf(warp_sum_f(qv) / C4 + eps);
#pragma unroll
    for (int j = 0; j < 12; ++j) {
        const int i = lane + 32 * j;
        if (i < nv) {
            const float4 ww = __ldg(reinterpret_cast<const float4*>(w) + i);
            const float4 bb = __ldg(reinterpret_cast<const float4*>(b) + i);
            uint2 u;
            u.x = pack_h2((v[j].x - mean) * rstd * ww.x + bb.x, (v[j].y - mean) * rstd * ww.y + bb.y);
            u.y = pack_h2((v[j].z - mean) * rstd * ww.z + bb.z, (v[j].w - mean) * rstd * ww.w + bb.w);
            reinterpret_cast<uint2*>(out + static_cast<size_t>(warp) * C4)[i] = u;
        }
    }
}

// latent[b][c] = mean_t x_region[b][t][c]   (AdaptiveAvgPool1d(1), SwinTransformerModule.py:831-832).  Block = 64 channels x
// 4 token groups (one block per image walked all T tokens of 768 channels with 256 threads: 19.6 us for 4.8 MB); the four
// partial sums are combined in fixed order.
__global__ void __launch_bounds__(256) token_mean_kernel(const float* __restrict__ x, float* __restrict__ out, int T,
                                                         int C) {
    griddep_launch_dependents();
    griddep_wait();
    __shared__ float part[4][64];
    const int b = blockIdx.x;
    const int cl = threadIdx.x & 63, tg = threadIdx.x >> 6;
    const int c = blockIdx.y * 64 + cl;
    float s = 0.f;
    if (c < C)
        for (int t = tg; t < T; t += 4) s += x[(static_cast<size_t>(b) * T + t) * C + c];
    part[tg][cl] = s;
    __syncthreads();
    if (tg == 0 && c < C) out[static_cast<size_t>(b) * C + c] = (((part[0][cl] + part[1][cl]) + part[2][cl]) + part[3][cl]) / T;
}

namespace {
constexpr int kTokPad = 256;
struct SwinWs {
    float *x, *x2;
    void *h, *q, *k, *vt, *mlp, *patches;
    size_t vt_bytes, total;
};
SwinWs carve_swin(const vitad_swin_weights& w, int batch, void* base) {
    uint8_t* p = static_cast<uint8_t*>(base);
    size_t used = 0;
    auto take = [&](size_t bytes) {
        void* r = p ? p + used : nullptr;
        used += (bytes + 255) & ~static_cast<size_t>(255);
        return r;
    };
    const size_t g = w.img / w.patch;
    const size_t elems = static_cast<size_t>(batch) * g * g * w.embed;  // rows*C of stage 0 (halves every stage)
    SwinWs s;
    s.x = static_cast<float*>(take(elems * 4));
    s.x2 = static_cast<float*>(take(elems * 2));
    s.h = take(elems * 2);      // LN out / attention out / merged operand (4C * rows/4 = elems)
    s.q = take(elems * 2);
    s.k = take(elems * 2);
    size_t vt_max = 0;
    for (int i = 0; i < w.stages; ++i) {
        const vitad_swin_stage& st = w.stage[i];
        const size_t nW = static_cast<size_t>(st.res / st.window) * (st.res / st.window);
        const size_t v = static_cast<size_t>(batch) * nW * st.heads * (st.dim / st.heads) * kTokPad * 2;
        vt_max = v > vt_max ? v : vt_max;
    }
    s.vt_bytes = vt_max;
    s.vt = take(vt_max);
    s.mlp = take(elems * 4 * 2);
    s.patches = take(static_cast<size_t>(batch) * g * g * 48 * 2);
    s.total = used;
    return s;
}
}  // namespace
}  // namespace vitad

using namespace vitad;

extern "C" size_t vitad_swin_workspace_bytes(const vitad_swin_weights* w, int batch) {
    if (!w || batch <= 0 || !w->stage) return 0;
    return carve_swin(*w, batch, nullptr).total;
}

extern "C" int vitad_swin_forward(const vitad_swin_weights* wp, const float* images, int batch, void* workspace,
                                  size_t workspace_bytes, float* out_tokens, float* out_latent, void* out_xaug,
                                  int ld_xaug, void* stream) {
    VITAD_NVTX("vitad_swin_forward");
    int rc = check_device_arch();
    if (rc) return rc;
    VITAD_REQUIRE(wp && images && workspace && out_tokens && wp->stage, VITAD_ERR_ARG, "null pointer");
    const vitad_swin_weights& w = *wp;
    VITAD_REQUIRE(w.patch == 4 && w.img % 4 == 0 && w.embed % 32 == 0 && w.stages >= 1 && w.stages <= 4,
                  VITAD_ERR_SHAPE, "unsupported Swin geometry (patch %d embed %d stages %d)", w.patch, w.embed, w.stages);
    VITAD_REQUIRE(batch > 0, VITAD_ERR_SHAPE, "empty batch");
    VITAD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, VITAD_ERR_ALIGN, "workspace alignment");
    SwinWs ws = carve_swin(w, batch, workspace);
    VITAD_REQUIRE(workspace_bytes >= ws.total, VITAD_ERR_WORKSPACE, "workspace %zu < %zu", workspace_bytes, ws.total);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int g = w.img / w.patch;

    // patch embedding: 4x4 conv as GEMM (K = 48) + bias, then LayerNorm(embed) in place (patch_norm)
    {
        const size_t total = static_cast<size_t>(batch) * g * g * 12;
        VITAD_CUDA_OK(launch_pdl(patchify4_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, s, images,
                                 static_cast<__half*>(ws.patches), w.img, total));
        g_launches.fetch_add(1);
    }
    vitad_linear_args a;
    memset(&a, 0, sizeof(a));
    a.a = ws.patches, a.w = w.patch_w, a.bias = w.patch_b, a.m = batch * g * g, a.n = w.embed, a.k = 48;
    a.lda = 48, a.ldw = 48, a.epilogue = VITAD_EPI_F32, a.out = ws.x, a.ldo = w.embed;
    if ((rc = vitad_linear_f16(&a, s))) return rc;
    if ((rc = vitad_layernorm(ws.x, w.patch_ln_w, w.patch_ln_b, nullptr, ws.x, a.m, w.embed, w.embed, 0, w.embed, a.m,
                              a.m, 0, 1e-5f, 0, s)))
        return rc;

    float* x = ws.x;
    float* xalt = ws.x2;
    int cur_T = -1;
    for (int si = 0; si < w.stages; ++si) {
        const vitad_swin_stage& st = w.stage[si];
        const int C = st.dim, H = st.res, L = H * H, rows = batch * L;
        const int nWside = H / st.window, nW = nWside * nWside, T = st.window * st.window, hd = C / st.heads;
        VITAD_REQUIRE(hd == 32 && H % st.window == 0 && T <= 208 && st.blocks, VITAD_ERR_SHAPE,
                      "stage %d: head dim %d / window %d / res %d unsupported", si, hd, st.window, H);
        if (T != cur_T) {  // padded key columns of the transposed-V buffer must read as zero
            VITAD_CUDA_OK(cudaMemsetAsync(ws.vt, 0, ws.vt_bytes, s));
            cur_T = T;
        }
        for (int bi = 0; bi < st.depth; ++bi) {
            const vitad_swin_block& B = st.blocks[bi];
            const bool shifted = B.shift > 0;
            VITAD_REQUIRE(!shifted || (st.tok2win[1] && st.win2tok[1] && st.region), VITAD_ERR_ARG,
                          "shifted block without maps");
            const int* t2w = nW > 1 ? st.tok2win[shifted ? 1 : 0] : nullptr;
            const int* w2t = nW > 1 ? st.win2tok[shifted ? 1 : 0] : nullptr;
            VITAD_REQUIRE(nW == 1 || (t2w && w2t), VITAD_ERR_ARG, "stage %d needs window maps", si);
            if ((rc = vitad_layernorm(x, B.ln1_w, B.ln1_b, ws.h, nullptr, rows, C, C, C, 0, rows, rows, 0, 1e-5f, 0, s)))
                return rc;
            memset(&a, 0, sizeof(a));
            a.a = ws.h, a.w = B.qkv_w, a.bias = B.qkv_b, a.m = rows, a.n = 3 * C, a.k = C, a.lda = C, a.ldw = C;
            a.epilogue = VITAD_EPI_QKV, a.q = ws.q, a.kmat = ws.k, a.vt = ws.vt;
            a.tokens = L, a.tokens_pad = kTokPad, a.heads = st.heads, a.q_scale = rsqrtf(static_cast<float>(hd));
            a.head_dim = hd, a.windows = nW, a.win_tokens = T, a.tok2win = t2w;
            if ((rc = vitad_linear_f16(&a, s))) return rc;
            vitad_attention_args at;
            memset(&at, 0, sizeof(at));
            at.q = ws.q, at.k = ws.k, at.vt = ws.vt, at.out = ws.h;
            at.batch_windows = batch * nW, at.heads = st.heads, at.tokens = T, at.tokens_pad = kTokPad;
            at.head_dim = hd, at.windows = nW, at.bias = B.attn_bias, at.region = shifted ? st.region : nullptr;
            at.win2tok = w2t;
            if ((rc = vitad_attention_f16(&at, s))) return rc;
            memset(&a, 0, sizeof(a));
            a.a = ws.h, a.w = B.proj_w, a.bias = B.proj_b, a.m = rows, a.n = C, a.k = C, a.lda = C, a.ldw = C;
            a.epilogue = VITAD_EPI_RESIDUAL_F32, a.out = x, a.resid = x, a.ldo = C;
            if ((rc = vitad_linear_f16(&a, s))) return rc;
            if ((rc = vitad_layernorm(x, B.ln2_w, B.ln2_b, ws.h, nullptr, rows, C, C, C, 0, rows, rows, 0, 1e-5f, 0, s)))
                return rc;
            memset(&a, 0, sizeof(a));
            a.a = ws.h, a.w = B.fc1_w, a.bias = B.fc1_b, a.m = rows, a.n = 4 * C, a.k = C, a.lda = C, a.ldw = C;
            a.epilogue = VITAD_EPI_BIAS_GELU_F16, a.out = ws.mlp, a.ldo = 4 * C;
            if ((rc = vitad_linear_f16(&a, s))) return rc;
            memset(&a, 0, sizeof(a));
            a.a = ws.mlp, a.w = B.fc2_w, a.bias = B.fc2_b, a.m = rows, a.n = C, a.k = 4 * C, a.lda = 4 * C, a.ldw = 4 * C;
            a.epilogue = VITAD_EPI_RESIDUAL_F32, a.out = x, a.resid = x, a.ldo = C;
            if ((rc = vitad_linear_f16(&a, s))) return rc;
        }
        if (st.merge_w) {  // PatchMerging: gather 2x2 -> LayerNorm(4C) -> Linear(4C -> 2C, no bias)
            const int rows_out = rows / 4;
            VITAD_CUDA_OK(launch_pdl(merge_layernorm_kernel, dim3((rows_out + 7) / 8), dim3(256), 0, s,
                                     static_cast<const float*>(x), st.merge_ln_w, st.merge_ln_b, static_cast<__half*>(ws.h),
                                     rows_out, H, C, 1e-5f));
            g_launches.fetch_add(1);
            memset(&a, 0, sizeof(a));
            a.a = ws.h, a.w = st.merge_w, a.bias = nullptr, a.m = rows_out, a.n = 2 * C, a.k = 4 * C, a.lda = 4 * C,
            a.ldw = 4 * C;
            a.epilogue = VITAD_EPI_F32, a.out = xalt, a.ldo = 2 * C;
            if ((rc = vitad_linear_f16(&a, s))) return rc;
            float* t = x;
            x = xalt;
            xalt = t;
        }
    }
    const vitad_swin_stage& last = w.stage[w.stages - 1];
    const int Lf = last.res * last.res, Cf = last.dim;
    if ((rc = vitad_layernorm(x, w.norm_w, w.norm_b, out_xaug, out_tokens, batch * Lf, Cf, Cf, ld_xaug, Cf, Lf, Lf, 0,
                              1e-5f, out_xaug ? 2 : 0, s)))
        return rc;
    if (out_latent) {
        VITAD_CUDA_OK(launch_pdl(token_mean_kernel, dim3(batch, (Cf + 63) / 64), dim3(256), 0, s,
                                 static_cast<const float*>(out_tokens), out_latent, Lf, Cf));
        g_launches.fetch_add(1);
    }
    return VITAD_OK;
}
