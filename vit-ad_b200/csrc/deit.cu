// vitad_deit_forward: the whole DeiT-B distilled 16/224 encoder forward as one C-ABI call
// (EncoderDeit.forward, src/classes/transformer/TransformerEncoder.py:145-173; timm 0.6.13
// VisionTransformerDistilled.forward_features).  Per block: LayerNorm -> QKV GEMM (head-major epilogue)
// -> fused attention -> proj GEMM (+residual) -> LayerNorm -> fc1 GEMM (+GELU) -> fc2 GEMM (+residual).
// The residual stream stays fp32 in HBM; GEMM operands are fp16.  In the standard forward (block_index 0) both
// LayerNorms of a block run inside the residual GEMM that produces their input (vitad_linear_resid_ln_f16, gemm_ln.cuh):
// proj also emits norm2(x), fc2 also emits the NEXT block's norm1(x) — 26 LayerNorm launches become 3.
#include <atomic>

#include "host_util.cuh"

namespace vitad {
extern std::atomic<int> g_fused_ln;  // gemm_ln.cu: residual GEMM + LayerNorm in one kernel (default on)
extern std::atomic<int> g_use_pair;  // host_util.cu: CTA-pair GEMM kernels (default on)
std::atomic<int> g_v_natural{1};     // V stored like K for the attention kernel's MN-major operand path (default on)
}

extern "C" void vitad_set_v_natural(int enable) { vitad::g_v_natural.store(enable ? 1 : 0); }

extern "C" int vitad_layernorm(const float*, const float*, const float*, void*, float*, int, int, int, int, int, int,
                               int, int, float, int, void*);
extern "C" int vitad_layernorm768_tree(const float*, const float*, const float*, void*, int, int, int, float, void*);
extern "C" int vitad_patchify(const float*, void*, int, int, int, int, void*);
extern "C" int vitad_patchify_u8(const uint8_t*, void*, int, int, int, int, void*);
extern "C" int vitad_prefix_tokens(const float*, const float*, float*, int, int, int, int, void*);

namespace {
constexpr int kTokPad = 256;  // key padding of the transposed-V buffer

struct Carve {
    uint8_t* p;
    size_t used = 0;
    void* take(size_t bytes) {
        void* r = p ? p + used : nullptr;
        used += (bytes + 255) & ~static_cast<size_t>(255);
        return r;
    }
};

struct DeitWs {
    float* x;
    void *h, *q, *k, *vt, *mlp, *patches;
    size_t vt_bytes, total;
};

DeitWs carve_ws(const vitad_deit_weights& w, int batch, void* base) {
    Carve c{static_cast<uint8_t*>(base)};
    const size_t rows = static_cast<size_t>(batch) * w.tokens;
    const int hd = w.dim / w.heads;
    DeitWs s;
    s.x = static_cast<float*>(c.take(rows * w.dim * 4));
    s.h = c.take(rows * w.dim * 2);
    s.q = c.take(rows * w.dim * 2);
    s.k = c.take(rows * w.dim * 2);
    s.vt_bytes = static_cast<size_t>(batch) * w.heads * hd * kTokPad * 2;
    s.vt = c.take(s.vt_bytes);
    s.mlp = c.take(rows * w.hidden * 2);
    s.patches = c.take(static_cast<size_t>(batch) * (w.tokens - w.prefix) * 3 * w.patch * w.patch * 2);
    s.total = c.used;
    return s;
}
}  // namespace

extern "C" size_t vitad_deit_workspace_bytes(const vitad_deit_weights* w, int batch) {
    if (!w || batch <= 0) return 0;
    return carve_ws(*w, batch, nullptr).total;
}

static int deit_forward_impl(const vitad_deit_weights* wp, const void* images, bool images_u8, int batch, int block_index,
                             void* workspace, size_t workspace_bytes, float* out_tokens, float* out_cls, void* out_xaug,
                             int ld_xaug, void* stream);

extern "C" int vitad_deit_forward(const vitad_deit_weights* wp, const float* images, int batch, int block_index,
                                  void* workspace, size_t workspace_bytes, float* out_tokens, float* out_cls,
                                  void* out_xaug, int ld_xaug, void* stream) {
    VITAD_NVTX("vitad_deit_forward");
    return deit_forward_impl(wp, images, false, batch, block_index, workspace, workspace_bytes, out_tokens, out_cls,
                             out_xaug, ld_xaug, stream);
}

extern "C" int vitad_deit_forward_u8(const vitad_deit_weights* wp, const uint8_t* images, int batch, int block_index,
                                     void* workspace, size_t workspace_bytes, float* out_tokens, float* out_cls,
                                     void* out_xaug, int ld_xaug, void* stream) {
    VITAD_NVTX("vitad_deit_forward_u8");
    return deit_forward_impl(wp, images, true, batch, block_index, workspace, workspace_bytes, out_tokens, out_cls,
                             out_xaug, ld_xaug, stream);
}

static int deit_forward_impl(const vitad_deit_weights* wp, const void* images, bool images_u8, int batch, int block_index,
                             void* workspace, size_t workspace_bytes, float* out_tokens, float* out_cls, void* out_xaug,
                             int ld_xaug, void* stream) {
    using namespace vitad;
    int rc = check_device_arch();
    if (rc) return rc;
    VITAD_REQUIRE(wp && images && workspace && out_tokens, VITAD_ERR_ARG, "null pointer");
    const vitad_deit_weights& w = *wp;
    VITAD_REQUIRE(w.dim == 768 && w.heads == 12 && w.patch == 16 && w.img % w.patch == 0 && w.depth >= 1 &&
                      w.depth <= VITAD_DEIT_MAX_DEPTH && w.layers,
                  VITAD_ERR_SHAPE, "unsupported DeiT geometry (dim %d heads %d patch %d depth %d)", w.dim, w.heads,
                  w.patch, w.depth);
    const int g = w.img / w.patch, P = g * g, T = P + w.prefix;
    VITAD_REQUIRE(T == w.tokens && T <= 208, VITAD_ERR_SHAPE, "tokens=%d inconsistent with img/patch/prefix", w.tokens);
    VITAD_REQUIRE(block_index >= 0 && block_index < w.depth, VITAD_ERR_SHAPE, "block_index %d out of range", block_index);
    VITAD_REQUIRE(batch > 0, VITAD_ERR_SHAPE, "empty batch");
    VITAD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, VITAD_ERR_ALIGN, "workspace must be 256-byte aligned");
    DeitWs ws = carve_ws(w, batch, workspace);
    VITAD_REQUIRE(workspace_bytes >= ws.total, VITAD_ERR_WORKSPACE, "workspace %zu < %zu bytes", workspace_bytes, ws.total);
    VITAD_REQUIRE(!out_xaug || ld_xaug >= w.dim + 16, VITAD_ERR_SHAPE, "xaug pitch");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int rows = batch * T, C = w.dim;

    // V in its natural layout (stored like K, consumed as an MN-major operand by the attention kernel) whenever the QKV
    // projection runs on the CTA-pair kernel; the transposed, zero-padded form otherwise
    const bool v_nat = vitad::g_v_natural.load() != 0 && vitad::g_use_pair.load() != 0 && rows > 128 && C / w.heads == 64;
    // patch embedding: gather patches -> GEMM with (+bias +pos_embed) epilogue into x[:, prefix:, :]
    if (!v_nat) VITAD_CUDA_OK(cudaMemsetAsync(ws.vt, 0, ws.vt_bytes, s));
    if (images_u8)
        rc = vitad_patchify_u8(static_cast<const uint8_t*>(images), ws.patches, batch, 3, w.img, w.patch, s);
    else
        rc = vitad_patchify(static_cast<const float*>(images), ws.patches, batch, 3, w.img, w.patch, s);
    if (rc) return rc;
    vitad_linear_args a;
    memset(&a, 0, sizeof(a));
    a.a = ws.patches, a.w = w.patch_w, a.bias = w.patch_b;
    a.m = batch * P, a.n = C, a.k = 3 * w.patch * w.patch, a.lda = a.k, a.ldw = a.k;
    a.epilogue = VITAD_EPI_PATCH_EMBED, a.out = ws.x, a.ldo = C, a.pos = w.pos, a.patches = P, a.prefix = w.prefix;
    if ((rc = vitad_linear_f16(&a, s))) return rc;
    if ((rc = vitad_prefix_tokens(w.prefix_tokens, w.pos, ws.x, batch, w.prefix, T, C, s))) return rc;

    const int last = block_index != 0 ? block_index : w.depth - 1;
    // block_index != 0 re-normalises the stream after every block (below), which breaks the fc2 -> next norm1 chain
    // The fused kernel runs one 4-CTA cluster per 256 rows: it pays once the batch fills most of the 33 cluster slots
    // (batch 32: 25 clusters); below that the separate launches spread the same work over more SMs (batch 1: 0.73 vs
    // 0.95 ms per image).  Both forms compute every row with the same arithmetic (ln_tree.cuh), bit for bit.
    const bool fused = vitad::g_fused_ln.load() != 0 && block_index == 0 && C == 768 && rows >= 16 * 256;
    auto block_norm = [&](const float* gw, const float* gb) {  // a block's norm1 / norm2 as its own launch
        return C == 768 ? vitad_layernorm768_tree(ws.x, gw, gb, ws.h, rows, C, C, 1e-6f, s)
                        : vitad_layernorm(ws.x, gw, gb, ws.h, nullptr, rows, C, C, C, 0, rows, rows, 0, 1e-6f, 0, s);
    };
    vitad_linear_ln_args f;
    for (int i = 0; i <= last; ++i) {
        const vitad_deit_layer& L = w.layers[i];
        if (!fused || i == 0)
            if ((rc = block_norm(L.ln1_w, L.ln1_b))) return rc;
        memset(&a, 0, sizeof(a));
        a.a = ws.h, a.w = L.qkv_w, a.bias = L.qkv_b, a.m = rows, a.n = 3 * C, a.k = C, a.lda = C, a.ldw = C;
        a.epilogue = VITAD_EPI_QKV, a.q = ws.q, a.kmat = ws.k, a.vt = ws.vt;
        a.tokens = T, a.tokens_pad = kTokPad, a.heads = w.heads, a.q_scale = 0.125f, a.v_natural = v_nat;
        if ((rc = vitad_linear_f16(&a, s))) return rc;
        vitad_attention_args at;
        memset(&at, 0, sizeof(at));
        at.q = ws.q, at.k = ws.k, at.out = ws.h;
        if (v_nat) at.v = ws.vt; else at.vt = ws.vt;
        at.batch_windows = batch, at.heads = w.heads, at.tokens = T, at.tokens_pad = kTokPad;
        at.head_dim = C / w.heads, at.windows = 1;
        if ((rc = vitad_attention_f16(&at, s))) return rc;
        if (fused) {
            // x += proj(attention) + bias;  h = norm2(x).  In place on ws.h: a cluster reads its 256 rows of the attention
            // output through the whole K loop before its epilogue overwrites the same rows with the normalised stream
            memset(&f, 0, sizeof(f));
            f.a = ws.h, f.w = L.proj_w, f.bias = L.proj_b, f.m = rows, f.k = C, f.lda = C, f.ldw = C;
            f.x = ws.x, f.gamma = L.ln2_w, f.beta = L.ln2_b, f.eps = 1e-6f, f.h = ws.h, f.ldh = C;
            if ((rc = vitad_linear_resid_ln_f16(&f, s))) return rc;
        } else {
            memset(&a, 0, sizeof(a));
            a.a = ws.h, a.w = L.proj_w, a.bias = L.proj_b, a.m = rows, a.n = C, a.k = C, a.lda = C, a.ldw = C;
            a.epilogue = VITAD_EPI_RESIDUAL_F32, a.out = ws.x, a.resid = ws.x, a.ldo = C;
            if ((rc = vitad_linear_f16(&a, s))) return rc;
            if ((rc = block_norm(L.ln2_w, L.ln2_b))) return rc;
        }
        memset(&a, 0, sizeof(a));
        a.a = ws.h, a.w = L.fc1_w, a.bias = L.fc1_b, a.m = rows, a.n = w.hidden, a.k = C, a.lda = C, a.ldw = C;
        a.epilogue = VITAD_EPI_BIAS_GELU_F16, a.out = ws.mlp, a.ldo = w.hidden;
        if ((rc = vitad_linear_f16(&a, s))) return rc;
        if (fused && i < last) {
            // x += fc2(gelu(fc1)) + bias;  h = norm1 of the NEXT block
            const vitad_deit_layer& Ln = w.layers[i + 1];
            memset(&f, 0, sizeof(f));
            f.a = ws.mlp, f.w = L.fc2_w, f.bias = L.fc2_b, f.m = rows, f.k = w.hidden, f.lda = w.hidden, f.ldw = w.hidden;
            f.x = ws.x, f.gamma = Ln.ln1_w, f.beta = Ln.ln1_b, f.eps = 1e-6f, f.h = ws.h, f.ldh = C;
            if ((rc = vitad_linear_resid_ln_f16(&f, s))) return rc;
            continue;
        }
        memset(&a, 0, sizeof(a));
        a.a = ws.mlp, a.w = L.fc2_w, a.bias = L.fc2_b, a.m = rows, a.n = C, a.k = w.hidden, a.lda = w.hidden,
        a.ldw = w.hidden;
        a.epilogue = VITAD_EPI_RESIDUAL_F32, a.out = ws.x, a.resid = ws.x, a.ldo = C;
        if ((rc = vitad_linear_f16(&a, s))) return rc;
        // block_index != 0: the final norm is applied after EVERY block, in place (TransformerEncoder.py:161-163)
        if (block_index != 0 && i < last)
            if ((rc = vitad_layernorm(ws.x, w.norm_w, w.norm_b, nullptr, ws.x, rows, C, C, 0, C, rows, rows, 0, 1e-6f, 0, s)))
                return rc;
    }
    // final norm straight into the outputs: patch tokens (prefix dropped, :168), MDN operand, cls token (:169)
    if ((rc = vitad_layernorm(ws.x, w.norm_w, w.norm_b, out_xaug, out_tokens, batch * P, C, C, ld_xaug, C, T, P,
                              w.prefix, 1e-6f, out_xaug ? 2 : 0, s)))
        return rc;
    if (out_cls)
        if ((rc = vitad_layernorm(ws.x, w.norm_w, w.norm_b, nullptr, out_cls, batch, C, C, 0, C, T, 1, 0, 1e-6f, 0, s)))
            return rc;
    return VITAD_OK;
}
