/*
 * vitad.h — C ABI of the B200-native scoring path of Miwri/vit-ad.
 *
 * The reference has no FFI of its own (it is pure Python/PyTorch); the boundary it offers is the
 * set of Python classes its scripts and Validators call.  Every entry point below
 * names the reference code it replaces (path:line relative to the reference root) so a maintainer
 * can bind it from those classes (see INTEGRATION.md for the ctypes stubs).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless it says "host";
 *   - the caller owns every buffer (inputs, outputs, workspaces); nothing here calls cudaMalloc;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *   - return value: VITAD_OK (0) or a negative vitad_status; vitad_last_error() has the message;
 *   - sm_100a only: on any other device every compute entry point returns VITAD_ERR_ARCH.
 *     There is no CPU path.
 */
#ifndef VITAD_H_
#define VITAD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum vitad_status {
    VITAD_OK = 0,
    VITAD_ERR_SHAPE = -1,   /* unsupported or inconsistent shape */
    VITAD_ERR_ALIGN = -2,   /* pointer or pitch not 16-byte aligned */
    VITAD_ERR_ARG = -3,     /* null pointer / bad enum */
    VITAD_ERR_ARCH = -4,    /* not an sm_100 device */
    VITAD_ERR_CUDA = -5,    /* CUDA runtime/driver call failed */
    VITAD_ERR_WORKSPACE = -6 /* workspace too small */
} vitad_status;

const char* vitad_last_error(void);
int vitad_abi_version(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
uint64_t vitad_launch_count(void);
/* GEMM-class kernels run on CTA pairs (tcgen05 cta_group::2, 256-row tiles) by default; 0 selects the
 * single-CTA 128-row kernels (kept for A/B measurements and for problems of <= 128 rows). */
void vitad_set_cta_pair(int enable);
/* Programmatic dependent launch between the kernels of the scoring chain (default on; 0 = plain stream order). */
void vitad_set_pdl(int enable);
/* Fused GMM kernel on clusters of four CTAs sharing the token block by TMA multicast (default on; 0 = CTA pairs). */
void vitad_set_gmm_cluster4(int enable);
/* 148 SMs hold 33 clusters of four, so 16 SMs idle under the 4-CTA kernel: the last `features` features of the fused GMM
 * projection run concurrently on those SMs as the CTA-pair kernel, launched on the same stream as an independent
 * programmatic dependent of the 4-CTA grid (it starts once that grid is fully resident and waits for it before exiting;
 * needs vitad_set_pdl(1)).  -1 = balanced split (default, 72 of 768 on a B200), 0 = off. */
void vitad_set_gmm_split(int features);
/* Diagnostics: force 8 or 16 epilogue warps in the CTA-pair GEMM kernels (0 = per-epilogue default). */
void vitad_set_epilogue_warps(int warps);
/* Optional in-library profiler: CUDA events around every launch site of this library.
 * vitad_profile_enable(1) clears and starts recording, vitad_profile_report() synchronises the device and
 * writes "name count total_us" lines (+ a "__span__" line: first start .. last end) into buf. */
void vitad_profile_enable(int on);
int vitad_profile_report(char* buf, int size);

/* ------------------------------------------------------------------------------------------
 * Dense projection  D = A · Wᵀ (+ fused epilogue), fp16 operands (IEEE half: same tensor-core rate as bf16, 3 more mantissa bits), fp32 accumulate (tcgen05/TMEM,
 * TMA-fed).  Replaces every nn.Linear / conv-as-GEMM on the encoder path:
 *   timm Attention.qkv / Attention.proj / Mlp.fc1 / Mlp.fc2 / PatchEmbed.proj, called through
 *   src/classes/transformer/TransformerEncoder.py:150-165 (EncoderDeit.forward).
 * A: [M,K] fp16 row-major (pitch lda elements), W: [N,K] fp16 row-major (nn.Linear layout).
 * K % 16 == 0, N % 32 == 0 (N % 8 for VITAD_EPI_F32), pitches % 8 == 0.
 * ------------------------------------------------------------------------------------------ */
typedef enum vitad_epilogue {
    VITAD_EPI_BIAS_F16 = 0,      /* out_f16 = acc + bias                                  */
    VITAD_EPI_BIAS_GELU_F16 = 1, /* out_f16 = gelu_erf(acc + bias)   (timm Mlp.act)       */
    VITAD_EPI_RESIDUAL_F32 = 2,   /* out_f32  = resid_f32 + acc + bias (timm Block residual) */
    VITAD_EPI_QKV = 3,            /* head-major q (pre-scaled), k, transposed v             */
    VITAD_EPI_PATCH_EMBED = 4,    /* + bias + pos_embed, written behind the prefix tokens   */
    VITAD_EPI_F32 = 5,            /* out_f32 = acc (+ bias if non-null)                     */
    VITAD_EPI_BIAS_RELU_F16 = 6,  /* out_f16 = relu(acc + bias)  (FastFlow subnet, NormalizingFlow.py:61-82) */
    VITAD_EPI_CONVT_RELU_F16 = 7, /* ConvTranspose2d(k3,s2,p1,op1)+BN+ReLU as one GEMM (CnnDecoder.py:47-117): row =
                                     input pixel (b,i,j) of a convt_w-wide grid, col = (phase a, phase c, channel);
                                     out_f16 NHWC [B,2H,2W,N/4] = relu(acc + bias) scattered to pixel (2i+a, 2j+c) */
    VITAD_EPI_RES16_RELU_F16 = 8, /* out_f16 = relu(acc + bias + resid16): the tail of a reverse-ResNet Bottleneck
                                     (ReverseResNet.py:86-103: conv1 + bn1, `out += identity`, relu).  res_grid = 0: the
                                     residual row is the output row; res_grid = g > 0: the identity path is a stride-2
                                     1x1 transposed convolution computed on the g x g input grid (:190-195), so output
                                     pixel (b,y,x) of the 2g x 2g grid adds row (b,y/2,x/2) when y and x are even and
                                     nothing otherwise (its BatchNorm shift is folded into bias) */
    VITAD_EPI_TANH_PIX4_F32 = 9   /* reverse-ResNet image head (CnnDecoder.py:189-194): row = pixel (b,J,I) of a
                                     convt_w-wide grid, col = c*16 + py*4 + px (48 live of 64); out fp32 NCHW
                                     [B,3,4W,4W][b][c][4J+py][4I+px] = tanh(acc + bias[col]) */
} vitad_epilogue;

typedef struct vitad_linear_args {
    const void* a;      /* fp16 [M,K] */
    const void* w;      /* fp16 [N,K] */
    const float* bias;  /* fp32 [N]   */
    int m, n, k;
    int lda, ldw;       /* pitches in elements */
    int epilogue;       /* vitad_epilogue */
    int block_n;        /* tile-width hint: 0 = library default, else 128, 192 or 256 (96: single-CTA kernel only) */
    void* out;          /* fp16 or fp32 [M,ldo], see epilogue */
    int ldo;
    const float* resid; /* RESIDUAL_F32: fp32 [M,ldo] (may alias out) */
    /* QKV: q,k fp16 [B,H,T,64]; vt fp16 [B,H,64,Tpad]; M = B*T, N = 3*H*64 */
    void* q;
    void* kmat;
    void* vt;
    int tokens, tokens_pad, heads;
    float q_scale;
    /* PATCH_EMBED: out fp32 [B,prefix+P,N]; pos fp32 [prefix+P,N]; M = B*P */
    const float* pos;
    int patches, prefix;
    /* QKV, optional (0 = DeiT defaults: head_dim 64, one window of `tokens`): Swin window attention.
     * q,k become [B*windows,H,win_tokens,head_dim], vt [B*windows,H,head_dim,tokens_pad]; tok2win int32 [tokens]
     * maps a token of the image to window*win_tokens + position (cyclic shift + window_partition,
     * SwinTransformerModule.py:360-384). */
    int head_dim, windows, win_tokens;
    const int* tok2win;
    int convt_w;        /* CONVT_RELU_F16 / TANH_PIX4_F32: width (= height) of the input pixel grid */
    const void* resid16; /* RES16_RELU_F16: fp16 residual rows, pitch ldr elements */
    int ldr, res_grid;
    /* Implicit 3x3 convolution (stride 1, zero padding 1) without an im2col pass, for BIAS_F16 / BIAS_RELU_F16 /
     * TANH_PIX4_F32.  conv_grid = g > 0: `a` is the zero-bordered NHWC activation [B, g+2, g+2, C] (pitch lda >= C,
     * C % 64 == 0), k = 9*C with W columns ordered (ty, tx, c), m = B*(g+2)*(g+2) GEMM rows of which the B*g*g interior
     * ones are written as plain pixel rows.  out_pad_grid = g > 0 (BIAS_*_F16, RES16_RELU_F16): the GEMM's m = B*g*g
     * plain pixel rows are written into such a zero-bordered layout (the caller zeroes the border once). */
    int conv_grid, out_pad_grid;
    /* Split-fp16 operands: fp32-grade products on the fp16 tensor cores (used by the reverse-ResNet decoder, whose 53
     * chained layers otherwise leave 2e-3 on the L2 map).  A value v is stored as hi = fp16(v), lo = fp16(v - hi).
     * split_c = C > 0 (C % 64 == 0): every tap of `a` holds [hi (C) | lo (C)] columns and every tap of `w`
     * [w_hi (C) | w_hi (C) | w_lo (C)], k = taps * 3 * C; the K loop pairs them as hi*w_hi + lo*w_hi + hi*w_lo (each
     * product exact in the fp32 accumulator; the dropped lo*w_lo term is 2^-22 relative).
     * a_taps > 1: `a` is an explicit im2col layout of that many taps (columns tap-major); taps = 9 with conv_grid.
     * split_out = 1 (BIAS_F16, BIAS_RELU_F16, CONVT_RELU_F16, RES16_RELU_F16): the epilogue writes [hi (N) | lo (N)]
     * rows (ldo >= 2N; per pixel for CONVT), and RES16 reads its residual in the same form (ldr >= 2N). */
    int split_c, a_taps, split_out;
    /* QKV: 1 = `vt` receives V in its natural layout [B*windows, H, win_tokens, head_dim] (like k) instead of the
     * transposed, padded form; pairs with vitad_attention_args.v. */
    int v_natural;
} vitad_linear_args;

int vitad_linear_f16(const vitad_linear_args* args, void* stream);

/* Residual projection + LayerNorm in one kernel (N = 768), the tail of both halves of a timm Block
 * (`x = x + drop_path(attn(norm1(x)))` / `x = x + drop_path(mlp(norm2(x)))` followed by the next norm; call sites
 * TransformerEncoder.py:150-165):
 *     x[r,:] += a[r,:] . w^T + bias          fp32 [m, 768] residual stream, in place (pitch 768)
 *     h[r,:]  = LayerNorm(x[r,:]) * gamma + beta   fp16 [m, ldh]
 * a fp16 [m, k] (pitch lda), w fp16 [768, k] (pitch ldw), k % 64 == 0.  A cluster of four CTAs owns 256 rows over all
 * 768 columns, so the row statistics never leave the chip (csrc/gemm_ln.cuh). */
typedef struct vitad_linear_ln_args {
    const void* a;
    const void* w;
    const float* bias;
    int m, k, lda, ldw;
    float* x;
    const float* gamma;
    const float* beta;
    float eps;
    void* h;
    int ldh;
} vitad_linear_ln_args;
int vitad_linear_resid_ln_f16(const vitad_linear_ln_args* args, void* stream);
/* Encoder forward: fuse each block's two LayerNorms into the preceding residual GEMMs when the batch has enough rows
 * to fill the 4-CTA clusters (default 1; 0 = always the separate vitad_linear_f16(RESIDUAL_F32) +
 * vitad_layernorm768_tree launches — the results are bit-identical either way). */
void vitad_set_fused_ln(int enable);
/* Encoder forward: V in its natural layout, consumed as an MN-major operand (default 1; 0 = transposed padded vT). */
void vitad_set_v_natural(int enable);
/* LayerNorm over C = 768 (x fp32 [rows, ldx] -> out fp16 [rows, ldh]) with exactly the arithmetic of the LayerNorm half of
 * vitad_linear_resid_ln_f16 (csrc/ln_tree.cuh): bit-identical rows, so the encoder can pick the fused kernel or the
 * separate launches by row count without changing a result. */
int vitad_layernorm768_tree(const float* x, const float* weight, const float* bias, void* out_f16, int rows, int ldx,
                            int ldh, float eps, void* stream);

/* ------------------------------------------------------------------------------------------
 * Row-wise encoder kernels (HBM-bound).
 * vitad_layernorm: nn.LayerNorm over the last dim (timm Block.norm1/norm2 and VisionTransformer.norm,
 *   eps 1e-6; Swin uses 1e-5).  x fp32 [*, c] with pitch ldx.  Output row r reads input row
 *   (r / out_tokens) * in_tokens + skip + r % out_tokens, so the final norm can drop the prefix tokens
 *   (x[:, 2:, :], TransformerEncoder.py:168).  out_f16 (pitch ld_f16) and/or out_f32 (pitch ld_f32).
 *   aug_ones > 0: fp16 columns [c, c+aug_ones) = 1 and the rest of the pitch = 0 (MDN bias columns).
 * vitad_patchify: images fp32 [B,C,S,S] -> fp16 [B*(S/P)^2, C*P*P], column order (c,i,j) = flattened
 *   Conv2d weight (timm PatchEmbed.proj as a GEMM operand).
 * vitad_prefix_tokens: x[b][t][:] = tokens[t] + pos[t] for t < prefix (cls, dist tokens).
 * ------------------------------------------------------------------------------------------ */
int vitad_layernorm(const float* x, const float* weight, const float* bias, void* out_f16, float* out_f32,
                    int rows, int c, int ldx, int ld_f16, int ld_f32, int in_tokens, int out_tokens, int skip,
                    float eps, int aug_ones, void* stream);
int vitad_patchify(const float* images, void* out_f16, int batch, int channels, int size, int patch, void* stream);
/* Same from uint8 pixels: value = u / 255 (torchvision ToTensor, GeneralDataset.py:46-53) before the fp16 rounding. */
int vitad_patchify_u8(const uint8_t* images, void* out_f16, int batch, int channels, int size, int patch, void* stream);
int vitad_prefix_tokens(const float* tokens, const float* pos, float* x, int batch, int prefix, int t, int c,
                        void* stream);

/* Fused softmax(Q K^T + bias + mask) V on tcgen05.
 *   DeiT: timm Attention.forward (windows = 1, bias/region/win2tok null).
 *   Swin: WindowAttention.forward + shift mask + window_reverse/roll (SwinTransformerModule.py:144-193,316-347,
 *         392-408).
 * q,k fp16 [BW,H,T,hd] (q pre-scaled), vt fp16 [BW,H,hd,tokens_pad] (zero beyond T), BW = batch*windows;
 * out fp16 [batch*windows*T, H*hd] in ORIGINAL token order: row = b*(windows*T) + win2tok[window*T + pos].
 * bias fp32 [H,T_key,T_query] (dense relative-position bias, key-major so query lanes read contiguous memory)
 * or null; region int8 [windows,T] region label per window
 * position (scores between different labels get -100) or null.  T <= 208, hd in {32, 64}. */
typedef struct vitad_attention_args {
    const void *q, *k, *vt;
    void* out;
    int batch_windows, heads, tokens, tokens_pad, head_dim, windows;
    const float* bias;
    const signed char* region;
    const int* win2tok;
    /* Alternative to vt (exactly one of the two): V in its natural layout fp16 [batch_windows, heads, tokens, 64], the
     * layout of k (head_dim 64, no bias / region).  It enters P.V as an MN-major tensor-core operand, so the QKV
     * projection (vitad_linear_args.v_natural) stores V like K and no padded transposed buffer exists. */
    const void* v;
} vitad_attention_args;
int vitad_attention_f16(const vitad_attention_args* args, void* stream);

/* ------------------------------------------------------------------------------------------
 * Whole DeiT-B distilled encoder forward = EncoderDeit.forward
 * (src/classes/transformer/TransformerEncoder.py:145-173).  All weight pointers are device pointers;
 * matrices are fp16 in nn.Linear layout [out,in] (patch_w = conv weight flattened to [dim, 3*P*P]),
 * vectors fp32.  block_index follows the reference: 0 = all blocks + final norm; i != 0 = blocks 0..i
 * with the final norm applied after every block.
 *   out_tokens fp32 [B, P, dim] (= patch_embedding), out_cls fp32 [B, dim] (= latent_space, may be null),
 *   out_xaug   fp16 [B*P, ld_xaug] (may be null): normalised tokens + two 1-columns, the MDN GEMM operand.
 * ------------------------------------------------------------------------------------------ */
#define VITAD_DEIT_MAX_DEPTH 24
typedef struct vitad_deit_layer {
    const float *ln1_w, *ln1_b;
    const void* qkv_w;
    const float* qkv_b;
    const void* proj_w;
    const float* proj_b;
    const float *ln2_w, *ln2_b;
    const void* fc1_w;
    const float* fc1_b;
    const void* fc2_w;
    const float* fc2_b;
} vitad_deit_layer;

typedef struct vitad_deit_weights {
    int img, patch, dim, heads, hidden, depth, tokens, prefix;
    const void* patch_w;
    const float* patch_b;
    const float* prefix_tokens; /* [prefix, dim]: cls, dist */
    const float* pos;           /* [tokens, dim] */
    const float *norm_w, *norm_b;
    const vitad_deit_layer* layers; /* host array of `depth` entries */
} vitad_deit_weights;

size_t vitad_deit_workspace_bytes(const vitad_deit_weights* w, int batch);
int vitad_deit_forward(const vitad_deit_weights* w, const float* images, int batch, int block_index, void* workspace,
                       size_t workspace_bytes, float* out_tokens, float* out_cls, void* out_xaug, int ld_xaug,
                       void* stream);
/* Same forward from uint8 images [B,3,S,S] (the /255 of ToTensor is folded into the patch gather). */
int vitad_deit_forward_u8(const vitad_deit_weights* w, const uint8_t* images, int batch, int block_index, void* workspace,
                          size_t workspace_bytes, float* out_tokens, float* out_cls, void* out_xaug, int ld_xaug,
                          void* stream);

/* ------------------------------------------------------------------------------------------
 * Whole EsViT Swin-T encoder forward = EncoderEsVit.forward (TransformerEncoder.py:269-273) =
 * SwinTransformer.forward_features (SwinTransformerModule.py:821-837), inference mode.
 * Matrices fp16 [out,in], vectors fp32.  Per block: attn_bias fp32 [heads,T_key,T_query] =
 * relative_position_bias_table gathered through relative_position_index (:169-178), key-major; shift = 0 or
 * window/2.  Per stage: window maps for the
 * unshifted [0] and shifted [1] partition (null when the stage is a single window) and the region labels of
 * create_attn_mask (:316-347) for the shifted partition; merge_* = PatchMerging (null on the last stage).
 *   out_tokens fp32 [B,49,768] (x_region = patch_embedding), out_latent fp32 [B,768] (avg-pool, may be null),
 *   out_xaug fp16 [B*49, ld_xaug] (may be null): MDN GEMM operand.
 * ------------------------------------------------------------------------------------------ */
typedef struct vitad_swin_block {
    const float *ln1_w, *ln1_b;
    const void* qkv_w;
    const float* qkv_b;
    const float* attn_bias;
    const void* proj_w;
    const float* proj_b;
    const float *ln2_w, *ln2_b;
    const void* fc1_w;
    const float* fc1_b;
    const void* fc2_w;
    const float* fc2_b;
    int shift;
} vitad_swin_block;

typedef struct vitad_swin_stage {
    int dim, heads, res, window, depth;
    const vitad_swin_block* blocks; /* host array */
    const int* tok2win[2];          /* int32 [res*res]: token -> window*T + pos */
    const int* win2tok[2];          /* int32 [res*res]: window*T + pos -> token  */
    const signed char* region;      /* int8 [windows*T], shifted partition */
    const float *merge_ln_w, *merge_ln_b;
    const void* merge_w;            /* fp16 [2*dim, 4*dim] or null */
} vitad_swin_stage;

typedef struct vitad_swin_weights {
    int img, patch, embed, stages;
    const void* patch_w;            /* fp16 [embed, 3*patch*patch] */
    const float *patch_b, *patch_ln_w, *patch_ln_b;
    const float *norm_w, *norm_b;
    const vitad_swin_stage* stage;  /* host array */
} vitad_swin_weights;

size_t vitad_swin_workspace_bytes(const vitad_swin_weights* w, int batch);
int vitad_swin_forward(const vitad_swin_weights* w, const float* images, int batch, void* workspace,
                       size_t workspace_bytes, float* out_tokens, float* out_latent, void* out_xaug, int ld_xaug,
                       void* stream);

/* ------------------------------------------------------------------------------------------
 * Mixture-density ("GMM") head: GaussianMixtureDensityNetwork.forward + log_likelihood +
 * get_probability_map (src/classes/MixtureDensityNetwork.py:35-97,151-171) and the score tail of
 * ValidatorMdn.valid_loop_transformer (src/pipeline/ValidatorMDN.py:133-172).
 *   vitad_gmm_plan            mixture chunking used by the packed layouts (n_kc chunks of kc slots, kcv valid)
 *   vitad_gmm_pack_weights    sigma/mu weights+biases (fp32, reference layout [dim*K, dim], k fastest) -> packed fp16
 *   vitad_gmm_log_pi          lp2[t][slot] = log2(softmax(pi(x)+gumbel)+1e-15); gumbel fp32 [tokens,K] is an
 *                             explicit input (the reference draws it inside gumbel_softmax, :62)
 *   vitad_gmm_patch_loglik    L[t] = mean_d logsumexp_k(log pi + log N(x_d; mu_dk, sigma_dk))   (:49-72,:86-88)
 *   vitad_gmm_finish          prob = exp(L - max over the whole batch) (:90-95); scores[b] = 1 - min_p prob
 * ------------------------------------------------------------------------------------------ */
int vitad_gmm_plan(int num_gaussians, int* n_kc, int* kc, int* kcv);
size_t vitad_gmm_packed_weight_bytes(int dim, int num_gaussians);
int vitad_gmm_pack_weights(const float* sigma_w, const float* sigma_b, const float* mu_w, const float* mu_b, int dim,
                           int num_gaussians, void* packed, void* stream);
/* xaug fp16 [tokens,784] = (fp16(x), 1, 1, 0...) from plain fp32 features x [tokens, ldx]. */
int vitad_gmm_make_operand(const float* x, int ldx, void* xaug, int tokens, int dim, void* stream);
int vitad_gmm_log_pi(const float* x, int ldx, const float* pi_w, const float* pi_b, const float* gumbel, float* lp2,
                     int tokens, int dim, int num_gaussians, void* stream);
/* The same with the Gumbel noise generated inside the kernel instead of read from memory.  The reference draws it from
 * torch's global generator in every call (F.gumbel_softmax, MixtureDensityNetwork.py:62), so its scores depend on the
 * call order; here element (token t, mixture k) of global batch `batch_index` gets
 *   g = -log(-log(u)),  u = ((w >> 9) + 0.5) * 2^-23  (exact in fp32, never 0 or 1),
 *   w = word (k/32)%4 of Philox4x32-10(key = seed, counter = (t, k%32, k/128, batch_index)),
 * a pure function of its arguments: a batch scores the same on whichever rank and in whichever order it runs
 * (ValidatorMdn, rank/world_size).  vitad_gumbel_noise writes that noise as fp32 [tokens, num_gaussians]. */
int vitad_gmm_log_pi_seeded(const float* x, int ldx, const float* pi_w, const float* pi_b, uint64_t seed,
                            uint32_t batch_index, float* lp2, int tokens, int dim, int num_gaussians, void* stream);
int vitad_gumbel_noise(uint64_t seed, uint32_t batch_index, float* out, int tokens, int num_gaussians, void* stream);
/* The same on the tensor cores (split-fp16 operands x = xh + xl, W = Wh + Wl): pi_packed fp16 [K, 3*dim] from
 * vitad_gmm_pack_pi (pi.weight fp32 [K, dim]); workspace of vitad_gmm_log_pi_workspace_bytes, 256-byte aligned.
 * 2x faster, but the tensor core's truncating fp32 accumulation is ~5x less accurate than vitad_gmm_log_pi
 * (1.1e-4 vs 1.9e-5 on the log2-probabilities, measured); the Python head uses vitad_gmm_log_pi. */
size_t vitad_gmm_pi_packed_bytes(int dim, int num_gaussians);
int vitad_gmm_pack_pi(const float* pi_w, int dim, int num_gaussians, void* packed, void* stream);
size_t vitad_gmm_log_pi_workspace_bytes(int tokens, int dim, int num_gaussians);
int vitad_gmm_log_pi_tc(const float* x, int ldx, const void* pi_packed, const float* pi_b, const float* gumbel, float* lp2,
                        int tokens, int dim, int num_gaussians, void* workspace, size_t workspace_bytes, void* stream);
int vitad_gmm_patch_loglik(const void* xaug, const void* packed, const float* lp2, const float* x, int ldx,
                           float* ll_ws, int ld_ws, float* L, int tokens, int dim, int num_gaussians, void* stream);
int vitad_gmm_finish(const float* L, float* prob, float* scores, int batch, int patches, void* stream);

/* ------------------------------------------------------------------------------------------
 * Normalizing-flow (FastFlow) head: NormalizingFlow.forward (src/classes/NormalizingFlow.py:118-145) over
 * FrEIA 0.2 AllInOneBlock steps (:84-116).  Per-step device pointers (packed by the host mirror,
 * vitad/nf.py):
 *   w0p fp16 [64, taps*384]   first subnet conv, rows = hidden channels (61 padded to 64), cols (tap, c_in)
 *   b0p fp32 [64]
 *   w2p fp16 [768, taps*64]   second subnet conv x 0.1, rows interleaved per 96-row tile: 48 s-channels, 48 t-channels
 *   b2p fp32 [768]            same order, x 0.1
 *   scale, offset fp32 [768]  0.1*softplus_{beta=.5}(global_scale), global_offset
 *   inv_perm int32 [768]      y channel j is stored at stream row inv_perm[j]  (w_perm as an index)
 * tokens fp32 [batch*grid*grid, 768] (= patch_embedding); outputs: one_minus_prob fp32 [batch*grid*grid]
 * (= 1 - exp(-0.5 mean_c z^2), feed to vitad_bilinear_up with align_corners=0) and loss_terms fp32 [batch]
 * (= 0.5*sum z^2 - log|det J|; the reference's loss is their mean).
 * ------------------------------------------------------------------------------------------ */
typedef struct vitad_nf_step {
    const void* w0p;
    const float* b0p;
    const void* w2p;
    const float* b2p;
    const float* scale;
    const float* offset;
    const int* inv_perm;
    int ksize; /* 1 or 3 */
} vitad_nf_step;

typedef struct vitad_nf_weights {
    int channels, grid, hidden_pad, steps;
    float clamp;        /* affine_clamping (2.0) */
    float logdet_const; /* sum over steps of grid*grid*sum(log scale) */
    const vitad_nf_step* step; /* host array of `steps` entries */
} vitad_nf_weights;

size_t vitad_nf_workspace_bytes(const vitad_nf_weights* w, int batch);
int vitad_nf_forward(const vitad_nf_weights* w, const float* tokens, int batch, void* workspace,
                     size_t workspace_bytes, float* one_minus_prob, float* loss_terms, void* stream);

/* ------------------------------------------------------------------------------------------
 * Anomaly-map tails.
 * vitad_bilinear_up: in fp32 [n,g,g] -> out fp32 [n,S,S], PyTorch bilinear semantics for both
 *   align_corners settings; value = post(interp(pre(in))) with pre/post = optional (1 - v); image_max
 *   (may be null) receives max over each produced map (maps are >= 0 on this path).
 *     ValidatorMDN.py:137-162,171-172 -> align_corners=1, pre=0, post=1
 *     NormalizingFlow.py:134-143 + ValidatorNF.py:137-142 -> align_corners=0, pre=1, post=0, image_max
 * vitad_l2_map_score: map = mean_c (recon - x)^2, image_max = amax(map)
 *     CnnAutoEncoder.py:49,68-74 (MSELoss 'none') + ValidatorRecon.py:109-116
 * ------------------------------------------------------------------------------------------ */
int vitad_bilinear_up(const float* in, float* out, float* image_max, int n, int grid_in, int size_out,
                      int align_corners, int pre_one_minus, int post_one_minus, void* stream);
int vitad_l2_map_score(const float* recon, const float* x, float* map, float* image_max, int n, int channels, int hw,
                       void* stream);

/* ------------------------------------------------------------------------------------------
 * Small CNN decoder of the reconstruction models: DecoderVanillaCNN.forward (src/classes/CnnDecoder.py:16-117) as used by
 * AutoEncoderDeit(decoder="cnn") (TransformerAutoEncoder.py:152-194): two Linear+ReLU, four ConvTranspose2d(k3,s2,p1,op1)
 * + BatchNorm2d (inference statistics, folded) + ReLU, and ConvTranspose2d(48->3) + BatchNorm2d + Tanh.
 * Packed weights (vitad.autoencoders builds them from the reference's state_dict):
 *   lin1_w fp16 [hidden, latent], lin1_b fp32;  lin2_w fp16 [grid0*grid0*chan[0], hidden] with rows in (h, w, c) order,
 *   lin2_b fp32 in the same order;  conv_w[l] fp16 [4*chan[l+1], 4*chan[l]]: row (a, c, co) = output phase and channel,
 *   column (di, dj, ci) = input tap and channel, BN scale folded, dead taps and padding channels zero;  conv_b[l] fp32
 *   [4*chan[l+1]];  last_w fp32 [3][3][last_cin][3] (ky, kx, ci, co), last_b fp32 [3].
 * chan[] are channel pitches (multiples of 32: 768, 384, 192, 96, 64 for the reference's 768..48).
 * latent fp32 [B, latent] -> recon fp32 NCHW [B, 3, 32*grid0, 32*grid0].
 * ------------------------------------------------------------------------------------------ */
typedef struct vitad_cnn_decoder_weights {
    int latent, hidden, grid0, last_cin;
    int chan[5];
    const void* lin1_w;
    const float* lin1_b;
    const void* lin2_w;
    const float* lin2_b;
    const void* conv_w[4];
    const float* conv_b[4];
    const float* last_w;
    const float* last_b;
} vitad_cnn_decoder_weights;
size_t vitad_cnn_decoder_workspace_bytes(const vitad_cnn_decoder_weights* w, int batch);
int vitad_cnn_decoder_forward(const vitad_cnn_decoder_weights* w, const float* latent, int batch, void* workspace,
                              size_t workspace_bytes, float* recon, void* stream);

/* ------------------------------------------------------------------------------------------
 * Input path: the loader's transforms.Resize((S,S)) + ToTensor (src/data_loader/GeneralDataset.py:38-59) on the device.
 * torchvision resizes PIL images with Pillow's BILINEAR filter (antialiased triangle filter, 8-bit fixed point, horizontal
 * pass then vertical pass with a uint8 intermediate; Pillow Resample.c); these entry points reproduce it bit for bit.
 *   vitad_resize_ksize   taps per output pixel for one axis
 *   vitad_resize_plan    HOST memory: plan[0..S) first source pixel, plan[S..2S) tap count, plan[2S + i*ksize + t]
 *                        int32 coefficients (22 fractional bits); the caller copies the plan to the device once per geometry
 *   vitad_resize_bilinear_u8  in uint8 HWC [B,H,W,3] -> tmp uint8 [B,H,S,3] -> out uint8 planar [B,3,S,S] (what
 *                        vitad_deit_forward_u8 reads; ToTensor's /255 is folded into its patch gather)
 * ------------------------------------------------------------------------------------------ */
int vitad_resize_ksize(int in_size, int out_size);
int vitad_resize_plan(int in_size, int out_size, int32_t* plan);
int vitad_resize_bilinear_u8(const uint8_t* in, int batch, int height, int width, int out_size, const int32_t* plan_h,
                             const int32_t* plan_v, uint8_t* tmp, uint8_t* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Reverse-ResNet decoder of the reconstruction models: DecoderResNetVariableEmbeddingSize.forward
 * (src/classes/CnnDecoder.py:158-196) over ReverseResNet (src/classes/resnet/ReverseResNet.py:106-242), the default
 * decoder of AutoEncoderDeit (TransformerAutoEncoder.py:155-181, get_model("ae_deit")):
 *   latent [B,768] -> Linear+ReLU (1536) -> Linear+ReLU (2048) -> nearest upsample of the 1x1 feature to grid0 x grid0
 *   -> 16 Bottleneck blocks (layer4: 3 at 7x7, layer3: 4 at 14x14, layer2: 6 at 28x28, layer1: 3 at 56x56; the last
 *   block of layers 4..2 doubles the grid) -> nearest upsample to 112 -> ConvTranspose2d(64->3, k7, s2, p3, op1)
 *   -> BatchNorm2d -> Tanh: fp32 NCHW [B,3,224,224].
 * Every convolution is a tcgen05 GEMM over NHWC fp16 activations with BatchNorm (inference statistics) folded:
 *   w3 fp16 [width, cin]                  conv3 (1x1 transposed conv = per-pixel Linear), b3 fp32 [width]
 *   w2 fp16 [width, 9*width]  (stride 1)  conv2 as a 3x3 convolution over im2col rows, column (ty, tx, ci) holds the
 *                                         transposed-conv weight [ci, co, 2-ty, 2-tx]; b2 fp32 [width]
 *      fp16 [4*width, 4*width] (stride 2) conv2 as the four-phase GEMM of VITAD_EPI_CONVT_RELU_F16; b2 fp32 [4*width]
 *   w1 fp16 [cout, width]                 conv1; b1 fp32 [cout] = bn1 shift (+ the identity path's BatchNorm shift)
 *   wup fp16 [cout, cin] or null          1x1 transposed conv of the identity path (`upsample`, :186-196); bup fp32
 *                                         [cout] = its BatchNorm shift when stride 1, zeros when stride 2 (the shift
 *                                         then belongs in b1: it reaches every output pixel, the convolution only
 *                                         the even ones)
 *   last_w fp16 [64, 9*64]                nearest-upsample + 7x7 stride-2 transposed conv + BatchNorm scale collapsed to
 *                                         a 3x3 convolution on the 56x56 grid that produces a 4x4 pixel block x 3
 *                                         channels per input pixel (row c*16+py*4+px, 48 live); last_b fp32 [64]
 * Channel counts must be multiples of 32.  Workspace: vitad_resnet_decoder_workspace_bytes.
 * ------------------------------------------------------------------------------------------ */
#define VITAD_RESNET_MAX_BLOCKS 24
typedef struct vitad_resnet_block {
    int cin, width, cout, stride; /* stride 1 or 2 */
    const void* w3;
    const float* b3;
    const void* w2;
    const float* b2;
    const void* w1;
    const float* b1;
    const void* wup;
    const float* bup;
} vitad_resnet_block;
typedef struct vitad_resnet_decoder_weights {
    int latent, hidden, feat, grid0, n_blocks, last_c;
    const void* fc1_w;
    const float* fc1_b;
    const void* fc2_w;
    const float* fc2_b;
    vitad_resnet_block blocks[VITAD_RESNET_MAX_BLOCKS];
    const void* last_w;
    const float* last_b;
    /* 1: split-fp16 arithmetic (vitad_linear_args.split_c): every packed weight matrix holds [w_hi | w_hi | w_lo] per tap
     * (3x the columns listed above), activations travel as [hi | lo] pairs, all channel counts % 64 == 0.  3x the tensor
     * work for fp32-grade products: the 53 chained layers otherwise leave ~2e-3 of the maximum on the L2 anomaly map,
     * outside the 1e-3 the reference's fp32 decoder meets.  0: plain fp16 operands. */
    int split;
} vitad_resnet_decoder_weights;
size_t vitad_resnet_decoder_workspace_bytes(const vitad_resnet_decoder_weights* w, int batch);
int vitad_resnet_decoder_forward(const vitad_resnet_decoder_weights* w, const float* latent, int batch, void* workspace,
                                 size_t workspace_bytes, float* recon, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VITAD_H_ */
