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

/* ------------------------------------------------------------------------------------------
 * Dense projection  D = A · Wᵀ (+ fused epilogue), bf16 operands, fp32 accumulate (tcgen05/TMEM,
 * TMA-fed).  Replaces every nn.Linear / conv-as-GEMM on the encoder path:
 *   timm Attention.qkv / Attention.proj / Mlp.fc1 / Mlp.fc2 / PatchEmbed.proj, called through
 *   src/classes/transformer/TransformerEncoder.py:150-165 (EncoderDeit.forward).
 * A: [M,K] bf16 row-major (pitch lda elements), W: [N,K] bf16 row-major (nn.Linear layout).
 * K % 16 == 0, N % 32 == 0 (N % 8 for VITAD_EPI_F32), pitches % 8 == 0.
 * ------------------------------------------------------------------------------------------ */
typedef enum vitad_epilogue {
    VITAD_EPI_BIAS_BF16 = 0,      /* out_bf16 = acc + bias                                  */
    VITAD_EPI_BIAS_GELU_BF16 = 1, /* out_bf16 = gelu_erf(acc + bias)   (timm Mlp.act)       */
    VITAD_EPI_RESIDUAL_F32 = 2,   /* out_f32  = resid_f32 + acc + bias (timm Block residual) */
    VITAD_EPI_QKV = 3,            /* head-major q (pre-scaled), k, transposed v             */
    VITAD_EPI_PATCH_EMBED = 4,    /* + bias + pos_embed, written behind the prefix tokens   */
    VITAD_EPI_F32 = 5             /* out_f32 = acc (+ bias if non-null)                     */
} vitad_epilogue;

typedef struct vitad_linear_args {
    const void* a;      /* bf16 [M,K] */
    const void* w;      /* bf16 [N,K] */
    const float* bias;  /* fp32 [N]   */
    int m, n, k;
    int lda, ldw;       /* pitches in elements */
    int epilogue;       /* vitad_epilogue */
    int block_n;        /* 0 = library default, else 128 or 256 */
    void* out;          /* bf16 or fp32 [M,ldo], see epilogue */
    int ldo;
    const float* resid; /* RESIDUAL_F32: fp32 [M,ldo] (may alias out) */
    /* QKV: q,k bf16 [B,H,T,64]; vt bf16 [B,H,64,Tpad]; M = B*T, N = 3*H*64 */
    void* q;
    void* kmat;
    void* vt;
    int tokens, tokens_pad, heads;
    float q_scale;
    /* PATCH_EMBED: out fp32 [B,prefix+P,N]; pos fp32 [prefix+P,N]; M = B*P */
    const float* pos;
    int patches, prefix;
} vitad_linear_args;

int vitad_linear_bf16(const vitad_linear_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VITAD_H_ */
