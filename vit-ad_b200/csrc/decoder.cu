// vitad_cnn_decoder_forward: the small CNN decoder of the reconstruction models (DecoderVanillaCNN,
// src/classes/CnnDecoder.py:16-117; used by AutoEncoderDeit(decoder="cnn"), TransformerAutoEncoder.py:152-194) as
// one C-ABI call:
//   latent [B,768] -> Linear(768,1536)+ReLU -> Linear(1536,768*7*7)+ReLU -> unflatten [768,7,7]
//   -> 4 x (ConvTranspose2d(k3,s2,p1,op1) + BatchNorm2d(eval) + ReLU): 768 -> 384 -> 192 -> 96 -> 48 channels, 7 -> 112 px
//   -> ConvTranspose2d(48 -> 3) + BatchNorm2d + Tanh: 224 x 224 reconstruction.
// Design:
//   * activations are NHWC fp16 between layers (the second Linear's rows are permuted at pack time so its output
//     already is [B,7,7,768]); BatchNorm (inference statistics) is folded into the packed weights and biases;
//   * a stride-2 transposed 3x3 convolution is four ordinary convolutions, one per output phase (y mod 2, x mod 2),
//     over the 2x2 input neighbourhood with 1/2/2/4 live taps.  All four run as ONE tcgen05 GEMM per layer:
//     A = im2col2x2 [B*H*W, 4*C_in], W = [4*C_out, 4*C_in] (dead taps are zero: 16/9 of the minimal FLOPs, still < 60
//     GFLOP per batch of 32), and the staged epilogue (SEpiConvT, gemm_staged.cuh) adds the folded bias, applies ReLU
//     and scatters each phase to its output pixel;
//   * the last layer has 3 output channels: a CUDA-core kernel (1296 FMAs per 2x2 output block) with the folded
//     BatchNorm and tanh, writing the reconstruction as fp32 NCHW (what the validator returns and the L2-map kernel reads).
//
// vitad_resnet_decoder_forward: the reverse-ResNet decoder (DecoderResNetVariableEmbeddingSize, CnnDecoder.py:158-196, over
// ReverseResNet.py:86-103,169-235; AutoEncoderDeit's default) as one C-ABI call: fc1, fc2, the 1x1 feature replicated to
// 7x7, 16 Bottleneck blocks, image head.  Every convolution is a tcgen05 GEMM on NHWC fp16 rows (vitad_linear_f16):
//   conv3 (1x1) bias+ReLU | conv2 (3x3, stride 1) as an implicit convolution over a zero-bordered activation for grids
//   >= 14 (im2col at 7x7), or the four-phase GEMM above for the stride-2 block of a layer | conv1 + identity + ReLU in one
//   epilogue, the stride-2 identity path computed on the input grid and added on the even output pixels | the nearest
//   upsample + 7x7 stride-2 transposed convolution + BatchNorm + tanh head collapsed to one implicit 3x3 convolution that
//   emits 4x4 pixel blocks.  See include/vitad.h for the packed layouts and DESIGN.md 4.6 for the derivations.
#include <algorithm>
#include <atomic>

#include "host_util.cuh"
#include "ptx.cuh"

extern "C" int vitad_linear_f16(const vitad_linear_args* args, void* stream);

namespace vitad {
extern std::atomic<uint64_t> g_launches;

__global__ void __launch_bounds__(256) cast_f16_kernel(const float* __restrict__ x, __half* __restrict__ y, size_t n4) {
    griddep_launch_dependents();
    griddep_wait();
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    uint2 u;
    u.x = pack_h2(v.x, v.y);
    u.y = pack_h2(v.z, v.w);
    reinterpret_cast<uint2*>(y)[i] = u;
}

// y[r] = [hi (cols) | lo (cols)] of x[r] (split-fp16 operand form, gemm_staged.cuh: split_h4); one thread = 4 columns
__global__ void __launch_bounds__(256) cast_split_f16_kernel(const float* __restrict__ x, __half* __restrict__ y, int cols,
                                                              size_t n4) {
    griddep_launch_dependents();
    griddep_wait();
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    const int per_row = cols >> 2;
    const size_t r = i / per_row;
    const int c = static_cast<int>(i - r * per_row) << 2;
    uint2 hi, lo;
    split_h4(v.x, v.y, v.z, v.w, hi, lo);
    __half* dst = y + r * (2 * static_cast<size_t>(cols)) + c;
    *reinterpret_cast<uint2*>(dst) = hi;
    *reinterpret_cast<uint2*>(dst + cols) = lo;
}

// A[(b,i,j)][(di,dj,c)] = in[b, i+di, j+dj, c] (NHWC fp16, C channels per pixel), zero outside the Wg x Wg grid.
// One thread moves 8 channels (16 bytes).
__global__ void __launch_bounds__(256) im2col2x2_kernel(const __half* __restrict__ in, __half* __restrict__ a, int Wg,
                                                        int C, size_t total8) {
    griddep_launch_dependents();
    griddep_wait();
    const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total8) return;
    const int c8 = C >> 3;
    const int c = static_cast<int>(idx % c8) << 3;
    size_t r = idx / c8;
    const int tap = static_cast<int>(r & 3);
    r >>= 2;  // pixel row (b, i, j)
    const int di = tap >> 1, dj = tap & 1;
    const int j = static_cast<int>(r % Wg);
    const int i = static_cast<int>((r / Wg) % Wg);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (i + di < Wg && j + dj < Wg)
        v = *reinterpret_cast<const uint4*>(in + (r + static_cast<size_t>(di) * Wg + dj) * C + c);
    *reinterpret_cast<uint4*>(a + (r * 4 + tap) * C + c) = v;
}

// Last layer: ConvTranspose2d(C_in -> 3, k3, s2, p1, op1) + folded BatchNorm + tanh.  in NHWC fp16 [B,Hi,Hi,Cp] (Cp = padded
// channel pitch, C_in live channels), w fp32 [3 ky][3 kx][C_in][3] with the BN scale folded, bias fp32 [3];
// out fp32 NCHW [B,3,2Hi,2Hi].  One thread = one input pixel = one 2x2 output block.
template <int CIN>
__global__ void __launch_bounds__(256) convt_last_kernel(const __half* __restrict__ in, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ out, int B,
                                                         int Hi, int Cp) {
    griddep_launch_dependents();
    __shared__ __align__(16) float ws[9 * CIN * 3];
    for (int i = threadIdx.x; i < 9 * CIN * 3; i += 256) ws[i] = w[i];
    griddep_wait();
    __syncthreads();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * Hi * Hi) return;
    const int j = idx % Hi;
    const int i = (idx / Hi) % Hi;
    const int b = idx / (Hi * Hi);
    float acc[2][2][3];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int co = 0; co < 3; ++co) acc[a][c][co] = bias[co];
#pragma unroll
    for (int tap = 0; tap < 4; ++tap) {
        const int di = tap >> 1, dj = tap & 1;
        if (i + di >= Hi || j + dj >= Hi) continue;
        const __half* px = in + (static_cast<size_t>(b * Hi + i + di) * Hi + j + dj) * Cp;
#pragma unroll
        for (int c8 = 0; c8 < CIN; c8 += 8) {
            const uint4 raw = *reinterpret_cast<const uint4*>(px + c8);
            const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
            float xv[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = __half22float2(h2[e]);
                xv[2 * e] = f.x;
                xv[2 * e + 1] = f.y;
            }
            // output phase (a, c) sees this tap through kernel element ky = a + 1 - 2 di, kx = c + 1 - 2 dj
#pragma unroll
            for (int a = di; a < 2; ++a) {
#pragma unroll
                for (int c = dj; c < 2; ++c) {
                    const int ky = a + 1 - 2 * di, kx = c + 1 - 2 * dj;
                    // 8 channels x 3 outputs = 24 consecutive weights: six broadcast LDS.128 instead of 24 scalar loads
                    const float4* wk4 = reinterpret_cast<const float4*>(ws + ((ky * 3 + kx) * CIN + c8) * 3);
                    float wk[24];
#pragma unroll
                    for (int q = 0; q < 6; ++q) {
                        const float4 t4 = wk4[q];
                        wk[4 * q + 0] = t4.x, wk[4 * q + 1] = t4.y, wk[4 * q + 2] = t4.z, wk[4 * q + 3] = t4.w;
                    }
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        acc[a][c][0] = fmaf(xv[e], wk[e * 3 + 0], acc[a][c][0]);
                        acc[a][c][1] = fmaf(xv[e], wk[e * 3 + 1], acc[a][c][1]);
                        acc[a][c][2] = fmaf(xv[e], wk[e * 3 + 2], acc[a][c][2]);
                    }
                }
            }
        }
    }
    const int Ho = 2 * Hi;
#pragma unroll
    for (int co = 0; co < 3; ++co)
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            float2 v;
            v.x = tanhf(acc[a][0][co]);
            v.y = tanhf(acc[a][1][co]);
            *reinterpret_cast<float2*>(out + ((static_cast<size_t>(b) * 3 + co) * Ho + 2 * i + a) * Ho + 2 * j) = v;
        }
}

// 3x3 im2col on a g x g grid per image, zero padding: in fp16 [M][cw] -> out fp16 [M][9*cw], column (tap, c),
// tap = ty*3+tx reads pixel (y+ty-1, x+tx-1).  One thread moves 8 channels (16 bytes).
__global__ void __launch_bounds__(256) dec_im2col3x3_kernel(const __half* __restrict__ in, __half* __restrict__ out, int cw,
                                                            int g, size_t total) {
    griddep_launch_dependents();
    griddep_wait();
    const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int v_per_tap = cw >> 3;
    const int v = static_cast<int>(idx % v_per_tap);
    const size_t r = idx / v_per_tap;
    const int tap = static_cast<int>(r % 9);
    const size_t t = r / 9;
    const int p = static_cast<int>(t % (static_cast<size_t>(g) * g));
    const int py = p / g, px = p - py * g;
    const int y = py + tap / 3 - 1, x = px + tap % 3 - 1;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (y >= 0 && y < g && x >= 0 && x < g)
        val = *reinterpret_cast<const uint4*>(in + (t - p + static_cast<size_t>(y) * g + x) * cw + v * 8);
    *reinterpret_cast<uint4*>(out + (t * 9 + tap) * cw + v * 8) = val;
}

// Zero the one-pixel border of a bordered NHWC activation [B, g+2, g+2, C] (the padding the implicit 3x3 convolution reads).
// Only the 4g+4 border pixels are touched, and the kernel is a link of the programmatic-launch chain (a cudaMemsetAsync of
// the whole buffer is not).  One thread clears 8 channels.
__global__ void __launch_bounds__(256) dec_zero_border_kernel(__half* __restrict__ buf, int g, int c8, size_t total) {
    griddep_launch_dependents();
    griddep_wait();
    const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int v = static_cast<int>(idx % c8);
    size_t r = idx / c8;
    const int P = g + 2, nb = 4 * g + 4;
    const int k = static_cast<int>(r % nb);
    const size_t b = r / nb;
    int y, x;
    if (k < P) y = 0, x = k;                          // top row
    else if (k < 2 * P) y = P - 1, x = k - P;         // bottom row
    else if (k < 2 * P + g) y = k - 2 * P + 1, x = 0; // left column
    else y = k - 2 * P - g + 1, x = P - 1;            // right column
    reinterpret_cast<uint4*>(buf)[((b * P + y) * P + x) * c8 + v] = make_uint4(0u, 0u, 0u, 0u);
}

// nn.Upsample(size=g, mode="nearest") of a 1x1 feature map (ReverseResNet.py:135,229): out[(b, p)][c] = in[b][c].
__global__ void __launch_bounds__(256) dec_replicate_kernel(const __half* __restrict__ in, __half* __restrict__ out, int pix,
                                                            int c8, size_t total) {
    griddep_launch_dependents();
    griddep_wait();
    const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int v = static_cast<int>(idx % c8);
    const size_t b = idx / c8 / pix;
    reinterpret_cast<uint4*>(out)[idx] = reinterpret_cast<const uint4*>(in)[b * c8 + v];
}

namespace {
struct Carve {
    uint8_t* p;
    size_t used = 0;
    void* take(size_t bytes) {
        void* r = p ? p + used : nullptr;
        used += (bytes + 255) & ~static_cast<size_t>(255);
        return r;
    }
};
struct DecWs {
    void *lat, *h1, *act[5], *col[4];
    size_t total;
};
DecWs carve(const vitad_cnn_decoder_weights& w, int batch, void* base) {
    Carve c{static_cast<uint8_t*>(base)};
    DecWs s;
    s.lat = c.take(static_cast<size_t>(batch) * w.latent * 2);
    s.h1 = c.take(static_cast<size_t>(batch) * w.hidden * 2);
    int g = w.grid0;
    s.act[0] = c.take(static_cast<size_t>(batch) * g * g * w.chan[0] * 2);
    for (int l = 0; l < 4; ++l) {
        s.col[l] = c.take(static_cast<size_t>(batch) * g * g * 4 * w.chan[l] * 2);
        g *= 2;
        s.act[l + 1] = c.take(static_cast<size_t>(batch) * g * g * w.chan[l + 1] * 2);
    }
    s.total = c.used;
    return s;
}
// conv2 of a stride-1 block runs as an implicit 3x3 convolution (no im2col pass) when the (g+2)^2/g^2 extra MMA work on
// border rows is cheaper than writing and re-reading the 9x larger im2col matrix
inline bool implicit_conv(const vitad_resnet_block& b, int g) { return b.stride == 1 && g >= 14 && b.width % 64 == 0; }

struct ResWs {
    void *lat, *h, *f, *x[2], *h1, *col, *h2, *up;
    size_t total;
};
ResWs carve_resnet(const vitad_resnet_decoder_weights& w, int batch, void* base) {
    const size_t S = w.split ? 2 : 1;  // split-fp16 activations: [hi | lo] per pixel
    size_t x_el = static_cast<size_t>(batch) * w.grid0 * w.grid0 * w.feat, h1_el = 0, col_el = 64, h2_el = 0, up_el = 0;
    int g = w.grid0;
    for (int i = 0; i < w.n_blocks; ++i) {
        const vitad_resnet_block& b = w.blocks[i];
        const size_t M = static_cast<size_t>(batch) * g * g;
        const size_t Mp = static_cast<size_t>(batch) * (g + 2) * (g + 2);  // zero-bordered layout
        const size_t Mo = b.stride == 2 ? 4 * M : M;
        x_el = std::max(x_el, std::max(M * b.cin, Mo * b.cout));
        if (implicit_conv(b, g)) {
            h1_el = std::max(h1_el, Mp * b.width);
        } else {
            h1_el = std::max(h1_el, M * b.width);
            col_el = std::max(col_el, M * b.width * (b.stride == 2 ? 4 : 9));
        }
        h2_el = std::max(h2_el, Mo * b.width);
        if (b.wup) up_el = std::max(up_el, M * b.cout);
        if (b.stride == 2) g *= 2;
    }
    x_el = std::max(x_el, static_cast<size_t>(batch) * (g + 2) * (g + 2) * w.last_c);  // the image head reads a bordered map
    Carve c{static_cast<uint8_t*>(base)};
    ResWs s;
    s.lat = c.take(static_cast<size_t>(batch) * w.latent * 2 * S);
    s.h = c.take(static_cast<size_t>(batch) * w.hidden * 2 * S);
    s.f = c.take(static_cast<size_t>(batch) * w.feat * 2 * S);
    s.x[0] = c.take(x_el * 2 * S);
    s.x[1] = c.take(x_el * 2 * S);
    s.h1 = c.take(h1_el * 2 * S);
    s.col = c.take(col_el * 2 * S);
    s.h2 = c.take(h2_el * 2 * S);
    s.up = c.take(up_el * 2 * S);
    s.total = c.used;
    return s;
}
int check_resnet(const vitad_resnet_decoder_weights& w) {
    VITAD_REQUIRE(w.latent % 16 == 0 && w.hidden % 32 == 0 && w.feat % 32 == 0 && w.grid0 > 0 && w.n_blocks > 0 &&
                      w.n_blocks <= VITAD_RESNET_MAX_BLOCKS && w.last_c == 64 && w.fc1_w && w.fc1_b && w.fc2_w && w.fc2_b &&
                      w.last_w && w.last_b,
                  VITAD_ERR_SHAPE, "unsupported reverse-ResNet decoder geometry");
    int cin = w.feat;
    for (int i = 0; i < w.n_blocks; ++i) {
        const vitad_resnet_block& b = w.blocks[i];
        VITAD_REQUIRE(b.cin == cin && b.width % 32 == 0 && b.cout % 32 == 0 && b.width > 0 && b.cout > 0 &&
                          (b.stride == 1 || b.stride == 2) && b.w3 && b.b3 && b.w2 && b.b2 && b.w1 && b.b1,
                      VITAD_ERR_SHAPE, "reverse-ResNet block %d: channels must chain and be multiples of 32", i);
        VITAD_REQUIRE((b.wup && b.bup) || (!b.wup && b.stride == 1 && b.cin == b.cout), VITAD_ERR_SHAPE,
                      "reverse-ResNet block %d: an identity path without a convolution needs stride 1 and cin == cout", i);
        cin = b.cout;
    }
    VITAD_REQUIRE(cin == w.last_c, VITAD_ERR_SHAPE, "the last block must end in %d channels", w.last_c);
    if (w.split) {
        VITAD_REQUIRE(w.split == 1 && w.latent % 64 == 0 && w.hidden % 64 == 0 && w.feat % 64 == 0, VITAD_ERR_SHAPE,
                      "split-fp16 decoder: every channel count must be a multiple of 64");
        for (int i = 0; i < w.n_blocks; ++i)
            VITAD_REQUIRE(w.blocks[i].cin % 64 == 0 && w.blocks[i].width % 64 == 0 && w.blocks[i].cout % 64 == 0,
                          VITAD_ERR_SHAPE, "split-fp16 decoder: block %d channels must be multiples of 64", i);
    }
    return VITAD_OK;
}
}  // namespace
}  // namespace vitad

using namespace vitad;

extern "C" size_t vitad_cnn_decoder_workspace_bytes(const vitad_cnn_decoder_weights* w, int batch) {
    if (!w || batch <= 0) return 0;
    return carve(*w, batch, nullptr).total;
}

extern "C" int vitad_cnn_decoder_forward(const vitad_cnn_decoder_weights* wp, const float* latent, int batch,
                                         void* workspace, size_t workspace_bytes, float* recon, void* stream) {
    VITAD_NVTX("vitad_cnn_decoder_forward");
    int rc = check_device_arch();
    if (rc) return rc;
    VITAD_REQUIRE(wp && latent && workspace && recon, VITAD_ERR_ARG, "null pointer");
    const vitad_cnn_decoder_weights& w = *wp;
    VITAD_REQUIRE(batch > 0 && w.latent % 16 == 0 && w.hidden % 32 == 0 && w.grid0 > 0 && w.last_cin == 48, VITAD_ERR_SHAPE,
                  "unsupported decoder geometry");
    for (int l = 0; l < 5; ++l)
        VITAD_REQUIRE(w.chan[l] % 32 == 0 && w.chan[l] > 0, VITAD_ERR_SHAPE, "channel pitches must be multiples of 32");
    VITAD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0 && aligned16(latent) && aligned16(recon), VITAD_ERR_ALIGN,
                  "decoder buffers alignment");
    DecWs ws = carve(w, batch, workspace);
    VITAD_REQUIRE(workspace_bytes >= ws.total, VITAD_ERR_WORKSPACE, "workspace %zu < %zu bytes", workspace_bytes, ws.total);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    {
        ProfScope prof("dec_cast", s);
        const size_t n4 = static_cast<size_t>(batch) * w.latent / 4;
        VITAD_CUDA_OK(launch_pdl(cast_f16_kernel, dim3(static_cast<unsigned>((n4 + 255) / 256)), dim3(256), 0, s, latent,
                                 static_cast<__half*>(ws.lat), n4));
        g_launches.fetch_add(1);
    }
    vitad_linear_args a;
    memset(&a, 0, sizeof(a));
    a.a = ws.lat, a.w = w.lin1_w, a.bias = w.lin1_b, a.m = batch, a.n = w.hidden, a.k = w.latent, a.lda = w.latent,
    a.ldw = w.latent, a.epilogue = VITAD_EPI_BIAS_RELU_F16, a.out = ws.h1, a.ldo = w.hidden;
    if ((rc = vitad_linear_f16(&a, s))) return rc;
    const int n2 = w.grid0 * w.grid0 * w.chan[0];
    memset(&a, 0, sizeof(a));
    a.a = ws.h1, a.w = w.lin2_w, a.bias = w.lin2_b, a.m = batch, a.n = n2, a.k = w.hidden, a.lda = w.hidden, a.ldw = w.hidden;
    a.epilogue = VITAD_EPI_BIAS_RELU_F16, a.out = ws.act[0], a.ldo = n2;
    if ((rc = vitad_linear_f16(&a, s))) return rc;
    int g = w.grid0;
    for (int l = 0; l < 4; ++l) {
        const int cin = w.chan[l], cout = w.chan[l + 1];
        const int M = batch * g * g;
        {
            ProfScope prof("dec_im2col", s);
            const size_t total8 = static_cast<size_t>(M) * 4 * (cin / 8);
            VITAD_CUDA_OK(launch_pdl(im2col2x2_kernel, dim3(static_cast<unsigned>((total8 + 255) / 256)), dim3(256), 0, s,
                                     static_cast<const __half*>(ws.act[l]), static_cast<__half*>(ws.col[l]), g, cin, total8));
            g_launches.fetch_add(1);
        }
        memset(&a, 0, sizeof(a));
        a.a = ws.col[l], a.w = w.conv_w[l], a.bias = w.conv_b[l], a.m = M, a.n = 4 * cout, a.k = 4 * cin, a.lda = 4 * cin,
        a.ldw = 4 * cin, a.epilogue = VITAD_EPI_CONVT_RELU_F16, a.out = ws.act[l + 1], a.ldo = 2 * cout, a.convt_w = g;
        if ((rc = vitad_linear_f16(&a, s))) return rc;
        g *= 2;
    }
    {
        ProfScope prof("dec_last", s);
        const int pixels = batch * g * g;
        VITAD_CUDA_OK(launch_pdl(convt_last_kernel<48>, dim3((pixels + 255) / 256), dim3(256), 0, s,
                                 static_cast<const __half*>(ws.act[4]), w.last_w, w.last_b, recon, batch, g, w.chan[4]));
        g_launches.fetch_add(1);
    }
    return VITAD_OK;
}

extern "C" size_t vitad_resnet_decoder_workspace_bytes(const vitad_resnet_decoder_weights* w, int batch) {
    if (!w || batch <= 0 || check_resnet(*w)) return 0;
    return carve_resnet(*w, batch, nullptr).total;
}

extern "C" int vitad_resnet_decoder_forward(const vitad_resnet_decoder_weights* wp, const float* latent, int batch,
                                            void* workspace, size_t workspace_bytes, float* recon, void* stream) {
    VITAD_NVTX("vitad_resnet_decoder_forward");
    int rc = check_device_arch();
    if (rc) return rc;
    VITAD_REQUIRE(wp && latent && workspace && recon && batch > 0, VITAD_ERR_ARG, "null pointer or empty batch");
    const vitad_resnet_decoder_weights& w = *wp;
    if ((rc = check_resnet(w))) return rc;
    VITAD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0 && aligned16(latent) && aligned16(recon), VITAD_ERR_ALIGN,
                  "decoder buffers alignment");
    ResWs ws = carve_resnet(w, batch, workspace);
    VITAD_REQUIRE(workspace_bytes >= ws.total, VITAD_ERR_WORKSPACE, "workspace %zu < %zu bytes", workspace_bytes, ws.total);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    auto blocks_for = [](size_t n) { return dim3(static_cast<unsigned>((n + 255) / 256)); };
    // Split-fp16 mode (w.split): activations are [hi | lo] pairs (S = 2 halves per pixel), packed weights hold
    // [w_hi | w_hi | w_lo] per tap (KM = 3), every GEMM runs the three exact partial products (include/vitad.h: split_c).
    const int S = w.split ? 2 : 1, KM = w.split ? 3 : 1;
    {
        ProfScope prof("dec_cast", s);
        const size_t n4 = static_cast<size_t>(batch) * w.latent / 4;
        if (w.split)
            VITAD_CUDA_OK(launch_pdl(cast_split_f16_kernel, blocks_for(n4), dim3(256), 0, s, latent, static_cast<__half*>(ws.lat),
                                     w.latent, n4));
        else
            VITAD_CUDA_OK(launch_pdl(cast_f16_kernel, blocks_for(n4), dim3(256), 0, s, latent, static_cast<__half*>(ws.lat), n4));
        g_launches.fetch_add(1);
    }
    vitad_linear_args a;
    // A: `taps` taps of C channels each (S halves per tap); W: [N, taps * KM * C]; out: N channels (S halves)
    auto gemm = [&](const void* A, int M, int C, int taps, const void* W, const float* bias, int N, int epi, void* out, int ldo) {
        memset(&a, 0, sizeof(a));
        a.a = A, a.w = W, a.bias = bias, a.m = M, a.n = N, a.k = taps * KM * C, a.lda = taps * S * C, a.ldw = taps * KM * C;
        a.epilogue = epi, a.out = out, a.ldo = ldo;
        a.a_taps = taps > 1 ? taps : 0;
        a.split_c = w.split ? C : 0;
        a.split_out = w.split;
    };
    // fc1, fc2 (CnnDecoder.py:171-179, 185-186)
    gemm(ws.lat, batch, w.latent, 1, w.fc1_w, w.fc1_b, w.hidden, VITAD_EPI_BIAS_RELU_F16, ws.h, S * w.hidden);
    if ((rc = vitad_linear_f16(&a, s))) return rc;
    gemm(ws.h, batch, w.hidden, 1, w.fc2_w, w.fc2_b, w.feat, VITAD_EPI_BIAS_RELU_F16, ws.f, S * w.feat);
    if ((rc = vitad_linear_f16(&a, s))) return rc;
    int g = w.grid0;
    {
        ProfScope prof("dec_replicate", s);
        const size_t total = static_cast<size_t>(batch) * g * g * (S * w.feat / 8);
        VITAD_CUDA_OK(launch_pdl(dec_replicate_kernel, blocks_for(total), dim3(256), 0, s, static_cast<const __half*>(ws.f),
                                 static_cast<__half*>(ws.x[0]), g * g, S * w.feat / 8, total));
        g_launches.fetch_add(1);
    }
    int cur = 0;
    int bordered_g = 0, bordered_w = 0;  // geometry whose zero border ws.h1 currently holds
    for (int i = 0; i < w.n_blocks; ++i) {
        const vitad_resnet_block& b = w.blocks[i];
        const int M = batch * g * g;
        const int Mp = batch * (g + 2) * (g + 2);
        const bool implicit = implicit_conv(b, g);
        const bool last = i == w.n_blocks - 1;
        const void* x = ws.x[cur];
        void* y = ws.x[cur ^ 1];
        // conv3 + bn3 + relu (ReverseResNet.py:89-91)
        if (implicit && (bordered_g != g || bordered_w != b.width)) {
            const size_t total = static_cast<size_t>(batch) * (4 * g + 4) * (S * b.width / 8);
            VITAD_CUDA_OK(launch_pdl(dec_zero_border_kernel, blocks_for(total), dim3(256), 0, s, static_cast<__half*>(ws.h1), g,
                                     S * b.width / 8, total));
            g_launches.fetch_add(1);
            bordered_g = g, bordered_w = b.width;
        }
        if (!implicit) bordered_g = 0;
        gemm(x, M, b.cin, 1, b.w3, b.b3, b.width, VITAD_EPI_BIAS_RELU_F16, ws.h1, S * b.width);
        a.out_pad_grid = implicit ? g : 0;
        if ((rc = vitad_linear_f16(&a, s))) return rc;
        // conv2 + bn2 + relu (:92-94)
        int Mo = M;
        if (implicit) {
            gemm(ws.h1, Mp, b.width, 9, b.w2, b.b2, b.width, VITAD_EPI_BIAS_RELU_F16, ws.h2, S * b.width);
            a.lda = S * b.width, a.conv_grid = g, a.a_taps = 0;
            if ((rc = vitad_linear_f16(&a, s))) return rc;
        } else if (b.stride == 1) {
            {
                ProfScope prof("dec_im2col3", s);
                const size_t total = static_cast<size_t>(M) * 9 * (S * b.width / 8);
                VITAD_CUDA_OK(launch_pdl(dec_im2col3x3_kernel, blocks_for(total), dim3(256), 0, s,
                                         static_cast<const __half*>(ws.h1), static_cast<__half*>(ws.col), S * b.width, g, total));
                g_launches.fetch_add(1);
            }
            gemm(ws.col, M, b.width, 9, b.w2, b.b2, b.width, VITAD_EPI_BIAS_RELU_F16, ws.h2, S * b.width);
            if ((rc = vitad_linear_f16(&a, s))) return rc;
        } else {
            {
                ProfScope prof("dec_im2col", s);
                const size_t total8 = static_cast<size_t>(M) * 4 * (S * b.width / 8);
                VITAD_CUDA_OK(launch_pdl(im2col2x2_kernel, blocks_for(total8), dim3(256), 0, s,
                                         static_cast<const __half*>(ws.h1), static_cast<__half*>(ws.col), g, S * b.width, total8));
                g_launches.fetch_add(1);
            }
            gemm(ws.col, M, b.width, 4, b.w2, b.b2, 4 * b.width, VITAD_EPI_CONVT_RELU_F16, ws.h2, S * 2 * b.width);
            a.convt_w = g;
            if ((rc = vitad_linear_f16(&a, s))) return rc;
            Mo = 4 * M;
        }
        // identity path (:98-99, 186-196): a 1x1 transposed convolution on the input grid, or the block input itself
        const void* resid = x;
        if (b.wup) {
            gemm(x, M, b.cin, 1, b.wup, b.bup, b.cout, VITAD_EPI_BIAS_F16, ws.up, S * b.cout);
            if ((rc = vitad_linear_f16(&a, s))) return rc;
            resid = ws.up;
        }
        // conv1 + bn1 + identity + relu (:95-101); the last block writes the zero-bordered layout the image head reads
        const int go = b.stride == 2 ? 2 * g : g;
        if (last) {
            const size_t total = static_cast<size_t>(batch) * (4 * go + 4) * (S * b.cout / 8);
            VITAD_CUDA_OK(launch_pdl(dec_zero_border_kernel, blocks_for(total), dim3(256), 0, s, static_cast<__half*>(y), go,
                                     S * b.cout / 8, total));
            g_launches.fetch_add(1);
        }
        gemm(ws.h2, Mo, b.width, 1, b.w1, b.b1, b.cout, VITAD_EPI_RES16_RELU_F16, y, S * b.cout);
        a.resid16 = resid, a.ldr = S * b.cout, a.res_grid = b.stride == 2 ? g : 0, a.out_pad_grid = last ? go : 0;
        if ((rc = vitad_linear_f16(&a, s))) return rc;
        cur ^= 1;
        g = go;
    }
    // image head (CnnDecoder.py:189-194) as an implicit 3x3 convolution producing 4x4 pixel blocks
    gemm(ws.x[cur], batch * (g + 2) * (g + 2), w.last_c, 9, w.last_w, w.last_b, 64, VITAD_EPI_TANH_PIX4_F32, recon, 0);
    a.lda = S * w.last_c, a.conv_grid = g, a.a_taps = 0, a.split_out = 0;
    a.convt_w = g;
    return vitad_linear_f16(&a, s);
}
