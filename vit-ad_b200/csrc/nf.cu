// Normalizing-flow (FastFlow) head: src/classes/NormalizingFlow.py:84-145 with FrEIA 0.2 AllInOneBlock steps.
//
// Per step i (kernel 3x3 on even, 1x1 on odd steps, :96-100):
//   x1,x2 = split(x, 384|384);  h = relu(conv_k(x1));  a = 0.1*conv_k(h);  s = clamp*tanh(a[:, :384])
//   y = cat(x1, x2*exp(s) + a[:, 384:]) * scale + offset;   x' = y[:, perm]   (w_perm is a 0/1 matrix)
//   logdet += sum(s) + H*W*sum(log scale)
// Design:
//   * the activation stream lives channel-major in HBM (xT fp32 [768][tokens]) so the channel permutation is a
//     ROW permutation of the epilogue's (token-coalesced) stores — the reference spends 4.6 GFLOP/img on a dense
//     768x768 conv of zeros and ones for it;
//   * both subnet convolutions are tcgen05 GEMMs over token-major fp16 operands; the 3x3 first convolution (384
//     channels) is an implicit convolution over a zero-bordered operand (vitad_linear_args::conv_grid, no im2col), the
//     3x3 second one (64 hidden channels) reads im2col rows; conv1 uses the bias+ReLU epilogue, conv2's weight rows are interleaved so one
//     96-column accumulator tile holds [s(48 channels) | t(48 channels)], and its epilogue (EpiNfCouple) does the
//     coupling, the global affine, the permuted store and the per-token sum of s;
//   * log-det partial sums are written per (channel tile, token) and reduced in fixed order (deterministic).
#include <atomic>

#include "gemm_pair.cuh"
#include "host_util.cuh"

namespace vitad {
extern std::atomic<uint64_t> g_launches;
extern std::atomic<int> g_use_pair;

constexpr int kNfTile = 96;   // accumulator columns per tile: 48 s-channels | 48 t-channels
constexpr int kNfHalf = 48;

struct EpiNfCouple {
    static constexpr bool kSplitColumns = false;  // a thread needs the s and the t column of its channel
    const float* xin;    // [C][ld]
    float* xout;         // [C][ld]
    const float* b2p;    // [C] bias in interleaved tile order, pre-scaled by 0.1
    const float* scale;  // [C] 0.1*softplus_{0.5}(global_scale)
    const float* offset; // [C]
    const int* inv_perm; // [C] y channel j is stored at stream row inv_perm[j]
    float* sjac;         // [C/2/48][ld] per-token partial sums of s
    int ld, M, c_half;
    float clamp;
    float ssum;

    __device__ __forceinline__ void tile_begin(int, int, int) { ssum = 0.f; }
    __device__ __forceinline__ void merge(int, float2*) {}
    // the 2 x 8 stream loads of an 8-channel chunk: rows of the channel-major stream, coalesced over the warp's 32 tokens
    __device__ __forceinline__ void load_x(int n_tile, int q, int r, float (&x1a)[8], float (&x2a)[8]) const {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = n_tile * kNfHalf + q + j;
            x1a[j] = __ldg(xin + static_cast<size_t>(c) * ld + r);
            x2a[j] = __ldg(xin + static_cast<size_t>(c_half + c) * ld + r);
        }
    }
    // one 8-channel chunk: accumulators + constants -> coupling, global affine, permuted stores, sum of s
    __device__ __forceinline__ void chunk(int n_tile, int q, int row, bool valid, uint32_t taddr, const float (&x1a)[8],
                                          const float (&x2a)[8]) {
        uint32_t as[8], at[8];
        tmem_ld_x8(taddr + q, as);
        tmem_ld_x8(taddr + kNfHalf + q, at);
        // per-channel constants: 128-bit uniform loads (the scalar form issued 8 LDG per channel)
        const int cb = n_tile * kNfHalf + q;
        float bs[8], bt[8], sc1[8], of1[8], sc2[8], of2[8];
        int ip1[8], ip2[8];
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            ld4(bs + 4 * v, b2p + n_tile * kNfTile + q + 4 * v);
            ld4(bt + 4 * v, b2p + n_tile * kNfTile + kNfHalf + q + 4 * v);
            ld4(sc1 + 4 * v, scale + cb + 4 * v);
            ld4(of1 + 4 * v, offset + cb + 4 * v);
            ld4(sc2 + 4 * v, scale + c_half + cb + 4 * v);
            ld4(of2 + 4 * v, offset + c_half + cb + 4 * v);
            ld4i(ip1 + 4 * v, inv_perm + cb + 4 * v);
            ld4i(ip2 + 4 * v, inv_perm + c_half + cb + 4 * v);
        }
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float a_s = __uint_as_float(as[j]) + bs[j];
            const float a_t = __uint_as_float(at[j]) + bt[j];
            // clamp * tanh(a_s) = clamp * (1 - 2 / (1 + e^{2 a_s})), two MUFU (absolute error ~1e-7; e^{2a} = inf / 0
            // saturate to +-1); exp(s) = 2^{s log2 e}
            const float e2 = ex2f(a_s * 2.885390081777927f);
            const float sv = clamp * (1.0f - __fdividef(2.0f, 1.0f + e2));
            const float y2 = x2a[j] * ex2f(sv * 1.4426950408889634f) + a_t;
            if (valid) {
                xout[static_cast<size_t>(ip1[j]) * ld + row] = x1a[j] * sc1[j] + of1[j];
                xout[static_cast<size_t>(ip2[j]) * ld + row] = y2 * sc2[j] + of2[j];
            }
            ssum += sv;
        }
    }
    // Six 8-channel chunks, software-pipelined: the stream loads of chunk i+1 are in flight while chunk i is computed (their
    // L2 latency was the largest stall of this epilogue: 38 % of its warp samples in the ncu source page).  16-channel chunks
    // with the same pipelining spilled (ptxas keeps this kernel at 168 registers) and ran 35 % slower.
    __device__ __forceinline__ void sub(int, int, int n_tile, int row, uint32_t taddr, int, int) {
        static_assert(kNfHalf % 8 == 0, "8-channel chunks");
        const bool valid = row < M;
        const int r = valid ? row : 0;
        float xa[2][2][8];
        load_x(n_tile, 0, r, xa[0][0], xa[0][1]);
#pragma unroll
        for (int i = 0; i < kNfHalf / 8; ++i) {
            if (i + 1 < kNfHalf / 8) load_x(n_tile, 8 * (i + 1), r, xa[(i + 1) & 1][0], xa[(i + 1) & 1][1]);
            chunk(n_tile, 8 * i, row, valid, taddr, xa[i & 1][0], xa[i & 1][1]);
        }
    }
    __device__ __forceinline__ static void ld4(float* dst, const float* src) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src));
        dst[0] = v.x, dst[1] = v.y, dst[2] = v.z, dst[3] = v.w;
    }
    __device__ __forceinline__ static void ld4i(int* dst, const int* src) {
        const int4 v = __ldg(reinterpret_cast<const int4*>(src));
        dst[0] = v.x, dst[1] = v.y, dst[2] = v.z, dst[3] = v.w;
    }
    __device__ __forceinline__ void tile_end(int, int n_tile, int row) {
        if (row < M) sjac[static_cast<size_t>(n_tile) * ld + row] = ssum;
    }
};

// in [rows][cols] fp32 row-major -> out [cols][ld] (transpose), 32x32 tiles through shared memory.
__global__ void __launch_bounds__(256) transpose_f32_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                            int rows, int cols, int ld_out) {
    griddep_launch_dependents();
    griddep_wait();
    __shared__ float tile[32][33];
    const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8)
        if (r0 + i < rows && c0 + tx < cols) tile[i][tx] = in[static_cast<size_t>(r0 + i) * cols + c0 + tx];
    __syncthreads();
    for (int i = ty; i < 32; i += 8)
        if (c0 + i < cols && r0 + tx < rows) out[static_cast<size_t>(c0 + i) * ld_out + r0 + tx] = tile[tx][i];
}

// xT [.. >= c1][ld] fp32 -> out fp16 [M][c1] (first c1 channels, token-major GEMM operand).  pad_g > 0: token (b, y, x) of
// the pad_g x pad_g grid goes to row (b, y+1, x+1) of the zero-bordered layout the implicit 3x3 convolution reads.
// Block = 32 tokens x 128 channels: 128-byte row segments in, 256 contiguous bytes per token out (a thread packs 16
// channels into two 16-byte stores; the 32 x 32 tile version wrote 64-byte pieces).
constexpr int kOpCh = 128;
__global__ void __launch_bounds__(256) stream_to_operand_kernel(const float* __restrict__ xT, int ld,
                                                                __half* __restrict__ out, int M, int c1, int pad_g) {
    griddep_launch_dependents();
    griddep_wait();
    __shared__ float tile[kOpCh][33];
    const int t0 = blockIdx.x * 32, c0 = blockIdx.y * kOpCh;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll 4
    for (int i = ty; i < kOpCh; i += 8)
        tile[i][tx] = (c0 + i < c1 && t0 + tx < M) ? xT[static_cast<size_t>(c0 + i) * ld + t0 + tx] : 0.f;
    __syncthreads();
    const int tl = threadIdx.x >> 3, cg = (threadIdx.x & 7) * 16;  // token of the block, first of this thread's 16 channels
    if (t0 + tl >= M || c0 + cg >= c1) return;
    int row = t0 + tl;
    if (pad_g > 0) {
        const int x = row % pad_g, t = row / pad_g;
        row = ((t / pad_g) * (pad_g + 2) + t % pad_g + 1) * (pad_g + 2) + x + 1;
    }
    uint32_t u[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) u[j] = pack_h2(tile[cg + 2 * j][tl], tile[cg + 2 * j + 1][tl]);
    uint4* dst = reinterpret_cast<uint4*>(out + static_cast<size_t>(row) * c1 + c0 + cg);
    dst[0] = make_uint4(u[0], u[1], u[2], u[3]);
    dst[1] = make_uint4(u[4], u[5], u[6], u[7]);
}

// 3x3 im2col on a g x g grid per image, zero padding: in fp16 [M][cw] -> out fp16 [M][9*cw], column (tap, c),
// tap = ky*3+kx.  One thread moves 8 channels (16 bytes).
__global__ void __launch_bounds__(256) im2col3x3_kernel(const __half* __restrict__ in, __half* __restrict__ out, int M,
                                                        int cw, int g, size_t total) {
    griddep_launch_dependents();
    griddep_wait();
    const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int v_per_tap = cw >> 3;
    const int v = static_cast<int>(idx % v_per_tap);
    size_t r = idx / v_per_tap;
    const int tap = static_cast<int>(r % 9);
    const int t = static_cast<int>(r / 9);
    const int p = t % (g * g);
    const int y = p / g + tap / 3 - 1, x = p % g + tap % 3 - 1;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (y >= 0 && y < g && x >= 0 && x < g) {
        const int src = t - p + y * g + x;
        val = *reinterpret_cast<const uint4*>(in + static_cast<size_t>(src) * cw + v * 8);
    }
    *reinterpret_cast<uint4*>(out + static_cast<size_t>(t) * 9 * cw + tap * cw + v * 8) = val;
}

// Per token: omp[t] = 1 - exp(-0.5*mean_c z^2) (NormalizingFlow.py:134-137), tok[0][t] = sum_c z^2, tok[1][t] = sum of the
// log-det partials.  Block = 32 tokens (lane) x 8 channel groups (warp): every load is a coalesced 128-byte row segment,
// each thread has C/8 independent loads in flight instead of one thread walking all C channels of a token; the 8 partial
// sums are combined in fixed order (deterministic, independent of the batch size).
__global__ void __launch_bounds__(256) nf_token_kernel(const float* __restrict__ zT, int ld, const float* __restrict__ sjac,
                                                       int n_partials, int M, int C, float* __restrict__ omp,
                                                       float* __restrict__ tok) {
    griddep_launch_dependents();
    griddep_wait();
    __shared__ float red_z[8][32], red_s[8][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int t = blockIdx.x * 32 + lane;
    float zz = 0.f, sv = 0.f;
    if (t < M) {
        const int c0 = w * (C / 8), c1 = w == 7 ? C : c0 + C / 8;
        float z4[4] = {0.f, 0.f, 0.f, 0.f};
        int c = c0;
        for (; c + 4 <= c1; c += 4) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float z = zT[static_cast<size_t>(c + j) * ld + t];
                z4[j] = fmaf(z, z, z4[j]);
            }
        }
        for (; c < c1; ++c) {
            const float z = zT[static_cast<size_t>(c) * ld + t];
            z4[0] = fmaf(z, z, z4[0]);
        }
        zz = (z4[0] + z4[1]) + (z4[2] + z4[3]);
        for (int k = w; k < n_partials; k += 8) sv += sjac[static_cast<size_t>(k) * ld + t];
    }
    red_z[w][lane] = zz;
    red_s[w][lane] = sv;
    __syncthreads();
    if (w == 0 && t < M) {
        float z = 0.f, sum = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) z += red_z[i][lane], sum += red_s[i][lane];
        omp[t] = 1.0f - expf(-0.5f * (z / C));
        tok[t] = z;
        tok[ld + t] = sum;
    }
}

// One warp per image: loss_term[b] = 0.5*sum z^2 - (sum of s partials + logdet_const) (NormalizingFlow.py:130-132).
__global__ void __launch_bounds__(32) nf_image_kernel(const float* __restrict__ tok, int ld, float logdet_const, int P,
                                                      float* __restrict__ loss_terms) {
    griddep_launch_dependents();
    griddep_wait();
    const int b = blockIdx.x;
    float z = 0.f, sv = 0.f;
    for (int p = threadIdx.x; p < P; p += 32) {
        z += tok[b * P + p];
        sv += tok[ld + b * P + p];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        z += __shfl_xor_sync(0xffffffffu, z, o);
        sv += __shfl_xor_sync(0xffffffffu, sv, o);
    }
    if (threadIdx.x == 0) loss_terms[b] = 0.5f * z - (sv + logdet_const);
}

static int launch_couple(const void* a2, const void* w2p, int M, int C, int K2, const EpiNfCouple& epi,
                         cudaStream_t stream) {
    CUtensorMap ta, tb;
    int rc = make_tmap_f16_2d(&ta, a2, M, K2, K2, kBlockM);
    if (rc) return rc;
    const int n_tiles = C / kNfTile;
    if (g_use_pair.load() && M > kBlockM) {
        using S = PairSmem<kNfTile>;
        rc = make_tmap_f16_2d(&tb, w2p, C, K2, K2, kNfTile / 2);
        if (rc) return rc;
        auto kern = gemm2_tc_kernel<kNfTile, 1, EpiNfCouple>;
        static bool attr_set = false;
        if (!attr_set) {
            VITAD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotalBytes));
            attr_set = true;
        }
        const int tiles = ((M + 2 * kBlockM - 1) / (2 * kBlockM)) * n_tiles;
        const int maxc = device_sm_count() / 2;
        const int clusters = tiles < maxc ? tiles : maxc;
        VITAD_CUDA_OK(launch_pdl(kern, dim3(2 * clusters), dim3(kGemmThreads), S::kTotalBytes, stream, ta, tb, M, n_tiles, K2, epi));
    } else {
        using S = GemmSmem<kNfTile>;
        rc = make_tmap_f16_2d(&tb, w2p, C, K2, K2, kNfTile);
        if (rc) return rc;
        auto kern = gemm_tc_kernel<kNfTile, 1, EpiNfCouple>;
        static bool attr_set = false;
        if (!attr_set) {
            VITAD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotalBytes));
            attr_set = true;
        }
        const int tiles = ((M + kBlockM - 1) / kBlockM) * n_tiles;
        const int grid = tiles < device_sm_count() ? tiles : device_sm_count();
        VITAD_CUDA_OK(launch_pdl(kern, dim3(grid), dim3(kGemmThreads), S::kTotalBytes, stream, ta, tb, M, n_tiles, K2, epi));
    }
    VITAD_CUDA_OK(cudaGetLastError());
    g_launches.fetch_add(1);
    return VITAD_OK;
}

namespace {
struct NfWs {
    float *xa, *xb, *sjac, *tok;
    void *x1h, *x1p, *h, *a2;
    int ld;
    size_t total;
};
NfWs carve_nf(const vitad_nf_weights& w, int batch, void* base) {
    uint8_t* p = static_cast<uint8_t*>(base);
    size_t used = 0;
    auto take = [&](size_t bytes) {
        void* r = p ? p + used : nullptr;
        used += (bytes + 255) & ~static_cast<size_t>(255);
        return r;
    };
    const size_t M = static_cast<size_t>(batch) * w.grid * w.grid;
    NfWs s;
    s.ld = static_cast<int>((M + 31) / 32 * 32);
    const int c1 = w.channels - w.channels / 2;
    s.xa = static_cast<float*>(take(static_cast<size_t>(w.channels) * s.ld * 4));
    s.xb = static_cast<float*>(take(static_cast<size_t>(w.channels) * s.ld * 4));
    s.sjac = static_cast<float*>(take(static_cast<size_t>(w.steps) * (w.channels / kNfTile) * s.ld * 4));
    s.tok = static_cast<float*>(take(static_cast<size_t>(2) * s.ld * 4));
    s.x1h = take(M * c1 * 2);
    s.x1p = take(static_cast<size_t>(batch) * (w.grid + 2) * (w.grid + 2) * c1 * 2);  // zero-bordered x1 of the 3x3 steps
    s.h = take(M * w.hidden_pad * 2);
    s.a2 = take(M * 9 * w.hidden_pad * 2);
    s.total = used;
    return s;
}
}  // namespace

}  // namespace vitad

using namespace vitad;

extern "C" size_t vitad_nf_workspace_bytes(const vitad_nf_weights* w, int batch) {
    if (!w || batch <= 0) return 0;
    return carve_nf(*w, batch, nullptr).total;
}

extern "C" int vitad_nf_forward(const vitad_nf_weights* wp, const float* tokens, int batch, void* workspace,
                                size_t workspace_bytes, float* one_minus_prob, float* loss_terms, void* stream) {
    VITAD_NVTX("vitad_nf_forward");
    int rc = check_device_arch();
    if (rc) return rc;
    VITAD_REQUIRE(wp && tokens && workspace && one_minus_prob && loss_terms, VITAD_ERR_ARG, "null pointer");
    const vitad_nf_weights& w = *wp;
    VITAD_REQUIRE(w.channels == 768 && w.hidden_pad == 64 && w.grid > 0 && w.steps > 0 && w.step, VITAD_ERR_SHAPE,
                  "unsupported flow geometry (channels %d hidden_pad %d)", w.channels, w.hidden_pad);
    VITAD_REQUIRE(batch > 0, VITAD_ERR_SHAPE, "empty batch");
    VITAD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, VITAD_ERR_ALIGN, "workspace alignment");
    NfWs ws = carve_nf(w, batch, workspace);
    VITAD_REQUIRE(workspace_bytes >= ws.total, VITAD_ERR_WORKSPACE, "workspace %zu < %zu", workspace_bytes, ws.total);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int P = w.grid * w.grid, M = batch * P, C = w.channels, c_half = C / 2, c1 = C - c_half;

    // tokens [M][C] -> channel-major stream (the reference's transpose(2,1).reshape, ValidatorNF.py:130-134)
    {
        ProfScope prof("nf_transpose", s);
        dim3 grid((M + 31) / 32, (C + 31) / 32);
        VITAD_CUDA_OK(launch_pdl(transpose_f32_kernel, grid, dim3(256), 0, s, tokens, ws.xa, M, C, ws.ld));
        VITAD_CUDA_OK(cudaGetLastError());
        g_launches.fetch_add(1);
    }
    float* xin = ws.xa;
    float* xout = ws.xb;
    const int Mp = batch * (w.grid + 2) * (w.grid + 2);
    bool bordered = false;  // ws.x1p's border has been zeroed in this call
    for (int i = 0; i < w.steps; ++i) {
        const vitad_nf_step& st = w.step[i];
        VITAD_REQUIRE(st.ksize == 1 || st.ksize == 3, VITAD_ERR_SHAPE, "subnet kernel size %d", st.ksize);
        // subnet conv 1 (+ReLU): 1x1 = plain GEMM; 3x3 = implicit convolution over the zero-bordered operand (no im2col)
        const bool k3 = st.ksize == 3;
        if (k3 && !bordered) {
            VITAD_CUDA_OK(cudaMemsetAsync(ws.x1p, 0, static_cast<size_t>(Mp) * c1 * 2, s));
            bordered = true;
        }
        {
            ProfScope prof("nf_operand", s);
            dim3 grid((M + 31) / 32, (c1 + kOpCh - 1) / kOpCh);
            VITAD_CUDA_OK(launch_pdl(stream_to_operand_kernel, grid, dim3(256), 0, s, static_cast<const float*>(xin), ws.ld,
                                     static_cast<__half*>(k3 ? ws.x1p : ws.x1h), M, c1, k3 ? w.grid : 0));
            VITAD_CUDA_OK(cudaGetLastError());
            g_launches.fetch_add(1);
        }
        vitad_linear_args la;
        memset(&la, 0, sizeof(la));
        la.a = k3 ? ws.x1p : ws.x1h, la.w = st.w0p, la.bias = st.b0p, la.m = k3 ? Mp : M, la.n = w.hidden_pad;
        la.k = k3 ? 9 * c1 : c1, la.lda = c1, la.ldw = la.k, la.conv_grid = k3 ? w.grid : 0;
        la.epilogue = VITAD_EPI_BIAS_RELU_F16, la.out = ws.h, la.ldo = w.hidden_pad;
        if ((rc = vitad_linear_f16(&la, s))) return rc;
        const void* a2 = ws.h;
        int k2 = w.hidden_pad;
        if (st.ksize == 3) {
            ProfScope prof("nf_im2col", s);
            const size_t total = static_cast<size_t>(M) * 9 * (w.hidden_pad / 8);
            VITAD_CUDA_OK(launch_pdl(im2col3x3_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, s,
                                     static_cast<const __half*>(ws.h), static_cast<__half*>(ws.a2), M, w.hidden_pad, w.grid, total));
            VITAD_CUDA_OK(cudaGetLastError());
            g_launches.fetch_add(1);
            a2 = ws.a2;
            k2 = 9 * w.hidden_pad;
        }
        EpiNfCouple epi{xin,     xout,  st.b2p, st.scale, st.offset, st.inv_perm,
                        ws.sjac + static_cast<size_t>(i) * (C / kNfTile) * ws.ld, ws.ld, M, c_half, w.clamp, 0.f};
        {
            ProfScope prof("nf_couple_gemm", s);
            if ((rc = launch_couple(a2, st.w2p, M, C, k2, epi, s))) return rc;
        }
        float* t = xin;
        xin = xout;
        xout = t;
    }
    {
        ProfScope prof("nf_finish", s);
        VITAD_CUDA_OK(launch_pdl(nf_token_kernel, dim3((M + 31) / 32), dim3(256), 0, s, static_cast<const float*>(xin), ws.ld,
                                 static_cast<const float*>(ws.sjac), w.steps * (C / kNfTile), M, C, one_minus_prob, ws.tok));
        VITAD_CUDA_OK(launch_pdl(nf_image_kernel, dim3(batch), dim3(32), 0, s, static_cast<const float*>(ws.tok), ws.ld,
                                 w.logdet_const, P, loss_terms));
        VITAD_CUDA_OK(cudaGetLastError());
        g_launches.fetch_add(2);
    }
    return VITAD_OK;
}
