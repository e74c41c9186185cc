// vitad_linear_f16: dispatch of the tcgen05 GEMM main loops with the encoder's fused epilogues.
//
// Tile choice (M > 128: CTA-pair kernel of gemm_staged.cuh).  The encoder's GEMMs are small (M = 6336 at batch
// 32, N in {768, 2304, 3072}): narrow tiles are bound by L2->SM operand traffic, wide tiles by wave quantisation
// (N = 768 at BLOCK_N = 256: 75 tiles for 74 SM pairs).  make_sched() runs the full waves on 256 x BLOCK_N tiles
// and cuts the tiles of the last, partial wave into 2/4/8 narrower ones so that wave is short and full;
// pick_block_n() compares BLOCK_N = 128 and 256 under that schedule.  M <= 128 uses the single-CTA kernel.
#include <atomic>

#include "gemm_epilogues.cuh"
#include "gemm_pair.cuh"
#include "gemm_staged.cuh"
#include "host_util.cuh"

namespace vitad {
extern std::atomic<uint64_t> g_launches;
extern std::atomic<int> g_use_pair;
std::atomic<int> g_epi_warps{0};  // 0 = per-epilogue default, 8 / 16 = forced (diagnostics)
static double g_cost192 = 1.10;

// Full waves of 256 x bn tiles, then the remaining tiles cut into `split` pieces each (<= one wave of pieces).
static TileSched make_sched(int m, int n, int bn, int clusters) {
    TileSched s;
    s.num_m = (m + 2 * kBlockM - 1) / (2 * kBlockM);
    const int tiles = s.num_m * ((n + bn - 1) / bn);
    const int rest = tiles % clusters;
    s.big_tiles = tiles - rest;
    s.tail_split = 1;
    // a piece is a whole number of 32-column epilogue units
    while (s.tail_split < 8 && rest * s.tail_split * 2 <= clusters && bn / (s.tail_split * 2) >= 32 &&
           (bn / (s.tail_split * 2)) % 32 == 0)
        s.tail_split *= 2;
    s.tail_tiles = rest * s.tail_split;
    s.tail_w = bn / s.tail_split;
    s.conv_pitch = 0;
    s.kpt = 0;
    s.tap_cols = 0;
    s.split_kb = 0;
    return s;
}
// Relative cost of a schedule in units of "one 256 x 256 tile": a cut piece still streams the whole 256-row A block,
// so a wave of pieces costs ~max(w/256, 0.55) of a full wave (measured: L2->SM delivery per SM, not MMA, bounds it).
static double sched_cost(const TileSched& s, int bn, int clusters) {
    const double full = static_cast<double>(s.big_tiles / clusters) * bn / 256.0;
    if (s.tail_tiles == 0) return full;
    const double piece = static_cast<double>(s.tail_w) / 256.0;
    return full + (piece > 0.55 * bn / 256.0 ? piece : 0.55 * bn / 256.0);
}

// CTA-pair launch with the staged epilogue (gemm_staged.cuh).
template <int BLOCK_N, int EPI_WARPS, class Epi>
static int launch_gemm_staged_w(const vitad_linear_args& a, const Epi& epi, cudaStream_t stream) {
    using S = StagedSmem<BLOCK_N, EPI_WARPS>;
    const int max_clusters = device_sm_count() / 2;
    TileSched sched = make_sched(a.m, a.n, BLOCK_N, max_clusters);
    // A's width and the K walk (TileSched): plain | implicit 3x3 convolution (row-shifted taps) | explicit im2col taps,
    // each optionally with split-fp16 operands (A rows [hi | lo] per tap, W rows [w_hi | w_hi | w_lo] per tap)
    const int taps = a.conv_grid > 0 ? 9 : (a.a_taps > 0 ? a.a_taps : 1);
    const int per_tap_k = a.k / taps;                                   // W columns per tap
    const int per_tap_a = a.split_c > 0 ? 2 * a.split_c : per_tap_k;    // A columns per tap
    const int a_cols = a.conv_grid > 0 ? per_tap_a : taps * per_tap_a;
    if (a.conv_grid > 0 || a.split_c > 0 || taps > 1) {
        sched.kpt = per_tap_k / kBlockK;
        sched.tap_cols = a.conv_grid > 0 ? 0 : per_tap_a;
        sched.split_kb = a.split_c > 0 ? a.split_c / kBlockK : 0;
        sched.conv_pitch = a.conv_grid > 0 ? a.conv_grid + 2 : 0;
    }
    CUtensorMap ta, tb, tbt;
    int rc = make_tmap_f16_2d(&ta, a.a, a.m, a_cols, a.lda, kBlockM);
    if (rc) return rc;
    rc = make_tmap_f16_2d(&tb, a.w, a.n, a.k, a.ldw, BLOCK_N / 2);
    if (rc) return rc;
    rc = make_tmap_f16_2d(&tbt, a.w, a.n, a.k, a.ldw, sched.tail_w / 2);
    if (rc) return rc;
    auto kern = gemm3_tc_kernel<BLOCK_N, EPI_WARPS, Epi>;
    static bool attr_set = false;
    if (!attr_set) {
        VITAD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotalBytes));
        attr_set = true;
    }
    const int tiles = sched.big_tiles + sched.tail_tiles;
    const int clusters = tiles < max_clusters ? tiles : max_clusters;
    VITAD_CUDA_OK(launch_pdl(kern, dim3(2 * clusters), dim3(64 + 32 * EPI_WARPS), S::kTotalBytes, stream, ta, tb, tbt, sched,
                             a.k, epi));
    VITAD_CUDA_OK(cudaGetLastError());
    g_launches.fetch_add(1);
    return VITAD_OK;
}

template <int BLOCK_N, class Epi>
static int launch_gemm_staged(const vitad_linear_args& a, const Epi& epi, cudaStream_t stream) {
    const int forced = g_epi_warps.load();
    const int warps = forced == 8 || forced == 16 ? forced : Epi::kDefaultEpiWarps;
    if (warps == 16) return launch_gemm_staged_w<BLOCK_N, 16>(a, epi, stream);
    return launch_gemm_staged_w<BLOCK_N, 8>(a, epi, stream);
}

template <int BLOCK_N>
static int dispatch_staged(const vitad_linear_args& a, cudaStream_t stream) {
    switch (a.epilogue) {
        case VITAD_EPI_BIAS_F16:
            if (a.conv_grid > 0 || a.out_pad_grid > 0)
                return launch_gemm_staged<BLOCK_N>(a, SEpiBiasHMap<0>{a.bias, static_cast<__half*>(a.out), a.ldo, a.m, a.n,
                                                                     a.conv_grid > 0 ? 2 : 1,
                                                                     a.conv_grid > 0 ? a.conv_grid : a.out_pad_grid, a.split_out}, stream);
            return launch_gemm_staged<BLOCK_N>(a, SEpiBiasH<0>{a.bias, static_cast<__half*>(a.out), a.ldo, a.m, a.n, a.split_out}, stream);
        case VITAD_EPI_BIAS_GELU_F16:
            return launch_gemm_staged<BLOCK_N>(a, SEpiBiasH<1>{a.bias, static_cast<__half*>(a.out), a.ldo, a.m, a.n, 0}, stream);
        case VITAD_EPI_BIAS_RELU_F16:
            if (a.conv_grid > 0 || a.out_pad_grid > 0)
                return launch_gemm_staged<BLOCK_N>(a, SEpiBiasHMap<2>{a.bias, static_cast<__half*>(a.out), a.ldo, a.m, a.n,
                                                                     a.conv_grid > 0 ? 2 : 1,
                                                                     a.conv_grid > 0 ? a.conv_grid : a.out_pad_grid, a.split_out}, stream);
            return launch_gemm_staged<BLOCK_N>(a, SEpiBiasH<2>{a.bias, static_cast<__half*>(a.out), a.ldo, a.m, a.n, a.split_out}, stream);
        case VITAD_EPI_RESIDUAL_F32:
            return launch_gemm_staged<BLOCK_N>(a, SEpiResidualF32{a.bias, a.resid, static_cast<float*>(a.out), a.ldo, a.m, a.n},
                                               stream);
        case VITAD_EPI_QKV: {
            const int hd = a.head_dim > 0 ? a.head_dim : 64;
            const int nw = a.windows > 0 ? a.windows : 1;
            const int wt = a.win_tokens > 0 ? a.win_tokens : a.tokens;
            SEpiQkv e{a.bias, static_cast<__half*>(a.q), static_cast<__half*>(a.kmat), static_cast<__half*>(a.vt),
                      a.tok2win, a.m, a.tokens, wt, a.tokens_pad, a.heads, hd, nw, a.q_scale, a.v_natural};
            return launch_gemm_staged<BLOCK_N>(a, e, stream);
        }
        case VITAD_EPI_PATCH_EMBED:
            return launch_gemm_staged<BLOCK_N>(
                a, SEpiPatchEmbed{a.bias, a.pos, static_cast<float*>(a.out), a.m, a.patches, a.prefix, a.n}, stream);
        case VITAD_EPI_F32:
            return launch_gemm_staged<BLOCK_N>(a, SEpiBiasF32{a.bias, static_cast<float*>(a.out), a.ldo, a.m, a.n}, stream);
        case VITAD_EPI_CONVT_RELU_F16:
            return launch_gemm_staged<BLOCK_N>(a, SEpiConvT{a.bias, static_cast<__half*>(a.out), a.m, a.n, a.convt_w, a.split_out}, stream);
        case VITAD_EPI_RES16_RELU_F16:
            return launch_gemm_staged<BLOCK_N>(a, SEpiResReluH{a.bias, static_cast<const __half*>(a.resid16),
                                                               static_cast<__half*>(a.out), a.ldo, a.ldr, a.m, a.n, a.res_grid,
                                                               a.out_pad_grid, a.split_out},
                                               stream);
        case VITAD_EPI_TANH_PIX4_F32:
            return launch_gemm_staged<BLOCK_N>(a, SEpiTanhPix4{a.bias, static_cast<float*>(a.out), a.m, a.convt_w, a.conv_grid > 0 ? 1 : 0},
                                               stream);
        default:
            set_error("unknown epilogue %d", a.epilogue);
            return VITAD_ERR_ARG;
    }
}

template <int BLOCK_N, class Epi>
static int launch_gemm(const vitad_linear_args& a, const Epi& epi, cudaStream_t stream) {
    using S = GemmSmem<BLOCK_N>;
    CUtensorMap ta, tb;
    int rc = make_tmap_f16_2d(&ta, a.a, a.m, a.k, a.lda, kBlockM);
    if (rc) return rc;
    rc = make_tmap_f16_2d(&tb, a.w, a.n, a.k, a.ldw, BLOCK_N);
    if (rc) return rc;
    auto kern = gemm_tc_kernel<BLOCK_N, 1, Epi>;
    static bool attr_set = false;
    if (!attr_set) {
        VITAD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotalBytes));
        attr_set = true;
    }
    const int num_m = (a.m + kBlockM - 1) / kBlockM;
    const int num_n = (a.n + BLOCK_N - 1) / BLOCK_N;
    const int tiles = num_m * num_n;
    const int grid = tiles < device_sm_count() ? tiles : device_sm_count();
    VITAD_CUDA_OK(launch_pdl(kern, dim3(grid), dim3(kGemmThreads), S::kTotalBytes, stream, ta, tb, a.m, num_n, a.k, epi));
    VITAD_CUDA_OK(cudaGetLastError());
    g_launches.fetch_add(1);
    return VITAD_OK;
}

template <int BLOCK_N>
static int dispatch_epilogue(const vitad_linear_args& a, cudaStream_t stream) {
    switch (a.epilogue) {
        case VITAD_EPI_BIAS_F16: {
            EpiBiasH<BLOCK_N, 0> e{a.bias, static_cast<__half*>(a.out), a.ldo, a.m, a.n};
            return launch_gemm<BLOCK_N>(a, e, stream);
        }
        case VITAD_EPI_BIAS_GELU_F16: {
            EpiBiasH<BLOCK_N, 1> e{a.bias, static_cast<__half*>(a.out), a.ldo, a.m, a.n};
            return launch_gemm<BLOCK_N>(a, e, stream);
        }
        case VITAD_EPI_BIAS_RELU_F16: {
            EpiBiasH<BLOCK_N, 2> e{a.bias, static_cast<__half*>(a.out), a.ldo, a.m, a.n};
            return launch_gemm<BLOCK_N>(a, e, stream);
        }
        case VITAD_EPI_RESIDUAL_F32: {
            EpiResidualF32<BLOCK_N> e{a.bias, a.resid, static_cast<float*>(a.out), a.ldo, a.m, a.n};
            return launch_gemm<BLOCK_N>(a, e, stream);
        }
        case VITAD_EPI_QKV: {
            const int hd = a.head_dim > 0 ? a.head_dim : 64;
            const int nw = a.windows > 0 ? a.windows : 1;
            const int wt = a.win_tokens > 0 ? a.win_tokens : a.tokens;
            EpiQkv<BLOCK_N> e{a.bias,
                              static_cast<__half*>(a.q),
                              static_cast<__half*>(a.kmat),
                              static_cast<__half*>(a.vt),
                              a.tok2win,
                              a.m,
                              a.tokens,
                              wt,
                              a.tokens_pad,
                              a.heads,
                              hd,
                              nw,
                              a.q_scale};
            return launch_gemm<BLOCK_N>(a, e, stream);
        }
        case VITAD_EPI_PATCH_EMBED: {
            EpiPatchEmbed<BLOCK_N> e{a.bias, a.pos, static_cast<float*>(a.out), a.m, a.patches, a.prefix, a.n};
            return launch_gemm<BLOCK_N>(a, e, stream);
        }
        case VITAD_EPI_F32: {
            EpiBiasF32<BLOCK_N> e{a.bias, static_cast<float*>(a.out), a.ldo, a.m, a.n};
            return launch_gemm<BLOCK_N>(a, e, stream);
        }
        default:
            set_error("unknown epilogue %d", a.epilogue);
            return VITAD_ERR_ARG;
    }
}

// Single-CTA kernel (M <= 128 or pairs disabled): minimise (waves) x (tile width); ties go to the wider tile.
static int pick_block_n_single(int m, int n) {
    const int workers = device_sm_count();
    const int num_m = (m + kBlockM - 1) / kBlockM;
    const int cand[3] = {256, 128, 96};
    int best = 256;
    long best_cost = -1;
    for (int bn : cand) {
        const long tiles = static_cast<long>(num_m) * ((n + bn - 1) / bn);
        const long cost = ((tiles + workers - 1) / workers) * bn;
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            best = bn;
        }
    }
    return best;
}
static int pick_block_n_pair(int m, int n) {
    const int clusters = device_sm_count() / 2;
    const double c256 = sched_cost(make_sched(m, n, 256, clusters), 256, clusters);
    const double c192 = sched_cost(make_sched(m, n, 192, clusters), 192, clusters) * g_cost192;  // narrower: more L2 traffic
    const double c128 = sched_cost(make_sched(m, n, 128, clusters), 128, clusters) * 1.25;
    int best = 256;
    double cost = c256;
    if (c192 < cost) best = 192, cost = c192;
    if (c128 < cost) best = 128, cost = c128;
    return best;
}

}  // namespace vitad

// Diagnostics: force 8 or 16 epilogue warps in the CTA-pair GEMM (0 = per-epilogue default).
extern "C" void vitad_set_epilogue_warps(int warps) { vitad::g_epi_warps.store(warps); }

#ifdef VITAD_TIMELINE
// Diagnostic builds only: point the CTA-pair GEMM's timeline stamps at a [grid][64] uint64 device buffer.
extern "C" int vitad_debug_timeline(void* device_buffer) {
    unsigned long long* p = static_cast<unsigned long long*>(device_buffer);
    VITAD_CUDA_OK(cudaMemcpyToSymbol(vitad::g_timeline, &p, sizeof(p)));
    return VITAD_OK;
}
#endif

extern "C" int vitad_linear_f16(const vitad_linear_args* args, void* stream) {
    using namespace vitad;
    VITAD_REQUIRE(args != nullptr, VITAD_ERR_ARG, "null args");
    const vitad_linear_args& a = *args;
    int rc = check_device_arch();
    if (rc) return rc;
    VITAD_REQUIRE(a.a && a.w, VITAD_ERR_ARG, "null operand");
    VITAD_REQUIRE(a.m > 0 && a.n > 0 && a.k > 0, VITAD_ERR_SHAPE, "empty GEMM %dx%dx%d", a.m, a.n, a.k);
    VITAD_REQUIRE(a.k % 16 == 0, VITAD_ERR_SHAPE, "K=%d must be a multiple of 16", a.k);
    VITAD_REQUIRE(a.conv_grid >= 0 && a.out_pad_grid >= 0 && !(a.conv_grid > 0 && a.out_pad_grid > 0), VITAD_ERR_ARG,
                  "conv_grid / out_pad_grid: at most one, non-negative");
    VITAD_REQUIRE(a.split_c >= 0 && a.a_taps >= 0 && (a.split_out == 0 || a.split_out == 1), VITAD_ERR_ARG,
                  "split_c / a_taps / split_out out of range");
    const int taps_chk = a.conv_grid > 0 ? 9 : (a.a_taps > 0 ? a.a_taps : 1);
    if (a.split_c > 0 || a.a_taps > 1)
        VITAD_REQUIRE((a.split_c == 0 || (a.split_c % kBlockK == 0 && a.k == taps_chk * 3 * a.split_c)) &&
                          (a.a_taps <= 1 || (a.conv_grid == 0 && a.k % (a.a_taps * kBlockK) == 0)),
                      VITAD_ERR_SHAPE, "split operands: K = taps * 3 * split_c with split_c %% 64 == 0; im2col taps: K %% (taps*64) == 0");
    if (a.split_out)
        VITAD_REQUIRE(a.epilogue == VITAD_EPI_BIAS_F16 || a.epilogue == VITAD_EPI_BIAS_RELU_F16 ||
                          a.epilogue == VITAD_EPI_CONVT_RELU_F16 || a.epilogue == VITAD_EPI_RES16_RELU_F16,
                      VITAD_ERR_ARG, "split_out: bias / ReLU / conv-transpose / residual fp16 epilogues only");
    if (a.conv_grid > 0) {
        const int P = a.conv_grid + 2;
        VITAD_REQUIRE(a.k % (9 * kBlockK) == 0 && a.m % (P * P) == 0 &&
                          (a.epilogue == VITAD_EPI_BIAS_F16 || a.epilogue == VITAD_EPI_BIAS_RELU_F16 ||
                           a.epilogue == VITAD_EPI_TANH_PIX4_F32),
                      VITAD_ERR_SHAPE, "implicit 3x3 convolution: K = 9*C with C %% 64 == 0, M = B*(g+2)^2, bias/ReLU/image-head epilogue");
        VITAD_REQUIRE(a.epilogue != VITAD_EPI_TANH_PIX4_F32 || a.convt_w == a.conv_grid, VITAD_ERR_SHAPE,
                      "image-head epilogue over an implicit convolution: convt_w must equal conv_grid");
    }
    if (a.out_pad_grid > 0)
        VITAD_REQUIRE(a.m % (a.out_pad_grid * a.out_pad_grid) == 0 &&
                          (a.epilogue == VITAD_EPI_BIAS_F16 || a.epilogue == VITAD_EPI_BIAS_RELU_F16 ||
                           a.epilogue == VITAD_EPI_RES16_RELU_F16),
                      VITAD_ERR_SHAPE, "out_pad_grid: M = B*g*g and a bias/ReLU/residual fp16 epilogue");
    const int a_width = (a.conv_grid > 0 ? 1 : taps_chk) * (a.split_c > 0 ? 2 * a.split_c : a.k / taps_chk);
    VITAD_REQUIRE(a.lda % 8 == 0 && a.ldw % 8 == 0 && a.lda >= a_width && a.ldw >= a.k, VITAD_ERR_ALIGN,
                  "pitches must cover the operands and be multiples of 8 elements (lda=%d >= %d, ldw=%d >= %d)", a.lda, a_width,
                  a.ldw, a.k);
    if (a.epilogue == VITAD_EPI_F32) {
        VITAD_REQUIRE(a.out && a.ldo >= a.n, VITAD_ERR_ARG, "bad fp32 output");
    } else {
        VITAD_REQUIRE(a.n % 32 == 0, VITAD_ERR_SHAPE, "N=%d must be a multiple of 32 for this epilogue", a.n);
        VITAD_REQUIRE(a.bias && aligned16(a.bias), VITAD_ERR_ARG, "bias missing or misaligned");
    }
    switch (a.epilogue) {
        case VITAD_EPI_BIAS_F16:
        case VITAD_EPI_BIAS_GELU_F16:
        case VITAD_EPI_BIAS_RELU_F16:
            VITAD_REQUIRE(a.out && aligned16(a.out) && a.ldo % 8 == 0 && a.ldo >= (a.split_out ? 2 : 1) * a.n, VITAD_ERR_ALIGN,
                          "fp16 output must be 16-byte aligned with pitch %% 8 == 0 (and >= 2N with split_out)");
            break;
        case VITAD_EPI_RESIDUAL_F32:
            VITAD_REQUIRE(a.out && a.resid && aligned16(a.out) && aligned16(a.resid) && a.ldo % 4 == 0 &&
                              a.ldo >= a.n,
                          VITAD_ERR_ALIGN, "fp32 residual/output must be 16-byte aligned");
            break;
        case VITAD_EPI_QKV:
            VITAD_REQUIRE(a.q && a.kmat && a.vt && aligned16(a.q) && aligned16(a.kmat), VITAD_ERR_ARG,
                          "q/k/vt missing or misaligned");
            VITAD_REQUIRE((a.head_dim == 0 || a.head_dim == 32 || a.head_dim == 64) && a.heads > 0 &&
                              a.n == 3 * a.heads * (a.head_dim > 0 ? a.head_dim : 64) && a.tokens > 0 &&
                              a.m % a.tokens == 0,
                          VITAD_ERR_SHAPE, "QKV epilogue needs N = 3*H*hd (hd 32 or 64) and M = B*tokens");
            VITAD_REQUIRE(!a.v_natural || (g_use_pair.load() && a.m > kBlockM), VITAD_ERR_SHAPE,
                          "v_natural needs the CTA-pair kernel (M > 128)");
            VITAD_REQUIRE((a.v_natural || a.tokens_pad >= (a.win_tokens > 0 ? a.win_tokens : a.tokens)) &&
                              (a.windows <= 1 || (a.tok2win && a.windows * a.win_tokens == a.tokens)),
                          VITAD_ERR_SHAPE, "QKV epilogue window layout (windows*win_tokens == tokens, map given)");
            break;
        case VITAD_EPI_CONVT_RELU_F16:
            VITAD_REQUIRE(a.out && aligned16(a.out) && a.n % 128 == 0 && a.convt_w > 0 && a.m % (a.convt_w * a.convt_w) == 0,
                          VITAD_ERR_SHAPE, "conv-transpose epilogue needs N = 4*Cp with Cp %% 32 == 0 and M = B*Wg*Wg");
            break;
        case VITAD_EPI_RES16_RELU_F16:
            VITAD_REQUIRE(a.out && a.resid16 && aligned16(a.out) && aligned16(a.resid16) && a.ldo % 8 == 0 &&
                              a.ldo >= (a.split_out ? 2 : 1) * a.n && a.ldr % 8 == 0 && a.ldr >= (a.split_out ? 2 : 1) * a.n &&
                              a.res_grid >= 0 &&
                              (a.res_grid == 0 || a.m % (4 * a.res_grid * a.res_grid) == 0),
                          VITAD_ERR_ALIGN, "fp16 residual epilogue: aligned out/resid16, pitches %% 8, M = B*(2*res_grid)^2");
            break;
        case VITAD_EPI_TANH_PIX4_F32:
            VITAD_REQUIRE(a.out && aligned16(a.out) && a.n == 64 && a.convt_w > 0 &&
                              (a.conv_grid > 0 || a.m % (a.convt_w * a.convt_w) == 0),
                          VITAD_ERR_SHAPE, "image-head epilogue needs N = 64 (48 live) and M = B*Wg*Wg");
            break;
        case VITAD_EPI_PATCH_EMBED:
            VITAD_REQUIRE(a.out && a.pos && aligned16(a.out) && aligned16(a.pos) && a.patches > 0 &&
                              a.m % a.patches == 0 && a.prefix >= 0,
                          VITAD_ERR_ARG, "patch-embed epilogue arguments");
            break;
        default:
            break;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // the conv-transpose scatter exists only as a staged epilogue of the CTA-pair kernel (which handles any M)
    const bool pair = (g_use_pair.load() && a.m > kBlockM) || a.epilogue == VITAD_EPI_CONVT_RELU_F16 ||
                      a.epilogue == VITAD_EPI_RES16_RELU_F16 || a.epilogue == VITAD_EPI_TANH_PIX4_F32 || a.conv_grid > 0 ||
                      a.out_pad_grid > 0 || a.split_c > 0 || a.a_taps > 1 || a.split_out;
    // block_n is a hint: widths the selected kernel does not instantiate fall back to the library's choice
    const bool hint_ok = pair ? (a.block_n == 128 || a.block_n == 192 || a.block_n == 256)
                              : (a.block_n == 96 || a.block_n == 128 || a.block_n == 256);
    const int bn = hint_ok ? a.block_n : (pair ? pick_block_n_pair(a.m, a.n) : pick_block_n_single(a.m, a.n));
    char pname[64];
    snprintf(pname, sizeof(pname), "gemm_epi%d_n%d_k%d_bn%d%s%s", a.epilogue, a.n, a.k, bn, a.conv_grid > 0 ? "_conv3x3" : "",
             a.split_c > 0 ? "_split" : "");
    ProfScope prof(pname, s);
    if (pair) {
        if (bn == 256) return dispatch_staged<256>(a, s);
        if (bn == 192) return dispatch_staged<192>(a, s);
        return dispatch_staged<128>(a, s);
    }
    if (bn == 256) return dispatch_epilogue<256>(a, s);
    if (bn == 128) return dispatch_epilogue<128>(a, s);
    if (bn == 96) return dispatch_epilogue<96>(a, s);
    set_error("block_n=%d unsupported (96, 128 or 256)", bn);
    return VITAD_ERR_SHAPE;
}
