// vitad_resize_bilinear_u8: the loader's transforms.Resize((S, S)) (src/data_loader/GeneralDataset.py:38-59) on the device.
// torchvision resizes the PIL image with Pillow's BILINEAR filter: a separable, antialiased triangle filter in 8-bit fixed
// point (coefficients int32 with 22 fractional bits, computed in double precision; horizontal pass, uint8 intermediate,
// vertical pass).  This file restates that arithmetic so the result is BIT-IDENTICAL to the CPU loader's:
//   vitad_resize_plan          host: per output pixel the first source pixel, the tap count and the integer coefficients
//   vitad_resize_bilinear_u8   device: horizontal pass HWC -> HWC (uint8), vertical pass HWC -> planar CHW (uint8), the
//                              layout vitad_deit_forward_u8 reads (ToTensor's /255 is folded into its patch gather).
// Both passes are HBM-bound byte kernels: one thread per output pixel (all 3 channels), consecutive threads on consecutive
// output x so a warp's source windows / destination bytes are contiguous.
#include <atomic>
#include <cmath>
#include <vector>

#include "host_util.cuh"
#include "ptx.cuh"

namespace vitad {
extern std::atomic<uint64_t> g_launches;

constexpr int kResizeBits = 32 - 8 - 2;  // Pillow Resample.c PRECISION_BITS

__device__ __forceinline__ uint8_t clip8(int v) {
    v >>= kResizeBits;
    return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// in [B*H][W][3] -> out [B*H][S][3]; plan = xmin[S] | count[S] | kk[S][ks]
__global__ void __launch_bounds__(256) resize_h_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                                       const int* __restrict__ plan, int ks, int rows, int W, int S) {
    griddep_launch_dependents();
    griddep_wait();
    const int xx = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = blockIdx.y;
    if (xx >= S || row >= rows) return;
    const int x0 = __ldg(plan + xx), n = __ldg(plan + S + xx);
    const int* kk = plan + 2 * S + xx * ks;
    const uint8_t* src = in + (static_cast<size_t>(row) * W + x0) * 3;
    int a0 = 1 << (kResizeBits - 1), a1 = a0, a2 = a0;
    for (int t = 0; t < n; ++t) {
        const int k = __ldg(kk + t);
        a0 += src[3 * t + 0] * k;
        a1 += src[3 * t + 1] * k;
        a2 += src[3 * t + 2] * k;
    }
    uint8_t* dst = out + (static_cast<size_t>(row) * S + xx) * 3;
    dst[0] = clip8(a0), dst[1] = clip8(a1), dst[2] = clip8(a2);
}

// in [B][H][S][3] -> out [B][3][S][S] (planar)
__global__ void __launch_bounds__(256) resize_v_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                                       const int* __restrict__ plan, int ks, int H, int S) {
    griddep_launch_dependents();
    griddep_wait();
    const int xx = blockIdx.x * blockDim.x + threadIdx.x;
    const int yy = blockIdx.y, b = blockIdx.z;
    if (xx >= S) return;
    const int y0 = __ldg(plan + yy), n = __ldg(plan + S + yy);
    const int* kk = plan + 2 * S + yy * ks;
    const uint8_t* src = in + ((static_cast<size_t>(b) * H + y0) * S + xx) * 3;
    int a0 = 1 << (kResizeBits - 1), a1 = a0, a2 = a0;
    for (int t = 0; t < n; ++t) {
        const int k = __ldg(kk + t);
        const uint8_t* p = src + static_cast<size_t>(t) * S * 3;
        a0 += p[0] * k;
        a1 += p[1] * k;
        a2 += p[2] * k;
    }
    const size_t plane = static_cast<size_t>(S) * S;
    uint8_t* dst = out + static_cast<size_t>(b) * 3 * plane + static_cast<size_t>(yy) * S + xx;
    dst[0] = clip8(a0), dst[plane] = clip8(a1), dst[2 * plane] = clip8(a2);
}
}  // namespace vitad

using namespace vitad;

// Taps per output pixel: Pillow precompute_coeffs, `ksize = (int)ceil(support) * 2 + 1` with support = max(in/out, 1).
extern "C" int vitad_resize_ksize(int in_size, int out_size) {
    if (in_size <= 0 || out_size <= 0) return 0;
    const double scale = static_cast<double>(in_size) / out_size;
    const double support = scale < 1.0 ? 1.0 : scale;
    return static_cast<int>(std::ceil(support)) * 2 + 1;
}

// Host-side plan (Pillow precompute_coeffs + normalize_coeffs_8bpc for the triangle filter over the whole image):
// plan[0..S) = first source pixel, plan[S..2S) = tap count, plan[2S + xx*ks + t] = coefficient t of output pixel xx.
extern "C" int vitad_resize_plan(int in_size, int out_size, int32_t* plan) {
    VITAD_REQUIRE(plan && in_size > 0 && out_size > 0, VITAD_ERR_ARG, "resize plan arguments");
    const int ks = vitad_resize_ksize(in_size, out_size);
    const double scale = static_cast<double>(in_size) / out_size;
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = 1.0 * filterscale;
    const double ss = 1.0 / filterscale;
    std::vector<double> w(ks);
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = (xx + 0.5) * scale;
        int xmin = static_cast<int>(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = static_cast<int>(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        double ww = 0.0;
        for (int x = 0; x < xmax; ++x) {
            double t = (x + xmin - center + 0.5) * ss;
            if (t < 0.0) t = -t;
            const double v = t < 1.0 ? 1.0 - t : 0.0;
            w[x] = v;
            ww += v;
        }
        for (int x = 0; x < xmax; ++x)
            if (ww != 0.0) w[x] /= ww;
        for (int x = xmax; x < ks; ++x) w[x] = 0.0;
        plan[xx] = xmin;
        plan[out_size + xx] = xmax;
        for (int x = 0; x < ks; ++x) {
            const double v = w[x] * (1 << kResizeBits);
            plan[2 * out_size + xx * ks + x] = w[x] < 0 ? static_cast<int>(-0.5 + v) : static_cast<int>(0.5 + v);
        }
    }
    return VITAD_OK;
}

extern "C" int vitad_resize_bilinear_u8(const uint8_t* in, int batch, int height, int width, int out_size,
                                        const int32_t* plan_h, const int32_t* plan_v, uint8_t* tmp, uint8_t* out,
                                        void* stream) {
    VITAD_NVTX("vitad_resize_bilinear_u8");
    int rc = check_device_arch();
    if (rc) return rc;
    VITAD_REQUIRE(in && out && tmp && plan_h && plan_v, VITAD_ERR_ARG, "null pointer");
    VITAD_REQUIRE(batch > 0 && height > 0 && width > 0 && out_size > 0 && batch <= 65535 && out_size <= 65535 &&
                      static_cast<long long>(batch) * height <= 0x7fffffffLL,
                  VITAD_ERR_SHAPE, "resize geometry");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int S = out_size;
    const int ks_h = vitad_resize_ksize(width, S), ks_v = vitad_resize_ksize(height, S);
    const int rows = batch * height;
    {
        ProfScope prof("resize_h", s);
        // grid.y is limited to 65535 rows per launch
        for (int r0 = 0; r0 < rows; r0 += 65535) {
            const int nr = rows - r0 < 65535 ? rows - r0 : 65535;
            VITAD_CUDA_OK(launch_pdl(resize_h_kernel, dim3((S + 255) / 256, nr), dim3(256), 0, s,
                                     in + static_cast<size_t>(r0) * width * 3, tmp + static_cast<size_t>(r0) * S * 3, plan_h,
                                     ks_h, nr, width, S));
            g_launches.fetch_add(1);
        }
    }
    {
        ProfScope prof("resize_v", s);
        VITAD_CUDA_OK(launch_pdl(resize_v_kernel, dim3((S + 255) / 256, S, batch), dim3(256), 0, s,
                                 static_cast<const uint8_t*>(tmp), out, plan_v, ks_v, height, S));
        g_launches.fetch_add(1);
    }
    return VITAD_OK;
}
