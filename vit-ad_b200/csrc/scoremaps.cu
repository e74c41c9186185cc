// Write-bandwidth-bound tail kernels that turn per-patch values into 224x224 anomaly maps and image scores:
//   * bilinear upsample g x g -> S x S in both PyTorch conventions, optional "1 - v" on input or output,
//     optional per-image max of the produced map
//       ValidatorMDN.py:137-162,171-172 (align_corners=True, 1 - interp(prob))
//       NormalizingFlow.py:134-143 + ValidatorNF.py:137-142 (align_corners=False, interp(1 - prob), amax)
//   * reconstruction L2 map: mean_c (recon - x)^2 and its per-image max
//       CnnAutoEncoder.py:49,68-74 + ValidatorRecon.py:109-116
// 128-bit stores, one thread per 4 output pixels; per-image max via warp shuffle + atomicMax on the
// (non-negative) float bit pattern, which is order independent and therefore deterministic.  A NaN anywhere in an image's
// map makes its score NaN, as torch.amax does in the reference (fmaxf alone would swallow it and an overflow upstream
// would turn into a plausible-looking AUROC): the quiet-NaN bit pattern is larger than every finite one, so the same
// atomicMax carries it.
#include <atomic>

#include "host_util.cuh"
#include "ptx.cuh"

namespace vitad {
extern std::atomic<uint64_t> g_launches;

constexpr unsigned int kNanBits = 0x7FC00000u;  // quiet NaN: above every finite (and infinite) non-negative bit pattern

__device__ __forceinline__ void src_index(int o, int in_size, float scale, bool align, int& i0, int& i1, float& lam) {
    // PyTorch area_pixel_compute_source_index + guard (UpSample.h)
    float s = align ? scale * o : fmaxf(scale * (o + 0.5f) - 0.5f, 0.f);
    i0 = min(static_cast<int>(s), in_size - 1);
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    lam = s - i0;
}

// in: [N, g, g] fp32; out: [N, S, S] fp32.  value = post(interp(pre(in))), pre/post = optional 1 - v.
__global__ void __launch_bounds__(256) bilinear_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                       unsigned int* __restrict__ img_max, int N, int g, int S,
                                                       int align, int pre_one_minus, int post_one_minus) {
    griddep_launch_dependents();
    griddep_wait();
    const int per_img = S * S / 4;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = blockIdx.y;
    float local_max = 0.f;
    bool saw_nan = false;
    if (idx < per_img) {
        const int oy = (idx * 4) / S, ox0 = (idx * 4) - oy * S;
        const float scale = align ? (S > 1 ? static_cast<float>(g - 1) / (S - 1) : 0.f) : static_cast<float>(g) / S;
        int y0, y1;
        float ly;
        src_index(oy, g, scale, align, y0, y1, ly);
        const float* p = in + static_cast<size_t>(n) * g * g;
        float r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int x0, x1;
            float lx;
            src_index(ox0 + j, g, scale, align, x0, x1, lx);
            float v00 = p[y0 * g + x0], v01 = p[y0 * g + x1], v10 = p[y1 * g + x0], v11 = p[y1 * g + x1];
            if (pre_one_minus) v00 = 1.f - v00, v01 = 1.f - v01, v10 = 1.f - v10, v11 = 1.f - v11;
            float v = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
            if (post_one_minus) v = 1.f - v;
            r[j] = v;
            local_max = fmaxf(local_max, v);
            saw_nan |= v != v;
        }
        *reinterpret_cast<float4*>(out + (static_cast<size_t>(n) * S + oy) * S + ox0) = make_float4(r[0], r[1], r[2], r[3]);
    }
    if (img_max) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) local_max = fmaxf(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
        saw_nan = __any_sync(0xffffffffu, saw_nan);
        if ((threadIdx.x & 31) == 0) atomicMax(img_max + n, saw_nan ? kNanBits : __float_as_uint(fmaxf(local_max, 0.f)));
    }
}

// recon, x: [N, C, HW] fp32 -> map [N, HW] = mean_c (recon - x)^2 ; img_max[n] = max over HW.
__global__ void __launch_bounds__(256) l2_map_kernel(const float* __restrict__ recon, const float* __restrict__ x,
                                                     float* __restrict__ map, unsigned int* __restrict__ img_max,
                                                     int C, int HW) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // float4 index within the image plane
    const int n = blockIdx.y;
    float local_max = 0.f;
    bool saw_nan = false;
    if (idx < HW / 4) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c = 0; c < C; ++c) {
            const size_t off = (static_cast<size_t>(n) * C + c) * HW;
            const float4 a = __ldg(reinterpret_cast<const float4*>(recon + off) + idx);
            const float4 b = __ldg(reinterpret_cast<const float4*>(x + off) + idx);
            const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z, dw = a.w - b.w;
            acc.x += dx * dx, acc.y += dy * dy, acc.z += dz * dz, acc.w += dw * dw;
        }
        acc.x /= C, acc.y /= C, acc.z /= C, acc.w /= C;
        reinterpret_cast<float4*>(map + static_cast<size_t>(n) * HW)[idx] = acc;
        local_max = fmaxf(fmaxf(acc.x, acc.y), fmaxf(acc.z, acc.w));
        saw_nan = acc.x != acc.x || acc.y != acc.y || acc.z != acc.z || acc.w != acc.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local_max = fmaxf(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
    saw_nan = __any_sync(0xffffffffu, saw_nan);
    if ((threadIdx.x & 31) == 0) atomicMax(img_max + n, saw_nan ? kNanBits : __float_as_uint(local_max));
}

}  // namespace vitad

using namespace vitad;

extern "C" int vitad_bilinear_up(const float* in, float* out, float* image_max, int n, int grid_in, int size_out,
                                 int align_corners, int pre_one_minus, int post_one_minus, void* stream) {
    VITAD_NVTX("vitad_bilinear_up");
    int rc = check_device_arch();
    if (rc) return rc;
    VITAD_REQUIRE(in && out && n > 0 && grid_in > 0, VITAD_ERR_ARG, "bilinear args");
    VITAD_REQUIRE(size_out % 4 == 0 && aligned16(out), VITAD_ERR_ALIGN, "output size %% 4 and 16-byte alignment");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (image_max) VITAD_CUDA_OK(cudaMemsetAsync(image_max, 0, sizeof(float) * n, s));
    const int per_img = size_out * size_out / 4;
    dim3 grid((per_img + 255) / 256, n);
    VITAD_CUDA_OK(launch_pdl(bilinear_kernel, grid, dim3(256), 0, s, in, out, reinterpret_cast<unsigned int*>(image_max), n,
                             grid_in, size_out, align_corners, pre_one_minus, post_one_minus));
    VITAD_CUDA_OK(cudaGetLastError());
    g_launches.fetch_add(1);
    return VITAD_OK;
}

extern "C" int vitad_l2_map_score(const float* recon, const float* x, float* map, float* image_max, int n, int channels,
                                  int hw, void* stream) {
    VITAD_NVTX("vitad_l2_map_score");
    int rc = check_device_arch();
    if (rc) return rc;
    VITAD_REQUIRE(recon && x && map && image_max && n > 0 && channels > 0, VITAD_ERR_ARG, "l2 map args");
    VITAD_REQUIRE(hw % 4 == 0 && aligned16(recon) && aligned16(x) && aligned16(map), VITAD_ERR_ALIGN,
                  "plane size %% 4 and 16-byte alignment");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    VITAD_CUDA_OK(cudaMemsetAsync(image_max, 0, sizeof(float) * n, s));
    dim3 grid((hw / 4 + 255) / 256, n);
    l2_map_kernel<<<grid, 256, 0, s>>>(recon, x, map, reinterpret_cast<unsigned int*>(image_max), channels, hw);
    VITAD_CUDA_OK(cudaGetLastError());
    g_launches.fetch_add(1);
    return VITAD_OK;
}
