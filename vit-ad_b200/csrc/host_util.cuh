// Host-side helpers shared by the C-ABI entry points: error codes, TMA tensor-map encoding
// (driver entry point fetched through the runtime, no -lcuda link), launch checks.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/vitad.h"

namespace vitad {

void set_error(const char* fmt, ...);

#define VITAD_REQUIRE(cond, code, ...)       \
    do {                                     \
        if (!(cond)) {                       \
            ::vitad::set_error(__VA_ARGS__); \
            return (code);                   \
        }                                    \
    } while (0)

#define VITAD_CUDA_OK(expr)                                                                      \
    do {                                                                                         \
        cudaError_t e__ = (expr);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            ::vitad::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                               __LINE__);                                                        \
            return VITAD_ERR_CUDA;                                                               \
        }                                                                                        \
    } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// 2-D fp16 row-major tensor [rows, cols] with row pitch `ld` elements; box = [box_rows, 64 cols],
// 128-byte swizzle (matches the shared-memory matrix descriptors of ptx.cuh).  Out-of-bounds box elements read as zero.
int make_tmap_f16_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols = 64);
// General 2-D map: element type fp16 (elem_bytes 2) or fp32 (4), box [box_rows, box_cols], 128-byte swizzle or none.
int make_tmap_2d(CUtensorMap* map, const void* base, int elem_bytes, uint64_t rows, uint64_t cols, uint64_t ld,
                  uint32_t box_rows, uint32_t box_cols, bool swizzle128);
// 3-D variant: [d2, rows, cols] with pitches ld (elements, rows) and ld2 (elements, slabs).
int make_tmap_f16_3d(CUtensorMap* map, const void* base, uint64_t d2, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint64_t ld2, uint32_t box_rows, uint32_t box_cols = 64);

// Launch with programmatic stream serialization (see ptx.cuh: griddep_*).  Only for kernels that call griddep_wait()
// before touching global memory.  vitad_set_pdl(0) turns the attribute off (plain stream order).
bool pdl_enabled();
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

int device_sm_count();
int check_device_arch();  // VITAD_OK only on compute capability 10.x

}  // namespace vitad

// ---- tracing: NVTX ranges (no-ops unless a tool such as nsys/ncu injects an NVTX library) around every
// C-ABI entry point (VITAD_NVTX) and every launch site (ProfScope), plus the optional in-library profiler (CUDA events
// around every launch site; off by default) ----
namespace vitad {
struct NvtxRange {
    explicit NvtxRange(const char* name);
    ~NvtxRange();
};
#define VITAD_NVTX(name) ::vitad::NvtxRange nvtx_range__(name)
struct ProfScope {
    ProfScope(const char* name, cudaStream_t stream);
    ~ProfScope();
    int idx;
    cudaStream_t stream;
    NvtxRange nvtx;
};
}  // namespace vitad
