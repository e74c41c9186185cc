#include "host_util.cuh"

#include <nvtx3/nvToolsExt.h>
#include <stdarg.h>

#include <atomic>
#include <mutex>

namespace vitad {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_use_pair{1};
std::atomic<int> g_use_pdl{1};
bool pdl_enabled() { return g_use_pdl.load() != 0; }

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    });
    return fn;
}

// Encoded tensor maps are cached by their defining tuple: the scoring chain re-launches the same ~200 (buffer, shape,
// box) combinations every batch (weights and workspaces keep their addresses), and cuTensorMapEncodeTiled costs
// 1-2 us of host time per call — a third of the launch path at batch 1.  A map holds no device state, so a cached copy
// stays valid as long as the same address is used with the same geometry.
namespace {
struct TmapKey {
    const void* base;
    uint64_t d2, rows, cols, ld, ld2;
    uint32_t box_rows, box_cols;
    bool operator==(const TmapKey& o) const {
        return base == o.base && d2 == o.d2 && rows == o.rows && cols == o.cols && ld == o.ld && ld2 == o.ld2 &&
               box_rows == o.box_rows && box_cols == o.box_cols;
    }
};
struct TmapSlot {
    TmapKey key;
    CUtensorMap map;
    bool used;
};
constexpr int kTmapSlots = 1024;  // direct-mapped, overwritten on collision
thread_local TmapSlot g_tmaps[kTmapSlots];
inline size_t tmap_hash(const TmapKey& k) {
    uint64_t h = reinterpret_cast<uint64_t>(k.base) * 0x9E3779B97F4A7C15ull;
    h ^= (k.rows * 0xC2B2AE3D27D4EB4Full) ^ (k.cols << 17) ^ (k.ld << 29) ^ (static_cast<uint64_t>(k.box_rows) << 41) ^
         (k.d2 << 7) ^ (k.ld2 << 11);
    h ^= h >> 31;
    return static_cast<size_t>(h % kTmapSlots);
}
}  // namespace

int make_tmap_f16_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols) {
    const TmapKey key{base, 0, rows, cols, ld, 0, box_rows, box_cols};
    TmapSlot& slot = g_tmaps[tmap_hash(key)];
    if (slot.used && slot.key == key) {
        *map = slot.map;
        return VITAD_OK;
    }
    auto enc = get_encode();
    VITAD_REQUIRE(enc != nullptr, VITAD_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    VITAD_REQUIRE(aligned16(base) && (ld * 2) % 16 == 0, VITAD_ERR_ALIGN,
                  "TMA operand needs a 16-byte aligned base and pitch (base=%p ld=%llu)", base,
                  (unsigned long long)ld);
    VITAD_REQUIRE(box_rows >= 1 && box_rows <= 256 && box_cols * 2 == 128, VITAD_ERR_SHAPE, "bad TMA box %ux%u",
                  box_rows, box_cols);
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstr[1] = {ld * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VITAD_REQUIRE(r == CUDA_SUCCESS, VITAD_ERR_CUDA, "cuTensorMapEncodeTiled(2d) failed: CUresult %d", (int)r);
    slot.key = key;
    slot.map = *map;
    slot.used = true;
    return VITAD_OK;
}

int make_tmap_2d(CUtensorMap* map, const void* base, int elem_bytes, uint64_t rows, uint64_t cols, uint64_t ld,
                  uint32_t box_rows, uint32_t box_cols, bool swizzle128) {
    // cached like the fp16 maps; d2 / ld2 carry the extra parameters in the key
    const TmapKey key{base, 0x8000000000000000ull | static_cast<uint64_t>(elem_bytes) << 8 | (swizzle128 ? 1u : 0u), rows, cols, ld,
                      ~0ull, box_rows, box_cols};
    TmapSlot& slot = g_tmaps[tmap_hash(key)];
    if (slot.used && slot.key == key) {
        *map = slot.map;
        return VITAD_OK;
    }
    auto enc = get_encode();
    VITAD_REQUIRE(enc != nullptr, VITAD_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    VITAD_REQUIRE((elem_bytes == 2 || elem_bytes == 4) && aligned16(base) && (ld * elem_bytes) % 16 == 0, VITAD_ERR_ALIGN,
                  "TMA operand needs a 16-byte aligned base and pitch (base=%p ld=%llu)", base, (unsigned long long)ld);
    VITAD_REQUIRE(box_rows >= 1 && box_rows <= 256 && box_cols * elem_bytes <= 128 && (box_cols * elem_bytes) % 16 == 0 &&
                      (!swizzle128 || box_cols * elem_bytes == 128),
                  VITAD_ERR_SHAPE, "bad TMA box %ux%u", box_rows, box_cols);
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstr[1] = {ld * elem_bytes};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                     const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VITAD_REQUIRE(r == CUDA_SUCCESS, VITAD_ERR_CUDA, "cuTensorMapEncodeTiled(2d any) failed: CUresult %d", (int)r);
    slot.key = key;
    slot.map = *map;
    slot.used = true;
    return VITAD_OK;
}

int make_tmap_f16_3d(CUtensorMap* map, const void* base, uint64_t d2, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint64_t ld2, uint32_t box_rows, uint32_t box_cols) {
    const TmapKey key{base, d2 == 0 ? ~0ull : d2, rows, cols, ld, ld2, box_rows, box_cols};
    TmapSlot& slot = g_tmaps[tmap_hash(key)];
    if (slot.used && slot.key == key) {
        *map = slot.map;
        return VITAD_OK;
    }
    auto enc = get_encode();
    VITAD_REQUIRE(enc != nullptr, VITAD_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    VITAD_REQUIRE(aligned16(base) && (ld * 2) % 16 == 0 && (ld2 * 2) % 16 == 0, VITAD_ERR_ALIGN,
                  "TMA operand needs 16-byte aligned base and pitches");
    VITAD_REQUIRE(box_rows >= 1 && box_rows <= 256 && box_cols * 2 == 128, VITAD_ERR_SHAPE, "bad TMA box %ux%u",
                  box_rows, box_cols);
    cuuint64_t gdim[3] = {cols, rows, d2};
    cuuint64_t gstr[2] = {ld * 2, ld2 * 2};
    cuuint32_t box[3] = {box_cols, box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VITAD_REQUIRE(r == CUDA_SUCCESS, VITAD_ERR_CUDA, "cuTensorMapEncodeTiled(3d) failed: CUresult %d", (int)r);
    slot.key = key;
    slot.map = *map;
    slot.used = true;
    return VITAD_OK;
}

int device_sm_count() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return sms;
}

int check_device_arch() {
    static int major = -1;
    if (major < 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
            major = -1;
            set_error("no CUDA device available (this library has no CPU path)");
            return VITAD_ERR_ARCH;
        }
    }
    VITAD_REQUIRE(major == 10, VITAD_ERR_ARCH, "device compute capability %d.x is not sm_100 (B200)", major);
    return VITAD_OK;
}

}  // namespace vitad

extern "C" const char* vitad_last_error(void) { return vitad::g_err; }
extern "C" int vitad_abi_version(void) { return 1; }
extern "C" uint64_t vitad_launch_count(void) { return vitad::g_launches.load(); }
extern "C" void vitad_set_cta_pair(int enable) { vitad::g_use_pair.store(enable ? 1 : 0); }
extern "C" void vitad_set_pdl(int enable) { vitad::g_use_pdl.store(enable ? 1 : 0); }

// ------------------------------------------------------------------------------------ profiler
#include <map>
#include <string>
#include <vector>
namespace vitad {
struct ProfRec {
    std::string name;
    cudaEvent_t e0, e1;
};
static std::atomic<int> g_prof_on{0};
static std::vector<ProfRec> g_prof;
NvtxRange::NvtxRange(const char* name) { nvtxRangePushA(name); }
NvtxRange::~NvtxRange() { nvtxRangePop(); }
ProfScope::ProfScope(const char* name, cudaStream_t s) : idx(-1), stream(s), nvtx(name) {
    if (!g_prof_on.load()) return;
    ProfRec r;
    r.name = name;
    cudaEventCreate(&r.e0);
    cudaEventCreate(&r.e1);
    cudaEventRecord(r.e0, s);
    g_prof.push_back(r);
    idx = static_cast<int>(g_prof.size()) - 1;
}
ProfScope::~ProfScope() {
    if (idx >= 0) cudaEventRecord(g_prof[idx].e1, stream);
}
}  // namespace vitad

extern "C" void vitad_profile_enable(int on) {
    using namespace vitad;
    for (auto& r : g_prof) {
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    g_prof.clear();
    g_prof_on.store(on ? 1 : 0);
}

// Writes "name count total_us\n" lines plus a "__span__" line (first start .. last end) into buf.
extern "C" int vitad_profile_report(char* buf, int size) {
    using namespace vitad;
    cudaDeviceSynchronize();
    std::map<std::string, std::pair<int, double>> agg;
    std::vector<std::string> order;
    for (auto& r : g_prof) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.e0, r.e1);
        if (!agg.count(r.name)) order.push_back(r.name);
        agg[r.name].first += 1;
        agg[r.name].second += ms * 1e3;
    }
    int off = 0;
    for (auto& n : order)
        off += snprintf(buf + off, off < size ? size - off : 0, "%s %d %.1f\n", n.c_str(), agg[n].first, agg[n].second);
    if (!g_prof.empty()) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, g_prof.front().e0, g_prof.back().e1);
        off += snprintf(buf + off, off < size ? size - off : 0, "__span__ 1 %.1f\n", ms * 1e3);
    }
    return off;
}
