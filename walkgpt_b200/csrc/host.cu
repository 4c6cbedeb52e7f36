// Host-side plumbing: error strings, device queries, TMA tensor-map encoding.
#include "host.h"

#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace wg {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || sym == nullptr) {
        set_error("cuTensorMapEncodeTiled is not available from the CUDA driver (%s)", cudaGetErrorString(e));
        return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(sym);
    return fn;
}

int make_tensor_map(CUtensorMap* out, const void* gptr, int elem_bytes, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box) {
    return make_tensor_map_sw(out, gptr, elem_bytes, rank, dims, strides_bytes, box, 128);
}

int make_tensor_map_sw(CUtensorMap* out, const void* gptr, int elem_bytes, int rank, const uint64_t* dims,
                       const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return WG_ERR_CUDA;
    cuuint64_t gdim[5];
    cuuint64_t gstr[4];
    cuuint32_t bx[5];
    cuuint32_t es[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
    }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    const CUtensorMapSwizzle sw = swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                      : swizzle_bytes == 0 ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B;
    CUresult r = fn(out, dt, (cuuint32_t)rank, const_cast<void*>(gptr), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank=%d dims=%llu,%llu stride0=%llu box=%u,%u ptr=%p)", (int)r,
                  rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                  (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), box[0], rank > 1 ? box[1] : 0, gptr);
        return WG_ERR_CUDA;
    }
    return WG_OK;
}

struct DevInfo {
    int sm_count = 0;
    int is_sm100 = 0;
    bool valid = false;
};

static DevInfo& dev_info() {
    static thread_local DevInfo info;
    static thread_local int cached_dev = -1;
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        info = DevInfo();
        return info;
    }
    if (dev != cached_dev || !info.valid) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, dev) == cudaSuccess) {
            info.sm_count = prop.multiProcessorCount;
            info.is_sm100 = (prop.major == 10 && prop.minor == 0) ? 1 : 0;
            info.valid = true;
            cached_dev = dev;
        } else {
            info = DevInfo();
        }
    }
    return info;
}

// ---------------------------------------------------------------------------------------- profiler
static std::atomic<long long> g_launches{0};
static std::atomic<int> g_prof_on{0};
struct ProfRec {
    const char* name;
    cudaEvent_t e0, e1;
    double flops, bytes;
    int launches;
};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof_recs;
static std::vector<cudaEvent_t> g_event_pool;

static cudaEvent_t get_event() {
    if (!g_event_pool.empty()) {
        cudaEvent_t e = g_event_pool.back();
        g_event_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

Prof::Prof(const char* name, cudaStream_t s, double flops, double bytes, int launches) : slot(-1), stream(s) {
    g_launches.fetch_add(launches, std::memory_order_relaxed);
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfRec r{name, get_event(), get_event(), flops, bytes, launches};
    cudaEventRecord(r.e0, s);
    g_prof_recs.push_back(r);
    slot = (int)g_prof_recs.size() - 1;
}
Prof::~Prof() {
    if (slot < 0) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (slot < (int)g_prof_recs.size()) cudaEventRecord(g_prof_recs[slot].e1, stream);
}

int device_sm_count() { return dev_info().sm_count; }
int device_is_sm100() { return dev_info().is_sm100; }

}  // namespace wg

extern "C" int wg_version(void) { return WG_ABI_VERSION; }
extern "C" const char* wg_last_error(void) { return wg::g_err; }
extern "C" int wg_device_check(int device) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) {
        wg::set_error("wg_device_check: no CUDA device %d (count=%d); this library has no CPU fallback", device, n);
        (void)cudaGetLastError();
        return WG_ERR_UNSUPPORTED;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        wg::set_error("wg_device_check: cudaGetDeviceProperties failed");
        return WG_ERR_CUDA;
    }
    if (!(prop.major == 10 && prop.minor == 0)) {
        wg::set_error("wg_device_check: device %d is sm_%d%d; kernels are built for sm_100a only", device, prop.major, prop.minor);
        return WG_ERR_UNSUPPORTED;
    }
    return WG_OK;
}

extern "C" long long wg_launch_count(int reset) {
    long long v = wg::g_launches.load();
    if (reset) wg::g_launches.store(0);
    return v;
}
extern "C" int wg_profile_enable(int on) {
    wg::g_prof_on.store(on ? 1 : 0);
    return WG_OK;
}
// Synchronises the recorded events and writes one line per kernel family:
//   "<name> <launches> <total_ms> <flops> <bytes>\n"; returns the number of bytes written (or needed if > cap).
extern "C" long long wg_profile_collect(char* out, long long cap) {
    using namespace wg;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    struct Agg { long long n = 0; double ms = 0, flops = 0, bytes = 0; };
    std::map<std::string, Agg> agg;
    for (auto& r : g_prof_recs) {
        float ms = 0.f;
        if (cudaEventSynchronize(r.e1) == cudaSuccess) cudaEventElapsedTime(&ms, r.e0, r.e1);
        Agg& a = agg[r.name];
        a.n += r.launches; a.ms += ms; a.flops += r.flops; a.bytes += r.bytes;
        g_event_pool.push_back(r.e0);
        g_event_pool.push_back(r.e1);
    }
    g_prof_recs.clear();
    std::string text;
    char line[256];
    for (auto& kv : agg) {
        snprintf(line, sizeof(line), "%s %lld %.6f %.6e %.6e\n", kv.first.c_str(), kv.second.n, kv.second.ms, kv.second.flops, kv.second.bytes);
        text += line;
    }
    if (out && cap > 0) {
        long long n = (long long)text.size() < cap - 1 ? (long long)text.size() : cap - 1;
        memcpy(out, text.data(), (size_t)n);
        out[n] = 0;
    }
    return (long long)text.size() + 1;
}
