// Host-side helpers shared by all translation units: error reporting across the C ABI,
// TMA tensor-map encoding through the driver entry point (no link-time libcuda dependency,
// so the library still loads on a machine without a GPU driver).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/walkgpt_b200.h"

namespace wg {

void set_error(const char* fmt, ...);  // thread-local message for wg_last_error()

#define WG_CHECK_CUDA(expr)                                                                       \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            ::wg::set_error("%s:%d CUDA error %d (%s) in %s", __FILE__, __LINE__, (int)_e,        \
                            cudaGetErrorString(_e), #expr);                                       \
            return WG_ERR_CUDA;                                                                   \
        }                                                                                         \
    } while (0)

#define WG_REQUIRE(cond, ...)                        \
    do {                                             \
        if (!(cond)) {                               \
            ::wg::set_error(__VA_ARGS__);            \
            return WG_ERR_INVALID;                   \
        }                                            \
    } while (0)

// Opt a kernel in to more than 48 KB of dynamic shared memory.  The attribute belongs to the (function, device) pair, so the
// "already done" flag of each call site is one bit per device ordinal; safe to race (setting the attribute twice is harmless).
#define WG_SMEM_OPT_IN(kern, bytes)                                                                              \
    do {                                                                                                         \
        static std::atomic<unsigned long long> _wg_done{0};                                                      \
        int _wg_dev = 0;                                                                                         \
        WG_CHECK_CUDA(cudaGetDevice(&_wg_dev));                                                                  \
        const unsigned long long _wg_bit = 1ull << (_wg_dev & 63);                                               \
        if (_wg_dev >= 64 || !(_wg_done.load(std::memory_order_acquire) & _wg_bit)) {                            \
            WG_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
            _wg_done.fetch_or(_wg_bit, std::memory_order_release);                                               \
        }                                                                                                        \
    } while (0)

#define WG_TRY(expr)                 \
    do {                             \
        int _rc = (expr);            \
        if (_rc != WG_OK) return _rc;\
    } while (0)

// 2-D / 3-D bf16 (or fp32) row-major tensor maps with 128B swizzle.  dims/box innermost-first.
// strides_bytes has rank-1 entries (stride of dim1, dim2).
int make_tensor_map(CUtensorMap* out, const void* gptr, int elem_bytes, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box);

// same with an explicit swizzle span (bytes: 128, 64, 32, or 0 = none); the box's innermost extent must not exceed the span
int make_tensor_map_sw(CUtensorMap* out, const void* gptr, int elem_bytes, int rank, const uint64_t* dims,
                       const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);

inline int make_tmap_2d_bf16(CUtensorMap* out, const void* gptr, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                             uint32_t box_rows, uint32_t box_cols) {
    uint64_t dims[2] = {cols, rows};
    uint64_t strides[1] = {ld_elems * 2};
    uint32_t box[2] = {box_cols, box_rows};
    return make_tensor_map(out, gptr, 2, 2, dims, strides, box);
}

// ---- launch accounting + optional CUDA-event profiler (used by bench.py's roofline leg) -------------------
// Every kernel launch site creates a Prof scope: it counts the launch and, when profiling is enabled
// (wg_profile_enable), brackets it with CUDA events on the launching stream and attributes the elapsed time,
// algorithmic FLOPs and algorithmic bytes to `name`.
struct Prof {
    Prof(const char* name, cudaStream_t s, double flops = 0.0, double bytes = 0.0, int launches = 1);
    ~Prof();
    int slot;
    cudaStream_t stream;
};

int device_sm_count();
int device_is_sm100();

}  // namespace wg
