// Host-side helpers shared by the C-ABI translation units: error reporting, launch accounting,
// tensor-map encoding through the driver entry point (the library does not link libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/qv2x.h"

namespace qv2x {

std::string& last_error_ref();
int set_error(int code, const char* fmt, ...);
extern std::atomic<long long> g_launch_count;
extern int g_debug_flags;
extern long long* g_trace_ptr;   // device buffer for igemm role timelines (qv2x_debug_trace), or nullptr
int num_sms();

#define QV2X_CUDA_OK(expr)                                                                            \
    do {                                                                                              \
        cudaError_t e__ = (expr);                                                                     \
        if (e__ != cudaSuccess)                                                                       \
            return ::qv2x::set_error(QV2X_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                                     __FILE__, __LINE__);                                             \
    } while (0)

// Descriptors carry their own size as first field (include/qv2x.h): reject callers built against another layout.
#define QV2X_CHECK_SIZE(ptr, type)                                                                              \
    QV2X_REQUIRE((ptr)->struct_size == sizeof(type), #type ".struct_size is %u, this library expects %zu "     \
                 "(set it to sizeof(" #type "); the caller was built against another revision of qv2x.h)",       \
                 (ptr)->struct_size, sizeof(type))

#define QV2X_REQUIRE(cond, ...)                                                  \
    do {                                                                         \
        if (!(cond)) return ::qv2x::set_error(QV2X_ERR_INVALID, __VA_ARGS__);    \
    } while (0)

// uint8 tensor map, `rank` dims (innermost first), byte strides for dims 1..rank-1.
int encode_tmap_u8(CUtensorMap* out, const void* gaddr, int rank, const uint64_t* dims, const uint64_t* strides,
                   const uint32_t* box, const uint32_t* elem_strides, int swizzle_bytes);

template <class T>
int upload(T** dptr, const T* host, size_t n) {
    QV2X_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(dptr), n * sizeof(T)));
    QV2X_CUDA_OK(cudaMemcpy(*dptr, host, n * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

}  // namespace qv2x
