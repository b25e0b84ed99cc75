#include "host_common.h"

#include <cudaTypedefs.h>

#include <mutex>

namespace qv2x {

std::string& last_error_ref() {
    static thread_local std::string s;
    return s;
}

int set_error(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error_ref() = buf;
    return code;
}

std::atomic<long long> g_launch_count{0};
int g_debug_flags = 0;
long long* g_trace_ptr = nullptr;

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
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

int encode_tmap_u8(CUtensorMap* out, const void* gaddr, int rank, const uint64_t* dims, const uint64_t* strides,
                   const uint32_t* box, const uint32_t* elem_strides, int swizzle_bytes) {
    auto fn = get_encode_fn();
    if (fn == nullptr) return set_error(QV2X_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                            : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                            : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                  : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, static_cast<cuuint32_t>(rank), const_cast<void*>(gaddr),
                    reinterpret_cast<const cuuint64_t*>(dims), reinterpret_cast<const cuuint64_t*>(strides),
                    reinterpret_cast<const cuuint32_t*>(box), reinterpret_cast<const cuuint32_t*>(elem_strides),
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        return set_error(QV2X_ERR_CUDA,
                         "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu] box [%u %u %u %u]",
                         static_cast<int>(r), rank, (unsigned long long)dims[0],
                         (unsigned long long)(rank > 1 ? dims[1] : 0), (unsigned long long)(rank > 2 ? dims[2] : 0),
                         (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], rank > 1 ? box[1] : 0,
                         rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    }
    return 0;
}

}  // namespace qv2x

extern "C" {

const char* qv2x_last_error(void) { return qv2x::last_error_ref().c_str(); }
int qv2x_version(void) { return 100; }
long long qv2x_launch_count(void) { return qv2x::g_launch_count.load(); }
void qv2x_set_debug_flags(int flags) { qv2x::g_debug_flags = flags; }
void qv2x_debug_trace(long long* d_buf) { qv2x::g_trace_ptr = d_buf; }

int qv2x_device_check(int device) {
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return qv2x::set_error(QV2X_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return qv2x::set_error(QV2X_ERR_DEVICE, "device %d is sm_%d%d; libqv2x needs sm_100 (tcgen05/TMEM)", device,
                               prop.major, prop.minor);
    return 0;
}

}  // extern "C"
