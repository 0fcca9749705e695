// Library-wide state of libspq_b200: error string, launch counter, device queries.
#include <stdarg.h>
#include <string.h>

#include "spq_common.cuh"

namespace spq {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int cached = -1;
    if (cached < 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
        cached = n;
    }
    return cached;
}

}  // namespace spq

extern "C" int spq_abi_version(void) { return SPQ_ABI_VERSION; }
extern "C" const char* spq_last_error(void) { return spq::g_err; }
extern "C" int64_t spq_launch_count(void) { return spq::g_launches.load(); }

extern "C" int spq_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    SPQ_CUDA_OK(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    SPQ_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return SPQ_OK;
}
