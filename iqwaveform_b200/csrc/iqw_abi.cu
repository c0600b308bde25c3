// iqw_abi.cu -- version / error plumbing of the C-ABI (include/iqw_b200.h).
#include <mutex>
#include "iqw_common.cuh"

namespace iqw {

char* last_error_buffer() {
    static thread_local char buf[512] = "";
    return buf;
}

int device_sm_count(int* sms) {
    static std::mutex m;
    static int cache[64] = {0};
    int dev = 0;
    IQW_CUDA_OK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(m);
    if (dev < 64 && cache[dev]) { *sms = cache[dev]; return IQW_OK; }
    int n = 0;
    IQW_CUDA_OK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    if (dev < 64) cache[dev] = n;
    *sms = n;
    return IQW_OK;
}

}  // namespace iqw

extern "C" int iqw_abi_version(void) { return IQW_ABI_VERSION; }
extern "C" const char* iqw_last_error(void) { return iqw::last_error_buffer(); }
