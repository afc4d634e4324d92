// Library-level entry points of the C ABI (version, device info, counters).
#include "common.cuh"

using namespace pmb;

extern "C" int pmb_version(void) { return 100; }

extern "C" int pmb_device_info(int *sm_count, int *cc, size_t *global_mem) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) return (int)e;
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc) *cc = prop.major * 10 + prop.minor;
    if (global_mem) *global_mem = prop.totalGlobalMem;
    return 0;
}

extern "C" long long pmb_launch_count(void) { return g_launch_count; }
extern "C" void pmb_launch_count_reset(void) { g_launch_count = 0; }

extern "C" const char *pmb_error_string(int code) {
    switch (code) {
        case 0: return "success";
        case PMB_E_BADARG: return "pymes_b200: bad argument";
        case PMB_E_WORKSPACE: return "pymes_b200: workspace missing or too small";
        case PMB_E_UNSUPPORTED: return "pymes_b200: unsupported configuration";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "pymes_b200: unknown error";
}
