// Version / diagnostics entry points of the C-ABI.
#include "xb_common.cuh"

extern "C" int xb_version(void) { return 100; }  // 0.1.0 (round 1)

extern "C" const char *xb_error_string(int code) {
    switch (code) {
        case XB_OK: return "ok";
        case XB_EINVAL: return "xb200: invalid argument (null pointer or bad size)";
        case XB_EALIGN: return "xb200: pointer or row size not aligned as required";
        case XB_ERANGE: return "xb200: argument outside the supported range";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "xb200: unknown error";
    }
}

extern "C" int xb_device_info(int *sm_count, int *cc_major, int *cc_minor) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) return (int)e;
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return XB_OK;
}
