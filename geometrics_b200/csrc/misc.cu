// misc.cu -- version / status / device queries of the C ABI (include/geom_b200.h).
#include "common.cuh"

extern "C" int geom_abi_version(void) { return GEOM_ABI_VERSION; }

extern "C" const char *geom_status_string(int status)
{
    switch (status) {
        case GEOM_OK: return "ok";
        case GEOM_ERR_INVALID_ARG: return "invalid argument (null pointer, negative or unsupported size)";
        case GEOM_ERR_WORKSPACE: return "workspace missing, misaligned or too small";
        case GEOM_ERR_UNSUPPORTED_DEVICE: return "device is not sm_100 (B200)";
        default: break;
    }
    if (status > 0) return cudaGetErrorString((cudaError_t)status);
    return "unknown geom status";
}

extern "C" int geom_device_info(int *sm_count, int *cc_major, int *cc_minor)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    int v = 0;
    if (sm_count) {
        if ((e = cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return (int)e;
        *sm_count = v;
    }
    if (cc_major) {
        if ((e = cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev)) != cudaSuccess) return (int)e;
        *cc_major = v;
    }
    if (cc_minor) {
        if ((e = cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev)) != cudaSuccess) return (int)e;
        *cc_minor = v;
    }
    return GEOM_OK;
}
