// misc.cu -- version / status / device queries of the C ABI (include/geom_b200.h).
#include <atomic>
#include <cstring>

#include "common.cuh"

namespace geom {
namespace debug {
static std::atomic<unsigned> g_knob[GEOM_DEBUG_KNOBS];  // float bits; zero-initialised = not set
float knob(int id)
{
    const unsigned u = g_knob[id].load(std::memory_order_relaxed);
    float f;
    memcpy(&f, &u, 4);
    return f;
}
}  // namespace debug
}  // namespace geom

extern "C" int geom_debug_set(int knob, float value)
{
    if (knob < 0 || knob >= GEOM_DEBUG_KNOBS) return GEOM_ERR_INVALID_ARG;
    unsigned u;
    memcpy(&u, &value, 4);
    geom::debug::g_knob[knob].store(u, std::memory_order_relaxed);
    return GEOM_OK;
}

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
