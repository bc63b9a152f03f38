// ABI bookkeeping: version, error text, launch counter.
#include "ofr_common.cuh"

namespace ofr {
std::atomic<uint64_t> g_launches{0};
}

extern "C" int ofr_abi_version(void) { return OFR_ABI_VERSION; }

extern "C" uint64_t ofr_launch_count(void) { return ofr::g_launches.load(std::memory_order_relaxed); }

extern "C" const char* ofr_error_string(int code) {
    switch (code) {
        case OFR_OK: return "ok";
        case OFR_EINVAL: return "invalid argument (null pointer, non-positive dimension or bad mode)";
        case OFR_ENOTSUP: return "parameter combination not supported";
        case OFR_EWORKSPACE: return "workspace missing, too small or not 256-byte aligned";
        default: break;
    }
    if (code >= OFR_ECUDA) return cudaGetErrorString((cudaError_t)(code - OFR_ECUDA));
    return "unknown ofr error code";
}
