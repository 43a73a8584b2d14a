// Library-level entry points of include/dsmnet_b200.h (version, error strings).
#include "common.cuh"

extern "C" int dsm_abi_version(void) { return DSM_ABI_VERSION; }

extern "C" const char* dsm_strerror(int code) {
    switch (code) {
        case 0: return "ok";
        case DSM_EINVAL: return "DSM_EINVAL: bad shape, null pointer or bad enum";
        case DSM_EUNSUPPORTED: return "DSM_EUNSUPPORTED: shape or mode outside what the sm_100a kernels are built for";
        case DSM_EALIGN: return "DSM_EALIGN: pointer not 16-byte aligned";
        case DSM_EDRIVER: return "DSM_EDRIVER: cuTensorMapEncodeTiled unavailable or rejected the tensor map";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "unknown dsm error";
}
