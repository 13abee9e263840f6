// api.cu -- version and error strings of libcimhead.so.
#include "common.cuh"

#include <atomic>

static std::atomic<unsigned> g_debug_flags{0u};
CIM_API void cim_set_debug_flags(unsigned flags) { g_debug_flags.store(flags, std::memory_order_relaxed); }
CIM_API unsigned cim_get_debug_flags(void) { return g_debug_flags.load(std::memory_order_relaxed); }

CIM_API int cim_abi_version(void) { return CIM_ABI_VERSION; }
CIM_API size_t cim_sizeof_mine_params(void) { return sizeof(cim_mine_params); }

CIM_API const char *cim_error_string(int code) {
    switch (code) {
        case CIM_OK: return "ok";
        case CIM_ERR_ARG: return "cimhead: invalid argument (null pointer, negative size or bad enum)";
        case CIM_ERR_SHAPE: return "cimhead: shape outside the supported range";
        case CIM_ERR_WORKSPACE: return "cimhead: workspace missing, misaligned or too small";
        case CIM_ERR_ALIGN: return "cimhead: pointer not aligned as documented";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "cimhead: unknown error code";
}
