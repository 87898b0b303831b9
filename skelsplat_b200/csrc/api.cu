// Library-level entry points: version, error strings.
#include <cstring>
#include "api_internal.h"

static thread_local char g_last_cuda_error[256] = "";

int ssb_set_cuda_error(cudaError_t e) {
    if (e == cudaSuccess) return SSB_OK;
    std::strncpy(g_last_cuda_error, cudaGetErrorString(e), sizeof(g_last_cuda_error) - 1);
    return SSB_ERR_CUDA;
}

extern "C" {
int ssb_version(void) { return 200; }
#ifndef SSB_SOURCE_HASH
#define SSB_SOURCE_HASH "unknown"
#endif
const char* ssb_source_hash(void) { return SSB_SOURCE_HASH; }
#ifndef SSB_SOURCE_MANIFEST
#define SSB_SOURCE_MANIFEST ""
#endif
const char* ssb_source_manifest(void) { return SSB_SOURCE_MANIFEST; }
// sizeof of the ABI structs (0: ssb_gaussians, 1: ssb_cameras, 2: ssb_opt_config): lets a binding verify its struct layout
int ssb_struct_size(int which) {
    switch (which) {
        case 0: return (int)sizeof(ssb_gaussians);
        case 1: return (int)sizeof(ssb_cameras);
        case 2: return (int)sizeof(ssb_opt_config);
        default: return -1;
    }
}
const char* ssb_last_cuda_error(void) { return g_last_cuda_error; }
const char* ssb_error_string(int code) {
    switch (code) {
        case SSB_OK: return "ok";
        case SSB_ERR_INVALID: return "invalid argument";
        case SSB_ERR_CAPACITY: return "capacity exceeded (P > 1024, r_capacity out of range, or buffer too small)";
        case SSB_ERR_CUDA: return "CUDA runtime error";
        case SSB_ERR_UNSUPPORTED: return "unsupported configuration (channel count not instantiated)";
        default: return "unknown error";
    }
}
}
