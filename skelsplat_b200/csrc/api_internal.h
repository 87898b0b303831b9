// Internal glue shared by the translation units of libskelsplat_b200.so.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/skelsplat_b200.h"

namespace ssb {
struct StateLayout {
    size_t off[SSB_F_COUNT];
    size_t total;
};
}  // namespace ssb

// Records the CUDA error string (if any) and maps it to a status code.
int ssb_set_cuda_error(cudaError_t e);
