// host_util.h -- host-side helpers shared by the C-ABI entry points.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace glass {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled resolved through the runtime (no link-time libcuda dependency).
PFN_encodeTiled get_encode_tiled();
int num_sms();
void count_launch(int n = 1);

}  // namespace glass
