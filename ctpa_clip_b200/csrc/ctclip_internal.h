// Internal host-side helpers shared by the translation units of libctclip_sm100.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ctclip_b200.h"

namespace ctclip {

// records a thread-local error message and returns `code`
int fail(int code, const char* fmt, ...);
// CTCLIP_OK iff the current device is compute capability 10.x
int require_sm100();
int sm_count();
// counts a kernel launch (reported by ctclip_launch_count)
void count_launch(int n = 1);
// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
int encode_tmap(CUtensorMap* map, CUtensorMapDataType dtype, int rank, void* base, const cuuint64_t* dims,
                const cuuint64_t* strides_bytes /* rank-1 */, const cuuint32_t* box, const cuuint32_t* elem_strides,
                CUtensorMapSwizzle swizzle);

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(CTCLIP_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
  count_launch();
  return CTCLIP_OK;
}

}  // namespace ctclip
