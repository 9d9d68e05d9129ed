// tng_internal.h — host-side helpers shared by the translation units of libtango_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/tango_b200.h"

namespace tng {
int set_error(int code, const char* fmt, ...);
int num_sms();
void count_launch();
// Encode a bf16 tiled tensor map with SWIZZLE_128B and zero OOB fill (driver entry point resolved at run time so
// the library loads on machines without libcuda).
int encode_tmap_bf16(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box);
// col_stats[(n * C + c) * 2 + {0, 1}] += sum / sum of squares of x[n, :, c] over the HW pixels of image n (x: [NB*HW, ld],
// fp32 or bf16). The per-channel form of the GroupNorm statistics: any grouping (also across a channel concat) is a sum
// of channels. Used when a GEMM cannot emit the statistics from its epilogue (partial tiles, split-K).
int launch_col_stats(const void* x, int dt, long long C, long long ld, long long NB, long long HW, double* col_stats,
                     cudaStream_t st);
inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(TNG_ECUDA, "%s launch: %s", what, cudaGetErrorString(e));
  return TNG_OK;
}
}  // namespace tng
