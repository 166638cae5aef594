// Error reporting, the count -> offset scan, and small utilities of libnpcd_b200.
#include <cub/cub.cuh>
#include <stdarg.h>
#include <string.h>

#include "common.cuh"
#include "npcd_b200.h"

namespace npcd {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: CUDA error: %s", what, cudaGetErrorString(e));
    return 2;
  }
  return 0;
}

struct CountOf {
  const int* counts;
  const int* ids;
  __host__ __device__ long long operator()(long long i) const { return (long long)counts[ids ? ids[i] : i]; }
};

__global__ void k_zero_first(long long* p) { p[0] = 0; }

}  // namespace npcd

extern "C" const char* npcd_last_error(void) { return npcd::g_err; }

extern "C" int npcd_abi_version(void) { return NPCD_B200_ABI_VERSION; }

extern "C" int npcd_scan_workspace_bytes(long long n, size_t* bytes) {
  using namespace npcd;
  NPCD_CHECK_ARG(bytes && n >= 0, "bad arguments");
  size_t tmp = 0;
  CountOf op{nullptr, nullptr};
  cub::CountingInputIterator<long long> cnt(0);
  cub::TransformInputIterator<long long, CountOf, cub::CountingInputIterator<long long>> it(cnt, op);
  cub::DeviceScan::InclusiveSum(nullptr, tmp, it, (long long*)nullptr, n > 0 ? n : 1);
  *bytes = tmp + 256;
  return 0;
}

// ray_offset[0] = 0, ray_offset[i+1] = sum_{j<=i} ray_count[ray_ids ? ray_ids[j] : j]
extern "C" int npcd_scan_counts(const int* ray_count, const int* ray_ids, long long n, long long* ray_offset, void* workspace,
                                size_t workspace_bytes, void* stream) {
  using namespace npcd;
  NPCD_CHECK_ARG(ray_offset && (n == 0 || (ray_count && workspace)), "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  k_zero_first<<<1, 1, 0, st>>>(ray_offset);
  if (n > 0) {
    CountOf op{ray_count, ray_ids};
    cub::CountingInputIterator<long long> cnt(0);
    cub::TransformInputIterator<long long, CountOf, cub::CountingInputIterator<long long>> it(cnt, op);
    size_t tmp = workspace_bytes;
    cudaError_t e = cub::DeviceScan::InclusiveSum(workspace, tmp, it, ray_offset + 1, n, st);
    if (e != cudaSuccess) {
      set_error("npcd_scan_counts: %s", cudaGetErrorString(e));
      return 2;
    }
  }
  return check_launch("npcd_scan_counts");
}
