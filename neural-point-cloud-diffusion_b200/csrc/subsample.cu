// Q3: train-mode valid-ray subsampling.  Replaces Aggregator.subsample_valid_rays (fields/aggregators/aggregator.py:78-119): every
// view keeps n = min(min over views of #rays with >= 1 shading sample, ray_subsamples) of its valid rays, chosen uniformly at random,
// reported in ascending ray order (the order of the reference's boolean ray mask, renderers/renderer.py:262-267).
// The reference does it with a global randperm over all valid rays, a stable argsort by instance and boolean masks (CPU randperm,
// three host syncs); here: one kernel counts the valid rays per view (+ global minimum, read by the host because it sizes the output),
// one kernel per view compacts the valid rays, runs a partial Fisher-Yates shuffle with a counter-based generator (any prefix of a
// Fisher-Yates shuffle is a uniform random subset, like the first n entries of the reference's per-instance shuffle) and rank-sorts the
// n picks.  The random stream differs from torch.randperm's; parity tests inject the reference's permutation on the host path instead.
#include "common.cuh"
#include "npcd_b200.h"

namespace npcd {

__global__ void __launch_bounds__(128) k_count_valid(const int* __restrict__ ray_count, int R, int* __restrict__ n_valid,
                                                     int* __restrict__ min_valid) {
  const int view = blockIdx.x;
  int c = 0;
  for (int r = threadIdx.x; r < R; r += blockDim.x) c += ray_count[(size_t)view * R + r] > 0;
  for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  __shared__ int part[4];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    const int t = part[0] + part[1] + part[2] + part[3];
    n_valid[view] = t;
    atomicMin(min_valid, t);
  }
}

__global__ void k_init_min(int* p) { *p = 0x7fffffff; }

// splitmix64 finaliser on (seed, view, draw): one independent 64-bit word per draw, no state
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

__global__ void __launch_bounds__(128) k_subsample_valid(const int* __restrict__ ray_count, int R, int n_keep, unsigned long long seed,
                                                         unsigned long long view_offset, int* __restrict__ ray_ids) {
  extern __shared__ int list[];  // [R] valid ray ids of this view, in ray order
  __shared__ int warp_base[5];
  __shared__ int n_valid_s;
  const int view = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int* rc = ray_count + (size_t)view * R;
  int base = 0;
  for (int r0 = 0; r0 < R; r0 += blockDim.x) {  // ordered compaction, 128 rays per round
    const int r = r0 + threadIdx.x;
    const bool v = r < R && rc[r] > 0;
    const uint32_t m = __ballot_sync(0xffffffffu, v);
    if (lane == 0) warp_base[warp + 1] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
      warp_base[0] = 0;
      for (int w = 1; w <= 4; ++w) warp_base[w] += warp_base[w - 1];
    }
    __syncthreads();
    if (v) list[base + warp_base[warp] + __popc(m & ((1u << lane) - 1u))] = r;
    base += warp_base[4];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    n_valid_s = base;
    const int n = min(n_keep, base);
    for (int i = 0; i < n; ++i) {  // partial Fisher-Yates: list[0..n) becomes a uniform random n-subset
      const uint64_t u = mix64(seed ^ mix64(((view_offset + (uint64_t)view) << 32) | (uint32_t)i));
      const int j = i + (int)(((u >> 32) * (uint64_t)(base - i)) >> 32);
      const int t = list[i];
      list[i] = list[j];
      list[j] = t;
    }
  }
  __syncthreads();
  const int n = min(n_keep, n_valid_s);
  for (int t = threadIdx.x; t < n; t += blockDim.x) {  // rank sort (n <= ray_subsamples, a few hundred at most)
    const int mine = list[t];
    int rank = 0;
    for (int k = 0; k < n; ++k) rank += list[k] < mine;
    ray_ids[(size_t)view * n_keep + rank] = view * R + mine;
  }
}

}  // namespace npcd

using namespace npcd;

extern "C" int npcd_count_valid_rays(const int* ray_count, long long n_views, int rays_per_view, int* n_valid, int* min_valid,
                                     void* stream) {
  NPCD_CHECK_ARG(n_views >= 0 && rays_per_view > 0, "bad sizes");
  NPCD_CHECK_ARG(min_valid, "null pointer");
  k_init_min<<<1, 1, 0, (cudaStream_t)stream>>>(min_valid);
  if (n_views > 0) {
    NPCD_CHECK_ARG(ray_count && n_valid, "null pointer");
    k_count_valid<<<(unsigned)n_views, 128, 0, (cudaStream_t)stream>>>(ray_count, rays_per_view, n_valid, min_valid);
  }
  return check_launch("npcd_count_valid_rays");
}

extern "C" int npcd_subsample_valid_rays(const int* ray_count, long long n_views, int rays_per_view, int n_keep,
                                         unsigned long long seed, long long view_offset, int* ray_ids, void* stream) {
  NPCD_CHECK_ARG(n_views >= 0 && rays_per_view > 0 && n_keep >= 0 && view_offset >= 0, "bad sizes");
  NPCD_CHECK_ARG((long long)rays_per_view * n_views < (1ll << 31), "ray ids must fit int32");
  if (n_views == 0 || n_keep == 0) return 0;
  NPCD_CHECK_ARG(ray_count && ray_ids, "null pointer");
  const size_t smem = (size_t)rays_per_view * sizeof(int);
  NPCD_CHECK_ARG(smem <= 200 * 1024, "rays_per_view too large for the shared-memory list");
  cudaFuncSetAttribute(k_subsample_valid, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_subsample_valid<<<(unsigned)n_views, 128, smem, (cudaStream_t)stream>>>(ray_count, rays_per_view, n_keep, seed,
                                                                               (unsigned long long)view_offset, ray_ids);
  return check_launch("npcd_subsample_valid_rays");
}
