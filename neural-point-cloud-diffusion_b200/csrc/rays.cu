// R1 + R2: ray generation and ray/cube limits.
// Replaces RaySampler.forward (renderers/ray_sampler.py:10-63), get_ray_limits_box (renderers/math_utils.py:46-97)
// and Renderer.get_ray_limits (renderers/renderer.py:36-47).  Bit-exact with oracle/pointnerf_oracle.py
// (generate_rays / ray_limits_box / get_ray_limits): same operation order, every op rounded to fp32.
#include "common.cuh"
#include "npcd_b200.h"

namespace npcd {

__device__ __forceinline__ float nanmax(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : fmaxf(a, b); }
__device__ __forceinline__ float nanmin(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : fminf(a, b); }

// One thread per ray.  limits_ord[0] = ordered-min of valid starts, [1] = ordered-max of valid ends.
__global__ void k_rays(const float* __restrict__ extr, const float* __restrict__ intr, int n_views, int res,
                       const long long* __restrict__ subset, int n_sub, float box, float* __restrict__ cam_out,
                       float* __restrict__ origins, float* __restrict__ dirs, float* __restrict__ start,
                       float* __restrict__ end, uint32_t* __restrict__ limits_ord) {
  const int rays_per_view = subset ? n_sub : res * res;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)n_views * rays_per_view;
  uint32_t my_min = 0xffffffffu, my_max = 0u;
  if (gid < total) {
    const int v = (int)(gid / rays_per_view);
    const int rr = (int)(gid % rays_per_view);
    const int pix = subset ? (int)subset[rr] : rr;
    const int i = pix / res, j = pix % res;
    const float* E = extr + (size_t)v * 16;
    const float* Kc = intr + (size_t)v * 9;
    const float fx = Kc[0], sk = Kc[1], cx = Kc[2], fy = Kc[4], cy = Kc[5];
    const float xc = (float)j + 0.5f, yc = (float)i + 0.5f;
    // ray_sampler.py:28-29
    const float t1 = __fsub_rn(xc, cx);
    const float t3 = __fdiv_rn(__fmul_rn(cy, sk), fy);
    const float t4 = __fadd_rn(t1, t3);
    const float t6 = __fdiv_rn(__fmul_rn(sk, yc), fy);
    const float xl = __fdiv_rn(__fsub_rn(t4, t6), fx);
    const float yl = __fdiv_rn(__fsub_rn(yc, cy), fy);
    // ray_sampler.py:35-45   Rt[a][b] = E[b][a]
    float cam[3], d[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float r0 = E[0 * 4 + a], r1 = E[1 * 4 + a], r2 = E[2 * 4 + a];
      cam[a] = -__fadd_rn(__fadd_rn(__fmul_rn(r0, E[3]), __fmul_rn(r1, E[7])), __fmul_rn(r2, E[11]));
      const float p = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r0, xl), __fmul_rn(r1, yl)), r2), cam[a]);
      d[a] = __fsub_rn(p, cam[a]);
    }
    float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
    nrm = fmaxf(nrm, 1e-12f);
#pragma unroll
    for (int a = 0; a < 3; ++a) d[a] = __fdiv_rn(d[a], nrm);
    // math_utils.py:64-95
    float tn[3], tf[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float inv = __fdiv_rn(1.0f, d[a]);
      const bool neg = inv < 0.0f;
      const float nearb = neg ? box : -box, farb = neg ? -box : box;
      tn[a] = __fmul_rn(__fsub_rn(nearb, cam[a]), inv);
      tf[a] = __fmul_rn(__fsub_rn(farb, cam[a]), inv);
    }
    float tmin = tn[0], tmax = tf[0];
    bool valid = !((tmin > tf[1]) || (tn[1] > tmax));
    tmin = nanmax(tmin, tn[1]);
    tmax = nanmin(tmax, tf[1]);
    valid = valid && !((tmin > tf[2]) || (tn[2] > tmax));
    tmin = nanmax(tmin, tn[2]);
    tmax = nanmin(tmax, tf[2]);
    if (!valid) { tmin = -1.0f; tmax = -2.0f; }
    if (tmax > tmin) { my_min = f2ord(tmin); my_max = f2ord(tmax); }
    start[gid] = tmin;
    end[gid] = tmax;
    dirs[gid * 3 + 0] = d[0]; dirs[gid * 3 + 1] = d[1]; dirs[gid * 3 + 2] = d[2];
    if (origins) { origins[gid * 3 + 0] = cam[0]; origins[gid * 3 + 1] = cam[1]; origins[gid * 3 + 2] = cam[2]; }
    if (rr == 0) { cam_out[v * 3 + 0] = cam[0]; cam_out[v * 3 + 1] = cam[1]; cam_out[v * 3 + 2] = cam[2]; }
  }
  // block reduce -> one atomic pair per block
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    my_min = min(my_min, __shfl_xor_sync(0xffffffffu, my_min, o));
    my_max = max(my_max, __shfl_xor_sync(0xffffffffu, my_max, o));
  }
  __shared__ uint32_t smin[32], smax[32];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { smin[w] = my_min; smax[w] = my_max; }
  __syncthreads();
  if (w == 0) {
    const int nw = blockDim.x >> 5;
    my_min = l < nw ? smin[l] : 0xffffffffu;
    my_max = l < nw ? smax[l] : 0u;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      my_min = min(my_min, __shfl_xor_sync(0xffffffffu, my_min, o));
      my_max = max(my_max, __shfl_xor_sync(0xffffffffu, my_max, o));
    }
    if (l == 0) {
      if (my_min != 0xffffffffu) atomicMin(&limits_ord[0], my_min);
      if (my_max != 0u) atomicMax(&limits_ord[1], my_max);
    }
  }
}

// renderer.py:40-43: rays that miss the cube inherit the global [min start, max end] of the rays that hit it.
__global__ void k_fix_invalid(float* __restrict__ start, float* __restrict__ end, long long total,
                              const uint32_t* __restrict__ limits_ord) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const uint32_t lo = limits_ord[0], hi = limits_ord[1];
  if (lo == 0xffffffffu) return;  // no valid ray at all: leave (-1,-2) like the reference
  if (!(end[gid] > start[gid])) {
    start[gid] = ord2f(lo);
    end[gid] = ord2f(hi);
  }
}

__global__ void k_init_ord(uint32_t* p) {
  p[0] = 0xffffffffu;
  p[1] = 0u;
}

}  // namespace npcd

extern "C" int npcd_rays_generate(const float* extr, const float* intr, int n_views, int resolution,
                                  const long long* ray_subset, int n_subset, float cube_scale, float* cam_centers,
                                  float* origins, float* dirs, float* ray_start, float* ray_end, void* limits_scratch,
                                  void* stream) {
  using namespace npcd;
  NPCD_CHECK_ARG(extr && intr && cam_centers && dirs && ray_start && ray_end && limits_scratch, "null pointer");
  NPCD_CHECK_ARG(n_views >= 0 && resolution > 0 && resolution <= 4096, "bad n_views/resolution");
  NPCD_CHECK_ARG(!ray_subset || n_subset > 0, "ray_subset given with n_subset <= 0");
  cudaStream_t st = (cudaStream_t)stream;
  const long long per_view = ray_subset ? n_subset : (long long)resolution * resolution;
  const long long total = per_view * n_views;
  if (total == 0) return 0;
  uint32_t* lim = (uint32_t*)limits_scratch;
  k_init_ord<<<1, 1, 0, st>>>(lim);
  const int bs = 256;
  const unsigned grid = (unsigned)((total + bs - 1) / bs);
  k_rays<<<grid, bs, 0, st>>>(extr, intr, n_views, resolution, ray_subset, n_subset, cube_scale, cam_centers, origins, dirs,
                             ray_start, ray_end, lim);
  k_fix_invalid<<<grid, bs, 0, st>>>(ray_start, ray_end, total, lim);
  return check_launch("npcd_rays_generate");
}
