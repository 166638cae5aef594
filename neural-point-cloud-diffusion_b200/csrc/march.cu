// Q2: ray marching + exact radius-kNN query.  Replaces Aggregator.query_keypoints (fields/aggregators/aggregator.py:25-76,
// exact branch :42-58) and torch_knnquery.VoxelGrid.query (call site aggregator.py:63), plus the sample generation feeding it
// (renderers/renderer.py:49-77, renderers/volume_renderer.py:63-70) -- the [B,T,R,128,3] sample tensor is never materialised.
//
// Two passes, no atomics on the data path, deterministic layout (ray-major, depth order = the reference's masked_select order):
//   k_march_count : warp per ray; lane l tests samples l, l+32, l+64, l+96: dilated-occupancy bit, then "any point within r"
//                   with early exit.  Emits a 128-bit validity mask per ray and count = min(popc, max_shading_pts).
//   (exclusive scan of counts -> ray_offset, cub::DeviceScan, scan.cu)
//   k_knn_fill    : thread per kept sample; finds its ray (binary search) and its depth index (rank-th set bit), recomputes
//                   the sample position bit-exactly and selects the <=8 nearest points within r in canonical (dist, index)
//                   order with a register-resident sorted list.
#include "common.cuh"
#include "npcd_b200.h"

namespace npcd {

struct GridView {
  const int* cell_start;      // [G^3+1]
  const float4* sorted_pts;   // [P]
  const uint32_t* occ_bits;   // [G^3/32]
};

// Visits every point stored in the 27 cells around (x,y,z); f(px,py,pz,orig_idx) returns true to stop early.
template <typename F>
__device__ __forceinline__ void visit_neighbourhood(const GridView& g, float x, float y, float z, F&& f) {
  const int cx = grid_coord(x), cy = grid_coord(y), cz = grid_coord(z);
  const int x0 = max(cx - 1, 0), x1 = min(cx + 1, kGrid - 1);
#pragma unroll 1
  for (int dz = -1; dz <= 1; ++dz) {
    const int qz = cz + dz;
    if (qz < 0 || qz >= kGrid) continue;
#pragma unroll 1
    for (int dy = -1; dy <= 1; ++dy) {
      const int qy = cy + dy;
      if (qy < 0 || qy >= kGrid) continue;
      const int base = (qz * kGrid + qy) * kGrid;
      const int s = __ldg(g.cell_start + base + x0), e = __ldg(g.cell_start + base + x1 + 1);
      for (int i = s; i < e; ++i) {
        const float4 p = __ldg(g.sorted_pts + i);
        if (f(p.x, p.y, p.z, __float_as_int(p.w))) return;
      }
    }
  }
}

__global__ void __launch_bounds__(256) k_march_count(const float* __restrict__ cam, const float* __restrict__ dirs,
                                                     const float* __restrict__ start, const float* __restrict__ end,
                                                     const float* __restrict__ jitter, long long n_rays, int rays_per_view,
                                                     int views_per_obj, int P, const int* __restrict__ cell_start,
                                                     const float4* __restrict__ sorted_pts,
                                                     const uint32_t* __restrict__ occ_bits, const float* __restrict__ aabb,
                                                     float radius, int max_shading, uint32_t* __restrict__ valid_bits,
                                                     int* __restrict__ ray_count) {
  const int lane = threadIdx.x & 31;
  const long long ray = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ray >= n_rays) return;
  const int view = (int)(ray / rays_per_view);
  const int obj = view / views_per_obj;
  GridView g{cell_start + (size_t)obj * (kGridCells + 1), sorted_pts + (size_t)obj * P, occ_bits + (size_t)obj * kGridWords};
  const float ox = cam[view * 3], oy = cam[view * 3 + 1], oz = cam[view * 3 + 2];
  const float dx = dirs[ray * 3], dy = dirs[ray * 3 + 1], dz = dirs[ray * 3 + 2];
  const float t0 = start[ray], t1 = end[ray];
  const float* jit = jitter ? jitter + ray * kDepthRes : nullptr;
  // Empty-space skipping: only the depth samples inside the object's dilated-occupancy box can have a neighbour.  The sample-index
  // range is conservative (+-2 samples absorb the rounding of the slab test and the train-mode jitter of < 1 step); every sample
  // inside it still takes the exact test below, so the result is bit-identical to testing all 128.
  int i_lo = 0, i_hi = kDepthRes - 1;
  if (aabb) {
    const float* bx = aabb + (size_t)obj * 6;
    float tmin = -INFINITY, tmax = INFINITY;
    bool miss = false;
    const float o3[3] = {ox, oy, oz}, d3[3] = {dx, dy, dz};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float lo = __ldg(bx + a), hi = __ldg(bx + 3 + a);
      if (fabsf(d3[a]) < 1e-12f) {
        miss |= (o3[a] < lo || o3[a] > hi);
      } else {
        const float inv = 1.0f / d3[a];
        const float ta = (lo - o3[a]) * inv, tb = (hi - o3[a]) * inv;  // +-inf bounds give +-inf (o is finite)
        tmin = fmaxf(tmin, fminf(ta, tb));
        tmax = fminf(tmax, fmaxf(ta, tb));
      }
    }
    const float span = t1 - t0;
    if (miss || tmin > tmax || !(span > 0.f)) {
      if (miss || tmin > tmax) i_hi = -1;  // the ray never enters the box (a degenerate span keeps the full range)
    } else {
      const float s = (float)(kDepthRes - 1) / span;
      const float flo = (tmin - t0) * s - 2.0f, fhi = (tmax - t0) * s + 2.0f;
      i_lo = flo <= 0.f ? 0 : (flo >= (float)kDepthRes ? kDepthRes : (int)flo);
      i_hi = fhi >= (float)(kDepthRes - 1) ? kDepthRes - 1 : (fhi < 0.f ? -1 : (int)fhi + 1);
      i_hi = min(i_hi, kDepthRes - 1);
    }
  }
  int total = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (j * 32 > i_hi || j * 32 + 31 < i_lo) {  // warp-uniform: no sample of this word can be valid
      if (lane == j) valid_bits[ray * 4 + j] = 0u;
      continue;
    }
    const int i = j * 32 + lane;
    const float t = sample_depth(t0, t1, i, jit);
    const float x = axpy_rn(ox, t, dx), y = axpy_rn(oy, t, dy), z = axpy_rn(oz, t, dz);
    bool hit = false;
    // samples outside the cube cannot be within r of a cell we index (points are clamped into border cells, so test anyway)
    const int c = (grid_coord(z) * kGrid + grid_coord(y)) * kGrid + grid_coord(x);
    if (i >= i_lo && i <= i_hi && ((__ldg(g.occ_bits + (c >> 5)) >> (c & 31)) & 1u)) {
      visit_neighbourhood(g, x, y, z, [&](float px, float py, float pz, int) {
        hit = dist_rn(x, y, z, px, py, pz) < radius;
        return hit;
      });
    }
    const uint32_t word = __ballot_sync(0xffffffffu, hit);
    total += __popc(word);
    if (lane == j) valid_bits[ray * 4 + j] = word;
  }
  if (lane == 0) ray_count[ray] = min(total, max_shading);
}

// <= 8 nearest points within `radius`, ascending (dist, index), written as two int4 (global index = base + p, -1 padded).
__device__ __forceinline__ void select_and_store(const GridView& g, float x, float y, float z, float radius, int base,
                                                 int* __restrict__ out) {
  unsigned long long best[kK];
#pragma unroll
  for (int j = 0; j < kK; ++j) best[j] = ~0ull;
  visit_neighbourhood(g, x, y, z, [&](float px, float py, float pz, int idx) {
    const float d = dist_rn(x, y, z, px, py, pz);
    if (d < radius) {
      unsigned long long cur = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)idx;
      if (cur < best[kK - 1]) {
#pragma unroll
        for (int j = 0; j < kK; ++j) {
          const unsigned long long b = best[j];
          const bool sw = cur < b;
          best[j] = sw ? cur : b;
          cur = sw ? b : cur;
        }
      }
    }
    return false;
  });
  int tmp[kK];
#pragma unroll
  for (int j = 0; j < kK; ++j) tmp[j] = best[j] == ~0ull ? -1 : base + (int)(best[j] & 0xffffffffu);
  reinterpret_cast<int4*>(out)[0] = make_int4(tmp[0], tmp[1], tmp[2], tmp[3]);
  reinterpret_cast<int4*>(out)[1] = make_int4(tmp[4], tmp[5], tmp[6], tmp[7]);
}

__global__ void __launch_bounds__(128) k_knn_points(const float* __restrict__ x, const int* __restrict__ query_obj, long long n,
                                                    int queries_per_obj, int P, const int* __restrict__ cell_start,
                                                    const float4* __restrict__ sorted_pts, float radius,
                                                    int* __restrict__ nbr_idx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int obj = query_obj ? query_obj[i] : (int)(i / queries_per_obj);
  GridView g{cell_start + (size_t)obj * (kGridCells + 1), sorted_pts + (size_t)obj * P, nullptr};
  select_and_store(g, x[i * 3], x[i * 3 + 1], x[i * 3 + 2], radius, obj * P, nbr_idx + i * kK);
}

__device__ __forceinline__ int nth_set_bit(uint32_t w, int n) {  // position of the n-th (0-based) set bit
  for (int k = 0; k < n; ++k) w &= w - 1;
  return __ffs(w) - 1;
}

__global__ void __launch_bounds__(128) k_knn_fill(const float* __restrict__ cam, const float* __restrict__ dirs,
                                                  const float* __restrict__ start, const float* __restrict__ end,
                                                  const float* __restrict__ jitter, const int* __restrict__ ray_ids,
                                                  long long n_sel, const long long* __restrict__ ray_offset,
                                                  const uint32_t* __restrict__ valid_bits, int rays_per_view, int views_per_obj,
                                                  int P, const int* __restrict__ cell_start,
                                                  const float4* __restrict__ sorted_pts, float radius,
                                                  long long capacity, int* __restrict__ nbr_idx,
                                                  float4* __restrict__ sample_pos, int* __restrict__ sample_ray) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long S = min(ray_offset[n_sel], capacity);
  if (s >= S) return;
  // upper_bound(ray_offset, s) - 1
  long long lo = 0, hi = n_sel;
  while (hi - lo > 1) {
    const long long mid = (lo + hi) >> 1;
    if (__ldg(ray_offset + mid) <= s) lo = mid; else hi = mid;
  }
  const long long sel = lo;
  const long long ray = ray_ids ? (long long)ray_ids[sel] : sel;
  int rank = (int)(s - ray_offset[sel]);
  int i = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t w = __ldg(valid_bits + ray * 4 + j);
    const int pc = __popc(w);
    if (rank >= 0 && rank < pc) { i = j * 32 + nth_set_bit(w, rank); rank = -1; }
    else if (rank >= 0) rank -= pc;
  }
  const int view = (int)(ray / rays_per_view);
  const int obj = view / views_per_obj;
  GridView g{cell_start + (size_t)obj * (kGridCells + 1), sorted_pts + (size_t)obj * P, nullptr};
  const float ox = cam[view * 3], oy = cam[view * 3 + 1], oz = cam[view * 3 + 2];
  const float dx = dirs[ray * 3], dy = dirs[ray * 3 + 1], dz = dirs[ray * 3 + 2];
  const float t = sample_depth(start[ray], end[ray], i, jitter ? jitter + ray * kDepthRes : nullptr);
  const float x = axpy_rn(ox, t, dx), y = axpy_rn(oy, t, dy), z = axpy_rn(oz, t, dz);

  select_and_store(g, x, y, z, radius, obj * P, nbr_idx + s * kK);
  // slot depth: nanmean over axes of (x - o) / d   (renderers/renderer.py:105)
  const float q0 = __fdiv_rn(__fsub_rn(x, ox), dx), q1 = __fdiv_rn(__fsub_rn(y, oy), dy), q2 = __fdiv_rn(__fsub_rn(z, oz), dz);
  float sum = 0.f, cnt = 0.f;
  if (q0 == q0) { sum = __fadd_rn(sum, q0); cnt += 1.f; }
  if (q1 == q1) { sum = __fadd_rn(sum, q1); cnt += 1.f; }
  if (q2 == q2) { sum = __fadd_rn(sum, q2); cnt += 1.f; }
  sample_pos[s] = make_float4(x, y, z, __fdiv_rn(sum, cnt));
  if (sample_ray) sample_ray[s] = (int)sel;
}

}  // namespace npcd

extern "C" int npcd_march_count(const float* cam_centers, const float* dirs, const float* ray_start, const float* ray_end,
                                const float* jitter, long long n_rays, int rays_per_view, int views_per_obj, int n_points,
                                const int* cell_start, const float* sorted_pts, const unsigned* occ_bits, const float* aabb,
                                float radius, int max_shading_pts, unsigned* valid_bits, int* ray_count, void* stream) {
  using namespace npcd;
  NPCD_CHECK_ARG(cam_centers && dirs && ray_start && ray_end && cell_start && sorted_pts && occ_bits && valid_bits && ray_count,
                 "null pointer");
  NPCD_CHECK_ARG(n_rays >= 0 && rays_per_view > 0 && views_per_obj > 0 && n_points > 0, "bad sizes");
  NPCD_CHECK_ARG(max_shading_pts > 0 && max_shading_pts <= kDepthRes, "max_shading_pts must be in [1,128]");
  NPCD_CHECK_ARG(radius > 0.f && radius <= 2.0f / kGrid, "radius must be in (0, 1/12] (grid cell edge)");
  if (n_rays == 0) return 0;
  const int wpb = 8;
  const unsigned grid = (unsigned)((n_rays + wpb - 1) / wpb);
  k_march_count<<<grid, wpb * 32, 0, (cudaStream_t)stream>>>(cam_centers, dirs, ray_start, ray_end, jitter, n_rays, rays_per_view,
                                                             views_per_obj, n_points, cell_start, (const float4*)sorted_pts,
                                                             occ_bits, aabb, radius, max_shading_pts, valid_bits, ray_count);
  return check_launch("npcd_march_count");
}

extern "C" int npcd_knn_fill(const float* cam_centers, const float* dirs, const float* ray_start, const float* ray_end,
                             const float* jitter, const int* ray_ids, long long n_sel, const long long* ray_offset,
                             const unsigned* valid_bits, int rays_per_view, int views_per_obj, int n_points,
                             const int* cell_start, const float* sorted_pts, float radius, long long capacity, int* nbr_idx,
                             float* sample_pos, int* sample_ray, void* stream) {
  using namespace npcd;
  NPCD_CHECK_ARG(cam_centers && dirs && ray_start && ray_end && ray_offset && valid_bits && cell_start && sorted_pts, "null pointer");
  NPCD_CHECK_ARG(capacity == 0 || (nbr_idx && sample_pos), "null output with capacity > 0");
  NPCD_CHECK_ARG(n_sel >= 0 && capacity >= 0, "bad sizes");
  if (n_sel == 0 || capacity == 0) return 0;
  const int bs = 128;
  const unsigned grid = (unsigned)((capacity + bs - 1) / bs);
  k_knn_fill<<<grid, bs, 0, (cudaStream_t)stream>>>(cam_centers, dirs, ray_start, ray_end, jitter, ray_ids, n_sel, ray_offset,
                                                    valid_bits, rays_per_view, views_per_obj, n_points, cell_start,
                                                    (const float4*)sorted_pts, radius, capacity, nbr_idx, (float4*)sample_pos,
                                                    sample_ray);
  return check_launch("npcd_knn_fill");
}

extern "C" int npcd_knn_points(const float* x, const int* query_obj, long long n, int queries_per_obj, int n_points,
                               const int* cell_start, const float* sorted_pts, float radius, int* nbr_idx, void* stream) {
  using namespace npcd;
  NPCD_CHECK_ARG(n >= 0 && n_points > 0 && (query_obj || queries_per_obj > 0), "bad sizes");
  NPCD_CHECK_ARG(radius > 0.f && radius <= 2.0f / kGrid, "radius must be in (0, 1/12] (grid cell edge)");
  if (n == 0) return 0;
  NPCD_CHECK_ARG(x && cell_start && sorted_pts && nbr_idx, "null pointer");
  k_knn_points<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(x, query_obj, n, queries_per_obj, n_points, cell_start,
                                                                             (const float4*)sorted_pts, radius, nbr_idx);
  return check_launch("npcd_knn_points");
}
