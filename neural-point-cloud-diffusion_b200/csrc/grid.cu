// Acceleration grid build ("set_pointset").  Replaces torch_knnquery.VoxelGrid.set_pointset
// (call sites npcd/models/pointnerf/pointnerf.py:67-75,116-124).
//
// Exact mode (the graded one): the grid is a pure acceleration structure for the exact radius query
// (fields/aggregators/aggregator.py:42-58 semantics): 24^3 cells of edge 1/12 (> r = 0.08, so every point within r
// of a sample lies in the 27 cells around the sample's cell, with ~4% slack that absorbs fp rounding), NO per-cell cap.
// Per object we emit
//   cell_start [G^3 + 1] int32   CSR offsets into the sorted point list
//   sorted_pts [P] float4        (x, y, z, original index as int bits), sorted by (cell, original index)
//   occ_bits   [G^3 / 32] u32    bit c set iff any of the 27 cells around c holds a point ("dilated occupancy")
//   aabb       [6] float         (optional) world-space bounding box of the cells with a set bit: (lo xyz, hi xyz), open-ended
//                                (+-inf) on a side that touches the cube border (samples outside the cube are clamped into border
//                                cells).  A sample outside this box cannot have a neighbour: the marcher skips it untested.
// One CTA per object; the whole build lives in shared memory (histogram 55 KB + bits 1.7 KB).
#include "common.cuh"
#include "npcd_b200.h"

namespace npcd {

__device__ __forceinline__ int cell_of(float x, float y, float z) {
  return (grid_coord(z) * kGrid + grid_coord(y)) * kGrid + grid_coord(x);
}

__global__ void __launch_bounds__(512) k_grid_build(const float* __restrict__ kp_pos, int P, int* __restrict__ cell_start,
                                                    float4* __restrict__ sorted_pts, uint32_t* __restrict__ occ_bits,
                                                    float* __restrict__ aabb) {
  extern __shared__ int smem[];
  int* hist = smem;                                  // [kGridCells + 1]
  uint32_t* bits = (uint32_t*)(smem + kGridCells + 1);  // [kGridWords]
  __shared__ int chunk_sum[512];
  __shared__ int box[6];
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const float* pts = kp_pos + (size_t)b * P * 3;
  for (int c = tid; c <= kGridCells; c += nt) hist[c] = 0;
  for (int c = tid; c < kGridWords; c += nt) bits[c] = 0u;
  if (tid < 3) { box[tid] = kGrid; box[3 + tid] = -1; }
  __syncthreads();
  for (int p = tid; p < P; p += nt) {
    const float x = pts[p * 3], y = pts[p * 3 + 1], z = pts[p * 3 + 2];
    atomicAdd(&hist[cell_of(x, y, z)], 1);
    if (is_far_sentinel(x)) continue;  // invisible point (voxel-compat mode): listed under the last cell, no occupancy, no box
    const int cx = grid_coord(x), cy = grid_coord(y), cz = grid_coord(z);
    atomicMin(&box[0], cx); atomicMin(&box[1], cy); atomicMin(&box[2], cz);
    atomicMax(&box[3], cx); atomicMax(&box[4], cy); atomicMax(&box[5], cz);
    for (int dz = -1; dz <= 1; ++dz)
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          const int qx = cx + dx, qy = cy + dy, qz = cz + dz;
          if (qx < 0 || qy < 0 || qz < 0 || qx >= kGrid || qy >= kGrid || qz >= kGrid) continue;
          const int c = (qz * kGrid + qy) * kGrid + qx;
          atomicOr(&bits[c >> 5], 1u << (c & 31));
        }
  }
  __syncthreads();
  // exclusive scan of hist over kGridCells entries: each thread scans a contiguous chunk
  const int per = (kGridCells + nt - 1) / nt;
  const int lo = min(tid * per, kGridCells), hi = min(lo + per, kGridCells);
  int s = 0;
  for (int c = lo; c < hi; ++c) s += hist[c];
  chunk_sum[tid] = s;
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int t = 0; t < nt; ++t) { const int v = chunk_sum[t]; chunk_sum[t] = run; run += v; }
  }
  __syncthreads();
  int run = chunk_sum[tid];
  for (int c = lo; c < hi; ++c) { const int v = hist[c]; hist[c] = run; run += v; }
  if (tid == 0) hist[kGridCells] = P;
  __syncthreads();
  int* cs = cell_start + (size_t)b * (kGridCells + 1);
  for (int c = tid; c <= kGridCells; c += nt) cs[c] = hist[c];
  uint32_t* ob = occ_bits + (size_t)b * kGridWords;
  for (int c = tid; c < kGridWords; c += nt) ob[c] = bits[c];
  if (aabb && tid < 3) {  // dilated cell range [lo - 1, hi + 1] in world units, with a small margin for the cell-boundary rounding
    const int lo_c = box[tid] - 1, hi_c = box[3 + tid] + 1;
    aabb[b * 6 + tid] = lo_c <= 0 ? -INFINITY : (float)lo_c * (2.0f / kGrid) - 1.0f - 1e-4f;
    aabb[b * 6 + 3 + tid] = hi_c >= kGrid - 1 ? INFINITY : (float)(hi_c + 1) * (2.0f / kGrid) - 1.0f + 1e-4f;
  }
  // stable placement: rank within the cell = #points with the same cell and a lower index (deterministic, P is small)
  float4* sp = sorted_pts + (size_t)b * P;
  for (int p = tid; p < P; p += nt) {
    const float x = pts[p * 3], y = pts[p * 3 + 1], z = pts[p * 3 + 2];
    const int c = cell_of(x, y, z);
    int rank = 0;
    for (int q = 0; q < p; ++q) rank += (cell_of(pts[q * 3], pts[q * 3 + 1], pts[q * 3 + 2]) == c);
    sp[hist[c] + rank] = make_float4(x, y, z, __int_as_float(p));
  }
}

// Sub-cell classification for the marcher: every grid cell is split into 4^3 sub-cells of edge 1/48 and gets two 64-bit masks
// (bit (sz*4+sy)*4+sx):
//   sure  : some point is closer than r - margin to EVERY position of the sub-cell  -> a sample in it is valid, no exact test needed
//   maybe : some point is closer than r + margin to SOME position of the sub-cell   -> a sample outside every maybe sub-cell is invalid
// Only samples in maybe-but-not-sure sub-cells (a shell of ~0.036 around the surface of the union of r-balls) take the exact test,
// so the classification never changes a result (the margin, 2e-5, is 100x the rounding of the fp32 distance and of the sub-cell
// assignment).  Masks are zeroed by the host wrapper (memset) before the kernel ORs into them.
constexpr int kFine = kGrid * 4;
constexpr int kMaskLayers = 12;  // fine z layers a point can reach: 2 (r + margin) * 48 + 3 <= 12 for r <= 1/12

// one thread per (point, fine z layer): tests the <= 12 x 12 sub-cells of that layer, ORs one mask pair per touched grid cell
__global__ void __launch_bounds__(128) k_grid_masks(const float* __restrict__ kp_pos, int P, float radius,
                                                    unsigned long long* __restrict__ masks /* [n_obj, cells, 2] = (sure, maybe) */) {
  const int b = blockIdx.y;
  const int item = blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= P * kMaskLayers) return;
  const int p = item / kMaskLayers, layer = item % kMaskLayers;
  const float* pts = kp_pos + (size_t)b * P * 3;
  unsigned long long* m = masks + (size_t)b * kGridCells * 2;
  const float margin = 2e-5f;
  const float r_maybe = radius + margin, r_sure = radius - margin;
  const float r2_maybe = r_maybe * r_maybe, r2_sure = r_sure > 0.f ? r_sure * r_sure : 0.f;
  const float h = 2.0f / kFine;
  const float c3[3] = {pts[p * 3], pts[p * 3 + 1], pts[p * 3 + 2]};
  if (is_far_sentinel(c3[0])) return;
  int lo[3], hi[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    lo[a] = min(max((int)floorf((c3[a] - r_maybe + 1.0f) * (0.5f * kFine)) - 1, 0), kFine - 1);
    hi[a] = min(max((int)floorf((c3[a] + r_maybe + 1.0f) * (0.5f * kFine)) + 1, 0), kFine - 1);
  }
  const int fz = lo[2] + layer;
  if (fz > hi[2]) return;
  const int cz = fz >> 2;
  const float zl = (float)fz * h - 1.0f, zh = (float)(fz + 1) * h - 1.0f;
  const float zn = fmaxf(fmaxf(zl - c3[2], c3[2] - zh), 0.f), zx = fmaxf(c3[2] - zl, zh - c3[2]);
  for (int cy = lo[1] >> 2; cy <= hi[1] >> 2; ++cy)
    for (int cx = lo[0] >> 2; cx <= hi[0] >> 2; ++cx) {
      unsigned long long sure = 0ull, maybe = 0ull;
      for (int fy = max(lo[1], cy * 4); fy <= min(hi[1], cy * 4 + 3); ++fy) {
        const float yl = (float)fy * h - 1.0f, yh = (float)(fy + 1) * h - 1.0f;
        const float yn = fmaxf(fmaxf(yl - c3[1], c3[1] - yh), 0.f), yx = fmaxf(c3[1] - yl, yh - c3[1]);
        for (int fx = max(lo[0], cx * 4); fx <= min(hi[0], cx * 4 + 3); ++fx) {
          const float xl = (float)fx * h - 1.0f, xh = (float)(fx + 1) * h - 1.0f;
          const float xn = fmaxf(fmaxf(xl - c3[0], c3[0] - xh), 0.f), xx = fmaxf(c3[0] - xl, xh - c3[0]);
          const int bit = ((fz & 3) * 4 + (fy & 3)) * 4 + (fx & 3);
          if (xn * xn + yn * yn + zn * zn < r2_maybe) maybe |= 1ull << bit;
          if (xx * xx + yx * yx + zx * zx < r2_sure) sure |= 1ull << bit;
        }
      }
      const int c = (cz * kGrid + cy) * kGrid + cx;
      if (sure) atomicOr(m + 2 * c, sure);
      if (maybe) atomicOr(m + 2 * c + 1, maybe);
    }
}

}  // namespace npcd

extern "C" int npcd_grid_build_masks(const float* kp_pos, int n_obj, int n_points, float radius, void* masks, void* stream) {
  using namespace npcd;
  NPCD_CHECK_ARG(kp_pos && masks, "null pointer");
  NPCD_CHECK_ARG(n_obj >= 0 && n_points > 0, "bad n_obj / n_points");
  NPCD_CHECK_ARG(radius > 0.f && radius <= 2.0f / kGrid, "radius must be in (0, 1/12] (grid cell edge)");
  if (n_obj == 0) return 0;
  cudaMemsetAsync(masks, 0, (size_t)n_obj * kGridCells * 2 * sizeof(unsigned long long), (cudaStream_t)stream);
  k_grid_masks<<<dim3((n_points * kMaskLayers + 127) / 128, n_obj), 128, 0, (cudaStream_t)stream>>>(kp_pos, n_points, radius,
                                                                                                      (unsigned long long*)masks);
  return check_launch("npcd_grid_build_masks");
}

extern "C" int npcd_grid_build(const float* kp_pos, int n_obj, int n_points, int* cell_start, float* sorted_pts,
                               unsigned* occ_bits, float* aabb, void* stream) {
  using namespace npcd;
  NPCD_CHECK_ARG(kp_pos && cell_start && sorted_pts && occ_bits, "null pointer");
  NPCD_CHECK_ARG(n_obj >= 0 && n_points > 0 && n_points <= (1 << 20), "bad n_obj / n_points");
  if (n_obj == 0) return 0;
  const size_t smem = (size_t)(kGridCells + 1 + kGridWords) * sizeof(int);
  cudaFuncSetAttribute(k_grid_build, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_grid_build<<<n_obj, 512, smem, (cudaStream_t)stream>>>(kp_pos, n_points, cell_start, (float4*)sorted_pts, occ_bits, aabb);
  return check_launch("npcd_grid_build");
}

extern "C" int npcd_grid_dims(int* cells, int* words) {
  if (cells) *cells = npcd::kGridCells;
  if (words) *words = npcd::kGridWords;
  return 0;
}
