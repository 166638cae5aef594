// Voxel-grid-compatible query mode (SURVEY.md section 8(a) row Q1, Appendix A.4 "Voxel mode"): the semantics the reference's DEPLOYED
// query path has through torch_knnquery.VoxelGrid (fields/aggregators/aggregator.py:59-76, options pointnerf.py:147-153), as far as
// the call sites, the option names and the TV loss's remark that the grid "looses keypoints" (losses/neural_point_cloud_tv_loss.py:42)
// pin them down -- the extension's source is not part of the reference (pip-from-git HEAD, no pin): PARITY UNPINNED, checked against
// oracle/pointnerf_oracle.py::query_keypoints_voxel only.
//
//   * voxels of voxel_size * voxel_scale (0.08) over `ranges` ([-1,1]^3: 25^3); a voxel stores at most max_points_per_voxel (4)
//     points, LOWEST index wins (upstream: atomics order); points beyond the cap or outside the ranges are invisible to queries;
//   * candidate shading samples = samples inside the ranges whose voxel lies in the kernel_size (3^3) dilation of the occupied
//     voxels; the first max_shading_pts (50) CANDIDATES of a ray take slots 0, 1, ...; a candidate without a stored point within
//     r leaves a HOLE (aggregator.py:66-70), and a sample followed by a hole gets alpha = 0 downstream (volume_renderer.py:35-38);
//   * neighbours = the <= 8 nearest stored points within r (r = voxel edge, so the 27 surrounding voxels hold them all).
//
// Built on the exact-mode kernels: npcd_voxel_select hands the STORED points (dropped ones moved to a far sentinel that grid.cu keeps
// out of its occupancy) to npcd_grid_build; npcd_march_count then marks the samples with a stored neighbour; npcd_voxel_filter
// intersects them with the capped candidate set; npcd_voxel_slots numbers the kept samples' slots for the compositor.
#include "common.cuh"
#include "npcd_b200.h"

namespace npcd {

constexpr int kVoxMax = 32;  // voxels per axis supported by the shared-memory bit set (32^3 bits = 4 KB)

__device__ __forceinline__ int vox_coord(float v, float lo, float vsize) {  // np.floor((v - lo) / vsize), fp32, separately rounded
  return (int)floorf(__fdiv_rn(__fsub_rn(v, lo), vsize));
}

__global__ void __launch_bounds__(256) k_voxel_select(const float* __restrict__ kp_pos, int P, float vsize, float lo, int n_vox,
                                                      int max_per_voxel, int dil, float* __restrict__ stored_pos,
                                                      uint32_t* __restrict__ vox_bits, int words) {
  __shared__ uint32_t bits[kVoxMax * kVoxMax * kVoxMax / 32];
  const int b = blockIdx.x;
  const float* pts = kp_pos + (size_t)b * P * 3;
  for (int w = threadIdx.x; w < words; w += blockDim.x) bits[w] = 0u;
  __syncthreads();
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    const float x = pts[p * 3], y = pts[p * 3 + 1], z = pts[p * 3 + 2];
    const int cx = vox_coord(x, lo, vsize), cy = vox_coord(y, lo, vsize), cz = vox_coord(z, lo, vsize);
    const bool inb = cx >= 0 && cy >= 0 && cz >= 0 && cx < n_vox && cy < n_vox && cz < n_vox;
    bool stored = false;
    if (inb) {
      int rank = 0;  // points of the same voxel with a lower index (P is a few hundred)
      for (int q = 0; q < p && rank < max_per_voxel; ++q)
        rank += (vox_coord(pts[q * 3], lo, vsize) == cx && vox_coord(pts[q * 3 + 1], lo, vsize) == cy &&
                 vox_coord(pts[q * 3 + 2], lo, vsize) == cz);
      stored = rank < max_per_voxel;
      if (rank == 0) {  // first point of its voxel: the voxel is occupied -> mark its dilation
        for (int dz = -dil; dz <= dil; ++dz)
          for (int dy = -dil; dy <= dil; ++dy)
            for (int dx = -dil; dx <= dil; ++dx) {
              const int qx = cx + dx, qy = cy + dy, qz = cz + dz;
              if (qx < 0 || qy < 0 || qz < 0 || qx >= n_vox || qy >= n_vox || qz >= n_vox) continue;
              const int c = (qx * n_vox + qy) * n_vox + qz;
              atomicOr(&bits[c >> 5], 1u << (c & 31));
            }
      }
    }
    float* o = stored_pos + ((size_t)b * P + p) * 3;
    o[0] = stored ? x : kFarSentinel; o[1] = stored ? y : kFarSentinel; o[2] = stored ? z : kFarSentinel;
  }
  __syncthreads();
  for (int w = threadIdx.x; w < words; w += blockDim.x) vox_bits[(size_t)b * words + w] = bits[w];
}

// warp per ray: candidate bits (capped at the first max_shading candidates), valid &= candidates, counts
__global__ void __launch_bounds__(256) k_voxel_filter(const float* __restrict__ cam, const float* __restrict__ dirs,
                                                      const float* __restrict__ start, const float* __restrict__ end,
                                                      const float* __restrict__ jitter, long long n_rays, int rays_per_view,
                                                      int views_per_obj, const uint32_t* __restrict__ vox_bits, int words, int n_vox,
                                                      float vsize, float lo, int max_shading, uint32_t* __restrict__ valid_bits,
                                                      uint32_t* __restrict__ cand_bits, int* __restrict__ ray_count) {
  const int lane = threadIdx.x & 31;
  const long long ray = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ray >= n_rays) return;
  const int view = (int)(ray / rays_per_view);
  const int obj = view / views_per_obj;
  const uint32_t* vb = vox_bits + (size_t)obj * words;
  const float ox = cam[view * 3], oy = cam[view * 3 + 1], oz = cam[view * 3 + 2];
  const float dx = dirs[ray * 3], dy = dirs[ray * 3 + 1], dz = dirs[ray * 3 + 2];
  const float t0 = start[ray], t1 = end[ray];
  const float* jit = jitter ? jitter + ray * kDepthRes : nullptr;
  int n_cand = 0, total = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int i = j * 32 + lane;
    const float t = sample_depth(t0, t1, i, jit);
    const float x = axpy_rn(ox, t, dx), y = axpy_rn(oy, t, dy), z = axpy_rn(oz, t, dz);
    const int cx = vox_coord(x, lo, vsize), cy = vox_coord(y, lo, vsize), cz = vox_coord(z, lo, vsize);
    bool cand = cx >= 0 && cy >= 0 && cz >= 0 && cx < n_vox && cy < n_vox && cz < n_vox;
    if (cand) {
      const int c = (cx * n_vox + cy) * n_vox + cz;
      cand = (__ldg(vb + (c >> 5)) >> (c & 31)) & 1u;
    }
    const uint32_t raw = __ballot_sync(0xffffffffu, cand);
    const int before = n_cand + __popc(raw & ((1u << lane) - 1u));  // candidates in front of this sample
    const uint32_t capped = __ballot_sync(0xffffffffu, cand && before < max_shading);
    n_cand += __popc(raw);
    const uint32_t v = __ldg(valid_bits + ray * 4 + j) & capped;
    total += __popc(v);
    if (lane == j) { valid_bits[ray * 4 + j] = v; cand_bits[ray * 4 + j] = capped; }
  }
  if (lane == 0) ray_count[ray] = total;
}

// thread per selected ray: slot (= index among the ray's candidates) of every kept sample, in the compact order of knn_fill
__global__ void __launch_bounds__(128) k_voxel_slots(const long long* __restrict__ ray_offset, const int* __restrict__ ray_ids,
                                                     const uint32_t* __restrict__ valid_bits, const uint32_t* __restrict__ cand_bits,
                                                     long long n_sel, long long capacity, unsigned char* __restrict__ slot) {
  const long long sel = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (sel >= n_sel) return;
  const long long ray = ray_ids ? (long long)ray_ids[sel] : sel;
  long long s = ray_offset[sel];
  const long long s_end = min(ray_offset[sel + 1], capacity);
  int cand_before = 0;
  for (int j = 0; j < 4 && s < s_end; ++j) {
    uint32_t v = __ldg(valid_bits + ray * 4 + j);
    const uint32_t c = __ldg(cand_bits + ray * 4 + j);
    while (v && s < s_end) {
      const int bit = __ffs(v) - 1;
      v &= v - 1;
      slot[s++] = (unsigned char)(cand_before + __popc(c & ((1u << bit) - 1u)));
    }
    cand_before += __popc(c);
  }
}

}  // namespace npcd

using namespace npcd;

extern "C" int npcd_voxel_dims(float voxel_size, float range_lo, float range_hi, int* n_vox, int* words) {
  NPCD_CHECK_ARG(n_vox && words && voxel_size > 0.f && range_hi > range_lo, "bad arguments");
  const int n = (int)lroundf((range_hi - range_lo) / voxel_size);
  NPCD_CHECK_ARG(n >= 1 && n <= kVoxMax, "at most 32 voxels per axis");
  *n_vox = n;
  *words = (n * n * n + 31) / 32;
  return 0;
}

extern "C" int npcd_voxel_select(const float* kp_pos, int n_obj, int n_points, float voxel_size, float range_lo, int n_vox,
                                 int max_points_per_voxel, int kernel_size, float* stored_pos, unsigned* vox_bits, void* stream) {
  NPCD_CHECK_ARG(kp_pos && stored_pos && vox_bits, "null pointer");
  NPCD_CHECK_ARG(n_obj >= 0 && n_points > 0, "bad n_obj / n_points");
  NPCD_CHECK_ARG(n_vox >= 1 && n_vox <= kVoxMax && voxel_size > 0.f, "bad voxel grid");
  NPCD_CHECK_ARG(max_points_per_voxel >= 1 && kernel_size >= 1 && (kernel_size & 1), "bad cap / kernel size (odd)");
  if (n_obj == 0) return 0;
  const int words = (n_vox * n_vox * n_vox + 31) / 32;
  k_voxel_select<<<n_obj, 256, 0, (cudaStream_t)stream>>>(kp_pos, n_points, voxel_size, range_lo, n_vox, max_points_per_voxel,
                                                          kernel_size / 2, stored_pos, vox_bits, words);
  return check_launch("npcd_voxel_select");
}

extern "C" int npcd_voxel_filter(const float* cam_centers, const float* dirs, const float* ray_start, const float* ray_end,
                                 const float* jitter, long long n_rays, int rays_per_view, int views_per_obj, const unsigned* vox_bits,
                                 int n_vox, float voxel_size, float range_lo, int max_shading_pts, unsigned* valid_bits,
                                 unsigned* cand_bits, int* ray_count, void* stream) {
  NPCD_CHECK_ARG(n_rays >= 0 && rays_per_view > 0 && views_per_obj > 0, "bad sizes");
  if (n_rays == 0) return 0;
  NPCD_CHECK_ARG(cam_centers && dirs && ray_start && ray_end && vox_bits && valid_bits && cand_bits && ray_count, "null pointer");
  NPCD_CHECK_ARG(n_vox >= 1 && n_vox <= kVoxMax && voxel_size > 0.f && max_shading_pts > 0, "bad voxel grid / cap");
  const int words = (n_vox * n_vox * n_vox + 31) / 32;
  k_voxel_filter<<<(unsigned)((n_rays + 7) / 8), 256, 0, (cudaStream_t)stream>>>(cam_centers, dirs, ray_start, ray_end, jitter, n_rays,
                                                                                 rays_per_view, views_per_obj, vox_bits, words, n_vox,
                                                                                 voxel_size, range_lo, max_shading_pts, valid_bits,
                                                                                 cand_bits, ray_count);
  return check_launch("npcd_voxel_filter");
}

extern "C" int npcd_voxel_slots(const long long* ray_offset, const int* ray_ids, const unsigned* valid_bits, const unsigned* cand_bits,
                                long long n_sel, long long capacity, unsigned char* slot, void* stream) {
  NPCD_CHECK_ARG(n_sel >= 0 && capacity >= 0, "bad sizes");
  if (n_sel == 0 || capacity == 0) return 0;
  NPCD_CHECK_ARG(ray_offset && valid_bits && cand_bits && slot, "null pointer");
  k_voxel_slots<<<(unsigned)((n_sel + 127) / 128), 128, 0, (cudaStream_t)stream>>>(ray_offset, ray_ids, valid_bits, cand_bits, n_sel,
                                                                                   capacity, slot);
  return check_launch("npcd_voxel_slots");
}
