// C1-C3: slot depths, alpha and front-to-back compositing, forward and backward.
// Replaces Renderer.get_depths_from_shading_pts (renderers/renderer.py:95-110), VolumeRenderer.get_alpha
// (renderers/volume_renderer.py:23-39) and Renderer.ray_march (renderers/renderer.py:120-185) plus the dense scatters that feed
// them (fields/field.py:62-75,130-143), on the COMPACT per-ray sample lists produced by march.cu (exact mode: a ray's kept
// samples occupy slots 0..n-1, aggregator.py:57-58).
//
// Reference quirks reproduced (SURVEY.md finding 7):
//   * slots >= n inherit the last valid depth (cummax), so delta_{n-1} = 0 and the LAST kept sample has alpha = 0;
//   * rays without samples have all slot depths = ray_end; their composite depth is 0/0 -> NaN -> +inf -> clamp;
//   * the clamp range is the GLOBAL [min, max] over every slot depth of every ray in the call (renderer.py:154-156).
// Warp per ray; lanes own consecutive samples; transmittance is a warp-shuffle product scan (chunks of 32 with carry).
#include "common.cuh"
#include "npcd_b200.h"

namespace npcd {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// inclusive product scan over the warp
__device__ __forceinline__ float warp_scan_mul(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float u = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v *= u;
  }
  return v;
}
// inclusive suffix sum over the warp (lane i gets sum_{k>=i})
__device__ __forceinline__ float warp_rscan_add(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float u = __shfl_down_sync(0xffffffffu, v, o);
    if (lane + o < 32) v += u;
  }
  return v;
}

struct SampleTerms {
  float alpha, f, edd, delta, t;  // f = 1 - alpha + 1e-10 ; edd = exp(-sigma*delta)
  float4 c;                       // r, g, b, sigma
};

// `slot` (voxel-compat mode, else null): slot index of every kept sample among its ray's candidates.  A kept sample whose successor
// does not sit in the very next slot is followed by a HOLE, whose depth is the running maximum = its own: delta = 0, alpha = 0
// (renderer.py:106-108 cummax, volume_renderer.py:35-38).
__device__ __forceinline__ SampleTerms load_terms(const float4* __restrict__ sample_pos, const float* __restrict__ sample_t,
                                                  const float4* __restrict__ rgbs, const unsigned char* __restrict__ slot,
                                                  long long off, int i, int n) {
  SampleTerms s;
  const bool valid = i < n;
  auto depth_of = [&](long long k) { return sample_t ? __ldg(sample_t + k) : __ldg(&sample_pos[k].w); };
  s.t = valid ? depth_of(off + i) : 0.f;
  bool has_next = i + 1 < n;
  if (has_next && slot) has_next = __ldg(slot + off + i + 1) == __ldg(slot + off + i) + 1;
  const float nxt = has_next ? depth_of(off + i + 1) : s.t;  // slot n inherits t_{n-1}: delta_{n-1} = 0
  s.c = valid ? __ldg(&rgbs[off + i]) : make_float4(0.f, 0.f, 0.f, 0.f);
  s.delta = nxt - s.t;
  s.edd = expf(-(s.c.w * s.delta));
  s.alpha = valid ? 1.0f - s.edd : 0.f;
  s.f = (1.0f - s.alpha) + 1e-10f;
  return s;
}

// Forward: a warp owns 32 CONSECUTIVE rays.  Lanes first fetch their own ray's offsets (coalesced); three of four rays of the eval
// workload carry no sample and are finished right there by their lane; the rays with samples are then composited one after the
// other by the whole warp (lanes own consecutive samples), each result parked in the owning lane, and all 32 outputs leave with
// coalesced stores.  Slot depths come from `sample_t` ([S], written by the kNN kernel next to the positions: 4 B per sample
// instead of a 16-byte sector for one float) or, when it is null, from sample_pos.w.
__global__ void __launch_bounds__(256) k_composite_fwd(const float4* __restrict__ sample_pos, const float* __restrict__ sample_t,
                                                       const float4* __restrict__ rgbs,
                                                       const long long* __restrict__ ray_offset,
                                                       const int* __restrict__ ray_ids, const float* __restrict__ ray_end,
                                                       const unsigned char* __restrict__ slot,
                                                       long long n_sel, int white_back, float* __restrict__ out_mask,
                                                       float* __restrict__ out_depth, float* __restrict__ out_rgb,
                                                       uint32_t* __restrict__ range_ord) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long sel = ((long long)blockIdx.x * (blockDim.x >> 5) + warp) * 32 + lane;
  const bool live = sel < n_sel;
  uint32_t dmin = 0xffffffffu, dmax = 0u;
  long long my_off = 0;
  int my_n = 0;
  if (live) {
    my_off = ray_offset[sel];
    my_n = (int)(ray_offset[sel + 1] - my_off);
  }
  // result of this lane's ray; a sample-free ray: mask 0, depth 0/0 -> NaN -> +inf (clamped later), colour = background
  float o_mask = 0.f, o_depth = __int_as_float(0x7f800000), o_r = white_back ? 1.f : 0.f, o_g = o_r, o_b = o_r;
  if (live && my_n == 0) dmin = dmax = f2ord(ray_end[ray_ids ? ray_ids[sel] : sel]);  // all slot depths = ray_end (renderer.py:109-110)
  uint32_t todo = __ballot_sync(0xffffffffu, live && my_n > 0);
  while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    const long long off = __shfl_sync(0xffffffffu, my_off, src);
    const int n = __shfl_sync(0xffffffffu, my_n, src);
    float carry = 1.f, sw = 0.f, swt = 0.f, r = 0.f, g = 0.f, b = 0.f;
    for (int base = 0; base < n; base += 32) {
      const int i = base + lane;
      const SampleTerms s = load_terms(sample_pos, sample_t, rgbs, slot, off, i, n);
      const float incl = warp_scan_mul(s.f, lane);
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 1.f;
      const float w = s.alpha * (carry * excl);
      sw += w; swt += w * s.t; r += w * s.c.x; g += w * s.c.y; b += w * s.c.z;
      carry *= __shfl_sync(0xffffffffu, incl, 31);
      if (i < n) { dmin = min(dmin, f2ord(s.t)); dmax = max(dmax, f2ord(s.t)); }
    }
    sw = warp_sum(sw); swt = warp_sum(swt); r = warp_sum(r); g = warp_sum(g); b = warp_sum(b);
    if (lane == src) {
      if (slot && __ldg(slot + off) != 0) {  // leading holes carry ray_end (renderer.py:109-110): it joins the clamp range
        const uint32_t e = f2ord(ray_end[ray_ids ? ray_ids[sel] : sel]);
        dmin = min(dmin, e); dmax = max(dmax, e);
      }
      float d = swt / sw;
      if (d != d) d = __int_as_float(0x7f800000);  // nan_to_num(nan -> +inf)  (renderer.py:154)
      const float bg = white_back ? 1.0f - sw : 0.f;
      o_mask = sw; o_depth = d; o_r = r + bg; o_g = g + bg; o_b = b + bg;
    }
  }
  if (live) {
    out_mask[sel] = o_mask;
    out_depth[sel] = o_depth;
    out_rgb[sel * 3 + 0] = o_r; out_rgb[sel * 3 + 1] = o_g; out_rgb[sel * 3 + 2] = o_b;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    dmin = min(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
    dmax = max(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
  }
  __shared__ uint32_t smin[8], smax[8];
  if (lane == 0) { smin[warp] = dmin; smax[warp] = dmax; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { dmin = min(dmin, smin[w]); dmax = max(dmax, smax[w]); }
    if (dmin != 0xffffffffu) { atomicMin(&range_ord[0], dmin); atomicMax(&range_ord[1], dmax); }
  }
}

__global__ void k_clamp_depth(float* __restrict__ depth, long long n, const uint32_t* __restrict__ range_ord,
                              unsigned char* __restrict__ clamped) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float lo = ord2f(range_ord[0]), hi = ord2f(range_ord[1]);
  const float d = depth[i];
  const float c = fminf(fmaxf(d, lo), hi);
  depth[i] = c;
  if (clamped) clamped[i] = (d < lo || d > hi) ? 1 : 0;  // gradient passes only strictly inside (torch.clamp semantics)
}

__global__ void k_init_range(uint32_t* p) { p[0] = 0xffffffffu; p[1] = 0u; }

// Backward.  Per ray, with dw_i = dL/dw_i:
//   dL/dc_i     = w_i * g_rgb
//   dw_i        = g_rgb . c_i  - (white_back ? sum(g_rgb) : 0) + g_mask + g_depth * (t_i - D) / M   [depth term if unclamped]
//   dL/dalpha_i = T_i * dw_i - (sum_{k>i} dw_k w_k) / f_i          (autograd of the exclusive cumprod)
//   dL/dsigma_i = dL/dalpha_i * delta_i * exp(-sigma_i delta_i)
__global__ void __launch_bounds__(256) k_composite_bwd(const float4* __restrict__ sample_pos, const float* __restrict__ sample_t,
                                                       const float4* __restrict__ rgbs,
                                                       const long long* __restrict__ ray_offset, long long n_sel,
                                                       int white_back, const float* __restrict__ g_rgb,
                                                       const float* __restrict__ g_mask, const float* __restrict__ g_depth,
                                                       const float* __restrict__ out_mask, const float* __restrict__ out_depth,
                                                       const unsigned char* __restrict__ clamped,
                                                       const unsigned char* __restrict__ slot, float4* __restrict__ g_rgbs) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long sel = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (sel >= n_sel) return;
  const long long off = ray_offset[sel];
  const int n = (int)(ray_offset[sel + 1] - off);
  if (n == 0) return;
  const float gr = g_rgb ? g_rgb[sel * 3] : 0.f, gg = g_rgb ? g_rgb[sel * 3 + 1] : 0.f, gb = g_rgb ? g_rgb[sel * 3 + 2] : 0.f;
  const float gm = (g_mask ? g_mask[sel] : 0.f) - (white_back ? (gr + gg + gb) : 0.f);
  const float M = out_mask[sel];
  float gd = 0.f, D = 0.f;
  if (g_depth && !(clamped && clamped[sel]) && M != 0.f) { gd = g_depth[sel] / M; D = out_depth[sel]; }
  // forward sweep: transmittance carry at the start of every 32-chunk (n <= 128 -> <= 4 chunks)
  float carries[kDepthRes / 32];
  {
    float carry = 1.f;
#pragma unroll
    for (int ch = 0; ch < kDepthRes / 32; ++ch) {
      carries[ch] = carry;
      if (ch * 32 < n) {
        const SampleTerms s = load_terms(sample_pos, sample_t, rgbs, slot, off, ch * 32 + lane, n);
        const float incl = warp_scan_mul(s.f, lane);
        carry *= __shfl_sync(0xffffffffu, incl, 31);
      }
    }
  }
  float suffix_carry = 0.f;  // sum_{k in later chunks} dw_k w_k
#pragma unroll
  for (int ch = kDepthRes / 32 - 1; ch >= 0; --ch) {
    if (ch * 32 < n) {
    const int i = ch * 32 + lane;
    const SampleTerms s = load_terms(sample_pos, sample_t, rgbs, slot, off, i, n);
    const float incl = warp_scan_mul(s.f, lane);
    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 1.f;
    const float T = carries[ch] * excl;
    const float w = s.alpha * T;
    const float dw = gr * s.c.x + gg * s.c.y + gb * s.c.z + gm + gd * (s.t - D);
    const float q = (i < n) ? dw * w : 0.f;
    const float incl_suffix = warp_rscan_add(q, lane);          // sum_{k>=i in chunk}
    const float suffix = incl_suffix - q + suffix_carry;        // sum_{k>i}
    suffix_carry += __shfl_sync(0xffffffffu, incl_suffix, 0);
    if (i < n) {
      const float dalpha = T * dw - suffix / s.f;
      g_rgbs[off + i] = make_float4(w * gr, w * gg, w * gb, dalpha * s.delta * s.edd);
    }
    }
  }
}

}  // namespace npcd

extern "C" int npcd_composite_fwd(const float* sample_pos, const float* sample_t, const float* rgbs, const long long* ray_offset,
                                  const int* ray_ids, const float* ray_end, const unsigned char* slot, long long n_sel,
                                  int white_back, float* out_mask, float* out_depth, float* out_rgb, void* range_scratch,
                                  int init_range, void* stream) {
  using namespace npcd;
  NPCD_CHECK_ARG(ray_offset && ray_end && out_mask && out_depth && out_rgb && range_scratch, "null pointer");
  NPCD_CHECK_ARG(n_sel >= 0, "bad n_sel");
  cudaStream_t st = (cudaStream_t)stream;
  uint32_t* rng = (uint32_t*)range_scratch;
  if (init_range) k_init_range<<<1, 1, 0, st>>>(rng);
  if (n_sel > 0) {
    const int rpb = 8 * 32;  // 8 warps x 32 rays
    k_composite_fwd<<<(unsigned)((n_sel + rpb - 1) / rpb), 256, 0, st>>>((const float4*)sample_pos, sample_t, (const float4*)rgbs,
                                                                        ray_offset, ray_ids, ray_end, slot, n_sel, white_back,
                                                                        out_mask, out_depth, out_rgb, rng);
  }
  return check_launch("npcd_composite_fwd");
}

// renderer.py:154-156: clamp to the global slot-depth range accumulated in range_scratch by the composite call(s).
extern "C" int npcd_clamp_depth(float* depth, long long n, const void* range_scratch, unsigned char* out_clamped, void* stream) {
  using namespace npcd;
  NPCD_CHECK_ARG(range_scratch && (n == 0 || depth), "null pointer");
  if (n <= 0) return 0;
  k_clamp_depth<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(depth, n, (const uint32_t*)range_scratch, out_clamped);
  return check_launch("npcd_clamp_depth");
}

extern "C" int npcd_composite_bwd(const float* sample_pos, const float* sample_t, const float* rgbs, const long long* ray_offset,
                                  const unsigned char* slot, long long n_sel, int white_back, const float* g_rgb,
                                  const float* g_mask, const float* g_depth, const float* out_mask, const float* out_depth,
                                  const unsigned char* clamped, float* g_rgbs, void* stream) {
  using namespace npcd;
  NPCD_CHECK_ARG(ray_offset && out_mask && out_depth && g_rgbs, "null pointer");
  NPCD_CHECK_ARG(n_sel >= 0, "bad n_sel");
  if (n_sel == 0) return 0;
  const int wpb = 8;
  k_composite_bwd<<<(unsigned)((n_sel + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
      (const float4*)sample_pos, sample_t, (const float4*)rgbs, ray_offset, n_sel, white_back, g_rgb, g_mask, g_depth, out_mask,
      out_depth, clamped, slot, (float4*)g_rgbs);
  return check_launch("npcd_composite_bwd");
}


// ---- decode post-processing of `eval_diffusion` (npcd/eval/diffusion_evaluation.py:169-173):
// unflatten_pred (npcd/utils/util.py:199-203: [n, R, 3] -> [n, 3, H, W]) fused with np.clip(., 0, 1) and np.round(. * 255) / 255
// (round-half-to-even, like numpy), so the images feed the Inception network without a host round trip.
namespace npcd {
__global__ void k_channels_to_images(const float* __restrict__ channels, long long n_views, int n_pix, int quantize,
                                     float* __restrict__ images) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // (view, pixel)
  if (i >= n_views * n_pix) return;
  const long long v = i / n_pix;
  const int p = (int)(i - v * n_pix);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float x = __ldg(channels + i * 3 + c);
    if (quantize) {
      if (x == x) x = fminf(fmaxf(x, 0.f), 1.f);  // np.clip keeps NaN; fminf/fmaxf would silently turn it into a bound
      x = __fdiv_rn(rintf(__fmul_rn(x, 255.f)), 255.f);
    }
    images[(v * 3 + c) * n_pix + p] = x;
  }
}
}  // namespace npcd

extern "C" int npcd_channels_to_images(const float* channels, long long n_views, int resolution, int quantize, float* images,
                                       void* stream) {
  using namespace npcd;
  NPCD_CHECK_ARG(n_views >= 0 && resolution > 0, "bad sizes");
  if (n_views == 0) return 0;
  NPCD_CHECK_ARG(channels && images, "null pointer");
  const long long n = n_views * resolution * resolution;
  k_channels_to_images<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(channels, n_views, resolution * resolution, quantize,
                                                                                     images);
  return check_launch("npcd_channels_to_images");
}
