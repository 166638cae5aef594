// Shared device/host helpers for libnpcd_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace npcd {

// Compile-time specialisation on the reference's hard-coded option tree
// (npcd/models/pointnerf/pointnerf.py:134-194).
constexpr int kK = 8;             // neighbours per shading sample            (pointnerf.py:170)
constexpr int kDepthRes = 128;    // depth samples per ray                    (pointnerf.py:184)
constexpr int kMaxShading = 50;   // default max shading samples per ray      (pointnerf.py:172)
constexpr int kHidden = 256;      // MLP width                                (pointnerf.py:161,176)
constexpr int kFreqs = 10;        // positional-encoding octaves              (pointnerf.py:174)
constexpr int kGrid = 24;         // acceleration grid: 24^3 cells of 1/12 > r = 0.08 over [-1,1]^3
constexpr int kGridCells = kGrid * kGrid * kGrid;
constexpr int kGridWords = kGridCells / 32;
// Position of a point that must stay invisible to every query (voxel-compat mode: points beyond a voxel's cap).  npcd_grid_build
// files such points under the last cell without marking occupancy; every distance test against them fails.
constexpr float kFarSentinel = 1e9f;
__host__ __device__ inline bool is_far_sentinel(float x) { return !(fabsf(x) < 1e8f); }

void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define NPCD_CHECK_ARG(cond, msg)                       \
  do {                                                  \
    if (!(cond)) {                                      \
      npcd::set_error("%s: %s", __func__, msg);         \
      return 1;                                         \
    }                                                   \
  } while (0)

// Monotone float <-> uint32 encoding so atomicMin/atomicMax order like the floats (NaN excluded by callers).
__host__ __device__ inline uint32_t f2ord(float f) {
#ifdef __CUDA_ARCH__
  uint32_t b = __float_as_uint(f);
#else
  uint32_t b;
  memcpy(&b, &f, 4);
#endif
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ inline float ord2f(uint32_t u) {
  uint32_t b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
  return __uint_as_float(b);
#else
  float f;
  memcpy(&f, &b, 4);
  return f;
#endif
}

// Depth of sample i on a ray; bit-exact restatement of math_utils.py:106-115 + renderer.py:74-76
// (every operation individually rounded, no FMA contraction).
__device__ __forceinline__ float sample_depth(float start, float end, int i, const float* jitter_row) {
  float span = __fsub_rn(end, start);
  float step = __fdiv_rn((float)i, (float)(kDepthRes - 1));
  float t = __fadd_rn(start, __fmul_rn(step, span));
  if (jitter_row) {
    float delta = __fdiv_rn(span, (float)(kDepthRes - 1));
    t = __fadd_rn(t, __fmul_rn(jitter_row[i], delta));
  }
  return t;
}

// x = o + t*d, multiply and add rounded separately (volume_renderer.py:70).
__device__ __forceinline__ float axpy_rn(float o, float t, float d) { return __fadd_rn(o, __fmul_rn(t, d)); }

// Squared distance in the oracle's op order: (dx*dx + dy*dy) + dz*dz.
__device__ __forceinline__ float dist_rn(float x, float y, float z, float px, float py, float pz) {
  float dx = __fsub_rn(x, px), dy = __fsub_rn(y, py), dz = __fsub_rn(z, pz);
  float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
  return __fsqrt_rn(d2);
}

__device__ __forceinline__ int grid_coord(float v) {
  int c = (int)floorf((v + 1.0f) * (0.5f * kGrid));
  return min(max(c, 0), kGrid - 1);
}

}  // namespace npcd
