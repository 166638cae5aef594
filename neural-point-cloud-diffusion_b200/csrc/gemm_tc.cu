// Generic fp32-accurate GEMM on the tcgen05 tensor cores:  C[M,N] = act((A[M,K] . B[N,K]^T) * out_scale + bias),  N <= 256.
// It carries the dense layers of the TRAINING path (row B* of SURVEY.md section 8: forward, dgrad and wgrad of the twelve
// nn.Linear layers in fields/aggregators/mlp.py:34,84 and fields/mlp.py:38-72, which the reference runs as fp32 cuBLAS GEMMs
// under autograd).  Same numerics as the fused inference kernels (mlp_tc.cu): both operands are split into fp16 hi/lo halves and
// three kind::f16 products (hi*hi + lo*hi + hi*lo) accumulate in fp32 TMEM.
//
// Operands arrive as pre-split "operand images" built by k_pack_rows (any fp32 matrix, optionally transposed, optionally
// multiplied by a LeakyReLU-derivative mask, scaled by a power of two read from device memory):
//   image[row tile of 128][K-block of 64][hi 16 KB | lo 16 KB], every 16 KB block in the K-major SWIZZLE_128B layout.
// One CTA per (128-row output tile, K split): warp 0 streams A / B K-blocks into a 2-stage ring with cp.async.bulk, warp 1 issues
// the MMAs (M=128, N=round_up(N,16), K=16), warps 2..5 drain the accumulator.  Split-K partials go to a workspace and are
// reduced in a fixed order (deterministic) by k_splitk_reduce.
#include "npcd_b200.h"
#include "tc_ptx.cuh"

namespace npcd {
namespace tc {

constexpr int kGemmThreads = 192;
constexpr int kGemmStages = 2;
constexpr int kStageBytes = 6 * kTileBytesA;  // A hi, A lo, B hi (2 row tiles), B lo (2 row tiles) = 96 KB
constexpr int kGemmSmem = kGemmStages * kStageBytes + 1024;

struct GemmParams {
  const uint8_t* a_img;
  const uint8_t* b_img;
  int M, N;
  int nkb;           // K-blocks in the images (K padded to a multiple of 64)
  int kb_per_split;  // K-blocks per CTA along grid.y
  int n_mma;         // N rounded up to a multiple of 16
  float* C;
  long long ldc;
  const float* bias;           // [N] device, optional
  const float* out_scale_dev;  // device scalar, optional (inverse of the operand scales)
  float slope;                 // LeakyReLU slope of the fused activation (1 = none)
  float* partial;              // split-K workspace [splits][tiles_m * 128][256], or NULL
};

__global__ void __launch_bounds__(kGemmThreads, 1) k_gemm_tc(const __grid_constant__ GemmParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* misc = smem + kGemmStages * kStageBytes;
  const uint32_t bars = smem_u32(misc);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 64);
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };  // full[2], empty[2], acc_ready
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kGemmStages; ++i) { mbar_init(bar(i), 1); mbar_init(bar(kGemmStages + i), 1); }
    mbar_init(bar(2 * kGemmStages), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tile_m = blockIdx.x, split = blockIdx.y;
  const int kb0 = split * P.kb_per_split, kb1 = min(P.nkb, kb0 + P.kb_per_split);
  const int b_tiles = P.n_mma > 128 ? 2 : 1;

  if (warp == 0) {
    int st = 0;
    uint32_t ph = 0;
    for (int kb = kb0; kb < kb1; ++kb) {
      mbar_wait(bar(kGemmStages + st), ph ^ 1);
      if (elect_one()) {
        uint8_t* s = smem + st * kStageBytes;
        mbar_expect_tx(bar(st), (uint32_t)(2 + 2 * b_tiles) * kTileBytesA);
        bulk_g2s(smem_u32(s), P.a_img + ((size_t)tile_m * P.nkb + kb) * (2 * kTileBytesA), 2 * kTileBytesA, bar(st));
        for (int t = 0; t < b_tiles; ++t) {
          const uint8_t* src = P.b_img + ((size_t)t * P.nkb + kb) * (2 * kTileBytesA);
          bulk_g2s(smem_u32(s + (2 + t) * kTileBytesA), src, kTileBytesA, bar(st));                // hi -> B hi region
          bulk_g2s(smem_u32(s + (4 + t) * kTileBytesA), src + kTileBytesA, kTileBytesA, bar(st));  // lo -> B lo region
        }
      }
      __syncwarp();
      if (++st == kGemmStages) { st = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    int st = 0;
    uint32_t ph = 0;
    const uint32_t idesc = make_idesc(128, (uint32_t)P.n_mma);
    const uint64_t desc0 = make_desc(smem_u32(smem));
    for (int kb = kb0; kb < kb1; ++kb) {
      mbar_wait(bar(st), ph);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t a_hi = desc0 + (uint64_t)(st * (kStageBytes >> 4)), a_lo = a_hi + (kTileBytesA >> 4);
        const uint64_t b_hi = a_hi + 2 * (kTileBytesA >> 4), b_lo = a_hi + 4 * (kTileBytesA >> 4);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_f16(tmem_base, a_hi + 2 * ks, b_hi + 2 * ks, idesc, (kb > kb0 || ks > 0) ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_f16(tmem_base, a_lo + 2 * ks, b_hi + 2 * ks, idesc, 1u);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_f16(tmem_base, a_hi + 2 * ks, b_lo + 2 * ks, idesc, 1u);
        umma_commit(bar(kGemmStages + st));
        if (kb == kb1 - 1) umma_commit(bar(2 * kGemmStages));
      }
      __syncwarp();
      if (++st == kGemmStages) { st = 0; ph ^= 1; }
    }
  } else if (kb1 > kb0) {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const long long grow = (long long)tile_m * 128 + row;
    mbar_wait(bar(2 * kGemmStages), 0);
    tc_fence_after();
    const float scale = P.out_scale_dev ? __ldg(P.out_scale_dev) : 1.0f;
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int c0 = 0; c0 < P.n_mma; c0 += 32) {
      uint32_t v[32];
      tmem_ld32_async(t_row + c0, v);  // columns beyond n_mma hold stale data and are never stored
      tmem_wait(v);
      if (grow < P.M) {
        if (P.partial) {
          float* dst = P.partial + ((size_t)split * gridDim.x * 128 + (size_t)grow) * 256 + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            if (c0 + j < P.N)
              *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        } else {
          float* dst = P.C + grow * P.ldc + c0;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (c0 + j < P.N) {
              float y = __uint_as_float(v[j]) * scale;
              if (P.bias) y += __ldg(P.bias + c0 + j);
              dst[j] = fmaxf(y, P.slope * y);
            }
          }
        }
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 256);
  }
}

// C[row, col] = act(scale * sum_split partial[split][row][col] + bias[col]); splits are summed in ascending order (deterministic)
__global__ void k_splitk_reduce(const float* __restrict__ partial, int splits, long long rows_pad, int M, int N, float* __restrict__ C,
                                long long ldc, const float* __restrict__ bias, const float* __restrict__ out_scale_dev, float slope) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)M * N) return;
  const long long row = i / N;
  const int col = (int)(i % N);
  float acc = 0.f;
  for (int s = 0; s < splits; ++s) acc += partial[((size_t)s * rows_pad + row) * 256 + col];
  float y = acc * (out_scale_dev ? __ldg(out_scale_dev) : 1.0f);
  if (bias) y += __ldg(bias + col);
  C[row * ldc + col] = fmaxf(y, slope * y);
}

// Operand image of X' (image rows r', K index k'):  X'[r', k'] = src[r', k'] (transpose = 0) or src[k', r'] (transpose = 1),
// times (mask_src > 0 ? 1 : slope) at the same source position if mask_src is given, times *scale_dev.  K padding and the rows up
// to the next multiple of 128 are zero-filled.
__global__ void k_pack_rows(const float* __restrict__ src, long long rows, int cols, long long ld, int transpose,
                            const float* __restrict__ mask_src, float slope, const float* __restrict__ scale_dev,
                            long long img_rows, long long img_k, int nkb, uint8_t* __restrict__ img) {
  const long long rows_pad = (img_rows + 127) & ~127ll;  // rows up to the tile boundary are zero-filled (wgrad reduces over rows)
  const long long n_chunks = rows_pad * (nkb * 8);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_chunks) return;
  // consecutive threads: consecutive image rows when transposed (coalesced source reads), consecutive chunks otherwise
  long long r;
  int c;
  if (transpose) { r = i % rows_pad; c = (int)(i / rows_pad); }
  else { r = i / (nkb * 8); c = (int)(i % (nkb * 8)); }
  const float scale = scale_dev ? __ldg(scale_dev) : 1.0f;
  float y[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const long long k = (long long)c * 8 + j;
    float v = 0.f;
    if (k < img_k && r < img_rows) {
      const long long off = transpose ? k * ld + r : r * ld + k;
      v = __ldg(src + off);
      if (mask_src) v *= (__ldg(mask_src + off) > 0.f) ? 1.0f : slope;
      v *= scale;
    }
    y[j] = v;
  }
  uint4 hi, lo;
  split8(y, hi, lo);
  uint8_t* p = img + ((size_t)(r >> 7) * nkb + (c >> 3)) * (2 * kTileBytesA) + swz((int)(r & 127), c & 7);
  *reinterpret_cast<uint4*>(p) = hi;
  *reinterpret_cast<uint4*>(p + kTileBytesA) = lo;
}

}  // namespace tc
}  // namespace npcd

using namespace npcd;

extern "C" int npcd_tc_image_bytes(long long rows, long long k, size_t* bytes) {
  NPCD_CHECK_ARG(bytes && rows >= 0 && k >= 0, "bad arguments");
  *bytes = (size_t)((rows + 127) / 128) * (size_t)((k + 63) / 64) * (2 * tc::kTileBytesA);
  return 0;
}

extern "C" int npcd_tc_pack_rows(const float* src, long long rows, int cols, long long ld, int transpose, const float* mask_src,
                                 float slope, const float* scale_dev, void* image, void* stream) {
  NPCD_CHECK_ARG(src && image && rows > 0 && cols > 0 && ld >= cols, "bad arguments");
  const long long img_rows = transpose ? cols : rows, img_k = transpose ? rows : cols;
  const int nkb = (int)((img_k + 63) / 64);
  const long long n_chunks = ((img_rows + 127) & ~127ll) * (nkb * 8);
  tc::k_pack_rows<<<(unsigned)((n_chunks + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, rows, cols, ld, transpose, mask_src, slope,
                                                                                       scale_dev, img_rows, img_k, nkb,
                                                                                       (uint8_t*)image);
  return check_launch("npcd_tc_pack_rows");
}

extern "C" int npcd_tc_gemm_workspace_bytes(int M, int split_k, size_t* bytes) {
  NPCD_CHECK_ARG(bytes && M > 0 && split_k >= 1, "bad arguments");
  *bytes = split_k > 1 ? (size_t)split_k * (size_t)((M + 127) / 128) * 128 * 256 * sizeof(float) : 0;
  return 0;
}

extern "C" int npcd_tc_gemm(const void* a_image, const void* b_image, int M, int N, long long K, float* C, long long ldc,
                            const float* bias, const float* out_scale_dev, float act_slope, int split_k, void* workspace,
                            size_t workspace_bytes, void* stream) {
  NPCD_CHECK_ARG(a_image && b_image && C, "null pointer");
  NPCD_CHECK_ARG(M > 0 && N > 0 && N <= 256 && K > 0 && ldc >= N && split_k >= 1, "bad sizes (N <= 256)");
  const int nkb = (int)((K + 63) / 64);
  if (split_k > nkb) split_k = nkb;
  const int kb_per_split = (nkb + split_k - 1) / split_k;
  split_k = (nkb + kb_per_split - 1) / kb_per_split;  // no empty splits
  const int tiles_m = (M + 127) / 128;
  size_t need = 0;
  npcd_tc_gemm_workspace_bytes(M, split_k, &need);
  NPCD_CHECK_ARG(split_k == 1 || (workspace && workspace_bytes >= need), "split-K workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaFuncSetAttribute(tc::k_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kGemmSmem);
  if (e != cudaSuccess) {
    set_error("npcd_tc_gemm: cannot opt in to %d bytes of shared memory: %s", tc::kGemmSmem, cudaGetErrorString(e));
    return 2;
  }
  tc::GemmParams P{};
  P.a_img = (const uint8_t*)a_image; P.b_img = (const uint8_t*)b_image;
  P.M = M; P.N = N; P.nkb = nkb; P.kb_per_split = kb_per_split; P.n_mma = (N + 15) / 16 * 16;
  P.C = C; P.ldc = ldc; P.bias = bias; P.out_scale_dev = out_scale_dev; P.slope = act_slope;
  P.partial = split_k > 1 ? (float*)workspace : nullptr;
  tc::k_gemm_tc<<<dim3(tiles_m, split_k), tc::kGemmThreads, tc::kGemmSmem, st>>>(P);
  int rc = check_launch("npcd_tc_gemm");
  if (rc || split_k == 1) return rc;
  const long long n = (long long)M * N;
  tc::k_splitk_reduce<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const float*)workspace, split_k, (long long)tiles_m * 128, M, N, C,
                                                                 ldc, bias, out_scale_dev, act_slope);
  return check_launch("npcd_tc_gemm(reduce)");
}
