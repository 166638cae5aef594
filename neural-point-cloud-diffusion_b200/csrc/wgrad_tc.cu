// Weight-gradient GEMMs on the tcgen05 tensor cores:  C[m, j] = scale * sum_rows A[row, m] * B[row, perm(j)]  and, from the same
// pass, the bias gradient  bias[m] = scale * sum_rows A[row, m]
// (row B* of SURVEY.md section 8: dW = dY^T . X and db = sum dY of every nn.Linear in fields/aggregators/mlp.py:34,84 and
// fields/mlp.py:38-72, which the reference obtains from autograd as fp32 cuBLAS GEMMs / reductions).
//
// Both operands are ROW-major operand images (rows x 64-column K-blocks, fp16 hi/lo, SWIZZLE_128B) -- exactly what the forward /
// dgrad kernels read as K-major operands and what they stash.  Here the reduction runs over ROWS, so the very same bytes are
// consumed as MN-major operands (instruction-descriptor bits 15/16): a 128-byte image row is 64 contiguous M (or N) elements, an
// 8-row swizzle atom is 8 consecutive K steps.  No transposed copies are ever built.
// The bias gradient rides along as two extra N = 16 MMAs per K step against a tile of ones (any layout of ones is ones).
//
// Grid = (M halves of 128, row splits, groups): several independent problems (the layers of one backward) share a launch.  Each
// CTA streams 64-row slabs of its row range through a 2-stage bulk-copy ring (A: 2 blocks x hi/lo x 8 KB, B: up to 4 blocks x
// hi/lo x 8 KB), accumulates hi*hi + lo*hi + hi*lo in TMEM and writes one fp32 partial; k_wgrad_reduce sums the partials in a
// fixed order (deterministic).  Rows of the A image beyond the row count must be zero (npcd_tc_pack_rows and the fused backward
// kernels guarantee it).
#include "npcd_b200.h"
#include "tc_ptx.cuh"

namespace npcd {
namespace tc {

constexpr int kWgThreads = 192;
constexpr int kWgStages = 2;
constexpr int kSlabBytes = 8192;                   // 64 rows of one K-block half (hi or lo)
constexpr int kWgStageBytes = 12 * kSlabBytes;     // A: 4 slabs, B: 8 slabs = 96 KB
constexpr int kWgOnesBytes = 2048;                 // 16 K rows x 128 B of fp16 1.0
constexpr int kWgSmem = kWgStages * kWgStageBytes + kWgOnesBytes + 1024;
constexpr int kWgMaxGroups = NPCD_WGRAD_MAX_GROUPS;

struct WgradGroup {
  const uint8_t* a_img;
  const uint8_t* b_img;
  int a_nkb, b_nkb;
  long long rows;
  const long long* rows_dev;  // optional device row count (overrides rows, clamped to it)
  float* partial;             // [gridDim.y][m_pad][256 + 1 bias column, padded to 260]
  int m_pad;                  // halves * 128
  // reduce
  float* C;
  long long ldc;
  int a_cols, n_out;
  const int* perm;
  const float* scale_dev;
  const float* bias_scale_dev;
  float* bias_out;
  int accumulate;
};
struct WgradParams {
  WgradGroup g[kWgMaxGroups];
  int flags;  // bit 0: swap LBO / SBO (descriptor probe)
};
constexpr int kWgRow = 260;  // floats per partial row: 256 C columns + the bias column + padding

// MN-major SWIZZLE_128B descriptor: 64-element (128 B) MN blocks `lbo` bytes apart, 8-row K groups `sbo` bytes apart
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(kWgThreads, 1) k_wgrad_tc(const __grid_constant__ WgradParams P) {
  const WgradGroup& G = P.g[blockIdx.z];
  const int h = blockIdx.x;
  if (h * 128 >= G.m_pad) return;  // this group has fewer M halves than the widest one in the launch
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* ones = smem + kWgStages * kWgStageBytes;
  uint8_t* misc = ones + kWgOnesBytes;
  const uint32_t bars = smem_u32(misc);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 64);
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };  // full[2], empty[2], acc_ready
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kWgStages; ++i) { mbar_init(bar(i), 1); mbar_init(bar(kWgStages + i), 1); }
    mbar_init(bar(2 * kWgStages), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < kWgOnesBytes / 4; i += kWgThreads) reinterpret_cast<uint32_t*>(ones)[i] = 0x3C003C00u;  // (1.0h, 1.0h)
  fence_proxy_async();
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const long long rows = G.rows_dev ? min(*G.rows_dev, G.rows) : G.rows;
  const long long n_slabs = (rows + 63) >> 6;
  const long long per = (n_slabs + gridDim.y - 1) / gridDim.y;
  const long long s0 = min(n_slabs, (long long)blockIdx.y * per), s1 = min(n_slabs, s0 + per);
  const int n_mma = G.b_nkb * 64;

  if (warp == 0) {
    int st = 0;
    uint32_t ph = 0;
    for (long long s = s0; s < s1; ++s) {
      mbar_wait(bar(kWgStages + st), ph ^ 1);
      if (elect_one()) {
        uint8_t* dst = smem + st * kWgStageBytes;
        mbar_expect_tx(bar(st), (uint32_t)(4 + 2 * G.b_nkb) * kSlabBytes);
        const long long tile = s >> 1;
        const size_t half_off = (size_t)(s & 1) * kSlabBytes;
        for (int j = 0; j < 2; ++j) {
          const int kb = min(2 * h + j, G.a_nkb - 1);  // a missing second block re-reads the last one (its output rows are dropped)
          const uint8_t* src = G.a_img + ((size_t)tile * G.a_nkb + kb) * (2 * kTileBytesA) + half_off;
          bulk_g2s(smem_u32(dst + j * kSlabBytes), src, kSlabBytes, bar(st));                      // hi
          bulk_g2s(smem_u32(dst + (2 + j) * kSlabBytes), src + kTileBytesA, kSlabBytes, bar(st));  // lo
        }
        for (int j = 0; j < G.b_nkb; ++j) {
          const uint8_t* src = G.b_img + ((size_t)tile * G.b_nkb + j) * (2 * kTileBytesA) + half_off;
          bulk_g2s(smem_u32(dst + (4 + j) * kSlabBytes), src, kSlabBytes, bar(st));
          bulk_g2s(smem_u32(dst + (4 + G.b_nkb + j) * kSlabBytes), src + kTileBytesA, kSlabBytes, bar(st));
        }
      }
      __syncwarp();
      if (++st == kWgStages) { st = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    int st = 0;
    uint32_t ph = 0;
    const uint32_t mn = (1u << 15) | (1u << 16);  // A and B MN-major
    const uint32_t idesc = make_idesc(128, (uint32_t)n_mma) | mn, idesc_ones = make_idesc(128, 16) | mn;
    const uint32_t lbo = (P.flags & 1) ? 1024u : (uint32_t)kSlabBytes, sbo = (P.flags & 1) ? (uint32_t)kSlabBytes : 1024u;
    const uint64_t b_ones = make_desc_mn(smem_u32(ones), lbo, sbo);
    for (long long s = s0; s < s1; ++s) {
      mbar_wait(bar(st), ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t base = smem_u32(smem + st * kWgStageBytes);
        const uint64_t a_hi = make_desc_mn(base, lbo, sbo), a_lo = make_desc_mn(base + 2 * kSlabBytes, lbo, sbo);
        const uint64_t b_hi = make_desc_mn(base + 4 * kSlabBytes, lbo, sbo);
        const uint64_t b_lo = make_desc_mn(base + (4 + G.b_nkb) * kSlabBytes, lbo, sbo);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {  // 16 rows = two 8-row atoms = 2048 B per K step
          const uint64_t o = (uint64_t)(ks * (2048 >> 4));
          const uint32_t acc = (s > s0 || ks > 0) ? 1u : 0u;
          umma_f16(tmem_base, a_hi + o, b_hi + o, idesc, acc);
          umma_f16(tmem_base, a_lo + o, b_hi + o, idesc, 1u);
          umma_f16(tmem_base, a_hi + o, b_lo + o, idesc, 1u);
          umma_f16(tmem_base + 256u, a_hi + o, b_ones, idesc_ones, acc);  // bias gradient: column sums of A
          umma_f16(tmem_base + 256u, a_lo + o, b_ones, idesc_ones, 1u);
        }
        umma_commit(bar(kWgStages + st));
        if (s == s1 - 1) umma_commit(bar(2 * kWgStages));
      }
      __syncwarp();
      if (++st == kWgStages) { st = 0; ph ^= 1; }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;  // output row within this M half
    float* dst = G.partial + ((size_t)blockIdx.y * G.m_pad + (size_t)h * 128 + row) * kWgRow;
    if (s1 > s0) {
      mbar_wait(bar(2 * kWgStages), 0);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
      for (int c0 = 0; c0 < n_mma; c0 += 32) {
        uint32_t v[32];
        tmem_ld32_async(t_row + c0, v);
        tmem_wait(v);
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                 __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
      }
      uint32_t v[32];
      tmem_ld32_async(t_row + 256, v);  // 16 identical bias columns (+ 16 unused ones)
      tmem_wait(v);
      dst[256] = __uint_as_float(v[0]);
      tc_fence_before();
    } else {
      for (int c0 = 0; c0 < n_mma; c0 += 4) *reinterpret_cast<float4*>(dst + c0) = make_float4(0.f, 0.f, 0.f, 0.f);
      dst[256] = 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

// C[m, j] = scale * sum_y partial[y][m][perm ? perm[j] : j]   (+= if accumulate); bias_out[m] likewise from the bias column;
// splits summed in ascending order.  grid = (blocks over (m, j), 1, groups)
__global__ void k_wgrad_reduce(const __grid_constant__ WgradParams P, int splits) {
  const WgradGroup& G = P.g[blockIdx.z];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_c = G.a_cols * G.n_out;
  const float scale = G.scale_dev ? __ldg(G.scale_dev) : 1.0f;
  if (i < n_c) {
    const int m = i / G.n_out, j = i % G.n_out;
    const int col = G.perm ? G.perm[j] : j;
    float acc = 0.f;
    if (col >= 0)
      for (int y = 0; y < splits; ++y) acc += G.partial[((size_t)y * G.m_pad + m) * kWgRow + col];
    acc *= scale;
    float* out = G.C + (size_t)m * G.ldc + j;
    *out = G.accumulate ? *out + acc : acc;
  } else if (G.bias_out && i < n_c + G.a_cols) {
    const int m = i - n_c;
    float acc = 0.f;
    for (int y = 0; y < splits; ++y) acc += G.partial[((size_t)y * G.m_pad + m) * kWgRow + 256];
    acc *= G.bias_scale_dev ? __ldg(G.bias_scale_dev) : scale;
    G.bias_out[m] = G.accumulate ? G.bias_out[m] + acc : acc;
  }
}

// column sums of an operand image: out[c] = scale * sum_rows (hi + lo)[row, c]; one block per (K-block, row split); fixed-order
// two-level sum (deterministic)
__global__ void k_image_colsum(const uint8_t* __restrict__ img, int nkb, long long rows, const long long* __restrict__ rows_dev,
                               float* __restrict__ partial /* [gridDim.y][nkb*64] */) {
  const long long n = rows_dev ? min(*rows_dev, rows) : rows;
  const int kb = blockIdx.x;
  const long long per = (n + gridDim.y - 1) / gridDim.y;
  const long long r0 = min(n, (long long)blockIdx.y * per), r1 = min(n, r0 + per);
  // thread t: 16-byte chunk c16 = t & 7 (8 columns), row lane t >> 3 (32 row lanes)
  const int c16 = threadIdx.x & 7, rl = threadIdx.x >> 3;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (long long r = r0 + rl; r < r1; r += 32) {
    const uint8_t* p = img + ((size_t)(r >> 7) * nkb + kb) * (2 * kTileBytesA) + swz((int)(r & 127), c16);
    const uint4 hi = __ldg(reinterpret_cast<const uint4*>(p)), lo = __ldg(reinterpret_cast<const uint4*>(p + kTileBytesA));
    const uint32_t hh[4] = {hi.x, hi.y, hi.z, hi.w}, ll[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hh[j]));
      const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&ll[j]));
      acc[2 * j] += a.x + b.x;
      acc[2 * j + 1] += a.y + b.y;
    }
  }
  __shared__ float red[32][65];
#pragma unroll
  for (int j = 0; j < 8; ++j) red[rl][c16 * 8 + j] = acc[j];
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
    for (int i = 0; i < 32; ++i) s += red[i][threadIdx.x];
    partial[(size_t)blockIdx.y * (nkb * 64) + kb * 64 + threadIdx.x] = s;
  }
}

__global__ void k_colsum_reduce(const float* __restrict__ partial, int splits, int width, int n_out, const int* __restrict__ perm,
                                const float* __restrict__ scale_dev, float* __restrict__ out, int accumulate) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_out) return;
  const int col = perm ? perm[j] : j;
  float acc = 0.f;
  if (col >= 0)
    for (int y = 0; y < splits; ++y) acc += partial[(size_t)y * width + col];
  acc *= scale_dev ? __ldg(scale_dev) : 1.0f;
  out[j] = accumulate ? out[j] + acc : acc;
}

}  // namespace tc
}  // namespace npcd

using namespace npcd;

extern "C" int npcd_tc_wgrad_workspace_bytes(int a_cols, int row_splits, size_t* bytes) {
  NPCD_CHECK_ARG(bytes && a_cols > 0 && row_splits >= 1, "bad arguments");
  const int halves = (a_cols + 127) / 128;
  *bytes = (size_t)row_splits * halves * 128 * tc::kWgRow * sizeof(float);
  return 0;
}

extern "C" int npcd_tc_wgrad_grouped(const npcd_wgrad_problem* problems, int n_problems, int row_splits, void* workspace,
                                     size_t workspace_bytes, int flags, void* stream) {
  NPCD_CHECK_ARG(problems && workspace, "null pointer");
  NPCD_CHECK_ARG(n_problems >= 1 && n_problems <= tc::kWgMaxGroups && row_splits >= 1 && row_splits <= 65535, "bad group count / splits");
  static thread_local tc::WgradParams P;
  memset(&P, 0, sizeof(P));
  P.flags = flags;
  size_t off = 0;
  int max_halves = 1, max_n = 1;
  for (int i = 0; i < n_problems; ++i) {
    const npcd_wgrad_problem& q = problems[i];
    NPCD_CHECK_ARG(q.a_image && q.b_image && q.C, "null pointer in a problem");
    NPCD_CHECK_ARG(q.a_cols > 0 && q.b_cols > 0 && q.b_cols <= 256 && q.rows >= 0 && q.n_out > 0 && q.ldc >= q.n_out,
                   "bad sizes (b_cols <= 256)");
    NPCD_CHECK_ARG(q.col_perm || q.n_out <= q.b_cols, "n_out > b_cols without a column map");
    tc::WgradGroup& G = P.g[i];
    const int halves = (q.a_cols + 127) / 128;
    G.a_img = (const uint8_t*)q.a_image; G.b_img = (const uint8_t*)q.b_image;
    G.a_nkb = (q.a_cols + 63) / 64; G.b_nkb = (q.b_cols + 63) / 64;
    G.rows = q.rows; G.rows_dev = q.rows_dev;
    G.partial = (float*)((uint8_t*)workspace + off);
    G.m_pad = halves * 128;
    G.C = q.C; G.ldc = q.ldc; G.a_cols = q.a_cols; G.n_out = q.n_out; G.perm = q.col_perm; G.scale_dev = q.out_scale_dev;
    G.bias_out = q.bias_out; G.bias_scale_dev = q.bias_scale_dev; G.accumulate = q.accumulate;
    size_t need = 0;
    npcd_tc_wgrad_workspace_bytes(q.a_cols, row_splits, &need);
    off += need;
    max_halves = halves > max_halves ? halves : max_halves;
    const int n = q.a_cols * q.n_out + q.a_cols;
    max_n = n > max_n ? n : max_n;
  }
  NPCD_CHECK_ARG(workspace_bytes >= off, "workspace too small (sum of npcd_tc_wgrad_workspace_bytes over the problems)");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaFuncSetAttribute(tc::k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kWgSmem);
  if (e != cudaSuccess) {
    set_error("npcd_tc_wgrad: cannot opt in to %d bytes of shared memory: %s", tc::kWgSmem, cudaGetErrorString(e));
    return 2;
  }
  tc::k_wgrad_tc<<<dim3(max_halves, row_splits, n_problems), tc::kWgThreads, tc::kWgSmem, st>>>(P);
  int rc = check_launch("npcd_tc_wgrad");
  if (rc) return rc;
  tc::k_wgrad_reduce<<<dim3((max_n + 255) / 256, 1, n_problems), 256, 0, st>>>(P, row_splits);
  return check_launch("npcd_tc_wgrad(reduce)");
}

extern "C" int npcd_tc_wgrad(const void* a_image, int a_cols, const void* b_image, int b_cols, long long rows,
                             const long long* rows_dev, float* C, long long ldc, int n_out, const int* col_perm,
                             const float* out_scale_dev, int accumulate, int row_splits, void* workspace, size_t workspace_bytes,
                             int flags, void* stream) {
  npcd_wgrad_problem q;
  memset(&q, 0, sizeof(q));
  q.a_image = a_image; q.a_cols = a_cols; q.b_image = b_image; q.b_cols = b_cols; q.rows = rows; q.rows_dev = rows_dev;
  q.C = C; q.ldc = ldc; q.n_out = n_out; q.col_perm = col_perm; q.out_scale_dev = out_scale_dev; q.bias_out = nullptr;
  q.accumulate = accumulate;
  return npcd_tc_wgrad_grouped(&q, 1, row_splits, workspace, workspace_bytes, flags, stream);
}

extern "C" int npcd_tc_image_colsum(const void* image, int cols, long long rows, const long long* rows_dev, float* out, int n_out,
                                    const int* col_perm, const float* out_scale_dev, int accumulate, int row_splits,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  NPCD_CHECK_ARG(image && out && workspace, "null pointer");
  NPCD_CHECK_ARG(cols > 0 && rows >= 0 && n_out > 0 && row_splits >= 1 && row_splits <= 65535, "bad sizes");
  NPCD_CHECK_ARG(col_perm || n_out <= cols, "n_out > cols without a column map");
  const int nkb = (cols + 63) / 64;
  NPCD_CHECK_ARG(workspace_bytes >= (size_t)row_splits * nkb * 64 * sizeof(float), "workspace too small (row_splits * ceil64(cols) floats)");
  cudaStream_t st = (cudaStream_t)stream;
  tc::k_image_colsum<<<dim3(nkb, row_splits), 256, 0, st>>>((const uint8_t*)image, nkb, rows, rows_dev, (float*)workspace);
  int rc = check_launch("npcd_tc_image_colsum");
  if (rc) return rc;
  tc::k_colsum_reduce<<<(n_out + 255) / 256, 256, 0, st>>>((const float*)workspace, row_splits, nkb * 64, n_out, col_perm,
                                                           out_scale_dev, out, accumulate);
  return check_launch("npcd_tc_image_colsum(reduce)");
}
