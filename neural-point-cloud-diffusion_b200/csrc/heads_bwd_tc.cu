// Fused backward of the per-sample heads (row B* of SURVEY.md section 8: the autograd of `local_field.8`, `shape_net`, `channel_net`
// and the output activations, fields/mlp.py:38-72 + fields/field.py:126-141), sm_100a only.
//
// Input: dL/d(r, g, b, sigma) per kept sample, the forward outputs (r, g, b, sigma) and the sign masks stashed by the training
// forward (k_field_tc<MODE_HEADS_TRAIN>).  Per 128-sample tile, all values times one power-of-two scale (scale_dev[0]):
//   g_rgb = dL/drgb * rgb (1 - rgb),  g_s = dL/dsigma * (1 - exp(-sigma))            (sigmoid' and softplus' from the outputs)
//   dP_c3 = (g_rgb . W_cout) * lrelu'(C4)                                             prologue, 3 -> 256 outer product
//   dP_c2 = (dP_c3 . W_c3) * lrelu'(C3);  dP_c1 = (dP_c2 . W_c2) * lrelu'(C2);  dP_c0 = (dP_c1 . W_c1) * lrelu'(C1)
//   dP_s  = (g_s * w_sout) * lrelu'(H)                                                second prologue, 1 -> 256
//   dF    = dP_c0 . W_c0 + dP_s . W_s0                                                two GEMMs into ONE TMEM accumulator
//   dL/dagg = dF . W_4                                                                 -> fp32 [S,256] for the pair backward
// Every dP operand (and dF) is bulk-copied to HBM as an operand image: with the stashed layer inputs they feed the weight-gradient
// GEMMs (wgrad_tc.cu) and the bias column sums; (g_rgb, g_s) leaves as fp32 [S,4] for the two narrow output layers.
// Engine and warp roles as in pair_bwd_tc.cu.
#include "npcd_b200.h"
#include "tc_ptx.cuh"

namespace npcd {
namespace tc {

constexpr int kHbThreads = 352;
constexpr int kHbTileW = 256 * 128;
constexpr int kHbStages = 3;
constexpr int kHbSmemA = 4 * 2 * kTileBytesA;
constexpr int kHbSmem = kHbSmemA + kHbStages * kHbTileW + 1024;
constexpr int hWFull = 0, hWEmpty = 3, hARdy = 6, hAFree = 10, hAccRdy = 14, hAccFree = 16, hStash = 18;

struct HeadsBwdParams {
  const uint8_t* w[6];  // packed transposes in order of use: W_c3^T, W_c2^T, W_c1^T, W_c0^T, W_s0^T, W_4^T
  float inv_scale[6];   // inv_scale[3] == inv_scale[4] (the two GEMMs share an accumulator)
  float chan_out_w[3][256];
  float shape_out_w[256];
  const float4* d_rgbs;
  const float4* rgbs;
  const uint32_t* mask[5];  // H, C1, C2, C3, C4
  const long long* n_samples_dev;
  long long capacity;
  const float* scale_dev;
  uint8_t* dp[6];  // dP_c3, dP_c2, dP_c1, dP_c0, dP_s, dF
  float4* g4;
  float* d_agg;
  int* error_flag;
};

__device__ __forceinline__ void hb_chunk_store(const float (&v)[32], float mul, uint32_t bits, int c0, uint8_t* sA, uint32_t rowbase,
                                               int x7) {
  uint8_t* kb_base = sA + (c0 >> 6) * (2 * kTileBytesA) + rowbase;
  const int c16_0 = (c0 & 63) >> 3;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float y[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = v[g * 8 + j] * mul;
      t = ((bits >> (g * 8 + j)) & 1u) ? t : 0.01f * t;
      y[j] = fminf(fmaxf(t, -60000.f), 60000.f);
    }
    uint4 hi, lo;
    split8(y, hi, lo);
    uint8_t* p = kb_base + (((c16_0 + g) ^ x7) << 4);
    *reinterpret_cast<uint4*>(p) = hi;
    *reinterpret_cast<uint4*>(p + kTileBytesA) = lo;
  }
}

__global__ void __launch_bounds__(kHbThreads, 1) k_heads_bwd_tc(const __grid_constant__ HeadsBwdParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sW = smem + kHbSmemA;
  uint8_t* misc = smem + kHbSmemA + kHbStages * kHbTileW;
  const uint32_t bars = smem_u32(misc);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 256);
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if ((smem_u32(smem) & 1023u) != 0u) {
    if (threadIdx.x == 0 && P.error_flag) atomicExch(P.error_flag, 1);
    return;
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < kHbStages; ++i) { mbar_init(bar(hWFull + i), 1); mbar_init(bar(hWEmpty + i), 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(bar(hARdy + i), 8); mbar_init(bar(hAFree + i), 1); mbar_init(bar(hStash + i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar(hAccRdy + i), 1); mbar_init(bar(hAccFree + i), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const long long S = min(*P.n_samples_dev, P.capacity);
  const int n_tiles = (int)((S + 127) / 128);

  if (warp == 0) {
    // ================================================= weight producer ==================================================
    int st = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int L = 0; L < 6; ++L) {
        for (int t = 0; t < 8; ++t) {
          mbar_wait(bar(hWEmpty + st), ph ^ 1);
          if (elect_one()) {
            mbar_expect_tx(bar(hWFull + st), kHbTileW);
            bulk_g2s(smem_u32(sW + st * kHbTileW), P.w[L] + (size_t)t * kHbTileW, kHbTileW, bar(hWFull + st));
          }
          __syncwarp();
          if (++st == kHbStages) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer ===================================================
    // accumulator of GEMM L: 0, 1, 0, 1, 1 (accumulates onto GEMM 3), 0
    int st = 0;
    uint32_t ph_w = 0, ph_ar = 0, ph_af = 0;
    const uint64_t desc_a0 = make_desc(smem_u32(sA));
    const uint64_t desc_w0 = make_desc(smem_u32(sW));
    constexpr uint32_t idesc = make_idesc(128, 256);
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int L = 0; L < 6; ++L) {
        const uint32_t ab = (L < 4) ? (uint32_t)(L & 1) : (L == 4 ? 1u : 0u);
        const uint32_t d_tmem = tmem_base + ab * 256u;
        if (L != 4) {
          mbar_wait(bar(hAccFree + ab), ((ph_af >> ab) & 1u) ^ 1u);
          ph_af ^= 1u << ab;
        }
        for (int kb = 0; kb < 4; ++kb) {
          mbar_wait(bar(hARdy + kb), (ph_ar >> kb) & 1u);
          ph_ar ^= 1u << kb;
          const uint64_t a_hi = desc_a0 + (uint64_t)(kb * 2 * (kTileBytesA >> 4)), a_lo = a_hi + (kTileBytesA >> 4);
          mbar_wait(bar(hWFull + st), ph_w);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t b = desc_w0 + (uint64_t)(st * (kHbTileW >> 4));
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_f16(d_tmem, a_hi + 2 * ks, b + 2 * ks, idesc, (L == 4 || (kb | ks) != 0) ? 1u : 0u);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_f16(d_tmem, a_lo + 2 * ks, b + 2 * ks, idesc, 1u);
            umma_commit(bar(hWEmpty + st));
          }
          __syncwarp();
          if (++st == kHbStages) { st = 0; ph_w ^= 1; }
          mbar_wait(bar(hWFull + st), ph_w);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t b = desc_w0 + (uint64_t)(st * (kHbTileW >> 4));
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_f16(d_tmem, a_hi + 2 * ks, b + 2 * ks, idesc, 1u);
            umma_commit(bar(hWEmpty + st));
            if (L == 3 || L == 5) umma_commit(bar(hAFree + kb));  // dP_c0 / dF consumed: the K-block may be rewritten
            if (kb == 3 && L != 3) umma_commit(bar(hAccRdy + ab));
          }
          __syncwarp();
          if (++st == kHbStages) { st = 0; ph_w ^= 1; }
        }
      }
    }
  } else if (warp == 10) {
    // ============================================ dP stash (operands of the weight gradients) ===========================
    uint32_t ph_ar = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int L = 0; L < 6; ++L) {
        for (int kb = 0; kb < 4; ++kb) {
          mbar_wait(bar(hARdy + kb), (ph_ar >> kb) & 1u);
          ph_ar ^= 1u << kb;
          if (elect_one()) {
            bulk_s2g(P.dp[L] + ((size_t)tile * 4 + kb) * (2 * kTileBytesA), smem_u32(sA + kb * 2 * kTileBytesA), 2 * kTileBytesA);
            bulk_commit();
            bulk_wait_read0();
            mbar_arrive(bar(hStash + kb));
          }
          __syncwarp();
        }
      }
    }
    if (elect_one()) bulk_wait_all0();
    __syncwarp();
  } else {
    // ============================================ prologue / epilogue threads ===========================================
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int x7 = row & 7;
    const uint32_t rowbase = (uint32_t)((row >> 3) * 1024 + x7 * 128);
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t ph_acc = 0, ph_sd = 0, sd_pending = 0;
    const float s_fwd = __ldg(P.scale_dev), s_inv = __ldg(P.scale_dev + 1);

    auto publish = [&](int kb) {
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(hARdy + kb));
      sd_pending |= 1u << kb;
    };
    auto stash_wait = [&](int kb) {
      if ((sd_pending >> kb) & 1u) {
        mbar_wait(bar(hStash + kb), (ph_sd >> kb) & 1u);
        ph_sd ^= 1u << kb;
        sd_pending &= ~(1u << kb);
      }
    };
    auto wait_acc = [&](uint32_t ab) {
      mbar_wait(bar(hAccRdy + ab), (ph_acc >> ab) & 1u);
      ph_acc ^= 1u << ab;
      tc_fence_after();
    };
    auto release_acc = [&](uint32_t ab) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(hAccFree + ab));
    };
    // dgrad epilogue: accumulator ab -> (acc * inv) * lrelu'(mask) -> next A operand, in place
    auto epilogue = [&](uint32_t ab, float inv, const uint32_t* mask, size_t r) {
      wait_acc(ab);
      const uint32_t t_acc = t_row + ab * 256u;
#pragma unroll 1
      for (int i = 0; i < 4; ++i) {
        const int ch = 2 * i + half;
        uint32_t v[32];
        tmem_ld32_async(t_acc + ch * 32, v);
        const uint32_t bits = mask ? __ldg(mask + r * 8 + ch) : 0xffffffffu;
        tmem_wait(v);
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        stash_wait(i);
        hb_chunk_store(f, inv, bits, ch * 32, sA, rowbase, x7);
        if (i == 3) release_acc(ab);
        publish(i);
      }
    };

    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const long long s = (long long)tile * 128 + row;
      const size_t r = (size_t)s;
      const bool valid = s < S;
      float gr[3] = {0.f, 0.f, 0.f}, gs = 0.f;
      if (valid) {
        const float4 d = __ldg(P.d_rgbs + s), o = __ldg(P.rgbs + s);
        gr[0] = d.x * o.x * (1.0f - o.x);
        gr[1] = d.y * o.y * (1.0f - o.y);
        gr[2] = d.z * o.z * (1.0f - o.z);
        gs = d.w * (1.0f - expf(-o.w));  // softplus'(x) = sigmoid(x) = 1 - exp(-softplus(x))
        if (half == 0) P.g4[s] = make_float4(gr[0], gr[1], gr[2], gs);
      }
      gr[0] *= s_fwd; gr[1] *= s_fwd; gr[2] *= s_fwd; gs *= s_fwd;
      // ---- first operand dP_c3 = (g_rgb . W_cout) * lrelu'(C4)
#pragma unroll 1
      for (int i = 0; i < 4; ++i) {
        const int ch = 2 * i + half;
        const uint32_t bits = __ldg(P.mask[4] + r * 8 + ch);
        float g[32];
#pragma unroll
        for (int j = 0; j < 32; ++j)
          g[j] = fmaf(gr[0], P.chan_out_w[0][ch * 32 + j], fmaf(gr[1], P.chan_out_w[1][ch * 32 + j], gr[2] * P.chan_out_w[2][ch * 32 + j]));
        if (it > 0) mbar_wait(bar(hAFree + i), 1u);  // second release of the previous tile (after its last GEMM)
        stash_wait(i);
        hb_chunk_store(g, 1.0f, bits, ch * 32, sA, rowbase, x7);
        publish(i);
      }
      epilogue(0u, P.inv_scale[0], P.mask[3], r);  // dP_c2
      epilogue(1u, P.inv_scale[1], P.mask[2], r);  // dP_c1
      epilogue(0u, P.inv_scale[2], P.mask[1], r);  // dP_c0
      // ---- second operand dP_s = (g_s * w_sout) * lrelu'(H): overwrites dP_c0 as GEMM 3 releases its K-blocks
#pragma unroll 1
      for (int i = 0; i < 4; ++i) {
        const int ch = 2 * i + half;
        const uint32_t bits = __ldg(P.mask[0] + r * 8 + ch);
        float g[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) g[j] = gs * P.shape_out_w[ch * 32 + j];
        mbar_wait(bar(hAFree + i), 0u);  // first release of this tile (GEMM 3 has read dP_c0)
        stash_wait(i);
        hb_chunk_store(g, 1.0f, bits, ch * 32, sA, rowbase, x7);
        publish(i);
      }
      epilogue(1u, P.inv_scale[3], nullptr, r);  // dF = dP_c0 . W_c0 + dP_s . W_s0  (local_field.8 is linear: no mask)
      // ---- dL/dagg = dF . W_4
      {
        wait_acc(0u);
        const float mul = P.inv_scale[5] * s_inv;
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
          const int ch = 2 * i + half;
          uint32_t v[32];
          tmem_ld32_async(t_row + ch * 32, v);
          tmem_wait(v);
          if (valid) {
            float4* dst = reinterpret_cast<float4*>(P.d_agg + r * kHidden + ch * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              dst[j] = make_float4(__uint_as_float(v[4 * j]) * mul, __uint_as_float(v[4 * j + 1]) * mul,
                                   __uint_as_float(v[4 * j + 2]) * mul, __uint_as_float(v[4 * j + 3]) * mul);
          }
        }
        release_acc(0u);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace tc
}  // namespace npcd

using namespace npcd;

extern "C" int npcd_heads_tc_bwd(const float* d_rgbs, const float* rgbs, const long long* n_samples_dev, long long capacity,
                                 const npcd_pair_stash_layout* layout, void* stash, const void* const* w_t_packed,
                                 const float* inv_scale, const float* chan_out_w, const float* shape_out_w, const float* scale_dev,
                                 int* error_flag, int num_sms, void* stream) {
  NPCD_CHECK_ARG(d_rgbs && rgbs && n_samples_dev && layout && stash && w_t_packed && inv_scale && chan_out_w && shape_out_w && scale_dev,
                 "null pointer");
  NPCD_CHECK_ARG(capacity > 0 && (capacity + 127) / 128 <= layout->h_tiles, "capacity does not match the stash layout");
  NPCD_CHECK_ARG(inv_scale[3] == inv_scale[4], "W_c0^T and W_s0^T must be packed with one common scale");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaFuncSetAttribute(tc::k_heads_bwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kHbSmem);
  if (e != cudaSuccess) {
    set_error("npcd_heads_tc_bwd: cannot opt in to %d bytes of shared memory: %s", tc::kHbSmem, cudaGetErrorString(e));
    return 2;
  }
  static thread_local tc::HeadsBwdParams P;
  memset(&P, 0, sizeof(P));
  uint8_t* base = (uint8_t*)stash;
  for (int i = 0; i < 6; ++i) {
    NPCD_CHECK_ARG(w_t_packed[i], "null weight pointer");
    P.w[i] = (const uint8_t*)w_t_packed[i];
    P.inv_scale[i] = inv_scale[i];
    P.dp[i] = base + layout->hdp[i];
  }
  for (int i = 0; i < 5; ++i) P.mask[i] = (const uint32_t*)(base + layout->hmask[i]);
  memcpy(P.chan_out_w, chan_out_w, sizeof(float) * 3 * 256);
  memcpy(P.shape_out_w, shape_out_w, sizeof(float) * 256);
  P.d_rgbs = (const float4*)d_rgbs;
  P.rgbs = (const float4*)rgbs;
  P.n_samples_dev = n_samples_dev;
  P.capacity = capacity;
  P.scale_dev = scale_dev;
  P.g4 = (float4*)(base + layout->g4);
  P.d_agg = (float*)(base + layout->d_agg);
  P.error_flag = error_flag;
  if (num_sms <= 0) num_sms = 148;
  const long long tiles = (capacity + 127) / 128;
  const unsigned grid = (unsigned)(tiles < num_sms ? tiles : num_sms);
  tc::k_heads_bwd_tc<<<grid, tc::kHbThreads, tc::kHbSmem, st>>>(P);
  return check_launch("npcd_heads_tc_bwd");
}
