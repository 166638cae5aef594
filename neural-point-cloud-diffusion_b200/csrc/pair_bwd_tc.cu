// Fused backward of the per-(sample, neighbour) MLP (row B* of SURVEY.md section 8: the autograd of G1/G2/M1/A1, i.e. of
// fields/aggregators/mlp.py:69-88,119-121 and the four hidden layers of `local_field`, utils/model.py:22-36), sm_100a only.
//
// Input: dL/d(aggregated feature) [S,256] and the stash written by the training forward (k_field_tc<MODE_PAIR_TRAIN>): per dense
// pair tile the LeakyReLU sign masks of the four layer outputs, the normalised inverse-distance weight, point index and sample
// index of every row.  Per tile, with G_4 = w_row * dL/dagg[sample(row)]:
//     dP_l = G_{l+1} * lrelu'(X_{l+1})          (epilogue, fp16 hi/lo split, written IN PLACE as the next A operand)
//     G_l  = dP_l . W_l                          (tcgen05, A = dP_l, B = W_l^T packed tiles, three fp16 products, fp32 TMEM)
// for l = 3, 2, 1; for l = 0 only the 32 feature columns of G_0 are formed (N = 32) and scatter-added into dL/dkp_feat.
// Every dP_l operand image is also bulk-copied to HBM: together with the stashed layer inputs X_l they are the operands of the
// weight-gradient GEMMs (wgrad_tc.cu, MN-major descriptors) and of the bias-gradient column sums.
// All dP values carry one power-of-two scale (scale_dev[0]; chosen from max|dL/dagg|) so the fp16 halves stay in range; the
// consumers multiply by scale_dev[1] = 1 / scale.
//
// Same engine as the forward kernel: persistent CTA per SM, warp 0 streams weight tiles (3-stage ring), warp 1 issues the MMAs
// into two alternating 256-column TMEM accumulators, warps 2..9 are the prologue / epilogue threads, warp 10 copies operand
// K-blocks to the stash.
#include "npcd_b200.h"
#include "tc_ptx.cuh"

namespace npcd {
namespace tc {

constexpr int kBwdThreads = 352;
constexpr int kBwdTileW = 256 * 128;
constexpr int kBwdStages = 3;
constexpr int kBwdSmemA = 4 * 2 * kTileBytesA;
constexpr int kBwdSmem = kBwdSmemA + kBwdStages * kBwdTileW + 1024;
// barriers
constexpr int bWFull = 0, bWEmpty = 3, bARdy = 6, bAFree = 10, bAccRdy = 14, bAccFree = 16, bStash = 18;  // 22 barriers

struct BwdParams {
  const uint8_t* w[4];  // packed W_l^T tiles (k_pack_weights of the [in, out] matrix): per out K-block a hi and a lo 32 KB tile
  float inv_scale[4];
  const float* d_agg;
  const uint32_t* mask[4];
  const float* wn;
  const int* idx;
  const int* samp;
  const long long* rows_dev;  // n_tiles * 128
  long long max_tiles;
  const float* scale_dev;  // [2]: scale, 1 / scale
  uint8_t* dp[4];
  float* d_feat;
  int* error_flag;
};

// 32 values of one row -> K-block (c0 >> 6), chunks (c0 & 63) / 8 .. + 3 of the A operand: y = v * mul, times 0.01 where the
// forward activation was not positive, clamped to the fp16 range, fp16 hi/lo split
__device__ __forceinline__ void bwd_chunk_store(const float (&v)[32], float mul, uint32_t bits, int c0, uint8_t* sA, uint32_t rowbase,
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

__global__ void __launch_bounds__(kBwdThreads, 1) k_pair_bwd_tc(const __grid_constant__ BwdParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sW = smem + kBwdSmemA;
  uint8_t* misc = smem + kBwdSmemA + kBwdStages * kBwdTileW;
  const uint32_t bars = smem_u32(misc);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 256);
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if ((smem_u32(smem) & 1023u) != 0u) {
    if (threadIdx.x == 0 && P.error_flag) atomicExch(P.error_flag, 1);
    return;
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < kBwdStages; ++i) { mbar_init(bar(bWFull + i), 1); mbar_init(bar(bWEmpty + i), 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(bar(bARdy + i), 8); mbar_init(bar(bAFree + i), 1); mbar_init(bar(bStash + i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar(bAccRdy + i), 1); mbar_init(bar(bAccFree + i), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_tiles = (int)min(*P.rows_dev >> 7, P.max_tiles);

  if (warp == 0) {
    // ================================================= weight producer ==================================================
    int st = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int l = 3; l >= 0; --l) {
        const uint32_t bytes = l == 0 ? 4096u : (uint32_t)kBwdTileW;  // layer 0: only the 32 feature rows of W_0^T (N = 32)
        for (int t = 0; t < 8; ++t) {
          mbar_wait(bar(bWEmpty + st), ph ^ 1);
          if (elect_one()) {
            mbar_expect_tx(bar(bWFull + st), bytes);
            bulk_g2s(smem_u32(sW + st * kBwdTileW), P.w[l] + (size_t)t * kBwdTileW, bytes, bar(bWFull + st));
          }
          __syncwarp();
          if (++st == kBwdStages) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer ===================================================
    int st = 0;
    uint32_t ph_w = 0, ph_ar = 0, ph_af = 0, lc = 0;
    const uint64_t desc_a0 = make_desc(smem_u32(sA));
    const uint64_t desc_w0 = make_desc(smem_u32(sW));
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int l = 3; l >= 0; --l, ++lc) {
        const uint32_t ab = lc & 1u;
        const uint32_t d_tmem = tmem_base + ab * 256u;
        const uint32_t idesc = l == 0 ? make_idesc(128, 32) : make_idesc(128, 256);
        mbar_wait(bar(bAccFree + ab), ((ph_af >> ab) & 1u) ^ 1u);
        ph_af ^= 1u << ab;
        for (int kb = 0; kb < 4; ++kb) {
          mbar_wait(bar(bARdy + kb), (ph_ar >> kb) & 1u);
          ph_ar ^= 1u << kb;
          const uint64_t a_hi = desc_a0 + (uint64_t)(kb * 2 * (kTileBytesA >> 4)), a_lo = a_hi + (kTileBytesA >> 4);
          mbar_wait(bar(bWFull + st), ph_w);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t b = desc_w0 + (uint64_t)(st * (kBwdTileW >> 4));
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_f16(d_tmem, a_hi + 2 * ks, b + 2 * ks, idesc, (kb | ks) != 0);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_f16(d_tmem, a_lo + 2 * ks, b + 2 * ks, idesc, 1u);
            umma_commit(bar(bWEmpty + st));
          }
          __syncwarp();
          if (++st == kBwdStages) { st = 0; ph_w ^= 1; }
          mbar_wait(bar(bWFull + st), ph_w);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t b = desc_w0 + (uint64_t)(st * (kBwdTileW >> 4));
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_f16(d_tmem, a_hi + 2 * ks, b + 2 * ks, idesc, 1u);
            umma_commit(bar(bWEmpty + st));
            if (l == 0) umma_commit(bar(bAFree + kb));  // the next tile's first operand may overwrite this K-block
            if (kb == 3) umma_commit(bar(bAccRdy + ab));
          }
          __syncwarp();
          if (++st == kBwdStages) { st = 0; ph_w ^= 1; }
        }
      }
    }
  } else if (warp == 10) {
    // ============================================ dP stash (operands of the weight gradients) ===========================
    uint32_t ph_ar = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int l = 3; l >= 0; --l) {
        for (int kb = 0; kb < 4; ++kb) {
          mbar_wait(bar(bARdy + kb), (ph_ar >> kb) & 1u);
          ph_ar ^= 1u << kb;
          if (elect_one()) {
            bulk_s2g(P.dp[l] + ((size_t)tile * 4 + kb) * (2 * kTileBytesA), smem_u32(sA + kb * 2 * kTileBytesA), 2 * kTileBytesA);
            bulk_commit();
            bulk_wait_read0();
            mbar_arrive(bar(bStash + kb));
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
    uint32_t ph_acc = 0, lc = 0, ph_sd = 0, sd_pending = 0;
    const float s_fwd = __ldg(P.scale_dev), s_inv = __ldg(P.scale_dev + 1);

    auto publish = [&](int kb) {
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(bARdy + kb));
      sd_pending |= 1u << kb;
    };
    auto stash_wait = [&](int kb) {
      if ((sd_pending >> kb) & 1u) {
        mbar_wait(bar(bStash + kb), (ph_sd >> kb) & 1u);
        ph_sd ^= 1u << kb;
        sd_pending &= ~(1u << kb);
      }
    };
    auto wait_acc = [&](uint32_t ab) {
      mbar_wait(bar(bAccRdy + ab), (ph_acc >> ab) & 1u);
      ph_acc ^= 1u << ab;
      tc_fence_after();
    };
    auto release_acc = [&](uint32_t ab) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(bAccFree + ab));
    };

    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const size_t r = (size_t)tile * 128 + row;
      const int samp = __ldg(P.samp + r);
      const float wrow = samp >= 0 ? __ldg(P.wn + r) * s_fwd : 0.f;
      // ---- first operand dP_3 = w_row * dL/dagg[sample] * lrelu'(X_4), 32-column chunks (2 i + half)
#pragma unroll 1
      for (int i = 0; i < 4; ++i) {
        const int ch = 2 * i + half;
        const uint32_t bits = __ldg(P.mask[3] + r * 8 + ch);
        float g[32];
        if (samp >= 0) {
          const float4* src = reinterpret_cast<const float4*>(P.d_agg + (size_t)samp * kHidden + ch * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 t = __ldg(src + j);
            g[4 * j] = t.x; g[4 * j + 1] = t.y; g[4 * j + 2] = t.z; g[4 * j + 3] = t.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) g[j] = 0.f;
        }
        if (it > 0) mbar_wait(bar(bAFree + i), (it - 1) & 1u);  // the previous tile's last layer has consumed this K-block
        stash_wait(i);
        bwd_chunk_store(g, wrow, bits, ch * 32, sA, rowbase, x7);
        publish(i);
      }
      // ---- dgrad epilogues of layers 3, 2, 1: dP_{l-1} = (acc * inv_l) * lrelu'(X_l), in place over the consumed operand
#pragma unroll 1
      for (int l = 3; l >= 1; --l, ++lc) {
        const uint32_t ab = lc & 1u;
        wait_acc(ab);
        const float inv = P.inv_scale[l];
        const uint32_t t_acc = t_row + ab * 256u;
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
          const int ch = 2 * i + half;
          uint32_t v[32];
          tmem_ld32_async(t_acc + ch * 32, v);
          const uint32_t bits = __ldg(P.mask[l - 1] + r * 8 + ch);
          tmem_wait(v);
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          stash_wait(i);
          bwd_chunk_store(f, inv, bits, ch * 32, sA, rowbase, x7);
          if (i == 3) release_acc(ab);
          publish(i);
        }
      }
      // ---- layer 0: the 32 feature columns of G_0 are dL/dfeat of this row's point (aggregator.py:121-144 gather, backward)
      {
        const uint32_t ab = lc & 1u;
        wait_acc(ab);
        uint32_t v[32];
        if (half == 0) {
          tmem_ld32_async(t_row + ab * 256u, v);
          tmem_wait(v);
        }
        release_acc(ab);
        ++lc;
        if (half == 0) {
          const int idx = __ldg(P.idx + r);
          if (idx >= 0) {
            const float mul = P.inv_scale[0] * s_inv;
            float* dst = P.d_feat + (size_t)idx * 32;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(v[j]) * mul),
                           "f"(__uint_as_float(v[j + 1]) * mul), "f"(__uint_as_float(v[j + 2]) * mul),
                           "f"(__uint_as_float(v[j + 3]) * mul)
                           : "memory");
          }
        }
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

// max |x| -> power-of-two scale: scale_out = {s, 1 / s} with s * max|x| in [2^target_exp, 2^(target_exp+1)) (1 if x == 0)
__global__ void k_absmax(const float* __restrict__ x, long long n, unsigned* __restrict__ amax_bits) {
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(__ldg(x + i)));
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f && m <= 3.0e38f) atomicMax(amax_bits, __float_as_uint(m));
}
__global__ void k_scale_from_amax(unsigned* amax_bits, int target_exp, float* scale_out) {
  const float m = __uint_as_float(*amax_bits);
  float s = 1.0f;
  if (m > 0.f) {
    int e;
    frexpf(m, &e);  // m = f * 2^e, f in [0.5, 1)  ->  m in [2^(e-1), 2^e)
    s = ldexpf(1.0f, target_exp - (e - 1));
  }
  scale_out[0] = s;
  scale_out[1] = 1.0f / s;
  *amax_bits = 0u;  // ready for the next call
}

}  // namespace tc
}  // namespace npcd

using namespace npcd;

extern "C" int npcd_absmax_scale(const float* x, long long n, int target_exp, void* scratch4, float* scale_out, void* stream) {
  NPCD_CHECK_ARG(scratch4 && scale_out && n >= 0 && (n == 0 || x), "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(scratch4, 0, 4, st);
  if (e != cudaSuccess) {
    set_error("npcd_absmax_scale: %s", cudaGetErrorString(e));
    return 2;
  }
  if (n > 0) {
    const long long want = (n + 1023) / 1024;
    tc::k_absmax<<<(unsigned)(want < 1184 ? want : 1184), 256, 0, st>>>(x, n, (unsigned*)scratch4);
  }
  tc::k_scale_from_amax<<<1, 1, 0, st>>>((unsigned*)scratch4, target_exp, scale_out);
  return check_launch("npcd_absmax_scale");
}

extern "C" int npcd_pair_tc_bwd(const float* d_agg, const npcd_pair_stash_layout* layout, void* stash, const void* const* w_t_packed,
                                const float* inv_scale, const float* scale_dev, float* d_kp_feat, int* error_flag, int num_sms,
                                void* stream) {
  NPCD_CHECK_ARG(d_agg && layout && stash && w_t_packed && inv_scale && scale_dev && d_kp_feat, "null pointer");
  NPCD_CHECK_ARG(layout->max_tiles > 0, "empty stash layout");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaFuncSetAttribute(tc::k_pair_bwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kBwdSmem);
  if (e != cudaSuccess) {
    set_error("npcd_pair_tc_bwd: cannot opt in to %d bytes of shared memory: %s", tc::kBwdSmem, cudaGetErrorString(e));
    return 2;
  }
  tc::BwdParams P{};
  uint8_t* base = (uint8_t*)stash;
  for (int l = 0; l < 4; ++l) {
    NPCD_CHECK_ARG(w_t_packed[l], "null weight pointer");
    P.w[l] = (const uint8_t*)w_t_packed[l];
    P.inv_scale[l] = inv_scale[l];
    P.mask[l] = (const uint32_t*)(base + layout->mask[l]);
    P.dp[l] = base + layout->dp[l];
  }
  P.d_agg = d_agg;
  P.wn = (const float*)(base + layout->wn);
  P.idx = (const int*)(base + layout->idx);
  P.samp = (const int*)(base + layout->samp);
  P.rows_dev = (const long long*)(base + layout->rows_dev);
  P.max_tiles = layout->max_tiles;
  P.scale_dev = scale_dev;
  P.d_feat = d_kp_feat;
  P.error_flag = error_flag;
  if (num_sms <= 0) num_sms = 148;
  const unsigned grid = (unsigned)(layout->max_tiles < num_sms ? layout->max_tiles : num_sms);
  tc::k_pair_bwd_tc<<<grid, tc::kBwdThreads, tc::kBwdSmem, st>>>(P);
  return check_launch("npcd_pair_tc_bwd");
}
