// G1/G2/M1/A1/M2/M3 on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
// Same contract as mlp_simt.cu (reference: fields/aggregators/mlp.py:69-88,119-121, fields/mlp.py:38-72, fields/field.py:126-141).
//
// Precision: the reference runs TRUE fp32 GEMMs (no TF32, train_pointnerf.py:16-17) and the parity bar is 1e-4 on RGB after ten
// chained layers, which single-pass TF32/BF16 cannot meet.  Every operand is therefore split into two fp16 halves
// (x = hi + lo, 22 significant bits) and each layer is accumulated in fp32 TMEM from THREE tcgen05.mma.kind::f16 products
// (hi*hi + lo*hi + hi*lo; the dropped lo*lo term is 2^-22 relative).  kind::f16 runs at twice the TF32 rate, so this costs
// 3 bf16-rate passes where 3xTF32 would cost 6.  Weights are pre-scaled by a power of two per layer (exact) so the lo halves
// stay out of the fp16 subnormal range; the inverse scale is folded into the epilogue FMA.
//
// Kernel shape (one persistent CTA per SM, 320 threads, warp-specialised):
//   warp 0      : producer.  Streams pre-swizzled 32 KB weight tiles (256 out x 64 k, hi or lo) from L2 into a 3-stage ring
//                 with cp.async.bulk (UBLKCP) completing on mbarriers.
//   warp 1      : allocates 256 TMEM columns, then one elected lane issues tcgen05.mma (M=128, N=256, K=16, cta_group::1),
//                 A and B both from shared memory (K-major, SWIZZLE_128B descriptors); tcgen05.commit frees ring slots and
//                 publishes the accumulator.
//   warps 2..9  : 256 prologue/epilogue threads, two per tile row (column halves).  tcgen05.ld the fp32 accumulator, apply
//                 scale+bias+LeakyReLU, split to fp16 hi/lo and write the NEXT layer's A operand straight into the swizzled
//                 shared-memory image (activations never leave the SM between layers).
// Tile = 128 rows: 16 samples x 8 neighbour slots (pair MLP) or 128 samples (heads).
#include <cuda_fp16.h>

#include "common.cuh"
#include "npcd_b200.h"

namespace npcd {
namespace tc {

constexpr int kThreadsTc = 320;
constexpr int kEpiThreads = 256;
constexpr int kTileBytesA = 128 * 128;        // one K-block (64 fp16) of 128 rows
constexpr int kTileBytesW = 256 * 128;        // one K-block of 256 output rows
constexpr int kStages = 3;
constexpr int kSmemA = 4 * 2 * kTileBytesA;   // 4 K-blocks x (hi, lo) = 128 KB (aliased by the fp32 aggregation staging)
constexpr int kSmemW = kStages * kTileBytesW; // 96 KB
constexpr int kSmemMisc = 3072;
constexpr int kSmemTotal = kSmemA + kSmemW + kSmemMisc;  // 232448 = the 227 KB per-CTA maximum
constexpr uint32_t kIdesc = (1u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);  // D=f32, A=B=f16, K-major, N=256, M=128
constexpr int kTmemCols = 256;

enum Epi { EPI_ACT = 0, EPI_LINEAR = 1, EPI_AGG = 2, EPI_DOT1 = 3, EPI_DOT3 = 4, EPI_DUMP = 5 };

// ---------------------------------------------------------------------------------------------------------------- PTX ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO(1)<<16 | SBO(1024B>>4)<<32 |
// version 1 <<46 | layout SWIZZLE_128B(2) <<61.  Rows are 128 B (64 fp16), 8-row groups are 1024 B apart.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

// byte offset of the 16-byte chunk holding columns [8*c16, 8*c16+8) of `row` inside a K-block tile
__device__ __forceinline__ uint32_t swz(int row, int c16) { return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((c16 ^ (row & 7)) << 4)); }

__device__ __forceinline__ float lrelu(float x) { return fmaxf(x, 0.01f * x); }

// split 8 fp32 values into fp16 hi / lo and store them as one 16-byte chunk each (columns col..col+7 of `row`)
__device__ __forceinline__ void store_split8(uint8_t* sA, int row, int col, const float (&y)[8]) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __half2 h = __floats2half2_rn(y[2 * j], y[2 * j + 1]);
    const float2 f = __half22float2(h);
    const __half2 l = __floats2half2_rn(y[2 * j] - f.x, y[2 * j + 1] - f.y);
    hi[j] = *reinterpret_cast<const uint32_t*>(&h);
    lo[j] = *reinterpret_cast<const uint32_t*>(&l);
  }
  uint8_t* p = sA + (col >> 6) * (2 * kTileBytesA) + swz(row, (col & 63) >> 3);
  *reinterpret_cast<uint4*>(p) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(p + kTileBytesA) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

struct Layer {
  const uint8_t* w;   // packed tiles: for each K-block: hi tile (32 KB) then lo tile (32 KB)
  const float* bias;  // [256]
  float inv_scale;    // weights were multiplied by 1/inv_scale (a power of two) when packed
  int ksteps;         // K / 16 (a multiple of 1; K-blocks = ceil(ksteps / 4))
  int epi;
};

struct Params {
  Layer layers[6];
  int n_layers;
  int mode;  // 0 = pair MLP, 1 = heads, 2 = probe (one linear layer, dump fp32)
  // pair
  const int* nbr_idx;
  const float4* sample_pos;
  const float* kp_pos;
  const float* kp_feat;
  float* agg;  // [S,256] (pair: out, heads/probe: in)
  // heads
  const float* shape_out_w;
  const float* shape_out_b;
  const float* chan_out_w;
  const float* chan_out_b;
  float4* rgbs;
  float* feat_out;  // optional [S,256] (heads: local_field output; probe: layer output)
  const long long* n_samples_dev;
  long long capacity;
  int* error_flag;
};

// ------------------------------------------------------------------------------------------------------------- kernel ----
__global__ void __launch_bounds__(kThreadsTc, 1) k_field_tc(const __grid_constant__ Params P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sW = smem + kSmemA;
  uint8_t* misc = smem + kSmemA + kSmemW;
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc);  // full[3], empty[3], a_ready, acc_ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 64);
  float* wts = reinterpret_cast<float*>(misc + 128);     // [128] raw inverse-distance weights (pair) / sigma partials (heads)
  float* part3 = reinterpret_cast<float*>(misc + 640);   // [128][3]
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + kStages), bar_a = smem_u32(bars + 2 * kStages),
                 bar_acc = smem_u32(bars + 2 * kStages + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if ((smem_u32(smem) & 1023u) != 0u) {  // SWIZZLE_128B atoms need 1024-byte aligned tiles
    if (threadIdx.x == 0 && P.error_flag) atomicExch(P.error_flag, 1);
    return;
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    mbar_init(bar_a, kEpiThreads);
    mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const long long S = min(*P.n_samples_dev, P.capacity);
  const int rows_per_tile = (P.mode == 0) ? 16 : 128;  // samples per tile
  const long long n_tiles = (S + rows_per_tile - 1) / rows_per_tile;

  if (warp == 0) {
    // ===================================================== producer =====================================================
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int l = 0; l < P.n_layers; ++l) {
          const int nkb = (P.layers[l].ksteps + 3) >> 2;
          const uint8_t* src = P.layers[l].w;
          for (int t = 0; t < 2 * nkb; ++t) {
            mbar_wait(bar_empty + 8 * st, ph ^ 1);
            mbar_expect_tx(bar_full + 8 * st, kTileBytesW);
            bulk_g2s(smem_u32(sW + st * kTileBytesW), src + (size_t)t * kTileBytesW, kTileBytesW, bar_full + 8 * st);
            if (++st == kStages) { st = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer ===================================================
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0, ph_a = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int l = 0; l < P.n_layers; ++l) {
          mbar_wait(bar_a, ph_a);
          ph_a ^= 1;
          tc_fence_after();
          const int ksteps = P.layers[l].ksteps;
          const int nkb = (ksteps + 3) >> 2;
          for (int kb = 0; kb < nkb; ++kb) {
            const int ks_n = min(4, ksteps - kb * 4);
            const uint32_t a_hi = smem_u32(sA + kb * 2 * kTileBytesA), a_lo = a_hi + kTileBytesA;
            // stage "hi": A_hi*W_hi + A_lo*W_hi
            mbar_wait(bar_full + 8 * st, ph);
            tc_fence_after();
            uint32_t b = smem_u32(sW + st * kTileBytesW);
            for (int ks = 0; ks < ks_n; ++ks)
              umma_f16(tmem_base, make_desc(a_hi + ks * 32), make_desc(b + ks * 32), kIdesc, (kb | ks) != 0);
            for (int ks = 0; ks < ks_n; ++ks) umma_f16(tmem_base, make_desc(a_lo + ks * 32), make_desc(b + ks * 32), kIdesc, 1u);
            umma_commit(bar_empty + 8 * st);
            if (++st == kStages) { st = 0; ph ^= 1; }
            // stage "lo": A_hi*W_lo
            mbar_wait(bar_full + 8 * st, ph);
            tc_fence_after();
            b = smem_u32(sW + st * kTileBytesW);
            for (int ks = 0; ks < ks_n; ++ks) umma_f16(tmem_base, make_desc(a_hi + ks * 32), make_desc(b + ks * 32), kIdesc, 1u);
            umma_commit(bar_empty + 8 * st);
            if (++st == kStages) { st = 0; ph ^= 1; }
          }
          umma_commit(bar_acc);
        }
      }
    }
  } else {
    // ============================================ prologue / epilogue threads ===========================================
    const int et = threadIdx.x - 64;         // 0..255
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;        // column half (0: cols 0..127, 1: cols 128..255)
    const int row = q * 32 + lane;           // tile row == TMEM lane
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t ph_acc = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      // ------------------------------------------------------ prologue ---------------------------------------------------
      if (P.mode == 0) {
        // pair MLP input, 112 columns: [feat 0..31 | x: d, sin*10, cos*10, 0,0,0 | y: ... | z: ... | 8 zeros]
        // (the column order is OURS; the first-layer weights are permuted to match when they are packed)
        const long long s = tile * 16 + (row >> 3);
        const int idx = (s < S) ? __ldg(P.nbr_idx + s * kK + (row & 7)) : -1;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        float px = 0.f, py = 0.f, pz = 0.f;
        if (idx >= 0) {
          x = __ldg(P.sample_pos + s);
          px = __ldg(P.kp_pos + (size_t)idx * 3); py = __ldg(P.kp_pos + (size_t)idx * 3 + 1); pz = __ldg(P.kp_pos + (size_t)idx * 3 + 2);
        }
        const float d3[3] = {x.x - px, x.y - py, x.z - pz};
        auto enc_group = [&](int c) {  // 24 columns starting at 32 + 24 c
          float v[24];
          v[0] = d3[c];
          float fr = 3.14159274101257324f;
#pragma unroll
          for (int i = 0; i < kFreqs; ++i) {
            float sn, cs;
            sincosf(d3[c] * fr, &sn, &cs);
            v[1 + i] = sn;
            v[1 + kFreqs + i] = cs;
            fr *= 2.0f;
          }
          v[21] = v[22] = v[23] = 0.f;
          if (idx < 0) {
#pragma unroll
            for (int i = 0; i < 24; ++i) v[i] = 0.f;
          }
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            const float y[8] = {v[ch * 8], v[ch * 8 + 1], v[ch * 8 + 2], v[ch * 8 + 3], v[ch * 8 + 4], v[ch * 8 + 5], v[ch * 8 + 6], v[ch * 8 + 7]};
            store_split8(sA, row, 32 + 24 * c + 8 * ch, y);
          }
        };
        if (half == 0) {
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            float y[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (idx >= 0) {
              const float4 f0 = __ldg(reinterpret_cast<const float4*>(P.kp_feat + (size_t)idx * 32 + ch * 8));
              const float4 f1 = __ldg(reinterpret_cast<const float4*>(P.kp_feat + (size_t)idx * 32 + ch * 8 + 4));
              y[0] = f0.x; y[1] = f0.y; y[2] = f0.z; y[3] = f0.w; y[4] = f1.x; y[5] = f1.y; y[6] = f1.z; y[7] = f1.w;
            }
            store_split8(sA, row, ch * 8, y);
          }
          enc_group(0);
          const float nrm = sqrtf(d3[0] * d3[0] + d3[1] * d3[1] + d3[2] * d3[2]);
          wts[row] = idx >= 0 ? 1.0f / (nrm + 1e-5f) : 0.f;
        } else {
          enc_group(1);
          enc_group(2);
          const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          store_split8(sA, row, 104, z);
        }
      } else {
        // heads / probe input: 128 consecutive rows of the fp32 [S,256] buffer, coalesced loads
        const long long s0 = tile * 128;
#pragma unroll 4
        for (int i = 0; i < 32; ++i) {
          const int f = i * kEpiThreads + et;
          const int r = f >> 6, c = (f & 63) * 4;
          const long long s = s0 + r;
          const float4 v = (s < S) ? __ldg(reinterpret_cast<const float4*>(P.agg + s * kHidden + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
          const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
          const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
          const __half2 l0 = __floats2half2_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2half2_rn(v.z - f1.x, v.w - f1.y);
          uint8_t* p = sA + (c >> 6) * (2 * kTileBytesA) + swz(r, (c & 63) >> 3) + (c & 7) * 2;
          *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
          *reinterpret_cast<uint2*>(p + kTileBytesA) =
              make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
        }
      }
      fence_proxy_async();
      mbar_arrive(bar_a);

      float sigma = 0.f;
      // ------------------------------------------------------ layers -----------------------------------------------------
      for (int l = 0; l < P.n_layers; ++l) {
        const Layer& L = P.layers[l];
        mbar_wait(bar_acc, ph_acc);
        ph_acc ^= 1;
        tc_fence_after();
        const float inv = L.inv_scale;
        if (L.epi == EPI_ACT || L.epi == EPI_LINEAR) {
#pragma unroll 1
          for (int ch = 0; ch < 4; ++ch) {
            const int c0 = half * 128 + ch * 32;
            uint32_t v[32];
            tmem_ld32(t_row + c0, v);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(L.bias + c0 + g * 8));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(L.bias + c0 + g * 8 + 4));
              float y[8] = {fmaf(__uint_as_float(v[g * 8 + 0]), inv, b0.x), fmaf(__uint_as_float(v[g * 8 + 1]), inv, b0.y),
                            fmaf(__uint_as_float(v[g * 8 + 2]), inv, b0.z), fmaf(__uint_as_float(v[g * 8 + 3]), inv, b0.w),
                            fmaf(__uint_as_float(v[g * 8 + 4]), inv, b1.x), fmaf(__uint_as_float(v[g * 8 + 5]), inv, b1.y),
                            fmaf(__uint_as_float(v[g * 8 + 6]), inv, b1.z), fmaf(__uint_as_float(v[g * 8 + 7]), inv, b1.w)};
              if (L.epi == EPI_ACT) {
#pragma unroll
                for (int j = 0; j < 8; ++j) y[j] = lrelu(y[j]);
              }
              if (P.feat_out && L.epi == EPI_LINEAR && P.mode == 1) {
                const long long s = tile * 128 + row;
                if (s < S) {
                  *reinterpret_cast<float4*>(P.feat_out + s * kHidden + c0 + g * 8) = make_float4(y[0], y[1], y[2], y[3]);
                  *reinterpret_cast<float4*>(P.feat_out + s * kHidden + c0 + g * 8 + 4) = make_float4(y[4], y[5], y[6], y[7]);
                }
              }
              store_split8(sA, row, c0 + g * 8, y);
            }
          }
          tc_fence_before();
          fence_proxy_async();
          mbar_arrive(bar_a);
        } else if (L.epi == EPI_DUMP) {
#pragma unroll 1
          for (int ch = 0; ch < 4; ++ch) {
            const int c0 = half * 128 + ch * 32;
            uint32_t v[32];
            tmem_ld32(t_row + c0, v);
            const long long s = tile * 128 + row;
            if (s < S) {
#pragma unroll
              for (int j = 0; j < 32; ++j) P.feat_out[s * kHidden + c0 + j] = fmaf(__uint_as_float(v[j]), inv, __ldg(L.bias + c0 + j));
            }
          }
          tc_fence_before();
        } else if (L.epi == EPI_AGG) {
          // bias + LeakyReLU, scale by the normalised inverse-distance weight of this pair, stage as fp32 in the (now free) A
          // region with a 16-byte XOR swizzle, then sum the 8 slot rows of every sample (fields/aggregators/mlp.py:86-88,119-121)
          epi_bar_sync();  // wts[] was written by other warps in the prologue
          float wsum = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) wsum += wts[(row & ~7) + j];
          const float wn = wts[row] / wsum;  // rows of padded samples have wts = 0 and wsum = 0 -> NaN * never read
          float* stage = reinterpret_cast<float*>(sA);
#pragma unroll 1
          for (int ch = 0; ch < 4; ++ch) {
            const int c0 = half * 128 + ch * 32;
            uint32_t v[32];
            tmem_ld32(t_row + c0, v);
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(L.bias + c0 + g * 4));
              const float w = (wts[row] > 0.f) ? wn : 0.f;
              const float4 o = make_float4(w * lrelu(fmaf(__uint_as_float(v[g * 4 + 0]), inv, b.x)),
                                           w * lrelu(fmaf(__uint_as_float(v[g * 4 + 1]), inv, b.y)),
                                           w * lrelu(fmaf(__uint_as_float(v[g * 4 + 2]), inv, b.z)),
                                           w * lrelu(fmaf(__uint_as_float(v[g * 4 + 3]), inv, b.w)));
              const int c4 = (c0 >> 2) + g;  // float4 chunk index within the 64-chunk row
              *reinterpret_cast<float4*>(stage + row * 256 + ((c4 ^ (row & 7)) << 2)) = o;
            }
          }
          tc_fence_before();
          epi_bar_sync();
#pragma unroll 1
          for (int i = 0; i < 4; ++i) {
            const int f = i * kEpiThreads + et;  // 0..1023: sample-in-tile * 64 + chunk
            const int sl = f >> 6, c4 = f & 63;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 t = *reinterpret_cast<const float4*>(stage + (sl * 8 + j) * 256 + ((c4 ^ j) << 2));
              acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
            }
            const long long s = tile * 16 + sl;
            if (s < S) *reinterpret_cast<float4*>(P.agg + s * kHidden + c4 * 4) = acc;
          }
          epi_bar_sync();  // staging (== A region) is overwritten by the next tile's prologue
        } else {
          // EPI_DOT1 (shape_net.2 -> softplus(x-1)) / EPI_DOT3 (channel_net.8 -> sigmoid): dot of the activated row with 1 or 3
          // output weight vectors; the two column-half threads of a row combine through shared memory.
          const int nout = (L.epi == EPI_DOT1) ? 1 : 3;
          const float* wo = (L.epi == EPI_DOT1) ? P.shape_out_w : P.chan_out_w;
          float part[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
          for (int ch = 0; ch < 4; ++ch) {
            const int c0 = half * 128 + ch * 32;
            uint32_t v[32];
            tmem_ld32(t_row + c0, v);
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(L.bias + c0 + g * 4));
              const float h0 = lrelu(fmaf(__uint_as_float(v[g * 4 + 0]), inv, b.x)), h1 = lrelu(fmaf(__uint_as_float(v[g * 4 + 1]), inv, b.y)),
                          h2 = lrelu(fmaf(__uint_as_float(v[g * 4 + 2]), inv, b.z)), h3 = lrelu(fmaf(__uint_as_float(v[g * 4 + 3]), inv, b.w));
              for (int o = 0; o < nout; ++o) {
                const float4 w = __ldg(reinterpret_cast<const float4*>(wo + o * kHidden + c0 + g * 4));
                part[o] = fmaf(h0, w.x, fmaf(h1, w.y, fmaf(h2, w.z, fmaf(h3, w.w, part[o]))));
              }
            }
          }
          tc_fence_before();
          if (half == 1) {
            if (L.epi == EPI_DOT1) wts[row] = part[0];
            else { part3[row * 3] = part[0]; part3[row * 3 + 1] = part[1]; part3[row * 3 + 2] = part[2]; }
          }
          epi_bar_sync();
          if (half == 0) {
            if (L.epi == EPI_DOT1) {
              const float xs = part[0] + wts[row] + __ldg(P.shape_out_b) - 1.0f;
              sigma = xs > 20.f ? xs : log1pf(expf(xs));
            } else {
              float rgb[3];
#pragma unroll
              for (int o = 0; o < 3; ++o) rgb[o] = 1.0f / (1.0f + expf(-(part[o] + part3[row * 3 + o] + __ldg(P.chan_out_b + o))));
              const long long s = tile * 128 + row;
              if (s < S) P.rgbs[s] = make_float4(rgb[0], rgb[1], rgb[2], sigma);
            }
          }
          epi_bar_sync();  // wts / part3 are reused
          if (L.epi == EPI_DOT1) {  // accumulator consumed, A (= feat) untouched: let the MMA warp start channel_net
            fence_proxy_async();
            mbar_arrive(bar_a);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// Pack an fp32 [256, k_in] nn.Linear weight into per-K-block pre-swizzled fp16 hi/lo tiles (the exact shared-memory image the
// MMA descriptors expect), multiplied by `scale` (a power of two).  perm[k'] = source column of packed column k' (or -1 = 0).
__global__ void k_pack_weights(const float* __restrict__ w, int k_in, const int* __restrict__ perm, int k_pad, float scale,
                               uint8_t* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // (n, k') pairs
  if (idx >= 256 * k_pad) return;
  const int n = idx / k_pad, kp = idx % k_pad;
  const int src = perm ? perm[kp] : (kp < k_in ? kp : -1);
  const float v = src >= 0 ? w[(size_t)n * k_in + src] * scale : 0.f;
  const __half hi = __float2half_rn(v);
  const __half lo = __float2half_rn(v - __half2float(hi));
  const int kb = kp >> 6, kk = kp & 63;
  uint8_t* tile = out + (size_t)kb * 2 * kTileBytesW;
  const size_t off = (size_t)(n >> 3) * 1024 + (n & 7) * 128 + (((kk >> 3) ^ (n & 7)) << 4) + (kk & 7) * 2;
  *reinterpret_cast<__half*>(tile + off) = hi;
  *reinterpret_cast<__half*>(tile + kTileBytesW + off) = lo;
}

}  // namespace tc
}  // namespace npcd

using namespace npcd;

extern "C" int npcd_tc_pack_weights(const float* w, int k_in, const int* perm, int k_pad, float scale, void* out, void* stream) {
  NPCD_CHECK_ARG(w && out, "null pointer");
  NPCD_CHECK_ARG(k_in > 0 && k_pad > 0 && k_pad % 16 == 0 && k_pad <= 256 && (perm || k_in <= k_pad), "bad sizes");
  const int n = 256 * k_pad;
  tc::k_pack_weights<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, k_in, perm, k_pad, scale, (uint8_t*)out);
  return check_launch("npcd_tc_pack_weights");
}

static int launch_tc(const tc::Params& P, long long tiles, int num_sms, cudaStream_t st, const char* what) {
  cudaError_t e = cudaFuncSetAttribute(tc::k_field_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemTotal);
  if (e != cudaSuccess) {
    set_error("%s: cannot opt in to %d bytes of shared memory: %s", what, tc::kSmemTotal, cudaGetErrorString(e));
    return 2;
  }
  if (num_sms <= 0) num_sms = 148;
  const unsigned grid = (unsigned)(tiles < num_sms ? tiles : num_sms);
  tc::k_field_tc<<<grid, tc::kThreadsTc, tc::kSmemTotal, st>>>(P);
  return check_launch(what);
}

static void fill_layer(tc::Layer& L, const npcd_tc_layer& src, int epi) {
  L.w = (const uint8_t*)src.packed_w;
  L.bias = src.bias;
  L.inv_scale = src.inv_scale;
  L.ksteps = src.k_pad / 16;
  L.epi = epi;
}

extern "C" int npcd_field_tc_fwd(const int* nbr_idx, const float* sample_pos, const float* kp_pos, const float* kp_feat,
                                 const long long* n_samples_dev, long long capacity, const npcd_mlp_tc_weights* W,
                                 float* agg_workspace, float* rgbs, float* feat_out, int stages, int* error_flag, int num_sms,
                                 void* stream) {
  NPCD_CHECK_ARG(n_samples_dev && W, "null pointer");
  NPCD_CHECK_ARG(capacity >= 0, "bad capacity");
  if (capacity == 0) return 0;
  NPCD_CHECK_ARG(nbr_idx && sample_pos && kp_pos && kp_feat && agg_workspace && rgbs, "null pointer");
  NPCD_CHECK_ARG(W->feat_dim == 32, "the tensor-core field kernel is specialised for feat_dim = 32 (configs/npcd_srncars.yaml:6)");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = 0;
  if (stages & 1) {
    tc::Params P{};
    for (int i = 0; i < 4; ++i) fill_layer(P.layers[i], W->pair[i], i < 3 ? tc::EPI_ACT : tc::EPI_AGG);
    P.n_layers = 4;
    P.mode = 0;
    P.nbr_idx = nbr_idx; P.sample_pos = (const float4*)sample_pos; P.kp_pos = kp_pos; P.kp_feat = kp_feat; P.agg = agg_workspace;
    P.n_samples_dev = n_samples_dev; P.capacity = capacity; P.error_flag = error_flag;
    rc = launch_tc(P, (capacity + 15) / 16, num_sms, st, "npcd_field_tc_fwd(pair)");
    if (rc) return rc;
  }
  if (stages & 2) {
    tc::Params P{};
    fill_layer(P.layers[0], W->agg, tc::EPI_LINEAR);
    fill_layer(P.layers[1], W->shape, tc::EPI_DOT1);
    for (int i = 0; i < 3; ++i) fill_layer(P.layers[2 + i], W->chan[i], tc::EPI_ACT);
    fill_layer(P.layers[5], W->chan[3], tc::EPI_DOT3);
    P.n_layers = 6;
    P.mode = 1;
    P.agg = agg_workspace; P.rgbs = (float4*)rgbs; P.feat_out = feat_out;
    P.shape_out_w = W->shape_out_w; P.shape_out_b = W->shape_out_b; P.chan_out_w = W->chan_out_w; P.chan_out_b = W->chan_out_b;
    P.n_samples_dev = n_samples_dev; P.capacity = capacity; P.error_flag = error_flag;
    rc = launch_tc(P, (capacity + 127) / 128, num_sms, st, "npcd_field_tc_fwd(heads)");
  }
  return rc;
}

// Probe / self-test: out[s, :] = x[s, :] @ W^T + b for one packed 256x256 layer (validates descriptors, swizzle, TMEM readback).
extern "C" int npcd_tc_linear_probe(const float* x, const long long* n_rows_dev, long long capacity, const npcd_tc_layer* layer,
                                    float* out, int* error_flag, int num_sms, void* stream) {
  NPCD_CHECK_ARG(x && n_rows_dev && layer && out, "null pointer");
  NPCD_CHECK_ARG(capacity > 0, "bad capacity");
  tc::Params P{};
  fill_layer(P.layers[0], *layer, tc::EPI_DUMP);
  P.n_layers = 1;
  P.mode = 2;
  P.agg = const_cast<float*>(x);
  P.feat_out = out;
  P.n_samples_dev = n_rows_dev; P.capacity = capacity; P.error_flag = error_flag;
  return launch_tc(P, (capacity + 127) / 128, num_sms, (cudaStream_t)stream, "npcd_tc_linear_probe");
}
