// G1/G2/M1/A1/M2/M3 -- fp32 SIMT implementation (correctness baseline and ncu reference for the tensor-core kernel in mlp_tc.cu).
// Replaces Aggregator.get_keypoint_data / mask_to_batch_ray_idx (fields/aggregators/aggregator.py:121-156), aggregators.MLP
// .get_local_feat / .aggregate_local_feat (fields/aggregators/mlp.py:69-88,119-121), PositionalEncoder1D
// (npcd/utils/positional_encoder.py:14-20), fields.MLP.get_shape / get_channels (fields/mlp.py:38-72) and the activations in
// Field.forward (fields/field.py:126-141).
//
// k_pair_mlp : tile = 16 samples x 8 neighbour slots = 128 rows.  Gather + posenc build the [K0,128] input in shared memory,
//              4 hidden layers (Linear+LeakyReLU) run in place, then the inverse-distance weighted 8->1 sum is taken IN REGISTERS
//              (each thread owns the 8 slot-rows of one sample).  Because the weights sum to 1, the last (linear) layer of
//              local_field commutes with the aggregation:  sum_j w_j (W5 h_j + b5) = W5 (sum_j w_j h_j) + b5, so it is applied
//              once per sample in k_heads instead of once per pair (-23% FLOPs, rounding-level difference only).
// k_heads    : tile = 128 samples: local_field.8, shape_net (-> softplus(x-1)), channel_net (-> sigmoid).
// Register tile 8 rows x 16 cols per thread, 256 threads, weights streamed in 16-row chunks with cp.async double buffering.
#include "common.cuh"
#include "npcd_b200.h"

namespace npcd {

constexpr int kTileRows = 128;
constexpr int kPitch = 132;   // activation row pitch (floats): [k][row], 16-byte aligned rows
constexpr int kChunk = 16;    // weight rows per cp.async stage
constexpr int kThreads = 256;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void load_chunk(float* wbuf, const float* __restrict__ wt, int k0) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int idx = threadIdx.x + q * kThreads;  // float4 index within the [16][256] chunk
    const int row = idx >> 6, c4 = idx & 63;
    cp_async16(wbuf + row * kHidden + c4 * 4, wt + (size_t)(k0 + row) * kHidden + c4 * 4);
  }
}

// acc[i][q*4+c] += sum_k act[k][ty*8+i] * Wt[k][tx*4 + 64*q + c]
__device__ __forceinline__ void gemm_tile(const float* __restrict__ act, const float* __restrict__ wt, int K, float* wbuf,
                                          float (&acc)[8][16]) {
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[i][j] = 0.f;
  const int nchunks = K / kChunk;
  load_chunk(wbuf, wt, 0);
  cp_async_commit();
  for (int c = 0; c < nchunks; ++c) {
    if (c + 1 < nchunks) {
      load_chunk(wbuf + ((c + 1) & 1) * kChunk * kHidden, wt, (c + 1) * kChunk);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* wb = wbuf + (c & 1) * kChunk * kHidden;
    const float* ab = act + (size_t)c * kChunk * kPitch + ty * 8;
#pragma unroll 4
    for (int kk = 0; kk < kChunk; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(ab + kk * kPitch);
      const float4 a1 = *reinterpret_cast<const float4*>(ab + kk * kPitch + 4);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float w[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(wb + kk * kHidden + tx * 4 + 64 * q);
        w[q * 4] = v.x; w[q * 4 + 1] = v.y; w[q * 4 + 2] = v.z; w[q * 4 + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
}

__device__ __forceinline__ float lrelu(float x) { return x > 0.f ? x : 0.01f * x; }

// act[n][rows] = lrelu?(acc + b[n])   (all threads have passed the final barrier of gemm_tile)
template <bool kAct>
__device__ __forceinline__ void store_tile(float* act, const float* __restrict__ bias, const float (&acc)[8][16]) {
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int n = tx * 4 + 64 * (j >> 2) + (j & 3);
    const float b = __ldg(bias + n);
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] = acc[i][j] + b; if (kAct) v[i] = lrelu(v[i]); }
    float* dst = act + (size_t)n * kPitch + ty * 8;
    *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kThreads, 1) k_pair_mlp(const int* __restrict__ nbr_idx, const float4* __restrict__ sample_pos,
                                                          const float* __restrict__ kp_pos, const float* __restrict__ kp_feat,
                                                          const long long* __restrict__ n_samples_dev, long long capacity,
                                                          npcd_mlp_simt_weights W, float* __restrict__ agg) {
  extern __shared__ __align__(16) float smem[];
  float* act = smem;                                  // [256][kPitch]
  float* wbuf = act + kHidden * kPitch;               // [2][16][256]
  float* wts = wbuf + 2 * kChunk * kHidden;           // [128] raw inverse-distance weights
  const long long S = min(*n_samples_dev, capacity);
  const int F = W.feat_dim;
  const int K0 = (F + 3 + 6 * kFreqs + kChunk - 1) / kChunk * kChunk;
  const long long n_tiles = (S + 15) / 16;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    // ---- prologue: gather + relative position + posenc (aggregators/mlp.py:69-83) ----
    {
      const int row = threadIdx.x >> 1, half = threadIdx.x & 1;
      const long long s = tile * 16 + (row >> 3);
      const int idx = (s < S) ? __ldg(nbr_idx + s * kK + (row & 7)) : -1;
      if (half == 0) {
        if (idx >= 0) {
          const float* f = kp_feat + (size_t)idx * F;
          for (int k = 0; k < F; ++k) act[k * kPitch + row] = __ldg(f + k);
        } else {
          for (int k = 0; k < F; ++k) act[k * kPitch + row] = 0.f;
        }
      } else {
        float* a = act + (size_t)F * kPitch + row;
        if (idx >= 0) {
          const float4 x = __ldg(sample_pos + s);
          const float d[3] = {x.x - __ldg(kp_pos + (size_t)idx * 3), x.y - __ldg(kp_pos + (size_t)idx * 3 + 1),
                              x.z - __ldg(kp_pos + (size_t)idx * 3 + 2)};
          const float nrm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
          wts[row] = 1.0f / (nrm + 1e-5f);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            a[c * kPitch] = d[c];
            float fr = 3.14159274101257324f;  // fl32(pi); 2^i scaling is exact (positional_encoder.py:14)
#pragma unroll
            for (int i = 0; i < kFreqs; ++i) {
              float sn, cs;
              sincosf(d[c] * fr, &sn, &cs);
              a[(3 + c * 2 * kFreqs + i) * kPitch] = sn;
              a[(3 + c * 2 * kFreqs + kFreqs + i) * kPitch] = cs;
              fr *= 2.0f;
            }
          }
        } else {
          wts[row] = 0.f;
          for (int k = 0; k < 3 + 6 * kFreqs; ++k) a[k * kPitch] = 0.f;
        }
        for (int k = F + 3 + 6 * kFreqs; k < K0; ++k) act[k * kPitch + row] = 0.f;
      }
    }
    __syncthreads();
    float acc[8][16];
    gemm_tile(act, W.pair_wt[0], K0, wbuf, acc);
    store_tile<true>(act, W.pair_b[0], acc);
    gemm_tile(act, W.pair_wt[1], kHidden, wbuf, acc);
    store_tile<true>(act, W.pair_b[1], acc);
    gemm_tile(act, W.pair_wt[2], kHidden, wbuf, acc);
    store_tile<true>(act, W.pair_b[2], acc);
    gemm_tile(act, W.pair_wt[3], kHidden, wbuf, acc);
    // ---- epilogue: bias + LeakyReLU + normalised inverse-distance aggregation over the 8 slots (mlp.py:86-88,119-121) ----
    {
      const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
      const long long s = tile * 16 + ty;
      float w[8], norm = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) { w[i] = wts[ty * 8 + i]; norm += w[i]; }
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = w[i] / norm;  // norm > 0: every kept sample has >= 1 neighbour
      if (s < S) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float o[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int n = tx * 4 + 64 * q + c;
            const float b = __ldg(W.pair_b[3] + n);
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) sum = fmaf(w[i], lrelu(acc[i][q * 4 + c] + b), sum);
            o[c] = sum;
          }
          *reinterpret_cast<float4*>(agg + s * kHidden + tx * 4 + 64 * q) = make_float4(o[0], o[1], o[2], o[3]);
        }
      }
    }
    __syncthreads();  // wts / act reused by the next tile
  }
}

// sum over the 16 lanes that share a row group (tx = lane & 15)
__device__ __forceinline__ float half_warp_sum(float v) {
#pragma unroll
  for (int o = 8; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(kThreads, 1) k_heads(const float* __restrict__ agg, const long long* __restrict__ n_samples_dev,
                                                       long long capacity, npcd_mlp_simt_weights W, float4* __restrict__ rgbs,
                                                       float* __restrict__ feat_out) {
  extern __shared__ __align__(16) float smem[];
  float* act = smem;
  float* wbuf = act + kHidden * kPitch;
  const long long S = min(*n_samples_dev, capacity);
  const long long n_tiles = (S + kTileRows - 1) / kTileRows;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long s0 = tile * kTileRows;
    // load agg tile transposed into act[k][row]
    for (int r = threadIdx.x >> 5; r < kTileRows; r += kThreads / 32) {
      const long long s = s0 + r;
      const int lane = threadIdx.x & 31;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = lane * 4 + h * 128;
        const float4 v = (s < S) ? __ldg(reinterpret_cast<const float4*>(agg + s * kHidden + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
        act[(k + 0) * kPitch + r] = v.x; act[(k + 1) * kPitch + r] = v.y;
        act[(k + 2) * kPitch + r] = v.z; act[(k + 3) * kPitch + r] = v.w;
      }
    }
    __syncthreads();
    float acc[8][16];
    gemm_tile(act, W.agg_wt, kHidden, wbuf, acc);          // local_field.8 (linear) -> feat
    store_tile<false>(act, W.agg_b, acc);
    if (feat_out) {
      for (int r = threadIdx.x >> 5; r < kTileRows; r += kThreads / 32) {
        const long long s = s0 + r;
        if (s < S)
          for (int k = threadIdx.x & 31; k < kHidden; k += 32) feat_out[s * kHidden + k] = act[k * kPitch + r];
      }
    }
    // ---- shape_net: 256 -> 256 -> 1, softplus(x - 1)  (fields/mlp.py:50, field.py:29-30) ----
    gemm_tile(act, W.shape_wt, kHidden, wbuf, acc);
    float sig[8];
    {
      float part[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int n = tx * 4 + 64 * (j >> 2) + (j & 3);
        const float b = __ldg(W.shape_b + n), wo = __ldg(W.shape_out_w + n);
#pragma unroll
        for (int i = 0; i < 8; ++i) part[i] = fmaf(lrelu(acc[i][j] + b), wo, part[i]);
      }
      const float bo = __ldg(W.shape_out_b);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float x = half_warp_sum(part[i]) + bo - 1.0f;
        sig[i] = x > 20.f ? x : log1pf(expf(x));
      }
    }
    // ---- channel_net: 4 x (256 -> 256, LeakyReLU) -> 3, sigmoid  (fields/mlp.py:71, field.py:139-140) ----
    gemm_tile(act, W.chan_wt[0], kHidden, wbuf, acc);
    store_tile<true>(act, W.chan_b[0], acc);
    gemm_tile(act, W.chan_wt[1], kHidden, wbuf, acc);
    store_tile<true>(act, W.chan_b[1], acc);
    gemm_tile(act, W.chan_wt[2], kHidden, wbuf, acc);
    store_tile<true>(act, W.chan_b[2], acc);
    gemm_tile(act, W.chan_wt[3], kHidden, wbuf, acc);
    {
      float part[3][8];
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int i = 0; i < 8; ++i) part[c][i] = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int n = tx * 4 + 64 * (j >> 2) + (j & 3);
        const float b = __ldg(W.chan_b[3] + n);
        const float w0 = __ldg(W.chan_out_w + n), w1 = __ldg(W.chan_out_w + kHidden + n), w2 = __ldg(W.chan_out_w + 2 * kHidden + n);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float h = lrelu(acc[i][j] + b);
          part[0][i] = fmaf(h, w0, part[0][i]);
          part[1][i] = fmaf(h, w1, part[1][i]);
          part[2][i] = fmaf(h, w2, part[2][i]);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float rgb[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float x = half_warp_sum(part[c][i]) + __ldg(W.chan_out_b + c);
          rgb[c] = 1.0f / (1.0f + expf(-x));
        }
        const long long s = s0 + ty * 8 + i;
        if (tx == 0 && s < S) rgbs[s] = make_float4(rgb[0], rgb[1], rgb[2], sig[i]);
      }
    }
    __syncthreads();
  }
}

constexpr size_t kSimtSmem = (size_t)(kHidden * kPitch + 2 * kChunk * kHidden + kTileRows) * sizeof(float);

}  // namespace npcd

extern "C" int npcd_field_simt_fwd(const int* nbr_idx, const float* sample_pos, const float* kp_pos, const float* kp_feat,
                                   const long long* n_samples_dev, long long capacity, const npcd_mlp_simt_weights* weights,
                                   float* agg_workspace, float* rgbs, float* feat_out, int stages, int num_sms, void* stream) {
  using namespace npcd;
  NPCD_CHECK_ARG(n_samples_dev && weights, "null pointer");
  NPCD_CHECK_ARG(capacity >= 0, "bad capacity");
  if (capacity == 0) return 0;
  NPCD_CHECK_ARG(nbr_idx && sample_pos && kp_pos && kp_feat && agg_workspace && rgbs, "null pointer");
  NPCD_CHECK_ARG(weights->feat_dim > 0 && weights->feat_dim + 3 + 6 * kFreqs <= kHidden, "feat_dim out of range");
  cudaStream_t st = (cudaStream_t)stream;
  cudaFuncSetAttribute(k_pair_mlp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSimtSmem);
  cudaFuncSetAttribute(k_heads, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSimtSmem);
  if (num_sms <= 0) num_sms = 148;
  const long long pair_tiles = (capacity + 15) / 16, head_tiles = (capacity + kTileRows - 1) / kTileRows;
  const unsigned g1 = (unsigned)(pair_tiles < num_sms ? pair_tiles : num_sms);
  const unsigned g2 = (unsigned)(head_tiles < num_sms ? head_tiles : num_sms);
  if (stages & 1)
    k_pair_mlp<<<g1, kThreads, kSimtSmem, st>>>(nbr_idx, (const float4*)sample_pos, kp_pos, kp_feat, n_samples_dev, capacity,
                                                *weights, agg_workspace);
  if (stages & 2) k_heads<<<g2, kThreads, kSimtSmem, st>>>(agg_workspace, n_samples_dev, capacity, *weights, (float4*)rgbs, feat_out);
  return check_launch("npcd_field_simt_fwd");
}
