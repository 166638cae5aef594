// the embedding-side training step.
//   * variational lookup + reparameterised sampling (npcd/models/pointnerf/embeddings/variational_embedding.py:36-58) and its
//     backward into COMPACT per-batch-slot row gradients (the reference materialises a dense [n_obj, P*2F] gradient, 308 MB for
//     SRN-cars, through nn.Embedding's backward);
//   * the KL term (npcd/losses/neural_point_cloud_kl_loss.py:36-37), forward and backward;
//   * Adam on the touched rows only, EXACTLY equal to the dense torch.optim.Adam the reference trainer runs over the whole table
//     (npcd/train/pointnerf_training.py:101-102,152): dense Adam keeps moving a row through its momentum on every later step even
//     when its gradient is zero, so each row remembers the last step it was brought up to (row_step) and the kernel replays the
//     missed zero-gradient steps in registers before it applies the current one.  npcd_embed_adam_rows with obj_idx == NULL brings
//     every row up to date (checkpointing / evaluation).
// All kernels are HBM-bound element-wise passes over B rows of P*2F floats (131 KB per object at P = 512, F = 32).
#include "common.cuh"
#include "npcd_b200.h"

namespace npcd {

// one thread per (slot b, point p, feature f); row layout [P][2F]: mean at [p][f], log-variance at [p][F+f]
__global__ void __launch_bounds__(256) k_embed_fwd(const float* __restrict__ table, const long long* __restrict__ obj_idx,
                                                   long long n_elem, int P, int F, const float* __restrict__ eps,
                                                   float* __restrict__ feats, float* __restrict__ mean_out,
                                                   float* __restrict__ logvar_out, float* __restrict__ std_out) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_elem) return;
  const int f = (int)(e % F);
  const long long bp = e / F;
  const int p = (int)(bp % P);
  const long long b = bp / P;
  const float* row = table + (size_t)obj_idx[b] * P * 2 * F + (size_t)p * 2 * F;
  const float mean = row[f], lv = row[F + f];
  const float sd = expf(0.5f * lv);
  if (feats) feats[e] = eps ? mean + sd * eps[e] : mean;
  if (mean_out) mean_out[e] = mean;
  if (logvar_out) logvar_out[e] = lv;
  if (std_out) std_out[e] = sd;
}

__global__ void __launch_bounds__(256) k_embed_bwd(const float* __restrict__ table, const long long* __restrict__ obj_idx,
                                                   long long n_elem, int P, int F, const float* __restrict__ eps,
                                                   const float* __restrict__ g_feats, const float* __restrict__ g_mean,
                                                   const float* __restrict__ g_logvar, const float* __restrict__ g_std,
                                                   float* __restrict__ d_rows) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_elem) return;
  const int f = (int)(e % F);
  const long long bp = e / F;
  const int p = (int)(bp % P);
  const long long b = bp / P;
  const float lv = table[(size_t)obj_idx[b] * P * 2 * F + (size_t)p * 2 * F + F + f];
  const float half_sd = 0.5f * expf(0.5f * lv);  // d std / d log_var
  const float gf = g_feats ? g_feats[e] : 0.f;
  float dm = gf + (g_mean ? g_mean[e] : 0.f);
  float dl = g_logvar ? g_logvar[e] : 0.f;
  if (eps) dl += gf * eps[e] * half_sd;
  if (g_std) dl += g_std[e] * half_sd;
  float* out = d_rows + (size_t)b * P * 2 * F + (size_t)p * 2 * F;
  out[f] = dm;
  out[F + f] = dl;
}

// KL term, one warp per point: kld = -0.5 * weight * sum_f (1 + lv - mean^2 - exp(lv))
template <bool kBackward>
__global__ void __launch_bounds__(256) k_kl(const float* __restrict__ mean, const float* __restrict__ logvar, long long n_pts, int F,
                                            float weight, const float* __restrict__ g_kld, float* __restrict__ kld,
                                            float* __restrict__ d_mean, float* __restrict__ d_logvar) {
  const int lane = threadIdx.x & 31;
  const long long p = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= n_pts) return;
  const float g = kBackward ? g_kld[p] * weight : 0.f;
  float acc = 0.f;
  for (int f = lane; f < F; f += 32) {
    const float m = mean[p * F + f], lv = logvar[p * F + f];
    if (kBackward) {
      d_mean[p * F + f] = g * m;
      d_logvar[p * F + f] = g * -0.5f * (1.0f - expf(lv));
    } else {
      acc += 1.0f + lv - m * m - expf(lv);
    }
  }
  if (!kBackward) {
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) kld[p] = -0.5f * acc * weight;
  }
}

// Lazy dense-equivalent Adam.  grid = (ceil(row_len / (4*256)), n_slots); a thread owns 4 consecutive elements of one row.
// slot -> row = obj_idx ? obj_idx[slot] : slot.  Duplicate objects inside a batch: the first slot holding a row owns it and sums the
// gradients of the later slots (the embedding backward accumulates duplicates).  d_rows == NULL: zero-gradient replay up to and
// including `step`.  row_step is only read here; k_set_row_step writes it afterwards.
__global__ void __launch_bounds__(256) k_adam_rows(float* __restrict__ table, float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                                                   const int* __restrict__ row_step, const long long* __restrict__ obj_idx, int n_slots,
                                                   long long row_len, const float* __restrict__ d_rows, int step, double lr,
                                                   double beta1, double beta2, float adam_eps) {
  const int slot = blockIdx.y;
  const long long row = obj_idx ? obj_idx[slot] : slot;
  if (obj_idx)
    for (int s = 0; s < slot; ++s)
      if (obj_idx[s] == row) return;  // an earlier slot owns this row
  const long long e0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (e0 >= row_len) return;
  const int n = (int)min((long long)4, row_len - e0);
  float w[4], m[4], v[4], g[4];
  const size_t base = (size_t)row * row_len + e0;
  const bool vec = (n == 4) && ((row_len & 3) == 0);  // 16-byte aligned: cudaMalloc'ed bases, row_len and e0 multiples of 4
  if (vec) {
    const float4 a = *reinterpret_cast<const float4*>(table + base), b = *reinterpret_cast<const float4*>(exp_avg + base),
                 c = *reinterpret_cast<const float4*>(exp_avg_sq + base);
    w[0] = a.x, w[1] = a.y, w[2] = a.z, w[3] = a.w;
    m[0] = b.x, m[1] = b.y, m[2] = b.z, m[3] = b.w;
    v[0] = c.x, v[1] = c.y, v[2] = c.z, v[3] = c.w;
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      w[i] = i < n ? table[base + i] : 0.f;
      m[i] = i < n ? exp_avg[base + i] : 0.f;
      v[i] = i < n ? exp_avg_sq[base + i] : 0.f;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) g[i] = 0.f;
  if (d_rows)
    for (int s = slot; s < n_slots; ++s)
      if (obj_idx[s] == row) {
        const float* src = d_rows + (size_t)s * row_len + e0;
        if (vec) {
          const float4 a = *reinterpret_cast<const float4*>(src);
          g[0] += a.x, g[1] += a.y, g[2] += a.z, g[3] += a.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (i < n) g[i] += src[i];
        }
      }
  const int last = row_step[row];
  // a row that no step ever touched has m = v = 0: its dense updates are exactly zero, nothing to replay
  int s0 = (last == 0) ? step : last + 1;
  double b1p = pow(beta1, (double)(s0 - 1)), b2p = pow(beta2, (double)(s0 - 1));
  const float one_m_b1 = (float)(1.0 - beta1), b2f = (float)beta2, one_m_b2 = (float)(1.0 - beta2);
  for (int s = s0; s <= step; ++s) {
    b1p *= beta1;
    b2p *= beta2;
    const bool with_grad = d_rows != nullptr && s == step;
    const float step_size = (float)(lr / (1.0 - b1p));
    const float bc2_sqrt = (float)sqrt(1.0 - b2p);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i < n) {
        const float gi = with_grad ? g[i] : 0.f;
        m[i] = m[i] + (gi - m[i]) * one_m_b1;          // exp_avg.lerp_(grad, 1 - beta1)
        v[i] = v[i] * b2f + one_m_b2 * gi * gi;         // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
        const float denom = sqrtf(v[i]) / bc2_sqrt + adam_eps;
        w[i] = w[i] - step_size * (m[i] / denom);       // param.addcdiv_(exp_avg, denom, value=-step_size)
      }
    }
  }
  if (vec) {
    *reinterpret_cast<float4*>(table + base) = make_float4(w[0], w[1], w[2], w[3]);
    *reinterpret_cast<float4*>(exp_avg + base) = make_float4(m[0], m[1], m[2], m[3]);
    *reinterpret_cast<float4*>(exp_avg_sq + base) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (i < n) {
        table[base + i] = w[i];
        exp_avg[base + i] = m[i];
        exp_avg_sq[base + i] = v[i];
      }
  }
}

__global__ void k_set_row_step(int* __restrict__ row_step, const long long* __restrict__ obj_idx, long long n_slots, int step) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n_slots) row_step[obj_idx ? obj_idx[s] : s] = step;
}

}  // namespace npcd

using namespace npcd;

extern "C" int npcd_embed_fwd(const float* table, const long long* obj_idx, int n_slots, int n_points, int feat_dim, const float* eps,
                              float* feats, float* mean, float* log_var, float* std, void* stream) {
  NPCD_CHECK_ARG(n_slots >= 0 && n_points > 0 && feat_dim > 0, "bad sizes");
  if (n_slots == 0) return 0;
  NPCD_CHECK_ARG(table && obj_idx, "null pointer");
  const long long n = (long long)n_slots * n_points * feat_dim;
  k_embed_fwd<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(table, obj_idx, n, n_points, feat_dim, eps, feats, mean,
                                                                             log_var, std);
  return check_launch("npcd_embed_fwd");
}

extern "C" int npcd_embed_bwd(const float* table, const long long* obj_idx, int n_slots, int n_points, int feat_dim, const float* eps,
                              const float* g_feats, const float* g_mean, const float* g_log_var, const float* g_std, float* d_rows,
                              void* stream) {
  NPCD_CHECK_ARG(n_slots >= 0 && n_points > 0 && feat_dim > 0, "bad sizes");
  if (n_slots == 0) return 0;
  NPCD_CHECK_ARG(table && obj_idx && d_rows, "null pointer");
  const long long n = (long long)n_slots * n_points * feat_dim;
  k_embed_bwd<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(table, obj_idx, n, n_points, feat_dim, eps, g_feats, g_mean,
                                                                             g_log_var, g_std, d_rows);
  return check_launch("npcd_embed_bwd");
}

extern "C" int npcd_kl_fwd(const float* mean, const float* log_var, long long n_points_total, int feat_dim, float weight, float* kld,
                           void* stream) {
  NPCD_CHECK_ARG(n_points_total >= 0 && feat_dim > 0, "bad sizes");
  if (n_points_total == 0) return 0;
  NPCD_CHECK_ARG(mean && log_var && kld, "null pointer");
  k_kl<false><<<(unsigned)((n_points_total + 7) / 8), 256, 0, (cudaStream_t)stream>>>(mean, log_var, n_points_total, feat_dim, weight,
                                                                                     nullptr, kld, nullptr, nullptr);
  return check_launch("npcd_kl_fwd");
}

extern "C" int npcd_kl_bwd(const float* mean, const float* log_var, long long n_points_total, int feat_dim, float weight,
                           const float* g_kld, float* d_mean, float* d_log_var, void* stream) {
  NPCD_CHECK_ARG(n_points_total >= 0 && feat_dim > 0, "bad sizes");
  if (n_points_total == 0) return 0;
  NPCD_CHECK_ARG(mean && log_var && g_kld && d_mean && d_log_var, "null pointer");
  k_kl<true><<<(unsigned)((n_points_total + 7) / 8), 256, 0, (cudaStream_t)stream>>>(mean, log_var, n_points_total, feat_dim, weight,
                                                                                    g_kld, nullptr, d_mean, d_log_var);
  return check_launch("npcd_kl_bwd");
}

extern "C" int npcd_embed_adam_rows(float* table, float* exp_avg, float* exp_avg_sq, int* row_step, const long long* obj_idx,
                                    int n_slots, long long row_len, const float* d_rows, int step, double lr, double beta1, double beta2,
                                    double adam_eps, void* stream) {
  NPCD_CHECK_ARG(n_slots >= 0 && row_len > 0 && step >= 1, "bad sizes");
  NPCD_CHECK_ARG(n_slots <= 65535, "at most 65535 rows per call");
  if (n_slots == 0) return 0;
  NPCD_CHECK_ARG(table && exp_avg && exp_avg_sq && row_step, "null pointer");
  NPCD_CHECK_ARG(d_rows == nullptr || obj_idx != nullptr, "row gradients need obj_idx");
  dim3 grid((unsigned)((row_len + 1023) / 1024), (unsigned)n_slots);
  // lr / betas are doubles: torch computes the bias corrections in double from the Python floats
  k_adam_rows<<<grid, 256, 0, (cudaStream_t)stream>>>(table, exp_avg, exp_avg_sq, row_step, obj_idx, n_slots, row_len, d_rows, step,
                                                      lr, beta1, beta2, (float)adam_eps);
  k_set_row_step<<<(unsigned)((n_slots + 255) / 256), 256, 0, (cudaStream_t)stream>>>(row_step, obj_idx, n_slots, step);
  return check_launch("npcd_embed_adam_rows");
}
