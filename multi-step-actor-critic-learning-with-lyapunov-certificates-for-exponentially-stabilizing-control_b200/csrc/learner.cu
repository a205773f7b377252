// Learner glue kernels: everything of RL/algorithm/msacl.py:227-460 that is not a dense layer (mlp_tc.cu) or an
// [B, n] window target (targets.cu): TanhGauss rsample / log_prob forward and analytic backward
// (RL/utils/act_distribution_cls.py:59-84), the critic loss gradient (:254-257), the policy loss gradient through the
// reparameterised sample, the entropy term and the clipped stability-advantage surrogate (:349-411), the entropy
// coefficient update (:425-438), bias gradients, and a multi-tensor Adam step (torch.optim.Adam defaults).  With these
// the whole model_update runs without autograd: a fixed sequence of launches on one stream (CUDA-graph friendly),
// every scalar (alpha, losses, entropy) stays on the device.
// All kernels are elementwise / row-wise over M = B * n rows: HBM- or latency-bound, a few hundred KB per call at the
// reference's replay_batch_size 256.
#include "common.cuh"

namespace msacl {

constexpr float kEps = 1e-6f;                       // act_distribution_cls.py:7
constexpr float kLogSqrt2Pi = 0.91893853320467267f;
constexpr int LA_MAX = 8;                           // max action dim handled in registers

__device__ __forceinline__ float block_sum_to_double(float v, double* dst) {
  // warp reduce, then one double atomic per warp (sums are over <= a few million rows)
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0 && v != 0.f) atomicAdd(dst, (double)v);
  return v;
}

// ---- bias gradients: out[z][c] = sum over the rows of split z of x[r][c]
// (blockIdx.y = 256-column chunk: layers wider than 256 units take several chunks)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, int64_t rows, int cols_total, int64_t ld,
                                                     float* __restrict__ out) {
  __shared__ float part[256];
  const int c0 = blockIdx.y * 256;
  const int cols = min(256, cols_total - c0);       // columns of this chunk
  const int rpp = 256 / cols;                       // rows per pass
  const int rsub = threadIdx.x / cols, c = threadIdx.x - rsub * cols;
  const int64_t per = (rows + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = (int64_t)blockIdx.x * per, r1 = min(rows, r0 + per);
  float s = 0.f;
  if (rsub < rpp)
    for (int64_t r = r0 + rsub; r < r1; r += rpp) s += x[r * ld + c0 + c];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x < cols) {
    float t = 0.f;
    for (int j = 0; j < rpp; ++j) t += part[j * cols + threadIdx.x];
    out[(int64_t)blockIdx.x * cols_total + c0 + threadIdx.x] = t;
  }
}

// ---- [a | b] row concatenation (ActionValue input, mlp.py:50-52)
__global__ void __launch_bounds__(256) concat2_kernel(const float* __restrict__ a, int da, const float* __restrict__ b, int db,
                                                      int64_t rows, float* __restrict__ out) {
  const int w = da + db;
  const int64_t total = rows * w, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t r = i / w;
    const int c = (int)(i - r * w);
    out[i] = c < da ? a[r * da + c] : b[r * db + (c - da)];
  }
}

// ---- TanhGaussDistribution.rsample (act_distribution_cls.py:59-71) on explicit N(0,1) draws
__global__ void __launch_bounds__(256)
tanh_gauss_rsample_kernel(int64_t rows, int A, const float* __restrict__ logits, const float* __restrict__ eps,
                          const float* __restrict__ lo, const float* __restrict__ hi, float min_ls, float max_ls,
                          float* __restrict__ act, float* __restrict__ logp) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += stride) {
    float lg = 0.f, lt = 0.f, lsc = 0.f;
    for (int j = 0; j < A; ++j) {
      const float mean = logits[r * 2 * A + j], ls = logits[r * 2 * A + A + j];
      const float sd = expf(fminf(fmaxf(ls, min_ls), max_ls));                 // mlp.py:134-135
      const float half = (hi[j] - lo[j]) / 2.0f, mid = (hi[j] + lo[j]) / 2.0f;
      const float u = mean + eps[r * A + j] * sd;                              // Normal.rsample: loc + eps * scale
      const float th = tanhf(u);
      act[r * A + j] = half * th + mid;
      const float diff = u - mean;
      const float g = ((-(diff * diff)) / (2.0f * (sd * sd)) - logf(sd)) - kLogSqrt2Pi;
      const float t = logf((1.0f + kEps) - th * th);
      const float sc = logf(half);
      lg = j == 0 ? g : lg + g; lt = j == 0 ? t : lt + t; lsc = j == 0 ? sc : lsc + sc;
    }
    logp[r] = (lg - lt) - lsc;
  }
}

// ---- TanhGaussDistribution.log_prob(action) (act_distribution_cls.py:73-84): forward, and d logp / d logits
// u = atanh((1 - EPS) (2a - (hi + lo)) / (hi - lo)); logp = sum_j N(u_j; mean_j, sd_j) - sum_j log(half_j (1 + EPS - tanh(u_j)^2))
__device__ __forceinline__ float lp_u_of_action(float a, float lo, float hi) {
  return atanhf(((1.0f - kEps) * (2.0f * a - (hi + lo))) / (hi - lo));
}

__global__ void __launch_bounds__(256)
tanh_gauss_log_prob_kernel(int64_t rows, int A, const float* __restrict__ logits, const float* __restrict__ act,
                           const float* __restrict__ lo, const float* __restrict__ hi, float min_ls, float max_ls,
                           float* __restrict__ logp) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += stride) {
    float lg = 0.f, lt = 0.f;
    for (int j = 0; j < A; ++j) {
      const float mean = logits[r * 2 * A + j], ls = logits[r * 2 * A + A + j];
      const float sd = expf(fminf(fmaxf(ls, min_ls), max_ls));
      const float u = lp_u_of_action(act[r * A + j], lo[j], hi[j]);
      const float th = tanhf(u);
      const float diff = u - mean;
      const float g = ((-(diff * diff)) / (2.0f * (sd * sd)) - logf(sd)) - kLogSqrt2Pi;
      const float t = logf(((hi[j] - lo[j]) / 2.0f) * ((1.0f + kEps) - th * th));
      lg = j == 0 ? g : lg + g; lt = j == 0 ? t : lt + t;
    }
    logp[r] = lg - lt;
  }
}

// dlogits[r][:] (+)= g[r] * d logp(act[r]) / d logits[r]: d/dmean = (u - mean)/var, d/dlog_std = pass * ((u - mean)^2/var - 1)
__global__ void __launch_bounds__(256)
tanh_gauss_log_prob_bwd_kernel(int64_t rows, int A, const float* __restrict__ logits, const float* __restrict__ act,
                               const float* __restrict__ lo, const float* __restrict__ hi, float min_ls, float max_ls,
                               const float* __restrict__ g, int accumulate, float* __restrict__ dlogits) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += stride) {
    const float gr = g[r];
    for (int j = 0; j < A; ++j) {
      const float mean = logits[r * 2 * A + j], ls = logits[r * 2 * A + A + j];
      const bool pass = ls >= min_ls && ls <= max_ls;                          // clamp passes the gradient inside [min, max]
      const float sd = expf(fminf(fmaxf(ls, min_ls), max_ls));
      const float var = sd * sd;
      const float diff = lp_u_of_action(act[r * A + j], lo[j], hi[j]) - mean;
      const float dm = gr * (diff / var);
      const float dl = pass ? gr * ((diff * diff) / var - 1.0f) : 0.f;
      if (accumulate) { dlogits[r * 2 * A + j] += dm; dlogits[r * 2 * A + A + j] += dl; }
      else { dlogits[r * 2 * A + j] = dm; dlogits[r * 2 * A + A + j] = dl; }
    }
  }
}

// ---- soft-TD backup with the entropy coefficient read from the device (msacl.py:249-252; alpha = exp(log_alpha))
__global__ void __launch_bounds__(256)
q_backup_dev_alpha_kernel(int64_t count, const float* __restrict__ rew, const float* __restrict__ done, const float* __restrict__ q1,
                          const float* __restrict__ q2, const float* __restrict__ logp, float gamma, const float* __restrict__ log_alpha,
                          float* __restrict__ backup) {
  const float alpha = expf(log_alpha[0]);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
    const float nq = fminf(q1[i], q2[i]);
    backup[i] = rew[i] + ((1.0f - done[i]) * gamma) * (nq - alpha * logp[i]);
  }
}

// ---- critic loss (msacl.py:254-257): loss_q = mean((q1 - y)^2) + mean((q2 - y)^2); dq = 2 (q - y) / count
// sums[0..3] += sum (q1-y)^2, sum (q2-y)^2, sum q1, sum q2
__global__ void __launch_bounds__(256)
q_loss_grad_kernel(int64_t count, const float* __restrict__ q1, const float* __restrict__ q2, const float* __restrict__ backup,
                   float* __restrict__ dq1, float* __restrict__ dq2, double* __restrict__ sums) {
  const float sc = 2.0f / (float)count;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
    const float a = q1[i], b = q2[i], y = backup[i];
    const float e1 = a - y, e2 = b - y;
    dq1[i] = sc * e1; dq2[i] = sc * e2;
    s0 += e1 * e1; s1 += e2 * e2; s2 += a; s3 += b;
  }
  block_sum_to_double(s0, &sums[0]); block_sum_to_double(s1, &sums[1]);
  block_sum_to_double(s2, &sums[2]); block_sum_to_double(s3, &sums[3]);
}

// ---- V = sum_j z_j^2 (mlp.py:86-88) backward: dz[r][:] = 2 z[r][:] dV[r]
__global__ void __launch_bounds__(256)
sumsq_bwd_kernel(int64_t rows, int cols, const float* __restrict__ z, const float* __restrict__ dv, float* __restrict__ dz) {
  const int64_t total = rows * cols, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) dz[i] = (2.0f * z[i]) * dv[i / cols];
}

// ---- policy update, stage 1 (msacl.py:365-369): min(Q1, Q2)(obs, a_new) routing and the scalar sums
// dq{1,2}[r] = d loss / d q = -(1/rows) to the smaller one (0.5 each on a tie, as torch.min's backward)
// sums[0] += sum (min_q - alpha logp_new), sums[1] += sum logp_new
__global__ void __launch_bounds__(256)
policy_q_route_kernel(int64_t rows, const float* __restrict__ q1, const float* __restrict__ q2, const float* __restrict__ logp_new,
                      const float* __restrict__ log_alpha, float* __restrict__ dq1, float* __restrict__ dq2, double* __restrict__ sums) {
  const float alpha = expf(log_alpha[0]);
  const float w = -1.0f / (float)rows;
  float s0 = 0.f, s1 = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += stride) {
    const float a = q1[r], b = q2[r], lp = logp_new[r];
    dq1[r] = a < b ? w : (a == b ? 0.5f * w : 0.f);
    dq2[r] = b < a ? w : (a == b ? 0.5f * w : 0.f);
    s0 += fminf(a, b) - alpha * lp;
    s1 += lp;
  }
  block_sum_to_double(s0, &sums[0]); block_sum_to_double(s1, &sums[1]);
}

// ---- policy update, stage 2: d loss_policy / d logits (msacl.py:349-411), loss = -mean(min_q - alpha logp_new) - L_lya
//   (a) through the critics: da = dXq1[:, D:] + dXq2[:, D:]  ->  du = da * half * (1 - tanh(u)^2)
//   (b) entropy term:  (alpha / rows) * d logp_new / d(u, log_std)   (the Gaussian part cancels analytically; tanh part + (-1) on log_std)
//   (c) clipped surrogate on the first step of every window: -(1/B) * [surr1 <= surr2 or ratio inside the clip] * adv * ratio
//       * d log_prob(old_act) / d logits
// sums[2] += sum_b min(surr1, surr2)
__global__ void __launch_bounds__(256)
policy_logits_grad_kernel(int64_t rows, int n_step, int D, int A, const float* __restrict__ logits, const float* __restrict__ eps,
                          const float* __restrict__ dxq1, const float* __restrict__ dxq2, const float* __restrict__ log_alpha,
                          const float* __restrict__ old_act, const float* __restrict__ old_logp, const float* __restrict__ adv,
                          float clip_coef, const float* __restrict__ lo, const float* __restrict__ hi, float min_ls, float max_ls,
                          float* __restrict__ dlogits, double* __restrict__ sums) {
  const float alpha = expf(log_alpha[0]);
  const float ent_w = alpha / (float)rows;
  const int64_t B = rows / n_step;
  const int xw = D + A;
  float s2 = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += stride) {
    const bool first = (r % n_step) == 0;
    float w_sur = 0.f;
    if (first) {                                                        // new log_prob of the stored action under the new policy
      float lg = 0.f, lt = 0.f;
      for (int j = 0; j < A; ++j) {
        const float mean = logits[r * 2 * A + j], ls = logits[r * 2 * A + A + j];
        const float sd = expf(fminf(fmaxf(ls, min_ls), max_ls));
        const float u = lp_u_of_action(old_act[r * A + j], lo[j], hi[j]);
        const float th = tanhf(u), diff = u - mean;
        const float g = ((-(diff * diff)) / (2.0f * (sd * sd)) - logf(sd)) - kLogSqrt2Pi;
        const float t = logf(((hi[j] - lo[j]) / 2.0f) * ((1.0f + kEps) - th * th));
        lg = j == 0 ? g : lg + g; lt = j == 0 ? t : lt + t;
      }
      const int64_t b = r / n_step;
      const float ratio = expf((lg - lt) - old_logp[r]);
      const float a_b = adv[b];
      const float surr1 = ratio * a_b;
      const float surr2 = fminf(fmaxf(ratio, 1.0f - clip_coef), 1.0f + clip_coef) * a_b;
      s2 += fminf(surr1, surr2);
      const bool inside = ratio >= 1.0f - clip_coef && ratio <= 1.0f + clip_coef;
      w_sur = (surr1 <= surr2 || inside) ? -(a_b * ratio) / (float)B : 0.f;
    }
    for (int j = 0; j < A; ++j) {
      const float mean = logits[r * 2 * A + j], ls = logits[r * 2 * A + A + j];
      const bool pass = ls >= min_ls && ls <= max_ls;
      const float sd = expf(fminf(fmaxf(ls, min_ls), max_ls));
      const float e = eps[r * A + j];
      const float th = tanhf(mean + e * sd);
      const float one_m = 1.0f - th * th;
      const float half = (hi[j] - lo[j]) / 2.0f;
      const float da = dxq1[r * xw + D + j] + dxq2[r * xw + D + j];
      float du = da * half * one_m;
      du += ent_w * ((2.0f * th * one_m) / ((1.0f + kEps) - th * th));
      float dm = du;
      float dl = pass ? (du * e * sd - ent_w) : 0.f;
      if (first) {
        const float var = sd * sd;
        const float diff = lp_u_of_action(old_act[r * A + j], lo[j], hi[j]) - mean;
        dm += w_sur * (diff / var);
        if (pass) dl += w_sur * ((diff * diff) / var - 1.0f);
      }
      dlogits[r * 2 * A + j] = dm;
      dlogits[r * 2 * A + A + j] = dl;
    }
  }
  block_sum_to_double(s2, &sums[2]);
}

// ---- entropy coefficient (msacl.py:425-438): loss_alpha = exp(log_alpha) (entropy - target), one Adam step on the scalar.
// sums[1] = sum logp_new over `rows` rows (entropy = -mean); state = {exp_avg, exp_avg_sq}; out[0] = entropy
__global__ void alpha_update_kernel(float* __restrict__ log_alpha, const double* __restrict__ sums, int64_t rows, float target_entropy,
                                    float* __restrict__ state, float one_m_b1, float b2, float one_m_b2, float step_size,
                                    float bc2_sqrt, float eps, float clamp_max, float* __restrict__ out, const float* __restrict__ dyn) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (dyn) { step_size = dyn[0]; bc2_sqrt = dyn[1]; }
  const float entropy = -(float)(sums[1] / (double)rows);
  const float g = expf(log_alpha[0]) * (entropy - target_entropy);
  float m = state[0], v = state[1];
  m = m + (g - m) * one_m_b1;
  v = v * b2 + (one_m_b2 * g) * g;
  state[0] = m; state[1] = v;
  const float denom = sqrtf(v) / bc2_sqrt + eps;
  float la = log_alpha[0] - step_size * (m / denom);
  la = fminf(la, clamp_max);
  log_alpha[0] = la;
  if (out) out[0] = entropy;
}

// ---- multi-tensor Adam (torch.optim.Adam defaults: no weight decay, no amsgrad), one launch per optimizer:
// blockIdx.y = tensor; the gradient of tensor t is the sum of nsplit[t] partials grad[t] + z * numel[t] (split-K
// weight gradients / row-split bias gradients), added in a fixed order (deterministic).
//   m <- m + (g - m)(1 - b1);  v <- v b2 + (1 - b2) g g;  p <- p - step_size * m / (sqrt(v)/bc2_sqrt + eps)
__global__ void __launch_bounds__(256)
adam_multi_kernel(float* const* __restrict__ params, const float* const* __restrict__ grads, float* const* __restrict__ exp_avg,
                  float* const* __restrict__ exp_avg_sq, const int64_t* __restrict__ numel, const int32_t* __restrict__ nsplit,
                  float one_m_b1, float b2, float one_m_b2, float step_size, float bc2_sqrt, float eps, const float* __restrict__ dyn) {
  if (dyn) { step_size = dyn[0]; bc2_sqrt = dyn[1]; }
  const int t = blockIdx.y;
  const int64_t n = numel[t];
  const int ns = nsplit[t];
  float* p = params[t];
  const float* g0 = grads[t];
  float* m_ = exp_avg[t];
  float* v_ = exp_avg_sq[t];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float g = g0[i];
    for (int z = 1; z < ns; ++z) g += g0[(int64_t)z * n + i];
    float m = m_[i], v = v_[i];
    m = m + (g - m) * one_m_b1;
    v = v * b2 + (one_m_b2 * g) * g;
    m_[i] = m; v_[i] = v;
    const float denom = sqrtf(v) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (m / denom);
  }
}

// Bias-correction scalars of an Adam step computed on the device (so that a captured CUDA graph advances them on every
// replay): step <- step + 1; dyn = {lr / (1 - beta1^step), sqrt(1 - beta2^step)} in float64, as torch does on the host.
__global__ void adam_tick_kernel(int32_t* __restrict__ step, float* __restrict__ dyn, double lr, double beta1, double beta2) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int32_t s = step[0] + 1;
  step[0] = s;
  dyn[0] = (float)(lr / (1.0 - pow(beta1, (double)s)));
  dyn[1] = (float)sqrt(1.0 - pow(beta2, (double)s));
}

// sum of split partials into one tensor (tests / gradient inspection)
__global__ void __launch_bounds__(256) reduce_splits_kernel(const float* __restrict__ parts, int64_t n, int ns, float* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float g = parts[i];
    for (int z = 1; z < ns; ++z) g += parts[(int64_t)z * n + i];
    out[i] = g;
  }
}

static inline unsigned grid_for(int64_t work, int threads = 256, int cap = kNumSMs * 8) {
  int64_t b = (work + threads - 1) / threads;
  if (b < 1) b = 1;
  return (unsigned)(b < cap ? b : cap);
}

}  // namespace msacl

using namespace msacl;

#define LCHECK(cond, name)                                              \
  if (!(cond)) { set_error(name ": bad argument"); return MSACL_ERR_BAD_ARG; }

extern "C" {

int msacl_colsum(const float* x, int64_t rows, int32_t cols, int64_t ld, int32_t splits, float* out, void* stream) {
  LCHECK(x && out && rows > 0 && cols > 0 && ld >= cols && splits >= 1, "colsum");
  colsum_kernel<<<dim3((unsigned)splits, (unsigned)((cols + 255) / 256)), 256, 0, (cudaStream_t)stream>>>(x, rows, cols, ld, out);
  return check_launch("colsum");
}

int msacl_concat2(const float* a, int32_t da, const float* b, int32_t db, int64_t rows, float* out, void* stream) {
  LCHECK(a && b && out && da > 0 && db > 0 && rows > 0, "concat2");
  concat2_kernel<<<grid_for(rows * (da + db)), 256, 0, (cudaStream_t)stream>>>(a, da, b, db, rows, out);
  return check_launch("concat2");
}

int msacl_tanh_gauss_rsample(int64_t rows, int32_t act_dim, const float* logits, const float* eps, const float* act_low,
                             const float* act_high, float min_log_std, float max_log_std, float* act, float* logp, void* stream) {
  LCHECK(rows > 0 && act_dim > 0 && act_dim <= LA_MAX && logits && eps && act_low && act_high && act && logp, "tanh_gauss_rsample");
  tanh_gauss_rsample_kernel<<<grid_for(rows), 256, 0, (cudaStream_t)stream>>>(rows, act_dim, logits, eps, act_low, act_high, min_log_std,
                                                                           max_log_std, act, logp);
  return check_launch("tanh_gauss_rsample");
}

int msacl_tanh_gauss_log_prob(int64_t rows, int32_t act_dim, const float* logits, const float* act, const float* act_low,
                              const float* act_high, float min_log_std, float max_log_std, float* logp, void* stream) {
  LCHECK(rows > 0 && act_dim > 0 && act_dim <= LA_MAX && logits && act && act_low && act_high && logp, "tanh_gauss_log_prob");
  tanh_gauss_log_prob_kernel<<<grid_for(rows), 256, 0, (cudaStream_t)stream>>>(rows, act_dim, logits, act, act_low, act_high, min_log_std,
                                                                            max_log_std, logp);
  return check_launch("tanh_gauss_log_prob");
}

int msacl_tanh_gauss_log_prob_bwd(int64_t rows, int32_t act_dim, const float* logits, const float* act, const float* act_low,
                                  const float* act_high, float min_log_std, float max_log_std, const float* grad_logp,
                                  int32_t accumulate, float* grad_logits, void* stream) {
  LCHECK(rows > 0 && act_dim > 0 && act_dim <= LA_MAX && logits && act && act_low && act_high && grad_logp && grad_logits,
         "tanh_gauss_log_prob_bwd");
  tanh_gauss_log_prob_bwd_kernel<<<grid_for(rows), 256, 0, (cudaStream_t)stream>>>(rows, act_dim, logits, act, act_low, act_high,
                                                                                min_log_std, max_log_std, grad_logp, accumulate, grad_logits);
  return check_launch("tanh_gauss_log_prob_bwd");
}

int msacl_q_backup_dev_alpha(int64_t count, const float* rew, const float* done, const float* next_q1, const float* next_q2,
                             const float* next_logp, float gamma, const float* log_alpha, float* backup, void* stream) {
  LCHECK(count > 0 && rew && done && next_q1 && next_q2 && next_logp && log_alpha && backup, "q_backup_dev_alpha");
  q_backup_dev_alpha_kernel<<<grid_for(count), 256, 0, (cudaStream_t)stream>>>(count, rew, done, next_q1, next_q2, next_logp, gamma,
                                                                            log_alpha, backup);
  return check_launch("q_backup_dev_alpha");
}

int msacl_q_loss_grad(int64_t count, const float* q1, const float* q2, const float* backup, float* dq1, float* dq2, double* sums,
                      void* stream) {
  LCHECK(count > 0 && q1 && q2 && backup && dq1 && dq2 && sums, "q_loss_grad");
  cudaMemsetAsync(sums, 0, 4 * sizeof(double), (cudaStream_t)stream);
  q_loss_grad_kernel<<<grid_for(count), 256, 0, (cudaStream_t)stream>>>(count, q1, q2, backup, dq1, dq2, sums);
  return check_launch("q_loss_grad");
}

int msacl_sumsq_bwd(int64_t rows, int32_t cols, const float* z, const float* dv, float* dz, void* stream) {
  LCHECK(rows > 0 && cols > 0 && z && dv && dz, "sumsq_bwd");
  sumsq_bwd_kernel<<<grid_for(rows * cols), 256, 0, (cudaStream_t)stream>>>(rows, cols, z, dv, dz);
  return check_launch("sumsq_bwd");
}

int msacl_policy_q_route(int64_t rows, const float* q1, const float* q2, const float* logp_new, const float* log_alpha, float* dq1,
                         float* dq2, double* sums, void* stream) {
  LCHECK(rows > 0 && q1 && q2 && logp_new && log_alpha && dq1 && dq2 && sums, "policy_q_route");
  cudaMemsetAsync(sums, 0, 3 * sizeof(double), (cudaStream_t)stream);
  policy_q_route_kernel<<<grid_for(rows), 256, 0, (cudaStream_t)stream>>>(rows, q1, q2, logp_new, log_alpha, dq1, dq2, sums);
  return check_launch("policy_q_route");
}

int msacl_policy_logits_grad(int64_t rows, int32_t n_step, int32_t obs_dim, int32_t act_dim, const float* logits, const float* eps,
                             const float* dxq1, const float* dxq2, const float* log_alpha, const float* old_act,
                             const float* old_logp, const float* adv, float clip_coef, const float* act_low, const float* act_high,
                             float min_log_std, float max_log_std, float* grad_logits, double* sums, void* stream) {
  LCHECK(rows > 0 && n_step > 0 && rows % n_step == 0 && obs_dim > 0 && act_dim > 0 && act_dim <= LA_MAX && logits && eps && dxq1 &&
             dxq2 && log_alpha && old_act && old_logp && adv && act_low && act_high && grad_logits && sums,
         "policy_logits_grad");
  policy_logits_grad_kernel<<<grid_for(rows), 256, 0, (cudaStream_t)stream>>>(rows, n_step, obs_dim, act_dim, logits, eps, dxq1, dxq2,
                                                                           log_alpha, old_act, old_logp, adv, clip_coef, act_low, act_high,
                                                                           min_log_std, max_log_std, grad_logits, sums);
  return check_launch("policy_logits_grad");
}

int msacl_alpha_update(float* log_alpha, const double* sums, int64_t rows, float target_entropy, float* adam_state,
                       float one_minus_beta1, float beta2, float one_minus_beta2, float step_size, float bc2_sqrt, float eps,
                       float clamp_max, float* entropy_out, const float* dyn, void* stream) {
  LCHECK(log_alpha && sums && rows > 0 && adam_state, "alpha_update");
  alpha_update_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(log_alpha, sums, rows, target_entropy, adam_state, one_minus_beta1, beta2,
                                                         one_minus_beta2, step_size, bc2_sqrt, eps, clamp_max, entropy_out, dyn);
  return check_launch("alpha_update");
}

int msacl_adam_multi(int32_t count, float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                     const int64_t* numel, const int32_t* nsplit, int64_t max_numel, float one_minus_beta1, float beta2,
                     float one_minus_beta2, float step_size, float bc2_sqrt, float eps, const float* dyn, void* stream) {
  LCHECK(count > 0 && params && grads && exp_avg && exp_avg_sq && numel && nsplit && max_numel > 0, "adam_multi");
  int64_t bx = (max_numel + 255) / 256;
  if (bx > 64) bx = 64;
  adam_multi_kernel<<<dim3((unsigned)bx, (unsigned)count), 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, numel, nsplit,
                                                                                        one_minus_beta1, beta2, one_minus_beta2, step_size, bc2_sqrt, eps, dyn);
  return check_launch("adam_multi");
}

int msacl_adam_tick(int32_t* step, float* dyn, double lr, double beta1, double beta2, void* stream) {
  LCHECK(step && dyn, "adam_tick");
  adam_tick_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(step, dyn, lr, beta1, beta2);
  return check_launch("adam_tick");
}

int msacl_reduce_splits(const float* parts, int64_t numel, int32_t nsplit, float* out, void* stream) {
  LCHECK(parts && out && numel > 0 && nsplit >= 1, "reduce_splits");
  reduce_splits_kernel<<<grid_for(numel), 256, 0, (cudaStream_t)stream>>>(parts, numel, nsplit, out);
  return check_launch("reduce_splits");
}

}  // extern "C"
