// Unfused vector-env API: dims / bounds / reset / step with same-step autoreset.
// One thread per env instance; SoA state rows are read/written coalesced; the [n][D] API
// tensors are written row-per-thread (a warp covers 32*D contiguous floats).
// HBM-bound: algorithmic bytes per env-step = 4*(2*SF + 2*D + A + 1) + 8*2*SD + 2 + 32 (counters).
#include <cstdarg>

#include "common.cuh"

namespace msacl {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return MSACL_ERR_CUDA;
  }
  return MSACL_OK;
}

template <int ID>
__global__ void __launch_bounds__(128) env_reset_kernel(msacl_env_state_t st) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= st.n) return;
  EnvRegs<ID> r;
  r.episode = st.episode[i];
  r.run = 0;
  r.reset(st.seed, st.env_base + (uint64_t)i);
  r.store(st, i);
}

__global__ void __launch_bounds__(128) quad_init_from_raw_kernel(msacl_env_state_t st) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= st.n) return;
  EnvRegs<kQuadTracking> r;
  r.load(st, i);
  double xd[3], Rd[9];
  float vd[3];
  quad_desired(r.sf, r.sf + 3, 0.0, xd, vd, Rd);
  const double Omd[3] = {0.0, 0.0, 0.0};
  quad_errors(r.sf, r.sf + 3, r.sf + 6, r.sf + 15, xd, vd, Rd, Omd, r.sf + 18);
  r.sd[0] = 0.0;
#pragma unroll
  for (int j = 0; j < 9; ++j) r.sd[1 + j] = Rd[j];
  r.store(st, i);
}

template <int ID>
__global__ void __launch_bounds__(128)
env_step_kernel(msacl_env_state_t st, const float* __restrict__ action, float* __restrict__ next_obs,
                float* __restrict__ reward, uint8_t* __restrict__ terminated, uint8_t* __restrict__ truncated,
                float* __restrict__ final_obs) {
  using E = Env<ID>;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= st.n) return;
  EnvRegs<ID> r;
  r.load(st, i);
  float a[E::A];
#pragma unroll
  for (int j = 0; j < E::A; ++j) a[j] = action[i * E::A + j];
  const float rew = E::step(r.sf, r.sd, a);
  const bool term = r.out_of_bounds();
  r.step += 1;
  // max_step <= 0: a bare env instance as the reference classes are outside a vector env -- no time limit and no
  // autoreset; the state keeps evolving from the terminal state exactly as <env>.step would (e.g. VanderPol.py:100-130)
  const bool vec = st.max_step > 0;
  const bool trunc = vec && r.step >= st.max_step;
  r.ep_return += rew;
  r.ep_len += 1;
  if (final_obs) store_row<E::D>(final_obs, i, r.obs());
  if (vec && (term || trunc)) {   // gymnasium 0.28.1 SyncVectorEnv: reset in the same step
    r.episode += 1;
    r.run = 0;
    r.reset(st.seed, st.env_base + (uint64_t)i);
  }
  store_row<E::D>(next_obs, i, r.obs());
  reward[i] = rew;
  terminated[i] = term ? 1 : 0;
  truncated[i] = trunc ? 1 : 0;
  r.store(st, i);
}

// One BaseSampler._n_step (RL/trainer/sampler/base.py:118-163,220) for policy outputs computed OUTSIDE the kernel: the
// path for policies the fused rollout kernels are not specialised for (any depth / width / activation; the reference
// builds arbitrary MLPs, RL/apprfunc/mlp.py:18-33).  Thread = env instance; same Philox streams, sampling arithmetic,
// reward / cost scaling, autoreset and n-step run / emit bookkeeping as the env phase of rollout_fused.cu / rollout_tc.cu.
template <int ID>
__global__ void __launch_bounds__(128)
rollout_step_kernel(msacl_env_state_t st, const float* __restrict__ logits, float min_log_std, float max_log_std, uint32_t step,
                    int n_step, float reward_scale, float cost_scale, const float* __restrict__ eps, int deterministic,
                    msacl_transitions_t out, double* stats) {
  using E = Env<ID>;
  constexpr int D = E::D, A = E::A;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float v[5] = {0.f, 0.f, 0.f, 0.f, 0.f};      // episodes, sum of returns, sum of lengths, terminated, truncated
  if (i < st.n) {
    EnvRegs<ID> e;
    e.load(st, i);
    float z[4] = {0.f, 0.f, 0.f, 0.f};
    if (!deterministic) {
      if (eps) {
#pragma unroll
        for (int j = 0; j < A; ++j) z[j] = eps[i * A + j];
      } else {
        action_noise4(st.seed, st.env_base + (uint64_t)i, step, z);
      }
    }
    if (out.obs) store_row<D>(out.obs, i, e.obs());
    float act[A];
    float lp_gauss = 0.f, lp_tanh = 0.f, lp_scale = 0.f;
#pragma unroll
    for (int j = 0; j < A; ++j) {      // TanhGaussDistribution.sample / mode (act_distribution_cls.py:45-57, 90-95)
      const float mean = logits[i * (2 * A) + j];
      const float ls = logits[i * (2 * A) + A + j];
      const float sd = expf(fminf(fmaxf(ls, min_log_std), max_log_std));
      const float half = (E::act_high(j) - E::act_low(j)) / 2.0f;
      const float mid = (E::act_high(j) + E::act_low(j)) / 2.0f;
      const float u = deterministic ? mean : (sd * z[j] + mean);
      const float th = tanhf(u);
      const float a_lim = half * th + mid;
      const float diff = u - mean;
      const float g = ((-(diff * diff)) / (2.0f * (sd * sd)) - logf(sd)) - 0.91893853320467267f;
      const float t = logf(1.000001f - th * th);
      const float sc = logf(half);
      lp_gauss = (j == 0) ? g : lp_gauss + g;
      lp_tanh = (j == 0) ? t : lp_tanh + t;
      lp_scale = (j == 0) ? sc : lp_scale + sc;
      act[j] = fminf(fmaxf(a_lim, E::act_low(j)), E::act_high(j));       // base.py:141-143
    }
    const float logp = (lp_gauss - lp_tanh) - lp_scale;
    const float rew = E::step(e.sf, e.sd, act);
    const bool term = e.out_of_bounds();
    e.step += 1;
    const bool trunc = e.step >= st.max_step;
    const bool done = term || trunc;
    e.ep_return += rew;
    e.ep_len += 1;
    const float cost = np_rowsum_sq<D>(e.obs()) * cost_scale;          // rew_plus_cost.py:20-21 on real_next_obs
    e.run = min(e.run + 1, n_step);
    const bool emit = e.run >= n_step;
    if (out.obs2) store_row<D>(out.obs2, i, e.obs());                  // pre-reset observation
    if (done) {
      v[0] = 1.f; v[1] = e.ep_return; v[2] = (float)e.ep_len;
      if (term) v[3] = 1.f; else v[4] = 1.f;
      e.episode += 1;
      e.run = 0;
      e.reset(st.seed, st.env_base + (uint64_t)i);
    }
    if (out.act) store_row<A>(out.act, i, act);
    if (out.rew) out.rew[i] = rew * reward_scale;
    if (out.cost) out.cost[i] = cost;
    if (out.done) out.done[i] = done ? 1 : 0;
    if (out.logp) out.logp[i] = logp;
    if (out.emit) out.emit[i] = emit ? 1 : 0;
    if (out.logits) {
#pragma unroll
      for (int j = 0; j < 2 * A; ++j) out.logits[i * (2 * A) + j] = logits[i * (2 * A) + j];
    }
    e.store(st, i);
  }
  if (stats) {                                   // one atomic per warp and statistic that saw an episode end
#pragma unroll
    for (int q = 0; q < 5; ++q) {
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
      if ((threadIdx.x & 31) == 0 && v[q] != 0.f) atomicAdd(&stats[q], (double)v[q]);
    }
  }
}

__global__ void action_noise_kernel(uint64_t seed, uint64_t env_base, int64_t n, int act_dim, uint32_t step, float* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float z[4];
  action_noise4(seed, env_base + (uint64_t)i, step, z);
  for (int j = 0; j < act_dim; ++j) out[i * act_dim + j] = z[j];
}

}  // namespace msacl

using namespace msacl;

__global__ void quad_polar_selftest_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t n, float theta2) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float R[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) R[j] = in[i * 9 + j];
  quad_polar_f32(R, theta2);
#pragma unroll
  for (int j = 0; j < 9; ++j) out[i * 9 + j] = R[j];
}

extern "C" {

const char* msacl_last_error(void) { return g_err; }
int msacl_abi_version(void) { return MSACL_ABI_VERSION; }

int msacl_env_dims(int env_id, int32_t dims[6]) {
  MSACL_DISPATCH_ENV(env_id, {
    using E = Env<ID>;
    dims[0] = E::D; dims[1] = E::A; dims[2] = E::SF; dims[3] = E::SD; dims[4] = E::OBS_OFF; dims[5] = E::CONTROL_STEP;
  });
  return MSACL_OK;
}

__global__ void bounds_kernel(int env_id, float* out) {
  // out: obs_low[12] obs_high[12] act_low[4] act_high[4]
  auto fill = [&](auto tag) {
    using E = decltype(tag);
    for (int j = 0; j < E::D; ++j) { out[j] = E::obs_low(j); out[12 + j] = E::obs_high(j); }
    for (int j = 0; j < E::A; ++j) { out[24 + j] = E::act_low(j); out[28 + j] = E::act_high(j); }
  };
  switch (env_id) {
    case kVanderPol: fill(Env<kVanderPol>{}); break;
    case kPendulum: fill(Env<kPendulum>{}); break;
    case kDuctedFan: fill(Env<kDuctedFan>{}); break;
    case kTwoLink: fill(Env<kTwoLink>{}); break;
    case kSingleTrackCar: fill(Env<kSingleTrackCar>{}); break;
    case kQuadTracking: fill(Env<kQuadTracking>{}); break;
  }
}

int msacl_env_bounds(int env_id, float* obs_low, float* obs_high, float* act_low, float* act_high) {
  int32_t dims[6];
  int rc = msacl_env_dims(env_id, dims);
  if (rc) return rc;
  float* d = nullptr;
  float h[32];
  if (cudaMalloc(&d, sizeof(h)) != cudaSuccess) { set_error("cudaMalloc failed (no CUDA device?)"); return MSACL_ERR_CUDA; }
  bounds_kernel<<<1, 1>>>(env_id, d);
  cudaError_t e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) { set_error("bounds query: %s", cudaGetErrorString(e)); return MSACL_ERR_CUDA; }
  memcpy(obs_low, h, dims[0] * sizeof(float));
  memcpy(obs_high, h + 12, dims[0] * sizeof(float));
  memcpy(act_low, h + 24, dims[1] * sizeof(float));
  memcpy(act_high, h + 28, dims[1] * sizeof(float));
  return MSACL_OK;
}

static int validate_state(const msacl_env_state_t* st) {
  if (!st || st->n <= 0 || st->stride < st->n || !st->sf || !st->step || !st->episode || !st->ep_return ||
      !st->ep_len || !st->run) {
    set_error("invalid env state descriptor");
    return MSACL_ERR_BAD_ARG;
  }
  return MSACL_OK;
}

int msacl_env_reset(const msacl_env_state_t* st, void* stream) {
  if (int rc = validate_state(st)) return rc;
  const int threads = 128;
  const unsigned blocks = (unsigned)((st->n + threads - 1) / threads);
  MSACL_DISPATCH_ENV(st->env_id, (env_reset_kernel<ID><<<blocks, threads, 0, (cudaStream_t)stream>>>(*st)));
  return check_launch("env_reset");
}

int msacl_quad_init_from_raw(const msacl_env_state_t* st, void* stream) {
  if (int rc = validate_state(st)) return rc;
  if (st->env_id != kQuadTracking || !st->sd) { set_error("quad_init_from_raw needs a QuadTracking state"); return MSACL_ERR_BAD_ARG; }
  const int threads = 128;
  const unsigned blocks = (unsigned)((st->n + threads - 1) / threads);
  quad_init_from_raw_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(*st);
  return check_launch("quad_init_from_raw");
}

int msacl_env_step(const msacl_env_state_t* st, const float* action, float* next_obs, float* reward,
                   uint8_t* terminated, uint8_t* truncated, float* final_obs, void* stream) {
  if (int rc = validate_state(st)) return rc;
  if (!action || !next_obs || !reward || !terminated || !truncated) { set_error("env_step: null buffer"); return MSACL_ERR_BAD_ARG; }
  const int threads = 128;
  const unsigned blocks = (unsigned)((st->n + threads - 1) / threads);
  MSACL_DISPATCH_ENV(st->env_id, {
    if (row_store_misaligned<Env<ID>::D>(next_obs) || row_store_misaligned<Env<ID>::D>(final_obs)) {
      set_error("env_step: next_obs / final_obs rows must be aligned to their vector width (16 B if obs_dim %% 4 == 0, 8 B if even)");
      return MSACL_ERR_BAD_ARG;
    }
    env_step_kernel<ID><<<blocks, threads, 0, (cudaStream_t)stream>>>(*st, action, next_obs, reward, terminated, truncated, final_obs);
  });
  return check_launch("env_step");
}

int msacl_rollout_step(const msacl_env_state_t* st, const float* logits, float min_log_std, float max_log_std, uint32_t step,
                       int32_t n_step, float reward_scale, float cost_scale, const float* eps, int32_t deterministic,
                       const msacl_transitions_t* out, double* stats, void* stream) {
  if (int rc = validate_state(st)) return rc;
  if (!logits || !out || n_step <= 0) { set_error("rollout_step: bad argument"); return MSACL_ERR_BAD_ARG; }
  if (st->max_step <= 0) { set_error("rollout_step: the sampler path needs max_step > 0 (bare-env mode is msacl_env_step only)"); return MSACL_ERR_BAD_ARG; }
  const int threads = 128;
  const unsigned blocks = (unsigned)((st->n + threads - 1) / threads);
  MSACL_DISPATCH_ENV(st->env_id, {
    if (row_store_misaligned<Env<ID>::D>(out->obs) || row_store_misaligned<Env<ID>::D>(out->obs2) || row_store_misaligned<Env<ID>::A>(out->act)) {
      set_error("rollout_step: transition obs/obs2/act rows must be aligned to their vector width (16 B if the row length is a multiple of 4 floats, 8 B if even)");
      return MSACL_ERR_BAD_ARG;
    }
    rollout_step_kernel<ID><<<blocks, threads, 0, (cudaStream_t)stream>>>(*st, logits, min_log_std, max_log_std, step, n_step, reward_scale,
                                                                          cost_scale, eps, deterministic, *out, stats);
  });
  return check_launch("rollout_step");
}

int msacl_selftest_quad_polar(const float* in, float* out, int64_t n, float theta2, void* stream) {
  if (!in || !out || n <= 0) { set_error("selftest_quad_polar: bad argument"); return MSACL_ERR_BAD_ARG; }
  quad_polar_selftest_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(in, out, n, theta2);
  return check_launch("selftest_quad_polar");
}

int msacl_action_noise(uint64_t seed, uint64_t env_base, int64_t n, int32_t act_dim, uint32_t step, float* out,
                       void* stream) {
  if (n <= 0 || act_dim < 1 || act_dim > 4 || !out) { set_error("action_noise: bad argument"); return MSACL_ERR_BAD_ARG; }
  const int threads = 128;
  const unsigned blocks = (unsigned)((n + threads - 1) / threads);
  action_noise_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(seed, env_base, n, act_dim, step, out);
  return check_launch("action_noise");
}

}  // extern "C"
