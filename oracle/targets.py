"""Oracle (test infrastructure): MSACL learner target / label math over [B, n] replay windows.

NumPy float32 restatement of RL/algorithm/msacl.py given the network outputs as inputs:
  coefficients        :153-164
  soft-TD backup      :243-252   (_q_update; elementwise over the window)
  Lyapunov risk       :280-329   (_lyapunov_update: clipped-IS cumprod, boundedness hinge,
                                  exponential-stability label, lambda-weighted decrease hinge)
  stability advantage :383-405   (_policy_update: lambda-weighted advantage, batch normalise
                                  with unbiased std, PPO-clipped surrogate)
Gradients w.r.t. the differentiable inputs (V(obs), V(obs2), new log-prob) are written out
analytically so the CUDA backward kernels can be checked without autograd.
"""
import numpy as np

f32 = np.float32


def coefficients(n_step, lya_eta=0.15, retrace_lambda=0.95, alpha1=1.0, alpha2=2.0):
    """msacl.py:153-164 -> (start_obs_norm_coef, lya_diff_coef, start_lya_coef), f32[n]."""
    k1 = np.arange(1, n_step + 1)
    base = f32(1 - lya_eta)
    start_obs_norm_coef = ((base ** k1).astype(f32) * f32(alpha2 / alpha1)).astype(f32) ** f32(0.5)
    lam = np.power(f32(retrace_lambda), np.arange(n_step)).astype(f32)
    lya_diff_coef = (lam / lam.sum(dtype=f32)).astype(f32)
    start_lya_coef = np.power(f32(1 - lya_eta), k1).astype(f32)
    return start_obs_norm_coef.astype(f32), lya_diff_coef, start_lya_coef


def q_backup(rew, done, next_q1, next_q2, next_logp, gamma, alpha):
    """msacl.py:249-252."""
    next_q = np.minimum(next_q1, next_q2)
    return (rew + (f32(1) - done) * f32(gamma) * (next_q - f32(alpha) * next_logp)).astype(f32)


def lyapunov_risk(obs, obs2, logp_new, logp_old, lya_obs, lya_obs2, coefs,
                  alpha1=1.0, alpha2=2.0, lya_diff_scale=10.0, lya_positive_scale=1.0):
    """msacl.py:280-329.  Returns dict with loss, the intermediate labels and the gradients
    d loss / d lya_obs and d loss / d lya_obs2 ([B,n])."""
    son_coef, diff_coef, sl_coef = coefs
    B, n = logp_new.shape
    ratio = np.exp((logp_new - logp_old).astype(f32)).astype(f32)
    clip_ratio = np.clip(ratio, f32(0), f32(1))
    isr = np.cumprod(clip_ratio, axis=1, dtype=f32)
    obs_pow2 = np.sum(obs.astype(f32) ** 2, axis=-1, dtype=f32)
    lo = f32(alpha1) * obs_pow2 - lya_obs
    up = lya_obs - f32(alpha2) * obs_pow2
    loss2 = (np.maximum(lo, 0).mean(dtype=f32) + np.maximum(up, 0).mean(dtype=f32)) * f32(lya_positive_scale)
    start_norm = np.sqrt(np.sum(obs[:, 0, :].astype(f32) ** 2, axis=-1, dtype=f32)).astype(f32)
    expanded = start_norm[:, None] * son_coef[None, :]
    obs2_norm = np.sqrt(np.sum(obs2.astype(f32) ** 2, axis=-1, dtype=f32)).astype(f32)
    esl = np.where(expanded - obs2_norm >= 0, f32(1), f32(-1)).astype(f32)
    start_lya = lya_obs[:, 0]
    inner = esl * (lya_obs2 - start_lya[:, None] * sl_coef[None, :])
    term = isr * np.maximum(inner, 0)
    lya_diff = np.sum(diff_coef[None, :] * term, axis=1, dtype=f32)
    loss3 = lya_diff.mean(dtype=f32) * f32(lya_diff_scale)
    loss = f32(loss2 + loss3)
    # analytic gradients
    inv_bn = f32(1.0 / (B * n))
    g_obs = (-(lo > 0).astype(f32) + (up > 0).astype(f32)) * inv_bn * f32(lya_positive_scale)
    active = (inner > 0).astype(f32)
    w = diff_coef[None, :] * isr * active * esl * f32(lya_diff_scale / B)
    g_obs2 = w.astype(f32)
    g_obs = g_obs.copy()
    g_obs[:, 0] += -np.sum(w * sl_coef[None, :], axis=1, dtype=f32)
    return dict(loss=loss, loss2=f32(loss2), loss3=f32(loss3), is_clip=isr, esl=esl, obs_pow2=obs_pow2,
                grad_lya_obs=g_obs.astype(f32), grad_lya_obs2=g_obs2)


def stability_advantage(lya_obs0, lya_obs2, coefs):
    """msacl.py:392-400 -> raw [B] and batch-normalised [B] (unbiased std + 1e-8)."""
    _, diff_coef, sl_coef = coefs
    adv = lya_obs0[:, None] * sl_coef[None, :] - lya_obs2
    raw = np.sum(diff_coef[None, :] * adv, axis=1, dtype=f32)
    mean = raw.mean(dtype=f32)
    std = raw.std(ddof=1, dtype=f32)
    return raw.astype(f32), ((raw - mean) / (std + f32(1e-8))).astype(f32)


def clipped_surrogate(new_logp0, old_logp0, adv, clip_coef=0.1):
    """msacl.py:385-387,402-405 -> (loss_policy_lya, d loss / d new_logp0 [B])."""
    ratio = np.exp((new_logp0 - old_logp0).astype(f32)).astype(f32)
    surr1 = ratio * adv
    clipped = np.clip(ratio, f32(1 - clip_coef), f32(1 + clip_coef))
    surr2 = clipped * adv
    loss = np.minimum(surr1, surr2).mean(dtype=f32)
    B = adv.shape[0]
    inside = (ratio >= f32(1 - clip_coef)) & (ratio <= f32(1 + clip_coef))
    use1 = surr1 <= surr2
    g = np.where(use1, adv * ratio, np.where(inside, adv * ratio, f32(0))) / f32(B)
    return f32(loss), g.astype(f32)


def polyak_update(target_params, params, tau):
    """Soft target update (msacl.py:445-460): `p_targ.mul_(polyak); p_targ.add_((1 - polyak) * p)` with polyak = 1 - tau.
    float32 arrays in, new float32 target arrays out; the Python scalars are applied as float32 (torch casts a Python
    scalar to the tensor dtype) and every operation rounds to float32 separately."""
    polyak = 1 - tau
    pk, om = np.float32(polyak), np.float32(1 - polyak)
    return [(np.asarray(t, np.float32) * pk + om * np.asarray(p, np.float32)).astype(np.float32)
            for t, p in zip(target_params, params)]
