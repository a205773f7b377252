"""Parity of the DEFAULT engine (tcgen05 split-bf16 actor) at the sizes BASELINE.json quotes:

    config 2  Pendulum / DuctedFan      65 536 envs on one GPU
    config 4  SingleTrackCar            2^22 envs (the single-GPU total; sharded over 2/4/8 in the scaling run)
    config 5  QuadTracking              2^21 envs per GPU
    config 3  TwoLink                   2^20 replay windows x n=20 through the MSACL target kernels

The fused rollout runs K steps at full size with its own Philox streams; a random subset of >= 4096 GLOBAL env ids
(incl. the first and the last tile) is then replayed through the NumPy oracle -- Philox is keyed by (seed, global env
id, step | episode), so any subset of a run is reproducible on the CPU:

  * logits (mean || log_std, the pre-tanh Gaussian parameters): the kernel's diagnostic output vs
    oracle.actor.mlp_forward on the kernel's own observation.  STATED TOLERANCE OF THE SPLIT-BF16 ENGINE:
    |d logit| <= 5e-5 * max|logit| over the batch (absolute, i.e. relative to the logit scale; measured ~1.5e-5).
  * action / log-prob: oracle TanhGauss sample from the oracle logits with the oracle's Philox draw -> 1e-4 * action range.
  * env outputs: the oracle env free-runs from the oracle's Philox reset and is stepped with the KERNEL's clipped action
    (removes the actor difference): next obs 1e-5 rel + 2e-5 abs per step (growing with the step index for the free-running
    quadrotor), scaled reward / cost as in test_gpu_rollout.REW_TOL, done flags exact away from the bounds, emit exact.
"""
import numpy as np
import pytest
import torch

from oracle import actor as oactor
from oracle import envs as oenv
from oracle import philox as ophx
from oracle import rollout as oroll
from oracle import targets as otg

pytestmark = pytest.mark.gpu

SUBSET = 4096
LOGIT_TOL = 5e-5


def _subset(n, rng):
    ids = np.unique(np.concatenate([np.arange(128), np.arange(n - 128, n), rng.choice(n, SUBSET, replace=False)]))
    return ids.astype(np.int64)


@pytest.mark.parametrize("name,n,env_base", [("Pendulum", 65536, 0), ("DuctedFan", 65536, 0), ("SingleTrackCar", 1 << 22, 0),
                                             ("QuadTracking", 1 << 21, 3 * (1 << 21))])
def test_tc_rollout_full_size_subset_replay(name, n, env_base):
    from msacl_b200.sampler import ActorWeights, FusedRollout
    K, seed, n_step = 4, 17, 3
    spec = oenv.SPECS[name]
    A = spec.act_dim
    w = oactor.init_policy_weights(spec.obs_dim, A, seed=2)
    ro = FusedRollout(name, n, K, n_step=n_step, seed=seed, env_base=env_base, engine="tc", history_chunks=1, record_logits=True)
    ro.state.reset()
    ro.global_step = 100                   # Philox step counter of the first step of this launch
    ro.run(ActorWeights(w))
    torch.cuda.synchronize()
    rng = np.random.default_rng(5)
    loc = _subset(n, rng)
    idx = torch.as_tensor(loc, device="cuda")
    got = {k: v[ro.tr.H:][:, idx].cpu().numpy() for k, v in ro.tr.fields().items()}
    logits = ro.tr.logits[:, idx].cpu().numpy()
    gids = (loc + env_base).astype(np.uint64)
    venv = oroll.VectorEnv(name, oroll.philox_reset(name, seed, gids, np.zeros(len(gids), np.int64)), seed=seed, env_ids=gids)
    span = float((spec.act_high - spec.act_low).max())
    from test_gpu_rollout import REW_TOL
    rt, at = REW_TOL[name == "QuadTracking"]
    alive = np.ones(len(gids), bool)      # envs whose done flags agreed so far (near-bound cases drop out)
    run = np.zeros(len(gids), np.int32)
    max_rel = 0.0
    for k in range(K):
        m = alive
        # the observation the kernel acted on == the oracle's current observation (reset: float32 round-off only)
        np.testing.assert_allclose(got["obs"][k][m], venv.obs[m], rtol=1e-5, atol=(2e-5 if name != "QuadTracking" else 2e-5 * (k + 1)))
        # ---- logits: stated tolerance of the tensor-core engine, on the kernel's own observation
        want_logits = oactor.mlp_forward(w, got["obs"][k])
        scale = np.abs(want_logits).max()
        err = np.abs(logits[k] - want_logits).max()
        max_rel = max(max_rel, err / scale)
        assert err <= LOGIT_TOL * scale, (k, err / scale)
        # ---- sampled action / log-prob from the oracle logits + the oracle's Philox draw
        eps = ophx.action_noise(seed, gids, 100 + k, A)
        mean, std = oactor.policy_forward(w, got["obs"][k])
        act, logp, _ = oactor.tanh_gauss_sample(mean, std, eps, spec.act_low, spec.act_high)
        act = np.clip(act, spec.act_low, spec.act_high)
        np.testing.assert_allclose(got["act"][k][m], act[m], rtol=0, atol=1e-4 * span)
        sat = (np.abs(act - (spec.act_high + spec.act_low) / 2) > 0.4995 * (spec.act_high - spec.act_low)).any(axis=1)
        ok = m & ~sat
        np.testing.assert_allclose(got["logp"][k][ok], logp[ok], rtol=1e-3, atol=5e-3)
        # ---- env: oracle stepped with the kernel's own clipped action
        rew_k, cost_k, obs2_k, term_k, trunc_k = oroll.env_outputs_for_action(name, venv.state, got["act"][k])
        np.testing.assert_allclose(got["obs2"][k][m], obs2_k[m], rtol=1e-5, atol=2e-5 * (k + 1))
        np.testing.assert_allclose(got["rew"][k][m], rew_k[m], rtol=rt * (k + 1), atol=at * (k + 1))
        np.testing.assert_allclose(got["cost"][k][m], cost_k[m], rtol=rt * (k + 1), atol=at * (k + 1))
        done_k = term_k | trunc_k
        near = (np.abs(obs2_k - spec.obs_low) < 1e-3).any(1) | (np.abs(obs2_k - spec.obs_high) < 1e-3).any(1)
        assert np.array_equal(got["done"][k].astype(bool)[m & ~near], done_k[m & ~near])
        alive = alive & (got["done"][k].astype(bool) == done_k)
        run = np.minimum(run + 1, n_step)
        assert np.array_equal(got["emit"][k].astype(bool)[alive], (run >= n_step)[alive])
        run = np.where(done_k, 0, run)
        venv.step(got["act"][k])          # advance the oracle (autoreset through the oracle's Philox restatement)
    assert alive.mean() > 0.995
    print(f"{name}: n={n} subset={len(gids)} max logits error / scale = {max_rel:.2e}")


def test_msacl_targets_full_size_subset_vs_oracle():
    """BASELINE config 3: the three MSACL target kernels over B = 2^20 TwoLink windows (n = 20); a random subset of 4096
    windows through the oracle (per-window outputs), and the batch-level sums through float64 reductions of the
    kernels' own per-window outputs."""
    from msacl_b200 import targets as tg
    B, n, D = 1 << 20, 20, 4
    g = torch.Generator(device="cuda").manual_seed(0)
    r = lambda *s: torch.randn(*s, device="cuda", generator=g)
    obs = r(B, n, D) * 0.4
    obs2 = obs + 0.05 * r(B, n, D)
    obs2[: B // 3] *= 0.6
    logp_new, logp_old = r(B, n) - 1.0, r(B, n) - 1.0
    v1, v2 = r(B, n).abs() * 0.5, r(B, n).abs() * 0.5
    rew, done = -torch.rand(B, n, device="cuda", generator=g) * 50, (torch.rand(B, n, device="cuda", generator=g) < 0.1).float()
    q1, q2, nlp = r(B, n), r(B, n), r(B, n)
    coef = tg.Coefficients(n)
    oc = otg.coefficients(n)
    idx = torch.as_tensor(np.random.default_rng(1).choice(B, SUBSET, replace=False), device="cuda")
    h = lambda t: t[idx].cpu().numpy()

    backup = tg.q_backup(rew, done, q1, q2, nlp, 0.99, 0.2)
    want = otg.q_backup(h(rew), h(done), h(q1), h(q2), h(nlp), 0.99, 0.2)
    np.testing.assert_allclose(h(backup), want, rtol=1e-6, atol=1e-6)

    out = tg.lyapunov_risk_raw(obs, obs2, logp_new, logp_old, v1, v2, coef, 10.0, 1.0, want_labels=True)
    sub = otg.lyapunov_risk(h(obs), h(obs2), h(logp_new), h(logp_old), h(v1), h(v2), oc, lya_diff_scale=10.0, lya_positive_scale=1.0)
    # ESL is a sign decision: exact wherever the margin is above float32 round-off
    o, o2 = h(obs), h(obs2)
    margin = np.abs(np.sqrt((o[:, 0] ** 2).sum(-1))[:, None] * oc[0][None] - np.sqrt((o2 ** 2).sum(-1)))
    ok = margin > 1e-5
    assert np.array_equal(h(out["esl"])[ok], sub["esl"][ok]) and ok.mean() > 0.999
    np.testing.assert_allclose(h(out["is_clip"]), sub["is_clip"], rtol=1e-5, atol=1e-7)
    # the loss over the subset alone (same kernel, B = 4096) against the oracle's scalar
    small = tg.lyapunov_risk_raw(obs[idx], obs2[idx], logp_new[idx], logp_old[idx], v1[idx], v2[idx], coef, 10.0, 1.0)
    np.testing.assert_allclose(float(small["loss"]), float(sub["loss"]), rtol=1e-5)
    # gradients are per window x step: full-size launch and subset launch must agree up to the batch-size factor
    np.testing.assert_allclose(h(out["grad_lya_obs2"]) * (B / SUBSET), small["grad_lya_obs2"].cpu().numpy(), rtol=1e-6, atol=1e-12)

    raw, adv = tg.stability_advantage(v1[:, 0].contiguous(), v2, coef)
    want_raw, _ = otg.stability_advantage(h(v1)[:, 0], h(v2), oc)
    np.testing.assert_allclose(h(raw), want_raw, rtol=2e-6, atol=1e-6)
    mean, std = raw.double().mean(), raw.double().std(unbiased=True)
    np.testing.assert_allclose(h(adv), ((raw[idx].double() - mean) / (std + 1e-8)).cpu().numpy(), rtol=1e-5, atol=1e-5)
