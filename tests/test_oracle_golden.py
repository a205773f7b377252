"""Pins the NumPy oracle against golden vectors produced by running the reference itself
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import load_golden, quad_state
from oracle import actor, envs, rollout, targets

BOX = envs.ENV_NAMES[:5]


@pytest.mark.parametrize("name", BOX)
def test_box_env_step_bit_exact(name):
    g = load_golden(f"env_step_{name}.npz")
    new, obs, rew, term, trunc = envs.env_step(name, {"obs": g["obs_in"], "step": g["step_in"]}, g["act"])
    # the oracle reproduces the reference's NumPy dtype flow, so float32 results are identical
    assert np.array_equal(obs, g["obs_out"])
    assert np.array_equal(rew, g["reward"])
    assert np.array_equal(term, g["term"]) and np.array_equal(trunc, g["trunc"])
    assert term.any() and trunc.any() and (~term).any()


def test_quad_env_step_bit_exact():
    g = load_golden("env_step_QuadTracking.npz")
    new, obs, rew, term, trunc = envs.env_step("QuadTracking", quad_state(g, "in"), g["act"])
    for k in ("x", "v", "R", "Om", "obs"):
        assert np.array_equal(new[k], g[f"{k}_out"]), k
    assert np.array_equal(new["t"], g["t_out"]) and np.array_equal(new["step"], g["step_out"])
    np.testing.assert_allclose(new["Rd_last"], g["Rd_last_out"], rtol=0, atol=1e-15)   # f64 BLAS order
    assert np.array_equal(rew, g["reward"])
    assert np.array_equal(term, g["term"]) and np.array_equal(trunc, g["trunc"])


def test_quad_reset_observation():
    g = load_golden("env_step_QuadTracking.npz")
    st = envs.quad_state_from_raw(g["x_reset"], g["v_reset"], g["R_reset"], g["Om_reset"])
    assert np.array_equal(st["obs"], g["obs_reset"])
    np.testing.assert_allclose(st["Rd_last"], g["Rd_last_reset"], rtol=0, atol=1e-15)


def _weights(g):
    return [(g[f"W{i}"], g[f"b{i}"]) for i in range(3)]


def _init_state(name, g, prefix="init_"):
    keys = ("x", "v", "R", "Om", "t", "t_last", "Rd_last", "obs", "step") if name == "QuadTracking" else ("obs", "step")
    return {k: g[prefix + k] for k in keys}


@pytest.mark.parametrize("name", envs.ENV_NAMES)
def test_sampler_pipeline_matches_reference(name):
    """Replays 40 reference `_n_step()` calls (env_num=4, n_step=5, max_step patched to 13):
    same weights, same N(0,1) draws, reference post-reset states injected on done."""
    g = load_golden(f"sampler_{name}.npz")
    n_step, ring_size = int(g["n_step"]), int(g["ring"])
    spec = envs.SPECS[name]
    T, N = g["step_eps"].shape[:2]
    post = {k[5:]: g[k] for k in g if k.startswith("post_")}
    t_box = {"t": 0}

    def reset_fn(done, episode):
        return {k: v[t_box["t"]] for k, v in post.items()}

    # the golden run shortened episodes via the instance attribute max_step
    object.__setattr__(spec, "max_step", int(g["max_step"]))
    try:
        venv = rollout.VectorEnv(name, _init_state(name, g), reset_fn=reset_fn)
        emitter = rollout.WindowEmitter(N, n_step)
        ring = rollout.ReplayRing(ring_size, n_step, spec.obs_dim, spec.act_dim)
        wi = 0
        max_dlogp = 0.0
        for t in range(T):
            t_box["t"] = t
            # teacher forcing: start every step from the reference's own state
            assert np.array_equal(venv.obs, g["step_obs"][t])
            tr = rollout.sampler_step(venv, _weights(g), g["step_eps"][t])
            # actor: BLAS summation order differs (torch MKL vs NumPy OpenBLAS) -> tolerance
            np.testing.assert_allclose(tr["act"], g["step_act"][t], rtol=2e-5, atol=2e-5)
            v = g["step_valid"][t]
            np.testing.assert_allclose(tr["logp"][v], g["step_logp"][t][v], rtol=1e-4, atol=2e-4)
            max_dlogp = max(max_dlogp, np.abs(tr["logp"][v] - g["step_logp"][t][v]).max(initial=0))
            # window logic is checked on the reference's own per-step values (exact compare)
            emit, wins = emitter.push(dict(tr, act=g["step_act"][t], logp=np.nan_to_num(g["step_logp"][t]),
                                           rew=np.nan_to_num(g["step_rew"][t]), cost=np.nan_to_num(g["step_cost"][t]),
                                           obs2=g["step_obs2"][t], done=g["step_done"][t] > 0))
            assert np.array_equal(emit, g["step_emit"][t])
            for w in wins:
                for k in rollout.WindowEmitter.FIELDS:
                    assert np.array_equal(w[k], g["win_" + k][wi]), (t, k)
                wi += 1
            ring.add_batch(wins)
            assert ring.ptr == g["ptr_after"][t] and ring.size == g["size_after"][t]
            # resync env state to the reference (actor rounding must not accumulate)
            venv.state = {k: post[k][t].copy() for k in post}
        assert wi == len(g["win_rew"])
        for k in ring.buf:
            assert np.array_equal(ring.buf[k], g["ring_" + k]), k
    finally:
        object.__setattr__(spec, "max_step", 1000)


@pytest.mark.parametrize("name", envs.ENV_NAMES)
def test_sampler_env_side_exact_given_reference_action(name):
    """With the reference's own clipped action, the vector step (autoreset, final_observation,
    reward/cost scaling, done flags) must be bit-identical."""
    g = load_golden(f"sampler_{name}.npz")
    spec = envs.SPECS[name]
    T, N = g["step_eps"].shape[:2]
    post = {k[5:]: g[k] for k in g if k.startswith("post_")}
    object.__setattr__(spec, "max_step", int(g["max_step"]))
    try:
        state = _init_state(name, g)
        n_done = 0
        for t in range(T):
            venv = rollout.VectorEnv(name, state, reset_fn=lambda d, e: {k: v[t] for k, v in post.items()})
            next_obs, reward, term, trunc, final_obs, _ = venv.step(g["step_act"][t])
            done = term | trunc
            real_next = np.where(done[:, None], final_obs, next_obs)
            assert np.array_equal(term, g["step_term"][t]) and np.array_equal(trunc, g["step_trunc"][t])
            assert np.array_equal(reward, g["step_raw_reward"][t])
            assert np.array_equal(real_next, g["step_obs2"][t])
            assert np.array_equal(next_obs, g["step_next_obs"][t])
            v = g["step_valid"][t]
            rew = (reward * 100.0).astype(np.float32)
            cost = (envs.np_pairwise_rowsum(real_next ** 2) * 100.0).astype(np.float32)
            assert np.array_equal(rew[v], g["step_rew"][t][v])
            assert np.array_equal(cost[v], g["step_cost"][t][v])
            n_done += int(done.sum())
            state = {k: post[k][t].copy() for k in post}
        assert n_done > 0
    finally:
        object.__setattr__(spec, "max_step", 1000)


def test_msacl_coefficients():
    g = load_golden("msacl_targets_TwoLink.npz")
    son, diff, sl = targets.coefficients(20)
    np.testing.assert_allclose(son, g["coef_start_obs_norm"], rtol=2e-6)
    np.testing.assert_allclose(diff, g["coef_lya_diff"], rtol=2e-6)
    np.testing.assert_allclose(sl, g["coef_start_lya"], rtol=2e-6)


def test_msacl_tanh_gauss_log_prob():
    g = load_golden("msacl_targets_TwoLink.npz")
    lp = actor.tanh_gauss_log_prob(g["pi_mean"], g["pi_std"], g["act"], g["act_low"], g["act_high"])
    np.testing.assert_allclose(lp, g["logp_new"], rtol=1e-5, atol=1e-4)


def test_msacl_lyapunov_risk_matches_reference_loss():
    g = load_golden("msacl_targets_TwoLink.npz")
    coefs = (g["coef_start_obs_norm"], g["coef_lya_diff"], g["coef_start_lya"])
    out = targets.lyapunov_risk(g["obs"], g["obs2"], g["logp_new"], g["logp"], g["lya_obs"], g["lya_obs2"], coefs)
    np.testing.assert_allclose(out["loss"], g["loss_lya"], rtol=2e-6)
    assert (out["esl"] > 0).any() and (out["esl"] < 0).any()


def test_msacl_lyapunov_gradients_match_reference_param_grads():
    torch = pytest.importorskip("torch")
    g = load_golden("msacl_targets_TwoLink.npz")
    coefs = (g["coef_start_obs_norm"], g["coef_lya_diff"], g["coef_start_lya"])
    out = targets.lyapunov_risk(g["obs"], g["obs2"], g["logp_new"], g["logp"], g["lya_obs"], g["lya_obs2"], coefs)
    # rebuild the Lyapunov net (D->256->256->256, tanh, V = sum of squares: mlp.py:72-88) in plain torch
    names = [str(n) for n in g["lya_param_names"]]
    params = {n: torch.tensor(g["lya_param_" + n], requires_grad=True) for n in names}

    def V(x):
        h = torch.tanh(x @ params["lya.0.weight"].T + params["lya.0.bias"])
        h = torch.tanh(h @ params["lya.2.weight"].T + params["lya.2.bias"])
        o = h @ params["lya.4.weight"].T + params["lya.4.bias"]
        return (o ** 2).sum(-1)

    # the reference evaluates V(obs) twice (msacl.py:289,317); both feed the same parameters
    v1 = V(torch.tensor(g["obs"])); v2 = V(torch.tensor(g["obs2"]))
    torch.autograd.backward([v1, v2], [torch.tensor(out["grad_lya_obs"]), torch.tensor(out["grad_lya_obs2"])])
    for n in names:
        ref = g["lya_grad_" + n]
        np.testing.assert_allclose(params[n].grad.numpy(), ref, rtol=2e-4, atol=2e-6 * max(1.0, np.abs(ref).max()))


def test_msacl_q_backup_matches_reference_loss():
    g = load_golden("msacl_targets_TwoLink.npz")
    backup = targets.q_backup(g["rew"], g["done"], g["next_q1"], g["next_q2"], g["next_logp"], float(g["gamma"]), float(g["alpha"]))
    loss = np.mean((g["q1"] - backup) ** 2, dtype=np.float32) + np.mean((g["q2"] - backup) ** 2, dtype=np.float32)
    np.testing.assert_allclose(loss, g["loss_q"], rtol=3e-6)


def test_msacl_policy_loss_matches_reference():
    g = load_golden("msacl_targets_TwoLink.npz")
    coefs = (g["coef_start_obs_norm"], g["coef_lya_diff"], g["coef_start_lya"])
    raw, adv = targets.stability_advantage(g["pol_lya_obs0"], g["pol_lya_obs2"], coefs)
    loss_lya, grad = targets.clipped_surrogate(g["pol_new_logp0"], g["logp"][:, 0], adv)
    loss = -g["policy_q_term"] - loss_lya
    np.testing.assert_allclose(loss, g["loss_policy"], rtol=1e-5, atol=1e-5)
    assert np.isfinite(grad).all()


def test_oracle_polyak_matches_torch_ops():
    """oracle.targets.polyak_update vs the reference's literal torch statements (msacl.py:445-460) on CPU tensors."""
    import torch
    from oracle import targets as otg
    g = torch.Generator().manual_seed(1)
    ps = [torch.randn(33, 7, generator=g), torch.randn(5, generator=g)]
    ts = [torch.randn(33, 7, generator=g), torch.randn(5, generator=g)]
    for tau in (0.005, 0.3):
        want = [t.clone() for t in ts]
        polyak = 1 - tau
        for p, pt in zip(ps, want):
            pt.data.mul_(polyak)
            pt.data.add_((1 - polyak) * p.data)
        got = otg.polyak_update([t.numpy() for t in ts], [p.numpy() for p in ps], tau)
        for w, gq in zip(want, got):
            assert np.array_equal(w.numpy(), gq)


def test_quad_polar_incl_reflection_branch():
    """NormalizeOrientMatrix (QuadTracking.py:308-315) incl. its det < 0 column flip (:312-314)."""
    g = load_golden("quad_polar.npz")
    assert (g["det_in"] < 0).sum() >= 40 and (g["det_in"] > 0).sum() >= 40
    got = envs._polar_svd(g["mat_in"])
    np.testing.assert_allclose(got, g["mat_out"], rtol=0, atol=2e-7)
    assert np.all(np.linalg.det(got.astype(np.float64)) > 0.99)


@pytest.mark.parametrize("name", ["VanderPol", "TwoLink", "QuadTracking"])
def test_evaluator_matches_reference(name):
    """oracle.evaluator vs a recorded run of the reference Evaluator.run_parallel_episodes (evaluator.py:141-204)."""
    from oracle import evaluator as oeval
    g = load_golden(f"evaluator_{name}.npz")
    spec = envs.SPECS[name]
    object.__setattr__(spec, "max_step", int(g["max_step"]))
    try:
        got, lens = oeval.run_parallel_episodes(name, _weights(g), _init_state(name, g), 0, float(g["reward_scale"]),
                                                float(g["cost_scale"]), return_lengths=True)
    finally:
        object.__setattr__(spec, "max_step", 1000)
    assert np.array_equal(lens, g["first_episode_len"])
    want = (float(g["trm"]), float(g["trs"]), float(g["tcm"]), float(g["tcs"]))
    np.testing.assert_allclose(got, want, rtol=2e-4)
