"""GPU parity of the general rollout engine: policies of arbitrary depth / width / activation (the reference builds any
MLP, RL/apprfunc/mlp.py:18-33) run as per-layer `msacl_gemm_tc` calls + one `msacl_rollout_step` launch per env step.

Tolerances (float32): the layers are split-bf16 (bf16x6) tensor-core GEMMs with FP32 accumulation, error ~7e-7 * sum|a||b|
per output, so the pre-tanh logits agree with the NumPy oracle to ~1e-5 relative: actions 2e-5 * range, log-prob as in
test_gpu_rollout.py (conditioned on tanh saturation); env outputs are compared with the oracle env stepped with the kernel's
own clipped action (next obs 1e-5 rel + 2e-5 abs, scaled reward / cost 1e-5 rel, 1e-4 QuadTracking).
"""
import numpy as np
import pytest
import torch

from oracle import actor as oactor
from oracle import envs as oenv
from oracle import rollout as oroll
from test_gpu_rollout import REW_TOL, _assert_logp, _sync_state

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,hidden,act", [("VanderPol", (64, 128, 32), "tanh"), ("TwoLink", (96,), "relu"),
                                             ("QuadTracking", (128, 64), "tanh"), ("SingleTrackCar", (300, 300), "relu")])
def test_general_step_vs_oracle_teacher_forced(name, hidden, act):
    from msacl_b200.sampler import FusedRollout, GeneralActor
    n, T, seed = 777, 10, 5
    spec = oenv.SPECS[name]
    w = oactor.init_policy_weights(spec.obs_dim, spec.act_dim, hidden=hidden, seed=2)
    mod = {"tanh": torch.nn.Tanh, "relu": torch.nn.ReLU}[act]
    hidden_act = {"tanh": lambda x: np.tanh(x).astype(np.float32), "relu": oactor.relu}[act]
    ga = GeneralActor(w, [mod() for _ in hidden] + [torch.nn.Identity()])
    ro = FusedRollout(name, n, 1, n_step=4, seed=seed, max_step=7)
    ro.state.reset()
    object.__setattr__(spec, "max_step", 7)
    try:
        ids = np.arange(n, dtype=np.uint64)
        venv = oroll.VectorEnv(name, oroll.philox_reset(name, seed, ids, np.zeros(n, np.int64)), seed=seed, env_ids=ids)
        emitter = oroll.WindowEmitter(n, 4)
        rng = np.random.default_rng(0)
        rngspan = spec.act_high - spec.act_low
        _sync_state(ro, name, venv.state)
        for t in range(T):
            eps = rng.standard_normal((n, spec.act_dim)).astype(np.float32)
            before = oroll.clone_state(venv.state)
            tr = oroll.sampler_step(venv, w, eps, hidden_act=hidden_act)
            emit, _ = emitter.push(tr)
            ro.run(ga, eps=torch.as_tensor(eps[None]).cuda())
            g = {k: v[ro.tr.H].cpu().numpy() for k, v in ro.tr.fields().items()}
            assert np.array_equal(g["obs"], tr["obs"])
            np.testing.assert_allclose(g["act"], tr["act"], rtol=0, atol=2e-5 * rngspan.max())
            _assert_logp(g["logp"], tr["logp"], tr["act"], spec)
            rew_k, cost_k, obs2_k, _, _ = oroll.env_outputs_for_action(name, before, g["act"])
            rt, at = REW_TOL[name == "QuadTracking"]
            np.testing.assert_allclose(g["obs2"], obs2_k, rtol=1e-5, atol=2e-5)
            np.testing.assert_allclose(g["rew"], rew_k, rtol=rt, atol=at)
            np.testing.assert_allclose(g["cost"], cost_k, rtol=rt, atol=at)
            near = (np.abs(tr["obs2"] - spec.obs_low) < 5e-3).any(1) | (np.abs(tr["obs2"] - spec.obs_high) < 5e-3).any(1)
            assert np.array_equal(g["done"].astype(bool)[~near], tr["done"][~near])
            assert np.array_equal(g["emit"].astype(bool), emit)
            _sync_state(ro, name, venv.state)
            ro.state.episode.copy_(torch.as_tensor(venv.episode.astype(np.int32)).cuda())
            ro.state.run.copy_(torch.as_tensor(emitter.run.astype(np.int32)).cuda())
        assert venv.episode.sum() > 0
    finally:
        object.__setattr__(spec, "max_step", 1000)


@pytest.mark.parametrize("name", ["Pendulum", "QuadTracking"])
def test_general_engine_matches_fused_engine_on_the_default_policy(name):
    """Same [256, 256] ReLU policy, same seeds, internal Philox noise: the unfused path (GEMM per layer + rollout_step)
    and the FP32 fused kernel are two implementations of the same K steps."""
    from msacl_b200.sampler import ActorWeights, FusedRollout, GeneralActor
    n, K = 1500, 12
    spec = oenv.SPECS[name]
    w = oactor.init_policy_weights(spec.obs_dim, spec.act_dim, seed=4)
    fa = ActorWeights(w)
    ga = GeneralActor(w, [torch.nn.ReLU(), torch.nn.ReLU(), torch.nn.Identity()])
    a = FusedRollout(name, n, K, n_step=5, seed=9, engine="ffma")
    b = FusedRollout(name, n, K, n_step=5, seed=9, engine="ffma")
    a.state.reset(); b.state.reset()
    a.run(fa); b.run(ga)
    fa_, fb_ = a.tr.fields(), b.tr.fields()
    span = float((spec.act_high - spec.act_low).max())
    # step 0 starts from identical states: actions agree to the actor tolerance; later steps drift with the closed loop
    assert (fa_["act"][a.tr.H] - fb_["act"][b.tr.H]).abs().max().item() <= 2e-5 * span
    assert torch.equal(fa_["obs"][a.tr.H], fb_["obs"][b.tr.H])
    assert torch.equal(fa_["emit"], fb_["emit"])
    tol = 2e-2 if name == "QuadTracking" else 5e-3
    assert (fa_["obs2"][a.tr.H:] - fb_["obs2"][b.tr.H:]).abs().max().item() <= tol
    assert torch.equal(a.state.step, b.state.step)
    np.testing.assert_allclose(a.stats[:5].cpu().numpy(), b.stats[:5].cpu().numpy(), rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("name,hidden,engine", [("Pendulum", (64, 64), "tc"), ("QuadTracking", (128, 32), "tc"), ("DuctedFan", (40, 200), "ffma")])
def test_narrow_relu_policy_runs_on_the_fused_kernels_zero_padded(name, hidden, engine):
    """Two ReLU hidden layers narrower than 256 are embedded in the fused kernels' 256-wide layers by zero padding: same logits
    as the narrow network (oracle), at fused-kernel speed."""
    from msacl_b200.sampler import ActorWeights, FusedRollout
    n, K, seed = 640, 5, 7
    spec = oenv.SPECS[name]
    w = oactor.init_policy_weights(spec.obs_dim, spec.act_dim, hidden=hidden, seed=5)
    aw = ActorWeights(w)
    assert aw.hidden_sizes == hidden
    ro = FusedRollout(name, n, K, n_step=3, seed=seed, engine=engine, record_logits=True)
    ro.state.reset()
    eps = np.random.default_rng(3).standard_normal((K, n, spec.act_dim)).astype(np.float32)
    ro.run(aw, eps=torch.as_tensor(eps).cuda())
    obs = ro.tr.obs[ro.tr.H:].cpu().numpy()
    logits = ro.tr.logits.cpu().numpy()
    tol = 5e-5 if engine == "tc" else 2e-5
    for k in range(K):          # the kernel's own observation through the narrow network (NumPy)
        want = oactor.mlp_forward(w, obs[k])
        assert np.abs(logits[k] - want).max() <= tol * max(np.abs(want).max(), 1.0)


def test_sampler_selects_the_general_engine_for_other_policies():
    """policy_hidden_sizes / activation other than the fused kernels' [256, 256] ReLU: the reference-compatible sampler
    falls back to the general engine (and refuses when a fused engine was requested by name)."""
    import msacl_b200
    from msacl_b200.algorithm import ApproxContainer
    from msacl_b200.sampler import DeviceWindowBatch, GeneralActor
    kw = dict(env_name="DuctedFan", env_num=64, sample_batch_size=8, reward_scale=100.0, cost_scale=100.0, noise_params=None,
              n_step=4, action_type="continu")
    spec = oenv.SPECS["DuctedFan"]
    nets = ApproxContainer(obs_dim=spec.obs_dim, act_dim=spec.act_dim, action_low_limit=spec.act_low, action_high_limit=spec.act_high,
                           policy_hidden_sizes=[64, 64, 64], policy_hidden_activation="gelu", q_learning_rate=1e-3,
                           lyapunov_learning_rate=1e-3, policy_learning_rate=1e-3, alpha_learning_rate=1e-3).cuda()
    s = msacl_b200.create_sampler(sampler_name="b200_nstep_off_sampler", networks=nets, **kw)
    for _ in range(2):
        batch, info = s.sample()
    assert isinstance(batch, DeviceWindowBatch) and isinstance(s._current_actor(), GeneralActor)
    assert 0 < batch.count() <= 64 * 8 and batch.count() == int(s.rollout.tr.emit[s.rollout.tr.H:].sum())
    tr = s.rollout.tr
    # the recorded action is the clipped tanh-Gauss sample of the torch policy's own output on the recorded observation
    with torch.no_grad():
        logits = nets.policy.policy(tr.obs[tr.H])
    mean = logits[:, :spec.act_dim]
    lo, hi = torch.as_tensor(spec.act_low).cuda(), torch.as_tensor(spec.act_high).cuda()
    assert torch.all(tr.act[tr.H] >= lo) and torch.all(tr.act[tr.H] <= hi)
    det = msacl_b200.create_sampler(sampler_name="b200_nstep_off_sampler", networks=nets, **kw)
    det.rollout.run(det._current_actor(), deterministic=True)
    dtr = det.rollout.tr
    with torch.no_grad():
        dl = nets.policy.policy(dtr.obs[dtr.H])
    mode = (hi - lo) / 2 * torch.tanh(dl[:, :spec.act_dim]) + (hi + lo) / 2
    assert (dtr.act[dtr.H] - mode).abs().max().item() <= 2e-5 * float((hi - lo).max())
    assert mean.shape == (64, spec.act_dim)
    with pytest.raises(ValueError):
        msacl_b200.create_sampler(sampler_name="b200_nstep_off_sampler", networks=nets, rollout_engine="tc", **kw).sample()
