"""B200Evaluator vs a NumPy restatement of Evaluator.run_parallel_episodes (RL/trainer/evaluator.py:141-204)
built on the oracle: same initial states, same weights, greedy (mode) actions, first episode only."""
import numpy as np
import pytest
import torch

from oracle import actor as oactor
from oracle import envs as oenv
from oracle import rollout as oroll

pytestmark = pytest.mark.gpu


def oracle_evaluate(name, weights, state, seed, max_step, reward_scale=100.0, cost_scale=100.0):
    from oracle import evaluator as oeval
    return oeval.run_parallel_episodes(name, weights, state, seed, reward_scale, cost_scale)


@pytest.mark.parametrize("name,engine", [("VanderPol", "ffma"), ("DuctedFan", "tc"), ("QuadTracking", "ffma")])
def test_evaluator_matches_oracle(name, engine):
    from msacl_b200.evaluator import B200Evaluator
    from msacl_b200.sampler import ActorWeights
    spec = oenv.SPECS[name]
    n, max_step, seed = 48, 30, 4
    w = oactor.init_policy_weights(spec.obs_dim, spec.act_dim, seed=3)
    ev = B200Evaluator(env_name=name, num_eval_episode=n, reward_scale=100.0, cost_scale=100.0, eval_env_seed=seed,
                       max_step=max_step, eval_chunk_steps=7, rollout_engine=engine)
    object.__setattr__(spec, "max_step", max_step)
    try:
        ost = oroll.philox_reset(name, seed, np.arange(n, dtype=np.uint64), np.zeros(n, np.int64))

        def init(state):       # same initial states on both sides (device reset == oracle reset up to round-off)
            if name == "QuadTracking":
                state.set_quad_state(ost["x"], ost["v"], ost["R"], ost["Om"], t=ost["t"], Rd_last=ost["Rd_last"], obs=ost["obs"], step=ost["step"])
            else:
                state.set_box_state(ost["obs"], ost["step"])

        got = ev.run_parallel_episodes(ActorWeights(w), state_init=init)
        want = oracle_evaluate(name, w, ost, seed, max_step)
    finally:
        object.__setattr__(spec, "max_step", 1000)
    # greedy closed-loop rollouts of up to 30 steps: trajectories agree to ~1e-4 relative
    np.testing.assert_allclose(got, want, rtol=2e-3, atol=1e-3 * max(1.0, abs(want[0])))


@pytest.mark.parametrize("name", ["VanderPol", "TwoLink", "QuadTracking"])
@pytest.mark.parametrize("engine", ["tc", "ffma"])
def test_evaluator_matches_reference_golden(name, engine):
    """The reference Evaluator's own run (tests/golden/evaluator_*.npz: its policy weights, its initial env states, its
    TRM / TRS / TCM / TCS) reproduced by B200Evaluator on the fused kernel.  Free-running greedy closed loop of up to 60
    steps: 2e-3 relative (split-bf16 actor 1e-4 * action range per step, amplified by the closed loop)."""
    from conftest import load_golden
    from msacl_b200.evaluator import B200Evaluator
    from msacl_b200.sampler import ActorWeights
    g = load_golden(f"evaluator_{name}.npz")
    n, max_step = int(g["episodes"]), int(g["max_step"])
    ev = B200Evaluator(env_name=name, num_eval_episode=n, reward_scale=float(g["reward_scale"]), cost_scale=float(g["cost_scale"]),
                       eval_env_seed=0, max_step=max_step, eval_chunk_steps=16, rollout_engine=engine)

    def init(state):
        if name == "QuadTracking":
            state.set_quad_state(g["init_x"], g["init_v"], g["init_R"], g["init_Om"], t=g["init_t"], Rd_last=g["init_Rd_last"],
                                 obs=g["init_obs"], step=g["init_step"])
        else:
            state.set_box_state(g["init_obs"], g["init_step"])

    got = ev.run_parallel_episodes(ActorWeights([(g[f"W{i}"], g[f"b{i}"]) for i in range(3)]), state_init=init)
    want = (float(g["trm"]), float(g["trs"]), float(g["tcm"]), float(g["tcs"]))
    np.testing.assert_allclose(got, want, rtol=2e-3, atol=1e-3 * abs(want[0]))
    assert (ev.last_first_episode_len.cpu().numpy() == g["first_episode_len"]).all()
