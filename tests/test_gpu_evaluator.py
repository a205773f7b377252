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
    spec = oenv.SPECS[name]
    n = state["obs"].shape[0]
    venv = oroll.VectorEnv(name, state, seed=seed, env_ids=np.arange(n, dtype=np.uint64))
    rets, costs = [[] for _ in range(n)], [[] for _ in range(n)]
    finished = np.zeros(n, bool)
    while not finished.all():
        obs = venv.obs.astype(np.float32)
        mean, _ = oactor.policy_forward(weights, obs)
        act = oactor.tanh_gauss_mode(mean, spec.act_low, spec.act_high)          # no clip in the evaluator (:158-160)
        next_obs, reward, term, trunc, final_obs, _ = venv.step(act)
        done = term | trunc
        real_next = np.where(done[:, None], final_obs, next_obs)
        rew = reward.astype(np.float32) * reward_scale
        cost = oenv.np_pairwise_rowsum(real_next.astype(np.float32) ** 2) * cost_scale
        for i in range(n):
            if not finished[i]:
                rets[i].append(rew[i]); costs[i].append(cost[i]); finished[i] = done[i]
    er, ec = [np.mean(r) for r in rets], [np.mean(c) for c in costs]
    return np.mean(er), np.std(er), np.mean(ec), np.std(ec)


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
