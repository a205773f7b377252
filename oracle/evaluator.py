"""Oracle (test infrastructure): Evaluator.run_parallel_episodes in NumPy
(RL/trainer/evaluator.py:141-204), built on the oracle vector env and actor.

Every instance is stepped with the distribution's mode() action (act_distribution_cls.py:90-95; the
evaluator does not clip, :158-160) until it has finished its FIRST episode; scaled reward and cost
(RL/utils/rew_plus_cost.py:18-21) are averaged over the steps of that episode per instance, then mean
and population std over instances are returned (TRM, TRS, TCM, TCS).  Pinned against a run of the
reference evaluator by tests/test_oracle_golden.py::test_evaluator_matches_reference.
"""
import numpy as np

from . import actor, envs, rollout

f32 = np.float32


def run_parallel_episodes(name, weights, state, seed=0, reward_scale=100.0, cost_scale=100.0, return_lengths=False):
    spec = envs.SPECS[name]
    n = state["obs"].shape[0]
    venv = rollout.VectorEnv(name, state, seed=seed, env_ids=np.arange(n, dtype=np.uint64))
    rets, costs = [[] for _ in range(n)], [[] for _ in range(n)]
    finished = np.zeros(n, bool)
    while not finished.all():
        obs = venv.obs.astype(f32)
        mean, _ = actor.policy_forward(weights, obs)
        act = actor.tanh_gauss_mode(mean, spec.act_low, spec.act_high)
        next_obs, reward, term, trunc, final_obs, _ = venv.step(act)
        done = term | trunc
        real_next = np.where(done[:, None], final_obs, next_obs)
        rew = reward.astype(f32) * f32(reward_scale)
        cost = envs.np_pairwise_rowsum(real_next.astype(f32) ** 2) * f32(cost_scale)
        for i in range(n):
            if not finished[i]:
                rets[i].append(rew[i]); costs[i].append(cost[i]); finished[i] = done[i]
    er, ec = [np.mean(r) for r in rets], [np.mean(c) for c in costs]
    out = (np.mean(er), np.std(er), np.mean(ec), np.std(ec))
    if return_lengths:
        return out, np.array([len(r) for r in rets], np.int32)
    return out
