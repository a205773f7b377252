"""Oracle (test infrastructure): the sampler-side pipeline in NumPy.

Restates, vectorised over env instances:
  * gymnasium 0.28.1 SyncVectorEnv same-step autoreset + info["final_observation"]
    (third-party; call sites RL/create_pkg/create_envs.py:24, RL/trainer/sampler/base.py:148,160)
  * BaseSampler._n_step (RL/trainer/sampler/base.py:118-222): actor -> sample -> clip ->
    envs.step -> real_next_obs -> rew_plus_cost (RL/utils/rew_plus_cost.py:18-21) ->
    per-env deque(maxlen=n) sliding windows, emitted when full, cleared on done
  * NstepReplayBuffer store / sample_batch (RL/trainer/buffer/nstep_replay_buffer.py:91-150)
"""
from __future__ import annotations

import numpy as np

from . import actor, envs, philox

f32 = np.float32


def clone_state(s):
    return {k: v.copy() for k, v in s.items()}


def select_state(mask, a, b):
    """Row-wise where(mask, a, b) over a state dict."""
    out = {}
    for k in a:
        m = mask.reshape((-1,) + (1,) * (a[k].ndim - 1))
        out[k] = np.where(m, a[k], b[k]).astype(a[k].dtype)
    return out


def philox_reset(name, seed, env_ids, episode):
    """Production reset distribution (see oracle.envs / oracle.philox)."""
    if name == "QuadTracking":
        u9, z3 = philox.quad_reset_draws(seed, env_ids, episode)
        return envs.quad_reset_from_draws(u9, z3)
    spec = envs.SPECS[name]
    return envs.box_reset_from_uniform(name, philox.box_reset_uniforms(seed, env_ids, episode, spec.obs_dim))


class VectorEnv:
    """N instances of one env with same-step autoreset.  `reset_fn(done_mask, episode)`
    returns a full state dict of fresh states (rows where done_mask is False are ignored)."""

    def __init__(self, name, state, reset_fn=None, seed=0, env_ids=None):
        self.name = name
        self.spec = envs.SPECS[name]
        self.state = clone_state(state)
        n = state["obs"].shape[0]
        self.env_ids = np.arange(n, dtype=np.uint64) if env_ids is None else np.asarray(env_ids, np.uint64)
        self.episode = np.zeros(n, dtype=np.int64)      # index of the CURRENT episode
        self.seed = seed
        self.reset_fn = reset_fn or (lambda done, ep: philox_reset(name, self.seed, self.env_ids, ep))
        self.ep_return = np.zeros(n, dtype=np.float64)
        self.ep_length = np.zeros(n, dtype=np.int64)

    @property
    def obs(self):
        return self.state["obs"]

    def step(self, action):
        new, obs, reward, term, trunc = envs.env_step(self.name, self.state, action)
        done = term | trunc
        final_obs = obs.copy()
        self.ep_return += reward
        self.ep_length += 1
        stats = dict(done=done.copy(), ret=np.where(done, self.ep_return, 0.0), length=np.where(done, self.ep_length, 0))
        if done.any():
            self.episode = self.episode + done.astype(np.int64)
            fresh = self.reset_fn(done, self.episode)
            new = select_state(done, fresh, new)
            self.ep_return = np.where(done, 0.0, self.ep_return)
            self.ep_length = np.where(done, 0, self.ep_length)
        self.state = new
        return new["obs"].copy(), reward, term, trunc, final_obs, stats


def sampler_step(venv: VectorEnv, weights, eps, reward_scale=100.0, cost_scale=100.0, hidden_act=actor.relu):
    """base.py:124-163 for one vector step; returns the transition dict (all f32 / bool)."""
    spec = venv.spec
    obs = venv.obs.astype(f32).copy()
    mean, std = actor.policy_forward(weights, obs, hidden_act=hidden_act)
    act, logp, _ = actor.tanh_gauss_sample(mean, std, eps, spec.act_low, spec.act_high)
    act_clip = np.clip(act, spec.act_low, spec.act_high).astype(f32)
    next_obs, reward, term, trunc, final_obs, stats = venv.step(act_clip)
    done = term | trunc
    real_next_obs = np.where(done[:, None], final_obs, next_obs).astype(f32)
    rew = (reward.astype(f32) * reward_scale).astype(f32)                  # rew_plus_cost.py:18
    cost = (envs.np_pairwise_rowsum(real_next_obs ** 2) * cost_scale).astype(f32)   # :20-21
    return dict(obs=obs, act=act_clip, rew=rew, cost=cost, obs2=real_next_obs,
                done=done, term=term, trunc=trunc, logp=logp.astype(f32), next_obs=next_obs,
                mean=mean, std=std, stats=stats)


def env_outputs_for_action(name, state, act_clip, reward_scale=100.0, cost_scale=100.0):
    """The env half of base.py:148-163 for a GIVEN clipped action (no actor): (scaled reward, scaled cost, real_next_obs,
    term, trunc) from `state`, which is not modified.  Parity tests feed it the kernel's own action so that the
    reward / cost comparison is free of the actor's summation-order difference."""
    _, obs, reward, term, trunc = envs.env_step(name, clone_state(state), np.asarray(act_clip, f32))
    rew = (reward.astype(f32) * f32(reward_scale)).astype(f32)
    cost = (envs.np_pairwise_rowsum(obs.astype(f32) ** 2) * f32(cost_scale)).astype(f32)
    return rew, cost, obs.astype(f32), term, trunc


class WindowEmitter:
    """base.py:95,178-217: per-env deque(maxlen=n); after appending the newest transition a
    window [n, .] is emitted iff the deque is full; the deque is cleared when the newest
    transition is done.  Emission order within a step is env order."""

    FIELDS = ("obs", "act", "rew", "cost", "obs2", "done", "logp")

    def __init__(self, num_envs, n_step):
        self.n = n_step
        self.hist = []                                   # last n transitions (dicts of arrays)
        self.run = np.zeros(num_envs, dtype=np.int32)    # deque length, capped at n

    def push(self, tr):
        self.hist.append({k: np.asarray(tr[k]).copy() for k in self.FIELDS})
        if len(self.hist) > self.n:
            self.hist.pop(0)
        self.run = np.minimum(self.run + 1, self.n)
        emit = self.run >= self.n
        windows = []
        for i in np.nonzero(emit)[0]:
            w = {}
            for k in self.FIELDS:
                w[k] = np.stack([h[k][i] for h in self.hist], axis=0).astype(f32)
            windows.append(w)
        self.run = np.where(tr["done"], 0, self.run).astype(np.int32)
        return emit, windows


class ReplayRing:
    """nstep_replay_buffer.py:43-150 (arrays [max_size, n, .], ptr/size arithmetic)."""

    def __init__(self, max_size, n_step, obs_dim, act_dim):
        self.max_size, self.n = max_size, n_step
        z = lambda *s: np.zeros((max_size, n_step, *s), dtype=f32)
        self.buf = dict(obs=z(obs_dim), act=z(act_dim), rew=z(), cost=z(), obs2=z(obs_dim), done=z(), logp=z())
        self.ptr = 0
        self.size = 0

    def store(self, w):
        for k in self.buf:
            self.buf[k][self.ptr] = w[k]
        self.ptr = (self.ptr + 1) % self.max_size
        self.size = min(self.size + 1, self.max_size)

    def add_batch(self, windows):
        for w in windows:
            self.store(w)

    def gather(self, idx):
        return {k: v[idx].copy() for k, v in self.buf.items()}
