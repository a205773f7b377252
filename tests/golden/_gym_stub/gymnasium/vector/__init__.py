import numpy as np
from ..spaces import Box


def _batch_space(space, n):
    return Box(np.stack([space.low] * n), np.stack([space.high] * n), dtype=space.dtype)


class SyncVectorEnv:
    """gymnasium 0.28.1 SyncVectorEnv: sequential envs, same-step autoreset."""

    def __init__(self, env_fns, observation_space=None, action_space=None, copy=True):
        self.envs = [fn() for fn in env_fns]
        self.num_envs = len(self.envs)
        self.single_observation_space = self.envs[0].observation_space
        self.single_action_space = self.envs[0].action_space
        self.observation_space = _batch_space(self.single_observation_space, self.num_envs)
        self.action_space = _batch_space(self.single_action_space, self.num_envs)
        self.is_vector_env = True

    def reset(self, seed=None, options=None):
        if seed is None:
            seeds = [None] * self.num_envs
        elif isinstance(seed, int):
            seeds = [seed + i for i in range(self.num_envs)]
        else:
            seeds = list(seed)
        obs = []
        for env, s in zip(self.envs, seeds):
            o, _ = env.reset(seed=s, options=options)
            obs.append(o)
        return np.stack(obs).astype(self.single_observation_space.dtype), {}

    def step(self, actions):
        n = self.num_envs
        obs_out = []
        rewards = np.zeros(n, dtype=np.float64)
        terms = np.zeros(n, dtype=np.bool_)
        truncs = np.zeros(n, dtype=np.bool_)
        final_obs = np.full(n, None, dtype=object)
        has_final = np.zeros(n, dtype=np.bool_)
        for i, (env, a) in enumerate(zip(self.envs, actions)):
            o, r, te, tr, info = env.step(a)
            rewards[i], terms[i], truncs[i] = r, te, tr
            if te or tr:
                final_obs[i] = o
                has_final[i] = True
                o, info = env.reset()
            obs_out.append(o)
        infos = {}
        if has_final.any():
            infos["final_observation"] = final_obs
            infos["_final_observation"] = has_final
        return (np.stack(obs_out).astype(self.single_observation_space.dtype),
                rewards, terms, truncs, infos)

    def close(self):
        pass
