"""Minimal stand-in for gymnasium==0.28.1, used ONLY by tests/golden/make_golden.py.

The reference (`/root/reference/RL/env/*.py`) does `import gymnasium as gym`; gymnasium is
not installed in this image, so the golden-vector generator puts this directory on
sys.path. It restates the published 0.28.1 semantics the reference relies on:
  * Env.reset(seed) -> np_random = Generator(PCG64(SeedSequence(seed)))
  * spaces.Box casting low/high to dtype
  * vector.SyncVectorEnv same-step autoreset with info["final_observation"]
  * wrappers.RecordEpisodeStatistics episode return/length accumulators
It is test infrastructure and is never imported by the product package.
"""
import numpy as np


class Env:
    observation_space = None
    action_space = None
    _np_random = None

    @property
    def np_random(self):
        if self._np_random is None:
            self._np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence()))
        return self._np_random

    @np_random.setter
    def np_random(self, v):
        self._np_random = v

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self._np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        return None

    @property
    def unwrapped(self):
        return self

    def close(self):
        pass


class Wrapper(Env):
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def observation_space(self):
        return self.env.observation_space

    @property
    def action_space(self):
        return self.env.action_space

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def step(self, action):
        return self.env.step(action)


from . import spaces, vector, wrappers  # noqa: E402,F401  (after Env/Wrapper: wrappers imports them)
