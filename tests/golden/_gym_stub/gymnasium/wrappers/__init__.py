import time
import numpy as np
from .. import Wrapper


class RecordEpisodeStatistics(Wrapper):
    def __init__(self, env, deque_size=100):
        super().__init__(env)
        self.episode_returns = np.zeros(1, dtype=np.float32)
        self.episode_lengths = np.zeros(1, dtype=np.int32)
        self.episode_count = 0

    def reset(self, **kwargs):
        out = self.env.reset(**kwargs)
        self.episode_returns = np.zeros(1, dtype=np.float32)
        self.episode_lengths = np.zeros(1, dtype=np.int32)
        return out

    def step(self, action):
        o, r, te, tr, info = self.env.step(action)
        self.episode_returns += r
        self.episode_lengths += 1
        if te or tr:
            info = dict(info)
            info["episode"] = {"r": self.episode_returns.copy(), "l": self.episode_lengths.copy(),
                               "t": np.array([time.perf_counter()], dtype=np.float32)}
            self.episode_count += 1
            self.episode_returns[:] = 0
            self.episode_lengths[:] = 0
        return o, r, te, tr, info
