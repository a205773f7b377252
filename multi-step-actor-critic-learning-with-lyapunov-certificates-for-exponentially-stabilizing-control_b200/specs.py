"""Static description of the six environments (names, ids, dims, Box bounds).

Mirrors the spaces the reference env classes build in their constructors
(RL/env/VanderPol.py:31-45, Pendulum.py:31-41, DuctedFan.py:34-45, TwoLink.py:42-53,
SingleTrackCar.py:76-87, QuadTracking.py:80-104) so that host code can size buffers without a
GPU; `tests/test_gpu_envs.py` cross-checks this table against the device constants.
"""
import math
from dataclasses import dataclass

import numpy as np

ENV_NAMES = ("VanderPol", "Pendulum", "DuctedFan", "TwoLink", "SingleTrackCar", "QuadTracking")
ENV_IDS = {name: i for i, name in enumerate(ENV_NAMES)}


@dataclass(frozen=True)
class EnvSpec:
    name: str
    env_id: int
    obs_dim: int
    act_dim: int
    sf_rows: int
    sd_rows: int
    obs_off: int
    control_step: int
    obs_low: np.ndarray
    obs_high: np.ndarray
    act_low: np.ndarray
    act_high: np.ndarray
    dt: float = 0.01
    max_step: int = 1000


def _f32(v):
    return np.asarray(v, dtype=np.float64).astype(np.float32)


def _box(name, high, act_high, control_step=5):
    high = _f32(high)
    d, a = len(high), len(act_high)
    return EnvSpec(name, ENV_IDS[name], d, a, d, 0, 0, control_step, -high, high, -_f32(act_high), _f32(act_high))


_pi = math.pi
SPECS = {
    "VanderPol": _box("VanderPol", [10.0, 10.0], [5.0]),
    "Pendulum": _box("Pendulum", [_pi, 10.0], [5.0]),
    "DuctedFan": _box("DuctedFan", [5.0, 5.0, _pi / 2, 5.0, 5.0, 5.0], [5.0, 5.0]),
    "TwoLink": _box("TwoLink", [_pi / 2, _pi / 2, 20.0, 20.0], [20.0, 20.0]),
    "SingleTrackCar": _box("SingleTrackCar", [1.0, 1.0, 1.066, 1.0, _pi / 2, _pi / 2, _pi / 3], [5.0, 5.0]),
    "QuadTracking": EnvSpec("QuadTracking", ENV_IDS["QuadTracking"], 12, 4, 30, 10, 18, 4,
                            -10.0 * np.ones(12, np.float32), 10.0 * np.ones(12, np.float32),
                            np.array([0.0 * (4.34 * 9.8), -10.0, -10.0, -10.0], dtype=np.float32),
                            np.array([2.0 * (4.34 * 9.8), 10.0, 10.0, 10.0], dtype=np.float32)),
}


def get_spec(env_name):
    try:
        return SPECS[env_name]
    except KeyError:
        # same failure mode as RL/env/make_env.py:33
        raise ValueError(f"Unknown custom env: {env_name}")
