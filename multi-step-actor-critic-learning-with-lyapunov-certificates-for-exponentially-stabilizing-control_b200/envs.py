"""GPU-resident vector environments -- drop-in for `RL.create_pkg.create_envs.create_envs`.

`create_envs(**args)` returns a `B200VectorEnv` that exposes the attribute / method surface
the reference's consumers use on a gymnasium SyncVectorEnv (RL/create_pkg/create_envs.py:9-34;
consumers: RL/trainer/sampler/base.py:57-61,98,141-148,160, RL/utils/init_args.py:33-43,
RL/trainer/evaluator.py:172): `single_observation_space`, `single_action_space`,
`observation_space`, `action_space`, `num_envs`, `reset(seed=None)`,
`step(actions) -> (obs, rewards, terminations, truncations, infos)` with
`infos["final_observation"]` on done envs.  All arithmetic runs in libmsacl_b200.so; NumPy
in -> NumPy out (host copies), torch CUDA in -> torch CUDA out (no copies).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .specs import get_spec


class Box:
    """Minimal stand-in for gymnasium.spaces.Box (low/high/shape/dtype only)."""

    def __init__(self, low, high, dtype=np.float32):
        self.low = np.asarray(low, dtype=dtype)
        self.high = np.asarray(high, dtype=dtype)
        self.shape = self.low.shape
        self.dtype = np.dtype(dtype)

    def sample(self):
        return np.random.uniform(self.low, self.high).astype(self.dtype)

    def __repr__(self):
        return f"Box({self.low}, {self.high}, {self.shape}, {self.dtype})"


class EnvStateBuffers:
    """Owns the SoA device state of n env instances and the C descriptor pointing at it."""

    def __init__(self, env_name, n, seed=0, env_base=0, device="cuda", max_step=None):
        self.spec = get_spec(env_name)
        self.n = int(n)
        self.device = torch.device(device)
        s = self.spec
        self.sf = torch.zeros(s.sf_rows, self.n, dtype=torch.float32, device=self.device)
        self.sd = torch.zeros(max(s.sd_rows, 1), self.n, dtype=torch.float64, device=self.device)
        self.step = torch.zeros(self.n, dtype=torch.int32, device=self.device)
        self.episode = torch.zeros(self.n, dtype=torch.int32, device=self.device)
        self.ep_return = torch.zeros(self.n, dtype=torch.float32, device=self.device)
        self.ep_len = torch.zeros(self.n, dtype=torch.int32, device=self.device)
        self.run = torch.zeros(self.n, dtype=torch.int32, device=self.device)
        self.desc = _lib.EnvState(
            env_id=s.env_id, max_step=int(s.max_step if max_step is None else max_step), n=self.n, stride=self.n,
            sf=self.sf.data_ptr(), sd=self.sd.data_ptr() if s.sd_rows else None, step=self.step.data_ptr(),
            episode=self.episode.data_ptr(), ep_return=self.ep_return.data_ptr(), ep_len=self.ep_len.data_ptr(),
            run=self.run.data_ptr(), seed=int(seed) & (2 ** 64 - 1), env_base=int(env_base))

    @property
    def obs(self):
        """[n, obs_dim] view-copy of the current observations (device)."""
        s = self.spec
        return self.sf[s.obs_off:s.obs_off + s.obs_dim].t().contiguous()

    def reset(self):
        _lib.check(_lib.load().msacl_env_reset(C.byref(self.desc), _lib.current_stream()))

    def set_box_state(self, obs, step=None):
        """Inject observations (box envs: state == obs)."""
        obs = torch.as_tensor(obs, dtype=torch.float32, device=self.device)
        self.sf.copy_(obs.t())
        if step is not None:
            self.step.copy_(torch.as_tensor(step, dtype=torch.int32, device=self.device))

    def set_quad_state(self, x, v, R, Om, t=None, Rd_last=None, obs=None, step=None):
        """Inject QuadTracking hidden state.  With t/Rd_last/obs omitted the desired-frame state is
        recomputed as reset() does (t = 0)."""
        dev = self.device
        f = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32, device=dev)
        self.sf[0:3] = f(x).t(); self.sf[3:6] = f(v).t()
        self.sf[6:15] = f(R).reshape(self.n, 9).t(); self.sf[15:18] = f(Om).t()
        if t is None:
            _lib.check(_lib.load().msacl_quad_init_from_raw(C.byref(self.desc), _lib.current_stream()))
        else:
            self.sd[0] = torch.as_tensor(np.asarray(t), dtype=torch.float64, device=dev)
            self.sd[1:10] = torch.as_tensor(np.asarray(Rd_last), dtype=torch.float64, device=dev).reshape(self.n, 9).t()
            self.sf[18:30] = f(obs).t()
        if step is not None:
            self.step.copy_(torch.as_tensor(np.asarray(step), dtype=torch.int32, device=dev))

    def get_quad_state(self):
        n = self.n
        return dict(x=self.sf[0:3].t().cpu().numpy(), v=self.sf[3:6].t().cpu().numpy(),
                    R=self.sf[6:15].t().reshape(n, 3, 3).cpu().numpy(), Om=self.sf[15:18].t().cpu().numpy(),
                    t=self.sd[0].cpu().numpy(), Rd_last=self.sd[1:10].t().reshape(n, 3, 3).cpu().numpy(),
                    obs=self.sf[18:30].t().cpu().numpy(), step=self.step.cpu().numpy())


class B200VectorEnv:
    def __init__(self, env_name, env_num, env_seed=0, env_base=0, device="cuda", max_step=None):
        self.env_name = env_name
        self.spec = get_spec(env_name)
        self.num_envs = int(env_num)
        s = self.spec
        self.single_observation_space = Box(s.obs_low, s.obs_high)
        self.single_action_space = Box(s.act_low, s.act_high)
        self.observation_space = Box(np.tile(s.obs_low, (self.num_envs, 1)), np.tile(s.obs_high, (self.num_envs, 1)))
        self.action_space = Box(np.tile(s.act_low, (self.num_envs, 1)), np.tile(s.act_high, (self.num_envs, 1)))
        self.is_vector_env = True
        self.state = EnvStateBuffers(env_name, env_num, seed=env_seed, env_base=env_base, device=device, max_step=max_step)
        dev = self.state.device
        n, d = self.num_envs, s.obs_dim
        self._next_obs = torch.empty(n, d, dtype=torch.float32, device=dev)
        self._final_obs = torch.empty(n, d, dtype=torch.float32, device=dev)
        self._reward = torch.empty(n, dtype=torch.float32, device=dev)
        self._term = torch.empty(n, dtype=torch.uint8, device=dev)
        self._trunc = torch.empty(n, dtype=torch.uint8, device=dev)
        self._resets = 0

    def reset(self, seed=None, options=None):
        st = self.state
        if seed is not None:
            st.desc.seed = int(seed) & (2 ** 64 - 1)
        if self._resets:          # every reset() call starts a fresh block of episodes so repeated resets differ
            st.episode.add_(1)
        self._resets += 1
        st.reset()
        return st.obs.cpu().numpy(), {}

    def step_device(self, actions):
        """actions: CUDA float32 [n, act_dim] (already clipped).  Returns device tensors
        (next_obs, reward, terminated u8, truncated u8, final_obs); valid until the next call."""
        st = self.state
        actions = actions.contiguous()
        _lib.check(_lib.load().msacl_env_step(C.byref(st.desc), actions.data_ptr(), self._next_obs.data_ptr(),
                                             self._reward.data_ptr(), self._term.data_ptr(), self._trunc.data_ptr(),
                                             self._final_obs.data_ptr(), _lib.current_stream()))
        return self._next_obs, self._reward, self._term, self._trunc, self._final_obs

    def step(self, actions):
        if isinstance(actions, torch.Tensor) and actions.is_cuda:
            nxt, rew, te, tr, fin = self.step_device(actions.to(torch.float32))
            done = (te | tr).bool()
            infos = {"final_observation": fin, "_final_observation": done} if bool(done.any()) else {}
            return nxt, rew, te.bool(), tr.bool(), infos
        a = torch.as_tensor(np.asarray(actions, dtype=np.float32)).to(self.state.device, non_blocking=True)
        nxt, rew, te, tr, fin = self.step_device(a)
        nxt_h = nxt.cpu().numpy()
        rew_h = rew.cpu().numpy().astype(np.float64)      # SyncVectorEnv returns a float64 reward array
        te_h = te.cpu().numpy().astype(bool)
        tr_h = tr.cpu().numpy().astype(bool)
        infos = {}
        done = te_h | tr_h
        if done.any():
            fin_h = fin.cpu().numpy()
            fo = np.full(self.num_envs, None, dtype=object)
            for i in np.nonzero(done)[0]:
                fo[i] = fin_h[i]
            infos = {"final_observation": fo, "_final_observation": done}
        return nxt_h, rew_h, te_h, tr_h, infos

    def close(self):
        pass


class B200Env:
    """Single-instance env with the gym.Env surface of the reference classes
    (`reset(seed, options) -> (obs, {})`, `step(a) -> (obs, reward, terminated, truncated, {})`;
    e.g. RL/env/VanderPol.py:69-130).  No autoreset, like the reference classes: the device state is stepped with
    max_step = -1 (bare-env mode of msacl_env_step), so a step() after terminated=True continues from the terminal
    state; the time limit is applied on the host from `current_step`, as the reference classes do."""

    def __init__(self, env_name, device="cuda"):
        self.spec = get_spec(env_name)
        s = self.spec
        self.observation_space = Box(s.obs_low, s.obs_high)
        self.action_space = Box(s.act_low, s.act_high)
        self.obs_dim, self.act_dim = s.obs_dim, s.act_dim
        self.dt, self.control_step, self.max_step = s.dt, s.control_step, s.max_step
        self._v = B200VectorEnv(env_name, 1, device=device, max_step=-1)
        self.current_step = 0

    def reset(self, seed=None, options=None):
        obs, _ = self._v.reset(seed=seed)
        self.current_step = 0
        return obs[0], {}

    def step(self, action):
        st = self._v.state
        nxt, rew, te, tr, fin = self._v.step_device(torch.as_tensor(np.asarray(action, np.float32)).reshape(1, -1).to(st.device))
        self.current_step += 1
        obs = fin.cpu().numpy()[0]
        return obs, np.float32(rew.item()), bool(te.item()), self.current_step >= self.max_step, {}


def create_envs(**args):
    """Drop-in for RL/create_pkg/create_envs.py:9-34 (keys env_name, env_seed, env_num)."""
    return B200VectorEnv(args.get("env_name"), args.get("env_num"), env_seed=args.get("env_seed") or 0,
                         env_base=args.get("env_base", 0), device=args.get("device", "cuda"))


def make_env(env_id, seed=0, idx=0, capture_video=False, run_name=""):
    """Drop-in for RL/env/make_env.py:10-41: returns a thunk building one env instance."""
    get_spec(env_id)   # raises ValueError for unknown ids, as the reference does

    def thunk():
        return B200Env(env_id)
    return thunk
