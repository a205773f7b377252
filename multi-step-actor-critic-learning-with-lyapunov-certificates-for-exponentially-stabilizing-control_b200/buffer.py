"""Device-resident n-step replay ring -- drop-in for `NstepReplayBuffer`
(RL/trainer/buffer/nstep_replay_buffer.py:40-150): same constructor kwargs, `store`,
`add_batch`, `sample_batch`, `size`, `ptr`, `len()`, `__get_RAM__()`.  Arrays keep the
reference layout [max_size, n_step, .]; ptr/size arithmetic is done by the device scan in
msacl_window_store (bit-exact slot assignment in the reference's append order).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .sampler import DeviceWindowBatch

FIELDS = ("obs", "act", "rew", "cost", "obs2", "done", "logp")


class B200NstepReplayBuffer:
    def __init__(self, **kwargs):
        self.obsv_dim = int(kwargs["obs_dim"])
        self.act_dim = int(kwargs["act_dim"])
        self.max_size = int(kwargs["buffer_max_size"])
        self.n_step = int(kwargs["n_step"])
        self.device = torch.device(kwargs.get("device", "cuda"))
        z = lambda *s: torch.zeros(self.max_size, self.n_step, *s, dtype=torch.float32, device=self.device)
        self.n_step_buf = {"obs": z(self.obsv_dim), "act": z(self.act_dim), "rew": z(), "cost": z(), "obs2": z(self.obsv_dim),
                           "done": z(), "logp": z()}
        self._ptr_size = torch.zeros(2, dtype=torch.int64, device=self.device)
        self._count = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._scratch = None
        self._ring = self._make_ring(self.n_step_buf, self.max_size)
        self._gen = torch.Generator(device=self.device)
        self._gen.manual_seed(int(kwargs.get("seed") or 0))

    def _make_ring(self, bufs, max_size):
        return _lib.Ring(max_size=max_size, n_step=self.n_step, obs_dim=self.obsv_dim, act_dim=self.act_dim,
                         **{k: bufs[k].data_ptr() for k in FIELDS})

    # ---- bookkeeping (device counters; reading them synchronises)
    @property
    def ptr(self):
        return int(self._ptr_size[0].item())

    @property
    def size(self):
        return int(self._ptr_size[1].item())

    def __len__(self):
        return self.size

    def __get_RAM__(self):
        size = self.size
        if size == 0:
            return 0.0
        per = sum(v[0].numel() * 4 for v in self.n_step_buf.values())
        return round(per * size / (1024 * 1024), 2)

    # ---- writes
    def store(self, obs, act, rew, cost, next_obs, done, logp):
        """Host path, one window (nstep_replay_buffer.py:91-119)."""
        p = self.ptr
        vals = dict(obs=obs, act=act, rew=rew, cost=cost, obs2=next_obs, done=done, logp=logp)
        for k, v in vals.items():
            self.n_step_buf[k][p] = torch.as_tensor(np.asarray(v, dtype=np.float32), device=self.device)
        self._ptr_size[0] = (p + 1) % self.max_size
        self._ptr_size[1] = min(self.size + 1, self.max_size)

    def add_batch(self, samples):
        if isinstance(samples, DeviceWindowBatch):
            return self.add_device_batch(samples)
        for s in samples:
            self.store(*s)

    def add_device_batch(self, batch: DeviceWindowBatch):
        tr = batch.tr
        nb = (tr.K * tr.n + 255) // 256
        if self._scratch is None or self._scratch.numel() < nb + 2:
            self._scratch = torch.zeros(nb + 2, dtype=torch.int64, device=self.device)
        desc = tr.desc(0)
        _lib.check(_lib.load().msacl_window_store(C.byref(desc), tr.H, tr.K, tr.n, C.byref(self._ring), self._ptr_size.data_ptr(),
                                                 self._count.data_ptr(), self._scratch.data_ptr(), _lib.current_stream()))
        return self._count

    # ---- reads
    def gather(self, idx):
        idx = torch.as_tensor(idx, dtype=torch.int64, device=self.device).contiguous()
        B = idx.numel()
        out = {k: torch.empty(B, *v.shape[1:], dtype=torch.float32, device=self.device) for k, v in self.n_step_buf.items()}
        dst = self._make_ring(out, B)
        _lib.check(_lib.load().msacl_ring_gather(C.byref(self._ring), idx.data_ptr(), B, C.byref(dst), _lib.current_stream()))
        return out

    def sample_batch(self, batch_size: int) -> dict:
        """Uniform sampling with replacement over the valid range (nstep_replay_buffer.py:138)."""
        idx = torch.randint(0, self.size, (int(batch_size),), device=self.device, generator=self._gen)
        return self.gather(idx)
