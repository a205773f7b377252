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
        self._host_ps = (0, 0)
        self._ring = self._make_ring(self.n_step_buf, self.max_size)
        self._gen = torch.Generator(device=self.device)
        self._gen.manual_seed(int(kwargs.get("seed") or 0))

    def _make_ring(self, bufs, max_size):
        return _lib.Ring(max_size=max_size, n_step=self.n_step, obs_dim=self.obsv_dim, act_dim=self.act_dim,
                         **{k: bufs[k].data_ptr() for k in FIELDS})

    # ---- bookkeeping: the counters live on the device (the window scan updates them); the host copy is refreshed
    #      lazily, i.e. reading .ptr / .size synchronises only after a device append, and sample_batch never reads them
    def _host_counters(self):
        if self._host_ps is None:
            self._host_ps = tuple(int(x) for x in self._ptr_size.tolist())
        return self._host_ps

    @property
    def ptr(self):
        return self._host_counters()[0]

    @property
    def size(self):
        return self._host_counters()[1]

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
        new = ((p + 1) % self.max_size, min(self.size + 1, self.max_size))
        self._ptr_size.copy_(torch.tensor(new, dtype=torch.int64))
        self._host_ps = new

    def add_batch(self, samples):
        if isinstance(samples, DeviceWindowBatch):
            return self.add_device_batch(samples)
        for s in samples:
            self.store(*s)

    def add_device_batch(self, batch: DeviceWindowBatch):
        batch.check_current()
        tr = batch.tr
        self._host_ps = None
        need = int(_lib.load().msacl_window_store_scratch_elems(tr.K, tr.n))      # exactly the size the header documents
        if self._scratch is None or self._scratch.numel() < need:
            self._scratch = torch.zeros(need, dtype=torch.int64, device=self.device)
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
        # idx ~ U{0..size-1} drawn on the device from the device-resident size (no host synchronisation)
        u = torch.rand(int(batch_size), device=self.device, generator=self._gen, dtype=torch.float64)
        size = self._ptr_size[1]
        idx = torch.minimum((u * size).to(torch.int64), torch.clamp(size - 1, min=0))
        return self.gather(idx)
