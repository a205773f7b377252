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
        self._seed = int(kwargs.get("seed") or 0) & (2 ** 64 - 1)
        self._draws = 0                  # sample_batch calls so far: keys the Philox index draws together with the seed
        self._idx = None

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
    def gather(self, idx, out=None):
        """out: optional preallocated {field: [B, n_step, .]} float32 CUDA tensors (e.g. views of a packed exchange buffer)."""
        idx = torch.as_tensor(idx, dtype=torch.int64, device=self.device).contiguous()
        B = idx.numel()
        if out is None:
            out = {k: torch.empty(B, *v.shape[1:], dtype=torch.float32, device=self.device) for k, v in self.n_step_buf.items()}
        dst = self._make_ring(out, B)
        _lib.check(_lib.load().msacl_ring_gather(C.byref(self._ring), idx.data_ptr(), B, C.byref(dst), _lib.current_stream()))
        return out

    def sample_batch(self, batch_size: int, out=None) -> dict:
        """Uniform sampling with replacement over the valid range (nstep_replay_buffer.py:138)."""
        # idx ~ U{0..size-1} drawn by the library from the device-resident size (no host synchronisation, two launches)
        B = int(batch_size)
        if out is None:
            out = {k: torch.empty(B, *v.shape[1:], dtype=torch.float32, device=self.device) for k, v in self.n_step_buf.items()}
        dst = self._make_ring(out, B)
        if self._idx is None or self._idx.numel() < B:
            self._idx = torch.empty(B, dtype=torch.int64, device=self.device)
        self._draws += 1
        _lib.check(_lib.load().msacl_ring_sample(C.byref(self._ring), self._ptr_size.data_ptr(), self._seed, self._draws, B, C.byref(dst),
                                                self._idx.data_ptr(), _lib.current_stream()))
        return out


class B200IndexedReplayBuffer:
    """Index-based n-step replay with the `NstepReplayBuffer` surface (`add_batch`, `sample_batch`,
    `size`, `ptr`, `len()`, `__get_RAM__()`; RL/trainer/buffer/nstep_replay_buffer.py:40-150).

    The payload is the sampler's own `[H + M*K, n, .]` transition store (what the rollout kernel writes anyway,
    4*(2D+A+3)+2 bytes per transition); a window is the flat position of its newest transition, kept in an int64 ring
    of `buffer_max_size` entries with the reference's append order and ptr / size arithmetic.  Appending a chunk costs
    8 bytes per emitted window instead of copying n_step rows twice, and a store of S transitions serves ~S windows
    instead of S / n_step -- a replay that scales with a 2e9 env-steps/s sampler.

    Retention: a window lives as long as its n_step slices are in the store, i.e. for M - 2 further launches
    (M = chunks of the sampler's store).  `sample_batch` draws uniformly from the most recent
    min(size, windows emitted by the last M - 2 launches) entries; with
    M >= chunks_for(buffer_max_size, env_num, horizon) that is the whole ring whenever every transition emits, and
    it is exactly the reference's `randint(0, size)` as long as buffer_max_size >= windows produced."""

    def __init__(self, **kwargs):
        self.obsv_dim = int(kwargs["obs_dim"])
        self.act_dim = int(kwargs["act_dim"])
        self.max_size = int(kwargs["buffer_max_size"])
        self.n_step = int(kwargs["n_step"])
        self.device = torch.device(kwargs.get("device", "cuda"))
        self.win_pos = torch.zeros(self.max_size, dtype=torch.int64, device=self.device)
        self._ptr_size = torch.zeros(2, dtype=torch.int64, device=self.device)
        self._count = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._scratch = None
        self._host_ps = (0, 0)
        self._tr = None                  # the sampler's TransitionBuffers, bound on the first add_batch
        self._launch_counts = None       # windows emitted by each of the last M - 2 launches (device, circular)
        self._adds = 0
        self._seed = int(kwargs.get("seed") or 0) & (2 ** 64 - 1)
        self._draws = 0                  # sample_batch calls so far: keys the Philox index draws together with the seed

    @staticmethod
    def chunks_for(buffer_max_size, num_envs, horizon):
        """Chunks of K = horizon slices the sampler's store needs so that buffer_max_size windows stay resident."""
        per_launch = max(1, int(num_envs) * int(horizon))
        return max(3, -(-int(buffer_max_size) // per_launch) + 2)

    def _host_counters(self):
        if self._host_ps is None:
            self._host_ps = tuple(int(x) for x in self._ptr_size.tolist())
        return self._host_ps

    ptr = property(lambda self: self._host_counters()[0])
    size = property(lambda self: self._host_counters()[1])

    def __len__(self):
        return self.size

    def __get_RAM__(self):
        """MB the stored windows occupy: their share of the transition store + the position ring."""
        if self._tr is None or self.size == 0:
            return 0.0
        per = 4 * (2 * self.obsv_dim + self.act_dim + 3) + 2 + 8
        return round(per * self.size / (1024 * 1024), 2)

    def store(self, *a, **k):
        raise NotImplementedError("the index-based buffer stores windows by reference to the sampler's device transition store; "
                                  "use B200NstepReplayBuffer for host-side store()")

    def add_batch(self, batch):
        if not isinstance(batch, DeviceWindowBatch):
            raise NotImplementedError("B200IndexedReplayBuffer.add_batch takes the DeviceWindowBatch returned by sampler.sample()")
        batch.check_current()
        tr = batch.tr
        if self._tr is None:
            if tr.M < 3:
                raise ValueError("index-based replay needs a transition store of >= 3 chunks (sampler kwarg history_chunks, "
                                 "or buffer_name='b200_indexed_replay_buffer' in the sampler's kwargs)")
            if tr.H < self.n_step - 1:
                raise ValueError("transition store history is shorter than n_step - 1")
            self._tr = tr
            self._launch_counts = torch.zeros(tr.M - 2, dtype=torch.int64, device=self.device)
        elif tr is not self._tr:
            raise ValueError("B200IndexedReplayBuffer is bound to one sampler's transition store")
        lib = _lib.load()
        need = int(lib.msacl_window_store_scratch_elems(tr.K, tr.n))
        if self._scratch is None or self._scratch.numel() < need:
            self._scratch = torch.zeros(need, dtype=torch.int64, device=self.device)
        base = tr.chunk_base_slice * tr.n
        emit_new = tr._full["emit"][tr.chunk_base_slice:]
        self._host_ps = None
        _lib.check(lib.msacl_window_index_store(emit_new.data_ptr(), tr.K, tr.n, base, self.win_pos.data_ptr(), self.max_size,
                                                self._ptr_size.data_ptr(), self._count.data_ptr(), self._scratch.data_ptr(),
                                                _lib.current_stream()))
        self._launch_counts[self._adds % (tr.M - 2)] = self._count[0]
        self._adds += 1
        return self._count

    add_device_batch = add_batch

    def valid_count(self):
        """Device scalar: number of most-recent ring entries whose slices are still resident."""
        return torch.minimum(self._ptr_size[1], self._launch_counts.sum())

    def gather(self, idx, out=None):
        """idx: ring slots (as the reference's fancy indexing of n_step_buf)."""
        idx = torch.as_tensor(idx, dtype=torch.int64, device=self.device).contiguous()
        B = idx.numel()
        n, D, A = self.n_step, self.obsv_dim, self.act_dim
        z = lambda *s: torch.empty(B, n, *s, dtype=torch.float32, device=self.device)
        if out is None:
            out = {"obs": z(D), "act": z(A), "rew": z(), "cost": z(), "obs2": z(D), "done": z(), "logp": z()}
        dst = _lib.Ring(max_size=B, n_step=n, obs_dim=D, act_dim=A, **{k: out[k].data_ptr() for k in FIELDS})
        desc = self._tr.full_desc()
        _lib.check(_lib.load().msacl_window_gather_indexed(C.byref(desc), self._tr.n, self.win_pos.data_ptr(), idx.data_ptr(), B,
                                                          C.byref(dst), _lib.current_stream()))
        return out

    def sample_batch(self, batch_size: int, out=None, return_slots=False) -> dict:
        """Uniform with replacement over the resident windows (nstep_replay_buffer.py:138-146): index draw (Philox keyed by
        the buffer seed and a call counter) and gather in ONE launch that reads ptr / size / residency from the device."""
        if self._tr is None:
            raise RuntimeError("sample_batch before any add_batch")
        B = int(batch_size)
        n, D, A = self.n_step, self.obsv_dim, self.act_dim
        z = lambda *s: torch.empty(B, n, *s, dtype=torch.float32, device=self.device)
        if out is None:
            out = {"obs": z(D), "act": z(A), "rew": z(), "cost": z(), "obs2": z(D), "done": z(), "logp": z()}
        dst = _lib.Ring(max_size=B, n_step=n, obs_dim=D, act_dim=A, **{k: out[k].data_ptr() for k in FIELDS})
        slots = torch.empty(B, dtype=torch.int64, device=self.device) if return_slots else None
        desc = self._tr.full_desc()
        self._draws += 1
        _lib.check(_lib.load().msacl_window_sample_indexed(C.byref(desc), self._tr.n, self.win_pos.data_ptr(), self.max_size,
                                                          self._ptr_size.data_ptr(), self._launch_counts.data_ptr(),
                                                          self._launch_counts.numel(), self._seed, self._draws, B, C.byref(dst),
                                                          None if slots is None else slots.data_ptr(), _lib.current_stream()))
        return (out, slots) if return_slots else out
