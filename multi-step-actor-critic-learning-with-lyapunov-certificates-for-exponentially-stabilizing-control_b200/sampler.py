"""Fused GPU rollout sampler -- drop-in for `NstepOffSampler`
(RL/trainer/sampler/nstep_off_sampler.py:8-29 on top of RL/trainer/sampler/base.py:48-323).

`sample()` runs `horizon = sample_batch_size` vector steps in ONE kernel launch
(msacl_rollout_fused): actor forward, TanhGauss sampling, clipping, env dynamics, reward/cost
scaling, same-step autoreset and the n-step deque bookkeeping all happen on the device with
the env state resident in registers.  The emitted n-step windows stay on the device
(`DeviceWindowBatch`) and are appended to `B200NstepReplayBuffer` by a device scatter.
"""
import ctypes as C
import time
from typing import NamedTuple

import numpy as np
import torch

from . import _lib
from .envs import EnvStateBuffers, B200VectorEnv
from .specs import get_spec

SAMPLER_TIME_TAG = "Time/Sampler time [ms]-RL iter"      # RL/utils/tensorboard_setup.py:38


class nStepExperience(NamedTuple):
    """Same fields as RL/trainer/sampler/base.py:33-45."""
    n_step_obs: np.ndarray
    n_step_act: np.ndarray
    n_step_rew: np.ndarray
    n_step_cost: np.ndarray
    n_step_obs2: np.ndarray
    n_step_done: np.ndarray
    n_step_log_prob: np.ndarray


class ActorWeights:
    """Device copy of a StochaPolicy MLP (obs -> h1 -> h2 -> 2*act, ReLU, h1, h2 <= 256: zero-padded to the kernels' 256-wide
    layers), packed for the fused rollout kernels."""

    HIDDEN = 256

    def __init__(self, layers, device="cuda", min_log_std=-20.0, max_log_std=1.0):
        (w1, b1), (w2, b2), (w3, b3) = layers
        dev = torch.device(device)
        t = lambda a: torch.as_tensor(np.asarray(a) if not isinstance(a, torch.Tensor) else a.detach(), dtype=torch.float32).to(dev).contiguous()
        w1, b1, w2, b2, w3, b3 = t(w1), t(b1), t(w2), t(b2), t(w3), t(b3)
        H = self.HIDDEN
        h1, h2 = w1.shape[0], w2.shape[0]
        if w2.shape[1] != h1 or w3.shape[1] != h2:
            raise ValueError("layer widths do not chain")
        if h1 > H or h2 > H:
            raise ValueError("the fused rollout kernels hold two hidden layers of at most 256 units (policy_hidden_sizes=[256, 256] "
                             "is the reference default); wider policies run on the general engine")
        if (h1, h2) != (H, H):
            # narrower ReLU layers are embedded in the 256-wide kernel by zero padding: the extra units have zero weights and
            # biases, relu(0) = 0, and adding exact zeros changes no partial sum -- the logits are those of the narrow network
            z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
            W1, B1, W2, B2, W3 = z(H, w1.shape[1]), z(H), z(H, H), z(H), z(w3.shape[0], H)
            W1[:h1] = w1; B1[:h1] = b1; W2[:h2, :h1] = w2; B2[:h2] = b2; W3[:, :h2] = w3
            w1, b1, w2, b2, w3 = W1, B1, W2, B2, W3
        self.hidden_sizes = (h1, h2)
        self.w1, self.b1, self.b2, self.w3, self.b3 = w1, b1, b2, w3, b3
        self.w2t = w2.t().contiguous()
        self.obs_dim = self.w1.shape[1]
        self.act_dim = self.w3.shape[0] // 2
        self.desc = _lib.Actor(w1=self.w1.data_ptr(), b1=self.b1.data_ptr(), w2t=self.w2t.data_ptr(), b2=self.b2.data_ptr(),
                               w3=self.w3.data_ptr(), b3=self.b3.data_ptr(), min_log_std=float(min_log_std),
                               max_log_std=float(max_log_std))

    def tc_images(self):
        """UMMA operand images (split-bf16 W1|b1 and W2 chunk images) for the tensor-core kernel."""
        if getattr(self, "_tc", None) is None:
            lib = _lib.load()
            n1, n2 = C.c_int64(0), C.c_int64(0)
            _lib.check(lib.msacl_tc_pack_bytes(C.byref(n1), C.byref(n2)))
            w1p = torch.empty(n1.value, dtype=torch.uint8, device=self.w1.device)
            w2p = torch.empty(n2.value, dtype=torch.uint8, device=self.w1.device)
            _lib.check(lib.msacl_tc_pack_actor(C.byref(self.desc), self.obs_dim, w1p.data_ptr(), w2p.data_ptr(), _lib.current_stream()))
            self._tc = (w1p, w2p)
        return self._tc

    @classmethod
    def from_policy(cls, policy, device="cuda"):
        """policy: a reference-style StochaPolicy (RL/apprfunc/mlp.py:111-136): `.policy` is an
        nn.Sequential of Linear/activation pairs, `.min_log_std`, `.max_log_std`."""
        seq = policy.policy if hasattr(policy, "policy") else policy
        lin = [m for m in seq if isinstance(m, torch.nn.Linear)]
        if len(lin) != 3:
            raise ValueError("expected a 2-hidden-layer policy MLP")
        # the rollout kernels hard-wire ReLU hidden activations and a linear output layer (the reference defaults,
        # example/msacl_train.py: policy_hidden_activation "relu"); any other module would make the sampler run a
        # different network than the learner, silently corrupting the stored act / logp
        other = [m for m in seq if not isinstance(m, torch.nn.Linear)]
        if len(other) != 3 or not all(isinstance(m, torch.nn.ReLU) for m in other[:2]) or not isinstance(other[2], torch.nn.Identity):
            raise ValueError("the fused rollout kernel is specialised for policy_hidden_activation='relu' with a linear "
                             f"output layer; got {[type(m).__name__ for m in other]}")
        return cls([(l.weight, l.bias) for l in lin], device=device,
                   min_log_std=getattr(policy, "min_log_std", -20.0), max_log_std=getattr(policy, "max_log_std", 1.0))

    @staticmethod
    def policy_version(policy):
        """Changes whenever a parameter of `policy` is updated in place (optimizer step, load_state_dict) or replaced."""
        return tuple((p.data_ptr(), p._version) for p in policy.parameters())


class GeneralActor:
    """Any StochaPolicy MLP (RL/apprfunc/mlp.py:18-33,111-136): Linear / activation pairs of arbitrary depth, width and
    activation -- the policies `ActorWeights` (the fused rollout kernels: two ReLU hidden layers of at most 256 units) cannot take.
    A forward pass is one `msacl_gemm_tc` per layer (split-bf16 tcgen05 GEMM at FP32-class precision; bias and ReLU / Tanh
    fused into the epilogue, any other activation module applied in place to the GEMM output); layer 1 reads the
    observations straight out of the structure-of-arrays env state (row stride 1, k stride = env pitch)."""

    _FUSED_ACT = {torch.nn.ReLU: 1, torch.nn.Tanh: 2, torch.nn.Identity: 0}

    def __init__(self, layers, activations, device="cuda", min_log_std=-20.0, max_log_std=1.0, precision=6):
        dev = torch.device(device)
        t = lambda a: torch.as_tensor(np.asarray(a) if not isinstance(a, torch.Tensor) else a.detach(), dtype=torch.float32).to(dev).contiguous()
        self.layers = [(t(w), t(b)) for w, b in layers]
        self.activations = list(activations)
        if len(self.activations) != len(self.layers):
            raise ValueError("one activation module per Linear layer (Identity for the output layer)")
        for (w0, _), (w1, _) in zip(self.layers[:-1], self.layers[1:]):
            if w1.shape[1] != w0.shape[0]:
                raise ValueError("layer widths do not chain")
        self.obs_dim = self.layers[0][0].shape[1]
        self.act_dim = self.layers[-1][0].shape[0] // 2
        self.min_log_std, self.max_log_std = float(min_log_std), float(max_log_std)
        self.precision = int(precision)
        self.device = dev
        self._ws = {}
        self._packed = None

    @classmethod
    def from_policy(cls, policy, device="cuda"):
        seq = policy.policy if hasattr(policy, "policy") else policy
        mods = list(seq)
        lin = [m for m in mods if isinstance(m, torch.nn.Linear)]
        acts = [m for m in mods if not isinstance(m, torch.nn.Linear)]
        if not lin or len(acts) != len(lin):
            raise ValueError("expected Linear / activation pairs (RL/apprfunc/mlp.py:18-33)")
        return cls([(l.weight, l.bias) for l in lin], acts, device=device,
                   min_log_std=getattr(policy, "min_log_std", -20.0), max_log_std=getattr(policy, "max_log_std", 1.0))

    def _workspace(self, n):
        if n not in self._ws:
            self._ws[n] = [torch.empty(n, w.shape[0], dtype=torch.float32, device=self.device) for w, _ in self.layers]
        return self._ws[n]

    def _packed_weights(self):
        """Pre-converted bf16 operand images of the layers' weights (msacl_gemm_pack_b), built once: an actor is a snapshot
        of the policy for one rollout (the sampler rebuilds it when a parameter changes), so every env step streams the same
        images instead of re-converting the weights in every CTA.  Layers wider than one 256-column tile stay unpacked."""
        if self._packed is None:
            from .learner import _desc
            lib = _lib.load()
            self._packed = []
            for w, _ in self.layers:
                if w.shape[0] > 256:
                    self._packed.append(None)
                    continue
                g = _desc(0, w.shape[1], 1, w, w.shape[1], 1, 128, w.shape[0], w.shape[1], 0, w.shape[0], precision=self.precision)
                buf = torch.empty(int(lib.msacl_gemm_packed_b_bytes(w.shape[1], self.precision)), dtype=torch.uint8, device=self.device)
                _lib.check(lib.msacl_gemm_pack_b(C.byref(g), buf.data_ptr(), _lib.current_stream()))
                self._packed.append(buf)
        return self._packed

    def forward_state(self, state):
        """logits [n, 2*act_dim] for the current observations of an `EnvStateBuffers`."""
        from .learner import _desc
        lib = _lib.load()
        n, spec = state.n, state.spec
        outs = self._workspace(n)
        packed = self._packed_weights()
        a_ptr, a_rs, a_ks = state.sf.data_ptr() + 4 * spec.obs_off * n, 1, n       # SoA: obs[r][k] = sf[obs_off + k][r]
        for (w, b), act, y, pk in zip(self.layers, self.activations, outs, packed):
            code = self._FUSED_ACT.get(type(act))
            g = _desc(a_ptr, a_rs, a_ks, w, w.shape[1], 1, n, w.shape[0], w.shape[1], y, w.shape[0], bias=b,
                      act=code or 0, precision=self.precision, b_packed=pk)
            _lib.check(lib.msacl_gemm_tc(C.byref(g), _lib.current_stream()))
            if code is None:                    # gelu / elu / selu / sigmoid ...: the module itself, on the device
                with torch.no_grad():
                    y.copy_(act(y))
            a_ptr, a_rs, a_ks = y.data_ptr(), w.shape[0], 1
        return outs[-1]


def actor_from_policy(policy, device="cuda", strict=False):
    """The packed actor the rollout engines take: `ActorWeights` for the policies the fused kernels hold (two ReLU hidden
    layers of at most 256 units; reference default [256, 256]), else (unless `strict`) a `GeneralActor` driving per-layer
    GEMMs + msacl_rollout_step."""
    try:
        return ActorWeights.from_policy(policy, device=device)
    except ValueError:
        if strict:
            raise
        return GeneralActor.from_policy(policy, device=device)


class TransitionBuffers:
    """Transition store of one rollout: for every field a `[H + M*K, n, .]` array, H = n_step - 1 history slices.

    Launch j writes the K slices of chunk j % M; the H slices in front of a chunk are the tail of the previous
    chunk, so n-step windows that straddle two launches are contiguous in memory without copying.  Only when the
    chunk index wraps (every M launches) the last H slices are carried back to the front.  `obs`, `act`, ... and
    `fields()` are views of the *current* window `[H + K, n, .]` (history first, the K new slices at `[H:]`), so
    consumers index them exactly as if the store were `[(H + K), n, .]` rolled after every launch (M = 1)."""

    NAMES = ("obs", "act", "rew", "cost", "obs2", "done", "logp", "emit")

    def __init__(self, spec, n, K, n_step, device, chunks=None, record_logits=False):
        self.H, self.K, self.n = n_step - 1, K, n
        slice_bytes = n * (4 * (2 * spec.obs_dim + spec.act_dim + 3) + 2)
        if chunks is None:
            chunks = 4
            if torch.device(device).type == "cuda":      # keep the store below a fifth of the free device memory
                free, _ = torch.cuda.mem_get_info(device)
                while chunks > 1 and (self.H + chunks * K) * slice_bytes > 0.2 * free:
                    chunks -= 1
        self.M = max(1, int(chunks))
        T = self.H + self.M * K
        f = lambda *s: torch.zeros(T, n, *s, dtype=torch.float32, device=device)
        b = lambda: torch.zeros(T, n, dtype=torch.uint8, device=device)
        self._full = dict(obs=f(spec.obs_dim), act=f(spec.act_dim), rew=f(), cost=f(), obs2=f(spec.obs_dim), done=b(),
                          logp=f(), emit=b())
        # optional diagnostic: the policy outputs (mean || log_std) every action was sampled from, [K, n, 2A] of the last launch
        self.logits = torch.zeros(K, n, 2 * spec.act_dim, dtype=torch.float32, device=device) if record_logits else None
        self._j = self.M - 1          # the first roll_history() wraps to chunk 0
        self.launches = 0             # number of roll_history() calls = rollout launches into this store
        # the carry-over copy runs on a side stream (copy engines) underneath the next rollout launch whenever that launch
        # cannot touch its source rows: see roll_history()
        self._side = torch.cuda.Stream(device) if torch.device(device).type == "cuda" else None
        self.pending_history = None

    def _view(self, name, join=True):
        if join:
            self.join_history()       # every reader of the window (history rows included) is ordered after a pending carry-over
        base = self._j * self.K
        return self._full[name][base:base + self.H + self.K]

    obs = property(lambda self: self._view("obs"))
    act = property(lambda self: self._view("act"))
    rew = property(lambda self: self._view("rew"))
    cost = property(lambda self: self._view("cost"))
    obs2 = property(lambda self: self._view("obs2"))
    done = property(lambda self: self._view("done"))
    logp = property(lambda self: self._view("logp"))
    emit = property(lambda self: self._view("emit"))

    def fields(self):
        return {k: self._view(k) for k in self.NAMES}

    def desc(self, t0=0, join=True):
        d = _lib.Transitions(**{k: self._view(k, join)[t0:].data_ptr() for k in self.NAMES})
        if self.logits is not None:
            d.logits = self.logits.data_ptr()
        return d

    def full_desc(self):
        """Base pointers of the whole [H + M*K, n, .] store (index-based replay: positions are absolute)."""
        self.join_history()
        return _lib.Transitions(**{k: v.data_ptr() for k, v in self._full.items()})

    @property
    def chunk_base_slice(self):
        """Slice index, in the whole store, of the first NEW slice of the current chunk."""
        return self._j * self.K + self.H

    def roll_history(self):
        """Advance to the next chunk (called once before every launch)."""
        self.launches += 1
        self._j += 1
        if self._j < self.M:
            return
        self._j = 0
        if self.H == 0:
            return
        src0 = self.M * self.K                          # tail of the last chunk -> history of chunk 0
        # The launch that follows writes rows [H, H + K) only; readers of rows [0, H) (window store / gather / materialize)
        # come after it in stream order.  If the source rows [M K, M K + H) are disjoint from the rows that launch writes,
        # the copy (contiguous D2D memcpys: copy engines, 1.6 ms for 2^21 quadrotor envs) overlaps the 15 ms rollout kernel
        # instead of delaying it; FusedRollout.run() makes the main stream wait for it right AFTER the launch.
        if self._side is not None and self.H + self.K <= src0:
            cur = torch.cuda.current_stream()
            ready = torch.cuda.Event()
            ready.record(cur)                           # the previous launch and all consumers of the old rows are enqueued before this
            self._side.wait_event(ready)
            with torch.cuda.stream(self._side):
                for v in self._full.values():
                    v[:self.H].copy_(v[src0:src0 + self.H])
                done = torch.cuda.Event()
                done.record(self._side)
            self.pending_history = done
            return
        for v in self._full.values():
            src = v[src0:src0 + self.H]
            v[:self.H].copy_(src.clone() if src0 < self.H else src)

    def join_history(self):
        """Order the current stream after an asynchronous carry-over copy (called right after the rollout launch)."""
        if self.pending_history is not None:
            torch.cuda.current_stream().wait_event(self.pending_history)
            self.pending_history = None


class DeviceWindowBatch:
    """Result of one `sample()`: the chunk's transitions + emit flags, still on the device."""

    def __init__(self, tr: TransitionBuffers, n_step):
        self.tr, self.n_step = tr, n_step
        self._launch = tr.launches          # the views of `tr` follow the current chunk: a batch is only valid until the next launch

    def check_current(self):
        if self.tr.launches != self._launch:
            raise RuntimeError("stale DeviceWindowBatch: the sampler has launched another rollout since this batch was "
                               "returned; add it to the buffer (or materialize() it) before the next sample()")

    def count(self):
        self.check_current()
        return int(self.tr.emit[self.tr.H:].sum().item())

    def __len__(self):
        return self.count()

    def materialize(self):
        """Host list of nStepExperience in the reference's order (step-major, env-minor)."""
        self.check_current()
        tr, ns = self.tr, self.n_step
        emit = self.tr.emit[tr.H:].cpu().numpy().astype(bool)
        f = {k: v.cpu().numpy() for k, v in tr.fields().items()}
        out = []
        for t, i in zip(*np.nonzero(emit)):
            sl = slice(t + tr.H - ns + 1, t + tr.H + 1)
            out.append(nStepExperience(f["obs"][sl, i], f["act"][sl, i], f["rew"][sl, i], f["cost"][sl, i], f["obs2"][sl, i],
                                       f["done"][sl, i].astype(np.float32), f["logp"][sl, i]))
        return out

    def __iter__(self):
        return iter(self.materialize())


class FusedRollout:
    """Owns env state + transition buffers and launches msacl_rollout_fused."""

    def __init__(self, env_name, num_envs, horizon, n_step=20, reward_scale=100.0, cost_scale=100.0, seed=0, env_base=0,
                 device="cuda", max_step=None, state=None, engine="ffma", history_chunks=None, record_logits=False):
        self.engine = engine
        self.spec = get_spec(env_name)
        self.state = state or EnvStateBuffers(env_name, num_envs, seed=seed, env_base=env_base, device=device, max_step=max_step)
        self.n, self.K, self.n_step = self.state.n, int(horizon), int(n_step)
        self.reward_scale, self.cost_scale = float(reward_scale), float(cost_scale)
        self.tr = TransitionBuffers(self.spec, self.n, self.K, self.n_step, self.state.device, chunks=history_chunks,
                                    record_logits=record_logits)
        self.stats = torch.zeros(32, dtype=torch.float64, device=self.state.device)   # [0:8) documented, rest diagnostic
        self.global_step = 0

    @staticmethod
    def reserve_sms(count):
        """Leave `count` SMs free for kernels that run concurrently with the (persistent, one CTA per SM) tensor-core rollout
        kernel, e.g. a side-stream NCCL collective.  0 restores the full grid."""
        _lib.check(_lib.load().msacl_rollout_tc_set_max_ctas(0 if count <= 0 else 148 - int(count)))

    def run(self, actor: ActorWeights, eps=None, deterministic=False, write=True, engine=None, timing=None):
        """One K-step chunk.  eps: optional CUDA float32 [K, n, act_dim] explicit N(0,1) draws.
        engine: "ffma" (FP32 FFMA actor) or "tc" (tcgen05 split-bf16 actor); default self.engine.
        timing: optional (start, end) torch.cuda.Event pair recorded on the current stream directly around the kernel
        launch (excludes the history carry-over copy of the transition store)."""
        if actor.obs_dim != self.spec.obs_dim or actor.act_dim != self.spec.act_dim:
            raise ValueError("actor dimensions do not match the environment")
        tr = self.tr
        if write:
            tr.roll_history()
            out = tr.desc(tr.H, join=False)       # the launch writes rows [H, H + K) only: it need not wait for the carry-over
        else:
            # nothing is recorded: the store keeps its chunk, and the n-step run counters restart so that no window can
            # straddle the unrecorded steps
            out = _lib.Transitions()
        if eps is not None:
            eps = eps.contiguous()
            assert tuple(eps.shape) == (self.K, self.n, self.spec.act_dim) and eps.is_cuda
        engine = engine or self.engine
        common = (self.K, self.global_step & 0xFFFFFFFF, self.n_step, self.reward_scale, self.cost_scale,
                  None if eps is None else eps.data_ptr(), 1 if deterministic else 0, C.byref(out), self.stats.data_ptr(),
                  _lib.current_stream())
        if isinstance(actor, GeneralActor):
            engine = "general"
        elif engine == "general":
            raise ValueError("rollout engine 'general' takes a GeneralActor")
        if engine == "tc":
            w1p, w2p = actor.tc_images()
        if timing is not None:
            timing[0].record()
        if engine == "general":
            # per step: policy forward (one GEMM launch per layer) + one msacl_rollout_step launch (sample, env step,
            # transition record, n-step bookkeeping) -- the unfused form of what the rollout kernels do in one launch
            lib = _lib.load()
            for k in range(self.K):
                logits = actor.forward_state(self.state)
                if write:
                    full = tr.desc(tr.H + k, join=False)
                    if tr.logits is not None:
                        full.logits = tr.logits[k].data_ptr()
                else:
                    full = _lib.Transitions()
                _lib.check(lib.msacl_rollout_step(C.byref(self.state.desc), logits.data_ptr(), actor.min_log_std, actor.max_log_std,
                                                  (self.global_step + k) & 0xFFFFFFFF, self.n_step, self.reward_scale, self.cost_scale,
                                                  None if eps is None else eps[k].data_ptr(), 1 if deterministic else 0,
                                                  C.byref(full), self.stats.data_ptr(), _lib.current_stream()))
        elif engine == "tc":
            _lib.check(_lib.load().msacl_rollout_fused_tc(C.byref(self.state.desc), C.byref(actor.desc), w1p.data_ptr(),
                                                         w2p.data_ptr(), *common))
        elif engine == "ffma":
            _lib.check(_lib.load().msacl_rollout_fused(C.byref(self.state.desc), C.byref(actor.desc), *common))
        else:
            raise ValueError(f"unknown rollout engine {engine!r}")
        if timing is not None:
            timing[1].record()
        tr.join_history()
        self.global_step += self.K
        if not write:
            self.state.run.zero_()
            return None
        return DeviceWindowBatch(tr, self.n_step)


class B200NstepOffSampler:
    """Reference-compatible constructor: B200NstepOffSampler(**kwargs) with the kwargs of
    RL/trainer/sampler/base.py:55-95 (env_name, env_num, sample_batch_size, reward_scale,
    cost_scale, noise_params, n_step, ...).  `.networks` may be assigned by the trainer
    (RL/trainer/nstep_off_serial_trainer.py:34) and is re-read on every `sample()`."""

    def __init__(self, **kwargs):
        self.env_id = kwargs["env_name"]
        self.num_envs = int(kwargs["env_num"])
        if kwargs.get("noise_params") is not None:
            raise RuntimeError("additive exploration noise is not part of the fused rollout (reference default: None)")
        if kwargs.get("action_type", "continu") != "continu":
            raise RuntimeError("Only continuous action space is supported!")
        self.horizon = int(kwargs["sample_batch_size"])
        self.sample_batch_size = self.horizon * self.num_envs
        self.n_step = int(kwargs.get("n_step", 1))
        self.reward_scale, self.cost_scale = kwargs["reward_scale"], kwargs["cost_scale"]
        self.device = torch.device(kwargs.get("device", "cuda"))
        self.envs = B200VectorEnv(self.env_id, self.num_envs, env_seed=kwargs.get("env_seed") or 0,
                                  env_base=kwargs.get("env_base", 0), device=self.device)
        self.obs_dim = self.envs.single_observation_space.shape
        self.act_dim = self.envs.single_action_space.shape
        self.networks = kwargs.get("networks")
        self.total_sample_number = 0
        # actor engine: "tc" = tcgen05 split-bf16 tensor cores (default), "ffma" = FP32 FFMA (bit-closer to torch fp32)
        chunks = kwargs.get("history_chunks")
        if chunks is None and kwargs.get("buffer_name") == "b200_indexed_replay_buffer":
            # the transition store IS the replay payload: retain enough chunks for buffer_max_size windows
            from .buffer import B200IndexedReplayBuffer
            chunks = B200IndexedReplayBuffer.chunks_for(int(kwargs["buffer_max_size"]), self.num_envs, self.horizon)
        self.rollout = FusedRollout(self.env_id, self.num_envs, self.horizon, self.n_step, self.reward_scale, self.cost_scale,
                                    device=self.device, state=self.envs.state, engine=kwargs.get("rollout_engine", "tc"),
                                    history_chunks=chunks)
        self._explicit_engine = kwargs.get("rollout_engine") in ("tc", "ffma")     # then an unsupported policy is an error
        self.envs.state.reset()        # base.py:98  envs.reset(seed=None)
        self._actor = None
        self._actor_cache = None

    @property
    def obs(self):
        return self.envs.state.obs

    def get_total_sample_num(self):
        return self.total_sample_number

    get_total_sample_number = get_total_sample_num

    def load_state_dict(self, state_dict):
        self.networks.load_state_dict(state_dict)

    def set_actor(self, actor: ActorWeights):
        self._actor = actor

    def _current_actor(self):
        """Packed device copy of `networks.policy`, rebuilt (transpose + split-bf16 operand images) only when a parameter
        changed since the last call -- at the reference's default env_num=4 the repack would otherwise dominate sample()."""
        if self.networks is None:
            return self._actor
        pol = self.networks.policy
        ver = (id(pol), ActorWeights.policy_version(pol))
        if self._actor_cache is None or self._actor_cache[0] != ver:
            # (a policy the fused kernels are not specialised for runs as per-layer GEMMs + msacl_rollout_step, unless the
            #  caller asked for a fused engine by name)
            self._actor_cache = (ver, actor_from_policy(pol, device=self.device, strict=self._explicit_engine))
        return self._actor_cache[1]

    def _sample(self):
        actor = self._current_actor()
        if actor is None:
            raise RuntimeError("sampler has no policy: assign `.networks` or call set_actor()")
        return self.rollout.run(actor)

    def sample(self):
        self.total_sample_number += self.sample_batch_size
        start = time.perf_counter()
        data = self._sample()
        torch.cuda.current_stream().synchronize()
        return data, {SAMPLER_TIME_TAG: (time.perf_counter() - start) * 1000}
