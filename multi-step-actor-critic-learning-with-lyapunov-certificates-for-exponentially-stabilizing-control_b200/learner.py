"""Autograd-free MSACL learner on hand-written kernels.

`FusedLearner` runs `MSACL._q_update / _lyapunov_update / _policy_update / _alpha_update / _target_update`
(RL/algorithm/msacl.py:227-460) as a fixed sequence of launches of libmsacl_b200.so:

  * every dense layer of the four networks (RL/apprfunc/mlp.py:18-52,72-88,111-136), forward, input-gradient and
    weight-gradient, is `msacl_gemm_tc`: a split-bf16 tcgen05 GEMM with the bias / activation / activation-derivative /
    sum-of-squares epilogue fused (csrc/mlp_tc.cu) -- no cuBLAS, no autograd graph;
  * TanhGauss rsample / log_prob and their analytic gradients, the critic loss gradient, the policy-loss gradient
    (reparameterised sample through min(Q1, Q2), entropy term, clipped stability-advantage surrogate) and the entropy
    coefficient step are the kernels of csrc/learner.cu; the [B, n] window targets are those of csrc/targets.cu;
  * Adam is one multi-tensor launch per optimizer (`msacl_adam_multi`), operating in place on the state tensors of the
    `torch.optim.Adam` objects of the reference's ApproxContainer (so `optimizer.state_dict()` stays meaningful);
  * alpha, the losses and the entropy never leave the device; `model_update` reads one small statistics vector back
    only on the iterations that return a tb_info dict (every `policy_frequency`-th).

The parameters remain the `nn.Linear` tensors of `networks` (same state-dict keys as the reference, checkpoints load
unchanged); the kernels read and update them in place.
"""
import ctypes as C
import math

import torch

from . import _lib
from . import targets as tg

ACT_CODE = {torch.nn.ReLU: 1, torch.nn.Tanh: 2, torch.nn.Identity: 0}
HID = 256


def _splits(rows):
    return max(1, min(74, -(-int(rows) // 256)))


class _Launcher:
    """Pre-built msacl_gemm_t descriptors (buffers are preallocated, so pointers are fixed); a call is one ctypes call."""

    def __init__(self):
        self.lib = _lib.load()

    def gemm(self, desc):
        _lib.check(self.lib.msacl_gemm_tc(C.byref(desc), _lib.current_stream()))


def _desc(a, a_rs, a_ks, b, b_rs, b_ks, m, n, k, c, ldc, split_k=1, c_split_stride=0, bias=None, act=0, mask=None, mask_ld=0,
          mask_act=0, row_sumsq=None, precision=0, b_packed=None):
    p = lambda t: None if t is None else (t if isinstance(t, int) else t.data_ptr())
    return _lib.Gemm(b_packed=p(b_packed), a=p(a), a_row_stride=a_rs, a_k_stride=a_ks, b=p(b), b_row_stride=b_rs, b_k_stride=b_ks, m=m, n=n, k=k, c=p(c),
                     ldc=ldc, split_k=split_k, c_split_stride=c_split_stride, bias=p(bias), act=act, mask_src=p(mask), mask_ld=mask_ld,
                     mask_act=mask_act, row_sumsq=p(row_sumsq), precision=precision)


class MLPWorkspace:
    """Activations / gradients of one (network, row count) pair and the GEMM descriptors over them."""

    def __init__(self, mlp, rows, train, need_dx, sumsq):
        dev = mlp.device
        f = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        self.rows, self.train = rows, train
        M, Din, Dout, a = rows, mlp.din, mlp.dout, mlp.act
        H1, H2 = mlp.h1, mlp.h2                       # hidden widths (reference default 256, 256; any width is accepted)
        (W1, b1), (W2, b2), (W3, b3) = mlp.layers
        import functools
        _desc = functools.partial(globals()["_desc"], precision=mlp.precision)
        self.h1, self.h2, self.y = f(M, H1), f(M, H2), f(M, Dout)
        self.v = f(M) if sumsq else None
        self.x = None
        # The weight operands of the forward / dgrad GEMMs are pre-converted once per call (msacl_gemm_pack_b) and streamed by
        # bulk copies instead of being re-converted from FP32 in each of the CTAs -- also for few row tiles, where the
        # streamed-weights kernel's pipelined loader wins on latency (model_update at replay batch 256: 1.14 -> 1.00 ms).
        self.packed = {}
        pack_ok = True

        def packed(tag, k, n):
            if not pack_ok or n > 256:               # a pre-packed operand covers one 256-wide column tile
                return None
            nbytes = int(mlp._l.lib.msacl_gemm_packed_b_bytes(k, mlp.precision))
            t = self.packed[tag] = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            return t

        # forward: H1 = act(X W1^T + b1), H2 = act(H1 W2^T + b2), Y = H2 W3^T + b3
        self.f1 = _desc(0, Din, 1, W1, Din, 1, M, H1, Din, self.h1, H1, bias=b1, act=a)
        self.f2 = _desc(self.h1, H1, 1, W2, H1, 1, M, H2, H1, self.h2, H2, bias=b2, act=a, b_packed=packed("f2", H1, H2))
        self.f3 = _desc(self.h2, H2, 1, W3, H2, 1, M, Dout, H2, self.y, Dout, bias=b3, act=0, row_sumsq=self.v,
                        b_packed=packed("f3", H2, Dout))
        if not (train or need_dx):
            return
        self.da2, self.da1 = f(M, H2), f(M, H1)
        # dgrad: dA2 = (dY W3) * act'(H2);  dA1 = (dA2 W2) * act'(H1);  dX = dA1 W1
        self.g3 = _desc(0, Dout, 1, W3, 1, H2, M, H2, Dout, self.da2, H2, mask=self.h2, mask_ld=H2, mask_act=a)
        self.g2 = _desc(self.da2, H2, 1, W2, 1, H1, M, H1, H2, self.da1, H1, mask=self.h1, mask_ld=H1, mask_act=a,
                        b_packed=packed("g2", H2, H1))
        self.dx = f(M, Din) if need_dx else None
        self.g1 = _desc(self.da1, H1, 1, W1, 1, Din, M, Din, H1, self.dx, Din) if need_dx else None
        if not train:
            return
        # wgrad (K = rows, split over CTAs; partials summed by the Adam kernel): dW = dY^T X, db = column sums of dY
        S = self.S = _splits(M)
        Sb = self.Sb = max(1, min(1184, -(-M // 64)))       # bias gradients: many short row blocks (memory-level parallelism)
        self.gw1, self.gb1 = f(S, H1, Din), f(Sb, H1)
        self.gw2, self.gb2 = f(S, H2, H1), f(Sb, H2)
        self.gw3, self.gb3 = f(S, Dout, H2), f(Sb, Dout)
        self.w3 = _desc(0, 1, Dout, self.h2, 1, H2, Dout, H2, M, self.gw3, H2, split_k=S, c_split_stride=Dout * H2)
        self.w2 = _desc(self.da2, 1, H2, self.h1, 1, H1, H2, H1, M, self.gw2, H1, split_k=S, c_split_stride=H2 * H1)
        self.w1 = _desc(self.da1, 1, H1, 0, 1, Din, H1, Din, M, self.gw1, Din, split_k=S, c_split_stride=H1 * Din)

    def grads(self):
        """[(partials tensor, nsplit)] in nn.Module.parameters() order: W1, b1, W2, b2, W3, b3."""
        return [(self.gw1, self.S), (self.gb1, self.Sb), (self.gw2, self.S), (self.gb2, self.Sb), (self.gw3, self.S), (self.gb3, self.Sb)]


class FusedMLP:
    """A reference-style 3-layer MLP (mlp.py:18-33: Linear-act-Linear-act-Linear-Identity, any hidden widths) driven by
    msacl_gemm_tc.  `seq` is the nn.Sequential whose parameters stay the single source of truth."""

    def __init__(self, seq, device, precision=6, sumsq_head=False):
        if precision not in (3, 6):
            raise ValueError("precision: 6 (bf16x6, FP32-class) or 3 (bf16x3)")
        self.precision = precision
        lin = [m for m in seq if isinstance(m, torch.nn.Linear)]
        other = [m for m in seq if not isinstance(m, torch.nn.Linear)]
        if len(lin) != 3 or lin[1].in_features != lin[0].out_features or lin[2].in_features != lin[1].out_features:
            raise ValueError("FusedMLP: the fused learner takes two hidden layers (any widths; reference default [256, 256])")
        self.h1, self.h2 = lin[0].out_features, lin[1].out_features
        if sumsq_head and lin[2].out_features > 256:
            raise ValueError("FusedMLP: the fused sum-of-squares head covers one 256-wide column tile (lyapunov_output_dim <= 256)")
        if type(other[0]) not in ACT_CODE or type(other[0]) is torch.nn.Identity or type(other[1]) is not type(other[0]) or \
                not isinstance(other[2], torch.nn.Identity):
            raise ValueError(f"FusedMLP: unsupported activations {[type(m).__name__ for m in other]} (relu / tanh hidden, linear output)")
        self.act = ACT_CODE[type(other[0])]
        self.device = torch.device(device)
        self._lin = lin
        self.din, self.dout = lin[0].in_features, lin[2].out_features
        self._ws = {}
        self._l = _Launcher()
        self._bind()

    def _key(self):
        return tuple(t.data_ptr() for l in self._lin for t in (l.weight, l.bias))

    def _bind(self):
        """(Re)capture the parameter storages.  `nn.Module.to()` replaces them -- e.g. the reference trainer's
        ModuleOnDevice shuffle (nstep_off_serial_trainer.py:78) moves the networks to the CPU and back every iteration --
        so descriptors built over the old storages must be dropped."""
        self.layers = [(l.weight.data, l.bias.data) for l in self._lin]
        for w, b in self.layers:
            if not (w.is_cuda and w.is_contiguous() and b.is_contiguous() and w.dtype == torch.float32):
                raise ValueError("FusedMLP needs contiguous CUDA float32 parameters (is the module on the GPU?)")
        self.params = [t for wb in self.layers for t in wb]
        self._bound = self._key()
        self._ws.clear()

    def workspace(self, tag, rows, train=False, need_dx=False, sumsq=False):
        if self._key() != self._bound:
            self._bind()
        key = (tag, int(rows), train, need_dx, sumsq)
        ws = self._ws.get(key)
        if ws is None:
            ws = self._ws[key] = MLPWorkspace(self, int(rows), train, need_dx, sumsq)
        return ws

    def forward(self, x, ws):
        """x: [rows, din] contiguous float32 CUDA.  Returns ws.y ([rows, dout]; ws.v = row sums of squares if requested)."""
        assert x.is_contiguous() and x.shape == (ws.rows, self.din) and x.dtype == torch.float32
        ws.x = x
        ws.f1.a = x.data_ptr()
        g = self._l.gemm
        self._pack(ws, ("f2", ws.f2), ("f3", ws.f3))
        g(ws.f1); g(ws.f2); g(ws.f3)
        return ws.y

    def _pack(self, ws, *descs):
        for tag, d in descs:
            if tag in ws.packed:
                _lib.check(self._l.lib.msacl_gemm_pack_b(C.byref(d), ws.packed[tag].data_ptr(), _lib.current_stream()))

    def backward(self, ws, dy, wgrad=True, need_dx=False):
        """dy: [rows, dout] gradient w.r.t. ws.y.  Fills the weight / bias gradient partials (wgrad) and/or ws.dx."""
        assert dy.is_contiguous() and dy.numel() == ws.rows * self.dout
        lib, st, l = self._l.lib, _lib.current_stream(), self._l
        ws.g3.a = dy.data_ptr()
        self._pack(ws, ("g2", ws.g2))
        l.gemm(ws.g3)
        l.gemm(ws.g2)
        if need_dx:
            l.gemm(ws.g1)
        if wgrad:
            ws.w3.a = dy.data_ptr()
            ws.w1.b = ws.x.data_ptr()
            l.gemm(ws.w3); l.gemm(ws.w2); l.gemm(ws.w1)
            _lib.check(lib.msacl_colsum(dy.data_ptr(), ws.rows, self.dout, self.dout, ws.Sb, ws.gb3.data_ptr(), st))
            _lib.check(lib.msacl_colsum(ws.da2.data_ptr(), ws.rows, self.h2, self.h2, ws.Sb, ws.gb2.data_ptr(), st))
            _lib.check(lib.msacl_colsum(ws.da1.data_ptr(), ws.rows, self.h1, self.h1, ws.Sb, ws.gb1.data_ptr(), st))
        return ws.dx

    def reduced_grads(self, ws):
        """Summed gradients [W1, b1, W2, b2, W3, b3] (tests / inspection)."""
        out = []
        for (parts, S), p in zip(ws.grads(), self.params):
            g = torch.empty_like(p)
            _lib.check(self._l.lib.msacl_reduce_splits(parts.data_ptr(), p.numel(), S, g.data_ptr(), _lib.current_stream()))
            out.append(g)
        return out


class FusedAdam:
    """One-launch Adam step over a parameter list, in place on the state of a torch.optim.Adam (defaults only)."""

    def __init__(self, optimizer, params):
        g = optimizer.param_groups[0]
        if g.get("weight_decay", 0) != 0 or g.get("amsgrad", False) or g.get("maximize", False):
            raise ValueError("FusedAdam implements torch.optim.Adam defaults (no weight decay / amsgrad / maximize)")
        self.opt, self.group, self.params = optimizer, g, list(params)
        dev = self.params[0].device
        for p in self.params:
            st = optimizer.state[p] if p in optimizer.state else None
            if not st:
                optimizer.state[p] = {"step": torch.tensor(0.0), "exp_avg": torch.zeros_like(p.data), "exp_avg_sq": torch.zeros_like(p.data)}
        self.step_count = int(optimizer.state[self.params[0]]["step"])
        # step counter and bias-correction scalars live on the device (msacl_adam_tick), so a captured CUDA graph of the update
        # advances them on every replay; `step_count` mirrors the counter on the host
        self._step_dev = torch.full((1,), self.step_count, dtype=torch.int32, device=dev)
        self._dyn = torch.zeros(2, dtype=torch.float32, device=dev)
        self._max = max(p.numel() for p in self.params)
        self._grad_key, self._g, self._s = None, None, None
        self._param_key = None
        self.lib = _lib.load()

    def _tables(self):
        key = tuple(p.data.data_ptr() for p in self.params)
        if key != self._param_key:                   # first call, or the module was moved (new parameter storages)
            dev = self.params[0].device
            t64 = lambda v: torch.tensor(v, dtype=torch.int64, device=dev)
            st = self.opt.state
            for p in self.params:                    # optimizer state follows the parameters' device
                for k in ("exp_avg", "exp_avg_sq"):
                    if st[p][k].device != dev:
                        st[p][k] = st[p][k].to(dev)
            self._p = t64(list(key))
            self._m = t64([st[p]["exp_avg"].data_ptr() for p in self.params])
            self._v = t64([st[p]["exp_avg_sq"].data_ptr() for p in self.params])
            self._n = t64([p.numel() for p in self.params])
            self._param_key = key

    def step(self, grads):
        """grads: [(partials tensor [nsplit, *param.shape], nsplit)] aligned with the parameter list."""
        self._tables()
        key = tuple((t.data_ptr(), s) for t, s in grads)
        if key != self._grad_key:
            dev = self.params[0].device
            self._g = torch.tensor([t.data_ptr() for t, _ in grads], dtype=torch.int64, device=dev)
            self._s = torch.tensor([s for _, s in grads], dtype=torch.int32, device=dev)
            self._keep = [t for t, _ in grads]
            self._grad_key = key
        b1, b2 = self.group["betas"]
        self.captured_lr = lr = self.group["lr"]
        eps = self.group["eps"]
        st = _lib.current_stream()
        _lib.check(self.lib.msacl_adam_tick(self._step_dev.data_ptr(), self._dyn.data_ptr(), lr, b1, b2, st))
        _lib.check(self.lib.msacl_adam_multi(len(self.params), self._p.data_ptr(), self._g.data_ptr(), self._m.data_ptr(), self._v.data_ptr(),
                                             self._n.data_ptr(), self._s.data_ptr(), self._max, 1 - b1, b2, 1 - b2, 0.0, 1.0, eps,
                                             self._dyn.data_ptr(), st))
        if not torch.cuda.is_current_stream_capturing():
            self.count_step()

    def count_step(self):
        """Host mirror of one executed step (called per launch, or per graph replay)."""
        self.step_count += 1
        for p in self.params:       # keep the torch optimizer's own step counters in sync (host tensors)
            self.opt.state[p]["step"].fill_(float(self.step_count))


class FusedLearner:
    """The update steps of B200MSACL on the kernels above.  `alg` supplies networks + hyper-parameters."""

    def __init__(self, alg):
        self.alg = alg
        net, dev = alg.networks, alg.device
        self.dev, self.lib = dev, _lib.load()
        pr = int(getattr(alg, "learner_precision", 6))
        self.P = FusedMLP(net.policy.policy, dev, pr)
        self.Q1, self.Q2 = FusedMLP(net.q1.q, dev, pr), FusedMLP(net.q2.q, dev, pr)
        self.Q1t, self.Q2t = FusedMLP(net.q1_target.q, dev, pr), FusedMLP(net.q2_target.q, dev, pr)
        self.L = FusedMLP(net.lyapunov.lya, dev, pr, sumsq_head=True)
        if self.P.act != 1 or self.Q1.act != 1:
            pass      # any supported activation works for the learner; the ROLLOUT kernels additionally require ReLU policies
        self.D, self.A = self.P.din, self.P.dout // 2
        self.min_ls, self.max_ls = float(net.policy.min_log_std), float(net.policy.max_log_std)
        self.adam_q1 = FusedAdam(net.q1_optimizer, list(net.q1.parameters()))
        self.adam_q2 = FusedAdam(net.q2_optimizer, list(net.q2.parameters()))
        self.adam_l = FusedAdam(net.lyapunov_optimizer, list(net.lyapunov.parameters()))
        self.adam_p = FusedAdam(net.policy_optimizer, list(net.policy.parameters()))
        self.alpha_state = torch.zeros(2, dtype=torch.float32, device=dev)
        self.alpha_steps = 0
        self._alpha_step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self._alpha_dyn = torch.zeros(2, dtype=torch.float32, device=dev)
        self._graphs, self._warm, self._static = {}, set(), {}
        self._side = [torch.cuda.Stream(dev) for _ in range(2)] if getattr(alg, "learner_streams", True) else []
        # device statistics: [0:4) critic sums, [4:7) Lyapunov loss parts, [8:11) policy sums
        self.stats = torch.zeros(16, dtype=torch.float64, device=dev)
        self.entropy = torch.zeros(1, dtype=torch.float32, device=dev)
        self._buf = {}

    # ---- small helpers
    @property
    def lo(self):      # fetched per call: module.to() replaces buffer storages
        return self.alg.networks.policy.act_low_lim.data

    @property
    def hi(self):
        return self.alg.networks.policy.act_high_lim.data

    def _tmp(self, name, *shape):
        key = (name,) + tuple(shape)
        t = self._buf.get(key)
        if t is None:
            t = self._buf[key] = torch.empty(*shape, dtype=torch.float32, device=self.dev)
        return t

    def _st(self):
        return _lib.current_stream()

    def _par(self, *fns):
        """Run independent launch sequences concurrently: fns[0] on the current stream, the others on side streams forked
        from / joined back into it with events (under CUDA-graph capture these become parallel branches of the graph).  A GEMM
        over a few thousand rows fills 40-80 of the 148 SMs, so two or three of them overlap almost perfectly."""
        if not self._side or len(fns) == 1:
            for fn in fns:
                fn()
            return
        cur = torch.cuda.current_stream()
        fork = torch.cuda.Event()
        fork.record(cur)
        joins = []
        for i, fn in enumerate(fns[1:]):
            side = self._side[i % len(self._side)]
            side.wait_event(fork)
            with torch.cuda.stream(side):
                fn()
                ev = torch.cuda.Event()
                ev.record(side)
            joins.append(ev)
        fns[0]()
        for ev in joins:
            cur.wait_event(ev)

    def _concat(self, a, b, name):
        rows = a.shape[0]
        out = self._tmp(name, rows, a.shape[1] + b.shape[1])
        _lib.check(self.lib.msacl_concat2(a.data_ptr(), a.shape[1], b.data_ptr(), b.shape[1], rows, out.data_ptr(), self._st()))
        return out

    def _rsample(self, logits, eps, name):
        rows = logits.shape[0]
        act, logp = self._tmp(name + "_a", rows, self.A), self._tmp(name + "_lp", rows)
        _lib.check(self.lib.msacl_tanh_gauss_rsample(rows, self.A, logits.data_ptr(), eps.data_ptr(), self.lo.data_ptr(), self.hi.data_ptr(),
                                                     self.min_ls, self.max_ls, act.data_ptr(), logp.data_ptr(), self._st()))
        return act, logp

    def _eps(self, eps, rows):
        if eps is None:
            return torch.randn(rows, self.A, dtype=torch.float32, device=self.dev)
        return eps.reshape(rows, self.A).contiguous().float()

    @staticmethod
    def _flat(d, B, n):
        f = lambda t: t.reshape(B * n, -1).contiguous().float() if t.dim() == 3 else t.reshape(B * n).contiguous().float()
        return {k: f(v) for k, v in d.items()}

    # ---- msacl.py:227-262
    def q_update(self, d, eps=None):
        B, n = d["rew"].shape
        M = B * n
        f = self._flat(d, B, n)
        lib, st = self.lib, self._st()
        xq = self._concat(f["obs"], f["act"], "xq")
        w1, w2 = self.Q1.workspace("q", M, train=True), self.Q2.workspace("q", M, train=True)
        wpn, wt1, wt2 = self.P.workspace("nograd", M), self.Q1t.workspace("t", M), self.Q2t.workspace("t", M)
        eps = self._eps(eps, M)
        r = {}

        def next_action():
            logits2 = self.P.forward(f["obs2"], wpn)
            r["next_act"], r["next_logp"] = self._rsample(logits2, eps, "next")
            r["xq2"] = self._concat(f["obs2"], r["next_act"], "xq2")

        self._par(next_action, lambda: self.Q1.forward(xq, w1), lambda: self.Q2.forward(xq, w2))
        q1, q2, next_logp, xq2 = w1.y, w2.y, r["next_logp"], r["xq2"]
        self._par(lambda: self.Q1t.forward(xq2, wt1), lambda: self.Q2t.forward(xq2, wt2))
        tq1, tq2 = wt1.y, wt2.y
        backup = self._tmp("backup", M)
        _lib.check(lib.msacl_q_backup_dev_alpha(M, f["rew"].data_ptr(), f["done"].data_ptr(), tq1.data_ptr(), tq2.data_ptr(),
                                                next_logp.data_ptr(), float(self.alg.gamma), self.alg.networks.log_alpha.data_ptr(),
                                                backup.data_ptr(), st))
        dq1, dq2 = self._tmp("dq1", M), self._tmp("dq2", M)
        _lib.check(lib.msacl_q_loss_grad(M, q1.data_ptr(), q2.data_ptr(), backup.data_ptr(), dq1.data_ptr(), dq2.data_ptr(),
                                         self.stats.data_ptr(), st))
        def upd(Q, w, dq, adam):
            Q.backward(w, dq)
            adam.step(w.grads())

        self._par(lambda: upd(self.Q1, w1, dq1, self.adam_q1), lambda: upd(self.Q2, w2, dq2, self.adam_q2))
        self._M_q = M

    # ---- msacl.py:265-336
    def lyapunov_update(self, d):
        B, n = d["rew"].shape
        M = B * n
        f = self._flat(d, B, n)
        lib, st, alg = self.lib, self._st(), self.alg
        logp = self._tmp("lya_logp", M)
        wpn = self.P.workspace("nograd", M)
        xl = self._tmp("xl", 2 * M, self.D)
        ws = self.L.workspace("train", 2 * M, train=True, sumsq=True)

        def new_logp():
            logits = self.P.forward(f["obs"], wpn)
            _lib.check(lib.msacl_tanh_gauss_log_prob(M, self.A, logits.data_ptr(), f["act"].data_ptr(), self.lo.data_ptr(), self.hi.data_ptr(),
                                                     self.min_ls, self.max_ls, logp.data_ptr(), self._st()))

        def lya_forward():                    # one forward of V over [obs; obs2] (the reference's second V(obs) is identical)
            xl[:M].copy_(f["obs"]); xl[M:].copy_(f["obs2"])
            self.L.forward(xl, ws)

        self._par(lya_forward, new_logp)
        z, V = ws.y, ws.v
        dV = self._tmp("dV", 2 * M)
        c = alg.coef
        _lib.check(lib.msacl_lyapunov_risk(B, n, self.D, f["obs"].data_ptr(), f["obs2"].data_ptr(), logp.data_ptr(), f["logp"].data_ptr(),
                                           V.data_ptr(), V[M:].data_ptr(), c.son.data_ptr(), c.diff.data_ptr(), c.sl.data_ptr(), c.alpha1,
                                           c.alpha2, float(alg.lya_diff_scale), float(alg.lya_positive_scale), self.stats[4:].data_ptr(),
                                           dV.data_ptr(), dV[M:].data_ptr(), None, None, st))
        dz = self._tmp("dz", 2 * M, self.L.dout)
        _lib.check(lib.msacl_sumsq_bwd(2 * M, self.L.dout, z.data_ptr(), dV.data_ptr(), dz.data_ptr(), st))
        self.L.backward(ws, dz)
        self.adam_l.step(ws.grads())
        self._B_l, self._n_l = B, n

    # ---- msacl.py:349-411 (+ :425-438)
    def policy_update(self, d, eps=None):
        B, n = d["rew"].shape
        M = B * n
        f = self._flat(d, B, n)
        lib, st, alg = self.lib, self._st(), self.alg
        la = alg.networks.log_alpha.data
        wp = self.P.workspace("train", M, train=True)
        logits = self.P.forward(f["obs"], wp)
        eps = self._eps(eps, M)
        new_act, new_logp = self._rsample(logits, eps, "new")
        xq = self._concat(f["obs"], new_act, "xqn")
        w1, w2 = self.Q1.workspace("dx", M, need_dx=True), self.Q2.workspace("dx", M, need_dx=True)
        xl = self._tmp("xlp", B + M, self.D)
        wl = self.L.workspace("nograd", B + M, sumsq=True)
        r = {}

        def advantage():      # stability advantage from V(obs_0) and V(obs2) (no gradient): one forward over [obs[:, 0]; obs2]
            xl[:B].copy_(d["obs"][:, 0].reshape(B, self.D)); xl[B:].copy_(f["obs2"])
            self.L.forward(xl, wl)
            _, r["adv"] = tg.stability_advantage(wl.v[:B], wl.v[B:].view(B, n), alg.coef)

        self._par(advantage, lambda: self.Q1.forward(xq, w1), lambda: self.Q2.forward(xq, w2))      # torch allocations stay on the main stream
        q1, q2, adv = w1.y, w2.y, r["adv"]
        dq1, dq2 = self._tmp("pdq1", M), self._tmp("pdq2", M)
        _lib.check(lib.msacl_policy_q_route(M, q1.data_ptr(), q2.data_ptr(), new_logp.data_ptr(), la.data_ptr(), dq1.data_ptr(),
                                            dq2.data_ptr(), self.stats[8:].data_ptr(), st))
        self._par(lambda: self.Q1.backward(w1, dq1, wgrad=False, need_dx=True), lambda: self.Q2.backward(w2, dq2, wgrad=False, need_dx=True))
        dx1, dx2 = w1.dx, w2.dx
        dlogits = self._tmp("dlogits", M, 2 * self.A)
        _lib.check(lib.msacl_policy_logits_grad(M, n, self.D, self.A, logits.data_ptr(), eps.data_ptr(), dx1.data_ptr(), dx2.data_ptr(),
                                                la.data_ptr(), f["act"].data_ptr(), f["logp"].data_ptr(), adv.data_ptr(), float(alg.clip_coef),
                                                self.lo.data_ptr(), self.hi.data_ptr(), self.min_ls, self.max_ls, dlogits.data_ptr(),
                                                self.stats[8:].data_ptr(), st))
        self.P.backward(wp, dlogits)
        self.adam_p.step(wp.grads())
        self._B_p, self._M_p = B, M

    def alpha_update(self):
        alg, net = self.alg, self.alg.networks
        g = net.alpha_optimizer.param_groups[0]
        b1, b2 = g["betas"]
        clamp = math.log(alg.alpha_bound) if alg.set_alpha_bound else float("inf")
        _lib.check(self.lib.msacl_adam_tick(self._alpha_step_dev.data_ptr(), self._alpha_dyn.data_ptr(), g["lr"], b1, b2, self._st()))
        _lib.check(self.lib.msacl_alpha_update(net.log_alpha.data.data_ptr(), self.stats[8:].data_ptr(), self._M_p, float(alg.target_entropy),
                                               self.alpha_state.data_ptr(), 1 - b1, b2, 1 - b2, 0.0, 1.0, g["eps"], clamp,
                                               self.entropy.data_ptr(), self._alpha_dyn.data_ptr(), self._st()))
        if not torch.cuda.is_current_stream_capturing():
            self.alpha_steps += 1

    # ---- the whole update as one CUDA graph per schedule variant
    def _sequence(self, data, do_target, do_policy, eps_list):
        """msacl.py:191-224 schedule.  eps_list: [q eps, policy eps ...] tensors or Nones (None -> torch.randn)."""
        alg = self.alg
        self.q_update(data, eps_list[0])
        if do_target:
            alg._target_update()
        self.lyapunov_update(data)
        if do_policy:
            for i in range(alg.policy_frequency):
                self.policy_update(data, eps_list[1 + i])
                if alg.auto_alpha:
                    self.alpha_update()

    def _count_replay(self, do_policy):
        for a in (self.adam_q1, self.adam_q2, self.adam_l):
            a.count_step()
        if do_policy:
            for _ in range(self.alg.policy_frequency):
                self.adam_p.count_step()
                if self.alg.auto_alpha:
                    self.alpha_steps += 1

    def update(self, data, do_target, do_policy, noise=None):
        """One model_update.  The first call of a schedule variant (batch shape, target / policy flags, explicit noise or
        not) runs eagerly (it builds the workspaces and pointer tables), the second is captured into a CUDA graph, later ones
        replay it: ~100 launches become one graph launch.  Inputs are staged into static buffers; alpha, the Adam step
        counters and the losses live on the device, so nothing in the graph depends on host state."""
        alg = self.alg
        B, n = data["rew"].shape
        n_eps = 1 + (alg.policy_frequency if do_policy else 0)
        eps_in = [None] * n_eps
        if noise is not None:
            eps_in = [noise() for _ in range(n_eps)]
        has_noise = eps_in[0] is not None
        lrs = tuple(a.group["lr"] for a in (self.adam_q1, self.adam_q2, self.adam_l, self.adam_p)) + (alg.networks.alpha_optimizer.param_groups[0]["lr"],)
        key = (B, n, bool(do_target), bool(do_policy), has_noise, lrs, self.P._key(), self.Q1._key(), self.L._key())
        if not getattr(alg, "learner_graph", True):
            return self._sequence(data, do_target, do_policy, eps_in)
        if key not in self._graphs:
            if key not in self._warm:                      # first time: eager (allocations, tables, smem attributes)
                if len(self._warm) > 64:                   # e.g. a trainer that moves the module every iteration: stay eager
                    self._warm.clear()
                self._warm.add(key)
                return self._sequence(data, do_target, do_policy, eps_in)
            static = {k: torch.empty_like(v, dtype=torch.float32).contiguous() for k, v in data.items()}
            static_eps = [torch.empty(B, n, self.A, dtype=torch.float32, device=self.dev) if has_noise else None for _ in range(n_eps)]
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            # no garbage collection while capturing: collecting an older learner's CUDAGraph objects (reference cycles
            # alg <-> learner) would call cudaGraphExecDestroy mid-capture and invalidate it
            import gc
            gc.collect()
            was_enabled = gc.isenabled()
            gc.disable()
            try:
                with torch.cuda.graph(g):
                    self._sequence(static, do_target, do_policy, static_eps)
            finally:
                if was_enabled:
                    gc.enable()
            self._graphs[key] = (g, static, static_eps)
        g, static, static_eps = self._graphs[key]
        for k, v in static.items():
            v.copy_(data[k])
        if has_noise:
            for dst, src in zip(static_eps, eps_in):
                dst.copy_(src.reshape(dst.shape))
        g.replay()
        self._count_replay(do_policy)

    def read_stats(self):
        """One device -> host read: the scalars of the reference's tb_info dict."""
        s = self.stats.tolist()
        M, B, n = self._M_q, self._B_l, self._n_l
        alg = self.alg
        loss_q = s[0] / M + s[1] / M
        loss_lya = (s[4] + s[5]) / (B * n) * alg.lya_positive_scale + s[6] / B * alg.lya_diff_scale
        loss_policy = -s[8] / self._M_p - s[10] / self._B_p
        entropy = -s[9] / self._M_p
        return dict(loss_q=loss_q, q1_mean=s[2] / M, q2_mean=s[3] / M, loss_lya=loss_lya, loss_policy=loss_policy, entropy=entropy)
