"""The autograd-free learner (csrc/mlp_tc.cu, csrc/learner.cu, learner.py) against float64 / PyTorch fp32 references
of the same ops, the NumPy oracle and the reference's recorded tensors.

Stated tolerances: the default "bf16x6" GEMM (3-term bf16 operand split, 6 cross products, FP32 accumulation) is
FP32-class: <= 1.5e-6 * sum|a||b| per output (measured 7e-7); the optional "bf16x3" mode carries ~2^-16 per product (asserted 5e-5);
gradients through three layers 1e-4 relative to the gradient scale (backward kernels vs float64 through the kernel's own
activations); elementwise distribution kernels 2e-5 (CUDA tanhf/logf/atanhf vs torch / NumPy)."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import actor as oactor

pytestmark = pytest.mark.gpu


def _gemm(**kw):
    import msacl_b200
    from msacl_b200 import _lib
    from msacl_b200.learner import _desc
    lib = msacl_b200.load_library()
    _lib.check(lib.msacl_gemm_tc(C.byref(_desc(**kw)), _lib.current_stream()))


def _rel(got, want):
    return float((got.double() - want).abs().max() / want.abs().max())


@pytest.mark.parametrize("prec,tol", [(6, 1.5e-6), (3, 5e-5)])
@pytest.mark.parametrize("m,n,k", [(128, 256, 256), (5120, 256, 4), (1000, 256, 6), (333, 1, 256), (4097, 4, 256), (260, 256, 16),
                                   (129, 8, 33), (640, 300, 100)])
def test_gemm_tc_forward_bias_act(m, n, k, prec, tol):
    """C = act(A B^T + bias) for the layer shapes of the learner (and ragged ones), all three activations, both precisions:
    bf16x6 (3-term operand split) is FP32-class (1.5e-6 of sum|a||b|: the tensor core aligns and truncates the addends of a k-step, measured 7e-7), bf16x3 5e-5."""
    g = torch.Generator(device="cuda").manual_seed(m + n + k)
    A = torch.randn(m, k, device="cuda", generator=g)
    Bm = torch.randn(n, k, device="cuda", generator=g) / max(1.0, k ** 0.5)
    bias = torch.randn(n, device="cuda", generator=g)
    for act, fn in ((0, lambda x: x), (1, torch.relu), (2, torch.tanh)):
        Cm = torch.full((m, n), 7.0, device="cuda")
        ss = torch.zeros(m, device="cuda") if n <= 256 else None
        _gemm(a=A, a_rs=k, a_ks=1, b=Bm, b_rs=k, b_ks=1, m=m, n=n, k=k, c=Cm, ldc=n, bias=bias, act=act, row_sumsq=ss, precision=prec)
        pre = A.double() @ Bm.double().t() + bias.double()
        want = fn(pre)
        scale = (A.double().abs() @ Bm.double().abs().t()).max()
        assert float((Cm.double() - want).abs().max()) <= tol * float(scale) + 2e-7 * float(want.abs().max()), (act, m, n, k)
        if ss is not None:
            np.testing.assert_allclose(ss.cpu().numpy(), (Cm.double() ** 2).sum(1).cpu().numpy(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("prec", [6, 3])
@pytest.mark.parametrize("transposed_b", [False, True])
def test_gemm_tc_prepacked_weights_bit_identical(prec, transposed_b):
    """Large row counts stream a pre-converted B operand (msacl_gemm_pack_b) with bulk copies: same bf16 images, so the result
    is bit-identical to the path that converts B inside every CTA -- forward (B = W) and dgrad (B = W^T read strided)."""
    import msacl_b200
    from msacl_b200 import _lib
    from msacl_b200.learner import _desc
    lib = msacl_b200.load_library()
    g = torch.Generator(device="cuda").manual_seed(11)
    m, n, k = 148 * 128 + 77, 256, 256
    A = torch.randn(m, k, device="cuda", generator=g)
    W = torch.randn(256, 256, device="cuda", generator=g) / 16
    bias = torch.randn(n, device="cuda", generator=g)
    b_rs, b_ks = (1, 256) if transposed_b else (256, 1)
    out = []
    for use_pack in (False, True):
        Cm = torch.empty(m, n, device="cuda")
        d = _desc(a=A, a_rs=k, a_ks=1, b=W, b_rs=b_rs, b_ks=b_ks, m=m, n=n, k=k, c=Cm, ldc=n, bias=bias, act=1, precision=prec)
        if use_pack:
            packed = torch.empty(int(lib.msacl_gemm_packed_b_bytes(k, prec)), dtype=torch.uint8, device="cuda")
            _lib.check(lib.msacl_gemm_pack_b(C.byref(d), packed.data_ptr(), _lib.current_stream()))
            d.b_packed = packed.data_ptr()
        _lib.check(lib.msacl_gemm_tc(C.byref(d), _lib.current_stream()))
        out.append(Cm)
    assert torch.equal(out[0], out[1])
    Wm = W.t() if transposed_b else W
    want = torch.relu(A.double() @ Wm.double().t() + bias.double())
    assert _rel(out[1], want) < (4e-6 if prec == 6 else 5e-5)


def test_gemm_tc_strided_operands_dgrad_mask_and_splitk_wgrad():
    """dgrad (B read transposed, activation-derivative mask in the epilogue) and wgrad (A and B read transposed, K = rows
    split over CTAs, partials reduced in order) vs float64."""
    import msacl_b200
    from msacl_b200 import _lib
    lib = msacl_b200.load_library()
    g = torch.Generator(device="cuda").manual_seed(5)
    rows, nin, nout = 3000, 256, 6
    dY = torch.randn(rows, nout, device="cuda", generator=g)
    W = torch.randn(nout, nin, device="cuda", generator=g) / 16
    H = torch.tanh(torch.randn(rows, nin, device="cuda", generator=g))
    dX = torch.empty(rows, nin, device="cuda")
    _gemm(a=dY, a_rs=nout, a_ks=1, b=W, b_rs=1, b_ks=nin, m=rows, n=nin, k=nout, c=dX, ldc=nin, mask=H, mask_ld=nin, mask_act=2)
    want = (dY.double() @ W.double()) * (1 - H.double() ** 2)
    assert _rel(dX, want) < 5e-5
    Hr = torch.relu(torch.randn(rows, nin, device="cuda", generator=g))
    _gemm(a=dY, a_rs=nout, a_ks=1, b=W, b_rs=1, b_ks=nin, m=rows, n=nin, k=nout, c=dX, ldc=nin, mask=Hr, mask_ld=nin, mask_act=1)
    want = (dY.double() @ W.double()) * (Hr.double() > 0)
    assert _rel(dX, want) < 5e-5
    # wgrad: dW[nout][nin] = dY^T X, split over 7 CTAs along the rows
    S = 7
    X = torch.randn(rows, nin, device="cuda", generator=g)
    parts = torch.full((S, nout, nin), 3.0, device="cuda")
    _gemm(a=dY, a_rs=1, a_ks=nout, b=X, b_rs=1, b_ks=nin, m=nout, n=nin, k=rows, c=parts, ldc=nin, split_k=S, c_split_stride=nout * nin)
    dW = torch.empty(nout, nin, device="cuda")
    _lib.check(lib.msacl_reduce_splits(parts.data_ptr(), nout * nin, S, dW.data_ptr(), _lib.current_stream()))
    want = dY.double().t() @ X.double()
    scale = (dY.double().abs().t() @ X.double().abs()).max()
    assert float((dW.double() - want).abs().max()) <= 5e-5 * float(scale)
    assert torch.equal(dW, parts.sum(0)) or _rel(dW, parts.double().sum(0)) < 1e-6
    # bias gradient: column sums over the same row splits
    cs = torch.empty(S, nout, device="cuda")
    _lib.check(lib.msacl_colsum(dY.data_ptr(), rows, nout, nout, S, cs.data_ptr(), _lib.current_stream()))
    np.testing.assert_allclose(cs.sum(0).cpu().numpy(), dY.double().sum(0).cpu().numpy(), rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("din,dout,act,rows,hidden", [(6, 1, "relu", 5120, (256, 256)), (4, 4, "relu", 777, (256, 256)),
                                                      (4, 256, "tanh", 2 * 640, (256, 256)), (16, 1, "relu", 130, (256, 256)),
                                                      (7, 2, "relu", 1500, (64, 128)), (12, 256, "tanh", 900, (300, 40)),
                                                      (3, 1, "relu", 20000, (512, 32))])
def test_fused_mlp_forward_backward_vs_autograd(din, dout, act, rows, hidden):
    """FusedMLP (three msacl_gemm_tc layers forward, dgrad + split-K wgrad backward) vs torch autograd on the same nn.Sequential;
    hidden widths are arbitrary (the reference builds them from `*_hidden_sizes`, mlp.py:18-33)."""
    from msacl_b200.algorithm import mlp
    from msacl_b200.learner import FusedMLP
    torch.manual_seed(rows)
    seq = mlp([din, *hidden, dout], {"relu": torch.nn.ReLU, "tanh": torch.nn.Tanh}[act]).cuda()
    x = torch.randn(rows, din, device="cuda")
    dy = torch.randn(rows, dout, device="cuda") / rows
    fm = FusedMLP(seq, "cuda")
    ws = fm.workspace("t", rows, train=True, need_dx=True, sumsq=(dout == 256))
    y = fm.forward(x, ws)
    xr = x.clone().requires_grad_(True)
    yr = seq(xr)
    assert _rel(y, yr.detach().double()) < 1e-4
    if dout == 256:
        np.testing.assert_allclose(ws.v.cpu().numpy(), (yr.detach() ** 2).sum(-1).cpu().numpy(), rtol=2e-4)
    dx = fm.backward(ws, dy, wgrad=True, need_dx=True)
    # (a) backward kernels vs float64 back-propagation through the kernel's OWN saved activations: same masks on both sides,
    #     so the comparison is not disturbed by hidden units whose pre-activation sits within round-off of the ReLU kink
    (W1, b1), (W2, b2), (W3, b3) = [(w.double(), b.double()) for w, b in fm.layers]
    h1, h2, xd, dyd = ws.h1.double(), ws.h2.double(), x.double(), dy.double()
    dact = (lambda h: (h > 0).double()) if act == "relu" else (lambda h: 1 - h * h)
    dA2 = (dyd @ W3) * dact(h2)
    dA1 = (dA2 @ W2) * dact(h1)
    want = [dA1.t() @ xd, dA1.sum(0), dA2.t() @ h1, dA2.sum(0), dyd.t() @ h2, dyd.sum(0)]
    assert _rel(dx, dA1 @ W1) < 1e-4
    for got, w, p in zip(fm.reduced_grads(ws), want, seq.parameters()):
        assert _rel(got, w) < 1e-4, tuple(p.shape)
    # (b) vs torch autograd end to end: all but the few rows with a unit on the kink agree to 2e-4 of the gradient scale
    yr.backward(dy)
    err = (dx.double() - xr.grad.double()).abs().max(dim=1).values / xr.grad.double().abs().max()
    assert float((err < 2e-4).double().mean()) > 0.99
    for got, p in zip(fm.reduced_grads(ws), seq.parameters()):
        assert _rel(got, p.grad.double()) < 2e-3, tuple(p.shape)


def test_tanh_gauss_kernels_vs_torch_oracle_and_reference_golden():
    """rsample / log_prob forward (a9: act_distribution_cls.py:59-84) vs the oracle and the reference's recorded log_prob;
    log_prob backward vs torch autograd of the same expression."""
    import msacl_b200
    from msacl_b200 import _lib
    from msacl_b200.algorithm import TanhGauss
    lib = msacl_b200.load_library()
    g = load_golden("msacl_targets_TwoLink.npz")
    lo, hi = torch.as_tensor(g["act_low"]).cuda(), torch.as_tensor(g["act_high"]).cuda()
    mean, std, act = (torch.as_tensor(g[k]).cuda().reshape(-1, 2) for k in ("pi_mean", "pi_std", "act"))
    rows, A = mean.shape
    logits = torch.cat([mean, std.log()], dim=-1).contiguous()
    st = _lib.current_stream()
    logp = torch.empty(rows, device="cuda")
    _lib.check(lib.msacl_tanh_gauss_log_prob(rows, A, logits.data_ptr(), act.contiguous().data_ptr(), lo.data_ptr(), hi.data_ptr(), -20.0, 1.0,
                                             logp.data_ptr(), st))
    np.testing.assert_allclose(logp.cpu().numpy(), g["logp_new"].reshape(-1), rtol=2e-5, atol=2e-5)       # the reference itself
    want = oactor.tanh_gauss_log_prob(mean.cpu().numpy(), std.cpu().numpy(), act.cpu().numpy(), g["act_low"], g["act_high"])
    np.testing.assert_allclose(logp.cpu().numpy(), want, rtol=2e-5, atol=2e-5)
    # backward vs autograd (torch fp32 reference of the same op)
    lg = logits.clone().requires_grad_(True)
    dist = TanhGauss(torch.cat([lg[:, :A], torch.clamp(lg[:, A:], -20.0, 1.0).exp()], dim=-1), lo, hi)
    gl = torch.randn(rows, device="cuda")
    (dist.log_prob(act) * gl).sum().backward()
    dlog = torch.empty_like(logits)
    _lib.check(lib.msacl_tanh_gauss_log_prob_bwd(rows, A, logits.data_ptr(), act.contiguous().data_ptr(), lo.data_ptr(), hi.data_ptr(),
                                                 -20.0, 1.0, gl.data_ptr(), 0, dlog.data_ptr(), st))
    np.testing.assert_allclose(dlog.cpu().numpy(), lg.grad.cpu().numpy(), rtol=2e-4, atol=2e-4 * float(lg.grad.abs().max()))
    # rsample forward vs the oracle
    eps = torch.randn(rows, A, device="cuda")
    a_out, lp_out = torch.empty(rows, A, device="cuda"), torch.empty(rows, device="cuda")
    _lib.check(lib.msacl_tanh_gauss_rsample(rows, A, logits.data_ptr(), eps.data_ptr(), lo.data_ptr(), hi.data_ptr(), -20.0, 1.0,
                                            a_out.data_ptr(), lp_out.data_ptr(), st))
    wa, wl, _ = oactor.tanh_gauss_sample(mean.cpu().numpy(), std.cpu().numpy(), eps.cpu().numpy(), g["act_low"], g["act_high"])
    np.testing.assert_allclose(a_out.cpu().numpy(), wa, rtol=0, atol=2e-5 * float((hi - lo).max()))
    ok = (np.abs(wa - (g["act_high"] + g["act_low"]) / 2) < 0.4995 * (g["act_high"] - g["act_low"])).all(axis=1)
    np.testing.assert_allclose(lp_out.cpu().numpy()[ok], wl[ok], rtol=1e-4, atol=1e-3)


def test_adam_multi_matches_torch_adam():
    """Three msacl_adam_multi steps (split gradient partials) vs torch.optim.Adam on identical gradients."""
    from msacl_b200.learner import FusedAdam
    torch.manual_seed(0)
    shapes = [(256, 6), (256,), (256, 256), (1, 256), (1,)]
    pa = [torch.nn.Parameter(torch.randn(*s, device="cuda")) for s in shapes]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    oa, ob = torch.optim.Adam(pa, lr=1e-3), torch.optim.Adam(pb, lr=1e-3)
    fa = FusedAdam(oa, pa)
    for step in range(3):
        S = 3
        parts = [torch.randn(S, *s, device="cuda") * (10.0 ** -step) for s in shapes]
        for p, g in zip(pb, parts):
            p.grad = (g[0] + g[1]) + g[2]
        ob.step()
        fa.step([(g, S) for g in parts])
        for a, b in zip(pa, pb):
            np.testing.assert_allclose(a.detach().cpu().numpy(), b.detach().cpu().numpy(), rtol=2e-6, atol=1e-7)
        assert float(oa.state[pa[0]]["step"]) == step + 1
    np.testing.assert_allclose(oa.state[pa[2]]["exp_avg_sq"].cpu().numpy(), ob.state[pb[2]]["exp_avg_sq"].cpu().numpy(), rtol=2e-6, atol=1e-12)


def test_learner_engine_falls_back_for_networks_outside_the_fused_family():
    """Three hidden layers / GELU: the default learner_engine switches to the autograd engine (with a warning) instead of
    failing at the first update; the reference accepts any `*_hidden_sizes` / `*_hidden_activation`."""
    import msacl_b200
    from msacl_b200.specs import get_spec
    spec = get_spec("VanderPol")
    D, A, n, B = spec.obs_dim, spec.act_dim, 5, 16
    alg = msacl_b200.create_alg(algorithm="msacl", env_name="VanderPol", obs_dim=D, act_dim=A, n_step=n, action_low_limit=spec.act_low,
                                action_high_limit=spec.act_high, q_learning_rate=1e-3, lyapunov_learning_rate=1e-3, policy_learning_rate=3e-4,
                                alpha_learning_rate=1e-3, value_hidden_sizes=[32, 32, 32], policy_hidden_activation="gelu")
    r = lambda *s: torch.randn(*s, device="cuda")
    data = dict(obs=r(B, n, D) * 0.3, obs2=r(B, n, D) * 0.3, act=torch.zeros(B, n, A, device="cuda"), rew=-torch.rand(B, n, device="cuda"),
                cost=torch.rand(B, n, device="cuda"), done=torch.zeros(B, n, device="cuda"), logp=r(B, n) - 1.0)
    with pytest.warns(UserWarning, match="learner_engine='torch'"):
        info = alg.model_update(data, 2)
    assert alg.engine_name == "torch" and info is not None and np.isfinite(list(info.values())).all()


@pytest.mark.parametrize("env,B,sizes", [("TwoLink", 48, None), ("QuadTracking", 20, None), ("DuctedFan", 33, None),
                                         ("SingleTrackCar", 24, dict(value_hidden_sizes=[128, 64], lyapunov_hidden_sizes=[64, 96],
                                                                     lyapunov_output_dim=32, policy_hidden_sizes=[96, 320]))])
def test_fused_learner_matches_torch_engine(env, B, sizes):
    """Two iterations (q + Lyapunov + 2 policy + alpha updates) of the fused learner vs the autograd engine from identical
    state, batch and rsample noise: same losses (2e-4) and the same parameters after the Adam steps."""
    import msacl_b200
    from msacl_b200.specs import get_spec
    spec = get_spec(env)
    D, A, n = spec.obs_dim, spec.act_dim, 20
    kw = dict(algorithm="msacl", env_name=env, obs_dim=D, act_dim=A, n_step=n, action_low_limit=spec.act_low, action_high_limit=spec.act_high,
              q_learning_rate=1e-3, lyapunov_learning_rate=1e-3, policy_learning_rate=3e-4, alpha_learning_rate=1e-3, lya_diff_scale=10.0,
              **(sizes or {}))
    torch.manual_seed(1)
    a = msacl_b200.create_alg(learner_engine="torch", **kw)
    b = msacl_b200.create_alg(learner_engine="fused", **kw)
    b.networks.load_state_dict(a.networks.state_dict())
    g = torch.Generator(device="cuda").manual_seed(2)
    r = lambda *s: torch.randn(*s, device="cuda", generator=g)
    lo, hi = (torch.as_tensor(x).cuda() for x in (spec.act_low, spec.act_high))
    obs = r(B, n, D) * 0.4
    data = dict(obs=obs, obs2=obs * 0.9 + 0.05 * r(B, n, D), act=(lo + (hi - lo) * torch.rand(B, n, A, device="cuda", generator=g)) * 0.97,
                rew=-torch.rand(B, n, device="cuda", generator=g) * 50, cost=torch.rand(B, n, device="cuda", generator=g),
                done=(torch.rand(B, n, device="cuda", generator=g) < 0.1).float(), logp=r(B, n) - 1.0)
    def compare_params(tag):
        sa, sb = a.networks.state_dict(), b.networks.state_dict()
        for k in sa:
            x, y = sa[k].cpu().numpy(), sb[k].cpu().numpy()
            bad = (~np.isclose(y, x, rtol=2e-4, atol=4e-6)).sum()
            # Adam's early steps are sign-like (lr * g / |g|): an entry whose gradient is at round-off level may step either
            # way, so a handful of entries per tensor is allowed to differ by up to 2 lr
            assert bad <= max(3, 0.002 * x.size), (tag, k, bad, x.size, np.abs(x - y).max())
            assert np.abs(x - y).max() <= 2.5e-3

    for it in (2, 3, 4):
        noise = [r(B, n, A) for _ in range(3)]
        ta = a.model_update(data, it, noise=[x.clone() for x in noise])
        tb = b.model_update(data, it, noise=[x.clone() for x in noise])
        assert (ta is None) == (tb is None)
        if ta is not None:
            for k in ta:
                if "time" not in k.lower():
                    np.testing.assert_allclose(tb[k], ta[k], rtol=1e-3, atol=1e-4, err_msg=f"{k} @ {it}")
        if it == 2:
            compare_params("after one update")
    # three updates later the two runs have drifted by a few sign-like steps at most
    sa, sb = a.networks.state_dict(), b.networks.state_dict()
    for k in sa:
        assert float((sa[k] - sb[k]).abs().max()) <= 4e-3, k
