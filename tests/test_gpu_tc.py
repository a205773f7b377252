"""tcgen05 building blocks: split-bf16 GEMM accumulated in TMEM vs a float64 matmul."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("splits,tol", [(1, 2e-2), (3, 1e-4)])
def test_tc_gemm_selftest(splits, tol):
    import msacl_b200
    from msacl_b200 import _lib
    lib = msacl_b200.load_library()
    g = torch.Generator().manual_seed(0)
    A = torch.randn(128, 256, generator=g).cuda()
    W = (torch.randn(256, 256, generator=g) / 16).cuda()
    D = torch.zeros(128, 256, device="cuda")
    _lib.check(lib.msacl_selftest_tc_gemm(A.data_ptr(), W.data_ptr(), D.data_ptr(), splits, _lib.current_stream()))
    torch.cuda.synchronize()
    want = (A.double() @ W.double().t()).cpu().numpy()
    got = D.cpu().numpy()
    err = np.abs(got - want).max() / np.abs(want).max()
    assert err < tol, err


# ---- tensor-core rollout kernel vs the FP32 FFMA kernel and the oracle
from oracle import actor as oactor  # noqa: E402
from oracle import envs as oenv  # noqa: E402


def _pair(name, n, K, seed=3, **kw):
    from msacl_b200.sampler import ActorWeights, FusedRollout
    spec = oenv.SPECS[name]
    w = oactor.init_policy_weights(spec.obs_dim, spec.act_dim, seed=1)
    a = FusedRollout(name, n, K, n_step=4, seed=seed, engine="ffma", **kw)
    b = FusedRollout(name, n, K, n_step=4, seed=seed, engine="tc", **kw)
    a.state.reset(); b.state.reset()
    return a, b, ActorWeights(w), w


@pytest.mark.parametrize("name", oenv.ENV_NAMES)
def test_tc_rollout_matches_ffma_rollout_one_step(name):
    """Same state, same noise: the split-bf16 tensor-core actor must agree with the FP32 actor to
    1e-4 * action range (logits ~3e-5 relative), and everything downstream accordingly."""
    n, K = 1000, 1
    spec = oenv.SPECS[name]
    a, b, aw, _ = _pair(name, n, K, max_step=9)
    rng = np.random.default_rng(0)
    span = float((spec.act_high - spec.act_low).max())
    for t in range(12):
        eps = torch.as_tensor(rng.standard_normal((K, n, spec.act_dim)).astype(np.float32)).cuda()
        a.run(aw, eps=eps); b.run(aw, eps=eps)
        fa = {k: v[a.tr.H].cpu().numpy() for k, v in a.tr.fields().items()}
        fb = {k: v[b.tr.H].cpu().numpy() for k, v in b.tr.fields().items()}
        assert np.array_equal(fa["obs"], fb["obs"])
        np.testing.assert_allclose(fb["act"], fa["act"], rtol=0, atol=1e-4 * span)
        tol = 5e-3 if name == "QuadTracking" else 1e-3
        np.testing.assert_allclose(fb["obs2"], fa["obs2"], rtol=1e-4, atol=tol)
        near = (np.abs(fa["obs2"] - spec.obs_low) < 1e-2).any(1) | (np.abs(fa["obs2"] - spec.obs_high) < 1e-2).any(1)
        assert np.array_equal(fb["done"][~near], fa["done"][~near])
        assert np.array_equal(fb["emit"], fa["emit"]) or near.any()
        sat = np.abs(fa["act"] - (spec.act_high + spec.act_low) / 2) > 0.4995 * (spec.act_high - spec.act_low)
        ok = ~sat.any(axis=1)
        np.testing.assert_allclose(fb["logp"][ok], fa["logp"][ok], rtol=1e-3, atol=5e-3)
        # re-synchronise the TC rollout to the FP32 one (teacher forcing)
        b.state.sf.copy_(a.state.sf); b.state.sd.copy_(a.state.sd); b.state.step.copy_(a.state.step)
        b.state.episode.copy_(a.state.episode); b.state.run.copy_(a.state.run)
        b.state.ep_return.copy_(a.state.ep_return); b.state.ep_len.copy_(a.state.ep_len)
    assert int(a.state.episode.sum().item()) > 0


# 70000 / 40001: shares of 3-4 / 2-3 tiles per CTA; 148 * 128 * 7 + 700: shares of 7 and 8 tiles = rounds of 3, 2, 2 / 3, 3, 2 tiles;
# 1000: eight CTAs with one tile each; 65536: BASELINE config 2 (68 CTAs with two rounds of 2 tiles, 80 with one round of 3)
@pytest.mark.parametrize("name,n", [("Pendulum", 70000), ("QuadTracking", 40001), ("TwoLink", 148 * 128 * 7 + 700), ("DuctedFan", 1000),
                                    ("Pendulum", 65536)])
def test_tc_rollout_multi_step_large(name, n):
    """K=6 steps in one launch on a grid whose CTAs walk several rounds of tiles (persistent loop, balanced shares, tails)."""
    K = 6
    spec = oenv.SPECS[name]
    a, b, aw, _ = _pair(name, n, K)
    a.run(aw); b.run(aw)
    fa = {k: v[a.tr.H:].cpu().numpy() for k, v in a.tr.fields().items()}
    fb = {k: v[b.tr.H:].cpu().numpy() for k, v in b.tr.fields().items()}
    span = float((spec.act_high - spec.act_low).max())
    np.testing.assert_allclose(fb["act"][0], fa["act"][0], rtol=0, atol=1e-4 * span)
    same = np.ones(n, bool)
    for k in range(K):
        same &= fa["done"][k] == fb["done"][k]
        np.testing.assert_allclose(fb["act"][k][same], fa["act"][k][same], rtol=0, atol=2e-3 * span)
        np.testing.assert_allclose(fb["obs2"][k][same], fa["obs2"][k][same], rtol=1e-3, atol=2e-2)
    assert same.mean() > 0.99


def test_tc_rollout_against_reference_golden_sampler_run():
    """The reference sampler's own recorded steps (weights, obs, eps -> action, log-prob, next obs)
    replayed through the tensor-core kernel.  Tolerance (split-bf16 actor): actions 1e-4 * range,
    next observation 1e-3 (5e-3 Quad), done flags exact away from the bounds."""
    from conftest import load_golden
    from msacl_b200.sampler import ActorWeights, FusedRollout
    for name in oenv.ENV_NAMES:
        g = load_golden(f"sampler_{name}.npz")
        spec = oenv.SPECS[name]
        T, N = g["step_eps"].shape[:2]
        aw = ActorWeights([(g[f"W{i}"], g[f"b{i}"]) for i in range(3)])
        ro = FusedRollout(name, N, 1, n_step=int(g["n_step"]), max_step=int(g["max_step"]), engine="tc")
        post = {k[5:]: g[k] for k in g if k.startswith("post_")}
        prev = {k[5:]: g[k] for k in g if k.startswith("init_")}
        span = float((spec.act_high - spec.act_low).max())
        for t in range(T):
            if name == "QuadTracking":
                ro.state.set_quad_state(prev["x"], prev["v"], prev["R"], prev["Om"], t=prev["t"], Rd_last=prev["Rd_last"],
                                        obs=prev["obs"], step=prev["step"])
            else:
                ro.state.set_box_state(prev["obs"], prev["step"])
            ro.run(aw, eps=torch.as_tensor(g["step_eps"][t][None]).cuda())
            out = {k: v[ro.tr.H].cpu().numpy() for k, v in ro.tr.fields().items()}
            np.testing.assert_allclose(out["act"], g["step_act"][t], rtol=0, atol=1e-4 * span)
            tol = 5e-3 if name == "QuadTracking" else 1e-3
            np.testing.assert_allclose(out["obs2"], g["step_obs2"][t], rtol=1e-4, atol=tol)
            near = (np.abs(g["step_obs2"][t] - spec.obs_low) < 1e-2).any(1) | (np.abs(g["step_obs2"][t] - spec.obs_high) < 1e-2).any(1)
            assert np.array_equal(out["done"].astype(bool)[~near], (g["step_done"][t] > 0)[~near])
            prev = {k: post[k][t] for k in post}


def test_tc_statistics_and_determinism():
    """Episode statistics are reduced per launch; two identical launches give identical results."""
    from msacl_b200.sampler import ActorWeights, FusedRollout
    name, n, K = "Pendulum", 5000, 12
    spec = oenv.SPECS[name]
    aw = ActorWeights(oactor.init_policy_weights(spec.obs_dim, spec.act_dim, seed=1))
    outs = []
    for rep in range(2):
        ro = FusedRollout(name, n, K, n_step=4, seed=9, engine="tc")
        ro.state.reset()
        ro.run(aw)
        f = {k: v[ro.tr.H:].clone() for k, v in ro.tr.fields().items()}
        outs.append((f, ro.stats.clone()))
    for k in outs[0][0]:
        assert torch.equal(outs[0][0][k], outs[1][0][k]), k
    st = outs[0][1].cpu().numpy()
    done = outs[0][0]["done"].cpu().numpy().astype(bool)
    assert st[0] == done.sum() and st[3] + st[4] == st[0] and st[0] > 0


def test_rollout_rejects_misaligned_transition_rows():
    """obs / obs2 / act rows are written with 16- or 8-byte vector stores: a misaligned pointer must be refused by
    the C ABI (both engines) instead of faulting on the device."""
    import ctypes as C
    from msacl_b200 import _lib
    from msacl_b200.sampler import ActorWeights, FusedRollout
    from msacl_b200.specs import get_spec
    name = "TwoLink"                       # obs_dim 4 -> float4 rows, act_dim 2 -> float2 rows
    spec = get_spec(name)
    lin = [torch.nn.Linear(spec.obs_dim, 256), torch.nn.Linear(256, 256), torch.nn.Linear(256, 2 * spec.act_dim)]
    aw = ActorWeights([(l.weight, l.bias) for l in lin])
    ro = FusedRollout(name, 256, 2, n_step=3, engine="tc")
    ro.state.reset()
    ro.tr.roll_history()
    lib = _lib.load()
    w1p, w2p = aw.tc_images()
    for field, shift in (("obs", 4), ("obs2", 8), ("act", 4)):
        out = ro.tr.desc(ro.tr.H)
        setattr(out, field, getattr(out, field) + shift)
        common = (2, 0, 3, 100.0, 100.0, None, 0, C.byref(out), ro.stats.data_ptr(), _lib.current_stream())
        rc_tc = lib.msacl_rollout_fused_tc(C.byref(ro.state.desc), C.byref(aw.desc), w1p.data_ptr(), w2p.data_ptr(), *common)
        rc_ff = lib.msacl_rollout_fused(C.byref(ro.state.desc), C.byref(aw.desc), *common)
        assert rc_tc != 0 and rc_ff != 0, field
        assert b"aligned" in lib.msacl_last_error()
    torch.cuda.synchronize()
