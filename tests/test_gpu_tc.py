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


@pytest.mark.parametrize("name,n", [("Pendulum", 70000), ("QuadTracking", 40001)])
def test_tc_rollout_multi_step_large(name, n):
    """K=6 steps in one launch on a grid that wraps over tile pairs (persistent loop, tails)."""
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
