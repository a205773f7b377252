"""GPU parity of the MSACL target kernels against the oracle (seeded) and the reference golden.
Tolerance: float32 elementwise 2e-6 relative; reductions (losses) 1e-5 relative."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import targets as otg

pytestmark = pytest.mark.gpu


def _cu(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def test_q_backup_vs_oracle_and_golden():
    from msacl_b200 import targets as tg
    g = load_golden("msacl_targets_TwoLink.npz")
    got = tg.q_backup(_cu(g["rew"]), _cu(g["done"]), _cu(g["next_q1"]), _cu(g["next_q2"]), _cu(g["next_logp"]),
                      float(g["gamma"]), float(g["alpha"])).cpu().numpy()
    want = otg.q_backup(g["rew"], g["done"], g["next_q1"], g["next_q2"], g["next_logp"], float(g["gamma"]), float(g["alpha"]))
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-5)
    loss = np.mean((g["q1"] - got) ** 2, dtype=np.float32) + np.mean((g["q2"] - got) ** 2, dtype=np.float32)
    np.testing.assert_allclose(loss, g["loss_q"], rtol=1e-5)


# (all-lanes-live kernel: n = 20 / 5 / 10 -> 5 passes per tile, 24 / 12 / 3 -> 3, 32 / 16 / 8 / 2 -> 1; warp-per-window kernel:
#  n = 1, 7, 18 and obs_dim 3)
@pytest.mark.parametrize("B,n,D", [(48, 20, 4), (1000, 20, 12), (7, 5, 2), (300, 32, 7), (1001, 16, 4), (333, 24, 6), (50, 12, 2),
                                   (129, 10, 7), (77, 8, 12), (64, 3, 4), (90, 2, 2), (40, 1, 4), (65, 7, 4), (33, 18, 6), (21, 20, 3)])
def test_lyapunov_risk_vs_oracle(B, n, D):
    from msacl_b200 import targets as tg
    rng = np.random.default_rng(B)
    obs = (rng.standard_normal((B, n, D)) * 0.5).astype(np.float32)
    obs2 = (obs * rng.uniform(0.5, 1.2, (B, 1, 1)) + 0.02 * rng.standard_normal((B, n, D))).astype(np.float32)
    lpn = rng.standard_normal((B, n)).astype(np.float32)
    lpo = (lpn + 0.3 * rng.standard_normal((B, n))).astype(np.float32)
    v1 = (np.sum(obs ** 2, -1) * rng.uniform(0.5, 3.0, (B, n))).astype(np.float32)
    v2 = (np.sum(obs2 ** 2, -1) * rng.uniform(0.5, 3.0, (B, n))).astype(np.float32)
    coefs = otg.coefficients(n)
    want = otg.lyapunov_risk(obs, obs2, lpn, lpo, v1, v2, coefs)
    out = tg.lyapunov_risk_raw(_cu(obs), _cu(obs2), _cu(lpn), _cu(lpo), _cu(v1), _cu(v2), tg.Coefficients(n))
    np.testing.assert_allclose(out["loss"].item(), want["loss"], rtol=1e-5)
    np.testing.assert_allclose(out["is_clip"].cpu().numpy(), want["is_clip"], rtol=1e-5, atol=1e-7)
    # ESL is a sign decision: compare where the margin is not at round-off level
    son = coefs[0]
    margin = np.abs(np.sqrt((obs[:, 0] ** 2).sum(-1))[:, None] * son[None] - np.sqrt((obs2 ** 2).sum(-1)))
    ok = margin > 1e-5
    assert np.array_equal(out["esl"].cpu().numpy()[ok], want["esl"][ok])
    rows = ok.all(axis=1)
    np.testing.assert_allclose(out["grad_lya_obs"].cpu().numpy()[rows], want["grad_lya_obs"][rows], rtol=2e-5, atol=1e-9)
    np.testing.assert_allclose(out["grad_lya_obs2"].cpu().numpy()[rows], want["grad_lya_obs2"][rows], rtol=2e-5, atol=1e-9)


def test_lyapunov_risk_reference_golden_loss_and_autograd():
    from msacl_b200 import targets as tg
    g = load_golden("msacl_targets_TwoLink.npz")
    v1 = _cu(g["lya_obs"]).requires_grad_(True)
    v2 = _cu(g["lya_obs2"]).requires_grad_(True)
    loss = tg.lyapunov_risk(_cu(g["obs"]), _cu(g["obs2"]), _cu(g["logp_new"]), _cu(g["logp"]), v1, v2, tg.Coefficients(20))
    np.testing.assert_allclose(loss.item(), g["loss_lya"], rtol=1e-5)
    loss.backward()
    coefs = (g["coef_start_obs_norm"], g["coef_lya_diff"], g["coef_start_lya"])
    want = otg.lyapunov_risk(g["obs"], g["obs2"], g["logp_new"], g["logp"], g["lya_obs"], g["lya_obs2"], coefs)
    np.testing.assert_allclose(v1.grad.cpu().numpy(), want["grad_lya_obs"], rtol=2e-5, atol=1e-9)
    np.testing.assert_allclose(v2.grad.cpu().numpy(), want["grad_lya_obs2"], rtol=2e-5, atol=1e-9)


def test_coefficients_match_reference():
    from msacl_b200 import targets as tg
    g = load_golden("msacl_targets_TwoLink.npz")
    c = tg.Coefficients(20)
    np.testing.assert_allclose(c.son.cpu().numpy(), g["coef_start_obs_norm"], rtol=2e-6)
    np.testing.assert_allclose(c.diff.cpu().numpy(), g["coef_lya_diff"], rtol=2e-6)
    np.testing.assert_allclose(c.sl.cpu().numpy(), g["coef_start_lya"], rtol=2e-6)


@pytest.mark.parametrize("B", [2, 48, 100000])
def test_stability_advantage_vs_oracle(B):
    from msacl_b200 import targets as tg
    n = 20
    rng = np.random.default_rng(B)
    v0 = rng.uniform(0.1, 3, B).astype(np.float32)
    v2 = rng.uniform(0.1, 3, (B, n)).astype(np.float32)
    raw, adv = otg.stability_advantage(v0, v2, otg.coefficients(n))
    got_raw, got = tg.stability_advantage(_cu(v0), _cu(v2), tg.Coefficients(n))
    np.testing.assert_allclose(got_raw.cpu().numpy(), raw, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(got.cpu().numpy(), adv, rtol=2e-4, atol=2e-5)


def test_policy_loss_reference_golden():
    from msacl_b200 import targets as tg
    g = load_golden("msacl_targets_TwoLink.npz")
    _, adv = tg.stability_advantage(_cu(g["pol_lya_obs0"]), _cu(g["pol_lya_obs2"]), tg.Coefficients(20))
    new_logp0 = _cu(g["pol_new_logp0"]).requires_grad_(True)
    loss_lya = tg.clipped_surrogate(new_logp0, _cu(g["logp"][:, 0]), adv, 0.1)
    np.testing.assert_allclose(-g["policy_q_term"] - loss_lya.item(), g["loss_policy"], rtol=1e-5, atol=1e-5)
    loss_lya.backward()
    coefs = (g["coef_start_obs_norm"], g["coef_lya_diff"], g["coef_start_lya"])
    _, oadv = otg.stability_advantage(g["pol_lya_obs0"], g["pol_lya_obs2"], coefs)
    _, ograd = otg.clipped_surrogate(g["pol_new_logp0"], g["logp"][:, 0], oadv, 0.1)
    np.testing.assert_allclose(new_logp0.grad.cpu().numpy(), ograd, rtol=1e-4, atol=1e-7)


def test_polyak_update_multi_tensor_bit_exact():
    """One-launch soft target update vs the reference's per-tensor `mul_` / `add_` pair (msacl.py:445-460)."""
    from msacl_b200.targets import PolyakUpdater
    g = torch.Generator(device="cuda").manual_seed(5)
    shapes = [(256, 6), (256,), (256, 256), (1, 256), (1,), (257, 3), (65536 + 3,)]
    src = [torch.randn(*s, device="cuda", generator=g) for s in shapes]
    dst = [torch.randn(*s, device="cuda", generator=g) for s in shapes]
    ref = [d.clone() for d in dst]
    up = PolyakUpdater(list(zip(src, dst)))
    for tau in (0.005, 0.005, 0.1):
        polyak = 1 - tau
        for p, pt in zip(src, ref):
            pt.mul_(polyak)
            pt.add_((1 - polyak) * p)
        want = otg.polyak_update([d.cpu().numpy() for d in dst], [p.cpu().numpy() for p in src], tau)   # oracle, before the step
        up.step(tau)
        for a, b, w in zip(dst, ref, want):
            assert torch.equal(a, b)
            assert np.array_equal(a.cpu().numpy(), w)
    with pytest.raises(ValueError):
        PolyakUpdater([(src[0], dst[1])])
