"""MSACL learner targets on the GPU (RL/algorithm/msacl.py:153-164, 243-252, 280-329, 383-405).

The network forwards stay in PyTorch/cuBLAS (plain library GEMMs); everything between the
network outputs and the scalar losses is done by the kernels in csrc/targets.cu, wrapped as
autograd Functions where the reference back-propagates through them.
"""

import numpy as np
import torch

from . import _lib


class Coefficients:
    """msacl.py:153-164, float32 on the device, built with the same torch expressions."""

    def __init__(self, n_step, lya_eta=0.15, retrace_lambda=0.95, alpha1=1.0, alpha2=2.0, device="cuda"):
        self.n = int(n_step)
        son = ((torch.tensor(1 - lya_eta) ** torch.arange(1, self.n + 1) * torch.tensor(alpha2 / alpha1)) ** 0.5)
        diff = torch.pow(retrace_lambda, torch.arange(self.n))
        diff = diff / torch.sum(diff)
        sl = torch.pow((1 - lya_eta), torch.arange(self.n) + 1)
        dev = torch.device(device)
        self.son, self.diff, self.sl = (x.to(torch.float32).to(dev).contiguous() for x in (son, diff, sl))
        self.alpha1, self.alpha2 = float(alpha1), float(alpha2)


def q_backup(rew, done, next_q1, next_q2, next_logp, gamma, alpha):
    """backup = rew + (1-done)*gamma*(min(Q1',Q2') - alpha*logp')   (msacl.py:249-252)."""
    args = [t.contiguous().float() for t in (rew, done, next_q1, next_q2, next_logp)]
    out = torch.empty_like(args[0])
    _lib.check(_lib.load().msacl_q_backup(out.numel(), *[a.data_ptr() for a in args], float(gamma), float(alpha),
                                         out.data_ptr(), _lib.current_stream()))
    return out


def lyapunov_risk_raw(obs, obs2, logp_new, logp_old, lya_obs, lya_obs2, coef: Coefficients, lya_diff_scale=10.0,
                      lya_positive_scale=1.0, want_labels=True):
    B, n, D = obs.shape
    f = lambda t: t.detach().contiguous().float()
    obs, obs2, logp_new, logp_old, v1, v2 = map(f, (obs, obs2, logp_new, logp_old, lya_obs, lya_obs2))
    parts = torch.empty(3, dtype=torch.float64, device=obs.device)
    g1, g2 = torch.empty_like(v1), torch.empty_like(v2)
    isc = torch.empty_like(v1) if want_labels else None
    esl = torch.empty_like(v1) if want_labels else None
    _lib.check(_lib.load().msacl_lyapunov_risk(
        B, n, D, obs.data_ptr(), obs2.data_ptr(), logp_new.data_ptr(), logp_old.data_ptr(), v1.data_ptr(), v2.data_ptr(),
        coef.son.data_ptr(), coef.diff.data_ptr(), coef.sl.data_ptr(), coef.alpha1, coef.alpha2, float(lya_diff_scale),
        float(lya_positive_scale), parts.data_ptr(), g1.data_ptr(), g2.data_ptr(),
        None if isc is None else isc.data_ptr(), None if esl is None else esl.data_ptr(), _lib.current_stream()))
    loss2 = (parts[0] + parts[1]) / (B * n) * lya_positive_scale
    loss3 = parts[2] / B * lya_diff_scale
    return dict(loss=(loss2 + loss3).float(), loss2=loss2.float(), loss3=loss3.float(), grad_lya_obs=g1, grad_lya_obs2=g2,
                is_clip=isc, esl=esl)


class _LyapunovRisk(torch.autograd.Function):
    @staticmethod
    def forward(ctx, lya_obs, lya_obs2, obs, obs2, logp_new, logp_old, coef, diff_scale, pos_scale):
        out = lyapunov_risk_raw(obs, obs2, logp_new, logp_old, lya_obs, lya_obs2, coef, diff_scale, pos_scale, want_labels=False)
        ctx.save_for_backward(out["grad_lya_obs"], out["grad_lya_obs2"])
        return out["loss"]

    @staticmethod
    def backward(ctx, grad_out):
        g1, g2 = ctx.saved_tensors
        return grad_out * g1, grad_out * g2, None, None, None, None, None, None, None


def lyapunov_risk(obs, obs2, logp_new, logp_old, lya_obs, lya_obs2, coef, lya_diff_scale=10.0, lya_positive_scale=1.0):
    """Differentiable Lyapunov risk (loss_lya of msacl.py:331): gradients flow to lya_obs / lya_obs2.
    The reference evaluates V(obs) twice (:289,:317); one forward feeding both uses gives the same
    parameter gradients."""
    return _LyapunovRisk.apply(lya_obs, lya_obs2, obs, obs2, logp_new, logp_old, coef, lya_diff_scale, lya_positive_scale)


def stability_advantage(lya_obs0, lya_obs2, coef: Coefficients):
    """-> (raw [B], batch-normalised [B]); msacl.py:392-400 (no gradient, as in the reference)."""
    v0, v2 = lya_obs0.detach().contiguous().float(), lya_obs2.detach().contiguous().float()
    B, n = v2.shape
    raw, adv = torch.empty_like(v0), torch.empty_like(v0)
    mom = torch.empty(2, dtype=torch.float64, device=v0.device)
    lib = _lib.load()
    s = _lib.current_stream()
    _lib.check(lib.msacl_stability_advantage(B, n, v0.data_ptr(), v2.data_ptr(), coef.diff.data_ptr(), coef.sl.data_ptr(),
                                            raw.data_ptr(), mom.data_ptr(), s))
    _lib.check(lib.msacl_advantage_normalize(B, raw.data_ptr(), mom.data_ptr(), adv.data_ptr(), s))
    return raw, adv


def clipped_surrogate(new_logp0, old_logp0, adv, clip_coef=0.1):
    """loss_policy_lya of msacl.py:385-387,402-405 (PPO-clipped, differentiable in new_logp0)."""
    ratio = torch.exp(new_logp0 - old_logp0)
    surr1 = ratio * adv
    surr2 = torch.clamp(ratio, 1 - clip_coef, 1 + clip_coef) * adv
    return torch.min(surr1, surr2).mean()


class PolyakUpdater:
    """Soft target update `p_targ <- p_targ * polyak + (1 - polyak) * p` over whole parameter lists in one kernel
    launch (msacl.py:445-460), bit-exact with the reference's per-tensor `mul_` / `add_` pair.  The device pointer
    tables are built once; parameters must keep their storage (optimizers update in place)."""

    def __init__(self, pairs):
        pairs = [(p, pt) for p, pt in pairs]
        if not pairs:
            raise ValueError("no parameters")
        for p, pt in pairs:
            if not (p.is_cuda and pt.is_cuda and p.dtype == pt.dtype == torch.float32 and p.is_contiguous() and pt.is_contiguous()
                    and p.numel() == pt.numel()):
                raise ValueError("PolyakUpdater needs matching contiguous CUDA float32 parameter pairs")
        self._pairs = pairs
        dev = pairs[0][0].device
        self._src = torch.tensor([p.data_ptr() for p, _ in pairs], dtype=torch.int64, device=dev)
        self._dst = torch.tensor([pt.data_ptr() for _, pt in pairs], dtype=torch.int64, device=dev)
        self._numel = torch.tensor([p.numel() for p, _ in pairs], dtype=torch.int64, device=dev)
        self._max = max(p.numel() for p, _ in pairs)
        self._ptrs = [(p.data_ptr(), pt.data_ptr()) for p, pt in pairs]

    def step(self, tau):
        if any((p.data_ptr(), pt.data_ptr()) != q for (p, pt), q in zip(self._pairs, self._ptrs)):
            raise RuntimeError("a parameter changed its storage since the PolyakUpdater was built")
        polyak = 1 - tau
        _lib.check(_lib.load().msacl_polyak_update(len(self._pairs), self._src.data_ptr(), self._dst.data_ptr(), self._numel.data_ptr(),
                                                  self._max, float(np.float32(polyak)), float(np.float32(1 - polyak)),
                                                  _lib.current_stream()))
