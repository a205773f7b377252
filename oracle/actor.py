"""Oracle (test infrastructure): StochaPolicy forward and TanhGaussDistribution in NumPy f32.

Follows RL/apprfunc/mlp.py:111-136 (StochaPolicy: Linear-ReLU-Linear-ReLU-Linear, then
mean || exp(clamp(log_std, min, max))) and RL/utils/act_distribution_cls.py:30-95
(TanhGaussDistribution.sample / log_prob / mode).  torch.distributions.Normal.log_prob is
restated from its published formula: -(v-mu)^2/(2 sigma^2) - log sigma - log sqrt(2 pi).
"""
import math

import numpy as np

f32 = np.float32
EPS = 1e-6                                  # act_distribution_cls.py:7
_ONE_PLUS_EPS = f32(1 + EPS)                # python float folded, then cast when it meets a tensor
_LOG_SQRT_2PI = f32(math.log(math.sqrt(2 * math.pi)))


def relu(x):
    return np.maximum(x, f32(0))


def mlp_forward(weights, x, hidden_act=relu):
    """weights: list of (W [out,in], b [out]) float32, torch nn.Linear convention."""
    h = x.astype(f32)
    for i, (W, b) in enumerate(weights):
        h = (h @ W.T.astype(f32) + b.astype(f32)).astype(f32)
        if i < len(weights) - 1:
            h = hidden_act(h)
    return h


def policy_forward(weights, obs, min_log_std=-20.0, max_log_std=1.0, hidden_act=relu):
    """mlp.py:132-136 -> (mean [N,A], std [N,A]).  `weights` may hold any number of layers and `hidden_act` any
    activation (mlp.py:18-33 builds Linear / activation pairs from `hidden_sizes` / `hidden_activation`)."""
    logits = mlp_forward(weights, obs, hidden_act)
    a = logits.shape[-1] // 2
    mean, log_std = logits[..., :a], logits[..., a:]
    std = np.exp(np.clip(log_std, f32(min_log_std), f32(max_log_std))).astype(f32)
    return mean.astype(f32), std


def _normal_log_prob(u, mean, std):
    var = std ** 2
    log_scale = np.log(std)
    return (-((u - mean) ** 2) / (2 * var) - log_scale - _LOG_SQRT_2PI).astype(f32)


def _sum_last(x):
    acc = x[..., 0].copy()
    for j in range(1, x.shape[-1]):
        acc = acc + x[..., j]
    return acc.astype(f32)


def tanh_gauss_sample(mean, std, eps, lo, hi):
    """act_distribution_cls.py:45-57 with the N(0,1) draw `eps` made explicit:
    torch.normal(mean, std) computes eps*std then +mean (two roundings)."""
    lo, hi = lo.astype(f32), hi.astype(f32)
    u = ((std * eps.astype(f32)).astype(f32) + mean).astype(f32)
    th = np.tanh(u).astype(f32)
    half = ((hi - lo) / f32(2)).astype(f32)
    mid = ((hi + lo) / f32(2)).astype(f32)
    act = (half * th + mid).astype(f32)
    logp = (_sum_last(_normal_log_prob(u, mean, std))
            - _sum_last(np.log(_ONE_PLUS_EPS - th ** 2).astype(f32))
            - _sum_last(np.broadcast_to(np.log(half), u.shape).astype(f32))).astype(f32)
    return act, logp, u


def tanh_gauss_log_prob(mean, std, act, lo, hi):
    """act_distribution_cls.py:73-84."""
    lo, hi = lo.astype(f32), hi.astype(f32)
    y = (f32(1 - EPS) * (f32(2) * act - (hi + lo)) / (hi - lo)).astype(f32)
    u = np.arctanh(y).astype(f32)
    half = ((hi - lo) / f32(2)).astype(f32)
    th = np.tanh(u).astype(f32)
    return (_sum_last(_normal_log_prob(u, mean, std))
            - _sum_last(np.log(half * (_ONE_PLUS_EPS - th ** 2)).astype(f32))).astype(f32)


def tanh_gauss_mode(mean, lo, hi):
    """act_distribution_cls.py:90-95."""
    lo, hi = lo.astype(f32), hi.astype(f32)
    return (((hi - lo) / f32(2)) * np.tanh(mean).astype(f32) + (hi + lo) / f32(2)).astype(f32)


def init_policy_weights(obs_dim, act_dim, hidden=(256, 256), seed=0):
    """torch nn.Linear default init distribution (U(+-1/sqrt(fan_in)) for W and b), drawn
    from NumPy so the bench/tests need no torch on the oracle side."""
    rng = np.random.default_rng(seed)
    sizes = [obs_dim, *hidden, 2 * act_dim]
    ws = []
    for i in range(len(sizes) - 1):
        bound = 1.0 / math.sqrt(sizes[i])
        W = rng.uniform(-bound, bound, size=(sizes[i + 1], sizes[i])).astype(f32)
        b = rng.uniform(-bound, bound, size=(sizes[i + 1],)).astype(f32)
        ws.append((W, b))
    return ws
