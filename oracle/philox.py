"""Oracle (test infrastructure): Philox4x32-10 counter RNG + the uniform / normal transforms
the CUDA rollout uses for action noise and episode resets.

The reference draws action noise from torch's global CPU generator
(act_distribution_cls.py:45-47) and reset states from a per-env PCG64 reseeded from Python's
global `random` (e.g. VanderPol.py:72-81); neither stream can be reproduced on a GPU, so the
product replaces them by counter-based Philox keyed by (seed, global env id, step | episode).
This file is the bit-exact integer restatement of that generator (Salmon et al., SC'11) and
the float transforms; parity tests inject identical draws on both sides.

Counter layout (shared with csrc/philox.cuh):
  key     = (seed_lo, seed_hi)
  counter = (env_lo, env_hi, index, stream)   stream 0: action noise, index = global step
                                              stream 1+b: reset block b, index = episode
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)

STREAM_NOISE = 0
STREAM_RESET = 1


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """All arguments broadcastable uint32-valued arrays; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def u01(x):
    """uint32 -> float32 in [0,1): top 24 bits, exact."""
    return ((x >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)


def box_muller(xa, xb):
    """Two uint32 -> two float32 N(0,1): r = sqrt(-2 ln u1), u1 in (0,1]; angle 2 pi u2."""
    u1 = (((xa >> np.uint32(8)).astype(np.float32) + np.float32(1.0)) * np.float32(2.0 ** -24)).astype(np.float32)
    u2 = u01(xb)
    r = np.sqrt(np.float32(-2.0) * np.log(u1)).astype(np.float32)
    ang = 2.0 * u2.astype(np.float64)          # in units of pi (device uses sincospif)
    return ((r * np.cos(np.pi * ang).astype(np.float32)).astype(np.float32),
            (r * np.sin(np.pi * ang).astype(np.float32)).astype(np.float32))


def _split64(v):
    v = np.asarray(v, dtype=np.uint64)
    return (v & MASK), (v >> np.uint64(32))


def action_noise(seed, env_ids, step, act_dim):
    """float32 [N, act_dim] N(0,1) for global env ids at global step `step` (act_dim <= 4)."""
    e_lo, e_hi = _split64(env_ids)
    s_lo, s_hi = int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF
    x = philox4x32_10(e_lo, e_hi, np.uint64(step), np.uint64(STREAM_NOISE), s_lo, s_hi)
    z0, z1 = box_muller(x[0], x[1])
    z2, z3 = box_muller(x[2], x[3])
    return np.stack([z0, z1, z2, z3], axis=1)[:, :act_dim].astype(np.float32)


def reset_words(seed, env_ids, episode, nblocks):
    """uint32 [N, 4*nblocks] raw words for the reset of (env, episode)."""
    e_lo, e_hi = _split64(env_ids)
    s_lo, s_hi = int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF
    ep = np.asarray(episode, dtype=np.uint64)
    out = []
    for b in range(nblocks):
        out.extend(philox4x32_10(e_lo, e_hi, ep, np.uint64(STREAM_RESET + b), s_lo, s_hi))
    return np.stack(out, axis=1)


def box_reset_uniforms(seed, env_ids, episode, obs_dim):
    w = reset_words(seed, env_ids, episode, (obs_dim + 3) // 4)
    return u01(w[:, :obs_dim])


def quad_reset_draws(seed, env_ids, episode):
    """9 uniforms (words 0..8) and 3 normals (Box-Muller on words 12..15)."""
    w = reset_words(seed, env_ids, episode, 4)
    u9 = u01(w[:, 0:9])
    z0, z1 = box_muller(w[:, 12], w[:, 13])
    z2, _ = box_muller(w[:, 14], w[:, 15])
    return u9, np.stack([z0, z1, z2], axis=1).astype(np.float32)
