"""Oracle (test infrastructure): batched NumPy restatement of the six reference envs.

Every function works on N env instances at once (leading axis) and reproduces the dtype
flow of the reference's per-instance scalar code under NumPy 2.x.  Reference citations are
relative to /root/reference/.

State layout (dict of arrays):
  five "box" envs : obs f32[N,D], step i32[N]
  QuadTracking    : x,v,Om f32[N,3], R f32[N,3,3], t f64[N], t_last f64[N],
                    Rd_last f64[N,3,3], obs f32[N,12], step i32[N]
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

f32 = np.float32
f64 = np.float64


@dataclass(frozen=True)
class EnvSpec:
    name: str
    env_id: int
    obs_dim: int
    act_dim: int
    obs_low: np.ndarray   # f32[D]   termination box
    obs_high: np.ndarray
    act_low: np.ndarray   # f32[A]
    act_high: np.ndarray
    reset_low: np.ndarray  # f32[D] (quad: unused)
    reset_high: np.ndarray
    q: np.ndarray          # f32[D]  state cost weights
    r: np.ndarray          # f32[A]  control cost weights
    control_step: int
    origin_radius: float
    max_step: int = 1000
    dt: float = 0.01


def _a(*v):
    return np.array(v, dtype=f32)


def _box_spec(name, env_id, low, high, alow, ahigh, q, r, noise):
    low = np.asarray(low, dtype=f64).astype(f32)
    high = np.asarray(high, dtype=f64).astype(f32)
    d = low.shape[0]
    if noise is None:      # Pendulum resets over the whole box (Pendulum.py:83-86)
        rl, rh = low.copy(), high.copy()
    else:                  # others: +-reset_noise (e.g. VanderPol.py:78-81)
        rl = (-noise * np.ones(d, dtype=f32)).astype(f32)
        rh = (noise * np.ones(d, dtype=f32)).astype(f32)
    return EnvSpec(name, env_id, d, len(alow), low, high, _a(*alow), _a(*ahigh), rl, rh,
                   _a(*q), _a(*r), 5, 1e-2)


_PI = math.pi
SPECS = {
    # VanderPol.py:23-66
    "VanderPol": _box_spec("VanderPol", 0, [-10.0, -10.0], [10.0, 10.0], [-5.0], [5.0],
                           [2.0, 1.0], [0.1], 5),
    # Pendulum.py:19-68
    "Pendulum": _box_spec("Pendulum", 1, [-_PI, -10.0], [_PI, 10.0], [-5.0], [5.0],
                          [2.0, 1.0], [0.1], None),
    # DuctedFan.py:21-73
    "DuctedFan": _box_spec("DuctedFan", 2, [-5.0, -5.0, -_PI / 2, -5.0, -5.0, -5.0],
                           [5.0, 5.0, _PI / 2, 5.0, 5.0, 5.0], [-5.0, -5.0], [5.0, 5.0],
                           [2.0, 2.0, 2.0, 1.0, 1.0, 1.0], [0.1, 0.1], 0.5),
    # TwoLink.py:21-82
    "TwoLink": _box_spec("TwoLink", 3, [-_PI / 2, -_PI / 2, -20.0, -20.0],
                         [_PI / 2, _PI / 2, 20.0, 20.0], [-20.0, -20.0], [20.0, 20.0],
                         [2.0, 2.0, 1.0, 1.0], [0.1, 0.1], 0.5),
    # SingleTrackCar.py:41-113
    "SingleTrackCar": _box_spec("SingleTrackCar", 4,
                                [-1.0, -1.0, -1.066, -1.0, -_PI / 2, -_PI / 2, -_PI / 3],
                                [1.0, 1.0, 1.066, 1.0, _PI / 2, _PI / 2, _PI / 3],
                                [-5.0, -5.0], [5.0, 5.0],
                                [2.0, 2.0, 1.0, 1.0, 1.0, 1.0, 1.0], [0.1, 0.1], 0.5),
}
# QuadTracking.py:38-111 : action box [0, 2 m g] x [-10,10]^3 ; obs box +-10 ; control_step 4
_QUAD_M = 4.34
_QUAD_G = np.array([0, 0, 9.8])
SPECS["QuadTracking"] = EnvSpec(
    "QuadTracking", 5, 12, 4, -10.0 * np.ones(12, f32), 10.0 * np.ones(12, f32),
    np.array([0.0 * (_QUAD_M * _QUAD_G[2]), -10.0, -10.0, -10.0], dtype=f32),
    np.array([2.0 * (_QUAD_M * _QUAD_G[2]), 10.0, 10.0, 10.0], dtype=f32),
    np.zeros(12, f32), np.zeros(12, f32), np.ones(12, f32),
    np.array([0.0001, 0.01, 0.01, 0.01], dtype=f32), 4, 0.1)

ENV_NAMES = ["VanderPol", "Pendulum", "DuctedFan", "TwoLink", "SingleTrackCar", "QuadTracking"]


# --------------------------------------------------------------------------------------
# derivatives of the five box envs (all return the array that the reference multiplies by dt)
# --------------------------------------------------------------------------------------
def _deriv_vanderpol(o, a):
    """VanderPol.py:89-97 (_dynamics): float32 throughout."""
    x, dx = o[:, 0], o[:, 1]
    u = a[:, 0]
    d_dx = 1.0 * (1 - x ** 2) * dx - x + u
    return np.stack([dx, d_dx], axis=1)


def _deriv_pendulum(o, a):
    """Pendulum.py:93-104: theta_dd = (m g L sin(theta) - b theta_d + u)/(m L^2), float32."""
    g, L, m, b = 9.81, 0.5, 0.15, 0.1
    u = a[:, 0]
    th, thd = o[:, 0], o[:, 1]
    d2 = (m * g * L * np.sin(th) - b * thd + u) / (m * L ** 2)
    return np.stack([thd, d2], axis=1)


def _deriv_ductedfan(o, a):
    """DuctedFan.py:99-112, float32."""
    m, g, r, d, J = 8.5, 9.81, 0.26, 0.95, 0.048
    th, dx, dy, dth = o[:, 2], o[:, 3], o[:, 4], o[:, 5]
    u1, u2 = a[:, 0], a[:, 1]
    d_dx = (-m * g * np.sin(th) - d * dx + u1 * np.cos(th) - u2 * np.sin(th)) / m
    d_dy = (m * g * (np.cos(th) - 1) - d * dy + u1 * np.sin(th) + u2 * np.cos(th)) / m
    d_dth = (r * u1) / J
    return np.stack([dx, dy, dth, d_dx, d_dy, d_dth], axis=1)


def _deriv_twolink(o, a):
    """TwoLink.py:100-144.  Matrix entries are float32 scalars stored in float64 arrays
    (M22 and C22 are Python numbers, which makes np.array() pick float64); G stays float32;
    the 2x2 solve runs in float64 (LAPACK dgesv = LU with partial pivoting, restated here)."""
    l1 = l2 = m1 = m2 = 1.0
    lc1, lc2 = l1 / 2, l2 / 2
    I1 = (1 / 12) * m1 * l1 ** 2
    I2 = (1 / 12) * m2 * l2 ** 2
    g = 9.81
    th1, th2, dth1, dth2 = o[:, 0], o[:, 1], o[:, 2], o[:, 3]
    c2 = np.cos(th2)
    M11 = I1 + I2 + m1 * lc1 ** 2 + m2 * (l1 ** 2 + lc2 ** 2 + 2 * l1 * lc2 * c2)   # f32
    M12 = I2 + m2 * (lc2 ** 2 + l1 * lc2 * c2)                                       # f32
    M22 = I2 + m2 * lc2 ** 2                                                         # python f64
    s2 = np.sin(th2)
    h = -m2 * l1 * lc2 * s2
    C11 = h * dth2
    C12 = h * dth2 + h * dth1
    C21 = -h * dth1
    G1 = (-(m1 * lc1 + m2 * l1) * g * np.sin(th1) - m2 * lc2 * g * np.sin(th1 + th2))
    G2 = -m2 * lc2 * g * np.sin(th1 + th2)
    # rhs = u - C @ dq - G   (float64 because C is a float64 array)
    dq1, dq2 = dth1.astype(f64), dth2.astype(f64)
    Cdq1 = C11.astype(f64) * dq1 + C12.astype(f64) * dq2
    Cdq2 = C21.astype(f64) * dq1 + 0.0 * dq2
    b1 = (a[:, 0].astype(f64) - Cdq1) - G1.astype(f64)
    b2 = (a[:, 1].astype(f64) - Cdq2) - G2.astype(f64)
    m11, m12 = M11.astype(f64), M12.astype(f64)
    # dgesv on [[m11,m12],[m12,M22]]: |m11| > |m12| always, so no row swap
    l21 = m12 / m11
    u22 = M22 - l21 * m12
    y2 = b2 - l21 * b1
    x2 = y2 / u22
    x1 = (b1 - m12 * x2) / m11
    return np.stack([dq1, dq2, x1, x2], axis=1)  # float64, as np.concatenate([dq, ddq]) is


# SingleTrackCar.py:41-63 constants
_CAR = dict(
    p_dy1=1.0489, p_ky1=-21.92, lf=0.3048 * 3.793293, lr=0.3048 * 4.667707,
    h=0.3048 * 2.01355, m=4.4482216152605 / 0.3048 * (74.91452),
    Iz=4.4482216152605 * 0.3048 * (1321.416), g=9.81,
    v_ref=1.0, omega_ref=0.0, a_ref=0.0, mu_scale=0.1)


def _deriv_car(o, a):
    """SingleTrackCar.py:131-288: xdot = f(x) + g(x) u.  Scalars are float32 (constants fold
    in Python float64 first, exactly as the reference's expressions associate), stored into
    float64 f / g arrays; g @ u and f + g@u are float64."""
    c = _CAR
    lf, lr, hh, m, Iz, gg = c["lf"], c["lr"], c["h"], c["m"], c["Iz"], c["g"]
    v_ref, omega_ref, a_ref = c["v_ref"], c["omega_ref"], c["a_ref"]
    mu = c["mu_scale"] * c["p_dy1"]
    C_Sf = -c["p_ky1"] / c["p_dy1"]
    C_Sr = -c["p_ky1"] / c["p_dy1"]
    sxe, sye, delta, ve, psi_e, psi_e_dot, beta = (o[:, i] for i in range(7))
    n = o.shape[0]
    v = ve + v_ref
    psi_dot = psi_e_dot + omega_ref
    f = np.zeros((n, 7))
    g = np.zeros((n, 7, 2))
    f[:, 0] = v * np.cos(psi_e + beta) - v_ref + omega_ref * sye
    f[:, 1] = v * np.sin(psi_e + beta) - omega_ref * sxe
    f[:, 3] = -a_ref
    f[:, 2] = 0.0
    kin = np.abs(v) < 0.1
    dyn = ~kin
    with np.errstate(all="ignore"):
        # dynamic branch (:176-196, :241-256)
        f5 = (
            -(mu * m / (v * Iz * (lr + lf)))
            * (lf ** 2 * C_Sf * gg * lr + lr ** 2 * C_Sr * gg * lf)
            * psi_dot
            + (mu * m / (Iz * (lr + lf)))
            * (lr * C_Sr * gg * lf - lf * C_Sf * gg * lr)
            * beta
            + (mu * m / (Iz * (lr + lf))) * (lf * C_Sf * gg * lr) * delta
        )
        f6 = (
            (
                (mu / (v ** 2 * (lr + lf))) * (C_Sr * gg * lf * lr - C_Sf * gg * lr * lf)
                - 1
            )
            * psi_dot
            - (mu / (v * (lr + lf))) * (C_Sr * gg * lf + C_Sf * gg * lr) * beta
            + mu / (v * (lr + lf)) * (C_Sf * gg * lr) * delta
        )
        g51 = (
            -(mu * m / (v * Iz * (lr + lf)))
            * (-(lf ** 2) * C_Sf * hh + lr ** 2 * C_Sr * hh)
            * psi_dot
            + (mu * m / (Iz * (lr + lf))) * (lr * C_Sr * hh + lf * C_Sf * hh) * beta
            - (mu * m / (Iz * (lr + lf))) * (lf * C_Sf * hh) * delta
        )
        g61 = (
            (mu / (v ** 2 * (lr + lf))) * (C_Sr * hh * lr + C_Sf * hh * lf) * psi_dot
            - (mu / (v * (lr + lf))) * (C_Sr * hh - C_Sf * hh) * beta
            - mu / (v * (lr + lf)) * C_Sf * hh * delta
        )
        # kinematic branch (:197-204, :257-277)
        lwb = lf + lr
        kf4 = v * np.cos(beta) / lwb * np.tan(delta) - omega_ref
        beta_dot = (1 / (1 + (np.tan(delta) * lr / lwb) ** 2) * lr / (lwb * np.cos(delta) ** 2))
        kg51 = 1 / lwb * (np.cos(beta) * np.tan(delta))
        kg50 = (1 / lwb * (-v * np.sin(beta) * np.tan(delta) * beta_dot
                           + v * np.cos(beta) / np.cos(delta) ** 2))
    f[:, 4] = np.where(dyn, psi_e_dot, kf4)
    f[:, 5] = np.where(dyn, f5, 0.0)
    f[:, 6] = np.where(dyn, f6, 0.0)
    g[:, 2, 0] = np.where(dyn, 1.0, 0.0)
    g[:, 3, 1] = np.where(dyn, 1.0, 0.0)
    g[:, 5, 1] = np.where(dyn, g51, kg51)
    g[:, 6, 1] = np.where(dyn, g61, 0.0)
    g[:, 5, 0] = np.where(dyn, 0.0, kg50)
    g[:, 6, 0] = np.where(dyn, 0.0, beta_dot)
    u = a.astype(f64)
    gu = g[:, :, 0] * u[:, None, 0] + g[:, :, 1] * u[:, None, 1]
    return f + gu


_DERIV = {"VanderPol": _deriv_vanderpol, "Pendulum": _deriv_pendulum, "DuctedFan": _deriv_ductedfan,
          "TwoLink": _deriv_twolink, "SingleTrackCar": _deriv_car}


def _seq_sum(x):
    """np.sum over a short (<8) last axis = left-to-right float32 adds."""
    acc = x[:, 0].copy()
    for j in range(1, x.shape[1]):
        acc = acc + x[:, j]
    return acc


def np_pairwise_rowsum(x):
    """Row sums with the association NumPy's pairwise_sum uses for 8 <= n <= 128 when the
    reduction axis is contiguous (numpy/core/src/umath/loops_utils.h.src): eight running
    lanes, tree-combined, then the tail added sequentially.  n < 8: plain sequential."""
    n = x.shape[1]
    if n < 8:
        return _seq_sum(x)
    r = [x[:, j].copy() for j in range(8)]
    i = 8
    while i < n - (n % 8):
        for j in range(8):
            r[j] = r[j] + x[:, i + j]
        i += 8
    res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
    while i < n:
        res = res + x[:, i]
        i += 1
    return res


def box_env_step(name, state, action):
    """One reference `step()` for the five box envs (e.g. VanderPol.py:100-130):
    control_step explicit-Euler sub-steps with the state re-rounded to float32 after each,
    quadratic cost, +1 bonus inside the origin box, strict out-of-box termination,
    truncation at max_step.  Returns (new_state, obs, reward f32, terminated, truncated)."""
    spec = SPECS[name]
    o = state["obs"].astype(f32).copy()
    a = action.astype(f32)
    deriv = _DERIV[name]
    for _ in range(spec.control_step):
        d = deriv(o, a)
        # `self.obs += d * self.dt` : product in d's dtype, add in the wider type, round to f32
        o = (o + d * spec.dt).astype(f32)
    obs_cost = _seq_sum(spec.q * o ** 2)
    control_cost = _seq_sum(spec.r * a ** 2)
    cost = obs_cost + control_cost
    reward = -cost
    near = np.all(np.abs(o) <= f32(spec.origin_radius), axis=1)
    reward = np.where(near, reward + f32(1), reward).astype(f32)
    term = np.any((o < spec.obs_low) | (o > spec.obs_high), axis=1)
    step = state["step"] + 1
    trunc = step >= spec.max_step
    return {"obs": o, "step": step.astype(np.int32)}, o.copy(), reward, term, trunc


# --------------------------------------------------------------------------------------
# QuadTracking
# --------------------------------------------------------------------------------------
_QUAD_J = np.diag([0.0820, 0.0845, 0.1377])
_QUAD_KX, _QUAD_KV = 69.44, 24.304
_QUAD_RACT = np.array([0.0001, 0.01, 0.01, 0.01], dtype=f32)


def _hat(w):
    """QuadTracking.py:298-305 VecToSo3, float32."""
    n = w.shape[0]
    m = np.zeros((n, 3, 3), dtype=f32)
    m[:, 0, 1] = -w[:, 2]; m[:, 0, 2] = w[:, 1]
    m[:, 1, 0] = w[:, 2];  m[:, 1, 2] = -w[:, 0]
    m[:, 2, 0] = -w[:, 1]; m[:, 2, 1] = w[:, 0]
    return m


def _vee(m):
    """QuadTracking.py:294-297 So3ToVec: picks [2,1],[0,2],[1,0] and rounds to float32."""
    return np.stack([m[:, 2, 1], m[:, 0, 2], m[:, 1, 0]], axis=1).astype(f32)


def _polar_svd(mat):
    """QuadTracking.py:308-315 NormalizeOrientMatrix: R <- U Vh of the float32 SVD, flipping
    the last column of U if det < 0."""
    U, _, Vh = np.linalg.svd(mat)
    R = U @ Vh
    neg = np.linalg.det(R) < 0
    if np.any(neg):
        U = U.copy()
        U[neg, :, -1] *= -1
        R = np.where(neg[:, None, None], U @ Vh, R)
    return R.astype(f32)


def quad_desired(x, v, t, t_last, Rd_last, have_last):
    """QuadTracking.py:122-149 _get_desired_states with the default trajectory (:29-36).
    x,v f32[N,3]; t,t_last f64[N]; Rd_last f64[N,3,3]; have_last: bool (False only inside reset).
    Returns xd f64, vd f32, Rd f64, Omega_d (f32, or f64 zeros inside reset)."""
    n = x.shape[0]
    z = np.zeros(n)
    xd = np.stack([0.4 * t, 0.4 * np.sin(t), 0.6 * np.cos(t)], axis=1)
    b1d = np.stack([np.cos(t), np.sin(t), z], axis=1)
    vd = np.stack([0.4 + z, 0.4 * np.cos(t), -0.6 * np.sin(t)], axis=1).astype(f32)
    ad = np.stack([z, -0.4 * np.sin(t), -0.6 * np.cos(t)], axis=1).astype(f32)
    ex = (x - xd).astype(f32)
    ev = (v - vd).astype(f32)
    fd = -(-_QUAD_KX * ex - _QUAD_KV * ev - _QUAD_M * _QUAD_G + _QUAD_M * ad)   # f64
    b3d = fd / np.sqrt(np.sum(fd * fd, axis=1))[:, None]
    c = np.cross(b3d, b1d)
    b2d = c / np.sqrt(np.sum(c * c, axis=1))[:, None]
    b1n = np.cross(b2d, b3d)
    Rd = np.stack([b1n, b2d, b3d], axis=2)        # columns
    if have_last:
        dt = t - t_last
        dt = np.where(dt < 1e-6, 1e-6, dt)        # safe_time_diff :340-345
        Rd_dot = ((Rd - Rd_last) / dt[:, None, None]).astype(f32)       # RDerive :346-350
        Om_d = _vee(np.transpose(Rd, (0, 2, 1)) @ Rd_dot.astype(f64))   # getOmega :351-354
    else:
        Om_d = np.zeros((n, 3))
    return xd, vd, Rd, Om_d


def quad_errors(x, v, R, Om, xd, vd, Rd, Om_d):
    """QuadTracking.py:317-338 cal_ex / cal_ev / cal_eR / cal_eOmega -> obs f32[N,12]."""
    ex = (x - xd).astype(f32)
    ev = (v - vd).astype(f32)
    RdT = np.transpose(Rd, (0, 2, 1))
    RT = np.transpose(R, (0, 2, 1))
    eR = (_vee(RdT @ R.astype(f64) - RT.astype(f64) @ Rd) * 0.5).astype(f32)
    eOm = (Om - ((RT.astype(f64) @ Rd) @ Om_d.astype(f64)[:, :, None])[:, :, 0]).astype(f32)
    return ex, ev, eR, eOm


def quad_state_from_raw(x, v, R, Om):
    """QuadTracking.py:152-202 reset(), given the already-drawn x, v, R, Omega (float32)."""
    n = x.shape[0]
    t0 = np.zeros(n)
    xd, vd, Rd, Om_d = quad_desired(x, v, t0, t0, None, False)
    ex, ev, eR, eOm = quad_errors(x, v, R, Om, xd, vd, Rd, Om_d)
    obs = np.hstack([ex, ev, eR, eOm]).astype(f32)
    return {"x": x.astype(f32).copy(), "v": v.astype(f32).copy(), "R": R.astype(f32).copy(),
            "Om": Om.astype(f32).copy(), "t": t0.copy(), "t_last": t0.copy(),
            "Rd_last": Rd.copy(), "obs": obs, "step": np.zeros(n, np.int32)}


def quad_env_step(state, action):
    """QuadTracking.py:205-285 step()."""
    spec = SPECS["QuadTracking"]
    a = action.astype(f32)
    x, v, R, Om = (state[k].copy() for k in ("x", "v", "R", "Om"))
    force = a[:, 0]
    M = a[:, 1:]
    Jinv = np.linalg.inv(_QUAD_J)
    dt = spec.dt
    for _ in range(spec.control_step):
        dx = v
        dv = _QUAD_G - (force[:, None] * R[:, :, 2] / _QUAD_M)                  # f64
        dR = R @ _hat(Om)                                                        # f32
        JO = Om.astype(f64) @ _QUAD_J.T                                          # J @ Omega, f64
        dOm = (M - np.cross(Om, JO)) @ Jinv.T                                    # f64
        x = (x + dx * dt).astype(f32)
        v = (v + dv * dt).astype(f32)
        R = (R + dR * dt).astype(f32)
        Om = (Om + dOm * dt).astype(f32)
        R = _polar_svd(R)
    t = state["t"] + dt * spec.control_step
    xd, vd, Rd, Om_d = quad_desired(x, v, t, state["t_last"], state["Rd_last"], True)
    ex, ev, eR, eOm = quad_errors(x, v, R, Om, xd, vd, Rd, Om_d)
    obs = np.hstack([ex, ev, eR, eOm]).astype(f32)
    reward = -(_seq_sum(f32(1.0) * ex ** 2) + _seq_sum(f32(1.0) * ev ** 2)
               + _seq_sum(f32(1.0) * eR ** 2) + _seq_sum(f32(1.0) * eOm ** 2)
               + _seq_sum(_QUAD_RACT * a ** 2))
    dist = np.max(np.abs(obs), axis=1)
    bonus = 10.0 * (1 - dist / f32(spec.origin_radius))          # reward type 1, :263-267
    reward = np.where(dist <= f32(spec.origin_radius), reward + bonus, reward).astype(f32)
    term = np.any((obs < spec.obs_low) | (obs > spec.obs_high), axis=1)
    step = state["step"] + 1
    trunc = step >= spec.max_step
    new = {"x": x, "v": v, "R": R, "Om": Om, "t": t, "t_last": t.copy(), "Rd_last": Rd,
           "obs": obs, "step": step.astype(np.int32)}
    return new, obs.copy(), reward, term, trunc


def env_step(name, state, action):
    if name == "QuadTracking":
        return quad_env_step(state, action)
    return box_env_step(name, state, action)


# --------------------------------------------------------------------------------------
# resets: production resets draw from counter-based Philox (oracle.philox); the
# distribution follows each env's reset() (uniform box; quad: R = exp(hat(0.01 z)))
# --------------------------------------------------------------------------------------
def box_reset_from_uniform(name, u):
    """u f32[N,D] in [0,1) -> obs = low + (high-low)*u, float32 (distribution of e.g.
    VanderPol.py:78-81 / Pendulum.py:83-86; the reference draws float64 then rounds)."""
    spec = SPECS[name]
    span = (spec.reset_high - spec.reset_low).astype(f32)
    o = (spec.reset_low + span * u.astype(f32)).astype(f32)
    return {"obs": o, "step": np.zeros(u.shape[0], np.int32)}


def rodrigues_f32(rv):
    """exp(hat(rv)) in float32 (QuadTracking.py:180-181 uses scipy Rotation.from_rotvec)."""
    rv = rv.astype(f32)
    th2 = _seq_sum(rv * rv)
    th = np.sqrt(th2)
    small = th < f32(1e-4)
    ths = np.where(small, f32(1), th)
    A = np.where(small, f32(1) - th2 / f32(6), np.sin(ths) / ths).astype(f32)
    B = np.where(small, f32(0.5) - th2 / f32(24), (f32(1) - np.cos(ths)) / (ths * ths)).astype(f32)
    K = _hat(rv)
    K2 = K @ K
    eye = np.eye(3, dtype=f32)[None]
    return (eye + A[:, None, None] * K + B[:, None, None] * K2).astype(f32)


def quad_reset_from_draws(u9, z3):
    """u9 f32[N,9] uniforms (x,v,Omega), z3 f32[N,3] normals (rotation vector / 0.01)."""
    lo, span = f32(-0.01), f32(0.02)
    x = (lo + span * u9[:, 0:3]).astype(f32)
    v = (lo + span * u9[:, 3:6]).astype(f32)
    Om = (lo + span * u9[:, 6:9]).astype(f32)
    R = rodrigues_f32(f32(0.01) * z3.astype(f32))
    return quad_state_from_raw(x, v, R, Om)
