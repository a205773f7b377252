"""HBM-roofline measurement of the MSACL target / replay kernels at BASELINE config 3 size
(TwoLink, B = 2^20 windows, n = 20).  Prints one JSON line per kernel."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import msacl_b200  # noqa: F401
from msacl_b200 import targets as tg
from msacl_b200.buffer import B200NstepReplayBuffer

peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except Exception:
    pass
HBM = float(peaks.get("hbm_gbs", 6650.0))
B, n, D, A = 1 << 20, 20, 4, 2
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
r = lambda *s: torch.randn(*s, device=dev, generator=g)
obs, obs2 = r(B, n, D) * 0.5, r(B, n, D) * 0.5
lpn, lpo = r(B, n), r(B, n)
v1, v2 = r(B, n).abs(), r(B, n).abs()
rew, done, q1, q2 = r(B, n), (torch.rand(B, n, device=dev, generator=g) < 0.1).float(), r(B, n), r(B, n)
coef = tg.Coefficients(n)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def line(name, ms, nbytes, units):
    gbs = nbytes / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": name, "ms": round(ms, 4), "windows_per_s": units / (ms * 1e-3), "algorithmic_bytes": nbytes,
                      "achieved_gbs": round(gbs, 1), "hbm_peak_gbs": HBM, "frac": round(gbs / HBM, 3)}))


# Each kernel is timed twice: "op" = the Python-level operator a user calls (allocations, the small torch epilogue ops
# and ctypes overhead included), "kernel" = the C-ABI call alone on preallocated outputs (what the HBM roofline bounds).
import ctypes as C
from msacl_b200 import _lib
lib = _lib.load()
st = _lib.current_stream()

ms = timeit(lambda: tg.lyapunov_risk_raw(obs, obs2, lpn, lpo, v1, v2, coef, want_labels=False))
line("lyapunov_risk (fwd + analytic bwd), op", ms, B * n * 4 * (2 * D + 4 + 2), B)
parts = torch.empty(3, dtype=torch.float64, device=dev)
g1, g2 = torch.empty_like(v1), torch.empty_like(v2)
ms = timeit(lambda: _lib.check(lib.msacl_lyapunov_risk(
    B, n, D, obs.data_ptr(), obs2.data_ptr(), lpn.data_ptr(), lpo.data_ptr(), v1.data_ptr(), v2.data_ptr(),
    coef.son.data_ptr(), coef.diff.data_ptr(), coef.sl.data_ptr(), coef.alpha1, coef.alpha2, 10.0, 1.0,
    parts.data_ptr(), g1.data_ptr(), g2.data_ptr(), None, None, st)))
line("lyapunov_risk (fwd + analytic bwd), kernel", ms, B * n * 4 * (2 * D + 4 + 2), B)

ms = timeit(lambda: tg.q_backup(rew, done, q1, q2, lpn, 0.99, 0.2))
line("q_backup, op", ms, B * n * 4 * 6, B)
bk = torch.empty_like(rew)
ms = timeit(lambda: _lib.check(lib.msacl_q_backup(B * n, rew.data_ptr(), done.data_ptr(), q1.data_ptr(), q2.data_ptr(), lpn.data_ptr(),
                                                  0.99, 0.2, bk.data_ptr(), st)))
line("q_backup, kernel", ms, B * n * 4 * 6, B)

v0 = v1[:, 0].contiguous()
ms = timeit(lambda: tg.stability_advantage(v0, v2, coef))
line("stability_advantage + normalize, op", ms, B * 4 * (n + 1 + 3), B)
raw, adv, mom = torch.empty_like(v0), torch.empty_like(v0), torch.empty(2, dtype=torch.float64, device=dev)


def _adv():
    _lib.check(lib.msacl_stability_advantage(B, n, v0.data_ptr(), v2.data_ptr(), coef.diff.data_ptr(), coef.sl.data_ptr(),
                                             raw.data_ptr(), mom.data_ptr(), st))
    _lib.check(lib.msacl_advantage_normalize(B, raw.data_ptr(), mom.data_ptr(), adv.data_ptr(), st))


ms = timeit(_adv)
line("stability_advantage + normalize, kernels", ms, B * 4 * (n + 1 + 3), B)

buf = B200NstepReplayBuffer(obs_dim=D, act_dim=A, buffer_max_size=B, n_step=n)
buf._ptr_size[1] = B
idx = torch.randint(0, B, (1 << 18,), device=dev, generator=g)
ms = timeit(lambda: buf.gather(idx))
line("ring_gather (2^18 windows), op", ms, idx.numel() * n * 4 * (2 * D + A + 4) * 2, idx.numel())
out = {k: torch.empty(idx.numel(), *v.shape[1:], dtype=torch.float32, device=dev) for k, v in buf.n_step_buf.items()}
dst = buf._make_ring(out, idx.numel())
ms = timeit(lambda: _lib.check(lib.msacl_ring_gather(C.byref(buf._ring), idx.data_ptr(), idx.numel(), C.byref(dst), st)))
line("ring_gather (2^18 windows), kernel", ms, idx.numel() * n * 4 * (2 * D + A + 4) * 2, idx.numel())

# n-step window store (deque windows + add_batch): every env emits a window at every step of a K-step chunk; the ring
# keeps the last `max_size` of them.  Algorithmic bytes = live windows x (read the n_step rows + write the ring entry).
from msacl_b200.sampler import DeviceWindowBatch, TransitionBuffers
from msacl_b200.specs import get_spec
for env_name in ("TwoLink", "QuadTracking"):
    spec = get_spec(env_name)
    N, K = 1 << 18, 8
    Dw, Aw = spec.obs_dim, spec.act_dim
    tr = TransitionBuffers(spec, N, K, n, torch.device(dev), chunks=1)
    tr.roll_history()
    for k, v in tr.fields().items():
        if v.dtype == torch.float32:
            v.copy_(torch.randn(v.shape, device=dev, generator=g))
    tr.emit.fill_(1)
    ring_size = 1 << 20
    buf2 = B200NstepReplayBuffer(obs_dim=Dw, act_dim=Aw, buffer_max_size=ring_size, n_step=n)
    batch = DeviceWindowBatch(tr, n)
    ms = timeit(lambda: buf2.add_batch(batch))
    live = min(N * K, ring_size)
    line(f"window_store {env_name} (count + scan + scatter, 2^20 live windows)", ms, live * n * 4 * (2 * Dw + Aw + 4) * 2 + N * K, live)
    del tr, buf2, batch
    torch.cuda.empty_cache()
