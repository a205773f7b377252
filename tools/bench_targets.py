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


ms = timeit(lambda: tg.lyapunov_risk_raw(obs, obs2, lpn, lpo, v1, v2, coef, want_labels=False))
line("lyapunov_risk (fwd + analytic bwd)", ms, B * n * 4 * (2 * D + 4 + 2), B)
ms = timeit(lambda: tg.q_backup(rew, done, q1, q2, lpn, 0.99, 0.2))
line("q_backup", ms, B * n * 4 * 6, B)
ms = timeit(lambda: tg.stability_advantage(v1[:, 0].contiguous(), v2, coef))
line("stability_advantage + normalize", ms, B * 4 * (n + 1 + 3), B)
buf = B200NstepReplayBuffer(obs_dim=D, act_dim=A, buffer_max_size=B, n_step=n)
buf._ptr_size[1] = B
idx = torch.randint(0, B, (1 << 18,), device=dev, generator=g)
ms = timeit(lambda: buf.gather(idx))
line("ring_gather (2^18 windows)", ms, idx.numel() * n * 4 * (2 * D + A + 4) * 2, idx.numel())
