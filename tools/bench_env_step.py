"""HBM-roofline measurement of the unfused vector-env step (msacl_env_step: SyncVectorEnv.step drop-in) for the six
envs.  One JSON line per env: env-steps/s, algorithmic bytes per env-step, achieved GB/s and fraction of the HBM peak."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import msacl_b200  # noqa: F401
from msacl_b200.envs import B200VectorEnv
from msacl_b200.specs import SPECS

peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except Exception:
    pass
HBM = float(peaks.get("hbm_gbs", 6551.0))
for name, spec in SPECS.items():
    n = 1 << (21 if name == "QuadTracking" else 23)
    env = B200VectorEnv(name, n, env_seed=0, device="cuda")
    env.state.reset()
    lo = torch.as_tensor(spec.act_low, device="cuda", dtype=torch.float32)
    hi = torch.as_tensor(spec.act_high, device="cuda", dtype=torch.float32)
    act = lo + (hi - lo) * (0.45 + 0.1 * torch.rand(n, spec.act_dim, device="cuda"))
    for _ in range(3):
        env.step_device(act)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    torch.cuda.synchronize(); a.record()
    for _ in range(reps):
        env.step_device(act)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    # state (sf float32 rows, sd float64 rows, 4 int32 counters + ep_return) read + written; action in;
    # next_obs, final_obs, reward, terminated, truncated out
    bytes_per = 2 * (4 * spec.sf_rows + 8 * spec.sd_rows + 20) + 4 * (spec.act_dim + 2 * spec.obs_dim + 1) + 2
    gbs = bytes_per * n / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": f"env_step_kernel<{name}>", "envs": n, "ms": round(ms, 4), "env_steps_per_s": n / (ms * 1e-3),
                      "algorithmic_bytes_per_env_step": bytes_per, "achieved_gbs": round(gbs, 1), "hbm_peak_gbs": HBM,
                      "frac": round(gbs / HBM, 3)}), flush=True)
    del env
    torch.cuda.empty_cache()
