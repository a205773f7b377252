"""A few unfused vector-env steps per env (for ncu):  python tools/prof_env_step.py TwoLink QuadTracking"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import msacl_b200  # noqa: F401
from msacl_b200.envs import B200VectorEnv
from msacl_b200.specs import SPECS

for name in sys.argv[1:]:
    spec = SPECS[name]
    n = 1 << (21 if name == "QuadTracking" else 23)
    env = B200VectorEnv(name, n, env_seed=0, device="cuda")
    env.state.reset()
    lo = torch.as_tensor(spec.act_low, device="cuda", dtype=torch.float32)
    hi = torch.as_tensor(spec.act_high, device="cuda", dtype=torch.float32)
    act = lo + (hi - lo) * (0.45 + 0.1 * torch.rand(n, spec.act_dim, device="cuda"))
    for _ in range(3):
        env.step_device(act)
    torch.cuda.synchronize()
    print(name, "ok")
    del env
    torch.cuda.empty_cache()
