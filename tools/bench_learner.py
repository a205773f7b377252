"""model_update wall time per iteration, fused learner (precision 6 / 3) vs the autograd + cuBLAS engine."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import msacl_b200
from msacl_b200.specs import get_spec

env = sys.argv[1] if len(sys.argv) > 1 else "TwoLink"
spec = get_spec(env)
D, A, n = spec.obs_dim, spec.act_dim, 20
dev = torch.device("cuda")
for B in (256, 4096, (1 << 20) // 20):
    for engine, prec in (("torch", 0), ("fused", 6), ("fused", 3)):
        alg = msacl_b200.create_alg(algorithm="msacl", env_name=env, obs_dim=D, act_dim=A, n_step=n, action_low_limit=spec.act_low,
                                    action_high_limit=spec.act_high, q_learning_rate=1e-3, lyapunov_learning_rate=1e-3,
                                    policy_learning_rate=3e-4, alpha_learning_rate=1e-3, lya_diff_scale=10.0, learner_engine=engine,
                                    learner_precision=prec or 6)
        g = torch.Generator(device=dev).manual_seed(1)
        r = lambda *s: torch.randn(*s, device=dev, generator=g)
        lo, hi = (torch.as_tensor(x, device=dev) for x in (spec.act_low, spec.act_high))
        data = dict(obs=r(B, n, D) * 0.4, act=(lo + (hi - lo) * torch.rand(B, n, A, device=dev, generator=g)) * 0.98,
                    rew=-torch.rand(B, n, device=dev, generator=g) * 50, cost=torch.rand(B, n, device=dev, generator=g),
                    done=(torch.rand(B, n, device=dev, generator=g) < 0.1).float(), logp=r(B, n) - 1.0)
        data["obs2"] = data["obs"] + 0.05 * r(B, n, D)
        for it in range(2, 8):
            alg.model_update(data, it)
        iters = 20 if B <= 4096 else 6
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for it in range(8, 8 + iters):
            alg.model_update(data, it)
        b.record(); torch.cuda.synchronize()
        print(json.dumps({"env": env, "replay_batch": B, "rows": B * n, "engine": engine, "precision": prec or None,
                          "ms_per_iter": round(a.elapsed_time(b) / iters, 4)}), flush=True)
        del alg, data
        torch.cuda.empty_cache()
