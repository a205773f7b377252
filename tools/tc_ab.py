"""Same-box A/B of rollout_tc_kernel build / run-time variants: python tools/tc_ab.py VAR=v1,v2,... [env:n:K ...]
(each variant = one value of the environment variable VAR read by the library at launch time)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import msacl_b200  # noqa: F401
from msacl_b200.sampler import ActorWeights, FusedRollout
from msacl_b200.specs import get_spec
var, vals = sys.argv[1].split("=")
vals = vals.split(",")
cases = sys.argv[2:] or ["Pendulum:65536:256", "TwoLink:1048576:16", "SingleTrackCar:2097152:16", "QuadTracking:2097152:16"]
for case in cases:
    env, n, K = case.split(":"); n, K = int(n), int(K)
    spec = get_spec(env)
    torch.manual_seed(0)
    lin = [torch.nn.Linear(spec.obs_dim, 256), torch.nn.Linear(256, 256), torch.nn.Linear(256, 2 * spec.act_dim)]
    aw = ActorWeights([(l.weight, l.bias) for l in lin])
    ro = FusedRollout(env, n, K, n_step=20, engine="tc")
    ro.state.reset()
    res = {}
    for rnd in range(2):                      # two interleaved rounds: box drift shows up as a difference between them
        for v in vals:
            os.environ[var] = v
            for _ in range(2): ro.run(aw)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); a.record()
            for _ in range(5): ro.run(aw)
            b.record(); torch.cuda.synchronize()
            res.setdefault(v, []).append(a.elapsed_time(b) / 5)
    print(env, n, K, " | ".join(f"{var}={v}: " + " / ".join(f"{ms:.3f} ms ({n * K / ms / 1e6:.3f}e9/s)" for ms in res[v]) for v in vals), flush=True)
    del ro
    torch.cuda.empty_cache()
