"""Small end-to-end exercise of every kernel for `compute-sanitizer --tool memcheck`."""
import sys
sys.path.insert(0, '/root/repo')
import torch
import msacl_b200
from msacl_b200 import targets as tg
from msacl_b200.buffer import B200NstepReplayBuffer
from msacl_b200.envs import B200VectorEnv
from msacl_b200.sampler import ActorWeights, FusedRollout
from msacl_b200.specs import get_spec
engines = sys.argv[1:] or ["ffma", "tc"]
for name in ("VanderPol", "SingleTrackCar", "QuadTracking"):
    spec = get_spec(name)
    torch.manual_seed(0)
    lin = [torch.nn.Linear(spec.obs_dim, 256), torch.nn.Linear(256, 256), torch.nn.Linear(256, 2 * spec.act_dim)]
    aw = ActorWeights([(l.weight, l.bias) for l in lin])
    v = B200VectorEnv(name, 77, max_step=3)
    v.reset()
    for _ in range(4):
        v.step_device(torch.zeros(77, spec.act_dim, device="cuda"))
    for eng in engines:
        ro = FusedRollout(name, 333, 5, n_step=3, engine=eng, max_step=4)
        ro.state.reset()
        buf = B200NstepReplayBuffer(obs_dim=spec.obs_dim, act_dim=spec.act_dim, buffer_max_size=500, n_step=3)
        for _ in range(2):
            buf.add_batch(ro.run(aw))
        b = buf.sample_batch(17)
        coef = tg.Coefficients(3)
        vv = b["rew"].abs()
        tg.lyapunov_risk_raw(b["obs"], b["obs2"], b["logp"], b["logp"], vv, vv, coef)
        tg.q_backup(b["rew"], b["done"], vv, vv, b["logp"], 0.99, 0.2)
        tg.stability_advantage(vv[:, 0].contiguous(), vv, coef)
    torch.cuda.synchronize()
    print(name, "ok")
