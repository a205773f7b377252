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

# round 2: general rollout engine, indexed replay, learner (tcgen05 GEMMs incl. the streamed-weights instantiation with
# 256-bit loads / stores, target kernels with the all-lanes-live Lyapunov risk at n = 20 / 16 / 7)
from msacl_b200.buffer import B200IndexedReplayBuffer
from msacl_b200.sampler import GeneralActor
spec = get_spec("TwoLink")
lin = [torch.nn.Linear(spec.obs_dim, 48), torch.nn.Linear(48, 300), torch.nn.Linear(300, 2 * spec.act_dim)]
ga = GeneralActor([(l.weight, l.bias) for l in lin], [torch.nn.Tanh(), torch.nn.Tanh(), torch.nn.Identity()])
ro = FusedRollout("TwoLink", 333, 6, n_step=4, max_step=5, history_chunks=4)
ro.state.reset()
ibuf = B200IndexedReplayBuffer(obs_dim=spec.obs_dim, act_dim=spec.act_dim, buffer_max_size=1500, n_step=4)
for _ in range(5):
    ibuf.add_batch(ro.run(ga))
ibuf.sample_batch(19)
for n in (20, 16, 7):
    B, D = 37, 4
    coef = tg.Coefficients(n)
    o = torch.randn(B, n, D, device="cuda")
    lp = torch.randn(B, n, device="cuda")
    vv = o.square().sum(-1)
    tg.lyapunov_risk_raw(o, o * 0.9, lp, lp - 0.1, vv, vv * 0.8, coef)
alg = msacl_b200.create_alg(algorithm="msacl", env_name="TwoLink", obs_dim=spec.obs_dim, act_dim=spec.act_dim, n_step=5,
                            action_low_limit=spec.act_low, action_high_limit=spec.act_high, q_learning_rate=1e-3,
                            lyapunov_learning_rate=1e-3, policy_learning_rate=3e-4, alpha_learning_rate=1e-3, learner_graph=False)
for B in (40, 4000):          # 4000 x 5 rows: >= 148 row tiles -> pre-packed weights, streamed instantiation
    r = lambda *s: torch.randn(*s, device="cuda")
    data = dict(obs=r(B, 5, spec.obs_dim) * 0.3, obs2=r(B, 5, spec.obs_dim) * 0.3, act=torch.zeros(B, 5, spec.act_dim, device="cuda"),
                rew=-torch.rand(B, 5, device="cuda"), cost=torch.rand(B, 5, device="cuda"), done=torch.zeros(B, 5, device="cuda"),
                logp=r(B, 5) - 1.0)
    alg.model_update(data, 2)
torch.cuda.synchronize()
print("round-2 kernels ok")
