import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import msacl_b200
from msacl_b200.sampler import ActorWeights, FusedRollout
from oracle import actor as oactor, envs as oenv
name = sys.argv[1]; n = int(sys.argv[2]); K = int(sys.argv[3])
spec = oenv.SPECS[name]
aw = ActorWeights(oactor.init_policy_weights(spec.obs_dim, spec.act_dim, seed=1))
a = FusedRollout(name, n, K, n_step=4, seed=3, engine="ffma"); b = FusedRollout(name, n, K, n_step=4, seed=3, engine="tc")
a.state.reset(); b.state.reset()
a.run(aw); b.run(aw); torch.cuda.synchronize()
d = (a.tr.act[a.tr.H] - b.tr.act[b.tr.H]).abs().max().item()
print(name, n, K, 'ok, max act diff step0', d)
