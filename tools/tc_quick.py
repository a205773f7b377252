import sys
sys.path.insert(0, '/root/repo')
import torch
import msacl_b200
from msacl_b200.sampler import ActorWeights, FusedRollout
from msacl_b200.specs import get_spec
name = sys.argv[1]; n = int(sys.argv[2]); K = int(sys.argv[3])
spec = get_spec(name)
torch.manual_seed(1)
lin = [torch.nn.Linear(spec.obs_dim, 256), torch.nn.Linear(256, 256), torch.nn.Linear(256, 2 * spec.act_dim)]
aw = ActorWeights([(l.weight, l.bias) for l in lin])
a = FusedRollout(name, n, K, n_step=4, seed=3, engine="ffma"); b = FusedRollout(name, n, K, n_step=4, seed=3, engine="tc")
a.state.reset(); b.state.reset()
a.run(aw); b.run(aw); torch.cuda.synchronize()
d = (a.tr.act[a.tr.H] - b.tr.act[b.tr.H]).abs().max().item()
print(name, n, K, 'ok, max act diff step0', d)
