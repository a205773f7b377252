import sys
sys.path.insert(0, '/root/repo')
import torch
import msacl_b200
from msacl_b200.sampler import ActorWeights, FusedRollout
from msacl_b200.specs import get_spec
WRITE = not bool(int(__import__('os').environ.get('NOWRITE', '0')))
for env in sys.argv[1:]:
    spec = get_spec(env)
    torch.manual_seed(0)
    pol = torch.nn.Sequential(torch.nn.Linear(spec.obs_dim, 256), torch.nn.ReLU(), torch.nn.Linear(256, 256), torch.nn.ReLU(), torch.nn.Linear(256, 2 * spec.act_dim)).cuda()
    lin = [m for m in pol if isinstance(m, torch.nn.Linear)]
    aw = ActorWeights([(l.weight, l.bias) for l in lin])
    ro = FusedRollout(env, 1 << 18, 8, n_step=20, engine="tc")
    ro.state.reset()
    for _ in range(2): ro.run(aw, write=WRITE)
    ro.stats.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ro.run(aw, write=WRITE); b.record(); torch.cuda.synchronize()
    s = ro.stats.cpu().numpy()
    n = max(s[18], 1)
    print(env, 'ms %.2f' % a.elapsed_time(b), 'env %.0f wait %.0f |' % (s[5] / max(s[7], 1), s[6] / max(s[7], 1)),
          'per tile-step: epi1 waitH1 %.0f loop %.0f (waitAfree %.0f) | epi2 waitH2 %.0f compute %.0f | '
          'MMA waitX %.0f waitEpi1 %.0f waitA %.0f waitB %.0f waitTmemBuf %.0f'
          % (s[8] / n, s[9] / n, s[10] / n, s[11] / n, s[12] / n, s[13] / n, s[14] / n, s[15] / n, s[16] / n, s[17] / n))
