import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from oracle import actor as oactor, envs as oenv, rollout as oroll
from test_gpu_rollout import _mk, _sync_state
name='QuadTracking'; n,T,seed=1000,12,3
spec=oenv.SPECS[name]
ro,aw,w=_mk(name,n,1,n_step=4,seed=seed,max_step=7)
ro.state.reset()
object.__setattr__(spec,'max_step',7)
ids=np.arange(n,dtype=np.uint64)
venv=oroll.VectorEnv(name, oroll.philox_reset(name,seed,ids,np.zeros(n,np.int64)), seed=seed, env_ids=ids)
rng=np.random.default_rng(0)
_sync_state(ro,name,venv.state)
for t in range(T):
    eps=rng.standard_normal((n,spec.act_dim)).astype(np.float32)
    tr=oroll.sampler_step(venv,w,eps)
    ro.run(aw, eps=torch.as_tensor(eps[None]).cuda())
    g={k:v[ro.tr.H].cpu().numpy() for k,v in ro.tr.fields().items()}
    bad=(g['obs']!=tr['obs'])
    print(t,'mismatch',bad.sum(),'nan',np.isnan(g['obs']).sum(),np.isnan(tr['obs']).sum(), 'maxdiff',np.nanmax(np.abs(g['obs']-tr['obs'])), 'rows',np.nonzero(bad.any(1))[0][:5], 'done', tr['done'].sum())
    if bad.any():
        i=np.nonzero(bad.any(1))[0][0]; print(g['obs'][i], tr['obs'][i])
    _sync_state(ro,name,venv.state)
    ro.state.episode.copy_(torch.as_tensor(venv.episode.astype(np.int32)).cuda())
