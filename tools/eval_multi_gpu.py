"""torchrun --nproc-per-node G tools/eval_multi_gpu.py : the evaluator sharded over G GPUs must return the statistics of
the same episode set as one GPU evaluating all of them (global instance ids key the resets)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import msacl_b200  # noqa: F401
from msacl_b200.evaluator import B200Evaluator
from msacl_b200.sampler import ActorWeights
from msacl_b200.specs import get_spec

rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl")
for env in ("Pendulum", "QuadTracking"):
    spec = get_spec(env)
    torch.manual_seed(0)
    lin = [torch.nn.Linear(spec.obs_dim, 256), torch.nn.Linear(256, 256), torch.nn.Linear(256, 2 * spec.act_dim)]
    aw = ActorWeights([(l.weight, l.bias) for l in lin])
    kw = dict(env_name=env, num_eval_episode=1001, reward_scale=100.0, cost_scale=100.0, max_step=120, device=f"cuda:{local}")
    sharded = B200Evaluator(**kw).run_parallel_episodes(aw)
    whole = B200Evaluator(distributed=False, **kw).run_parallel_episodes(aw)
    err = max(abs(a - b) / max(1.0, abs(b)) for a, b in zip(sharded, whole))
    print(f"rank {rank} {env}: sharded {sharded} whole {whole} max rel diff {err:.2e}", flush=True)
    assert err < 1e-9
dist.barrier()
dist.destroy_process_group()
