"""Greedy policy evaluation on the fused rollout kernel -- drop-in for `Evaluator.run_parallel_episodes`
(RL/trainer/evaluator.py:141-204): `num_eval_episode` env instances are reset and stepped with the
distribution's mode() action (act_distribution_cls.py:90-95) until every instance has finished its FIRST
episode; per instance the scaled reward / cost are averaged over the steps of that episode, and the mean
and (population) std over instances are returned: (TRM, TRS, TCM, TCS).
With torch.distributed initialised (one process per GPU) the `num_eval_episode` instances are sharded over the ranks
(global instance ids key the resets, so the set of episodes does not depend on the number of GPUs) and the statistics
are combined with one all-reduce of the float64 moments; every rank returns the same four numbers.
"""
import torch
import torch.distributed as dist

from . import distributed as mdist
from .sampler import ActorWeights, FusedRollout, actor_from_policy


class B200Evaluator:
    def __init__(self, **kwargs):
        self.env_id = kwargs["env_name"]
        self.num_eval_episode = int(kwargs["num_eval_episode"])
        self.reward_scale, self.cost_scale = kwargs["reward_scale"], kwargs["cost_scale"]
        self.device = torch.device(kwargs.get("device", "cuda"))
        self.chunk = int(kwargs.get("eval_chunk_steps", 50))
        self.engine = kwargs.get("rollout_engine", "tc")
        self.max_step = kwargs.get("max_step")
        self.networks = kwargs.get("networks")
        self.seed = int(kwargs.get("eval_env_seed") or 0)
        self.distributed = bool(kwargs.get("distributed", True))     # shard the episodes over torch.distributed ranks
        self._resets = 0

    def load_state_dict(self, state_dict):
        self.networks.load_state_dict(state_dict)

    def run_parallel_episodes(self, actor: ActorWeights = None, state_init=None):
        actor = actor or actor_from_policy(self.networks.policy, device=self.device)
        lo, hi = 0, self.num_eval_episode
        sharded = self.distributed and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        if sharded:
            lo, hi = mdist.shard_env_range(self.num_eval_episode, dist.get_rank(), dist.get_world_size())
        n = max(hi - lo, 1)       # (a rank without instances still takes part in the all-reduce with one dummy env)
        ro = FusedRollout(self.env_id, n, self.chunk, n_step=1, reward_scale=self.reward_scale, cost_scale=self.cost_scale,
                          seed=self.seed, env_base=lo, device=self.device, max_step=self.max_step, engine=self.engine)
        ro.state.episode.fill_(self._resets)          # envs.reset(seed=None): a fresh set of initial states per call
        self._resets += 1
        ro.state.reset()
        if state_init is not None:
            state_init(ro.state)
        finished = torch.zeros(n, dtype=torch.bool, device=self.device)
        ret_sum = torch.zeros(n, dtype=torch.float64, device=self.device)
        cost_sum, count = torch.zeros_like(ret_sum), torch.zeros_like(ret_sum)
        limit = (self.max_step or ro.spec.max_step) + self.chunk
        steps = 0
        while not bool(finished.all()) and steps < limit:
            ro.run(actor, deterministic=True)
            tr = ro.tr
            done = tr.done[tr.H:].bool()
            cum = torch.cumsum(done.to(torch.int32), dim=0)
            mask = (~finished)[None, :] & ((cum - done.to(torch.int32)) == 0)     # steps up to and incl. the first done
            ret_sum += (tr.rew[tr.H:].double() * mask).sum(0)
            cost_sum += (tr.cost[tr.H:].double() * mask).sum(0)
            count += mask.sum(0)
            finished |= cum[-1] > 0
            steps += self.chunk
        ep_ret, ep_cost = ret_sum / count, cost_sum / count
        self.last_first_episode_len = count.to(torch.int32)      # steps of each instance's first episode (this rank's shard)
        if hi - lo == 0:                                     # dummy instance of an empty shard: contributes nothing
            ep_ret, ep_cost = ep_ret[:0], ep_cost[:0]
        if not sharded:
            return (float(ep_ret.mean()), float(ep_ret.std(unbiased=False)), float(ep_cost.mean()), float(ep_cost.std(unbiased=False)))
        (trm, trs), (tcm, tcs) = mdist.mean_std_over_ranks(ep_ret, ep_cost)
        return (trm, trs, tcm, tcs)

    def run_evaluation(self, iteration=0):
        return self.run_parallel_episodes()
