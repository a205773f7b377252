"""Multi-GPU plumbing: env instances shard across ranks with NO collective on the step path
(every env is independent; RNG is keyed by the global env id, so results do not depend on the
number of GPUs).  Collectives (NCCL on GPUs, gloo in the CPU tests) are used only for the
per-chunk episode statistics and for assembling a global replay batch from per-rank
sub-batches.
"""
import torch
import torch.distributed as dist


def shard_env_range(total_envs, rank, world):
    """Contiguous env-id range [lo, hi) owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(int(total_envs), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_reduce_stats(stats):
    """Sum the [8] float64 rollout statistics vector (episodes, sum return, sum length,
    terminated, truncated, ...) over ranks."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        stats = stats.clone()
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def all_gather_replay_batch(sub_batch):
    """Concatenate per-rank replay sub-batches {field: [B/G, n, ...]} along the batch axis in
    rank order -> {field: [B, n, ...]} on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return sub_batch
    world = dist.get_world_size()
    out = {}
    for k in sorted(sub_batch):
        t = sub_batch[k].contiguous()
        full = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(full, t)
        out[k] = full
    return out


def exchange_batch_and_stats(sub_batch, stats):
    """One collective per chunk instead of eight: the per-rank replay sub-batch {field: [B/G, n, ...]} (float32) and
    the [8] float64 statistics vector are packed into one buffer, all-gathered once, and unpacked into the global
    batch {field: [B, n, ...]} (rank order, as `all_gather_replay_batch`) and the rank-summed statistics (as
    `all_reduce_stats`; summed in rank order, so the result is identical on every rank).  NVSwitch makes the cost of
    these small exchanges launch latency, not bandwidth, hence the single bucket."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return sub_batch, stats
    world = dist.get_world_size()
    keys = sorted(sub_batch)
    parts = [sub_batch[k].contiguous().float().reshape(-1) for k in keys]
    st32 = stats.detach().to(torch.float64).contiguous().view(torch.float32)         # 8 doubles -> 16 floats, bit-exact
    packed = torch.cat(parts + [st32.to(parts[0].device)])
    flat = torch.empty(world * packed.numel(), dtype=torch.float32, device=packed.device)
    dist.all_gather_into_tensor(flat, packed)
    full = flat.view(world, packed.numel())
    out, off = {}, 0
    for k, p in zip(keys, parts):
        shape = tuple(sub_batch[k].shape)
        out[k] = full[:, off:off + p.numel()].reshape((world * shape[0],) + shape[1:])
        off += p.numel()
    stats_all = full[:, off:off + st32.numel()].contiguous().view(torch.float64)      # [world, 8]
    return out, stats_all.sum(dim=0)


def broadcast_parameters(module_or_tensors, src=0):
    """Learner -> rollout ranks: broadcast a module's parameters (e.g. `networks.policy`, 285 KB) from rank `src`
    in ONE collective (flattened bucket) and copy them back in place, so that every rank samples with the same actor."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    tensors = [p.data for p in module_or_tensors.parameters()] if hasattr(module_or_tensors, "parameters") else list(module_or_tensors)
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.broadcast(flat, src=src)
    off = 0
    for t in tensors:
        t.copy_(flat[off:off + t.numel()].view_as(t))
        off += t.numel()


def mean_std_over_ranks(*per_instance):
    """Mean and population std of per-instance values whose instances are sharded over ranks (evaluator: one value per
    evaluation episode).  One all-reduce of the float64 moments [N, sum x_j, sum x_j^2 ...]; single-process: two-pass
    torch mean/std.  Returns [(mean, std), ...] as Python floats, identical on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [(float(x.mean()), float(x.std(unbiased=False))) for x in per_instance]
    dev = per_instance[0].device
    m = [torch.tensor(float(per_instance[0].numel()), dtype=torch.float64, device=dev)]
    for x in per_instance:
        x = x.double()
        m += [x.sum(), (x * x).sum()]
    m = torch.stack(m)
    dist.all_reduce(m, op=dist.ReduceOp.SUM)
    n = m[0]
    out = []
    for j in range(len(per_instance)):
        mean = m[1 + 2 * j] / n
        var = torch.clamp(m[2 + 2 * j] / n - mean * mean, min=0.0)
        out.append((float(mean), float(var.sqrt())))
    return out


def episode_summary(stats):
    """Host dict from a (reduced) statistics vector."""
    s = stats.detach().cpu().tolist()
    ep = max(s[0], 1.0)
    return {"episodes": int(s[0]), "mean_return": s[1] / ep, "mean_length": s[2] / ep, "terminated": int(s[3]),
            "truncated": int(s[4])}
