"""world_size-2 gloo tests of the multi-GPU host logic (no GPU needed)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import msacl_b200  # noqa: F401
from msacl_b200 import distributed as mdist


def test_shard_ranges_cover_and_balance():
    for total, world in [(16, 2), (17, 4), (1 << 24, 8), (5, 8)]:
        spans = [mdist.shard_env_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    stats = torch.tensor([1.0 + rank, 10.0 * (rank + 1), 5.0, float(rank), 1.0, 0, 0, 0], dtype=torch.float64)
    red = mdist.all_reduce_stats(stats)
    sub = {"obs": torch.full((3, 4, 2), float(rank)), "rew": torch.full((3, 4), float(rank) + 0.5)}
    full = mdist.all_gather_replay_batch(sub)
    # the single-bucket exchange must give exactly what the two separate collectives give
    full2, red2 = mdist.exchange_batch_and_stats(sub, stats)
    assert sorted(full2) == sorted(full) and all(torch.equal(full2[k], full[k]) for k in full)
    assert torch.equal(red2, red) and red2.dtype == torch.float64
    # learner -> rollout ranks weight broadcast (one bucket)
    torch.manual_seed(100 + rank)
    pol = torch.nn.Sequential(torch.nn.Linear(3, 8), torch.nn.ReLU(), torch.nn.Linear(8, 2))
    torch.manual_seed(100)
    want = torch.nn.Sequential(torch.nn.Linear(3, 8), torch.nn.ReLU(), torch.nn.Linear(8, 2))
    mdist.broadcast_parameters(pol, src=0)
    assert all(torch.equal(a, b) for a, b in zip(pol.parameters(), want.parameters()))
    # evaluator statistics: instances sharded over ranks, moments combined with one all-reduce
    allv = torch.arange(10, dtype=torch.float64) ** 1.5
    lo, hi = mdist.shard_env_range(10, rank, world)
    (m, sd), (m2, sd2) = mdist.mean_std_over_ranks(allv[lo:hi], -2.0 * allv[lo:hi])
    assert abs(m - float(allv.mean())) < 1e-12 and abs(sd - float(allv.std(unbiased=False))) < 1e-12
    assert abs(m2 + 2.0 * float(allv.mean())) < 1e-12 and abs(sd2 - 2.0 * float(allv.std(unbiased=False))) < 1e-12
    # overlapped exchange protocol (packed buffer, results consumed one call later); synchronous on CPU / gloo
    ex = mdist.BatchExchange({"obs": (2,), "rew": ()}, 3, 4, "cpu")
    got_prev = []
    for step in range(3):
        views = ex.views()
        views["obs"].fill_(float(10 * step + rank)); views["rew"].fill_(float(10 * step + rank) + 0.5)   # "gather kernel" output
        batch, st = ex.exchange(views, stats * (step + 1))
        got_prev.append((batch["obs"][:, 0, 0].tolist(), batch["rew"][:, 0].tolist(), st.tolist()))
        host, st_h = ex.host_views(ex.recv[0 if step == 0 else 1 - ((step) & 1)].clone())
        assert torch.equal(host["obs"], batch["obs"]) and torch.equal(st_h, st)
    # call 0 returns its own result, call t > 0 the result of call t - 1 (one step stale)
    for step, src in enumerate((0, 0, 1)):
        assert got_prev[step][0] == [10.0 * src] * 3 + [10.0 * src + 1] * 3
        assert got_prev[step][1] == [10.0 * src + 0.5] * 3 + [10.0 * src + 1.5] * 3
        assert got_prev[step][2] == (red * (src + 1)).tolist()
    q.put((rank, red.tolist(), {k: v.tolist() for k, v in full.items()}, mdist.episode_summary(red)))
    dist.destroy_process_group()


def test_gloo_world2_stats_and_replay_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(60) for p in procs]
    for rank, red, full, summ in res:
        assert red[:5] == [3.0, 30.0, 10.0, 1.0, 2.0]
        assert torch.tensor(full["obs"]).shape == (6, 4, 2)
        assert torch.tensor(full["obs"])[:3].eq(0).all() and torch.tensor(full["obs"])[3:].eq(1).all()
        assert torch.tensor(full["rew"])[:, 0].tolist() == [0.5] * 3 + [1.5] * 3
        assert summ["episodes"] == 3 and summ["mean_return"] == 10.0
