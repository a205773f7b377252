"""Bit-exact checks of the n-step window scatter + replay ring (integer bookkeeping and payload)
against the oracle's restatement of base.py:178-217 + nstep_replay_buffer.py, and against the
reference's own recorded ring (golden)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import actor as oactor
from oracle import envs as oenv
from oracle import rollout as oroll

pytestmark = pytest.mark.gpu
FIELDS = ("obs", "act", "rew", "cost", "obs2", "done", "logp")


@pytest.mark.parametrize("name,ring_size", [("VanderPol", 997), ("TwoLink", 100000), ("QuadTracking", 5003)])
def test_window_store_matches_oracle_ring(name, ring_size):
    from msacl_b200.buffer import B200NstepReplayBuffer
    from msacl_b200.sampler import ActorWeights, FusedRollout
    n, K, chunks, n_step = 300, 7, 6, 5
    spec = oenv.SPECS[name]
    ro = FusedRollout(name, n, K, n_step=n_step, seed=8, max_step=11)
    ro.state.reset()
    aw = ActorWeights(oactor.init_policy_weights(spec.obs_dim, spec.act_dim, seed=2))
    buf = B200NstepReplayBuffer(obs_dim=spec.obs_dim, act_dim=spec.act_dim, buffer_max_size=ring_size, n_step=n_step)
    emitter = oroll.WindowEmitter(n, n_step)
    ring = oroll.ReplayRing(ring_size, n_step, spec.obs_dim, spec.act_dim)
    total = 0
    for c in range(chunks):
        batch = ro.run(aw)
        cnt = buf.add_batch(batch)
        f = {k: v[ro.tr.H:].cpu().numpy() for k, v in ro.tr.fields().items()}
        host_windows = batch.materialize()
        wins_all = []
        for k in range(K):
            tr = {fld: f[fld][k] for fld in FIELDS}
            tr["done"] = tr["done"].astype(bool)
            emit, wins = emitter.push(tr)
            assert np.array_equal(emit, f["emit"][k].astype(bool))
            wins_all += wins
        ring.add_batch(wins_all)
        total += len(wins_all)
        assert int(cnt.item()) == len(wins_all) == len(host_windows)
        for hw, ow in zip(host_windows[:50], wins_all[:50]):
            assert np.array_equal(hw.n_step_obs, ow["obs"]) and np.array_equal(hw.n_step_log_prob, ow["logp"])
        assert buf.ptr == ring.ptr and buf.size == ring.size
    assert total > ring_size or ring_size > 5000     # the small rings must have wrapped
    for k in FIELDS:
        assert np.array_equal(buf.n_step_buf[k].cpu().numpy(), ring.buf[k]), k
    # gather = fancy indexing
    idx = np.random.default_rng(0).integers(0, ring.size, size=257)
    got = buf.gather(idx)
    want = ring.gather(idx)
    for k in FIELDS:
        assert np.array_equal(got[k].cpu().numpy(), want[k]), k
    b = buf.sample_batch(64)
    assert set(b) == set(FIELDS) and b["obs"].shape == (64, n_step, spec.obs_dim) and b["rew"].shape == (64, n_step)
    # sample_batch = library-side index draw over [0, size) + gather: the batch is the ring at the drawn slots
    big = buf.sample_batch(5000)
    idx = buf._idx[:5000].cpu().numpy()
    assert idx.min() >= 0 and idx.max() < ring.size and abs(idx.mean() - (ring.size - 1) / 2) < 4 * ring.size / np.sqrt(12 * 5000)
    want = ring.gather(idx)
    for k in FIELDS:
        assert np.array_equal(big[k].cpu().numpy(), want[k]), k


@pytest.mark.parametrize("chunks", [1, 2, None])
def test_ring_from_reference_golden_transitions(chunks):
    """Feed the reference sampler's recorded per-step transitions through the device scatter and
    compare the ring with the reference NstepReplayBuffer's arrays bit for bit.  chunks = number of K-slice chunks
    of the transition store (1: history copied before every launch; 2 / default 4: carried over only when the chunk
    index wraps -- the run is long enough to wrap several times)."""
    from msacl_b200 import _lib
    from msacl_b200.buffer import B200NstepReplayBuffer
    from msacl_b200.sampler import DeviceWindowBatch, TransitionBuffers
    from msacl_b200.specs import get_spec
    for name in oenv.ENV_NAMES:
        g = load_golden(f"sampler_{name}.npz")
        spec = get_spec(name)
        T, N = g["step_eps"].shape[:2]
        n_step = int(g["n_step"])
        buf = B200NstepReplayBuffer(obs_dim=spec.obs_dim, act_dim=spec.act_dim, buffer_max_size=int(g["ring"]), n_step=n_step)
        K = 4
        tr = TransitionBuffers(spec, N, K, n_step, torch.device("cuda"), chunks=chunks)
        assert tr.M == (chunks or 4) and T // K > 2 * tr.M
        for c in range(T // K):
            tr.roll_history()
            sl = slice(c * K, (c + 1) * K)
            put = lambda dst, a: dst[tr.H:].copy_(torch.as_tensor(np.nan_to_num(a[sl])).to(dst.dtype).cuda())
            put(tr.obs, g["step_obs"]); put(tr.act, g["step_act"]); put(tr.rew, g["step_rew"]); put(tr.cost, g["step_cost"])
            put(tr.obs2, g["step_obs2"]); put(tr.logp, g["step_logp"])
            tr.done[tr.H:].copy_(torch.as_tensor((g["step_done"][sl] > 0).astype(np.uint8)).cuda())
            tr.emit[tr.H:].copy_(torch.as_tensor(g["step_emit"][sl].astype(np.uint8)).cuda())
            buf.add_batch(DeviceWindowBatch(tr, n_step))
            assert buf.ptr == g["ptr_after"][(c + 1) * K - 1] and buf.size == g["size_after"][(c + 1) * K - 1]
        for k in FIELDS:
            assert np.array_equal(buf.n_step_buf[k].cpu().numpy(), g["ring_" + k]), (name, k)


def test_host_store_path_and_ram():
    from msacl_b200.buffer import B200NstepReplayBuffer
    buf = B200NstepReplayBuffer(obs_dim=2, act_dim=1, buffer_max_size=3, n_step=4)
    assert buf.__get_RAM__() == 0.0 and len(buf) == 0
    for i in range(5):
        buf.store(np.full((4, 2), i, np.float32), np.full((4, 1), i, np.float32), np.full(4, i), np.full(4, i),
                  np.full((4, 2), i, np.float32), np.zeros(4), np.full(4, -i))
    assert buf.ptr == 2 and buf.size == 3
    assert buf.n_step_buf["rew"][:, 0].cpu().tolist() == [3.0, 4.0, 2.0]
    assert buf.__get_RAM__() == round(3 * 4 * (2 + 1 + 1 + 1 + 2 + 1 + 1) * 4 / 2 ** 20, 2)


def test_chunked_transition_store_equals_rolled_store():
    """Same rollout, history copied before every launch (1 chunk) vs carried over only on wrap (3 chunks): the new
    transition slices, the stored windows and the ring bookkeeping must be bit-identical over several wraps."""
    from msacl_b200.buffer import B200NstepReplayBuffer
    from msacl_b200.sampler import ActorWeights, FusedRollout
    from msacl_b200.specs import get_spec
    name, n, K, n_step = "Pendulum", 300, 3, 5
    spec = get_spec(name)
    torch.manual_seed(3)
    lin = [torch.nn.Linear(spec.obs_dim, 256), torch.nn.Linear(256, 256), torch.nn.Linear(256, 2 * spec.act_dim)]
    aw = ActorWeights([(l.weight, l.bias) for l in lin])
    ros = [FusedRollout(name, n, K, n_step=n_step, seed=11, engine="tc", history_chunks=m) for m in (1, 3)]
    bufs = [B200NstepReplayBuffer(obs_dim=spec.obs_dim, act_dim=spec.act_dim, buffer_max_size=4096, n_step=n_step) for _ in ros]
    for ro in ros:
        ro.state.reset()
    assert [ro.tr.M for ro in ros] == [1, 3]
    for launch in range(8):
        batches = [ro.run(aw) for ro in ros]
        fa, fb = (ro.tr.fields() for ro in ros)
        for k in fa:
            assert torch.equal(fa[k], fb[k]), (launch, k)          # history + new slices of the current window
        cnt = [int(b.add_batch(x).item()) for b, x in zip(bufs, batches)]
        assert cnt[0] == cnt[1] and (bufs[0].ptr, bufs[0].size) == (bufs[1].ptr, bufs[1].size)
    assert bufs[0].size > 0
    for k in FIELDS:
        assert torch.equal(bufs[0].n_step_buf[k], bufs[1].n_step_buf[k]), k


@pytest.mark.parametrize("name,ring_size,chunks", [("Pendulum", 100000, 8), ("QuadTracking", 100000, 8), ("TwoLink", 1500, 5)])
def test_indexed_replay_equals_reference_layout_ring(name, ring_size, chunks):
    """Index-based store (window = position of its newest transition in the sampler's transition store) vs the
    reference-layout ring fed by the same launches: ptr / size / per-call counts identical, and every resident slot
    gathers bit-identical [n_step, .] payloads -- across chunk wraps of the store (history carry-over) and, for the small
    ring, across wraps of the position ring.  The reference-layout ring is itself bit-exact vs the oracle / the reference's
    recorded ring (tests above)."""
    from msacl_b200.buffer import B200IndexedReplayBuffer, B200NstepReplayBuffer
    from msacl_b200.sampler import ActorWeights, FusedRollout
    n, K, n_step = 300, 7, 5
    spec = oenv.SPECS[name]
    ro = FusedRollout(name, n, K, n_step=n_step, seed=8, max_step=11, engine="tc", history_chunks=chunks)
    ro.state.reset()
    aw = ActorWeights(oactor.init_policy_weights(spec.obs_dim, spec.act_dim, seed=2))
    kw = dict(obs_dim=spec.obs_dim, act_dim=spec.act_dim, buffer_max_size=ring_size, n_step=n_step)
    ref, ixb = B200NstepReplayBuffer(**kw), B200IndexedReplayBuffer(**kw)
    rng = np.random.default_rng(0)
    per_launch = []
    for launch in range(3 * chunks + 2):
        batch = ro.run(aw)
        c1 = int(ref.add_batch(batch).item())
        c2 = int(ixb.add_batch(batch).item())
        per_launch.append(c2)
        assert c1 == c2 and (ref.ptr, ref.size) == (ixb.ptr, ixb.size)
        resident = min(ixb.size, sum(per_launch[-(chunks - 2):]))
        assert int(ixb.valid_count().item()) == resident
        # the `resident` most recent slots, walking back from ptr - 1
        back = rng.integers(0, resident, size=257)
        slots = (ixb.ptr - 1 - back) % ring_size
        a, b = ref.gather(slots), ixb.gather(slots)
        for k in FIELDS:
            assert torch.equal(a[k], b[k]), (launch, k)
    assert ref.size > 0 and launch >= 2 * chunks
    s = ixb.sample_batch(64)
    assert set(s) == set(FIELDS) and s["obs"].shape == (64, n_step, spec.obs_dim) and s["done"].dtype == torch.float32
    # the one-launch sample_batch: the drawn slots lie among the `resident` most recent entries, the batch is exactly the
    # gather of those slots, the draws are uniform over them and differ from call to call
    big, slots = ixb.sample_batch(20000, return_slots=True)
    back = (ixb.ptr - 1 - slots.cpu().numpy()) % ring_size
    assert back.min() >= 0 and back.max() < resident
    want = ixb.gather(slots)
    for k in FIELDS:
        assert torch.equal(big[k], want[k]), k
    hist = np.bincount(back, minlength=resident)
    assert hist.min() > 0 or resident > 4000
    assert abs(back.mean() - (resident - 1) / 2) < 4 * resident / np.sqrt(12 * 20000)
    _, slots2 = ixb.sample_batch(20000, return_slots=True)
    assert not torch.equal(slots, slots2)
    with pytest.raises(NotImplementedError):
        ixb.store(None)


def test_indexed_replay_through_registries():
    """create_sampler / create_buffer with buffer_name='b200_indexed_replay_buffer': the sampler sizes its transition
    store from buffer_max_size, the buffer binds to it on the first add_batch."""
    import msacl_b200
    from msacl_b200.buffer import B200IndexedReplayBuffer
    kw = dict(env_name="DuctedFan", env_num=64, sample_batch_size=8, reward_scale=100.0, cost_scale=100.0, noise_params=None,
              n_step=4, obs_dim=6, act_dim=2, buffer_max_size=2000, buffer_name="b200_indexed_replay_buffer", env_seed=1)
    sampler = msacl_b200.create_sampler(**kw)
    buf = msacl_b200.create_buffer(**kw)
    assert isinstance(buf, B200IndexedReplayBuffer)
    assert sampler.rollout.tr.M == B200IndexedReplayBuffer.chunks_for(2000, 64, 8) == 6
    sampler.set_actor(__import__("msacl_b200.sampler", fromlist=["ActorWeights"]).ActorWeights(
        oactor.init_policy_weights(6, 2, seed=0)))
    for _ in range(9):
        data, _ = sampler.sample()
        buf.add_batch(data)
    assert 0 < buf.size <= 2000 and len(buf) == buf.size and buf.__get_RAM__() > 0
    out = buf.sample_batch(32)
    assert out["obs"].shape == (32, 4, 6) and torch.isfinite(out["rew"]).all()
