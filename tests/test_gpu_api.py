"""Drop-in behaviour of the Python shims (reference API surface) on the GPU."""
import numpy as np
import pytest
import torch

from oracle import envs as oenv
from oracle import rollout as oroll

pytestmark = pytest.mark.gpu


class _Policy(torch.nn.Module):
    """Stand-in with the attribute surface of the reference StochaPolicy (RL/apprfunc/mlp.py:111-136)."""

    def __init__(self, obs_dim, act_dim):
        super().__init__()
        self.policy = torch.nn.Sequential(torch.nn.Linear(obs_dim, 256), torch.nn.ReLU(), torch.nn.Linear(256, 256),
                                          torch.nn.ReLU(), torch.nn.Linear(256, 2 * act_dim), torch.nn.Identity())
        self.min_log_std, self.max_log_std = -20.0, 1.0


class _Networks(torch.nn.Module):
    def __init__(self, obs_dim, act_dim):
        super().__init__()
        self.policy = _Policy(obs_dim, act_dim)


def test_create_envs_surface_and_numpy_step():
    import msacl_b200
    envs = msacl_b200.create_envs(env_name="DuctedFan", env_num=33, env_seed=5)
    assert envs.single_observation_space.shape == (6,) and envs.single_action_space.shape == (2,)
    assert envs.action_space.low.shape == (33, 2) and envs.action_space.high.dtype == np.float32
    assert envs.num_envs == 33
    obs, info = envs.reset(seed=None)
    assert obs.shape == (33, 6) and obs.dtype == np.float32 and info == {}
    assert np.all(np.abs(obs) <= 0.5)
    act = np.random.default_rng(0).uniform(-5, 5, size=(33, 2)).astype(np.float32)
    nxt, rew, term, trunc, infos = envs.step(act)
    assert nxt.shape == (33, 6) and nxt.dtype == np.float32
    assert rew.shape == (33,) and rew.dtype == np.float64            # SyncVectorEnv returns a float64 reward array
    assert term.dtype == np.bool_ and trunc.dtype == np.bool_
    _, o_obs, o_rew, o_term, o_trunc = oenv.env_step("DuctedFan", {"obs": obs, "step": np.zeros(33, np.int32)}, act)
    np.testing.assert_allclose(nxt, o_obs, rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(rew, o_rew, rtol=1e-5, atol=1e-5)
    assert not term.any() and not trunc.any() and infos == {}


def test_vector_env_final_observation_on_done():
    from msacl_b200.envs import B200VectorEnv
    envs = B200VectorEnv("VanderPol", 16, env_seed=1, max_step=2)
    obs, _ = envs.reset()
    a = np.zeros((16, 1), np.float32)
    _, _, te, tr, infos = envs.step(a)
    assert not tr.any() and infos == {} or te.any()
    nxt, _, te, tr, infos = envs.step(a)
    assert tr.all()                                                     # time limit (max_step=2)
    fo = infos["final_observation"]
    assert fo.dtype == object and infos["_final_observation"].all()
    assert all(f.shape == (2,) and f.dtype == np.float32 for f in fo)
    assert np.all(np.abs(nxt) <= 5.0)                                   # fresh reset states (+-reset_noise)
    assert not np.allclose(np.stack(list(fo)), nxt)
    # torch-in / torch-out path stays on the device
    t_out = envs.step(torch.zeros(16, 1, device="cuda"))
    assert t_out[0].is_cuda and t_out[2].dtype == torch.bool


@pytest.mark.parametrize("name", ["Pendulum", "QuadTracking"])
def test_single_env_class_surface(name):
    from msacl_b200.envs import make_env
    env = make_env(name, 0, 0, False, "x")()
    spec = oenv.SPECS[name]
    assert env.observation_space.shape == (spec.obs_dim,) and env.action_space.shape == (spec.act_dim,)
    assert (env.obs_dim, env.act_dim, env.dt, env.max_step) == (spec.obs_dim, spec.act_dim, 0.01, 1000)
    assert env.control_step == spec.control_step
    obs, info = env.reset(seed=3)
    assert obs.shape == (spec.obs_dim,) and obs.dtype == np.float32 and info == {}
    a = ((spec.act_low + spec.act_high) / 2).astype(np.float32)
    o2, r, te, tr, info = env.step(a)
    assert o2.shape == (spec.obs_dim,) and isinstance(r, np.float32) and isinstance(te, bool) and tr is False and info == {}
    with pytest.raises(ValueError, match="Unknown custom env"):
        make_env("Acrobot")


def test_registry_errors():
    import msacl_b200
    with pytest.raises(KeyError):
        msacl_b200.create_sampler(sampler_name="on_sampler", env_name="VanderPol", env_num=4)
    with pytest.raises(KeyError):
        msacl_b200.create_buffer(buffer_name="prioritized_replay_buffer", obs_dim=2, act_dim=1, buffer_max_size=8, n_step=2)


@pytest.mark.parametrize("engine", ["tc", "ffma"])
def test_sampler_buffer_roundtrip_like_the_trainer(engine):
    """The call sequence of NstepOffSerialTrainer.step (nstep_off_serial_trainer.py:75-104)."""
    import msacl_b200
    kw = dict(env_name="TwoLink", env_num=512, env_seed=3, sample_batch_size=12, action_type="continu", reward_scale=100.0,
              cost_scale=100.0, noise_params=None, target_value=0.0, n_step=5, gamma=0.99, retrace_lambda=0.95,
              obs_dim=4, act_dim=2, buffer_max_size=4000, rollout_engine=engine)
    sampler = msacl_b200.create_sampler(**kw)
    buffer = msacl_b200.create_buffer(**kw)
    sampler.networks = _Networks(4, 2).cuda()          # the trainer overwrites .networks (…trainer.py:34)
    assert sampler.horizon == 12 and sampler.sample_batch_size == 12 * 512
    data, tb = sampler.sample()
    assert "Time/Sampler time [ms]-RL iter" in tb and tb["Time/Sampler time [ms]-RL iter"] > 0
    assert sampler.get_total_sample_number() == 12 * 512
    n_windows = len(data)
    tr = sampler.rollout.tr
    done = tr.done[tr.H:].cpu().numpy().astype(bool)
    # windows appear from the 5th transition of an episode on; recompute the count on the host
    run = np.zeros(512, np.int32); expect = 0
    for k in range(12):
        run = np.minimum(run + 1, 5); expect += int((run >= 5).sum()); run[done[k]] = 0
    assert n_windows == expect > 0
    first = data.materialize()[0]
    assert first.n_step_obs.shape == (5, 4) and first.n_step_act.shape == (5, 2) and first.n_step_done.dtype == np.float32
    buffer.add_batch(data)
    assert buffer.size == min(expect, 4000) and len(buffer) == buffer.size and buffer.ptr == expect % 4000
    assert buffer.__get_RAM__() > 0
    batch = buffer.sample_batch(64)
    assert set(batch) == {"obs", "act", "rew", "cost", "obs2", "done", "logp"}
    assert batch["obs"].shape == (64, 5, 4) and batch["logp"].shape == (64, 5) and batch["obs"].is_cuda
    # consecutive rows of a window are consecutive transitions of one env: obs[k+1] == obs2[k] unless done[k]
    o, o2, d = batch["obs"].cpu().numpy(), batch["obs2"].cpu().numpy(), batch["done"].cpu().numpy()
    cont = d[:, :-1] == 0
    assert np.array_equal(o[:, 1:][cont], o2[:, :-1][cont])
    assert (d[:, :-1] == 0).all()                      # a window never continues past a done (deque cleared)
    # new weights are picked up on the next call
    with torch.no_grad():
        for p in sampler.networks.parameters():
            p.zero_()
    data2, _ = sampler.sample()
    act = sampler.rollout.tr.act[sampler.rollout.tr.H:]
    logp = sampler.rollout.tr.logp[sampler.rollout.tr.H:]
    assert torch.isfinite(act).all() and torch.isfinite(logp).all()
    # zero weights -> mean 0, std 1: actions = 20*tanh(eps) -> roughly symmetric
    assert abs(float(act.mean())) < 0.5


def test_noise_params_rejected():
    import msacl_b200
    with pytest.raises(RuntimeError):
        msacl_b200.create_sampler(env_name="VanderPol", env_num=4, sample_batch_size=2, reward_scale=1.0, cost_scale=1.0,
                                  noise_params={"std": 0.1}, n_step=2)


@pytest.mark.parametrize("engine", ["tc", "ffma"])
def test_reference_default_config_sizes(engine):
    """BASELINE config 1 sizes (example/msacl_train.py defaults): env_num=4, sample_batch_size=20, n_step=20."""
    import msacl_b200
    kw = dict(env_name="VanderPol", env_num=4, env_seed=1, sample_batch_size=20, action_type="continu", reward_scale=100.0,
              cost_scale=100.0, noise_params=None, target_value=0.0, n_step=20, gamma=0.99, retrace_lambda=0.95,
              obs_dim=2, act_dim=1, buffer_max_size=1000, rollout_engine=engine)
    sampler = msacl_b200.create_sampler(**kw)
    buffer = msacl_b200.create_buffer(**kw)
    sampler.networks = _Networks(2, 1).cuda()
    total = 0
    for it in range(4):
        data, _ = sampler.sample()
        total += len(data)
        buffer.add_batch(data)
    assert sampler.get_total_sample_number() == 4 * 80
    assert buffer.size == total and 0 < total <= 4 * 80 - 4 * 19
    b = buffer.sample_batch(256)
    assert b["obs"].shape == (256, 20, 2) and torch.isfinite(b["rew"]).all()
