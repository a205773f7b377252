"""Trainer glue + checkpoint parity on the GPU:
 (a) the reference's own NstepOffSerialTrainer, built through the reference's registries in a scratch copy of the unmodified
     reference package, drives the B200 sampler / buffer / learner for 50 iterations and writes apprfunc_{iter}.pkl, which
     loads into the reference's own MSACL ApproxContainer (tools/reference_dropin.py);
 (b) msacl_b200's trainer (same schedule without the per-iteration GPU<->CPU module shuffle) trains, checkpoints, logs the
     reference's TensorBoard tags and resumes from its checkpoint."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_reference_trainer_drives_b200_modules_and_checkpoint_round_trip():
    import reference_dropin as rd
    if rd.reference_package() is None:
        pytest.skip("baseline/_ref/RL not present on this box")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "reference_dropin.py"), "train", "50"], capture_output=True,
                         text=True, timeout=900)
    assert res.returncode == 0, (res.stdout[-1500:], res.stderr[-3000:])
    out = json.loads(res.stdout.strip().splitlines()[-1])
    assert out["trainer_class"] == "RL.trainer.nstep_off_serial_trainer.NstepOffSerialTrainer"       # the reference's trainer
    assert (out["sampler_class"], out["buffer_class"], out["alg_class"]) == ("B200NstepOffSampler", "B200NstepReplayBuffer", "B200MSACL")
    assert out["learner_engine"] == "fused" and out["iterations"] == 51
    assert "apprfunc_0.pkl" in out["checkpoints"] and "apprfunc_25.pkl" in out["checkpoints"] and "apprfunc_51.pkl" in out["checkpoints"]
    assert out["checkpoint_loads_into_reference_MSACL"] and out["all_finite"] and out["params_moved_max"] > 1e-4
    assert out["buffer_size"] >= 5000 and out["total_sample_number"] >= 64 * 20 * 51 and out["tensorboard_files"]
    print("reference trainer loop:", out["wall_ms_per_iteration"], "ms / iteration")


def test_b200_trainer_trains_checkpoints_logs_and_resumes(tmp_path):
    import msacl_b200
    from msacl_b200.specs import get_spec
    from msacl_b200.trainer import tb_tags
    spec = get_spec("Pendulum")
    save = str(tmp_path / "run")
    kw = dict(algorithm="msacl", env_name="Pendulum", obs_dim=spec.obs_dim, act_dim=spec.act_dim, n_step=20,
              action_low_limit=spec.act_low, action_high_limit=spec.act_high, q_learning_rate=1e-3, lyapunov_learning_rate=1e-3,
              policy_learning_rate=3e-4, alpha_learning_rate=1e-3, lya_diff_scale=10.0, env_num=64, env_seed=1, sample_batch_size=20,
              action_type="continu", reward_scale=100.0, cost_scale=100.0, noise_params=None, target_value=0.0, buffer_max_size=100000,
              buffer_warm_size=3000, replay_batch_size=128, max_iteration=20, policy_frequency=2, log_save_interval=4,
              apprfunc_save_interval=10, eval_interval=10, save_folder=save, num_eval_episode=8, eval_env_seed=5, ini_network_dir=None,
              buffer_name="nstep_replay_buffer", trainer="nstep_off_serial_trainer", sampler_name="nstep_off_sampler", verbose=False)
    alg = msacl_b200.create_alg(**kw)
    trainer = msacl_b200.create_trainer(alg, msacl_b200.create_sampler(**kw), msacl_b200.create_buffer(**kw),
                                        msacl_b200.create_evaluator(**kw), **kw)
    assert trainer.buffer.size >= 3000
    trainer.train()
    files = sorted(os.listdir(os.path.join(save, "apprfunc")))
    assert {"apprfunc_0.pkl", "apprfunc_10.pkl", "apprfunc_20.pkl", "apprfunc_21.pkl"} <= set(files)
    assert any(f.endswith("_opt.pkl") for f in files)                       # best model kept after the evaluations at 10 and 20
    sd = torch.load(os.path.join(save, "apprfunc", "apprfunc_21.pkl"), map_location="cpu")
    assert set(sd) == set(alg.networks.state_dict()) and "policy.policy.0.weight" in sd and "lyapunov.lya.4.bias" in sd
    for k, v in alg.networks.state_dict().items():
        assert torch.equal(v.cpu(), sd[k])
    trainer.writer.flush()
    # scalars under the reference's tags
    try:
        from tensorboard.backend.event_processing import event_accumulator
        ea = event_accumulator.EventAccumulator(save)
        ea.Reload()
        keys = set(ea.scalars.Keys())
    except ImportError:
        keys = {line.split(",")[0] for line in open(os.path.join(save, "scalars.csv"))}
    for t in ("alg_time", "sampler_time", "loss_critic", "loss_lyapunov", "loss_actor", "TRM of RL iteration", "TCM of RL iteration",
              "Buffer RAM of RL iteration"):
        assert tb_tags[t] in keys, t
    # resume: a new run initialised from the checkpoint starts from identical parameters
    kw2 = dict(kw, ini_network_dir=os.path.join(save, "apprfunc", "apprfunc_21.pkl"), save_folder=str(tmp_path / "run2"), max_iteration=2)
    alg2 = msacl_b200.create_alg(**kw2)
    tr2 = msacl_b200.create_trainer(alg2, msacl_b200.create_sampler(**kw2), msacl_b200.create_buffer(**kw2), None, **kw2)
    for k, v in alg2.networks.state_dict().items():
        assert torch.equal(v.cpu(), sd[k]), k
    tr2.train()
    assert all(np.isfinite(v.float().cpu().numpy()).all() for v in alg2.networks.state_dict().values())
