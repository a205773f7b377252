"""Prove the drop-in INSIDE the real reference tree.

A scratch copy of the UNMODIFIED reference package `RL` (baseline/_ref/RL, installed by __graft_entry__.build();
/root/reference/RL in the build container) gets the one-line modules INTEGRATION.md lists; then the reference's own
registries resolve them and -- on a GPU -- the reference's own `NstepOffSerialTrainer` drives the B200 sampler / buffer /
learner for a few dozen iterations, writes `apprfunc_{iter}.pkl`, and the checkpoint is loaded back into a fresh
reference-built algorithm.

    python tools/reference_dropin.py registries          # CPU: registry resolution only
    python tools/reference_dropin.py train [iterations]   # GPU: reference trainer loop + checkpoint round trip

Prints one JSON line.  Nothing of the reference is copied into this repository; the scratch tree lives under /tmp.
"""
import json
import os
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SHIMS = {
    "RL/trainer/sampler/b200_nstep_off_sampler.py": "from msacl_b200.sampler import B200NstepOffSampler  # noqa: F401\n",
    "RL/trainer/buffer/b200_nstep_replay_buffer.py": "from msacl_b200.buffer import B200NstepReplayBuffer  # noqa: F401\n",
    "RL/trainer/buffer/b200_indexed_replay_buffer.py": "from msacl_b200.buffer import B200IndexedReplayBuffer  # noqa: F401\n",
    "RL/algorithm/msacl_b200.py": "from msacl_b200.algorithm import B200MSACL as MSACL_B200, ApproxContainer  # noqa: F401\n",
    "RL/trainer/b200_nstep_off_serial_trainer.py": "from msacl_b200.trainer import B200NstepOffSerialTrainer  # noqa: F401\n",
}


def reference_package():
    for cand in (os.path.join(ROOT, "baseline", "_ref", "RL"), "/root/reference/RL"):
        if os.path.isdir(cand):
            return cand
    return None


def build_tree():
    src = reference_package()
    if src is None:
        raise SystemExit("no reference package available (baseline/_ref/RL or /root/reference/RL)")
    tmp = tempfile.mkdtemp(prefix="msacl_dropin_")
    shutil.copytree(src, os.path.join(tmp, "RL"))
    for rel, text in SHIMS.items():
        with open(os.path.join(tmp, rel), "w") as fh:
            fh.write(text)
    return tmp, src


def setup_imports(tmp):
    import numpy as np
    if not hasattr(np, "float_"):
        np.float_ = np.float64          # RL/utils/common_utils.py:50 uses the removed NumPy 1.x name
    for p in (tmp, os.path.join(ROOT, "tests", "golden", "_gym_stub"), ROOT):
        sys.path.insert(0, p)
    import msacl_b200  # noqa: F401   (the alias module at the repo root)


def reference_args(env_name, save_folder, **over):
    """Keyword set of example/msacl_train.py:20-174 (reference defaults) + the registry ids of the B200 modules."""
    args = dict(
        env_name=env_name, algorithm="msacl_b200", enable_cuda=True, seed=3, env_num=64, env_seed=1, capture_video=False,
        target_value=0.0, reward_scale=100.0, cost_scale=100.0, is_render=False, is_parallel_eval=True,
        value_func_name="ActionValue", value_func_type="MLP", value_hidden_sizes=[256, 256], value_hidden_activation="relu",
        value_output_activation="linear", lyapunov_func_name="LyapunovValue", lyapunov_func_type="MLP",
        lyapunov_hidden_sizes=[256, 256], lyapunov_hidden_activation="tanh", lyapunov_output_dim=256,
        lyapunov_output_activation="linear", lyapunov_single_input_dim=False, policy_func_name="StochaPolicy",
        policy_func_type="MLP", policy_act_distribution="TanhGaussDistribution", policy_hidden_sizes=[256, 256],
        policy_hidden_activation="relu", policy_min_log_std=-20, policy_max_log_std=1, q_learning_rate=1e-3,
        lyapunov_learning_rate=1e-3, policy_learning_rate=3e-4, alpha_learning_rate=1e-3, lya_diff_scale=10.0, lya_zero_scale=1.0,
        lya_positive_scale=1.0, gamma=0.99, retrace_lambda=0.95, tau=0.005, disable_auto_alpha=False, alpha=1.0,
        set_alpha_bound=False, alpha_bound=2.0, n_step=20, policy_frequency=2, target_network_frequency=1, anneal_lr=False,
        alpha1=1, alpha2=2, lya_eta=0.15, clip_coef=0.1, trainer="nstep_off_serial_trainer", max_iteration=50, ini_network_dir=None,
        sampler_name="b200_nstep_off_sampler", sample_batch_size=20, noise_params=None, buffer_name="b200_nstep_replay_buffer",
        buffer_warm_size=5000, buffer_max_size=200000, replay_batch_size=256, sample_interval=1, evaluator_name="evaluator",
        num_eval_episode=2, eval_interval=25, eval_save=False, eval_env_seed=12345, save_folder=save_folder,
        apprfunc_save_interval=25, log_save_interval=10)
    args.update(over)
    return args


def cmd_registries():
    tmp, src = build_tree()
    try:
        setup_imports(tmp)
        import io
        from contextlib import redirect_stdout
        with redirect_stdout(io.StringIO()):
            from RL.create_pkg import create_alg, create_buffer, create_sampler, create_trainer
        import msacl_b200.algorithm as A
        import msacl_b200.buffer as Bf
        import msacl_b200.sampler as S
        import msacl_b200.trainer as T
        out = {
            "reference_package": src,
            "sampler": create_sampler.registry["b200_nstep_off_sampler"].entry_point is S.B200NstepOffSampler,
            "buffer": create_buffer.registry["b200_nstep_replay_buffer"].entry_point is Bf.B200NstepReplayBuffer,
            "buffer_indexed": create_buffer.registry["b200_indexed_replay_buffer"].entry_point is Bf.B200IndexedReplayBuffer,
            "algorithm": create_alg.registry["msacl_b200"].entry_point is A.B200MSACL,
            "approx_container": create_alg.registry["msacl_b200"].approx_container_cls is A.ApproxContainer,
            "trainer": create_trainer.registry["b200_nstep_off_serial_trainer"].entry_point is T.B200NstepOffSerialTrainer,
            "reference_ids_still_registered": all(k in create_sampler.registry for k in ("nstep_off_sampler",)) and
                                              "nstep_replay_buffer" in create_buffer.registry and "msacl" in create_alg.registry and
                                              "nstep_off_serial_trainer" in create_trainer.registry,
        }
        # unknown ids fail exactly as in the reference (KeyError from the registries)
        try:
            create_sampler.create_sampler(sampler_name="no_such_sampler")
            out["unknown_id_raises"] = False
        except KeyError:
            out["unknown_id_raises"] = True
        print(json.dumps(out))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def cmd_train(iterations):
    tmp, src = build_tree()
    save = tempfile.mkdtemp(prefix="msacl_dropin_run_")
    try:
        setup_imports(tmp)
        import io
        import time
        from contextlib import redirect_stdout
        import torch
        log = io.StringIO()
        with redirect_stdout(log):
            from RL.create_pkg.create_alg import create_alg
            from RL.create_pkg.create_buffer import create_buffer
            from RL.create_pkg.create_envs import create_envs
            from RL.create_pkg.create_evaluator import create_evaluator
            from RL.create_pkg.create_sampler import create_sampler
            from RL.create_pkg.create_trainer import create_trainer
            from RL.utils.init_args import init_args
            args = reference_args("VanderPol", save, max_iteration=iterations)
            envs = create_envs(**args)                     # the reference's own (CPU, gymnasium) envs: only probed for the spaces
            args = init_args(envs, **args)                 # init_args.py:16-74: use_gpu, obs_dim, act limits, save folder, seed
            alg = create_alg(**args)                       # -> msacl_b200.algorithm.B200MSACL through the reference registry
            sampler = create_sampler(**args)               # -> B200NstepOffSampler
            buffer = create_buffer(**args)                 # -> B200NstepReplayBuffer
            evaluator = create_evaluator(**args)           # the reference's own Evaluator (CPU envs), policy = alg.networks
            trainer = create_trainer(alg, sampler, buffer, evaluator, **args)      # the reference's NstepOffSerialTrainer
            before = {k: v.detach().clone() for k, v in alg.networks.state_dict().items()}
            t0 = time.perf_counter()
            while trainer.iteration <= iterations:         # NstepOffSerialTrainer.train (:150-155) without the final browser step
                trainer.step()
                trainer.iteration += 1
            torch.cuda.synchronize()
            wall = time.perf_counter() - t0
            trainer.save_apprfunc()
            trainer.writer.flush()
        files = sorted(os.listdir(os.path.join(save, "apprfunc")))
        after = alg.networks.state_dict()
        moved = {k: float((after[k].float().cpu() - before[k].float().cpu()).abs().max()) for k in after}
        # checkpoint round trip through the REFERENCE's own algorithm class ("msacl"): same state-dict keys
        last = os.path.join(save, "apprfunc", "apprfunc_{}.pkl".format(trainer.iteration))
        sd = torch.load(last, map_location="cpu")
        with redirect_stdout(log):
            ref_args = dict(args, algorithm="msacl", use_gpu=False, enable_cuda=False)
            ref_alg = create_alg(**ref_args)               # RL/algorithm/msacl.py MSACL + its ApproxContainer
        missing, unexpected = ref_alg.networks.load_state_dict(sd, strict=True), None
        same = all(torch.equal(ref_alg.networks.state_dict()[k].cpu(), sd[k].cpu()) for k in sd)
        tb = [f for f in os.listdir(save) if f.startswith("events.out.tfevents") or f == "scalars.csv"]
        print(json.dumps({
            "reference_package": src, "trainer_class": type(trainer).__module__ + "." + type(trainer).__name__,
            "sampler_class": type(sampler).__name__, "buffer_class": type(buffer).__name__, "alg_class": type(alg).__name__,
            "learner_engine": getattr(alg, "engine_name", None), "iterations": trainer.iteration, "wall_ms_per_iteration": wall * 1e3 / trainer.iteration,
            "buffer_size": int(buffer.size), "checkpoints": files, "tensorboard_files": tb, "params_moved_max": max(moved.values()),
            "all_finite": all(bool(torch.isfinite(v.float()).all()) for v in after.values()),
            "checkpoint_loads_into_reference_MSACL": bool(same), "state_keys": len(sd),
            "total_sample_number": sampler.get_total_sample_number()}))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
        shutil.rmtree(save, ignore_errors=True)


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "registries"
    if mode == "registries":
        cmd_registries()
    else:
        cmd_train(int(sys.argv[2]) if len(sys.argv) > 2 else 50)
