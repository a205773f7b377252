"""n-step off-policy serial trainer -- drop-in for `NstepOffSerialTrainer`
(RL/trainer/nstep_off_serial_trainer.py:21-163): same constructor signature and kwargs, same `step()` schedule
(sample every `sample_interval` iterations -> add_batch -> sample_batch -> model_update -> periodic log / checkpoint /
evaluation), same artefacts:

  * `<save_folder>/apprfunc/apprfunc_{iteration}.pkl` and `apprfunc_{iteration}_opt.pkl`: `torch.save(networks.state_dict())`
    with the reference's state-dict keys (:112-133, :157-163) -- files written here load into the reference's
    ApproxContainer and vice versa (tensors are saved from the device they live on; `torch.load(map_location=...)` as usual);
  * TensorBoard scalars under the reference's tags (RL/utils/tensorboard_setup.py:13-40), written with
    torch.utils.tensorboard when the `tensorboard` package is importable, otherwise into `<save_folder>/scalars.csv`.

What is NOT reproduced, on purpose: the per-iteration `ModuleOnDevice(networks, "cpu")` shuffle around `sampler.sample()`
and the evaluator (:78, :118) -- the fused sampler / evaluator read the policy weights where they live (on the GPU) --
and the `.cuda()` copy of the replay batch (:87-89), which is already on the device.
"""
import os
import time
from math import inf

import torch

# RL/utils/tensorboard_setup.py:13-40 -- a naming contract with the reference's plotting / CSV tools
tb_tags = {
    "TRM of RL iteration": "Evaluation/1-1. TRM-RL iter",
    "TRS of RL iteration": "Evaluation/1-1. TRS-RL iter",
    "TRM of total time": "Evaluation/2-1. TRM-Total time [s]",
    "TRM of collected samples": "Evaluation/3-1. TRM-Collected samples",
    "TRM of replay samples": "Evaluation/4-1. TRM-Replay samples",
    "TCM of RL iteration": "Evaluation/1-2. TCM-RL iter",
    "TCS of RL iteration": "Evaluation/1-2. TCS-RL iter",
    "TCM of total time": "Evaluation/2-2. TCM-Total time [s]",
    "TCM of collected samples": "Evaluation/3-2. TCM-Collected samples",
    "TCM of replay samples": "Evaluation/4-2. TCM-Replay samples",
    "Buffer RAM of RL iteration": "RAM/RAM [MB]-RL iter",
    "loss_actor": "Loss/Actor loss-RL iter",
    "loss_critic": "Loss/Critic loss-RL iter",
    "loss_entropy": "Loss/Entropy loss-RL iter",
    "loss_lyapunov": "Loss/Lyapunov loss-RL iter",
    "alg_time": "Time/Algorithm time [ms]-RL iter",
    "sampler_time": "Time/Sampler time [ms]-RL iter",
}


class _CsvWriter:
    """Fallback scalar sink with SummaryWriter's add_scalar / flush surface."""

    def __init__(self, log_dir):
        os.makedirs(log_dir, exist_ok=True)
        self.fh = open(os.path.join(log_dir, "scalars.csv"), "a")

    def add_scalar(self, tag, value, step):
        self.fh.write(f"{tag},{step},{float(value)}\n")

    def flush(self):
        self.fh.flush()

    def close(self):
        self.fh.close()


def _make_writer(log_dir):
    try:
        from torch.utils.tensorboard import SummaryWriter
        return SummaryWriter(log_dir=log_dir, flush_secs=20)
    except Exception:
        return _CsvWriter(log_dir)


def add_scalars(tb_info, writer, step):
    for key, value in tb_info.items():
        writer.add_scalar(key, value, step)


class _RunningAverage:
    """RL/utils/log_data.py:5-35 (LogData): running mean per key, popped at every log interval."""

    def __init__(self):
        self.data, self.counter = {}, {}

    def add_average(self, d):
        for k, v in d.items():
            if k not in self.data:
                self.data[k], self.counter[k] = v, 1
            else:
                self.data[k] = (self.data[k] * self.counter[k] + v) / (self.counter[k] + 1)
                self.counter[k] += 1

    def pop(self):
        out = self.data.copy()
        self.data, self.counter = {}, {}
        return out


class B200NstepOffSerialTrainer:
    def __init__(self, alg, sampler, buffer, evaluator, **kwargs):
        self.alg, self.sampler, self.buffer, self.evaluator = alg, sampler, buffer, evaluator
        if kwargs.get("buffer_name") == "prioritized_replay_buffer":
            raise RuntimeError("prioritized replay is not part of the MSACL path (reference default: nstep_replay_buffer)")
        self.networks = self.alg.networks
        self.sampler.networks = self.networks
        if self.evaluator is not None:
            self.evaluator.networks = self.networks
        if kwargs.get("ini_network_dir") is not None:
            self.networks.load_state_dict(torch.load(kwargs["ini_network_dir"], map_location=next(self.networks.parameters()).device))
        self.replay_batch_size = kwargs["replay_batch_size"]
        self.max_iteration = kwargs["max_iteration"]
        self.policy_frequency = kwargs["policy_frequency"]
        self.sample_interval = kwargs.get("sample_interval", 1)
        self.log_save_interval = kwargs["log_save_interval"]
        self.apprfunc_save_interval = kwargs["apprfunc_save_interval"]
        self.save_folder = kwargs["save_folder"]
        self.eval_interval = kwargs["eval_interval"]
        self.verbose = bool(kwargs.get("verbose", True))
        self.best_tar = -inf
        self.iteration = 0
        os.makedirs(os.path.join(self.save_folder, "apprfunc"), exist_ok=True)
        self.writer = _make_writer(self.save_folder)
        add_scalars({tb_tags["alg_time"]: 0, tb_tags["sampler_time"]: 0}, self.writer, 0)
        self.writer.flush()
        while self.buffer.size < kwargs["buffer_warm_size"]:          # :62-64 pre-sampling
            samples, _ = self.sampler.sample()
            self.buffer.add_batch(samples)
        self.sampler_tb_dict = _RunningAverage()
        self.start_time = time.time()

    def _say(self, *a):
        if self.verbose:
            print(*a)

    def step(self):
        if self.iteration % self.sample_interval == 0:
            sampler_samples, sampler_tb_dict = self.sampler.sample()
            self.buffer.add_batch(sampler_samples)
            self.sampler_tb_dict.add_average(sampler_tb_dict)
        replay_samples = self.buffer.sample_batch(self.replay_batch_size)
        self.networks.train()
        if self.iteration % self.policy_frequency == 0:
            alg_tb_dict = self.alg.model_update(replay_samples, self.iteration)
            if self.iteration % self.log_save_interval == 0:
                self._say("Iter = ", self.iteration, "save training data!")
                add_scalars(alg_tb_dict, self.writer, step=self.iteration)
        else:
            self.alg.model_update(replay_samples, self.iteration)
        self.networks.eval()
        if self.iteration % self.log_save_interval == 0:
            self._say("Iter = ", self.iteration, "save average sampling time!")
            add_scalars(self.sampler_tb_dict.pop(), self.writer, step=self.iteration)
        if self.iteration % self.apprfunc_save_interval == 0:
            self.save_apprfunc()
        if self.evaluator is not None and self.iteration % self.eval_interval == 0 and self.iteration > 0:
            trm, trs, tcm, tcs = self.evaluator.run_evaluation(self.iteration)
            if trm >= self.best_tar and self.iteration >= self.max_iteration / 5:
                self.best_tar = trm
                self._say("Eval_Iter: {}, Highest total average return = {}! Current total average cost = {}".format(
                    str(self.iteration), str(self.best_tar), str(tcm)))
                folder = os.path.join(self.save_folder, "apprfunc")
                for filename in os.listdir(folder):
                    if filename.endswith("_opt.pkl"):
                        os.remove(os.path.join(folder, filename))
                torch.save(self.networks.state_dict(), os.path.join(folder, "apprfunc_{}_opt.pkl".format(self.iteration)))
            elapsed = int(time.time() - self.start_time)
            w = self.writer
            w.add_scalar(tb_tags["Buffer RAM of RL iteration"], self.buffer.__get_RAM__(), self.iteration)
            w.add_scalar(tb_tags["TRM of RL iteration"], trm, self.iteration)
            w.add_scalar(tb_tags["TRS of RL iteration"], trs, self.iteration)
            w.add_scalar(tb_tags["TRM of total time"], trm, elapsed)
            w.add_scalar(tb_tags["TCM of RL iteration"], tcm, self.iteration)
            w.add_scalar(tb_tags["TCS of RL iteration"], tcs, self.iteration)
            w.add_scalar(tb_tags["TCM of total time"], tcm, elapsed)

    def train(self):
        while self.iteration <= self.max_iteration:
            self.step()
            self.iteration += 1
        self.save_apprfunc()
        self.writer.flush()

    def save_apprfunc(self):
        torch.save(self.networks.state_dict(), os.path.join(self.save_folder, "apprfunc", "apprfunc_{}.pkl".format(self.iteration)))


NstepOffSerialTrainer = B200NstepOffSerialTrainer
