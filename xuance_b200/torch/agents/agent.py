"""Agent base class - mirror of xuance/torch/agents/base/agent.py:23-370 (the parts on the hot path's callers).

Keeps the constructor contract (config Namespace + vector envs or explicit spaces + callback), the observation /
reward normalisation (agent.py:262-294), ``_build_representation`` / ``_build_learner`` through the registries,
``save_model`` / ``load_model`` (+ obs_rms.npy) and ``finish``.  TensorBoard / wandb logging is optional: scalars
are kept in ``self.logged`` and forwarded to a SummaryWriter only if ``config.logger == 'tensorboard'``."""
import os
import random
from abc import ABC, abstractmethod
from argparse import Namespace

import numpy as np
import torch
import torch.nn as nn

from ... import _lib
from ...common import BaseCallback, space2shape
from ...common.statistic_tools import RunningMeanStd
from ..rl_models import REGISTRY_Representation, ActivationFunctions
from ..utils import init_distributed_mode

EPS = 1e-8
InitializeFunctions = {"orthogonal": nn.init.orthogonal_, "xavier_uniform": nn.init.xavier_uniform_,
                       "kaiming_uniform": nn.init.kaiming_uniform_, None: None}
NormalizeFunctions = {"LayerNorm": nn.LayerNorm, "BatchNorm": nn.BatchNorm1d, "BatchNorm2d": nn.BatchNorm2d,
                      "GroupNorm": nn.GroupNorm, "InstanceNorm2d": nn.InstanceNorm2d}


def set_seed(seed):
    """xuance/torch/utils/operations.py set_seed: python, numpy, torch (+cuda)."""
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)
    np.random.seed(seed)
    random.seed(seed)


def set_device(device):
    if isinstance(device, int):
        return "cuda:%d" % device
    if device in ("gpu", "cuda", "GPU"):
        return "cuda:0"
    return device


class Agent(ABC):
    def __init__(self, config: Namespace, envs=None, observation_space=None, action_space=None, callback=None):
        set_seed(config.seed)
        self.config = config
        self.use_rnn = getattr(config, "use_rnn", False)
        self.use_actions_mask = getattr(config, "use_actions_mask", False)
        self.distributed_training = getattr(config, "distributed_training", False)
        if self.distributed_training:
            self.rank, self.world_size, local_rank = init_distributed_mode(getattr(config, "master_port", None))
            if torch.cuda.is_available():
                config.device = "cuda:%d" % local_rank
        else:
            self.world_size, self.rank = 1, 0
        self.gamma = config.gamma
        self.start_training = getattr(config, "start_training", 1)
        self.training_frequency = getattr(config, "training_frequency", 1)
        self.n_epochs = getattr(config, "n_epochs", 1)
        self.device = self.config.device = set_device(self.config.device)
        _lib.use_device(self.device)
        self.train_envs = envs
        self.render = getattr(config, "render", False)
        self.fps = getattr(config, "fps", 50)
        if envs is None:
            if observation_space is None or action_space is None:
                raise ValueError("Please provide the observation_space and action_space when the envs is not provided.")
            # config.parallels is the GLOBAL env count; under distributed training each rank owns an equal shard
            assert config.parallels % self.world_size == 0, "parallels must be divisible by the world size"
            self.n_envs = config.parallels // self.world_size
            self.observation_space, self.action_space = observation_space, action_space
            self.episode_length = config.episode_length = getattr(config, "episode_length", None)
        else:
            self.train_envs.reset()
            self.n_envs = self.train_envs.num_envs
            self.episode_length = self.config.episode_length = self.train_envs.max_episode_steps
            self.observation_space = self.train_envs.observation_space
            self.action_space = self.train_envs.action_space
        self.current_step = 0
        self.current_episode = np.zeros((self.n_envs,), np.int32)
        self.obs_rms = RunningMeanStd(shape=space2shape(self.observation_space))
        self.ret_rms = RunningMeanStd(shape=())
        self.returns = np.zeros((self.n_envs,), np.float32)
        self.use_obsnorm = getattr(config, "use_obsnorm", False)
        self.use_rewnorm = getattr(config, "use_rewnorm", False)
        self.obsnorm_range = getattr(config, "obsnorm_range", 5)
        self.rewnorm_range = getattr(config, "rewnorm_range", 5)
        self.normalize_fn = NormalizeFunctions[config.normalize] if hasattr(config, "normalize") else None
        self.initializer = InitializeFunctions[getattr(config, "initializer", "orthogonal")]
        self.activation = ActivationFunctions[config.activation]
        self.callback = callback or BaseCallback()
        self.logged = {}
        self.writer = None
        self.use_wandb = False
        if getattr(config, "logger", None) == "tensorboard" and self.rank == 0:
            try:
                from torch.utils.tensorboard import SummaryWriter
                log_dir = os.path.join(os.getcwd(), getattr(config, "log_dir", "logs"), f"seed_{config.seed}")
                os.makedirs(log_dir, exist_ok=True)
                self.writer = SummaryWriter(log_dir)
            except Exception:
                self.writer = None
        self.model_dir_save = os.path.join(os.getcwd(), getattr(config, "model_dir", "models"), f"seed_{config.seed}")
        self.model_dir_load = getattr(config, "model_dir", "models")
        self.model = None
        self.learner = None
        self.memory = None

    # ---------------------------------------------------------------- logging
    def log_infos(self, info: dict, x_index: int):
        for k, v in info.items():
            if v is None:
                continue
            self.logged[k] = v
            if self.writer is not None:
                try:
                    if isinstance(v, dict):
                        self.writer.add_scalars(k, v, x_index)
                    else:
                        self.writer.add_scalar(k, v, x_index)
                except Exception:
                    pass

    # ---------------------------------------------------------------- checkpoints (agent.py:199-229)
    def save_model(self, model_name, save_buffer=False):
        """agent.py:199-213.  ``save_buffer=True`` (extension) also writes this rank's replay /
        rollout buffer - HBM arrays, ring cursors, PER trees - to ``<model_name>.buffer.rank<r>``."""
        if save_buffer and self.memory is not None and hasattr(self.memory, "state_dict"):
            os.makedirs(self.model_dir_save, exist_ok=True)
            torch.save(self.memory.state_dict(),
                       os.path.join(self.model_dir_save, "%s.buffer.rank%d" % (model_name, self.rank)))
        if self.distributed_training and self.rank > 0:
            return
        os.makedirs(self.model_dir_save, exist_ok=True)
        self.learner.save_model(os.path.join(self.model_dir_save, model_name))
        if self.use_obsnorm:
            np.save(os.path.join(self.model_dir_save, "obs_rms.npy"),
                    {'count': self.obs_rms.count, 'mean': self.obs_rms.mean, 'var': self.obs_rms.var})

    def load_model(self, path, model=None, load_buffer=False):
        """agent.py:215-229.  ``load_buffer=True`` restores the buffer written by ``save_model(save_buffer=True)`` when
        the file for this rank sits next to the loaded ``.pth``."""
        path_loaded = self.learner.load_model(path, model)
        if self.use_obsnorm:
            p = os.path.join(os.path.dirname(str(path_loaded)), "obs_rms.npy")
            if os.path.exists(p):
                d = np.load(p, allow_pickle=True).item()
                self.obs_rms.count, self.obs_rms.mean, self.obs_rms.var = d['count'], d['mean'], d['var']
        if load_buffer:
            p = "%s.buffer.rank%d" % (str(path_loaded), self.rank)
            if not os.path.exists(p):
                raise FileNotFoundError("no buffer checkpoint %s" % p)
            self.memory.load_state_dict(torch.load(p, map_location="cpu", weights_only=False))

    # ---------------------------------------------------------------- normalisation (agent.py:262-294)
    def _process_observation(self, observations):
        if not self.use_obsnorm:
            return observations
        return np.clip((observations - self.obs_rms.mean) / (self.obs_rms.std + EPS),
                       -self.obsnorm_range, self.obsnorm_range)

    def _process_reward(self, rewards):
        if not self.use_rewnorm:
            return rewards
        std = np.clip(self.ret_rms.std, 0.1, 100)
        return np.clip(rewards / std, -self.rewnorm_range, self.rewnorm_range)

    def _build_representation(self, representation_key, input_space, config):
        if representation_key not in REGISTRY_Representation:
            raise AttributeError(f"{representation_key} is not registered in REGISTRY_Representation.")
        kw = dict(input_shape=space2shape(input_space),
                  hidden_sizes=getattr(config, "representation_hidden_size", None),
                  normalize=self.normalize_fn, initialize=nn.init.orthogonal_,
                  activation=ActivationFunctions[config.activation],
                  kernels=getattr(config, "kernels", None), strides=getattr(config, "strides", None),
                  filters=getattr(config, "filters", None), fc_hidden_sizes=getattr(config, "fc_hidden_sizes", None),
                  device=self.device)
        rep = REGISTRY_Representation[representation_key](**kw)
        if hasattr(rep, "set_compute"):
            rep.tc_planes = getattr(config, "tc_planes", 3)       # planes per float32 operand of compute="tc" (3 = exact to 2^-24)
            rep.set_compute(getattr(config, "compute", "fp32"))
        return rep

    @abstractmethod
    def _build_model(self, *args, **kwargs):
        raise NotImplementedError

    def _build_learner(self, *args, **kwargs):
        from ..learners import REGISTRY_Learners
        return REGISTRY_Learners[self.config.learner](*args)

    @abstractmethod
    def get_actions(self, *args, **kwargs):
        raise NotImplementedError

    @abstractmethod
    def train(self, *args, **kwargs):
        raise NotImplementedError

    @abstractmethod
    def test(self, *args, **kwargs):
        raise NotImplementedError

    def finish(self):
        if self.writer is not None:
            self.writer.close()
