"""Learner base class - mirror of xuance/torch/learners/base/drl_learner.py:12-214.

Same constructor contract (config fields read the same way), ``update`` abstract, ``save_model`` /
``load_model`` writing/reading the reference's checkpoint dict ({'policy', 'optimizer', 'rng_state',
'cuda_rng_state'}, drl_learner.py:64-93).  Distributed training is the torchrun process group; there are no
per-module DDP wrappers, the learner all-reduces its flat gradient bucket once per update."""
import os
from abc import ABC, abstractmethod
from pathlib import Path

import torch
import torch.distributed as dist


class Learner(ABC):
    def __init__(self, config, model, callback):
        self.value_normalizer = None
        self.config = config
        self.distributed_training = bool(getattr(config, "distributed_training", False))
        self.episode_length = getattr(config, "episode_length", None)
        self.learning_rate = getattr(config, "learning_rate", None)
        self.use_linear_lr_decay = getattr(config, "use_linear_lr_decay", False)
        self.end_factor_lr_decay = getattr(config, "end_factor_lr_decay", 1.0)
        self.gamma = getattr(config, "gamma", 0.99)
        self.use_cnn = getattr(config, "use_cnn", False)
        self.use_rnn = getattr(config, "use_rnn", False)
        self.use_actions_mask = getattr(config, "use_actions_mask", False)
        self.model = model
        self.optimizer = None
        self.scheduler = None
        self.callback = callback
        if self.distributed_training and dist.is_available() and dist.is_initialized():
            self.world_size = dist.get_world_size()
            self.rank = dist.get_rank()
        elif self.distributed_training:
            self.world_size = int(os.environ.get("WORLD_SIZE", "1"))
            self.rank = int(os.environ.get("RANK", "0"))
        else:
            self.world_size, self.rank = 1, 0
        self.use_grad_clip = config.use_grad_clip
        self.grad_clip_norm = config.grad_clip_norm
        self.device = torch.device(config.device if not isinstance(config.device, int) else "cuda:%d" % config.device)
        if self.device.type != "cuda":
            raise RuntimeError("xuance_b200 learners run on CUDA devices only (config.device=%r); there is no "
                               "CPU fallback" % (config.device,))
        self.model_dir = getattr(config, "model_dir", "./models")
        self.total_iters = self.estimate_total_iterations()
        self.iterations = 0

    def estimate_total_iterations(self):
        """drl_learner.py:57-62."""
        start_training = getattr(self.config, "start_training", 0)
        training_frequency = getattr(self.config, "training_frequency", 1)
        return (self.config.running_steps - start_training) // (training_frequency * self.config.parallels)

    # ---------------------------------------------------------------- checkpoints (reference format)
    def _opt_state(self):
        if isinstance(self.optimizer, dict):
            return {k: v.state_dict() for k, v in self.optimizer.items()}
        if isinstance(self.optimizer, list):
            return [o.state_dict() for o in self.optimizer]
        return self.optimizer.state_dict()

    def save_model(self, model_path):
        os.makedirs(os.path.dirname(os.path.abspath(model_path)), exist_ok=True)
        torch.save({'policy': self.model.state_dict(), 'optimizer': self._opt_state(),
                    'rng_state': torch.get_rng_state(),
                    'cuda_rng_state': torch.cuda.get_rng_state_all()}, model_path)

    def load_model(self, path, model=None):
        """drl_learner.py:95-157: a file, or a directory holding seed_* folders -> newest -> final_train_model.pth
        (or any .pth).  Restores policy, optimizer and RNG state."""
        target = os.path.join(path, model) if model is not None else path
        if os.path.isfile(target):
            model_path = target
        else:
            if not os.path.isdir(path):
                raise RuntimeError(f"The path '{path}' is not a valid directory or file!")
            folders = sorted(f for f in os.listdir(path) if "seed_" in f)
            if not folders:
                raise RuntimeError(f"No model files with 'seed_' found in '{path}'!")
            d = Path(os.path.join(path, folders[-1]))
            names = list(d.glob("*.pth"))
            if not names:
                raise FileNotFoundError(f"No .pth file found in {d}")
            finals = [f for f in names if "final_train_model.pth" in str(f)]
            model_path = str(finals[0] if finals else sorted(names)[-1])
        ckpt = torch.load(str(model_path), map_location=self.device, weights_only=True)
        self.model.load_state_dict(ckpt['policy'])
        opt = ckpt.get('optimizer')
        if opt is not None:
            if isinstance(self.optimizer, dict):
                for k, v in self.optimizer.items():
                    if k in opt:
                        v.load_state_dict(opt[k])
            elif isinstance(self.optimizer, list):
                for o, s in zip(self.optimizer, opt):
                    o.load_state_dict(s)
            elif self.optimizer is not None:
                self.optimizer.load_state_dict(opt)
        if 'rng_state' in ckpt:
            torch.set_rng_state(ckpt['rng_state'].cpu())
        print(f"Successfully load model from '{model_path}'.")
        return model_path

    @abstractmethod
    def update(self, *args):
        raise NotImplementedError
