"""Thin runner layer (reference: xuance/engine/__init__.py:33-131 ``get_runner`` and engine/run_drl.py RunnerDRL):
config -> vector envs -> agent -> run('train' | 'test' | 'benchmark').  Orchestration only - no compute."""
import time
from copy import deepcopy

import numpy as np

from .common.common_tools import get_arguments
from .environment import make_envs
from .torch.agents import REGISTRY_Agents


class RunnerDRL:
    def __init__(self, config):
        self.config = config
        self.envs = make_envs(config)
        self.agent = REGISTRY_Agents[config.agent](config, self.envs)
        self.n_envs = self.envs.num_envs
        self.rank = getattr(self.agent, "rank", 0)

    def run(self, mode="train", **kwargs):
        if mode == "train":
            steps = kwargs.get("running_steps", self.config.running_steps)
            info = self.agent.train(max(1, steps // self.n_envs))
            self.agent.save_model("final_train_model.pth")
            return info
        if mode == "test":
            return self.agent.test(test_episodes=kwargs.get("test_episodes", self.config.test_episode))
        if mode == "benchmark":
            steps = kwargs.get("running_steps", self.config.running_steps)
            interval = kwargs.get("eval_interval", self.config.eval_interval)
            episodes = kwargs.get("test_episodes", self.config.test_episode)
            cfg_test = deepcopy(self.config)
            cfg_test.parallels = 1
            cfg_test.distributed_training = False
            curve, t0 = [], time.time()
            for epoch in range(max(1, steps // interval)):
                self.agent.train(max(1, interval // self.n_envs))
                scores = self.agent.test(test_episodes=episodes, test_envs=make_envs(cfg_test), close_envs=True)
                curve.append((self.agent.current_step, float(np.mean(scores)), float(np.std(scores))))
            return {"learning_curve": curve, "seconds": time.time() - t0}
        raise ValueError(f"unknown mode {mode}")

    def finish(self):
        self.envs.close()
        self.agent.finish()


def get_runner(algo, env, env_id=None, config_path=None, parser_args=None):
    config = get_arguments(algo, env, env_id, config_path, parser_args)
    if getattr(config, "dl_toolbox", "torch") != "torch":
        raise ValueError("xuance_b200 is torch-only (the tensorflow / mindspore backends are removed)")
    return RunnerDRL(config)
