"""DQN / PER-DQN agents - mirrors of xuance/torch/agents/qlearning_family/dqn_agent.py:14-52 and
perdqn_agent.py:17-109 (beta annealing, priorities fed back after every update)."""
import torch

from ...common import PerOffPolicyBuffer
from ..rl_models import DeepQNetwork, DuelingDeepQNetwork
from .off_policy import OffPolicyAgent


class DQN_Agent(OffPolicyAgent):
    def __init__(self, config, envs=None, observation_space=None, action_space=None, callback=None):
        super().__init__(config, envs, observation_space, action_space, callback)
        self.model = self._build_model()
        self.memory = self._build_memory()
        self.learner = self._build_learner(self.config, self.model, self.callback)

    def _build_model(self):
        rep = self._build_representation(self.config.representation, self.observation_space, self.config)
        return DeepQNetwork(representation=rep, hidden_size=self.config.q_hidden_size, action_space=self.action_space,
                            normalizer=self.normalize_fn, initializer=self.initializer, activation=self.activation,
                            device=self.device).to(self.device)


class DuelDQN_Agent(DQN_Agent):
    """dueldqn_agent.py:10-45: DQN_Agent with the dueling network."""

    def _build_model(self):
        rep = self._build_representation(self.config.representation, self.observation_space, self.config)
        return DuelingDeepQNetwork(representation=rep, hidden_size=self.config.q_hidden_size, action_space=self.action_space,
                                   normalizer=self.normalize_fn, initializer=self.initializer, activation=self.activation,
                                   device=self.device).to(self.device)


class PerDQN_Agent(DQN_Agent):
    def __init__(self, config, envs=None, observation_space=None, action_space=None, callback=None):
        self.PER_beta0 = config.PER_beta0
        self.PER_beta = config.PER_beta0
        super().__init__(config, envs, observation_space, action_space, callback)

    def _build_memory(self, auxiliary_info_shape=None):
        self.atari = getattr(self.config, "env_name", None) == "Atari"
        return PerOffPolicyBuffer(observation_space=self.observation_space, action_space=self.action_space,
                                  auxiliary_shape=auxiliary_info_shape, n_envs=self.n_envs,
                                  buffer_size=self.buffer_size, batch_size=self.batch_size,
                                  alpha=self.config.PER_alpha, device=self.device)

    def train_epochs(self, n_epochs=1):
        """perdqn_agent.py:43-50."""
        train_info = {}
        for e in range(n_epochs):
            samples = self.memory.sample(self.PER_beta)
            td_error, train_info = self.learner.update(sync=(e == n_epochs - 1), **samples)
            self.memory.update_priorities(samples['step_choices'], td_error)
        train_info["epsilon-greedy"] = self.e_greedy
        return train_info

    def _after_update(self, train_steps):
        """perdqn_agent.py:72 - beta moves only when an update happened, by (1 - beta0) / train_steps."""
        self.PER_beta = min(1.0, self.PER_beta + (1.0 - self.PER_beta0) / max(1, train_steps))
