"""SAC agent - mirror of xuance/torch/agents/policy_gradient/sac_agent.py:19-103."""
from copy import deepcopy

import numpy as np
import torch

from ..rl_models import SAC_GaussianActor, TwinActionValueCritic, SoftActorCritic, ActivationFunctions, ActionOutput
from .off_policy import OffPolicyAgent


class SAC_Agent(OffPolicyAgent):
    def __init__(self, config, envs=None, observation_space=None, action_space=None, callback=None):
        super().__init__(config, envs, observation_space, action_space, callback)
        self.model = self._build_model()
        self.memory = self._build_memory()
        self.learner = self._build_learner(self.config, self.model, self.callback)

    def _build_model(self):
        if not hasattr(self.action_space, "low"):
            raise NotImplementedError("discrete SAC is outside the hot-path scope (BASELINE config 4 is continuous)")
        rep = self._build_representation(self.config.representation, self.observation_space, self.config)
        actor = SAC_GaussianActor(representation=rep, actor_hidden_size=self.config.actor_hidden_size,
                                  action_space=self.action_space, normalizer=self.normalize_fn,
                                  initializer=self.initializer, activation=self.activation,
                                  activation_action=ActivationFunctions[self.config.activation_action],
                                  device=self.device)
        critic = TwinActionValueCritic(representation=deepcopy(rep), action_space=self.action_space,
                                       critic_hidden_size=self.config.critic_hidden_size, normalizer=self.normalize_fn,
                                       initializer=self.initializer, activation=self.activation, device=self.device)
        return SoftActorCritic(actor=actor, critic=critic).to(self.device)

    @torch.no_grad()
    def get_actions(self, observations, test_mode=False):
        if isinstance(observations, np.ndarray):
            observations = torch.from_numpy(observations).to(self.device)
        actions = self.model.act(observations, deterministic=test_mode)
        return ActionOutput(env_actions=actions.cpu().numpy())
