"""Device-facing view of a vector env - counterpart of xuance/torch/utils/tensor_env.py:8-51 (``TensorEnvWrapper``).

The reference wrapper moves actions D->H, steps the host envs and uploads observations / rewards / flags as float32
(uint8 frames become 4x the bytes).  Here the observation keeps its dtype (uint8 frames stay uint8:
7.2 MB instead of 28.9 MB per step for 256 Atari envs), goes through ONE pinned staging block (or straight from the
shared-memory block of ``ShmSubprocVecEnv``) and one async H2D copy; the HBM buffers' ``store`` takes the tensors as is."""
import numpy as np
import torch


class TensorEnvWrapper:
    def __init__(self, envs, device):
        self.envs = envs
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("TensorEnvWrapper needs a CUDA device (no CPU fallback)")
        self.num_envs = envs.num_envs
        self.observation_space, self.action_space = envs.observation_space, envs.action_space
        self.max_episode_steps = envs.max_episode_steps
        self._stage = [torch.from_numpy(np.zeros_like(envs.buf_obs)).pin_memory() for _ in range(2)]
        self._i = 0
        self.buf_obs = self._to_device(envs.buf_obs)

    def _to_device(self, obs):
        t = torch.from_numpy(obs) if isinstance(obs, np.ndarray) else obs
        if t.is_pinned():
            return t.to(self.device, non_blocking=True)
        st = self._stage[self._i]
        self._i ^= 1
        st.copy_(t)
        return st.to(self.device, non_blocking=True)

    def reset(self):
        obs, infos = self.envs.reset()
        self.buf_obs = self._to_device(obs)
        return self.buf_obs, infos

    def step(self, actions):
        if isinstance(actions, torch.Tensor):
            actions = actions.detach().cpu().numpy()
        obs, rew, term, trunc, infos = self.envs.step(actions)
        self.buf_obs = self._to_device(obs)
        dev = self.device
        return (self.buf_obs, torch.from_numpy(rew).to(dev, non_blocking=True),
                torch.from_numpy(term).to(dev, non_blocking=True), torch.from_numpy(trunc).to(dev, non_blocking=True), infos)

    def render(self, mode):
        return self.envs.render(mode)

    def close(self):
        self.envs.close()
