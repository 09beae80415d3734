from .agent import Agent, set_seed, set_device
from .on_policy import OnPolicyAgent
from .ppo_agent import PPO_Agent

REGISTRY_Agents = {"PPO": PPO_Agent, "PPO_Clip": PPO_Agent}
