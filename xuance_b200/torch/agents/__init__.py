from .agent import Agent, set_seed, set_device
from .on_policy import OnPolicyAgent
from .off_policy import OffPolicyAgent
from .ppo_agent import PPO_Agent
from .dqn_agent import DQN_Agent, PerDQN_Agent, DuelDQN_Agent

REGISTRY_Agents = {"PPO": PPO_Agent, "PPO_Clip": PPO_Agent, "DQN": DQN_Agent, "PerDQN": PerDQN_Agent, "DuelDQN": DuelDQN_Agent}
try:
    from .sac_agent import SAC_Agent
    REGISTRY_Agents["SAC"] = SAC_Agent
except ImportError:
    pass
from .marl import MARLAgents, OffPolicyMARLAgents, QMIX_Agents, VDN_Agents
REGISTRY_Agents["QMIX"] = QMIX_Agents
REGISTRY_Agents["VDN"] = VDN_Agents
