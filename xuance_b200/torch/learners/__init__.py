from .learner import Learner
from .ppo_learner import PPO_Learner

REGISTRY_Learners = {"PPO_Learner": PPO_Learner, "PPOCLIP_Learner": PPO_Learner}
