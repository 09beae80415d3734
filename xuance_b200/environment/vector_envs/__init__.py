from .vector_env import VecEnv, AlreadySteppingError, NotSteppingError
from .dummy_vec_env import DummyVecEnv, DummyVecEnv_Atari
from .subproc_vec_env import SubprocVecEnv, SubprocVecEnv_Atari
from .shm_vec_env import ShmSubprocVecEnv, ShmSubprocVecEnv_Atari
from .dummy_vec_maenv import DummyVecMultiAgentEnv
from .subproc_vec_maenv import SubprocVecMultiAgentEnv

# reference: xuance/environment/vector_envs/__init__.py:36-49
REGISTRY_VEC_ENV = {"DummyVecEnv": DummyVecEnv, "Dummy_Atari": DummyVecEnv_Atari, "SubprocVecEnv": SubprocVecEnv,
                    "Subproc_Atari": SubprocVecEnv_Atari, "Dummy_Gym": DummyVecEnv, "Subproc_Gym": SubprocVecEnv,
                    "ShmSubprocVecEnv": ShmSubprocVecEnv, "ShmSubproc_Atari": ShmSubprocVecEnv_Atari,
                    "DummyVecMultiAgentEnv": DummyVecMultiAgentEnv, "Dummy_StarCraft2": DummyVecMultiAgentEnv,
                    "SubprocVecMultiAgentEnv": SubprocVecMultiAgentEnv, "Subproc_StarCraft2": SubprocVecMultiAgentEnv}
