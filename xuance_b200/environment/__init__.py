"""Environment construction (reference: xuance/environment/__init__.py:12-76 ``make_envs``)."""
from .envs import REGISTRY_ENV, XuanCeEnvWrapper, CartPoleEnv, SyntheticAtariEnv
from .vector_envs import (REGISTRY_VEC_ENV, VecEnv, DummyVecEnv, DummyVecEnv_Atari, SubprocVecEnv, SubprocVecEnv_Atari,
                          ShmSubprocVecEnv, ShmSubprocVecEnv_Atari)
from .tensor_env import TensorEnvWrapper


def make_envs(config):
    """``config.env_id`` selects the raw env, ``config.vectorize`` the vector wrapper, ``config.parallels`` the count.
    Under distributed training every rank builds its own ``parallels / world_size`` envs with rank-distinct seeds
    (the reference hard-codes rank = 1 here, appendix B #15)."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1")) if getattr(config, "distributed_training", False) else 1
    rank = int(os.environ.get("RANK", "0")) if world > 1 else 0
    n = config.parallels // world
    env_cls = REGISTRY_ENV[config.env_id]
    base_seed = getattr(config, "env_seed", 1) + rank * n

    def thunk(i):
        return lambda: XuanCeEnvWrapper(env_cls(seed=base_seed + i))

    vec_cls = REGISTRY_VEC_ENV[getattr(config, "vectorize", "DummyVecEnv")]
    return vec_cls([thunk(i) for i in range(n)], base_seed)
