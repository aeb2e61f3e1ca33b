"""bullet_envs_b200 -- B200-native batched simulator for the snakeRL ``SnakeGymEnv.step()`` hot path."""
from ._abi import default_params, gait_params  # noqa: F401
from .spaces import Box  # noqa: F401
from .urdf_model import ImportRules, build_model  # noqa: F401


def __getattr__(name):  # torch is imported lazily (the first import can take a minute on a fresh box)
    if name == "SnakeVecEnv":
        from .vec_env import SnakeVecEnv
        return SnakeVecEnv
    if name in ("SnakeGymEnv", "Snake"):
        from . import gym_env
        return getattr(gym_env, name)
    if name in ("RolloutBuffer", "compute_gae"):
        import importlib
        return getattr(importlib.import_module(__name__ + ".rollout"), name)
    if name == "dist":
        import importlib
        return importlib.import_module(__name__ + ".dist")
    raise AttributeError(name)
