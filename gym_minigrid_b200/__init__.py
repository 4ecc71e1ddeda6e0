"""gym_minigrid_b200 -- B200-native batched MiniGrid: a drop-in for the hot path of
rohitrango/gym-minigrid (MiniGridEnv.step / gen_obs / Grid.encode / reset+_gen_grid).

Importing the package registers the env ids (like `import gym_minigrid` does, __init__.py:1-5)
but does not touch CUDA; `make()` loads libmgb200.so and fails loudly without it or without a GPU.
"""
from ._lib import MgbError  # noqa: F401
from .register import env_list, make, register, spec  # noqa: F401

__all__ = ["env_list", "make", "register", "spec", "MgbError"]
