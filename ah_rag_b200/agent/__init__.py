from .environment import GraphEnvironment, VecGraphEnvironment  # noqa: F401
from .gym_env import AHRAGGymEnv, VecAHRAGGymEnv  # noqa: F401
