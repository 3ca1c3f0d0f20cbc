"""Drop-in `PPO` package: same import surface as the reference (`from PPO import PPO, ActorCritic, RND, Memory`,
/root/reference/PPO/__init__.py:1-4), backed by the sm_100a kernels in libprl_b200.so."""
from .ActorCritic import ActorCritic  # noqa: F401
from .RND import RND  # noqa: F401
from .Memory import Memory  # noqa: F401
from .PPO import PPO  # noqa: F401
