"""multi_car_racing_b200 -- B200-native batched MultiCarRacing-v0 step + render path.

Drop-in surface of igilitschenski/multi_car_racing (gym_multi_car_racing):

    import multi_car_racing_b200 as mcr
    env = mcr.make("MultiCarRacing-v0", num_agents=2, direction='CCW', use_random_direction=True,
                   backwards_flag=True, h_ratio=0.25, use_ego_color=False)
    obs = env.reset()                               # (num_agents, 96, 96, 3) uint8
    obs, reward, done, info = env.step(action)      # action (num_agents, 3)

    venv = mcr.BatchedMultiCarRacing(batch_envs=1024, num_agents=2)   # torch tensors, on device
    vec = mcr.MultiCarRacingVecEnv(1024, num_agents=2)                 # gym(nasium) VectorEnv protocol

If a `gym` (or `gymnasium`) package is importable the id "MultiCarRacing-v0" is also registered
there with the reference's max_episode_steps=1000 / reward_threshold=900
(reference gym_multi_car_racing/__init__.py:5-10).
"""
from ._lib import McrError, LIB_PATH  # noqa: F401
from .env import BatchedMultiCarRacing, MultiCarRacing, MultiCarRacingVecEnv, TimeLimit, Box  # noqa: F401

__all__ = ["make", "MultiCarRacing", "BatchedMultiCarRacing", "MultiCarRacingVecEnv", "TimeLimit", "Box", "McrError", "ENV_ID"]

ENV_ID = "MultiCarRacing-v0"
MAX_EPISODE_STEPS = 1000
REWARD_THRESHOLD = 900


def make(id=ENV_ID, **kwargs):
    """gym.make(id, **kwargs) for the one id this package provides: the env wrapped in the
    registration's TimeLimit(1000)."""
    if id != ENV_ID:
        raise ValueError("unknown environment id %r (this package provides %r)" % (id, ENV_ID))
    return TimeLimit(MultiCarRacing(**kwargs), MAX_EPISODE_STEPS)


REGISTERED_WITH = []       # the gym packages the id was registered with at import time


def _register():
    """register(id='MultiCarRacing-v0', max_episode_steps=1000, reward_threshold=900) with every gym package that is
    importable (reference gym_multi_car_racing/__init__.py:5-10).  A missing package is skipped silently; a package that
    is present but refuses the registration (other than "already registered") is reported as a warning, not swallowed."""
    import importlib
    import warnings
    for modname in ("gym", "gymnasium"):
        try:
            mod = importlib.import_module(modname + ".envs.registration")
        except ImportError:
            continue
        try:
            mod.register(id=ENV_ID, entry_point="multi_car_racing_b200:MultiCarRacing",
                         max_episode_steps=MAX_EPISODE_STEPS, reward_threshold=REWARD_THRESHOLD)
            REGISTERED_WITH.append(modname)
        except Exception as e:      # gym raises gym.error.Error on a duplicate id: a re-import, fine
            if "registered" in str(e).lower() or "re-register" in str(e).lower():
                REGISTERED_WITH.append(modname)
            else:
                warnings.warn("multi_car_racing_b200: could not register %s with %s: %r" % (ENV_ID, modname, e))


_register()
