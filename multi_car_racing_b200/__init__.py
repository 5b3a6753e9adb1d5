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


def _register():
    for modname in ("gym", "gymnasium"):
        try:
            mod = __import__(modname + ".envs.registration", fromlist=["register"])
            mod.register(id=ENV_ID, entry_point="multi_car_racing_b200:MultiCarRacing",
                         max_episode_steps=MAX_EPISODE_STEPS, reward_threshold=REWARD_THRESHOLD)
        except Exception:
            pass


_register()
