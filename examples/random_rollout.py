"""Quick start: 1024 MultiCarRacing-v0 envs x 2 agents on one B200 under a random policy.

    python examples/random_rollout.py [steps]

Shows the three surfaces: the batched device API (torch tensors in / out, no host copies), the
gym(nasium)-style vector env, and the reference's single-env API (numpy in / out)."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

import multi_car_racing_b200 as mcr

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300

# ---- batched, device resident ------------------------------------------------------------------
venv = mcr.BatchedMultiCarRacing(1024, num_agents=2, auto_reset='next_step', seed=0)
obs = venv.reset(device_tracks=True)                      # (1024, 2, 96, 96, 3) uint8 on cuda:0
act = torch.empty((1024, 2, 3), device=venv.device)
torch.cuda.synchronize(); t0 = time.perf_counter()
ret = torch.zeros((1024, 2), dtype=torch.float64, device=venv.device)
for _ in range(steps):
    act.uniform_(0, 1); act[..., 0].mul_(2).sub_(1)       # steer in [-1, 1], gas / brake in [0, 1]
    obs, reward, done, _ = venv.step(act)                 # done: bit0 terminated, bit1 TimeLimit
    ret += reward
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("batched: %d agent-frames in %.3f s = %.2e agent-frames/s, mean return %.1f" % (
    steps * 2048, dt, steps * 2048 / dt, float(ret.mean())))
frame = venv.render('rgb_array')[0, 0].cpu().numpy()      # (400, 600, 3) picture of env 0, agent 0
print("rgb_array frame", frame.shape, frame.dtype)

# ---- gym(nasium) VectorEnv protocol -----------------------------------------------------------------
vec = mcr.MultiCarRacingVecEnv(64, num_agents=2, seed=1)
obs, info = vec.reset()
obs, reward, terminated, truncated, info = vec.step(torch.rand((64, 2, 3), device=obs.device))
print("vector env:", tuple(obs.shape), tuple(reward.shape), terminated.dtype, truncated.dtype)

# ---- the reference's own API --------------------------------------------------------------------------
env = mcr.make("MultiCarRacing-v0", num_agents=2, direction='CCW', use_random_direction=True,
               backwards_flag=True, h_ratio=0.25, use_ego_color=False, verbose=0)
obs = env.reset()                                         # (2, 96, 96, 3) uint8 numpy
obs, reward, done, info = env.step(np.array([[0.0, 1.0, 0.0], [0.1, 0.5, 0.0]]))
print("single env:", obs.shape, reward, done, info, "on grass:", env.unwrapped.driving_on_grass)
