"""Run N steady-state steps of the bench workload (no auto reset) -- target for ncu."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import multi_car_racing_b200 as mcr
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
np.random.seed(0)
venv = mcr.BatchedMultiCarRacing(B, num_agents=2, auto_reset=False, max_episode_steps=0, seed=0)
venv.reset(device_tracks=True)
g = torch.Generator(device=venv.device); g.manual_seed(0)
tape = torch.rand((64, B, 2, 3), device=venv.device, generator=g); tape[..., 0] = tape[..., 0] * 2 - 1
for s in range(steps):
    venv.step(tape[s % 64])
torch.cuda.synchronize()
print("done", venv.status())
