"""Bench-like workload (uniform random actions incl. brake, as bench.py) for ncu launch lists."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import multi_car_racing_b200 as mcr
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 150
np.random.seed(1234)
venv = mcr.BatchedMultiCarRacing(B, num_agents=2, auto_reset=False, max_episode_steps=0, seed=1234)
venv.reset(device_tracks=True)
g = torch.Generator(device=venv.device); g.manual_seed(1234)
tape = torch.rand((128, B, 2, 3), device=venv.device, generator=g); tape[..., 0] = tape[..., 0] * 2 - 1
for s in range(steps):
    venv.step(tape[s % 128])
torch.cuda.synchronize()
print("done", venv.status(), "coupled envs:", int((venv.buffers["n_manifold"] > 0).sum().item()))
