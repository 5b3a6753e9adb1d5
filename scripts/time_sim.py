import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import multi_car_racing_b200 as mcr
from multi_car_racing_b200 import _lib
for B in [int(x) for x in sys.argv[1:]]:
    np.random.seed(0)
    venv = mcr.BatchedMultiCarRacing(B, num_agents=2, auto_reset=False, max_episode_steps=0, seed=0)
    venv.reset(device_tracks=True)
    g = torch.Generator(device=venv.device); g.manual_seed(0)
    tape = torch.rand((64, B, 2, 3), device=venv.device, generator=g); tape[..., 0] = tape[..., 0] * 2 - 1
    for s in range(60): venv.step(tape[s % 64])
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ts, tr = [], []
    for s in range(100):
        venv.step_split(tape[s % 64], ev)
        torch.cuda.synchronize()
        ts.append(ev[0].elapsed_time(ev[1])); tr.append(ev[1].elapsed_time(ev[2]))
    print("B=%d sim %.1f us (min %.1f)  render %.1f us (min %.1f)" % (B, 1e3*np.mean(ts), 1e3*np.min(ts), 1e3*np.mean(tr), 1e3*np.min(tr)))
    del venv
