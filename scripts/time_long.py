import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import multi_car_racing_b200 as mcr
B = 1024
np.random.seed(1234)
venv = mcr.BatchedMultiCarRacing(B, num_agents=2, auto_reset=False, max_episode_steps=0, seed=1234)
venv.reset(device_tracks=True)
g = torch.Generator(device=venv.device); g.manual_seed(1234)
tape = torch.rand((128, B, 2, 3), device=venv.device, generator=g); tape[..., 0] = tape[..., 0] * 2 - 1
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
rows = []
for s in range(1000):
    venv.step_split(tape[s % 128], ev)
    torch.cuda.synchronize()
    rows.append((ev[0].elapsed_time(ev[1]) * 1e3, ev[1].elapsed_time(ev[2]) * 1e3, int((venv.buffers["n_manifold"] > 0).sum().item())))
rows = np.array(rows)
for a in range(0, 1000, 100):
    r = rows[a:a + 100]
    print("steps %4d-%4d  sim %.1f us (max %.1f)  render %.1f us  coupled envs avg %.1f max %d" % (a, a + 99, r[:, 0].mean(), r[:, 0].max(), r[:, 1].mean(), r[:, 2].mean(), r[:, 2].max()))
