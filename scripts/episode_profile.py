"""Per-step device time across a whole episode incl. the TimeLimit mass reset at step 1000 (bench workload)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import multi_car_racing_b200 as mcr
B = 1024
np.random.seed(1234)
venv = mcr.BatchedMultiCarRacing(B, num_agents=2, auto_reset="next_step", max_episode_steps=1000, seed=1234)
venv.reset(device_tracks=True)
g = torch.Generator(device=venv.device); g.manual_seed(1234)
tape = torch.rand((128, B, 2, 3), device=venv.device, generator=g); tape[..., 0] = tape[..., 0] * 2 - 1
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
N = 1100
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(N)]
ncpl = []
for s in range(N):
    flush.zero_()
    ev[s][0].record(); venv.step(tape[s % 128]); ev[s][1].record()
    if s % 10 == 0 or 995 <= s <= 1060:
        ncpl.append((s, int((venv.buffers["n_manifold"] > 0).sum().item()), int(venv.done_out.ne(0).sum().item())))
torch.cuda.synchronize()
t = np.array([a.elapsed_time(b) for a, b in ev]) * 1e3
print("mean by 50-step window:", " ".join("%d:%.0f" % (i, t[i:i + 50].mean()) for i in range(0, N, 50)))
print("steps 0..60:", " ".join("%.0f" % x for x in t[0:60]))
print("steps 995..1060:", " ".join("%.0f" % x for x in t[995:1060]))
print("(step, envs with car-car manifolds, envs done):", ncpl[:12], "...", [x for x in ncpl if 995 <= x[0] <= 1060][::5])
