"""Cycles per part of a coupled-island velocity iteration (instrumented build, lane 0 of every coupled warp).
    python -m multi_car_racing_b200.build --phase-clocks && MCR_LIB_PATH=multi_car_racing_b200/libmcr_clk.so python scripts/coupled_phases.py [B] [A] [warm]"""
import os, sys, ctypes
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import multi_car_racing_b200 as mcr
from multi_car_racing_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
A = int(sys.argv[2]) if len(sys.argv) > 2 else 2
WARM = int(sys.argv[3]) if len(sys.argv) > 3 else 900
L = _lib.load()
np.random.seed(1234)
venv = mcr.BatchedMultiCarRacing(B, num_agents=A, auto_reset="next_step", max_episode_steps=1000, seed=1234)
venv.reset(device_tracks=True)
g = torch.Generator(device=venv.device); g.manual_seed(1234)
tape = torch.rand((128, B, A, 3), device=venv.device, generator=g); tape[..., 0] = tape[..., 0] * 2 - 1
for s in range(WARM): venv.step(tape[s % 128])
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 8)()
L.mcr_debug_coupled_clocks(None, 1)
for s in range(50): venv.step(tape[(WARM + s) % 128])
torch.cuda.synchronize()
L.mcr_debug_coupled_clocks(buf, 1)
v = np.array(list(buf), np.float64); n = max(v[7], 1)
nm = venv.buffers["n_manifold"].cpu().numpy()
print("coupled envs in the last step: %d (max manifolds %d); %d warp-iterations sampled" % ((nm > 0).sum(), nm.max(), n))
for k, nmk in enumerate(["joints sweep", "push + syncwarp", "contacts + syncwarp", "pull"]):
    print("  %-22s %7.0f cycles per iteration" % (nmk, v[k] / n))
print("  total                  %7.0f" % (v[:4].sum() / n))
