"""Cycles per part of head_kernel (instrumented build, lane 0 of every warp = env).
    python -m multi_car_racing_b200.build --phase-clocks && MCR_LIB_PATH=multi_car_racing_b200/libmcr_clk.so python scripts/head_phases.py [B] [A]"""
import os, sys, ctypes
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import multi_car_racing_b200 as mcr
from multi_car_racing_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
A = int(sys.argv[2]) if len(sys.argv) > 2 else 2
L = _lib.load()
np.random.seed(1234)
venv = mcr.BatchedMultiCarRacing(B, num_agents=A, auto_reset="next_step", max_episode_steps=1000, seed=1234)
venv.reset(device_tracks=True)
g = torch.Generator(device=venv.device); g.manual_seed(1234)
tape = torch.rand((128, B, A, 3), device=venv.device, generator=g); tape[..., 0] = tape[..., 0] * 2 - 1
for s in range(100): venv.step(tape[s % 128])
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 8)()
L.mcr_debug_head_clocks(None, 1)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for s in range(50):
    flush.zero_(); venv.step(tape[(100 + s) % 128])
torch.cuda.synchronize()
L.mcr_debug_head_clocks(buf, 1)
v = np.array(list(buf), np.float64); n = max(v[7], 1)
for k, nm in [(4, "entry + auto-reset check"), (5, "car-car narrow phase"), (0, "pre_car: loads"), (1, "pre_car: controls + tyre model"),
              (2, "pre_car: integrate + joints_init"), (3, "pre_car: stores")]:
    print("  %-36s %7.0f cycles per env" % (nm, v[k] / n))
print("  total                                %7.0f   (%d warp samples)" % (v[:6].sum() / n, n))
