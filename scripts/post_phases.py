"""Cycles per part of post_kernel (instrumented build, thread 0 of every CTA).
    python -m multi_car_racing_b200.build --phase-clocks && MCR_LIB_PATH=multi_car_racing_b200/libmcr_clk.so python scripts/post_phases.py [B] [A]"""
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
L.mcr_debug_post_clocks(None, 1)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for s in range(50):
    flush.zero_(); venv.step(tape[(100 + s) % 128])
torch.cuda.synchronize()
L.mcr_debug_post_clocks(buf, 1)
v = np.array(list(buf), np.float64); n = max(v[7], 1)
for k, nm in enumerate(["grid-dependency wait", "loads", "integrate + position iterations", "transforms + sleep", "stores", "camera + heading"]):
    print("  %-34s %7.0f cycles per CTA" % (nm, v[k] / n))
print("  total                              %7.0f   (%d CTA samples)" % (v[:6].sum() / n, n))
hist = (ctypes.c_ulonglong * 64)()
L.mcr_debug_post_clocks(hist, -1)
hv = np.array(list(hist), np.float64)
if hv.sum() > 0:
    print("position iterations per car-step (all cars):", " ".join("%d:%.4f" % (k, hv[k] / hv.sum()) for k in range(64) if hv[k] > 0))
wt = (ctypes.c_ulonglong * 4096)()
if L.mcr_debug_post_clocks(wt, -2) == 0:
    w = np.array(list(wt), np.float64).reshape(4, 1024)[:, :(2 * B * A + 31) // 32] / 1e3
    t0 = w[0].min()
    for k, nm in enumerate(["entry", "past contacts / stripes flags", "past the sweep flag", "end (car published)"]):
        x = np.sort(w[k] - t0)
        print("  post warps, %-30s min %.1f  median %.1f  p90 %.1f  p99 %.1f  max %.1f us" % (nm, x[0], x[len(x) // 2], x[int(len(x) * .9)], x[int(len(x) * .99)], x[-1]))
    dur = np.sort(w[3] - w[2])
    print("  per-warp time from the sweep flag to the end: median %.1f p90 %.1f p99 %.1f max %.1f us; slowest warps:" % (dur[len(dur) // 2], dur[int(len(dur) * .9)], dur[int(len(dur) * .99)], dur[-1]),
          np.argsort(w[3] - w[2])[-8:].tolist())
