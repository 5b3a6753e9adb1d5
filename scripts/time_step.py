"""Whole-step device time of the bench workload per episode phase, CUDA graph replay vs direct launches.
    python scripts/time_step.py [B] [steps]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import multi_car_racing_b200 as mcr
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for mode in ("graph", "direct"):
    if mode == "direct": os.environ["MCR_NO_GRAPH"] = "1"
    else: os.environ.pop("MCR_NO_GRAPH", None)
    np.random.seed(1234)
    venv = mcr.BatchedMultiCarRacing(B, num_agents=2, auto_reset="next_step", max_episode_steps=1000, seed=1234)
    venv.reset(device_tracks=True)
    g = torch.Generator(device=venv.device); g.manual_seed(1234)
    tape = torch.rand((128, B, 2, 3), device=venv.device, generator=g); tape[..., 0] = tape[..., 0] * 2 - 1
    for s in range(50): venv.step(tape[s % 128])
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(STEPS)]
    for s in range(STEPS):
        flush.zero_()
        ev[s][0].record(); venv.step(tape[(50 + s) % 128]); ev[s][1].record()
    torch.cuda.synchronize()
    t = np.array([a.elapsed_time(b) for a, b in ev]) * 1e3
    w = max(1, STEPS // 4)
    print(mode, "mean %.1f us | quarters %s | max %.1f | launches/step %.1f" % (
        t.mean(), " ".join("%.1f" % t[i:i + w].mean() for i in range(0, STEPS, w)), t.max(), venv.launch_count / (STEPS + 50)))
    # back-to-back (no flush, no events between steps): pure pipeline rate
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(300): venv.step(tape[s % 128])
    e1.record(); torch.cuda.synchronize()
    print(mode, "back-to-back %.1f us/step" % (e0.elapsed_time(e1) * 1e3 / 300))
    del venv
