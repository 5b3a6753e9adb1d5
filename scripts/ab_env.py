"""A/B a tuning environment variable on the bench workload inside one process (same box, same clocks):
    python scripts/ab_env.py VAR v1 v2 ... [--steps K]     whole-step device time per value, interleaved rounds."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import multi_car_racing_b200 as mcr
var, vals = sys.argv[1], [v for v in sys.argv[2:] if not v.startswith("--")]
STEPS = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 300
B = 1024
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
res = {v: [] for v in vals}
for rnd in range(3):
    for v in vals:
        os.environ[var] = v
        np.random.seed(1234)
        venv = mcr.BatchedMultiCarRacing(B, num_agents=2, auto_reset="next_step", max_episode_steps=1000, seed=1234)
        venv.reset(device_tracks=True)
        g = torch.Generator(device=venv.device); g.manual_seed(1234)
        tape = torch.rand((128, B, 2, 3), device=venv.device, generator=g); tape[..., 0] = tape[..., 0] * 2 - 1
        for s in range(60): venv.step(tape[s % 128])
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(STEPS)]
        for s in range(STEPS):
            flush.zero_()
            ev[s][0].record(); venv.step(tape[(60 + s) % 128]); ev[s][1].record()
        torch.cuda.synchronize()
        res[v].append(np.mean([a.elapsed_time(b) for a, b in ev]) * 1e3)
        del venv
for v in vals:
    print("%s=%-8s step %s us (mean %.1f)" % (var, v, " ".join("%.1f" % x for x in res[v]), np.mean(res[v])))
