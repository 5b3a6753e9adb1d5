"""Device time of render_kernel alone (mcr_render, post_step=0) on the bench workload at three episode phases:
zoomed-out first frames (t = 0.02 and 0.5 s) and a mid-episode step; 256 MiB L2 flush between timed launches.
    python scripts/render_perf.py [B] [A] [reps]        (NCU=1: two untimed launches per phase, for ncu -k render_kernel)"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import multi_car_racing_b200 as mcr
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
A = int(sys.argv[2]) if len(sys.argv) > 2 else 2
REPS = int(sys.argv[3]) if len(sys.argv) > 3 else 30
np.random.seed(1234)
venv = mcr.BatchedMultiCarRacing(B, num_agents=A, auto_reset=False, max_episode_steps=0, seed=1234)
venv.reset(device_tracks=True)
g = torch.Generator(device=venv.device); g.manual_seed(1234)
tape = torch.rand((128, B, A, 3), device=venv.device, generator=g); tape[..., 0] = tape[..., 0] * 2 - 1
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
done = 0
out = {}
for name, upto in (("t=0.02", 0), ("t=0.5", 24), ("mid(170)", 170)):
    while done < upto:
        venv.step(tape[done % 128]); done += 1
    torch.cuda.synchronize()
    if os.environ.get("NCU"):                       # ncu --profile-from-start off: only these launches are captured
        if os.environ["NCU"] in ("all", name):
            torch.cuda.cudart().cudaProfilerStart()
            venv.render_only(); torch.cuda.synchronize()
            torch.cuda.cudart().cudaProfilerStop()
        continue
    for _ in range(3): venv.render_only()
    ts = []
    for _ in range(REPS):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); venv.render_only(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts = np.array(ts)
    gbs = B * A * 27648 / (ts.mean() * 1e-6) / 1e9
    out[name] = ts.mean()
    print("%-9s render %.1f us mean, %.1f min  -> %.0f GB/s algorithmic (%d frames)" % (name, ts.mean(), ts.min(), gbs, B * A))
import hashlib
print("obs sha1", hashlib.sha1(venv.obs.cpu().numpy().tobytes()).hexdigest(), "status", venv.status().tolist())
