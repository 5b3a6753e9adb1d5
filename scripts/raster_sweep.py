"""BASELINE.json configs[4]: rasteriser HBM sweep, num_agents=2, batch in {1k, 4k, 16k, 65k} envs on one
GPU: render_kernel time (CUDA events, L2 flushed between launches) on frozen mid-episode states
(t >= 1 s) and on the zoomed-out first frame (t = 0.02 s), GB/s of observation bytes against the
measured HBM peak; plus the whole-step time at that batch.
    python scripts/raster_sweep.py [out.json] [B ...]"""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import multi_car_racing_b200 as mcr
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/raster_sweep.json"
Bs = [int(x) for x in sys.argv[2:]] or [1024, 4096, 16384, 65536]
A = 2
peak = 6536.4
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timed(fn, reps):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        flush.zero_(); a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return float(np.mean([a.elapsed_time(b) for a, b in ev]))


rows = []
for B in Bs:
    np.random.seed(1)
    t0 = time.perf_counter()
    venv = mcr.BatchedMultiCarRacing(B, num_agents=A, auto_reset="next_step", max_episode_steps=1000, seed=7)
    venv.reset(device_tracks=True)
    torch.cuda.synchronize()
    t_reset = time.perf_counter() - t0
    reps = 20 if B <= 16384 else 8
    ms_zoomed_out = timed(venv.render_only, reps)
    g = torch.Generator(device=venv.device); g.manual_seed(1)
    tape = torch.rand((16, B, A, 3), device=venv.device, generator=g); tape[..., 0] = tape[..., 0] * 2 - 1
    for s in range(80):
        venv.step(tape[s % 16])
    ms_mid = timed(venv.render_only, reps)
    k = [0]
    def one_step():
        venv.step(tape[k[0] % 16]); k[0] += 1
    ms_step = timed(one_step, reps)
    frames, nbytes = B * A, B * A * 96 * 96 * 3
    row = {"batch_envs": B, "num_agents": A, "frames": frames, "obs_bytes": nbytes,
           "render_ms_mid_episode": ms_mid, "render_GBs_mid_episode": nbytes / ms_mid / 1e6, "frac_mid_episode": nbytes / ms_mid / 1e6 / peak,
           "render_ms_t0.02": ms_zoomed_out, "render_GBs_t0.02": nbytes / ms_zoomed_out / 1e6, "frac_t0.02": nbytes / ms_zoomed_out / 1e6 / peak,
           "step_ms": ms_step, "step_agent_frames_per_s": frames / ms_step * 1e3, "reset_device_tracks_s": t_reset,
           "status": venv.status().tolist()}
    rows.append(row)
    print(json.dumps(row), flush=True)
    del venv, tape
    torch.cuda.empty_cache()
json.dump({"hbm_peak_GBs": peak, "l2": "256 MiB flush before every timed launch", "rows": rows}, open(out, "w"), indent=1)
