"""Wall time of reset(device_tracks=True) and of the constructor at several batch sizes.
    python scripts/time_reset.py [B ...]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import multi_car_racing_b200 as mcr
for B in [int(x) for x in sys.argv[1:]] or [1024, 8192, 65536]:
    np.random.seed(1)
    t0 = time.perf_counter()
    venv = mcr.BatchedMultiCarRacing(B, num_agents=2, auto_reset="next_step", max_episode_steps=1000, seed=7)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    venv.reset(device_tracks=True); torch.cuda.synchronize(); t2 = time.perf_counter()
    venv.reset(device_tracks=True); torch.cuda.synchronize(); t3 = time.perf_counter()
    a = torch.zeros((B, 2, 3), device=venv.device)
    for _ in range(8): venv.step(a)
    torch.cuda.synchronize(); t4 = time.perf_counter()
    print("B=%6d  ctor %.3f s  first reset %.3f s  second reset %.3f s  8 steps right after (spare tracks generating beside them) %.1f ms  status %s"
          % (B, t1 - t0, t2 - t1, t3 - t2, (t4 - t3) * 1e3, venv.status().tolist()), flush=True)
    del venv; torch.cuda.empty_cache()
