"""Device vs host track generation time.   python scripts/time_trackgen.py [n]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import multi_car_racing_b200 as mcr
from multi_car_racing_b200.track import TrackGenerator
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
venv = mcr.BatchedMultiCarRacing(n, num_agents=2, auto_reset=False, max_episode_steps=0)
rngs = [np.random.RandomState(i) for i in range(n)]
t0 = time.perf_counter(); res = venv.generate_tracks_device(list(range(n)), rngs); torch.cuda.synchronize(); t1 = time.perf_counter()
print("device: %d tracks in %.3f s end to end (incl. RNG state shuttling and node read-back), mean attempts %.2f" % (n, t1 - t0, res[:, 1].mean()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
stride = int(venv.L.mcr_trackgen_scratch_bytes())
mt = np.stack([np.concatenate([r.get_state()[1], [r.get_state()[2]]]).astype(np.uint32) for r in rngs])
d_mt = torch.from_numpy(mt.view(np.int32)).cuda(); d_slot = torch.arange(n, dtype=torch.int32, device="cuda")
d_scr = torch.empty((n, stride), dtype=torch.uint8, device="cuda"); d_res = torch.zeros((n, 4), dtype=torch.int32, device="cuda")
e0.record(); venv.L.mcr_tracks_generate_device(venv._h, n, d_mt.data_ptr(), d_slot.data_ptr(), d_scr.data_ptr(), d_res.data_ptr(), venv._stream()); e1.record()
torch.cuda.synchronize()
print("device kernel alone: %.2f ms for %d tracks" % (e0.elapsed_time(e1), n))
gen = TrackGenerator(); m = min(n, 256)
t0 = time.perf_counter()
for i in range(m): gen.generate(np.random.RandomState(i))
print("host generator: %.3f ms per track, one core" % (1e3 * (time.perf_counter() - t0) / m))
