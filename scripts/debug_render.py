import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import numpy as np, torch
import mcr_oracle as oracle
import multi_car_racing_b200 as mcr
from test_gpu_parity import _setup
from helpers import action_tape
venv, worlds, tracks, obs0, oobs0 = _setup(oracle, mcr, B=3, A=2, seed=1)
d = (obs0 != oobs0).any(-1)
print("reset diff pixels per frame", d.reshape(6, -1).sum(1))
ys, xs = np.nonzero(d[0, 0])
print("rows", np.unique(ys)[:20], "cols", np.unique(xs)[:20])
for y, x in list(zip(ys, xs))[:10]:
    print(y, x, obs0[0, 0, y, x], oobs0[0, 0, y, x])
os.makedirs("gpurun_out", exist_ok=True)
frames_g, frames_o = [obs0], [oobs0]
tape = action_tape(1, 100, 3, 2)
for s in range(100):
    obs, rew, done, _ = venv.step(torch.from_numpy(tape[s]).to(venv.device))
    oo = np.stack([w.step(tape[s, e].astype(np.float64))[0] for e, w in enumerate(worlds)])
    if s in (0, 5, 30, 60, 99):
        frames_g.append(obs.cpu().numpy().copy()); frames_o.append(oo)
        print("step", s, "diff px", (frames_g[-1] != oo).any(-1).reshape(6, -1).sum(1))
np.savez_compressed("gpurun_out/render_debug.npz", g=np.stack(frames_g), o=np.stack(frames_o))
print("status", venv.status())
