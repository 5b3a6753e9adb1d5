"""Where one step's time goes INSIDE the CUDA graph: %globaltimer stamps written by the kernels
themselves (start of each kernel, latest CTA end of sweep / render), averaged over steps.
    python scripts/timeline.py [B] [steps] [warm steps] [num_agents]"""
import os, sys, signal
signal.signal(signal.SIGPIPE, signal.SIG_DFL)     # BrokenPipe-safe when piped into head
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import multi_car_racing_b200 as mcr
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 200
WARM = int(sys.argv[3]) if len(sys.argv) > 3 else 60      # e.g. 900: late-episode steps (cars collide more)
A = int(sys.argv[4]) if len(sys.argv) > 4 else 2
names = ["head", "contacts", "stripes", "sweep", "coupled", "post", "score", "render", "render_end", "post2", "render2", "sweep_end_percar", "sweep_end_packed", "coupled_vel_end", "coupled_pos_end", "fill", "post_end", "project_end", "head_end",
         "fill_first_in~", "fill_first_go~", "project_first_end~", "project_last_go", "fill_last_in"]     # "~": stored complemented (earliest stamp)
np.random.seed(1234)
venv = mcr.BatchedMultiCarRacing(B, num_agents=A, auto_reset="next_step", max_episode_steps=1000, seed=1234)
venv.reset(device_tracks=True)
g = torch.Generator(device=venv.device); g.manual_seed(1234)
tape = torch.rand((128, B, A, 3), device=venv.device, generator=g); tape[..., 0] = tape[..., 0] * 2 - 1
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for s in range(WARM): venv.step(tape[s % 128])
tl = venv.buffers["timeline"].view(torch.int64)
acc = np.zeros(len(names)); n = 0; tot = 0.0
for s in range(STEPS):
    if not os.environ.get("NOFLUSH"): flush.zero_()      # NOFLUSH=1: leave the previous step's lines in L2
    tl.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    act = tape[(WARM + s) % 128]
    if os.environ.get("ACTION_INPLACE"):              # the action already lies in the library's staging buffer: no device-to-device copy in the step
        stage = venv.buffers["action_stage"].view(torch.float32).view(-1)[:B * A * 3].view(B, A, 3)
        stage.copy_(act); act = stage
    e0.record(); venv.step(act); e1.record()
    torch.cuda.synchronize()
    ti = tl.cpu().numpy()[:len(names)].copy()
    for k, nm_ in enumerate(names):
        if nm_.endswith("~") and ti[k] != 0: ti[k] = ~ti[k]
    t = ti.astype(np.float64)
    acc += np.where(t > 0, (t - t[0]) / 1e3, np.nan); n += 1; tot += e0.elapsed_time(e1) * 1e3
nm = venv.buffers["n_manifold"].cpu().numpy()
print("envs with car-car manifolds in the last step: %d of %d (max manifolds %d)" % ((nm > 0).sum(), B, nm.max()))
print("step (events) %.1f us; kernel start stamps relative to head_kernel start, us:" % (tot / n))
for k, v in sorted(zip(names, acc / n), key=lambda kv: (np.isnan(kv[1]), kv[1])):
    print("  %-18s %8.1f" % (k, v) if not np.isnan(v) else "  %-18s  (did not run in every step)" % k)

