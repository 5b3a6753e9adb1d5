import sys; sys.path.insert(0, "/root/repo")
import numpy as np, torch
import multi_car_racing_b200 as mcr
for (B, A, steps, kw) in [(1024, 2, 3500, {}), (256, 8, 2500, dict(use_ego_color=True)), (64, 16, 1500, {}), (512, 1, 2500, dict(use_random_direction=False, backwards_flag=False)), (1024, 2, 1200, dict(obs_format="gray", auto_reset=True))]:
    np.random.seed(1)
    ar = kw.pop("auto_reset", "next_step")
    venv = mcr.BatchedMultiCarRacing(B, num_agents=A, auto_reset=ar, max_episode_steps=1000, seed=5, **kw)
    venv.reset(device_tracks=True)
    g = torch.Generator(device=venv.device); g.manual_seed(3)
    tape = torch.rand((64, B, A, 3), device=venv.device, generator=g); tape[..., 0] = tape[..., 0] * 2 - 1
    dones = 0; rsum = 0.0
    for s in range(steps):
        obs, rew, done, _ = venv.step(tape[s % 64])
        if s % 50 == 0:
            dones += int((done != 0).sum().item()); rsum += float(rew.sum().item())
    torch.cuda.synchronize()
    body = venv.buffers["body"]
    ok = bool(torch.isfinite(body).all().item()) and bool(torch.isfinite(venv.buffers["reward"]).all().item())
    print("B=%d A=%d steps=%d %s: finite=%s status=%s max|pos|=%.1f obs mean %.1f sampled dones %d" % (
        B, A, steps, kw, ok, venv.status().tolist(), float(body[0, 6:8].abs().max().item()), float(obs.float().mean().item()), dones))
    del venv
