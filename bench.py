#!/usr/bin/env python
"""bench.py -- agent-frames/s of the batched MultiCarRacing-v0 step+render path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one env.step over one rank's batch: contacts + physics + render/post-step for
batch_envs x num_agents cars (BASELINE.json configs[1]: num_agents=2, batch=1024 envs per GPU,
random policy).  One agent-frame = one (96,96,3) uint8 observation produced by a full step.

  value     device-resident throughput (actions already in HBM), CUDA events, max over ranks
  e2e       the same through BatchedMultiCarRacing.step_host: actions from pinned host memory,
            observations/rewards/dones copied back to pinned host memory every step
  roofline  the rasteriser (project_kernel + fill_kernel, timed alone with CUDA events -- no reward /
            done block beside it): algorithmic bytes (27 648 B per agent-frame) / launch time vs
            MEASURED_PEAKS.json hbm_gbs; fill_kernel's own share from the %globaltimer stamps the
            kernels write (kernel_ms.fill)
  cpu_baseline  the CPU oracle (a port of the reference's algorithm, NOT the original
            Box2D+pyglet, which is not installable here) timed on host cores

--impl reference times that CPU port with every host core (one process per core) and prints
the same JSON line with "impl": "reference".
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

OBS_BYTES = 96 * 96 * 3
METRIC = "agent_frames_per_sec"
UNIT = "agent-frames/s"
NUM_AGENTS = 2
BATCH_ENVS = 1024


def workload_config(A, B, world, obs_format="rgb", extra=None, ego=False):
    cfg = _workload_config(A, B, world, obs_format, ego)
    if extra:
        cfg.update(extra)
    return cfg


def _workload_config(A, B, world, obs_format="rgb", ego=False):
    return {"workload": "MultiCarRacing-v0 step+render, num_agents=%d, batch=%d envs per GPU, random policy, "
                        "use_random_direction=True, device-side next-step auto reset (done or 1000 steps)%s" % (
                            A, B, ("" if obs_format == "rgb" else ", obs_format=%s (NOT the reference's layout)" % obs_format) +
                            (", use_ego_color=True" if ego else "")),
            "batch_envs_per_gpu": B, "num_agents": A, "l2": "256 MiB flush between timed steps",
            "parallelism": "env-sharded x%d, no data-path collective" % world}


# --------------------------------------------------------------------------------------------
# CPU arm: the oracle port (bench.py is one of the few places allowed to execute oracle/)
# --------------------------------------------------------------------------------------------
def _cpu_worker(args):
    seed, n_envs, num_agents, steps, budget_s = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import mcr_oracle as mo
    worlds = []
    rs = np.random.RandomState(seed)
    for e in range(n_envs):
        tr, _ = mo.generate_track(np.random.RandomState(seed * 1000 + e))
        w = mo.OracleWorld(num_agents)
        cw = rs.uniform() < 0.5
        w.set_track(tr, cw)
        order = rs.permutation(num_agents)
        w.spawn(mo.spawn_poses([tuple(r) for r in tr.nodes], {i: order[i] for i in range(num_agents)}, 'CW' if cw else 'CCW'))
        w.step(None)
        worlds.append(w)
    acts = np.empty((steps, n_envs, num_agents, 3))
    acts[..., 0] = rs.uniform(-1, 1, acts.shape[:-1])
    acts[..., 1] = rs.uniform(0, 1, acts.shape[:-1])
    acts[..., 2] = rs.uniform(0, 1, acts.shape[:-1])
    t0 = time.perf_counter()
    done_steps = 0
    for s in range(steps):
        for e, w in enumerate(worlds):
            w.step(acts[s, e])
        done_steps += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return done_steps * n_envs * num_agents, dt


def cpu_port_throughput(procs, n_envs_per_proc, steps, budget_s, num_agents=NUM_AGENTS):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mcr_oracle as mo
    mo.build()
    if procs == 1:
        frames, dt = _cpu_worker((1, n_envs_per_proc, num_agents, steps, budget_s))
        return frames / dt, frames, dt
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, [(i + 1, n_envs_per_proc, num_agents, steps, budget_s) for i in range(procs)])
        wall = time.perf_counter() - t0
    frames = sum(r[0] for r in res)
    slowest = max(r[1] for r in res)
    return frames / slowest, frames, wall


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    steps = max(1, int(args.steps))
    warm = max(0, int(args.warmup))
    envs_per_proc = 4
    # bounded sample: each "step" of this arm advances cores x 4 envs x 2 agents by one env step
    budget = float(os.environ.get("MCR_REF_BUDGET_S", "60"))
    if warm:
        cpu_port_throughput(cores, envs_per_proc, warm, budget)
    value, frames, wall = cpu_port_throughput(cores, envs_per_proc, steps, budget)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": int(args.gpus), "steps": steps,
        "warmup": warm, "ms_per_step": 1e3 * wall / max(1, frames // (cores * envs_per_proc * NUM_AGENTS)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
        "config": workload_config(NUM_AGENTS, BATCH_ENVS, int(args.gpus), extra={
            "reference_arm_sample": "a step of THIS arm advances %d envs (%d processes x %d), not the %d of the workload: "
                                    "a bounded sample of the same per-env work, throughput in the same unit" % (
                                        cores * envs_per_proc, cores, envs_per_proc, BATCH_ENVS),
            "envs_stepped_per_step": cores * envs_per_proc}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d agent-frames: %d processes x %d envs x %d agents, oracle/mcr_oracle.c (CPU port of the "
                                   "reference algorithm; the original Box2D+pyglet stack is not installable offline)" % (
                                       frames, cores, envs_per_proc, NUM_AGENTS)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------------
# clocks sampler
# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw"

    def __init__(self, index):
        self.samples, self.proc, self.index = [], None, index
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def wait_first(self, timeout):
        t0 = time.perf_counter()
        while self.proc is not None and not self.samples and time.perf_counter() - t0 < timeout:
            time.sleep(0.02)

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [l for (t, l) in self.samples if t0 <= t <= t1] or [l for (_, l) in self.samples]
        for l in rows:
            f = [x.strip() for x in l.split(",")]
            try:
                sm.append(float(f[0])); smax = float(f[1])
            except Exception:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(rows)}


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from multi_car_racing_b200 import build as mcr_build

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if rank == 0:
        mcr_build.build()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    import multi_car_racing_b200 as mcr
    from multi_car_racing_b200.dist import bind_to_gpu_numa_node
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    numa_node = bind_to_gpu_numa_node(local_rank)     # before any pinned allocation
    K, W = max(1, int(args.steps)), max(3, int(args.warmup))
    B, A = int(args.batch_envs), int(args.num_agents)
    frames_per_step = B * A

    np.random.seed(1234 + rank)
    venv = mcr.BatchedMultiCarRacing(B, num_agents=A, use_random_direction=True, device=dev, auto_reset='next_step',
                                     max_episode_steps=1000, seed=1234 + rank * B, obs_format=args.obs_format,
                                     use_ego_color=bool(args.use_ego_color))
    obs_bytes = int(np.prod(venv.obs_shape)) * (2 if args.obs_format == "rgb_chw_f16" else 1)
    venv.reset(device_tracks=True)
    gen = torch.Generator(device=dev); gen.manual_seed(1234 + rank)
    TAPE = 128
    tape = torch.rand((TAPE, B, A, 3), device=dev, generator=gen)
    tape[..., 0] = tape[..., 0] * 2 - 1          # steer in [-1, 1], gas/brake in [0, 1]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident timed region --------------------------------------------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.wait_first(3.0)
    for s in range(W):
        venv.step(tape[s % TAPE])
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    launches0 = venv.launch_count
    t_wall0 = time.perf_counter()
    for s in range(K):
        flush.zero_()                              # evict L2 between timed iterations (not timed)
        ev[s][0].record()
        venv.step(tape[(W + s) % TAPE])
        ev[s][1].record()
    barrier()
    t_wall1 = time.perf_counter()
    launches = venv.launch_count - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    status = venv.status().tolist()

    # ---- per-kernel split (same work, no auto reset) for the roofline: simulate, then the rasteriser ALONE
    #      (mcr_render without the reward / done block that rides beside it in the real step) -------------------
    KS = min(K, 200)
    kev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(KS)]
    tl = venv.buffers["timeline"].view(torch.int64)
    fill_us = []
    for s in range(KS):
        flush.zero_()
        tl.zero_()
        kev[s][0].record()
        venv.simulate_only(tape[s % TAPE])
        kev[s][1].record()
        venv.render_only()
        kev[s][2].record()
        if s < 32:                                 # fill_kernel's own window: its first CTA's start .. its last CTA's end
            torch.cuda.synchronize(dev)
            t = tl.cpu().numpy()
            if t[15] > 0 and t[8] > t[15]:
                fill_us.append((t[8] - t[15]) / 1e3)
    torch.cuda.synchronize(dev)
    k_ms = [sum(e[i].elapsed_time(e[i + 1]) for e in kev) / KS for i in range(2)]
    fill_ms = (sum(fill_us) / len(fill_us) / 1e3) if fill_us else None

    # ---- end-to-end through the host-buffer API ----------------------------------------------------
    KE = min(K, 200)
    e2e_s = pipe_s = float("nan")
    if not args.no_e2e:
      hb = venv.host_buffers()
      host_tape = tape[:16].cpu().numpy()
      for s in range(3):
          venv.step_host(host_tape[s % 16])
      barrier()
      t0 = time.perf_counter()
      for s in range(KE):
          hb["action"].numpy()[...] = host_tape[s % 16]
          venv.step_host(hb["action"].numpy())
      barrier()
      e2e_s = time.perf_counter() - t0

      # ---- the same with the copy of step k overlapping the compute of step k+1 (not the headline: a
      #      synchronous caller cannot use it; reported as e2e_pipelined) ---------------------------------------
      for s in range(3):
          venv.step_host_async(host_tape[s % 16])
          if s:
              venv.step_host_wait()
      venv.step_host_wait()
      barrier()
      t0 = time.perf_counter()
      for s in range(KE):
          venv.step_host_async(host_tape[s % 16])
          if s:
              venv.step_host_wait()
      venv.step_host_wait()
      barrier()
      pipe_s = time.perf_counter() - t0

    # ---- reduce over ranks: MAX time, SUM frames -------------------------------------------------------
    times = torch.tensor([dev_ms, e2e_s * 1e3, pipe_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max, pipe_ms_max = times.tolist()
    total_frames = frames_per_step * K * world
    value = total_frames / (dev_ms_max * 1e-3)
    e2e_value = frames_per_step * KE * world / (e2e_ms_max * 1e-3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        render_ms = k_ms[1]
        achieved = frames_per_step * obs_bytes / (render_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None          # not measurable inside this run: taken from the committed ncu capture
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "render_traffic.json")))
            traffic = tj.get("dram_bytes_per_launch")
            traffic_src = "static: profiles/render_traffic.json (%s)" % tj.get("source", "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum")
        except Exception:
            pass
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, frames, wall = cpu_port_throughput(1, 4, 2000, 12.0)
            cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": "%d agent-frames: 4 envs x 2 agents, single thread, oracle/mcr_oracle.c (CPU port of the "
                             "reference algorithm; original Box2D+pyglet not installable offline)" % frames}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 rigid bodies + f64 tyre/reward (u8 frames)", "data": "synthetic",
            "config": workload_config(A, B, world, args.obs_format, ego=bool(args.use_ego_color)),
            "clocks": clocks,
            "e2e": None if args.no_e2e else {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * A * 3 * 4,
                    "d2h_bytes_per_step": B * A * obs_bytes + B * A * 8 + B, "steps": KE},
            "e2e_pipelined": None if args.no_e2e else {"value": frames_per_step * KE * world / (pipe_ms_max * 1e-3), "unit": UNIT, "steps": KE,
                              "note": "step_host_async / step_host_wait: the device-to-host copy of step k overlaps the "
                                      "compute of step k+1 (results one call late); not usable by a synchronous policy loop"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "rasteriser = project_kernel + fill_kernel (fill_kernel writes the frames)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_kind,
                         "algorithmic_bytes_per_launch": frames_per_step * obs_bytes,
                         "kernel_ms": {"simulate": k_ms[0], "render": k_ms[1], "fill": fill_ms},
                         "fill_kernel_alone": None if not fill_ms else {
                             "achieved": frames_per_step * obs_bytes / (fill_ms * 1e-3) / 1e9,
                             "frac": frames_per_step * obs_bytes / (fill_ms * 1e-3) / 1e9 / peak}},
            "cpu_baseline": cpu,
            "status_words": status,
            "host": {"numa_node_rank0": numa_node, "cpus_rank0": len(os.sched_getaffinity(0))},
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch-envs", dest="batch_envs", type=int, default=BATCH_ENVS)
    ap.add_argument("--num-agents", dest="num_agents", type=int, default=NUM_AGENTS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", dest="no_e2e", action="store_true", help="skip the host-buffer legs (large-batch sweeps)")
    ap.add_argument("--use-ego-color", dest="use_ego_color", action="store_true", help="BASELINE.json configs[3]")
    ap.add_argument("--obs-format", dest="obs_format", default="rgb", choices=["rgb", "gray", "rgb_chw", "gray_stack", "rgb_chw_f16"],
                    help="rasteriser store layout; only 'rgb' is the reference's observation (the headline config)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
