#!/usr/bin/env python
"""Generate tests/golden/*.npz by executing the UNMODIFIED reference class on stub third-party
modules (tests/golden/refstub.py).  Run in the build container only:

    python tests/golden/make_golden.py            # needs /root/reference

Each fixture records one seeded episode: constructor kwargs, both RNG seeds, the action tape,
and what the reference returned / held after every step -- step rewards, done, env.reward,
tile_visited_count, driving_backward, driving_on_grass, hull poses, a SHA-1 of every observation and every 20th
observation in full.  tests/test_golden.py replays the tapes through the CPU oracle (CPU suite)
and through the CUDA path (GPU suite).

Also writes track_kat.npz: known answers of the reference's _create_track for plain
RandomState(seed) streams (the AST-extracted method executed with the stubs).
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)
REF = os.environ.get("MCR_REFERENCE", "/root/reference")


def load_reference():
    import refstub
    refstub.install()
    sys.path.insert(0, REF)
    import gym_multi_car_racing.multi_car_racing as ref
    return ref


def tape(kind, seed, steps, A):
    rs = np.random.RandomState(seed)
    a = np.zeros((steps, A, 3), np.float64)
    if kind == "random":
        a[..., 0] = rs.uniform(-1, 1, (steps, A))
        a[..., 1] = rs.uniform(0, 1, (steps, A))
        a[..., 2] = np.where(rs.uniform(0, 1, (steps, A)) < 0.25, rs.uniform(0, 1, (steps, A)), 0.0)
    elif kind == "drive":          # mostly sensible driving: small steering noise, steady gas
        a[..., 0] = rs.uniform(-0.3, 0.3, (steps, A))
        a[..., 1] = 0.4 + 0.2 * rs.uniform(0, 1, (steps, A))
    elif kind == "flatout":        # straight off the track and out of the playfield: done + -100
        a[..., 1] = 0.3
        a[..., 0] = 0.0
    elif kind == "spin":           # full throttle with steering: spins, drives backwards (flag triangle in the HUD)
        a[..., 1] = 1.0
        a[..., 0] = 0.02
    return a.astype(np.float32).astype(np.float64)   # exactly representable as float32 (the GPU ABI's f32 path)


CASES = [
    dict(name="default_a2", kwargs=dict(num_agents=2), np_seed=1, env_seed=11, kind="random", steps=300),
    dict(name="carracing_a1", kwargs=dict(num_agents=1, use_random_direction=False, backwards_flag=False),
         np_seed=2, env_seed=22, kind="drive", steps=300),
    dict(name="ego_a4", kwargs=dict(num_agents=4, use_ego_color=True, h_ratio=0.3, direction='CW', use_random_direction=False),
         np_seed=3, env_seed=33, kind="random", steps=150),
    dict(name="outfield_a2", kwargs=dict(num_agents=2), np_seed=4, env_seed=44, kind="flatout", steps=400),
    dict(name="spin_a2", kwargs=dict(num_agents=2, backwards_flag=True), np_seed=6, env_seed=66, kind="spin", steps=250),
    dict(name="two_episodes_a2", kwargs=dict(num_agents=2), np_seed=5, env_seed=55, kind="drive", steps=120, episodes=2),
]


def run_case(ref, case):
    A = case["kwargs"]["num_agents"]
    np.random.seed(case["np_seed"])
    env = ref.MultiCarRacing(verbose=0, **case["kwargs"])
    env.seed(case["env_seed"])
    out = dict(rewards=[], dones=[], env_reward=[], counts=[], backward=[], grass=[], poses=[], sha=[], frames=[], frame_idx=[],
               reset_sha=[], reset_frames=[], track_len=[], direction=[], car_order=[])
    acts = tape(case["kind"], case["env_seed"], case["steps"], A)
    for ep in range(case.get("episodes", 1)):
        obs = env.reset()
        out["reset_sha"].append(hashlib.sha1(np.ascontiguousarray(obs).tobytes()).digest())
        out["reset_frames"].append(np.ascontiguousarray(obs))
        out["track_len"].append(len(env.track))
        out["direction"].append(env.episode_direction)
        out["car_order"].append([int(env.car_order[i]) for i in range(A)])
        for s in range(case["steps"]):
            obs, rew, done, info = env.step(acts[s])
            assert info == {}
            obs = np.ascontiguousarray(obs)
            out["rewards"].append(np.array(rew, np.float64))
            out["dones"].append(bool(done))
            out["env_reward"].append(np.array(env.reward, np.float64))
            out["counts"].append(np.array(env.tile_visited_count, np.int32))
            out["backward"].append(np.array(env.driving_backward, np.uint8))
            out["grass"].append(np.array(env.driving_on_grass, np.uint8))
            out["poses"].append(np.array([[c.hull.position[0], c.hull.position[1], c.hull.angle] for c in env.cars], np.float32))
            out["sha"].append(hashlib.sha1(obs.tobytes()).digest())
            if s % 20 == 0:
                out["frames"].append(obs)
                out["frame_idx"].append(len(out["sha"]) - 1)
            if done:
                break
    n = len(out["rewards"])
    np.savez_compressed(
        os.path.join(HERE, case["name"] + ".npz"),
        kwargs=np.array(repr(case["kwargs"])), np_seed=case["np_seed"], env_seed=case["env_seed"],
        episodes=case.get("episodes", 1), steps_per_episode=case["steps"], actions=acts.astype(np.float32),
        rewards=np.array(out["rewards"]), dones=np.array(out["dones"]), env_reward=np.array(out["env_reward"]),
        counts=np.array(out["counts"]), backward=np.array(out["backward"]), grass=np.array(out["grass"]), poses=np.array(out["poses"]),
        sha=np.frombuffer(b"".join(out["sha"]), np.uint8).reshape(n, 20),
        frames=np.array(out["frames"]), frame_idx=np.array(out["frame_idx"], np.int32),
        reset_sha=np.frombuffer(b"".join(out["reset_sha"]), np.uint8).reshape(-1, 20),
        reset_frames=np.array(out["reset_frames"]), track_len=np.array(out["track_len"], np.int32),
        direction=np.array(out["direction"]), car_order=np.array(out["car_order"], np.int32))
    print("%-18s steps=%d done=%s counts=%s reward=%s backward_frames=%d" % (
        case["name"], n, out["dones"][-1], out["counts"][-1], np.round(out["env_reward"][-1], 3),
        int(np.array(out["backward"]).sum())))


def track_kats(ref):
    """_create_track known answers for RandomState(seed), seeds 0..7."""
    import Box2D
    rows = {}
    for seed in range(8):
        env = ref.MultiCarRacing.__new__(ref.MultiCarRacing)
        env.np_random = np.random.RandomState(seed)
        env.num_agents, env.verbose = 2, 0
        env.world = Box2D.b2World((0, 0))
        env.fd_tile = ref.fixtureDef(shape=ref.polygonShape(vertices=[(0, 0), (1, 0), (1, -1), (0, -1)]))
        attempts = 0
        while True:
            attempts += 1
            env.road_poly = []
            if env._create_track():
                break
        rows["nodes_%d" % seed] = np.array(env.track, np.float64)
        rows["quads_%d" % seed] = np.array([p for p, c in env.road_poly], np.float64)
        rows["rgb_%d" % seed] = np.array([c for p, c in env.road_poly], np.float64)
        rows["attempts_%d" % seed] = attempts
    np.savez_compressed(os.path.join(HERE, "track_kat.npz"), **rows)
    print("track_kat: T =", [len(rows["nodes_%d" % s]) for s in range(8)])


if __name__ == "__main__":
    ref = load_reference()
    track_kats(ref)
    for case in CASES:
        run_case(ref, case)
