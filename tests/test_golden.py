"""Golden fixtures produced by the UNMODIFIED reference class running on stub third-party
modules (tests/golden/make_golden.py, tests/golden/refstub.py).  They pin the restatement of
the reference's own Python -- FrictionDetector reward rule, reset() RNG order and spawn grid,
_create_track, step()'s reward/backward/done block, camera maths, draw lists, HUD geometry,
read-back flip -- in the CPU oracle (CPU suite) and in the CUDA path (GPU suite)."""
import ast
import glob
import hashlib
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")) if "track_kat" not in p)


def _load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return z, ast.literal_eval(str(z["kwargs"]))


def _replay(env_cls, z, kwargs, check_state):
    """Drive env_cls exactly like make_golden.run_case drove the reference."""
    np.random.seed(int(z["np_seed"]))
    env = env_cls(verbose=0, **kwargs)
    env.seed(int(z["env_seed"]))
    acts, steps = z["actions"].astype(np.float64), int(z["steps_per_episode"])
    frame_at = {int(i): k for k, i in enumerate(z["frame_idx"])}
    n = 0
    for ep in range(int(z["episodes"])):
        obs = env.reset()
        assert len(env.track) == int(z["track_len"][ep]), "track length"
        assert str(env.episode_direction) == str(z["direction"][ep]), "episode direction (global numpy RNG order)"
        assert [int(env.car_order[i]) for i in range(kwargs["num_agents"])] == list(z["car_order"][ep]), "car order"
        assert np.array_equal(obs, z["reset_frames"][ep]), "first observation of episode %d" % ep
        for s in range(steps):
            if n >= len(z["rewards"]):
                break
            obs, rew, done, info = env.step(acts[s])
            assert np.array_equal(np.asarray(rew, np.float64), z["rewards"][n]), "step reward at step %d" % n
            assert bool(done) == bool(z["dones"][n]), "done at step %d" % n
            if check_state:
                assert np.array_equal(np.asarray(env.reward, np.float64), z["env_reward"][n]), "env.reward at %d" % n
                assert [int(v) for v in env.tile_visited_count] == list(z["counts"][n]), "tile_visited_count at %d" % n
                assert np.array_equal(np.asarray(env.driving_backward, np.uint8), z["backward"][n]), "driving_backward at %d" % n
                assert np.array_equal(np.asarray(env.driving_on_grass, np.uint8), z["grass"][n]), "driving_on_grass at %d" % n
            if n in frame_at:
                assert np.array_equal(obs, z["frames"][frame_at[n]]), "observation pixels at step %d" % n
            assert hashlib.sha1(np.ascontiguousarray(obs).tobytes()).digest() == z["sha"][n].tobytes(), "observation hash at step %d" % n
            n += 1
            if done:
                break
    assert n == len(z["rewards"])
    return env


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_run(oracle, name):
    z, kwargs = _load(name)
    _replay(oracle.OracleMultiCarRacing, z, kwargs, check_state=True)


def test_oracle_track_generator_known_answers(oracle):
    z = np.load(os.path.join(GOLDEN, "track_kat.npz"))
    for seed in range(8):
        rng = np.random.RandomState(seed)
        tr, attempts = oracle.generate_track(rng)
        assert attempts == int(z["attempts_%d" % seed])
        assert np.array_equal(tr.nodes, z["nodes_%d" % seed])
        assert np.array_equal(tr.quads, z["quads_%d" % seed])
        assert np.allclose(tr.quad_rgb.astype(np.float64), z["rgb_%d" % seed], atol=1e-7)


def test_native_track_generator_known_answers(mcr):
    """The product's C++ generator (mcr_track_generate) against the reference's _create_track."""
    from multi_car_racing_b200.track import TrackGenerator
    z = np.load(os.path.join(GOLDEN, "track_kat.npz"))
    gen = TrackGenerator()
    for seed in range(8):
        rng = np.random.RandomState(seed)
        tr = gen.generate(rng)
        assert tr.attempts == int(z["attempts_%d" % seed])
        assert np.array_equal(tr.nodes, z["nodes_%d" % seed])
        assert np.array_equal(tr.quads, z["quads_%d" % seed])
        assert np.allclose(tr.quad_rgb.astype(np.float64), z["rgb_%d" % seed], atol=1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_reproduces_reference_run(mcr, name):
    z, kwargs = _load(name)
    _replay(mcr.MultiCarRacing, z, kwargs, check_state=True)
