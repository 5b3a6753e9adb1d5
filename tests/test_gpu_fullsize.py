"""Parity at BASELINE.json's full batch size (configs[1]: 1024 envs x 2 agents) through
size-independent properties: (1) a sample of envs spread over the batch is bit-exact against the
oracle stepping on the same tracks; (2) an env's trajectory does not depend on what else is in the
batch (the 1024-env run and a 6-env run of the same envs agree bit for bit); (3) two runs from the
same seeds give identical checksums of all frames, rewards and state (determinism, no racing
atomics); (4) configs[3] (8 agents, use_ego_color) at 64 envs against the oracle."""
import hashlib

import numpy as np
import pytest

from helpers import action_tape, make_oracle_worlds, gpu_state, oracle_state

pytestmark = pytest.mark.gpu


def _run(mcr, B, A, steps, tracks, orders, directions, tape, sample, **kw):
    import torch
    venv = mcr.BatchedMultiCarRacing(B, num_agents=A, auto_reset=False, max_episode_steps=0, **kw)
    obs = venv.reset(tracks=tracks, car_orders=orders, directions=directions)
    frames = [obs[sample].cpu().numpy()]
    rewards = []
    sha = hashlib.sha1()
    dev_tape = torch.from_numpy(tape).to(venv.device)
    for s in range(steps):
        obs, rew, done, _ = venv.step(dev_tape[s])
        frames.append(obs[sample].cpu().numpy())
        rewards.append(rew[sample].cpu().numpy())
        sha.update(obs.cpu().numpy().tobytes()); sha.update(rew.cpu().numpy().tobytes()); sha.update(done.cpu().numpy().tobytes())
    sha.update(venv.buffers["body"].cpu().numpy().tobytes())
    assert not venv.status().any()
    return venv, np.stack(frames), np.stack(rewards), sha.hexdigest()


def test_full_batch_sample_parity_independence_determinism(oracle, mcr):
    B, A, steps = 1024, 2, 60
    n_unique = 16                                  # 16 distinct tracks, tiled over the batch
    uniq = [oracle.generate_track(np.random.RandomState(4000 + i))[0] for i in range(n_unique)]
    rs = np.random.RandomState(9)
    tracks = [uniq[e % n_unique] for e in range(B)]
    orders = [rs.permutation(A) for _ in range(B)]
    directions = ['CW' if rs.uniform() < 0.5 else 'CCW' for _ in range(B)]
    tape = action_tape(31, steps, B, A)
    sample = np.array([0, 1, 337, 512, 1000, 1023])
    venv, frames, rewards, digest = _run(mcr, B, A, steps, tracks, orders, directions, tape, sample)
    # (1) oracle parity of the sampled envs, every step
    worlds = make_oracle_worlds(oracle, [tracks[e] for e in sample], [orders[e] for e in sample], [directions[e] for e in sample], A)
    o0 = np.stack([w.step(None)[0] for w in worlds])
    assert np.array_equal(frames[0], o0)
    for s in range(steps):
        oo = [w.step(tape[s, e].astype(np.float64)) for e, w in zip(sample, worlds)]
        assert np.array_equal(rewards[s], np.stack([x[1] for x in oo])), "rewards, step %d" % s
        assert np.array_equal(frames[s + 1], np.stack([x[0] for x in oo])), "pixels, step %d" % s
    g, o = gpu_state(venv), oracle_state(worlds)
    assert np.array_equal(g["bodies"][sample], o["bodies"])
    # (2) independence of the batch composition
    _, frames6, rewards6, _ = _run(mcr, len(sample), A, steps, [tracks[e] for e in sample], [orders[e] for e in sample],
                                   [directions[e] for e in sample], np.ascontiguousarray(tape[:, sample]), np.arange(len(sample)))
    assert np.array_equal(frames, frames6) and np.array_equal(rewards, rewards6)
    # (3) determinism of the whole batch
    del venv
    _, _, _, digest2 = _run(mcr, B, A, steps, tracks, orders, directions, tape, sample)
    assert digest == digest2


def test_config_8_agents_ego_color(oracle, mcr):
    """BASELINE.json configs[3]: num_agents=8, use_ego_color=True (spawn rows 0-3, 8 cars per view)."""
    B, A, steps = 64, 8, 40
    uniq = [oracle.generate_track(np.random.RandomState(4100 + i))[0] for i in range(4)]
    rs = np.random.RandomState(10)
    tracks = [uniq[e % 4] for e in range(B)]
    orders = [rs.permutation(A) for _ in range(B)]
    directions = ['CW' if rs.uniform() < 0.5 else 'CCW' for _ in range(B)]
    tape = action_tape(32, steps, B, A)
    sample = np.array([0, 21, 63])
    venv, frames, rewards, _ = _run(mcr, B, A, steps, tracks, orders, directions, tape, sample, use_ego_color=True)
    worlds = make_oracle_worlds(oracle, [tracks[e] for e in sample], [orders[e] for e in sample], [directions[e] for e in sample], A,
                                use_ego_color=True)
    assert np.array_equal(frames[0], np.stack([w.step(None)[0] for w in worlds]))
    for s in range(steps):
        oo = [w.step(tape[s, e].astype(np.float64)) for e, w in zip(sample, worlds)]
        assert np.array_equal(rewards[s], np.stack([x[1] for x in oo])), "rewards, step %d" % s
        assert np.array_equal(frames[s + 1], np.stack([x[0] for x in oo])), "pixels, step %d" % s
