"""Property tests from SURVEY.md section 4 (the reference has no tests of its own): invariants of FrictionDetector's
visit bookkeeping (mcr:88-123) and of the HUD (mcr:634-674), checked on the CPU oracle with hypothesis-drawn tracks,
agent counts and action tapes, and on the CUDA path at a larger batch (GPU-marked)."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from helpers import action_tape, visited_bits

HUD_COLOURS = {(0, 0, 0), (255, 255, 255), (0, 0, 255), (51, 0, 255), (0, 255, 0), (255, 0, 0)}


def _tile_reward_total(visited, A, T):
    """What the visits of a (T, A) flag matrix are worth in total: the j-th visitor of a tile gets (1 - j/A) * 1000/T."""
    k = visited.sum(axis=1).astype(np.int64)
    per_tile = np.array([sum(1.0 - j / A for j in range(n)) for n in range(A + 1)])
    return float(per_tile[k].sum() * 1000.0 / T)


def _check_hud(obs):
    """rows 84..95 of every frame are the HUD bar: black except for the indicator quads, the score glyphs (white) and the flag"""
    hud = obs[..., 84:, :, :].reshape(-1, 3)
    cols = {tuple(c) for c in np.unique(hud, axis=0)}
    assert cols <= HUD_COLOURS, "world colours leak into the HUD rows: %s" % sorted(cols - HUD_COLOURS)


@settings(max_examples=12, deadline=None)
@given(seed=st.integers(0, 10 ** 6), A=st.sampled_from([1, 2, 3, 4]), steps=st.integers(20, 160), cw=st.booleans(),
       brake_p=st.sampled_from([0.0, 0.1, 0.5]))
def test_visit_bookkeeping_invariants_oracle(oracle, seed, A, steps, cw, brake_p):
    tr, _ = oracle.generate_track(np.random.RandomState(seed))
    w = oracle.OracleWorld(A)
    w.set_track(tr, cw)
    order = np.random.RandomState(seed + 1).permutation(A)
    w.spawn(oracle.spawn_poses([tuple(r) for r in tr.nodes], {i: order[i] for i in range(A)}, 'CW' if cw else 'CCW'))
    obs0 = w.step(None)[0]
    _check_hud(obs0)
    tape = action_tape(seed % 1000, steps, 1, A, brake_p=brake_p)
    prev_flags = np.array(w.visited()[0]).copy()
    step_sum = np.zeros(A)
    n_real = 0
    for s in range(steps):
        obs, rew, done = w.step(tape[s, 0].astype(np.float64))
        n_real += 1
        step_sum += rew
        flags = np.array(w.visited()[0])
        assert (flags >= prev_flags).all(), "a visit flag never clears within an episode"
        prev_flags = flags.copy()
        if s % 20 == 0:
            _check_hud(obs)
        if done:
            break
    reward, counts, _ = w.scores()
    reward, counts = np.array(reward), np.array(counts)
    assert np.array_equal(counts, flags.sum(axis=0)), "an agent is counted (and rewarded) once per tile"
    assert (counts <= tr.T).all()
    # conservation: env.reward summed over agents = what the visit flags are worth - 0.1 per agent per step
    assert abs(reward.sum() - (_tile_reward_total(flags, A, tr.T) - 0.1 * A * n_real)) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("A,B", [(2, 64), (8, 16)])
def test_visit_bookkeeping_invariants_cuda(mcr, A, B):
    import torch
    np.random.seed(11)
    venv = mcr.BatchedMultiCarRacing(B, num_agents=A, auto_reset=False, max_episode_steps=0, seed=4242)
    obs = venv.reset(device_tracks=True)
    _check_hud(obs.cpu().numpy())
    steps = 300
    tape = action_tape(8, steps, B, A, brake_p=0.1)
    alive = np.ones(B, bool)
    n_real = np.zeros(B, np.int64)
    for s in range(steps):
        obs, rew, done, _ = venv.step(torch.from_numpy(tape[s]).to(venv.device))
        n_real += alive
        alive &= (done.cpu().numpy() & 1) == 0
        if s % 50 == 0:
            _check_hud(obs.cpu().numpy())
        if s == 149:
            mid_flags = venv.buffers["visited"].cpu().numpy().copy()
    vis = venv.buffers["visited"].cpu().numpy()
    assert ((vis & mid_flags) == mid_flags).all(), "visit flags only ever get set"
    reward = venv.buffers["reward"].cpu().numpy().reshape(B, A)
    counts = venv.buffers["visit_count"].cpu().numpy().reshape(B, A)
    T = venv.buffers["trk_T"].cpu().numpy()
    for e in range(B):
        flags = visited_bits(vis[e], int(T[e]), A)
        assert np.array_equal(counts[e], flags.sum(axis=0)) and (counts[e] <= T[e]).all()
        # steps after `done` keep costing 0.1 (no auto reset here): every step call counts
        assert abs(reward[e].sum() - (_tile_reward_total(flags, A, int(T[e])) - 0.1 * A * steps)) < 1e-6, "env %d" % e
    assert not venv.status().any()
