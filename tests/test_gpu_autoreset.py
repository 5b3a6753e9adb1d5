"""Device-side auto reset (both conventions) against the CPU oracle.

The device picks the next track slot, direction and car order itself (counter-based RNG); the
test reads that choice back (env_track / env_cw, car order recovered from the spawn positions),
builds a fresh oracle world for it and demands bit-exact state, rewards and pixels afterwards.
Reference behaviour being reproduced: reset() = respawn + step(None) (mcr:340-408), TimeLimit at
max_episode_steps (reference __init__.py:5-10)."""
import hashlib
import itertools

import numpy as np
import pytest

from helpers import action_tape, make_oracle_worlds, gpu_state, oracle_state

pytestmark = pytest.mark.gpu


def _oracle_for_reset(oracle, venv, env, A):
    """Oracle world matching what the device respawned in `env` (after its step(None))."""
    slot = int(venv.buffers["env_track"][env].item())
    cw = bool(venv.buffers["env_cw"][env].item())
    tr = venv.tracks[slot]
    d = 'CW' if cw else 'CCW'
    g = gpu_state(venv)
    for order in itertools.permutations(range(A)):
        w = make_oracle_worlds(oracle, [tr], [np.array(order)], [d], A)[0]
        obs = w.step(None)[0]
        if np.array_equal(w.bodies(), g["bodies"][env]):
            return w, obs
    raise AssertionError("no car order reproduces the respawned env %d (slot %d, %s)" % (env, slot, d))


@pytest.mark.parametrize("mode", ["same_step", "next_step"])
def test_auto_reset_matches_oracle(oracle, mcr, mode):
    import torch
    B, A, LIMIT = 3, 2, 6
    np.random.seed(3)
    venv = mcr.BatchedMultiCarRacing(B, num_agents=A, auto_reset=mode, max_episode_steps=LIMIT, seed=11, pool_tracks=5)
    venv.reset()
    tape = action_tape(21, 3 * LIMIT + 4, B, A)
    worlds = [None] * B
    steps_in_episode = 0
    for s in range(len(tape)):
        obs, rew, done, _ = venv.step(torch.from_numpy(tape[s]).to(venv.device))
        obs, rew, done = obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy()
        steps_in_episode += 1
        if mode == "same_step":
            if steps_in_episode == LIMIT:
                assert np.all(done & 2), "TimeLimit bit at step %d" % s
                for e in range(B):
                    worlds[e], oobs = _oracle_for_reset(oracle, venv, e, A)
                    assert np.array_equal(obs[e], oobs), "reset frame, env %d" % e
                steps_in_episode = 0
                continue
        else:
            if steps_in_episode == LIMIT:
                assert np.all(done & 2), "TimeLimit bit at step %d" % s
                if worlds[0] is not None:      # terminal observation stays the stepped one
                    oo = [w.step(tape[s, e].astype(np.float64)) for e, w in enumerate(worlds)]
                    assert np.array_equal(obs, np.stack([x[0] for x in oo]))
                    assert np.array_equal(rew, np.stack([x[1] for x in oo]))
                continue
            if steps_in_episode == LIMIT + 1:  # the reset step: action ignored, reward 0, done 0
                assert np.all(done == 0) and np.all(rew == 0.0)
                for e in range(B):
                    worlds[e], oobs = _oracle_for_reset(oracle, venv, e, A)
                    assert np.array_equal(obs[e], oobs), "reset frame, env %d" % e
                steps_in_episode = 0
                continue
        assert np.all(done == 0)
        if worlds[0] is None:
            continue
        oo = [w.step(tape[s, e].astype(np.float64)) for e, w in enumerate(worlds)]
        assert np.array_equal(rew, np.stack([x[1] for x in oo])), "reward, step %d" % s
        assert np.array_equal(obs, np.stack([x[0] for x in oo])), "pixels, step %d" % s
        g, o = gpu_state(venv), oracle_state(worlds)
        for k in ("bodies", "wheels", "joints", "reward", "counts"):
            assert np.array_equal(g[k], o[k]), "%s, step %d" % (k, s)
    assert not venv.status().any()


def test_out_of_field_done_then_next_step_reset(mcr):
    """done without TimeLimit: drive one env off the playfield by teleporting its hull."""
    import torch
    B, A = 2, 2
    np.random.seed(4)
    venv = mcr.BatchedMultiCarRacing(B, num_agents=A, auto_reset="next_step", max_episode_steps=0, seed=5)
    venv.reset()
    act = torch.zeros((B, A, 3), device=venv.device)
    venv.step(act)
    body = venv.buffers["body"]            # (5, 10, N): push every body of env 1 / car 0 far out in x
    car = 1 * A + 0
    for i in range(5):
        body[i, 0, car] += 1000.0
        body[i, 6, car] += 1000.0
    obs, rew, done, _ = venv.step(act)
    done = done.cpu().numpy(); rew = rew.cpu().numpy()
    assert done[1] & 1 and not done[0]
    assert rew[1, 0] == -100.0
    obs, rew, done, _ = venv.step(act)
    done = done.cpu().numpy(); rew = rew.cpu().numpy()
    assert done[1] == 0 and np.all(rew[1] == 0.0)
    x = venv.buffers["body"][0, 6, car].item()
    assert abs(x) < 400.0, "env 1 was respawned inside the playfield"
    assert int(venv.buffers["steps"][car].item()) == 0 and int(venv.buffers["steps"][0].item()) == 3


def test_vector_env_adapter(mcr):
    """gym(nasium) VectorEnv protocol over the batched env: shapes, dtypes, terminated / truncated
    split, next-step autoreset, and the same numbers as BatchedMultiCarRacing driven directly."""
    import torch
    vec = mcr.MultiCarRacingVecEnv(6, num_agents=2, seed=40, max_episode_steps=25)
    ref = mcr.BatchedMultiCarRacing(6, num_agents=2, seed=40, max_episode_steps=25, auto_reset='next_step')
    np.random.seed(5)                      # reset() draws direction / car order from the global numpy RNG (mcr:351-357)
    obs, info = vec.reset()
    np.random.seed(5)
    robs = ref.reset()
    assert info == {} and obs.shape == (6, 2, 96, 96, 3) and obs.dtype == torch.uint8 and torch.equal(obs, robs)
    assert vec.single_action_space.shape == (2, 3) and vec.single_observation_space.shape == (2, 96, 96, 3)
    assert vec.action_space.shape == (6, 2, 3) and vec.num_envs == 6
    tape = action_tape(12, 60, 6, 2)
    n_trunc = 0
    for s in range(60):
        a = torch.from_numpy(tape[s]).to(obs.device)
        if s % 2:
            vec.step_async(a)
            o, r, d, _ = vec.step_wait()
            term = trunc = None
        else:
            o, r, term, trunc, _ = vec.step(a)
            d = term | trunc
            assert term.dtype == torch.bool and trunc.dtype == torch.bool and r.shape == (6, 2) and r.dtype == torch.float64
            n_trunc += int(trunc.sum())
        ro, rr, rd, _ = ref.step(a)
        assert torch.equal(o, ro) and torch.equal(r, rr) and torch.equal(d, rd != 0)
    assert n_trunc > 0
    assert vec.render('rgb_array').shape == (6, 2, 400, 600, 3)
    with pytest.raises(RuntimeError):
        vec.step_wait()


def test_graph_replay_equals_direct_launches(mcr, monkeypatch):
    """mcr_step replays a captured CUDA graph after two eager steps; MCR_NO_GRAPH=1 (read at mcr_create)
    issues every step directly.  Same seeds -> bit-identical frames, rewards, dones and state over 80 steps
    with the next-step auto reset firing (max_episode_steps=30)."""
    import torch

    def run(no_graph):
        if no_graph:
            monkeypatch.setenv("MCR_NO_GRAPH", "1")
        else:
            monkeypatch.delenv("MCR_NO_GRAPH", raising=False)
        np.random.seed(8)
        venv = mcr.BatchedMultiCarRacing(16, num_agents=2, seed=77, max_episode_steps=30, auto_reset='next_step')
        venv.reset()
        tape = action_tape(13, 80, 16, 2)
        sha = hashlib.sha1()
        for s in range(80):
            obs, rew, done, _ = venv.step(torch.from_numpy(tape[s]).to(venv.device))
            sha.update(obs.cpu().numpy().tobytes()); sha.update(rew.cpu().numpy().tobytes()); sha.update(done.cpu().numpy().tobytes())
        sha.update(venv.buffers["body"].cpu().numpy().tobytes())
        return sha.hexdigest(), venv.launch_count

    (a, la), (b, lb) = run(False), run(True)
    assert a == b and la == lb


def test_pipelined_host_api_matches_synchronous(mcr):
    """step_host_async / step_host_wait (copy of step k under the compute of step k+1, double-buffered) returns
    exactly what step_host returns, one call later."""
    np.random.seed(4)
    a = mcr.BatchedMultiCarRacing(8, num_agents=2, seed=21, max_episode_steps=20, auto_reset='next_step')
    np.random.seed(4)
    b = mcr.BatchedMultiCarRacing(8, num_agents=2, seed=21, max_episode_steps=20, auto_reset='next_step')
    np.random.seed(6); a.reset()
    np.random.seed(6); b.reset()
    tape = action_tape(15, 50, 8, 2)
    want = []
    for s in range(50):
        o, r, d, _ = a.step_host(tape[s])
        want.append((o.copy(), r.copy(), d.copy()))
    got = []
    for s in range(50):
        b.step_host_async(tape[s])
        if s:
            o, r, d, _ = b.step_host_wait()
            got.append((o.copy(), r.copy(), d.copy()))
    o, r, d, _ = b.step_host_wait()
    got.append((o.copy(), r.copy(), d.copy()))
    with pytest.raises(RuntimeError):
        b.step_host_wait()
    assert len(got) == 50
    for s, (w, g) in enumerate(zip(want, got)):
        assert all(np.array_equal(x, y) for x, y in zip(w, g)), "step %d" % s


def test_pipelined_host_api_with_a_busy_stream(mcr):
    """two step_host_async() calls issued while the GPU is still busy must not share one pinned
    action buffer (step k would run with step k+1's action).  A long kernel is queued first so that both
    host -> device action copies are still pending when the second call writes its action."""
    import torch
    np.random.seed(9)
    a = mcr.BatchedMultiCarRacing(4, num_agents=2, seed=5, max_episode_steps=0, auto_reset=False)
    np.random.seed(9)
    b = mcr.BatchedMultiCarRacing(4, num_agents=2, seed=5, max_episode_steps=0, auto_reset=False)
    np.random.seed(3); a.reset()
    np.random.seed(3); b.reset()
    tape = action_tape(33, 6, 4, 2)
    want = []
    for s in range(6):
        o, r, d, _ = a.step_host(tape[s])
        want.append((o.copy(), r.copy(), d.copy()))
    got = []
    for s in range(0, 6, 2):
        torch.cuda._sleep(200_000_000)                 # ~0.1 s of GPU time in front of the two steps
        b.step_host_async(tape[s])
        b.step_host_async(tape[s + 1])
        for _ in range(2):
            o, r, d, _ = b.step_host_wait()
            got.append((o.copy(), r.copy(), d.copy()))
    for s, (w, g) in enumerate(zip(want, got)):
        assert all(np.array_equal(x, y) for x, y in zip(w, g)), "step %d" % s


def test_vecenv_truncated_excludes_terminated(mcr):
    """gym's TimeLimit sets truncated = not done -- an env that terminates on its last allowed step
    is terminated only."""
    import torch
    np.random.seed(1)
    v = mcr.MultiCarRacingVecEnv(4, num_agents=2, seed=3, max_episode_steps=5)
    v.reset()
    act = torch.zeros((4, 2, 3), device=v.venv.device)
    for s in range(5):
        obs, rew, term, trunc, _ = v.step(act)
    assert trunc.all() and not term.any()
    # force a termination on the step that also hits the limit: place env 0's first car outside the playfield
    v.step(act)                                           # next-step auto reset: the new episodes start here
    for s in range(4):
        v.step(act)
    body = v.venv.buffers["body"]
    body[0, 6, 0] = 400.0                                 # hull origin x of car 0 (env 0): |x| > PLAYFIELD
    body[0, 0, 0] = 400.0
    obs, rew, term, trunc, _ = v.step(act)
    assert bool(term[0]) and not bool(trunc[0]), "terminated on the limit step must not also be truncated"
    assert trunc[1:].all() and not term[1:].any()


@pytest.mark.parametrize("mode,device_tracks", [("next_step", True), ("same_step", True), ("next_step", False)])
def test_fresh_tracks_follow_each_envs_own_stream(mcr, mode, device_tracks):
    """SURVEY 8f #1 / reference :359-364: every reset() generates a NEW track on the env's own np_random stream.
    Device-side auto reset must do the same: the k-th episode of env i runs on the k-th successful track of the
    host generator seeded like env.seed(SEED + i) -- T, Q, border pattern and palette identical, float64 nodes
    within 1e-9 (CUDA libm vs glibc), whatever the timing (episodes of 4 steps are far shorter than a track
    generation, so the in-place generation path is exercised too)."""
    import torch
    from multi_car_racing_b200.track import TrackGenerator, np_random
    B, A, LIMIT, SEED, EPISODES = 8, 2, 4, 321, 4
    venv = mcr.BatchedMultiCarRacing(B, num_agents=A, auto_reset=mode, max_episode_steps=LIMIT, seed=SEED)
    assert venv.fresh_tracks == 1 and venv.pool_tracks == 2 * B
    np.random.seed(0)
    venv.reset(device_tracks=device_tracks)
    gen = TrackGenerator()
    host = []
    for i in range(B):
        rng, _ = np_random(SEED + i)
        host.append([gen.generate(rng) for _ in range(EPISODES)])

    def check(ep):
        slots = venv.buffers["env_track"].cpu().numpy()
        buf = {k: venv.buffers[k].cpu().numpy() for k in ("trk_T", "trk_Q", "trk_quad_tile", "trk_quad_col", "trk_node", "trk_quad")}
        for e in range(B):
            s, tr = int(slots[e]), host[e][ep]
            assert s == e + B * (ep % 2), "env %d episode %d runs on slot %d" % (e, ep, s)
            assert buf["trk_T"][s] == tr.T and buf["trk_Q"][s] == tr.Q, "env %d episode %d: T/Q" % (e, ep)
            assert np.array_equal(buf["trk_quad_tile"][s, :tr.Q], tr.quad_tile), "border pattern, env %d episode %d" % (e, ep)
            assert np.abs(buf["trk_node"][s, :tr.T] - tr.nodes[:, 1:4]).max() <= 1e-9
            assert np.abs(buf["trk_quad"][s, :tr.Q].reshape(tr.Q, 4, 2) - tr.quads).max() <= 1e-4

    check(0)
    act = torch.zeros((B, A, 3), device=venv.device)
    calls = LIMIT if mode == "same_step" else LIMIT + 1
    for ep in range(1, EPISODES):
        for _ in range(calls):
            venv.step(act)
        assert (venv.buffers["steps"].cpu().numpy() == 0).all(), "every env has just been respawned"
        check(ep)
    assert not venv.status().any()
    # the streams live on the device now, one track ahead of the running episode at most
    ahead = venv.buffers["trk_produced"].cpu().numpy() - venv.buffers["trk_consumed"].cpu().numpy()
    assert ((ahead >= 0) & (ahead <= 1)).all()
