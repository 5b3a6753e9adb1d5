"""On-device track generation against the native host generator, which is itself
pinned bit-for-bit to the reference's _create_track by tests/test_golden.py.

Same MT19937 stream, same draws, same float64 operation order; only libm differs (CUDA vs glibc:
<= 1-2 ulp in sin/cos/atan2).  Bars: every discrete result identical (T, Q, attempts, idx_range,
border pattern, palette, tile ids, RNG stream position); float64 nodes within 1e-9 absolute;
fp32 geometry (what Box2D / GL see) within 2 fp32 ulps and identical for >= 99 % of the values."""
import numpy as np
import pytest

from helpers import action_tape, make_oracle_worlds, gpu_state, oracle_state, visited_bits

pytestmark = pytest.mark.gpu


def test_device_generator_matches_host_generator(mcr):
    from multi_car_racing_b200.track import TrackGenerator
    n = 24
    venv = mcr.BatchedMultiCarRacing(n, num_agents=3, pool_tracks=2 * n, auto_reset=False, max_episode_steps=0)
    gen = TrackGenerator()
    host_rng = [np.random.RandomState(500 + i) for i in range(n)]
    dev_rng = [np.random.RandomState(500 + i) for i in range(n)]
    host = []
    for i in range(n):
        tr = gen.generate(host_rng[i])
        venv.load_track(i, tr)
        host.append(tr)
    res = venv.generate_tracks_device(list(range(n, 2 * n)), dev_rng)
    buf = {k: v.cpu().numpy() for k, v in venv.buffers.items() if k.startswith("trk_")}
    same32, total32 = 0, 0
    for i in range(n):
        h, d, tr = i, n + i, host[i]
        T, Q = tr.T, tr.Q
        assert res[i, 0] == T and res[i, 1] == tr.attempts and tuple(res[i, 2:4]) == tuple(tr.idx_range)
        assert buf["trk_T"][d] == T and buf["trk_Q"][d] == Q
        assert host_rng[i].uniform() == dev_rng[i].uniform(), "RandomState streams must stay in lock step"
        assert np.array_equal(buf["trk_quad_tile"][d, :Q], buf["trk_quad_tile"][h, :Q])      # border pattern + tile ids
        assert np.array_equal(buf["trk_quad_col"][d, :Q], buf["trk_quad_col"][h, :Q])        # palette
        dt = venv.tracks[d]
        assert np.abs(dt.nodes - tr.nodes).max() <= 1e-9
        assert np.abs(buf["trk_node"][d, :T] - buf["trk_node"][h, :T]).max() <= 1e-9
        assert np.abs(buf["trk_slot_pose"][d] - buf["trk_slot_pose"][h]).max() <= 1e-9
        for key, cnt in (("trk_quad", Q), ("trk_tile", T), ("trk_tile_aabb", T)):
            a, b = buf[key][d, :cnt], buf[key][h, :cnt]
            assert np.abs(a - b).max() <= 6.2e-5, key                     # 2 ulps of fp32 at |x| < 512
            same32 += int((a == b).sum()); total32 += a.size
        nch = (Q + 7) // 8
        assert np.abs(buf["trk_chunk"][d, :nch] - buf["trk_chunk"][h, :nch]).max() <= 1e-3
        assert np.array_equal(dt.quad_rgb, tr.quad_rgb) and np.array_equal(dt.quad_tile, tr.quad_tile)
    assert same32 >= 0.99 * total32, "only %d of %d fp32 values identical" % (same32, total32)


def test_reset_with_device_tracks_steps_like_the_oracle(oracle, mcr):
    """Whole path on device-generated tracks: reset(device_tracks=True), then the CUDA step against
    the oracle stepping on the SAME tracks (read back from the pool) -- bit-exact state and pixels."""
    import torch
    B, A = 4, 2
    np.random.seed(77)
    venv = mcr.BatchedMultiCarRacing(B, num_agents=A, auto_reset=False, max_episode_steps=0, seed=900)
    obs0 = venv.reset(device_tracks=True).cpu().numpy()
    tracks = [venv.tracks[e] for e in range(B)]
    orders = [np.array([venv.car_order[e][i] for i in range(A)]) for e in range(B)]
    worlds = make_oracle_worlds(oracle, tracks, orders, venv.episode_direction, A)
    # the oracle's spawn poses use glibc sin/cos on the same nodes: overwrite with the device's own
    # grid so that both sides start from identical fp32 poses
    poses = venv._pose.cpu().numpy()
    for e, w in enumerate(worlds):
        w.spawn(poses[e])
    oobs0 = np.stack([w.step(None)[0] for w in worlds])
    assert np.array_equal(obs0, oobs0)
    tape = action_tape(5, 120, B, A)
    for s in range(120):
        obs, rew, done, _ = venv.step(torch.from_numpy(tape[s]).to(venv.device))
        oo = [w.step(tape[s, e].astype(np.float64)) for e, w in enumerate(worlds)]
        assert np.array_equal(rew.cpu().numpy(), np.stack([x[1] for x in oo])), "rewards, step %d" % s
        if s % 10 == 0:
            assert np.array_equal(obs.cpu().numpy(), np.stack([x[0] for x in oo])), "pixels, step %d" % s
    g, o = gpu_state(venv), oracle_state(worlds)
    assert np.array_equal(g["bodies"], o["bodies"])
    for e, tr in enumerate(tracks):
        assert np.array_equal(visited_bits(g["visited"][e], tr.T, A), o["visited"][e])


def test_full_pool_device_reset_and_auto_reset(mcr):
    """pool_tracks > batch_envs: every slot is generated on the device; stepping with the device-side
    auto reset then only ever sees valid tracks (status words stay 0, observations change)."""
    import torch
    np.random.seed(3)
    venv = mcr.BatchedMultiCarRacing(8, num_agents=2, pool_tracks=24, auto_reset='next_step', max_episode_steps=30, seed=11)
    venv.reset(device_tracks=True)
    assert all(t is not None for t in venv.tracks)
    T = venv.buffers["trk_T"].cpu().numpy()
    assert (T >= 200).all() and (T <= venv.max_tiles).all()
    tape = action_tape(9, 70, 8, 2)
    seen_reset = 0
    for s in range(70):
        obs, rew, done, _ = venv.step(torch.from_numpy(tape[s]).to(venv.device))
        seen_reset += int((done.cpu().numpy() != 0).sum())
    assert seen_reset >= 8 and not venv.status().any()


def test_device_generator_reports_capacity_errors(mcr):
    """A pool whose slots are too small for the generated track: the kernel reports the error code per track
    (-3: more tiles than max_tiles) and the Python side raises instead of stepping on a half-written slot."""
    venv = mcr.BatchedMultiCarRacing(2, num_agents=1, max_tiles=64, max_quads=128, auto_reset=False, max_episode_steps=0)
    with pytest.raises(mcr.McrError, match="code -3"):
        venv.generate_tracks_device([0, 1], [np.random.RandomState(1), np.random.RandomState(2)])
