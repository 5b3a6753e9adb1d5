"""GPU parity: the CUDA path (through the C-ABI) against the CPU oracle, same seeded inputs.

Bars (BASELINE.json north_star): tile-visit flags / counts bit-exact, rewards bit-exact
(float64, same accumulation order), car pose within 1e-4 relative after 1000 steps (in
practice bit-exact: both sides run the same un-contracted fp32 operation sequence),
observations bit-exact against the oracle rasteriser."""
import numpy as np
import pytest

from helpers import action_tape, make_oracle_worlds, gpu_state, oracle_state, visited_bits

pytestmark = pytest.mark.gpu


def _setup(oracle, mcr, B, A, seed, directions=None, **kw):
    tracks, orders = [], []
    rs = np.random.RandomState(1000 + seed)
    for e in range(B):
        tr, _ = oracle.generate_track(np.random.RandomState(seed * 100 + e))
        tracks.append(tr)
        orders.append(rs.permutation(A))
    if directions is None:
        directions = ['CW' if rs.uniform() < 0.5 else 'CCW' for _ in range(B)]
    venv = mcr.BatchedMultiCarRacing(B, num_agents=A, auto_reset=False, max_episode_steps=0, **kw)
    obs0 = venv.reset(tracks=tracks, car_orders=orders, directions=directions).cpu().numpy()
    okw = dict(h_ratio=kw.get("h_ratio", 0.25), backwards_flag=kw.get("backwards_flag", True),
               use_ego_color=kw.get("use_ego_color", False), particles=kw.get("particles", False))
    worlds = make_oracle_worlds(oracle, tracks, orders, directions, A, **okw)
    oobs0 = np.stack([w.step(None)[0] for w in worlds])
    return venv, worlds, tracks, obs0, oobs0


def _compare_state(venv, worlds, tracks, step, exact=True):
    g, o = gpu_state(venv), oracle_state(worlds)
    A = venv.num_agents
    for e, tr in enumerate(tracks):
        assert np.array_equal(visited_bits(g["visited"][e], tr.T, A), o["visited"][e]), "visited flags, step %d env %d" % (step, e)
        assert np.array_equal(g["touched"][e][:tr.T], o["touched"][e]), "tile colour flags, step %d env %d" % (step, e)
    assert np.array_equal(g["counts"], o["counts"]), "tile_visited_count, step %d" % step
    assert np.array_equal(g["reward"], o["reward"]), "reward (float64, bit-exact), step %d" % step
    assert np.array_equal(g["backward"], o["backward"]), "driving_backward, step %d" % step
    assert np.array_equal(g["grass"], o["grass"]), "driving_on_grass, step %d" % step
    if exact:
        for k in ("bodies", "wheels", "joints"):
            assert np.array_equal(g[k], o[k]), "%s differ at step %d: max abs %g" % (k, step, np.abs(g[k].astype(np.float64) - o[k]).max())
    else:
        pose_g, pose_o = g["bodies"][..., :3].astype(np.float64), o["bodies"][..., :3].astype(np.float64)
        rel = np.abs(pose_g - pose_o) / np.maximum(np.abs(pose_o), 1.0)
        assert rel.max() <= 1e-4, "pose relative error %g > 1e-4 at step %d" % (rel.max(), step)


def test_reset_matches_oracle(oracle, mcr):
    venv, worlds, tracks, obs0, oobs0 = _setup(oracle, mcr, B=3, A=2, seed=1)
    _compare_state(venv, worlds, tracks, -1)
    assert np.array_equal(obs0, oobs0)
    assert not venv.status().any()


@pytest.mark.parametrize("A,B,seed", [(1, 2, 2), (2, 4, 3), (4, 2, 4), (8, 1, 5), (16, 1, 6)])
def test_step_parity_300(oracle, mcr, A, B, seed):
    import torch
    venv, worlds, tracks, obs0, oobs0 = _setup(oracle, mcr, B=B, A=A, seed=seed)
    tape = action_tape(seed, 300, B, A)
    for s in range(300):
        obs, rew, done, _ = venv.step(torch.from_numpy(tape[s]).to(venv.device))
        oo = [w.step(tape[s, e].astype(np.float64)) for e, w in enumerate(worlds)]
        assert np.array_equal(rew.cpu().numpy(), np.stack([x[1] for x in oo])), "step_reward, step %d" % s
        assert np.array_equal(done.cpu().numpy() & 1, np.array([x[2] for x in oo], np.uint8)), "done, step %d" % s
        if s % 10 == 0 or s > 290:
            _compare_state(venv, worlds, tracks, s)
            assert np.array_equal(obs.cpu().numpy(), np.stack([x[0] for x in oo])), "observation pixels, step %d" % s
    assert not venv.status().any()


@pytest.mark.parametrize("level", ["0", "1", "2"])
def test_flag_handoff_levels(oracle, mcr, level, monkeypatch):
    """The step's critical chain hands over by per-car / per-frame ready flags (DevBuffers::ready, csrc/mcr_internal.h):
    level 0 = plain stream order, 1 = sweep -> post -> project -> fill by flags, 2 = also contacts / stripes -> post (the
    default for small batches).  Every level, replayed from its CUDA graph with cars that touch (both classes of envs in
    flight), reproduces the oracle bit for bit."""
    import torch
    monkeypatch.setenv("MCR_FLAG_HANDOFF", level)
    B, A, seed = 6, 2, 21
    venv, worlds, tracks, obs0, oobs0 = _setup(oracle, mcr, B=B, A=A, seed=seed)
    tape = action_tape(seed, 120, B, A)
    tape[:, :, :, 0] *= 0.2; tape[:, :, 1, 1] = 1.0; tape[:, :, 1, 2] = 0.0; tape[:, :, 0, 1] = 0.1    # the rear car runs into the front one
    touched_cars = False
    for s in range(120):
        obs, rew, done, _ = venv.step(torch.from_numpy(tape[s]).to(venv.device))
        oo = [w.step(tape[s, e].astype(np.float64)) for e, w in enumerate(worlds)]
        touched_cars = touched_cars or bool((venv.buffers["n_manifold"] > 0).any().item())
        assert np.array_equal(rew.cpu().numpy(), np.stack([x[1] for x in oo])), "step_reward, step %d" % s
        assert np.array_equal(obs.cpu().numpy(), np.stack([x[0] for x in oo])), "observation pixels, step %d" % s
        if s % 20 == 0 or s == 119:
            _compare_state(venv, worlds, tracks, s)
    assert not venv.status().any()


def test_pose_after_1000_steps(oracle, mcr):
    """north_star: tile visits + rewards bit-exact and pose within 1e-4 relative after 1000 steps."""
    import torch
    B, A = 8, 2
    venv, worlds, tracks, _, _ = _setup(oracle, mcr, B=B, A=A, seed=7)
    tape = action_tape(7, 1000, B, A, brake_p=0.1)
    total_g = np.zeros((B, A)); total_o = np.zeros((B, A))
    for s in range(1000):
        obs, rew, done, _ = venv.step(torch.from_numpy(tape[s]).to(venv.device))
        total_g += rew.cpu().numpy()
        for e, w in enumerate(worlds):
            total_o[e] += w.step(tape[s, e].astype(np.float64), render=False)[1]
    assert np.array_equal(total_g, total_o)
    _compare_state(venv, worlds, tracks, 1000, exact=False)
    _compare_state(venv, worlds, tracks, 1000, exact=True)


@pytest.mark.parametrize("kw", [dict(use_ego_color=True), dict(h_ratio=0.5, backwards_flag=False)])
def test_render_options(oracle, mcr, kw):
    import torch
    venv, worlds, tracks, obs0, oobs0 = _setup(oracle, mcr, B=2, A=3, seed=11, **kw)
    assert np.array_equal(obs0, oobs0)
    tape = action_tape(11, 80, 2, 3)
    for s in range(80):
        obs, _, _, _ = venv.step(torch.from_numpy(tape[s]).to(venv.device))
        oo = [w.step(tape[s, e].astype(np.float64)) for e, w in enumerate(worlds)]
        if s % 8 == 0:
            assert np.array_equal(obs.cpu().numpy(), np.stack([x[0] for x in oo])), "pixels, step %d" % s


def _reformat(rgb, fmt):
    """numpy restatement of the fused store layouts (include/mcr.h MCR_OBS_*) applied to oracle frames."""
    if fmt == "gray":
        c = rgb.astype(np.int64)
        return ((299 * c[..., 0] + 587 * c[..., 1] + 114 * c[..., 2] + 500) // 1000).astype(np.uint8)
    if fmt == "rgb_chw":
        return np.ascontiguousarray(np.moveaxis(rgb, -1, -3))
    if fmt == "rgb_chw_f16":          # uint8 -> float normalisation fused into the store: fp32 division, RN to fp16
        return (np.moveaxis(rgb, -1, -3).astype(np.float32) / np.float32(255)).astype(np.float16)
    return rgb


@pytest.mark.parametrize("fmt", ["gray", "rgb_chw", "rgb_chw_f16"])
def test_fused_observation_formats(oracle, mcr, fmt):
    """grayscale / planar layouts written straight from the rasteriser's registers;
    bit-exact against the oracle's RGB frame pushed through the same integer formula.  Covers the
    reset frame, eager steps and CUDA-graph replayed steps (the graph bakes the layout in)."""
    import torch
    venv, worlds, tracks, obs0, oobs0 = _setup(oracle, mcr, B=2, A=2, seed=21, obs_format=fmt)
    assert obs0.shape == oobs0.shape[:2] + ({"gray": (96, 96), "rgb_chw": (3, 96, 96), "rgb_chw_f16": (3, 96, 96)}[fmt])
    assert obs0.dtype == (np.float16 if fmt == "rgb_chw_f16" else np.uint8)
    assert np.array_equal(obs0, _reformat(oobs0, fmt))
    tape = action_tape(21, 40, 2, 2)
    for s in range(40):
        obs, _, _, _ = venv.step(torch.from_numpy(tape[s]).to(venv.device))
        oo = np.stack([w.step(tape[s, e].astype(np.float64))[0] for e, w in enumerate(worlds)])
        if s % 5 == 0 or s < 4:
            assert np.array_equal(obs.cpu().numpy(), _reformat(oo, fmt)), "%s pixels, step %d" % (fmt, s)
    hobs, _, _, _ = venv.step_host(tape[0])
    assert hobs.shape == obs.shape


@pytest.mark.parametrize("K", [4, 3])
def test_frame_stack_ring(oracle, mcr, K):
    """SURVEY 8f #4 frame stack: obs_format='gray_stack' keeps the last K luma frames of every agent as a ring
    written by the rasteriser's store (slot = episode step % K, every slot on the first frame of an episode).
    stacked_obs() (oldest -> newest) must equal the oracle's frames pushed through the luma formula and a
    host-side deque -- across eager steps, CUDA-graph replay and a next-step auto reset."""
    import collections
    import torch
    B, A = 2, 2
    tracks = [oracle.generate_track(np.random.RandomState(700 + e))[0] for e in range(B)]
    orders = [np.array([0, 1]), np.array([1, 0])]
    directions = ['CCW', 'CW']
    venv = mcr.BatchedMultiCarRacing(B, num_agents=A, auto_reset=False, max_episode_steps=0, obs_format="gray_stack", frame_stack=K)
    venv.reset(tracks=tracks, car_orders=orders, directions=directions)
    worlds = make_oracle_worlds(oracle, tracks, orders, directions, A)
    f0 = _reformat(np.stack([w.step(None)[0] for w in worlds]), "gray")
    dq = collections.deque([f0] * K, maxlen=K)
    assert venv.obs.shape == (B, A, K, 96, 96)
    assert np.array_equal(venv.stacked_obs().cpu().numpy(), np.stack(dq, axis=2)), "reset fills every slot"
    tape = action_tape(71, 3 * K + 5, B, A)
    for s in range(len(tape)):
        venv.step(torch.from_numpy(tape[s]).to(venv.device))
        dq.append(_reformat(np.stack([w.step(tape[s, e].astype(np.float64))[0] for e, w in enumerate(worlds)]), "gray"))
        assert np.array_equal(venv.stacked_obs().cpu().numpy(), np.stack(dq, axis=2)), "stack after step %d" % s


def test_render_modes_between_steps(oracle, mcr):
    """render('state_pixels') and render('rgb_array') (600 x 400) called between steps:
    the tiled viewport rasteriser against the oracle's full-frame fill, bit-exact, zoomed-out first
    frames and mid-episode ones, with the live score label / backward flag."""
    import torch
    venv, worlds, tracks, obs0, oobs0 = _setup(oracle, mcr, B=2, A=2, seed=31)
    for mode in ("state_pixels", "rgb_array"):
        got = venv.render(mode).cpu().numpy()
        want = np.stack([w.render(mode) for w in worlds])
        assert got.shape == want.shape and np.array_equal(got, want), "%s after reset" % mode
    tape = action_tape(31, 90, 2, 2)
    for s in range(90):
        venv.step(torch.from_numpy(tape[s]).to(venv.device))
        for e, w in enumerate(worlds):
            w.step(tape[s, e].astype(np.float64), render=False)
        if s in (0, 3, 20, 49, 50, 89):
            for mode in ("state_pixels", "rgb_array"):
                got = venv.render(mode).cpu().numpy()
                want = np.stack([w.render(mode) for w in worlds])
                bad = int((got != want).any(axis=-1).sum())
                assert bad == 0, "%s, step %d: %d pixels differ" % (mode, s, bad)


def test_rgb_array_skid_particles(oracle, mcr):
    """particles=True: the skid traces of gym car_dynamics.Car (Car.step "Skid trace" block, Car.particles ring of
    30 polylines of <= 30 points) are kept on the device and drawn by render('rgb_array') before
    each car, like Car.draw(viewer, draw_particles=True) (mcr:564); state_pixels frames never show them.
    Bit-exact against the oracle: particle lists (order, grass flag, points) and the 600 x 400 frames, over
    enough hard-braking steps for the ring to wrap."""
    import torch
    venv, worlds, tracks, obs0, oobs0 = _setup(oracle, mcr, B=2, A=2, seed=51, particles=True)
    tape = action_tape(51, 400, 2, 2, brake_p=0.3)
    N = 4
    wrapped = False
    for s in range(400):
        obs, _, _, _ = venv.step(torch.from_numpy(tape[s]).to(venv.device))
        oo = [w.step(tape[s, e].astype(np.float64)) for e, w in enumerate(worlds)]
        if s % 40 == 39:
            assert np.array_equal(obs.cpu().numpy(), np.stack([x[0] for x in oo])), "state pixels, step %d" % s
            hdr = venv.buffers["prt_hdr"].cpu().numpy()
            meta = venv.buffers["prt_meta"].cpu().numpy()
            pts = venv.buffers["prt_pts"].cpu().numpy().reshape(30, 30, 2, N)
            for e, w in enumerate(worlds):
                for c in range(2):
                    car = e * 2 + c
                    want = w.particles(c)
                    head, count = int(hdr[0, car]), int(hdr[1, car])
                    assert count == len(want), "particle count, step %d car %d" % (s, car)
                    wrapped |= head != 0
                    for i, (grass, p) in enumerate(want):
                        slot = (head + i) % 30
                        ln = int(meta[slot, car]) & 0xff
                        assert ln == len(p) and bool(int(meta[slot, car]) & 256) == grass
                        assert np.array_equal(pts[slot, :ln, :, car], p), "particle points, step %d car %d" % (s, car)
            got = venv.render('rgb_array').cpu().numpy()
            want_img = np.stack([w.render('rgb_array') for w in worlds])
            bad = int((got != want_img).any(axis=-1).sum())
            assert bad == 0, "rgb_array with particles, step %d: %d pixels differ" % (s, bad)
    assert wrapped, "the test tape must make the particle ring wrap"


def test_viewport_tiling_partial_tiles_ego_color(oracle, mcr):
    """mcr_render_viewport on viewports that do not divide into 96 x 96 tiles (partial tiles on the right
    and top edges, a single partial tile, a human-sized window), 3 agents with use_ego_color and a
    non-default h_ratio: bit-exact against the oracle's full-frame fill."""
    import torch
    venv, worlds, tracks, obs0, oobs0 = _setup(oracle, mcr, B=2, A=3, seed=41, use_ego_color=True, h_ratio=0.4)
    tape = action_tape(41, 70, 2, 3)
    for s in range(70):
        venv.step(torch.from_numpy(tape[s]).to(venv.device))
        for e, w in enumerate(worlds):
            w.step(tape[s, e].astype(np.float64), render=False)
        if s in (1, 69):
            for vw, vh in ((200, 120), (52, 40), (1000, 800)):
                got = venv.render_viewport(vw, vh).cpu().numpy()
                want = np.stack([w.render_viewport(vw, vh) for w in worlds])
                bad = int((got != want).any(axis=-1).sum())
                assert got.shape == want.shape and bad == 0, "%dx%d, step %d: %d pixels differ" % (vw, vh, s, bad)
    with pytest.raises(mcr.McrError):
        venv.render_viewport(50, 40)            # width must be a multiple of 4


def test_single_env_dropin_matches_oracle_env(oracle, mcr):
    """The reference-shaped API end to end: same seeds -> same tracks, spawn, rewards, frames."""
    np.random.seed(5)
    ref = oracle.OracleMultiCarRacing(num_agents=2, verbose=0)
    ref.seed(42)
    o0 = ref.reset()
    np.random.seed(5)
    env = mcr.MultiCarRacing(num_agents=2, verbose=0)
    assert env.seed(42) == [42]
    g0 = env.reset()
    assert g0.shape == (2, 96, 96, 3) and g0.dtype == np.uint8
    assert np.array_equal(g0, o0)
    assert env.episode_direction == ref.episode_direction
    assert len(env.track) == len(ref.track)
    rs = np.random.RandomState(9)
    for s in range(60):
        a = np.stack([rs.uniform(-1, 1, 2), rs.uniform(0, 1, 2), np.zeros(2)], 1)
        go, gr, gd, gi = env.step(a)
        oo, orr, od, _ = ref.step(a)
        assert gr.dtype == np.float64 and gr.shape == (2,) and gi == {} and isinstance(gd, bool)
        assert np.array_equal(gr, orr) and gd == od
        assert np.array_equal(go, oo), "pixels at step %d" % s
    assert env.tile_visited_count == [int(v) for v in ref.tile_visited_count]


def test_car_car_collisions(oracle, mcr):
    """Rear row drives into the braking front row: hull-hull and wheel-hull manifolds, block solver,
    warm starting by contact id, merged islands -- state bit-exact against the oracle every step."""
    import torch
    B, A, steps = 2, 4, 260
    venv, worlds, tracks, obs0, oobs0 = _setup(oracle, mcr, B=B, A=A, seed=21, directions=['CCW', 'CW'])
    rs = np.random.RandomState(3)
    saw_manifolds = 0
    for s in range(steps):
        a = np.zeros((B, A, 3), np.float32)
        for e in range(B):
            order = venv.car_order[e]
            for c in range(A):
                front = order[c] < 2
                a[e, c] = (rs.uniform(-0.2, 0.2), 0.0, 1.0) if front else (rs.uniform(-0.2, 0.2), 1.0, 0.0)
        obs, rew, done, _ = venv.step(torch.from_numpy(a).to(venv.device))
        oo = [w.step(a[e].astype(np.float64)) for e, w in enumerate(worlds)]
        assert np.array_equal(rew.cpu().numpy(), np.stack([x[1] for x in oo])), "step_reward, step %d" % s
        nman = venv.buffers["n_manifold"].cpu().numpy()
        assert list(nman) == [len(w.manifolds()) for w in worlds], "manifold count, step %d" % s
        saw_manifolds = max(saw_manifolds, int(nman.max()))
        if s % 5 == 0 or nman.max() > 0:
            _compare_state(venv, worlds, tracks, s)
        if s % 20 == 0:
            assert np.array_equal(obs.cpu().numpy(), np.stack([x[0] for x in oo])), "pixels, step %d" % s
    assert saw_manifolds >= 2, "the tape never produced a car-car contact"
    assert not venv.status().any()


def test_collisions_keep_cars_apart(mcr):
    """Physical sanity of the solid contacts: with collisions the pushing car never overlaps the
    pushed one; without them (collisions=False) it drives straight through."""
    import torch
    mins = {}
    for coll in (True, False):
        np.random.seed(0)
        venv = mcr.BatchedMultiCarRacing(1, num_agents=4, auto_reset=False, max_episode_steps=0, seed=5,
                                         use_random_direction=False, collisions=coll)
        venv.reset(car_orders=[np.arange(4)])
        a = torch.zeros((1, 4, 3), device=venv.device)
        a[0, 0, 2] = 1.0; a[0, 1, 2] = 1.0; a[0, 2, 1] = 1.0; a[0, 3, 1] = 1.0
        dmin = 1e9
        for s in range(200):
            venv.step(a)
            p = venv.bodies()[0, :, 0, 6:8].cpu().numpy()
            dmin = min(dmin, float(np.linalg.norm(p[2] - p[0])))
        mins[coll] = dmin
    assert mins[True] > 4.0 and mins[False] < 2.0, mins


@pytest.mark.parametrize("A,B,seed", [(8, 2, 41), (2, 64, 43)])
def test_1000_steps_random_policy_with_collisions(oracle, mcr, A, B, seed):
    """the 1000-step bar with car-car contacts in play -- BASELINE.json configs[3]'s 8 agents
    (use_ego_color=True; manifolds in most steps) and a 64-env batch of 2 agents.  Rewards and done every step,
    manifold counts, full state bit-exact every 100 steps and at the end, pixels at a few steps."""
    import torch
    kw = dict(use_ego_color=True) if A == 8 else {}
    venv, worlds, tracks, obs0, oobs0 = _setup(oracle, mcr, B=B, A=A, seed=seed, **kw)
    assert np.array_equal(obs0, oobs0)
    tape = action_tape(seed, 1000, B, A, brake_p=0.15)
    steps_with_manifolds = 0
    for s in range(1000):
        obs, rew, done, _ = venv.step(torch.from_numpy(tape[s]).to(venv.device))
        want_pixels = s in (0, 499, 999)
        oo = [w.step(tape[s, e].astype(np.float64), render=want_pixels) for e, w in enumerate(worlds)]
        assert np.array_equal(rew.cpu().numpy(), np.stack([x[1] for x in oo])), "step_reward, step %d" % s
        assert np.array_equal(done.cpu().numpy() & 1, np.array([x[2] for x in oo], np.uint8)), "done, step %d" % s
        if s % 100 == 99:
            nman = venv.buffers["n_manifold"].cpu().numpy()
            assert list(nman) == [len(w.manifolds()) for w in worlds], "manifold count, step %d" % s
            steps_with_manifolds += int(nman.max() > 0)
            _compare_state(venv, worlds, tracks, s)
        if want_pixels:
            assert np.array_equal(obs.cpu().numpy(), np.stack([x[0] for x in oo])), "pixels, step %d" % s
    if A == 8:
        assert steps_with_manifolds >= 1, "eight cars on one track never touched"
    assert not venv.status().any()
