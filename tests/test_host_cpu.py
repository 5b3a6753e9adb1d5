"""CPU suite: host-side logic of the product (no GPU, no compute calls through the C-ABI's
device entry points) -- the C-ABI library loads and exports every symbol of include/mcr.h, the
native restatements (MT19937, _create_track, spawn grid, Box2D mass data) agree with the
oracle bit for bit, and the Python surface fails loudly without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol(mcr):
    from multi_car_racing_b200 import _lib
    L = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "mcr.h")).read()
    declared = set(re.findall(r"\b(mcr_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"mcr_handle", "mcr_config"}
    assert declared, "no declarations parsed"
    for sym in sorted(declared):
        assert hasattr(L, sym), "libmcr.so does not export %s" % sym
    assert set(_lib.EXPORTS) == declared
    assert L.mcr_abi_version() == 3


def test_config_validation_and_error_strings(mcr):
    from multi_car_racing_b200 import _lib
    L = _lib.load()
    h = ctypes.c_void_p()
    bad = _lib.McrConfig(1, 99, 512, 1024, 1, 1, 0, 0, 0.25, 0, 1, 0, 0, 0)
    assert L.mcr_create(ctypes.byref(bad), ctypes.byref(h)) < 0
    assert b"num_agents" in L.mcr_last_error()
    bad = _lib.McrConfig(0, 2, 512, 1024, 1, 1, 0, 0, 0.25, 0, 1, 0, 0, 0)
    assert L.mcr_create(ctypes.byref(bad), ctypes.byref(h)) < 0
    ok = _lib.McrConfig(4, 2, 512, 1024, 4, 1, 0, 1000, 0.25, 0, 1, 0, 0, 0)
    assert L.mcr_create(ctypes.byref(ok), ctypes.byref(h)) == 0
    # buffers are described but not bound: device entry points must refuse, not crash
    n = L.mcr_buffer_count(h)
    names = []
    for i in range(n):
        name, dt, nd = ctypes.c_char_p(), ctypes.c_int32(), ctypes.c_int32()
        dims = (ctypes.c_int64 * 4)()
        assert L.mcr_buffer_spec(h, i, ctypes.byref(name), ctypes.byref(dt), ctypes.byref(nd), ctypes.byref(dims)) == 0
        names.append(name.value.decode())
        assert all(dims[k] > 0 for k in range(nd.value))
    assert {"body", "joint", "wheel", "visited", "touched", "trk_quad", "trk_tile", "scratch"} <= set(names)
    assert L.mcr_contacts(h, None, None) < 0 and b"not bound" in L.mcr_last_error()
    assert L.mcr_bind_buffer(h, 0, ctypes.c_void_p(8)) < 0          # misaligned
    assert L.mcr_launch_count(h) == 0
    assert L.mcr_destroy(h) == 0


def test_car_constants_match_oracle(mcr, oracle):
    from multi_car_racing_b200 import _lib
    L = _lib.load()
    h = ctypes.c_void_p()
    cfg = _lib.McrConfig(1, 2, 512, 1024, 1, 1, 0, 0, 0.25, 0, 1, 0, 0, 0)
    assert L.mcr_create(ctypes.byref(cfg), ctypes.byref(h)) == 0
    m = np.empty(12, np.float32)
    L.mcr_get_mass(h, m.ctypes.data)
    w = oracle.OracleWorld(2)
    assert np.array_equal(m, w.mass())
    # SURVEY Appendix B
    assert abs(m[0] - 7.06) < 1e-5 and abs(m[2] - 18.2123) < 1e-3 and abs(m[6] - 0.06048) < 1e-6 and m[10] == 0 and m[11] == 0
    for which in range(5):
        o = np.empty(16, np.float32)
        n = L.mcr_get_shape(h, which, o.ctypes.data)
        assert np.array_equal(o[:2 * n].reshape(n, 2), w.shape(which))
    L.mcr_destroy(h)


def test_native_track_generator_matches_oracle(mcr, oracle):
    from multi_car_racing_b200.track import TrackGenerator
    gen = TrackGenerator()
    for seed in range(100, 130):
        r1, r2 = np.random.RandomState(seed), np.random.RandomState(seed)
        a = gen.generate(r1)
        b, attempts = oracle.generate_track(r2)
        assert a.attempts == attempts and a.idx_range == b.idx_range
        assert np.array_equal(a.nodes, b.nodes) and np.array_equal(a.quads, b.quads)
        assert np.array_equal(a.quad_rgb, b.quad_rgb) and np.array_equal(a.quad_tile, b.quad_tile)
        assert r1.uniform() == r2.uniform(), "RandomState streams must stay in lock step"


def test_gym_seeding_and_mt_seed(mcr, oracle):
    from multi_car_racing_b200 import _lib
    from multi_car_racing_b200.track import np_random, seed_key
    L = _lib.load()
    for seed in (0, 1, 42, 2 ** 40 + 7):
        r1, s1 = np_random(seed)
        r2, s2 = oracle.np_random(seed)
        assert s1 == s2 and r1.uniform() == r2.uniform()
        key, _ = seed_key(seed)
        st = np.zeros(625, np.uint32)
        k = np.array(key, np.uint32)
        assert L.mcr_mt_seed(st.ctypes.data, k.ctypes.data, len(k)) == 0
        rr = np.random.RandomState()
        rr.seed(key)
        assert np.array_equal(rr.get_state()[1], st[:624]) and st[624] == 624
    with pytest.raises(ValueError):
        np_random(-1)


def test_spawn_grid_matches_oracle(mcr, oracle):
    from multi_car_racing_b200.track import TrackGenerator
    gen = TrackGenerator()
    tr, _ = oracle.generate_track(np.random.RandomState(3))
    for A in (1, 2, 3, 8, 16):
        for cw in (0, 1):
            order = np.random.RandomState(A).permutation(A)
            p1 = gen.spawn_poses(tr.nodes, order, cw)
            p2 = oracle.spawn_poses([tuple(r) for r in tr.nodes], {i: order[i] for i in range(A)}, 'CW' if cw else 'CCW')
            assert np.array_equal(p1, p2)


def test_obs_format_abi(mcr):
    """mcr_set_obs_format / mcr_obs_bytes are host-only calls: exercised without a GPU."""
    import ctypes
    from multi_car_racing_b200 import _lib
    L = _lib.load()
    cfg = _lib.McrConfig(2, 2, 512, 1024, 2, 1, 0, 1000, 0.25, 0, 1, 0, 1, 0, 0)
    h = ctypes.c_void_p()
    assert L.mcr_create(ctypes.byref(cfg), ctypes.byref(h)) == 0
    assert L.mcr_obs_bytes(h) == 96 * 96 * 3
    assert L.mcr_set_obs_format(h, _lib.OBS_FORMATS["gray"]) == 0 and L.mcr_obs_bytes(h) == 96 * 96
    assert L.mcr_set_obs_format(h, _lib.OBS_FORMATS["rgb_chw"]) == 0 and L.mcr_obs_bytes(h) == 96 * 96 * 3
    assert L.mcr_set_obs_format(h, _lib.OBS_FORMATS["rgb_chw_f16"]) == 0 and L.mcr_obs_bytes(h) == 2 * 96 * 96 * 3
    assert L.mcr_set_obs_format(h, _lib.OBS_FORMATS["gray_stack"]) == 0 and L.mcr_obs_bytes(h) == 4 * 96 * 96
    assert L.mcr_set_frame_stack(h, 6) == 0 and L.mcr_obs_bytes(h) == 6 * 96 * 96
    assert L.mcr_set_frame_stack(h, 0) < 0 and L.mcr_set_frame_stack(h, 17) < 0
    assert L.mcr_set_obs_format(h, 7) < 0 and b"unknown format" in L.mcr_last_error()
    L.mcr_destroy(h)


def test_python_surface_without_gpu(mcr):
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a CPU-only box")
    with pytest.raises(mcr.McrError):
        mcr.BatchedMultiCarRacing(4)
    with pytest.raises(mcr.McrError):
        mcr.make("MultiCarRacing-v0", num_agents=1)
    with pytest.raises(ValueError):
        mcr.make("CarRacing-v0")


def test_box_and_timelimit_shims(mcr):
    box = mcr.Box(np.array([-1, 0, 0]), np.array([1, 1, 1]), dtype=np.float32)
    assert box.shape == (3,) and box.dtype == np.float32
    box.seed(0)
    for _ in range(10):
        assert box.contains(box.sample())
    obs = mcr.Box(0, 255, shape=(96, 96, 3), dtype=np.uint8)
    assert obs.shape == (96, 96, 3)

    class Dummy:
        def __init__(self):
            self.n = 0
            self.marker = "x"

        def reset(self):
            self.n = 0
            return "obs"

        def step(self, a):
            self.n += 1
            return "obs", 0.0, False, {}
    tl = mcr.TimeLimit(Dummy(), 3)
    with pytest.raises(AssertionError):
        tl.step(None)
    tl.reset()
    assert tl.marker == "x"
    assert [tl.step(None)[2] for _ in range(3)] == [False, False, True]
    tl.reset()
    o, r, d, info = tl.step(None)
    assert not d and info == {}
    tl.step(None)
    assert tl.step(None)[3] == {'TimeLimit.truncated': True}


def test_numa_binding_helper_is_safe_without_topology():
    from multi_car_racing_b200.dist import _parse_cpulist, bind_to_gpu_numa_node
    assert _parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert _parse_cpulist("") == set()
    before = os.sched_getaffinity(0)
    assert bind_to_gpu_numa_node(0) is None or isinstance(bind_to_gpu_numa_node(0), int)   # no GPU / no sysfs entry: no-op
    os.sched_setaffinity(0, before)


def test_env_sharding():
    from multi_car_racing_b200.dist import shard_envs
    for total, world in ((8192, 8), (1000, 3), (5, 8)):
        blocks = [shard_envs(total, r, world) for r in range(world)]
        assert sum(c for _, c in blocks) == total
        nxt = 0
        for first, count in blocks:
            assert first == nxt
            nxt += count


def test_reset_draws_match_numpy_global_rng(mcr):
    """mcr_reset_draws restates what reset() draws from the GLOBAL numpy RNG (reference :351-357):
    np.random.choice(['CW','CCW']) and np.random.choice(ids, size=A, replace=False) per env, in turn.
    Same values, and np.random ends at the same stream position."""
    from multi_car_racing_b200 import _lib
    L = _lib.load()
    for A, rand_dir, seed in [(1, 1, 0), (2, 1, 1), (2, 0, 2), (3, 1, 3), (8, 1, 4), (16, 1, 5), (5, 0, 6)]:
        n = 200
        np.random.seed(seed)
        want_cw, want_order = [], []
        for e in range(n):
            d = str(np.random.choice(['CW', 'CCW'])) if rand_dir else 'CCW'
            want_cw.append(d == 'CW')
            want_order.append(np.random.choice([i for i in range(A)], size=A, replace=False))
        tail_ref = np.random.uniform(size=3)
        np.random.seed(seed)
        st = np.random.get_state()
        mt = np.empty(625, np.uint32); mt[:624] = st[1]; mt[624] = st[2]
        cws = np.empty(n, np.uint8); orders = np.empty((n, A), np.int32)
        assert L.mcr_reset_draws(mt.ctypes.data, n, A, rand_dir, 0, cws.ctypes.data, orders.ctypes.data) == 0
        np.random.set_state((st[0], mt[:624].copy(), int(mt[624]), st[3], st[4]))
        assert np.array_equal(cws.astype(bool), np.array(want_cw)), (A, rand_dir)
        assert np.array_equal(orders, np.array(want_order)), (A, rand_dir)
        assert np.array_equal(np.random.uniform(size=3), tail_ref), "stream position after the draws"


def test_batched_seeding_matches_np_random(mcr):
    """mcr_mt_seed_batch == gym seeding.np_random (sha512 -> init_by_array) for a run of seeds."""
    from multi_car_racing_b200 import _lib
    from multi_car_racing_b200.track import np_random, seed_key
    L = _lib.load()
    seeds = [0, 1, 2, 12345, 2 ** 31 - 1, 2 ** 40 + 7] + list(range(100, 130))
    keys = np.zeros((len(seeds), 2), np.uint32); lens = np.empty(len(seeds), np.int32)
    for i, sd in enumerate(seeds):
        k, _ = seed_key(sd)
        keys[i, :len(k)] = k; lens[i] = len(k)
    states = np.empty((len(seeds), 625), np.uint32)
    assert L.mcr_mt_seed_batch(states.ctypes.data, keys.ctypes.data, lens.ctypes.data, len(seeds), 2) == 0
    for i, sd in enumerate(seeds):
        rng, _ = np_random(sd)
        st = rng.get_state()
        assert np.array_equal(states[i, :624], st[1]) and states[i, 624] == st[2], "seed %d" % sd


def test_gym_registration_path(monkeypatch):
    """reference gym_multi_car_racing/__init__.py:5-10: register(id, entry_point, max_episode_steps=1000,
    reward_threshold=900).  gym is not installable here, so a stand-in `gym.envs.registration` module records the call
    the package makes at import time; a package that refuses the registration produces a warning instead of silence."""
    import importlib, sys, types, warnings
    calls = []
    gym = types.ModuleType("gym"); envs = types.ModuleType("gym.envs"); reg = types.ModuleType("gym.envs.registration")
    reg.register = lambda **kw: calls.append(kw)
    gym.envs = envs; envs.registration = reg
    for name, mod in (("gym", gym), ("gym.envs", envs), ("gym.envs.registration", reg)):
        monkeypatch.setitem(sys.modules, name, mod)
    import multi_car_racing_b200 as pkg
    pkg = importlib.reload(pkg)
    assert calls == [dict(id="MultiCarRacing-v0", entry_point="multi_car_racing_b200:MultiCarRacing",
                          max_episode_steps=1000, reward_threshold=900)]
    assert "gym" in pkg.REGISTERED_WITH
    entry_mod, entry_cls = calls[0]["entry_point"].split(":")
    assert getattr(importlib.import_module(entry_mod), entry_cls) is pkg.MultiCarRacing

    def refuse(**kw):
        raise RuntimeError("registry is read-only")
    reg.register = refuse
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        pkg = importlib.reload(pkg)
    assert any("could not register" in str(x.message) for x in w) and "gym" not in pkg.REGISTERED_WITH
    monkeypatch.undo()
    importlib.reload(pkg)
