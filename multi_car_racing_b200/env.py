"""Host-side mirror of the reference's env surface for the B200 step+render path.

Two classes:
  * BatchedMultiCarRacing -- B independent MultiCarRacing-v0 envs on one GPU; torch tensors in
    and out, state held in a torch SoA that libmcr.so's kernels read and write in place.
  * MultiCarRacing        -- the reference's single-env API (same kwargs, defaults, shapes,
    dtypes and RNG usage; reference gym_multi_car_racing/multi_car_racing.py:125-674) on top of
    a batch of one: numpy observations (num_agents, 96, 96, 3) uint8, rewards (num_agents,)
    float64, done bool, info {}.

All compute goes through the C-ABI of include/mcr.h; this module owns memory (torch), streams
and RNG plumbing only.  It raises if CUDA or libmcr.so is unavailable -- there is no fallback.
"""
import ctypes

import numpy as np

from . import _lib
from .track import TrackGenerator, np_random, seed_key, MAX_TILES_DEFAULT, MAX_QUADS_DEFAULT

STATE_W = 96
STATE_H = 96
VIDEO_W = 600
VIDEO_H = 400
FPS = 50
PLAYFIELD = 2000 / 6.0

_TORCH_DTYPES = None


def _torch():
    import torch
    global _TORCH_DTYPES
    if _TORCH_DTYPES is None:
        _TORCH_DTYPES = {_lib.MCR_U8: torch.uint8, _lib.MCR_I32: torch.int32, _lib.MCR_U32: torch.int32,
                         _lib.MCR_F32: torch.float32, _lib.MCR_F64: torch.float64, _lib.MCR_I16: torch.int16}
    return torch


class Box:
    """Minimal stand-in for gym.spaces.Box (gym is an optional dependency)."""

    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        if shape is None:
            low = np.asarray(low)
            shape = low.shape
        self.shape = tuple(shape)
        self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()
        self._rng = np.random.RandomState()

    def seed(self, seed=None):
        self._rng.seed(seed)
        return [seed]

    def sample(self):
        if self.dtype.kind == "f":
            return self._rng.uniform(self.low, self.high).astype(self.dtype)
        return self._rng.randint(self.low, self.high.astype(np.int64) + 1).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high)

    def __repr__(self):
        return "Box(%s, %s, %s, %s)" % (self.low.min(), self.high.max(), self.shape, self.dtype)


def _make_box(low, high, shape=None, dtype=np.float32):
    try:  # use the real thing when a gym is installed, so isinstance checks downstream work
        from gym import spaces  # type: ignore
        return spaces.Box(low, high, shape=shape, dtype=dtype) if shape is not None else spaces.Box(low, high, dtype=dtype)
    except Exception:
        return Box(low, high, shape=shape, dtype=dtype)


class BatchedMultiCarRacing:
    """B independent MultiCarRacing-v0 environments stepped by three CUDA kernels.

    Constructor kwargs mirror MultiCarRacing.__init__ (reference :131-133); `batch_envs`,
    `device`, the pool/capacity knobs and `max_episode_steps` (the gym registration's TimeLimit,
    reference __init__.py:8) are additions.  `obs_format` selects what the rasteriser stores:
    'rgb' (B,A,96,96,3) as the reference returns it, 'gray' (B,A,96,96) ITU-R 601 luma, or
    'rgb_chw' (B,A,3,96,96), 'gray_stack' (B,A,K,96,96) -- a ring of the last K = `frame_stack` luma
    frames, see stacked_obs() -- or 'rgb_chw_f16' (B,A,3,96,96) float16 = value / 255: the learner's
    first pre-processing stages fused into the store.
    `fresh_tracks=R` (default 1 when auto_reset is on and no explicit `pool_tracks` is given): every
    auto-reset episode runs on a NEW track generated on the GPU from the env's own RandomState stream,
    as the reference's reset() does (:359-364) -- env e owns R + 1 pool slots, R tracks are kept ready
    ahead of the running episode.  `fresh_tracks=0` with `pool_tracks=P` is the shared-pool mode: auto
    reset draws one of P tracks generated at reset().
    `particles=True` keeps the cars' skid traces (gym car_dynamics Car.particles) so that
    render('rgb_array') draws them like the reference does in its non-state modes (:564).
    """

    def __init__(self, batch_envs, num_agents=2, verbose=0, direction='CCW', use_random_direction=True,
                 backwards_flag=True, h_ratio=0.25, use_ego_color=False, device=None,
                 max_tiles=MAX_TILES_DEFAULT, max_quads=MAX_QUADS_DEFAULT, pool_tracks=None,
                 max_episode_steps=1000, auto_reset=True, seed=None, collisions=True, obs_format='rgb',
                 particles=False, frame_stack=4, fresh_tracks=None):
        torch = _torch()
        if not torch.cuda.is_available():
            raise _lib.McrError("multi_car_racing_b200 needs a CUDA device (B200, sm_100a); none is visible")
        self.L = _lib.load()
        self.device = torch.device(device if device is not None else "cuda:0")
        if self.device.type != "cuda":
            raise _lib.McrError("device must be a CUDA device")
        self.batch_envs, self.num_agents = int(batch_envs), int(num_agents)
        self.verbose = verbose
        self.direction, self.use_random_direction = direction, use_random_direction
        self.backwards_flag, self.h_ratio, self.use_ego_color = backwards_flag, h_ratio, use_ego_color
        self.max_episode_steps = int(max_episode_steps or 0)
        # auto_reset: False | True / 'same_step' (the finished env's obs is replaced by its next
        # episode's first frame before step() returns) | 'next_step' (EnvPool convention: the terminal
        # obs is returned; the NEXT step() ignores that env's action and returns the reset frame with
        # reward 0 -- one kernel pass per step, the fast path bench.py measures)
        if auto_reset not in (False, True, 'same_step', 'next_step', 0, 1, None):
            raise ValueError("auto_reset must be False, True, 'same_step' or 'next_step'")
        self.auto_reset = auto_reset if isinstance(auto_reset, str) else bool(auto_reset)
        self._step_flags = 2 if self.auto_reset == 'next_step' else (1 if self.auto_reset else 0)
        if fresh_tracks is None:
            fresh_tracks = 1 if (self.auto_reset and not pool_tracks) else 0
        self.fresh_tracks = int(fresh_tracks)
        if self.fresh_tracks:
            if pool_tracks and int(pool_tracks) != self.batch_envs * (self.fresh_tracks + 1):
                raise ValueError("fresh_tracks=R uses a ring of R + 1 slots per env: leave pool_tracks unset")
            self.pool_tracks = self.batch_envs * (self.fresh_tracks + 1)
        else:
            self.pool_tracks = int(pool_tracks) if pool_tracks else self.batch_envs
        if self.pool_tracks < self.batch_envs:
            raise ValueError("pool_tracks must be >= batch_envs")
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        cfg = _lib.McrConfig(self.batch_envs, self.num_agents, max_tiles, max_quads, self.pool_tracks,
                             int(bool(backwards_flag)), int(bool(use_ego_color)), self.max_episode_steps,
                             float(h_ratio), dev_index, int(bool(use_random_direction)),
                             int(direction == 'CW'), int(bool(collisions)), int(seed if seed is not None else 0) & (2 ** 64 - 1),
                             int(bool(particles)), self.fresh_tracks)
        self.particles = bool(particles)
        self._h = ctypes.c_void_p()
        _lib.check(self.L.mcr_create(ctypes.byref(cfg), ctypes.byref(self._h)), "mcr_create")
        self.max_tiles, self.max_quads = max_tiles, max_quads
        # ---- torch-owned SoA state, bound once ------------------------------------------------
        self.buffers = {}
        with torch.cuda.device(self.device):
            n = self.L.mcr_buffer_count(self._h)
            for i in range(n):
                name, dt, nd = ctypes.c_char_p(), ctypes.c_int32(), ctypes.c_int32()
                dims = (ctypes.c_int64 * 4)()
                _lib.check(self.L.mcr_buffer_spec(self._h, i, ctypes.byref(name), ctypes.byref(dt), ctypes.byref(nd),
                                                  ctypes.byref(dims)), "mcr_buffer_spec")
                shape = tuple(int(dims[k]) for k in range(nd.value))
                t = torch.zeros(shape, dtype=_TORCH_DTYPES[dt.value], device=self.device)
                self.buffers[name.value.decode()] = t
                _lib.check(self.L.mcr_bind_buffer(self._h, i, t.data_ptr()), "mcr_bind_buffer")
            B, A = self.batch_envs, self.num_agents
            if obs_format not in _lib.OBS_FORMATS:
                raise ValueError("obs_format must be one of %s" % sorted(_lib.OBS_FORMATS))
            self.obs_format = obs_format
            _lib.check(self.L.mcr_set_obs_format(self._h, _lib.OBS_FORMATS[obs_format]), "mcr_set_obs_format")
            self.frame_stack = int(frame_stack)
            _lib.check(self.L.mcr_set_frame_stack(self._h, self.frame_stack), "mcr_set_frame_stack")
            self.obs_shape = {"rgb": (STATE_H, STATE_W, 3), "gray": (STATE_H, STATE_W), "rgb_chw": (3, STATE_H, STATE_W),
                              "gray_stack": (self.frame_stack, STATE_H, STATE_W), "rgb_chw_f16": (3, STATE_H, STATE_W)}[obs_format]
            self.obs_dtype = torch.float16 if obs_format == "rgb_chw_f16" else torch.uint8
            assert int(np.prod(self.obs_shape)) * (2 if self.obs_dtype == torch.float16 else 1) == int(self.L.mcr_obs_bytes(self._h))
            self.obs = torch.zeros((B, A) + self.obs_shape, dtype=self.obs_dtype, device=self.device)
            self.reward_out = torch.zeros((B, A), dtype=torch.float64, device=self.device)
            self.done_out = torch.zeros((B,), dtype=torch.uint8, device=self.device)
            self._slot = torch.zeros((B,), dtype=torch.int32, device=self.device)
            self._cw = torch.zeros((B,), dtype=torch.uint8, device=self.device)
            self._pose = torch.zeros((B, A, 3), dtype=torch.float64, device=self.device)
        self._gen = TrackGenerator(max_tiles, max_quads)
        self.tracks = _TrackTable(self)                  # HostTrack / DeviceTrack per pool slot
        self.episode_direction = [direction] * self.batch_envs
        self.car_order = [None] * self.batch_envs
        self.action_space = _make_box(np.array([-1, 0, 0]), np.array([+1, +1, +1]), dtype=np.float32)
        if self.obs_dtype == torch.float16:
            self.observation_space = _make_box(0.0, 1.0, shape=self.obs_shape, dtype=np.float16)
        else:
            self.observation_space = _make_box(0, 255, shape=self.obs_shape, dtype=np.uint8)
        self.seed(seed)

    # ---- plumbing ---------------------------------------------------------------------------
    def _stream(self):
        return ctypes.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                self.L.mcr_destroy(self._h)
                self._h = ctypes.c_void_p()
        except Exception:
            pass

    def close(self):
        try:
            if getattr(self, "_h", None) and self._h.value and self.status().any():
                import warnings
                warnings.warn("multi_car_racing_b200: device status words set at close(): %s" % self.status().tolist())
        except Exception:
            pass
        self.__del__()

    @property
    def launch_count(self):
        return int(self.L.mcr_launch_count(self._h))

    def seed(self, seed=None):
        """Per-env RandomState streams: env i is seeded like the reference's env.seed(seed + i)
        (gym seeding.np_random: sha512 of the seed -> init_by_array).  The MT19937 states are built in one
        native call; `np_randoms` materialises numpy RandomState objects from them on first use."""
        B = self.batch_envs
        keys = np.zeros((B, 2), np.uint32)
        lens = np.empty((B,), np.int32)
        seeds = []
        for i in range(B):
            k, sd = seed_key(None if seed is None else int(seed) + i)
            keys[i, :len(k)] = k
            lens[i] = len(k)
            seeds.append(sd)
        self._mt_host = np.empty((B, 625), np.uint32)
        _lib.check(self.L.mcr_mt_seed_batch(self._mt_host.ctypes.data, keys.ctypes.data, lens.ctypes.data, B, 2), "mcr_mt_seed_batch")
        self._np_randoms = None
        self._mt_on_device = False       # True: buffers['mt_state'] is ahead of the host copies
        return seeds

    @property
    def np_randoms(self):
        """numpy RandomState of every env (the reference's env.np_random).  After a device-side generation the
        streams live on the GPU (buffers['mt_state'], R tracks ahead when fresh_tracks=R); they are read back here."""
        if self._mt_on_device:
            self._mt_host = self.buffers["mt_state"].cpu().numpy().view(np.uint32).copy()
            self._mt_on_device = False
            self._np_randoms = None
        if self._np_randoms is None:
            self._np_randoms = []
            for row in self._mt_host:
                rng = np.random.RandomState()
                rng.set_state(("MT19937", row[:624].copy(), int(row[624]), 0, 0.0))
                self._np_randoms.append(rng)
        return self._np_randoms

    @np_randoms.setter
    def np_randoms(self, rngs):
        self._np_randoms = list(rngs)
        self._mt_on_device = False

    def _mt_states(self):
        """(B, 625) uint32 MT19937 states of the envs' streams as they stand on the host."""
        if self._np_randoms is not None:
            for i, rng in enumerate(self._np_randoms):
                st = rng.get_state()
                self._mt_host[i, :624] = st[1]
                self._mt_host[i, 624] = st[2]
        return self._mt_host

    def _reset_draws(self, car_orders=None, directions=None):
        """reset()'s draws from the GLOBAL numpy RNG for every env in turn (reference :351-357), in one native
        call on np.random's MT19937 state (same stream positions as B reference resets)."""
        B, A = self.batch_envs, self.num_agents
        cws = np.empty((B,), np.uint8)
        orders = np.empty((B, A), np.int32)
        if car_orders is None and (directions is None or not self.use_random_direction):
            st = np.random.get_state()
            mt = np.empty(625, np.uint32)
            mt[:624] = st[1]; mt[624] = st[2]
            rand_dir = self.use_random_direction and directions is None
            _lib.check(self.L.mcr_reset_draws(mt.ctypes.data, B, A, int(rand_dir), int(self.direction == 'CW'),
                                              cws.ctypes.data, orders.ctypes.data), "mcr_reset_draws")
            np.random.set_state((st[0], mt[:624].copy(), int(mt[624]), st[3], st[4]))
            if directions is not None:
                cws[:] = [d == 'CW' for d in directions]
            elif not self.use_random_direction:
                cws[:] = [d == 'CW' for d in self.episode_direction]
        else:
            for e in range(B):                      # mixed injection: the reference's own calls, env by env
                if directions is not None:
                    d = directions[e]
                elif self.use_random_direction:
                    d = str(np.random.choice(['CW', 'CCW']))
                else:
                    d = self.episode_direction[e]
                cws[e] = d == 'CW'
                orders[e] = np.asarray(car_orders[e]) if car_orders is not None else \
                    np.random.choice([i for i in range(A)], size=A, replace=False)
        return cws, orders

    def load_track(self, slot, track):
        """Upload a HostTrack (or any object with nodes/quads/quad_rgb/quad_tile) into a pool slot."""
        nodes = np.ascontiguousarray(track.nodes, np.float64)
        quads = np.ascontiguousarray(track.quads, np.float64).reshape(-1, 8)
        rgb = np.ascontiguousarray(track.quad_rgb, np.float32)
        qt = np.ascontiguousarray(track.quad_tile, np.int32)
        with _torch().cuda.device(self.device):
            _lib.check(self.L.mcr_load_track(self._h, int(slot), len(nodes), nodes.ctypes.data, len(quads),
                                             quads.ctypes.data, rgb.ctypes.data, qt.ctypes.data, self._stream()),
                       "mcr_load_track")
        self.tracks[slot] = track

    # ---- reset, reference :340-408 ---------------------------------------------------------------
    def reset(self, tracks=None, car_orders=None, directions=None, device_tracks=False):
        """Reset every env.  RNG usage per env follows the reference: direction and car order from
        the GLOBAL numpy RNG (:351-357), the track from the env's own RandomState (:359-364).
        `tracks` / `car_orders` / `directions` inject host-made values (parity tests).
        device_tracks=True generates all tracks on the GPU (mcr_tracks_generate_device, one thread
        per track on the env's own MT19937 stream) instead of one by one on the host -- same draws,
        same arithmetic, CUDA libm instead of glibc (coordinates agree to ~1e-12, not bit for bit);
        nothing in that path loops over envs in Python."""
        torch = _torch()
        B, A, P = self.batch_envs, self.num_agents, self.pool_tracks
        if device_tracks and tracks is not None:
            raise ValueError("device_tracks=True generates the tracks itself; do not pass `tracks`")
        cws, orders = self._reset_draws(car_orders, directions)
        self.episode_direction = np.where(cws != 0, 'CW', 'CCW')
        self.car_order = orders
        ring = self.fresh_tracks > 0
        poses = None
        if device_tracks and (ring or P == B):
            self._generate_episode0_tracks()
        elif device_tracks:
            # shared pool larger than the batch: every slot in rounds of B, slot s draws from env (s % B)'s stream
            rngs = self.np_randoms
            for s0 in range(0, P, B):
                slots = list(range(s0, min(s0 + B, P)))
                if s0 == 0 or any(self.tracks.host.get(s) is None for s in slots):
                    self.generate_tracks_device(slots, [rngs[s % B] for s in slots])
        else:
            rngs = self.np_randoms
            poses = np.empty((B, A, 3), np.float64)
            for e in range(B):
                tr = tracks[e] if tracks is not None else self._gen.generate(rngs[e], self.verbose)
                self.load_track(e, tr)
                poses[e] = self._gen.spawn_poses(tr.nodes, orders[e], cws[e])
            if not ring:
                for s in range(B, P):        # extra pool slots (the shared-pool auto reset draws from the whole pool)
                    if self.tracks.host.get(s) is None:
                        self.load_track(s, self._gen.generate(rngs[s % B], 0))
        with torch.cuda.device(self.device):
            if ring and not (device_tracks and self._mt_on_device):
                # the env streams continue on the device: the refill kernel generates the next episodes' tracks from them
                self.buffers["mt_state"].copy_(torch.from_numpy(self._mt_states().view(np.int32)))
                self._mt_on_device = True
                self._np_randoms = None
            self._slot.copy_(torch.arange(B, dtype=torch.int32))
            self._cw.copy_(torch.from_numpy(cws))
            if poses is None:
                # spawn grid poses were evaluated per slot / direction / grid position by the generator
                sp = self.buffers["trk_slot_pose"]                     # (P, 2, A, 3)
                e_idx = torch.arange(B, device=self.device).view(B, 1).expand(B, A)
                cw_idx = torch.from_numpy(cws.astype(np.int64)).to(self.device).view(B, 1).expand(B, A)
                self._pose.copy_(sp[e_idx, cw_idx, torch.from_numpy(orders.astype(np.int64)).to(self.device)])
            else:
                self._pose.copy_(torch.from_numpy(poses))
            _lib.check(self.L.mcr_reset(self._h, None, self._slot.data_ptr(), self._cw.data_ptr(),
                                        self._pose.data_ptr(), self.obs.data_ptr(), self._stream()), "mcr_reset")
        return self.obs

    def _generate_episode0_tracks(self):
        """All B first-episode tracks in one launch: env e's stream (uploaded to buffers['mt_state']) -> pool slot e."""
        torch = _torch()
        B = self.batch_envs
        stride = int(self.L.mcr_trackgen_scratch_bytes())
        with torch.cuda.device(self.device):
            self.buffers["mt_state"].copy_(torch.from_numpy(self._mt_states().view(np.int32)))
            scratch = self.buffers["tg_scratch"] if self.fresh_tracks else torch.empty((B, stride), dtype=torch.uint8, device=self.device)
            d_res = torch.zeros((B, 4), dtype=torch.int32, device=self.device)
            _lib.check(self.L.mcr_tracks_generate_device(self._h, B, self.buffers["mt_state"].data_ptr(), None, scratch.data_ptr(),
                                                         d_res.data_ptr(), self._stream()), "mcr_tracks_generate_device")
            res = d_res.cpu().numpy()
        if (res[:, 0] <= 0).any():
            bad = int(np.argmax(res[:, 0] <= 0))
            raise _lib.McrError("device track generation failed for env %d: code %d (-3: more tiles than max_tiles, "
                                "-4: more quads than max_quads, -5: no valid track in 64 attempts)" % (bad, res[bad, 0]))
        self._mt_on_device = True
        self._np_randoms = None
        self.tracks.host.clear()
        if self.verbose == 1:
            for i in range(B):
                print("Track generation: %i..%i -> %i-tiles track" % (res[i, 2], res[i, 3], res[i, 3] - res[i, 2]))
        return res

    def generate_tracks_device(self, slots, rngs):
        """Generate len(slots) tracks on the GPU into the given pool slots, each on its own numpy
        RandomState (MT19937) stream, which is advanced exactly as the host generator would.
        Fills self.tracks[slot] with DeviceTrack views.  Returns the (n, 4) int32 result rows
        (T, attempts, i1, i2)."""
        torch = _torch()
        n = len(slots)
        mt = np.empty((n, 625), np.uint32)
        states = []
        for i, rng in enumerate(rngs):
            st = rng.get_state()
            assert st[0] == "MT19937"
            mt[i, :624] = st[1]
            mt[i, 624] = st[2]
            states.append(st)
        stride = int(self.L.mcr_trackgen_scratch_bytes())
        with torch.cuda.device(self.device):
            d_mt = torch.from_numpy(mt.view(np.int32)).to(self.device)
            d_slot = torch.tensor(list(slots), dtype=torch.int32, device=self.device)
            d_scratch = torch.empty((n, stride), dtype=torch.uint8, device=self.device)
            d_res = torch.zeros((n, 4), dtype=torch.int32, device=self.device)
            _lib.check(self.L.mcr_tracks_generate_device(self._h, n, d_mt.data_ptr(), d_slot.data_ptr(), d_scratch.data_ptr(),
                                                         d_res.data_ptr(), self._stream()), "mcr_tracks_generate_device")
            res = d_res.cpu().numpy()
            if (res[:, 0] <= 0).any():
                bad = int(np.argmax(res[:, 0] <= 0))
                raise _lib.McrError("device track generation failed for slot %d: code %d (-3: more tiles than max_tiles, "
                                    "-4: more quads than max_quads, -5: no valid track in 64 attempts)" % (slots[bad], res[bad, 0]))
            mt_out = d_mt.cpu().numpy().view(np.uint32)
            # path nodes (alpha, beta, x, y) of every track: rows [i1, i1 + T) of its scratch block
            path = d_scratch[:, :2600 * 32].view(torch.float64).view(n, 2600, 4)
            rows = (d_res[:, 2].long().view(n, 1) + torch.arange(self.max_tiles, device=self.device).view(1, -1)).clamp_(max=2599)
            nodes = torch.gather(path, 1, rows.view(n, -1, 1).expand(n, self.max_tiles, 4)).cpu().numpy()
        for i, rng in enumerate(rngs):
            st = states[i]
            rng.set_state((st[0], mt_out[i, :624].copy(), int(mt_out[i, 624]), st[3], st[4]))
            T = int(res[i, 0])
            self.tracks[slots[i]] = DeviceTrack(self, slots[i], nodes[i, :T].copy(), (int(res[i, 2]), int(res[i, 3])), int(res[i, 1]))
        if self.verbose == 1:
            for i in range(n):
                print("Track generation: %i..%i -> %i-tiles track" % (res[i, 2], res[i, 3], res[i, 3] - res[i, 2]))
        return res

    # ---- step, reference :410-509 ------------------------------------------------------------------
    def step(self, action):
        """action: (B, A, 3) torch tensor on this device (float32 or float64), or anything
        np.reshape can bring to that shape.  Returns (obs u8 (B,A,96,96,3), reward f64 (B,A),
        done u8 (B,) [bit0 done, bit1 TimeLimit], {}) -- views of buffers reused every step."""
        torch = _torch()
        B, A = self.batch_envs, self.num_agents
        if not isinstance(action, torch.Tensor):
            action = torch.from_numpy(np.ascontiguousarray(np.reshape(action, (B, A, -1)))).to(self.device)
        if action.dtype not in (torch.float32, torch.float64):
            action = action.to(torch.float32)
        action = action.reshape(B, A, -1)
        if action.shape[-1] != 3 or action.device != self.device:
            raise ValueError("action must have shape (batch_envs, num_agents, 3) on %s" % self.device)
        action = action.contiguous()
        dt = _lib.MCR_F32 if action.dtype == torch.float32 else _lib.MCR_F64
        with torch.cuda.device(self.device):
            _lib.check(self.L.mcr_step(self._h, action.data_ptr(), dt, self.obs.data_ptr(), self.reward_out.data_ptr(),
                                       self.done_out.data_ptr(), self._step_flags, self._stream()), "mcr_step")
        return self.obs, self.reward_out, self.done_out, {}

    def stacked_obs(self):
        """obs_format='gray_stack': the ring (B, A, K, 96, 96) gathered oldest -> newest.  The store writes the
        frame of an episode's step s into slot s % K (every slot on the first frame of an episode), so slot
        (s - K + 1 + j) % K holds the j-th oldest frame."""
        torch = _torch()
        if self.obs_format != "gray_stack":
            raise ValueError("stacked_obs() needs obs_format='gray_stack'")
        B, A, K = self.batch_envs, self.num_agents, self.frame_stack
        s = self.buffers["steps"].view(B, A).long()
        idx = (s.unsqueeze(-1) - (K - 1) + torch.arange(K, device=self.device).view(1, 1, K)) % K
        return torch.gather(self.obs, 2, idx.view(B, A, K, 1, 1).expand(B, A, K, STATE_H, STATE_W))

    def step_host(self, action):
        """The reference-facing call with HOST buffers: `action` is a (B, A, 3) float32/float64
        numpy array (ideally a view of pinned memory, see host_buffers()); returns numpy views
        of pinned host buffers (obs, reward, done) valid until the next call.  Host->device and
        device->host copies and the synchronisation happen inside this call."""
        torch = _torch()
        B, A = self.batch_envs, self.num_agents
        hb = self.host_buffers()
        a = np.reshape(action, (B, A, 3))
        if a.dtype == np.float64:
            hb["action64"].numpy()[...] = a
            src = hb["action64"]
        else:
            if a.ctypes.data != hb["action"].data_ptr():
                hb["action"].numpy()[...] = a
            src = hb["action"]
        dt = _lib.MCR_F64 if src.dtype == torch.float64 else _lib.MCR_F32
        with torch.cuda.device(self.device):
            # one native call: action in, step, and the results out in ranges of envs as they are rendered
            _lib.check(self.L.mcr_step_host(self._h, src.data_ptr(), dt, self.obs.data_ptr(), self.reward_out.data_ptr(),
                                            self.done_out.data_ptr(), hb["obs"].data_ptr(), hb["reward"].data_ptr(),
                                            hb["done"].data_ptr(), self._step_flags, self._stream()), "mcr_step_host")
            torch.cuda.current_stream(self.device).synchronize()
        return hb["obs"].numpy(), hb["reward"].numpy(), hb["done"].numpy(), {}

    # ---- pipelined host API: the copy of step k's results overlaps the compute of step k+1 -------------
    def step_host_async(self, action):
        """Enqueue one step whose results go to pinned host memory, without waiting for them: host ->
        device copy of the action and the step on the compute stream, device -> host copies of
        (obs, reward, done) on a copy stream behind it.  Two steps may be in flight (double-buffered
        device and host result buffers); step_host_wait() returns the oldest one.  For consumers whose
        next action does not depend on the newest observation (asynchronous / delayed policies, data
        collection): throughput is then bound by the PCIe link alone instead of copy + compute."""
        torch = _torch()
        B, A = self.batch_envs, self.num_agents
        if getattr(self, "_pipe", None) is None:
            with torch.cuda.device(self.device):
                self._pipe = [dict(
                    obs=torch.zeros((B, A) + self.obs_shape, dtype=self.obs_dtype, device=self.device),
                    reward=torch.zeros((B, A), dtype=torch.float64, device=self.device),
                    done=torch.zeros((B,), dtype=torch.uint8, device=self.device),
                    h_obs=torch.zeros((B, A) + self.obs_shape, dtype=self.obs_dtype).pin_memory(),
                    h_reward=torch.zeros((B, A), dtype=torch.float64).pin_memory(),
                    h_done=torch.zeros((B,), dtype=torch.uint8).pin_memory(),
                    # per-slot action staging: the host buffer of a slot is only rewritten after the step that
                    # read it was retired by step_host_wait(), and the device copy is this slot's own
                    h_action=torch.zeros((B, A, 3), dtype=torch.float32).pin_memory(),
                    h_action64=torch.zeros((B, A, 3), dtype=torch.float64).pin_memory(),
                    d_action=torch.zeros((B, A, 3), dtype=torch.float32, device=self.device),
                    d_action64=torch.zeros((B, A, 3), dtype=torch.float64, device=self.device),
                    computed=torch.cuda.Event(), copied=torch.cuda.Event(), busy=False) for _ in range(2)]
                self._pipe_copy_stream = torch.cuda.Stream(device=self.device)
            self._pipe_issue, self._pipe_retire = 0, 0
        if self._pipe_issue - self._pipe_retire >= 2:
            raise RuntimeError("two steps are already in flight: call step_host_wait() first")
        a = np.reshape(action, (B, A, 3))
        p = self._pipe[self._pipe_issue & 1]
        with torch.cuda.device(self.device):
            cur = torch.cuda.current_stream(self.device)
            # slot p was last used two calls ago and has been retired since (the in-flight check above), so its
            # host -> device action copy has completed: the pinned buffer can be rewritten
            if a.dtype == np.float64:
                p["h_action64"].numpy()[...] = a
                src, dev_a, dt = p["h_action64"], p["d_action64"], _lib.MCR_F64
            else:
                p["h_action"].numpy()[...] = a
                src, dev_a, dt = p["h_action"], p["d_action"], _lib.MCR_F32
            if p["busy"]:
                cur.wait_event(p["copied"])            # the buffers' previous results have reached the host
            dev_a.copy_(src, non_blocking=True)
            _lib.check(self.L.mcr_step(self._h, dev_a.data_ptr(), dt, p["obs"].data_ptr(), p["reward"].data_ptr(),
                                       p["done"].data_ptr(), self._step_flags, self._stream()), "mcr_step")
            p["computed"].record(cur)
            cs = self._pipe_copy_stream
            cs.wait_event(p["computed"])
            with torch.cuda.stream(cs):
                p["h_obs"].copy_(p["obs"], non_blocking=True)
                p["h_reward"].copy_(p["reward"], non_blocking=True)
                p["h_done"].copy_(p["done"], non_blocking=True)
                p["copied"].record(cs)
            p["busy"] = True
        self._pipe_issue += 1

    def step_host_wait(self):
        """Results of the oldest step in flight: numpy views of pinned host buffers (valid until the
        second step_host_async() call from now)."""
        if getattr(self, "_pipe", None) is None or self._pipe_issue == self._pipe_retire:
            raise RuntimeError("step_host_wait() without a step in flight")
        p = self._pipe[self._pipe_retire & 1]
        p["copied"].synchronize()
        self._pipe_retire += 1
        return p["h_obs"].numpy(), p["h_reward"].numpy(), p["h_done"].numpy(), {}

    def host_buffers(self):
        """Pinned host staging buffers of step_host (allocated on first use)."""
        if getattr(self, "_host", None) is None:
            torch = _torch()
            B, A = self.batch_envs, self.num_agents
            self._host = dict(
                action=torch.zeros((B, A, 3), dtype=torch.float32).pin_memory(),
                action64=torch.zeros((B, A, 3), dtype=torch.float64).pin_memory(),
                obs=torch.zeros((B, A) + self.obs_shape, dtype=self.obs_dtype).pin_memory(),
                reward=torch.zeros((B, A), dtype=torch.float64).pin_memory(),
                done=torch.zeros((B,), dtype=torch.uint8).pin_memory())
            # the library's own action staging buffer (what its CUDA graph reads): copying the host
            # action straight into it saves mcr_step's device-to-device copy
            stage = self.buffers["action_stage"]
            self._dev_action64 = stage.view(B, A, 3)
            self._dev_action = stage.view(torch.float32).reshape(-1)[:B * A * 3].view(B, A, 3)
        return self._host

    def step_split(self, action, events=None, fused=True):
        """mcr_step without auto reset, issued through the split entry points; `events`
        (torch.cuda.Event list) are recorded around the launches (bench.py's per-kernel timing):
        fused=True  -> [simulate, render]            (3 events; what mcr_step launches)
        fused=False -> [contacts, physics, render]   (4 events).  Identical results to step()."""
        torch = _torch()
        dt = _lib.MCR_F32 if action.dtype == torch.float32 else _lib.MCR_F64
        st = self._stream()
        cur = torch.cuda.current_stream(self.device)
        k = 0
        if events: events[k].record(cur); k += 1
        if fused:
            _lib.check(self.L.mcr_simulate(self._h, None, action.data_ptr(), dt, st), "mcr_simulate")
        else:
            _lib.check(self.L.mcr_contacts(self._h, None, st), "mcr_contacts")
            if events: events[k].record(cur); k += 1
            _lib.check(self.L.mcr_physics(self._h, None, action.data_ptr(), dt, st), "mcr_physics")
        if events: events[k].record(cur); k += 1
        _lib.check(self.L.mcr_render(self._h, None, self.obs.data_ptr(), self.reward_out.data_ptr(),
                                     self.done_out.data_ptr(), 1, st), "mcr_render")
        if events: events[k].record(cur)
        return self.obs, self.reward_out, self.done_out, {}

    def render(self, mode='state_pixels'):
        """render(mode) between steps, reference :511-604: 'state_pixels' -> (B, A, 96, 96, 3),
        'rgb_array' -> (B, A, 400, 600, 3) uint8 on the device, showing the envs as they are now
        (current score and backward flags).  Same rasteriser, tiled over the larger viewport; skid
        particles are drawn when the env was built with particles=True.  'human' opens windows in the
        reference; there is no display here."""
        assert mode in ['human', 'state_pixels', 'rgb_array']
        if mode == 'human':
            raise NotImplementedError("mode='human' needs a display; use 'rgb_array' (same picture at 600x400)")
        vw, vh = (STATE_W, STATE_H) if mode == 'state_pixels' else (VIDEO_W, VIDEO_H)
        return self.render_viewport(vw, vh)

    def render_viewport(self, vw, vh):
        """The render() picture for any glViewport size (width a multiple of 4): (B, A, vh, vw, 3) uint8."""
        torch = _torch()
        cache = self.__dict__.setdefault("_render_buffers", {})
        out = cache.get((vw, vh))
        if out is None:
            out = cache[(vw, vh)] = torch.zeros((self.batch_envs, self.num_agents, vh, vw, 3), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.L.mcr_render_viewport(self._h, None, int(vw), int(vh), out.data_ptr(), self._stream()), "mcr_render_viewport")
        return out

    # ---- split entry points (bench / ncu / tests) ------------------------------------------------
    def contacts_only(self):
        _lib.check(self.L.mcr_contacts(self._h, None, self._stream()), "mcr_contacts")

    def physics_only(self, action):
        dt = _lib.MCR_F32 if action.dtype == _torch().float32 else _lib.MCR_F64
        _lib.check(self.L.mcr_physics(self._h, None, action.data_ptr(), dt, self._stream()), "mcr_physics")

    def simulate_only(self, action):
        dt = _lib.MCR_F32 if action.dtype == _torch().float32 else _lib.MCR_F64
        _lib.check(self.L.mcr_simulate(self._h, None, action.data_ptr(), dt, self._stream()), "mcr_simulate")

    def render_only(self):
        _lib.check(self.L.mcr_render(self._h, None, self.obs.data_ptr(), None, None, 0, self._stream()), "mcr_render")

    # ---- state views for tests / users ----------------------------------------------------------
    def bodies(self):
        """(B, A, 5, 10) float32: cx cy angle vx vy w px py sin cos for hull + 4 wheels."""
        b = self.buffers["body"]                      # (5, 10, N)
        return b.permute(2, 0, 1).reshape(self.batch_envs, self.num_agents, 5, 10)

    STATUS_NAMES = ("contact event overflow", "NaN in a body state", "rasteriser overflow", "car-car manifold overflow")

    def status(self):
        """Sticky device status words (events dropped, NaN, rasteriser overflow, manifold overflow): all zero
        means every contact / manifold / polygon of every step so far was processed."""
        return self.buffers["status"].cpu().numpy()

    def check_status(self):
        """Raise if a kernel ever had to drop work (results then differ from the reference).  One small
        device -> host read: call it at episode or logging boundaries, not per step.  close() warns."""
        st = self.status()
        if st.any():
            raise _lib.McrError("device status words are set: " + ", ".join(
                "%s (%d)" % (n, v) for n, v in zip(self.STATUS_NAMES, st.tolist()) if v))

    def mass(self):
        out = np.empty(12, np.float32)
        _lib.check(self.L.mcr_get_mass(self._h, out.ctypes.data), "mcr_get_mass")
        return out

    def shape(self, which):
        out = np.empty(16, np.float32)
        n = _lib.check(self.L.mcr_get_shape(self._h, which, out.ctypes.data), "mcr_get_shape")
        return out[:2 * n].reshape(n, 2).copy()


class _TrackTable:
    """venv.tracks[slot]: the HostTrack loaded into a pool slot, or a DeviceTrack view of whatever the GPU generated
    there (read back on demand -- device-side resets never touch the host)."""

    def __init__(self, venv):
        self._venv, self.host = venv, {}

    def __getitem__(self, slot):
        slot = int(slot)
        t = self.host.get(slot)
        v = self._venv
        if t is not None and v.fresh_tracks and int(v.buffers["trk_consumed"][slot % v.batch_envs].item()) > 0:
            t = None                     # the env's ring has moved on: the slot may hold a device-generated track by now
        return t if t is not None else DeviceTrack(v, slot)

    def __setitem__(self, slot, track):
        self.host[int(slot)] = track

    def __len__(self):
        return self._venv.pool_tracks

    def __iter__(self):
        return (self[s] for s in range(len(self)))


class DeviceTrack:
    """A track generated on the GPU: nodes (alpha, beta, x, y) float64 on the host; the road_poly
    quads are read back from the pool on demand (fp32 -- what GL and Box2D see of them)."""

    def __init__(self, venv, slot, nodes=None, idx_range=None, attempts=None):
        self._venv, self.slot, self.idx_range, self.attempts = venv, slot, idx_range, attempts
        if nodes is None:
            # from the pool: (beta, x, y) per tile; alpha (only used while generating) is not kept there
            T = int(venv.buffers["trk_T"][slot].item())
            bxy = venv.buffers["trk_node"][slot, :T].cpu().numpy()
            nodes = np.concatenate([np.full((T, 1), np.nan), bxy], axis=1)
        self.nodes = nodes

    @property
    def T(self):
        return len(self.nodes)

    @property
    def Q(self):
        return int(self._venv.buffers["trk_Q"][self.slot].item())

    @property
    def quads(self):
        return self._venv.buffers["trk_quad"][self.slot, :self.Q].cpu().numpy().astype(np.float64).reshape(-1, 4, 2)

    @property
    def quad_tile(self):
        return self._venv.buffers["trk_quad_tile"][self.slot, :self.Q].cpu().numpy().astype(np.int32)

    @property
    def quad_pal(self):
        return self._venv.buffers["trk_quad_col"][self.slot, :self.Q].cpu().numpy()

    @property
    def quad_rgb(self):
        """road_poly colours as _create_track leaves them (reference :316-317, :331-332), float32."""
        c = [np.float32(0.4 + 0.01 * i) for i in range(3)]
        table = {3: (c[0],) * 3, 4: (c[1],) * 3, 5: (c[2],) * 3, 6: (1.0, 1.0, 1.0), 7: (1.0, 0.0, 0.0)}
        return np.array([table[int(p)] for p in self.quad_pal], np.float32)


class _Vec2(tuple):
    @property
    def x(self):
        return self[0]

    @property
    def y(self):
        return self[1]


class _JointView:
    def __init__(self, env, car, wheel):
        self._e, self._c, self._w = env, car, wheel

    @property
    def angle(self):
        b = self._e._batch.bodies()[0, self._c].cpu().numpy()
        return float(np.float32(b[1 + self._w, 2]) - np.float32(b[0, 2]))


class _WheelView:
    def __init__(self, env, car, wheel):
        self._e, self._c, self._w = env, car, wheel
        self.joint = _JointView(env, car, wheel)
        self.car_id = car

    @property
    def omega(self):
        return float(self._e._batch.buffers["wheel"][self._w, 0, self._c].item())

    @property
    def phase(self):
        return float(self._e._batch.buffers["wheel"][self._w, 1, self._c].item())


class _HullView:
    def __init__(self, env, car):
        self._e, self._c = env, car

    def _row(self):
        return self._e._batch.bodies()[0, self._c, 0].cpu().numpy()

    @property
    def position(self):
        r = self._row()
        return _Vec2((float(r[6]), float(r[7])))

    @property
    def angle(self):
        return float(self._row()[2])

    @property
    def linearVelocity(self):
        r = self._row()
        return _Vec2((float(r[3]), float(r[4])))

    @property
    def angularVelocity(self):
        return float(self._row()[5])


class _CarView:
    def __init__(self, env, car):
        self.hull = _HullView(env, car)
        self.wheels = [_WheelView(env, car, w) for w in range(4)]


class MultiCarRacing:
    """Single-env drop-in for the reference class (same kwargs/defaults, reference :131-133).

    step()/reset() return host numpy arrays exactly shaped like the reference's
    (reference :408, :509, :518).  `device` is the only extra kwarg.
    """
    metadata = {'render.modes': ['human', 'rgb_array', 'state_pixels'], 'video.frames_per_second': FPS}

    def __init__(self, num_agents=2, verbose=1, direction='CCW', use_random_direction=True,
                 backwards_flag=True, h_ratio=0.25, use_ego_color=False, device=None,
                 max_tiles=MAX_TILES_DEFAULT, max_quads=MAX_QUADS_DEFAULT, collisions=True):
        self.num_agents = num_agents
        self.verbose = verbose
        self.use_random_direction = use_random_direction
        self.episode_direction = direction
        if self.use_random_direction:                       # reference :156-157 (global numpy RNG)
            self.episode_direction = str(np.random.choice(['CW', 'CCW']))
        self.backwards_flag, self.h_ratio, self.use_ego_color = backwards_flag, h_ratio, use_ego_color
        self._batch = BatchedMultiCarRacing(1, num_agents=num_agents, verbose=verbose, direction=direction,
                                            use_random_direction=use_random_direction, backwards_flag=backwards_flag,
                                            h_ratio=h_ratio, use_ego_color=use_ego_color, device=device,
                                            max_tiles=max_tiles, max_quads=max_quads, pool_tracks=1,
                                            max_episode_steps=0, auto_reset=False, collisions=collisions, particles=True)
        self.action_space = self._batch.action_space
        self.observation_space = self._batch.observation_space
        self.car_order = None
        self.track = None
        self.road_poly = []
        self.state = None
        self._has_reset = False
        self.cars = [_CarView(self, c) for c in range(num_agents)]
        self.seed()

    def seed(self, seed=None):
        self.np_random, seed = np_random(seed)
        self._batch.np_randoms = [self.np_random]
        return [seed]

    def reset(self):
        if self.use_random_direction:                       # reference :351-352
            self.episode_direction = str(np.random.choice(['CW', 'CCW']))
        ids = [i for i in range(self.num_agents)]           # reference :355-357
        shuffle_ids = np.random.choice(ids, size=self.num_agents, replace=False)
        self.car_order = {i: shuffle_ids[i] for i in range(self.num_agents)}
        tr = self._batch._gen.generate(self.np_random, self.verbose)
        self.track = [tuple(r) for r in tr.nodes]
        self.road_poly = [([tuple(p) for p in q], tuple(c)) for q, c in zip(tr.quads, tr.quad_rgb)]
        obs = self._batch.reset(tracks=[tr], car_orders=[shuffle_ids], directions=[self.episode_direction])
        self._has_reset = True
        self.state = obs[0].cpu().numpy()
        return self.state

    def step(self, action):
        if action is None:
            raise ValueError("step(None) is reset()'s internal call (reference :408); call reset()")
        action = np.reshape(action, (self.num_agents, -1))   # reference :420 (raises on a wrong size)
        if action.shape[1] != 3:
            raise ValueError("cannot reshape action into (num_agents, 3)")
        if action.dtype != np.float32:
            action = action.astype(np.float64)
        obs, rew, done, _ = self._batch.step(action.reshape(1, self.num_agents, 3))
        self.state = obs[0].cpu().numpy()
        return self.state, rew[0].cpu().numpy(), bool(done[0].item() & 1), {}

    def render(self, mode='human'):
        assert mode in ['human', 'state_pixels', 'rgb_array']
        if not self._has_reset:
            return None                                      # reference :538 ("reset() not called yet")
        return self._batch.render(mode)[0].cpu().numpy()

    def close(self):
        pass

    # attributes users of the reference poke at
    @property
    def reward(self):
        return self._batch.buffers["reward"].cpu().numpy().copy()

    @property
    def prev_reward(self):
        return self._batch.buffers["prev_reward"].cpu().numpy().copy()

    @property
    def tile_visited_count(self):
        return [int(v) for v in self._batch.buffers["visit_count"].cpu().numpy()]

    @property
    def driving_backward(self):
        return self._batch.buffers["backward"].cpu().numpy().astype(bool)

    @property
    def driving_on_grass(self):
        """reference :153, :469-472: the hull position is inside none of the road_poly polygons"""
        return self._batch.buffers["on_grass"].cpu().numpy().astype(bool)

    @property
    def t(self):
        return float(self._batch.buffers["time"][0].item())


class MultiCarRacingVecEnv:
    """Vector-env adapter over BatchedMultiCarRacing for learners that speak the
    gym(nasium) VectorEnv protocol: `num_envs`, `single_observation_space`, `single_action_space`,
    `reset(seed=None) -> (obs, info)`, `step(actions) -> (obs, reward, terminated, truncated, info)`
    and the older `step_async` / `step_wait` pair.  Tensors stay on the device; an env that ended
    (terminated: mcr:497-507, truncated: the registration's TimeLimit) is reset by the NEXT step
    (gymnasium's NEXT_STEP autoreset mode == mcr_step flag bit1).  One "env" of the vector is one
    MultiCarRacing-v0 instance with `num_agents` cars: observations (num_envs, A, ...), actions
    (num_envs, A, 3), rewards (num_envs, A)."""

    def __init__(self, num_envs, num_agents=2, device=None, seed=None, **kwargs):
        kwargs.setdefault("max_episode_steps", 1000)
        self._device_tracks = kwargs.pop("device_tracks", False)
        self.venv = BatchedMultiCarRacing(num_envs, num_agents=num_agents, device=device, auto_reset='next_step',
                                          seed=seed, **kwargs)
        self.num_envs, self.num_agents = int(num_envs), int(num_agents)
        A = self.num_agents
        self.single_action_space = _make_box(np.tile(np.array([-1, 0, 0], np.float32), (A, 1)),
                                             np.tile(np.array([+1, +1, +1], np.float32), (A, 1)), dtype=np.float32)
        hi, odt = (1.0, np.float16) if self.venv.obs_format == "rgb_chw_f16" else (255, np.uint8)
        self.single_observation_space = _make_box(0, hi, shape=(A,) + self.venv.obs_shape, dtype=odt)
        self.action_space = _make_box(np.tile(np.array([-1, 0, 0], np.float32), (self.num_envs, A, 1)),
                                      np.tile(np.array([+1, +1, +1], np.float32), (self.num_envs, A, 1)), dtype=np.float32)
        self.observation_space = _make_box(0, hi, shape=(self.num_envs, A) + self.venv.obs_shape, dtype=odt)
        self.metadata = {'render.modes': ['rgb_array', 'state_pixels'], 'video.frames_per_second': FPS,
                         'autoreset_mode': 'next_step'}
        self._pending_actions = None

    def reset(self, seed=None, options=None):
        if seed is not None:
            self.venv.seed(seed)
        return self.venv.reset(device_tracks=self._device_tracks), {}

    def step(self, actions):
        obs, reward, done, _ = self.venv.step(actions)
        terminated = (done & 1).bool()
        # gym's TimeLimit sets TimeLimit.truncated = not done: an env that terminated on its last allowed
        # step is terminated, not truncated
        return obs, reward, terminated, (done & 2).bool() & ~terminated, {}

    def step_async(self, actions):
        self._pending_actions = actions

    def step_wait(self):
        if self._pending_actions is None:
            raise RuntimeError("step_wait() without step_async()")
        actions, self._pending_actions = self._pending_actions, None
        obs, reward, terminated, truncated, info = self.step(actions)
        return obs, reward, terminated | truncated, info      # gym 0.17 VectorEnv.step_wait -> (obs, rews, dones, infos)

    def render(self, mode='rgb_array'):
        return self.venv.render(mode)

    def close(self):
        self.venv.close()


class TimeLimit:
    """gym.wrappers.TimeLimit semantics (gym 0.17.2) for make(): done at max_episode_steps with
    info['TimeLimit.truncated'] = not done (reference __init__.py:5-10)."""

    def __init__(self, env, max_episode_steps):
        self.env = env
        self._max_episode_steps = max_episode_steps
        self._elapsed_steps = None

    def __getattr__(self, name):
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env

    def reset(self, **kw):
        self._elapsed_steps = 0
        return self.env.reset(**kw)

    def step(self, action):
        assert self._elapsed_steps is not None, "Cannot call env.step() before calling reset()"
        obs, rew, done, info = self.env.step(action)
        self._elapsed_steps += 1
        if self._elapsed_steps >= self._max_episode_steps:
            info['TimeLimit.truncated'] = not done
            done = True
        return obs, rew, done, info
