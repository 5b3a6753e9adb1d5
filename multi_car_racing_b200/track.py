"""Host-side track generation and seeding for the B200 path.

The generator itself is native (mcr_track_generate in libmcr.so, C++ restatement of
MultiCarRacing._create_track, reference multi_car_racing.py:183-338); this module only moves
numpy RandomState (MT19937) streams in and out of it so that `env.seed(s)` produces the same
tracks as the reference, and implements gym 0.17.2's seeding.np_random hashing
(reference multi_car_racing.py:169-171).
"""
import ctypes
import hashlib
import os
import struct

import numpy as np

from . import _lib

MAX_TILES_DEFAULT = 512
MAX_QUADS_DEFAULT = 1024


def _bigint_from_bytes(b):
    b = b + b"\0" * (4 - len(b) % 4)
    words = struct.unpack("%dI" % (len(b) // 4), b)
    return sum(w << (32 * i) for i, w in enumerate(words))


def seed_key(seed):
    """gym.utils.seeding: create_seed -> hash_seed (sha512, 8 bytes) -> _int_list_from_bigint."""
    if seed is not None and not (isinstance(seed, (int, np.integer)) and 0 <= seed):
        raise ValueError("Seed must be a non-negative integer or omitted, not {}".format(seed))
    if seed is None:
        seed = _bigint_from_bytes(os.urandom(8))
    seed = int(seed) % 2 ** 64
    big = _bigint_from_bytes(hashlib.sha512(str(seed).encode("utf8")).digest()[:8])
    key = []
    if big == 0:
        key = [0]
    while big > 0:
        big, mod = divmod(big, 2 ** 32)
        key.append(mod)
    return key, seed


def np_random(seed=None):
    """Drop-in for gym.utils.seeding.np_random: (RandomState, seed)."""
    key, seed = seed_key(seed)
    rng = np.random.RandomState()
    rng.seed(key)
    return rng, seed


class HostTrack:
    """What _create_track leaves on the env (float64, host)."""
    __slots__ = ("nodes", "quads", "quad_rgb", "quad_tile", "idx_range", "attempts")

    def __init__(self, nodes, quads, quad_rgb, quad_tile, idx_range, attempts):
        self.nodes, self.quads, self.quad_rgb, self.quad_tile = nodes, quads, quad_rgb, quad_tile
        self.idx_range, self.attempts = idx_range, attempts

    @property
    def T(self):
        return len(self.nodes)

    @property
    def Q(self):
        return len(self.quads)


class TrackGenerator:
    """Runs the reset() retry loop (reference :359-364) on a numpy RandomState's own stream."""

    def __init__(self, max_tiles=MAX_TILES_DEFAULT, max_quads=MAX_QUADS_DEFAULT):
        self.L = _lib.load()
        self.max_tiles, self.max_quads = max_tiles, max_quads
        self._nodes = np.empty((max_tiles, 4), np.float64)
        self._quads = np.empty((max_quads, 8), np.float64)
        self._rgb = np.empty((max_quads, 3), np.float32)
        self._qt = np.empty((max_quads,), np.int32)
        self._mt = np.empty(625, np.uint32)
        self._rng_range = (ctypes.c_int32 * 2)()

    def generate(self, rng, verbose=0):
        kind, key, pos, has_gauss, cached = rng.get_state()
        assert kind == "MT19937"
        self._mt[:624] = key
        self._mt[624] = pos
        attempts = 0
        q = ctypes.c_int32(0)
        while True:
            attempts += 1
            T = self.L.mcr_track_generate(
                self._mt.ctypes.data, self.max_tiles, self.max_quads, self._nodes.ctypes.data,
                self._quads.ctypes.data, self._rgb.ctypes.data, self._qt.ctypes.data, ctypes.byref(q),
                ctypes.cast(self._rng_range, ctypes.c_void_p))
            _lib.check(T, "mcr_track_generate")
            if T > 0:
                if verbose == 1:
                    print("Track generation: %i..%i -> %i-tiles track" % (
                        self._rng_range[0], self._rng_range[1], self._rng_range[1] - self._rng_range[0]))
                break
            if verbose == 1:
                print("retry to generate track (normal if there are not many of this messages)")
        rng.set_state((kind, self._mt[:624].copy(), int(self._mt[624]), has_gauss, cached))
        Q = q.value
        return HostTrack(self._nodes[:T].copy(), self._quads[:Q].reshape(Q, 4, 2).copy(), self._rgb[:Q].copy(),
                         self._qt[:Q].copy(), (self._rng_range[0], self._rng_range[1]), attempts)

    def spawn_poses(self, nodes, car_order, cw):
        A = len(car_order)
        nodes = np.ascontiguousarray(nodes, np.float64)
        order = np.ascontiguousarray(car_order, np.int32)
        out = np.empty((A, 3), np.float64)
        _lib.check(self.L.mcr_spawn_poses(nodes.ctypes.data, len(nodes), order.ctypes.data, A, int(bool(cw)),
                                          out.ctypes.data), "mcr_spawn_poses")
        return out
