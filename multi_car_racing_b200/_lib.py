"""ctypes binding of libmcr.so (include/mcr.h).  Fails loudly when the library is missing --
there is no CPU fallback in the product path."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MCR_LIB_PATH") or os.path.join(_HERE, "libmcr.so")     # MCR_LIB_PATH: A/B another build

MCR_U8, MCR_I32, MCR_U32, MCR_F32, MCR_F64, MCR_I16 = range(6)
STATE_W = 96
STATE_H = 96
OBS_BYTES = STATE_W * STATE_H * 3
MAX_AGENTS = 16

# every symbol include/mcr.h declares
EXPORTS = [
    "mcr_create", "mcr_destroy", "mcr_last_error", "mcr_abi_version", "mcr_buffer_count",
    "mcr_buffer_spec", "mcr_bind_buffer", "mcr_track_generate", "mcr_mt_seed", "mcr_spawn_poses",
    "mcr_load_track", "mcr_reset", "mcr_step", "mcr_simulate", "mcr_contacts", "mcr_physics", "mcr_render",
    "mcr_get_mass", "mcr_get_shape", "mcr_launch_count", "mcr_set_obs_format", "mcr_obs_bytes",
    "mcr_tracks_generate_device", "mcr_trackgen_scratch_bytes", "mcr_render_viewport", "mcr_set_frame_stack", "mcr_mt_seed_batch", "mcr_reset_draws", "mcr_step_host",
]
OBS_FORMATS = {"rgb": 0, "gray": 1, "rgb_chw": 2, "gray_stack": 3, "rgb_chw_f16": 4}     # MCR_OBS_* of include/mcr.h


class McrConfig(ctypes.Structure):
    _fields_ = [
        ("batch_envs", ctypes.c_int32), ("num_agents", ctypes.c_int32),
        ("max_tiles", ctypes.c_int32), ("max_quads", ctypes.c_int32),
        ("pool_tracks", ctypes.c_int32), ("backwards_flag", ctypes.c_int32),
        ("use_ego_color", ctypes.c_int32), ("max_episode_steps", ctypes.c_int32),
        ("h_ratio", ctypes.c_double), ("device", ctypes.c_int32),
        ("use_random_direction", ctypes.c_int32), ("direction_cw", ctypes.c_int32),
        ("collisions", ctypes.c_int32), ("seed", ctypes.c_uint64), ("particles", ctypes.c_int32),
        ("fresh_tracks", ctypes.c_int32),
    ]


class McrError(RuntimeError):
    pass


_lib = None


def load():
    """Load libmcr.so.  Raises if it has not been built (python -m multi_car_racing_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise McrError(
            "libmcr.so is missing at %s -- build it with `python -m multi_car_racing_b200.build` "
            "(nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
    L.mcr_create.restype = i32
    L.mcr_create.argtypes = [ctypes.POINTER(McrConfig), ctypes.POINTER(vp)]
    L.mcr_destroy.restype = i32
    L.mcr_destroy.argtypes = [vp]
    L.mcr_last_error.restype = ctypes.c_char_p
    L.mcr_last_error.argtypes = []
    L.mcr_abi_version.restype = i32
    L.mcr_buffer_count.restype = i32
    L.mcr_buffer_count.argtypes = [vp]
    L.mcr_buffer_spec.restype = i32
    L.mcr_buffer_spec.argtypes = [vp, i32, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(i32),
                                  ctypes.POINTER(i32), ctypes.POINTER(i64 * 4)]
    L.mcr_bind_buffer.restype = i32
    L.mcr_bind_buffer.argtypes = [vp, i32, vp]
    L.mcr_track_generate.restype = i32
    L.mcr_track_generate.argtypes = [vp, i32, i32, vp, vp, vp, vp, ctypes.POINTER(i32), vp]
    L.mcr_mt_seed.restype = i32
    L.mcr_mt_seed.argtypes = [vp, vp, i32]
    L.mcr_mt_seed_batch.restype = i32
    L.mcr_mt_seed_batch.argtypes = [vp, vp, vp, i32, i32]
    L.mcr_reset_draws.restype = i32
    L.mcr_reset_draws.argtypes = [vp, i32, i32, i32, i32, vp, vp]
    L.mcr_spawn_poses.restype = i32
    L.mcr_spawn_poses.argtypes = [vp, i32, vp, i32, i32, vp]
    L.mcr_load_track.restype = i32
    L.mcr_load_track.argtypes = [vp, i32, i32, vp, i32, vp, vp, vp, vp]
    L.mcr_reset.restype = i32
    L.mcr_reset.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    L.mcr_step.restype = i32
    L.mcr_step.argtypes = [vp, vp, i32, vp, vp, vp, i32, vp]
    L.mcr_step_host.restype = i32
    L.mcr_step_host.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp, vp, i32, vp]
    L.mcr_contacts.restype = i32
    L.mcr_contacts.argtypes = [vp, vp, vp]
    L.mcr_physics.restype = i32
    L.mcr_physics.argtypes = [vp, vp, vp, i32, vp]
    L.mcr_simulate.restype = i32
    L.mcr_simulate.argtypes = [vp, vp, vp, i32, vp]
    L.mcr_render.restype = i32
    L.mcr_render.argtypes = [vp, vp, vp, vp, vp, i32, vp]
    L.mcr_get_mass.restype = i32
    L.mcr_get_mass.argtypes = [vp, vp]
    L.mcr_get_shape.restype = i32
    L.mcr_get_shape.argtypes = [vp, i32, vp]
    L.mcr_launch_count.restype = i64
    L.mcr_launch_count.argtypes = [vp]
    L.mcr_set_obs_format.restype = i32
    L.mcr_set_obs_format.argtypes = [vp, i32]
    L.mcr_set_frame_stack.restype = i32
    L.mcr_set_frame_stack.argtypes = [vp, i32]
    L.mcr_obs_bytes.restype = i64
    L.mcr_obs_bytes.argtypes = [vp]
    L.mcr_tracks_generate_device.restype = i32
    L.mcr_tracks_generate_device.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    L.mcr_trackgen_scratch_bytes.restype = i64
    L.mcr_trackgen_scratch_bytes.argtypes = []
    L.mcr_render_viewport.restype = i32
    L.mcr_render_viewport.argtypes = [vp, vp, i32, i32, vp, vp]
    if L.mcr_abi_version() != 3 and not os.environ.get("MCR_LIB_PATH"):
        raise McrError("libmcr.so ABI version mismatch")
    _lib = L
    return L


def check(rc, what=""):
    if rc < 0:
        msg = load().mcr_last_error()
        raise McrError("%s failed (%d): %s" % (what or "libmcr call", rc, msg.decode() if msg else ""))
    return rc
