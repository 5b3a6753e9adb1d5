/*
 * mcr.h -- C-ABI of libmcr.so, the B200-native batched MultiCarRacing-v0 step+render path.
 *
 * The reference (igilitschenski/multi_car_racing) is pure Python and has no FFI of its own;
 * the boundary it exposes is gym.Env.  Each entry point below names the reference interface
 * it replaces ("mcr" = gym_multi_car_racing/multi_car_racing.py in the reference tree).
 * INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C, no torch types.  Every pointer named d_* is a CUDA DEVICE pointer owned by the
 *     caller (the Python host allocates them as torch tensors); h_* are HOST pointers.
 *   - every device entry point only enqueues work on `stream` (a cudaStream_t passed as
 *     void*) and returns; nothing synchronises unless stated.
 *   - return value: 0 = OK, < 0 = error; mcr_last_error() returns a static description.
 *   - one handle per device; a handle is not thread-safe.
 *   - state lives in caller-owned SoA buffers described by mcr_buffer_spec(); the host binds
 *     them once with mcr_bind_buffer().  Layouts are documented in DESIGN.md §3.
 */
#ifndef MCR_H
#define MCR_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCR_ABI_VERSION 3
#define MCR_STATE_W 96
#define MCR_STATE_H 96
#define MCR_MAX_AGENTS 16
#define MCR_OBS_BYTES (MCR_STATE_W * MCR_STATE_H * 3)

typedef struct mcr_handle_t* mcr_handle;

/* Constructor kwargs of MultiCarRacing.__init__ (mcr:131-133) plus the batch geometry. */
typedef struct mcr_config {
    int32_t batch_envs;        /* B: independent environments on this device            */
    int32_t num_agents;        /* A: cars per environment (mcr:131 num_agents)           */
    int32_t max_tiles;         /* T_max: capacity of one track slot (tiles)              */
    int32_t max_quads;         /* Q_max: capacity of one track slot (road_poly entries)  */
    int32_t pool_tracks;       /* P: number of track slots in the device pool            */
    int32_t backwards_flag;    /* mcr:132 backwards_flag                                  */
    int32_t use_ego_color;     /* mcr:133 use_ego_color                                   */
    int32_t max_episode_steps; /* TimeLimit of the gym registration (__init__.py:8); 0=off */
    double  h_ratio;           /* mcr:132 h_ratio                                         */
    int32_t device;            /* CUDA device ordinal                                     */
    int32_t use_random_direction; /* mcr:132; only used by the device-side auto reset     */
    int32_t direction_cw;      /* mcr:131 direction == 'CW'; auto reset when not random   */
    int32_t collisions;        /* 1: car-car rigid contacts (Box2D polygon contacts); 0: cars pass through each other */
    uint64_t seed;             /* stream id of the device-side auto-reset RNG             */
    int32_t particles;         /* 1: keep the skid traces of gym car_dynamics.Car (Car.particles, "Skid trace" block of
                                  Car.step) so that the non-state render modes draw them (mcr:564); 0: no bookkeeping */
    int32_t fresh_tracks;      /* R > 0: the device-side auto reset gives every episode a NEW track generated on the env's
                                  own RandomState stream, like reset() does (mcr:359-364): env e owns pool slots
                                  e + batch_envs * j, j = 0..R (pool_tracks must be batch_envs * (R + 1)); a low-priority
                                  refill kernel keeps R tracks ready ahead of the episode in progress, and an env whose
                                  next track is not ready generates it in place (the track sequence of an env never
                                  depends on timing).  0: auto reset draws from the shared pool filled at reset(). */
} mcr_config;

/* Observation layouts the rasteriser can store (mcr_set_obs_format).  The reference returns
 * MCR_OBS_RGB_HWC (mcr:599-604); the others are the usual first stage of a learner's
 * pre-processing fused into the store, so the frame never makes an extra HBM round trip:
 *   MCR_OBS_RGB_HWC  [96][96][3] u8   (27 648 B per agent-frame, the default)
 *   MCR_OBS_GRAY     [96][96]    u8   ( 9 216 B) ITU-R 601 luma, (299 R + 587 G + 114 B + 500) / 1000
 *   MCR_OBS_RGB_CHW  [3][96][96] u8   (27 648 B) planar, what a convolution stack wants
 *   MCR_OBS_GRAY_STACK [K][96][96] u8 (K x 9 216 B) frame stack of the last K luma frames as a ring: the frame of the
 *                    episode's step s is stored in slot s % K, the first frame of an episode (s = 0) in every slot
 *                    (what gym's FrameStack wrapper does on reset); K = mcr_set_frame_stack(), default 4.  The
 *                    caller's d_obs persists between steps -- only one slot per agent is written per step.
 *   MCR_OBS_RGB_CHW_F16 [3][96][96] fp16 (55 296 B) planar, value / 255 rounded to nearest fp16: the learner's
 *                    uint8 -> float normalisation fused into the store (README.md:82-95 downstream learners) */
enum { MCR_OBS_RGB_HWC = 0, MCR_OBS_GRAY = 1, MCR_OBS_RGB_CHW = 2, MCR_OBS_GRAY_STACK = 3, MCR_OBS_RGB_CHW_F16 = 4 };

/* dtype codes used by mcr_buffer_spec */
enum { MCR_U8 = 0, MCR_I32 = 1, MCR_U32 = 2, MCR_F32 = 3, MCR_F64 = 4, MCR_I16 = 5 };

/* ---- lifetime ----------------------------------------------------------------------- */
/* replaces MultiCarRacing.__init__ (mcr:131-166): validates config, computes the rigid-body
 * constants of car_dynamics.Car.  Allocates nothing on the device. */
int mcr_create(const mcr_config* cfg, mcr_handle* out);
/* replaces MultiCarRacing.close / _destroy (mcr:173-181, 606-611) */
int mcr_destroy(mcr_handle h);
const char* mcr_last_error(void);
int mcr_abi_version(void);

/* ---- caller-owned state buffers ------------------------------------------------------- */
/* Number of named buffers the handle needs. */
int mcr_buffer_count(mcr_handle h);
/* Describe buffer i: name (static string), dtype code, rank and dims (up to 4). */
int mcr_buffer_spec(mcr_handle h, int i, const char** name, int32_t* dtype, int32_t* ndim, int64_t dims[4]);
/* Bind a device allocation (>= prod(dims)*sizeof(dtype) bytes, 16-byte aligned). */
int mcr_bind_buffer(mcr_handle h, int i, void* d_ptr);

/* ---- tracks (replaces what _create_track leaves on the env, mcr:183-338) -------------- */
/* Host-side generator: one attempt of _create_track for a legacy MT19937 RandomState whose
 * 624-word state + position the caller passes in/out (mt_state[625], mt_state[624] = pos),
 * so host RNG streams stay identical to numpy's RandomState.  Returns T (>0) on success,
 * 0 when the attempt failed (the reference returns False and retries), < 0 on error.
 * Outputs (host): nodes[T][4] (alpha,beta,x,y) f64, quads[Q][8] f64, quad_rgb[Q][3] f32,
 * quad_tile[Q] i32 (-1 = border quad), *out_q = Q, idx_range[2] = (i1, i2) of mcr:277. */
int mcr_track_generate(uint32_t* mt_state, int32_t max_tiles, int32_t max_quads, double* h_nodes,
                       double* h_quads, float* h_quad_rgb, int32_t* h_quad_tile, int32_t* out_q,
                       int32_t* idx_range);
/* MT19937 init_by_array, the seeding RandomState.seed(list) performs (gym seeding.np_random,
 * used at mcr:169-171). */
int mcr_mt_seed(uint32_t* mt_state, const uint32_t* key, int32_t key_len);
/* The same for n streams at once: states[n][625], keys[n][key_stride] of which key_len[i] words are used. */
int mcr_mt_seed_batch(uint32_t* h_states, const uint32_t* h_keys, const int32_t* h_key_len, int32_t n, int32_t key_stride);
/* reset()'s draws from the GLOBAL numpy RNG (mcr:351-357) for n envs in a row: direction
 * (np.random.choice(['CW','CCW']) when use_random_direction, else default_cw) and car order
 * (np.random.choice(ids, size=A, replace=False)).  mt_state[625] = np.random.get_state() in/out, so the
 * global stream advances exactly as n reference reset() calls would advance it.  h_cw[n], h_order[n][A]. */
int mcr_reset_draws(uint32_t* mt_state, int32_t n, int32_t A, int32_t use_random_direction, int32_t default_cw,
                    uint8_t* h_cw, int32_t* h_order);
/* Spawn grid of reset() (mcr:366-393): poses[A][3] = (angle, x, y) f64 for car_order[A]. */
int mcr_spawn_poses(const double* h_nodes, int32_t T, const int32_t* car_order, int32_t A, int32_t cw,
                    double* h_poses);
/* Upload one generated track into pool slot `slot` (synchronous host->device copies on
 * `stream`; builds the Box2D tile polygons and AABBs on the host first). */
int mcr_load_track(mcr_handle h, int32_t slot, int32_t T, const double* h_nodes, int32_t Q,
                   const double* h_quads, const float* h_quad_rgb, const int32_t* h_quad_tile, void* stream);

/* On-device batched generation (replaces the `while True: success = self._create_track()` loop of
 * reset(), mcr:359-364, for n tracks at once, one GPU thread per track): track i consumes the
 * MT19937 stream d_mt_state[i][625] (in/out, same layout as mcr_track_generate's) exactly like the
 * host generator -- same uniforms, same float64 operation order -- and is written straight into
 * pool slot d_slot[i] (tiles, AABBs, road_poly quads, palette, culling chunks, spawn grid).  CUDA's
 * fp64 sin/cos/atan2 differ from glibc's by <= 1-2 ulp, so node coordinates agree with the host
 * generator to ~1e-12 rather than bit for bit; the host generator remains the bit-exact path.
 * d_scratch: n * mcr_trackgen_scratch_bytes() bytes; afterwards track i's path nodes
 * (alpha, beta, x, y) f64 are rows [i1, i1 + T) of its scratch block.
 * d_result[i][4] = { T (> 0) or an error code (< 0), attempts, i1, i2 }.  d_slot = NULL: track i goes to slot i. */
int mcr_tracks_generate_device(mcr_handle h, int32_t n, uint32_t* d_mt_state, const int32_t* d_slot,
                               void* d_scratch, int32_t* d_result, void* stream);
int64_t mcr_trackgen_scratch_bytes(void);

/* ---- reset (replaces MultiCarRacing.reset, mcr:340-408) -------------------------------- */
/* For every env with d_env_mask[env] != 0 (NULL = all): bind track slot d_track_slot[env],
 * direction d_cw[env], create the cars at d_spawn_pose[env][A][3] (f64: angle,x,y), zero the
 * bookkeeping and run the implicit step(None) (mcr:408) writing the first observation into
 * d_obs[env].  d_track_slot/d_cw/d_spawn_pose are [B]/[B]/[B][A][3] device arrays. */
int mcr_reset(mcr_handle h, const uint8_t* d_env_mask, const int32_t* d_track_slot, const uint8_t* d_cw,
              const double* d_spawn_pose, uint8_t* d_obs, void* stream);

/* ---- step (replaces MultiCarRacing.step, mcr:410-509) ---------------------------------- */
/* d_action: [B][A][3] (steer, gas, brake), dtype f32 (action_dtype=MCR_F32) or f64.
 * d_obs: [B][A][96][96][3] u8.  d_reward: [B][A] f64.  d_done: [B] u8.
 * flags: bit0 = same-step device-side auto reset of finished envs (done or max_episode_steps
 * reached) from the track pool before returning: their d_obs is the new episode's first frame,
 * reward/done are the terminal ones (a second, masked pass of the kernels).
 * bit1 = next-step auto reset (EnvPool convention, exclusive with bit0): an env that reported
 * done at step k keeps its terminal observation; the call for step k+1 ignores that env's action,
 * respawns it and runs reset()'s implicit step(None) (mcr:408) inside the same single pass, and
 * returns the new episode's first frame with reward 0, done 0.
 * Bit1 of d_done[env] marks TimeLimit truncation.
 * d_action is read by the step's first kernel in `stream` order (no staging copy: the replayed CUDA graph's kernel node
 * is pointed at it): like any stream-ordered argument it must stay valid, and unchanged by other streams / the host,
 * until the step has run. */
int mcr_step(mcr_handle h, const void* d_action, int32_t action_dtype, uint8_t* d_obs, double* d_reward,
             uint8_t* d_done, int32_t flags, void* stream);

/* The same step with HOST buffers (what a caller holding numpy arrays uses): h_action [B][A][3] in pinned (or pageable)
 * host memory is copied in, and obs / reward / done are copied out to h_obs / h_reward / h_done (pinned) as part of the
 * call's stream work -- the frames leave in ranges of envs as soon as each range is rendered, so most of the step hides
 * behind the PCIe transfer.  d_obs / d_reward / d_done are the device staging buffers (also valid results).  Nothing
 * synchronises: the host buffers are complete once `stream` has drained. */
int mcr_step_host(mcr_handle h, const void* h_action, int32_t action_dtype, uint8_t* d_obs, double* d_reward, uint8_t* d_done,
                  uint8_t* h_obs, double* h_reward, uint8_t* h_done, int32_t flags, void* stream);

/* Split entry points (benchmarks / ncu / tests): the same results as mcr_step without auto reset, issued as
 * separate stages.  mcr_simulate = mcr_contacts (side stream) beside mcr_physics; mcr_render with
 * post_step = 1 also runs the reward / done block.  mcr_step itself issues one pipelined CUDA graph. */
int mcr_simulate(mcr_handle h, const uint8_t* d_env_mask, const void* d_action, int32_t action_dtype,
                 void* stream);                                                      /* mcr:84-123 + mcr:421-428 */
int mcr_contacts(mcr_handle h, const uint8_t* d_env_mask, void* stream);            /* FrictionDetector + b2 Collide, mcr:84-123 */
int mcr_physics(mcr_handle h, const uint8_t* d_env_mask, const void* d_action, int32_t action_dtype,
                void* stream);                                                       /* Car.step + world.Step, mcr:421-428 */
int mcr_render(mcr_handle h, const uint8_t* d_env_mask, uint8_t* d_obs, double* d_reward, uint8_t* d_done,
               int32_t post_step, void* stream);                                     /* render + mcr:433-507 */

/* render(mode) outside step(), mcr:511-604, for any glViewport size: (96, 96) = 'state_pixels',
 * (600, 400) = 'rgb_array'.  d_out: [B][A][vh][vw][3] u8.  Shows the env as it is NOW (current
 * reward in the score label, current backward flags), like a render() call between steps.  Skid
 * particles (only drawn in the non-state modes, mcr:564) are drawn when mcr_config.particles is set. */
int mcr_render_viewport(mcr_handle h, const uint8_t* d_env_mask, int32_t vw, int32_t vh, uint8_t* d_out, void* stream);

/* Select the layout every later mcr_reset / mcr_step / mcr_render call writes into d_obs (one of
 * MCR_OBS_*).  mcr_obs_bytes() = bytes per agent-frame of the current layout. */
int mcr_set_obs_format(mcr_handle h, int32_t format);
/* Ring depth K of MCR_OBS_GRAY_STACK (1..16, default 4). */
int mcr_set_frame_stack(mcr_handle h, int32_t k);
int64_t mcr_obs_bytes(mcr_handle h);

/* Car-constant readback for parity tests: 12 floats hull(mass,invMass,I,invI,lc.x,lc.y),
 * wheel(same). */
int mcr_get_mass(mcr_handle h, float* h_out12);
/* Polygon readback: which = 0..3 hull fixtures, 4 = wheel.  Returns the vertex count. */
int mcr_get_shape(mcr_handle h, int32_t which, float* h_out_xy16);
/* Number of kernels launched through this handle so far (bench.py's gpu_launches). */
int64_t mcr_launch_count(mcr_handle h);

#ifdef __cplusplus
}
#endif
#endif /* MCR_H */
