// mcr_internal.h -- layouts shared by the host API and the three kernels of libmcr.so.
// Product code: nothing here (or anywhere under multi_car_racing_b200/) touches oracle/.
#pragma once
#include <stdint.h>
#include "../../include/mcr.h"

// ---------------------------------------------------------------------------------------
// Box2D 2.3.x constants (b2Settings.h) used by the specialised solver
// ---------------------------------------------------------------------------------------
#define B2_PI 3.14159265359f
#define B2_LINEAR_SLOP 0.005f
#define B2_ANGULAR_SLOP (2.0f / 180.0f * B2_PI)
#define B2_POLYGON_RADIUS (2.0f * B2_LINEAR_SLOP)
#define B2_MAX_ANGULAR_CORRECTION (8.0f / 180.0f * B2_PI)
#define B2_MAX_TRANSLATION 2.0f
#define B2_MAX_ROTATION (0.5f * B2_PI)
#define B2_TIME_TO_SLEEP 0.5f
#define B2_LINEAR_SLEEP_TOL 0.01f
#define B2_ANGULAR_SLEEP_TOL (2.0f / 180.0f * B2_PI)
#define B2_EPS 1.1920928955078125e-7f

#define MCR_VEL_ITERS 180   // world.Step(1/FPS, 6*30, 2*30), mcr:428
#define MCR_POS_ITERS 60
#define MCR_MAXV 8
#define MCR_QUAD_CHUNK 8        // road_poly quads per culling chunk (bounding circle)
#define MCR_SCRATCH_FIELDS 101  // 31 velocity/impulse values + 4 joints x 15 constants + 10 sleep hand-over
#define MCR_MAX_MANIFOLDS 32    // touching car-car fixture pairs per env
#define MCR_MANIFOLD_WORDS 16

// body SoA: body[(b * BODY_FIELDS + f) * N + car], b = 0 hull, 1..4 wheels
enum { BF_CX = 0, BF_CY, BF_A, BF_VX, BF_VY, BF_W, BF_PX, BF_PY, BF_QS, BF_QC, BODY_FIELDS };
// joint SoA: joint[(w * JOINT_FIELDS + f) * N + car]
enum { JF_IX = 0, JF_IY, JF_IZ, JF_MOTOR, JOINT_FIELDS };
// wheel SoA (f64): wheel[(w * WHEEL_FIELDS + f) * N + car]
enum { WF_OMEGA = 0, WF_PHASE, WHEEL_FIELDS };
// ctrl SoA (f64): ctrl[f * N + car]
enum { CF_GAS = 0, CF_BRAKE, CF_STEER, CTRL_FIELDS };
enum { LIM_INACTIVE = 0, LIM_LOWER = 1, LIM_UPPER = 2, LIM_EQUAL = 3 };

// buffer ids (order == mcr_buffer_spec index)
enum {
    BUF_BODY = 0, BUF_SLEEP_TIME, BUF_AWAKE, BUF_JOINT, BUF_LIMIT_STATE, BUF_WHEEL, BUF_CTRL,
    BUF_ON_ROAD, BUF_ON_ROAD_NEXT, BUF_REWARD, BUF_PREV_REWARD, BUF_VISIT_COUNT, BUF_BACKWARD,
    BUF_TIME, BUF_STEPS, BUF_CAMERA, BUF_STRIPE, BUF_HEADING,
    BUF_ENV_TRACK, BUF_ENV_CW, BUF_ENV_EPISODE,
    BUF_VISITED, BUF_TOUCHED, BUF_RESET_MASK, BUF_STATUS, BUF_SCRATCH, BUF_MANIFOLD, BUF_N_MANIFOLD, BUF_SCORE_SNAP, BUF_BACKWARD_SNAP, BUF_PENDING, BUF_ACTION_STAGE, BUF_CAMERA_VP, BUF_TIMELINE, BUF_ON_GRASS, BUF_PRT_PTS, BUF_PRT_META, BUF_PRT_HDR, BUF_SKID_START, BUF_SKID_META,
    BUF_TRK_T, BUF_TRK_Q, BUF_TRK_NODE, BUF_TRK_TILE, BUF_TRK_TILE_AABB, BUF_TRK_QUAD, BUF_TRK_QUAD_COL,
    BUF_TRK_QUAD_TILE, BUF_TRK_SLOT_POSE, BUF_TRK_CHUNK, BUF_TRK_QUAD64,
    BUF_DL_HDR, BUF_DL_META, BUF_DL_EDGE, BUF_DL_OCT,
    BUF_MT_STATE, BUF_TRK_CONSUMED, BUF_TRK_PRODUCED, BUF_TRK_LOCK, BUF_TG_SCRATCH, BUF_READY,
    BUF_COUNT
};

// status words
enum { ST_EVENT_OVERFLOW = 0, ST_NAN = 1, ST_RASTER_OVERFLOW = 2, ST_MANIFOLD_OVERFLOW = 3, ST_TRACK_ERROR = 4, STATUS_WORDS = 8 };

// palette indices (rgb values in raster.cu)
enum {
    PAL_BLACK = 0, PAL_GRASS, PAL_GRASS_LIGHT, PAL_ROAD0, PAL_ROAD1, PAL_ROAD2, PAL_WHITE, PAL_RED,
    PAL_WHEEL_WHITE, PAL_CAR0, /* 8 car colours: PAL_CAR0 .. PAL_CAR0+7 */
    PAL_IND_BLUE = PAL_CAR0 + 8, PAL_IND_BLUE2, PAL_IND_GREEN, PAL_FLAG_BLUE, PAL_MUD, PAL_COUNT
};

struct Poly8 { int n; float x[MCR_MAXV]; float y[MCR_MAXV]; float nx[MCR_MAXV]; float ny[MCR_MAXV]; };   // vertices + unit edge normals

// car_dynamics.Car rigid-body constants, computed on the host at mcr_create (fp32, Box2D's
// ComputeMass / ResetMassData) and passed to the kernels by value.
struct CarConst {
    float hull_invMass, hull_invI, hull_lcx, hull_lcy;
    float wheel_invMass, wheel_invI;
    float anchor_x[4], anchor_y[4];       // localAnchorA on the hull
    float max_motor_torque, lower, upper;
    Poly8 hull_poly[4];
    Poly8 wheel_poly;
    float hull_mass, hull_I, wheel_mass, wheel_I;
};

struct DevBuffers {
    float* body; float* sleep_time; uint8_t* awake; float* joint; uint8_t* limit_state;
    double* wheel; double* ctrl; uint8_t* on_road; uint8_t* on_road_next;
    double* reward; double* prev_reward; int32_t* visit_count; uint8_t* backward;
    double* time; int32_t* steps;        // per car (all cars of an env hold identical values)
    float* camera;                       // [6][N] world -> viewport-pixel affine of each car's view
    float* stripe;                       // [4*2][N] wheel stripe y extents (NaN = hidden)
    double* heading;                     // [N] car_angle of mcr:449-456
    int32_t* env_track; uint8_t* env_cw; uint32_t* env_episode;
    uint32_t* visited; uint8_t* touched; uint8_t* reset_mask; int32_t* status;
    float* scratch;                      // [MCR_SCRATCH_FIELDS][N] solver hand-over between pre/sweep/post kernels
    float* manifold;                     // [B][MCR_MAX_MANIFOLDS][MCR_MANIFOLD_WORDS] persistent car-car contact manifolds
    int32_t* n_manifold;                 // [B]
    double* score_snap; uint8_t* backward_snap;   // [N] env.reward / driving_backward as the render of this step sees them
    uint8_t* pending;                    // [B] done flags of the previous step (next-step auto reset)
    double* action_stage;                // [N][3] f64-sized staging copy of the step's action (CUDA-graph replay reads it)
    float* camera_vp;                    // [6][N] camera affine of the last mcr_render_viewport call
    uint8_t* on_grass;                   // [N] driving_on_grass, mcr:469-472
    float* prt_pts;                      // [PRT_MAX][PRT_PTS][2][N] skid-trace points (ring slot, point, xy)
    int32_t* prt_meta;                   // [PRT_MAX][N] length | grass << 8 of the particle in a ring slot
    int32_t* prt_hdr;                    // [2][N] ring head (oldest particle), particle count
    float* skid_start;                   // [4][2][N] wheel.skid_start
    int32_t* skid_meta;                  // [4][N] see above
    unsigned long long* timeline;        // [TL_COUNT] %globaltimer stamps (ns) of the last step's kernels, see TL_*
    int32_t* trk_T; int32_t* trk_Q; double* trk_node; float* trk_tile; float* trk_tile_aabb;
    float* trk_quad; uint8_t* trk_quad_col; int16_t* trk_quad_tile; double* trk_slot_pose;
    float* trk_chunk;                    // [P][Qmax/8][4] bounding circle (cx, cy, r, 0) of 8 consecutive road_poly quads
    double* trk_quad64;                  // [P][Qmax][8] road_poly vertices in float64 (what shapely's polygons hold, mcr:336-337)
    // per-frame display lists, project_kernel -> fill_kernel (raster.cu): painter-ordered polygons that can fill a pixel
    int32_t* dl_hdr;                     // [N][4] entries, span slots, "playfield covers the frame", score glyphs (4 x i8)
    uint32_t* dl_meta;                   // [N][dl_cap][2] y0 | rows << 8 | palette << 16 | ne << 24 ; first span slot
    float* dl_edge;                      // [N][dl_cap][4][4] canonical edges (ax, ay, by, slope)
    float* dl_oct;                       // [N][A][4][4] edges 4..7 of car c's hull octagon
    // fresh track per episode on the device (mcr_config.fresh_tracks = R > 0): env e's own MT19937 stream and its ring of
    // R + 1 pool slots e + B * j (trackgen.cuh)
    uint32_t* mt_state;                  // [B][625] numpy RandomState of every env (624 words + position)
    int32_t* trk_consumed; int32_t* trk_produced; int32_t* trk_lock;   // [B]
    unsigned char* tg_scratch;           // [B][mcr_trackgen_scratch_bytes()] generator scratch (one element when fresh_tracks = 0)
    // per-car / per-frame hand-off flags of the step's critical chain (release / acquire at GPU scope): the consumer kernel
    // is launched early (programmatic dependent launch, the producer triggers at its start) and each of its CTAs only
    // waits for the cars / the frame it reads, not for the producer's slowest CTA:
    //   ready[car]           == ready[READY_EPOCH]: post_kernel stored car's final pose, camera and snapshots in this pass
    //                        (contacts_kernel counts the passes; several CTAs read the flag, nobody takes it back)
    //   ready[N + f]         project_kernel stored frame f's display list (taken back by the fill_kernel CTA that consumed it)
    //   ready[2N + car]      sweep_kernel stored car's solved velocities      } set only inside mcr_step's pipeline, taken back
    //   ready[3N + car]      contacts_kernel is done with car's env           } by the post_kernel lane pair of the car
    //   ready[(4+k)N + car]  stripe_kernel stored the stripe of wheel k       }
    // A consumer only triggers the launch of ITS consumer after its own waits: whatever a waiting grid depends on is
    // complete by the time it can be placed, so a grid that does not fit the GPU never holds back its own producers.
    int32_t* ready;
};

// half extents, in the hull frame and about the hull origin, of a rectangle that holds every fixture point of a car with
// 0.5 m to spare (joint slack, polygon radii, rounding): head_kernel's narrow-phase prefilter; checked at mcr_create
#define CAR_OBB_EX 2.21f
#define CAR_OBB_EY 3.1f
#define READY_EPOCH(N) (8 * (size_t)(N))
#define READY_WORDS(N) (8 * (size_t)(N) + 32)
struct Dims { int B, A, N, Tmax, Qmax, P; int particles; int dl_cap; };
// display-list capacity per frame: every candidate of a state frame (playfield, 100 checker squares, road_poly, 12 parts per car, 9 HUD polygons)
#define MCR_DL_CAP(Qmax, A) ((1 + 100 + (Qmax) + 12 * (A) + 9 + 7) & ~7)
// skid traces (gym car_dynamics Car.particles): per car a ring of PRT_MAX polylines of <= PRT_PTS wheel positions
#define PRT_MAX 30
#define PRT_PTS 30
// skid_meta[wheel] bits: 0 skid_start valid, 1 skid_particle valid, 2 its grass flag, 8-15 its length, 16-23 its ring slot + 1 (0 = popped from Car.particles)
// timeline slots: kernel start stamps (block 0, thread 0) and the latest CTA end of the rasteriser
enum { TL_HEAD = 0, TL_CONTACTS, TL_STRIPES, TL_SWEEP, TL_COUPLED, TL_POST, TL_SCORE, TL_RENDER, TL_RENDER_END, TL_POST2, TL_RENDER2, TL_SWEEP_END, TL_SWEEP_END_PACKED, TL_COUPLED_VEL_END, TL_COUPLED_POS_END, TL_FILL = 15 /* fill_kernel start (cls != 2) */,
       TL_POST_END = 16, TL_PROJECT_END = 17, TL_HEAD_END = 18 /* latest CTA end of post (cls != 2) / project / head */,
       // hand-off diagnostics; "first" slots hold ~t (atomicMax of the complement = earliest stamp)
       TL_FILL_FIRST_IN = 19 /* first fill CTA placed */, TL_FILL_FIRST_GO = 20 /* first fill CTA past its flag */, TL_PROJECT_FIRST_END = 21,
       TL_PROJECT_LAST_GO = 22 /* last project CTA past post's flags */, TL_FILL_LAST_IN = 23 /* last fill CTA placed */, TL_COUNT = 24 };
#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long mcr_globaltimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void tl_stamp_any(const unsigned long long* tl_base, int slot) { const_cast<unsigned long long*>(tl_base)[slot] = mcr_globaltimer(); }
__device__ __forceinline__ void tl_stamp(const unsigned long long* tl_base, int slot) {
    if ((blockIdx.x | blockIdx.y | threadIdx.x) == 0) const_cast<unsigned long long*>(tl_base)[slot] = mcr_globaltimer();
}
// hand-off flags (DevBuffers::ready): release store after the producer's data stores, acquire load before the consumer's
// data loads, both at GPU scope (the acquire also drops the SM's L1 lines, so weak loads after it see the producer's data)
__device__ __forceinline__ void flag_release(int32_t* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
// polling uses relaxed loads (no L1 invalidation per poll); one acquire fence after the flags have been seen orders the
// data loads behind them and drops the SM's L1 lines once
__device__ __forceinline__ int flag_peek(const int32_t* p) { int v; asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void flag_fence_acquire() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void flag_wait(const int32_t* p, unsigned ns = 40) { while (flag_peek(p) == 0) __nanosleep(ns); flag_fence_acquire(); }
#endif

// Programmatic dependent launch for the kernels of the critical chain (head -> sweep -> post -> render): the
// dependent kernel's CTAs are launched while its predecessor in the stream is still draining and wait in
// cudaGridDependencySynchronize() (first statement of the kernel), which takes the launch latency of the big
// grids off the chain.  MCR_NO_PDL=1 falls back to plain stream order.
#ifdef __CUDACC__
#include <cstdlib>
template <typename... KArgs, typename... Args>
static inline cudaError_t mcr_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    static const bool use_pdl = std::getenv("MCR_NO_PDL") == nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = use_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

// kernel launchers (each returns the number of kernels it launched, or < 0 on error)
int launch_contacts(const Dims& d, const DevBuffers& b, const CarConst& cc, const uint8_t* mask, void* stream, int set_flags = 0);   // set_flags: see DevBuffers::ready
// noact[env] != 0: that env takes the action=None path of mcr:421 this step (next-step auto reset)
int launch_physics(const Dims& d, const DevBuffers& b, const CarConst& cc, const uint8_t* mask, const uint8_t* noact,
                   const void* action, int action_dtype, double h_ratio, int collisions, void* stream);
// wheel-stripe extents for the rasteriser (needs pre_kernel's phase; launch_physics issues it itself)
int launch_stripes(const Dims& d, const DevBuffers& b, const uint8_t* mask, void* stream, int set_flags = 0);
int launch_carcontacts(const Dims& d, const DevBuffers& b, const CarConst& cc, const uint8_t* mask, void* stream);
int launch_coupled(const Dims& d, const DevBuffers& b, const CarConst& cc, const uint8_t* mask, int early_exit, void* stream);
// cls selects envs by their car-car contact state this step: 0 = all, 1 = only envs without
// manifolds (per-car solver), 2 = only envs with manifolds (coupled_kernel) -- the two classes
// flow through post / score / render on separate streams
int launch_physics_post(const Dims& d, const DevBuffers& b, const CarConst& cc, const uint8_t* mask, const uint8_t* noact,
                        int has_action, double h_ratio, int cls, void* stream, int wait_sweep = 0);   // wait_sweep: right behind sweep_kernel, which swept every car this launch takes
int launch_presweep(const Dims& d, const DevBuffers& b, const CarConst& cc, const uint8_t* mask, const uint8_t* noact,
                    const void* action, int action_dtype, int collisions, int with_sweep, void* stream, int set_flags = 0);
// score_reward != NULL (only when render_runs_score(cls)): the reward / done block (launch_score's work) runs inside the launch
int launch_render(const Dims& d, const DevBuffers& b, const CarConst& cc, const uint8_t* mask, uint8_t* obs,
                  int backwards_flag, int use_ego_color, int cls, int obs_format, int stack_k, void* stream,
                  const uint8_t* score_noact = nullptr, double* score_reward = nullptr, uint8_t* score_done = nullptr, int max_episode_steps = 0,
                  int wait_post = 0);   // wait_post: launched right behind post_kernel in the same stream -- the projector waits for post's per-car ready flags
bool render_runs_score(int cls);
int launch_score(const Dims& d, const DevBuffers& b, const uint8_t* mask, const uint8_t* noact, double* reward, uint8_t* done,
                 int max_episode_steps, int cls, void* stream);
bool render_is_split();
int launch_project(const Dims& d, const DevBuffers& b, const CarConst& cc, const uint8_t* mask, int backwards_flag, int use_ego_color,
                   int cls, void* stream, const uint8_t* score_noact = nullptr, double* score_reward = nullptr, uint8_t* score_done = nullptr,
                   int max_episode_steps = 0, int wait_post = 0);
int launch_fill(const Dims& d, const DevBuffers& b, const uint8_t* mask, uint8_t* obs, int cls, int obs_format, int stack_k,
                int env0, int nenv, bool pdl, void* stream, int wait_flags = 0);
// render(mode) for a vw x vh viewport (rgb_array: 600 x 400): camera_kernel + tiled render_kernel<true>
int launch_render_viewport(const Dims& d, const DevBuffers& b, const CarConst& cc, const uint8_t* mask, uint8_t* out, float* cam,
                           int vw, int vh, double h_ratio, int backwards_flag, int use_ego_color, void* stream);
const uint8_t (*mcr_host_palette())[4];
int launch_trackgen(const Dims& d, const DevBuffers& b, int n, uint32_t* mt_state, const int32_t* slots, void* scratch,
                    int32_t* result, int max_attempts, void* stream);
int launch_spawn(const Dims& d, const DevBuffers& b, const CarConst& cc, const uint8_t* mask,
                 const int32_t* track_slot, const uint8_t* cw, const double* spawn_pose, void* stream);
struct AutoResetCfg { int use_random_direction, direction_cw; unsigned long long seed; int fresh; /* R: ring of R + 1 slots per env, 0 = shared pool */ };
int launch_ring_refill(const Dims& d, const DevBuffers& b, int R, void* stream);
// auto reset (reset_flags != NULL: respawn the flagged envs, write reset_mask) + carcontacts + pre in one launch
int launch_head(const Dims& d, const DevBuffers& b, const CarConst& cc, const uint8_t* mask, const uint8_t* reset_flags,
                const AutoResetCfg& ar, const void* action, int action_dtype, int collisions, void* stream);
// index of the `action` argument if func is head_kernel / pre_kernel (any instantiation), else -1; *nargs = argument count
int head_action_arg(const void* func, int* nargs);
int pre_action_arg(const void* func, int* nargs);
int launch_auto_reset(const Dims& d, const DevBuffers& b, const CarConst& cc, const uint8_t* done,
                      const AutoResetCfg& cfg, void* stream);
