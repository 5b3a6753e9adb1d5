// reset.cu -- car creation for reset() and the device-side auto-reset selector.
//
// Replaces (reference: gym_multi_car_racing/multi_car_racing.py):
//   mcr:341-350  zero reward / prev_reward / tile_visited_count / t / flags
//   mcr:400-406  car_dynamics.Car(world, angle, x, y): hull at (x, y, angle); the four wheels
//                at (x + wx*SIZE, y + wy*SIZE) -- offsets NOT rotated by the spawn angle -- with
//                the same angle; zero velocities, fresh joints (impulses 0, limit inactive)
//   mcr:351-357  episode direction and car order draws (auto-reset path: a counter-based hash
//                RNG replaces the global numpy RNG; the host-driven mcr_reset path takes the
//                reference's own draws from the caller)
// The implicit step(None) of mcr:408 is issued by the host API right after this kernel.
#include "mcr_internal.h"
#include <cuda_runtime.h>

#define SP_THREADS 128

__device__ __forceinline__ void rot_set_d(float a, float& s, float& c) {
    double ds, dc;
    sincos((double)a, &ds, &dc);
    s = (float)ds; c = (float)dc;
}

__device__ void spawn_car(const Dims& d, const DevBuffers& b, const CarConst& cc, int car, double ang, double ix, double iy) {
    const int N = d.N;
    const double SIZE = 0.02;
    const double WHEELPOS[4][2] = {{-55, +80}, {+55, +80}, {-55, -82}, {+55, -82}};
    const float fang = (float)ang;
    float qs, qc; rot_set_d(fang, qs, qc);
    for (int i = 0; i < 5; ++i) {
        float px, py, cx, cy;
        if (i == 0) {
            px = (float)ix; py = (float)iy;
            cx = (qc * cc.hull_lcx - qs * cc.hull_lcy) + px;      // sweep.c = b2Mul(xf, localCenter)
            cy = (qs * cc.hull_lcx + qc * cc.hull_lcy) + py;
        } else {
            px = (float)(ix + WHEELPOS[i - 1][0] * SIZE); py = (float)(iy + WHEELPOS[i - 1][1] * SIZE);
            cx = px; cy = py;                                      // wheel centre of mass is its origin
        }
        float* p = b.body + (size_t)(i * BODY_FIELDS) * N + car;
        p[(size_t)BF_CX * N] = cx; p[(size_t)BF_CY * N] = cy; p[(size_t)BF_A * N] = fang;
        p[(size_t)BF_VX * N] = 0.0f; p[(size_t)BF_VY * N] = 0.0f; p[(size_t)BF_W * N] = 0.0f;
        p[(size_t)BF_PX * N] = px; p[(size_t)BF_PY * N] = py; p[(size_t)BF_QS * N] = qs; p[(size_t)BF_QC * N] = qc;
        b.sleep_time[(size_t)i * N + car] = 0.0f;
        b.awake[(size_t)i * N + car] = 1;
    }
    for (int k = 0; k < 4; ++k) {
        float* p = b.joint + (size_t)(k * JOINT_FIELDS) * N + car;
        p[(size_t)JF_IX * N] = 0.0f; p[(size_t)JF_IY * N] = 0.0f; p[(size_t)JF_IZ * N] = 0.0f; p[(size_t)JF_MOTOR * N] = 0.0f;
        b.limit_state[(size_t)k * N + car] = LIM_INACTIVE;
        b.wheel[(size_t)(k * WHEEL_FIELDS + WF_OMEGA) * N + car] = 0.0;
        b.wheel[(size_t)(k * WHEEL_FIELDS + WF_PHASE) * N + car] = 0.0;
        b.on_road[(size_t)k * N + car] = 0;
        b.on_road_next[(size_t)k * N + car] = 0;
    }
    for (int f = 0; f < CTRL_FIELDS; ++f) b.ctrl[(size_t)f * N + car] = 0.0;
    b.reward[car] = 0.0; b.prev_reward[car] = 0.0; b.visit_count[car] = 0; b.backward[car] = 0;
    b.time[car] = 0.0; b.steps[car] = 0;
}

__global__ void __launch_bounds__(SP_THREADS)
spawn_kernel(Dims d, DevBuffers b, CarConst cc, const uint8_t* __restrict__ mask, const int32_t* __restrict__ track_slot,
             const uint8_t* __restrict__ cw, const double* __restrict__ spawn_pose) {
    const int env = blockIdx.x;
    if (mask && !mask[env]) return;
    const int tid = threadIdx.x;
    if (tid == 0) { b.env_track[env] = track_slot[env]; b.env_cw[env] = cw[env]; b.n_manifold[env] = 0; b.pending[env] = 0; }
    for (int i = tid; i < d.Tmax; i += SP_THREADS) {
        b.visited[(size_t)env * d.Tmax + i] = 0u;
        b.touched[(size_t)env * d.Tmax + i] = 0;
    }
    if (tid < d.A) {
        const double* sp = spawn_pose + ((size_t)env * d.A + tid) * 3;
        spawn_car(d, b, cc, env * d.A + tid, sp[0], sp[1], sp[2]);
    }
}

int launch_spawn(const Dims& d, const DevBuffers& b, const CarConst& cc, const uint8_t* mask,
                 const int32_t* track_slot, const uint8_t* cw, const double* spawn_pose, void* stream) {
    spawn_kernel<<<d.B, SP_THREADS, 0, (cudaStream_t)stream>>>(d, b, cc, mask, track_slot, cw, spawn_pose);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// ---------------------------------------------------------------------------------------
// Device-side auto reset: envs whose done flag is set pick the next track of the pool, a
// direction and a car order, and respawn -- no host round trip.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__global__ void __launch_bounds__(SP_THREADS)
auto_reset_kernel(Dims d, DevBuffers b, CarConst cc, const uint8_t* __restrict__ done, AutoResetCfg cfg) {
    const int env = blockIdx.x;
    const int tid = threadIdx.x;
    const bool r = done[env] != 0;
    if (tid == 0) b.reset_mask[env] = r ? 1 : 0;
    if (!r) return;
    __shared__ int s_slot, s_cw;
    __shared__ int s_order[MCR_MAX_AGENTS];
    if (tid == 0) {
        const uint32_t episode = b.env_episode[env] + 1u;
        b.env_episode[env] = episode;
        uint64_t st = splitmix64(cfg.seed ^ ((uint64_t)(uint32_t)env << 32) ^ (uint64_t)episode);
        const int slot = (int)(splitmix64(st) % (uint64_t)d.P); st = splitmix64(st + 1);
        int cwv = cfg.direction_cw;
        if (cfg.use_random_direction) { cwv = (int)(splitmix64(st) & 1ull); st = splitmix64(st + 2); }
        for (int i = 0; i < d.A; ++i) s_order[i] = i;
        for (int i = d.A - 1; i > 0; --i) {          // Fisher-Yates
            st = splitmix64(st + 3);
            const int j = (int)(st % (uint64_t)(i + 1));
            const int tmp = s_order[i]; s_order[i] = s_order[j]; s_order[j] = tmp;
        }
        s_slot = slot; s_cw = cwv;
        b.env_track[env] = slot; b.env_cw[env] = (uint8_t)cwv; b.n_manifold[env] = 0;
    }
    __syncthreads();
    for (int i = tid; i < d.Tmax; i += SP_THREADS) {
        b.visited[(size_t)env * d.Tmax + i] = 0u;
        b.touched[(size_t)env * d.Tmax + i] = 0;
    }
    if (tid < d.A) {
        // spawn pose of grid position `order` on this track under this direction (host-computed
        // at track load with the reference's arithmetic, mcr:366-393)
        const double* sp = b.trk_slot_pose + (((size_t)s_slot * 2 + s_cw) * d.A + s_order[tid]) * 3;
        spawn_car(d, b, cc, env * d.A + tid, sp[0], sp[1], sp[2]);
    }
}

int launch_auto_reset(const Dims& d, const DevBuffers& b, const CarConst& cc, const uint8_t* done,
                      const AutoResetCfg& cfg, void* stream) {
    auto_reset_kernel<<<d.B, SP_THREADS, 0, (cudaStream_t)stream>>>(d, b, cc, done, cfg);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}
