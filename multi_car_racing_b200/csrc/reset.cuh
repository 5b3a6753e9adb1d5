// reset.cuh -- car creation (car_dynamics.Car at a spawn pose, mcr:400-406, and the zeroing of
// mcr:341-350) and the device-side auto-reset draw, shared by reset.cu's kernels and the fused
// head_kernel (carcontacts.cu).
#pragma once
#include "mcr_internal.h"
#include <cuda_runtime.h>
#include "trackgen.cuh"

__device__ __forceinline__ void rot_set_d(float a, float& s, float& c) {
    double ds, dc;
    sincos((double)a, &ds, &dc);
    s = (float)ds; c = (float)dc;
}

__device__ __forceinline__ void spawn_car(const Dims& d, const DevBuffers& b, const CarConst& cc, int car, double ang, double ix, double iy) {
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
    b.reward[car] = 0.0; b.prev_reward[car] = 0.0; b.visit_count[car] = 0; b.backward[car] = 0; b.on_grass[car] = 0;
    if (d.particles) {                       // Car.__init__: particles = [], skid_start = skid_particle = None
        b.prt_hdr[car] = 0; b.prt_hdr[(size_t)N + car] = 0;
        for (int k = 0; k < 4; ++k) b.skid_meta[(size_t)k * N + car] = 0;
    }
    b.time[car] = 0.0; b.steps[car] = 0;
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// What the auto reset draws for episode `episode` of env `env`: pool slot, direction and the car
// order (grid position of every car) -- a pure function of (seed, env, episode), so every thread
// of a block / warp can evaluate it redundantly instead of broadcasting it.
struct ResetDraw { int slot, cw; int order[MCR_MAX_AGENTS]; };
__device__ __forceinline__ void auto_reset_draw(const Dims& d, const AutoResetCfg& cfg, int env, uint32_t episode, ResetDraw& r) {
    uint64_t st = splitmix64(cfg.seed ^ ((uint64_t)(uint32_t)env << 32) ^ (uint64_t)episode);
    r.slot = (int)(splitmix64(st) % (uint64_t)d.P); st = splitmix64(st + 1);
    r.cw = cfg.direction_cw;
    if (cfg.use_random_direction) { r.cw = (int)(splitmix64(st) & 1ull); st = splitmix64(st + 2); }
    for (int i = 0; i < d.A; ++i) r.order[i] = i;
    for (int i = d.A - 1; i > 0; --i) {          // Fisher-Yates
        st = splitmix64(st + 3);
        const int j = (int)(st % (uint64_t)(i + 1));
        const int tmp = r.order[i]; r.order[i] = r.order[j]; r.order[j] = tmp;
    }
}

// Respawn env `env` with `nthreads` cooperating threads (tid = 0 .. nthreads - 1, nthreads >= A); the
// caller orders the episode-counter read before this call and synchronises its threads afterwards.
__device__ __forceinline__ void auto_reset_env(const Dims& d, const DevBuffers& b, const CarConst& cc, const AutoResetCfg& cfg,
                                               int env, uint32_t episode, int tid, int nthreads) {
    ResetDraw r; auto_reset_draw(d, cfg, env, episode, r);
    if (cfg.fresh) {
        // a fresh track on the env's own RandomState stream (mcr:359-364): the next slot of the env's ring, generated
        // ahead of time by the refill kernel (or right here if the episode was shorter than a track generation)
        if (tid == 0) b.env_track[env] = ring_next_slot(d, b, env, cfg.fresh);
        if (nthreads <= 32) __syncwarp(); else __syncthreads();
        r.slot = ring_ld(b.env_track + env);
    }
    if (tid == 0) { b.env_episode[env] = episode; b.env_track[env] = r.slot; b.env_cw[env] = (uint8_t)r.cw; b.n_manifold[env] = 0; }
    for (int i = tid; i < d.Tmax; i += nthreads) {
        b.visited[(size_t)env * d.Tmax + i] = 0u;
        b.touched[(size_t)env * d.Tmax + i] = 0;
    }
    if (tid < d.A) {
        // spawn pose of grid position `order` on this track under this direction (evaluated at track
        // load / generation with the reference's arithmetic, mcr:366-393)
        const double* sp = b.trk_slot_pose + (((size_t)r.slot * 2 + r.cw) * d.A + r.order[tid]) * 3;
        spawn_car(d, b, cc, env * d.A + tid, sp[0], sp[1], sp[2]);
    }
}
