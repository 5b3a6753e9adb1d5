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
#include "reset.cuh"

#define SP_THREADS 128

__global__ void __launch_bounds__(SP_THREADS)
spawn_kernel(Dims d, DevBuffers b, CarConst cc, const uint8_t* __restrict__ mask, const int32_t* __restrict__ track_slot,
             const uint8_t* __restrict__ cw, const double* __restrict__ spawn_pose) {
    const int env = blockIdx.x;
    if (mask && !mask[env]) return;
    const int tid = threadIdx.x;
    if (tid == 0) {
        b.env_track[env] = track_slot[env]; b.env_cw[env] = cw[env]; b.n_manifold[env] = 0; b.pending[env] = 0;
        b.trk_consumed[env] = 0; b.trk_produced[env] = 0; b.trk_lock[env] = 0;     // track ring: this episode is ring index 0
    }
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
__global__ void __launch_bounds__(SP_THREADS)
auto_reset_kernel(Dims d, DevBuffers b, CarConst cc, const uint8_t* __restrict__ done, AutoResetCfg cfg) {
    const int env = blockIdx.x;
    const int tid = threadIdx.x;
    const bool r = done[env] != 0;
    if (tid == 0) b.reset_mask[env] = r ? 1 : 0;
    if (!r) return;
    const uint32_t episode = b.env_episode[env] + 1u;
    __syncthreads();                                  // every thread has read the counter before thread 0 bumps it
    auto_reset_env(d, b, cc, cfg, env, episode, tid, SP_THREADS);
}

int launch_auto_reset(const Dims& d, const DevBuffers& b, const CarConst& cc, const uint8_t* done,
                      const AutoResetCfg& cfg, void* stream) {
    auto_reset_kernel<<<d.B, SP_THREADS, 0, (cudaStream_t)stream>>>(d, b, cc, done, cfg);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}
