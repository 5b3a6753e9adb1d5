// trackgen.cu -- on-device batched track generation: reset()'s retry loop around
// MultiCarRacing._create_track (reference gym_multi_car_racing/multi_car_racing.py:183-338, 359-364),
// one THREAD per track, written straight into a slot of the device track pool.
//
// Each track consumes its own numpy-legacy MT19937 stream (625 words in/out: the host RandomState
// of that env stays in lock step), draws the same uniforms in the same order as the reference and
// runs the same float64 operation sequence as the host generator (mcr_track_generate in api.cu).
// The only difference from the host path is libm: CUDA's fp64 sin / cos / atan2 are within 1-2 ulp
// of glibc's, so node coordinates agree to ~1e-12 (not bit for bit), and after the fp32 conversion
// Box2D and GL apply (b2Vec2 / glVertex3f) nearly every tile vertex is identical.  The host
// generator stays the bit-exact reference path; this kernel is the throughput path (a full reset
// of 8192 envs is 2.5 s of one host core, or one launch here).
//
// Per-track scratch (global memory, caller-owned): the 2500-node pursuit path (alpha, beta, x, y)
// and the per-tile border flags.
#include "mcr_internal.h"
#include <cuda_runtime.h>
#include "trackgen.cuh"

#define TG_THREADS 32

extern "C" int64_t mcr_trackgen_scratch_bytes(void) { return (int64_t)TG_SCRATCH_BYTES; }

__global__ void __launch_bounds__(TG_THREADS)
trackgen_kernel(Dims d, DevBuffers b, int n_tracks, uint32_t* __restrict__ mt_state, const int32_t* __restrict__ slots,
                unsigned char* __restrict__ scratch, size_t scratch_stride, int32_t* __restrict__ result, int max_attempts) {
    const int i = blockIdx.x * TG_THREADS + threadIdx.x;
    if (i >= n_tracks) return;
    const int slot = slots ? slots[i] : i;
    int32_t* res = result + (size_t)i * 4;      // T (or error), attempts, i1, i2
    if (slot < 0 || slot >= d.P) { res[0] = -1; res[1] = 0; return; }
    tg_generate_track(d, b, mt_state + (size_t)i * 625, slot, scratch + (size_t)i * scratch_stride, max_attempts, res);
}

int launch_trackgen(const Dims& d, const DevBuffers& b, int n, uint32_t* mt_state, const int32_t* slots, void* scratch,
                    int32_t* result, int max_attempts, void* stream) {
    const size_t stride = (size_t)mcr_trackgen_scratch_bytes();
    trackgen_kernel<<<(n + TG_THREADS - 1) / TG_THREADS, TG_THREADS, 0, (cudaStream_t)stream>>>(
        d, b, n, mt_state, slots, (unsigned char*)scratch, stride, result, max_attempts);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// Top up every env's ring to R tracks ahead of the episode in progress.  Runs on the handle's low-priority refill
// stream beside the steps; a thread that finds its env's generator lock taken leaves (a starved consumer is generating).
__global__ void __launch_bounds__(TG_THREADS)
ring_refill_kernel(Dims d, DevBuffers b, int R) {
    const int env = blockIdx.x * TG_THREADS + threadIdx.x;
    if (env >= d.B) return;
    for (;;) {
        if (ring_ld(b.trk_produced + env) >= ring_ld(b.trk_consumed + env) + R) return;
        if (atomicCAS(b.trk_lock + env, 0, 1) != 0) return;
        if (ring_ld(b.trk_produced + env) < ring_ld(b.trk_consumed + env) + R) ring_produce_one(d, b, env, R);
        atomicExch(b.trk_lock + env, 0);
    }
}

int launch_ring_refill(const Dims& d, const DevBuffers& b, int R, void* stream) {
    ring_refill_kernel<<<(d.B + TG_THREADS - 1) / TG_THREADS, TG_THREADS, 0, (cudaStream_t)stream>>>(d, b, R);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}
