// sim.cu -- kernels 1+2: FrictionDetector/Collide (contacts_kernel) and Car.step + world.Step
// (pre_kernel -> sweep_kernel -> post_kernel).
//
// Replaces, for every env of the batch (reference: gym_multi_car_racing/multi_car_racing.py):
//   mcr:84-123   FrictionDetector.BeginContact/EndContact/_contact; Box2D b2Contact::Update for
//                sensor fixtures: touching = b2TestOverlap = (GJK distance < rA + rB = 0.02),
//                evaluated at the START of world.Step on the poses the previous step left
//   mcr:421-424  car.steer(-a0) / car.gas(a1) / car.brake(a2)      (gym car_dynamics.Car)
//   mcr:426-427  car.step(1/FPS)            -- fp64 tyre model on fp32 body state
//   mcr:428      world.Step(1/FPS, 180, 60) -- b2Island::Solve specialised to the fixed
//                topology {hull + 4 wheels + 4 revolute joints (limit + motor)}; island joint
//                order [j3, j2, j1, j0]; sleeping; SynchronizeTransform.
// Numerics: IEEE fp32, no FMA contraction (-fmad=false), same operation order as Box2D;
// b2Rot::Set is (float)sin((double)a).  The wheel's local centre of mass and localAnchorB are
// exactly 0 so every rB term of b2RevoluteJoint is an exact +-0 and is dropped.
//
// Mapping.
//   contacts_kernel  one warp per env: world-space fixture polygons in shared memory, tiles strided
//                    over lanes (coalesced float4 AABB loads), AABB reject against per-car then
//                    per-fixture boxes, surviving (tile, fixture) pairs queued and tested one pair
//                    per lane with the exact predicate; new visits replayed in the contact-list order
//                    of a fresh b2World (tile descending, car descending) so the float64 reward sums
//                    are bit-reproducible -- no racing atomicAdds.  Runs on a side stream beside the
//                    solver; it only reads the step's start poses.
//   pre_kernel       one THREAD per car, SoA-coalesced: controls, the float64 tyre model, velocity
//                    integration, InitVelocityConstraints + warm start (the split entry points; mcr_step's
//                    head_kernel, carcontacts.cu, runs the same arithmetic with four lanes per car).
//   sweep_kernel     the 180 Gauss-Seidel sweeps are a serial dependency chain through the hull velocity
//                    (latency bound): one THREAD per car for the cars without an active joint limit (32 per
//                    warp), one WARP per car -- straight-line code specialised to the limit pattern -- for
//                    the rest; a warp leaves the loop when its cars sit on an exact fixed point / cycle.
//   post_kernel      two threads per car in two warps (solver warp: position integration, <= 60 position
//                    iterations, SynchronizeTransform, sleeping; view warp: the per-view values the
//                    rasteriser needs -- camera affine, heading, snapshots).
//   The kernels of mcr_step's chain hand over by per-car ready flags (DevBuffers::ready, mcr_internal.h).
#include "mcr_internal.h"
#include <cuda_runtime.h>
#include <cstdlib>

#define AABB_MARGIN 0.05f
#define FIX_STRIDE 21   // 8 x + 8 y + 4 aabb + (n | active<<8), odd stride spreads banks
#define PAIRS_PER_CAR 128
#define CANDS_PER_CAR 64


#include "solver.cuh"
#include "pre.cuh"

// ---------------------------------------------------------------------------------------
// contacts
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float seg_dist2(float px, float py, float ax, float ay, float bx, float by) {
    float ex = bx - ax, ey = by - ay, wx = px - ax, wy = py - ay;
    float t = wx * ex + wy * ey;
    if (t <= 0.0f) return wx * wx + wy * wy;
    float l2 = ex * ex + ey * ey;
    if (t >= l2) { float ux = px - bx, uy = py - by; return ux * ux + uy * uy; }
    float cr = ex * wy - ey * wx;
    return (cr * cr) / l2;
}

// a = tile quad (4 verts), b = fixture polygon (nb verts); both CCW, world space.
// Same predicate as b2TestOverlap: distance(a, b) < 2 * b2_polygonRadius (+ 10 eps).
__device__ bool poly_touch(const float* ax, const float* ay, const float* bx, const float* by, int nb) {
    bool separated = false;
    for (int i = 0; i < 4 && !separated; ++i) {
        int i2 = i + 1 < 4 ? i + 1 : 0;
        float px = ax[i], py = ay[i], ex = ax[i2] - px, ey = ay[i2] - py;
        bool all_out = true;
        for (int k = 0; k < nb; ++k) {
            float c = ex * (by[k] - py) - ey * (bx[k] - px);
            if (!(c < 0.0f)) { all_out = false; break; }
        }
        separated = all_out;
    }
    for (int i = 0; i < nb && !separated; ++i) {
        int i2 = i + 1 < nb ? i + 1 : 0;
        float px = bx[i], py = by[i], ex = bx[i2] - px, ey = by[i2] - py;
        bool all_out = true;
        for (int k = 0; k < 4; ++k) {
            float c = ex * (ay[k] - py) - ey * (ax[k] - px);
            if (!(c < 0.0f)) { all_out = false; break; }
        }
        separated = all_out;
    }
    if (!separated) return true;
    float d2 = 3.402823466e+38f;
    for (int i = 0; i < 4; ++i)
        for (int k = 0; k < nb; ++k) {
            int k2 = k + 1 < nb ? k + 1 : 0;
            d2 = fminf(d2, seg_dist2(ax[i], ay[i], bx[k], by[k], bx[k2], by[k2]));
        }
    for (int k = 0; k < nb; ++k)
        for (int i = 0; i < 4; ++i) {
            int i2 = i + 1 < 4 ? i + 1 : 0;
            d2 = fminf(d2, seg_dist2(bx[k], by[k], ax[i], ay[i], ax[i2], ay[i2]));
        }
    float dist = sqrtf(d2);
    const float rr = B2_POLYGON_RADIUS + B2_POLYGON_RADIUS;
    if (dist > rr && dist > B2_EPS) return (dist - rr) < 10.0f * B2_EPS;
    return true;
}

struct ContactSmem {      // carved from dynamic shared memory, sizes depend on A
    float* fixt;          // [A*8][FIX_STRIDE]
    float* carbox;        // [A][4]
    uint32_t* pairs;      // [PAIRS_PER_CAR*A]  tile << 8 | fixture
    uint32_t* cands;      // [CANDS_PER_CAR*A]  tile << 8 | car  (new-visit candidates)
    uint32_t* road_bits;  // [2]
    int* counts;          // [2] pairs, cands
};

__device__ __forceinline__ ContactSmem carve(unsigned char* base, int A) {
    ContactSmem S;
    S.fixt = (float*)base; base += (size_t)A * 8 * FIX_STRIDE * 4;
    S.carbox = (float*)base; base += (size_t)A * 4 * 4;
    S.pairs = (uint32_t*)base; base += (size_t)A * PAIRS_PER_CAR * 4;
    S.cands = (uint32_t*)base; base += (size_t)A * CANDS_PER_CAR * 4;
    S.road_bits = (uint32_t*)base; base += 8;
    S.counts = (int*)base;
    return S;
}

static size_t sim_smem_bytes(int A) {
    return (size_t)A * 8 * FIX_STRIDE * 4 + (size_t)A * 16 + (size_t)A * PAIRS_PER_CAR * 4 + (size_t)A * CANDS_PER_CAR * 4 + 16;
}

__device__ void contacts_warp(const Dims& d, const DevBuffers& b, const CarConst& cc, int env, int lane, unsigned char* smem) {
    const int A = d.A, N = d.N, nfix = A * 8;
    ContactSmem S = carve(smem, A);
    if (lane < 2) { S.counts[lane] = 0; S.road_bits[lane] = 0u; }
    // ---- phase 1: world-space fixture polygons + AABBs -----------------------------------------
    for (int f = lane; f < nfix; f += 32) {
        const int c = f >> 3, fi = f & 7, car = env * A + c;
        const int body = fi < 4 ? 1 + fi : 0;
        const Poly8& P = fi < 4 ? cc.wheel_poly : cc.hull_poly[fi - 4];
        const float* bp = b.body + (size_t)(body * BODY_FIELDS) * N + car;
        const float px = bp[(size_t)BF_PX * N], py = bp[(size_t)BF_PY * N];
        const float qs = bp[(size_t)BF_QS * N], qc = bp[(size_t)BF_QC * N];
        float* o = S.fixt + (size_t)f * FIX_STRIDE;
        float lx = 3.402823466e+38f, ly = lx, hx = -lx, hy = -lx;
        for (int i = 0; i < P.n; ++i) {
            float x = (qc * P.x[i] - qs * P.y[i]) + px;
            float y = (qs * P.x[i] + qc * P.y[i]) + py;
            o[i] = x; o[8 + i] = y;
            lx = fminf(lx, x); ly = fminf(ly, y); hx = fmaxf(hx, x); hy = fmaxf(hy, y);
        }
        o[16] = lx; o[17] = ly; o[18] = hx; o[19] = hy;
        // wheels are always awake at Collide time (Car.step woke them); the hull may sleep
        const int active = fi < 4 ? 1 : (b.awake[car] != 0);   // awake[0 * N + car]
        o[20] = __int_as_float(P.n | (active << 8));
    }
    __syncwarp();
    for (int c = lane; c < A; c += 32) {
        float lx = 3.402823466e+38f, ly = lx, hx = -lx, hy = -lx;
        for (int fi = 0; fi < 8; ++fi) {
            const float* o = S.fixt + (size_t)(c * 8 + fi) * FIX_STRIDE;
            lx = fminf(lx, o[16]); ly = fminf(ly, o[17]); hx = fmaxf(hx, o[18]); hy = fmaxf(hy, o[19]);
        }
        S.carbox[c * 4 + 0] = lx; S.carbox[c * 4 + 1] = ly; S.carbox[c * 4 + 2] = hx; S.carbox[c * 4 + 3] = hy;
    }
    __syncwarp();
    // ---- phase 2: broad phase, tiles strided over lanes ------------------------------------------
    const int slot = b.env_track[env];
    const int T = b.trk_T[slot];
    const float4* aabbs = (const float4*)(b.trk_tile_aabb + (size_t)slot * d.Tmax * 4);
    const float* tiles = b.trk_tile + (size_t)slot * d.Tmax * 8;
    const int pair_cap = PAIRS_PER_CAR * A, cand_cap = CANDS_PER_CAR * A;
    for (int t = lane; t < T; t += 32) {
        const float4 ta = aabbs[t];
        for (int c = 0; c < A; ++c) {
            const float* cb = S.carbox + c * 4;
            if (cb[0] - ta.z > AABB_MARGIN || cb[1] - ta.w > AABB_MARGIN || ta.x - cb[2] > AABB_MARGIN || ta.y - cb[3] > AABB_MARGIN) continue;
            for (int fi = 0; fi < 8; ++fi) {
                const float* o = S.fixt + (size_t)(c * 8 + fi) * FIX_STRIDE;
                if (o[16] - ta.z > AABB_MARGIN || o[17] - ta.w > AABB_MARGIN || ta.x - o[18] > AABB_MARGIN || ta.y - o[19] > AABB_MARGIN) continue;
                const int slot_p = atomicAdd(&S.counts[0], 1);
                if (slot_p < pair_cap) S.pairs[slot_p] = ((uint32_t)t << 8) | (uint32_t)(c * 8 + fi);
                else atomicExch(&b.status[ST_EVENT_OVERFLOW], 1);
            }
        }
    }
    __syncwarp();
    // ---- phase 3: exact predicate, one queued pair per lane ---------------------------------------
    uint8_t* touched = b.touched + (size_t)env * d.Tmax;
    int npairs = S.counts[0]; if (npairs > pair_cap) npairs = pair_cap;
    for (int i = lane; i < npairs; i += 32) {
        const uint32_t pr = S.pairs[i];
        const int t = (int)(pr >> 8), f = (int)(pr & 0xffu), c = f >> 3, fi = f & 7;
        const float4 v0 = *(const float4*)(tiles + (size_t)t * 8);
        const float4 v1 = *(const float4*)(tiles + (size_t)t * 8 + 4);
        const float tx[4] = {v0.x, v0.z, v1.x, v1.z}, ty[4] = {v0.y, v0.w, v1.y, v1.w};
        const float* o = S.fixt + (size_t)f * FIX_STRIDE;
        const int meta = __float_as_int(o[20]);
        // a hull fixture's contact only ever recolours the tile (mcr:102-108): nothing to learn once it is grey
        if (fi >= 4 && touched[t]) continue;
        if (!poly_touch(tx, ty, o, o + 8, meta & 0xff)) continue;
        if (fi < 4) atomicOr(&S.road_bits[(c * 4 + fi) >> 5], 1u << ((c * 4 + fi) & 31));   // len(wheel.tiles) > 0
        if (!(meta >> 8)) continue;              // sleeping body: contact not updated
        touched[t] = 1;                          // tile.color = ROAD_COLOR, mcr:102-104 (any body, idempotent)
        if (fi >= 4) continue;                   // hull.userData is None, mcr:108
        // only a FIRST visit changes anything in the replay below (mcr:113); visited[] is not written before
        // phase 4, so the test is safe here and keeps the serial replay list to the step's new visits
        if ((b.visited[(size_t)env * d.Tmax + t] >> c) & 1u) continue;
        const int slot_c = atomicAdd(&S.counts[1], 1);
        if (slot_c < cand_cap) S.cands[slot_c] = ((uint32_t)t << 8) | (uint32_t)c;
        else atomicExch(&b.status[ST_EVENT_OVERFLOW], 1);
    }
    __syncwarp();
    // ---- phase 4: flags for the next Car.step; replay visits in contact-list order ------------------
    for (int i = lane; i < A * 4; i += 32) {
        const int c = i >> 2, k = i & 3;
        b.on_road_next[(size_t)k * N + env * A + c] = (uint8_t)((S.road_bits[i >> 5] >> (i & 31)) & 1u);
    }
    if (lane == 0) {
        int n = S.counts[1]; if (n > cand_cap) n = cand_cap;
        uint32_t* ev = S.cands;
        for (int i = 1; i < n; ++i) {            // insertion sort, descending (tile, car); n is tiny
            const uint32_t k = ev[i]; int j = i - 1;
            while (j >= 0 && ev[j] < k) { ev[j + 1] = ev[j]; --j; }
            ev[j + 1] = k;
        }
        uint32_t* visited = b.visited + (size_t)env * d.Tmax;
        uint32_t prev = 0xffffffffu;
        for (int i = 0; i < n; ++i) {
            const uint32_t k = ev[i];
            if (k == prev) continue;             // another wheel of the same car on the same tile
            prev = k;
            const int t = (int)(k >> 8), c = (int)(k & 0xffu);
            uint32_t vis = visited[t];
            if ((vis >> c) & 1u) continue;       // mcr:113
            vis |= 1u << c;
            visited[t] = vis;
            const int car = env * A + c;
            b.visit_count[car] += 1;             // mcr:115
            const int past = __popc(vis) - 1;    // mcr:118-120
            const double reward_factor = 1 - ((double)past / (double)A);
            b.reward[car] += reward_factor * 1000.0 / (double)T;
        }
    }
}

// ---------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------
#define CT_WARPS 4

__global__ void __launch_bounds__(CT_WARPS * 32)
contacts_kernel(Dims d, DevBuffers b, CarConst cc, const uint8_t* __restrict__ mask, size_t smem_per_warp, int set_flags) {
    extern __shared__ __align__(16) unsigned char sim_smem[];
    tl_stamp(b.timeline, TL_CONTACTS);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int env = blockIdx.x * CT_WARPS + warp;
    if (env >= d.B) return;
    if (mask && !mask[env]) return;
    contacts_warp(d, b, cc, env, lane, sim_smem + (size_t)warp * smem_per_warp);
    if (set_flags) {                               // publish the env's cars to post_kernel (ready[3N + car])
        __syncwarp();
        if (lane < d.A) flag_release(b.ready + 3 * d.N + env * d.A + lane, 1);
    }
}

// pre/post are per-thread latency chains (fp64 division / sqrt / sincos); block size 32..128 measured equal

#define PRE_BLOCK 128
static int pre_block() { const char* e = getenv("MCR_PRE_BLOCK"); const int v = e ? atoi(e) : PRE_BLOCK; return (v == 32 || v == 64 || v == 128) ? v : PRE_BLOCK; }   // tuning knob

template <typename ActT>
__global__ void __launch_bounds__(128)
pre_kernel(Dims d, DevBuffers b, CarConst cc, const uint8_t* __restrict__ mask, const uint8_t* __restrict__ noact,
           const ActT* __restrict__ action) {
    const int car = blockIdx.x * blockDim.x + threadIdx.x;
    // a new pass: post_kernel publishes a car by writing the pass number into ready[car] (READY_EPOCH, see DevBuffers); counted
    // by the first kernel of the main chain, which everything that reads the number follows in stream order
    if (car == 0) b.ready[READY_EPOCH(d.N)] += 1;
    if (car >= d.N) return;
    const int env = car / d.A;
    if (mask && !mask[env]) return;
    pre_car<ActT>(d, b, cc, car, env, action && !(noact && noact[env]), action);
}

// pre_kernel's `action` argument for the captured step graph (api.cu patches it per step instead of copying the action)
int pre_action_arg(const void* func, int* nargs) {
    if (nargs) *nargs = 6;
    return (func == (const void*)pre_kernel<float> || func == (const void*)pre_kernel<double>) ? 5 : -1;
}

#define SWEEP_BLOCK 128
static int sweep_block() { const char* e = getenv("MCR_SWEEP_BLOCK"); const int v = e ? atoi(e) : SWEEP_BLOCK; return (v == 32 || v == 64 || v == 128) ? v : SWEEP_BLOCK; }   // tuning knob

__device__ __forceinline__ void sweep_load(const DevBuffers& b, int car, int N, VelState& s, JointC (&J)[4]) {
    const float* sc = b.scratch + car;
#pragma unroll
    for (int i = 0; i < 5; ++i) { s.vx[i] = sc[(size_t)(SC_VX + i) * N]; s.vy[i] = sc[(size_t)(SC_VY + i) * N]; s.w[i] = sc[(size_t)(SC_W + i) * N]; }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        s.jix[k] = sc[(size_t)(SC_JIX + k) * N]; s.jiy[k] = sc[(size_t)(SC_JIY + k) * N];
        s.jiz[k] = sc[(size_t)(SC_JIZ + k) * N]; s.jmot[k] = sc[(size_t)(SC_JMOT + k) * N];
        const float* q = sc + (size_t)(SC_JOINT + k * SC_JOINT_FIELDS) * N;
        JointC& j = J[k];
        j.rAx = q[(size_t)0 * N]; j.rAy = q[(size_t)1 * N]; j.k11 = q[(size_t)2 * N]; j.k12 = q[(size_t)3 * N];
        j.k22 = q[(size_t)4 * N]; j.ezx = q[(size_t)5 * N]; j.ezy = q[(size_t)6 * N]; j.ezz = q[(size_t)7 * N];
        j.det22 = q[(size_t)8 * N]; j.cfx = q[(size_t)9 * N]; j.cfy = q[(size_t)10 * N]; j.cfz = q[(size_t)11 * N];
        j.det33 = q[(size_t)12 * N]; j.motorMass = q[(size_t)13 * N]; j.motorSpeed = q[(size_t)14 * N];
        j.limit = b.limit_state[(size_t)k * N + car];
    }
}

__device__ __forceinline__ void sweep_store(const DevBuffers& b, int car, int N, const VelState& s) {
    float* so = b.scratch + car;
#pragma unroll
    for (int i = 0; i < 5; ++i) { so[(size_t)(SC_VX + i) * N] = s.vx[i]; so[(size_t)(SC_VY + i) * N] = s.vy[i]; so[(size_t)(SC_W + i) * N] = s.w[i]; }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        so[(size_t)(SC_JIX + k) * N] = s.jix[k]; so[(size_t)(SC_JIY + k) * N] = s.jiy[k];
        so[(size_t)(SC_JIZ + k) * N] = s.jiz[k]; so[(size_t)(SC_JMOT + k) * N] = s.jmot[k];
    }
}

// The 180 Gauss-Seidel sweeps are a serial chain (349 cycles per sweep for one warp alone on a
// scheduler, scripts/micro/chain_latency.cu); what a car costs is latency, what a WARP costs is issue
// slots.  One warp per car made every scheduler issue ~170 instructions per sweep for each of its
// ~3.5 resident warps (570-800 cycles per sweep measured), so the grid has two parts:
//   CTAs [0, packed)      one THREAD per car for the cars whose four limits are inactive (99 %):
//                         32 cars per warp, straight-line code, N/32 warps spread one per scheduler.
//                         A warp leaves when ALL its cars sit on an exact fixed point / 2-cycle /
//                         4-cycle (more multiples of four sweeps leave such a car bit-identical).
//   CTAs [packed, ...)    one WARP per car for the few cars with an active limit: code specialised
//                         to the limit pattern, no divergence inside the sweep, own early exit.
__global__ void __launch_bounds__(SWEEP_BLOCK)
sweep_kernel(Dims d, DevBuffers b, CarConst cc, const uint8_t* __restrict__ mask, int early_exit, int packed_ctas, int set_flags) {
    cudaGridDependencySynchronize();               // programmatic dependent launch behind head_kernel
    cudaTriggerProgrammaticLaunchCompletion();     // post_kernel may be placed from here on: its lanes wait for their own car's flag (ready[2N + car])
    const int N = d.N;
    tl_stamp(b.timeline, TL_SWEEP);
    const float h = (float)(1.0 / 50);
    Masses m; m.mA = cc.hull_invMass; m.iA = cc.hull_invI; m.mB = cc.wheel_invMass; m.iB = cc.wheel_invI;
    m.maxMotorImpulse = h * cc.max_motor_torque;
    const bool packed = (int)blockIdx.x < packed_ctas;
    const int car = packed ? blockIdx.x * blockDim.x + threadIdx.x
                           : (blockIdx.x - packed_ctas) * (blockDim.x / 32) + (threadIdx.x >> 5);
    bool mine = car < N;
    int pat = 0;
    if (mine) {
        const int env = car / d.A;
        if (mask && !mask[env]) mine = false;
        else if (b.n_manifold[env] > 0) mine = false;      // solved by coupled_kernel
        else {
#pragma unroll
            for (int k = 0; k < 4; ++k) pat |= (b.limit_state[(size_t)k * N + car] != LIM_INACTIVE ? 1 : 0) << k;
        }
    }
    if (packed) {
        mine = mine && pat == 0;
        const unsigned peers = __ballot_sync(0xffffffffu, mine);
        if (!mine) return;
        VelState s; JointC J[4];
        sweep_load(b, car, N, s, J);
        solve_velocity<0>(s, J, m, early_exit != 0, peers);
        sweep_store(b, car, N, s);
        if (set_flags) flag_release(b.ready + 2 * N + car, 1);
        if ((threadIdx.x & 31) == 0) atomicMax(b.timeline + TL_SWEEP_END_PACKED, mcr_globaltimer());
    } else {
        if (!mine || pat == 0) return;
        // every lane runs the same scalar arithmetic (the loads broadcast); lane 0 stores
        VelState s; JointC J[4];
        sweep_load(b, car, N, s, J);
        switch (pat) {
            case 1: solve_velocity<1>(s, J, m, early_exit != 0); break;
            case 2: solve_velocity<2>(s, J, m, early_exit != 0); break;
            case 3: solve_velocity<3>(s, J, m, early_exit != 0); break;
            default: solve_velocity<-1>(s, J, m, early_exit != 0); break;   // a rear joint at its limit: rare
        }
        if ((threadIdx.x & 31) == 0) { sweep_store(b, car, N, s); if (set_flags) flag_release(b.ready + 2 * N + car, 1); atomicMax(b.timeline + TL_SWEEP_END, mcr_globaltimer()); }
    }
}

// Diagnostics build (-DMCR_PHASE_CLOCKS): thread 0 of every post CTA adds the cycles of each part to g_post_clk
// (0 grid-dependency wait, 1 loads, 2 integrate + position iterations, 3 transforms + sleep, 4 stores, 5 camera; 7 = CTAs).
#ifdef MCR_PHASE_CLOCKS
__device__ unsigned long long g_post_clk[8];
__device__ unsigned long long g_post_iter_hist[64];      // position iterations a car needed (all cars, not just thread 0)
__device__ unsigned long long g_post_warp_t[4][1024];    // per warp of the last cls != 2 launch: %globaltimer at entry, past the early flags, past the sweep flag, at the end
#define PK_T0() long long pk_t_ = clock64()
#define PK(k) do { if (threadIdx.x == 0) { const long long n_ = clock64(); atomicAdd(&g_post_clk[k], (unsigned long long)(n_ - pk_t_)); pk_t_ = n_; } } while (0)
extern "C" int mcr_debug_post_clocks(unsigned long long* out8, int reset) {
    if (out8 && cudaMemcpyFromSymbol(out8, g_post_clk, sizeof(g_post_clk)) != cudaSuccess) return -1;
    if (reset) { unsigned long long z[8] = {}; if (cudaMemcpyToSymbol(g_post_clk, z, sizeof(z)) != cudaSuccess) return -1; }
    if (reset == -2) return out8 && cudaMemcpyFromSymbol(out8, g_post_warp_t, sizeof(g_post_warp_t)) == cudaSuccess ? 0 : -1;   // out8: 4096 entries
    if (reset < 0) {                              // reset = -1: out8 is a 64-entry buffer for the iteration histogram instead
        if (out8 && cudaMemcpyFromSymbol(out8, g_post_iter_hist, sizeof(g_post_iter_hist)) != cudaSuccess) return -1;
        unsigned long long z[64] = {}; if (cudaMemcpyToSymbol(g_post_iter_hist, z, sizeof(z)) != cudaSuccess) return -1;
    }
    return 0;
}
#else
#define PK_T0() do {} while (0)
#define PK(k) do {} while (0)
#endif

// Two THREADS per car in two different warps of the CTA (warps [0, P) = solver, warps [P, 2P) = view, P = blockDim / 64;
// solver warp p and view warp P + p hold the same 32 cars, lane = car): the solver thread integrates the positions, runs
// the position iterations, synchronises the transforms, takes the sleep decision and stores the bodies and joints; the
// view thread meanwhile evaluates what only depends on the solved VELOCITIES -- the camera's angle (atan2 + two sincos,
// mcr:544-556) and the heading of the backward test (mcr:449-456) --, takes the score / backward snapshots, advances the
// env clock, and finishes the camera once the solver thread hands it the final hull pose (shared memory, a named barrier of
// the two warps).  Both are serial fp64-trig chains; on two warp schedulers they really run side by side (as two lanes
// of one warp -- the form before -- the divergent halves only interleaved: 10 us per warp, now the longer of the two).
__global__ void __launch_bounds__(128)
post_kernel(Dims d, DevBuffers b, CarConst cc, const uint8_t* __restrict__ mask, const uint8_t* __restrict__ noact, int has_action,
            double h_ratio, int cls, int wait_sweep) {
    PK_T0();
    // wait_sweep (the step's main chain, launched right behind sweep_kernel, which triggers at its start): each lane pair
    // waits below for its own car's flags only -- 1: the sweep's; 2: also the contact pass's and the wheel stripes' (the
    // launch then has no other dependency than the sweep and is placed while the sweep runs).
    // Otherwise: programmatic dependent launch behind the sweep / coupled kernel.
    if (!wait_sweep) cudaGridDependencySynchronize();
    PK(0);
    const int lane_ = threadIdx.x & 31, wrp_ = threadIdx.x >> 5, pairs_ = blockDim.x >> 6;
    const bool view = wrp_ >= pairs_;
    const int pair_ = view ? wrp_ - pairs_ : wrp_;
    const int car_raw = blockIdx.x * (pairs_ * 32) + pair_ * 32 + lane_;
    __shared__ float s_pose[2][3][32];             // hull origin x, y and angle, solver warp -> view warp
    // named barriers of a warp pair (64 threads; every thread of both warps executes them, live or not)
#define PAIR_SYNC(k) asm volatile("bar.sync %0, 64;" :: "r"(1 + 3 * pair_ + (k)) : "memory")
#define PAIR_ARRIVE(k) asm volatile("bar.arrive %0, 64;" :: "r"(1 + 3 * pair_ + (k)) : "memory")
    tl_stamp(b.timeline, cls == 2 ? TL_POST2 : TL_POST);
    bool live = car_raw < d.N;
    const int car = live ? car_raw : d.N - 1;
    const int env = car / d.A;
    if (mask && !mask[env]) live = false;
    const bool coupled = b.n_manifold[env] > 0;
    if (cls && (cls == 2) != coupled) live = false;
    const int N = d.N;
    const float* sc = b.scratch + car;
#ifdef MCR_PHASE_CLOCKS
    const int wid_ = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
#define PWT(k) do { if (cls != 2 && (threadIdx.x & 31) == 0 && wid_ < 1024) g_post_warp_t[k][wid_] = mcr_globaltimer(); } while (0)
#else
#define PWT(k) do {} while (0)
#endif
    PWT(0);
    if (wait_sweep > 1 && live) {
        // the contact pass and the stripes end long before the sweep: their five flags are polled together, first
        for (;;) {
            int ok = flag_peek(b.ready + 3 * N + car);
#pragma unroll
            for (int k = 0; k < 4; ++k) ok &= flag_peek(b.ready + (size_t)(4 + k) * N + car);
            if (ok) break;
            __nanosleep(200);
        }
        flag_fence_acquire();
    }
    // what does not come from the sweep is loaded before the wait for it (the step's L2-cold loads leave the critical chain)
    float cx[5], cy[5], ang[5], slp[5];
    bool awake[5];
    float motorMassK[4];
    int lim[4];
    uint8_t on_road_next[4];
    double t_old = 0.0, reward_now = 0.0;
    uint8_t backward_now = 0;
    int steps_now = 0;
    if (live && !view) {
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const float* p = b.body + (size_t)(i * BODY_FIELDS) * N + car;
            cx[i] = p[(size_t)BF_CX * N]; cy[i] = p[(size_t)BF_CY * N]; ang[i] = p[(size_t)BF_A * N];
            slp[i] = b.sleep_time[(size_t)i * N + car];
            awake[i] = b.awake[(size_t)i * N + car] != 0;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            motorMassK[k] = sc[(size_t)(SC_JOINT + k * SC_JOINT_FIELDS + 13) * N];
            lim[k] = b.limit_state[(size_t)k * N + car];
            on_road_next[k] = b.on_road_next[(size_t)k * N + car];      // stored at the end
        }
    }
    int epoch = 0;
    if (live && view) {
        t_old = b.time[car]; reward_now = b.reward[car]; backward_now = b.backward[car]; steps_now = b.steps[car];
        epoch = flag_peek(b.ready + READY_EPOCH(N));       // (the contact pass that counted it is complete: flags above / stream order)
    }
    PWT(1);
    if (wait_sweep) {
        if (live) { while (flag_peek(b.ready + 2 * N + car) == 0) __nanosleep(60); flag_fence_acquire(); }
        PAIR_SYNC(0);                              // both threads of a car have seen the flags before they are taken back
    }
    PWT(2);
    if (live && !view) {
        // (also without waits -- the producers are complete then: a car that changes class between steps must not find
        // the flags of an earlier pass)
        b.ready[2 * N + car] = 0; b.ready[3 * N + car] = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) b.ready[(size_t)(4 + k) * N + car] = 0;
    }
    // the projector may be launched once every CTA is here (everything this kernel waits for is done by then, so the
    // projector's grid cannot hold back a launch this kernel depends on): its CTAs wait for the ready flags of the cars
    // they draw (set below) -- a frame is projected as soon as its own env is done, not when this kernel's slowest car is
    cudaTriggerProgrammaticLaunchCompletion();
    const float h = (float)(1.0 / 50);
    const float mA = cc.hull_invMass, iA = cc.hull_invI, mB = cc.wheel_invMass, iB = cc.wheel_invI;
    (void)mA; (void)iA; (void)mB; (void)iB;
    float hull_px = 0.0f, hull_py = 0.0f, hull_ang = 0.0f;       // the solver thread's result the view thread needs
    float vx[5], vy[5], w[5], qs[5], qc[5], px[5], py[5];
    float jix[4], jiy[4], jiz[4], jmot[4];

    if (live && !view) {
        // ================================ solver thread ==============================================
#pragma unroll
        for (int i = 0; i < 5; ++i) { vx[i] = sc[(size_t)(SC_VX + i) * N]; vy[i] = sc[(size_t)(SC_VY + i) * N]; w[i] = sc[(size_t)(SC_W + i) * N]; }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            jix[k] = sc[(size_t)(SC_JIX + k) * N]; jiy[k] = sc[(size_t)(SC_JIY + k) * N];
            jiz[k] = sc[(size_t)(SC_JIZ + k) * N]; jmot[k] = sc[(size_t)(SC_JMOT + k) * N];
        }
        PK(1);
        if (coupled) {
            // coupled_kernel already integrated, position-solved and took the sleep decision for this car
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                slp[i] = sc[(size_t)(SC_SLP + i) * N]; awake[i] = sc[(size_t)(SC_AWAKE + i) * N] != 0.0f;
                rot_set(ang[i], qs[i], qc[i]);
                float lx = i == 0 ? cc.hull_lcx : 0.0f, ly = i == 0 ? cc.hull_lcy : 0.0f;
                px[i] = cx[i] - (qc[i] * lx - qs[i] * ly);
                py[i] = cy[i] - (qs[i] * lx + qc[i] * ly);
            }
        } else {
            if (!awake[0]) { awake[0] = true; slp[0] = 0.0f; }           // island DFS wakes the hull
            // integrate positions
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                float tx = h * vx[i], ty = h * vy[i];
                if (tx * tx + ty * ty > B2_MAX_TRANSLATION * B2_MAX_TRANSLATION) {
                    float ratio = B2_MAX_TRANSLATION / sqrtf(tx * tx + ty * ty);
                    vx[i] = ratio * vx[i]; vy[i] = ratio * vy[i];
                }
                float rotn = h * w[i];
                if (rotn * rotn > B2_MAX_ROTATION * B2_MAX_ROTATION) {
                    float ratio = B2_MAX_ROTATION / fabsf(rotn);
                    w[i] *= ratio;
                }
                cx[i] += h * vx[i]; cy[i] += h * vy[i];
                ang[i] += h * w[i];
            }
            // SolvePositionConstraints, up to 60 iterations with Box2D's early exit
            bool positionSolved = false;
#ifdef MCR_PHASE_CLOCKS
            int n_iter_ = 0;
#endif
            for (int it = 0; it < MCR_POS_ITERS; ++it) {
#ifdef MCR_PHASE_CLOCKS
                ++n_iter_;
#endif
                bool jointsOkay = true;
                // an iteration that leaves every position bit-identical is a fixed point of the remaining
                // ones (e.g. a limit error that sits exactly at the angular slop): stopping there is exact
                float p_cx[5], p_cy[5], p_an[5];
#pragma unroll
                for (int i = 0; i < 5; ++i) { p_cx[i] = cx[i]; p_cy[i] = cy[i]; p_an[i] = ang[i]; }
                jointsOkay = joints_solve_pos(cc, cx, cy, ang, lim, motorMassK);
                if (jointsOkay) { positionSolved = true; break; }
                unsigned moved = 0u;
#pragma unroll
                for (int i = 0; i < 5; ++i)
                    moved |= (__float_as_uint(p_cx[i]) ^ __float_as_uint(cx[i])) | (__float_as_uint(p_cy[i]) ^ __float_as_uint(cy[i])) |
                             (__float_as_uint(p_an[i]) ^ __float_as_uint(ang[i]));
                if (moved == 0u) break;
            }
#ifdef MCR_PHASE_CLOCKS
            atomicAdd(&g_post_iter_hist[n_iter_ < 63 ? n_iter_ : 63], 1ull);
#endif
            PK(2);
            // SynchronizeTransform + sleep
            float minSleepTime = 3.402823466e+38f;
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                rot_set(ang[i], qs[i], qc[i]);
                float lx = i == 0 ? cc.hull_lcx : 0.0f, ly = i == 0 ? cc.hull_lcy : 0.0f;
                px[i] = cx[i] - (qc[i] * lx - qs[i] * ly);
                py[i] = cy[i] - (qs[i] * lx + qc[i] * ly);
                if (w[i] * w[i] > B2_ANGULAR_SLEEP_TOL * B2_ANGULAR_SLEEP_TOL ||
                    vx[i] * vx[i] + vy[i] * vy[i] > B2_LINEAR_SLEEP_TOL * B2_LINEAR_SLEEP_TOL) {
                    slp[i] = 0.0f; minSleepTime = 0.0f;
                } else {
                    slp[i] += h; minSleepTime = fminf(minSleepTime, slp[i]);
                }
            }
            if (minSleepTime >= B2_TIME_TO_SLEEP && positionSolved) {
#pragma unroll
                for (int i = 0; i < 5; ++i) { awake[i] = false; slp[i] = 0.0f; vx[i] = 0.0f; vy[i] = 0.0f; w[i] = 0.0f; }
            }
        }   // !coupled
        PK(3);
        hull_px = px[0]; hull_py = py[0]; hull_ang = ang[0];
    }
    if (!view) {
        // hand the hull pose to the view warp and go on storing (arrive, no wait)
        s_pose[pair_][0][lane_] = hull_px; s_pose[pair_][1][lane_] = hull_py; s_pose[pair_][2][lane_] = hull_ang;
        __threadfence_block();
        PAIR_ARRIVE(1);
    }
    if (live && !view) {
        // ---- store ---------------------------------------------------------------------------
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            float* p = b.body + (size_t)(i * BODY_FIELDS) * N + car;
            p[(size_t)BF_CX * N] = cx[i]; p[(size_t)BF_CY * N] = cy[i]; p[(size_t)BF_A * N] = ang[i];
            p[(size_t)BF_VX * N] = vx[i]; p[(size_t)BF_VY * N] = vy[i]; p[(size_t)BF_W * N] = w[i];
            p[(size_t)BF_PX * N] = px[i]; p[(size_t)BF_PY * N] = py[i];
            p[(size_t)BF_QS * N] = qs[i]; p[(size_t)BF_QC * N] = qc[i];
            b.sleep_time[(size_t)i * N + car] = slp[i];
            b.awake[(size_t)i * N + car] = awake[i] ? 1 : 0;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float* p = b.joint + (size_t)(k * JOINT_FIELDS) * N + car;
            p[(size_t)JF_IX * N] = jix[k]; p[(size_t)JF_IY * N] = jiy[k]; p[(size_t)JF_IZ * N] = jiz[k];
            p[(size_t)JF_MOTOR * N] = jmot[k];
            // the contact pass of THIS step decides the friction of the NEXT Car.step
            b.on_road[(size_t)k * N + car] = on_road_next[k];
        }
        PK(4);
    }

    // ================================ view lane ====================================================
    // what the camera and the backward test read of the hull: its FINAL velocity.  The view lane restates the few
    // operations that produce it (integration clamp, island-wide sleep) from the same inputs, bit for bit.
    double angle_fast = 0.0, cos_a = 1.0, sin_a = 0.0, cos_t = 1.0, sin_t = 0.0;   // cos / sin of the deg -> rad rounded angle and of the angle itself
    bool fast = false, view_ready = false;
    double t_new = 0.0;
    if (live && view) {
        float hvxf = sc[(size_t)(SC_VX + 0) * N], hvyf = sc[(size_t)(SC_VY + 0) * N];
        const bool counted = has_action && !(noact && noact[env]);
        bool asleep;
        if (coupled) {
            asleep = sc[(size_t)(SC_AWAKE + 0) * N] == 0.0f;      // coupled_kernel: final velocities are in scratch, zeroed if asleep
        } else {
            // integration clamp of the hull (b2_maxTranslation), then the island's sleep decision: velocities are zeroed when
            // every body has been below the sleep tolerances for b2_timeToSleep AND the position solve converged.  The
            // view lane cannot know the latter early, so for the (rare) cars about to fall asleep it waits for the solver
            // lane's verdict instead (fast = false path below is then taken with the final angle anyway).
            const float tx = h * hvxf, ty = h * hvyf;
            if (tx * tx + ty * ty > B2_MAX_TRANSLATION * B2_MAX_TRANSLATION) {
                const float ratio = B2_MAX_TRANSLATION / sqrtf(tx * tx + ty * ty);
                hvxf = ratio * hvxf; hvyf = ratio * hvyf;
            }
            asleep = false;
        }
        const double hvx = hvxf, hvy = hvyf;
        fast = !asleep && sqrt(hvx * hvx + hvy * hvy) > 0.5;
        if (fast) {
            // a hull faster than 0.5 is far above b2_linearSleepTolerance (0.01): the island cannot fall asleep this step,
            // so this velocity is final whatever the solver lane decides
            angle_fast = atan2(hvx, hvy);
            const float fdeg = (float)(57.29577951308232 * angle_fast);
            const double rad = (double)fdeg * (3.14159265358979323846 / 180.0);
            cos_a = cos(rad); sin_a = sin(rad);
            cos_t = cos(angle_fast); sin_t = sin(angle_fast);
            view_ready = true;
        }
        // what the reference's render() of this step sees: env.reward before `reward -= 0.1` (score label)
        // and the backward flag of the previous step (mcr:431 precedes mcr:436-495)
        b.score_snap[car] = reward_now;
        b.backward_snap[car] = backward_now;
        t_new = t_old + 1.0 / 50;                                  // mcr:429
        b.time[car] = t_new;
        if (counted) b.steps[car] = steps_now + 1;                 // TimeLimit counts step() calls only
    }
    float f_px = 0.0f, f_py = 0.0f, f_ang = 0.0f;
    if (view) {
        PAIR_SYNC(1);                              // the solver warp's hull poses are in shared memory
        f_px = s_pose[pair_][0][lane_]; f_py = s_pose[pair_][1][lane_]; f_ang = s_pose[pair_][2][lane_];
    }
    if (live && view) {
        // camera, mcr:540-556 + Transform.enable + glViewport(0,0,96,96) under glOrtho(0,1000,0,800)
        const double SCALE = 6.0, ZOOM = 2.7, WINDOW_W = 1000, WINDOW_H = 800;
        const double zoom = 0.1 * SCALE * fmax(1 - t_new, 0.0) + ZOOM * SCALE * fmin(t_new, 1.0);
        const double scroll_x = f_px, scroll_y = f_py;
        double angle = -(double)f_ang;
        if (fast) angle = angle_fast;
        if (!view_ready) {
            const float fdeg = (float)(57.29577951308232 * angle);
            const double rad = (double)fdeg * (3.14159265358979323846 / 180.0);
            cos_a = cos(rad); sin_a = sin(rad);
            cos_t = cos(angle); sin_t = sin(angle);
        }
        const double tx = WINDOW_W / 2 - (scroll_x * zoom * cos_t - scroll_y * zoom * sin_t);
        const double ty = WINDOW_H * h_ratio - (scroll_x * zoom * sin_t + scroll_y * zoom * cos_t);
        const float ftx = (float)tx, fty = (float)ty, fzoom = (float)zoom;
        const double cs = cos_a, sn = sin_a;
        const double SX = 96.0 / 1000.0, SY = 96.0 / 800.0;
        b.camera[(size_t)0 * N + car] = (float)(cs * (double)fzoom * SX);
        b.camera[(size_t)1 * N + car] = (float)(-sn * (double)fzoom * SX);
        b.camera[(size_t)2 * N + car] = (float)((double)ftx * SX);
        b.camera[(size_t)3 * N + car] = (float)(sn * (double)fzoom * SY);
        b.camera[(size_t)4 * N + car] = (float)(cs * (double)fzoom * SY);
        b.camera[(size_t)5 * N + car] = (float)((double)fty * SY);
        // car_angle of the backward test, mcr:449-456
        const double PI = 3.141592653589793;
        double car_angle = fast ? -angle_fast : (double)f_ang;
        car_angle = fmod(car_angle + 2 * PI, 2 * PI);
        if (car_angle != 0 && car_angle < 0) car_angle += 2 * PI;
        b.heading[car] = car_angle;
    }
    PK(5);
    // both threads of the car have stored: publish it (ready[car], see DevBuffers)
    PAIR_SYNC(2);                                  // (barrier + release store: the release is cumulative over the solver thread's stores)
    if (live && view) flag_release(b.ready + car, epoch);
    PWT(3);
#undef PAIR_SYNC
#undef PAIR_ARRIVE
    if (cls != 2 && threadIdx.x == 0) atomicMax(b.timeline + TL_POST_END, mcr_globaltimer());
#ifdef MCR_PHASE_CLOCKS
    if (threadIdx.x == 0) atomicAdd(&g_post_clk[7], 1ull);
#endif
}

// Car.draw's wheel stripe (gym car_dynamics: a1 = phase, a2 = phase + 1.2, ...), evaluated once per
// wheel for all A views: 4 fp64 sin/cos of a possibly large argument per wheel.  One THREAD per
// (wheel, car) instead of a serial tail of post_kernel; phase is final once pre_kernel has run, so
// mcr_step issues this beside the solver (side stream) and the rasteriser finds the result ready.
__global__ void __launch_bounds__(128)
stripe_kernel(Dims d, DevBuffers b, const uint8_t* __restrict__ mask, int set_flags) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int N = d.N;
    if (idx >= 4 * N) return;
    tl_stamp(b.timeline, TL_STRIPES);
    const int k = idx / N, car = idx - k * N;
    if (mask && !mask[car / d.A]) return;
    const double SIZE = 0.02;
    const double phase = b.wheel[(size_t)(k * WHEEL_FIELDS + WF_PHASE) * N + car];
    const double a1 = phase, a2 = phase + 1.2;
    double s1 = sin(a1), s2 = sin(a2), c1 = cos(a1), c2 = cos(a2);
    float y1 = __int_as_float(0x7fc00000), y2 = 0.0f;
    if (!(s1 > 0 && s2 > 0)) {
        if (s1 > 0) c1 = sign_d(c1);
        if (s2 > 0) c2 = sign_d(c2);
        y1 = (float)(+27 * c1 * SIZE); y2 = (float)(+27 * c2 * SIZE);
    }
    b.stripe[(size_t)(k * 2 + 0) * N + car] = y1;
    b.stripe[(size_t)(k * 2 + 1) * N + car] = y2;
    if (set_flags) flag_release(b.ready + (size_t)(4 + k) * N + car, 1);      // ready[(4 + wheel) N + car], taken back by post_kernel
}

int launch_stripes(const Dims& d, const DevBuffers& b, const uint8_t* mask, void* stream, int set_flags) {
    stripe_kernel<<<(4 * d.N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d, b, mask, set_flags);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// ---------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------
int launch_contacts(const Dims& d, const DevBuffers& b, const CarConst& cc, const uint8_t* mask, void* stream, int set_flags) {
    static bool configured[64] = {};       // per device: function attributes belong to the device's context
    const size_t per_warp = (sim_smem_bytes(d.A) + 15) & ~(size_t)15;
    const size_t smem = per_warp * CT_WARPS;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
    if (!configured[dev]) {                // once, for the largest num_agents a handle may ask for
        const size_t smem_max = ((sim_smem_bytes(MCR_MAX_AGENTS) + 15) & ~(size_t)15) * CT_WARPS;
        if (cudaFuncSetAttribute(contacts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max) != cudaSuccess) return -1;
        configured[dev] = true;
    }
    contacts_kernel<<<(d.B + CT_WARPS - 1) / CT_WARPS, CT_WARPS * 32, smem, (cudaStream_t)stream>>>(d, b, cc, mask, per_warp, set_flags);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_physics(const Dims& d, const DevBuffers& b, const CarConst& cc, const uint8_t* mask, const uint8_t* noact,
                   const void* action, int action_dtype, double h_ratio, int collisions, void* stream) {
    static const int early_exit = getenv("MCR_NO_EARLY_EXIT") ? 0 : 1;   // diagnostics only
    cudaStream_t s = (cudaStream_t)stream;
    int launched = 0;
    if (collisions && d.A > 1) { if (launch_carcontacts(d, b, cc, mask, stream) < 0) return -1; ++launched; }
    const int pb = pre_block(), nb = (d.N + pb - 1) / pb;
    if (action_dtype == MCR_F64) pre_kernel<double><<<nb, pb, 0, s>>>(d, b, cc, mask, noact, (const double*)action);
    else pre_kernel<float><<<nb, pb, 0, s>>>(d, b, cc, mask, noact, (const float*)action);
    const int sb = sweep_block();
    const int packed_ctas = (d.N + sb - 1) / sb, percar_ctas = (d.N + sb / 32 - 1) / (sb / 32);
    stripe_kernel<<<(4 * d.N + 127) / 128, 128, 0, s>>>(d, b, mask, 0);
    sweep_kernel<<<packed_ctas + percar_ctas, sb, 0, s>>>(d, b, cc, mask, early_exit, packed_ctas, 0);
    launched += 3;
    if (collisions && d.A > 1) { if (launch_coupled(d, b, cc, mask, early_exit, stream) < 0) return -1; ++launched; }
    return cudaGetLastError() == cudaSuccess ? launched : -1;
}

// carcontacts -> pre (with_sweep = 0) or the sweep alone (with_sweep = 1): the pieces mcr_step's
// two-chain pipeline issues itself (coupled_kernel goes to its own stream there).
int launch_presweep(const Dims& d, const DevBuffers& b, const CarConst& cc, const uint8_t* mask, const uint8_t* noact,
                    const void* action, int action_dtype, int collisions, int with_sweep, void* stream, int set_flags) {
    static const int early_exit = (getenv("MCR_NO_EARLY_EXIT") ? 0 : 1);   // diagnostics only
    cudaStream_t s = (cudaStream_t)stream;
    int launched = 0;
    if (!with_sweep) {
        if (collisions && d.A > 1) { if (launch_carcontacts(d, b, cc, mask, stream) < 0) return -1; ++launched; }
        const int pb = pre_block(), nb = (d.N + pb - 1) / pb;
        if (action_dtype == MCR_F64) pre_kernel<double><<<nb, pb, 0, s>>>(d, b, cc, mask, noact, (const double*)action);
        else pre_kernel<float><<<nb, pb, 0, s>>>(d, b, cc, mask, noact, (const float*)action);
        ++launched;
    } else {
        const int sb = sweep_block();
        const int packed_ctas = (d.N + sb - 1) / sb, percar_ctas = (d.N + sb / 32 - 1) / (sb / 32);
        mcr_launch_pdl(sweep_kernel, dim3(packed_ctas + percar_ctas), dim3(sb), 0, s, d, b, cc, mask, early_exit, packed_ctas, set_flags);
        ++launched;
    }
    return cudaGetLastError() == cudaSuccess ? launched : -1;
}

// post_kernel must not start before the contacts pass of the same step has finished reading the
// start poses and writing on_road_next (the API joins the side stream before calling this).
int launch_physics_post(const Dims& d, const DevBuffers& b, const CarConst& cc, const uint8_t* mask, const uint8_t* noact,
                        int has_action, double h_ratio, int cls, void* stream, int wait_sweep) {
    const int pb = pre_block() < 128 ? 64 : 128, cars = pb / 2, nb = (d.N + cars - 1) / cars;      // two threads per car: a solver warp and a view warp per 32 cars
    mcr_launch_pdl(post_kernel, dim3(nb), dim3(pb), 0, (cudaStream_t)stream, d, b, cc, mask, noact, has_action, h_ratio, cls, wait_sweep);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}
