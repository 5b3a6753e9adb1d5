// contacts.cu -- kernel 2: b2ContactManager::Collide for the sensor tiles + FrictionDetector.
//
// Replaces (reference: gym_multi_car_racing/multi_car_racing.py):
//   mcr:84-123   FrictionDetector.BeginContact/EndContact/_contact
//   Box2D        b2Contact::Update for sensor fixtures: touching = b2TestOverlap(shapes) =
//                (GJK distance < rA + rB = 0.02), evaluated at the START of world.Step on the
//                poses the previous step left behind.
// Because road_visited makes BeginContact idempotent and tile colour only ever moves towards
// ROAD_COLOR, no per-pair contact state is needed: per step and env we evaluate the overlap
// predicate for {4 wheels + 4 hull fixtures} x A cars against every tile whose AABB is near.
//
// Mapping: one warp per env.  Phase 1: lanes build the world-space fixture polygons in shared
// memory.  Phase 2: lanes stride over the T tiles (coalesced float4 AABB loads), AABB-reject,
// run the exact predicate on survivors, update the per-tile visit bitmask (lane-private, no
// atomics).  Phase 3: new visits are appended to a shared-memory event list that lane 0
// replays in the contact-list order of a fresh b2World (tile descending, car descending) so
// that the float64 reward sums are bit-reproducible -- no racing atomicAdds.
#include "mcr_internal.h"
#include <cuda_runtime.h>

#define CT_WARPS 4
#define AABB_MARGIN 0.05f
#define FIX_STRIDE 21   // 8 x + 8 y + 4 aabb + (n | active<<8), padded to an odd stride (bank spread)

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

__global__ void __launch_bounds__(CT_WARPS * 32)
contacts_kernel(Dims d, DevBuffers b, CarConst cc, const uint8_t* __restrict__ mask) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int env = blockIdx.x * CT_WARPS + warp;
    if (env >= d.B) return;
    if (mask && !mask[env]) return;
    const int A = d.A, N = d.N, nfix = A * 8, maxev = 16 * A;
    // per-warp shared memory carve-up
    const size_t per_warp = (size_t)nfix * FIX_STRIDE * 4 + (size_t)maxev * 12 + 16;
    unsigned char* base = smem_raw + (size_t)warp * ((per_warp + 15) & ~(size_t)15);
    float* fixt = (float*)base;
    double* ev_inc = (double*)(base + (((size_t)nfix * FIX_STRIDE * 4 + 7) & ~(size_t)7));
    uint32_t* ev_key = (uint32_t*)(ev_inc + maxev);
    int* ev_count = (int*)(ev_key + maxev);
    if (lane == 0) *ev_count = 0;

    // ---- phase 1: world-space fixture polygons --------------------------------------------
    for (int f = lane; f < nfix; f += 32) {
        const int c = f >> 3, fi = f & 7, car = env * A + c;
        const int body = fi < 4 ? 1 + fi : 0;
        const Poly8& P = fi < 4 ? cc.wheel_poly : cc.hull_poly[fi - 4];
        const float* bp = b.body + (size_t)(body * BODY_FIELDS) * N + car;
        const float px = bp[(size_t)BF_PX * N], py = bp[(size_t)BF_PY * N];
        const float qs = bp[(size_t)BF_QS * N], qc = bp[(size_t)BF_QC * N];
        float* o = fixt + (size_t)f * FIX_STRIDE;
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

    // ---- phase 2: tiles ----------------------------------------------------------------------
    const int slot = b.env_track[env];
    const int T = b.trk_T[slot];
    const float4* aabbs = (const float4*)(b.trk_tile_aabb + (size_t)slot * d.Tmax * 4);
    const float* tiles = b.trk_tile + (size_t)slot * d.Tmax * 8;
    uint32_t* visited = b.visited + (size_t)env * d.Tmax;
    uint8_t* touched = b.touched + (size_t)env * d.Tmax;
    unsigned long long road_bits = 0ull;   // bit (c*4 + wheel): wheel touches >= 1 tile
    for (int t = lane; t < T; t += 32) {
        const float4 ta = aabbs[t];
        float tx[4], ty[4];
        bool have_tile = false;
        bool tile_touched = false;
        uint32_t vis = 0; bool vis_loaded = false, vis_dirty = false;
        for (int c = A - 1; c >= 0; --c) {
            bool wheel_touch = false;
            for (int fi = 7; fi >= 0; --fi) {
                const float* o = fixt + (size_t)(c * 8 + fi) * FIX_STRIDE;
                if (o[16] - ta.z > AABB_MARGIN || o[17] - ta.w > AABB_MARGIN ||
                    ta.x - o[18] > AABB_MARGIN || ta.y - o[19] > AABB_MARGIN) continue;
                if (!have_tile) {
                    const float4 v0 = *(const float4*)(tiles + (size_t)t * 8);
                    const float4 v1 = *(const float4*)(tiles + (size_t)t * 8 + 4);
                    tx[0] = v0.x; ty[0] = v0.y; tx[1] = v0.z; ty[1] = v0.w;
                    tx[2] = v1.x; ty[2] = v1.y; tx[3] = v1.z; ty[3] = v1.w;
                    have_tile = true;
                }
                const int meta = __float_as_int(o[20]);
                if (!poly_touch(tx, ty, o, o + 8, meta & 0xff)) continue;
                if (fi < 4) road_bits |= 1ull << (c * 4 + fi);
                if (!(meta >> 8)) continue;          // sleeping body: contact not updated
                tile_touched = true;                 // mcr:102-104
                if (fi < 4) wheel_touch = true;      // hull.userData is None, mcr:108
            }
            if (wheel_touch) {                       // mcr:110-120
                if (!vis_loaded) { vis = visited[t]; vis_loaded = true; }
                if (!((vis >> c) & 1u)) {
                    vis |= 1u << c; vis_dirty = true;
                    const int past = __popc(vis) - 1;
                    const double reward_factor = 1 - ((double)past / (double)A);
                    const double inc = reward_factor * 1000.0 / (double)T;
                    const int slot_e = atomicAdd(ev_count, 1);
                    if (slot_e < maxev) { ev_key[slot_e] = ((uint32_t)t << 8) | (uint32_t)c; ev_inc[slot_e] = inc; }
                    else atomicExch(&b.status[ST_EVENT_OVERFLOW], 1);
                }
            }
        }
        if (vis_dirty) visited[t] = vis;
        if (tile_touched) touched[t] = 1;
    }
    // ---- phase 3: len(wheel.tiles) > 0 flags for the next Car.step ---------------------------
    uint32_t lo = __reduce_or_sync(0xffffffffu, (uint32_t)road_bits);
    uint32_t hi = __reduce_or_sync(0xffffffffu, (uint32_t)(road_bits >> 32));
    for (int i = lane; i < A * 4; i += 32) {
        const int c = i >> 2, k = i & 3;
        const uint32_t bit = i < 32 ? (lo >> i) & 1u : (hi >> (i - 32)) & 1u;
        b.on_road_next[(size_t)k * N + env * A + c] = (uint8_t)bit;
    }
    __syncwarp();
    // ---- phase 4: replay new visits in contact-list order (tile desc, car desc) ---------------
    if (lane == 0) {
        int n = *ev_count; if (n > maxev) n = maxev;
        for (int i = 1; i < n; ++i) {       // insertion sort, descending key; n is tiny
            uint32_t k = ev_key[i]; double v = ev_inc[i]; int j = i - 1;
            while (j >= 0 && ev_key[j] < k) { ev_key[j + 1] = ev_key[j]; ev_inc[j + 1] = ev_inc[j]; --j; }
            ev_key[j + 1] = k; ev_inc[j + 1] = v;
        }
        for (int i = 0; i < n; ++i) {
            const int car = env * A + (int)(ev_key[i] & 0xffu);
            b.reward[car] += ev_inc[i];
            b.visit_count[car] += 1;
        }
    }
}

static size_t contacts_smem_bytes(int A) {
    const int nfix = A * 8, maxev = 16 * A;
    size_t per_warp = (size_t)nfix * FIX_STRIDE * 4 + (size_t)maxev * 12 + 16;
    per_warp = (per_warp + 15) & ~(size_t)15;
    return per_warp * CT_WARPS + 16;
}

int launch_contacts(const Dims& d, const DevBuffers& b, const CarConst& cc, const uint8_t* mask, void* stream) {
    static int configured_for = -1;
    const size_t smem = contacts_smem_bytes(d.A);
    if (smem > 48 * 1024 && configured_for != d.A) {
        if (cudaFuncSetAttribute(contacts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
        configured_for = d.A;
    }
    dim3 grid((d.B + CT_WARPS - 1) / CT_WARPS), block(CT_WARPS * 32);
    contacts_kernel<<<grid, block, smem, (cudaStream_t)stream>>>(d, b, cc, mask);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}
