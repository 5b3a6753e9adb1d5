// trackgen.cuh -- device-side _create_track (reference gym_multi_car_racing/multi_car_racing.py:183-338, 359-364) and the
// per-env track ring built on it.  Included by trackgen.cu (batched generation, ring refill) and, through reset.cuh, by
// the kernels that respawn envs (a starved env generates its next track in place).  Everything here is __noinline__:
// the callers' hot paths only pay for a call that is almost never taken.
#pragma once
#include "mcr_internal.h"
#include <cuda_runtime.h>

#define TG_PATH_MAX 2600

#define TG_SCRATCH_BYTES ((size_t)TG_PATH_MAX * 4 * 8 + TG_PATH_MAX)

namespace {

struct DevMT {
    uint32_t* mt;      // 624 state words + position
    __device__ void regen() {
        const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MATRIX_A = 0x9908b0dfu;
        int kk; uint32_t y;
        for (kk = 0; kk < 624 - 397; ++kk) { y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER); mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u); }
        for (; kk < 623; ++kk) { y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER); mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u); }
        y = (mt[623] & UPPER) | (mt[0] & LOWER); mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
        mt[624] = 0;
    }
    __device__ uint32_t next32() {
        if (mt[624] >= 624) regen();
        uint32_t y = mt[mt[624]++];
        y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= (y >> 18);
        return y;
    }
    __device__ double next_double() { const uint32_t a = next32() >> 5, b = next32() >> 6; return (a * 67108864.0 + b) / 9007199254740992.0; }
    __device__ double uniform(double lo, double hi) { return lo + (hi - lo) * next_double(); }
};

__device__ __forceinline__ double npsign(double x) { return x > 0 ? 1.0 : (x < 0 ? -1.0 : 0.0); }
__device__ __forceinline__ int pyidx(int i, int n) { return i < 0 ? i + n : i; }

// b2PolygonShape::Set for 4 input vertices (weld, gift-wrap from the right-most vertex, CCW);
// returns the hull size (4 for every proper tile quad).
__device__ int b2_quad_set(const float* in, float* ox, float* oy) {
    float px[4], py[4]; int n = 0;
    for (int i = 0; i < 4; ++i) {
        const float vx = in[2 * i], vy = in[2 * i + 1];
        bool uniq = true;
        for (int j = 0; j < n; ++j) { const float dx = vx - px[j], dy = vy - py[j]; if (dx * dx + dy * dy < 0.5f * B2_LINEAR_SLOP) { uniq = false; break; } }
        if (uniq) { px[n] = vx; py[n] = vy; ++n; }
    }
    if (n < 3) return 0;
    int i0 = 0; float x0 = px[0];
    for (int i = 1; i < n; ++i) { const float x = px[i]; if (x > x0 || (x == x0 && py[i] < py[i0])) { i0 = i; x0 = x; } }
    int hull[5]; int m = 0, ih = i0;
    for (;;) {
        hull[m++] = ih;
        int ie = 0;
        for (int j = 1; j < n; ++j) {
            if (ie == ih) { ie = j; continue; }
            const float rx = px[ie] - px[hull[m - 1]], ry = py[ie] - py[hull[m - 1]];
            const float vx = px[j] - px[hull[m - 1]], vy = py[j] - py[hull[m - 1]];
            const float c = rx * vy - ry * vx;
            if (c < 0.0f) ie = j;
            if (c == 0.0f && vx * vx + vy * vy > rx * rx + ry * ry) ie = j;
        }
        ih = ie;
        if (ie == i0) break;
        if (m > n) return 0;
    }
    for (int i = 0; i < m; ++i) { ox[i] = px[hull[i]]; oy[i] = py[hull[i]]; }
    return m;
}

struct TgOut {
    int T, Q, i1, i2;
};

// One attempt of _create_track (mcr:183-338); same statement order as mcr_track_generate (api.cu).
// Returns T > 0, 0 = attempt failed (retry), < 0 = capacity error.
__device__ __noinline__ int tg_attempt(DevMT& rng, const Dims& d, const DevBuffers& b, int slot, double* path, unsigned char* border, TgOut& out) {
    const double PI = 3.14159265358979323846, SCALE = 6.0, TRACK_RAD = 900 / SCALE, TRACK_DETAIL_STEP = 21 / SCALE;
    const double TRACK_TURN_RATE = 0.31, TRACK_WIDTH = 40 / SCALE, BORDER = 8 / SCALE;
    const int CHECKPOINTS = 12, BORDER_MIN_COUNT = 4;
    double cp_a[CHECKPOINTS], cp_x[CHECKPOINTS], cp_y[CHECKPOINTS];
    double start_alpha = 0;
    for (int c = 0; c < CHECKPOINTS; ++c) {
        double alpha = 2 * PI * c / CHECKPOINTS + rng.uniform(0, 2 * PI * 1 / CHECKPOINTS);
        double rad = rng.uniform(TRACK_RAD / 3, TRACK_RAD);
        if (c == 0) { alpha = 0; rad = 1.5 * TRACK_RAD; }
        if (c == CHECKPOINTS - 1) { alpha = 2 * PI * c / CHECKPOINTS; start_alpha = 2 * PI * (-0.5) / CHECKPOINTS; rad = 1.5 * TRACK_RAD; }
        cp_a[c] = alpha; cp_x[c] = rad * cos(alpha); cp_y[c] = rad * sin(alpha);
    }
    double x = 1.5 * TRACK_RAD, y = 0, beta = 0;
    long long dest_i = 0; int laps = 0, no_freeze = 2500, np = 0; bool visited_other_side = false;
    for (;;) {
        double alpha = atan2(y, x);
        if (visited_other_side && alpha > 0) { ++laps; visited_other_side = false; }
        if (alpha < 0) { visited_other_side = true; alpha += 2 * PI; }
        double dest_alpha, dest_x, dest_y;
        for (;;) {
            bool failed = true;
            for (;;) {
                const int ci = (int)(dest_i % CHECKPOINTS);
                dest_alpha = cp_a[ci]; dest_x = cp_x[ci]; dest_y = cp_y[ci];
                if (alpha <= dest_alpha) { failed = false; break; }
                ++dest_i;
                if (dest_i % CHECKPOINTS == 0) break;
            }
            if (!failed) break;
            alpha -= 2 * PI;
        }
        const double r1x = cos(beta), r1y = sin(beta);
        const double p1x = -r1y, p1y = r1x;
        const double dest_dx = dest_x - x, dest_dy = dest_y - y;
        double proj = r1x * dest_dx + r1y * dest_dy;
        while (beta - alpha > 1.5 * PI) beta -= 2 * PI;
        while (beta - alpha < -1.5 * PI) beta += 2 * PI;
        const double prev_beta = beta;
        proj *= SCALE;
        if (proj > 0.3) beta -= fmin(TRACK_TURN_RATE, fabs(0.001 * proj));
        if (proj < -0.3) beta += fmin(TRACK_TURN_RATE, fabs(0.001 * proj));
        x += p1x * TRACK_DETAIL_STEP;
        y += p1y * TRACK_DETAIL_STEP;
        if (np >= TG_PATH_MAX) return -2;
        path[4 * np] = alpha; path[4 * np + 1] = prev_beta * 0.5 + beta * 0.5; path[4 * np + 2] = x; path[4 * np + 3] = y; ++np;
        if (laps > 4) break;
        if (--no_freeze == 0) break;
    }
    int i1 = -1, i2 = -1;
    int i = np;
    for (;;) {
        --i;
        if (i == 0) return 0;   // "return False  # Failed"
        const bool pass = path[4 * i] > start_alpha && path[4 * (i - 1)] <= start_alpha;
        if (pass && i2 == -1) i2 = i;
        else if (pass && i1 == -1) { i1 = i; break; }
    }
    const int n = (i2 - 1) - i1;
    if (n <= 0) return 0;
    const double* tr = path + 4 * (size_t)i1;
    {
        const double fb = tr[1], fpx = cos(fb), fpy = sin(fb);
        const double a = fpx * (tr[2] - tr[4 * (n - 1) + 2]), bq = fpy * (tr[3] - tr[4 * (n - 1) + 3]);
        const double glued = sqrt(a * a + bq * bq);
        if (glued > TRACK_DETAIL_STEP) return 0;
    }
    if (n > d.Tmax) return -3;
    for (int k = 0; k < n; ++k) {
        bool good = true; double oneside = 0;
        for (int neg = 0; neg < BORDER_MIN_COUNT; ++neg) {
            const double b1 = tr[4 * pyidx(k - neg - 0, n) + 1], b2 = tr[4 * pyidx(k - neg - 1, n) + 1];
            good &= fabs(b1 - b2) > TRACK_TURN_RATE * 0.2;
            oneside += npsign(b1 - b2);
        }
        good &= fabs(oneside) == BORDER_MIN_COUNT;
        border[k] = good ? 1 : 0;
    }
    for (int k = 0; k < n; ++k)                          // in place, like the reference (mcr:300-302)
        for (int neg = 0; neg < BORDER_MIN_COUNT; ++neg) border[pyidx(k - neg, n)] |= border[k];
    // ---- pool slot ---------------------------------------------------------------------------
    float* quad = b.trk_quad + (size_t)slot * d.Qmax * 8;
    double* quad64 = b.trk_quad64 + (size_t)slot * d.Qmax * 8;
    uint8_t* qcol = b.trk_quad_col + (size_t)slot * d.Qmax;
    int16_t* qtile = b.trk_quad_tile + (size_t)slot * d.Qmax;
    float* tile = b.trk_tile + (size_t)slot * d.Tmax * 8;
    float* aabb = b.trk_tile_aabb + (size_t)slot * d.Tmax * 4;
    double* node = b.trk_node + (size_t)slot * d.Tmax * 3;
    int q = 0;
    for (int k = 0; k < n; ++k) {
        const double* n1 = tr + 4 * (size_t)k; const double* n2 = tr + 4 * (size_t)pyidx(k - 1, n);
        const double b1 = n1[1], x1 = n1[2], y1 = n1[3], b2 = n2[1], x2 = n2[2], y2 = n2[3];
        const double c1 = cos(b1), s1 = sin(b1), c2 = cos(b2), s2 = sin(b2);
        if (q >= d.Qmax) return -4;
        float* qv = quad + (size_t)q * 8;
        double* qd = quad64 + (size_t)q * 8;
        qd[0] = x1 - TRACK_WIDTH * c1; qd[1] = y1 - TRACK_WIDTH * s1;
        qd[2] = x1 + TRACK_WIDTH * c1; qd[3] = y1 + TRACK_WIDTH * s1;
        qd[4] = x2 + TRACK_WIDTH * c2; qd[5] = y2 + TRACK_WIDTH * s2;
        qd[6] = x2 - TRACK_WIDTH * c2; qd[7] = y2 - TRACK_WIDTH * s2;
        for (int v = 0; v < 8; ++v) qv[v] = (float)qd[v];
        qcol[q] = (uint8_t)(PAL_ROAD0 + k % 3);           // 0.4 + 0.01 * (k % 3), mcr:316-317
        qtile[q] = (int16_t)k;
        {   // fd_tile.shape.vertices = ..., mcr:318 -> Box2D polygon + AABB
            float hx[4], hy[4];
            const int m = b2_quad_set(qv, hx, hy);
            if (m != 4) return 0;                         // degenerate tile: treat the attempt as failed
            float lx = hx[0], ly = hy[0], ux = lx, uy = ly;
            for (int v = 0; v < 4; ++v) {
                tile[(size_t)k * 8 + 2 * v] = hx[v]; tile[(size_t)k * 8 + 2 * v + 1] = hy[v];
                lx = fminf(lx, hx[v]); ly = fminf(ly, hy[v]); ux = fmaxf(ux, hx[v]); uy = fmaxf(uy, hy[v]);
            }
            aabb[(size_t)k * 4] = lx; aabb[(size_t)k * 4 + 1] = ly; aabb[(size_t)k * 4 + 2] = ux; aabb[(size_t)k * 4 + 3] = uy;
        }
        ++q;
        if (border[k]) {
            const double side = npsign(b2 - b1);
            if (q >= d.Qmax) return -4;
            float* bv = quad + (size_t)q * 8;
            double* bd = quad64 + (size_t)q * 8;
            bd[0] = x1 + side * TRACK_WIDTH * c1; bd[1] = y1 + side * TRACK_WIDTH * s1;
            bd[2] = x1 + side * (TRACK_WIDTH + BORDER) * c1; bd[3] = y1 + side * (TRACK_WIDTH + BORDER) * s1;
            bd[4] = x2 + side * (TRACK_WIDTH + BORDER) * c2; bd[5] = y2 + side * (TRACK_WIDTH + BORDER) * s2;
            bd[6] = x2 + side * TRACK_WIDTH * c2; bd[7] = y2 + side * TRACK_WIDTH * s2;
            for (int v = 0; v < 8; ++v) bv[v] = (float)bd[v];
            qcol[q] = (uint8_t)(k % 2 == 0 ? PAL_WHITE : PAL_RED);
            qtile[q] = (int16_t)-1;
            ++q;
        }
        node[(size_t)k * 3] = n1[1]; node[(size_t)k * 3 + 1] = n1[2]; node[(size_t)k * 3 + 2] = n1[3];
    }
    out.T = n; out.Q = q; out.i1 = i1; out.i2 = i2;
    return n;
}


// reset()'s `while True: success = self._create_track()` loop (mcr:359-364) for one track on the MT19937 stream `mt`
// (624 words + position), written into pool slot `slot` together with the culling chunks and the spawn grid.
// res[4] = { T (> 0) or an error code (< 0), attempts, i1, i2 }.  Returns res[0].
__device__ __noinline__ int tg_generate_track(const Dims& d, const DevBuffers& b, uint32_t* mt, int slot, unsigned char* scratch,
                                              int max_attempts, int32_t* res) {
    DevMT rng; rng.mt = mt;
    double* path = reinterpret_cast<double*>(scratch);
    unsigned char* border = scratch + (size_t)TG_PATH_MAX * 4 * 8;
    TgOut out; out.T = 0; out.Q = 0; out.i1 = 0; out.i2 = 0;
    int attempts = 0, rc = 0;
    while (attempts < max_attempts) {          // reset(): `while True: success = self._create_track()`, mcr:359-364
        ++attempts;
        rc = tg_attempt(rng, d, b, slot, path, border, out);
        if (rc != 0) break;
    }
    res[0] = rc > 0 ? out.T : (rc < 0 ? rc : -5); res[1] = attempts; res[2] = out.i1; res[3] = out.i2;
    if (rc <= 0) return res[0];
    b.trk_T[slot] = out.T; b.trk_Q[slot] = out.Q;
    // ---- bounding circle of every MCR_QUAD_CHUNK consecutive road_poly quads (rasteriser culling) ----
    const float* quad = b.trk_quad + (size_t)slot * d.Qmax * 8;
    const int nchunk = d.Qmax / MCR_QUAD_CHUNK;
    float* chunk = b.trk_chunk + (size_t)slot * nchunk * 4;
    for (int c = 0; c < nchunk; ++c) {
        const int q0 = c * MCR_QUAD_CHUNK, q1 = min(out.Q, q0 + MCR_QUAD_CHUNK);
        float cx = 0.0f, cy = 0.0f, cr = 0.0f;
        if (q0 < out.Q) {
            double sx = 0, sy = 0; int nvert = 0;
            for (int q = q0; q < q1; ++q) for (int k = 0; k < 4; ++k) { sx += quad[(size_t)q * 8 + 2 * k]; sy += quad[(size_t)q * 8 + 2 * k + 1]; ++nvert; }
            const double mx = sx / nvert, my = sy / nvert;
            double r2 = 0;
            for (int q = q0; q < q1; ++q) for (int k = 0; k < 4; ++k) {
                const double dx = quad[(size_t)q * 8 + 2 * k] - (double)(float)mx, dy = quad[(size_t)q * 8 + 2 * k + 1] - (double)(float)my;
                r2 = fmax(r2, dx * dx + dy * dy);
            }
            cx = (float)mx; cy = (float)my; cr = (float)(sqrt(r2) * 1.0001 + 1e-3);
        }
        chunk[(size_t)c * 4] = cx; chunk[(size_t)c * 4 + 1] = cy; chunk[(size_t)c * 4 + 2] = cr; chunk[(size_t)c * 4 + 3] = 0.0f;
    }
    // ---- spawn pose of every grid position, both directions (reset() spawn grid, mcr:366-393) ------
    const double PI = 3.14159265358979323846;
    const double* tr = path + 4 * (size_t)out.i1;
    const int A = d.A, T = out.T;
    double* pose = b.trk_slot_pose + (size_t)slot * 2 * A * 3;
    const double pos_x = tr[2], pos_y = tr[3];
    for (int cw = 0; cw < 2; ++cw)
        for (int c = 0; c < A; ++c) {
            const int line_number = c / 2, side = 2 * (c % 2) - 1;
            int idx = -line_number * 5;
            if (idx < 0) idx += T;
            if (idx < 0 || idx >= T) { res[0] = -6; return -6; }
            const double dx = tr[4 * (size_t)idx + 2] - pos_x, dy = tr[4 * (size_t)idx + 3] - pos_y;
            double angle = tr[4 * (size_t)idx + 1];
            if (cw) angle -= PI;
            const double norm_theta = angle - PI / 2;
            double* o = pose + ((size_t)cw * A + c) * 3;
            o[0] = angle;
            o[1] = pos_x + dx + (3.0 * sin(norm_theta) * side);
            o[2] = pos_y + dy + (3.0 * cos(norm_theta) * side);
        }
    return res[0];
}

// ---- per-env track ring (fresh tracks on auto reset): env e owns pool slots e + B * j, j = 0 .. R; episode n of the
// env runs on ring index n % (R + 1).  trk_consumed = index of the episode in progress, trk_produced = last track
// generated (single writer each); trk_lock serialises the generators (refill kernel / starved consumer).
__device__ __forceinline__ int ring_ld(const int32_t* p) { return *reinterpret_cast<const volatile int32_t*>(p); }

__device__ __noinline__ void ring_produce_one(const Dims& d, const DevBuffers& b, int env, int R) {
    // caller holds trk_lock[env]
    const int k = ring_ld(b.trk_produced + env) + 1;
    const int slot = env + d.B * (k % (R + 1));
    int32_t res[4];
    unsigned char* scratch = b.tg_scratch + (size_t)env * TG_SCRATCH_BYTES;
    const int rc = tg_generate_track(d, b, b.mt_state + (size_t)env * 625, slot, scratch, 64, res);
    if (rc <= 0) atomicMax(b.status + ST_TRACK_ERROR, -rc);   // capacity error / no valid track: the slot keeps its last track
    __threadfence();
    *reinterpret_cast<volatile int32_t*>(b.trk_produced + env) = k;
    __threadfence();
}

// Called by ONE thread when env starts its next episode: returns the pool slot of that episode's track.  If the refill
// has not delivered it yet (episodes shorter than a track generation), it is generated right here -- the sequence of
// tracks of an env never depends on timing.
__device__ __forceinline__ int ring_next_slot(const Dims& d, const DevBuffers& b, int env, int R) {
    const int c = b.trk_consumed[env] + 1;
    if (ring_ld(b.trk_produced + env) < c) {
        if (atomicCAS(b.trk_lock + env, 0, 1) == 0) {
            while (ring_ld(b.trk_produced + env) < c) ring_produce_one(d, b, env, R);
            atomicExch(b.trk_lock + env, 0);
        } else {
            // the refill kernel is generating for this env right now (it holds the lock, so it is running): wait for it
            while (ring_ld(b.trk_produced + env) < c) __nanosleep(200);
        }
        __threadfence();
    }
    b.trk_consumed[env] = c;
    return env + d.B * (c % (R + 1));
}

}  // namespace
