// raster.cu -- kernel 3: software top-down rasteriser + the post-step reward/done block.
//
// Replaces, per agent view (reference: gym_multi_car_racing/multi_car_racing.py):
//   mcr:511-604  render("state_pixels") / _render_window: camera from ego pose/velocity/t,
//                glViewport(0,0,96,96) under glOrtho(0,1000,0,800), colour-buffer read-back,
//                row flip, alpha drop
//   mcr:613-632  render_road (playfield quad, 100 checker quads, road + border quads)
//   gym Car.draw (wheels, wheel stripes from phase, 4 hull fixtures), all cars in every view
//   mcr:634-674  render_indicators (HUD bar, 7 indicator quads, score, backward flag)
//   mcr:433-507  reward -= 0.1, step_reward, driving_backward, done / out-of-field (post_step)
// OpenGL's fixed-function fill is restated as point sampling at pixel centres with half-open
// spans, every edge evaluated from its lower to its upper endpoint (shared edges are
// watertight and bit-reproducible); painter's order is the reference's draw order.
//
// Mapping: one CTA (256 threads) per agent-frame.
//   1. candidate polygons are generated one per thread in painter's order, projected with the
//      fp32 affine camera, culled against the 96x96 viewport and compacted IN ORDER into a
//      shared-memory display list (block scan);
//   2. span generation: warp per polygon, lane per row -> (x0,x1) byte pairs in a shared span
//      pool + a per-row polygon bitmask;
//   3. fill: warp per row walks that row's bitmask top-down with early exit once all 96 pixels
//      are resolved, writing 1-byte palette indices into a 96x96 shared tile;
//   4. the tile is expanded to RGB and written to HBM with 16-byte vector stores (1728 uint4
//      per frame) -- the only HBM traffic that scales with the frame count.
#include "mcr_internal.h"
#include <cuda_runtime.h>

#define RS_THREADS 256
#define RS_WARPS (RS_THREADS / 32)
#define LIST_CAP 384
#define MASK_WORDS (LIST_CAP / 32)
#define SPAN_POOL 4096
#define N_CHECKER 400   // range(-20, 20, 2) squared
#define SW MCR_STATE_W
#define SH MCR_STATE_H

// initial values are documentation; launch_render overwrites them with mcr_host_palette()
__constant__ uint8_t c_palette[PAL_COUNT][4] = {
    {0, 0, 0, 0},       // PAL_BLACK
    {102, 204, 102, 0}, // PAL_GRASS        (0.4, 0.8, 0.4)
    {102, 230, 102, 0}, // PAL_GRASS_LIGHT  (0.4, 0.9, 0.4): 0.9f*255 = 229.5 in fp32 -> 230
    {102, 102, 102, 0}, // PAL_ROAD0        0.40
    {105, 105, 105, 0}, // PAL_ROAD1        0.41
    {107, 107, 107, 0}, // PAL_ROAD2        0.42
    {255, 255, 255, 0}, // PAL_WHITE
    {255, 0, 0, 0},     // PAL_RED
    {77, 77, 77, 0},    // PAL_WHEEL_WHITE  (0.3, 0.3, 0.3)
    {204, 0, 0, 0}, {0, 0, 204, 0}, {0, 204, 0, 0}, {0, 204, 204, 0},      // CAR_COLORS, mcr:67-70
    {204, 204, 204, 0}, {0, 0, 0, 0}, {204, 0, 204, 0}, {204, 204, 0, 0},
    {0, 0, 255, 0},     // PAL_IND_BLUE     (0, 0, 1)
    {51, 0, 255, 0},    // PAL_IND_BLUE2    (0.2, 0, 1)
    {0, 255, 0, 0},     // PAL_IND_GREEN
    {0, 0, 255, 0},     // PAL_FLAG_BLUE    c3B (0, 0, 255)
};

// Host copy of the palette, derived from the reference's float colours with the GL rule
// u8 = floor(c * 255 + 0.5) evaluated in fp32; launch_render checks c_palette against it.
static const float h_palette_f[PAL_COUNT][3] = {
    {0, 0, 0}, {0.4f, 0.8f, 0.4f}, {0.4f, 0.9f, 0.4f}, {0.4f, 0.4f, 0.4f}, {0.41f, 0.41f, 0.41f}, {0.42f, 0.42f, 0.42f},
    {1, 1, 1}, {1, 0, 0}, {0.3f, 0.3f, 0.3f},
    {0.8f, 0.0f, 0.0f}, {0.0f, 0.0f, 0.8f}, {0.0f, 0.8f, 0.0f}, {0.0f, 0.8f, 0.8f},
    {0.8f, 0.8f, 0.8f}, {0.0f, 0.0f, 0.0f}, {0.8f, 0.0f, 0.8f}, {0.8f, 0.8f, 0.0f},
    {0, 0, 1}, {0.2f, 0, 1}, {0, 1, 0}, {0, 0, 1}};

const uint8_t (*mcr_host_palette())[4] {
    static uint8_t pal[PAL_COUNT][4];
    for (int i = 0; i < PAL_COUNT; ++i) {
        for (int c = 0; c < 3; ++c) pal[i][c] = (uint8_t)(int)floorf(h_palette_f[i][c] * 255.0f + 0.5f);
        pal[i][3] = 0;
    }
    return pal;
}

// 3x5 digit font for the score label (documented deviation D3: the reference uses the
// platform font through pyglet).  Row-major, msb = left column.
__constant__ uint8_t c_font[11][5] = {
    {7, 5, 5, 5, 7}, {2, 6, 2, 2, 7}, {7, 1, 7, 4, 7}, {7, 1, 7, 1, 7}, {5, 5, 7, 1, 1}, {7, 4, 7, 1, 7},
    {7, 4, 7, 5, 7}, {7, 1, 1, 1, 1}, {7, 5, 7, 5, 7}, {7, 5, 7, 1, 7}, {0, 0, 7, 0, 0}};

struct Affine { float m00, m01, m02, m10, m11, m12; };

struct __align__(16) RasterSmem {
    float vx[MCR_MAXV][LIST_CAP];
    float vy[MCR_MAXV][LIST_CAP];
    uint16_t off[LIST_CAP];
    uint8_t n[LIST_CAP], col[LIST_CAP], y0[LIST_CAP], y1[LIST_CAP];
    uchar2 span[SPAN_POOL];
    uint32_t rowmask[SH][MASK_WORDS];
    uint8_t img[SH * SW];            // palette indices, row 0 = TOP row of the observation
    Affine M;
    int list_count, pool_count, first_bad;
    int warp_cnt[RS_WARPS], warp_rows[RS_WARPS];
    double red_d[RS_WARPS]; int red_i[RS_WARPS];
    char glyph[4];
    uint32_t pal32[32];
};

__device__ __forceinline__ double sign_d(double x) { return x > 0 ? 1.0 : (x < 0 ? -1.0 : 0.0); }
__device__ __forceinline__ double py_mod(double a, double m) {
    double r = fmod(a, m);
    if (r != 0 && ((r < 0) != (m < 0))) r += m;
    return r;
}

struct View {              // everything candidate generation needs, in registers/params
    int env, agent, A, N, Q, slot;
    const float* body; const double* wheel; const float* stripe;
    const float* quad; const uint8_t* quad_col; const int16_t* quad_tile; const uint8_t* touched;
    int use_ego_color, backward_flag_on;
};

__device__ __forceinline__ void xf_pt(const Affine& M, float x, float y, float& ox, float& oy) {
    ox = (M.m00 * x + M.m01 * y) + M.m02;
    oy = (M.m10 * x + M.m11 * y) + M.m12;
}

// Generate candidate polygon i (painter's order).  Returns vertex count (0 = nothing to draw).
__device__ int gen_candidate(int i, const View& V, const Affine& M, const CarConst& cc, float* px, float* py, int& col) {
    const double PLAYFIELD = 2000 / 6.0;
    if (i == 0) {                                   // playfield, mcr:615-619
        const float pf = (float)PLAYFIELD;
        xf_pt(M, -pf, +pf, px[0], py[0]); xf_pt(M, +pf, +pf, px[1], py[1]);
        xf_pt(M, +pf, -pf, px[2], py[2]); xf_pt(M, -pf, -pf, px[3], py[3]);
        col = PAL_GRASS; return 4;
    }
    i -= 1;
    if (i < N_CHECKER) {                            // 20 x 20 checker quads, mcr:620-627
        const double k = PLAYFIELD / 20.0;
        const int x = -20 + 2 * (i / 20), y = -20 + 2 * (i % 20);
        xf_pt(M, (float)(k * x + k), (float)(k * y + 0), px[0], py[0]);
        xf_pt(M, (float)(k * x + 0), (float)(k * y + 0), px[1], py[1]);
        xf_pt(M, (float)(k * x + 0), (float)(k * y + k), px[2], py[2]);
        xf_pt(M, (float)(k * x + k), (float)(k * y + k), px[3], py[3]);
        col = PAL_GRASS_LIGHT; return 4;
    }
    i -= N_CHECKER;
    if (i < V.Q) {                                  // road_poly, mcr:628-631
        const float4 a = *(const float4*)(V.quad + (size_t)i * 8);
        const float4 b = *(const float4*)(V.quad + (size_t)i * 8 + 4);
        xf_pt(M, a.x, a.y, px[0], py[0]); xf_pt(M, a.z, a.w, px[1], py[1]);
        xf_pt(M, b.x, b.y, px[2], py[2]); xf_pt(M, b.z, b.w, px[3], py[3]);
        const int tl = V.quad_tile[i];
        col = (tl >= 0 && V.touched[tl]) ? PAL_ROAD0 : V.quad_col[i];   // tile.color reset, mcr:102-104
        return 4;
    }
    i -= V.Q;
    if (i < 12 * V.A) {                             // Car.draw for every car, mcr:559-564
        const int c = i / 12, part = i % 12, car = V.env * V.A + c, N = V.N;
        if (part < 8) {
            const int wl = part >> 1;
            const float* bp = V.body + (size_t)((1 + wl) * BODY_FIELDS) * N + car;
            const float bx = bp[(size_t)BF_PX * N], by = bp[(size_t)BF_PY * N];
            const float qs = bp[(size_t)BF_QS * N], qc = bp[(size_t)BF_QC * N];
            float lx[4], ly[4];
            if ((part & 1) == 0) {
                for (int k = 0; k < 4; ++k) { lx[k] = cc.wheel_poly.x[k]; ly[k] = cc.wheel_poly.y[k]; }
                col = PAL_BLACK;
            } else {
                // stripe quad y-extents were evaluated once per wheel by the physics kernel
                // (Car.draw: a1 = phase, a2 = phase + 1.2 ...); NaN = "not drawn this frame"
                const float sy1 = V.stripe[(size_t)(wl * 2 + 0) * N + car], sy2 = V.stripe[(size_t)(wl * 2 + 1) * N + car];
                if (sy1 != sy1) return 0;
                const float hw = cc.wheel_poly.x[0] < 0 ? -cc.wheel_poly.x[0] : cc.wheel_poly.x[0];   // (float)(WHEEL_W*SIZE)
                lx[0] = -hw; ly[0] = sy1; lx[1] = +hw; ly[1] = sy1;
                lx[2] = +hw; ly[2] = sy2; lx[3] = -hw; ly[3] = sy2;
                col = PAL_WHEEL_WHITE;
            }
            for (int k = 0; k < 4; ++k) {
                const float wx = (qc * lx[k] - qs * ly[k]) + bx, wy = (qs * lx[k] + qc * ly[k]) + by;
                xf_pt(M, wx, wy, px[k], py[k]);
            }
            return 4;
        } else {
            const int f = 3 - (part - 8);           // body.fixtures iterates newest first
            const float* bp = V.body + (size_t)(0 * BODY_FIELDS) * N + car;
            const float bx = bp[(size_t)BF_PX * N], by = bp[(size_t)BF_PY * N];
            const float qs = bp[(size_t)BF_QS * N], qc = bp[(size_t)BF_QC * N];
            const Poly8& P = cc.hull_poly[f];
            for (int k = 0; k < P.n; ++k) {
                const float wx = (qc * P.x[k] - qs * P.y[k]) + bx, wy = (qs * P.x[k] + qc * P.y[k]) + by;
                xf_pt(M, wx, wy, px[k], py[k]);
            }
            if (V.use_ego_color) col = (c == V.agent) ? PAL_CAR0 + 0 : PAL_CAR0 + 1;   // mcr:560-563
            else col = PAL_CAR0 + (c % 8);                                              // mcr:402
            return P.n;
        }
    }
    i -= 12 * V.A;
    {                                               // render_indicators, mcr:634-674
        const double Wd = 1000, Hd = 800, s = Wd / 40.0, h = Hd / 40.0;
        const int car = V.env * V.A + V.agent, N = V.N;
        double wx[4], wy[4]; int n = 4;
        if (i == 0) {
            wx[0] = Wd; wy[0] = 0; wx[1] = Wd; wy[1] = 5 * h; wx[2] = 0; wy[2] = 5 * h; wx[3] = 0; wy[3] = 0;
            col = PAL_BLACK;
        } else if (i <= 5) {
            double place, val;
            if (i == 1) {
                const double lvx = V.body[(size_t)BF_VX * N + car], lvy = V.body[(size_t)BF_VY * N + car];
                place = 5; val = 0.02 * sqrt(lvx * lvx + lvy * lvy); col = PAL_WHITE;
            } else {
                const int wl = i - 2;
                place = 7 + wl; val = 0.01 * V.wheel[(size_t)(wl * WHEEL_FIELDS + WF_OMEGA) * N + car];
                col = wl < 2 ? PAL_IND_BLUE : PAL_IND_BLUE2;
            }
            wx[0] = (place + 0) * s; wy[0] = h + h * val; wx[1] = (place + 1) * s; wy[1] = h + h * val;
            wx[2] = (place + 1) * s; wy[2] = h;           wx[3] = (place + 0) * s; wy[3] = h;
        } else if (i <= 7) {
            double place, val;
            if (i == 6) {
                const float a0 = V.body[(size_t)(0 * BODY_FIELDS + BF_A) * N + car];
                const float a1 = V.body[(size_t)(1 * BODY_FIELDS + BF_A) * N + car];
                place = 20; val = -10.0 * (double)((a1 - a0) - 0.0f); col = PAL_IND_GREEN;
            } else {
                place = 30; val = -0.8 * (double)V.body[(size_t)BF_W * N + car]; col = PAL_RED;
            }
            wx[0] = (place + 0) * s;   wy[0] = 4 * h; wx[1] = (place + val) * s; wy[1] = 4 * h;
            wx[2] = (place + val) * s; wy[2] = 2 * h; wx[3] = (place + 0) * s;   wy[3] = 2 * h;
        } else {
            if (!V.backward_flag_on) return 0;
            wx[0] = Wd - 100; wy[0] = 30; wx[1] = Wd - 75; wy[1] = 70; wx[2] = Wd - 50; wy[2] = 30; n = 3;
            col = PAL_FLAG_BLUE;
        }
        for (int k = 0; k < n; ++k) {
            px[k] = (float)wx[k] * (float)(96.0 / 1000.0);
            py[k] = (float)wy[k] * (float)(96.0 / 800.0);
        }
        return n;
    }
}

__device__ void flush_list(RasterSmem& S, int tid) {
    const int warp = tid >> 5, lane = tid & 31;
    __syncthreads();
    const int n = S.list_count;
    // ---- span generation: warp per polygon, lane per row ---------------------------------
    for (int p = warp; p < n; p += RS_WARPS) {
        const int nv = S.n[p], y0 = S.y0[p], y1 = S.y1[p], off = S.off[p];
        for (int y = y0 + lane; y < y1; y += 32) {
            const float yc = (float)y + 0.5f;
            float xl = 3.402823466e+38f, xr = -3.402823466e+38f;
            for (int i = 0; i < nv; ++i) {
                const int k = i + 1 < nv ? i + 1 : 0;
                float ax = S.vx[i][p], ay = S.vy[i][p], bx = S.vx[k][p], by = S.vy[k][p];
                if (ay == by) continue;
                if (ay > by) { float t = ax; ax = bx; bx = t; t = ay; ay = by; by = t; }
                if (!(yc >= ay && yc < by)) continue;
                const float x = ax + (yc - ay) * ((bx - ax) / (by - ay));
                xl = fminf(xl, x); xr = fmaxf(xr, x);
            }
            int x0 = 0, x1 = 0;
            if (xl < xr) {
                xl = fminf(fmaxf(xl, -1.0f), (float)SW + 1.0f);
                xr = fminf(fmaxf(xr, -1.0f), (float)SW + 1.0f);
                x0 = (int)ceilf(xl - 0.5f); if (x0 < 0) x0 = 0;
                x1 = (int)ceilf(xr - 0.5f); if (x1 > SW) x1 = SW;
            }
            S.span[off + (y - y0)] = make_uchar2((unsigned char)x0, (unsigned char)x1);
            if (x0 < x1) atomicOr(&S.rowmask[y][p >> 5], 1u << (p & 31));
        }
    }
    __syncthreads();
    // ---- fill: warp per row, top-most polygon first, early exit when the row is resolved ----
    for (int y = warp; y < SH; y += RS_WARPS) {
        int c0 = -1, c1 = -1, c2 = -1;
        const int xb = lane * 3;
        bool row_done = false;
        for (int wd = MASK_WORDS - 1; wd >= 0 && !row_done; --wd) {
            uint32_t m = S.rowmask[y][wd];
            while (m) {
                const int bit = 31 - __clz(m);
                m &= ~(1u << bit);
                const int p = wd * 32 + bit;
                const uchar2 sp = S.span[S.off[p] + (y - S.y0[p])];
                const int col = S.col[p];
                if (c0 < 0 && xb + 0 >= sp.x && xb + 0 < sp.y) c0 = col;
                if (c1 < 0 && xb + 1 >= sp.x && xb + 1 < sp.y) c1 = col;
                if (c2 < 0 && xb + 2 >= sp.x && xb + 2 < sp.y) c2 = col;
                if (__all_sync(0xffffffffu, (c0 | c1 | c2) >= 0)) { row_done = true; break; }
            }
        }
        uint8_t* row = S.img + (SH - 1 - y) * SW + xb;   // GL row y -> observation row 95 - y, mcr:602
        if (c0 >= 0) row[0] = (uint8_t)c0;
        if (c1 >= 0) row[1] = (uint8_t)c1;
        if (c2 >= 0) row[2] = (uint8_t)c2;
    }
    __syncthreads();
    for (int i = tid; i < SH * MASK_WORDS; i += RS_THREADS) (&S.rowmask[0][0])[i] = 0;
    if (tid == 0) { S.list_count = 0; S.pool_count = 0; }
    __syncthreads();
}

__global__ void __launch_bounds__(RS_THREADS)
render_kernel(Dims d, DevBuffers b, CarConst cc, const uint8_t* __restrict__ mask, uint8_t* __restrict__ obs,
              double* __restrict__ out_reward, uint8_t* __restrict__ out_done, int post_step,
              int backwards_flag, int use_ego_color, int max_episode_steps) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RasterSmem& S = *reinterpret_cast<RasterSmem*>(smem_raw);
    const int frame = blockIdx.x;
    const int env = frame / d.A, agent = frame % d.A;
    if (mask && !mask[env]) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = d.N, car = frame;
    const int slot = b.env_track[env];
    const int T = b.trk_T[slot], Q = b.trk_Q[slot];

    // ---- camera (mcr:540-556) was evaluated by the physics kernel for this car ---------------
    if (tid < 6) (&S.M.m00)[tid] = b.camera[(size_t)tid * N + car];
    if (tid < PAL_COUNT) S.pal32[tid] = (uint32_t)c_palette[tid][0] | ((uint32_t)c_palette[tid][1] << 8) | ((uint32_t)c_palette[tid][2] << 16);
    if (tid == 0) {
        S.list_count = 0; S.pool_count = 0;
        // score label "%04i" % reward  (mcr:665; drawn before reward -= 0.1)
        const double rw = b.reward[car];
        int val = (int)rw;
        const bool neg = rw < 0 && val != 0;
        int mag = val < 0 ? -val : val;
        char digs[12]; int nd = 0;
        do { digs[nd++] = (char)(mag % 10); mag /= 10; } while (mag > 0);
        char buf[16]; int len = 0;
        const int width = nd + (neg ? 1 : 0), pad = width < 4 ? 4 - width : 0;
        if (neg) buf[len++] = 10;
        for (int i = 0; i < pad; ++i) buf[len++] = 0;
        for (int i = nd - 1; i >= 0; --i) buf[len++] = digs[i];
        for (int i = 0; i < 4; ++i) S.glyph[i] = i < len ? buf[i] : (char)-1;
    }
    for (int i = tid; i < SH * SW / 4; i += RS_THREADS) ((uint32_t*)S.img)[i] = 0;   // glClear -> black
    for (int i = tid; i < SH * MASK_WORDS; i += RS_THREADS) (&S.rowmask[0][0])[i] = 0;
    __syncthreads();

    View V;
    V.env = env; V.agent = agent; V.A = d.A; V.N = N; V.Q = Q; V.slot = slot;
    V.body = b.body; V.wheel = b.wheel; V.stripe = b.stripe;
    V.quad = b.trk_quad + (size_t)slot * d.Qmax * 8;
    V.quad_col = b.trk_quad_col + (size_t)slot * d.Qmax;
    V.quad_tile = b.trk_quad_tile + (size_t)slot * d.Qmax;
    V.touched = b.touched + (size_t)env * d.Tmax;
    V.use_ego_color = use_ego_color;
    V.backward_flag_on = (b.backward[car] != 0) && backwards_flag;
    const Affine M = S.M;

    // ---- candidates -> ordered display list -> flush -----------------------------------------
    const int NC = 1 + N_CHECKER + Q + 12 * d.A + 9;
    int base = 0;
    while (base < NC) {
        const int i = base + tid;
        float px[MCR_MAXV], py[MCR_MAXV];
        int nv = 0, col = 0, y0 = 0, y1 = 0;
        if (i < NC) nv = gen_candidate(i, V, M, cc, px, py, col);
        bool valid = nv >= 3;
        if (valid) {
            float ymin = py[0], ymax = py[0], xmin = px[0], xmax = px[0];
            for (int k = 1; k < nv; ++k) {
                ymin = fminf(ymin, py[k]); ymax = fmaxf(ymax, py[k]);
                xmin = fminf(xmin, px[k]); xmax = fmaxf(xmax, px[k]);
            }
            if (!(ymax > 0.0f) || !(ymin < (float)SH) || !(xmax > 0.0f) || !(xmin < (float)SW)) valid = false;
            else {
                y0 = (int)ceilf(fmaxf(ymin, 0.0f) - 0.5f); if (y0 < 0) y0 = 0;
                y1 = (int)ceilf(fminf(ymax, (float)SH) - 0.5f); if (y1 > SH) y1 = SH;
                if (y1 <= y0) valid = false;
            }
        }
        const int rows = valid ? y1 - y0 : 0;
        // block-wide exclusive scan of (valid, rows)
        int cnt_inc = valid ? 1 : 0, rows_inc = rows;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int a = __shfl_up_sync(0xffffffffu, cnt_inc, o), r = __shfl_up_sync(0xffffffffu, rows_inc, o);
            if (lane >= o) { cnt_inc += a; rows_inc += r; }
        }
        if (lane == 31) { S.warp_cnt[warp] = cnt_inc; S.warp_rows[warp] = rows_inc; }
        if (tid == 0) S.first_bad = RS_THREADS;
        __syncthreads();
        int cnt_before = 0, rows_before = 0;
#pragma unroll
        for (int wq = 0; wq < RS_WARPS; ++wq)
            if (wq < warp) { cnt_before += S.warp_cnt[wq]; rows_before += S.warp_rows[wq]; }
        const int slot_rel = cnt_before + cnt_inc - (valid ? 1 : 0);
        const int row_rel = rows_before + rows_inc - rows;
        const int lc = S.list_count, pc = S.pool_count;
        const bool fits = (lc + slot_rel < LIST_CAP) && (pc + row_rel + rows <= SPAN_POOL);
        if (valid && !fits) atomicMin(&S.first_bad, tid);
        __syncthreads();
        const int first_bad = S.first_bad;
        if (valid && tid < first_bad) {
            const int sl = lc + slot_rel;
            for (int k = 0; k < nv; ++k) { S.vx[k][sl] = px[k]; S.vy[k][sl] = py[k]; }
            S.n[sl] = (uint8_t)nv; S.col[sl] = (uint8_t)col; S.y0[sl] = (uint8_t)y0; S.y1[sl] = (uint8_t)y1;
            S.off[sl] = (uint16_t)(pc + row_rel);
        }
        __syncthreads();
        if (tid == 0) {
            // totals of the accepted prefix
            int acc_cnt = 0, acc_rows = 0;
            if (first_bad == RS_THREADS) {
                for (int wq = 0; wq < RS_WARPS; ++wq) { acc_cnt += S.warp_cnt[wq]; acc_rows += S.warp_rows[wq]; }
                S.list_count = lc + acc_cnt; S.pool_count = pc + acc_rows;
            }
        }
        if (first_bad < RS_THREADS) {
            // the thread that did not fit publishes the accepted totals (its exclusive prefix)
            if (tid == first_bad) { S.list_count = lc + slot_rel; S.pool_count = pc + row_rel; }
            base += first_bad;
            flush_list(S, tid);
        } else {
            base += RS_THREADS;
            __syncthreads();
        }
    }
    flush_list(S, tid);

    // ---- score label glyphs (D3), rows 87..91, cols 2..13 ---------------------------------------
    if (tid < 60) {
        const int ch = tid / 15, ry = (tid % 15) / 3, rx = tid % 3;
        const int g = S.glyph[ch];
        if (g >= 0 && (c_font[g][ry] & (4 >> rx))) S.img[(87 + ry) * SW + (2 + 3 * ch + rx)] = PAL_WHITE;
    }
    __syncthreads();

    // ---- expand palette -> RGB and store the frame with 16-byte vector stores --------------------
    // 16 pixels = 48 bytes = 3 x uint4; byte stream r0 g0 b0 r1 g1 b1 ... from packed 0x00BBGGRR
    {
        uint4* dst = reinterpret_cast<uint4*>(obs + (size_t)frame * MCR_OBS_BYTES);
        for (int g = tid; g < SH * SW / 16; g += RS_THREADS) {
            const uint4 idx4 = reinterpret_cast<const uint4*>(S.img)[g];
            const uint32_t iw[4] = {idx4.x, idx4.y, idx4.z, idx4.w};
            uint32_t o[12];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t c0 = S.pal32[iw[q] & 0xff], c1 = S.pal32[(iw[q] >> 8) & 0xff];
                const uint32_t c2 = S.pal32[(iw[q] >> 16) & 0xff], c3 = S.pal32[iw[q] >> 24];
                o[3 * q + 0] = c0 | (c1 << 24);
                o[3 * q + 1] = (c1 >> 8) | (c2 << 16);
                o[3 * q + 2] = (c2 >> 16) | (c3 << 8);
            }
            dst[3 * g + 0] = make_uint4(o[0], o[1], o[2], o[3]);
            dst[3 * g + 1] = make_uint4(o[4], o[5], o[6], o[7]);
            dst[3 * g + 2] = make_uint4(o[8], o[9], o[10], o[11]);
        }
    }

    // ---- post-step block, mcr:433-507 (skipped for reset()'s step(None)) -----------------------
    if (!post_step) return;
    // nearest track node: argmin_i || pos - track[i].xy ||, first minimum wins (np.argmin)
    const double posx = b.body[(size_t)BF_PX * N + car], posy = b.body[(size_t)BF_PY * N + car];
    const double* node = b.trk_node + (size_t)slot * d.Tmax * 3;
    double bestd = 1.0e300; int besti = 0x7fffffff;
    for (int i = tid; i < T; i += RS_THREADS) {
        const double dx = posx - node[(size_t)i * 3 + 1], dy = posy - node[(size_t)i * 3 + 2];
        const double dd = sqrt(dx * dx + dy * dy);
        if (dd < bestd || (dd == bestd && i < besti)) { bestd = dd; besti = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double od = __shfl_down_sync(0xffffffffu, bestd, o);
        const int oi = __shfl_down_sync(0xffffffffu, besti, o);
        if (od < bestd || (od == bestd && oi < besti)) { bestd = od; besti = oi; }
    }
    if (lane == 0) { S.red_d[warp] = bestd; S.red_i[warp] = besti; }
    __syncthreads();
    if (tid == 0) {
        for (int wq = 1; wq < RS_WARPS; ++wq) {
            const double od = S.red_d[wq]; const int oi = S.red_i[wq];
            if (od < bestd || (od == bestd && oi < besti)) { bestd = od; besti = oi; }
        }
        const double PI = 3.141592653589793, PLAYFIELD = 2000 / 6.0;
        double reward = b.reward[car] - 0.1;                     // mcr:436
        double step_reward = reward - b.prev_reward[car];        // mcr:443
        const double car_angle = b.heading[car];              // mcr:449-456, evaluated in the physics kernel
        double desired = node[(size_t)besti * 3 + 0];
        if (b.env_cw[env]) desired += PI;
        desired = py_mod(desired + 2 * PI, 2 * PI);
        double diff = fabs(desired - car_angle);
        if (diff > PI) diff = fabs(diff - 2 * PI);
        uint8_t backward = 0;
        if (diff > PI / 2) { backward = 1; step_reward -= 0 * diff; }   // K_BACKWARD = 0, mcr:78
        b.backward[car] = backward;
        b.reward[car] = reward;
        b.prev_reward[car] = reward;
        // done: ANY agent finished all tiles or left the playfield (evaluated identically by
        // every agent's CTA of this env, so the plain store below is race-free in value)
        uint8_t done = 0;
        for (int c = 0; c < d.A; ++c) {
            const int oc = env * d.A + c;
            if (b.visit_count[oc] == T) done = 1;
            const double x = b.body[(size_t)BF_PX * N + oc], y = b.body[(size_t)BF_PY * N + oc];
            if (fabs(x) > PLAYFIELD || fabs(y) > PLAYFIELD) { done = 1; if (c == agent) step_reward = -100; }
        }
        if (max_episode_steps > 0 && b.steps[car] >= max_episode_steps) done |= 2;   // TimeLimit
        out_reward[car] = step_reward;
        if (agent == 0) out_done[env] = done;
    }
}

int launch_render(const Dims& d, const DevBuffers& b, const CarConst& cc, const uint8_t* mask, uint8_t* obs,
                  double* reward, uint8_t* done, int post_step, int backwards_flag,
                  int use_ego_color, int max_episode_steps, void* stream) {
    static bool configured = false;
    const size_t smem = sizeof(RasterSmem);
    if (!configured) {
        if (cudaMemcpyToSymbol(c_palette, mcr_host_palette(), sizeof(uint8_t) * PAL_COUNT * 4) != cudaSuccess) return -1;
        if (cudaFuncSetAttribute(render_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
        configured = true;
    }
    render_kernel<<<d.N, RS_THREADS, smem, (cudaStream_t)stream>>>(d, b, cc, mask, obs, reward, done, post_step,
                                                                  backwards_flag, use_ego_color, max_episode_steps);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}
