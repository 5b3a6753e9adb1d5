// raster.cu -- kernel 3: software top-down rasteriser (render_kernel) and the post-step
// reward/done block (score_kernel, runs beside the rasteriser on the side stream).
//
// Replaces, per agent view (reference: gym_multi_car_racing/multi_car_racing.py):
//   mcr:511-604  render("state_pixels") / _render_window: camera from ego pose/velocity/t,
//                glViewport(0,0,96,96) under glOrtho(0,1000,0,800), colour-buffer read-back,
//                row flip, alpha drop
//   mcr:613-632  render_road (playfield quad, 20x20 checker quads, road + border quads)
//   gym Car.draw (wheels, wheel stripes from phase, 4 hull fixtures), all cars in every view
//   mcr:634-674  render_indicators (HUD bar, 7 indicator quads, score, backward flag)
//   mcr:433-507  reward -= 0.1, step_reward, driving_backward, done / out-of-field (post_step)
// OpenGL's fixed-function fill is restated as point sampling at pixel centres with half-open
// spans, every edge evaluated from its lower to its upper endpoint (shared edges are
// watertight and bit-reproducible); painter's order is the reference's draw order.
//
// Mapping: one CTA of 288 threads (= 96 rows x 3 segments of 32 pixels) per agent-frame.
//   1. candidates: one polygon per thread in painter's order (checker squares are enumerated
//      only over the index range the camera can see), projected with the fp32 affine camera,
//      culled against the viewport and compacted IN ORDER (block scan) into a shared-memory
//      display list of canonical edges (lower endpoint, upper y, slope) -- the one IEEE
//      division per edge is paid once per polygon, not once per row;
//   2. spans: one thread per (polygon,row) slot (binary search slot -> polygon) writes the
//      half-open pixel span [x0,x1) into a shared pool and sets the polygon's bit in the row's
//      bitmask;
//   3. fill: one thread per (row, 32-pixel segment) walks the row's bitmask top-most polygon
//      first with a 32-bit "still uncovered" mask; the segment's 32 palette indices live in 8
//      registers and are updated with byte-select masks (no per-pixel loop, no image in shared
//      memory), the walk stops as soon as the segment is resolved;
//   4. each thread expands its 32 pixels to RGB and stores them as 6 x uint4 (1728 uint4 per
//      frame) -- the only HBM traffic that scales with the frame count.
#include "mcr_internal.h"
#include <cuda_runtime.h>
#include <cstring>
#include <cstddef>
#include <algorithm>
#include <cstdlib>

#define RS_THREADS 288
#define RS_WARPS (RS_THREADS / 32)
#define LIST_CAP 256
#define MASK_WORDS (LIST_CAP / 32)
#define SPAN_POOL 4096
#define N_CHECKER_AXIS 20   // range(-20, 20, 2)
#define CAR_PARTS 12        // 4 x (wheel, stripe) + 4 hull fixtures
#define MAX_CHUNKS 256      // road_poly chunks of MCR_QUAD_CHUNK quads (Qmax <= 2048)
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
    {102, 102, 0, 0},   // PAL_MUD          (0.4, 0.4, 0.0) skid traces on grass
};

// Host copy of the palette, derived from the reference's float colours with the GL rule
// u8 = floor(c * 255 + 0.5) evaluated in fp32; launch_render uploads it into c_palette.
static const float h_palette_f[PAL_COUNT][3] = {
    {0, 0, 0}, {0.4f, 0.8f, 0.4f}, {0.4f, 0.9f, 0.4f}, {0.4f, 0.4f, 0.4f}, {0.41f, 0.41f, 0.41f}, {0.42f, 0.42f, 0.42f},
    {1, 1, 1}, {1, 0, 0}, {0.3f, 0.3f, 0.3f},
    {0.8f, 0.0f, 0.0f}, {0.0f, 0.0f, 0.8f}, {0.0f, 0.8f, 0.0f}, {0.0f, 0.8f, 0.8f},
    {0.8f, 0.8f, 0.8f}, {0.0f, 0.0f, 0.0f}, {0.8f, 0.0f, 0.8f}, {0.8f, 0.8f, 0.0f},
    {0, 0, 1}, {0.2f, 0, 1}, {0, 1, 0}, {0, 0, 1}, {0.4f, 0.4f, 0.0f}};

const uint8_t (*mcr_host_palette())[4] {
    static uint8_t pal[PAL_COUNT][4];
    for (int i = 0; i < PAL_COUNT; ++i) {
        for (int c = 0; c < 3; ++c) pal[i][c] = (uint8_t)(int)floorf(h_palette_f[i][c] * 255.0f + 0.5f);
        // ITU-R 601 luma in integers (MCR_OBS_GRAY): exact, so tests can restate it with numpy
        pal[i][3] = (uint8_t)((299 * pal[i][0] + 587 * pal[i][1] + 114 * pal[i][2] + 500) / 1000);
    }
    return pal;
}

// fp16 bits of c / 255 (fp32 division, round to nearest even) per palette colour: r | g << 16, b
__constant__ uint32_t c_pal16[PAL_COUNT][2];

static uint16_t host_f2h(float f) {          // IEEE fp32 -> fp16, round to nearest even (values in [0, 1])
    uint32_t x; std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    int32_t e = (int32_t)((x >> 23) & 0xffu) - 127 + 15;
    uint32_t m = x & 0x7fffffu;
    if (e >= 31) return (uint16_t)(sign | 0x7c00u);
    if (e <= 0) {
        if (e < -10) return (uint16_t)sign;
        m |= 0x800000u;
        const int sh = 14 - e;
        uint32_t h = m >> sh;
        const uint32_t rem = m & ((1u << sh) - 1u), half = 1u << (sh - 1);
        if (rem > half || (rem == half && (h & 1u))) ++h;
        return (uint16_t)(sign | h);
    }
    uint32_t h = ((uint32_t)e << 10) | (m >> 13);
    const uint32_t rem = m & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) ++h;
    return (uint16_t)(sign | h);
}

// 3x5 digit font for the score label (documented deviation D3: the reference uses the
// platform font through pyglet).  Row-major, msb = left column.
__constant__ uint8_t c_font[11][5] = {
    {7, 5, 5, 5, 7}, {2, 6, 2, 2, 7}, {7, 1, 7, 4, 7}, {7, 1, 7, 1, 7}, {5, 5, 7, 1, 1}, {7, 4, 7, 1, 7},
    {7, 4, 7, 5, 7}, {7, 1, 1, 1, 1}, {7, 5, 7, 5, 7}, {7, 5, 7, 1, 7}, {0, 0, 7, 0, 0}};

struct Affine { float m00, m01, m02, m10, m11, m12; };

// Diagnostics build (python -m multi_car_racing_b200.build --phase-clocks -> libmcr_clk.so): thread 0 of every
// render CTA adds the SM clock cycles it spent in each phase to g_phase_clk (scripts/render_phases.py).
#ifdef MCR_PHASE_CLOCKS
__device__ unsigned long long g_phase_clk[16];
#define PHASE_T0() long long ph_t_ = clock64()
#define PHASE(k) do { if (threadIdx.x == 0) { const long long n_ = clock64(); atomicAdd(&g_phase_clk[k], (unsigned long long)(n_ - ph_t_)); ph_t_ = n_; } } while (0)
extern "C" int mcr_debug_phase_clocks(unsigned long long* out16, int reset) {
    if (out16 && cudaMemcpyFromSymbol(out16, g_phase_clk, sizeof(g_phase_clk)) != cudaSuccess) return -1;
    if (reset) { unsigned long long z[16] = {}; if (cudaMemcpyToSymbol(g_phase_clk, z, sizeof(z)) != cudaSuccess) return -1; }
    return 0;
}
#else
#define PHASE_T0() do {} while (0)
#define PHASE(k) do {} while (0)
#endif

// One display-list entry = 4 canonical edges.  The hull octagon takes two consecutive
// entries (ne = 8 on the first, a row-less continuation after it).
struct __align__(16) RasterSmem {
    // canonical edge = (x at the lower endpoint, y of the lower endpoint (+inf: unused / horizontal edge), y of the
    // upper endpoint, slope (bx - ax) / (by - ay)): one LDS.128 per edge.  Entries LIST_CAP + c hold edges 4..7
    // of car c's hull octagon (the only polygon with more than four edges).
    float4 edge[4][LIST_CAP + MCR_MAX_AGENTS];
    // fill_kernel stages the finished frame in `edge` + `stage_tail` (27 648 bytes, contiguous) for its bulk store: the
    // edges are dead once the last list's spans are computed, so a warp can stage its rows as soon as its own fill is
    // done, while other warps still read the spans and masks
    unsigned char stage_tail[MCR_OBS_BYTES - sizeof(float4) * 4 * (LIST_CAP + MCR_MAX_AGENTS)];
    int base[LIST_CAP];           // span pool offset of viewport row 0 of this polygon: first slot - y0
    uint8_t ne[LIST_CAP], col[LIST_CAP];   // ne = 4, or 8 + c for car c's octagon
    // slot -> polygon without a search: polygon p's first slot sets a bit in startbits (unless it is word aligned),
    // slot32_owner[k] = owner of slot 32 k; owner(32 k + l) = slot32_owner[k] + popc(startbits[k] & bits 0..l)
    uint32_t startbits[SPAN_POOL / 32];
    uint8_t slot32_owner[SPAN_POOL / 32];
    uchar2 span[SPAN_POOL];
    uint32_t rowmask[MASK_WORDS][SH * 3];   // per (32-pixel segment, row) = fill thread: polygons whose span touches it (word-major: conflict-free)
    Affine M;
    int first_bad;
    int ck_j0x, ck_nx, ck_j0y, ck_ny;   // visible checker index range
    int grass_full;                     // the playfield quad covers every pixel centre of this tile
    uint32_t chunk_ballot[RS_WARPS];
    int n_vis_chunks;
    uint8_t vis_chunk[MAX_CHUNKS];    // ids of the road_poly chunks whose bounding circle touches the viewport, ascending
    int warp_cnt[2][RS_WARPS], warp_rows[2][RS_WARPS];
    int bc_cnt, bc_rows;
    float red_f[RS_WARPS]; int cand[32]; int n_cand;
    signed char glyph[4];
    uint32_t pal32[32];      // r | g << 8 | b << 16
    uint32_t palY[32];       // luma (MCR_OBS_GRAY)
    uint2 pal16[32];         // fp16 bits of r / 255, g / 255 (x) and b / 255 (y): MCR_OBS_RGB_CHW_F16
    uint32_t prmt_sel[256];  // coverage byte -> two PRMT selectors (low half: pixels 0-3, high half: pixels 4-7):
                             // byte i comes from operand b (4 + i) if its bit is set, else from a (i)
};

__device__ __forceinline__ double py_mod(double a, double m) {
    double r = fmod(a, m);
    if (r != 0 && ((r < 0) != (m < 0))) r += m;
    return r;
}

struct View {
    int env, agent, A, N, Q;
    int n_checker, ck_j0x, ck_j0y, ck_ny, grass_full;
    int c_road, c_cars, c_hud;           // first candidate index of each class (warp aligned)
    int n_road;                          // 8 * visible chunks
    const uint8_t* vis_chunk;
    const float* body; const double* wheel; const float* stripe;
    const float* quad; const uint8_t* quad_col; const int16_t* quad_tile; const uint8_t* touched;
    int use_ego_color, backward_flag_on;
    int car_slots;                       // candidate slots per car: 12 parts (+ PRT_MAX * (PRT_PTS - 1) skid-trace segments before them)
    const float* prt_pts; const int32_t* prt_meta; const int32_t* prt_hdr;
    float hud_sx, hud_sy;                // window (1000 x 800) -> viewport pixels: (float)(vw / 1000.0), (float)(vh / 800.0)
};

__device__ __forceinline__ void xf_pt(const Affine& M, float x, float y, float& ox, float& oy) {
    ox = (M.m00 * x + M.m01 * y) + M.m02;
    oy = (M.m10 * x + M.m11 * y) + M.m12;
}

// Candidate polygon i in painter's order.  Returns the vertex count (0 = nothing to draw).
__device__ int gen_candidate(int i, const View& V, const Affine& M, const CarConst& cc, float (&px)[MCR_MAXV],
                             float (&py)[MCR_MAXV], int& col, int& aux) {
    const double PLAYFIELD = 2000 / 6.0;
    // candidate classes start on warp boundaries so that a warp executes one kind of polygon
    if (i < V.c_road) {
        if (i == 0) {                               // playfield, mcr:615-619
            if (V.grass_full) return 0;             // covers the whole tile: the pixels start out as grass instead
            const float pf = (float)PLAYFIELD;
            xf_pt(M, -pf, +pf, px[0], py[0]); xf_pt(M, +pf, +pf, px[1], py[1]);
            xf_pt(M, +pf, -pf, px[2], py[2]); xf_pt(M, -pf, -pf, px[3], py[3]);
            col = PAL_GRASS; return 4;
        }
        i -= 1;
        if (i >= V.n_checker) return 0;
        {                                           // checker quads that can be in view, mcr:620-627
            const double k = PLAYFIELD / 20.0;
            const int x = -20 + 2 * (V.ck_j0x + i / V.ck_ny), y = -20 + 2 * (V.ck_j0y + i % V.ck_ny);
            xf_pt(M, (float)(k * x + k), (float)(k * y + 0), px[0], py[0]);
            xf_pt(M, (float)(k * x + 0), (float)(k * y + 0), px[1], py[1]);
            xf_pt(M, (float)(k * x + 0), (float)(k * y + k), px[2], py[2]);
            xf_pt(M, (float)(k * x + k), (float)(k * y + k), px[3], py[3]);
            col = PAL_GRASS_LIGHT; return 4;
        }
    }
    if (i < V.c_cars) {                             // road_poly (visible chunks only), mcr:628-631
        i -= V.c_road;
        if (i >= V.n_road) return 0;
        i = (int)V.vis_chunk[i >> 3] * MCR_QUAD_CHUNK + (i & (MCR_QUAD_CHUNK - 1));
        if (i >= V.Q) return 0;
        const float4 a = *(const float4*)(V.quad + (size_t)i * 8);
        const float4 b = *(const float4*)(V.quad + (size_t)i * 8 + 4);
        xf_pt(M, a.x, a.y, px[0], py[0]); xf_pt(M, a.z, a.w, px[1], py[1]);
        xf_pt(M, b.x, b.y, px[2], py[2]); xf_pt(M, b.z, b.w, px[3], py[3]);
        const int tl = V.quad_tile[i];
        col = (tl >= 0 && V.touched[tl]) ? PAL_ROAD0 : V.quad_col[i];   // tile.color reset, mcr:102-104
        return 4;
    }
    if (i < V.c_hud) {                              // Car.draw for every car, mcr:559-564
        i -= V.c_cars;
        if (i >= V.car_slots * V.A) return 0;
        const int c = i / V.car_slots, car = V.env * V.A + c, N = V.N;
        int part = i - c * V.car_slots;
        if (V.car_slots > CAR_PARTS) {
            // Car.draw(viewer, draw_particles=True): the skid traces come first, one candidate per polyline segment.
            // D6: a glLineWidth(5) line-strip segment is the parallelogram spanning +-2.5 viewport pixels along the
            // minor axis (y for |dx| >= |dy|, else x), filled with the polygon rule.
            const int nseg = PRT_MAX * (PRT_PTS - 1);
            if (part < nseg) {
                const int pi = part / (PRT_PTS - 1), seg = part - pi * (PRT_PTS - 1);
                const int head = V.prt_hdr[car], count = V.prt_hdr[(size_t)N + car];
                if (pi >= count) return 0;
                int slot = head + pi; if (slot >= PRT_MAX) slot -= PRT_MAX;
                const int meta = V.prt_meta[(size_t)slot * N + car];
                if (seg + 1 >= (meta & 0xff)) return 0;
                const float* pt = V.prt_pts + ((size_t)(slot * PRT_PTS + seg) * 2) * N + car;
                float x0, y0, x1, y1;
                xf_pt(M, pt[0], pt[N], x0, y0);
                xf_pt(M, pt[(size_t)2 * N], pt[(size_t)3 * N], x1, y1);
                const float hw = 2.5f;
                if (fabsf(x1 - x0) >= fabsf(y1 - y0)) {
                    px[0] = x0; py[0] = y0 - hw; px[1] = x1; py[1] = y1 - hw; px[2] = x1; py[2] = y1 + hw; px[3] = x0; py[3] = y0 + hw;
                } else {
                    px[0] = x0 - hw; py[0] = y0; px[1] = x1 - hw; py[1] = y1; px[2] = x1 + hw; py[2] = y1; px[3] = x0 + hw; py[3] = y0;
                }
                col = (meta & 256) ? PAL_MUD : PAL_BLACK;
                return 4;
            }
            part -= nseg;
        }
        if (part < 8) {
            const int wl = part >> 1;
            const float* bp = V.body + (size_t)((1 + wl) * BODY_FIELDS) * N + car;
            const float bx = bp[(size_t)BF_PX * N], by = bp[(size_t)BF_PY * N];
            const float qs = bp[(size_t)BF_QS * N], qc = bp[(size_t)BF_QC * N];
            float lx[4], ly[4];
            if ((part & 1) == 0) {
#pragma unroll
                for (int k = 0; k < 4; ++k) { lx[k] = cc.wheel_poly.x[k]; ly[k] = cc.wheel_poly.y[k]; }
                col = PAL_BLACK;
            } else {
                // stripe quad y-extents were evaluated once per wheel by the physics kernel
                // (Car.draw: a1 = phase, a2 = phase + 1.2 ...); NaN = "not drawn this frame"
                const float sy1 = V.stripe[(size_t)(wl * 2 + 0) * N + car], sy2 = V.stripe[(size_t)(wl * 2 + 1) * N + car];
                if (sy1 != sy1) return 0;
                const float hw = fabsf(cc.wheel_poly.x[0]);            // (float)(WHEEL_W*SIZE)
                lx[0] = -hw; ly[0] = sy1; lx[1] = +hw; ly[1] = sy1;
                lx[2] = +hw; ly[2] = sy2; lx[3] = -hw; ly[3] = sy2;
                col = PAL_WHEEL_WHITE;
            }
#pragma unroll
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
#pragma unroll
            for (int k = 0; k < MCR_MAXV; ++k) {
                if (k < P.n) {
                    const float wx = (qc * P.x[k] - qs * P.y[k]) + bx, wy = (qs * P.x[k] + qc * P.y[k]) + by;
                    xf_pt(M, wx, wy, px[k], py[k]);
                }
            }
            if (V.use_ego_color) col = (c == V.agent) ? PAL_CAR0 + 0 : PAL_CAR0 + 1;   // mcr:560-563
            else col = PAL_CAR0 + (c % 8);                                              // mcr:402
            aux = c;
            return P.n;
        }
    }
    i -= V.c_hud;
    if (i >= 9) return 0;
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
            wx[0] = Wd - 100; wy[0] = 30; wx[1] = Wd - 75; wy[1] = 70; wx[2] = Wd - 50; wy[2] = 30;
            wx[3] = 0; wy[3] = 0; n = 3;
            col = PAL_FLAG_BLUE;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            px[k] = (float)wx[k] * V.hud_sx;
            py[k] = (float)wy[k] * V.hud_sy;
        }
        return n;
    }
}

// Evaluate one canonical edge on the row through yc; identical arithmetic to the CPU
// restatement: x = ax + (yc - ay) * ((bx - ax) / (by - ay)), edge taken lower -> upper.
__device__ __forceinline__ void edge_row(const RasterSmem& S, int e, int p, float yc, float& xl, float& xr) {
    const float4 ed = S.edge[e][p];                 // (ax, ay, by, slope)
    if (yc >= ed.y && yc < ed.z) {
        const float x = ed.x + (yc - ed.y) * ed.w;
        xl = fminf(xl, x); xr = fmaxf(xr, x);
    }
}

// (ox, oy) = viewport pixel of the tile's lower-left corner, vw = viewport width: every edge is
// evaluated in VIEWPORT coordinates (identical arithmetic for any tiling), only the resulting
// integer span is moved into the tile.  The 96 x 96 state frame is the single tile (0, 0).
template <bool VP>
__device__ __forceinline__ void flush_list(RasterSmem& S, int tid, uint32_t (&pix)[8], int n, int nslots, bool last,
                                           int ox, int oy, int vw) {
    PHASE_T0();
    // only the mask words this list can set are read by the fill below
    for (int i = tid; i < ((n + 31) >> 5) * (SH * 3); i += RS_THREADS) (&S.rowmask[0][0])[i] = 0;
    __syncthreads();
    if (nslots < 0) nslots = S.bc_rows;                // fill_kernel: written by the thread that loaded the chunk's last entry
    PHASE(5);
    // ---- spans: one thread per (polygon,row) slot ---------------------------------------------
    for (int s = tid; s < nslots; s += RS_THREADS) {
        // RS_THREADS is a multiple of 32: the lanes of a warp share the word s >> 5
        const int p = (int)S.slot32_owner[s >> 5] + __popc(S.startbits[s >> 5] & (0xffffffffu >> (31 - (s & 31))));
        const int row = s - S.base[p];
        const float yc = (float)(VP ? row + oy : row) + 0.5f;
        float xl = 3.402823466e+38f, xr = -3.402823466e+38f;
#pragma unroll
        for (int e = 0; e < 4; ++e) edge_row(S, e, p, yc, xl, xr);
        const int ne = S.ne[p];
        if (ne >= 8) {
#pragma unroll
            for (int e = 0; e < 4; ++e) edge_row(S, e, LIST_CAP + (ne - 8), yc, xl, xr);
        }
        int x0 = 0, x1 = 0;
        if (xl < xr) {
            const int VW = VP ? vw : SW;
            xl = fminf(fmaxf(xl, -1.0f), (float)VW + 1.0f);
            xr = fminf(fmaxf(xr, -1.0f), (float)VW + 1.0f);
            x0 = (int)ceilf(xl - 0.5f); if (x0 < 0) x0 = 0;
            x1 = (int)ceilf(xr - 0.5f); if (x1 > VW) x1 = VW;
            if (VP) {   // viewport span -> tile columns
                x0 -= ox; x1 -= ox;
                x0 = x0 < 0 ? 0 : (x0 > SW ? SW : x0);
                x1 = x1 < 0 ? 0 : (x1 > SW ? SW : x1);
            }
        }
        S.span[s] = make_uchar2((unsigned char)x0, (unsigned char)x1);
        if (x0 < x1) {
            const uint32_t bit = 1u << (p & 31);
            const int g0 = x0 >> 5, g1 = (x1 - 1) >> 5;
            for (int g = g0; g <= g1; ++g) atomicOr(&S.rowmask[p >> 5][g * SH + row], bit);
        }
    }
    __syncthreads();
    PHASE(6);
    // ---- fill: one thread per (row, 32-pixel segment), top-most polygon first ------------------
    // pix[k] holds the palette indices of pixels 4k..4k+3 of this thread's segment.  Polygons of a
    // later flush are later in painter's order, so they overwrite what earlier flushes left.
    {
        // thread = (segment, row) with the row fastest: a warp walks 32 consecutive rows of one segment
        // column, whose polygon lists are nearly the same (little divergence in the loop below)
        const int seg = tid / SH, y = tid - SH * seg;
        const int xs = 32 * seg;
        uint32_t uncovered = 0xffffffffu;
        for (int wd = (n - 1) >> 5; wd >= 0 && uncovered; --wd) {
            uint32_t m = S.rowmask[wd][tid];               // tid == seg * SH + y
            while (m && uncovered) {
                const int bit = 31 - __clz(m);
                m &= ~(1u << bit);
                const int p = wd * 32 + bit;
                const uchar2 sp = S.span[S.base[p] + y];
                int lo = (int)sp.x - xs, hi = (int)sp.y - xs;
                lo = lo < 0 ? 0 : lo; hi = hi > 32 ? 32 : hi;
                if (hi > lo) {
                    const uint32_t cover = (uint32_t)((1ull << hi) - (1ull << lo));
                    const uint32_t fresh = cover & uncovered;
                    uncovered &= ~cover;
                    const uint32_t c4 = (uint32_t)S.col[p] * 0x01010101u;
                    if (fresh == 0xffffffffu) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) pix[k] = c4;
                    } else if (fresh) {
#pragma unroll
                        for (int k2 = 0; k2 < 4; ++k2) {
                            // bit i of a coverage nibble selects byte i of c4 over byte i of pix[k]: one PRMT per
                            // four pixels, the selectors of eight pixels from one table load
                            const uint32_t sel = S.prmt_sel[(fresh >> (8 * k2)) & 0xFFu];
                            pix[2 * k2] = __byte_perm(pix[2 * k2], c4, sel);
                            pix[2 * k2 + 1] = __byte_perm(pix[2 * k2 + 1], c4, sel >> 16);
                        }
                    }
                }
            }
        }
    }
    PHASE(7);
    if (last) return;                                  // nothing reads the list or the masks again
    __syncthreads();
    if (tid < SPAN_POOL / 32) S.startbits[tid] = 0;
    __syncthreads();
}

// coverage byte -> two PRMT selectors (low half: pixels 0-3, high half: pixels 4-7): byte i comes from operand b
// (4 + i) if its bit is set, else from a (i)
__device__ __forceinline__ uint32_t prmt_selector(int byte) {
    const uint32_t lo = byte & 15u, hi = (uint32_t)byte >> 4;
    const uint32_t slo = 0x3210u | ((lo & 1u) << 2) | ((lo & 2u) << 5) | ((lo & 4u) << 8) | ((lo & 8u) << 11);
    const uint32_t shi = 0x3210u | ((hi & 1u) << 2) | ((hi & 2u) << 5) | ((hi & 4u) << 8) | ((hi & 8u) << 11);
    return slo | (shi << 16);
}

// Character i (0..3) of the score label "%04i" % reward (mcr:665; drawn before reward -= 0.1) as a glyph index
// (0-9, 10 = '-', -1 = none): straight-line, one thread per character.
__device__ __forceinline__ signed char score_glyph(double rw, int i) {
    const int val = (int)rw;
    const bool neg = rw < 0 && val != 0;
    const unsigned mag = (unsigned)(val < 0 ? -(long long)val : (long long)val);
    int nd = 1;
    for (unsigned p = 10u; nd < 10 && mag >= p; p *= 10u) ++nd;      // decimal digits of |val| (p overflows only after nd = 10)
    // "%04i": zero padded to a width of 4 (the sign counts), longer numbers are cut after four characters
    int e;                                                           // power of ten of the digit shown at position i
    if (!neg) e = nd <= 4 ? 3 - i : nd - 1 - i;
    else { if (i == 0) return 10; e = nd <= 3 ? 3 - i : nd - i; }
    unsigned p10 = 1u;
    for (int k = 0; k < e; ++k) p10 *= 10u;
    return (signed char)((mag / p10) % 10u);
}

// Front of a frame (fused kernel) (NT threads; SM has the fields used below):
// camera -> S.M, the score label's glyphs, the visible checker index range, "the playfield quad covers the tile",
// and the ascending list of road_poly chunks (8 consecutive quads, bounding circle) that can touch the tile.
// Ends with a __syncthreads().  rw = env.reward as the label shows it.
template <int NT, class SM>
__device__ __forceinline__ void frame_front(SM& S, const Dims& d, const DevBuffers& b, const float* __restrict__ camera, double rw,
                                            int car, int slot, int tid, float tx0, float ty0, float tx1, float ty1) {
    const int N = d.N, warp = tid >> 5, lane = tid & 31;
    // ---- camera (mcr:540-556) was evaluated by the physics kernel for this car ---------------
    if (tid < 6) (&S.M.m00)[tid] = camera[(size_t)tid * N + car];
    if (tid >= 64 && tid < 68) S.glyph[tid - 64] = score_glyph(rw, tid - 64);
    if (tid == 96) {
        // Checker squares the camera can see: invert the affine for the four viewport corners
        // (+ a two-unit margin, far above fp32 error) and keep the index ranges that overlap.
        const float m00 = camera[(size_t)0 * N + car], m01 = camera[(size_t)1 * N + car], m02 = camera[(size_t)2 * N + car];
        const float m10 = camera[(size_t)3 * N + car], m11 = camera[(size_t)4 * N + car], m12 = camera[(size_t)5 * N + car];
        const float det = m00 * m11 - m01 * m10;
        int j0x = 0, j1x = N_CHECKER_AXIS - 1, j0y = 0, j1y = N_CHECKER_AXIS - 1, grass_full = 0;
        if (fabsf(det) > 1e-12f) {
            const float inv = 1.0f / det;
            float wxmin = 3.0e38f, wxmax = -3.0e38f, wymin = 3.0e38f, wymax = -3.0e38f;
#pragma unroll
            for (int cnr = 0; cnr < 4; ++cnr) {
                const float u = ((cnr & 1) ? tx1 : tx0) - m02, v = ((cnr & 2) ? ty1 : ty0) - m12;
                const float wx = (m11 * u - m01 * v) * inv, wy = (-m10 * u + m00 * v) * inv;
                wxmin = fminf(wxmin, wx); wxmax = fmaxf(wxmax, wx); wymin = fminf(wymin, wy); wymax = fmaxf(wymax, wy);
            }
            // the tile lies inside the playfield square by >= 4 world units (>= 0.2 px at the smallest zoom, far above
            // the fp32 error of the fill): every pixel centre of the tile is inside the playfield quad
            const float pfm = (float)(2000 / 6.0) - 4.0f;
            grass_full = (wxmin > -pfm && wxmax < pfm && wymin > -pfm && wymax < pfm) ? 1 : 0;
            const float ik = 20.0f / (float)(2000 / 6.0), margin = 2.0f;
            // square j spans k*(2j-20) .. k*(2j-19) on its axis
            const float a0 = floorf(((wxmin - margin) * ik + 19.0f) * 0.5f - 1e-3f), a1 = ceilf(((wxmax + margin) * ik + 20.0f) * 0.5f + 1e-3f);
            const float c0 = floorf(((wymin - margin) * ik + 19.0f) * 0.5f - 1e-3f), c1 = ceilf(((wymax + margin) * ik + 20.0f) * 0.5f + 1e-3f);
            if (a0 == a0 && a1 == a1 && c0 == c0 && c1 == c1) {
                j0x = (int)fmaxf(0.0f, fminf(a0, 20.0f)); j1x = (int)fminf((float)(N_CHECKER_AXIS - 1), fmaxf(a1, -1.0f));
                j0y = (int)fmaxf(0.0f, fminf(c0, 20.0f)); j1y = (int)fminf((float)(N_CHECKER_AXIS - 1), fmaxf(c1, -1.0f));
            }
        }
        S.ck_j0x = j0x; S.ck_nx = j1x >= j0x ? j1x - j0x + 1 : 0;
        S.ck_j0y = j0y; S.ck_ny = j1y >= j0y ? j1y - j0y + 1 : 0;
        S.grass_full = grass_full;
    }
    // ---- road_poly chunk culling: bounding circle of every 8 consecutive quads vs the viewport ------
    // (chunks past the track's last quad have radius 0 in the pool: no dependency on trk_Q here)
    const int nchunks = d.Qmax / MCR_QUAD_CHUNK;
    bool vis = false;
    if (tid < nchunks) {
        const float4 cc4 = *(const float4*)(b.trk_chunk + ((size_t)slot * nchunks + tid) * 4);
        const float m00 = camera[(size_t)0 * N + car], m01 = camera[(size_t)1 * N + car], m02 = camera[(size_t)2 * N + car];
        const float m10 = camera[(size_t)3 * N + car], m11 = camera[(size_t)4 * N + car], m12 = camera[(size_t)5 * N + car];
        const float cxp = (m00 * cc4.x + m01 * cc4.y) + m02, cyp = (m10 * cc4.x + m11 * cc4.y) + m12;
        // |M v| <= ||M||_F |v|: conservative pixel radius, plus a 2-pixel margin
        const float rp = cc4.z * sqrtf(m00 * m00 + m01 * m01 + m10 * m10 + m11 * m11) * 1.001f + 2.0f;
        vis = (cxp + rp >= tx0) && (cxp - rp <= tx1) && (cyp + rp >= ty0) && (cyp - rp <= ty1);
        if (!(rp == rp) || !(cxp == cxp) || !(cyp == cyp)) vis = true;
        if (!(cc4.z > 0.0f)) vis = false;
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, vis);
    if (lane == 0) S.chunk_ballot[warp] = bal;
    __syncthreads();
    if (vis) {
        int pos = __popc(bal & ((1u << lane) - 1u));
        for (int wq = 0; wq < warp; ++wq) pos += __popc(S.chunk_ballot[wq]);
        S.vis_chunk[pos] = (uint8_t)tid;
    }
    if (tid == 0) {
        int n = 0;
        for (int wq = 0; wq < NT / 32; ++wq) n += __popc(S.chunk_ballot[wq]);
        S.n_vis_chunks = n;
    }
    __syncthreads();
}

// Rows [y0, y1) of this tile a candidate polygon can touch (false: it fills nothing here).  world: the polygon is
// drawn before the HUD bar, which hides viewport rows [0, hud_rows).
template <bool VP>
__device__ __forceinline__ bool cand_rows(int nv, const float (&px)[MCR_MAXV], const float (&py)[MCR_MAXV], bool world, int hud_rows,
                                          float tx0, float ty0, float tx1, float ty1, int oy, int VH, int& y0, int& y1) {
    if (nv < 3) return false;
    float ymin = py[0], ymax = py[0], xmin = px[0], xmax = px[0];
#pragma unroll
    for (int k = 1; k < MCR_MAXV; ++k) {
        if (k < nv) {
            ymin = fminf(ymin, py[k]); ymax = fmaxf(ymax, py[k]);
            xmin = fminf(xmin, px[k]); xmax = fmaxf(xmax, px[k]);
        }
    }
    if (!(ymax > ty0) || !(ymin < ty1) || !(xmax > tx0) || !(xmin < tx1)) return false;
    // rows of the VIEWPORT the polygon covers (as the full-frame fill takes them), then this tile's share
    y0 = (int)ceilf(fmaxf(ymin, 0.0f) - 0.5f); if (y0 < 0) y0 = 0;
    y1 = (int)ceilf(fminf(ymax, (float)VH) - 0.5f); if (y1 > VH) y1 = VH;
    if (world) y0 = max(y0, hud_rows);
    if (VP) { y0 = max(y0, oy) - oy; y1 = min(y1, min(oy + SH, VH)) - oy; }
    if (y1 <= y0) return false;
    // same test along x: a polygon whose bounding box holds no pixel-centre column fills nothing
    // (every span lies inside [xmin, xmax]); most road quads of a zoomed-out frame go here
    // (such a box is < 1 px wide, so |x| <= vw + 1 and the interpolated crossings stay within a few
    // ulp(vw) ~ 1e-4 of [xmin, xmax]: the 1e-3 margin keeps the cull conservative, hence exact)
    const float cx0 = ceilf((fmaxf(xmin, tx0) - 1e-3f) - 0.5f), cx1 = ceilf((fminf(xmax, tx1) + 1e-3f) - 0.5f);
    return cx1 > cx0;
}

// Canonical edge k of a polygon: lower endpoint first, slope hoisted out of the row loop (one IEEE division per edge).
__device__ __forceinline__ float4 canon_edge(int k, int nv, const float (&px)[MCR_MAXV], const float (&py)[MCR_MAXV]) {
    const int k2 = (k + 1 < nv) ? k + 1 : 0;
    float ax = px[k], ay = py[k], bx = px[k2], by = py[k2];
    if (k >= nv || ay == by) return make_float4(0.0f, 3.402823466e+38f, 3.402823466e+38f, 0.0f);
    if (ay > by) { float t = ax; ax = bx; bx = t; t = ay; ay = by; by = t; }
    return make_float4(ax, ay, by, (bx - ax) / (by - ay));
}

// Score label glyphs, palette expansion and the store of this thread's 32 pixels (6 x uint4 for the
// reference's RGB HWC layout).  Shared by the fused kernel (viewport modes) and fill_kernel (step path).
template <bool VP>
__device__ __forceinline__ void finish_frame(RasterSmem& S, int tid, uint32_t (&pix)[8], uint8_t* __restrict__ obs, int frame,
                                             int obs_format, int ox, int oy, int VW, int VH, int stack_k = 1, int stack_step = 0) {
    const int my_seg = tid / SH, my_y = tid - SH * my_seg;
    // ---- score label glyphs (D3): observation rows 87..91 (= GL rows 8..4), cols 2..13 -------------
    if (VP) {
        // the same glyph cells, scaled: viewport pixel -> state cell by nearest neighbour (D3)
        const int Y = VH - 1 - (oy + my_y);                       // image row from the top
        const int sy = (int)floor(((double)Y + 0.5) * 96.0 / (double)VH) - 87;
        if (sy >= 0 && sy < 5 && oy + my_y < VH) {
#pragma unroll 1
            for (int i = 0; i < 32; ++i) {
                const int X = ox + 32 * my_seg + i;
                const int sx = (int)floor(((double)X + 0.5) * 96.0 / (double)VW) - 2;
                if (sx < 0 || sx >= 12 || X >= VW) continue;
                const int g = S.glyph[sx / 3];
                const int bits = g >= 0 ? c_font[g][sy] : 0;
                if (bits & (4 >> (sx % 3))) pix[i >> 2] = (pix[i >> 2] & ~(0xFFu << (8 * (i & 3)))) | ((uint32_t)PAL_WHITE << (8 * (i & 3)));
            }
        }
    } else if (my_seg == 0 && my_y >= 4 && my_y <= 8) {
        const int ry = (SH - 1 - my_y) - 87;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            const int g = S.glyph[ch];
            const int bits = g >= 0 ? c_font[g][ry] : 0;
#pragma unroll
            for (int rx = 0; rx < 3; ++rx) {
                const int x = 2 + 3 * ch + rx;
                if (bits & (4 >> rx)) pix[x >> 2] = (pix[x >> 2] & ~(0xFFu << (8 * (x & 3)))) | ((uint32_t)PAL_WHITE << (8 * (x & 3)));
            }
        }
    }

    // ---- expand palette -> RGB and store this thread's 32 pixels as 6 x uint4 ------------------------
    // 4 pixels = 12 bytes = 3 words; byte stream r0 g0 b0 r1 g1 b1 ... from packed 0x00BBGGRR
    const int out_row = SH - 1 - my_y;
    if (VP) {
        // RGB rows of VW pixels; the row stride VW * 3 is only 4-byte aligned in general -> word stores
        const int gy = oy + my_y, gx = ox + 32 * my_seg;
        if (gy < VH && gx < VW) {
            const int npx = min(32, VW - gx);                     // multiple of 4 (vw % 4 == 0 is checked by the API)
            uint32_t* dst = reinterpret_cast<uint32_t*>(obs + ((size_t)frame * VH + (size_t)(VH - 1 - gy)) * ((size_t)VW * 3) + (size_t)gx * 3);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t c0 = S.pal32[pix[k] & 0xff], c1 = S.pal32[(pix[k] >> 8) & 0xff];
                const uint32_t c2 = S.pal32[(pix[k] >> 16) & 0xff], c3 = S.pal32[pix[k] >> 24];
                const uint32_t w0 = c0 | (c1 << 24), w1 = (c1 >> 8) | (c2 << 16), w2 = (c2 >> 16) | (c3 << 8);
                if (4 * k < npx) { dst[3 * k] = w0; dst[3 * k + 1] = w1; dst[3 * k + 2] = w2; }   // 4 pixels = 12 bytes
            }
        }
    } else if (obs_format == MCR_OBS_RGB_HWC) {
        uint4* dst = reinterpret_cast<uint4*>(obs + (size_t)frame * MCR_OBS_BYTES + (size_t)out_row * (SW * 3) + my_seg * 96);
        uint32_t o[24];
        const uint32_t p0 = pix[0];
        const bool uniform = __byte_perm(p0, 0, 0x0000) == p0 &&
                             ((p0 ^ pix[1]) | (p0 ^ pix[2]) | (p0 ^ pix[3]) | (p0 ^ pix[4]) | (p0 ^ pix[5]) | (p0 ^ pix[6]) | (p0 ^ pix[7])) == 0u;
        if (uniform) {
            // one colour across the 32 pixels (grass, road, HUD bar: about half of all segments): the 12-byte
            // r g b pattern repeats, one palette lookup instead of 32
            const uint32_t c = S.pal32[p0 & 0xff];
            const uint32_t w0 = __byte_perm(c, c, 0x4210), w1 = __byte_perm(c, c, 0x5421), w2 = __byte_perm(c, c, 0x6542);
            const uint4 q0 = make_uint4(w0, w1, w2, w0), q1 = make_uint4(w1, w2, w0, w1), q2 = make_uint4(w2, w0, w1, w2);
            dst[0] = q0; dst[1] = q1; dst[2] = q2; dst[3] = q0; dst[4] = q1; dst[5] = q2;
            return;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t c0 = S.pal32[pix[k] & 0xff], c1 = S.pal32[(pix[k] >> 8) & 0xff];
            const uint32_t c2 = S.pal32[(pix[k] >> 16) & 0xff], c3 = S.pal32[pix[k] >> 24];
            // (byte permutes: r0 g0 b0 r1 | g1 b1 r2 g2 | b2 r3 g3 b3 -- one instruction per word instead of shift, shift, or)
            o[3 * k + 0] = __byte_perm(c0, c1, 0x4210);
            o[3 * k + 1] = __byte_perm(c1, c2, 0x5421);
            o[3 * k + 2] = __byte_perm(c2, c3, 0x6542);
        }
#pragma unroll
        for (int v = 0; v < 6; ++v) dst[v] = make_uint4(o[4 * v], o[4 * v + 1], o[4 * v + 2], o[4 * v + 3]);
    } else if (obs_format == MCR_OBS_RGB_CHW_F16) {
        // planar fp16, value / 255: 32 pixels = 64 bytes = 4 x uint4 per plane
        uint2 v[32];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            v[4 * k + 0] = S.pal16[pix[k] & 0xff]; v[4 * k + 1] = S.pal16[(pix[k] >> 8) & 0xff];
            v[4 * k + 2] = S.pal16[(pix[k] >> 16) & 0xff]; v[4 * k + 3] = S.pal16[pix[k] >> 24];
        }
        uint16_t* plane0 = reinterpret_cast<uint16_t*>(obs) + (size_t)frame * (3 * SW * SH) + (size_t)out_row * SW + my_seg * 32;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            uint32_t o[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const uint32_t a = c == 2 ? v[2 * k].y : v[2 * k].x, bq = c == 2 ? v[2 * k + 1].y : v[2 * k + 1].x;
                o[k] = c == 1 ? __byte_perm(a, bq, 0x7632) : __byte_perm(a, bq, 0x5410);
            }
            uint4* dst = reinterpret_cast<uint4*>(plane0 + (size_t)c * (SW * SH));
#pragma unroll
            for (int q = 0; q < 4; ++q) dst[q] = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
        }
    } else {
        // planar layouts: one byte per pixel and plane, 32 pixels = 2 x uint4 per plane
        // MCR_OBS_GRAY_STACK: a ring of stack_k luma frames per agent; the frame of step s goes to slot s % stack_k and
        // the first frame of an episode (s == 0) to every slot, like gym's FrameStack does on reset
        const bool stack = obs_format == MCR_OBS_GRAY_STACK;
        const int planes = (obs_format == MCR_OBS_GRAY || stack) ? 1 : 3;
        const size_t frame_bytes = (size_t)(stack ? stack_k : planes) * SW * SH;
        if (stack) {
            uint32_t o[8];
#pragma unroll
            for (int k = 0; k < 8; ++k)
                o[k] = S.palY[pix[k] & 0xff] | (S.palY[(pix[k] >> 8) & 0xff] << 8) | (S.palY[(pix[k] >> 16) & 0xff] << 16) | (S.palY[pix[k] >> 24] << 24);
            const int s0 = stack_step == 0 ? 0 : stack_step % stack_k, s1 = stack_step == 0 ? stack_k : s0 + 1;
            for (int sl = s0; sl < s1; ++sl) {
                uint4* dst = reinterpret_cast<uint4*>(obs + (size_t)frame * frame_bytes + (size_t)sl * (SW * SH) + (size_t)out_row * SW + my_seg * 32);
                dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
                dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
            }
            return;
        }
        for (int c = 0; c < planes; ++c) {
            const uint32_t* tab = obs_format == MCR_OBS_GRAY ? S.palY : S.pal32;
            const int sh = 8 * c;
            uint32_t o[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t c0 = (tab[pix[k] & 0xff] >> sh) & 0xffu, c1 = (tab[(pix[k] >> 8) & 0xff] >> sh) & 0xffu;
                const uint32_t c2 = (tab[(pix[k] >> 16) & 0xff] >> sh) & 0xffu, c3 = (tab[pix[k] >> 24] >> sh) & 0xffu;
                o[k] = c0 | (c1 << 8) | (c2 << 16) | (c3 << 24);
            }
            uint4* dst = reinterpret_cast<uint4*>(obs + (size_t)frame * frame_bytes + (size_t)c * (SW * SH) + (size_t)out_row * SW + my_seg * 32);
            dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
            dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
        }
    }
}

// VP = false: the 96 x 96 observation of step() (one CTA per agent-frame, camera and score snapshots
// taken by post_kernel).  VP = true: render(mode) for any viewport (rgb_array 600 x 400, mcr:566-575)
// as a grid of 96 x 96 tiles, blockIdx.y = tile; camera from vp.camera, live score / backward flag
// (what a render() call outside step() shows), RGB rows of vp.vw pixels.
struct VpParams { int vw, vh, tiles_x; float hud_sx, hud_sy; const float* camera; };

template <bool VP>
__global__ void __launch_bounds__(RS_THREADS, 4)
render_kernel(Dims d, DevBuffers b, CarConst cc, const uint8_t* __restrict__ mask, uint8_t* __restrict__ obs,
              int backwards_flag, int use_ego_color, int cls, int obs_format, int stack_k, VpParams vp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RasterSmem& S = *reinterpret_cast<RasterSmem*>(smem_raw);
    PHASE_T0();
    if (!VP) cudaGridDependencySynchronize();      // programmatic dependent launch behind post_kernel (mcr_launch_pdl)
    if (!VP) tl_stamp(b.timeline, cls == 2 ? TL_RENDER2 : TL_RENDER);
    // VP = false: grid (B, A) -- env and agent come from the block index, no integer division per thread
    const int frame = VP ? (int)blockIdx.x : (int)(blockIdx.x * d.A + blockIdx.y);
    const int ox = VP ? (int)(blockIdx.y % vp.tiles_x) * SW : 0, oy = VP ? (int)(blockIdx.y / vp.tiles_x) * SH : 0;
    const int VW = VP ? vp.vw : SW, VH = VP ? vp.vh : SH;
    const float* __restrict__ camera = VP ? vp.camera : b.camera;
    // this tile's rectangle in viewport pixels (partial tiles at the right / top edge)
    const float tx0 = (float)ox, ty0 = (float)oy, tx1 = (float)min(ox + SW, VW), ty1 = (float)min(oy + SH, VH);
    const int env = VP ? frame / d.A : (int)blockIdx.x, agent = VP ? frame % d.A : (int)blockIdx.y;
    if (mask && !mask[env]) return;
    if (cls && (cls == 2) != (b.n_manifold[env] > 0)) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = d.N, car = frame;
    const int slot = b.env_track[env];
    const int Q = b.trk_Q[slot];

    PHASE(0);
    if (tid >= 32 && tid < 32 + PAL_COUNT) {
        const int i = tid - 32;
        S.pal32[i] = (uint32_t)c_palette[i][0] | ((uint32_t)c_palette[i][1] << 8) | ((uint32_t)c_palette[i][2] << 16);
        S.palY[i] = (uint32_t)c_palette[i][3];
        S.pal16[i] = make_uint2(c_pal16[i][0], c_pal16[i][1]);
    }
    if (tid < 256) S.prmt_sel[tid] = prmt_selector(tid);
    if (tid >= 128 && tid < 128 + SPAN_POOL / 32) S.startbits[tid - 128] = 0;   // (the row masks are zeroed per list, in flush_list)
    frame_front<RS_THREADS>(S, d, b, camera, VP ? b.reward[car] : b.score_snap[car], car, slot, tid, tx0, ty0, tx1, ty1);
    PHASE(2);

    View V;
    V.env = env; V.agent = agent; V.A = d.A; V.N = N; V.Q = Q;
    V.ck_j0x = S.ck_j0x; V.ck_j0y = S.ck_j0y; V.ck_ny = S.ck_ny > 0 ? S.ck_ny : 1;
    V.n_checker = S.ck_nx * S.ck_ny; V.grass_full = S.grass_full;
    V.n_road = S.n_vis_chunks * MCR_QUAD_CHUNK; V.vis_chunk = S.vis_chunk;
    V.c_road = (1 + V.n_checker + 31) & ~31;
    V.c_cars = V.c_road + ((V.n_road + 31) & ~31);
    V.car_slots = CAR_PARTS + ((VP && d.particles) ? PRT_MAX * (PRT_PTS - 1) : 0);
    V.prt_pts = b.prt_pts; V.prt_meta = b.prt_meta; V.prt_hdr = b.prt_hdr;
    V.c_hud = V.c_cars + ((V.car_slots * d.A + 31) & ~31);
    V.body = b.body; V.wheel = b.wheel; V.stripe = b.stripe;
    V.quad = b.trk_quad + (size_t)slot * d.Qmax * 8;
    V.quad_col = b.trk_quad_col + (size_t)slot * d.Qmax;
    V.quad_tile = b.trk_quad_tile + (size_t)slot * d.Qmax;
    V.touched = b.touched + (size_t)env * d.Tmax;
    V.use_ego_color = use_ego_color;
    V.backward_flag_on = ((VP ? b.backward[car] : b.backward_snap[car]) != 0) && backwards_flag;   // step(): flag of the PREVIOUS step (render precedes mcr:445-495)
    V.hud_sx = VP ? vp.hud_sx : (float)(96.0 / 1000.0); V.hud_sy = VP ? vp.hud_sy : (float)(96.0 / 800.0);
    const Affine M = S.M;

    uint32_t pix[8];                       // this thread's 32 pixels (palette indices); glClear -> black
#pragma unroll
    for (int k = 0; k < 8; ++k) pix[k] = (V.grass_full ? PAL_GRASS : PAL_BLACK) * 0x01010101u;
    // The HUD bar (mcr:637-642, drawn after the world) covers viewport rows [0, hud_rows) over the full width:
    // nothing of the world shows there, so world polygons start at row hud_rows (same picture, fewer spans).
    int hud_rows = 0;
    {
        const float bar_top = (float)(5 * (800 / 40.0)) * V.hud_sy, bar_right = (float)1000.0 * V.hud_sx;
        if (bar_right >= (float)VW) hud_rows = min(VH, max(0, (int)ceilf(bar_top - 0.5f)));
    }

    // ---- candidates -> ordered display list -> flush -----------------------------------------
    const int NC = V.c_hud + 9;
    int base = 0, lc = 0, pc = 0, round = 0;   // display-list / span-pool fill (same in every thread)
    while (base < NC) {
        const int i = base + tid;
        float px[MCR_MAXV], py[MCR_MAXV];
        int nv = 0, col = 0, y0 = 0, y1 = 0, aux = 0;
        if (i < NC) nv = gen_candidate(i, V, M, cc, px, py, col, aux);
        const bool valid = cand_rows<VP>(nv, px, py, i < V.c_hud, hud_rows, tx0, ty0, tx1, ty1, oy, VH, y0, y1);
        const int rows = valid ? y1 - y0 : 0;
        const int ents = valid ? 1 : 0;
        // block-wide exclusive scan of (entries, rows)
        const int cnt_inc = __popc(__ballot_sync(0xffffffffu, valid) & (0xffffffffu >> (31 - lane)));
        int rows_inc = rows;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int r = __shfl_up_sync(0xffffffffu, rows_inc, o);
            if (lane >= o) rows_inc += r;
        }
        const int par = round & 1; ++round;             // double-buffered warp totals: one barrier per round
        if (lane == 31) { S.warp_cnt[par][warp] = cnt_inc; S.warp_rows[par][warp] = rows_inc; }
        __syncthreads();
        PHASE(3);
        int cnt_before = 0, rows_before = 0, cnt_total = 0, rows_total = 0;
#pragma unroll
        for (int wq = 0; wq < RS_WARPS; ++wq) {
            const int c = S.warp_cnt[par][wq], r = S.warp_rows[par][wq];
            if (wq < warp) { cnt_before += c; rows_before += r; }
            cnt_total += c; rows_total += r;
        }
        const int slot_rel = cnt_before + cnt_inc - ents;
        const int row_rel = rows_before + rows_inc - rows;
        // (lc, pc) are tracked in registers: every thread sees the same totals
        const bool all_fit = (lc + cnt_total <= LIST_CAP) && (pc + rows_total <= SPAN_POOL);
        int first_bad = RS_THREADS;
        if (!all_fit) {                                 // rare (zoomed-out frames): find the first candidate that does not fit
            if (tid == 0) S.first_bad = RS_THREADS;
            __syncthreads();
            const bool fits = (lc + slot_rel + ents <= LIST_CAP) && (pc + row_rel + rows <= SPAN_POOL);
            if (valid && !fits) atomicMin(&S.first_bad, tid);
            __syncthreads();
            first_bad = S.first_bad;
        }
        if (valid && tid < first_bad) {
            const int sl = lc + slot_rel;
            // canonical edges: lower endpoint first, slope hoisted out of the row loop
#pragma unroll
            for (int k = 0; k < MCR_MAXV; ++k)
                if (k < 4 || nv > 4) S.edge[k & 3][k < 4 ? sl : LIST_CAP + aux] = canon_edge(k, nv, px, py);
            S.ne[sl] = (uint8_t)(nv > 4 ? 8 + aux : 4); S.col[sl] = (uint8_t)col;
            const int first = pc + row_rel;                 // this polygon's slots: [first, first + rows)
            S.base[sl] = first - y0;
            if (first & 31) atomicOr(&S.startbits[first >> 5], 1u << (first & 31));
            for (int k = (first + 31) >> 5; (k << 5) < first + rows; ++k) S.slot32_owner[k] = (uint8_t)sl;
        }
        PHASE(4);
        if (first_bad < RS_THREADS) {
            // accepted prefix = everything before the first candidate that did not fit
            if (tid == first_bad) { S.bc_cnt = lc + slot_rel; S.bc_rows = pc + row_rel; }
            __syncthreads();
            lc = S.bc_cnt; pc = S.bc_rows;
            base += first_bad;
            flush_list<VP>(S, tid, pix, lc, pc, false, ox, oy, VW);
            lc = 0; pc = 0;
        } else {
            lc += cnt_total; pc += rows_total;
            base += RS_THREADS;
        }
    }
    flush_list<VP>(S, tid, pix, lc, pc, true, ox, oy, VW);

    finish_frame<VP>(S, tid, pix, obs, frame, obs_format, ox, oy, VW, VH, stack_k, VP ? 0 : b.steps[car]);
    if (!VP && cls != 2 && tid == 0) atomicMax(b.timeline + TL_RENDER_END, mcr_globaltimer());
    PHASE(8);
#ifdef MCR_PHASE_CLOCKS
    if (threadIdx.x == 0) atomicAdd(&g_phase_clk[15], 1ull);
#endif
}


// ---------------------------------------------------------------------------------------
// score_kernel: mcr:433-507 for one car per warp.  Runs concurrently with render_kernel (which
// reads the snapshots post_kernel took of env.reward / driving_backward, because the reference
// renders BEFORE this block).
// ---------------------------------------------------------------------------------------
#define SCORE_WARPS 4

// One car of the block above on one warp; s_cand[32] / s_ncand are that warp's shared scratch.
__device__ __forceinline__ void score_car(const Dims& d, const DevBuffers& b, const uint8_t* __restrict__ mask, const uint8_t* __restrict__ noact,
                                          double* __restrict__ out_reward, uint8_t* __restrict__ out_done, int max_episode_steps, int cls,
                                          int car, int lane, int* s_cand_w, int* s_ncand_w) {
    const int env = car / d.A, agent = car - env * d.A;
    if (mask && !mask[env]) return;
    if (cls && (cls == 2) != (b.n_manifold[env] > 0)) return;
    if (noact && noact[env]) {
        // this env was respawned at the start of the step and took reset()'s step(None): the whole
        // `if action is not None` block of mcr:433-507 is skipped, the caller sees reward 0, done 0
        if (lane == 0) { out_reward[car] = 0.0; if (agent == 0) { out_done[env] = 0; b.pending[env] = 0; } }
        return;
    }
    const int N = d.N;
    const int slot = b.env_track[env];
    const int T = b.trk_T[slot];
    // nearest track node: argmin_i || pos - track[i].xy ||, first minimum wins (np.argmin over
    // sqrt(dx^2 + dy^2) in float64).  Pass 1 finds the minimal squared distance in fp32 (error
    // < 1e-3 near the minimum); pass 2 queues every node within 0.05 of it; lane 0 then evaluates
    // those few exactly as the reference does.
    const float posxf = b.body[(size_t)BF_PX * N + car], posyf = b.body[(size_t)BF_PY * N + car];
    const double* node = b.trk_node + (size_t)slot * d.Tmax * 3;
    float best2 = 3.0e38f;
#pragma unroll 4
    for (int i = lane; i < T; i += 32) {
        const float dx = posxf - (float)node[(size_t)i * 3 + 1], dy = posyf - (float)node[(size_t)i * 3 + 2];
        best2 = fminf(best2, dx * dx + dy * dy);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best2 = fminf(best2, __shfl_xor_sync(0xffffffffu, best2, o));
    if (lane == 0) (*s_ncand_w) = 0;
    __syncwarp();
    for (int i = lane; i < T; i += 32) {
        const float dx = posxf - (float)node[(size_t)i * 3 + 1], dy = posyf - (float)node[(size_t)i * 3 + 2];
        if (dx * dx + dy * dy <= best2 + 0.05f) { const int q = atomicAdd(&(*s_ncand_w), 1); if (q < 32) s_cand_w[q] = i; }
    }
    __syncwarp();
    // driving_on_grass, mcr:469-472: the hull position is strictly inside none of the road_poly quads (shapely
    // Point.within(Polygon), float64).  Lanes take the culling chunks (8 quads, fp32 bounding circle whose radius
    // is padded far beyond the fp32 / float64 vertex difference) and test the quads of the chunks in reach exactly.
    {
        const double posx = posxf, posy = posyf;
        const int Q = b.trk_Q[slot], nchunks = (Q + MCR_QUAD_CHUNK - 1) / MCR_QUAD_CHUNK;
        const float4* chunk = (const float4*)(b.trk_chunk + (size_t)slot * (d.Qmax / MCR_QUAD_CHUNK) * 4);
        const double* quad64 = b.trk_quad64 + (size_t)slot * d.Qmax * 8;
        // pass 1: the chunks in reach (warp-wide bitmask); pass 2: one cross product per lane -- lane = (quad of the
        // chunk, edge) -- and two ballots decide "all four edges on the same strict side" for the chunk's 8 quads
        unsigned reach[MAX_CHUNKS / 32];
#pragma unroll
        for (int w = 0; w < MAX_CHUNKS / 32; ++w) {
            const int c = 32 * w + lane;
            bool hit = false;
            if (c < nchunks) {
                const float4 cc4 = chunk[c];
                const float dx = posxf - cc4.x, dy = posyf - cc4.y;
                hit = !(dx * dx + dy * dy > cc4.z * cc4.z + 1.0f);
            }
            reach[w] = __ballot_sync(0xffffffffu, hit);
        }
        bool inside = false;
#pragma unroll
        for (int w = 0; w < MAX_CHUNKS / 32; ++w) {
            unsigned m = reach[w];
            while (m) {
                const int c = 32 * w + (__ffs((int)m) - 1);
                m &= m - 1u;
                const int q = c * MCR_QUAD_CHUNK + (lane >> 2), k = lane & 3, k2 = (k + 1) & 3;
                double cr = 0.0;
                if (q < Q) {
                    const double* v = quad64 + (size_t)q * 8;
                    const double ex = v[2 * k2] - v[2 * k], ey = v[2 * k2 + 1] - v[2 * k + 1];
                    cr = ex * (posy - v[2 * k + 1]) - ey * (posx - v[2 * k]);
                }
                const unsigned pm = __ballot_sync(0xffffffffu, cr > 0), nm = __ballot_sync(0xffffffffu, cr < 0);
                const unsigned allp = pm & (pm >> 1) & (pm >> 2) & (pm >> 3) & 0x11111111u;
                const unsigned alln = nm & (nm >> 1) & (nm >> 2) & (nm >> 3) & 0x11111111u;
                inside = inside || (allp | alln) != 0u;
            }
        }
        const bool any_inside = __any_sync(0xffffffffu, inside);
        if (lane == 0) b.on_grass[car] = any_inside ? 0 : 1;
    }
    if (lane == 0) {
        const double posx = posxf, posy = posyf;
        double bestd = 1.0e300; int besti = 0x7fffffff;
        const int nc = (*s_ncand_w);
        if (nc <= 32) {
            for (int q = 0; q < nc; ++q) {
                const int i = s_cand_w[q];
                const double dx = posx - node[(size_t)i * 3 + 1], dy = posy - node[(size_t)i * 3 + 2];
                const double dd = sqrt(dx * dx + dy * dy);
                if (dd < bestd || (dd == bestd && i < besti)) { bestd = dd; besti = i; }
            }
        } else {                                    // pathological (many equidistant nodes): exact scan
            for (int i = 0; i < T; ++i) {
                const double dx = posx - node[(size_t)i * 3 + 1], dy = posy - node[(size_t)i * 3 + 2];
                const double dd = sqrt(dx * dx + dy * dy);
                if (dd < bestd) { bestd = dd; besti = i; }
            }
        }
        const double PI = 3.141592653589793, PLAYFIELD = 2000 / 6.0;
        double reward = b.reward[car] - 0.1;                     // mcr:436
        double step_reward = reward - b.prev_reward[car];        // mcr:443
        const double car_angle = b.heading[car];                 // mcr:449-456, evaluated in the physics kernel
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
        // every agent's CTA of this env; only agent 0 stores it)
        uint8_t done = 0;
        for (int c = 0; c < d.A; ++c) {
            const int oc = env * d.A + c;
            if (b.visit_count[oc] == T) done = 1;
            const double x = b.body[(size_t)BF_PX * N + oc], y = b.body[(size_t)BF_PY * N + oc];
            if (fabs(x) > PLAYFIELD || fabs(y) > PLAYFIELD) { done = 1; if (c == agent) step_reward = -100; }
        }
        if (max_episode_steps > 0 && b.steps[car] >= max_episode_steps) done |= 2;   // TimeLimit
        out_reward[car] = step_reward;
        if (agent == 0) { out_done[env] = done; b.pending[env] = done; }
    }
}

__global__ void __launch_bounds__(SCORE_WARPS * 32)
score_kernel(Dims d, DevBuffers b, const uint8_t* __restrict__ mask, const uint8_t* __restrict__ noact,
             double* __restrict__ out_reward, uint8_t* __restrict__ out_done, int max_episode_steps, int cls) {
    tl_stamp(b.timeline, TL_SCORE);
    __shared__ int s_cand[SCORE_WARPS][32];
    __shared__ int s_ncand[SCORE_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int car = blockIdx.x * SCORE_WARPS + warp;
    if (car >= d.N) return;
    score_car(d, b, mask, noact, out_reward, out_done, max_episode_steps, cls, car, lane, s_cand[warp], &s_ncand[warp]);
}


// ---------------------------------------------------------------------------------------
// The step path renders in two kernels (the fused render_kernel above stays for the viewport modes):
//   project_kernel  front end, one CTA of 256 threads per agent-frame: candidates in painter's order -> camera ->
//                   cull -> ordered compaction (block scan) -> canonical edges, written as the frame's display list
//                   to global memory (stays in L2).  Latency bound (three dependent load levels), light on
//                   registers and shared memory: many CTAs per SM hide the latency the fused kernel's idle warps
//                   used to wait out at block barriers.
//   fill_kernel     back end, one CTA of 288 threads per agent-frame: display list -> shared memory in chunks that
//                   fit the list / span pool -> spans -> fill -> palette expansion -> 6 x uint4 per thread.
// Same arithmetic as the fused kernel, polygon by polygon (cand_rows, canon_edge, flush_list, finish_frame are shared).
// ---------------------------------------------------------------------------------------
#ifndef FILL_CTAS
#define FILL_CTAS 4
#endif
#define PJ_THREADS 256
#define PJ_WARPS (PJ_THREADS / 32)

// checker index range the camera can see + "the playfield quad covers the whole frame" (see frame_front)
struct CheckerRange { int j0x, nx, j0y, ny, grass_full; };
__device__ __forceinline__ CheckerRange checker_range(float m00, float m01, float m02, float m10, float m11, float m12,
                                                      float tx0, float ty0, float tx1, float ty1) {
    const float det = m00 * m11 - m01 * m10;
    int j0x = 0, j1x = N_CHECKER_AXIS - 1, j0y = 0, j1y = N_CHECKER_AXIS - 1, grass_full = 0;
    if (fabsf(det) > 1e-12f) {
        const float inv = 1.0f / det;
        float wxmin = 3.0e38f, wxmax = -3.0e38f, wymin = 3.0e38f, wymax = -3.0e38f;
#pragma unroll
        for (int cnr = 0; cnr < 4; ++cnr) {
            const float u = ((cnr & 1) ? tx1 : tx0) - m02, v = ((cnr & 2) ? ty1 : ty0) - m12;
            const float wx = (m11 * u - m01 * v) * inv, wy = (-m10 * u + m00 * v) * inv;
            wxmin = fminf(wxmin, wx); wxmax = fmaxf(wxmax, wx); wymin = fminf(wymin, wy); wymax = fmaxf(wymax, wy);
        }
        const float pfm = (float)(2000 / 6.0) - 4.0f;
        grass_full = (wxmin > -pfm && wxmax < pfm && wymin > -pfm && wymax < pfm) ? 1 : 0;
        const float ik = 20.0f / (float)(2000 / 6.0), margin = 2.0f;
        const float a0 = floorf(((wxmin - margin) * ik + 19.0f) * 0.5f - 1e-3f), a1 = ceilf(((wxmax + margin) * ik + 20.0f) * 0.5f + 1e-3f);
        const float c0 = floorf(((wymin - margin) * ik + 19.0f) * 0.5f - 1e-3f), c1 = ceilf(((wymax + margin) * ik + 20.0f) * 0.5f + 1e-3f);
        if (a0 == a0 && a1 == a1 && c0 == c0 && c1 == c1) {
            j0x = (int)fmaxf(0.0f, fminf(a0, 20.0f)); j1x = (int)fminf((float)(N_CHECKER_AXIS - 1), fmaxf(a1, -1.0f));
            j0y = (int)fmaxf(0.0f, fminf(c0, 20.0f)); j1y = (int)fminf((float)(N_CHECKER_AXIS - 1), fmaxf(c1, -1.0f));
        }
    }
    CheckerRange r;
    r.j0x = j0x; r.nx = j1x >= j0x ? j1x - j0x + 1 : 0; r.j0y = j0y; r.ny = j1y >= j0y ? j1y - j0y + 1 : 0; r.grass_full = grass_full;
    return r;
}

// PJ_W warps per agent-frame (one CTA per frame).  The frame's candidates are cut into passes of 32 in painter's order;
// pass p is worked by warp p % PJ_W -- loads, camera transform, cull and the row scan of different passes run side by
// side -- and only the hand-off of the running (entries, span slots) prefix goes in pass order: the warp of pass p
// waits for the prefix pass p - 1 published in shared memory, publishes its own and writes its entries.  Dependencies
// only point to earlier passes and every warp takes its passes in increasing order, so the wait always ends.
#ifdef MCR_PHASE_CLOCKS
#define PJCLK_T0() long long pj_t_ = clock64()
#define PJCLK(k) do { if (threadIdx.x == 0) { const long long n_ = clock64(); atomicAdd(&g_phase_clk[k], (unsigned long long)(n_ - pj_t_)); pj_t_ = n_; } } while (0)
#else
#define PJCLK_T0() do {} while (0)
#define PJCLK(k) do {} while (0)
#endif
// small batches (one wave of CTAs) are latency bound and finish earlier with two warps per frame; large batches are
// throughput bound and do less redundant set-up work with one
#define PJ_W_SMALL 2
#define PJ_SMALL_FRAMES 8192
#define PJ_ONE_WAVE 2304        // frames whose CTAs are all resident at once (148 SMs x 16 CTAs, less what post_kernel still holds)
#define PJ_MAX_PASS 96           // (1 + 100) / 32 + 2048 / 32 + (12 * 16 + 9) / 32 + slack

// The reward / done block of the step (score_car, mcr:433-507) rides in the same launch: CTAs past the last frame take
// PJ_W cars each, one warp per car.  It only needs what post_kernel left, like the projector, so it used to be a kernel
// of its own on a side stream -- but the event record that forked it sat between post_kernel and this kernel in the
// main stream and cost the programmatic dependent launch (6-7 us of launch latency on the critical chain).
struct ScoreArgs { const uint8_t* noact; double* out_reward; uint8_t* out_done; int max_episode_steps; int enabled; };

template <int PJ_W>
__global__ void __launch_bounds__(PJ_W * 32, 32 / PJ_W)
project_kernel(Dims d, DevBuffers b, CarConst cc, const uint8_t* __restrict__ mask, int backwards_flag, int use_ego_color, int cls,
               ScoreArgs sa, int wait_post) {
    __shared__ uint8_t s_vis_chunk[MAX_CHUNKS];
    __shared__ int s_nvis;
    __shared__ unsigned long long s_pre[PJ_MAX_PASS + 1];   // prefix BEFORE pass p: entries << 32 | span slots; ~0 = not published yet
    __shared__ __align__(16) uint8_t s_touched[2048];   // this env's "tile colour was reset" flags (Tmax <= Qmax <= 2048): loaded
                                                        // beside the camera, so the road passes do not wait for a dependent load
    PJCLK_T0();
    // wait_post (the step's main chain): post_kernel triggered this launch at its start and publishes every car with a
    // ready flag; the CTA waits below for the cars of its own env only.  Otherwise: plain programmatic dependent launch
    // behind whatever precedes in the stream.  Either way fill_kernel may be launched from here on -- its CTAs wait for
    // their own frame's flag (set at the end of this kernel).
    if (!wait_post || (int)blockIdx.x >= d.N) cudaGridDependencySynchronize();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if ((int)blockIdx.x >= d.N) {                  // score CTAs
        __shared__ int s_cand[PJ_W][32];
        __shared__ int s_ncand[PJ_W];
        if ((int)blockIdx.x == d.N && threadIdx.x == 0) tl_stamp_any(b.timeline, TL_SCORE);
        const int car = ((int)blockIdx.x - d.N) * PJ_W + warp;
        if (car < d.N) score_car(d, b, mask, sa.noact, sa.out_reward, sa.out_done, sa.max_episode_steps, cls, car, lane, s_cand[warp], &s_ncand[warp]);
        return;
    }
    const int frame = (int)blockIdx.x;
    PJCLK(9);
    const int env = frame / d.A, agent = frame - env * d.A;
    if (mask && !mask[env]) return;
    if (cls && (cls == 2) != (b.n_manifold[env] > 0)) return;      // (written by the head of the chain: complete long ago)
    const int N = d.N, car = frame;
    if (wait_post) {
        // every car of the env is drawn in this frame: lanes 0..A-1 of each warp wait for one car each
        const int epoch = flag_peek(b.ready + READY_EPOCH(N));
        if (lane < d.A) { while (flag_peek(b.ready + env * d.A + lane) != epoch) __nanosleep(40); }
        flag_fence_acquire();
        __syncwarp();
    }
    // fill_kernel may be launched once every CTA of this grid is here (or has left).  Not earlier: its CTAs wait for
    // registers until CTAs of this kernel retire, and the hardware places grids in launch order -- launched at the top of
    // this kernel (measured, profiles/README r02) the pending grid held score_kernel, which is launched when post_kernel
    // completes, back by 30 us; a persistent score grid launched ahead of it and waiting for post_kernel's flags took
    // project_kernel's one-wave residency instead.  Here post_kernel is complete, so score_kernel is launched first.
    // Only while this grid is a single wave: with more frames, fill_kernel's CTAs would be placed while later waves of
    // this kernel are still to come and sit on their shared memory and registers waiting for their frame (measured:
    // the rasteriser alone 7-10 % slower from 4096 envs up); there fill_kernel starts when this grid has completed.
    // (and only on the step's chain: for a stand-alone render pass -- every CTA of this grid ends at about the same time --
    // the early launch buys nothing and the flag costs fill_kernel's CTAs an L2 round trip each)
    const bool flag_to_fill = wait_post && d.N <= PJ_ONE_WAVE;      // launch_fill takes the same decision (fill_kernel's wait_flags)
    if (flag_to_fill) cudaTriggerProgrammaticLaunchCompletion();
    tl_stamp(b.timeline, cls == 2 ? TL_RENDER2 : TL_RENDER);
    if (cls != 2 && threadIdx.x == 0) atomicMax(b.timeline + TL_PROJECT_LAST_GO, mcr_globaltimer());
    const int slot = b.env_track[env];
    Affine M;
    M.m00 = b.camera[(size_t)0 * N + car]; M.m01 = b.camera[(size_t)1 * N + car]; M.m02 = b.camera[(size_t)2 * N + car];
    M.m10 = b.camera[(size_t)3 * N + car]; M.m11 = b.camera[(size_t)4 * N + car]; M.m12 = b.camera[(size_t)5 * N + car];
    const int Q = b.trk_Q[slot];
    for (int i = threadIdx.x; i <= PJ_MAX_PASS; i += PJ_W * 32) s_pre[i] = i == 0 ? 0ull : ~0ull;
    if ((d.Tmax & 15) == 0) {
        const uint4* src = reinterpret_cast<const uint4*>(b.touched + (size_t)env * d.Tmax);
        for (int i = threadIdx.x; i < d.Tmax / 16; i += PJ_W * 32) reinterpret_cast<uint4*>(s_touched)[i] = src[i];
    } else {
        for (int i = threadIdx.x; i < d.Tmax; i += PJ_W * 32) s_touched[i] = b.touched[(size_t)env * d.Tmax + i];
    }
    // ---- road_poly chunk culling (bounding circle of every 8 consecutive quads vs the frame), warp 0, lane per chunk;
    // chunks past the track's last quad have radius 0 in the pool, so the loop bound does not wait for trk_Q
    if (warp == 0) {
        int n_vis = 0;
        const int nchunks = d.Qmax / MCR_QUAD_CHUNK;
        const float rscale = sqrtf(M.m00 * M.m00 + M.m01 * M.m01 + M.m10 * M.m10 + M.m11 * M.m11) * 1.001f;
        for (int c0 = 0; c0 < nchunks; c0 += 128) {          // four circles per lane in flight: one memory round trip per 128 chunks
            float4 cc4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = c0 + 32 * u + lane;
                cc4[u] = c < nchunks ? *(const float4*)(b.trk_chunk + ((size_t)slot * nchunks + c) * 4) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = c0 + 32 * u + lane;
                const float cxp = (M.m00 * cc4[u].x + M.m01 * cc4[u].y) + M.m02, cyp = (M.m10 * cc4[u].x + M.m11 * cc4[u].y) + M.m12;
                const float rp = cc4[u].z * rscale + 2.0f;   // |M v| <= ||M||_F |v|: conservative pixel radius, plus a 2-pixel margin
                bool vis = (cxp + rp >= 0.0f) && (cxp - rp <= (float)SW) && (cyp + rp >= 0.0f) && (cyp - rp <= (float)SH);
                if (!(rp == rp) || !(cxp == cxp) || !(cyp == cyp)) vis = true;
                if (!(cc4[u].z > 0.0f)) vis = false;
                const uint32_t bal = __ballot_sync(0xffffffffu, vis);
                if (vis) s_vis_chunk[n_vis + __popc(bal & ((1u << lane) - 1u))] = (uint8_t)c;
                n_vis += __popc(bal);
            }
        }
        if (lane == 0) s_nvis = n_vis;
    }
    const CheckerRange ck = checker_range(M.m00, M.m01, M.m02, M.m10, M.m11, M.m12, 0.0f, 0.0f, (float)SW, (float)SH);
    __syncthreads();
    PJCLK(10);
    const uint8_t* vis_chunk = s_vis_chunk;
    const int n_vis = s_nvis;

    View V;
    V.env = env; V.agent = agent; V.A = d.A; V.N = N; V.Q = Q;
    V.ck_j0x = ck.j0x; V.ck_j0y = ck.j0y; V.ck_ny = ck.ny > 0 ? ck.ny : 1;
    V.n_checker = ck.nx * ck.ny; V.grass_full = ck.grass_full;
    V.n_road = n_vis * MCR_QUAD_CHUNK; V.vis_chunk = vis_chunk;
    V.c_road = (1 + V.n_checker + 31) & ~31;
    V.c_cars = V.c_road + ((V.n_road + 31) & ~31);
    V.car_slots = CAR_PARTS;
    V.prt_pts = b.prt_pts; V.prt_meta = b.prt_meta; V.prt_hdr = b.prt_hdr;
    V.c_hud = V.c_cars + ((V.car_slots * d.A + 31) & ~31);
    V.body = b.body; V.wheel = b.wheel; V.stripe = b.stripe;
    V.quad = b.trk_quad + (size_t)slot * d.Qmax * 8;
    V.quad_col = b.trk_quad_col + (size_t)slot * d.Qmax;
    V.quad_tile = b.trk_quad_tile + (size_t)slot * d.Qmax;
    V.touched = s_touched;
    V.use_ego_color = use_ego_color;
    V.backward_flag_on = (b.backward_snap[car] != 0) && backwards_flag;     // step(): flag of the PREVIOUS step (render precedes mcr:445-495)
    V.hud_sx = (float)(96.0 / 1000.0); V.hud_sy = (float)(96.0 / 800.0);
    int hud_rows = 0;                              // see render_kernel
    {
        const float bar_top = (float)(5 * (800 / 40.0)) * V.hud_sy, bar_right = (float)1000.0 * V.hud_sx;
        if (bar_right >= (float)SW) hud_rows = min(SH, max(0, (int)ceilf(bar_top - 0.5f)));
    }
    uint2* __restrict__ meta = reinterpret_cast<uint2*>(b.dl_meta) + (size_t)frame * d.dl_cap;
    float4* __restrict__ edge = reinterpret_cast<float4*>(b.dl_edge) + (size_t)frame * d.dl_cap * 4;
    float4* __restrict__ oct = reinterpret_cast<float4*>(b.dl_oct) + (size_t)frame * d.A * 4;

    // passes in painter's order: [0, P0) playfield + checker squares, [P0, P0 + P1) road_poly, [P0 + P1, NP) cars + HUD
    const int NC = V.c_hud + 9;
    const int P0 = V.c_road >> 5, P1 = (V.c_cars - V.c_road) >> 5, NP = min(P0 + P1 + ((NC - V.c_cars + 31) >> 5), PJ_MAX_PASS);
    // road_poly quads of the visible chunks (mcr:628-631), four chunks per pass: a warp loads the quads of its NEXT road
    // pass before it works on the current pass
    struct RoadQuad { int q; float4 a, b; int tile, col; };
    auto road_load = [&](int p) {
        RoadQuad r; r.q = -1; r.a = make_float4(0, 0, 0, 0); r.b = r.a; r.tile = -1; r.col = 0;
        if (p >= P0 && p < P0 + P1) {
            const int k = p * 32 + lane - V.c_road;
            if (k < V.n_road) { r.q = (int)vis_chunk[k >> 3] * MCR_QUAD_CHUNK + (k & (MCR_QUAD_CHUNK - 1)); if (r.q >= Q) r.q = -1; }
            if (r.q >= 0) {
                r.a = *(const float4*)(V.quad + (size_t)r.q * 8); r.b = *(const float4*)(V.quad + (size_t)r.q * 8 + 4);
                r.tile = V.quad_tile[r.q]; r.col = V.quad_col[r.q];
            }
        }
        return r;
    };
    RoadQuad nextq = road_load(warp);
    for (int p = warp; p < NP; p += PJ_W) {
        const int i = p * 32 + lane;               // candidate index (the classes start on multiples of 32)
        float px[MCR_MAXV], py[MCR_MAXV];
        int nv = 0, col = 0, y0 = 0, y1 = 0, aux = 0;
        const RoadQuad rq = nextq;
        nextq = road_load(p + PJ_W);
        if (p >= P0 && p < P0 + P1) {
            const int q = rq.q;
            if (q >= 0) {
                const float4 qa = rq.a, qb = rq.b;
                const int tl = rq.tile;
                col = rq.col;
                if (tl >= 0 && V.touched[tl]) col = PAL_ROAD0;                // tile.color reset, mcr:102-104
                xf_pt(M, qa.x, qa.y, px[0], py[0]); xf_pt(M, qa.z, qa.w, px[1], py[1]);
                xf_pt(M, qb.x, qb.y, px[2], py[2]); xf_pt(M, qb.z, qb.w, px[3], py[3]);
                nv = 4;
            }
        } else if (i < NC) {
            nv = gen_candidate(i, V, M, cc, px, py, col, aux);
        }
        const bool valid = cand_rows<false>(nv, px, py, i < V.c_hud, hud_rows, 0.0f, 0.0f, (float)SW, (float)SH, 0, SH, y0, y1);
        const uint32_t bal = __ballot_sync(0xffffffffu, valid);
        const int rows = valid ? y1 - y0 : 0;
        int rows_inc = rows;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int r = __shfl_up_sync(0xffffffffu, rows_inc, o);
            if (lane >= o) rows_inc += r;
        }
        const int pass_rows = __shfl_sync(0xffffffffu, rows_inc, 31);
        PJCLK(11);
        // ---- in-order hand-off of the running prefix -------------------------------------------------------------
        int lc = 0, pc = 0;
        if (lane == 0) {
            // one 64-bit word per pass carries the prefix AND its "published" state, read and written with shared-memory
            // atomics (nothing else has to be ordered around it; compute-sanitizer racecheck accepts atomics only)
            unsigned long long v;
            while ((v = atomicAdd(&s_pre[p], 0ull)) == ~0ull) __nanosleep(20);
            lc = (int)(v >> 32); pc = (int)(v & 0xffffffffull);
            atomicExch(&s_pre[p + 1], ((unsigned long long)(unsigned)(lc + __popc(bal)) << 32) | (unsigned long long)(unsigned)(pc + pass_rows));
        }
        lc = __shfl_sync(0xffffffffu, lc, 0); pc = __shfl_sync(0xffffffffu, pc, 0);
        PJCLK(12);
        if (valid) {
            const int sl = lc + __popc(bal & ((1u << lane) - 1u));          // < dl_cap: the list holds every candidate
            const int first = pc + rows_inc - rows;                         // this polygon's span slots: [first, first + rows)
#pragma unroll
            for (int k = 0; k < 4; ++k) edge[(size_t)sl * 4 + k] = canon_edge(k, nv, px, py);
            if (nv > 4) {
#pragma unroll
                for (int k = 4; k < MCR_MAXV; ++k) oct[(size_t)aux * 4 + (k - 4)] = canon_edge(k, nv, px, py);
            }
            meta[sl] = make_uint2((uint32_t)y0 | ((uint32_t)rows << 8) | ((uint32_t)col << 16) | ((uint32_t)(nv > 4 ? 8 + aux : 4) << 24),
                                  (uint32_t)first);
        }
        if (p == NP - 1 && lane == 0)
            *reinterpret_cast<int4*>(b.dl_hdr + (size_t)frame * 4) = make_int4(lc + __popc(bal), pc + pass_rows, V.grass_full, 0);
        PJCLK(13);
    }
    // the frame's display list is complete when every warp has stored its passes: publish it (ready[N + frame])
    if (flag_to_fill) {
        __syncthreads();                           // (barrier + release store: the release is cumulative over the other threads' stores)
        if (threadIdx.x == 0) flag_release(b.ready + N + frame, 1);
    }
    if (cls != 2 && threadIdx.x == 0) { const unsigned long long t = mcr_globaltimer(); atomicMax(b.timeline + TL_PROJECT_END, t); atomicMax(b.timeline + TL_PROJECT_FIRST_END, ~t); }
#ifdef MCR_PHASE_CLOCKS
    if (threadIdx.x == 0) atomicAdd(&g_phase_clk[14], 1ull);
#endif
}

// One CTA per agent-frame.  (Measured and dropped, profiles/README r02: persistent CTAs with a frame queue or a static
// stride, register prefetch of the next frame's list, 5 CTAs per SM at 40 registers -- none beat this plain form.)
__global__ void __launch_bounds__(RS_THREADS, FILL_CTAS)
fill_kernel(Dims d, DevBuffers b, const uint8_t* __restrict__ mask, uint8_t* __restrict__ obs, int cls, int obs_format, int stack_k, int env0,
            int wait_flags) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RasterSmem& S = *reinterpret_cast<RasterSmem*>(smem_raw);
    const int env = env0 + (int)blockIdx.x, agent = (int)blockIdx.y, frame = env * d.A + agent;   // env0: launches over a range of envs (step_host chunks)
    const int tid = threadIdx.x;
    PHASE_T0();
    // tables that do not depend on project_kernel's output: built while it drains (programmatic dependent launch)
    if (tid >= 32 && tid < 32 + PAL_COUNT) {
        const int i = tid - 32;
        S.pal32[i] = (uint32_t)c_palette[i][0] | ((uint32_t)c_palette[i][1] << 8) | ((uint32_t)c_palette[i][2] << 16);
        S.palY[i] = (uint32_t)c_palette[i][3];
        S.pal16[i] = make_uint2(c_pal16[i][0], c_pal16[i][1]);
    }
    if (tid < 256) S.prmt_sel[tid] = prmt_selector(tid);
    if (tid >= 128 && tid < 128 + SPAN_POOL / 32) S.startbits[tid - 128] = 0;
    if (mask && !mask[env]) return;
    if (cls && (cls == 2) != (b.n_manifold[env] > 0)) return;
    // wait_flags (project_kernel's grid is one wave and triggers this launch before it ends): wait for this frame's display
    // list only, and take the flag back.  Otherwise project_kernel has completed when this grid is placed, and the plain
    // grid dependency spares every CTA the flag's L2 round trip in front of its loads (10 % of the rasteriser from 4096 envs up).
    if (!wait_flags) cudaGridDependencySynchronize();
    else if (tid == 0) {
        if (cls != 2) { const unsigned long long t = mcr_globaltimer(); atomicMax(b.timeline + TL_FILL_FIRST_IN, ~t); atomicMax(b.timeline + TL_FILL_LAST_IN, t); }
        flag_wait(b.ready + d.N + frame); b.ready[d.N + frame] = 0;
        if (cls != 2) atomicMax(b.timeline + TL_FILL_FIRST_GO, ~mcr_globaltimer());
    }
    if (wait_flags) __syncthreads();
    if (cls != 2) tl_stamp(b.timeline, TL_FILL);
    const uint2* __restrict__ meta = reinterpret_cast<const uint2*>(b.dl_meta) + (size_t)frame * d.dl_cap;
    const float4* __restrict__ edge = reinterpret_cast<const float4*>(b.dl_edge) + (size_t)frame * d.dl_cap * 4;
    const float4* __restrict__ oct = reinterpret_cast<const float4*>(b.dl_oct) + (size_t)frame * d.A * 4;
    // header, metadata and the first edges are loaded side by side (no load waits for the entry count)
    const int4 hdr = *reinterpret_cast<const int4*>(b.dl_hdr + (size_t)frame * 4);
    uint2 mt = make_uint2(0u, 0u);
    if (tid < LIST_CAP) mt = meta[min(tid, d.dl_cap - 1)];
    const float4 ed0 = edge[min(tid, d.dl_cap * 4 - 1)], ed1 = edge[min(tid + RS_THREADS, d.dl_cap * 4 - 1)];
    // edges 4..7 of every car's hull octagon go to their fixed place behind the list, whether or not the octagon made it
    // into the list (no load waits for the metadata)
    if (tid < d.A * 4) S.edge[tid & 3][LIST_CAP + (tid >> 2)] = oct[tid];
    const int n = hdr.x;
    if (tid >= 64 && tid < 68) S.glyph[tid - 64] = score_glyph(b.score_snap[frame], tid - 64);
    uint32_t pix[8];                       // this thread's 32 pixels (palette indices); glClear -> black
#pragma unroll
    for (int k = 0; k < 8; ++k) pix[k] = (hdr.z ? PAL_GRASS : PAL_BLACK) * 0x01010101u;
    PHASE(0);
    int e0 = 0;
    uint32_t rs0 = 0u;                     // first span slot of the chunk (the frame's first polygon starts at slot 0)
    const bool whole = n <= LIST_CAP && hdr.y <= SPAN_POOL;     // the whole frame fits one list (the header holds both totals)
    do {
        // this chunk = the longest run of entries from e0 that fits the shared-memory list and the span pool
        if (e0 > 0) {
            rs0 = meta[e0].y;
            if (tid < LIST_CAP) mt = meta[min(e0 + tid, d.dl_cap - 1)];
        }
        // the common case -- the whole frame fits one list (the header holds its entry and span-slot totals): no count, no barrier
        const bool fits = tid < LIST_CAP && e0 + tid < n && mt.y + ((mt.x >> 8) & 0xffu) - rs0 <= (uint32_t)SPAN_POOL;
        const int cnt = whole ? n : __syncthreads_count(fits);         // slot offsets are monotone: the entries that fit are a prefix
        if (tid < cnt) {
            const int y0 = (int)(mt.x & 0xffu), rows = (int)((mt.x >> 8) & 0xffu), ne = (int)(mt.x >> 24);
            const int first = (int)(mt.y - rs0);
            S.base[tid] = first - y0; S.ne[tid] = (uint8_t)ne; S.col[tid] = (uint8_t)((mt.x >> 16) & 0xffu);
            if (first & 31) atomicOr(&S.startbits[first >> 5], 1u << (first & 31));
            for (int k = (first + 31) >> 5; (k << 5) < first + rows; ++k) S.slot32_owner[k] = (uint8_t)tid;
            if (tid == cnt - 1) S.bc_rows = first + rows;
        }
        if (e0 == 0) {
            if (tid < cnt * 4) S.edge[tid & 3][tid >> 2] = ed0;
            if (tid + RS_THREADS < cnt * 4) S.edge[(tid + RS_THREADS) & 3][(tid + RS_THREADS) >> 2] = ed1;
            for (int i = tid + 2 * RS_THREADS; i < cnt * 4; i += RS_THREADS) S.edge[i & 3][i >> 2] = edge[i];
        } else {
            for (int i = tid; i < cnt * 4; i += RS_THREADS) S.edge[i & 3][i >> 2] = edge[(size_t)e0 * 4 + i];
        }
        PHASE(3);
        const bool last = e0 + cnt >= n;
        flush_list<false>(S, tid, pix, cnt, -1, last, 0, 0, SW);
        e0 += cnt;
    } while (e0 < n);
#ifdef MCR_FILL_LANE_BULK
    // A/B build (python -m multi_car_racing_b200.build --lane-bulk -> libmcr_lb.so): every thread sends its own 96 staged bytes
    // with its own cp.async.bulk -- no block barrier in front of the store, 288 small TMA requests per frame instead of one
    if (obs_format == MCR_OBS_RGB_HWC) {
        unsigned char* stage = reinterpret_cast<unsigned char*>(&S.edge[0][0]);
        finish_frame<false>(S, tid, pix, stage, 0, obs_format, 0, 0, SW, SH, 1, 0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const int my_seg = tid / SH, my_y = tid - SH * my_seg;
        const unsigned off = (unsigned)((SH - 1 - my_y) * (SW * 3) + my_seg * 96);
        const unsigned long long gdst = (unsigned long long)(obs + (size_t)frame * MCR_OBS_BYTES + off);
        const unsigned ssrc = (unsigned)__cvta_generic_to_shared(stage + off);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(gdst), "r"(ssrc), "r"(96u) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    } else
#endif
#ifndef MCR_FILL_STG_STORE
    // The reference's RGB HWC frame is staged in shared memory (over the display list / span pool / row masks, which
    // are dead by now) and leaves as ONE cp.async.bulk (TMA, UBLKCP.G.S) of 27 648 contiguous bytes issued by thread 0.
    // Each warp-level STG.128 of the direct store touches 32 rows at 16 bytes -- half a sector each: the bulk copy
    // writes whole sectors (L2 write requests 4.8 M -> 0.68 M, sectors 5.2 M -> 2.7 M per 2048 frames) and is 5-8 % faster
    // end to end (profiles/README r02; A/B build: python -m multi_car_racing_b200.build --stg-store -> libmcr_stg.so).
    if (obs_format == MCR_OBS_RGB_HWC || obs_format == MCR_OBS_GRAY || obs_format == MCR_OBS_RGB_CHW) {   // one contiguous block per frame
        unsigned char* stage = reinterpret_cast<unsigned char*>(&S.edge[0][0]);
        static_assert(offsetof(RasterSmem, stage_tail) == offsetof(RasterSmem, edge) + sizeof(((RasterSmem*)0)->edge) &&
                      offsetof(RasterSmem, base) - offsetof(RasterSmem, edge) >= MCR_OBS_BYTES, "staging area = edge + stage_tail");
        finish_frame<false>(S, tid, pix, stage, 0, obs_format, 0, 0, SW, SH, 1, 0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            const unsigned frame_bytes = obs_format == MCR_OBS_GRAY ? (unsigned)(SW * SH) : (unsigned)MCR_OBS_BYTES;
            const unsigned long long gdst = (unsigned long long)(obs + (size_t)frame * frame_bytes);
            const unsigned ssrc = (unsigned)__cvta_generic_to_shared(stage);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(gdst), "r"(ssrc), "r"(frame_bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    } else
#endif
    finish_frame<false>(S, tid, pix, obs, frame, obs_format, 0, 0, SW, SH, stack_k, b.steps[frame]);
    if (cls != 2 && tid == 0) atomicMax(b.timeline + TL_RENDER_END, mcr_globaltimer());
#ifdef MCR_PHASE_CLOCKS
    if (threadIdx.x == 0) atomicAdd(&g_phase_clk[15], 1ull);
#endif
}


// One-time set-up per DEVICE (function attributes and __constant__ data belong to the device's context; a
// process that drives several GPUs has one handle per device).
static bool configure_render() {
    static bool configured[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
    if (!configured[dev]) {
        const size_t smem = sizeof(RasterSmem);
        if (cudaMemcpyToSymbol(c_palette, mcr_host_palette(), sizeof(uint8_t) * PAL_COUNT * 4) != cudaSuccess) return false;
        uint32_t p16[PAL_COUNT][2];
        for (int i = 0; i < PAL_COUNT; ++i) {
            const uint8_t* c = mcr_host_palette()[i];
            p16[i][0] = (uint32_t)host_f2h((float)c[0] / 255.0f) | ((uint32_t)host_f2h((float)c[1] / 255.0f) << 16);
            p16[i][1] = (uint32_t)host_f2h((float)c[2] / 255.0f);
        }
        if (cudaMemcpyToSymbol(c_pal16, p16, sizeof(p16)) != cudaSuccess) return false;
        if (cudaFuncSetAttribute(render_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false;
        if (cudaFuncSetAttribute(render_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false;
        if (cudaFuncSetAttribute(fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false;
        configured[dev] = true;
    }
    return true;
}

static void launch_project_impl(const Dims& d, const DevBuffers& b, const CarConst& cc, const uint8_t* mask, int backwards_flag,
                                int use_ego_color, int cls, const ScoreArgs& sa, int wait_post, cudaStream_t stream) {
    if (d.N <= PJ_SMALL_FRAMES) {
        const int score_ctas = sa.enabled ? (d.N + PJ_W_SMALL - 1) / PJ_W_SMALL : 0;
        mcr_launch_pdl(project_kernel<PJ_W_SMALL>, dim3(d.N + score_ctas), dim3(PJ_W_SMALL * 32), 0, stream,
                       d, b, cc, mask, backwards_flag, use_ego_color, cls, sa, wait_post);
    } else {
        const int score_ctas = sa.enabled ? d.N : 0;
        mcr_launch_pdl(project_kernel<1>, dim3(d.N + score_ctas), dim3(32), 0, stream,
                       d, b, cc, mask, backwards_flag, use_ego_color, cls, sa, wait_post);
    }
}

int launch_render(const Dims& d, const DevBuffers& b, const CarConst& cc, const uint8_t* mask, uint8_t* obs,
                  int backwards_flag, int use_ego_color, int cls, int obs_format, int stack_k, void* stream,
                  const uint8_t* score_noact, double* score_reward, uint8_t* score_done, int max_episode_steps, int wait_post) {
    if (!configure_render()) return -1;
    // score_reward != NULL: the step's reward / done block runs inside this launch (project_kernel's extra CTAs); callers
    // must check render_runs_score(cls) first -- the fused kernel has no such CTAs
    const ScoreArgs sa{score_noact, score_reward, score_done, max_episode_steps, score_reward != nullptr ? 1 : 0};
    static const bool fused = std::getenv("MCR_RENDER_FUSED") != nullptr;     // diagnostics: the single-kernel rasteriser (A/B)
    // cls 2 = the few envs with touching cars, at the end of the step's longest chain: one wave of the fused kernel
    // (~18 us) ends earlier there than project + fill (two latency-bound launches, ~30 us for a handful of frames)
    if (fused || cls == 2) {
        mcr_launch_pdl(render_kernel<false>, dim3(d.B, d.A), dim3(RS_THREADS), sizeof(RasterSmem), (cudaStream_t)stream,
                       d, b, cc, mask, obs, backwards_flag, use_ego_color, cls, obs_format, stack_k, VpParams{});
        return cudaGetLastError() == cudaSuccess ? 1 : -1;
    }
    launch_project_impl(d, b, cc, mask, backwards_flag, use_ego_color, cls, sa, wait_post, (cudaStream_t)stream);
    mcr_launch_pdl(fill_kernel, dim3(d.B, d.A), dim3(RS_THREADS), sizeof(RasterSmem), (cudaStream_t)stream,
                   d, b, mask, obs, cls, obs_format, stack_k, 0, (wait_post && d.N <= PJ_ONE_WAVE) ? 1 : 0);
    return cudaGetLastError() == cudaSuccess ? 2 : -1;
}

bool render_is_split() {
    static const bool fused = std::getenv("MCR_RENDER_FUSED") != nullptr;
    return !fused;
}

// The two halves of launch_render for callers that interleave something with the fill (mcr_step_host: the device-to-host
// copy of a range of envs starts as soon as that range is filled).
// MCR_SCORE_IN_PROJECT=1 (diagnostics): the reward / done block as extra CTAs of project_kernel instead of its own kernel
// on the side stream.  Measured slower (profiles/README r02): the projector is on the critical chain and gets 3 us longer,
// and the gap it was meant to close (post_kernel -> projector) is the tail of post_kernel's slowest cars, not launch latency.
bool render_runs_score(int cls) {
    static const bool on = std::getenv("MCR_SCORE_IN_PROJECT") != nullptr;
    return on && render_is_split() && cls != 2;
}

int launch_project(const Dims& d, const DevBuffers& b, const CarConst& cc, const uint8_t* mask, int backwards_flag, int use_ego_color,
                   int cls, void* stream, const uint8_t* score_noact, double* score_reward, uint8_t* score_done, int max_episode_steps,
                   int wait_post) {
    if (!configure_render()) return -1;
    const ScoreArgs sa{score_noact, score_reward, score_done, max_episode_steps, score_reward != nullptr ? 1 : 0};
    launch_project_impl(d, b, cc, mask, backwards_flag, use_ego_color, cls, sa, wait_post, (cudaStream_t)stream);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// wait_flags: project_kernel was launched with wait_post (the step's chain) -- see fill_kernel
int launch_fill(const Dims& d, const DevBuffers& b, const uint8_t* mask, uint8_t* obs, int cls, int obs_format, int stack_k,
                int env0, int nenv, bool pdl, void* stream, int wait_flags) {
    if (!configure_render()) return -1;
    if (pdl) mcr_launch_pdl(fill_kernel, dim3(nenv, d.A), dim3(RS_THREADS), sizeof(RasterSmem), (cudaStream_t)stream,
                            d, b, mask, obs, cls, obs_format, stack_k, env0, (wait_flags && d.N <= PJ_ONE_WAVE) ? 1 : 0);
    else fill_kernel<<<dim3(nenv, d.A), RS_THREADS, sizeof(RasterSmem), (cudaStream_t)stream>>>(d, b, mask, obs, cls, obs_format, stack_k, env0, (wait_flags && d.N <= PJ_ONE_WAVE) ? 1 : 0);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// camera of mcr:540-556 for a viewport of vw x vh pixels (post_kernel evaluates the 96 x 96 one); same
// float64 expression order, time and poses as they are when render() is called
__global__ void camera_kernel(Dims d, DevBuffers b, const uint8_t* __restrict__ mask, float* __restrict__ cam, double h_ratio, int vw, int vh) {
    const int car = blockIdx.x * blockDim.x + threadIdx.x;
    if (car >= d.N) return;
    if (mask && !mask[car / d.A]) return;
    const int N = d.N;
    const double SCALE = 6.0, ZOOM = 2.7, WINDOW_W = 1000, WINDOW_H = 800;
    const double t = b.time[car];
    const double zoom = 0.1 * SCALE * fmax(1 - t, 0.0) + ZOOM * SCALE * fmin(t, 1.0);
    const double scroll_x = b.body[(size_t)BF_PX * N + car], scroll_y = b.body[(size_t)BF_PY * N + car];
    double angle = -(double)b.body[(size_t)BF_A * N + car];
    const double hvx = b.body[(size_t)BF_VX * N + car], hvy = b.body[(size_t)BF_VY * N + car];
    if (sqrt(hvx * hvx + hvy * hvy) > 0.5) angle = atan2(hvx, hvy);
    const double tx = WINDOW_W / 2 - (scroll_x * zoom * cos(angle) - scroll_y * zoom * sin(angle));
    const double ty = WINDOW_H * h_ratio - (scroll_x * zoom * sin(angle) + scroll_y * zoom * cos(angle));
    const float ftx = (float)tx, fty = (float)ty, fdeg = (float)(57.29577951308232 * angle), fzoom = (float)zoom;
    const double rad = (double)fdeg * (3.14159265358979323846 / 180.0);
    const double cs = cos(rad), sn = sin(rad);
    const double SX = (double)vw / 1000.0, SY = (double)vh / 800.0;
    cam[(size_t)0 * N + car] = (float)(cs * (double)fzoom * SX);
    cam[(size_t)1 * N + car] = (float)(-sn * (double)fzoom * SX);
    cam[(size_t)2 * N + car] = (float)((double)ftx * SX);
    cam[(size_t)3 * N + car] = (float)(sn * (double)fzoom * SY);
    cam[(size_t)4 * N + car] = (float)(cs * (double)fzoom * SY);
    cam[(size_t)5 * N + car] = (float)((double)fty * SY);
}

int launch_render_viewport(const Dims& d, const DevBuffers& b, const CarConst& cc, const uint8_t* mask, uint8_t* out, float* cam,
                           int vw, int vh, double h_ratio, int backwards_flag, int use_ego_color, void* stream) {
    if (!configure_render()) return -1;
    camera_kernel<<<(d.N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d, b, mask, cam, h_ratio, vw, vh);
    VpParams vp;
    vp.vw = vw; vp.vh = vh; vp.tiles_x = (vw + SW - 1) / SW;
    vp.hud_sx = (float)((double)vw / 1000.0); vp.hud_sy = (float)((double)vh / 800.0); vp.camera = cam;
    const int tiles = vp.tiles_x * ((vh + SH - 1) / SH);
    render_kernel<true><<<dim3(d.N, tiles), RS_THREADS, sizeof(RasterSmem), (cudaStream_t)stream>>>(
        d, b, cc, mask, out, backwards_flag, use_ego_color, 0, MCR_OBS_RGB_HWC, 1, vp);
    return cudaGetLastError() == cudaSuccess ? 2 : -1;
}

int launch_score(const Dims& d, const DevBuffers& b, const uint8_t* mask, const uint8_t* noact, double* reward, uint8_t* done,
                 int max_episode_steps, int cls, void* stream) {
    // score_kernel runs beside render_kernel.  An SM can only change its shared-memory carve-out when it is
    // idle, so score CTAs that were placed first (small carve-out) held render_kernel's CTAs back by ~7 us.
    // Asking for 12 KB of (unused) dynamic shared memory per CTA -- 16 resident CTAs x 12 KB lands in the same
    // carve-out class as the rasteriser's 4 x 37 KB -- makes both kernels want the same configuration
    // (measured: render start 98.4 -> 91.3 us into the step; 0 and 9.5 KB do not).  With the programmatic
    // dependent launch of render_kernel the rasteriser's CTAs are placed first and score fills in as they retire.
    const size_t score_smem = 12288;
    score_kernel<<<(d.N + SCORE_WARPS - 1) / SCORE_WARPS, SCORE_WARPS * 32, score_smem, (cudaStream_t)stream>>>(d, b, mask, noact, reward, done, max_episode_steps, cls);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}
