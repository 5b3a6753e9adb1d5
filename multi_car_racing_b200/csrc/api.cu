// api.cu -- host side of libmcr.so: the C-ABI of include/mcr.h.
//
// Host restatements (product code, independent of oracle/):
//   * Box2D b2PolygonShape::Set / ComputeMass and b2Body::ResetMassData for the fixed
//     car_dynamics.Car geometry (gym 0.17.2) -> CarConst
//   * numpy legacy RandomState (MT19937) + MultiCarRacing._create_track (mcr:183-338) and the
//     reset() spawn grid (mcr:366-393), float64 with the reference's operation order
// "mcr" = gym_multi_car_racing/multi_car_racing.py of the reference.
#include "mcr_internal.h"
#include <cuda_runtime.h>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <algorithm>
#include <string>
#include <vector>

static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
    return code;
}
#define CUDA_OK(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return fail(-100, "%s: %s", #expr, cudaGetErrorString(e_)); } while (0)

// ---------------------------------------------------------------------------------------
// Box2D polygon set-up (fp32)
// ---------------------------------------------------------------------------------------
namespace {

struct P2 { float x, y; };
inline P2 operator-(P2 a, P2 b) { return P2{a.x - b.x, a.y - b.y}; }
inline P2 operator+(P2 a, P2 b) { return P2{a.x + b.x, a.y + b.y}; }
inline P2 operator*(float s, P2 a) { return P2{s * a.x, s * a.y}; }
inline float dot(P2 a, P2 b) { return a.x * b.x + a.y * b.y; }
inline float cross(P2 a, P2 b) { return a.x * b.y - a.y * b.x; }

// b2PolygonShape::Set: weld, gift-wrap from the right-most vertex, counter-clockwise.
std::vector<P2> b2_polygon_set(const std::vector<P2>& in) {
    std::vector<P2> ps;
    for (const P2& v : in) {
        bool uniq = true;
        for (const P2& u : ps) { P2 dd = v - u; if (dot(dd, dd) < 0.5f * B2_LINEAR_SLOP) { uniq = false; break; } }
        if (uniq) ps.push_back(v);
    }
    const int n = (int)ps.size();
    if (n < 3) return {};
    int i0 = 0; float x0 = ps[0].x;
    for (int i = 1; i < n; ++i) {
        const float x = ps[i].x;
        if (x > x0 || (x == x0 && ps[i].y < ps[i0].y)) { i0 = i; x0 = x; }
    }
    std::vector<int> hull; int ih = i0;
    for (;;) {
        hull.push_back(ih);
        int ie = 0;
        for (int j = 1; j < n; ++j) {
            if (ie == ih) { ie = j; continue; }
            const P2 r = ps[ie] - ps[hull.back()], v = ps[j] - ps[hull.back()];
            const float c = cross(r, v);
            if (c < 0.0f) ie = j;
            if (c == 0.0f && dot(v, v) > dot(r, r)) ie = j;
        }
        ih = ie;
        if (ie == i0) break;
        if ((int)hull.size() > n) return {};
    }
    std::vector<P2> out;
    for (int id : hull) out.push_back(ps[id]);
    return out;
}

struct Mass { float mass; P2 center; float I; };

// b2PolygonShape::ComputeMass
Mass b2_polygon_mass(const std::vector<P2>& v, float density) {
    const int n = (int)v.size();
    P2 center{0.0f, 0.0f}; float area = 0.0f, I = 0.0f;
    P2 s{0.0f, 0.0f};
    for (int i = 0; i < n; ++i) s = s + v[i];
    s = (1.0f / n) * s;
    const float k_inv3 = 1.0f / 3.0f;
    for (int i = 0; i < n; ++i) {
        const P2 e1 = v[i] - s, e2 = (i + 1 < n ? v[i + 1] : v[0]) - s;
        const float D = cross(e1, e2);
        const float tri = 0.5f * D;
        area += tri;
        center = center + (tri * k_inv3) * (e1 + e2);
        const float intx2 = e1.x * e1.x + e2.x * e1.x + e2.x * e2.x;
        const float inty2 = e1.y * e1.y + e2.y * e1.y + e2.y * e2.y;
        I += (0.25f * k_inv3 * D) * (intx2 + inty2);
    }
    Mass m;
    m.mass = density * area;
    center = (1.0f / area) * center;
    m.center = center + s;
    m.I = density * I;
    m.I += m.mass * (dot(m.center, m.center) - dot(center, center));
    return m;
}

void to_poly8(const std::vector<P2>& v, Poly8& out) {
    out.n = (int)v.size();
    for (int i = 0; i < MCR_MAXV; ++i) { out.x[i] = i < out.n ? v[i].x : 0.0f; out.y[i] = i < out.n ? v[i].y : 0.0f; out.nx[i] = 0.0f; out.ny[i] = 0.0f; }
    // b2PolygonShape::Set: m_normals[i] = b2Cross(edge, 1.0f), normalised
    for (int i = 0; i < out.n; ++i) {
        const int i2 = i + 1 < out.n ? i + 1 : 0;
        const P2 e = v[i2] - v[i];
        P2 nn{1.0f * e.y, -1.0f * e.x};
        const float len = std::sqrt(nn.x * nn.x + nn.y * nn.y);
        if (len >= B2_EPS) { const float inv = 1.0f / len; nn.x *= inv; nn.y *= inv; }
        out.nx[i] = nn.x; out.ny[i] = nn.y;
    }
}

// car_dynamics.py geometry (gym 0.17.2), units of SIZE
const double kSIZE = 0.02;
const double kWHEEL_R = 27, kWHEEL_W = 14;
const double kWHEELPOS[4][2] = {{-55, +80}, {+55, +80}, {-55, -82}, {+55, -82}};
const std::vector<std::vector<std::pair<double, double>>> kHULL = {
    {{-60, +130}, {+60, +130}, {+60, +110}, {-60, +110}},
    {{-15, +120}, {+15, +120}, {+20, +20}, {-20, 20}},
    {{+25, +20}, {+50, -10}, {+50, -40}, {+20, -90}, {-20, -90}, {-50, -40}, {-50, -10}, {-25, +20}},
    {{-50, -120}, {+50, -120}, {+50, -90}, {-50, -90}}};

bool build_car_const(CarConst& cc) {
    std::vector<P2> hull[4];
    for (int f = 0; f < 4; ++f) {
        std::vector<P2> raw;
        for (auto& p : kHULL[f]) raw.push_back(P2{(float)(p.first * kSIZE), (float)(p.second * kSIZE)});
        hull[f] = b2_polygon_set(raw);
        if (hull[f].size() != kHULL[f].size()) return false;
        to_poly8(hull[f], cc.hull_poly[f]);
    }
    const double front_k = 1.0;
    const double wp[4][2] = {{-kWHEEL_W, +kWHEEL_R}, {+kWHEEL_W, +kWHEEL_R}, {+kWHEEL_W, -kWHEEL_R}, {-kWHEEL_W, -kWHEEL_R}};
    std::vector<P2> raw;
    for (int i = 0; i < 4; ++i) raw.push_back(P2{(float)(wp[i][0] * front_k * kSIZE), (float)(wp[i][1] * front_k * kSIZE)});
    std::vector<P2> wheel = b2_polygon_set(raw);
    if (wheel.size() != 4) return false;
    to_poly8(wheel, cc.wheel_poly);
    // b2Body::ResetMassData -- the fixture list is walked newest first
    {
        float mass = 0.0f, I = 0.0f; P2 lc{0.0f, 0.0f};
        for (int f = 3; f >= 0; --f) {
            const Mass md = b2_polygon_mass(hull[f], 1.0f);
            mass += md.mass; lc = lc + md.mass * md.center; I += md.I;
        }
        const float inv = 1.0f / mass;
        lc = inv * lc;
        I -= mass * dot(lc, lc);
        cc.hull_mass = mass; cc.hull_invMass = inv; cc.hull_I = I; cc.hull_invI = 1.0f / I;
        cc.hull_lcx = lc.x; cc.hull_lcy = lc.y;
    }
    {
        const Mass md = b2_polygon_mass(wheel, 0.1f);
        const float mass = md.mass; P2 lc = md.mass * md.center; float I = md.I;
        const float inv = 1.0f / mass;
        lc = inv * lc;
        I -= mass * dot(lc, lc);
        cc.wheel_mass = mass; cc.wheel_invMass = inv; cc.wheel_I = I; cc.wheel_invI = 1.0f / I;
        // the solver drops every rB term: that is only exact if the wheel's centre of mass is its origin
        if (lc.x != 0.0f || lc.y != 0.0f) return false;
    }
    for (int w = 0; w < 4; ++w) { cc.anchor_x[w] = (float)(kWHEELPOS[w][0] * kSIZE); cc.anchor_y[w] = (float)(kWHEELPOS[w][1] * kSIZE); }
    {   // the rectangle of head_kernel's narrow-phase prefilter must hold the car with its margin (a wheel turns about its anchor)
        float ex = 0.0f, ey = 0.0f;
        for (int f = 0; f < 4; ++f) for (auto& v : hull[f]) { ex = std::max(ex, std::fabs(v.x)); ey = std::max(ey, std::fabs(v.y)); }
        const float wheel_diag = (float)(std::sqrt(kWHEEL_R * kWHEEL_R + kWHEEL_W * kWHEEL_W) * kSIZE);
        for (int w = 0; w < 4; ++w) { ex = std::max(ex, std::fabs(cc.anchor_x[w]) + wheel_diag); ey = std::max(ey, std::fabs(cc.anchor_y[w]) + wheel_diag); }
        if (ex + 0.45f > CAR_OBB_EX || ey + 0.45f > CAR_OBB_EY) return false;
    }
    cc.max_motor_torque = (float)(180 * 900 * kSIZE * kSIZE);
    cc.lower = (float)-0.4; cc.upper = (float)+0.4;
    // the sweep kernels have no e_equalLimits path (b2RevoluteJoint: |upper - lower| < 2 * angularSlop)
    if (std::fabs(cc.upper - cc.lower) < 2.0f * B2_ANGULAR_SLOP) return false;
    return true;
}

// ---------------------------------------------------------------------------------------
// numpy legacy RandomState: MT19937
// ---------------------------------------------------------------------------------------
struct MT {
    uint32_t* mt; uint32_t& pos;
    explicit MT(uint32_t* st) : mt(st), pos(st[624]) {}
    void regen() {
        const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MATRIX_A = 0x9908b0dfu;
        int kk; uint32_t y;
        for (kk = 0; kk < 624 - 397; ++kk) { y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER); mt[kk] = mt[kk + 397] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A); }
        for (; kk < 623; ++kk) { y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER); mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A); }
        y = (mt[623] & UPPER) | (mt[0] & LOWER); mt[623] = mt[396] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A);
        pos = 0;
    }
    uint32_t next32() {
        if (pos >= 624) regen();
        uint32_t y = mt[pos++];
        y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= (y >> 18);
        return y;
    }
    double next_double() { const uint32_t a = next32() >> 5, b = next32() >> 6; return (a * 67108864.0 + b) / 9007199254740992.0; }
    double uniform(double lo, double hi) { return lo + (hi - lo) * next_double(); }
};

inline double npsign(double x) { return x > 0 ? 1.0 : (x < 0 ? -1.0 : 0.0); }
inline int pyidx(int i, int n) { return i < 0 ? i + n : i; }

}  // namespace

extern "C" int mcr_mt_seed(uint32_t* st, const uint32_t* key, int32_t key_len) {
    if (!st || !key || key_len < 1) return fail(-1, "mcr_mt_seed: bad arguments");
    // init_genrand(19650218)
    st[0] = 19650218u;
    for (int i = 1; i < 624; ++i) st[i] = 1812433253u * (st[i - 1] ^ (st[i - 1] >> 30)) + (uint32_t)i;
    int i = 1, j = 0;
    int k = 624 > key_len ? 624 : key_len;
    for (; k; --k) {
        st[i] = (st[i] ^ ((st[i - 1] ^ (st[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
        ++i; ++j;
        if (i >= 624) { st[0] = st[623]; i = 1; }
        if (j >= key_len) j = 0;
    }
    for (k = 623; k; --k) {
        st[i] = (st[i] ^ ((st[i - 1] ^ (st[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
        ++i;
        if (i >= 624) { st[0] = st[623]; i = 1; }
    }
    st[0] = 0x80000000u;
    st[624] = 624;
    return 0;
}

extern "C" int mcr_mt_seed_batch(uint32_t* states, const uint32_t* keys, const int32_t* key_len, int32_t n, int32_t key_stride) {
    if (!states || !keys || !key_len || n < 1 || key_stride < 1) return fail(-1, "mcr_mt_seed_batch: bad arguments");
    for (int i = 0; i < n; ++i) {
        if (key_len[i] < 1 || key_len[i] > key_stride) return fail(-1, "mcr_mt_seed_batch: key %d has length %d", i, key_len[i]);
        const int rc = mcr_mt_seed(states + (size_t)i * 625, keys + (size_t)i * key_stride, key_len[i]);
        if (rc) return rc;
    }
    return 0;
}

// reset()'s draws from the GLOBAL numpy RandomState for n envs in a row (mcr:351-357), on a copy of its MT19937 state:
//   np.random.choice(['CW', 'CCW'])              -> randint(0, 2): one 32-bit draw, masked (legacy _rand_int64, rng = 1)
//   np.random.choice(ids, size=A, replace=False) -> permutation(A)[:A]: Fisher-Yates from the top, j = random_interval(i)
//                                                   (masked rejection on 32-bit draws)
extern "C" int mcr_reset_draws(uint32_t* mt_state, int32_t n, int32_t A, int32_t use_random_direction, int32_t default_cw,
                               uint8_t* h_cw, int32_t* h_order) {
    if (!mt_state || !h_cw || !h_order || n < 1 || A < 1) return fail(-1, "mcr_reset_draws: bad arguments");
    MT rng(mt_state);
    for (int e = 0; e < n; ++e) {
        int cw = default_cw ? 1 : 0;
        if (use_random_direction) cw = (rng.next32() & 1u) ? 0 : 1;        // index 0 = 'CW', 1 = 'CCW'
        h_cw[e] = (uint8_t)cw;
        int32_t* o = h_order + (size_t)e * A;
        for (int i = 0; i < A; ++i) o[i] = i;
        for (int i = A - 1; i > 0; --i) {
            uint32_t mask = (uint32_t)i;
            mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
            uint32_t j;
            do { j = rng.next32() & mask; } while (j > (uint32_t)i);
            const int32_t t = o[i]; o[i] = o[j]; o[j] = t;
        }
    }
    return 0;
}

// One attempt of _create_track (mcr:183-338).
extern "C" int mcr_track_generate(uint32_t* mt_state, int32_t max_tiles, int32_t max_quads, double* h_nodes,
                                  double* h_quads, float* h_quad_rgb, int32_t* h_quad_tile, int32_t* out_q,
                                  int32_t* idx_range) {
    if (!mt_state || !h_nodes || !h_quads || !h_quad_rgb || !h_quad_tile || !out_q) return fail(-1, "mcr_track_generate: null argument");
    MT rng(mt_state);
    const double PI = M_PI, SCALE = 6.0, TRACK_RAD = 900 / SCALE, TRACK_DETAIL_STEP = 21 / SCALE;
    const double TRACK_TURN_RATE = 0.31, TRACK_WIDTH = 40 / SCALE, BORDER = 8 / SCALE;
    const int CHECKPOINTS = 12, BORDER_MIN_COUNT = 4;
    struct CP { double alpha, x, y; } cp[CHECKPOINTS];
    double start_alpha = 0;
    for (int c = 0; c < CHECKPOINTS; ++c) {
        double alpha = 2 * PI * c / CHECKPOINTS + rng.uniform(0, 2 * PI * 1 / CHECKPOINTS);
        double rad = rng.uniform(TRACK_RAD / 3, TRACK_RAD);
        if (c == 0) { alpha = 0; rad = 1.5 * TRACK_RAD; }
        if (c == CHECKPOINTS - 1) { alpha = 2 * PI * c / CHECKPOINTS; start_alpha = 2 * PI * (-0.5) / CHECKPOINTS; rad = 1.5 * TRACK_RAD; }
        cp[c] = CP{alpha, rad * std::cos(alpha), rad * std::sin(alpha)};
    }
    struct Node { double alpha, beta, x, y; };
    std::vector<Node> path; path.reserve(2600);
    double x = 1.5 * TRACK_RAD, y = 0, beta = 0;
    long dest_i = 0; int laps = 0, no_freeze = 2500; bool visited_other_side = false;
    for (;;) {
        double alpha = std::atan2(y, x);
        if (visited_other_side && alpha > 0) { ++laps; visited_other_side = false; }
        if (alpha < 0) { visited_other_side = true; alpha += 2 * PI; }
        double dest_alpha, dest_x, dest_y;
        for (;;) {
            bool failed = true;
            for (;;) {
                const CP& d = cp[dest_i % CHECKPOINTS];
                dest_alpha = d.alpha; dest_x = d.x; dest_y = d.y;
                if (alpha <= dest_alpha) { failed = false; break; }
                ++dest_i;
                if (dest_i % CHECKPOINTS == 0) break;
            }
            if (!failed) break;
            alpha -= 2 * PI;
        }
        const double r1x = std::cos(beta), r1y = std::sin(beta);
        const double p1x = -r1y, p1y = r1x;
        const double dest_dx = dest_x - x, dest_dy = dest_y - y;
        double proj = r1x * dest_dx + r1y * dest_dy;
        while (beta - alpha > 1.5 * PI) beta -= 2 * PI;
        while (beta - alpha < -1.5 * PI) beta += 2 * PI;
        const double prev_beta = beta;
        proj *= SCALE;
        if (proj > 0.3) beta -= std::fmin(TRACK_TURN_RATE, std::fabs(0.001 * proj));
        if (proj < -0.3) beta += std::fmin(TRACK_TURN_RATE, std::fabs(0.001 * proj));
        x += p1x * TRACK_DETAIL_STEP;
        y += p1y * TRACK_DETAIL_STEP;
        path.push_back(Node{alpha, prev_beta * 0.5 + beta * 0.5, x, y});
        if (laps > 4) break;
        if (--no_freeze == 0) break;
    }
    int i1 = -1, i2 = -1;
    int i = (int)path.size();
    for (;;) {
        --i;
        if (i == 0) return 0;   // "return False  # Failed"
        const bool pass = path[i].alpha > start_alpha && path[i - 1].alpha <= start_alpha;
        if (pass && i2 == -1) i2 = i;
        else if (pass && i1 == -1) { i1 = i; break; }
    }
    if (idx_range) { idx_range[0] = i1; idx_range[1] = i2; }
    const int n = (i2 - 1) - i1;
    if (n <= 0) return 0;
    const Node* tr = path.data() + i1;
    {
        const double fb = tr[0].beta, fpx = std::cos(fb), fpy = std::sin(fb);
        const double a = fpx * (tr[0].x - tr[n - 1].x), bq = fpy * (tr[0].y - tr[n - 1].y);
        const double glued = std::sqrt(a * a + bq * bq);
        if (glued > TRACK_DETAIL_STEP) return 0;
    }
    if (n > max_tiles) return fail(-2, "track has %d tiles, max_tiles is %d", n, max_tiles);
    std::vector<char> border(n, 0);
    for (int k = 0; k < n; ++k) {
        bool good = true; double oneside = 0;
        for (int neg = 0; neg < BORDER_MIN_COUNT; ++neg) {
            const double b1 = tr[pyidx(k - neg - 0, n)].beta, b2 = tr[pyidx(k - neg - 1, n)].beta;
            good &= std::fabs(b1 - b2) > TRACK_TURN_RATE * 0.2;
            oneside += npsign(b1 - b2);
        }
        good &= std::fabs(oneside) == BORDER_MIN_COUNT;
        border[k] = good;
    }
    for (int k = 0; k < n; ++k)
        for (int neg = 0; neg < BORDER_MIN_COUNT; ++neg) border[pyidx(k - neg, n)] |= border[k];
    int q = 0;
    auto put = [&](const double v[8], float r, float g, float bl, int tile) -> bool {
        if (q >= max_quads) return false;
        std::memcpy(h_quads + (size_t)q * 8, v, sizeof(double) * 8);
        h_quad_rgb[3 * q] = r; h_quad_rgb[3 * q + 1] = g; h_quad_rgb[3 * q + 2] = bl;
        h_quad_tile[q] = tile; ++q; return true;
    };
    for (int k = 0; k < n; ++k) {
        const Node& n1 = tr[k]; const Node& n2 = tr[pyidx(k - 1, n)];
        const double b1 = n1.beta, x1 = n1.x, y1 = n1.y, b2 = n2.beta, x2 = n2.x, y2 = n2.y;
        const double road[8] = {x1 - TRACK_WIDTH * std::cos(b1), y1 - TRACK_WIDTH * std::sin(b1),
                                x1 + TRACK_WIDTH * std::cos(b1), y1 + TRACK_WIDTH * std::sin(b1),
                                x2 + TRACK_WIDTH * std::cos(b2), y2 + TRACK_WIDTH * std::sin(b2),
                                x2 - TRACK_WIDTH * std::cos(b2), y2 - TRACK_WIDTH * std::sin(b2)};
        const double c = 0.01 * (k % 3);
        {   // D8: a tile quad that b2PolygonShape::Set cannot keep as four vertices (fd_tile.shape.vertices = ..., mcr:318:
            // the reference's CreateStaticBody asserts) fails the attempt, like the device generator does (trackgen.cuh)
            std::vector<P2> raw;
            for (int v = 0; v < 4; ++v) raw.push_back(P2{(float)road[2 * v], (float)road[2 * v + 1]});
            if (b2_polygon_set(raw).size() != 4) return 0;
        }
        if (!put(road, (float)(0.4 + c), (float)(0.4 + c), (float)(0.4 + c), k)) return fail(-3, "max_quads too small");
        if (border[k]) {
            const double side = npsign(b2 - b1);
            const double bq[8] = {x1 + side * TRACK_WIDTH * std::cos(b1), y1 + side * TRACK_WIDTH * std::sin(b1),
                                  x1 + side * (TRACK_WIDTH + BORDER) * std::cos(b1), y1 + side * (TRACK_WIDTH + BORDER) * std::sin(b1),
                                  x2 + side * (TRACK_WIDTH + BORDER) * std::cos(b2), y2 + side * (TRACK_WIDTH + BORDER) * std::sin(b2),
                                  x2 + side * TRACK_WIDTH * std::cos(b2), y2 + side * TRACK_WIDTH * std::sin(b2)};
            const bool white = k % 2 == 0;
            if (!put(bq, 1.0f, white ? 1.0f : 0.0f, white ? 1.0f : 0.0f, -1)) return fail(-3, "max_quads too small");
        }
    }
    for (int k = 0; k < n; ++k) { h_nodes[4 * k] = tr[k].alpha; h_nodes[4 * k + 1] = tr[k].beta; h_nodes[4 * k + 2] = tr[k].x; h_nodes[4 * k + 3] = tr[k].y; }
    *out_q = q;
    return n;
}

// reset() spawn grid, mcr:366-393
extern "C" int mcr_spawn_poses(const double* nodes, int32_t T, const int32_t* car_order, int32_t A, int32_t cw, double* poses) {
    if (!nodes || !car_order || !poses || T < 1 || A < 1) return fail(-1, "mcr_spawn_poses: bad arguments");
    const double PI = M_PI;
    const int LINE_SPACING = 5; const double LATERAL_SPACING = 3;
    const double pos_x = nodes[2], pos_y = nodes[3];
    for (int c = 0; c < A; ++c) {
        const int line_number = car_order[c] / 2;
        const int side = 2 * (car_order[c] % 2) - 1;
        int idx = -line_number * LINE_SPACING;
        if (idx < 0) idx += T;
        if (idx < 0 || idx >= T) return fail(-2, "mcr_spawn_poses: track too short for %d agents", A);
        const double dx = nodes[4 * idx + 2] - pos_x, dy = nodes[4 * idx + 3] - pos_y;
        double angle = nodes[4 * idx + 1];
        if (cw) angle -= PI;
        const double norm_theta = angle - PI / 2;
        poses[3 * c + 0] = angle;
        poses[3 * c + 1] = pos_x + dx + (LATERAL_SPACING * std::sin(norm_theta) * side);
        poses[3 * c + 2] = pos_y + dy + (LATERAL_SPACING * std::cos(norm_theta) * side);
    }
    return 0;
}

// ---------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------
struct BufSpec { const char* name; int dtype; int ndim; int64_t dims[4]; };

struct mcr_handle_t {
    mcr_config cfg;
    Dims d;
    CarConst cc;
    DevBuffers buf;
    void* ptr[BUF_COUNT];
    BufSpec spec[BUF_COUNT];
    int64_t launches;
    uint8_t palette[PAL_COUNT][4];
    bool palette_ready;
    cudaStream_t side;           // contacts run here, beside the solver
    cudaStream_t cap;            // capture origin for the step graph (the caller's stream may be the legacy default stream, which cannot be captured)
    cudaEvent_t ev_fork, ev_join;
    cudaStream_t side2;          // the chain of envs with touching cars: coupled -> post -> score -> render
    cudaStream_t side3;          // score_kernel of the touching-car envs, beside their rasteriser
    cudaStream_t copy;           // mcr_step_host: device-to-host copies of the envs already filled, beside the fill of the rest
    cudaEvent_t ev_chunk[8], ev_copy;
    cudaStream_t refill;         // lowest priority: ring_refill_kernel tops up the envs' track rings beside the steps
    cudaEvent_t ev_refill, ev_refill_go;
    int64_t steps_since_refill;
    cudaEvent_t ev_pre, ev_contacts, ev_chain2, ev_score, ev_post2, ev_score2;
    bool side_ready;
    // mcr_step replays a captured CUDA graph of its launches (one graph per argument tuple)
    struct StepGraph { int32_t dtype, flags; uint8_t* obs; double* reward; uint8_t* done; uint8_t* h_obs; double* h_reward; uint8_t* h_done;
                       cudaGraphExec_t exec; int64_t launches;
                       // the kernel nodes that read the step's action: their pointer argument is patched per step
                       // (cudaGraphExecKernelNodeSetParams) instead of copying the action into action_stage first -- the
                       // copy was 3.5 us of every step (profiles/README r02).  `graph` stays alive: the nodes' argument
                       // storage belongs to it.
                       cudaGraph_t graph; std::vector<std::pair<cudaGraphNode_t, int>> action_nodes; const void* cur_action; };
    std::vector<StepGraph> graphs;
    int64_t eager_steps;
    bool use_graphs;
    int obs_format;              // MCR_OBS_*
    int stack_k;                 // ring depth of MCR_OBS_GRAY_STACK
};


static void destroy_step_graph(mcr_handle_t::StepGraph& g) {
    if (g.exec) cudaGraphExecDestroy(g.exec);
    if (g.graph) cudaGraphDestroy(g.graph);
    g.exec = nullptr; g.graph = nullptr; g.action_nodes.clear();
}

// The kernel nodes of a captured step that take `stage` as their action argument.
static void find_action_nodes(cudaGraph_t graph, const void* stage, std::vector<std::pair<cudaGraphNode_t, int>>& out) {
    size_t n = 0;
    if (cudaGraphGetNodes(graph, nullptr, &n) != cudaSuccess || n == 0) { (void)cudaGetLastError(); return; }
    std::vector<cudaGraphNode_t> nodes(n);
    if (cudaGraphGetNodes(graph, nodes.data(), &n) != cudaSuccess) { (void)cudaGetLastError(); return; }
    for (size_t i = 0; i < n; ++i) {
        cudaGraphNodeType ty;
        if (cudaGraphNodeGetType(nodes[i], &ty) != cudaSuccess || ty != cudaGraphNodeTypeKernel) continue;
        cudaKernelNodeParams kp = {};
        if (cudaGraphKernelNodeGetParams(nodes[i], &kp) != cudaSuccess || !kp.kernelParams) { (void)cudaGetLastError(); continue; }
        int nargs = 0, idx = head_action_arg(kp.func, &nargs);
        if (idx < 0) idx = pre_action_arg(kp.func, &nargs);
        if (idx < 0) continue;
        if (*reinterpret_cast<const void* const*>(kp.kernelParams[idx]) == stage) out.emplace_back(nodes[i], idx);
    }
}

// Point the action argument of the step graph's kernels at `action` (no-op when it already is).
static int patch_action(mcr_handle_t::StepGraph& g, const void* action) {
    if (g.cur_action == action) return 0;
    for (auto& na : g.action_nodes) {
        cudaKernelNodeParams kp = {};
        CUDA_OK(cudaGraphKernelNodeGetParams(na.first, &kp));
        int nargs = 0;
        if (head_action_arg(kp.func, &nargs) < 0) (void)pre_action_arg(kp.func, &nargs);
        void* args[16];
        for (int i = 0; i < nargs && i < 16; ++i) args[i] = kp.kernelParams[i];
        const void* a = action;
        args[na.second] = (void*)&a;
        kp.kernelParams = args; kp.extra = nullptr;
        CUDA_OK(cudaGraphExecKernelNodeSetParams(g.exec, na.first, &kp));
    }
    g.cur_action = action;
    return 0;
}

static void set_spec(mcr_handle_t* h, int id, const char* name, int dtype, std::initializer_list<int64_t> dims) {
    BufSpec& s = h->spec[id];
    s.name = name; s.dtype = dtype; s.ndim = (int)dims.size();
    int k = 0; for (int64_t v : dims) s.dims[k++] = v;
    for (; k < 4; ++k) s.dims[k] = 1;
}

extern "C" int mcr_abi_version(void) { return MCR_ABI_VERSION; }
extern "C" const char* mcr_last_error(void) { return g_err; }

extern "C" int mcr_create(const mcr_config* cfg, mcr_handle* out) {
    if (!cfg || !out) return fail(-1, "mcr_create: null argument");
    if (cfg->batch_envs < 1) return fail(-1, "batch_envs must be >= 1");
    if (cfg->num_agents < 1 || cfg->num_agents > MCR_MAX_AGENTS) return fail(-1, "num_agents must be in [1, %d]", MCR_MAX_AGENTS);
    if (cfg->max_tiles < 16 || cfg->max_tiles > 32768) return fail(-1, "max_tiles out of range");
    if (cfg->max_quads < cfg->max_tiles || cfg->max_quads > 2048 || cfg->max_quads % MCR_QUAD_CHUNK)
        return fail(-1, "max_quads must be a multiple of %d in [max_tiles, 2048]", MCR_QUAD_CHUNK);
    if (cfg->pool_tracks < 1) return fail(-1, "pool_tracks must be >= 1");
    if (cfg->fresh_tracks < 0 || cfg->fresh_tracks > 7) return fail(-1, "fresh_tracks (spare tracks per env) must be in [0, 7]");
    if (cfg->fresh_tracks > 0 && (int64_t)cfg->pool_tracks != (int64_t)cfg->batch_envs * (cfg->fresh_tracks + 1))
        return fail(-1, "fresh_tracks = R needs pool_tracks == batch_envs * (R + 1): env e owns slots e + batch_envs * j");
    mcr_handle_t* h = new mcr_handle_t();
    h->cfg = *cfg;
    h->d = Dims{cfg->batch_envs, cfg->num_agents, cfg->batch_envs * cfg->num_agents, cfg->max_tiles, cfg->max_quads, cfg->pool_tracks, cfg->particles ? 1 : 0,
                MCR_DL_CAP(cfg->max_quads, cfg->num_agents)};
    if (!build_car_const(h->cc)) { delete h; return fail(-2, "car geometry set-up failed"); }
    std::memset(&h->buf, 0, sizeof(h->buf));
    std::memset(h->ptr, 0, sizeof(h->ptr));
    h->launches = 0; h->palette_ready = false; h->side_ready = false;
    h->obs_format = MCR_OBS_RGB_HWC; h->stack_k = 4;
    h->steps_since_refill = 0;
    h->eager_steps = 0; h->use_graphs = std::getenv("MCR_NO_GRAPH") == nullptr;
    const int64_t N = h->d.N, B = h->d.B, A = h->d.A, T = h->d.Tmax, Q = h->d.Qmax, P = h->d.P;
    set_spec(h, BUF_BODY, "body", MCR_F32, {5, BODY_FIELDS, N});
    set_spec(h, BUF_SLEEP_TIME, "sleep_time", MCR_F32, {5, N});
    set_spec(h, BUF_AWAKE, "awake", MCR_U8, {5, N});
    set_spec(h, BUF_JOINT, "joint", MCR_F32, {4, JOINT_FIELDS, N});
    set_spec(h, BUF_LIMIT_STATE, "limit_state", MCR_U8, {4, N});
    set_spec(h, BUF_WHEEL, "wheel", MCR_F64, {4, WHEEL_FIELDS, N});
    set_spec(h, BUF_CTRL, "ctrl", MCR_F64, {CTRL_FIELDS, N});
    set_spec(h, BUF_ON_ROAD, "on_road", MCR_U8, {4, N});
    set_spec(h, BUF_ON_ROAD_NEXT, "on_road_next", MCR_U8, {4, N});
    set_spec(h, BUF_REWARD, "reward", MCR_F64, {N});
    set_spec(h, BUF_PREV_REWARD, "prev_reward", MCR_F64, {N});
    set_spec(h, BUF_VISIT_COUNT, "visit_count", MCR_I32, {N});
    set_spec(h, BUF_BACKWARD, "backward", MCR_U8, {N});
    set_spec(h, BUF_TIME, "time", MCR_F64, {N});
    set_spec(h, BUF_STEPS, "steps", MCR_I32, {N});
    set_spec(h, BUF_CAMERA, "camera", MCR_F32, {6, N});
    set_spec(h, BUF_STRIPE, "stripe", MCR_F32, {8, N});
    set_spec(h, BUF_HEADING, "heading", MCR_F64, {N});
    set_spec(h, BUF_ENV_TRACK, "env_track", MCR_I32, {B});
    set_spec(h, BUF_ENV_CW, "env_cw", MCR_U8, {B});
    set_spec(h, BUF_ENV_EPISODE, "env_episode", MCR_U32, {B});
    set_spec(h, BUF_VISITED, "visited", MCR_U32, {B, T});
    set_spec(h, BUF_TOUCHED, "touched", MCR_U8, {B, T});
    set_spec(h, BUF_RESET_MASK, "reset_mask", MCR_U8, {B});
    set_spec(h, BUF_STATUS, "status", MCR_I32, {STATUS_WORDS});
    set_spec(h, BUF_SCRATCH, "scratch", MCR_F32, {MCR_SCRATCH_FIELDS, N});
    set_spec(h, BUF_MANIFOLD, "manifold", MCR_F32, {B, MCR_MAX_MANIFOLDS, MCR_MANIFOLD_WORDS});
    set_spec(h, BUF_N_MANIFOLD, "n_manifold", MCR_I32, {B});
    set_spec(h, BUF_SCORE_SNAP, "score_snap", MCR_F64, {N});
    set_spec(h, BUF_BACKWARD_SNAP, "backward_snap", MCR_U8, {N});
    set_spec(h, BUF_PENDING, "pending_reset", MCR_U8, {B});
    set_spec(h, BUF_ACTION_STAGE, "action_stage", MCR_F64, {N, 3});
    set_spec(h, BUF_CAMERA_VP, "camera_vp", MCR_F32, {6, N});
    set_spec(h, BUF_TIMELINE, "timeline", MCR_F64, {TL_COUNT});      // u64 nanosecond stamps (8-byte slots)
    set_spec(h, BUF_ON_GRASS, "on_grass", MCR_U8, {N});
    const int64_t Np = cfg->particles ? N : 1;                           // skid traces are only kept on request
    set_spec(h, BUF_PRT_PTS, "prt_pts", MCR_F32, {PRT_MAX * PRT_PTS, 2, Np});
    set_spec(h, BUF_PRT_META, "prt_meta", MCR_I32, {PRT_MAX, Np});
    set_spec(h, BUF_PRT_HDR, "prt_hdr", MCR_I32, {2, Np});
    set_spec(h, BUF_SKID_START, "skid_start", MCR_F32, {4, 2, Np});
    set_spec(h, BUF_SKID_META, "skid_meta", MCR_I32, {4, Np});
    set_spec(h, BUF_TRK_T, "trk_T", MCR_I32, {P});
    set_spec(h, BUF_TRK_Q, "trk_Q", MCR_I32, {P});
    set_spec(h, BUF_TRK_NODE, "trk_node", MCR_F64, {P, T, 3});
    set_spec(h, BUF_TRK_TILE, "trk_tile", MCR_F32, {P, T, 8});
    set_spec(h, BUF_TRK_TILE_AABB, "trk_tile_aabb", MCR_F32, {P, T, 4});
    set_spec(h, BUF_TRK_QUAD, "trk_quad", MCR_F32, {P, Q, 8});
    set_spec(h, BUF_TRK_QUAD_COL, "trk_quad_col", MCR_U8, {P, Q});
    set_spec(h, BUF_TRK_QUAD_TILE, "trk_quad_tile", MCR_I16, {P, Q});
    set_spec(h, BUF_TRK_SLOT_POSE, "trk_slot_pose", MCR_F64, {P, 2, A, 3});
    set_spec(h, BUF_TRK_CHUNK, "trk_chunk", MCR_F32, {P, Q / MCR_QUAD_CHUNK, 4});
    set_spec(h, BUF_TRK_QUAD64, "trk_quad64", MCR_F64, {P, Q, 8});
    set_spec(h, BUF_DL_HDR, "dl_hdr", MCR_I32, {N, 4});
    set_spec(h, BUF_DL_META, "dl_meta", MCR_U32, {N, h->d.dl_cap, 2});
    set_spec(h, BUF_DL_EDGE, "dl_edge", MCR_F32, {N, h->d.dl_cap, 16});
    set_spec(h, BUF_DL_OCT, "dl_oct", MCR_F32, {N, A, 16});
    set_spec(h, BUF_MT_STATE, "mt_state", MCR_U32, {B, 625});
    set_spec(h, BUF_TRK_CONSUMED, "trk_consumed", MCR_I32, {B});
    set_spec(h, BUF_TRK_PRODUCED, "trk_produced", MCR_I32, {B});
    set_spec(h, BUF_TRK_LOCK, "trk_lock", MCR_I32, {B});
    set_spec(h, BUF_TG_SCRATCH, "tg_scratch", MCR_U8, {cfg->fresh_tracks > 0 ? B : 1, mcr_trackgen_scratch_bytes()});
    set_spec(h, BUF_READY, "ready", MCR_I32, {(int64_t)READY_WORDS(N)});
    *out = h;
    return 0;
}

extern "C" int mcr_destroy(mcr_handle h) {
    if (h) for (auto& g : h->graphs) destroy_step_graph(g);
    if (h && h->side_ready) {
        cudaStreamDestroy(h->side); cudaStreamDestroy(h->cap); cudaEventDestroy(h->ev_fork); cudaEventDestroy(h->ev_join);
        cudaStreamDestroy(h->side3); cudaEventDestroy(h->ev_post2); cudaEventDestroy(h->ev_score2);
        cudaStreamDestroy(h->copy); for (int i = 0; i < 8; ++i) cudaEventDestroy(h->ev_chunk[i]); cudaEventDestroy(h->ev_copy);
        cudaStreamSynchronize(h->refill); cudaStreamDestroy(h->refill); cudaEventDestroy(h->ev_refill); cudaEventDestroy(h->ev_refill_go);
        cudaStreamDestroy(h->side2); cudaEventDestroy(h->ev_pre); cudaEventDestroy(h->ev_contacts); cudaEventDestroy(h->ev_chain2); cudaEventDestroy(h->ev_score);
    }
    delete h;
    return 0;
}

extern "C" int mcr_buffer_count(mcr_handle h) { return h ? BUF_COUNT : fail(-1, "null handle"); }

extern "C" int mcr_buffer_spec(mcr_handle h, int i, const char** name, int32_t* dtype, int32_t* ndim, int64_t dims[4]) {
    if (!h || i < 0 || i >= BUF_COUNT) return fail(-1, "mcr_buffer_spec: bad index");
    const BufSpec& s = h->spec[i];
    if (name) *name = s.name;
    if (dtype) *dtype = s.dtype;
    if (ndim) *ndim = s.ndim;
    if (dims) for (int k = 0; k < 4; ++k) dims[k] = s.dims[k];
    return 0;
}

extern "C" int mcr_bind_buffer(mcr_handle h, int i, void* p) {
    if (!h || i < 0 || i >= BUF_COUNT) return fail(-1, "mcr_bind_buffer: bad index");
    if (!p || ((uintptr_t)p & 15)) return fail(-1, "mcr_bind_buffer(%s): pointer must be non-null and 16-byte aligned", h->spec[i].name);
    h->ptr[i] = p;
    DevBuffers& b = h->buf;
    switch (i) {
        case BUF_BODY: b.body = (float*)p; break;
        case BUF_SLEEP_TIME: b.sleep_time = (float*)p; break;
        case BUF_AWAKE: b.awake = (uint8_t*)p; break;
        case BUF_JOINT: b.joint = (float*)p; break;
        case BUF_LIMIT_STATE: b.limit_state = (uint8_t*)p; break;
        case BUF_WHEEL: b.wheel = (double*)p; break;
        case BUF_CTRL: b.ctrl = (double*)p; break;
        case BUF_ON_ROAD: b.on_road = (uint8_t*)p; break;
        case BUF_ON_ROAD_NEXT: b.on_road_next = (uint8_t*)p; break;
        case BUF_REWARD: b.reward = (double*)p; break;
        case BUF_PREV_REWARD: b.prev_reward = (double*)p; break;
        case BUF_VISIT_COUNT: b.visit_count = (int32_t*)p; break;
        case BUF_BACKWARD: b.backward = (uint8_t*)p; break;
        case BUF_TIME: b.time = (double*)p; break;
        case BUF_STEPS: b.steps = (int32_t*)p; break;
        case BUF_CAMERA: b.camera = (float*)p; break;
        case BUF_STRIPE: b.stripe = (float*)p; break;
        case BUF_HEADING: b.heading = (double*)p; break;
        case BUF_ENV_TRACK: b.env_track = (int32_t*)p; break;
        case BUF_ENV_CW: b.env_cw = (uint8_t*)p; break;
        case BUF_ENV_EPISODE: b.env_episode = (uint32_t*)p; break;
        case BUF_VISITED: b.visited = (uint32_t*)p; break;
        case BUF_TOUCHED: b.touched = (uint8_t*)p; break;
        case BUF_RESET_MASK: b.reset_mask = (uint8_t*)p; break;
        case BUF_STATUS: b.status = (int32_t*)p; break;
        case BUF_SCRATCH: b.scratch = (float*)p; break;
        case BUF_MANIFOLD: b.manifold = (float*)p; break;
        case BUF_N_MANIFOLD: b.n_manifold = (int32_t*)p; break;
        case BUF_SCORE_SNAP: b.score_snap = (double*)p; break;
        case BUF_BACKWARD_SNAP: b.backward_snap = (uint8_t*)p; break;
        case BUF_PENDING: b.pending = (uint8_t*)p; break;
        case BUF_ACTION_STAGE: b.action_stage = (double*)p; break;
        case BUF_CAMERA_VP: b.camera_vp = (float*)p; break;
        case BUF_TIMELINE: b.timeline = (unsigned long long*)p; break;
        case BUF_ON_GRASS: b.on_grass = (uint8_t*)p; break;
        case BUF_PRT_PTS: b.prt_pts = (float*)p; break;
        case BUF_PRT_META: b.prt_meta = (int32_t*)p; break;
        case BUF_PRT_HDR: b.prt_hdr = (int32_t*)p; break;
        case BUF_SKID_START: b.skid_start = (float*)p; break;
        case BUF_SKID_META: b.skid_meta = (int32_t*)p; break;
        case BUF_TRK_QUAD64: b.trk_quad64 = (double*)p; break;
        case BUF_TRK_T: b.trk_T = (int32_t*)p; break;
        case BUF_TRK_Q: b.trk_Q = (int32_t*)p; break;
        case BUF_TRK_NODE: b.trk_node = (double*)p; break;
        case BUF_TRK_TILE: b.trk_tile = (float*)p; break;
        case BUF_TRK_TILE_AABB: b.trk_tile_aabb = (float*)p; break;
        case BUF_TRK_QUAD: b.trk_quad = (float*)p; break;
        case BUF_TRK_QUAD_COL: b.trk_quad_col = (uint8_t*)p; break;
        case BUF_TRK_QUAD_TILE: b.trk_quad_tile = (int16_t*)p; break;
        case BUF_TRK_SLOT_POSE: b.trk_slot_pose = (double*)p; break;
        case BUF_TRK_CHUNK: b.trk_chunk = (float*)p; break;
        case BUF_DL_HDR: b.dl_hdr = (int32_t*)p; break;
        case BUF_DL_META: b.dl_meta = (uint32_t*)p; break;
        case BUF_DL_EDGE: b.dl_edge = (float*)p; break;
        case BUF_DL_OCT: b.dl_oct = (float*)p; break;
        case BUF_MT_STATE: b.mt_state = (uint32_t*)p; break;
        case BUF_TRK_CONSUMED: b.trk_consumed = (int32_t*)p; break;
        case BUF_TRK_PRODUCED: b.trk_produced = (int32_t*)p; break;
        case BUF_TRK_LOCK: b.trk_lock = (int32_t*)p; break;
        case BUF_TG_SCRATCH: b.tg_scratch = (unsigned char*)p; break;
        case BUF_READY: b.ready = (int32_t*)p; break;
    }
    return 0;
}

static int check_bound(mcr_handle h) {
    if (!h) return fail(-1, "null handle");
    for (int i = 0; i < BUF_COUNT; ++i)
        if (!h->ptr[i]) return fail(-3, "buffer '%s' is not bound (mcr_bind_buffer)", h->spec[i].name);
    return 0;
}

extern "C" int mcr_get_mass(mcr_handle h, float* o) {
    if (!h || !o) return fail(-1, "mcr_get_mass: null argument");
    const CarConst& c = h->cc;
    o[0] = c.hull_mass; o[1] = c.hull_invMass; o[2] = c.hull_I; o[3] = c.hull_invI; o[4] = c.hull_lcx; o[5] = c.hull_lcy;
    o[6] = c.wheel_mass; o[7] = c.wheel_invMass; o[8] = c.wheel_I; o[9] = c.wheel_invI; o[10] = 0.0f; o[11] = 0.0f;
    return 0;
}

extern "C" int mcr_get_shape(mcr_handle h, int32_t which, float* o) {
    if (!h || !o || which < 0 || which > 4) return fail(-1, "mcr_get_shape: bad argument");
    const Poly8& P = which < 4 ? h->cc.hull_poly[which] : h->cc.wheel_poly;
    for (int i = 0; i < P.n; ++i) { o[2 * i] = P.x[i]; o[2 * i + 1] = P.y[i]; }
    return P.n;
}

extern "C" int64_t mcr_launch_count(mcr_handle h) { return h ? h->launches : -1; }

extern "C" int mcr_set_obs_format(mcr_handle h, int32_t format) {
    if (!h) return fail(-1, "null handle");
    if (format < MCR_OBS_RGB_HWC || format > MCR_OBS_RGB_CHW_F16)
        return fail(-1, "mcr_set_obs_format: unknown format %d", format);
    if (format != h->obs_format) {
        // captured step graphs bake the layout in
        for (auto& g : h->graphs) destroy_step_graph(g);
        h->graphs.clear();
        h->obs_format = format;
    }
    return 0;
}

extern "C" int mcr_set_frame_stack(mcr_handle h, int32_t k) {
    if (!h) return fail(-1, "null handle");
    if (k < 1 || k > 16) return fail(-1, "mcr_set_frame_stack: depth must be in [1, 16]");
    if (k != h->stack_k) {
        for (auto& g : h->graphs) destroy_step_graph(g);    // captured step graphs bake the depth in
        h->graphs.clear();
        h->stack_k = k;
    }
    return 0;
}

extern "C" int64_t mcr_obs_bytes(mcr_handle h) {
    if (!h) return -1;
    switch (h->obs_format) {
        case MCR_OBS_GRAY: return MCR_STATE_W * MCR_STATE_H;
        case MCR_OBS_GRAY_STACK: return (int64_t)h->stack_k * MCR_STATE_W * MCR_STATE_H;
        case MCR_OBS_RGB_CHW_F16: return 2 * MCR_OBS_BYTES;
        default: return MCR_OBS_BYTES;
    }
}

// ---------------------------------------------------------------------------------------
// tracks
// ---------------------------------------------------------------------------------------
static inline uint8_t col_u8(float c) { return (uint8_t)(int)std::floor(c * 255.0f + 0.5f); }

extern "C" int mcr_load_track(mcr_handle h, int32_t slot, int32_t T, const double* nodes, int32_t Q, const double* quads,
                              const float* quad_rgb, const int32_t* quad_tile, void* stream) {
    int rc = check_bound(h); if (rc) return rc;
    const Dims& d = h->d;
    if (slot < 0 || slot >= d.P) return fail(-1, "mcr_load_track: slot %d out of range", slot);
    if (T < 1 || T > d.Tmax) return fail(-1, "mcr_load_track: T=%d exceeds max_tiles=%d", T, d.Tmax);
    if (Q < T || Q > d.Qmax) return fail(-1, "mcr_load_track: Q=%d exceeds max_quads=%d", Q, d.Qmax);
    CUDA_OK(cudaSetDevice(h->cfg.device));
    if (!h->palette_ready) { std::memcpy(h->palette, mcr_host_palette(), sizeof(h->palette)); h->palette_ready = true; }
    std::vector<double> node3((size_t)T * 3);
    for (int i = 0; i < T; ++i) { node3[3 * i] = nodes[4 * i + 1]; node3[3 * i + 1] = nodes[4 * i + 2]; node3[3 * i + 2] = nodes[4 * i + 3]; }
    std::vector<float> quadf((size_t)Q * 8), tile((size_t)T * 8, 0.0f), aabb((size_t)T * 4, 0.0f);
    std::vector<uint8_t> qcol(Q); std::vector<int16_t> qtile(Q);
    std::vector<char> seen(T, 0);
    for (int q = 0; q < Q; ++q) {
        for (int k = 0; k < 8; ++k) quadf[(size_t)q * 8 + k] = (float)quads[(size_t)q * 8 + k];   // glVertex3f / b2Vec2
        const uint8_t r = col_u8(quad_rgb[3 * q]), g = col_u8(quad_rgb[3 * q + 1]), bl = col_u8(quad_rgb[3 * q + 2]);
        int pal = -1;
        for (int p = 0; p < PAL_COUNT; ++p) if (h->palette[p][0] == r && h->palette[p][1] == g && h->palette[p][2] == bl) { pal = p; break; }
        if (pal < 0) return fail(-4, "mcr_load_track: quad %d colour (%d,%d,%d) is not in the palette", q, r, g, bl);
        qcol[q] = (uint8_t)pal;
        const int t = quad_tile[q];
        if (t >= T) return fail(-4, "mcr_load_track: quad %d refers to tile %d >= T", q, t);
        qtile[q] = (int16_t)(t < 0 ? -1 : t);
        if (t >= 0) {
            std::vector<P2> raw;
            for (int k = 0; k < 4; ++k) raw.push_back(P2{quadf[(size_t)q * 8 + 2 * k], quadf[(size_t)q * 8 + 2 * k + 1]});
            const std::vector<P2> poly = b2_polygon_set(raw);      // fd_tile.shape.vertices = ..., mcr:318
            if (poly.size() != 4) return fail(-5, "mcr_load_track: tile %d is not a proper quad after b2PolygonShape::Set", t);
            float lx = poly[0].x, ly = poly[0].y, hx = lx, hy = ly;
            for (int k = 0; k < 4; ++k) {
                tile[(size_t)t * 8 + 2 * k] = poly[k].x; tile[(size_t)t * 8 + 2 * k + 1] = poly[k].y;
                lx = std::fmin(lx, poly[k].x); ly = std::fmin(ly, poly[k].y); hx = std::fmax(hx, poly[k].x); hy = std::fmax(hy, poly[k].y);
            }
            aabb[(size_t)t * 4] = lx; aabb[(size_t)t * 4 + 1] = ly; aabb[(size_t)t * 4 + 2] = hx; aabb[(size_t)t * 4 + 3] = hy;
            seen[t] = 1;
        }
    }
    for (int t = 0; t < T; ++t) if (!seen[t]) return fail(-4, "mcr_load_track: tile %d has no quad", t);
    // bounding circle of every MCR_QUAD_CHUNK consecutive road_poly quads (rasteriser culling)
    const int nchunk = d.Qmax / MCR_QUAD_CHUNK;
    std::vector<float> chunk((size_t)nchunk * 4, 0.0f);
    for (int c = 0; c * MCR_QUAD_CHUNK < Q; ++c) {
        const int q0 = c * MCR_QUAD_CHUNK, q1 = std::min(Q, q0 + MCR_QUAD_CHUNK);
        double sx = 0, sy = 0; int nvert = 0;
        for (int q = q0; q < q1; ++q) for (int k = 0; k < 4; ++k) { sx += quadf[(size_t)q * 8 + 2 * k]; sy += quadf[(size_t)q * 8 + 2 * k + 1]; ++nvert; }
        const double mx = sx / nvert, my = sy / nvert;
        double r2 = 0;
        for (int q = q0; q < q1; ++q) for (int k = 0; k < 4; ++k) {
            const double dx = quadf[(size_t)q * 8 + 2 * k] - (double)(float)mx, dy = quadf[(size_t)q * 8 + 2 * k + 1] - (double)(float)my;
            r2 = std::max(r2, dx * dx + dy * dy);
        }
        chunk[(size_t)c * 4] = (float)mx; chunk[(size_t)c * 4 + 1] = (float)my;
        chunk[(size_t)c * 4 + 2] = (float)(std::sqrt(r2) * 1.0001 + 1e-3);
    }
    // spawn pose of every grid position for the device-side auto reset
    const int A = d.A;
    std::vector<double> slot_pose((size_t)2 * A * 3);
    std::vector<int32_t> ident(A);
    for (int c = 0; c < A; ++c) ident[c] = c;
    for (int cw = 0; cw < 2; ++cw) {
        rc = mcr_spawn_poses(nodes, T, ident.data(), A, cw, slot_pose.data() + (size_t)cw * A * 3);
        if (rc) return rc;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const DevBuffers& b = h->buf;
    const int32_t Ti = T, Qi = Q;
    CUDA_OK(cudaMemcpyAsync(b.trk_T + slot, &Ti, 4, cudaMemcpyHostToDevice, s));
    CUDA_OK(cudaMemcpyAsync(b.trk_Q + slot, &Qi, 4, cudaMemcpyHostToDevice, s));
    CUDA_OK(cudaMemcpyAsync(b.trk_node + (size_t)slot * d.Tmax * 3, node3.data(), node3.size() * 8, cudaMemcpyHostToDevice, s));
    CUDA_OK(cudaMemcpyAsync(b.trk_tile + (size_t)slot * d.Tmax * 8, tile.data(), tile.size() * 4, cudaMemcpyHostToDevice, s));
    CUDA_OK(cudaMemcpyAsync(b.trk_tile_aabb + (size_t)slot * d.Tmax * 4, aabb.data(), aabb.size() * 4, cudaMemcpyHostToDevice, s));
    CUDA_OK(cudaMemcpyAsync(b.trk_quad + (size_t)slot * d.Qmax * 8, quadf.data(), quadf.size() * 4, cudaMemcpyHostToDevice, s));
    CUDA_OK(cudaMemcpyAsync(b.trk_quad64 + (size_t)slot * d.Qmax * 8, quads, (size_t)Q * 8 * 8, cudaMemcpyHostToDevice, s));
    CUDA_OK(cudaMemcpyAsync(b.trk_quad_col + (size_t)slot * d.Qmax, qcol.data(), qcol.size(), cudaMemcpyHostToDevice, s));
    CUDA_OK(cudaMemcpyAsync(b.trk_quad_tile + (size_t)slot * d.Qmax, qtile.data(), qtile.size() * 2, cudaMemcpyHostToDevice, s));
    CUDA_OK(cudaMemcpyAsync(b.trk_chunk + (size_t)slot * nchunk * 4, chunk.data(), chunk.size() * 4, cudaMemcpyHostToDevice, s));
    CUDA_OK(cudaMemcpyAsync(b.trk_slot_pose + (size_t)slot * 2 * A * 3, slot_pose.data(), slot_pose.size() * 8, cudaMemcpyHostToDevice, s));
    CUDA_OK(cudaStreamSynchronize(s));   // the staging vectors die with this frame
    return 0;
}

// ---------------------------------------------------------------------------------------
// reset / step
// ---------------------------------------------------------------------------------------
#define LAUNCH(expr) do { int n_ = (expr); if (n_ < 0) return fail(-101, "kernel launch failed in %s: %s", #expr, cudaGetErrorString(cudaGetLastError())); h->launches += n_; } while (0)

extern "C" int mcr_tracks_generate_device(mcr_handle h, int32_t n, uint32_t* d_mt_state, const int32_t* d_slot,
                                          void* d_scratch, int32_t* d_result, void* stream) {
    int rc = check_bound(h); if (rc) return rc;
    if (n < 1 || !d_mt_state || !d_scratch || !d_result) return fail(-1, "mcr_tracks_generate_device: bad arguments");
    CUDA_OK(cudaSetDevice(h->cfg.device));
    LAUNCH(launch_trackgen(h->d, h->buf, n, d_mt_state, d_slot, d_scratch, d_result, 64, stream));
    return 0;
}

static int ensure_side(mcr_handle h) {
    if (!h->side_ready) {
        CUDA_OK(cudaSetDevice(h->cfg.device));
        int prio_lo = 0, prio_hi = 0;
        CUDA_OK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CUDA_OK(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
        CUDA_OK(cudaStreamCreateWithFlags(&h->cap, cudaStreamNonBlocking));
        CUDA_OK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
        CUDA_OK(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
        CUDA_OK(cudaStreamCreateWithFlags(&h->side2, cudaStreamNonBlocking));
        CUDA_OK(cudaEventCreateWithFlags(&h->ev_pre, cudaEventDisableTiming));
        CUDA_OK(cudaEventCreateWithFlags(&h->ev_contacts, cudaEventDisableTiming));
        CUDA_OK(cudaEventCreateWithFlags(&h->ev_chain2, cudaEventDisableTiming));
        CUDA_OK(cudaEventCreateWithFlags(&h->ev_score, cudaEventDisableTiming));
        CUDA_OK(cudaStreamCreateWithFlags(&h->side3, cudaStreamNonBlocking));
        CUDA_OK(cudaEventCreateWithFlags(&h->ev_post2, cudaEventDisableTiming));
        CUDA_OK(cudaEventCreateWithFlags(&h->ev_score2, cudaEventDisableTiming));
        CUDA_OK(cudaStreamCreateWithFlags(&h->copy, cudaStreamNonBlocking));
        for (int i = 0; i < 8; ++i) CUDA_OK(cudaEventCreateWithFlags(&h->ev_chunk[i], cudaEventDisableTiming));
        CUDA_OK(cudaEventCreateWithFlags(&h->ev_copy, cudaEventDisableTiming));
        CUDA_OK(cudaStreamCreateWithPriority(&h->refill, cudaStreamNonBlocking, prio_lo));
        CUDA_OK(cudaEventCreateWithFlags(&h->ev_refill, cudaEventDisableTiming));
        CUDA_OK(cudaEventCreateWithFlags(&h->ev_refill_go, cudaEventDisableTiming));
        h->side_ready = true;
    }
    return 0;
}

// contacts || (pre -> sweep), then post.  The contact pass only reads the step's start poses, so it
// runs on the handle's side stream while the solver occupies the main one; post_kernel (which
// overwrites the poses and consumes on_road_next) waits for both.
static int simulate(mcr_handle h, const uint8_t* mask, const void* action, int32_t action_dtype, void* stream,
                    const uint8_t* noact = nullptr) {
    cudaStream_t s = (cudaStream_t)stream;
    { int rc_ = ensure_side(h); if (rc_) return rc_; }
    CUDA_OK(cudaEventRecord(h->ev_fork, s));
    CUDA_OK(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
    LAUNCH(launch_contacts(h->d, h->buf, h->cc, mask, h->side));
    CUDA_OK(cudaEventRecord(h->ev_join, h->side));
    LAUNCH(launch_physics(h->d, h->buf, h->cc, mask, noact, action, action_dtype, h->cfg.h_ratio, h->cfg.collisions, s));
    CUDA_OK(cudaStreamWaitEvent(s, h->ev_join, 0));
    LAUNCH(launch_physics_post(h->d, h->buf, h->cc, mask, noact, action != nullptr, h->cfg.h_ratio, 0, s));
    return 0;
}

extern "C" int mcr_contacts(mcr_handle h, const uint8_t* mask, void* stream) {
    int rc = check_bound(h); if (rc) return rc;
    LAUNCH(launch_contacts(h->d, h->buf, h->cc, mask, stream));
    return 0;
}

extern "C" int mcr_physics(mcr_handle h, const uint8_t* mask, const void* action, int32_t action_dtype, void* stream) {
    int rc = check_bound(h); if (rc) return rc;
    if (action && action_dtype != MCR_F32 && action_dtype != MCR_F64) return fail(-1, "action dtype must be MCR_F32 or MCR_F64");
    LAUNCH(launch_physics(h->d, h->buf, h->cc, mask, nullptr, action, action_dtype, h->cfg.h_ratio, h->cfg.collisions, stream));
    LAUNCH(launch_physics_post(h->d, h->buf, h->cc, mask, nullptr, action != nullptr, h->cfg.h_ratio, 0, stream));
    return 0;
}

extern "C" int mcr_simulate(mcr_handle h, const uint8_t* mask, const void* action, int32_t action_dtype, void* stream) {
    int rc = check_bound(h); if (rc) return rc;
    if (action && action_dtype != MCR_F32 && action_dtype != MCR_F64) return fail(-1, "action dtype must be MCR_F32 or MCR_F64");
    return simulate(h, mask, action, action_dtype, stream);
}

// render || score: the rasteriser only reads the snapshots post_kernel took, so the reward / done
// block runs beside it on the side stream; the caller's stream waits for both.
static int render_and_score(mcr_handle h, const uint8_t* mask, uint8_t* obs, double* reward, uint8_t* done,
                            int post_step, void* stream, const uint8_t* noact = nullptr) {
    cudaStream_t s = (cudaStream_t)stream;
    if (post_step) {
        int rc = ensure_side(h); if (rc) return rc;
        CUDA_OK(cudaEventRecord(h->ev_fork, s));
        CUDA_OK(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
        LAUNCH(launch_score(h->d, h->buf, mask, noact, reward, done, h->cfg.max_episode_steps, 0, h->side));
        CUDA_OK(cudaEventRecord(h->ev_join, h->side));
    }
    LAUNCH(launch_render(h->d, h->buf, h->cc, mask, obs, h->cfg.backwards_flag, h->cfg.use_ego_color, 0, h->obs_format, h->stack_k, s));
    if (post_step) CUDA_OK(cudaStreamWaitEvent(s, h->ev_join, 0));
    return 0;
}

extern "C" int mcr_render(mcr_handle h, const uint8_t* mask, uint8_t* obs, double* reward, uint8_t* done, int32_t post_step, void* stream) {
    int rc = check_bound(h); if (rc) return rc;
    if (!obs) return fail(-1, "mcr_render: d_obs is null");
    if (post_step && (!reward || !done)) return fail(-1, "mcr_render: post_step needs d_reward and d_done");
    return render_and_score(h, mask, obs, reward, done, post_step, stream);
}

extern "C" int mcr_render_viewport(mcr_handle h, const uint8_t* mask, int32_t vw, int32_t vh, uint8_t* out, void* stream) {
    int rc = check_bound(h); if (rc) return rc;
    if (!out) return fail(-1, "mcr_render_viewport: d_out is null");
    if (vw < 8 || vh < 8 || vw > 4096 || vh > 4096 || (vw & 3)) return fail(-1, "mcr_render_viewport: viewport must be 8..4096 pixels, width a multiple of 4");
    LAUNCH(launch_render_viewport(h->d, h->buf, h->cc, mask, out, h->buf.camera_vp, vw, vh, h->cfg.h_ratio,
                                  h->cfg.backwards_flag, h->cfg.use_ego_color, stream));
    return 0;
}

struct HostOut { uint8_t* h_obs; double* h_reward; uint8_t* h_done; };   // pinned host destinations of mcr_step_host
static int pipeline(mcr_handle h, const uint8_t* mask, const uint8_t* noact, const void* action, int32_t action_dtype,
                    uint8_t* obs, double* reward, uint8_t* done, int post_step, void* stream, const uint8_t* reset_flags = nullptr,
                    const HostOut* ho = nullptr);

// Top up the track rings on the refill stream, ordered behind what `after` holds so far (mcr_reset: the spawn kernel
// that zeroes the ring counters; steps: nothing to wait for, the kernel only looks at the counters).
static int kick_refill(mcr_handle h, cudaStream_t after) {
    if (after) {
        CUDA_OK(cudaEventRecord(h->ev_refill_go, after));
        CUDA_OK(cudaStreamWaitEvent(h->refill, h->ev_refill_go, 0));
    }
    LAUNCH(launch_ring_refill(h->d, h->buf, h->cfg.fresh_tracks, h->refill));
    CUDA_OK(cudaEventRecord(h->ev_refill, h->refill));
    h->steps_since_refill = 0;
    return 0;
}

extern "C" int mcr_reset(mcr_handle h, const uint8_t* mask, const int32_t* track_slot, const uint8_t* cw,
                         const double* spawn_pose, uint8_t* obs, void* stream) {
    int rc = check_bound(h); if (rc) return rc;
    if (!track_slot || !cw || !spawn_pose || !obs) return fail(-1, "mcr_reset: null argument");
    const bool fresh = h->cfg.fresh_tracks > 0;
    if (fresh) {
        rc = ensure_side(h); if (rc) return rc;
        // a refill of the previous episodes may still be writing the rings this reset re-initialises
        CUDA_OK(cudaStreamWaitEvent((cudaStream_t)stream, h->ev_refill, 0));
    }
    LAUNCH(launch_spawn(h->d, h->buf, h->cc, mask, track_slot, cw, spawn_pose, stream));
    if (fresh) { rc = kick_refill(h, (cudaStream_t)stream); if (rc) return rc; }
    // the implicit step(None), mcr:408
    return pipeline(h, mask, nullptr, nullptr, MCR_F32, obs, nullptr, nullptr, 0, stream);
}

// One full pass (contacts, physics, post-step block, render) as two chains that only meet at the end:
//   side   : contacts_kernel (reads the start poses) -> stripe_kernel        ............. score_kernel
//   main   : head (auto reset + carcontacts + pre) -> sweep -> post(cls 1) -> render(cls 1)
//   side2  :                   \-> coupled_kernel ---------> post(cls 2) -> score(cls 2) -> render(cls 2)
// cls 1 = envs without car-car manifolds (per-car solver), cls 2 = envs with touching cars.  The
// coupled solver (joints + contacts of a merged island in lock step, 100-250 us when any env of the
// batch has touching cars) no longer stalls the other envs' post/render; the classes are disjoint
// sets of envs, so the chains never touch the same state.
static int pipeline(mcr_handle h, const uint8_t* mask, const uint8_t* noact, const void* action, int32_t action_dtype,
                    uint8_t* obs, double* reward, uint8_t* done, int post_step, void* stream, const uint8_t* reset_flags,
                    const HostOut* ho) {
    cudaStream_t s = (cudaStream_t)stream;
    { int rc_ = ensure_side(h); if (rc_) return rc_; }
    const Dims& d = h->d; const DevBuffers& b = h->buf; const CarConst& cc = h->cc;
    const bool split = h->cfg.collisions && d.A > 1;
    static const int early_exit = std::getenv("MCR_NO_EARLY_EXIT") ? 0 : 1;   // diagnostics only
    const AutoResetCfg ar{h->cfg.use_random_direction, h->cfg.direction_cw, (unsigned long long)h->cfg.seed, h->cfg.fresh_tracks};
    // The fused head kernel (warp per env: auto reset + car-car narrow phase + the per-car head on lanes 0..A-1) takes two
    // launch latencies off the chain.  MCR_HEAD_SPLIT=1 runs the three stages as packed kernels instead (thread per car
    // for the per-car head): measured slower at every batch size, 1 k to 16 k envs (profiles/README r02) -- diagnostics only.
    static const char* head_env = std::getenv("MCR_HEAD_SPLIT");
    const bool head_split = head_env && head_env[0] == '1';
    // Hand-offs of the main chain by per-car / per-frame ready flags instead of grid-wide dependencies (DevBuffers::ready):
    //   1  sweep -> post -> project -> fill: the consumer is launched early and each CTA waits for its own cars / frame
    //   2  also contacts / stripes -> post, so that post_kernel's only stream dependency is the sweep and it is placed while
    //      the sweep runs (a second dependency costs the programmatic launch).  Only while post_kernel's grid is small: it
    //      then waits, resident, for kernels that are launched after it.
    // MCR_FLAG_HANDOFF=0|1|2 overrides (A/B); read per pass, i.e. once per graph capture.
    int handoff = d.N <= 2304 ? 2 : 1;
    if (const char* e = std::getenv("MCR_FLAG_HANDOFF")) handoff = std::min(handoff, std::max(0, atoi(e)));
    if (reset_flags) {
        // next-step auto reset: the flagged envs are respawned first, so the contact pass (which reads the start poses)
        // has to follow it
        if (head_split) LAUNCH(launch_auto_reset(d, b, cc, reset_flags, ar, s));
        else LAUNCH(launch_head(d, b, cc, mask, reset_flags, ar, action, action_dtype, h->cfg.collisions, s));
        CUDA_OK(cudaEventRecord(h->ev_fork, s));
        CUDA_OK(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
        LAUNCH(launch_contacts(d, b, cc, mask, h->side, handoff == 2));
        if (head_split) LAUNCH(launch_presweep(d, b, cc, mask, noact, action, action_dtype, h->cfg.collisions, 0, s));
    } else {
        CUDA_OK(cudaEventRecord(h->ev_fork, s));
        CUDA_OK(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
        LAUNCH(launch_contacts(d, b, cc, mask, h->side, handoff == 2));
        if (noact || head_split) LAUNCH(launch_presweep(d, b, cc, mask, noact, action, action_dtype, h->cfg.collisions, 0, s));
        else LAUNCH(launch_head(d, b, cc, mask, nullptr, ar, action, action_dtype, h->cfg.collisions, s));
    }
    CUDA_OK(cudaEventRecord(h->ev_pre, s));
    // wheel stripes for the rasteriser: they need pre_kernel's phase and nothing else, so they ride on
    // the side stream behind the contact pass; ev_contacts (which every post_kernel waits for) covers both
    CUDA_OK(cudaStreamWaitEvent(h->side, h->ev_pre, 0));
    LAUNCH(launch_stripes(d, b, mask, h->side, handoff == 2));
    CUDA_OK(cudaEventRecord(h->ev_contacts, h->side));
    if (split) {
        CUDA_OK(cudaStreamWaitEvent(h->side2, h->ev_pre, 0));
        LAUNCH(launch_coupled(d, b, cc, mask, early_exit, h->side2));
        CUDA_OK(cudaStreamWaitEvent(h->side2, h->ev_contacts, 0));
        LAUNCH(launch_physics_post(d, b, cc, mask, noact, action != nullptr, h->cfg.h_ratio, 2, h->side2));
        // this chain is the long pole whenever an env has touching cars (coupled_kernel: 220 us): render first,
        // the reward / done block after it costs nothing on the main chain's clock only if it is short, so it
        // goes on the third side stream beside the rasteriser
        if (post_step) {
            CUDA_OK(cudaEventRecord(h->ev_post2, h->side2));
            CUDA_OK(cudaStreamWaitEvent(h->side3, h->ev_post2, 0));
            LAUNCH(launch_score(d, b, mask, noact, reward, done, h->cfg.max_episode_steps, 2, h->side3));
            CUDA_OK(cudaEventRecord(h->ev_score2, h->side3));
        }
        LAUNCH(launch_render(d, b, cc, mask, obs, h->cfg.backwards_flag, h->cfg.use_ego_color, 2, h->obs_format, h->stack_k, h->side2));
        if (post_step) CUDA_OK(cudaStreamWaitEvent(h->side2, h->ev_score2, 0));
        CUDA_OK(cudaEventRecord(h->ev_chain2, h->side2));
    }
    LAUNCH(launch_presweep(d, b, cc, mask, noact, action, action_dtype, h->cfg.collisions, 1, s, handoff >= 1));
    if (handoff < 2) CUDA_OK(cudaStreamWaitEvent(s, h->ev_contacts, 0));
    const int cls = split ? 1 : 0;
    const int flag_handoff = handoff >= 1;
    LAUNCH(launch_physics_post(d, b, cc, mask, noact, action != nullptr, h->cfg.h_ratio, cls, s, handoff));
    // the reward / done block: inside the projector's launch (nothing between post_kernel and it in this stream, so the
    // programmatic dependent launch holds), or -- fused rasteriser -- as its own kernel on the side stream
    const bool score_in_render = post_step && render_runs_score(cls);
    if (post_step && !score_in_render) {
        CUDA_OK(cudaEventRecord(h->ev_join, s));
        CUDA_OK(cudaStreamWaitEvent(h->side, h->ev_join, 0));
        LAUNCH(launch_score(d, b, mask, noact, reward, done, h->cfg.max_episode_steps, cls, h->side));
        CUDA_OK(cudaEventRecord(h->ev_score, h->side));
    }
    double* sc_reward = score_in_render ? reward : nullptr;
    const size_t frame_bytes = (size_t)mcr_obs_bytes(h);
    const bool chunked = ho && render_is_split() && d.B >= 64;
    if (chunked) {
        // mcr_step_host: the frames go to the host in up to 8 ranges of envs; the copy of a range starts as soon as it is
        // filled, beside the fill of the next ranges -- only the physics and the first range's fill are not hidden
        // behind the PCIe transfer.  (The touching-car chain's frames lie scattered in every range: wait for it first.)
        LAUNCH(launch_project(d, b, cc, mask, h->cfg.backwards_flag, h->cfg.use_ego_color, cls, s, noact, sc_reward, done, h->cfg.max_episode_steps, flag_handoff));
        if (split) CUDA_OK(cudaStreamWaitEvent(h->copy, h->ev_chain2, 0));
        // ranges grow geometrically (B/16, B/16, B/8, B/4, B/2): the first copy starts after a sixteenth of the fill, and
        // the later, larger copies keep the per-copy set-up cost off the link
        int env0 = 0;
        for (int c = 0; c < 5 && env0 < d.B; ++c) {
            const int want = std::max(1, c == 0 ? d.B / 16 : (d.B >> (5 - c)));
            const int nenv = c == 4 ? d.B - env0 : std::min(want, d.B - env0);
            LAUNCH(launch_fill(d, b, mask, obs, cls, h->obs_format, h->stack_k, env0, nenv, true, s, flag_handoff));
            CUDA_OK(cudaEventRecord(h->ev_chunk[c], s));
            CUDA_OK(cudaStreamWaitEvent(h->copy, h->ev_chunk[c], 0));
            const size_t off = (size_t)env0 * d.A * frame_bytes;
            CUDA_OK(cudaMemcpyAsync(ho->h_obs + off, obs + off, (size_t)nenv * d.A * frame_bytes, cudaMemcpyDeviceToHost, h->copy));
            env0 += nenv;
        }
        CUDA_OK(cudaEventRecord(h->ev_copy, h->copy));
    } else {
        LAUNCH(launch_render(d, b, cc, mask, obs, h->cfg.backwards_flag, h->cfg.use_ego_color, cls, h->obs_format, h->stack_k, s,
                             noact, sc_reward, done, h->cfg.max_episode_steps, flag_handoff));
    }
    if (post_step && !score_in_render) CUDA_OK(cudaStreamWaitEvent(s, h->ev_score, 0));
    if (handoff == 2) CUDA_OK(cudaStreamWaitEvent(s, h->ev_contacts, 0));      // the side stream joins here instead of in front of post_kernel
    if (split) CUDA_OK(cudaStreamWaitEvent(s, h->ev_chain2, 0));
    if (ho) {
        if (!chunked) CUDA_OK(cudaMemcpyAsync(ho->h_obs, obs, (size_t)d.N * frame_bytes, cudaMemcpyDeviceToHost, s));
        if (post_step) {
            CUDA_OK(cudaMemcpyAsync(ho->h_reward, reward, (size_t)d.N * sizeof(double), cudaMemcpyDeviceToHost, s));
            CUDA_OK(cudaMemcpyAsync(ho->h_done, done, (size_t)d.B, cudaMemcpyDeviceToHost, s));
        }
        if (chunked) CUDA_OK(cudaStreamWaitEvent(s, h->ev_copy, 0));
    }
    return 0;
}

static int step_enqueue(mcr_handle h, const void* action, int32_t action_dtype, uint8_t* obs, double* reward,
                        uint8_t* done, int32_t flags, void* stream, const HostOut* ho = nullptr) {
    int rc;
    if (flags & 2) {
        // next-step auto reset: envs whose previous step ended the episode respawn now and take
        // reset()'s step(None) inside this very pass -- one pass, no masked second pass
        // (the respawn itself is the first stage of the pipeline's head kernel; it writes reset_mask)
        return pipeline(h, nullptr, h->buf.reset_mask, action, action_dtype, obs, reward, done, 1, stream, h->buf.pending, ho);
    }
    // same-step auto reset rewrites the frames of the envs that ended: their host copy has to follow the second pass
    const HostOut* ho1 = (flags & 1) ? nullptr : ho;
    rc = pipeline(h, nullptr, nullptr, action, action_dtype, obs, reward, done, 1, stream, nullptr, ho1); if (rc) return rc;
    if (flags & 1) {
        AutoResetCfg ar{h->cfg.use_random_direction, h->cfg.direction_cw, (unsigned long long)h->cfg.seed, h->cfg.fresh_tracks};
        LAUNCH(launch_auto_reset(h->d, h->buf, h->cc, done, ar, stream));
        rc = pipeline(h, h->buf.reset_mask, nullptr, nullptr, MCR_F32, obs, nullptr, nullptr, 0, stream); if (rc) return rc;
        if (ho) {
            cudaStream_t s = (cudaStream_t)stream;
            CUDA_OK(cudaMemcpyAsync(ho->h_obs, obs, (size_t)h->d.N * (size_t)mcr_obs_bytes(h), cudaMemcpyDeviceToHost, s));
            CUDA_OK(cudaMemcpyAsync(ho->h_reward, reward, (size_t)h->d.N * sizeof(double), cudaMemcpyDeviceToHost, s));
            CUDA_OK(cudaMemcpyAsync(ho->h_done, done, (size_t)h->d.B, cudaMemcpyDeviceToHost, s));
        }
    }
    return 0;
}

// The step is 9 dependent launches on two streams; issued one by one they cost more host time than
// the GPU needs to run them (launch bound at ~0.1 ms per step).  After two eager steps (one-time
// kernel attributes and constant uploads are done by then) the launch sequence is captured once per
// argument tuple into a CUDA graph that reads the action from the bound `action_stage` buffer; a
// step is then one device-to-device copy of the action plus one cudaGraphLaunch.
static int step_impl(mcr_handle h, const void* action, bool action_on_host, int32_t action_dtype, uint8_t* obs, double* reward,
                     uint8_t* done, int32_t flags, void* stream, const HostOut* ho);

extern "C" int mcr_step(mcr_handle h, const void* action, int32_t action_dtype, uint8_t* obs, double* reward,
                        uint8_t* done, int32_t flags, void* stream) {
    return step_impl(h, action, false, action_dtype, obs, reward, done, flags, stream, nullptr);
}

extern "C" int mcr_step_host(mcr_handle h, const void* h_action, int32_t action_dtype, uint8_t* d_obs, double* d_reward, uint8_t* d_done,
                             uint8_t* h_obs, double* h_reward, uint8_t* h_done, int32_t flags, void* stream) {
    if (!h_obs || !h_reward || !h_done) return fail(-1, "mcr_step_host: null host buffer");
    const HostOut ho{h_obs, h_reward, h_done};
    return step_impl(h, h_action, true, action_dtype, d_obs, d_reward, d_done, flags, stream, &ho);
}

static int step_impl(mcr_handle h, const void* action, bool action_on_host, int32_t action_dtype, uint8_t* obs, double* reward,
                     uint8_t* done, int32_t flags, void* stream, const HostOut* ho) {
    int rc = check_bound(h); if (rc) return rc;
    if (!action || !obs || !reward || !done) return fail(-1, "mcr_step: null argument");
    if (action_dtype != MCR_F32 && action_dtype != MCR_F64) return fail(-1, "action dtype must be MCR_F32 or MCR_F64");
    if ((flags & 3) == 3) return fail(-1, "mcr_step: flags bit0 (same-step) and bit1 (next-step) auto reset are exclusive");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t abytes = (size_t)h->d.N * 3 * (action_dtype == MCR_F64 ? 8 : 4);
    if (action_on_host) {
        // the action goes straight into the library's staging buffer (what the kernels and the captured graph read)
        CUDA_OK(cudaMemcpyAsync(h->buf.action_stage, action, abytes, cudaMemcpyHostToDevice, s));
        action = h->buf.action_stage;
    }
    const bool refill_now = h->cfg.fresh_tracks > 0 && (flags & 3) && ++h->steps_since_refill >= 4;
    if (!h->use_graphs || h->eager_steps < 2) {
        ++h->eager_steps;
        rc = step_enqueue(h, action, action_dtype, obs, reward, done, flags, stream, ho);
        if (!rc && refill_now) rc = kick_refill(h, nullptr);
        return rc;
    }
    mcr_handle_t::StepGraph* g = nullptr;
    for (auto& c : h->graphs)
        if (c.dtype == action_dtype && c.flags == flags && c.obs == obs && c.reward == reward && c.done == done &&
            c.h_obs == (ho ? ho->h_obs : nullptr) && c.h_reward == (ho ? ho->h_reward : nullptr) && c.h_done == (ho ? ho->h_done : nullptr)) { g = &c; break; }
    if (!g) {
        if (h->graphs.size() >= 16) { destroy_step_graph(h->graphs.front()); h->graphs.erase(h->graphs.begin()); }
        rc = ensure_side(h); if (rc) return rc;
        const int64_t before = h->launches;
        CUDA_OK(cudaStreamBeginCapture(h->cap, cudaStreamCaptureModeRelaxed));
        rc = step_enqueue(h, h->buf.action_stage, action_dtype, obs, reward, done, flags, h->cap, ho);
        cudaGraph_t graph = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(h->cap, &graph);
        const int64_t per_step = h->launches - before;
        h->launches = before;
        cudaGraphExec_t exec = nullptr;
        cudaError_t ie = cudaErrorUnknown;
        if (!rc && ce == cudaSuccess && graph) ie = cudaGraphInstantiate(&exec, graph, 0);
        if (ie != cudaSuccess && graph) { cudaGraphDestroy(graph); graph = nullptr; }
        if (ie != cudaSuccess) {
            // graphs are an optimisation of the launch path only: issue this and all later steps directly
            (void)cudaGetLastError();
            h->use_graphs = false;
            rc = step_enqueue(h, action, action_dtype, obs, reward, done, flags, stream, ho);
            if (!rc && refill_now) rc = kick_refill(h, nullptr);
            return rc;
        }
        h->graphs.push_back(mcr_handle_t::StepGraph{action_dtype, flags, obs, reward, done, ho ? ho->h_obs : nullptr,
                                                    ho ? ho->h_reward : nullptr, ho ? ho->h_done : nullptr, exec, per_step,
                                                    graph, {}, (const void*)h->buf.action_stage});
        g = &h->graphs.back();
        static const bool no_patch = std::getenv("MCR_ACTION_COPY") != nullptr;      // A/B: the device-to-device copy instead
        if (!no_patch) find_action_nodes(graph, h->buf.action_stage, g->action_nodes);
    }
    if (!g->action_nodes.empty()) {
        rc = patch_action(*g, action); if (rc) return rc;
    } else if (action != (const void*)h->buf.action_stage) {
        CUDA_OK(cudaMemcpyAsync(h->buf.action_stage, action, abytes, cudaMemcpyDeviceToDevice, s));
    }
    CUDA_OK(cudaGraphLaunch(g->exec, s));
    h->launches += g->launches;
    if (refill_now) { rc = kick_refill(h, nullptr); if (rc) return rc; }
    return 0;
}
