/*
 * mcr_oracle.c -- CPU ORACLE for the MultiCarRacing-v0 step+render hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (multi_car_racing_b200/) never imports, links or calls anything in oracle/.
 *
 * What it restates (reference = /root/reference/gym_multi_car_racing/multi_car_racing.py,
 * "mcr" below):
 *   - MultiCarRacing.step                       mcr:410-509
 *   - FrictionDetector._contact                 mcr:88-123
 *   - _render_window / render_road / render_indicators   mcr:520-604, 613-674
 * and the un-vendored third-party arithmetic those lines call into (NOT present under
 * /root/reference; restated from the published algorithms):
 *   - gym 0.17.2  gym/envs/box2d/car_dynamics.py  (Car.__init__/gas/brake/steer/step/draw)
 *   - box2d-py 2.3.5 (Box2D 2.3.x C++): b2PolygonShape::Set/ComputeMass, b2Body mass data,
 *     b2World::Step -> Collide (sensor overlap) / b2Island::Solve / b2RevoluteJoint
 *     (limit + motor), sleeping.
 *   - gym rendering.Transform + OpenGL fixed function point-sampled polygon fill.
 *
 * PARITY STATUS: "parity unpinned" for the third-party restatements (no Box2D / gym /
 * pyglet install exists in the build container and the reference ships no tests or golden
 * vectors).  The reference's OWN code (reward rule, spawn grid, draw lists, HUD geometry,
 * track generator) is pinned by tests/golden/ fixtures produced by executing the unmodified
 * reference module on top of stub third-party modules (tests/golden/make_golden.py).
 *
 * Numerics: all rigid-body math is IEEE fp32 without contraction (compile with
 * -ffp-contract=off), tyre model / reward / camera math is fp64 exactly where the reference
 * uses Python floats.  b2Rot::Set's sinf/cosf are restated as (float)sin((double)a) so that
 * the CPU and the CUDA path share one definition (differs from glibc sinf by <= 1 ulp).
 *
 * Documented deviations (also in DESIGN.md):
 *   D1  same-step shared-tile tie break: events are processed tile-descending, then
 *       car-id descending (what a fresh b2World yields at spawn).
 *   D2  b2TestOverlap (GJK distance < rA+rB) restated as SAT-intersect OR
 *       min vertex/edge distance <= 0.02 (same predicate, different fp32 rounding).
 *   D3  score label glyphs: baked 3x5 digit font (reference uses the platform font).
 *   D4  car-car rigid contacts (b2CollidePolygons + b2ContactSolver) follow Box2D 2.3.1+'s brute-force
 *       b2FindMaxSeparation; manifold order = (car a < b, fixture a, fixture b) ascending.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#if defined(__GNUC__)
#define ORC_API __attribute__((visibility("default")))
#else
#define ORC_API
#endif

/* ------------------------------------------------------------------------------------ */
/* Box2D common constants (b2Settings.h, 2.3.x)                                          */
/* ------------------------------------------------------------------------------------ */
#define B2_PI 3.14159265359f
#define B2_LINEAR_SLOP 0.005f
#define B2_ANGULAR_SLOP (2.0f / 180.0f * B2_PI)
#define B2_POLYGON_RADIUS (2.0f * B2_LINEAR_SLOP)
#define B2_MAX_LINEAR_CORRECTION 0.2f
#define B2_MAX_ANGULAR_CORRECTION (8.0f / 180.0f * B2_PI)
#define B2_MAX_TRANSLATION 2.0f
#define B2_MAX_TRANSLATION_SQ (B2_MAX_TRANSLATION * B2_MAX_TRANSLATION)
#define B2_MAX_ROTATION (0.5f * B2_PI)
#define B2_MAX_ROTATION_SQ (B2_MAX_ROTATION * B2_MAX_ROTATION)
#define B2_BAUMGARTE 0.2f
#define B2_TIME_TO_SLEEP 0.5f
#define B2_LINEAR_SLEEP_TOL 0.01f
#define B2_ANGULAR_SLEEP_TOL (2.0f / 180.0f * B2_PI)
#define B2_EPSILON FLT_EPSILON
#define B2_VELOCITY_THRESHOLD 1.0f
#define B2_MAX_MANIFOLD_POINTS 2

#define MAXV 8 /* hull poly 3 has 8 vertices; pybox2d allows 16 */

typedef struct { float x, y; } V2;
typedef struct { float s, c; } Rot;

static inline V2 v2(float x, float y) { V2 r; r.x = x; r.y = y; return r; }
static inline V2 vadd(V2 a, V2 b) { return v2(a.x + b.x, a.y + b.y); }
static inline V2 vsub(V2 a, V2 b) { return v2(a.x - b.x, a.y - b.y); }
static inline V2 vscale(float s, V2 a) { return v2(s * a.x, s * a.y); }
static inline V2 vneg(V2 a) { return v2(-a.x, -a.y); }
static inline float vdot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
static inline float vcross(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
static inline V2 cross_sv(float s, V2 a) { return v2(-s * a.y, s * a.x); } /* b2Cross(s, v) */
static inline V2 cross_vs(V2 a, float s) { return v2(s * a.y, -s * a.x); } /* b2Cross(v, s) */
static inline float vlen(V2 a) { return sqrtf(a.x * a.x + a.y * a.y); }
static inline float clampf(float a, float lo, float hi) { return fmaxf(lo, fminf(a, hi)); }
/* b2Rot::Set -- see header note on sin/cos */
static inline Rot rot_set(float a) { Rot q; q.s = (float)sin((double)a); q.c = (float)cos((double)a); return q; }
static inline V2 rmul(Rot q, V2 v) { return v2(q.c * v.x - q.s * v.y, q.s * v.x + q.c * v.y); }
static inline V2 rmulT(Rot q, V2 v) { return v2(q.c * v.x + q.s * v.y, -q.s * v.x + q.c * v.y); }
static inline V2 xmul(V2 p, Rot q, V2 v) { /* b2Mul(b2Transform, v) */
    return v2((q.c * v.x - q.s * v.y) + p.x, (q.s * v.x + q.c * v.y) + p.y);
}

/* ------------------------------------------------------------------------------------ */
/* b2PolygonShape::Set (welding, gift-wrapped hull, normals) and ComputeMass             */
/* ------------------------------------------------------------------------------------ */
typedef struct { int n; V2 v[MAXV]; V2 nrm[MAXV]; V2 centroid; } Poly;

static void poly_set(Poly* P, const V2* in, int count) {
    V2 ps[MAXV]; int n = 0;
    if (count > MAXV) count = MAXV;
    for (int i = 0; i < count; ++i) {
        int unique = 1;
        for (int j = 0; j < n; ++j) {
            V2 d = vsub(in[i], ps[j]);
            if (vdot(d, d) < 0.5f * B2_LINEAR_SLOP) { unique = 0; break; }
        }
        if (unique) ps[n++] = in[i];
    }
    if (n < 3) { /* degenerate; Box2D falls back to a unit box -- never happens here */
        P->n = 4; P->v[0] = v2(-1, -1); P->v[1] = v2(1, -1); P->v[2] = v2(1, 1); P->v[3] = v2(-1, 1);
    } else {
        int i0 = 0; float x0 = ps[0].x;
        for (int i = 1; i < n; ++i) {
            float x = ps[i].x;
            if (x > x0 || (x == x0 && ps[i].y < ps[i0].y)) { i0 = i; x0 = x; }
        }
        int hull[MAXV]; int m = 0; int ih = i0;
        for (;;) {
            hull[m] = ih;
            int ie = 0;
            for (int j = 1; j < n; ++j) {
                if (ie == ih) { ie = j; continue; }
                V2 r = vsub(ps[ie], ps[hull[m]]);
                V2 v = vsub(ps[j], ps[hull[m]]);
                float c = vcross(r, v);
                if (c < 0.0f) ie = j;
                if (c == 0.0f && vdot(v, v) > vdot(r, r)) ie = j;
            }
            ++m; ih = ie;
            if (ie == i0) break;
        }
        P->n = m;
        for (int i = 0; i < m; ++i) P->v[i] = ps[hull[i]];
    }
    for (int i = 0; i < P->n; ++i) {
        int i2 = i + 1 < P->n ? i + 1 : 0;
        V2 e = vsub(P->v[i2], P->v[i]);
        V2 nn = cross_vs(e, 1.0f);
        float len = vlen(nn);
        if (len >= B2_EPSILON) { float inv = 1.0f / len; nn.x *= inv; nn.y *= inv; }
        P->nrm[i] = nn;
    }
    /* ComputeCentroid */
    {
        V2 c = v2(0.0f, 0.0f); float area = 0.0f; V2 pRef = v2(0.0f, 0.0f);
        const float inv3 = 1.0f / 3.0f;
        for (int i = 0; i < P->n; ++i) {
            V2 p1 = pRef, p2 = P->v[i], p3 = i + 1 < P->n ? P->v[i + 1] : P->v[0];
            V2 e1 = vsub(p2, p1), e2 = vsub(p3, p1);
            float D = vcross(e1, e2);
            float ta = 0.5f * D;
            area += ta;
            c.x += ta * inv3 * (p1.x + p2.x + p3.x);
            c.y += ta * inv3 * (p1.y + p2.y + p3.y);
        }
        c.x *= 1.0f / area; c.y *= 1.0f / area;
        P->centroid = c;
    }
}

typedef struct { float mass; V2 center; float I; } MassData;

static void poly_mass(const Poly* P, float density, MassData* md) {
    V2 center = v2(0.0f, 0.0f); float area = 0.0f, I = 0.0f;
    V2 s = v2(0.0f, 0.0f);
    for (int i = 0; i < P->n; ++i) s = vadd(s, P->v[i]);
    s = vscale(1.0f / P->n, s);
    const float k_inv3 = 1.0f / 3.0f;
    for (int i = 0; i < P->n; ++i) {
        V2 e1 = vsub(P->v[i], s);
        V2 e2 = i + 1 < P->n ? vsub(P->v[i + 1], s) : vsub(P->v[0], s);
        float D = vcross(e1, e2);
        float ta = 0.5f * D;
        area += ta;
        center = vadd(center, vscale(ta * k_inv3, vadd(e1, e2)));
        float ex1 = e1.x, ey1 = e1.y, ex2 = e2.x, ey2 = e2.y;
        float intx2 = ex1 * ex1 + ex2 * ex1 + ex2 * ex2;
        float inty2 = ey1 * ey1 + ey2 * ey1 + ey2 * ey2;
        I += (0.25f * k_inv3 * D) * (intx2 + inty2);
    }
    md->mass = density * area;
    center = vscale(1.0f / area, center);
    md->center = vadd(center, s);
    md->I = density * I;
    md->I += md->mass * (vdot(md->center, md->center) - vdot(center, center));
}

/* ------------------------------------------------------------------------------------ */
/* car_dynamics (gym 0.17.2) constants; arithmetic in the same order as the Python source */
/* ------------------------------------------------------------------------------------ */
static const double SIZE = 0.02;
#define ENGINE_POWER (100000000 * SIZE * SIZE)
#define WHEEL_MOMENT_OF_INERTIA (4000 * SIZE * SIZE)
#define FRICTION_LIMIT (1000000 * SIZE * SIZE)
static const double WHEEL_R = 27, WHEEL_W = 14;
static const double WHEELPOS[4][2] = { {-55, +80}, {+55, +80}, {-55, -82}, {+55, -82} };
static const double HULL_POLY1[4][2] = { {-60, +130}, {+60, +130}, {+60, +110}, {-60, +110} };
static const double HULL_POLY2[4][2] = { {-15, +120}, {+15, +120}, {+20, +20}, {-20, 20} };
static const double HULL_POLY3[8][2] = { {+25, +20}, {+50, -10}, {+50, -40}, {+20, -90},
                                          {-20, -90}, {-50, -40}, {-50, -10}, {-25, +20} };
static const double HULL_POLY4[4][2] = { {-50, -120}, {+50, -120}, {+50, -90}, {-50, -90} };

typedef struct {
    V2 c; float a;          /* sweep: world centre of mass, angle */
    V2 v; float w;
    V2 force; float torque;
    V2 p; Rot q;            /* transform (body origin) */
    V2 localCenter;
    float mass, invMass, I, invI;
    float sleepTime; int awake;
} Body;

enum { LIM_INACTIVE = 0, LIM_LOWER = 1, LIM_UPPER = 2, LIM_EQUAL = 3 };

typedef struct {
    V2 localAnchorA;       /* on hull */
    float impulse[3]; float motorImpulse; int limitState;
    float motorSpeed, maxMotorTorque, lower, upper, refAngle;
    /* solver temporaries */
    V2 rA, rB; float K[3][3]; /* K[col][row] == b2Mat33 ex,ey,ez */
    float motorMass;
} Joint;

typedef struct {
    Body b[5];             /* 0 hull, 1..4 wheels (creation order, car_dynamics.Car.__init__) */
    Joint j[4];
    /* wheel python-side attributes (float64) */
    double gas[4], brake[4], steer[4], phase[4], omega[4];
    int ntiles[4];         /* len(wheel.tiles) */
    int hull_color;        /* palette index */
    /* skid traces (gym car_dynamics Car.step "Skid trace" block, Car._create_particle, Car.particles): at most
     * PRT_MAX particles per car in creation order, each a polyline of <= PRT_PTS wheel positions */
    int n_prt; uint32_t prt_next_id;
    struct { uint32_t id; int len; int grass; V2 pt[30]; } prt[30];
    /* per wheel: skid_start (b2Vec2 snapshot or None), skid_particle (id of the particle the wheel extends, its
     * length and grass flag are kept here too: a particle popped from Car.particles is still extended) */
    int skid_start_valid[4]; V2 skid_start[4];
    int skid_ref_valid[4]; uint32_t skid_ref_id[4]; int skid_ref_len[4], skid_ref_grass[4];
} Car;
#define PRT_MAX 30
#define PRT_PTS 30

#define MAX_AGENTS 16

/* palette (u8 rgb) */
typedef struct { uint8_t r, g, b; } RGB;
static inline uint8_t col_u8(float c) { return (uint8_t)(int)floorf(c * 255.0f + 0.5f); }
static inline RGB rgbf(float r, float g, float b) { RGB c; c.r = col_u8(r); c.g = col_u8(g); c.b = col_u8(b); return c; }

/* persistent car-car contact manifolds (b2Manifold + ids, for warm starting) */
typedef struct { V2 localPoint; uint32_t id; float normalImpulse, tangentImpulse; } MPoint;
typedef struct {
    int carA, fixA, carB, fixB;   /* fixture A belongs to the lower car index (lower proxy id on a fresh world) */
    int type;                     /* 0 = e_faceA, 1 = e_faceB */
    int pointCount;
    V2 localNormal, localPoint;
    MPoint pt[2];
} Manifold;
#define MAX_MANIFOLDS 64

typedef struct OrcWorld {
    int A;
    int T, Q;
    double* track;       /* T*4 alpha,beta,x,y */
    float* quad;         /* Q*8 fp32 render verts (as passed to glVertex3f) */
    double* quad64;      /* Q*8 float64 road_poly verts (what shapely's Polygon holds, mcr:336-337) */
    float* quad_rgb;     /* Q*3 float colours */
    int* quad_tile;      /* Q: tile index or -1 */
    Poly* tile_poly;     /* T Box2D polygon of each tile */
    float* tile_aabb;    /* T*4 lo.x lo.y hi.x hi.y */
    uint8_t* visited;    /* T*A road_visited */
    uint8_t* touched;    /* T colour has been reset to ROAD_COLOR */
    int cw;
    Car car[MAX_AGENTS];
    Poly hull_poly[4], wheel_poly;
    float hull_mass, hull_invMass, hull_I, hull_invI; V2 hull_lc;
    float wheel_mass, wheel_invMass, wheel_I, wheel_invI; V2 wheel_lc;
    double reward[MAX_AGENTS], prev_reward[MAX_AGENTS];
    int tile_visited_count[MAX_AGENTS];
    uint8_t driving_backward[MAX_AGENTS];
    uint8_t driving_on_grass[MAX_AGENTS];   /* mcr:153, 350, 469-472 */
    int draw_particles;  /* 1: the non-state render modes draw the skid traces (the reference always does; the CUDA
                            path keeps them on request only) */
    double t;
    float inv_dt0;
    /* config */
    double h_ratio; int backwards_flag, use_ego_color;
    int collisions;      /* car-car rigid contacts on/off */
    /* EXPOSURE VARIANTS (tests/test_oracle_exposure.py): what the two orders this restatement could not check against a
     * real Box2D would change if they were the other way round.  0 = the documented choice (D1). */
    int var_tie_ascending;    /* same-step shared-tile tie break: LOWER car id first */
    int var_joint_ascending;  /* island joint order [j0, j1, j2, j3] */
    Manifold manifolds[MAX_MANIFOLDS]; int n_manifolds; int manifold_overflow;
    int vel_iters_used;  /* diagnostics: last fixed-point iteration index */
} OrcWorld;

/* ------------------------------------------------------------------------------------ */
static void build_shapes(OrcWorld* W) {
    V2 tmp[MAXV];
    for (int i = 0; i < 4; ++i) tmp[i] = v2((float)(HULL_POLY1[i][0] * SIZE), (float)(HULL_POLY1[i][1] * SIZE));
    poly_set(&W->hull_poly[0], tmp, 4);
    for (int i = 0; i < 4; ++i) tmp[i] = v2((float)(HULL_POLY2[i][0] * SIZE), (float)(HULL_POLY2[i][1] * SIZE));
    poly_set(&W->hull_poly[1], tmp, 4);
    for (int i = 0; i < 8; ++i) tmp[i] = v2((float)(HULL_POLY3[i][0] * SIZE), (float)(HULL_POLY3[i][1] * SIZE));
    poly_set(&W->hull_poly[2], tmp, 8);
    for (int i = 0; i < 4; ++i) tmp[i] = v2((float)(HULL_POLY4[i][0] * SIZE), (float)(HULL_POLY4[i][1] * SIZE));
    poly_set(&W->hull_poly[3], tmp, 4);
    /* wheel: front_k = 1.0 for all wheels in gym 0.17.2 */
    double front_k = 1.0;
    double wp[4][2] = { {-WHEEL_W, +WHEEL_R}, {+WHEEL_W, +WHEEL_R}, {+WHEEL_W, -WHEEL_R}, {-WHEEL_W, -WHEEL_R} };
    for (int i = 0; i < 4; ++i) tmp[i] = v2((float)(wp[i][0] * front_k * SIZE), (float)(wp[i][1] * front_k * SIZE));
    poly_set(&W->wheel_poly, tmp, 4);

    /* b2Body::ResetMassData: fixtures iterate in reverse creation order */
    {
        float mass = 0.0f, I = 0.0f; V2 lc = v2(0.0f, 0.0f);
        for (int f = 3; f >= 0; --f) {
            MassData md; poly_mass(&W->hull_poly[f], 1.0f, &md);
            mass += md.mass;
            lc = vadd(lc, vscale(md.mass, md.center));
            I += md.I;
        }
        float inv = 1.0f / mass;
        lc = vscale(inv, lc);
        I -= mass * vdot(lc, lc);
        W->hull_mass = mass; W->hull_invMass = inv; W->hull_I = I; W->hull_invI = 1.0f / I; W->hull_lc = lc;
    }
    {
        MassData md; poly_mass(&W->wheel_poly, 0.1f, &md);
        float mass = md.mass; V2 lc = vscale(md.mass, md.center); float I = md.I;
        float inv = 1.0f / mass;
        lc = vscale(inv, lc);
        I -= mass * vdot(lc, lc);
        W->wheel_mass = mass; W->wheel_invMass = inv; W->wheel_I = I; W->wheel_invI = 1.0f / I; W->wheel_lc = lc;
    }
}

ORC_API OrcWorld* orc_create(int num_agents, double h_ratio, int backwards_flag, int use_ego_color) {
    if (num_agents < 1 || num_agents > MAX_AGENTS) return NULL;
    OrcWorld* W = (OrcWorld*)calloc(1, sizeof(OrcWorld));
    W->A = num_agents; W->h_ratio = h_ratio; W->backwards_flag = backwards_flag; W->use_ego_color = use_ego_color;
    W->collisions = 1;
    build_shapes(W);
    return W;
}

static void free_track(OrcWorld* W) {
    free(W->track); free(W->quad); free(W->quad64); free(W->quad_rgb); free(W->quad_tile); free(W->tile_poly);
    free(W->tile_aabb); free(W->visited); free(W->touched);
    W->track = NULL; W->quad = NULL; W->quad64 = NULL; W->quad_rgb = NULL; W->quad_tile = NULL; W->tile_poly = NULL;
    W->tile_aabb = NULL; W->visited = NULL; W->touched = NULL;
}

ORC_API void orc_destroy(OrcWorld* W) { if (!W) return; free_track(W); free(W); }

ORC_API void orc_set_collisions(OrcWorld* W, int on) { W->collisions = on; }
/* Vertex count b2PolygonShape::Set leaves of a tile quad given as 8 floats (fd_tile.shape.vertices = ..., mcr:318). */
ORC_API int orc_tile_vertex_count(const float* xy8) {
    V2 vv[4], ps[4]; Poly P; int n = 0;
    for (int k = 0; k < 4; ++k) vv[k] = v2(xy8[2 * k], xy8[2 * k + 1]);
    for (int i = 0; i < 4; ++i) {               /* the weld step of Set: fewer than 3 vertices left -> b2Assert in Box2D */
        int unique = 1;
        for (int j = 0; j < n; ++j) { V2 d = vsub(vv[i], ps[j]); if (vdot(d, d) < 0.5f * B2_LINEAR_SLOP) { unique = 0; break; } }
        if (unique) ps[n++] = vv[i];
    }
    if (n < 3) return n;
    poly_set(&P, vv, 4);
    return P.n;
}
ORC_API void orc_set_variant(OrcWorld* W, int tie_ascending, int joint_ascending) {
    W->var_tie_ascending = tie_ascending; W->var_joint_ascending = joint_ascending;
}

/* Load a track: what _create_track leaves behind (mcr:310-335).  quad_verts are the float64
 * vertices of road_poly in draw order, quad_tile[q] = tile index for road quads, -1 for
 * red/white border quads. */
ORC_API int orc_set_track(OrcWorld* W, int T, const double* track_abxy, int Q, const double* quad_verts,
                          const float* quad_rgb, const int* quad_tile, int cw) {
    free_track(W);
    W->T = T; W->Q = Q; W->cw = cw;
    W->track = (double*)malloc(sizeof(double) * 4 * T); memcpy(W->track, track_abxy, sizeof(double) * 4 * T);
    W->quad = (float*)malloc(sizeof(float) * 8 * Q);
    for (int i = 0; i < 8 * Q; ++i) W->quad[i] = (float)quad_verts[i];
    W->quad64 = (double*)malloc(sizeof(double) * 8 * Q); memcpy(W->quad64, quad_verts, sizeof(double) * 8 * Q);
    W->quad_rgb = (float*)malloc(sizeof(float) * 3 * Q); memcpy(W->quad_rgb, quad_rgb, sizeof(float) * 3 * Q);
    W->quad_tile = (int*)malloc(sizeof(int) * Q); memcpy(W->quad_tile, quad_tile, sizeof(int) * Q);
    W->tile_poly = (Poly*)calloc(T, sizeof(Poly));
    W->tile_aabb = (float*)malloc(sizeof(float) * 4 * T);
    W->visited = (uint8_t*)calloc((size_t)T * W->A, 1);
    W->touched = (uint8_t*)calloc(T, 1);
    int nt = 0;
    for (int q = 0; q < Q; ++q) {
        int t = quad_tile[q];
        if (t < 0) continue;
        if (t >= T) return -1;
        V2 vv[4];
        for (int k = 0; k < 4; ++k) vv[k] = v2(W->quad[8 * q + 2 * k], W->quad[8 * q + 2 * k + 1]);
        poly_set(&W->tile_poly[t], vv, 4);
        const Poly* P = &W->tile_poly[t];
        float lx = P->v[0].x, ly = P->v[0].y, hx = lx, hy = ly;
        for (int k = 1; k < P->n; ++k) {
            lx = fminf(lx, P->v[k].x); ly = fminf(ly, P->v[k].y);
            hx = fmaxf(hx, P->v[k].x); hy = fmaxf(hy, P->v[k].y);
        }
        W->tile_aabb[4 * t + 0] = lx; W->tile_aabb[4 * t + 1] = ly;
        W->tile_aabb[4 * t + 2] = hx; W->tile_aabb[4 * t + 3] = hy;
        ++nt;
    }
    return nt == T ? 0 : -2;
}

static void body_init(Body* b, float x, float y, float angle, float mass, float invMass, float I, float invI, V2 lc) {
    memset(b, 0, sizeof(*b));
    b->p = v2(x, y); b->q = rot_set(angle); b->a = angle;
    b->mass = mass; b->invMass = invMass; b->I = I; b->invI = invI; b->localCenter = lc;
    b->c = xmul(b->p, b->q, lc);
    b->awake = 1; b->sleepTime = 0.0f;
}

/* reset(): zero the bookkeeping (mcr:341-350) and create the cars (mcr:400-406 ->
 * car_dynamics.Car.__init__).  init = A x (angle, x, y) float64 as computed at mcr:384-393. */
ORC_API void orc_spawn(OrcWorld* W, const double* init) {
    memset(W->visited, 0, (size_t)W->T * W->A);
    memset(W->touched, 0, W->T);
    W->t = 0.0;
    W->n_manifolds = 0;
    for (int c = 0; c < W->A; ++c) {
        Car* car = &W->car[c];
        memset(car, 0, sizeof(*car));
        double ang = init[3 * c + 0], ix = init[3 * c + 1], iy = init[3 * c + 2];
        body_init(&car->b[0], (float)ix, (float)iy, (float)ang, W->hull_mass, W->hull_invMass, W->hull_I, W->hull_invI, W->hull_lc);
        for (int w = 0; w < 4; ++w) {
            double wx = WHEELPOS[w][0], wy = WHEELPOS[w][1];
            body_init(&car->b[1 + w], (float)(ix + wx * SIZE), (float)(iy + wy * SIZE), (float)ang,
                      W->wheel_mass, W->wheel_invMass, W->wheel_I, W->wheel_invI, W->wheel_lc);
            Joint* J = &car->j[w];
            J->localAnchorA = v2((float)(wx * SIZE), (float)(wy * SIZE));
            J->maxMotorTorque = (float)(180 * 900 * SIZE * SIZE);
            J->motorSpeed = 0.0f;
            J->lower = (float)-0.4; J->upper = (float)+0.4;
            J->refAngle = car->b[1 + w].a - car->b[0].a; /* pybox2d: bodyB.angle - bodyA.angle */
            J->limitState = LIM_INACTIVE;
        }
        car->hull_color = c % 8;
        W->reward[c] = 0.0; W->prev_reward[c] = 0.0; W->tile_visited_count[c] = 0; W->driving_backward[c] = 0; W->driving_on_grass[c] = 0;
    }
}

/* ------------------------------------------------------------------------------------ */
/* Car.gas / brake / steer and Car.step (gym 0.17.2 car_dynamics.py)                     */
/* ------------------------------------------------------------------------------------ */
static double sign_d(double x) { return x > 0 ? 1.0 : (x < 0 ? -1.0 : 0.0); }

static void car_controls(Car* car, double a0, double a1, double a2) {
    /* mcr:422-424: car.steer(-a0); car.gas(a1); car.brake(a2) */
    double s = -a0;
    car->steer[0] = s; car->steer[1] = s;
    double g = a1 < 0 ? 0 : (a1 > 1 ? 1 : a1);
    for (int w = 2; w < 4; ++w) {
        double diff = g - car->gas[w];
        if (diff > 0.1) diff = 0.1;
        car->gas[w] += diff;
    }
    for (int w = 0; w < 4; ++w) car->brake[w] = a2;
}

static void car_step(Car* car, double dt) {
    for (int w = 0; w < 4; ++w) {
        Body* wb = &car->b[1 + w]; Body* hb = &car->b[0]; Joint* J = &car->j[w];
        double jangle = (double)((wb->a - hb->a) - J->refAngle); /* b2RevoluteJoint::GetJointAngle, fp32 */
        double dir = sign_d(car->steer[w] - jangle);
        double val = fabs(car->steer[w] - jangle);
        J->motorSpeed = (float)(dir * fmin(50.0 * val, 3.0));

        double friction_limit = FRICTION_LIMIT * 0.6;
        if (car->ntiles[w] > 0) friction_limit = fmax(friction_limit, FRICTION_LIMIT * 1.0);

        /* GetWorldVector((0,1)) / ((1,0)) in fp32 */
        V2 forw = rmul(wb->q, v2(0.0f, 1.0f));
        V2 side = rmul(wb->q, v2(1.0f, 0.0f));
        double vx = wb->v.x, vy = wb->v.y;
        double vf = (double)forw.x * vx + (double)forw.y * vy;
        double vs = (double)side.x * vx + (double)side.y * vy;

        car->omega[w] += dt * ENGINE_POWER * car->gas[w] / WHEEL_MOMENT_OF_INERTIA / (fabs(car->omega[w]) + 5.0);
        if (car->brake[w] >= 0.9) {
            car->omega[w] = 0;
        } else if (car->brake[w] > 0) {
            double BRAKE_FORCE = 15;
            double d2 = -sign_d(car->omega[w]);
            double v2_ = BRAKE_FORCE * car->brake[w];
            if (fabs(v2_) > fabs(car->omega[w])) v2_ = fabs(car->omega[w]);
            car->omega[w] += d2 * v2_;
        }
        car->phase[w] += car->omega[w] * dt;

        double wheel_rad = 1.0 * WHEEL_R * SIZE;
        double vr = car->omega[w] * wheel_rad;
        double f_force = -vf + vr;
        double p_force = -vs;
        f_force *= 205000 * SIZE * SIZE;
        p_force *= 205000 * SIZE * SIZE;
        double force = sqrt(f_force * f_force + p_force * p_force);
        /* Skid trace (gym car_dynamics.Car.step) */
        {
            const int grass = car->ntiles[w] == 0;
            if (fabs(force) > 2.0 * friction_limit) {
                if (car->skid_ref_valid[w] && car->skid_ref_grass[w] == grass && car->skid_ref_len[w] < PRT_PTS) {
                    for (int i = 0; i < car->n_prt; ++i)          /* still in Car.particles: the drawn copy grows too */
                        if (car->prt[i].id == car->skid_ref_id[w]) { car->prt[i].pt[car->prt[i].len++] = wb->p; break; }
                    car->skid_ref_len[w] += 1;
                } else if (!car->skid_start_valid[w]) {
                    car->skid_start_valid[w] = 1; car->skid_start[w] = wb->p;
                } else {
                    /* _create_particle(skid_start, position, grass): append, then pop(0) while more than 30 */
                    if (car->n_prt == PRT_MAX) { memmove(&car->prt[0], &car->prt[1], sizeof(car->prt[0]) * (PRT_MAX - 1)); car->n_prt = PRT_MAX - 1; }
                    const int i = car->n_prt++;
                    car->prt[i].id = car->prt_next_id++; car->prt[i].len = 2; car->prt[i].grass = grass;
                    car->prt[i].pt[0] = car->skid_start[w]; car->prt[i].pt[1] = wb->p;
                    car->skid_ref_valid[w] = 1; car->skid_ref_id[w] = car->prt[i].id; car->skid_ref_len[w] = 2; car->skid_ref_grass[w] = grass;
                    car->skid_start_valid[w] = 0;
                }
            } else {
                car->skid_start_valid[w] = 0;
                car->skid_ref_valid[w] = 0;
            }
        }
        if (fabs(force) > friction_limit) {
            f_force /= force; p_force /= force;
            force = friction_limit;
            f_force *= force; p_force *= force;
        }
        car->omega[w] -= dt * f_force * wheel_rad / WHEEL_MOMENT_OF_INERTIA;
        /* ApplyForceToCenter(force, wake=True) */
        float fx = (float)(p_force * (double)side.x + f_force * (double)forw.x);
        float fy = (float)(p_force * (double)side.y + f_force * (double)forw.y);
        if (!wb->awake) { wb->awake = 1; wb->sleepTime = 0.0f; }
        wb->force.x += fx; wb->force.y += fy;
    }
}

/* ------------------------------------------------------------------------------------ */
/* Collide(): sensor contacts wheel/hull fixture vs tile (b2TestOverlap restated, D2)    */
/* ------------------------------------------------------------------------------------ */
static float seg_dist2(V2 p, V2 a, V2 b) {
    V2 e = vsub(b, a), w = vsub(p, a);
    float t = vdot(w, e);
    if (t <= 0.0f) return vdot(w, w);
    float l2 = vdot(e, e);
    if (t >= l2) { V2 w2 = vsub(p, b); return vdot(w2, w2); }
    float cr = vcross(e, w);
    return (cr * cr) / l2;
}

/* both polygons counter-clockwise, world coordinates */
static int poly_touch(const V2* a, int na, const V2* b, int nb) {
    /* SAT with un-normalised edge normals: is there a separating edge? */
    int separated = 0;
    for (int i = 0; i < na && !separated; ++i) {
        V2 p = a[i], e = vsub(a[i + 1 < na ? i + 1 : 0], p);
        int all_out = 1;
        for (int k = 0; k < nb; ++k) { if (!(vcross(e, vsub(b[k], p)) < 0.0f)) { all_out = 0; break; } }
        if (all_out) separated = 1;
    }
    for (int i = 0; i < nb && !separated; ++i) {
        V2 p = b[i], e = vsub(b[i + 1 < nb ? i + 1 : 0], p);
        int all_out = 1;
        for (int k = 0; k < na; ++k) { if (!(vcross(e, vsub(a[k], p)) < 0.0f)) { all_out = 0; break; } }
        if (all_out) separated = 1;
    }
    if (!separated) return 1;
    float d2 = FLT_MAX;
    for (int i = 0; i < na; ++i)
        for (int k = 0; k < nb; ++k) d2 = fminf(d2, seg_dist2(a[i], b[k], b[k + 1 < nb ? k + 1 : 0]));
    for (int k = 0; k < nb; ++k)
        for (int i = 0; i < na; ++i) d2 = fminf(d2, seg_dist2(b[k], a[i], a[i + 1 < na ? i + 1 : 0]));
    float d = sqrtf(d2);
    const float rr = B2_POLYGON_RADIUS + B2_POLYGON_RADIUS;
    /* b2Distance(useRadii): distance' = d > rr && d > eps ? d - rr : 0 ; touching = distance' < 10 eps */
    if (d > rr && d > B2_EPSILON) return (d - rr) < 10.0f * B2_EPSILON;
    return 1;
}

static void world_poly(const Body* b, const Poly* P, V2* out, float* aabb) {
    float lx = FLT_MAX, ly = FLT_MAX, hx = -FLT_MAX, hy = -FLT_MAX;
    for (int i = 0; i < P->n; ++i) {
        out[i] = xmul(b->p, b->q, P->v[i]);
        lx = fminf(lx, out[i].x); ly = fminf(ly, out[i].y); hx = fmaxf(hx, out[i].x); hy = fmaxf(hy, out[i].y);
    }
    aabb[0] = lx; aabb[1] = ly; aabb[2] = hx; aabb[3] = hy;
}

#define AABB_MARGIN 0.05f /* conservative pre-reject; exact predicate threshold is 0.02 */

/* One touching (tile, fixture) pair as b2Contact::Update would report it. */
typedef struct { int tile, car, fixture /* 0..3 wheels, 4..7 hull fixtures */, active; } Pair;
#define MAX_PAIRS 4096

/* All currently touching pairs in the contact-list order of a fresh world:
 * tile descending, then car descending, then fixture descending (D1). */
static int collect_pairs(OrcWorld* W, Pair* out, int max) {
    int A = W->A, T = W->T, n = 0;
    V2 wv[MAX_AGENTS][8][MAXV]; float wa[MAX_AGENTS][8][4]; int wn[MAX_AGENTS][8];
    int active[MAX_AGENTS][8];
    for (int c = 0; c < A; ++c) {
        Car* car = &W->car[c];
        for (int w = 0; w < 4; ++w) {
            world_poly(&car->b[1 + w], &W->wheel_poly, wv[c][w], wa[c][w]); wn[c][w] = W->wheel_poly.n;
            active[c][w] = car->b[1 + w].awake;
        }
        for (int f = 0; f < 4; ++f) {
            world_poly(&car->b[0], &W->hull_poly[f], wv[c][4 + f], wa[c][4 + f]); wn[c][4 + f] = W->hull_poly[f].n;
            active[c][4 + f] = car->b[0].awake;
        }
    }
    for (int t = T - 1; t >= 0; --t) {
        const float* ta = &W->tile_aabb[4 * t]; const Poly* TP = &W->tile_poly[t];
        for (int cc_ = A - 1; cc_ >= 0; --cc_) {
            const int c = W->var_tie_ascending ? A - 1 - cc_ : cc_;
            for (int f = 7; f >= 0; --f) {
                const float* fa = wa[c][f];
                if (fa[0] - ta[2] > AABB_MARGIN || fa[1] - ta[3] > AABB_MARGIN ||
                    ta[0] - fa[2] > AABB_MARGIN || ta[1] - fa[3] > AABB_MARGIN) continue;
                if (!poly_touch(TP->v, TP->n, wv[c][f], wn[c][f])) continue;
                if (n < max) { out[n].tile = t; out[n].car = c; out[n].fixture = f; out[n].active = active[c][f]; }
                ++n;
            }
        }
    }
    return n < max ? n : max;
}

/* b2ContactManager::Collide + FrictionDetector (mcr:88-123).  road_visited makes BeginContact
 * idempotent and tile colour only ever moves to ROAD_COLOR, so no per-pair state is kept. */
static void collide(OrcWorld* W) {
    int A = W->A, T = W->T;
    static __thread Pair pairs[MAX_PAIRS];
    int n = collect_pairs(W, pairs, MAX_PAIRS);
    int touching_now[MAX_AGENTS][4];
    for (int c = 0; c < A; ++c) for (int w = 0; w < 4; ++w) touching_now[c][w] = 0;
    for (int i = 0; i < n; ++i) {
        int t = pairs[i].tile, c = pairs[i].car, f = pairs[i].fixture;
        if (f < 4) touching_now[c][f] += 1;
        /* A sleeping body's contacts are not updated (b2ContactManager::Collide) */
        if (!pairs[i].active) continue;
        W->touched[t] = 1;                      /* mcr:102-104 */
        if (f >= 4) continue;                   /* hull: userData None, mcr:108 */
        if (!W->visited[(size_t)t * A + c]) {   /* mcr:113-120 */
            W->visited[(size_t)t * A + c] = 1;
            W->tile_visited_count[c] += 1;
            int past = -1;
            for (int k = 0; k < A; ++k) past += W->visited[(size_t)t * A + k];
            double reward_factor = 1 - ((double)past / (double)A);
            W->reward[c] += reward_factor * 1000.0 / (double)T;
        }
    }
    for (int c = 0; c < A; ++c)
        for (int w = 0; w < 4; ++w)
            if (W->car[c].b[1 + w].awake) W->car[c].ntiles[w] = touching_now[c][w];
}

/* ------------------------------------------------------------------------------------ */
/* b2RevoluteJoint (limit + motor), Box2D 2.3.x                                          */
/* ------------------------------------------------------------------------------------ */
static void solve22(float K[3][3], float bx, float by, float* ox, float* oy) {
    float a11 = K[0][0], a12 = K[1][0], a21 = K[0][1], a22 = K[1][1];
    float det = a11 * a22 - a12 * a21;
    if (det != 0.0f) det = 1.0f / det;
    *ox = det * (a22 * bx - a12 * by);
    *oy = det * (a11 * by - a21 * bx);
}

static void solve33(float K[3][3], const float b[3], float x[3]) {
    const float* ex = K[0]; const float* ey = K[1]; const float* ez = K[2];
    /* det = dot(ex, cross(ey, ez)) */
    float cx = ey[1] * ez[2] - ey[2] * ez[1], cy = ey[2] * ez[0] - ey[0] * ez[2], cz = ey[0] * ez[1] - ey[1] * ez[0];
    float det = ex[0] * cx + ex[1] * cy + ex[2] * cz;
    if (det != 0.0f) det = 1.0f / det;
    x[0] = det * (b[0] * cx + b[1] * cy + b[2] * cz);
    /* dot(ex, cross(b, ez)) */
    float dx = b[1] * ez[2] - b[2] * ez[1], dy = b[2] * ez[0] - b[0] * ez[2], dz = b[0] * ez[1] - b[1] * ez[0];
    x[1] = det * (ex[0] * dx + ex[1] * dy + ex[2] * dz);
    /* dot(ex, cross(ey, b)) */
    float fx = ey[1] * b[2] - ey[2] * b[1], fy = ey[2] * b[0] - ey[0] * b[2], fz = ey[0] * b[1] - ey[1] * b[0];
    x[2] = det * (ex[0] * fx + ex[1] * fy + ex[2] * fz);
}

static void joint_init(Joint* J, Body* A, Body* B, float dtRatio) {
    float aA = A->a, aB = B->a;
    V2 vA = A->v, vB = B->v; float wA = A->w, wB = B->w;
    Rot qA = rot_set(aA), qB = rot_set(aB);
    J->rA = rmul(qA, vsub(J->localAnchorA, A->localCenter));
    J->rB = rmul(qB, vsub(v2(0.0f, 0.0f), B->localCenter));
    float mA = A->invMass, mB = B->invMass, iA = A->invI, iB = B->invI;
    V2 rA = J->rA, rB = J->rB;
    J->K[0][0] = mA + mB + rA.y * rA.y * iA + rB.y * rB.y * iB;
    J->K[1][0] = -rA.y * rA.x * iA - rB.y * rB.x * iB;
    J->K[2][0] = -rA.y * iA - rB.y * iB;
    J->K[0][1] = J->K[1][0];
    J->K[1][1] = mA + mB + rA.x * rA.x * iA + rB.x * rB.x * iB;
    J->K[2][1] = rA.x * iA + rB.x * iB;
    J->K[0][2] = J->K[2][0];
    J->K[1][2] = J->K[2][1];
    J->K[2][2] = iA + iB;
    J->motorMass = iA + iB;
    if (J->motorMass > 0.0f) J->motorMass = 1.0f / J->motorMass;
    /* enableLimit, not fixedRotation */
    float jointAngle = aB - aA - J->refAngle;
    if (fabsf(J->upper - J->lower) < 2.0f * B2_ANGULAR_SLOP) {
        J->limitState = LIM_EQUAL;
    } else if (jointAngle <= J->lower) {
        if (J->limitState != LIM_LOWER) J->impulse[2] = 0.0f;
        J->limitState = LIM_LOWER;
    } else if (jointAngle >= J->upper) {
        if (J->limitState != LIM_UPPER) J->impulse[2] = 0.0f;
        J->limitState = LIM_UPPER;
    } else {
        J->limitState = LIM_INACTIVE;
        J->impulse[2] = 0.0f;
    }
    /* warm starting */
    J->impulse[0] *= dtRatio; J->impulse[1] *= dtRatio; J->impulse[2] *= dtRatio;
    J->motorImpulse *= dtRatio;
    V2 P = v2(J->impulse[0], J->impulse[1]);
    vA = vsub(vA, vscale(mA, P));
    wA -= iA * (vcross(rA, P) + J->motorImpulse + J->impulse[2]);
    vB = vadd(vB, vscale(mB, P));
    wB += iB * (vcross(rB, P) + J->motorImpulse + J->impulse[2]);
    A->v = vA; A->w = wA; B->v = vB; B->w = wB;
}

static void joint_solve_vel(Joint* J, Body* A, Body* B, float dt) {
    V2 vA = A->v, vB = B->v; float wA = A->w, wB = B->w;
    float mA = A->invMass, mB = B->invMass, iA = A->invI, iB = B->invI;
    /* motor */
    if (J->limitState != LIM_EQUAL) {
        float Cdot = wB - wA - J->motorSpeed;
        float impulse = -J->motorMass * Cdot;
        float oldImpulse = J->motorImpulse;
        float maxImpulse = dt * J->maxMotorTorque;
        J->motorImpulse = clampf(J->motorImpulse + impulse, -maxImpulse, maxImpulse);
        impulse = J->motorImpulse - oldImpulse;
        wA -= iA * impulse;
        wB += iB * impulse;
    }
    if (J->limitState != LIM_INACTIVE) {
        V2 Cdot1 = vsub(vsub(vadd(vB, cross_sv(wB, J->rB)), vA), cross_sv(wA, J->rA));
        float Cdot2 = wB - wA;
        float Cd[3] = { Cdot1.x, Cdot1.y, Cdot2 }, imp[3];
        solve33(J->K, Cd, imp);
        imp[0] = -imp[0]; imp[1] = -imp[1]; imp[2] = -imp[2];
        if (J->limitState == LIM_EQUAL) {
            J->impulse[0] += imp[0]; J->impulse[1] += imp[1]; J->impulse[2] += imp[2];
        } else if (J->limitState == LIM_LOWER) {
            float newImpulse = J->impulse[2] + imp[2];
            if (newImpulse < 0.0f) {
                float rx = -Cdot1.x + J->impulse[2] * J->K[2][0], ry = -Cdot1.y + J->impulse[2] * J->K[2][1];
                float redx, redy; solve22(J->K, rx, ry, &redx, &redy);
                imp[0] = redx; imp[1] = redy; imp[2] = -J->impulse[2];
                J->impulse[0] += redx; J->impulse[1] += redy; J->impulse[2] = 0.0f;
            } else {
                J->impulse[0] += imp[0]; J->impulse[1] += imp[1]; J->impulse[2] += imp[2];
            }
        } else { /* upper */
            float newImpulse = J->impulse[2] + imp[2];
            if (newImpulse > 0.0f) {
                float rx = -Cdot1.x + J->impulse[2] * J->K[2][0], ry = -Cdot1.y + J->impulse[2] * J->K[2][1];
                float redx, redy; solve22(J->K, rx, ry, &redx, &redy);
                imp[0] = redx; imp[1] = redy; imp[2] = -J->impulse[2];
                J->impulse[0] += redx; J->impulse[1] += redy; J->impulse[2] = 0.0f;
            } else {
                J->impulse[0] += imp[0]; J->impulse[1] += imp[1]; J->impulse[2] += imp[2];
            }
        }
        V2 P = v2(imp[0], imp[1]);
        vA = vsub(vA, vscale(mA, P));
        wA -= iA * (vcross(J->rA, P) + imp[2]);
        vB = vadd(vB, vscale(mB, P));
        wB += iB * (vcross(J->rB, P) + imp[2]);
    } else {
        V2 Cdot = vsub(vsub(vadd(vB, cross_sv(wB, J->rB)), vA), cross_sv(wA, J->rA));
        float ix, iy; solve22(J->K, -Cdot.x, -Cdot.y, &ix, &iy);
        J->impulse[0] += ix; J->impulse[1] += iy;
        V2 P = v2(ix, iy);
        vA = vsub(vA, vscale(mA, P));
        wA -= iA * vcross(J->rA, P);
        vB = vadd(vB, vscale(mB, P));
        wB += iB * vcross(J->rB, P);
    }
    A->v = vA; A->w = wA; B->v = vB; B->w = wB;
}

static int joint_solve_pos(Joint* J, Body* A, Body* B) {
    V2 cA = A->c, cB = B->c; float aA = A->a, aB = B->a;
    float angularError = 0.0f, positionError = 0.0f;
    if (J->limitState != LIM_INACTIVE) {
        float angle = aB - aA - J->refAngle;
        float limitImpulse = 0.0f;
        if (J->limitState == LIM_EQUAL) {
            float C = clampf(angle - J->lower, -B2_MAX_ANGULAR_CORRECTION, B2_MAX_ANGULAR_CORRECTION);
            limitImpulse = -J->motorMass * C;
            angularError = fabsf(C);
        } else if (J->limitState == LIM_LOWER) {
            float C = angle - J->lower;
            angularError = -C;
            C = clampf(C + B2_ANGULAR_SLOP, -B2_MAX_ANGULAR_CORRECTION, 0.0f);
            limitImpulse = -J->motorMass * C;
        } else {
            float C = angle - J->upper;
            angularError = C;
            C = clampf(C - B2_ANGULAR_SLOP, 0.0f, B2_MAX_ANGULAR_CORRECTION);
            limitImpulse = -J->motorMass * C;
        }
        aA -= A->invI * limitImpulse;
        aB += B->invI * limitImpulse;
    }
    {
        Rot qA = rot_set(aA), qB = rot_set(aB);
        V2 rA = rmul(qA, vsub(J->localAnchorA, A->localCenter));
        V2 rB = rmul(qB, vsub(v2(0.0f, 0.0f), B->localCenter));
        V2 C = vsub(vsub(vadd(cB, rB), cA), rA);
        positionError = vlen(C);
        float mA = A->invMass, mB = B->invMass, iA = A->invI, iB = B->invI;
        float K[3][3];
        K[0][0] = mA + mB + iA * rA.y * rA.y + iB * rB.y * rB.y;
        K[0][1] = -iA * rA.x * rA.y - iB * rB.x * rB.y;
        K[1][0] = K[0][1];
        K[1][1] = mA + mB + iA * rA.x * rA.x + iB * rB.x * rB.x;
        float ix, iy; solve22(K, C.x, C.y, &ix, &iy);
        ix = -ix; iy = -iy;
        V2 imp = v2(ix, iy);
        cA = vsub(cA, vscale(mA, imp));
        aA -= iA * vcross(rA, imp);
        cB = vadd(cB, vscale(mB, imp));
        aB += iB * vcross(rB, imp);
    }
    A->c = cA; A->a = aA; B->c = cB; B->a = aB;
    return positionError <= B2_LINEAR_SLOP && angularError <= B2_ANGULAR_SLOP;
}

/* ------------------------------------------------------------------------------------ */
/* b2Island::Solve for one car island.  Island joint order for a Car is [j3, j2, j1, j0]  */
/* (joint edges are prepended).                                               */
/* ------------------------------------------------------------------------------------ */
static const int JOINT_ORDER[4] = { 3, 2, 1, 0 };

static int state_equal(const Car* a, const Car* b) {
    for (int i = 0; i < 5; ++i) {
        if (memcmp(&a->b[i].v, &b->b[i].v, sizeof(V2)) || memcmp(&a->b[i].w, &b->b[i].w, sizeof(float))) return 0;
    }
    for (int i = 0; i < 4; ++i) {
        if (memcmp(a->j[i].impulse, b->j[i].impulse, sizeof(float) * 3) ||
            memcmp(&a->j[i].motorImpulse, &b->j[i].motorImpulse, sizeof(float))) return 0;
    }
    return 1;
}

/* ------------------------------------------------------------------------------------ */
/* Car-car rigid contacts: b2CollidePolygons + b2ContactSolver (Box2D 2.3.x)              */
/* Filtering (car_dynamics): hull cat 0x0001 mask 0xFFFF, wheel cat 0x0020 mask 0x0001 ->  */
/* hull-hull and wheel-(other car's) hull collide, wheel-wheel does not.                    */
/* ------------------------------------------------------------------------------------ */
static const Poly* fixture_poly(const OrcWorld* W, int f) { return f < 4 ? &W->wheel_poly : &W->hull_poly[f - 4]; }
static Body* fixture_body(OrcWorld* W, int car, int f) { return f < 4 ? &W->car[car].b[1 + f] : &W->car[car].b[0]; }

typedef struct { V2 v; uint32_t id; } ClipVertex;
static uint32_t cf_key(int indexA, int indexB, int typeA, int typeB) {
    return (uint32_t)(indexA & 0xff) | ((uint32_t)(indexB & 0xff) << 8) | ((uint32_t)(typeA & 0xff) << 16) | ((uint32_t)(typeB & 0xff) << 24);
}
enum { CF_VERTEX = 0, CF_FACE = 1 };
static inline V2 xmulT(V2 p, Rot q, V2 v) { /* b2MulT(b2Transform, v) */
    float px = v.x - p.x, py = v.y - p.y;
    return v2(q.c * px + q.s * py, -q.s * px + q.c * py);
}

/* b2FindMaxSeparation (brute-force version of Box2D 2.3.1+) */
static float find_max_separation(int* edgeIndex, const Poly* poly1, V2 p1, Rot q1, const Poly* poly2, V2 p2, Rot q2) {
    /* xf = b2MulT(xf2, xf1) */
    Rot q; q.s = q2.c * q1.s - q2.s * q1.c; q.c = q2.c * q1.c + q2.s * q1.s;
    V2 p = rmulT(q2, vsub(p1, p2));
    int bestIndex = 0; float maxSeparation = -FLT_MAX;
    for (int i = 0; i < poly1->n; ++i) {
        V2 n = rmul(q, poly1->nrm[i]);
        V2 v1 = xmul(p, q, poly1->v[i]);
        float si = FLT_MAX;
        for (int j = 0; j < poly2->n; ++j) {
            float sij = vdot(n, vsub(poly2->v[j], v1));
            if (sij < si) si = sij;
        }
        if (si > maxSeparation) { maxSeparation = si; bestIndex = i; }
    }
    *edgeIndex = bestIndex;
    return maxSeparation;
}

static void find_incident_edge(ClipVertex c[2], const Poly* poly1, Rot q1, int edge1, const Poly* poly2, V2 p2, Rot q2) {
    V2 normal1 = rmulT(q2, rmul(q1, poly1->nrm[edge1]));
    int index = 0; float minDot = FLT_MAX;
    for (int i = 0; i < poly2->n; ++i) {
        float d = vdot(normal1, poly2->nrm[i]);
        if (d < minDot) { minDot = d; index = i; }
    }
    int i1 = index, i2 = i1 + 1 < poly2->n ? i1 + 1 : 0;
    c[0].v = xmul(p2, q2, poly2->v[i1]); c[0].id = cf_key(edge1, i1, CF_FACE, CF_VERTEX);
    c[1].v = xmul(p2, q2, poly2->v[i2]); c[1].id = cf_key(edge1, i2, CF_FACE, CF_VERTEX);
}

static int clip_segment_to_line(ClipVertex vOut[2], const ClipVertex vIn[2], V2 normal, float offset, int vertexIndexA) {
    int numOut = 0;
    float distance0 = vdot(normal, vIn[0].v) - offset;
    float distance1 = vdot(normal, vIn[1].v) - offset;
    if (distance0 <= 0.0f) vOut[numOut++] = vIn[0];
    if (distance1 <= 0.0f) vOut[numOut++] = vIn[1];
    if (distance0 * distance1 < 0.0f) {
        float interp = distance0 / (distance0 - distance1);
        vOut[numOut].v = vadd(vIn[0].v, vscale(interp, vsub(vIn[1].v, vIn[0].v)));
        vOut[numOut].id = cf_key(vertexIndexA, (int)((vIn[0].id >> 8) & 0xff), CF_VERTEX, CF_FACE);
        ++numOut;
    }
    return numOut;
}


static void collide_polygons(Manifold* m, const Poly* polyA, V2 pA, Rot qA, const Poly* polyB, V2 pB, Rot qB) {
    m->pointCount = 0;
    const float totalRadius = B2_POLYGON_RADIUS + B2_POLYGON_RADIUS;
    int edgeA = 0; float separationA = find_max_separation(&edgeA, polyA, pA, qA, polyB, pB, qB);
    if (separationA > totalRadius) return;
    int edgeB = 0; float separationB = find_max_separation(&edgeB, polyB, pB, qB, polyA, pA, qA);
    if (separationB > totalRadius) return;
    const Poly *poly1, *poly2; V2 p1, p2; Rot q1, q2; int edge1, flip;
    const float k_tol = 0.1f * B2_LINEAR_SLOP;
    if (separationB > separationA + k_tol) { poly1 = polyB; poly2 = polyA; p1 = pB; q1 = qB; p2 = pA; q2 = qA; edge1 = edgeB; m->type = 1; flip = 1; }
    else { poly1 = polyA; poly2 = polyB; p1 = pA; q1 = qA; p2 = pB; q2 = qB; edge1 = edgeA; m->type = 0; flip = 0; }
    ClipVertex incidentEdge[2];
    find_incident_edge(incidentEdge, poly1, q1, edge1, poly2, p2, q2);
    int iv1 = edge1, iv2 = edge1 + 1 < poly1->n ? edge1 + 1 : 0;
    V2 v11 = poly1->v[iv1], v12 = poly1->v[iv2];
    V2 localTangent = vsub(v12, v11);
    { float len = vlen(localTangent); if (len >= B2_EPSILON) { float inv = 1.0f / len; localTangent.x *= inv; localTangent.y *= inv; } }
    V2 localNormal = cross_vs(localTangent, 1.0f);
    V2 planePoint = vscale(0.5f, vadd(v11, v12));
    V2 tangent = rmul(q1, localTangent);
    V2 normal = cross_vs(tangent, 1.0f);
    v11 = xmul(p1, q1, v11); v12 = xmul(p1, q1, v12);
    float frontOffset = vdot(normal, v11);
    float sideOffset1 = -vdot(tangent, v11) + totalRadius;
    float sideOffset2 = vdot(tangent, v12) + totalRadius;
    ClipVertex clipPoints1[2], clipPoints2[2];
    int np = clip_segment_to_line(clipPoints1, incidentEdge, vneg(tangent), sideOffset1, iv1);
    if (np < 2) return;
    np = clip_segment_to_line(clipPoints2, clipPoints1, tangent, sideOffset2, iv2);
    if (np < 2) return;
    m->localNormal = localNormal; m->localPoint = planePoint;
    int pointCount = 0;
    for (int i = 0; i < 2; ++i) {
        float separation = vdot(normal, clipPoints2[i].v) - frontOffset;
        if (separation <= totalRadius) {
            MPoint* cp = &m->pt[pointCount];
            cp->localPoint = xmulT(p2, q2, clipPoints2[i].v);
            uint32_t id = clipPoints2[i].id;
            if (flip) id = cf_key((int)((id >> 8) & 0xff), (int)(id & 0xff), (int)((id >> 24) & 0xff), (int)((id >> 16) & 0xff));
            cp->id = id;
            ++pointCount;
        }
    }
    m->pointCount = pointCount;
}

/* b2ContactManager::Collide for the solid car-car pairs (start-of-step poses).  Pair order =
 * (car a < car b, fixture of a ascending, fixture of b ascending): deterministic stand-in for
 * the history-dependent contact-list order of Box2D (D1). */
static void collide_cars(OrcWorld* W) {
    Manifold* old = W->manifolds; int nold = W->n_manifolds;
    static __thread Manifold fresh[MAX_MANIFOLDS]; int nnew = 0;
    const float r = B2_POLYGON_RADIUS;
    /* awake flags as they are when Collide() starts (a wake caused by one pair does not change how
     * a later pair of the same pass is treated; deterministic stand-in, see D1) */
    int awake0[MAX_AGENTS][8];
    for (int a = 0; a < W->A; ++a) for (int f = 0; f < 8; ++f) awake0[a][f] = fixture_body(W, a, f)->awake;
    for (int a = 0; a < W->A; ++a) for (int b = a + 1; b < W->A; ++b)
        for (int fa = 0; fa < 8; ++fa) for (int fb = 0; fb < 8; ++fb) {
            if (fa < 4 && fb < 4) continue;                    /* wheel-wheel filtered out */
            Body* bA = fixture_body(W, a, fa); Body* bB = fixture_body(W, b, fb);
            const Poly* PA = fixture_poly(W, fa); const Poly* PB = fixture_poly(W, fb);
            const Manifold* prev = NULL;
            for (int i = 0; i < nold; ++i) if (old[i].carA == a && old[i].carB == b && old[i].fixA == fa && old[i].fixB == fb) { prev = &old[i]; break; }
            if (!awake0[a][fa] && !awake0[b][fb]) {            /* neither body active: contact not updated */
                if (prev && prev->pointCount > 0 && nnew < MAX_MANIFOLDS) fresh[nnew++] = *prev;
                continue;
            }
            /* b2PolygonShape::ComputeAABB (+ radius) and b2TestOverlap(aabb, aabb): a non-overlap means
             * a separation > 2 r, for which b2CollidePolygons returns no points */
            V2 va[MAXV], vb[MAXV]; float aa[4], ab[4];
            world_poly(bA, PA, va, aa); world_poly(bB, PB, vb, ab);
            if ((ab[0] - r) - (aa[2] + r) > 0.0f || (ab[1] - r) - (aa[3] + r) > 0.0f ||
                (aa[0] - r) - (ab[2] + r) > 0.0f || (aa[1] - r) - (ab[3] + r) > 0.0f) continue;
            Manifold m; memset(&m, 0, sizeof(m));
            m.carA = a; m.fixA = fa; m.carB = b; m.fixB = fb;
            collide_polygons(&m, PA, bA->p, bA->q, PB, bB->p, bB->q);
            int wasTouching = prev && prev->pointCount > 0;
            for (int i = 0; i < m.pointCount; ++i) {           /* b2Contact::Update: carry impulses by feature id */
                m.pt[i].normalImpulse = 0.0f; m.pt[i].tangentImpulse = 0.0f;
                if (prev) for (int j = 0; j < prev->pointCount; ++j)
                    if (prev->pt[j].id == m.pt[i].id) { m.pt[i].normalImpulse = prev->pt[j].normalImpulse; m.pt[i].tangentImpulse = prev->pt[j].tangentImpulse; break; }
            }
            if ((m.pointCount > 0) != wasTouching) {
                if (!bA->awake) { bA->awake = 1; bA->sleepTime = 0.0f; }
                if (!bB->awake) { bB->awake = 1; bB->sleepTime = 0.0f; }
            }
            if (m.pointCount > 0) { if (nnew < MAX_MANIFOLDS) fresh[nnew++] = m; else W->manifold_overflow = 1; }
        }
    memcpy(W->manifolds, fresh, sizeof(Manifold) * nnew);
    W->n_manifolds = nnew;
}

typedef struct {
    Body *A, *B; Manifold* m;
    V2 normal; float friction;
    int pointCount;               /* may drop to 1 when the block solver finds the points redundant */
    V2 rA[2], rB[2]; float normalMass[2], tangentMass[2], velocityBias[2], normalImpulse[2], tangentImpulse[2];
    float K[2][2], NM[2][2];      /* b2Mat22 as [col][row] */
} VC;

static void world_manifold(const Manifold* m, V2 pA, Rot qA, V2 pB, Rot qB, V2* normal, V2 points[2]) {
    const float radiusA = B2_POLYGON_RADIUS, radiusB = B2_POLYGON_RADIUS;
    if (m->type == 0) {
        *normal = rmul(qA, m->localNormal);
        V2 planePoint = xmul(pA, qA, m->localPoint);
        for (int i = 0; i < m->pointCount; ++i) {
            V2 clipPoint = xmul(pB, qB, m->pt[i].localPoint);
            V2 cA = vadd(clipPoint, vscale(radiusA - vdot(vsub(clipPoint, planePoint), *normal), *normal));
            V2 cB = vsub(clipPoint, vscale(radiusB, *normal));
            points[i] = vscale(0.5f, vadd(cA, cB));
        }
    } else {
        *normal = rmul(qB, m->localNormal);
        V2 planePoint = xmul(pB, qB, m->localPoint);
        for (int i = 0; i < m->pointCount; ++i) {
            V2 clipPoint = xmul(pA, qA, m->pt[i].localPoint);
            V2 cB = vadd(clipPoint, vscale(radiusB - vdot(vsub(clipPoint, planePoint), *normal), *normal));
            V2 cA = vsub(clipPoint, vscale(radiusA, *normal));
            points[i] = vscale(0.5f, vadd(cA, cB));
        }
        *normal = vneg(*normal);
    }
}

static void contact_init(VC* vc) {
    Body *A = vc->A, *B = vc->B; const Manifold* m = vc->m;
    float mA = A->invMass, mB = B->invMass, iA = A->invI, iB = B->invI;
    Rot qA = rot_set(A->a), qB = rot_set(B->a);
    V2 pA = vsub(A->c, rmul(qA, A->localCenter)), pB = vsub(B->c, rmul(qB, B->localCenter));
    V2 points[2];
    world_manifold(m, pA, qA, pB, qB, &vc->normal, points);
    vc->pointCount = m->pointCount;
    vc->friction = sqrtf(0.2f * 0.2f);            /* b2MixFriction of the default fixture frictions */
    for (int j = 0; j < vc->pointCount; ++j) {
        vc->normalImpulse[j] = m->pt[j].normalImpulse;   /* dtRatio == 1 */
        vc->tangentImpulse[j] = m->pt[j].tangentImpulse;
        vc->rA[j] = vsub(points[j], A->c); vc->rB[j] = vsub(points[j], B->c);
        float rnA = vcross(vc->rA[j], vc->normal), rnB = vcross(vc->rB[j], vc->normal);
        float kNormal = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
        vc->normalMass[j] = kNormal > 0.0f ? 1.0f / kNormal : 0.0f;
        V2 tangent = cross_vs(vc->normal, 1.0f);
        float rtA = vcross(vc->rA[j], tangent), rtB = vcross(vc->rB[j], tangent);
        float kTangent = mA + mB + iA * rtA * rtA + iB * rtB * rtB;
        vc->tangentMass[j] = kTangent > 0.0f ? 1.0f / kTangent : 0.0f;
        vc->velocityBias[j] = 0.0f;               /* restitution 0 */
    }
    if (vc->pointCount == 2) {
        float rn1A = vcross(vc->rA[0], vc->normal), rn1B = vcross(vc->rB[0], vc->normal);
        float rn2A = vcross(vc->rA[1], vc->normal), rn2B = vcross(vc->rB[1], vc->normal);
        float k11 = mA + mB + iA * rn1A * rn1A + iB * rn1B * rn1B;
        float k22 = mA + mB + iA * rn2A * rn2A + iB * rn2B * rn2B;
        float k12 = mA + mB + iA * rn1A * rn2A + iB * rn1B * rn2B;
        const float k_maxConditionNumber = 1000.0f;
        if (k11 * k11 < k_maxConditionNumber * (k11 * k22 - k12 * k12)) {
            vc->K[0][0] = k11; vc->K[0][1] = k12; vc->K[1][0] = k12; vc->K[1][1] = k22;
            float a = k11, b = k12, c = k12, d = k22;
            float det = a * d - b * c;
            if (det != 0.0f) det = 1.0f / det;
            vc->NM[0][0] = det * d; vc->NM[1][0] = -det * b; vc->NM[0][1] = -det * c; vc->NM[1][1] = det * a;
        } else {
            vc->pointCount = 1;
        }
    }
}

static void contact_warm_start(VC* vc) {
    Body *A = vc->A, *B = vc->B;
    float mA = A->invMass, mB = B->invMass, iA = A->invI, iB = B->invI;
    V2 tangent = cross_vs(vc->normal, 1.0f);
    for (int j = 0; j < vc->pointCount; ++j) {
        V2 P = vadd(vscale(vc->normalImpulse[j], vc->normal), vscale(vc->tangentImpulse[j], tangent));
        A->w -= iA * vcross(vc->rA[j], P); A->v = vsub(A->v, vscale(mA, P));
        B->w += iB * vcross(vc->rB[j], P); B->v = vadd(B->v, vscale(mB, P));
    }
}

static void contact_apply2(VC* vc, V2* vA, float* wA, V2* vB, float* wB, float dx, float dy) {
    Body *A = vc->A, *B = vc->B;
    float mA = A->invMass, mB = B->invMass, iA = A->invI, iB = B->invI;
    V2 P1 = vscale(dx, vc->normal), P2 = vscale(dy, vc->normal);
    *vA = vsub(*vA, vscale(mA, vadd(P1, P2)));
    *wA -= iA * (vcross(vc->rA[0], P1) + vcross(vc->rA[1], P2));
    *vB = vadd(*vB, vscale(mB, vadd(P1, P2)));
    *wB += iB * (vcross(vc->rB[0], P1) + vcross(vc->rB[1], P2));
}

static void contact_solve_vel(VC* vc) {
    Body *A = vc->A, *B = vc->B;
    float mA = A->invMass, mB = B->invMass, iA = A->invI, iB = B->invI;
    V2 vA = A->v, vB = B->v; float wA = A->w, wB = B->w;
    V2 normal = vc->normal, tangent = cross_vs(normal, 1.0f);
    for (int j = 0; j < vc->pointCount; ++j) {      /* friction first */
        V2 dv = vsub(vsub(vadd(vB, cross_sv(wB, vc->rB[j])), vA), cross_sv(wA, vc->rA[j]));
        float vt = vdot(dv, tangent) - 0.0f;
        float lambda = vc->tangentMass[j] * (-vt);
        float maxFriction = vc->friction * vc->normalImpulse[j];
        float newImpulse = clampf(vc->tangentImpulse[j] + lambda, -maxFriction, maxFriction);
        lambda = newImpulse - vc->tangentImpulse[j];
        vc->tangentImpulse[j] = newImpulse;
        V2 P = vscale(lambda, tangent);
        vA = vsub(vA, vscale(mA, P)); wA -= iA * vcross(vc->rA[j], P);
        vB = vadd(vB, vscale(mB, P)); wB += iB * vcross(vc->rB[j], P);
    }
    if (vc->pointCount == 1) {
        V2 dv = vsub(vsub(vadd(vB, cross_sv(wB, vc->rB[0])), vA), cross_sv(wA, vc->rA[0]));
        float vn = vdot(dv, normal);
        float lambda = -vc->normalMass[0] * (vn - vc->velocityBias[0]);
        float newImpulse = fmaxf(vc->normalImpulse[0] + lambda, 0.0f);
        lambda = newImpulse - vc->normalImpulse[0];
        vc->normalImpulse[0] = newImpulse;
        V2 P = vscale(lambda, normal);
        vA = vsub(vA, vscale(mA, P)); wA -= iA * vcross(vc->rA[0], P);
        vB = vadd(vB, vscale(mB, P)); wB += iB * vcross(vc->rB[0], P);
    } else {                                        /* block solver */
        float ax = vc->normalImpulse[0], ay = vc->normalImpulse[1];
        V2 dv1 = vsub(vsub(vadd(vB, cross_sv(wB, vc->rB[0])), vA), cross_sv(wA, vc->rA[0]));
        V2 dv2 = vsub(vsub(vadd(vB, cross_sv(wB, vc->rB[1])), vA), cross_sv(wA, vc->rA[1]));
        float vn1 = vdot(dv1, normal), vn2 = vdot(dv2, normal);
        float bx = vn1 - vc->velocityBias[0], by = vn2 - vc->velocityBias[1];
        bx -= vc->K[0][0] * ax + vc->K[1][0] * ay;
        by -= vc->K[0][1] * ax + vc->K[1][1] * ay;
        for (;;) {
            float xx = -(vc->NM[0][0] * bx + vc->NM[1][0] * by), xy = -(vc->NM[0][1] * bx + vc->NM[1][1] * by);
            if (xx >= 0.0f && xy >= 0.0f) {
                contact_apply2(vc, &vA, &wA, &vB, &wB, xx - ax, xy - ay);
                vc->normalImpulse[0] = xx; vc->normalImpulse[1] = xy; break;
            }
            xx = -vc->normalMass[0] * bx; xy = 0.0f;
            vn1 = 0.0f; vn2 = vc->K[0][1] * xx + by;
            if (xx >= 0.0f && vn2 >= 0.0f) {
                contact_apply2(vc, &vA, &wA, &vB, &wB, xx - ax, xy - ay);
                vc->normalImpulse[0] = xx; vc->normalImpulse[1] = xy; break;
            }
            xx = 0.0f; xy = -vc->normalMass[1] * by;
            vn1 = vc->K[1][0] * xy + bx; vn2 = 0.0f;
            if (xy >= 0.0f && vn1 >= 0.0f) {
                contact_apply2(vc, &vA, &wA, &vB, &wB, xx - ax, xy - ay);
                vc->normalImpulse[0] = xx; vc->normalImpulse[1] = xy; break;
            }
            xx = 0.0f; xy = 0.0f; vn1 = bx; vn2 = by;
            if (vn1 >= 0.0f && vn2 >= 0.0f) {
                contact_apply2(vc, &vA, &wA, &vB, &wB, xx - ax, xy - ay);
                vc->normalImpulse[0] = xx; vc->normalImpulse[1] = xy; break;
            }
            break;   /* no solution, give up */
        }
    }
    A->v = vA; A->w = wA; B->v = vB; B->w = wB;
}

/* b2ContactSolver::SolvePositionConstraints for one contact; updates *minSeparation */
static void contact_solve_pos(VC* vc, float* minSeparation) {
    Body *A = vc->A, *B = vc->B; const Manifold* m = vc->m;
    float mA = A->invMass, mB = B->invMass, iA = A->invI, iB = B->invI;
    V2 cA = A->c, cB = B->c; float aA = A->a, aB = B->a;
    for (int j = 0; j < m->pointCount; ++j) {
        Rot qA = rot_set(aA), qB = rot_set(aB);
        V2 pA = vsub(cA, rmul(qA, A->localCenter)), pB = vsub(cB, rmul(qB, B->localCenter));
        V2 normal, point; float separation;
        if (m->type == 0) {
            normal = rmul(qA, m->localNormal);
            V2 planePoint = xmul(pA, qA, m->localPoint);
            V2 clipPoint = xmul(pB, qB, m->pt[j].localPoint);
            separation = vdot(vsub(clipPoint, planePoint), normal) - B2_POLYGON_RADIUS - B2_POLYGON_RADIUS;
            point = clipPoint;
        } else {
            normal = rmul(qB, m->localNormal);
            V2 planePoint = xmul(pB, qB, m->localPoint);
            V2 clipPoint = xmul(pA, qA, m->pt[j].localPoint);
            separation = vdot(vsub(clipPoint, planePoint), normal) - B2_POLYGON_RADIUS - B2_POLYGON_RADIUS;
            point = clipPoint;
            normal = vneg(normal);
        }
        V2 rA = vsub(point, cA), rB = vsub(point, cB);
        *minSeparation = fminf(*minSeparation, separation);
        float C = clampf(B2_BAUMGARTE * (separation + B2_LINEAR_SLOP), -B2_MAX_LINEAR_CORRECTION, 0.0f);
        float rnA = vcross(rA, normal), rnB = vcross(rB, normal);
        float K = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
        float impulse = K > 0.0f ? -C / K : 0.0f;
        V2 P = vscale(impulse, normal);
        cA = vsub(cA, vscale(mA, P)); aA -= iA * vcross(rA, P);
        cB = vadd(cB, vscale(mB, P)); aB += iB * vcross(rB, P);
    }
    A->c = cA; A->a = aA; B->c = cB; B->a = aB;
}

/* b2Island::Solve for one island = the cars in `cars` plus the touching manifolds among them. */
static void solve_island(OrcWorld* W, const int* cars, int ncars, Manifold** mans, int nmans, float h, int velIters,
                         int posIters, float dtRatio) {
    static __thread VC vcs[MAX_MANIFOLDS];
    for (int ci = 0; ci < ncars; ++ci) {
        Car* car = &W->car[cars[ci]];
        for (int i = 0; i < 5; ++i) {                 /* DFS wakes every body of the island */
            Body* b = &car->b[i];
            if (!b->awake) { b->awake = 1; b->sleepTime = 0.0f; }
        }
        for (int i = 0; i < 5; ++i) {                 /* gravity (0,0), damping 0 */
            Body* b = &car->b[i];
            b->v.x += h * (b->invMass * b->force.x);
            b->v.y += h * (b->invMass * b->force.y);
            b->w += h * b->invI * b->torque;
        }
    }
    for (int i = 0; i < nmans; ++i) {
        vcs[i].m = mans[i];
        vcs[i].A = fixture_body(W, mans[i]->carA, mans[i]->fixA);
        vcs[i].B = fixture_body(W, mans[i]->carB, mans[i]->fixB);
        contact_init(&vcs[i]);
    }
    for (int i = 0; i < nmans; ++i) contact_warm_start(&vcs[i]);
    for (int ci = 0; ci < ncars; ++ci) {
        Car* car = &W->car[cars[ci]];
        for (int k = 0; k < 4; ++k) { int j = W->var_joint_ascending ? k : JOINT_ORDER[k]; joint_init(&car->j[j], &car->b[0], &car->b[1 + j], dtRatio); }
    }
    int fixed_at = -1;
    for (int it = 0; it < velIters; ++it) {
        Car before; if (fixed_at < 0 && ncars == 1) before = W->car[cars[0]];
        for (int ci = 0; ci < ncars; ++ci) {
            Car* car = &W->car[cars[ci]];
            for (int k = 0; k < 4; ++k) { int j = W->var_joint_ascending ? k : JOINT_ORDER[k]; joint_solve_vel(&car->j[j], &car->b[0], &car->b[1 + j], h); }
        }
        for (int i = 0; i < nmans; ++i) contact_solve_vel(&vcs[i]);
        if (fixed_at < 0 && ncars == 1 && state_equal(&before, &W->car[cars[0]])) fixed_at = it; /* diagnostics only */
    }
    { int fa = fixed_at < 0 ? velIters : fixed_at; if (fa > W->vel_iters_used) W->vel_iters_used = fa; }
    for (int i = 0; i < nmans; ++i)                   /* StoreImpulses */
        for (int j = 0; j < vcs[i].pointCount; ++j) {
            vcs[i].m->pt[j].normalImpulse = vcs[i].normalImpulse[j];
            vcs[i].m->pt[j].tangentImpulse = vcs[i].tangentImpulse[j];
        }
    for (int ci = 0; ci < ncars; ++ci) {
        Car* car = &W->car[cars[ci]];
        for (int i = 0; i < 5; ++i) {
            Body* b = &car->b[i];
            V2 tr = vscale(h, b->v);
            if (vdot(tr, tr) > B2_MAX_TRANSLATION_SQ) { float ratio = B2_MAX_TRANSLATION / vlen(tr); b->v = vscale(ratio, b->v); }
            float rotn = h * b->w;
            if (rotn * rotn > B2_MAX_ROTATION_SQ) { float ratio = B2_MAX_ROTATION / fabsf(rotn); b->w *= ratio; }
            b->c.x += h * b->v.x; b->c.y += h * b->v.y;
            b->a += h * b->w;
        }
    }
    int positionSolved = 0;
    for (int it = 0; it < posIters; ++it) {
        float minSeparation = 0.0f;
        for (int i = 0; i < nmans; ++i) contact_solve_pos(&vcs[i], &minSeparation);
        int contactsOkay = minSeparation >= -3.0f * B2_LINEAR_SLOP;
        int jointsOkay = 1;
        for (int ci = 0; ci < ncars; ++ci) {
            Car* car = &W->car[cars[ci]];
            for (int k = 0; k < 4; ++k) {
                int j = W->var_joint_ascending ? k : JOINT_ORDER[k];
                int ok = joint_solve_pos(&car->j[j], &car->b[0], &car->b[1 + j]);
                jointsOkay = jointsOkay && ok;
            }
        }
        if (contactsOkay && jointsOkay) { positionSolved = 1; break; }
    }
    float minSleepTime = FLT_MAX;
    const float linTolSqr = B2_LINEAR_SLEEP_TOL * B2_LINEAR_SLEEP_TOL;
    const float angTolSqr = B2_ANGULAR_SLEEP_TOL * B2_ANGULAR_SLEEP_TOL;
    for (int ci = 0; ci < ncars; ++ci) {
        Car* car = &W->car[cars[ci]];
        for (int i = 0; i < 5; ++i) {
            Body* b = &car->b[i];
            b->q = rot_set(b->a);                     /* SynchronizeTransform */
            V2 rc = rmul(b->q, b->localCenter);
            b->p = vsub(b->c, rc);
            if (b->w * b->w > angTolSqr || vdot(b->v, b->v) > linTolSqr) { b->sleepTime = 0.0f; minSleepTime = 0.0f; }
            else { b->sleepTime += h; minSleepTime = fminf(minSleepTime, b->sleepTime); }
        }
    }
    if (minSleepTime >= B2_TIME_TO_SLEEP && positionSolved) {
        for (int ci = 0; ci < ncars; ++ci) {
            Car* car = &W->car[cars[ci]];
            for (int i = 0; i < 5; ++i) {
                Body* b = &car->b[i];
                b->awake = 0; b->sleepTime = 0.0f; b->v = v2(0.0f, 0.0f); b->w = 0.0f; b->force = v2(0.0f, 0.0f); b->torque = 0.0f;
            }
        }
    }
}

static void world_solve(OrcWorld* W, float dt, int velIters, int posIters) {
    float inv_dt = dt > 0.0f ? 1.0f / dt : 0.0f;
    float dtRatio = W->inv_dt0 * dt;
    W->vel_iters_used = 0;
    if (W->collisions) collide_cars(W); else W->n_manifolds = 0;
    /* islands: cars linked by touching manifolds (union-find, root = lowest car index) */
    int root[MAX_AGENTS];
    for (int c = 0; c < W->A; ++c) root[c] = c;
    for (int i = 0; i < W->n_manifolds; ++i) {
        int a = W->manifolds[i].carA, b = W->manifolds[i].carB;
        while (root[a] != a) a = root[a];
        while (root[b] != b) b = root[b];
        if (a != b) { if (a < b) root[b] = a; else root[a] = b; }
    }
    for (int c = 0; c < W->A; ++c) { int r = c; while (root[r] != r) r = root[r]; root[c] = r; }
    /* islands are seeded from awake bodies in reverse creation order: last car first */
    int done[MAX_AGENTS] = {0};
    for (int c = W->A - 1; c >= 0; --c) {
        if (done[c]) continue;
        int cars[MAX_AGENTS], ncars = 0, any_awake = 0;
        for (int k = W->A - 1; k >= 0; --k) if (root[k] == root[c]) { cars[ncars++] = k; done[k] = 1; }
        Manifold* mans[MAX_MANIFOLDS]; int nmans = 0;
        for (int i = 0; i < W->n_manifolds; ++i) if (root[W->manifolds[i].carA] == root[c]) mans[nmans++] = &W->manifolds[i];
        for (int k = 0; k < ncars; ++k) for (int i = 0; i < 5; ++i) any_awake |= W->car[cars[k]].b[i].awake;
        if (!any_awake) continue;
        solve_island(W, cars, ncars, mans, nmans, dt, velIters, posIters, dtRatio);
    }
    W->inv_dt0 = inv_dt;
    /* ClearForces */
    for (int c = 0; c < W->A; ++c)
        for (int i = 0; i < 5; ++i) { W->car[c].b[i].force = v2(0.0f, 0.0f); W->car[c].b[i].torque = 0.0f; }
}

static void world_step(OrcWorld* W, float dt, int velIters, int posIters) {
    collide(W);
    world_solve(W, dt, velIters, posIters);
}

/* ------------------------------------------------------------------------------------ */
/* Software rasteriser: GL fixed-function point sampling restated            */
/* ------------------------------------------------------------------------------------ */
#define STATE_W 96
#define STATE_H 96

/* img[(h - 1 - y) * w * 3 + x * 3] (row flip, mcr:602); w x h = the glViewport size of the render mode
 * (state_pixels 96 x 96, rgb_array 600 x 400, mcr:566-575) */
typedef struct { uint8_t* img; int w, h; } Canvas;

/* Fill a convex polygon given in viewport pixel coordinates (origin bottom-left).  A pixel is
 * covered iff its centre lies in [xl, xr) on a row whose centre lies in [ylo, yhi) of two
 * edges; every edge is evaluated from its lower to its higher endpoint, so an edge shared
 * by two polygons yields bit-identical crossings (each pixel centre belongs to exactly one). */
static void fill_poly(Canvas* cv, const float* px, const float* py, int n, RGB col) {
    const int VW = cv->w, VH = cv->h;
    float ymin = py[0], ymax = py[0];
    for (int i = 1; i < n; ++i) { ymin = fminf(ymin, py[i]); ymax = fmaxf(ymax, py[i]); }
    if (!(ymax > 0.0f) || !(ymin < (float)VH)) return;
    int y0 = (int)ceilf(fmaxf(ymin, 0.0f) - 0.5f); if (y0 < 0) y0 = 0;
    int y1 = (int)ceilf(fminf(ymax, (float)VH) - 0.5f); if (y1 > VH) y1 = VH;
    for (int y = y0; y < y1; ++y) {
        float yc = (float)y + 0.5f;
        float xl = FLT_MAX, xr = -FLT_MAX;
        for (int i = 0; i < n; ++i) {
            int k = i + 1 < n ? i + 1 : 0;
            float ax = px[i], ay = py[i], bx = px[k], by = py[k];
            if (ay == by) continue;
            if (ay > by) { float tx = ax, ty = ay; ax = bx; ay = by; bx = tx; by = ty; }
            if (!(yc >= ay && yc < by)) continue;
            float x = ax + (yc - ay) * ((bx - ax) / (by - ay));
            xl = fminf(xl, x); xr = fmaxf(xr, x);
        }
        if (!(xl < xr)) continue;
        xl = fminf(fmaxf(xl, -1.0f), (float)VW + 1.0f);
        xr = fminf(fmaxf(xr, -1.0f), (float)VW + 1.0f);
        int x0 = (int)ceilf(xl - 0.5f); if (x0 < 0) x0 = 0;
        int x1 = (int)ceilf(xr - 0.5f); if (x1 > VW) x1 = VW;
        uint8_t* row = cv->img + (size_t)(VH - 1 - y) * VW * 3;
        for (int x = x0; x < x1; ++x) { row[3 * x] = col.r; row[3 * x + 1] = col.g; row[3 * x + 2] = col.b; }
    }
}

typedef struct { float m00, m01, m02, m10, m11, m12; } Affine; /* world -> viewport pixels */

static inline void xf_pt(const Affine* M, float x, float y, float* ox, float* oy) {
    *ox = (M->m00 * x + M->m01 * y) + M->m02;
    *oy = (M->m10 * x + M->m11 * y) + M->m12;
}

static void fill_world_poly(Canvas* cv, const Affine* M, const V2* v, int n, RGB col) {
    float px[MAXV], py[MAXV];
    for (int i = 0; i < n; ++i) xf_pt(M, v[i].x, v[i].y, &px[i], &py[i]);
    fill_poly(cv, px, py, n, col);
}

/* window (1000 x 800) coordinates -> viewport pixels: glViewport(0,0,w,h) under the
 * unchanged glOrtho(0,1000,0,800) projection (mcr:576-586) */
static void fill_window_poly(Canvas* cv, const double* wx, const double* wy, int n, RGB col) {
    float px[MAXV], py[MAXV];
    for (int i = 0; i < n; ++i) {
        float fx = (float)wx[i], fy = (float)wy[i]; /* glVertex3f */
        px[i] = fx * (float)((double)cv->w / 1000.0);
        py[i] = fy * (float)((double)cv->h / 800.0);
    }
    fill_poly(cv, px, py, n, col);
}

static const uint8_t FONT3x5[11][5] = { /* rows top->bottom, 3 bits per row (msb = left) */
    {7, 5, 5, 5, 7}, {2, 6, 2, 2, 7}, {7, 1, 7, 4, 7}, {7, 1, 7, 1, 7}, {5, 5, 7, 1, 1},
    {7, 4, 7, 1, 7}, {7, 4, 7, 5, 7}, {7, 1, 1, 1, 1}, {7, 5, 7, 5, 7}, {7, 5, 7, 1, 7},
    {0, 0, 7, 0, 0} /* '-' */
};

/* pyglet Label at x=20, y=50 (window) -> 3x5 glyphs at cols 2..13, rows 87..91 from the top of the
 * 96 x 96 state frame (D3); larger viewports show the same glyph cells scaled (nearest neighbour:
 * viewport pixel (X, Y) takes state cell (floor((X + .5) * 96 / w), floor((Y + .5) * 96 / h))) */
static void draw_label_vp(uint8_t* img, int vw, int vh, const char* text) {
    uint8_t cell[5][12]; memset(cell, 0, sizeof(cell));
    for (int ch = 0; ch < 4 && text[ch]; ++ch) {
        int g = text[ch] == '-' ? 10 : (text[ch] >= '0' && text[ch] <= '9' ? text[ch] - '0' : -1);
        if (g < 0) continue;
        for (int ry = 0; ry < 5; ++ry)
            for (int rx = 0; rx < 3; ++rx)
                if (FONT3x5[g][ry] & (4 >> rx)) cell[ry][3 * ch + rx] = 1;
    }
    if (vw == STATE_W && vh == STATE_H) {
        for (int ry = 0; ry < 5; ++ry) for (int cx = 0; cx < 12; ++cx) if (cell[ry][cx]) {
            uint8_t* p = img + ((size_t)(87 + ry) * STATE_W + (2 + cx)) * 3;
            p[0] = 255; p[1] = 255; p[2] = 255;
        }
        return;
    }
    for (int Y = 0; Y < vh; ++Y) {          /* Y, X: image row from the top, column */
        int sy = (int)floor(((double)Y + 0.5) * 96.0 / (double)vh) - 87;
        if (sy < 0 || sy >= 5) continue;
        for (int X = 0; X < vw; ++X) {
            int sx = (int)floor(((double)X + 0.5) * 96.0 / (double)vw) - 2;
            if (sx < 0 || sx >= 12 || !cell[sy][sx]) continue;
            uint8_t* p = img + ((size_t)Y * vw + X) * 3;
            p[0] = 255; p[1] = 255; p[2] = 255;
        }
    }
}
static void draw_label(uint8_t* img, const char* text) { draw_label_vp(img, STATE_W, STATE_H, text); }

static const float CAR_COLORS[8][3] = { {0.8f, 0.0f, 0.0f}, {0.0f, 0.0f, 0.8f}, {0.0f, 0.8f, 0.0f}, {0.0f, 0.8f, 0.8f},
                                        {0.8f, 0.8f, 0.8f}, {0.0f, 0.0f, 0.0f}, {0.8f, 0.0f, 0.8f}, {0.8f, 0.8f, 0.0f} };

static void render_view_vp(OrcWorld* W, int agent, uint8_t* img, int vw, int vh) {
    Canvas cv; cv.img = img; cv.w = vw; cv.h = vh;
    memset(img, 0, (size_t)vw * vh * 3); /* glClear, default clear colour */
    const double SCALE = 6.0, ZOOM = 2.7, WINDOW_W = 1000, WINDOW_H = 800;
    const double PLAYFIELD = 2000 / SCALE;
    /* camera, mcr:540-556 */
    double t = W->t;
    double zoom = 0.1 * SCALE * fmax(1 - t, 0) + ZOOM * SCALE * fmin(t, 1);
    const Body* hull = &W->car[agent].b[0];
    double scroll_x = hull->p.x, scroll_y = hull->p.y;
    double angle = -(double)hull->a;
    double vx = hull->v.x, vy = hull->v.y;
    if (sqrt(vx * vx + vy * vy) > 0.5) angle = atan2(vx, vy);
    double tx = WINDOW_W / 2 - (scroll_x * zoom * cos(angle) - scroll_y * zoom * sin(angle));
    double ty = WINDOW_H * W->h_ratio - (scroll_x * zoom * sin(angle) + scroll_y * zoom * cos(angle));
    /* Transform.enable: glTranslatef(tx,ty,0); glRotatef(RAD2DEG*angle,0,0,1); glScalef(zoom,zoom,1) */
    float ftx = (float)tx, fty = (float)ty, fdeg = (float)(57.29577951308232 * angle), fzoom = (float)zoom;
    double rad = (double)fdeg * (3.14159265358979323846 / 180.0);
    double cs = cos(rad), sn = sin(rad);
    const double SX = (double)vw / 1000.0, SY = (double)vh / 800.0;
    Affine M;
    M.m00 = (float)(cs * (double)fzoom * SX); M.m01 = (float)(-sn * (double)fzoom * SX); M.m02 = (float)((double)ftx * SX);
    M.m10 = (float)(sn * (double)fzoom * SY); M.m11 = (float)(cs * (double)fzoom * SY);  M.m12 = (float)((double)fty * SY);

    /* render_road, mcr:613-632 */
    {
        float pf = (float)PLAYFIELD;
        V2 q[4] = { v2(-pf, +pf), v2(+pf, +pf), v2(+pf, -pf), v2(-pf, -pf) };
        fill_world_poly(&cv, &M, q, 4, rgbf(0.4f, 0.8f, 0.4f));
        double k = PLAYFIELD / 20.0;
        RGB light = rgbf(0.4f, 0.9f, 0.4f);
        for (int x = -20; x < 20; x += 2)
            for (int y = -20; y < 20; y += 2) {
                V2 c[4] = { v2((float)(k * x + k), (float)(k * y + 0)), v2((float)(k * x + 0), (float)(k * y + 0)),
                            v2((float)(k * x + 0), (float)(k * y + k)), v2((float)(k * x + k), (float)(k * y + k)) };
                fill_world_poly(&cv, &M, c, 4, light);
            }
        for (int qi = 0; qi < W->Q; ++qi) {
            const float* qv = &W->quad[8 * qi];
            V2 c[4] = { v2(qv[0], qv[1]), v2(qv[2], qv[3]), v2(qv[4], qv[5]), v2(qv[6], qv[7]) };
            RGB col;
            int tl = W->quad_tile[qi];
            if (tl >= 0 && W->touched[tl]) col = rgbf(0.4f, 0.4f, 0.4f);
            else col = rgbf(W->quad_rgb[3 * qi], W->quad_rgb[3 * qi + 1], W->quad_rgb[3 * qi + 2]);
            fill_world_poly(&cv, &M, c, 4, col);
        }
    }
    /* cars: Car.draw(viewer, draw_particles=False), all cars in id order, mcr:559-564 */
    for (int c = 0; c < W->A; ++c) {
        const Car* car = &W->car[c];
        /* Car.draw(viewer, draw_particles = mode != "state_pixels"), mcr:564: skid traces first.  D6: a
         * glLineWidth(5) GL_LINE_STRIP segment is the parallelogram spanning +-2.5 viewport pixels along the
         * minor axis (y for |dx| >= |dy|, else x), filled with the polygon rule. */
        if (W->draw_particles && !(vw == STATE_W && vh == STATE_H)) {
            for (int i = 0; i < car->n_prt; ++i) {
                RGB pc = car->prt[i].grass ? rgbf(0.4f, 0.4f, 0.0f) : rgbf(0.0f, 0.0f, 0.0f);
                for (int k = 0; k + 1 < car->prt[i].len; ++k) {
                    float x0, y0, x1, y1;
                    xf_pt(&M, car->prt[i].pt[k].x, car->prt[i].pt[k].y, &x0, &y0);
                    xf_pt(&M, car->prt[i].pt[k + 1].x, car->prt[i].pt[k + 1].y, &x1, &y1);
                    const float hw = 2.5f;
                    float qx[4], qy[4];
                    if (fabsf(x1 - x0) >= fabsf(y1 - y0)) {
                        qx[0] = x0; qy[0] = y0 - hw; qx[1] = x1; qy[1] = y1 - hw; qx[2] = x1; qy[2] = y1 + hw; qx[3] = x0; qy[3] = y0 + hw;
                    } else {
                        qx[0] = x0 - hw; qy[0] = y0; qx[1] = x1 - hw; qy[1] = y1; qx[2] = x1 + hw; qy[2] = y1; qx[3] = x0 + hw; qy[3] = y0;
                    }
                    fill_poly(&cv, qx, qy, 4, pc);
                }
            }
        }
        RGB hullcol;
        if (W->use_ego_color) hullcol = (c == agent) ? rgbf(0.8f, 0.0f, 0.0f) : rgbf(0.0f, 0.0f, 0.8f);
        else hullcol = rgbf(CAR_COLORS[car->hull_color][0], CAR_COLORS[car->hull_color][1], CAR_COLORS[car->hull_color][2]);
        for (int w = 0; w < 4; ++w) {
            const Body* wb = &car->b[1 + w];
            V2 path[MAXV];
            for (int i = 0; i < W->wheel_poly.n; ++i) path[i] = xmul(wb->p, wb->q, W->wheel_poly.v[i]);
            fill_world_poly(&cv, &M, path, W->wheel_poly.n, rgbf(0.0f, 0.0f, 0.0f));
            double a1 = car->phase[w], a2 = car->phase[w] + 1.2;
            double s1 = sin(a1), s2 = sin(a2), c1 = cos(a1), c2 = cos(a2);
            if (s1 > 0 && s2 > 0) continue;
            if (s1 > 0) c1 = sign_d(c1);
            if (s2 > 0) c2 = sign_d(c2);
            V2 wp[4] = { v2((float)(-WHEEL_W * SIZE), (float)(+WHEEL_R * c1 * SIZE)), v2((float)(+WHEEL_W * SIZE), (float)(+WHEEL_R * c1 * SIZE)),
                         v2((float)(+WHEEL_W * SIZE), (float)(+WHEEL_R * c2 * SIZE)), v2((float)(-WHEEL_W * SIZE), (float)(+WHEEL_R * c2 * SIZE)) };
            for (int i = 0; i < 4; ++i) path[i] = xmul(wb->p, wb->q, wp[i]);
            fill_world_poly(&cv, &M, path, 4, rgbf(0.3f, 0.3f, 0.3f));
        }
        const Body* hb = &car->b[0];
        for (int f = 3; f >= 0; --f) { /* body.fixtures iterates newest first */
            V2 path[MAXV];
            for (int i = 0; i < W->hull_poly[f].n; ++i) path[i] = xmul(hb->p, hb->q, W->hull_poly[f].v[i]);
            fill_world_poly(&cv, &M, path, W->hull_poly[f].n, hullcol);
        }
    }
    /* render_indicators, mcr:634-674 */
    {
        double Wd = WINDOW_W, Hd = WINDOW_H, s = Wd / 40.0, h = Hd / 40.0;
        double bx[4] = { Wd, Wd, 0, 0 }, by[4] = { 0, 5 * h, 5 * h, 0 };
        fill_window_poly(&cv, bx, by, 4, rgbf(0, 0, 0));
        const Car* car = &W->car[agent];
        double lvx = car->b[0].v.x, lvy = car->b[0].v.y;
        double true_speed = sqrt(lvx * lvx + lvy * lvy);
        struct { double place, val; float r, g, b; } vi[5] = {
            { 5, 0.02 * true_speed, 1, 1, 1 },
            { 7, 0.01 * car->omega[0], 0.0f, 0, 1 }, { 8, 0.01 * car->omega[1], 0.0f, 0, 1 },
            { 9, 0.01 * car->omega[2], 0.2f, 0, 1 }, { 10, 0.01 * car->omega[3], 0.2f, 0, 1 } };
        for (int i = 0; i < 5; ++i) {
            double place = vi[i].place, val = vi[i].val;
            double qx[4] = { (place + 0) * s, (place + 1) * s, (place + 1) * s, (place + 0) * s };
            double qy[4] = { h + h * val, h + h * val, h, h };
            fill_window_poly(&cv, qx, qy, 4, rgbf(vi[i].r, vi[i].g, vi[i].b));
        }
        double jangle0 = (double)((car->b[1].a - car->b[0].a) - car->j[0].refAngle);
        struct { double place, val; float r, g, b; } hi[2] = {
            { 20, -10.0 * jangle0, 0, 1, 0 }, { 30, -0.8 * (double)car->b[0].w, 1, 0, 0 } };
        for (int i = 0; i < 2; ++i) {
            double place = hi[i].place, val = hi[i].val;
            double qx[4] = { (place + 0) * s, (place + val) * s, (place + val) * s, (place + 0) * s };
            double qy[4] = { 4 * h, 4 * h, 2 * h, 2 * h };
            fill_window_poly(&cv, qx, qy, 4, rgbf(hi[i].r, hi[i].g, hi[i].b));
        }
        /* score label "%04i" % reward (D3: baked 3x5 font, cols 2..13, rows 87..91 from top) */
        {
            int val = (int)W->reward[agent]; /* Python %i truncates toward zero */
            char buf[16]; int neg = W->reward[agent] < 0 && val != 0;
            int mag = val < 0 ? -val : val;
            /* "%04i": width 4 including sign, zero padded */
            int nd = 0; char digs[12];
            do { digs[nd++] = (char)('0' + mag % 10); mag /= 10; } while (mag > 0);
            int width = nd + (neg ? 1 : 0); int pad = width < 4 ? 4 - width : 0; int len = 0;
            if (neg) buf[len++] = '-';
            for (int i = 0; i < pad; ++i) buf[len++] = '0';
            for (int i = nd - 1; i >= 0; --i) buf[len++] = digs[i];
            buf[len] = 0;
            draw_label_vp(img, vw, vh, buf);
        }
        if (W->driving_backward[agent] && W->backwards_flag) {
            double fx[3] = { Wd - 100, Wd - 75, Wd - 50 }, fy[3] = { 30, 70, 30 };
            RGB blue; blue.r = 0; blue.g = 0; blue.b = 255;
            fill_window_poly(&cv, fx, fy, 3, blue);
        }
    }
}

static void render_view(OrcWorld* W, int agent, uint8_t* img) { render_view_vp(W, agent, img, STATE_W, STATE_H); }

/* ------------------------------------------------------------------------------------ */
/* MultiCarRacing.step  (mcr:410-509).  action == NULL  <=>  step(None)                  */
/* ------------------------------------------------------------------------------------ */
static double py_mod(double a, double b) { double r = fmod(a, b); if (r != 0 && ((r < 0) != (b < 0))) r += b; return r; }

ORC_API void orc_step(OrcWorld* W, const double* action, uint8_t* obs, double* step_reward, int* done_out, int render) {
    const double FPS = 50;
    const double PLAYFIELD = 2000 / 6.0;
    const double PI = 3.141592653589793;
    int A = W->A;
    if (action) for (int c = 0; c < A; ++c) car_controls(&W->car[c], action[3 * c], action[3 * c + 1], action[3 * c + 2]);
    for (int c = 0; c < A; ++c) car_step(&W->car[c], 1.0 / FPS);
    world_step(W, (float)(1.0 / FPS), 6 * 30, 2 * 30);
    W->t += 1.0 / FPS;
    if (render && obs) for (int c = 0; c < A; ++c) render_view(W, c, obs + (size_t)c * STATE_W * STATE_H * 3);
    int done = 0;
    for (int c = 0; c < A; ++c) step_reward[c] = 0.0;
    if (action) {
        for (int c = 0; c < A; ++c) W->reward[c] -= 0.1;
        for (int c = 0; c < A; ++c) step_reward[c] = W->reward[c] - W->prev_reward[c];
        for (int c = 0; c < A; ++c) {
            const Body* hull = &W->car[c].b[0];
            double vx = hull->v.x, vy = hull->v.y, car_angle;
            if (sqrt(vx * vx + vy * vy) > 0.5) car_angle = -atan2(vx, vy);
            else car_angle = (double)hull->a;
            car_angle = py_mod(car_angle + 2 * PI, 2 * PI);
            double px = hull->p.x, py = hull->p.y;
            int best = 0; double bestd = 0;
            for (int i = 0; i < W->T; ++i) {
                double dx = px - W->track[4 * i + 2], dy = py - W->track[4 * i + 3];
                double d = sqrt(dx * dx + dy * dy);
                if (i == 0 || d < bestd) { bestd = d; best = i; }
            }
            /* on_grass = not any(Point(car_pos).within(Polygon(road_poly[i]))), mcr:469-472: strictly inside one
             * of the (convex) road / border quads, float64 */
            {
                int inside_any = 0;
                for (int q = 0; q < W->Q && !inside_any; ++q) {
                    const double* v = &W->quad64[8 * q];
                    int pos = 0, neg = 0;
                    for (int k = 0; k < 4; ++k) {
                        int k2 = (k + 1) & 3;
                        double ex = v[2 * k2] - v[2 * k], ey = v[2 * k2 + 1] - v[2 * k + 1];
                        double cr = ex * (py - v[2 * k + 1]) - ey * (px - v[2 * k]);
                        if (cr > 0) ++pos; else if (cr < 0) ++neg;
                    }
                    if (pos == 4 || neg == 4) inside_any = 1;
                }
                W->driving_on_grass[c] = (uint8_t)!inside_any;
            }
            double desired = W->track[4 * best + 1];
            if (W->cw) desired += PI;
            desired = py_mod(desired + 2 * PI, 2 * PI);
            double diff = fabs(desired - car_angle);
            if (diff > PI) diff = fabs(diff - 2 * PI);
            if (diff > PI / 2) { W->driving_backward[c] = 1; step_reward[c] -= 0 * diff; }
            else W->driving_backward[c] = 0;
        }
        for (int c = 0; c < A; ++c) W->prev_reward[c] = W->reward[c];
        for (int c = 0; c < A; ++c) if (W->tile_visited_count[c] == W->T) done = 1;
        for (int c = 0; c < A; ++c) {
            double x = W->car[c].b[0].p.x, y = W->car[c].b[0].p.y;
            if (fabs(x) > PLAYFIELD || fabs(y) > PLAYFIELD) { done = 1; step_reward[c] = -100; }
        }
    }
    *done_out = done;
}

ORC_API void orc_render(OrcWorld* W, uint8_t* obs) {
    for (int c = 0; c < W->A; ++c) render_view(W, c, obs + (size_t)c * STATE_W * STATE_H * 3);
}
/* render(mode) for any viewport: state_pixels (96, 96), rgb_array (600, 400), mcr:566-575.  Skid
 * particles (drawn in the non-state modes, mcr:564) are drawn when orc_set_particles(1) (wide lines restated as D6). */
ORC_API void orc_render_vp(OrcWorld* W, int vw, int vh, uint8_t* out) {
    for (int c = 0; c < W->A; ++c) render_view_vp(W, c, out + (size_t)c * vw * vh * 3, vw, vh);
}

/* ------------------------------------------------------------------------------------ */
/* state access for parity tests                                                         */
/* ------------------------------------------------------------------------------------ */
/* bodies: A x 5 x 9 floats: p.x p.y angle v.x v.y w c.x c.y awake */
ORC_API void orc_get_bodies(const OrcWorld* W, float* out) {
    for (int c = 0; c < W->A; ++c)
        for (int i = 0; i < 5; ++i) {
            const Body* b = &W->car[c].b[i]; float* o = out + ((size_t)c * 5 + i) * 9;
            o[0] = b->p.x; o[1] = b->p.y; o[2] = b->a; o[3] = b->v.x; o[4] = b->v.y; o[5] = b->w;
            o[6] = b->c.x; o[7] = b->c.y; o[8] = (float)b->awake;
        }
}
/* wheels: A x 4 x 6 doubles: omega phase gas brake steer ntiles */
ORC_API void orc_get_wheels(const OrcWorld* W, double* out) {
    for (int c = 0; c < W->A; ++c)
        for (int w = 0; w < 4; ++w) {
            const Car* car = &W->car[c]; double* o = out + ((size_t)c * 4 + w) * 6;
            o[0] = car->omega[w]; o[1] = car->phase[w]; o[2] = car->gas[w]; o[3] = car->brake[w]; o[4] = car->steer[w];
            o[5] = (double)car->ntiles[w];
        }
}
/* joints: A x 4 x 6 floats: impulse.x impulse.y impulse.z motorImpulse limitState motorSpeed */
ORC_API void orc_get_joints(const OrcWorld* W, float* out) {
    for (int c = 0; c < W->A; ++c)
        for (int w = 0; w < 4; ++w) {
            const Joint* J = &W->car[c].j[w]; float* o = out + ((size_t)c * 4 + w) * 6;
            o[0] = J->impulse[0]; o[1] = J->impulse[1]; o[2] = J->impulse[2]; o[3] = J->motorImpulse;
            o[4] = (float)J->limitState; o[5] = J->motorSpeed;
        }
}
ORC_API void orc_get_visited(const OrcWorld* W, uint8_t* visited /*T*A*/, uint8_t* touched /*T*/) {
    memcpy(visited, W->visited, (size_t)W->T * W->A); memcpy(touched, W->touched, W->T);
}
ORC_API void orc_get_scores(const OrcWorld* W, double* reward, int* counts, uint8_t* backward) {
    for (int c = 0; c < W->A; ++c) { reward[c] = W->reward[c]; counts[c] = W->tile_visited_count[c]; backward[c] = W->driving_backward[c]; }
}
ORC_API void orc_set_particles(OrcWorld* W, int on) { W->draw_particles = on; }
ORC_API void orc_get_grass(const OrcWorld* W, uint8_t* grass) { for (int c = 0; c < W->A; ++c) grass[c] = W->driving_on_grass[c]; }
/* skid traces of car c: out[i] = (len, grass, 30 x (x, y)) floats per particle, returns the particle count */
ORC_API int orc_get_particles(const OrcWorld* W, int c, float* out) {
    const Car* car = &W->car[c];
    for (int i = 0; i < car->n_prt; ++i) {
        float* o = out + (size_t)i * 62; o[0] = (float)car->prt[i].len; o[1] = (float)car->prt[i].grass;
        for (int k = 0; k < PRT_PTS; ++k) { o[2 + 2 * k] = k < car->prt[i].len ? car->prt[i].pt[k].x : 0.0f; o[3 + 2 * k] = k < car->prt[i].len ? car->prt[i].pt[k].y : 0.0f; }
    }
    return car->n_prt;
}
ORC_API double orc_get_time(const OrcWorld* W) { return W->t; }
ORC_API int orc_get_fixed_point_iter(const OrcWorld* W) { return W->vel_iters_used; }
/* mass constants: hull mass, invMass, I, invI, lc.x, lc.y, wheel mass, invMass, I, invI, lc.x, lc.y */
ORC_API void orc_get_mass(const OrcWorld* W, float* out) {
    out[0] = W->hull_mass; out[1] = W->hull_invMass; out[2] = W->hull_I; out[3] = W->hull_invI; out[4] = W->hull_lc.x; out[5] = W->hull_lc.y;
    out[6] = W->wheel_mass; out[7] = W->wheel_invMass; out[8] = W->wheel_I; out[9] = W->wheel_invI; out[10] = W->wheel_lc.x; out[11] = W->wheel_lc.y;
}
/* polygon dump: which = 0..3 hull fixtures, 4 wheel; returns vertex count */
ORC_API int orc_get_shape(const OrcWorld* W, int which, float* out_xy) {
    const Poly* P = which < 4 ? &W->hull_poly[which] : &W->wheel_poly;
    for (int i = 0; i < P->n; ++i) { out_xy[2 * i] = P->v[i].x; out_xy[2 * i + 1] = P->v[i].y; }
    return P->n;
}
ORC_API int orc_get_tile_poly(const OrcWorld* W, int t, float* out_xy) {
    const Poly* P = &W->tile_poly[t];
    for (int i = 0; i < P->n; ++i) { out_xy[2 * i] = P->v[i].x; out_xy[2 * i + 1] = P->v[i].y; }
    return P->n;
}
/* direct state injection (used by the stubbed-reference harness and by tests) */
ORC_API void orc_set_backward(OrcWorld* W, const uint8_t* flags) { for (int c = 0; c < W->A; ++c) W->driving_backward[c] = flags[c]; }
ORC_API void orc_set_hull_color(OrcWorld* W, int car, int palette_index) { W->car[car].hull_color = palette_index; }

/* ------------------------------------------------------------------------------------ */
/* Hooks for tests/golden/make_golden.py: the UNMODIFIED reference module runs on stub      */
/* Box2D / gym / pyglet modules whose arithmetic is this file, so the reference's own      */
/* Python (reward rule, spawn, draw lists, HUD geometry, RNG order) pins the restatement.   */
/* ------------------------------------------------------------------------------------ */
ORC_API void orc_ext_steer(OrcWorld* W, int c, double s) { W->car[c].steer[0] = s; W->car[c].steer[1] = s; }
ORC_API void orc_ext_gas(OrcWorld* W, int c, double gas) {
    double g = gas < 0 ? 0 : (gas > 1 ? 1 : gas);
    for (int w = 2; w < 4; ++w) { double diff = g - W->car[c].gas[w]; if (diff > 0.1) diff = 0.1; W->car[c].gas[w] += diff; }
}
ORC_API void orc_ext_brake(OrcWorld* W, int c, double b) { for (int w = 0; w < 4; ++w) W->car[c].brake[w] = b; }
/* Car.step(dt) with len(wheel.tiles) supplied by the caller (the reference's listener owns the sets) */
ORC_API void orc_ext_car_step(OrcWorld* W, int c, const int* ntiles4, double dt) {
    for (int w = 0; w < 4; ++w) W->car[c].ntiles[w] = ntiles4[w];
    car_step(&W->car[c], dt);
}
/* touching pairs in contact-list order: out[i] = {tile, car, fixture, active} */
ORC_API int orc_ext_collide_pairs(OrcWorld* W, int* out, int max) {
    static __thread Pair pairs[MAX_PAIRS];
    int n = collect_pairs(W, pairs, max < MAX_PAIRS ? max : MAX_PAIRS);
    for (int i = 0; i < n; ++i) { out[4 * i] = pairs[i].tile; out[4 * i + 1] = pairs[i].car; out[4 * i + 2] = pairs[i].fixture; out[4 * i + 3] = pairs[i].active; }
    return n;
}
ORC_API void orc_ext_solve(OrcWorld* W, double dt, int velIters, int posIters) { world_solve(W, (float)dt, velIters, posIters); }
/* GL fixed-function polygon fill on a 96x96 RGB canvas (row 0 = top), viewport pixel coordinates */
ORC_API void orc_raster_fill(uint8_t* img, const float* px, const float* py, int n, float r, float g, float b) {
    Canvas cv; cv.img = img; cv.w = STATE_W; cv.h = STATE_H;
    fill_poly(&cv, px, py, n, rgbf(r, g, b));
}
ORC_API void orc_raster_fill_u8(uint8_t* img, const float* px, const float* py, int n, int r, int g, int b) {
    Canvas cv; cv.img = img; cv.w = STATE_W; cv.h = STATE_H; RGB c; c.r = (uint8_t)r; c.g = (uint8_t)g; c.b = (uint8_t)b;
    fill_poly(&cv, px, py, n, c);
}
ORC_API void orc_raster_text(uint8_t* img, const char* text) { draw_label(img, text); }

/* manifold readback for parity tests: out[i] = carA, fixA, carB, fixB, type, pointCount, then per point
 * (id, localPoint.x, localPoint.y, normalImpulse, tangentImpulse) x 2, localNormal.xy, localPoint.xy -> 20 floats */
ORC_API int orc_get_manifolds(const OrcWorld* W, float* out, int max) {
    int n = W->n_manifolds < max ? W->n_manifolds : max;
    for (int i = 0; i < n; ++i) {
        const Manifold* m = &W->manifolds[i]; float* o = out + 20 * i;
        o[0] = (float)m->carA; o[1] = (float)m->fixA; o[2] = (float)m->carB; o[3] = (float)m->fixB; o[4] = (float)m->type; o[5] = (float)m->pointCount;
        for (int j = 0; j < 2; ++j) {
            o[6 + 5 * j] = j < m->pointCount ? (float)(m->pt[j].id & 0xffff) : 0.0f;
            o[7 + 5 * j] = j < m->pointCount ? m->pt[j].localPoint.x : 0.0f; o[8 + 5 * j] = j < m->pointCount ? m->pt[j].localPoint.y : 0.0f;
            o[9 + 5 * j] = j < m->pointCount ? m->pt[j].normalImpulse : 0.0f; o[10 + 5 * j] = j < m->pointCount ? m->pt[j].tangentImpulse : 0.0f;
        }
        o[16] = m->localNormal.x; o[17] = m->localNormal.y; o[18] = m->localPoint.x; o[19] = m->localPoint.y;
    }
    return W->n_manifolds;
}
