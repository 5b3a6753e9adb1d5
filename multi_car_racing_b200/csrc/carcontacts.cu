// carcontacts.cu -- car-car rigid contacts: Box2D's narrow phase for the solid fixture pairs and
// the coupled-island solve.
//
// Replaces the part of world.Step(1/50, 180, 60) (reference multi_car_racing.py:428) that only
// matters when cars touch: b2CollidePolygons (b2FindMaxSeparation, b2FindIncidentEdge,
// b2ClipSegmentToLine), b2Contact::Update's impulse matching by feature id, and b2ContactSolver
// (friction + normal constraints with the 2-point block solver, position correction) running
// inside the same 180 velocity / <= 60 position iterations as the cars' revolute joints.
// Filtering (gym car_dynamics): hull-hull and wheel-(other car's) hull collide, wheel-wheel not.
//
//   carcontacts_kernel  one warp per env.  Lanes enumerate the (car a < car b, fixture, fixture)
//                       pairs in a fixed order, AABB-reject, run b2CollidePolygons on survivors,
//                       carry the previous step's impulses by contact id, and compact the touching
//                       manifolds in order (ballot prefix) into the env's manifold list.
//   coupled_kernel      one warp per env that has >= 1 manifold, lane = car.  Joints of different
//                       cars never share a body, so the lanes sweep their own car's joints in
//                       lock step (identical to Box2D's sequential order), then lane 0 solves the
//                       contacts on a shared-memory mirror of the body velocities.  Islands (cars
//                       linked by manifolds) keep their own position-iteration exit and sleep
//                       decision through lane masks.  Envs without manifolds never enter here:
//                       their cars take the per-car fast path of sim.cu.
#include "solver.cuh"
// Diagnostics build (-DMCR_PHASE_CLOCKS): lane 0 of every head_kernel warp adds the SM cycles of each part to g_head_clk
// (0-3 pre_car: loads, controls + tyre model, integrate + joints_init, stores; 4 entry + auto-reset check, 5 narrow phase; 7 = warps)
#ifdef MCR_PHASE_CLOCKS
__device__ unsigned long long g_head_clk[8];
#define PRE_CLK_T0() long long pc_t_ = clock64()
#define PRE_CLK(k) do { if ((threadIdx.x & 31) == 0) { const long long n_ = clock64(); atomicAdd(&g_head_clk[k], (unsigned long long)(n_ - pc_t_)); pc_t_ = n_; } } while (0)
#define HEAD_CLK_DECL long long hc_t_ = clock64()
#define HEAD_CLK(k) do { if ((threadIdx.x & 31) == 0) { const long long n_ = clock64(); atomicAdd(&g_head_clk[k], (unsigned long long)(n_ - hc_t_)); hc_t_ = n_; } } while (0)
extern "C" int mcr_debug_head_clocks(unsigned long long* out8, int reset) {
    if (out8 && cudaMemcpyFromSymbol(out8, g_head_clk, sizeof(g_head_clk)) != cudaSuccess) return -1;
    if (reset) { unsigned long long z[8] = {}; if (cudaMemcpyToSymbol(g_head_clk, z, sizeof(z)) != cudaSuccess) return -1; }
    return 0;
}
#else
#define HEAD_CLK_DECL do {} while (0)
#define HEAD_CLK(k) do {} while (0)
#endif
#include "pre.cuh"
#include "reset.cuh"

#define CC_WARPS 4
#define MAXM MCR_MAX_MANIFOLDS
#define MW MCR_MANIFOLD_WORDS

struct V2 { float x, y; };
__device__ __forceinline__ V2 mk(float x, float y) { V2 r; r.x = x; r.y = y; return r; }
__device__ __forceinline__ V2 operator+(V2 a, V2 b) { return mk(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ V2 operator-(V2 a, V2 b) { return mk(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ V2 operator*(float s, V2 a) { return mk(s * a.x, s * a.y); }
__device__ __forceinline__ V2 neg(V2 a) { return mk(-a.x, -a.y); }
__device__ __forceinline__ float dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
__device__ __forceinline__ float cross(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
__device__ __forceinline__ V2 cross_sv(float s, V2 a) { return mk(-s * a.y, s * a.x); }
__device__ __forceinline__ V2 cross_vs(V2 a, float s) { return mk(s * a.y, -s * a.x); }
struct XF { V2 p; float s, c; };   // b2Transform
__device__ __forceinline__ V2 rmul(float s, float c, V2 v) { return mk(c * v.x - s * v.y, s * v.x + c * v.y); }
__device__ __forceinline__ V2 rmulT(float s, float c, V2 v) { return mk(c * v.x + s * v.y, -s * v.x + c * v.y); }
__device__ __forceinline__ V2 xmul(const XF& x, V2 v) { return mk((x.c * v.x - x.s * v.y) + x.p.x, (x.s * v.x + x.c * v.y) + x.p.y); }
__device__ __forceinline__ V2 xmulT(const XF& x, V2 v) {
    const float px = v.x - x.p.x, py = v.y - x.p.y;
    return mk(x.c * px + x.s * py, -x.s * px + x.c * py);
}

enum { CF_VERTEX = 0, CF_FACE = 1 };
__device__ __forceinline__ uint32_t cf_key(int indexA, int indexB, int typeA, int typeB) {
    return (uint32_t)(indexA & 0xff) | ((uint32_t)(indexB & 0xff) << 8) | ((uint32_t)(typeA & 0xff) << 16) | ((uint32_t)(typeB & 0xff) << 24);
}
struct ClipVertex { V2 v; uint32_t id; };

struct Manifold {
    uint32_t key;            // carA | fixA << 8 | carB << 16 | fixB << 24
    int type, pointCount;    // 0 = e_faceA, 1 = e_faceB
    uint32_t id[2];
    V2 lp[2]; float ni[2], ti[2];
    V2 localNormal, localPoint;
};

__device__ __forceinline__ void manifold_store(float* w, const Manifold& m) {
    w[0] = __uint_as_float(m.key); w[1] = __int_as_float(m.type | (m.pointCount << 8));
    w[2] = __uint_as_float(m.id[0]); w[3] = __uint_as_float(m.id[1]);
    w[4] = m.lp[0].x; w[5] = m.lp[0].y; w[6] = m.ni[0]; w[7] = m.ti[0];
    w[8] = m.lp[1].x; w[9] = m.lp[1].y; w[10] = m.ni[1]; w[11] = m.ti[1];
    w[12] = m.localNormal.x; w[13] = m.localNormal.y; w[14] = m.localPoint.x; w[15] = m.localPoint.y;
}
__device__ __forceinline__ void manifold_load(const float* w, Manifold& m) {
    m.key = __float_as_uint(w[0]); const int tp = __float_as_int(w[1]); m.type = tp & 0xff; m.pointCount = tp >> 8;
    m.id[0] = __float_as_uint(w[2]); m.id[1] = __float_as_uint(w[3]);
    m.lp[0] = mk(w[4], w[5]); m.ni[0] = w[6]; m.ti[0] = w[7];
    m.lp[1] = mk(w[8], w[9]); m.ni[1] = w[10]; m.ti[1] = w[11];
    m.localNormal = mk(w[12], w[13]); m.localPoint = mk(w[14], w[15]);
}

// b2FindMaxSeparation (brute-force form of Box2D 2.3.1+)
__device__ float find_max_separation(int& edgeIndex, const Poly8& poly1, const XF& xf1, const Poly8& poly2, const XF& xf2) {
    const float qs = xf2.c * xf1.s - xf2.s * xf1.c, qc = xf2.c * xf1.c + xf2.s * xf1.s;   // b2MulT(xf2, xf1)
    XF xf; xf.s = qs; xf.c = qc; xf.p = rmulT(xf2.s, xf2.c, xf1.p - xf2.p);
    int bestIndex = 0; float maxSeparation = -3.402823466e+38f;
    for (int i = 0; i < poly1.n; ++i) {
        const V2 n = rmul(qs, qc, mk(poly1.nx[i], poly1.ny[i]));
        const V2 v1 = xmul(xf, mk(poly1.x[i], poly1.y[i]));
        float si = 3.402823466e+38f;
        for (int j = 0; j < poly2.n; ++j) {
            const float sij = dot(n, mk(poly2.x[j], poly2.y[j]) - v1);
            if (sij < si) si = sij;
        }
        if (si > maxSeparation) { maxSeparation = si; bestIndex = i; }
    }
    edgeIndex = bestIndex;
    return maxSeparation;
}

__device__ int clip_segment_to_line(ClipVertex vOut[2], const ClipVertex vIn[2], V2 normal, float offset, int vertexIndexA) {
    int numOut = 0;
    const float distance0 = dot(normal, vIn[0].v) - offset;
    const float distance1 = dot(normal, vIn[1].v) - offset;
    if (distance0 <= 0.0f) vOut[numOut++] = vIn[0];
    if (distance1 <= 0.0f) vOut[numOut++] = vIn[1];
    if (distance0 * distance1 < 0.0f) {
        const float interp = distance0 / (distance0 - distance1);
        vOut[numOut].v = vIn[0].v + interp * (vIn[1].v - vIn[0].v);
        vOut[numOut].id = cf_key(vertexIndexA, (int)((vIn[0].id >> 8) & 0xff), CF_VERTEX, CF_FACE);
        ++numOut;
    }
    return numOut;
}

// b2CollidePolygons
__device__ void collide_polygons(Manifold& m, const Poly8& polyA, const XF& xfA, const Poly8& polyB, const XF& xfB) {
    m.pointCount = 0;
    const float totalRadius = B2_POLYGON_RADIUS + B2_POLYGON_RADIUS;
    int edgeA = 0; const float separationA = find_max_separation(edgeA, polyA, xfA, polyB, xfB);
    if (separationA > totalRadius) return;
    int edgeB = 0; const float separationB = find_max_separation(edgeB, polyB, xfB, polyA, xfA);
    if (separationB > totalRadius) return;
    const float k_tol = 0.1f * B2_LINEAR_SLOP;
    const bool flip = separationB > separationA + k_tol;
    const Poly8& poly1 = flip ? polyB : polyA; const Poly8& poly2 = flip ? polyA : polyB;
    const XF xf1 = flip ? xfB : xfA, xf2 = flip ? xfA : xfB;
    const int edge1 = flip ? edgeB : edgeA;
    m.type = flip ? 1 : 0;
    // b2FindIncidentEdge
    ClipVertex incidentEdge[2];
    {
        const V2 normal1 = rmulT(xf2.s, xf2.c, rmul(xf1.s, xf1.c, mk(poly1.nx[edge1], poly1.ny[edge1])));
        int index = 0; float minDot = 3.402823466e+38f;
        for (int i = 0; i < poly2.n; ++i) {
            const float dd = dot(normal1, mk(poly2.nx[i], poly2.ny[i]));
            if (dd < minDot) { minDot = dd; index = i; }
        }
        const int i1 = index, i2 = i1 + 1 < poly2.n ? i1 + 1 : 0;
        incidentEdge[0].v = xmul(xf2, mk(poly2.x[i1], poly2.y[i1])); incidentEdge[0].id = cf_key(edge1, i1, CF_FACE, CF_VERTEX);
        incidentEdge[1].v = xmul(xf2, mk(poly2.x[i2], poly2.y[i2])); incidentEdge[1].id = cf_key(edge1, i2, CF_FACE, CF_VERTEX);
    }
    const int iv1 = edge1, iv2 = edge1 + 1 < poly1.n ? edge1 + 1 : 0;
    V2 v11 = mk(poly1.x[iv1], poly1.y[iv1]), v12 = mk(poly1.x[iv2], poly1.y[iv2]);
    V2 localTangent = v12 - v11;
    {
        const float len = sqrtf(localTangent.x * localTangent.x + localTangent.y * localTangent.y);
        if (len >= B2_EPS) { const float inv = 1.0f / len; localTangent.x *= inv; localTangent.y *= inv; }
    }
    const V2 localNormal = cross_vs(localTangent, 1.0f);
    const V2 planePoint = 0.5f * (v11 + v12);
    const V2 tangent = rmul(xf1.s, xf1.c, localTangent);
    const V2 normal = cross_vs(tangent, 1.0f);
    v11 = xmul(xf1, v11); v12 = xmul(xf1, v12);
    const float frontOffset = dot(normal, v11);
    const float sideOffset1 = -dot(tangent, v11) + totalRadius;
    const float sideOffset2 = dot(tangent, v12) + totalRadius;
    ClipVertex clipPoints1[2], clipPoints2[2];
    int np = clip_segment_to_line(clipPoints1, incidentEdge, neg(tangent), sideOffset1, iv1);
    if (np < 2) return;
    np = clip_segment_to_line(clipPoints2, clipPoints1, tangent, sideOffset2, iv2);
    if (np < 2) return;
    m.localNormal = localNormal; m.localPoint = planePoint;
    int pointCount = 0;
    for (int i = 0; i < 2; ++i) {
        const float separation = dot(normal, clipPoints2[i].v) - frontOffset;
        if (separation <= totalRadius) {
            m.lp[pointCount] = xmulT(xf2, clipPoints2[i].v);
            uint32_t id = clipPoints2[i].id;
            if (flip) id = cf_key((int)((id >> 8) & 0xff), (int)(id & 0xff), (int)((id >> 24) & 0xff), (int)((id >> 16) & 0xff));
            m.id[pointCount] = id;
            ++pointCount;
        }
    }
    m.pointCount = pointCount;
}

__device__ __forceinline__ XF body_xf(const DevBuffers& b, int N, int car, int body) {
    const float* bp = b.body + (size_t)(body * BODY_FIELDS) * N + car;
    XF x; x.p = mk(bp[(size_t)BF_PX * N], bp[(size_t)BF_PY * N]); x.s = bp[(size_t)BF_QS * N]; x.c = bp[(size_t)BF_QC * N];
    return x;
}

// Narrow phase of one env by one warp (lanes over the fixture pairs of every car pair); s_old = this
// warp's MAXM * MW floats of shared memory.
// MANY = more than three cars per env: the car pairs worth a fixture-pair sweep are selected first.  The two
// variants are separate kernels on purpose: the step's head is a cold-code latency chain, and the selection
// code costs the two-car configuration 3 us when it is merely present.
template <bool MANY>
__device__ void carcontacts_warp(const Dims& d, const DevBuffers& b, const CarConst& cc, int env, int lane, float* s_old_w) {
    const int A = d.A, N = d.N;
    float* gman = b.manifold + (size_t)env * MAXM * MW;
    if (A < 2) { if (lane == 0) b.n_manifold[env] = 0; return; }
    const int nold = b.n_manifold[env];
    const int ncp = A * (A - 1) / 2;
    if (nold == 0) {
        // The common case costs two loads: with no manifold to carry over, fixtures can only touch if two
        // hull origins are within reach of each other.  Every fixture vertex lies within 2.9 m of its car's
        // hull origin (hull polygon <= 2.87 m, wheel anchor + wheel half diagonal + joint slack <= 2.7 m), so
        // 7 m between origins rules every pair out (2 x 2.9 + polygon radii + a wide margin).
        bool close = false;
        for (int cp = lane; cp < ncp; cp += 32) {
            int a = 0, rem = cp;
            while (rem >= A - 1 - a) { rem -= A - 1 - a; ++a; }
            const int carA = env * A + a, carB = env * A + a + 1 + rem;
            const float dx = b.body[(size_t)BF_PX * N + carA] - b.body[(size_t)BF_PX * N + carB];
            const float dy = b.body[(size_t)BF_PY * N + carA] - b.body[(size_t)BF_PY * N + carB];
            close = close || !(dx * dx + dy * dy > 49.0f);
        }
        if (!__any_sync(0xffffffffu, close)) { if (lane == 0) b.n_manifold[env] = 0; return; }
    }
    for (int i = lane; i < nold * MW; i += 32) s_old_w[i] = gman[i];
    __syncwarp();
    const float r = B2_POLYGON_RADIUS;
    // Car pairs worth a fixture-pair sweep, in pair order: hull origins within reach (see above), or a manifold
    // of the pair to carry over (sleeping bodies keep theirs wherever they are).  With 8 or 16 cars per env most
    // of the A (A - 1) / 2 pairs are far apart.
    uint32_t pair_bits[4] = {0u, 0u, 0u, 0u};          // MCR_MAX_AGENTS = 16 -> at most 120 pairs
    if (!MANY) pair_bits[0] = (1u << ncp) - 1u;        // 2 or 3 cars: every pair
    else for (int c0 = 0; c0 < ncp; c0 += 32) {
        const int cp = c0 + lane;
        bool keep = false;
        if (cp < ncp) {
            int a = 0, rem = cp;
            while (rem >= A - 1 - a) { rem -= A - 1 - a; ++a; }
            const int bc = a + 1 + rem, carA = env * A + a, carB = env * A + bc;
            const float dx = b.body[(size_t)BF_PX * N + carA] - b.body[(size_t)BF_PX * N + carB];
            const float dy = b.body[(size_t)BF_PY * N + carA] - b.body[(size_t)BF_PY * N + carB];
            keep = !(dx * dx + dy * dy > 49.0f);
            for (int i = 0; i < nold && !keep; ++i) {
                const uint32_t key = __float_as_uint(s_old_w[i * MW]);
                keep = (int)(key & 0xff) == a && (int)((key >> 16) & 0xff) == bc;
            }
        }
        pair_bits[c0 >> 5] = __ballot_sync(0xffffffffu, keep);
    }
    const int nclose = __popc(pair_bits[0]) + __popc(pair_bits[1]) + __popc(pair_bits[2]) + __popc(pair_bits[3]);
    int nnew = 0;
    for (int base = 0; base < nclose * 64; base += 32) {
        const int idx = base + lane;
        const int rank = idx >> 6, fa = (idx >> 3) & 7, fb = idx & 7;
        // the rank-th kept pair -> its pair index cp (uniform over the two 32-lane halves of a pair's 64 tests)
        int cp = ncp;
        if (!MANY) cp = rank;
        else {
            int left = rank;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const int cnt = __popc(pair_bits[w]);
                if (cp == ncp && left >= 0 && left < cnt) {
                    uint32_t m = pair_bits[w];
                    for (int q = 0; q < left; ++q) m &= m - 1u;      // drop the `left` lowest kept pairs
                    cp = 32 * w + (__ffs((int)m) - 1);
                }
                left -= cnt;
            }
        }
        // decode the car pair (a < b) from its rank cp in the order (0,1),(0,2),...,(1,2),...
        int a = 0, rem = cp;
        if (cp < ncp) { while (rem >= A - 1 - a) { rem -= A - 1 - a; ++a; } }
        const int bcar = a + 1 + rem;
        Manifold m; m.pointCount = 0;
        bool emit = false;
        if (cp < ncp && !(fa < 4 && fb < 4)) {
            const int carA = env * A + a, carB = env * A + bcar;
            const int bodyA = fa < 4 ? 1 + fa : 0, bodyB = fb < 4 ? 1 + fb : 0;
            const Poly8& PA = fa < 4 ? cc.wheel_poly : cc.hull_poly[fa - 4];
            const Poly8& PB = fb < 4 ? cc.wheel_poly : cc.hull_poly[fb - 4];
            const uint32_t key = (uint32_t)a | ((uint32_t)fa << 8) | ((uint32_t)bcar << 16) | ((uint32_t)fb << 24);
            int prev = -1;
            for (int i = 0; i < nold; ++i) if (__float_as_uint(s_old_w[i * MW]) == key) { prev = i; break; }
            // wheels are always awake at Collide time (Car.step woke them); hull flags are the start-of-step ones
            const bool awakeA = fa < 4 ? true : (b.awake[carA] != 0), awakeB = fb < 4 ? true : (b.awake[carB] != 0);
            if (!awakeA && !awakeB) {
                if (prev >= 0) { manifold_load(&s_old_w[prev * MW], m); emit = m.pointCount > 0; }
            } else {
                const XF xfA = body_xf(b, N, carA, bodyA), xfB = body_xf(b, N, carB, bodyB);
                float alx = 3.402823466e+38f, aly = alx, ahx = -alx, ahy = -alx, blx = alx, bly = alx, bhx = -alx, bhy = -alx;
                for (int i = 0; i < PA.n; ++i) {
                    const V2 v = xmul(xfA, mk(PA.x[i], PA.y[i]));
                    alx = fminf(alx, v.x); aly = fminf(aly, v.y); ahx = fmaxf(ahx, v.x); ahy = fmaxf(ahy, v.y);
                }
                for (int i = 0; i < PB.n; ++i) {
                    const V2 v = xmul(xfB, mk(PB.x[i], PB.y[i]));
                    blx = fminf(blx, v.x); bly = fminf(bly, v.y); bhx = fmaxf(bhx, v.x); bhy = fmaxf(bhy, v.y);
                }
                const bool apart = (blx - r) - (ahx + r) > 0.0f || (bly - r) - (ahy + r) > 0.0f ||
                                   (alx - r) - (bhx + r) > 0.0f || (aly - r) - (bhy + r) > 0.0f;
                if (!apart) {
                    m.key = key;
                    collide_polygons(m, PA, xfA, PB, xfB);
                    for (int i = 0; i < m.pointCount; ++i) {          // b2Contact::Update: impulses follow the feature id
                        m.ni[i] = 0.0f; m.ti[i] = 0.0f;
                        if (prev >= 0) {
                            Manifold o; manifold_load(&s_old_w[prev * MW], o);
                            for (int j = 0; j < o.pointCount; ++j) if (o.id[j] == m.id[i]) { m.ni[i] = o.ni[j]; m.ti[i] = o.ti[j]; break; }
                        }
                    }
                    emit = m.pointCount > 0;
                }
            }
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, emit);
        if (emit) {
            const int pos = nnew + __popc(bal & ((1u << lane) - 1u));
            if (pos < MAXM) {
                if (m.pointCount < 2) { m.id[1] = 0u; m.lp[1] = mk(0.0f, 0.0f); m.ni[1] = 0.0f; m.ti[1] = 0.0f; }
                manifold_store(gman + (size_t)pos * MW, m);
            } else atomicExch(&b.status[ST_MANIFOLD_OVERFLOW], 1);
        }
        nnew += __popc(bal);
    }
    if (lane == 0) b.n_manifold[env] = nnew < MAXM ? nnew : MAXM;
}

__global__ void __launch_bounds__(CC_WARPS * 32)
carcontacts_kernel(Dims d, DevBuffers b, CarConst cc, const uint8_t* __restrict__ mask) {
    __shared__ float s_old[CC_WARPS][MAXM * MW];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int env = blockIdx.x * CC_WARPS + warp;
    if (env >= d.B) return;
    if (mask && !mask[env]) return;
    if (d.A > 3) carcontacts_warp<true>(d, b, cc, env, lane, s_old[warp]); else carcontacts_warp<false>(d, b, cc, env, lane, s_old[warp]);
}

// The head of mcr_step's pipeline as ONE launch, one warp per env: the next-step auto reset of the
// envs whose previous step ended the episode (reset.cuh), the car-car narrow phase (above), then the
// per-car head of the step (pre.cuh) on lanes 0 .. A-1.  Three dependent latency-bound launches
// (auto_reset -> carcontacts -> pre, 3 + 7 + 9 us plus two launch gaps) become one.
template <typename ActT, bool MANY>
__global__ void __launch_bounds__(CC_WARPS * 32, 2)      // (two CTAs per SM hold the bench batch; the register cap this implies leaves the per-car head without spills)
head_kernel(Dims d, DevBuffers b, CarConst cc, const uint8_t* __restrict__ mask, const uint8_t* __restrict__ reset_flags,
            AutoResetCfg ar, const ActT* __restrict__ action, int collisions) {
    __shared__ float s_old[CC_WARPS][MAXM * MW];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int env = blockIdx.x * CC_WARPS + warp;
    if (blockIdx.x == 0 && threadIdx.x == 0) b.ready[READY_EPOCH(d.N)] += 1;      // a new pass, see pre_kernel
    if (env >= d.B) return;
    if (mask && !mask[env]) return;
    bool respawned = false;
    tl_stamp(b.timeline, TL_HEAD);
    HEAD_CLK_DECL;
    // The control loads of the three stages, issued together (they used to be dependent round trips: reset flag ->
    // manifold count -> hull origins, each from DRAM when the L2 is cold), and the lines pre_car will read prefetched
    // into the L2 beside them.
    const int car_l = env * d.A + (lane < d.A ? lane : 0);
    // (volatile asm: the loads are issued HERE -- the compiler otherwise sinks them to their first use, behind the reset check)
    unsigned rf_u = 0u; int nold_early = 0; float hx_early = 0.0f, hy_early = 0.0f, hs_early = 0.0f, hc_early = 1.0f;
    const bool cc_on = collisions && d.A > 1;
    if (reset_flags) asm volatile("ld.global.u8 %0, [%1];" : "=r"(rf_u) : "l"(reset_flags + env));
    if (cc_on) {
        asm volatile("ld.global.s32 %0, [%1];" : "=r"(nold_early) : "l"(b.n_manifold + env));
        if (lane < d.A) {
            asm volatile("ld.global.f32 %0, [%1];" : "=f"(hx_early) : "l"(b.body + (size_t)BF_PX * d.N + car_l));
            asm volatile("ld.global.f32 %0, [%1];" : "=f"(hy_early) : "l"(b.body + (size_t)BF_PY * d.N + car_l));
            asm volatile("ld.global.f32 %0, [%1];" : "=f"(hs_early) : "l"(b.body + (size_t)BF_QS * d.N + car_l));
            asm volatile("ld.global.f32 %0, [%1];" : "=f"(hc_early) : "l"(b.body + (size_t)BF_QC * d.N + car_l));
        }
    }
    const uint8_t rf_early = (uint8_t)rf_u;
    if (warp == 0) {
        // one warp per CTA prefetches for the CTA's CC_WARPS envs (consecutive cars: the same 128-byte line of every field,
        // but for a batch boundary), one field per lane and round -- every redundant request queues in front of real loads
        const size_t N_ = (size_t)d.N;
        const int car0 = env * d.A;
#define PF_L2(p) asm volatile("prefetch.global.L2 [%0];" :: "l"(p))
        for (int f = lane; f < 5 * BODY_FIELDS; f += 32) PF_L2(b.body + f * N_ + car0);
        if (lane < 4 * JOINT_FIELDS) PF_L2(b.joint + lane * N_ + car0);
        if (lane < 4 * WHEEL_FIELDS) PF_L2(b.wheel + lane * N_ + car0);
        if (lane < CTRL_FIELDS) PF_L2(b.ctrl + lane * N_ + car0);
        if (lane < 5) { PF_L2(b.sleep_time + lane * N_ + car0); PF_L2(b.awake + lane * N_ + car0); }
        if (lane < 4) { PF_L2(b.limit_state + lane * N_ + car0); PF_L2(b.on_road + lane * N_ + car0); }
        if (action && lane == 0) PF_L2(action + (size_t)car0 * 3);
#undef PF_L2
    }
    if (reset_flags) {
        respawned = rf_early != 0;
        if (lane == 0) b.reset_mask[env] = respawned ? 1 : 0;
        if (respawned) {
            const uint32_t episode = b.env_episode[env] + 1u;
            __syncwarp();                              // every lane has read the counter before lane 0 bumps it
            auto_reset_env(d, b, cc, ar, env, episode, lane, 32);
        }
        __syncwarp();                                  // the respawned poses are visible to every lane
    }
    HEAD_CLK(4);
    int coupled_known = cc_on ? -1 : 0;
    if (cc_on) {
        // The common case is decided right here, on the preloaded values: no manifold to carry over and every pair of
        // hull origins more than 7 m apart (see carcontacts_warp) -- no fixture can touch.  carcontacts_warp is a real
        // call whose struct arguments are copied to local memory first (~1.4 KB per thread: 5.8 of the kernel's 9.7 us
        // when it was called for every env, scripts/head_phases.py); only the envs that need it pay for that now.
        bool need = respawned || nold_early != 0;      // (a respawned env was just rewritten: the narrow phase loads it itself)
        if (!__any_sync(0xffffffffu, need)) {
            const int A = d.A, ncp = A * (A - 1) / 2;
            bool close = false;
            for (int cp0 = 0; cp0 < ncp; cp0 += 32) {
                const int cp = cp0 + lane;
                int a = 0, rem = cp < ncp ? cp : 0;
                while (rem >= A - 1 - a) { rem -= A - 1 - a; ++a; }
                const int bc = a + 1 + rem;
                const float ax = __shfl_sync(0xffffffffu, hx_early, a), ay = __shfl_sync(0xffffffffu, hy_early, a);
                const float bx = __shfl_sync(0xffffffffu, hx_early, bc), by = __shfl_sync(0xffffffffu, hy_early, bc);
                const float as_ = __shfl_sync(0xffffffffu, hs_early, a), ac = __shfl_sync(0xffffffffu, hc_early, a);
                const float bs = __shfl_sync(0xffffffffu, hs_early, bc), bcq = __shfl_sync(0xffffffffu, hc_early, bc);
                const float dx = bx - ax, dy = by - ay;
                bool near = !(dx * dx + dy * dy > 49.0f);
                if (near) {
                    // Inside 7 m: the two cars' bounding rectangles in their hull frames.  Every fixture point lies within
                    // |x| <= 1.71 (wheel anchor 1.1 + wheel half diagonal 0.61), -2.4 <= y <= 2.6 (hull polygons) of the
                    // hull origin; with 0.5 m for joint slack, polygon radii and rounding: half extents 2.21 x 3.1.  Cars on
                    // the start grid stand 6.67 m apart side by side -- inside the 7 m circle test for the first part of
                    // every episode, which sent most envs through the fixture-pair sweep (5 of the head's 8 us).
                    const float EX = CAR_OBB_EX, EY = CAR_OBB_EY;
                    // axes: A's (ac, as_), (-as_, ac); B's (bcq, bs), (-bs, bcq)
                    const float c = ac * bcq + as_ * bs, sn = ac * bs - as_ * bcq;      // cos / sin of the relative angle
                    const float ca = fabsf(c), sa = fabsf(sn);
                    const float dAx = dx * ac + dy * as_, dAy = -dx * as_ + dy * ac;     // centre offset in A's frame
                    const float dBx = dx * bcq + dy * bs, dBy = -dx * bs + dy * bcq;     // ... in B's frame
                    const bool sep = fabsf(dAx) > EX + (EX * ca + EY * sa) || fabsf(dAy) > EY + (EX * sa + EY * ca) ||
                                     fabsf(dBx) > EX + (EX * ca + EY * sa) || fabsf(dBy) > EY + (EX * sa + EY * ca);
                    near = !sep;
                }
                close = close || (cp < ncp && near);
            }
            need = __any_sync(0xffffffffu, close);
        }
        if (need) carcontacts_warp<MANY>(d, b, cc, env, lane, s_old[warp]);
        else if (lane == 0) b.n_manifold[env] = 0;
        coupled_known = need ? -1 : 0;                 // (the common case: pre_car4 does not wait for the store above)
    }
    __syncwarp();                                      // n_manifold[env]
    HEAD_CLK(5);
    // a respawned env takes reset()'s implicit step(None) (mcr:408): its action is ignored
    if (4 * d.A <= 32) {
        // four lanes per car (wheel / joint k = lane & 3): a quarter of the wheel and joint code per warp
        const unsigned gmask = 4 * d.A == 32 ? 0xffffffffu : ((1u << (4 * d.A)) - 1u);
        if (lane < 4 * d.A) pre_car4<ActT>(d, b, cc, env * d.A + (lane >> 2), env, lane & 3, gmask, action != nullptr && !respawned, action, coupled_known);
    } else if (lane < d.A) pre_car<ActT>(d, b, cc, env * d.A + lane, env, action != nullptr && !respawned, action);
    if (threadIdx.x == 0) atomicMax(b.timeline + TL_HEAD_END, mcr_globaltimer());
#ifdef MCR_PHASE_CLOCKS
    if (lane == 0) atomicAdd(&g_head_clk[7], 1ull);
#endif
}

// ---------------------------------------------------------------------------------------
// coupled island solve
// ---------------------------------------------------------------------------------------
struct BodyMirror { float vx, vy, w, cx, cy, a; };
struct VC {               // b2ContactVelocityConstraint + what the position solver needs
    int bA, bB;           // body slots in the mirror: car * 5 + body
    int root;             // island (lowest car index)
    int type, pointCount, manifoldPoints;
    float nx, ny, friction;
    float rAx[2], rAy[2], rBx[2], rBy[2], normalMass[2], tangentMass[2], ni[2], ti[2];
    float k11, k12, k22, n11, n12, n21, n22;
    float mA, iA, mB, iB, lcAx, lcAy, lcBx, lcBy;
    V2 localNormal, localPoint, lp[2];
};

__device__ void contact_init(VC& vc, const Manifold& m, const BodyMirror* bm, const CarConst& cc) {
    const int fa = (m.key >> 8) & 0xff, fb = (m.key >> 24) & 0xff, ca = m.key & 0xff, cb = (m.key >> 16) & 0xff;
    vc.bA = ca * 5 + (fa < 4 ? 1 + fa : 0); vc.bB = cb * 5 + (fb < 4 ? 1 + fb : 0);
    vc.mA = fa < 4 ? cc.wheel_invMass : cc.hull_invMass; vc.iA = fa < 4 ? cc.wheel_invI : cc.hull_invI;
    vc.mB = fb < 4 ? cc.wheel_invMass : cc.hull_invMass; vc.iB = fb < 4 ? cc.wheel_invI : cc.hull_invI;
    vc.lcAx = fa < 4 ? 0.0f : cc.hull_lcx; vc.lcAy = fa < 4 ? 0.0f : cc.hull_lcy;
    vc.lcBx = fb < 4 ? 0.0f : cc.hull_lcx; vc.lcBy = fb < 4 ? 0.0f : cc.hull_lcy;
    vc.type = m.type; vc.pointCount = m.pointCount; vc.manifoldPoints = m.pointCount;
    vc.localNormal = m.localNormal; vc.localPoint = m.localPoint; vc.lp[0] = m.lp[0]; vc.lp[1] = m.lp[1];
    const BodyMirror& A = bm[vc.bA]; const BodyMirror& B = bm[vc.bB];
    const float mA = vc.mA, mB = vc.mB, iA = vc.iA, iB = vc.iB;
    XF xfA, xfB;
    rot_set(A.a, xfA.s, xfA.c); rot_set(B.a, xfB.s, xfB.c);
    xfA.p = mk(A.cx, A.cy) - rmul(xfA.s, xfA.c, mk(vc.lcAx, vc.lcAy));
    xfB.p = mk(B.cx, B.cy) - rmul(xfB.s, xfB.c, mk(vc.lcBx, vc.lcBy));
    // b2WorldManifold::Initialize
    const float radiusA = B2_POLYGON_RADIUS, radiusB = B2_POLYGON_RADIUS;
    V2 normal, points[2];
    if (m.type == 0) {
        normal = rmul(xfA.s, xfA.c, m.localNormal);
        const V2 planePoint = xmul(xfA, m.localPoint);
        for (int i = 0; i < m.pointCount; ++i) {
            const V2 clipPoint = xmul(xfB, m.lp[i]);
            const V2 cA = clipPoint + (radiusA - dot(clipPoint - planePoint, normal)) * normal;
            const V2 cB = clipPoint - radiusB * normal;
            points[i] = 0.5f * (cA + cB);
        }
    } else {
        normal = rmul(xfB.s, xfB.c, m.localNormal);
        const V2 planePoint = xmul(xfB, m.localPoint);
        for (int i = 0; i < m.pointCount; ++i) {
            const V2 clipPoint = xmul(xfA, m.lp[i]);
            const V2 cB = clipPoint + (radiusB - dot(clipPoint - planePoint, normal)) * normal;
            const V2 cA = clipPoint - radiusA * normal;
            points[i] = 0.5f * (cA + cB);
        }
        normal = neg(normal);
    }
    vc.nx = normal.x; vc.ny = normal.y;
    vc.friction = sqrtf(0.2f * 0.2f);               // b2MixFriction of the default fixture frictions
    for (int j = 0; j < vc.pointCount; ++j) {
        vc.ni[j] = m.ni[j]; vc.ti[j] = m.ti[j];      // dtRatio == 1
        const V2 rA = points[j] - mk(A.cx, A.cy), rB = points[j] - mk(B.cx, B.cy);
        vc.rAx[j] = rA.x; vc.rAy[j] = rA.y; vc.rBx[j] = rB.x; vc.rBy[j] = rB.y;
        const float rnA = cross(rA, normal), rnB = cross(rB, normal);
        const float kNormal = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
        vc.normalMass[j] = kNormal > 0.0f ? 1.0f / kNormal : 0.0f;
        const V2 tangent = cross_vs(normal, 1.0f);
        const float rtA = cross(rA, tangent), rtB = cross(rB, tangent);
        const float kTangent = mA + mB + iA * rtA * rtA + iB * rtB * rtB;
        vc.tangentMass[j] = kTangent > 0.0f ? 1.0f / kTangent : 0.0f;
    }
    if (vc.pointCount == 2) {
        const V2 r1A = mk(vc.rAx[0], vc.rAy[0]), r1B = mk(vc.rBx[0], vc.rBy[0]), r2A = mk(vc.rAx[1], vc.rAy[1]), r2B = mk(vc.rBx[1], vc.rBy[1]);
        const float rn1A = cross(r1A, normal), rn1B = cross(r1B, normal), rn2A = cross(r2A, normal), rn2B = cross(r2B, normal);
        const float k11 = mA + mB + iA * rn1A * rn1A + iB * rn1B * rn1B;
        const float k22 = mA + mB + iA * rn2A * rn2A + iB * rn2B * rn2B;
        const float k12 = mA + mB + iA * rn1A * rn2A + iB * rn1B * rn2B;
        const float k_maxConditionNumber = 1000.0f;
        if (k11 * k11 < k_maxConditionNumber * (k11 * k22 - k12 * k12)) {
            vc.k11 = k11; vc.k12 = k12; vc.k22 = k22;
            const float a = k11, bq = k12, c = k12, dd = k22;
            float det = a * dd - bq * c;
            if (det != 0.0f) det = 1.0f / det;
            vc.n11 = det * dd; vc.n12 = -det * bq; vc.n21 = -det * c; vc.n22 = det * a;   // ex.x, ey.x, ex.y, ey.y
        } else {
            vc.pointCount = 1;
        }
    }
}

__device__ void contact_warm_start(const VC& vc, BodyMirror* bm) {
    BodyMirror& A = bm[vc.bA]; BodyMirror& B = bm[vc.bB];
    const V2 normal = mk(vc.nx, vc.ny), tangent = cross_vs(normal, 1.0f);
    for (int j = 0; j < vc.pointCount; ++j) {
        const V2 P = vc.ni[j] * normal + vc.ti[j] * tangent;
        A.w -= vc.iA * cross(mk(vc.rAx[j], vc.rAy[j]), P); A.vx -= vc.mA * P.x; A.vy -= vc.mA * P.y;
        B.w += vc.iB * cross(mk(vc.rBx[j], vc.rBy[j]), P); B.vx += vc.mB * P.x; B.vy += vc.mB * P.y;
    }
}

__device__ __forceinline__ void apply2(const VC& vc, V2& vA, float& wA, V2& vB, float& wB, float dx, float dy) {
    const V2 normal = mk(vc.nx, vc.ny);
    const V2 P1 = dx * normal, P2 = dy * normal;
    vA = vA - vc.mA * (P1 + P2);
    wA -= vc.iA * (cross(mk(vc.rAx[0], vc.rAy[0]), P1) + cross(mk(vc.rAx[1], vc.rAy[1]), P2));
    vB = vB + vc.mB * (P1 + P2);
    wB += vc.iB * (cross(mk(vc.rBx[0], vc.rBy[0]), P1) + cross(mk(vc.rBx[1], vc.rBy[1]), P2));
}

// b2ContactSolver::SolveVelocityConstraints for one manifold on the velocities of its two bodies (registers)
__device__ __forceinline__ void contact_solve_vel_reg(VC& vc, V2& vA, float& wA, V2& vB, float& wB) {
    const float mA = vc.mA, mB = vc.mB, iA = vc.iA, iB = vc.iB;
    const V2 normal = mk(vc.nx, vc.ny), tangent = cross_vs(normal, 1.0f);
    for (int j = 0; j < vc.pointCount; ++j) {
        const V2 rA = mk(vc.rAx[j], vc.rAy[j]), rB = mk(vc.rBx[j], vc.rBy[j]);
        const V2 dv = ((vB + cross_sv(wB, rB)) - vA) - cross_sv(wA, rA);
        const float vt = dot(dv, tangent) - 0.0f;
        float lambda = vc.tangentMass[j] * (-vt);
        const float maxFriction = vc.friction * vc.ni[j];
        const float newImpulse = clampf(vc.ti[j] + lambda, -maxFriction, maxFriction);
        lambda = newImpulse - vc.ti[j];
        vc.ti[j] = newImpulse;
        const V2 P = lambda * tangent;
        vA = vA - mA * P; wA -= iA * cross(rA, P);
        vB = vB + mB * P; wB += iB * cross(rB, P);
    }
    if (vc.pointCount == 1) {
        const V2 rA = mk(vc.rAx[0], vc.rAy[0]), rB = mk(vc.rBx[0], vc.rBy[0]);
        const V2 dv = ((vB + cross_sv(wB, rB)) - vA) - cross_sv(wA, rA);
        const float vn = dot(dv, normal);
        float lambda = -vc.normalMass[0] * (vn - 0.0f);
        const float newImpulse = fmaxf(vc.ni[0] + lambda, 0.0f);
        lambda = newImpulse - vc.ni[0];
        vc.ni[0] = newImpulse;
        const V2 P = lambda * normal;
        vA = vA - mA * P; wA -= iA * cross(rA, P);
        vB = vB + mB * P; wB += iB * cross(rB, P);
    } else {
        const float ax = vc.ni[0], ay = vc.ni[1];
        const V2 r1A = mk(vc.rAx[0], vc.rAy[0]), r1B = mk(vc.rBx[0], vc.rBy[0]), r2A = mk(vc.rAx[1], vc.rAy[1]), r2B = mk(vc.rBx[1], vc.rBy[1]);
        const V2 dv1 = ((vB + cross_sv(wB, r1B)) - vA) - cross_sv(wA, r1A);
        const V2 dv2 = ((vB + cross_sv(wB, r2B)) - vA) - cross_sv(wA, r2A);
        float vn1 = dot(dv1, normal), vn2 = dot(dv2, normal);
        float bx = vn1 - 0.0f, by = vn2 - 0.0f;
        bx -= vc.k11 * ax + vc.k12 * ay;
        by -= vc.k12 * ax + vc.k22 * ay;
        for (;;) {
            float xx = -(vc.n11 * bx + vc.n12 * by), xy = -(vc.n21 * bx + vc.n22 * by);
            if (xx >= 0.0f && xy >= 0.0f) { apply2(vc, vA, wA, vB, wB, xx - ax, xy - ay); vc.ni[0] = xx; vc.ni[1] = xy; break; }
            xx = -vc.normalMass[0] * bx; xy = 0.0f;
            vn1 = 0.0f; vn2 = vc.k12 * xx + by;
            if (xx >= 0.0f && vn2 >= 0.0f) { apply2(vc, vA, wA, vB, wB, xx - ax, xy - ay); vc.ni[0] = xx; vc.ni[1] = xy; break; }
            xx = 0.0f; xy = -vc.normalMass[1] * by;
            vn1 = vc.k12 * xy + bx; vn2 = 0.0f;
            if (xy >= 0.0f && vn1 >= 0.0f) { apply2(vc, vA, wA, vB, wB, xx - ax, xy - ay); vc.ni[0] = xx; vc.ni[1] = xy; break; }
            xx = 0.0f; xy = 0.0f; vn1 = bx; vn2 = by;
            if (vn1 >= 0.0f && vn2 >= 0.0f) { apply2(vc, vA, wA, vB, wB, xx - ax, xy - ay); vc.ni[0] = xx; vc.ni[1] = xy; break; }
            break;
        }
    }
}

__device__ void contact_solve_vel(VC& vc, BodyMirror* bm) {
    BodyMirror& A = bm[vc.bA]; BodyMirror& B = bm[vc.bB];
    V2 vA = mk(A.vx, A.vy), vB = mk(B.vx, B.vy); float wA = A.w, wB = B.w;
    contact_solve_vel_reg(vc, vA, wA, vB, wB);
    A.vx = vA.x; A.vy = vA.y; A.w = wA; B.vx = vB.x; B.vy = vB.y; B.w = wB;
}

__device__ void contact_solve_pos(const VC& vc, BodyMirror* bm, float& minSeparation) {
    BodyMirror& A = bm[vc.bA]; BodyMirror& B = bm[vc.bB];
    const float mA = vc.mA, mB = vc.mB, iA = vc.iA, iB = vc.iB;
    V2 cA = mk(A.cx, A.cy), cB = mk(B.cx, B.cy); float aA = A.a, aB = B.a;
    XF xfA, xfB;
    float aA_set = 0.0f, aB_set = 0.0f;
    for (int j = 0; j < vc.manifoldPoints; ++j) {
        // b2Rot::Set per point (fp64 sincos, D5): an angle the previous point left untouched (zero impulse: the point is
        // not penetrating) gives the same rotation bit for bit
        if (j == 0 || __float_as_uint(aA) != __float_as_uint(aA_set)) { rot_set(aA, xfA.s, xfA.c); aA_set = aA; }
        if (j == 0 || __float_as_uint(aB) != __float_as_uint(aB_set)) { rot_set(aB, xfB.s, xfB.c); aB_set = aB; }
        xfA.p = cA - rmul(xfA.s, xfA.c, mk(vc.lcAx, vc.lcAy));
        xfB.p = cB - rmul(xfB.s, xfB.c, mk(vc.lcBx, vc.lcBy));
        V2 normal, point; float separation;
        if (vc.type == 0) {
            normal = rmul(xfA.s, xfA.c, vc.localNormal);
            const V2 planePoint = xmul(xfA, vc.localPoint);
            const V2 clipPoint = xmul(xfB, vc.lp[j]);
            separation = dot(clipPoint - planePoint, normal) - B2_POLYGON_RADIUS - B2_POLYGON_RADIUS;
            point = clipPoint;
        } else {
            normal = rmul(xfB.s, xfB.c, vc.localNormal);
            const V2 planePoint = xmul(xfB, vc.localPoint);
            const V2 clipPoint = xmul(xfA, vc.lp[j]);
            separation = dot(clipPoint - planePoint, normal) - B2_POLYGON_RADIUS - B2_POLYGON_RADIUS;
            point = clipPoint;
            normal = neg(normal);
        }
        const V2 rA = point - cA, rB = point - cB;
        minSeparation = fminf(minSeparation, separation);
        const float C = clampf(0.2f * (separation + B2_LINEAR_SLOP), -0.2f, 0.0f);   // b2_baumgarte, b2_maxLinearCorrection
        const float rnA = cross(rA, normal), rnB = cross(rB, normal);
        const float K = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
        const float impulse = K > 0.0f ? -C / K : 0.0f;
        const V2 P = impulse * normal;
        cA = cA - mA * P; aA -= iA * cross(rA, P);
        cB = cB + mB * P; aB += iB * cross(rB, P);
    }
    A.cx = cA.x; A.cy = cA.y; A.a = aA; B.cx = cB.x; B.cy = cB.y; B.a = aB;
}

// Diagnostics build (-DMCR_PHASE_CLOCKS): lane 0 of every coupled warp adds the cycles of each part of a velocity
// iteration to g_coupled_clk (0 joints, 1 push + sync, 2 contacts + sync, 3 pull; 7 = iterations).
#ifdef MCR_PHASE_CLOCKS
__device__ unsigned long long g_coupled_clk[8];
#define CK_T0() long long ck_t_ = clock64()
#define CK(k) do { if (threadIdx.x == 0) { const long long n_ = clock64(); atomicAdd(&g_coupled_clk[k], (unsigned long long)(n_ - ck_t_)); ck_t_ = n_; if (k == 3) atomicAdd(&g_coupled_clk[7], 1ull); } } while (0)
extern "C" int mcr_debug_coupled_clocks(unsigned long long* out8, int reset) {
    if (out8 && cudaMemcpyFromSymbol(out8, g_coupled_clk, sizeof(g_coupled_clk)) != cudaSuccess) return -1;
    if (reset) { unsigned long long z[8] = {}; if (cudaMemcpyToSymbol(g_coupled_clk, z, sizeof(z)) != cudaSuccess) return -1; }
    return 0;
}
#else
#define CK_T0() do {} while (0)
#define CK(k) do {} while (0)
#endif

__global__ void __launch_bounds__(32)
coupled_kernel(Dims d, DevBuffers b, CarConst cc, const uint8_t* __restrict__ mask, int early_exit) {
    __shared__ BodyMirror s_bm[MCR_MAX_AGENTS * 5];
    __shared__ VC s_vc[MAXM];
    __shared__ int s_root[MCR_MAX_AGENTS];
    __shared__ float s_minsep[MCR_MAX_AGENTS];
    __shared__ int s_changed;
    __shared__ unsigned char s_list[MAXM];                 // manifold indices grouped by island root, ascending inside an island
    __shared__ int s_lstart[MCR_MAX_AGENTS + 1];
    tl_stamp(b.timeline, TL_COUPLED);
    const int env = blockIdx.x, lane = threadIdx.x;
    if (mask && !mask[env]) return;
    const int nman = b.n_manifold[env];
    if (nman == 0) return;
    const int A = d.A, N = d.N;
    const bool mine = lane < A;
    const int car = env * A + (mine ? lane : A - 1);
    const float h = (float)(1.0 / 50);
    float* gman = b.manifold + (size_t)env * MAXM * MW;

    // ---- load this lane's car (pre_kernel left the force-integrated velocities in scratch) ----------
    float cx[5], cy[5], ang[5], slp[5];
    bool awake[5];
    VelState s;
    float* sc = b.scratch + car;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const float* p = b.body + (size_t)(i * BODY_FIELDS) * N + car;
        cx[i] = p[(size_t)BF_CX * N]; cy[i] = p[(size_t)BF_CY * N]; ang[i] = p[(size_t)BF_A * N];
        s.vx[i] = sc[(size_t)(SC_VX + i) * N]; s.vy[i] = sc[(size_t)(SC_VY + i) * N]; s.w[i] = sc[(size_t)(SC_W + i) * N];
        slp[i] = b.sleep_time[(size_t)i * N + car];
        awake[i] = b.awake[(size_t)i * N + car] != 0;
        if (!awake[i]) { awake[i] = true; slp[i] = 0.0f; }          // the island DFS wakes every body
    }
    float motorSpeed[4];
    int lim[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        s.jix[k] = sc[(size_t)(SC_JIX + k) * N]; s.jiy[k] = sc[(size_t)(SC_JIY + k) * N];
        s.jiz[k] = sc[(size_t)(SC_JIZ + k) * N]; s.jmot[k] = sc[(size_t)(SC_JMOT + k) * N];
        motorSpeed[k] = sc[(size_t)(SC_JOINT + k * SC_JOINT_FIELDS + 14) * N];
        lim[k] = b.limit_state[(size_t)k * N + car];
    }
    // The contact solver only reads and writes the bodies some manifold refers to (`touched`, set below): inside the
    // iteration loops only those are exchanged through the shared-memory mirror (typically one body per car).
    unsigned touched = 0x1fu;
    auto push_vel = [&]() {
        if (mine) {
#pragma unroll
            for (int i = 0; i < 5; ++i) if ((touched >> i) & 1u) { BodyMirror& m = s_bm[lane * 5 + i]; m.vx = s.vx[i]; m.vy = s.vy[i]; m.w = s.w[i]; }
        }
    };
    auto pull_vel = [&]() {
#pragma unroll
        for (int i = 0; i < 5; ++i) if ((touched >> i) & 1u) { const BodyMirror& m = s_bm[(mine ? lane : A - 1) * 5 + i]; s.vx[i] = m.vx; s.vy[i] = m.vy; s.w[i] = m.w; }
    };
    auto push_pos = [&]() {
        if (mine) {
#pragma unroll
            for (int i = 0; i < 5; ++i) if ((touched >> i) & 1u) { BodyMirror& m = s_bm[lane * 5 + i]; m.cx = cx[i]; m.cy = cy[i]; m.a = ang[i]; }
        }
    };
    auto pull_pos = [&]() {
#pragma unroll
        for (int i = 0; i < 5; ++i) if ((touched >> i) & 1u) { const BodyMirror& m = s_bm[(mine ? lane : A - 1) * 5 + i]; cx[i] = m.cx; cy[i] = m.cy; ang[i] = m.a; }
    };

    // ---- islands: union-find over cars linked by manifolds, root = lowest car index -------------------
    if (lane == 0) {
        for (int c = 0; c < A; ++c) s_root[c] = c;
        for (int i = 0; i < nman; ++i) {
            const uint32_t key = __float_as_uint(gman[(size_t)i * MW]);
            int a = key & 0xff, bq = (key >> 16) & 0xff;
            while (s_root[a] != a) a = s_root[a];
            while (s_root[bq] != bq) bq = s_root[bq];
            if (a != bq) { if (a < bq) s_root[bq] = a; else s_root[a] = bq; }
        }
        for (int c = 0; c < A; ++c) { int r = c; while (s_root[r] != r) r = s_root[r]; s_root[c] = r; }
    }
    push_vel(); push_pos();
    __syncwarp();
    const int root = s_root[mine ? lane : A - 1];
    uint32_t imask = 0u;
    for (int k = 0; k < A; ++k) if (s_root[k] == root) imask |= 1u << k;

    // ---- b2ContactSolver: InitializeVelocityConstraints (lane per manifold), WarmStart (sequential) -----
    if (lane < nman) {
        Manifold m; manifold_load(gman + (size_t)lane * MW, m);
        contact_init(s_vc[lane], m, s_bm, cc);
        s_vc[lane].root = s_root[m.key & 0xff];
    }
    __syncwarp();
    // Islands do not share bodies, so their contact constraints are independent: every island's ROOT lane solves
    // its own manifolds (in manifold order, as Box2D does inside one island) while the other roots solve theirs.
    if (lane == 0) {
        int pos = 0;
        for (int r = 0; r < A; ++r) {
            s_lstart[r] = pos;
            for (int i = 0; i < nman; ++i) if (s_vc[i].root == r) s_list[pos++] = (unsigned char)i;
        }
        s_lstart[A] = pos;
    }
    __syncwarp();
    const int my_first = mine ? s_lstart[lane] : 0, my_count = mine ? s_lstart[lane + 1] - s_lstart[lane] : 0;
    for (int k = 0; k < my_count; ++k) contact_warm_start(s_vc[s_list[my_first + k]], s_bm);
    __syncwarp();
    {   // bodies of this lane's car that a manifold refers to (mirror slot = car * 5 + body)
        unsigned t = 0u;
        const int me = mine ? lane : A - 1;
        for (int i = 0; i < nman; ++i) {
            const int a = s_vc[i].bA, bq = s_vc[i].bB;
            if (a / 5 == me) t |= 1u << (a % 5);
            if (bq / 5 == me) t |= 1u << (bq % 5);
        }
        touched = t;
    }
    pull_vel();
    // ---- joints: InitVelocityConstraints + warm start, then the 180 sweeps in lock step ----------------
    JointC J[4];
    joints_init(cc, ang, motorSpeed, s.vx, s.vy, s.w, s.jix, s.jiy, s.jiz, s.jmot, lim, J);
    Masses m; m.mA = cc.hull_invMass; m.iA = cc.hull_invI; m.mB = cc.wheel_invMass; m.iB = cc.wheel_invI;
    m.maxMotorImpulse = h * cc.max_motor_torque;
    // joints_init fixed every joint's limit state for the step: when no car of the env has an active limit
    // (the common case) the whole warp runs the straight-line sweep<0> instead of the branching sweep<-1>
    // (696 -> 540 cycles per sweep, clock64 in situ).  The contact solve itself (~550 cycles for one two-point
    // manifold) is a dependent chain too: moving it from lane 0 / shared memory to a redundant solve in
    // registers fed by shuffles was measured and is not faster (809 cycles).
    const bool no_limits = __all_sync(0xffffffffu, !mine || (lim[0] | lim[1] | lim[2] | lim[3]) == LIM_INACTIVE);
    for (int it = 0; it < MCR_VEL_ITERS; it += 4) {
        const VelState before = s;
        float ci_before[2][4];                         // contact impulses of "my" manifold (lane < nman)
        if (lane < nman) { ci_before[0][0] = s_vc[lane].ni[0]; ci_before[0][1] = s_vc[lane].ni[1]; ci_before[0][2] = s_vc[lane].ti[0]; ci_before[0][3] = s_vc[lane].ti[1]; }
#pragma unroll 1
        for (int rep = 0; rep < 4; ++rep) {
            CK_T0();
            if (no_limits) sweep<0>(s, J, m); else sweep<-1>(s, J, m);
            CK(0);
            // (measured and dropped, profiles/README r02: fetching the manifold's two bodies from their owners' registers
            // with shuffles instead of this shared-memory mirror -- 12 shuffles per manifold -- was 17 % slower)
            push_vel();
            __syncwarp();
            CK(1);
            for (int k = 0; k < my_count; ++k) contact_solve_vel(s_vc[s_list[my_first + k]], s_bm);
            __syncwarp();
            CK(2);
            pull_vel();
            CK(3);
        }
        bool changed = mine && state_diff(before, s) != 0u;
        if (lane < nman) {
            changed = changed || __float_as_uint(ci_before[0][0]) != __float_as_uint(s_vc[lane].ni[0]) ||
                      __float_as_uint(ci_before[0][1]) != __float_as_uint(s_vc[lane].ni[1]) ||
                      __float_as_uint(ci_before[0][2]) != __float_as_uint(s_vc[lane].ti[0]) ||
                      __float_as_uint(ci_before[0][3]) != __float_as_uint(s_vc[lane].ti[1]);
        }
        if (early_exit && !__any_sync(0xffffffffu, changed)) break;
    }
    if (lane == 0) { atomicMax(b.timeline + TL_COUPLED_VEL_END, mcr_globaltimer()); }
    // StoreImpulses
    if (lane < nman) {
        float* w = gman + (size_t)lane * MW;
        for (int j = 0; j < s_vc[lane].pointCount; ++j) { w[6 + 4 * j] = s_vc[lane].ni[j]; w[7 + 4 * j] = s_vc[lane].ti[j]; }
    }
    // ---- integrate positions ------------------------------------------------------------------------------
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        float tx = h * s.vx[i], ty = h * s.vy[i];
        if (tx * tx + ty * ty > B2_MAX_TRANSLATION * B2_MAX_TRANSLATION) {
            float ratio = B2_MAX_TRANSLATION / sqrtf(tx * tx + ty * ty);
            s.vx[i] = ratio * s.vx[i]; s.vy[i] = ratio * s.vy[i];
        }
        float rotn = h * s.w[i];
        if (rotn * rotn > B2_MAX_ROTATION * B2_MAX_ROTATION) {
            float ratio = B2_MAX_ROTATION / fabsf(rotn);
            s.w[i] *= ratio;
        }
        cx[i] += h * s.vx[i]; cy[i] += h * s.vy[i];
        ang[i] += h * s.w[i];
    }
    // ---- position iterations, per island: contacts first, then joints; exit when both are satisfied ----------
    float motorMassK[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) motorMassK[k] = J[k].motorMass;
    bool active = mine, positionSolved = false;
    for (int it = 0; it < MCR_POS_ITERS; ++it) {
        const uint32_t act = __ballot_sync(0xffffffffu, active);
        if (act == 0u) break;
        float p_cx[5], p_cy[5], p_an[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) { p_cx[i] = cx[i]; p_cy[i] = cy[i]; p_an[i] = ang[i]; }
        push_pos();
        if (lane < A) s_minsep[lane] = 0.0f;
        __syncwarp();
        if (my_count > 0 && ((act >> lane) & 1u)) {       // the island's root lane is active iff the island is
            float ms = s_minsep[lane];
            for (int k = 0; k < my_count; ++k) contact_solve_pos(s_vc[s_list[my_first + k]], s_bm, ms);
            s_minsep[lane] = ms;
        }
        __syncwarp();
        bool ok = true, moved = false;
        if (active) {
            pull_pos();
            ok = joints_solve_pos(cc, cx, cy, ang, lim, motorMassK);
            unsigned mv = 0u;
#pragma unroll
            for (int i = 0; i < 5; ++i)
                mv |= (__float_as_uint(p_cx[i]) ^ __float_as_uint(cx[i])) | (__float_as_uint(p_cy[i]) ^ __float_as_uint(cy[i])) |
                      (__float_as_uint(p_an[i]) ^ __float_as_uint(ang[i]));
            moved = mv != 0u;
        }
        const uint32_t okb = __ballot_sync(0xffffffffu, ok || !active), mvb = __ballot_sync(0xffffffffu, moved);
        if (active) {
            const bool contactsOkay = s_minsep[root] >= -3.0f * B2_LINEAR_SLOP;
            const bool jointsOkay = (okb & imask) == imask;
            if (contactsOkay && jointsOkay) { positionSolved = true; active = false; }
            else if ((mvb & imask) == 0u) active = false;   // bit-identical iteration: a fixed point of the remaining ones
        }
        __syncwarp();
    }
    if (lane == 0) { atomicMax(b.timeline + TL_COUPLED_POS_END, mcr_globaltimer()); }
    // ---- sleep, island-wide ------------------------------------------------------------------------------------
    float laneMin = 3.402823466e+38f;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        if (s.w[i] * s.w[i] > B2_ANGULAR_SLEEP_TOL * B2_ANGULAR_SLEEP_TOL ||
            s.vx[i] * s.vx[i] + s.vy[i] * s.vy[i] > B2_LINEAR_SLEEP_TOL * B2_LINEAR_SLEEP_TOL) { slp[i] = 0.0f; laneMin = 0.0f; }
        else { slp[i] += h; laneMin = fminf(laneMin, slp[i]); }
    }
    float islandMin = 3.402823466e+38f;
    for (int k = 0; k < A; ++k) {
        const float v = __shfl_sync(0xffffffffu, laneMin, k);
        if ((imask >> k) & 1u) islandMin = fminf(islandMin, v);
    }
    if (islandMin >= B2_TIME_TO_SLEEP && positionSolved) {
#pragma unroll
        for (int i = 0; i < 5; ++i) { awake[i] = false; slp[i] = 0.0f; s.vx[i] = 0.0f; s.vy[i] = 0.0f; s.w[i] = 0.0f; }
    }
    // ---- hand over to post_kernel (it commits the sleep state after the tile-contact pass has finished) ----------
    if (mine) {
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            float* p = b.body + (size_t)(i * BODY_FIELDS) * N + car;
            p[(size_t)BF_CX * N] = cx[i]; p[(size_t)BF_CY * N] = cy[i]; p[(size_t)BF_A * N] = ang[i];
            sc[(size_t)(SC_VX + i) * N] = s.vx[i]; sc[(size_t)(SC_VY + i) * N] = s.vy[i]; sc[(size_t)(SC_W + i) * N] = s.w[i];
            sc[(size_t)(SC_SLP + i) * N] = slp[i]; sc[(size_t)(SC_AWAKE + i) * N] = awake[i] ? 1.0f : 0.0f;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            sc[(size_t)(SC_JIX + k) * N] = s.jix[k]; sc[(size_t)(SC_JIY + k) * N] = s.jiy[k];
            sc[(size_t)(SC_JIZ + k) * N] = s.jiz[k]; sc[(size_t)(SC_JMOT + k) * N] = s.jmot[k];
            b.limit_state[(size_t)k * N + car] = (uint8_t)lim[k];
        }
    }
}

int launch_carcontacts(const Dims& d, const DevBuffers& b, const CarConst& cc, const uint8_t* mask, void* stream) {
    carcontacts_kernel<<<(d.B + CC_WARPS - 1) / CC_WARPS, CC_WARPS * 32, 0, (cudaStream_t)stream>>>(d, b, cc, mask);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_head(const Dims& d, const DevBuffers& b, const CarConst& cc, const uint8_t* mask, const uint8_t* reset_flags,
                const AutoResetCfg& ar, const void* action, int action_dtype, int collisions, void* stream) {
    const int nb = (d.B + CC_WARPS - 1) / CC_WARPS;
    cudaStream_t s = (cudaStream_t)stream;
    const bool many = d.A > 3;
    if (action_dtype == MCR_F64) {
        if (many) head_kernel<double, true><<<nb, CC_WARPS * 32, 0, s>>>(d, b, cc, mask, reset_flags, ar, (const double*)action, collisions);
        else head_kernel<double, false><<<nb, CC_WARPS * 32, 0, s>>>(d, b, cc, mask, reset_flags, ar, (const double*)action, collisions);
    } else {
        if (many) head_kernel<float, true><<<nb, CC_WARPS * 32, 0, s>>>(d, b, cc, mask, reset_flags, ar, (const float*)action, collisions);
        else head_kernel<float, false><<<nb, CC_WARPS * 32, 0, s>>>(d, b, cc, mask, reset_flags, ar, (const float*)action, collisions);
    }
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// head_kernel's `action` argument for the captured step graph (api.cu patches it per step instead of copying the action)
int head_action_arg(const void* func, int* nargs) {
    const bool mine = func == (const void*)head_kernel<double, true> || func == (const void*)head_kernel<double, false> ||
                      func == (const void*)head_kernel<float, true> || func == (const void*)head_kernel<float, false>;
    if (nargs) *nargs = 8;
    return mine ? 6 : -1;
}

int launch_coupled(const Dims& d, const DevBuffers& b, const CarConst& cc, const uint8_t* mask, int early_exit, void* stream) {
    coupled_kernel<<<d.B, 32, 0, (cudaStream_t)stream>>>(d, b, cc, mask, early_exit);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}
