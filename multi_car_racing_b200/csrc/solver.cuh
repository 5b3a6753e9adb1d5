// solver.cuh -- device code shared by the per-car solver kernels (sim.cu) and the coupled-island
// kernel (carcontacts.cu): b2RevoluteJoint (limit + motor) restated for the fixed car topology.
// Every function evaluates the same IEEE fp32 operation sequence as Box2D 2.3.x (no contraction).
#pragma once
#include "mcr_internal.h"
#include <cuda_runtime.h>

__device__ __forceinline__ void rot_set(float a, float& s, float& c) {
    double ds, dc;
    sincos((double)a, &ds, &dc);
    s = (float)ds; c = (float)dc;
}
__device__ __forceinline__ float clampf(float a, float lo, float hi) { return fmaxf(lo, fminf(a, hi)); }
__device__ __forceinline__ double sign_d(double x) { return x > 0 ? 1.0 : (x < 0 ? -1.0 : 0.0); }

struct JointC {            // per-step constants of one revolute joint
    float rAx, rAy;
    float k11, k12, k22;   // K.ex.x, K.ey.x (= K.ex.y), K.ey.y
    float ezx, ezy, ezz;   // K.ez
    float det22;           // 1/det of the 2x2 block (0 if singular)
    float cfx, cfy, cfz;   // cross(K.ey, K.ez): row-independent part of b2Mat33::Solve33
    float det33;           // 1/det of K (0 if singular)
    float motorMass, motorSpeed;
    int limit;
};

__device__ __forceinline__ void solve22(const JointC& J, float bx, float by, float& ox, float& oy) {
    ox = J.det22 * (J.k22 * bx - J.k12 * by);
    oy = J.det22 * (J.k11 * by - J.k12 * bx);
}

// b2Mat33::Solve33 with ex=(k11,k12,ezx) ey=(k12,k22,ezy) ez=(ezx,ezy,ezz); cross(ey,ez) and the
// determinant do not depend on the right-hand side and are evaluated once per step (same values).
__device__ __forceinline__ void solve33_init(JointC& J) {
    const float ex0 = J.k11, ex1 = J.k12, ex2 = J.ezx;
    const float ey0 = J.k12, ey1 = J.k22, ey2 = J.ezy;
    const float ez0 = J.ezx, ez1 = J.ezy, ez2 = J.ezz;
    J.cfx = ey1 * ez2 - ey2 * ez1; J.cfy = ey2 * ez0 - ey0 * ez2; J.cfz = ey0 * ez1 - ey1 * ez0;
    float det = ex0 * J.cfx + ex1 * J.cfy + ex2 * J.cfz;
    if (det != 0.0f) det = 1.0f / det;
    J.det33 = det;
}
__device__ __forceinline__ void solve33(const JointC& J, float b0, float b1, float b2, float& x0, float& x1, float& x2) {
    const float ex0 = J.k11, ex1 = J.k12, ex2 = J.ezx;
    const float ey0 = J.k12, ey1 = J.k22, ey2 = J.ezy;
    const float ez0 = J.ezx, ez1 = J.ezy, ez2 = J.ezz;
    const float det = J.det33;
    x0 = det * (b0 * J.cfx + b1 * J.cfy + b2 * J.cfz);
    float dx = b1 * ez2 - b2 * ez1, dy = b2 * ez0 - b0 * ez2, dz = b0 * ez1 - b1 * ez0;
    x1 = det * (ex0 * dx + ex1 * dy + ex2 * dz);
    float fx = ey1 * b2 - ey2 * b1, fy = ey2 * b0 - ey0 * b2, fz = ey0 * b1 - ey1 * b0;
    x2 = det * (ex0 * fx + ex1 * fy + ex2 * fz);
}

struct Masses { float mA, iA, mB, iB, maxMotorImpulse; };

// b2RevoluteJoint::SolveVelocityConstraints for one joint.  LIMIT_ACTIVE is the joint's limit
// state (at lower / at upper), fixed for the whole step by InitVelocityConstraints, so the 180
// sweeps run straight-line code; the "release" case of the limit complementarity is a select.
// (e_equalLimits cannot occur: upper - lower = 0.8 rad, checked at mcr_create.)
template <bool LIMIT_ACTIVE>
__device__ __forceinline__ void joint_sweep(const JointC& j, const Masses& m, float& vAx, float& vAy, float& wA,
                                            float& vBx, float& vBy, float& wB, float& jix, float& jiy, float& jiz,
                                            float& jmot) {
    {   // motor
        float Cdot = wB - wA - j.motorSpeed;
        float impulse = -j.motorMass * Cdot;
        float oldImpulse = jmot;
        jmot = clampf(jmot + impulse, -m.maxMotorImpulse, m.maxMotorImpulse);
        impulse = jmot - oldImpulse;
        wA -= m.iA * impulse;
        wB += m.iB * impulse;
    }
    // Cdot1 = vB + cross(wB, rB) - vA - cross(wA, rA),  cross(s, r) = (-s*r.y, s*r.x),  rB = 0
    const float C1x = vBx - vAx - (-wA * j.rAy);
    const float C1y = vBy - vAy - (wA * j.rAx);
    if (LIMIT_ACTIVE) {
        const float Cdot2 = wB - wA;
        float i0, i1, i2;
        solve33(j, C1x, C1y, Cdot2, i0, i1, i2);
        i0 = -i0; i1 = -i1; i2 = -i2;
        const float newImpulse = jiz + i2;
        const bool release = (j.limit == LIM_LOWER) ? (newImpulse < 0.0f) : (newImpulse > 0.0f);
        const float rx = -C1x + jiz * j.ezx, ry = -C1y + jiz * j.ezy;
        float redx, redy; solve22(j, rx, ry, redx, redy);
        i0 = release ? redx : i0; i1 = release ? redy : i1; i2 = release ? -jiz : i2;
        jix += i0; jiy += i1; jiz = release ? 0.0f : newImpulse;
        vAx -= m.mA * i0; vAy -= m.mA * i1;
        wA -= m.iA * ((j.rAx * i1 - j.rAy * i0) + i2);
        vBx += m.mB * i0; vBy += m.mB * i1;
        wB += m.iB * i2;
    } else {
        float ix, iy; solve22(j, -C1x, -C1y, ix, iy);
        jix += ix; jiy += iy;
        vAx -= m.mA * ix; vAy -= m.mA * iy;
        wA -= m.iA * (j.rAx * iy - j.rAy * ix);
        vBx += m.mB * ix; vBy += m.mB * iy;
    }
}

struct VelState { float vx[5], vy[5], w[5], jix[4], jiy[4], jiz[4], jmot[4]; };

// One Gauss-Seidel sweep over the island's joints in Box2D's order [j3, j2, j1, j0].
// PAT bit k = joint k has an active limit.  PAT < 0: decide per joint at run time.
template <int PAT>
__device__ __forceinline__ void sweep(VelState& s, const JointC (&J)[4], const Masses& m) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        const int k = 3 - kk, bi = 1 + k;
        const bool act = PAT >= 0 ? ((PAT >> k) & 1) != 0 : (J[k].limit != LIM_INACTIVE);
        if (act) joint_sweep<true>(J[k], m, s.vx[0], s.vy[0], s.w[0], s.vx[bi], s.vy[bi], s.w[bi], s.jix[k], s.jiy[k], s.jiz[k], s.jmot[k]);
        else joint_sweep<false>(J[k], m, s.vx[0], s.vy[0], s.w[0], s.vx[bi], s.vy[bi], s.w[bi], s.jix[k], s.jiy[k], s.jiz[k], s.jmot[k]);
    }
}

__device__ __forceinline__ unsigned state_diff(const VelState& a, const VelState& b) {
    unsigned d0 = 0u, d1 = 0u, d2 = 0u, d3 = 0u;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        d0 |= __float_as_uint(a.vx[i]) ^ __float_as_uint(b.vx[i]);
        d1 |= __float_as_uint(a.vy[i]) ^ __float_as_uint(b.vy[i]);
        d2 |= __float_as_uint(a.w[i]) ^ __float_as_uint(b.w[i]);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        d0 |= __float_as_uint(a.jix[k]) ^ __float_as_uint(b.jix[k]);
        d1 |= __float_as_uint(a.jiy[k]) ^ __float_as_uint(b.jiy[k]);
        d2 |= __float_as_uint(a.jiz[k]) ^ __float_as_uint(b.jiz[k]);
        d3 |= __float_as_uint(a.jmot[k]) ^ __float_as_uint(b.jmot[k]);
    }
    return (d0 | d1) | (d2 | d3);
}

// Box2D runs all 180 sweeps.  A sweep is a deterministic map of (velocities, accumulated
// impulses): when SWEEP_CHECK sweeps leave that state bit-identical (a fixed point or a cycle whose
// length divides SWEEP_CHECK, checked at multiples of it so the phase matches sweep 180) the remaining sweeps
// cannot change it, so stopping there is exact.
template <int PAT>
__device__ __forceinline__ void solve_velocity(VelState& s, const JointC (&J)[4], const Masses& m, bool early_exit,
                                               unsigned peers = 0u) {
    // one sweep per loop trip keeps the loop body (~3 KB of SASS) inside the L0 instruction cache
    // SWEEP_CHECK sweeps between two comparisons (a divisor of MCR_VEL_ITERS: a state that repeats after SWEEP_CHECK sweeps
    // repeats after every multiple of it, so the phase matches sweep 180)
// (A/B builds: python -m multi_car_racing_b200.build --sweep-check N.  4 -> 12: the state copy + 62 compare instructions +
// vote per check were 9 % of the chain: sweep 45.9 -> 41.8 us, step 127.0 -> 124.9 us; 20 measured equal to 12.)
#ifndef SWEEP_CHECK
#define SWEEP_CHECK 12
#endif
    static_assert(MCR_VEL_ITERS % SWEEP_CHECK == 0, "SWEEP_CHECK must divide the sweep count");
#pragma unroll 1
    for (int it = 0; it < MCR_VEL_ITERS; it += SWEEP_CHECK) {
        const VelState before = s;
#ifndef SWEEP_UNROLL
#define SWEEP_UNROLL 1
#endif
        constexpr int kSweepUnroll = SWEEP_UNROLL;     // (A/B builds: --sweep-unroll N)
#pragma unroll kSweepUnroll
        for (int r = 0; r < SWEEP_CHECK; ++r) sweep<PAT>(s, J, m);
        // peers != 0: the lanes of `peers` hold one car each and leave together
        const bool settled = state_diff(before, s) == 0u;
        if (early_exit && (peers ? __all_sync(peers, settled) : settled)) break;
    }
}

// scratch SoA between pre / sweep / post: scratch[f * N + car]
enum { SC_VX = 0, SC_VY = 5, SC_W = 10, SC_JIX = 15, SC_JIY = 19, SC_JIZ = 23, SC_JMOT = 27, SC_JOINT = 31, SC_JOINT_FIELDS = 15,
       SC_SLP = SC_JOINT + 4 * SC_JOINT_FIELDS, SC_AWAKE = SC_SLP + 5, SC_FIELDS = SC_AWAKE + 5 };
static_assert(SC_FIELDS == MCR_SCRATCH_FIELDS, "scratch layout");


// b2RevoluteJoint::InitVelocityConstraints (incl. warm start) for the four joints of one car, in the
// island's joint order [j3, j2, j1, j0].
__device__ __forceinline__ void joints_init(const CarConst& cc, const float (&ang)[5], const float (&motorSpeed)[4],
                                            float (&vx)[5], float (&vy)[5], float (&w)[5], const float (&jix)[4],
                                            const float (&jiy)[4], float (&jiz)[4], const float (&jmot)[4], int (&lim)[4],
                                            JointC (&J)[4]) {
    const float mA = cc.hull_invMass, iA = cc.hull_invI, mB = cc.wheel_invMass, iB = cc.wheel_invI;
        float sA, cA; rot_set(ang[0], sA, cA);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const int k = 3 - kk;
            const int bi = 1 + k;
            float lx = cc.anchor_x[k] - cc.hull_lcx, ly = cc.anchor_y[k] - cc.hull_lcy;
            float rAx = cA * lx - sA * ly, rAy = sA * lx + cA * ly;
            JointC& j = J[k];
            j.rAx = rAx; j.rAy = rAy;
            j.k11 = mA + mB + rAy * rAy * iA;
            j.k12 = -rAy * rAx * iA;
            j.ezx = -rAy * iA;
            j.k22 = mA + mB + rAx * rAx * iA;
            j.ezy = rAx * iA;
            j.ezz = iA + iB;
            float det = j.k11 * j.k22 - j.k12 * j.k12;
            if (det != 0.0f) det = 1.0f / det;
            j.det22 = det;
            float mm = iA + iB;
            if (mm > 0.0f) mm = 1.0f / mm;
            j.motorMass = mm;
            j.motorSpeed = motorSpeed[k];
            solve33_init(j);
            float jointAngle = ang[bi] - ang[0] - 0.0f;
            if (jointAngle <= cc.lower) {              // e_equalLimits is excluded at mcr_create
                if (lim[k] != LIM_LOWER) jiz[k] = 0.0f;
                lim[k] = LIM_LOWER;
            } else if (jointAngle >= cc.upper) {
                if (lim[k] != LIM_UPPER) jiz[k] = 0.0f;
                lim[k] = LIM_UPPER;
            } else {
                lim[k] = LIM_INACTIVE;
                jiz[k] = 0.0f;
            }
            j.limit = lim[k];
            // warm start (dtRatio == 1 exactly: 50.0f * 0.02f rounds to 1.0f; impulses are 0 on the first step)
            float Px = jix[k], Py = jiy[k];
            vx[0] -= mA * Px; vy[0] -= mA * Py;
            w[0] -= iA * ((rAx * Py - rAy * Px) + jmot[k] + jiz[k]);
            vx[bi] += mB * Px; vy[bi] += mB * Py;
            w[bi] += iB * (jmot[k] + jiz[k]);
        }
}

// b2RevoluteJoint::SolvePositionConstraints for the four joints of one car (one position iteration).
__device__ __forceinline__ bool joints_solve_pos(const CarConst& cc, float (&cx)[5], float (&cy)[5], float (&ang)[5],
                                                 const int (&lim)[4], const float (&motorMassK)[4]) {
    const float mA = cc.hull_invMass, iA = cc.hull_invI, mB = cc.wheel_invMass, iB = cc.wheel_invI;
    bool jointsOkay = true;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        const int k = 3 - kk;
        const int bi = 1 + k;
        float aA = ang[0], aB = ang[bi];
        float angularError = 0.0f;
        if (lim[k] != LIM_INACTIVE) {
            float angle = aB - aA - 0.0f;
            float limitImpulse = 0.0f;
            if (lim[k] == LIM_EQUAL) {
                float C = clampf(angle - cc.lower, -B2_MAX_ANGULAR_CORRECTION, B2_MAX_ANGULAR_CORRECTION);
                limitImpulse = -motorMassK[k] * C;
                angularError = fabsf(C);
            } else if (lim[k] == LIM_LOWER) {
                float C = angle - cc.lower;
                angularError = -C;
                C = clampf(C + B2_ANGULAR_SLOP, -B2_MAX_ANGULAR_CORRECTION, 0.0f);
                limitImpulse = -motorMassK[k] * C;
            } else {
                float C = angle - cc.upper;
                angularError = C;
                C = clampf(C - B2_ANGULAR_SLOP, 0.0f, B2_MAX_ANGULAR_CORRECTION);
                limitImpulse = -motorMassK[k] * C;
            }
            aA -= iA * limitImpulse;
            aB += iB * limitImpulse;
        }
        float sA, cA; rot_set(aA, sA, cA);
        float lx = cc.anchor_x[k] - cc.hull_lcx, ly = cc.anchor_y[k] - cc.hull_lcy;
        float rAx = cA * lx - sA * ly, rAy = sA * lx + cA * ly;
        float Cx = cx[bi] - cx[0] - rAx, Cy = cy[bi] - cy[0] - rAy;
        float positionError = sqrtf(Cx * Cx + Cy * Cy);
        float K11 = mA + mB + iA * rAy * rAy;
        float K12 = -iA * rAx * rAy;
        float K22 = mA + mB + iA * rAx * rAx;
        float det = K11 * K22 - K12 * K12;
        if (det != 0.0f) det = 1.0f / det;
        float ix = -(det * (K22 * Cx - K12 * Cy));
        float iy = -(det * (K11 * Cy - K12 * Cx));
        cx[0] -= mA * ix; cy[0] -= mA * iy;
        aA -= iA * (rAx * iy - rAy * ix);
        cx[bi] += mB * ix; cy[bi] += mB * iy;
        ang[0] = aA; ang[bi] = aB;
        bool ok = positionError <= B2_LINEAR_SLOP && angularError <= B2_ANGULAR_SLOP;
        jointsOkay = jointsOkay && ok;
    }
    return jointsOkay;
}

