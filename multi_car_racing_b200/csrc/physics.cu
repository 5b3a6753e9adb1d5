// physics.cu -- kernel 1: Car.step (tyre model) + world.Step solve for one car per thread.
//
// Replaces, for every car of the batch (reference: gym_multi_car_racing/multi_car_racing.py):
//   mcr:421-424  car.steer(-a0) / car.gas(a1) / car.brake(a2)      (gym car_dynamics.Car)
//   mcr:426-427  car.step(1/FPS)            -- fp64 tyre model on fp32 body state
//   mcr:428      world.Step(1/FPS, 180, 60) -- b2Island::Solve specialised to the fixed
//                topology {hull + 4 wheels + 4 revolute joints (limit + motor)}; island joint
//                order [j3, j2, j1, j0]; sleeping; SynchronizeTransform.
// Numerics: IEEE fp32, no FMA contraction (-fmad=false), same operation order as Box2D;
// b2Rot::Set is (float)sin((double)a).  The wheel's local centre of mass and localAnchorB are
// exactly 0 so every rB term of b2RevoluteJoint is an exact +-0 and is dropped.
//
// Layout: one thread per car, SoA state => every load/store below is a coalesced 128-byte
// line per warp.  The 180 Gauss-Seidel iterations are a serial dependency chain through the
// hull velocity, so the kernel is latency bound: blocks are one warp wide to spread the batch
// over as many SMs as possible.
#include "mcr_internal.h"
#include <cuda_runtime.h>

#define PHYS_BLOCK 32

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
    float motorMass, motorSpeed;
    int limit;
};

__device__ __forceinline__ void solve22(const JointC& J, float bx, float by, float& ox, float& oy) {
    ox = J.det22 * (J.k22 * bx - J.k12 * by);
    oy = J.det22 * (J.k11 * by - J.k12 * bx);
}

// b2Mat33::Solve33 with ex=(k11,k12,ezx) ey=(k12,k22,ezy) ez=(ezx,ezy,ezz)
__device__ __forceinline__ void solve33(const JointC& J, float b0, float b1, float b2, float& x0, float& x1, float& x2) {
    const float ex0 = J.k11, ex1 = J.k12, ex2 = J.ezx;
    const float ey0 = J.k12, ey1 = J.k22, ey2 = J.ezy;
    const float ez0 = J.ezx, ez1 = J.ezy, ez2 = J.ezz;
    float cx = ey1 * ez2 - ey2 * ez1, cy = ey2 * ez0 - ey0 * ez2, cz = ey0 * ez1 - ey1 * ez0;
    float det = ex0 * cx + ex1 * cy + ex2 * cz;
    if (det != 0.0f) det = 1.0f / det;
    x0 = det * (b0 * cx + b1 * cy + b2 * cz);
    float dx = b1 * ez2 - b2 * ez1, dy = b2 * ez0 - b0 * ez2, dz = b0 * ez1 - b1 * ez0;
    x1 = det * (ex0 * dx + ex1 * dy + ex2 * dz);
    float fx = ey1 * b2 - ey2 * b1, fy = ey2 * b0 - ey0 * b2, fz = ey0 * b1 - ey1 * b0;
    x2 = det * (ex0 * fx + ex1 * fy + ex2 * fz);
}

template <typename ActT>
__global__ void __launch_bounds__(PHYS_BLOCK)
physics_kernel(Dims d, DevBuffers b, CarConst cc, const uint8_t* __restrict__ mask, const ActT* __restrict__ action, double h_ratio) {
    const int car = blockIdx.x * PHYS_BLOCK + threadIdx.x;
    if (car >= d.N) return;
    const int env = car / d.A;
    if (mask && !mask[env]) return;
    const int N = d.N;

    // ---- load state ----------------------------------------------------------------
    float cx[5], cy[5], ang[5], vx[5], vy[5], w[5], qs[5], qc[5], slp[5];
    bool awake[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const float* p = b.body + (size_t)(i * BODY_FIELDS) * N + car;
        cx[i] = p[(size_t)BF_CX * N]; cy[i] = p[(size_t)BF_CY * N]; ang[i] = p[(size_t)BF_A * N];
        vx[i] = p[(size_t)BF_VX * N]; vy[i] = p[(size_t)BF_VY * N]; w[i] = p[(size_t)BF_W * N];
        qs[i] = p[(size_t)BF_QS * N]; qc[i] = p[(size_t)BF_QC * N];
        slp[i] = b.sleep_time[(size_t)i * N + car];
        awake[i] = b.awake[(size_t)i * N + car] != 0;
    }
    float jix[4], jiy[4], jiz[4], jmot[4];
    int lim[4];
    double omega[4], phase[4];
    bool on_road[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float* p = b.joint + (size_t)(k * JOINT_FIELDS) * N + car;
        jix[k] = p[(size_t)JF_IX * N]; jiy[k] = p[(size_t)JF_IY * N]; jiz[k] = p[(size_t)JF_IZ * N];
        jmot[k] = p[(size_t)JF_MOTOR * N];
        lim[k] = b.limit_state[(size_t)k * N + car];
        omega[k] = b.wheel[(size_t)(k * WHEEL_FIELDS + WF_OMEGA) * N + car];
        phase[k] = b.wheel[(size_t)(k * WHEEL_FIELDS + WF_PHASE) * N + car];
        on_road[k] = b.on_road[(size_t)k * N + car] != 0;
    }
    double gas = b.ctrl[(size_t)CF_GAS * N + car];
    double brake = b.ctrl[(size_t)CF_BRAKE * N + car];
    double steer = b.ctrl[(size_t)CF_STEER * N + car];

    // ---- controls, mcr:421-424 -------------------------------------------------------
    if (action) {
        double a0 = (double)action[(size_t)car * 3 + 0];
        double a1 = (double)action[(size_t)car * 3 + 1];
        double a2 = (double)action[(size_t)car * 3 + 2];
        steer = -a0;
        double g = a1 < 0 ? 0 : (a1 > 1 ? 1 : a1);
        double diff = g - gas;
        if (diff > 0.1) diff = 0.1;
        gas += diff;
        brake = a2;
    }

    // ---- Car.step(dt): tyre model (float64), per wheel ------------------------------------
    const double SIZE = 0.02;
    const double ENGINE_POWER = 100000000 * SIZE * SIZE;
    const double WHEEL_MOI = 4000 * SIZE * SIZE;
    const double FRICTION_LIMIT = 1000000 * SIZE * SIZE;
    const double dt = 1.0 / 50;
    float motorSpeed[4], Fx[4], Fy[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int bi = 1 + k;
        const double steer_w = k < 2 ? steer : 0.0;
        const double gas_w = k >= 2 ? gas : 0.0;
        double jangle = (double)(ang[bi] - ang[0]);
        double dir = sign_d(steer_w - jangle);
        double val = fabs(steer_w - jangle);
        motorSpeed[k] = (float)(dir * fmin(50.0 * val, 3.0));
        double friction_limit = FRICTION_LIMIT * 0.6;
        if (on_road[k]) friction_limit = fmax(friction_limit, FRICTION_LIMIT * 1.0);
        // GetWorldVector((0,1)) = (-s, c); ((1,0)) = (c, s)   (fp32 products with exact 0/1)
        float forw_x = qc[bi] * 0.0f - qs[bi] * 1.0f, forw_y = qs[bi] * 0.0f + qc[bi] * 1.0f;
        float side_x = qc[bi] * 1.0f - qs[bi] * 0.0f, side_y = qs[bi] * 1.0f + qc[bi] * 0.0f;
        double wvx = vx[bi], wvy = vy[bi];
        double vf = (double)forw_x * wvx + (double)forw_y * wvy;
        double vs = (double)side_x * wvx + (double)side_y * wvy;
        omega[k] += dt * ENGINE_POWER * gas_w / WHEEL_MOI / (fabs(omega[k]) + 5.0);
        if (brake >= 0.9) {
            omega[k] = 0;
        } else if (brake > 0) {
            double bdir = -sign_d(omega[k]);
            double bval = 15 * brake;
            if (fabs(bval) > fabs(omega[k])) bval = fabs(omega[k]);
            omega[k] += bdir * bval;
        }
        phase[k] += omega[k] * dt;
        const double wheel_rad = 1.0 * 27 * SIZE;
        double vr = omega[k] * wheel_rad;
        double f_force = -vf + vr;
        double p_force = -vs;
        f_force *= 205000 * SIZE * SIZE;
        p_force *= 205000 * SIZE * SIZE;
        double force = sqrt(f_force * f_force + p_force * p_force);
        if (fabs(force) > friction_limit) {
            f_force /= force; p_force /= force;
            force = friction_limit;
            f_force *= force; p_force *= force;
        }
        omega[k] -= dt * f_force * wheel_rad / WHEEL_MOI;
        Fx[k] = (float)(p_force * (double)side_x + f_force * (double)forw_x);
        Fy[k] = (float)(p_force * (double)side_y + f_force * (double)forw_y);
        if (!awake[bi]) { awake[bi] = true; slp[bi] = 0.0f; }   // ApplyForceToCenter(wake=True)
    }

    // ---- b2Island::Solve ---------------------------------------------------------------
    const float h = (float)(1.0 / 50);
    if (!awake[0]) { awake[0] = true; slp[0] = 0.0f; }           // island DFS wakes the hull
    const float mA = cc.hull_invMass, iA = cc.hull_invI, mB = cc.wheel_invMass, iB = cc.wheel_invI;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        vx[1 + k] += h * (mB * Fx[k]);
        vy[1 + k] += h * (mB * Fy[k]);
    }
    // InitVelocityConstraints, joints in island order 3,2,1,0
    JointC J[4];
    {
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
            float jointAngle = ang[bi] - ang[0] - 0.0f;
            if (fabsf(cc.upper - cc.lower) < 2.0f * B2_ANGULAR_SLOP) {
                lim[k] = LIM_EQUAL;
            } else if (jointAngle <= cc.lower) {
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
    // SolveVelocityConstraints x 180 (no early exit in Box2D)
    const float maxMotorImpulse = h * cc.max_motor_torque;
    for (int it = 0; it < MCR_VEL_ITERS; ++it) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const int k = 3 - kk;
            const int bi = 1 + k;
            const JointC& j = J[k];
            float vAx = vx[0], vAy = vy[0], wA = w[0], vBx = vx[bi], vBy = vy[bi], wB = w[bi];
            if (j.limit != LIM_EQUAL) {
                float Cdot = wB - wA - j.motorSpeed;
                float impulse = -j.motorMass * Cdot;
                float oldImpulse = jmot[k];
                jmot[k] = clampf(jmot[k] + impulse, -maxMotorImpulse, maxMotorImpulse);
                impulse = jmot[k] - oldImpulse;
                wA -= iA * impulse;
                wB += iB * impulse;
            }
            // Cdot1 = vB + cross(wB, rB) - vA - cross(wA, rA),  cross(s, r) = (-s*r.y, s*r.x)
            float C1x = vBx - vAx - (-wA * j.rAy);
            float C1y = vBy - vAy - (wA * j.rAx);
            if (j.limit != LIM_INACTIVE) {
                float Cdot2 = wB - wA;
                float i0, i1, i2;
                solve33(j, C1x, C1y, Cdot2, i0, i1, i2);
                i0 = -i0; i1 = -i1; i2 = -i2;
                if (j.limit == LIM_EQUAL) {
                    jix[k] += i0; jiy[k] += i1; jiz[k] += i2;
                } else {
                    float newImpulse = jiz[k] + i2;
                    bool release = (j.limit == LIM_LOWER) ? (newImpulse < 0.0f) : (newImpulse > 0.0f);
                    if (release) {
                        float rx = -C1x + jiz[k] * j.ezx, ry = -C1y + jiz[k] * j.ezy;
                        float redx, redy; solve22(j, rx, ry, redx, redy);
                        i0 = redx; i1 = redy; i2 = -jiz[k];
                        jix[k] += redx; jiy[k] += redy; jiz[k] = 0.0f;
                    } else {
                        jix[k] += i0; jiy[k] += i1; jiz[k] += i2;
                    }
                }
                vAx -= mA * i0; vAy -= mA * i1;
                wA -= iA * ((j.rAx * i1 - j.rAy * i0) + i2);
                vBx += mB * i0; vBy += mB * i1;
                wB += iB * i2;
            } else {
                float ix, iy; solve22(j, -C1x, -C1y, ix, iy);
                jix[k] += ix; jiy[k] += iy;
                vAx -= mA * ix; vAy -= mA * iy;
                wA -= iA * (j.rAx * iy - j.rAy * ix);
                vBx += mB * ix; vBy += mB * iy;
            }
            vx[0] = vAx; vy[0] = vAy; w[0] = wA; vx[bi] = vBx; vy[bi] = vBy; w[bi] = wB;
        }
    }
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
    for (int it = 0; it < MCR_POS_ITERS; ++it) {
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
                    limitImpulse = -J[k].motorMass * C;
                    angularError = fabsf(C);
                } else if (lim[k] == LIM_LOWER) {
                    float C = angle - cc.lower;
                    angularError = -C;
                    C = clampf(C + B2_ANGULAR_SLOP, -B2_MAX_ANGULAR_CORRECTION, 0.0f);
                    limitImpulse = -J[k].motorMass * C;
                } else {
                    float C = angle - cc.upper;
                    angularError = C;
                    C = clampf(C - B2_ANGULAR_SLOP, 0.0f, B2_MAX_ANGULAR_CORRECTION);
                    limitImpulse = -J[k].motorMass * C;
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
        if (jointsOkay) { positionSolved = true; break; }
    }
    // SynchronizeTransform + sleep
    float px[5], py[5];
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
        b.limit_state[(size_t)k * N + car] = (uint8_t)lim[k];
        b.wheel[(size_t)(k * WHEEL_FIELDS + WF_OMEGA) * N + car] = omega[k];
        b.wheel[(size_t)(k * WHEEL_FIELDS + WF_PHASE) * N + car] = phase[k];
        // the contact pass of THIS step (already run) decides the friction of the NEXT Car.step
        b.on_road[(size_t)k * N + car] = b.on_road_next[(size_t)k * N + car];
    }
    b.ctrl[(size_t)CF_GAS * N + car] = gas;
    b.ctrl[(size_t)CF_BRAKE * N + car] = brake;
    b.ctrl[(size_t)CF_STEER * N + car] = steer;
    // ---- per-view values the rasteriser needs, evaluated once here (fp64 trig is serial latency) --
    const double t = b.time[car] + 1.0 / 50;                      // mcr:429
    b.time[car] = t;
    if (action) b.steps[car] += 1;                                // TimeLimit counts step() calls only
    {   // camera, mcr:540-556 + Transform.enable + glViewport(0,0,96,96) under glOrtho(0,1000,0,800)
        const double SCALE = 6.0, ZOOM = 2.7, WINDOW_W = 1000, WINDOW_H = 800;
        const double zoom = 0.1 * SCALE * fmax(1 - t, 0.0) + ZOOM * SCALE * fmin(t, 1.0);
        const double scroll_x = px[0], scroll_y = py[0];
        double angle = -(double)ang[0];
        const double hvx = vx[0], hvy = vy[0];
        const bool fast = sqrt(hvx * hvx + hvy * hvy) > 0.5;
        double at = 0.0;
        if (fast) { at = atan2(hvx, hvy); angle = at; }
        const double tx = WINDOW_W / 2 - (scroll_x * zoom * cos(angle) - scroll_y * zoom * sin(angle));
        const double ty = WINDOW_H * h_ratio - (scroll_x * zoom * sin(angle) + scroll_y * zoom * cos(angle));
        const float ftx = (float)tx, fty = (float)ty, fdeg = (float)(57.29577951308232 * angle), fzoom = (float)zoom;
        const double rad = (double)fdeg * (3.14159265358979323846 / 180.0);
        const double cs = cos(rad), sn = sin(rad);
        const double SX = 96.0 / 1000.0, SY = 96.0 / 800.0;
        b.camera[(size_t)0 * N + car] = (float)(cs * (double)fzoom * SX);
        b.camera[(size_t)1 * N + car] = (float)(-sn * (double)fzoom * SX);
        b.camera[(size_t)2 * N + car] = (float)((double)ftx * SX);
        b.camera[(size_t)3 * N + car] = (float)(sn * (double)fzoom * SY);
        b.camera[(size_t)4 * N + car] = (float)(cs * (double)fzoom * SY);
        b.camera[(size_t)5 * N + car] = (float)((double)fty * SY);
        // car_angle of the backward test, mcr:449-456
        const double PI = 3.141592653589793;
        double car_angle = fast ? -at : (double)ang[0];
        car_angle = fmod(car_angle + 2 * PI, 2 * PI);
        if (car_angle != 0 && car_angle < 0) car_angle += 2 * PI;
        b.heading[car] = car_angle;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {   // Car.draw wheel stripe, evaluated once per wheel for all A views
        const double a1 = phase[k], a2 = phase[k] + 1.2;
        double s1 = sin(a1), s2 = sin(a2), c1 = cos(a1), c2 = cos(a2);
        float y1 = __int_as_float(0x7fc00000), y2 = 0.0f;
        if (!(s1 > 0 && s2 > 0)) {
            if (s1 > 0) c1 = sign_d(c1);
            if (s2 > 0) c2 = sign_d(c2);
            y1 = (float)(+27 * c1 * SIZE); y2 = (float)(+27 * c2 * SIZE);
        }
        b.stripe[(size_t)(k * 2 + 0) * N + car] = y1;
        b.stripe[(size_t)(k * 2 + 1) * N + car] = y2;
    }
}

int launch_physics(const Dims& d, const DevBuffers& b, const CarConst& cc, const uint8_t* mask,
                   const void* action, int action_dtype, double h_ratio, void* stream) {
    dim3 grid((d.N + PHYS_BLOCK - 1) / PHYS_BLOCK), block(PHYS_BLOCK);
    cudaStream_t s = (cudaStream_t)stream;
    if (action_dtype == MCR_F64)
        physics_kernel<double><<<grid, block, 0, s>>>(d, b, cc, mask, (const double*)action, h_ratio);
    else
        physics_kernel<float><<<grid, block, 0, s>>>(d, b, cc, mask, (const float*)action, h_ratio);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}
