// pre.cuh -- the per-car head of a step: controls (mcr:421-424), Car.step's float64 tyre model
// (mcr:426-427, gym car_dynamics) and the start of b2Island::Solve (integrate forces,
// InitVelocityConstraints + warm start), handed to the sweep kernels through `scratch`.
// One THREAD per car.  Shared by pre_kernel (sim.cu) and the fused head_kernel (carcontacts.cu).
#pragma once
#include "solver.cuh"

// Diagnostics build: the including file may define PRE_CLK(k) (k = 0 loads, 1 controls + tyre model, 2 integrate + joints_init, 3 stores)
#ifndef PRE_CLK
#define PRE_CLK_T0() do {} while (0)
#define PRE_CLK(k) do {} while (0)
#endif

// Skid trace of gym car_dynamics Car.step for the four wheels of one car (only with mcr_config.particles;
// kept out of line so that the step's hot path does not carry its code): skid_bits / grass_bits = per wheel
// |force| > 2 friction_limit / no tile under the wheel; (wx, wy) = wheel.position (the wheel body's origin).
static __device__ __noinline__ void skid_traces(const Dims& d, const DevBuffers& b, int car, unsigned skid_bits, unsigned grass_bits,
                                         float wx0, float wy0, float wx1, float wy1, float wx2, float wy2, float wx3, float wy3) {
    const int N = d.N;
    for (int k = 0; k < 4; ++k) {
        const float wx = k == 0 ? wx0 : (k == 1 ? wx1 : (k == 2 ? wx2 : wx3)), wy = k == 0 ? wy0 : (k == 1 ? wy1 : (k == 2 ? wy2 : wy3));
        const bool grass = (grass_bits >> k) & 1u;
        int meta = b.skid_meta[(size_t)k * N + car];
        if ((skid_bits >> k) & 1u) {
            const bool ref_valid = (meta & 2) != 0, ref_grass = (meta & 4) != 0;
            const int ref_len = (meta >> 8) & 0xff, ref_slot = ((meta >> 16) & 0xff) - 1;
            if (ref_valid && ref_grass == grass && ref_len < PRT_PTS) {
                if (ref_slot >= 0) {             // still in Car.particles: the drawn polyline grows
                    float* pt = b.prt_pts + ((size_t)(ref_slot * PRT_PTS + ref_len) * 2) * N + car;
                    pt[0] = wx; pt[N] = wy;
                    b.prt_meta[(size_t)ref_slot * N + car] = (ref_len + 1) | (ref_grass ? 256 : 0);
                }
                meta = (meta & ~0xff00) | ((ref_len + 1) << 8);
            } else if (!(meta & 1)) {
                b.skid_start[(size_t)(k * 2 + 0) * N + car] = wx; b.skid_start[(size_t)(k * 2 + 1) * N + car] = wy;
                meta |= 1;
            } else {
                // _create_particle(skid_start, position, grass): append, pop(0) while more than PRT_MAX
                int head = b.prt_hdr[car], count = b.prt_hdr[(size_t)N + car], slot;
                if (count < PRT_MAX) { slot = head + count; if (slot >= PRT_MAX) slot -= PRT_MAX; ++count; }
                else {
                    slot = head; head = head + 1 == PRT_MAX ? 0 : head + 1;
                    for (int k2 = 0; k2 < 4; ++k2) {     // wheels still extending the popped particle keep doing so, unseen
                        const size_t mi = (size_t)k2 * N + car;
                        if (k2 != k) { const int m2 = b.skid_meta[mi]; if (((m2 >> 16) & 0xff) - 1 == slot) b.skid_meta[mi] = m2 & ~0xff0000; }
                    }
                }
                b.prt_hdr[car] = head; b.prt_hdr[(size_t)N + car] = count;
                float* pt = b.prt_pts + ((size_t)(slot * PRT_PTS) * 2) * N + car;
                pt[0] = b.skid_start[(size_t)(k * 2 + 0) * N + car]; pt[N] = b.skid_start[(size_t)(k * 2 + 1) * N + car];
                pt[(size_t)2 * N] = wx; pt[(size_t)3 * N] = wy;
                b.prt_meta[(size_t)slot * N + car] = 2 | (grass ? 256 : 0);
                meta = 2 | (grass ? 4 : 0) | (2 << 8) | ((slot + 1) << 16);      // skid_start = None
            }
        } else {
            meta = 0;                             // skid_start = None, skid_particle = None
        }
        b.skid_meta[(size_t)k * N + car] = meta;
    }
}

// take_action = false: the action=None path of mcr:421 (reset()'s implicit step, next-step auto reset)
template <typename ActT>
__device__ __forceinline__ void pre_car(const Dims& d, const DevBuffers& b, const CarConst& cc, int car, int env,
                                        bool take_action, const ActT* __restrict__ action) {
    const int N = d.N;
    PRE_CLK_T0();

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

    PRE_CLK(0);
    // ---- controls, mcr:421-424 -------------------------------------------------------
    if (take_action) {
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
    unsigned skid_bits = 0u;
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
        skid_bits |= (fabs(force) > 2.0 * friction_limit ? 1u : 0u) << k;     // Car.step's "Skid trace" condition
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

    if (d.particles) {
        const unsigned grass_bits = (on_road[0] ? 0u : 1u) | (on_road[1] ? 0u : 2u) | (on_road[2] ? 0u : 4u) | (on_road[3] ? 0u : 8u);
        skid_traces(d, b, car, skid_bits, grass_bits, cx[1], cy[1], cx[2], cy[2], cx[3], cy[3], cx[4], cy[4]);
    }

    PRE_CLK(1);
    // ---- b2Island::Solve ---------------------------------------------------------------
    const float h = (float)(1.0 / 50);
    const float mA = cc.hull_invMass, iA = cc.hull_invI, mB = cc.wheel_invMass, iB = cc.wheel_invI;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        vx[1 + k] += h * (mB * Fx[k]);
        vy[1 + k] += h * (mB * Fy[k]);
    }
    // InitVelocityConstraints, joints in island order 3,2,1,0 (coupled envs do it after the contact warm start)
    // Cars of an env with car-car manifolds are solved together by coupled_kernel (contact warm start
    // comes before the joints' in b2Island::Solve), so their joints are NOT initialised here.
    const bool coupled = b.n_manifold[env] > 0;
    JointC J[4];
    if (!coupled) joints_init(cc, ang, motorSpeed, vx, vy, w, jix, jiy, jiz, jmot, lim, J);
    else {
#pragma unroll
        for (int k = 0; k < 4; ++k) { J[k] = JointC(); J[k].motorSpeed = motorSpeed[k]; }
    }
    PRE_CLK(2);
    // ---- hand over to sweep_kernel / post_kernel ---------------------------------------------------
    float* sc = b.scratch + car;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        sc[(size_t)(SC_VX + i) * N] = vx[i]; sc[(size_t)(SC_VY + i) * N] = vy[i]; sc[(size_t)(SC_W + i) * N] = w[i];
        if (i > 0) {   // wheels woken by ApplyForceToCenter; the hull's flag is read by contacts_kernel right now
            b.sleep_time[(size_t)i * N + car] = slp[i];
            b.awake[(size_t)i * N + car] = awake[i] ? 1 : 0;
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        sc[(size_t)(SC_JIX + k) * N] = jix[k]; sc[(size_t)(SC_JIY + k) * N] = jiy[k];
        sc[(size_t)(SC_JIZ + k) * N] = jiz[k]; sc[(size_t)(SC_JMOT + k) * N] = jmot[k];
        const JointC& j = J[k];
        float* q = sc + (size_t)(SC_JOINT + k * SC_JOINT_FIELDS) * N;
        q[(size_t)0 * N] = j.rAx; q[(size_t)1 * N] = j.rAy; q[(size_t)2 * N] = j.k11; q[(size_t)3 * N] = j.k12;
        q[(size_t)4 * N] = j.k22; q[(size_t)5 * N] = j.ezx; q[(size_t)6 * N] = j.ezy; q[(size_t)7 * N] = j.ezz;
        q[(size_t)8 * N] = j.det22; q[(size_t)9 * N] = j.cfx; q[(size_t)10 * N] = j.cfy; q[(size_t)11 * N] = j.cfz;
        q[(size_t)12 * N] = j.det33; q[(size_t)13 * N] = j.motorMass; q[(size_t)14 * N] = j.motorSpeed;
        b.limit_state[(size_t)k * N + car] = (uint8_t)lim[k];
        b.wheel[(size_t)(k * WHEEL_FIELDS + WF_OMEGA) * N + car] = omega[k];
        b.wheel[(size_t)(k * WHEEL_FIELDS + WF_PHASE) * N + car] = phase[k];
    }
    b.ctrl[(size_t)CF_GAS * N + car] = gas;
    b.ctrl[(size_t)CF_BRAKE * N + car] = brake;
    b.ctrl[(size_t)CF_STEER * N + car] = steer;
    PRE_CLK(3);
}


// The same head with FOUR lanes per car (lane k of the group = wheel k = joint k; the four lanes are consecutive lanes
// of one warp, gmask = the lanes of the warp that execute this function).  The per-car form above runs the four wheels
// and the four joints as unrolled straight-line code on one lane -- ~3000 instructions per warp, and the step's head is
// bound by fetching them (the code is cold every step: 6-13 cycles per instruction, whatever the L2 holds; scripts/
// head_phases.py); here a warp executes a quarter of the wheel / joint code.  Same operations on the same values in the
// same order: the hull's warm-start chain (joint order 3, 2, 1, 0) is evaluated by every lane of the group from the
// other lanes' shuffled terms.
// coupled_in: 0 / 1 = the env has no / has car-car manifolds (known to the caller), -1 = read n_manifold[env]
template <typename ActT>
__device__ __forceinline__ void pre_car4(const Dims& d, const DevBuffers& b, const CarConst& cc, int car, int env, int k, unsigned gmask,
                                         bool take_action, const ActT* __restrict__ action, int coupled_in = -1) {
    const int N = d.N, bi = 1 + k;
    const int l0 = (threadIdx.x & 31) & ~3;                  // first lane of the group
    PRE_CLK_T0();
    // ---- load state: the hull and this lane's wheel / joint --------------------------------------------------
    const float* ph = b.body + car;
    const float* pw = b.body + (size_t)(bi * BODY_FIELDS) * N + car;
    const float ang0 = ph[(size_t)BF_A * N];
    float vx0 = ph[(size_t)BF_VX * N], vy0 = ph[(size_t)BF_VY * N], w0 = ph[(size_t)BF_W * N];
    const float cxw = pw[(size_t)BF_CX * N], cyw = pw[(size_t)BF_CY * N], angw = pw[(size_t)BF_A * N];
    float vxw = pw[(size_t)BF_VX * N], vyw = pw[(size_t)BF_VY * N], ww = pw[(size_t)BF_W * N];
    const float qsw = pw[(size_t)BF_QS * N], qcw = pw[(size_t)BF_QC * N];
    float slpw = b.sleep_time[(size_t)bi * N + car];
    bool awakew = b.awake[(size_t)bi * N + car] != 0;
    const float* pj = b.joint + (size_t)(k * JOINT_FIELDS) * N + car;
    const float jix = pj[(size_t)JF_IX * N], jiy = pj[(size_t)JF_IY * N], jmot = pj[(size_t)JF_MOTOR * N];
    float jiz = pj[(size_t)JF_IZ * N];
    int lim = b.limit_state[(size_t)k * N + car];
    double omega = b.wheel[(size_t)(k * WHEEL_FIELDS + WF_OMEGA) * N + car];
    double phase = b.wheel[(size_t)(k * WHEEL_FIELDS + WF_PHASE) * N + car];
    const bool on_road = b.on_road[(size_t)k * N + car] != 0;
    double gas = b.ctrl[(size_t)CF_GAS * N + car];
    double brake = b.ctrl[(size_t)CF_BRAKE * N + car];
    double steer = b.ctrl[(size_t)CF_STEER * N + car];
    PRE_CLK(0);
    // ---- controls, mcr:421-424 (every lane of the group, same values) ------------------------------------------
    if (take_action) {
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
    // ---- Car.step(dt): tyre model (float64) of wheel k ----------------------------------------------------------
    const double SIZE = 0.02;
    const double ENGINE_POWER = 100000000 * SIZE * SIZE;
    const double WHEEL_MOI = 4000 * SIZE * SIZE;
    const double FRICTION_LIMIT = 1000000 * SIZE * SIZE;
    const double dt = 1.0 / 50;
    float motorSpeed, Fx, Fy;
    bool skid;
    {
        const double steer_w = k < 2 ? steer : 0.0;
        const double gas_w = k >= 2 ? gas : 0.0;
        double jangle = (double)(angw - ang0);
        double dir = sign_d(steer_w - jangle);
        double val = fabs(steer_w - jangle);
        motorSpeed = (float)(dir * fmin(50.0 * val, 3.0));
        double friction_limit = FRICTION_LIMIT * 0.6;
        if (on_road) friction_limit = fmax(friction_limit, FRICTION_LIMIT * 1.0);
        float forw_x = qcw * 0.0f - qsw * 1.0f, forw_y = qsw * 0.0f + qcw * 1.0f;
        float side_x = qcw * 1.0f - qsw * 0.0f, side_y = qsw * 1.0f + qcw * 0.0f;
        double wvx = vxw, wvy = vyw;
        double vf = (double)forw_x * wvx + (double)forw_y * wvy;
        double vs = (double)side_x * wvx + (double)side_y * wvy;
        omega += dt * ENGINE_POWER * gas_w / WHEEL_MOI / (fabs(omega) + 5.0);
        if (brake >= 0.9) {
            omega = 0;
        } else if (brake > 0) {
            double bdir = -sign_d(omega);
            double bval = 15 * brake;
            if (fabs(bval) > fabs(omega)) bval = fabs(omega);
            omega += bdir * bval;
        }
        phase += omega * dt;
        const double wheel_rad = 1.0 * 27 * SIZE;
        double vr = omega * wheel_rad;
        double f_force = -vf + vr;
        double p_force = -vs;
        f_force *= 205000 * SIZE * SIZE;
        p_force *= 205000 * SIZE * SIZE;
        double force = sqrt(f_force * f_force + p_force * p_force);
        skid = fabs(force) > 2.0 * friction_limit;
        if (fabs(force) > friction_limit) {
            f_force /= force; p_force /= force;
            force = friction_limit;
            f_force *= force; p_force *= force;
        }
        omega -= dt * f_force * wheel_rad / WHEEL_MOI;
        Fx = (float)(p_force * (double)side_x + f_force * (double)forw_x);
        Fy = (float)(p_force * (double)side_y + f_force * (double)forw_y);
        if (!awakew) { awakew = true; slpw = 0.0f; }              // ApplyForceToCenter(wake=True)
    }
    if (d.particles) {
        // (uniform over the warp) the wheels' flags and positions gathered on the group's first lane
        const unsigned sk = __ballot_sync(gmask, skid), gr = __ballot_sync(gmask, !on_road);
        const float x1 = __shfl_sync(gmask, cxw, l0 + 1), y1 = __shfl_sync(gmask, cyw, l0 + 1);
        const float x2 = __shfl_sync(gmask, cxw, l0 + 2), y2 = __shfl_sync(gmask, cyw, l0 + 2);
        const float x3 = __shfl_sync(gmask, cxw, l0 + 3), y3 = __shfl_sync(gmask, cyw, l0 + 3);
        if (k == 0) skid_traces(d, b, car, (sk >> l0) & 15u, (gr >> l0) & 15u, cxw, cyw, x1, y1, x2, y2, x3, y3);
    }
    PRE_CLK(1);
    // ---- b2Island::Solve ---------------------------------------------------------------------------------------
    const float h = (float)(1.0 / 50);
    const float mA = cc.hull_invMass, iA = cc.hull_invI, mB = cc.wheel_invMass, iB = cc.wheel_invI;
    vxw += h * (mB * Fx);
    vyw += h * (mB * Fy);
    const bool coupled = coupled_in >= 0 ? coupled_in != 0 : b.n_manifold[env] > 0;
    JointC j = JointC();
    j.motorSpeed = motorSpeed;
    if (!coupled) {
        // joints_init for joint k; the hull's part of the warm start in the island's joint order 3, 2, 1, 0
        float sA, cA; rot_set(ang0, sA, cA);
        const float lx = cc.anchor_x[k] - cc.hull_lcx, ly = cc.anchor_y[k] - cc.hull_lcy;
        const float rAx = cA * lx - sA * ly, rAy = sA * lx + cA * ly;
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
        solve33_init(j);
        const float jointAngle = angw - ang0 - 0.0f;
        if (jointAngle <= cc.lower) {
            if (lim != LIM_LOWER) jiz = 0.0f;
            lim = LIM_LOWER;
        } else if (jointAngle >= cc.upper) {
            if (lim != LIM_UPPER) jiz = 0.0f;
            lim = LIM_UPPER;
        } else {
            lim = LIM_INACTIVE;
            jiz = 0.0f;
        }
        j.limit = lim;
        // warm start: this joint's terms, then the hull chain over the four joints
        const float hull_w_term = (rAx * jiy - rAy * jix) + jmot + jiz;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const int src = l0 + 3 - kk;
            const float Px = __shfl_sync(gmask, jix, src), Py = __shfl_sync(gmask, jiy, src), wt = __shfl_sync(gmask, hull_w_term, src);
            vx0 -= mA * Px; vy0 -= mA * Py;
            w0 -= iA * wt;
        }
        vxw += mB * jix; vyw += mB * jiy;
        ww += iB * (jmot + jiz);
    }
    PRE_CLK(2);
    // ---- hand over to sweep_kernel / post_kernel ---------------------------------------------------------------
    float* sc = b.scratch + car;
    sc[(size_t)(SC_VX + bi) * N] = vxw; sc[(size_t)(SC_VY + bi) * N] = vyw; sc[(size_t)(SC_W + bi) * N] = ww;
    b.sleep_time[(size_t)bi * N + car] = slpw;
    b.awake[(size_t)bi * N + car] = awakew ? 1 : 0;
    sc[(size_t)(SC_JIX + k) * N] = jix; sc[(size_t)(SC_JIY + k) * N] = jiy;
    sc[(size_t)(SC_JIZ + k) * N] = jiz; sc[(size_t)(SC_JMOT + k) * N] = jmot;
    float* q = sc + (size_t)(SC_JOINT + k * SC_JOINT_FIELDS) * N;
    q[(size_t)0 * N] = j.rAx; q[(size_t)1 * N] = j.rAy; q[(size_t)2 * N] = j.k11; q[(size_t)3 * N] = j.k12;
    q[(size_t)4 * N] = j.k22; q[(size_t)5 * N] = j.ezx; q[(size_t)6 * N] = j.ezy; q[(size_t)7 * N] = j.ezz;
    q[(size_t)8 * N] = j.det22; q[(size_t)9 * N] = j.cfx; q[(size_t)10 * N] = j.cfy; q[(size_t)11 * N] = j.cfz;
    q[(size_t)12 * N] = j.det33; q[(size_t)13 * N] = j.motorMass; q[(size_t)14 * N] = j.motorSpeed;
    b.limit_state[(size_t)k * N + car] = (uint8_t)lim;
    b.wheel[(size_t)(k * WHEEL_FIELDS + WF_OMEGA) * N + car] = omega;
    b.wheel[(size_t)(k * WHEEL_FIELDS + WF_PHASE) * N + car] = phase;
    if (k == 0) {
        sc[(size_t)(SC_VX + 0) * N] = vx0; sc[(size_t)(SC_VY + 0) * N] = vy0; sc[(size_t)(SC_W + 0) * N] = w0;
        b.ctrl[(size_t)CF_GAS * N + car] = gas;
        b.ctrl[(size_t)CF_BRAKE * N + car] = brake;
        b.ctrl[(size_t)CF_STEER * N + car] = steer;
    }
    PRE_CLK(3);
}
