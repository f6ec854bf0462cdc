// models.cuh -- device-side single-agent dynamics library (sm_100a).
//
// Each model provides the continuous ODE xdot = f(x, u), the zero-order-hold RK4 step the
// reference integrates with, and the analytic continuous-time Jacobian entries, which the
// caller discretises with forward Euler (A = I + dt*dfdx, B = dt*dfdu) exactly as the
// reference does.  Everything is templated on the model id so state vectors live in
// registers with fully static indexing.
//
// Behaviour follows (not copies) the reference:
//   ODEs + Jacobians   dpilqr/bbdynamics.cpp:108-711, Bike5D dpilqr/dynamics.py:254-277
//   RK4, 5 sub-steps   dpilqr/bbdynamics.cpp:39-93  (Bike5D: ONE step, dynamics.py:18-38,74)
//   Euler Jacobians    dpilqr/bbdynamics.cpp:95-106, dynamics.py:112-114
#pragma once
#include <cuda_runtime.h>

namespace dpilqr {

enum ModelId : int {
    kDoubleInt4D = 0,
    kDoubleInt6D = 1,
    kCar3D = 2,
    kUnicycle4D = 3,
    kQuad6D = 4,
    kHuman6D = 5,
    kHumanLin6D = 6,
    kQuad12D = 7,
    kBike5D = 8,
    kModelCount = 9
};

__host__ __device__ constexpr int model_nx(int m)
{
    return m == kDoubleInt4D ? 4 : m == kDoubleInt6D ? 6 : m == kCar3D ? 3 : m == kUnicycle4D ? 4
         : m == kQuad6D ? 6 : m == kHuman6D ? 6 : m == kHumanLin6D ? 6 : m == kQuad12D ? 12
         : m == kBike5D ? 5 : -1;
}
__host__ __device__ constexpr int model_nu(int m)
{
    return m == kDoubleInt4D ? 2 : m == kDoubleInt6D ? 3 : m == kCar3D ? 2 : m == kUnicycle4D ? 2
         : m == kQuad6D ? 3 : m == kHuman6D ? 3 : m == kHumanLin6D ? 3 : m == kQuad12D ? 4
         : m == kBike5D ? 2 : -1;
}

// Pseudo model ids for kernels instantiated per (s, c) size class rather than per model: teams that mix models of
// one size class (zero-padded heterogeneous teams, reference scripts/examples.py:73-131) dispatch per agent inside.
constexpr int kMixed4 = 100;  // DoubleInt4D | Unicycle4D
constexpr int kMixed6 = 101;  // DoubleInt6D | Quad6D | Human6D | HumanLin6D
__host__ __device__ constexpr int class_nx(int mc) { return mc == kMixed4 ? 4 : mc == kMixed6 ? 6 : model_nx(mc); }
__host__ __device__ constexpr int class_nu(int mc) { return mc == kMixed4 ? 2 : mc == kMixed6 ? 3 : model_nu(mc); }

constexpr double kGravity = 9.80665;
// Quadcopter12D rigid-body constants: thrust/mass gain, torque/inertia gains and the
// gyroscopic coupling ratios (I_j - I_k) / I_i of the airframe the reference models.
constexpr double kThrustGain = 2000.0 / 63.0;
constexpr double kTauX = 625000000000000000.0 / 10982593196059.0;
constexpr double kTauY = 5000000000000000000.0 / 92848985528431.0;
constexpr double kTauZ = 10000000000000000000.0 / 271597947137541.0;
constexpr double kGyroX = 85899976080679.0 / 175721491136944.0;
constexpr double kGyroY = 95876456000597.0 / 185697971056862.0;
constexpr double kGyroZ = 9976479919918.0 / 271597947137541.0;

// ------------------------------------------------------------------------------------------
// sin and cos for the ODE right-hand sides.  RK4 calls them sixty times per agent and step, and the library sincos
// (about 125 instructions with its large-argument machinery) dominated the rollout kernel.  This is the textbook
// scheme in about 35: two-constant Cody-Waite reduction by pi/2 with FMAs, then the fdlibm minimax polynomials on
// [-pi/4, pi/4] (error about one ulp).  Arguments beyond 1e5 (runaway line-search candidates) and NaN take the
// library path.
// ------------------------------------------------------------------------------------------
static __device__ __noinline__ void ode_sincos_library(double x, double *sn, double *cs) { sincos(x, sn, cs); }

// Polynomial and reduction constants live in constant memory: as immediates every one of them costs two UMOV
// instructions per use inside the rolled integrator loop (a fifth of all issued instructions, measured), as
// constant-bank operands they ride along in the DFMA for free.
static __constant__ double kSinCosC[16] = {
    6.36619772367581382433e-01,   // 0  2/pi
    1.5707963267948966,           // 1  pi/2 high
    6.123233995736766e-17,        // 2  pi/2 low
    1.58969099521155010221e-10,   // 3  sin
    -2.50507602534068634195e-08,  // 4
    2.75573137070700676789e-06,   // 5
    -1.98412698298579493134e-04,  // 6
    8.33333333332248946124e-03,   // 7
    -1.66666666666666324348e-01,  // 8
    -1.13596475577881948265e-11,  // 9  cos
    2.08757232129817482790e-09,   // 10
    -2.75573143513906633035e-07,  // 11
    2.48015872894767294178e-05,   // 12
    -1.38888888888741095749e-03,  // 13
    4.16666666666666019037e-02,   // 14
    0.0};

__device__ __forceinline__ void ode_sincos(double x, double *sn, double *cs)
{
    if (!(fabs(x) < 1.0e5)) {  // rare; out of line so the hot path stays compact
        ode_sincos_library(x, sn, cs);
        return;
    }
    const double *K = kSinCosC;
    const double kd = rint(x * K[0]);
    const int k = __double2int_rn(kd);
    double r = fma(-kd, K[1], x);
    r = fma(-kd, K[2], r);
    const double z = r * r;
    double ps = fma(z, K[3], K[4]);
    ps = fma(z, ps, K[5]);
    ps = fma(z, ps, K[6]);
    ps = fma(z, ps, K[7]);
    ps = fma(z, ps, K[8]);
    const double s = fma(r * z, ps, r);
    double pc = fma(z, K[9], K[10]);
    pc = fma(z, pc, K[11]);
    pc = fma(z, pc, K[12]);
    pc = fma(z, pc, K[13]);
    pc = fma(z, pc, K[14]);
    const double c = fma(z * z, pc, fma(-0.5, z, 1.0));
    double ss = (k & 1) ? c : s;
    double cc = (k & 1) ? s : c;
    if (k & 2) ss = -ss;
    if ((k + 1) & 2) cc = -cc;
    *sn = ss;
    *cs = cc;
}

// Three angles at once, as one straight-line block: the compiler interleaves the three dependency chains (evaluated
// one after the other -- each behind its own range-check branch -- they cost three times the latency).
__device__ __forceinline__ void ode_sincos3(const double (&x)[3], double (&sn)[3], double (&cs)[3])
{
    if (!(fabs(x[0]) < 1.0e5 && fabs(x[1]) < 1.0e5 && fabs(x[2]) < 1.0e5)) {
        for (int i = 0; i < 3; ++i) ode_sincos_library(x[i], &sn[i], &cs[i]);
        return;
    }
    const double *K = kSinCosC;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double kd = rint(x[i] * K[0]);
        const int k = __double2int_rn(kd);
        double r = fma(-kd, K[1], x[i]);
        r = fma(-kd, K[2], r);
        const double z = r * r;
        double ps = fma(z, K[3], K[4]);
        ps = fma(z, ps, K[5]);
        ps = fma(z, ps, K[6]);
        ps = fma(z, ps, K[7]);
        ps = fma(z, ps, K[8]);
        const double s = fma(r * z, ps, r);
        double pc = fma(z, K[9], K[10]);
        pc = fma(z, pc, K[11]);
        pc = fma(z, pc, K[12]);
        pc = fma(z, pc, K[13]);
        pc = fma(z, pc, K[14]);
        const double c = fma(z * z, pc, fma(-0.5, z, 1.0));
        double ss = (k & 1) ? c : s;
        double cc = (k & 1) ? s : c;
        if (k & 2) ss = -ss;
        if ((k + 1) & 2) cc = -cc;
        sn[i] = ss;
        cs[i] = cc;
    }
}

// ------------------------------------------------------------------------------------------
// xdot = f(x, u)
// ------------------------------------------------------------------------------------------
template <int M>
__device__ __forceinline__ void model_f_inline(const double (&x)[model_nx(M)], const double (&u)[model_nu(M)],
                                               double (&xd)[model_nx(M)])
{
    if constexpr (M == kDoubleInt4D) {
        xd[0] = x[2]; xd[1] = x[3]; xd[2] = u[0]; xd[3] = u[1];
    } else if constexpr (M == kDoubleInt6D) {
        xd[0] = x[3]; xd[1] = x[4]; xd[2] = x[5]; xd[3] = u[0]; xd[4] = u[1]; xd[5] = u[2];
    } else if constexpr (M == kCar3D) {
        double sn, cs;
        ode_sincos(x[2], &sn, &cs);
        xd[0] = u[0] * cs; xd[1] = u[0] * sn; xd[2] = u[1];
    } else if constexpr (M == kUnicycle4D) {
        double sn, cs;
        ode_sincos(x[3], &sn, &cs);
        xd[0] = x[2] * cs; xd[1] = x[2] * sn; xd[2] = u[0]; xd[3] = u[1];
    } else if constexpr (M == kQuad6D) {
        xd[0] = x[3]; xd[1] = x[4]; xd[2] = x[5];
        xd[3] = kGravity * tan(u[2]);
        xd[4] = -kGravity * tan(u[1]);
        xd[5] = u[0] - kGravity;
    } else if constexpr (M == kHuman6D) {
        // planar unicycle at constant height whose heading is a *control*
        double sn, cs;
        ode_sincos(u[0], &sn, &cs);
        xd[0] = x[3] * cs; xd[1] = x[3] * sn; xd[2] = 0.0; xd[3] = u[1]; xd[4] = 0.0; xd[5] = 0.0;
    } else if constexpr (M == kHumanLin6D) {
        xd[0] = x[3]; xd[1] = x[4]; xd[2] = 0.0; xd[3] = u[0]; xd[4] = u[1]; xd[5] = 0.0;
    } else if constexpr (M == kQuad12D) {
        // x = [p(3), yaw, pitch, roll, v_body(3), w_body(3)], u = [tau(3), thrust]
        double sy, cy, sp, cp, sr, cr;
        ode_sincos(x[3], &sy, &cy);
        ode_sincos(x[4], &sp, &cp);
        ode_sincos(x[5], &sr, &cr);
        const double icp = 1.0 / cp;
        const double tp = sp * icp;
        const double v0 = x[6], v1 = x[7], v2 = x[8];
        const double w0 = x[9], w1 = x[10], w2 = x[11];
        const double srsp = sr * sp, crsp = cr * sp;
        xd[0] = v0 * (cy * cp) + v1 * (srsp * cy - sy * cr) + v2 * (sr * sy + crsp * cy);
        xd[1] = v0 * (sy * cp) + v1 * (srsp * sy + cr * cy) + v2 * (crsp * sy - sr * cy);
        xd[2] = v1 * (sr * cp) - v0 * sp + v2 * (cr * cp);
        const double wq = w1 * sr + w2 * cr;
        xd[3] = wq * icp;
        xd[4] = w1 * cr - w2 * sr;
        xd[5] = w0 + wq * tp;
        xd[6] = v1 * w2 - v2 * w1 + kGravity * sp;
        xd[7] = v2 * w0 - v0 * w2 - kGravity * (sr * cp);
        xd[8] = kThrustGain * u[3] + v0 * w1 - v1 * w0 - kGravity * (cr * cp);
        xd[9] = kTauX * u[0] - kGyroX * (w1 * w2);
        xd[10] = kTauY * u[1] + kGyroY * (w0 * w2);
        xd[11] = kTauZ * u[2] - kGyroZ * (w0 * w1);
    } else if constexpr (M == kBike5D) {
        double sn, cs;
        ode_sincos(x[3], &sn, &cs);
        xd[0] = x[2] * cs; xd[1] = x[2] * sn; xd[2] = u[0]; xd[3] = x[2] * tan(x[4]); xd[4] = u[1];
    }
}

// Out-of-line variant (kept for experiments; see model_f).
template <int M>
__device__ __noinline__ void model_f_outlined(const double (&x)[model_nx(M)], const double (&u)[model_nu(M)],
                                              double (&xd)[model_nx(M)])
{
    model_f_inline<M>(x, u, xd);
}

template <int M>
__device__ __forceinline__ void model_f(const double (&x)[model_nx(M)], const double (&u)[model_nu(M)],
                                        double (&xd)[model_nx(M)])
{
    // Out-of-line evaluation passes the state through local memory, which thrashes once shared memory has taken
    // most of the L1 (measured: 7.6 % L1 hit rate on the spill traffic of the rollout kernel), so everything inlines.
    model_f_inline<M>(x, u, xd);
}

// ------------------------------------------------------------------------------------------
// x <- Phi_dt(x, u): classic RK4 with zero-order-hold controls.
// ------------------------------------------------------------------------------------------
template <int M>
__device__ __forceinline__ void model_step(double dt, double (&x)[model_nx(M)], const double (&u)[model_nu(M)])
{
    constexpr int NX = model_nx(M);
    constexpr int kSub = (M == kBike5D) ? 1 : 5;
    const double h = (M == kBike5D) ? dt : dt / 5;
    const double hh = h / 2.0;
    const double h6 = h / 6.0;
    // One rolled loop over the 4 * kSub stage evaluations: the body holds ONE inlined copy of the ODE (the
    // twenty-fold unrolled form is about 30 kB of SASS for Quadcopter12D and misses the instruction cache).
    // Stage s of a sub-step:  k = f(xs);  acc += w_s k  (w = 1, 2, 2, 1);  xs = x + c_s k  (c = h/2, h/2, h);
    // after stage 3:  x += h/6 acc.  Same operations, in the same order, as the textbook form.
    double acc[NX], xs[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) { acc[i] = 0.0; xs[i] = x[i]; }
#pragma unroll 1
    for (int ev = 0; ev < 4 * kSub; ++ev) {
        const int stage = ev & 3;
        double k[NX];
        model_f<M>(xs, u, k);
        const double w = (stage == 0 || stage == 3) ? 1.0 : 2.0;
        const double cs = (stage == 2) ? h : hh;
        if (stage == 3) {
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                x[i] += h6 * (acc[i] + k[i]);
                xs[i] = x[i];
                acc[i] = 0.0;
            }
        } else {
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                acc[i] = fma(w, k[i], acc[i]);
                xs[i] = fma(cs, k[i], x[i]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Quadcopter12D integrated by a TEAM of three warps (the rollout kernel): lane q of every warp of the team works on
// agent q of the team's 32 agents, and the warp's ROLE owns a slice of the state and of the RK4 arithmetic --
//   role 0: Euler angles (x3..5): the three sin/cos pairs and the angle rates -- the serial spine of the ODE
//           (angles -> sincos -> rates -> angles);
//   role 1: position p (x0..2) = R(angles) v;
//   role 2: body velocity and rates v, w (x6..11).
// Per ODE evaluation role 2 publishes v, w and role 0 the six trig values of the stage state in a double-buffered
// shared-memory scratch [2][12][32] (component-major: conflict-free), then ONE named barrier of 96 threads, then
// every role evaluates its slice.  Same expressions as model_f_inline<kQuad12D> / model_step; a warp only ever runs
// its own role's code (no divergence), so an evaluation costs the spine instead of the whole ODE -- the rollout is
// bound by the latency of 20 dependent ODE evaluations per step, and a batch of 4096 ten-drone problems is only
// 41 k agents, a seventh of the threads the GPU holds.
// ------------------------------------------------------------------------------------------
constexpr int kTeamRoles = 3;
constexpr int kTeamThreads = 32 * kTeamRoles;
constexpr int kTeamScratch = 2 * 12 * 32;  // doubles per team
__host__ __device__ constexpr int team_role_offset(int role) { return role == 0 ? 3 : role == 1 ? 0 : 6; }  // first state component
__host__ __device__ constexpr int team_role_count(int role) { return role == 2 ? 6 : 3; }

__device__ __forceinline__ void team_barrier(int id)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kTeamThreads) : "memory");
}

// x: the role's state slice (3 components, 6 for role 2), advanced in place by one step of length dt
template <int ROLE>
__device__ __forceinline__ void quad12_step_team(double dt, int q, int bar_id, double (&x)[6], const double (&u)[4], double *scratch)
{
    constexpr int NC = team_role_count(ROLE);
    const double h = dt / 5;
    const double hh = h / 2.0;
    const double h6 = h / 6.0;
    double acc[NC], xs[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) { acc[i] = 0.0; xs[i] = x[i]; }
#pragma unroll 1
    for (int ev = 0; ev < 20; ++ev) {
        const int stage = ev & 3;
        double *xsh = scratch + (ev & 1) * (12 * 32) + q;  // [12][32]: 0..2 v, 3..5 w, 6..11 sy cy sp cp sr cr
        double sp, cp, sr, cr, icp = 0.0;
        if constexpr (ROLE == 0) {
            double sn3[3], cs3[3];
            ode_sincos3(xs, sn3, cs3);
            sp = sn3[1]; cp = cs3[1]; sr = sn3[2]; cr = cs3[2];
            xsh[6 * 32] = sn3[0]; xsh[7 * 32] = cs3[0]; xsh[8 * 32] = sp; xsh[9 * 32] = cp; xsh[10 * 32] = sr; xsh[11 * 32] = cr;
            icp = 1.0 / cp;  // issued before the barrier: the division completes in its shadow
        } else if constexpr (ROLE == 2) {
#pragma unroll
            for (int i = 0; i < 6; ++i) xsh[i * 32] = xs[i];
        }
        team_barrier(bar_id);
        double k[NC];
        if constexpr (ROLE == 0) {
            const double w0 = xsh[3 * 32], w1 = xsh[4 * 32], w2 = xsh[5 * 32];
            const double tp = sp * icp;
            const double wq = w1 * sr + w2 * cr;
            k[0] = wq * icp;
            k[1] = w1 * cr - w2 * sr;
            k[2] = w0 + wq * tp;
        } else if constexpr (ROLE == 1) {
            const double sy = xsh[6 * 32], cy = xsh[7 * 32];
            sp = xsh[8 * 32]; cp = xsh[9 * 32]; sr = xsh[10 * 32]; cr = xsh[11 * 32];
            const double v0 = xsh[0], v1 = xsh[1 * 32], v2 = xsh[2 * 32];
            const double srsp = sr * sp, crsp = cr * sp;
            k[0] = v0 * (cy * cp) + v1 * (srsp * cy - sy * cr) + v2 * (sr * sy + crsp * cy);
            k[1] = v0 * (sy * cp) + v1 * (srsp * sy + cr * cy) + v2 * (crsp * sy - sr * cy);
            k[2] = v1 * (sr * cp) - v0 * sp + v2 * (cr * cp);
        } else {
            sp = xsh[8 * 32]; cp = xsh[9 * 32]; sr = xsh[10 * 32]; cr = xsh[11 * 32];
            const double v0 = xs[0], v1 = xs[1], v2 = xs[2];
            const double w0 = xs[3], w1 = xs[4], w2 = xs[5];
            k[0] = v1 * w2 - v2 * w1 + kGravity * sp;
            k[1] = v2 * w0 - v0 * w2 - kGravity * (sr * cp);
            k[2] = kThrustGain * u[3] + v0 * w1 - v1 * w0 - kGravity * (cr * cp);
            k[3] = kTauX * u[0] - kGyroX * (w1 * w2);
            k[4] = kTauY * u[1] + kGyroY * (w0 * w2);
            k[5] = kTauZ * u[2] - kGyroZ * (w0 * w1);
        }
        const double w = (stage == 0 || stage == 3) ? 1.0 : 2.0;
        const double cst = (stage == 2) ? h : hh;
        if (stage == 3) {
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                x[i] += h6 * (acc[i] + k[i]);
                xs[i] = x[i];
                acc[i] = 0.0;
            }
        } else {
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                acc[i] = fma(w, k[i], acc[i]);
                xs[i] = fma(cst, k[i], x[i]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Continuous-time Jacobian entries.  `Sink` receives the structurally non-zero entries through
// sink.a(row, col, value) / sink.b(row, col, value); everything else is zero.
// ------------------------------------------------------------------------------------------
template <int M, class Sink>
__device__ __forceinline__ void model_jacobian(const double (&x)[model_nx(M)], const double (&u)[model_nu(M)], Sink &sink)
{
    if constexpr (M == kDoubleInt4D) {
        sink.a(0, 2, 1.0); sink.a(1, 3, 1.0);
        sink.b(2, 0, 1.0); sink.b(3, 1, 1.0);
    } else if constexpr (M == kDoubleInt6D) {
        sink.a(0, 3, 1.0); sink.a(1, 4, 1.0); sink.a(2, 5, 1.0);
        sink.b(3, 0, 1.0); sink.b(4, 1, 1.0); sink.b(5, 2, 1.0);
    } else if constexpr (M == kCar3D) {
        double sn, cs;
        sincos(x[2], &sn, &cs);
        sink.a(0, 2, -u[0] * sn); sink.a(1, 2, u[0] * cs);
        sink.b(0, 0, cs); sink.b(1, 0, sn); sink.b(2, 1, 1.0);
    } else if constexpr (M == kUnicycle4D) {
        double sn, cs;
        sincos(x[3], &sn, &cs);
        sink.a(0, 2, cs); sink.a(0, 3, -x[2] * sn);
        sink.a(1, 2, sn); sink.a(1, 3, x[2] * cs);
        sink.b(2, 0, 1.0); sink.b(3, 1, 1.0);
    } else if constexpr (M == kQuad6D) {
        const double t1 = tan(u[1]), t2 = tan(u[2]);
        sink.a(0, 3, 1.0); sink.a(1, 4, 1.0); sink.a(2, 5, 1.0);
        sink.b(3, 2, kGravity * (t2 * t2) + kGravity);
        sink.b(4, 1, -kGravity * (t1 * t1) - kGravity);
        sink.b(5, 0, 1.0);
    } else if constexpr (M == kHuman6D) {
        double sn, cs;
        sincos(u[0], &sn, &cs);
        sink.a(0, 3, cs); sink.a(1, 3, sn);
        sink.b(0, 0, -x[3] * sn); sink.b(1, 0, x[3] * cs); sink.b(3, 1, 1.0);
    } else if constexpr (M == kHumanLin6D) {
        sink.a(0, 3, 1.0); sink.a(1, 4, 1.0);
        sink.b(3, 0, 1.0); sink.b(4, 1, 1.0);
    } else if constexpr (M == kQuad12D) {
        double sy, cy, sp, cp, sr, cr;
        sincos(x[3], &sy, &cy);
        sincos(x[4], &sp, &cp);
        sincos(x[5], &sr, &cr);
        const double icp = 1.0 / cp;
        const double tp = sp * icp;
        const double sec2 = tp * tp + 1.0;
        const double v0 = x[6], v1 = x[7], v2 = x[8];
        const double w0 = x[9], w1 = x[10], w2 = x[11];
        // rotation matrix body -> world, R = Rz(yaw) Ry(pitch) Rx(roll)
        const double r00 = cy * cp, r01 = sr * sp * cy - sy * cr, r02 = sr * sy + sp * cr * cy;
        const double r10 = sy * cp, r11 = sr * sy * sp + cr * cy, r12 = sy * sp * cr - sr * cy;
        const double r20 = -sp, r21 = sr * cp, r22 = cr * cp;
        // d(R v)/d yaw = [-row1; row0; 0]
        sink.a(0, 3, -(v0 * r10 + v1 * r11 + v2 * r12));
        sink.a(1, 3, v0 * r00 + v1 * r01 + v2 * r02);
        // d(R v)/d pitch
        const double q = -v0 * sp + v1 * (sr * cp) + v2 * (cr * cp);
        sink.a(0, 4, cy * q);
        sink.a(1, 4, sy * q);
        sink.a(2, 4, -v0 * cp - v1 * (sr * sp) - v2 * (sp * cr));
        // d(R v)/d roll
        sink.a(0, 5, v1 * r02 - v2 * r01);
        sink.a(1, 5, v1 * r12 - v2 * r11);
        sink.a(2, 5, v1 * r22 - v2 * r21);
        // d(R v)/d v = R
        sink.a(0, 6, r00); sink.a(0, 7, r01); sink.a(0, 8, r02);
        sink.a(1, 6, r10); sink.a(1, 7, r11); sink.a(1, 8, r12);
        sink.a(2, 6, r20); sink.a(2, 7, r21); sink.a(2, 8, r22);
        // Euler-angle kinematics
        const double wq = w1 * sr + w2 * cr;  // d/d roll of which is wr
        const double wr = w1 * cr - w2 * sr;
        sink.a(3, 4, wq * sp * (icp * icp));
        sink.a(3, 5, wr * icp);
        sink.a(3, 10, sr * icp);
        sink.a(3, 11, cr * icp);
        sink.a(4, 5, -wq);
        sink.a(4, 10, cr);
        sink.a(4, 11, -sr);
        sink.a(5, 4, wq * sec2);
        sink.a(5, 5, wr * tp);
        sink.a(5, 9, 1.0);
        sink.a(5, 10, sr * tp);
        sink.a(5, 11, cr * tp);
        // body-frame translational dynamics
        sink.a(6, 4, kGravity * cp);
        sink.a(6, 7, w2); sink.a(6, 8, -w1); sink.a(6, 10, -v2); sink.a(6, 11, v1);
        sink.a(7, 4, kGravity * (sr * sp));
        sink.a(7, 5, -kGravity * (cr * cp));
        sink.a(7, 6, -w2); sink.a(7, 8, w0); sink.a(7, 9, v2); sink.a(7, 11, -v0);
        sink.a(8, 4, kGravity * (sp * cr));
        sink.a(8, 5, kGravity * (sr * cp));
        sink.a(8, 6, w1); sink.a(8, 7, -w0); sink.a(8, 9, -v1); sink.a(8, 10, v0);
        // Euler's rotation equations
        sink.a(9, 10, -kGyroX * w2); sink.a(9, 11, -kGyroX * w1);
        sink.a(10, 9, kGyroY * w2); sink.a(10, 11, kGyroY * w0);
        sink.a(11, 9, -kGyroZ * w1); sink.a(11, 10, -kGyroZ * w0);
        sink.b(8, 3, kThrustGain);
        sink.b(9, 0, kTauX); sink.b(10, 1, kTauY); sink.b(11, 2, kTauZ);
    } else if constexpr (M == kBike5D) {
        double sn, cs;
        sincos(x[3], &sn, &cs);
        const double tf = tan(x[4]);
        sink.a(0, 2, cs); sink.a(0, 3, -x[2] * sn);
        sink.a(1, 2, sn); sink.a(1, 3, x[2] * cs);
        sink.a(3, 2, tf); sink.a(3, 4, x[2] * (tf * tf + 1.0));
        sink.b(2, 0, 1.0); sink.b(4, 1, 1.0);
    }
}

// Sink writing the Euler-discretised dense blocks A (NX x NX) and B (NX x NU), row-major with
// leading dimensions lda/ldb, into memory the caller has pre-set to A = I, B = 0.
struct EulerDenseSink {
    double *A;
    double *B;
    int lda, ldb;
    double dt;
    __device__ __forceinline__ void a(int r, int c, double v) { A[r * lda + c] = (r == c ? 1.0 : 0.0) + dt * v; }
    __device__ __forceinline__ void b(int r, int c, double v) { B[r * ldb + c] = dt * v; }
};

// Runtime dispatch helper: calls fn.template operator()<M>() for the model id.
template <class Fn>
__device__ __forceinline__ void dispatch_model(int model, Fn &&fn)
{
    switch (model) {
    case kDoubleInt4D: fn.template operator()<kDoubleInt4D>(); break;
    case kDoubleInt6D: fn.template operator()<kDoubleInt6D>(); break;
    case kCar3D: fn.template operator()<kCar3D>(); break;
    case kUnicycle4D: fn.template operator()<kUnicycle4D>(); break;
    case kQuad6D: fn.template operator()<kQuad6D>(); break;
    case kHuman6D: fn.template operator()<kHuman6D>(); break;
    case kHumanLin6D: fn.template operator()<kHumanLin6D>(); break;
    case kQuad12D: fn.template operator()<kQuad12D>(); break;
    case kBike5D: fn.template operator()<kBike5D>(); break;
    default: break;
    }
}

// Dispatch inside a size class: a real model id is a compile-time constant, a mixed class switches per agent.
template <int MC, class Fn>
__device__ __forceinline__ void dispatch_class(int model, Fn &&fn)
{
    if constexpr (MC == kMixed4) {
        if (model == kUnicycle4D) fn.template operator()<kUnicycle4D>();
        else fn.template operator()<kDoubleInt4D>();
    } else if constexpr (MC == kMixed6) {
        switch (model) {
        case kQuad6D: fn.template operator()<kQuad6D>(); break;
        case kHuman6D: fn.template operator()<kHuman6D>(); break;
        case kHumanLin6D: fn.template operator()<kHumanLin6D>(); break;
        default: fn.template operator()<kDoubleInt6D>(); break;
        }
    } else {
        fn.template operator()<MC>();
    }
}

}  // namespace dpilqr
