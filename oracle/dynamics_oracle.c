/*
 * dynamics_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU oracle, never shipped).
 *
 * Plain-C restatement of the reference's single-agent dynamics library:
 *   - zero-order-hold RK4 with 5 fixed sub-steps   (reference dpilqr/bbdynamics.cpp:39-93)
 *   - forward-Euler discretisation of the Jacobians (reference dpilqr/bbdynamics.cpp:95-106)
 *   - the eight ODEs + analytic Jacobians           (reference dpilqr/bbdynamics.cpp:108-711)
 *   - the Bike5D ODE + Jacobian and its *single*-step RK4
 *                                                   (reference dpilqr/dynamics.py:18-38,74,254-277)
 *
 * Model numbering follows the reference enum (dpilqr/bbdynamicswrap.pyx:8-16);
 * Bike5D (a SymbolicModel in the reference) is appended as 8.
 *
 * Built with `gcc -O2 -ffp-contract=off` (the reference host build has no FMA
 * contraction).  Pinned against the compiled reference
 * (oracle/_ref) and the golden vectors in tests/golden by tests/test_oracle.py.
 */
#include <math.h>
#include <string.h>

#define ORACLE_MAX_NX 12
#define ORACLE_G 9.80665

enum {
    M_DOUBLE_INT_4D = 0,
    M_DOUBLE_INT_6D = 1,
    M_CAR_3D = 2,
    M_UNICYCLE_4D = 3,
    M_QUAD_6D = 4,
    M_HUMAN_6D = 5,
    M_HUMAN_LIN_6D = 6,
    M_QUAD_12D = 7,
    M_BIKE_5D = 8,
    M_COUNT = 9
};

static const int k_nx[M_COUNT] = {4, 6, 3, 4, 6, 6, 6, 12, 5};
static const int k_nu[M_COUNT] = {2, 3, 2, 2, 3, 3, 3, 4, 2};

/* Quadcopter12D rigid-body constants (reference bbdynamics.cpp:507-510). */
static const double kThrustGain = 2000.0 / 63.0;
static const double kTauX = 625000000000000000.0 / 10982593196059.0;
static const double kTauY = 5000000000000000000.0 / 92848985528431.0;
static const double kTauZ = 10000000000000000000.0 / 271597947137541.0;
static const double kGyroX = 85899976080679.0 / 175721491136944.0;
static const double kGyroY = 95876456000597.0 / 185697971056862.0;
static const double kGyroZ = 9976479919918.0 / 271597947137541.0;

int oracle_model_nx(int model) { return (model >= 0 && model < M_COUNT) ? k_nx[model] : -1; }
int oracle_model_nu(int model) { return (model >= 0 && model < M_COUNT) ? k_nu[model] : -1; }

/* Continuous-time derivative xdot = f(x, u). */
int oracle_f(int model, const double *x, const double *u, double *xd)
{
    switch (model) {
    case M_DOUBLE_INT_4D: /* bbdynamics.cpp:108-117 */
        xd[0] = x[2]; xd[1] = x[3]; xd[2] = u[0]; xd[3] = u[1];
        return 0;
    case M_DOUBLE_INT_6D: /* bbdynamics.cpp:150-161 */
        xd[0] = x[3]; xd[1] = x[4]; xd[2] = x[5];
        xd[3] = u[0]; xd[4] = u[1]; xd[5] = u[2];
        return 0;
    case M_CAR_3D: /* bbdynamics.cpp:230-239 */
        xd[0] = u[0] * cos(x[2]);
        xd[1] = u[0] * sin(x[2]);
        xd[2] = u[1];
        return 0;
    case M_UNICYCLE_4D: /* bbdynamics.cpp:264-274 */
        xd[0] = x[2] * cos(x[3]);
        xd[1] = x[2] * sin(x[3]);
        xd[2] = u[0];
        xd[3] = u[1];
        return 0;
    case M_QUAD_6D: /* bbdynamics.cpp:417-429 */
        xd[0] = x[3]; xd[1] = x[4]; xd[2] = x[5];
        xd[3] = ORACLE_G * tan(u[2]);
        xd[4] = -ORACLE_G * tan(u[1]);
        xd[5] = u[0] - ORACLE_G;
        return 0;
    case M_HUMAN_6D: /* bbdynamics.cpp:308-329: heading is a control */
        xd[0] = x[3] * cos(u[0]);
        xd[1] = x[3] * sin(u[0]);
        xd[2] = 0.0;
        xd[3] = u[1];
        xd[4] = 0.0;
        xd[5] = 0.0;
        return 0;
    case M_HUMAN_LIN_6D: /* bbdynamics.cpp:393-406 */
        xd[0] = x[3]; xd[1] = x[4]; xd[2] = 0.0;
        xd[3] = u[0]; xd[4] = u[1]; xd[5] = 0.0;
        return 0;
    case M_QUAD_12D: { /* bbdynamics.cpp:493-511 */
        const double sy = sin(x[3]), cy = cos(x[3]); /* psi   */
        const double sp = sin(x[4]), cp = cos(x[4]); /* theta */
        const double sr = sin(x[5]), cr = cos(x[5]); /* phi   */
        const double tp = tan(x[4]);
        const double v0 = x[6], v1 = x[7], v2 = x[8];
        const double w0 = x[9], w1 = x[10], w2 = x[11];
        xd[0] = v0 * cy * cp + v1 * (sr * sp * cy - sy * cr) + v2 * (sr * sy + sp * cr * cy);
        xd[1] = v0 * sy * cp + v1 * (sr * sy * sp + cr * cy) + v2 * (-sr * cy + sy * sp * cr);
        xd[2] = -v0 * sp + v1 * sr * cp + v2 * cr * cp;
        xd[3] = w1 * sr / cp + w2 * cr / cp;
        xd[4] = w1 * cr - w2 * sr;
        xd[5] = w0 + w1 * sr * tp + w2 * cr * tp;
        xd[6] = v1 * w2 - v2 * w1 + ORACLE_G * sp;
        xd[7] = -v0 * w2 + v2 * w0 - ORACLE_G * sr * cp;
        xd[8] = kThrustGain * u[3] + v0 * w1 - v1 * w0 - ORACLE_G * cr * cp;
        xd[9] = kTauX * u[0] - kGyroX * w1 * w2;
        xd[10] = kTauY * u[1] + kGyroY * w0 * w2;
        xd[11] = kTauZ * u[2] - kGyroZ * w0 * w1;
        return 0;
    }
    case M_BIKE_5D: /* dynamics.py:258-268 */
        xd[0] = x[2] * cos(x[3]);
        xd[1] = x[2] * sin(x[3]);
        xd[2] = u[0];
        xd[3] = x[2] * tan(x[4]);
        xd[4] = u[1];
        return 0;
    default:
        return -1;
    }
}

/* One classic RK4 step of size h, in place on x (bbdynamics.cpp:62-79 / dynamics.py:30-36). */
static void rk4_step(int model, int nx, double h, double *x, const double *u)
{
    double k0[ORACLE_MAX_NX], k1[ORACLE_MAX_NX], k2[ORACLE_MAX_NX], k3[ORACLE_MAX_NX];
    double xs[ORACLE_MAX_NX];
    int i;
    oracle_f(model, x, u, k0);
    for (i = 0; i < nx; ++i) xs[i] = x[i] + (h / 2.0) * k0[i];
    oracle_f(model, xs, u, k1);
    for (i = 0; i < nx; ++i) xs[i] = x[i] + (h / 2.0) * k1[i];
    oracle_f(model, xs, u, k2);
    for (i = 0; i < nx; ++i) xs[i] = x[i] + h * k2[i];
    oracle_f(model, xs, u, k3);
    for (i = 0; i < nx; ++i) x[i] += h * (k0[i] + 2.0 * k1[i] + 2.0 * k2[i] + k3[i]) / 6.0;
}

/* x_new = Phi_dt(x, u): 5 sub-steps for the native models, 1 step for Bike5D. */
int oracle_integrate(int model, double dt, const double *x, const double *u, double *x_new)
{
    int nx = oracle_model_nx(model), j;
    if (nx < 0) return -1;
    memcpy(x_new, x, sizeof(double) * (size_t)nx);
    if (model == M_BIKE_5D) {
        /* python rk4_integration(f, x, u, dt, dh=dt): while t < h - 1e-8 -> exactly one step
         * (dynamics.py:18-38); there the 0.5*k*step product is formed left to right. */
        double k0[5], k1[5], k2[5], k3[5], xs[5];
        int i;
        oracle_f(model, x_new, u, k0);
        for (i = 0; i < 5; ++i) xs[i] = x_new[i] + 0.5 * k0[i] * dt;
        oracle_f(model, xs, u, k1);
        for (i = 0; i < 5; ++i) xs[i] = x_new[i] + 0.5 * k1[i] * dt;
        oracle_f(model, xs, u, k2);
        for (i = 0; i < 5; ++i) xs[i] = x_new[i] + k2[i] * dt;
        oracle_f(model, xs, u, k3);
        for (i = 0; i < 5; ++i) x_new[i] += dt * (k0[i] + 2.0 * k1[i] + 2.0 * k2[i] + k3[i]) / 6.0;
        return 0;
    }
    for (j = 0; j < 5; ++j) rk4_step(model, nx, dt / 5, x_new, u);
    return 0;
}

/* A = I + dt * df/dx (nx x nx row-major), B = dt * df/du (nx x nu row-major). */
int oracle_linearize(int model, double dt, const double *x, const double *u, double *A, double *B)
{
    int nx = oracle_model_nx(model), nu = oracle_model_nu(model), i;
    if (nx < 0) return -1;
    memset(A, 0, sizeof(double) * (size_t)(nx * nx));
    memset(B, 0, sizeof(double) * (size_t)(nx * nu));
#define AC(r, c) A[(r) * nx + (c)]
#define BC(r, c) B[(r) * nu + (c)]
    switch (model) {
    case M_DOUBLE_INT_4D: /* bbdynamics.cpp:119-148 */
        AC(0, 2) = 1; AC(1, 3) = 1;
        BC(2, 0) = 1; BC(3, 1) = 1;
        break;
    case M_DOUBLE_INT_6D: /* bbdynamics.cpp:163-228 */
        AC(0, 3) = 1; AC(1, 4) = 1; AC(2, 5) = 1;
        BC(3, 0) = 1; BC(4, 1) = 1; BC(5, 2) = 1;
        break;
    case M_CAR_3D: /* bbdynamics.cpp:241-262 */
        AC(0, 2) = -u[0] * sin(x[2]);
        AC(1, 2) = u[0] * cos(x[2]);
        BC(0, 0) = cos(x[2]);
        BC(1, 0) = sin(x[2]);
        BC(2, 1) = 1;
        break;
    case M_UNICYCLE_4D: /* bbdynamics.cpp:276-306 */
        AC(0, 2) = cos(x[3]);
        AC(0, 3) = -x[2] * sin(x[3]);
        AC(1, 2) = sin(x[3]);
        AC(1, 3) = x[2] * cos(x[3]);
        BC(2, 0) = 1; BC(3, 1) = 1;
        break;
    case M_QUAD_6D: /* bbdynamics.cpp:431-491 */
        AC(0, 3) = 1; AC(1, 4) = 1; AC(2, 5) = 1;
        BC(3, 2) = ORACLE_G * pow(tan(u[2]), 2) + ORACLE_G;
        BC(4, 1) = -ORACLE_G * pow(tan(u[1]), 2) - ORACLE_G;
        BC(5, 0) = 1;
        break;
    case M_HUMAN_6D: /* bbdynamics.cpp:331-391 */
        AC(0, 3) = cos(u[0]);
        AC(1, 3) = sin(u[0]);
        BC(0, 0) = -x[3] * sin(u[0]);
        BC(1, 0) = x[3] * cos(u[0]);
        BC(3, 1) = 1;
        break;
    case M_HUMAN_LIN_6D: /* bbdynamics.cpp:408-415: DoubleInt6D with the z channels cut */
        AC(0, 3) = 1; AC(1, 4) = 1;
        BC(3, 0) = 1; BC(4, 1) = 1;
        break;
    case M_QUAD_12D: { /* bbdynamics.cpp:513-711 */
        const double sy = sin(x[3]), cy = cos(x[3]);
        const double sp = sin(x[4]), cp = cos(x[4]);
        const double sr = sin(x[5]), cr = cos(x[5]);
        const double tp = tan(x[4]);
        const double v0 = x[6], v1 = x[7], v2 = x[8];
        const double w0 = x[9], w1 = x[10], w2 = x[11];
        /* world-frame velocity rows */
        AC(0, 3) = -v0 * sy * cp + v1 * (-sr * sy * sp - cr * cy) + v2 * (sr * cy - sy * sp * cr);
        AC(0, 4) = -v0 * sp * cy + v1 * sr * cy * cp + v2 * cr * cy * cp;
        AC(0, 5) = v1 * (sr * sy + sp * cr * cy) + v2 * (-sr * sp * cy + sy * cr);
        AC(0, 6) = cy * cp;
        AC(0, 7) = sr * sp * cy - sy * cr;
        AC(0, 8) = sr * sy + sp * cr * cy;
        AC(1, 3) = v0 * cy * cp + v1 * (sr * sp * cy - sy * cr) + v2 * (sr * sy + sp * cr * cy);
        AC(1, 4) = -v0 * sy * sp + v1 * sr * sy * cp + v2 * sy * cr * cp;
        AC(1, 5) = v1 * (-sr * cy + sy * sp * cr) + v2 * (-sr * sy * sp - cr * cy);
        AC(1, 6) = sy * cp;
        AC(1, 7) = sr * sy * sp + cr * cy;
        AC(1, 8) = -sr * cy + sy * sp * cr;
        AC(2, 4) = -v0 * cp - v1 * sr * sp - v2 * sp * cr;
        AC(2, 5) = v1 * cr * cp - v2 * sr * cp;
        AC(2, 6) = -sp;
        AC(2, 7) = sr * cp;
        AC(2, 8) = cr * cp;
        /* Euler-angle kinematics */
        AC(3, 4) = w1 * sr * sp / pow(cp, 2) + w2 * sp * cr / pow(cp, 2);
        AC(3, 5) = w1 * cr / cp - w2 * sr / cp;
        AC(3, 10) = sr / cp;
        AC(3, 11) = cr / cp;
        AC(4, 5) = -w1 * sr - w2 * cr;
        AC(4, 10) = cr;
        AC(4, 11) = -sr;
        AC(5, 4) = w1 * (pow(tp, 2) + 1) * sr + w2 * (pow(tp, 2) + 1) * cr;
        AC(5, 5) = w1 * cr * tp - w2 * sr * tp;
        AC(5, 9) = 1;
        AC(5, 10) = sr * tp;
        AC(5, 11) = cr * tp;
        /* body-frame translational dynamics */
        AC(6, 4) = ORACLE_G * cp;
        AC(6, 7) = w2;
        AC(6, 8) = -w1;
        AC(6, 10) = -v2;
        AC(6, 11) = v1;
        AC(7, 4) = ORACLE_G * sr * sp;
        AC(7, 5) = -ORACLE_G * cr * cp;
        AC(7, 6) = -w2;
        AC(7, 8) = w0;
        AC(7, 9) = v2;
        AC(7, 11) = -v0;
        AC(8, 4) = ORACLE_G * sp * cr;
        AC(8, 5) = ORACLE_G * sr * cp;
        AC(8, 6) = w1;
        AC(8, 7) = -w0;
        AC(8, 9) = -v1;
        AC(8, 10) = v0;
        /* Euler's rotation equations */
        AC(9, 10) = -kGyroX * w2;
        AC(9, 11) = -kGyroX * w1;
        AC(10, 9) = kGyroY * w2;
        AC(10, 11) = kGyroY * w0;
        AC(11, 9) = -kGyroZ * w1;
        AC(11, 10) = -kGyroZ * w0;
        BC(8, 3) = kThrustGain;
        BC(9, 0) = kTauX;
        BC(10, 1) = kTauY;
        BC(11, 2) = kTauZ;
        break;
    }
    case M_BIKE_5D: /* sympy Jacobians of dynamics.py:258-271 */
        AC(0, 2) = cos(x[3]);
        AC(0, 3) = -x[2] * sin(x[3]);
        AC(1, 2) = sin(x[3]);
        AC(1, 3) = x[2] * cos(x[3]);
        AC(3, 2) = tan(x[4]);
        AC(3, 4) = x[2] * (pow(tan(x[4]), 2) + 1);
        BC(2, 0) = 1; BC(4, 1) = 1;
        break;
    default:
        return -1;
    }
#undef AC
#undef BC
    /* forward-Euler discretisation, bbdynamics.cpp:95-106 / dynamics.py:112-114 */
    for (i = 0; i < nx * nx; ++i) {
        A[i] *= dt;
        if (i % (nx + 1) == 0) A[i] += 1.0;
    }
    for (i = 0; i < nx * nu; ++i) B[i] *= dt;
    return 0;
}

/* Joint (multi-agent) wrappers: agents are concatenated with uniform strides
 * (reference dynamics.py:159-186).  `models` has one entry per agent. */
int oracle_integrate_joint(const int *models, int n_agents, int s, int c, double dt,
                           const double *x, const double *u, double *x_new)
{
    int i, rc = 0;
    for (i = 0; i < n_agents; ++i)
        rc |= oracle_integrate(models[i], dt, x + i * s, u + i * c, x_new + i * s);
    return rc;
}
