// psi_ode.cuh — per-thread adaptive ODE integrators (register-resident state, FP64 scalar).
//
// Explicit pairs: Dormand-Prince 5(4) (named by the north star) and Tsitouras 5(4) (the
// reference's own explicit solver: OdeSolver::ExplicitRk(Tsit45), ode/mod.rs:59-84, 382-400).
// Stiff: see psi_stiff.cuh.
//
// The reference delegates stepping to diffsol =0.16.1 (third party, not under /root/reference);
// these integrators implement the published methods, so parity with the reference's CPU solvers
// is tolerance-based (both sides converge on the same solution), never step-for-step.
// Step-size control: error norm = RMS of e_i / (atol + rtol*max(|y_i|,|ynew_i|)), I-controller
// fac = 0.9 * err^(-1/5) clamped to [0.2, 10]; the factor is evaluated in FP32 on the SFU
// (MUFU.LG2/EX2) because it only steers h and must not occupy the FP64 pipe.
#pragma once
#include "psi_common.cuh"

namespace psi {

// The tableau coefficients live in the constant bank so DFMA reads them as c[3][..] operands; as
// constexpr literals the compiler re-materialises each 64-bit immediate with two UMOVs per use
// (88 UMOV per Dopri5 step in the first build).  The constexpr copies stay for compile-time zero tests.
#define PSI_DP5_A \
    {0, 0, 0, 0, 0, 0}, {1.0 / 5, 0, 0, 0, 0, 0}, {3.0 / 40, 9.0 / 40, 0, 0, 0, 0}, {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0, 0}, \
    {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0, 0}, \
    {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656, 0}, \
    {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84}
#define PSI_DP5_E \
    35.0 / 384 - 5179.0 / 57600, 0.0, 500.0 / 1113 - 7571.0 / 16695, 125.0 / 192 - 393.0 / 640, -2187.0 / 6784 + 92097.0 / 339200, \
    11.0 / 84 - 187.0 / 2100, -1.0 / 40
#define PSI_TS5_A \
    {0, 0, 0, 0, 0, 0}, {0.161, 0, 0, 0, 0, 0}, {-0.008480655492356989, 0.335480655492357, 0, 0, 0, 0}, \
    {2.8971530571054935, -6.359448489975075, 4.3622954328695815, 0, 0, 0}, \
    {5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525, 0, 0}, \
    {5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383, 0}, \
    {0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774}
#define PSI_TS5_E \
    -0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995, -0.1447110071732629, 0.5823571654525552, \
    -0.45808210592918697, 0.015151515151515152
static __constant__ double kDp5A[7][6] = {PSI_DP5_A};
static __constant__ double kDp5E[7] = {PSI_DP5_E};
static __constant__ double kTs5A[7][6] = {PSI_TS5_A};
static __constant__ double kTs5E[7] = {PSI_TS5_E};

struct Dopri5 {
    static constexpr int S = 7;
    __host__ __device__ static constexpr double c(int s) {
        constexpr double C[7] = {0.0, 1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1.0, 1.0};
        return C[s];
    }
    __host__ __device__ static constexpr double a(int s, int j) {
        constexpr double A[7][6] = {PSI_DP5_A};
        return A[s][j];
    }
    PSI_DEV static double ca(int s, int j) { return kDp5A[s][j]; }
    PSI_DEV static double ce(int j) { return kDp5E[j]; }
    // b - bhat
    __host__ __device__ static constexpr double e(int j) {
        constexpr double E[7] = {PSI_DP5_E};
        return E[j];
    }
};

struct Tsit5 {
    static constexpr int S = 7;
    __host__ __device__ static constexpr double c(int s) {
        constexpr double C[7] = {0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0};
        return C[s];
    }
    __host__ __device__ static constexpr double a(int s, int j) {
        constexpr double A[7][6] = {PSI_TS5_A};
        return A[s][j];
    }
    PSI_DEV static double ca(int s, int j) { return kTs5A[s][j]; }
    PSI_DEV static double ce(int j) { return kTs5E[j]; }
    __host__ __device__ static constexpr double e(int j) {
        constexpr double E[7] = {PSI_TS5_E};
        return E[j];
    }
};

template <int N>
struct OdeState {
    double t;
    double y[N];
    double h;          // next trial step; <= 0 means "pick automatically" (restart)
    double k1[N];      // f(t, y) when have_k1 (FSAL)
    bool have_k1;
    // Restart memory: the step size the controller had settled on shortly after the previous restart
    // (periodic dosing => the same transient recurs), tried first at the next restart instead of
    // ramping up again from the Hairer starting step.  <= 0: none yet.
    double h_post;
    int since_restart;
};
#ifndef PSI_RESTART_REUSE
#define PSI_RESTART_REUSE 1
#endif

template <int N>
PSI_DEV double rms_scaled(const double* v, const double* y, double rtol, double atol) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double q = v[i] * rcp_approx(fma(rtol, fabs(y[i]), atol));
        s = fma(q, q, s);
    }
    return sqrt(s * (1.0 / N));
}

// Hairer-Norsett-Wanner II.4 starting step (one extra RHS evaluation).
template <int N, class F>
PSI_DEV double initial_step(F& f, double t, const double* y, const double* f0, double span, double rtol, double atol,
                            Counters& cnt) {
    const double d0 = rms_scaled<N>(y, y, rtol, atol);
    const double d1 = rms_scaled<N>(f0, y, rtol, atol);
    double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
    h0 = fmin(h0, span);
    double y1[N], f1[N], df[N];
#pragma unroll
    for (int i = 0; i < N; ++i) y1[i] = y[i] + h0 * f0[i];
    f(t + h0, y1, f1);
    cnt.evals++;
#pragma unroll
    for (int i = 0; i < N; ++i) df[i] = f1[i] - f0[i];
    const double d2 = rms_scaled<N>(df, y, rtol, atol) / h0;
    const double dm = fmax(d1, d2);
    const double h1 = (dm <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : (double)__powf((float)(0.01 / dm), 0.2f);
    return fmin(fmin(100.0 * h0, h1), span);
}

// Integrate from st.t to exactly tstop.  Returns ST_OK or ST_SOLVER_FAILURE.
template <class TAB, int N, class F>
PSI_DEV int erk_integrate_to(OdeState<N>& st, double tstop, F& f, const RunOpts& opt, Counters& cnt) {
    const double rtol = opt.rtol, atol = opt.atol;
    double k[TAB::S][N];
    int iters = 0;
    if (!(st.t < tstop)) return ST_OK;
    // Restart work (first slope, first step size) happens at most once per call and only here: inside the step loop both
    // conditions are invariantly false (k1 is FSAL-carried or kept on a rejection, h stays > 0), and a test left there gets
    // if-converted into a predicated right-hand side that every step issues (ncu: 7 evaluations per step instead of 6).
    if (!st.have_k1) {
        f(st.t, st.y, st.k1);
        cnt.evals++;
        st.have_k1 = true;
    }
    if (!(st.h > 0.0)) {
        st.since_restart = 0;
        if (PSI_RESTART_REUSE && st.h_post > 0.0) st.h = st.h_post;
        else st.h = (opt.h0 > 0.0) ? opt.h0 : initial_step<N>(f, st.t, st.y, st.k1, tstop - st.t, rtol, atol, cnt);
    }
    while (st.t < tstop) {
        if (++iters > opt.max_steps) return ST_SOLVER_FAILURE;
        const double rem = tstop - st.t;
        const bool last = st.h >= rem;
        const double h = last ? rem : st.h;
#pragma unroll
        for (int i = 0; i < N; ++i) k[0][i] = st.k1[i];
        double ynew[N];
#pragma unroll
        for (int s = 1; s < TAB::S; ++s) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < s; ++j)
                    if (TAB::a(s, j) != 0.0) acc = fma(TAB::ca(s, j), k[j][i], acc);
                ynew[i] = fma(h, acc, st.y[i]);
            }
            f(st.t + TAB::c(s) * h, ynew, k[s]);
        }
        cnt.evals += TAB::S - 1;
        // ynew = 5th-order solution (FSAL: stage S-1 is f(t+h, ynew)).
        // Error norm: the embedded difference sum stays in FP64 (it is a cancellation of O(1) terms down
        // to ~tol); the per-component weight 1/sc uses the one-MUFU reciprocal and the controller runs
        // in FP32 on the SFU: none of this steers anything but h.
        double err2 = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double e = 0.0;
#pragma unroll
            for (int j = 0; j < TAB::S; ++j)
                if (TAB::e(j) != 0.0) e = fma(TAB::ce(j), k[j][i], e);
            const double sc = fma(rtol, max_abs(st.y[i], ynew[i]), atol);
            const double q = (h * e) * rcp_approx(sc);
            err2 = fma(q, q, err2);
        }
        if (!(err2 <= 1e300)) {
            // non-finite error estimate: treat as a rejected step with the maximum shrink
            cnt.rejected++;
            st.h = h * 0.2;
            if (st.h < 1e-14 * fmax(1.0, fabs(st.t))) return ST_SOLVER_FAILURE;
            continue;
        }
        // fac = 0.9 * err^(-1/5), err = sqrt(err2 / N)  =>  0.9 * (err2 / N)^(-1/10)
        const float e2 = (float)err2 * (1.0f / N);
        float fac = (e2 <= 1e-30f) ? 10.0f : 0.9f * powf_fast(e2, -0.1f);
        fac = fminf(10.0f, fmaxf(0.2f, fac));
        if (err2 <= (double)N) {
            cnt.steps++;
            st.t = last ? tstop : st.t + h;
#pragma unroll
            for (int i = 0; i < N; ++i) { st.y[i] = ynew[i]; st.k1[i] = k[TAB::S - 1][i]; }
            // a step clipped by tstop must not shrink the controller's step estimate
            const double hn = h * (double)fac;
            st.h = (last && hn < st.h) ? st.h : hn;
            if (PSI_RESTART_REUSE && ++st.since_restart == 2) st.h_post = st.h;
        } else {
            cnt.rejected++;
            st.h = h * (double)fminf(1.0f, fac);
            if (st.h < 1e-14 * fmax(1.0, fabs(st.t))) return ST_SOLVER_FAILURE;
        }
    }
    return ST_OK;
}

}  // namespace psi
