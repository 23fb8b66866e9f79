// psi_bdf.cuh — per-thread variable-order BDF/NDF (orders 1-5) and the ESDIRK3(4) tableau: the two
// reference solver names that had no device counterpart (OdeSolver::Bdf — the DEFAULT of every reference
// ODE — and OdeSolver::Sdirk(SdirkTableau::Esdirk34), ode/mod.rs:59-84, 367-447).
//
// The reference delegates to diffsol =0.16.1 (third party, not under /root/reference).  diffsol documents its
// `bdf` as the variable-order NDF scheme of Shampine & Reichelt (ode15s) in the formulation of SciPy's BDF:
// a fixed-leading-coefficient Nordsieck-like array of backward differences D, NDF coefficients kappa, a
// simplified Newton iteration on (I - c J) with a lazily refreshed Jacobian, the order selected every
// order+1 equal steps from the error estimates of order-1 / order / order+1.  That published algorithm
// is what is implemented here (Byrne & Hindmarsh 1975; Shampine & Reichelt 1997; Virtanen et al. 2020);
// parity with the reference is tolerance-based (parity unpinned at diffsol).
//
//   D[0] = y_n, D[j] = j-th backward difference of y (scaled for the current h), j <= order + 2
//   predictor   y0 = sum_{j<=k} D[j]          psi = sum_{j=1..k} gamma_j D[j] / alpha_k
//   corrector   solve  d - c f(t+h, y0 + d) + psi = 0,  c = h / alpha_k   (simplified Newton, <= 4 iterations)
//   error       error_const_k * d
// Stops: every event time is a hard stop (the reference calls set_stop_time per event, ode/mod.rs:741-817):
// the differences are rescaled so the step lands on tstop exactly (SciPy's t_bound handling); the step the
// controller had before the clip is restored at the next call, so observation times do not erode the step.
//
// ESDIRK3(4): 4 stages, explicit first stage, stiffly accurate, L-stable, stage order 2, gamma =
// 0.43586652150845899942 (root of g^3 - 3 g^2 + 3/2 g - 1/6), advancing order 3 with a 4th-order embedded
// solution — the ESDIRK34 of Jørgensen, Kristensen & Thomsen, "A family of ESDIRK integration methods" (2018),
// which diffsol's `esdirk34` cites.  The coefficients below were re-derived from the order conditions
// (stage order 2; sum b = 1, b.c = 1/2, b.c^2 = 1/3; embedded: + b.c^3 = 1/4, b.A.c^2 = 1/12, which pins c3)
// with 40-digit arithmetic (scripts/derive_esdirk34.py) and checked by observed order in tests.
#pragma once
#include "psi_stiff.cuh"

namespace psi {

struct Esdirk34Tab {
    static constexpr int S = 4;
    static constexpr int EST_ORDER = 3;
    __host__ __device__ static constexpr double gamma() { return 0.43586652150845899942; }
    __host__ __device__ static constexpr double c(int s) {
        constexpr double C[4] = {0.0, 0.87173304301691799884, 0.46823874485184439562, 1.0};
        return C[s];
    }
    __host__ __device__ static constexpr double a(int s, int j) {   // strictly lower part; diagonal = gamma (0 for stage 0)
        constexpr double A[4][3] = {{0, 0, 0},
                                    {0.43586652150845899942, 0, 0},
                                    {0.14073777472470619619, -0.10836555138132079998, 0},
                                    {0.10239940061991099768, -0.37687845225555610609, 0.83861253012718610899}};
        return A[s][j];
    }
    __host__ __device__ static constexpr double b(int j) {
        constexpr double B[4] = {0.10239940061991099768, -0.37687845225555610609, 0.83861253012718610899, 0.43586652150845899942};
        return B[j];
    }
    __host__ __device__ static constexpr double e(int j) {         // b - bhat (bhat = the 4th-order weights)
        constexpr double E[4] = {0.10239940061991099768 - 0.15702489786032493710, -0.37687845225555610609 - 0.11733044137043884870,
                                 0.83861253012718610899 - 0.61667803039212146435, 0.43586652150845899942 - 0.10896663037711474985};
        return E[j];
    }
    static constexpr bool EXPLICIT_FIRST = true;
};

template <int N, class F>
PSI_DEV int esdirk34_integrate_to(OdeState<N>& st, double tstop, F& f, const RunOpts& opt, Counters& cnt) {
    return dirk_integrate_to<Esdirk34Tab, N>(st, tstop, f, opt, cnt);
}

// ---------------------------------------------------------------------------------------------
// BDF / NDF, orders 1..5
// ---------------------------------------------------------------------------------------------
constexpr int BDF_MAX_ORDER = 5;
static __constant__ double kBdfKappa[6] = {0.0, -0.1850, -1.0 / 9.0, -0.0823, -0.0415, 0.0};
static __constant__ double kBdfGamma[6] = {0.0, 1.0, 1.5, 11.0 / 6.0, 25.0 / 12.0, 137.0 / 60.0};

// Solver memory that must survive between two integrate-to-stop calls of one occasion.
template <int N>
struct BdfState {
    double D[BDF_MAX_ORDER + 3][N];
    double J[N * N];
    double h_before_clip;      // step size the controller held before the last step was clipped to a stop time (<= 0: none)
    int order;
    int n_equal;
    bool current_jac;
};
// Empty placeholder for the one-step methods.
struct NoSolverMem {};

template <int N>
PSI_DEV double bdf_rms(const double* v, const double* scale_y, double rtol, double atol) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double q = v[i] / fma(rtol, fabs(scale_y[i]), atol);
        s = fma(q, q, s);
    }
    return sqrt(s * (1.0 / N));
}

// Step-size change h -> factor * h:  D[:k+1] <- (R U)^T D[:k+1]  with
//   R[i][j] = prod_{m=1..i} (m - 1 - factor j) / m   (R[0][j] = 1, R[i>=1][0] = 0),   U = R(factor = 1).
// U depends on nothing but the order: it is a constant table (kBdfU, computed once below).  R is never stored: with
// W = R^T D (column j of R is generated by its recurrence while it is consumed) the update is D_new = U^T W, so the
// scratch is (order+1) x N doubles instead of three 6 x 6 matrices — the first version kept 1.8 KB of dynamically indexed
// local memory per thread and ran 65x slower per step than the register-resident Dopri5 (profiles/r02_tuning.md).
// kBdfU[i][j] = prod_{m=1..i} (m - 1 - j) / m = (-1)^i C(j, i): upper triangular (the factor m - 1 - j vanishes at m = j + 1)
static __constant__ double kBdfU[BDF_MAX_ORDER + 1][BDF_MAX_ORDER + 1] = {
    {1, 1, 1, 1, 1, 1},
    {0, -1, -2, -3, -4, -5},
    {0, 0, 1, 3, 6, 10},
    {0, 0, 0, -1, -4, -10},
    {0, 0, 0, 0, 1, 5},
    {0, 0, 0, 0, 0, -1}};
static __constant__ double kBdfInv[BDF_MAX_ORDER + 1] = {0.0, 1.0, 0.5, 1.0 / 3.0, 0.25, 0.2};
// Out of line on purpose: it is called from five places of the step loop and every inlined copy would carry its own scratch.
template <int N>
__device__ __noinline__ void bdf_change_D(BdfState<N>& B, int order, double factor) {
    double W[BDF_MAX_ORDER + 1][N];
    // W[j] = sum_i R[i][j] D[i]
#pragma unroll
    for (int j = 0; j <= BDF_MAX_ORDER; ++j) {
        if (j > order) break;
        double r = 1.0;                                   // R[0][j]
        double acc[N];
#pragma unroll
        for (int q = 0; q < N; ++q) acc[q] = B.D[0][q];
#pragma unroll
        for (int i = 1; i <= BDF_MAX_ORDER; ++i) {
            if (i > order) break;
            r = (j == 0) ? 0.0 : r * (((double)(i - 1) - factor * (double)j) * kBdfInv[i]);
#pragma unroll
            for (int q = 0; q < N; ++q) acc[q] = fma(r, B.D[i][q], acc[q]);
        }
#pragma unroll
        for (int q = 0; q < N; ++q) W[j][q] = acc[q];
    }
    // D[j] = sum_m U[m][j] W[m]
#pragma unroll
    for (int j = 0; j <= BDF_MAX_ORDER; ++j) {
        if (j > order) break;
        double acc[N];
#pragma unroll
        for (int q = 0; q < N; ++q) acc[q] = 0.0;
#pragma unroll
        for (int m = 0; m <= BDF_MAX_ORDER; ++m) {
            if (m > order || m > j) break;                // U[m][j] = 0 for m > j
            const double u = kBdfU[m][j];
#pragma unroll
            for (int q = 0; q < N; ++q) acc[q] = fma(u, W[m][q], acc[q]);
        }
#pragma unroll
        for (int q = 0; q < N; ++q) B.D[j][q] = acc[q];
    }
}

template <int N, class F>
PSI_DEV int bdf_integrate_to(OdeState<N>& st, BdfState<N>& B, double tstop, F& f, const RunOpts& opt, Counters& cnt) {
    constexpr int NEWTON_MAXITER = 4;
    constexpr double MIN_FACTOR = 0.2, MAX_FACTOR = 10.0;
    const double rtol = opt.rtol, atol = opt.atol;
    const double newton_tol = fmax(10.0 * 2.220446049250313e-16 / rtol, fmin(0.03, sqrt(rtol)));
    SmallLU<N> lu;
    double c_lu = 0.0;      // the c of the current factorisation (0: none)
    int iters = 0;
    while (st.t < tstop) {
        // ---- (re)start: after a bolus / infusion boundary / new occasion the history is void ------------------
        if (!st.have_k1) {
            f(st.t, st.y, st.k1);
            cnt.evals++;
            st.have_k1 = true;
            // Hairer's starting step for a first-order method (exponent 1/2), as SciPy's select_initial_step(order = 1)
            double h0;
            {
                const double d0 = bdf_rms<N>(st.y, st.y, rtol, atol), d1 = bdf_rms<N>(st.k1, st.y, rtol, atol);
                h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
                h0 = fmin(h0, tstop - st.t);
                double y1[N], f1[N], df[N];
#pragma unroll
                for (int i = 0; i < N; ++i) y1[i] = fma(h0, st.k1[i], st.y[i]);
                f(st.t + h0, y1, f1);
                cnt.evals++;
#pragma unroll
                for (int i = 0; i < N; ++i) df[i] = f1[i] - st.k1[i];
                const double d2 = bdf_rms<N>(df, st.y, rtol, atol) / h0;
                const double dm = fmax(d1, d2);
                const double h1 = (dm <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : sqrt(0.01 / dm);
                st.h = (opt.h0 > 0.0) ? opt.h0 : fmin(100.0 * h0, h1);
            }
#pragma unroll
            for (int i = 0; i < N; ++i) { B.D[0][i] = st.y[i]; B.D[1][i] = st.k1[i] * st.h; }
            for (int j = 2; j < BDF_MAX_ORDER + 3; ++j)
#pragma unroll
                for (int i = 0; i < N; ++i) B.D[j][i] = 0.0;
            B.order = 1;
            B.n_equal = 0;
            B.h_before_clip = -1.0;
            f.jacobian(st.t, st.y, B.J);
            cnt.evals++;
            B.current_jac = true;
            c_lu = 0.0;
        } else if (B.h_before_clip > st.h) {
            // the previous step was clipped to a stop time: give the controller its step back, at most 10x per attempt
            const double factor = fmin(MAX_FACTOR, B.h_before_clip / st.h);
            bdf_change_D<N>(B, B.order, factor);
            st.h *= factor;
            B.n_equal = 0;
            if (st.h >= B.h_before_clip * (1.0 - 1e-12)) B.h_before_clip = -1.0;
            c_lu = 0.0;
        }
        // ---- one step (SciPy BDF._step_impl) --------------------------------------------------------------------
        const int order = B.order;
        const double alpha = (1.0 - kBdfKappa[order]) * kBdfGamma[order];
        const double error_const = kBdfKappa[order] * kBdfGamma[order] + 1.0 / (double)(order + 1);
        bool accepted = false;
        double d[N], ynew[N], scale_y[N];
        double safety = 0.9, error_norm = 0.0, h = st.h;
        int n_iter = 0;
        while (!accepted) {
            if (++iters > opt.max_steps) return ST_SOLVER_FAILURE;
            h = st.h;
            double t_new = st.t + h;
            const double rem = tstop - st.t;
            // land exactly on the stop: clip, or stretch by <= 1e-9 relative so that rounding never leaves a sliver of a few
            // ulps to be stepped separately (a sliver step used to drag the step size below the failure threshold)
            if (h >= rem * (1.0 - 1e-9)) {
                if (h != rem) {
                    if (h > rem * (1.0 + 1e-9) && B.h_before_clip <= 0.0) B.h_before_clip = st.h;
                    bdf_change_D<N>(B, order, rem / st.h);
                    B.n_equal = 0;
                    c_lu = 0.0;
                }
                t_new = tstop;
                h = rem;
                st.h = h;
            } else if (st.h < 2.3e-15 * fmax(1e-3, fabs(st.t))) {
                // SciPy's min_step = 10 ulp(t): the controller (not a stop time) drove the step to nothing.  Hairer's starting
                // step is legitimately ~1e-13 when a restart finds y ~ atol and |f| large (an infusion switching on at an
                // emptied compartment): that must not count as a failure, the step grows 10x per accepted step from there.
                return ST_SOLVER_FAILURE;
            }
            double ypred[N], psi[N];
#pragma unroll
            for (int i = 0; i < N; ++i) { ypred[i] = 0.0; psi[i] = 0.0; }
            for (int j = 0; j <= order; ++j)
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    ypred[i] += B.D[j][i];
                    if (j >= 1) psi[i] = fma(kBdfGamma[j], B.D[j][i], psi[i]);
                }
            const double ialpha = 1.0 / alpha;
#pragma unroll
            for (int i = 0; i < N; ++i) psi[i] *= ialpha;
            const double c = h * ialpha;
            bool converged = false;
            while (true) {
                if (c_lu != c) {
#pragma unroll
                    for (int i = 0; i < N; ++i)
#pragma unroll
                        for (int j = 0; j < N; ++j) lu.a[i * N + j] = ((i == j) ? 1.0 : 0.0) - c * B.J[i * N + j];
                    lu.factor();
                    c_lu = c;
                }
                // simplified Newton (SciPy solve_bdf_system); the displacement norm is scaled by the predictor
#pragma unroll
                for (int i = 0; i < N; ++i) { d[i] = 0.0; ynew[i] = ypred[i]; }
                double dy_norm_old = -1.0;
                converged = false;
                n_iter = 0;
                if (!lu.singular) {
                    for (int k = 0; k < NEWTON_MAXITER; ++k) {
                        double fv[N];
                        f(t_new, ynew, fv);
                        cnt.evals++; cnt.newton++;
                        n_iter = k + 1;
                        bool finite = true;
#pragma unroll
                        for (int i = 0; i < N; ++i) finite = finite && (fabs(fv[i]) <= 1.7976931348623157e308);
                        if (!finite) break;
                        double dy[N];
#pragma unroll
                        for (int i = 0; i < N; ++i) dy[i] = c * fv[i] - psi[i] - d[i];
                        lu.solve(dy);
                        const double dy_norm = bdf_rms<N>(dy, ypred, rtol, atol);
                        const bool have_rate = dy_norm_old >= 0.0;
                        const double rate = have_rate ? dy_norm / dy_norm_old : 0.0;
                        double rate_pow = rate;                      // rate^(NEWTON_MAXITER - k)
                        for (int e = 1; e < NEWTON_MAXITER - k; ++e) rate_pow *= rate;
                        if (have_rate && (rate >= 1.0 || rate_pow / (1.0 - rate) * dy_norm > newton_tol)) break;
#pragma unroll
                        for (int i = 0; i < N; ++i) { ynew[i] += dy[i]; d[i] += dy[i]; }
                        if (dy_norm == 0.0 || (have_rate && rate / (1.0 - rate) * dy_norm < newton_tol)) { converged = true; break; }
                        dy_norm_old = dy_norm;
                    }
                }
                if (converged || B.current_jac) break;
                f.jacobian(t_new, ypred, B.J);      // stale Jacobian: refresh at the predictor and retry once
                cnt.evals++;
                B.current_jac = true;
                c_lu = 0.0;
            }
            if (!converged) {
                cnt.rejected++;
                st.h *= 0.5;
                bdf_change_D<N>(B, order, 0.5);
                B.n_equal = 0;
                c_lu = 0.0;
                B.h_before_clip = -1.0;
                continue;
            }
            safety = 0.9 * (double)(2 * NEWTON_MAXITER + 1) / (double)(2 * NEWTON_MAXITER + n_iter);
            double err[N];
#pragma unroll
            for (int i = 0; i < N; ++i) { err[i] = error_const * d[i]; scale_y[i] = ynew[i]; }
            error_norm = bdf_rms<N>(err, scale_y, rtol, atol);
            if (!(error_norm <= 1.0)) {
                cnt.rejected++;
                const double factor = (error_norm == error_norm) ? fmax(MIN_FACTOR, safety * (double)__powf((float)error_norm, -1.0f / (float)(order + 1))) : MIN_FACTOR;
                st.h *= factor;
                bdf_change_D<N>(B, order, factor);
                B.n_equal = 0;
                B.h_before_clip = -1.0;
                c_lu = 0.0;
            } else {
                accepted = true;
            }
        }
        // ---- accept -----------------------------------------------------------------------------------------------------
        cnt.steps++;
        B.n_equal++;
        st.t = (h == tstop - st.t) ? tstop : st.t + h;
        B.current_jac = false;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            st.y[i] = ynew[i];
            B.D[order + 2][i] = d[i] - B.D[order + 1][i];
            B.D[order + 1][i] = d[i];
        }
        for (int j = order; j >= 0; --j)
#pragma unroll
            for (int i = 0; i < N; ++i) B.D[j][i] += B.D[j + 1][i];
        if (B.n_equal < order + 1) continue;
        // ---- order / step selection every order + 1 equal steps ----------------------------------------------------
        double f_m = 0.0, f_k, f_p = 0.0;      // candidate factors for order-1, order, order+1 (error_norm^(-1/(q+1)))
        f_k = (error_norm > 1e-30) ? (double)__powf((float)error_norm, -1.0f / (float)(order + 1)) : 1e30;     // FP32 on the SFU: it only steers h
        if (order > 1) {
            double e[N];
            const double ec = kBdfKappa[order - 1] * kBdfGamma[order - 1] + 1.0 / (double)order;
#pragma unroll
            for (int i = 0; i < N; ++i) e[i] = ec * B.D[order][i];
            const double en = bdf_rms<N>(e, scale_y, rtol, atol);
            f_m = (en > 1e-30) ? (double)__powf((float)en, -1.0f / (float)order) : 1e30;
        }
        if (order < BDF_MAX_ORDER) {
            double e[N];
            const double ec = kBdfKappa[order + 1] * kBdfGamma[order + 1] + 1.0 / (double)(order + 2);
#pragma unroll
            for (int i = 0; i < N; ++i) e[i] = ec * B.D[order + 2][i];
            const double en = bdf_rms<N>(e, scale_y, rtol, atol);
            f_p = (en > 1e-30) ? (double)__powf((float)en, -1.0f / (float)(order + 2)) : 1e30;
        }
        // numpy argmax over [order-1, order, order+1]: the first maximum wins (a missing neighbour has factor 0)
        int new_order = order - 1;
        double best = f_m;
        if (f_k > best) { best = f_k; new_order = order; }
        if (f_p > best) { best = f_p; new_order = order + 1; }
        if (new_order < 1) new_order = 1;
        const double factor = fmin(MAX_FACTOR, safety * best);
        B.order = new_order;
        // a step that was clipped to the stop time says nothing about the step the controller wants next
        if (B.h_before_clip > 0.0) B.h_before_clip = fmax(B.h_before_clip, st.h * factor);
        st.h *= factor;
        bdf_change_D<N>(B, new_order, factor);
        B.n_equal = 0;
        c_lu = 0.0;
    }
    return ST_OK;
}

}  // namespace psi
