// psi_stiff.cuh — per-thread implicit integrators for stiff models, with a register-resident
// dense LU (N <= ~8, fully unrolled, partial pivoting by conditional row swaps so no register
// array is ever indexed dynamically) and an analytic Jacobian generated from the DSL IR.
//
//   SDIRK4   Hairer & Wanner, "Solving ODEs II", Table IV.6.5: 5 stages, gamma = 1/4, order 4,
//            L-stable, embedded order 3.
//   TR-BDF2  Bank et al. 1985 as a 3-stage ESDIRK (gamma = 2 - sqrt 2) with the Hosea-Shampine
//            1996 third-order error estimate — one of the reference's own solver choices
//            (OdeSolver::Sdirk(SdirkTableau::TrBdf2), ode/mod.rs:59-84; diffsol `tr_bdf2`).
//
// Stage equations are solved for the stage derivatives K_i by simplified Newton with one
// factorisation of (I - h*gamma*J) per step; the error estimate is filtered through the same
// factorisation (Shampine) so it stays meaningful on stiff components.
// The reference's default stiff solver is diffsol's BDF (third party, not under
// /root/reference); parity is tolerance-based against SciPy Radau goldens.
#pragma once
#include "psi_ode.cuh"

namespace psi {

// LU of an N x N matrix held in registers.  swaps[k] bit i: rows k and i swapped at column k (one mask per
// column, statically indexed under full unrolling, so any N <= 32 is correct).
template <int N>
struct SmallLU {
    static_assert(N >= 1 && N <= 32, "SmallLU keeps one 32-bit swap mask per column");
    double a[N * N];
    unsigned int swaps[N];
    bool singular;

    PSI_DEV void factor() {
#pragma unroll
        for (int k = 0; k < N; ++k) swaps[k] = 0u;
        singular = false;
#pragma unroll
        for (int k = 0; k < N; ++k) {
            // bring the largest |a_ik|, i >= k, into row k by a chain of conditional swaps
#pragma unroll
            for (int i = k + 1; i < N; ++i) {
                const bool sw = fabs(a[i * N + k]) > fabs(a[k * N + k]);
                swaps[k] |= sw ? (1u << i) : 0u;
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    const double u = a[k * N + j], v = a[i * N + j];
                    a[k * N + j] = sw ? v : u;
                    a[i * N + j] = sw ? u : v;
                }
            }
            const double piv = a[k * N + k];
            if (piv == 0.0 || piv != piv) singular = true;
            const double ip = rcp_nr(piv);      // ~1 ulp, no IEEE fix-up path (psi_common.cuh)
            a[k * N + k] = ip;                 // store the reciprocal pivot
#pragma unroll
            for (int i = k + 1; i < N; ++i) {
                const double l = a[i * N + k] * ip;
                a[i * N + k] = l;
#pragma unroll
                for (int j = k + 1; j < N; ++j) a[i * N + j] = fma(-l, a[k * N + j], a[i * N + j]);
            }
        }
    }
    PSI_DEV void solve(double* b) const {
#pragma unroll
        for (int k = 0; k < N; ++k) {
#pragma unroll
            for (int i = k + 1; i < N; ++i) {
                const bool sw = (swaps[k] >> i) & 1u;
                const double u = b[k], v = b[i];
                b[k] = sw ? v : u;
                b[i] = sw ? u : v;
            }
#pragma unroll
            for (int i = k + 1; i < N; ++i) b[i] = fma(-a[i * N + k], b[k], b[i]);
        }
#pragma unroll
        for (int k = N - 1; k >= 0; --k) {
            double s = b[k];
#pragma unroll
            for (int j = k + 1; j < N; ++j) s = fma(-a[k * N + j], b[j], s);
            b[k] = s * a[k * N + k];
        }
    }
};

// N = 2 (the effect-compartment PD models of the C4 kind): the explicit inverse by the adjugate instead of pivoted
// elimination.  The conditional row swaps of the generic LU are 64-bit selects in `factor` and in every one of the six
// `solve` calls of a RODAS4 step — ncu on C4: FSEL was 12 % of the executed instructions — while a 2 x 2 inverse is one
// determinant (an FMA pair), one reciprocal and four products, and a solve is two FMA pairs with no data-dependent
// permutation.  E = I/(h gamma) - J is diagonally dominated by 1/(h gamma) for the steps the controller accepts; a
// vanishing or non-finite determinant is reported as singular exactly like a zero pivot.
template <>
struct SmallLU<2> {
    double a[4];          // in: the matrix (row-major); after factor(): its inverse
    bool singular;
    PSI_DEV void factor() {
        const double det = fma(a[0], a[3], -(a[1] * a[2]));
        singular = (det == 0.0) || (det != det);
        const double id = rcp_nr(det);
        const double i00 = a[3] * id, i01 = -a[1] * id, i10 = -a[2] * id, i11 = a[0] * id;
        a[0] = i00; a[1] = i01; a[2] = i10; a[3] = i11;
    }
    PSI_DEV void solve(double* b) const {
        const double b0 = b[0], b1 = b[1];
        b[0] = fma(a[0], b0, a[1] * b1);
        b[1] = fma(a[2], b0, a[3] * b1);
    }
};

struct Sdirk4Tab {
    static constexpr int S = 5;
    static constexpr int EST_ORDER = 3;
    __host__ __device__ static constexpr double gamma() { return 0.25; }
    __host__ __device__ static constexpr double c(int s) {
        constexpr double C[5] = {1.0 / 4, 3.0 / 4, 11.0 / 20, 1.0 / 2, 1.0};
        return C[s];
    }
    __host__ __device__ static constexpr double a(int s, int j) {   // strictly lower part; diagonal = gamma
        constexpr double A[5][4] = {{0, 0, 0, 0},
                                    {1.0 / 2, 0, 0, 0},
                                    {17.0 / 50, -1.0 / 25, 0, 0},
                                    {371.0 / 1360, -137.0 / 2720, 15.0 / 544, 0},
                                    {25.0 / 24, -49.0 / 48, 125.0 / 16, -85.0 / 12}};
        return A[s][j];
    }
    __host__ __device__ static constexpr double b(int j) {
        constexpr double B[5] = {25.0 / 24, -49.0 / 48, 125.0 / 16, -85.0 / 12, 1.0 / 4};
        return B[j];
    }
    __host__ __device__ static constexpr double e(int j) {         // b - bhat
        constexpr double E[5] = {25.0 / 24 - 59.0 / 48, -49.0 / 48 + 17.0 / 96, 125.0 / 16 - 225.0 / 32, -85.0 / 12 + 85.0 / 12, 1.0 / 4};
        return E[j];
    }
    static constexpr bool EXPLICIT_FIRST = false;
};

struct TrBdf2Tab {
    static constexpr int S = 3;
    static constexpr int EST_ORDER = 2;   // controller exponent 1/(2+1)
    // gamma_TR = 2 - sqrt2; diagonal d = gamma_TR/2; w = sqrt2/4
    __host__ __device__ static constexpr double gamma() { return 0.29289321881345248; }   // d = 1 - sqrt2/2
    __host__ __device__ static constexpr double c(int s) {
        constexpr double C[3] = {0.0, 0.58578643762690495, 1.0};
        return C[s];
    }
    __host__ __device__ static constexpr double a(int s, int j) {
        constexpr double d = 0.29289321881345248, w = 0.35355339059327376;
        constexpr double A[3][2] = {{0, 0}, {d, 0}, {w, w}};
        return A[s][j];
    }
    __host__ __device__ static constexpr double b(int j) {
        constexpr double B[3] = {0.35355339059327376, 0.35355339059327376, 0.29289321881345248};
        return B[j];
    }
    __host__ __device__ static constexpr double e(int j) {
        // b - bhat, bhat = ((1-w)/3, (3w+1)/3, d/3)
        constexpr double d = 0.29289321881345248, w = 0.35355339059327376;
        constexpr double E[3] = {w - (1.0 - w) / 3.0, w - (3.0 * w + 1.0) / 3.0, d - d / 3.0};
        return E[j];
    }
    static constexpr bool EXPLICIT_FIRST = true;   // first stage is f(t, y) (ESDIRK)
};

template <class TAB, int N, class F>
PSI_DEV int dirk_integrate_to(OdeState<N>& st, double tstop, F& f, const RunOpts& opt, Counters& cnt) {
    const double rtol = opt.rtol, atol = opt.atol;
    constexpr double g = TAB::gamma();
    double K[TAB::S][N];
    SmallLU<N> lu;
    double J[N * N];
    bool have_jac = false;
    double h_lu = 0.0;
    int iters = 0;
    while (st.t < tstop) {
        if (++iters > opt.max_steps) return ST_SOLVER_FAILURE;
        if (!st.have_k1) { f(st.t, st.y, st.k1); cnt.evals++; st.have_k1 = true; }
        if (!(st.h > 0.0)) st.h = (opt.h0 > 0.0) ? opt.h0 : initial_step<N>(f, st.t, st.y, st.k1, tstop - st.t, rtol, atol, cnt);
        const double rem = tstop - st.t;
        const bool last = st.h >= rem;
        const double h = last ? rem : st.h;
        if (!have_jac) { f.jacobian(st.t, st.y, J); have_jac = true; h_lu = 0.0; cnt.evals++; }
        if (h != h_lu) {
#pragma unroll
            for (int i = 0; i < N; ++i)
#pragma unroll
                for (int j = 0; j < N; ++j) lu.a[i * N + j] = ((i == j) ? 1.0 : 0.0) - h * g * J[i * N + j];
            lu.factor();
            h_lu = h;
        }
        bool ok = !lu.singular;
        double ys[N];
        // ---- stages ------------------------------------------------------------------------------
#pragma unroll
        for (int s = 0; s < TAB::S; ++s) {
            if (!ok) break;
            if (TAB::EXPLICIT_FIRST && s == 0) {
#pragma unroll
                for (int i = 0; i < N; ++i) K[0][i] = st.k1[i];
                continue;
            }
            double base[N];
#pragma unroll
            for (int i = 0; i < N; ++i) {
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < s; ++j)
                    if (TAB::a(s, j) != 0.0) acc = fma(TAB::a(s, j), K[j][i], acc);
                base[i] = fma(h, acc, st.y[i]);
                K[s][i] = (s == 0) ? st.k1[i] : K[s - 1][i];       // predictor
            }
            const double ts = st.t + TAB::c(s) * h;
            double prev_norm = 1e300;
            bool conv = false;
#pragma unroll 1
            for (int it = 0; it < 8; ++it) {
                double res[N];
#pragma unroll
                for (int i = 0; i < N; ++i) ys[i] = fma(h * g, K[s][i], base[i]);
                f(ts, ys, res);
                cnt.evals++; cnt.newton++;
#pragma unroll
                for (int i = 0; i < N; ++i) res[i] = res[i] - K[s][i];    // -G(K)
                lu.solve(res);
                double nrm2 = 0.0;
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    K[s][i] += res[i];
                    const double q = (h * res[i]) / (atol + rtol * fabs(st.y[i]));
                    nrm2 = fma(q, q, nrm2);
                }
                const double nrm = sqrt(nrm2 * (1.0 / N));
                if (!(nrm == nrm)) break;
                if (nrm < 0.03) { conv = true; break; }
                if (it > 0 && nrm > 2.0 * prev_norm) break;          // diverging
                prev_norm = nrm;
            }
            if (!conv) ok = false;
        }
        if (!ok) {
            cnt.rejected++;
            st.h = h * 0.25;
            have_jac = false;                                         // refresh J at the same point
            if (st.h < 1e-14 * fmax(1.0, fabs(st.t))) return ST_SOLVER_FAILURE;
            continue;
        }
        // ---- solution + filtered error estimate ---------------------------------------------------
        double ynew[N], ev[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double acc = 0.0, e = 0.0;
#pragma unroll
            for (int j = 0; j < TAB::S; ++j) {
                acc = fma(TAB::b(j), K[j][i], acc);
                if (TAB::e(j) != 0.0) e = fma(TAB::e(j), K[j][i], e);
            }
            ynew[i] = fma(h, acc, st.y[i]);
            ev[i] = h * e;
        }
        lu.solve(ev);
        double err2 = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double q = ev[i] / (atol + rtol * fmax(fabs(st.y[i]), fabs(ynew[i])));
            err2 = fma(q, q, err2);
        }
        const double err = sqrt(err2 * (1.0 / N));
        if (!(err == err)) {
            cnt.rejected++; st.h = h * 0.25; have_jac = false;
            if (st.h < 1e-14 * fmax(1.0, fabs(st.t))) return ST_SOLVER_FAILURE;
            continue;
        }
        float fac = (err <= 1e-30) ? 8.0f : 0.9f * __powf((float)err, -1.0f / (float)(TAB::EST_ORDER + 1));
        fac = fminf(8.0f, fmaxf(0.2f, fac));
        if (err <= 1.0) {
            cnt.steps++;
            st.t = last ? tstop : st.t + h;
#pragma unroll
            for (int i = 0; i < N; ++i) st.y[i] = ynew[i];
            // both tableaux are stiffly accurate with c_S = 1: K_S = f(t+h, ynew)
#pragma unroll
            for (int i = 0; i < N; ++i) st.k1[i] = K[TAB::S - 1][i];
            have_jac = false;
            // avoid refactorising for marginal changes of h
            if (fac > 1.0f && fac < 1.2f) fac = 1.0f;
            const double hn = h * (double)fac;
            st.h = (last && hn < st.h) ? st.h : hn;
        } else {
            cnt.rejected++;
            st.h = h * (double)fminf(1.0f, fac);
            if (st.h < 1e-14 * fmax(1.0, fabs(st.t))) return ST_SOLVER_FAILURE;
        }
    }
    return ST_OK;
}

template <int N, class F>
PSI_DEV int sdirk4_integrate_to(OdeState<N>& st, double tstop, F& f, const RunOpts& opt, Counters& cnt) {
    return dirk_integrate_to<Sdirk4Tab, N>(st, tstop, f, opt, cnt);
}
template <int N, class F>
PSI_DEV int trbdf2_integrate_to(OdeState<N>& st, double tstop, F& f, const RunOpts& opt, Counters& cnt) {
    return dirk_integrate_to<TrBdf2Tab, N>(st, tstop, f, opt, cnt);
}

// ---------------------------------------------------------------------------------------------
// RODAS4 (Hairer & Wanner, "Solving ODEs II", IV.7; coefficients of the authors' RODAS code, METH=1):
// 6-stage stiffly accurate Rosenbrock method of order 4 with an embedded order-3 solution.
// Linearly implicit: per step ONE Jacobian, ONE LU of (I/(h gamma) - J), six right-hand sides and six
// triangular solves, and no Newton iteration, so every lane of a warp does the same fixed work per step
// (the SDIRK methods above diverge on their per-stage Newton counts).  The coefficients were
// cross-checked by convergence order (4.00 observed) against SciPy Radau before use.
//   (I/(h g) - J) k_i = f(t + alpha_i h, y + sum_j a_ij k_j) + (1/h) sum_j c_ij k_j + h d_i df/dt
//   y_new = y + sum a_5j k_j + k_5 + k_6,   error estimate = k_6
// df/dt is only formed (one extra RHS evaluation) for right-hand sides that depend on time explicitly.
// ---------------------------------------------------------------------------------------------
static __constant__ double kRodasA[6][5] = {
    {0, 0, 0, 0, 0},
    {0.1544000000000000e+01, 0, 0, 0, 0},
    {0.9466785280815826e+00, 0.2557011698983284e+00, 0, 0, 0},
    {0.3314825187068521e+01, 0.2896124015972201e+01, 0.9986419139977817e+00, 0, 0},
    {0.1221224509226641e+01, 0.6019134481288629e+01, 0.1253708332932087e+02, -0.6878860361058950e+00, 0},
    {0, 0, 0, 0, 0}};   // stage 6 continues from stage 5: u += k_5
static __constant__ double kRodasC[6][5] = {
    {0, 0, 0, 0, 0},
    {-0.5668800000000000e+01, 0, 0, 0, 0},
    {-0.2430093356833875e+01, -0.2063599157091915e+00, 0, 0, 0},
    {-0.1073529058151375e+00, -0.9594562251023355e+01, -0.2047028614809616e+02, 0, 0},
    {0.7496443313967647e+01, -0.1024680431464352e+02, -0.3399990352819905e+02, 0.1170890893206160e+02, 0},
    {0.8083246795921522e+01, -0.7981132988064893e+01, -0.3152159432874371e+02, 0.1631930543123136e+02, -0.6058818238834054e+01}};
static __constant__ double kRodasAlpha[6] = {0.0, 0.386, 0.21, 0.63, 1.0, 1.0};
static __constant__ double kRodasD[6] = {0.25, -0.1043, 0.1035, -0.3620000000000023e-01, 0.0, 0.0};

template <int N, bool TIME_DEP, class F>
PSI_DEV int rodas4_integrate_to(OdeState<N>& st, double tstop, F& f, const RunOpts& opt, Counters& cnt) {
    const double rtol = opt.rtol, atol = opt.atol;
    constexpr double g = 0.25;
    double K[6][N];
    SmallLU<N> lu;
    [[maybe_unused]] double T[N];
    int iters = 0;
    while (st.t < tstop) {
        if (++iters > opt.max_steps) return ST_SOLVER_FAILURE;
        if (!st.have_k1) {
            if (TIME_DEP || !(st.h > 0.0)) f(st.t, st.y, st.k1);      // (the first step size needs it; else see below)
            cnt.evals++;
            st.have_k1 = true;
            if constexpr (TIME_DEP) {       // df/dt at (t_n, y_n) by a forward difference
                const double dlt = 1.4901161193847656e-08 * fmax(1e-5, fabs(st.t)) + 1e-10;
                f(st.t + dlt, st.y, T);
                cnt.evals++;
                const double idl = 1.0 / dlt;
#pragma unroll
                for (int i = 0; i < N; ++i) T[i] = (T[i] - st.k1[i]) * idl;
            }
        }
        if (!(st.h > 0.0)) {
            st.since_restart = 0;
            if (PSI_RESTART_REUSE && st.h_post > 0.0) st.h = st.h_post;
            else st.h = (opt.h0 > 0.0) ? opt.h0 : initial_step<N>(f, st.t, st.y, st.k1, tstop - st.t, rtol, atol, cnt);
        }
        const double rem = tstop - st.t;
        const bool last = st.h >= rem;
        const double h = last ? rem : st.h;
        const double ih = rcp_nr(h);
        const double ihg = ih * (1.0 / g);
        // E = I/(h g) - J, factored in registers (J is re-evaluated instead of stored: a Jacobian costs
        // about one RHS, N*N registers cost occupancy)
        // f(t_n, y_n) is evaluated next to the Jacobian, also on the attempt after a rejection (1 in 10) where it is already
        // known: the two share their sub-expressions (the derive block, the Michaelis-Menten reciprocal), which a slope
        // computed behind the `have_k1` test above cannot
        if constexpr (!TIME_DEP) f(st.t, st.y, st.k1);
        f.jacobian(st.t, st.y, lu.a);
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j) lu.a[i * N + j] = ((i == j) ? ihg : 0.0) - lu.a[i * N + j];
        lu.factor();
        bool bad = lu.singular;
        double u[N];
#pragma unroll
        for (int s = 0; s < 6; ++s) {
            double rhs[N];
            if (s == 0) {
#pragma unroll
                for (int i = 0; i < N; ++i) rhs[i] = st.k1[i];
            } else {
                if (s < 5) {
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        double acc = st.y[i];
#pragma unroll
                        for (int j = 0; j < s; ++j) acc = fma(kRodasA[s][j], K[j][i], acc);
                        u[i] = acc;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < N; ++i) u[i] += K[4][i];
                }
                f(st.t + kRodasAlpha[s] * h, u, rhs);
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    double acc = 0.0;
#pragma unroll
                    for (int j = 0; j < s; ++j) acc = fma(kRodasC[s][j], K[j][i], acc);
                    rhs[i] = fma(ih, acc, rhs[i]);
                }
            }
            if constexpr (TIME_DEP) {
                if (s < 4) {
#pragma unroll
                    for (int i = 0; i < N; ++i) rhs[i] = fma(h * kRodasD[s], T[i], rhs[i]);
                }
            }
            lu.solve(rhs);
#pragma unroll
            for (int i = 0; i < N; ++i) K[s][i] = rhs[i];
        }
        cnt.evals += 6;        // 5 right-hand sides + the Jacobian
        double err2 = 0.0;
        double ynew[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            ynew[i] = u[i] + K[5][i];
            const double sc = fma(rtol, max_abs(st.y[i], ynew[i]), atol);
            const double q = K[5][i] * rcp_approx(sc);
            err2 = fma(q, q, err2);
        }
        if (bad || !(err2 <= 1e300)) {
            cnt.rejected++;
            st.h = h * 0.25;
            if (st.h < 1e-14 * fmax(1.0, fabs(st.t))) return ST_SOLVER_FAILURE;
            continue;
        }
        // fac = 0.9 * err^(-1/4), err = sqrt(err2 / N)
        const float e2 = (float)err2 * (1.0f / N);
        float fac = (e2 <= 1e-30f) ? 6.0f : 0.9f * powf_fast(e2, -0.125f);
        fac = fminf(6.0f, fmaxf(0.2f, fac));
        if (err2 <= (double)N) {
            cnt.steps++;
            st.t = last ? tstop : st.t + h;
#pragma unroll
            for (int i = 0; i < N; ++i) st.y[i] = ynew[i];
            st.have_k1 = false;                       // not FSAL: f(t_n, y_n) is evaluated at the new point
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
