// psi_engine.cuh — the fused psi kernel: one thread per (subject, support point) pair walks the
// subject's flattened timeline, simulates (closed form or adaptive ODE), evaluates the outputs at
// every observation, accumulates the log-likelihood and stores psi column-major.
//
// Mapping: threadIdx/blockIdx.x run over support points (a warp = 32 consecutive support points of
// ONE subject, so all lanes walk the same timeline and only parameter-dependent control flow
// diverges: lag-shifted bolus times and adaptive step counts); blockIdx.y runs over subjects.
// Support points are SoA (P x nspp) => every parameter load is a coalesced 256-B warp transaction;
// event / covariate / infusion records are warp-uniform broadcast loads that live in L1.
//
// Semantics follow the reference's DSL runtime (the device route for user models):
//   event loop          dsl/native.rs:1797-1868 (analytical), 1178-1513 (ODE); == the generic loop
//                       equation/mod.rs:480-516 + simulate_event :300-358 on the reference fixtures
//   route properties    dsl/native.rs:926-1025   (lag at the bolus time, fa at the lagged time)
//   initial state       dsl/native.rs:886-924    (derive at t=0, init only on occasion 0)
//   observations        dsl/native.rs:1044-1086  (derive at t_obs with active route inputs)
//   analytical interval dsl/native.rs:1871-1937  (== analytical/mod.rs:299-370)
//   ODE run_events      ode/mod.rs:609-824, closure.rs:103-195 (stop at infusion boundaries,
//                       restart after bolus / boundary)
// `M` is the model policy struct emitted by the DSL -> CUDA-C code generator (host/dsl_emit.cpp).
#pragma once
#include "psi_analytical.cuh"
#include "psi_ode.cuh"
#include "psi_stiff.cuh"
#include "psi_bdf.cuh"
#include "psi_sde.cuh"

// records of a subject's timeline program a CTA stages in shared memory (48 B each): 96 records = 4.5 KB
#ifndef PSI_PROG_STAGE
#define PSI_PROG_STAGE 96
#endif

namespace psi {

// ODE right-hand side: derive/covariates refreshed at every call at absolute t (native.rs:1221-1268)
template <class M>
struct OdeRhs {
    PairCtx<M>& c;
    PSI_DEV void operator()(double t, const double* x, double* dx) {
        if constexpr (M::RHS_USES_COV) fill_cov<M>(*c.pop, c.occ, t, c.cov);
        if constexpr (M::HAS_DERIVE && M::DERIVE_DEPS != 0 && M::RHS_USES_DERIVED) M::derive(t, x, c.p, c.cov, c.rate, c.d);
        M::dynamics(t, x, c.p, c.cov, c.rate, c.d, dx);
    }
    PSI_DEV void jacobian(double t, const double* x, double* J) {
        if constexpr (M::RHS_USES_COV) fill_cov<M>(*c.pop, c.occ, t, c.cov);
        M::jacobian(t, x, c.p, c.cov, c.rate, J);
    }
};

// The solver is a compile-time parameter of the kernel so the explicit kernels do not pay the
// register cost of the implicit ones (Jacobian + LU) and vice versa.
// Memory a solver keeps between two integrate-to-stop calls of one occasion (only the multistep method has any).
template <int SOLVER, int N> struct SolverMem { using type = NoSolverMem; };
template <int N> struct SolverMem<SOLVER_BDF, N> { using type = BdfState<N>; };

template <class M, int SOLVER>
PSI_DEV int ode_advance(OdeState<M::NSTATE>& st, typename SolverMem<SOLVER, M::NSTATE>::type& mem, double tstop, OdeRhs<M>& f, const RunOpts& opt,
                        Counters& cnt) {
    if constexpr (SOLVER == SOLVER_BDF) return bdf_integrate_to<M::NSTATE>(st, mem, tstop, f, opt, cnt);
    else if constexpr (SOLVER == SOLVER_ESDIRK34) return esdirk34_integrate_to<M::NSTATE>(st, tstop, f, opt, cnt);
    else if constexpr (SOLVER == SOLVER_TSIT5) return erk_integrate_to<Tsit5, M::NSTATE>(st, tstop, f, opt, cnt);
    else if constexpr (SOLVER == SOLVER_SDIRK4) return sdirk4_integrate_to<M::NSTATE>(st, tstop, f, opt, cnt);
    else if constexpr (SOLVER == SOLVER_TRBDF2) return trbdf2_integrate_to<M::NSTATE>(st, tstop, f, opt, cnt);
    else if constexpr (SOLVER == SOLVER_RODAS4) return rodas4_integrate_to<M::NSTATE, M::RHS_TIME_DEP>(st, tstop, f, opt, cnt);
    else return erk_integrate_to<Dopri5, M::NSTATE>(st, tstop, f, opt, cnt);
}

// One (subject, support point) pair.  Returns the summed log-likelihood; `status` != 0 on error.
template <class M, int SOLVER>
PSI_DEV double run_pair(const PopView& pop, const RunOpts& opt, int subj, PairCtx<M>& c, int& status, Counters& cnt,
                        double* __restrict__ pred, long long pred_ld) {
    constexpr int NS = M::NSTATE;
    constexpr int NR = AtLeast1<M::NROUTE>::v;
    double ll = 0.0;
    c.pop = &pop;

    // time-invariant derive is hoisted out of everything (derive reads only parameters/constants)
    if constexpr (M::HAS_DERIVE && M::DERIVE_DEPS == 0) {
        double zx[AtLeast1<NS>::v];
#pragma unroll
        for (int k = 0; k < AtLeast1<NS>::v; ++k) zx[k] = 0.0;
        c.zero_rate();
        M::derive(0.0, zx, c.p, c.cov, c.rate, c.d);
    } else {
#pragma unroll
        for (int k = 0; k < AtLeast1<M::NDER>::v; ++k) c.d[k] = 0.0;
    }

    // analytical: hoist the kernel coefficients too when nothing they depend on varies
    [[maybe_unused]] AKernel<(M::KIND == 1 ? M::AKERNEL : 0)> ak;
    constexpr bool AK_HOISTED = (M::KIND == 1) && (!M::HAS_DERIVE || M::DERIVE_DEPS == 0 || !M::KP_USES_DERIVED);
    if constexpr (AK_HOISTED) {
        double kp[8];
        M::kparams(c.p, c.d, kp);
        ak.setup_kp(kp, status);
    }

    const int occ0 = __ldg(pop.occ_offsets + subj), occ1 = __ldg(pop.occ_offsets + subj + 1);
    for (int occ = occ0; occ < occ1; ++occ) {
        c.occ = occ;
        const InfRange inf = occ_infusions(pop, occ);

        // ---- initial state (native.rs:886-924) -------------------------------------------------
        double x[AtLeast1<NS>::v];
#pragma unroll
        for (int k = 0; k < AtLeast1<NS>::v; ++k) x[k] = 0.0;
        if constexpr (M::HAS_INIT) {
            if (__ldg(pop.occ_index + occ) == 0) {
                c.zero_rate();
                c.refresh(0.0, x);
                M::init(0.0, x, c.p, c.cov, c.rate, c.d, x);
            }
        }

        // ---- closed-form models without lag: execute the occasion's timeline program --------------------
        // (psi_types.h EV_STEP: the host has already walked events / boundaries / infusions once per occasion;
        //  the arithmetic per record is exactly what the generic walk below performs, so results are identical)
        if constexpr (M::KIND == 1 && !M::HAS_LAG) {
            const int pg0 = __ldg(pop.prog_offsets + occ), pg1 = __ldg(pop.prog_offsets + occ + 1);
            for (int q = pg0; q < pg1; ++q) {
                const EventRec e = load_event_any(c.prog + (q - c.prog_first));
                const int kind = ev_kind(e.meta);
                if (kind == EV_STEP) {
                    const double dt = e.a;
                    if constexpr (NR == 1) { c.rate[0] = e.b; }
                    else if constexpr (NR == 2) { c.rate[0] = e.b; c.rate[1] = e.w; }
                    else if constexpr (NR == 3) { c.rate[0] = e.b; c.rate[1] = e.w; c.rate[2] = e.sigma; }
                    else {
#pragma unroll
                        for (int k = 0; k < NR; ++k) c.rate[k] = __ldg(pop.prog_rates + (long long)e.obs_row * NR + k);
                    }
                    if constexpr (!AK_HOISTED) {
                        const double tder = opt.cov_time == COVTIME_INTERVAL_LENGTH ? dt : e.time;
                        bool done = false;
                        if constexpr (M::NCOV == 1 && NR == 1 && M::HAS_DERIVE) {
                            if (pop.prog_cov) {       // the host has interpolated the covariate for both derive-time conventions
                                c.cov[0] = opt.cov_time == COVTIME_INTERVAL_LENGTH ? e.sigma : e.w;
                                M::derive(tder, x, c.p, c.cov, c.rate, c.d);
                                done = true;
                            }
                        }
                        if (!done) c.refresh(tder, x);
                        double kp[8];
                        M::kparams(c.p, c.d, kp);
                        ak.setup_kp(kp, status);
                    }
                    ak.step(x, dt, c.rate[0]);
                    cnt.evals++;
                } else if (kind == EV_BOLUS) {
                    const int route = ev_index(e.meta);
                    double amount = e.a;
                    if constexpr (M::HAS_FA) {
                        double zx[AtLeast1<NS>::v];
#pragma unroll
                        for (int k = 0; k < AtLeast1<NS>::v; ++k) zx[k] = 0.0;
                        c.zero_rate();
                        c.refresh(e.time, zx);
                        const double fa = M::fa(route, e.time, zx, c.p, c.cov, c.rate, c.d);
                        if (fa != 1.0) amount *= fa;
                    }
                    const int dest = M::bolus_dest(route);
                    if (dest < 0) { if (status == ST_OK) status = ST_UNSUPPORTED_INPUT_ROUTE_KIND; }
                    add_at<AtLeast1<NS>::v>(x, dest, amount);
                } else {
                    if constexpr (M::OBS_USES_RATE) active_rates<NR>(inf, e.time, c.rate);
                    if constexpr (M::OBS_NEEDS_REFRESH) c.refresh(e.time, x);
                    double y[AtLeast1<M::NOUT>::v];
#pragma unroll
                    for (int k = 0; k < AtLeast1<M::NOUT>::v; ++k) y[k] = 0.0;
                    M::outputs(e.time, x, c.p, c.cov, c.rate, c.d, y);
                    const double yp = pick<AtLeast1<M::NOUT>::v>(y, ev_index(e.meta));
                    if (pred && e.obs_row >= 0) pred[(long long)e.obs_row * pred_ld] = yp;
                    if (opt.want_ll && ev_has_value(e.meta)) {
                        if (opt.diagonal) ll += resid_log_likelihood(opt, ev_index(e.meta), e.a, yp);
                        else ll += obs_log_likelihood(e, yp, status);
                    }
                }
            }
            continue;
        }

        // ---- event cursor with per-thread lag ---------------------------------------------------
        auto lag_of = [&](int route, double tb) -> double {
            if constexpr (M::HAS_LAG) {
                double zx[AtLeast1<NS>::v];
#pragma unroll
                for (int k = 0; k < AtLeast1<NS>::v; ++k) zx[k] = 0.0;
                c.zero_rate();
                c.refresh(tb, zx);
                return M::lag(route, tb, zx, c.p, c.cov, c.rate, c.d);
            } else {
                return 0.0;
            }
        };
        EventCursor<M, decltype(lag_of)> cur(pop, occ, lag_of);

        // ---- ODE solver state -------------------------------------------------------------------
        // (A flat per-lane loop — one step attempt OR one event per iteration, so lanes in different inter-event
        //  intervals share the step code — was measured on B200 and lost: C2 43.1 -> 49.3 ms, C4 104.2 -> 114.7 ms.
        //  Lanes reach an event in different iterations, so the event code runs several times per warp instead of once;
        //  with work-balanced columns the re-convergence at events costs less than that.  profiles/r02_tuning.md)
        [[maybe_unused]] OdeState<AtLeast1<NS>::v> st;
        [[maybe_unused]] OdeRhs<M> rhs{c};
        [[maybe_unused]] typename SolverMem<(M::KIND == 0 ? SOLVER : 0), AtLeast1<NS>::v>::type solver_mem;
        [[maybe_unused]] int bnd = 0, bnd_end = 0, abc = 0;
        if constexpr (M::KIND == 1) {
            abc = __ldg(pop.bnd_offsets + occ);
            bnd_end = __ldg(pop.bnd_offsets + occ + 1);
        }
        if constexpr (M::KIND == 0) {
            st.t = __ldg(pop.occ_t0 + occ);
            st.h = -1.0;
            st.have_k1 = false;
            st.h_post = -1.0;
            st.since_restart = 0;
            bnd = __ldg(pop.bnd_offsets + occ);
            bnd_end = __ldg(pop.bnd_offsets + occ + 1);
        }

        EventRec e;
        double te;
        bool have = cur.next(e, te);
        while (have) {
            const int kind = ev_kind(e.meta);
            if (kind == EV_BOLUS) {
                const int route = ev_index(e.meta);
                double amount = e.a;
                if constexpr (M::HAS_FA) {                       // native.rs:991-1018, at the lagged time
                    double zx[AtLeast1<NS>::v];
#pragma unroll
                    for (int k = 0; k < AtLeast1<NS>::v; ++k) zx[k] = 0.0;
                    c.zero_rate();
                    c.refresh(te, zx);
                    const double fa = M::fa(route, te, zx, c.p, c.cov, c.rate, c.d);
                    if (fa != 1.0) amount *= fa;
                }
                const int dest = M::bolus_dest(route);           // native.rs:1027-1042
                if (dest < 0) { if (status == ST_OK) status = ST_UNSUPPORTED_INPUT_ROUTE_KIND; }
                add_at<AtLeast1<NS>::v>(x, dest, amount);
                if constexpr (M::KIND == 0) st.have_k1 = false, st.h = -1.0;   // pending_reinit (ode/mod.rs:687)
            } else if (kind == EV_OBS) {
                // observation_prediction (native.rs:1044-1086)
                if constexpr (M::OBS_USES_RATE) active_rates<NR>(inf, te, c.rate);
                if constexpr (M::OBS_NEEDS_REFRESH) c.refresh(te, x);
                double y[AtLeast1<M::NOUT>::v];
#pragma unroll
                for (int k = 0; k < AtLeast1<M::NOUT>::v; ++k) y[k] = 0.0;
                M::outputs(te, x, c.p, c.cov, c.rate, c.d, y);
                const double yp = pick<AtLeast1<M::NOUT>::v>(y, ev_index(e.meta));
                if (pred && e.obs_row >= 0) pred[(long long)e.obs_row * pred_ld] = yp;
                if (opt.want_ll && ev_has_value(e.meta)) {
                    if (opt.diagonal) ll += resid_log_likelihood(opt, ev_index(e.meta), e.a, yp);
                    else ll += obs_log_likelihood(e, yp, status);
                }
            }
            // ---- advance to the next event ------------------------------------------------------
            EventRec en;
            double tn;
            have = cur.next(en, tn);
            if (have) {
                if constexpr (M::KIND == 1) {
                    // Analytical interval (native.rs:1871-1937): split at interior infusion
                    // boundaries, dedup at 1e-12, derive at the sub-interval end, kernel step.
                    // The reference collects the infusion boundaries strictly inside (te, tn) per call, sorts
                    // and de-duplicates them at 1e-12 (analytical/mod.rs:311-327); the flattener has already
                    // sorted the occasion's boundaries, so a per-thread cursor walks them instead of
                    // rescanning every infusion for every sub-interval.
                    if (te != tn) {
                        double last = te;
                        while (abc < bnd_end && __ldg(pop.bnds + abc) <= te) ++abc;
                        while (true) {
                            while (abc < bnd_end) {          // drop candidates that the 1e-12 dedup would remove
                                const double bq = __ldg(pop.bnds + abc);
                                if (bq <= last || fabs(bq - last) < 1e-12) ++abc; else break;
                            }
                            double nxt;
                            const double bq = (abc < bnd_end) ? __ldg(pop.bnds + abc) : psi_inf();
                            if (bq < tn) { nxt = bq; ++abc; }
                            else if (tn > last && !(fabs(tn - last) < 1e-12)) nxt = tn;
                            else break;
                            const double dt = nxt - last;
                            interval_rates<NR>(inf, last, nxt, c.rate);
                            if constexpr (!AK_HOISTED) {
                                // derive time: sub-interval END (DSL, native.rs:1903-1916) or the
                                // sub-interval LENGTH (analytical! macro quirk)
                                c.refresh(opt.cov_time == COVTIME_INTERVAL_LENGTH ? dt : nxt, x);
                                double kp[8];
                                M::kparams(c.p, c.d, kp);
                                ak.setup_kp(kp, status);
                            }
                            ak.step(x, dt, c.rate[0]);           // built-in kernels read rateiv[0] only (A.3)
                            cnt.evals++;
                            last = nxt;
                            if (nxt == tn) break;
                        }
                    }
                } else if constexpr (M::KIND == 0) {
                    // ODE::run_events advance loop (ode/mod.rs:718-819)
#pragma unroll
                    for (int k = 0; k < NS; ++k) st.y[k] = x[k];
                    while (tn > st.t) {
                        while (bnd < bnd_end && __ldg(pop.bnds + bnd) <= st.t) ++bnd;
                        double stop = tn;
                        bool is_bnd = false;
                        if (bnd < bnd_end) {
                            const double b = __ldg(pop.bnds + bnd);
                            if (b <= tn) { stop = b; is_bnd = true; ++bnd; }
                        }
                        segment_rates<NR>(inf, st.t, c.rate);    // constant on [st.t, stop): right-continuous at
                                                                 // st.t, left-continuous at stop (closure.rs:43-51)
                        const int rc = ode_advance<M, SOLVER>(st, solver_mem, stop, rhs, opt, cnt);
                        if (rc != ST_OK) { if (status == ST_OK) status = rc; st.t = tn; break; }
                        if (is_bnd) st.have_k1 = false;          // RHS discontinuity: refresh dy (ode/mod.rs:568-586)
                    }
#pragma unroll
                    for (int k = 0; k < NS; ++k) x[k] = st.y[k];
                }
            }
            e = en;
            te = tn;
        }
    }
    return ll;
}

// ---------------------------------------------------------------------------------------------
// Kernel entry: grid = (ceil(ncols/128), min(nsub, 65535)), block = 128.
// spp is SoA: spp[k*spp_ld + j]; psi is column-major: out.ll[i + j*out.ld_ll].
// ---------------------------------------------------------------------------------------------
template <class M, int SOLVER, int MODE>
__device__ __forceinline__ void psi_pair_loop(const PopView& pop, const double* __restrict__ spp, long long ncols,
                                                long long spp_ld, const RunOpts& opt, const OutView& out) {
    const int nsub = (opt.nsub_limit > 0 && opt.nsub_limit < pop.nsub) ? opt.nsub_limit : pop.nsub;
    Counters cnt;
    // MODE (compile time; psi_kernel_body picks the instantiation): the three index spaces of a launch
    //   0 matrix       blockIdx.x * blockDim.x + threadIdx.x = column slot, blockIdx.y strides over subjects
    //   1 warp tasks   few support points (a 128-column CTA per subject would idle most lanes): the warps of a 1-D grid
    //                  take (subject, 32-column chunk) tasks in order, so a CTA mixes subjects
    //   2 diagonal     log_likelihood_batch: thread q = subject q with parameter row q
    // One source, three instantiations: with the mode a run-time value the Dopri5 kernel of C2 executed 5.8 % more
    // instructions per launch than with a dedicated copy per mode (ncu, profiles/r02_tuning.md).
    constexpr bool kWarpTasks = MODE == 1, kDiagonal = MODE == 2;
    // Only ONE 32-bit value stays live across a pair (it); the bounds and everything else about the index space are
    // recomputed from the launch parameters and special registers when needed — a pair of an ODE model runs for
    // ~10^5 instructions under an 80-register cap, so every value held across it costs spills.
    // Closed-form models without lag stage the subject's timeline program in shared memory (below): then EVERY thread of
    // the CTA walks the subject loop — also the ones past the last column — because the staging uses block barriers.
    constexpr bool kStaged = (M::KIND == 1 && !M::HAS_LAG);
    auto it_end = [&]() -> int {
        if (kWarpTasks) return (int)((ncols + 31) >> 5) * nsub;
        if (kDiagonal) return ((long long)blockIdx.x * blockDim.x + threadIdx.x < ncols) ? 1 : 0;
        const bool in_range = (long long)blockIdx.x * blockDim.x + threadIdx.x < ncols;
        return (in_range || kStaged) ? nsub : 0;
    };
    auto it_step = [&]() -> int { return kWarpTasks ? (int)(gridDim.x * (blockDim.x >> 5)) : (kDiagonal ? 1 : (int)gridDim.y); };
    int it = kWarpTasks ? (int)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) : (kDiagonal ? 0 : (int)blockIdx.y);
    for (; it < it_end(); it += it_step()) {
        const bool diag = kDiagonal;
        int subj;
        long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        if (kWarpTasks) {
            const int nchunk = (int)((ncols + 31) >> 5);
            subj = it / nchunk;
            q = (long long)(it - subj * nchunk) * 32 + (threadIdx.x & 31);
            if (q >= ncols) continue;
        } else {
            subj = diag ? (int)q : it;
        }
        // Closed-form models: in the matrix index space a CTA is 128 columns of ONE subject, so its threads execute the
        // same timeline program.  Stage it in shared memory once per CTA (one coalesced copy) instead of letting every warp
        // chase the records through L1 / L2 one dependent load at a time (ncu on C1 before this: 8.7 long-scoreboard
        // stall cycles per issued instruction, L1 hit rate 76 %).  Longer programs stay in global memory.
        [[maybe_unused]] const EventRec* prog_ptr = pop.prog;
        [[maybe_unused]] int prog_first = 0;
        if constexpr (kStaged) {
            if (!kWarpTasks && !diag) {                         // uniform over the CTA: every thread takes the barriers
                __shared__ double2 sprog[PSI_PROG_STAGE * 3];
                const int p0 = __ldg(pop.prog_offsets + __ldg(pop.occ_offsets + subj));
                const int p1 = __ldg(pop.prog_offsets + __ldg(pop.occ_offsets + subj + 1));
                __syncthreads();                                    // the previous subject's readers are done with the buffer
                if (p1 - p0 <= PSI_PROG_STAGE) {
                    const double2* src = reinterpret_cast<const double2*>(pop.prog + p0);
#ifdef PSI_HOST_SIM                                                 // one "thread" at a time: each stages the whole program
                    for (int i = 0; i < (p1 - p0) * 3; ++i) sprog[i] = src[i];
#else
                    for (int i = threadIdx.x; i < (p1 - p0) * 3; i += blockDim.x) sprog[i] = __ldg(src + i);
#endif
                    prog_ptr = reinterpret_cast<const EventRec*>(sprog);
                    prog_first = p0;
                }
                __syncthreads();
                if (q >= ncols) continue;                           // a thread past the last column only helped to stage
            }
        }
        // work-balanced warps: slot q -> column col_perm[q] (columns ordered by probed step counts); the
        // parameter loads become a gather (P loads per pair, nothing against hundreds of solver steps)
        const long long j = out.col_perm ? (long long)__ldg(out.col_perm + q) : q;
        PairCtx<M> c;
#pragma unroll
        for (int k = 0; k < M::NP; ++k) c.p[k] = __ldg(spp + (long long)k * spp_ld + j);
        M::prologue(c.p);
        c.prog = prog_ptr;
        c.prog_first = prog_first;
        int status = ST_OK;
        double* pred = (opt.want_pred && out.pred && !diag) ? out.pred + j : nullptr;
        const unsigned int work0 = cnt.steps + cnt.rejected;
        double ll = run_pair<M, SOLVER>(pop, opt, subj, c, status, cnt, pred, out.ld_pred);
        if (out.col_work) atomicAdd(out.col_work + j, cnt.steps + cnt.rejected - work0);
        if (diag) {
            // a failed simulation scores -inf (likelihood/mod.rs:134-137)
            if (out.ll) out.ll[q] = (status != ST_OK) ? -psi_inf() : ll;
            continue;
        }
        if (status != ST_OK) {
            ll = psi_nan();
            report_error(out, (long long)subj + (j + out.col_base) * (long long)pop.nsub, status);
        }
        if (out.npeers > 0) {
            const long long at = (long long)subj + (j + out.col_base) * out.ld_ll;
#pragma unroll
            for (int r = 0; r < 8; ++r)
                if (r < out.npeers) out.ll_peers[r][at] = ll;      // posted 8-byte stores, one per rank
        } else if (out.ll) {
            out.ll[(long long)subj + j * out.ld_ll] = ll;
        }
    }
    flush_counters(out, cnt);
}

template <class M, int SOLVER>
__device__ __forceinline__ void psi_kernel_body(const PopView& pop, const double* __restrict__ spp, long long ncols,
                                                long long spp_ld, const RunOpts& opt, const OutView& out) {
    if (opt.warp_tasks) psi_pair_loop<M, SOLVER, 1>(pop, spp, ncols, spp_ld, opt, out);
    else if (opt.diagonal) psi_pair_loop<M, SOLVER, 2>(pop, spp, ncols, spp_ld, opt, out);
    else psi_pair_loop<M, SOLVER, 0>(pop, spp, ncols, spp_ld, opt, out);
}

template <class M, int SOLVER>
__device__ __forceinline__ void psi_dispatch(const PopView& pop, const double* __restrict__ spp, long long ncols,
                                             long long spp_ld, const RunOpts& opt, const OutView& out) {
    if constexpr (M::KIND == 2) psi_sde_kernel_body<M>(pop, spp, ncols, spp_ld, opt, out);
    else psi_kernel_body<M, SOLVER>(pop, spp, ncols, spp_ld, opt, out);
}

}  // namespace psi

// The emitted translation unit instantiates:  PSI_DEFINE_ENTRY(Model, SOLVER, psi_entry_<id>)
#ifndef PSI_MAX_THREADS
#define PSI_MAX_THREADS 128
#endif
// Resident CTAs of 128 threads per SM the compiler must allow (register cap = 65536 / (128 * MIN_BLOCKS)), chosen per
// model by the emitter (MODEL::MIN_BLOCKS) from measurements on B200 (scripts/tune.py, profiles/):
//   6 (<= 80 registers)  default.  The 3-compartment closed form gains 1.66x and the stiff kernels 1.4x over the
//                        unconstrained build (140-180 registers of hoisted loop invariants, 3 warps per scheduler);
//                        the 96-register Dopri5 kernel of the 3-state model loses 2 %.  C3: 6 -> 29.3 ms, 5 -> 30.6, 8 -> 30.7.
//  10 (<= 48 registers)  one-compartment closed forms: latency / issue bound, no spills in the step loop.
//                        C1: 6 -> 0.1145 ms, 8 -> 0.0932, 10 -> 0.0876, 12 -> 0.0965.
//   8 (<= 64 registers)  two-compartment closed forms (between the two).
// -DPSI_MIN_BLOCKS=n (NVRTC flag, scripts/tune.py) overrides the model's choice.
#ifdef PSI_MIN_BLOCKS
#define PSI_ENTRY_MIN_BLOCKS(MODEL) PSI_MIN_BLOCKS
#else
#define PSI_ENTRY_MIN_BLOCKS(MODEL) MODEL::MIN_BLOCKS
#endif
#define PSI_DEFINE_ENTRY(MODEL, SOLVER, NAME)                                                                 \
    extern "C" __global__ void __launch_bounds__(PSI_MAX_THREADS, PSI_ENTRY_MIN_BLOCKS(MODEL))                 \
    NAME(psi::PopView pop, const double* __restrict__ spp, long long ncols, long long spp_ld, psi::RunOpts opt, \
         psi::OutView out) {                                                                                  \
        psi::psi_dispatch<MODEL, SOLVER>(pop, spp, ncols, spp_ld, opt, out);                                  \
    }
