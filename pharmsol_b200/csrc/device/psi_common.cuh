// psi_common.cuh — device helpers shared by the analytical / ODE / SDE psi kernels:
// covariate interpolation, the event cursor (with the per-thread lag merge), the likelihood
// epilogue and the error / counter plumbing.  sm_100a, FP64 scalar; no tensor cores by design.
//
// Reference semantics restated (file:line relative to /root/reference):
//   covariates        src/data/covariate.rs:189-241, dsl/native.rs:805-812 (missing -> NaN)
//   event order       src/data/event.rs:292-304; lag/fa src/data/structs.rs:611-690,
//                     dsl/native.rs:926-1025 (+ sort_events :2686-2702)
//   likelihood        src/simulator/likelihood/{distributions.rs:31-103, prediction.rs:105-125}
#pragma once
#include "psi_types.h"

#define PSI_DEV __device__ __forceinline__

namespace psi {

PSI_DEV double psi_nan() { return __longlong_as_double(0x7ff8000000000000LL); }
PSI_DEV double psi_inf() { return __longlong_as_double(0x7ff0000000000000LL); }

// reciprocal to ~2^-20 with ONE MUFU.RCP64H and no Newton refinement: enough for the weights of an
// error norm, and it moves the work from the FP64 pipe (the bound of these kernels) to the XU pipe.
PSI_DEV double rcp_approx(double x) {
#ifdef PSI_HOST_SIM      // tests/hostsim: the device headers compiled for the host (test infrastructure)
    return 1.0 / x;
#else
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
#endif
}

// x^p for the step-size controllers, x > 0 a normal float: MUFU.LG2 * p -> MUFU.EX2 with flush-to-zero and none of __powf's
// denormal scaling (ncu: the controller's __powf was 5 % of the Dopri5 kernel's instructions; it only steers h).
PSI_DEV float powf_fast(float x, float p) {
#ifdef PSI_HOST_SIM
    return __powf(x, p);
#else
    float l, r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(x));
    l *= p;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(l));
    return r;
#endif
}

// max(|a|, |b|) of two finite doubles on the INTEGER pipe: non-negative doubles order like their bit patterns.  The error
// norms of the explicit pairs need it once per component per step, and DSETP / FP64 selects would add to the pipe that
// bounds those kernels.  A NaN input gives a NaN-or-larger pattern; the step is then rejected by the non-finite error test.
PSI_DEV double max_abs(double a, double b) {
    // on the 32-bit halves: a 64-bit `& 0x7fff...` is recognised as fabs and comes back as an FP64-pipe DADD
    const unsigned int ah = (unsigned int)__double2hiint(a) & 0x7fffffffu, bh = (unsigned int)__double2hiint(b) & 0x7fffffffu;
    const unsigned int al = (unsigned int)__double2loint(a), bl = (unsigned int)__double2loint(b);
    const bool a_wins = ah > bh || (ah == bh && al > bl);
    return __hiloint2double((int)(a_wins ? ah : bh), (int)(a_wins ? al : bl));
}

// min / max of two doubles of which at most one is negative (and neither is -0), on the integer pipe: such doubles order
// like their bit patterns read as signed 64-bit integers.  +NaN orders above +inf.
PSI_DEV double max_pos(double a, double b) { return __double_as_longlong(a) > __double_as_longlong(b) ? a : b; }
PSI_DEV double min_pos(double a, double b) { return __double_as_longlong(a) < __double_as_longlong(b) ? a : b; }

// Single-MUFU FP32 helpers without the denormal fix-ups of __logf / sqrtf / rsqrtf (arguments are normal or +-0 / inf)
PSI_DEV float lg2_ftz(float x) {
#ifdef PSI_HOST_SIM
    return log2f(x);
#else
    float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#endif
}
PSI_DEV float sqrt_ftz(float x) {
#ifdef PSI_HOST_SIM
    return sqrtf(x);
#else
    float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#endif
}
PSI_DEV float rsqrt_ftz(float x) {
#ifdef PSI_HOST_SIM
    return 1.0f / sqrtf(x);
#else
    float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#endif
}

// Reciprocal / quotient to within 1 ulp without the IEEE fix-up path: MUFU.RCP64H seed + one third-order step (3 DFMA).  CUDA's
// `a / b` costs ~14 instructions plus a divergence-scoped branch to a slow path (ncu on the RODAS4 kernel: FSEL + FSETP +
// BSSY / BSYNC / BRA are 25 % of the executed instructions, most of them the fix-ups of ~10 divisions per step).  Used
// where the last bit does not matter: the emitted ODE dynamics / Jacobian (the solver's tolerance dominates), the LU
// pivots and 1/h.  A zero, infinite or denormal (ftz) divisor gives NaN (0 x inf inside the refinement) where IEEE division
// gives inf or 0: either way a non-finite right-hand side, which every solver rejects; NaN propagates.
// The seed is good to 2^-19.9 (measured, scripts/micro/rcp_check.cu: 3e8 log-uniform doubles); with e = 1 - x r the
// third-order update r (1 + e + e^2) leaves e^3 = 2^-60 before rounding: within 1 ulp of 1/x on every sample, in three
// dependent FMAs where two Newton steps take four.
PSI_DEV double rcp_nr(double x) {
    const double r = rcp_approx(x);
    const double e = fma(-x, r, 1.0);
    return fma(fma(e, e, e), r, r);
}
PSI_DEV double fdiv(double a, double b) { return a * rcp_nr(b); }

// ---------------------------------------------------------------------------------------------
// Covariates.  The host flattener emits, per (occasion, covariate), a leading sentinel segment
// (-inf, first.from) carrying the first observation's value, then the reference's segments with
// the last one open-ended, so one scan reproduces Covariate::interpolate exactly.
// ---------------------------------------------------------------------------------------------
PSI_DEV double cov_value(const CovSeg* __restrict__ segs, int n, double t) {
    double v = psi_nan();   // no segments (covariate missing for this occasion) or NaN time
    for (int i = 0; i < n; ++i) {
        const double2 ft = __ldg(reinterpret_cast<const double2*>(segs + i));       // from, to
        if (ft.x <= t && t < ft.y) {
            const double2 si = __ldg(reinterpret_cast<const double2*>(segs + i) + 1);   // slope, intercept
            v = si.x * t + si.y;
            break;
        }
    }
    return v;
}

template <class M>
PSI_DEV void fill_cov(const PopView& pop, int occ, double t, double* cov) {
    if constexpr (M::NCOV > 0) {
#pragma unroll
        for (int c = 0; c < M::NCOV; ++c) {
            const int o0 = __ldg(pop.cov_offsets + occ * M::NCOV + c);
            const int o1 = __ldg(pop.cov_offsets + occ * M::NCOV + c + 1);
            cov[c] = cov_value(pop.cov_segs + o0, o1 - o0, t);
        }
    }
}

// pow(x, c) for the exponents PK models actually use (allometric 0.75 / 0.25, square roots, small integers): square
// roots and multiplications instead of the ~150-instruction general pow.  sqrt is correctly rounded, so these are
// within 1.5 ulp — the same class as CUDA's pow (2 ulp) and the reference's libm powf.
PSI_DEV double pow_half(double x) { return sqrt(x); }
PSI_DEV double pow_quarter(double x) { return sqrt(sqrt(x)); }
PSI_DEV double pow_three_quarters(double x) { const double s = sqrt(x); return s * sqrt(s); }
PSI_DEV double pow_three_halves(double x) { return x * sqrt(x); }
PSI_DEV double pow_2(double x) { return x * x; }
PSI_DEV double pow_3(double x) { return (x * x) * x; }
PSI_DEV double pow_4(double x) { const double q = x * x; return q * q; }

template <int N> struct AtLeast1 { static constexpr int v = N > 0 ? N : 1; };

// Per-pair context shared by the model callbacks.
template <class M>
struct PairCtx {
    double p[AtLeast1<M::NPX>::v];     // parameters, then the pair-invariant slots written by M::prologue
    double cov[AtLeast1<M::NCOV>::v];
    double d[AtLeast1<M::NDER>::v];
    double rate[AtLeast1<M::NROUTE>::v];
    const PopView* pop;
    int occ;
    // timeline program of the pair's subject: record q lives at prog[q - prog_first] — the global array (prog_first = 0)
    // or the CTA's shared-memory copy of this subject's records (psi_engine.cuh psi_kernel_body)
    const EventRec* prog;
    int prog_first;

    // SharedNativeModel::refresh_derived (dsl/native.rs:828-857): covariates at `t`, then derive.
    PSI_DEV void refresh(double t, const double* x) {
        fill_cov<M>(*pop, occ, t, cov);
        if constexpr (M::HAS_DERIVE) {
            if constexpr (M::DERIVE_DEPS != 0) M::derive(t, x, p, cov, rate, d);
        }
    }
    PSI_DEV void zero_rate() {
#pragma unroll
        for (int k = 0; k < AtLeast1<M::NROUTE>::v; ++k) rate[k] = 0.0;
    }
};

// ---------------------------------------------------------------------------------------------
// Error / counter plumbing.  No per-pair status array (C3 would need 2 GB): the first failing
// pair (lowest linear index) and its code are kept in one 64-bit word by atomicMin; failing
// pairs get NaN in psi.  matrix.rs:96-104 aborts on the first error; the host maps the code back.
// ---------------------------------------------------------------------------------------------
PSI_DEV void report_error(const OutView& out, long long pair, int code) {
    if (code != ST_OK && out.first_error)
        atomicMin(out.first_error, ((unsigned long long)pair << 8) | (unsigned long long)(code & 0xff));
}

struct Counters {
    unsigned int steps = 0, rejected = 0, evals = 0, newton = 0;
};
PSI_DEV void flush_counters(const OutView& out, const Counters& c) {
    if (!out.counters) return;
    unsigned int v[4] = {c.steps, c.rejected, c.evals, c.newton};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        unsigned long long s = v[k];
        // butterfly over the active lanes of the warp, one atomic per warp
        const unsigned mask = __activemask();
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(mask, s, off);
        // with a partial mask the xor butterfly still sums correctly only for full warps; use the
        // safe path otherwise
        if (mask == 0xffffffffu) {
            if ((threadIdx.x & 31) == 0 && s) atomicAdd(out.counters + k, s);
        } else if (v[k]) {
            atomicAdd(out.counters + k, (unsigned long long)v[k]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Likelihood epilogue: one observation.
//   Censor::None : -0.5 ln 2pi - ln sigma - (o-p)^2/(2 sigma^2) = c - d*d*w   (c, w precomputed)
//   BLOQ / ALOQ  : ln Phi / ln(1-Phi) with statrs' cdf = 0.5 erfc((mu-x)/(sigma sqrt2)) and the
//                  reference's |z| > 37 asymptotic branch.
// ---------------------------------------------------------------------------------------------
PSI_DEV double obs_log_likelihood(const EventRec& e, double pred, int& status) {
    const int cens = ev_cens(e.meta);
    const int host_status = ev_status(e.meta);
    if (host_status != ST_OK) {   // sigma < 0 / non-finite / missing error model: decided on the host
        if (status == ST_OK) status = host_status;
        return psi_nan();
    }
    const double obs = e.a;
    double ll;
    if (cens == CENS_NONE) {
        const double d = obs - pred;
        ll = e.b - (d * d) * e.w;
    } else {
        const double sigma = e.sigma;
        if (!(sigma > 0.0) || pred != pred) {      // statrs Normal::new(mean, sd) rejects these
            if (status == ST_OK) status = ST_NEGATIVE_SIGMA;
            return psi_nan();
        }
        const double cdf = 0.5 * erfc((pred - obs) / (sigma * 1.4142135623730951));
        const double tail = (cens == CENS_BLOQ) ? cdf : 1.0 - cdf;
        if (tail <= 0.0) {
            const double z = (obs - pred) / sigma;
            const bool asym = (cens == CENS_BLOQ) ? (z < -37.0) : (z > 37.0);
            if (!asym) {
                if (status == ST_OK) status = ST_NEGATIVE_SIGMA;   // distributions.rs:66, :99
                return psi_nan();
            }
            const double d = obs - pred;
            ll = (e.b - (d * d) * e.w) - log(fabs(z));
        } else {
            ll = log(tail);
        }
    }
    if (!isfinite(ll) && status == ST_OK) status = ST_NON_FINITE_LIKELIHOOD;   // prediction.rs:120-124
    return ll;
}

// ResidualErrorModels::log_likelihood (data/residual_error.rs:178-197, 265-271, 399-426):
// sigma from the PREDICTION with the sqrt(eps) cutoff; a missing model scores -inf.
PSI_DEV double resid_log_likelihood(const RunOpts& opt, int outeq, double obs, double pred) {
    int kind = RESID_MISSING;
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int k = 0; k < PSI_MAX_RESID; ++k)
        if (k == outeq && k < opt.nresid) { kind = opt.resid[k].kind; a = opt.resid[k].a; b = opt.resid[k].b; }
    if (kind == RESID_MISSING) return -psi_inf();
    double raw;
    if (kind == RESID_PROPORTIONAL) raw = b * fabs(pred);
    else if (kind == RESID_COMBINED) raw = sqrt(a * a + (b * b) * (pred * pred));
    else raw = a;
    const double sigma = fmax(raw, 1.4901161193847656e-08);     // f64::EPSILON.sqrt()
    const double nr = (obs - pred) / sigma;
    return -0.5 * (1.8378770664093453 + 2.0 * log(sigma) + nr * nr);
}

// ---------------------------------------------------------------------------------------------
// Infusion helpers (warp-uniform data).
// ---------------------------------------------------------------------------------------------
struct InfRange {
    const InfRec* p;
    int n;
};
PSI_DEV InfRange occ_infusions(const PopView& pop, int occ) {
    const int a = __ldg(pop.inf_offsets + occ), b = __ldg(pop.inf_offsets + occ + 1);
    return InfRange{pop.infs + a, b - a};
}
PSI_DEV void load_inf(const InfRec* r, double& time, double& dur, double& rate, int& input) {
    const double2 td = __ldg(reinterpret_cast<const double2*>(r));
    const double2 ai = __ldg(reinterpret_cast<const double2*>(r) + 1);
    time = td.x; dur = td.y; rate = ai.x;
    input = (int)(__double_as_longlong(ai.y) & 0xffffffffLL);
}
template <int NROUTE>
PSI_DEV void add_rate(double* rate, int input, double r) {
#pragma unroll
    for (int k = 0; k < NROUTE; ++k) rate[k] += (k == input) ? r : 0.0;
}
// interval_route_inputs (dsl/native.rs:2667-2684) == analytical/mod.rs:337-357:
// rate active on [cur, next] iff cur >= start && next <= start + duration
template <int NROUTE>
PSI_DEV void interval_rates(const InfRange& inf, double cur, double next, double* rate) {
#pragma unroll
    for (int k = 0; k < NROUTE; ++k) rate[k] = 0.0;
    for (int i = 0; i < inf.n; ++i) {
        double s, d, a; int input;
        load_inf(inf.p + i, s, d, a, input);
        if (cur >= s && next <= s + d) add_rate<NROUTE>(rate, input, a);
    }
}
// active_route_inputs (dsl/native.rs:2651-2665) == SDE drift rule (sde/mod.rs:124-133):
// start <= t <= start + duration, closed at both ends
template <int NROUTE>
PSI_DEV void active_rates(const InfRange& inf, double t, double* rate) {
#pragma unroll
    for (int k = 0; k < NROUTE; ++k) rate[k] = 0.0;
    for (int i = 0; i < inf.n; ++i) {
        double s, d, a; int input;
        load_inf(inf.p + i, s, d, a, input);
        if (t >= s && t <= s + d) add_rate<NROUTE>(rate, input, a);
    }
}
// ODE InfusionSchedule (ode/closure.rs:103-195): right-continuous cumulative rate at `t`
// (= the rate on the segment that starts at t); infusions with duration <= 0 are skipped.
template <int NROUTE>
PSI_DEV void segment_rates(const InfRange& inf, double t, double* rate) {
#pragma unroll
    for (int k = 0; k < NROUTE; ++k) rate[k] = 0.0;
    for (int i = 0; i < inf.n; ++i) {
        double s, d, a; int input;
        load_inf(inf.p + i, s, d, a, input);
        if (d > 0.0 && s <= t && t < s + d) add_rate<NROUTE>(rate, input, a);
    }
}

// ---------------------------------------------------------------------------------------------
// Event record load (warp-uniform address -> broadcast)
// ---------------------------------------------------------------------------------------------
PSI_DEV EventRec load_event(const EventRec* __restrict__ p) {
    EventRec e;
    const double2 a = __ldg(reinterpret_cast<const double2*>(p));
    const double2 b = __ldg(reinterpret_cast<const double2*>(p) + 1);
    const double2 c = __ldg(reinterpret_cast<const double2*>(p) + 2);
    e.time = a.x; e.a = a.y; e.b = b.x; e.w = b.y; e.sigma = c.x;
    const long long m = __double_as_longlong(c.y);
    e.meta = (int)(m & 0xffffffffLL);
    e.obs_row = (int)(m >> 32);
    return e;
}

// The same record through a generic pointer (the shared-memory copy of a subject's timeline program).
PSI_DEV EventRec load_event_any(const EventRec* p) {
    EventRec e;
    const double2 a = *reinterpret_cast<const double2*>(p);
    const double2 b = *(reinterpret_cast<const double2*>(p) + 1);
    const double2 c = *(reinterpret_cast<const double2*>(p) + 2);
    e.time = a.x; e.a = a.y; e.b = b.x; e.w = b.y; e.sigma = c.x;
    const long long m = __double_as_longlong(c.y);
    e.meta = (int)(m & 0xffffffffLL);
    e.obs_row = (int)(m >> 32);
    return e;
}

// ---------------------------------------------------------------------------------------------
// Event cursor.  Without lag the occasion's events are walked in their stored order.  With lag
// (parameter dependent => per thread) the boluses form a second stream whose times are shifted
// by lag(p, t_bolus, cov); the two streams are merged with the reference's comparator
// (time, then Observation < Bolus < Infusion, stable).  If the lagged bolus times are not
// non-decreasing in the original order (pathological lag functions) the cursor switches to an
// O(nb^2) selection that needs no per-thread storage, so the result is always the stable sort.
// ---------------------------------------------------------------------------------------------
template <class M, class LagFn>
struct EventCursor {
    const PopView& pop;
    LagFn lag_of;          // (route, t_bolus) -> lag
    int ev, ev_end;        // stream A cursor (all events when !HAS_LAG; non-bolus events otherwise)
    int bol, bol_begin, bol_end;   // stream B
    double tb;             // lagged time of the head of stream B (+inf when exhausted)
    int bsel;              // selected bolus (slow path)
    bool slow;
    double last_t; int last_i;   // slow path: last emitted (time, original index)

    PSI_DEV EventCursor(const PopView& p, int occ, LagFn f) : pop(p), lag_of(f) {
        ev = __ldg(pop.ev_offsets + occ);
        ev_end = __ldg(pop.ev_offsets + occ + 1);
        bol = bol_begin = bol_end = 0; tb = psi_inf(); bsel = -1; slow = false; last_t = -psi_inf(); last_i = -1;
        if constexpr (M::HAS_LAG) {
            bol_begin = bol = __ldg(pop.bol_offsets + occ);
            bol_end = __ldg(pop.bol_offsets + occ + 1);
            // pre-pass: are lagged times non-decreasing in original order?
            double prev = -psi_inf();
            for (int b = bol_begin; b < bol_end; ++b) {
                const double t = lagged_time(b);
                if (t < prev) slow = true;
                prev = t;
            }
            skip_boluses();
            advance_b_head(true);
        }
    }
    PSI_DEV double lagged_time(int b) const {
        const EventRec* e = pop.events + __ldg(pop.bol_event + b);
        const double t = __ldg(&e->time);
        const int meta = __ldg(&e->meta);
        const double l = lag_of(ev_index(meta), t);
        return (l != 0.0) ? t + l : t;     // structs.rs:636-639 / native.rs:984-987
    }
    PSI_DEV void skip_boluses() {
        while (ev < ev_end && ev_kind(__ldg(&pop.events[ev].meta)) == EV_BOLUS) ++ev;
    }
    PSI_DEV void advance_b_head(bool first) {
        if (!slow) {
            if (!first) ++bol;
            tb = (bol < bol_end) ? lagged_time(bol) : psi_inf();
            bsel = bol;
        } else {
            // smallest (time, index) strictly greater than (last_t, last_i)
            double best_t = psi_inf(); int best = -1;
            for (int b = bol_begin; b < bol_end; ++b) {
                const double t = lagged_time(b);
                const bool after = (t > last_t) || (t == last_t && b > last_i);
                if (after && (best < 0 || t < best_t)) { best_t = t; best = b; }
            }
            tb = (best >= 0) ? best_t : psi_inf();
            bsel = best;
        }
    }
    // Is there another event?  (time of the next event without consuming it)
    PSI_DEV bool peek_time(double& t) const {
        if constexpr (!M::HAS_LAG) {
            if (ev >= ev_end) return false;
            t = __ldg(&pop.events[ev].time);
            return true;
        } else {
            const bool has_a = ev < ev_end;
            const bool has_b = bsel >= 0 && bsel < bol_end && tb < psi_inf();
            if (!has_a && !has_b) return false;
            double ta = has_a ? __ldg(&pop.events[ev].time) : psi_inf();
            if (has_b && (!has_a || take_b(ta))) t = tb; else t = ta;
            return true;
        }
    }
    PSI_DEV bool take_b(double ta) const {
        // equal times: observation (A) first, then bolus (B), then infusion (A)
        if (tb < ta) return true;
        if (tb > ta) return false;
        return ev_kind(__ldg(&pop.events[ev].meta)) == EV_INFUSION;
    }
    // Consume the next event.  `time` is the (possibly lagged) event time.
    PSI_DEV bool next(EventRec& e, double& time) {
        if constexpr (!M::HAS_LAG) {
            if (ev >= ev_end) return false;
            e = load_event(pop.events + ev);
            time = e.time;
            ++ev;
            return true;
        } else {
            const bool has_a = ev < ev_end;
            const bool has_b = bsel >= 0 && bsel < bol_end && tb < psi_inf();
            if (!has_a && !has_b) return false;
            const double ta = has_a ? __ldg(&pop.events[ev].time) : psi_inf();
            if (has_b && (!has_a || take_b(ta))) {
                e = load_event(pop.events + __ldg(pop.bol_event + bsel));
                time = tb;
                last_t = tb; last_i = bsel;
                advance_b_head(false);
            } else {
                e = load_event(pop.events + ev);
                time = e.time;
                ++ev;
                skip_boluses();
            }
            return true;
        }
    }
};

// select y[idx] / add to x[idx] without dynamic register indexing
template <int N>
PSI_DEV double pick(const double* y, int idx) {
    double v = psi_nan();
#pragma unroll
    for (int k = 0; k < N; ++k) v = (k == idx) ? y[k] : v;
    return v;
}
template <int N>
PSI_DEV void add_at(double* x, int idx, double a) {
#pragma unroll
    for (int k = 0; k < N; ++k) x[k] += (k == idx) ? a : 0.0;
}

}  // namespace psi
