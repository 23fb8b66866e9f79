// psi_sde.cuh — SDE models: one CTA per (subject, support point) pair, particles strided over the
// CTA's threads; Euler-Maruyama with a counter-based Philox4x32-10 stream; block reductions for
// the weight sums; block scan + binary search for resampling.
//
// Two likelihood modes, because the reference has two:
//   SDE_MEAN_PREDICTION  what `log_likelihood_matrix` runs today: no weighting, no resampling;
//                        Gaussian log-pdf of the particle-MEAN prediction
//                        (equation/mod.rs:468-477 -> sde/mod.rs:387-433).
//   SDE_PARTICLE_FILTER  `SDE::estimate_log_likelihood` (sde/mod.rs:689-736): per observation
//                        q_i = exp(loglik_i), likelihood *= mean(q), resample (sde/mod.rs:526-577,
//                        747-767).  The product is accumulated as a sum of logs.
// Two steppers:
//   EM_REFERENCE_ADAPTIVE  sde/em.rs:134-167 as written: dt starts at 0.1 in every interval (the
//                          first attempt may overshoot tf), one full step vs two half steps with
//                          INDEPENDENT noise, err = max|y1-y2|/(1e-2+1e-2|x|), accept -> take y2.
//   EM_FIXED_STEP          plain Euler-Maruyama with n = ceil((tf-ti)/em_dt) equal steps.
// Resampling restates `sysresample` (sde/mod.rs:747-767): u_j = (j + U_j)/m with an independent
// U_j per j, ancestor = first k with cumsum[k] >= u_j (clamped to m-1 where the reference would
// index out of bounds).
// The reference RNG is an unseeded thread-local ChaCha (rand 0.10), so parity is statistical only.
#pragma once
#include "psi_ode.cuh"

namespace psi {

// ---- Philox4x32-10 (Salmon et al., SC'11) -------------------------------------------------------
struct Philox {
    unsigned int k0, k1;
    PSI_DEV void operator()(unsigned int c0, unsigned int c1, unsigned int c2, unsigned int c3, unsigned int* r) const {
        unsigned int a = k0, b = k1;
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            const unsigned int h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
            const unsigned int h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
            const unsigned int n0 = h1 ^ c1 ^ a, n2 = h0 ^ c3 ^ b;
            c0 = n0; c1 = l1; c2 = n2; c3 = l0;
            a += 0x9E3779B9u; b += 0xBB67AE85u;
        }
        r[0] = c0; r[1] = c1; r[2] = c2; r[3] = c3;
    }
};

// Per-particle normal stream: one Philox block = 4 normals.  Everything is indexed statically so the normals stay in
// registers (a runtime-indexed buffer would live in local memory).
// Precision of the noise (RunOpts::sde_normals): the default draws Box-Muller normals in FP32 (32-bit Philox words rounded
// to FP32 uniforms) on the SFU (MUFU.LG2 / SQRT / SIN / COS): |z| <= 6.76, 24-bit mantissa — narrower than the f64 `Normal` the reference samples
// (sde/em.rs:104-120), chosen because the stepper is bound by integer / SFU issue and an FP64 Box-Muller triples the cost
// of a draw.  SDE_NORMALS_FP64 evaluates the same transform in FP64 from 32-bit uniforms (|z| <= 6.66); the particle-filter
// likelihood does not distinguish the two (tests/test_gpu_sde_parity.py compares them at equal seed counts).
struct NormalStream {
    Philox ph;
    unsigned int c0, c2, c3;   // particle slot, interval sequence, pair
    unsigned int ctr;          // draw-block counter within the interval
    int fp64;                  // RunOpts::sde_normals
    PSI_DEV void reset(unsigned int particle, unsigned int seq, unsigned int pair) {
        c0 = particle; c2 = seq; c3 = pair; ctr = 0;
    }
    PSI_DEV void block4(double* z) {
        unsigned int r[4];
        ph(c0, ctr++, c2, c3, r);
        if (fp64) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const double u1 = ((double)r[2 * i] + 0.5) * (1.0 / 4294967296.0);      // (0,1), 32 bits
                const double u2 = ((double)r[2 * i + 1] + 0.5) * (1.0 / 4294967296.0);
                const double rad = sqrt(-2.0 * log(u1));
                double sn, cs;
                sincos(6.283185307179586 * u2, &sn, &cs);
                z[2 * i] = rad * cs;
                z[2 * i + 1] = rad * sn;
            }
            return;
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            // a 32-bit word rounded to FP32: u1 in [2^-33, 1], the angle in (0, 2 pi]; -2 ln u = -2 ln 2 * lg2 u
            const float u1 = fmaf((float)r[2 * i], 2.3283064365386963e-10f, 1.1641532182693481e-10f);
            const float a2 = fmaf((float)r[2 * i + 1], 1.4629180792671596e-09f, 7.314590396335798e-10f);
            const float rad = sqrt_ftz(-1.3862943611198906f * lg2_ftz(u1));
            float sn, cs;
            __sincosf(a2, &sn, &cs);
            z[2 * i] = (double)(rad * cs);
            z[2 * i + 1] = (double)(rad * sn);
        }
    }
    // COUNT normals into z[0..COUNT) (z must hold 4*ceil(COUNT/4) values)
    template <int COUNT>
    PSI_DEV void fill(double* z) {
#pragma unroll
        for (int b = 0; b < (COUNT + 3) / 4; ++b) block4(z + 4 * b);
    }
};
PSI_DEV double philox_uniform(const Philox& ph, unsigned int c0, unsigned int c1, unsigned int c2, unsigned int c3) {
    unsigned int r[4];
    ph(c0, c1, c2, c3, r);
    // 53-bit uniform in [0,1)
    const unsigned long long m = ((unsigned long long)r[0] << 21) ^ (unsigned long long)(r[1] >> 11);
    return (double)(m & ((1ull << 53) - 1)) * (1.0 / 9007199254740992.0);
}

// ---- block reductions / scan (blockDim.x == 128 => 4 warps) ---------------------------------------
PSI_DEV double block_sum(double v, double* smem4) {
#ifndef PSI_HOST_SIM      // (the host build runs one thread per CTA: there are no other lanes to add)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
#endif
    const int w = threadIdx.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) smem4[w] = v;
    __syncthreads();
    double s = 0.0;
    const int nw = (blockDim.x + 31) >> 5;
    for (int i = 0; i < nw; ++i) s += smem4[i];
    return s;
}

// sqrt(x) for x > 0 to ~1e-12 relative: MUFU.RSQ64H seed + one Newton step (the noise amplitude
// sqrt(dt) does not need the last bits; a full FP64 sqrt is ~25 instructions on the bound pipe)
PSI_DEV double sqrt_fast(double x) {
    double r;
#ifdef PSI_HOST_SIM
    r = 1.0 / sqrt(x);
#else
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#endif
    const double hx = 0.5 * x;
    r = fma(r, fma(-hx * r, r, 0.5), r);       // r <- r (1.5 - 0.5 x r^2)
    return x * r;
}

// drift rule of the reference: every infusion with start <= t <= start + duration (sde/mod.rs:124-133); `until` is the
// first start / end after t, up to which the set cannot change
template <int NR>
struct RateWindow { double rate[NR]; double until; };
template <int NR>
__device__ __noinline__ RateWindow<NR> sde_rate_window(const InfRec* p, int n, double t) {
    RateWindow<NR> w;
    w.until = psi_inf();
#pragma unroll
    for (int k = 0; k < NR; ++k) w.rate[k] = 0.0;
    for (int i = 0; i < n; ++i) {
        double s, d, a; int input;
        load_inf(p + i, s, d, a, input);
        const double e = s + d;
        if (t >= s && t <= e) add_rate<NR>(w.rate, input, a);
        if (s > t) w.until = fmin(w.until, s);
        if (e >= t) w.until = fmin(w.until, e);
    }
    return w;
}

template <class M>
struct SdeStep {
    static constexpr int NS = M::NSTATE;
    static constexpr int NR = AtLeast1<M::NROUTE>::v;
    PairCtx<M>& c;
    InfRange inf;
    // the active infusion rates are piecewise constant: c.rate is valid for rate_from <= t < rate_until
    double rate_from, rate_until;
    PSI_DEV SdeStep(PairCtx<M>& cc, InfRange r) : c(cc), inf(r), rate_from(1.0), rate_until(0.0) {}
    PSI_DEV void invalidate_rates() { rate_from = 1.0; rate_until = 0.0; }
    // The window test is all the hot loop sees (1 % of the evaluations on C5 leave it); the recomputation is an out-of-line
    // call so that the compiler cannot fold its first statements into selects executed by every evaluation.
    PSI_DEV void rates_at(double t) {
        if (t >= rate_from && t < rate_until) return;
        const RateWindow<NR> w = sde_rate_window<NR>(inf.p, inf.n, t);
#pragma unroll
        for (int k = 0; k < NR; ++k) c.rate[k] = w.rate[k];
        rate_from = t;
        rate_until = w.until;     // at t == until the set is recomputed (closed interval ends)
    }
    // drift + diffusion at (t, x): derive/covariates refreshed at absolute t (native.rs:2330-2420)
    // CHECK = false: the caller has established that t lies in the current rate window
    template <bool CHECK = true>
    PSI_DEV void eval(double t, const double* x, double* dx, double* g) {
        if constexpr (CHECK) rates_at(t);
        if constexpr (M::RHS_USES_COV) fill_cov<M>(*c.pop, c.occ, t, c.cov);
        if constexpr (M::HAS_DERIVE && M::DERIVE_DEPS != 0) M::derive(t, x, c.p, c.cov, c.rate, c.d);
        M::drift(t, x, c.p, c.cov, c.rate, c.d, dx);
#pragma unroll
        for (int k = 0; k < NS; ++k) g[k] = 0.0;
        M::diffusion(t, x, c.p, c.cov, c.rate, c.d, g);
    }
    // em.rs:104-120
    template <bool CHECK = true>
    PSI_DEV void em_step(double t, double dt, double sqdt, double* x, const double* z) {
        double dx[NS], g[NS];
        eval<CHECK>(t, x, dx, g);
#pragma unroll
        for (int k = 0; k < NS; ++k) x[k] = fma(dx[k], dt, fma(g[k] * z[k], sqdt, x[k]));
    }
    // The full step y1 and the first half step of y2 both start from (t, x): one drift / diffusion evaluation serves
    // both (the reference evaluates it twice with identical arguments, em.rs:134-150).
    template <bool CHECK = true>
    PSI_DEV void em_first(double t, double dt, double sq, double sqh, const double* x, double* y1, double* y2, const double* z) {
        double dx[NS], g[NS];
        eval<CHECK>(t, x, dx, g);
        const double hdt = dt * 0.5;
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            y1[k] = fma(dx[k], dt, fma(g[k] * z[k], sq, x[k]));
            y2[k] = fma(dx[k], hdt, fma(g[k] * z[NS + k], sqh, x[k]));
        }
    }
    // Normal bookkeeping of the reference stepper.  One attempt consumes NEED = 3*NSTATE normals (three INDEPENDENT draws
    // per state, em.rs:104-120) and a Philox block yields 4, so a particle's attempts run in a cycle of PERIOD attempts
    // over which whole blocks are used up exactly; within the cycle the leftover normals of a block are carried into the
    // next attempt (NSTATE = 1: 4 attempts per 3 blocks instead of 4 — a quarter of the generator work, which is what
    // bounds this kernel).  All indices below are compile-time, so the carry stays in registers.
    static constexpr int NEED = 3 * NS;
    static constexpr int PERIOD = (NEED % 4 == 0) ? 1 : ((NEED % 2 == 0) ? 2 : 4);
    // em.rs:134-167, one attempt of one particle at cycle position PHASE
    template <int PHASE>
    PSI_DEV void attempt(double tf, double* x, double& t, double& dt, NormalStream& rng, double* carry, Counters& cnt) {
        constexpr int L = 4 * ((NEED * PHASE + 3) / 4) - NEED * PHASE;   // normals left over by the attempts before this one
        constexpr int NB = (NEED - L + 3) / 4, L2 = L + 4 * NB - NEED;    // blocks to draw now, normals left over after
        double z[L + 4 * NB + 1];
#pragma unroll
        for (int i = 0; i < L; ++i) z[i] = carry[i];
#pragma unroll
        for (int b = 0; b < NB; ++b) rng.block4(z + L + 4 * b);
#pragma unroll
        for (int i = 0; i < L2; ++i) carry[i] = z[NEED + i];
        double y1[NS], y2[NS];
        const double sq = sqrt_fast(dt), sqh = sq * 0.70710678118654752;
        // full step and first half step share drift / diffusion at (t, x).  One window test covers both evaluation times
        // (99 % of the attempts on C5); outside it each evaluation looks its rates up
        const double th = fma(dt, 0.5, t);
        if (t >= rate_from && th < rate_until) {
            em_first<false>(t, dt, sq, sqh, x, y1, y2, z);
            em_step<false>(th, dt * 0.5, sqh, y2, z + 2 * NS);
        } else {
            em_first(t, dt, sq, sqh, x, y1, y2, z);
            em_step(th, dt * 0.5, sqh, y2, z + 2 * NS);
        }
        cnt.evals += 3;                                  // algorithmic count (the reference evaluates the pair three times)
        // err only steers dt: the weight 1/tol uses the one-MUFU reciprocal and the new step the FP32 rsqrt; err, nd and
        // tf - t are compared on the integer pipe (err >= 0, nd >= 0; tf - t < 0 only on the overshooting first attempt)
        double err = fabs(y1[0] - y2[0]) * rcp_approx(fma(1e-2, fabs(x[0]), 1e-2));
#pragma unroll
        for (int q = 1; q < NS; ++q) {
            const double tol = fma(1e-2, fabs(x[q]), 1e-2);
            err = max_pos(err, fabs(y1[q] - y2[q]) * rcp_approx(tol));
        }
        double nd = dt * (double)(0.9f * rsqrt_ftz((float)err));
        nd = min_pos(max_pos(nd, 1e-6), 0.1);
        if (err <= 1.0) {
            cnt.steps++;
            t += dt;
#pragma unroll
            for (int q = 0; q < NS; ++q) x[q] = y2[q];
            dt = min_pos(nd, tf - t);
        } else {
            cnt.rejected++;
            dt = nd;
        }
    }
    template <int PHASE>
    PSI_DEV void cycle(double tf, double* x, double& t, double& dt, int& guard, bool& active, int k, double* buf, int np,
                       NormalStream& rng, double* carry, Counters& cnt) {
        if constexpr (PHASE < PERIOD) {
            if (active) {
                attempt<PHASE>(tf, x, t, dt, rng, carry, cnt);
                if (!(t < tf) || ++guard > 4000000) {
#pragma unroll
                    for (int s = 0; s < NS; ++s) buf[(long long)s * np + k] = x[s];
                    active = false;
                }
            }
            cycle<PHASE + 1>(tf, x, t, dt, guard, active, k, buf, np, rng, carry, cnt);
        }
    }
    // The reference stepper over ALL particles of the CTA with lane-level dynamic scheduling: the number of
    // attempts per particle is random (the error estimate is noise-dominated), so a static particle->lane
    // map leaves a quarter of the lanes idle (ncu: 24.0 active threads per instruction).  Here every lane
    // runs cycles of PERIOD attempts and fetches the next particle from a CTA-wide counter at the first cycle
    // boundary after its own finishes — every lane of a warp is then at the same cycle position, so the block
    // generation of a position is skipped by the whole warp where the carry covers it; a particle that ends
    // mid-cycle idles its lane for at most PERIOD-1 attempts out of the hundreds an interval takes.  Results do
    // not depend on the schedule: the Philox stream, and the position in it, are functions of the particle alone.
    PSI_DEV void solve_reference_all(double t0, double tf, double* buf, int np, int* next, NormalStream& rng,
                                     unsigned int seq, unsigned int pair, Counters& cnt) {
        double x[NS];
        double carry[4];
        double t = t0, dt = 0.1;
        int k = -1, guard = 0;
        bool active = false;
        while (true) {
            if (!active) {
                k = atomicAdd(next, 1);
                if (k >= np) break;
#pragma unroll
                for (int s = 0; s < NS; ++s) x[s] = buf[(long long)s * np + k];
                rng.reset((unsigned int)k, seq, pair);
                t = t0; dt = 0.1; guard = 0; active = true;
            }
            cycle<0>(tf, x, t, dt, guard, active, k, buf, np, rng, carry, cnt);
        }
    }
    PSI_DEV void solve_fixed(double t0, double tf, double hmax, double* x, NormalStream& rng, Counters& cnt) {
        const int n = (int)fmax(1.0, ceil((tf - t0) / hmax - 1e-9));
        const double dt = (tf - t0) / n, sq = sqrt(dt);
        // one Philox block yields 4 normals: take G = 4 / NS steps per block when NS < 4
        constexpr int G = (NS >= 4) ? 1 : 4 / NS;
        for (int i = 0; i < n; i += G) {
            double z[4 * ((G * NS + 3) / 4)];
            rng.template fill<G * NS>(z);
#pragma unroll
            for (int q = 0; q < G; ++q)
                if (i + q < n) em_step(fma((double)(i + q), dt, t0), dt, sq, x, z + q * NS);
        }
        cnt.steps += n; cnt.evals += n;
    }
};

// Workspace per CTA (shared memory when it fits, else global scratch): 2 state buffers [NSTATE][np] that swap roles at
// every resampling, the weights in the first np slots of the idle one, and anc[np] (launch_geometry.hpp)
template <class M>
__device__ __forceinline__ void psi_sde_kernel_body(const PopView& pop, const double* __restrict__ spp, long long ncols,
                                                    long long spp_ld, const RunOpts& opt, const OutView& out) {
    if constexpr (M::KIND == 2) {
        constexpr int NS = M::NSTATE;
        constexpr int NR = AtLeast1<M::NROUTE>::v;
        __shared__ double red[8];
        __shared__ int next_particle;
        const int np = opt.nparticles;
        const int tid = threadIdx.x, B = blockDim.x;
        // particle workspace: the CTA's dynamic shared memory when it fits (1 000 particles x 1 state = 28 KB), else a slab of
        // global scratch (L2 resident)
#ifdef PSI_HOST_SIM
        double* sde_smem = nullptr;
#else
        extern __shared__ double sde_smem[];
#endif
        double* ws = out.scratch_in_smem ? sde_smem : out.scratch + (long long)blockIdx.x * out.scratch_stride;
        double* bufA = ws;
        double* bufB = ws + (long long)NS * np;
        int* anc = reinterpret_cast<int*>(ws + 2ll * NS * np);
        Counters cnt;
        // log_likelihood_batch (likelihood/mod.rs:119-177): CTA q evaluates subject q with parameter row q; the score is the
        // residual-error likelihood of the particle-MEAN predictions (estimate_predictions of an SDE, sde/mod.rs:387-433)
        const bool diag = opt.diagonal != 0;
        const long long npairs = diag ? (ncols < pop.nsub ? ncols : (long long)pop.nsub) : (long long)pop.nsub * ncols;
        // Pairs differ in cost (the adaptive stepper's attempt count depends on the parameters): after its first pair a CTA
        // draws the next one from a launch-wide ticket instead of striding, so no CTA ends long after the others.  psi does
        // not depend on the assignment (streams are keyed by the global pair).
        __shared__ long long next_pair;
        long long lpair = blockIdx.x;
        while (lpair < npairs) {
            const int subj = diag ? (int)lpair : (int)(lpair % pop.nsub);
            const long long j = diag ? lpair : lpair / pop.nsub;
            // the random streams are keyed by the GLOBAL pair index, so psi does not depend on how the columns are
            // sharded over GPUs or chunked by the host call
            const long long pair = (long long)subj + (j + out.col_base) * (long long)pop.nsub;
            PairCtx<M> c;
            c.pop = &pop;
#pragma unroll
            for (int k = 0; k < M::NP; ++k) c.p[k] = __ldg(spp + (long long)k * spp_ld + j);
            M::prologue(c.p);
            Philox ph{(unsigned int)(opt.seed & 0xffffffffull) ^ (unsigned int)(pair >> 32), (unsigned int)(opt.seed >> 32)};
            NormalStream rng; rng.ph = ph; rng.fp64 = opt.sde_normals;
            int status = ST_OK;
            double ll = 0.0;
            unsigned int seq = 0;
            if constexpr (M::HAS_DERIVE && M::DERIVE_DEPS == 0) {
                double zx[NS];
#pragma unroll
                for (int k = 0; k < NS; ++k) zx[k] = 0.0;
                c.zero_rate();
                M::derive(0.0, zx, c.p, c.cov, c.rate, c.d);
            } else {
#pragma unroll
                for (int k = 0; k < AtLeast1<M::NDER>::v; ++k) c.d[k] = 0.0;
            }
            double* cur_buf = bufA;
            double* alt_buf = bufB;
            const int occ0 = __ldg(pop.occ_offsets + subj), occ1 = __ldg(pop.occ_offsets + subj + 1);
            for (int occ = occ0; occ < occ1; ++occ) {
                c.occ = occ;
                const InfRange inf = occ_infusions(pop, occ);
                SdeStep<M> stepper(c, inf);
                double x0[NS];
#pragma unroll
                for (int k = 0; k < NS; ++k) x0[k] = 0.0;
                if constexpr (M::HAS_INIT) {
                    if (__ldg(pop.occ_index + occ) == 0) {
                        c.zero_rate();
                        c.refresh(0.0, x0);
                        M::init(0.0, x0, c.p, c.cov, c.rate, c.d, x0);
                    }
                }
                __syncthreads();
                for (int k = tid; k < np; k += B)
#pragma unroll
                    for (int s = 0; s < NS; ++s) cur_buf[(long long)s * np + k] = x0[s];
                auto lag_of = [&](int route, double tb) -> double {
                    if constexpr (M::HAS_LAG) {
                        double zx[NS];
#pragma unroll
                        for (int k = 0; k < NS; ++k) zx[k] = 0.0;
                        c.zero_rate();
                        c.refresh(tb, zx);
                        return M::lag(route, tb, zx, c.p, c.cov, c.rate, c.d);
                    } else {
                        return 0.0;
                    }
                };
                EventCursor<M, decltype(lag_of)> cur(pop, occ, lag_of);
                EventRec e;
                double te;
                bool have = cur.next(e, te);
                while (have) {
                    const int kind = ev_kind(e.meta);
                    if (kind == EV_BOLUS) {
                        const int route = ev_index(e.meta);
                        double amount = e.a;
                        if constexpr (M::HAS_FA) {
                            double zx[NS];
#pragma unroll
                            for (int k = 0; k < NS; ++k) zx[k] = 0.0;
                            c.zero_rate();
                            c.refresh(te, zx);
                            const double fa = M::fa(route, te, zx, c.p, c.cov, c.rate, c.d);
                            if (fa != 1.0) amount *= fa;
                        }
                        const int dest = M::bolus_dest(route);
                        if (dest < 0) { if (status == ST_OK) status = ST_UNSUPPORTED_INPUT_ROUTE_KIND; }
                        else for (int k = tid; k < np; k += B) cur_buf[(long long)dest * np + k] += amount;
                    } else if (kind == EV_OBS) {
                        const bool pf = opt.want_ll && !diag && (opt.sde_mode == SDE_PARTICLE_FILTER) && ev_has_value(e.meta);
                        double ysum = 0.0, qsum = 0.0;
                        double* qv = alt_buf;            // weights, then their running sum, until the ancestors are known
                        int lstat = ST_OK;
                        for (int k = tid; k < np; k += B) {
                            double x[NS];
#pragma unroll
                            for (int s = 0; s < NS; ++s) x[s] = cur_buf[(long long)s * np + k];
                            active_rates<NR>(inf, te, c.rate);
                            stepper.invalidate_rates();
                            c.refresh(te, x);
                            double y[AtLeast1<M::NOUT>::v];
#pragma unroll
                            for (int q = 0; q < AtLeast1<M::NOUT>::v; ++q) y[q] = 0.0;
                            M::outputs(te, x, c.p, c.cov, c.rate, c.d, y);
                            const double yp = pick<AtLeast1<M::NOUT>::v>(y, ev_index(e.meta));
                            ysum += yp;
                            if (pf) {
                                const double q = exp(obs_log_likelihood(e, yp, lstat));
                                qv[k] = q;
                                qsum += q;
                            }
                        }
                        const double ymean = block_sum(ysum, red) / (double)np;
                        if (out.pred && opt.want_pred && e.obs_row >= 0 && tid == 0) out.pred[(long long)e.obs_row * out.ld_pred + j] = ymean;
                        if (pf) {
                            lstat = __syncthreads_or(lstat != ST_OK) ? ST_NON_FINITE_LIKELIHOOD : ST_OK;   // reference panics
                            if (lstat != ST_OK && status == ST_OK) status = lstat;
                            const double sum_q = block_sum(qsum, red);
                            ll += log(sum_q / (double)np);
                            // inclusive scan of w = q / sum_q over particle index, tiles of B
                            __shared__ double carry_s;
                            __shared__ double wtot[4];
                            if (tid == 0) carry_s = 0.0;
                            __syncthreads();
                            for (int base = 0; base < np; base += B) {
                                const int k = base + tid;
                                double v = (k < np) ? qv[k] / sum_q : 0.0;
#pragma unroll
                                for (int off = 1; off < 32; off <<= 1) {
                                    const double n = __shfl_up_sync(0xffffffffu, v, off);
                                    if ((tid & 31) >= off) v += n;
                                }
                                if ((tid & 31) == 31) wtot[tid >> 5] = v;
                                __syncthreads();
                                double pre = carry_s;
                                for (int w = 0; w < (tid >> 5); ++w) pre += wtot[w];
                                v += pre;
                                if (k < np) qv[k] = v;
                                __syncthreads();
                                if (tid == B - 1) carry_s = v;
                                __syncthreads();
                            }
                            // ancestors
                            ++seq;
                            for (int k = tid; k < np; k += B) {
                                const double u = ((double)k + philox_uniform(ph, (unsigned int)k, 0xffffffffu, seq, (unsigned int)pair)) / (double)np;
                                int lo = 0, hi = np - 1;          // first index with qv[idx] >= u
                                while (lo < hi) {
                                    const int mid = (lo + hi) >> 1;
                                    if (qv[mid] < u) lo = mid + 1; else hi = mid;
                                }
                                anc[k] = lo;
                            }
                            __syncthreads();
                            for (int k = tid; k < np; k += B) {
                                const int a = anc[k];
#pragma unroll
                                for (int s = 0; s < NS; ++s) alt_buf[(long long)s * np + k] = cur_buf[(long long)s * np + a];
                            }
                            __syncthreads();
                            double* t2 = cur_buf; cur_buf = alt_buf; alt_buf = t2;
                        } else if (opt.want_ll && ev_has_value(e.meta) && diag) {
                            ll += resid_log_likelihood(opt, ev_index(e.meta), e.a, ymean);
                        } else if (opt.want_ll && ev_has_value(e.meta) && opt.sde_mode == SDE_MEAN_PREDICTION) {
                            ll += obs_log_likelihood(e, ymean, status);
                        }
                    }
                    EventRec en;
                    double tn;
                    have = cur.next(en, tn);
                    if (have && te != tn) {
                        ++seq;
                        __syncthreads();
                        stepper.invalidate_rates();     // lag / fa / outputs above reuse c.rate as scratch
                        if (opt.em_mode == EM_FIXED_STEP) {
                            for (int k = tid; k < np; k += B) {
                                double x[NS];
#pragma unroll
                                for (int s = 0; s < NS; ++s) x[s] = cur_buf[(long long)s * np + k];
                                rng.reset((unsigned int)k, seq, (unsigned int)pair);
                                stepper.solve_fixed(te, tn, opt.em_dt, x, rng, cnt);
#pragma unroll
                                for (int s = 0; s < NS; ++s) cur_buf[(long long)s * np + k] = x[s];
                            }
                        } else {
                            if (tid == 0) next_particle = 0;
                            __syncthreads();
                            stepper.solve_reference_all(te, tn, cur_buf, np, &next_particle, rng, seq, (unsigned int)pair, cnt);
                        }
                        __syncthreads();
                    }
                    e = en;
                    te = tn;
                }
            }
            if (tid == 0 && diag) {
                if (out.ll) out.ll[lpair] = (status != ST_OK) ? -psi_inf() : ll;      // a failed simulation scores -inf (mod.rs:134-137)
            } else if (tid == 0) {
                if (status != ST_OK) { ll = psi_nan(); report_error(out, pair, status); }
                if (out.npeers > 0) {
                    const long long at = (long long)subj + (j + out.col_base) * out.ld_ll;
                    for (int r = 0; r < out.npeers; ++r) out.ll_peers[r][at] = ll;
                } else if (out.ll) {
                    out.ll[(long long)subj + j * out.ld_ll] = ll;
                }
            }
            if (tid == 0) next_pair = out.pair_ticket ? (long long)gridDim.x + (long long)atomicAdd(out.pair_ticket, 1ull) : lpair + gridDim.x;
            __syncthreads();
            lpair = next_pair;
            __syncthreads();               // (a subject without occasions has no other barrier before the next draw)
            flush_counters(out, cnt);      // per pair: the 32-bit per-thread counters would wrap over a long grid-stride loop
            cnt = Counters{};
        }
    }
}

}  // namespace psi
