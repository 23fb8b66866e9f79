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
#include "psi_common.cuh"

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

// Per-particle normal stream: Box-Muller on 32-bit uniforms in FP32 (SFU), 4 normals per block.
struct NormalStream {
    Philox ph;
    unsigned int c0, c2, c3;   // particle slot, interval sequence, pair
    unsigned int ctr;          // draw-block counter within the interval
    float buf[4];
    int have;
    PSI_DEV void reset(unsigned int particle, unsigned int seq, unsigned int pair) {
        c0 = particle; c2 = seq; c3 = pair; ctr = 0; have = 0;
    }
    PSI_DEV double next() {
        if (have == 0) {
            unsigned int r[4];
            ph(c0, ctr++, c2, c3, r);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float u1 = ((float)(r[2 * i] >> 8) + 0.5f) * (1.0f / 16777216.0f);     // (0,1)
                const float u2 = ((float)(r[2 * i + 1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
                const float rad = sqrtf(-2.0f * __logf(u1));
                float s, c;
                __sincosf(6.2831853071795865f * u2, &s, &c);
                buf[2 * i] = rad * c;
                buf[2 * i + 1] = rad * s;
            }
            have = 4;
        }
        return (double)buf[--have];
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
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    const int w = threadIdx.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) smem4[w] = v;
    __syncthreads();
    double s = 0.0;
    const int nw = (blockDim.x + 31) >> 5;
    for (int i = 0; i < nw; ++i) s += smem4[i];
    return s;
}

template <class M>
struct SdeStep {
    PairCtx<M>& c;
    InfRange inf;
    // drift + diffusion at (t, x): derive/covariates refreshed at absolute t (native.rs:2330-2420)
    PSI_DEV void eval(double t, const double* x, double* dx, double* g) {
        constexpr int NR = AtLeast1<M::NROUTE>::v;
        active_rates<NR>(inf, t, c.rate);
        if constexpr (M::RHS_USES_COV) fill_cov<M>(*c.pop, c.occ, t, c.cov);
        if constexpr (M::HAS_DERIVE && M::DERIVE_DEPS != 0) M::derive(t, x, c.p, c.cov, c.rate, c.d);
        M::drift(t, x, c.p, c.cov, c.rate, c.d, dx);
#pragma unroll
        for (int k = 0; k < M::NSTATE; ++k) g[k] = 0.0;
        M::diffusion(t, x, c.p, c.cov, c.rate, c.d, g);
    }
    // em.rs:104-120
    PSI_DEV void em_step(double t, double dt, double sqdt, double* x, NormalStream& rng) {
        double dx[M::NSTATE], g[M::NSTATE];
        eval(t, x, dx, g);
#pragma unroll
        for (int k = 0; k < M::NSTATE; ++k) x[k] += dx[k] * dt + g[k] * rng.next() * sqdt;
    }
    // em.rs:134-167
    PSI_DEV void solve_reference(double t0, double tf, double* x, NormalStream& rng, Counters& cnt) {
        double t = t0, dt = 0.1;
        int guard = 0;
        while (t < tf) {
            if (++guard > 4000000) break;
            double y1[M::NSTATE], y2[M::NSTATE];
#pragma unroll
            for (int k = 0; k < M::NSTATE; ++k) { y1[k] = x[k]; y2[k] = x[k]; }
            const double sq = sqrt(dt), sqh = sqrt(dt / 2.0);
            em_step(t, dt, sq, y1, rng);
            em_step(t, dt / 2.0, sqh, y2, rng);
            em_step(t + dt / 2.0, dt / 2.0, sqh, y2, rng);
            cnt.evals += 3;
            double err = 0.0;
#pragma unroll
            for (int k = 0; k < M::NSTATE; ++k) {
                const double tol = 1e-2 + 1e-2 * fabs(x[k]);
                err = fmax(err, fabs(y1[k] - y2[k]) / tol);
            }
            double nd = dt * 0.9 * sqrt(1.0 / err);
            nd = fmin(fmax(nd, 1e-6), 0.1);
            if (err <= 1.0) {
                cnt.steps++;
                t += dt;
#pragma unroll
                for (int k = 0; k < M::NSTATE; ++k) x[k] = y2[k];
                dt = fmin(nd, tf - t);
            } else {
                cnt.rejected++;
                dt = nd;
            }
        }
    }
    PSI_DEV void solve_fixed(double t0, double tf, double hmax, double* x, NormalStream& rng, Counters& cnt) {
        const int n = (int)fmax(1.0, ceil((tf - t0) / hmax - 1e-9));
        const double dt = (tf - t0) / n, sq = sqrt(dt);
        for (int i = 0; i < n; ++i) em_step(t0 + i * dt, dt, sq, x, rng);
        cnt.steps += n; cnt.evals += n;
    }
};

// Workspace per CTA (global memory, L2 resident): 2 state buffers [NSTATE][np], q[np], anc[np]
template <class M>
__device__ __forceinline__ void psi_sde_kernel_body(const PopView& pop, const double* __restrict__ spp, long long ncols,
                                                    long long spp_ld, const RunOpts& opt, const OutView& out) {
    if constexpr (M::KIND == 2) {
        constexpr int NS = M::NSTATE;
        constexpr int NR = AtLeast1<M::NROUTE>::v;
        __shared__ double red[8];
        const int np = opt.nparticles;
        const int tid = threadIdx.x, B = blockDim.x;
        double* ws = out.scratch + (long long)blockIdx.x * out.scratch_stride;
        double* bufA = ws;
        double* bufB = ws + (long long)NS * np;
        double* qv = ws + 2ll * NS * np;
        int* anc = reinterpret_cast<int*>(ws + 2ll * NS * np + np);
        Counters cnt;
        const long long npairs = (long long)pop.nsub * ncols;
        for (long long pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
            const int subj = (int)(pair % pop.nsub);
            const long long j = pair / pop.nsub;
            PairCtx<M> c;
            c.pop = &pop;
#pragma unroll
            for (int k = 0; k < M::NP; ++k) c.p[k] = __ldg(spp + (long long)k * spp_ld + j);
            Philox ph{(unsigned int)(opt.seed & 0xffffffffull) ^ (unsigned int)(pair >> 32), (unsigned int)(opt.seed >> 32)};
            NormalStream rng; rng.ph = ph;
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
                SdeStep<M> stepper{c, inf};
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
                        const bool pf = opt.want_ll && (opt.sde_mode == SDE_PARTICLE_FILTER) && ev_has_value(e.meta);
                        double ysum = 0.0, qsum = 0.0;
                        int lstat = ST_OK;
                        for (int k = tid; k < np; k += B) {
                            double x[NS];
#pragma unroll
                            for (int s = 0; s < NS; ++s) x[s] = cur_buf[(long long)s * np + k];
                            active_rates<NR>(inf, te, c.rate);
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
                        for (int k = tid; k < np; k += B) {
                            double x[NS];
#pragma unroll
                            for (int s = 0; s < NS; ++s) x[s] = cur_buf[(long long)s * np + k];
                            rng.reset((unsigned int)k, seq, (unsigned int)pair);
                            if (opt.em_mode == EM_FIXED_STEP) stepper.solve_fixed(te, tn, opt.em_dt, x, rng, cnt);
                            else stepper.solve_reference(te, tn, x, rng, cnt);
#pragma unroll
                            for (int s = 0; s < NS; ++s) cur_buf[(long long)s * np + k] = x[s];
                        }
                        __syncthreads();
                    }
                    e = en;
                    te = tn;
                }
            }
            if (tid == 0) {
                if (status != ST_OK) { ll = psi_nan(); report_error(out, pair, status); }
                if (out.ll) out.ll[(long long)subj + j * out.ld_ll] = ll;
            }
            __syncthreads();
        }
        flush_counters(out, cnt);
    }
}

}  // namespace psi
