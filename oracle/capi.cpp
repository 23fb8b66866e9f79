// capi.cpp — ctypes-facing C API of the CPU oracle (TEST INFRASTRUCTURE ONLY; see pharmsol_oracle.hpp).
//
// The psi driver restates likelihood/matrix.rs:52-106: F-order (column-major) output of shape
// (nsub, nspp); parallel over subject rows (rayon -> OpenMP, matrix.rs:79-83), serial over support
// points (:86-95); the first error aborts the whole matrix (:96-104).
#include <atomic>
#include <chrono>
#include <cstdio>
#include <memory>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "pharmsol_oracle.hpp"

namespace orc { Model make_model(const std::string& name); }

using namespace orc;

static thread_local std::string g_last_error;

struct Data { std::vector<Subject> subjects; };

extern "C" {

const char* orc_last_error() { return g_last_error.c_str(); }
int orc_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// ---- subjects ---------------------------------------------------------------------------------
void* orc_sb_new(const char* id) { return new SubjectBuilder(id); }
void orc_sb_bolus(void* b, double t, double amt, const char* input) { ((SubjectBuilder*)b)->bolus(t, amt, input); }
void orc_sb_infusion(void* b, double t, double amt, const char* input, double dur) { ((SubjectBuilder*)b)->infusion(t, amt, input, dur); }
void orc_sb_observation(void* b, double t, double v, const char* outeq) { ((SubjectBuilder*)b)->observation(t, v, outeq); }
void orc_sb_censored_observation(void* b, double t, double v, const char* outeq, int cens) {
    ((SubjectBuilder*)b)->censored_observation(t, v, outeq, (Censor)cens);
}
void orc_sb_missing_observation(void* b, double t, const char* outeq) { ((SubjectBuilder*)b)->missing_observation(t, outeq); }
void orc_sb_observation_with_error(void* b, double t, double v, const char* outeq, double c0, double c1, double c2, double c3, int cens) {
    ((SubjectBuilder*)b)->observation_with_error(t, v, outeq, ErrorPoly{c0, c1, c2, c3}, (Censor)cens);
}
void orc_sb_repeat(void* b, long n, double delta) { ((SubjectBuilder*)b)->repeat((size_t)n, delta); }
void orc_sb_reset(void* b) { ((SubjectBuilder*)b)->reset(); }
void orc_sb_covariate(void* b, const char* name, double t, double v) { ((SubjectBuilder*)b)->covariate(name, t, v); }
void* orc_sb_build(void* b) {
    auto* sb = (SubjectBuilder*)b;
    auto* s = new Subject(sb->build());
    delete sb;
    return s;
}
void orc_subject_set_covariate_fixed(void* s, int occasion, const char* name, int fixed) {
    auto& occ = ((Subject*)s)->occasions.at((size_t)occasion);
    auto it = occ.covariates.map.find(name);
    if (it != occ.covariates.map.end()) { it->second.fixed = fixed != 0; it->second.build_segments(); }
}
void orc_subject_free(void* s) { delete (Subject*)s; }
// flattened view of the (sorted) events of one occasion, for event-order tests
long orc_subject_n_occasions(void* s) { return (long)((Subject*)s)->occasions.size(); }
long orc_subject_n_events(void* s, int occ) { return (long)((Subject*)s)->occasions.at((size_t)occ).events.size(); }
void orc_subject_event(void* s, int occ, long i, int* kind, double* time, double* amount) {
    const Event& e = ((Subject*)s)->occasions.at((size_t)occ).events.at((size_t)i);
    *kind = (int)e.kind; *time = e.time; *amount = e.kind == EventKind::Observation ? e.value : e.amount;
}
int orc_covariate_interpolate(void* s, int occ, const char* name, double t, double* out) {
    const Covariate* c = ((Subject*)s)->occasions.at((size_t)occ).covariates.get_covariate(name);
    if (!c) return 1;
    return c->interpolate(t, *out) ? 0 : 2;
}

void* orc_data_new() { return new Data(); }
void orc_data_add(void* d, void* s) { ((Data*)d)->subjects.push_back(*(Subject*)s); }
long orc_data_len(void* d) { return (long)((Data*)d)->subjects.size(); }
void orc_data_free(void* d) { delete (Data*)d; }

// ---- models -----------------------------------------------------------------------------------
void* orc_model_new(const char* name) {
    try { return new Model(make_model(name)); }
    catch (const std::exception& e) { g_last_error = e.what(); return nullptr; }
}
void orc_model_free(void* m) { delete (Model*)m; }
void orc_model_set_solver(void* m, int solver, double rtol, double atol) {
    auto* mm = (Model*)m; mm->solver = (OdeSolver)solver; mm->rtol = rtol; mm->atol = atol;
}
void orc_model_set_particles(void* m, int n) { ((Model*)m)->nparticles = n; }
int orc_model_kind(void* m) { return (int)((Model*)m)->kind; }

// ---- error models -----------------------------------------------------------------------------
void* orc_em_new(int nout) { auto* e = new AssayErrorModels(); e->models.resize((size_t)nout); return e; }
void orc_em_set(void* e, int outeq, int kind, double factor, double c0, double c1, double c2, double c3) {
    auto& m = ((AssayErrorModels*)e)->models.at((size_t)outeq);
    m.kind = (ErrKind)kind; m.factor = factor; m.poly = ErrorPoly{c0, c1, c2, c3};
}
void orc_em_free(void* e) { delete (AssayErrorModels*)e; }

// ---- scalar helpers for anchor tests ------------------------------------------------------------
double orc_lognormpdf(double o, double p, double s) { return lognormpdf(o, p, s); }
int orc_lognormcdf(double o, double p, double s, double* out) {
    try { *out = lognormcdf(o, p, s); return 0; } catch (const Error& e) { g_last_error = e.what(); return e.code; }
}
int orc_lognormccdf(double o, double p, double s, double* out) {
    try { *out = lognormccdf(o, p, s); return 0; } catch (const Error& e) { g_last_error = e.what(); return e.code; }
}
int orc_kernel_step(const char* kernel, const double* x, int nx, const double* p, int np, double dt, double rate, double* out) {
    try {
        auto k = analytical_kernel_by_name(kernel);
        V xo = k(V(x, x + nx), V(p, p + np), dt, V{rate}, Covariates());
        for (int i = 0; i < nx; ++i) out[i] = xo[(size_t)i];
        return 0;
    } catch (const Error& e) { g_last_error = e.what(); return e.code; }
}

// ---- per-pair ---------------------------------------------------------------------------------
// predictions for one (subject, support point): writes up to `cap` values; *n = number of predictions.
int orc_predictions(void* model, void* subject, const double* p, int np, unsigned long long seed,
                    double* out, long cap, long* n, long* stats /* nsteps,nrej,nrhs or null */) {
    try {
        SolveStats st;
        auto preds = estimate_predictions(*(Model*)model, *(Subject*)subject, V(p, p + np), seed, &st);
        *n = (long)preds.size();
        for (long i = 0; i < *n && i < cap; ++i) out[i] = preds[(size_t)i].prediction;
        if (stats) { stats[0] = st.nsteps; stats[1] = st.nrej; stats[2] = st.nrhs; }
        return 0;
    } catch (const Error& e) { g_last_error = e.what(); return e.code; }
      catch (const std::exception& e) { g_last_error = e.what(); return OtherError; }
}
int orc_log_likelihood(void* model, void* subject, const double* p, int np, void* em, unsigned long long seed, double* out) {
    try { *out = estimate_log_likelihood_dense(*(Model*)model, *(Subject*)subject, V(p, p + np), *(AssayErrorModels*)em, seed); return 0; }
    catch (const Error& e) { g_last_error = e.what(); return e.code; }
    catch (const std::exception& e) { g_last_error = e.what(); return OtherError; }
}
int orc_sde_pf_log_likelihood(void* model, void* subject, const double* p, int np, void* em, unsigned long long seed, double* out) {
    try { *out = sde_particle_filter_log_likelihood(*(Model*)model, *(Subject*)subject, V(p, p + np), *(AssayErrorModels*)em, seed); return 0; }
    catch (const Error& e) { g_last_error = e.what(); return e.code; }
    catch (const std::exception& e) { g_last_error = e.what(); return OtherError; }
}

double orc_residual_sigma(int kind, double a, double b, double prediction) {
    ResidualErrorModel r; r.kind = (ResidualErrorModel::Kind)kind; r.a = a; r.b = b; return r.sigma(prediction);
}
double orc_residual_log_likelihood(int kind, double a, double b, double observation, double prediction) {
    ResidualErrorModel r; r.kind = (ResidualErrorModel::Kind)kind; r.a = a; r.b = b; return r.log_likelihood(observation, prediction);
}
// ---- log_likelihood_batch: likelihood/mod.rs:119-177 ------------------------------------------------
// params: row-major (nsub x np), one row per subject.  resid: nres x {kind, a, b}.  out[nsub];
// a subject whose simulation fails scores -inf (mod.rs:134-137).
int orc_log_likelihood_batch(void* model, void* data, const double* params, long nrows, int np, const double* resid, int nres,
                             double* out) {
    const Model& m = *(Model*)model;
    const Data& d = *(Data*)data;
    if (nrows != (long)d.subjects.size()) {
        g_last_error = "parameters has " + std::to_string(nrows) + " rows but there are " + std::to_string(d.subjects.size()) + " subjects";
        return OtherError;
    }
    ResidualErrorModels rem;
    for (int k = 0; k < nres; ++k) {
        ResidualErrorModel r; r.kind = (ResidualErrorModel::Kind)(int)resid[3 * k]; r.a = resid[3 * k + 1]; r.b = resid[3 * k + 2];
        rem.models.push_back(r);
    }
    for (long i = 0; i < nrows; ++i) {
        try {
            auto preds = estimate_predictions(m, d.subjects[(size_t)i], V(params + i * np, params + (i + 1) * np), 0, nullptr);
            std::vector<std::pair<size_t, std::pair<double, double>>> rows;
            for (const auto& p : preds) if (p.has_obs) rows.push_back({p.outeq, {p.observation, p.prediction}});
            out[i] = rem.total(rows);
        } catch (const std::exception&) { out[i] = -std::numeric_limits<double>::infinity(); }
    }
    return 0;
}

// ---- psi matrix: likelihood/matrix.rs:52-106 ----------------------------------------------------
// spp: row-major (nspp x np).  out: column-major (nsub x nspp).  sde_mode: 0 = mean prediction
// (what log_likelihood_matrix does), 1 = particle filter.  Returns 0 or the first error's code;
// *first_err_pair = i + j*nsub of (one of) the failing pairs.  *seconds = wall time of the loop.
int orc_log_likelihood_matrix(void* model, void* data, const double* spp, long nspp, int np, void* em,
                              double* out, int nthreads, unsigned long long seed, int sde_mode,
                              long* first_err_pair, double* seconds, long* stats) {
    const Model& m = *(Model*)model;
    const Data& d = *(Data*)data;
    const AssayErrorModels& e = *(AssayErrorModels*)em;
    const long nsub = (long)d.subjects.size();
    std::vector<V> rows((size_t)nspp);
    for (long j = 0; j < nspp; ++j) rows[(size_t)j] = V(spp + j * np, spp + (j + 1) * np);
    std::atomic<int> err{0};
    std::atomic<long> err_pair{-1};
    std::string err_msg;
    long tot_steps = 0, tot_rej = 0, tot_rhs = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : tot_steps, tot_rej, tot_rhs)
    for (long i = 0; i < nsub; ++i) {
        if (err.load(std::memory_order_relaxed)) continue;
        const Subject& s = d.subjects[(size_t)i];
        for (long j = 0; j < nspp; ++j) {
            try {
                uint64_t pair_seed = seed * 0x9E3779B97F4A7C15ULL + (uint64_t)(i + j * nsub);
                double ll;
                if (m.kind == EqnKind::SDE && sde_mode == 1) ll = sde_particle_filter_log_likelihood(m, s, rows[(size_t)j], e, pair_seed);
                else {
                    SolveStats st;
                    ll = estimate_log_likelihood_dense(m, s, rows[(size_t)j], e, pair_seed, &st);
                    tot_steps += st.nsteps; tot_rej += st.nrej; tot_rhs += st.nrhs;
                }
                out[i + j * nsub] = ll;
            } catch (const Error& ex) {
                int expected = 0;
                if (err.compare_exchange_strong(expected, ex.code)) {
                    err_pair = i + j * nsub;
#pragma omp critical
                    err_msg = ex.what();
                }
                break;
            } catch (const std::exception& ex) {
                int expected = 0;
                if (err.compare_exchange_strong(expected, (int)OtherError)) {
                    err_pair = i + j * nsub;
#pragma omp critical
                    err_msg = ex.what();
                }
                break;
            }
        }
    }
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    if (first_err_pair) *first_err_pair = err_pair.load();
    if (stats) { stats[0] = tot_steps; stats[1] = tot_rej; stats[2] = tot_rhs; }
    if (err.load()) g_last_error = err_msg;
    return err.load();
}

}  // extern "C"
