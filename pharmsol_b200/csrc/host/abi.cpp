// abi.cpp — the extern "C" surface declared in include/pharmsol_cuda.h.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <thread>
#include <vector>

#include <unistd.h>

#include "../../../include/pharmsol_cuda.h"
#include "runtime.hpp"

using namespace pharmsol;

// A context is one device (`c`) or, from pharmsol_cuda_ctx_create_multi, `c` (= device_ids[0]) plus one Ctx per further
// device: streams, events, staging buffers and the status word are per device; the caller-facing lock is c.mu.
struct pcu_ctx {
    Ctx c;
    std::vector<std::unique_ptr<Ctx>> more;
    int ndev() const { return 1 + (int)more.size(); }
    Ctx& dev(int k) { return k == 0 ? c : *more[(size_t)k - 1]; }
    bool peer_enabled = false;
    // Host-buffer calls from several host threads (rayon over subjects: `Equation: Sync`, equation/mod.rs:377) do not queue
    // behind one lock: a single-device context grows up to kMaxLanes extra Ctx bundles on the same device (own streams,
    // events, staging and status buffers), and a call takes whichever is free, so small matrices from different threads
    // overlap on the GPU instead of running one after the other.
    static constexpr int kMaxLanes = 4;
    std::vector<std::unique_ptr<Ctx>> lanes;
    std::mutex lanes_mu;
};
struct pcu_model { Model m; };
struct pcu_subject_builder { SubjectBuilder b; std::string error; explicit pcu_subject_builder(const char* id) : b(id) {} };
struct pcu_subject { Subject s; };
struct pcu_data { Data d; };
struct pcu_population { Population p; AssayErrorModels em; bool has_em = false; };

namespace {

template <class F>
int32_t guarded(F&& f) {
    try {
        return f();
    } catch (const PharmsolError& e) {
        set_last_error(e.what());
        return e.code;
    } catch (const dsl::DslError& e) {
        set_last_error(e.what());
        return PCU_ERR_COMPILE;
    } catch (const CudaError& e) {
        set_last_error(e.what());
        return PCU_ERR_CUDA;
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return PCU_ERR_OTHER;
    }
}

AssayErrorModels to_models(const pcu_error_model* ems, int32_t n) {
    AssayErrorModels out;
    for (int32_t i = 0; i < n; ++i) {
        AssayErrorModel m;
        m.kind = (ErrKind)ems[i].kind;
        m.factor = ems[i].factor;
        m.poly = ErrorPoly{ems[i].c0, ems[i].c1, ems[i].c2, ems[i].c3};
        out.models.push_back(m);
    }
    return out;
}

cudaStream_t pick_stream(Ctx& c, void* stream) { return stream ? static_cast<cudaStream_t>(stream) : c.stream; }

// `prefetched`: the error word / counters were copied into c.err_host by an async copy queued behind the kernels
// (host-buffer calls: one synchronisation instead of two round trips — it matters for 100-microsecond calls).
int32_t collect(Ctx& c, int32_t* code, int64_t* pair, bool prefetched = false) {
    unsigned long long host[5];
    if (prefetched && c.err_host) std::memcpy(host, c.err_host, sizeof host);
    else cuda_check(cudaMemcpy(host, c.err_ctr.p, sizeof host, cudaMemcpyDeviceToHost), "read error word");
    for (int k = 0; k < 4; ++k) c.last_counters[k] = host[1 + k];
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c.ev0, c.ev1) == cudaSuccess) c.last_kernel_ms = ms;
    if (host[0] == ~0ull) {
        if (code) *code = 0;
        if (pair) *pair = -1;
        return PCU_OK;
    }
    const int32_t ec = (int32_t)(host[0] & 0xff);
    // device pair index = i + j_local*nsub; make it global with the shard's first column
    const int64_t local = (int64_t)(host[0] >> 8);
    const int64_t gp = local + c.pending_first_col * c.pending_nsub;
    if (code) *code = ec;
    if (pair) *pair = gp;
    set_last_error("psi evaluation failed for pair " + std::to_string(gp) + " with status " + std::to_string(ec));
    return ec;
}

// The builder entry points return nothing (they mirror the reference's chained builder), so a failure — a NULL label,
// an absurd repeat count, out of memory — is remembered in the builder and reported by ..._build(), which then returns
// NULL with the message in pharmsol_cuda_last_error_message.  Nothing may unwind through the C boundary.
template <class F>
static void builder_op(pcu_subject_builder* b, F&& f) {
    if (!b) { set_last_error("NULL subject builder"); return; }
    if (!b->error.empty()) return;
    try { f(); }
    catch (const std::exception& e) { b->error = e.what(); }
    catch (...) { b->error = "unknown error in the subject builder"; }
}
static const char* label_or_throw(const char* s, const char* what) {
    if (!s) throw PharmsolError(PCU_ERR_INVALID_ARGUMENT, std::string("NULL ") + what + " label");
    return s;
}

}  // namespace

extern "C" {

int32_t pharmsol_cuda_abi_version(void) { return PHARMSOL_CUDA_ABI_VERSION; }

int32_t pharmsol_cuda_device_count(int32_t* n) {
    return guarded([&] {
        int c = 0;
        cuda_check(cudaGetDeviceCount(&c), "cudaGetDeviceCount");
        if (n) *n = c;
        return PCU_OK;
    });
}

static void init_ctx(Ctx& c, int32_t device) {
    cuda_check(cudaSetDevice(device), "cudaSetDevice");
    cudaDeviceProp prop;
    cuda_check(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties");
    if (prop.major < 10) throw CudaError(std::string("device `") + prop.name + "` is sm_" + std::to_string(prop.major * 10 + prop.minor) + "; this backend is built for sm_100a only");
    c.device = device;
    c.sm_count = prop.multiProcessorCount;
    cuda_check(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking), "cudaStreamCreate");
    cuda_check(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking), "cudaStreamCreate");
    for (auto& e : c.chunk_ev) cuda_check(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
    cuda_check(cudaEventCreate(&c.ev0), "cudaEventCreate");
    cuda_check(cudaEventCreate(&c.ev1), "cudaEventCreate");
    cuda_check(cudaEventCreateWithFlags(&c.reset_ev, cudaEventDisableTiming), "cudaEventCreate");
    c.err_ctr.reserve(5 * sizeof(unsigned long long));
    cuda_check(cudaMallocHost((void**)&c.err_host, 16 * sizeof(unsigned long long)), "cudaMallocHost");
    std::memset(c.err_host, 0, 16 * sizeof(unsigned long long));
    c.err_host[8] = ~0ull;
    cuda_check(cudaMallocHost((void**)&c.small_host, Ctx::kSmallIn + Ctx::kSmallOut), "cudaMallocHost");
}

struct LaneLock {
    Ctx* c = nullptr;
    std::unique_lock<std::mutex> lk;
};
// The primary bundle if it is free, else a free (or new) lane, else wait for the primary.
static LaneLock acquire_lane(pcu_ctx* ctx) {
    LaneLock r;
    r.lk = std::unique_lock<std::mutex>(ctx->c.mu, std::try_to_lock);
    if (r.lk.owns_lock()) { r.c = &ctx->c; return r; }
    if (ctx->ndev() == 1) {
        std::lock_guard<std::mutex> g(ctx->lanes_mu);
        for (auto& l : ctx->lanes) {
            std::unique_lock<std::mutex> lk(l->mu, std::try_to_lock);
            if (lk.owns_lock()) { r.c = l.get(); r.lk = std::move(lk); return r; }
        }
        if ((int)ctx->lanes.size() < pcu_ctx::kMaxLanes) {
            std::unique_ptr<Ctx> l(new Ctx());
            init_ctx(*l, ctx->c.device);
            r.lk = std::unique_lock<std::mutex>(l->mu);
            r.c = l.get();
            ctx->lanes.push_back(std::move(l));
            return r;
        }
    }
    r.lk = std::unique_lock<std::mutex>(ctx->c.mu);
    r.c = &ctx->c;
    return r;
}
static void publish_lane_stats(pcu_ctx* ctx, const Ctx& lane) {
    if (&lane == &ctx->c) return;
    std::lock_guard<std::mutex> g(ctx->lanes_mu);
    ctx->c.last_kernel_ms = lane.last_kernel_ms;
    for (int k = 0; k < 4; ++k) ctx->c.last_counters[k] = lane.last_counters[k];
}

int32_t pharmsol_cuda_ctx_create(int32_t device, pcu_ctx** out) {
    return guarded([&] {
        if (!out) return (int32_t)PCU_ERR_INVALID_ARGUMENT;
        std::unique_ptr<pcu_ctx> c(new pcu_ctx());      // ~Ctx releases whatever was created if a later step throws
        init_ctx(c->c, device);
        *out = c.release();
        return (int32_t)PCU_OK;
    });
}

// ctx_create(const int* device_ids, int n_dev, ...).  One host process drives every listed device: the
// host-buffer entry points split the support-point columns into n_dev contiguous blocks (matrix.rs:60 F-order => one
// contiguous slab of `out` per device), each device copies its slab straight into the caller's matrix, and
// pharmsol_cuda_log_likelihood_matrix_replicated leaves the whole psi resident on every device (gathered over
// NVLink).  A device may be listed more than once (two shards on one GPU: the single-GPU test of this path).
int32_t pharmsol_cuda_ctx_create_multi(const int32_t* device_ids, int32_t n_dev, pcu_ctx** out) {
    return guarded([&] {
        if (!out || !device_ids || n_dev < 1 || n_dev > 64) return (int32_t)PCU_ERR_INVALID_ARGUMENT;
        std::unique_ptr<pcu_ctx> c(new pcu_ctx());
        init_ctx(c->c, device_ids[0]);
        for (int32_t k = 1; k < n_dev; ++k) {
            c->more.emplace_back(new Ctx());
            init_ctx(*c->more.back(), device_ids[k]);
        }
        // peer access between every pair of distinct devices (NVLink / NVSwitch): peer stores from the psi kernel and
        // device-to-device pushes by the copy engines both need it
        bool all = true;
        for (int a = 0; a < n_dev; ++a) {
            for (int b = 0; b < n_dev; ++b) {
                if (device_ids[a] == device_ids[b]) continue;
                int can = 0;
                cuda_check(cudaDeviceCanAccessPeer(&can, device_ids[a], device_ids[b]), "cudaDeviceCanAccessPeer");
                if (!can) { all = false; continue; }
                cuda_check(cudaSetDevice(device_ids[a]), "cudaSetDevice");
                const cudaError_t e = cudaDeviceEnablePeerAccess(device_ids[b], 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) (void)cudaGetLastError();
                else cuda_check(e, "cudaDeviceEnablePeerAccess");
            }
        }
        c->peer_enabled = all;
        cuda_check(cudaSetDevice(device_ids[0]), "cudaSetDevice");
        *out = c.release();
        return (int32_t)PCU_OK;
    });
}
int32_t pharmsol_cuda_ctx_num_devices(pcu_ctx* ctx) { return ctx ? ctx->ndev() : 0; }
int32_t pharmsol_cuda_ctx_num_lanes(pcu_ctx* ctx) {
    if (!ctx) return 0;
    std::lock_guard<std::mutex> g(ctx->lanes_mu);
    return 1 + (int32_t)ctx->lanes.size();
}
int32_t pharmsol_cuda_ctx_device_id(pcu_ctx* ctx, int32_t k) { return (ctx && k >= 0 && k < ctx->ndev()) ? ctx->dev(k).device : -1; }
void pharmsol_cuda_ctx_destroy(pcu_ctx* ctx) { delete ctx; }
const char* pharmsol_cuda_last_error_message(void) { return last_error().c_str(); }
int64_t pharmsol_cuda_launch_count(pcu_ctx* ctx) {
    if (!ctx) return 0;
    int64_t n = 0;
    for (int k = 0; k < ctx->ndev(); ++k) n += ctx->dev(k).launches;
    std::lock_guard<std::mutex> g(ctx->lanes_mu);
    for (auto& l : ctx->lanes) n += l->launches;
    return n;
}
double pharmsol_cuda_last_kernel_ms(pcu_ctx* ctx) { return ctx ? ctx->c.last_kernel_ms : 0.0; }
int32_t pharmsol_cuda_last_counters(pcu_ctx* ctx, uint64_t out[4]) {
    if (!ctx || !out) return PCU_ERR_INVALID_ARGUMENT;
    for (int k = 0; k < 4; ++k) out[k] = ctx->c.last_counters[k];
    return PCU_OK;
}
int32_t pharmsol_cuda_host_alloc(size_t bytes, void** out) {
    return guarded([&] { cuda_check(cudaMallocHost(out, bytes), "cudaMallocHost"); return PCU_OK; });
}
int32_t pharmsol_cuda_host_free(void* p) {
    return guarded([&] { cuda_check(cudaFreeHost(p), "cudaFreeHost"); return PCU_OK; });
}

// ---- data ------------------------------------------------------------------------------------------------
pcu_subject_builder* pharmsol_subject_builder_new(const char* id) {
    try { return new pcu_subject_builder(id ? id : ""); } catch (const std::exception& e) { set_last_error(e.what()); return nullptr; }
}
void pharmsol_subject_builder_bolus(pcu_subject_builder* b, double t, double a, const char* input) {
    builder_op(b, [&] { b->b.bolus(t, a, label_or_throw(input, "input")); });
}
void pharmsol_subject_builder_infusion(pcu_subject_builder* b, double t, double a, const char* input, double dur) {
    builder_op(b, [&] { b->b.infusion(t, a, label_or_throw(input, "input"), dur); });
}
void pharmsol_subject_builder_observation(pcu_subject_builder* b, double t, double v, const char* outeq) {
    builder_op(b, [&] { b->b.observation(t, v, label_or_throw(outeq, "output")); });
}
void pharmsol_subject_builder_censored_observation(pcu_subject_builder* b, double t, double v, const char* outeq, int32_t cens) {
    builder_op(b, [&] {
        if (cens < 0 || cens > 2) throw PharmsolError(PCU_ERR_INVALID_ARGUMENT, "censoring must be 0 (none), 1 (BLOQ) or 2 (ALOQ)");
        b->b.censored_observation(t, v, label_or_throw(outeq, "output"), (Censor)cens);
    });
}
void pharmsol_subject_builder_missing_observation(pcu_subject_builder* b, double t, const char* outeq) {
    builder_op(b, [&] { b->b.missing_observation(t, label_or_throw(outeq, "output")); });
}
void pharmsol_subject_builder_observation_with_error(pcu_subject_builder* b, double t, double v, const char* outeq, double c0, double c1,
                                                     double c2, double c3, int32_t cens) {
    builder_op(b, [&] {
        if (cens < 0 || cens > 2) throw PharmsolError(PCU_ERR_INVALID_ARGUMENT, "censoring must be 0 (none), 1 (BLOQ) or 2 (ALOQ)");
        b->b.observation_with_error(t, v, label_or_throw(outeq, "output"), ErrorPoly{c0, c1, c2, c3}, (Censor)cens);
    });
}
void pharmsol_subject_builder_covariate(pcu_subject_builder* b, const char* name, double t, double v) {
    builder_op(b, [&] { b->b.covariate(label_or_throw(name, "covariate"), t, v); });
}
void pharmsol_subject_builder_repeat(pcu_subject_builder* b, int64_t n, double delta) {
    builder_op(b, [&] {
        if (n < 0 || n > 10000000) throw PharmsolError(PCU_ERR_INVALID_ARGUMENT, "repeat count " + std::to_string(n) + " is outside [0, 10000000]");
        b->b.repeat((size_t)n, delta);
    });
}
void pharmsol_subject_builder_reset(pcu_subject_builder* b) { builder_op(b, [&] { b->b.reset(); }); }
pcu_subject* pharmsol_subject_builder_build(pcu_subject_builder* b) {
    if (!b) { set_last_error("NULL subject builder"); return nullptr; }
    std::unique_ptr<pcu_subject_builder> owned(b);      // consumed either way
    if (!b->error.empty()) { set_last_error("subject `" + b->b.id + "`: " + b->error); return nullptr; }
    try { return new pcu_subject{b->b.build()}; }
    catch (const std::exception& e) { set_last_error(e.what()); return nullptr; }
}
int32_t pharmsol_subject_set_covariate_fixed(pcu_subject* s, int32_t occasion, const char* name, int32_t fixed) {
    if (!s || !name || occasion < 0 || (size_t)occasion >= s->s.occasions.size()) return PCU_ERR_INVALID_ARGUMENT;
    auto& covs = s->s.occasions[(size_t)occasion].covariates;
    auto it = covs.find(name);
    if (it == covs.end()) return PCU_ERR_MISSING_COVARIATE;
    it->second.fixed = fixed != 0;
    return PCU_OK;
}
void pharmsol_subject_free(pcu_subject* s) { delete s; }
pcu_data* pharmsol_data_new(void) {
    try { return new pcu_data(); } catch (const std::exception& e) { set_last_error(e.what()); return nullptr; }
}
int32_t pharmsol_data_add_subject(pcu_data* d, const pcu_subject* s) {
    return guarded([&] {
        if (!d || !s) return (int32_t)PCU_ERR_INVALID_ARGUMENT;
        d->d.subjects.push_back(s->s);
        return (int32_t)PCU_OK;
    });
}
int64_t pharmsol_data_len(const pcu_data* d) { return d ? (int64_t)d->d.subjects.size() : 0; }
void pharmsol_data_free(pcu_data* d) { delete d; }
int32_t pharmsol_data_read_pmetrics(const char* path, pcu_data** out) {
    return guarded([&] {
        if (!path || !out) return (int32_t)PCU_ERR_INVALID_ARGUMENT;
        auto* d = new pcu_data();
        try { d->d = read_pmetrics_file(path); } catch (...) { delete d; throw; }
        *out = d;
        return (int32_t)PCU_OK;
    });
}
int32_t pharmsol_data_from_pmetrics_text(const char* text, size_t len, pcu_data** out) {
    return guarded([&] {
        if (!text || !out) return (int32_t)PCU_ERR_INVALID_ARGUMENT;
        auto* d = new pcu_data();
        try { d->d = read_pmetrics_text(std::string(text, len)); } catch (...) { delete d; throw; }
        *out = d;
        return (int32_t)PCU_OK;
    });
}
int32_t pharmsol_data_expand(const pcu_data* d, double idelta, double tad, pcu_data** out) {
    return guarded([&] {
        if (!d || !out) return (int32_t)PCU_ERR_INVALID_ARGUMENT;
        auto* n = new pcu_data();
        try { n->d = expand_data(d->d, idelta, tad); } catch (...) { delete n; throw; }
        *out = n;
        return (int32_t)PCU_OK;
    });
}
int64_t pharmsol_data_describe_json(const pcu_data* d, char* buf, size_t cap) {
    if (!d) return -1;
    std::string s;
    try { s = describe_data_json(d->d); } catch (const std::exception& e) { set_last_error(e.what()); return -1; }
    if (buf && cap > 0) {
        const size_t n = std::min(cap - 1, s.size());
        std::memcpy(buf, s.data(), n);
        buf[n] = '\0';
    }
    return (int64_t)s.size();
}

// ---- models ----------------------------------------------------------------------------------------------
static pcu_model* make_model(const std::string& source) {
    std::unique_ptr<pcu_model> m(new pcu_model());
    m->m.cm = dsl::compile_source(source);
    m->m.dsl_source = source;
    std::memset(&m->m.opts, 0, sizeof m->m.opts);
    m->m.opts.rtol = 1e-4; m->m.opts.atol = 1e-4;            // ode/mod.rs:40-41
    m->m.opts.h0 = 0.0;
    m->m.opts.em_dt = 0.05;
    m->m.opts.seed = 0x5eed5eedULL;
    m->m.opts.solver = psi::SOLVER_DOPRI5;
    m->m.opts.cov_time = psi::COVTIME_INTERVAL_END;
    m->m.opts.max_steps = 200000;
    m->m.opts.balance = 1;
    m->m.opts.nparticles = m->m.cm.particles > 0 ? m->m.cm.particles : 1000;
    m->m.opts.sde_mode = psi::SDE_MEAN_PREDICTION;
    m->m.opts.em_mode = psi::EM_REFERENCE_ADAPTIVE;
    m->m.info_json = m->m.cm.model_info_json();
    std::vector<std::pair<int, std::string>> entries;
    if (m->m.cm.kind == dsl::ModelKind::Ode) for (int s = 0; s < psi::SOLVER_COUNT; ++s) entries.emplace_back(s, entry_name(m->m.cm.id, s));
    else entries.emplace_back(0, entry_name(m->m.cm.id, 0));
    m->m.source_cache = m->m.cm.cuda_source(entries, false);
    return m.release();
}
int32_t pharmsol_cuda_model_from_dsl(pcu_ctx*, const char* source, size_t len, pcu_model** out) {
    return guarded([&] {
        if (!source || !out) return (int32_t)PCU_ERR_INVALID_ARGUMENT;
        *out = make_model(std::string(source, len));
        return (int32_t)PCU_OK;
    });
}
// ---- CUDA-target artifact (.pkm) ------------------------------------------------------------------------------
int32_t pharmsol_cuda_model_export_artifact(pcu_model* m, const char* path, const int32_t* solvers, int32_t nsolvers) {
    return guarded([&] {
        if (!m || !path || nsolvers < 0 || (nsolvers > 0 && !solvers)) return (int32_t)PCU_ERR_INVALID_ARGUMENT;
        std::vector<int> list;
        if (nsolvers == 0) list.push_back(effective_solver(m->m));
        for (int i = 0; i < nsolvers; ++i) {
            if (solvers[i] < 0 || solvers[i] >= psi::SOLVER_COUNT) return (int32_t)PCU_ERR_INVALID_ARGUMENT;
            const int sv = m->m.cm.kind == dsl::ModelKind::Ode ? solvers[i] : 0;
            if (std::find(list.begin(), list.end(), sv) == list.end()) list.push_back(sv);
        }
        write_artifact(m->m, path, list);
        return (int32_t)PCU_OK;
    });
}
int32_t pharmsol_cuda_model_load_artifact(pcu_ctx*, const char* path, pcu_model** out) {
    return guarded([&] {
        if (!path || !out) return (int32_t)PCU_ERR_INVALID_ARGUMENT;
        ArtifactFile a = read_artifact(path);
        std::unique_ptr<pcu_model> m(make_model(a.dsl_source));
        try {
            apply_artifact_settings(a.settings, m->m.opts);
        } catch (const PharmsolError& e) { throw PharmsolError(PCU_ERR_OTHER, std::string("artifact ") + path + ": " + e.what());
        } catch (const std::exception&) { throw PharmsolError(PCU_ERR_OTHER, std::string("artifact ") + path + " has malformed settings"); }
        // Device code compiled against another engine build (other kernel-parameter layout) is not trusted: the
        // model then takes the usual registry / cache / NVRTC route from the DSL source it carries.
        if (a.engine_matches) m->m.artifact_cubins = std::move(a.cubins);
        *out = m.release();
        return (int32_t)PCU_OK;
    });
}
int64_t pharmsol_cuda_artifact_info_json(const char* path, char* buf, size_t cap) {
    int64_t need = -1;
    guarded([&] {
        if (!path) return (int32_t)PCU_ERR_INVALID_ARGUMENT;
        ArtifactFile a = read_artifact(path);
        std::string id;
        try { id = dsl::compile_source(a.dsl_source).id; } catch (...) {}
        const std::string js = artifact_info_json(a, id);
        need = (int64_t)js.size();
        if (buf && cap > 0) {
            const size_t n = std::min(js.size(), cap - 1);
            std::memcpy(buf, js.data(), n);
            buf[n] = '\0';
        }
        return (int32_t)PCU_OK;
    });
    return need;
}
// ---- native (host) artifact with the reference's frozen symbols --------------------------------------------------
const char* pharmsol_cuda_model_host_source(pcu_model* m) {
    if (!m) return "";
    std::lock_guard<std::mutex> lk(m->m.mu);
    try {
        if (m->m.host_source_cache.empty()) m->m.host_source_cache = m->m.cm.host_source();
    } catch (const std::exception& e) { set_last_error(e.what()); return ""; }
    return m->m.host_source_cache.c_str();
}
int32_t pharmsol_cuda_model_export_host_artifact(pcu_model* m, const char* path) {
    return guarded([&] {
        if (!m || !path) return (int32_t)PCU_ERR_INVALID_ARGUMENT;
        const std::string src = m->m.cm.host_source();
        const std::string out = path;
        const std::string tmp_src = out + ".src." + std::to_string((long long)::getpid()) + ".cpp";
        const std::string tmp_out = out + ".tmp." + std::to_string((long long)::getpid());
        const std::string log = out + ".log." + std::to_string((long long)::getpid());
        {
            FILE* f = std::fopen(tmp_src.c_str(), "wb");
            if (!f) throw PharmsolError(PCU_ERR_OTHER, "cannot write " + tmp_src);
            std::fwrite(src.data(), 1, src.size(), f);
            std::fclose(f);
        }
        // the reference shells out to `cargo build` for its cdylib (dsl/aot.rs:146-300); the host twin is C++, so the
        // system C++ compiler plays that part ($PHARMSOL_B200_CXX, else $CXX, else g++)
        const char* cxx = std::getenv("PHARMSOL_B200_CXX");
        if (!cxx || !*cxx) cxx = "g++";
        auto quote = [](const std::string& a) { std::string q = "'"; for (char ch : a) { if (ch == '\'') q += "'\\''"; else q += ch; } return q + "'"; };
        const std::string cmd = quote(cxx) + " -std=c++17 -O2 -fPIC -shared -fvisibility=default -ffp-contract=off -o " + quote(tmp_out) + " " + quote(tmp_src) + " > " + quote(log) + " 2>&1";
        const int rc = std::system(cmd.c_str());
        std::string text;
        if (FILE* f = std::fopen(log.c_str(), "rb")) { char buf[4096]; size_t n; while ((n = std::fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, n); std::fclose(f); }
        std::remove(log.c_str());
        std::remove(tmp_src.c_str());
        if (rc != 0) { std::remove(tmp_out.c_str()); throw dsl::DslError("host compiler failed for the native artifact (" + std::string(cxx) + "):\n" + text); }
        if (std::rename(tmp_out.c_str(), out.c_str()) != 0) { std::remove(tmp_out.c_str()); throw PharmsolError(PCU_ERR_OTHER, "cannot write " + out); }
        return (int32_t)PCU_OK;
    });
}
void pharmsol_cuda_model_destroy(pcu_model* m) { delete m; }
int32_t pharmsol_cuda_model_kind(const pcu_model* m) { return m ? (int32_t)m->m.cm.kind : -1; }
int32_t pharmsol_cuda_model_nparams(const pcu_model* m) { return m ? (int32_t)m->m.cm.parameters.size() : -1; }
int32_t pharmsol_cuda_model_nstates(const pcu_model* m) { return m ? m->m.cm.state_len : -1; }
int32_t pharmsol_cuda_model_nouteqs(const pcu_model* m) { return m ? m->m.cm.output_len : -1; }
const char* pharmsol_cuda_model_info_json(const pcu_model* m) { return m ? m->m.info_json.c_str() : ""; }
const char* pharmsol_cuda_model_cuda_source(const pcu_model* m) { return m ? m->m.source_cache.c_str() : ""; }
const char* pharmsol_cuda_model_id(const pcu_model* m) { return m ? m->m.cm.id.c_str() : ""; }
int32_t pharmsol_cuda_model_set_solver(pcu_model* m, int32_t solver, double rtol, double atol) {
    if (!m || solver < 0 || solver >= psi::SOLVER_COUNT || !(rtol > 0) || !(atol > 0)) return PCU_ERR_INVALID_ARGUMENT;
    m->m.opts.solver = solver; m->m.opts.rtol = rtol; m->m.opts.atol = atol;
    return PCU_OK;
}
int32_t pharmsol_cuda_model_set_max_steps(pcu_model* m, int32_t n) {
    if (!m || n <= 0) return PCU_ERR_INVALID_ARGUMENT;
    m->m.opts.max_steps = n;
    return PCU_OK;
}
int32_t pharmsol_cuda_model_set_particles(pcu_model* m, uint32_t n, uint64_t seed, int32_t sde_mode, int32_t em_mode, double em_dt) {
    if (!m || n == 0 || sde_mode < 0 || sde_mode > 1 || em_mode < 0 || em_mode > 1) return PCU_ERR_INVALID_ARGUMENT;
    m->m.opts.nparticles = (int32_t)n; m->m.opts.seed = seed; m->m.opts.sde_mode = sde_mode; m->m.opts.em_mode = em_mode;
    if (em_dt > 0) m->m.opts.em_dt = em_dt;
    return PCU_OK;
}
int32_t pharmsol_cuda_model_set_sde_normals(pcu_model* m, int32_t precision) {
    if (!m || precision < 0 || precision > 1) return PCU_ERR_INVALID_ARGUMENT;
    m->m.opts.sde_normals = precision;
    return PCU_OK;
}
int32_t pharmsol_cuda_model_set_cov_time(pcu_model* m, int32_t mode) {
    if (!m || mode < 0 || mode > 1) return PCU_ERR_INVALID_ARGUMENT;
    m->m.opts.cov_time = mode;
    return PCU_OK;
}
int32_t pharmsol_cuda_model_compile(pcu_ctx* ctx, pcu_model* m, int32_t* source_out) {
    return guarded([&] {
        if (!ctx || !m) return (int32_t)PCU_ERR_INVALID_ARGUMENT;
        cuda_check(cudaSetDevice(ctx->c.device), "cudaSetDevice");
        KernelRef k = get_kernel(m->m, effective_solver(m->m));
        if (source_out) *source_out = k.source;
        return (int32_t)PCU_OK;
    });
}
int32_t pharmsol_cuda_model_precompile_to_cache(pcu_model* m, int32_t solver) {
    return guarded([&] {
        if (!m) return (int32_t)PCU_ERR_INVALID_ARGUMENT;
        const std::string name = entry_name(m->m.cm.id, solver);
        auto cubin = nvrtc_compile_cubin(m->m.cm.cuda_source({{solver, name}}, false), name);
        const std::string path = cubin_cache_path(m->m.cm.id, solver);
        // write-then-rename, like get_kernel: a concurrent reader never sees a half-written cubin
        const std::string tmp = path + ".tmp." + std::to_string((long long)::getpid());
        FILE* f = std::fopen(tmp.c_str(), "wb");
        if (!f) throw PharmsolError(PCU_ERR_OTHER, "cannot write " + tmp);
        const size_t nw = std::fwrite(cubin.data(), 1, cubin.size(), f);
        if (std::fclose(f) != 0 || nw != cubin.size() || std::rename(tmp.c_str(), path.c_str()) != 0) {
            std::remove(tmp.c_str());
            throw PharmsolError(PCU_ERR_OTHER, "cannot write " + path);
        }
        return (int32_t)PCU_OK;
    });
}

// ---- population --------------------------------------------------------------------------------------------
int32_t pharmsol_cuda_population_create(pcu_ctx* ctx, const pcu_model* m, const pcu_data* d, const pcu_error_model* ems,
                                        int32_t n_ems, pcu_population** out) {
    return guarded([&] {
        if (!ctx || !m || !d || !out) return (int32_t)PCU_ERR_INVALID_ARGUMENT;
        cuda_check(cudaSetDevice(ctx->c.device), "cudaSetDevice");
        auto* p = new pcu_population();
        try {
            p->p.data = d->d;
            p->p.labels = m->m.cm.labels();
            p->p.device = ctx->c.device;
            if (ems && n_ems > 0) { p->em = to_models(ems, n_ems); p->has_em = true; }
            p->p.flat = flatten_population(p->p.data, p->p.labels, p->has_em ? &p->em : nullptr);
            for (int k = 1; k < ctx->ndev(); ++k) {       // one replica per further device of a multi-device context
                p->p.replicas.emplace_back(new PopReplica());
                p->p.replicas.back()->device = ctx->dev(k).device;
            }
            p->p.upload();
        } catch (...) { delete p; throw; }
        *out = p;
        return (int32_t)PCU_OK;
    });
}
int32_t pharmsol_cuda_population_set_error_models(pcu_population* pop, const pcu_error_model* ems, int32_t n) {
    return guarded([&] {
        if (!pop) return (int32_t)PCU_ERR_INVALID_ARGUMENT;
        cuda_check(cudaSetDevice(pop->p.device), "cudaSetDevice");
        pop->has_em = ems && n > 0;
        if (pop->has_em) pop->em = to_models(ems, n);
        pop->p.flat = flatten_population(pop->p.data, pop->p.labels, pop->has_em ? &pop->em : nullptr);
        // The upload overwrites the events in place with a blocking copy on the legacy stream, which does not order
        // against the contexts' non-blocking streams or a caller's stream: drain the device first so a psi kernel still
        // in flight from the asynchronous *_device / *_peers entry points never reads a half-updated timeline.
        cuda_check(cudaDeviceSynchronize(), "synchronize before re-flattening the population");
        for (auto& r : pop->p.replicas) {      // the replicas of a multi-device context live on other devices
            cuda_check(cudaSetDevice(r->device), "cudaSetDevice");
            cuda_check(cudaDeviceSynchronize(), "synchronize before re-flattening the population");
        }
        pop->p.upload();
        return (int32_t)PCU_OK;
    });
}
void pharmsol_cuda_population_destroy(pcu_population* pop) {
    if (!pop) return;
    cudaSetDevice(pop->p.device);
    pop->p.dev.release();
    delete pop;
}
int64_t pharmsol_cuda_population_nsubjects(const pcu_population* pop) { return pop ? pop->p.flat.nsub : 0; }
int64_t pharmsol_cuda_population_nobservations(const pcu_population* pop) { return pop ? pop->p.flat.nobs_total : 0; }
int32_t pharmsol_cuda_population_obs_offsets(const pcu_population* pop, int64_t* out) {
    if (!pop || !out) return PCU_ERR_INVALID_ARGUMENT;
    for (size_t i = 0; i < pop->p.flat.obs_offsets.size(); ++i) out[i] = pop->p.flat.obs_offsets[i];
    return PCU_OK;
}
int32_t pharmsol_cuda_population_observation_table(const pcu_population* pop, double* time, double* value, int32_t* outeq, int32_t* occasion,
                                                   int32_t* censoring) {
    if (!pop) return PCU_ERR_INVALID_ARGUMENT;
    const auto& f = pop->p.flat;
    for (size_t oc = 0; oc + 1 < f.ev_offsets.size(); ++oc) {
        for (int32_t e = f.ev_offsets[oc]; e < f.ev_offsets[oc + 1]; ++e) {
            const psi::EventRec& r = f.events[(size_t)e];
            if (psi::ev_kind(r.meta) != psi::EV_OBS || r.obs_row < 0) continue;
            const size_t row = (size_t)r.obs_row;
            if (time) time[row] = r.time;
            if (value) value[row] = psi::ev_has_value(r.meta) ? r.a : std::numeric_limits<double>::quiet_NaN();
            if (outeq) outeq[row] = psi::ev_index(r.meta);
            if (occasion) occasion[row] = f.occ_index[oc];
            if (censoring) censoring[row] = psi::ev_cens(r.meta);
        }
    }
    return PCU_OK;
}
int64_t pharmsol_cuda_population_device_bytes(const pcu_population* pop) { return pop ? (int64_t)pop->p.dev.cap : 0; }

// ---- hot path ------------------------------------------------------------------------------------------------
int32_t pharmsol_cuda_upload_support_points(pcu_ctx* ctx, const double* spp, int64_t nspp, int32_t np, double* spp_soa_dev,
                                            int64_t ld, void* stream) {
    return guarded([&] {
        if (!ctx || !spp || !spp_soa_dev || nspp < 0 || np <= 0 || ld < nspp) return (int32_t)PCU_ERR_INVALID_ARGUMENT;
        std::lock_guard<std::mutex> lk(ctx->c.mu);
        cuda_check(cudaSetDevice(ctx->c.device), "cudaSetDevice");
        cudaStream_t s = pick_stream(ctx->c, stream);
        ctx->c.spp_rows.reserve((size_t)nspp * np * 8);
        cuda_check(cudaMemcpyAsync(ctx->c.spp_rows.p, spp, (size_t)nspp * np * 8, cudaMemcpyHostToDevice, s), "H2D support points");
        launch_transpose(ctx->c.spp_rows.as<double>(), spp_soa_dev, nspp, np, ld, s);
        ctx->c.launches += 1;
        return (int32_t)PCU_OK;
    });
}

int32_t pharmsol_cuda_log_likelihood_matrix_device(pcu_ctx* ctx, pcu_model* m, pcu_population* pop, const double* spp_soa_dev,
                                                   int64_t ncols, int64_t ld_spp, double* out_dev, int64_t ld_out,
                                                   int64_t first_col, void* stream) {
    return guarded([&] {
        if (!ctx || !m || !pop || !spp_soa_dev || !out_dev || ld_out < pop->p.flat.nsub) return (int32_t)PCU_ERR_INVALID_ARGUMENT;
        std::lock_guard<std::mutex> lk(ctx->c.mu);
        cuda_check(cudaSetDevice(ctx->c.device), "cudaSetDevice");
        launch_psi(ctx->c, m->m, pop->p, spp_soa_dev, ncols, ld_spp, out_dev, ld_out, nullptr, 0, first_col, pick_stream(ctx->c, stream));
        return (int32_t)PCU_OK;
    });
}
int32_t pharmsol_cuda_log_likelihood_matrix_peers(pcu_ctx* ctx, pcu_model* m, pcu_population* pop, const double* spp_soa_dev, int64_t ncols,
                                                  int64_t ld_spp, double* const* out_full_peers, int32_t npeers, int64_t ld_out,
                                                  int64_t first_col, void* stream) {
    return guarded([&] {
        if (!ctx || !m || !pop || !spp_soa_dev || !out_full_peers || npeers < 1 || npeers > 8 || ld_out < pop->p.flat.nsub) return (int32_t)PCU_ERR_INVALID_ARGUMENT;
        for (int r = 0; r < npeers; ++r) if (!out_full_peers[r]) return (int32_t)PCU_ERR_INVALID_ARGUMENT;
        std::lock_guard<std::mutex> lk(ctx->c.mu);
        cuda_check(cudaSetDevice(ctx->c.device), "cudaSetDevice");
        launch_psi(ctx->c, m->m, pop->p, spp_soa_dev, ncols, ld_spp, nullptr, ld_out, nullptr, 0, first_col, pick_stream(ctx->c, stream), nullptr, true,
                   out_full_peers, npeers);
        return (int32_t)PCU_OK;
    });
}
// Column-sharded psi with a copy-engine push gather (one process per GPU: the peers' matrices are mapped into this
// process by CUDA IPC / symmetric memory; one process for all GPUs: cudaDeviceEnablePeerAccess).
int32_t pharmsol_cuda_log_likelihood_matrix_push(pcu_ctx* ctx, pcu_model* m, pcu_population* pop, const double* spp_soa_dev, int64_t ncols,
                                                 int64_t ld_spp, double* const* out_full_peers, int32_t npeers, int32_t self, int64_t ld_out,
                                                 int64_t first_col, void* stream) {
    return guarded([&] {
        if (!ctx || !m || !pop || !spp_soa_dev || !out_full_peers || npeers < 1 || npeers > 64 || self < 0 || self >= npeers || ld_out < pop->p.flat.nsub)
            return (int32_t)PCU_ERR_INVALID_ARGUMENT;
        for (int r = 0; r < npeers; ++r) if (!out_full_peers[r]) return (int32_t)PCU_ERR_INVALID_ARGUMENT;
        std::lock_guard<std::mutex> lk(ctx->c.mu);
        cuda_check(cudaSetDevice(ctx->c.device), "cudaSetDevice");
        launch_psi_push(ctx->c, m->m, pop->p, pop->p.view, spp_soa_dev, ncols, ld_spp, out_full_peers, npeers, self, ld_out, first_col, pick_stream(ctx->c, stream));
        return (int32_t)PCU_OK;
    });
}
int32_t pharmsol_cuda_predictions_device(pcu_ctx* ctx, pcu_model* m, pcu_population* pop, const double* spp_soa_dev, int64_t ncols,
                                         int64_t ld_spp, double* pred_dev, int64_t ld_pred, double* ll_dev, int64_t ld_out, void* stream) {
    return guarded([&] {
        if (!ctx || !m || !pop || !spp_soa_dev || !pred_dev) return (int32_t)PCU_ERR_INVALID_ARGUMENT;
        std::lock_guard<std::mutex> lk(ctx->c.mu);
        cuda_check(cudaSetDevice(ctx->c.device), "cudaSetDevice");
        launch_psi(ctx->c, m->m, pop->p, spp_soa_dev, ncols, ld_spp, ll_dev, ld_out, pred_dev, ld_pred, 0, pick_stream(ctx->c, stream));
        return (int32_t)PCU_OK;
    });
}
int32_t pharmsol_cuda_status_batch_begin(pcu_ctx* ctx, void* stream) {
    return guarded([&] {
        if (!ctx) return (int32_t)PCU_ERR_INVALID_ARGUMENT;
        std::lock_guard<std::mutex> lk(ctx->c.mu);
        Ctx& c = ctx->c;
        cuda_check(cudaSetDevice(c.device), "cudaSetDevice");
        cudaStream_t s = pick_stream(c, stream);
        c.err_ctr.reserve(5 * sizeof(unsigned long long));
        if (c.status_clean_on && c.status_clean_on != s) cuda_check(cudaStreamWaitEvent(s, c.reset_ev, 0), "wait for queued status reset");
        cuda_check(cudaMemcpyAsync(c.err_ctr.p, c.err_host + 8, 5 * sizeof(unsigned long long), cudaMemcpyHostToDevice, s), "reset status");
        cuda_check(cudaEventRecord(c.ev0, s), "event record");
        c.status_clean_on = nullptr;
        c.status_batch = true;
        return (int32_t)PCU_OK;
    });
}
int32_t pharmsol_cuda_collect_errors(pcu_ctx* ctx, int32_t* code, int64_t* pair) {
    return guarded([&] {
        if (!ctx) return (int32_t)PCU_ERR_INVALID_ARGUMENT;
        std::lock_guard<std::mutex> lk(ctx->c.mu);
        cuda_check(cudaSetDevice(ctx->c.device), "cudaSetDevice");
        cuda_check(cudaDeviceSynchronize(), "synchronize");
        ctx->c.status_batch = false;
        return collect(ctx->c, code, pair);
    });
}

// ---- host-buffer calls ---------------------------------------------------------------------------------------
// One device's share of a host-buffer psi call: H2D of its support-point rows, on-device transpose, the column chunks
// pipelined against their copy-back (psi is column-major, so a block of columns is one contiguous slab of `out`; a
// chunk is copied device -> host on the copy stream while the next one is computed and only the LAST copy is exposed).
//  * closed-form models (kernel short next to the copy): up to 16 equal chunks of >= 2 MB and >= 128 columns — for a
//    1000 x 1000 one-compartment matrix the 8 MB copy-back is otherwise longer than the kernel;
//  * adaptive ODE / SDE models (kernel long next to the copy): two chunks, 7/8 + 1/8 of the columns
//    (profiles/r01_tuning.md).
// `first_col` = global index of the shard's first column (error pairs and SDE random streams are global).
static int32_t matrix_host_shard(Ctx& c, Model& m, Population& pop, const psi::PopView& view, const double* spp_rows, int64_t ncols, int32_t np,
                                 double* out, int64_t first_col, bool exponentiate, int32_t* code, int64_t* pair) {
    cuda_check(cudaSetDevice(c.device), "cudaSetDevice");
    c.status_batch = false;      // a host-buffer call resets and reads the status itself
    const int64_t nsub = pop.flat.nsub;
    c.spp_soa.reserve((size_t)ncols * np * 8);
    c.out.reserve((size_t)nsub * ncols * 8);
    c.spp_rows.reserve((size_t)ncols * np * 8);
    cuda_check(cudaMemcpyAsync(c.spp_rows.p, spp_rows, (size_t)ncols * np * 8, cudaMemcpyHostToDevice, c.stream), "H2D support points");
    launch_transpose(c.spp_rows.as<double>(), c.spp_soa.as<double>(), ncols, np, ncols, c.stream);
    c.launches += 1;
    const std::vector<int64_t> cuts = column_chunks(m, nsub, ncols, Ctx::kMaxChunks);
    // All the launches are queued first and the copies second: a copy into PAGEABLE caller memory (a Rust Array2, a
    // numpy array) blocks the issuing thread until it is done, and issued chunk by chunk in one loop it would serialise
    // with the launches (C3 shard: 63 ms = 32 ms kernels + 31 ms copies; queued this way the copies of the early chunks
    // run under the kernels of the later ones).  For pinned buffers the order makes no difference.
    for (size_t k = 0; k + 1 < cuts.size(); ++k) {
        const int64_t c0 = cuts[k], c1 = cuts[k + 1];
        double* slab = c.out.as<double>() + c0 * nsub;
        launch_psi(c, m, pop, c.spp_soa.as<double>() + c0, c1 - c0, ncols, slab, nsub, nullptr, 0, first_col + c0, c.stream, nullptr, k == 0, nullptr, 0, &view);
        if (exponentiate) { launch_exp_inplace(slab, nsub * (c1 - c0), c.stream); c.launches += 1; }
        cuda_check(cudaEventRecord(c.chunk_ev[k], c.stream), "chunk event");
    }
    cuda_check(cudaMemcpyAsync(c.err_host, c.err_ctr.p, 5 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c.stream), "D2H status");
    for (size_t k = 0; k + 1 < cuts.size(); ++k) {
        const int64_t c0 = cuts[k], c1 = cuts[k + 1];
        cuda_check(cudaStreamWaitEvent(c.copy_stream, c.chunk_ev[k], 0), "chunk wait");
        cuda_check(cudaMemcpyAsync(out + c0 * nsub, c.out.as<double>() + c0 * nsub, (size_t)(nsub * (c1 - c0)) * 8, cudaMemcpyDeviceToHost, c.copy_stream), "D2H psi");
    }
    cuda_check(cudaMemcpyAsync(c.err_host, c.err_ctr.p, 5 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c.stream), "D2H status");
    cuda_check(cudaStreamSynchronize(c.stream), "synchronize");
    cuda_check(cudaStreamSynchronize(c.copy_stream), "synchronize");
    return collect(c, code, pair, true);
}

// Contiguous column blocks of a multi-device call: block k = [k*per, min((k+1)*per, nspp)).
static int64_t shard_columns(int64_t nspp, int ndev) { return (nspp + ndev - 1) / ndev; }
static bool use_all_devices(const pcu_ctx* ctx, int64_t nspp) { return ctx->ndev() > 1 && nspp >= 32ll * ctx->ndev(); }

struct ShardResult {
    int32_t rc = PCU_OK, code = 0;
    int64_t pair = -1;
    std::string msg;
};
// Run fn(k) for every device of the context concurrently (device 0 on the calling thread, the others on their own host
// threads: copies from pageable caller memory block the issuing thread, so one thread per device keeps the devices
// busy at the same time) and fold the results: the FIRST failing pair over all shards wins (matrix.rs:96-104);
// library-level failures (CUDA, compile) win over pair errors.
extern "C++" {
template <class F>
static int32_t for_each_device(pcu_ctx* ctx, F&& fn, int32_t* code, int64_t* pair) {
    const int n = ctx->ndev();
    std::vector<ShardResult> res((size_t)n);
    auto run = [&](int k) {
        ShardResult& r = res[(size_t)k];
        r.rc = guarded([&] { return fn(k, &r.code, &r.pair); });
        if (r.rc != PCU_OK) r.msg = last_error();
    };
    std::vector<std::thread> workers;
    workers.reserve((size_t)n);
    int inline_from = n;                                  // devices whose host thread could not be started run on this one
    for (int k = 1; k < n; ++k) {
        try { workers.emplace_back(run, k); }
        catch (...) { inline_from = k; break; }
    }
    run(0);
    for (int k = inline_from; k < n; ++k) run(k);
    for (auto& w : workers) w.join();
    cudaSetDevice(ctx->c.device);
    int32_t out_code = 0;
    int64_t out_pair = -1;
    const ShardResult* hard = nullptr;
    const ShardResult* first = nullptr;
    for (const auto& r : res) {
        if (r.rc != PCU_OK && r.code == 0 && !hard) hard = &r;          // not a pair error
        if (r.code != 0 && (!first || r.pair < first->pair)) first = &r;
    }
    if (hard) { set_last_error(hard->msg); if (code) *code = 0; if (pair) *pair = -1; return hard->rc; }
    if (first) { out_code = first->code; out_pair = first->pair; set_last_error(first->msg); }
    if (code) *code = out_code;
    if (pair) *pair = out_pair;
    return first ? first->rc : (int32_t)PCU_OK;
}
}  // extern "C++"

static int32_t matrix_host(pcu_ctx* ctx, pcu_model* m, pcu_population* pop, const double* spp, int64_t nspp, int32_t np,
                           double* out, int32_t* code, int64_t* pair, bool exponentiate) {
    return guarded([&] {
        if (!ctx || !m || !pop || !spp || !out || nspp < 0) return (int32_t)PCU_ERR_INVALID_ARGUMENT;
        if (np != (int32_t)m->m.cm.parameters.size()) {
            // dsl/native.rs:688-698 validate_support_point
            throw PharmsolError(PCU_ERR_OTHER, "model `" + m->m.cm.name + "` expects " + std::to_string(m->m.cm.parameters.size()) +
                                                   " parameter value(s), got " + std::to_string(np));
        }
        const int64_t nsub = pop->p.flat.nsub;
        if (nspp == 0 || nsub == 0) { if (code) *code = 0; if (pair) *pair = -1; return (int32_t)PCU_OK; }
        if (use_all_devices(ctx, nspp)) {
            std::lock_guard<std::mutex> lk(ctx->c.mu);
            cuda_check(cudaSetDevice(ctx->c.device), "cudaSetDevice");
            if ((int)pop->p.replicas.size() + 1 != ctx->ndev()) throw PharmsolError(PCU_ERR_OTHER, "population was created for another context (device list differs)");
            const int64_t per = shard_columns(nspp, ctx->ndev());
            return for_each_device(ctx, [&](int k, int32_t* scode, int64_t* spair) {
                const int64_t lo = std::min<int64_t>(nspp, k * per), hi = std::min<int64_t>(nspp, lo + per);
                if (hi <= lo) return (int32_t)PCU_OK;
                return matrix_host_shard(ctx->dev(k), m->m, pop->p, pop->p.view_on(k), spp + lo * np, hi - lo, np, out + lo * nsub, lo, exponentiate, scode, spair);
            }, code, pair);
        }
        LaneLock lane = acquire_lane(ctx);      // one of up to 1 + kMaxLanes bundles on the device: concurrent callers overlap
        Ctx& c = *lane.c;
        struct Publish { pcu_ctx* ctx; Ctx& c; ~Publish() { publish_lane_stats(ctx, c); } } publish{ctx, c};
        cuda_check(cudaSetDevice(c.device), "cudaSetDevice");
        c.status_batch = false;      // a host-buffer call resets and reads the status itself
        if (c.small_host && (size_t)nspp * np * 8 <= Ctx::kSmallIn && (size_t)nsub * nspp * 8 <= Ctx::kSmallOut) {
            // Latency-bound call (an optimiser's cost function: one or a few support points).  Support points are
            // transposed on the host into pinned memory; results and status come back through pinned memory with one
            // synchronize; the status reset for the NEXT call is queued after this one has been read back, off the
            // critical path.  For the very smallest calls (<= 2 KB each way) the kernel reads the support points and
            // stores psi directly in the pinned host buffer (unified addressing), which removes both DMA copies;
            // measured on the 32 x 64 criterion shapes, beyond a few KB the copy engines are faster than PCIe stores
            // from the SMs (profiles/r01_tuning.md).
            c.spp_soa.reserve((size_t)nspp * np * 8);
            c.out.reserve((size_t)nsub * nspp * 8);
            double* in = c.small_host;
            double* res = c.small_host + Ctx::kSmallIn / 8;
            for (int64_t j = 0; j < nspp; ++j)
                for (int32_t k = 0; k < np; ++k) in[(size_t)k * nspp + j] = spp[(size_t)j * np + k];
            constexpr size_t kInPlace = 2048;
            const bool in_place_in = (size_t)nspp * np * 8 <= kInPlace;
            const bool in_place_out = !exponentiate && (size_t)nsub * nspp * 8 <= kInPlace;
            if (!in_place_in) cuda_check(cudaMemcpyAsync(c.spp_soa.p, in, (size_t)nspp * np * 8, cudaMemcpyHostToDevice, c.stream), "H2D support points");
            launch_psi(c, m->m, pop->p, in_place_in ? in : c.spp_soa.as<double>(), nspp, nspp, in_place_out ? res : c.out.as<double>(), nsub, nullptr, 0, 0,
                       c.stream, nullptr, true);
            if (exponentiate) { launch_exp_inplace(c.out.as<double>(), nsub * nspp, c.stream); c.launches += 1; }
            if (!in_place_out) cuda_check(cudaMemcpyAsync(res, c.out.p, (size_t)nsub * nspp * 8, cudaMemcpyDeviceToHost, c.stream), "D2H psi");
            cuda_check(cudaMemcpyAsync(c.err_host, c.err_ctr.p, 5 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c.stream), "D2H status");
            cuda_check(cudaStreamSynchronize(c.stream), "synchronize");
            std::memcpy(out, res, (size_t)nsub * nspp * 8);
            const int32_t rc = collect(c, code, pair, true);
            cuda_check(cudaMemcpyAsync(c.err_ctr.p, c.err_host + 8, 5 * sizeof(unsigned long long), cudaMemcpyHostToDevice, c.stream), "reset status");
            cuda_check(cudaEventRecord(c.reset_ev, c.stream), "event record");
            c.status_clean_on = c.stream;
            return rc;
        }
        return matrix_host_shard(c, m->m, pop->p, pop->p.view, spp, nspp, np, out, 0, exponentiate, code, pair);
    });
}

int32_t pharmsol_cuda_log_likelihood_matrix(pcu_ctx* ctx, pcu_model* m, pcu_population* pop, const double* spp, int64_t nspp,
                                            int32_t np, double* out, int32_t* code, int64_t* pair) {
    return matrix_host(ctx, m, pop, spp, nspp, np, out, code, pair, false);
}
int32_t pharmsol_cuda_psi(pcu_ctx* ctx, pcu_model* m, pcu_population* pop, const double* spp, int64_t nspp, int32_t np, double* out,
                          int32_t* code, int64_t* pair) {
    return matrix_host(ctx, m, pop, spp, nspp, np, out, code, pair, true);
}

// psi resident and REPLICATED on every device of the context (north_star: "psi shards gathered over NVLink"): device k
// evaluates its block of columns chunk by chunk into its own full matrix and the copy engines push every finished
// chunk to the other devices while the next chunk is computed (`gather` = PCU_GATHER_COPY_ENGINE), or the psi kernel
// stores every result straight into all the matrices (PCU_GATHER_PEER_STORES).  dev_out[k] = the library-owned
// (nsub x nspp) column-major matrix on device k, valid until the next replicated call on this context.
int32_t pharmsol_cuda_log_likelihood_matrix_replicated(pcu_ctx* ctx, pcu_model* m, pcu_population* pop, const double* spp, int64_t nspp,
                                                       int32_t np, int32_t gather, double** dev_out, int32_t* code, int64_t* pair) {
    return guarded([&] {
        if (!ctx || !m || !pop || !spp || !dev_out || nspp < 0 || gather < 0 || gather > 1) return (int32_t)PCU_ERR_INVALID_ARGUMENT;
        if (np != (int32_t)m->m.cm.parameters.size())
            throw PharmsolError(PCU_ERR_OTHER, "model `" + m->m.cm.name + "` expects " + std::to_string(m->m.cm.parameters.size()) +
                                                   " parameter value(s), got " + std::to_string(np));
        std::lock_guard<std::mutex> lk(ctx->c.mu);
        const int n = ctx->ndev();
        if (n > 8) throw PharmsolError(PCU_ERR_OTHER, "the replicated psi supports at most 8 devices");
        if (gather == PCU_GATHER_PEER_STORES && !ctx->peer_enabled && n > 1) {
            bool distinct = false;
            for (int k = 1; k < n; ++k) distinct = distinct || ctx->dev(k).device != ctx->c.device;
            if (distinct) throw PharmsolError(PCU_ERR_OTHER, "PCU_GATHER_PEER_STORES needs peer access between all devices of the context; use PCU_GATHER_COPY_ENGINE");
        }
        if ((int)pop->p.replicas.size() + 1 != n) throw PharmsolError(PCU_ERR_OTHER, "population was created for another context (device list differs)");
        const int64_t nsub = pop->p.flat.nsub;
        for (int k = 0; k < n; ++k) dev_out[k] = nullptr;
        if (nspp == 0 || nsub == 0) { if (code) *code = 0; if (pair) *pair = -1; return (int32_t)PCU_OK; }
        std::vector<double*> peers((size_t)n);
        for (int k = 0; k < n; ++k) {
            Ctx& c = ctx->dev(k);
            cuda_check(cudaSetDevice(c.device), "cudaSetDevice");
            c.full.reserve((size_t)nsub * nspp * 8);
            peers[(size_t)k] = c.full.as<double>();
        }
        const int64_t per = shard_columns(nspp, n);
        // phase 1 (concurrent): upload + launches + pushes, all asynchronous on each device's stream
        int32_t rc = for_each_device(ctx, [&](int k, int32_t*, int64_t*) {
            Ctx& c = ctx->dev(k);
            cuda_check(cudaSetDevice(c.device), "cudaSetDevice");
            c.status_batch = false;
            const int64_t lo = std::min<int64_t>(nspp, k * per), hi = std::min<int64_t>(nspp, lo + per);
            if (hi <= lo) return (int32_t)PCU_OK;
            const int64_t ncols = hi - lo;
            c.spp_rows.reserve((size_t)ncols * np * 8);
            c.spp_soa.reserve((size_t)ncols * np * 8);
            cuda_check(cudaMemcpyAsync(c.spp_rows.p, spp + lo * np, (size_t)ncols * np * 8, cudaMemcpyHostToDevice, c.stream), "H2D support points");
            launch_transpose(c.spp_rows.as<double>(), c.spp_soa.as<double>(), ncols, np, ncols, c.stream);
            c.launches += 1;
            if (gather == PCU_GATHER_PEER_STORES) {
                launch_psi(c, m->m, pop->p, c.spp_soa.as<double>(), ncols, ncols, nullptr, nsub, nullptr, 0, lo, c.stream, nullptr, true, peers.data(), n,
                           &pop->p.view_on(k));
            } else {
                launch_psi_push(c, m->m, pop->p, pop->p.view_on(k), c.spp_soa.as<double>(), ncols, ncols, peers.data(), n, k, nsub, lo, c.stream);
            }
            cuda_check(cudaMemcpyAsync(c.err_host, c.err_ctr.p, 5 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c.stream), "D2H status");
            return (int32_t)PCU_OK;
        }, nullptr, nullptr);
        // phase 2: every device's stream drained == every matrix complete (each stream ends after its own pushes / stores)
        int32_t first_code = 0;
        int64_t first_pair = -1;
        std::string first_msg;
        for (int k = 0; k < n; ++k) {
            Ctx& c = ctx->dev(k);
            cuda_check(cudaSetDevice(c.device), "cudaSetDevice");
            cuda_check(cudaStreamSynchronize(c.stream), "synchronize");
            const int64_t lo = std::min<int64_t>(nspp, k * per);
            if (rc != PCU_OK || lo >= nspp) continue;
            int32_t sc = 0; int64_t sp = -1;
            if (collect(c, &sc, &sp, true) != PCU_OK && sc != 0 && (first_code == 0 || sp < first_pair)) { first_code = sc; first_pair = sp; first_msg = last_error(); }
        }
        cuda_check(cudaSetDevice(ctx->c.device), "cudaSetDevice");
        if (rc != PCU_OK) return rc;
        for (int k = 0; k < n; ++k) dev_out[k] = peers[(size_t)k];
        if (code) *code = first_code;
        if (pair) *pair = first_pair;
        if (first_code) set_last_error(first_msg);
        return first_code;
    });
}

// estimate_predictions for one device's block of columns.  The (nobs x ncols) block is produced in column chunks of at
// most ~256 MB through two device buffers: chunk k is copied back (a strided 2-D copy into the caller's row-major
// matrix) while chunk k+1 is computed, so a C3-sized grid (10^5 rows x 6,250 columns = 5 GB) needs 0.5 GB of HBM.
static int32_t predictions_shard(Ctx& c, Model& m, Population& pop, const psi::PopView& view, const double* spp_rows, int64_t ncols, int32_t np,
                                 double* out, int64_t ld_out, int64_t first_col, int32_t* code, int64_t* pair) {
    cuda_check(cudaSetDevice(c.device), "cudaSetDevice");
    c.status_batch = false;
    const int64_t nobs = pop.flat.nobs_total;
    int64_t budget = 256ll << 20;
    if (const char* kb = std::getenv("PHARMSOL_B200_PRED_CHUNK_KB")) budget = std::max<int64_t>(1, std::atoll(kb)) << 10;      // test knob
    int64_t chunk = std::max<int64_t>(128, (budget / std::max<int64_t>(1, nobs * 8)) / 128 * 128);
    chunk = std::min<int64_t>(chunk, (ncols + 127) / 128 * 128);
    c.spp_rows.reserve((size_t)ncols * np * 8);
    c.spp_soa.reserve((size_t)ncols * np * 8);
    c.pred.reserve((size_t)nobs * chunk * 8 * 2);
    cuda_check(cudaMemcpyAsync(c.spp_rows.p, spp_rows, (size_t)ncols * np * 8, cudaMemcpyHostToDevice, c.stream), "H2D support points");
    launch_transpose(c.spp_rows.as<double>(), c.spp_soa.as<double>(), ncols, np, ncols, c.stream);
    c.launches += 1;
    int k = 0;
    for (int64_t c0 = 0; c0 < ncols; c0 += chunk, ++k) {
        const int64_t c1 = std::min<int64_t>(ncols, c0 + chunk);
        double* buf = c.pred.as<double>() + (size_t)(k & 1) * nobs * chunk;
        // the buffer is free again once the copy issued two chunks ago has finished
        if (k >= 2) cuda_check(cudaStreamWaitEvent(c.stream, c.chunk_ev[Ctx::kMaxChunks - 1 - (k & 1)], 0), "buffer wait");
        // predictions need no error model: run with the likelihood output disabled
        launch_psi(c, m, pop, c.spp_soa.as<double>() + c0, c1 - c0, ncols, nullptr, pop.flat.nsub, buf, chunk, first_col + c0, c.stream, nullptr, k == 0, nullptr, 0, &view);
        c.status_batch = true;
        cuda_check(cudaEventRecord(c.chunk_ev[k & 1], c.stream), "chunk event");
        cuda_check(cudaStreamWaitEvent(c.copy_stream, c.chunk_ev[k & 1], 0), "chunk wait");
        cuda_check(cudaMemcpy2DAsync(out + first_col + c0, (size_t)ld_out * 8, buf, (size_t)chunk * 8, (size_t)(c1 - c0) * 8, (size_t)nobs, cudaMemcpyDeviceToHost,
                                     c.copy_stream), "D2H predictions");
        cuda_check(cudaEventRecord(c.chunk_ev[Ctx::kMaxChunks - 1 - (k & 1)], c.copy_stream), "copy event");
    }
    c.status_batch = false;
    cuda_check(cudaMemcpyAsync(c.err_host, c.err_ctr.p, 5 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c.stream), "D2H status");
    cuda_check(cudaStreamSynchronize(c.stream), "synchronize");
    cuda_check(cudaStreamSynchronize(c.copy_stream), "synchronize");
    return collect(c, code, pair, true);
}

int32_t pharmsol_cuda_predictions(pcu_ctx* ctx, pcu_model* m, pcu_population* pop, const double* spp, int64_t nspp, int32_t np, double* out) {
    return guarded([&] {
        if (!ctx || !m || !pop || !spp || !out || nspp < 0) return (int32_t)PCU_ERR_INVALID_ARGUMENT;
        if (np != (int32_t)m->m.cm.parameters.size())
            throw PharmsolError(PCU_ERR_OTHER, "model `" + m->m.cm.name + "` expects " + std::to_string(m->m.cm.parameters.size()) +
                                                   " parameter value(s), got " + std::to_string(np));
        std::lock_guard<std::mutex> lk(ctx->c.mu);
        const int64_t nobs = pop->p.flat.nobs_total;
        if (nspp == 0 || nobs == 0) return (int32_t)PCU_OK;
        int32_t code = 0; int64_t pair = -1;
        if (use_all_devices(ctx, nspp)) {
            if ((int)pop->p.replicas.size() + 1 != ctx->ndev()) throw PharmsolError(PCU_ERR_OTHER, "population was created for another context (device list differs)");
            const int64_t per = shard_columns(nspp, ctx->ndev());
            return for_each_device(ctx, [&](int k, int32_t* scode, int64_t* spair) {
                const int64_t lo = std::min<int64_t>(nspp, k * per), hi = std::min<int64_t>(nspp, lo + per);
                if (hi <= lo) return (int32_t)PCU_OK;
                return predictions_shard(ctx->dev(k), m->m, pop->p, pop->p.view_on(k), spp + lo * np, hi - lo, np, out, nspp, lo, scode, spair);
            }, &code, &pair);
        }
        return predictions_shard(ctx->c, m->m, pop->p, pop->p.view, spp, nspp, np, out, nspp, 0, &code, &pair);
    });
}

int32_t pharmsol_cuda_log_likelihood_batch(pcu_ctx* ctx, pcu_model* m, pcu_population* pop, const double* params, int64_t nrows, int32_t np,
                                           const pcu_residual_error_model* models, int32_t n_models, double* out) {
    return guarded([&] {
        if (!ctx || !m || !pop || !params || !out || nrows < 0 || n_models < 0 || (n_models > 0 && !models)) return (int32_t)PCU_ERR_INVALID_ARGUMENT;
        const int64_t nsub = pop->p.flat.nsub;
        if (nrows != nsub)      // likelihood/mod.rs:128-134
            throw PharmsolError(PCU_ERR_OTHER, "parameters has " + std::to_string(nrows) + " rows but there are " + std::to_string(nsub) + " subjects");
        if (np != (int32_t)m->m.cm.parameters.size())
            throw PharmsolError(PCU_ERR_OTHER, "model `" + m->m.cm.name + "` expects " + std::to_string(m->m.cm.parameters.size()) +
                                                   " parameter value(s), got " + std::to_string(np));
        if (n_models > psi::PSI_MAX_RESID) throw PharmsolError(PCU_ERR_OTHER, "at most " + std::to_string(psi::PSI_MAX_RESID) + " residual error models");
        if (nsub == 0) return (int32_t)PCU_OK;
        std::lock_guard<std::mutex> lk(ctx->c.mu);
        Ctx& c = ctx->c;
        cuda_check(cudaSetDevice(c.device), "cudaSetDevice");
        c.status_batch = false;
        psi::RunOpts opt = m->m.opts;
        opt.diagonal = 1;
        opt.nresid = n_models;
        for (int k = 0; k < psi::PSI_MAX_RESID; ++k) {
            opt.resid[k].kind = k < n_models ? models[k].kind : psi::RESID_MISSING;
            opt.resid[k].pad = 0;
            opt.resid[k].a = k < n_models ? models[k].a : 0.0;
            opt.resid[k].b = k < n_models ? models[k].b : 0.0;
        }
        c.spp_rows.reserve((size_t)nsub * np * 8);
        c.spp_soa.reserve((size_t)nsub * np * 8);
        c.out.reserve((size_t)nsub * 8);
        cuda_check(cudaMemcpyAsync(c.spp_rows.p, params, (size_t)nsub * np * 8, cudaMemcpyHostToDevice, c.stream), "H2D parameters");
        launch_transpose(c.spp_rows.as<double>(), c.spp_soa.as<double>(), nsub, np, nsub, c.stream);
        c.launches += 1;
        launch_psi(c, m->m, pop->p, c.spp_soa.as<double>(), nsub, nsub, c.out.as<double>(), nsub, nullptr, 0, 0, c.stream, &opt);
        cuda_check(cudaMemcpyAsync(out, c.out.p, (size_t)nsub * 8, cudaMemcpyDeviceToHost, c.stream), "D2H log-likelihoods");
        cuda_check(cudaStreamSynchronize(c.stream), "synchronize");
        int32_t code = 0; int64_t pair = -1;
        collect(c, &code, &pair);      // failures are -inf entries, never an error (mod.rs:134-137)
        return (int32_t)PCU_OK;
    });
}

int32_t pharmsol_cuda_measure_fp64_peak(pcu_ctx* ctx, double* tflops, double* clock_mhz) {
    return guarded([&] {
        if (!ctx || !tflops) return (int32_t)PCU_ERR_INVALID_ARGUMENT;
        std::lock_guard<std::mutex> lk(ctx->c.mu);
        cuda_check(cudaSetDevice(ctx->c.device), "cudaSetDevice");
        *tflops = measure_fp64_peak(ctx->c, clock_mhz);
        return (int32_t)PCU_OK;
    });
}

}  // extern "C"
