// hostsim_api.cpp — TEST INFRASTRUCTURE ONLY (see cuda_shim.h).  Drives an emitted model translation unit that was
// compiled for the host: flattens the population with the product's own flattener (csrc/host/data.cpp), transposes the
// support points to SoA, walks the launch grid one "thread" at a time and calls the kernel entry as a plain function.
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

#include "dsl.hpp"
#include "launch_geometry.hpp"

struct hs_dim3 { unsigned x = 1, y = 1, z = 1; };
hs_dim3 threadIdx, blockIdx, blockDim, gridDim;

using namespace pharmsol;
typedef void (*entry_fn)(psi::PopView, const double*, long long, long long, psi::RunOpts, psi::OutView);

struct HostSim {
    dsl::CompiledModel cm;
    Data data;
    std::string error;
};

static Data parse_ops(const std::string& text) {
    Data d;
    std::istringstream is(text);
    std::string line;
    std::unique_ptr<SubjectBuilder> b;
    std::vector<std::tuple<int, std::string, int>> fixed;
    auto finish = [&] {
        if (!b) return;
        Subject s = b->build();
        for (auto& f : fixed) {
            auto& covs = s.occasions.at((size_t)std::get<0>(f)).covariates;
            auto it = covs.find(std::get<1>(f));
            if (it != covs.end()) it->second.fixed = std::get<2>(f) != 0;
        }
        d.subjects.push_back(std::move(s));
        b.reset();
        fixed.clear();
    };
    while (std::getline(is, line)) {
        if (line.empty()) continue;
        std::istringstream ls(line);
        std::string k;
        ls >> k;
        if (k == "S") { finish(); std::string id; ls >> id; b.reset(new SubjectBuilder(id)); continue; }
        if (!b) throw PharmsolError(psi::ST_OTHER, "op before subject");
        double t = 0, a = 0, c = 0; std::string lab;
        if (k == "B") { ls >> t >> a >> lab; b->bolus(t, a, lab); }
        else if (k == "I") { ls >> t >> a >> lab >> c; b->infusion(t, a, lab, c); }
        else if (k == "O") { ls >> t >> a >> lab; b->observation(t, a, lab); }
        else if (k == "M") { ls >> t >> lab; b->missing_observation(t, lab); }
        else if (k == "C") { int ce; ls >> t >> a >> lab >> ce; b->censored_observation(t, a, lab, (Censor)ce); }
        else if (k == "E") { ErrorPoly p; int ce; ls >> t >> a >> lab >> p.c0 >> p.c1 >> p.c2 >> p.c3 >> ce; b->observation_with_error(t, a, lab, p, (Censor)ce); }
        else if (k == "V") { ls >> lab >> t >> a; b->covariate(lab, t, a); }
        else if (k == "R") { long n; ls >> n >> a; b->repeat((size_t)n, a); }
        else if (k == "X") { b->reset(); }
        else if (k == "F") { int occ, fx; ls >> occ >> lab >> fx; fixed.emplace_back(occ, lab, fx); }
        else throw PharmsolError(psi::ST_OTHER, "unknown op " + k);
    }
    finish();
    return d;
}

extern "C" {

void* hs_new(const char* src) {
    auto* h = new HostSim();
    try { h->cm = dsl::compile_source(src); } catch (const std::exception& e) { h->error = e.what(); }
    return h;
}
void hs_free(void* p) { delete static_cast<HostSim*>(p); }
const char* hs_error(void* p) { return static_cast<HostSim*>(p)->error.c_str(); }
const char* hs_model_id(void* p) { return static_cast<HostSim*>(p)->cm.id.c_str(); }
int hs_model_kind(void* p) { return (int)static_cast<HostSim*>(p)->cm.kind; }
int hs_set_data(void* p, const char* ops) {
    auto* h = static_cast<HostSim*>(p);
    try { h->data = parse_ops(ops); return 0; } catch (const std::exception& e) { h->error = e.what(); return 15; }
}
long hs_nobs(void* p, const double* ems, int nem) {
    auto* h = static_cast<HostSim*>(p);
    try { return (long)flatten_population(h->data, h->cm.labels(), nullptr).nobs_total; } catch (const std::exception& e) { h->error = e.what(); return -1; }
}

// ems: nem rows of {kind, factor, c0, c1, c2, c3}; opts: {rtol, atol, cov_time, max_steps, h0, nparticles, seed, sde_mode, em_mode, em_dt, sde_normals}
// SDE models run with ONE thread per CTA (blockDim.x = 1): every cooperative loop of the kernel degenerates to a serial
// loop and the block reductions to one lane, so the same source runs unchanged (sums are formed in a different order
// than with 128 threads: results agree with the device to rounding, not bit for bit).
// out_ll: F-order (nsub x nspp); out_pred: (nobs x nspp) row-major or NULL; counters[4]
int hs_run(void* p, int solver, const double* spp_rows, long nspp, int np, const double* ems, int nem, const double* o, double* out_ll, double* out_pred,
           int* code, long* pair, unsigned long long* counters) {
    auto* h = static_cast<HostSim*>(p);
    try {
        AssayErrorModels em;
        for (int i = 0; i < nem; ++i) {
            AssayErrorModel m;
            m.kind = (ErrKind)(int)ems[6 * i];
            m.factor = ems[6 * i + 1];
            m.poly = ErrorPoly{ems[6 * i + 2], ems[6 * i + 3], ems[6 * i + 4], ems[6 * i + 5]};
            em.models.push_back(m);
        }
        FlatPopulation f = flatten_population(h->data, h->cm.labels(), nem > 0 ? &em : nullptr);
        psi::PopView v{};
        v.occ_offsets = f.occ_offsets.data(); v.occ_index = f.occ_index.data(); v.ev_offsets = f.ev_offsets.data(); v.events = f.events.data();
        v.bol_offsets = f.bol_offsets.data(); v.bol_event = f.bol_event.data(); v.inf_offsets = f.inf_offsets.data(); v.infs = f.infs.data();
        v.bnd_offsets = f.bnd_offsets.data(); v.bnds = f.bnds.data(); v.cov_offsets = f.cov_offsets.data(); v.cov_segs = f.cov_segs.data();
        v.occ_t0 = f.occ_t0.data(); v.nsub = f.nsub; v.ncov = f.ncov; v.max_events = f.max_events;
        v.prog_offsets = f.has_prog ? f.prog_offsets.data() : nullptr; v.prog = f.has_prog ? f.prog.data() : nullptr;
        v.prog_rates = f.has_prog ? f.prog_rates.data() : nullptr;
        v.prog_cov = f.prog_cov ? 1 : 0;
        std::vector<double> soa((size_t)np * nspp);
        for (long j = 0; j < nspp; ++j)
            for (int k = 0; k < np; ++k) soa[(size_t)k * nspp + j] = spp_rows[(size_t)j * np + k];
        psi::RunOpts opt;
        std::memset(&opt, 0, sizeof opt);
        opt.rtol = o[0]; opt.atol = o[1]; opt.cov_time = (int)o[2]; opt.max_steps = (int)o[3]; opt.h0 = o[4];
        opt.solver = solver; opt.want_ll = out_ll ? 1 : 0; opt.want_pred = out_pred ? 1 : 0;
        opt.nparticles = (int)o[5]; opt.seed = (unsigned long long)o[6]; opt.sde_mode = (int)o[7]; opt.em_mode = (int)o[8]; opt.em_dt = o[9]; opt.sde_normals = (int)o[10];
        unsigned long long status[5] = {~0ull, 0, 0, 0, 0};
        psi::OutView out{};
        out.ll = out_ll; out.ld_ll = f.nsub; out.pred = out_pred; out.ld_pred = nspp; out.first_error = status; out.counters = status + 1;
        const int s_eff = h->cm.kind == dsl::ModelKind::Ode ? solver : 0;
        const std::string name = "psi_entry_" + h->cm.id + "_s" + std::to_string(s_eff);
        entry_fn fn = (entry_fn)dlsym(RTLD_DEFAULT, name.c_str());
        if (!fn) { h->error = "entry " + name + " is not linked into this hostsim module"; return 15; }
        if (h->cm.kind == dsl::ModelKind::Sde) {
            const int np_ = opt.nparticles > 0 ? opt.nparticles : 1;
            const long long stride = pharmsol::sde_workspace_doubles(h->cm.state_len, np_);
            std::vector<double> scratch((size_t)stride);
            out.scratch = scratch.data();
            out.scratch_stride = 0;              // one CTA at a time: every CTA reuses the same slab
            blockDim.x = 1; gridDim.x = (unsigned)((long long)f.nsub * nspp); gridDim.y = 1;
            threadIdx.x = 0; blockIdx.y = 0;
            for (unsigned bx = 0; bx < gridDim.x; ++bx) { blockIdx.x = bx; fn(v, soa.data(), nspp, nspp, opt, out); }
            for (int k = 0; k < 4; ++k) counters[k] = status[1 + k];
            if (status[0] == ~0ull) { *code = 0; *pair = -1; }
            else { *code = (int)(status[0] & 0xff); *pair = (long)(status[0] >> 8); }
            return 0;
        }
        const pharmsol::LaunchGeometry g = pharmsol::psi_launch_geometry(f.nsub, nspp, false, 148, 128);
        opt.warp_tasks = g.warp_tasks;
        blockDim.x = g.block; gridDim.x = g.grid_x; gridDim.y = g.grid_y;
        for (unsigned by = 0; by < g.grid_y; ++by)
            for (unsigned bx = 0; bx < g.grid_x; ++bx)
                for (unsigned tx = 0; tx < g.block; ++tx) {
                    blockIdx.x = bx; blockIdx.y = by; threadIdx.x = tx;
                    fn(v, soa.data(), nspp, nspp, opt, out);
                }
        for (int k = 0; k < 4; ++k) counters[k] = status[1 + k];
        if (status[0] == ~0ull) { *code = 0; *pair = -1; }
        else { *code = (int)(status[0] & 0xff); *pair = (long)(status[0] >> 8); }
        return 0;
    } catch (const PharmsolError& e) {
        h->error = e.what();
        return e.code;
    } catch (const std::exception& e) {
        h->error = e.what();
        return 15;
    }
}
}
