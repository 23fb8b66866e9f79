// artifact.cpp — the CUDA-target model artifact (`.pkm`), SURVEY §8 f.2.
//
// The reference's native-AoT artifact is a cdylib carrying an API-version symbol, the model-info JSON and the
// compiled model functions (src/dsl/aot.rs:303-353, src/dsl/compiled_backend_abi.rs:6-33); `load_aot_model`
// checks the version, reads the info and binds the functions.  The CUDA counterpart is one container file with the
// same three parts: an API version, the model-info JSON (+ the run settings), and the device code — the sm_100a
// cubin(s) of the psi kernel for the chosen solver(s), compiled by NVRTC without a GPU.  The DSL source travels too,
// so the host-side metadata is rebuilt at load time and a cubin compiled against another engine version (different
// kernel-parameter layout) is recompiled instead of trusted.
//
// Layout (little endian): "PKMCUDA\0" | u32 api_version | u32 nsections | sections{u32 tag, u32 aux, u64 len, bytes}
// | u64 FNV-1a of everything before it.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

#include <cmath>

#include "runtime.hpp"

namespace pharmsol {

namespace {
constexpr char kMagic[8] = {'P', 'K', 'M', 'C', 'U', 'D', 'A', '\0'};
unsigned long long fnv(const char* p, size_t n) {
    unsigned long long h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) { h ^= (unsigned char)p[i]; h *= 1099511628211ull; }
    return h;
}
template <class T> void put(std::string& s, T v) { s.append(reinterpret_cast<const char*>(&v), sizeof v); }
void put_section(std::string& s, uint32_t tag, uint32_t aux, const char* p, size_t n) {
    put<uint32_t>(s, tag); put<uint32_t>(s, aux); put<uint64_t>(s, (uint64_t)n); s.append(p, n);
}
std::string settings_text(const psi::RunOpts& o) {
    std::ostringstream os;
    os.precision(17);
    os << "solver=" << o.solver << "\nrtol=" << o.rtol << "\natol=" << o.atol << "\nmax_steps=" << o.max_steps << "\nnparticles=" << o.nparticles
       << "\nseed=" << o.seed << "\nsde_mode=" << o.sde_mode << "\nem_mode=" << o.em_mode << "\nem_dt=" << o.em_dt << "\ncov_time=" << o.cov_time << "\nsde_normals=" << o.sde_normals << "\n";
    return os.str();
}
}  // namespace

std::string artifact_settings_json(const std::string& text) {
    std::istringstream is(text);
    std::string line, out = "{";
    bool first = true;
    while (std::getline(is, line)) {
        const size_t eq = line.find('=');
        if (eq == std::string::npos) continue;
        out += (first ? "\"" : ", \"") + line.substr(0, eq) + "\": " + line.substr(eq + 1);
        first = false;
    }
    return out + "}";
}

void apply_artifact_settings(const std::string& text, psi::RunOpts& o) {
    std::istringstream is(text);
    std::string line;
    while (std::getline(is, line)) {
        const size_t eq = line.find('=');
        if (eq == std::string::npos) continue;
        const std::string k = line.substr(0, eq), v = line.substr(eq + 1);
        if (k == "solver") o.solver = std::stoi(v);
        else if (k == "rtol") o.rtol = std::stod(v);
        else if (k == "atol") o.atol = std::stod(v);
        else if (k == "max_steps") o.max_steps = std::stoi(v);
        else if (k == "nparticles") o.nparticles = std::stoi(v);
        else if (k == "seed") o.seed = std::stoull(v);
        else if (k == "sde_mode") o.sde_mode = std::stoi(v);
        else if (k == "em_mode") o.em_mode = std::stoi(v);
        else if (k == "em_dt") o.em_dt = std::stod(v);
        else if (k == "cov_time") o.cov_time = std::stoi(v);
        else if (k == "sde_normals") o.sde_normals = std::stoi(v);
    }
    // the same ranges the pharmsol_cuda_model_set_* entry points enforce: the checksum only detects corruption, a
    // well-formed file can still carry values no setter would accept
    const bool ok = o.solver >= 0 && o.solver < psi::SOLVER_COUNT && o.rtol > 0.0 && o.atol > 0.0 && std::isfinite(o.rtol) && std::isfinite(o.atol) &&
                    o.max_steps > 0 && o.nparticles > 0 && o.sde_mode >= 0 && o.sde_mode <= 1 && o.em_mode >= 0 && o.em_mode <= 1 &&
                    o.em_dt > 0.0 && std::isfinite(o.em_dt) && o.cov_time >= 0 && o.cov_time <= 1 && o.sde_normals >= 0 && o.sde_normals <= 1;
    if (!ok) throw PharmsolError(psi::ST_OTHER, "artifact has invalid settings");
}

void write_artifact(Model& m, const std::string& path, const std::vector<int>& solvers) {
    std::string s(kMagic, 8);
    put<uint32_t>(s, PKM_API_VERSION);
    put<uint32_t>(s, (uint32_t)(4 + solvers.size()));
    put_section(s, PKM_INFO, 0, m.info_json.data(), m.info_json.size());
    put_section(s, PKM_SOURCE, 0, m.dsl_source.data(), m.dsl_source.size());
    const std::string st = settings_text(m.opts);
    put_section(s, PKM_SETTINGS, 0, st.data(), st.size());
    const std::string eng = std::string("sm_100a ") + engine_fingerprint();
    put_section(s, PKM_ENGINE, 0, eng.data(), eng.size());
    for (int solver : solvers) {
        const std::string name = entry_name(m.cm.id, solver);
        std::vector<char> cubin;
        {
            std::lock_guard<std::mutex> lk(m.mu);
            auto it = m.artifact_cubins.find(solver);
            if (it != m.artifact_cubins.end()) cubin = it->second;
        }
        if (cubin.empty()) cubin = nvrtc_compile_cubin(m.cm.cuda_source({{solver, name}}, false), name);
        put_section(s, PKM_CUBIN, (uint32_t)solver, cubin.data(), cubin.size());
    }
    put<uint64_t>(s, fnv(s.data(), s.size()));
    const std::string tmp = path + ".tmp";
    {
        std::ofstream f(tmp, std::ios::binary);
        if (!f) throw PharmsolError(psi::ST_OTHER, "cannot write " + path);
        f.write(s.data(), (std::streamsize)s.size());
        if (!f) throw PharmsolError(psi::ST_OTHER, "short write to " + path);
    }
    if (std::rename(tmp.c_str(), path.c_str()) != 0) throw PharmsolError(psi::ST_OTHER, "cannot move artifact into place: " + path);
}

ArtifactFile read_artifact(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw PharmsolError(psi::ST_OTHER, "cannot open artifact " + path);
    std::string s((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    if (s.size() < 24 || std::memcmp(s.data(), kMagic, 8) != 0) throw PharmsolError(psi::ST_OTHER, path + " is not a pharmsol CUDA artifact");
    uint64_t sum;
    std::memcpy(&sum, s.data() + s.size() - 8, 8);
    if (sum != fnv(s.data(), s.size() - 8)) throw PharmsolError(psi::ST_OTHER, "artifact " + path + " is corrupt (checksum mismatch)");
    ArtifactFile a;
    uint32_t nsec;
    std::memcpy(&a.api_version, s.data() + 8, 4);
    std::memcpy(&nsec, s.data() + 12, 4);
    if (a.api_version != PKM_API_VERSION)      // aot.rs:395-407 ApiVersionMismatch
        throw PharmsolError(psi::ST_OTHER, "artifact API version mismatch: expected " + std::to_string(PKM_API_VERSION) + ", found " + std::to_string(a.api_version));
    size_t off = 16;
    const size_t end = s.size() - 8;
    for (uint32_t i = 0; i < nsec; ++i) {
        if (off + 16 > end) throw PharmsolError(psi::ST_OTHER, "artifact " + path + " is truncated");
        uint32_t tag, aux; uint64_t len;
        std::memcpy(&tag, s.data() + off, 4); std::memcpy(&aux, s.data() + off + 4, 4); std::memcpy(&len, s.data() + off + 8, 8);
        off += 16;
        if (len > end - off) throw PharmsolError(psi::ST_OTHER, "artifact " + path + " is truncated");
        const char* p = s.data() + off;
        switch (tag) {
            case PKM_INFO: a.info_json.assign(p, len); break;
            case PKM_SOURCE: a.dsl_source.assign(p, len); break;
            case PKM_SETTINGS: a.settings.assign(p, len); break;
            case PKM_ENGINE: a.engine.assign(p, len); break;
            case PKM_CUBIN: a.cubins[(int)aux] = std::vector<char>(p, p + len); break;
            default: break;      // unknown sections are skipped (forward compatible within one API version)
        }
        off += len;
    }
    if (a.dsl_source.empty() || a.info_json.empty()) throw PharmsolError(psi::ST_OTHER, "artifact " + path + " lacks the model sections");
    a.engine_matches = a.engine == std::string("sm_100a ") + engine_fingerprint();
    return a;
}

std::string artifact_info_json(const ArtifactFile& a, const std::string& id) {
    std::ostringstream os;
    os << "{\"format\": \"pharmsol-cuda-pkm\", \"api_version\": " << a.api_version << ", \"engine\": \"" << a.engine << "\", \"engine_matches\": "
       << (a.engine_matches ? "true" : "false") << ", \"settings\": " << artifact_settings_json(a.settings) << ", \"kernels\": [";
    bool first = true;
    for (const auto& kv : a.cubins) {
        os << (first ? "" : ", ") << "{\"solver\": " << kv.first << ", \"entry\": \"" << (id.empty() ? std::string() : entry_name(id, kv.first)) << "\", \"cubin_bytes\": " << kv.second.size() << "}";
        first = false;
    }
    os << "], \"model\": " << a.info_json << "}";
    return os.str();
}

}  // namespace pharmsol
