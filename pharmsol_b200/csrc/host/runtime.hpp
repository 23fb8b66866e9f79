// runtime.hpp — context / model / population objects behind the C ABI.
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "data.hpp"
#include "dsl.hpp"

namespace pharmsol {

void set_last_error(const std::string& msg);
const std::string& last_error();

struct CudaError : std::runtime_error {
    explicit CudaError(const std::string& m) : std::runtime_error(m) {}
};
void cuda_check(cudaError_t e, const char* what);

// A grow-only device buffer.
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    void reserve(size_t bytes);
    void release();
    template <class T> T* as() const { return static_cast<T*>(p); }
};

enum KernelSource : int { SRC_AOT = 0, SRC_CACHE = 1, SRC_NVRTC = 2, SRC_ARTIFACT = 3 };
struct KernelRef {
    const void* fn = nullptr;      // function pointer (AOT) or cudaKernel_t (runtime-loaded library)
    int source = SRC_AOT;
    std::string name;
};

struct Ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    static constexpr int kMaxChunks = 16;
    cudaEvent_t chunk_ev[kMaxChunks] = {};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t reset_ev = nullptr;           // recorded behind the status reset that the latency path queues for the next call
    std::mutex mu;
    DevBuf err_ctr;          // [0] first_error (u64)  [1..4] counters
    unsigned long long* err_host = nullptr;   // pinned: [0..4] mirror of err_ctr filled by an async copy queued behind the kernels, [8..12] reset template
    double* small_host = nullptr;             // pinned staging for latency-bound calls: [0, kSmallIn) support points SoA, then psi
    static constexpr size_t kSmallIn = 64u << 10, kSmallOut = 256u << 10;
    DevBuf spp_rows, spp_soa, out, pred, scratch;
    DevBuf sde_tickets;      // 64 pair tickets (u64), one per SDE launch in rotation: launches of one context may overlap on different streams
    unsigned sde_ticket_next = 0;
    // device-resident, replicated psi of a multi-device context (pharmsol_cuda_log_likelihood_matrix_replicated) and the
    // streams / events of the copy-engine push gather: finished column chunks are pushed to the peers over NVLink by
    // cudaMemcpyAsync on these streams while the next chunk is computed
    DevBuf full;
    static constexpr int kPeerStreams = 4;
    cudaStream_t peer_streams[kPeerStreams] = {};
    cudaEvent_t peer_ev[kPeerStreams] = {};
    DevBuf col_work, col_idx, col_sort;      // work-balanced column order (ODE): probe counts, permutation, cub scratch
    int64_t launches = 0;
    bool status_batch = false;                // several launches share one error word / counter set until the next collect
    cudaStream_t status_clean_on = nullptr;   // stream on which a status reset is already queued behind the previous call (latency path)
    double last_kernel_ms = 0.0;
    unsigned long long last_counters[4] = {0, 0, 0, 0};
    int64_t pending_nsub = 0, pending_first_col = 0;
    ~Ctx();
};

struct Model {
    dsl::CompiledModel cm;
    psi::RunOpts opts;
    std::string source_cache;                 // generated CUDA C (for inspection)
    std::string host_source_cache;            // generated host twin (frozen compiled-backend ABI), built on demand
    std::string info_json;
    std::map<int, KernelRef> kernels;         // by solver id
    std::string dsl_source;                   // as given to pharmsol_cuda_model_from_dsl (travels in the artifact)
    std::map<int, std::vector<char>> artifact_cubins;   // device code that arrived in a .pkm artifact, by solver id
    std::mutex mu;
};

// The flattened population on one more device of a multi-device context (the primary copy lives in Population itself).
struct PopReplica {
    int device = 0;
    DevBuf dev;
    psi::PopView view{};
};
struct Population {
    Data data;
    ModelLabels labels;
    FlatPopulation flat;
    DevBuf dev;
    psi::PopView view{};
    int device = 0;
    std::vector<std::unique_ptr<PopReplica>> replicas;      // devices 1.. of the creating context, in context order
    void upload();                                            // primary device + every replica
    void upload_into(DevBuf& dev, psi::PopView& view) const;  // current device
    const psi::PopView& view_on(int ctx_index) const { return ctx_index == 0 ? view : replicas[(size_t)ctx_index - 1]->view; }
    ~Population();
};

// --- module management --------------------------------------------------------------------------------
// Look up / build the kernel for (model, solver): AOT registry -> cubin cache -> NVRTC.
KernelRef get_kernel(Model& m, int solver);
// NVRTC compile to cubin (no device needed); throws dsl::DslError with the compile log on failure.
std::vector<char> nvrtc_compile_cubin(const std::string& source, const std::string& name);
std::string cubin_cache_path(const std::string& id, int solver);
std::string engine_fingerprint();

// --- CUDA-target model artifact (.pkm), artifact.cpp ---------------------------------------------------
constexpr uint32_t PKM_API_VERSION = 1;
enum PkmSection : uint32_t { PKM_INFO = 1, PKM_SOURCE = 2, PKM_SETTINGS = 3, PKM_ENGINE = 4, PKM_CUBIN = 5 };
struct ArtifactFile {
    uint32_t api_version = 0;
    std::string info_json, dsl_source, settings, engine;
    std::map<int, std::vector<char>> cubins;
    bool engine_matches = false;
};
void write_artifact(Model& m, const std::string& path, const std::vector<int>& solvers);
ArtifactFile read_artifact(const std::string& path);
std::string artifact_info_json(const ArtifactFile& a, const std::string& model_id);
void apply_artifact_settings(const std::string& text, psi::RunOpts& o);
std::string entry_name(const std::string& id, int solver);
int effective_solver(const Model& m);

// --- launches -------------------------------------------------------------------------------------------
void launch_psi(Ctx& ctx, Model& m, Population& pop, const double* spp_soa_dev, int64_t ncols, int64_t ld_spp,
                double* out_dev, int64_t ld_out, double* pred_dev, int64_t ld_pred, int64_t first_col, cudaStream_t stream,
                const psi::RunOpts* opts_override = nullptr, bool reset_status = true, double* const* peers = nullptr, int npeers = 0,
                const psi::PopView* view = nullptr);
// Column chunks of one shard for pipelining (copy-back or peer push behind the compute): closed-form models up to
// `max_chunks` equal chunks of >= 2 MB and >= 128 columns; adaptive ODE / SDE models 7/8 + 1/8 (every extra launch costs
// a probe + sort of the work-balanced column order and a kernel tail).  Returns the cut points, first 0, last ncols.
std::vector<int64_t> column_chunks(const Model& m, int64_t nsub, int64_t ncols, int max_chunks);
// Copy-engine push gather: evaluate [first_col, first_col + ncols) chunk by chunk into peers[self] (a FULL column-major
// matrix) and push every finished chunk to the other peers with device-to-device copies on the context's peer streams;
// `stream` ends up ordered after the last push.
void launch_psi_push(Ctx& ctx, Model& m, Population& pop, const psi::PopView& view, const double* spp_soa_dev, int64_t ncols, int64_t ld_spp,
                     double* const* peers, int npeers, int self, int64_t ld_out, int64_t first_col, cudaStream_t stream);
void launch_transpose(const double* rows, double* soa, int64_t nspp, int nparams, int64_t ld, cudaStream_t stream);
void launch_exp_inplace(double* p, int64_t n, cudaStream_t stream);
double measure_fp64_peak(Ctx& ctx, double* clock_mhz);

}  // namespace pharmsol
