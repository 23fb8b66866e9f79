/* pharmsol_cuda.h — C ABI of the B200-native psi-matrix backend for LAPKB/pharmsol.
 *
 * This is the drop-in boundary a `simulator::cuda` backend inside pharmsol binds over FFI (see
 * INTEGRATION.md for the Rust `extern "C"` block and the `impl Equation` glue).  All citations are
 * file:line relative to the pharmsol source tree (v0.28.8).
 *
 * Conventions
 *   - every function returns int32_t: 0 = ok, otherwise a PCU_ERR_* code that maps 1:1 onto a
 *     `PharmsolError` / `ErrorModelError` variant (src/error/mod.rs:14-49);
 *   - out-parameters by pointer; host buffers are caller-owned; device buffers live behind the
 *     opaque handles and are freed by the matching *_destroy / *_free;
 *   - handles are safe to use from several host threads, matching `Equation: Sync`
 *     (src/simulator/equation/mod.rs:377); concurrent host-buffer calls on one context run on separate
 *     stream "lanes" (pharmsol_cuda_ctx_num_lanes) instead of queueing behind one lock;
 *   - nothing here ever aborts or traps: the reference's `panic!`s on this path (imaginary roots,
 *     missing covariates, particle-filter likelihood errors) become status codes;
 *   - there is NO CPU fallback: every compute entry point fails with PCU_ERR_CUDA when no sm_100
 *     device / driver is present.
 *   - all arithmetic is FP64.
 */
#ifndef PHARMSOL_CUDA_H
#define PHARMSOL_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PHARMSOL_CUDA_ABI_VERSION 2

/* ---- status codes (== PharmsolError variants, src/error/mod.rs:14-49) ------------------------ */
enum {
    PCU_OK = 0,
    PCU_ERR_NON_FINITE_LIKELIHOOD = 1,      /* PharmsolError::NonFiniteLikelihood          prediction.rs:120-124 */
    PCU_ERR_NEGATIVE_SIGMA = 2,             /* ErrorModelError::NegativeSigma              error_model.rs:1073   */
    PCU_ERR_NON_FINITE_SIGMA = 3,           /* ErrorModelError::NonFiniteSigma             error_model.rs:1075   */
    PCU_ERR_INVALID_OUTPUT_EQUATION = 4,    /* ErrorModelError::InvalidOutputEquation      error_model.rs:679    */
    PCU_ERR_NONE_ERROR_MODEL = 5,           /* ErrorModelError::NoneErrorModel             error_model.rs:682    */
    PCU_ERR_MISSING_ERROR_MODEL = 6,        /* ErrorModelError::MissingErrorModel          error_model.rs:1067   */
    PCU_ERR_SOLVER_FAILURE = 7,             /* PharmsolError::DiffsolError (from_solver_error)                  */
    PCU_ERR_INPUT_OUT_OF_RANGE = 8,         /* PharmsolError::InputOutOfRange                                    */
    PCU_ERR_OUTEQ_OUT_OF_RANGE = 9,         /* PharmsolError::OuteqOutOfRange                                    */
    PCU_ERR_UNKNOWN_INPUT_LABEL = 10,       /* PharmsolError::UnknownInputLabel                                  */
    PCU_ERR_UNKNOWN_OUTPUT_LABEL = 11,      /* PharmsolError::UnknownOutputLabel                                 */
    PCU_ERR_IMAGINARY_ROOTS = 12,           /* replaces panic!  two_compartment_models.rs:20-22, three_...:32-34 */
    PCU_ERR_UNSUPPORTED_INPUT_ROUTE_KIND = 13, /* PharmsolError::UnsupportedInputRouteKind                       */
    PCU_ERR_MISSING_COVARIATE = 14,
    PCU_ERR_OTHER = 15,                     /* PharmsolError::OtherError (message in last_error_message)         */
    /* library-level */
    PCU_ERR_CUDA = 64,                      /* CUDA runtime / driver failure, or no device                       */
    PCU_ERR_COMPILE = 65,                   /* DSL parse / analysis error or NVRTC compile error                 */
    PCU_ERR_INVALID_ARGUMENT = 66
};

/* ---- enums ------------------------------------------------------------------------------------ */
enum { PCU_KIND_ODE = 0, PCU_KIND_ANALYTICAL = 1, PCU_KIND_SDE = 2 };        /* EqnKind, equation/mod.rs:580-586 */
enum { PCU_CENSOR_NONE = 0, PCU_CENSOR_BLOQ = 1, PCU_CENSOR_ALOQ = 2 };      /* Censor, data/event.rs:559-567    */
enum { PCU_ERRMODEL_NONE = 0, PCU_ERRMODEL_ADDITIVE = 1, PCU_ERRMODEL_PROPORTIONAL = 2 }; /* error_model.rs:786 */
/* OdeSolver (ode/mod.rs:59-84).  The reference offers Bdf | Sdirk(TrBdf2|Esdirk34) | ExplicitRk(Tsit45)
 * through diffsol; this backend offers two explicit pairs, two SDIRK methods and a Rosenbrock method (RODAS4), all
 * register-resident per thread. */
enum { PCU_SOLVER_DOPRI5 = 0, PCU_SOLVER_TSIT45 = 1, PCU_SOLVER_SDIRK4 = 2, PCU_SOLVER_TRBDF2 = 3, PCU_SOLVER_RODAS4 = 4,
       PCU_SOLVER_BDF = 5,        /* OdeSolver::Bdf (the reference default): variable-order NDF/BDF 1-5, the algorithm diffsol's `bdf` documents */
       PCU_SOLVER_ESDIRK34 = 6 }; /* OdeSolver::Sdirk(SdirkTableau::Esdirk34) */
/* Analytical `derive` time semantics: sub-interval END (DSL runtime,
 * dsl/native.rs:1903-1916) or sub-interval LENGTH (analytical! macro, analytical/mod.rs:362-364). */
enum { PCU_COVTIME_INTERVAL_END = 0, PCU_COVTIME_INTERVAL_LENGTH = 1 };
/* SDE likelihood mode: what log_likelihood_matrix does (mean prediction,
 * sde/mod.rs:387-433) or SDE::estimate_log_likelihood (particle filter, sde/mod.rs:526-577). */
enum { PCU_SDE_MEAN_PREDICTION = 0, PCU_SDE_PARTICLE_FILTER = 1 };
enum { PCU_EM_REFERENCE_ADAPTIVE = 0, PCU_EM_FIXED_STEP = 1 };               /* sde/em.rs:134-167 vs fixed dt     */

typedef struct pcu_ctx pcu_ctx;
typedef struct pcu_model pcu_model;
typedef struct pcu_subject_builder pcu_subject_builder;
typedef struct pcu_subject pcu_subject;
typedef struct pcu_data pcu_data;
typedef struct pcu_population pcu_population;

/* AssayErrorModel for one output equation (data/error_model.rs:786-812): sigma is computed from
 * the OBSERVATION: alpha = c0 + c1 o + c2 o^2 + c3 o^3; additive sqrt(alpha^2 + factor^2);
 * proportional factor * alpha (error_model.rs:1045-1080). */
typedef struct pcu_error_model {
    int32_t kind;        /* PCU_ERRMODEL_* */
    int32_t pad;
    double factor;       /* lambda (additive) or gamma (proportional) */
    double c0, c1, c2, c3;
} pcu_error_model;

/* ResidualErrorModel for one output equation (data/residual_error.rs:69-139): sigma is computed from the
 * PREDICTION (parametric algorithms: SAEM, FOCE).  kind: PCU_RESID_*; constant: a; proportional: b;
 * combined: sqrt(a^2 + b^2 f^2); exponential: sigma = a.  sigma is floored at sqrt(DBL_EPSILON). */
enum { PCU_RESID_MISSING = 0, PCU_RESID_CONSTANT = 1, PCU_RESID_PROPORTIONAL = 2, PCU_RESID_COMBINED = 3, PCU_RESID_EXPONENTIAL = 4 };
typedef struct pcu_residual_error_model {
    int32_t kind;
    int32_t pad;
    double a, b;
} pcu_residual_error_model;

/* ---- library / context -------------------------------------------------------------------------- */
int32_t pharmsol_cuda_abi_version(void);
int32_t pharmsol_cuda_device_count(int32_t* n);
int32_t pharmsol_cuda_ctx_create(int32_t device, pcu_ctx** out);
/* SURVEY §8b / north_star "support-point columns shard across the 8 GPUs of one box": ONE host process drives every
 * listed device.  With such a context
 *   - pharmsol_cuda_population_create replicates the flattened population on every device;
 *   - the host-buffer entry points (log_likelihood_matrix, psi, predictions) split the support-point columns into
 *     n_dev contiguous blocks — psi is F-order (matrix.rs:60), so each block is one contiguous slab of the caller's
 *     matrix — and every device copies its slab straight to its offset in `out` (no gather needed for a host result);
 *   - pharmsol_cuda_log_likelihood_matrix_replicated leaves the full psi resident on EVERY device, gathered over NVLink;
 *   - the *_device / *_peers / *_push entry points and the latency path keep running on device_ids[0].
 * The first failing pair over all devices is reported (matrix.rs:96-104); SDE random streams are keyed by the global
 * pair, so results do not depend on the device count.  A device may be listed more than once (several column shards on
 * one GPU).  rayon callers (`Equation: Sync`) share one context: calls are serialised by its lock. */
int32_t pharmsol_cuda_ctx_create_multi(const int32_t* device_ids, int32_t n_dev, pcu_ctx** out);
int32_t pharmsol_cuda_ctx_num_devices(pcu_ctx* ctx);
int32_t pharmsol_cuda_ctx_device_id(pcu_ctx* ctx, int32_t k);             /* -1 if k is out of range */
/* Concurrency of one context (`Equation: Sync`: one shared handle is used from all rayon threads, equation/mod.rs:377): the
 * host-buffer calls of a single-device context do not queue behind one lock — a caller that finds the context busy takes
 * (or creates, up to 4) another "lane" on the same device with its own streams, staging and status buffers, so matrices
 * requested by different host threads overlap on the GPU.  Returns how many lanes exist (>= 1). */
int32_t pharmsol_cuda_ctx_num_lanes(pcu_ctx* ctx);
void    pharmsol_cuda_ctx_destroy(pcu_ctx* ctx);
/* message of the last failure on this thread (valid until the next failing call on the thread) */
const char* pharmsol_cuda_last_error_message(void);
/* kernel launches issued by this context so far; device time (ms, CUDA events on the launch stream)
 * and work counters {accepted steps, rejected steps, rhs/kernel evaluations, Newton iterations}
 * of the most recent psi launch */
int64_t pharmsol_cuda_launch_count(pcu_ctx* ctx);
double  pharmsol_cuda_last_kernel_ms(pcu_ctx* ctx);
int32_t pharmsol_cuda_last_counters(pcu_ctx* ctx, uint64_t out[4]);
/* pinned host memory for the caller's support-point / psi buffers */
int32_t pharmsol_cuda_host_alloc(size_t bytes, void** out);
int32_t pharmsol_cuda_host_free(void* p);

/* ---- data: Subject::builder (src/data/builder.rs:84-362) ------------------------------------------ */
pcu_subject_builder* pharmsol_subject_builder_new(const char* id);
void pharmsol_subject_builder_bolus(pcu_subject_builder* b, double time, double amount, const char* input);
void pharmsol_subject_builder_infusion(pcu_subject_builder* b, double time, double amount, const char* input, double duration);
void pharmsol_subject_builder_observation(pcu_subject_builder* b, double time, double value, const char* outeq);
void pharmsol_subject_builder_censored_observation(pcu_subject_builder* b, double time, double value, const char* outeq, int32_t censoring);
void pharmsol_subject_builder_missing_observation(pcu_subject_builder* b, double time, const char* outeq);
void pharmsol_subject_builder_observation_with_error(pcu_subject_builder* b, double time, double value, const char* outeq,
                                                     double c0, double c1, double c2, double c3, int32_t censoring);
void pharmsol_subject_builder_covariate(pcu_subject_builder* b, const char* name, double time, double value);
void pharmsol_subject_builder_repeat(pcu_subject_builder* b, int64_t n, double delta);
void pharmsol_subject_builder_reset(pcu_subject_builder* b);            /* next occasion */
/* consumes the builder; NULL (message in pharmsol_cuda_last_error_message) if any builder call before it failed:
 * NULL label, censoring outside 0..2, repeat count outside [0, 10^7], out of memory */
pcu_subject* pharmsol_subject_builder_build(pcu_subject_builder* b);
/* Covariate::set_fixed (data/covariate.rs:243-248): carry-forward instead of linear interpolation */
int32_t pharmsol_subject_set_covariate_fixed(pcu_subject* s, int32_t occasion, const char* name, int32_t fixed);
void pharmsol_subject_free(pcu_subject* s);
pcu_data* pharmsol_data_new(void);                                      /* Data::new (data/structs.rs:38) */
int32_t pharmsol_data_add_subject(pcu_data* d, const pcu_subject* s);   /* copies */
int64_t pharmsol_data_len(const pcu_data* d);
/* read_pmetrics (src/data/parser/pmetrics/mod.rs:164-239, row.rs:269-381, 593-672): Pmetrics CSV -> Data.
 * ID / EVID / TIME required; DUR DOSE ADDL II INPUT OUT OUTEQ CENS C0..C3 optional; every other column is a
 * covariate (`name!` = carry-forward); ADDL/II doses are expanded, EVID = 4 starts a new occasion,
 * OUT = -99 is a missing observation.  Errors: PCU_ERR_OTHER with the message in last_error_message. */
int32_t pharmsol_data_read_pmetrics(const char* path, pcu_data** out);
int32_t pharmsol_data_from_pmetrics_text(const char* text, size_t len, pcu_data** out);
/* Data::expand (src/data/structs.rs:155-260): a copy with missing observations added every `idelta` from 0 to the
 * last dose end + `tad` (per occasion, every output label of the dataset) — dense prediction grids. */
int32_t pharmsol_data_expand(const pcu_data* d, double idelta, double tad, pcu_data** out);
/* JSON description (subjects -> occasions -> events, covariates) for inspection; returns the number of bytes
 * needed (excluding the terminator) and writes at most cap - 1 bytes + NUL into buf (buf may be NULL). */
int64_t pharmsol_data_describe_json(const pcu_data* d, char* buf, size_t cap);
void pharmsol_data_free(pcu_data* d);

/* ---- models ----------------------------------------------------------------------------------------- */
/* Compile a pharmsol-dsl source (authoring shorthand or canonical `model {}` form) to a device model:
 * DSL -> IR -> CUDA C -> (AOT registry | on-disk cubin cache | NVRTC) -> module.
 * Counterpart of compile_module_source_to_runtime(..., RuntimeCompilationTarget::*) src/dsl/runtime.rs:118-245.
 * Parsing and code generation happen here; the device module is built lazily at the first launch
 * (or by pharmsol_cuda_model_compile). */
int32_t pharmsol_cuda_model_from_dsl(pcu_ctx* ctx, const char* source, size_t len, pcu_model** out);
void    pharmsol_cuda_model_destroy(pcu_model* m);
int32_t pharmsol_cuda_model_kind(const pcu_model* m);                   /* PCU_KIND_* */
int32_t pharmsol_cuda_model_nparams(const pcu_model* m);
int32_t pharmsol_cuda_model_nstates(const pcu_model* m);
int32_t pharmsol_cuda_model_nouteqs(const pcu_model* m);
const char* pharmsol_cuda_model_info_json(const pcu_model* m);          /* NativeModelInfo mirror, dsl/model_info.rs:17-92 */
const char* pharmsol_cuda_model_cuda_source(const pcu_model* m);        /* the generated CUDA C translation unit */
const char* pharmsol_cuda_model_id(const pcu_model* m);
/* with_solver / with_tolerances (ode/mod.rs:135-166); defaults Dopri5, rtol = atol = 1e-4 (ode/mod.rs:40-41) */
int32_t pharmsol_cuda_model_set_solver(pcu_model* m, int32_t solver, double rtol, double atol);
int32_t pharmsol_cuda_model_set_max_steps(pcu_model* m, int32_t max_steps);
/* with_particles (dsl/native.rs:2162) + stream seed + likelihood / stepper modes */
int32_t pharmsol_cuda_model_set_particles(pcu_model* m, uint32_t nparticles, uint64_t seed, int32_t sde_mode,
                                          int32_t em_mode, double em_dt);
/* Precision of the SDE noise draws: PCU_SDE_NORMALS_FP32 (default) = Box-Muller in FP32 on 24-bit Philox uniforms
 * (|z| <= 5.77, SFU pipe); PCU_SDE_NORMALS_FP64 = the same transform in FP64 on 32-bit uniforms (|z| <= 6.66), ~the
 * f64 `Normal` of the reference (sde/em.rs:104-120).  State, drift, diffusion and likelihood arithmetic is FP64 either way. */
enum { PCU_SDE_NORMALS_FP32 = 0, PCU_SDE_NORMALS_FP64 = 1 };
int32_t pharmsol_cuda_model_set_sde_normals(pcu_model* m, int32_t precision);
int32_t pharmsol_cuda_model_set_cov_time(pcu_model* m, int32_t cov_time);
/* build (or fetch) the device module now; returns the NVRTC log in last_error_message on failure.
 * *source_out (optional): 0 = ahead-of-time (nvcc, linked in), 1 = cubin cache, 2 = NVRTC, 3 = .pkm artifact */
int32_t pharmsol_cuda_model_compile(pcu_ctx* ctx, pcu_model* m, int32_t* source_out);
/* NVRTC-only: compile to a cubin without touching a device (used by the build step) */
int32_t pharmsol_cuda_model_precompile_to_cache(pcu_model* m, int32_t solver);

/* ---- CUDA-target model artifact (.pkm) ----------------------------------------------------------------
 * Counterpart of the native-AoT artifact: compile_module_source_to_aot / load_aot_model / read_aot_model_info
 * (src/dsl/aot.rs:146-353, symbols src/dsl/compiled_backend_abi.rs:6-33) and RuntimeCompilationTarget /
 * load_runtime_artifact (src/dsl/runtime.rs:118-137).  One file carries the API version, the model-info JSON, the
 * run settings, the DSL source and the sm_100a cubin of the psi kernel for each requested solver (NVRTC, no GPU
 * needed to export).  `solvers` = PCU_SOLVER_* list (ODE models; ignored otherwise); nsolvers = 0 exports the
 * model's current solver.  A loaded model launches the shipped device code directly (pharmsol_cuda_model_compile
 * reports source 3); if the artifact was built by another engine version the cubin is ignored and the model is
 * rebuilt from the source it carries.  Errors: unreadable / corrupt / API-version mismatch -> PCU_ERR_OTHER. */
int32_t pharmsol_cuda_model_export_artifact(pcu_model* m, const char* path, const int32_t* solvers, int32_t nsolvers);
int32_t pharmsol_cuda_model_load_artifact(pcu_ctx* ctx, const char* path, pcu_model** out);
/* read_aot_model_info: JSON {format, api_version, engine, engine_matches, settings, kernels[], model}; returns the
 * byte count needed (excluding the terminator), -1 on error; writes at most cap - 1 bytes + NUL into buf. */
int64_t pharmsol_cuda_artifact_info_json(const char* path, char* buf, size_t cap);

/* ---- native (host) artifact with the reference's frozen compiled-backend ABI ---------------------------------------
 * SURVEY §8 f.2 as written: a cdylib that the reference's own `load_aot_model` (src/dsl/aot.rs:316-353) can open. The
 * DSL emitter writes the HOST twin of the model it emits for the device — the same function bodies, C++ instead of CUDA
 * C — exporting the frozen symbols of src/dsl/compiled_backend_abi.rs:6-33:
 *     uint32_t pharmsol_dsl_api_version(void)                    == AOT_API_VERSION = 2 (aot.rs:43, 404-417)
 *     const uint8_t* pharmsol_dsl_model_info_json_ptr(void), size_t pharmsol_dsl_model_info_json_len(void)
 *                                                                 CompiledModelInfoEnvelope{abi_version, model, functions}
 *     void pharmsol_dsl_kernel_{outputs,derive,dynamics,init,drift,diffusion,route_lag,route_bioavailability}
 *          (double t, const double* states, const double* params, const double* covariates, const double* routes,
 *           const double* derived, double* out)                   CompiledModelFunction, dsl/native.rs:45-53
 * (only the roles the model has; `outputs` always).  `*_host_source` returns the generated C++ translation unit (owned
 * by the model); `*_export_host_artifact` compiles it with the system C++ compiler ($PHARMSOL_B200_CXX, default g++) the
 * way the reference shells out to cargo.  No GPU is involved. */
const char* pharmsol_cuda_model_host_source(pcu_model* m);
int32_t pharmsol_cuda_model_export_host_artifact(pcu_model* m, const char* path);

/* ---- population: flattened Data resident in HBM ------------------------------------------------------ */
/* Resolve labels against the model's routes / outputs (equation/mod.rs:192-273, dsl/native.rs:663-770),
 * precompute the per-observation sigma terms from `error_models` (may be NULL for predictions only),
 * flatten into the SoA buffers of csrc/device/psi_types.h and upload. */
int32_t pharmsol_cuda_population_create(pcu_ctx* ctx, const pcu_model* m, const pcu_data* d,
                                        const pcu_error_model* error_models, int32_t n_error_models,
                                        pcu_population** out);
int32_t pharmsol_cuda_population_set_error_models(pcu_population* pop, const pcu_error_model* error_models, int32_t n);
void    pharmsol_cuda_population_destroy(pcu_population* pop);
int64_t pharmsol_cuda_population_nsubjects(const pcu_population* pop);
int64_t pharmsol_cuda_population_nobservations(const pcu_population* pop);   /* prediction rows */
/* prefix sums of per-subject observation counts, nsub+1 entries (ragged prediction layout) */
int32_t pharmsol_cuda_population_obs_offsets(const pcu_population* pop, int64_t* out);
/* per prediction row (same order as the rows of pharmsol_cuda_predictions): the Observation the row belongs to —
 * time, observed value (NaN = missing), output equation, occasion index, censoring (Prediction, likelihood/prediction.rs:18-27).
 * Any pointer may be NULL. */
int32_t pharmsol_cuda_population_observation_table(const pcu_population* pop, double* time, double* value, int32_t* outeq,
                                                   int32_t* occasion, int32_t* censoring);
int64_t pharmsol_cuda_population_device_bytes(const pcu_population* pop);

/* ---- the hot path ------------------------------------------------------------------------------------- */
/* log_likelihood_matrix (src/simulator/likelihood/matrix.rs:52-106).
 *   support_points  host, row-major (nspp x nparams): rows = support points, cols = parameters in model order
 *   out             host, column-major / F-order (nsub x nspp)          (matrix.rs:60)
 * The first failing pair aborts the result like matrix.rs:96-104: the call returns that pair's code,
 * *first_error_pair = i + j*nsub (either pointer may be NULL), and `out` holds NaN for failing pairs. */
int32_t pharmsol_cuda_log_likelihood_matrix(pcu_ctx* ctx, pcu_model* m, pcu_population* pop,
                                            const double* support_points, int64_t nspp, int32_t nparams,
                                            double* out, int32_t* first_error_code, int64_t* first_error_pair);
/* Same with everything resident in HBM (no copies in the call):
 *   spp_soa_dev   device, parameter-major SoA: spp[k*ld_spp + j]
 *   out_dev       device, column-major: out[i + j*ld_out], ld_out >= nsub
 *   stream        cudaStream_t to launch on (NULL = the context's own non-blocking stream; pass
 *                 cudaStreamLegacy / cudaStreamPerThread for the default streams); the call is asynchronous,
 *                 errors of the launch are collected by pharmsol_cuda_collect_errors after a sync. */
int32_t pharmsol_cuda_log_likelihood_matrix_device(pcu_ctx* ctx, pcu_model* m, pcu_population* pop,
                                                   const double* spp_soa_dev, int64_t ncols, int64_t ld_spp,
                                                   double* out_dev, int64_t ld_out, int64_t first_col, void* stream);
/* Column-sharded psi with the all-gather FUSED into the kernel: every result is stored directly into the full
 * column-major psi of every rank through peer-mapped pointers (NVLink / NVSwitch), at global column
 * first_col + j, so no separate collective copies slabs afterwards.
 *   out_full_peers   host array of npeers (<= 8) DEVICE pointers, one full (ld_out x nspp_total) matrix per rank,
 *                    each dereferenceable from this device (CUDA IPC / symmetric memory / cudaDeviceEnablePeerAccess)
 * The caller synchronises the ranks (a device barrier) before any rank reads its matrix. */
int32_t pharmsol_cuda_log_likelihood_matrix_peers(pcu_ctx* ctx, pcu_model* m, pcu_population* pop,
                                                  const double* spp_soa_dev, int64_t ncols, int64_t ld_spp,
                                                  double* const* out_full_peers, int32_t npeers, int64_t ld_out,
                                                  int64_t first_col, void* stream);
/* Same sharding, gather by the COPY ENGINES instead of by stores from the SMs: the shard is evaluated chunk by chunk
 * into out_full_peers[self] (this rank's full matrix) and every finished chunk is pushed to the other ranks' matrices
 * with device-to-device copies over NVLink while the next chunk is computed; `stream` ends up ordered after the last
 * push.  Closed-form models produce psi at GB/s rates, where 8-byte peer stores (one NVLink packet each) or an NCCL
 * kernel queued behind the psi CTAs cost 14-40 % of the step; bulk copies on the copy engines overlap with the compute. */
int32_t pharmsol_cuda_log_likelihood_matrix_push(pcu_ctx* ctx, pcu_model* m, pcu_population* pop,
                                                 const double* spp_soa_dev, int64_t ncols, int64_t ld_spp,
                                                 double* const* out_full_peers, int32_t npeers, int32_t self, int64_t ld_out,
                                                 int64_t first_col, void* stream);
/* Multi-device context: host support points in, the whole psi resident on every device out.
 *   gather    PCU_GATHER_COPY_ENGINE (chunked pushes, see *_push) or PCU_GATHER_PEER_STORES (see *_peers)
 *   dev_out   host array of pharmsol_cuda_ctx_num_devices(ctx) pointers; dev_out[k] receives the library-owned
 *             column-major (nsub x nspp) matrix on device k, valid until the next replicated call on the context
 * Returns after every device's matrix is complete. */
enum { PCU_GATHER_COPY_ENGINE = 0, PCU_GATHER_PEER_STORES = 1 };
int32_t pharmsol_cuda_log_likelihood_matrix_replicated(pcu_ctx* ctx, pcu_model* m, pcu_population* pop,
                                                       const double* support_points, int64_t nspp, int32_t nparams, int32_t gather,
                                                       double** dev_out, int32_t* first_error_code, int64_t* first_error_pair);
int32_t pharmsol_cuda_collect_errors(pcu_ctx* ctx, int32_t* first_error_code, int64_t* first_error_pair);
/* Several asynchronous *_device / *_peers launches, one collect: after this call (which resets the status on
 * `stream`) the launches of the context share one error word and counter set until the next
 * pharmsol_cuda_collect_errors, which reports the first failing pair over all of them (matrix.rs:96-104). */
int32_t pharmsol_cuda_status_batch_begin(pcu_ctx* ctx, void* stream);
/* row-major host support points -> SoA device buffer (H2D + on-device transpose) */
int32_t pharmsol_cuda_upload_support_points(pcu_ctx* ctx, const double* support_points, int64_t nspp, int32_t nparams,
                                            double* spp_soa_dev, int64_t ld_spp, void* stream);
/* estimate_predictions for every (subject, support point) pair (equation/mod.rs:526-532):
 *   out  host, (nobs_total x nspp) row-major: out[row*nspp + j], row = obs_offsets[i] + k-th observation of subject i
 * Missing observations are included (they are prediction slots). */
int32_t pharmsol_cuda_predictions(pcu_ctx* ctx, pcu_model* m, pcu_population* pop,
                                  const double* support_points, int64_t nspp, int32_t nparams, double* out);
int32_t pharmsol_cuda_predictions_device(pcu_ctx* ctx, pcu_model* m, pcu_population* pop,
                                         const double* spp_soa_dev, int64_t ncols, int64_t ld_spp,
                                         double* pred_dev, int64_t ld_pred, double* ll_dev_or_null, int64_t ld_out, void* stream);
/* log_likelihood_batch (src/simulator/likelihood/mod.rs:119-177): one parameter row per subject
 *   parameters  host, row-major (nrows x nparams); nrows must equal the number of subjects (else PCU_ERR_OTHER)
 *   out         host, nsub log-likelihoods; -inf for a subject whose simulation fails or whose output has no model
 * The population may be created without assay error models (they are not used here). */
int32_t pharmsol_cuda_log_likelihood_batch(pcu_ctx* ctx, pcu_model* m, pcu_population* pop, const double* parameters,
                                           int64_t nrows, int32_t nparams, const pcu_residual_error_model* models,
                                           int32_t n_models, double* out);
/* psi / log_psi deprecated wrappers (matrix.rs:117-150): exp of the log matrix, on device */
int32_t pharmsol_cuda_psi(pcu_ctx* ctx, pcu_model* m, pcu_population* pop, const double* support_points, int64_t nspp,
                          int32_t nparams, double* out, int32_t* first_error_code, int64_t* first_error_pair);

/* measured FP64 FMA throughput of this device (register-resident DFMA chains), TFLOP/s */
int32_t pharmsol_cuda_measure_fp64_peak(pcu_ctx* ctx, double* tflops, double* sm_clock_mhz);

#ifdef __cplusplus
}
#endif
#endif /* PHARMSOL_CUDA_H */
