// psi_types.h — plain-data layouts shared by the host (flattener, launcher) and the device engine.
//
// Everything the kernels read is a flat device buffer described here.  The population layout is
// "struct of arrays of small records": per-thread data (support points) is SoA so a warp's loads
// coalesce; per-subject data (the event timeline) is read at a warp-uniform address, so it is kept
// as fixed-size records that one or two 128-bit broadcast loads fetch.
#pragma once
#if defined(__CUDACC_RTC__)
// NVRTC has no standard headers: define the fixed-width types the layouts use.
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
#else
#include <stdint.h>
#endif
#if !defined(__CUDACC__) && !defined(__CUDACC_RTC__)
#define __host__
#define __device__
#endif

namespace psi {

// ---- event kinds / flags (order = the reference's tie-break rank, data/event.rs:292-304) -------
enum : int { EV_OBS = 0, EV_BOLUS = 1, EV_INFUSION = 2,
             EV_STEP = 3 };   // timeline-program record only (PopView::prog): one propagation sub-interval, see below
enum : int { CENS_NONE = 0, CENS_BLOQ = 1, CENS_ALOQ = 2 };

// status codes == PharmsolError variants (src/error/mod.rs:14-49); shared with pharmsol_cuda.h
enum : int {
    ST_OK = 0,
    ST_NON_FINITE_LIKELIHOOD = 1,
    ST_NEGATIVE_SIGMA = 2,
    ST_NON_FINITE_SIGMA = 3,
    ST_INVALID_OUTPUT_EQUATION = 4,
    ST_NONE_ERROR_MODEL = 5,
    ST_MISSING_ERROR_MODEL = 6,
    ST_SOLVER_FAILURE = 7,
    ST_INPUT_OUT_OF_RANGE = 8,
    ST_OUTEQ_OUT_OF_RANGE = 9,
    ST_UNKNOWN_INPUT_LABEL = 10,
    ST_UNKNOWN_OUTPUT_LABEL = 11,
    ST_IMAGINARY_ROOTS = 12,
    ST_UNSUPPORTED_INPUT_ROUTE_KIND = 13,
    ST_MISSING_COVARIATE = 14,
    ST_OTHER = 15,
};

// One event of a subject's timeline (48 B, 16-B aligned).
//   bolus:       a = amount
//   infusion:    a = amount, b = duration
//   observation: a = observed value (NaN if missing), b = c = -0.5*ln(2*pi) - ln(sigma),
//                w = 1/(2 sigma^2), sigma = sigma (for censored rows); these depend only on the
//                observation and the error model (data/error_model.rs:1045-1080), so the host
//                computes them once at flatten time.
struct __attribute__((aligned(16))) EventRec {
    double time;
    double a;
    double b;
    double w;
    double sigma;
    int32_t meta;     // bits 0-1 kind | 2-3 censoring | 4 has_value | 8-15 input/outeq index | 16-23 host status
    int32_t obs_row;  // row of this observation in the predictions output (global over the population), -1 otherwise
};
// Timeline program (closed-form models without lag: nothing about the event walk depends on the support point).  The host
// flattener runs the reference's interval logic ONCE per occasion — consecutive events, interior infusion boundaries
// with the 1e-12 de-duplication (analytical/mod.rs:311-327), the rates active on each sub-interval (:337-357) — and
// emits a flat list of records the device executes in order:
//   EV_BOLUS / EV_OBS   the event record unchanged (infusion events carry no action and are dropped)
//   EV_STEP             time = sub-interval end, a = dt (the same IEEE difference the device formed per pair),
//                       b / w / sigma = rate of route 0 / 1 / 2 on the sub-interval; with more than 3 routes obs_row
//                       indexes PopView::prog_rates (route_len doubles per step).  One route and one covariate
//                       (PopView::prog_cov): w / sigma = the covariate interpolated at the sub-interval end / at t = dt
//                       (the two derive-time conventions, COVTIME_*), so the step needs no segment scan
// so a pair costs one broadcast load + arithmetic per record: no cursor, no boundary scan, no infusion scan.
__host__ __device__ inline int ev_kind(int meta) { return meta & 3; }
__host__ __device__ inline int ev_cens(int meta) { return (meta >> 2) & 3; }
__host__ __device__ inline int ev_has_value(int meta) { return (meta >> 4) & 1; }
__host__ __device__ inline int ev_index(int meta) { return (meta >> 8) & 0xff; }
__host__ __device__ inline int ev_status(int meta) { return (meta >> 16) & 0xff; }
__host__ __device__ inline int ev_pack(int kind, int cens, int has_value, int index, int status) {
    return (kind & 3) | ((cens & 3) << 2) | ((has_value & 1) << 4) | ((index & 0xff) << 8) | ((status & 0xff) << 16);
}

// One covariate interpolation segment (data/covariate.rs:26-66): value = slope*t + intercept on
// [from, to); carry-forward segments have slope 0.  `to` = +inf for the unbounded last segment.
struct __attribute__((aligned(16))) CovSeg {
    double from, to, slope, intercept;
};

// One infusion of an occasion (warp-uniform), already label-resolved.
struct __attribute__((aligned(16))) InfRec {
    double time, duration, rate;      // rate = amount / duration, divided once on the host (the same IEEE quotient
                                      // the reference forms per use: analytical/mod.rs:356, sde/mod.rs:130)
    int32_t input;
    int32_t pad;
};

// Device view of a flattened population (pointers into one device allocation).
struct PopView {
    const int32_t* occ_offsets;   // [nsub+1]   occasions of subject i = [occ_offsets[i], occ_offsets[i+1])
    const int32_t* occ_index;     // [nocc]     occasion.index() (init runs only when 0)
    const int32_t* ev_offsets;    // [nocc+1]
    const EventRec* events;       // [nev]      sorted (time, Obs<Bolus<Infusion), stable
    const int32_t* bol_offsets;   // [nocc+1]   boluses of the occasion, in event order (lag merge stream)
    const int32_t* bol_event;     // [nbol]     index into events[]
    const int32_t* inf_offsets;   // [nocc+1]
    const InfRec* infs;           // [ninf]
    const int32_t* bnd_offsets;   // [nocc+1]   sorted unique start/end times of infusions with duration > 0
    const double* bnds;           // [nbnd]
    const int32_t* cov_offsets;   // [nocc*ncov+1]  segments of covariate c in occasion o = [cov_offsets[o*ncov+c], ...+1)
    const CovSeg* cov_segs;
    const double* occ_t0;         // [nocc]     occasion.initial_time() (ODE t0, data/structs.rs:782-793)
    const int32_t* prog_offsets;  // [nocc+1]   timeline program of the occasion (models without lag; nullptr otherwise)
    const EventRec* prog;         // [nprog]
    const double* prog_rates;     // [nstep * route_len]  only when route_len > 3
    int32_t nsub;
    int32_t ncov;
    int32_t max_events;           // max events in any occasion
    int32_t prog_cov;             // 1: models with ONE covariate and one route carry the interpolated covariate in the EV_STEP
                                  //    records (w = value at the sub-interval end, sigma = value at t = dt), see data.cpp
};

// 0-4 are this backend's own integrators; 5-6 carry the reference's remaining solver names (ode/mod.rs:59-84):
// Bdf = variable-order NDF/BDF 1-5 (diffsol `bdf`, the default of every reference ODE), Esdirk34 = diffsol `esdirk34`.
enum : int { SOLVER_DOPRI5 = 0, SOLVER_TSIT5 = 1, SOLVER_SDIRK4 = 2, SOLVER_TRBDF2 = 3, SOLVER_RODAS4 = 4, SOLVER_BDF = 5, SOLVER_ESDIRK34 = 6,
             SOLVER_COUNT = 7 };
enum : int { COVTIME_INTERVAL_END = 0, COVTIME_INTERVAL_LENGTH = 1 };
enum : int { SDE_MEAN_PREDICTION = 0, SDE_PARTICLE_FILTER = 1 };
enum : int { EM_REFERENCE_ADAPTIVE = 0, EM_FIXED_STEP = 1 };
enum : int { SDE_NORMALS_FP32 = 0, SDE_NORMALS_FP64 = 1 };

// ResidualErrorModel (data/residual_error.rs:69-139): prediction-based sigma, evaluated on the device.
enum : int { RESID_MISSING = 0, RESID_CONSTANT = 1, RESID_PROPORTIONAL = 2, RESID_COMBINED = 3, RESID_EXPONENTIAL = 4 };
struct ResidErr {
    int32_t kind;
    int32_t pad;
    double a, b;       // Constant{a} | Proportional{b} | Combined{a, b} | Exponential{sigma = a}
};
constexpr int PSI_MAX_RESID = 8;

struct RunOpts {
    double rtol, atol;        // ODE tolerances (reference default 1e-4 / 1e-4, ode/mod.rs:40-41)
    double h0;                // initial step (<= 0: automatic)
    double em_dt;             // fixed-step EM step size
    uint64_t seed;            // Philox key for SDE models
    int32_t solver;           // SOLVER_*
    int32_t cov_time;         // COVTIME_* (analytical derive time semantics)
    int32_t max_steps;        // per integrate-to-stop call
    int32_t nparticles;       // SDE
    int32_t sde_mode;         // SDE_*
    int32_t em_mode;          // EM_*
    int32_t want_pred;        // write per-observation predictions
    int32_t want_ll;          // accumulate the log-likelihood (0 for estimate_predictions: no error model needed)
    int32_t nsub_limit;       // > 0: only the first nsub_limit subjects (work probe)
    int32_t balance;          // ODE: order the columns by probed step counts so the lanes of a warp do similar work
    // log_likelihood_batch (likelihood/mod.rs:119-177): thread q evaluates subject q with parameter row q and
    // scores it with the prediction-based residual error models below
    int32_t diagonal;
    int32_t warp_tasks;       // few support points (< 128): warp w of the 1-D grid takes (subject, 32-column chunk) task w
    int32_t nresid;
    int32_t sde_normals;      // SDE noise: 0 = FP32 Box-Muller on 24-bit uniforms (default), 1 = FP64 Box-Muller on 32-bit uniforms
    ResidErr resid[PSI_MAX_RESID];
};

// Output buffers.
struct OutView {
    double* ll;                 // column-major (nsub x ncols): ll[i + j*ld_ll]   (matrix.rs:60 F-order)
    int64_t ld_ll;
    double* pred;               // (nobs_total x ncols) row-major: pred[row*ld_pred + j]  (coalesced over j)
    int64_t ld_pred;
    unsigned long long* first_error;   // atomicMin of (pair_index << 8 | code); init = ~0ull
    unsigned long long* counters;      // [0] accepted steps [1] rejected steps [2] rhs/kernel evals [3] Newton iters
    double* scratch;                   // SDE particle workspace: one slab of scratch_stride doubles per CTA
    int64_t scratch_stride;
    const int32_t* col_perm;           // optional: thread slot q evaluates column col_perm[q] (work-balanced warps)
    unsigned int* col_work;            // optional (probe launch): per-column solver attempts, atomically accumulated
    // Fused all-gather: when npeers > 0 every result is stored straight into the FULL column-major psi of every rank
    // (peer pointers mapped over NVLink / NVSwitch) at its global column, instead of into a local slab that a
    // separate collective copies afterwards.  ll is ignored then.
    double* ll_peers[8];
    int32_t npeers;
    int32_t scratch_in_smem;           // SDE: the particle workspace is the CTA's dynamic shared memory (it fits), not `scratch`
    unsigned long long* pair_ticket;   // SDE, optional: zeroed per launch; CTAs draw their next pair from it (dynamic balance) instead of striding
    int64_t col_base;                  // global index of local column 0 (column shards / pipelined chunks): error pair = i + (j + col_base) * nsub
};

}  // namespace psi
