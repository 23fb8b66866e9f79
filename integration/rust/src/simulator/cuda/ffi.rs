// UNVERIFIED GLUE SOURCE — written against include/pharmsol_cuda.h, NOT compiled: there is no Rust toolchain in the
// environment this backend was built in (no cargo / rustc / registry).  It shows where the C ABI plugs into pharmsol
// (src/simulator/cuda/, feature `cuda`); see INTEGRATION.md.  Everything below the FFI line (the shared library, the
// CUDA kernels, the DSL -> CUDA-C generator) is built and tested; the Python mirror pharmsol_b200/api.py drives the
// same entry points through ctypes.
// Cargo.toml:  [features] cuda = ["dsl-core"]          build.rs: println!("cargo:rustc-link-lib=dylib=pharmsol_cuda");
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_void};

#[repr(C)] pub struct pcu_ctx { _p: [u8; 0] }
#[repr(C)] pub struct pcu_model { _p: [u8; 0] }
#[repr(C)] pub struct pcu_subject_builder { _p: [u8; 0] }
#[repr(C)] pub struct pcu_subject { _p: [u8; 0] }
#[repr(C)] pub struct pcu_data { _p: [u8; 0] }
#[repr(C)] pub struct pcu_population { _p: [u8; 0] }

#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct pcu_error_model { pub kind: i32, pub pad: i32, pub factor: f64, pub c0: f64, pub c1: f64, pub c2: f64, pub c3: f64 }

extern "C" {
    pub fn pharmsol_cuda_abi_version() -> i32;
    pub fn pharmsol_cuda_device_count(n: *mut i32) -> i32;
    pub fn pharmsol_cuda_ctx_create(device: i32, out: *mut *mut pcu_ctx) -> i32;
    pub fn pharmsol_cuda_ctx_destroy(ctx: *mut pcu_ctx);
    pub fn pharmsol_cuda_last_error_message() -> *const c_char;

    pub fn pharmsol_subject_builder_new(id: *const c_char) -> *mut pcu_subject_builder;
    pub fn pharmsol_subject_builder_bolus(b: *mut pcu_subject_builder, t: f64, amount: f64, input: *const c_char);
    pub fn pharmsol_subject_builder_infusion(b: *mut pcu_subject_builder, t: f64, amount: f64, input: *const c_char, dur: f64);
    pub fn pharmsol_subject_builder_observation_with_error(b: *mut pcu_subject_builder, t: f64, v: f64, outeq: *const c_char,
                                                           c0: f64, c1: f64, c2: f64, c3: f64, censoring: i32);
    pub fn pharmsol_subject_builder_observation(b: *mut pcu_subject_builder, t: f64, v: f64, outeq: *const c_char);
    pub fn pharmsol_subject_builder_censored_observation(b: *mut pcu_subject_builder, t: f64, v: f64, outeq: *const c_char, censoring: i32);
    pub fn pharmsol_subject_builder_missing_observation(b: *mut pcu_subject_builder, t: f64, outeq: *const c_char);
    pub fn pharmsol_subject_builder_covariate(b: *mut pcu_subject_builder, name: *const c_char, t: f64, v: f64);
    pub fn pharmsol_subject_builder_reset(b: *mut pcu_subject_builder);
    pub fn pharmsol_subject_builder_build(b: *mut pcu_subject_builder) -> *mut pcu_subject;
    pub fn pharmsol_subject_set_covariate_fixed(s: *mut pcu_subject, occasion: i32, name: *const c_char, fixed: i32) -> i32;
    pub fn pharmsol_subject_free(s: *mut pcu_subject);
    pub fn pharmsol_data_new() -> *mut pcu_data;
    pub fn pharmsol_data_add_subject(d: *mut pcu_data, s: *const pcu_subject) -> i32;
    pub fn pharmsol_data_free(d: *mut pcu_data);

    pub fn pharmsol_cuda_model_from_dsl(ctx: *mut pcu_ctx, src: *const c_char, len: usize, out: *mut *mut pcu_model) -> i32;
    pub fn pharmsol_cuda_model_destroy(m: *mut pcu_model);
    pub fn pharmsol_cuda_model_set_solver(m: *mut pcu_model, solver: i32, rtol: f64, atol: f64) -> i32;
    pub fn pharmsol_cuda_model_set_particles(m: *mut pcu_model, n: u32, seed: u64, sde_mode: i32, em_mode: i32, em_dt: f64) -> i32;
    pub fn pharmsol_cuda_model_set_cov_time(m: *mut pcu_model, mode: i32) -> i32;
    pub fn pharmsol_cuda_model_compile(ctx: *mut pcu_ctx, m: *mut pcu_model, source_out: *mut i32) -> i32;
    pub fn pharmsol_cuda_model_export_artifact(m: *mut pcu_model, path: *const c_char, solvers: *const i32, nsolvers: i32) -> i32;
    pub fn pharmsol_cuda_model_load_artifact(ctx: *mut pcu_ctx, path: *const c_char, out: *mut *mut pcu_model) -> i32;
    pub fn pharmsol_cuda_artifact_info_json(path: *const c_char, buf: *mut c_char, cap: usize) -> i64;

    pub fn pharmsol_cuda_population_create(ctx: *mut pcu_ctx, m: *const pcu_model, d: *const pcu_data,
                                           ems: *const pcu_error_model, n: i32, out: *mut *mut pcu_population) -> i32;
    pub fn pharmsol_cuda_population_set_error_models(p: *mut pcu_population, ems: *const pcu_error_model, n: i32) -> i32;
    pub fn pharmsol_cuda_population_destroy(p: *mut pcu_population);
    pub fn pharmsol_cuda_population_nobservations(p: *const pcu_population) -> i64;

    pub fn pharmsol_cuda_log_likelihood_matrix(ctx: *mut pcu_ctx, m: *mut pcu_model, pop: *mut pcu_population,
        support_points: *const f64, nspp: i64, nparams: i32, out: *mut f64, first_error_code: *mut i32, first_error_pair: *mut i64) -> i32;
    pub fn pharmsol_cuda_predictions(ctx: *mut pcu_ctx, m: *mut pcu_model, pop: *mut pcu_population,
        support_points: *const f64, nspp: i64, nparams: i32, out: *mut f64) -> i32;
    pub fn pharmsol_cuda_log_likelihood_matrix_device(ctx: *mut pcu_ctx, m: *mut pcu_model, pop: *mut pcu_population,
        spp_soa_dev: *const f64, ncols: i64, ld_spp: i64, out_dev: *mut f64, ld_out: i64, first_col: i64, stream: *mut c_void) -> i32;
    pub fn pharmsol_cuda_collect_errors(ctx: *mut pcu_ctx, code: *mut i32, pair: *mut i64) -> i32;
    pub fn pharmsol_cuda_status_batch_begin(ctx: *mut pcu_ctx, stream: *mut c_void) -> i32;
    // fused all-gather: results stored straight into every rank's full psi through peer-mapped pointers
    pub fn pharmsol_cuda_log_likelihood_matrix_peers(ctx: *mut pcu_ctx, m: *mut pcu_model, pop: *mut pcu_population,
        spp_soa_dev: *const f64, ncols: i64, ld_spp: i64, out_full_peers: *const *mut f64, npeers: i32, ld_out: i64, first_col: i64,
        stream: *mut c_void) -> i32;
    // likelihood/mod.rs:119-177
    pub fn pharmsol_cuda_log_likelihood_batch(ctx: *mut pcu_ctx, m: *mut pcu_model, pop: *mut pcu_population, parameters: *const f64,
        nrows: i64, nparams: i32, models: *const pcu_residual_error_model, n_models: i32, out: *mut f64) -> i32;
    // data/parser/pmetrics
    pub fn pharmsol_data_read_pmetrics(path: *const c_char, out: *mut *mut pcu_data) -> i32;
}
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct pcu_residual_error_model { pub kind: i32, pub pad: i32, pub a: f64, pub b: f64 }   // 1 constant(a) 2 proportional(b) 3 combined(a,b) 4 exponential(a)
