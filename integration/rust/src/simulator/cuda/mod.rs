//! `pharmsol::simulator::cuda` — the Rust side of the B200 psi-matrix backend (cargo feature `cuda`).
//!
//! STATUS: complete source, NOT compiled — the authoring environment has no cargo / rustc / crates registry, so this file
//! has never been through the borrow checker.  It is written against `include/pharmsol_cuda.h` (bound in `ffi.rs`, which
//! is generated from that header and checked for completeness) and against the pharmsol v0.28.8 API it plugs into
//! (`src/data/structs.rs`, `src/data/event.rs`, `src/data/error_model.rs`, `src/simulator/likelihood/matrix.rs`).  Everything
//! below the FFI line — the shared library, the kernels, the DSL -> CUDA-C generator — is built and tested from C
//! (`examples/native_matrix.c`) and Python (`pharmsol_b200/api.py` drives the same entry points in the same order).
//!
//! What lives here:
//!   * `CudaEquation`                 a model compiled from pharmsol-dsl source (or loaded from a `.pkm`), one or MANY devices
//!   * `CudaPopulation`               the flattened `Data` + `AssayErrorModels` resident in HBM (built once, reused per call)
//!   * `log_likelihood_matrix_cuda`   `matrix.rs:52-106` with the same signature and the same F-order result
//!   * `estimate_predictions_cuda`, `log_likelihood_batch_cuda`
use std::collections::hash_map::DefaultHasher;
use std::ffi::{CStr, CString};
use std::hash::{Hash, Hasher};
use std::path::Path;
use std::sync::Mutex;

use ndarray::{Array2, ShapeBuilder};

use crate::data::error_model::{AssayErrorModel, AssayErrorModels};
use crate::data::event::{Censor, Event};
use crate::simulator::equation::ode::{ExplicitRkTableau, OdeSolver, SdirkTableau};
use crate::simulator::likelihood::{Prediction, SubjectPredictions};
use crate::{Data, PharmsolError, Subject};

pub mod ffi;

// ---------------------------------------------------------------------------------------------------------------------
// errors: the C status codes are PharmsolError variants (include/pharmsol_cuda.h PCU_ERR_*; src/error/mod.rs:14-49)
// ---------------------------------------------------------------------------------------------------------------------
fn last_error_message() -> String {
    unsafe {
        let p = ffi::pharmsol_cuda_last_error_message();
        if p.is_null() { String::new() } else { CStr::from_ptr(p).to_string_lossy().into_owned() }
    }
}

fn status_to_error(rc: i32, pair: i64) -> PharmsolError {
    use crate::data::error_model::ErrorModelError as E;
    let msg = last_error_message();
    match rc {
        1 => PharmsolError::NonFiniteLikelihood(f64::NAN),
        2 => PharmsolError::ErrorModelError(E::NegativeSigma),
        3 => PharmsolError::ErrorModelError(E::NonFiniteSigma),
        4 => PharmsolError::ErrorModelError(E::InvalidOutputEquation(0)),
        5 => PharmsolError::ErrorModelError(E::NoneErrorModel(0)),
        6 => PharmsolError::ErrorModelError(E::MissingErrorModel),
        7 => PharmsolError::DiffsolError(format!("solver failure at pair {pair}: {msg}")),
        // 8-11, 13: label / range errors are raised at flatten time with the offending label in the message
        12 => PharmsolError::OtherError(format!("imaginary roots in the analytical kernel (pair {pair})")),   // the reference panics here
        _ => PharmsolError::OtherError(msg),
    }
}

fn check(rc: i32) -> Result<(), PharmsolError> {
    if rc == 0 { Ok(()) } else { Err(status_to_error(rc, -1)) }
}

fn cstring(s: &str) -> CString {
    CString::new(s.replace('\0', " ")).expect("interior NUL removed")
}

// ---------------------------------------------------------------------------------------------------------------------
// handles (freed on drop; the library locks per context, so sharing `&CudaEquation` across rayon threads is sound)
// ---------------------------------------------------------------------------------------------------------------------
struct Ctx(*mut ffi::pcu_ctx);
impl Drop for Ctx { fn drop(&mut self) { unsafe { ffi::pharmsol_cuda_ctx_destroy(self.0) } } }
struct Model(*mut ffi::pcu_model);
impl Drop for Model { fn drop(&mut self) { unsafe { ffi::pharmsol_cuda_model_destroy(self.0) } } }

/// Flattened `Data` + error models on the device(s) of the creating context.
pub struct CudaPopulation { ptr: *mut ffi::pcu_population, nsub: usize, nobs: usize, key: u64 }
impl Drop for CudaPopulation { fn drop(&mut self) { unsafe { ffi::pharmsol_cuda_population_destroy(self.ptr) } } }
unsafe impl Send for CudaPopulation {}
unsafe impl Sync for CudaPopulation {}

/// `OdeSolver` (ode/mod.rs:59-84) -> PCU_SOLVER_*: every reference solver has its own device counterpart.
fn solver_code(s: &OdeSolver) -> i32 {
    match s {
        OdeSolver::Bdf => 5,                                        // PCU_SOLVER_BDF: variable-order NDF/BDF 1-5
        OdeSolver::Sdirk(SdirkTableau::TrBdf2) => 3,                // PCU_SOLVER_TRBDF2
        OdeSolver::Sdirk(SdirkTableau::Esdirk34) => 6,              // PCU_SOLVER_ESDIRK34
        OdeSolver::ExplicitRk(ExplicitRkTableau::Tsit45) => 1,      // PCU_SOLVER_TSIT45
    }
}

/// A model compiled for the device from pharmsol-dsl source.  The macro surface (`analytical!/ode!/sde!`) stores host
/// `fn` pointers that cannot run on a GPU, so device models are authored in DSL (or loaded from a `.pkm`).
pub struct CudaEquation {
    ctx: Ctx,
    model: Model,
    nparams: usize,
    nouteqs: usize,
    solver: OdeSolver,
    rtol: f64,
    atol: f64,
    // one resident population per (Data, AssayErrorModels) content hash: the reference rebuilds its per-subject caches
    // from `subject.hash()` (ode/mod.rs:293); here the whole flattened population is the cached object
    pops: Mutex<Vec<std::sync::Arc<CudaPopulation>>>,
}
unsafe impl Send for CudaEquation {}
unsafe impl Sync for CudaEquation {}   // Equation: 'static + Clone + Sync (equation/mod.rs:377); the C handles lock internally

impl CudaEquation {
    /// One device.
    pub fn from_dsl(source: &str, device: i32) -> Result<Self, PharmsolError> { Self::from_dsl_on(source, &[device]) }

    /// SURVEY §8b `ctx_create(device_ids, n_dev)`: one process drives every listed GPU; `log_likelihood_matrix_cuda`
    /// then splits the support-point columns over them and each device copies its slab straight into the result.
    pub fn from_dsl_on(source: &str, devices: &[i32]) -> Result<Self, PharmsolError> {
        let mut ctx = std::ptr::null_mut();
        check(unsafe { ffi::pharmsol_cuda_ctx_create_multi(devices.as_ptr(), devices.len() as i32, &mut ctx) })?;
        let ctx = Ctx(ctx);
        let mut model = std::ptr::null_mut();
        check(unsafe { ffi::pharmsol_cuda_model_from_dsl(ctx.0, source.as_ptr() as *const _, source.len(), &mut model) })?;
        Self::wrap(ctx, Model(model))
    }

    /// `load_aot_model` (dsl/aot.rs:316-353) for the CUDA target: the library checks API version, checksum and engine build.
    pub fn from_artifact(path: &Path, devices: &[i32]) -> Result<Self, PharmsolError> {
        let mut ctx = std::ptr::null_mut();
        check(unsafe { ffi::pharmsol_cuda_ctx_create_multi(devices.as_ptr(), devices.len() as i32, &mut ctx) })?;
        let ctx = Ctx(ctx);
        let mut model = std::ptr::null_mut();
        let p = cstring(&path.to_string_lossy());
        check(unsafe { ffi::pharmsol_cuda_model_load_artifact(ctx.0, p.as_ptr(), &mut model) })?;
        Self::wrap(ctx, Model(model))
    }

    fn wrap(ctx: Ctx, model: Model) -> Result<Self, PharmsolError> {
        let nparams = unsafe { ffi::pharmsol_cuda_model_nparams(model.0) } as usize;
        let nouteqs = unsafe { ffi::pharmsol_cuda_model_nouteqs(model.0) } as usize;
        let eq = Self { ctx, model, nparams, nouteqs, solver: OdeSolver::default(), rtol: 1e-4, atol: 1e-4, pops: Mutex::new(Vec::new()) };
        eq.apply_solver()?;                                        // the reference default: Bdf, rtol = atol = 1e-4 (ode/mod.rs:40-41)
        Ok(eq)
    }

    fn apply_solver(&self) -> Result<(), PharmsolError> {
        if unsafe { ffi::pharmsol_cuda_model_kind(self.model.0) } != 0 { return Ok(()); }      // PCU_KIND_ODE only
        check(unsafe { ffi::pharmsol_cuda_model_set_solver(self.model.0, solver_code(&self.solver), self.rtol, self.atol) })
    }
    pub fn with_solver(mut self, s: OdeSolver) -> Result<Self, PharmsolError> { self.solver = s; self.apply_solver()?; Ok(self) }
    pub fn with_tolerances(mut self, rtol: f64, atol: f64) -> Result<Self, PharmsolError> { self.rtol = rtol; self.atol = atol; self.apply_solver()?; Ok(self) }
    pub fn with_particles(self, n: u32, seed: u64, particle_filter: bool) -> Result<Self, PharmsolError> {
        check(unsafe { ffi::pharmsol_cuda_model_set_particles(self.model.0, n, seed, particle_filter as i32, 0 /* reference EM */, 0.0) })?;
        Ok(self)
    }

    /// `compile_module_source_to_aot` (dsl/aot.rs:146-300) for the CUDA target (NVRTC only; no GPU needed) ...
    pub fn export_artifact(&self, path: &Path, solvers: &[OdeSolver]) -> Result<(), PharmsolError> {
        let codes: Vec<i32> = solvers.iter().map(solver_code).collect();
        let p = cstring(&path.to_string_lossy());
        check(unsafe { ffi::pharmsol_cuda_model_export_artifact(self.model.0, p.as_ptr(), codes.as_ptr(), codes.len() as i32) })
    }
    /// ... and for the HOST target: a cdylib with the frozen `pharmsol_dsl_*` symbols that `load_aot_model` itself opens.
    pub fn export_native_artifact(&self, path: &Path) -> Result<(), PharmsolError> {
        let p = cstring(&path.to_string_lossy());
        check(unsafe { ffi::pharmsol_cuda_model_export_host_artifact(self.model.0, p.as_ptr()) })
    }

    // -----------------------------------------------------------------------------------------------------------------
    // src/data -> SoA device buffers, once per (Data, error models)
    // -----------------------------------------------------------------------------------------------------------------
    fn content_key(data: &Data, ems: Option<&AssayErrorModels>) -> u64 {
        let mut h = DefaultHasher::new();
        for s in data.iter() { s.hash().hash(&mut h); }             // Subject::hash (data/structs.rs:483)
        if let Some(e) = ems { e.hash().hash(&mut h); }            // AssayErrorModels::hash (error_model.rs:410)
        h.finish()
    }

    pub fn population(&self, data: &Data, ems: Option<&AssayErrorModels>) -> Result<std::sync::Arc<CudaPopulation>, PharmsolError> {
        let key = Self::content_key(data, ems);
        if let Some(p) = self.pops.lock().unwrap().iter().find(|p| p.key == key) { return Ok(p.clone()); }
        let d = unsafe { ffi::pharmsol_data_new() };
        struct DataGuard(*mut ffi::pcu_data);
        impl Drop for DataGuard { fn drop(&mut self) { unsafe { ffi::pharmsol_data_free(self.0) } } }
        let _guard = DataGuard(d);
        for subject in data.iter() {
            let id = cstring(subject.id());
            let b = unsafe { ffi::pharmsol_subject_builder_new(id.as_ptr()) };
            let mut fixed: Vec<(i32, CString)> = Vec::new();
            for (k, occasion) in subject.occasions().iter().enumerate() {
                if k > 0 { unsafe { ffi::pharmsol_subject_builder_reset(b) } }
                for (name, cov) in occasion.covariates().covariates() {                  // data/covariate.rs:184, 336
                    let cname = cstring(&name);
                    for (t, v) in cov.observations() { unsafe { ffi::pharmsol_subject_builder_covariate(b, cname.as_ptr(), t, v) } }
                    if cov.fixed() { fixed.push((k as i32, cname)); }
                }
                for event in occasion.events() {
                    match event {                                                          // data/event.rs:107-575
                        Event::Bolus(x) => {
                            let l = cstring(x.input().as_str());
                            unsafe { ffi::pharmsol_subject_builder_bolus(b, x.time(), x.amount(), l.as_ptr()) }
                        }
                        Event::Infusion(x) => {
                            let l = cstring(x.input().as_str());
                            unsafe { ffi::pharmsol_subject_builder_infusion(b, x.time(), x.amount(), l.as_ptr(), x.duration()) }
                        }
                        Event::Observation(o) => {
                            let l = cstring(o.outeq().as_str());
                            let cens = match o.censoring() { Censor::None => 0, Censor::BLOQ => 1, Censor::ALOQ => 2 };
                            match (o.value(), o.errorpoly()) {
                                (None, _) => unsafe { ffi::pharmsol_subject_builder_missing_observation(b, o.time(), l.as_ptr()) },
                                (Some(v), Some(p)) => unsafe {
                                    ffi::pharmsol_subject_builder_observation_with_error(b, o.time(), v, l.as_ptr(), p.c0(), p.c1(), p.c2(), p.c3(), cens)
                                },
                                (Some(v), None) => unsafe { ffi::pharmsol_subject_builder_censored_observation(b, o.time(), v, l.as_ptr(), cens) },
                            }
                        }
                    }
                }
            }
            let s = unsafe { ffi::pharmsol_subject_builder_build(b) };                     // consumes the builder
            if s.is_null() { return Err(PharmsolError::OtherError(last_error_message())); }
            for (occ, name) in &fixed { check(unsafe { ffi::pharmsol_subject_set_covariate_fixed(s, *occ, name.as_ptr(), 1) })?; }
            let rc = unsafe { ffi::pharmsol_data_add_subject(d, s) };
            unsafe { ffi::pharmsol_subject_free(s) };
            check(rc)?;
        }
        // AssayErrorModels by output-equation slot (error_model.rs:352, 786-812)
        let dense: Vec<ffi::pcu_error_model> = match ems {
            None => Vec::new(),
            Some(ems) => (0..self.nouteqs).map(|k| match ems.error_model(k) {
                Ok(AssayErrorModel::Additive { lambda, poly }) => ffi::pcu_error_model { kind: 1, pad: 0, factor: lambda.value(), c0: poly.c0(), c1: poly.c1(), c2: poly.c2(), c3: poly.c3() },
                Ok(AssayErrorModel::Proportional { gamma, poly }) => ffi::pcu_error_model { kind: 2, pad: 0, factor: gamma.value(), c0: poly.c0(), c1: poly.c1(), c2: poly.c2(), c3: poly.c3() },
                _ => ffi::pcu_error_model::default(),                                       // AssayErrorModel::None / no model for this outeq
            }).collect(),
        };
        let mut pop = std::ptr::null_mut();
        check(unsafe { ffi::pharmsol_cuda_population_create(self.ctx.0, self.model.0, d, dense.as_ptr(), dense.len() as i32, &mut pop) })?;
        let p = std::sync::Arc::new(CudaPopulation {
            ptr: pop,
            nsub: unsafe { ffi::pharmsol_cuda_population_nsubjects(pop) } as usize,
            nobs: unsafe { ffi::pharmsol_cuda_population_nobservations(pop) } as usize,
            key,
        });
        let mut cache = self.pops.lock().unwrap();
        if cache.len() >= 8 { cache.remove(0); }
        cache.push(p.clone());
        Ok(p)
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// the hot path
// ---------------------------------------------------------------------------------------------------------------------
/// `log_likelihood_matrix` (likelihood/matrix.rs:52-106): same arguments, same F-order `(nsub, nspp)` result, the first
/// failing pair aborts with its error (matrix.rs:96-104).  With a multi-device `CudaEquation` the columns are sharded
/// inside the library.  The population is flattened and uploaded on the first call for a (Data, error models) pair and
/// reused afterwards; the support points go up and psi comes back on every call.
pub fn log_likelihood_matrix_cuda(eq: &CudaEquation, subjects: &Data, support_points: &Array2<f64>,
                                  error_models: &AssayErrorModels, _progress: bool) -> Result<Array2<f64>, PharmsolError> {
    let pop = eq.population(subjects, Some(error_models))?;
    let spp = support_points.as_standard_layout();                    // rows = support points, cols = params in model order
    let (nspp, np) = spp.dim();
    let mut out: Array2<f64> = Array2::zeros((pop.nsub, nspp).f());   // F-order, matrix.rs:60
    let (mut code, mut pair) = (0i32, -1i64);
    let rc = unsafe {
        ffi::pharmsol_cuda_log_likelihood_matrix(eq.ctx.0, eq.model.0, pop.ptr, spp.as_ptr(), nspp as i64, np as i32, out.as_mut_ptr(), &mut code, &mut pair)
    };
    if rc != 0 { return Err(status_to_error(rc, pair)); }
    Ok(out)
}

/// `psi` (matrix.rs:138-150): exp of the above, exponentiated on the device.
pub fn psi_cuda(eq: &CudaEquation, subjects: &Data, support_points: &Array2<f64>, error_models: &AssayErrorModels) -> Result<Array2<f64>, PharmsolError> {
    let pop = eq.population(subjects, Some(error_models))?;
    let spp = support_points.as_standard_layout();
    let (nspp, np) = spp.dim();
    let mut out: Array2<f64> = Array2::zeros((pop.nsub, nspp).f());
    let (mut code, mut pair) = (0i32, -1i64);
    let rc = unsafe { ffi::pharmsol_cuda_psi(eq.ctx.0, eq.model.0, pop.ptr, spp.as_ptr(), nspp as i64, np as i32, out.as_mut_ptr(), &mut code, &mut pair) };
    if rc != 0 { return Err(status_to_error(rc, pair)); }
    Ok(out)
}

/// `Equation::estimate_predictions` (equation/mod.rs:526-532) for one subject and one parameter vector.
pub fn estimate_predictions_cuda(eq: &CudaEquation, subject: &Subject, params: &[f64]) -> Result<SubjectPredictions, PharmsolError> {
    let data = Data::new(vec![subject.clone()]);
    let pop = eq.population(&data, None)?;
    let mut pred = vec![0.0f64; pop.nobs];
    check(unsafe { ffi::pharmsol_cuda_predictions(eq.ctx.0, eq.model.0, pop.ptr, params.as_ptr(), 1, params.len() as i32, pred.as_mut_ptr()) })?;
    // rows follow the observations in event order, occasion by occasion (missing observations included)
    let mut out: Vec<Prediction> = Vec::with_capacity(pop.nobs);
    let mut row = 0usize;
    for occasion in subject.occasions() {
        for event in occasion.events() {
            if let Event::Observation(o) = event {
                out.push(o.to_prediction(pred[row], Vec::new()));          // data/event.rs:698 (state not returned by the device path)
                row += 1;
            }
        }
    }
    Ok(SubjectPredictions::from(out))
}

/// `Equation::estimate_log_likelihood_dense` (equation/mod.rs:468-477): a 1 x 1 matrix call on the library's latency path
/// (pinned staging, one kernel, one synchronize); the population of a subject is cached by content hash like the matrix's.
pub fn estimate_log_likelihood_cuda(eq: &CudaEquation, subject: &Subject, params: &[f64], ems: &AssayErrorModels) -> Result<f64, PharmsolError> {
    let data = Data::new(vec![subject.clone()]);
    let spp = Array2::from_shape_vec((1, params.len()), params.to_vec()).map_err(PharmsolError::NdarrayShapeError)?;
    Ok(log_likelihood_matrix_cuda(eq, &data, &spp, ems, false)?[[0, 0]])
}

/// `log_likelihood_batch` (likelihood/mod.rs:119-177): row i of `parameters` belongs to subject i; prediction-based sigma.
pub fn log_likelihood_batch_cuda(eq: &CudaEquation, subjects: &Data, parameters: &Array2<f64>,
                                 models: &[ffi::pcu_residual_error_model]) -> Result<Vec<f64>, PharmsolError> {
    let pop = eq.population(subjects, None)?;
    let prm = parameters.as_standard_layout();
    let (nrows, np) = prm.dim();
    let mut out = vec![0.0f64; pop.nsub];
    check(unsafe {
        ffi::pharmsol_cuda_log_likelihood_batch(eq.ctx.0, eq.model.0, pop.ptr, prm.as_ptr(), nrows as i64, np as i32, models.as_ptr(), models.len() as i32, out.as_mut_ptr())
    })?;
    Ok(out)
}

/// The whole psi resident on EVERY device of a multi-device equation (gathered over NVLink by the copy engines):
/// returns one device pointer per GPU, owned by the library until the next replicated call.
pub fn log_likelihood_matrix_replicated_cuda(eq: &CudaEquation, subjects: &Data, support_points: &Array2<f64>,
                                             error_models: &AssayErrorModels) -> Result<Vec<*mut f64>, PharmsolError> {
    let pop = eq.population(subjects, Some(error_models))?;
    let spp = support_points.as_standard_layout();
    let (nspp, np) = spp.dim();
    let n = unsafe { ffi::pharmsol_cuda_ctx_num_devices(eq.ctx.0) } as usize;
    let mut ptrs: Vec<*mut f64> = vec![std::ptr::null_mut(); n];
    let (mut code, mut pair) = (0i32, -1i64);
    let rc = unsafe {
        ffi::pharmsol_cuda_log_likelihood_matrix_replicated(eq.ctx.0, eq.model.0, pop.ptr, spp.as_ptr(), nspp as i64, np as i32, 0 /* PCU_GATHER_COPY_ENGINE */,
                                                            ptrs.as_mut_ptr(), &mut code, &mut pair)
    };
    if rc != 0 { return Err(status_to_error(rc, pair)); }
    Ok(ptrs)
}
