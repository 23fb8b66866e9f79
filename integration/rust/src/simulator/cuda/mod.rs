// UNVERIFIED GLUE SOURCE — written against include/pharmsol_cuda.h, NOT compiled: there is no Rust toolchain in the
// environment this backend was built in (no cargo / rustc / registry).  It shows where the C ABI plugs into pharmsol
// (src/simulator/cuda/, feature `cuda`); see INTEGRATION.md.  Everything below the FFI line (the shared library, the
// CUDA kernels, the DSL -> CUDA-C generator) is built and tested; the Python mirror pharmsol_b200/api.py drives the
// same entry points through ctypes.
pub mod ffi;

/// A model compiled for the device from pharmsol-dsl source (the macro surface stores host `fn` pointers that
/// cannot run on a GPU — SURVEY F9 — so `analytical!/ode!/sde!` gain a `dsl_source()` twin or are authored in DSL).
pub struct CudaEquation { ctx: CtxHandle, model: ModelHandle, info: NativeModelInfo, kind: EqnKind }
unsafe impl Sync for CudaEquation {}   // handles are internally locked (one mutex per context), Equation: Sync (equation/mod.rs:377)
unsafe impl Send for CudaEquation {}

impl CudaEquation {
    pub fn from_dsl(source: &str, device: i32) -> Result<Self, PharmsolError> { /* ctx_create + model_from_dsl (+ info JSON) */ }
    /// dsl/aot.rs:316-353 `load_aot_model` for the CUDA target: API version + checksum are checked by the library.
    pub fn from_artifact(path: &std::path::Path, device: i32) -> Result<Self, PharmsolError> { /* ctx_create + pharmsol_cuda_model_load_artifact */ }
    /// dsl/aot.rs:146-300 `compile_module_source_to_aot`: writes the `.pkm` (NVRTC only, no GPU needed).
    pub fn export_artifact(&self, path: &std::path::Path, solvers: &[OdeSolver]) -> Result<(), PharmsolError> { /* pharmsol_cuda_model_export_artifact */ }
    pub fn with_solver(self, s: OdeSolver) -> Self      { /* model_set_solver; Bdf -> PCU_SOLVER_RODAS4, Tsit45 -> 1, TrBdf2 -> 3, Esdirk34 -> 2 */ self }
    pub fn with_tolerances(self, rtol: f64, atol: f64) -> Self { /* model_set_solver */ self }

    /// src/data -> SoA device buffers: replay every Occasion/Event through the builder ABI once per (Data, error models)
    /// and keep the `pcu_population` (content hash of Data + AssayErrorModels as the cache key, like `subject.hash()`
    /// in ode/mod.rs:293).
    fn population(&self, data: &Data, ems: &AssayErrorModels) -> Result<PopHandle, PharmsolError> {
        let d = unsafe { ffi::pharmsol_data_new() };
        for subject in data.subjects_slice() {
            let b = unsafe { ffi::pharmsol_subject_builder_new(cstr(subject.id())) };
            for (k, occasion) in subject.occasions().iter().enumerate() {
                if k > 0 { unsafe { ffi::pharmsol_subject_builder_reset(b) } }
                for (name, cov) in occasion.covariates().covariates() {          // data/covariate.rs:189-212
                    for (t, v) in cov.observations() { unsafe { ffi::pharmsol_subject_builder_covariate(b, cstr(name), *t, *v) } }
                }
                for event in occasion.events() {
                    match event {                                                  // data/event.rs:107-575
                        Event::Bolus(x)    => unsafe { ffi::pharmsol_subject_builder_bolus(b, x.time(), x.amount(), cstr(x.input().as_str())) },
                        Event::Infusion(x) => unsafe { ffi::pharmsol_subject_builder_infusion(b, x.time(), x.amount(), cstr(x.input().as_str()), x.duration()) },
                        Event::Observation(o) => { /* value None -> missing_observation; errorpoly Some -> observation_with_error; censoring -> 0/1/2 */ }
                    }
                }
            }
            let s = unsafe { ffi::pharmsol_subject_builder_build(b) };
            /* fixed covariates: pharmsol_subject_set_covariate_fixed(s, occasion, name, 1) */
            check(unsafe { ffi::pharmsol_data_add_subject(d, s) })?;
            unsafe { ffi::pharmsol_subject_free(s) };
        }
        let dense: Vec<ffi::pcu_error_model> = ems.bind_to(&self.info.outputs)?   // error_model.rs bind_to: label -> outeq slot
            .iter().map(|m| m.into()).collect();                                    // Additive{lambda,poly} / Proportional{gamma,poly} / None
        let mut pop = std::ptr::null_mut();
        check(unsafe { ffi::pharmsol_cuda_population_create(self.ctx.0, self.model.0, d, dense.as_ptr(), dense.len() as i32, &mut pop) })?;
        unsafe { ffi::pharmsol_data_free(d) };
        Ok(PopHandle(pop))
    }
}

/// matrix.rs:52-106 with the same signature and the same F-order result.
pub fn log_likelihood_matrix_cuda(eq: &CudaEquation, subjects: &Data, support_points: &Array2<f64>,
                                  error_models: &AssayErrorModels, _progress: bool) -> Result<Array2<f64>, PharmsolError> {
    let pop = eq.population(subjects, error_models)?;
    let spp = support_points.as_standard_layout();                    // rows = support points, cols = params in model order
    let (nspp, np) = spp.dim();
    let mut out: Array2<f64> = Array2::zeros((subjects.len(), nspp).f());   // F-order, matrix.rs:60
    let (mut code, mut pair) = (0i32, -1i64);
    let rc = unsafe { ffi::pharmsol_cuda_log_likelihood_matrix(eq.ctx.0, eq.model.0, pop.0, spp.as_ptr(), nspp as i64, np as i32,
                                                               out.as_mut_ptr(), &mut code, &mut pair) };
    if rc != 0 { return Err(PharmsolError::from_cuda_status(rc, pair, last_error_message())); }   // first error wins, matrix.rs:96-104
    Ok(out)
}

impl EquationTypes for CudaEquation { type S = V; type P = SubjectPredictions; }
impl Equation for CudaEquation {
    fn kind() -> EqnKind { /* per instance in practice: EqnKind is #[repr(C)] == PCU_KIND_* (equation/mod.rs:580-586) */ }
    fn estimate_log_likelihood_dense(&self, subject: &Subject, p: &[f64], ems: &AssayErrorModels) -> Result<f64, PharmsolError> {
        let data = Data::new(vec![subject.clone()]);
        Ok(log_likelihood_matrix_cuda(self, &data, &Array2::from_shape_vec((1, p.len()), p.to_vec()).unwrap(), ems, false)?[[0, 0]])
    }
    fn estimate_predictions_dense(&self, subject: &Subject, p: &[f64]) -> Result<SubjectPredictions, PharmsolError> {
        /* pharmsol_cuda_predictions -> Vec<f64>, zipped with subject observations into Prediction{time, obs, pred, outeq, ...} */
    }
    /* nstates / nouteqs from model info; simulate_subject = predictions + optional likelihood */
}
