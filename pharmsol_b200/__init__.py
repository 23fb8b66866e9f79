"""pharmsol_b200 — B200-native (sm_100a) psi-matrix backend for LAPKB/pharmsol.

The package holds only what the hot path needs: ``csrc/`` (hand-written CUDA kernels, the DSL ->
CUDA-C code generator and the C ABI of ``include/pharmsol_cuda.h``), the ctypes binding and a
host-side mirror of pharmsol's ``Subject`` / ``Data`` / ``Equation`` / ``log_likelihood_matrix``
interface.  There is no CPU implementation here: importing works without a GPU (parsing, code
generation and NVRTC compilation do too), every numeric call needs an sm_100 device.
"""
from ._lib import PharmsolError, LIB_PATH, device_count  # noqa: F401
from .api import (  # noqa: F401
    Analytical, AssayErrorModel, AssayErrorModels, Censor, CovTime, Data, EmMode, EqnKind, Equation, ErrorPoly, ODE, OdeSolver,
    ParameterOrder, Prediction, ResidentPsi, ResidualErrorModel, ResidualErrorModels, SDE, SdeMode, Subject, SubjectBuilder, SubjectPredictions, RuntimeArtifactFormat, RuntimeBackend, RuntimeCompilationTarget, analytical,
    NativeArtifact, compile_module_source_to_native_aot, compile_module_source_to_aot, compile_module_source_to_runtime, load_aot_model, load_runtime_artifact, read_aot_model_info, log_likelihood_batch, log_likelihood_matrix, log_psi, ode, psi, read_pmetrics, sde,
)

__all__ = [n for n in dir() if not n.startswith("_")]
