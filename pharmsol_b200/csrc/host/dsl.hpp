// dsl.hpp — pharmsol-dsl front end (subset) and the CUDA-C code generator.
//
// Accepts the two surface forms of the reference's DSL and lowers them to one model description:
//   * line-oriented authoring shorthand     pharmsol-dsl/src/authoring.rs:362-900
//   * canonical `model name { ... }` blocks  pharmsol-dsl/src/parser.rs:300-1261
// then emits a CUDA-C "model policy" struct (device functions derive / dynamics / outputs / init /
// lag / fa / drift / diffusion / kparams / jacobian) that psi_engine.cuh is instantiated with.  The
// emitter plays the role of the reference's Rust-source AOT emitter (src/dsl/rust_backend.rs:29-490)
// with CUDA C as the target; the function roles are the frozen ABI's eight roles
// (src/dsl/compiled_backend_abi.rs:6-33).  Expression typing follows pharmsol-dsl/src/analyze.rs
// (Int/Real/Bool, `/` and `^` always real, integer-valued literals are Int) and the intrinsic set
// is pharmsol-dsl/src/analysis.rs:663-700.
#pragma once
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "data.hpp"

namespace pharmsol {
namespace dsl {

struct DslError : std::runtime_error {
    int pos;
    DslError(const std::string& m, int p = -1) : std::runtime_error(m), pos(p) {}
};

enum class ModelKind : int { Ode = 0, Analytical = 1, Sde = 2 };   // == EqnKind repr(C) (equation/mod.rs:580-586)

struct Expr;
using ExprP = std::shared_ptr<Expr>;
struct Expr {
    enum Kind { Num, BoolLit, Name, Index, Call, Unary, Binary, IfElse } kind = Num;
    double num = 0.0;
    bool bval = false;
    std::string name;            // identifier / operator / callee
    std::vector<ExprP> args;
    int pos = 0;
};

struct Stmt {
    enum Kind { Assign, Let, If, For } kind = Assign;
    // Assign / Let
    std::string callee;          // "" (plain name), "ddt", "noise", "init", "out"
    std::string target;
    ExprP index;                 // target[index]
    ExprP value;
    // If
    ExprP cond;
    std::vector<Stmt> then_body, else_body;
    // For (exclusive range)
    std::string var;
    ExprP lo, hi;
    std::vector<Stmt> body;
    int pos = 0;
};

struct StateDecl { std::string name; int len = 1; bool is_array = false; int offset = 0; };
struct CovDecl { std::string name; std::string interpolation; };
struct RouteDecl {
    std::string name;
    bool has_kind = false;
    RouteKind kind = RouteKind::Bolus;
    std::string dest;
    ExprP dest_index;
    ExprP lag, fa;
    int index = 0;               // dense input slot
    int dest_offset = 0;
    int declaration_index = 0;
};

struct ModelAst {
    std::string name;
    ModelKind kind = ModelKind::Ode;
    bool authoring = false;
    std::vector<std::string> params;
    std::vector<std::pair<std::string, ExprP>> constants;
    std::vector<CovDecl> covariates;
    std::vector<StateDecl> states;
    std::vector<RouteDecl> routes;
    std::vector<std::string> derived_decl, outputs_decl;
    std::vector<Stmt> derive, dynamics, outputs, init, diffusion;
    std::string structure;
    int particles = 0;
};

// Parse either surface form (auto-detected: canonical sources start with `model`).
ModelAst parse_model(const std::string& source);

// Everything the runtime needs to know about a compiled model (mirror of NativeModelInfo,
// src/dsl/model_info.rs:17-92) plus the generated CUDA-C.
struct CompiledModel {
    std::string name;
    ModelKind kind = ModelKind::Ode;
    std::vector<std::string> parameters, derived, covariates, states, outputs;
    std::vector<RouteInfo> routes;
    std::vector<std::string> covariate_interpolation;            // "" | "linear" | "locf", parallel to `covariates`
    std::vector<std::pair<std::string, int>> state_decls;       // declared states (arrays once) with their dense offsets
    int state_len = 0, derived_len = 0, output_len = 0, route_len = 0;
    int analytical_kernel = -1;
    int particles = 0;
    bool has_lag = false, has_fa = false, has_init = false, has_derive = false;
    std::string struct_body;     // the body of the policy struct (hashed for the module cache)
    std::string id;              // 16 hex digits: FNV-1a of struct_body
    ModelLabels labels() const;
    // Full translation unit.  `entries`: list of (solver, entry symbol) to instantiate.
    std::string cuda_source(const std::vector<std::pair<int, std::string>>& entries, bool aot_register) const;
    std::string model_info_json() const;
    // Host twin exporting the reference's frozen compiled-backend ABI (compiled_backend_abi.rs:6-33): C++ source of a cdylib.
    std::string host_source() const;
    std::vector<std::string> injection_lines;   // the `out[dest] += rate[k]` lines fused into the device dynamics (absent from the host twin)
};

CompiledModel compile_model(const ModelAst& ast);
// parse + analyse + emit; a failure is re-thrown rendered like the reference's Diagnostic::render (pharmsol-dsl/src/diagnostic.rs:220-266):
// `error[DSL1000|DSL2000]: message`, `  --> line L, column C`, then `  = note / help / suggestion` lines
CompiledModel compile_source(const std::string& src);

int analytical_kernel_index(const std::string& name);                        // -1 if unknown
const std::vector<std::string>& analytical_kernel_params(int kernel);

}  // namespace dsl
}  // namespace pharmsol
