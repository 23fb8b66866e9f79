// dsl_emit.cpp — semantic analysis and the CUDA-C emitter (the `simulator::cuda` codegen target).
//
// Mirrors src/dsl/rust_backend.rs:29-490 (one function per ModelFunctionKind, same 7-argument
// shape as the frozen CompiledModelFunction ABI, src/dsl/native.rs:45-53) with these additions
// for the device:
//   * dependency flags (does `derive` read t / covariates / states / rates?) so the engine can
//     hoist time-invariant work out of the event loop;
//   * infusion-rate injection into ddt(destination) unless the model reads rate(route) itself
//     (src/dsl/model_info.rs:154, native.rs:814-825);
//   * a forward-mode analytic Jacobian d(dynamics)/d(state) (through `derive`) for the stiff
//     solvers — the reference has none (Newton uses a linear-RHS trick, ode/closure.rs:360-375).
// Typing: Int / Real / Bool; integer-valued literals are Int; `/`, `^`, pow are always Real;
// max/min -> fmax/fmin; Real->Int casts saturate (analyze.rs:2751-2817, rust_backend.rs:277-466).
#include <cstring>
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <functional>
#include <set>
#include <sstream>

#include "dsl.hpp"

namespace pharmsol {
namespace dsl {

static const char* kKernelNames[12] = {
    "one_compartment", "one_compartment_cl", "one_compartment_cl_with_absorption", "one_compartment_with_absorption",
    "two_compartments", "two_compartments_cl", "two_compartments_cl_with_absorption", "two_compartments_with_absorption",
    "three_compartments", "three_compartments_cl", "three_compartments_cl_with_absorption", "three_compartments_with_absorption"};

int analytical_kernel_index(const std::string& name) {
    for (int i = 0; i < 12; ++i) if (name == kKernelNames[i]) return i;
    return -1;
}
const std::vector<std::string>& analytical_kernel_params(int k) {
    // pharmsol-dsl/src/analysis.rs:240-255
    static const std::vector<std::string> P[12] = {
        {"ke"}, {"cl", "v"}, {"ka", "cl", "v"}, {"ka", "ke"},
        {"ke", "kcp", "kpc"}, {"cl", "q", "vc", "vp"}, {"ka", "cl", "q", "vc", "vp"}, {"ke", "ka", "kcp", "kpc"},
        {"k10", "k12", "k13", "k21", "k31"}, {"cl", "q2", "q3", "vc", "v2", "v3"},
        {"ka", "cl", "q2", "q3", "vc", "v2", "v3"}, {"ka", "k10", "k12", "k13", "k21", "k31"}};
    return P[k];
}
static int kernel_state_count(int k) {
    static const int N[12] = {1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4};
    return N[k];
}

namespace {

enum class Ty { Int, Real, Bool };
struct Val {
    std::string code;
    Ty ty = Ty::Real;
    bool is_const = false;
    double cval = 0.0;
};

std::string fmt_real(double v) {
    if (std::isnan(v)) return "psi::psi_nan()";
    if (std::isinf(v)) return v > 0 ? "psi::psi_inf()" : "(-psi::psi_inf())";
    char buf[64];
    std::snprintf(buf, sizeof buf, "%.17g", v);
    std::string s = buf;
    if (s.find_first_of(".eE") == std::string::npos) s += ".0";
    return s;
}
std::string fmt_int(long long v) { return std::to_string(v) + "LL"; }

enum DepBits { DEP_T = 1, DEP_COV = 2, DEP_STATE = 4, DEP_RATE = 8, DEP_DERIVED = 16 };

enum class Role { Derive, Dynamics, Outputs, Init, Lag, Fa, Diffusion };

// "did you mean" (analyze.rs:1990-2027): the closest declared name by edit distance, case-insensitive ties first
static std::string nearest_name(const std::string& want, const std::vector<std::string>& have) {
    auto lower = [](std::string v) { for (auto& ch : v) ch = (char)std::tolower((unsigned char)ch); return v; };
    std::string best;
    size_t best_d = std::max<size_t>(2, want.size() / 3) + 1;
    for (const auto& h : have) {
        if (h == want) continue;
        if (lower(h) == lower(want)) return h;
        std::vector<size_t> prev(h.size() + 1), cur(h.size() + 1);
        for (size_t j = 0; j <= h.size(); ++j) prev[j] = j;
        for (size_t i = 1; i <= want.size(); ++i) {
            cur[0] = i;
            for (size_t j = 1; j <= h.size(); ++j)
                cur[j] = std::min({prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (want[i - 1] == h[j - 1] ? 0 : 1)});
            std::swap(prev, cur);
        }
        if (prev[h.size()] < best_d) { best_d = prev[h.size()]; best = h; }
    }
    return best;
}
static std::string suggest(const std::string& want, const std::vector<std::string>& have) {
    const std::string n = nearest_name(want, have);
    return n.empty() ? std::string() : "\n  suggestion: did you mean `" + n + "`?";
}

struct Ctx {
    const ModelAst& m;
    std::map<std::string, int> param_ix, cov_ix, derived_ix, output_ix, const_ix;
    std::map<std::string, StateDecl> state_by_name;
    std::vector<Val> const_vals;
    std::vector<std::string> derived_names;
    int state_len = 0;
    explicit Ctx(const ModelAst& mm) : m(mm) {}
};

struct Scope {
    std::map<std::string, Ty> locals;             // emitted as L_<name>
    std::map<std::string, long long> loop_vars;   // unrolled for-loop bindings
    int deps = 0;
    std::set<int> routes_read;
    Role role = Role::Derive;
    // forward-mode AD: names of the derivative variables in scope (empty => derivative is zero)
    bool ad = false;
    int ad_wrt = -1;                              // state offset the derivative is taken against
    std::set<std::string> ad_locals;              // locals that have a DL_<name> variable
    std::set<int> ad_derived;                     // derived slots that have a DD_<i> variable
    int if_depth = 0;                             // inside a conditional every assignment keeps its AD variable
};

struct Emitter {
    Ctx& c;
    explicit Emitter(Ctx& cc) : c(cc) {}

    // ---- pair-invariant hoisting ---------------------------------------------------------------------
    // A sub-expression that reads only parameters and constants has one value per (subject, support
    // point) pair.  If it contains a division, a power or a function call it is evaluated ONCE per pair in
    // `prologue` into an extra slot p[NP + k] instead of in every right-hand-side / output evaluation, and a
    // division by such an expression becomes a multiplication by its hoisted reciprocal (FP64 division is
    // ~20 instructions on the pipe that bounds these kernels).  The hoisted value is bit-identical to the
    // in-place one; only the reciprocal rewrite changes results, by at most 1 ulp per division.
    std::vector<std::string> slots;               // code of slot k (may reference earlier slots)
    static constexpr int kMaxSlots = 12;
    int np() const { return (int)c.param_ix.size(); }
    bool pair_invariant(const ExprP& e, const Scope& s) const {
        switch (e->kind) {
            case Expr::Num: case Expr::BoolLit: return true;
            case Expr::Name: {
                const std::string& n = e->name;
                if (s.loop_vars.count(n)) return true;
                if (s.locals.count(n)) return false;
                if (n == "t" || n == "time") return false;
                return c.param_ix.count(n) || c.const_ix.count(n);
            }
            case Expr::Index: return false;
            case Expr::Call: if (e->name == "rate") return false; [[fallthrough]];
            default:
                for (const auto& a : e->args) if (!pair_invariant(a, s)) return false;
                return true;
        }
    }
    static bool expensive(const ExprP& e) {
        if (e->kind == Expr::Call) return true;
        if (e->kind == Expr::Binary && (e->name == "/" || e->name == "^")) return true;
        for (const auto& a : e->args) if (expensive(a)) return true;
        return false;
    }
    // slot holding `code` (deduplicated); "" when the slot budget is exhausted
    std::string slot_for(const std::string& code) {
        for (size_t k = 0; k < slots.size(); ++k) if (slots[k] == code) return "p[" + std::to_string(np() + (int)k) + "]";
        if ((int)slots.size() >= kMaxSlots) return "";
        slots.push_back(code);
        return "p[" + std::to_string(np() + (int)slots.size() - 1) + "]";
    }
    // code of 1 / y when y is pair-invariant (constant folded, or a hoisted reciprocal); "" otherwise
    std::string reciprocal(const ExprP& ey, const Val& y, const Scope& s) {
        if (y.is_const) return (y.cval != 0.0) ? fmt_real(1.0 / y.cval) : "";
        if (!pair_invariant(ey, s)) return "";
        return slot_for("(1.0 / " + y.code + ")");
    }

    static Val real(const Val& v) {
        if (v.ty == Ty::Real) return v;
        Val r; r.ty = Ty::Real; r.is_const = v.is_const; r.cval = v.cval;
        if (v.is_const) r.code = fmt_real(v.cval);
        else if (v.ty == Ty::Bool) r.code = "((" + v.code + ") ? 1.0 : 0.0)";
        else r.code = "((double)(" + v.code + "))";
        return r;
    }
    static Val boolean(const Val& v) {
        if (v.ty == Ty::Bool) return v;
        Val r; r.ty = Ty::Bool;
        r.code = "((" + v.code + ") != " + (v.ty == Ty::Int ? "0LL" : "0.0") + ")";
        return r;
    }
    static Val mkconst(double x, Ty ty) {
        Val v; v.ty = ty; v.is_const = true; v.cval = x;
        v.code = ty == Ty::Int ? fmt_int((long long)x) : (ty == Ty::Bool ? (x != 0 ? "true" : "false") : fmt_real(x));
        return v;
    }
    int route_slot(const std::string& label, int pos) const {
        for (const auto& r : c.m.routes) if (r.name == label) return r.index;
        if (!label.empty() && std::all_of(label.begin(), label.end(), [](char ch) { return std::isdigit((unsigned char)ch) != 0; }))
            throw DslError("bare numeric route labels are not allowed in the DSL; use `input_" + label + "` instead", pos);
        std::vector<std::string> have;
        for (const auto& r : c.m.routes) have.push_back(r.name);
        throw DslError("unknown route `" + label + "` in rate(...)" + suggest(label, have), pos);
    }
    long long const_int(const ExprP& e, Scope& s) {
        Val v = expr(e, s);
        if (!v.is_const) throw DslError("expected a compile-time integer expression", e->pos);
        return (long long)v.cval;
    }

    Val name(const ExprP& e, Scope& s) {
        const std::string& n = e->name;
        // resolution order (analyze.rs:1272-1408): local, t/time, parameter, constant, covariate, state, derived
        auto lv = s.loop_vars.find(n);
        if (lv != s.loop_vars.end()) return mkconst((double)lv->second, Ty::Int);
        auto lo = s.locals.find(n);
        if (lo != s.locals.end()) { Val v; v.ty = lo->second; v.code = "L_" + n; return v; }
        if (n == "t" || n == "time") { s.deps |= DEP_T; Val v; v.code = "t"; return v; }
        auto pi = c.param_ix.find(n);
        if (pi != c.param_ix.end()) { Val v; v.code = "p[" + std::to_string(pi->second) + "]"; return v; }
        auto ci = c.const_ix.find(n);
        if (ci != c.const_ix.end()) return c.const_vals[(size_t)ci->second];
        auto vi = c.cov_ix.find(n);
        if (vi != c.cov_ix.end()) { s.deps |= DEP_COV; Val v; v.code = "cov[" + std::to_string(vi->second) + "]"; return v; }
        auto si = c.state_by_name.find(n);
        if (si != c.state_by_name.end()) {
            if (si->second.is_array) throw DslError("state array `" + n + "` must be indexed", e->pos);
            s.deps |= DEP_STATE;
            Val v; v.code = "x[" + std::to_string(si->second.offset) + "]"; return v;
        }
        auto di = c.derived_ix.find(n);
        if (di != c.derived_ix.end()) { s.deps |= DEP_DERIVED; Val v; v.code = "d[" + std::to_string(di->second) + "]"; return v; }
        for (const auto& r : c.m.routes)
            if (r.name == n) throw DslError("unknown identifier `" + n + "`\n  help: route inputs are read through `rate(" + n + ")`", e->pos);
        if (c.output_ix.count(n))
            throw DslError("unknown identifier `" + n + "`\n  help: outputs are assignment targets inside the `outputs` block and are not available as expression values", e->pos);
        std::vector<std::string> have;
        for (const auto& kv : s.locals) have.push_back(kv.first);
        for (const auto& kv : c.param_ix) have.push_back(kv.first);
        for (const auto& kv : c.const_ix) have.push_back(kv.first);
        for (const auto& kv : c.cov_ix) have.push_back(kv.first);
        for (const auto& kv : c.state_by_name) have.push_back(kv.first);
        for (const auto& kv : c.derived_ix) have.push_back(kv.first);
        throw DslError("unknown identifier `" + n + "`" + suggest(n, have), e->pos);
    }

    Val expr(const ExprP& e, Scope& s) {
        // In the right-hand side of an ODE — evaluated 6-7 times per step on the FP64 pipe that bounds those kernels — a
        // cheap pair-invariant sum / product (`ke + k12`) is hoisted as well: one slot instead of one operation per evaluation.
        const bool cheap_ok = e->kind == Expr::Binary && c.m.kind == ModelKind::Ode && s.role == Role::Dynamics &&
                              (e->name == "+" || e->name == "-" || e->name == "*");
        if ((e->kind == Expr::Binary || e->kind == Expr::Call || e->kind == Expr::Unary || e->kind == Expr::IfElse) && (expensive(e) || cheap_ok) && pair_invariant(e, s)) {
            Val v = expr_raw(e, s);
            if (v.is_const || v.ty != Ty::Real) return v;
            const std::string sl = slot_for(v.code);
            if (sl.empty()) return v;
            Val r; r.code = sl; return r;
        }
        return expr_raw(e, s);
    }
    Val expr_raw(const ExprP& e, Scope& s) {
        switch (e->kind) {
            case Expr::Num: {
                // a literal with zero fractional part is an Int constant even if written 70.0 (analyze.rs:2751-2763)
                const bool is_int = std::floor(e->num) == e->num && std::fabs(e->num) < 9e15;
                return mkconst(e->num, is_int ? Ty::Int : Ty::Real);
            }
            case Expr::BoolLit: return mkconst(e->bval ? 1 : 0, Ty::Bool);
            case Expr::Name: return name(e, s);
            case Expr::Index: {
                auto si = c.state_by_name.find(e->name);
                if (si == c.state_by_name.end()) throw DslError("`" + e->name + "` is not a state array", e->pos);
                const long long ix = const_int(e->args[0], s);
                if (ix < 0 || ix >= si->second.len) throw DslError("index out of range for `" + e->name + "`", e->pos);
                s.deps |= DEP_STATE;
                Val v; v.code = "x[" + std::to_string(si->second.offset + ix) + "]"; return v;
            }
            case Expr::Unary: {
                Val a = expr(e->args[0], s);
                if (e->name == "!") { Val b = boolean(a); Val r; r.ty = Ty::Bool; r.code = "(!" + b.code + ")";
                    if (a.is_const) return mkconst(a.cval != 0 ? 0 : 1, Ty::Bool); return r; }
                if (a.ty == Ty::Bool) a = real(a);
                if (e->name == "+") return a;
                if (a.is_const) return mkconst(-a.cval, a.ty);
                Val r; r.ty = a.ty; r.code = "(-" + a.code + ")"; return r;
            }
            case Expr::Binary: return binary(e, s);
            case Expr::Call: return call(e, s);
            case Expr::IfElse: {
                Val cnd = boolean(expr(e->args[0], s));
                Val a = expr(e->args[1], s), b = expr(e->args[2], s);
                if (a.ty == Ty::Bool && b.ty == Ty::Bool) { Val r; r.ty = Ty::Bool; r.code = "(" + cnd.code + " ? " + a.code + " : " + b.code + ")"; return r; }
                if (a.ty == Ty::Int && b.ty == Ty::Int) { Val r; r.ty = Ty::Int; r.code = "(" + cnd.code + " ? " + a.code + " : " + b.code + ")"; return r; }
                a = real(a); b = real(b);
                Val r; r.code = "(" + cnd.code + " ? " + a.code + " : " + b.code + ")"; return r;
            }
        }
        throw DslError("bad expression", e->pos);
    }

    // pow(x, c) with a literal exponent that has a cheap exact-arithmetic form (device/psi_common.cuh pow_*)
    static std::string special_pow(const Val& x, const Val& y) {
        if (!y.is_const || x.is_const) return "";
        const double c = y.cval;
        const char* fn = c == 0.5 ? "pow_half" : c == 0.25 ? "pow_quarter" : c == 0.75 ? "pow_three_quarters" : c == 1.5 ? "pow_three_halves"
                       : c == 2.0 ? "pow_2" : c == 3.0 ? "pow_3" : c == 4.0 ? "pow_4" : nullptr;
        if (c == 1.0) return x.code;
        return fn ? std::string("psi::") + fn + "(" + x.code + ")" : "";
    }

    Val binary(const ExprP& e, Scope& s) {
        const std::string& op = e->name;
        Val a = expr(e->args[0], s), b = expr(e->args[1], s);
        if (op == "&&" || op == "||") {
            Val x = boolean(a), y = boolean(b);
            Val r; r.ty = Ty::Bool; r.code = "(" + x.code + " " + op + " " + y.code + ")"; return r;
        }
        if (op == "==" || op == "!=" || op == "<" || op == "<=" || op == ">" || op == ">=") {
            Val r; r.ty = Ty::Bool;
            if (a.ty == Ty::Bool && b.ty == Ty::Bool) { r.code = "(" + a.code + " " + op + " " + b.code + ")"; return r; }
            if (a.ty == Ty::Int && b.ty == Ty::Int) { r.code = "(" + a.code + " " + op + " " + b.code + ")"; return r; }
            Val x = real(a), y = real(b);
            r.code = "(" + x.code + " " + op + " " + y.code + ")"; return r;
        }
        if (a.ty == Ty::Bool) a = real(a);
        if (b.ty == Ty::Bool) b = real(b);
        if (op == "+" || op == "-" || op == "*") {
            if (a.ty == Ty::Int && b.ty == Ty::Int) {
                if (a.is_const && b.is_const) {
                    const long long x = (long long)a.cval, y = (long long)b.cval;
                    return mkconst((double)(op == "+" ? x + y : op == "-" ? x - y : x * y), Ty::Int);
                }
                Val r; r.ty = Ty::Int; r.code = "(" + a.code + " " + op + " " + b.code + ")"; return r;
            }
            Val x = real(a), y = real(b);
            if (x.is_const && y.is_const) return mkconst(op == "+" ? x.cval + y.cval : op == "-" ? x.cval - y.cval : x.cval * y.cval, Ty::Real);
            if (op == "*") {
                // (-a) * b is written -(a * b) — the same value bit for bit — so that the compiler shares a * b with the
                // other statements that use it (`dx(depot) = -ka * depot`, `dx(central) = ka * depot - ...`)
                auto negated = [](const ExprP& q, const Val& v) {
                    return q->kind == Expr::Unary && q->name == "-" && !v.is_const && v.code.size() > 3 && v.code.compare(0, 2, "(-") == 0 && v.code.back() == ')';
                };
                const bool nx = negated(e->args[0], x), ny = negated(e->args[1], y);
                if (nx != ny) {
                    const std::string xi = nx ? x.code.substr(2, x.code.size() - 3) : x.code, yi = ny ? y.code.substr(2, y.code.size() - 3) : y.code;
                    Val r; r.code = "(-(" + xi + " * " + yi + "))"; return r;
                }
            }
            Val r; r.code = "(" + x.code + " " + op + " " + y.code + ")"; return r;
        }
        if (op == "/") {
            Val x = real(a), y = real(b);
            if (x.is_const && y.is_const) return mkconst(x.cval / y.cval, Ty::Real);
            const std::string rc = reciprocal(e->args[1], y, s);
            // ODE / SDE dynamics and the Jacobian pass: ~1-ulp quotient without the IEEE fix-up path (psi::fdiv); everything a
            // prediction or likelihood is computed from directly (derive, outputs, init, lag, fa) keeps the exact division
            const bool fast = s.ad || s.role == Role::Dynamics || s.role == Role::Diffusion;
            Val r; r.code = !rc.empty() ? "(" + x.code + " * " + rc + ")" : fast ? "psi::fdiv(" + x.code + ", " + y.code + ")" : "(" + x.code + " / " + y.code + ")"; return r;
        }
        if (op == "^") {
            Val x = real(a), y = real(b);
            if (x.is_const && y.is_const) return mkconst(std::pow(x.cval, y.cval), Ty::Real);
            const std::string sp = special_pow(x, y);
            Val r; r.code = sp.empty() ? "pow(" + x.code + ", " + y.code + ")" : sp; return r;
        }
        throw DslError("unknown operator `" + op + "`", e->pos);
    }

    Val call(const ExprP& e, Scope& s) {
        const std::string& f = e->name;
        if (f == "rate") {
            if (e->args.size() != 1 || (e->args[0]->kind != Expr::Name && e->args[0]->kind != Expr::Num)) throw DslError("rate(route) expects a route name", e->pos);
            std::string label = e->args[0]->kind == Expr::Name ? e->args[0]->name : std::to_string((long long)e->args[0]->num);
            const int slot = route_slot(label, e->pos);
            s.deps |= DEP_RATE;
            s.routes_read.insert(slot);
            Val v; v.code = "rate[" + std::to_string(slot) + "]"; return v;
        }
        std::vector<Val> a;
        for (const auto& x : e->args) a.push_back(expr(x, s));
        auto need = [&](size_t n) { if (a.size() != n) throw DslError("`" + f + "` expects " + std::to_string(n) + " argument(s)", e->pos); };
        auto unary_real = [&](const char* cfun) { need(1); Val x = real(a[0]); Val r; r.code = std::string(cfun) + "(" + x.code + ")"; return r; };
        if (f == "abs") {
            need(1);
            if (a[0].ty == Ty::Int) { Val r; r.ty = Ty::Int; r.code = "llabs(" + a[0].code + ")"; return r; }
            return unary_real("fabs");
        }
        if (f == "ceil") return unary_real("ceil");
        if (f == "floor") return unary_real("floor");
        if (f == "round") return unary_real("round");        // half away from zero == Rust f64::round
        if (f == "exp") return unary_real("exp");
        if (f == "ln" || f == "log") return unary_real("log");
        if (f == "log10") return unary_real("log10");
        if (f == "log2") return unary_real("log2");
        if (f == "sin") return unary_real("sin");
        if (f == "cos") return unary_real("cos");
        if (f == "tan") return unary_real("tan");
        if (f == "sqrt") return unary_real("sqrt");
        if (f == "pow") {
            need(2);
            Val x = real(a[0]), y = real(a[1]);
            if (x.is_const && y.is_const) return mkconst(std::pow(x.cval, y.cval), Ty::Real);
            const std::string sp = special_pow(x, y);
            Val r; r.code = sp.empty() ? "pow(" + x.code + ", " + y.code + ")" : sp; return r;
        }
        if (f == "max" || f == "min") {
            need(2);
            if (a[0].ty == Ty::Int && a[1].ty == Ty::Int) {
                Val r; r.ty = Ty::Int;
                r.code = "((" + a[0].code + (f == "max" ? ") > (" : ") < (") + a[1].code + ") ? (" + a[0].code + ") : (" + a[1].code + "))";
                return r;
            }
            Val x = real(a[0]), y = real(a[1]);
            Val r; r.code = (f == "max" ? "fmax(" : "fmin(") + x.code + ", " + y.code + ")"; return r;
        }
        throw DslError("unknown function `" + f + "`" + suggest(f, {"abs", "ceil", "exp", "floor", "ln", "log", "log10", "log2", "max", "min", "pow", "round", "sin", "cos", "tan", "sqrt", "rate"}), e->pos);
    }

    // ---- forward-mode derivative of a Real-valued expression w.r.t. state offset s.ad_wrt --------
    // returns "" when the derivative is identically zero
    std::string deriv(const ExprP& e, Scope& s) {
        switch (e->kind) {
            case Expr::Num: case Expr::BoolLit: return "";
            case Expr::Name: {
                const std::string& n = e->name;
                if (s.loop_vars.count(n)) return "";
                if (s.locals.count(n)) return s.ad_locals.count(n) ? "DL_" + n : "";
                if (n == "t" || n == "time" || c.param_ix.count(n) || c.const_ix.count(n) || c.cov_ix.count(n)) return "";
                auto si = c.state_by_name.find(n);
                if (si != c.state_by_name.end()) return si->second.offset == s.ad_wrt ? "1.0" : "";
                auto di = c.derived_ix.find(n);
                if (di != c.derived_ix.end()) return s.ad_derived.count(di->second) ? "DD_" + std::to_string(di->second) : "";
                return "";
            }
            case Expr::Index: {
                auto si = c.state_by_name.find(e->name);
                const long long ix = const_int(e->args[0], s);
                return (si->second.offset + ix == s.ad_wrt) ? "1.0" : "";
            }
            case Expr::Unary: {
                if (e->name == "!") return "";
                std::string d = deriv(e->args[0], s);
                if (d.empty()) return "";
                return e->name == "-" ? "(-" + d + ")" : d;
            }
            case Expr::IfElse: {
                std::string da = deriv(e->args[1], s), db = deriv(e->args[2], s);
                if (da.empty() && db.empty()) return "";
                Val cnd = boolean(expr(e->args[0], s));
                return "(" + cnd.code + " ? " + (da.empty() ? "0.0" : da) + " : " + (db.empty() ? "0.0" : db) + ")";
            }
            case Expr::Binary: {
                const std::string& op = e->name;
                if (op == "&&" || op == "||" || op == "==" || op == "!=" || op == "<" || op == "<=" || op == ">" || op == ">=") return "";
                std::string da = deriv(e->args[0], s), db = deriv(e->args[1], s);
                if (da.empty() && db.empty()) return "";
                Val a = real(expr(e->args[0], s)), b = real(expr(e->args[1], s));
                if (op == "+") return da.empty() ? db : db.empty() ? da : "(" + da + " + " + db + ")";
                if (op == "-") return da.empty() ? "(-" + db + ")" : db.empty() ? da : "(" + da + " - " + db + ")";
                if (op == "*") {
                    std::string t1 = da.empty() ? "" : "(" + da + " * " + b.code + ")";
                    std::string t2 = db.empty() ? "" : "(" + a.code + " * " + db + ")";
                    return t1.empty() ? t2 : t2.empty() ? t1 : "(" + t1 + " + " + t2 + ")";
                }
                if (op == "/") {
                    // (a/b)' = (a' - (a/b) b') / b: one reciprocal of b, shared with the quotient itself, instead of 1/b and 1/b^2
                    const std::string rc = reciprocal(e->args[1], b, s);
                    if (db.empty()) return rc.empty() ? "psi::fdiv(" + da + ", " + b.code + ")" : "(" + da + " * " + rc + ")";
                    const std::string q = "psi::fdiv(" + a.code + ", " + b.code + ")";
                    const std::string num = da.empty() ? "(-(" + q + " * " + db + "))" : "(" + da + " - (" + q + " * " + db + "))";
                    return "psi::fdiv(" + num + ", " + b.code + ")";
                }
                if (op == "^") return deriv_pow(a, b, da, db);
                return "";
            }
            case Expr::Call: {
                const std::string& f = e->name;
                if (f == "rate" || f == "ceil" || f == "floor" || f == "round") return "";
                std::vector<std::string> d;
                bool any = false;
                for (const auto& x : e->args) { d.push_back(deriv(x, s)); any = any || !d.back().empty(); }
                if (!any) return "";
                std::vector<Val> a;
                for (const auto& x : e->args) a.push_back(real(expr(x, s)));
                const std::string& u = a[0].code; const std::string& du = d[0];
                if (f == "exp") return "(exp(" + u + ") * " + du + ")";
                if (f == "ln" || f == "log") return "(" + du + " / " + u + ")";
                if (f == "log10") return "(" + du + " / (" + u + " * 2.302585092994046))";
                if (f == "log2") return "(" + du + " / (" + u + " * 0.6931471805599453))";
                if (f == "sqrt") return "(" + du + " / (2.0 * sqrt(" + u + ")))";
                if (f == "sin") return "(cos(" + u + ") * " + du + ")";
                if (f == "cos") return "(-sin(" + u + ") * " + du + ")";
                if (f == "tan") return "(" + du + " / (cos(" + u + ") * cos(" + u + ")))";
                if (f == "abs") return "((" + u + " < 0.0 ? -1.0 : 1.0) * " + du + ")";
                if (f == "pow") return deriv_pow(a[0], a[1], d[0], d[1]);
                if (f == "max" || f == "min") {
                    const std::string cmp = f == "max" ? " >= " : " <= ";
                    return "((" + a[0].code + cmp + a[1].code + ") ? " + (d[0].empty() ? "0.0" : d[0]) + " : " + (d[1].empty() ? "0.0" : d[1]) + ")";
                }
                return "";
            }
        }
        return "";
    }
    static std::string deriv_pow(const Val& a, const Val& b, const std::string& da, const std::string& db) {
        std::string t1, t2;
        if (!da.empty()) t1 = "(" + b.code + " * pow(" + a.code + ", " + b.code + " - 1.0) * " + da + ")";
        if (!db.empty()) t2 = "(pow(" + a.code + ", " + b.code + ") * log(" + a.code + ") * " + db + ")";
        return t1.empty() ? t2 : t2.empty() ? t1 : "(" + t1 + " + " + t2 + ")";
    }
};

// Collect the names assigned as plain targets anywhere in a statement list (derive targets).
void collect_targets(const std::vector<Stmt>& ss, std::vector<std::string>& out) {
    for (const auto& s : ss) {
        if (s.kind == Stmt::Assign && s.callee.empty()) {
            bool seen = false;
            for (auto& o : out) if (o == s.target) seen = true;
            if (!seen) out.push_back(s.target);
        } else if (s.kind == Stmt::If) { collect_targets(s.then_body, out); collect_targets(s.else_body, out); }
        else if (s.kind == Stmt::For) collect_targets(s.body, out);
    }
}
// output names assigned anywhere in an outputs block (`out(x) = ...` or plain `x = ...`)
void collect_output_targets(const std::vector<Stmt>& ss, std::vector<std::string>& out) {
    for (const auto& s : ss) {
        if (s.kind == Stmt::Assign && (s.callee == "out" || s.callee.empty())) out.push_back(s.target);
        else if (s.kind == Stmt::If) { collect_output_targets(s.then_body, out); collect_output_targets(s.else_body, out); }
        else if (s.kind == Stmt::For) collect_output_targets(s.body, out);
    }
}
// names of the states touched by `ddt(state) = ...` statements anywhere in a block
void collect_state_targets(const std::vector<Stmt>& ss, std::vector<std::string>& out) {
    for (const auto& s : ss) {
        if (s.kind == Stmt::Assign && s.callee == "ddt") out.push_back(s.target);
        else if (s.kind == Stmt::If) { collect_state_targets(s.then_body, out); collect_state_targets(s.else_body, out); }
        else if (s.kind == Stmt::For) collect_state_targets(s.body, out);
    }
}
void collect_lets(const std::vector<Stmt>& ss, std::vector<std::string>& out) {
    for (const auto& s : ss) {
        if (s.kind == Stmt::Let) { bool seen = false; for (auto& o : out) if (o == s.target) seen = true; if (!seen) out.push_back(s.target); }
        else if (s.kind == Stmt::If) { collect_lets(s.then_body, out); collect_lets(s.else_body, out); }
        else if (s.kind == Stmt::For) collect_lets(s.body, out);
    }
}

struct BodyEmitter {
    Emitter& em;
    Ctx& c;
    std::ostringstream os;
    int indent = 2;
    BodyEmitter(Emitter& e, Ctx& cc) : em(e), c(cc) {}
    void line(const std::string& s) { for (int i = 0; i < indent; ++i) os << "    "; os << s << "\n"; }

    // resolve the storage an assignment writes, as C lvalue
    std::string lvalue(const Stmt& st, Scope& s, int* slot_out = nullptr, char* space_out = nullptr) {
        auto state_off = [&](const std::string& n, const ExprP& idx) -> int {
            auto si = c.state_by_name.find(n);
            if (si == c.state_by_name.end()) throw DslError("unknown state `" + n + "`", st.pos);
            long long ix = 0;
            if (idx) ix = em.const_int(idx, s);
            else if (si->second.is_array) throw DslError("state array `" + n + "` must be indexed", st.pos);
            if (ix < 0 || ix >= si->second.len) throw DslError("index out of range for `" + n + "`", st.pos);
            return si->second.offset + (int)ix;
        };
        int slot = -1; char space = '?';
        std::string lv;
        if (st.callee == "ddt" || st.callee == "dx") { slot = state_off(st.target, st.index); space = 'o'; lv = "out[" + std::to_string(slot) + "]"; }
        else if (st.callee == "noise") { slot = state_off(st.target, st.index); space = 'o'; lv = "out[" + std::to_string(slot) + "]"; }
        else if (st.callee == "init") { slot = state_off(st.target, st.index); space = 'o'; lv = "out[" + std::to_string(slot) + "]"; }
        else if (st.callee == "out") {
            auto oi = c.output_ix.find(st.target);
            if (oi == c.output_ix.end()) throw DslError("unknown output `" + st.target + "`", st.pos);
            slot = oi->second; space = 'o'; lv = "out[" + std::to_string(slot) + "]";
        } else {
            // plain name: local if known, else role default
            if (s.locals.count(st.target)) { space = 'l'; lv = "L_" + st.target; }
            else if (s.role == Role::Derive) {
                auto di = c.derived_ix.find(st.target);
                if (di == c.derived_ix.end()) throw DslError("`" + st.target + "` is not a derived value", st.pos);
                slot = di->second; space = 'd'; lv = "d[" + std::to_string(slot) + "]";
            } else if (s.role == Role::Init) { slot = state_off(st.target, st.index); space = 'o'; lv = "out[" + std::to_string(slot) + "]"; }
            else if (s.role == Role::Outputs) {
                auto oi = c.output_ix.find(st.target);
                if (oi == c.output_ix.end()) throw DslError("unknown output `" + st.target + "`", st.pos);
                slot = oi->second; space = 'o'; lv = "out[" + std::to_string(slot) + "]";
            } else throw DslError("cannot assign `" + st.target + "` here", st.pos);
        }
        if (slot_out) *slot_out = slot;
        if (space_out) *space_out = space;
        return lv;
    }

    void stmts(const std::vector<Stmt>& ss, Scope& s) {
        for (const auto& st : ss) {
            switch (st.kind) {
                case Stmt::Let: {
                    Val v = em.expr(st.value, s);
                    Val r = v.ty == Ty::Bool ? v : Emitter::real(v);
                    // locals are pre-declared at function scope (zero-initialised, rust_backend.rs:90-120)
                    s.locals[st.target] = r.ty == Ty::Bool ? Ty::Bool : Ty::Real;
                    if (s.ad) {
                        std::string d = r.ty == Ty::Bool ? "" : em.deriv(st.value, s);
                        if (d.empty() && s.if_depth == 0) s.ad_locals.erase(st.target);
                        else { line("DL_" + st.target + " = " + (d.empty() ? "0.0" : d) + ";"); s.ad_locals.insert(st.target); }
                    } else {
                        line("L_" + st.target + " = " + r.code + ";");
                    }
                    break;
                }
                case Stmt::Assign: {
                    int slot; char space;
                    const std::string lv = lvalue(st, s, &slot, &space);
                    Val v = Emitter::real(em.expr(st.value, s));
                    if (s.ad) {
                        std::string d = em.deriv(st.value, s);
                        const bool drop = d.empty() && s.if_depth == 0;   // identically zero: no AD variable needed
                        if (space == 'l') { if (drop) s.ad_locals.erase(st.target); else { line("DL_" + st.target + " = " + (d.empty() ? "0.0" : d) + ";"); s.ad_locals.insert(st.target); } }
                        else if (space == 'd') { if (drop) s.ad_derived.erase(slot); else { line("DD_" + std::to_string(slot) + " = " + (d.empty() ? "0.0" : d) + ";"); s.ad_derived.insert(slot); } }
                        else line("J[" + std::to_string(slot) + " * NSTATE + " + std::to_string(s.ad_wrt) + "] = " + (d.empty() ? "0.0" : d) + ";");
                    } else {
                        line(lv + " = " + v.code + ";");
                    }
                    break;
                }
                case Stmt::If: {
                    Val cnd = Emitter::boolean(em.expr(st.cond, s));
                    line("if (" + cnd.code + ") {");
                    ++s.if_depth;
                    ++indent; stmts(st.then_body, s); --indent;
                    if (!st.else_body.empty()) {
                        line("} else {");
                        ++indent; stmts(st.else_body, s); --indent;
                    }
                    --s.if_depth;
                    line("}");
                    break;
                }
                case Stmt::For: {
                    const long long lo = em.const_int(st.lo, s), hi = em.const_int(st.hi, s);
                    for (long long k = lo; k < hi; ++k) {
                        s.loop_vars[st.var] = k;
                        stmts(st.body, s);
                    }
                    s.loop_vars.erase(st.var);
                    break;
                }
            }
        }
    }
};

// ---- definite assignment (analyze.rs:790-900, 1352-1360, 2334-2411) --------------------------------------------------
// A derived value may only be read where it has been assigned on every control-flow path: an `if` without `else`
// and a `for` body (which may run zero times) contribute nothing; `if`/`else` contributes the intersection.
struct Flow {
    std::set<std::string> avail;        // derived values readable here
    std::set<std::string> targets;      // derive / output targets assigned on every path so far
    std::set<std::string> locals;
};
static std::set<std::string> intersect(const std::set<std::string>& a, const std::set<std::string>& b) {
    std::set<std::string> out;
    for (const auto& x : a) if (b.count(x)) out.insert(x);
    return out;
}
static void flow_reads(const ExprP& e, const Ctx& c, const Flow& f) {
    if (!e) return;
    if (e->kind == Expr::Name) {
        if (!f.locals.count(e->name) && c.derived_ix.count(e->name) && !f.avail.count(e->name))
            throw DslError("derived value `" + e->name + "` is not definitely assigned at this point", e->pos);
        return;
    }
    if (e->kind == Expr::Call && e->name == "rate") return;       // the argument is a route label, not a value
    for (const auto& a : e->args) flow_reads(a, c, f);
}
static void flow_stmts(const std::vector<Stmt>& ss, const Ctx& c, Role role, Flow& f) {
    for (const auto& st : ss) {
        switch (st.kind) {
            case Stmt::Let:
                flow_reads(st.value, c, f);
                f.locals.insert(st.target);
                break;
            case Stmt::Assign: {
                flow_reads(st.index, c, f);
                flow_reads(st.value, c, f);
                const bool plain = st.callee.empty() && !f.locals.count(st.target);
                if (role == Role::Derive && !st.callee.empty()) throw DslError("derive assignments must target a bare identifier", st.pos);
                if (role == Role::Derive && plain) { f.avail.insert(st.target); f.targets.insert(st.target); }
                else if (role == Role::Outputs && (st.callee == "out" || plain)) f.targets.insert(st.target);
                break;
            }
            case Stmt::If: {
                flow_reads(st.cond, c, f);
                Flow a = f, b = f;
                flow_stmts(st.then_body, c, role, a);
                if (!st.else_body.empty()) {
                    flow_stmts(st.else_body, c, role, b);
                    if (role == Role::Derive) f.avail = intersect(a.avail, b.avail);
                    if (role == Role::Derive || role == Role::Outputs) f.targets = intersect(a.targets, b.targets);
                }
                break;
            }
            case Stmt::For: {
                flow_reads(st.lo, c, f);
                flow_reads(st.hi, c, f);
                Flow body = f;
                body.locals.insert(st.var);
                flow_stmts(st.body, c, role, body);
                break;
            }
        }
    }
}

std::string fnv1a_hex(const std::string& s) {
    unsigned long long h = 1469598103934665603ull;
    for (unsigned char ch : s) { h ^= ch; h *= 1099511628211ull; }
    char buf[32];
    std::snprintf(buf, sizeof buf, "%016llx", h);
    return buf;
}

const char* kSig = "(double t, const double* x, const double* p, const double* cov, const double* rate, const double* d, double* out)";

}  // namespace

CompiledModel compile_model(const ModelAst& ast_in) {
    ModelAst ast = ast_in;
    CompiledModel cm;
    cm.name = ast.name;
    cm.kind = ast.kind;
    cm.particles = ast.particles;
    Ctx c(ast);

    // ---- symbol tables -------------------------------------------------------------------------
    // Global names (analyze.rs:19-46, 1739-1800): reserved words are refused, every declared name is unique across
    // parameters / constants / covariates / states / derived / routes / outputs, except that a route and an
    // output may share a label.
    static const char* kReserved[] = {"abs", "bioavailability", "carry_forward", "ceil", "ddt", "exp", "floor", "lag", "linear", "ln", "locf", "log",
                                      "log10", "log2", "max", "min", "noise", "pow", "rate", "round", "sin", "cos", "tan", "sqrt", "t", "time"};
    std::map<std::string, char> all_names;      // 'p' 'k' 'v' 's' 'd' 'r' 'o'
    auto declare = [&](const std::string& n, char kind) {
        for (const char* r : kReserved) if (n == r) throw DslError("`" + n + "` is reserved by the DSL and cannot be used as a symbol name\n  suggestion: rename `" + n + "` to `" + n + "_value`");
        auto it = all_names.find(n);
        if (it == all_names.end()) { all_names[n] = kind; return; }
        const bool overlap_ok = (it->second == 'r' && kind == 'o') || (it->second == 'o' && kind == 'r') || (it->second == 'r' && kind == 'r');
        if (!overlap_ok) throw DslError("symbol name `" + n + "` collides with existing `" + n + "`");
    };
    auto check_dup = [&](const std::string& n, const char* what, char kind) {
        if (c.param_ix.count(n) || c.cov_ix.count(n) || c.state_by_name.count(n) || c.const_ix.count(n))
            throw DslError(std::string("duplicate ") + what + " `" + n + "`\n  note: " + what + " `" + n + "` first declared here");
        declare(n, kind);
    };
    auto digits = [](const std::string& t) { return !t.empty() && std::all_of(t.begin(), t.end(), [](char ch) { return std::isdigit((unsigned char)ch) != 0; }); };
    auto suffix_of = [&](const std::string& label, const char* prefix) -> std::string {      // canonical_numeric_suffix, analyze.rs:2448-2451
        const size_t n = std::strlen(prefix);
        return label.compare(0, n, prefix) == 0 && digits(label.substr(n)) ? label.substr(n) : std::string();
    };
    for (const auto& p : ast.params) { check_dup(p, "parameter", 'p'); c.param_ix[p] = (int)cm.parameters.size(); cm.parameters.push_back(p); }
    for (const auto& v : ast.covariates) { check_dup(v.name, "covariate", 'v'); c.cov_ix[v.name] = (int)cm.covariates.size(); cm.covariates.push_back(v.name); cm.covariate_interpolation.push_back(v.interpolation); }
    int off = 0;
    for (auto& st : ast.states) {
        check_dup(st.name, "state", 's');
        st.offset = off;
        c.state_by_name[st.name] = st;
        cm.state_decls.emplace_back(st.name, st.offset);
        if (st.is_array) for (int i = 0; i < st.len; ++i) cm.states.push_back(st.name + "[" + std::to_string(i) + "]");
        else cm.states.push_back(st.name);
        off += st.len;
    }
    cm.state_len = c.state_len = off;
    Emitter em(c);
    for (const auto& kv : ast.constants) {
        check_dup(kv.first, "constant", 'k');
        Scope s;
        Val v = em.expr(kv.second, s);
        if (!v.is_const) throw DslError("constant `" + kv.first + "` is not a compile-time value");
        c.const_ix[kv.first] = (int)c.const_vals.size();
        c.const_vals.push_back(v);
    }
    // derived = declared names, then undeclared plain targets of the derive block in first-assignment order
    std::vector<std::string> dnames = ast.derived_decl;
    {
        std::vector<std::string> targets;
        collect_targets(ast.derive, targets);
        for (auto& t : targets) { bool seen = false; for (auto& d : dnames) if (d == t) seen = true; if (!seen) dnames.push_back(t); }
    }
    for (auto& d : dnames) {
        if (c.param_ix.count(d)) throw DslError("derived name `" + d + "` collides with primary parameter `" + d + "`");      // analyze.rs:725-748
        if (c.cov_ix.count(d) || c.state_by_name.count(d)) throw DslError("derived `" + d + "` conflicts with another name");
        declare(d, 'd');
        c.derived_ix[d] = (int)cm.derived.size();
        cm.derived.push_back(d);
    }
    cm.derived_len = (int)cm.derived.size();
    for (auto& o : ast.outputs_decl) {
        // validate_output_label_name, analyze.rs:1813-1822
        if (digits(o)) throw DslError("bare numeric output labels are not allowed in the DSL; use `outeq_" + o + "` instead");
        const std::string sfx = suffix_of(o, "input_");
        if (!sfx.empty()) throw DslError("`" + o + "` is a route label and cannot be used as an output target; use `outeq_" + sfx + "` here");
        if (c.output_ix.count(o)) throw DslError("duplicate output `" + o + "`");
        declare(o, 'o');
        c.output_ix[o] = (int)cm.outputs.size(); cm.outputs.push_back(o);
    }
    cm.output_len = (int)cm.outputs.size();
    if (cm.output_len == 0) throw DslError("model `" + ast.name + "` declares no outputs");

    // ---- routes: dense slots (execution.rs:575-590; metadata.rs:926-957) ----------------------------
    {
        bool uses_kinds = false;
        for (auto& r : ast.routes) if (r.has_kind) uses_kinds = true;
        int nb = 0, ni = 0, decl = 0, maxslot = -1;
        std::map<std::string, std::set<int>> route_kinds;      // label -> kinds seen (-1 = kind-less), analyze.rs:594-628
        for (auto& r : ast.routes) {
            if (digits(r.name)) throw DslError("bare numeric route labels are not allowed in the DSL; use `input_" + r.name + "` instead");
            const std::string sfx = suffix_of(r.name, "outeq_");
            if (!sfx.empty()) throw DslError("`" + r.name + "` is an output label and cannot be used as a route; use `input_" + sfx + "` here");
            const int rk = r.has_kind ? (int)r.kind : -1;
            auto seen = route_kinds.find(r.name);
            if (seen != route_kinds.end() && (rk == -1 || seen->second.count(-1) || seen->second.count(rk))) throw DslError("duplicate route `" + r.name + "`");
            route_kinds[r.name].insert(rk);
            declare(r.name, 'r');
            r.declaration_index = decl;
            if (uses_kinds && r.has_kind) r.index = (r.kind == RouteKind::Bolus) ? nb++ : ni++;
            else r.index = decl;
            ++decl;
            auto si = c.state_by_name.find(r.dest);
            if (si == c.state_by_name.end()) {
                std::vector<std::string> have;
                for (const auto& kv : c.state_by_name) have.push_back(kv.first);
                throw DslError("route `" + r.name + "` targets unknown state `" + r.dest + "`" + suggest(r.dest, have));
            }
            long long ix = 0;
            Scope s;
            if (r.dest_index) ix = em.const_int(r.dest_index, s);
            else if (si->second.is_array) throw DslError("route destination `" + r.dest + "` must be indexed");
            r.dest_offset = si->second.offset + (int)ix;
            if (r.has_kind && r.kind == RouteKind::Infusion && (r.lag || r.fa)) throw DslError("lag and bioavailability are bolus-only route properties (route `" + r.name + "`)");
            maxslot = std::max(maxslot, r.index);
            RouteInfo ri; ri.name = r.name; ri.has_kind = r.has_kind; ri.kind = r.kind; ri.index = r.index; ri.destination = r.dest_offset; ri.has_lag = (bool)r.lag; ri.has_bioavailability = (bool)r.fa;
            ri.declaration_index = r.declaration_index; ri.destination_name = r.dest;
            cm.routes.push_back(ri);
            if (r.lag) cm.has_lag = true;
            if (r.fa) cm.has_fa = true;
        }
        cm.route_len = maxslot + 1;
    }

    // ---- analytical structure ---------------------------------------------------------------------
    // ---- definite assignment -------------------------------------------------------------------------------
    Flow derive_flow;
    flow_stmts(ast.derive, c, Role::Derive, derive_flow);
    {
        auto check_block = [&](const std::vector<Stmt>& ss, Role role) {
            Flow f;
            f.avail = derive_flow.avail;
            flow_stmts(ss, c, role, f);
            return f;
        };
        check_block(ast.dynamics, Role::Dynamics);
        check_block(ast.init, Role::Init);
        check_block(ast.diffusion, Role::Diffusion);
        Flow of = check_block(ast.outputs, Role::Outputs);
        Flow rf;
        rf.avail = derive_flow.avail;
        for (const auto& r : ast.routes) { flow_reads(r.lag, c, rf); flow_reads(r.fa, c, rf); }
        std::vector<std::string> assigned;
        collect_output_targets(ast.outputs, assigned);
        for (const auto& o : cm.outputs) {
            if (of.targets.count(o)) continue;
            const bool ever = std::find(assigned.begin(), assigned.end(), o) != assigned.end();
            if (!ever) throw DslError("output `" + o + "` is declared in `outputs = ...` but never assigned");
            throw DslError("output `" + o + "` is not definitely assigned on all control-flow paths");
        }
    }
    std::vector<std::pair<bool, int>> kp_bind;   // (is_derived, index) per kernel parameter
    if (ast.kind == ModelKind::Analytical) {
        if (ast.structure.empty()) throw DslError("analytical model `" + ast.name + "` does not declare a structure");
        cm.analytical_kernel = analytical_kernel_index(ast.structure);
        if (cm.analytical_kernel < 0) throw DslError("unknown analytical structure `" + ast.structure + "`");
        if (kernel_state_count(cm.analytical_kernel) != cm.state_len)
            throw DslError("structure `" + ast.structure + "` needs " + std::to_string(kernel_state_count(cm.analytical_kernel)) + " state(s), model declares " + std::to_string(cm.state_len));
        for (const auto& n : analytical_kernel_params(cm.analytical_kernel)) {
            auto pi = c.param_ix.find(n);
            auto di = c.derived_ix.find(n);
            if (pi != c.param_ix.end() && di == c.derived_ix.end()) kp_bind.emplace_back(false, pi->second);
            else if (di != c.derived_ix.end() && pi == c.param_ix.end()) {
                if (!derive_flow.avail.count(n))
                    throw DslError("derived value `" + n + "` is not definitely assigned on all control-flow paths before analytical structure `" + ast.structure + "` uses it");
                kp_bind.emplace_back(true, di->second);
            }
            else throw DslError("analytical structure `" + ast.structure + "` requires `" + n + "` as a parameter or derived value");
        }
    } else if (!ast.structure.empty()) {
        throw DslError("`structure` is only valid for analytical models");
    }
    if (ast.kind == ModelKind::Sde && ast.diffusion.empty()) throw DslError("sde model `" + ast.name + "` has no diffusion (noise) statements");

    // ---- emit function bodies ------------------------------------------------------------------------
    std::ostringstream S;
    auto emit_fn = [&](const char* fname, Role role, const std::vector<Stmt>& body, Scope& scope, bool derive_sig) {
        std::vector<std::string> lets;
        collect_lets(body, lets);
        BodyEmitter be(em, c);
        scope.role = role;
        be.stmts(body, scope);
        if (derive_sig) S << "    PSI_DEV static void " << fname << "(double t, const double* x, const double* p, const double* cov, const double* rate, double* d) {\n";
        else S << "    PSI_DEV static void " << fname << kSig << " {\n";
        for (auto& l : lets) S << "        " << (scope.locals[l] == Ty::Bool ? "bool" : "double") << " L_" << l << " = " << (scope.locals[l] == Ty::Bool ? "false" : "0.0") << ";\n";
        if (role == Role::Dynamics) {
            // every state derivative starts at zero (the analyzer requires all to be assigned,
            // analyze.rs:2414-2432; zeroing keeps partial models and the rate injection well defined)
            for (int k = 0; k < cm.state_len; ++k) S << "        out[" << k << "] = 0.0;\n";
        }
        S << be.os.str();
    };

    // derive
    Scope sc_derive;
    cm.has_derive = !ast.derive.empty();
    if (cm.has_derive) { emit_fn("derive", Role::Derive, ast.derive, sc_derive, true); S << "    }\n"; }
    const int derive_deps = sc_derive.deps & (DEP_T | DEP_COV | DEP_STATE | DEP_RATE);

    // dynamics / drift (+ injected infusion rates)
    Scope sc_dyn;
    const bool has_dyn = ast.kind != ModelKind::Analytical;
    if (has_dyn) {
        if (ast.dynamics.empty()) throw DslError("model `" + ast.name + "` has no dynamics");
        emit_fn(ast.kind == ModelKind::Sde ? "drift" : "dynamics", Role::Dynamics, ast.dynamics, sc_dyn, false);
        for (const auto& r : ast.routes) {
            const bool carries_infusion = !r.has_kind || r.kind == RouteKind::Infusion;
            if (carries_infusion && !sc_dyn.routes_read.count(r.index)) {
                std::ostringstream inj;
                inj << "        out[" << r.dest_offset << "] += rate[" << r.index << "];   // infusion(" << r.name << ") -> state " << r.dest_offset << "\n";
                S << inj.str();
                cm.injection_lines.push_back(inj.str());
            }
        }
        for (auto& ri : cm.routes) ri.inject_input_to_destination = !sc_dyn.routes_read.count(ri.index);    // model_info.rs:151-154
        S << "    }\n";
        // every state must be touched by the block (validate_state_coverage, analyze.rs:2414-2432)
        std::vector<std::string> touched;
        collect_state_targets(ast.dynamics, touched);
        for (const auto& st : ast.states) {
            bool ok = false;
            for (const auto& t : touched) ok = ok || t == st.name;
            if (!ok) throw DslError(std::string(ast.kind == ModelKind::Sde ? "drift" : "dynamics") + " block does not assign `" + st.name + "`");
        }
    }
    // diffusion
    Scope sc_diff;
    if (ast.kind == ModelKind::Sde) { emit_fn("diffusion", Role::Diffusion, ast.diffusion, sc_diff, false); S << "    }\n"; }
    // outputs
    Scope sc_out;
    if (ast.outputs.empty()) throw DslError("model `" + ast.name + "` has no output equations");
    emit_fn("outputs", Role::Outputs, ast.outputs, sc_out, false); S << "    }\n";
    // init
    Scope sc_init;
    cm.has_init = !ast.init.empty();
    if (cm.has_init) { emit_fn("init", Role::Init, ast.init, sc_init, false); S << "    }\n"; }
    // lag / fa: per-route scalar functions selected by a switch (no per-thread arrays)
    auto emit_route_fn = [&](const char* fname, bool is_lag, const char* dflt) {
        S << "    PSI_DEV static double " << fname << "(int route, double t, const double* x, const double* p, const double* cov, const double* rate, const double* d) {\n";
        S << "        switch (route) {\n";
        for (const auto& r : ast.routes) {
            const ExprP& e = is_lag ? r.lag : r.fa;
            if (!e) continue;
            Scope s;
            Val v = Emitter::real(em.expr(e, s));
            S << "            case " << r.index << ": return " << v.code << ";\n";
        }
        S << "            default: return " << dflt << ";\n        }\n    }\n";
    };
    if (cm.has_lag) emit_route_fn("lag", true, "0.0");
    if (cm.has_fa) emit_route_fn("fa", false, "1.0");
    // bolus destinations (native.rs:572-597): bolus routes and kind-less routes
    S << "    PSI_DEV static int bolus_dest(int route) {\n        switch (route) {\n";
    for (const auto& r : ast.routes)
        if (!r.has_kind || r.kind == RouteKind::Bolus) S << "            case " << r.index << ": return " << r.dest_offset << ";\n";
    S << "            default: return -1;\n        }\n    }\n";
    // analytical parameter projection (native.rs:2736-2771)
    bool kp_uses_derived = false;
    if (ast.kind == ModelKind::Analytical) {
        S << "    PSI_DEV static void kparams(const double* p, const double* d, double* kp) {\n";
        for (size_t k = 0; k < kp_bind.size(); ++k) {
            S << "        kp[" << k << "] = " << (kp_bind[k].first ? "d[" : "p[") << kp_bind[k].second << "];\n";
            kp_uses_derived = kp_uses_derived || kp_bind[k].first;
        }
        S << "    }\n";
    }
    // Jacobian (ODE only): value pass (derive + dynamics into scratch), then one AD pass per state
    if (ast.kind == ModelKind::Ode) {
        S << "    PSI_DEV static void jacobian(double t, const double* x, const double* p, const double* cov, const double* rate, double* J) {\n";
        S << "        double d[" << std::max(1, cm.derived_len) << "] = {0.0};\n";
        S << "        double out[" << std::max(1, cm.state_len) << "] = {0.0};\n";
        S << "        for (int q = 0; q < NSTATE * NSTATE; ++q) J[q] = 0.0;\n";
        std::vector<std::string> lets;
        collect_lets(ast.derive, lets);
        collect_lets(ast.dynamics, lets);
        Scope vs;
        BodyEmitter vb(em, c);
        vs.role = Role::Derive; vb.stmts(ast.derive, vs);
        vs.role = Role::Dynamics; vb.stmts(ast.dynamics, vs);
        for (auto& l : lets) S << "        " << (vs.locals[l] == Ty::Bool ? "bool" : "double") << " L_" << l << " = " << (vs.locals[l] == Ty::Bool ? "false" : "0.0") << "; double DL_" << l << " = 0.0;\n";
        for (int k = 0; k < cm.derived_len; ++k) S << "        double DD_" << k << " = 0.0;\n";
        S << vb.os.str();
        for (int j = 0; j < cm.state_len; ++j) {
            S << "        {   // d/dx[" << j << "]\n";
            for (int k = 0; k < cm.derived_len; ++k) S << "            DD_" << k << " = 0.0;\n";
            for (auto& l : lets) if (vs.locals[l] != Ty::Bool) S << "            DL_" << l << " = 0.0;\n";
            Scope as = vs;
            as.ad = true; as.ad_wrt = j; as.ad_locals.clear(); as.ad_derived.clear();
            BodyEmitter ab(em, c);
            ab.indent = 3;
            as.role = Role::Derive; ab.stmts(ast.derive, as);
            as.role = Role::Dynamics; ab.stmts(ast.dynamics, as);
            S << ab.os.str();
            S << "        }\n";
        }
        S << "        (void)out; (void)d;\n    }\n";
    }

    // ---- flags -----------------------------------------------------------------------------------------
    const bool dyn_reads_derived = (sc_dyn.deps & DEP_DERIVED) != 0 || (sc_diff.deps & DEP_DERIVED) != 0;
    const bool rhs_uses_cov = (sc_dyn.deps & DEP_COV) || (sc_diff.deps & DEP_COV) || (dyn_reads_derived && (derive_deps & DEP_COV));
    // explicit time dependence of the right-hand side (t itself or covariates, directly or through derive):
    // Rosenbrock methods then need the df/dt term
    const bool rhs_time_dep = (sc_dyn.deps & (DEP_T | DEP_COV)) != 0 || (dyn_reads_derived && (derive_deps & (DEP_T | DEP_COV)) != 0);
    std::ostringstream H;
    H << "    static constexpr int KIND = " << (int)ast.kind << ";\n";
    H << "    static constexpr int NP = " << cm.parameters.size() << ", NCOV = " << cm.covariates.size() << ", NSTATE = " << cm.state_len
      << ", NROUTE = " << cm.route_len << ", NDER = " << cm.derived_len << ", NOUT = " << cm.output_len << ";\n";
    H << "    static constexpr int NPX = " << (cm.parameters.size() + em.slots.size()) << ";   // parameters + pair-invariant slots filled by prologue()\n";
    H << "    static constexpr int AKERNEL = " << cm.analytical_kernel << ";\n";
    H << "    static constexpr int DERIVE_DEPS = " << derive_deps << ";   // 1 t | 2 covariates | 4 states | 8 rates\n";
    H << "    static constexpr bool HAS_DERIVE = " << (cm.has_derive ? "true" : "false") << ", HAS_INIT = " << (cm.has_init ? "true" : "false")
      << ", HAS_LAG = " << (cm.has_lag ? "true" : "false") << ", HAS_FA = " << (cm.has_fa ? "true" : "false") << ";\n";
    H << "    static constexpr bool RHS_USES_COV = " << (rhs_uses_cov ? "true" : "false") << ", RHS_USES_DERIVED = " << (dyn_reads_derived ? "true" : "false")
      << ", KP_USES_DERIVED = " << (kp_uses_derived ? "true" : "false") << ", RHS_TIME_DEP = " << (rhs_time_dep ? "true" : "false") << ";\n";
    // resident CTAs per SM the kernel is compiled for (psi_engine.cuh PSI_DEFINE_ENTRY): measured per kernel family
    {
        int min_blocks = 6;
        if (ast.kind == ModelKind::Analytical && !cm.has_lag) {
            const int ak = cm.analytical_kernel;          // 0-3 one compartment, 4-7 two compartments, 8-11 three
            // the 48-register cap is for the plain kernels only: derive blocks, covariates or bioavailability
            // expressions need room, and spilling them would cost more than the occupancy buys
            const bool plain = !cm.has_derive && cm.covariates.empty() && !cm.has_fa;
            if (ak >= 0 && ak <= 3) min_blocks = plain ? 10 : 8;
            else if (ak >= 4 && ak <= 7) min_blocks = plain ? 8 : 6;
        }
        H << "    static constexpr int MIN_BLOCKS = " << min_blocks << ";\n";
    }
    // do the outputs (directly or through derive) read rate(route)?  If not, an observation needs no infusion scan
    H << "    static constexpr bool OBS_USES_RATE = " << ((((sc_out.deps | sc_derive.deps) & DEP_RATE) != 0) ? "true" : "false") << ";\n";
    // do the outputs read covariates or derived values at all?  If not, an observation needs neither the covariate
    // interpolation nor the derive block (the reference refreshes them unconditionally, native.rs:1044-1086; the
    // prediction is the same)
    H << "    static constexpr bool OBS_NEEDS_REFRESH = " << (((sc_out.deps & (DEP_DERIVED | DEP_COV)) != 0) ? "true" : "false") << ";\n";
    // pair-invariant slots: evaluated once per (subject, support point) pair right after the parameter load
    S << "    PSI_DEV static void prologue(double* p) {\n";
    for (size_t k = 0; k < em.slots.size(); ++k) S << "        p[" << (cm.parameters.size() + k) << "] = " << em.slots[k] << ";\n";
    S << "    }\n";
    cm.struct_body = H.str() + S.str();
    cm.id = fnv1a_hex(cm.struct_body);
    return cm;
}

ModelLabels CompiledModel::labels() const {
    ModelLabels L;
    L.routes = routes;
    L.outputs = outputs;
    L.covariates = covariates;
    L.route_len = route_len;
    L.nout = output_len;
    return L;
}

std::string CompiledModel::cuda_source(const std::vector<std::pair<int, std::string>>& entries, bool aot_register) const {
    std::ostringstream o;
    o << "// generated by pharmsol-b200 dsl_emit from model `" << name << "` (id " << id << ") — do not edit\n";
    o << "#include \"psi_engine.cuh\"\n";
    o << "namespace {\nstruct Model_" << id << " {\n" << struct_body << "};\n}  // namespace\n";
    for (const auto& e : entries) o << "PSI_DEFINE_ENTRY(Model_" << id << ", " << e.first << ", " << e.second << ")\n";
    if (aot_register) {
        o << "extern \"C\" void psi_aot_register(const char* id, int solver, const void* fn, const char* name);\n";
        o << "namespace { struct Reg_" << id << " { Reg_" << id << "() {\n";
        for (const auto& e : entries) o << "    psi_aot_register(\"" << id << "\", " << e.first << ", (const void*)&" << e.second << ", \"" << e.second << "\");\n";
        o << "} } reg_" << id << "; }\n";
    }
    return o.str();
}

// The host twin of the emitted model: the same function bodies as the CUDA translation unit, compiled by the system C++
// compiler into a cdylib that exports the reference's frozen compiled-backend ABI (src/dsl/compiled_backend_abi.rs:6-33;
// loader src/dsl/aot.rs:316-353, 404-470): pharmsol_dsl_api_version() == 2, the CompiledModelInfoEnvelope JSON, and one
// `extern "C" fn(t, states, params, covariates, routes, derived, out)` per function role the model has.
//  * dynamics / drift do NOT add the route inputs: the reference runtime injects them itself for routes with
//    inject_input_to_destination (native.rs apply_route_inputs_to_rates), while the device dynamics has them fused in;
//  * `out` may alias `states` (init) or `derived` (derive), as compiled_backend_abi.rs:147-180 allows;
//  * route_lag / route_bioavailability write only the slots of routes that declare the property (the caller pre-fills
//    0.0 / 1.0, native.rs:941-1018);
//  * the pair-invariant slots the device code hoists into p[NP..] are recomputed per call from `params`.
std::string CompiledModel::host_source() const {
    std::string body = struct_body;
    for (const auto& inj : injection_lines) {
        const size_t at = body.find(inj);
        if (at != std::string::npos) body.erase(at, inj.size());
    }
    auto b = [](bool v) { return v ? "true" : "false"; };
    const bool ode = kind == ModelKind::Ode, sde = kind == ModelKind::Sde;
    std::ostringstream env;
    env << "{\"abi_version\": 2, \"model\": " << model_info_json() << ", \"functions\": {\"derive\": " << b(has_derive) << ", \"dynamics\": " << b(ode)
        << ", \"outputs\": true, \"init\": " << b(has_init) << ", \"drift\": " << b(sde) << ", \"diffusion\": " << b(sde) << ", \"route_lag\": " << b(has_lag)
        << ", \"route_bioavailability\": " << b(has_fa) << "}}";
    std::ostringstream o;
    o << "// generated by pharmsol-b200 dsl_emit from model `" << name << "` (id " << id << ") — host twin, frozen compiled-backend ABI; do not edit\n";
    o << "#include <cmath>\n#include <cstddef>\n#include <cstdint>\n#include <cstdlib>\n#include <limits>\n";
    o << "#define PSI_DEV inline\n";
    o << "namespace psi {\n"
         "inline double psi_nan() { return std::numeric_limits<double>::quiet_NaN(); }\n"
         "inline double psi_inf() { return std::numeric_limits<double>::infinity(); }\n"
         "inline double pow_half(double x) { return std::sqrt(x); }\n"
         "inline double pow_quarter(double x) { return std::sqrt(std::sqrt(x)); }\n"
         "inline double pow_three_quarters(double x) { const double s = std::sqrt(x); return s * std::sqrt(s); }\n"
         "inline double pow_three_halves(double x) { return x * std::sqrt(x); }\n"
         "inline double pow_2(double x) { return x * x; }\n"
         "inline double pow_3(double x) { return (x * x) * x; }\n"
         "inline double pow_4(double x) { const double q = x * x; return q * q; }\n"
         "inline double fdiv(double a, double b) { return a / b; }\n"
         "}  // namespace psi\n";
    o << "namespace {\nstruct Model {\n" << body << "};\n";
    o << "inline void load_params(const double* params, double* p) {\n    for (int k = 0; k < Model::NP; ++k) p[k] = params[k];\n    Model::prologue(p);\n}\n";
    o << "const char kModelInfoJson[] = R\"PKMJSON(" << env.str() << ")PKMJSON\";\n}  // namespace\n";
    o << "extern \"C\" {\n";
    o << "uint32_t pharmsol_dsl_api_version(void) { return 2u; }\n";
    o << "const uint8_t* pharmsol_dsl_model_info_json_ptr(void) { return reinterpret_cast<const uint8_t*>(kModelInfoJson); }\n";
    o << "size_t pharmsol_dsl_model_info_json_len(void) { return sizeof(kModelInfoJson) - 1; }\n";
    const char* sig = "(double t, const double* states, const double* params, const double* covariates, const double* routes, const double* derived, double* out)";
    const char* pre = "    double p[Model::NPX > 0 ? Model::NPX : 1];\n    load_params(params, p);\n";
    if (has_derive) o << "void pharmsol_dsl_kernel_derive" << sig << " {\n" << pre << "    (void)derived;\n    Model::derive(t, states, p, covariates, routes, out);\n}\n";
    if (ode) o << "void pharmsol_dsl_kernel_dynamics" << sig << " {\n" << pre << "    Model::dynamics(t, states, p, covariates, routes, derived, out);\n}\n";
    o << "void pharmsol_dsl_kernel_outputs" << sig << " {\n" << pre << "    Model::outputs(t, states, p, covariates, routes, derived, out);\n}\n";
    if (has_init) o << "void pharmsol_dsl_kernel_init" << sig << " {\n" << pre << "    Model::init(t, states, p, covariates, routes, derived, out);\n}\n";
    if (sde) {
        o << "void pharmsol_dsl_kernel_drift" << sig << " {\n" << pre << "    Model::drift(t, states, p, covariates, routes, derived, out);\n}\n";
        o << "void pharmsol_dsl_kernel_diffusion" << sig << " {\n" << pre << "    Model::diffusion(t, states, p, covariates, routes, derived, out);\n}\n";
    }
    auto route_fn = [&](const char* symbol, const char* member, bool lag) {
        o << "void " << symbol << sig << " {\n" << pre;
        for (const auto& r : routes)
            if (lag ? r.has_lag : r.has_bioavailability)
                o << "    out[" << r.index << "] = Model::" << member << "(" << r.index << ", t, states, p, covariates, routes, derived);\n";
        o << "}\n";
    };
    if (has_lag) route_fn("pharmsol_dsl_kernel_route_lag", "lag", true);
    if (has_fa) route_fn("pharmsol_dsl_kernel_route_bioavailability", "fa", false);
    o << "}  // extern \"C\"\n";
    return o.str();
}

static std::string json_list(const std::vector<std::string>& v) {
    std::string s = "[";
    for (size_t i = 0; i < v.size(); ++i) s += (i ? ", \"" : "\"") + v[i] + "\"";
    return s + "]";
}
static std::string camel(const std::string& snake) {
    std::string out;
    bool up = true;
    for (char ch : snake) {
        if (ch == '_') { up = true; continue; }
        out += up ? (char)std::toupper((unsigned char)ch) : ch;
        up = false;
    }
    return out;
}

// The serde JSON of NativeModelInfo (src/dsl/model_info.rs:17-101): enum variants by name (`"Ode"`, `"Locf"`,
// `"OneCompartment"`), options as null.  `id` (module-cache key) and `analytical_structure` (the DSL spelling) are
// extra keys a serde reader ignores.
std::string CompiledModel::model_info_json() const {
    std::ostringstream o;
    static const char* kinds[3] = {"Ode", "Analytical", "Sde"};
    auto b = [](bool v) { return v ? "true" : "false"; };
    o << "{\"name\": \"" << name << "\", \"kind\": \"" << kinds[(int)kind] << "\", \"id\": \"" << id << "\", \"parameters\": " << json_list(parameters)
      << ", \"derived\": " << json_list(derived) << ", \"covariates\": [";
    for (size_t i = 0; i < covariates.size(); ++i) {
        const std::string& ip = i < covariate_interpolation.size() ? covariate_interpolation[i] : std::string();
        o << (i ? ", " : "") << "{\"name\": \"" << covariates[i] << "\", \"index\": " << i << ", \"interpolation\": " << (ip.empty() ? std::string("null") : "\"" + camel(ip) + "\"") << "}";
    }
    o << "], \"states\": [";
    for (size_t i = 0; i < state_decls.size(); ++i)
        o << (i ? ", " : "") << "{\"name\": \"" << state_decls[i].first << "\", \"offset\": " << state_decls[i].second << "}";
    o << "], \"routes\": [";
    for (size_t i = 0; i < routes.size(); ++i) {
        const auto& r = routes[i];
        o << (i ? ", " : "") << "{\"name\": \"" << r.name << "\", \"declaration_index\": " << r.declaration_index << ", \"index\": " << r.index << ", \"kind\": "
          << (r.has_kind ? (r.kind == RouteKind::Bolus ? "\"Bolus\"" : "\"Infusion\"") : "null") << ", \"destination_offset\": " << r.destination
          << ", \"destination_name\": \"" << r.destination_name << "\", \"has_lag\": " << b(r.has_lag) << ", \"has_bioavailability\": " << b(r.has_bioavailability)
          << ", \"inject_input_to_destination\": " << b(r.inject_input_to_destination) << "}";
    }
    o << "], \"outputs\": [";
    for (size_t i = 0; i < outputs.size(); ++i) o << (i ? ", " : "") << "{\"name\": \"" << outputs[i] << "\", \"index\": " << i << "}";
    o << "], \"state_len\": " << state_len << ", \"derived_len\": " << derived_len << ", \"output_len\": " << output_len << ", \"route_len\": " << route_len
      << ", \"analytical\": " << (analytical_kernel >= 0 ? "\"" + camel(kKernelNames[analytical_kernel]) + "\"" : std::string("null"))
      << ", \"analytical_structure\": " << (analytical_kernel >= 0 ? std::string("\"") + kKernelNames[analytical_kernel] + "\"" : std::string("null"))
      << ", \"particles\": " << (kind == ModelKind::Sde && particles > 0 ? std::to_string(particles) : std::string("null")) << "}";
    return o.str();
}

static DslError render_diagnostic(const DslError& e, const char* code, const std::string& src) {
    const std::string what = e.what();
    const size_t nl = what.find('\n');
    std::string out = std::string("error[") + code + "]: " + what.substr(0, nl);
    if (e.pos >= 0) {       // line_info, diagnostic.rs:679-698 (columns count characters, not bytes)
        const size_t off = std::min((size_t)e.pos, src.size());
        size_t line = 1, line_start = 0;
        for (size_t i = 0; i < off; ++i) if (src[i] == '\n') { ++line; line_start = i + 1; }
        size_t column = 1;
        for (size_t i = line_start; i < off; ++i) if (((unsigned char)src[i] & 0xC0) != 0x80) ++column;
        out += "\n  --> line " + std::to_string(line) + ", column " + std::to_string(column);
    }
    if (nl != std::string::npos) {      // "  note: ..." / "  help: ..." / "  suggestion: ..." -> "  = note: ..."
        std::istringstream rest(what.substr(nl + 1));
        std::string l;
        while (std::getline(rest, l)) {
            const size_t a = l.find_first_not_of(' ');
            out += "\n  = " + (a == std::string::npos ? std::string() : l.substr(a));
        }
    }
    return DslError(out, e.pos);
}

CompiledModel compile_source(const std::string& src) {
    ModelAst ast;
    try { ast = parse_model(src); } catch (const DslError& e) { throw render_diagnostic(e, "DSL1000", src); }
    try { return compile_model(ast); } catch (const DslError& e) { throw render_diagnostic(e, "DSL2000", src); }
}

}  // namespace dsl
}  // namespace pharmsol
