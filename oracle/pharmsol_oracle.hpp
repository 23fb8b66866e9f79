// pharmsol_oracle.hpp — CPU restatement of the pharmsol psi-matrix hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build, load or
// call it, and there only as the checker (or the timed CPU baseline), never as the thing shipped.
//
// Every function cites the reference file:line (relative to /root/reference) that it restates.
// The reference is pure Rust and cannot be compiled in this environment (no cargo/rustc), so this
// is a line-by-line restatement, kept deliberately literal: same expression trees, same order of
// operations, same event ordering, same error conditions.
//
// PARITY PINNING STATUS
//   * pinned by the reference's own literal anchors: lognormpdf(0,0,1) = -0.9189385332046727
//     (likelihood/distributions.rs:112-118), cdf(mean) = ln 0.5 (:141-150), the `seq_eq`
//     accumulation fixture == 2.5 exactly (analytical/mod.rs:492-527), the event-ordering /
//     lag / bioavailability unit fixtures (data/structs.rs:1148-1345), covariate segment
//     fixtures (data/covariate.rs:459-828) and the differential analytical<->ODE fixtures
//     (analytical/*_models.rs tests), all re-run in tests/test_oracle_*.py;
//   * pinned independently by scipy.linalg.expm / mpmath / SciPy Radau golden vectors
//     committed under tests/golden/ (generator scripts committed beside them);
//   * "parity unpinned" at two third-party boundaries that are not under /root/reference:
//     diffsol =0.16.1 (ODE time stepping: BDF / TR-BDF2 / ESDIRK34 / Tsit45 and its step
//     controller) and rand 0.10 / rand_distr 0.6 (unseeded thread-local ChaCha stream for SDE).
//     For those the oracle restates the *published* algorithm (Tsitouras 5(4), Dormand-Prince
//     5(4), Euler-Maruyama as written in sde/em.rs) and parity is tolerance / statistical.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <limits>
#include <map>
#include <optional>
#include <random>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace orc {

using V = std::vector<double>;

// ---------------------------------------------------------------------------------------------
// Errors — one code per PharmsolError / ErrorModelError variant on the hot path
// (src/error/mod.rs:14-49, src/data/error_model.rs ErrorModelError).  Codes are shared with
// include/pharmsol_cuda.h so tests can compare status codes directly.
// ---------------------------------------------------------------------------------------------
enum ErrCode : int {
    OK = 0,
    NonFiniteLikelihood = 1,
    NegativeSigma = 2,
    NonFiniteSigma = 3,
    InvalidOutputEquation = 4,
    NoneErrorModel = 5,
    MissingErrorModel = 6,
    SolverFailure = 7,
    InputOutOfRange = 8,
    OuteqOutOfRange = 9,
    UnknownInputLabel = 10,
    UnknownOutputLabel = 11,
    ImaginaryRoots = 12,  // the reference panics (two_compartment_models.rs:20-22, three_...:32-34)
    UnsupportedInputRouteKind = 13,
    MissingCovariate = 14,  // fetch_cov! panics (lib.rs:433-443)
    OtherError = 15,
    MissingObservation = 16,
};

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// ---------------------------------------------------------------------------------------------
// Data model — src/data/{event,structs,covariate,error_model}.rs
// ---------------------------------------------------------------------------------------------
struct ErrorPoly { double c0 = 0, c1 = 0, c2 = 0, c3 = 0; };      // data/error_model.rs:17
enum class Censor { None = 0, BLOQ = 1, ALOQ = 2 };                 // data/event.rs:559-567
enum class EventKind { Observation = 0, Bolus = 1, Infusion = 2 }; // rank order event.rs:293-299

// One struct for the three Event variants (data/event.rs:107,121,203,575).
struct Event {
    EventKind kind = EventKind::Observation;
    double time = 0.0;
    // doses
    double amount = 0.0;
    double duration = 0.0;
    std::string label;      // InputLabel / OutputLabel (public label from data)
    long index = -1;        // resolved dense input / outeq index (set_input / set_outeq)
    // observations
    bool has_value = false;
    double value = 0.0;
    bool has_poly = false;
    ErrorPoly poly;
    Censor cens = Censor::None;
    int occasion = 0;
};

// data/event.rs:292-304 — time by total_cmp, then Observation < Bolus < Infusion.
inline bool total_less(double a, double b) {
    // f64::total_cmp restated for the values that can occur (no NaN payload games needed):
    auto key = [](double x) {
        int64_t bits;
        static_assert(sizeof(bits) == sizeof(x), "");
        std::memcpy(&bits, &x, sizeof(bits));
        bits ^= (int64_t)((uint64_t)(bits >> 63) >> 1);
        return bits;
    };
    return key(a) < key(b);
}
inline bool event_less(const Event& a, const Event& b) {
    if (total_less(a.time, b.time)) return true;
    if (total_less(b.time, a.time)) return false;
    return (int)a.kind < (int)b.kind;
}

// data/covariate.rs:26-66, 165-241
struct CovariateSegment {
    double from = 0.0;
    bool has_to = false;
    double to = 0.0;
    bool linear = false;
    double slope = 0.0, intercept = 0.0, value = 0.0;
    bool in_interval(double t) const { return from <= t && (!has_to || t < to); }
};

struct Covariate {
    std::string name;
    std::vector<std::pair<double, double>> observations;
    std::vector<CovariateSegment> segments;
    bool fixed = false;

    // covariate.rs:135-147
    void add_observation(double time, double value) {
        for (auto& o : observations)
            if (o.first == time) { o.second = value; build_segments(); return; }
        observations.emplace_back(time, value);
        build_segments();
    }
    // covariate.rs:189-212
    void build_segments() {
        std::stable_sort(observations.begin(), observations.end(),
                         [](auto& l, auto& r) { return total_less(l.first, r.first); });
        segments.clear();
        for (size_t i = 0; i < observations.size(); ++i) {
            const auto& cur = observations[i];
            bool has_next = i + 1 < observations.size();
            CovariateSegment s;
            s.from = cur.first;
            s.has_to = has_next;
            if (has_next) s.to = observations[i + 1].first;
            if (fixed) {
                s.linear = false; s.value = cur.second;
            } else if (has_next) {
                const auto& nxt = observations[i + 1];
                double slope = (nxt.second - cur.second) / (nxt.first - cur.first);
                s.linear = true; s.slope = slope; s.intercept = cur.second - slope * cur.first;
            } else {
                s.linear = false; s.value = cur.second;
            }
            segments.push_back(s);
        }
    }
    // covariate.rs:216-241.  Returns false for CovariateError::MissingSegments.
    bool interpolate(double time, double& out) const {
        if (segments.empty()) return false;
        for (const auto& s : segments) {
            if (s.in_interval(time)) {
                out = s.linear ? s.slope * time + s.intercept : s.value;
                return true;
            }
        }
        if (!observations.empty()) {
            if (time < observations.front().first) { out = observations.front().second; return true; }
            if (time >= observations.back().first) { out = observations.back().second; return true; }
        }
        return false;
    }
};

// data/covariate.rs Covariates: BTreeMap<String, Covariate>
struct Covariates {
    std::map<std::string, Covariate> map;
    const Covariate* get_covariate(const std::string& name) const {
        auto it = map.find(name);
        return it == map.end() ? nullptr : &it->second;
    }
    void add_observation(const std::string& name, double t, double v, bool fixed = false) {
        auto it = map.find(name);
        if (it == map.end()) {
            Covariate c; c.name = name; c.fixed = fixed;
            it = map.emplace(name, c).first;
        }
        it->second.add_observation(t, v);
    }
};

// fetch_cov! (src/lib.rs:433-443): missing covariate panics, interpolate().unwrap() panics.
inline double fetch_cov(const Covariates& cov, double t, const char* name) {
    const Covariate* c = cov.get_covariate(name);
    if (!c) throw Error(MissingCovariate, std::string("Covariate ") + name + " not found");
    double v;
    if (!c->interpolate(t, v)) throw Error(MissingCovariate, "MissingSegments");
    return v;
}

struct Occasion {                       // data/structs.rs:556-560
    std::vector<Event> events;
    Covariates covariates;
    int index = 0;
    void sort() { std::stable_sort(events.begin(), events.end(), event_less); }  // structs.rs:669-671
    void add_event(const Event& e) { events.push_back(e); sort(); }              // structs.rs:713-716
    double initial_time() const {                                                // structs.rs:782-793
        if (events.empty()) return 0.0;
        double m = events[0].time;
        for (auto& e : events) if (e.time < m) m = e.time;
        return m;
    }
};

struct Subject {                        // data/structs.rs:352
    std::string id;
    std::vector<Occasion> occasions;
};

// data/builder.rs:84-362
struct SubjectBuilder {
    std::string id;
    std::vector<Occasion> occasions;
    Occasion current;
    Covariates covariates;
    std::optional<Event> last;

    explicit SubjectBuilder(std::string i) : id(std::move(i)) { current.index = 0; }
    SubjectBuilder& event(Event e) {
        e.occasion = current.index;
        last = e;
        current.add_event(e);
        return *this;
    }
    SubjectBuilder& bolus(double t, double amt, const std::string& input) {
        Event e; e.kind = EventKind::Bolus; e.time = t; e.amount = amt; e.label = input; return event(e);
    }
    SubjectBuilder& infusion(double t, double amt, const std::string& input, double dur) {
        Event e; e.kind = EventKind::Infusion; e.time = t; e.amount = amt; e.duration = dur; e.label = input;
        return event(e);
    }
    SubjectBuilder& observation(double t, double v, const std::string& outeq) {
        Event e; e.kind = EventKind::Observation; e.time = t; e.has_value = true; e.value = v; e.label = outeq;
        return event(e);
    }
    SubjectBuilder& censored_observation(double t, double v, const std::string& outeq, Censor c) {
        Event e; e.kind = EventKind::Observation; e.time = t; e.has_value = true; e.value = v; e.label = outeq;
        e.cens = c; return event(e);
    }
    SubjectBuilder& missing_observation(double t, const std::string& outeq) {
        Event e; e.kind = EventKind::Observation; e.time = t; e.has_value = false; e.label = outeq; return event(e);
    }
    SubjectBuilder& observation_with_error(double t, double v, const std::string& outeq, ErrorPoly p, Censor c) {
        Event e; e.kind = EventKind::Observation; e.time = t; e.has_value = true; e.value = v; e.label = outeq;
        e.has_poly = true; e.poly = p; e.cens = c; return event(e);
    }
    // builder.rs:263-322
    SubjectBuilder& repeat(size_t n, double delta) {
        if (!last) return *this;
        Event proto = *last;
        for (size_t i = 1; i <= n; ++i) {
            Event e = proto;
            e.time = proto.time + delta * (double)i;
            event(e);
        }
        return *this;
    }
    // builder.rs:331-342
    SubjectBuilder& reset() {
        int block_index = current.index + 1;
        current.sort();
        current.covariates = covariates;
        occasions.push_back(current);
        current = Occasion();
        current.index = block_index;
        covariates = Covariates();
        last.reset();
        return *this;
    }
    SubjectBuilder& covariate(const std::string& name, double t, double v) {
        covariates.add_observation(name, t, v);
        return *this;
    }
    Subject build() {
        reset();
        Subject s; s.id = id; s.occasions = occasions; return s;
    }
};

// ---------------------------------------------------------------------------------------------
// Error models — data/error_model.rs:150, 677-686, 786-812, 1045-1080
// ---------------------------------------------------------------------------------------------
enum class ErrKind { None = 0, Additive = 1, Proportional = 2 };
struct AssayErrorModel {
    ErrKind kind = ErrKind::None;
    double factor = 0.0;   // lambda (additive) or gamma (proportional); Factor::value()
    ErrorPoly poly;
};
struct AssayErrorModels { std::vector<AssayErrorModel> models; };

// ---------------------------------------------------------------------------------------------
// Predictions & likelihood — likelihood/{prediction,subject,distributions}.rs
// ---------------------------------------------------------------------------------------------
struct Prediction {                      // likelihood/prediction.rs:18-27
    double time = 0.0;
    bool has_obs = false;
    double observation = 0.0;
    double prediction = 0.0;
    size_t outeq = 0;
    bool has_poly = false;
    ErrorPoly errorpoly;
    V state;
    int occasion = 0;
    Censor censoring = Censor::None;
};

constexpr double LOG_2PI = 1.8378770664093453;   // distributions.rs:12

// ---------------------------------------------------------------------------------------------
// ResidualErrorModel(s): prediction-based sigma for parametric algorithms — data/residual_error.rs:69-426
// ---------------------------------------------------------------------------------------------
struct ResidualErrorModel {
    enum Kind { Missing = 0, Constant = 1, Proportional = 2, Combined = 3, Exponential = 4 } kind = Missing;
    double a = 0.0, b = 0.0;     // Constant{a} | Proportional{b} | Combined{a,b} | Exponential{sigma = a}
    double sigma(double prediction) const {                       // residual_error.rs:178-197
        double raw = 0.0;
        switch (kind) {
            case Constant: raw = a; break;
            case Proportional: raw = b * std::fabs(prediction); break;
            case Combined: raw = std::sqrt(a * a + (b * b) * (prediction * prediction)); break;
            case Exponential: raw = a; break;
            default: break;
        }
        return std::fmax(raw, std::sqrt(std::numeric_limits<double>::epsilon()));
    }
    double log_likelihood(double observation, double prediction) const {   // residual_error.rs:265-271
        const double s = sigma(prediction);
        const double nr = (observation - prediction) / s;
        return -0.5 * (std::log(6.283185307179586476925286766559) + 2.0 * std::log(s) + nr * nr);
    }
};
struct ResidualErrorModels {
    std::vector<ResidualErrorModel> models;     // indexed by outeq; Missing = no model
    // residual_error.rs:413-426: missing model for an outeq => -inf
    double total(const std::vector<std::pair<size_t, std::pair<double, double>>>& outeq_obs_pred) const {
        double t = 0.0;
        for (const auto& r : outeq_obs_pred) {
            if (r.first >= models.size() || models[r.first].kind == ResidualErrorModel::Missing) return -std::numeric_limits<double>::infinity();
            t += models[r.first].log_likelihood(r.second.first, r.second.second);
        }
        return t;
    }
};

// distributions.rs:31-34
inline double lognormpdf(double obs, double pred, double sigma) {
    double diff = obs - pred;
    return -0.5 * LOG_2PI - std::log(sigma) - (diff * diff) / (2.0 * sigma * sigma);
}
// statrs 0.19 Normal::cdf(x) = 0.5 * erfc((mean - x) / (std_dev * sqrt(2))) (third-party, restated)
inline double normal_cdf(double x, double mean, double sd) {
    return 0.5 * std::erfc((mean - x) / (sd * std::sqrt(2.0)));
}
// distributions.rs:53-71.  Normal::new fails for sigma <= 0 / non-finite -> NegativeSigma.
inline double lognormcdf(double obs, double pred, double sigma) {
    if (!(sigma > 0.0) || !std::isfinite(sigma) || std::isnan(pred)) throw Error(NegativeSigma, "Normal::new");
    double cdf = normal_cdf(obs, pred, sigma);
    if (cdf <= 0.0) {
        double z = (obs - pred) / sigma;
        if (z < -37.0) return lognormpdf(obs, pred, sigma) - std::log(std::fabs(z));
        throw Error(NegativeSigma, "lognormcdf numerical issue");
    }
    return std::log(cdf);
}
// distributions.rs:89-103
inline double lognormccdf(double obs, double pred, double sigma) {
    if (!(sigma > 0.0) || !std::isfinite(sigma) || std::isnan(pred)) throw Error(NegativeSigma, "Normal::new");
    double sf = 1.0 - normal_cdf(obs, pred, sigma);
    if (sf <= 0.0) {
        double z = (obs - pred) / sigma;
        if (z > 37.0) return lognormpdf(obs, pred, sigma) - std::log(z);
        throw Error(NegativeSigma, "lognormccdf numerical issue");
    }
    return std::log(sf);
}

// error_model.rs:1045-1080 (AssayErrorModel::sigma) wrapped by :677-686 (AssayErrorModels::sigma)
inline double sigma_for(const AssayErrorModels& em, const Prediction& p) {
    size_t outeq = p.outeq;
    if (outeq >= em.models.size()) throw Error(InvalidOutputEquation, "InvalidOutputEquation");
    const AssayErrorModel& m = em.models[outeq];
    if (m.kind == ErrKind::None) throw Error(NoneErrorModel, "NoneErrorModel");
    if (!p.has_obs) throw Error(MissingObservation, "MissingObservation");
    ErrorPoly ep = p.has_poly ? p.errorpoly : m.poly;
    double o = p.observation;
    double alpha = ep.c0 + ep.c1 * o + ep.c2 * (o * o) + ep.c3 * (o * o * o);
    double sigma;
    if (m.kind == ErrKind::Additive) sigma = std::sqrt(alpha * alpha + m.factor * m.factor);
    else sigma = m.factor * alpha;
    if (sigma < 0.0) throw Error(NegativeSigma, "NegativeSigma");
    if (!std::isfinite(sigma)) throw Error(NonFiniteSigma, "NonFiniteSigma");
    return sigma;
}

// prediction.rs:105-125
inline double prediction_log_likelihood(const Prediction& p, const AssayErrorModels& em) {
    if (!p.has_obs) return 0.0;
    double sigma = sigma_for(em, p);
    double ll;
    switch (p.censoring) {
        case Censor::None: ll = lognormpdf(p.observation, p.prediction, sigma); break;
        case Censor::BLOQ: ll = lognormcdf(p.observation, p.prediction, sigma); break;
        default:           ll = lognormccdf(p.observation, p.prediction, sigma); break;
    }
    if (std::isfinite(ll)) return ll;
    throw Error(NonFiniteLikelihood, "NonFiniteLikelihood");
}

// subject.rs:63-78
inline double subject_log_likelihood(const std::vector<Prediction>& preds, const AssayErrorModels& em) {
    if (preds.empty()) return 0.0;
    double total = 0.0;
    for (const auto& p : preds)
        if (p.has_obs) total += prediction_log_likelihood(p, em);
    return total;
}

// ---------------------------------------------------------------------------------------------
// Model metadata (only what label resolution needs) — equation/metadata.rs:236-275, 926-957
// ---------------------------------------------------------------------------------------------
enum class RouteKind { Bolus = 0, Infusion = 1 };
struct Route {
    std::string name;
    RouteKind kind;
    int input_index = 0;        // separate ordinal spaces per kind (metadata.rs:926-957)
    int destination = -1;       // destination state index (descriptive for analytical, A.3)
    bool inject = false;        // inject_input_to_destination
};
struct Metadata {
    bool present = false;
    std::vector<Route> routes;
    std::vector<std::string> outputs;

    void add_route(const std::string& name, RouteKind kind, int dest, bool inject = false) {
        int idx = 0;
        for (auto& r : routes) if (r.kind == kind) ++idx;
        routes.push_back(Route{name, kind, idx, dest, inject});
        present = true;
    }
    int ndrugs() const {
        int b = 0, i = 0;
        for (auto& r : routes) (r.kind == RouteKind::Bolus ? b : i)++;
        return std::max(b, i);
    }
};

inline bool is_bare_numeric_label(const std::string& s) {
    if (s.empty()) return false;
    for (char c : s) if (c < '0' || c > '9') return false;
    return true;
}
// InputLabel::index / OutputLabel::index (data/event.rs:140-142, 222-224): str::parse::<usize>
inline long parse_usize(const std::string& s) {
    if (s.empty()) return -1;
    size_t i = 0;
    if (s[0] == '+') i = 1;  // Rust usize::from_str accepts a leading '+'
    if (i >= s.size()) return -1;
    long v = 0;
    for (; i < s.size(); ++i) {
        if (s[i] < '0' || s[i] > '9') return -1;
        v = v * 10 + (s[i] - '0');
        if (v > (1L << 40)) return -1;
    }
    return v;
}

// ---------------------------------------------------------------------------------------------
// Closure types — src/simulator/mod.rs:41-197
// ---------------------------------------------------------------------------------------------
using DiffEq = std::function<void(const V& x, const V& p, double t, V& dx, const V& bolus, const V& rateiv, const Covariates& cov)>;
using AnalyticalEq = std::function<V(const V& x, const V& p, double t, const V& rateiv, const Covariates& cov)>;
using SecEq = std::function<void(V& p, double t, const Covariates& cov)>;
using LagFa = std::function<std::map<int, double>(const V& p, double t, const Covariates& cov)>;
using Init = std::function<void(const V& p, double t, const Covariates& cov, V& x)>;
using Out = std::function<void(const V& x, const V& p, double t, const Covariates& cov, V& y)>;
using Drift = std::function<void(const V& x, const V& p, double t, V& dx, const V& rateiv, const Covariates& cov)>;
using Diffusion = std::function<void(const V& p, V& d)>;

enum class EqnKind { ODE = 0, Analytical = 1, SDE = 2 };   // equation/mod.rs:580-586
enum class OdeSolver { Tsit45 = 0, Dopri5 = 1 };

struct Model {
    std::string name;
    EqnKind kind = EqnKind::Analytical;
    int nstates = 0, ndrugs = 0, nout = 0;
    Metadata metadata;
    AnalyticalEq eq;
    SecEq seq_eq;
    DiffEq diffeq;
    Drift drift;
    Diffusion diffusion;
    LagFa lag, fa;
    Init init;
    Out out;
    // ODE options (ode/mod.rs:40-41, 135-166)
    OdeSolver solver = OdeSolver::Tsit45;
    double rtol = 1e-4, atol = 1e-4;
    // SDE (sde/mod.rs)
    int nparticles = 1;
    std::vector<int> injected_bolus_destination;   // -1 = none (sde/mod.rs:46-79)
};

// ---------------------------------------------------------------------------------------------
// Label resolution — equation/mod.rs:192-273
// ---------------------------------------------------------------------------------------------
inline long resolve_input_label(const Model& m, const std::string& label, RouteKind kind) {
    if (m.metadata.present) {
        auto find = [&](RouteKind k) -> const Route* {
            for (auto& r : m.metadata.routes) if (r.kind == k && r.name == label) return &r;
            if (!is_bare_numeric_label(label)) return nullptr;
            std::string alias = "input_" + label;
            for (auto& r : m.metadata.routes) if (r.kind == k && r.name == alias) return &r;
            return nullptr;
        };
        if (const Route* r = find(kind)) return r->input_index;
        RouteKind other = kind == RouteKind::Bolus ? RouteKind::Infusion : RouteKind::Bolus;
        if (find(other)) throw Error(UnsupportedInputRouteKind, "UnsupportedInputRouteKind " + label);
        throw Error(UnknownInputLabel, "unknown input label " + label);
    }
    long idx = parse_usize(label);
    if (idx < 0) throw Error(UnknownInputLabel, "unknown input label " + label);
    return idx;
}
inline long resolve_output_label(const Model& m, const std::string& label) {
    if (m.metadata.present) {
        for (size_t i = 0; i < m.metadata.outputs.size(); ++i) if (m.metadata.outputs[i] == label) return (long)i;
        if (is_bare_numeric_label(label)) {
            std::string alias = "outeq_" + label;
            for (size_t i = 0; i < m.metadata.outputs.size(); ++i) if (m.metadata.outputs[i] == alias) return (long)i;
        }
        throw Error(UnknownOutputLabel, "unknown output label " + label);
    }
    long idx = parse_usize(label);
    if (idx < 0) throw Error(UnknownOutputLabel, "unknown output label " + label);
    return idx;
}

// data/structs.rs:611-690 (process_events = clone + add_lagtime + add_bioavailability) after
// equation/mod.rs:247-273 (resolve labels on the clone).
inline std::vector<Event> resolve_occasion_events(const Model& m, const Occasion& occ, const V& params) {
    Occasion resolved = occ;
    for (auto& e : resolved.events) {
        switch (e.kind) {
            case EventKind::Bolus:       e.index = resolve_input_label(m, e.label, RouteKind::Bolus); break;
            case EventKind::Infusion:    e.index = resolve_input_label(m, e.label, RouteKind::Infusion); break;
            case EventKind::Observation: e.index = resolve_output_label(m, e.label); break;
        }
    }
    const Covariates& cov = occ.covariates;
    // add_lagtime (structs.rs:611-646)
    bool shifted = false;
    for (auto& e : resolved.events) {
        if (e.kind != EventKind::Bolus) continue;
        if (e.index < 0) continue;
        auto lagtime = m.lag ? m.lag(params, e.time, cov) : std::map<int, double>{};
        auto it = lagtime.find((int)e.index);
        if (it != lagtime.end() && it->second != 0.0) { e.time += it->second; shifted = true; }
    }
    if (shifted) resolved.sort();
    // add_bioavailability (structs.rs:648-667): fa evaluated at the (already lagged) bolus time
    for (auto& e : resolved.events) {
        if (e.kind != EventKind::Bolus) continue;
        if (e.index < 0) continue;
        auto fa = m.fa ? m.fa(params, e.time, cov) : std::map<int, double>{};
        auto it = fa.find((int)e.index);
        if (it != fa.end()) e.amount = e.amount * it->second;
    }
    return resolved.events;
}

// data/event.rs:691-705
inline Prediction to_prediction(const Event& obs, double pred, const V& state) {
    Prediction p;
    p.time = obs.time; p.has_obs = obs.has_value; p.observation = obs.value; p.prediction = pred;
    p.outeq = (size_t)obs.index; p.has_poly = obs.has_poly; p.errorpoly = obs.poly; p.state = state;
    p.occasion = obs.occasion; p.censoring = obs.cens;
    return p;
}

// =============================================================================================
// Analytical kernels — equation/analytical/*.rs.  `t` is the sub-interval length dt.
// =============================================================================================
// one_compartment_models.rs:12-19
inline V one_compartment(const V& x, const V& p, double t, const V& rateiv, const Covariates&) {
    V xout = x;
    double ke = p[0];
    xout[0] = x[0] * std::exp(-ke * t) + rateiv[0] / ke * (1.0 - std::exp(-ke * t));
    return xout;
}
// one_compartment_models.rs:32-44
inline V one_compartment_with_absorption(const V& x, const V& p, double t, const V& rateiv, const Covariates&) {
    V xout = x;
    double ka = p[0], ke = p[1];
    xout[0] = x[0] * std::exp(-ka * t);
    xout[1] = x[1] * std::exp(-ke * t) + rateiv[0] / ke * (1.0 - std::exp(-ke * t))
            + ((ka * x[0]) / (ka - ke)) * (std::exp(-ke * t) - std::exp(-ka * t));
    return xout;
}
// two_compartment_models.rs:14-48
inline V two_compartments(const V& x, const V& p, double t, const V& rateiv, const Covariates&) {
    double ke = p[0], kcp = p[1], kpc = p[2];
    double s0 = (ke + kcp + kpc);
    double sq = s0 * s0 - 4.0 * ke * kpc;
    if (sq < 0.0) throw Error(ImaginaryRoots, "Imaginary solutions, program stopped!");
    sq = std::sqrt(sq);
    double l1 = (ke + kcp + kpc + sq) / 2.0;
    double l2 = (ke + kcp + kpc - sq) / 2.0;
    double e1 = std::exp(-l1 * t), e2 = std::exp(-l2 * t);
    double m11 = (l1 - kpc) * e1 + (kpc - l2) * e2;
    double m12 = -kpc * e1 + kpc * e2;
    double m21 = -kcp * e1 + kcp * e2;
    double m22 = (l1 - ke - kcp) * e1 + (ke + kcp - l2) * e2;
    double nz0 = (m11 * x[0] + m12 * x[1]) / (l1 - l2);
    double nz1 = (m21 * x[0] + m22 * x[1]) / (l1 - l2);
    double iv0 = ((l1 - kpc) / l1) * (1.0 - e1) + ((kpc - l2) / l2) * (1.0 - e2);
    double iv1 = (-kcp / l1) * (1.0 - e1) + (kcp / l2) * (1.0 - e2);
    double f = rateiv[0] / (l1 - l2);
    return V{nz0 + iv0 * f, nz1 + iv1 * f};
}
// two_compartment_models.rs:61-112
inline V two_compartments_with_absorption(const V& x, const V& p, double t, const V& rateiv, const Covariates&) {
    double ke = p[0], ka = p[1], kcp = p[2], kpc = p[3];
    V xout = x;
    double s0 = (ke + kcp + kpc);
    double sq = s0 * s0 - 4.0 * ke * kpc;
    if (sq < 0.0) throw Error(ImaginaryRoots, "Imaginary solutions, program stopped!");
    sq = std::sqrt(sq);
    double l1 = (ke + kcp + kpc + sq) / 2.0;
    double l2 = (ke + kcp + kpc - sq) / 2.0;
    double e1 = std::exp(-l1 * t), e2 = std::exp(-l2 * t);
    double m11 = (l1 - kpc) * e1 + (kpc - l2) * e2;
    double m12 = -kpc * e1 + kpc * e2;
    double m21 = -kcp * e1 + kcp * e2;
    double m22 = (l1 - ke - kcp) * e1 + (ke + kcp - l2) * e2;
    double nz0 = (m11 * x[1] + m12 * x[2]) / (l1 - l2);
    double nz1 = (m21 * x[1] + m22 * x[2]) / (l1 - l2);
    double iv0 = ((l1 - kpc) / l1) * (1.0 - e1) + ((kpc - l2) / l2) * (1.0 - e2);
    double iv1 = (-kcp / l1) * (1.0 - e1) + (kcp / l2) * (1.0 - e2);
    double f = rateiv[0] / (l1 - l2);
    double ea = std::exp(-ka * t);
    double ab0 = ((l1 - kpc) / (ka - l1)) * (e1 - ea) + ((kpc - l2) / (ka - l2)) * (e2 - ea);
    double ab1 = (-kcp / (ka - l1)) * (e1 - ea) + (kcp / (ka - l2)) * (e2 - ea);
    double g = ka * x[0] / (l1 - l2);
    xout[0] = x[0] * ea;
    xout[1] = (nz0 + iv0 * f) + ab0 * g;
    xout[2] = (nz1 + iv1 * f) + ab1 * g;
    return xout;
}

// Shared root/coefficient block of three_compartment_models.rs:23-77 / 139-193.
struct ThreeCptCoeffs {
    double l1, l2, l3, e1, e2, e3;
    double c[28];   // c[1]..c[27]
};
inline ThreeCptCoeffs three_cpt_coeffs(double k10, double k12, double k13, double k21, double k31, double t) {
    ThreeCptCoeffs r;
    double a = k10 + k12 + k13 + k21 + k31;
    double b = k10 * k21 + k13 * k21 + k10 * k31 + k12 * k31 + k21 * k31;
    double c = k10 * k21 * k31;
    double m = (3.0 * b - a * a) / 3.0;
    double n = (2.0 * (a * a * a) - 9.0 * a * b + 27.0 * c) / 27.0;
    double q = (n * n) / 4.0 + (m * m * m) / 27.0;
    if (q > 0.0) throw Error(ImaginaryRoots, "Imaginary solutions, program stopped!");
    double alpha = std::sqrt(-q);
    double beta = -n / 2.0;
    double gamma = std::sqrt(beta * beta + alpha * alpha);
    double theta = std::atan2(alpha, beta);
    double g3 = std::pow(gamma, 1.0 / 3.0);
    double s3 = std::sqrt(3.0);
    double l1 = a / 3.0 + g3 * (std::cos(theta / 3.0) + s3 * std::sin(theta / 3.0));
    double l2 = a / 3.0 + g3 * (std::cos(theta / 3.0) - s3 * std::sin(theta / 3.0));
    double l3 = a / 3.0 - (2.0 * g3 * std::cos(theta / 3.0));
    r.l1 = l1; r.l2 = l2; r.l3 = l3;
    r.e1 = std::exp(-(l1 * t)); r.e2 = std::exp(-(l2 * t)); r.e3 = std::exp(-(l3 * t));
    double d1 = (l2 - l1) * (l3 - l1), d2 = (l1 - l2) * (l3 - l2), d3 = (l1 - l3) * (l2 - l3);
    double* C = r.c;
    C[1] = (k21 - l1) * (k31 - l1) / d1;  C[2] = (k21 - l2) * (k31 - l2) / d2;  C[3] = (k21 - l3) * (k31 - l3) / d3;
    C[4] = k21 * (k31 - l1) / d1;         C[5] = k21 * (k31 - l2) / d2;         C[6] = k21 * (k31 - l3) / d3;
    C[7] = k31 * (k21 - l1) / d1;         C[8] = k31 * (k21 - l2) / d2;         C[9] = k31 * (k21 - l3) / d3;
    C[10] = k12 * (k31 - l1) / d1;        C[11] = k12 * (k31 - l2) / d2;        C[12] = k12 * (k31 - l3) / d3;
    C[13] = ((k10 + k12 + k13 - l1) * (k31 - l1) - (k13 * k31)) / d1;
    C[14] = ((k10 + k12 + k13 - l2) * (k31 - l2) - (k13 * k31)) / d2;
    C[15] = ((k10 + k12 + k13 - l3) * (k31 - l3) - (k13 * k31)) / d3;
    C[16] = k12 * k31 / d1;               C[17] = k12 * k31 / d2;               C[18] = k12 * k31 / d3;
    C[19] = k13 * (k21 - l1) / d1;        C[20] = k13 * (k21 - l2) / d2;        C[21] = k13 * (k21 - l3) / d3;
    C[22] = k21 * k13 / d1;               C[23] = k21 * k13 / d2;               C[24] = k21 * k13 / d3;
    C[25] = ((k10 + k12 + k13 - l1) * (k21 - l1) - (k12 * k21)) / d1;
    C[26] = ((k10 + k12 + k13 - l2) * (k21 - l2) - (k12 * k21)) / d2;
    C[27] = ((k10 + k12 + k13 - l3) * (k21 - l3) - (k12 * k21)) / d3;
    return r;
}
// nalgebra Matrix3 * Vector3: column-major axpy accumulation => (m_i0*v0 + m_i1*v1) + m_i2*v2
inline void three_cpt_apply(const ThreeCptCoeffs& r, double x1, double x2, double x3, double rate, double out[3]) {
    const double* C = r.c; double e1 = r.e1, e2 = r.e2, e3 = r.e3;
    double m[9] = {
        C[1] * e1 + C[2] * e2 + C[3] * e3,    C[4] * e1 + C[5] * e2 + C[6] * e3,    C[7] * e1 + C[8] * e2 + C[9] * e3,
        C[10] * e1 + C[11] * e2 + C[12] * e3, C[13] * e1 + C[14] * e2 + C[15] * e3, C[16] * e1 + C[17] * e2 + C[18] * e3,
        C[19] * e1 + C[20] * e2 + C[21] * e3, C[22] * e1 + C[23] * e2 + C[24] * e3, C[25] * e1 + C[26] * e2 + C[27] * e3};
    double iv[3] = {
        ((1.0 - e1) * C[1] / r.l1) + ((1.0 - e2) * C[2] / r.l2) + ((1.0 - e3) * C[3] / r.l3),
        ((1.0 - e1) * C[10] / r.l1) + ((1.0 - e2) * C[11] / r.l2) + ((1.0 - e3) * C[12] / r.l3),
        ((1.0 - e1) * C[19] / r.l1) + ((1.0 - e2) * C[20] / r.l2) + ((1.0 - e3) * C[21] / r.l3)};
    for (int i = 0; i < 3; ++i) {
        double nz = m[3 * i] * x1 + m[3 * i + 1] * x2 + m[3 * i + 2] * x3;
        out[i] = nz + iv[i] * rate;
    }
}
// three_compartment_models.rs:17-109
inline V three_compartments(const V& x, const V& p, double t, const V& rateiv, const Covariates&) {
    ThreeCptCoeffs r = three_cpt_coeffs(p[0], p[1], p[2], p[3], p[4], t);
    double o[3];
    three_cpt_apply(r, x[0], x[1], x[2], rateiv[0], o);
    return V{o[0], o[1], o[2]};
}
// three_compartment_models.rs:126-240
inline V three_compartments_with_absorption(const V& x, const V& p, double t, const V& rateiv, const Covariates&) {
    double ka = p[0];
    ThreeCptCoeffs r = three_cpt_coeffs(p[1], p[2], p[3], p[4], p[5], t);
    V xout = x;
    double o[3];
    three_cpt_apply(r, x[1], x[2], x[3], rateiv[0], o);
    double ea = std::exp(-ka * t);
    const double* C = r.c;
    double ab[3] = {
        (r.e1 - ea) * C[1] / (ka - r.l1) + (r.e2 - ea) * C[2] / (ka - r.l2) + (r.e3 - ea) * C[3] / (ka - r.l3),
        (r.e1 - ea) * C[10] / (ka - r.l1) + (r.e2 - ea) * C[11] / (ka - r.l2) + (r.e3 - ea) * C[12] / (ka - r.l3),
        (r.e1 - ea) * C[19] / (ka - r.l1) + (r.e2 - ea) * C[20] / (ka - r.l2) + (r.e3 - ea) * C[21] / (ka - r.l3)};
    xout[0] = x[0] * ea;
    for (int i = 0; i < 3; ++i) xout[i + 1] = o[i] + ab[i] * ka * x[0];
    return xout;
}
// *_cl_models.rs converters
inline V one_compartment_cl(const V& x, const V& p, double t, const V& r, const Covariates& c) {
    return one_compartment(x, V{p[0] / p[1]}, t, r, c);                        // one_compartment_cl_models.rs:16-22
}
inline V one_compartment_cl_with_absorption(const V& x, const V& p, double t, const V& r, const Covariates& c) {
    return one_compartment_with_absorption(x, V{p[0], p[1] / p[2]}, t, r, c);  // :39-46
}
inline V two_compartments_cl(const V& x, const V& p, double t, const V& r, const Covariates& c) {
    double cl = p[0], q = p[1], vc = p[2], vp = p[3];                          // two_compartment_cl_models.rs:16-26
    return two_compartments(x, V{cl / vc, q / vc, q / vp}, t, r, c);
}
inline V two_compartments_cl_with_absorption(const V& x, const V& p, double t, const V& r, const Covariates& c) {
    double ka = p[0], cl = p[1], q = p[2], vc = p[3], vp = p[4];               // :43-54
    return two_compartments_with_absorption(x, V{cl / vc, ka, q / vc, q / vp}, t, r, c);
}
inline V three_compartments_cl(const V& x, const V& p, double t, const V& r, const Covariates& c) {
    double cl = p[0], q2 = p[1], q3 = p[2], vc = p[3], v2 = p[4], v3 = p[5];   // three_compartment_cl_models.rs:16-30
    return three_compartments(x, V{cl / vc, q2 / vc, q3 / vc, q2 / v2, q3 / v3}, t, r, c);
}
inline V three_compartments_cl_with_absorption(const V& x, const V& p, double t, const V& r, const Covariates& c) {
    double ka = p[0], cl = p[1], q2 = p[2], q3 = p[3], vc = p[4], v2 = p[5], v3 = p[6];   // :47-68
    return three_compartments_with_absorption(x, V{ka, cl / vc, q2 / vc, q3 / vc, q2 / v2, q3 / v3}, t, r, c);
}
// pm_* wrappers: analytical/mod.rs:61-90 (drop index 0, run, pad a leading 0)
inline V wrap_pmetrics_analytical(const V& x, const V& p, double t, const V& rateiv, const Covariates& cov,
                                  const AnalyticalEq& native) {
    V cx(x.begin() + (x.empty() ? 0 : 1), x.end());
    V cr(rateiv.begin() + (rateiv.empty() ? 0 : 1), rateiv.end());
    V o = native(cx, p, t, cr, cov);
    V padded; padded.push_back(0.0); padded.insert(padded.end(), o.begin(), o.end());
    return padded;
}

inline AnalyticalEq analytical_kernel_by_name(const std::string& n) {
    if (n == "one_compartment") return one_compartment;
    if (n == "one_compartment_with_absorption") return one_compartment_with_absorption;
    if (n == "two_compartments") return two_compartments;
    if (n == "two_compartments_with_absorption") return two_compartments_with_absorption;
    if (n == "three_compartments") return three_compartments;
    if (n == "three_compartments_with_absorption") return three_compartments_with_absorption;
    if (n == "one_compartment_cl") return one_compartment_cl;
    if (n == "one_compartment_cl_with_absorption") return one_compartment_cl_with_absorption;
    if (n == "two_compartments_cl") return two_compartments_cl;
    if (n == "two_compartments_cl_with_absorption") return two_compartments_cl_with_absorption;
    if (n == "three_compartments_cl") return three_compartments_cl;
    if (n == "three_compartments_cl_with_absorption") return three_compartments_cl_with_absorption;
    throw Error(OtherError, "unknown analytical kernel " + n);
}

// =============================================================================================
// Analytical equation — equation/analytical/mod.rs:299-426
// =============================================================================================
// analytical/mod.rs:299-370
inline void analytical_solve(const Model& m, V& x, const V& parameters, const Covariates& cov,
                             const std::vector<Event>& infusions, double ti, double tf) {
    if (ti == tf) return;
    std::vector<double> ts{ti, tf};
    for (const auto& inf : infusions) {
        double t0 = inf.time, t1 = t0 + inf.duration;
        if (t0 > ti && t0 < tf) ts.push_back(t0);
        if (t1 > ti && t1 < tf) ts.push_back(t1);
    }
    std::stable_sort(ts.begin(), ts.end());
    {   // Vec::dedup_by(|a, b| (a - b).abs() < 1e-12): drop `a` when close to the retained `b`
        std::vector<double> d;
        for (double v : ts) if (d.empty() || !(std::fabs(v - d.back()) < 1e-12)) d.push_back(v);
        ts.swap(d);
    }
    double current_t = ts[0];
    V parameters_v = parameters;
    V rateiv((size_t)m.ndrugs, 0.0);
    for (size_t k = 1; k < ts.size(); ++k) {
        double next_t = ts[k];
        std::fill(rateiv.begin(), rateiv.end(), 0.0);
        for (const auto& inf : infusions) {
            double s = inf.time, e = s + inf.duration;
            if (current_t >= s && next_t <= e) {
                if (inf.index < 0) throw Error(UnknownInputLabel, "unknown input label");
                if (inf.index >= m.ndrugs) throw Error(InputOutOfRange, "InputOutOfRange");
                rateiv[(size_t)inf.index] += inf.amount / inf.duration;
            }
        }
        if (m.seq_eq) m.seq_eq(parameters_v, next_t, cov);
        double dt = next_t - current_t;
        x = m.eq(x, parameters_v, dt, rateiv, cov);
        current_t = next_t;
    }
}

// =============================================================================================
// Explicit Runge-Kutta steppers for the ODE family.
// THIRD-PARTY BOUNDARY: the reference delegates stepping to diffsol =0.16.1 (Cargo.toml:57),
// absent from /root/reference.  Restated here from the published methods:
//   Tsit45  — Ch. Tsitouras, "Runge-Kutta pairs of order 5(4) satisfying only the first column
//             simplifying assumption", Comput. Math. Appl. 62 (2011) 770-775.
//   Dopri5  — Dormand & Prince, J. Comput. Appl. Math. 6 (1980) 19-26 (Hairer-Norsett-Wanner
//             DOPRI5 controller: safety 0.9, growth in [0.2, 10]).
// Parity at this boundary is tolerance-based ("parity unpinned").
// =============================================================================================
struct RkTableau {
    int stages;            // including the FSAL stage
    double c[7];
    double a[7][7];
    double b[7];           // propagating weights (= last row for FSAL pairs)
    double e[7];           // error weights (b - bhat)
};

inline const RkTableau& tableau_dopri5() {
    static RkTableau t = [] {
        RkTableau r{}; r.stages = 7;
        double c[7] = {0, 1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1, 1};
        for (int i = 0; i < 7; ++i) r.c[i] = c[i];
        r.a[1][0] = 1.0 / 5;
        r.a[2][0] = 3.0 / 40; r.a[2][1] = 9.0 / 40;
        r.a[3][0] = 44.0 / 45; r.a[3][1] = -56.0 / 15; r.a[3][2] = 32.0 / 9;
        r.a[4][0] = 19372.0 / 6561; r.a[4][1] = -25360.0 / 2187; r.a[4][2] = 64448.0 / 6561; r.a[4][3] = -212.0 / 729;
        r.a[5][0] = 9017.0 / 3168; r.a[5][1] = -355.0 / 33; r.a[5][2] = 46732.0 / 5247; r.a[5][3] = 49.0 / 176; r.a[5][4] = -5103.0 / 18656;
        r.a[6][0] = 35.0 / 384; r.a[6][1] = 0; r.a[6][2] = 500.0 / 1113; r.a[6][3] = 125.0 / 192; r.a[6][4] = -2187.0 / 6784; r.a[6][5] = 11.0 / 84;
        for (int j = 0; j < 7; ++j) r.b[j] = r.a[6][j];
        double bh[7] = {5179.0 / 57600, 0, 7571.0 / 16695, 393.0 / 640, -92097.0 / 339200, 187.0 / 2100, 1.0 / 40};
        for (int j = 0; j < 7; ++j) r.e[j] = r.b[j] - bh[j];
        return r;
    }();
    return t;
}
inline const RkTableau& tableau_tsit45() {
    static RkTableau t = [] {
        RkTableau r{}; r.stages = 7;
        double c[7] = {0, 0.161, 0.327, 0.9, 0.9800255409045097, 1, 1};
        for (int i = 0; i < 7; ++i) r.c[i] = c[i];
        r.a[1][0] = 0.161;
        r.a[2][0] = -0.008480655492356989; r.a[2][1] = 0.335480655492357;
        r.a[3][0] = 2.8971530571054935; r.a[3][1] = -6.359448489975075; r.a[3][2] = 4.3622954328695815;
        r.a[4][0] = 5.325864828439257; r.a[4][1] = -11.748883564062828; r.a[4][2] = 7.4955393428898365; r.a[4][3] = -0.09249506636175525;
        r.a[5][0] = 5.86145544294642; r.a[5][1] = -12.92096931784711; r.a[5][2] = 8.159367898576159; r.a[5][3] = -0.071584973281401; r.a[5][4] = -0.028269050394068383;
        r.a[6][0] = 0.09646076681806523; r.a[6][1] = 0.01; r.a[6][2] = 0.4798896504144996; r.a[6][3] = 1.379008574103742; r.a[6][4] = -3.290069515436081; r.a[6][5] = 2.324710524099774;
        for (int j = 0; j < 7; ++j) r.b[j] = r.a[6][j];
        double e[7] = {-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995, -0.1447110071732629,
                       0.5823571654525552, -0.45808210592918697, 0.015151515151515152};
        for (int j = 0; j < 7; ++j) r.e[j] = e[j];
        return r;
    }();
    return t;
}

// Minimal adaptive ERK "solver" exposing the operations ode/mod.rs uses on a diffsol solver:
// state (t, y), set_stop_time + step-until-TstopReached, and a restart hook.
struct ErkSolver {
    const RkTableau* tab;
    std::function<void(double, const V&, V&)> rhs;
    double rtol, atol;
    double t = 0.0;
    V y;
    double h;             // current trial step (h0 = 1e-3, ode/mod.rs:349)
    bool have_k1 = false;
    V k[7];
    long nsteps = 0, nrej = 0, nrhs = 0;
    long max_steps = 5'000'000;

    ErkSolver(const RkTableau* tb, std::function<void(double, const V&, V&)> f, double rt, double at, double t0, V y0)
        : tab(tb), rhs(std::move(f)), rtol(rt), atol(at), t(t0), y(std::move(y0)), h(1e-3) {
        for (auto& kk : k) kk.assign(y.size(), 0.0);
    }
    void restart() { have_k1 = false; }      // reinitialize_at_boundary (ode/mod.rs:568-586)

    // integrate to exactly `tstop` (set_stop_time + step() loop of ode/mod.rs:741-760)
    void integrate_to(double tstop) {
        const size_t n = y.size();
        V ynew(n), ytmp(n);
        while (t < tstop) {
            if (++nsteps > max_steps) throw Error(SolverFailure, "max steps exceeded");
            double hh = std::min(h, tstop - t);
            bool last = (hh >= tstop - t);
            if (!have_k1) { rhs(t, y, k[0]); ++nrhs; have_k1 = true; }
            for (int s = 1; s < tab->stages; ++s) {
                for (size_t i = 0; i < n; ++i) {
                    double acc = 0.0;
                    for (int j = 0; j < s; ++j) acc += tab->a[s][j] * k[j][i];
                    ytmp[i] = y[i] + hh * acc;
                }
                rhs(t + tab->c[s] * hh, ytmp, k[s]); ++nrhs;
            }
            // ytmp now holds the 5th-order solution (FSAL: last stage evaluated at y_{n+1})
            ynew = ytmp;
            double err2 = 0.0;
            for (size_t i = 0; i < n; ++i) {
                double e = 0.0;
                for (int j = 0; j < tab->stages; ++j) e += tab->e[j] * k[j][i];
                e *= hh;
                double sc = atol + rtol * std::max(std::fabs(y[i]), std::fabs(ynew[i]));
                err2 += (e / sc) * (e / sc);
            }
            double err = std::sqrt(err2 / (double)std::max<size_t>(n, 1));
            if (!std::isfinite(err)) throw Error(SolverFailure, "non-finite error estimate");
            double fac = (err == 0.0) ? 10.0 : 0.9 * std::pow(err, -0.2);
            fac = std::min(10.0, std::max(0.2, fac));
            if (err <= 1.0) {
                t = last ? tstop : t + hh;
                y = ynew;
                k[0] = k[tab->stages - 1];   // FSAL
                if (!last || fac < 1.0) h = hh * fac; else h = std::max(h, hh * fac);
            } else {
                ++nrej;
                h = hh * std::min(1.0, fac);
                if (h < 1e-14 * std::max(1.0, std::fabs(t))) throw Error(SolverFailure, "step size underflow");
            }
        }
    }
};

// ode/closure.rs:16-99 InfusionTrack + :103-195 InfusionSchedule
struct InfusionTrack {
    int input;
    std::vector<double> event_times, cumulative_rates;
    double rate_at_left(double time) const {
        if (event_times.empty()) return 0.0;
        // binary_search semantic: find any equal element, walk left over equals
        auto lo = std::lower_bound(event_times.begin(), event_times.end(), time);
        size_t idx = (size_t)(lo - event_times.begin());
        if (idx == 0) return 0.0;
        return cumulative_rates[idx - 1];
    }
    double rate_at_right(double time) const {
        if (event_times.empty()) return 0.0;
        auto hi = std::upper_bound(event_times.begin(), event_times.end(), time);
        size_t idx = (size_t)(hi - event_times.begin());
        if (idx == 0) return 0.0;
        return cumulative_rates[idx - 1];
    }
    double rate_at(double time, const std::optional<double>& left_time) const {
        if (left_time && *left_time == time) return rate_at_left(time);
        return rate_at_right(time);
    }
};
struct InfusionSchedule {
    std::vector<InfusionTrack> tracks;
    std::vector<double> boundary_times;
    std::optional<double> left_continuity_time;

    InfusionSchedule(int ndrugs, const std::vector<Event>& events) {
        if (ndrugs == 0) return;
        std::vector<std::vector<std::pair<double, double>>> per_input((size_t)ndrugs);
        bool saw = false;
        for (const auto& inf : events) {
            if (inf.kind != EventKind::Infusion) continue;
            saw = true;
            if (inf.duration <= 0.0) continue;                       // closure.rs:127-129
            if (inf.index < 0) throw Error(UnknownInputLabel, "unknown input label");
            if (inf.index >= ndrugs) throw Error(InputOutOfRange, "InputOutOfRange");
            double rate = inf.amount / inf.duration;
            double end = inf.time + inf.duration;
            per_input[(size_t)inf.index].push_back({inf.time, rate});
            per_input[(size_t)inf.index].push_back({end, -rate});
            boundary_times.push_back(inf.time);
            boundary_times.push_back(end);
        }
        std::sort(boundary_times.begin(), boundary_times.end());
        boundary_times.erase(std::unique(boundary_times.begin(), boundary_times.end()), boundary_times.end());
        if (!saw) return;
        for (int i = 0; i < ndrugs; ++i) {
            auto& ev = per_input[(size_t)i];
            if (ev.empty()) continue;
            std::stable_sort(ev.begin(), ev.end(), [](auto& a, auto& b) { return a.first < b.first; });
            InfusionTrack tr; tr.input = i;
            double cur = 0.0;
            for (auto& [time, delta] : ev) { cur += delta; tr.event_times.push_back(time); tr.cumulative_rates.push_back(cur); }
            tracks.push_back(std::move(tr));
        }
    }
    void fill_rate_vector(double time, V& rateiv) const {     // closure.rs:185-194
        std::fill(rateiv.begin(), rateiv.end(), 0.0);
        for (const auto& tr : tracks) {
            double r = tr.rate_at(time, left_continuity_time);
            if (r != 0.0) rateiv[(size_t)tr.input] = r;
        }
    }
};

// ode/mod.rs:601-604
inline bool stop_time_reached(double stop_time, double state_t) {
    double tol = std::numeric_limits<double>::epsilon() * std::max(std::fabs(state_t), 1.0) * 64.0;
    return std::fabs(stop_time - state_t) <= tol;
}

struct SolveStats { long nsteps = 0, nrej = 0, nrhs = 0; };

// initial_state — analytical/mod.rs:409-426, ode/mod.rs:536-549
inline V initial_state(const Model& m, const V& params, const Covariates& cov, int occasion_index) {
    V x((size_t)m.nstates, 0.0);
    if (occasion_index == 0 && m.init) m.init(params, 0.0, cov, x);
    return x;
}

// ode/mod.rs:306-461 + run_events :609-824
inline std::vector<Prediction> ode_simulate_subject(const Model& m, const Subject& subject, const V& params,
                                                    SolveStats* stats = nullptr) {
    std::vector<Prediction> output;
    const int nstates = m.nstates, ndrugs = m.ndrugs;
    V zero_bolus((size_t)ndrugs, 0.0), zero_rateiv((size_t)ndrugs, 0.0), bolus_v((size_t)ndrugs, 0.0);
    V with_b((size_t)nstates), without_b((size_t)nstates), y_out((size_t)m.nout);

    for (const auto& occasion : subject.occasions) {
        const Covariates& cov = occasion.covariates;
        std::vector<Event> events = resolve_occasion_events(m, occasion, params);
        InfusionSchedule sched(ndrugs, events);
        V rate_scratch((size_t)ndrugs, 0.0);
        auto rhs = [&](double t, const V& y, V& dy) {                    // PmRhs::call_inplace (closure.rs)
            sched.fill_rate_vector(t, rate_scratch);
            std::fill(dy.begin(), dy.end(), 0.0);
            m.diffeq(y, params, t, dy, zero_bolus, rate_scratch, cov);
        };
        ErkSolver solver(m.solver == OdeSolver::Tsit45 ? &tableau_tsit45() : &tableau_dopri5(), rhs, m.rtol, m.atol,
                         occasion.initial_time(), initial_state(m, params, cov, occasion.index));

        const auto& bt = sched.boundary_times;
        size_t cursor = 0;
        bool pending_reinit = false;
        for (size_t index = 0; index < events.size(); ++index) {
            const Event& ev = events[index];
            switch (ev.kind) {
                case EventKind::Bolus: {                                   // ode/mod.rs:644-688
                    if (ev.index < 0) throw Error(UnknownInputLabel, "unknown input label");
                    if (ev.index >= ndrugs) throw Error(InputOutOfRange, "InputOutOfRange");
                    std::fill(bolus_v.begin(), bolus_v.end(), 0.0);
                    bolus_v[(size_t)ev.index] = ev.amount;
                    std::fill(with_b.begin(), with_b.end(), 0.0);
                    std::fill(without_b.begin(), without_b.end(), 0.0);
                    m.diffeq(solver.y, params, ev.time, without_b, zero_bolus, zero_rateiv, cov);
                    m.diffeq(solver.y, params, ev.time, with_b, bolus_v, zero_rateiv, cov);
                    for (int i = 0; i < nstates; ++i) solver.y[(size_t)i] += with_b[(size_t)i] - without_b[(size_t)i];
                    pending_reinit = true;
                    break;
                }
                case EventKind::Infusion: break;
                case EventKind::Observation: {                             // ode/mod.rs:692-715
                    std::fill(y_out.begin(), y_out.end(), 0.0);
                    m.out(solver.y, params, ev.time, cov, y_out);
                    if (ev.index < 0) throw Error(UnknownOutputLabel, "unknown output label");
                    if ((size_t)ev.index >= y_out.size()) throw Error(OuteqOutOfRange, "OuteqOutOfRange");
                    output.push_back(to_prediction(ev, y_out[(size_t)ev.index], solver.y));
                    break;
                }
            }
            if (index + 1 < events.size()) {
                double next_event_time = events[index + 1].time;
                while (next_event_time > solver.t) {
                    while (cursor < bt.size() && bt[cursor] <= solver.t) ++cursor;
                    double stop_time = next_event_time; bool is_boundary = false;
                    if (cursor < bt.size() && bt[cursor] <= next_event_time) { stop_time = bt[cursor]; is_boundary = true; ++cursor; }
                    sched.left_continuity_time = is_boundary ? std::optional<double>(stop_time) : std::nullopt;
                    if (stop_time > solver.t) {
                        if (pending_reinit) { solver.restart(); pending_reinit = false; }
                        solver.integrate_to(stop_time);
                        sched.left_continuity_time.reset();
                        if (is_boundary) pending_reinit = true;
                    } else {
                        // StopTimeAtCurrentTime branch (ode/mod.rs:781-812)
                        sched.left_continuity_time.reset();
                        if (stop_time_reached(stop_time, solver.t)) {
                            if (is_boundary) pending_reinit = true;
                            if (stop_time < next_event_time) continue;
                            break;
                        }
                        throw Error(SolverFailure, "StopTimeAtCurrentTime");
                    }
                }
            }
        }
        if (stats) { stats->nsteps += solver.nsteps; stats->nrej += solver.nrej; stats->nrhs += solver.nrhs; }
    }
    return output;
}

// Generic loop for Analytical — equation/mod.rs:480-516 + simulate_event :300-358 +
// process_observation analytical/mod.rs:373-407.
inline std::vector<Prediction> analytical_simulate_subject(const Model& m, const Subject& subject, const V& params) {
    std::vector<Prediction> output;
    for (const auto& occasion : subject.occasions) {
        const Covariates& cov = occasion.covariates;
        V x = initial_state(m, params, cov, occasion.index);
        std::vector<Event> infusions;
        std::vector<Event> events = resolve_occasion_events(m, occasion, params);
        for (size_t index = 0; index < events.size(); ++index) {
            const Event& ev = events[index];
            switch (ev.kind) {
                case EventKind::Bolus:
                    if (ev.index < 0) throw Error(UnknownInputLabel, "unknown input label");
                    if (ev.index >= m.ndrugs) throw Error(InputOutOfRange, "InputOutOfRange");
                    x[(size_t)ev.index] += ev.amount;                       // State::add_bolus ode/mod.rs:268-273
                    break;
                case EventKind::Infusion: infusions.push_back(ev); break;
                case EventKind::Observation: {
                    V y((size_t)m.nout, 0.0);
                    m.out(x, params, ev.time, cov, y);
                    if (ev.index < 0) throw Error(UnknownOutputLabel, "unknown output label");
                    if ((size_t)ev.index >= y.size()) throw Error(OuteqOutOfRange, "OuteqOutOfRange");
                    output.push_back(to_prediction(ev, y[(size_t)ev.index], x));
                    break;
                }
            }
            if (index + 1 < events.size()) analytical_solve(m, x, params, cov, infusions, ev.time, events[index + 1].time);
        }
    }
    return output;
}

// =============================================================================================
// SDE — sde/em.rs (whole file) + sde/mod.rs:102-175, 491-661, 712-767
// THIRD-PARTY BOUNDARY: rand::rng() (thread-local, unseeded ChaCha) -> here a seeded
// std::mt19937_64 per (subject, support point); parity is statistical only.
// =============================================================================================
struct Rng {
    std::mt19937_64 gen;
    std::normal_distribution<double> normal{0.0, 1.0};
    std::uniform_real_distribution<double> unif{0.0, 1.0};
    explicit Rng(uint64_t seed) : gen(seed) {}
    double n() { return normal(gen); }
    double u() { return unif(gen); }
};

// em.rs:104-120
inline void euler_maruyama_step(const std::function<void(double, const V&, V&)>& drift,
                                const std::function<void(double, const V&, V&)>& diffusion,
                                double time, double dt, V& state, Rng& rng) {
    size_t n = state.size();
    V drift_term(n, 0.0), diffusion_term(n, 0.0);
    drift(time, state, drift_term);
    diffusion(time, state, diffusion_term);
    for (size_t i = 0; i < n; ++i) state[i] += drift_term[i] * dt + diffusion_term[i] * rng.n() * std::sqrt(dt);
}
// em.rs:134-167 (EM::solve) with rtol = atol = 1e-2 (sde/mod.rs:172), max_step 0.1, min_step 1e-6
inline V em_solve(const std::function<void(double, const V&, V&)>& drift,
                  const std::function<void(double, const V&, V&)>& diffusion, V state, double t0, double tf, Rng& rng) {
    const double rtol = 1e-2, atol = 1e-2, max_step = 0.1, min_step = 1e-6, safety = 0.9;
    double t = t0, dt = max_step;
    auto new_step = [&](double dt_, double error) {
        double nd = dt_ * safety * std::pow(1.0 / error, 0.5);
        return std::min(std::max(nd, min_step), max_step);      // f64::clamp
    };
    while (t < tf) {
        V y1 = state, y2 = state;
        euler_maruyama_step(drift, diffusion, t, dt, y1, rng);
        euler_maruyama_step(drift, diffusion, t, dt / 2.0, y2, rng);
        euler_maruyama_step(drift, diffusion, t + dt / 2.0, dt / 2.0, y2, rng);
        double err = 0.0;
        for (size_t i = 0; i < state.size(); ++i) {
            double tol = atol + rtol * std::fabs(state[i]);
            err = std::max(err, std::fabs(y1[i] - y2[i]) / tol);
        }
        if (err <= 1.0) {
            t += dt;
            state = y2;
            dt = new_step(dt, err);
            dt = std::min(dt, tf - t);
        } else {
            dt = new_step(dt, err);
        }
    }
    return state;
}

// sde/mod.rs:102-155
inline V simulate_sde_event(const Model& m, const V& x, const V& params, const Covariates& cov,
                            const std::vector<Event>& infusions, double ti, double tf, Rng& rng) {
    if (ti == tf) return x;
    auto drift = [&](double time, const V& state, V& out) {
        V rateiv((size_t)m.ndrugs, 0.0);
        for (const auto& inf : infusions)
            if (time >= inf.time && time <= inf.duration + inf.time) rateiv[(size_t)inf.index] += inf.amount / inf.duration;
        V o(state.size(), 0.0);
        m.drift(state, params, time, o, rateiv, cov);
        out = o;
    };
    auto diffusion = [&](double, const V&, V& out) {
        V o(out.size(), 0.0);
        m.diffusion(params, o);
        out = o;
    };
    return em_solve(drift, diffusion, x, ti, tf, rng);
}

// sde/mod.rs:747-767
inline std::vector<size_t> sysresample(const V& q, Rng& rng) {
    size_t m = q.size();
    V qc(m, 0.0);
    qc[0] = q[0];
    for (size_t i = 1; i < m; ++i) qc[i] = qc[i - 1] + q[i];
    V u(m);
    for (size_t i = 0; i < m; ++i) u[i] = ((double)i + rng.u()) / (double)m;
    std::vector<size_t> idx(m, 0);
    size_t k = 0;
    for (size_t j = 0; j < m; ++j) {
        while (k + 1 < m && qc[k] < u[j]) ++k;   // (reference indexes out of bounds if rounding leaves qc[m-1] < u)
        idx[j] = k;
    }
    return idx;
}

struct SdeResult {
    std::vector<Prediction> mean_predictions;   // Predictions::get_predictions (sde/mod.rs:394-419)
    bool has_likelihood = false;
    double likelihood = 1.0;                    // product of per-observation mean weights (equation/mod.rs:514)
};

// Generic loop specialised for SDE (equation/mod.rs:480-516 + sde/mod.rs:526-661).
// error_models == nullptr  -> "mean prediction" mode (what log_likelihood_matrix runs)
// error_models != nullptr  -> particle filter (weights, resampling; SDE::estimate_log_likelihood)
inline SdeResult sde_simulate_subject(const Model& m, const Subject& subject, const V& params,
                                      const AssayErrorModels* error_models, uint64_t seed) {
    Rng rng(seed);
    SdeResult res;
    const size_t N = (size_t)m.nparticles;
    std::vector<std::vector<double>> pred_cols;   // per observation: N predictions
    std::vector<Event> obs_events;
    std::vector<double> likelihood;
    for (const auto& occasion : subject.occasions) {
        const Covariates& cov = occasion.covariates;
        std::vector<V> x(N);
        for (size_t i = 0; i < N; ++i) x[i] = initial_state(m, params, cov, occasion.index);   // sde/mod.rs:579-599
        std::vector<Event> infusions;
        std::vector<Event> events = resolve_occasion_events(m, occasion, params);
        for (size_t index = 0; index < events.size(); ++index) {
            const Event& ev = events[index];
            switch (ev.kind) {
                case EventKind::Bolus: {                                            // sde/mod.rs:614-632
                    if (ev.index < 0) throw Error(UnknownInputLabel, "unknown input label");
                    if (ev.index >= m.ndrugs) throw Error(InputOutOfRange, "InputOutOfRange");
                    long dest = ev.index;
                    if ((size_t)ev.index < m.injected_bolus_destination.size() && m.injected_bolus_destination[(size_t)ev.index] >= 0)
                        dest = m.injected_bolus_destination[(size_t)ev.index];
                    for (auto& p : x) p[(size_t)dest] += ev.amount;
                    break;
                }
                case EventKind::Infusion: infusions.push_back(ev); break;
                case EventKind::Observation: {                                      // sde/mod.rs:526-577
                    std::vector<Prediction> pred(N);
                    std::vector<double> col(N);
                    for (size_t i = 0; i < N; ++i) {
                        V y((size_t)m.nout, 0.0);
                        m.out(x[i], params, ev.time, cov, y);
                        pred[i] = to_prediction(ev, y[(size_t)ev.index], x[i]);
                        col[i] = pred[i].prediction;
                    }
                    pred_cols.push_back(col);
                    obs_events.push_back(ev);
                    if (error_models) {
                        V q; q.reserve(N);
                        for (auto& p : pred) q.push_back(std::exp(prediction_log_likelihood(p, *error_models)));
                        double sum_q = 0.0; for (double v : q) sum_q += v;
                        V w(N); for (size_t i = 0; i < N; ++i) w[i] = q[i] / sum_q;
                        auto idx = sysresample(w, rng);
                        std::vector<V> a(N);
                        for (size_t i = 0; i < N; ++i) a[i] = x[idx[i]];
                        x = a;
                        likelihood.push_back(sum_q / (double)N);
                    }
                    break;
                }
            }
            if (index + 1 < events.size()) {
                double ti = ev.time, tf = events[index + 1].time;
                for (auto& p : x) p = simulate_sde_event(m, p, params, cov, infusions, ti, tf, rng);   // sde/mod.rs:491-517
            }
        }
    }
    for (size_t c = 0; c < pred_cols.size(); ++c) {
        double s = 0.0; for (double v : pred_cols[c]) s += v;
        Prediction p = to_prediction(obs_events[c], s / (double)N, V{});
        res.mean_predictions.push_back(p);
    }
    if (error_models) {
        res.has_likelihood = true;
        double prod = 1.0; for (double v : likelihood) prod *= v;
        res.likelihood = prod;
    }
    return res;
}

// =============================================================================================
// Per-pair entry points and the psi matrix driver
// =============================================================================================
// Equation::estimate_predictions_dense (equation/mod.rs:459-466; ODE override ode/mod.rs:859-866)
inline std::vector<Prediction> estimate_predictions(const Model& m, const Subject& s, const V& params,
                                                    uint64_t seed = 0, SolveStats* stats = nullptr) {
    switch (m.kind) {
        case EqnKind::Analytical: return analytical_simulate_subject(m, s, params);
        case EqnKind::ODE:        return ode_simulate_subject(m, s, params, stats);
        default:                  return sde_simulate_subject(m, s, params, nullptr, seed).mean_predictions;
    }
}
// Equation::estimate_log_likelihood_dense (equation/mod.rs:468-477): predictions -> Σ log-lik.
// For SDE this is the mean-prediction likelihood (sde/mod.rs:394-432), NOT the particle filter.
inline double estimate_log_likelihood_dense(const Model& m, const Subject& s, const V& params,
                                            const AssayErrorModels& em, uint64_t seed = 0, SolveStats* stats = nullptr) {
    return subject_log_likelihood(estimate_predictions(m, s, params, seed, stats), em);
}
// SDE::estimate_log_likelihood (sde/mod.rs:689-704, 712-736): particle filter, ln(Π mean weights)
inline double sde_particle_filter_log_likelihood(const Model& m, const Subject& s, const V& params,
                                                 const AssayErrorModels& em, uint64_t seed) {
    SdeResult r = sde_simulate_subject(m, s, params, &em, seed);
    return r.likelihood > 0.0 ? std::log(r.likelihood) : -std::numeric_limits<double>::infinity();
}

}  // namespace orc
