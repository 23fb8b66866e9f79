// data.hpp — host-side mirror of pharmsol's data types on the psi path and the SoA flattener.
//
// Mirrors (file:line relative to /root/reference/src/data):
//   Subject / Occasion / Event{Bolus,Infusion,Observation} / Censor   structs.rs:352,556; event.rs:107-575
//   SubjectBuilder (bolus, infusion, observation, ..., repeat, reset)   builder.rs:84-362
//   Covariate / Covariates (piecewise-linear + carry-forward segments)  covariate.rs:26-241
//   ErrorPoly / AssayErrorModel / AssayErrorModels (sigma from the OBSERVATION)  error_model.rs:17,150,677,786,1045
// and the label -> dense index resolution of equation/metadata.rs:236-275, dsl/native.rs:663-770,
// which happens ONCE here at flatten time instead of per (subject, support point) pair.
#pragma once
#include <cstdint>
#include <map>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../device/psi_types.h"

namespace pharmsol {

struct PharmsolError : std::runtime_error {
    int code;
    PharmsolError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

struct ErrorPoly { double c0 = 0, c1 = 0, c2 = 0, c3 = 0; };
enum class Censor : int { None = 0, BLOQ = 1, ALOQ = 2 };
enum class EventKind : int { Observation = 0, Bolus = 1, Infusion = 2 };

struct Event {
    EventKind kind = EventKind::Observation;
    double time = 0.0;
    double amount = 0.0;      // dose amount
    double duration = 0.0;    // infusion
    std::string label;        // InputLabel / OutputLabel
    bool has_value = false;   // observation: Some(value)
    double value = 0.0;
    bool has_poly = false;
    ErrorPoly poly;
    Censor cens = Censor::None;
    int occasion = 0;
};

struct Covariate {
    std::string name;
    std::vector<std::pair<double, double>> observations;
    bool fixed = false;       // Pmetrics `!` suffix / set_fixed: carry-forward everywhere
    void add_observation(double t, double v);
};

struct Occasion {
    std::vector<Event> events;
    std::map<std::string, Covariate> covariates;
    int index = 0;
    void sort();
    void add_event(const Event& e) { events.push_back(e); sort(); }
    double initial_time() const;
};

struct Subject {
    std::string id;
    std::vector<Occasion> occasions;
};

struct SubjectBuilder {
    std::string id;
    std::vector<Occasion> occasions;
    Occasion current;
    std::map<std::string, Covariate> covariates;
    std::optional<Event> last;

    explicit SubjectBuilder(std::string i) : id(std::move(i)) {}
    SubjectBuilder& event(Event e);
    SubjectBuilder& bolus(double t, double amount, const std::string& input);
    SubjectBuilder& infusion(double t, double amount, const std::string& input, double duration);
    SubjectBuilder& observation(double t, double value, const std::string& outeq);
    SubjectBuilder& censored_observation(double t, double value, const std::string& outeq, Censor c);
    SubjectBuilder& missing_observation(double t, const std::string& outeq);
    SubjectBuilder& observation_with_error(double t, double value, const std::string& outeq, ErrorPoly p, Censor c);
    SubjectBuilder& repeat(size_t n, double delta);
    SubjectBuilder& reset();
    SubjectBuilder& covariate(const std::string& name, double t, double value);
    Subject build();
};

struct Data {
    std::vector<Subject> subjects;
};

// Pmetrics CSV -> Data (data/parser/pmetrics/{mod.rs, row.rs}); throws PharmsolError(ST_OTHER, message)
Data read_pmetrics_text(const std::string& text);
Data read_pmetrics_file(const std::string& path);
std::string describe_data_json(const Data& d);
Data expand_data(const Data& d, double idelta, double tad);      // Data::expand (data/structs.rs:155-260)

enum class ErrKind : int { None = 0, Additive = 1, Proportional = 2 };
struct AssayErrorModel {
    ErrKind kind = ErrKind::None;
    double factor = 0.0;      // lambda (additive) / gamma (proportional)
    ErrorPoly poly;
};
struct AssayErrorModels {
    std::vector<AssayErrorModel> models;   // dense, by output index
};

// What the flattener needs to know about the model (names for label resolution).
enum class RouteKind : int { Bolus = 0, Infusion = 1 };
struct RouteInfo {
    std::string name;
    bool has_kind = true;     // canonical `routes { oral -> depot }` declares no kind
    RouteKind kind = RouteKind::Bolus;
    int index = 0;            // dense input slot
    int destination = -1;     // destination state offset
    bool has_lag = false, has_bioavailability = false;   // dsl/model_info.rs route properties
    int declaration_index = 0;
    std::string destination_name;
    bool inject_input_to_destination = true;            // false when dynamics / drift read the route input themselves
};
struct ModelLabels {
    std::vector<RouteInfo> routes;
    std::vector<std::string> outputs;
    std::vector<std::string> covariates;
    int route_len = 0;
    int nout = 0;
};

// Host copy of the flattened population; `upload` puts it into one device allocation.
struct FlatPopulation {
    std::vector<int32_t> occ_offsets, occ_index, ev_offsets, bol_offsets, bol_event, inf_offsets, bnd_offsets, cov_offsets;
    std::vector<psi::EventRec> events;
    std::vector<psi::InfRec> infs;
    std::vector<double> bnds, occ_t0;
    std::vector<psi::CovSeg> cov_segs;
    std::vector<int32_t> obs_offsets;   // [nsub+1] prefix sums of observation counts (prediction rows)
    // timeline program (psi_types.h): built when no route of the model declares a lag, i.e. when event times do not
    // depend on the support point; the closed-form kernels execute it instead of walking events / boundaries / infusions
    std::vector<int32_t> prog_offsets;
    std::vector<psi::EventRec> prog;
    std::vector<double> prog_rates;
    bool has_prog = false;
    bool prog_cov = false;              // EV_STEP records carry the single covariate's value (w: at the end, sigma: at t = dt)
    int32_t nsub = 0, ncov = 0, max_events = 0;
    int64_t nobs_total = 0;
    bool has_lagged_candidates = false;
};

// Resolve labels, compute per-observation sigma constants, build every offset table.
// Throws PharmsolError(UnknownInputLabel / UnknownOutputLabel / UnsupportedInputRouteKind / ...)
// exactly where resolve_occasion_events / resolve_events would (before any simulation).
FlatPopulation flatten_population(const Data& data, const ModelLabels& labels, const AssayErrorModels* error_models);

// AssayErrorModel::sigma (error_model.rs:1045-1080).  Returns a psi::ST_* code (0 = ok).
int assay_sigma(const AssayErrorModels& em, int outeq, double obs, bool has_poly, const ErrorPoly& poly, double& sigma);

}  // namespace pharmsol
