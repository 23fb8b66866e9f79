// data.cpp — Subject builder, covariate segments, error-model sigma and the SoA flattener.
// See data.hpp for the reference lines each piece mirrors.
#include "data.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

namespace pharmsol {

using namespace psi;

// f64::total_cmp order (data/event.rs:301-303)
static inline int64_t total_key(double x) {
    int64_t b;
    std::memcpy(&b, &x, sizeof b);
    b ^= (int64_t)((uint64_t)(b >> 63) >> 1);
    return b;
}
static bool event_less(const Event& a, const Event& b) {
    const int64_t ka = total_key(a.time), kb = total_key(b.time);
    if (ka != kb) return ka < kb;
    return (int)a.kind < (int)b.kind;   // Observation < Bolus < Infusion
}

void Occasion::sort() { std::stable_sort(events.begin(), events.end(), event_less); }

double Occasion::initial_time() const {
    if (events.empty()) return 0.0;
    double m = events[0].time;
    for (const auto& e : events) if (e.time < m) m = e.time;
    return m;
}

void Covariate::add_observation(double t, double v) {
    for (auto& o : observations)
        if (o.first == t) { o.second = v; return; }
    observations.emplace_back(t, v);
}

SubjectBuilder& SubjectBuilder::event(Event e) {
    e.occasion = current.index;
    last = e;
    current.add_event(e);
    return *this;
}
SubjectBuilder& SubjectBuilder::bolus(double t, double amount, const std::string& input) {
    Event e; e.kind = EventKind::Bolus; e.time = t; e.amount = amount; e.label = input; return event(e);
}
SubjectBuilder& SubjectBuilder::infusion(double t, double amount, const std::string& input, double duration) {
    Event e; e.kind = EventKind::Infusion; e.time = t; e.amount = amount; e.duration = duration; e.label = input; return event(e);
}
SubjectBuilder& SubjectBuilder::observation(double t, double value, const std::string& outeq) {
    Event e; e.kind = EventKind::Observation; e.time = t; e.has_value = true; e.value = value; e.label = outeq; return event(e);
}
SubjectBuilder& SubjectBuilder::censored_observation(double t, double value, const std::string& outeq, Censor c) {
    Event e; e.kind = EventKind::Observation; e.time = t; e.has_value = true; e.value = value; e.label = outeq; e.cens = c; return event(e);
}
SubjectBuilder& SubjectBuilder::missing_observation(double t, const std::string& outeq) {
    Event e; e.kind = EventKind::Observation; e.time = t; e.label = outeq; return event(e);
}
SubjectBuilder& SubjectBuilder::observation_with_error(double t, double value, const std::string& outeq, ErrorPoly p, Censor c) {
    Event e; e.kind = EventKind::Observation; e.time = t; e.has_value = true; e.value = value; e.label = outeq;
    e.has_poly = true; e.poly = p; e.cens = c; return event(e);
}
SubjectBuilder& SubjectBuilder::repeat(size_t n, double delta) {
    if (!last) return *this;
    const Event proto = *last;
    for (size_t i = 1; i <= n; ++i) {
        Event e = proto;
        e.time = proto.time + delta * (double)i;
        event(e);
    }
    return *this;
}
SubjectBuilder& SubjectBuilder::reset() {
    const int next = current.index + 1;
    current.sort();
    current.covariates = covariates;
    occasions.push_back(current);
    current = Occasion();
    current.index = next;
    covariates.clear();
    last.reset();
    return *this;
}
SubjectBuilder& SubjectBuilder::covariate(const std::string& name, double t, double value) {
    auto it = covariates.find(name);
    if (it == covariates.end()) {
        Covariate c; c.name = name;
        it = covariates.emplace(name, c).first;
    }
    it->second.add_observation(t, value);
    return *this;
}
Subject SubjectBuilder::build() {
    reset();
    Subject s; s.id = id; s.occasions = occasions;
    return s;
}

int assay_sigma(const AssayErrorModels& em, int outeq, double obs, bool has_poly, const ErrorPoly& poly, double& sigma) {
    sigma = std::numeric_limits<double>::quiet_NaN();
    if (outeq < 0 || (size_t)outeq >= em.models.size()) return ST_INVALID_OUTPUT_EQUATION;   // error_model.rs:679-681
    const AssayErrorModel& m = em.models[(size_t)outeq];
    if (m.kind == ErrKind::None) return ST_NONE_ERROR_MODEL;                                  // :682-684
    const ErrorPoly ep = has_poly ? poly : m.poly;
    const double alpha = ep.c0 + ep.c1 * obs + ep.c2 * (obs * obs) + ep.c3 * (obs * obs * obs);
    if (m.kind == ErrKind::Additive) sigma = std::sqrt(alpha * alpha + m.factor * m.factor);
    else sigma = m.factor * alpha;
    if (sigma < 0.0) return ST_NEGATIVE_SIGMA;
    if (!std::isfinite(sigma)) return ST_NON_FINITE_SIGMA;
    return ST_OK;
}

static bool is_bare_numeric(const std::string& s) {
    if (s.empty()) return false;
    for (char c : s) if (c < '0' || c > '9') return false;
    return true;
}

// dsl/native.rs:663-770 resolution order: exact (name, kind); kind-less route of that name;
// `input_<n>` alias for bare numeric labels; never positional.
static int resolve_input(const ModelLabels& L, const std::string& label, RouteKind kind) {
    auto find = [&](bool want_kind, RouteKind k) -> const RouteInfo* {
        for (const auto& r : L.routes)
            if (r.has_kind == want_kind && (!want_kind || r.kind == k) && r.name == label) return &r;
        if (!is_bare_numeric(label)) return nullptr;
        const std::string alias = "input_" + label;
        for (const auto& r : L.routes)
            if (r.has_kind == want_kind && (!want_kind || r.kind == k) && r.name == alias) return &r;
        return nullptr;
    };
    const RouteInfo* r = find(true, kind);
    if (!r) r = find(false, kind);
    if (!r) {
        const RouteKind other = kind == RouteKind::Bolus ? RouteKind::Infusion : RouteKind::Bolus;
        if (find(true, other)) throw PharmsolError(ST_UNSUPPORTED_INPUT_ROUTE_KIND, "input `" + label + "` is declared with the other route kind");
        std::string avail;
        for (const auto& q : L.routes) avail += (avail.empty() ? "" : ", ") + q.name;
        throw PharmsolError(ST_UNKNOWN_INPUT_LABEL, "unknown input label `" + label + "` (available: " + avail + ")");
    }
    if (r->index >= L.route_len) throw PharmsolError(ST_INPUT_OUT_OF_RANGE, "input out of range");
    return r->index;
}
static int resolve_output(const ModelLabels& L, const std::string& label) {
    for (size_t i = 0; i < L.outputs.size(); ++i) if (L.outputs[i] == label) return (int)i;
    if (is_bare_numeric(label)) {
        const std::string alias = "outeq_" + label;
        for (size_t i = 0; i < L.outputs.size(); ++i) if (L.outputs[i] == alias) return (int)i;
    }
    std::string avail;
    for (const auto& q : L.outputs) avail += (avail.empty() ? "" : ", ") + q;
    throw PharmsolError(ST_UNKNOWN_OUTPUT_LABEL, "unknown output label `" + label + "` (available: " + avail + ")");
}

FlatPopulation flatten_population(const Data& data, const ModelLabels& labels, const AssayErrorModels* em) {
    FlatPopulation f;
    f.nsub = (int32_t)data.subjects.size();
    f.ncov = (int32_t)labels.covariates.size();
    f.occ_offsets.push_back(0);
    f.ev_offsets.push_back(0);
    f.bol_offsets.push_back(0);
    f.inf_offsets.push_back(0);
    f.bnd_offsets.push_back(0);
    f.cov_offsets.push_back(0);
    f.obs_offsets.push_back(0);
    const double inf = std::numeric_limits<double>::infinity();
    int32_t obs_row = 0;
    f.has_prog = true;
    for (const auto& r : labels.routes) if (r.has_lag) f.has_prog = false;      // lagged bolus times depend on the support point
    const int nroute = std::max(1, labels.route_len);
    f.prog_cov = f.has_prog && nroute == 1 && labels.covariates.size() == 1;
    f.prog_offsets.push_back(0);
    for (const auto& subj : data.subjects) {
        for (const auto& occ : subj.occasions) {
            const size_t ev_first = f.events.size(), inf_first = f.infs.size(), prog_first = f.prog.size(), seg_first = f.cov_segs.size();
            f.occ_index.push_back(occ.index);
            f.occ_t0.push_back(occ.initial_time());
            std::vector<double> bounds;
            for (const auto& ev : occ.events) {
                EventRec r{};
                r.time = ev.time;
                r.obs_row = -1;
                switch (ev.kind) {
                    case EventKind::Bolus: {
                        const int idx = resolve_input(labels, ev.label, RouteKind::Bolus);
                        r.a = ev.amount;
                        r.meta = ev_pack(EV_BOLUS, 0, 0, idx, 0);
                        f.bol_event.push_back((int32_t)f.events.size());
                        break;
                    }
                    case EventKind::Infusion: {
                        const int idx = resolve_input(labels, ev.label, RouteKind::Infusion);
                        r.a = ev.amount; r.b = ev.duration;
                        r.meta = ev_pack(EV_INFUSION, 0, 0, idx, 0);
                        InfRec ir{}; ir.time = ev.time; ir.duration = ev.duration; ir.rate = ev.amount / ev.duration; ir.input = idx;
                        f.infs.push_back(ir);
                        if (ev.duration > 0.0) { bounds.push_back(ev.time); bounds.push_back(ev.time + ev.duration); }   // closure.rs:127-144
                        break;
                    }
                    case EventKind::Observation: {
                        const int idx = resolve_output(labels, ev.label);
                        if (idx >= labels.nout) throw PharmsolError(ST_OUTEQ_OUT_OF_RANGE, "outeq out of range");
                        int status = ST_OK;
                        double sigma = std::numeric_limits<double>::quiet_NaN();
                        r.a = ev.has_value ? ev.value : std::numeric_limits<double>::quiet_NaN();
                        if (ev.has_value && em) {
                            status = assay_sigma(*em, idx, ev.value, ev.has_poly, ev.poly, sigma);
                            if (status == ST_OK) {
                                r.b = -0.5 * 1.8378770664093453 - std::log(sigma);   // distributions.rs:12, 33
                                r.w = 1.0 / (2.0 * sigma * sigma);
                                r.sigma = sigma;
                            }
                        } else if (ev.has_value && !em) {
                            status = ST_MISSING_ERROR_MODEL;
                        }
                        r.meta = ev_pack(EV_OBS, (int)ev.cens, ev.has_value ? 1 : 0, idx, status);
                        r.obs_row = obs_row++;
                        break;
                    }
                }
                f.events.push_back(r);
            }
            f.max_events = std::max<int32_t>(f.max_events, (int32_t)occ.events.size());
            std::sort(bounds.begin(), bounds.end());
            bounds.erase(std::unique(bounds.begin(), bounds.end()), bounds.end());
            f.bnds.insert(f.bnds.end(), bounds.begin(), bounds.end());
            if (f.has_prog) {
                // The walk psi_engine.cuh performs per pair (event -> next event, split at interior boundaries), once.
                size_t abc = 0;
                std::vector<double> rate((size_t)nroute);
                for (size_t k = ev_first; k < f.events.size(); ++k) {
                    const EventRec& e = f.events[k];
                    if (ev_kind(e.meta) != EV_INFUSION) f.prog.push_back(e);
                    if (k + 1 >= f.events.size()) break;
                    const double te = e.time, tn = f.events[k + 1].time;
                    if (te == tn) continue;
                    double last = te;
                    while (abc < bounds.size() && bounds[abc] <= te) ++abc;
                    while (true) {
                        while (abc < bounds.size()) {       // candidates the reference's 1e-12 dedup would remove
                            const double bq = bounds[abc];
                            if (bq <= last || std::fabs(bq - last) < 1e-12) ++abc; else break;
                        }
                        double nxt;
                        const double bq = abc < bounds.size() ? bounds[abc] : inf;
                        if (bq < tn) { nxt = bq; ++abc; }
                        else if (tn > last && !(std::fabs(tn - last) < 1e-12)) nxt = tn;
                        else break;
                        std::fill(rate.begin(), rate.end(), 0.0);
                        for (size_t q = inf_first; q < f.infs.size(); ++q) {      // interval_route_inputs, analytical/mod.rs:337-357
                            const InfRec& ir = f.infs[q];
                            if (last >= ir.time && nxt <= ir.time + ir.duration && ir.input >= 0 && ir.input < nroute) rate[(size_t)ir.input] += ir.rate;
                        }
                        EventRec s{};
                        s.time = nxt;
                        s.a = nxt - last;
                        s.b = rate[0];
                        s.w = nroute > 1 ? rate[1] : 0.0;
                        s.sigma = nroute > 2 ? rate[2] : 0.0;
                        s.meta = ev_pack(EV_STEP, 0, 0, 0, 0);
                        s.obs_row = -1;
                        if (nroute > 3) {
                            s.obs_row = (int32_t)(f.prog_rates.size() / (size_t)nroute);
                            f.prog_rates.insert(f.prog_rates.end(), rate.begin(), rate.end());
                        }
                        f.prog.push_back(s);
                        last = nxt;
                        if (nxt == tn) break;
                    }
                }
            }
            f.prog_offsets.push_back((int32_t)f.prog.size());
            // covariate segments (covariate.rs:189-212) with a leading sentinel for t < first
            for (const auto& cname : labels.covariates) {
                auto it = occ.covariates.find(cname);
                if (it != occ.covariates.end() && !it->second.observations.empty()) {
                    auto obs = it->second.observations;
                    std::stable_sort(obs.begin(), obs.end(), [](auto& l, auto& r) { return total_key(l.first) < total_key(r.first); });
                    f.cov_segs.push_back(CovSeg{-inf, obs.front().first, 0.0, obs.front().second});
                    for (size_t i = 0; i < obs.size(); ++i) {
                        const bool has_next = i + 1 < obs.size();
                        CovSeg s{};
                        s.from = obs[i].first;
                        s.to = has_next ? obs[i + 1].first : inf;
                        if (it->second.fixed || !has_next) { s.slope = 0.0; s.intercept = obs[i].second; }
                        else {
                            const double slope = (obs[i + 1].second - obs[i].second) / (obs[i + 1].first - obs[i].first);
                            s.slope = slope; s.intercept = obs[i].second - slope * obs[i].first;
                        }
                        f.cov_segs.push_back(s);
                    }
                }
                f.cov_offsets.push_back((int32_t)f.cov_segs.size());
            }
            if (f.has_prog && f.prog_cov) {
                // Covariate::interpolate (covariate.rs:216-241) on the segments just built: first segment with from <= t < to
                auto cov_at = [&](double t) {
                    for (size_t q = seg_first; q < f.cov_segs.size(); ++q) {
                        const CovSeg& sg = f.cov_segs[q];
                        if (sg.from <= t && t < sg.to) return sg.slope * t + sg.intercept;
                    }
                    return std::numeric_limits<double>::quiet_NaN();
                };
                for (size_t q = prog_first; q < f.prog.size(); ++q) {
                    EventRec& r = f.prog[q];
                    if (ev_kind(r.meta) != EV_STEP) continue;
                    r.w = cov_at(r.time);       // COVTIME_INTERVAL_END
                    r.sigma = cov_at(r.a);      // COVTIME_INTERVAL_LENGTH (analytical! macro quirk: derive at t = dt)
                }
            }
            f.ev_offsets.push_back((int32_t)f.events.size());
            f.bol_offsets.push_back((int32_t)f.bol_event.size());
            f.inf_offsets.push_back((int32_t)f.infs.size());
            f.bnd_offsets.push_back((int32_t)f.bnds.size());
        }
        f.occ_offsets.push_back((int32_t)f.occ_index.size());
        f.obs_offsets.push_back(obs_row);
    }
    f.nobs_total = obs_row;
    return f;
}

}  // namespace pharmsol
