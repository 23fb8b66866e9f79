// pmetrics.cpp — Pmetrics CSV -> Data (the `src/data` side of the psi path: datasets flatten into the SoA
// device buffers without going through per-row Subject builders in the host language).
//
// Restates /root/reference/src/data/parser/pmetrics/{mod.rs:164-239, 254-316, 341-442} (header handling, field
// parsing, "." / "NA" / "" = missing, OUT = -99 = missing observation, CENS vocabulary) and
// row.rs:{144-214 validate, 269-381 into_events with ADDL/II expansion, 593-672 build_data: group by ID,
// split occasions at EVID = 4, per-occasion covariates with conflict detection, `!` = fixed covariate,
// subjects sorted by ID}.
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <limits>
#include <map>
#include <set>
#include <sstream>

#include "data.hpp"

namespace pharmsol {

namespace {

[[noreturn]] void fail(const std::string& msg) { throw PharmsolError(psi::ST_OTHER, msg); }

std::string lower(std::string s) { for (auto& c : s) c = (char)std::tolower((unsigned char)c); return s; }
std::string upper(std::string s) { for (auto& c : s) c = (char)std::toupper((unsigned char)c); return s; }

// RFC-4180-style record splitting (quoted fields, doubled quotes); `#` starts a comment line
std::vector<std::vector<std::string>> read_records(const std::string& text) {
    std::vector<std::vector<std::string>> recs;
    size_t i = 0;
    const size_t n = text.size();
    while (i < n) {
        if (text[i] == '#') { while (i < n && text[i] != '\n') ++i; if (i < n) ++i; continue; }
        std::vector<std::string> rec;
        std::string field;
        bool quoted = false, any = false;
        while (i < n) {
            const char c = text[i];
            if (quoted) {
                if (c == '"') { if (i + 1 < n && text[i + 1] == '"') { field += '"'; i += 2; } else { quoted = false; ++i; } }
                else { field += c; ++i; }
                continue;
            }
            if (c == '"' && field.empty()) { quoted = true; any = true; ++i; continue; }
            if (c == ',') { rec.push_back(field); field.clear(); any = true; ++i; continue; }
            if (c == '\r') { ++i; continue; }
            if (c == '\n') { ++i; break; }
            field += c; any = true; ++i;
        }
        if (any || !field.empty()) { rec.push_back(field); recs.push_back(rec); }
    }
    return recs;
}

bool is_missing(const std::string& s) { return s.empty() || s == "." || s == "NA"; }

double parse_f64(const std::string& s, const std::string& what) {
    char* end = nullptr;
    const double v = std::strtod(s.c_str(), &end);
    if (end == s.c_str() || *end != '\0') fail("invalid float literal `" + s + "` in column " + what);
    return v;
}
long long parse_i64(const std::string& s, const std::string& what) {
    char* end = nullptr;
    const long long v = std::strtoll(s.c_str(), &end, 10);
    if (end == s.c_str() || *end != '\0') fail("invalid digit found in `" + s + "` in column " + what);
    return v;
}

const char* const kCore[15] = {"ID", "EVID", "TIME", "DUR", "DOSE", "ADDL", "II", "INPUT", "OUT", "OUTEQ", "CENS", "C0", "C1", "C2", "C3"};
int core_index(const std::string& header) {
    const std::string u = upper(header);
    for (int k = 0; k < 15; ++k) if (u == kCore[k]) return k;
    return -1;
}

struct Row {
    std::string id;
    long long evid = 0;
    double time = 0.0;
    std::optional<double> dur, dose, ii, out, c0, c1, c2, c3;
    std::optional<long long> addl;
    std::optional<std::string> input, outeq;
    std::optional<Censor> cens;
    std::vector<std::pair<std::string, double>> covs;   // key keeps the trailing `!` of fixed covariates
};

std::string fmt_time(double t) { std::ostringstream o; o << t; return o.str(); }

// DataRow::validate (row.rs:144-214)
void validate(const Row& r) {
    auto finite = [&](const std::optional<double>& v, const char* f) {
        if (v && !std::isfinite(*v)) fail(std::string("non-finite value in ") + f + " for subject " + r.id);
    };
    if (!std::isfinite(r.time)) fail("non-finite value in TIME for subject " + r.id);
    finite(r.dose, "DOSE"); finite(r.dur, "DUR"); finite(r.ii, "II"); finite(r.out, "OUT");
    finite(r.c0, "C0"); finite(r.c1, "C1"); finite(r.c2, "C2"); finite(r.c3, "C3");
    for (const auto& c : r.covs) if (!std::isfinite(c.second)) fail("non-finite value in covariate " + c.first + " for subject " + r.id);
    if (r.addl && *r.addl != 0) {
        if (!(r.evid == 1 || r.evid == 4)) fail("nonzero ADDL for " + r.id + " at time " + fmt_time(r.time) + " requires a dose row");
        if (!(r.ii && *r.ii > 0.0)) fail("nonzero ADDL for " + r.id + " at time " + fmt_time(r.time) + " requires a positive II");
    }
    if (r.evid == 4 && (!r.dose || !r.input)) fail("EVID=4 row for " + r.id + " at time " + fmt_time(r.time) + " must contain a dose and INPUT");
    if ((r.evid == 1 || r.evid == 4) && r.dur && *r.dur < 0.0) fail("negative DUR for " + r.id + " at time " + fmt_time(r.time));
    if (!(r.evid == 0 || r.evid == 1 || r.evid == 4)) fail("Unsupported EVID=" + std::to_string(r.evid) + " for subject " + r.id + " at time " + fmt_time(r.time));
}

// DataRow::into_events (row.rs:269-381)
void into_events(const Row& r, std::vector<Event>& out) {
    validate(r);
    if (r.evid == 0) {
        if (!r.outeq) fail("observation for " + r.id + " at time " + fmt_time(r.time) + " is missing OUTEQ");
        Event e;
        e.kind = EventKind::Observation; e.time = r.time; e.label = *r.outeq;
        e.has_value = r.out.has_value(); e.value = r.out.value_or(0.0);
        if (r.c0 && r.c1 && r.c2 && r.c3) { e.has_poly = true; e.poly = ErrorPoly{*r.c0, *r.c1, *r.c2, *r.c3}; }
        e.cens = r.cens.value_or(Censor::None);
        out.push_back(e);
        return;
    }
    if (!r.input) fail("dose for " + r.id + " at time " + fmt_time(r.time) + " is missing INPUT");
    if (!r.dose) fail("dose for " + r.id + " at time " + fmt_time(r.time) + " is missing DOSE");
    Event e;
    e.time = r.time; e.amount = *r.dose; e.label = *r.input;
    if (r.dur.value_or(0.0) > 0.0) { e.kind = EventKind::Infusion; e.duration = *r.dur; }
    else e.kind = EventKind::Bolus;
    if (r.addl && r.ii && *r.addl != 0) {
        // checked_abs + try_reserve in the reference (row.rs:333-347): refuse what cannot be materialised
        const long long reps = (*r.addl == std::numeric_limits<long long>::min()) ? std::numeric_limits<long long>::max() : std::llabs(*r.addl);
        if (reps > 20000000LL) fail("ADDL for " + r.id + " at time " + fmt_time(r.time) + " is too large to expand");
        const double interval = std::fabs(*r.ii), direction = *r.addl > 0 ? 1.0 : -1.0;
        for (long long k = 1; k <= reps; ++k) {
            const double offset = direction * interval * (double)k;
            if (!std::isfinite(e.time + offset)) fail("non-finite value in expanded TIME for subject " + r.id);
            Event rep = e;
            rep.time += offset;
            out.push_back(rep);
        }
    }
    out.push_back(e);
}

}  // namespace

Data read_pmetrics_text(const std::string& text) {
    const auto recs = read_records(text);
    static const std::vector<std::string> no_header;
    const auto& hdr = recs.empty() ? no_header : recs[0];
    // ---- headers (mod.rs:176-216) ----------------------------------------------------------------------
    std::vector<int> core_of(hdr.size(), -1);
    std::vector<std::string> cov_key(hdr.size());
    std::set<int> seen_core;
    std::map<std::string, bool> cov_forms;
    for (size_t k = 0; k < hdr.size(); ++k) {
        const int ci = core_index(hdr[k]);
        if (ci >= 0) {
            if (!seen_core.insert(ci).second) fail("duplicate core header `" + lower(kCore[ci]) + "`");
            core_of[k] = ci;
            continue;
        }
        const std::string& h = hdr[k];
        const bool fixed = !h.empty() && h.back() == '!';
        const std::string base = fixed ? h.substr(0, h.size() - 1) : h;
        bool bad = base.empty() || base.find('!') != std::string::npos || core_index(base) >= 0;
        for (char c : base) if (std::iscntrl((unsigned char)c)) bad = true;
        if (bad) fail("reserved or ambiguous covariate column `" + h + "`");
        const std::string name = lower(base);
        auto it = cov_forms.find(name);
        if (it != cov_forms.end())
            fail(it->second == fixed ? "duplicate covariate column `" + name + "`" : "covariate `" + name + "` is declared both with and without trailing !");
        cov_forms[name] = fixed;
        cov_key[k] = fixed ? name + "!" : name;
    }
    for (int req : {0, 1, 2}) if (!seen_core.count(req)) fail(std::string("missing required core header `") + kCore[req] + "`");

    // ---- rows -----------------------------------------------------------------------------------------------
    std::map<std::string, std::vector<Row>> by_id;     // BTreeMap order == the reference's final sort by ID
    for (size_t ri = 1; ri < recs.size(); ++ri) {
        const auto& rec = recs[ri];
        if (rec.size() != hdr.size()) fail("CSV error: record " + std::to_string(ri) + " has " + std::to_string(rec.size()) + " fields, expected " + std::to_string(hdr.size()));
        Row r;
        bool have_id = false, have_evid = false, have_time = false;
        for (size_t k = 0; k < rec.size(); ++k) {
            const std::string& s = rec[k];
            const int ci = core_of[k];
            if (ci < 0) { if (!is_missing(s)) r.covs.emplace_back(cov_key[k], parse_f64(s, cov_key[k])); continue; }
            switch (ci) {
                case 0: r.id = s; have_id = true; break;
                case 1: r.evid = parse_i64(s, "EVID"); have_evid = true; break;
                case 2: r.time = parse_f64(s, "TIME"); have_time = true; break;
                case 3: if (!is_missing(s)) r.dur = parse_f64(s, "DUR"); break;
                case 4: if (!is_missing(s)) r.dose = parse_f64(s, "DOSE"); break;
                case 5: if (!is_missing(s)) r.addl = parse_i64(s, "ADDL"); break;
                case 6: if (!is_missing(s)) r.ii = parse_f64(s, "II"); break;
                case 7: if (!is_missing(s)) r.input = s; break;
                case 8: if (!is_missing(s)) { const double v = parse_f64(s, "OUT"); if (v != -99.0) r.out = v; } break;   // mod.rs:296
                case 9: if (!is_missing(s)) r.outeq = s; break;
                case 10:
                    if (!is_missing(s)) {
                        if (s == "1" || s == "bloq") r.cens = Censor::BLOQ;
                        else if (s == "0" || s == "none") r.cens = Censor::None;
                        else if (s == "-1" || s == "aloq") r.cens = Censor::ALOQ;
                        else fail("Expected one of 1/-1/0 or bloq/aloq/none), got " + s);
                    }
                    break;
                case 11: if (!is_missing(s)) r.c0 = parse_f64(s, "C0"); break;
                case 12: if (!is_missing(s)) r.c1 = parse_f64(s, "C1"); break;
                case 13: if (!is_missing(s)) r.c2 = parse_f64(s, "C2"); break;
                case 14: if (!is_missing(s)) r.c3 = parse_f64(s, "C3"); break;
            }
        }
        if (!have_id || !have_evid || !have_time) fail("CSV error: missing ID / EVID / TIME");
        if (r.id.empty()) fail("empty subject ID at record " + std::to_string(ri));
        by_id[r.id].push_back(std::move(r));
    }

    // ---- build_data (row.rs:593-672) -----------------------------------------------------------------------------
    Data data;
    for (auto& kv : by_id) {
        const std::vector<Row>& rows = kv.second;
        std::vector<std::pair<size_t, size_t>> blocks;
        size_t start = 0;
        for (size_t i = 0; i < rows.size(); ++i)
            if (rows[i].evid == 4) { if (start < i) blocks.emplace_back(start, i); start = i; }
        if (start < rows.size()) blocks.emplace_back(start, rows.size());
        Subject subj;
        subj.id = kv.first;
        for (size_t b = 0; b < blocks.size(); ++b) {
            Occasion occ;
            occ.index = (int)b;
            std::map<std::string, std::vector<std::pair<double, double>>> observed;
            for (size_t i = blocks[b].first; i < blocks[b].second; ++i) {
                into_events(rows[i], occ.events);
                for (const auto& c : rows[i].covs) {
                    auto& obs = observed[c.first];
                    bool dup = false;
                    for (const auto& o : obs)
                        if (o.first == rows[i].time) {
                            if (o.second != c.second)
                                fail("conflicting covariate `" + c.first + "` values for subject `" + subj.id + "` occasion " + std::to_string(b) + " at time " + fmt_time(rows[i].time));
                            dup = true;
                        }
                    if (!dup) obs.emplace_back(rows[i].time, c.second);
                }
            }
            for (auto& e : occ.events) e.occasion = (int)b;
            for (const auto& ov : observed) {            // Covariates::from_row_observations (covariate.rs:316-333)
                const bool fixed = !ov.first.empty() && ov.first.back() == '!';
                Covariate cov;
                cov.name = fixed ? ov.first.substr(0, ov.first.size() - 1) : ov.first;
                cov.fixed = fixed;
                for (const auto& o : ov.second) cov.add_observation(o.first, o.second);
                occ.covariates[cov.name] = cov;
            }
            occ.sort();
            subj.occasions.push_back(std::move(occ));
        }
        data.subjects.push_back(std::move(subj));
    }
    return data;
}

Data read_pmetrics_file(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) fail("CSV error: cannot open " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return read_pmetrics_text(ss.str());
}

// Data::expand (data/structs.rs:155-260): add missing observations on a regular grid (every `idelta` from 0 to the
// last dose end + `tad`, per occasion, for every output label of the dataset) unless an observation of that output
// already exists at that time (compared in integer microseconds) — dense prediction grids.
Data expand_data(const Data& d, double idelta, double tad) {
    if (!(idelta > 0.0)) return d;
    const unsigned long long step_us = (unsigned long long)std::llround(idelta * 1e6);
    if (step_us == 0) return d;
    std::set<std::string> outeqs;                         // sorted + deduplicated (structs.rs:294-304)
    for (const auto& s : d.subjects) for (const auto& o : s.occasions) for (const auto& e : o.events)
        if (e.kind == EventKind::Observation) outeqs.insert(e.label);
    Data out;
    for (const auto& s : d.subjects) {
        Subject ns; ns.id = s.id;
        for (const auto& o : s.occasions) {
            double last = 0.0; bool any = false;
            for (const auto& e : o.events) {
                if (e.kind == EventKind::Observation) continue;
                const double t = e.kind == EventKind::Bolus ? e.time : e.time + e.duration;
                if (!any || t > last) { last = t; any = true; }
            }
            last = (any ? last : 0.0) + tad;
            std::set<std::pair<unsigned long long, std::string>> existing;
            for (const auto& e : o.events)
                if (e.kind == EventKind::Observation) existing.insert({(unsigned long long)std::llround(e.time * 1e6), e.label});
            Occasion no; no.index = o.index; no.covariates = o.covariates;
            const double last_us_f = std::round(last * 1e6);
            const unsigned long long last_us = last_us_f < 0.0 ? 0ull : (unsigned long long)last_us_f;   // Rust `as u64` saturates at 0
            for (unsigned long long key = 0; key <= last_us; key += step_us) {
                for (const auto& lab : outeqs) {
                    if (existing.count({key, lab})) continue;
                    Event e; e.kind = EventKind::Observation; e.time = (double)key / 1e6; e.label = lab; e.has_value = false; e.occasion = o.index;
                    no.events.push_back(e);
                }
            }
            no.events.insert(no.events.end(), o.events.begin(), o.events.end());
            no.sort();
            ns.occasions.push_back(std::move(no));
        }
        out.subjects.push_back(std::move(ns));
    }
    return out;
}

// JSON description of a dataset (inspection / tests): subjects -> occasions -> events + covariates.
std::string describe_data_json(const Data& d) {
    std::ostringstream o;
    o.precision(17);
    auto str = [&](const std::string& s) { o << '"'; for (char c : s) { if (c == '"' || c == '\\') o << '\\'; o << c; } o << '"'; };
    o << "[";
    for (size_t i = 0; i < d.subjects.size(); ++i) {
        const auto& s = d.subjects[i];
        o << (i ? ", " : "") << "{\"id\": "; str(s.id); o << ", \"occasions\": [";
        for (size_t k = 0; k < s.occasions.size(); ++k) {
            const auto& oc = s.occasions[k];
            o << (k ? ", " : "") << "{\"index\": " << oc.index << ", \"events\": [";
            for (size_t e = 0; e < oc.events.size(); ++e) {
                const auto& ev = oc.events[e];
                o << (e ? ", " : "") << "{\"kind\": \"" << (ev.kind == EventKind::Observation ? "observation" : ev.kind == EventKind::Bolus ? "bolus" : "infusion")
                  << "\", \"time\": " << ev.time << ", \"label\": "; str(ev.label);
                if (ev.kind == EventKind::Observation) {
                    o << ", \"value\": "; if (ev.has_value) o << ev.value; else o << "null";
                    o << ", \"censoring\": " << (int)ev.cens;
                    if (ev.has_poly) o << ", \"errorpoly\": [" << ev.poly.c0 << ", " << ev.poly.c1 << ", " << ev.poly.c2 << ", " << ev.poly.c3 << "]";
                } else {
                    o << ", \"amount\": " << ev.amount;
                    if (ev.kind == EventKind::Infusion) o << ", \"duration\": " << ev.duration;
                }
                o << "}";
            }
            o << "], \"covariates\": {";
            bool first = true;
            for (const auto& cv : oc.covariates) {
                o << (first ? "" : ", "); first = false; str(cv.first); o << ": {\"fixed\": " << (cv.second.fixed ? "true" : "false") << ", \"observations\": [";
                for (size_t q = 0; q < cv.second.observations.size(); ++q)
                    o << (q ? ", " : "") << "[" << cv.second.observations[q].first << ", " << cv.second.observations[q].second << "]";
                o << "]}";
            }
            o << "}}";
        }
        o << "]}";
    }
    o << "]";
    return o.str();
}

}  // namespace pharmsol
