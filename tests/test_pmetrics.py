"""Pmetrics CSV -> Data: the reference's parser tests re-run against the native C++ reader
(src/data/parser/pmetrics/tests.rs and mod.rs tests; file:line cited per test).  CPU only: parsing needs no device."""
import os

import numpy as np
import pytest

CORE = "ID,EVID,TIME,DUR,DOSE,ADDL,II,INPUT,OUT,OUTEQ,CENS,C0,C1,C2,C3"


def parse(ps, text):
    return ps.Data.from_pmetrics(text=text).native().describe()


def inp(covs, rows):
    return ",".join([CORE] + list(covs)) + "\n" + rows


def test_addl_expansion_file(ps, tmp_path):
    # mod.rs tests::test_addl with src/tests/data/addl_test.csv (contents restated)
    f = tmp_path / "addl.csv"
    f.write_text("ID,EVID,TIME,DUR,DOSE,ADDL,II,INPUT,OUT,OUTEQ,C0,C1,C2,C3\n1,1,0,0,600,-10,12,1,.,.,.,.,.,.\n1,0,9,.,.,.,.,.,100,100,.,.,.,.\n"
                 "2,1,0,0,600,10,12,1,.,.,.,.,.,.\n2,0,9,.,.,.,.,.,100,100,.,.,.,.\n")
    d = ps.read_pmetrics(str(f)).native().describe()
    t1 = [e["time"] for e in d[0]["occasions"][0]["events"]]
    t2 = [e["time"] for e in d[1]["occasions"][0]["events"]]
    assert t1 == [-120.0, -108.0, -96.0, -84.0, -72.0, -60.0, -48.0, -36.0, -24.0, -12.0, 0.0, 9.0]
    assert t2 == [0.0, 9.0, 12.0, 24.0, 36.0, 48.0, 60.0, 72.0, 84.0, 96.0, 108.0, 120.0]


def test_labels_are_preserved_as_strings(ps):
    # mod.rs tests: named and numeric INPUT / OUTEQ labels
    d = parse(ps, CORE + "\npt1,1,0,1,100,.,.,iv,.,.,.,.,.,.,.\npt1,0,1,.,.,.,.,.,42,cp,0,.,.,.,.\n")
    ev = d[0]["occasions"][0]["events"]
    assert ev[0]["kind"] == "infusion" and ev[0]["label"] == "iv" and ev[1]["kind"] == "observation" and ev[1]["label"] == "cp"
    d = parse(ps, CORE + "\npt1,1,0,.,100,.,.,1,.,.,.,.,.,.,.\npt1,0,1,.,.,.,.,.,42,1,0,.,.,.,.\n")
    ev = d[0]["occasions"][0]["events"]
    assert ev[0]["kind"] == "bolus" and ev[0]["label"] == "1" and ev[1]["label"] == "1"


def test_duplicate_and_conflicting_headers_are_rejected(ps):
    # tests.rs:221-238
    for hdr in [CORE + ",WT,wt\n", CORE + ",WT!,wt!\n", CORE + ",WT,wt!\n", "id," + CORE + "\n", CORE + ",wt!!\n", CORE + ",wt!x\n"]:
        with pytest.raises(ps.PharmsolError):
            parse(ps, hdr)


def test_required_core_headers(ps):
    # tests.rs:241-254
    for text, missing in [("", "ID"), ("EVID,TIME\n", "ID"), ("ID,TIME\n", "EVID"), ("ID,EVID\n", "TIME")]:
        with pytest.raises(ps.PharmsolError) as e:
            parse(ps, text)
        assert f"missing required core header `{missing}`" in str(e.value)


def test_unused_core_headers_may_be_omitted(ps):
    # tests.rs:257-270
    assert parse(ps, "ID,EVID,TIME,DOSE,INPUT\ns,1,0,100,iv\n")[0]["occasions"][0]["events"][0]["kind"] == "bolus"
    assert parse(ps, "ID,EVID,TIME,OUT,OUTEQ\ns,0,0,1.5,cp\n")[0]["occasions"][0]["events"][0]["kind"] == "observation"


def test_mixed_case_covariate_headers_are_normalized(ps):
    # tests.rs:273-283
    cov = parse(ps, inp(["WT!", "Ka"], "s,1,0,0,1,.,.,iv,.,.,.,.,.,.,.,70,0.5\n"))[0]["occasions"][0]["covariates"]
    assert cov["wt"]["fixed"] is True and cov["ka"]["fixed"] is False


def test_negative_addl_reset_starts_occasion_at_earliest_dose(ps):
    # tests.rs:444-463
    d = parse(ps, inp([], "s,1,0,0,1,.,.,iv,.,.,.,.,.,.,.\ns,4,0,0,2,-2,1,iv,.,.,.,.,.,.,.\n"))
    occ = d[0]["occasions"][1]
    assert [e["time"] for e in occ["events"]] == [-2.0, -1.0, 0.0] and occ["index"] == 1


def test_addl_validation(ps):
    # tests.rs:507-539
    for ii in [".", "0", "-1"]:
        with pytest.raises(ps.PharmsolError) as e:
            parse(ps, CORE + f"\ns,1,0,0,1,2,{ii},iv,.,.,.,.,.,.,.\n")
        assert "requires a positive II" in str(e.value)
    with pytest.raises(ps.PharmsolError) as e:
        parse(ps, CORE + f"\ns,1,0,0,1,{-2**63},1,iv,.,.,.,.,.,.,.\n")
    assert "too large to expand" in str(e.value)
    with pytest.raises(ps.PharmsolError) as e:
        parse(ps, CORE + "\ns,1,0,0,1,2,1e308,iv,.,.,.,.,.,.,.\n")
    assert "expanded TIME" in str(e.value)


def test_covariate_rows(ps):
    # tests.rs:567-620
    same = parse(ps, inp(["wt"], "s,1,0,0,1,.,.,iv,.,.,.,.,.,.,.,70\ns,0,0,.,.,.,.,.,1,cp,0,.,.,.,.,70\n"))
    assert same[0]["occasions"][0]["covariates"]["wt"]["observations"] == [[0.0, 70.0]]
    with pytest.raises(ps.PharmsolError) as e:
        parse(ps, inp(["wt"], "s,1,0,0,1,.,.,iv,.,.,.,.,.,.,.,70\ns,0,0,.,.,.,.,.,1,cp,0,.,.,.,.,71\n"))
    msg = str(e.value)
    assert "conflicting covariate `wt` values" in msg and "subject `s` occasion 0" in msg and "time 0" in msg
    two = parse(ps, inp(["wt"], "s,1,0,0,1,.,.,iv,.,.,.,.,.,.,.,70\ns,0,24,.,.,.,.,.,1,cp,0,.,.,.,.,72\n"))
    assert two[0]["occasions"][0]["covariates"]["wt"]["observations"] == [[0.0, 70.0], [24.0, 72.0]]


def test_missing_observation_placeholders(ps):
    # tests.rs:623-642 and mod.rs:296 (OUT = -99)
    d = parse(ps, CORE + "\ns,0,0,.,.,.,.,.,.,cp,0,.,.,.,.\ns,0,1,.,.,.,.,.,NA,cp,0,.,.,.,.\ns,0,2,.,.,.,.,.,,cp,0,.,.,.,.\ns,0,3,.,.,.,.,.,-99,cp,0,.,.,.,.\n")
    assert [e["value"] for e in d[0]["occasions"][0]["events"]] == [None, None, None, None]


def test_evid_rules(ps):
    # tests.rs:657-676
    with pytest.raises(ps.PharmsolError) as e:
        parse(ps, inp(["wt"], "s,2,0,.,.,.,.,.,.,.,.,.,.,.,.,70\n"))
    assert "Unsupported EVID=2" in str(e.value)
    with pytest.raises(ps.PharmsolError) as e:
        parse(ps, CORE + "\ns,4,0,.,.,.,.,.,.,.,.,.,.,.,.\n")
    assert "must contain a dose" in str(e.value)


def test_censoring_errorpoly_and_subject_order(ps):
    d = parse(ps, CORE + "\nb,1,0,.,100,.,.,1,.,.,.,.,.,.,.\nb,0,1,.,.,.,.,.,0.2,1,1,0.1,0.2,0,0\nb,0,2,.,.,.,.,.,9.0,1,aloq,.,.,.,.\na,1,0,.,50,.,.,1,.,.,.,.,.,.,.\n")
    assert [s["id"] for s in d] == ["a", "b"]                      # row.rs:668-669: subjects sorted by ID
    ev = d[1]["occasions"][0]["events"]
    assert ev[1]["censoring"] == 1 and ev[1]["errorpoly"] == [0.1, 0.2, 0.0, 0.0] and ev[2]["censoring"] == 2


def test_parsed_data_feeds_the_oracle_and_the_builder_ops(ps, oracle):
    """The parsed dataset mirrors into builder ops (what the GPU flattener and the oracle both consume)."""
    text = inp(["wt"], "1,1,0,0.5,500,.,.,iv,.,.,.,.,.,.,.,70\n1,0,0.5,.,.,.,.,.,3.1,cp,0,.,.,.,.,70\n1,0,4,.,.,.,.,.,1.5,cp,0,.,.,.,.,72\n"
                       "1,4,0,1,300,.,.,iv,.,.,.,.,.,.,.,75\n1,0,2,.,.,.,.,.,2.2,cp,0,.,.,.,.,75\n")
    data = ps.Data.from_pmetrics(text=text)
    assert len(data) == 1
    ops = data.subjects[0].ops
    assert ("reset",) in ops and ("infusion", 0.0, 500.0, "iv", 0.5) in ops and ("infusion", 0.0, 300.0, "iv", 1.0) in ops
    s = oracle.Subject(ops)
    assert s.n_occasions() == 2
    pr = oracle.Model("one_cpt_iv").predictions(s, [0.3, 100.0])
    assert len(pr) == 3 and np.all(np.isfinite(pr))


def _obs_times(occ):
    return sorted(e["time"] for e in occ["events"] if e["kind"] == "observation")


def test_expand_grid_reaches_last_dose_plus_tad(ps):
    # data/structs.rs:1700-1723
    d = ps.Data([ps.Subject.builder("s1").bolus(0.0, 100.0, 0).observation(0.0, 5.0, 0).build()]).expand(1.0, 3.0)
    assert _obs_times(d.native().describe()[0]["occasions"][0]) == [0.0, 1.0, 2.0, 3.0]
    # the original observation keeps its value, the grid points are missing observations
    ev = [e for e in d.native().describe()[0]["occasions"][0]["events"] if e["kind"] == "observation"]
    assert sum(e["value"] is not None for e in ev) == 1


def test_expand_last_time_is_per_occasion(ps):
    # data/structs.rs:1726-1760
    s = (ps.Subject.builder("s1").bolus(0.0, 100.0, 0).observation(0.0, 5.0, 0).reset()
         .bolus(10.0, 100.0, 0).observation(10.0, 5.0, 0).build())
    occ = ps.Data([s]).expand(5.0, 0.0).native().describe()[0]["occasions"]
    assert _obs_times(occ[0]) == [0.0] and _obs_times(occ[1]) == [0.0, 5.0, 10.0]
    assert ps.Data([s]).expand(0.0, 5.0).native().describe()[0]["occasions"][0]["events"].__len__() == 2     # idelta <= 0: unchanged
