"""Extract every DSL source literal from the reference tree (tests, examples, benches, src unit tests, pharmsol-dsl)
together with the outcome its test expects (accepted / rejected + the first diagnostic substring the test asserts)
into tests/golden/dsl_corpus.json.  Run in the build container (needs /root/reference); the fixture travels."""
import glob
import json
import os
import re

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pat = re.compile(r'r#"(.*?)"#', re.S)
files = sorted(glob.glob(f"{REF}/tests/**/*.rs", recursive=True) + glob.glob(f"{REF}/examples/*.rs") + glob.glob(f"{REF}/benches/**/*.rs", recursive=True)
               + glob.glob(f"{REF}/src/**/*.rs", recursive=True) + glob.glob(f"{REF}/pharmsol-dsl/**/*.rs", recursive=True))
# literals that are templates (format! placeholders), parse-only inputs with undeclared names, or expected failures the
# heuristic cannot see: file:line -> override
OVERRIDE = {
    "pharmsol-dsl/src/parser.rs:1826": "skip", "pharmsol-dsl/src/parser.rs:1833": "skip",                      # parse_module only, `ke` undeclared
    "pharmsol-dsl/tests/dsl_authoring_edge_cases.rs:1092": "skip", "pharmsol-dsl/tests/dsl_authoring_edge_cases.rs:1170": "skip",   # format! templates
    "tests/authoring_parity_corpus.rs:110": "reject",
}
# models without any output equation analyse in the reference but cannot be evaluated; this backend refuses them at compile time
NO_OUTPUT = {"pharmsol-dsl/tests/dsl_authoring_edge_cases.rs:%d" % n for n in (825, 857, 956, 1008, 1123, 1207)}
rows, seen = [], set()
for f in files:
    txt = open(f).read()
    for m in pat.finditer(txt):
        src = m.group(1)
        if "//!" in src or "///" in src:
            src = "\n".join(re.sub(r"^\s*//[/!] ?", "", l) for l in src.split("\n"))
        if ("kind" not in src) or ("model " not in src and "name =" not in src and "name=" not in src) or src in seen:
            continue
        seen.add(src)
        line = txt[:m.start()].count("\n") + 1
        key = f"{f.replace(REF + '/', '')}:{line}"
        nxt = txt[m.end():m.end() + 900].split("#[test]")[0]
        expect = "reject" if re.search(r"expect_err|unwrap_err|is_err\(\)|should fail|must fail", nxt) else "accept"
        expect = OVERRIDE.get(key, expect)
        if expect == "skip":
            continue
        if key in NO_OUTPUT:
            expect = "reject-no-output"
        msg = None
        if expect == "reject":
            mm = re.search(r'contains\(\s*"((?:[^"\\]|\\.)*)"', nxt)
            msg = mm.group(1) if mm else None
        rows.append({"where": key, "expect": expect, "diagnostic": msg, "source": src})
json.dump(rows, open(os.path.join(ROOT, "tests", "golden", "dsl_corpus.json"), "w"), indent=1)
print(len(rows), "sources;", sum(r["expect"] == "accept" for r in rows), "accepted,", sum(r["expect"].startswith("reject") for r in rows), "rejected")
