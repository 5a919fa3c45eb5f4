#!/usr/bin/env python3
"""Extract the reference's only recorded simulator output into a committed fixture.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Run in the build container, where
/root/reference exists; the GPU box never reads /root/reference, it reads the JSON this
script writes (tests/golden/fit_curves_golden.json).

Source: /root/reference/docs/src/fitting_data/parameters.xlsx_fit.xlsx, written by the
reference's `savefit` (src/fitting.jl:320-431).  Sheet 1 columns are, per antibody
concentration j: times[j], refdata[j] (experimental), means(TotalBoundOutputter) of
`update_pars_and_run_spr_sim!` (src/fitting.jl:331-345) -- i.e. the mean over nsims=250
(docs/src/fitting.md:34,75) replicates of run_spr_sim! at the best-fit internal parameters
of sheet 2 with tsave = the data times, tstop = last(times), tstop_AtoB = 150, default
geometry (N=1000, L from DEFAULT_SIM_ANTIGENCONCEN, DIM=3, resampled positions).
For concentration j the first parameter is logkon + log10(abc_j/abc_1) (fitting.jl:342).

No openpyxl in this image: the sheet XML is parsed with zipfile + regex.
"""
import hashlib
import json
import os
import re
import sys
import zipfile

SRC = "/root/reference/docs/src/fitting_data/parameters.xlsx_fit.xlsx"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden",
                   "fit_curves_golden.json")


def cells(xml):
    out = {}
    for m in re.finditer(r'<c r="([A-Z]+)(\d+)"([^>]*?)(?:/>|>(.*?)</c>)', xml, re.S):
        col, row, attrs, body = m.group(1), int(m.group(2)), m.group(3), m.group(4)
        if body is None:
            continue
        v = re.search(r"<v>(.*?)</v>", body, re.S)
        if v is None:
            continue
        out[(col, row)] = (v.group(1), 't="s"' in attrs)
    return out


def extract(src):
    """the savefit workbook at `src` (the reference's recorded one, or one written by api.savefit) as a dict"""
    raw = open(src, "rb").read()
    z = zipfile.ZipFile(src)
    shared = re.findall(r"<si><t[^>]*>(.*?)</t></si>", z.read("xl/sharedStrings.xml").decode(), re.S)
    s1 = cells(z.read("xl/worksheets/sheet1.xml").decode())
    s2 = cells(z.read("xl/worksheets/sheet2.xml").decode())

    def val(sheet, col, row):
        v, is_s = sheet[(col, row)]
        return shared[int(v)] if is_s else float(v)

    header = [val(s1, c, 1) for c in "ABCDEF"]
    assert header[0] == "times" and header[3] == "times", header
    conc = [float(re.search(r"([\d.]+) nM", header[1]).group(1)),
            float(re.search(r"([\d.]+) nM", header[4]).group(1))]
    nrow = max(r for (_, r) in s1)
    cols = {c: [val(s1, c, r) for r in range(2, nrow + 1)] for c in "ABCDEF"}
    assert cols["A"] == cols["D"]

    internal = {}
    physical = {}
    fitness = None
    for r in range(2, 9):
        if ("A", r) in s2 and ("B", r) in s2:
            name = val(s2, "A", r)
            if name == "Bivalent Fitness":
                fitness = val(s2, "B", r)
            else:
                internal[name] = val(s2, "B", r)
        if ("D", r) in s2 and ("E", r) in s2:
            physical[val(s2, "D", r)] = val(s2, "E", r)

    gold = {
        "source": "docs/src/fitting_data/parameters.xlsx_fit.xlsx (reference tree)",
        "source_sha256": hashlib.sha256(raw).hexdigest(),
        "written_by": "src/fitting.jl:320-431 (savefit) -> update_pars_and_run_spr_sim! (fitting.jl:227-241)",
        "nsims": 250,
        "N": 1000,
        "DIM": 3,
        "tstop_AtoB": 150.0,
        "antigenconcen_internal_uM": 500 / 149 / 26795 * 1000000,
        "antibody_nM": conc,
        "internal_params": internal,
        "physical_params": physical,
        "fitness": fitness,
        "times": cols["A"],
        "data": [cols["B"], cols["E"]],
        "sim_mean": [cols["C"], cols["F"]],
    }
    return gold


def main():
    gold = extract(SRC)
    with open(OUT, "w") as f:
        json.dump(gold, f, indent=0)
    internal = gold["internal_params"]
    print("wrote", os.path.normpath(OUT), "rows", len(gold["times"]), "params", internal, gold["physical_params"], gold["fitness"])
    # sanity: value*nsims*N/CP should be an integer count
    CP = 10.0 ** internal["logCP"]
    x = gold["sim_mean"][0][140] / CP * 250 * 1000
    print("row 142 count check:", gold["times"][140], gold["sim_mean"][0][140], x)


if __name__ == "__main__":
    sys.exit(main())
