"""CPU: the fitting-side oracle (oracle/fit_oracle.py: surrogate lookup + l2 loss of fitting.jl:38-91) against
closed forms, the aligned-data CSV reader (spr_data.jl:31-62) and the parameter conversions against the
reference's recorded best fit (tests/golden/fit_curves_golden.json)."""
import json
import math
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def affine_table(size, ranges, tsave, coef):
    """table that is multilinear in the axis values (so multilinear interpolation is exact)"""
    ax = [np.linspace(r[0], r[1], n) for r, n in zip(ranges, size[:4])] + [np.asarray(tsave)]
    g = np.meshgrid(*ax, indexing="ij")
    return coef[0] + coef[1] * g[0] + coef[2] * g[1] * g[2] + coef[3] * g[3] + coef[4] * g[4] * g[0]


def test_interpolant_is_exact_on_multilinear_tables_and_loss_matches_closed_form():
    from oracle import fit_oracle as F
    size = (4, 3, 5, 3, 11)
    ranges = [(-3.0, 1.0), (-2.0, 0.0), (-1.0, 1.0), (5.0, 25.0)]
    tsave = np.linspace(0.0, 50.0, 11)
    coef = (0.3, 0.02, 0.01, 0.004, 0.0007)
    tab = affine_table(size, ranges, tsave, coef)
    f = lambda x1, x2, x3, x4, t: coef[0] + coef[1] * x1 + coef[2] * x2 * x3 + coef[3] * x4 + coef[4] * t * x1
    rng = np.random.default_rng(3)
    times = [np.sort(rng.uniform(0.0, 50.0, 17)), np.sort(rng.uniform(0.0, 50.0, 9))]
    abcs = [25.0, 100.0]
    refdata = [rng.normal(1.0, 0.2, 17), rng.normal(2.0, 0.2, 9)]
    for _ in range(20):
        op = [rng.uniform(-3.0, 1.0 - math.log10(4.0)), rng.uniform(-2, 0), rng.uniform(-1, 1), rng.uniform(5, 25),
              rng.uniform(0, 1)]
        want = 0.0
        for j, abc in enumerate(abcs):
            x1 = op[0] + math.log10(abc / abcs[0])
            for t, d in zip(times[j], refdata[j]):
                want += (10.0 ** op[4] * f(x1, op[1], op[2], op[3], t) - d) ** 2
        got = F.surrogate_sprdata_error(op, tab, ranges, tsave, times, refdata, abcs)
        assert got == pytest.approx(want, rel=1e-12)
    # nodes reproduce the table, the last node of every axis is inside, one step beyond is a BoundsError
    assert F.itp(tab, (4, 3, 5, 3, 11)) == pytest.approx(tab[3, 2, 4, 2, 10], rel=1e-15)
    assert F.itp(tab, (1, 1, 1, 1, 1)) == tab[0, 0, 0, 0, 0]
    with pytest.raises(IndexError):
        F.itp(tab, (4.0001, 1, 1, 1, 1))
    with pytest.raises(IndexError):
        F.itp(tab, (1, 1, 1, 1, 0.999))


def test_oracle_agrees_with_the_package_interpolant():
    """api.Surrogate.itp (the host mirror of surrogate.jl:66-70) and the oracle are independent restatements"""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "sprfittingpaper2023.jl_b200"))
    from oracle import fit_oracle as F
    from sprb200 import api
    rng = np.random.default_rng(11)
    tab = rng.random((3, 4, 2, 5, 6))
    sur = api.Surrogate(tab)
    for _ in range(50):
        q = [rng.uniform(1, n) for n in tab.shape]
        assert sur.itp(*q) == pytest.approx(F.itp(tab, q), rel=1e-13)


def test_get_aligned_data_reads_the_reference_csv_layout(tmp_path):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "sprfittingpaper2023.jl_b200"))
    from sprb200 import api
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "fit_curves_golden.json")))
    t = gold["times"]
    d0, d1 = gold["data"]
    # the reference's file: header "time (s), 25.000000,time (s), 100.000000", one row per time, ragged tail allowed
    fn = tmp_path / "Data_FC4_test_RBD-13.8_aligned.csv"
    with open(fn, "w") as f:
        f.write("time (s), 25.000000,time (s), 100.000000\n")
        for k in range(len(t)):
            if k < len(t) - 3:
                f.write("%.18e,%.18e,%.18e,%.18e\n" % (t[k], d0[k], t[k], d1[k]))
            else:
                f.write("%.18e,%.18e,,\n" % (t[k], d0[k]))          # missing cells in the second curve
    ad = api.get_aligned_data(str(fn))
    assert ad.antigenconcen == 13.8 and list(ad.antibodyconcens) == [25.0, 100.0]
    assert np.array_equal(ad.times[0], np.array(t)) and np.array_equal(ad.refdata[0], np.array(d0))
    assert len(ad.times[1]) == len(t) - 3 and np.array_equal(ad.refdata[1], np.array(d1[:-3]))
    assert api.get_aligned_data(str(fn), 99.0).antigenconcen == 99.0


def test_parameter_conversions_match_the_recorded_fit():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "sprfittingpaper2023.jl_b200"))
    from sprb200 import api
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "fit_curves_golden.json")))
    ip, pp = gold["internal_params"], gold["physical_params"]
    op = [ip["logkon"], ip["logkoff"], ip["logkonb"], ip["reach"], ip["logCP"]]
    phys = api.optpars_to_physpars(op, gold["antibody_nM"][0], 13.8, gold["antigenconcen_internal_uM"])
    for got, name in zip(phys, ("kon", "koff", "konb", "reach", "CP")):
        assert got == pytest.approx(pp[name], rel=1e-12), name
    sp = api.SurrogateParams((-5.0, 2.0), (-4.0, 0.0), (-3.0, 1.5), (2.0, 35.0))
    api.checkranges([(-4, 1), (-3, 0), (-3, 1), (2, 35), (1, 3)], sp)
    with pytest.raises(ValueError):
        api.checkranges([(-6, 1), (-3, 0), (-3, 1), (2, 35), (1, 3)], sp)
    # a 4x concentration range needs log10(4) of head room at the top of the logkon axis (fitting.jl:103-114)
    lo, hi = api.rescale_logkon_max((-5.0, 2.0), sp.logkon_range, [25.0, 100.0])
    assert lo == -5.0 and hi == pytest.approx(2.0 - math.log10(4.0))
    assert api.rescale_logkon_max((-5.0, 1.0), sp.logkon_range, [25.0, 100.0]) == (-5.0, 1.0)


@pytest.mark.parametrize("method", ["xnes", "de"])
def test_population_optimisers_find_a_known_minimum(method):
    """the two stand-ins for the reference's Optimization.jl methods (xNES is its default, fitting.jl:26-37) on
    an anisotropic, coupled quadratic in the fit's 5-D box; the objective takes a whole population per call,
    as the GPU loss does"""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "sprfittingpaper2023.jl_b200"))
    from sprb200 import api
    lb = np.array([-4.0, -2.5, -1.5, 5.0, 1.0])
    ub = np.array([-1.0, -0.5, 1.0, 25.0, 3.0])
    xstar = np.array([-2.88, -1.39, -0.11, 15.3, 2.11])
    calls = []

    def f(pop):
        assert pop.ndim == 2 and pop.shape[1] == 5 and np.all(pop >= lb) and np.all(pop <= ub)
        calls.append(len(pop))
        d = (pop - xstar) / (ub - lb)
        return 1e3 * (d[:, 0] ** 2 + 10 * d[:, 1] ** 2 + 0.1 * d[:, 2] ** 2 + 3 * (d[:, 3] - d[:, 2]) ** 2 + d[:, 4] ** 2) + 79.0

    rng = np.random.default_rng(1)
    sol = api.minimise_xnes(f, lb, ub, rng) if method == "xnes" else api.minimise_de(f, lb, ub, rng)
    assert sol.minimum == pytest.approx(79.0, abs=1e-6)
    assert np.abs(sol.u - xstar).max() < 1e-3
    assert sol.evals == sum(calls) and set(calls) == ({8} if method == "xnes" else {64})
    # a minimum on the boundary of the box is found too (candidates are reflected into the box)
    g = lambda pop: ((pop - lb) ** 2).sum(axis=1)
    rng = np.random.default_rng(2)
    sol = api.minimise_xnes(g, lb, ub, rng, maxiters=3000) if method == "xnes" else api.minimise_de(g, lb, ub, rng)
    assert sol.minimum < 1e-6


def test_monovalent_fit_recovers_manufactured_parameters(built_lib):
    """the reference's own test (test/monovalent_fit.jl): data manufactured with monovalent_total_bound at three
    antibody concentrations on three time grids, random start, parameters recovered to 1e-10 (:45); plus the closed
    form against the two-state master equation and the loss/gradient against finite differences."""
    from sprb200 import api
    kon, koff, CP, toff, tstop = 1e-2, 1e-1, 8.5, 150.0, 600.0
    tv = np.linspace(0.0, tstop, 601)
    times = [tv, tv[0::2], tv[1::2]]
    abcs = [25.0, 100.0, 150.0]
    data = [api.monovalent_total_bound((kon, koff, CP), (abc, toff), t) for abc, t in zip(abcs, times)]
    # closed form == integral of dA/dt = kon Ab [t <= toff] (1 - A) - koff A, A(0) = 0 (RK4, 1e-2 steps)
    for abc in (25.0, 150.0):
        A, h = 0.0, 0.01
        for on, nsteps in ((1.0, 15000), (0.0, 5000)):               # injection until toff = 150, wash-out until 200
            f = lambda A: kon * abc * on * (1.0 - A) - koff * A
            for _ in range(nsteps):
                k1 = f(A); k2 = f(A + 0.5 * h * k1); k3 = f(A + 0.5 * h * k2); k4 = f(A + h * k3)
                A += h * (k1 + 2 * k2 + 2 * k3 + k4) / 6.0
        assert abs(CP * A - api.monovalent_total_bound((kon, koff, CP), (abc, toff), 200.0)) < 1e-9
    assert api.monovalent_total_bound((kon, koff, CP), (25.0, toff), 0.0) == 0.0
    ad = api.AlignedData(times, data, abcs, CP)
    assert api.monovalent_sprdata_error(np.log10([kon, koff, CP]), ad, toff) == 0.0
    x = np.array([-1.7, -1.2, 0.7])
    want = sum(float(np.sum((api.monovalent_total_bound(tuple(10.0 ** x), (abc, toff), t) - d) ** 2))
               for abc, t, d in zip(abcs, times, data))
    assert math.isclose(api.monovalent_sprdata_error(x, ad, toff), want, rel_tol=1e-14)
    rng = np.random.default_rng(11)
    lb, ub = [-8.0] * 3, [8.0] * 3
    for _ in range(3):
        u0 = np.log10(rng.random(3))                                  # the reference's start: log10.(rand(3))
        sol = api.monovalent_fit_spr_data(ad, toff, u0, lb, ub)
        assert np.max(np.abs(10.0 ** sol.u - np.array([kon, koff, CP]))) < 1e-10
        assert sol.minimum < 1e-18


def test_get_aligned_data_reference_known_answer(built_lib, tmp_path):
    """the reference's own test (test/aligned_data.jl:6-15): a two-curve file with a ragged second curve and trailing
    empty rows; expected times, data, antigen concentration (from the file name) and antibody concentrations (header)"""
    from sprb200 import api
    times = [[5.0, 6.0, 7.0, 8.0, 9.0, 10.0, 11.0], [5.0, 6.0, 7.0, 8.0]]
    refdata = [[.821, 1.12, 1.22, 1.52, 1.42, 1.82, 1.92], [3.80, 4.30, 4.80, 5.50]]
    fn = tmp_path / "tester-13.8_aligned.csv"
    with open(fn, "w") as f:
        f.write("time (s),25,time (s),100\n")
        for k in range(7):
            second = "%.2E,%.2E" % (times[1][k], refdata[1][k]) if k < 4 else ","
            f.write("%.2E,%.2E,%s\n" % (times[0][k], refdata[0][k], second))
        f.write(",,,\n" * 5)
    ad = api.get_aligned_data(str(fn))
    assert [list(t) for t in ad.times] == times
    assert [list(d) for d in ad.refdata] == refdata
    assert ad.antigenconcen == 13.8
    assert list(ad.antibodyconcens) == [25.0, 100.0]


def test_savefit_workbook_is_read_by_the_golden_extractor(built_lib, tmp_path):
    """fitting.jl:320-431: api.savefit writes the reference's two-sheet workbook; the extractor that reads the
    reference's own recorded workbook into tests/golden/ (oracle/make_golden.py) reads it back identically"""
    from oracle import make_golden
    from sprb200 import api
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "fit_curves_golden.json")))
    t = np.array(gold["times"])
    ad = api.AlignedData([t, t], gold["data"], gold["antibody_nM"], 13.8)
    ip = gold["internal_params"]
    sol = api.OptSolution(np.array([ip["logkon"], ip["logkoff"], ip["logkonb"], ip["reach"], ip["logCP"]]), gold["fitness"], 0, 0)
    surpars = api.SurrogateParams((-5.0, 2.0), (-4.0, 0.0), (-3.0, 1.5), (2.0, 35.0))
    simpars = api.SimParams(tstop=600.0, tstop_AtoB=150.0, tsave=t)
    sur = api.Surrogate(np.zeros((2, 2, 2, 2, 2)), surpars=surpars, simpars=simpars)
    fn = api.savefit(sol, ad, sur, simpars, str(tmp_path / "parameters.xlsx"), curves=gold["sim_mean"])
    assert fn.endswith("parameters.xlsx_fit.xlsx")
    back = make_golden.extract(fn)
    for k in ("antibody_nM", "internal_params", "fitness", "times", "data", "sim_mean"):
        assert back[k] == gold[k], k
    for k, v in gold["physical_params"].items():                  # the reference's own conversion, to rounding
        assert math.isclose(back["physical_params"][k], v, rel_tol=1e-12), k
    # with a monovalent fit: four columns per concentration and the extra parameter block
    mono = api.OptSolution(np.log10([1e-4, 0.05, 100.0]), 123.0, 1, 1)
    fn2 = api.savefit(sol, ad, sur, simpars, str(tmp_path / "with_mono"), monofit=mono, curves=gold["sim_mean"])
    import zipfile
    z = zipfile.ZipFile(fn2)
    shared = z.read("xl/sharedStrings.xml").decode()
    assert "Monovalent Fit 25.0 nM" in shared and "Monovalent fit parameters (physical):" in shared and "Monofit return code:" in shared
    assert 'r="H587"' in z.read("xl/worksheets/sheet1.xml").decode()
    assert "SPR and Fit Curves" in z.read("xl/workbook.xml").decode() and "Fit Parameters" in z.read("xl/workbook.xml").decode()
