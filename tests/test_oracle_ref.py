"""CPU: pins oracle/spr_ref.c (the restatement of the reference algorithm) against everything the
reference tree offers for this path: its recorded mean curves (docs xlsx -> tests/golden), the exact
N=2 chain, the monovalent closed form (monovalent_fitting.jl:22-28), its own terminator property test
(test/callbacks.jl:21-37) and periodic_dist_sq."""
import json
import os

import numpy as np
import pytest

from oracle import exact
from oracle import ref_oracle as R

GOLD = os.path.join(os.path.dirname(__file__), "golden", "fit_curves_golden.json")


@pytest.fixture(scope="module")
def gold():
    return json.load(open(GOLD))


def test_golden_fixture_is_the_reference_file(gold):
    assert gold["source_sha256"].startswith("64078c2f")
    assert len(gold["times"]) == 586 and gold["times"][0] == 5.0 and gold["times"][-1] == 599.0
    # value * nsims * N / CP is an integer count (means of 250 replicates of integer (B+C))
    CP = 10 ** gold["internal_params"]["logCP"]
    for j in range(2):
        x = np.array(gold["sim_mean"][j]) / CP * 250 * 1000
        assert np.abs(x - np.round(x)).max() < 1e-6


@pytest.mark.parametrize("j", [0, 1])
def test_oracle_matches_reference_golden_curves(oracle_libs, gold, j):
    """two-sample z test per time bin, oracle (400 replicates) vs the reference's recorded means (250)"""
    p = gold["internal_params"]
    times = np.array(gold["times"])
    CP = 10 ** p["logCP"]
    kon = 10 ** (p["logkon"] + np.log10(gold["antibody_nM"][j] / gold["antibody_nM"][0]))
    sim = R.make_sim(tstop=times[-1], tstop_AtoB=150.0, tsave=times, nsims=400)
    o = R.run(kon, 10 ** p["logkoff"], 10 ** p["logkonb"], p["reach"], sim, CP=CP, seed=11 + j)
    z = (o["mean"] - np.array(gold["sim_mean"][j])) / np.sqrt(o["var"] / o["n"] + o["var"] / 250)
    assert np.abs(z).max() < 4.0
    assert (z ** 2).mean() < 2.5          # bins are strongly autocorrelated; 1.0 expected


def test_oracle_two_particle_chain_exact(oracle_libs):
    tsave = np.arange(0, 61.0, 1.0)
    locs = np.array([[1.0, 1.0, 1.0], [2.0, 1.0, 1.0]])
    sim = R.make_sim(N=2, DIM=3, tstop=60.0, tstop_AtoB=30.0, tsave=tsave, L=100.0, initlocs=locs,
                     resample_initlocs=False, nsims=100000)
    for obs, col in ((0, 0), (1, 1)):
        o = R.run(0.05, 0.1, 0.5, 5.0, sim, CP=2.0, seed=7, observable=obs)   # CP/N = 1
        ex = exact.two_particle_ctmc(0.05, 0.1, 0.5, 30.0, tsave)[col]
        se = np.sqrt(o["var"] / o["n"])
        z = (o["mean"] - ex)[1:] / se[1:]
        assert np.abs(z).max() < 4.0


def test_oracle_monovalent_limit(oracle_libs):
    """reach = 0: independent particles, E[(B+C)/N] = monovalent_total_bound, Var = p(1-p)/N"""
    sim = R.make_sim(tstop=600.0, tstop_AtoB=150.0, nsims=100)
    o = R.run(0.01, 0.02, 0.0001, 0.0, sim, seed=3)
    ex = exact.monovalent_total_bound(0.01, 0.02, 1.0, 1.0, 150.0, sim.tsave)
    z = (o["mean"] - ex)[1:] / np.sqrt(ex[1:] * (1 - ex[1:]) / 1000 / o["n"])
    assert np.abs(z).max() < 4.0


def test_oracle_variance_terminator_property(oracle_libs):
    """the reference's own test (test/callbacks.jl:21-37)"""
    sim = R.make_sim(tstop_AtoB=150.0, dt=1.0)
    term = R.make_term("variance", ssetol=0.025, minsims=15, maxsims=4000)
    o = R.run(0.001, 0.01, 0.01, 10.0, sim, term=term, seed=5)
    sem = np.sqrt(o["var"] / o["n"])
    assert np.all(sem <= 0.025 * o["mean"] + 1e-15) or o["n"] == 4000
    assert 15 <= o["n"] <= 4000


def test_oracle_raw_copynumbers_and_save_semantics(oracle_libs):
    sim = R.make_sim(N=50, tstop=20.0, tstop_AtoB=10.0, tsave=np.array([0.0, 1.0, 5.0, 10.0, 19.0, 25.0, 30.0]),
                     nsims=20)
    o = R.run(0.3, 0.2, 1.0, 60.0, sim, seed=1, want_raw=True)
    raw = o["raw"]
    assert raw.shape == (20, 3, 7)
    assert np.all(raw[:, 0] + raw[:, 1] + 2 * raw[:, 2] == 50)     # A + B + 2C = N
    assert np.all(raw[:, :, 0] == np.array([50, 0, 0]))            # t = 0 slot: everything A
    # mean over replicates of (B+C)/N equals the outputter mean
    assert np.allclose((raw[:, 1] + raw[:, 2]).mean(axis=0) / 50, o["mean"], rtol=1e-12)


def test_periodic_dist_sq(oracle_libs):
    import ctypes as C
    L = 10.0
    a = np.array([0.5, 9.5, 5.0]); b = np.array([9.5, 0.5, 5.0])
    d = R.lib().spr_ref_periodic_dist_sq(a.ctypes.data_as(C.POINTER(C.c_double)),
                                         b.ctypes.data_as(C.POINTER(C.c_double)), 3, L)
    assert d == pytest.approx(2.0)
    assert R.default_L() == pytest.approx(236.68625663347206, rel=1e-12)


def test_build_slice_index_order(oracle_libs):
    """linear index -> (kon fastest ... reach slowest), surrogate.jl:284-288"""
    g = R.make_grid((-3, 2), (-4, -1), (-3, 1.5), (2, 35), (2, 3, 2, 2))
    sim = R.make_sim(N=20, tstop=5.0, tstop_AtoB=2.0, tsave=np.array([0.0, 2.0, 5.0]), nsims=3)
    out, nobs, nev = R.build_slice(g, sim, R.make_term("number"), 1, 24, seed=1, nthreads=2)
    assert out.shape == (24, 3) and np.all(nobs == 3)
    assert np.all(out[:, 0] == 0.0)
    # kon is the fastest axis: index 2 (kon high) binds more than index 1 (kon low) at equal other parameters
    assert out[1, 1] > out[0, 1]
