"""-m gpu: statistical parity of the CUDA path with the reference, through the C ABI.

North-star criterion: the GPU mean bound curve lies within 3 standard errors of the reference at every
time bin, >= 10^4 replicates.  "The reference" is, in decreasing order of authority:
  (1) the reference's own recorded curves (tests/golden/fit_curves_golden.json, 250 replicates),
  (2) exact answers derived from its formulas (N=2 chain, monovalent closed form),
  (3) oracle/spr_ref.c, the CPU restatement of its algorithm (independent RNG: a two-sample test).
z = (m_gpu - m_ref) / sqrt(SE_gpu^2 + SE_ref^2) per bin.

NO SEED IS SELECTED.  Every seed derives from ONE base seed (env SPRB200_PARITY_SEED, default 20231); the
assertions are built to hold for any base seed with a false-alarm probability below 1 % per case while a
bias of 0.3 % of the curve level fails them (profiles/r2_parity_null.txt holds the measured null).  Why not
"max|z| < 3 for one seed pair": over ~600 strongly autocorrelated bins that bound is exceeded by pure chance
for about one seed pair in six (measured: 4 of 24 CPU-vs-CPU pairs), so a single pair can only pass it with
hand-picked seeds.  Instead each case runs K = 4 independent (GPU sample, oracle sample) pairs, >= 10^4 GPU
replicates and >= 10^4 (C1) / 2x10^3 (stress cases) oracle replicates per pair, and asserts

  A. pooled curve (all pairs): max_t |z_pool(t)| < 4.75 -- Bonferroni over <= 601 bins: P < 0.13 % whatever
     the correlation; with 4 x 10^4 replicates per side a 0.3 % bias is a 6-sigma excursion;
  B. |mean of z over all pairs and bins| < 4 / sqrt(K) = 2: the mean of unit-variance variables has variance
     <= 1 whatever their correlation, so this is a >= 4-sigma event under the null; a uniform bias of one SE
     of a single pair (0.15 % of the level) already gives mean z = 1 per pair;
  C. mean of z^2 over all pairs and bins < 3.5 (null expectation 1; even if every curve moved as ONE degree of
     freedom the sum over 4 pairs is chi^2_4 and P(> 14) = 0.7 %);
  D. the north-star rule itself in a seed-free form: a real bias pushes EVERY pair beyond 3 SE, chance pushes one
     pair in six, so at least one of the K pairs must lie within 3 SE at every bin (null P(all 4 exceed) =
     0.1 %), and no pair may exceed 5 SE anywhere (Bonferroni 0.03 % per pair);
  E. events per trajectory agree within 0.5 % (same process => same event count; SE ~ 0.05 %);
  F. second moments: the replicate-to-replicate variance summed over the bins and pairs, GPU / oracle, within 10 % of 1
     (what `vars`/`sems` of the outputter and the VarianceTerminator's stopping index are made of; the null scatter of
     this ratio is 0.7 % with 10^4 and 2 % with 2x10^3 oracle replicates per pair -- tests/tools/parity_replay_cpu.py
     replays this test on the CPU with the bit-identical mirror).
"""
import json
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDF = os.path.join(os.path.dirname(__file__), "golden", "fit_curves_golden.json")
NREP_GPU = 10000
K_PAIRS = 4
BASE_SEED = int(os.environ.get("SPRB200_PARITY_SEED", "20231"))
ORACLE_STREAMS = 32          # the oracle sample of a pair is 32 independent streams, whatever the core count


def seeds_of(case_index, k):
    """(GPU seed, oracle seed) of pair k of a case: a fixed function of the base seed, nothing chosen"""
    return BASE_SEED * 100003 + 64 * case_index + k, BASE_SEED * 100003 + 64 * case_index + 32 + k


def gpu_moments(engine, sim, point, nsims, seed, N, pidx=0, obs=0):
    import sprb200
    out = engine.simulate(sim, [point], sprb200.make_run(nsims=nsims, seed=seed, param_index_base=pidx, observable=obs))
    n = int(out["n_used"][0])
    mean = out["sum"][0] / n / N
    var = (out["sumsq"][0].astype(float) - out["sum"][0].astype(float) ** 2 / n) / (n - 1) / N ** 2
    return mean, var, n, int(out["n_events"][0])


def ref_moments(point, N, DIM, L, tstop, toff, tsave, nrep, seed, copies=None, observable=0):
    """reference-algorithm oracle, replicates split over `copies` independent streams run on all host cores"""
    from oracle import ref_oracle as R
    copies = copies or ORACLE_STREAMS
    per = max(2, -(-nrep // copies))
    sim = R.make_sim(N=N, DIM=DIM, tstop=tstop, tstop_AtoB=toff, tsave=tsave, L=L, nsims=per)

    def one(c):
        return R.run(*point, sim, seed=seed, stream=1000 + c, observable=observable)
    with ThreadPoolExecutor(min(copies, os.cpu_count() or 1)) as ex:
        res = list(ex.map(one, range(copies)))
    n = per * copies
    m = np.mean([r["mean"] for r in res], axis=0)
    # pooled unbiased variance from per-copy mean/variance
    ss = sum((per - 1) * r["var"] + per * (r["mean"] - m) ** 2 for r in res)
    v = ss / (n - 1)
    ev = sum(r["nevents"] for r in res)
    return m, v, n, ev


def parity_statistics(pairs):
    """pairs: list of (gm, gv, gn, rm, rv, rn).  Returns the statistics A-D of the module docstring."""
    num = np.zeros_like(pairs[0][0])
    den = np.zeros_like(pairs[0][0])
    zs, maxabs = [], []
    gvsum = rvsum = 0.0
    for gm, gv, gn, rm, rv, rn in pairs:
        gvsum += float(np.sum(gv))
        rvsum += float(np.sum(rv))
        se2 = gv / gn + rv / rn
        ok = se2 > 0
        assert np.all(gm[~ok] == rm[~ok])          # deterministic bins (t = 0) must agree exactly
        z = (gm - rm)[ok] / np.sqrt(se2[ok])
        zs.append(z)
        maxabs.append(float(np.abs(z).max()))
        num += gm - rm
        den += se2
    okp = den > 0
    zpool = num[okp] / np.sqrt(den[okp])
    allz = np.concatenate(zs)
    return dict(pool_max=float(np.abs(zpool).max()), mean_z=float(allz.mean()), mean_z2=float((allz ** 2).mean()),
                pair_max=maxabs, n_within_3=int(sum(m < 3.0 for m in maxabs)),
                frac_bins_beyond_3=float((np.abs(allz) > 3.0).mean()),
                var_ratio=gvsum / rvsum if rvsum > 0 else 1.0)


def assert_parity(st, K, label):
    assert st["pool_max"] < 4.75, (label, st)                       # A
    assert abs(st["mean_z"]) < 4.0 / np.sqrt(K), (label, st)        # B
    assert st["mean_z2"] < 3.5, (label, st)                         # C
    assert st["n_within_3"] >= 1 and max(st["pair_max"]) < 5.0, (label, st)   # D
    assert abs(st["var_ratio"] - 1.0) < 0.10, (label, st)           # F


@pytest.fixture(scope="module")
def gold():
    return json.load(open(GOLDF))


@pytest.mark.parametrize("j", [0, 1])
def test_gpu_vs_reference_recorded_curves(engine, gold, j):
    """the ONE simulator output recorded in the reference tree (250 replicates per curve,
    docs/src/fitting_data/parameters.xlsx_fit.xlsx).  The reference sample is fixed, so only the GPU seed
    varies: K seeds, each within the Bonferroni bound 4.1 over 586 bins of the recorded curve (SE of the
    recorded mean estimated from the GPU variance / 250), and their average z^2 < 3."""
    import sprb200
    from sprb200 import _lib
    p = gold["internal_params"]
    times = np.array(gold["times"])
    CP = 10 ** p["logCP"]
    kon = 10 ** (p["logkon"] + np.log10(gold["antibody_nM"][j] / gold["antibody_nM"][0]))
    sim = sprb200.SimSpec(1000, 3, _lib.domain_length(), times[-1], 150.0, times)
    z2 = []
    for k in range(K_PAIRS):
        gseed, _ = seeds_of(20 + j, k)
        mean, var, n, _ = gpu_moments(engine, sim, [kon, 10 ** p["logkoff"], 10 ** p["logkonb"], p["reach"]], NREP_GPU,
                                      gseed, 1000, j)
        z = (CP * mean - np.array(gold["sim_mean"][j])) / (CP * np.sqrt(var / n + var / 250))
        assert np.abs(z).max() < 4.1, (k, np.abs(z).max())
        z2.append((z ** 2).mean())
    assert np.mean(z2) < 3.0, z2


def test_gpu_two_particle_chain_exact(engine):
    import sprb200
    from oracle import exact
    locs = np.array([[1.0, 1.0, 1.0], [2.0, 1.0, 1.0]])
    tsave = np.arange(0.0, 61.0, 1.0)
    sim = sprb200.SimSpec(2, 3, 100.0, 60.0, 30.0, tsave, resample_initlocs=False, initlocs=locs)
    for obs, col in ((0, 0), (1, 1)):
        mean, var, n, _ = gpu_moments(engine, sim, [0.05, 0.1, 0.5, 5.0], 200000, seeds_of(30, obs)[0], 1, obs=obs)
        ex = exact.two_particle_ctmc(0.05, 0.1, 0.5, 30.0, tsave)[col]
        z = (mean - ex)[1:] / np.sqrt(var[1:] / n)
        assert np.abs(z).max() < 4.2, (obs, np.abs(z).max())       # Bonferroni over 60 bins: P < 0.2 %
    # far apart: no pair channel, each particle is monovalent
    far = np.array([[1.0, 1.0, 1.0], [40.0, 1.0, 1.0]])
    sim = sprb200.SimSpec(2, 3, 100.0, 60.0, 30.0, tsave, resample_initlocs=False, initlocs=far)
    mean, var, n, _ = gpu_moments(engine, sim, [0.05, 0.1, 0.5, 5.0], 100000, seeds_of(30, 2)[0], 2)
    ex = exact.monovalent_total_bound(0.05, 0.1, 1.0, 1.0, 30.0, tsave)
    z = (mean - ex)[1:] / np.sqrt(var[1:] / n)
    assert np.abs(z).max() < 4.2


def test_gpu_monovalent_limit(engine):
    """konb = 0 (and separately reach = 0): E[(B+C)/N] = monovalent_total_bound, Var = p(1-p)/N"""
    import sprb200
    from sprb200 import _lib
    from oracle import exact
    tsave = np.arange(0.0, 601.0, 1.0)
    sim = sprb200.SimSpec(1000, 3, _lib.domain_length(), 600.0, 150.0, tsave)
    ex = exact.monovalent_total_bound(0.01, 0.02, 1.0, 1.0, 150.0, tsave)
    for c, point in enumerate(([0.01, 0.02, 0.0, 30.0], [0.01, 0.02, 5.0, 0.0])):
        mean, var, n, _ = gpu_moments(engine, sim, point, NREP_GPU, seeds_of(31, c)[0], 1000)
        z = (mean - ex)[1:] / np.sqrt(ex[1:] * (1 - ex[1:]) / 1000 / n)
        assert np.abs(z).max() < 4.75                                 # Bonferroni over 600 bins
        assert np.allclose(var[50:], ex[50:] * (1 - ex[50:]) / 1000, rtol=0.1)


CASES = {
    # name: (kon', koff, konb, reach, N, DIM, L, tstop, toff, tsave, oracle replicates per pair)
    "C1-golden-25nM": (10 ** -2.8789331867029184, 10 ** -1.3886105070021626, 10 ** -0.107804350729586,
                       15.292951569975912, 1000, 3, None, 600.0, 150.0, np.arange(0.0, 601.0, 1.0), 10000),
    "C1-default-reach30-100nM": (4 * 10 ** -2.8789331867029184, 10 ** -1.3886105070021626, 10 ** -0.107804350729586,
                                 30.0, 1000, 3, None, 600.0, 150.0, np.arange(0.0, 601.0, 1.0), 10000),
    "C4a-stress-reach50": (1.0, 0.1, 1.0, 50.0, 1000, 3, None, 600.0, 150.0, np.arange(0.0, 601.0, 1.0), 2000),
    "C4b-stress-2x-density": (1.0, 0.1, 1.0, 50.0, 1000, 3, 187.86, 300.0, 150.0, np.arange(0.0, 301.0, 1.0), 2000),
    "2D-surface": (0.05, 0.02, 0.5, 25.0, 600, 2, None, 400.0, 100.0, np.arange(0.0, 401.0, 2.0), 10000),
    "high-rate-corner": (100.0, 1.0, 31.6, 35.0, 1000, 3, None, 60.0, 20.0, np.arange(0.0, 61.0, 0.5), 2000),
}


@pytest.mark.parametrize("name", list(CASES))
def test_gpu_parity_with_reference_algorithm(engine, oracle_libs, name):
    """statistics A-E of the module docstring on K = 4 seed pairs; seeds derive from the base seed"""
    import sprb200
    from sprb200 import _lib
    kon, koff, konb, reach, N, DIM, L, tstop, toff, tsave, nref = CASES[name]
    L = _lib.domain_length(N, DIM) if L is None else L
    sim = sprb200.SimSpec(N, DIM, L, tstop, toff, tsave)
    ci = list(CASES).index(name)
    pairs, gev, rev = [], [], []
    for k in range(K_PAIRS):
        gseed, rseed = seeds_of(ci, k)
        gm, gv, gn, ge = gpu_moments(engine, sim, [kon, koff, konb, reach], NREP_GPU, gseed, N, 11)
        rm, rv, rn, re_ = ref_moments([kon, koff, konb, reach], N, DIM, L, tstop, toff, tsave, nref, seed=rseed)
        assert gn >= 10000 and rn >= nref
        pairs.append((gm, gv, gn, rm, rv, rn))
        gev.append(ge / gn)
        rev.append(re_ / rn)
    st = parity_statistics(pairs)
    print("\nparity %-26s base seed %d: %s" % (name, BASE_SEED, json.dumps(st)))
    assert_parity(st, K_PAIRS, name)
    assert np.mean(gev) == pytest.approx(np.mean(rev), rel=0.005)    # E


def test_gpu_total_A_observable(engine, oracle_libs):
    """TotalAOutputter (callbacks.jl:111): A = N - B - 2C; same statistics on the A curve"""
    import sprb200
    from sprb200 import _lib
    tsave = np.arange(0.0, 201.0, 2.0)
    L = _lib.domain_length(500, 3)
    sim = sprb200.SimSpec(500, 3, L, 200.0, 80.0, tsave)
    pairs = []
    for k in range(K_PAIRS):
        gseed, rseed = seeds_of(40, k)
        gm, gv, gn, _ = gpu_moments(engine, sim, [0.05, 0.03, 0.6, 28.0], 5000, gseed, 500, obs=1)
        rm, rv, rn, _ = ref_moments([0.05, 0.03, 0.6, 28.0], 500, 3, L, 200.0, 80.0, tsave, 3000, seed=rseed, observable=1)
        assert gm[0] == 1.0
        pairs.append((gm, gv, gn, rm, rv, rn))
    st = parity_statistics(pairs)
    print("\nparity TotalA: %s" % json.dumps(st))
    assert_parity(st, K_PAIRS, "TotalA")


# ------------------------------------------------------------------------------------------------------------
# BASELINE configs 2 and 3: seeded samples of the C2 grid (examples/make_surrogate.jl) and of the paper-scale
# C3 grid (examples/cluster_scripts/make_surrogate_metadata.jl), built through sprb200_build_indices exactly as
# a surrogate build does (RNG parameter index = linear index - 1), against oracle/spr_ref.c.
GRIDS = {
    "C2": dict(ranges=((-3.0, 2.0), (-4.0, -1.0), (-3.0, 1.5), (2.0, 35.0)), size=(10, 10, 10, 10)),
    "C3": dict(ranges=((-5.0, 2.0), (-4.0, 0.0), (-3.0, 1.5), (2.0, 35.0)), size=(42, 40, 30, 30)),
}
NGRID_POINTS = 24


def grid_sample(name):
    """24 grid indices: the two extreme corners plus 22 drawn uniformly from the base seed"""
    g = GRIDS[name]
    P = int(np.prod(g["size"]))
    rng = np.random.default_rng(BASE_SEED + (2 if name == "C2" else 3))
    idx = np.concatenate([[1, P], rng.choice(np.arange(2, P), size=NGRID_POINTS - 2, replace=False)])
    return np.sort(idx).astype(np.int64)


def oracle_points(pts, N, L, tsave, nreps, seed, term=None):
    """mean/variance of every point from the oracle: point k gets nreps[k] replicates split over 8 streams"""
    from oracle import ref_oracle as R
    jobs = [(k, c) for k in range(len(pts)) for c in range(8)]

    def one(job):
        k, c = job
        per = max(2, -(-int(nreps[k]) // 8))
        sim = R.make_sim(N=N, DIM=3, tstop=600.0, tstop_AtoB=150.0, tsave=tsave, L=L, nsims=per)
        return R.run(*pts[k], sim, seed=seed + k, stream=c, term=term)
    with ThreadPoolExecutor(os.cpu_count() or 1) as ex:
        res = list(ex.map(one, jobs))
    out = []
    for k in range(len(pts)):
        rs = res[8 * k:8 * k + 8]
        per = rs[0]["n"]
        n = per * 8
        m = np.mean([r["mean"] for r in rs], axis=0)
        v = sum((per - 1) * r["var"] + per * (r["mean"] - m) ** 2 for r in rs) / (n - 1)
        out.append((m, v, n, sum(r["nevents"] for r in rs) / n))
    return out


@pytest.mark.parametrize("gname", ["C2", "C3"])
def test_gpu_grid_sample_parity(engine, oracle_libs, gname):
    """24 seeded points of the grid, FIXED 2000 GPU replicates per point against the oracle (replicates per
    point scaled to its cost, >= 200).  Per point the same z statistics; over the sample: no point beyond 5 SE
    anywhere, the points beyond 3 SE somewhere are a minority (chance: one in six), mean z^2 < 2 and
    |mean z| < 4/sqrt(24) over all points and bins, events per trajectory within 1 % in total."""
    import sprb200
    from sprb200 import _lib
    g = GRIDS[gname]
    grid = _lib.make_grid(*g["ranges"], g["size"])
    idx = grid_sample(gname)
    pts = _lib.grid_points(grid, idx)
    L = _lib.domain_length()
    tsave = np.linspace(0.0, 600.0, 601)
    sim = sprb200.SimSpec(1000, 3, L, 600.0, 150.0, tsave)
    NG = 2000
    run = sprb200.make_run(nsims=NG, seed=BASE_SEED + 7)
    # the moments of the same streams the slice builder uses: parameter index = linear index - 1
    gm, gv, gev = [], [], []
    rows, nu, nev = engine.build_indices(grid, sim, run, idx)
    assert np.all(nu == NG)
    for k, i in enumerate(idx):
        out = engine.simulate(sim, [pts[k]], sprb200.make_run(nsims=NG, seed=BASE_SEED + 7, param_index_base=int(i) - 1))
        assert np.array_equal(rows[k], out["mean"][0])               # build == simulate, bit for bit
        s1, s2 = out["sum"][0].astype(float), out["sumsq"][0].astype(float)
        gm.append(out["mean"][0]); gv.append((s2 - s1 ** 2 / NG) / (NG - 1) / 1000.0 ** 2); gev.append(int(out["n_events"][0]) / NG)
    # oracle replicates per point: about 6 core-seconds each, 200..2000
    cost = np.array(gev) / 1.0e6 + 0.003                             # core-seconds per replicate (oracle: ~1e6 events/s + setup)
    nreps = np.clip((6.0 / cost).astype(int), 200, 2000)
    ref = oracle_points(pts, 1000, L, tsave, nreps, seed=BASE_SEED * 7 + 1)
    allz, pmax = [], []
    for k in range(len(idx)):
        rm, rv, rn, _ = ref[k]
        se2 = gv[k] / NG + rv / rn
        assert np.all(gm[k][se2 == 0] == rm[se2 == 0])
        # the z statistic is Gaussian only where both samples hold enough bound particles in total (the grid
        # reaches kon' = 1e-5: a handful of binding events per thousand replicates): >= 50 counts on either side
        ok = (se2 > 0) & (gm[k] * 1000.0 * NG >= 50.0) & (rm * 1000.0 * rn >= 50.0)
        z = (gm[k] - rm)[ok] / np.sqrt(se2[ok])
        allz.append(z)
        pmax.append(float(np.abs(z).max()) if len(z) else 0.0)
    allz = np.concatenate(allz)
    st = dict(worst=max(pmax), n_beyond_3=int(sum(m >= 3.0 for m in pmax)), mean_z=float(allz.mean()),
              mean_z2=float((allz ** 2).mean()))
    print("\ngrid parity %s, %d points, base seed %d: %s" % (gname, len(idx), BASE_SEED, json.dumps(st)))
    assert st["worst"] < 5.0, (st, pmax)
    assert st["n_beyond_3"] <= 10, (st, pmax)           # Binomial(24, 1/6): P(> 10) < 0.1 %
    assert abs(st["mean_z"]) < 4.0 / np.sqrt(len(idx)), st
    assert st["mean_z2"] < 2.0, st
    rev = np.array([r[3] for r in ref])
    w = np.array(gev) > 50                                # events per trajectory, point by point (SE ~ 1/sqrt(n ev))
    assert np.allclose(np.array(gev)[w], rev[w], rtol=0.05), (np.array(gev)[w] / rev[w])
    assert np.sum(gev) == pytest.approx(np.sum(rev), rel=0.01)


@pytest.mark.parametrize("gname", ["C2", "C3"])
def test_gpu_grid_sample_variance_terminator(engine, oracle_libs, gname):
    """the reference's default VarianceTerminator(0.01, 15, 250) on the same grid samples: the number of
    replicates each point uses (callbacks.jl:204-217) is a random stopping time, compared in distribution:
    points that always run to maxsims or always stop at minsims must agree exactly with the oracle, the total
    over the sample within 8 %, the points in between within a factor 3 each (median factor < 1.6); mean curves
    within 5 SE."""
    import sprb200
    from sprb200 import _lib
    from oracle import ref_oracle as R
    g = GRIDS[gname]
    grid = _lib.make_grid(*g["ranges"], g["size"])
    idx = grid_sample(gname)
    pts = _lib.grid_points(grid, idx)
    L = _lib.domain_length()
    tsave = np.linspace(0.0, 600.0, 601)
    sim = sprb200.SimSpec(1000, 3, L, 600.0, 150.0, tsave)
    run = sprb200.make_run(variance=(0.01, 15, 250), seed=BASE_SEED + 11)
    rows, nu, nev = engine.build_indices(grid, sim, run, idx)
    term = R.make_term("variance", 0.01, 15, 250)
    rsim = R.make_sim(N=1000, DIM=3, tstop=600.0, tstop_AtoB=150.0, tsave=tsave, L=L)

    def one(k):
        return R.run(*pts[k], rsim, term=term, seed=BASE_SEED * 11 + 3, stream=int(idx[k]))
    with ThreadPoolExecutor(os.cpu_count() or 1) as ex:
        ref = list(ex.map(one, range(len(idx))))
    rn = np.array([r["n"] for r in ref])
    print("\nn_used %s gpu %s\n       %s ref %s" % (gname, nu.tolist(), gname, rn.tolist()))
    assert nu.min() >= 15 and nu.max() <= 250
    # a point whose relative SEM is far from the 1 % threshold stops at the same n in any sample
    sure_max = (nu == 250) & (rn == 250)
    sure_min = (nu == 15) & (rn == 15)
    assert np.sum((nu == 250) != (rn == 250)) <= 3 and np.sum((nu == 15) != (rn == 15)) <= 3
    assert nu.sum() == pytest.approx(rn.sum(), rel=0.08)
    # in between, the stopping index is a first-passage time with a broad distribution (the variance estimate of 20-100
    # replicates is itself 15-30 % noisy): two independent samples agree within a factor, not in rank
    mid = ~(sure_max | sure_min)
    if mid.any():
        ratio = nu[mid].astype(float) / rn[mid].astype(float)
        assert np.all((ratio > 1.0 / 3.0) & (ratio < 3.0)), (nu[mid].tolist(), rn[mid].tolist())
        assert abs(np.median(np.log(ratio))) < np.log(1.6), (nu[mid].tolist(), rn[mid].tolist())
    # the means of those adaptive samples agree as well (SE from the oracle's own variance estimate)
    # (a Gaussian bound needs enough replicates for the variance estimate and enough counts per bin: points that
    # stop near minsims = 15 are Student-t with ~14 degrees of freedom, bound 9 instead of 5)
    for k in range(len(idx)):
        se2 = ref[k]["var"] / rn[k] * (1.0 + rn[k] / max(nu[k], 1))
        ok = (se2 > 0) & (ref[k]["mean"] * 1000.0 * rn[k] >= 50.0) & (rows[k] * 1000.0 * nu[k] >= 50.0)
        if not ok.any():
            continue
        z = (rows[k] - ref[k]["mean"])[ok] / np.sqrt(se2[ok])
        assert np.abs(z).max() < (5.0 if min(nu[k], rn[k]) >= 100 else 9.0), (k, idx[k], np.abs(z).max())
