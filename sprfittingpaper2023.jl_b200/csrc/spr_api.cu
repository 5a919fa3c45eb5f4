// spr_api.cu -- the C ABI of libsprb200.so (include/sprb200.h): argument checking, work
// partitioning (shared-memory capacity classes by reach, expensive-first ticket order), replicate
// staging for the VarianceTerminator, device arenas and streams.  No CPU fallback: every entry
// point that simulates needs a CUDA device and fails with SPRB200_ERR_NO_DEVICE otherwise.
#include "../../include/sprb200.h"
#include "spr_device.cuh"
#include "spr_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace spr;

namespace {

constexpr int NSTREAM = 1;   // every launch of a handle is ordered on one non-default stream
thread_local std::string g_create_error;

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            want = bytes;
            e = cudaMalloc(&p, want);
        }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T *as() const { return static_cast<T *>(p); }
};

}  // namespace

struct sprb200_ctx {
    int device = 0;
    int num_sms = 0;
    int smem_optin = 0;      // largest dynamic shared memory of one CTA
    int smem_per_sm = 0;     // shared memory of one SM (every resident CTA also reserves 1 KB of it)
    cudaStream_t stream[NSTREAM] = {};
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    std::string err;
    // arenas
    DevBuf tsave, pts, pidx, order, curve, curve2, sum, sumsq, nused, nevents, nscanned, nstar, active, counters,
        rows, locs, goff, gnbr, misc, table, arena, desc, defer, tickets, slotrec, stage, sur, fitdata, fitpars, fitloss,
        mrows, mtable, midx, mnused, mnev, gnused, gnev;
    // multi-GPU (sprb200_comm_init): NCCL communicator of this rank, dlopen'ed libnccl
    void *nccl_comm = nullptr;
    int comm_rank = 0, comm_nranks = 1;
    // stats of the last call
    uint64_t st_traj = 0, st_events = 0, st_launches = 0;
    double st_ms = 0.0, st_table_ms = 0.0, st_traj_ms = 0.0;
    std::vector<cudaEvent_t> evpool;          // timing events of the K2 / K1 launches of the last call
    size_t evused = 0;
    cudaEvent_t next_event() {
        if (evused == evpool.size()) {
            cudaEvent_t e = nullptr;
            if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
            evpool.push_back(e);
        }
        return evpool[evused++];
    }
    size_t curve_budget = (size_t)8 << 30;    // SPRB200_CURVE_BYTES: replicate curves resident per batch
    size_t table_budget = (size_t)16 << 30;   // SPRB200_TABLE_BYTES: neighbour-table arena (K2 -> K1)
    double est_scale = 1.0;   // SPRB200_CAP_SCALE (test hook): scales the a-priori table size used for chunking
    int max_warps = 0;        // SPRB200_MAX_WARPS (experiment hook): cap on resident trajectories per SM
    // resident surrogate (sprb200_surrogate_load)
    SprGrid sur_grid = {};
    int sur_T = 0;
    double sur_t0 = 0.0, sur_t1 = 0.0;
    long long sur_P = 0;
    int stage_cap_override = 0;   // SPRB200_STAGE_CAP (test hook): staging stride of K2's distance pass
    int use_smem_table = 1;       // SPRB200_SMEM_TABLE=0 disables the rows-in-shared-memory variant of underfilled launches
    int use_tmem = 1;             // SPRB200_TMEM=0 disables the tensor-memory variant of K1 (row offsets in TMEM, 32 warps per SM)
};

namespace {

int fail(sprb200_ctx *h, int code, const std::string &msg) {
    if (h) h->err = msg;
    else g_create_error = msg;
    return code;
}
#define CUDA_TRY(h, expr)                                                                          \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            cudaGetLastError();                                                                    \
            return fail(h, e__ == cudaErrorMemoryAllocation ? SPRB200_ERR_ALLOC : SPRB200_ERR_CUDA, \
                        std::string(#expr) + ": " + cudaGetErrorString(e__));                      \
        }                                                                                          \
    } while (0)

int check_sim(sprb200_ctx *h, const SprSim *sim) {
    if (!sim) return fail(h, SPRB200_ERR_BAD_ARG, "sim is NULL");
    if (sim->N < 1 || sim->N > 65535) return fail(h, SPRB200_ERR_BAD_ARG, "N must be in 1..65535");
    if (sim->DIM != 2 && sim->DIM != 3) return fail(h, SPRB200_ERR_BAD_ARG, "DIM must be 2 or 3");
    if (!(sim->L > 0.0) || !std::isfinite(sim->L)) return fail(h, SPRB200_ERR_BAD_ARG, "L must be positive and finite");
    if (sim->T < 1 || !sim->tsave) return fail(h, SPRB200_ERR_BAD_ARG, "tsave must hold at least one time");
    for (int s = 0; s < sim->T; ++s) {
        if (std::isnan(sim->tsave[s])) return fail(h, SPRB200_ERR_BAD_ARG, "tsave contains NaN");
        if (s && sim->tsave[s] < sim->tsave[s - 1]) return fail(h, SPRB200_ERR_BAD_ARG, "tsave must be ascending");
    }
    if (std::isnan(sim->tstop) || std::isnan(sim->tstop_AtoB)) return fail(h, SPRB200_ERR_BAD_ARG, "tstop/tstop_AtoB is NaN");
    // an infinite tstop with positive rates never ends (the reference's tsave = 0:dt:tstop needs a finite one too)
    if (!std::isfinite(sim->tstop)) return fail(h, SPRB200_ERR_BAD_ARG, "tstop must be finite");
    if (!sim->resample_initlocs && !sim->initlocs)
        return fail(h, SPRB200_ERR_BAD_ARG, "initlocs required when resample_initlocs == 0");
    return SPRB200_OK;
}

int check_run(sprb200_ctx *h, const SprRun *run) {
    if (!run) return fail(h, SPRB200_ERR_BAD_ARG, "run is NULL");
    if (run->term_kind == SPRB200_TERM_FIXED) {
        if (run->nsims < 1 || run->nsims > 1000000) return fail(h, SPRB200_ERR_BAD_ARG, "nsims must be in 1..1000000");
    } else if (run->term_kind == SPRB200_TERM_VARIANCE) {
        if (run->minsims < 2 || run->maxsims < run->minsims || run->maxsims > 1000000)
            return fail(h, SPRB200_ERR_BAD_ARG, "need 2 <= minsims <= maxsims <= 1000000");
        if (!(run->ssetol >= 0.0)) return fail(h, SPRB200_ERR_BAD_ARG, "ssetol must be >= 0");
    } else
        return fail(h, SPRB200_ERR_BAD_ARG, "unknown term_kind");
    if (run->replicate_base < 0 || (long long)run->replicate_base + 1000000 > 0x7fffffffLL)
        return fail(h, SPRB200_ERR_BAD_ARG, "replicate_base out of range");
    if (run->observable != SPRB200_OBS_TOTAL_BOUND && run->observable != SPRB200_OBS_TOTAL_A)
        return fail(h, SPRB200_ERR_BAD_ARG, "unknown observable");
    return SPRB200_OK;
}

int check_points(sprb200_ctx *h, const PointDev *pts, int64_t npts) {
    for (int64_t k = 0; k < npts; ++k) {
        const PointDev &p = pts[k];
        if (!(p.kon_eff >= 0.0) || !(p.koff >= 0.0) || !(p.konb >= 0.0) || !(p.reach >= 0.0) ||
            !std::isfinite(p.kon_eff) || !std::isfinite(p.koff) || !std::isfinite(p.konb) || !std::isfinite(p.reach))
            return fail(h, SPRB200_ERR_BAD_ARG, "rates and reach must be finite and >= 0 (point " + std::to_string(k) + ")");
    }
    return SPRB200_OK;
}

// A-priori size of one neighbour-table record at this reach (mean + 3 sigma of the directed entry
// count): only used to cut the ticket range into chunks that fit the arena.  K2 allocates the exact
// size; a trajectory that does not fit is deferred to the next chunk, never dropped.
size_t record_estimate(int N, int DIM, double L, double reach, double scale) {
    const double PI = 3.14159265358979323846;
    double frac = DIM == 3 ? (4.0 / 3.0) * PI * reach * reach * reach / (L * L * L) : PI * reach * reach / (L * L);
    if (frac > 1.0) frac = 1.0;
    const double E = (double)N * (double)(N - 1) * frac;
    double ent = (E + 3.0 * std::sqrt(2.0 * E) + 32.0) * scale;
    const double maxe = (double)N * (double)(N - 1);
    if (ent > maxe) ent = maxe;
    // packed rows: entries_per_word(N) entries per word, every row rounded up to a whole word
    const double epw = (double)entries_per_word(N);
    const double words = ent / epw + (double)N * (epw - 1.0) / epw * (E > 0.0 ? std::min(1.0, E / (double)N) : 0.0);
    return record_bytes(N, (unsigned long long)std::ceil(words));
}

// rough SSA event count of one trajectory: the ranking that starts expensive points first inside one GPU's share
// (10^4 evaluations per C2 build sit on the timed path; the deal between GPUs uses events_estimate below)
double cost_estimate(const PointDev &p, int N, int DIM, double L, double tstop, double toff) {
    const double PI = 3.14159265358979323846;
    const double ta = std::min(std::max(toff, 0.0), tstop), td = tstop - ta;
    const double kon = p.kon_eff, koff = p.koff, konb = p.konb;
    double frac = DIM == 3 ? (4.0 / 3.0) * PI * p.reach * p.reach * p.reach / (L * L * L) : PI * p.reach * p.reach / (L * L);
    const double deg = std::min((double)(N - 1), (double)(N - 1) * std::min(frac, 1.0));
    const double bound = kon * ta / (1.0 + kon * ta);                 // fraction that ever binds
    const double cyc = (kon + koff) > 0 ? kon * koff / (kon + koff) : 0.0;
    double ev = N * (std::min(1.0, kon * ta) + 2.0 * cyc * ta + bound * std::min(1.0, koff * td));
    const double pre = (konb + koff) > 0 ? konb / (konb + koff) : 0.0; // re-formation probability
    ev += N * std::min(1.0, deg) * bound * 2.0 * koff * tstop * pre * (1.0 + pre * 4.0);
    return (ev + 600.0) * (1.0 + deg / 32.0);
}

// ---- expected SSA events of one trajectory from the mean-field kinetics of the four channels ----------------
// (what the multi-GPU deal balances; the in-rank ordering above only needs a ranking and keeps the cheap formula.)
// Fractions a, b, c of the particles in states A, B and in a complex obey
//     da/dt = -kon a + koff (1 - a) - p a b,     db/dt = kon a + koff (1 - a) - 2 koff b - p a b,     c = 1 - a - b
// with p = konb x (neighbours within reach); the event rate per particle is kon a + koff (1 - a) + p a b.  Particles
// without any neighbour (a fraction exp(-degree) of a uniform configuration) only flip A <-> B (closed form); the others
// see the conditional mean degree.  On 80,000 points of the C2 ranges this is within 2 % of the events the kernels
// execute in every degree class, 1 % - 99 % quantiles of true/estimate 0.87 - 1.14 (tests/tools/cost_model_check.py:
// oracle/spr_mirror.c counts the events bit for bit), and the slowest of 8 ranks gets +0.1 % of the mean instead
// of +1.6 % with the cheap formula.
struct OdeGrid {
    // geometric steps h0 q^k, k < NSTEP, that sum to the phase length
    static constexpr int NSTEP = 16;
    double h0 = 0.0, q = 1.0;
    explicit OdeGrid(double dur) {
        if (!(dur > 0.0)) return;
        h0 = std::min(1e-3, dur / NSTEP);
        if (h0 * NSTEP >= dur) { h0 = dur / NSTEP; return; }
        double lo = 1.0, hi = 16.0;
        for (int it = 0; it < 60; ++it) {
            const double m = 0.5 * (lo + hi);
            const double sum = h0 * (std::pow(m, NSTEP) - 1.0) / (m - 1.0);
            if (sum > dur) hi = m; else lo = m;
        }
        q = 0.5 * (lo + hi);
    }
};

struct PairOde {
    double k1, k2, p;
    double rate(double a, double b) const { return k1 * a + k2 * (1.0 - a) + p * a * b; }
    // one backward-Euler step (two Newton iterations from the old state; the Jacobian's determinant is >= 1)
    void step(double a, double b, double h, double *ao, double *bo) const {
        double an = a, bn = b;
        for (int it = 0; it < 2; ++it) {
            const double pab = p * an * bn;
            const double F1 = an - a - h * (-k1 * an + k2 * (1.0 - an) - pab);
            const double F2 = bn - b - h * (k1 * an + k2 * (1.0 - an) - 2.0 * k2 * bn - pab);
            const double J11 = 1.0 + h * (k1 + k2 + p * bn), J12 = h * p * an;
            const double J21 = -h * (k1 - k2 - p * bn), J22 = 1.0 + h * (2.0 * k2 + p * an);
            const double inv = 1.0 / (J11 * J22 - J12 * J21);
            an = std::min(1.0, std::max(0.0, an - (F1 * J22 - F2 * J12) * inv));
            bn = std::min(1.0, std::max(0.0, bn - (J11 * F2 - J21 * F1) * inv));
        }
        *ao = an; *bo = bn;
    }
    // events per particle over one phase; Richardson extrapolation of the step (second order, still damping the
    // fast modes), Simpson's rule for the rate
    double phase(const OdeGrid &g, double *a, double *b) const {
        double E = 0.0, h = g.h0;
        if (!(h > 0.0)) return 0.0;
        for (int s = 0; s < OdeGrid::NSTEP; ++s) {
            double a1, b1, am, bm, a2, b2;
            step(*a, *b, h, &a1, &b1);
            step(*a, *b, 0.5 * h, &am, &bm);
            step(am, bm, 0.5 * h, &a2, &b2);
            const double ae = std::min(1.0, std::max(0.0, 2.0 * a2 - a1)), be = std::min(1.0, std::max(0.0, 2.0 * b2 - b1));
            E += h * (rate(*a, *b) + 4.0 * rate(am, bm) + rate(ae, be)) / 6.0;
            *a = ae; *b = be;
            h *= g.q;
        }
        return E;
    }
};

double events_estimate(const PointDev &p, int N, int DIM, double L, double tstop, double toff, const OdeGrid &gon,
                       const OdeGrid &goff) {
    const double PI = 3.14159265358979323846;
    const double ta = std::min(std::max(toff, 0.0), tstop), td = tstop - ta;
    const double k1 = p.kon_eff, k2 = p.koff;
    double frac = DIM == 3 ? (4.0 / 3.0) * PI * p.reach * p.reach * p.reach / (L * L * L) : PI * p.reach * p.reach / (L * L);
    const double deg = std::min((double)(N - 1), (double)(N - 1) * std::min(frac, 1.0));
    const double phi = -std::expm1(-deg);                    // particles with at least one neighbour
    // isolated particles: a' = -k1 a + k2 (1 - a), rate k2 + (k1 - k2) a
    double E0 = 0.0, a_off = 1.0;
    const double lam = k1 + k2;
    if (lam > 0.0 && ta > 0.0) {
        const double astar = k2 / lam, rel = -std::expm1(-lam * ta) / lam;
        E0 = k2 * ta + (k1 - k2) * (astar * ta + (1.0 - astar) * rel);
        a_off = astar + (1.0 - astar) * std::exp(-lam * ta);
    }
    E0 += (1.0 - a_off) * -std::expm1(-k2 * td);
    double E1 = E0;
    if (phi > 0.0 && p.konb > 0.0) {
        PairOde ode{k1, k2, p.konb * deg / phi};
        double a = 1.0, b = 0.0;
        E1 = ode.phase(gon, &a, &b);
        ode.k1 = 0.0;
        E1 += ode.phase(goff, &a, &b);
    }
    const double ev = (double)N * ((1.0 - phi) * E0 + phi * E1);
    return std::isfinite(ev) && ev >= 0.0 ? ev : -1.0;
}

// cost of one replicate as the deal sees it: events (+ the save slots) x neighbour work per event
double deal_cost(const PointDev &p, int N, int DIM, double L, double tstop, double toff, const OdeGrid &gon, const OdeGrid &goff) {
    const double ev = events_estimate(p, N, DIM, L, tstop, toff, gon, goff);
    if (ev < 0.0) return cost_estimate(p, N, DIM, L, tstop, toff);
    const double PI = 3.14159265358979323846;
    double frac = DIM == 3 ? (4.0 / 3.0) * PI * p.reach * p.reach * p.reach / (L * L * L) : PI * p.reach * p.reach / (L * L);
    const double deg = std::min((double)(N - 1), (double)(N - 1) * std::min(frac, 1.0));
    return (ev + 600.0) * (1.0 + deg / 32.0);
}

// deal_cost of many points on the host threads (a paper-scale grid has 1.5 million of them)
template <typename GetPoint>
void deal_costs(int64_t npts, const SprSim *sim, GetPoint get, double *out) {
    const double ta = std::min(std::max(sim->tstop_AtoB, 0.0), sim->tstop);
    const OdeGrid gon(ta), goff(sim->tstop - ta);
    auto work = [&](int64_t b, int64_t e) {
        for (int64_t k = b; k < e; ++k) out[k] = deal_cost(get(k), sim->N, sim->DIM, sim->L, sim->tstop, sim->tstop_AtoB, gon, goff);
    };
    int nth = (int)std::min<int64_t>(std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency())), npts / 2048 + 1);
    if (nth <= 1) { work(0, npts); return; }
    std::vector<std::thread> th;
    const int64_t per = (npts + nth - 1) / nth;
    for (int t = 0; t < nth; ++t) th.emplace_back(work, std::min<int64_t>(npts, t * per), std::min<int64_t>(npts, (t + 1) * per));
    for (auto &t : th) t.join();
}

struct RunPlan {
    // inputs
    const SprSim *sim;
    const SprRun *run;
    int observable;        // OBS_*
    bool raw;              // keep curves (simulate_raw)
};

// one fixed-position neighbour table (K2b) in record format; the caller owns `rec`
struct FixedGraph {
    unsigned char *rec = nullptr;
    long long words = 0;
    unsigned int maxdeg = 0;
};

// needs sim->initlocs already uploaded to h->locs
int build_fixed_graph(sprb200_ctx *h, const SprSim *sim, double reach, FixedGraph *g) {
    const int N = sim->N, DIM = sim->DIM;
    cudaStream_t s0 = h->stream[0];
    CUDA_TRY(h, h->goff.ensure(sizeof(uint32_t) * (N + 1)));
    CUDA_TRY(h, h->misc.ensure(64));
    CUDA_TRY(h, cudaMemsetAsync(h->misc.p, 0, 64, s0));
    CUDA_TRY(h, launch_fixed_graph(h->locs.as<double>(), N, DIM, sim->L, reach, h->goff.as<uint32_t>(), nullptr,
                                   h->misc.as<unsigned long long>(),
                                   reinterpret_cast<unsigned int *>(h->misc.as<unsigned long long>() + 1), s0));
    unsigned long long res[2] = {0, 0};
    CUDA_TRY(h, cudaMemcpyAsync(res, h->misc.p, sizeof(res), cudaMemcpyDeviceToHost, s0));
    CUDA_TRY(h, cudaStreamSynchronize(s0));
    h->st_launches += 2;
    g->words = (long long)res[0];
    g->maxdeg = (unsigned int)(res[1] & 0xffffffffu);
    const size_t bytes = record_bytes(N, (unsigned long long)g->words);
    void *p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        return fail(h, SPRB200_ERR_ALLOC, "cudaMalloc(fixed graph record)");
    }
    g->rec = static_cast<unsigned char *>(p);
    uint32_t *words = reinterpret_cast<uint32_t *>(g->rec + align16z(sizeof(uint32_t) * (size_t)(N + 1)));
    cudaError_t e = cudaMemcpyAsync(g->rec, h->goff.p, sizeof(uint32_t) * (N + 1), cudaMemcpyDeviceToDevice, s0);
    if (e == cudaSuccess && g->words > 0)
        e = launch_fixed_graph(h->locs.as<double>(), N, DIM, sim->L, reach, h->goff.as<uint32_t>(), words, nullptr, nullptr, s0);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s0);
    if (e != cudaSuccess) {
        cudaGetLastError();
        cudaFree(g->rec);
        g->rec = nullptr;
        return fail(h, SPRB200_ERR_CUDA, std::string("fixed graph fill: ") + cudaGetErrorString(e));
    }
    h->st_launches += 1;
    return SPRB200_OK;
}

// CSR form (offsets in entries, neighbour indices) of a packed record copied to the host: test hook only
void unpack_record(int N, const uint32_t *offw, const uint32_t *words, std::vector<uint32_t> *off, std::vector<uint32_t> *nbr) {
    const int eb = entry_bits(N), epw = entries_per_word(N);
    const uint32_t empty = (1u << eb) - 1u;
    off->assign((size_t)N + 1, 0u);
    nbr->clear();
    for (int i = 0; i < N; ++i) {
        for (uint32_t w = offw[i]; w < offw[i + 1]; ++w)
            for (int e = 0; e < epw; ++e) {
                const uint32_t j = (words[w] >> (e * eb)) & empty;
                if (j != empty) nbr->push_back(j);
            }
        (*off)[(size_t)i + 1] = (uint32_t)nbr->size();
    }
}

// Simulates `npts` points (already on the host as PointDev + pidx).  On return the device arenas hold
// sum/sumsq [npts][T], nused [npts], nevents [npts] (and curve/curve2 when plan.raw for a single
// batch).  `consume` is called once per batch with (first point, count) after the batch's moments are
// final, while they are still resident at the start of the arenas.
template <typename Consume>
int run_points(sprb200_ctx *h, const RunPlan &plan, const PointDev *pts, const uint32_t *pidx, int64_t npts,
               Consume &&consume) {
    const SprSim *sim = plan.sim;
    const SprRun *run = plan.run;
    const int N = sim->N, DIM = sim->DIM, T = sim->T;
    const bool variance = run->term_kind == SPRB200_TERM_VARIANCE;
    const bool fixedpos = !sim->resample_initlocs;
    const int maxrep = variance ? run->maxsims : run->nsims;
    cudaStream_t s0 = h->stream[0];

    h->st_traj = h->st_events = h->st_launches = 0;
    h->st_ms = h->st_table_ms = h->st_traj_ms = 0.0;
    h->evused = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_k2, ev_k1;
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, h->tsave.ensure(sizeof(double) * T));
    CUDA_TRY(h, cudaMemcpyAsync(h->tsave.p, sim->tsave, sizeof(double) * T, cudaMemcpyHostToDevice, s0));
    CUDA_TRY(h, h->counters.ensure(sizeof(unsigned long long) * CTR_COUNT));

    // ---- K2 (resampled positions) must fit one CTA
    int k2_ctas = 0;
    if (!fixedpos) {
        if (table_smem_bytes(N) > h->smem_optin)
            return fail(h, SPRB200_ERR_TOO_LARGE, "neighbour-table build for N=" + std::to_string(N) + " needs " +
                                                      std::to_string(table_smem_bytes(N)) +
                                                      " B of shared memory, limit " + std::to_string(h->smem_optin));
        k2_ctas = table_ctas_per_sm(N, DIM);
        if (k2_ctas < 1) return fail(h, SPRB200_ERR_TOO_LARGE, "neighbour-table kernel does not fit on an SM");
    }
    // staging rows of K2's distance pass: one slice per resident CTA, stride set by the widest reach of the call
    int stage_cap = 8;
    if (!fixedpos) {
        double maxreach = 0.0;
        for (int64_t k = 0; k < npts; ++k) maxreach = std::max(maxreach, pts[k].reach);
        stage_cap = table_stage_cap(N, DIM, sim->L, maxreach);
        if (h->stage_cap_override > 0) stage_cap = std::min(stage_cap, std::max(1, h->stage_cap_override));
        // at most 2 GB of staging: a denser table than the stride allows takes K2's second-pass path
        const size_t rows_total = (size_t)k2_ctas * h->num_sms * (size_t)N;
        stage_cap = (int)std::max<size_t>(8, std::min<size_t>((size_t)stage_cap, ((size_t)2 << 30) / (rows_total * sizeof(uint16_t))));
        CUDA_TRY(h, h->stage.ensure(sizeof(uint16_t) * (size_t)k2_ctas * h->num_sms * (size_t)N * (size_t)stage_cap));
    }
    {   // the narrowest trajectory layout must fit one warp's slice
        TrajLayout probe;
        if (!make_traj_layout(N, false, false, h->smem_optin, h->smem_per_sm, &probe))
            return fail(h, SPRB200_ERR_TOO_LARGE, "per-trajectory state for N=" + std::to_string(N) + " needs " +
                                                      std::to_string(probe.slice) + " B of shared memory, limit " +
                                                      std::to_string(h->smem_optin));
    }

    // ---- fixed positions: one shared neighbour table per distinct reach (K2b)
    std::map<double, FixedGraph> graphs;
    auto free_graphs = [&]() {
        for (auto &kv : graphs)
            if (kv.second.rec) cudaFree(kv.second.rec);
        graphs.clear();
    };
    if (fixedpos) {
        CUDA_TRY(h, h->locs.ensure(sizeof(double) * N * DIM));
        CUDA_TRY(h, cudaMemcpyAsync(h->locs.p, sim->initlocs, sizeof(double) * N * DIM, cudaMemcpyHostToDevice, s0));
        for (int64_t k = 0; k < npts; ++k) {
            const double reach = pts[k].reach;
            if (graphs.count(reach)) continue;
            FixedGraph g;
            const int rc = build_fixed_graph(h, sim, reach, &g);
            if (rc) { free_graphs(); return rc; }
            graphs[reach] = g;
        }
    }

    // ---- batching so that the replicate curves fit the budget
    const size_t planes = plan.observable == OBS_RAW_BC ? 2 : 1;
    const size_t bytes_per_point = (size_t)maxrep * T * sizeof(uint16_t) * planes;
    int64_t PB = (int64_t)std::max<size_t>(1, h->curve_budget / std::max<size_t>(bytes_per_point, 1));
    if (plan.raw) PB = npts;   // raw output needs every curve resident at once
    PB = std::min<int64_t>(PB, npts);
    PB = std::min<int64_t>(PB, (int64_t)0x7fffffff / std::max(maxrep, 1));
    if (PB < 1) PB = 1;

    CUDA_TRY(h, h->pts.ensure(sizeof(PointDev) * PB));
    CUDA_TRY(h, h->pidx.ensure(sizeof(uint32_t) * PB));
    CUDA_TRY(h, h->order.ensure(sizeof(int32_t) * PB));
    CUDA_TRY(h, h->curve.ensure((size_t)PB * maxrep * T * sizeof(uint16_t)));
    if (planes == 2) CUDA_TRY(h, h->curve2.ensure((size_t)PB * maxrep * T * sizeof(uint16_t)));
    CUDA_TRY(h, h->sum.ensure(sizeof(long long) * PB * T));
    CUDA_TRY(h, h->sumsq.ensure(sizeof(unsigned long long) * PB * T));
    CUDA_TRY(h, h->nused.ensure(sizeof(int32_t) * PB));
    CUDA_TRY(h, h->nevents.ensure(sizeof(unsigned long long) * PB));
    CUDA_TRY(h, h->nscanned.ensure(sizeof(int32_t) * PB));
    CUDA_TRY(h, h->nstar.ensure(sizeof(int32_t) * PB));
    CUDA_TRY(h, h->active.ensure(sizeof(int32_t) * PB));
    if (fixedpos) CUDA_TRY(h, h->slotrec.ensure(sizeof(unsigned long long) * PB));

    CUDA_TRY(h, cudaEventRecord(h->ev_begin, s0));
    std::vector<int32_t> h_nstar, h_active, h_order;
    std::vector<unsigned long long> h_nev, h_slotrec;
    std::vector<size_t> h_est;
    std::vector<long long> deferred, h_defer;
    const size_t budget = std::max(h->table_budget, (size_t)1 << 16);

    for (int64_t b0 = 0; b0 < npts; b0 += PB) {
        const int64_t nb = std::min<int64_t>(PB, npts - b0);
        CUDA_TRY(h, cudaMemcpyAsync(h->pts.p, pts + b0, sizeof(PointDev) * nb, cudaMemcpyHostToDevice, s0));
        CUDA_TRY(h, cudaMemcpyAsync(h->pidx.p, pidx + b0, sizeof(uint32_t) * nb, cudaMemcpyHostToDevice, s0));
        CUDA_TRY(h, cudaMemsetAsync(h->nevents.p, 0, sizeof(unsigned long long) * nb, s0));
        if (variance) {
            CUDA_TRY(h, cudaMemsetAsync(h->sum.p, 0, sizeof(long long) * nb * T, s0));
            CUDA_TRY(h, cudaMemsetAsync(h->sumsq.p, 0, sizeof(unsigned long long) * nb * T, s0));
            CUDA_TRY(h, cudaMemsetAsync(h->nscanned.p, 0, sizeof(int32_t) * nb, s0));
            CUDA_TRY(h, cudaMemsetAsync(h->nstar.p, 0, sizeof(int32_t) * nb, s0));
        }
        unsigned int fixed_maxdeg = 0;
        long long fixed_maxent = 0;
        if (fixedpos) {
            h_slotrec.resize(nb);
            for (int64_t k = 0; k < nb; ++k) {
                const FixedGraph &g = graphs[pts[b0 + k].reach];
                h_slotrec[k] = (unsigned long long)(uintptr_t)g.rec;
                fixed_maxdeg = std::max(fixed_maxdeg, g.maxdeg);
                fixed_maxent = std::max(fixed_maxent, g.words);
            }
            CUDA_TRY(h, cudaMemcpyAsync(h->slotrec.p, h_slotrec.data(), sizeof(unsigned long long) * nb,
                                        cudaMemcpyHostToDevice, s0));
        }
        // active slots of this batch (all of them at first)
        h_active.resize(nb);
        for (int64_t k = 0; k < nb; ++k) h_active[k] = (int32_t)k;

        const int rbase = run->replicate_base;
        int rep_lo = 0;                                  // relative to replicate_base
        int rep_hi = variance ? run->minsims : run->nsims;
        while (!h_active.empty()) {
            // ---- work order of this stage: every active slot, expensive points first
            const int nrep = rep_hi - rep_lo;
            {
                std::vector<std::pair<double, int32_t>> keyed;
                keyed.reserve(h_active.size());
                for (int32_t sl : h_active)
                    keyed.emplace_back(-cost_estimate(pts[b0 + sl], N, DIM, sim->L, sim->tstop, sim->tstop_AtoB), sl);
                std::sort(keyed.begin(), keyed.end());
                h_order.resize(keyed.size());
                h_est.resize(keyed.size());
                for (size_t k = 0; k < keyed.size(); ++k) {
                    h_order[k] = keyed[k].second;
                    h_est[k] = fixedpos ? 0 : record_estimate(N, DIM, sim->L, pts[b0 + keyed[k].second].reach, h->est_scale);
                }
            }
            CUDA_TRY(h, cudaMemcpyAsync(h->order.p, h_order.data(), sizeof(int32_t) * h_order.size(),
                                        cudaMemcpyHostToDevice, s0));
            const long long ntick_total = (long long)h_order.size() * nrep;
            h->st_traj += (uint64_t)ntick_total;

            if (!fixedpos) {
                // one arena allocation per stage (growing it later would synchronise the device)
                size_t total_est = 0;
                for (size_t k = 0; k < h_est.size() && total_est < budget; ++k) total_est += h_est[k] * (size_t)nrep;
                total_est += total_est / 64 + ((size_t)1 << 20);
                CUDA_TRY(h, h->arena.ensure(std::min(total_est, budget)));
            }
            // ---- chunks of tickets whose tables fit the arena: K2 builds the tables, K1 runs the trajectories
            long long tk = 0;
            deferred.clear();
            while (tk < ntick_total || !deferred.empty()) {
                long long n = 0;
                const long long tk0 = tk;
                const bool explicit_list = !deferred.empty();
                size_t want = 0;                 // arena bytes this chunk is expected to need
                if (explicit_list) {
                    n = (long long)deferred.size();
                    CUDA_TRY(h, h->tickets.ensure(sizeof(long long) * deferred.size()));
                    CUDA_TRY(h, cudaMemcpyAsync(h->tickets.p, deferred.data(), sizeof(long long) * deferred.size(),
                                                cudaMemcpyHostToDevice, s0));
                    for (long long t : deferred) want += 2 * h_est[(size_t)(t / nrep)] + 4096;
                    want = std::max(want, record_bytes(N, (unsigned long long)N * (unsigned long long)((N - 1 + entries_per_word(N) - 1) / entries_per_word(N))));
                } else if (fixedpos) {
                    n = ntick_total;
                    tk = ntick_total;
                } else {
                    long long tk1 = tk;
                    while (tk1 < ntick_total) {
                        const size_t eb = std::max<size_t>(h_est[(size_t)(tk1 / nrep)], 16);
                        const long long room = (long long)((budget > want ? budget - want : 0) / eb);
                        const long long take = std::min<long long>(nrep - tk1 % nrep, room);
                        if (take <= 0) break;
                        tk1 += take;
                        want += (size_t)take * eb;
                    }
                    if (tk1 == tk) { tk1 = tk + 1; want = budget; }   // a single record beyond the estimate: K2 decides
                    n = tk1 - tk;
                    tk = tk1;
                }
                CUDA_TRY(h, cudaMemsetAsync(h->counters.p, 0, sizeof(unsigned long long) * CTR_COUNT, s0));
                unsigned long long ctr[CTR_COUNT] = {};
                if (!fixedpos) {
                    want = std::min(want, budget);
                    CUDA_TRY(h, h->arena.ensure(want));
                    CUDA_TRY(h, h->desc.ensure(sizeof(unsigned long long) * (size_t)n));
                    CUDA_TRY(h, h->defer.ensure(sizeof(long long) * (size_t)n));
                    TableArgs ta{};
                    ta.N = N; ta.L = sim->L;
                    ta.pts = h->pts.as<PointDev>(); ta.pidx = h->pidx.as<uint32_t>(); ta.order = h->order.as<int32_t>();
                    ta.tk0 = tk0; ta.ntickets = n;
                    ta.ticket_list = explicit_list ? h->tickets.as<long long>() : nullptr;
                    ta.rep0 = rbase + rep_lo; ta.nrep = nrep;
                    ta.seed_lo = (uint32_t)(run->seed & 0xffffffffu); ta.seed_hi = (uint32_t)(run->seed >> 32);
                    ta.ctr = h->counters.as<unsigned long long>();
                    ta.arena = h->arena.as<unsigned char>();
                    ta.arena_bytes = std::min<unsigned long long>(h->arena.cap, budget);
                    ta.desc = h->desc.as<unsigned long long>();
                    ta.defer_list = h->defer.as<long long>();
                    ta.spos_out = nullptr;
                    ta.stage = h->stage.as<uint16_t>();
                    ta.stage_cap = stage_cap;
                    ta.epw = entries_per_word(N); ta.eb = entry_bits(N);
                    const long long nblk = std::min<long long>(n, (long long)k2_ctas * h->num_sms);
                    cudaEvent_t ea = h->next_event(), eb = h->next_event();
                    if (ea && eb) cudaEventRecord(ea, s0);
                    CUDA_TRY(h, launch_table(ta, DIM, (int)nblk, s0));
                    if (ea && eb) { cudaEventRecord(eb, s0); ev_k2.emplace_back(ea, eb); }
                    h->st_launches += 1;
                    CUDA_TRY(h, cudaMemcpyAsync(ctr, h->counters.p, sizeof(ctr), cudaMemcpyDeviceToHost, s0));
                    CUDA_TRY(h, cudaStreamSynchronize(s0));
                } else {
                    ctr[CTR_MAX_DEGREE] = fixed_maxdeg;
                    ctr[CTR_MAX_ENTRIES] = (unsigned long long)fixed_maxent;
                }
                const long long ndefer = (long long)ctr[CTR_NDEFER];
                if (ndefer > 0) {
                    if (explicit_list && ndefer >= n) {
                        free_graphs();
                        return fail(h, SPRB200_ERR_TOO_LARGE, "a neighbour table does not fit the table arena (" +
                                                                  std::to_string(budget) + " B; SPRB200_TABLE_BYTES)");
                    }
                    h_defer.resize((size_t)ndefer);
                    CUDA_TRY(h, cudaMemcpyAsync(h_defer.data(), h->defer.p, sizeof(long long) * (size_t)ndefer,
                                                cudaMemcpyDeviceToHost, s0));
                    CUDA_TRY(h, cudaStreamSynchronize(s0));
                }
                if (ndefer < n) {
                    TrajLayout lay;
                    const bool wide_sn = ctr[CTR_MAX_DEGREE] > 127ULL, wide_off = ctr[CTR_MAX_ENTRIES] > 65535ULL;
                    // the tensor-memory variant has a fixed CTA shape: only for launches that fill the GPU
                    const bool want_tmem = h->use_tmem && h->max_warps == 0 && n >= (long long)32 * h->num_sms;
                    if (!make_traj_layout(N, wide_sn, wide_off, h->smem_optin, h->smem_per_sm, &lay, want_tmem)) {
                        free_graphs();
                        return fail(h, SPRB200_ERR_TOO_LARGE,
                                    "per-trajectory state (N=" + std::to_string(N) + ", max degree " +
                                        std::to_string(ctr[CTR_MAX_DEGREE]) + ", " + std::to_string(ctr[CTR_MAX_ENTRIES]) +
                                        " table words) needs " + std::to_string(lay.slice) +
                                        " B of shared memory, limit " + std::to_string(h->smem_optin));
                    }
                    if (h->max_warps > 0 && lay.cta_per_sm * lay.wpc > h->max_warps) {
                        if (h->max_warps < lay.wpc) lay.wpc = h->max_warps;
                        lay.cta_per_sm = std::max(1, h->max_warps / lay.wpc);
                    }
                    // a launch that cannot fill the GPU (C1 / C4 / post-fit re-simulation: a few hundred to a few
                    // thousand trajectories) runs narrower CTAs, so that the block scheduler spreads the warps
                    // evenly over all SMs and their four schedulers instead of stacking 7 per CTA
                    // ... and, when every trajectory of the launch can be resident WITH its neighbour rows in shared
                    // memory, one-warp CTAs that copy the rows in once: each row load is then an LDS instead of an L2
                    // round trip, which is what a half-empty GPU is waiting for (per-event latency, not issue slots)
                    if (!lay.tmem && !wide_sn && !wide_off && lay.eb == 10 && h->use_smem_table && h->max_warps == 0 &&
                        ctr[CTR_MAX_ENTRIES] > 0) {
                        const size_t tbl = align16z(sizeof(uint32_t) * (size_t)ctr[CTR_MAX_ENTRIES]);
                        const size_t per = (size_t)lay.slice + tbl;
                        if (per <= (size_t)h->smem_optin) {
                            const long long c = std::min<long long>(K1_MAX_WARPS, (long long)(h->smem_per_sm / (per + 1024)));
                            if (c >= 1 && n <= c * h->num_sms) {
                                lay.o_tbl = lay.slice;
                                lay.tbl_words = (int)ctr[CTR_MAX_ENTRIES];
                                lay.slice = (int)per;
                                lay.wpc = 1;
                                lay.cta_per_sm = (int)c;
                            }
                        }
                    }
                    if (!lay.tmem && lay.tbl_words == 0) {
                        const long long full = (long long)lay.cta_per_sm * lay.wpc * h->num_sms;
                        if (n < full) {
                            const long long per_sm = (n + h->num_sms - 1) / h->num_sms;      // warps per SM if spread evenly
                            int w = (int)std::max<long long>(1, per_sm / 4);                 // at least 4 CTAs per SM
                            if (w < lay.wpc) {
                                lay.cta_per_sm = std::min(32, (lay.cta_per_sm * lay.wpc) / w);
                                lay.wpc = w;
                            }
                        }
                    }
                    TrajArgs a{};
                    a.N = N; a.T = T; a.tstop = sim->tstop;
                    // tstop_AtoB <= 0: the bath is off from t = 0 on; the clock must not move backwards
                    // (the reference keeps tc = 0 and only zeroes kon, forward_simulator.jl:174-181)
                    a.toff = std::max(sim->tstop_AtoB, 0.0);
                    a.tsave = h->tsave.as<double>();
                    a.pts = h->pts.as<PointDev>(); a.pidx = h->pidx.as<uint32_t>(); a.order = h->order.as<int32_t>();
                    a.tk0 = tk0; a.ntickets = n;
                    a.ticket_list = explicit_list ? h->tickets.as<long long>() : nullptr;
                    a.rep0 = rbase + rep_lo; a.nrep = nrep; a.repstride = maxrep; a.rep_rowbase = rbase;
                    a.seed_lo = (uint32_t)(run->seed & 0xffffffffu); a.seed_hi = (uint32_t)(run->seed >> 32);
                    a.ctr = h->counters.as<unsigned long long>();
                    a.desc = fixedpos ? nullptr : h->desc.as<unsigned long long>();
                    a.slot_rec = fixedpos ? h->slotrec.as<unsigned long long>() : nullptr;
                    a.curve = h->curve.as<uint16_t>();
                    a.curve2 = planes == 2 ? h->curve2.as<uint16_t>() : nullptr;
                    a.nevents = h->nevents.as<unsigned long long>();
                    a.observable = plan.observable;
                    a.lay = lay;
                    const long long nblk = std::min<long long>((n + lay.wpc - 1) / lay.wpc,
                                                               (long long)lay.cta_per_sm * h->num_sms);
                    cudaEvent_t ea = h->next_event(), eb = h->next_event();
                    if (ea && eb) cudaEventRecord(ea, s0);
                    CUDA_TRY(h, launch_traj(a, (int)nblk, s0));
                    if (ea && eb) { cudaEventRecord(eb, s0); ev_k1.emplace_back(ea, eb); }
                    h->st_launches += 1;
                    // the arena, descriptors and ticket list are reused by the next chunk: stream order keeps
                    // this launch ahead of the next K2
                }
                if (ndefer > 0) deferred.assign(h_defer.begin(), h_defer.begin() + ndefer);
                else deferred.clear();
            }
            // ---- K3
            if (!variance) {
                CUDA_TRY(h, launch_reduce_fixed(h->curve.as<uint16_t>(), (int)nb, maxrep, run->nsims, T,
                                                h->sum.as<long long>(), h->sumsq.as<unsigned long long>(), s0));
                h->st_launches += 1;
                h_active.clear();
            } else {
                CUDA_TRY(h, cudaMemcpyAsync(h->active.p, h_active.data(), sizeof(int32_t) * h_active.size(),
                                            cudaMemcpyHostToDevice, s0));
                CUDA_TRY(h, launch_variance_scan(h->curve.as<uint16_t>(), h->active.as<int32_t>(), (int)h_active.size(),
                                                 maxrep, T, rep_hi, run->ssetol, run->minsims, run->maxsims,
                                                 h->sum.as<long long>(), h->sumsq.as<unsigned long long>(),
                                                 h->nscanned.as<int32_t>(), h->nstar.as<int32_t>(), s0));
                h->st_launches += 1;
                h_nstar.resize(nb);
                CUDA_TRY(h, cudaMemcpyAsync(h_nstar.data(), h->nstar.p, sizeof(int32_t) * nb, cudaMemcpyDeviceToHost, s0));
                CUDA_TRY(h, cudaStreamSynchronize(s0));
                std::vector<int32_t> next;
                for (int32_t sl : h_active)
                    if (h_nstar[sl] == 0) next.push_back(sl);
                h_active.swap(next);
                rep_lo = rep_hi;
                // replicate stages grow geometrically: the stopping index is exact (sequential scan),
                // only replicates past it inside the last stage are surplus work
                rep_hi = std::min(run->maxsims, std::max(rep_hi + 1, rep_hi + rep_hi / 2));
                if (rep_lo >= run->maxsims && !h_active.empty())
                    { free_graphs(); return fail(h, SPRB200_ERR_INTERNAL, "variance scan did not terminate"); }
            }
        }
        if (!variance) {
            // nused = nsims for every slot
            std::vector<int32_t> nu(nb, run->nsims);
            CUDA_TRY(h, cudaMemcpyAsync(h->nused.p, nu.data(), sizeof(int32_t) * nb, cudaMemcpyHostToDevice, s0));
            CUDA_TRY(h, cudaStreamSynchronize(s0));
        } else {
            CUDA_TRY(h, cudaMemcpyAsync(h->nused.p, h->nstar.p, sizeof(int32_t) * nb, cudaMemcpyDeviceToDevice, s0));
        }
        h_nev.resize(nb);
        CUDA_TRY(h, cudaMemcpyAsync(h_nev.data(), h->nevents.p, sizeof(unsigned long long) * nb, cudaMemcpyDeviceToHost, s0));
        CUDA_TRY(h, cudaStreamSynchronize(s0));
        for (int64_t k = 0; k < nb; ++k) h->st_events += h_nev[k];
        int rc = consume(b0, nb, h_nev.data());
        if (rc != SPRB200_OK) { free_graphs(); return rc; }
    }
    CUDA_TRY(h, cudaEventRecord(h->ev_end, s0));
    CUDA_TRY(h, cudaStreamSynchronize(s0));
    float ms = 0.f;
    CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev_begin, h->ev_end));
    h->st_ms = ms;
    for (auto &pr : ev_k2) { float v = 0.f; if (cudaEventElapsedTime(&v, pr.first, pr.second) == cudaSuccess) h->st_table_ms += v; }
    for (auto &pr : ev_k1) { float v = 0.f; if (cudaEventElapsedTime(&v, pr.first, pr.second) == cudaSuccess) h->st_traj_ms += v; }
    cudaGetLastError();
    free_graphs();
    return SPRB200_OK;
}

int grid_total(const SprGrid *g, int64_t *P) {
    if (!g) return SPRB200_ERR_BAD_ARG;
    int64_t p = 1;
    for (int k = 0; k < 4; ++k) {
        if (g->n[k] < 1) return SPRB200_ERR_BAD_ARG;
        p *= g->n[k];
        if (p > (int64_t)0xffffffffLL) return SPRB200_ERR_BAD_ARG;
    }
    *P = p;
    return SPRB200_OK;
}

void grid_point(const SprGrid *g, int64_t lin0, PointDev *out) {
    const int i1 = (int)(lin0 % g->n[0]);
    const int i2 = (int)((lin0 / g->n[0]) % g->n[1]);
    const int i3 = (int)((lin0 / ((int64_t)g->n[0] * g->n[1])) % g->n[2]);
    const int i4 = (int)(lin0 / ((int64_t)g->n[0] * g->n[1] * g->n[2]));
    out->kon_eff = std::pow(10.0, sprb200_grid_value(g->lo[0], g->hi[0], g->n[0], i1));
    out->koff = std::pow(10.0, sprb200_grid_value(g->lo[1], g->hi[1], g->n[1], i2));
    out->konb = std::pow(10.0, sprb200_grid_value(g->lo[2], g->hi[2], g->n[2], i3));
    out->reach = sprb200_grid_value(g->lo[3], g->hi[3], g->n[3], i4);
}

// shared implementation of the slice builders: rows [count][T] land in d_rows (device);
// idx1[k] are 1-based linear grid indices (kon fastest ... reach slowest, surrogate.jl:284)
int build_rows(sprb200_handle h, const SprGrid *grid, const SprSim *sim, const SprRun *run, const int64_t *idx1,
               int64_t count, double *d_rows, int32_t *d_nused, uint64_t *d_nevents, int32_t *h_nused,
               uint64_t *h_nevents) {
    if (!h) return SPRB200_ERR_BAD_ARG;
    int rc = check_sim(h, sim);
    if (rc) return rc;
    rc = check_run(h, run);
    if (rc) return rc;
    int64_t P = 0;
    if (grid_total(grid, &P)) return fail(h, SPRB200_ERR_BAD_ARG, "bad grid");
    if (count < 1 || !idx1) return fail(h, SPRB200_ERR_BAD_ARG, "empty index set");
    std::vector<PointDev> pts((size_t)count);
    std::vector<uint32_t> pidx((size_t)count);
    for (int64_t k = 0; k < count; ++k) {
        if (idx1[k] < 1 || idx1[k] > P) return fail(h, SPRB200_ERR_BAD_ARG, "index outside the grid");
        const int64_t lin0 = idx1[k] - 1;
        grid_point(grid, lin0, &pts[k]);
        pidx[k] = (uint32_t)lin0;
    }
    rc = check_points(h, pts.data(), count);
    if (rc) return rc;
    SprRun r = *run;
    r.observable = SPRB200_OBS_TOTAL_BOUND;
    r.replicate_base = 0;
    RunPlan plan{sim, &r, OBS_TOTAL_BOUND, false};
    const int T = sim->T, N = sim->N;
    return run_points(h, plan, pts.data(), pidx.data(), count, [&](int64_t b0, int64_t nb, const unsigned long long *nev) -> int {
        cudaStream_t s0 = h->stream[0];
        CUDA_TRY(h, launch_means(h->sum.as<long long>(), h->nused.as<int32_t>(), (int)nb, T, N, d_rows + (size_t)b0 * T, s0));
        h->st_launches += 1;
        if (d_nused) CUDA_TRY(h, cudaMemcpyAsync(d_nused + b0, h->nused.p, sizeof(int32_t) * nb, cudaMemcpyDeviceToDevice, s0));
        if (d_nevents) CUDA_TRY(h, cudaMemcpyAsync(d_nevents + b0, h->nevents.p, sizeof(uint64_t) * nb, cudaMemcpyDeviceToDevice, s0));
        if (h_nused) CUDA_TRY(h, cudaMemcpyAsync(h_nused + b0, h->nused.p, sizeof(int32_t) * nb, cudaMemcpyDeviceToHost, s0));
        if (h_nevents) for (int64_t k = 0; k < nb; ++k) h_nevents[b0 + k] = nev[k];
        CUDA_TRY(h, cudaStreamSynchronize(s0));
        return SPRB200_OK;
    });
}

std::vector<int64_t> strided(int64_t idxstart, int64_t stride, int64_t count) {
    std::vector<int64_t> v((size_t)std::max<int64_t>(count, 0));
    for (int64_t k = 0; k < count; ++k) v[k] = idxstart + k * stride;
    return v;
}


// ---- cost-balanced deal of grid points to ranks (multi-GPU builds) ------------------------------
// The points, sorted by the a-priori cost estimate (descending, ties by index), are handed out in snake
// order 0..w-1, w-1..0, ...: every rank simulates the same mix of cheap and expensive points and the
// counts differ by at most one.  Deterministic and device-free: every rank computes the same lists.
// The reference's own split is by contiguous index ranges (examples/cluster_scripts/queue_job.sh:7-17).
// Replicates the VarianceTerminator is expected to use at a point: the stopping rule needs SEM <= ssetol * mean in
// every bin, and the hardest bin is the one with the fewest bound particles c (relative spread of one replicate
// 1/sqrt(c)): n = 1 / (ssetol^2 c), clamped to [minsims, maxsims].  c is taken at the first save time after 0
// (association just started) and at tstop (dissociation, slowed by the fraction of antibodies held by both arms);
// on 48 sampled points of the C2 and C3 grids this reproduces the measured count at 47
// (tests/test_gpu_parity.py prints them).  Only the relative cost of the points matters here.
double replicate_estimate(const PointDev &p, int DIM, const SprSim *sim, const SprRun *run) {
    if (!run || run->term_kind != SPRB200_TERM_VARIANCE) return 1.0;
    const double PI = 3.14159265358979323846;
    const int N = sim->N;
    double t1 = sim->tstop / 600.0;
    if (sim->tsave)
        for (int s = 0; s < sim->T; ++s)
            if (sim->tsave[s] > 0.0) { t1 = sim->tsave[s]; break; }
    const double toff = std::min(std::max(sim->tstop_AtoB, 0.0), sim->tstop);
    const double kon = p.kon_eff, koff = p.koff, konb = p.konb;
    const double c1 = (double)N * -std::expm1(-kon * std::min(t1, toff));
    double frac = DIM == 3 ? (4.0 / 3.0) * PI * p.reach * p.reach * p.reach / (sim->L * sim->L * sim->L)
                           : PI * p.reach * p.reach / (sim->L * sim->L);
    const double dmin = std::min(1.0, (double)(N - 1) * std::min(frac, 1.0));
    const double ptoff = (kon + koff) > 0.0 ? kon / (kon + koff) * -std::expm1(-(kon + koff) * toff) : 0.0;
    const double fC = (konb * dmin + 2.0 * koff) > 0.0 ? konb * dmin / (konb * dmin + 2.0 * koff) : 0.0;
    const double cend = (double)N * ptoff * std::exp(-koff * (1.0 - fC) * (sim->tstop - toff));
    const double c = std::min(c1, cend);
    const double tol2 = std::max(run->ssetol * run->ssetol, 1e-300);
    const double n = c > 0.0 ? 1.0 / (tol2 * c) : (double)run->maxsims;
    return std::min((double)run->maxsims, std::max((double)run->minsims, n));
}

// the last deal is kept: a build asks for it once per call (twice through sprb200_deal_indices), and a caller that
// repeats a build asks for the same one again
constexpr int64_t DEAL_ODE_MAX_POINTS = 400000;

struct DealMemo {
    std::mutex mu;
    std::vector<unsigned char> key;
    std::vector<std::vector<int64_t>> deal;
} g_deal_memo;

std::vector<unsigned char> deal_key(const SprGrid *grid, const SprSim *sim, const SprRun *run, int nranks) {
    std::vector<unsigned char> k;
    auto put = [&](const void *p, size_t n) { k.insert(k.end(), (const unsigned char *)p, (const unsigned char *)p + n); };
    put(grid, sizeof(*grid));
    put(&sim->N, sizeof(sim->N)); put(&sim->DIM, sizeof(sim->DIM)); put(&sim->L, sizeof(sim->L));
    put(&sim->tstop, sizeof(sim->tstop)); put(&sim->tstop_AtoB, sizeof(sim->tstop_AtoB));
    double t1 = -1.0;                                         // what replicate_estimate reads of tsave
    if (sim->tsave)
        for (int s = 0; s < sim->T; ++s)
            if (sim->tsave[s] > 0.0) { t1 = sim->tsave[s]; break; }
    put(&t1, sizeof(t1));
    const int variance = run && run->term_kind == SPRB200_TERM_VARIANCE ? 1 : 0;
    put(&variance, sizeof(variance));
    if (variance) { put(&run->ssetol, sizeof(run->ssetol)); put(&run->minsims, sizeof(run->minsims)); put(&run->maxsims, sizeof(run->maxsims)); }
    put(&nranks, sizeof(nranks));
    return k;
}

int deal_points(const SprGrid *grid, const SprSim *sim, const SprRun *run, int nranks, std::vector<std::vector<int64_t>> *out) {
    int64_t P = 0;
    if (grid_total(grid, &P) || nranks < 1) return SPRB200_ERR_BAD_ARG;
    if (nranks == 1) {                                        // nothing to balance
        out->assign(1, std::vector<int64_t>((size_t)P));
        for (int64_t i = 0; i < P; ++i) (*out)[0][(size_t)i] = i + 1;
        return SPRB200_OK;
    }
    const std::vector<unsigned char> key = deal_key(grid, sim, run, nranks);
    std::lock_guard<std::mutex> lk(g_deal_memo.mu);
    if (key == g_deal_memo.key) {
        *out = g_deal_memo.deal;
        return SPRB200_OK;
    }
    // The kinetic estimate costs 3 us per point; the cheap formula's errors average out once a rank gets enough points
    // (paper-scale grid, 189,000 points per rank: slowest rank +0.5 % with it; 10,000 per rank: +1.6 %).
    std::vector<double> cost((size_t)P);
    if (P <= DEAL_ODE_MAX_POINTS) {
        deal_costs(P, sim, [&](int64_t i) { PointDev p; grid_point(grid, i, &p); return p; }, cost.data());
    } else {
        for (int64_t i = 0; i < P; ++i) {
            PointDev p;
            grid_point(grid, i, &p);
            cost[(size_t)i] = cost_estimate(p, sim->N, sim->DIM, sim->L, sim->tstop, sim->tstop_AtoB);
        }
    }
    std::vector<std::pair<double, int64_t>> keyed((size_t)P);
    for (int64_t i = 0; i < P; ++i) {
        PointDev p;
        grid_point(grid, i, &p);
        keyed[(size_t)i] = std::make_pair(-cost[(size_t)i] * replicate_estimate(p, sim->DIM, sim, run), i);
    }
    std::sort(keyed.begin(), keyed.end());
    out->assign((size_t)nranks, std::vector<int64_t>());
    for (auto &v : *out) v.reserve((size_t)(P / nranks + 1));
    for (int64_t k = 0; k < P; ++k) {
        const int64_t lap = k / nranks, within = k % nranks;
        const int r = (int)((lap & 1) ? nranks - 1 - within : within);
        (*out)[(size_t)r].push_back(keyed[(size_t)k].second + 1);   // 1-based linear index
    }
    for (auto &v : *out) std::sort(v.begin(), v.end());
    g_deal_memo.key = key;
    g_deal_memo.deal = *out;
    return SPRB200_OK;
}

// ---- NCCL, bound at run time (dlopen): the library has no link-time dependency on it, a single-GPU
// caller never needs it, and inside a process that already loaded NCCL (PyTorch) the same copy is used.
struct NcclId { char internal[128]; };
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, NcclId /* ncclUniqueId, by value */, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mutex;

bool nccl_load(std::string *err) {
    std::lock_guard<std::mutex> lk(g_nccl_mutex);
    if (g_nccl.lib) return true;
    const char *names[] = {std::getenv("SPRB200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void *lib = nullptr;
    for (const char *nm : names) {
        if (!nm || !*nm) continue;
        lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) { *err = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "not found"); return false; }
    g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(dlsym(lib, "ncclGetUniqueId"));
    g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(dlsym(lib, "ncclCommInitRank"));
    g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
    g_nccl.AllGather = reinterpret_cast<decltype(g_nccl.AllGather)>(dlsym(lib, "ncclAllGather"));
    g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllGather) {
        *err = "libnccl.so.2 lacks ncclGetUniqueId/ncclCommInitRank/ncclCommDestroy/ncclAllGather";
        dlclose(lib);
        return false;
    }
    g_nccl.lib = lib;
    return true;
}
#define NCCL_TRY(h, expr)                                                                         \
    do {                                                                                          \
        int r__ = (expr);                                                                         \
        if (r__ != 0)                                                                             \
            return fail(h, SPRB200_ERR_NCCL, std::string(#expr) + ": " +                          \
                                                 (g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "NCCL error")); \
    } while (0)

// handles owned by sprb200_build_table_multi (one per device, kept for the life of the process)
std::map<int, sprb200_ctx *> g_multi_handles;
std::mutex g_multi_mutex;
}  // namespace

// ================================================================================================
extern "C" {

int sprb200_abi_version(void) { return SPRB200_ABI_VERSION; }

int sprb200_create(int device_ordinal, sprb200_handle *out) {
    if (!out) return fail(nullptr, SPRB200_ERR_BAD_ARG, "out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev < 1) {
        cudaGetLastError();
        return fail(nullptr, SPRB200_ERR_NO_DEVICE,
                    std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                        " (libsprb200 has no CPU fallback)");
    }
    if (device_ordinal < 0 || device_ordinal >= ndev) return fail(nullptr, SPRB200_ERR_NO_DEVICE, "bad device ordinal");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device_ordinal) != cudaSuccess) return fail(nullptr, SPRB200_ERR_CUDA, "cudaGetDeviceProperties");
    if (prop.major != 10)
        return fail(nullptr, SPRB200_ERR_NO_DEVICE, std::string("device is ") + prop.name + " (sm_" + std::to_string(prop.major) +
                                                        std::to_string(prop.minor) + "); this library is built for sm_100a only");
    sprb200_ctx *h = new (std::nothrow) sprb200_ctx();
    if (!h) return fail(nullptr, SPRB200_ERR_ALLOC, "out of host memory");
    h->device = device_ordinal;
    h->num_sms = prop.multiProcessorCount;
    if (cudaSetDevice(device_ordinal) != cudaSuccess) { delete h; return fail(nullptr, SPRB200_ERR_CUDA, "cudaSetDevice"); }
    h->smem_optin = (int)prop.sharedMemPerBlockOptin;
    h->smem_per_sm = (int)prop.sharedMemPerMultiprocessor;
    for (int s = 0; s < NSTREAM; ++s) {
        if (cudaStreamCreateWithFlags(&h->stream[s], cudaStreamNonBlocking) != cudaSuccess) { delete h; return fail(nullptr, SPRB200_ERR_CUDA, "cudaStreamCreate"); }
    }
    cudaEventCreate(&h->ev_begin);
    cudaEventCreate(&h->ev_end);
    if (const char *env = std::getenv("SPRB200_CURVE_BYTES")) {
        const long long v = std::atoll(env);
        if (v > 0) h->curve_budget = (size_t)v;
    }
    if (const char *env = std::getenv("SPRB200_TABLE_BYTES")) {
        const long long v = std::atoll(env);
        if (v > 0) h->table_budget = (size_t)v;
    }
    if (const char *env = std::getenv("SPRB200_MAX_WARPS")) h->max_warps = std::atoi(env);
    if (const char *env = std::getenv("SPRB200_STAGE_CAP")) h->stage_cap_override = std::atoi(env);
    if (const char *env = std::getenv("SPRB200_TMEM")) h->use_tmem = std::atoi(env);
    if (const char *env = std::getenv("SPRB200_SMEM_TABLE")) h->use_smem_table = std::atoi(env);
    if (const char *env = std::getenv("SPRB200_CAP_SCALE")) {
        const double v = std::atof(env);
        if (v > 0.0 && v <= 1.0) h->est_scale = v;
    }
    *out = h;
    return SPRB200_OK;
}

void sprb200_destroy(sprb200_handle h) {
    if (!h) return;
    cudaSetDevice(h->device);
    for (int s = 0; s < NSTREAM; ++s) {
        if (h->stream[s]) cudaStreamSynchronize(h->stream[s]);
    }
    if (h->nccl_comm && g_nccl.CommDestroy) {
        g_nccl.CommDestroy(h->nccl_comm);
        h->nccl_comm = nullptr;
    }
    DevBuf *bufs[] = {&h->tsave, &h->pts, &h->pidx, &h->order, &h->curve, &h->curve2, &h->sum, &h->sumsq, &h->nused,
                      &h->nevents, &h->nscanned, &h->nstar, &h->active, &h->counters, &h->rows, &h->locs,
                      &h->goff, &h->gnbr, &h->misc, &h->table, &h->arena, &h->desc, &h->defer, &h->tickets, &h->slotrec, &h->stage, &h->sur, &h->fitdata, &h->fitpars, &h->fitloss,
                      &h->mrows, &h->mtable, &h->midx, &h->mnused, &h->mnev, &h->gnused, &h->gnev};
    for (DevBuf *b : bufs) b->release();
    for (int s = 0; s < NSTREAM; ++s) {
        if (h->stream[s]) cudaStreamDestroy(h->stream[s]);
    }
    for (cudaEvent_t e : h->evpool) cudaEventDestroy(e);
    if (h->ev_begin) cudaEventDestroy(h->ev_begin);
    if (h->ev_end) cudaEventDestroy(h->ev_end);
    delete h;
}

const char *sprb200_last_error(sprb200_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int sprb200_last_stats(sprb200_handle h, uint64_t *trajectories, uint64_t *events, uint64_t *kernel_launches,
                       double *device_ms) {
    if (!h) return SPRB200_ERR_BAD_ARG;
    if (trajectories) *trajectories = h->st_traj;
    if (events) *events = h->st_events;
    if (kernel_launches) *kernel_launches = h->st_launches;
    if (device_ms) *device_ms = h->st_ms;
    return SPRB200_OK;
}

int sprb200_last_kernel_ms(sprb200_handle h, double *table_ms, double *traj_ms) {
    if (!h) return SPRB200_ERR_BAD_ARG;
    if (table_ms) *table_ms = h->st_table_ms;
    if (traj_ms) *traj_ms = h->st_traj_ms;
    return SPRB200_OK;
}

int sprb200_estimate_cost(const SprSim *sim, const SprPoint *pts, int64_t npts, double *cost_out) {
    if (!sim || !pts || !cost_out || npts < 0 || sim->N < 1 || (sim->DIM != 2 && sim->DIM != 3) || !(sim->L > 0.0))
        return SPRB200_ERR_BAD_ARG;
    deal_costs(npts, sim, [&](int64_t k) { return PointDev{pts[k].kon_eff, pts[k].koff, pts[k].konb, pts[k].reach}; }, cost_out);
    return SPRB200_OK;
}

int sprb200_traj_layout(int32_t N, int32_t wide_state, int32_t wide_offsets, int32_t out[4]) {
    if (!out || N < 1 || N > 65535) return SPRB200_ERR_BAD_ARG;
    TrajLayout lay;
    // B200: 227 KB of dynamic shared memory per CTA, 228 KB per SM
    const bool ok = make_traj_layout(N, wide_state != 0, wide_offsets != 0, 227 * 1024, 228 * 1024, &lay);
    out[0] = lay.slice;
    out[1] = lay.wpc;
    out[2] = lay.cta_per_sm;
    out[3] = table_smem_bytes(N);
    return ok ? SPRB200_OK : SPRB200_ERR_TOO_LARGE;
}

double sprb200_domain_length(int32_t N, int32_t DIM, double antigenconcen, int32_t convert_units) {
    const double agc = convert_units ? 6.02214076e-7 * antigenconcen : antigenconcen;
    return std::pow((double)N / agc, 1.0 / (double)DIM);
}

double sprb200_grid_value(double lo, double hi, int32_t n, int32_t i0) {
    if (n <= 1) return lo;
    if (i0 == n - 1) return hi;
    return lo + (double)i0 * ((hi - lo) / (double)(n - 1));
}

int sprb200_grid_point(const SprGrid *grid, int64_t idx, SprPoint *out) {
    int64_t P = 0;
    if (!out || grid_total(grid, &P) || idx < 1 || idx > P) return SPRB200_ERR_BAD_ARG;
    PointDev p;
    grid_point(grid, idx - 1, &p);
    out->kon_eff = p.kon_eff; out->koff = p.koff; out->konb = p.konb; out->reach = p.reach;
    return SPRB200_OK;
}

int sprb200_simulate(sprb200_handle h, const SprSim *sim, const SprPoint *pts, int64_t npts, const SprRun *run,
                     const SprOut *out) {
    if (!h) return SPRB200_ERR_BAD_ARG;
    int rc = check_sim(h, sim);
    if (rc) return rc;
    rc = check_run(h, run);
    if (rc) return rc;
    if (!pts || npts < 1 || !out) return fail(h, SPRB200_ERR_BAD_ARG, "pts/out is NULL or npts < 1");
    if (run->param_index_base + (uint64_t)npts > 0x100000000ULL) return fail(h, SPRB200_ERR_BAD_ARG, "parameter index exceeds 2^32");
    std::vector<PointDev> p((size_t)npts);
    std::vector<uint32_t> pidx((size_t)npts);
    for (int64_t k = 0; k < npts; ++k) {
        p[k] = PointDev{pts[k].kon_eff, pts[k].koff, pts[k].konb, pts[k].reach};
        pidx[k] = (uint32_t)(run->param_index_base + (uint64_t)k);
    }
    rc = check_points(h, p.data(), npts);
    if (rc) return rc;
    RunPlan plan{sim, run, run->observable == SPRB200_OBS_TOTAL_A ? OBS_TOTAL_A : OBS_TOTAL_BOUND, false};
    const int T = sim->T, N = sim->N;
    return run_points(h, plan, p.data(), pidx.data(), npts, [&](int64_t b0, int64_t nb, const unsigned long long *nev) -> int {
        cudaStream_t s0 = h->stream[0];
        if (out->sum) CUDA_TRY(h, cudaMemcpyAsync(out->sum + (size_t)b0 * T, h->sum.p, sizeof(int64_t) * nb * T, cudaMemcpyDeviceToHost, s0));
        if (out->sumsq) CUDA_TRY(h, cudaMemcpyAsync(out->sumsq + (size_t)b0 * T, h->sumsq.p, sizeof(uint64_t) * nb * T, cudaMemcpyDeviceToHost, s0));
        if (out->n_used) CUDA_TRY(h, cudaMemcpyAsync(out->n_used + b0, h->nused.p, sizeof(int32_t) * nb, cudaMemcpyDeviceToHost, s0));
        if (out->n_events) for (int64_t k = 0; k < nb; ++k) out->n_events[b0 + k] = nev[k];
        if (out->mean) {
            CUDA_TRY(h, h->rows.ensure(sizeof(double) * nb * T));
            CUDA_TRY(h, launch_means(h->sum.as<long long>(), h->nused.as<int32_t>(), (int)nb, T, N, h->rows.as<double>(), s0));
            h->st_launches += 1;
            CUDA_TRY(h, cudaMemcpyAsync(out->mean + (size_t)b0 * T, h->rows.p, sizeof(double) * nb * T, cudaMemcpyDeviceToHost, s0));
        }
        CUDA_TRY(h, cudaStreamSynchronize(s0));
        return SPRB200_OK;
    });
}

int sprb200_simulate_raw(sprb200_handle h, const SprSim *sim, const SprPoint *pts, int64_t npts, const SprRun *run,
                         uint16_t *B_out, uint16_t *C_out, uint64_t *n_events) {
    if (!h) return SPRB200_ERR_BAD_ARG;
    int rc = check_sim(h, sim);
    if (rc) return rc;
    rc = check_run(h, run);
    if (rc) return rc;
    if (run->term_kind != SPRB200_TERM_FIXED) return fail(h, SPRB200_ERR_BAD_ARG, "raw output needs term_kind FIXED");
    if (!pts || npts < 1) return fail(h, SPRB200_ERR_BAD_ARG, "pts is NULL or npts < 1");
    if (run->param_index_base + (uint64_t)npts > 0x100000000ULL) return fail(h, SPRB200_ERR_BAD_ARG, "parameter index exceeds 2^32");
    std::vector<PointDev> p((size_t)npts);
    std::vector<uint32_t> pidx((size_t)npts);
    for (int64_t k = 0; k < npts; ++k) {
        p[k] = PointDev{pts[k].kon_eff, pts[k].koff, pts[k].konb, pts[k].reach};
        pidx[k] = (uint32_t)(run->param_index_base + (uint64_t)k);
    }
    rc = check_points(h, p.data(), npts);
    if (rc) return rc;
    RunPlan plan{sim, run, OBS_RAW_BC, true};
    const size_t per = (size_t)run->nsims * sim->T;
    return run_points(h, plan, p.data(), pidx.data(), npts, [&](int64_t b0, int64_t nb, const unsigned long long *nev) -> int {
        cudaStream_t s0 = h->stream[0];
        if (B_out) CUDA_TRY(h, cudaMemcpyAsync(B_out + (size_t)b0 * per, h->curve.p, sizeof(uint16_t) * nb * per, cudaMemcpyDeviceToHost, s0));
        if (C_out) CUDA_TRY(h, cudaMemcpyAsync(C_out + (size_t)b0 * per, h->curve2.p, sizeof(uint16_t) * nb * per, cudaMemcpyDeviceToHost, s0));
        if (n_events) for (int64_t k = 0; k < nb; ++k) n_events[b0 + k] = nev[k];
        CUDA_TRY(h, cudaStreamSynchronize(s0));
        return SPRB200_OK;
    });
}

int64_t sprb200_neighbor_table(sprb200_handle h, const SprSim *sim, double reach, uint64_t seed, uint64_t param_index,
                               uint32_t replicate, double *positions_out, int32_t *off_out, int32_t *nbr_out,
                               int64_t nbr_cap) {
    if (!h) return SPRB200_ERR_BAD_ARG;
    int rc = check_sim(h, sim);
    if (rc) return rc;
    if (!(reach >= 0.0) || !std::isfinite(reach) || !off_out) return fail(h, SPRB200_ERR_BAD_ARG, "bad reach/off_out");
    const int N = sim->N, DIM = sim->DIM;
    cudaStream_t s0 = h->stream[0];
    CUDA_TRY(h, cudaSetDevice(h->device));
    std::vector<uint32_t> offw(N + 1), words, off, nbr;
    const size_t o_words = align16z(sizeof(uint32_t) * (size_t)(N + 1));
    const unsigned long long maxwords =
        (unsigned long long)N * (unsigned long long)((N - 1 + entries_per_word(N) - 1) / entries_per_word(N));
    if (!sim->resample_initlocs) {
        CUDA_TRY(h, h->locs.ensure(sizeof(double) * N * DIM));
        CUDA_TRY(h, cudaMemcpyAsync(h->locs.p, sim->initlocs, sizeof(double) * N * DIM, cudaMemcpyHostToDevice, s0));
        FixedGraph g;
        rc = build_fixed_graph(h, sim, reach, &g);
        if (rc) return rc;
        words.resize((size_t)std::max<long long>(g.words, 1));
        cudaError_t e = cudaMemcpyAsync(offw.data(), g.rec, sizeof(uint32_t) * (N + 1), cudaMemcpyDeviceToHost, s0);
        if (e == cudaSuccess && g.words > 0)
            e = cudaMemcpyAsync(words.data(), g.rec + o_words, sizeof(uint32_t) * g.words, cudaMemcpyDeviceToHost, s0);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s0);
        cudaFree(g.rec);
        if (e != cudaSuccess) { cudaGetLastError(); return fail(h, SPRB200_ERR_CUDA, std::string("fixed graph copy: ") + cudaGetErrorString(e)); }
        if (positions_out) std::memcpy(positions_out, sim->initlocs, sizeof(double) * N * DIM);
    } else {
        if (param_index > 0xffffffffULL) return fail(h, SPRB200_ERR_BAD_ARG, "parameter index exceeds 2^32");
        if (table_smem_bytes(N) > h->smem_optin || table_ctas_per_sm(N, DIM) < 1)
            return fail(h, SPRB200_ERR_TOO_LARGE, "neighbour-table build does not fit in shared memory");
        PointDev p{0, 0, 0, reach};
        const uint32_t pi = (uint32_t)param_index;
        const int32_t ord = 0;
        const size_t recmax = record_bytes(N, maxwords);
        CUDA_TRY(h, h->pts.ensure(sizeof(PointDev)));
        CUDA_TRY(h, h->pidx.ensure(sizeof(uint32_t)));
        CUDA_TRY(h, h->order.ensure(sizeof(int32_t)));
        CUDA_TRY(h, h->counters.ensure(sizeof(unsigned long long) * CTR_COUNT));
        CUDA_TRY(h, h->arena.ensure(recmax));
        CUDA_TRY(h, h->desc.ensure(sizeof(unsigned long long)));
        CUDA_TRY(h, h->defer.ensure(sizeof(long long)));
        CUDA_TRY(h, h->locs.ensure(sizeof(uint64_t) * N + sizeof(double) * N * DIM));
        CUDA_TRY(h, cudaMemcpyAsync(h->pts.p, &p, sizeof(p), cudaMemcpyHostToDevice, s0));
        CUDA_TRY(h, cudaMemcpyAsync(h->pidx.p, &pi, sizeof(pi), cudaMemcpyHostToDevice, s0));
        CUDA_TRY(h, cudaMemcpyAsync(h->order.p, &ord, sizeof(ord), cudaMemcpyHostToDevice, s0));
        CUDA_TRY(h, cudaMemsetAsync(h->counters.p, 0, sizeof(unsigned long long) * CTR_COUNT, s0));
        TableArgs ta{};
        ta.N = N; ta.L = sim->L;
        ta.pts = h->pts.as<PointDev>(); ta.pidx = h->pidx.as<uint32_t>(); ta.order = h->order.as<int32_t>();
        ta.tk0 = 0; ta.ntickets = 1; ta.ticket_list = nullptr;
        ta.rep0 = (int)replicate; ta.nrep = 1;
        ta.seed_lo = (uint32_t)(seed & 0xffffffffu); ta.seed_hi = (uint32_t)(seed >> 32);
        ta.ctr = h->counters.as<unsigned long long>();
        ta.arena = h->arena.as<unsigned char>(); ta.arena_bytes = h->arena.cap;
        ta.desc = h->desc.as<unsigned long long>(); ta.defer_list = h->defer.as<long long>();
        ta.spos_out = h->locs.as<uint64_t>();
        ta.stage_cap = table_stage_cap(N, DIM, sim->L, reach);
        if (h->stage_cap_override > 0) ta.stage_cap = std::max(1, std::min(ta.stage_cap, h->stage_cap_override));
        ta.epw = entries_per_word(N); ta.eb = entry_bits(N);
        CUDA_TRY(h, h->stage.ensure(sizeof(uint16_t) * (size_t)N * (size_t)ta.stage_cap));
        ta.stage = h->stage.as<uint16_t>();
        CUDA_TRY(h, launch_table(ta, DIM, 1, s0));
        unsigned long long ctr[CTR_COUNT] = {};
        CUDA_TRY(h, cudaMemcpyAsync(ctr, h->counters.p, sizeof(ctr), cudaMemcpyDeviceToHost, s0));
        CUDA_TRY(h, cudaStreamSynchronize(s0));
        if (ctr[CTR_NDEFER] != 0) return fail(h, SPRB200_ERR_INTERNAL, "table did not fit a full-size arena");
        const long long nwords = (long long)ctr[CTR_MAX_ENTRIES];
        words.resize((size_t)std::max<long long>(nwords, 1));
        std::vector<uint64_t> spos((size_t)N);
        CUDA_TRY(h, cudaMemcpyAsync(offw.data(), h->arena.p, sizeof(uint32_t) * (N + 1), cudaMemcpyDeviceToHost, s0));
        if (nwords > 0)
            CUDA_TRY(h, cudaMemcpyAsync(words.data(), h->arena.as<unsigned char>() + o_words, sizeof(uint32_t) * nwords, cudaMemcpyDeviceToHost, s0));
        CUDA_TRY(h, cudaMemcpyAsync(spos.data(), h->locs.p, sizeof(uint64_t) * N, cudaMemcpyDeviceToHost, s0));
        CUDA_TRY(h, cudaStreamSynchronize(s0));
        if (positions_out) {
            const double scale = sim->L * 0x1.0p-21;
            for (int i = 0; i < N; ++i)
                for (int d = 0; d < DIM; ++d)
                    positions_out[(size_t)i * DIM + d] = (double)((spos[i] >> (21 * d)) & 0x1fffffu) * scale;
        }
    }
    unpack_record(N, offw.data(), words.data(), &off, &nbr);
    const long long total = (long long)nbr.size();
    for (int i = 0; i <= N; ++i) off_out[i] = (int32_t)off[i];
    if (nbr_out)
        for (long long q = 0; q < total && q < nbr_cap; ++q) nbr_out[q] = (int32_t)nbr[q];
    return total;
}

int sprb200_build_slice(sprb200_handle h, const SprGrid *grid, const SprSim *sim, const SprRun *run, int64_t idxstart,
                        int64_t idxend, double *out_TxNidx, int32_t *n_used, uint64_t *n_events) {
    if (!h) return SPRB200_ERR_BAD_ARG;
    if (!out_TxNidx || idxend < idxstart) return fail(h, SPRB200_ERR_BAD_ARG, "bad slice arguments");
    if (!sim || sim->T < 1) return fail(h, SPRB200_ERR_BAD_ARG, "sim is NULL");
    const int64_t count = idxend - idxstart + 1;
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, h->table.ensure(sizeof(double) * (size_t)count * sim->T));
    const std::vector<int64_t> idx = strided(idxstart, 1, count);
    int rc = build_rows(h, grid, sim, run, idx.data(), count, h->table.as<double>(), nullptr, nullptr, n_used, n_events);
    if (rc) return rc;
    // rows [count][T] are exactly the T x nidx column-major slice (time fastest)
    CUDA_TRY(h, cudaMemcpyAsync(out_TxNidx, h->table.p, sizeof(double) * (size_t)count * sim->T, cudaMemcpyDeviceToHost, h->stream[0]));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream[0]));
    return SPRB200_OK;
}

int sprb200_build_table(sprb200_handle h, const SprGrid *grid, const SprSim *sim, const SprRun *run, double *out_table,
                        int32_t *n_used, uint64_t *n_events) {
    if (!h) return SPRB200_ERR_BAD_ARG;
    int64_t P = 0;
    if (grid_total(grid, &P)) return fail(h, SPRB200_ERR_BAD_ARG, "bad grid");
    if (!out_table || !sim || sim->T < 1) return fail(h, SPRB200_ERR_BAD_ARG, "bad table arguments");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t elems = (size_t)P * sim->T;
    CUDA_TRY(h, h->rows.ensure(sizeof(double) * elems));
    const std::vector<int64_t> idx = strided(1, 1, P);
    int rc = build_rows(h, grid, sim, run, idx.data(), P, h->rows.as<double>(), nullptr, nullptr, n_used, n_events);
    if (rc) return rc;
    CUDA_TRY(h, h->table.ensure(sizeof(double) * elems));
    CUDA_TRY(h, launch_scatter(h->rows.as<double>(), 0, 1, P, sim->T, P, h->table.as<double>(), h->stream[0]));
    h->st_launches += 1;
    CUDA_TRY(h, cudaMemcpyAsync(out_table, h->table.p, sizeof(double) * elems, cudaMemcpyDeviceToHost, h->stream[0]));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream[0]));
    return SPRB200_OK;
}

int sprb200_merge_slice(const SprGrid *grid, int32_t T, const double *slice_TxNidx, int64_t idxstart, int64_t idxend,
                        double *table) {
    int64_t P = 0;
    if (grid_total(grid, &P) || !slice_TxNidx || !table || T < 1 || idxstart < 1 || idxend > P || idxend < idxstart)
        return SPRB200_ERR_BAD_ARG;
    for (int64_t idx = idxstart; idx <= idxend; ++idx) {
        const double *col = slice_TxNidx + (size_t)(idx - idxstart) * T;
        for (int32_t s = 0; s < T; ++s) table[(size_t)s * P + (size_t)(idx - 1)] = col[s];
    }
    return SPRB200_OK;
}

int sprb200_build_slice_device(sprb200_handle h, const SprGrid *grid, const SprSim *sim, const SprRun *run,
                               int64_t idxstart, int64_t stride, int64_t count, double *d_out_NidxxT, int32_t *d_n_used,
                               uint64_t *d_n_events) {
    if (!h) return SPRB200_ERR_BAD_ARG;
    if (!d_out_NidxxT) return fail(h, SPRB200_ERR_BAD_ARG, "d_out is NULL");
    if (count < 1 || stride < 1) return fail(h, SPRB200_ERR_BAD_ARG, "bad index range");
    const std::vector<int64_t> idx = strided(idxstart, stride, count);
    return build_rows(h, grid, sim, run, idx.data(), count, d_out_NidxxT, d_n_used, d_n_events, nullptr, nullptr);
}

int sprb200_build_indices_device(sprb200_handle h, const SprGrid *grid, const SprSim *sim, const SprRun *run,
                                 const int64_t *idx, int64_t count, double *d_out_NidxxT, int32_t *d_n_used,
                                 uint64_t *d_n_events) {
    if (!h) return SPRB200_ERR_BAD_ARG;
    if (!d_out_NidxxT) return fail(h, SPRB200_ERR_BAD_ARG, "d_out is NULL");
    return build_rows(h, grid, sim, run, idx, count, d_out_NidxxT, d_n_used, d_n_events, nullptr, nullptr);
}

int sprb200_build_indices(sprb200_handle h, const SprGrid *grid, const SprSim *sim, const SprRun *run,
                          const int64_t *idx, int64_t count, double *out_NidxxT, int32_t *n_used, uint64_t *n_events) {
    if (!h) return SPRB200_ERR_BAD_ARG;
    if (!out_NidxxT || count < 1 || !sim || sim->T < 1) return fail(h, SPRB200_ERR_BAD_ARG, "bad arguments");
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, h->table.ensure(sizeof(double) * (size_t)count * sim->T));
    int rc = build_rows(h, grid, sim, run, idx, count, h->table.as<double>(), nullptr, nullptr, n_used, n_events);
    if (rc) return rc;
    CUDA_TRY(h, cudaMemcpyAsync(out_NidxxT, h->table.p, sizeof(double) * (size_t)count * sim->T, cudaMemcpyDeviceToHost, h->stream[0]));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream[0]));
    return SPRB200_OK;
}

int sprb200_scatter_rows_device(sprb200_handle h, const double *d_rows, const int64_t *idx, int64_t count, int32_t T,
                                int64_t P, double *d_table) {
    if (!h) return SPRB200_ERR_BAD_ARG;
    if (!d_rows || !d_table || !idx || count < 1 || T < 1) return fail(h, SPRB200_ERR_BAD_ARG, "bad scatter arguments");
    for (int64_t k = 0; k < count; ++k)
        if (idx[k] < 1 || idx[k] > P) return fail(h, SPRB200_ERR_BAD_ARG, "index outside the table");
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, h->misc.ensure(sizeof(int64_t) * (size_t)count));
    CUDA_TRY(h, cudaMemcpyAsync(h->misc.p, idx, sizeof(int64_t) * (size_t)count, cudaMemcpyHostToDevice, h->stream[0]));
    CUDA_TRY(h, launch_scatter_idx(d_rows, h->misc.as<long long>(), count, T, P, d_table, h->stream[0]));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream[0]));
    return SPRB200_OK;
}

int sprb200_surrogate_load(sprb200_handle h, const SprGrid *grid, int32_t T, double t_first, double t_last,
                           const double *table, int32_t table_on_device) {
    if (!h) return SPRB200_ERR_BAD_ARG;
    int64_t P = 0;
    if (grid_total(grid, &P)) return fail(h, SPRB200_ERR_BAD_ARG, "bad grid");
    if (!table || T < 1 || !std::isfinite(t_first) || !std::isfinite(t_last) || (T > 1 && !(t_last > t_first)))
        return fail(h, SPRB200_ERR_BAD_ARG, "bad surrogate arguments");
    for (int k = 0; k < 4; ++k)
        if (!std::isfinite(grid->lo[k]) || !std::isfinite(grid->hi[k]) || (grid->n[k] > 1 && !(grid->hi[k] > grid->lo[k])))
            return fail(h, SPRB200_ERR_BAD_ARG, "surrogate axis ranges must be finite and increasing");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t bytes = sizeof(double) * (size_t)P * (size_t)T;
    CUDA_TRY(h, h->sur.ensure(bytes));
    CUDA_TRY(h, cudaMemcpyAsync(h->sur.p, table, bytes, table_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                                h->stream[0]));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream[0]));
    h->sur_grid = *grid;
    h->sur_T = T;
    h->sur_t0 = t_first;
    h->sur_t1 = t_last;
    h->sur_P = P;
    return SPRB200_OK;
}

int sprb200_fit_loss(sprb200_handle h, const SprAligned *data, const double *optpars, int64_t npop, double *loss) {
    if (!h) return SPRB200_ERR_BAD_ARG;
    if (h->sur_T < 1) return fail(h, SPRB200_ERR_BAD_ARG, "no surrogate loaded (sprb200_surrogate_load)");
    if (!data || data->nconc < 1 || !data->npts || !data->times || !data->values || !data->antibodyconcens)
        return fail(h, SPRB200_ERR_BAD_ARG, "bad aligned data");
    if (!optpars || !loss || npop < 1 || npop > 0x7fffffffLL) return fail(h, SPRB200_ERR_BAD_ARG, "bad population arguments");
    long long ns = 0;
    for (int j = 0; j < data->nconc; ++j) {
        if (data->npts[j] < 0) return fail(h, SPRB200_ERR_BAD_ARG, "negative sample count");
        if (!(data->antibodyconcens[j] > 0.0)) return fail(h, SPRB200_ERR_BAD_ARG, "antibody concentrations must be positive");
        ns += data->npts[j];
    }
    cudaStream_t s0 = h->stream[0];
    CUDA_TRY(h, cudaSetDevice(h->device));
    // packed upload: times | values | abcs | npts
    const size_t o_val = sizeof(double) * (size_t)ns, o_abc = 2 * o_val, o_npts = o_abc + sizeof(double) * data->nconc;
    const size_t total = o_npts + sizeof(int) * data->nconc;
    std::vector<unsigned char> pack(total + 8);
    if (ns) {
        std::memcpy(pack.data(), data->times, o_val);
        std::memcpy(pack.data() + o_val, data->values, o_val);
    }
    std::memcpy(pack.data() + o_abc, data->antibodyconcens, sizeof(double) * data->nconc);
    for (int j = 0; j < data->nconc; ++j) reinterpret_cast<int *>(pack.data() + o_npts)[j] = data->npts[j];
    CUDA_TRY(h, h->fitdata.ensure(total + 8));
    CUDA_TRY(h, h->fitpars.ensure(sizeof(double) * 5 * (size_t)npop));
    CUDA_TRY(h, h->fitloss.ensure(sizeof(double) * (size_t)npop));
    CUDA_TRY(h, cudaMemcpyAsync(h->fitdata.p, pack.data(), total, cudaMemcpyHostToDevice, s0));
    CUDA_TRY(h, cudaMemcpyAsync(h->fitpars.p, optpars, sizeof(double) * 5 * (size_t)npop, cudaMemcpyHostToDevice, s0));
    FitArgs a{};
    a.table = h->sur.as<double>();
    a.P = h->sur_P;
    for (int k = 0; k < 4; ++k) { a.n[k] = h->sur_grid.n[k]; a.lo[k] = h->sur_grid.lo[k]; a.hi[k] = h->sur_grid.hi[k]; }
    a.T = h->sur_T; a.t_first = h->sur_t0; a.t_last = h->sur_t1;
    a.nconc = data->nconc; a.nsamples = ns;
    a.times = h->fitdata.as<double>();
    a.values = reinterpret_cast<const double *>(h->fitdata.as<unsigned char>() + o_val);
    a.abcs = reinterpret_cast<const double *>(h->fitdata.as<unsigned char>() + o_abc);
    a.npts = reinterpret_cast<const int *>(h->fitdata.as<unsigned char>() + o_npts);
    a.optpars = h->fitpars.as<double>();
    a.loss = h->fitloss.as<double>();
    CUDA_TRY(h, launch_fit_loss(a, npop, s0));
    CUDA_TRY(h, cudaMemcpyAsync(loss, h->fitloss.p, sizeof(double) * (size_t)npop, cudaMemcpyDeviceToHost, s0));
    CUDA_TRY(h, cudaStreamSynchronize(s0));
    return SPRB200_OK;
}

int sprb200_scatter_table_device(sprb200_handle h, const double *d_rows, int64_t idxstart, int64_t stride, int64_t count,
                                 int32_t T, int64_t P, double *d_table) {
    if (!h) return SPRB200_ERR_BAD_ARG;
    if (!d_rows || !d_table || idxstart < 1 || stride < 1 || count < 1 || T < 1 || idxstart + (count - 1) * stride > P)
        return fail(h, SPRB200_ERR_BAD_ARG, "bad scatter arguments");
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, launch_scatter(d_rows, idxstart - 1, stride, count, T, P, d_table, h->stream[0]));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream[0]));
    return SPRB200_OK;
}

// ---- multi-GPU -------------------------------------------------------------------------------------
int64_t sprb200_deal_indices(const SprGrid *grid, const SprSim *sim, const SprRun *run, int32_t nranks, int32_t rank,
                             int64_t *idx_out, int64_t cap) {
    if (!grid || !sim || nranks < 1 || rank < 0 || rank >= nranks || sim->N < 1 || (sim->DIM != 2 && sim->DIM != 3) ||
        !(sim->L > 0.0))
        return SPRB200_ERR_BAD_ARG;
    std::vector<std::vector<int64_t>> deal;
    if (deal_points(grid, sim, run, nranks, &deal)) return SPRB200_ERR_BAD_ARG;
    const std::vector<int64_t> &mine = deal[(size_t)rank];
    if (idx_out)
        for (size_t k = 0; k < mine.size() && (int64_t)k < cap; ++k) idx_out[k] = mine[k];
    return (int64_t)mine.size();
}

int sprb200_build_table_multi(const int32_t *devices, int32_t ndev, const SprGrid *grid, const SprSim *sim,
                              const SprRun *run, double *out_table, int32_t *n_used, uint64_t *n_events) {
    if (!devices || ndev < 1 || ndev > 64) return fail(nullptr, SPRB200_ERR_BAD_ARG, "need 1..64 devices");
    int64_t P = 0;
    if (grid_total(grid, &P)) return fail(nullptr, SPRB200_ERR_BAD_ARG, "bad grid");
    if (!out_table || !sim || sim->T < 1) return fail(nullptr, SPRB200_ERR_BAD_ARG, "bad table arguments");
    for (int a = 0; a < ndev; ++a)
        for (int b = a + 1; b < ndev; ++b)
            if (devices[a] == devices[b]) return fail(nullptr, SPRB200_ERR_BAD_ARG, "device listed twice");
    std::lock_guard<std::mutex> lk(g_multi_mutex);     // one multi-GPU build at a time per process
    std::vector<sprb200_ctx *> hs((size_t)ndev, nullptr);
    for (int r = 0; r < ndev; ++r) {
        auto it = g_multi_handles.find(devices[r]);
        if (it == g_multi_handles.end()) {
            sprb200_handle nh = nullptr;
            const int rc = sprb200_create(devices[r], &nh);
            if (rc) return rc;      // text is in the thread-local create error
            it = g_multi_handles.emplace(devices[r], nh).first;
        }
        hs[(size_t)r] = it->second;
    }
    std::vector<std::vector<int64_t>> deal;
    if (deal_points(grid, sim, run, ndev, &deal)) return fail(nullptr, SPRB200_ERR_BAD_ARG, "bad grid");
    const int T = sim->T;
    sprb200_ctx *h0 = hs[0];
    // the gathered slabs and the table live on the first device; every other device sends its slab with
    // one peer copy (NVLink when peer access is available, staged by the driver otherwise)
    if (cudaSetDevice(h0->device) != cudaSuccess) return fail(nullptr, SPRB200_ERR_CUDA, "cudaSetDevice");
    const size_t elems = (size_t)P * (size_t)T;
    {
        cudaError_t e = h0->mrows.ensure(sizeof(double) * elems);
        if (e == cudaSuccess) e = h0->mtable.ensure(sizeof(double) * elems);
        if (e == cudaSuccess) e = h0->midx.ensure(sizeof(int64_t) * (size_t)P);
        if (e != cudaSuccess) { cudaGetLastError(); return fail(nullptr, SPRB200_ERR_ALLOC, "device memory for the gathered table"); }
    }
    for (int r = 1; r < ndev; ++r) {
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, h0->device, hs[(size_t)r]->device) == cudaSuccess && can) {
            cudaDeviceEnablePeerAccess(hs[(size_t)r]->device, 0);
            cudaGetLastError();     // already enabled is fine
        }
    }
    std::vector<int64_t> offs((size_t)ndev + 1, 0);
    for (int r = 0; r < ndev; ++r) offs[(size_t)r + 1] = offs[(size_t)r] + (int64_t)deal[(size_t)r].size();
    std::vector<int> rcs((size_t)ndev, SPRB200_OK);
    std::vector<std::vector<int32_t>> nus((size_t)ndev);
    std::vector<std::vector<uint64_t>> nevs((size_t)ndev);
    auto worker = [&](int r) {
        sprb200_ctx *h = hs[(size_t)r];
        const std::vector<int64_t> &idx = deal[(size_t)r];
        const int64_t count = (int64_t)idx.size();
        if (count == 0) return;
        auto body = [&]() -> int {
            CUDA_TRY(h, cudaSetDevice(h->device));
            CUDA_TRY(h, h->table.ensure(sizeof(double) * (size_t)count * T));
            nus[(size_t)r].assign((size_t)count, 0);
            nevs[(size_t)r].assign((size_t)count, 0);
            int rc = build_rows(h, grid, sim, run, idx.data(), count, h->table.as<double>(), nullptr, nullptr,
                                nus[(size_t)r].data(), nevs[(size_t)r].data());
            if (rc) return rc;
            CUDA_TRY(h, cudaMemcpyPeerAsync(h0->mrows.as<double>() + (size_t)offs[(size_t)r] * T, h0->device, h->table.p,
                                            h->device, sizeof(double) * (size_t)count * T, h->stream[0]));
            CUDA_TRY(h, cudaStreamSynchronize(h->stream[0]));
            return SPRB200_OK;
        };
        rcs[(size_t)r] = body();
    };
    {
        std::vector<std::thread> th;
        for (int r = 1; r < ndev; ++r) th.emplace_back(worker, r);
        worker(0);
        for (auto &t : th) t.join();
    }
    for (int r = 0; r < ndev; ++r)
        if (rcs[(size_t)r]) return fail(nullptr, rcs[(size_t)r], "device " + std::to_string(devices[r]) + ": " + hs[(size_t)r]->err);
    // slabs -> FirstMoment order on the first device, one copy to the caller's table
    std::vector<int64_t> idx_all;
    idx_all.reserve((size_t)P);
    for (int r = 0; r < ndev; ++r) idx_all.insert(idx_all.end(), deal[(size_t)r].begin(), deal[(size_t)r].end());
    auto tail = [&]() -> int {
        sprb200_ctx *h = h0;
        cudaStream_t s0 = h->stream[0];
        CUDA_TRY(h, cudaSetDevice(h->device));
        CUDA_TRY(h, cudaMemcpyAsync(h->midx.p, idx_all.data(), sizeof(int64_t) * (size_t)P, cudaMemcpyHostToDevice, s0));
        CUDA_TRY(h, launch_scatter_idx(h->mrows.as<double>(), h->midx.as<long long>(), P, T, P, h->mtable.as<double>(), s0));
        h->st_launches += 1;
        CUDA_TRY(h, cudaMemcpyAsync(out_table, h->mtable.p, sizeof(double) * elems, cudaMemcpyDeviceToHost, s0));
        CUDA_TRY(h, cudaStreamSynchronize(s0));
        return SPRB200_OK;
    };
    const int rc = tail();
    if (rc) return fail(nullptr, rc, h0->err);
    for (int r = 0; r < ndev; ++r)
        for (size_t k = 0; k < deal[(size_t)r].size(); ++k) {
            const int64_t i0 = deal[(size_t)r][k] - 1;
            if (n_used) n_used[i0] = nus[(size_t)r][k];
            if (n_events) n_events[i0] = nevs[(size_t)r][k];
        }
    return SPRB200_OK;
}

int sprb200_multi_stats(const int32_t *devices, int32_t ndev, uint64_t *events, double *device_ms, double *traj_ms) {
    if (!devices || ndev < 1) return SPRB200_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lk(g_multi_mutex);
    for (int r = 0; r < ndev; ++r) {
        auto it = g_multi_handles.find(devices[r]);
        if (it == g_multi_handles.end()) return SPRB200_ERR_BAD_ARG;
        if (events) events[r] = it->second->st_events;
        if (device_ms) device_ms[r] = it->second->st_ms;
        if (traj_ms) traj_ms[r] = it->second->st_traj_ms;
    }
    return SPRB200_OK;
}

void sprb200_multi_release(void) {
    std::lock_guard<std::mutex> lk(g_multi_mutex);
    for (auto &kv : g_multi_handles) sprb200_destroy(kv.second);
    g_multi_handles.clear();
}

int sprb200_comm_unique_id(unsigned char id_out[128]) {
    if (!id_out) return fail(nullptr, SPRB200_ERR_BAD_ARG, "id_out is NULL");
    std::string err;
    if (!nccl_load(&err)) return fail(nullptr, SPRB200_ERR_NCCL, err);
    NcclId id;
    const int r = g_nccl.GetUniqueId(&id);
    if (r != 0) return fail(nullptr, SPRB200_ERR_NCCL, std::string("ncclGetUniqueId: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error"));
    std::memcpy(id_out, id.internal, 128);
    return SPRB200_OK;
}

int sprb200_comm_init(sprb200_handle h, const unsigned char id[128], int32_t nranks, int32_t rank) {
    if (!h) return SPRB200_ERR_BAD_ARG;
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(h, SPRB200_ERR_BAD_ARG, "bad rank / nranks");
    if (h->nccl_comm) return fail(h, SPRB200_ERR_BAD_ARG, "this handle already has a communicator");
    h->comm_rank = rank;
    h->comm_nranks = nranks;
    if (nranks == 1) return SPRB200_OK;       // nothing to exchange
    if (!id) return fail(h, SPRB200_ERR_BAD_ARG, "id is NULL");
    std::string err;
    if (!nccl_load(&err)) return fail(h, SPRB200_ERR_NCCL, err);
    CUDA_TRY(h, cudaSetDevice(h->device));
    NcclId nid;
    std::memcpy(nid.internal, id, 128);
    NCCL_TRY(h, g_nccl.CommInitRank(&h->nccl_comm, nranks, nid, rank));
    return SPRB200_OK;
}

int sprb200_comm_destroy(sprb200_handle h) {
    if (!h) return SPRB200_ERR_BAD_ARG;
    if (h->nccl_comm) {
        cudaSetDevice(h->device);
        cudaStreamSynchronize(h->stream[0]);
        g_nccl.CommDestroy(h->nccl_comm);
        h->nccl_comm = nullptr;
    }
    h->comm_rank = 0;
    h->comm_nranks = 1;
    return SPRB200_OK;
}

int sprb200_build_table_dist(sprb200_handle h, const SprGrid *grid, const SprSim *sim, const SprRun *run,
                             double *out_table, double *d_table, int32_t *n_used, uint64_t *n_events) {
    if (!h) return SPRB200_ERR_BAD_ARG;
    int64_t P = 0;
    if (grid_total(grid, &P)) return fail(h, SPRB200_ERR_BAD_ARG, "bad grid");
    if (!sim || sim->T < 1) return fail(h, SPRB200_ERR_BAD_ARG, "sim is NULL");
    const int nr = h->comm_nranks, me = h->comm_rank;
    if (nr > 1 && !h->nccl_comm) return fail(h, SPRB200_ERR_BAD_ARG, "no communicator (sprb200_comm_init)");
    if (P < nr) return fail(h, SPRB200_ERR_BAD_ARG, "fewer grid points than ranks");
    std::vector<std::vector<int64_t>> deal;
    if (deal_points(grid, sim, run, nr, &deal)) return fail(h, SPRB200_ERR_BAD_ARG, "bad grid");
    const int T = sim->T;
    const std::vector<int64_t> &mine = deal[(size_t)me];
    const int64_t count = (int64_t)mine.size();
    int64_t maxcount = 0;
    for (auto &v : deal) maxcount = std::max<int64_t>(maxcount, (int64_t)v.size());
    cudaStream_t s0 = h->stream[0];
    CUDA_TRY(h, cudaSetDevice(h->device));
    // this rank's slab [maxcount][T] (rows past `count` are padding), its replicate and event counts
    CUDA_TRY(h, h->table.ensure(sizeof(double) * (size_t)maxcount * T));
    CUDA_TRY(h, h->mnused.ensure(sizeof(int32_t) * (size_t)maxcount));
    CUDA_TRY(h, h->mnev.ensure(sizeof(uint64_t) * (size_t)maxcount));
    if (count < maxcount) {
        CUDA_TRY(h, cudaMemsetAsync(h->table.as<double>() + (size_t)count * T, 0, sizeof(double) * (size_t)(maxcount - count) * T, s0));
        CUDA_TRY(h, cudaMemsetAsync(h->mnused.as<int32_t>() + count, 0, sizeof(int32_t) * (size_t)(maxcount - count), s0));
        CUDA_TRY(h, cudaMemsetAsync(h->mnev.as<uint64_t>() + count, 0, sizeof(uint64_t) * (size_t)(maxcount - count), s0));
    }
    int rc = build_rows(h, grid, sim, run, mine.data(), count, h->table.as<double>(), h->mnused.as<int32_t>(),
                        h->mnev.as<uint64_t>(), nullptr, nullptr);
    if (rc) return rc;
    // run_points keeps its own statistics; the gather below is part of the same call
    const double *d_rows_all = h->table.as<double>();
    const int32_t *d_nu_all = h->mnused.as<int32_t>();
    const uint64_t *d_nev_all = h->mnev.as<uint64_t>();
    if (nr > 1) {
        CUDA_TRY(h, h->mrows.ensure(sizeof(double) * (size_t)maxcount * T * nr));
        NCCL_TRY(h, g_nccl.AllGather(h->table.p, h->mrows.p, sizeof(double) * (size_t)maxcount * T, 0 /* ncclInt8 */,
                                     h->nccl_comm, s0));
        d_rows_all = h->mrows.as<double>();
        // the replicate and event counts are gathered on EVERY rank whatever outputs it asked for: the sequence of
        // collectives must not depend on per-rank arguments (12 bytes per point)
        CUDA_TRY(h, h->gnused.ensure(sizeof(int32_t) * (size_t)maxcount * nr));
        NCCL_TRY(h, g_nccl.AllGather(h->mnused.p, h->gnused.p, sizeof(int32_t) * (size_t)maxcount, 0, h->nccl_comm, s0));
        d_nu_all = h->gnused.as<int32_t>();
        CUDA_TRY(h, h->gnev.ensure(sizeof(uint64_t) * (size_t)maxcount * nr));
        NCCL_TRY(h, g_nccl.AllGather(h->mnev.p, h->gnev.p, sizeof(uint64_t) * (size_t)maxcount, 0, h->nccl_comm, s0));
        d_nev_all = h->gnev.as<uint64_t>();
        h->st_launches += 3;
    }
    // gathered rows -> FirstMoment order (padding rows carry index 0 and are skipped)
    std::vector<int64_t> idx_all((size_t)maxcount * nr, 0);
    for (int r = 0; r < nr; ++r)
        for (size_t k = 0; k < deal[(size_t)r].size(); ++k) idx_all[(size_t)r * maxcount + k] = deal[(size_t)r][k];
    CUDA_TRY(h, h->midx.ensure(sizeof(int64_t) * idx_all.size()));
    CUDA_TRY(h, cudaMemcpyAsync(h->midx.p, idx_all.data(), sizeof(int64_t) * idx_all.size(), cudaMemcpyHostToDevice, s0));
    double *d_out = d_table;
    if (!d_out) {
        CUDA_TRY(h, h->mtable.ensure(sizeof(double) * (size_t)P * T));
        d_out = h->mtable.as<double>();
    }
    CUDA_TRY(h, launch_scatter_idx(d_rows_all, h->midx.as<long long>(), (long long)idx_all.size(), T, P, d_out, s0));
    h->st_launches += 1;
    if (out_table) CUDA_TRY(h, cudaMemcpyAsync(out_table, d_out, sizeof(double) * (size_t)P * T, cudaMemcpyDeviceToHost, s0));
    std::vector<int32_t> nu_h;
    std::vector<uint64_t> nev_h;
    if (n_used) {
        nu_h.resize(idx_all.size());
        CUDA_TRY(h, cudaMemcpyAsync(nu_h.data(), d_nu_all, sizeof(int32_t) * idx_all.size(), cudaMemcpyDeviceToHost, s0));
    }
    if (n_events) {
        nev_h.resize(idx_all.size());
        CUDA_TRY(h, cudaMemcpyAsync(nev_h.data(), d_nev_all, sizeof(uint64_t) * idx_all.size(), cudaMemcpyDeviceToHost, s0));
    }
    CUDA_TRY(h, cudaStreamSynchronize(s0));
    for (size_t k = 0; k < idx_all.size(); ++k) {
        if (idx_all[k] < 1) continue;
        if (n_used) n_used[idx_all[k] - 1] = nu_h[k];
        if (n_events) n_events[idx_all[k] - 1] = nev_h[k];
    }
    return SPRB200_OK;
}

}  // extern "C"
