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
#include <map>
#include <string>
#include <vector>

using namespace spr;

namespace {

constexpr int NSTREAM = 4;
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
    int smem_optin = 0;
    cudaStream_t stream[NSTREAM] = {};
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr, ev_ready = nullptr, ev_done[NSTREAM] = {};
    std::string err;
    // arenas
    DevBuf tsave, pts, pidx, order, curve, curve2, sum, sumsq, nused, nevents, nscanned, nstar, active, counters,
        ovf, rows, locs, goff, gnbr, misc, table, nbrscratch, nbrscratch2;
    // stats of the last call
    uint64_t st_traj = 0, st_events = 0, st_launches = 0;
    double st_ms = 0.0;
    size_t curve_budget = (size_t)8 << 30;
    double cap_scale = 1.0;   // SPRB200_CAP_SCALE (test hook): shrinks the a-priori table capacity
    // Tables of at least this many entries live in global memory (L2-resident, one slice per CTA) instead
    // of shared memory: the extra L2 latency per row is outweighed by 16 instead of 5-14 resident
    // trajectories per SM.  Measured on B200 (profiles/README.md): global wins for every reach class of
    // the surrogate grids, so the default is 0; SPRB200_NBR_GLOBAL_MIN_CAP overrides (1e9 = always shared).
    long long nbr_global_min_cap = 0;
    size_t nbr_scratch_budget = (size_t)6 << 30;
    int max_bps = 0;          // SPRB200_MAX_BPS (experiment hook): cap on resident trajectories per SM
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

// capacity (directed neighbour entries) that a trajectory at this reach needs with overwhelming
// probability: mean + 8 sigma of 2 x (pair count); overflowing trajectories are re-run wider.
int capacity_for(int N, int DIM, double L, double reach, double cap_scale = 1.0) {
    const double PI = 3.14159265358979323846;
    double frac = DIM == 3 ? (4.0 / 3.0) * PI * reach * reach * reach / (L * L * L) : PI * reach * reach / (L * L);
    if (frac > 1.0) frac = 1.0;
    const double E = (double)N * (double)(N - 1) * frac;
    double cap = E + 8.0 * std::sqrt(2.0 * E) + 64.0;
    const double maxe = (double)N * (double)(N - 1);
    if (cap > maxe) cap = maxe;
    cap *= cap_scale;
    long long c = (long long)std::ceil(cap);
    c = (c + 511) / 512 * 512;
    if (c < 1024) c = 1024;
    if (c > 0x7fffffffLL) c = 0x7fffffffLL;
    return (int)c;
}

// rough SSA event count of one trajectory (only used to start expensive points first)
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

struct RunPlan {
    // inputs
    const SprSim *sim;
    const SprRun *run;
    int observable;        // OBS_*
    bool raw;              // keep curves (simulate_raw)
};

struct ClassWork {
    SmemLayout lay;
    std::vector<int32_t> slots;   // sorted expensive-first
};

// table placement for a class: shared memory if it fits and the table is small, else global memory
bool choose_layout(const sprb200_ctx *h, int N, int DIM, int cap, SmemLayout *out) {
    const SmemLayout ls = make_layout(N, DIM, cap, false);
    const SmemLayout lg = make_layout(N, DIM, cap, true);
    if (ls.total <= h->smem_optin && (long long)cap < h->nbr_global_min_cap) { *out = ls; return true; }
    if (lg.total <= h->smem_optin) { *out = lg; return true; }
    if (ls.total <= h->smem_optin) { *out = ls; return true; }
    *out = lg;
    return false;
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
    const int maxrep = variance ? run->maxsims : run->nsims;
    cudaStream_t s0 = h->stream[0];

    h->st_traj = h->st_events = h->st_launches = 0;
    h->st_ms = 0.0;
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, h->tsave.ensure(sizeof(double) * T));
    CUDA_TRY(h, cudaMemcpyAsync(h->tsave.p, sim->tsave, sizeof(double) * T, cudaMemcpyHostToDevice, s0));

    // ---- fixed positions: one shared neighbour table per distinct reach (K2b)
    struct FixedGraph { uint32_t *off; uint16_t *nbr; long long entries; int id; };
    std::map<double, FixedGraph> graphs;
    std::vector<void *> graph_allocs;
    auto free_graphs = [&]() {
        for (void *p : graph_allocs) cudaFree(p);
        graph_allocs.clear();
    };
    if (!sim->resample_initlocs) {
        CUDA_TRY(h, h->locs.ensure(sizeof(double) * N * DIM));
        CUDA_TRY(h, cudaMemcpyAsync(h->locs.p, sim->initlocs, sizeof(double) * N * DIM, cudaMemcpyHostToDevice, s0));
        CUDA_TRY(h, h->misc.ensure(64));
        for (int64_t k = 0; k < npts; ++k) {
            const double reach = pts[k].reach;
            if (graphs.count(reach)) continue;
            FixedGraph g{nullptr, nullptr, 0, (int)graphs.size() + 1};
            cudaError_t e = cudaMalloc(&g.off, sizeof(uint32_t) * (N + 1));
            if (e != cudaSuccess) { free_graphs(); return fail(h, SPRB200_ERR_ALLOC, "cudaMalloc(fixed graph offsets)"); }
            graph_allocs.push_back(g.off);
            e = launch_fixed_graph(h->locs.as<double>(), N, DIM, sim->L, reach, g.off, nullptr, 0,
                                   h->misc.as<unsigned long long>(), s0);
            unsigned long long total = 0;
            if (e == cudaSuccess) e = cudaMemcpyAsync(&total, h->misc.p, 8, cudaMemcpyDeviceToHost, s0);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s0);
            if (e != cudaSuccess) { free_graphs(); return fail(h, SPRB200_ERR_CUDA, std::string("fixed graph: ") + cudaGetErrorString(e)); }
            h->st_launches += 2;
            g.entries = (long long)total;
            e = cudaMalloc(&g.nbr, sizeof(uint16_t) * std::max<long long>(g.entries, 1));
            if (e != cudaSuccess) { free_graphs(); return fail(h, SPRB200_ERR_ALLOC, "cudaMalloc(fixed graph entries)"); }
            graph_allocs.push_back(g.nbr);
            e = launch_fixed_graph(h->locs.as<double>(), N, DIM, sim->L, reach, g.off, g.nbr, g.entries, nullptr, s0);
            if (e != cudaSuccess) { free_graphs(); return fail(h, SPRB200_ERR_CUDA, std::string("fixed graph fill: ") + cudaGetErrorString(e)); }
            h->st_launches += 1;
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

    CUDA_TRY(h, cudaEventRecord(h->ev_begin, s0));
    std::vector<int32_t> h_nstar, h_active, h_order;
    std::vector<unsigned long long> h_nev;

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
        // active slots of this batch (all of them at first)
        h_active.resize(nb);
        for (int64_t k = 0; k < nb; ++k) h_active[k] = (int32_t)k;

        const int rbase = run->replicate_base;
        int rep_lo = 0;                                  // relative to replicate_base
        int rep_hi = variance ? run->minsims : run->nsims;
        while (!h_active.empty()) {
            // ---- partition the active slots into shared-memory capacity classes
            std::map<std::pair<int, int>, ClassWork> classes;   // key: (capacity, fixed-graph id)
            for (int32_t sl : h_active) {
                const PointDev &p = pts[b0 + sl];
                int cap, gid = 0;
                if (sim->resample_initlocs) cap = capacity_for(N, DIM, sim->L, p.reach, h->cap_scale);
                else {
                    const FixedGraph &g = graphs[p.reach];
                    cap = (int)std::min<long long>(std::max<long long>((g.entries + 63) / 64 * 64, 64), 0x7fffffff);
                    gid = g.id;
                }
                ClassWork &cw = classes[std::make_pair(cap, gid)];
                cw.slots.push_back(sl);
            }
            // ---- one persistent launch per class, round-robin over the streams
            const size_t nclass = classes.size();
            CUDA_TRY(h, h->counters.ensure(sizeof(unsigned long long) * 2 * nclass + 64));
            CUDA_TRY(h, cudaMemsetAsync(h->counters.p, 0, sizeof(unsigned long long) * 2 * nclass + 64, s0));
            const unsigned int ovf_cap = 1u << 16;
            CUDA_TRY(h, h->ovf.ensure(sizeof(long long) * ovf_cap * nclass));
            // order arrays for all classes packed into one upload
            h_order.clear();
            std::vector<size_t> class_off;
            // widest tables first: those classes run the fewest trajectories per SM and finish last, the
            // cheap small-reach classes then fill the SMs as the wide ones drain
            std::vector<std::pair<const std::pair<int, int>, ClassWork> *> ordered;
            for (auto &kv : classes) ordered.push_back(&kv);
            std::reverse(ordered.begin(), ordered.end());
            for (auto *kvp : ordered) {
                ClassWork &cw = kvp->second;
                std::vector<std::pair<double, int32_t>> keyed;
                keyed.reserve(cw.slots.size());
                for (int32_t sl : cw.slots)
                    keyed.emplace_back(-cost_estimate(pts[b0 + sl], N, DIM, sim->L, sim->tstop, sim->tstop_AtoB), sl);
                std::sort(keyed.begin(), keyed.end());
                class_off.push_back(h_order.size());
                for (auto &kk : keyed) h_order.push_back(kk.second);
            }
            CUDA_TRY(h, cudaMemcpyAsync(h->order.p, h_order.data(), sizeof(int32_t) * h_order.size(),
                                        cudaMemcpyHostToDevice, s0));
            CUDA_TRY(h, cudaEventRecord(h->ev_ready, s0));

            struct Launched { int cap; size_t ci; SmemLayout lay; TrajArgs args; };
            std::vector<Launched> launched;
            struct Prep { int cap; ClassWork *cw; SmemLayout lay; long long nblocks; size_t scratch_off; };
            std::vector<Prep> preps;
            size_t scratch_total = 0;
            for (auto *kvp : ordered) {
                const int cap = kvp->first.first;
                ClassWork &cw = kvp->second;
                SmemLayout lay;
                if (!choose_layout(h, N, DIM, cap, &lay)) {
                    free_graphs();
                    return fail(h, SPRB200_ERR_TOO_LARGE,
                                "per-trajectory state (N=" + std::to_string(N) + ", table of " + std::to_string(cap) +
                                    " entries) needs " + std::to_string(lay.total) + " B of shared memory, limit " +
                                    std::to_string(h->smem_optin));
                }
                int bps = traj_blocks_per_sm(lay, DIM);
                if (bps < 1) { free_graphs(); return fail(h, SPRB200_ERR_TOO_LARGE, "trajectory kernel does not fit on an SM"); }
                if (h->max_bps > 0 && bps > h->max_bps) bps = h->max_bps;
                const long long ntick = (long long)cw.slots.size() * (rep_hi - rep_lo);
                long long nblocks = std::min<long long>(ntick, (long long)bps * h->num_sms);
                size_t off = 0;
                if (lay.nbr_global) {
                    const size_t per = (size_t)lay.cap * sizeof(uint16_t);
                    const long long fit = (long long)std::max<size_t>(1, h->nbr_scratch_budget / per);
                    nblocks = std::min(nblocks, fit);
                    off = scratch_total;
                    scratch_total += (size_t)nblocks * per;
                    scratch_total = (scratch_total + 255) & ~(size_t)255;
                }
                preps.push_back(Prep{cap, &cw, lay, nblocks, off});
            }
            if (scratch_total) CUDA_TRY(h, h->nbrscratch.ensure(scratch_total));
            size_t ci = 0;
            for (auto &pr : preps) {
                const int cap = pr.cap;
                ClassWork &cw = *pr.cw;
                const SmemLayout &lay = pr.lay;
                TrajArgs a{};
                a.N = N; a.T = T; a.L = sim->L; a.tstop = sim->tstop; a.toff = sim->tstop_AtoB;
                a.tsave = h->tsave.as<double>();
                a.pts = h->pts.as<PointDev>();
                a.pidx = h->pidx.as<uint32_t>();
                a.order = h->order.as<int32_t>() + class_off[ci];
                a.rep0 = rbase + rep_lo; a.nrep = rep_hi - rep_lo; a.repstride = maxrep; a.rep_rowbase = rbase;
                a.ntickets = (long long)cw.slots.size() * a.nrep;
                a.ticket_list = nullptr;
                a.seed_lo = (uint32_t)(run->seed & 0xffffffffu); a.seed_hi = (uint32_t)(run->seed >> 32);
                a.ticket_ctr = h->counters.as<unsigned long long>() + 2 * ci;
                a.curve = h->curve.as<uint16_t>();
                a.curve2 = planes == 2 ? h->curve2.as<uint16_t>() : nullptr;
                a.nevents = h->nevents.as<unsigned long long>();
                a.observable = plan.observable;
                a.ovf_list = h->ovf.as<long long>() + (size_t)ovf_cap * ci;
                a.ovf_count = reinterpret_cast<unsigned int *>(h->counters.as<unsigned long long>() + 2 * ci + 1);
                a.ovf_cap = ovf_cap;
                a.nbr_scratch = lay.nbr_global ? reinterpret_cast<uint16_t *>(h->nbrscratch.as<unsigned char>() + pr.scratch_off) : nullptr;
                if (!sim->resample_initlocs) {
                    const FixedGraph &g = graphs[pts[b0 + cw.slots[0]].reach];
                    a.g_off = g.off; a.g_nbr = g.nbr;
                }
                a.lay = lay;
                cudaStream_t st = h->stream[ci % NSTREAM];
                if (st != s0) CUDA_TRY(h, cudaStreamWaitEvent(st, h->ev_ready, 0));
                CUDA_TRY(h, launch_traj(a, DIM, (int)pr.nblocks, st));
                h->st_launches += 1;
                h->st_traj += (uint64_t)a.ntickets;
                launched.push_back(Launched{cap, ci, lay, a});
                ++ci;
            }
            for (int s = 1; s < NSTREAM; ++s) {
                CUDA_TRY(h, cudaEventRecord(h->ev_done[s], h->stream[s]));
                CUDA_TRY(h, cudaStreamWaitEvent(s0, h->ev_done[s], 0));
            }
            // ---- capacity overflows: re-run those tickets with a wider table
            std::vector<unsigned long long> h_cnt(2 * nclass);
            CUDA_TRY(h, cudaMemcpyAsync(h_cnt.data(), h->counters.p, sizeof(unsigned long long) * 2 * nclass,
                                        cudaMemcpyDeviceToHost, s0));
            CUDA_TRY(h, cudaStreamSynchronize(s0));
            for (auto &lc : launched) {
                unsigned int novf = (unsigned int)(h_cnt[2 * lc.ci + 1] & 0xffffffffu);
                int cap = lc.cap;
                while (novf > 0) {
                    if (novf > ovf_cap) { free_graphs(); return fail(h, SPRB200_ERR_INTERNAL, "too many capacity overflows"); }
                    const long long maxe = (long long)N * (N - 1);
                    if (cap >= maxe) { free_graphs(); return fail(h, SPRB200_ERR_INTERNAL, "overflow at full capacity"); }
                    cap = (int)std::min<long long>(maxe + 64, (long long)cap + cap / 4 + 512);
                    SmemLayout lay;
                    if (!choose_layout(h, N, DIM, cap, &lay)) {
                        free_graphs();
                        return fail(h, SPRB200_ERR_TOO_LARGE, "per-trajectory state does not fit in shared memory after overflow");
                    }
                    // move the overflow list aside (misc) and reuse the class's counters
                    CUDA_TRY(h, h->misc.ensure(sizeof(long long) * novf + 64));
                    CUDA_TRY(h, cudaMemcpyAsync(h->misc.p, lc.args.ovf_list, sizeof(long long) * novf,
                                                cudaMemcpyDeviceToDevice, s0));
                    CUDA_TRY(h, cudaMemsetAsync(lc.args.ticket_ctr, 0, 2 * sizeof(unsigned long long), s0));
                    TrajArgs a = lc.args;
                    a.lay = lay;
                    a.ticket_list = h->misc.as<long long>();
                    a.ntickets = novf;
                    const int bps = traj_blocks_per_sm(lay, DIM);
                    if (bps < 1) { free_graphs(); return fail(h, SPRB200_ERR_TOO_LARGE, "trajectory kernel does not fit on an SM"); }
                    long long nblk = std::min<long long>(novf, (long long)bps * h->num_sms);
                    if (lay.nbr_global) {
                        const size_t per = (size_t)lay.cap * sizeof(uint16_t);
                        nblk = std::min<long long>(nblk, (long long)std::max<size_t>(1, h->nbr_scratch_budget / per));
                        CUDA_TRY(h, h->nbrscratch2.ensure((size_t)nblk * per));
                        a.nbr_scratch = h->nbrscratch2.as<uint16_t>();
                    } else
                        a.nbr_scratch = nullptr;
                    CUDA_TRY(h, launch_traj(a, DIM, (int)nblk, s0));
                    h->st_launches += 1;
                    unsigned long long cnt2[2] = {0, 0};
                    CUDA_TRY(h, cudaMemcpyAsync(cnt2, lc.args.ticket_ctr, sizeof(cnt2), cudaMemcpyDeviceToHost, s0));
                    CUDA_TRY(h, cudaStreamSynchronize(s0));
                    novf = (unsigned int)(cnt2[1] & 0xffffffffu);
                }
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
    h->smem_optin = max_dynamic_smem(device_ordinal);
    for (int s = 0; s < NSTREAM; ++s) {
        if (cudaStreamCreateWithFlags(&h->stream[s], cudaStreamNonBlocking) != cudaSuccess) { delete h; return fail(nullptr, SPRB200_ERR_CUDA, "cudaStreamCreate"); }
        cudaEventCreateWithFlags(&h->ev_done[s], cudaEventDisableTiming);
    }
    cudaEventCreate(&h->ev_begin);
    cudaEventCreate(&h->ev_end);
    cudaEventCreateWithFlags(&h->ev_ready, cudaEventDisableTiming);
    if (const char *env = std::getenv("SPRB200_CURVE_BYTES")) {
        const long long v = std::atoll(env);
        if (v > 0) h->curve_budget = (size_t)v;
    }
    if (const char *env = std::getenv("SPRB200_NBR_GLOBAL_MIN_CAP")) {
        const long long v = std::atoll(env);
        if (v >= 0) h->nbr_global_min_cap = v;
    }
    if (const char *env = std::getenv("SPRB200_MAX_BPS")) h->max_bps = std::atoi(env);
    if (const char *env = std::getenv("SPRB200_CAP_SCALE")) {
        const double v = std::atof(env);
        if (v > 0.0 && v <= 1.0) h->cap_scale = v;
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
    DevBuf *bufs[] = {&h->tsave, &h->pts, &h->pidx, &h->order, &h->curve, &h->curve2, &h->sum, &h->sumsq, &h->nused,
                      &h->nevents, &h->nscanned, &h->nstar, &h->active, &h->counters, &h->ovf, &h->rows, &h->locs,
                      &h->goff, &h->gnbr, &h->misc, &h->table, &h->nbrscratch, &h->nbrscratch2};
    for (DevBuf *b : bufs) b->release();
    for (int s = 0; s < NSTREAM; ++s) {
        if (h->ev_done[s]) cudaEventDestroy(h->ev_done[s]);
        if (h->stream[s]) cudaStreamDestroy(h->stream[s]);
    }
    if (h->ev_begin) cudaEventDestroy(h->ev_begin);
    if (h->ev_end) cudaEventDestroy(h->ev_end);
    if (h->ev_ready) cudaEventDestroy(h->ev_ready);
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
    std::vector<uint32_t> off(N + 1);
    std::vector<uint16_t> nbr;
    long long total = 0;
    if (!sim->resample_initlocs) {
        CUDA_TRY(h, h->locs.ensure(sizeof(double) * N * DIM));
        CUDA_TRY(h, cudaMemcpyAsync(h->locs.p, sim->initlocs, sizeof(double) * N * DIM, cudaMemcpyHostToDevice, s0));
        CUDA_TRY(h, h->goff.ensure(sizeof(uint32_t) * (N + 1)));
        CUDA_TRY(h, h->misc.ensure(64));
        CUDA_TRY(h, launch_fixed_graph(h->locs.as<double>(), N, DIM, sim->L, reach, h->goff.as<uint32_t>(), nullptr, 0,
                                       h->misc.as<unsigned long long>(), s0));
        unsigned long long tot = 0;
        CUDA_TRY(h, cudaMemcpyAsync(&tot, h->misc.p, 8, cudaMemcpyDeviceToHost, s0));
        CUDA_TRY(h, cudaStreamSynchronize(s0));
        total = (long long)tot;
        CUDA_TRY(h, h->gnbr.ensure(sizeof(uint16_t) * std::max<long long>(total, 1)));
        CUDA_TRY(h, launch_fixed_graph(h->locs.as<double>(), N, DIM, sim->L, reach, h->goff.as<uint32_t>(),
                                       h->gnbr.as<uint16_t>(), total, nullptr, s0));
        nbr.resize((size_t)std::max<long long>(total, 1));
        CUDA_TRY(h, cudaMemcpyAsync(off.data(), h->goff.p, sizeof(uint32_t) * (N + 1), cudaMemcpyDeviceToHost, s0));
        CUDA_TRY(h, cudaMemcpyAsync(nbr.data(), h->gnbr.p, sizeof(uint16_t) * total, cudaMemcpyDeviceToHost, s0));
        CUDA_TRY(h, cudaStreamSynchronize(s0));
        if (positions_out) std::memcpy(positions_out, sim->initlocs, sizeof(double) * N * DIM);
    } else {
        if (param_index > 0xffffffffULL) return fail(h, SPRB200_ERR_BAD_ARG, "parameter index exceeds 2^32");
        int cap = capacity_for(N, DIM, sim->L, reach);
        for (;;) {
            SmemLayout lay = make_layout(N, DIM, cap);
            if (lay.total > h->smem_optin) return fail(h, SPRB200_ERR_TOO_LARGE, "neighbour table does not fit in shared memory");
            PointDev p{0, 0, 0, reach};
            uint32_t pi = (uint32_t)param_index;
            CUDA_TRY(h, h->pts.ensure(sizeof(PointDev)));
            CUDA_TRY(h, h->pidx.ensure(sizeof(uint32_t)));
            CUDA_TRY(h, cudaMemcpyAsync(h->pts.p, &p, sizeof(p), cudaMemcpyHostToDevice, s0));
            CUDA_TRY(h, cudaMemcpyAsync(h->pidx.p, &pi, sizeof(pi), cudaMemcpyHostToDevice, s0));
            CUDA_TRY(h, h->goff.ensure(sizeof(uint32_t) * (N + 1)));
            CUDA_TRY(h, h->gnbr.ensure(sizeof(uint16_t) * (size_t)lay.cap));
            CUDA_TRY(h, h->locs.ensure(sizeof(uint64_t) * N + sizeof(double) * N * DIM));
            CUDA_TRY(h, h->misc.ensure(64));
            TrajArgs a{};
            a.N = N; a.T = sim->T; a.L = sim->L;
            a.pts = h->pts.as<PointDev>(); a.pidx = h->pidx.as<uint32_t>();
            a.seed_lo = (uint32_t)(seed & 0xffffffffu); a.seed_hi = (uint32_t)(seed >> 32);
            a.lay = lay;
            CUDA_TRY(h, launch_table_dump(a, DIM, replicate, 0, h->goff.as<uint32_t>(), h->gnbr.as<uint16_t>(),
                                          h->locs.as<uint64_t>(), h->misc.as<long long>(), s0));
            CUDA_TRY(h, cudaMemcpyAsync(&total, h->misc.p, 8, cudaMemcpyDeviceToHost, s0));
            CUDA_TRY(h, cudaStreamSynchronize(s0));
            if (total >= 0) break;
            const long long maxe = (long long)N * (N - 1);
            if (cap >= maxe) return fail(h, SPRB200_ERR_INTERNAL, "overflow at full capacity");
            cap = (int)std::min<long long>(maxe + 64, (long long)cap + cap / 4 + 512);
        }
        nbr.resize((size_t)std::max<long long>(total, 1));
        std::vector<uint64_t> spos((size_t)N);
        CUDA_TRY(h, cudaMemcpyAsync(off.data(), h->goff.p, sizeof(uint32_t) * (N + 1), cudaMemcpyDeviceToHost, s0));
        CUDA_TRY(h, cudaMemcpyAsync(nbr.data(), h->gnbr.p, sizeof(uint16_t) * total, cudaMemcpyDeviceToHost, s0));
        CUDA_TRY(h, cudaMemcpyAsync(spos.data(), h->locs.p, sizeof(uint64_t) * N, cudaMemcpyDeviceToHost, s0));
        CUDA_TRY(h, cudaStreamSynchronize(s0));
        if (positions_out) {
            const double scale = sim->L * 0x1.0p-21;
            for (int i = 0; i < N; ++i)
                for (int d = 0; d < DIM; ++d)
                    positions_out[(size_t)i * DIM + d] = (double)((spos[i] >> (21 * d)) & 0x1fffffu) * scale;
        }
    }
    for (int i = 0; i <= N; ++i) off_out[i] = (int32_t)off[i];
    if (nbr_out)
        for (long long q = 0; q < total && q < nbr_cap; ++q) nbr_out[q] = nbr[q];
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

}  // extern "C"
