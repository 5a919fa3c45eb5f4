// spr_kernels.cu -- hand-written sm_100a kernels for the SSA ensemble (DESIGN.md sections 3-5).
//
// K1  spr_traj_kernel      one trajectory per warp (one single-warp CTA), persistent CTAs pulling
//                          (parameter point, replicate) tickets from a global counter.  Replaces the
//                          replicate loop + event loop of run_spr_sim!
//                          (/root/reference/src/forward_simulator.jl:129-361).
// K2  build_table_*        neighbour table within `reach` on the periodic box from a cell list, in
//                          the prologue of K1; replaces setup_spr_sim! (forward_simulator.jl:37-89).
// K2b fixed_graph_*        the same table for user-supplied FP64 positions (resample_initlocs =
//                          false), built once per call with the reference's periodic_dist_sq (:1-8).
// K3  reduce_fixed / variance_scan   replicate reduction and the VarianceTerminator's sequential
//                          stopping rule (/root/reference/src/callbacks.jl:23-33, 204-217).
// K4  means / scatter      mean curves in slice order and the [P][T] -> [T][P] permutation into the
//                          "FirstMoment" layout (/root/reference/src/surrogate.jl:232, 292, 315).
//
// Nothing here is a dense contraction: no tensor cores.  The trajectory is a dependent chain of
// shared-memory reads/writes, shuffles and ALU ops; throughput comes from keeping as many
// trajectories resident per SM as shared memory allows.
#include "spr_device.cuh"
#include "spr_kernels.cuh"
#include <type_traits>

namespace spr {

namespace {

constexpr uint32_t FULL = 0xffffffffu;

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

template <int DIM>
__device__ __forceinline__ uint64_t gen_position(uint32_t i, uint32_t rep, uint32_t pidx, uint32_t k0, uint32_t k1) {
    const u32x4 r = philox4x32_10(i, STREAM_POSITIONS, rep, pidx, k0, k1);
    uint64_t p = (uint64_t)(r.x >> (32 - POS_BITS)) | ((uint64_t)(r.y >> (32 - POS_BITS)) << POS_BITS);
    if (DIM == 3) p |= (uint64_t)(r.z >> (32 - POS_BITS)) << (2 * POS_BITS);
    return p;
}

__device__ __forceinline__ uint32_t pos_axis(uint64_t p, int d) { return (uint32_t)(p >> (d * POS_BITS)) & POS_MASK; }

template <int DIM>
__device__ __forceinline__ uint32_t cell_of(uint64_t p, uint32_t nc) {
    uint32_t c = 0;
#pragma unroll
    for (int d = DIM - 1; d >= 0; --d) c = c * nc + ((pos_axis(p, d) * nc) >> POS_BITS);
    return c;
}

// Squared min-image distance in lattice units, exact in 64-bit integers (each axis < 2^20, so the sum
// is < 2^42), compared with r2lat = floor(reach^2 / (L 2^-21)^2): the reference's test
// periodic_dist_sq <= reach^2 (forward_simulator.jl:1-8, 54-55) evaluated without rounding on the lattice.
template <int DIM>
__device__ __forceinline__ bool within_reach(uint64_t a, uint64_t b, unsigned long long r2lat) {
    unsigned long long acc = 0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
        const uint32_t df = (pos_axis(a, d) - pos_axis(b, d)) & POS_MASK;
        const uint32_t nd = (0u - df) & POS_MASK;
        const uint32_t m = df < nd ? df : nd;
        acc += (unsigned long long)m * m;
    }
    return acc <= r2lat;
}

// r2lat of the specification: floor(reach^2 / scale^2) with scale = L 2^-21, clamped to 2^62
__host__ __device__ inline unsigned long long reach2_lattice(double L, double reach) {
    const double scale = L * 0x1.0p-21;
    const double q = (reach * reach) / (scale * scale);   // separate IEEE operations (no contraction)
    if (!(q < 4611686018427387904.0)) return 4611686018427387904ULL;
    return (unsigned long long)q;                        // truncation == floor for q >= 0
}

// in-place exclusive prefix sum of a[0..n) (values < 2^31 in total), returns the total.
// Each lane scans a contiguous block serially, lane totals are combined with a warp scan.
template <typename TT>
__device__ __forceinline__ int warp_exclusive_scan_inplace(TT *a, int n, int lane, int *overflow_limit_total) {
    const int per = (n + 31) >> 5;
    const int b = lane * per;
    const int e = min(b + per, n);
    int s = 0;
    for (int i = b; i < e; ++i) s += (int)a[i];
    const int incl = warp_incl_scan(s, lane, 32);
    const int total = __shfl_sync(FULL, incl, 31);
    if (overflow_limit_total && total > *overflow_limit_total) return total;  // caller bails out
    int run = incl - s;
    for (int i = b; i < e; ++i) {
        const int v = (int)a[i];
        a[i] = (TT)run;
        run += v;
    }
    return total;
}

struct Smem {
    uint16_t *nbr;
    void *off;
    uint32_t *sn;
    uint16_t *listA, *listB, *pos;
    uint16_t *h16;     // = listA; listB and pos are h16 + oB, h16 + oP (index offsets in u16 units)
    int oB, oP;
    uint32_t *listC;   // last word of the array listA lives in: complexes grow downwards from it
    uint64_t *spos;
    uint16_t *cstart;
    uint16_t *cfill;
    uint4 *rng;
};

template <bool NBRG>
__device__ __forceinline__ Smem carve(unsigned char *base, const SmemLayout &l, uint16_t *nbr_scratch) {
    Smem s;
    s.nbr = NBRG ? nbr_scratch + (size_t)blockIdx.x * (size_t)l.cap : reinterpret_cast<uint16_t *>(base);
    s.cfill = reinterpret_cast<uint16_t *>(base + l.o_cfill);
    s.off = base + l.o_off;
    s.sn = reinterpret_cast<uint32_t *>(base + l.o_sn);
    s.listA = reinterpret_cast<uint16_t *>(base + l.o_listA);
    s.listB = reinterpret_cast<uint16_t *>(base + l.o_listB);
    s.pos = reinterpret_cast<uint16_t *>(base + l.o_pos);
    s.h16 = s.listA;
    s.oB = (l.o_listB - l.o_listA) / 2;
    s.oP = (l.o_pos - l.o_listA) / 2;
    s.listC = reinterpret_cast<uint32_t *>(base + l.o_listC);
    s.spos = reinterpret_cast<uint64_t *>(base + l.o_spos);
    s.cstart = reinterpret_cast<uint16_t *>(base + l.o_cstart);
    s.rng = reinterpret_cast<uint4 *>(base + l.o_rng);
    return s;
}

// ------------------------------------------------------------------------------------------------
// K2: neighbour table of one trajectory with freshly sampled positions.
//   positions: particle i (0..N-1) gets lattice coordinates Philox(i, STREAM_POSITIONS, rep, pidx),
//              x = k * L * 2^-32 in [0, L)          (forward_simulator.jl:42-47: L*rand())
//   numbering: particles are renumbered by (cell id, i) with a stable counting sort
//   table:     neighbours of particle i in (dz,dy,dx) lexicographic cell order, ascending sorted
//              index inside a cell                   (forward_simulator.jl:52-61, 71-86)
// Returns the number of directed entries, or -1 when it exceeds `cap`.
// ------------------------------------------------------------------------------------------------
template <int DIM, typename OffT>
__device__ int build_table_resampled(const Smem &sm, int N, double L, double reach, int cap, uint32_t rep,
                                     uint32_t pidx, uint32_t k0, uint32_t k1, int lane) {
    OffT *off = reinterpret_cast<OffT *>(sm.off);
    const int nc = cells_per_axis(N, DIM, L, reach);
    int ncell = nc;
#pragma unroll
    for (int d = 1; d < DIM; ++d) ncell *= nc;
    const unsigned long long r2lat = reach2_lattice(L, reach);

    for (int c = lane; c <= ncell; c += 32) sm.cstart[c] = 0;
    __syncwarp();
    // pass 1: particles per cell (count of cell c kept at cstart[c+1])
    for (int base = 0; base < N; base += 32) {
        const int i = base + lane;
        uint32_t cell = 0x10000u + (uint32_t)lane;
        if (i < N) cell = cell_of<DIM>(gen_position<DIM>((uint32_t)i, rep, pidx, k0, k1), (uint32_t)nc);
        const uint32_t mask = __match_any_sync(FULL, cell);
        if (i < N && (__ffs(mask) - 1) == lane) sm.cstart[cell + 1] += (uint16_t)__popc(mask);
        __syncwarp();
    }
    // exclusive scan: cstart[c] = first sorted index of cell c, cstart[ncell] = N
    {
        // counts sit at cstart[1..ncell]; scanning them in place (inclusive into the same slots)
        const int per = (ncell + 31) >> 5;
        const int b = lane * per, e = min(b + per, ncell);
        int s = 0;
        for (int c = b; c < e; ++c) s += sm.cstart[c + 1];
        const int incl = warp_incl_scan(s, lane, 32);
        int run = incl - s;
        for (int c = b; c < e; ++c) {
            run += sm.cstart[c + 1];
            sm.cstart[c + 1] = (uint16_t)run;
        }
    }
    __syncwarp();
    for (int c = lane; c < ncell; c += 32) sm.cfill[c] = sm.cstart[c];
    __syncwarp();
    // pass 2: stable scatter into cell order
    for (int base = 0; base < N; base += 32) {
        const int i = base + lane;
        uint32_t cell = 0x10000u + (uint32_t)lane;
        uint64_t pk = 0;
        if (i < N) {
            pk = gen_position<DIM>((uint32_t)i, rep, pidx, k0, k1);
            cell = cell_of<DIM>(pk, (uint32_t)nc);
        }
        const uint32_t mask = __match_any_sync(FULL, cell);
        int slot = 0;
        if (i < N) slot = sm.cfill[cell] + __popc(mask & lanemask_lt());
        __syncwarp();
        if (i < N) {
            sm.spos[slot] = pk;
            if ((__ffs(mask) - 1) == lane) sm.cfill[cell] += (uint16_t)__popc(mask);
        }
        __syncwarp();
    }
    // pass 3: the table, one particle per warp iteration.  The candidates of particle i are the
    // particles of the 3^DIM cells around it in (dz,dy,dx) order; the three x-cells of a (dz,dy) row are
    // contiguous in the sorted numbering except across the periodic wrap, so they form nr <= 2*3^(DIM-1)
    // index ranges.  One lane per range computes (first index, length); a warp scan turns the lengths
    // into positions in the flattened candidate list, the ranges are expanded into a 256-entry window
    // (it lives in the idle rng block), and all 32 lanes then test consecutive candidates and append
    // the hits in order with a ballot prefix: no second counting pass and no idle lanes.  Consecutive
    // particles of the same cell reuse the expanded list.
    uint16_t *cand = reinterpret_cast<uint16_t *>(sm.rng);
    constexpr int WIN = 256;
    const int nrow = nc >= 3 ? (DIM == 3 ? 9 : 3) : 1;
    int running = 0;
    if (lane == 0) off[0] = 0;
    int prev_cell = -1, start = 0, len = 0, jb = 0, ctot = 0, mxl = 0, nr = 1;
    bool cand_ok = false;
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
        const uint64_t ki = sm.spos[i];
        const int cx = (int)((pos_axis(ki, 0) * (uint32_t)nc) >> POS_BITS);
        const int cy = (int)((pos_axis(ki, 1) * (uint32_t)nc) >> POS_BITS);
        const int cz = DIM == 3 ? (int)((pos_axis(ki, 2) * (uint32_t)nc) >> POS_BITS) : 0;
        const int cell = (cz * nc + cy) * nc + cx;
        if (cell != prev_cell) {
            prev_cell = cell;
            cand_ok = false;
            const bool edge = nc >= 3 && (cx == 0 || cx == nc - 1);
            nr = nc < 3 ? 1 : (edge ? 2 * nrow : nrow);
            len = 0;
            jb = 0;
            if (lane < nr) {
                int ca = 0, cb = 1;
                if (nc >= 3) {
                    const int r = edge ? lane >> 1 : lane;
                    const int half = edge ? lane & 1 : 0;
                    int yy = cy + (DIM == 3 ? r % 3 : r) - 1;
                    yy = yy < 0 ? yy + nc : (yy >= nc ? yy - nc : yy);
                    int zz = 0;
                    if (DIM == 3) {
                        zz = cz + r / 3 - 1;
                        zz = zz < 0 ? zz + nc : (zz >= nc ? zz - nc : zz);
                    }
                    const int rb = (zz * nc + yy) * nc;
                    if (cx == 0) {
                        if (half == 0) { ca = rb + nc - 1; cb = rb + nc; } else { ca = rb; cb = rb + 2; }
                    } else if (cx == nc - 1) {
                        if (half == 0) { ca = rb + nc - 2; cb = rb + nc; } else { ca = rb; cb = rb + 1; }
                    } else {
                        ca = rb + cx - 1; cb = rb + cx + 2;
                    }
                }
                jb = sm.cstart[ca];
                len = (int)sm.cstart[cb] - jb;
            }
            const int incl = warp_incl_scan(len, lane, 32);
            start = incl - len;
            ctot = __shfl_sync(FULL, incl, 31);
            mxl = __reduce_max_sync(FULL, len);
        }
#pragma unroll 1
        for (int w0 = 0; w0 < ctot; w0 += WIN) {
            const int wend = min(ctot - w0, WIN);
            if (nr > 1 && !cand_ok) {
                __syncwarp();
#pragma unroll 1
                for (int tt = 0; tt < mxl; ++tt) {
                    const int pq = start + tt - w0;
                    if (tt < len && (unsigned)pq < (unsigned)WIN) cand[pq] = (uint16_t)(jb + tt);
                }
                __syncwarp();
                cand_ok = ctot <= WIN;      // a single window stays valid for the next particles of the cell
            }
#pragma unroll 1
            for (int c0 = 0; c0 < wend; c0 += 32) {
                const int f = c0 + lane;
                int j = 0;
                bool hit = false;
                if (f < wend) {
                    j = nr > 1 ? (int)cand[f] : w0 + f;       // a single range [0, N): the candidate is its own index
                    hit = j != i && within_reach<DIM>(ki, sm.spos[j], r2lat);
                }
                const uint32_t bal = __ballot_sync(FULL, hit);
                const int nh = __popc(bal);
                if (running + nh > cap) return -1;
                if (hit) sm.nbr[running + __popc(bal & lanemask_lt())] = (uint16_t)j;
                running += nh;
            }
        }
        if (lane == 0) off[i + 1] = (OffT)running;
    }
    __syncwarp();
    return running;
}

// ------------------------------------------------------------------------------------------------
// K1: the trajectory.  All scalars (time, counts, list lengths) are warp-uniform registers; the 32
// lanes share the neighbour updates, list scans, bin writes and random-number generation.
// ------------------------------------------------------------------------------------------------

// Neighbour updates are split into a load (row offset, degree, this lane's neighbour) and an apply
// (read-modify-write of the neighbour's state word), so that the loads of both members of a pair
// event are in flight together.  When the partner of a pair event (`other`) is met, its own state
// change `extra` rides on the same store: pair events need no separate state write.
struct NbrRow {
    int o0, dg, j;   // first entry, degree, neighbour handled by this lane (valid if lane < dg)
};

template <typename OffT>
__device__ __forceinline__ NbrRow load_row(const uint16_t *nbr_lane, const OffT *off, int m, int lane) {
    NbrRow r;
    r.o0 = (int)off[m];
    r.dg = (int)off[m + 1] - r.o0;
    r.j = lane < r.dg ? (int)nbr_lane[r.o0] : 0;     // nbr_lane = nbr + lane
    return r;
}

template <bool PAIR>
__device__ __forceinline__ void apply_row(const Smem &sm, const NbrRow &r, uint32_t delta, int other, uint32_t extra,
                                          int lane) {
    if (lane < r.dg) {
        uint32_t d = delta;
        if (PAIR && r.j == other) d += extra;
        sm.sn[r.j] += d;
    }
    if (__builtin_expect(r.dg > 32, 0)) {
#pragma unroll 1
        for (int q = lane + 32; q < r.dg; q += 32) {
            const int j = sm.nbr[r.o0 + q];
            uint32_t d = delta;
            if (PAIR && j == other) d += extra;
            sm.sn[j] += d;
        }
    }
}

// inclusive scan over the first nact (<= 32) lanes, unrolled with warp-uniform early exit
__device__ __forceinline__ int scan_small(int v, int lane, int nact) {
    if (nact > 1) {
        int o = __shfl_up_sync(FULL, v, 1);
        if (lane >= 1) v += o;
        if (nact > 2) {
            o = __shfl_up_sync(FULL, v, 2);
            if (lane >= 2) v += o;
            if (nact > 4) {
                o = __shfl_up_sync(FULL, v, 4);
                if (lane >= 4) v += o;
                if (nact > 8) {
                    o = __shfl_up_sync(FULL, v, 8);
                    if (lane >= 8) v += o;
                    if (nact > 16) {
                        o = __shfl_up_sync(FULL, v, 16);
                        if (lane >= 16) v += o;
                    }
                }
            }
        }
    }
    return v;
}

template <int DIM, typename OffT, bool NBRG>
__global__ void __launch_bounds__(32) spr_traj_kernel(const TrajArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    const Smem sm = carve<NBRG>(smem_raw, a.lay, a.nbr_scratch);
    const OffT *off = reinterpret_cast<const OffT *>(sm.off);
    const uint16_t *nbr_lane = sm.nbr + lane;
    const int N = a.N, T = a.T;
    const double tstop = a.tstop, toff = a.toff;
    const double *__restrict__ tsave = a.tsave;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);

    for (;;) {
        long long ticket = 0;
        if (lane == 0) ticket = (long long)atomicAdd(a.ticket_ctr, 1ULL);
        ticket = __shfl_sync(FULL, ticket, 0);
        if (ticket >= a.ntickets) break;
        const long long tk = a.ticket_list ? a.ticket_list[ticket] : ticket;
        const int slot = a.order[tk / a.nrep];
        const int rep = a.rep0 + (int)(tk % a.nrep);
        const PointDev pt = a.pts[slot];
        const uint32_t pidx = a.pidx[slot];

        // ---------------- neighbour table
        int entries;
        if (a.g_off) {
            OffT *offw = reinterpret_cast<OffT *>(sm.off);
            for (int i = lane; i <= N; i += 32) offw[i] = (OffT)a.g_off[i];
            entries = (int)a.g_off[N];
            for (int q = lane; q < entries; q += 32) sm.nbr[q] = a.g_nbr[q];
        } else {
            entries = build_table_resampled<DIM, OffT>(sm, N, a.L, pt.reach, a.lay.cap, (uint32_t)rep, pidx,
                                                       a.seed_lo, a.seed_hi, lane);
        }
        __syncwarp();
        if (entries < 0) {  // table larger than this launch's capacity: hand the ticket back
            if (lane == 0) {
                const unsigned int w = atomicAdd(a.ovf_count, 1u);
                if (w < a.ovf_cap) a.ovf_list[w] = tk;
            }
            continue;
        }
        // ---------------- initial state: everything A (forward_simulator.jl:148-149)
        for (int i = lane; i < N; i += 32) {
            const uint32_t deg = (uint32_t)off[i + 1] - (uint32_t)off[i];
            sm.sn[i] = ST_A | (deg << SH_NA);
            sm.listA[i] = (uint16_t)i;
            sm.pos[i] = (uint16_t)i;
        }
        __syncwarp();

        int cntA = N, cntB = 0, cntC = 0, nAB = 0;
        double t = 0.0;
        double kon = pt.kon_eff;
        const double koff = pt.koff, konb = pt.konb;
        const double kc = __dmul_rn(2.0, koff);           // C -> A+B rate (forward_simulator.jl:135)
        bool on = true;
        bool last = false;
        int sidx = 0;
        double tp = tsave[0];
        // fast-path deadline: an event strictly before it needs no save-slot write, no switch-off and is
        // not the last one
        double tfast = fmin(fmin(tp, toff), tstop);
        uint32_t ndraw = 0, nskip = 0;
        const size_t row = (size_t)slot * a.repstride + (size_t)(rep - a.rep_rowbase);
        uint16_t *crow = a.curve + row * (size_t)T;
        uint16_t *crow2 = a.curve2 ? a.curve2 + row * (size_t)T : nullptr;
        const int obs = a.observable;

        for (;;) {
            // ---- randoms of this draw: {E = -log U (double), w2, w3}, generated 32 draws at a time
            if ((ndraw & 31u) == 0u) {
                const u32x4 r = philox4x32_10(ndraw + (uint32_t)lane, STREAM_EVENTS, (uint32_t)rep, pidx,
                                              a.seed_lo, a.seed_hi);
                const uint64_t k53 = (((uint64_t)r.x << 21) | (uint64_t)(r.y >> 11)) + 1ULL;
                const double En = neg_log_u53(k53);
                uint4 v;
                v.x = (uint32_t)__double2loint(En);
                v.y = (uint32_t)__double2hiint(En);
                v.z = r.z;
                v.w = r.w;
                __syncwarp();
                sm.rng[lane] = v;
                __syncwarp();
            }
            const uint4 rv = sm.rng[ndraw & 31u];
            ++ndraw;
            const double E = __hiloint2double((int)rv.y, (int)rv.x);

            // ---- total propensity (exact integer counts x rates)
            const double cA = __dmul_rn(kon, (double)cntA);
            const double cB = __dadd_rn(cA, __dmul_rn(koff, (double)cntB));
            const double cP = __dadd_rn(cB, __dmul_rn(konb, (double)nAB));
            const double a0 = __dadd_rn(cP, __dmul_rn(kc, (double)cntC));
            double tn;
            if (rcp_in_range(a0)) tn = __dadd_rn(t, event_dt_fast(E, a0));
            else tn = a0 > 0.0 ? __dadd_rn(t, __ddiv_rn(E, a0)) : INF;

            if (!(tn < tfast)) {
                // ---- slow path: antibody bath removed at tstop_AtoB (forward_simulator.jl:174-181,
                // 199-202), save slots strictly before the next event keep the current state (:188-194)
                bool switched = false;
                if (on && tn >= toff) {
                    tn = toff;
                    switched = true;
                }
                if (tn > tp) {
                    const int xb = cntB + cntC;
                    const uint16_t x1 = (uint16_t)(obs == OBS_TOTAL_A ? cntA : (obs == OBS_RAW_BC ? cntB : xb));
                    int cnt;
                    do {
                        const int s = sidx + lane;
                        const bool pred = (s < T) && (tsave[s] < tn);
                        cnt = __popc(__ballot_sync(FULL, pred));
                        if (pred) {
                            crow[s] = x1;
                            if (crow2) crow2[s] = (uint16_t)cntC;
                        }
                        sidx += cnt;
                    } while (cnt == 32);
                    tp = sidx < T ? tsave[sidx] : INF;
                }
                if (switched) {
                    t = toff;
                    kon = 0.0;
                    on = false;
                    ++nskip;
                    tfast = fmin(tp, tstop);
                    continue;
                }
                if (!(tn < INF)) {   // nothing can fire any more; all slots are filled
                    ++nskip;
                    break;
                }
                last = tn > tstop;   // the first event past tstop is still applied (:173, 183)
                tfast = fmin(fmin(tp, on ? toff : INF), tstop);
            }
            t = tn;

            // ---- channel selection
            const double x = __dmul_rn(__dmul_rn((double)rv.z, 0x1.0p-32), a0);
            if (x < cB) {
                // A --> B (forward_simulator.jl:204-224) when x < cA, else B --> A (:228-253): the same
                // code with the two lists swapped
                const bool ab = x < cA;
                const int ofrom = ab ? 0 : sm.oB, oto = ab ? sm.oB : 0;
                const int nfrom = ab ? cntA : cntB, nto = ab ? cntB : cntA;
                const int k = (int)__umulhi(rv.w, (uint32_t)nfrom);
                const int m = sm.h16[ofrom + k];
                const int lastm = sm.h16[ofrom + nfrom - 1];
                sm.h16[ofrom + k] = (uint16_t)lastm;
                sm.h16[sm.oP + lastm] = (uint16_t)k;
                sm.h16[oto + nto] = (uint16_t)m;
                sm.h16[sm.oP + m] = (uint16_t)nto;
                const int dAB = ab ? 1 : -1;
                cntA -= dAB;
                cntB += dAB;
                const uint32_t v = sm.sn[m];
                nAB += dAB * (sn_nA(v) - sn_nB(v));
                sm.sn[m] = (v & ~3u) | (ab ? ST_B : ST_A);          // idempotent: every lane stores the same word
                const NbrRow row = load_row<OffT>(nbr_lane, off, m, lane);
                const uint32_t dlt = (1u << SH_NB) - (1u << SH_NA);
                apply_row<false>(sm, row, ab ? dlt : 0u - dlt, 0, 0u, lane);
                __syncwarp();
            } else if (x < cP) {
                // A + B --> C (:256-289): pair number k of the nAB active pairs, enumerated along the
                // shorter of the A and B lists (weights nB resp. nA), then along the neighbour list.
                int k = (int)__umulhi(rv.w, (uint32_t)nAB);
                const bool sideA = cntA <= cntB;
                const uint16_t *list = sideA ? sm.listA : sm.listB;
                const int len = sideA ? cntA : cntB;
                const int sh = sideA ? SH_NB : SH_NA;
                const uint32_t want = sideA ? ST_B : ST_A;
                int xsel = 0, ix = 0;
                for (int c0 = 0; c0 < len; c0 += 32) {
                    const int idx = c0 + lane;
                    const int nact = min(32, len - c0);
                    int xm = 0, w = 0;
                    if (idx < len) {
                        xm = list[idx];
                        w = (int)((sm.sn[xm] >> sh) & CNT_MASK);
                    }
                    const int incl = scan_small(w, lane, nact);
                    const int tot = nact > 1 ? __shfl_sync(FULL, incl, nact - 1) : __shfl_sync(FULL, w, 0);
                    if (k < tot) {
                        const uint32_t hit = __ballot_sync(FULL, k < incl);
                        const int Lh = __ffs(hit) - 1;
                        xsel = __shfl_sync(FULL, xm, Lh);
                        k -= __shfl_sync(FULL, incl - w, Lh);
                        ix = c0 + Lh;
                        break;
                    }
                    k -= tot;
                }
                // k-th neighbour of xsel in state `want`
                int ysel = 0;
                {
                    const int o0 = (int)off[xsel], o1 = (int)off[xsel + 1];
                    for (int q0 = o0; q0 < o1; q0 += 32) {
                        const int q = q0 + lane;
                        int j = 0;
                        bool ok = false;
                        if (q < o1) {
                            j = sm.nbr[q];
                            ok = sn_state(sm.sn[j]) == want;
                        }
                        uint32_t bal = __ballot_sync(FULL, ok);
                        const int c = __popc(bal);
                        if (k < c) {
                            for (int r = 0; r < k; ++r) bal &= bal - 1u;
                            ysel = __shfl_sync(FULL, j, __ffs(bal) - 1);
                            break;
                        }
                        k -= c;
                    }
                }
                const int ma = sideA ? xsel : ysel;   // the A member
                const int mb = sideA ? ysel : xsel;   // the B member
                const int ia = sideA ? ix : (int)sm.pos[ma];
                const int ib = sideA ? (int)sm.pos[mb] : ix;
                // remove from the A and B lists, append to the complex list
                {
                    const int lastA = sm.listA[cntA - 1];
                    sm.listA[ia] = (uint16_t)lastA;
                    sm.pos[lastA] = (uint16_t)ia;
                    --cntA;
                    const int lastB = sm.listB[cntB - 1];
                    sm.listB[ib] = (uint16_t)lastB;
                    sm.pos[lastB] = (uint16_t)ib;
                    --cntB;
                    *(sm.listC - cntC) = (uint32_t)ma | ((uint32_t)mb << 16);
                    ++cntC;
                }
                const uint32_t va = sm.sn[ma], vb = sm.sn[mb];
                const NbrRow rowa = load_row<OffT>(nbr_lane, off, ma, lane);
                const NbrRow rowb = load_row<OffT>(nbr_lane, off, mb, lane);
                nAB -= sn_nB(va) + sn_nA(vb) - 1;
                __syncwarp();   // every lane has read va, vb before the partner words change
                // a's neighbours lose an A neighbour, and b (one of them) drops B(2) -> C(0);
                // b's neighbours lose a B neighbour, and a drops A(1) -> C(0)
                apply_row<true>(sm, rowa, 0u - (1u << SH_NA), mb, 0u - 2u, lane);
                __syncwarp();
                apply_row<true>(sm, rowb, 0u - (1u << SH_NB), ma, 0u - 1u, lane);
                __syncwarp();
            } else {
                // C --> A + B (:291-344): complex k; the freed end is chosen by a fair coin (:300-307)
                const int k = (int)__umulhi(rv.w, (uint32_t)cntC);
                const uint32_t pr = *(sm.listC - k);
                *(sm.listC - k) = *(sm.listC - (cntC - 1));
                --cntC;
                int ma = (int)(pr & 0xffffu), mb = (int)(pr >> 16);
                if (rv.w & 1u) {
                    const int tmp = ma;
                    ma = mb;
                    mb = tmp;
                }
                sm.listA[cntA] = (uint16_t)ma;
                sm.pos[ma] = (uint16_t)cntA;
                ++cntA;
                sm.listB[cntB] = (uint16_t)mb;
                sm.pos[mb] = (uint16_t)cntB;
                ++cntB;
                const uint32_t va = sm.sn[ma], vb = sm.sn[mb];
                const NbrRow rowa = load_row<OffT>(nbr_lane, off, ma, lane);
                const NbrRow rowb = load_row<OffT>(nbr_lane, off, mb, lane);
                nAB += sn_nB(va) + sn_nA(vb) + 1;
                __syncwarp();
                // a's neighbours gain an A neighbour, and b becomes B: C(0) -> B(2);
                // b's neighbours gain a B neighbour, and a becomes A: C(0) -> A(1)
                apply_row<true>(sm, rowa, 1u << SH_NA, mb, 2u, lane);
                __syncwarp();
                apply_row<true>(sm, rowb, 1u << SH_NB, ma, 1u, lane);
                __syncwarp();
            }
            if (last) break;
        }
        // ---- remaining save slots take the final state (forward_simulator.jl:349-354)
        {
            const int xb = cntB + cntC;
            const uint16_t x1 = (uint16_t)(obs == OBS_TOTAL_A ? cntA : (obs == OBS_RAW_BC ? cntB : xb));
            for (int s = sidx + lane; s < T; s += 32) {
                crow[s] = x1;
                if (crow2) crow2[s] = (uint16_t)cntC;
            }
        }
        if (lane == 0) atomicAdd(a.nevents + slot, (unsigned long long)(ndraw - nskip));
        __syncwarp();
    }
}

// one-trajectory table dump (test hook): same device functions as K1
template <int DIM, typename OffT>
__global__ void __launch_bounds__(32) spr_table_dump_kernel(const TrajArgs a, uint32_t rep, int slot,
                                                            uint32_t *off_out, uint16_t *nbr_out,
                                                            uint64_t *spos_out, long long *total_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    const Smem sm = carve<false>(smem_raw, a.lay, nullptr);
    const PointDev pt = a.pts[slot];
    const int entries = build_table_resampled<DIM, OffT>(sm, a.N, a.L, pt.reach, a.lay.cap, rep, a.pidx[slot],
                                                         a.seed_lo, a.seed_hi, lane);
    __syncwarp();
    if (lane == 0) *total_out = entries;
    if (entries < 0) return;
    const OffT *off = reinterpret_cast<const OffT *>(sm.off);
    for (int i = lane; i <= a.N; i += 32) off_out[i] = (uint32_t)off[i];
    for (int q = lane; q < entries; q += 32) nbr_out[q] = sm.nbr[q];
    for (int q = lane; q < a.N; q += 32) spos_out[q] = sm.spos[q];
}

// ------------------------------------------------------------------------------------------------
// K2b: neighbour table for fixed FP64 positions, the reference's arithmetic (d += min(|dx|, L-|dx|)^2
// with separate roundings, forward_simulator.jl:1-8), neighbours in ascending particle index.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double ref_dist_sq(const double *p, const double *q, int DIM, double L) {
    double d = 0.0;
    for (int k = 0; k < DIM; ++k) {
        const double dax = fabs(__dadd_rn(p[k], -q[k]));
        const double m = fmin(dax, __dadd_rn(L, -dax));
        d = __dadd_rn(d, __dmul_rn(m, m));
    }
    return d;
}

__global__ void fixed_graph_count_kernel(const double *locs, int N, int DIM, double L, double reach2,
                                         uint32_t *deg) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    uint32_t c = 0;
    for (int j = 0; j < N; ++j)
        if (j != i && ref_dist_sq(locs + (size_t)i * DIM, locs + (size_t)j * DIM, DIM, L) <= reach2) ++c;
    deg[i] = c;
}

__global__ void fixed_graph_scan_kernel(uint32_t *deg_off, int N, unsigned long long *total) {
    // single block; serial chunks + shared scan (N <= 65535, run once per call)
    __shared__ unsigned long long part[1024];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int per = (N + nt - 1) / nt;
    const int b = tid * per, e = min(b + per, N);
    unsigned long long s = 0;
    for (int i = b; i < e; ++i) s += deg_off[i];
    part[tid] = s;
    __syncthreads();
    if (tid == 0) {
        unsigned long long run = 0;
        for (int k = 0; k < nt; ++k) {
            const unsigned long long v = part[k];
            part[k] = run;
            run += v;
        }
        *total = run;
        deg_off[N] = (uint32_t)(run > 0xffffffffULL ? 0xffffffffULL : run);
    }
    __syncthreads();
    unsigned long long run = part[tid];
    for (int i = b; i < e; ++i) {
        const uint32_t v = deg_off[i];
        deg_off[i] = (uint32_t)run;
        run += v;
    }
}

__global__ void fixed_graph_fill_kernel(const double *locs, int N, int DIM, double L, double reach2,
                                        const uint32_t *off, uint16_t *nbr, long long cap) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    long long w = off[i];
    for (int j = 0; j < N; ++j)
        if (j != i && ref_dist_sq(locs + (size_t)i * DIM, locs + (size_t)j * DIM, DIM, L) <= reach2) {
            if (w < cap) nbr[w] = (uint16_t)j;
            ++w;
        }
}

// ------------------------------------------------------------------------------------------------
// K3: replicate reduction (fixed count) -- sum and sum of squares per (slot, bin) over nrep curves
// ------------------------------------------------------------------------------------------------
__global__ void reduce_fixed_kernel(const uint16_t *__restrict__ curve, int repstride, int nrep, int T,
                                    long long *__restrict__ sum, unsigned long long *__restrict__ sumsq) {
    const int slot = blockIdx.x;
    const uint16_t *base = curve + (size_t)slot * repstride * T;
    for (int s = threadIdx.x; s < T; s += blockDim.x) {
        long long S = 0;
        unsigned long long Q = 0;
        for (int r = 0; r < nrep; ++r) {
            const unsigned int v = base[(size_t)r * T + s];
            S += v;
            Q += (unsigned long long)v * v;
        }
        sum[(size_t)slot * T + s] = S;
        sumsq[(size_t)slot * T + s] = Q;
    }
}

// K3 (VarianceTerminator): continue the sequential scan of replicates nscanned+1 .. navail of every
// active slot; nstar = first n >= minsims with std <= ssetol*mean*sqrt(n) in every bin, or maxsims
// (callbacks.jl:204-217).  With integer moments S = sum x, Q = sum x^2 the bin test
// std > ssetol*mean*sqrt(n) is  n*Q - S^2 > ssetol^2 * S^2 * (n-1)  (scale CP/N cancels).
__global__ void variance_scan_kernel(const uint16_t *__restrict__ curve, const int32_t *__restrict__ active,
                                     int repstride, int T, int navail, double ssetol, int minsims,
                                     int maxsims, long long *__restrict__ sum,
                                     unsigned long long *__restrict__ sumsq, int32_t *__restrict__ nscanned,
                                     int32_t *__restrict__ nstar) {
    const int slot = active[blockIdx.x];
    const uint16_t *base = curve + (size_t)slot * repstride * T;
    const int n0 = nscanned[slot];
    const double tol2 = ssetol * ssetol;
    // each thread owns bins threadIdx.x + k*blockDim.x; running moments stay in global between stages
    int found = 0;
    int n = n0;
    while (n < navail) {
        ++n;
        int viol = 0;
        for (int s = threadIdx.x; s < T; s += blockDim.x) {
            const unsigned int v = base[(size_t)(n - 1) * T + s];
            const long long S = sum[(size_t)slot * T + s] + v;
            const unsigned long long Q = sumsq[(size_t)slot * T + s] + (unsigned long long)v * v;
            sum[(size_t)slot * T + s] = S;
            sumsq[(size_t)slot * T + s] = Q;
            if (n >= minsims && n < maxsims) {
                const double dS = (double)S;
                const double num = (double)((unsigned long long)n * Q - (unsigned long long)(S * S));
                if (num > tol2 * dS * dS * (double)(n - 1)) viol = 1;
            }
        }
        if (n < minsims) continue;                 // uniform across the block
        if (n >= maxsims) { found = n; break; }
        const int any = __syncthreads_or(viol);
        if (!any) { found = n; break; }
    }
    if (threadIdx.x == 0) {
        nscanned[slot] = n;
        nstar[slot] = found;   // 0 = needs more replicates
    }
}

// K4
__global__ void means_kernel(const long long *__restrict__ sum, const int32_t *__restrict__ nused, int T, int N,
                             double *__restrict__ rows) {
    const int slot = blockIdx.x;
    const double den = (double)nused[slot] * (double)N;
    for (int s = threadIdx.x; s < T; s += blockDim.x)
        rows[(size_t)slot * T + s] = (double)sum[(size_t)slot * T + s] / den;
}

// rows [count][T] -> table[s*P + idx] with a 32x32 shared tile (coalesced on both sides when stride==1)
__global__ void scatter_kernel(const double *__restrict__ rows, long long idx0, long long stride, long long count,
                               int T, long long P, double *__restrict__ table) {
    __shared__ double tile[32][33];
    const long long k0 = (long long)blockIdx.x * 32;
    const int s0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const long long k = k0 + r;
        const int s = s0 + threadIdx.x;
        if (k < count && s < T) tile[r][threadIdx.x] = rows[(size_t)k * T + s];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int s = s0 + r;
        const long long k = k0 + threadIdx.x;
        if (k < count && s < T) table[(size_t)s * P + (size_t)(idx0 + k * stride)] = tile[threadIdx.x][r];
    }
}

__global__ void scatter_idx_kernel(const double *__restrict__ rows, const long long *__restrict__ idx1, long long count,
                                   int T, long long P, double *__restrict__ table) {
    __shared__ double tile[32][33];
    const long long k0 = (long long)blockIdx.x * 32;
    const int s0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const long long k = k0 + r;
        const int s = s0 + threadIdx.x;
        if (k < count && s < T) tile[r][threadIdx.x] = rows[(size_t)k * T + s];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int s = s0 + r;
        const long long k = k0 + threadIdx.x;
        if (k < count && s < T) table[(size_t)s * P + (size_t)(idx1[k] - 1)] = tile[threadIdx.x][r];
    }
}

template <int DIM, typename OffT, bool NBRG>
cudaError_t set_smem_attr(int bytes) {
    static int configured = -1;
    if (bytes > configured) {
        cudaError_t e = cudaFuncSetAttribute(spr_traj_kernel<DIM, OffT, NBRG>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e != cudaSuccess) return e;
        if (!NBRG) {
            e = cudaFuncSetAttribute(spr_table_dump_kernel<DIM, OffT>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
            if (e != cudaSuccess) return e;
        }
        configured = bytes;
    }
    return cudaSuccess;
}

inline int align16(int x) { return (x + 15) & ~15; }

}  // namespace

SmemLayout make_layout(int N, int DIM, int cap, bool nbr_global) {
    (void)DIM;
    SmemLayout l;
    if (cap < 1024) cap = 1024;
    cap = (cap + 63) & ~63;
    l.cap = cap;
    l.wide_off = cap > 65535 ? 1 : 0;
    l.nbr_global = nbr_global ? 1 : 0;
    int o = 0;
    // cfill (u16 per cell, <= min(N,1024) cells, setup only) aliases nbr, or off when nbr is in global memory
    l.o_cfill = 0;
    if (!nbr_global) o += align16(cap * 2);
    l.o_off = o;
    if (nbr_global) l.o_cfill = o;
    o += align16((N + 1) * (l.wide_off ? 4 : 2));
    l.o_region = o;
    // state view: the A list (u16, from the front) and the complex list (u32 pairs, from the back) share
    // one array of 2N bytes: cntA + 2 cntC <= N always
    int s = o;
    l.o_sn = s;      s += align16(N * 4);
    l.o_listA = s;
    const int ac_words = (2 * N + 3) / 4;
    l.o_listC = s + 4 * (ac_words - 1);          // address of the LAST word
    s += align16(4 * ac_words);
    l.o_listB = s;   s += align16(N * 2);
    l.o_pos = s;     s += align16(N * 2);
    // setup view: packed lattice positions (8 B per particle) and the cell starts
    int u = o;
    l.o_spos = u;    u += align16(N * 8);
    l.o_cstart = u;  u += align16((1024 + 2) * 2);
    o = s > u ? s : u;
    l.o_rng = o;
    o += 32 * 16;
    l.total = o;
    return l;
}

int max_dynamic_smem(int device) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) != cudaSuccess) return 0;
    return v;
}

// (DIM, offset width, table placement) -> template instantiation
template <typename F>
cudaError_t dispatch(int DIM, int wide, int nbrg, F &&f) {
    using I3 = std::integral_constant<int, 3>;
    using I2 = std::integral_constant<int, 2>;
    using TT = std::true_type;
    using FF = std::false_type;
    if (DIM == 3) {
        if (wide) return nbrg ? f(I3{}, uint32_t{}, TT{}) : f(I3{}, uint32_t{}, FF{});
        return nbrg ? f(I3{}, uint16_t{}, TT{}) : f(I3{}, uint16_t{}, FF{});
    }
    if (wide) return nbrg ? f(I2{}, uint32_t{}, TT{}) : f(I2{}, uint32_t{}, FF{});
    return nbrg ? f(I2{}, uint16_t{}, TT{}) : f(I2{}, uint16_t{}, FF{});
}

cudaError_t launch_traj(const TrajArgs &a, int DIM, int nblocks, cudaStream_t st) {
    return dispatch(DIM, a.lay.wide_off, a.lay.nbr_global, [&](auto dim_c, auto off_t, auto g_c) -> cudaError_t {
        constexpr int D_ = decltype(dim_c)::value;
        using O_ = decltype(off_t);
        constexpr bool G_ = decltype(g_c)::value;
        cudaError_t e = set_smem_attr<D_, O_, G_>(a.lay.total);
        if (e != cudaSuccess) return e;
        spr_traj_kernel<D_, O_, G_><<<nblocks, 32, a.lay.total, st>>>(a);
        return cudaGetLastError();
    });
}

int traj_blocks_per_sm(const SmemLayout &lay, int DIM) {
    int nb = 0;
    cudaError_t e = dispatch(DIM, lay.wide_off, lay.nbr_global, [&](auto dim_c, auto off_t, auto g_c) -> cudaError_t {
        constexpr int D_ = decltype(dim_c)::value;
        using O_ = decltype(off_t);
        constexpr bool G_ = decltype(g_c)::value;
        cudaError_t e2 = set_smem_attr<D_, O_, G_>(lay.total);
        if (e2 != cudaSuccess) return e2;
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, spr_traj_kernel<D_, O_, G_>, 32, lay.total);
    });
    return e == cudaSuccess ? nb : 0;
}

cudaError_t launch_table_dump(const TrajArgs &a, int DIM, uint32_t replicate, int slot, uint32_t *d_off_out,
                              uint16_t *d_nbr_out, uint64_t *d_spos_out, long long *d_total_out,
                              cudaStream_t st) {
    return dispatch(DIM, a.lay.wide_off, 0, [&](auto dim_c, auto off_t, auto) -> cudaError_t {
        constexpr int D_ = decltype(dim_c)::value;
        using O_ = decltype(off_t);
        cudaError_t e = set_smem_attr<D_, O_, false>(a.lay.total);
        if (e != cudaSuccess) return e;
        spr_table_dump_kernel<D_, O_><<<1, 32, a.lay.total, st>>>(a, replicate, slot, d_off_out, d_nbr_out,
                                                                 d_spos_out, d_total_out);
        return cudaGetLastError();
    });
}

cudaError_t launch_fixed_graph(const double *d_locs, int N, int DIM, double L, double reach, uint32_t *d_deg_off,
                               uint16_t *d_nbr, long long nbr_cap, unsigned long long *d_total,
                               cudaStream_t st) {
    const double reach2 = reach * reach;
    const int nb = (N + 127) / 128;
    if (d_nbr == nullptr) {
        fixed_graph_count_kernel<<<nb, 128, 0, st>>>(d_locs, N, DIM, L, reach2, d_deg_off);
        fixed_graph_scan_kernel<<<1, 1024, 0, st>>>(d_deg_off, N, d_total);
    } else {
        fixed_graph_fill_kernel<<<nb, 128, 0, st>>>(d_locs, N, DIM, L, reach2, d_deg_off, d_nbr, nbr_cap);
    }
    return cudaGetLastError();
}

cudaError_t launch_reduce_fixed(const uint16_t *curve, int nslots, int repstride, int nrep, int T, long long *sum,
                                unsigned long long *sumsq, cudaStream_t st) {
    if (nslots <= 0) return cudaSuccess;
    reduce_fixed_kernel<<<nslots, 128, 0, st>>>(curve, repstride, nrep, T, sum, sumsq);
    return cudaGetLastError();
}

cudaError_t launch_variance_scan(const uint16_t *curve, const int32_t *active, int nactive, int repstride, int T,
                                 int navail, double ssetol, int minsims, int maxsims, long long *sum,
                                 unsigned long long *sumsq, int32_t *nscanned, int32_t *nstar,
                                 cudaStream_t st) {
    if (nactive <= 0) return cudaSuccess;
    variance_scan_kernel<<<nactive, 128, 0, st>>>(curve, active, repstride, T, navail, ssetol, minsims, maxsims,
                                                  sum, sumsq, nscanned, nstar);
    return cudaGetLastError();
}

cudaError_t launch_means(const long long *sum, const int32_t *nused, int nslots, int T, int N, double *rows,
                         cudaStream_t st) {
    if (nslots <= 0) return cudaSuccess;
    means_kernel<<<nslots, 128, 0, st>>>(sum, nused, T, N, rows);
    return cudaGetLastError();
}

cudaError_t launch_scatter(const double *rows, long long idx0, long long stride, long long count, int T,
                           long long P, double *table, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    dim3 grid((unsigned)((count + 31) / 32), (unsigned)((T + 31) / 32));
    dim3 block(32, 8);
    scatter_kernel<<<grid, block, 0, st>>>(rows, idx0, stride, count, T, P, table);
    return cudaGetLastError();
}

cudaError_t launch_scatter_idx(const double *rows, const long long *idx1, long long count, int T, long long P,
                               double *table, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    dim3 grid((unsigned)((count + 31) / 32), (unsigned)((T + 31) / 32));
    dim3 block(32, 8);
    scatter_idx_kernel<<<grid, block, 0, st>>>(rows, idx1, count, T, P, table);
    return cudaGetLastError();
}

}  // namespace spr
