// spr_kernels.cuh -- argument blocks and launch wrappers shared by spr_kernels.cu and spr_api.cu
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

namespace spr {

struct PointDev {          // one parameter point as the kernels see it
    double kon_eff, koff, konb, reach;
};

// ------------------------------------------------------------------------------------------------
// Work items.  A *ticket* is one trajectory: tk = (position in `order`) * nrep + (replicate - rep0).
// A launch processes local tickets 0..ntickets-1; local ticket q is tk = ticket_list ? ticket_list[q]
// : tk0 + q.  Both kernels are persistent and pull local tickets from a global counter.
//
// Neighbour-table record of one trajectory, written by K2 into the HBM arena and read by K1:
//     [ off u32[N+1] | pad to 16 B | words u32[off[N]] | pad to 16 B ]
// off[i] = first WORD of the row of particle i; a row holds its neighbours packed entries_per_word(N) to a
// 32-bit word (10-bit entries, 3 per word, when N <= 1023; else 16-bit, 2 per word), the last word padded
// with all-ones entries.
// `desc[q]` holds the device address of the record of local ticket q, or 0 when the arena was full
// (the ticket is then listed in `defer_list` and neither kernel simulates it in this launch).
// ------------------------------------------------------------------------------------------------

// slots of the counter block (unsigned long long each)
enum { CTR_K2_TICKET = 0, CTR_CURSOR = 1, CTR_MAX_ENTRIES = 2, CTR_MAX_DEGREE = 3, CTR_NDEFER = 4, CTR_K1_TICKET = 5,
       CTR_COUNT = 8 };

struct TableArgs {                      // K2: spr_table_kernel
    int N;
    double L;
    const PointDev *pts;                // [nslots]
    const uint32_t *pidx;               // [nslots] RNG parameter index
    const int32_t *order;               // work order -> slot (expensive points first)
    long long tk0, ntickets;
    const long long *ticket_list;       // optional explicit tickets
    int rep0, nrep;
    uint32_t seed_lo, seed_hi;
    unsigned long long *ctr;            // counter block (CTR_*)
    unsigned char *arena;
    unsigned long long arena_bytes;
    unsigned long long *desc;           // [ntickets]
    long long *defer_list;              // [ntickets]
    uint16_t *stage;                    // [nblocks][N][stage_cap] staging rows of the distance pass
    int stage_cap;                      // entries per staging row (>= the stride any trajectory of the launch uses)
    int epw, eb;                        // packed rows: entries per word, bits per entry (entries_per_word / entry_bits)
    uint64_t *spos_out;                 // test hook: packed sorted lattice positions of local ticket 0, or nullptr
};

enum { OBS_TOTAL_BOUND = 0, OBS_TOTAL_A = 1, OBS_RAW_BC = 2 };

// Byte offsets inside one warp's shared-memory slice of K1 (one warp = one trajectory).
struct TrajLayout {
    int o_off;      // OffT[N+1]   first word of every packed neighbour row
    int o_sn;       // SnT[N]      state | nA << 2 | nB << (2 + CB)
    int o_ab;       // u16[N]      state-A particles from the front, state-B from the back, one member of every
                    //             complex in the gap between them (A + B + 2C = N)
    int o_pos;      // u16[N]      A/B: index in its list; C: index of the partner
    int slice;      // bytes per warp
    int wide_sn;    // SnT = uint32_t (15-bit neighbour counts) instead of uint16_t (7-bit)
    int wide_off;   // OffT = uint32_t (more than 65535 words of packed rows)
    int eb;         // bits per packed table entry (10 or 16)
    int tmem;       // row offsets in tensor memory instead of o_off (6 B of shared memory per particle, 32 warps per SM)
    int o_tbl;      // u32[tbl_words]  the packed rows themselves, when tbl_words > 0 (underfilled launches, one warp per CTA)
    int tbl_words;
    int wpc;        // warps (trajectories) per CTA
    int cta_per_sm;
};

constexpr int K1_WPC = 7;          // warps per CTA at full occupancy: 4 CTAs x 7 warps = 28 trajectories per SM
constexpr int K1_MAX_WARPS = 28;   // 72 registers per thread
constexpr int K1_WPC_TM = 8;       // tensor-memory variant: 4 CTAs x 8 warps = 32 trajectories per SM, 64 registers
constexpr int K1_TMEM_COLS = 64;   // tensor-memory columns per CTA (two blocks of 32 per lane quadrant)

// packed table entries: 10 bits (3 per word) while every particle index and the EMPTY marker fit, else 16 (2 per word)
// (SPRB200_ENTRY_BITS=16 forces the 16-bit rows: experiment hook)
inline bool force_wide_entries() {
    static const bool wide = [] { const char *e = std::getenv("SPRB200_ENTRY_BITS"); return e && std::atoi(e) == 16; }();
    return wide;
}
inline int entry_bits(int N) { return (N <= 1023 && !force_wide_entries()) ? 10 : 16; }
inline int entries_per_word(int N) { return entry_bits(N) == 10 ? 3 : 2; }

struct TrajArgs {                       // K1: spr_traj_kernel
    int N, T;
    double tstop, toff;
    const double *tsave;                // [T] device
    const PointDev *pts;
    const uint32_t *pidx;
    const int32_t *order;
    long long tk0, ntickets;
    const long long *ticket_list;
    int rep0, nrep, repstride;          // curve row stride per slot
    int rep_rowbase;                    // curve row of replicate r is r - rep_rowbase
    uint32_t seed_lo, seed_hi;
    unsigned long long *ctr;
    const unsigned long long *desc;     // [ntickets] record addresses (resampled positions) ...
    const unsigned long long *slot_rec; // ... or one record per slot (fixed positions), else nullptr
    uint16_t *curve;                    // [nslots][repstride][T]
    uint16_t *curve2;                   // second plane (OBS_RAW_BC: C counts) or nullptr
    unsigned long long *nevents;        // [nslots]
    int observable;
    TrajLayout lay;
};

struct FitArgs {                        // K5: fit_loss_kernel
    const double *table;                // FirstMoment [T][P] on the device
    long long P;
    int n[4], T;
    double lo[4], hi[4];                // axis ranges (log10 kon', log10 koff, log10 konb, reach)
    double t_first, t_last;             // tsave[1], tsave[end]
    int nconc;
    long long nsamples;
    const int *npts;                    // [nconc]
    const double *times, *values;       // [nsamples] concatenated per concentration
    const double *abcs;                 // [nconc] antibody concentrations
    const double *optpars;              // [npop][5]
    double *loss;                       // [npop]
};
cudaError_t launch_fit_loss(const FitArgs &a, long long npop, cudaStream_t st);

inline size_t align16z(size_t x) { return (x + 15) & ~(size_t)15; }
inline size_t record_bytes(int N, unsigned long long words) {
    return align16z(sizeof(uint32_t) * (size_t)(N + 1)) + align16z(sizeof(uint32_t) * (size_t)words);
}

// K2
int table_smem_bytes(int N);
int table_ctas_per_sm(int N, int DIM);
int table_stage_cap(int N, int DIM, double L, double reach);   // staging stride for one reach value
cudaError_t launch_table(const TableArgs &a, int DIM, int nblocks, cudaStream_t st);
constexpr int K2_THREADS = 256;

// K1: picks the warps-per-CTA that keeps the most trajectories resident per SM; false if one slice does not fit
bool make_traj_layout(int N, bool wide_sn, bool wide_off, int smem_optin, int smem_per_sm, TrajLayout *out, bool tmem = false);
cudaError_t launch_traj(const TrajArgs &a, int nblocks, cudaStream_t st);

// fixed-position neighbour table (K2b): FP64, the reference's periodic_dist_sq.
// d_words == nullptr: degrees -> exclusive word offsets of the packed rows in d_off[0..N], total words in *d_total;
// else: fill the packed rows using the offsets.
cudaError_t launch_fixed_graph(const double *d_locs, int N, int DIM, double L, double reach,
                               uint32_t *d_off /* N+1 */, uint32_t *d_words,
                               unsigned long long *d_total, unsigned int *d_maxdeg, cudaStream_t st);

// K3: replicate reduction / sequential-stop scan
cudaError_t launch_reduce_fixed(const uint16_t *curve, int nslots, int repstride, int nrep, int T,
                                long long *sum, unsigned long long *sumsq, cudaStream_t st);
cudaError_t launch_variance_scan(const uint16_t *curve, const int32_t *active, int nactive,
                                 int repstride, int T, int navail, double ssetol, int minsims,
                                 int maxsims, long long *sum, unsigned long long *sumsq,
                                 int32_t *nscanned, int32_t *nstar, cudaStream_t st);
// K4: means in row order [slot][T] or scattered into FirstMoment order [T][P]
cudaError_t launch_means(const long long *sum, const int32_t *nused, int nslots, int T, int N,
                         double *rows, cudaStream_t st);
cudaError_t launch_scatter(const double *rows, long long idx0, long long stride, long long count,
                           int T, long long P, double *table, cudaStream_t st);
// rows [count][T] -> table[s*P + idx1[k]-1] for an explicit (1-based) index list
cudaError_t launch_scatter_idx(const double *rows, const long long *idx1, long long count, int T, long long P,
                               double *table, cudaStream_t st);
}  // namespace spr
