// spr_kernels.cuh -- argument blocks and launch wrappers shared by spr_kernels.cu and spr_api.cu
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace spr {

struct PointDev {          // one parameter point as the kernel sees it
    double kon_eff, koff, konb, reach;
};

// Byte offsets of the per-trajectory shared-memory arrays (one warp = one CTA = one trajectory).
// [nbr | off | region | rng]; `region` is time-shared: during neighbour-table construction it holds
// the cell-sorted lattice positions and the cell starts, afterwards the particle/list state.
struct SmemLayout {
    int o_off;     // OffT[N+1]   CSR offsets
    int o_region;
    int o_sn;      // u32[N]      state | nA<<2 | nB<<17
    int o_listA;   // u16[N]      particles in state A
    int o_listB;   // u16[N]      particles in state B
    int o_pos;     // u16[N]      index of a particle in listA/listB
    int o_listC;   // u32         LAST word of the array shared with listA; complexes (a | b<<16) grow down
    int o_spos;    // u64[N]      (setup) packed 21-bit lattice positions, cell-sorted
    int o_cstart;  // u16[ncell+1](setup) first sorted index of every cell
    int o_rng;     // 32 x 16 B   block of pre-generated event randoms
    int total;     // bytes of dynamic shared memory
    int cap;       // nbr capacity (directed entries, u16 each)
    int wide_off;  // OffT = uint32_t when cap > 65535
    int o_cfill;   // u16[ncell]  (setup) aliases nbr (table in shared memory) or off (table in global memory)
    int nbr_global;// neighbour table lives in a per-CTA slice of TrajArgs::nbr_scratch (L2-resident)
                   // instead of shared memory: more trajectories per SM for wide tables
};

enum { OBS_TOTAL_BOUND = 0, OBS_TOTAL_A = 1, OBS_RAW_BC = 2 };

struct TrajArgs {
    int N, T;
    double L, tstop, toff;
    const double *tsave;            // [T] device
    const PointDev *pts;            // [nslots]
    const uint32_t *pidx;           // [nslots] RNG parameter index
    const int32_t *order;           // [norder] work order -> slot (expensive points first)
    long long ntickets;             // norder * nrep (or length of ticket_list)
    const long long *ticket_list;   // optional explicit tickets (re-run after capacity overflow)
    int rep0, nrep, repstride;      // replicates rep0 .. rep0+nrep-1; curve row stride per slot
    int rep_rowbase;                // curve row of replicate r is r - rep_rowbase
    uint32_t seed_lo, seed_hi;
    unsigned long long *ticket_ctr; // global work counter
    uint16_t *curve;                // [nslots][repstride][T]
    uint16_t *curve2;               // second plane (OBS_RAW_BC: C counts) or nullptr
    unsigned long long *nevents;    // [nslots]
    int observable;
    long long *ovf_list;            // tickets whose neighbour table exceeded `cap`
    unsigned int *ovf_count;
    unsigned int ovf_cap;
    uint16_t *nbr_scratch;          // [nblocks][lay.cap] when lay.nbr_global
    const uint32_t *g_off;          // fixed-position mode: shared CSR (else nullptr)
    const uint16_t *g_nbr;
    SmemLayout lay;
};

SmemLayout make_layout(int N, int DIM, int cap, bool nbr_global = false);
int max_dynamic_smem(int device);
// launches the persistent trajectory kernel: `nblocks` single-warp CTAs
cudaError_t launch_traj(const TrajArgs &a, int DIM, int nblocks, cudaStream_t st);
int traj_blocks_per_sm(const SmemLayout &lay, int DIM);

// fixed-position neighbour table (K2b): FP64, the reference's periodic_dist_sq
cudaError_t launch_fixed_graph(const double *d_locs, int N, int DIM, double L, double reach,
                               uint32_t *d_deg_off /* N+1 */, uint16_t *d_nbr, long long nbr_cap,
                               unsigned long long *d_total, cudaStream_t st);

// one-trajectory neighbour-table dump (test hook)
cudaError_t launch_table_dump(const TrajArgs &a, int DIM, uint32_t replicate, int slot,
                              uint32_t *d_off_out, uint16_t *d_nbr_out, uint64_t *d_spos_out,
                              long long *d_total_out, cudaStream_t st);

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
