/*
 * sprb200.h -- C ABI of libsprb200.so: the B200 (sm_100a) replacement for the data-parallel hot
 * path of SPRFittingPaper2023.jl: run_spr_sim! (the ensemble of spatial stochastic particle
 * simulations) swept over the (kon, koff, konb, reach) grid that builds the surrogate.
 *
 * Every entry point is plain C: POD structs, raw pointers and sizes, integer error codes.  No
 * C++/torch types cross the boundary, nothing throws, no pointer is retained after a call returns.
 * All paths below are relative to the reference tree (/root/reference).
 *
 * Reference interface replaced                               | entry point here
 * -----------------------------------------------------------+------------------------------------
 * run_spr_sim!(outputter, biopars, simpars, terminator)      | sprb200_simulate (reduced moments)
 *   src/forward_simulator.jl:105-366                         | sprb200_simulate_raw (copynumbers)
 * TotalBoundOutputter / TotalAOutputter accumulation         | SprOut.sum / sumsq / n_used
 *   src/callbacks.jl:23-33, 105-114; means/vars/sems :54-80  |   (x = B+C or A, integer moments)
 * SimNumberTerminator / VarianceTerminator                   | SprRun.term_kind, nsims, ssetol,
 *   src/callbacks.jl:152-223                                 |   minsims, maxsims
 * setup_spr_sim! (neighbour tables) :37-89                   | sprb200_neighbor_table (test hook)
 * build_surrogate_slice(meta, idxstart, idxend)              | sprb200_build_slice
 *   src/surrogate.jl:273-300                                 |
 * build_surrogate_serial(size, surpars, simpars)             | sprb200_build_table
 *   src/surrogate.jl:203-242                                 |
 * merge_surrogate_slices (column scatter) :302-318           | sprb200_merge_slice
 * SimParams(...) L from antigen concentration                | sprb200_domain_length
 *   src/parameters.jl:124-150, src/utils.jl:10               |
 * surrogate_sprdata_error(optpars, pars)                      | sprb200_surrogate_load,
 *   src/fitting.jl:38-91 (5-D interpolation + l2 loss)       | sprb200_fit_loss (whole population)
 * the index-range split over independent processes         | sprb200_build_table_multi (all GPUs
 *   examples/cluster_scripts/queue_job.sh:7-17 +             |   of one process), sprb200_comm_* +
 *   merge_surrogate_slices src/surrogate.jl:302-332          |   sprb200_build_table_dist (one
 *                                                            |   process per GPU, NCCL gather),
 *                                                            | sprb200_deal_indices; device-resident
 *                                                            | slabs: sprb200_build_slice_device,
 *                                                            | sprb200_scatter_table_device
 */
#ifndef SPRB200_H
#define SPRB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPRB200_ABI_VERSION 2

/* error codes (all negative); 0 = success */
enum {
    SPRB200_OK = 0,
    SPRB200_ERR_BAD_ARG = -1,      /* NULL/invalid argument, unsorted tsave, N out of range ... */
    SPRB200_ERR_NO_DEVICE = -2,    /* no CUDA device / not an sm_100 part / bad ordinal */
    SPRB200_ERR_CUDA = -3,         /* CUDA runtime error, text in sprb200_last_error */
    SPRB200_ERR_TOO_LARGE = -4,    /* per-trajectory state does not fit in shared memory */
    SPRB200_ERR_ALLOC = -5,        /* host or device allocation failed */
    SPRB200_ERR_INTERNAL = -6,
    SPRB200_ERR_NCCL = -7          /* libnccl.so.2 missing or an NCCL call failed (multi-GPU entry points only) */
};

typedef struct sprb200_ctx *sprb200_handle;

/* One parameter point.  kon enters the dynamics only as kon*antibodyconcen
 * (forward_simulator.jl:132); CP only scales the outputter value (callbacks.jl:28) and is applied
 * by the caller; antigenconcen acts only through SprSim.L. */
typedef struct {
    double kon_eff; /* kon * antibodyconcen, s^-1 */
    double koff;    /* s^-1 */
    double konb;    /* s^-1 */
    double reach;   /* nm */
} SprPoint;

/* Mirror of SimParams (parameters.jl:92-109). */
typedef struct {
    int32_t N;                 /* particles, 1..65535 */
    int32_t DIM;               /* 2 or 3 */
    double L;                  /* periodic box edge */
    double tstop;
    double tstop_AtoB;         /* may be +Inf */
    int32_t T;                 /* number of save times */
    const double *tsave;       /* host, T ascending values */
    int32_t resample_initlocs; /* !=0: fresh uniform positions every replicate (:42-47,138-142) */
    const double *initlocs;    /* host, N*DIM doubles, DIM contiguous per particle (memory of
                                  Vector{SVector{DIM,Float64}}); required iff !resample_initlocs */
} SprSim;

enum { SPRB200_TERM_FIXED = 0, SPRB200_TERM_VARIANCE = 1 };
enum { SPRB200_OBS_TOTAL_BOUND = 0, SPRB200_OBS_TOTAL_A = 1 };

typedef struct {
    int32_t term_kind;         /* FIXED: exactly nsims replicates (SimNumberTerminator);
                                  VARIANCE: VarianceTerminator(ssetol, minsims, maxsims) with the
                                  reference's sequential stopping rule (callbacks.jl:204-217) */
    int32_t nsims;
    double ssetol;
    int32_t minsims;
    int32_t maxsims;
    int32_t observable;        /* SPRB200_OBS_*: x = B+C (callbacks.jl:28) or x = A (:111) */
    int32_t replicate_base;    /* index of the first replicate (RNG counter); 0 unless a caller
                                  continues an earlier run of the same parameter index */
    uint64_t seed;             /* Philox key */
    uint64_t param_index_base; /* RNG parameter index of pts[0]; point k uses base+k.  Streams are
                                  keyed by (seed, parameter index, replicate): results do not depend
                                  on batching, GPU count or scheduling. */
} SprRun;

/* Reduced per-point output; every pointer is a caller-allocated HOST buffer or NULL (skipped). */
typedef struct {
    int64_t *sum;       /* [npts][T]  sum over used replicates of x */
    uint64_t *sumsq;    /* [npts][T]  sum of x^2 */
    int32_t *n_used;    /* [npts]     replicates entering the sums (the terminator's nobs) */
    uint64_t *n_events; /* [npts]     SSA events simulated for the point (all replicates run) */
    double *mean;       /* [npts][T]  sum / (n_used * N): the reference's means(o) for CP = 1 */
} SprOut;

/* Mirror of SurrogateParams ranges + surrogate_size[1:4] (surrogate.jl:9-24, 208-211):
 * axis k takes n[k] values range(lo[k], hi[k], length=n[k]); axes 0..2 are log10 of
 * kon*[Ab], koff, konb; axis 3 is reach. */
typedef struct {
    double lo[4];
    double hi[4];
    int32_t n[4];
} SprGrid;

/* ---- lifecycle -------------------------------------------------------------------------- */
int sprb200_abi_version(void);
/* Creates a context on CUDA device `device_ordinal` (its own non-default streams and device
 * arenas).  One handle per GPU per host thread; a handle is not thread-safe. */
int sprb200_create(int device_ordinal, sprb200_handle *out);
void sprb200_destroy(sprb200_handle h);
/* Text of the last error on this handle (or of the last failed create when h == NULL). */
const char *sprb200_last_error(sprb200_handle h);
/* Counters of the last simulate/build call: trajectories run, events simulated, kernel launches,
 * device milliseconds between the first and last kernel (CUDA events on the handle's stream). */
int sprb200_last_stats(sprb200_handle h, uint64_t *trajectories, uint64_t *events,
                       uint64_t *kernel_launches, double *device_ms);

/* Device milliseconds of the last call spent in the neighbour-table kernel (K2) and in the trajectory
 * kernel (K1), summed over their launches (CUDA events on the handle's stream). */
int sprb200_last_kernel_ms(sprb200_handle h, double *table_ms, double *traj_ms);

/* Shared-memory plan of the trajectory kernel on a B200 for N particles (no device needed):
 * out[0] = bytes of shared memory per trajectory, out[1] = trajectories (warps) per CTA, out[2] = CTAs per
 * SM, out[3] = bytes of shared memory of one table-build CTA.  wide_state / wide_offsets select the
 * 32-bit state words (a degree above 127) / offsets (more than 65,535 table entries).
 * SPRB200_ERR_TOO_LARGE when one trajectory does not fit.  (This is the shared-memory plan; launches that fill the
 * GPU with N <= 1023 keep the row offsets in tensor memory instead: 6 B of shared memory per particle, 8 warps per CTA,
 * 4 CTAs per SM.  SPRB200_TMEM=0 in the environment switches that variant off.) */
int sprb200_traj_layout(int32_t N, int32_t wide_state, int32_t wide_offsets, int32_t out[4]);

/* Relative cost (expected SSA events x neighbour work of one replicate) of each point, the figure the library's own
 * multi-GPU deal (sprb200_deal_indices) balances on grids of up to 400,000 points: the events come from the mean-field
 * kinetics of the four reaction channels (a two-variable ODE per point, about 3 us; within a few per cent of the
 * executed events over the C2 ranges, tests/tools/cost_model_check.py).  Device-free; sim->tsave may be NULL here. */
int sprb200_estimate_cost(const SprSim *sim, const SprPoint *pts, int64_t npts, double *cost_out);

/* ---- helpers --------------------------------------------------------------------------- */
/* L = (N / agc)^(1/DIM), agc = 6.02214076e-7*antigenconcen if convert_units (parameters.jl:138-139) */
double sprb200_domain_length(int32_t N, int32_t DIM, double antigenconcen, int32_t convert_units);
/* value i0 (0-based) of range(lo, hi, length=n) */
double sprb200_grid_value(double lo, double hi, int32_t n, int32_t i0);
/* parameter point of 1-based linear index idx (kon fastest ... reach slowest, surrogate.jl:284) */
int sprb200_grid_point(const SprGrid *grid, int64_t idx, SprPoint *out);

/* ---- forward simulation (run_spr_sim!) -------------------------------------------------- */
/* Simulates npts parameter points x replicates; blocking; host buffers in, host buffers out. */
int sprb200_simulate(sprb200_handle h, const SprSim *sim, const SprPoint *pts, int64_t npts,
                     const SprRun *run, const SprOut *out);

/* Raw copynumbers for arbitrary outputter/terminator replay on the host: FIXED nsims replicates,
 * B_out/C_out are uint16 [npts][nsims][T] (A = N - B - 2C; rows of copynumbers,
 * forward_simulator.jl:189-191).  Either pointer may be NULL. */
int sprb200_simulate_raw(sprb200_handle h, const SprSim *sim, const SprPoint *pts, int64_t npts,
                         const SprRun *run, uint16_t *B_out, uint16_t *C_out, uint64_t *n_events);

/* Test hook for the neighbour tables: builds the table of one trajectory exactly as the kernel
 * does (resampled positions of (seed, param_index, replicate) when sim->resample_initlocs, else
 * sim->initlocs) and returns it in CSR form.  positions_out: N*DIM doubles (lattice positions in
 * [0,L) in the kernel's sorted numbering) or NULL; off_out: N+1; nbr_out: up to nbr_cap entries.
 * Returns the number of directed entries (2 x pairs) or a negative error. */
int64_t sprb200_neighbor_table(sprb200_handle h, const SprSim *sim, double reach, uint64_t seed,
                               uint64_t param_index, uint32_t replicate, double *positions_out,
                               int32_t *off_out, int32_t *nbr_out, int64_t nbr_cap);

/* ---- surrogate build -------------------------------------------------------------------- */
/* build_surrogate_slice: means of 1-based linear indices idxstart..idxend (inclusive), CP = 1,
 * [Ab] = 1; out_TxNidx is T x nidx column-major (time fastest), i.e. out[n*T + s].
 * n_used / n_events: [nidx] or NULL.  run->param_index_base is ignored: point idx always uses RNG
 * parameter index idx-1, so any slicing reproduces the full table bit for bit. */
int sprb200_build_slice(sprb200_handle h, const SprGrid *grid, const SprSim *sim, const SprRun *run,
                        int64_t idxstart, int64_t idxend, double *out_TxNidx, int32_t *n_used,
                        uint64_t *n_events);
/* build_surrogate_serial: the full "FirstMoment" array (n1,n2,n3,n4,T) column-major (kon fastest,
 * time slowest), out_table[(s*P) + idx-1], P = n1*n2*n3*n4. */
int sprb200_build_table(sprb200_handle h, const SprGrid *grid, const SprSim *sim, const SprRun *run,
                        double *out_table, int32_t *n_used, uint64_t *n_events);
/* merge_surrogate_slices: scatters a T x nidx slice into the (.., T) table (host, no GPU work). */
int sprb200_merge_slice(const SprGrid *grid, int32_t T, const double *slice_TxNidx, int64_t idxstart,
                        int64_t idxend, double *table);

/* ---- device-resident variants (one rank per GPU; NCCL gather done by the caller) ----------- */
/* Same as sprb200_build_slice for an arbitrary strided index set idx = idxstart + k*stride,
 * k = 0..count-1 (1-based idx).  d_out_NidxxT is a DEVICE pointer to count*T doubles, row k = mean
 * curve of index k ([count][T], time fastest); d_n_used (int32[count]) and d_n_events
 * (uint64[count]) are device pointers or NULL.  Asynchronous work is complete on return. */
int sprb200_build_slice_device(sprb200_handle h, const SprGrid *grid, const SprSim *sim,
                               const SprRun *run, int64_t idxstart, int64_t stride, int64_t count,
                               double *d_out_NidxxT, int32_t *d_n_used, uint64_t *d_n_events);
/* Device transpose/scatter of gathered slabs into FirstMoment order: for k in 0..count-1,
 * d_table[s*P + (idxstart-1 + k*stride)] = d_rows[k*T + s]. */
int sprb200_scatter_table_device(sprb200_handle h, const double *d_rows, int64_t idxstart,
                                 int64_t stride, int64_t count, int32_t T, int64_t P, double *d_table);

/* Arbitrary index sets (block-cyclic sharding over ranks): idx[k] are 1-based linear grid indices in
 * host memory; row k of the [count][T] output is the mean curve of idx[k]. */
int sprb200_build_indices_device(sprb200_handle h, const SprGrid *grid, const SprSim *sim,
                                 const SprRun *run, const int64_t *idx, int64_t count,
                                 double *d_out_NidxxT, int32_t *d_n_used, uint64_t *d_n_events);
/* Host-buffer variant of the above (out_NidxxT, n_used, n_events are host pointers; the latter two
 * may be NULL). */
int sprb200_build_indices(sprb200_handle h, const SprGrid *grid, const SprSim *sim, const SprRun *run,
                          const int64_t *idx, int64_t count, double *out_NidxxT, int32_t *n_used,
                          uint64_t *n_events);
/* d_table[s*P + idx[k]-1] = d_rows[k*T + s] for a host index list (after the NCCL gather). */
int sprb200_scatter_rows_device(sprb200_handle h, const double *d_rows, const int64_t *idx,
                                int64_t count, int32_t T, int64_t P, double *d_table);

/* ---- multi-GPU builds (the reference's "distributed backend" is the split of the linear index range over
 * independent processes, examples/cluster_scripts/queue_job.sh:7-17, merged afterwards by
 * merge_surrogate_slices, surrogate.jl:302-332).  Every (point, replicate) trajectory is independent and
 * its random stream is keyed by (seed, linear index - 1, replicate): the table is byte-identical for
 * any number of GPUs.  The grid points are dealt to the GPUs in snake order of the a-priori cost estimate
 * (equal cost mix and equal counts per GPU). ------------------------------------------------------- */
/* The 1-based linear indices `rank` of `nranks` simulates (ascending).  Device-free.  `run` (may be NULL = FIXED)
 * only matters for the VarianceTerminator, whose expected replicate count per point (15 ... 250 with the defaults)
 * enters the cost.  Writes at most `cap` indices to idx_out (may be NULL) and returns the count, or a negative
 * error. */
int64_t sprb200_deal_indices(const SprGrid *grid, const SprSim *sim, const SprRun *run, int32_t nranks,
                             int32_t rank, int64_t *idx_out, int64_t cap);
/* build_surrogate_serial on ndev GPUs of THIS process: one host thread and one context per device (created
 * on first use, kept until sprb200_multi_release), every device simulates its share, the slabs are
 * gathered on devices[0] with peer copies (NVLink), permuted there into FirstMoment order and copied
 * to out_table ((n1,n2,n3,n4,T) column-major, host).  n_used / n_events: [P] host or NULL.
 * Error text: sprb200_last_error(NULL). */
int sprb200_build_table_multi(const int32_t *devices, int32_t ndev, const SprGrid *grid, const SprSim *sim,
                              const SprRun *run, double *out_table, int32_t *n_used, uint64_t *n_events);
/* Per-device counters of the last sprb200_build_table_multi call: events simulated, device milliseconds of the
 * whole share and of the trajectory kernel alone (each [ndev] or NULL). */
int sprb200_multi_stats(const int32_t *devices, int32_t ndev, uint64_t *events, double *device_ms, double *traj_ms);
/* Destroys the contexts sprb200_build_table_multi created. */
void sprb200_multi_release(void);

/* One process per GPU (MPI / torchrun / one Julia worker per GPU).  NCCL is bound at run time (dlopen of
 * libnccl.so.2; SPRB200_ERR_NCCL if absent) and used for exactly one exchange: the all-gather of the
 * mean-curve slabs.  Rank 0 calls sprb200_comm_unique_id and hands the 128 bytes to every rank by any
 * means (MPI_Bcast, a file, torch.distributed); every rank then calls sprb200_comm_init (collective). */
int sprb200_comm_unique_id(unsigned char id_out[128]);
int sprb200_comm_init(sprb200_handle h, const unsigned char id[128], int32_t nranks, int32_t rank);
int sprb200_comm_destroy(sprb200_handle h);
/* build_surrogate_serial across the communicator (collective: every rank calls it with the same arguments).
 * Each rank simulates its share (sprb200_deal_indices), one ncclAllGather assembles the slabs on every
 * rank and the table is permuted on the device into FirstMoment order.  out_table (host, P*T doubles) and
 * d_table (device, P*T doubles) may each be NULL; n_used / n_events: [P] host or NULL (the output pointers may
 * differ between ranks: the collectives issued do not depend on them).  With nranks == 1
 * no communicator is needed.  sprb200_last_stats / sprb200_last_kernel_ms report this rank's share. */
int sprb200_build_table_dist(sprb200_handle h, const SprGrid *grid, const SprSim *sim, const SprRun *run,
                             double *out_table, double *d_table, int32_t *n_used, uint64_t *n_events);

/* ---- the consumer of the table: surrogate_sprdata_error (src/fitting.jl:38-91) ------------------- */
/* Mirror of AlignedData (src/spr_data.jl:9-18) in concatenated form. */
typedef struct {
    int32_t nconc;                 /* number of antibody concentrations (curves) */
    const int32_t *npts;           /* [nconc] samples per curve */
    const double *times;           /* [sum npts] times[i] of every curve, concatenated */
    const double *values;          /* [sum npts] refdata[i], concatenated */
    const double *antibodyconcens; /* [nconc]; the first one is the reference concentration (:66) */
} SprAligned;
/* Keeps a surrogate resident on the device: `table` is the FirstMoment array (n1,n2,n3,n4,T) column-major
 * (what sprb200_build_table writes) on the host, or on the device when table_on_device != 0; tsave must be
 * uniform from t_first to t_last (surrogate.jl:96-108 assumes the same). */
int sprb200_surrogate_load(sprb200_handle h, const SprGrid *grid, int32_t T, double t_first, double t_last,
                           const double *table, int32_t table_on_device);
/* loss[k] = sum over curves j and samples i of (10^p5 * itp(q1_j, q2, q3, q4, s_i) - refdata[j][i])^2 for
 * candidate k = optpars[5k .. 5k+4] = [logkon, logkoff, logkonb, reach, logCP]; itp is the 5-D
 * multilinear interpolation of the loaded table at fractional 1-based indices (surrogate.jl:66-70),
 * logkon is shifted by log10(abc_j / abc_1) per curve.  A candidate or sample outside the table gives
 * +Inf (Interpolations.jl throws BoundsError there).  Host pointers; blocking. */
int sprb200_fit_loss(sprb200_handle h, const SprAligned *data, const double *optpars, int64_t npop, double *loss);

#ifdef __cplusplus
}
#endif
#endif /* SPRB200_H */
