/*
 * oracle/spr_ref.c -- CPU restatement of the reference's forward simulator and sweep.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (sprfittingpaper2023.jl_b200/)
 * may include, link or call this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, and only as the checker or as the
 * timed CPU baseline ("port").
 *
 * What it restates (all paths relative to /root/reference, Julia, cannot run in this image):
 *   src/forward_simulator.jl:1-8      periodic_dist_sq            -> ref_periodic_dist_sq
 *   src/forward_simulator.jl:37-89    setup_spr_sim!              -> ref_setup
 *   src/forward_simulator.jl:105-366  run_spr_sim!                -> spr_ref_run
 *   src/callbacks.jl:12-80            TotalBoundOutputter         -> Welford "Variance" per bin
 *   src/callbacks.jl:91-137           TotalAOutputter             -> observable = 1
 *   src/callbacks.jl:152-223          SimNumber/VarianceTerminator-> term_kind 0 / 1
 *   src/surrogate.jl:203-300          build_surrogate_serial/slice-> spr_ref_build_slice
 *   src/parameters.jl:4,124-150, src/utils.jl:10   L from antigen concentration -> spr_ref_default_L
 *
 * Same algorithm as the reference: next-reaction SSA on an indexed binary min-heap of ABSOLUTE
 * firing times (DataStructures.MutableBinaryHeap restated as heap[]/hpos[]/key[]), O(N^2) pair
 * search every replicate when resample_initlocs, identical save-slot semantics, identical
 * outputter arithmetic (OnlineStats.Variance: mu/sigma2 smoothing with weight 1/n, Bessel on
 * read-out) and terminator rule.
 *
 * Parity status: the reference never seeds its RNG (no golden trajectories exist), so this
 * oracle is pinned STATISTICALLY: against the reference's recorded mean curves
 * (tests/golden/fit_curves_golden.json, 2x586 means of 250 reference replicates), the exact
 * N=2 five-state CTMC and the monovalent closed form (tests/test_oracle_*.py).  It uses its own
 * generator (xoshiro256++), independent of the product's Philox streams, so GPU-vs-oracle is a
 * genuine two-sample test.
 */
#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------ RNG (oracle-private) */
typedef struct { uint64_t s[4]; } rng_t;
static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t splitmix64(uint64_t *x) {
    uint64_t z = (*x += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
static void rng_seed(rng_t *r, uint64_t seed, uint64_t stream) {
    uint64_t x = seed ^ (0xd1342543de82ef95ULL * (stream + 1));
    for (int i = 0; i < 4; i++) r->s[i] = splitmix64(&x);
}
static inline uint64_t rng_next(rng_t *r) {
    uint64_t *s = r->s;
    uint64_t result = rotl64(s[0] + s[3], 23) + s[0];
    uint64_t t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
    s[2] ^= t; s[3] = rotl64(s[3], 45);
    return result;
}
/* rand(): uniform in [0,1) with 53 bits, as Julia's rand(Float64) */
static inline double rng_rand(rng_t *r) { return (double)(rng_next(r) >> 11) * 0x1.0p-53; }
/* randexp(): unit-rate exponential (Julia uses a ziggurat; any exact sampler is equivalent) */
static inline double rng_randexp(rng_t *r) { return -log(1.0 - rng_rand(r)); }

/* ------------------------------------------------------------------ public structs */
typedef struct {
    double kon, koff, konb, reach, CP, antigenconcen, antibodyconcen; /* parameters.jl:15-30 */
} RefBioPars;

typedef struct {
    int32_t N, DIM;              /* parameters.jl:92-109 */
    double tstop, tstop_AtoB;
    const double *tsave;
    int32_t T;
    double L;
    const double *initlocs;      /* N x DIM, DIM contiguous per particle */
    int32_t resample_initlocs;
    int32_t nsims;
} RefSimPars;

typedef struct {
    int32_t kind;                /* 0 = SimNumberTerminator (nsims), 1 = VarianceTerminator */
    double ssetol;               /* callbacks.jl:186-195 */
    int32_t minsims, maxsims;
} RefTerm;

/* utils.jl:10 and parameters.jl:4,138-139 */
double spr_ref_muM_to_inv_cubic_nm(double x) { return 6.02214076e-7 * x; }
double spr_ref_default_antigenconcen(void) { return 500.0 / 149 / 26795 * 1000000; }
double spr_ref_L(int N, int DIM, double antigenconcen, int convert_units) {
    double agc = convert_units ? spr_ref_muM_to_inv_cubic_nm(antigenconcen) : antigenconcen;
    return pow((double)N / agc, 1.0 / DIM);
}

/* forward_simulator.jl:1-8 */
double spr_ref_periodic_dist_sq(const double *p1, const double *p2, int DIM, double L) {
    double d = 0.0;
    for (int i = 0; i < DIM; i++) {
        double daxes = fabs(p1[i] - p2[i]);
        double m = fmin(daxes, L - daxes);
        d += m * m;
    }
    return d;
}

/* ------------------------------------------------------------------ neighbour tables */
typedef struct {
    int N, DIM;
    double *locs;     /* N*DIM (working copy; reference copies simpars.initlocs, :109) */
    int *rp1, *rp2;   /* rpair: directed pairs, first half i<j, second half flipped (:66) */
    int numpairs, cap;
    int *Acnt;        /* degree */
    int *off;         /* CSR offsets into nh/rx (restates Vector{Vector{Int}} nhbrs/rxids) */
    int *nh, *rx;
    int *next;
} nhbr_t;

static void nhbr_init(nhbr_t *n, int N, int DIM) {
    memset(n, 0, sizeof(*n));
    n->N = N; n->DIM = DIM;
    n->locs = (double *)malloc(sizeof(double) * N * DIM);
    n->Acnt = (int *)calloc(N, sizeof(int));
    n->off = (int *)calloc(N + 1, sizeof(int));
    n->next = (int *)calloc(N, sizeof(int));
    n->cap = 1024;
    n->rp1 = (int *)malloc(sizeof(int) * n->cap);
    n->rp2 = (int *)malloc(sizeof(int) * n->cap);
    n->nh = (int *)malloc(sizeof(int) * n->cap);
    n->rx = (int *)malloc(sizeof(int) * n->cap);
}
static void nhbr_free(nhbr_t *n) {
    free(n->locs); free(n->Acnt); free(n->off); free(n->next);
    free(n->rp1); free(n->rp2); free(n->nh); free(n->rx);
}
static void nhbr_reserve(nhbr_t *n, int need) {
    if (need <= n->cap) return;
    while (n->cap < need) n->cap *= 2;
    n->rp1 = (int *)realloc(n->rp1, sizeof(int) * n->cap);
    n->rp2 = (int *)realloc(n->rp2, sizeof(int) * n->cap);
    n->nh = (int *)realloc(n->nh, sizeof(int) * n->cap);
    n->rx = (int *)realloc(n->rx, sizeof(int) * n->cap);
}

/* setup_spr_sim! (forward_simulator.jl:37-89); 0-based indices, reaction ids 0..half-1 */
static void ref_setup(nhbr_t *n, double reach, double L, int resample, rng_t *rng) {
    const int N = n->N, DIM = n->DIM;
    if (resample) { /* :42-47 */
        for (int i = 0; i < N * DIM; i++) n->locs[i] = rng_rand(rng);
        for (int i = 0; i < N * DIM; i++) n->locs[i] *= L;
    }
    const double reachsq = reach * reach;
    memset(n->Acnt, 0, sizeof(int) * N);
    int np = 0;
    for (int i = 0; i < N; i++) { /* :52-61 */
        for (int j = i + 1; j < N; j++) {
            double sqd = spr_ref_periodic_dist_sq(n->locs + (size_t)i * DIM, n->locs + (size_t)j * DIM, DIM, L);
            if (sqd <= reachsq) {
                nhbr_reserve(n, 2 * (np + 1));
                n->rp1[np] = i; n->rp2[np] = j; np++;
                n->Acnt[i]++; n->Acnt[j]++;
            }
        }
    }
    /* :66 append flipped copies */
    for (int p = 0; p < np; p++) { n->rp1[np + p] = n->rp2[p]; n->rp2[np + p] = n->rp1[p]; }
    n->numpairs = 2 * np;
    const int half = np;
    n->off[0] = 0;
    for (int i = 0; i < N; i++) { n->off[i + 1] = n->off[i] + n->Acnt[i]; n->next[i] = 0; }
    int ii = 0; /* :76-86 */
    for (int p = 0; p < n->numpairs; p++) {
        int ridx = n->rp1[p];
        int idx = n->next[ridx];
        n->nh[n->off[ridx] + idx] = n->rp2[p];
        n->rx[n->off[ridx] + idx] = ii;
        n->next[ridx] = idx + 1;
        ii = (ii == half - 1) ? 0 : ii + 1;
    }
}

/* exported for the neighbour-table unit tests: returns number of undirected pairs, fills
 * pairs_out (2*maxpairs ints) with (i,j), i<j, in the reference's scan order */
int spr_ref_pairs(const double *locs, int N, int DIM, double L, double reach, int *pairs_out, int maxpairs) {
    const double reachsq = reach * reach;
    int np = 0;
    for (int i = 0; i < N; i++)
        for (int j = i + 1; j < N; j++)
            if (spr_ref_periodic_dist_sq(locs + (size_t)i * DIM, locs + (size_t)j * DIM, DIM, L) <= reachsq) {
                if (np < maxpairs) { pairs_out[2 * np] = i; pairs_out[2 * np + 1] = j; }
                np++;
            }
    return np;
}

/* ------------------------------------------------------------------ indexed min-heap
 * restates DataStructures.MutableBinaryHeap{Float64,FasterForward}: update!(h, handle, v),
 * top_with_handle(h).  handle == index into key[]. */
typedef struct {
    int M, cap;
    double *key;
    int *heap; /* position -> handle */
    int *hpos; /* handle -> position */
} heap_t;

static void heap_alloc(heap_t *h, int M) {
    if (M > h->cap) {
        h->cap = M + M / 2 + 16;
        h->key = (double *)realloc(h->key, sizeof(double) * h->cap);
        h->heap = (int *)realloc(h->heap, sizeof(int) * h->cap);
        h->hpos = (int *)realloc(h->hpos, sizeof(int) * h->cap);
    }
    h->M = M;
}
static inline void heap_up(heap_t *h, int p) {
    int hd = h->heap[p];
    double k = h->key[hd];
    while (p > 0) {
        int par = (p - 1) >> 1;
        int ph = h->heap[par];
        if (h->key[ph] <= k) break;
        h->heap[p] = ph; h->hpos[ph] = p;
        p = par;
    }
    h->heap[p] = hd; h->hpos[hd] = p;
}
static inline void heap_down(heap_t *h, int p) {
    int hd = h->heap[p];
    double k = h->key[hd];
    const int M = h->M;
    for (;;) {
        int c = 2 * p + 1;
        if (c >= M) break;
        if (c + 1 < M && h->key[h->heap[c + 1]] < h->key[h->heap[c]]) c++;
        int ch = h->heap[c];
        if (h->key[ch] >= k) break;
        h->heap[p] = ch; h->hpos[ch] = p;
        p = c;
    }
    h->heap[p] = hd; h->hpos[hd] = p;
}
/* build from key[0..M) (the MutableBinaryHeap(tvec) constructor, :126/:165) */
static void heap_build(heap_t *h) {
    for (int i = 0; i < h->M; i++) { h->heap[i] = i; h->hpos[i] = i; }
    for (int p = h->M / 2 - 1; p >= 0; p--) heap_down(h, p);
}
static inline void heap_update(heap_t *h, int handle, double v) {
    double old = h->key[handle];
    h->key[handle] = v;
    if (v < old) heap_up(h, h->hpos[handle]);
    else if (v > old) heap_down(h, h->hpos[handle]);
}

/* ------------------------------------------------------------------ run_spr_sim! */
typedef struct { double mu, s2; int64_t n; } var_t; /* OnlineStats.Variance */
static inline void var_fit(var_t *o, double x) {
    double mu = o->mu;
    o->n += 1;
    double g = 1.0 / (double)o->n;
    o->mu = mu + g * (x - mu);                       /* smooth(mu, x, g) */
    o->s2 = o->s2 + g * ((x - o->mu) * (x - mu) - o->s2); /* smooth(s2, (x-mu_new)(x-mu_old), g) */
}
static inline double var_var(const var_t *o) { return o->n > 1 ? o->s2 * ((double)o->n / (double)(o->n - 1)) : 1.0; }

/*
 * observable: 0 = TotalBoundOutputter x=(CP/N)(B+C) (callbacks.jl:28), 1 = TotalAOutputter x=(CP/N)A (:111)
 * mean_out[T], var_out[T] (unbiased; may be NULL), nobs_out, nevents_out (heap pops that changed state)
 * raw_out: optional int32 [maxreps][3][T] copynumbers of every replicate (A,B,C rows, :189-191)
 * returns replicates run, or <0 on bad arguments.
 */
int64_t spr_ref_run(const RefBioPars *bio, const RefSimPars *sim, const RefTerm *term, uint64_t seed,
                    uint64_t stream, int observable, double *mean_out, double *var_out,
                    uint64_t *nevents_out, int32_t *raw_out, int64_t raw_maxreps) {
    const int N = sim->N, DIM = sim->DIM, numsave = sim->T;
    if (N < 1 || (DIM != 2 && DIM != 3) || numsave < 1) return -1;
    const double tstop = sim->tstop, tstop_AtoB = sim->tstop_AtoB, L = sim->L;
    const double *tsave = sim->tsave;
    const int resample = sim->resample_initlocs;
    rng_t rng;
    rng_seed(&rng, seed, stream);

    nhbr_t nb;
    nhbr_init(&nb, N, DIM);
    if (sim->initlocs) memcpy(nb.locs, sim->initlocs, sizeof(double) * N * DIM); /* :109 */
    else if (!resample) { nhbr_free(&nb); return -2; }
    ref_setup(&nb, bio->reach, L, resample, &rng); /* :116 */
    int numpairs = nb.numpairs, half = numpairs / 2;
    const int twoN = 2 * N;

    int *states = (int *)malloc(sizeof(int) * N);
    int *copyn = (int *)malloc(sizeof(int) * 3 * numsave);
    heap_t times;
    memset(&times, 0, sizeof(times));
    heap_alloc(&times, twoN + numpairs);
    for (int i = 0; i < times.M; i++) times.key[i] = 0.0; /* tvec = zeros (:125) */
    heap_build(&times);
    int tvec_len = twoN + numpairs;

    var_t *bind = (var_t *)calloc(numsave, sizeof(var_t));
    uint64_t nevents = 0;
    int64_t nrun = 0;
    int notdone = 1;
    const double scale = bio->CP / (double)N;

    for (;;) {
        /* isnotdone (:129; callbacks.jl:160-162, 199-201) */
        if (term->kind == 0) { if (!(nrun < sim->nsims)) break; }
        else if (!notdone) break;

        double tkon = 1.0 / (bio->kon * bio->antibodyconcen); /* :132-135 */
        const double tkoff = 1.0 / bio->koff;
        const double tkonb = 1.0 / bio->konb;
        const double tkc = 1.0 / (2 * bio->koff);

        if (resample) { /* :138-142 */
            ref_setup(&nb, bio->reach, L, 1, &rng);
            numpairs = nb.numpairs; half = numpairs / 2;
        }
        double tc = 0.0;
        int A = N, B = 0, C = 0;
        for (int i = 0; i < N; i++) states[i] = 1;
        memset(copyn, 0, sizeof(int) * 3 * numsave);
        int sidx = 0;
        double tp = tsave[0];

        /* :157-170 initial reaction times */
        int rebuild = resample && (tvec_len != twoN + numpairs);
        if (rebuild) { heap_alloc(&times, twoN + numpairs); tvec_len = twoN + numpairs; }
        if (rebuild) {
            for (int i = 0; i < N; i++) times.key[i] = tkon * rng_randexp(&rng);
            for (int i = N; i < tvec_len; i++) times.key[i] = INFINITY;
            heap_build(&times);
        } else {
            /* reference updates every handle in index order */
            double *tv = (double *)malloc(sizeof(double) * N);
            for (int i = 0; i < N; i++) tv[i] = tkon * rng_randexp(&rng);
            for (int i = 0; i < N; i++) heap_update(&times, i, tv[i]);
            for (int i = N; i < tvec_len; i++) heap_update(&times, i, INFINITY);
            free(tv);
        }
        int turnoff = 0;
        const int *rp1 = nb.rp1, *rp2 = nb.rp2, *off = nb.off, *nh = nb.nh, *rx = nb.rx;

        while (tc <= tstop) { /* :173 */
            if (tc >= tstop_AtoB && !turnoff) { /* :174-181 */
                tkon = INFINITY;
                for (int i = 0; i < N; i++) heap_update(&times, i, INFINITY);
                turnoff = 1;
            }
            int rxidx = times.heap[0]; /* top_with_handle (:183) */
            double tnext = times.key[rxidx];
            tc = tnext;
            while (tc > tp && sidx < numsave) { /* :188-194 */
                copyn[0 * numsave + sidx] = A;
                copyn[1 * numsave + sidx] = B;
                copyn[2 * numsave + sidx] = C;
                sidx++;
                tp = (sidx < numsave) ? tsave[sidx] : nextafter(tstop, INFINITY);
            }
            if (isinf(tc)) break; /* nothing can fire any more: all slots were just filled */

            if (rxidx < N) {
                int m = rxidx;
                if (tc >= tstop_AtoB) { /* :199-202 */
                    heap_update(&times, m, INFINITY);
                } else { /* A --> B :204-224 */
                    A--; B++; nevents++;
                    states[m] = 2;
                    heap_update(&times, m, INFINITY);
                    heap_update(&times, N + m, tc + tkoff * rng_randexp(&rng));
                    for (int q = off[m]; q < off[m + 1]; q++) {
                        if (states[nh[q]] == 1) heap_update(&times, twoN + rx[q], tc + tkonb * rng_randexp(&rng));
                        else heap_update(&times, twoN + rx[q], INFINITY);
                    }
                }
            } else if (rxidx < twoN) { /* B --> A :228-253 */
                int m = rxidx - N;
                A++; B--; nevents++;
                states[m] = 1;
                heap_update(&times, N + m, INFINITY);
                if (!turnoff) heap_update(&times, m, tc + tkon * rng_randexp(&rng));
                else heap_update(&times, m, INFINITY);
                for (int q = off[m]; q < off[m + 1]; q++) {
                    if (states[nh[q]] == 2) heap_update(&times, twoN + rx[q], tc + tkonb * rng_randexp(&rng));
                    else heap_update(&times, twoN + rx[q], INFINITY);
                }
            } else if (rxidx < twoN + half) { /* A + B --> C :256-289 */
                int cur = rxidx - twoN;
                A--; B--; C++; nevents++;
                int m1 = rp1[cur], m2 = rp2[cur];
                if (states[m1] == 2) { int t = m1; m1 = m2; m2 = t; }
                states[m1] = 0; states[m2] = 0;
                heap_update(&times, m1, INFINITY);
                heap_update(&times, m2, INFINITY);
                heap_update(&times, m1 + N, INFINITY);
                heap_update(&times, m2 + N, INFINITY);
                for (int q = off[m1]; q < off[m1 + 1]; q++) heap_update(&times, twoN + rx[q], INFINITY);
                for (int q = off[m2]; q < off[m2 + 1]; q++) heap_update(&times, twoN + rx[q], INFINITY);
                heap_update(&times, twoN + half + cur, tc + tkc * rng_randexp(&rng));
            } else { /* C --> A + B :291-344 */
                int cur = rxidx - (twoN + half);
                A++; B++; C--; nevents++;
                heap_update(&times, cur + twoN + half, INFINITY);
                double pAB = rng_rand(&rng);
                int mA = rp1[cur], mB = rp2[cur];
                if (pAB >= 0.5) { int t = mA; mA = mB; mB = t; }
                states[mA] = 1; states[mB] = 2;
                if (!turnoff) heap_update(&times, mA, tc + tkon * rng_randexp(&rng));
                else heap_update(&times, mA, INFINITY);
                heap_update(&times, N + mA, INFINITY);
                heap_update(&times, mB, INFINITY);
                heap_update(&times, N + mB, tc + tkoff * rng_randexp(&rng));
                for (int q = off[mA]; q < off[mA + 1]; q++) {
                    if (states[nh[q]] == 2) heap_update(&times, twoN + rx[q], tc + tkonb * rng_randexp(&rng));
                    else heap_update(&times, twoN + rx[q], INFINITY);
                }
                for (int q = off[mB]; q < off[mB + 1]; q++) {
                    int nb_ = nh[q];
                    if (states[nb_] == 1 && nb_ != mA) heap_update(&times, twoN + rx[q], tc + tkonb * rng_randexp(&rng));
                    else if (states[nb_] != 1) heap_update(&times, twoN + rx[q], INFINITY);
                }
            }
        }
        while (sidx < numsave) { /* :349-354 */
            copyn[0 * numsave + sidx] = A;
            copyn[1 * numsave + sidx] = B;
            copyn[2 * numsave + sidx] = C;
            sidx++;
        }
        /* outputter(tsave, copynumbers, biopars, simpars) (:357) */
        for (int s = 0; s < numsave; s++) {
            double x = observable == 0 ? scale * (double)(copyn[numsave + s] + copyn[2 * numsave + s])
                                       : scale * (double)copyn[s];
            var_fit(&bind[s], x);
        }
        if (raw_out && nrun < raw_maxreps) {
            int32_t *dst = raw_out + (size_t)nrun * 3 * numsave;
            for (int q = 0; q < 3 * numsave; q++) dst[q] = copyn[q];
        }
        nrun++;
        /* update!(terminator, ...) (:360; callbacks.jl:204-217) */
        if (term->kind == 1) {
            int64_t ns = bind[0].n;
            if (ns < term->minsims) notdone = 1;
            else if (ns >= term->maxsims) notdone = 0;
            else {
                int any = 0;
                for (int s = 0; s < numsave && !any; s++) {
                    double sd = sqrt(var_var(&bind[s]));
                    if (sd > term->ssetol * bind[s].mu * sqrt((double)bind[s].n)) any = 1;
                }
                notdone = any;
            }
        }
    }
    for (int s = 0; s < numsave; s++) {
        if (mean_out) mean_out[s] = bind[s].mu;
        if (var_out) var_out[s] = var_var(&bind[s]);
    }
    if (nevents_out) *nevents_out = nevents;
    free(bind); free(states); free(copyn);
    free(times.key); free(times.heap); free(times.hpos);
    nhbr_free(&nb);
    return nrun;
}

/* ------------------------------------------------------------------ surrogate sweep
 * build_surrogate_slice (surrogate.jl:273-300): linear index idx (1-based, kon fastest ...
 * reach slowest, CartesianIndices(size[1:4])) -> 10^logkon, 10^logkoff, 10^logkonb, reach with
 * grids range(lo, hi, length=n) (surrogate.jl:247-250), CP = 1, antibodyconcen = 1;
 * output column n = means(tbo) of index idxstart+n-1, T x nidx, time fastest (:283,292).
 * Parallelism = the reference's own: independent index ranges (examples/cluster_scripts/
 * queue_job.sh:10-17), here one pthread per core pulling indices from an atomic counter.
 */
typedef struct {
    double lo[4], hi[4];
    int32_t n[4];
} RefGrid;

/* Julia's range(lo, hi, length=n)[i] for Float64 is computed in twice precision; to double
 * rounding it equals lo + (i-1)*(hi-lo)/(n-1) for the ranges used here. */
double spr_ref_grid_value(double lo, double hi, int n, int i0) {
    if (n == 1) return lo;
    if (i0 == n - 1) return hi;
    return lo + (double)i0 * ((hi - lo) / (double)(n - 1));
}

typedef struct {
    const RefGrid *grid;
    const RefSimPars *sim;
    const RefTerm *term;
    int64_t idxstart, nidx;
    uint64_t seed;
    double *out;          /* T x nidx */
    int64_t *nobs;        /* nidx (may be NULL) */
    uint64_t *nevents;    /* nidx (may be NULL) */
    atomic_llong next;
} slice_job_t;

static void *slice_worker(void *arg) {
    slice_job_t *job = (slice_job_t *)arg;
    const RefGrid *g = job->grid;
    for (;;) {
        long long k = atomic_fetch_add(&job->next, 1);
        if (k >= job->nidx) break;
        int64_t lin = job->idxstart - 1 + k; /* 0-based linear index */
        int i1 = (int)(lin % g->n[0]);
        int i2 = (int)((lin / g->n[0]) % g->n[1]);
        int i3 = (int)((lin / ((int64_t)g->n[0] * g->n[1])) % g->n[2]);
        int i4 = (int)(lin / ((int64_t)g->n[0] * g->n[1] * g->n[2]));
        RefBioPars bio;
        bio.kon = pow(10.0, spr_ref_grid_value(g->lo[0], g->hi[0], g->n[0], i1)); /* setpars! :257-264 */
        bio.koff = pow(10.0, spr_ref_grid_value(g->lo[1], g->hi[1], g->n[1], i2));
        bio.konb = pow(10.0, spr_ref_grid_value(g->lo[2], g->hi[2], g->n[2], i3));
        bio.reach = spr_ref_grid_value(g->lo[3], g->hi[3], g->n[3], i4);
        bio.CP = 1.0; bio.antigenconcen = spr_ref_default_antigenconcen(); bio.antibodyconcen = 1.0;
        uint64_t nev = 0;
        int64_t n = spr_ref_run(&bio, job->sim, job->term, job->seed, (uint64_t)lin, 0,
                                job->out + (size_t)k * job->sim->T, NULL, &nev, NULL, 0);
        if (job->nobs) job->nobs[k] = n;
        if (job->nevents) job->nevents[k] = nev;
    }
    return NULL;
}

int spr_ref_build_slice(const RefGrid *grid, const RefSimPars *sim, const RefTerm *term,
                        int64_t idxstart, int64_t idxend, uint64_t seed, int nthreads,
                        double *out_TxNidx, int64_t *nobs, uint64_t *nevents) {
    int64_t total = (int64_t)grid->n[0] * grid->n[1] * grid->n[2] * grid->n[3];
    if (idxstart < 1 || idxend > total || idxend < idxstart) return -1;
    slice_job_t job;
    job.grid = grid; job.sim = sim; job.term = term;
    job.idxstart = idxstart; job.nidx = idxend - idxstart + 1;
    job.seed = seed; job.out = out_TxNidx; job.nobs = nobs; job.nevents = nevents;
    atomic_init(&job.next, 0);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 1024) nthreads = 1024;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, slice_worker, &job);
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th);
    return 0;
}

/* explicit list of parameter points (used by the bench's bounded CPU sample) */
typedef struct {
    const RefBioPars *pts;
    const RefSimPars *sim;
    const RefTerm *term;
    int64_t npts;
    uint64_t seed;
    const uint64_t *streams; /* per-point stream id (may be NULL -> index) */
    double *out;
    int64_t *nobs;
    uint64_t *nevents;
    atomic_llong next;
} points_job_t;

static void *points_worker(void *arg) {
    points_job_t *job = (points_job_t *)arg;
    for (;;) {
        long long k = atomic_fetch_add(&job->next, 1);
        if (k >= job->npts) break;
        uint64_t nev = 0;
        int64_t n = spr_ref_run(&job->pts[k], job->sim, job->term, job->seed,
                                job->streams ? job->streams[k] : (uint64_t)k, 0,
                                job->out ? job->out + (size_t)k * job->sim->T : NULL, NULL, &nev, NULL, 0);
        if (job->nobs) job->nobs[k] = n;
        if (job->nevents) job->nevents[k] = nev;
    }
    return NULL;
}

int spr_ref_run_points(const RefBioPars *pts, int64_t npts, const RefSimPars *sim, const RefTerm *term,
                       uint64_t seed, const uint64_t *streams, int nthreads, double *out_TxNpts,
                       int64_t *nobs, uint64_t *nevents) {
    points_job_t job;
    job.pts = pts; job.sim = sim; job.term = term; job.npts = npts; job.seed = seed;
    job.streams = streams; job.out = out_TxNpts; job.nobs = nobs; job.nevents = nevents;
    atomic_init(&job.next, 0);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 1024) nthreads = 1024;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, points_worker, &job);
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th);
    return 0;
}

/* the same sweep with a replicate count per point and the wall seconds each point took on its thread while
 * all other threads were busy (bench.py: stratified, bounded sample of a grid; FIXED terminator only when
 * nsims_per_point is given) */
typedef struct {
    points_job_t base;
    const int32_t *nsims;
    double *seconds;
} timed_job_t;

static void *timed_worker(void *arg) {
    timed_job_t *tj = (timed_job_t *)arg;
    points_job_t *job = &tj->base;
    for (;;) {
        long long k = atomic_fetch_add(&job->next, 1);
        if (k >= job->npts) break;
        RefSimPars sim = *job->sim;
        if (tj->nsims) sim.nsims = tj->nsims[k];
        struct timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        uint64_t nev = 0;
        int64_t n = spr_ref_run(&job->pts[k], &sim, job->term, job->seed,
                                job->streams ? job->streams[k] : (uint64_t)k, 0,
                                job->out ? job->out + (size_t)k * sim.T : NULL, NULL, &nev, NULL, 0);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        if (job->nobs) job->nobs[k] = n;
        if (job->nevents) job->nevents[k] = nev;
        if (tj->seconds) tj->seconds[k] = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    }
    return NULL;
}

int spr_ref_run_points_timed(const RefBioPars *pts, int64_t npts, const RefSimPars *sim, const RefTerm *term,
                             uint64_t seed, const uint64_t *streams, const int32_t *nsims_per_point, int nthreads,
                             double *out_TxNpts, int64_t *nobs, uint64_t *nevents, double *seconds) {
    timed_job_t tj;
    tj.base.pts = pts; tj.base.sim = sim; tj.base.term = term; tj.base.npts = npts; tj.base.seed = seed;
    tj.base.streams = streams; tj.base.out = out_TxNpts; tj.base.nobs = nobs; tj.base.nevents = nevents;
    tj.nsims = nsims_per_point; tj.seconds = seconds;
    atomic_init(&tj.base.next, 0);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 1024) nthreads = 1024;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, timed_worker, &tj);
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th);
    return 0;
}
