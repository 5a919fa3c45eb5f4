/*
 * oracle/spr_mirror.c -- sequential CPU mirror of the *B200 trajectory specification*.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Never included, linked or called by the product.
 *
 * Purpose: the reference (Julia, unseeded RNG) defines the model only in distribution, so parity with
 * it is statistical (oracle/spr_ref.c).  The product's own integer and index work -- Philox streams,
 * lattice positions, cell-list neighbour tables, list bookkeeping, channel selection, save-slot
 * filling, moment sums, stopping index -- is specified exactly in DESIGN.md section 3; this file
 * re-implements that specification independently in plain sequential C so that the CUDA kernels can
 * be checked BIT FOR BIT (tests/test_gpu_mirror.py).  It shares no code with the product: Philox,
 * the deterministic -log(U) and every loop are written again here.
 *
 * The model itself is the reference's (all paths under /root/reference):
 *   channels and rates          src/forward_simulator.jl:198-344 (A->B kon*[Ab], B->A koff,
 *                               A+B->C konb per within-reach pair, C->A+B 2*koff with a fair coin)
 *   switch-off at tstop_AtoB    src/forward_simulator.jl:174-181, 199-202
 *   save-slot semantics         src/forward_simulator.jl:188-194, 349-354
 *   periodic min-image distance src/forward_simulator.jl:1-8
 *   observables                 src/callbacks.jl:28 (B+C), :111 (A)
 *   stopping rule               src/callbacks.jl:204-217
 * Direct-method (Gillespie) sampling of the same continuous-time Markov chain replaces the
 * reference's next-reaction heap: every waiting time there is exponential with a state-constant
 * rate, so both sample the same process.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- Philox4x32-10 */
static void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                          uint32_t out[4]) {
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
void spr_mirror_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
}

/* ---------------------------------------------------------------- -log(k * 2^-53), k in [1, 2^53] */
double spr_mirror_neg_log_u53(uint64_t k) {
    double d = (double)k;
    uint64_t bits;
    memcpy(&bits, &d, 8);
    int e = (int)((bits >> 52) & 0x7ff) - 1023;
    uint64_t mb = (bits & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL;
    double m;
    memcpy(&m, &mb, 8);
    if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
    double s = (m + -1.0) / (m + 1.0);
    double z = s * s;
    double p = 1.0 / 23.0;
    p = fma(p, z, 1.0 / 21.0);
    p = fma(p, z, 1.0 / 19.0);
    p = fma(p, z, 1.0 / 17.0);
    p = fma(p, z, 1.0 / 15.0);
    p = fma(p, z, 1.0 / 13.0);
    p = fma(p, z, 1.0 / 11.0);
    p = fma(p, z, 1.0 / 9.0);
    p = fma(p, z, 1.0 / 7.0);
    p = fma(p, z, 1.0 / 5.0);
    p = fma(p, z, 1.0 / 3.0);
    double t2 = s + s;
    double t2z = t2 * z;
    double logm = fma(t2z, p, t2);
    return fma((double)(53 - e), 0.6931471805599453, -logm);
}

/* ---------------------------------------------------------------- waiting time E / a0:
 * E times the correctly rounded reciprocal of a0 (two IEEE operations); a0 == 0 gives +inf. */
double spr_mirror_next_time(double t, double E, double a0) {
    const double r = 1.0 / a0;
    return t + E * r;
}

/* ---------------------------------------------------------------- cell grid */
int spr_mirror_cells_per_axis(int N, int DIM, double L, double reach) {
    int cmax = 1;
    long long lim = N < 1024 ? (N < 1 ? 1 : N) : 1024;
    for (;;) {
        long long c = cmax + 1, v = 1;
        for (int d = 0; d < DIM; d++) v *= c;
        if (v > lim) break;
        cmax++;
    }
    int nc = cmax;
    if (reach > 0.0) {
        double q = L / reach;
        if (q < (double)cmax) nc = (int)q;
    }
    if (nc < 3) nc = 1;
    return nc;
}

#define POS_BITS 21
#define POS_MASK ((1u << POS_BITS) - 1u)

typedef struct {
    int N, DIM;
    uint32_t *spos;   /* N*DIM lattice coordinates (21 bits per axis), sorted numbering */
    int *off;         /* N+1 */
    int *nbr;
    int nbr_cap, entries;
} table_t;

/* exact squared min-image distance in lattice units vs r2lat = floor(reach^2 / (L 2^-21)^2) */
static int within(const uint32_t *a, const uint32_t *b, int DIM, uint64_t r2lat) {
    uint64_t acc = 0;
    for (int d = 0; d < DIM; d++) {
        uint32_t df = (a[d] - b[d]) & POS_MASK;
        uint32_t nd = (0u - df) & POS_MASK;
        uint64_t m = df < nd ? df : nd;
        acc += m * m;
    }
    return acc <= r2lat;
}

static uint64_t reach2_lattice(double L, double reach) {
    double scale = L * 0x1.0p-21;
    double num = reach * reach;
    double den = scale * scale;
    double q = num / den;
    if (!(q < 4611686018427387904.0)) return 4611686018427387904ULL;
    return (uint64_t)q;
}

static void table_push(table_t *t, int j) {
    if (t->entries == t->nbr_cap) {
        t->nbr_cap = t->nbr_cap ? 2 * t->nbr_cap : 1024;
        t->nbr = (int *)realloc(t->nbr, sizeof(int) * t->nbr_cap);
    }
    t->nbr[t->entries++] = j;
}

/* resampled positions -> cell-sorted numbering -> neighbour table */
static void build_table_resampled(table_t *t, double L, double reach, uint32_t rep, uint32_t pidx, uint32_t k0,
                                  uint32_t k1) {
    const int N = t->N, DIM = t->DIM;
    const int nc = spr_mirror_cells_per_axis(N, DIM, L, reach);
    int ncell = 1;
    for (int d = 0; d < DIM; d++) ncell *= nc;
    const uint64_t r2lat = reach2_lattice(L, reach);
    uint32_t *raw = (uint32_t *)malloc(sizeof(uint32_t) * N * DIM);
    int *cell = (int *)malloc(sizeof(int) * N);
    int *cstart = (int *)calloc(ncell + 1, sizeof(int));
    for (int i = 0; i < N; i++) {
        uint32_t r[4];
        philox4x32_10((uint32_t)i, 0u /* positions stream */, rep, pidx, k0, k1, r);
        int c = 0;
        for (int d = 0; d < DIM; d++) raw[i * DIM + d] = r[d] >> (32 - POS_BITS);   /* top 21 bits */
        for (int d = DIM - 1; d >= 0; d--) c = c * nc + (int)((raw[i * DIM + d] * (uint32_t)nc) >> POS_BITS);
        cell[i] = c;
        cstart[c + 1]++;
    }
    for (int c = 0; c < ncell; c++) cstart[c + 1] += cstart[c];
    int *fill = (int *)malloc(sizeof(int) * (ncell + 1));
    memcpy(fill, cstart, sizeof(int) * (ncell + 1));
    for (int i = 0; i < N; i++) { /* stable: ascending original index inside a cell */
        int slot = fill[cell[i]]++;
        for (int d = 0; d < DIM; d++) t->spos[slot * DIM + d] = raw[i * DIM + d];
    }
    t->entries = 0;
    const int lo = nc >= 3 ? -1 : 0, hi = nc >= 3 ? 1 : 0;
    for (int i = 0; i < N; i++) {
        t->off[i] = t->entries;
        const uint32_t *ki = t->spos + i * DIM;
        int ci[3] = {0, 0, 0};
        for (int d = 0; d < DIM; d++) ci[d] = (int)((ki[d] * (uint32_t)nc) >> POS_BITS);
        const int zlo = DIM == 3 ? lo : 0, zhi = DIM == 3 ? hi : 0;
        for (int dz = zlo; dz <= zhi; dz++) {
            int cz = 0;
            if (DIM == 3) { cz = ci[2] + dz; cz = cz < 0 ? cz + nc : (cz >= nc ? cz - nc : cz); }
            for (int dy = lo; dy <= hi; dy++) {
                int cy = ci[1] + dy; cy = cy < 0 ? cy + nc : (cy >= nc ? cy - nc : cy);
                for (int dx = lo; dx <= hi; dx++) {
                    int cx = ci[0] + dx; cx = cx < 0 ? cx + nc : (cx >= nc ? cx - nc : cx);
                    int c = (cz * nc + cy) * nc + cx;
                    for (int j = cstart[c]; j < cstart[c + 1]; j++)
                        if (j != i && within(ki, t->spos + j * DIM, DIM, r2lat)) table_push(t, j);
                }
            }
        }
    }
    t->off[N] = t->entries;
    free(raw); free(cell); free(cstart); free(fill);
}

/* fixed FP64 positions: the reference's arithmetic and scan order (forward_simulator.jl:1-8, 52-61) */
static void build_table_fixed(table_t *t, const double *locs, double L, double reach) {
    const int N = t->N, DIM = t->DIM;
    const double reach2 = reach * reach;
    t->entries = 0;
    for (int i = 0; i < N; i++) {
        t->off[i] = t->entries;
        for (int j = 0; j < N; j++) {
            if (j == i) continue;
            double d = 0.0;
            for (int k = 0; k < DIM; k++) {
                double dax = fabs(locs[i * DIM + k] - locs[j * DIM + k]);
                double m = fmin(dax, L - dax);
                d = d + m * m;
            }
            if (d <= reach2) table_push(t, j);
        }
    }
    t->off[N] = t->entries;
}

/* ---------------------------------------------------------------- one trajectory */
enum { ST_C = 0, ST_A = 1, ST_B = 2 };

typedef struct {
    int32_t N, DIM;
    double L, tstop, toff;
    int32_t T;
    const double *tsave;
    int32_t resample;
    const double *initlocs;
} MirrorSim;

static void bump(const table_t *t, int *cnt, int m, int d) {
    for (int q = t->off[m]; q < t->off[m + 1]; q++) cnt[t->nbr[q]] += d;
}

/* observable: 0 -> curve1 = B+C ; 1 -> curve1 = A ; 2 -> curve1 = B, curve2 = C */
int spr_mirror_trajectory(const MirrorSim *sim, double kon_eff, double koff, double konb, double reach,
                          uint64_t seed, uint32_t pidx, uint32_t rep, int observable, uint16_t *curve1,
                          uint16_t *curve2, uint64_t *nevents_out, int32_t *off_out, int32_t *nbr_out,
                          int64_t nbr_cap, double *pos_out) {
    const int N = sim->N, DIM = sim->DIM, T = sim->T;
    const uint32_t k0 = (uint32_t)(seed & 0xffffffffu), k1 = (uint32_t)(seed >> 32);
    table_t tb;
    memset(&tb, 0, sizeof(tb));
    tb.N = N; tb.DIM = DIM;
    tb.spos = (uint32_t *)calloc((size_t)N * DIM, sizeof(uint32_t));
    tb.off = (int *)calloc(N + 1, sizeof(int));
    if (sim->resample) build_table_resampled(&tb, sim->L, reach, rep, pidx, k0, k1);
    else build_table_fixed(&tb, sim->initlocs, sim->L, reach);
    if (off_out) for (int i = 0; i <= N; i++) off_out[i] = tb.off[i];
    if (nbr_out) for (int q = 0; q < tb.entries && q < nbr_cap; q++) nbr_out[q] = tb.nbr[q];
    if (pos_out) {
        const double scale = sim->L * 0x1.0p-21;
        for (int q = 0; q < N * DIM; q++)
            pos_out[q] = sim->resample ? (double)tb.spos[q] * scale : sim->initlocs[q];
    }
    if (!curve1) { /* table only */
        int e = tb.entries;
        free(tb.spos); free(tb.off); free(tb.nbr);
        return e;
    }
    int *state = (int *)malloc(sizeof(int) * N);
    int *nA = (int *)malloc(sizeof(int) * N);
    int *nB = (int *)calloc(N, sizeof(int));
    int *listA = (int *)malloc(sizeof(int) * N);
    int *listB = (int *)malloc(sizeof(int) * N);
    int *pos = (int *)malloc(sizeof(int) * N);
    uint32_t *listC = (uint32_t *)malloc(sizeof(uint32_t) * (N / 2 + 1));
    for (int i = 0; i < N; i++) {
        state[i] = ST_A; nA[i] = tb.off[i + 1] - tb.off[i];
        listA[i] = i; pos[i] = i;
    }
    int cntA = N, cntB = 0, cntC = 0, nAB = 0;
    double t = 0.0, kon = kon_eff;
    const double kc = 2.0 * koff;
    int on = 1, sidx = 0;
    const double *tsave = sim->tsave;
    double tp = tsave[0];
    uint32_t ndraw = 0;
    uint64_t nev = 0;
    const double tstop = sim->tstop, toff = sim->toff;

    for (;;) {
        uint32_t r[4];
        philox4x32_10(ndraw, 1u /* events stream */, rep, pidx, k0, k1, r);
        ndraw++;
        const uint64_t k53 = (((uint64_t)r[0] << 21) | (uint64_t)(r[1] >> 11)) + 1ULL;
        const double E = spr_mirror_neg_log_u53(k53);
        const double cA = kon * (double)cntA;
        const double cB = fma(koff, (double)cntB, cA);
        const double cP = fma(konb, (double)nAB, cB);
        const double a0 = fma(kc, (double)cntC, cP);
        double tn = spr_mirror_next_time(t, E, a0);
        /* an event strictly before tfast = min(next save time, switch-off time while the bath is on, tstop) needs no
         * save-slot write, no switch-off and is not the last one; every other event takes the checks below */
        const double tfast = fmin(fmin(tp, on ? toff : INFINITY), tstop);
        const int fast = tn < tfast;
        int lastev = 0;
        if (!fast) {
            int switched = 0;
            if (on && tn >= toff) { tn = toff; switched = 1; }
            if (tn > tp) {
                while (sidx < T && tsave[sidx] < tn) {
                    curve1[sidx] = (uint16_t)(observable == 1 ? cntA : (observable == 2 ? cntB : cntB + cntC));
                    if (curve2) curve2[sidx] = (uint16_t)cntC;
                    sidx++;
                }
                tp = sidx < T ? tsave[sidx] : INFINITY;
            }
            if (switched) { t = toff; kon = 0.0; on = 0; continue; }
            if (!(tn < INFINITY)) break;
            lastev = tn > tstop;
        }
        t = tn;
        nev++;
        const double x = ((double)r[2] * 0x1.0p-32) * a0;
        if (x < cA) { /* A -> B */
            int k = (int)(((uint64_t)r[3] * (uint64_t)(uint32_t)cntA) >> 32);
            int m = listA[k], last = listA[cntA - 1];
            listA[k] = last; pos[last] = k; cntA--;
            listB[cntB] = m; pos[m] = cntB; cntB++;
            nAB += nA[m] - nB[m];
            state[m] = ST_B;
            bump(&tb, nA, m, -1); bump(&tb, nB, m, +1);
        } else if (x < cB) { /* B -> A */
            int k = (int)(((uint64_t)r[3] * (uint64_t)(uint32_t)cntB) >> 32);
            int m = listB[k], last = listB[cntB - 1];
            listB[k] = last; pos[last] = k; cntB--;
            listA[cntA] = m; pos[m] = cntA; cntA++;
            nAB += nB[m] - nA[m];
            state[m] = ST_A;
            bump(&tb, nA, m, +1); bump(&tb, nB, m, -1);
        } else if (x < cP) { /* A + B -> C */
            int k = (int)(((uint64_t)r[3] * (uint64_t)(uint32_t)nAB) >> 32);
            const int sideA = cntA <= cntB;
            const int *list = sideA ? listA : listB;
            const int len = sideA ? cntA : cntB;
            const int *w = sideA ? nB : nA;
            const int want = sideA ? ST_B : ST_A;
            int ix = 0;
            while (ix < len && k >= w[list[ix]]) { k -= w[list[ix]]; ix++; }
            if (ix >= len) { ix = -1; } /* cannot happen: nAB == sum of weights */
            int xsel = list[ix], ysel = -1;
            for (int q = tb.off[xsel]; q < tb.off[xsel + 1]; q++) {
                int j = tb.nbr[q];
                if (state[j] == want) { if (k == 0) { ysel = j; break; } k--; }
            }
            int ma = sideA ? xsel : ysel, mb = sideA ? ysel : xsel;
            int ia = sideA ? ix : pos[ma], ib = sideA ? pos[mb] : ix;
            int lastA = listA[cntA - 1]; listA[ia] = lastA; pos[lastA] = ia; cntA--;
            int lastB = listB[cntB - 1]; listB[ib] = lastB; pos[lastB] = ib; cntB--;
            listC[cntC++] = (uint32_t)ma | ((uint32_t)mb << 16);
            nAB -= nB[ma] + nA[mb] - 1;
            state[ma] = ST_C; state[mb] = ST_C;
            bump(&tb, nA, ma, -1); bump(&tb, nB, mb, -1);
        } else { /* C -> A + B */
            int k = (int)(((uint64_t)r[3] * (uint64_t)(uint32_t)cntC) >> 32);
            uint32_t pr = listC[k];
            listC[k] = listC[cntC - 1]; cntC--;
            int ma = (int)(pr & 0xffffu), mb = (int)(pr >> 16);
            if (r[3] & 1u) { int tmp = ma; ma = mb; mb = tmp; }
            listA[cntA] = ma; pos[ma] = cntA; cntA++;
            listB[cntB] = mb; pos[mb] = cntB; cntB++;
            nAB += nB[ma] + nA[mb] + 1;
            state[ma] = ST_A; state[mb] = ST_B;
            bump(&tb, nA, ma, +1); bump(&tb, nB, mb, +1);
        }
        if (lastev) break;
    }
    for (int s = sidx; s < T; s++) {
        curve1[s] = (uint16_t)(observable == 1 ? cntA : (observable == 2 ? cntB : cntB + cntC));
        if (curve2) curve2[s] = (uint16_t)cntC;
    }
    if (nevents_out) *nevents_out = nev;
    int e = tb.entries;
    free(state); free(nA); free(nB); free(listA); free(listB); free(pos); free(listC);
    free(tb.spos); free(tb.off); free(tb.nbr);
    return e;
}

/* ---------------------------------------------------------------- a parameter point: replicates,
 * integer moments, stopping index (same rule as the product's K3, callbacks.jl:204-217) */
int spr_mirror_point(const MirrorSim *sim, double kon_eff, double koff, double konb, double reach, uint64_t seed,
                     uint32_t pidx, int term_kind, int nsims, double ssetol, int minsims, int maxsims,
                     int observable, int64_t *sum, uint64_t *sumsq, int32_t *nused, uint64_t *nevents) {
    const int T = sim->T;
    uint16_t *curve = (uint16_t *)malloc(sizeof(uint16_t) * T);
    for (int s = 0; s < T; s++) { sum[s] = 0; sumsq[s] = 0; }
    uint64_t evtot = 0;
    int n = 0;
    const int nmax = term_kind == 0 ? nsims : maxsims;
    const double tol2 = ssetol * ssetol;
    while (n < nmax) {
        uint64_t ev = 0;
        spr_mirror_trajectory(sim, kon_eff, koff, konb, reach, seed, pidx, (uint32_t)n, observable, curve, NULL, &ev,
                              NULL, NULL, 0, NULL);
        evtot += ev;
        n++;
        int viol = 0;
        for (int s = 0; s < T; s++) {
            uint64_t v = curve[s];
            sum[s] += (int64_t)v;
            sumsq[s] += v * v;
            if (term_kind == 1 && n >= minsims && n < maxsims) {
                double dS = (double)sum[s];
                /* n*Q - S^2 in 128 bits, converted as hi * 2^64 + lo */
                unsigned __int128 d = (unsigned __int128)(uint64_t)n * sumsq[s] -
                                      (unsigned __int128)(uint64_t)sum[s] * (uint64_t)sum[s];
                uint64_t hi = (uint64_t)(d >> 64), lo = (uint64_t)d;
                double num = (double)lo;
                if (hi) num = (double)hi * 18446744073709551616.0 + num;
                if (num > tol2 * dS * dS * (double)(n - 1)) viol = 1;
            }
        }
        if (term_kind == 1 && n >= minsims && (n >= maxsims || !viol)) break;
    }
    *nused = n;
    if (nevents) *nevents = evtot;
    free(curve);
    return 0;
}
