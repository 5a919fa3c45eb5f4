/* c_client.c -- a plain C11 client of include/sprb200.h (the third client of the ABI next to the ctypes binding
 * and the Julia shim).  Built by tests/test_c_client.py with  gcc -std=c11 -Wall -Wextra -pedantic.
 *
 *   c_client layout            prints sizeof / offsetof of every struct of the header (no GPU needed); the test
 *                              compares them with the ctypes and the Julia declarations
 *   c_client run <out.bin>     one sprb200_simulate and one sprb200_build_slice call on device 0; writes
 *                              sum[2][T] (int64), n_events[2] (uint64), slice[4][T] (double) for the test to
 *                              compare bit for bit with the same calls through ctypes
 */
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sprb200.h"

#define FIELD(S, f) printf("  \"%s.%s\": [%zu, %zu],\n", #S, #f, offsetof(S, f), sizeof(((S *)0)->f))

static int layout(void) {
    printf("{\n");
    printf("  \"sizeof.SprPoint\": %zu,\n  \"sizeof.SprSim\": %zu,\n  \"sizeof.SprRun\": %zu,\n  \"sizeof.SprOut\": %zu,\n"
           "  \"sizeof.SprGrid\": %zu,\n  \"sizeof.SprAligned\": %zu,\n",
           sizeof(SprPoint), sizeof(SprSim), sizeof(SprRun), sizeof(SprOut), sizeof(SprGrid), sizeof(SprAligned));
    FIELD(SprPoint, kon_eff); FIELD(SprPoint, koff); FIELD(SprPoint, konb); FIELD(SprPoint, reach);
    FIELD(SprSim, N); FIELD(SprSim, DIM); FIELD(SprSim, L); FIELD(SprSim, tstop); FIELD(SprSim, tstop_AtoB);
    FIELD(SprSim, T); FIELD(SprSim, tsave); FIELD(SprSim, resample_initlocs); FIELD(SprSim, initlocs);
    FIELD(SprRun, term_kind); FIELD(SprRun, nsims); FIELD(SprRun, ssetol); FIELD(SprRun, minsims);
    FIELD(SprRun, maxsims); FIELD(SprRun, observable); FIELD(SprRun, replicate_base); FIELD(SprRun, seed);
    FIELD(SprRun, param_index_base);
    FIELD(SprOut, sum); FIELD(SprOut, sumsq); FIELD(SprOut, n_used); FIELD(SprOut, n_events); FIELD(SprOut, mean);
    FIELD(SprGrid, lo); FIELD(SprGrid, hi); FIELD(SprGrid, n);
    FIELD(SprAligned, nconc); FIELD(SprAligned, npts); FIELD(SprAligned, times); FIELD(SprAligned, values);
    FIELD(SprAligned, antibodyconcens);
    printf("  \"abi\": %d\n}\n", SPRB200_ABI_VERSION);
    return 0;
}

#define NT 41

static int run(const char *path) {
    sprb200_handle h = NULL;
    int rc = sprb200_create(0, &h);
    if (rc != SPRB200_OK) {
        fprintf(stderr, "sprb200_create: %d %s\n", rc, sprb200_last_error(NULL));
        return 2;
    }
    double tsave[NT];
    for (int s = 0; s < NT; ++s) tsave[s] = 2.0 * s;
    SprSim sim;
    memset(&sim, 0, sizeof(sim));
    sim.N = 300; sim.DIM = 3; sim.L = sprb200_domain_length(300, 3, 500.0 / 149.0 / 26795.0 * 1000000.0, 1);
    sim.tstop = 80.0; sim.tstop_AtoB = 30.0; sim.T = NT; sim.tsave = tsave; sim.resample_initlocs = 1; sim.initlocs = NULL;
    SprPoint pts[2] = {{0.05, 0.05, 0.8, 25.0}, {1.0, 0.2, 3.0, 40.0}};
    SprRun r;
    memset(&r, 0, sizeof(r));
    r.term_kind = SPRB200_TERM_FIXED; r.nsims = 50; r.observable = SPRB200_OBS_TOTAL_BOUND; r.seed = 12345u; r.param_index_base = 7u;
    int64_t sum[2 * NT];
    uint64_t nev[2];
    SprOut out;
    memset(&out, 0, sizeof(out));
    out.sum = sum; out.n_events = nev;
    rc = sprb200_simulate(h, &sim, pts, 2, &r, &out);
    if (rc != SPRB200_OK) { fprintf(stderr, "sprb200_simulate: %d %s\n", rc, sprb200_last_error(h)); return 3; }
    if (sum[0] != 0 || sum[NT] != 0) { fprintf(stderr, "t = 0 slot must be empty\n"); return 4; }
    SprGrid g = {{-2.0, -2.0, -1.0, 10.0}, {0.0, -1.0, 0.5, 30.0}, {2, 2, 2, 2}};
    SprRun rv = r;
    rv.term_kind = SPRB200_TERM_VARIANCE; rv.ssetol = 0.05; rv.minsims = 5; rv.maxsims = 30;
    double slice[4 * NT];
    int32_t nused[4];
    rc = sprb200_build_slice(h, &g, &sim, &rv, 5, 8, slice, nused, NULL);
    if (rc != SPRB200_OK) { fprintf(stderr, "sprb200_build_slice: %d %s\n", rc, sprb200_last_error(h)); return 5; }
    for (int k = 0; k < 4; ++k)
        if (nused[k] < 5 || nused[k] > 30 || !(slice[k * NT + NT - 1] >= 0.0 && slice[k * NT + NT - 1] <= 1.0)) return 6;
    /* error path: a NULL tsave must be refused with BAD_ARG, not crash */
    SprSim bad = sim;
    bad.tsave = NULL;
    if (sprb200_simulate(h, &bad, pts, 2, &r, &out) != SPRB200_ERR_BAD_ARG) return 7;
    FILE *f = fopen(path, "wb");
    if (!f) return 8;
    fwrite(sum, sizeof(sum), 1, f);
    fwrite(nev, sizeof(nev), 1, f);
    fwrite(slice, sizeof(slice), 1, f);
    fwrite(nused, sizeof(nused), 1, f);
    fclose(f);
    sprb200_destroy(h);
    printf("c_client ok: %llu + %llu events\n", (unsigned long long)nev[0], (unsigned long long)nev[1]);
    return 0;
}

int main(int argc, char **argv) {
    if (argc >= 2 && strcmp(argv[1], "layout") == 0) return layout();
    if (argc >= 3 && strcmp(argv[1], "run") == 0) return run(argv[2]);
    fprintf(stderr, "usage: c_client layout | run <out.bin>\n");
    return 1;
}
