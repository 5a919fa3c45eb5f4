// host emulation of K2's particle_row (extracted from spr_kernels.cu) against an O(N^2) scan
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>
#include <random>
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
struct uint4 { uint32_t x, y, z, w; };
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
using std::min;
#include "extract.inc"

static int cells_per_axis(int N, int DIM, double L, double reach) {
    int cmax = 1; const long long lim = N < 1024 ? (N < 1 ? 1 : N) : 1024;
    while (true) { long long c = cmax + 1, v = 1; for (int d = 0; d < DIM; ++d) v *= c; if (v > lim) break; ++cmax; }
    int nc = cmax;
    if (reach > 0.0) { const double q = L / reach; if (q < (double)cmax) nc = (int)q; }
    if (nc < 3) nc = 1;
    return nc;
}

template <int DIM> int run(int N, double L, double reach, unsigned seed, int capv) {
    std::mt19937 rng(seed);
    const int nc = cells_per_axis(N, DIM, L, reach);
    int ncell = 1; for (int d = 0; d < DIM; ++d) ncell *= nc;
    std::vector<uint32_t> k(3 * N);
    for (int i = 0; i < N; ++i) { k[3*i] = rng() >> 11; k[3*i+1] = rng() >> 11; k[3*i+2] = DIM == 3 ? rng() >> 11 : 0; }
    // a few coincident / boundary positions
    if (N > 4) { k[3] = k[0]; k[4] = k[1]; k[5] = k[2]; k[6] = 0; k[7] = 0; k[8] = 0; k[9] = (1u<<21)-1; k[10] = (1u<<21)-1; k[11] = DIM==3 ? (1u<<21)-1 : 0; }
    auto cellof = [&](int i) { uint32_t cx = (k[3*i]*(uint32_t)nc) >> 21, cy = (k[3*i+1]*(uint32_t)nc) >> 21, cz = DIM==3 ? (k[3*i+2]*(uint32_t)nc) >> 21 : 0; return nc >= 3 ? (int)((cz*nc+cy)*nc+cx) : 0; };
    std::vector<int> order(N); for (int i = 0; i < N; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cellof(a) < cellof(b); });
    std::vector<uint4> p(N);
    std::vector<uint32_t> cs(ncell + 3, 0);
    for (int s = 0; s < N; ++s) { int i = order[s]; p[s] = uint4{k[3*i] << 11, k[3*i+1] << 11, k[3*i+2] << 11, (uint32_t)i}; cs[cellof(i) + 2]++; }
    for (int c = 0; c < ncell; ++c) cs[c + 2] += cs[c + 1];   // cs[c+1] = first sorted index of cell c
    K2Smem s{}; s.p = p.data(); s.cs = cs.data();
    const double h = L * 0x1.0p-21; const double q = (reach * reach) / (h * h);
    unsigned long long r2lat = !(q < 4611686018427387904.0) ? 4611686018427387904ULL : (unsigned long long)q;
    const unsigned long long r2s = r2lat >= (3ULL << 40) ? (3ULL << 62) : (r2lat << 22);
    const double rq = sqrt((double)r2s) + 2.0;
    const uint32_t rmax = rq >= 4294967295.0 ? 0xffffffffu : (uint32_t)rq;
    uint32_t bnd[40];
    for (int c = 0; c <= nc && nc >= 3; ++c) bnd[c] = (uint32_t)((((unsigned long long)c << 32) + (unsigned long long)(nc - 1)) / (unsigned long long)nc);
    int bad = 0; long long tot = 0;
    std::vector<uint16_t> stage((size_t)N * capv + 64, 0xdead);
    for (int i = 0; i < N; ++i) {
        // brute force in sorted numbering
        std::vector<int> ref;
        for (int j = 0; j < N; ++j) {
            if (j == i) continue;
            unsigned long long acc = 0;
            const uint32_t a3[3] = {p[i].x, p[i].y, p[i].z}, b3[3] = {p[j].x, p[j].y, p[j].z};
            for (int d = 0; d < DIM; ++d) { uint32_t dd = b3[d] - a3[d]; uint32_t m = (int32_t)dd < 0 ? 0u - dd : dd; acc += (unsigned long long)m * m; }
            if (acc <= r2s) ref.push_back(j);
        }
        RowPacker none{nullptr, 0u, 0, 1, 16};
        uint16_t *row = stage.data() + (size_t)i * capv;
        const int c = particle_row<DIM, 0>(i, N, nc, s, r2s, rmax, bnd, row, capv, none);
        std::vector<int> got;
        for (int kk = 0; kk < c && kk < capv; ++kk) got.push_back(row[kk]);
        if (c != (int)ref.size()) { ++bad; if (bad < 5) printf("count mismatch i=%d got %d ref %zu\n", i, c, ref.size()); continue; }
        if (c <= capv) {
            std::vector<int> g2 = got; std::sort(g2.begin(), g2.end());
            if (g2 != ref) { ++bad; if (bad < 5) printf("set mismatch i=%d\n", i); }
            if (std::find(got.begin(), got.end(), i) != got.end()) { ++bad; if (bad < 5) printf("self listed i=%d\n", i); }
            // order: printed for the comparison between the skip levels
            for (int j : got) printf("E %d %d\n", i, j);
        }
        // MODE 1: packed words
        std::vector<uint32_t> w(N + 8, 0);
        RowPacker pk{w.data(), 0u, 0, 3, 10};
        const int c1 = particle_row<DIM, 1>(i, N, nc, s, r2s, rmax, bnd, nullptr, 0, pk);
        pk.flush();
        if (c1 != (int)ref.size()) { ++bad; if (bad < 5) printf("mode1 count mismatch i=%d %d %zu\n", i, c1, ref.size()); }
        if (N <= 1023) {
            std::vector<int> g1;
            for (int e = 0; e < c1; ++e) g1.push_back((w[e / 3] >> (10 * (e % 3))) & 1023);
            if (c <= capv && g1 != got) { ++bad; if (bad < 5) printf("mode1 order mismatch i=%d\n", i); }
        }
        tot += c;
    }
    fprintf(stderr, "DIM %d N %d L %g reach %g nc %d: mean degree %.2f, bad %d\n", DIM, N, L, reach, nc, (double)tot / N, bad);
    return bad;
}

int main() {
    int bad = 0;
    const double L = 236.7;
    for (double reach : {0.0, 2.0, 9.33, 16.7, 23.67, 24.0, 29.6, 35.0, 39.45, 60.0, 78.9, 79.0, 120.0, 500.0})
        for (unsigned seed = 1; seed <= 2; ++seed) bad += run<3>(1000, L, reach, seed, 64);
    for (double reach : {5.0, 35.0, 78.0}) bad += run<3>(300, L, reach, 7, 1000);
    for (double reach : {2.0, 10.0, 35.0, 70.0, 160.0, 400.0}) bad += run<2>(1000, 1500.0, reach, 3, 64);
    for (double reach : {10.0, 35.0}) bad += run<2>(60, 300.0, reach, 4, 64);
    for (double reach : {10.0, 35.0}) bad += run<3>(2, 100.0, reach, 5, 8);
    bad += run<3>(1, 100.0, 10.0, 6, 8);
    fprintf(stderr, "TOTAL bad %d\n", bad);
    return bad != 0;
}
