// spr_device.cuh -- device-side building blocks of the sm_100a SSA trajectory kernel.
//
// Everything here is part of the *trajectory specification* (DESIGN.md section 3): the CPU mirror
// in oracle/spr_mirror.c re-implements the same integer and IEEE-754 sequences independently, so
// that GPU output can be compared bit for bit.  All floating-point steps therefore use explicit
// round-to-nearest intrinsics (__dmul_rn/__dadd_rn/__fma_rn/__ddiv_rn): nvcc never contracts or
// re-associates those.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace spr {

// ---------------------------------------------------------------- Philox4x32-10 (Salmon et al. 2011)
// key = (seed_lo, seed_hi); counter = (draw index, stream id, replicate, parameter index).
struct u32x4 { uint32_t x, y, z, w; };

__device__ __forceinline__ u32x4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return u32x4{c0, c1, c2, c3};
}

enum : uint32_t { STREAM_POSITIONS = 0u, STREAM_EVENTS = 1u };

// ---------------------------------------------------------------- deterministic -log(U)
// U = k * 2^-53, k in [1, 2^53].  Only +,*,/,fma on doubles and integer bit operations, in a fixed
// order: bit-identical on any IEEE-754 machine (the mirror runs the same sequence with C fma()).
// log(m) = 2 atanh(s), s = (m-1)/(m+1), m in [sqrt(1/2), sqrt(2)], series to s^23 (|err| < 1e-18).
__device__ __forceinline__ double neg_log_u53(uint64_t k) {
    const double d = __ull2double_rn(k);                  // exact, k <= 2^53
    const long long bits = __double_as_longlong(d);
    int e = (int)((bits >> 52) & 0x7ff) - 1023;
    double m = __longlong_as_double((bits & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);
    if (m > 1.4142135623730951) { m = __dmul_rn(m, 0.5); e += 1; }
    const double s = __ddiv_rn(__dadd_rn(m, -1.0), __dadd_rn(m, 1.0));
    const double z = __dmul_rn(s, s);
    double p = 1.0 / 23.0;
    p = __fma_rn(p, z, 1.0 / 21.0);
    p = __fma_rn(p, z, 1.0 / 19.0);
    p = __fma_rn(p, z, 1.0 / 17.0);
    p = __fma_rn(p, z, 1.0 / 15.0);
    p = __fma_rn(p, z, 1.0 / 13.0);
    p = __fma_rn(p, z, 1.0 / 11.0);
    p = __fma_rn(p, z, 1.0 / 9.0);
    p = __fma_rn(p, z, 1.0 / 7.0);
    p = __fma_rn(p, z, 1.0 / 5.0);
    p = __fma_rn(p, z, 1.0 / 3.0);
    const double t2 = __dadd_rn(s, s);
    const double logm = __fma_rn(__dmul_rn(t2, z), p, t2);
    return __fma_rn((double)(53 - e), 0.6931471805599453, -logm);
}

// ---------------------------------------------------------------- waiting time dt = E / a0
// dt = E * RN(1 / a0): __drcp_rn (MUFU.RCP64H seed + 5 FMAs, correctly rounded, special cases in a
// subroutine that is never reached for normal a0) and one multiply -- two IEEE operations, which the mirror
// writes as E * (1.0 / a0).  a0 == 0 (nothing can fire) gives +inf.  Three dependent FMAs and five
// instructions shorter than the exponent-flip seed + four Newton steps it replaced (specification v3).

// ---------------------------------------------------------------- lattice positions
// 21 bits per axis (top bits of one Philox word each), packed k0 | k1 << 21 | k2 << 42;
// x_d = k_d * L * 2^-21 in [0, L).
constexpr int POS_BITS = 21;
constexpr uint32_t POS_MASK = (1u << POS_BITS) - 1u;

// ---------------------------------------------------------------- state word of one particle
// sn = state | nA << 2 | nB << 17 ; state codes follow the reference (forward_simulator.jl:123):
// 1 = A (free antigen), 2 = B (singly bound), 0 = C (member of a doubly bound pair).
// nA / nB = number of neighbours (within reach) currently in state A / B.
constexpr uint32_t ST_C = 0u, ST_A = 1u, ST_B = 2u;
constexpr int SH_NA = 2, SH_NB = 17;
constexpr uint32_t CNT_MASK = 0x7fffu;
__device__ __forceinline__ uint32_t sn_state(uint32_t v) { return v & 3u; }
__device__ __forceinline__ int sn_nA(uint32_t v) { return (int)((v >> SH_NA) & CNT_MASK); }
__device__ __forceinline__ int sn_nB(uint32_t v) { return (int)(v >> SH_NB); }

// ---------------------------------------------------------------- cell grid (shared with host)
// nc cells per axis: min(floor(L/reach), cmax) where cmax is the largest c with c^DIM <= min(N,1024);
// fewer than 3 cells per axis -> 1 (all-pairs).  Cell of lattice coordinate k: (k*nc) >> 21.
__host__ __device__ inline int cells_per_axis(int N, int DIM, double L, double reach) {
    int cmax = 1;
    const long long lim = N < 1024 ? (N < 1 ? 1 : N) : 1024;
    while (true) {
        long long c = cmax + 1, v = 1;
        for (int d = 0; d < DIM; ++d) v *= c;
        if (v > lim) break;
        ++cmax;
    }
    int nc = cmax;
    if (reach > 0.0) {
        const double q = L / reach;            // IEEE division
        if (q < (double)cmax) nc = (int)q;     // floor for q >= 0
    }
    if (nc < 3) nc = 1;
    return nc;
}

// warp inclusive scan over the first `nact` lanes' values (others must pass 0)
__device__ __forceinline__ int warp_incl_scan(int v, int lane, int nact) {
    for (int d = 1; d < nact; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

}  // namespace spr
