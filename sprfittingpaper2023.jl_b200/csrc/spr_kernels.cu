// spr_kernels.cu -- hand-written sm_100a kernels for the SSA ensemble (DESIGN.md sections 3-5).
//
// K2  spr_table_kernel     neighbour table within `reach` on the periodic box from a cell list: one
//                          256-thread CTA per trajectory (one thread per particle), persistent CTAs
//                          pulling tickets from a global counter, exactly-sized records bump-allocated
//                          in an HBM arena.  Replaces setup_spr_sim!
//                          (/root/reference/src/forward_simulator.jl:37-89).
// K1  spr_traj_kernel      one trajectory per warp, up to 25 warps resident per SM (9 bytes of shared
//                          memory per particle), persistent, same ticket scheme.  Replaces the
//                          replicate loop + event loop of run_spr_sim! (forward_simulator.jl:129-361).
// K2b fixed_graph_*        the table for user-supplied FP64 positions (resample_initlocs = false),
//                          built once per call with the reference's periodic_dist_sq (:1-8).
// K3  reduce_fixed / variance_scan   replicate reduction and the VarianceTerminator's sequential
//                          stopping rule (/root/reference/src/callbacks.jl:23-33, 204-217).
// K4  means / scatter      mean curves in slice order and the [P][T] -> [T][P] permutation into the
//                          "FirstMoment" layout (/root/reference/src/surrogate.jl:232, 292, 315).
//
// Nothing here is a dense contraction: no tensor cores.  The trajectory is a dependent chain of
// shared-memory reads/writes, shuffles and ALU ops; throughput comes from keeping as many
// trajectories resident per SM as shared memory and registers allow.
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

// r2lat of the specification: floor(reach^2 / scale^2) with scale = L 2^-21, clamped to 2^62
__host__ __device__ inline unsigned long long reach2_lattice(double L, double reach) {
    const double scale = L * 0x1.0p-21;
    const double q = (reach * reach) / (scale * scale);   // separate IEEE operations (no contraction)
    if (!(q < 4611686018427387904.0)) return 4611686018427387904ULL;
    return (unsigned long long)q;                        // truncation == floor for q >= 0
}

// ------------------------------------------------------------------------------------------------
// K2: neighbour table of one trajectory with freshly sampled positions.
//   positions: particle i (0..N-1) gets lattice coordinates k_d = top 21 bits of word d of
//              Philox(i, STREAM_POSITIONS, rep, pidx), x_d = k_d L 2^-21 in [0, L)
//              (forward_simulator.jl:42-47: L*rand())
//   numbering: particles are renumbered by (cell id, i) -- a stable counting sort
//   table:     neighbours of particle i in (dz,dy,dx) lexicographic order of the cell offsets,
//              ascending sorted index inside a cell  (forward_simulator.jl:52-61, 71-86)
// Coordinates are kept pre-shifted (k << 11): the difference of two of them wraps at 2^32 exactly as the
// lattice wraps at 2^21, so the min-image distance per axis is |int32(xj - xi)| = 2^11 min(|dk|, 2^21-|dk|)
// -- the reference's min(|dx|, L-|dx|) (forward_simulator.jl:1-8) evaluated without rounding on the
// lattice; the sum of squares (< 3 2^62) is compared with 2^22 floor(reach^2 / (L 2^-21)^2) in 64-bit integers.
//
// One 256-thread CTA per trajectory, one thread per particle:
//   1. positions + arrival rank in the cell (shared-memory atomics), 2. cell starts (warp scan),
//   3. stable placement (rank among the cell's particles by original index), 4. ONE distance pass that
//   stages each particle's hits in a fixed-stride row of a per-CTA global scratch (L2-resident),
//   5. degree scan + exactly-sized record from the arena, 6. compaction of the staged rows into the
//   record.  A particle with more neighbours than the staging stride falls back to a second distance pass.
// ------------------------------------------------------------------------------------------------
#ifndef SPRB200_K2_MIN_CTAS
#define SPRB200_K2_MIN_CTAS 6
#endif
#ifndef SPRB200_K2_SKIP
#define SPRB200_K2_SKIP 1          // 0: visit all 3^DIM neighbouring cells (same rows, more candidates)
#endif
#ifndef SPRB200_K2_UNROLL2
#define SPRB200_K2_UNROLL2 0       // 1: two candidates per iteration of the distance loop (measured 2 % slower at 40 registers)
#endif
#ifndef SPRB200_K2_SIGNED
#define SPRB200_K2_SIGNED 1        // 0: |d| and an unsigned square (three more instructions per candidate)
#endif
constexpr int K2_MIN_CTAS = SPRB200_K2_MIN_CTAS;

struct K2Smem {
    uint4 *p;                 // [N] cell-sorted pre-shifted lattice coordinates (x, y, z, original index)
    uint64_t *tmp;            // [N] packed coordinates in generation order (dead after the sort) ...
    uint32_t *off;            // ... [N+1] degrees, then CSR offsets (same memory)
    uint32_t *cs;             // [ncell+2]: cs[c+1] = first sorted index of cell c, cs[ncell+1] = N
    uint16_t *seg;            // [N] original indices grouped by cell in arrival order
    uint16_t *rk;             // [N] arrival rank of particle i in its cell
};

__host__ __device__ inline int k2_align16(int x) { return (x + 15) & ~15; }
__host__ __device__ inline int k2_o_tmp(int N) { return 16 * N; }
__host__ __device__ inline int k2_o_cs(int N) { return k2_o_tmp(N) + k2_align16(8 * N > 4 * (N + 1) ? 8 * N : 4 * (N + 1)); }
__host__ __device__ inline int k2_o_seg(int N) { (void)N; return k2_o_cs(N) + k2_align16(4 * (1024 + 3)); }
__host__ __device__ inline int k2_o_rk(int N) { return k2_o_seg(N) + k2_align16(2 * N); }
__host__ __device__ inline int k2_bytes(int N) { return k2_o_rk(N) + k2_align16(2 * N); }

// staging stride (entries per particle) for mean degree dbar: mean + 6 sigma, multiple of 8
__host__ __device__ inline int k2_stage_cap(int N, int DIM, double L, double reach) {
    const double PI = 3.14159265358979323846;
    double frac = DIM == 3 ? (4.0 / 3.0) * PI * reach * reach * reach / (L * L * L) : PI * reach * reach / (L * L);
    if (frac > 1.0) frac = 1.0;
    const double dbar = (double)(N - 1) * frac;
    double c = dbar + 6.0 * sqrt(dbar) + 8.0;
    if (c > (double)(N - 1)) c = (double)(N - 1);
    int cap = ((int)c + 7) & ~7;
    return cap < 8 ? 8 : cap;
}

// Packed neighbour rows (the record K1 reads): EPW entries of EB bits per 32-bit word, rows start on a word,
// the last word of a row is padded with the all-ones entry EMPTY.  EB = 10 (3 per word) when N <= 1023,
// else 16 (2 per word).
struct RowPacker {
    uint32_t *w;        // next word of the row
    uint32_t acc;       // word under construction
    int fill, epw, eb;  // entries in acc
    __device__ __forceinline__ void push(uint32_t j) {
        acc |= j << (fill * eb);
        if (++fill == epw) {
            *w++ = acc;
            acc = 0u;
            fill = 0;
        }
    }
    __device__ __forceinline__ void flush() {
        if (fill) {
            const uint32_t empty = (1u << eb) - 1u;
            for (; fill < epw; ++fill) acc |= empty << (fill * eb);
            *w++ = acc;
        }
    }
};

// squared min-image distance on the pre-shifted lattice: the 32-bit wrap-around difference read as a signed number
// IS the periodic image, and its signed 64-bit square the exact square (|d| <= 2^31)
template <int DIM>
__device__ __forceinline__ unsigned long long lattice_dist_sq(const uint4 c, uint32_t xi, uint32_t yi, uint32_t zi) {
#if SPRB200_K2_SIGNED
    const int32_t dx = (int32_t)(c.x - xi), dy = (int32_t)(c.y - yi);
    long long acc = (long long)dx * dx + (long long)dy * dy;
    if (DIM == 3) {
        const int32_t dz = (int32_t)(c.z - zi);
        acc += (long long)dz * dz;
    }
    return (unsigned long long)acc;
#else
    uint32_t d = c.x - xi;
    uint32_t m = (int32_t)d < 0 ? 0u - d : d;
    unsigned long long acc = (unsigned long long)m * m;
    d = c.y - yi;
    m = (int32_t)d < 0 ? 0u - d : d;
    acc += (unsigned long long)m * m;
    if (DIM == 3) {
        d = c.z - zi;
        m = (int32_t)d < 0 ? 0u - d : d;
        acc += (unsigned long long)m * m;
    }
    return acc;
#endif
}

// MODE 0: count + stage (hits beyond `cap` are only counted); MODE 1: pack every hit into the record
template <int MODE>
__device__ __forceinline__ void take_hit(int j, uint16_t *row, int cap, RowPacker &pk, int &cnt) {
    if (MODE == 1) pk.push((uint32_t)j);
    else if (cnt < cap) row[cnt] = (uint16_t)j;
    ++cnt;
}

template <int DIM, int MODE>
__device__ __forceinline__ int scan_range(int i, uint32_t xi, uint32_t yi, uint32_t zi, int j0, int j1,
                                          const uint4 *__restrict__ p, unsigned long long r2s, uint16_t *row, int cap,
                                          RowPacker &pk, int cnt) {
    int j = j0;
#if SPRB200_K2_UNROLL2
    // two candidates per iteration: both loads and both distance chains are in flight together
#pragma unroll 1
    for (; j + 1 < j1; j += 2) {
        const uint4 c0 = p[j], c1 = p[j + 1];
        const unsigned long long a0 = lattice_dist_sq<DIM>(c0, xi, yi, zi), a1 = lattice_dist_sq<DIM>(c1, xi, yi, zi);
        if (a0 <= r2s && j != i) take_hit<MODE>(j, row, cap, pk, cnt);
        if (a1 <= r2s && j + 1 != i) take_hit<MODE>(j + 1, row, cap, pk, cnt);
    }
    if (j < j1) {
        if (lattice_dist_sq<DIM>(p[j], xi, yi, zi) <= r2s && j != i) take_hit<MODE>(j, row, cap, pk, cnt);
    }
#else
#pragma unroll 1
    for (; j < j1; ++j) {
        if (lattice_dist_sq<DIM>(p[j], xi, yi, zi) <= r2s && j != i) take_hit<MODE>(j, row, cap, pk, cnt);
    }
#endif
    return cnt;
}

// all within-reach neighbours of sorted particle i, in the order of the specification.
// A neighbouring cell along an axis is only visited when the particle is within reach of the face it shares with
// that cell (`bnd[c]` = first pre-shifted coordinate of cell c, `rmax` >= floor(sqrt(r2s))): every particle of a
// skipped cell is farther than the reach along that axis alone, so the rows -- and their order -- are unchanged.
template <int DIM, int MODE>
__device__ __forceinline__ int particle_row(int i, int N, int nc, const K2Smem &s, unsigned long long r2s,
                                            uint32_t rmax, const uint32_t *bnd, uint16_t *row, int cap, RowPacker &pk) {
    const uint4 me = s.p[i];
    const uint32_t xi = me.x, yi = me.y, zi = DIM == 3 ? me.z : 0u;
    if (nc < 3) return scan_range<DIM, MODE>(i, xi, yi, zi, 0, N, s.p, r2s, row, cap, pk, 0);
    const uint32_t *cs1 = s.cs + 1;                     // cs1[c] = first sorted index of cell c
    const int cx = (int)__umulhi(xi, (uint32_t)nc);     // (k << 11) * nc >> 32 == (k * nc) >> 21
    const int cy = (int)__umulhi(yi, (uint32_t)nc);
    const int cz = DIM == 3 ? (int)__umulhi(zi, (uint32_t)nc) : 0;
#if SPRB200_K2_SKIP
    const int lox = xi - bnd[cx] <= rmax ? 1 : 0, hix = bnd[cx + 1] - 1u - xi <= rmax ? 1 : 0;
    const int loy = yi - bnd[cy] <= rmax ? 1 : 0, hiy = bnd[cy + 1] - 1u - yi <= rmax ? 1 : 0;
    const int loz = DIM == 3 ? (zi - bnd[cz] <= rmax ? 1 : 0) : 0, hiz = DIM == 3 ? (bnd[cz + 1] - 1u - zi <= rmax ? 1 : 0) : 0;
#else
    const int lox = 1, hix = 1, loy = 1, hiy = 1, loz = DIM == 3 ? 1 : 0, hiz = loz;
#endif
    // the x-cells of a (dz,dy) row are contiguous in the sorted numbering except across the wrap
    int ax1, bx1, ax2 = 0, bx2 = 0;
    if (cx == 0) {
        if (lox) {
            ax1 = nc - 1; bx1 = nc; bx2 = 1 + hix;
        } else {
            ax1 = 0; bx1 = 1 + hix;
        }
    } else if (cx == nc - 1) {
        ax1 = cx - lox; bx1 = nc; bx2 = hix;
    } else {
        ax1 = cx - lox; bx1 = cx + 1 + hix;
    }
    const int ny = 1 + loy + hiy, nrows = ny * (1 + loz + hiz);
    const int yy0 = cy - loy;
    int zz = cz - loz;
    zz = zz < 0 ? zz + nc : zz;
    int iy = 0, cnt = 0;
#pragma unroll 1
    for (int r = 0; r < nrows; ++r) {                   // (dz, dy) ascending, like the full 3 x 3 sweep
        int yy = yy0 + iy;
        yy = yy < 0 ? yy + nc : (yy >= nc ? yy - nc : yy);
        const int rb = (zz * nc + yy) * nc;
        cnt = scan_range<DIM, MODE>(i, xi, yi, zi, (int)cs1[rb + ax1], (int)cs1[rb + bx1], s.p, r2s, row, cap, pk, cnt);
        if (bx2 > ax2)
            cnt = scan_range<DIM, MODE>(i, xi, yi, zi, (int)cs1[rb + ax2], (int)cs1[rb + bx2], s.p, r2s, row, cap, pk, cnt);
        if (++iy == ny) {
            iy = 0;
            zz = zz + 1 >= nc ? 0 : zz + 1;
        }
    }
    return cnt;
}

template <int DIM>
__device__ __forceinline__ uint32_t k2_cell(uint32_t kx, uint32_t ky, uint32_t kz, uint32_t nc) {
    const uint32_t cx = (kx * nc) >> POS_BITS, cy = (ky * nc) >> POS_BITS;
    const uint32_t cz = DIM == 3 ? (kz * nc) >> POS_BITS : 0u;
    return (cz * nc + cy) * nc + cx;
}

template <int DIM>
__global__ void __launch_bounds__(K2_THREADS, K2_MIN_CTAS) spr_table_kernel(const TableArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ long long s_ticket;
    __shared__ unsigned long long s_base;
    __shared__ unsigned int s_maxdeg;
    __shared__ uint32_t s_bnd[40];             // [nc+1] first pre-shifted coordinate of every cell along an axis (nc <= 32)
    const int N = a.N, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    K2Smem s;
    s.p = reinterpret_cast<uint4 *>(smem_raw);
    s.tmp = reinterpret_cast<uint64_t *>(smem_raw + k2_o_tmp(N));
    s.off = reinterpret_cast<uint32_t *>(s.tmp);
    s.cs = reinterpret_cast<uint32_t *>(smem_raw + k2_o_cs(N));
    s.seg = reinterpret_cast<uint16_t *>(smem_raw + k2_o_seg(N));
    s.rk = reinterpret_cast<uint16_t *>(smem_raw + k2_o_rk(N));
    uint16_t *stage = a.stage + (size_t)blockIdx.x * (size_t)N * (size_t)a.stage_cap;

    for (;;) {
        __syncthreads();                          // everyone is done with the shared scalars of the previous ticket
        if (tid == 0) {
            s_ticket = (long long)atomicAdd(a.ctr + CTR_K2_TICKET, 1ULL);
            s_maxdeg = 0u;
        }
        __syncthreads();
        const long long q = s_ticket;
        if (q >= a.ntickets) break;
        const long long tk = a.ticket_list ? a.ticket_list[q] : a.tk0 + q;
        const int slot = a.order[tk / a.nrep];
        const uint32_t rep = (uint32_t)(a.rep0 + (int)(tk % a.nrep));
        const double reach = a.pts[slot].reach;
        const uint32_t pidx = a.pidx[slot];
        const int nc = cells_per_axis(N, DIM, a.L, reach);
        int ncell = nc;
#pragma unroll
        for (int d = 1; d < DIM; ++d) ncell *= nc;
        // 2^22 r2lat, saturated at the largest possible sum of squares
        const unsigned long long r2lat = reach2_lattice(a.L, reach);
        const unsigned long long r2s = r2lat >= (3ULL << 40) ? (3ULL << 62) : (r2lat << 22);
        const int cap = min(a.stage_cap, k2_stage_cap(N, DIM, a.L, reach));
        // >= floor(sqrt(r2s)) (the double square root is within 1 of it), saturated: no cell is skipped then
        const double rq = sqrt((double)r2s) + 2.0;
        const uint32_t rmax = rq >= 4294967295.0 ? 0xffffffffu : (uint32_t)rq;

        // ---- 1. positions, cell counts (kept at cs[c+2]) and arrival ranks
        for (int c = tid; c < ncell + 2; c += K2_THREADS) s.cs[c] = 0u;
        // ceil(c 2^32 / nc) mod 2^32: cell c holds the coordinates [bnd[c], bnd[c+1]) (bnd[nc] wraps to 0)
        if (tid <= nc && nc >= 3) s_bnd[tid] = (uint32_t)((((unsigned long long)tid << 32) + (unsigned long long)(nc - 1)) / (unsigned long long)nc);
        __syncthreads();
        for (int i = tid; i < N; i += K2_THREADS) {
            const u32x4 r = philox4x32_10((uint32_t)i, STREAM_POSITIONS, rep, pidx, a.seed_lo, a.seed_hi);
            const uint32_t kx = r.x >> (32 - POS_BITS), ky = r.y >> (32 - POS_BITS);
            const uint32_t kz = DIM == 3 ? r.z >> (32 - POS_BITS) : 0u;
            s.tmp[i] = (uint64_t)kx | ((uint64_t)ky << POS_BITS) | ((uint64_t)kz << (2 * POS_BITS));
            if (nc >= 3) s.rk[i] = (uint16_t)atomicAdd(&s.cs[k2_cell<DIM>(kx, ky, kz, (uint32_t)nc) + 2], 1u);
        }
        __syncthreads();
        // ---- 2. cell starts: cs[c+1] = first sorted index of cell c (inclusive scan of the counts one slot up)
        if (warp == 0 && nc >= 3) {
            const int per = (ncell + 31) >> 5;
            const int b = lane * per, e = min(b + per, ncell);
            int sum = 0;
            for (int c = b; c < e; ++c) sum += (int)s.cs[c + 2];
            int run = warp_incl_scan(sum, lane, 32) - sum;
            for (int c = b; c < e; ++c) {
                run += (int)s.cs[c + 2];
                s.cs[c + 2] = (uint32_t)run;
            }
        }
        __syncthreads();
        // ---- 3. stable placement: group original indices by cell, then rank inside the cell by original index
        if (nc >= 3) {
            for (int i = tid; i < N; i += K2_THREADS) {
                const uint64_t pk = s.tmp[i];
                const uint32_t c = k2_cell<DIM>((uint32_t)pk & POS_MASK, (uint32_t)(pk >> POS_BITS) & POS_MASK,
                                                (uint32_t)(pk >> (2 * POS_BITS)) & POS_MASK, (uint32_t)nc);
                s.seg[s.cs[c + 1] + s.rk[i]] = (uint16_t)i;
            }
            __syncthreads();
        }
        for (int i = tid; i < N; i += K2_THREADS) {
            const uint64_t pk = s.tmp[i];
            const uint32_t kx = (uint32_t)pk & POS_MASK, ky = (uint32_t)(pk >> POS_BITS) & POS_MASK;
            const uint32_t kz = (uint32_t)(pk >> (2 * POS_BITS)) & POS_MASK;
            int dst = i;
            if (nc >= 3) {
                const uint32_t c = k2_cell<DIM>(kx, ky, kz, (uint32_t)nc);
                const int st = (int)s.cs[c + 1], en = (int)s.cs[c + 2];
                dst = st;
                for (int w = st; w < en; ++w) dst += (int)s.seg[w] < i ? 1 : 0;
            }
            s.p[dst] = make_uint4(kx << 11, ky << 11, kz << 11, (uint32_t)i);
        }
        __syncthreads();
        // ---- 4. the distance pass (tmp is dead: off aliases it)
        unsigned int mydeg = 0u;
        for (int i = tid; i < N; i += K2_THREADS) {
            RowPacker none{nullptr, 0u, 0, 1, 16};
            const int c = particle_row<DIM, 0>(i, N, nc, s, r2s, rmax, s_bnd, stage + (size_t)i * (size_t)a.stage_cap, cap, none);
            s.seg[i] = (uint16_t)c;                       // degree (seg is dead after the sort)
            s.off[i + 1] = (uint32_t)((c + a.epw - 1) / a.epw);   // words of the packed row
            mydeg = max(mydeg, (unsigned int)c);
        }
        mydeg = __reduce_max_sync(FULL, mydeg);
        if (lane == 0) atomicMax(&s_maxdeg, mydeg);
        __syncthreads();
        // ---- 5. offsets and the record
        if (warp == 0) {
            // inclusive scan of the row lengths (words) in place: off[0] = 0, off[i+1] = words of particles 0..i
            const int per = (N + 31) >> 5;
            const int b = lane * per, e = min(b + per, N);
            unsigned int sum = 0;
            for (int i = b; i < e; ++i) sum += s.off[i + 1];
            unsigned int run = (unsigned int)warp_incl_scan((int)sum, lane, 32) - sum;
            for (int i = b; i < e; ++i) {
                run += s.off[i + 1];
                s.off[i + 1] = run;
            }
            if (lane == 0) s.off[0] = 0u;
            __syncwarp();
            if (lane == 0) {
                // exactly-sized record from the arena (bump allocation)
                const unsigned long long total = s.off[N];
                const unsigned long long bytes = (unsigned long long)(((size_t)(N + 1) * 4 + 15) & ~(size_t)15) +
                                                 ((total * 4ULL + 15ULL) & ~15ULL);
                const unsigned long long at = atomicAdd(a.ctr + CTR_CURSOR, bytes);
                if (at + bytes <= a.arena_bytes) {
                    s_base = at;
                    a.desc[q] = (unsigned long long)(uintptr_t)(a.arena + at);
                    atomicMax(a.ctr + CTR_MAX_ENTRIES, total);
                    atomicMax(a.ctr + CTR_MAX_DEGREE, (unsigned long long)s_maxdeg);
                } else {
                    s_base = ~0ULL;
                    a.desc[q] = 0ULL;
                    const unsigned long long w = atomicAdd(a.ctr + CTR_NDEFER, 1ULL);
                    a.defer_list[w] = tk;
                }
            }
        }
        __syncthreads();
        const unsigned long long at = s_base;
        if (at == ~0ULL) continue;
        // ---- 6. the record: word offsets, then the staged rows packed (or recomputed when a row overflowed its stride)
        uint32_t *goff = reinterpret_cast<uint32_t *>(a.arena + at);
        uint32_t *gw = reinterpret_cast<uint32_t *>(a.arena + at + (((size_t)(N + 1) * 4 + 15) & ~(size_t)15));
        for (int i = tid; i <= N; i += K2_THREADS) goff[i] = s.off[i];
        if ((int)s_maxdeg <= cap) {
            for (int i = tid; i < N; i += K2_THREADS) {
                const int dg = (int)s.seg[i];
                const uint16_t *src = stage + (size_t)i * (size_t)a.stage_cap;
                RowPacker pk{gw + s.off[i], 0u, 0, a.epw, a.eb};
                for (int k = 0; k < dg; ++k) pk.push((uint32_t)src[k]);
                pk.flush();
            }
        } else {
            for (int i = tid; i < N; i += K2_THREADS) {
                RowPacker pk{gw + s.off[i], 0u, 0, a.epw, a.eb};
                particle_row<DIM, 1>(i, N, nc, s, r2s, rmax, s_bnd, nullptr, 0, pk);
                pk.flush();
            }
        }
        if (a.spos_out && q == 0) {
            for (int i = tid; i < N; i += K2_THREADS) {
                const uint4 c = s.p[i];
                a.spos_out[i] = (uint64_t)(c.x >> 11) | ((uint64_t)(c.y >> 11) << POS_BITS) |
                                (DIM == 3 ? (uint64_t)(c.z >> 11) << (2 * POS_BITS) : 0ULL);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K1: the trajectory.  All scalars (time, counts, list lengths) are warp-uniform registers; the 32
// lanes share the neighbour updates, list scans, save-slot writes and random-number generation.
// State word of a particle: state | nA << 2 | nB << (2 + CB), CB = 7 (uint16_t) or 15 (uint32_t) bits
// per neighbour count; state codes follow the reference (forward_simulator.jl:123).
//
// Shared memory of one trajectory, 8 bytes per particle (TrajLayout):
//   off[N+1]  first word of every packed neighbour row        sn[N]  state words
//   ab[N]     A list from the front, B list from the back and -- inside the gap between them, which always
//             holds 2C free slots because A + B + 2C = N -- the complex list (one member per complex)
//   pos[N]    A/B: index in its list; C: the partner
// The neighbour rows themselves stay in the record in global memory (L2): EPW entries of EB bits per word.
// ------------------------------------------------------------------------------------------------
template <typename SnT> struct SnTraits;
template <> struct SnTraits<uint16_t> { static constexpr int SH_A = 2, SH_B = 9; static constexpr uint32_t MASK = 0x7fu; };
template <> struct SnTraits<uint32_t> { static constexpr int SH_A = 2, SH_B = 17; static constexpr uint32_t MASK = 0x7fffu; };

// EB = 10: three entries per word, lanes 0..29 take the entries of 10 words per pass; EB = 16: two per word, 32
// lanes take 16 words per pass.  A lane's word (lw) and shift (ls) inside a pass are constants of the lane.
// TS: the packed rows of the trajectory were copied into the warp's shared-memory slice (launches that cannot fill
// the GPU: every row load is then an LDS instead of an L2 round trip); else they are read from the record in global memory.
template <int EB_, bool TS_> struct PackTraits {
    static constexpr int EPW = EB_ == 10 ? 3 : 2;
    static constexpr int WPP = 32 / EPW;               // words per pass
    static constexpr uint32_t EMPTY = (1u << EB_) - 1u;
    static __device__ __forceinline__ uint32_t ld(const uint32_t *w, uint32_t i) { return TS_ ? w[i] : __ldg(w + i); }
};

// Neighbour updates are split into a load (row offset, length, this lane's neighbour) and an apply
// (read-modify-write of the neighbour's state word), so that the loads of both members of a pair
// event are in flight together.  When the partner of a pair event (`other`) is met, its own state
// change `extra` rides on the same store: pair events need no separate state write.
struct NbrRow {
    int o0, nw;      // first word, number of words
    uint32_t j;      // neighbour handled by this lane in the first pass, EMPTY if none
};

// table words are written by K2 (an earlier launch) and only read here
template <typename P>
__device__ __forceinline__ uint32_t ld_entry(const uint32_t *words, int o0, int nw, int wi, int ls) {
    uint32_t w = 0xffffffffu;                          // every entry EMPTY
    if (wi < nw) w = P::ld(words, (uint32_t)(o0 + wi));
    return (w >> ls) & P::EMPTY;
}

// Where the rows start.  TM = false: off[N+1] in the warp's shared-memory slice (2 or 4 B per particle).
// TM = true: the warp's 32 x 32-word block of TENSOR MEMORY holds (first word | words << 16) of particle p at
// (lane p & 31, column p >> 5): tcgen05.ld fetches the column of p (one word per lane), a shuffle broadcasts the lane
// of p -- three instructions like the two LDS + subtract of the shared-memory form, and 2 B per particle of shared
// memory become free: 6 B per particle, 32 trajectories per SM at 64 registers (N <= 1024, rows < 65,536 words).
template <bool TM, typename OffT>
struct RowIndex {
    const OffT *off;     // TM = false
    uint32_t taddr;      // TM = true: tensor-memory address of the warp's block (lane quadrant << 16 | first column)
    __device__ __forceinline__ void bounds(int m, int &o0, int &nw) const {
        if (TM) {
            uint32_t v;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr + (uint32_t)(m >> 5)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            v = __shfl_sync(FULL, v, m);          // the source lane is taken modulo 32
            o0 = (int)(v & 0xffffu);
            nw = (int)(v >> 16);
        } else {
            o0 = (int)off[m];
            nw = (int)off[m + 1] - o0;
        }
    }
};

template <bool TM, typename OffT, typename P>
__device__ __forceinline__ NbrRow load_row(const uint32_t *words, const RowIndex<TM, OffT> &ri, int m, int lw, int ls) {
    NbrRow r;
    ri.bounds(m, r.o0, r.nw);
    r.j = ld_entry<P>(words, r.o0, r.nw, lw, ls);
    return r;
}

// sn[j] += delta for every neighbour j of the row; the neighbour `other` (the partner of a pair event)
// also takes `extra`, its own state change.  apply_row covers the first pass (one neighbour per lane),
// apply_row_tail the rest (more than WPP words: rare).
template <bool PAIR, typename SnT, typename P>
__device__ __forceinline__ void apply_row(SnT *sn, const NbrRow &r, uint32_t delta, uint32_t other, uint32_t extra) {
    if (r.j != P::EMPTY) {
        uint32_t d = delta;
        if (PAIR && r.j == other) d += extra;
        sn[r.j] = (SnT)((uint32_t)sn[r.j] + d);
    }
}

template <bool PAIR, typename SnT, typename P>
__device__ __forceinline__ void apply_row_tail(SnT *sn, const uint32_t *words, const NbrRow &r, uint32_t delta,
                                               uint32_t other, uint32_t extra, int lw, int ls) {
#pragma unroll 1
    for (int wi = lw + P::WPP; wi < r.nw; wi += P::WPP) {
        const uint32_t j = ld_entry<P>(words, r.o0, r.nw, wi, ls);
        if (j != P::EMPTY) {
            uint32_t d = delta;
            if (PAIR && j == other) d += extra;
            sn[j] = (SnT)((uint32_t)sn[j] + d);
        }
    }
}

// both rows of a pair event: every phase touches each state word at most once and the phases are separated
// by __syncwarp(), so the order of the additions is immaterial; ONE uniform branch covers both tails
template <typename SnT, typename P>
__device__ __forceinline__ void apply_pair(SnT *sn, const uint32_t *words, const NbrRow &ra, const NbrRow &rb,
                                           uint32_t da, uint32_t db, int ma, int mb, uint32_t ea, uint32_t eb, int lw, int ls,
                                           bool tails) {
    apply_row<true, SnT, P>(sn, ra, da, (uint32_t)mb, ea);
    __syncwarp();
    apply_row<true, SnT, P>(sn, rb, db, (uint32_t)ma, eb);
    __syncwarp();
    if (tails && max(ra.nw, rb.nw) > P::WPP) {
        apply_row_tail<true, SnT, P>(sn, words, ra, da, (uint32_t)mb, ea, lw, ls);
        __syncwarp();
        apply_row_tail<true, SnT, P>(sn, words, rb, db, (uint32_t)ma, eb, lw, ls);
        __syncwarp();
    }
}

// inclusive scan over the first nact (<= 32) lanes; the upper levels are skipped for short lists
__device__ __forceinline__ int scan_small(int v, int lane, int nact) {
    int o = __shfl_up_sync(FULL, v, 1);
    if (lane >= 1) v += o;
    o = __shfl_up_sync(FULL, v, 2);
    if (lane >= 2) v += o;
    if (nact > 4) {
        o = __shfl_up_sync(FULL, v, 4);
        if (lane >= 4) v += o;
        o = __shfl_up_sync(FULL, v, 8);
        if (lane >= 8) v += o;
        o = __shfl_up_sync(FULL, v, 16);
        if (lane >= 16) v += o;
    }
    return v;
}

// The complex list lives in ab[cs .. cs+C), inside the gap between the A list (front) and the B list (back).
// [lo, hi) is the gap as it will be once the list appends of the current event are done; the block is moved
// (rarely: it is re-aligned to the side the gap is drifting away from) only when it would collide -- the callers
// test the one side that can collide before calling.
__device__ __noinline__ int cl_make_room(uint16_t *ab, int cs, int C, int lo, int hi, int lane) {
    if (C == 0) return lo;
    if (cs >= lo && cs + C <= hi) return cs;
    const int ns = cs < lo ? hi - C : lo;
    __syncwarp();                      // the uniform list writes of this event are visible to every lane
    if (ns < cs) {
#pragma unroll 1
        for (int c0 = 0; c0 < C; c0 += 32) {
            const int i = c0 + lane;
            uint16_t v = 0;
            if (i < C) v = ab[cs + i];
            __syncwarp();
            if (i < C) ab[ns + i] = v;
            __syncwarp();
        }
    } else {
#pragma unroll 1
        for (int c0 = ((C - 1) >> 5) << 5; c0 >= 0; c0 -= 32) {
            const int i = c0 + lane;
            uint16_t v = 0;
            if (i < C) v = ab[cs + i];
            __syncwarp();
            if (i < C) ab[ns + i] = v;
            __syncwarp();
        }
    }
    return ns;
}

// up to K1_WPC warps per CTA, 4 CTAs per SM: 28 trajectories per SM at 72 registers and 8 B of shared memory per
// particle (N = 1000); small batches run with fewer warps per CTA so that they spread over all SMs
template <typename SnT, typename OffT, int EB, bool TM, bool TS>
__global__ void __launch_bounds__(TM ? 32 * K1_WPC_TM : 32 * K1_WPC, 4) spr_traj_kernel(const TrajArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_tmem;
    using ST = SnTraits<SnT>;
    using PT = PackTraits<EB, TS>;
    constexpr int SH_A = ST::SH_A, SH_B = ST::SH_B;
    constexpr uint32_t CMASK = ST::MASK;
    constexpr uint32_t EMPTY = PT::EMPTY;
    const int lane = threadIdx.x & 31;
    // this lane's word and shift inside a pass of WPP words (lanes beyond EPW*WPP never hold an entry)
    const int lw = lane < PT::EPW * PT::WPP ? lane / PT::EPW : (1 << 28);
    const int ls = (lane % PT::EPW) * EB;
    // The warp's slice of shared memory.  The warp index goes through a shuffle so that ptxas sees a warp-UNIFORM
    // base address: every load of trajectory state at a uniform index is then uniform too, the data-dependent
    // branches below become uniform branches, and the shuffles / votes / __syncwarp inside them need no
    // convergence check (with threadIdx.x >> 5 used directly the same source compiles to 25 BRA.DIV checks, 64
    // BSSY/BSYNC pairs and 27 % more instructions).
    unsigned char *slice = smem_raw + (size_t)__shfl_sync(FULL, threadIdx.x >> 5, 0) * (size_t)a.lay.slice;
    RowIndex<TM, OffT> ri;
    ri.off = reinterpret_cast<const OffT *>(slice + a.lay.o_off);
    ri.taddr = 0u;
    if (TM) {
        // one warp allocates the CTA's tensor-memory columns (32 per group of four warps: a warp reaches only the 32
        // lanes of its quadrant, warps w and w + 4 share a quadrant and take different column blocks)
        if ((threadIdx.x >> 5) == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                         :: "r"((uint32_t)__cvta_generic_to_shared(&s_tmem)), "n"(K1_TMEM_COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;");
        const uint32_t w = (uint32_t)__shfl_sync(FULL, threadIdx.x >> 5, 0);
        ri.taddr = s_tmem + (((w & 3u) * 32u) << 16) + (w >> 2) * 32u;
    }
    OffT *off = reinterpret_cast<OffT *>(slice + a.lay.o_off);
    SnT *sn = reinterpret_cast<SnT *>(slice + a.lay.o_sn);
    uint16_t *ab = reinterpret_cast<uint16_t *>(slice + a.lay.o_ab);    // A from the front, B from the back, C in the gap
    uint16_t *pos = reinterpret_cast<uint16_t *>(slice + a.lay.o_pos);
    const int N = a.N, T = a.T;
    const int NB = N - 1;                                            // B-list element k lives at ab[NB - k]
    const double tstop = a.tstop, toff = a.toff;
    const double *__restrict__ tsave = a.tsave;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);

    for (;;) {
        long long q = 0;
        if (lane == 0) q = (long long)atomicAdd(a.ctr + CTR_K1_TICKET, 1ULL);
        q = __shfl_sync(FULL, q, 0);
        if (q >= a.ntickets) break;
        const long long tk = a.ticket_list ? a.ticket_list[q] : a.tk0 + q;
        const int slot = a.order[tk / a.nrep];
        const int rep = a.rep0 + (int)(tk % a.nrep);
        const unsigned long long rec = a.slot_rec ? a.slot_rec[slot] : a.desc[q];
        if (rec == 0ULL) continue;                                   // deferred by K2 (arena full)
        const PointDev pt = a.pts[slot];
        const uint32_t pidx = a.pidx[slot];

        // ---- neighbour table: row offsets into shared memory, packed rows stay in global memory
        const uint32_t *__restrict__ goff = reinterpret_cast<const uint32_t *>((uintptr_t)rec);
        const uint32_t *gwords =
            reinterpret_cast<const uint32_t *>((uintptr_t)rec + (((size_t)(N + 1) * 4 + 15) & ~(size_t)15));
        const uint32_t *words = gwords;
        if (TS) {
            // the launch cannot fill the GPU: the rows go into the warp's slice once (coalesced), every row load of
            // the trajectory is then a shared-memory load
            uint32_t *tw = reinterpret_cast<uint32_t *>(slice + a.lay.o_tbl);
            const int nwt = (int)goff[N];
            for (int i = lane; i < nwt; i += 32) tw[i] = __ldg(gwords + i);
            words = tw;
        }
        if (TM) {
#pragma unroll 1
            for (int c = 0; c < 32; ++c) {
                const int p = c * 32 + lane;
                uint32_t v = 0u;
                if (p < N) {
                    const uint32_t g0 = goff[p];
                    v = g0 | ((goff[p + 1] - g0) << 16);
                }
                asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" :: "r"(ri.taddr + (uint32_t)c), "r"(v));
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        } else {
            for (int i = lane; i <= N; i += 32) off[i] = (OffT)goff[i];
        }
        __syncwarp();
        // ---- initial state: everything A (forward_simulator.jl:148-149); nA = degree = valid entries of the row
        int maxnw = 0;
        for (int i = lane; i < N; i += 32) {
            const int o0 = (int)goff[i], nw = (int)goff[i + 1] - o0;
            maxnw = max(maxnw, nw);
            uint32_t deg = 0u;
            if (nw > 0) {
                // every word but the last is full; the last one holds 1..EPW entries
                const uint32_t wl = __ldg(gwords + (uint32_t)(o0 + nw - 1));
                deg = (uint32_t)(nw - 1) * PT::EPW;
#pragma unroll
                for (int e = 0; e < PT::EPW; ++e) deg += ((wl >> (e * EB)) & EMPTY) != EMPTY ? 1u : 0u;
            }
            sn[i] = (SnT)(ST_A | (deg << SH_A));
            ab[i] = (uint16_t)i;
            pos[i] = (uint16_t)i;
        }
        __syncwarp();
        // rows longer than one pass (more than 30 resp. 32 neighbours) exist in few trajectories: one uniform flag
        // instead of a degree test per event
        const bool tails = __reduce_max_sync(FULL, maxnw) > PT::WPP;

        int cntA = N, cntB = 0, cntC = 0, nAB = 0;
        int cs = 0;                                       // first slot of the complex list in ab
        double t = 0.0;
        double kon = pt.kon_eff;
        const double koff = pt.koff, konb = pt.konb;
        const double kc = __dmul_rn(2.0, koff);           // C -> A+B rate (forward_simulator.jl:135)
        bool on = true;
        int sidx = 0;
        double tp = tsave[0];
        // fast-path deadline: an event strictly before it needs no save-slot write, no switch-off and is
        // not the last one
        double tfast = fmin(fmin(tp, toff), tstop);
        uint32_t ndraw = 0, nskip = 0;                     // draws of the finished 32-blocks; draws that fired no event
        const size_t row = (size_t)slot * a.repstride + (size_t)(rep - a.rep_rowbase);
        uint16_t *crow = a.curve + row * (size_t)T;
        uint16_t *crow2 = a.curve2 ? a.curve2 + row * (size_t)T : nullptr;
        const int obs = a.observable;
        // randoms of 32 consecutive draws, one draw per lane: {E = -log U (double), w2, w3}
        uint32_t rEl = 0, rEh = 0, rz = 0, rw = 0;
        bool done = false;
        while (!done) {
            // 32 draws at a time, lane l holds draw ndraw + l (ndraw is a multiple of 32 here)
            {
                const u32x4 r = philox4x32_10(ndraw + (uint32_t)lane, STREAM_EVENTS, (uint32_t)rep, pidx,
                                              a.seed_lo, a.seed_hi);
                const uint64_t k53 = (((uint64_t)r.x << 21) | (uint64_t)(r.y >> 11)) + 1ULL;
                const double En = neg_log_u53(k53);
                rEl = (uint32_t)__double2loint(En);
                rEh = (uint32_t)__double2hiint(En);
                rz = r.z;
                rw = r.w;
            }
#pragma unroll 1
          for (int src = 0; src < 32; ++src) {
            const double E = __hiloint2double((int)__shfl_sync(FULL, rEh, src), (int)__shfl_sync(FULL, rEl, src));
            const uint32_t vz = __shfl_sync(FULL, rz, src);
            const uint32_t vw = __shfl_sync(FULL, rw, src);

            // ---- total propensity (exact integer counts x rates)
            const double cA = __dmul_rn(kon, (double)cntA);
            const double cB = __fma_rn(koff, (double)cntB, cA);
            const double cP = __fma_rn(konb, (double)nAB, cB);
            const double a0 = __fma_rn(kc, (double)cntC, cP);
            // waiting time E / a0 as E * RN(1 / a0) (IEEE reciprocal: MUFU.RCP64H seed + 5 FMAs); a0 == 0 gives
            // +inf (nothing can fire), which the slow path sorts out
            double tn = __dadd_rn(t, __dmul_rn(E, __drcp_rn(a0)));

            // (a fast-path test in the multiplied form E < (tfast - t) a0 takes the reciprocal chain off the critical
            // path but costs three instructions: measured 1 % slower, profiles/r2_k1_variants.txt)
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
                    done = true;
                    ndraw += (uint32_t)src + 1u - 32u;   // the block is left after draw src (the loop's tail adds 32)
                    src = 32;        // leaves the draw loop through its own condition (single-exit loop)
                    continue;
                }
                if (tn > tstop) {    // the first event past tstop is still applied (:173, 183) and is the last one:
                    done = true;     // it is applied below, then the draw loop ends through its own condition
                    ndraw += (uint32_t)src + 1u - 32u;
                    src = 31;
                }
                tfast = fmin(fmin(tp, on ? toff : INF), tstop);
            }
            t = tn;

            // ---- channel selection
            const double x = __dmul_rn(__dmul_rn((double)vz, 0x1.0p-32), a0);
            if (x < cB) {
                // A --> B (forward_simulator.jl:204-224) when x < cA, else B --> A (:228-253): the same
                // code with the two lists swapped (A list: ab[k], B list: ab[NB - k])
                const bool isab = x < cA;
                const int nfrom = isab ? cntA : cntB, nto = isab ? cntB : cntA;
                const int k = (int)__umulhi(vw, (uint32_t)nfrom);
                const int fk = isab ? k : NB - k;                       // slot of element k of the source list
                const int fl = isab ? nfrom - 1 : NB - (nfrom - 1);     // slot of its last element
                const int tl = isab ? NB - nto : nto;                   // slot appended to in the destination list
                const int m = ab[fk];
                const int lastm = ab[fl];
                ab[fk] = (uint16_t)lastm;
                pos[lastm] = (uint16_t)k;
                const int dAB = isab ? 1 : -1;
                cntA -= dAB;
                cntB += dAB;
                // the gap moved one slot towards the list that shrank: only the complex block's far end can collide
                if (cntC != 0 && (isab ? cs + cntC > N - cntB : cs < cntA)) cs = cl_make_room(ab, cs, cntC, cntA, N - cntB, lane);
                ab[tl] = (uint16_t)m;
                pos[m] = (uint16_t)nto;
                const uint32_t v = sn[m];
                nAB += dAB * ((int)((v >> SH_A) & CMASK) - (int)((v >> SH_B) & CMASK));
                sn[m] = (SnT)((v & ~3u) | (isab ? ST_B : ST_A));       // idempotent: every lane stores the same word
                const uint32_t dlt = (1u << SH_B) - (1u << SH_A);
                const NbrRow rowm = load_row<TM, OffT, PT>(words, ri, m, lw, ls);
                apply_row<false, SnT, PT>(sn, rowm, isab ? dlt : 0u - dlt, 0u, 0u);
                if (tails && rowm.nw > PT::WPP)
                    apply_row_tail<false, SnT, PT>(sn, words, rowm, isab ? dlt : 0u - dlt, 0u, 0u, lw, ls);
                __syncwarp();
            } else if (x < cP) {
                // A + B --> C (:256-289): pair number k of the nAB active pairs, enumerated along the
                // shorter of the A and B lists (weights nB resp. nA), then along the neighbour list.
                int k = (int)__umulhi(vw, (uint32_t)nAB);
                const bool sideA = cntA <= cntB;
                const int len = sideA ? cntA : cntB;
                const int sh = sideA ? SH_B : SH_A;
                const uint32_t want = sideA ? ST_B : ST_A;
                int xsel = 0, ix = 0;
                for (int c0 = 0; c0 < len; c0 += 32) {
                    const int idx = c0 + lane;
                    const int nact = min(32, len - c0);
                    int xm = 0, w = 0;
                    if (idx < len) {
                        xm = ab[sideA ? idx : NB - idx];
                        w = (int)(((uint32_t)sn[xm] >> sh) & CMASK);
                    }
                    const int incl = scan_small(w, lane, nact);
                    const int tot = __shfl_sync(FULL, incl, nact - 1);
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
                // k-th neighbour of xsel in state `want`; the first pass of xsel's row stays in registers for
                // the update below
                int ysel = 0;
                NbrRow rowx;
                ri.bounds(xsel, rowx.o0, rowx.nw);
                rowx.j = EMPTY;
                for (int w0 = 0; w0 < rowx.nw; w0 += PT::WPP) {
                    const uint32_t j = ld_entry<PT>(words, rowx.o0, rowx.nw, w0 + lw, ls);
                    if (w0 == 0) rowx.j = j;
                    const bool ok = j != EMPTY && ((uint32_t)sn[j != EMPTY ? j : 0u] & 3u) == want;
                    const uint32_t bal = __ballot_sync(FULL, ok);
                    const int c = __popc(bal);
                    if (k < c) {
                        // the lane holding set bit number k of the ballot
                        const uint32_t pick = __ballot_sync(FULL, ok && __popc(bal & lanemask_lt()) == k);
                        ysel = (int)__shfl_sync(FULL, j, __ffs(pick) - 1);
                        break;
                    }
                    k -= c;
                }
                const int ma = sideA ? xsel : ysel;   // the A member
                const int mb = sideA ? ysel : xsel;   // the B member
                const int ia = sideA ? ix : (int)pos[ma];
                const int ib = sideA ? (int)pos[mb] : ix;
                // remove from the A and B lists, append to the complex list (partner kept in pos)
                {
                    const int lastA = ab[cntA - 1];
                    ab[ia] = (uint16_t)lastA;
                    pos[lastA] = (uint16_t)ia;
                    --cntA;
                    const int lastB = ab[NB - (cntB - 1)];
                    ab[NB - ib] = (uint16_t)lastB;
                    pos[lastB] = (uint16_t)ib;
                    --cntB;
                    if (cntC == 0) cs = cntA;         // the gap just opened: [cntA, N - cntB) holds 2 slots
                    ab[cs + cntC] = (uint16_t)ma;
                    pos[ma] = (uint16_t)mb;
                    pos[mb] = (uint16_t)ma;
                    ++cntC;
                }
                const uint32_t va = sn[ma], vb = sn[mb];
                // a's neighbours lose an A neighbour, and b (one of them) drops B(2) -> C(0);
                // b's neighbours lose a B neighbour, and a drops A(1) -> C(0)
                const NbrRow rowy = load_row<TM, OffT, PT>(words, ri, ysel, lw, ls);
                NbrRow rowa, rowb;                  // selected by value: both stay in registers
                rowa.o0 = sideA ? rowx.o0 : rowy.o0; rowa.nw = sideA ? rowx.nw : rowy.nw; rowa.j = sideA ? rowx.j : rowy.j;
                rowb.o0 = sideA ? rowy.o0 : rowx.o0; rowb.nw = sideA ? rowy.nw : rowx.nw; rowb.j = sideA ? rowy.j : rowx.j;
                nAB -= (int)((va >> SH_B) & CMASK) + (int)((vb >> SH_A) & CMASK) - 1;
                __syncwarp();   // every lane has read va, vb before the partner words change
                apply_pair<SnT, PT>(sn, words, rowa, rowb, 0u - (1u << SH_A), 0u - (1u << SH_B), ma, mb, 0u - 2u, 0u - 1u, lw, ls, tails);
            } else {
                // C --> A + B (:291-344): complex k; the freed end is chosen by a fair coin (:300-307)
                const int k = (int)__umulhi(vw, (uint32_t)cntC);
                int ma = ab[cs + k];                 // the member that was A when the complex formed
                int mb = pos[ma];                    // its partner
                ab[cs + k] = ab[cs + cntC - 1];
                --cntC;
                if (cntC != 0 && (cs <= cntA || cs + cntC >= N - cntB)) cs = cl_make_room(ab, cs, cntC, cntA + 1, N - cntB - 1, lane);
                if (vw & 1u) {
                    const int tmp = ma;
                    ma = mb;
                    mb = tmp;
                }
                ab[cntA] = (uint16_t)ma;
                pos[ma] = (uint16_t)cntA;
                ++cntA;
                ab[NB - cntB] = (uint16_t)mb;
                pos[mb] = (uint16_t)cntB;
                ++cntB;
                const uint32_t va = sn[ma], vb = sn[mb];
                // a's neighbours gain an A neighbour, and b becomes B: C(0) -> B(2);
                // b's neighbours gain a B neighbour, and a becomes A: C(0) -> A(1)
                const NbrRow rowa = load_row<TM, OffT, PT>(words, ri, ma, lw, ls);
                const NbrRow rowb = load_row<TM, OffT, PT>(words, ri, mb, lw, ls);
                nAB += (int)((va >> SH_B) & CMASK) + (int)((vb >> SH_A) & CMASK) + 1;
                __syncwarp();
                apply_pair<SnT, PT>(sn, words, rowa, rowb, 1u << SH_A, 1u << SH_B, ma, mb, 2u, 1u, lw, ls, tails);
            }
          }
          ndraw += 32u;
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
    if (TM) {
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();                 // every warp of the CTA is done with its block
        if ((threadIdx.x >> 5) == 0)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(s_tmem), "n"(K1_TMEM_COLS));
    }
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
                                         uint32_t *deg, unsigned int *maxdeg) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    uint32_t c = 0;
    for (int j = 0; j < N; ++j)
        if (j != i && ref_dist_sq(locs + (size_t)i * DIM, locs + (size_t)j * DIM, DIM, L) <= reach2) ++c;
    deg[i] = c;
    if (maxdeg) atomicMax(maxdeg, c);
}

__global__ void fixed_graph_scan_kernel(uint32_t *deg_off, int N, int epw, unsigned long long *total) {
    // single block; serial chunks + shared scan (N <= 65535, run once per call).  In: degrees; out: exclusive
    // offsets in WORDS of the packed rows (epw entries per word), total words in *total
    __shared__ unsigned long long part[1024];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int per = (N + nt - 1) / nt;
    const int b = tid * per, e = min(b + per, N);
    unsigned long long s = 0;
    for (int i = b; i < e; ++i) s += (deg_off[i] + (uint32_t)epw - 1u) / (uint32_t)epw;
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
        const uint32_t v = (deg_off[i] + (uint32_t)epw - 1u) / (uint32_t)epw;
        deg_off[i] = (uint32_t)run;
        run += v;
    }
}

__global__ void fixed_graph_fill_kernel(const double *locs, int N, int DIM, double L, double reach2,
                                        const uint32_t *off, uint32_t *words, int epw, int eb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    RowPacker pk{words + off[i], 0u, 0, epw, eb};
    for (int j = 0; j < N; ++j)
        if (j != i && ref_dist_sq(locs + (size_t)i * DIM, locs + (size_t)j * DIM, DIM, L) <= reach2) pk.push((uint32_t)j);
    pk.flush();
}

// ------------------------------------------------------------------------------------------------
// K3: replicate reduction (fixed count) -- sum and sum of squares per (slot, bin) over nrep curves
// ------------------------------------------------------------------------------------------------
__global__ void reduce_fixed_kernel(const uint16_t *__restrict__ curve, int repstride, int nrep, int T, int nsplit,
                                    long long *__restrict__ sum, unsigned long long *__restrict__ sumsq) {
    const int slot = blockIdx.x;
    // blockIdx.y takes a contiguous share of the replicates (nsplit > 1 only when one point has very many
    // replicates; integer sums, so the atomic accumulation stays bit-exact)
    const int per = (nrep + nsplit - 1) / nsplit;
    const int r0 = blockIdx.y * per, r1 = min(r0 + per, nrep);
    const uint16_t *base = curve + (size_t)slot * repstride * T;
    for (int s = threadIdx.x; s < T; s += blockDim.x) {
        long long S = 0;
        unsigned long long Q = 0;
        for (int r = r0; r < r1; ++r) {
            const unsigned int v = base[(size_t)r * T + s];
            S += v;
            Q += (unsigned long long)v * v;
        }
        if (nsplit == 1) {
            sum[(size_t)slot * T + s] = S;
            sumsq[(size_t)slot * T + s] = Q;
        } else if (r1 > r0) {
            atomicAdd(reinterpret_cast<unsigned long long *>(sum) + (size_t)slot * T + s, (unsigned long long)S);
            atomicAdd(sumsq + (size_t)slot * T + s, Q);
        }
    }
}

// n*Q - S^2 (>= 0 by Cauchy-Schwarz) as a double.  The product n*Q passes 2^64 once n*N exceeds about 3e9
// (maxsims up to 10^6 with N up to 65535), so the difference is formed in 128 bits; it is converted as
// hi * 2^64 + lo with explicit roundings, the sequence oracle/spr_mirror.c repeats (hi == 0 in every
// realistic configuration: then this is the plain conversion of a 64-bit integer).
__device__ __forceinline__ double var_numerator(unsigned long long n, unsigned long long Q, unsigned long long S) {
    const unsigned __int128 d = (unsigned __int128)n * Q - (unsigned __int128)S * S;
    const unsigned long long hi = (unsigned long long)(d >> 64), lo = (unsigned long long)d;
    const double dlo = __ull2double_rn(lo);
    if (hi == 0ULL) return dlo;
    return __dadd_rn(__dmul_rn(__ull2double_rn(hi), 18446744073709551616.0), dlo);
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
                if (var_numerator((unsigned long long)n, Q, (unsigned long long)S) > tol2 * dS * dS * (double)(n - 1)) viol = 1;
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
        if (k < count && s < T) {
            const long long i1 = idx1[k];       // < 1: padding row of a gathered slab, no table column
            if (i1 >= 1) table[(size_t)s * P + (size_t)(i1 - 1)] = tile[threadIdx.x][r];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K5: the consumer of the table: surrogate_sprdata_error (/root/reference/src/fitting.jl:38-91) for a whole
// optimiser population at once.  One CTA per candidate [logkon, logkoff, logkonb, reach, logCP]; each
// thread takes samples (concentration j, time t) in a strided loop: 5-D multilinear interpolation of
// the FirstMoment table at fractional 1-based indices (Interpolations.jl BSpline(Linear()),
// surrogate.jl:66-70) = 32 gathers, squared residual against the data, fixed-order block reduction.
// A candidate or time outside the table gives +Inf (Interpolations throws BoundsError there; the
// reference keeps the optimiser inside with checkranges / rescale_logkon_max, fitting.jl:93-114).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool lut_index(double q, int n, int *i0, double *w) {
    // fractional 1-based index q in [1, n] -> 0-based lower corner and weight of the upper corner
    if (!(q >= 1.0 && q <= (double)n)) return false;
    if (n == 1) { *i0 = 0; *w = 0.0; return true; }
    int i = (int)floor(q);
    if (i > n - 1) i = n - 1;
    *i0 = i - 1;
    *w = q - (double)i;
    return true;
}

__global__ void __launch_bounds__(128) fit_loss_kernel(const FitArgs a) {
    __shared__ double red[128];
    __shared__ double cw[8];
    __shared__ long long coff[8];
    __shared__ int s_bad;
    const int cand = blockIdx.x, tid = threadIdx.x;
    const double *op = a.optpars + (size_t)cand * 5;
    const long long n1 = a.n[0], n2 = a.n[1], n3 = a.n[2];
    if (tid == 0) {
        int i2, i3, i4;
        double w2, w3, w4;
        const double q2 = (op[1] - a.lo[1]) * (double)(a.n[1] - 1) / (a.hi[1] - a.lo[1]) + 1.0;
        const double q3 = (op[2] - a.lo[2]) * (double)(a.n[2] - 1) / (a.hi[2] - a.lo[2]) + 1.0;
        const double q4 = (op[3] - a.lo[3]) * (double)(a.n[3] - 1) / (a.hi[3] - a.lo[3]) + 1.0;
        const bool ok = lut_index(q2, a.n[1], &i2, &w2) && lut_index(q3, a.n[2], &i3, &w3) && lut_index(q4, a.n[3], &i4, &w4);
        s_bad = ok ? 0 : 1;
        if (ok) {
            for (int c = 0; c < 8; ++c) {
                const int b2 = c & 1, b3 = (c >> 1) & 1, b4 = (c >> 2) & 1;
                // an upper corner with zero weight may lie outside the table (q == n): clamp its index
                const int j2 = min(i2 + b2, a.n[1] - 1), j3 = min(i3 + b3, a.n[2] - 1), j4 = min(i4 + b4, a.n[3] - 1);
                cw[c] = (b2 ? w2 : 1.0 - w2) * (b3 ? w3 : 1.0 - w3) * (b4 ? w4 : 1.0 - w4);
                coff[c] = n1 * ((long long)j2 + n2 * ((long long)j3 + n3 * (long long)j4));
            }
        }
    }
    __syncthreads();
    double err = 0.0;
    bool bad = s_bad != 0;
    if (!bad) {
        const double cp = pow(10.0, op[4]);
        const double tspan = a.t_last - a.t_first;
        int j = 0;
        long long first = 0;          // index of the first sample of concentration j
        for (long long k = tid; k < a.nsamples; k += blockDim.x) {
            while (k >= first + a.npts[j]) { first += a.npts[j]; ++j; }
            // rescale the on rate for the antibody concentration of this curve (fitting.jl:71)
            const double logkon = op[0] + log10(a.abcs[j] / a.abcs[0]);
            const double q1 = (logkon - a.lo[0]) * (double)(a.n[0] - 1) / (a.hi[0] - a.lo[0]) + 1.0;
            const double sq = (a.times[k] - a.t_first) * (double)(a.T - 1) / tspan + 1.0;
            int i1, is;
            double w1, ws;
            if (!lut_index(q1, a.n[0], &i1, &w1) || !lut_index(sq, a.T, &is, &ws)) { bad = true; break; }
            const int i1b = min(i1 + 1, a.n[0] - 1), isb = min(is + 1, a.T - 1);
            const double *t0 = a.table + (size_t)is * (size_t)a.P, *t1 = a.table + (size_t)isb * (size_t)a.P;
            double v = 0.0;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const long long o = coff[c];
                const double lo_t = (1.0 - w1) * t0[o + i1] + w1 * t0[o + i1b];
                const double hi_t = (1.0 - w1) * t1[o + i1] + w1 * t1[o + i1b];
                v += cw[c] * ((1.0 - ws) * lo_t + ws * hi_t);
            }
            const double r = cp * v - a.values[k];
            err += r * r;
        }
    }
    red[tid] = err;
    if (bad) atomicExch(&s_bad, 1);
    __syncthreads();
    for (int st = 64; st > 0; st >>= 1) {
        if (tid < st) red[tid] += red[tid + st];
        __syncthreads();
    }
    if (tid == 0) a.loss[cand] = s_bad ? __longlong_as_double(0x7ff0000000000000LL) : red[0];
}

inline int align16(int x) { return (x + 15) & ~15; }

template <int DIM>
cudaError_t set_table_attr(int bytes) {
    static int configured = -1;
    if (bytes > configured) {
        cudaError_t e = cudaFuncSetAttribute(spr_table_kernel<DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e != cudaSuccess) return e;
        configured = bytes;
    }
    return cudaSuccess;
}

template <typename SnT, typename OffT, int EB, bool TM, bool TS>
cudaError_t set_traj_attr(int bytes) {
    static int configured = -1;
    if (bytes > configured) {
        cudaError_t e = cudaFuncSetAttribute(spr_traj_kernel<SnT, OffT, EB, TM, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e != cudaSuccess) return e;
        configured = bytes;
    }
    return cudaSuccess;
}

template <typename F>
cudaError_t dispatch_traj(int wide_sn, int wide_off, F &&f) {
    if (wide_sn) return wide_off ? f(uint32_t{}, uint32_t{}) : f(uint32_t{}, uint16_t{});
    return wide_off ? f(uint16_t{}, uint32_t{}) : f(uint16_t{}, uint16_t{});
}

}  // namespace

int table_smem_bytes(int N) { return k2_bytes(N); }
int table_stage_cap(int N, int DIM, double L, double reach) { return k2_stage_cap(N, DIM, L, reach); }

int table_ctas_per_sm(int N, int DIM) {
    int nb = 0;
    const int bytes = k2_bytes(N);
    cudaError_t e;
    if (DIM == 3) {
        e = set_table_attr<3>(bytes);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, spr_table_kernel<3>, K2_THREADS, bytes);
    } else {
        e = set_table_attr<2>(bytes);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, spr_table_kernel<2>, K2_THREADS, bytes);
    }
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    return nb;
}

cudaError_t launch_table(const TableArgs &a, int DIM, int nblocks, cudaStream_t st) {
    const int bytes = k2_bytes(a.N);
    if (DIM == 3) {
        cudaError_t e = set_table_attr<3>(bytes);
        if (e != cudaSuccess) return e;
        spr_table_kernel<3><<<nblocks, K2_THREADS, bytes, st>>>(a);
    } else {
        cudaError_t e = set_table_attr<2>(bytes);
        if (e != cudaSuccess) return e;
        spr_table_kernel<2><<<nblocks, K2_THREADS, bytes, st>>>(a);
    }
    return cudaGetLastError();
}

bool make_traj_layout(int N, bool wide_sn, bool wide_off, int smem_optin, int smem_per_sm, TrajLayout *out, bool tmem) {
    TrajLayout l{};
    l.wide_sn = wide_sn ? 1 : 0;
    l.wide_off = wide_off ? 1 : 0;
    l.eb = entry_bits(N);
    l.tmem = (tmem && !wide_sn && !wide_off && l.eb == 10 && N <= 1024) ? 1 : 0;
    int o = 0;
    l.o_off = o;  if (!l.tmem) o += align16((N + 1) * (wide_off ? 4 : 2));
    l.o_sn = o;   o += align16(N * (wide_sn ? 4 : 2));
    l.o_ab = o;   o += align16(N * 2);
    l.o_pos = o;  o += align16(N * 2);
    l.slice = o;
    // warps per CTA: most resident trajectories per SM (every CTA also costs 1 KB of reserved shared memory);
    // 72 registers allow 28 warps per SM
    int best = 0, best_w = 0, best_c = 0;
    if (l.tmem) {
        // fixed shape: 8 warps per CTA (two column blocks of 32 per lane quadrant), 4 CTAs per SM, 64 registers
        l.wpc = K1_WPC_TM;
        l.cta_per_sm = 4;
        *out = l;
        return (long long)l.wpc * l.slice <= smem_optin && 4LL * ((long long)l.wpc * l.slice + 1024) <= smem_per_sm;
    }
    for (int w = 1; w <= K1_WPC; ++w) {
        const long long bytes = (long long)w * l.slice;
        if (bytes > smem_optin) break;
        int c = (int)(smem_per_sm / (bytes + 1024));
        if (c > 32) c = 32;
        const int regcap = K1_MAX_WARPS / w;
        if (c > regcap) c = regcap;
        if (c * w >= best && c * w > 0) { best = c * w; best_w = w; best_c = c; }
    }
    l.wpc = best_w;
    l.cta_per_sm = best_c;
    *out = l;
    return best > 0;
}

cudaError_t launch_traj(const TrajArgs &a, int nblocks, cudaStream_t st) {
    return dispatch_traj(a.lay.wide_sn, a.lay.wide_off, [&](auto sn_t, auto off_t) -> cudaError_t {
        using S_ = decltype(sn_t);
        using O_ = decltype(off_t);
        const int bytes = a.lay.slice * a.lay.wpc;
        if (a.lay.tmem) {
            // row offsets in tensor memory: 16-bit states and offsets, 10-bit entries only (N <= 1023)
            if (!std::is_same<S_, uint16_t>::value || !std::is_same<O_, uint16_t>::value || a.lay.eb != 10) return cudaErrorInvalidValue;
            cudaError_t e = set_traj_attr<uint16_t, uint16_t, 10, true, false>(bytes);
            if (e != cudaSuccess) return e;
            spr_traj_kernel<uint16_t, uint16_t, 10, true, false><<<nblocks, 32 * a.lay.wpc, bytes, st>>>(a);
        } else if (a.lay.tbl_words > 0) {
            // rows in shared memory (underfilled launches): the common instantiation only
            if (!std::is_same<S_, uint16_t>::value || !std::is_same<O_, uint16_t>::value || a.lay.eb != 10) return cudaErrorInvalidValue;
            cudaError_t e = set_traj_attr<uint16_t, uint16_t, 10, false, true>(bytes);
            if (e != cudaSuccess) return e;
            spr_traj_kernel<uint16_t, uint16_t, 10, false, true><<<nblocks, 32 * a.lay.wpc, bytes, st>>>(a);
        } else if (a.lay.eb == 10) {
            cudaError_t e = set_traj_attr<S_, O_, 10, false, false>(bytes);
            if (e != cudaSuccess) return e;
            spr_traj_kernel<S_, O_, 10, false, false><<<nblocks, 32 * a.lay.wpc, bytes, st>>>(a);
        } else {
            cudaError_t e = set_traj_attr<S_, O_, 16, false, false>(bytes);
            if (e != cudaSuccess) return e;
            spr_traj_kernel<S_, O_, 16, false, false><<<nblocks, 32 * a.lay.wpc, bytes, st>>>(a);
        }
        return cudaGetLastError();
    });
}

cudaError_t launch_fixed_graph(const double *d_locs, int N, int DIM, double L, double reach, uint32_t *d_off,
                               uint32_t *d_words, unsigned long long *d_total, unsigned int *d_maxdeg, cudaStream_t st) {
    const double reach2 = reach * reach;
    const int nb = (N + 127) / 128;
    const int eb = entry_bits(N), epw = entries_per_word(N);
    if (d_words == nullptr) {
        fixed_graph_count_kernel<<<nb, 128, 0, st>>>(d_locs, N, DIM, L, reach2, d_off, d_maxdeg);
        fixed_graph_scan_kernel<<<1, 1024, 0, st>>>(d_off, N, epw, d_total);
    } else {
        fixed_graph_fill_kernel<<<nb, 128, 0, st>>>(d_locs, N, DIM, L, reach2, d_off, d_words, epw, eb);
    }
    return cudaGetLastError();
}

cudaError_t launch_fit_loss(const FitArgs &a, long long npop, cudaStream_t st) {
    if (npop <= 0) return cudaSuccess;
    fit_loss_kernel<<<(unsigned)npop, 128, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_reduce_fixed(const uint16_t *curve, int nslots, int repstride, int nrep, int T, long long *sum,
                                unsigned long long *sumsq, cudaStream_t st) {
    if (nslots <= 0) return cudaSuccess;
    // enough blocks to fill the GPU even for a single point with many replicates
    int nsplit = 1;
    if ((long long)nslots < 1024 && nrep > 256) {
        nsplit = (nrep + 255) / 256;
        const int want = (1024 + nslots - 1) / nslots;
        if (nsplit > want) nsplit = want;
        if (nsplit > 65535) nsplit = 65535;
    }
    if (nsplit > 1) {
        cudaError_t e = cudaMemsetAsync(sum, 0, sizeof(long long) * (size_t)nslots * T, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(sumsq, 0, sizeof(unsigned long long) * (size_t)nslots * T, st);
        if (e != cudaSuccess) return e;
    }
    reduce_fixed_kernel<<<dim3((unsigned)nslots, (unsigned)nsplit), 128, 0, st>>>(curve, repstride, nrep, T, nsplit, sum, sumsq);
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
