// slos_kernels.cuh -- SLOS state-vector evolution and the distribution epilogues.
//
//  K3  slos_layer_kernel     : one photon-addition layer of SLOSBackend.calculate
//                              (lightworks/emulator/backends/slos.py:96-103, a_i_dagger :108-123,
//                              add_dicts :131-138) in gather form over rank-indexed layers.
//  K4a reorder_to_colex_kernel: SLOS (lexicographic) order -> fock_basis (colex) order.
//  K4b marginalise_kernel    : threshold + loss-mode marginalisation with first-seen bookkeeping
//                              (permanent.py:148-157, slos.py:73-79).
//
// Every layer k holds one complex128 amplitude per Fock state of k photons in m modes, stored in the
// reference's own SLOS order (dict insertion order = lexicographic order of the ascending mode list
// x_0 <= ... <= x_{k-1}), so the final layer needs no permutation.  The reference applies, for the k-th
// input photon (in mode i_k),
//     out[t] = sum_j sqrt(t_j) * in[t - e_j] * U[j, i_k]          (j ascending, slos.py:98-103)
// The kernel evaluates exactly that sum per child state t, in the same j order and with the same rounding
// sequence (no FMA contraction): (sqrt(t_j) * value) * U[j,i] then complex add, so amplitudes agree with the
// reference's dict arithmetic bit for bit in practice (e.g. the exact 0 of the HOM dip,
// tests/emulator/backend_test.py:236).
//
// Indexing.  With G(v, r) = C(m-v+r, r+1) (non-decreasing lists of length r+1 over values >= v) the lex rank
// is  rank = sum_i [ G(x_{i-1}, r_i) - G(x_i, r_i) ],  r_i = k-1-i,  x_{-1} = 0.  Removing the element at
// position p gives the parent (k-1 photons) with
//     rank(parent) = rank - Q[p+1],   Q[j] = sum_{i<j} [ H(x_{i-1}, r_i) - H(x_i, r_i) ],  H(v, r) = C(m-v+r-1, r+1)
// (Pascal's rule applied to the difference of the two rank sums), i.e. one prefix table per state.  A lane
// walks ranks base + 32*it + lane: consecutive lanes hold consecutive children, so stores and most parent
// gathers are coalesced; advancing by 32 re-derives only the trailing positions whose block the step leaves.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fock.cuh"

namespace lwb {

constexpr int SLOS_THREADS = 256;
#ifndef LWB_SLOS_ITERS
#define LWB_SLOS_ITERS 16
#endif
#ifndef LWB_SLOS_MINB
#define LWB_SLOS_MINB 4
#endif
constexpr int SLOS_ITERS = LWB_SLOS_ITERS;  // children per lane

struct SlosLayerArgs {
    const uint64_t* bin;    // binomial table (device, LWB_BIN_B stride)
    int m, k;               // modes, photons in the CHILD layer
    uint64_t S;             // number of child states
    const double2* parent;  // layer k-1 (lex order)
    void* child;            // layer k (lex order): double2 amplitudes, or double |a|^2 when write_probs
    const double2* U;       // m x m
    int in_mode;            // i_k
    const double* sqrt_tab; // sqrt_tab[c] = pow(c, 0.5) computed on the host (slos.py:121)
    int write_probs;        // final layer only: store abs(a)**2 (slos.py:74) instead of the amplitude
};

// complex multiply exactly as CPython's complex_mul: (ac - bd) + (ad + bc)i, no fma
__device__ __forceinline__ double2 cmul_py(double2 a, double2 b) {
    double2 r;
    r.x = __dsub_rn(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y));
    r.y = __dadd_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x));
    return r;
}

// K = photons in the child layer (exact: every loop bound and table row is static);
// IDX = uint32_t when the layer (and every table entry) fits 32 bits, else uint64_t.
template <int K, typename IDX>
__global__ void __launch_bounds__(SLOS_THREADS, LWB_SLOS_MINB) slos_layer_kernel(SlosLayerArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int m = a.m;
    const int mv = m + 1;  // table width over v in [0, m]
    // smem: ucol [m] double2 | sq [K+1] double | cnt[K][mv] | Gt[K][mv] | Ht[K][mv]   (rows indexed by r = rem)
    double2* ucol = reinterpret_cast<double2*>(smem_raw);
    double* sq = reinterpret_cast<double*>(ucol + m);
    IDX* cnt = reinterpret_cast<IDX*>(sq + ((K + 2) & ~1));
    IDX* Gt = cnt + K * mv;
    IDX* Ht = Gt + K * mv;
    for (int e = threadIdx.x; e < m; e += SLOS_THREADS) ucol[e] = a.U[(size_t)e * m + a.in_mode];
    for (int e = threadIdx.x; e <= K; e += SLOS_THREADS) sq[e] = a.sqrt_tab[e];
    for (int e = threadIdx.x; e < K * mv; e += SLOS_THREADS) {
        const int r = e / mv, v = e - r * mv;
        cnt[e] = (IDX)binom(a.bin, m - v + r - 1, r);      // lists of length r+1 over >= v that START with v
        Gt[e] = (IDX)binom(a.bin, m - v + r, r + 1);       // all lists of length r+1 over values >= v
        Ht[e] = (IDX)binom(a.bin, m - v + r - 1, r + 1);
    }
    __syncthreads();

    const uint64_t warp_base = ((uint64_t)blockIdx.x * (SLOS_THREADS / 32) + (threadIdx.x >> 5)) * (32ull * SLOS_ITERS);
    const int lane = threadIdx.x & 31;
    if (warp_base >= a.S) return;

    int x[K];
    IDX Q[K + 1];   // Q[j] = sum_{i<j} dq(i)
    IDX O[K];       // O[i] = offset of the state inside the block of states sharing x_0..x_{i-1}
    IDX W[K];       // W[i] = size of that block = G(x_{i-1}, r_i)
    IDX rank = (IDX)(warp_base + lane);
    int level = 0;  // positions >= level must be (re)derived from offset `off`
    IDX off = rank;
    Q[0] = 0;

#pragma unroll 1
    for (int it = 0; it < SLOS_ITERS; ++it) {
        if ((uint64_t)rank >= a.S) break;
        // ---- (re)derive positions >= level: lex unrank of `off` inside the block of the kept prefix ----
        {
            int prev = 0;
#pragma unroll
            for (int pos = 0; pos < K; ++pos) {
                if (pos >= level) {
                    const int r = K - 1 - pos;  // static after unrolling
                    W[pos] = Gt[r * mv + prev];
                    O[pos] = off;
                    int v = prev;
                    if (r == 0) {
                        v = prev + (int)off;  // one list per value
                        off = 0;
                    } else {
                        const IDX* c = cnt + r * mv;
                        IDX cv = c[v];
                        while (off >= cv) { off -= cv; ++v; cv = c[v]; }
                    }
                    x[pos] = v;
                    Q[pos + 1] = Q[pos] + (Ht[r * mv + prev] - Ht[r * mv + v]);
                }
                prev = x[pos];
            }
        }
        // ---- gather parents: one term per distinct occupied mode, ascending (slos.py:98-103) ----
        double2 acc = make_double2(0.0, 0.0);
        int run = 0;
#pragma unroll
        for (int p = 0; p < K; ++p) {
            run += 1;
            bool last_of_run = true;
            if (p + 1 < K) last_of_run = x[p + 1 < K ? p + 1 : p] != x[p];
            if (last_of_run) {
                const IDX prank = rank - Q[p + 1];
                LWB_BOUNDS((uint64_t)prank < a.S);  // a parent layer is never larger than its child layer
                double2 t = a.parent[prank];
                if (run > 1) {  // key[mode] ** 0.5 * value ; multiplying by sqrt(1) = 1.0 is exact, skip it
                    const double s = sq[run];
                    t.x = __dmul_rn(s, t.x);
                    t.y = __dmul_rn(s, t.y);
                }
                const double2 term = cmul_py(t, ucol[x[p]]);
                // the first term is added to +0.0: x + 0.0 == x for every x except -0.0, and a -0.0 first
                // term can only differ from the reference (which assigns it) in the sign of an exact zero
                acc.x = __dadd_rn(acc.x, term.x);
                acc.y = __dadd_rn(acc.y, term.y);
                run = 0;
            }
        }
        if (a.write_probs) {
            const double ab = hypot(acc.x, acc.y);  // abs(p) ** 2, slos.py:74
            reinterpret_cast<double*>(a.child)[rank] = ab * ab;
        } else {
            reinterpret_cast<double2*>(a.child)[rank] = acc;
        }
        // ---- advance by 32: keep the longest prefix whose block still contains the new rank ----
        rank += 32;
        level = 0;
        off = rank;
#pragma unroll
        for (int i = 1; i < K; ++i) {
            const IDX o = O[i] + 32;
            if (o < W[i] && level == i - 1) {  // blocks nest: level i is usable only if level i-1 is
                level = i;
                off = o;
            }
        }
#pragma unroll
        for (int i = 0; i < K; ++i)
            if (i < level) O[i] += 32;  // offsets of the kept levels move with the rank
    }
}

// ---------------------------------------------------------------------------------------------
// out[r] (r = colex / fock_basis rank) = f(in[SLOS-order position of that state]);
// mode 0: copy complex amplitude; mode 1: |a|^2 from complex; mode 2: copy real.
struct ReorderArgs {
    const uint64_t* bin;
    int m, n;
    uint64_t S;
    const double* in;
    double* out;
    int mode;
};

template <int KMAX>
__global__ void __launch_bounds__(SLOS_THREADS) reorder_to_colex_kernel(ReorderArgs a) {
    const uint64_t r = (uint64_t)blockIdx.x * SLOS_THREADS + threadIdx.x;
    if (r >= a.S) return;
    int x[KMAX];
    {   // colex greedy unrank with static indexing
        uint64_t rem = r;
        int y = a.m + a.n - 2;
#pragma unroll
        for (int i = KMAX; i >= 1; --i) {
            if (i <= a.n) {
                while (binom(a.bin, y, i) > rem) --y;
                rem -= binom(a.bin, y, i);
                x[i - 1] = y - i + 1;
                y -= 1;
            }
        }
    }
    uint64_t p = 0;  // lex rank
    int prev = 0;
#pragma unroll
    for (int i = 0; i < KMAX; ++i) {
        if (i < a.n) {
            const int rem = a.n - 1 - i;
            p += binom(a.bin, a.m - prev + rem, rem + 1) - binom(a.bin, a.m - x[i] + rem, rem + 1);
            prev = x[i];
        }
    }
    if (a.mode == 0) {
        reinterpret_cast<double2*>(a.out)[r] = reinterpret_cast<const double2*>(a.in)[p];
    } else if (a.mode == 1) {
        const double2 v = reinterpret_cast<const double2*>(a.in)[p];
        const double ab = hypot(v.x, v.y);
        a.out[r] = ab * ab;
    } else {
        a.out[r] = a.in[p];
    }
}

// |a|^2, same order: probs[r] = hypot(re, im)^2
static __global__ void abs2_kernel(const double2* __restrict__ amps, uint64_t S, double* __restrict__ probs) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= S) return;
    const double2 v = amps[r];
    const double ab = hypot(v.x, v.y);
    probs[r] = ab * ab;
}

// ---------------------------------------------------------------------------------------------
// Loss-mode marginalisation (permanent.py:148-157, slos.py:73-79).
// probs: per FULL state (m_tot modes, n photons) in colex order.  Measured keys = states of kk photons
// in the first n_modes modes, kk = k_min..n; key index = key_off[kk] + colex rank among (n_modes, kk).
// One thread per key walks every loss-mode completion in ascending full rank (= reference enumeration
// order, so the floating-point summation order matches), adds p where p > threshold, and records the
// first contributing position in the requested order (order 0: colex rank, 1: SLOS/lex rank).
struct MargArgs {
    const uint64_t* bin;
    int m_tot, n_modes, n, k_min;
    const uint64_t* key_off;  // [n+2] device
    uint64_t n_keys;
    const double* probs;
    double threshold;
    int order;
    double* marg;        // [n_keys]
    uint64_t* first_pos; // [n_keys], UINT64_MAX when no state passed the threshold
};

template <int KMAX>
__global__ void __launch_bounds__(SLOS_THREADS) marginalise_kernel(MargArgs a) {
    const uint64_t key = (uint64_t)blockIdx.x * SLOS_THREADS + threadIdx.x;
    if (key >= a.n_keys) return;
    int kk = a.k_min;
    while (kk < a.n && key >= a.key_off[kk + 1]) ++kk;
    const uint64_t krank = key - a.key_off[kk];
    int x[KMAX];  // full ascending mode list: kk measured photons then n-kk loss photons
    {
        int tmp[KMAX];
        colex_unrank(a.bin, krank, kk, a.n_modes, tmp);
        for (int i = 0; i < kk; ++i) x[i] = tmp[i];
    }
    uint64_t base = 0;
    for (int i = 1; i <= kk; ++i) base += binom(a.bin, x[i - 1] + i - 1, i);
    const int nl = a.n - kk, lm = a.m_tot - a.n_modes;
    int y[KMAX];  // loss part, ascending, values in [0, lm)
    for (int i = 0; i < nl; ++i) y[i] = 0;
    double sum = 0.0;
    uint64_t first = UINT64_MAX;
    bool any = false;
    if (nl > 0 && lm == 0) {
        a.marg[key] = 0.0;
        a.first_pos[key] = UINT64_MAX;
        return;
    }
    for (;;) {
        uint64_t r = base;
        for (int i = 0; i < nl; ++i) r += binom(a.bin, a.n_modes + y[i] + (kk + i + 1) - 1, kk + i + 1);
        const double p = a.probs[r];
        if (p > a.threshold) {
            sum = any ? sum + p : p;
            any = true;
            uint64_t pos = r;
            if (a.order == 1) {
                for (int i = 0; i < nl; ++i) x[kk + i] = a.n_modes + y[i];
                pos = lex_rank(a.bin, x, a.n, a.m_tot);
            }
            if (pos < first) first = pos;
        }
        if (nl == 0 || !colex_next(y, nl, lm)) break;
    }
    a.marg[key] = sum;
    a.first_pos[key] = any ? first : UINT64_MAX;
}

}  // namespace lwb
