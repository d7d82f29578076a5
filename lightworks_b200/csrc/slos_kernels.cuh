// slos_kernels.cuh -- SLOS state-vector evolution and the distribution epilogues.
//
//  K3  slos_layer_kernel     : one photon-addition layer of SLOSBackend.calculate
//                              (lightworks/emulator/backends/slos.py:96-103, a_i_dagger :108-123,
//                              add_dicts :131-138) in gather form over rank-indexed layers.
//  K4a reorder_kernel        : colex (fock_basis) order -> SLOS dict-insertion (lexicographic) order,
//                              optionally |a|^2.
//  K4b marginalise_kernel    : threshold + loss-mode marginalisation with first-seen bookkeeping
//                              (permanent.py:148-157, slos.py:73-79).
//
// Layer k holds one complex128 amplitude per Fock state of k photons in m modes, indexed by the
// colex rank of fock.cuh.  The reference applies, for input photon number k (in mode i_k),
//     out[t] = sum_j sqrt(t_j) * in[t - e_j] * U[j, i_k]          (j ascending, slos.py:98-103)
// The kernel evaluates exactly that sum per child state t, in the same j order and with the same
// rounding sequence (no FMA contraction): (sqrt(t_j) * value) * U[j,i] then complex add, so the
// amplitudes agree with the reference's dict arithmetic bit for bit (e.g. the exact 0 of the HOM
// dip, tests/emulator/backend_test.py:236).
//
// Parent rank without re-ranking: with A_i = C(x_i+i-1, i) and B_i = C(x_i+i-2, i-1) (1-based i over
// the ascending mode list x), removing the element at position p gives
//     rank(t - e_{x_p}) = sum_{i<p} A_i + sum_{i>p} B_i .
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fock.cuh"

namespace lwb {

constexpr int SLOS_THREADS = 256;

struct SlosLayerArgs {
    const uint64_t* bin;    // binomial table (device, LWB_BIN_B stride)
    int m, k;               // modes, photons in the CHILD layer
    uint64_t S;             // number of child states
    const double2* parent;  // layer k-1
    double2* child;         // layer k
    const double2* U;       // m x m
    int in_mode;            // i_k
    const double* sqrt_tab; // sqrt_tab[c] = pow(c, 0.5) computed on the host (slos.py:121)
};

// complex multiply exactly as CPython's complex_mul: (ac - bd) + (ad + bc)i, no fma
__device__ __forceinline__ double2 cmul_py(double2 a, double2 b) {
    double2 r;
    r.x = __dsub_rn(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y));
    r.y = __dadd_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x));
    return r;
}

template <int KMAX>
__global__ void __launch_bounds__(SLOS_THREADS) slos_layer_kernel(SlosLayerArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // smem: binomials [m+k][k+1] u64, U column [m] double2, sqrt table [k+1]
    const int k = a.k, m = a.m;
    const int bw = k + 1, bh = m + k;
    uint64_t* sb = reinterpret_cast<uint64_t*>(smem_raw);
    double2* ucol = reinterpret_cast<double2*>(sb + (((size_t)bh * bw + 1) & ~(size_t)1));  // keep 16-byte alignment
    double* sq = reinterpret_cast<double*>(ucol + m);
    for (int e = threadIdx.x; e < bh * bw; e += SLOS_THREADS) sb[e] = a.bin[(e / bw) * LWB_BIN_B + (e % bw)];
    for (int e = threadIdx.x; e < m; e += SLOS_THREADS) ucol[e] = a.U[(size_t)e * m + a.in_mode];
    for (int e = threadIdx.x; e <= k; e += SLOS_THREADS) sq[e] = a.sqrt_tab[e];
    __syncthreads();

    const uint64_t r = (uint64_t)blockIdx.x * SLOS_THREADS + threadIdx.x;
    if (r >= a.S) return;

    // unrank (colex greedy) -> x[1..k] ascending, and A_i, B_i on the fly
    int x[KMAX + 1];
    uint64_t A[KMAX + 1], B[KMAX + 2];
    {
        uint64_t rem = r;
        int y = m + k - 2;
#pragma unroll
        for (int i = KMAX; i >= 1; --i) {
            if (i <= k) {
                while (sb[y * bw + i] > rem) --y;
                const uint64_t c = sb[y * bw + i];
                rem -= c;
                x[i] = y - i + 1;
                A[i] = c;
                B[i] = (y >= 1) ? sb[(y - 1) * bw + (i - 1)] : (i == 1 ? 1 : 0);
                y -= 1;
                if (y < 0) y = 0;
            }
        }
    }
    // suffix sums of B (positions > p) and running prefix of A
    uint64_t SB[KMAX + 2];
    SB[k + 1 <= KMAX + 1 ? k + 1 : KMAX + 1] = 0;
#pragma unroll
    for (int i = KMAX; i >= 1; --i)
        if (i <= k) SB[i] = SB[i + 1] + B[i];

    double2 acc = make_double2(0.0, 0.0);
    bool first = true;
    uint64_t pa = 0;  // sum_{i<p} A_i
    int run = 0;
#pragma unroll
    for (int p = 1; p <= KMAX; ++p) {
        if (p <= k) {
            run += 1;
            const bool last_of_run = (p == k) || (x[p + 1] != x[p]);
            if (last_of_run) {
                // remove the element at position p (the last of its run): prefix uses A of positions < p
                const uint64_t prank = pa + SB[p + 1];
                const double2 v = a.parent[prank];
                const double s = sq[run];
                double2 t;
                t.x = __dmul_rn(s, v.x);
                t.y = __dmul_rn(s, v.y);
                const double2 term = cmul_py(t, ucol[x[p]]);
                if (first) { acc = term; first = false; }
                else { acc.x = __dadd_rn(acc.x, term.x); acc.y = __dadd_rn(acc.y, term.y); }
                run = 0;
            }
            pa += A[p];
        }
    }
    a.child[r] = acc;
}

// ---------------------------------------------------------------------------------------------
// out[p] (p = SLOS / lexicographic position) = f(in[colex rank of that state]);
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
__global__ void __launch_bounds__(SLOS_THREADS) reorder_lex_kernel(ReorderArgs a) {
    const uint64_t p = (uint64_t)blockIdx.x * SLOS_THREADS + threadIdx.x;
    if (p >= a.S) return;
    int x[KMAX];
    // lex_unrank with static indexing
    {
        uint64_t rank = p;
        int prev = 0;
#pragma unroll
        for (int i = 0; i < KMAX; ++i) {
            if (i < a.n) {
                const int rem = a.n - 1 - i;
                int v = prev;
                for (;;) {
                    const uint64_t cnt = binom(a.bin, a.m - v + rem - 1, rem);
                    if (rank < cnt) break;
                    rank -= cnt;
                    ++v;
                }
                x[i] = v;
                prev = v;
            }
        }
    }
    uint64_t r = 0;
#pragma unroll
    for (int i = 1; i <= KMAX; ++i)
        if (i <= a.n) r += binom(a.bin, x[i - 1] + i - 1, i);
    if (a.mode == 0) {
        reinterpret_cast<double2*>(a.out)[p] = reinterpret_cast<const double2*>(a.in)[r];
    } else if (a.mode == 1) {
        const double2 v = reinterpret_cast<const double2*>(a.in)[r];
        const double ab = hypot(v.x, v.y);  // abs(p) ** 2, slos.py:74
        a.out[p] = ab * ab;
    } else {
        a.out[p] = a.in[r];
    }
}

// |a|^2 in place order (colex): probs[r] = hypot(re, im)^2
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
