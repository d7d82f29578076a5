// fock.cuh -- exact integer Fock-state indexing shared by host and device code.
//
// Replaces lightworks/emulator/utils/state.py:22-35 (fock_basis/_sums) with closed-form
// rank/unrank so states can be enumerated on-device and sharded by rank range.
//
// A Fock state of n photons in m modes is held as its ascending "mode list"
// x[0] <= x[1] <= ... <= x[n-1] (mode index of every photon).  Two orders matter:
//
//  * REFERENCE (permanent) order = order produced by _sums(): outer loop on the LAST
//    mode's occupation.  On mode lists this is colexicographic order, whose rank is the
//    combinatorial number system:   rank = sum_{i=1..n} C(x[i-1] + i - 1, i).
//  * SLOS order = dict insertion order of slos.py:96-103 = lexicographic order on the
//    ascending mode list.
//
// Binomials come from a table bin[a*LWB_BIN_B + b] = C(a,b) (saturating at 2^64-1).
#pragma once
#include <stdint.h>

#ifndef __CUDACC__
#define __host__
#define __device__
#endif

#define LWB_BIN_A 160 /* a < 160 : m + n - 1 <= 159 */
#define LWB_BIN_B 48  /* b < 48  : n <= 47          */
#define LWB_MAX_MODES 128
#define LWB_MAX_PHOTONS 40

/* Debug build (make TAG=bounds EXTRA=-DLWB_DEBUG_BOUNDS): device-side index asserts in the kernels that compute their
 * own addresses.  compute-sanitizer is closed on the measurement pool, so tools/bounds_check.sh runs the small
 * configurations (tools/sanitize_small.py) against this build instead. */
#ifdef LWB_DEBUG_BOUNDS
#include <assert.h>
#define LWB_BOUNDS(cond) assert(cond)
#else
#define LWB_BOUNDS(cond) ((void)0)
#endif

#define LWB_MAX_PEERS 15 /* extra destinations of a probability store: the other GPUs of one NVLink domain */

namespace lwb {

// Extra destinations of every probability a kernel stores: the same element of the peers' buffers (device
// pointers into peer memory over NVLink, already offset like the local pointer).  n == 0: local store only.
struct MirrorSet {
    int n;
    double* p[LWB_MAX_PEERS];
};

// Host-side table fill (exact until saturation).
inline void fill_binomials(uint64_t* bin) {
    for (int a = 0; a < LWB_BIN_A; ++a)
        for (int b = 0; b < LWB_BIN_B; ++b) {
            uint64_t v;
            if (b == 0) v = 1;
            else if (a == 0) v = 0;
            else {
                uint64_t p = bin[(a - 1) * LWB_BIN_B + b - 1], q = bin[(a - 1) * LWB_BIN_B + b];
                v = (p > UINT64_MAX - q) ? UINT64_MAX : p + q;
            }
            bin[a * LWB_BIN_B + b] = v;
        }
}

__host__ __device__ inline uint64_t binom(const uint64_t* bin, int a, int b) {
    return (a < 0 || b < 0) ? 0 : bin[a * LWB_BIN_B + b];
}

// occupation vector (m entries) -> ascending mode list; returns n
__host__ __device__ inline int state_to_modes(const uint8_t* s, int m, int* x) {
    int n = 0;
    for (int i = 0; i < m; ++i)
        for (int k = 0; k < s[i]; ++k) x[n++] = i;
    return n;
}

__host__ __device__ inline void modes_to_state(const int* x, int n, int m, uint8_t* s) {
    for (int i = 0; i < m; ++i) s[i] = 0;
    for (int i = 0; i < n; ++i) s[x[i]] += 1;
}

// ---- reference (colex) order -------------------------------------------------
__host__ __device__ inline uint64_t colex_rank(const uint64_t* bin, const int* x, int n) {
    uint64_t r = 0;
    for (int i = 1; i <= n; ++i) r += binom(bin, x[i - 1] + i - 1, i);
    return r;
}

// Greedy unrank: for i = n..1 pick the largest y with C(y,i) <= rank; x[i-1] = y-i+1.
__host__ __device__ inline void colex_unrank(const uint64_t* bin, uint64_t rank, int n, int m, int* x) {
    // y_i = x[i-1] + i - 1 is strictly increasing in i; C(i-1, i) = 0 so the scan always stops.
    int y = m + n - 2;  // largest possible y for i = n is (m-1)+n-1
    for (int i = n; i >= 1; --i) {
        while (binom(bin, y, i) > rank) --y;
        rank -= binom(bin, y, i);
        x[i - 1] = y - i + 1;
        y -= 1;
    }
}

// Successor in colex order; returns false when x was the last state.
__host__ __device__ inline bool colex_next(int* x, int n, int m) {
    for (int p = 0; p < n; ++p) {
        int lim = (p == n - 1) ? m - 1 : x[p + 1];
        if (x[p] < lim) {
            x[p] += 1;
            for (int q = 0; q < p; ++q) x[q] = 0;
            return true;
        }
    }
    return false;
}

// ---- SLOS (lexicographic) order ------------------------------------------------
// rank = #lists lexicographically smaller.  Lists with prefix x[0..i-1] and i-th element v
// (x[i-1] <= v < x[i]) followed by anything >= v:  C((m - v) + r - 1, r), r = n-1-i free slots
// after position i ... summed over v by the hockey-stick identity.
__host__ __device__ inline uint64_t lex_rank(const uint64_t* bin, const int* x, int n, int m) {
    uint64_t r = 0;
    int prev = 0;
    for (int i = 0; i < n; ++i) {
        int rem = n - 1 - i;  // elements after position i
        // sum_{v=prev}^{x[i]-1} C(m - v + rem - 1, rem) = C(m - prev + rem, rem + 1) - C(m - x[i] + rem, rem + 1)
        r += binom(bin, m - prev + rem, rem + 1) - binom(bin, m - x[i] + rem, rem + 1);
        prev = x[i];
    }
    return r;
}

// Successor in SLOS (lexicographic) order; returns false when x was the last state.
__host__ __device__ inline bool lex_next(int* x, int n, int m) {
    for (int i = n - 1; i >= 0; --i) {
        if (x[i] < m - 1) {
            const int v = x[i] + 1;
            for (int q = i; q < n; ++q) x[q] = v;
            return true;
        }
    }
    return false;
}

__host__ __device__ inline void lex_unrank(const uint64_t* bin, uint64_t rank, int n, int m, int* x) {
    int prev = 0;
    for (int i = 0; i < n; ++i) {
        int rem = n - 1 - i;
        int v = prev;
        for (;;) {
            uint64_t cnt = binom(bin, m - v + rem - 1, rem);  // lists continuing with element v here
            if (rank < cnt) break;
            rank -= cnt;
            ++v;
        }
        x[i] = v;
        prev = v;
    }
}

}  // namespace lwb
