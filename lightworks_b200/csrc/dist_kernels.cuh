// dist_kernels.cuh -- dense output distributions on the device and their "merge-convolution"
// (SURVEY.md section 8f rank 1).
//
// Replaces the nested dict loops of annotated_state_pdist_calc
// (lightworks/emulator/simulation/probability_distribution.py:135-164):
//     for output, p1 in pdist.items():
//         for result, p2 in unique_results[d]:
//             new_pdist[output.merge(result)] += p1 * p2          # State.merge = mode-wise sum of occupations
// i.e. the distribution of the SUM of two independent occupation vectors (photons that carry different labels do
// not interfere, so their output distributions convolve).
//
// Layout: a distribution over the measured modes is a dense float64 vector over the key space
//     K(n_modes, kmax) = all occupation vectors of n_modes modes with 0..kmax photons,
//     index(s) = off[k] + colex_rank(s),  k = photons of s,  off[k] = sum_{j<k} C(n_modes + j - 1, j),
// (the photon-number blocks of lossy circuits side by side, each in the reference's fock_basis order), so
// K(n_modes, k) is a prefix of K(n_modes, k') for k < k' and `acc += alpha * x` is a prefix axpy.
//
// convolve_kernel is the GATHER form: one thread per output state t walks the sub-multisets a of t that can be
// non-zero in A (by photon number; the smaller part is enumerated as index combinations of the sorted photon list),
// ranks a and t - a, and accumulates A[a] * B[t - a].  Same products as the reference's double loop, but no atomics,
// a fixed summation order, and no State objects.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fock.cuh"

namespace lwb {

constexpr int DIST_THREADS = 256;

struct ConvolveArgs {
    const uint64_t* bin;
    int n_modes;
    int ka_min, ka_max;  // photon-number blocks of A that can be non-zero
    int kb_min, kb_max;
    const double* A;     // over K(n_modes, ka_max)
    const double* B;     // over K(n_modes, kb_max)
    double* out;         // over K(n_modes, ka_max + kb_max)
    uint64_t size_out;
    uint64_t off[LWB_MAX_PHOTONS + 2];  // off[k] for k = 0 .. ka_max + kb_max + 1
};

template <int KMAX>
__global__ void __launch_bounds__(DIST_THREADS) convolve_kernel(ConvolveArgs a) {
    const uint64_t o = (uint64_t)blockIdx.x * DIST_THREADS + threadIdx.x;
    if (o >= a.size_out) return;
    int kt = 0;
    while (o >= a.off[kt + 1]) ++kt;
    const int ka_lo = max(a.ka_min, kt - a.kb_max), ka_hi = min(a.ka_max, kt - a.kb_min);
    if (ka_lo > ka_hi) { a.out[o] = 0.0; return; }
    int x[KMAX], c[KMAX];
    colex_unrank(a.bin, o - a.off[kt], kt, a.n_modes, x);
    double acc = 0.0;
    for (int ka = ka_lo; ka <= ka_hi; ++ka) {
        // sub-multisets of t with ka photons <-> index combinations c_0 < ... < c_{q-1} of the sorted photon list that
        // take the copies of a mode from the left ("canonical"); the smaller of the two parts is enumerated
        const int kb = kt - ka;
        const bool pick_a = ka <= kb;
        const int q = pick_a ? ka : kb;
        for (int j = 0; j < q; ++j) c[j] = j;
        for (;;) {
            bool canonical = true;
            for (int j = 0; j < q; ++j)
                if (c[j] > 0 && x[c[j]] == x[c[j] - 1] && (j == 0 || c[j - 1] != c[j] - 1)) { canonical = false; break; }
            if (canonical) {
                // colex ranks of the picked photons and of the rest: sum over photons, ascending, of C(mode + i - 1, i)
                uint64_t r_sel = 0, r_rest = 0;
                int is = 0, ir = 1;
                for (int p = 0; p < kt; ++p) {
                    if (is < q && c[is] == p) { ++is; r_sel += binom(a.bin, x[p] + is - 1, is); }
                    else { r_rest += binom(a.bin, x[p] + ir - 1, ir); ++ir; }
                }
                const uint64_t ra = pick_a ? r_sel : r_rest, rb = pick_a ? r_rest : r_sel;
                acc += a.A[a.off[ka] + ra] * a.B[a.off[kb] + rb];
            }
            int j = q - 1;  // next combination in lexicographic order
            while (j >= 0 && c[j] == kt - q + j) --j;
            if (j < 0) break;
            ++c[j];
            for (int l = j + 1; l < q; ++l) c[l] = c[l - 1] + 1;
        }
    }
    a.out[o] = acc;
}

// acc[i] += alpha * x[i] over the (prefix) key space of x
static __global__ void dist_axpy_kernel(double* __restrict__ acc, double alpha, const double* __restrict__ x, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) acc[i] += alpha * x[i];
}

// acc[i] += alpha_j * x_j[i] for j = 0 .. n-1 IN THAT ORDER (each x_j over its own prefix key space): the same sequence
// of operations per element as n dist_axpy_kernel launches, in one launch
struct DistAxpyTerm {
    const double* x;
    uint64_t n;
    double alpha;
};
static __global__ void dist_axpy_many_kernel(double* __restrict__ acc, const DistAxpyTerm* __restrict__ terms, int n_terms,
                                             uint64_t n_max) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_max) return;
    double a = acc[i];
    for (int j = 0; j < n_terms; ++j) {
        const DistAxpyTerm t = terms[j];
        if (i < t.n) a += t.alpha * t.x[i];
    }
    acc[i] = a;
}

// dst[i] = src[i] > threshold ? src[i] : 0   (strict, per full state: permanent.py:151, slos.py:74)
static __global__ void dist_threshold_copy_kernel(double* __restrict__ dst, const double* __restrict__ src, double threshold,
                                                  uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const double p = src[i];
        dst[i] = p > threshold ? p : 0.0;
    }
}

// permanent.py:159-161: vacuum = 1 - sum when the marginalised entries sum to less than 1.  `sum` = tile_off[n_tiles]
// of the scan over entries [1, size).
static __global__ void dist_vacuum_kernel(double* dist, const double* sum) {
    if (threadIdx.x == 0 && blockIdx.x == 0) dist[0] = *sum < 1.0 ? 1.0 - *sum : 0.0;
}

}  // namespace lwb
