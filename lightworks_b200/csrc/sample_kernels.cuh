// sample_kernels.cuh -- on-device sampling from a probability vector.
//
// Replaces SamplerRunner._gen_samples_from_dist (lightworks/emulator/simulation/sampler.py:329-349), i.e.
//     rng.choice(range(len(dist)), p=probs, size=N)
// whose numpy implementation is  cdf = p.cumsum(); cdf /= cdf[-1]; idx = cdf.searchsorted(rng.random(N), 'right').
// The uniform variates stay the caller's (numpy's Generator on the host: N doubles), so a given seed draws the
// same samples as the reference unless a variate falls within rounding distance of a CDF edge (the parallel
// scan adds in a different order than numpy's sequential cumsum).  The distribution itself never leaves the GPU.
//
// States the reference drops from its dict (p <= sampler_probability_threshold, permanent.py:151 / slos.py:74)
// enter the CDF as zeros, which leaves every other state's interval exactly where the reference has it.
//
//  scan_tile_sums_kernel  : sum of every tile of SCAN_TILE thresholded probabilities          (read S doubles)
//  scan_tile_offsets_kernel: exclusive scan of the tile sums, one block
//  scan_tile_cdf_kernel   : inclusive scan inside every tile + tile offset -> cdf             (read S, write S)
//  sample_search_kernel   : first i with cdf[i] / cdf[S-1] > u, one thread per variate (binary search)
// All HBM-bound: 24 bytes per state for the CDF, ~27 dependent 8-byte probes per sample.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fock.cuh"

namespace lwb {

// ---------------------------------------------------------------------------------------------------------
// Output-state selection of SamplerRunner._sample_N_outputs (simulation/sampler.py:275-327) as a device mask over
// the full basis: a state keeps its probability iff, after optional threshold detection (s_i -> min(s_i, 1),
// :288-290), every heralded mode holds its required photon number (:292-294), the herald-free state holds at least
// min_detection photons (:299-300) and every post-selection rule "photons in these modes is one of these numbers"
// holds (sdk/utils/post_selection.py:105-138; rule modes already mapped to full-circuit modes by the caller).
// Sampling the masked full-state distribution and then mapping each drawn state (threshold, herald removal) on the
// host has the same law as the reference's sampling of the merged dict.
constexpr int MASK_MAX_RULES = 16;
struct SelectArgs {
    const uint64_t* bin;
    int m, n;               // modes, photons
    uint64_t S;
    int order;              // 0: fock_basis (colex) ranks, 1: SLOS (lex) ranks
    int threshold_detect;   // detector without photon counting
    int min_detection;
    int n_heralds;
    uint8_t herald_mode[LWB_MAX_MODES];
    uint8_t herald_count[LWB_MAX_MODES];
    int n_rules;
    unsigned long long rule_modes[MASK_MAX_RULES][2];    // bit j: mode j belongs to the rule (m <= 128)
    unsigned long long rule_counts[MASK_MAX_RULES];      // bit c: c photons allowed (c < 64)
    double* probs;          // in place: excluded states are set to 0
};

template <int KMAX>
static __global__ void __launch_bounds__(256) select_outputs_kernel(SelectArgs a) {
    const uint64_t r = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (r >= a.S) return;
    int x[KMAX];
    if (a.order == 0) colex_unrank(a.bin, r, a.n, a.m, x);
    else lex_unrank(a.bin, r, a.n, a.m, x);
    // runs of equal modes = (mode, occupation) pairs
    bool ok = true;
    int detected = 0;                    // photons seen outside the heralded modes
    int rule_sum[MASK_MAX_RULES];
#pragma unroll
    for (int q = 0; q < MASK_MAX_RULES; ++q) rule_sum[q] = 0;
    unsigned long long seen_herald = 0;  // heralded modes found occupied (index into the herald list)
    for (int i = 0; i < a.n;) {
        int j = i;
        while (j < a.n && x[j] == x[i]) ++j;
        const int mode = x[i];
        const int occ = a.threshold_detect ? 1 : (j - i);
        int hidx = -1;
        for (int hq = 0; hq < a.n_heralds; ++hq)
            if (a.herald_mode[hq] == mode) hidx = hq;
        if (hidx >= 0) {
            if (occ != a.herald_count[hidx]) ok = false;
            seen_herald |= 1ull << (hidx & 63);
        } else {
            detected += occ;
        }
        for (int q = 0; q < a.n_rules; ++q)
            if ((a.rule_modes[q][mode >> 6] >> (mode & 63)) & 1ull) rule_sum[q] += occ;
        i = j;
    }
    for (int hq = 0; hq < a.n_heralds; ++hq)  // heralded modes that stayed empty must be required empty
        if (!((seen_herald >> (hq & 63)) & 1ull) && a.herald_count[hq] != 0) ok = false;
    if (detected < a.min_detection) ok = false;
    for (int q = 0; q < a.n_rules; ++q)
        if (rule_sum[q] > 63 || !((a.rule_counts[q] >> rule_sum[q]) & 1ull)) ok = false;
    if (!ok) a.probs[r] = 0.0;
}

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ double keep_above(double p, double threshold) { return p > threshold ? p : 0.0; }

__device__ __forceinline__ double block_sum_256(double v, double* red /* [8] shared */) {
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) t += red[w];
    return t;
}

static __global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums_kernel(const double* __restrict__ p, uint64_t S,
                                                                              double threshold,
                                                                              double* __restrict__ tile_sum) {
    __shared__ double red[SCAN_THREADS / 32];
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE;
    double v = 0.0;
#pragma unroll
    for (int it = 0; it < SCAN_ITEMS; ++it) {
        const uint64_t i = base + (uint64_t)it * SCAN_THREADS + threadIdx.x;
        if (i < S) v += keep_above(p[i], threshold);
    }
    const double t = block_sum_256(v, red);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = t;
}

// tile_off[t] = sum of tile sums before t, tile_off[n_tiles] = total; one block of 1024 threads
static __global__ void __launch_bounds__(1024) scan_tile_offsets_kernel(const double* __restrict__ tile_sum,
                                                                         uint64_t n_tiles, double* __restrict__ tile_off) {
    __shared__ double part[1024];
    const uint64_t chunk = (n_tiles + 1023) / 1024;
    const uint64_t b = threadIdx.x * chunk, e = b + chunk < n_tiles ? b + chunk : n_tiles;
    double sum = 0.0;
    for (uint64_t i = b; i < e; ++i) sum += tile_sum[i];
    part[threadIdx.x] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        const double v = (int)threadIdx.x >= off ? part[threadIdx.x - off] : 0.0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    double run = part[threadIdx.x] - sum;
    for (uint64_t i = b; i < e; ++i) {
        const double v = tile_sum[i];
        tile_off[i] = run;
        run += v;
    }
    if (threadIdx.x == 1023) tile_off[n_tiles] = part[1023];
}

static __global__ void __launch_bounds__(SCAN_THREADS) scan_tile_cdf_kernel(const double* __restrict__ p, uint64_t S,
                                                                             double threshold,
                                                                             const double* __restrict__ tile_off,
                                                                             double* __restrict__ cdf) {
    __shared__ double tile[SCAN_TILE + SCAN_TILE / 32];  // padded: thread t owns [t*16, t*16+16)
    __shared__ double wsum[SCAN_THREADS / 32];
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE;
    auto slot = [](int e) { return e + (e >> 5); };
#pragma unroll
    for (int it = 0; it < SCAN_ITEMS; ++it) {
        const int e = it * SCAN_THREADS + threadIdx.x;
        const uint64_t i = base + e;
        tile[slot(e)] = i < S ? keep_above(p[i], threshold) : 0.0;
    }
    __syncthreads();
    double loc[SCAN_ITEMS];
    double run = 0.0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        run += tile[slot(threadIdx.x * SCAN_ITEMS + k)];
        loc[k] = run;
    }
    // exclusive prefix of the thread totals: warp scan, then the warp totals
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double incl = run;
    for (int s = 1; s < 32; s <<= 1) {
        const double v = __shfl_up_sync(0xffffffffu, incl, s);
        if (lane >= s) incl += v;
    }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    double wbase = 0.0;
    for (int ww = 0; ww < w; ++ww) wbase += wsum[ww];
    const double excl = tile_off[blockIdx.x] + (wbase + (incl - run));
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) tile[slot(threadIdx.x * SCAN_ITEMS + k)] = excl + loc[k];
    __syncthreads();
#pragma unroll
    for (int it = 0; it < SCAN_ITEMS; ++it) {
        const int e = it * SCAN_THREADS + threadIdx.x;
        const uint64_t i = base + e;
        if (i < S) cdf[i] = tile[slot(e)];
    }
}

// numpy: cdf /= cdf[-1]; searchsorted(cdf, u, side='right') = number of entries <= u = first i with cdf[i] > u.
// A distribution that sums to zero gives index S for every sample (the host entry points report the error).
static __global__ void sample_search_kernel(const double* __restrict__ cdf, uint64_t S, const double* __restrict__ u,
                                            uint64_t n, uint64_t* __restrict__ idx) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double total = cdf[S - 1];
    const double x = u[k];
    uint64_t lo = 0, hi = S;  // invariant: cdf[i]/total <= x for i < lo, > x for i >= hi
    if (!(total > 0.0)) lo = S;
    while (lo < hi) {
        const uint64_t mid = lo + ((hi - lo) >> 1);
        if (cdf[mid] / total > x) hi = mid; else lo = mid + 1;
    }
    idx[k] = lo;
}

}  // namespace lwb
