// misc_kernels.cuh -- on-device Fock enumeration and the FP64 peak micro-benchmark.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fock.cuh"

namespace lwb {

// States [r0, r0+cnt) of the reference's fock_basis(m, n) order (emulator/utils/state.py:22-35):
// each thread unranks the first state of its slice and walks the colex successor.
static __global__ void fock_basis_kernel(const uint64_t* __restrict__ bin, int m, int n, uint64_t r0, uint64_t cnt,
                                  uint64_t per_thread, uint8_t* __restrict__ states) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t k = t * per_thread;
    if (k >= cnt) return;
    const uint64_t k_end = (k + per_thread < cnt) ? k + per_thread : cnt;
    int x[LWB_MAX_PHOTONS + 1];
    colex_unrank(bin, r0 + k, n, m, x);
    for (; k < k_end; ++k) {
        uint8_t* s = states + k * (uint64_t)m;
        for (int i = 0; i < m; ++i) s[i] = 0;
        for (int i = 0; i < n; ++i) s[x[i]] += 1;
        colex_next(x, n, m);
    }
}

// Deterministic sum of `count` doubles: block b adds the contiguous slice [b, b+1) * ceil(count / grid) with a fixed
// thread-strided order and a fixed shared-memory tree; a second launch (grid 1) adds the block partials the same way.
constexpr int SUM_THREADS = 256;
static __global__ void __launch_bounds__(SUM_THREADS) sum_slices_kernel(const double* __restrict__ x, uint64_t count,
                                                                       double* __restrict__ partials) {
    __shared__ double sh[SUM_THREADS];
    const uint64_t per = (count + gridDim.x - 1) / gridDim.x;
    const uint64_t b = per * blockIdx.x, e = (b + per < count) ? b + per : count;
    double acc = 0.0;
    for (uint64_t i = b + threadIdx.x; i < e; i += SUM_THREADS) acc += x[i];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = SUM_THREADS / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partials[blockIdx.x] = sh[0];
}

// FP64 roofline denominator: 8 independent DFMA chains per thread, no memory traffic.
constexpr int DFMA_PER_ITER = 8;
static __global__ void dfma_peak_kernel(double* out, int iters, double a) {
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    const double b = 1e-9;
#pragma unroll 8
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 123.456) out[0] = s;  // keep the chains alive
}

// FP64 pipe issue-rate probes (register-operand patterns).  Each thread runs 8 independent dependency
// chains; `mode` picks the instruction pattern.  Rates are reported as if every instruction were a
// DFMA (2 flops), so 1.0 x lwb_fp64_peak = one FP64 warp-instruction per 2 cycles per scheduler.
//   1: DFMA x = a_i * b_i + x      (3 distinct register operands)
//   2: DFMA x = a   * b_i + x      (operand A shared by consecutive instructions: reuse cache)
//   3: DMUL x = x * a_i            (2 register operands)
//   4: DADD x = x + a_i            (2 register operands)
//   5: complex multiply chain p *= v_i written as in glynn.cuh (2 DMUL + 2 DFMA)
template <int MODE>
static __global__ void fp64_probe_kernel(double* out, int iters, const double* __restrict__ src) {
    double a[8], b[8], x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {  // lane-dependent so nothing is promoted to uniform registers
        a[i] = src[i] + threadIdx.x * 1e-9; b[i] = src[8 + i] + threadIdx.x * 1e-12; x[i] = src[16 + i] + threadIdx.x * 1e-7;
    }
    const double a0 = a[0];
#pragma unroll 4
    for (int it = 0; it < iters; ++it) {
        if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = fma(a[i], b[i], x[i]);
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = fma(a0, b[i], x[i]);
        } else if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = x[i] * a[i];
        } else if (MODE == 4) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = x[i] + a[i];
        } else {
            // 4 independent complex products (x[2c], x[2c+1]) *= (a[2c], b[2c]) then *= (a[2c+1], b[2c+1])
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                double pr = x[2 * c], pi = x[2 * c + 1];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const double ar = a[2 * c + k], ai = b[2 * c + k];
                    double nr = pr * ar, ni = pr * ai;
                    nr = fma(-pi, ai, nr);
                    ni = fma(pi, ar, ni);
                    pr = nr;
                    pi = ni;
                }
                x[2 * c] = pr;
                x[2 * c + 1] = pi;
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 123.456) out[0] = s;
}

}  // namespace lwb
