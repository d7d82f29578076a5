// glynn.cuh -- the Glynn / BBFG permanent inner loop, one aligned Gray-code chunk per lane.
//
// Replaces thewalrus.perm (call site lightworks/emulator/backends/permanent.py:81) and the
// np.ix_ gather of `partition` (permanent.py:166-178): rows are never gathered into an n x n
// copy, they are read in place from a shared-memory row table through a per-permanent
// row-index list (the occupation-expanded output modes).
//
//   perm(A) = 2^-(n-1) * sum_{delta in {+-1}^n, delta_{n-1}=+1} (prod_i delta_i) * prod_j ( sum_i delta_i a_ij )
//
// The 2^(n-1) sign vectors are walked in Gray order g = 0..2^(n-1)-1, bits of g^(g>>1) = rows
// with delta=-1.  The space is cut into aligned chunks of 2^C steps; chunk q covers
// g in [q*2^C, (q+1)*2^C).  Inside a chunk the flipped row at step t -> t+1 is ctz(t+1), the same
// for every chunk, so when the lanes of a warp work on chunks of the SAME permanent every row
// read is a shared-memory broadcast; only the update sign of row C-1 (parity of q) and the
// start vector differ per lane.
//
// Scaling: the lane keeps w_j = (1/2) sum_i delta_i a_ij, so a flip is w_j +- a_kj with the
// matrix stored unscaled (no pre-doubling pass) and perm = 2 * sum_g sign_g prod_j w_j.  All
// scalings are powers of two, i.e. exact.
//
// FP64 instruction count per Gray step: 2N (adds) + 4(N-1) (complex multiplies) + 2 = 6N-2;
// the roofline model counts it as 8N-4 flops.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lwb {

__host__ __device__ constexpr int cx_ctz(int v) { int k = 0; while (!((v >> k) & 1)) ++k; return k; }

template <typename T> struct vec2;
template <> struct vec2<double> { using type = double2; };
template <> struct vec2<float> { using type = float2; };

// Complex multiply-accumulate forms (all 2 mul + 2 fma; they differ in which operands the instruction pairs share,
// i.e. in the operand-reuse hits ptxas can give the 3-operand DFMAs - see profiles/r02_fp64_reuse_study.md).
#ifndef LWB_CMUL_FORM
#define LWB_CMUL_FORM 0
#endif
template <typename T>
__device__ __forceinline__ void cmul(T& pr, T& pi, T ar, T ai) {
#if LWB_CMUL_FORM == 0
    // (pr + i pi) * (ar + i ai): the mul pair shares pr, the fma pair shares pi
    T nr = pr * ar;
    T ni = pr * ai;
    nr = fma(-pi, ai, nr);
    ni = fma(pi, ar, ni);
#elif LWB_CMUL_FORM == 1
    // mul pair shares pi, fma pair shares pr
    T t1 = pi * ai;
    T t2 = pi * ar;
    T nr = fma(pr, ar, -t1);
    T ni = fma(pr, ai, t2);
#elif LWB_CMUL_FORM == 2
    // pairs share the FACTOR: (mul, fma) on ar, (mul, fma) on ai
    T nr = pr * ar;
    T ni = pi * ar;
    nr = fma(-pi, ai, nr);
    ni = fma(pr, ai, ni);
#else
    T ni = pr * ai;
    T nr = pr * ar;
    ni = fma(pi, ar, ni);
    nr = fma(-pi, ai, nr);
#endif
    pr = nr;
    pi = ni;
}

// product of w[0..N-1] as LWB_CHAINS interleaved chains (ILP = LWB_CHAINS, 4*LWB_CHAINS live temporaries)
#ifndef LWB_CHAINS
#define LWB_CHAINS 2
#endif
template <int N, typename T>
__device__ __forceinline__ void cprod(const T (&wr)[N], const T (&wi)[N], T& pr, T& pi) {
#ifdef LWB_PROD_TREE
    // pairwise tree: N-1 multiplies like the chains, depth ceil(log2 N)
    T cr[N], ci[N];
#pragma unroll
    for (int j = 0; j < N; ++j) { cr[j] = wr[j]; ci[j] = wi[j]; }
#pragma unroll
    for (int step = 1; step < N; step <<= 1) {
#pragma unroll
        for (int j = 0; j + step < N; j += 2 * step) cmul(cr[j], ci[j], cr[j + step], ci[j + step]);
    }
    pr = cr[0];
    pi = ci[0];
#else
    constexpr int K = (N >= 2 * LWB_CHAINS) ? LWB_CHAINS : 1;
    T cr[K], ci[K];
#pragma unroll
    for (int c = 0; c < K; ++c) {
        const int lo = (N * c) / K, hi = (N * (c + 1)) / K;
        cr[c] = wr[lo];
        ci[c] = wi[lo];
#pragma unroll
        for (int j = lo + 1; j < hi; ++j) cmul(cr[c], ci[c], wr[j], wi[j]);
    }
#pragma unroll
    for (int c = 1; c < K; ++c) cmul(cr[0], ci[0], cr[c], ci[c]);
    pr = cr[0];
    pi = ci[0];
#endif
}

// Shared-memory row reads go through volatile inline PTX: the row table is loop-invariant, and
// with plain loads the compiler hoists every row out of the Gray loop (LICM) and spills.
__device__ __forceinline__ double2 lds_c(const double2* p, int j) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y)
                 : "r"((uint32_t)__cvta_generic_to_shared(p + j)));
    return v;
}
__device__ __forceinline__ float2 lds_c(const float2* p, int j) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y)
                 : "r"((uint32_t)__cvta_generic_to_shared(p + j)));
    return v;
}
__device__ __forceinline__ int lds_i(const int* p) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)));
    return v;
}

template <int N, typename T, bool SUB>
__device__ __forceinline__ void row_addsub(T (&wr)[N], T (&wi)[N], const typename vec2<T>::type* __restrict__ row) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
        typename vec2<T>::type a = lds_c(row, j);
        if (SUB) { wr[j] -= a.x; wi[j] -= a.y; }
        else     { wr[j] += a.x; wi[j] += a.y; }
    }
}

template <int N, typename T>
__device__ __forceinline__ void row_fma(T (&wr)[N], T (&wi)[N], const typename vec2<T>::type* __restrict__ row, T s) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
        typename vec2<T>::type a = lds_c(row, j);
        wr[j] = fma(s, a.x, wr[j]);
        wi[j] = fma(s, a.y, wi[j]);
    }
}

// Dynamic flips (sign known only at run time).  A DFMA with the +-1.0 sign as third register operand costs 3 issue
// cycles on B200 (register-file bandwidth), a DADD 2.  Round 2: the row table is followed by its NEGATED copy
// (neg_off elements further), so a flip with a run-time sign is a plain add from one of the two copies - the sign only
// selects the address.  w + (-a) rounds exactly like fma(-1, a, w), so results are bit-identical to the DFMA form.
// (Branching to the add / subtract forms instead was measured SLOWER in round 1: it splits the unrolled block.)
template <int N, typename T>
__device__ __forceinline__ void row_sgn(T (&wr)[N], T (&wi)[N], const typename vec2<T>::type* __restrict__ row, bool negative,
                                        int neg_off) {
    row_addsub<N, T, false>(wr, wi, row + (negative ? neg_off : 0));
}

// One chunk of 2^C Gray steps.
//   rows   : shared-memory row table (row r starts at rows + r*N)
//   rowidx : shared memory, N ints: rowidx[i] = (table row of matrix row i) * N
//   q      : chunk index
// Adds sign-corrected  sum_{t} (+-) prod_j w_j  of this chunk to (acc_re, acc_im).
template <int N, int C, int U, typename T, bool NEG = true>
__device__ __forceinline__ void glynn_chunk(const typename vec2<T>::type* __restrict__ rows,
                                            const int* __restrict__ rowidx, uint32_t q, T& acc_re, T& acc_im, int neg_off) {
    auto flip = [&](T (&wr_)[N], T (&wi_)[N], const typename vec2<T>::type* r_, bool negative) {
        if constexpr (NEG) row_sgn<N, T>(wr_, wi_, r_, negative, neg_off);
        else row_fma<N, T>(wr_, wi_, r_, negative ? T(-1) : T(1));
    };
    static_assert(U <= C && C <= N - 1, "bad chunk shape");
    using V = typename vec2<T>::type;
    T wr[N], wi[N];
    const uint64_t g0 = (uint64_t)q << C;
    const uint64_t gr0 = g0 ^ (g0 >> 1);

    // ---- start vector w = 1/2 sum_i delta_i a_i ----
    {
        const V* r = rows + lds_i(rowidx + N - 1);  // delta_{N-1} = +1
#pragma unroll
        for (int j = 0; j < N; ++j) { V a = lds_c(r, j); wr[j] = a.x; wi[j] = a.y; }
    }
#pragma unroll
    for (int i = 0; i < N - 1; ++i) {
        const V* r = rows + lds_i(rowidx + i);
        if (i < C - 1) row_addsub<N, T, false>(wr, wi, r);  // low gray bits of an aligned chunk start are 0
        else flip(wr, wi, r, ((gr0 >> i) & 1) != 0);
    }
#pragma unroll
    for (int j = 0; j < N; ++j) { wr[j] *= T(0.5); wi[j] *= T(0.5); }

    // row pointers of the statically flipped low rows
    const V* lowrow[U > 0 ? U : 1];
#pragma unroll
    for (int k = 0; k < U; ++k) lowrow[k] = rows + lds_i(rowidx + k);

    T sr = T(0), si = T(0);
    constexpr uint32_t NBLK = 1u << (C - U);
    constexpr int BLK = 1 << U;
#pragma unroll 1
    for (uint32_t ob = 0; ob < NBLK; ++ob) {
#pragma unroll
        for (int t = 0; t < BLK; ++t) {
            T pr, pi;
            cprod<N, T>(wr, wi, pr, pi);
            if ((t & 1) && BLK > 1) { sr -= pr; si -= pi; }
            else if (BLK > 1 || !(ob & 1)) { sr += pr; si += pi; }
            else { sr -= pr; si -= pi; }  // BLK == 1: parity of ob
            if (t + 1 < BLK) {
                const int t1 = t + 1;
                const int k = cx_ctz(t1);
                if (k <= U - 2) {
                    // new gray bit k = bit k ^ bit k+1 of t1, both inside the block: static
                    const bool newbit = ((t1 >> k) ^ (t1 >> (k + 1))) & 1;
                    if (newbit) row_addsub<N, T, true>(wr, wi, lowrow[k]);
                    else row_addsub<N, T, false>(wr, wi, lowrow[k]);
                } else {
                    // k == U-1 (middle of the block): bit U of the global index decides.  (A branch to the add/sub
                    // forms for the warp-uniform case was measured SLOWER here: it splits the unrolled block.)
                    const uint32_t hi = (U == C) ? (q & 1u) : (ob & 1u);
                    flip(wr, wi, lowrow[k], !hi);  // newbit = 1 ^ hi
                }
            }
        }
        // block boundary: flip row k = U + ctz(ob+1) unless the chunk ends
        if (ob + 1 < NBLK) {
            const uint32_t o1 = ob + 1;
            const int k = U + (__ffs(o1) - 1);
            uint32_t newbit;
            if (k == C - 1) newbit = 1u ^ (q & 1u);
            else newbit = ((o1 >> (k - U)) ^ (o1 >> (k - U + 1))) & 1u;
            flip(wr, wi, rows + lds_i(rowidx + k), newbit != 0);
        }
    }
    const T s0 = (__popcll(gr0) & 1) ? T(-1) : T(1);
    acc_re = fma(s0, sr, acc_re);
    acc_im = fma(s0, si, acc_im);
}

// A contiguous RANGE of chunks [q_begin, q_begin + n_chunks) walked by one lane without re-deriving the
// start vector: inside a chunk the flips are the uniform ruler sequence (broadcast row reads, as in
// glynn_chunk); between chunk q and q+1 the Gray code flips row C + ctz(q+1), which differs per lane and is
// read with a lane-specific address once every 2^C steps.  Term signs: parity(popcount(gray(g))) = g & 1,
// so the sign alternates with the step index and needs no per-chunk correction (C >= 1).
// Used by the single-large-permanent kernel: every lane of the GPU owns an equal share of the Gray space
// regardless of how 2^(n-1) divides by the lane count (no tail quantisation, one setup per lane).
template <int N, int C, int U, typename T, bool NEG = true>
__device__ __forceinline__ void glynn_range(const typename vec2<T>::type* __restrict__ rows,
                                            const int* __restrict__ rowidx, uint64_t q_begin, uint64_t n_chunks,
                                            T& acc_re, T& acc_im, int neg_off) {
    auto flip = [&](T (&wr_)[N], T (&wi_)[N], const typename vec2<T>::type* r_, bool negative) {
        if constexpr (NEG) row_sgn<N, T>(wr_, wi_, r_, negative, neg_off);
        else row_fma<N, T>(wr_, wi_, r_, negative ? T(-1) : T(1));
    };
    static_assert(U <= C && C <= N - 1 && C >= 1, "bad chunk shape");
    using V = typename vec2<T>::type;
    if (n_chunks == 0) return;
    T wr[N], wi[N];
    const uint64_t g0 = q_begin << C;
    const uint64_t gr0 = g0 ^ (g0 >> 1);
    {
        const V* r = rows + lds_i(rowidx + N - 1);
#pragma unroll
        for (int j = 0; j < N; ++j) { V a = lds_c(r, j); wr[j] = a.x; wi[j] = a.y; }
    }
#pragma unroll
    for (int i = 0; i < N - 1; ++i) {
        const V* r = rows + lds_i(rowidx + i);
        if (i < C - 1) row_addsub<N, T, false>(wr, wi, r);
        else flip(wr, wi, r, ((gr0 >> i) & 1) != 0);
    }
#pragma unroll
    for (int j = 0; j < N; ++j) { wr[j] *= T(0.5); wi[j] *= T(0.5); }
    const V* lowrow[U > 0 ? U : 1];
#pragma unroll
    for (int k = 0; k < U; ++k) lowrow[k] = rows + lds_i(rowidx + k);

    T sr = T(0), si = T(0);
    constexpr uint32_t NBLK = 1u << (C - U);
    constexpr int BLK = 1 << U;
#pragma unroll 1
    for (uint64_t c = 0; c < n_chunks; ++c) {
        const uint64_t q = q_begin + c;
#pragma unroll 1
        for (uint32_t ob = 0; ob < NBLK; ++ob) {
#pragma unroll
            for (int t = 0; t < BLK; ++t) {
                T pr, pi;
                cprod<N, T>(wr, wi, pr, pi);
                if ((t & 1) && BLK > 1) { sr -= pr; si -= pi; }
                else if (BLK > 1 || !(ob & 1)) { sr += pr; si += pi; }
                else { sr -= pr; si -= pi; }
                if (t + 1 < BLK) {
                    const int t1 = t + 1;
                    const int k = cx_ctz(t1);
                    if (k <= U - 2) {
                        const bool newbit = ((t1 >> k) ^ (t1 >> (k + 1))) & 1;
                        if (newbit) row_addsub<N, T, true>(wr, wi, lowrow[k]);
                        else row_addsub<N, T, false>(wr, wi, lowrow[k]);
                    } else {
                        const uint32_t hi = (U == C) ? (uint32_t)(q & 1u) : (ob & 1u);
                        flip(wr, wi, lowrow[k], !hi);
                    }
                }
            }
            if (ob + 1 < NBLK) {
                const uint32_t o1 = ob + 1;
                const int k = U + (__ffs(o1) - 1);
                uint32_t newbit;
                if (k == C - 1) newbit = 1u ^ (uint32_t)(q & 1u);
                else newbit = ((o1 >> (k - U)) ^ (o1 >> (k - U + 1))) & 1u;
                flip(wr, wi, rows + lds_i(rowidx + k), newbit != 0);
            }
        }
        if (c + 1 < n_chunks) {  // chunk boundary: lane-specific row
            const uint64_t q1 = q + 1;
            const int z = __ffsll((long long)q1) - 1;
            const int k = C + z;
            const uint32_t newbit = (uint32_t)((q1 >> z) ^ (q1 >> (z + 1))) & 1u;
            flip(wr, wi, rows + lds_i(rowidx + k), newbit != 0);
        }
    }
    acc_re += sr;
    acc_im += si;
}

}  // namespace lwb
