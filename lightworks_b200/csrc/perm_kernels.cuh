// perm_kernels.cuh -- batched and single-permanent kernels built on glynn_chunk.
//
//  K1s perm_states_kernel   : amplitudes <out|U|in> for inputs x outputs given as Fock states, or for
//                             a rank range of the reference's fock_basis order enumerated on-device.
//                             Replaces the loops at permanent.py:145-157, simulator.py:98-104,
//                             analyzer.py:120-151 (per-pair partition + perm + factorial scaling).
//  K1m perm_matrices_kernel : permanents of B dense n x n matrices (thewalrus.perm semantics).
//  K2  perm_single_kernel   : one large permanent, Gray space split over every lane of the GPU
//                             (and over ranks through [q_begin, q_end)).
#pragma once
#include "fock.cuh"
#include "glynn.cuh"

namespace lwb {

constexpr int K1_THREADS = 256;
#ifndef LWB_U
/* K1: statically unrolled low Gray bits.  Measured on B200 (profiles/r01_variant_sweep.md): 4 is best up to
 * n = 14, 3 beyond (register pressure); capping registers with __launch_bounds__ minBlocks only spills. */
#define LWB_U(N) ((N) <= 14 ? 4 : 3)
#endif
#ifndef LWB_U2
#define LWB_U2(N) ((N) >= 27 ? 1 : 2)  /* K2: measured best on B200 (register pressure grows with n) */
#endif
#ifndef LWB_MINB
#define LWB_MINB 1
#endif
#ifndef LWB_K2_COFF
#define LWB_K2_COFF 19  /* K2 chunk bits C = clamp(N - LWB_K2_COFF, 3, 7); 19 vs 21: n = 24 runs 13 % faster, others equal (profiles/r01_k2_chunk_sweep.log) */
#endif

// Work split shared by all kernels: LOG_L lanes-per-permanent exponent and chunk bits.
template <int N, int LOG_L>
struct K1Shape {
    static constexpr int L = 1 << LOG_L;
    static constexpr int C = N - 1 - LOG_L;          // steps per lane = 2^C
    static constexpr int U = C < LWB_U(N) ? C : LWB_U(N);  // statically unrolled low bits
    static constexpr int GPB = K1_THREADS / L;       // groups (permanents in flight) per block
};

// complex64 drift control.  The incrementally updated column sums w_j pick up a rounding error that grows like
// sqrt(steps) * eps * n; for float (eps 6e-8) a 2^14-step chunk would leave ~1.5e-4, above the 1e-4 the complex64
// variant promises.  A float lane therefore restarts from a freshly derived start vector every 2^LWB_F32_RESTART steps
// (cost: n^2 FMAs per restart, ~1 % at 2^9) and adds the sub-chunk sums in double.  Double keeps one start per chunk.
#ifndef LWB_F32_RESTART
#define LWB_F32_RESTART 9
#endif
// The negated copy of the row table (glynn.cuh row_sgn) pays from 10 photons on: below, a permanent has so few Gray
// steps that staging twice the shared memory per matrix costs more than the 3-cycle DFMAs it removes.
__host__ __device__ constexpr bool k1_neg(int n) { return n >= 10; }

template <int N, int C, int U, typename T>
__device__ __forceinline__ void glynn_chunk_acc(const typename vec2<T>::type* __restrict__ rows,
                                                const int* __restrict__ rowidx, uint32_t q, double& dr, double& di, int neg_off) {
    constexpr bool SPLIT = sizeof(T) == 4 && C > LWB_F32_RESTART;
    constexpr int R = SPLIT ? LWB_F32_RESTART : C;
    constexpr int UR = U < R ? U : R;
#pragma unroll 1
    for (uint32_t i = 0; i < (1u << (C - R)); ++i) {
        T sr = T(0), si = T(0);
        glynn_chunk<N, R, UR, T, k1_neg(N)>(rows, rowidx, (q << (C - R)) + i, sr, si, neg_off);
        dr += (double)sr;
        di += (double)si;
    }
}

struct PermStatesArgs {
    const double2* U;          // m x m complex128, row-major, device
    int m;
    const uint8_t* in_states;  // [n_in][m]
    const uint8_t* out_states; // [n_out][m] or nullptr -> basis mode
    uint64_t r0;               // basis mode: first rank
    uint64_t n_out;
    const uint64_t* bin;       // binomial table (device)
    double* out;               // [n_in][n_out] complex (2 doubles) or real
    int out_probs;             // 0: amplitudes, 1: |amp|^2
    MirrorSet mir;             // out_probs only: peers that receive every probability as it is produced
    // Batch across unitaries (blockIdx.z = task): task z uses U + z * u_stride (complex entries), inputs
    // in_states + z * in_stride (bytes) and writes out + z * out_stride (doubles); outputs are shared.  0 = no batch.
    uint64_t u_stride, in_stride, out_stride;
};

template <typename T>
__device__ __forceinline__ T group_sum(T v, int log_l) {
    for (int s = 0; s < log_l; ++s) v += __shfl_xor_sync(0xffffffffu, v, 1 << s);
    return v;
}

template <int N, int LOG_L, typename T>
__global__ void __launch_bounds__(K1_THREADS, LWB_MINB) perm_states_kernel(PermStatesArgs a) {
    using S = K1Shape<N, LOG_L>;
    using V = typename vec2<T>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NT = k1_neg(N) ? 2 : 1;
    V* Vs = reinterpret_cast<V*>(smem_raw);                        // [NT][m][N]: the row table (and its negated copy)
    int* rowidx_all = reinterpret_cast<int*>(Vs + (size_t)NT * a.m * N);  // [GPB][N]
    const int neg_off = a.m * N;
    __shared__ int incol[N];
    __shared__ double fin_s;

    const int tid = threadIdx.x;
    const int m = a.m;
    a.U += (size_t)blockIdx.z * a.u_stride;
    a.in_states += (size_t)blockIdx.z * a.in_stride;
    a.out += (size_t)blockIdx.z * a.out_stride;
    const uint8_t* ins = a.in_states + (size_t)blockIdx.y * m;
    if (tid == 0) {
        int c = 0;
        double f = 1.0;
        for (int i = 0; i < m; ++i)
            for (int k = 0; k < ins[i]; ++k) {
                if (c < N) incol[c] = i;
                ++c;
                f *= (double)(k + 1);
            }
        fin_s = f;
    }
    __syncthreads();
    for (int e = tid; e < m * N; e += K1_THREADS) {
        int r = e / N, c = e - r * N;
        double2 u = a.U[(size_t)r * m + incol[c]];
        V v;
        v.x = (T)u.x;
        v.y = (T)u.y;
        Vs[e] = v;
        if constexpr (k1_neg(N)) {
            v.x = -v.x;
            v.y = -v.y;
            Vs[neg_off + e] = v;
        }
    }

    const int g_in_blk = tid >> LOG_L;
    const int lane_g = tid & (S::L - 1);
    int* rowidx = rowidx_all + g_in_blk * N;
    const uint64_t total_groups = (uint64_t)gridDim.x * S::GPB;
    const uint64_t gid = (uint64_t)blockIdx.x * S::GPB + g_in_blk;
    const uint64_t it_begin = (a.n_out * gid) / total_groups;
    const uint64_t it_end = (a.n_out * (gid + 1)) / total_groups;
    const uint64_t max_len = (a.n_out + total_groups - 1) / total_groups;
    const uint64_t my_len = it_end - it_begin;

    double fout = 1.0;
    if (lane_g == 0) {
        int x[N];
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = 0;
        if (my_len > 0) {
            if (a.out_states) {
                const uint8_t* os = a.out_states + it_begin * m;
                int c = 0;
                for (int i = 0; i < m; ++i)
                    for (int k = 0; k < os[i]; ++k) {
                        if (c < N) x[c] = i;
                        ++c;
                        fout *= (double)(k + 1);
                    }
            } else {
                colex_unrank(a.bin, a.r0 + it_begin, N, m, x);
            }
        }
#pragma unroll
        for (int i = 0; i < N; ++i) rowidx[i] = x[i] * N;
    }
    __syncthreads();

    const double fin = fin_s;
    double* outp = a.out + (size_t)blockIdx.y * a.n_out * (a.out_probs ? 1 : 2);

    for (uint64_t it = 0; it < max_len; ++it) {
        const bool active = it < my_len;
        double dr = 0.0, di = 0.0;
        glynn_chunk_acc<N, S::C, S::U, T>(Vs, rowidx, (uint32_t)lane_g, dr, di, neg_off);
        dr = group_sum<double>(dr, LOG_L);
        di = group_sum<double>(di, LOG_L);
        __syncwarp();
        if (lane_g == 0 && active) {
            if (!a.out_states) {  // factorials of the current occupation (runs of equal modes)
                fout = 1.0;
                int run = 1;
#pragma unroll
                for (int i = 1; i < N; ++i) {
                    run = (rowidx[i] == rowidx[i - 1]) ? run + 1 : 1;
                    fout *= (double)run;
                }
            }
            const double scale = sqrt(fin * fout);
            const double re = (2.0 * dr) / scale, im = (2.0 * di) / scale;
            const uint64_t o = it_begin + it;
            LWB_BOUNDS(o < a.n_out);
            if (a.out_probs) {
                const double pr_out = re * re + im * im;
                outp[o] = pr_out;
                for (int g = 0; g < a.mir.n; ++g) a.mir.p[g][(size_t)blockIdx.y * a.n_out + o] = pr_out;
            } else { outp[2 * o] = re; outp[2 * o + 1] = im; }
            // advance to the next output state
            if (it + 1 < my_len) {
                if (a.out_states) {
                    const uint8_t* os = a.out_states + (o + 1) * m;
                    int c = 0;
                    fout = 1.0;
                    for (int i = 0; i < m; ++i)
                        for (int k = 0; k < os[i]; ++k) {
                            if (c < N) rowidx[c] = i * N;
                            ++c;
                            fout *= (double)(k + 1);
                        }
                } else {
                    // colex successor on the mode list (units of N)
                    for (int p = 0; p < N; ++p) {
                        const int lim = (p == N - 1) ? (m - 1) * N : rowidx[p + 1];
                        if (rowidx[p] < lim) {
                            rowidx[p] += N;
                            for (int q = 0; q < p; ++q) rowidx[q] = 0;
                            break;
                        }
                    }
                }
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
struct PermMatricesArgs {
    const double2* A;  // [B][n][n] complex128 row-major, device
    uint64_t B;
    double* out;       // [B][2]
};

template <int N, int LOG_L, typename T>
__global__ void __launch_bounds__(K1_THREADS, LWB_MINB) perm_matrices_kernel(PermMatricesArgs a) {
    using S = K1Shape<N, LOG_L>;
    using V = typename vec2<T>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    V* Ms_all = reinterpret_cast<V*>(smem_raw);  // [GPB][2][N*N]: every group's matrix and its negated copy
    __shared__ int ident[N];
    const int tid = threadIdx.x;
    if (tid < N) ident[tid] = tid * N;
    const int g_in_blk = tid >> LOG_L;
    const int lane_g = tid & (S::L - 1);
    constexpr int NT = k1_neg(N) ? 2 : 1;
    V* Ms = Ms_all + (size_t)g_in_blk * NT * N * N;
    constexpr int neg_off = N * N;
    const uint64_t total_groups = (uint64_t)gridDim.x * S::GPB;
    const uint64_t gid = (uint64_t)blockIdx.x * S::GPB + g_in_blk;
    const uint64_t iters = (a.B + total_groups - 1) / total_groups;
    __syncthreads();
    for (uint64_t it = 0; it < iters; ++it) {
        const uint64_t b = gid + it * total_groups;
        const bool active = b < a.B;
        __syncwarp();
        if (active) {
            const double2* src = a.A + b * (uint64_t)(N * N);
            for (int e = lane_g; e < N * N; e += S::L) {
                double2 u = src[e];
                V v;
                v.x = (T)u.x;
                v.y = (T)u.y;
                Ms[e] = v;
                if constexpr (k1_neg(N)) {
                    v.x = -v.x;
                    v.y = -v.y;
                    Ms[neg_off + e] = v;
                }
            }
        } else if (it == 0) {
            for (int e = lane_g; e < NT * N * N; e += S::L) { V v; v.x = T(0); v.y = T(0); Ms[e] = v; }
        }
        __syncwarp();
        double dr = 0.0, di = 0.0;
        glynn_chunk_acc<N, S::C, S::U, T>(Ms, ident, (uint32_t)lane_g, dr, di, neg_off);
        dr = group_sum<double>(dr, LOG_L);
        di = group_sum<double>(di, LOG_L);
        if (lane_g == 0 && active) {
            a.out[2 * b] = 2.0 * dr;
            a.out[2 * b + 1] = 2.0 * di;
        }
    }
}

// ---------------------------------------------------------------------------------------------
struct PermSingleArgs {
    const double2* A;   // [n][n]
    uint64_t q_begin, q_end;  // chunk range handled by this launch (this rank)
    double* partials;   // [gridDim.x][2]
    unsigned int* done; // blocks finished (zero before the launch; the last block resets it)
    double* out;        // [2]: 2 * sum of the partials, written by the last block to finish
};

template <int N>
struct K2Shape {
    // chunk bits: every lane gets a contiguous range of ~28+ chunks (glynn_range), so C only has to keep the
    // lane-specific boundary flips rare; small n cannot fill the GPU anyway (use the batched kernel there)
    static constexpr int C0 = N - LWB_K2_COFF;
    static constexpr int CC = C0 < 3 ? 3 : (C0 > 7 ? 7 : C0);
    static constexpr int C = CC > N - 1 ? (N - 1 < 1 ? 1 : N - 1) : CC;
    static constexpr int U = C < LWB_U2(N) ? C : LWB_U2(N);
};

template <int N, typename T>
__global__ void __launch_bounds__(K1_THREADS, LWB_MINB) perm_single_kernel(PermSingleArgs a) {
    using S = K2Shape<N>;
    using V = typename vec2<T>::type;
    // the matrix and, when both fit the 48 KB of static shared memory (n <= 37 in double), its negated copy.
    // Rows are RS = N | 1 elements apart: at a chunk boundary every lane flips its own row, and with an even N
    // (n = 24: 96 words, a multiple of the 32 banks) all those rows start in the same bank - an 8-way conflict on each of
    // the N loads, every 2^C steps (measured: n = 24 ran 38 % slower per step than n = 26).  An odd stride spreads
    // consecutive rows over all banks.
    constexpr int RS = N | 1;
    constexpr bool NEG = 2 * N * RS * sizeof(V) + 1024 <= 48 * 1024;
    constexpr int neg_off = N * RS;
    __shared__ __align__(16) V Ms[(NEG ? 2 : 1) * N * RS];
    __shared__ int ident[N];
    __shared__ double red[2][K1_THREADS / 32];
    const int tid = threadIdx.x;
    if (tid < N) ident[tid] = tid * RS;
    for (int e = tid; e < N * N; e += K1_THREADS) {
        double2 u = a.A[e];
        V v;
        v.x = (T)u.x;
        v.y = (T)u.y;
        const int at = (e / N) * RS + e % N;
        Ms[at] = v;
        if constexpr (NEG) {
            v.x = -v.x;
            v.y = -v.y;
            Ms[neg_off + at] = v;
        }
    }
    __syncthreads();
    double dr = 0.0, di = 0.0;
    if constexpr (N == 1) {
        if (blockIdx.x == 0 && tid == 0 && a.q_begin < a.q_end) { dr = 0.5 * (double)Ms[0].x; di = 0.5 * (double)Ms[0].y; }
    } else {
        // equal contiguous shares of the chunk range for every lane of the grid
        const uint64_t lanes = (uint64_t)gridDim.x * K1_THREADS;
        const uint64_t lane = (uint64_t)blockIdx.x * K1_THREADS + tid;
        const uint64_t total = a.q_end - a.q_begin;
        const uint64_t b = a.q_begin + (total / lanes) * lane + (lane < total % lanes ? lane : total % lanes);
        const uint64_t cnt = total / lanes + (lane < total % lanes ? 1 : 0);
        // complex64: restart from a fresh start vector every 2^LWB_F32_RESTART steps (drift control, see glynn_chunk_acc)
        constexpr uint64_t PIECE = sizeof(T) == 4 ? (1ull << (LWB_F32_RESTART > S::C ? LWB_F32_RESTART - S::C : 0)) : ~0ull;
#pragma unroll 1
        for (uint64_t done = 0; done < cnt; done += PIECE) {
            T sr = T(0), si = T(0);
            glynn_range<N, S::C, S::U, T, NEG>(Ms, ident, b + done, (cnt - done < PIECE) ? cnt - done : PIECE, sr, si, neg_off);
            dr += (double)sr;
            di += (double)si;
        }
    }
    __syncwarp();
    dr = group_sum<double>(dr, 5);
    di = group_sum<double>(di, 5);
    if ((tid & 31) == 0) { red[0][tid >> 5] = dr; red[1][tid >> 5] = di; }
    __syncthreads();
    __shared__ bool is_last;
    if (tid == 0) {
        double r = 0.0, i = 0.0;
        for (int w = 0; w < K1_THREADS / 32; ++w) { r += red[0][w]; i += red[1][w]; }
        a.partials[2 * blockIdx.x] = r;
        a.partials[2 * blockIdx.x + 1] = i;
        __threadfence();  // the partial is visible before the counter moves
        is_last = atomicAdd(a.done, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    // Last block to finish: fixed-order sum of every block's partial (deterministic whatever the finishing order),
    // out = 2 * sum  (perm = 2 * sum_g sign_g prod_j w_j).  No second launch: at n = 24 the separate reduction
    // kernel and its launch gap were ~8 % of the 0.11 ms.
    __threadfence();
    double r = 0.0, i = 0.0;
    for (int k = tid; k < (int)gridDim.x; k += K1_THREADS) {
        r += __ldcg(a.partials + 2 * k);
        i += __ldcg(a.partials + 2 * k + 1);
    }
    r = group_sum<double>(r, 5);
    i = group_sum<double>(i, 5);
    __syncthreads();
    if ((tid & 31) == 0) { red[0][tid >> 5] = r; red[1][tid >> 5] = i; }
    __syncthreads();
    if (tid == 0) {
        double sr = 0.0, si = 0.0;
        for (int w = 0; w < K1_THREADS / 32; ++w) { sr += red[0][w]; si += red[1][w]; }
        a.out[0] = 2.0 * sr;
        a.out[1] = 2.0 * si;
        *a.done = 0u;  // ready for the next launch on this stream
    }
}

// fixed-order final reduction: out = 2 * sum(partials)   (perm = 2 * sum_g sign_g prod_j w_j)
static __global__ void reduce_partials_kernel(const double* partials, int n, double* out) {
    __shared__ double sr[256], si[256];
    double r = 0.0, i = 0.0;
    for (int k = threadIdx.x; k < n; k += 256) { r += partials[2 * k]; i += partials[2 * k + 1]; }
    sr[threadIdx.x] = r;
    si[threadIdx.x] = i;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) { sr[threadIdx.x] += sr[threadIdx.x + s]; si[threadIdx.x] += si[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[0] = 2.0 * sr[0]; out[1] = 2.0 * si[0]; }
}

}  // namespace lwb
