// k1x_common.cuh -- data structures, the two walks (generic table-driven, static Gray chunk) and the fused kernel of
// the multiplicity-aware permanent (see k1x.h); shared by k1x.cu (host pipeline, bucketing kernels, n <= 6) and the
// k1xs_g*.cu instantiation units (n = 7..16).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fock.cuh"
#include "glynn.cuh"

namespace lwb {

#ifndef K1X_UNROLL
#define K1X_UNROLL 2
#endif
// permanents (= threads) per block of the walk kernel.  Measured on B200 (tools/k1xs_sweep.sh): n <= 10 fits two
// 256-thread blocks per SM in 128 registers; beyond that the walks want ~250 registers and two 128-thread blocks
// stagger their prologues and tails better than one 256-thread block (6 % at n = 12, 3 % at n = 14).
#ifndef K1X_BLOCK
#define K1X_BLOCK(N) 128  /* round 2: n <= 10 runs 128 threads x 3 blocks/SM (168 registers, no spills): -1.1 % at C3 */
#endif
constexpr int k1x_unroll = K1X_UNROLL;
__host__ __device__ constexpr int k1x_threads(int n) { return K1X_BLOCK(n); }
constexpr int K1X_MAX_THREADS = 256;
constexpr int K1X_MAX_N = 16;       // kernels instantiated for 2 <= n <= 16
constexpr int K1X_CL_THREADS = 1024;  // one state per thread in the bucketing passes
constexpr int K1X_MAX_TYPES = 256;    // partitions of n: p(16) = 231
constexpr int K1X_STATIC_MIN_N = 7;   // fused kernels with static walks exist for 7 <= n <= 16 (k1xs_g*.cu)

struct TypeMeta {
    int d;         // distinct rows
    int T;         // steps of the program: every term (generic walk) or every tuple of the bunched rows (static walk)
    int prog_off;  // first Step of this type in the program array
    int factor;    // 2 when the v -> s-v symmetry halves the sum, else 1
    int ns;        // singly occupied rows (the last ns slots)
    int klass;     // -1: generic table-driven walk; C >= 2: k1x_static_walk<N, C>, C = ns - 1
    int s[K1X_MAX_N];  // multiplicities, descending
};
struct Step {
    double coef;  // (-1)^{sum v} prod C(s_i, v_i) of the tuple BEFORE the transition
    int slot;     // row slot whose v changes after this term (unused for the last term)
    int dir;      // +1 / -1
};
struct K1xArgs {
    const uint64_t* bin;
    const double2* U;
    int m;
    const uint8_t* in_state;
    uint64_t r0;
    const TypeMeta* meta;
    const Step* prog;
    // bucket layout, all computed on the device (k1x_layout_kernel): position p of the layout holds type layout[p];
    // its states are bucket[offsets[type] .. + counts[type]) and its blocks are [bstart[p], bstart[p+1])
    const int* layout;
    const uint32_t* bstart;   // [n_types + 1]
    const uint32_t* offsets;  // [n_types], indexed by type
    const uint32_t* counts;   // [n_types], indexed by type
    int n_types;
    const uint32_t* bucket;
    uint32_t cnt;  // states of the rank range (debug bounds checks)
    double* probs;
    MirrorSet mir;  // peers that receive every probability as it is produced (multi-GPU: no gather step afterwards)
};

// per-N tuning of the fused kernel: low Gray bits unrolled in the static walks, blocks per SM asked of ptxas
// (measured on B200, tools/k1xs_sweep.sh: n = 10 gains 9 % from 2 blocks/SM with U = 3; n >= 12 loses 25 % to spills)
#ifndef K1XS_U
#define K1XS_U(N) ((N) <= 10 ? 3 : (N) <= 12 ? 4 : 3)
#endif
#ifndef K1XS_MINB
#define K1XS_MINB(N) ((N) <= 10 ? 2 : 0)
#endif

template <int TB, typename T>
struct K1xLane {   // what the prologue hands to the walk of one permanent
    uint32_t vs_base;  // shared-memory address of the U[:, input columns] table, row stride NP complex entries
    uint32_t neg_bytes;  // byte distance to the NEGATED copy of the table (run-time signs select the copy: see row_sgn)
    const int* srow;   // [N][TB]: row offset (in complex entries) of slot sl for lane tid at srow[sl * TB + tid]
    int tid;
    __device__ __forceinline__ uint32_t slot_addr(int sl) const {
        return vs_base + (uint32_t)(2 * sizeof(T)) * (uint32_t)srow[sl * TB + tid];
    }
};

__device__ __forceinline__ double2 k1x_ld(uint32_t addr, int j, double) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr + 16u * j));
    return v;
}
__device__ __forceinline__ float2 k1x_ld(uint32_t addr, int j, float) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr + 8u * j));
    return v;
}

// start tuple v = 0 (every copy of every row '+'):  w_j = 1/2 sum_i s_i a_{r_i j}
template <int N, typename T>
__device__ __forceinline__ void k1x_start(const K1xLane<k1x_threads(N), T>& ln, const TypeMeta& tm, T (&wr)[N], T (&wi)[N]) {
#pragma unroll
    for (int j = 0; j < N; ++j) { wr[j] = T(0); wi[j] = T(0); }
    for (int sl = 0; sl < tm.d; ++sl) {
        const T c = T(0.5) * (T)tm.s[sl];
        const uint32_t ra = ln.slot_addr(sl);
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const auto v = k1x_ld(ra, j, T());
            wr[j] = fma(c, v.x, wr[j]);
            wi[j] = fma(c, v.y, wi[j]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Generic table-driven walk: every tuple of every row is one Step of the program.  The step descriptor and this
// lane's row address are fetched one term ahead, and the row of a flip is loaded BEFORE the product chain of the
// same term, so the shared-memory latency hides behind ~40 FP64 instructions instead of stalling the adds.
template <int N, typename T>
__device__ __forceinline__ void k1x_generic_walk(const K1xLane<k1x_threads(N), T>& ln, const TypeMeta& tm, const Step* __restrict__ prog,
                                                 double& sr, double& si) {
    using V = typename vec2<T>::type;
    T wr[N], wi[N];
    k1x_start<N, T>(ln, tm, wr, wi);
    const int n_steps = tm.T;
    sr = 0.0;
    si = 0.0;
    Step st = prog[0];  // same address for every lane of the block: one broadcast load
    uint32_t rowaddr = ln.slot_addr(st.slot);
#pragma unroll k1x_unroll
    for (int t = 0; t < n_steps; ++t) {
        const Step nx = prog[t + 1 < n_steps ? t + 1 : t];
        const uint32_t nrowaddr = ln.slot_addr(nx.slot);
        V rv[N];
#pragma unroll
        for (int j = 0; j < N; ++j) rv[j] = k1x_ld(rowaddr, j, T());
        T pr, pi;
        cprod<N, T>(wr, wi, pr, pi);
        sr = fma(st.coef, (double)pr, sr);  // the weighted sum is kept in double for both precisions
        si = fma(st.coef, (double)pi, si);
        // v_slot += dir  =>  (s - 2v)/2 decreases by dir   (the update after the last term is harmless)
#pragma unroll
        for (int j = 0; j < N; ++j) {
            if (st.dir > 0) { wr[j] -= rv[j].x; wi[j] -= rv[j].y; }
            else { wr[j] += rv[j].x; wi[j] += rv[j].y; }
        }
        st = nx;
        rowaddr = nrowaddr;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Static walk for patterns with >= 3 singly occupied rows.  The walk factorises: the tuples of the BUNCHED rows
// (few: prod (s_i + 1)) are walked by the pre-computed program, and for each of them the 2^C sign patterns of C of
// the singles (the remaining single is the fixed '+' row that halves the sum) are walked by a binary Gray chunk with
// its low U bits statically unrolled (flip signs known at compile time -> 2-operand DADDs): chunk number t starts
// where chunk t-1 ended, so it is exactly Gray chunk q = t of glynn_chunk (only the parity of t matters) and needs
// no new start vector.  The program's coefficient carries the binomials and signs of the bunched tuple and the
// parity sign of the chunk.
template <int N, int C, typename T>
__device__ __forceinline__ void k1x_static_walk(const K1xLane<k1x_threads(N), T>& ln, const TypeMeta& tm, const Step* __restrict__ prog,
                                                double& sr, double& si) {
    static_assert(C >= 2 && C <= N - 1, "static walk needs at least 3 single rows");
    constexpr int U = C < K1XS_U(N) ? C : K1XS_U(N);
    const int d = tm.d;  // distinct rows; slots [0, d-C-1) bunched, then C+1 singles in ascending mode order
    T wr[N], wi[N];
    k1x_start<N, T>(ln, tm, wr, wi);
    // Gray bit k <-> slot d-1-k: the highest modes flip fastest (the ones neighbouring lanes share, so their
    // shared-memory reads mostly broadcast); slot d-1-C is the fixed '+' single.  (Keeping the two fastest rows in
    // registers instead was measured: no gain, the kernel is not bound by shared-memory bandwidth.)
    uint32_t lowaddr[U];
#pragma unroll
    for (int k = 0; k < U; ++k) lowaddr[k] = ln.slot_addr(d - 1 - k);
    const int n_steps = tm.T;
    sr = 0.0;
    si = 0.0;
    constexpr uint32_t NBLK = 1u << (C - U);
    constexpr int BLK = 1 << U;
#pragma unroll 1
    for (int t = 0; t < n_steps; ++t) {
        const Step st = prog[t];
        const uint32_t par = (uint32_t)t & 1u;
        T cr = T(0), ci = T(0);
#pragma unroll 1
        for (uint32_t ob = 0; ob < NBLK; ++ob) {
#pragma unroll
            for (int tt = 0; tt < BLK; ++tt) {
                T pr, pi;
                cprod<N, T>(wr, wi, pr, pi);
                if (tt & 1) { cr -= pr; ci -= pi; }
                else { cr += pr; ci += pi; }
                if (tt + 1 < BLK) {
                    const int t1 = tt + 1;
                    const int k = cx_ctz(t1);
                    if (k <= U - 2) {
                        const bool newbit = ((t1 >> k) ^ (t1 >> (k + 1))) & 1;
#pragma unroll
                        for (int j = 0; j < N; ++j) {
                            const auto v = k1x_ld(lowaddr[k], j, T());
                            if (newbit) { wr[j] -= v.x; wi[j] -= v.y; }
                            else { wr[j] += v.x; wi[j] += v.y; }
                        }
                    } else {  // k == U-1: bit U of the step index decides
                        const uint32_t hi = (U == C) ? par : (ob & 1u);
                        const uint32_t ra = lowaddr[k] + (hi ? 0u : ln.neg_bytes);  // sign -1: add the negated row
#pragma unroll
                        for (int j = 0; j < N; ++j) {
                            const auto v = k1x_ld(ra, j, T());
                            wr[j] += v.x;
                            wi[j] += v.y;
                        }
                    }
                }
            }
            if (ob + 1 < NBLK) {
                const uint32_t o1 = ob + 1;
                const int k = U + (__ffs(o1) - 1);
                const uint32_t newbit = (k == C - 1) ? (1u ^ par) : (((o1 >> (k - U)) ^ (o1 >> (k - U + 1))) & 1u);
                const uint32_t ra = ln.slot_addr(d - 1 - k) + (newbit ? ln.neg_bytes : 0u);
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    const auto v = k1x_ld(ra, j, T());
                    wr[j] += v.x;
                    wi[j] += v.y;
                }
            }
        }
        sr = fma(st.coef, (double)cr, sr);  // chunk sums join the double accumulator
        si = fma(st.coef, (double)ci, si);
        if (t + 1 < n_steps) {  // next tuple of the bunched rows: v_slot += dir  =>  w -= dir * row
            const uint32_t ra = ln.slot_addr(st.slot) + (st.dir > 0 ? ln.neg_bytes : 0u);
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const auto v = k1x_ld(ra, j, T());
                wr[j] += v.x;
                wi[j] += v.y;
            }
        }
    }
}

// klass C >= 2 -> k1x_static_walk<N, C>, anything else -> the generic walk (block-uniform branch)
template <int N, int C, typename T>
struct K1xDispatch {
    static __device__ __forceinline__ void run(int klass, const K1xLane<k1x_threads(N), T>& ln, const TypeMeta& tm, const Step* prog,
                                               double& sr, double& si) {
        if (klass == C) k1x_static_walk<N, C, T>(ln, tm, prog, sr, si);
        else K1xDispatch<N, C - 1, T>::run(klass, ln, tm, prog, sr, si);
    }
};
template <int N, typename T>
struct K1xDispatch<N, 1, T> {
    static __device__ __forceinline__ void run(int, const K1xLane<k1x_threads(N), T>& ln, const TypeMeta& tm, const Step* prog, double& sr,
                                               double& si) {
        k1x_generic_walk<N, T>(ln, tm, prog, sr, si);
    }
};

// One block = up to 256 permanents of ONE multiplicity pattern, one thread per permanent.  Block b finds its
// pattern by a binary search of the device-built block table, so the whole pipeline needs no host round trip and a
// single launch covers every pattern (grid = upper bound, surplus blocks leave at once).
template <int N, bool STATIC, typename T>
__device__ __forceinline__ void k1x_block(const K1xArgs& a) {
    using V = typename vec2<T>::type;
    constexpr int TB = k1x_threads(N);
    if (blockIdx.x >= a.bstart[a.n_types]) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NP = N | 1;  // odd row stride (in complex entries): lanes reading different rows spread over the banks
    V* Vs = reinterpret_cast<V*>(smem_raw);                         // [2][m][NP]: the table and its negated copy
    int* srow = reinterpret_cast<int*>(Vs + (size_t)2 * a.m * NP);  // [N][TB] row offsets per slot
    __shared__ int incol[N];
    __shared__ double fin_s;
    __shared__ TypeMeta tm;
    __shared__ uint32_t range[2];
    __shared__ int type_s;
    __shared__ unsigned char sin[LWB_MAX_MODES];
    const int tid = threadIdx.x;
    const int m = a.m;
    // the prologue's global reads are issued in parallel (one round trip each), not as a serial chain on thread 0:
    // a block lives ~50 us and a chain of ~30 dependent L2 reads would idle it for 10 % of that
    for (int i = tid; i < m; i += TB) sin[i] = a.in_state[i];
    for (int p = tid; p < a.n_types; p += TB) {  // the layout position whose block range holds blockIdx.x
        const uint32_t b0 = a.bstart[p], b1 = a.bstart[p + 1];
        if (b0 <= blockIdx.x && blockIdx.x < b1) {
            const int type = a.layout[p];
            const uint32_t first = a.offsets[type] + (blockIdx.x - b0) * TB;
            const uint32_t last = a.offsets[type] + a.counts[type];
            range[0] = first;
            range[1] = first + TB < last ? first + TB : last;
            type_s = type;
        }
    }
    __syncthreads();
    if (tid == 0) {
        int c = 0;
        double f = 1.0;
        for (int i = 0; i < m; ++i)
            for (int k = 0; k < sin[i]; ++k) {
                if (c < N) incol[c] = i;
                ++c;
                f *= (double)(k + 1);
            }
        fin_s = f;
    }
    if (tid >= 32 && tid < 32 + (int)(sizeof(TypeMeta) / 4))
        reinterpret_cast<int*>(&tm)[tid - 32] = reinterpret_cast<const int*>(a.meta + type_s)[tid - 32];
    __syncthreads();
    for (int e = tid; e < m * N; e += TB) {
        const int r = e / N, c = e - r * N;
        const double2 u = a.U[(size_t)r * m + incol[c]];
        V v;
        v.x = (T)u.x;
        v.y = (T)u.y;
        Vs[r * NP + c] = v;
        v.x = -v.x;
        v.y = -v.y;
        Vs[(m + r) * NP + c] = v;
    }
    __syncthreads();
    const uint32_t item = range[0] + tid;
    if (item >= range[1]) return;
    const uint32_t k_out = a.bucket[item];

    // this thread's output state -> distinct modes ordered by decreasing multiplicity (= the pattern's slot order)
    double fout = 1.0;
    {
        int x[N];
        colex_unrank(a.bin, a.r0 + k_out, N, m, x);
        int mode[N], mult[N];
        int nd = 0, run = 1;
        for (int i = 1; i <= N; ++i) {
            if (i < N && x[i] == x[i - 1]) { ++run; fout *= (double)run; continue; }
            int p = nd++;
            while (p > 0 && mult[p - 1] < run) { mult[p] = mult[p - 1]; mode[p] = mode[p - 1]; --p; }
            mult[p] = run;
            mode[p] = x[i - 1];
            run = 1;
        }
        for (int sl = 0; sl < nd; ++sl) srow[sl * TB + tid] = mode[sl] * NP;
    }
    K1xLane<TB, T> ln;
    ln.vs_base = (uint32_t)__cvta_generic_to_shared(Vs);
    ln.neg_bytes = (uint32_t)(m * NP * sizeof(V));
    ln.srow = srow;
    ln.tid = tid;
    double sr, si;
    K1xDispatch<N, STATIC ? N - 1 : 1, T>::run(tm.klass, ln, tm, a.prog + tm.prog_off, sr, si);
    const double scale = sqrt(fin_s * fout);
    const double f = (double)tm.factor;
    const double re = (f * sr) / scale, im = (f * si) / scale;
    const double pr_out = re * re + im * im;
    LWB_BOUNDS(k_out < a.cnt);
    a.probs[k_out] = pr_out;
    for (int g = 0; g < a.mir.n; ++g) a.mir.p[g][k_out] = pr_out;  // fire-and-forget stores into peer memory
}

// (launch bounds: an explicit minBlocks of 1 makes ptxas schedule the walk ~25 % slower than none on B200)
template <int N, bool STATIC, typename T>
__global__ void __launch_bounds__(k1x_threads(N)) k1x_fused_kernel(K1xArgs a) { k1x_block<N, STATIC, T>(a); }
#ifndef K1XS_BLOCKS_PER_SM
#define K1XS_BLOCKS_PER_SM(N) 3
#endif
template <int N, bool STATIC, typename T>
__global__ void __launch_bounds__(k1x_threads(N), K1XS_BLOCKS_PER_SM(N)) k1x_fused_kernel_2(K1xArgs a) { k1x_block<N, STATIC, T>(a); }

// complex64 walks keep the 1e-4 promise only while one Gray chunk stays short (drift of the float column sums,
// see glynn_chunk_acc): the float kernels exist for n <= K1X_C64_MAX_N, beyond it complex64 uses the expanded-row kernel
constexpr int K1X_C64_MAX_N = 12;

typedef void (*K1xFn)(K1xArgs);
struct K1xStaticEntry {
    int n;
    K1xFn fn[2];  // [dtype]: complex128, complex64 (nullptr above K1X_C64_MAX_N)
};
template <int N, typename T>
constexpr K1xFn k1x_static_fn() {
    if constexpr (sizeof(T) == 4 && N > K1X_C64_MAX_N) return nullptr;
    else if constexpr (K1XS_MINB(N) == 2) return k1x_fused_kernel_2<N, true, T>;
    else return k1x_fused_kernel<N, true, T>;
}
#define LWB_K1XS_ENTRY(N) {N, {k1x_static_fn<N, double>(), k1x_static_fn<N, float>()}}

const K1xStaticEntry* k1xs_g0_table(int* count);
const K1xStaticEntry* k1xs_g1_table(int* count);
const K1xStaticEntry* k1xs_g2_table(int* count);
const K1xStaticEntry* k1xs_g3_table(int* count);
#define LWB_K1XS_GROUPS {k1xs_g0_table, k1xs_g1_table, k1xs_g2_table, k1xs_g3_table}

}  // namespace lwb
