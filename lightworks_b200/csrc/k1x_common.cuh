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
#define K1X_BLOCK(N) ((N) <= 10 ? 256 : 128)
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

template <int TB>
struct K1xLane {   // what the prologue hands to the walk of one permanent
    uint32_t vs_base;  // shared-memory address of the U[:, input columns] table, row stride NP * 16 bytes
    const int* srow;   // [N][TB]: row offset (in double2) of slot sl for lane tid at srow[sl * TB + tid]
    int tid;
    __device__ __forceinline__ uint32_t slot_addr(int sl) const { return vs_base + 16u * (uint32_t)srow[sl * TB + tid]; }
};

__device__ __forceinline__ double2 k1x_ld(uint32_t addr, int j) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr + 16u * j));
    return v;
}

// start tuple v = 0 (every copy of every row '+'):  w_j = 1/2 sum_i s_i a_{r_i j}
template <int N>
__device__ __forceinline__ void k1x_start(const K1xLane<k1x_threads(N)>& ln, const TypeMeta& tm, double (&wr)[N], double (&wi)[N]) {
#pragma unroll
    for (int j = 0; j < N; ++j) { wr[j] = 0.0; wi[j] = 0.0; }
    for (int sl = 0; sl < tm.d; ++sl) {
        const double c = 0.5 * (double)tm.s[sl];
        const uint32_t ra = ln.slot_addr(sl);
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const double2 v = k1x_ld(ra, j);
            wr[j] = fma(c, v.x, wr[j]);
            wi[j] = fma(c, v.y, wi[j]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Generic table-driven walk: every tuple of every row is one Step of the program.  The step descriptor and this
// lane's row address are fetched one term ahead, and the row of a flip is loaded BEFORE the product chain of the
// same term, so the shared-memory latency hides behind ~40 FP64 instructions instead of stalling the adds.
template <int N>
__device__ __forceinline__ void k1x_generic_walk(const K1xLane<k1x_threads(N)>& ln, const TypeMeta& tm, const Step* __restrict__ prog,
                                                 double& sr, double& si) {
    double wr[N], wi[N];
    k1x_start<N>(ln, tm, wr, wi);
    const int T = tm.T;
    sr = 0.0;
    si = 0.0;
    Step st = prog[0];  // same address for every lane of the block: one broadcast load
    uint32_t rowaddr = ln.slot_addr(st.slot);
#pragma unroll k1x_unroll
    for (int t = 0; t < T; ++t) {
        const Step nx = prog[t + 1 < T ? t + 1 : t];
        const uint32_t nrowaddr = ln.slot_addr(nx.slot);
        double2 rv[N];
#pragma unroll
        for (int j = 0; j < N; ++j) rv[j] = k1x_ld(rowaddr, j);
        double pr, pi;
        cprod<N, double>(wr, wi, pr, pi);
        sr = fma(st.coef, pr, sr);
        si = fma(st.coef, pi, si);
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
template <int N, int C>
__device__ __forceinline__ void k1x_static_walk(const K1xLane<k1x_threads(N)>& ln, const TypeMeta& tm, const Step* __restrict__ prog,
                                                double& sr, double& si) {
    static_assert(C >= 2 && C <= N - 1, "static walk needs at least 3 single rows");
    constexpr int U = C < K1XS_U(N) ? C : K1XS_U(N);
    const int d = tm.d;  // distinct rows; slots [0, d-C-1) bunched, then C+1 singles in ascending mode order
    double wr[N], wi[N];
    k1x_start<N>(ln, tm, wr, wi);
    // Gray bit k <-> slot d-1-k: the highest modes flip fastest (the ones neighbouring lanes share, so their
    // shared-memory reads mostly broadcast); slot d-1-C is the fixed '+' single.  (Keeping the two fastest rows in
    // registers instead was measured: no gain, the kernel is not bound by shared-memory bandwidth.)
    uint32_t lowaddr[U];
#pragma unroll
    for (int k = 0; k < U; ++k) lowaddr[k] = ln.slot_addr(d - 1 - k);
    const int T = tm.T;
    sr = 0.0;
    si = 0.0;
    constexpr uint32_t NBLK = 1u << (C - U);
    constexpr int BLK = 1 << U;
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
        const Step st = prog[t];
        const uint32_t par = (uint32_t)t & 1u;
        double cr = 0.0, ci = 0.0;
#pragma unroll 1
        for (uint32_t ob = 0; ob < NBLK; ++ob) {
#pragma unroll
            for (int tt = 0; tt < BLK; ++tt) {
                double pr, pi;
                cprod<N, double>(wr, wi, pr, pi);
                if (tt & 1) { cr -= pr; ci -= pi; }
                else { cr += pr; ci += pi; }
                if (tt + 1 < BLK) {
                    const int t1 = tt + 1;
                    const int k = cx_ctz(t1);
                    if (k <= U - 2) {
                        const bool newbit = ((t1 >> k) ^ (t1 >> (k + 1))) & 1;
#pragma unroll
                        for (int j = 0; j < N; ++j) {
                            const double2 v = k1x_ld(lowaddr[k], j);
                            if (newbit) { wr[j] -= v.x; wi[j] -= v.y; }
                            else { wr[j] += v.x; wi[j] += v.y; }
                        }
                    } else {  // k == U-1: bit U of the step index decides
                        const uint32_t hi = (U == C) ? par : (ob & 1u);
                        const double sg = hi ? 1.0 : -1.0;
#pragma unroll
                        for (int j = 0; j < N; ++j) {
                            const double2 v = k1x_ld(lowaddr[k], j);
                            wr[j] = fma(sg, v.x, wr[j]);
                            wi[j] = fma(sg, v.y, wi[j]);
                        }
                    }
                }
            }
            if (ob + 1 < NBLK) {
                const uint32_t o1 = ob + 1;
                const int k = U + (__ffs(o1) - 1);
                const uint32_t newbit = (k == C - 1) ? (1u ^ par) : (((o1 >> (k - U)) ^ (o1 >> (k - U + 1))) & 1u);
                const double sg = newbit ? -1.0 : 1.0;
                const uint32_t ra = ln.slot_addr(d - 1 - k);
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    const double2 v = k1x_ld(ra, j);
                    wr[j] = fma(sg, v.x, wr[j]);
                    wi[j] = fma(sg, v.y, wi[j]);
                }
            }
        }
        sr = fma(st.coef, cr, sr);
        si = fma(st.coef, ci, si);
        if (t + 1 < T) {  // next tuple of the bunched rows: v_slot += dir  =>  w -= dir * row
            const double sg = st.dir > 0 ? -1.0 : 1.0;
            const uint32_t ra = ln.slot_addr(st.slot);
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const double2 v = k1x_ld(ra, j);
                wr[j] = fma(sg, v.x, wr[j]);
                wi[j] = fma(sg, v.y, wi[j]);
            }
        }
    }
}

// klass C >= 2 -> k1x_static_walk<N, C>, anything else -> the generic walk (block-uniform branch)
template <int N, int C>
struct K1xDispatch {
    static __device__ __forceinline__ void run(int klass, const K1xLane<k1x_threads(N)>& ln, const TypeMeta& tm, const Step* prog,
                                               double& sr, double& si) {
        if (klass == C) k1x_static_walk<N, C>(ln, tm, prog, sr, si);
        else K1xDispatch<N, C - 1>::run(klass, ln, tm, prog, sr, si);
    }
};
template <int N>
struct K1xDispatch<N, 1> {
    static __device__ __forceinline__ void run(int, const K1xLane<k1x_threads(N)>& ln, const TypeMeta& tm, const Step* prog, double& sr,
                                               double& si) {
        k1x_generic_walk<N>(ln, tm, prog, sr, si);
    }
};

// One block = up to 256 permanents of ONE multiplicity pattern, one thread per permanent.  Block b finds its
// pattern by a binary search of the device-built block table, so the whole pipeline needs no host round trip and a
// single launch covers every pattern (grid = upper bound, surplus blocks leave at once).
template <int N, bool STATIC>
__device__ __forceinline__ void k1x_block(const K1xArgs& a) {
    constexpr int TB = k1x_threads(N);
    if (blockIdx.x >= a.bstart[a.n_types]) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NP = N | 1;  // odd row stride (in 16-byte units): lanes reading different rows spread over the banks
    double2* Vs = reinterpret_cast<double2*>(smem_raw);         // [m][NP]
    int* srow = reinterpret_cast<int*>(Vs + (size_t)a.m * NP);  // [N][TB] row offsets per slot
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
        Vs[r * NP + c] = a.U[(size_t)r * m + incol[c]];
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
    K1xLane<TB> ln;
    ln.vs_base = (uint32_t)__cvta_generic_to_shared(Vs);
    ln.srow = srow;
    ln.tid = tid;
    double sr, si;
    K1xDispatch<N, STATIC ? N - 1 : 1>::run(tm.klass, ln, tm, a.prog + tm.prog_off, sr, si);
    const double scale = sqrt(fin_s * fout);
    const double f = (double)tm.factor;
    const double re = (f * sr) / scale, im = (f * si) / scale;
    const double pr_out = re * re + im * im;
    a.probs[k_out] = pr_out;
    for (int g = 0; g < a.mir.n; ++g) a.mir.p[g][k_out] = pr_out;  // fire-and-forget stores into peer memory
}

// (launch bounds: an explicit minBlocks of 1 makes ptxas schedule the walk ~25 % slower than none on B200)
template <int N, bool STATIC>
__global__ void __launch_bounds__(k1x_threads(N)) k1x_fused_kernel(K1xArgs a) { k1x_block<N, STATIC>(a); }
template <int N, bool STATIC>
__global__ void __launch_bounds__(k1x_threads(N), 2 * (256 / k1x_threads(N))) k1x_fused_kernel_2(K1xArgs a) { k1x_block<N, STATIC>(a); }

typedef void (*K1xFn)(K1xArgs);
struct K1xStaticEntry {
    int n;
    K1xFn fn;
};
template <int N>
constexpr K1xFn k1x_static_fn() {
    if constexpr (K1XS_MINB(N) == 2) return k1x_fused_kernel_2<N, true>;
    else return k1x_fused_kernel<N, true>;
}
#define LWB_K1XS_ENTRY(N) {N, k1x_static_fn<N>()}

const K1xStaticEntry* k1xs_g0_table(int* count);
const K1xStaticEntry* k1xs_g1_table(int* count);
const K1xStaticEntry* k1xs_g2_table(int* count);
const K1xStaticEntry* k1xs_g3_table(int* count);
#define LWB_K1XS_GROUPS {k1xs_g0_table, k1xs_g1_table, k1xs_g2_table, k1xs_g3_table}

}  // namespace lwb
