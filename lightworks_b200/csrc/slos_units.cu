// slos_units.cu -- K3u / K3r: SLOS layer kernels over "prefix units" (round 2).  K3u (slos_units_kernel) derives the
// trailing indices per child; K3r (slos_units4_kernel, the default) reads them from host-built records - see the
// comment above that kernel.  Both share the units, the streams, the staged parent block and the host plan below.
//
// Same recurrence as slos_layer_kernel (slos_kernels.cuh; SLOSBackend.calculate, lightworks/emulator/backends/
// slos.py:96-103): for the k-th input photon in mode i
//     out[t] = sum_j sqrt(t_j) * in[t - e_j] * U[j, i]        (j ascending over the occupied modes of t)
// with every layer stored in the reference's dict order = lexicographic order of the ascending mode list
// x_0 <= ... <= x_{k-1}.  The gather-form kernel spends 612 warp-instructions per 32 children re-deriving positions
// and reads every parent through scattered 16-byte gathers (14.8 GB of L1/L2 traffic at 24 modes / 10 photons).
//
// Structure used here.  Cut the lex-ordered tree of mode lists at a frontier of nodes ("units"): a unit is a prefix
// P = (x_0..x_{k-d-1}) whose subtree - the d trailing photons in modes >= l = last(P), a "d-simplex" of
// T = C(m-l+d-1, d) consecutive children - is at most SLOS_TMAX states.  Inside one unit
//   * removing a PREFIX photon (mode v) from every child gives parents that are CONSECUTIVE in the parent layer and
//     end-aligned with the unit: parent(c) = in[pend_v - T + c]  -> one coalesced stream per distinct prefix mode;
//   * removing a TRAILING photon gives a parent inside the block B(P) = (P, d-1 trailing photons >= l) of the parent
//     layer, C(m-l+d-2, d-1) consecutive amplitudes, staged in shared memory once per unit; its index is the
//     child's index minus a prefix sum of binomial differences (the same identity as slos_layer_kernel, applied to the
//     sub-system of modes l..m-1);
//   * the trailing modes of child c come from a table that depends on (d, c) only: the d-simplex over modes >= l is
//     the END of the d-simplex over all modes in lex order, so one table per d serves every unit.
// Units, stream offsets and tables depend on (m, k) only: built once on the host and cached in the context.
// Terms are added in ascending mode order with the reference's rounding sequence (no FMA contraction), like K3.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "fock.cuh"
#include "lwb_internal.h"
#include "slos_units.h"

namespace lwb {

constexpr int SU_THREADS = 256;
constexpr int SU_DMAX = 8;        // trailing photons per unit (their modes are packed into one 64-bit word)
constexpr int SU_NPMAX = 16;      // distinct prefix modes per unit
#ifndef LWB_SU_TMAX
#define LWB_SU_TMAX 8192
#endif
#ifndef LWB_SU_MINB
#define LWB_SU_MINB 3
#endif
constexpr uint32_t SU_TMAX = LWB_SU_TMAX;   // children per unit (large layers)
#ifndef LWB_SU_TMAX_SMALL
#define LWB_SU_TMAX_SMALL 4096
#endif
// Layers below this many children use smaller units: more CTAs, smaller parent blocks (measured: 16 modes / 8 photons
// 0.117 -> 0.091 ms), while the largest layers prefer long units (24 modes / 10 photons: 1.97 vs 2.25 ms).
constexpr uint64_t SU_SMALL_LAYER = 16u << 20;
static uint32_t unit_tmax(uint64_t children) { return children < SU_SMALL_LAYER ? (uint32_t)LWB_SU_TMAX_SMALL : SU_TMAX; }
#ifndef LWB_SU_ITEM
#define LWB_SU_ITEM 4096
#endif
constexpr uint32_t SU_ITEM = LWB_SU_ITEM;   // a CTA takes consecutive units until it holds this many children ...
constexpr uint32_t SU_ITEM_MIN = 256;  // ... fewer on small layers, so that every SM gets several CTAs
static uint32_t item_size(uint64_t children) {
    const uint64_t per = children / (148 * 8);
    return (uint32_t)std::min<uint64_t>(SU_ITEM, std::max<uint64_t>(SU_ITEM_MIN, per));
}

struct SlosUnit {
    uint64_t start;             // rank of the unit's first child in the child layer
    uint64_t pblock;            // rank of the first amplitude of B(P) in the parent layer
    uint64_t sbase[SU_NPMAX];   // stream s: parent(c) = parent[sbase[s] / 16 + c]: BYTE offset of (pend - T)
    uint32_t T, Tp;             // children; amplitudes of B(P)
    uint32_t lut;               // trailing modes of child c = lut_table[lut + c]
    uint32_t rbase;             // record kernel: child c uses record rbase + c (= rec_end(d) - T + c)
    uint32_t pad0;
    uint8_t d, ell, np, has_ell;  // trailing photons; last prefix mode (0 if none); distinct prefix modes; l in prefix?
    uint8_t pm[SU_NPMAX];       // distinct prefix modes, ascending
    uint8_t pmult[SU_NPMAX];    // their multiplicities in the prefix
};

struct SlosUnitsArgs {
    const SlosUnit* units;
    const uint2* items;      // [n_items] (first unit, one past last unit)
    const uint64_t* lut;
    const uint32_t* ht;      // [SU_DMAX][m + 1]: C(m - v + r - 1, r + 1) mod 2^32
    int m, k;
    const double2* parent;
    void* child;
    const double2* U;
    int in_mode;
    const double* sqrt_tab;
    int write_probs;
    uint32_t max_tp;
    uint64_t n_children, n_parents;  // layer sizes (debug bounds checks)
};

__device__ __forceinline__ double2 su_cmul_py(double2 a, double2 b) {  // CPython's complex_mul, no fma
    double2 r;
    r.x = __dsub_rn(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y));
    r.y = __dadd_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x));
    return r;
}

constexpr int SU_R = 2;  // children per thread and pass (strided by the block size: every access stays coalesced)

__device__ __forceinline__ void su_cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src));
}
__device__ __forceinline__ void su_cp_async4(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src));
}
__device__ __forceinline__ void su_cp_async_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Trailing part of one child, D trailing photons (compile-time: exact unrolling, constant shifts).  Terms in ascending
// mode order; the run of mode l, if the prefix holds l, was merged into the last prefix term by the caller.
template <int D>
__device__ __forceinline__ void su_trailing(uint64_t y, uint32_t c, int ell, bool has_ell, const uint32_t* __restrict__ ht,
                                            int mv, const double2* __restrict__ pb, const double2* __restrict__ ucol,
                                            const double* __restrict__ sq, double2& acc) {
    const uint32_t lo = (uint32_t)y, hi = (uint32_t)(y >> 32);
    uint32_t Q = 0;
    int prev = ell, run = 0;
#pragma unroll
    for (int i = 0; i < D; ++i) {
        const int yi = (int)(((i < 4 ? lo : hi) >> (8 * (i & 3))) & 0xffu);
        const int r = D - 1 - i;
        Q += ht[r * mv + prev] - ht[r * mv + yi];
        run += 1;
        bool last_of_run = true;
        if (i + 1 < D) last_of_run = (int)((((i + 1) < 4 ? lo : hi) >> (8 * ((i + 1) & 3))) & 0xffu) != yi;
        if (last_of_run) {
            if (!(has_ell && yi == ell)) {
                LWB_BOUNDS(c - Q < LWB_SU_TMAX);  // inside the staged parent block (the caller checks c - Q < Tp)
                double2 t = pb[c - Q];
                if (run > 1) {
                    const double f = sq[run];
                    t.x = __dmul_rn(f, t.x);
                    t.y = __dmul_rn(f, t.y);
                }
                const double2 term = su_cmul_py(t, ucol[yi]);
                acc.x = __dadd_rn(acc.x, term.x);
                acc.y = __dadd_rn(acc.y, term.y);
            }
            run = 0;
        }
        prev = yi;
    }
}

struct SuSmem {
    const double2* ucol;
    const double2* pb;
    const double* sq;
    const uint32_t* ht;
    int mv;
};

// One pass over NJ * SU_THREADS children of the current unit.  Every global load is unconditional (indices clamped to the
// unit) and issued ahead of its use: the trailing-mode table, then the prefix streams two streams ahead of the
// arithmetic - the kernel is bound by the latency of these loads, not by their bandwidth (profiles/r02_k3u_*).
template <int NJ>
__device__ __forceinline__ void su_pass(const SlosUnitsArgs& a, const SlosUnit& un, const SuSmem& sm, uint32_t base, int tid,
                                        bool* pb_ready) {
    const uint32_t T = un.T;
    const int d = un.d, np = un.np, ell = un.ell;
    const bool has_ell = un.has_ell != 0;
    const uint32_t ell4 = (uint32_t)ell * 0x01010101u;
    uint32_t c[NJ];
    uint64_t y[NJ];
    double2 acc[NJ];
    int lead[NJ];
    const uint64_t* lut = a.lut + un.lut;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const uint32_t cj = base + j * SU_THREADS + tid;
        c[j] = cj < T ? cj : T - 1;  // clamped: loads stay inside the unit, only the store is predicated
        y[j] = lut[c[j]];
    }
    // ---- prefix modes, ascending: one end-aligned coalesced stream each.  Loads run two streams ahead of the
    // arithmetic through three statically rotated buffers; sbase is a BYTE offset, so an address is one 64-bit add. ----
    const char* pj[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        pj[j] = reinterpret_cast<const char*>(a.parent + c[j]);
        acc[j] = make_double2(0.0, 0.0);
        // trailing photons sitting in mode l extend the prefix's run of l (unused bytes of y are 0xff)
        lead[j] = has_ell ? (__popc(__vcmpeq4((uint32_t)y[j], ell4)) + __popc(__vcmpeq4((uint32_t)(y[j] >> 32), ell4))) >> 3 : 0;
    }
    double2 b0[NJ], b1[NJ], b2[NJ];
    auto load = [&](double2 (&buf)[NJ], int s) {
        if (s < np) {
            const uint64_t off = un.sbase[s];
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                LWB_BOUNDS(off / 16 + c[j] < a.n_parents);
                buf[j] = *reinterpret_cast<const double2*>(pj[j] + off);
            }
        }
    };
    auto use = [&](const double2 (&buf)[NJ], int s) {
        if (s >= np) return;
        const double2 uv = sm.ucol[un.pm[s]];
        const int pmul = un.pmult[s];
        if (s < np - 1) {  // multiplicity is the prefix's own: uniform for the whole unit
            const bool scale = pmul > 1;
            const double f = sm.sq[pmul];
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                double2 v = buf[j];
                if (scale) { v.x = __dmul_rn(f, v.x); v.y = __dmul_rn(f, v.y); }
                const double2 term = su_cmul_py(v, uv);
                acc[j].x = __dadd_rn(acc[j].x, term.x);
                acc[j].y = __dadd_rn(acc[j].y, term.y);
            }
        } else {  // the last prefix mode is l: trailing photons in l add to its multiplicity, child by child
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int mult = pmul + lead[j];
                double2 v = buf[j];
                if (mult > 1) {
                    const double f = sm.sq[mult];
                    v.x = __dmul_rn(f, v.x);
                    v.y = __dmul_rn(f, v.y);
                }
                const double2 term = su_cmul_py(v, uv);
                acc[j].x = __dadd_rn(acc[j].x, term.x);
                acc[j].y = __dadd_rn(acc[j].y, term.y);
            }
        }
    };
    load(b0, 0);
    load(b1, 1);
    for (int s = 0; s < np; s += 3) {
        load(b2, s + 2);
        use(b0, s);
        load(b0, s + 3);
        use(b1, s + 1);
        load(b1, s + 4);
        use(b2, s + 2);
    }
    if (!*pb_ready) {  // first pass of the unit: the parent block was copied asynchronously behind the streams
        su_cp_async_wait();
        __syncthreads();
        *pb_ready = true;
    }
    // ---- trailing modes, ascending: parents inside the staged block B(P) ----
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        switch (d) {  // block-uniform
            case 1: su_trailing<1>(y[j], c[j], ell, has_ell, sm.ht, sm.mv, sm.pb, sm.ucol, sm.sq, acc[j]); break;
            case 2: su_trailing<2>(y[j], c[j], ell, has_ell, sm.ht, sm.mv, sm.pb, sm.ucol, sm.sq, acc[j]); break;
            case 3: su_trailing<3>(y[j], c[j], ell, has_ell, sm.ht, sm.mv, sm.pb, sm.ucol, sm.sq, acc[j]); break;
            case 4: su_trailing<4>(y[j], c[j], ell, has_ell, sm.ht, sm.mv, sm.pb, sm.ucol, sm.sq, acc[j]); break;
            case 5: su_trailing<5>(y[j], c[j], ell, has_ell, sm.ht, sm.mv, sm.pb, sm.ucol, sm.sq, acc[j]); break;
            case 6: su_trailing<6>(y[j], c[j], ell, has_ell, sm.ht, sm.mv, sm.pb, sm.ucol, sm.sq, acc[j]); break;
            case 7: su_trailing<7>(y[j], c[j], ell, has_ell, sm.ht, sm.mv, sm.pb, sm.ucol, sm.sq, acc[j]); break;
            default: su_trailing<8>(y[j], c[j], ell, has_ell, sm.ht, sm.mv, sm.pb, sm.ucol, sm.sq, acc[j]); break;
        }
        if (base + j * SU_THREADS + tid < T) {
            const uint64_t rank = un.start + c[j];
            LWB_BOUNDS(rank < a.n_children);
            if (a.write_probs)  // abs(a) ** 2 (slos.py:74); x*x + y*y differs from hypot(x, y)^2 by <= 2 ulp
                reinterpret_cast<double*>(a.child)[rank] = __dadd_rn(__dmul_rn(acc[j].x, acc[j].x), __dmul_rn(acc[j].y, acc[j].y));
            else
                reinterpret_cast<double2*>(a.child)[rank] = acc[j];
        }
    }
}

__global__ void __launch_bounds__(SU_THREADS, LWB_SU_MINB) slos_units_kernel(SlosUnitsArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int m = a.m, mv = m + 1;
    double2* ucol = reinterpret_cast<double2*>(smem_raw);                 // [m]
    double2* pb = ucol + m;                                               // [max_tp] parent block
    double* sq = reinterpret_cast<double*>(pb + a.max_tp);                // [k + 1] (+ pad)
    uint32_t* ht = reinterpret_cast<uint32_t*>(sq + ((a.k + 2) & ~1));    // [SU_DMAX][mv]
    __shared__ __align__(16) SlosUnit un2[2];  // current unit and the prefetched next one
    const int tid = threadIdx.x;
    for (int e = tid; e < m; e += SU_THREADS) ucol[e] = a.U[(size_t)e * m + a.in_mode];
    for (int e = tid; e <= a.k; e += SU_THREADS) sq[e] = a.sqrt_tab[e];
    for (int e = tid; e < SU_DMAX * mv; e += SU_THREADS) ht[e] = a.ht[e];
    const uint2 item = a.items[blockIdx.x];
    constexpr int UW = (int)(sizeof(SlosUnit) / 4);
    if (tid < UW) su_cp_async4(reinterpret_cast<uint32_t*>(&un2[0]) + tid, reinterpret_cast<const uint32_t*>(a.units + item.x) + tid);
    SuSmem sm{ucol, pb, sq, ht, mv};
    int cur = 0;
    for (uint32_t u = item.x; u < item.y; ++u, cur ^= 1) {
        su_cp_async_wait();
        __syncthreads();  // unit descriptor `cur` has arrived; everybody is done with the previous unit's `pb`
        const SlosUnit& un = un2[cur];
        if (u + 1 < item.y && tid < UW)  // prefetch the next descriptor behind this unit's work
            su_cp_async4(reinterpret_cast<uint32_t*>(&un2[cur ^ 1]) + tid, reinterpret_cast<const uint32_t*>(a.units + u + 1) + tid);
        // parent block -> shared memory, asynchronously: only the trailing part needs it, after the prefix streams
        const uint32_t T = un.T, Tp = un.Tp;
        const double2* pblk = a.parent + un.pblock;
        for (uint32_t i = tid; i < Tp; i += SU_THREADS) su_cp_async16(pb + i, pblk + i);
        bool pb_ready = false;
        for (uint32_t base = 0; base < T; base += SU_THREADS * SU_R) {
            if (T - base > SU_THREADS) su_pass<2>(a, un, sm, base, tid, &pb_ready);
            else su_pass<1>(a, un, sm, base, tid, &pb_ready);
        }
    }
}

// ================================================================================================================
// K3u record kernel (option "slos_units" = 4): the same units, but no child derives an index or a multiplicity.
// Every trailing part (d <= 8 photons over the free modes) has a host-built RECORD of one 32-bit word per run of equal
// modes: the position of the parent "trailing part minus one photon of the run", counted from the END of the parent
// block B(P), and the byte offset of the weight w[mode][run length] in shared memory, where
// w[mode][mult] = U[mode, in_mode] * sqrt(mult) and w[.][0] = 0.  Records are "counted from the end", hence shared by
// every unit with the same d whatever its first free mode l: the child c of a unit with T children uses record
// rec_end(d) - T + c.  Unused words of a record point at w[0][0] = 0 and at a valid parent, so the trailing loop has no
// per-lane branch.  The prefix modes are the end-aligned global streams of the unit kernel above; a thread handles
// two children 256 apart so that one stream address and one weight serve both.  The only data-dependent step left is
// the run of l: trailing photons in l (the first run of the record, when its mode is l) join the last prefix stream.
// A term is parent * w, products and sums rounded one by one as in the gather kernel and in the same order, so the
// amplitudes of collision-free paths are bit-identical to it; where a multiplicity > 1 occurs, sqrt(mult) is applied
// to the weight instead of the parent, which changes the rounding (~1e-16 relative), not the value.
constexpr int SU_WMAX = 2048;     // entries of w: byte offsets fit 16 bits
constexpr uint32_t SU_NEVER = 0x80000000u;
#ifndef LWB_SU4_J
#define LWB_SU4_J 2
#endif
#ifndef LWB_SU4_PF
#define LWB_SU4_PF 2
#endif
constexpr int SU4_J = LWB_SU4_J;    // children per thread and pass, SU_THREADS apart
constexpr int SU4_PF = LWB_SU4_PF;  // prefix streams loaded ahead
#ifndef LWB_SU4_MINB
#define LWB_SU4_MINB LWB_SU_MINB
#endif

struct SuTables {
    const uint4* rec_a;     // words 0-3 of every record: bits 0-15 parent position from the end (1 = last), 16-31 weight offset
    const uint4* rec_b;     // words 4-7 (classes with d > 4)
};

// acc += v * w with every product and sum rounded separately, like the reference's complex arithmetic: two terms that
// cancel by symmetry (Hong-Ou-Mandel: the reference's own tests ask for an exact 0) then cancel exactly, which a fused
// multiply-add chain does not do.
__device__ __forceinline__ void su_add_term(double2& acc, const double2 v, const double2 w) {
    const double2 t = su_cmul_py(v, w);
    acc.x = __dadd_rn(acc.x, t.x);
    acc.y = __dadd_rn(acc.y, t.y);
}

#ifndef LWB_SU4_LD
#define LWB_SU4_LD 0
#endif
// a prefix-stream amplitude: read once by this thread, never again by this CTA
__device__ __forceinline__ double2 su4_stream_load(const char* p) {
#if LWB_SU4_LD == 1
    return __ldcs(reinterpret_cast<const double2*>(p));
#elif LWB_SU4_LD == 2
    return __ldg(reinterpret_cast<const double2*>(p));
#elif LWB_SU4_LD == 3
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
#else
    return *reinterpret_cast<const double2*>(p);
#endif
}

// J children per thread (c0, c0 + 256, ...) of the unit `un`; children past T - 1 recompute the last one, store nothing
template <int J>
__device__ __forceinline__ void su4_children(const SlosUnitsArgs& a, const SuTables& tb, const SlosUnit& un, uint32_t c0,
                                             const char* wbase, const char* pbase, const uint64_t* sptr, const uint32_t* pw,
                                             uint32_t ell_w16, uint32_t k1_16) {
    const uint32_t T = un.T, be = un.Tp * 16u;
    const int d = un.d, np = un.np;
    uint32_t cj[J];
    uint4 ra[J], rb[J];
    double2 acc[J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
        cj[j] = min(c0 + j * SU_THREADS, T - 1);
        ra[j] = __ldg(tb.rec_a + un.rbase + cj[j]);
        acc[j] = make_double2(0.0, 0.0);
    }
    // ---- prefix modes: end-aligned global streams, SU4_PF loads ahead ----
    double2 buf[SU4_PF][J];
    uint32_t dj[J];
#pragma unroll
    for (int j = 0; j < J; ++j) dj[j] = (cj[j] - c0) * 16u;
    auto load = [&](double2 (&bf)[J], int s_) {
        if (s_ < np) {
            const char* p = reinterpret_cast<const char*>(sptr[s_]) + (uint64_t)c0 * 16u;
#pragma unroll
            for (int j = 0; j < J; ++j) bf[j] = su4_stream_load(p + dj[j]);
        }
    };
#pragma unroll
    for (int i = 0; i < SU4_PF; ++i) load(buf[i], i);
    // the run of l: a first run of the record sitting in l joins the last prefix stream
    uint32_t lead[J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const uint32_t t = (ra[j].x >> 16) - ell_w16;     // (run length) * 16 when the first run's mode is l
        const bool merged = t < k1_16;
        lead[j] = merged ? t : 0u;
        if (merged) ra[j].x &= 0xffffu;                   // ... and then that run has no term of its own
    }
    auto use = [&](const double2 (&bf)[J], int s_) {
        if (s_ < np) {
            const uint32_t wo = pw[s_];
            if (s_ == np - 1) {
#pragma unroll
                for (int j = 0; j < J; ++j) su_add_term(acc[j], bf[j], *reinterpret_cast<const double2*>(wbase + wo + lead[j]));
            } else {
                const double2 w = *reinterpret_cast<const double2*>(wbase + wo);
#pragma unroll
                for (int j = 0; j < J; ++j) su_add_term(acc[j], bf[j], w);
            }
        }
    };
    for (int s_ = 0; s_ < np; s_ += SU4_PF) {
#pragma unroll
        for (int i = 0; i < SU4_PF; ++i) {
            use(buf[i], s_ + i);
            load(buf[i], s_ + i + SU4_PF);
        }
    }
    // ---- trailing runs: parents in the staged block, positions and weights from the record ----
    auto term = [&](double2& ac, uint32_t word) {
        LWB_BOUNDS((word & 0xffffu) >= 1 && (word & 0xffffu) * 16u <= be);
        su_add_term(ac, *reinterpret_cast<const double2*>(pbase + (be - (word & 0xffffu) * 16u)),
                    *reinterpret_cast<const double2*>(wbase + (word >> 16)));
    };
    if (d > 4) {  // second half of the records: in flight while the first half is used
#pragma unroll
        for (int j = 0; j < J; ++j) rb[j] = __ldg(tb.rec_b + un.rbase + cj[j]);
    }
#pragma unroll
    for (int j = 0; j < J; ++j) {
        term(acc[j], ra[j].x);
        if (d > 1) term(acc[j], ra[j].y);
        if (d > 2) term(acc[j], ra[j].z);
        if (d > 3) term(acc[j], ra[j].w);
    }
    if (d > 4) {
#pragma unroll
        for (int j = 0; j < J; ++j) {
            term(acc[j], rb[j].x);
            if (d > 5) term(acc[j], rb[j].y);
            if (d > 6) term(acc[j], rb[j].z);
            if (d > 7) term(acc[j], rb[j].w);
        }
    }
#pragma unroll
    for (int j = 0; j < J; ++j) {
        if (c0 + j * SU_THREADS >= T) break;
        const uint64_t rank = un.start + cj[j];
        LWB_BOUNDS(rank < a.n_children);
        if (a.write_probs)
            reinterpret_cast<double*>(a.child)[rank] = __dadd_rn(__dmul_rn(acc[j].x, acc[j].x), __dmul_rn(acc[j].y, acc[j].y));
        else
            reinterpret_cast<double2*>(a.child)[rank] = acc[j];
    }
}

__global__ void __launch_bounds__(SU_THREADS, LWB_SU4_MINB) slos_units4_kernel(SlosUnitsArgs a, SuTables tb) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int m = a.m, k1 = a.k + 1;
    double2* wtab = reinterpret_cast<double2*>(smem_raw);                 // [m][k + 1]
    double2* pb = wtab + m * k1;                                          // [max_tp] parent block
    __shared__ __align__(16) SlosUnit un2[2];
    __shared__ uint64_t sptr[SU_NPMAX];                                   // prefix stream s: address of its first amplitude
    __shared__ uint32_t pw[SU_NPMAX];                                     // ... and byte offset of its weight in wtab
    const int tid = threadIdx.x;
    for (int i = tid; i < m * k1; i += SU_THREADS) {
        const int mode = i / k1, mult = i - mode * k1;
        const double2 u = a.U[(size_t)mode * m + a.in_mode];
        const double f = a.sqrt_tab[mult];
        wtab[i] = mult ? make_double2(__dmul_rn(u.x, f), __dmul_rn(u.y, f)) : make_double2(0.0, 0.0);
    }
    const uint2 item = a.items[blockIdx.x];
    constexpr int UW = (int)(sizeof(SlosUnit) / 4);
    if (tid < UW) su_cp_async4(reinterpret_cast<uint32_t*>(&un2[0]) + tid, reinterpret_cast<const uint32_t*>(a.units + item.x) + tid);
    int cur = 0;
    const char* wbase = reinterpret_cast<const char*>(wtab);
    const char* pbase = reinterpret_cast<const char*>(pb);
    const uint32_t k1_16 = (uint32_t)k1 * 16u;
    for (uint32_t u = item.x; u < item.y; ++u, cur ^= 1) {
        su_cp_async_wait();
        __syncthreads();  // descriptor `cur` has arrived; everybody is done with the previous unit's `pb`, `sptr`, `pw`
        const SlosUnit& un = un2[cur];
        const uint32_t T = un.T, Tp = un.Tp;
        const double2* pblk = a.parent + un.pblock;
        for (uint32_t i = tid; i < Tp; i += SU_THREADS) su_cp_async16(pb + i, pblk + i);
        if (tid < un.np) {
            pw[tid] = (uint32_t)(un.pm[tid] * k1 + un.pmult[tid]) * 16u;
            sptr[tid] = reinterpret_cast<uint64_t>(a.parent) + un.sbase[tid];
        }
        // weight offset of (l, 0): a record whose first run is in l merges with the last prefix stream
        const uint32_t ell_w16 = un.has_ell ? (uint32_t)un.ell * k1_16 : SU_NEVER;
        su_cp_async_wait();
        __syncthreads();  // parent block, `sptr` and `pw` staged
        if (u + 1 < item.y && tid < UW)
            su_cp_async4(reinterpret_cast<uint32_t*>(&un2[cur ^ 1]) + tid, reinterpret_cast<const uint32_t*>(a.units + u + 1) + tid);
        for (uint32_t c0 = tid; c0 < T; c0 += SU4_J * SU_THREADS) {
            // warp-uniform choice: some lane of this warp has a second child, or none has
            if (SU4_J > 1 && (c0 & ~31u) + SU_THREADS < T)
                su4_children<SU4_J>(a, tb, un, c0, wbase, pbase, sptr, pw, ell_w16, k1_16);
            else
                su4_children<1>(a, tb, un, c0, wbase, pbase, sptr, pw, ell_w16, k1_16);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host plan: units, items and tables of one layer (m modes, k photons in the child layer)
struct SlosLayerPlan {
    int m = 0, k = 0;
    uint32_t n_units = 0, n_items = 0, max_tp = 0;
    SlosUnit* d_units = nullptr;
    uint2* d_items = nullptr;
    uint64_t* d_lut = nullptr;
    uint32_t* d_ht = nullptr;
    uint4* d_rec_a = nullptr;    // record kernel; has_tables = false when not representable (the unit kernel serves the layer)
    uint4* d_rec_b = nullptr;
    SuTables tables;
    bool has_tables = false;
};

struct SlosUnitsState {
    std::map<std::pair<int, int>, SlosLayerPlan> plans;
    size_t bytes = 0;   // device memory held by the cached plans
};

// A process that sweeps many circuit shapes would otherwise keep every plan (C3's ten layers: ~30 MB) for ever: once the
// cache passes this size it is emptied before the next plan is built (LWB_SLOS_PLAN_CACHE_MB overrides the 1 GiB).
static size_t plan_cache_cap() {
    static const size_t cap = [] {
        const char* e = getenv("LWB_SLOS_PLAN_CACHE_MB");
        const long mb = e ? atol(e) : 1024;
        return (size_t)(mb > 0 ? mb : 1024) << 20;
    }();
    return cap;
}

static void free_plans(SlosUnitsState* st) {
    for (auto& kv : st->plans) {   // cudaFree waits for the kernels that still read a plan
        cudaFree(kv.second.d_units);
        cudaFree(kv.second.d_items);
        cudaFree(kv.second.d_lut);
        cudaFree(kv.second.d_ht);
        cudaFree(kv.second.d_rec_a);
        cudaFree(kv.second.d_rec_b);
    }
    st->plans.clear();
    st->bytes = 0;
}

SlosUnitsState* slos_units_create() { return new SlosUnitsState(); }
void slos_units_destroy(SlosUnitsState* st) {
    if (!st) return;
    free_plans(st);
    delete st;
}

bool slos_units_supported(int m, int k) { return m >= 1 && m <= LWB_MAX_MODES && k >= 1 && k <= LWB_MAX_PHOTONS; }

namespace {
uint64_t simplex(const uint64_t* bin, int n_free, int d) { return binom(bin, n_free + d - 1, d); }  // C(n_free+d-1, d)

struct Builder {
    const uint64_t* bin;
    int m, k;
    uint32_t tmax = SU_TMAX;
    std::vector<SlosUnit> units;
    int nf_max[SU_DMAX + 1];   // per d: largest n_free among the units
    int x[LWB_MAX_PHOTONS + 2];
    uint64_t next_start = 0;
    bool ok = true;

    void emit(int depth) {  // prefix x[0..depth), d = k - depth trailing photons
        const int d = k - depth, ell = depth ? x[depth - 1] : 0, n_free = m - ell;
        SlosUnit u;
        memset(&u, 0, sizeof u);
        u.start = next_start;
        u.T = (uint32_t)simplex(bin, n_free, d);
        u.Tp = (uint32_t)simplex(bin, n_free, d - 1);
        u.d = (uint8_t)d;
        u.ell = (uint8_t)ell;
        next_start += u.T;
        nf_max[d] = std::max(nf_max[d], n_free);
        // parent block: (P, l, ..., l) with d-1 trailing photons, rank in the parent layer (m, k-1)
        int y[LWB_MAX_PHOTONS + 2];
        for (int i = 0; i < depth; ++i) y[i] = x[i];
        for (int i = depth; i < k - 1; ++i) y[i] = ell;
        u.pblock = lex_rank(bin, y, k - 1, m);
        // streams: remove one photon of each distinct prefix mode; the parents of the unit's children are the
        // LAST T states under the shortened prefix
        int np = 0;
        for (int i = 0; i < depth;) {
            int j = i;
            while (j < depth && x[j] == x[i]) ++j;
            if (np >= SU_NPMAX) { ok = false; return; }
            u.pm[np] = (uint8_t)x[i];
            u.pmult[np] = (uint8_t)(j - i);
            int q = 0;
            for (int t = 0; t < depth; ++t)
                if (t != i) y[q++] = x[t];          // drop one copy of mode x[i]
            for (int t = 0; t < d; ++t) y[q++] = m - 1;  // last state of the shortened prefix's subtree
            const uint64_t pend = lex_rank(bin, y, k - 1, m) + 1;
            u.sbase[np] = (pend - u.T) * 16;  // byte offset into the parent layer
            ++np;
            i = j;
        }
        u.np = (uint8_t)np;
        u.has_ell = (uint8_t)(depth > 0 ? 1 : 0);  // the last prefix mode IS l
        units.push_back(u);
    }

    void walk(int depth) {
        if (!ok) return;
        const int d = k - depth, ell = depth ? x[depth - 1] : 0;
        const uint64_t t = simplex(bin, m - ell, d);
        if (d <= SU_DMAX && t <= tmax && d >= 1) { emit(depth); return; }
        for (int v = ell; v < m; ++v) {
            x[depth] = v;
            walk(depth + 1);
        }
    }
};

// Records of the record kernel (see slos_units4_kernel): per d the d-simplex over the modes [m - nf_max[d], m) in lex
// order, one record each.  Sets rbase of every unit.
struct SuHostTables {
    std::vector<uint4> rec_a, rec_b;
    uint32_t rec_off[SU_DMAX + 1] = {0}, rec_size[SU_DMAX + 1] = {0};
};

bool build_tables(const uint64_t* bin, int m, int k, Builder& b, SuHostTables& t) {
    if (m * (k + 1) > SU_WMAX) return false;
    const int k1 = k + 1;
    for (int d = 1; d <= SU_DMAX; ++d) {
        if (!b.nf_max[d]) continue;
        const int nf = b.nf_max[d], lo = m - nf;
        const uint64_t size = simplex(bin, nf, d), psize = simplex(bin, nf, d - 1);
        if (size > 0x7fffffffull || psize > 0xffff) return false;
        t.rec_off[d] = (uint32_t)t.rec_a.size();
        t.rec_size[d] = (uint32_t)size;
        int y[SU_DMAX];
        for (int i = 0; i < d; ++i) y[i] = lo;
        for (uint64_t c = 0; c < size; ++c) {
            uint32_t w[SU_DMAX];
            for (int q = 0; q < SU_DMAX; ++q) w[q] = 1;  // unused: position 1 from the end (valid), weight w[0][0] = 0
            int n_runs = 0;
            for (int q = 0; q < d;) {
                int j = q;
                while (j < d && y[j] == y[q]) ++j;
                int z[SU_DMAX], n = 0;
                for (int i = 0; i < d; ++i)
                    if (i != q) z[n++] = y[i] - lo;   // the trailing part minus one photon of this run
                const uint64_t r = d > 1 ? lex_rank(bin, z, d - 1, nf) : 0;
                const uint32_t pe = (uint32_t)(psize - r);  // counted from the end, the parent itself included
                const uint32_t w16 = (uint32_t)(y[q] * k1 + (j - q)) * 16u;
                w[n_runs++] = pe | (w16 << 16);
                q = j;
            }
            t.rec_a.push_back(make_uint4(w[0], w[1], w[2], w[3]));
            t.rec_b.push_back(make_uint4(w[4], w[5], w[6], w[7]));
            for (int i = d - 1; i >= 0; --i)  // lex successor
                if (y[i] < m - 1) {
                    const int v = y[i] + 1;
                    for (int q = i; q < d; ++q) y[q] = v;
                    break;
                }
        }
    }
    for (SlosUnit& un : b.units) {
        if (un.T > t.rec_size[un.d]) return false;
        un.rbase = t.rec_off[un.d] + t.rec_size[un.d] - un.T;
    }
    return true;
}

// Replays slos_units4_kernel's index arithmetic for one unit on the host and compares every parent position it would
// read with the exact lex rank of the child minus that photon (and every weight offset with the true multiplicity).
int check_unit_v4(const uint64_t* bin, int m, int k, const SlosUnit& un, const SuHostTables& tb) {
    const int d = un.d, np = un.np, k1 = k + 1;
    int x[LWB_MAX_PHOTONS + 2], depth = 0;
    for (int s = 0; s < np; ++s)
        for (int q = 0; q < un.pmult[s]; ++q) x[depth++] = un.pm[s];
    if (depth + d != k) return LWB_E_ARG;
    if ((un.has_ell != 0) != (np > 0)) return LWB_E_ARG;   // the kernel merges into the LAST prefix stream
    if (np > 0 && un.pm[np - 1] != un.ell) return LWB_E_ARG;
    const uint32_t T = un.T, Tp = un.Tp, k1_16 = (uint32_t)k1 * 16u;
    const uint32_t ell_w16 = un.has_ell ? (uint32_t)un.ell * k1_16 : SU_NEVER;
    if ((size_t)un.rbase + T > tb.rec_a.size()) return LWB_E_ARG;
    for (uint32_t c = 0; c < T; ++c) {
        const uint4 ra = tb.rec_a[un.rbase + c], rb = tb.rec_b[un.rbase + c];
        uint32_t word[SU_DMAX] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
        int tq = 0;
        for (int q = 0; q < SU_DMAX; ++q) {  // the trailing part spelled by the runs of the record
            const uint32_t w16 = word[q] >> 16;
            if (w16 % 16) return LWB_E_ARG;
            const int wi = (int)(w16 / 16);
            if (!wi) continue;
            if (q >= d) return LWB_E_ARG;     // the kernel reads d words only
            for (int r = 0; r < wi % k1; ++r) {
                if (tq >= d) return LWB_E_ARG;
                x[depth + tq++] = wi / k1;
            }
        }
        if (tq != d) return LWB_E_ARG;
        for (int i = 1; i < k; ++i)
            if (x[i] < x[i - 1] || x[i] >= m) return LWB_E_ARG;
        if (lex_rank(bin, x, k, m) != un.start + c) return LWB_E_ARG;
        const uint32_t t = (word[0] >> 16) - ell_w16;
        const bool merged = t < k1_16;
        const uint32_t lead = merged ? t : 0u;
        if (merged) word[0] &= 0xffffu;
        // every (weight, parent position) the kernel would use, in its order
        int seen_modes[LWB_MAX_PHOTONS + 2] = {0}, n_seen = 0;
        auto verify = [&](uint32_t w16, uint64_t parent_rank) -> bool {
            const int mode = (int)(w16 / 16) / k1, mult = (int)(w16 / 16) % k1;
            int true_mult = 0, at = -1;
            for (int i = 0; i < k; ++i)
                if (x[i] == mode) { ++true_mult; at = i; }
            if (true_mult != mult || at < 0) return false;
            if (n_seen && seen_modes[n_seen - 1] >= mode) return false;  // ascending, each mode once
            seen_modes[n_seen++] = mode;
            int y[LWB_MAX_PHOTONS + 2], n = 0;
            for (int i = 0; i < k; ++i)
                if (i != at) y[n++] = x[i];
            return lex_rank(bin, y, k - 1, m) == parent_rank;
        };
        for (int s = 0; s < np; ++s) {
            const uint32_t w16 = (uint32_t)(un.pm[s] * k1 + un.pmult[s]) * 16u + (s == np - 1 ? lead : 0u);
            if (un.sbase[s] % 16 || !verify(w16, un.sbase[s] / 16 + c)) return LWB_E_ARG;
        }
        for (int q = 0; q < d; ++q) {
            const uint32_t pe = word[q] & 0xffffu, w16 = word[q] >> 16;
            if (pe == 0 || pe > Tp) return LWB_E_ARG;  // read even for a zero weight
            if (w16 && !verify(w16, un.pblock + Tp - pe)) return LWB_E_ARG;
        }
        int distinct = 0;
        for (int i = 0; i < k; ++i) distinct += i == 0 || x[i] != x[i - 1];
        if (n_seen != distinct) return LWB_E_ARG;  // no mode of the child was skipped
    }
    return 0;
}
}  // namespace

static int build_plan(SlosUnitsState* st, int m, int k, SlosLayerPlan** out, std::string* err) {
    auto key = std::make_pair(m, k);
    auto it = st->plans.find(key);
    if (it != st->plans.end()) { *out = &it->second; return 0; }
    if (st->bytes > plan_cache_cap()) free_plans(st);
    const uint64_t* bin = host_bin();
    Builder b;
    b.bin = bin; b.m = m; b.k = k; b.tmax = unit_tmax(binom(bin, m + k - 1, k));
    for (int d = 0; d <= SU_DMAX; ++d) b.nf_max[d] = 0;
    b.walk(0);
    if (!b.ok || b.units.empty() || b.units.size() >= (1ull << 31)) { *err = "slos units: plan not representable"; return LWB_E_LIMIT; }
    // trailing-mode tables: for every d the d-simplex over modes [m - nf_max[d], m) in lex order, one byte per photon
    std::vector<uint64_t> lut;
    uint64_t lut_base[SU_DMAX + 1] = {0}, lut_size[SU_DMAX + 1] = {0};
    for (int d = 1; d <= SU_DMAX; ++d) {
        if (!b.nf_max[d]) continue;
        lut_base[d] = lut.size();
        lut_size[d] = simplex(bin, b.nf_max[d], d);
        int y[SU_DMAX];
        const int lo = m - b.nf_max[d];
        for (int i = 0; i < d; ++i) y[i] = lo;
        for (uint64_t c = 0; c < lut_size[d]; ++c) {
            uint64_t w = ~0ull;  // unused bytes stay 0xff (never equal to a mode)
            for (int i = 0; i < d; ++i) w = (w & ~(0xffull << (8 * i))) | ((uint64_t)y[i] << (8 * i));
            lut.push_back(w);
            for (int i = d - 1; i >= 0; --i)  // lex successor
                if (y[i] < m - 1) {
                    const int v = y[i] + 1;
                    for (int q = i; q < d; ++q) y[q] = v;
                    break;
                }
        }
    }
    SlosLayerPlan pl;
    pl.m = m; pl.k = k;
    for (auto& u : b.units) {
        u.lut = (uint32_t)(lut_base[u.d] + lut_size[u.d] - u.T);
        pl.max_tp = std::max(pl.max_tp, u.Tp);
    }
    std::vector<uint2> items;
    const uint32_t item = item_size(b.next_start);
    for (uint32_t u = 0; u < b.units.size();) {
        uint32_t e = u, tot = 0;
        while (e < b.units.size() && (e == u || tot + b.units[e].T <= item)) tot += b.units[e++].T;
        items.push_back(make_uint2(u, e));
        u = e;
    }
    std::vector<uint32_t> ht((size_t)SU_DMAX * (m + 1));
    for (int r = 0; r < SU_DMAX; ++r)
        for (int v = 0; v <= m; ++v) ht[(size_t)r * (m + 1) + v] = (uint32_t)binom(bin, m - v + r - 1, r + 1);
    pl.n_units = (uint32_t)b.units.size();
    pl.n_items = (uint32_t)items.size();
    SuHostTables ht4;   // record kernel: records and the units' rbase (before the units are uploaded)
    memset(&pl.tables, 0, sizeof pl.tables);
    const bool has4 = build_tables(bin, m, k, b, ht4);
#define SU_CU(call)                                                                   \
    do {                                                                              \
        cudaError_t e_ = (call);                                                      \
        if (e_ != cudaSuccess) { *err = std::string(#call ": ") + cudaGetErrorString(e_); return LWB_E_CUDA; } \
    } while (0)
    SU_CU(cudaMalloc(&pl.d_units, b.units.size() * sizeof(SlosUnit)));
    SU_CU(cudaMalloc(&pl.d_items, items.size() * sizeof(uint2)));
    SU_CU(cudaMalloc(&pl.d_lut, std::max<size_t>(lut.size(), 1) * 8));
    SU_CU(cudaMalloc(&pl.d_ht, ht.size() * 4));
    SU_CU(cudaMemcpy(pl.d_units, b.units.data(), b.units.size() * sizeof(SlosUnit), cudaMemcpyHostToDevice));
    SU_CU(cudaMemcpy(pl.d_items, items.data(), items.size() * sizeof(uint2), cudaMemcpyHostToDevice));
    SU_CU(cudaMemcpy(pl.d_lut, lut.data(), lut.size() * 8, cudaMemcpyHostToDevice));
    SU_CU(cudaMemcpy(pl.d_ht, ht.data(), ht.size() * 4, cudaMemcpyHostToDevice));
    if (has4) {
        const size_t nrec = std::max<size_t>(ht4.rec_a.size(), 1);
        SU_CU(cudaMalloc(&pl.d_rec_a, nrec * sizeof(uint4)));
        SU_CU(cudaMalloc(&pl.d_rec_b, nrec * sizeof(uint4)));
        SU_CU(cudaMemcpy(pl.d_rec_a, ht4.rec_a.data(), ht4.rec_a.size() * sizeof(uint4), cudaMemcpyHostToDevice));
        SU_CU(cudaMemcpy(pl.d_rec_b, ht4.rec_b.data(), ht4.rec_b.size() * sizeof(uint4), cudaMemcpyHostToDevice));
        pl.tables.rec_a = pl.d_rec_a;
        pl.tables.rec_b = pl.d_rec_b;
        pl.has_tables = true;
    }
    st->bytes += b.units.size() * sizeof(SlosUnit) + items.size() * sizeof(uint2) + std::max<size_t>(lut.size(), 1) * 8 +
                 ht.size() * 4 + (has4 ? 2 * std::max<size_t>(ht4.rec_a.size(), 1) * sizeof(uint4) : 0);
    st->plans[key] = pl;
    *out = &st->plans[key];
    return 0;
}

int slos_units_layer(SlosUnitsState* st, cudaStream_t stream, int m, int k, const double2* parent, void* child,
                     const double2* U, int in_mode, const double* sqrt_tab, int write_probs, uint64_t* launches,
                     std::string* err, int version) {
    SlosLayerPlan* pl = nullptr;
    int rc = build_plan(st, m, k, &pl, err);
    if (rc) return rc;
    SlosUnitsArgs a;
    a.units = pl->d_units; a.items = pl->d_items; a.lut = pl->d_lut; a.ht = pl->d_ht;
    a.m = m; a.k = k; a.parent = parent; a.child = child; a.U = U; a.in_mode = in_mode; a.sqrt_tab = sqrt_tab;
    a.write_probs = write_probs; a.max_tp = pl->max_tp;
    a.n_children = binom(host_bin(), m + k - 1, k);
    a.n_parents = k >= 1 ? binom(host_bin(), m + k - 2, k - 1) : 1;
    const size_t smem = (size_t)m * 16 + (size_t)pl->max_tp * 16 + (size_t)((k + 2) & ~1) * 8 + (size_t)SU_DMAX * (m + 1) * 4;
    if (smem > 200 * 1024) { *err = "slos units: parent block does not fit shared memory"; return LWB_E_LIMIT; }
    const size_t smem4 = ((size_t)m * (k + 1) + (size_t)pl->max_tp) * 16;  // weights + parent block
    if (version >= 4 && pl->has_tables && smem4 <= 200 * 1024) {  // record kernel
        if (smem4 + 1024 > 48 * 1024)
            SU_CU(cudaFuncSetAttribute(slos_units4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4));
        slos_units4_kernel<<<pl->n_items, SU_THREADS, smem4, stream>>>(a, pl->tables);
        *launches += 1;
        SU_CU(cudaGetLastError());
        return 0;
    }
    if (smem + 1024 > 48 * 1024) SU_CU(cudaFuncSetAttribute(slos_units_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    slos_units_kernel<<<pl->n_items, SU_THREADS, smem, stream>>>(a);
    *launches += 1;
    SU_CU(cudaGetLastError());
    return 0;
}

// Host-only self-check of a plan (tests, no device): for every child of every unit rebuild its mode list from the
// unit's prefix and the trailing-mode table and verify, with the exact lex ranks of fock.cuh, the three index rules
// the kernel relies on: child rank = start + c; prefix-removal parent = sbase[s] + c; trailing-removal parent =
// pblock + c - (prefix sum of binomial differences).  Returns 0 when every index agrees.
int slos_units_plan_check(int m, int k, uint64_t* children_checked) {
    if (!slos_units_supported(m, k)) return LWB_E_LIMIT;
    const uint64_t* bin = host_bin();
    Builder b;
    b.bin = bin; b.m = m; b.k = k; b.tmax = unit_tmax(binom(bin, m + k - 1, k));
    for (int d = 0; d <= SU_DMAX; ++d) b.nf_max[d] = 0;
    b.walk(0);
    if (!b.ok) return LWB_E_LIMIT;
    std::vector<uint32_t> ht((size_t)SU_DMAX * (m + 1));
    for (int r = 0; r < SU_DMAX; ++r)
        for (int v = 0; v <= m; ++v) ht[(size_t)r * (m + 1) + v] = (uint32_t)binom(bin, m - v + r - 1, r + 1);
    SuHostTables t4;
    const bool has4 = build_tables(bin, m, k, b, t4);
    uint64_t checked = 0, expect_start = 0;
    for (const SlosUnit& u : b.units) {
        if (u.start != expect_start) return LWB_E_ARG;
        if (has4) {  // the record kernel's rules, replayed with the kernel's own arithmetic
            int rc4 = check_unit_v4(bin, m, k, u, t4);
            if (rc4) return rc4;
        }
        expect_start += u.T;
        int x[LWB_MAX_PHOTONS + 2], depth = 0;
        for (int s = 0; s < u.np; ++s)
            for (int q = 0; q < u.pmult[s]; ++q) x[depth++] = u.pm[s];
        if (depth + u.d != k) return LWB_E_ARG;
        for (int i = 0; i < u.d; ++i) x[depth + i] = u.ell;  // first child: every trailing photon in mode l
        for (uint32_t c = 0; c < u.T; ++c) {
            if (lex_rank(bin, x, k, m) != u.start + c) return LWB_E_ARG;
            int y[LWB_MAX_PHOTONS + 2];
            for (int s = 0, at = 0; s < u.np; at += u.pmult[s], ++s) {  // remove one photon of prefix mode pm[s]
                int q = 0;
                for (int t = 0; t < k; ++t)
                    if (t != at) y[q++] = x[t];
                if (u.sbase[s] % 16 || lex_rank(bin, y, k - 1, m) != u.sbase[s] / 16 + c) return LWB_E_ARG;
            }
            uint32_t Q = 0;
            int prev = u.ell;
            for (int i = 0; i < u.d; ++i) {
                const int yi = x[depth + i], r = u.d - 1 - i;
                Q += ht[(size_t)r * (m + 1) + prev] - ht[(size_t)r * (m + 1) + yi];
                const bool last_of_run = (i == u.d - 1) || x[depth + i + 1] != yi;
                if (last_of_run && !(u.has_ell && yi == u.ell)) {
                    int q = 0;
                    for (int t = 0; t < k; ++t)
                        if (t != depth + i) y[q++] = x[t];
                    const uint32_t local = c - Q;
                    if (local >= u.Tp || lex_rank(bin, y, k - 1, m) != u.pblock + local) return LWB_E_ARG;
                }
                prev = yi;
            }
            ++checked;
            for (int i = u.d - 1; i >= 0; --i)  // lex successor of the trailing part
                if (x[depth + i] < m - 1) {
                    const int v = x[depth + i] + 1;
                    for (int q = i; q < u.d; ++q) x[depth + q] = v;
                    break;
                }
        }
    }
    if (expect_start != binom(bin, m + k - 1, k)) return LWB_E_ARG;
    *children_checked = checked;
    return 0;
}

// host-only description of a plan (tests): number of units / items, children covered, largest parent block
int slos_units_plan_info(int m, int k, uint64_t* n_units, uint64_t* n_items, uint64_t* children, uint64_t* max_tp) {
    if (!slos_units_supported(m, k)) return LWB_E_LIMIT;
    Builder b;
    b.bin = host_bin(); b.m = m; b.k = k; b.tmax = unit_tmax(binom(b.bin, m + k - 1, k));
    for (int d = 0; d <= SU_DMAX; ++d) b.nf_max[d] = 0;
    b.walk(0);
    if (!b.ok) return LWB_E_LIMIT;
    uint64_t tp = 0;
    for (auto& u : b.units) tp = std::max<uint64_t>(tp, u.Tp);
    uint64_t items = 0;
    const uint32_t item = item_size(b.next_start);
    for (size_t u = 0; u < b.units.size();) {
        size_t e = u;
        uint32_t tot = 0;
        while (e < b.units.size() && (e == u || tot + b.units[e].T <= item)) tot += b.units[e++].T;
        ++items;
        u = e;
    }
    *n_units = b.units.size();
    *n_items = items;
    *children = b.next_start;
    *max_tp = tp;
    return 0;
}

}  // namespace lwb
