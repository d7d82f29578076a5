// k1x.cu -- multiplicity-aware batched permanents (see k1x.h for the formula and the plan).
#include "k1x.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "../../include/lwb200.h"
#include "fock.cuh"
#include "glynn.cuh"
#include "k1x_common.cuh"

namespace lwb {

// ---- type key: parts >= 2 of the partition, descending, 5 bits each ------------------------------------
__host__ __device__ inline uint64_t pack_key(const int* parts_desc, int count) {
    uint64_t k = 0;
    for (int i = 0; i < count; ++i) k |= (uint64_t)parts_desc[i] << (5 * i);
    return k;
}

// multiplicity pattern of an ascending mode list -> key (parts >= 2 sorted descending)
template <int NMAX>
__device__ __forceinline__ uint64_t key_of_modes(const int* x, int n) {
    int parts[NMAX / 2 + 1];
    int np = 0, run = 1;
    for (int i = 1; i <= n; ++i) {
        if (i < n && x[i] == x[i - 1]) { ++run; continue; }
        if (run >= 2) {  // insertion into the descending list
            int p = np++;
            while (p > 0 && parts[p - 1] < run) { parts[p] = parts[p - 1]; --p; }
            parts[p] = run;
        }
        run = 1;
    }
    return pack_key(parts, np);
}

__device__ __forceinline__ int find_type(const uint64_t* keys, int n_types, uint64_t key) {
    int lo = 0, hi = n_types - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (keys[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

struct ClassifyArgs {
    const uint64_t* bin;
    int m, n;
    uint64_t r0, cnt;
    const uint64_t* keys;      // sorted
    const int* key_type;       // type id of keys[i]
    int n_types;
    uint32_t nb;               // number of classify blocks
    uint32_t* H;               // [n_types][nb]: pass 1 writes per-block counts; scanned in place to exclusive prefixes
    const uint32_t* offsets;   // pass 2: first bucket slot of every type
    uint32_t* bucket;          // pass 2 output
    unsigned char* types;      // [cnt] type of every state: written by pass 1, read by pass 2
    int fill;
};

// Stable bucketing of the rank range by multiplicity pattern, two passes over the same layout (thread <-> one
// rank).  Pass 1 (fill = 0): per-block counts H[type][block].  After the per-type exclusive scan of H, pass 2
// (fill = 1) writes every state's range-relative index to bucket[offsets[type] + H[type][block] + (rank of the
// state among the block's states of that type)], so every bucket lists its states in ascending rank order:
// neighbouring lanes of the permanent kernel then hold nearly identical states and their row reads coalesce
// into shared-memory broadcasts.
template <int NMAX>
__global__ void __launch_bounds__(K1X_CL_THREADS) k1x_classify_kernel(ClassifyArgs a) {
    __shared__ unsigned short whist[K1X_CL_THREADS / 32][K1X_MAX_TYPES];
    for (int e = threadIdx.x; e < (K1X_CL_THREADS / 32) * K1X_MAX_TYPES; e += K1X_CL_THREADS)
        (&whist[0][0])[e] = 0;
    __syncthreads();
    const uint64_t k = (uint64_t)blockIdx.x * K1X_CL_THREADS + threadIdx.x;
    const bool valid = k < a.cnt;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int t = K1X_MAX_TYPES + lane;  // invalid lanes never match anybody
    if (valid) {
        if (a.fill) {
            t = a.types[k];
        } else {
            int x[NMAX];
            colex_unrank(a.bin, a.r0 + k, a.n, a.m, x);
            t = a.key_type[find_type(a.keys, a.n_types, key_of_modes<NMAX>(x, a.n))];
            a.types[k] = (unsigned char)t;
        }
    }
    const unsigned peers = __match_any_sync(0xffffffffu, t);
    const int in_warp = __popc(peers & ((1u << lane) - 1u));
    if (valid && in_warp == 0) whist[w][t] = (unsigned short)__popc(peers);
    __syncthreads();
    if (!a.fill) {
        for (int ty = threadIdx.x; ty < a.n_types; ty += K1X_CL_THREADS) {
            unsigned int tot = 0;
            for (int ww = 0; ww < K1X_CL_THREADS / 32; ++ww) tot += whist[ww][ty];
            a.H[(size_t)ty * a.nb + blockIdx.x] = tot;
        }
        return;
    }
    if (valid) {
        unsigned int pos = a.offsets[t] + a.H[(size_t)t * a.nb + blockIdx.x] + in_warp;
        for (int ww = 0; ww < w; ++ww) pos += whist[ww][t];
        a.bucket[pos] = (uint32_t)k;
    }
}

// In-place exclusive scan of every row H[type][0..nb) and the row totals; one block per type.
__global__ void __launch_bounds__(1024) k1x_scan_kernel(uint32_t* H, uint32_t nb, unsigned int* totals) {
    __shared__ unsigned int part[1024];
    uint32_t* row = H + (size_t)blockIdx.x * nb;
    const uint32_t chunk = (nb + 1023) / 1024;
    const uint32_t b = threadIdx.x * chunk, e = min(nb, b + chunk);
    unsigned int sum = 0;
    for (uint32_t i = b; i < e; ++i) sum += row[i];
    part[threadIdx.x] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {  // Hillis-Steele inclusive scan of the 1024 chunk sums
        unsigned int v = threadIdx.x >= off ? part[threadIdx.x - off] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    unsigned int run = part[threadIdx.x] - sum;  // exclusive prefix of this chunk
    for (uint32_t i = b; i < e; ++i) {
        const unsigned int v = row[i];
        row[i] = run;
        run += v;
    }
    if (threadIdx.x == 1023) totals[blockIdx.x] = part[1023];
}

// Bucket layout from the per-type totals: exclusive scans, in layout order, of the state counts (-> offsets[type])
// and of the block counts ceil(count / threads per block) (-> bstart[position]); one block, thread <-> layout position.
__global__ void __launch_bounds__(K1X_MAX_TYPES) k1x_layout_kernel(const unsigned int* counts, const int* layout,
                                                                     int n_types, uint32_t threads, uint32_t* offsets,
                                                                     uint32_t* bstart) {
    __shared__ uint32_t so[K1X_MAX_TYPES], sb[K1X_MAX_TYPES];
    const int p = threadIdx.x;
    const int type = p < n_types ? layout[p] : 0;
    const uint32_t c = p < n_types ? counts[type] : 0u;
    const uint32_t nb = (c + threads - 1) / threads;
    so[p] = c;
    sb[p] = nb;
    __syncthreads();
    for (int off = 1; off < K1X_MAX_TYPES; off <<= 1) {
        const uint32_t vo = p >= off ? so[p - off] : 0u, vb = p >= off ? sb[p - off] : 0u;
        __syncthreads();
        so[p] += vo;
        sb[p] += vb;
        __syncthreads();
    }
    if (p < n_types) {
        offsets[type] = so[p] - c;
        bstart[p] = sb[p] - nb;
        if (p == n_types - 1) bstart[n_types] = sb[p];
    }
}

// ------------------------------------------------------------------------------------------------------------
struct Plan {
    int n = 0;
    int n_types = 0;
    std::vector<TypeMeta> meta;
    uint64_t* d_keys = nullptr;
    int* d_key_type = nullptr;
    TypeMeta* d_meta = nullptr;
    Step* d_prog = nullptr;
    int* d_layout = nullptr;  // type ids, heaviest first: the order of the buckets and of the blocks
};

struct K1xState {
    std::map<int, Plan> plans;
    unsigned int* d_counts = nullptr;
    uint32_t* d_H = nullptr;
    size_t H_cap = 0;
    unsigned char* d_types = nullptr;
    uint32_t* d_offsets = nullptr;
    uint32_t* d_bstart = nullptr;
    uint32_t* d_bucket = nullptr;
    size_t bucket_cap = 0;
};

K1xState* k1x_create() { return new K1xState(); }

void k1x_destroy(K1xState* st) {
    if (!st) return;
    for (auto& kv : st->plans) {
        cudaFree(kv.second.d_keys);
        cudaFree(kv.second.d_key_type);
        cudaFree(kv.second.d_meta);
        cudaFree(kv.second.d_prog);
        cudaFree(kv.second.d_layout);
    }
    cudaFree(st->d_counts);
    cudaFree(st->d_H);
    cudaFree(st->d_types);
    cudaFree(st->d_offsets);
    cudaFree(st->d_bstart);
    cudaFree(st->d_bucket);
    delete st;
}

bool k1x_supported(int n) { return n >= 2 && n <= K1X_MAX_N; }

static void partitions(int n, int maxpart, std::vector<int>& cur, std::vector<std::vector<int>>& out);
static void build_program(const std::vector<int>& s, TypeMeta& tm, std::vector<Step>& prog, bool allow_static);

// Host-only description of the walk programs for n photons (no device needed): number of multiplicity patterns,
// and for pattern index t (partitions of n in descending lexicographic order) its multiplicities, term count,
// symmetry factor and the sum of its coefficients' absolute values (= 2^n / factor... a checksum of the walk).
int k1x_plan_info(int n, int t, int* n_types, int* s_out, int* d_out, int* T_out, int* factor_out, double* abs_coef_sum) {
    if (!k1x_supported(n)) return LWB_E_LIMIT;
    std::vector<std::vector<int>> parts;
    std::vector<int> cur;
    partitions(n, n, cur, parts);
    *n_types = (int)parts.size();
    if (t < 0 || t >= (int)parts.size()) return LWB_E_ARG;
    TypeMeta tm;
    std::vector<Step> prog;
    build_program(parts[t], tm, prog, false);
    for (int i = 0; i < tm.d; ++i) s_out[i] = tm.s[i];
    *d_out = tm.d;
    *T_out = tm.T;
    *factor_out = tm.factor;
    double a = 0.0;
    for (const Step& st : prog) a += std::fabs(st.coef);
    *abs_coef_sum = a;
    return 0;
}

static void partitions(int n, int maxpart, std::vector<int>& cur, std::vector<std::vector<int>>& out) {
    if (n == 0) { out.push_back(cur); return; }
    for (int p = std::min(n, maxpart); p >= 1; --p) {
        cur.push_back(p);
        partitions(n - p, p, cur, out);
        cur.pop_back();
    }
}

static double binom_d(int a, int b) {
    double r = 1.0;
    for (int i = 1; i <= b; ++i) r = r * (double)(a - b + i) / (double)i;
    return std::floor(r + 0.5);
}

// reflected mixed-radix Gray walk (Knuth 7.2.1.1 Algorithm H) over the digits with radix > 1.
// Generic program (klass -1): every tuple of every row is a Step.  Static program (>= 3 single rows): only the
// bunched rows are digits; each Step stands for a whole binary Gray chunk over C = ns - 1 single rows, and its
// coefficient also carries the sign of the chunk's first term (the chunk parity alternates with the step index).
static void build_program(const std::vector<int>& s, TypeMeta& tm, std::vector<Step>& prog, bool allow_static = true) {
    const int d = (int)s.size();
    int ns = 0;
    for (int v : s) ns += (v == 1);
    const bool is_static = allow_static && ns >= 3;
    std::vector<int> R(d);
    for (int i = 0; i < d; ++i) R[i] = s[i] + 1;
    tm.factor = 1;
    tm.ns = ns;
    tm.klass = is_static ? ns - 1 : -1;
    if (is_static) {
        for (int i = 0; i < d; ++i)
            if (s[i] == 1) R[i] = 1;  // single rows are walked by the unrolled binary chunk
        tm.factor = 2;                // ... one of them fixed to '+'
    } else {
        int h = -1;
        for (int i = d - 1; i >= 0; --i)
            if (s[i] % 2 == 1) { h = i; break; }
        if (h >= 0) { R[h] = (s[h] + 1) / 2; tm.factor = 2; }
    }
    // fastest Gray digit = last slot: among equal multiplicities the slots are in ascending mode order and, in
    // the reference's rank order, high modes change slowest from one state to the next, so the rows flipped most
    // often are the ones neighbouring lanes share
    std::vector<int> digits;
    for (int i = d - 1; i >= 0; --i)
        if (R[i] > 1) digits.push_back(i);
    const int nd = (int)digits.size();
    std::vector<int> a(nd, 0), o(nd, 1), f(nd + 1);
    for (int j = 0; j <= nd; ++j) f[j] = j;
    std::vector<int> v(d, 0);
    tm.d = d;
    tm.prog_off = (int)prog.size();
    for (int i = 0; i < K1X_MAX_N; ++i) tm.s[i] = i < d ? s[i] : 0;
    int T = 0;
    for (;;) {
        double coef = 1.0;
        int sum = 0;
        for (int i = 0; i < d; ++i) { coef *= binom_d(s[i], v[i]); sum += v[i]; }
        if (is_static) sum += T & 1;  // chunk T starts from the pattern with (T & 1) single rows negated
        Step st;
        st.coef = (sum & 1) ? -coef : coef;
        st.slot = 0;
        st.dir = 0;
        int j = f[0];
        f[0] = 0;
        if (j == nd) { prog.push_back(st); ++T; break; }
        a[j] += o[j];
        st.slot = digits[j];
        st.dir = o[j];
        v[digits[j]] += o[j];
        if (a[j] == 0 || a[j] == R[digits[j]] - 1) { o[j] = -o[j]; f[j] = f[j + 1]; f[j + 1] = j + 1; }
        prog.push_back(st);
        ++T;
    }
    tm.T = T;
}

static int make_plan(K1xState* st, int n, Plan** out, std::string* err) {
    auto it = st->plans.find(n);
    if (it != st->plans.end()) { *out = &it->second; return 0; }
    Plan pl;
    pl.n = n;
    std::vector<std::vector<int>> parts;
    std::vector<int> cur;
    partitions(n, n, cur, parts);
    pl.n_types = (int)parts.size();
    if (pl.n_types > K1X_MAX_TYPES) { *err = "too many multiplicity patterns"; return LWB_E_LIMIT; }
    std::vector<Step> prog;
    std::vector<std::pair<uint64_t, int>> keyed;
    pl.meta.resize(pl.n_types);
    for (int t = 0; t < pl.n_types; ++t) {
        build_program(parts[t], pl.meta[t], prog, n >= K1X_STATIC_MIN_N);
        int big[K1X_MAX_N];
        int nb = 0;
        for (int p : parts[t])
            if (p >= 2) big[nb++] = p;
        keyed.push_back({pack_key(big, nb), t});
    }
    std::sort(keyed.begin(), keyed.end());
    std::vector<uint64_t> keys;
    std::vector<int> key_type;
    for (auto& kv : keyed) { keys.push_back(kv.first); key_type.push_back(kv.second); }
    std::vector<int> layout(pl.n_types);
    for (int t = 0; t < pl.n_types; ++t) layout[t] = t;
    auto terms = [&](int t) { return (double)pl.meta[t].T * (pl.meta[t].klass >= 0 ? (double)(1 << pl.meta[t].klass) : 1.0); };
#ifdef K1X_LAYOUT_BY_CLASS
    std::stable_sort(layout.begin(), layout.end(), [&](int x, int y) {
        return pl.meta[x].klass != pl.meta[y].klass ? pl.meta[x].klass > pl.meta[y].klass : terms(x) > terms(y);
    });
#else
    std::stable_sort(layout.begin(), layout.end(), [&](int x, int y) { return terms(x) > terms(y); });
#endif
#define K1X_CU(call)                                                                  \
    do {                                                                              \
        cudaError_t e_ = (call);                                                      \
        if (e_ != cudaSuccess) { *err = std::string(#call ": ") + cudaGetErrorString(e_); return LWB_E_CUDA; } \
    } while (0)
    K1X_CU(cudaMalloc(&pl.d_keys, keys.size() * 8));
    K1X_CU(cudaMalloc(&pl.d_key_type, key_type.size() * 4));
    K1X_CU(cudaMalloc(&pl.d_meta, pl.meta.size() * sizeof(TypeMeta)));
    K1X_CU(cudaMalloc(&pl.d_prog, prog.size() * sizeof(Step)));
    K1X_CU(cudaMalloc(&pl.d_layout, layout.size() * 4));
    K1X_CU(cudaMemcpy(pl.d_layout, layout.data(), layout.size() * 4, cudaMemcpyHostToDevice));
    K1X_CU(cudaMemcpy(pl.d_keys, keys.data(), keys.size() * 8, cudaMemcpyHostToDevice));
    K1X_CU(cudaMemcpy(pl.d_key_type, key_type.data(), key_type.size() * 4, cudaMemcpyHostToDevice));
    K1X_CU(cudaMemcpy(pl.d_meta, pl.meta.data(), pl.meta.size() * sizeof(TypeMeta), cudaMemcpyHostToDevice));
    K1X_CU(cudaMemcpy(pl.d_prog, prog.data(), prog.size() * sizeof(Step), cudaMemcpyHostToDevice));
    st->plans[n] = pl;
    *out = &st->plans[n];
    return 0;
}

template <int N>
struct K1xTable {  // n below the static range: the generic walk only
    static void fill(K1xFn (*t)[2]) {
        t[N][0] = k1x_fused_kernel<N, false, double>;
        t[N][1] = k1x_fused_kernel<N, false, float>;
        K1xTable<N - 1>::fill(t);
    }
};
template <>
struct K1xTable<1> {
    static void fill(K1xFn (*)[2]) {}
};

bool k1x_supported_dtype(int n, int dtype) { return k1x_supported(n) && (dtype == 0 || n <= K1X_C64_MAX_N); }

int k1x_basis_probs(K1xState* st, cudaStream_t stream, const uint64_t* d_bin, const double2* d_U, int m,
                    const uint8_t* d_in_state, int n, uint64_t r0, uint64_t cnt, double* d_probs, uint64_t* launches,
                    std::string* err, const MirrorSet* mirrors, int dtype) {
    static K1xFn table[K1X_MAX_N + 1][2];
    static std::once_flag table_once;  // contexts of several GPUs may make their first call concurrently
    std::call_once(table_once, [] {
        K1xTable<K1X_STATIC_MIN_N - 1>::fill(table);
        const K1xStaticEntry* (*groups[])(int*) = LWB_K1XS_GROUPS;
        for (auto g : groups) {
            int c = 0;
            const K1xStaticEntry* e = g(&c);
            for (int i = 0; i < c; ++i) { table[e[i].n][0] = e[i].fn[0]; table[e[i].n][1] = e[i].fn[1]; }
        }
    });
    if (dtype < 0 || dtype > 1 || !k1x_supported(n) || !table[n][dtype] || cnt >= (1ull << 32)) { *err = "k1x: unsupported size"; return LWB_E_LIMIT; }
    if (cnt == 0) return 0;
    Plan* pl = nullptr;
    int rc = make_plan(st, n, &pl, err);
    if (rc) return rc;
    if (!st->d_counts) {
        K1X_CU(cudaMalloc(&st->d_counts, K1X_MAX_TYPES * 4));
        K1X_CU(cudaMalloc(&st->d_offsets, K1X_MAX_TYPES * 4));
        K1X_CU(cudaMalloc(&st->d_bstart, (K1X_MAX_TYPES + 1) * 4));
    }
    if (st->bucket_cap < cnt) {
        if (st->d_bucket) K1X_CU(cudaFree(st->d_bucket));
        st->d_bucket = nullptr;
        K1X_CU(cudaMalloc(&st->d_bucket, cnt * 4));
        if (st->d_types) K1X_CU(cudaFree(st->d_types));
        st->d_types = nullptr;
        K1X_CU(cudaMalloc(&st->d_types, cnt));
        st->bucket_cap = cnt;
    }
    // pass 1: per-block counts; per-type exclusive scan over the blocks and totals; bucket/block layout
    const uint32_t nb = (uint32_t)((cnt + K1X_CL_THREADS - 1) / K1X_CL_THREADS);
    const size_t h_entries = (size_t)pl->n_types * nb;
    if (st->H_cap < h_entries) {
        if (st->d_H) K1X_CU(cudaFree(st->d_H));
        st->d_H = nullptr;
        K1X_CU(cudaMalloc(&st->d_H, h_entries * 4));
        st->H_cap = h_entries;
    }
    ClassifyArgs ca;
    ca.bin = d_bin; ca.m = m; ca.n = n; ca.r0 = r0; ca.cnt = cnt; ca.keys = pl->d_keys; ca.key_type = pl->d_key_type;
    ca.n_types = pl->n_types; ca.nb = nb; ca.H = st->d_H; ca.offsets = st->d_offsets; ca.bucket = st->d_bucket; ca.types = st->d_types; ca.fill = 0;
    k1x_classify_kernel<K1X_MAX_N><<<nb, K1X_CL_THREADS, 0, stream>>>(ca);
    k1x_scan_kernel<<<pl->n_types, 1024, 0, stream>>>(st->d_H, nb, st->d_counts);
    const int tb = k1x_threads(n);
    k1x_layout_kernel<<<1, K1X_MAX_TYPES, 0, stream>>>(st->d_counts, pl->d_layout, pl->n_types, (uint32_t)tb, st->d_offsets, st->d_bstart);
    // pass 2: stable fill of the buckets
    ca.fill = 1;
    k1x_classify_kernel<K1X_MAX_N><<<nb, K1X_CL_THREADS, 0, stream>>>(ca);
    // permanents: one block per tb permanents of one pattern; the grid is the upper bound
    // sum_t ceil(count_t / tb) <= ceil(cnt / tb) + n_types  (the block table lives on the device)
    K1xArgs ka;
    ka.bin = d_bin; ka.U = d_U; ka.m = m; ka.in_state = d_in_state; ka.r0 = r0; ka.meta = pl->d_meta; ka.prog = pl->d_prog;
    ka.layout = pl->d_layout; ka.bstart = st->d_bstart; ka.offsets = st->d_offsets; ka.counts = st->d_counts;
    ka.n_types = pl->n_types; ka.bucket = st->d_bucket; ka.cnt = (uint32_t)cnt; ka.probs = d_probs;
    ka.mir.n = 0;
    if (mirrors) ka.mir = *mirrors;
    const size_t smem = (size_t)2 * m * (n | 1) * (dtype ? 8 : 16) + (size_t)n * tb * 4;  // table + negated copy, slot offsets
    K1xFn fn = table[n][dtype];
    if (smem > 48 * 1024) K1X_CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint64_t grid = (cnt + tb - 1) / tb + (uint64_t)pl->n_types;
    fn<<<(unsigned)grid, tb, smem, stream>>>(ka);
    *launches += 5;
    K1X_CU(cudaGetLastError());
    return 0;
}

}  // namespace lwb
