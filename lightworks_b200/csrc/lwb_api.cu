// lwb_api.cu -- C ABI (include/lwb200.h) over the kernels.  Host logic only: argument checks,
// buffer staging, kernel-variant dispatch, launch geometry.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/lwb200.h"
#include "lwb_internal.h"
#include "dist_kernels.cuh"
#include "fock.cuh"
#include "k1x.h"
#include "misc_kernels.cuh"
#include "perm_kernels.cuh"
#include "sample_kernels.cuh"
#include "slos_kernels.cuh"
#include "variants.h"

using namespace lwb;

// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
namespace lwb {
int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
int set_error(int code, const char* msg) { return fail(code, "%s", msg); }  // for the host-only units (source_stats.cpp)
int g_source_string_keys = 0;  // lwb_set_option("source_string_keys", 1): force the string-key path of lwb_source_statistics
}  // namespace lwb

static uint64_t g_bin[LWB_BIN_A * LWB_BIN_B];
static std::once_flag g_bin_once;
const uint64_t* lwb::host_bin() {
    std::call_once(g_bin_once, [] { fill_binomials(g_bin); });
    return g_bin;
}

int lwb::ensure(lwb_handle h, Buf& b, size_t bytes) {
    if (bytes <= b.cap) return 0;
    if (b.p) CU(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    size_t want = std::max(bytes, (size_t)4096);
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) return fail(LWB_E_NOMEM, "cudaMalloc(%zu): %s", want, cudaGetErrorString(e));
    b.cap = want;
    (void)h;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Kernel variant tables live in the k1_g*.cu / k2_g*.cu instantiation units (variants.h).
static std::vector<K1Variant> g_k1;
static const K2Variant* g_k2[LWB_K2_MAX_N + 1];
static int g_tune_logl[LWB_K1_MAX_N + 1];
static int g_multiplicity_aware = 1;  // lwb_set_option("multiplicity_aware", 0/1)
static int g_slos_units = 4;          // lwb_set_option("slos_units", v): 0 gather kernel (K3), 1 unit kernel (K3u), 4 record kernel (K3r)
static std::once_flag g_tables_once;

static void load_tables() {
    std::call_once(g_tables_once, [] {
        for (int i = 0; i <= LWB_K1_MAX_N; ++i) g_tune_logl[i] = -1;
        const K1Variant* (*k1g[])(int*) = LWB_K1_GROUPS;
        for (auto fn : k1g) {
            int c = 0;
            const K1Variant* t = fn(&c);
            for (int i = 0; i < c; ++i) g_k1.push_back(t[i]);
        }
        for (auto& p : g_k2) p = nullptr;
        const K2Variant* (*k2g[])(int*) = LWB_K2_GROUPS;
        for (auto fn : k2g) {
            int c = 0;
            const K2Variant* t = fn(&c);
            for (int i = 0; i < c; ++i) g_k2[t[i].n] = &t[i];
        }
    });
}

static const K1Variant* pick_k1(int n) {
    load_tables();
    const K1Variant* def = nullptr;
    for (const auto& v : g_k1) {
        if (v.n != n) continue;
        if (!def) def = &v;
        if (g_tune_logl[n] >= 0 && v.log_l == g_tune_logl[n]) return &v;
    }
    return def;
}

// ---- launch helpers ----------------------------------------------------------------------------
int lwb::sum_doubles_dev(lwb_handle h, const double* d_x, uint64_t count, double* d_out) {
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((count + 4095) / 4096, (uint64_t)h->sm_count * 4));
    CK(ensure(h, h->bPart, (size_t)std::max(grid, 1024u) * 16));
    double* partials = (double*)h->bPart.p;
    sum_slices_kernel<<<grid, SUM_THREADS, 0, h->stream>>>(d_x, count, partials);
    sum_slices_kernel<<<1, SUM_THREADS, 0, h->stream>>>(partials, grid, d_out);
    h->launches += 2;
    CU(cudaGetLastError());
    return 0;
}

template <typename F>
static int resident_blocks(lwb_handle h, F fn, size_t smem, int* out) {
    if (smem > 48 * 1024) CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, K1_THREADS, smem));
    if (per_sm < 1) return fail(LWB_E_LIMIT, "kernel does not fit on an SM (smem %zu)", smem);
    *out = per_sm * h->sm_count;
    return 0;
}

// ------------------------------------------------------------------------------------------------
extern "C" {

const char* lwb_last_error(void) { return g_err.c_str(); }
int lwb_version(void) { return 100; }

int lwb_create(int device, lwb_handle* out) {
    if (!out) return fail(LWB_E_ARG, "lwb_create: null out");
    int count = 0;
    CU(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return fail(LWB_E_ARG, "lwb_create: device %d of %d", device, count);
    CU(cudaSetDevice(device));
    auto* h = new lwb_context();
    h->device = device;
    CU(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device));
    CU(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    h->stream = h->own_stream;
    CU(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (auto& e : h->chunk_done) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CU(cudaMalloc(&h->d_bin, sizeof g_bin));
    CU(cudaMemcpy(h->d_bin, host_bin(), sizeof g_bin, cudaMemcpyHostToDevice));
    load_tables();
    h->k1x = k1x_create();
    h->slos_units = slos_units_create();
    *out = h;
    return 0;
}

int lwb_destroy(lwb_handle h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (Buf* b : {&h->bU, &h->bIn, &h->bOut, &h->bRes, &h->bPart, &h->bTmp, &h->bLayerA, &h->bLayerB, &h->bProbs,
                   &h->bMarg, &h->bFirst, &h->bSmall, &h->bCdf, &h->bTile, &h->bUni, &h->bIdx, &h->bTerms})
        if (b->p) cudaFree(b->p);
    k1x_destroy(h->k1x);
    slos_units_destroy(h->slos_units);
    for (auto& d : h->dists)
        if (d.p) cudaFree(d.p);
    for (auto& kv : h->dist_pool) cudaFree(kv.second);
    if (h->d_bin) cudaFree(h->d_bin);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    for (auto& e : h->chunk_done) if (e) cudaEventDestroy(e);
    delete h;
    return 0;
}

int lwb_set_stream(lwb_handle h, void* s, int use_own) {
    if (!h) return fail(LWB_E_ARG, "null handle");
    if (!h->dist_pool.empty()) CU(cudaStreamSynchronize(h->stream));  // parked buffers are ordered on the OLD stream
    h->stream = use_own ? h->own_stream : (cudaStream_t)s;  // s == NULL is the legacy default stream
    return 0;
}

int lwb_synchronize(lwb_handle h) {
    if (!h) return fail(LWB_E_ARG, "null handle");
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

int lwb_device_info(lwb_handle h, int* sm_count, int* sm_clock_khz, char* name, int name_len) {
    if (!h) return fail(LWB_E_ARG, "null handle");
    cudaDeviceProp p;
    CU(cudaGetDeviceProperties(&p, h->device));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (sm_clock_khz) CU(cudaDeviceGetAttribute(sm_clock_khz, cudaDevAttrClockRate, h->device));
    if (name && name_len > 0) {
        strncpy(name, p.name, (size_t)name_len - 1);
        name[name_len - 1] = 0;
    }
    return 0;
}

int lwb_set_tuning(int n, int log_l) {
    load_tables();
    if (n < 1 || n > LWB_K1_MAX_N) return fail(LWB_E_ARG, "lwb_set_tuning: n=%d", n);
    if (log_l >= 0) {
        bool ok = false;
        for (const auto& v : g_k1) ok |= (v.n == n && v.log_l == log_l);
        if (!ok) return fail(LWB_E_ARG, "lwb_set_tuning: no variant n=%d log_l=%d", n, log_l);
    }
    g_tune_logl[n] = log_l;
    return 0;
}

int lwb_set_option(const char* name, int value) {
    if (!name) return fail(LWB_E_ARG, "null option name");
    if (!strcmp(name, "multiplicity_aware")) { g_multiplicity_aware = value ? 1 : 0; return 0; }
    if (!strcmp(name, "slos_units")) { g_slos_units = value < 0 ? 0 : (int)value; return 0; }
    if (!strcmp(name, "source_string_keys")) { lwb::g_source_string_keys = value ? 1 : 0; return 0; }
    return fail(LWB_E_ARG, "unknown option %s", name);
}

int lwb_multiplicity_plan(int n, int pattern, int* n_patterns, int* multiplicities, int* n_distinct, int* n_terms,
                          int* symmetry_factor, double* abs_coef_sum) {
    if (!n_patterns || !multiplicities || !n_distinct || !n_terms || !symmetry_factor || !abs_coef_sum)
        return fail(LWB_E_ARG, "null argument");
    int rc = k1x_plan_info(n, pattern, n_patterns, multiplicities, n_distinct, n_terms, symmetry_factor, abs_coef_sum);
    if (rc) return fail(rc, "no multiplicity-aware plan for n=%d pattern=%d", n, pattern);
    return 0;
}

int lwb_slos_plan_check(int m, int k, uint64_t* units, uint64_t* items, uint64_t* children) {
    if (!units || !items || !children) return fail(LWB_E_ARG, "null argument");
    uint64_t tp = 0, ch = 0;
    int rc = slos_units_plan_info(m, k, units, items, &ch, &tp);
    if (rc) return fail(rc, "no prefix-unit plan for m=%d k=%d", m, k);
    rc = slos_units_plan_check(m, k, children);
    if (rc) return fail(rc, "prefix-unit plan of m=%d k=%d failed its index self-check", m, k);
    return 0;
}

int lwb_launch_count(lwb_handle h, uint64_t* count) {
    if (!h || !count) return fail(LWB_E_ARG, "null");
    *count = h->launches;
    return 0;
}

// ---- Fock indexing (host) ----------------------------------------------------------------------
}  // extern "C"
int lwb::check_mn(int m, int n) {
    if (m < 1 || m > LWB_MAX_MODES) return fail(LWB_E_LIMIT, "modes m=%d outside [1,%d]", m, LWB_MAX_MODES);
    if (n < 0 || n > LWB_MAX_PHOTONS) return fail(LWB_E_LIMIT, "photons n=%d outside [0,%d]", n, LWB_MAX_PHOTONS);
    if (m + n - 1 >= LWB_BIN_A) return fail(LWB_E_LIMIT, "m+n too large");
    if (binom(host_bin(), m + n - 1, n) == UINT64_MAX) return fail(LWB_E_LIMIT, "Fock space exceeds 2^64");
    return 0;
}
extern "C" {

int lwb_fock_count(int m, int n, uint64_t* count) {
    CK(check_mn(m, n));
    *count = binom(host_bin(), m + n - 1, n);
    return 0;
}

static int photons_of(const uint8_t* s, int m) {
    int n = 0;
    for (int i = 0; i < m; ++i) n += s[i];
    return n;
}

int lwb_fock_rank(int m, const uint8_t* state, uint64_t* rank) {
    int n = photons_of(state, m);
    CK(check_mn(m, n));
    int x[LWB_MAX_PHOTONS + 1];
    state_to_modes(state, m, x);
    *rank = colex_rank(host_bin(), x, n);
    return 0;
}

int lwb_fock_unrank(int m, int n, uint64_t rank, uint8_t* state) {
    CK(check_mn(m, n));
    if (rank >= binom(host_bin(), m + n - 1, n)) return fail(LWB_E_ARG, "rank out of range");
    int x[LWB_MAX_PHOTONS + 1];
    colex_unrank(host_bin(), rank, n, m, x);
    modes_to_state(x, n, m, state);
    return 0;
}

int lwb_slos_rank(int m, const uint8_t* state, uint64_t* rank) {
    int n = photons_of(state, m);
    CK(check_mn(m, n));
    int x[LWB_MAX_PHOTONS + 1];
    state_to_modes(state, m, x);
    *rank = lex_rank(host_bin(), x, n, m);
    return 0;
}

int lwb_slos_unrank(int m, int n, uint64_t rank, uint8_t* state) {
    CK(check_mn(m, n));
    if (rank >= binom(host_bin(), m + n - 1, n)) return fail(LWB_E_ARG, "rank out of range");
    int x[LWB_MAX_PHOTONS + 1];
    lex_unrank(host_bin(), rank, n, m, x);
    modes_to_state(x, n, m, state);
    return 0;
}

int lwb_fock_basis_dev(lwb_handle h, int m, int n, uint64_t r0, uint64_t cnt, uint8_t* d_states) {
    if (!h) return fail(LWB_E_ARG, "null handle");
    CK(check_mn(m, n));
    uint64_t S = binom(host_bin(), m + n - 1, n);
    if (r0 > S || cnt > S - r0) return fail(LWB_E_ARG, "rank range [%llu,+%llu) outside %llu",
                                            (unsigned long long)r0, (unsigned long long)cnt, (unsigned long long)S);
    if (cnt == 0) return 0;
    CU(cudaSetDevice(h->device));
    const int threads = 128;
    const uint64_t per_thread = 32;
    uint64_t blocks = (cnt + threads * per_thread - 1) / (threads * per_thread);
    fock_basis_kernel<<<(unsigned)blocks, threads, 0, h->stream>>>(h->d_bin, m, n, r0, cnt, per_thread, d_states);
    h->launches++;
    CU(cudaGetLastError());
    return 0;
}

int lwb_fock_basis(lwb_handle h, int m, int n, uint64_t r0, uint64_t cnt, uint8_t* states) {
    if (!h) return fail(LWB_E_ARG, "null handle");
    std::lock_guard<std::mutex> lk(h->mu);
    CK(check_mn(m, n));
    CU(cudaSetDevice(h->device));
    CK(ensure(h, h->bOut, cnt * (uint64_t)m));
    CK(lwb_fock_basis_dev(h, m, n, r0, cnt, (uint8_t*)h->bOut.p));
    CU(cudaMemcpyAsync(states, h->bOut.p, cnt * (uint64_t)m, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

int lwb_perm_matrices_dev(lwb_handle h, const double* d_A, uint64_t B, int n, int dtype, double* d_out) {
    if (!h) return fail(LWB_E_ARG, "null handle");
    if (dtype != LWB_C128 && dtype != LWB_C64) return fail(LWB_E_ARG, "dtype %d", dtype);
    if (n < 1 || n > LWB_K1_MAX_N) return fail(LWB_E_LIMIT, "batched permanents support 1 <= n <= %d (got %d)", LWB_K1_MAX_N, n);
    if (B == 0) return 0;
    CU(cudaSetDevice(h->device));
    // every group of a block stages its own n x n matrix: of the compiled variants (default first) take the first
    // whose tile leaves room for two blocks per SM, unless a variant was requested explicitly
    const size_t esz = dtype == LWB_C128 ? 16 : 8;
    const size_t nt = k1_neg(n) ? 2 : 1;
    const K1Variant* v = pick_k1(n);
    if (g_tune_logl[n] < 0) {
        const K1Variant* best = nullptr;
        for (const auto& c : g_k1) {
            if (c.n != n) continue;
            const size_t need = (size_t)(K1_THREADS >> c.log_l) * nt * n * n * esz;  // matrix (+ negated copy) per group
            if (need <= (nt == 2 ? 160u : 100u) * 1024) { best = &c; break; }
            if (!best || need < (size_t)(K1_THREADS >> best->log_l) * nt * n * n * esz) best = &c;
        }
        v = best;
    }
    const int gpb = K1_THREADS >> v->log_l;
    const size_t smem = (size_t)gpb * nt * n * n * esz;
    if (smem > 220 * 1024) return fail(LWB_E_LIMIT, "matrix batch tile needs %zu B of shared memory", smem);
    auto fn = v->matrices[dtype];
    int resident = 0;
    CK(resident_blocks(h, fn, smem, &resident));
    uint64_t need = (B + gpb - 1) / gpb;
    unsigned grid = (unsigned)std::min<uint64_t>(need, (uint64_t)resident);
    PermMatricesArgs a{(const double2*)d_A, B, d_out};
    fn<<<grid, K1_THREADS, smem, h->stream>>>(a);
    h->launches++;
    CU(cudaGetLastError());
    return 0;
}

int lwb_perm_matrices(lwb_handle h, const double* A, uint64_t B, int n, int dtype, double* out) {
    if (!h || !A || !out) return fail(LWB_E_ARG, "null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    if (n < 1 || n > LWB_K1_MAX_N) return fail(LWB_E_LIMIT, "batched permanents support 1 <= n <= %d (got %d)", LWB_K1_MAX_N, n);
    if (B == 0) return 0;
    CU(cudaSetDevice(h->device));
    const size_t in_bytes = B * (uint64_t)n * n * 16, out_bytes = B * 16;
    CK(ensure(h, h->bU, in_bytes));
    CK(ensure(h, h->bRes, out_bytes));
    CU(cudaMemcpyAsync(h->bU.p, A, in_bytes, cudaMemcpyHostToDevice, h->stream));
    CK(lwb_perm_matrices_dev(h, (const double*)h->bU.p, B, n, dtype, (double*)h->bRes.p));
    CU(cudaMemcpyAsync(out, h->bRes.p, out_bytes, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

}  // extern "C"
int lwb::launch_states(lwb_handle h, const double* d_U, int m, const uint8_t* d_ins, int n_in,
                       const uint8_t* d_outs, uint64_t r0, uint64_t n_out, int n, int dtype,
                       int want_probs, double* d_out) {
    if (dtype != LWB_C128 && dtype != LWB_C64) return fail(LWB_E_ARG, "dtype %d", dtype);
    if (n < 1 || n > LWB_K1_MAX_N) return fail(LWB_E_LIMIT, "batched permanents support 1 <= n <= %d photons (got %d)", LWB_K1_MAX_N, n);
    if (m < 1 || m > LWB_MAX_MODES) return fail(LWB_E_LIMIT, "modes m=%d outside [1,%d]", m, LWB_MAX_MODES);
    if (n_in < 1 || n_in > 65535) return fail(LWB_E_ARG, "n_in=%d outside [1,65535]", n_in);
    if (n_out == 0) return 0;
    if (g_multiplicity_aware && !d_outs && n_in == 1 && want_probs && k1x_supported_dtype(n, dtype) &&
        n >= 7 && n_out < (1ull << 32)) {  // by n and dtype only: a rank-range shard is bit-identical to the same slice of the whole
        // full-basis rank range: bunched output states repeat rows, walk prod(s_i+1)/2 terms instead of 2^(n-1)
        std::string err;
        int rc = k1x_basis_probs(h->k1x, h->stream, h->d_bin, (const double2*)d_U, m, d_ins, n, r0, n_out, d_out,
                                 &h->launches, &err, &h->mirror, dtype);
        if (rc) return fail(rc, "%s", err.c_str());
        return 0;
    }
    const K1Variant* v = pick_k1(n);
    const int gpb = K1_THREADS >> v->log_l;
    const size_t smem = (size_t)(k1_neg(n) ? 2 : 1) * m * n * (dtype == LWB_C128 ? 16 : 8) + (size_t)gpb * n * sizeof(int);  // rows (+ negated rows)
    auto fn = v->states[dtype];
    int resident = 0;
    CK(resident_blocks(h, fn, smem, &resident));
    uint64_t need = (n_out + gpb - 1) / gpb;
    uint64_t gx = std::max<uint64_t>(1, (uint64_t)resident / (uint64_t)n_in);
    gx = std::min<uint64_t>(need, gx);
    PermStatesArgs a{(const double2*)d_U, m, d_ins, d_outs, r0, n_out, h->d_bin, d_out, want_probs, h->mirror, 0, 0, 0};
    if (!want_probs) a.mir.n = 0;
    unsigned gz = 1;
    if (h->batch_tasks > 1) {  // lwb_perm_amplitudes_batch: one launch for every unitary of the batch
        gz = (unsigned)h->batch_tasks;
        a.u_stride = (uint64_t)m * m;
        a.in_stride = h->batch_shared_inputs ? 0 : (uint64_t)n_in * m;
        a.out_stride = (uint64_t)n_in * n_out * (want_probs ? 1 : 2);
        a.mir.n = 0;
        gx = std::max<uint64_t>(1, gx / gz);
    }
    dim3 grid((unsigned)gx, (unsigned)n_in, gz);
    fn<<<grid, K1_THREADS, smem, h->stream>>>(a);
    h->launches++;
    CU(cudaGetLastError());
    return 0;
}
extern "C" {

int lwb_perm_amplitudes_dev(lwb_handle h, const double* d_U, int m, const uint8_t* d_ins, int n_in,
                            const uint8_t* d_outs, uint64_t n_out, int n_photons, int dtype,
                            int want_probs, double* d_out) {
    if (!h || !d_outs) return fail(LWB_E_ARG, "null argument");
    CU(cudaSetDevice(h->device));
    return launch_states(h, d_U, m, d_ins, n_in, d_outs, 0, n_out, n_photons, dtype, want_probs, d_out);
}

int lwb_perm_amplitudes(lwb_handle h, const double* U, int m, const uint8_t* ins, int n_in,
                        const uint8_t* outs, uint64_t n_out, int dtype, int want_probs, double* out) {
    if (!h || !U || !ins || !outs || !out) return fail(LWB_E_ARG, "null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    if (m < 1 || m > LWB_MAX_MODES) return fail(LWB_E_LIMIT, "modes m=%d outside [1,%d]", m, LWB_MAX_MODES);
    if (n_in < 1) return fail(LWB_E_ARG, "n_in=%d", n_in);
    if (n_out == 0) return 0;
    const int n = photons_of(ins, m);
    for (int i = 0; i < n_in; ++i)
        if (photons_of(ins + (size_t)i * m, m) != n) return fail(LWB_E_PHOTONS, "input %d holds a different photon number", i);
    for (uint64_t j = 0; j < n_out; ++j)
        if (photons_of(outs + j * m, m) != n)
            return fail(LWB_E_PHOTONS, "output %llu holds %d photons, inputs hold %d", (unsigned long long)j,
                        photons_of(outs + j * m, m), n);
    const size_t per = want_probs ? 8 : 16;
    if (n == 0) {  // perm of the 0x0 matrix = 1 (thewalrus), factorials = 1
        for (uint64_t k = 0; k < (uint64_t)n_in * n_out; ++k) {
            if (want_probs) out[k] = 1.0;
            else { out[2 * k] = 1.0; out[2 * k + 1] = 0.0; }
        }
        return 0;
    }
    CU(cudaSetDevice(h->device));
    CK(ensure(h, h->bU, (size_t)m * m * 16));
    CK(ensure(h, h->bIn, (size_t)n_in * m));
    CK(ensure(h, h->bOut, n_out * (uint64_t)m));
    CK(ensure(h, h->bRes, (size_t)n_in * n_out * per));
    CU(cudaMemcpyAsync(h->bU.p, U, (size_t)m * m * 16, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->bIn.p, ins, (size_t)n_in * m, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->bOut.p, outs, n_out * (uint64_t)m, cudaMemcpyHostToDevice, h->stream));
    CK(launch_states(h, (const double*)h->bU.p, m, (const uint8_t*)h->bIn.p, n_in, (const uint8_t*)h->bOut.p, 0,
                     n_out, n, dtype, want_probs, (double*)h->bRes.p));
    CU(cudaMemcpyAsync(out, h->bRes.p, (size_t)n_in * n_out * per, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

// Batch across unitaries (sdk/tasks/batch.py:45-58, run serially by emulator/backends/backend.py:102-103): B tasks that
// differ in U (and optionally in their inputs) but share photon number and output states, ONE kernel launch.
int lwb_perm_amplitudes_batch(lwb_handle h, const double* U, int n_tasks, int m, const uint8_t* ins, int n_in,
                              int shared_inputs, const uint8_t* outs, uint64_t n_out, int dtype, int want_probs, double* out) {
    if (!h || !U || !ins || !outs || !out) return fail(LWB_E_ARG, "null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    if (m < 1 || m > LWB_MAX_MODES) return fail(LWB_E_LIMIT, "modes m=%d outside [1,%d]", m, LWB_MAX_MODES);
    if (n_tasks < 1 || n_tasks > 65535) return fail(LWB_E_ARG, "n_tasks=%d outside [1,65535]", n_tasks);
    if (n_in < 1) return fail(LWB_E_ARG, "n_in=%d", n_in);
    if (n_out == 0) return 0;
    const int n = photons_of(ins, m);
    const size_t n_in_rows = (size_t)(shared_inputs ? 1 : n_tasks) * n_in;
    for (size_t i = 0; i < n_in_rows; ++i)
        if (photons_of(ins + i * m, m) != n) return fail(LWB_E_PHOTONS, "input %zu holds a different photon number", i);
    for (uint64_t j = 0; j < n_out; ++j)
        if (photons_of(outs + j * m, m) != n)
            return fail(LWB_E_PHOTONS, "output %llu holds %d photons, inputs hold %d", (unsigned long long)j,
                        photons_of(outs + j * m, m), n);
    const size_t per = want_probs ? 8 : 16;
    const size_t total = (size_t)n_tasks * n_in * n_out;
    if (n == 0) {
        for (size_t k = 0; k < total; ++k) {
            if (want_probs) out[k] = 1.0;
            else { out[2 * k] = 1.0; out[2 * k + 1] = 0.0; }
        }
        return 0;
    }
    CU(cudaSetDevice(h->device));
    CK(ensure(h, h->bU, (size_t)n_tasks * m * m * 16));
    CK(ensure(h, h->bIn, n_in_rows * m));
    CK(ensure(h, h->bOut, n_out * (uint64_t)m));
    CK(ensure(h, h->bRes, total * per));
    CU(cudaMemcpyAsync(h->bU.p, U, (size_t)n_tasks * m * m * 16, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->bIn.p, ins, n_in_rows * m, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->bOut.p, outs, n_out * (uint64_t)m, cudaMemcpyHostToDevice, h->stream));
    h->batch_tasks = n_tasks;
    h->batch_shared_inputs = shared_inputs ? 1 : 0;
    const int rc = launch_states(h, (const double*)h->bU.p, m, (const uint8_t*)h->bIn.p, n_in, (const uint8_t*)h->bOut.p, 0,
                                 n_out, n, dtype, want_probs, (double*)h->bRes.p);
    h->batch_tasks = 0;
    if (rc) return rc;
    CU(cudaMemcpyAsync(out, h->bRes.p, total * per, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

int lwb_perm_basis_probs_dev(lwb_handle h, const double* d_U, int m, const uint8_t* d_in_state,
                             int n_photons, uint64_t r0, uint64_t cnt, int dtype, double* d_probs) {
    if (!h) return fail(LWB_E_ARG, "null handle");
    CK(check_mn(m, n_photons));
    uint64_t S = binom(host_bin(), m + n_photons - 1, n_photons);
    if (r0 > S || cnt > S - r0) return fail(LWB_E_ARG, "rank range outside the Fock space");
    CU(cudaSetDevice(h->device));
    return launch_states(h, d_U, m, d_in_state, 1, nullptr, r0, cnt, n_photons, dtype, 1, d_probs);
}

int lwb_perm_basis_probs(lwb_handle h, const double* U, int m, const uint8_t* in_state, uint64_t r0,
                         uint64_t cnt, int dtype, double* probs) {
    if (!h || !U || !in_state || !probs) return fail(LWB_E_ARG, "null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    if (m < 1 || m > LWB_MAX_MODES) return fail(LWB_E_LIMIT, "modes m=%d outside [1,%d]", m, LWB_MAX_MODES);
    const int n = photons_of(in_state, m);
    CK(check_mn(m, n));
    uint64_t S = binom(host_bin(), m + n - 1, n);
    if (r0 > S || cnt > S - r0) return fail(LWB_E_ARG, "rank range outside the Fock space");
    if (cnt == 0) return 0;
    if (n == 0) { probs[0] = 1.0; return 0; }
    CU(cudaSetDevice(h->device));
    CK(ensure(h, h->bU, (size_t)m * m * 16));
    CK(ensure(h, h->bIn, (size_t)m));
    CK(ensure(h, h->bRes, cnt * 8));
    CU(cudaMemcpyAsync(h->bU.p, U, (size_t)m * m * 16, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->bIn.p, in_state, (size_t)m, cudaMemcpyHostToDevice, h->stream));
    // Large ranges are cut into a few pieces of decreasing size so the device-to-host copy of piece c overlaps the
    // kernels of piece c+1 and only the small last copy is exposed (the copy is the only host<->device traffic of
    // this path: 8 bytes per state).  Few pieces: every piece pays the bucketing passes of the multiplicity-aware walk.
    static const double cut[5] = {0.0, 0.40, 0.70, 0.90, 1.0};
    // ... but only into PINNED host memory: a device-to-host copy into pageable memory returns when it is done, so the
    // pieces would run one after the other and each would pay the bucketing passes again
    cudaPointerAttributes pa;
    const bool pinned = cudaPointerGetAttributes(&pa, probs) == cudaSuccess && pa.type == cudaMemoryTypeHost;
    cudaGetLastError();  // unregistered host memory reports an error on old drivers
    const int chunks = (pinned && cnt >= (4u << 20)) ? 4 : 1;
    for (int c = 0; c < chunks; ++c) {
        const uint64_t b = chunks == 1 ? 0 : (uint64_t)((double)cnt * cut[c]);
        const uint64_t e = chunks == 1 ? cnt : (c + 1 == chunks ? cnt : (uint64_t)((double)cnt * cut[c + 1]));
        CK(launch_states(h, (const double*)h->bU.p, m, (const uint8_t*)h->bIn.p, 1, nullptr, r0 + b, e - b, n, dtype, 1,
                         (double*)h->bRes.p + b));
        if (chunks == 1) {
            CU(cudaMemcpyAsync(probs, h->bRes.p, cnt * 8, cudaMemcpyDeviceToHost, h->stream));
        } else {
            CU(cudaEventRecord(h->chunk_done[c], h->stream));
            CU(cudaStreamWaitEvent(h->copy_stream, h->chunk_done[c], 0));
            CU(cudaMemcpyAsync(probs + b, (double*)h->bRes.p + b, (e - b) * 8, cudaMemcpyDeviceToHost, h->copy_stream));
        }
    }
    CU(cudaStreamSynchronize(h->stream));
    if (chunks > 1) CU(cudaStreamSynchronize(h->copy_stream));
    return 0;
}

int lwb_perm_single_dev(lwb_handle h, const double* d_A, int n, int dtype, int slice, int n_slices, double* d_out2) {
    if (!h) return fail(LWB_E_ARG, "null handle");
    if (dtype != LWB_C128 && dtype != LWB_C64) return fail(LWB_E_ARG, "dtype %d", dtype);
    if (n < 1 || n > LWB_K2_MAX_N) return fail(LWB_E_LIMIT, "single permanent supports 1 <= n <= %d (got %d)", LWB_K2_MAX_N, n);
    if (n_slices < 1 || slice < 0 || slice >= n_slices) return fail(LWB_E_ARG, "slice %d of %d", slice, n_slices);
    CU(cudaSetDevice(h->device));
    load_tables();
    if (!g_k2[n]) return fail(LWB_E_LIMIT, "no single-permanent kernel for n=%d", n);
    const int C = g_k2[n]->chunk_bits;
    const uint64_t Q = (n - 1 - C) > 0 ? 1ull << (n - 1 - C) : 1ull;  // chunks of 2^C Gray steps
    const uint64_t qb = Q * (uint64_t)slice / (uint64_t)n_slices, qe = Q * (uint64_t)(slice + 1) / (uint64_t)n_slices;
    auto fn = g_k2[n]->single[dtype];
    int resident = 0;
    CK(resident_blocks(h, fn, 0, &resident));
    // every lane walks a contiguous range of chunks; aim for >= 4 chunks per lane before adding blocks
    uint64_t need = (qe - qb + 4 * K1_THREADS - 1) / (4 * K1_THREADS);
    unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(need, (uint64_t)resident));
    // partials [grid][2] followed by the finished-blocks counter (zeroed once: the last block of every launch resets it)
    const size_t need_bytes = (size_t)std::max(grid, 1024u) * 16 + 16;
    if (h->bPart.cap < need_bytes) {
        CK(ensure(h, h->bPart, need_bytes));
        CU(cudaMemsetAsync(h->bPart.p, 0, h->bPart.cap, h->stream));
    }
    unsigned int* done = (unsigned int*)((char*)h->bPart.p + h->bPart.cap - 16);
    if (qe > qb) {
        PermSingleArgs a{(const double2*)d_A, qb, qe, (double*)h->bPart.p, done, d_out2};
        fn<<<grid, K1_THREADS, 0, h->stream>>>(a);
    } else {
        reduce_partials_kernel<<<1, 256, 0, h->stream>>>((const double*)h->bPart.p, 0, d_out2);  // empty slice: 0
    }
    h->launches++;
    CU(cudaGetLastError());
    return 0;
}

int lwb_perm_single(lwb_handle h, const double* A, int n, int dtype, int slice, int n_slices, double* out2) {
    if (!h || !A || !out2) return fail(LWB_E_ARG, "null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    if (n < 1 || n > LWB_K2_MAX_N) return fail(LWB_E_LIMIT, "single permanent supports 1 <= n <= %d (got %d)", LWB_K2_MAX_N, n);
    CU(cudaSetDevice(h->device));
    CK(ensure(h, h->bU, (size_t)n * n * 16));
    CK(ensure(h, h->bRes, 16));
    CU(cudaMemcpyAsync(h->bU.p, A, (size_t)n * n * 16, cudaMemcpyHostToDevice, h->stream));
    CK(lwb_perm_single_dev(h, (const double*)h->bU.p, n, dtype, slice, n_slices, (double*)h->bRes.p));
    CU(cudaMemcpyAsync(out2, h->bRes.p, 16, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

int lwb_fp64_probe(lwb_handle h, int mode, double* flops_per_s) {
    if (!h || !flops_per_s) return fail(LWB_E_ARG, "null argument");
    if (mode < 1 || mode > 5) return fail(LWB_E_ARG, "mode %d", mode);
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    CK(ensure(h, h->bTmp, 8 * 1024));
    double host[24];
    for (int i = 0; i < 24; ++i) host[i] = (i < 8) ? 1.0 + 1e-7 * i : (i < 16 ? 1e-9 * (i - 7) : 0.5 + 0.01 * i);
    if (mode == 3) for (int i = 0; i < 8; ++i) host[i] = 1.0 + 1e-12 * i;
    if (mode == 5) for (int i = 0; i < 8; ++i) { host[i] = 0.9999999; host[8 + i] = 1e-4; }
    CU(cudaMemcpyAsync((char*)h->bTmp.p + 4096, host, sizeof host, cudaMemcpyHostToDevice, h->stream));
    const int iters = 2048, blocks = h->sm_count * 4, threads = 256;
    const double per_iter = (mode == 5) ? 32.0 : 8.0;  // FP64 instructions per thread per iteration
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        CU(cudaEventRecord(e0, h->stream));
        double* o = (double*)h->bTmp.p;
        const double* src = (const double*)((char*)h->bTmp.p + 4096);
        switch (mode) {
            case 1: fp64_probe_kernel<1><<<blocks, threads, 0, h->stream>>>(o, iters, src); break;
            case 2: fp64_probe_kernel<2><<<blocks, threads, 0, h->stream>>>(o, iters, src); break;
            case 3: fp64_probe_kernel<3><<<blocks, threads, 0, h->stream>>>(o, iters, src); break;
            case 4: fp64_probe_kernel<4><<<blocks, threads, 0, h->stream>>>(o, iters, src); break;
            default: fp64_probe_kernel<5><<<blocks, threads, 0, h->stream>>>(o, iters, src); break;
        }
        h->launches++;
        CU(cudaEventRecord(e1, h->stream));
        CU(cudaEventSynchronize(e1));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        double flops = 2.0 * per_iter * (double)iters * blocks * threads / (ms * 1e-3);
        if (rep > 0) best = std::max(best, flops);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *flops_per_s = best;
    return 0;
}

int lwb_fp64_peak(lwb_handle h, double* flops_per_s) {
    if (!h || !flops_per_s) return fail(LWB_E_ARG, "null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    CK(ensure(h, h->bTmp, 8 * 1024));
    const int iters = 4096, blocks = h->sm_count * 8, threads = 256;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CU(cudaEventRecord(e0, h->stream));
        dfma_peak_kernel<<<blocks, threads, 0, h->stream>>>((double*)h->bTmp.p, iters, 1.0000001);
        h->launches++;
        CU(cudaEventRecord(e1, h->stream));
        CU(cudaEventSynchronize(e1));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        double flops = 2.0 * DFMA_PER_ITER * (double)iters * blocks * threads / (ms * 1e-3);
        if (rep > 0) best = std::max(best, flops);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *flops_per_s = best;
    return 0;
}

}  // extern "C"

// ---- SLOS ----------------------------------------------------------------------------------------
template <typename Args, typename K4, typename K8, typename K12, typename K16, typename K24, typename K40>
static int launch_kmax(lwb_handle h, int n, uint64_t threads_needed, size_t smem, const Args& a, K4 k4, K8 k8, K12 k12,
                       K16 k16, K24 k24, K40 k40) {
    const unsigned grid = (unsigned)((threads_needed + SLOS_THREADS - 1) / SLOS_THREADS);
    if (grid == 0) return 0;
#define LWB_GO(fn)                                                                                   \
    do {                                                                                             \
        if (smem > 48 * 1024) CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        fn<<<grid, SLOS_THREADS, smem, h->stream>>>(a);                                              \
    } while (0)
    if (n <= 4) LWB_GO(k4);
    else if (n <= 8) LWB_GO(k8);
    else if (n <= 12) LWB_GO(k12);
    else if (n <= 16) LWB_GO(k16);
    else if (n <= 24) LWB_GO(k24);
    else LWB_GO(k40);
#undef LWB_GO
    h->launches++;
    CU(cudaGetLastError());
    return 0;
}

typedef void (*SlosFn)(SlosLayerArgs);
template <int K>
struct SlosTable {
    static void fill(SlosFn (*t)[2]) {
        t[K][0] = slos_layer_kernel<K, uint32_t>;
        t[K][1] = slos_layer_kernel<K, uint64_t>;
        SlosTable<K - 1>::fill(t);
    }
};
template <>
struct SlosTable<0> {
    static void fill(SlosFn (*)[2]) {}
};
static SlosFn g_slos[LWB_MAX_PHOTONS + 1][2];
static std::once_flag g_slos_once;

static int launch_slos_layer(lwb_handle h, const SlosLayerArgs& a) {
    if (g_slos_units && slos_units_supported(a.m, a.k)) {
        std::string err;
        const int rc = slos_units_layer(h->slos_units, h->stream, a.m, a.k, a.parent, a.child, a.U, a.in_mode, a.sqrt_tab,
                                        a.write_probs, &h->launches, &err, g_slos_units >= 4 ? 4 : 3);
        if (rc == 0) return 0;
        if (rc != LWB_E_LIMIT) return fail(rc, "%s", err.c_str());  // LWB_E_LIMIT: shape outside K3u, use the gather kernel
    }
    std::call_once(g_slos_once, [] { SlosTable<LWB_MAX_PHOTONS>::fill(g_slos); });
    const int k = a.k, m = a.m;
    const int wide = a.S + 32ull * SLOS_ITERS >= (1ull << 32) ? 1 : 0;  // a lane overshoots S by up to 32 * SLOS_ITERS
    const size_t isz = wide ? 8 : 4;
    const size_t smem = (size_t)m * 16 + (size_t)((k + 2) & ~1) * 8 + 3 * (size_t)k * (m + 1) * isz;
    const uint64_t warps = (a.S + 32ull * SLOS_ITERS - 1) / (32ull * SLOS_ITERS);
    const unsigned grid = (unsigned)((warps + SLOS_THREADS / 32 - 1) / (SLOS_THREADS / 32));
    SlosFn fn = g_slos[k][wide];
    if (smem > 48 * 1024) CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fn<<<grid, SLOS_THREADS, smem, h->stream>>>(a);
    h->launches++;
    CU(cudaGetLastError());
    return 0;
}

// Runs every SLOS layer (all layers in SLOS / lexicographic order).  The last layer is written to `final_dst`
// (S(m,n) complex amplitudes, or S(m,n) doubles |a|^2 when final_probs).
static int slos_layers(lwb_handle h, const double* d_U, int m, const uint8_t* in_state, int n, void* final_dst,
                       int final_probs) {
    const uint64_t* hb = host_bin();
    const uint64_t Sp1 = n >= 1 ? binom(hb, m + n - 2, n - 1) : 1;  // layer n-1
    const uint64_t Sp2 = n >= 2 ? binom(hb, m + n - 3, n - 2) : 1;  // layer n-2
    CK(ensure(h, h->bLayerA, std::max<uint64_t>(Sp1, 1) * 16));
    CK(ensure(h, h->bLayerB, std::max<uint64_t>(Sp2, 1) * 16));
    CK(ensure(h, h->bSmall, 4096));
    double vf = 1.0;  // vector_factorial, slos.py:126-128
    int modes[LWB_MAX_PHOTONS + 1];
    int c = 0;
    for (int i = 0; i < m; ++i)
        for (int k = 0; k < in_state[i]; ++k) { modes[c++] = i; vf *= (double)(k + 1); }
    double host_small[64 + 2];
    host_small[0] = 1.0 / sqrt(vf);  // 1 / np.sqrt(vector_factorial), slos.py:92
    host_small[1] = 0.0;
    for (int k = 0; k <= n; ++k) host_small[2 + k] = pow((double)k, 0.5);  // key[mode] ** 0.5, slos.py:121
    CU(cudaMemcpyAsync(h->bSmall.p, host_small, sizeof(double) * (size_t)(n + 3), cudaMemcpyHostToDevice, h->stream));
    // ping-pong so that layer n-1 lands in A (the larger buffer): layer j lives in A iff (n-1-j) is even
    double2* bufs[2] = {(double2*)h->bLayerA.p, (double2*)h->bLayerB.p};
    auto buf_of = [&](int j) { return bufs[((n - 1 - j) % 2 + 2) % 2]; };
    CU(cudaMemcpyAsync(n == 0 ? final_dst : (void*)buf_of(0), h->bSmall.p, 16, cudaMemcpyDeviceToDevice, h->stream));
    for (int k = 1; k <= n; ++k) {
        SlosLayerArgs a;
        a.bin = h->d_bin; a.m = m; a.k = k; a.S = binom(hb, m + k - 1, k);
        a.parent = buf_of(k - 1);
        a.child = (k == n) ? final_dst : (void*)buf_of(k);
        a.U = (const double2*)d_U; a.in_mode = modes[k - 1];
        a.sqrt_tab = (const double*)h->bSmall.p + 2;
        a.write_probs = (k == n) ? final_probs : 0;
        CK(launch_slos_layer(h, a));
    }
    return 0;
}

static int slos_check(int m, const uint8_t* in_state, int* n_out) {
    if (m < 1 || m > LWB_MAX_MODES) return fail(LWB_E_LIMIT, "modes m=%d outside [1,%d]", m, LWB_MAX_MODES);
    int n = 0;
    for (int i = 0; i < m; ++i) n += in_state[i];
    CK(check_mn(m, n));
    *n_out = n;
    return 0;
}

// out_mode: 0 complex amplitudes, 1 probabilities; order: 1 = SLOS (native), 0 = fock_basis (reordered)
static int slos_run_dev(lwb_handle h, const double* d_U, int m, const uint8_t* in_state, int order, int out_mode,
                        double* d_out) {
    int n = 0;
    CK(slos_check(m, in_state, &n));
    if (order != 0 && order != 1) return fail(LWB_E_ARG, "order %d", order);
    CU(cudaSetDevice(h->device));
    const uint64_t S = binom(host_bin(), m + n - 1, n);
    if (order == 1 || n == 0) {
        if (n == 0 && out_mode == 1) {
            const double one = 1.0;
            CU(cudaMemcpyAsync(d_out, &one, 8, cudaMemcpyHostToDevice, h->stream));
            CU(cudaStreamSynchronize(h->stream));
            return 0;
        }
        return slos_layers(h, d_U, m, in_state, n, d_out, out_mode);
    }
    CK(ensure(h, h->bTmp, S * 16));
    CK(slos_layers(h, d_U, m, in_state, n, h->bTmp.p, 0));
    ReorderArgs ra{h->d_bin, m, n, S, (const double*)h->bTmp.p, d_out, out_mode};
    CK(launch_kmax(h, n, S, 0, ra, reorder_to_colex_kernel<4>, reorder_to_colex_kernel<8>, reorder_to_colex_kernel<12>,
                   reorder_to_colex_kernel<16>, reorder_to_colex_kernel<24>, reorder_to_colex_kernel<40>));
    return 0;
}

static int slos_run_host(lwb_handle h, const double* U, int m, const uint8_t* in_state, int order, int out_mode,
                         double* out) {
    if (!h || !U || !in_state || !out) return fail(LWB_E_ARG, "null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    int n = 0;
    CK(slos_check(m, in_state, &n));
    CU(cudaSetDevice(h->device));
    const uint64_t S = binom(host_bin(), m + n - 1, n);
    const size_t per = out_mode == 0 ? 16 : 8;
    CK(ensure(h, h->bU, (size_t)m * m * 16));
    CK(ensure(h, h->bRes, S * per));
    CU(cudaMemcpyAsync(h->bU.p, U, (size_t)m * m * 16, cudaMemcpyHostToDevice, h->stream));
    CK(slos_run_dev(h, (const double*)h->bU.p, m, in_state, order, out_mode, (double*)h->bRes.p));
    CU(cudaMemcpyAsync(out, h->bRes.p, S * per, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

// Steps 2a / 2b of full_probability_distribution from per-full-state probabilities already on the device
// (fock_basis order for the permanent backend, and for SLOS with loss modes; SLOS order for lossless SLOS).
int lwb::pdist_finish(lwb_handle h, const double* d_probs, int m_tot, int n_modes, int n, double threshold, int backend,
                      uint8_t* keys, double* vals, uint64_t cap, uint64_t* count_out) {
    const uint64_t* hb = host_bin();
    const int loss = m_tot - n_modes;
    const int order = backend == LWB_BACKEND_SLOS ? 1 : 0;
    const uint64_t S = binom(hb, m_tot + n - 1, n);
    if (loss == 0) {
        // 2a. no loss modes: every full state is its own key; filter p > threshold in enumeration order
        std::vector<double> p(S);
        CU(cudaMemcpyAsync(p.data(), d_probs, S * 8, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        int x[LWB_MAX_PHOTONS + 1];
        for (int i = 0; i < n; ++i) x[i] = 0;  // first state in both orders: all photons in mode 0
        uint64_t cnt = 0;
        for (uint64_t r = 0; r < S; ++r) {
            if (p[r] > threshold) {
                if (cnt < cap) {
                    modes_to_state(x, n, n_modes, keys + cnt * (uint64_t)n_modes);
                    vals[cnt] = p[r];
                }
                ++cnt;
            }
            if (order == 0) colex_next(x, n, m_tot);
            else lex_next(x, n, m_tot);
        }
        *count_out = cnt;
        if (cnt > cap) return fail(LWB_E_LIMIT, "distribution has %llu entries, capacity %llu", (unsigned long long)cnt,
                                   (unsigned long long)cap);
        return 0;
    }

    // 2b. loss modes: marginalise on the device, order the keys by first appearance on the host
    const int k_min = backend == LWB_BACKEND_PERMANENT ? 1 : 0;  // permanent.py:148-149 skips zero-real-photon states
    uint64_t key_off[LWB_MAX_PHOTONS + 3];
    uint64_t nk = 0;
    for (int kk = 0; kk <= n + 1; ++kk) {
        key_off[kk] = nk;
        if (kk >= k_min && kk <= n) nk += binom(hb, n_modes + kk - 1, kk);
    }
    CK(ensure(h, h->bMarg, nk * 8));
    CK(ensure(h, h->bFirst, nk * 8));
    CK(ensure(h, h->bSmall, 4096));
    CU(cudaMemcpyAsync((char*)h->bSmall.p + 2048, key_off, sizeof(uint64_t) * (size_t)(n + 2), cudaMemcpyHostToDevice, h->stream));
    MargArgs ma;
    ma.bin = h->d_bin; ma.m_tot = m_tot; ma.n_modes = n_modes; ma.n = n; ma.k_min = k_min;
    ma.key_off = (const uint64_t*)((char*)h->bSmall.p + 2048); ma.n_keys = nk; ma.probs = d_probs;
    ma.threshold = threshold; ma.order = order; ma.marg = (double*)h->bMarg.p; ma.first_pos = (uint64_t*)h->bFirst.p;
    CK(launch_kmax(h, n, nk, 0, ma, marginalise_kernel<4>, marginalise_kernel<8>, marginalise_kernel<12>,
                   marginalise_kernel<16>, marginalise_kernel<24>, marginalise_kernel<40>));
    std::vector<double> marg(nk);
    std::vector<uint64_t> first(nk);
    CU(cudaMemcpyAsync(marg.data(), h->bMarg.p, nk * 8, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(first.data(), h->bFirst.p, nk * 8, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    std::vector<uint64_t> idx;
    for (uint64_t k = 0; k < nk; ++k)
        if (first[k] != UINT64_MAX) idx.push_back(k);
    std::sort(idx.begin(), idx.end(), [&](uint64_t a, uint64_t b) { return first[a] < first[b]; });
    uint64_t cnt = 0;
    // sum(pdist.values()) in dict order, permanent.py:159.  CPython >= 3.12 adds exact floats with Neumaier
    // compensation (bltinmodule.c cs_add); the vacuum entry 1 - total below inherits that rounding, so it is restated.
    double total = 0.0, comp = 0.0;
    auto py_add = [&](double x) {
        const double t = total + x;
        if (std::fabs(total) >= std::fabs(x)) comp += (total - t) + x;
        else comp += (x - t) + total;
        total = t;
    };
    for (uint64_t k : idx) {
        int kk = k_min;
        while (kk < n && k >= key_off[kk + 1]) ++kk;
        if (cnt < cap) {
            int x[LWB_MAX_PHOTONS + 1];
            colex_unrank(hb, k - key_off[kk], kk, n_modes, x);
            modes_to_state(x, kk, n_modes, keys + cnt * (uint64_t)n_modes);
            vals[cnt] = marg[k];
        }
        py_add(marg[k]);
        ++cnt;
    }
    total += comp;
    if (backend == LWB_BACKEND_PERMANENT && total < 1.0) {  // permanent.py:160-161 (loss_modes > 0 here)
        if (cnt < cap) {
            memset(keys + cnt * (uint64_t)n_modes, 0, (size_t)n_modes);
            vals[cnt] = 1.0 - total;
        }
        ++cnt;
    }
    *count_out = cnt;
    if (cnt > cap) return fail(LWB_E_LIMIT, "distribution has %llu entries, capacity %llu", (unsigned long long)cnt,
                               (unsigned long long)cap);
    return 0;
}

extern "C" {

int lwb_slos_amplitudes_dev(lwb_handle h, const double* d_U, int m, const uint8_t* in_state, int order, double* d_amps) {
    if (!h || !in_state) return fail(LWB_E_ARG, "null argument");
    return slos_run_dev(h, d_U, m, in_state, order, 0, d_amps);
}
int lwb_slos_amplitudes(lwb_handle h, const double* U, int m, const uint8_t* in_state, int order, double* amps) {
    return slos_run_host(h, U, m, in_state, order, 0, amps);
}
int lwb_slos_basis_probs_dev(lwb_handle h, const double* d_U, int m, const uint8_t* in_state, int order, double* d_probs) {
    if (!h || !in_state) return fail(LWB_E_ARG, "null argument");
    return slos_run_dev(h, d_U, m, in_state, order, 1, d_probs);
}
int lwb_slos_basis_probs(lwb_handle h, const double* U, int m, const uint8_t* in_state, int order, double* probs) {
    return slos_run_host(h, U, m, in_state, order, 1, probs);
}

// ---- full distribution -----------------------------------------------------------------------------
int lwb_pdist(lwb_handle h, const double* U_full, int m_tot, int n_modes, const uint8_t* in_state, double threshold,
              int backend, int dtype, uint8_t* keys, double* vals, uint64_t cap, uint64_t* count_out) {
    if (!h || !U_full || !in_state || !keys || !vals || !count_out) return fail(LWB_E_ARG, "null argument");
    if (backend != LWB_BACKEND_PERMANENT && backend != LWB_BACKEND_SLOS) return fail(LWB_E_ARG, "backend %d", backend);
    if (n_modes < 1 || n_modes > m_tot) return fail(LWB_E_ARG, "n_modes=%d outside [1,%d]", n_modes, m_tot);
    if (m_tot > LWB_MAX_MODES) return fail(LWB_E_LIMIT, "modes m=%d outside [1,%d]", m_tot, LWB_MAX_MODES);
    std::lock_guard<std::mutex> lk(h->mu);
    const uint64_t* hb = host_bin();
    const int loss = m_tot - n_modes;
    int n = 0;
    for (int i = 0; i < n_modes; ++i) n += in_state[i];
    *count_out = 0;
    if (n == 0) {  // permanent.py:137-138, slos.py:64-65
        *count_out = 1;
        if (cap < 1) return fail(LWB_E_LIMIT, "capacity");
        memset(keys, 0, (size_t)n_modes);
        vals[0] = 1.0;
        return 0;
    }
    CK(check_mn(m_tot, n));
    if (backend == LWB_BACKEND_PERMANENT && n > LWB_K1_MAX_N)
        return fail(LWB_E_LIMIT, "permanent backend supports up to %d photons (got %d)", LWB_K1_MAX_N, n);
    CU(cudaSetDevice(h->device));
    const uint64_t S = binom(hb, m_tot + n - 1, n);
    uint8_t full_in[LWB_MAX_MODES];
    memset(full_in, 0, sizeof full_in);
    memcpy(full_in, in_state, (size_t)n_modes);  // permanent.py:142-143, slos.py:69-70

    // 1. per-full-state probabilities on the device, colex order
    CK(ensure(h, h->bU, (size_t)m_tot * m_tot * 16));
    CK(ensure(h, h->bProbs, S * 8));
    CU(cudaMemcpyAsync(h->bU.p, U_full, (size_t)m_tot * m_tot * 16, cudaMemcpyHostToDevice, h->stream));
    if (backend == LWB_BACKEND_PERMANENT) {
        CK(ensure(h, h->bIn, (size_t)m_tot));
        CU(cudaMemcpyAsync(h->bIn.p, full_in, (size_t)m_tot, cudaMemcpyHostToDevice, h->stream));
        CK(launch_states(h, (const double*)h->bU.p, m_tot, (const uint8_t*)h->bIn.p, 1, nullptr, 0, S, n, dtype, 1,
                         (double*)h->bProbs.p));
    } else {
        // without loss modes the SLOS order is produced directly; with loss the marginaliser wants colex
        CK(slos_run_dev(h, (const double*)h->bU.p, m_tot, full_in, loss == 0 ? 1 : 0, 1, (double*)h->bProbs.p));
    }

    return pdist_finish(h, (const double*)h->bProbs.p, m_tot, n_modes, n, threshold, backend, keys, vals, cap, count_out);
}

}  // extern "C"

// ---- on-device sampling (sampler.py:329-349) --------------------------------------------------------------
// d_total (optional, device): receives cdf[S-1], the sum of the kept probabilities
static int sample_dev(lwb_handle h, const double* d_probs, uint64_t S, double threshold, const double* d_uniform,
                      uint64_t n_samples, uint64_t* d_index, double* d_total) {
    if (S == 0) return fail(LWB_E_ARG, "empty distribution");
    CU(cudaSetDevice(h->device));
    const uint64_t n_tiles = (S + SCAN_TILE - 1) / SCAN_TILE;
    if (n_tiles > 0x7fffffffull) return fail(LWB_E_LIMIT, "distribution too large");
    CK(ensure(h, h->bCdf, S * 8));
    CK(ensure(h, h->bTile, (2 * n_tiles + 1) * 8));
    double* tile_sum = (double*)h->bTile.p;
    double* tile_off = tile_sum + n_tiles;
    double* cdf = (double*)h->bCdf.p;
    scan_tile_sums_kernel<<<(unsigned)n_tiles, SCAN_THREADS, 0, h->stream>>>(d_probs, S, threshold, tile_sum);
    scan_tile_offsets_kernel<<<1, 1024, 0, h->stream>>>(tile_sum, n_tiles, tile_off);
    scan_tile_cdf_kernel<<<(unsigned)n_tiles, SCAN_THREADS, 0, h->stream>>>(d_probs, S, threshold, tile_off, cdf);
    h->launches += 3;
    if (n_samples > 0) {
        const uint64_t blocks = (n_samples + 255) / 256;
        if (blocks > 0x7fffffffull) return fail(LWB_E_LIMIT, "too many samples");
        sample_search_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>(cdf, S, d_uniform, n_samples, d_index);
        h->launches++;
    }
    if (d_total) CU(cudaMemcpyAsync(d_total, cdf + (S - 1), 8, cudaMemcpyDeviceToDevice, h->stream));
    CU(cudaGetLastError());
    return 0;
}

// uniforms in, indices + total out around a device-resident distribution
static int sample_roundtrip(lwb_handle h, const double* d_probs, uint64_t S, double threshold, const double* uniform,
                            uint64_t n_samples, uint64_t* index, double* total) {
    CK(ensure(h, h->bUni, (n_samples + 1) * 8));
    CK(ensure(h, h->bIdx, (n_samples + 1) * 8));
    if (n_samples) CU(cudaMemcpyAsync(h->bUni.p, uniform, n_samples * 8, cudaMemcpyHostToDevice, h->stream));
    double* d_total = (double*)h->bUni.p + n_samples;
    CK(sample_dev(h, d_probs, S, threshold, (const double*)h->bUni.p, n_samples, (uint64_t*)h->bIdx.p, d_total));
    double tot = 0.0;
    if (n_samples) CU(cudaMemcpyAsync(index, h->bIdx.p, n_samples * 8, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(&tot, d_total, 8, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (total) *total = tot;
    if (!(tot > 0.0)) return fail(LWB_E_ARG, "no probability above the threshold: nothing to sample from");
    return 0;
}

extern "C" {

int lwb_sample_dev(lwb_handle h, const double* d_probs, uint64_t S, double threshold, const double* d_uniform,
                   uint64_t n_samples, uint64_t* d_index, double* d_total) {
    if (!h || !d_probs || (n_samples && (!d_uniform || !d_index))) return fail(LWB_E_ARG, "null argument");
    return sample_dev(h, d_probs, S, threshold, d_uniform, n_samples, d_index, d_total);
}

int lwb_sample(lwb_handle h, const double* probs, uint64_t S, double threshold, const double* uniform,
               uint64_t n_samples, uint64_t* index, double* total) {
    if (!h || !probs || (n_samples && (!uniform || !index))) return fail(LWB_E_ARG, "null argument");
    if (S == 0) return fail(LWB_E_ARG, "empty distribution");
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    CK(ensure(h, h->bProbs, S * 8));
    CU(cudaMemcpyAsync(h->bProbs.p, probs, S * 8, cudaMemcpyHostToDevice, h->stream));
    return sample_roundtrip(h, (const double*)h->bProbs.p, S, threshold, uniform, n_samples, index, total);
}

// basis distribution of one input on the device (h->bProbs), shared by the sampling entry points
static int basis_probs_for_sampling(lwb_handle h, int backend, const double* U, int m, const uint8_t* in_state, int* n_out,
                                    uint64_t* S_out) {
    if (backend != LWB_BACKEND_PERMANENT && backend != LWB_BACKEND_SLOS) return fail(LWB_E_ARG, "backend %d", backend);
    int n = 0;
    CK(slos_check(m, in_state, &n));  // size limits shared by both backends
    if (backend == LWB_BACKEND_PERMANENT && n > LWB_K1_MAX_N)
        return fail(LWB_E_LIMIT, "permanent backend supports up to %d photons (got %d)", LWB_K1_MAX_N, n);
    CU(cudaSetDevice(h->device));
    const uint64_t S = binom(host_bin(), m + n - 1, n);
    CK(ensure(h, h->bU, (size_t)m * m * 16));
    CK(ensure(h, h->bProbs, S * 8));
    CU(cudaMemcpyAsync(h->bU.p, U, (size_t)m * m * 16, cudaMemcpyHostToDevice, h->stream));
    if (n == 0) {
        const double one = 1.0;
        CU(cudaMemcpyAsync(h->bProbs.p, &one, 8, cudaMemcpyHostToDevice, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    } else if (backend == LWB_BACKEND_PERMANENT) {
        CK(ensure(h, h->bIn, (size_t)m));
        CU(cudaMemcpyAsync(h->bIn.p, in_state, (size_t)m, cudaMemcpyHostToDevice, h->stream));
        CK(launch_states(h, (const double*)h->bU.p, m, (const uint8_t*)h->bIn.p, 1, nullptr, 0, S, n, LWB_C128, 1,
                         (double*)h->bProbs.p));
    } else {
        CK(slos_run_dev(h, (const double*)h->bU.p, m, in_state, 1, 1, (double*)h->bProbs.p));
    }
    *n_out = n;
    *S_out = S;
    return 0;
}

int lwb_basis_sample(lwb_handle h, int backend, const double* U, int m, const uint8_t* in_state, double threshold,
                     const double* uniform, uint64_t n_samples, uint64_t* index, double* total) {
    if (!h || !U || !in_state || (n_samples && (!uniform || !index))) return fail(LWB_E_ARG, "null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    int n = 0;
    uint64_t S = 0;
    CK(basis_probs_for_sampling(h, backend, U, m, in_state, &n, &S));
    return sample_roundtrip(h, (const double*)h->bProbs.p, S, threshold, uniform, n_samples, index, total);
}

int lwb_basis_sample_selected(lwb_handle h, int backend, const double* U, int m, const uint8_t* in_state, double threshold,
                              int threshold_detect, int min_detection, const int32_t* herald_modes,
                              const int32_t* herald_counts, int n_heralds, const int32_t* rule_modes,
                              const int32_t* rule_mode_offsets, const int32_t* rule_counts, const int32_t* rule_count_offsets,
                              int n_rules, const double* uniform, uint64_t n_samples, uint64_t* index, double* total) {
    if (!h || !U || !in_state || (n_samples && (!uniform || !index))) return fail(LWB_E_ARG, "null argument");
    if (n_heralds < 0 || n_heralds > 64 || (n_heralds && (!herald_modes || !herald_counts))) return fail(LWB_E_ARG, "heralds");
    if (n_rules < 0 || n_rules > MASK_MAX_RULES) return fail(LWB_E_LIMIT, "at most %d post-selection rules (got %d)", MASK_MAX_RULES, n_rules);
    if (n_rules && (!rule_modes || !rule_mode_offsets || !rule_counts || !rule_count_offsets)) return fail(LWB_E_ARG, "rules");
    std::lock_guard<std::mutex> lk(h->mu);
    int n = 0;
    uint64_t S = 0;
    CK(basis_probs_for_sampling(h, backend, U, m, in_state, &n, &S));
    if (n > 0) {
        SelectArgs sa;
        memset(&sa, 0, sizeof sa);
        sa.bin = h->d_bin; sa.m = m; sa.n = n; sa.S = S; sa.order = backend == LWB_BACKEND_SLOS ? 1 : 0;
        sa.threshold_detect = threshold_detect ? 1 : 0; sa.min_detection = min_detection; sa.n_heralds = n_heralds;
        for (int q = 0; q < n_heralds; ++q) {
            if (herald_modes[q] < 0 || herald_modes[q] >= m || herald_counts[q] < 0 || herald_counts[q] > 255)
                return fail(LWB_E_ARG, "herald %d: mode %d count %d", q, herald_modes[q], herald_counts[q]);
            sa.herald_mode[q] = (uint8_t)herald_modes[q];
            sa.herald_count[q] = (uint8_t)herald_counts[q];
        }
        sa.n_rules = n_rules;
        for (int q = 0; q < n_rules; ++q) {
            for (int e = rule_mode_offsets[q]; e < rule_mode_offsets[q + 1]; ++e) {
                if (rule_modes[e] < 0 || rule_modes[e] >= m) return fail(LWB_E_ARG, "rule %d: mode %d", q, rule_modes[e]);
                sa.rule_modes[q][rule_modes[e] >> 6] |= 1ull << (rule_modes[e] & 63);
            }
            for (int e = rule_count_offsets[q]; e < rule_count_offsets[q + 1]; ++e)
                if (rule_counts[e] >= 0 && rule_counts[e] < 64) sa.rule_counts[q] |= 1ull << rule_counts[e];
        }
        sa.probs = (double*)h->bProbs.p;
        CK(launch_kmax(h, n, S, 0, sa, select_outputs_kernel<4>, select_outputs_kernel<8>, select_outputs_kernel<12>,
                       select_outputs_kernel<16>, select_outputs_kernel<24>, select_outputs_kernel<40>));
    }
    return sample_roundtrip(h, (const double*)h->bProbs.p, S, threshold, uniform, n_samples, index, total);
}

}  // extern "C"

// ---- dense distributions and their merge-convolution (probability_distribution.py:25-164) -----------------
static uint64_t keyspace_off(int n_modes, int k) {  // states with fewer than k photons
    uint64_t o = 0;
    for (int j = 0; j < k; ++j) o += binom(host_bin(), n_modes + j - 1, j);
    return o;
}
static int keyspace_check(int n_modes, int kmax) {
    if (n_modes < 1 || n_modes > LWB_MAX_MODES) return fail(LWB_E_LIMIT, "modes m=%d outside [1,%d]", n_modes, LWB_MAX_MODES);
    if (kmax < 0 || kmax > LWB_MAX_PHOTONS) return fail(LWB_E_LIMIT, "photons n=%d outside [0,%d]", kmax, LWB_MAX_PHOTONS);
    if (n_modes + kmax - 1 >= LWB_BIN_A) return fail(LWB_E_LIMIT, "m+n too large");
    uint64_t o = 0;
    for (int j = 0; j <= kmax; ++j) {
        const uint64_t c = binom(host_bin(), n_modes + j - 1, j);
        if (c == UINT64_MAX || o > (1ull << 40) - c) return fail(LWB_E_LIMIT, "key space too large");
        o += c;
    }
    return 0;
}
static int new_slot(lwb_handle h, int n_modes, int kmin, int kmax, int* slot) {
    CK(keyspace_check(n_modes, kmax));
    lwb_context::DistSlot d;
    d.n_modes = n_modes; d.kmin = kmin; d.kmax = kmax; d.size = keyspace_off(n_modes, kmax + 1);
    for (size_t i = 0; i < h->dist_pool.size(); ++i)  // a parked buffer of exactly this size
        if (h->dist_pool[i].first == d.size) {
            d.p = h->dist_pool[i].second;
            h->dist_pool_bytes -= d.size * 8;
            h->dist_pool[i] = h->dist_pool.back();
            h->dist_pool.pop_back();
            break;
        }
    if (!d.p) {
        cudaError_t e = cudaMalloc((void**)&d.p, d.size * 8);
        if (e != cudaSuccess && !h->dist_pool.empty()) {  // out of memory with buffers parked: release them and retry
            cudaGetLastError();
            for (auto& kv : h->dist_pool) cudaFree(kv.second);
            h->dist_pool.clear();
            h->dist_pool_bytes = 0;
            e = cudaMalloc((void**)&d.p, d.size * 8);
        }
        if (e != cudaSuccess) return fail(LWB_E_NOMEM, "cudaMalloc(%llu): %s", (unsigned long long)(d.size * 8), cudaGetErrorString(e));
    }
    for (size_t i = 0; i < h->dists.size(); ++i)
        if (!h->dists[i].p) { h->dists[i] = d; *slot = (int)i; return 0; }
    h->dists.push_back(d);
    *slot = (int)h->dists.size() - 1;
    return 0;
}
static int get_slot(lwb_handle h, int slot, lwb_context::DistSlot** d) {
    if (!h) return fail(LWB_E_ARG, "null handle");
    if (slot < 0 || slot >= (int)h->dists.size() || !h->dists[slot].p) return fail(LWB_E_ARG, "no distribution in slot %d", slot);
    *d = &h->dists[slot];
    return 0;
}

int lwb::new_dist_slot(lwb_handle h, int n_modes, int kmin, int kmax, int* slot) { return new_slot(h, n_modes, kmin, kmax, slot); }

// full_probability_distribution (permanent.py:116-163 / slos.py:43-80) into the dense key-space vector d_dst
// (K(n_modes, n) entries, on h's device), everything enqueued on h->stream, no synchronisation.
int lwb::dist_from_circuit_enqueue(lwb_handle h, const double* d_U_full, int m_tot, int n_modes, const uint8_t* in_state,
                                   double threshold, int backend, int dtype, double* d_dst, uint64_t dst_size) {
    const uint64_t* hb = host_bin();
    const int loss = m_tot - n_modes;
    int n = 0;
    for (int i = 0; i < n_modes; ++i) n += in_state[i];
    CK(check_mn(m_tot, n));
    if (backend == LWB_BACKEND_PERMANENT && n > LWB_K1_MAX_N)
        return fail(LWB_E_LIMIT, "permanent backend supports up to %d photons (got %d)", LWB_K1_MAX_N, n);
    CU(cudaMemsetAsync(d_dst, 0, dst_size * 8, h->stream));
    if (n == 0) {  // permanent.py:137-138, slos.py:64-65
        const double one = 1.0;
        CU(cudaMemcpyAsync(d_dst, &one, 8, cudaMemcpyHostToDevice, h->stream));
        return 0;
    }
    uint8_t full_in[LWB_MAX_MODES];
    memset(full_in, 0, sizeof full_in);
    memcpy(full_in, in_state, (size_t)n_modes);  // permanent.py:142-143, slos.py:69-70
    const uint64_t S = binom(hb, m_tot + n - 1, n);
    CK(ensure(h, h->bProbs, S * 8));
    if (backend == LWB_BACKEND_PERMANENT) {
        CK(ensure(h, h->bIn, (size_t)LWB_MAX_MODES));
        CU(cudaMemcpyAsync(h->bIn.p, full_in, (size_t)m_tot, cudaMemcpyHostToDevice, h->stream));  // pageable: staged before return
        CK(launch_states(h, d_U_full, m_tot, (const uint8_t*)h->bIn.p, 1, nullptr, 0, S, n, dtype, 1, (double*)h->bProbs.p));
    } else {
        CK(slos_run_dev(h, d_U_full, m_tot, full_in, 0, 1, (double*)h->bProbs.p));  // fock_basis order
    }
    if (loss == 0) {
        dist_threshold_copy_kernel<<<(unsigned)((S + 255) / 256), 256, 0, h->stream>>>(d_dst + keyspace_off(n_modes, n),
                                                                                       (const double*)h->bProbs.p, threshold, S);
        h->launches++;
        CU(cudaGetLastError());
        return 0;
    }
    // loss modes: marginalise straight into the key space (the marginal table IS the key-space layout)
    const int k_min = backend == LWB_BACKEND_PERMANENT ? 1 : 0;  // permanent.py:148-149 skips zero-real-photon states
    uint64_t key_off[LWB_MAX_PHOTONS + 3];
    uint64_t nk = 0;
    for (int kk = 0; kk <= n + 1; ++kk) {
        key_off[kk] = nk;
        if (kk >= k_min && kk <= n) nk += binom(hb, n_modes + kk - 1, kk);
    }
    CK(ensure(h, h->bFirst, nk * 8));
    CK(ensure(h, h->bSmall, 4096));
    CU(cudaMemcpyAsync((char*)h->bSmall.p + 2048, key_off, sizeof(uint64_t) * (size_t)(n + 2), cudaMemcpyHostToDevice, h->stream));
    double* marg = d_dst + (k_min ? 1 : 0);
    MargArgs ma;
    ma.bin = h->d_bin; ma.m_tot = m_tot; ma.n_modes = n_modes; ma.n = n; ma.k_min = k_min;
    ma.key_off = (const uint64_t*)((char*)h->bSmall.p + 2048); ma.n_keys = nk; ma.probs = (const double*)h->bProbs.p;
    ma.threshold = threshold; ma.order = 0; ma.marg = marg; ma.first_pos = (uint64_t*)h->bFirst.p;
    CK(launch_kmax(h, n, nk, 0, ma, marginalise_kernel<4>, marginalise_kernel<8>, marginalise_kernel<12>,
                   marginalise_kernel<16>, marginalise_kernel<24>, marginalise_kernel<40>));
    if (backend == LWB_BACKEND_PERMANENT) {  // permanent.py:159-161: vacuum = 1 - sum if sum < 1
        const uint64_t n_tiles = (nk + SCAN_TILE - 1) / SCAN_TILE;
        CK(ensure(h, h->bTile, (2 * n_tiles + 1) * 8));
        double* tile_sum = (double*)h->bTile.p;
        double* tile_off = tile_sum + n_tiles;
        scan_tile_sums_kernel<<<(unsigned)n_tiles, SCAN_THREADS, 0, h->stream>>>(marg, nk, -1.0, tile_sum);
        scan_tile_offsets_kernel<<<1, 1024, 0, h->stream>>>(tile_sum, n_tiles, tile_off);
        dist_vacuum_kernel<<<1, 32, 0, h->stream>>>(d_dst, tile_off + n_tiles);
        h->launches += 3;
        CU(cudaGetLastError());
    }
    return 0;
}

extern "C" {

int lwb_keyspace_size(int n_modes, int kmax, uint64_t* size) {
    if (!size) return fail(LWB_E_ARG, "null argument");
    CK(keyspace_check(n_modes, kmax));
    *size = keyspace_off(n_modes, kmax + 1);
    return 0;
}

int lwb_keyspace_index(int n_modes, const uint8_t* state, uint64_t* index) {
    if (!state || !index) return fail(LWB_E_ARG, "null argument");
    const int k = photons_of(state, n_modes);
    CK(keyspace_check(n_modes, k));
    int x[LWB_MAX_PHOTONS + 1];
    state_to_modes(state, n_modes, x);
    *index = keyspace_off(n_modes, k) + colex_rank(host_bin(), x, k);
    return 0;
}

int lwb_keyspace_state(int n_modes, int kmax, uint64_t index, uint8_t* state) {
    if (!state) return fail(LWB_E_ARG, "null argument");
    CK(keyspace_check(n_modes, kmax));
    if (index >= keyspace_off(n_modes, kmax + 1)) return fail(LWB_E_ARG, "index outside the key space");
    int k = 0;
    while (index >= keyspace_off(n_modes, k + 1)) ++k;
    int x[LWB_MAX_PHOTONS + 1];
    colex_unrank(host_bin(), index - keyspace_off(n_modes, k), k, n_modes, x);
    modes_to_state(x, k, n_modes, state);
    return 0;
}

int lwb_keyspace_states(int n_modes, int kmax, const uint64_t* indices, uint64_t count, uint8_t* states) {
    if (count && (!indices || !states)) return fail(LWB_E_ARG, "null argument");
    CK(keyspace_check(n_modes, kmax));
    uint64_t off[LWB_MAX_PHOTONS + 2];
    for (int k = 0; k <= kmax + 1; ++k) off[k] = keyspace_off(n_modes, k);
    for (uint64_t c = 0; c < count; ++c) {
        const uint64_t index = indices[c];
        if (index >= off[kmax + 1]) return fail(LWB_E_ARG, "index outside the key space");
        int k = 0;
        while (index >= off[k + 1]) ++k;
        int x[LWB_MAX_PHOTONS + 1];
        colex_unrank(host_bin(), index - off[k], k, n_modes, x);
        modes_to_state(x, k, n_modes, states + c * (uint64_t)n_modes);
    }
    return 0;
}

int lwb_slos_states(int m, int n, const uint64_t* ranks, uint64_t count, uint8_t* states) {
    if (count && (!ranks || !states)) return fail(LWB_E_ARG, "null argument");
    CK(check_mn(m, n));
    const uint64_t S = binom(host_bin(), m + n - 1, n);
    for (uint64_t c = 0; c < count; ++c) {
        if (ranks[c] >= S) return fail(LWB_E_ARG, "rank out of range");
        int x[LWB_MAX_PHOTONS + 1];
        lex_unrank(host_bin(), ranks[c], n, m, x);
        modes_to_state(x, n, m, states + c * (uint64_t)m);
    }
    return 0;
}

int lwb_dist_create(lwb_handle h, int n_modes, int kmax, const double* dense, int* slot) {
    if (!h || !slot) return fail(LWB_E_ARG, "null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    CK(new_slot(h, n_modes, 0, kmax, slot));
    auto& d = h->dists[*slot];
    if (dense) {
        CU(cudaMemcpyAsync(d.p, dense, d.size * 8, cudaMemcpyHostToDevice, h->stream));
        CU(cudaStreamSynchronize(h->stream));  // the caller may reuse `dense` on return
    } else {
        CU(cudaMemsetAsync(d.p, 0, d.size * 8, h->stream));  // stream-ordered like every later use of the slot: no wait
    }
    return 0;
}

int lwb_dist_free(lwb_handle h, int slot) {
    if (!h) return fail(LWB_E_ARG, "null handle");
    std::lock_guard<std::mutex> lk(h->mu);
    lwb_context::DistSlot* d = nullptr;
    CK(get_slot(h, slot, &d));
    CU(cudaSetDevice(h->device));
    // park the buffer (up to 4 GiB in total): its next user is enqueued on the same stream, after every pending reader
    if (h->dist_pool_bytes + d->size * 8 <= (4ull << 30) && h->dist_pool.size() < 4096) {
        h->dist_pool.push_back({d->size, d->p});
        h->dist_pool_bytes += d->size * 8;
    } else {
        CU(cudaStreamSynchronize(h->stream));
        CU(cudaFree(d->p));
    }
    *d = lwb_context::DistSlot();
    return 0;
}

int lwb_dist_info(lwb_handle h, int slot, int* n_modes, int* kmax, uint64_t* size, void** device_ptr) {
    lwb_context::DistSlot* d = nullptr;
    CK(get_slot(h, slot, &d));
    if (n_modes) *n_modes = d->n_modes;
    if (kmax) *kmax = d->kmax;
    if (size) *size = d->size;
    if (device_ptr) *device_ptr = d->p;
    return 0;
}

int lwb_dist_download(lwb_handle h, int slot, double* dense) {
    lwb_context::DistSlot* d = nullptr;
    CK(get_slot(h, slot, &d));
    if (!dense) return fail(LWB_E_ARG, "null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    CU(cudaMemcpyAsync(dense, d->p, d->size * 8, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

int lwb_dist_axpy(lwb_handle h, int slot_acc, double alpha, int slot_x) {
    if (!h) return fail(LWB_E_ARG, "null handle");
    std::lock_guard<std::mutex> lk(h->mu);
    lwb_context::DistSlot *acc = nullptr, *x = nullptr;
    CK(get_slot(h, slot_acc, &acc));
    CK(get_slot(h, slot_x, &x));
    if (acc->n_modes != x->n_modes || acc->kmax < x->kmax)
        return fail(LWB_E_ARG, "accumulator (%d modes, <= %d photons) cannot hold (%d modes, <= %d photons)", acc->n_modes,
                    acc->kmax, x->n_modes, x->kmax);
    CU(cudaSetDevice(h->device));
    dist_axpy_kernel<<<(unsigned)((x->size + 255) / 256), 256, 0, h->stream>>>(acc->p, alpha, x->p, x->size);
    h->launches++;
    acc->kmin = std::min(acc->kmin, x->kmin);
    CU(cudaGetLastError());
    return 0;
}

int lwb_dist_axpy_many(lwb_handle h, int slot_acc, int n_terms, const double* alphas, const int* slots) {
    if (!h || (n_terms > 0 && (!alphas || !slots))) return fail(LWB_E_ARG, "null argument");
    if (n_terms <= 0) return 0;
    std::lock_guard<std::mutex> lk(h->mu);
    lwb_context::DistSlot* acc = nullptr;
    CK(get_slot(h, slot_acc, &acc));
    std::vector<DistAxpyTerm> terms((size_t)n_terms);
    uint64_t n_max = 0;
    int kmin = acc->kmin;
    for (int j = 0; j < n_terms; ++j) {
        lwb_context::DistSlot* x = nullptr;
        CK(get_slot(h, slots[j], &x));
        if (x == acc) return fail(LWB_E_ARG, "term %d is the accumulator itself", j);
        if (acc->n_modes != x->n_modes || acc->kmax < x->kmax)
            return fail(LWB_E_ARG, "accumulator (%d modes, <= %d photons) cannot hold term %d (%d modes, <= %d photons)",
                        acc->n_modes, acc->kmax, j, x->n_modes, x->kmax);
        terms[(size_t)j] = DistAxpyTerm{x->p, x->size, alphas[j]};
        n_max = std::max(n_max, x->size);
        kmin = std::min(kmin, x->kmin);
    }
    CU(cudaSetDevice(h->device));
    // the term table travels through a per-call device buffer: pageable source, so it is staged before the copy returns;
    // the buffer is reused by the next call only after this call's kernel (same stream)
    CK(ensure(h, h->bTerms, terms.size() * sizeof(DistAxpyTerm)));
    CU(cudaMemcpyAsync(h->bTerms.p, terms.data(), terms.size() * sizeof(DistAxpyTerm), cudaMemcpyHostToDevice, h->stream));
    dist_axpy_many_kernel<<<(unsigned)((n_max + 255) / 256), 256, 0, h->stream>>>(acc->p, (const DistAxpyTerm*)h->bTerms.p, n_terms,
                                                                                 n_max);
    h->launches++;
    acc->kmin = kmin;
    CU(cudaGetLastError());
    return 0;
}

int lwb_dist_convolve(lwb_handle h, int slot_a, int slot_b, int* slot_out) {
    if (!h || !slot_out) return fail(LWB_E_ARG, "null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    lwb_context::DistSlot *pa = nullptr, *pb = nullptr;
    CK(get_slot(h, slot_a, &pa));
    CK(get_slot(h, slot_b, &pb));
    if (pa->n_modes != pb->n_modes) return fail(LWB_E_ARG, "mode counts differ (%d, %d)", pa->n_modes, pb->n_modes);
    CU(cudaSetDevice(h->device));
    const lwb_context::DistSlot A = *pa, B = *pb;  // new_slot may move the vector
    CK(new_slot(h, A.n_modes, A.kmin + B.kmin, A.kmax + B.kmax, slot_out));
    const auto& o = h->dists[*slot_out];
    ConvolveArgs ca;
    ca.bin = h->d_bin; ca.n_modes = A.n_modes; ca.ka_min = A.kmin; ca.ka_max = A.kmax; ca.kb_min = B.kmin; ca.kb_max = B.kmax;
    ca.A = A.p; ca.B = B.p; ca.out = o.p; ca.size_out = o.size;
    for (int k = 0; k <= o.kmax + 1; ++k) ca.off[k] = keyspace_off(A.n_modes, k);
    const int rc = launch_kmax(h, o.kmax, o.size, 0, ca, convolve_kernel<4>, convolve_kernel<8>, convolve_kernel<12>,
                               convolve_kernel<16>, convolve_kernel<24>, convolve_kernel<40>);
    if (rc) {  // leave no slot behind on failure
        cudaFree(h->dists[*slot_out].p);
        h->dists[*slot_out] = lwb_context::DistSlot();
        *slot_out = -1;
    }
    return rc;
}

int lwb_dist_sample(lwb_handle h, int slot, const double* uniform, uint64_t n_samples, uint64_t* index, double* total) {
    lwb_context::DistSlot* d = nullptr;
    CK(get_slot(h, slot, &d));
    if (n_samples && (!uniform || !index)) return fail(LWB_E_ARG, "null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    return sample_roundtrip(h, d->p, d->size, -1.0, uniform, n_samples, index, total);
}

int lwb_dist_from_circuit(lwb_handle h, const double* U_full, int m_tot, int n_modes, const uint8_t* in_state,
                          double threshold, int backend, int dtype, int* slot) {
    return lwb_dists_from_circuit(h, nullptr, U_full, m_tot, n_modes, in_state, 1, threshold, backend, dtype, slot);
}

}  // extern "C"
