// comm.cu -- multi-GPU communicators behind the C ABI (include/lwb200.h, "multi-GPU" section).
//
// The path shards where the reference loops over independent units:
//   * output Fock states of one full distribution (permanent.py:145-157): contiguous rank ranges of the fock_basis
//     order holding EQUAL WORK of the multiplicity-aware walk (lwb_shard_cuts: exact cumulative-work bisection);
//   * the Gray-code space of one large permanent (lwb_perm_single's slices) + a 16-byte sum;
//   * the unique sub-inputs of an imperfect source (probability_distribution.py:127-133): round-robin over the GPUs.
//
// Two communicator kinds share the code:
//   LOCAL  one process drives n GPUs (lwb_comm_init_local): one lwb_handle per device, peer access enabled, every
//          launch is asynchronous so a single host thread keeps all devices busy.  This is what
//          lightworks_b200.install(n_gpus=N) gives `Backend("permanent")`.
//   RANK   one process per GPU (lwb_comm_init_rank), bootstrapped with an NCCL unique id the host framework
//          broadcasts (torchrun / MPI / a file).  NCCL is loaded with dlopen("libnccl.so.2") so the library has no
//          link-time dependency and shares the NCCL a host framework already loaded.
//
// Exchange step.  The only data the ranks exchange is the probability shards (8 bytes per permanent) and one sum.
// When the GPUs can map each other's memory (NVLink peer access; CUDA IPC across processes) the permanent kernels
// write every probability straight into the peers' copies of the distribution as it is produced (MirrorSet in the
// kernel arguments: fire-and-forget 8-byte stores over NVLink, hidden behind ~40 000 flops each), so the step ends
// with the scalar all-reduce of the sum, which is also the barrier that tells every rank its copy is complete.
// Without peer mapping the shards travel by NCCL broadcasts (one group call, in place, unequal lengths).
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "lwb_internal.h"

using namespace lwb;

// ------------------------------------------------------------------------------------------------
// NCCL through dlopen: the handful of entry points used, declared with their public C ABI (nccl.h, NCCL 2.x)
namespace {
typedef struct ncclComm* nccl_comm_t;
struct nccl_uid { char internal[128]; };
enum { NCCL_SUM = 0, NCCL_MIN = 3 };
enum { NCCL_UINT8 = 1, NCCL_INT32 = 2, NCCL_FLOAT64 = 8 };
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(nccl_uid*) = nullptr;
    int (*CommInitRank)(nccl_comm_t*, int, nccl_uid, int) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string err;
};
NcclApi g_nccl;
std::once_flag g_nccl_once;

int load_nccl() {
    std::call_once(g_nccl_once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (g_nccl.lib) break;
        }
        if (!g_nccl.lib) { g_nccl.err = std::string("dlopen(libnccl.so.2): ") + dlerror(); return; }
#define LWB_SYM(field, name)                                                       \
    *(void**)(&g_nccl.field) = dlsym(g_nccl.lib, name);                            \
    if (!g_nccl.field) { g_nccl.err = std::string("NCCL symbol missing: ") + name; return; }
        LWB_SYM(GetUniqueId, "ncclGetUniqueId")
        LWB_SYM(CommInitRank, "ncclCommInitRank")
        LWB_SYM(CommDestroy, "ncclCommDestroy")
        LWB_SYM(AllReduce, "ncclAllReduce")
        LWB_SYM(AllGather, "ncclAllGather")
        LWB_SYM(Broadcast, "ncclBroadcast")
        LWB_SYM(GroupStart, "ncclGroupStart")
        LWB_SYM(GroupEnd, "ncclGroupEnd")
        LWB_SYM(GetErrorString, "ncclGetErrorString")
#undef LWB_SYM
    });
    if (!g_nccl.err.empty()) return fail(LWB_E_COMM, "%s", g_nccl.err.c_str());
    return 0;
}
#define NC(call)                                                                                          \
    do {                                                                                                  \
        int r_ = (call);                                                                                  \
        if (r_ != 0) return fail(LWB_E_COMM, "%s: %s (%s:%d)", #call, g_nccl.GetErrorString(r_), __FILE__, __LINE__); \
    } while (0)
}  // namespace

// ------------------------------------------------------------------------------------------------
// Equal-work shard boundaries (host, exact integers): the C++ form of lightworks_b200/sharding.py, which stays as
// the readable statement of the same arithmetic (tests compare the two cut for cut).
namespace {
constexpr int WB_A = 320, WB_B = 48;  // C(a, b) for a < 320, min(b, a - b) < 48
uint64_t g_wbin[WB_A][WB_B];
std::once_flag g_wbin_once;
const uint64_t SAT = UINT64_MAX;

uint64_t wbinom(int a, int b) {
    std::call_once(g_wbin_once, [] {
        for (int x = 0; x < WB_A; ++x)
            for (int y = 0; y < WB_B; ++y) {
                uint64_t v;
                if (y == 0) v = 1;
                else if (x == 0) v = 0;
                else {
                    const uint64_t p = g_wbin[x - 1][y - 1], q = g_wbin[x - 1][y];
                    v = (p == SAT || q == SAT || p > SAT - q) ? SAT : p + q;
                }
                g_wbin[x][y] = v;
            }
    });
    if (b < 0 || a < 0 || b > a) return 0;
    if (a - b < b) b = a - b;
    if (a >= WB_A || b >= WB_B) return SAT;
    return g_wbin[a][b];
}

// Over all ways to put i photons into v modes: (sum of prod(s_j + 1), the same sum over states whose occupations are
// all even).  Generating functions (1 - z)^(-2v) and ((1 + z^2) / (1 - z^2)^2)^v  (sharding._free_work).
bool free_work(int i, int v, unsigned __int128* total, unsigned __int128* even) {
    if (v == 0) { *total = *even = (i == 0) ? 1 : 0; return true; }
    const uint64_t t = wbinom(i + 2 * v - 1, i);
    if (t == SAT) return false;
    *total = t;
    unsigned __int128 e = 0;
    if (i % 2 == 0) {
        const int h = i / 2;
        for (int a = 0; a <= std::min(v, h); ++a) {
            const uint64_t c1 = wbinom(v, a), c2 = wbinom(h - a + 2 * v - 1, 2 * v - 1);
            if (c1 == SAT || c2 == SAT) return false;
            e += (unsigned __int128)c1 * c2;
        }
    }
    *even = e;
    return true;
}

// Exact work (terms of the multiplicity-aware walk) of the rank prefix [0, rank) of fock_basis(m, n)
// (sharding.cumulative_work without the per-state overhead).
bool prefix_work(int m, int n, uint64_t rank, uint64_t total_states, unsigned __int128* out) {
    unsigned __int128 a, e;
    if (rank == 0) { *out = 0; return true; }
    if (rank >= total_states) {
        if (!free_work(n, m, &a, &e)) return false;
        *out = e + (a - e) / 2;
        return true;
    }
    int y[LWB_MAX_PHOTONS + 1];
    colex_unrank(host_bin(), rank, n, m, y);
    unsigned __int128 work = 0;
    for (int i = n; i >= 1; --i) {  // states with x_j = y_j for j > i and x_i < y_i
        const int v = y[i - 1];     // free photons live in modes 0 .. v-1
        if (v == 0) continue;
        unsigned __int128 f = 1;
        bool odd = false;
        for (int j = i; j < n;) {   // runs of equal modes among the fixed photons y[i..n)
            int k = j;
            while (k < n && y[k] == y[j]) ++k;
            const int c = k - j;
            f *= (unsigned)(c + 1);
            odd |= (c % 2 == 1);
            j = k;
        }
        if (!free_work(i, v, &a, &e)) return false;
        work += odd ? (f * a) / 2 : f * (e + (a - e) / 2);
    }
    *out = work;
    return true;
}
}  // namespace

// per-state cost of classification, unranking and block prologue in units of one Glynn term (sharding.STATE_OVERHEAD_TERMS)
static const double kStateOverheadTerms = 150.0;

extern "C" int lwb_shard_cuts(int m, int n, int n_fractions, const double* fractions, uint64_t* cuts) {
    if (!fractions || !cuts || n_fractions < 1) return fail(LWB_E_ARG, "null argument");
    CK(check_mn(m, n));
    const uint64_t total = binom(host_bin(), m + n - 1, n);
    auto uniform = [&]() {
        for (int k = 0; k < n_fractions; ++k) {
            const double r = std::nearbyint((double)total * fractions[k]);  // Python's round(): half to even
            cuts[k] = r <= 0 ? 0 : (r >= (double)total ? total : (uint64_t)r);
        }
        return 0;
    };
    if (n < 7 || !k1x_supported(n) || total >= (1ull << 32)) return uniform();  // expanded-row kernel: uniform cost per state
    unsigned __int128 w;
    if (!prefix_work(m, n, total, total, &w)) return uniform();
    auto cum = [&](uint64_t rank, double* out) {
        unsigned __int128 x;
        if (!prefix_work(m, n, rank, total, &x)) return false;
        *out = (double)x + kStateOverheadTerms * (double)rank;
        return true;
    };
    const double w_total = (double)w + kStateOverheadTerms * (double)total;
    uint64_t prev = 0;
    for (int k = 0; k < n_fractions; ++k) {
        const double f = fractions[k];
        if (f <= 0) { cuts[k] = 0; prev = 0; continue; }
        if (f >= 1) { cuts[k] = total; prev = total; continue; }
        const double target = w_total * f;
        uint64_t lo = prev, hi = total;
        while (lo < hi) {
            const uint64_t mid = lo + (hi - lo) / 2;
            double c;
            if (!cum(mid, &c)) return uniform();
            if (c < target) lo = mid + 1; else hi = mid;
        }
        cuts[k] = lo;
        prev = lo;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
enum { COMM_LOCAL = 0, COMM_RANK = 1 };

struct lwb_comm_s {
    int kind = COMM_LOCAL;
    int rank = 0, world = 1;
    std::vector<lwb_handle> h;      // LOCAL: one per device (owned); RANK: the caller's handle (borrowed)
    nccl_comm_t nccl = nullptr;     // RANK
    // the distribution buffers: full[g] is a pointer, valid on THIS process' device(s), to the copy of the whole
    // distribution that lives on GPU g.  LOCAL: all are ours.  RANK: full[rank] is ours, the others are IPC mappings.
    // Two halves, used alternately: a rank can run at most one sharded call ahead of a peer (the closing all-reduce
    // of call k needs every rank's kernels of call k), so the peer stores of call k+1 land in the other half and
    // never overwrite a distribution a peer's stream is still consuming.
    std::vector<double*> base;      // allocation on / mapping of GPU g (2 * cap bytes)
    std::vector<double*> full;      // base[g] + parity * cap / 8: the half the current call writes
    int parity = 0;
    size_t cap = 0;                 // bytes of each half
    bool peer_ok = false;           // kernels may store into the peers' buffers
    int use_peer_stores = 1;        // lwb_comm_set_option("peer_stores", 0/1)
    std::vector<double*> d_sum;     // per handle: [0] shard sum (all-reduced in place), [1..2] permanent partials
    void* d_stage = nullptr;        // RANK: staging for the handle exchange
    std::vector<uint64_t> cuts;     // cached shard boundaries of (cuts_m, cuts_n)
    int cuts_m = -1, cuts_n = -1;
    std::mutex mu;
};

static int comm_cuts(lwb_comm c, int m, int n) {
    if (c->cuts_m == m && c->cuts_n == n) return 0;
    std::vector<double> fr(c->world + 1);
    for (int k = 0; k <= c->world; ++k) fr[k] = (double)k / (double)c->world;
    c->cuts.assign(c->world + 1, 0);
    CK(lwb_shard_cuts(m, n, c->world + 1, fr.data(), c->cuts.data()));
    c->cuts_m = m;
    c->cuts_n = n;
    return 0;
}

static int rank_barrier(lwb_comm c) {  // RANK: every rank's stream work up to here has finished
    lwb_handle h = c->h[0];
    CU(cudaMemsetAsync(c->d_sum[0] + 3, 0, 8, h->stream));
    NC(g_nccl.AllReduce(c->d_sum[0] + 3, c->d_sum[0] + 3, 1, NCCL_FLOAT64, NCCL_SUM, c->nccl, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

static void select_half(lwb_comm c, int parity) {
    c->parity = parity;
    for (int g = 0; g < c->world; ++g) c->full[g] = c->base[g] ? c->base[g] + (size_t)parity * (c->cap / 8) : nullptr;
}

static int release_buffers(lwb_comm c) {
    if (c->base.empty()) return 0;
    if (c->kind == COMM_LOCAL) {
        for (int g = 0; g < c->world; ++g)
            if (c->base[g]) { cudaSetDevice(c->h[g]->device); cudaStreamSynchronize(c->h[g]->stream); cudaFree(c->base[g]); }
    } else {
        cudaSetDevice(c->h[0]->device);
        cudaStreamSynchronize(c->h[0]->stream);
        for (int g = 0; g < c->world; ++g)
            if (g != c->rank && c->base[g]) cudaIpcCloseMemHandle(c->base[g]);
        if (c->nccl) rank_barrier(c);  // nobody frees memory a peer still has mapped and may be writing
        if (c->base[c->rank]) cudaFree(c->base[c->rank]);
    }
    c->base.clear();
    c->full.clear();
    c->cap = 0;
    return 0;
}

// Collective (RANK: every rank calls it with the same size): make sure each GPU holds two halves of `bytes`, and
// switch to the other half.
static int ensure_buffers(lwb_comm c, size_t bytes) {
    if (c->cap >= bytes && !c->base.empty()) { select_half(c, c->parity ^ 1); return 0; }
    CK(release_buffers(c));
    bytes = (bytes + 4095) & ~(size_t)4095;
    c->base.assign(c->world, nullptr);
    c->full.assign(c->world, nullptr);
    if (c->kind == COMM_LOCAL) {
        for (int g = 0; g < c->world; ++g) {
            CU(cudaSetDevice(c->h[g]->device));
            cudaError_t e = cudaMalloc((void**)&c->base[g], 2 * bytes);
            if (e != cudaSuccess) return fail(LWB_E_NOMEM, "cudaMalloc(%zu) on device %d: %s", 2 * bytes, c->h[g]->device, cudaGetErrorString(e));
        }
        c->cap = bytes;
        select_half(c, 0);
        return 0;
    }
    lwb_handle h = c->h[0];
    CU(cudaSetDevice(h->device));
    cudaError_t e = cudaMalloc((void**)&c->base[c->rank], 2 * bytes);
    if (e != cudaSuccess) return fail(LWB_E_NOMEM, "cudaMalloc(%zu): %s", 2 * bytes, cudaGetErrorString(e));
    c->cap = bytes;
    // exchange CUDA IPC handles of the buffers through NCCL (64 bytes per rank) and map the peers' buffers
    int ok = 1;
    cudaIpcMemHandle_t mine;
    if (cudaIpcGetMemHandle(&mine, c->base[c->rank]) != cudaSuccess) { ok = 0; cudaGetLastError(); memset(&mine, 0, sizeof mine); }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    std::vector<cudaIpcMemHandle_t> all(c->world);
    CU(cudaMemcpyAsync((char*)c->d_stage + 64 * c->rank, &mine, 64, cudaMemcpyHostToDevice, h->stream));
    NC(g_nccl.AllGather((char*)c->d_stage + 64 * c->rank, c->d_stage, 64, NCCL_UINT8, c->nccl, h->stream));
    CU(cudaMemcpyAsync(all.data(), c->d_stage, 64 * (size_t)c->world, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    for (int g = 0; g < c->world && ok; ++g) {
        if (g == c->rank) continue;
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, all[g], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); break; }
        c->base[g] = (double*)p;
    }
    // peer stores only if EVERY rank mapped every peer
    int* d_flag = (int*)((char*)c->d_stage + 64 * c->world);
    CU(cudaMemcpyAsync(d_flag, &ok, 4, cudaMemcpyHostToDevice, h->stream));
    NC(g_nccl.AllReduce(d_flag, d_flag, 1, NCCL_INT32, NCCL_MIN, c->nccl, h->stream));
    int all_ok = 0;
    CU(cudaMemcpyAsync(&all_ok, d_flag, 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    c->peer_ok = all_ok != 0;
    if (!c->peer_ok)
        for (int g = 0; g < c->world; ++g)
            if (g != c->rank && c->base[g]) { cudaIpcCloseMemHandle(c->base[g]); c->base[g] = nullptr; }
    select_half(c, 0);
    return 0;
}

extern "C" {

int lwb_comm_unique_id(uint8_t* id128) {
    if (!id128) return fail(LWB_E_ARG, "null argument");
    CK(load_nccl());
    nccl_uid id;
    NC(g_nccl.GetUniqueId(&id));
    memcpy(id128, id.internal, 128);
    return 0;
}

int lwb_comm_init_rank(lwb_handle h, int rank, int world, const uint8_t* id128, lwb_comm* out) {
    if (!h || !id128 || !out) return fail(LWB_E_ARG, "null argument");
    if (world < 1 || rank < 0 || rank >= world || world > LWB_MAX_PEERS + 1) return fail(LWB_E_ARG, "rank %d of %d", rank, world);
    CK(load_nccl());
    CU(cudaSetDevice(h->device));
    auto* c = new lwb_comm_s();
    c->kind = COMM_RANK;
    c->rank = rank;
    c->world = world;
    c->h.push_back(h);
    nccl_uid id;
    memcpy(id.internal, id128, 128);
    int r = g_nccl.CommInitRank(&c->nccl, world, id, rank);
    if (r != 0) { delete c; return fail(LWB_E_COMM, "ncclCommInitRank: %s", g_nccl.GetErrorString(r)); }
    c->d_sum.assign(1, nullptr);
    CU(cudaMalloc((void**)&c->d_sum[0], 64));
    CU(cudaMalloc(&c->d_stage, 64 * (size_t)world + 64));
    *out = c;
    return 0;
}

int lwb_comm_init_local(int n_gpus, lwb_comm* out) {
    if (!out) return fail(LWB_E_ARG, "null argument");
    int count = 0;
    CU(cudaGetDeviceCount(&count));
    if (n_gpus < 1 || n_gpus > count || n_gpus > LWB_MAX_PEERS + 1) return fail(LWB_E_ARG, "n_gpus=%d with %d visible devices", n_gpus, count);
    auto* c = new lwb_comm_s();
    c->kind = COMM_LOCAL;
    c->world = n_gpus;
    c->d_sum.assign(n_gpus, nullptr);
    for (int d = 0; d < n_gpus; ++d) {
        lwb_handle hd = nullptr;
        int rc = lwb_create(d, &hd);
        if (rc) { for (auto x : c->h) lwb_destroy(x); delete c; return rc; }
        c->h.push_back(hd);
        CU(cudaMalloc((void**)&c->d_sum[d], 64));
    }
    bool ok = true;
    for (int i = 0; i < n_gpus; ++i)
        for (int j = 0; j < n_gpus; ++j) {
            if (i == j) continue;
            int can = 0;
            CU(cudaDeviceCanAccessPeer(&can, i, j));
            if (!can) { ok = false; continue; }
            CU(cudaSetDevice(i));
            cudaError_t e = cudaDeviceEnablePeerAccess(j, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) { ok = false; cudaGetLastError(); }
        }
    c->peer_ok = ok;
    *out = c;
    return 0;
}

int lwb_comm_destroy(lwb_comm c) {
    if (!c) return 0;
    release_buffers(c);
    for (size_t g = 0; g < c->d_sum.size(); ++g)
        if (c->d_sum[g]) { cudaSetDevice(c->h[g]->device); cudaFree(c->d_sum[g]); }
    if (c->d_stage) cudaFree(c->d_stage);
    if (c->kind == COMM_LOCAL)
        for (auto x : c->h) lwb_destroy(x);
    if (c->nccl) g_nccl.CommDestroy(c->nccl);
    delete c;
    return 0;
}

int lwb_comm_info(lwb_comm c, int* rank, int* world, int* kind, int* peer_stores) {
    if (!c) return fail(LWB_E_ARG, "null communicator");
    if (rank) *rank = c->rank;
    if (world) *world = c->world;
    if (kind) *kind = c->kind;
    if (peer_stores) *peer_stores = (c->peer_ok && c->use_peer_stores) ? 1 : 0;
    return 0;
}

/* debugging aid: the calling thread's pending CUDA error (cleared), 0 if none */
int lwb_debug_pending_cuda_error(char* msg, int len) {
    cudaError_t e = cudaGetLastError();
    if (msg && len > 0) { strncpy(msg, e == cudaSuccess ? "" : cudaGetErrorString(e), (size_t)len - 1); msg[len - 1] = 0; }
    return (int)e;
}

int lwb_comm_set_option(lwb_comm c, const char* name, int value) {
    if (!c || !name) return fail(LWB_E_ARG, "null argument");
    if (!strcmp(name, "peer_stores")) { c->use_peer_stores = value ? 1 : 0; return 0; }
    return fail(LWB_E_ARG, "unknown communicator option %s", name);
}

int lwb_comm_handle(lwb_comm c, int index, lwb_handle* out) {
    if (!c || !out || index < 0 || index >= (int)c->h.size()) return fail(LWB_E_ARG, "bad argument");
    *out = c->h[index];
    return 0;
}

int lwb_comm_shard(lwb_comm c, int m, int n, int rank, uint64_t* r0, uint64_t* r1) {
    if (!c || !r0 || !r1 || rank < 0 || rank >= c->world) return fail(LWB_E_ARG, "bad argument");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(comm_cuts(c, m, n));
    *r0 = c->cuts[rank];
    *r1 = c->cuts[rank + 1];
    return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// One rank's part of a sharded full-basis distribution, enqueued on its handle's stream: kernels (mirrored into the
// peers when possible) and the deterministic shard sum into d_sum[0].
static int shard_launch(lwb_comm c, int g /* rank whose shard */, lwb_handle h, const double* d_U, int m,
                        const uint8_t* d_in, int n, int dtype, int gather_all, double* d_sum) {
    const uint64_t r0 = c->cuts[g], cnt = c->cuts[g + 1] - c->cuts[g];
    double* mine = c->full[c->kind == COMM_LOCAL ? g : c->rank];
    h->mirror.n = 0;
    if (c->peer_ok && c->use_peer_stores) {
        for (int p = 0; p < c->world; ++p) {
            if (p == g) continue;
            if (!gather_all && p != 0) continue;  // LOCAL root gather: only device 0 collects
            h->mirror.p[h->mirror.n++] = c->full[p] + r0;
        }
    }
    int rc = 0;
    if (cnt) rc = launch_states(h, d_U, m, d_in, 1, nullptr, r0, cnt, n, dtype, 1, mine + r0);
    h->mirror.n = 0;
    if (rc) return rc;
    if (cnt) return sum_doubles_dev(h, mine + r0, cnt, d_sum);
    CU(cudaMemsetAsync(d_sum, 0, 8, h->stream));
    return 0;
}

static int comm_check_state(int m, const uint8_t* in_state, int* n_out) {
    if (m < 1 || m > LWB_MAX_MODES) return fail(LWB_E_LIMIT, "modes m=%d outside [1,%d]", m, LWB_MAX_MODES);
    int n = 0;
    for (int i = 0; i < m; ++i) n += in_state[i];
    CK(check_mn(m, n));
    if (n < 1 || n > LWB_K1_MAX_N) return fail(LWB_E_LIMIT, "sharded distributions support 1..%d photons (got %d)", LWB_K1_MAX_N, n);
    *n_out = n;
    return 0;
}

extern "C" {

// RANK communicators: device-resident inputs of THIS rank; asynchronous on the handle's stream.
int lwb_comm_perm_basis_probs_dev(lwb_comm c, const double* d_U, int m, const uint8_t* d_in_state, int n_photons,
                                  int dtype, double** d_probs, double** d_sum) {
    if (!c || !d_U || !d_in_state) return fail(LWB_E_ARG, "null argument");
    if (c->kind != COMM_RANK) return fail(LWB_E_ARG, "device-pointer entry point needs a per-rank communicator");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(check_mn(m, n_photons));
    if (n_photons < 1 || n_photons > LWB_K1_MAX_N) return fail(LWB_E_LIMIT, "photons n=%d", n_photons);
    lwb_handle h = c->h[0];
    CU(cudaSetDevice(h->device));
    const uint64_t S = binom(host_bin(), m + n_photons - 1, n_photons);
    CK(comm_cuts(c, m, n_photons));
    CK(ensure_buffers(c, S * 8));
    CK(shard_launch(c, c->rank, h, d_U, m, d_in_state, n_photons, dtype, 1, c->d_sum[0]));
    if (!(c->peer_ok && c->use_peer_stores) && c->world > 1) {
        // no peer mapping: the shards travel as in-place broadcasts of unequal length, one group call
        double* full = c->full[c->rank];
        NC(g_nccl.GroupStart());
        for (int g = 0; g < c->world; ++g) {
            const uint64_t cnt = c->cuts[g + 1] - c->cuts[g];
            if (cnt) NC(g_nccl.Broadcast(full + c->cuts[g], full + c->cuts[g], cnt, NCCL_FLOAT64, g, c->nccl, h->stream));
        }
        NC(g_nccl.GroupEnd());
    }
    // sum of the shard sums = normalisation check (permanent.py:159) AND the barrier after which every rank's copy
    // of the distribution is complete (a rank contributes only after its kernels, hence its peer stores, finished)
    if (c->world > 1) NC(g_nccl.AllReduce(c->d_sum[0], c->d_sum[0], 1, NCCL_FLOAT64, NCCL_SUM, c->nccl, h->stream));
    if (d_probs) *d_probs = c->full[c->rank];
    if (d_sum) *d_sum = c->d_sum[0];
    return 0;
}

// Host-buffer entry point for both kinds.  probs (optional): LOCAL -> the whole distribution (S doubles);
// RANK -> the whole distribution when gather_host != 0, else only this rank's shard is copied, to probs + r0.
int lwb_comm_perm_basis_probs(lwb_comm c, const double* U, int m, const uint8_t* in_state, int dtype, int gather_host,
                              double* probs, double* sum_out) {
    if (!c || !U || !in_state) return fail(LWB_E_ARG, "null argument");
    int n = 0;
    CK(comm_check_state(m, in_state, &n));
    const uint64_t S = binom(host_bin(), m + n - 1, n);
    if (c->kind == COMM_RANK) {
        lwb_handle h = c->h[0];
        {
            std::lock_guard<std::mutex> lk(h->mu);
            CU(cudaSetDevice(h->device));
            CK(ensure(h, h->bU, (size_t)m * m * 16));
            CK(ensure(h, h->bIn, (size_t)m));
            CU(cudaMemcpyAsync(h->bU.p, U, (size_t)m * m * 16, cudaMemcpyHostToDevice, h->stream));
            CU(cudaMemcpyAsync(h->bIn.p, in_state, (size_t)m, cudaMemcpyHostToDevice, h->stream));
            double *d_p = nullptr, *d_s = nullptr;
            CK(lwb_comm_perm_basis_probs_dev(c, (const double*)h->bU.p, m, (const uint8_t*)h->bIn.p, n, dtype, &d_p, &d_s));
            if (probs) {
                if (gather_host) CU(cudaMemcpyAsync(probs, d_p, S * 8, cudaMemcpyDeviceToHost, h->stream));
                else {
                    const uint64_t r0 = c->cuts[c->rank], cnt = c->cuts[c->rank + 1] - r0;
                    if (cnt) CU(cudaMemcpyAsync(probs + r0, d_p + r0, cnt * 8, cudaMemcpyDeviceToHost, h->stream));
                }
            }
            double s = 0.0;
            CU(cudaMemcpyAsync(&s, d_s, 8, cudaMemcpyDeviceToHost, h->stream));
            CU(cudaStreamSynchronize(h->stream));
            if (sum_out) *sum_out = s;
        }
        return 0;
    }
    // LOCAL: one host thread, every launch asynchronous; device 0 collects
    std::lock_guard<std::mutex> lk(c->mu);
    CK(comm_cuts(c, m, n));
    CK(ensure_buffers(c, S * 8));
    const bool peer = c->peer_ok && c->use_peer_stores;
    for (int g = 0; g < c->world; ++g) {
        lwb_handle h = c->h[g];
        CU(cudaSetDevice(h->device));
        CK(ensure(h, h->bU, (size_t)m * m * 16));
        CK(ensure(h, h->bIn, (size_t)m));
        CU(cudaMemcpyAsync(h->bU.p, U, (size_t)m * m * 16, cudaMemcpyHostToDevice, h->stream));
        CU(cudaMemcpyAsync(h->bIn.p, in_state, (size_t)m, cudaMemcpyHostToDevice, h->stream));
        CK(shard_launch(c, g, h, (const double*)h->bU.p, m, (const uint8_t*)h->bIn.p, n, dtype, 0, c->d_sum[g]));
        const uint64_t r0 = c->cuts[g], cnt = c->cuts[g + 1] - r0;
        if (!peer && g != 0 && cnt)  // no peer stores: copy the finished shard to device 0
            CU(cudaMemcpyPeerAsync(c->full[0] + r0, c->h[0]->device, c->full[g] + r0, h->device, cnt * 8, h->stream));
    }
    double total = 0.0;
    for (int g = 0; g < c->world; ++g) {
        lwb_handle h = c->h[g];
        CU(cudaSetDevice(h->device));
        double s = 0.0;
        CU(cudaMemcpyAsync(&s, c->d_sum[g], 8, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        total += s;
    }
    if (sum_out) *sum_out = total;
    if (probs) {
        lwb_handle h0 = c->h[0];
        CU(cudaSetDevice(h0->device));
        CU(cudaMemcpyAsync(probs, c->full[0], S * 8, cudaMemcpyDeviceToHost, h0->stream));
        CU(cudaStreamSynchronize(h0->stream));
    }
    return 0;
}

// full_probability_distribution (permanent.py:116-163) with the permanents sharded over the communicator; the
// threshold / loss-mode marginalisation / dict ordering run on this rank's (LOCAL: device 0's) complete copy.
int lwb_comm_pdist(lwb_comm c, const double* U_full, int m_tot, int n_modes, const uint8_t* in_state, double threshold,
                   int dtype, uint8_t* keys, double* vals, uint64_t cap, uint64_t* count_out) {
    if (!c || !U_full || !in_state || !keys || !vals || !count_out) return fail(LWB_E_ARG, "null argument");
    if (n_modes < 1 || n_modes > m_tot || m_tot > LWB_MAX_MODES) return fail(LWB_E_ARG, "n_modes=%d of %d", n_modes, m_tot);
    int n = 0;
    for (int i = 0; i < n_modes; ++i) n += in_state[i];
    lwb_handle h = c->h[0];
    if (n == 0 || n > LWB_K1_MAX_N)  // vacuum shortcut and size errors: the single-device entry point answers
        return lwb_pdist(h, U_full, m_tot, n_modes, in_state, threshold, LWB_BACKEND_PERMANENT, dtype, keys, vals, cap, count_out);
    uint8_t full_in[LWB_MAX_MODES];
    memset(full_in, 0, sizeof full_in);
    memcpy(full_in, in_state, (size_t)n_modes);  // permanent.py:142-143
    CK(lwb_comm_perm_basis_probs(c, U_full, m_tot, full_in, dtype, 0, nullptr, nullptr));
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    const double* d_probs = c->full[c->kind == COMM_LOCAL ? 0 : c->rank];
    return pdist_finish(h, d_probs, m_tot, n_modes, n, threshold, LWB_BACKEND_PERMANENT, keys, vals, cap, count_out);
}

// Below this size one GPU finishes the permanent sooner than the 16-byte all-reduce returns (measured on 8 B200s:
// n = 22 39 us on one GPU vs 46 us sliced, n = 24 104 vs 84 us): every rank then walks the whole Gray space itself -
// same kernel, same result on every rank, no exchange.
static const int kCommSingleMinN = 24;

// One large permanent: the Gray-code space in `world` slices, partial sums added (16 bytes).
int lwb_comm_perm_single_dev(lwb_comm c, const double* d_A, int n, int dtype, double* d_out2) {
    if (!c || !d_A || !d_out2) return fail(LWB_E_ARG, "null argument");
    if (c->kind != COMM_RANK) return fail(LWB_E_ARG, "device-pointer entry point needs a per-rank communicator");
    lwb_handle h = c->h[0];
    if (n < kCommSingleMinN || c->world == 1) return lwb_perm_single_dev(h, d_A, n, dtype, 0, 1, d_out2);
    CK(lwb_perm_single_dev(h, d_A, n, dtype, c->rank, c->world, d_out2));
    NC(g_nccl.AllReduce(d_out2, d_out2, 2, NCCL_FLOAT64, NCCL_SUM, c->nccl, h->stream));
    return 0;
}

int lwb_comm_perm_single(lwb_comm c, const double* A, int n, int dtype, double* out2) {
    if (!c || !A || !out2) return fail(LWB_E_ARG, "null argument");
    if (n < 1 || n > LWB_K2_MAX_N) return fail(LWB_E_LIMIT, "single permanent supports 1 <= n <= %d (got %d)", LWB_K2_MAX_N, n);
    if (c->kind == COMM_RANK) {
        lwb_handle h = c->h[0];
        std::lock_guard<std::mutex> lk(h->mu);
        CU(cudaSetDevice(h->device));
        CK(ensure(h, h->bU, (size_t)n * n * 16));
        CU(cudaMemcpyAsync(h->bU.p, A, (size_t)n * n * 16, cudaMemcpyHostToDevice, h->stream));
        CK(lwb_comm_perm_single_dev(c, (const double*)h->bU.p, n, dtype, c->d_sum[0] + 1));
        CU(cudaMemcpyAsync(out2, c->d_sum[0] + 1, 16, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        return 0;
    }
    std::lock_guard<std::mutex> lk(c->mu);
    const int parts = n < kCommSingleMinN ? 1 : c->world;  // small permanents: device 0 alone
    for (int g = 0; g < parts; ++g) {
        lwb_handle h = c->h[g];
        CU(cudaSetDevice(h->device));
        CK(ensure(h, h->bU, (size_t)n * n * 16));
        CU(cudaMemcpyAsync(h->bU.p, A, (size_t)n * n * 16, cudaMemcpyHostToDevice, h->stream));
        CK(lwb_perm_single_dev(h, (const double*)h->bU.p, n, dtype, g, parts, c->d_sum[g] + 1));
    }
    double re = 0.0, im = 0.0;
    for (int g = 0; g < parts; ++g) {  // fixed order: deterministic
        lwb_handle h = c->h[g];
        CU(cudaSetDevice(h->device));
        double part[2];
        CU(cudaMemcpyAsync(part, c->d_sum[g] + 1, 16, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        re += part[0];
        im += part[1];
    }
    out2[0] = re;
    out2[1] = im;
    return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// The unique sub-inputs of an imperfect source (probability_distribution.py:127-133) are independent
// full_probability_distribution calls: input k runs on device k mod n_gpus, all of them asynchronously, and the
// dense result (a few MB) is copied over NVLink into a slot of `h`, where the convolutions run.
extern "C" int lwb_dists_from_circuit(lwb_handle h, lwb_comm comm, const double* U_full, int m_tot, int n_modes,
                                      const uint8_t* in_states, int count, double threshold, int backend, int dtype,
                                      int* slots) {
    if (!h || !U_full || !in_states || !slots || count < 0) return fail(LWB_E_ARG, "null argument");
    if (backend != LWB_BACKEND_PERMANENT && backend != LWB_BACKEND_SLOS) return fail(LWB_E_ARG, "backend %d", backend);
    if (n_modes < 1 || n_modes > m_tot) return fail(LWB_E_ARG, "n_modes=%d outside [1,%d]", n_modes, m_tot);
    if (m_tot > LWB_MAX_MODES) return fail(LWB_E_LIMIT, "modes m=%d outside [1,%d]", m_tot, LWB_MAX_MODES);
    if (comm && comm->kind != COMM_LOCAL) return fail(LWB_E_ARG, "sub-input round-robin needs a single-process communicator");
    std::lock_guard<std::mutex> lk(h->mu);
    std::vector<lwb_handle> workers{h};
    if (comm)
        for (int g = 0; g < comm->world; ++g)
            if (comm->h[g]->device != h->device) workers.push_back(comm->h[g]);
    const int nw = (int)workers.size();
    const int loss = m_tot - n_modes;
    int nmax = 0;
    for (int k = 0; k < count; ++k) {
        int n = 0;
        for (int i = 0; i < n_modes; ++i) n += in_states[(size_t)k * n_modes + i];
        nmax = std::max(nmax, n);
    }
    CK(check_mn(m_tot, nmax));
    uint64_t kmax_size = 0;
    CK(lwb_keyspace_size(n_modes, nmax, &kmax_size));
    for (int g = 0; g < nw; ++g) {
        lwb_handle w = workers[g];
        CU(cudaSetDevice(w->device));
        CK(ensure(w, w->bU, (size_t)m_tot * m_tot * 16));
        CU(cudaMemcpyAsync(w->bU.p, U_full, (size_t)m_tot * m_tot * 16, cudaMemcpyHostToDevice, w->stream));
        if (g > 0) CK(ensure(w, w->bRes, kmax_size * 8));
    }
    if (nw > 1) {  // other devices' streams will write into slots that may be recycled buffers: drain h's pending readers
        CU(cudaSetDevice(h->device));
        CU(cudaStreamSynchronize(h->stream));
    }
    int made = 0, rc = 0;
    for (int k = 0; k < count && rc == 0; ++k) {
        const uint8_t* ins = in_states + (size_t)k * n_modes;
        int n = 0;
        for (int i = 0; i < n_modes; ++i) n += ins[i];
        CU(cudaSetDevice(h->device));
        rc = new_dist_slot(h, n_modes, loss ? 0 : n, n, &slots[k]);
        if (rc) break;
        ++made;
        const auto d = h->dists[slots[k]];
        lwb_handle w = workers[k % nw];
        if (w == h) {
            rc = dist_from_circuit_enqueue(h, (const double*)h->bU.p, m_tot, n_modes, ins, threshold, backend, dtype, d.p, d.size);
        } else {
            cudaSetDevice(w->device);
            rc = dist_from_circuit_enqueue(w, (const double*)w->bU.p, m_tot, n_modes, ins, threshold, backend, dtype,
                                           (double*)w->bRes.p, d.size);
            if (!rc && cudaMemcpyPeerAsync(d.p, h->device, w->bRes.p, w->device, d.size * 8, w->stream) != cudaSuccess)
                rc = fail(LWB_E_CUDA, "cudaMemcpyPeerAsync: %s", cudaGetErrorString(cudaGetLastError()));
        }
    }
    for (int g = 0; g < nw; ++g) {  // U_full / in_states are the caller's: everything has to be done before returning
        cudaSetDevice(workers[g]->device);
        cudaError_t e = cudaStreamSynchronize(workers[g]->stream);
        if (e != cudaSuccess && rc == 0) rc = fail(LWB_E_CUDA, "cudaStreamSynchronize: %s", cudaGetErrorString(e));
    }
    cudaSetDevice(h->device);
    if (rc) {  // leave no slot behind on failure
        const std::string msg = lwb_last_error();
        for (int k = 0; k < made; ++k) {
            auto& d = h->dists[slots[k]];
            if (d.p) cudaFree(d.p);
            d = lwb_context::DistSlot();
            slots[k] = -1;
        }
        return fail(rc, "%s", msg.c_str());
    }
    return 0;
}
