// lwb_internal.h -- private declarations shared by the units behind the C ABI (lwb_api.cu, comm.cu).
#pragma once
#include <cuda_runtime.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/lwb200.h"
#include "fock.cuh"
#include "k1x.h"
#include "slos_units.h"

namespace lwb {
int fail(int code, const char* fmt, ...);  // sets the calling thread's lwb_last_error() text, returns code
}
using lwb::fail;

#define CU(call)                                                                                 \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess) return fail(LWB_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define CK(expr)                 \
    do {                         \
        int rc_ = (expr);        \
        if (rc_ != 0) return rc_; \
    } while (0)

#ifndef LWB_K1_MAX_N
#define LWB_K1_MAX_N 24  /* batched / full-basis permanents: photons */
#define LWB_K2_MAX_N 40  /* single large permanent */
#endif

struct Buf {
    void* p = nullptr;
    size_t cap = 0;
};

struct lwb_context {
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // D2H of finished chunks overlaps the next chunk's kernel
    cudaEvent_t chunk_done[8] = {};
    uint64_t* d_bin = nullptr;
    Buf bU, bIn, bOut, bRes, bPart, bTmp, bLayerA, bLayerB, bProbs, bMarg, bFirst, bSmall, bCdf, bTile, bUni, bIdx, bTerms;
    uint64_t launches = 0;
    lwb::K1xState* k1x = nullptr;
    lwb::SlosUnitsState* slos_units = nullptr;  // cached K3u plans per (m, k)
    int batch_tasks = 0, batch_shared_inputs = 0;  // set around lwb_perm_amplitudes_batch's launch (blockIdx.z = task)
    lwb::MirrorSet mirror = {};  // set by the communicator around a sharded launch: peers receiving every probability
    struct DistSlot {  // dense distribution over the key space K(n_modes, kmax), device resident
        double* p = nullptr;
        int n_modes = 0, kmin = 0, kmax = 0;
        uint64_t size = 0;
    };
    std::vector<DistSlot> dists;
    // freed distribution buffers by size in doubles: lwb_dist_free parks them here instead of cudaFree (which
    // synchronises the device) and new slots reuse them - everything is ordered on the handle's stream
    std::vector<std::pair<uint64_t, double*>> dist_pool;
    uint64_t dist_pool_bytes = 0;
    std::mutex mu;
};

namespace lwb {
const uint64_t* host_bin();
int ensure(lwb_handle h, Buf& b, size_t bytes);
int check_mn(int m, int n);
// probabilities / amplitudes of n_in inputs x (explicit outputs | rank range [r0, r0 + n_out) of fock_basis) on h->stream
int launch_states(lwb_handle h, const double* d_U, int m, const uint8_t* d_ins, int n_in, const uint8_t* d_outs,
                  uint64_t r0, uint64_t n_out, int n, int dtype, int want_probs, double* d_out);
int pdist_finish(lwb_handle h, const double* d_probs, int m_tot, int n_modes, int n, double threshold, int backend,
                 uint8_t* keys, double* vals, uint64_t cap, uint64_t* count_out);
// lwb_dist_from_circuit without slot bookkeeping or synchronisation: zero-fills d_dst (the key space K(n_modes, n))
// and enqueues the whole chain on h->stream.  d_U_full is device memory on h's device.
int dist_from_circuit_enqueue(lwb_handle h, const double* d_U_full, int m_tot, int n_modes, const uint8_t* in_state,
                              double threshold, int backend, int dtype, double* d_dst, uint64_t dst_size);
int new_dist_slot(lwb_handle h, int n_modes, int kmin, int kmax, int* slot);
int sum_doubles_dev(lwb_handle h, const double* d_x, uint64_t count, double* d_out);  // deterministic tree sum
}  // namespace lwb
