// k1x.h -- multiplicity-aware batched permanents (K1x), host interface used by lwb_api.cu.
//
// For an output state that occupies d distinct modes r_1..r_d with multiplicities s_1..s_d the n x n matrix of
// `partition` (lightworks/emulator/backends/permanent.py:166-178) has repeated rows.  Glynn's sum over sign
// vectors then only depends on how many copies v_i of each distinct row carry a minus sign:
//
//   perm = sum_{v, 0<=v_i<=s_i} (-1)^{sum v} prod_i C(s_i, v_i) prod_j w_j(v),   w_j(v) = 1/2 sum_i (s_i - 2 v_i) a_{r_i j}
//
// i.e. prod_i (s_i + 1) terms instead of 2^n, halved by the v -> s - v symmetry whenever some s_i is odd.  The
// tuples v are walked in reflected mixed-radix Gray order (one row update per term, as in the binary case).
// States are bucketed by their multiplicity pattern ("type", a partition of n) so that every thread of a block
// runs the same pre-computed walk ("program": row slot, direction and integer coefficient per term) on its own
// permanent; rows are ordered by decreasing multiplicity, which the permanent does not care about.
// Bunched states dominate full distributions (24 modes / 10 photons: 233 terms on average instead of 512).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "fock.cuh"

namespace lwb {

struct K1xState;  // cached plans + scratch buffers (owned by the context)
K1xState* k1x_create();
void k1x_destroy(K1xState* st);
bool k1x_supported(int n);
bool k1x_supported_dtype(int n, int dtype);  // dtype 0 complex128, 1 complex64 (n <= 12: drift of the float column sums)
int k1x_plan_info(int n, int t, int* n_types, int* s_out, int* d_out, int* T_out, int* factor_out, double* abs_coef_sum);

// probs[k] = |<basis[r0+k]|U|in>|^2 for k in [0, cnt), all pointers on the device; enqueues five kernels on
// `stream` and returns without synchronising (bucket and block layout are computed on the device).
// Returns 0, or a negative LWB_E_* code with *err set.
int k1x_basis_probs(K1xState* st, cudaStream_t stream, const uint64_t* d_bin, const double2* d_U, int m,
                    const uint8_t* d_in_state, int n, uint64_t r0, uint64_t cnt, double* d_probs, uint64_t* launches,
                    std::string* err, const MirrorSet* mirrors = nullptr, int dtype = 0);

}  // namespace lwb
