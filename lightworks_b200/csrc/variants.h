// variants.h -- kernel variant tables shared between the instantiation units and lwb_api.cu
#pragma once
#include "perm_kernels.cuh"

#ifndef LWB_K1_MAX_N
#define LWB_K1_MAX_N 24
#define LWB_K2_MAX_N 40
#endif

namespace lwb {

struct K1Variant {
    int n, log_l;
    void (*states[2])(PermStatesArgs);      // [dtype]
    void (*matrices[2])(PermMatricesArgs);  // [dtype]
};
#define LWB_K1_ENTRY(N, LL)                                                             \
    {N, LL, {perm_states_kernel<N, LL, double>, perm_states_kernel<N, LL, float>},      \
     {perm_matrices_kernel<N, LL, double>, perm_matrices_kernel<N, LL, float>}}

struct K2Variant {
    int n, chunk_bits;
    void (*single[2])(PermSingleArgs);  // [dtype]
};
#define LWB_K2_ENTRY(N) {N, K2Shape<N>::C, {perm_single_kernel<N, double>, perm_single_kernel<N, float>}}

// The first K1 entry found for an N (in group order) is its default variant.
const K1Variant* k1_g0_table(int* count);
const K1Variant* k1_g1_table(int* count);
const K1Variant* k1_g2_table(int* count);
const K1Variant* k1_g3_table(int* count);
const K1Variant* k1_g4_table(int* count);
const K1Variant* k1_g5_table(int* count);
const K1Variant* k1_g6_table(int* count);
const K1Variant* k1_g7_table(int* count);
const K2Variant* k2_g0_table(int* count);
const K2Variant* k2_g1_table(int* count);
const K2Variant* k2_g2_table(int* count);
const K2Variant* k2_g3_table(int* count);
#define LWB_K1_GROUPS {k1_g0_table, k1_g1_table, k1_g2_table, k1_g3_table, k1_g4_table, k1_g5_table, k1_g6_table, k1_g7_table}
#define LWB_K2_GROUPS {k2_g0_table, k2_g1_table, k2_g2_table, k2_g3_table}

}  // namespace lwb
