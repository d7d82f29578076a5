// generated instantiation unit: K1 variants (N, LOG_L) compiled in this translation unit
// (first entry per N = default; the rest are selectable with lwb_set_tuning for sweeps)
#include "variants.h"
#include "perm_kernels.cuh"
namespace lwb {
static const K1Variant table[] = {
    LWB_K1_ENTRY(1, 0),
    LWB_K1_ENTRY(2, 0),
    LWB_K1_ENTRY(3, 0),
    LWB_K1_ENTRY(4, 0),
    LWB_K1_ENTRY(5, 0),
    LWB_K1_ENTRY(6, 0),
    LWB_K1_ENTRY(7, 0),
    LWB_K1_ENTRY(8, 1),
    LWB_K1_ENTRY(6, 1),
    LWB_K1_ENTRY(7, 1),
    LWB_K1_ENTRY(8, 2),
};
const K1Variant* k1_g0_table(int* count) { *count = (int)(sizeof(table) / sizeof(table[0])); return table; }
}  // namespace lwb
