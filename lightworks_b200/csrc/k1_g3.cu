// generated instantiation unit: K1 variants (N, LOG_L) compiled in this translation unit
// (first entry per N = default; the rest are selectable with lwb_set_tuning for sweeps)
#include "variants.h"
#include "perm_kernels.cuh"
namespace lwb {
static const K1Variant table[] = {
    LWB_K1_ENTRY(13, 3),
    LWB_K1_ENTRY(14, 3),
    LWB_K1_ENTRY(13, 4),
    LWB_K1_ENTRY(14, 4),
};
const K1Variant* k1_g3_table(int* count) { *count = (int)(sizeof(table) / sizeof(table[0])); return table; }
}  // namespace lwb
