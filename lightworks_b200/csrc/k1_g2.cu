// generated instantiation unit: K1 variants (N, LOG_L) compiled in this translation unit
// (first entry per N = default; the rest are selectable with lwb_set_tuning for sweeps)
#include "variants.h"
#include "perm_kernels.cuh"
namespace lwb {
static const K1Variant table[] = {
    LWB_K1_ENTRY(11, 2),
    LWB_K1_ENTRY(12, 2),
    LWB_K1_ENTRY(11, 3),
    LWB_K1_ENTRY(12, 3),
};
const K1Variant* k1_g2_table(int* count) { *count = (int)(sizeof(table) / sizeof(table[0])); return table; }
}  // namespace lwb
