// generated instantiation unit: K2 (single large permanent) for a range of N
#include "variants.h"
#include "perm_kernels.cuh"
namespace lwb {
static const K2Variant table[] = {
    LWB_K2_ENTRY(21),
    LWB_K2_ENTRY(22),
    LWB_K2_ENTRY(23),
    LWB_K2_ENTRY(24),
    LWB_K2_ENTRY(25),
    LWB_K2_ENTRY(26),
    LWB_K2_ENTRY(27),
    LWB_K2_ENTRY(28),
};
const K2Variant* k2_g1_table(int* count) { *count = (int)(sizeof(table) / sizeof(table[0])); return table; }
}  // namespace lwb
