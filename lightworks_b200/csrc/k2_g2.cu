// generated instantiation unit: K2 (single large permanent) for a range of N
#include "variants.h"
#include "perm_kernels.cuh"
namespace lwb {
static const K2Variant table[] = {
    LWB_K2_ENTRY(29),
    LWB_K2_ENTRY(30),
    LWB_K2_ENTRY(31),
    LWB_K2_ENTRY(32),
    LWB_K2_ENTRY(33),
    LWB_K2_ENTRY(34),
};
const K2Variant* k2_g2_table(int* count) { *count = (int)(sizeof(table) / sizeof(table[0])); return table; }
}  // namespace lwb
