// generated instantiation unit: K2 (single large permanent) for a range of N
#include "variants.h"
#include "perm_kernels.cuh"
namespace lwb {
static const K2Variant table[] = {
    LWB_K2_ENTRY(35),
    LWB_K2_ENTRY(36),
    LWB_K2_ENTRY(37),
    LWB_K2_ENTRY(38),
    LWB_K2_ENTRY(39),
    LWB_K2_ENTRY(40),
};
const K2Variant* k2_g3_table(int* count) { *count = (int)(sizeof(table) / sizeof(table[0])); return table; }
}  // namespace lwb
