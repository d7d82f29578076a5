// generated instantiation unit: K2 (single large permanent) for a range of N
#include "variants.h"
#include "perm_kernels.cuh"
namespace lwb {
static const K2Variant table[] = {
    LWB_K2_ENTRY(1),
    LWB_K2_ENTRY(2),
    LWB_K2_ENTRY(3),
    LWB_K2_ENTRY(4),
    LWB_K2_ENTRY(5),
    LWB_K2_ENTRY(6),
    LWB_K2_ENTRY(7),
    LWB_K2_ENTRY(8),
    LWB_K2_ENTRY(9),
    LWB_K2_ENTRY(10),
    LWB_K2_ENTRY(11),
    LWB_K2_ENTRY(12),
    LWB_K2_ENTRY(13),
    LWB_K2_ENTRY(14),
    LWB_K2_ENTRY(15),
    LWB_K2_ENTRY(16),
    LWB_K2_ENTRY(17),
    LWB_K2_ENTRY(18),
    LWB_K2_ENTRY(19),
    LWB_K2_ENTRY(20),
};
const K2Variant* k2_g0_table(int* count) { *count = (int)(sizeof(table) / sizeof(table[0])); return table; }
}  // namespace lwb
