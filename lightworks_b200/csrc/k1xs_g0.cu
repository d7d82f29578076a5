// instantiation unit: fused multiplicity-aware kernels (generic + static walks) for n = 7, 8, 9, 10
#include "k1x_common.cuh"
namespace lwb {
static const K1xStaticEntry table[] = {
    LWB_K1XS_ENTRY(7),
    LWB_K1XS_ENTRY(8),
    LWB_K1XS_ENTRY(9),
    LWB_K1XS_ENTRY(10),
};
const K1xStaticEntry* k1xs_g0_table(int* count) { *count = (int)(sizeof(table) / sizeof(table[0])); return table; }
}  // namespace lwb
