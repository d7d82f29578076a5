// instantiation unit: fused multiplicity-aware kernels (generic + static walks) for n = 11, 12, 13
#include "k1x_common.cuh"
namespace lwb {
static const K1xStaticEntry table[] = {
    LWB_K1XS_ENTRY(11),
    LWB_K1XS_ENTRY(12),
    LWB_K1XS_ENTRY(13),
};
const K1xStaticEntry* k1xs_g1_table(int* count) { *count = (int)(sizeof(table) / sizeof(table[0])); return table; }
}  // namespace lwb
