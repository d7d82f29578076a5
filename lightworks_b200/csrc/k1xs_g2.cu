// instantiation unit: fused multiplicity-aware kernels (generic + static walks) for n = 14, 15
#include "k1x_common.cuh"
namespace lwb {
static const K1xStaticEntry table[] = {
    LWB_K1XS_ENTRY(14),
    LWB_K1XS_ENTRY(15),
};
const K1xStaticEntry* k1xs_g2_table(int* count) { *count = (int)(sizeof(table) / sizeof(table[0])); return table; }
}  // namespace lwb
