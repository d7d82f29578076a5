// instantiation unit: fused multiplicity-aware kernels (generic + static walks) for n = 16
#include "k1x_common.cuh"
namespace lwb {
static const K1xStaticEntry table[] = {
    LWB_K1XS_ENTRY(16),
};
const K1xStaticEntry* k1xs_g3_table(int* count) { *count = (int)(sizeof(table) / sizeof(table[0])); return table; }
}  // namespace lwb
