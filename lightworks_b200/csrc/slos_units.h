// slos_units.h -- K3u host interface (slos_units.cu): SLOS layer kernel over prefix units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace lwb {
struct SlosUnitsState;  // cached per-(m, k) plans, owned by the context
SlosUnitsState* slos_units_create();
void slos_units_destroy(SlosUnitsState* st);
bool slos_units_supported(int m, int k);
// One layer (k photons in the child layer) on `stream`; all pointers on the device.  Returns 0 or a negative LWB_E_*
// code with *err set (LWB_E_LIMIT: fall back to the gather-form kernel).
int slos_units_layer(SlosUnitsState* st, cudaStream_t stream, int m, int k, const double2* parent, void* child,
                     const double2* U, int in_mode, const double* sqrt_tab, int write_probs, uint64_t* launches,
                     std::string* err, int version = 3);  // version 4: record kernel (host-built records)
int slos_units_plan_check(int m, int k, uint64_t* children_checked);
int slos_units_plan_info(int m, int k, uint64_t* n_units, uint64_t* n_items, uint64_t* children, uint64_t* max_tp);
}  // namespace lwb
