// Fused on-chip kernels (wc_fused.cu): one unit's coefficients stay in shared memory between the
// transform, the arg-max, the mask and the packing, so HBM sees the input once and the pairs once.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "wc_kernels.h"

namespace wc {

enum { FUSED_FULL = 0, FUSED_KEYS_ONLY = 1, FUSED_GIVEN_THRESH = 2 };

// Which fused kernel handles a box: 0 = none (generic path), 1 = one CTA per unit,
// 8 = one 8-CTA cluster per unit.
int  fused_class(int nx, int ny, int nz, int dtype, const void* device_ptr);
bool fused_decode_available();
int  fused_decode_class(int nx, int ny, int nz, int out_dtype, const void* out_device_ptr);

cudaError_t launch_fused_compress(int cluster, int mode, const UnitDev* units, UnitState* states,
                                  const int* unit_list, int n_list, double one_minus_keep,
                                  const u64* global_key, int sm_count, cudaStream_t st,
                                  LaunchStats* ls, int* work_counter = nullptr);
cudaError_t launch_fused_decompress(int cluster, const DecUnitDev* dec, const InvUnitDev* inv,
                                    const int* unit_list, int n_list, int* err, int sm_count,
                                    cudaStream_t st, LaunchStats* ls, int* work_counter = nullptr);

cudaError_t debug_phase_cycles(unsigned long long out[6], bool reset);

} // namespace wc
