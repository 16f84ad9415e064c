// placeholder until the fused kernels land: everything goes through the generic path
#include "wc_common.cuh"
#include "wc_fused.h"

namespace wc {
int  fused_class(int, int, int) { return 0; }
bool fused_decode_available() { return false; }
cudaError_t launch_fused_compress(int, int, const UnitDev*, UnitState*, const int*, int, double,
                                  const u64*, int, cudaStream_t, LaunchStats*) { return cudaSuccess; }
cudaError_t launch_fused_decompress(int, const DecUnitDev*, const InvUnitDev*, const int*, int, int*,
                                    int, cudaStream_t, LaunchStats*) { return cudaSuccess; }
} // namespace wc
