// Fused on-chip kernels (wc_fused.cu): one unit's coefficients stay in shared memory between the
// transform, the arg-max, the mask and the packing, so HBM sees the input once and the pairs once.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "wc_kernels.h"

namespace wc {

enum { FUSED_FULL = 0, FUSED_KEYS_ONLY = 1, FUSED_GIVEN_THRESH = 2,
       FUSED_MINMAX = 16 };   // flag, or-ed in: also record per-unit min / max of the narrowed inputs

// Which fused compress kernel handles a box: 0 = none (generic path); one CTA per unit (<= 32768 cells) or
// one 8-CTA cluster per unit (<= 262144 cells), each as a runtime-geometry kernel (512 threads) and as a
// literal-geometry kernel for the most common AMR box (1024 threads: the literal strides and trip counts
// bring it under 64 registers).
// Units of at most 4096 cells (16^3 and smaller) go to variants with a 16 KB coefficient array, so that
// four CTAs share an SM and overlap each other's barriers and load latencies; 8^3 boxes get one-warp CTAs,
// 32 per SM.
enum { FUSED_CLS_NONE = 0, FUSED_CLS_R1 = 1, FUSED_CLS_R1S = 2, FUSED_CLS_R8 = 8, FUSED_CLS_R4 = 24,
       FUSED_CLS_R2 = 22,      // clusters of 4 / 2: boxes whose half-height is not a multiple of 8 (40^3 ...)
       FUSED_CLS_CUBE32 = 101,
       FUSED_CLS_CUBE64 = 108, FUSED_CLS_CUBE16 = 116, FUSED_CLS_CUBE8 = 117,
       FUSED_CLS_RBIG = 200,    // decompress only: any number of y-slabs of <= 32768 cells (128^3 ...), one launch per slab count
       // x-slab classes (wc_xslab.cu): ANY shape the y-slab classes refuse (odd dimensions, nz % 4 != 0, rows that are not
       // 16-byte multiples, unaligned pointers) whose x-slab of 1 / 2 / 4 / 8 fits a CTA; compress = one cluster per unit
       FUSED_CLS_XS1 = 301, FUSED_CLS_XS2 = 302, FUSED_CLS_XS4 = 304, FUSED_CLS_XS8 = 308,
       FUSED_CLS_XS1S = 300 };  // one CTA, at most 8448 words of C: 128 threads, five CTAs per SM
int  fused_class(int nx, int ny, int nz, int dtype, const void* device_ptr);
int  fused_decode_class(int nx, int ny, int nz, int out_dtype, const void* out_device_ptr);
size_t fused_decode_table_entries(int fused_cls, int nx, int ny, int nz);   // int2 entries of a unit's segment table
int fused_decode_slabs_of(int fused_cls, int nx, int ny, int nz);         // y-slabs (work items) per unit
// Launch key of a FUSED_CLS_RBIG unit: its slab count, plus BIG_CUBE128 for the 128^3 cube (which has a literal-geometry
// decode kernel of its own).  One launch takes units of one key; the launchers take the key as their s_rt argument.
enum { BIG_SLAB_MASK = 0xfffff, BIG_CUBE128 = 1 << 20 };
int big_run_key(int nx, int ny, int nz);
bool fused_decode_needs_table(int fused_cls);               // slab-decoded classes cannot decode without one

// x-slab classes (wc_xslab.cu)
int xs_slabs(int nx, int ny, int nz);                       // 1 / 2 / 4 / 8 x-slabs, 0 = the box does not fit
int xs_class_of(int nx, int ny, int nz);                    // FUSED_CLS_XS* or FUSED_CLS_NONE
int xs_class_slabs(int fused_cls);                          // slabs of an x-slab class, 0 for every other class
cudaError_t launch_xs_compress(int fused_cls, int mode, const UnitDev* units, UnitState* states, const int* unit_list,
                               int n_list, double one_minus_keep, const u64* global_key, int sm_count, cudaStream_t st,
                               LaunchStats* ls);
cudaError_t launch_xs_decompress(int fused_cls, const DecUnitDev* dec, const InvUnitDev* inv, const int* unit_list,
                                 int n_list, int* err, int sm_count, cudaStream_t st, LaunchStats* ls, int* work_counter);

cudaError_t launch_fused_compress(int fused_cls, int mode, const UnitDev* units, UnitState* states,
                                  const int* unit_list, int n_list, double one_minus_keep,
                                  const u64* global_key, int sm_count, cudaStream_t st,
                                  LaunchStats* ls, int* work_counter = nullptr);
cudaError_t launch_fused_decompress(int fused_cls, const DecUnitDev* dec, const InvUnitDev* inv,
                                    const int* unit_list, int n_list, int* err, int sm_count,
                                    cudaStream_t st, LaunchStats* ls, int* work_counter = nullptr,
                                    bool build_tables = true, int stage = 0, int s_rt = 0);
// stage (32^3 cubes): 0 = pairs straight from global memory, 1 = table-less units decode from a shared-memory staging
// area filled by TMA bulk copies one item ahead (k_staged_decompress), 2 = units with a segment table as well

// Big boxes (no cluster holds them): y-slab forward transform into the unit's coefficient scratch + arg-max key;
// threshold and packing then run on the scratch like for every generic unit.  One launch per slab count.
int big_forward_slabs(int nx, int ny, int nz, int dtype, const void* in_device_ptr);
cudaError_t launch_big_forward(const UnitDev* units, UnitState* states, const int* unit_list, int n_list, int s_rt,
                               int* work_counter, int sm_count, cudaStream_t st, LaunchStats* ls);

// Ordered packing of generic / big-box units from their coefficient scratch in ONE pass (k_big_pack: TMA bulk loads,
// chunk position by decoupled look-back).  items[i] = (unit, chunk) for every chunk of big_pack_chunks(n) per unit, unit
// after unit; status: one word per item (zeroed by the launcher).
int big_pack_chunks(long long ncoef);
cudaError_t launch_big_pack(const UnitDev* units, UnitState* states, const int2* items, int n_items, u64* status,
                            int* work_counter, int sm_count, cudaStream_t st, LaunchStats* ls);

// Streamed segment index for packed streams without tables (see k_seg_index3), and the device-side preparation
// of a dense stream (k_dec_prepare).
int fused_decode_slabs(int fused_cls);
cudaError_t launch_seg_index3(int fused_cls, const DecUnitDev* dec, const InvUnitDev* inv, const int* list, int n,
                              int* work_counter, int* err, int sm_count, cudaStream_t st, LaunchStats* ls, int s_rt = 0);
// s_rt: slab count of the FUSED_CLS_RBIG units of this launch (every listed unit has the same)
cudaError_t launch_dec_prepare(DecUnitDev* dec, int n_units, const wc_pair* dense, const int32_t* npairs,
                               unsigned long long* chain /* (n_units + 1023) / 1024 + 1 words */, int* err, cudaStream_t st,
                               LaunchStats* ls);

#ifdef WC_PHASE_PROFILE
cudaError_t debug_phase_cycles(unsigned long long out[8], bool reset);
#endif

} // namespace wc
