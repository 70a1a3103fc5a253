// gb25_f64_stubs.cu — compiled only into libgb25cuda_f64.so (-DGB25_F64).  The Float64 build runs the operator-per-kernel
// generation (gb25_kernels.cu + gb25_tend_generic.cuh); the register-blocked, TMA-staged / packed-FP32x2 and persistent
// kernels are Float32 designs and are left out.  Their launchers are never reached (use_fused is forced off, tma_available
// and launch_barotropic_persistent say no); they exist so that the shared host code links.
#ifdef GB25_F64
#include "gb25_internal.h"

void launch_tracer_tendency_v2(Handle*) {}
void launch_momentum_tendency_v2(Handle*) {}
void launch_aux_columns(Handle*) {}
void launch_generic_list(Handle*, bool, bool) {}
void launch_momentum_tendency_tma(Handle*, const Ab2Spec*) {}
void launch_tracer_tendency_tma(Handle*, const Ab2Spec*) {}
bool tma_available(Handle*) { return false; }
void tma_free(Handle*) {}
bool launch_barotropic_persistent(Handle*, float) { return false; }
void baro_plan_free(Handle*) {}
int baro_check_timeout(Handle*) { return 0; }
void baro_plan_prepare(Handle*) {}
KernelTable kernel_table_tend_v2() { return {nullptr, 0}; }
KernelTable kernel_table_tend_tma() { return {nullptr, 0}; }
KernelTable kernel_table_baro() { return {nullptr, 0}; }
#endif
