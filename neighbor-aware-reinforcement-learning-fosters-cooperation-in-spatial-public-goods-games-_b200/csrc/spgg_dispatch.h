// Kernel pickers shared by the translation units of libspgg_b200.  The library is compiled as three units in
// parallel (__graft_entry__.build): spgg_capi.cu (C ABI, host logic, fast / resident / helper kernels),
// spgg_inst_general.cu (every instantiation of k_step / k_gmax) and spgg_inst_lean.cu (k_step_lean /
// k_gmax_lean / k_build_valtab).  A kernel is compiled in the unit that instantiates it; the others launch it
// through the function pointer returned here.
#pragma once
#include "spgg_kernels.cuh"

namespace spgg {

enum { MODE_F32_I8 = 0, MODE_F32_F = 1, MODE_F64 = 2 };

typedef void (*step_fn_t)(KArgs);
typedef void (*gmax_fn_t)(GArgs);

// spgg_inst_general.cu
step_fn_t pick_step(int mode, int M, int action, int replay);
gmax_fn_t pick_gmax_general(int mode, int M);
// spgg_inst_lean.cu
step_fn_t pick_lean(int mode, int M, int action, int replay);
gmax_fn_t pick_gmax_lean(int mode, int M);
cudaError_t launch_build_valtab(const RepConst *rc_all, double *tab, int n_rep);
#ifndef SPGG_LEAN_TR_MAX
#define SPGG_LEAN_TR_MAX 16
#endif
constexpr int LEAN_TR_MAX = SPGG_LEAN_TR_MAX;   // k_step_lean addresses the shared-memory layout of 16-row tiles

}  // namespace spgg
