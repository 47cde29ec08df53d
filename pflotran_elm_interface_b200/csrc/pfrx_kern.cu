// pfrx_kern.cu -- one translation unit per padded system size N.
// Compiled as  nvcc -DPFRX_N=<N> -DPFRX_L0=<L> [-DPFRX_L1=<L> [-DPFRX_L2=<L>]]
// by build.py; exports  pfrx_kernel_<N>(lanes)  to pfrx_api.cu.
#include "pfrx_tpc.cuh"

#ifndef PFRX_N
#error "compile with -DPFRX_N=<N>"
#endif

#define PFRX_CAT2(a, b) a##b
#define PFRX_CAT(a, b) PFRX_CAT2(a, b)

typedef void (*pfrx_kernel_fn)(DevCfg, DevState, int64_t, double, DevSummary *);

extern "C" pfrx_kernel_fn PFRX_CAT(pfrx_kernel_, PFRX_N)(int lanes) {
  if (lanes == 0) return pfrx_rstep_tpc_kernel<PFRX_N>;  // thread-per-cell variant
#ifdef PFRX_L0
  if (lanes == PFRX_L0) return pfrx_rstep_kernel<PFRX_N, PFRX_L0>;
#endif
#ifdef PFRX_L1
  if (lanes == PFRX_L1) return pfrx_rstep_kernel<PFRX_N, PFRX_L1>;
#endif
#ifdef PFRX_L2
  if (lanes == PFRX_L2) return pfrx_rstep_kernel<PFRX_N, PFRX_L2>;
#endif
  return nullptr;
}

typedef void (*pfrx_reaction_fn)(DevCfg, DevState, int64_t, int, double *, double *, double);
extern "C" pfrx_reaction_fn PFRX_CAT(pfrx_reaction_kernel_, PFRX_N)(void) { return pfrx_reaction_tpc_kernel<PFRX_N>; }

typedef void (*pfrx_constraint_fn)(DevCfg, DevState, int64_t, DevCons, int *, int *);
extern "C" pfrx_constraint_fn PFRX_CAT(pfrx_constraint_kernel_, PFRX_N)(void) { return pfrx_constraint_tpc_kernel<PFRX_N>; }

typedef void (*pfrx_auxvars_fn)(DevCfg, DevState, int64_t, const double *, int);
extern "C" pfrx_auxvars_fn PFRX_CAT(pfrx_auxvars_kernel_, PFRX_N)(void) { return pfrx_auxvars_tpc_kernel<PFRX_N>; }
