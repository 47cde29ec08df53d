// pfrx_specw_kernel.cuh -- entry point of a multi-warp specialised kernel; included by
// the generated file AFTER the per-role functions are defined.
#pragma once

extern "C" __global__ void __launch_bounds__(32 * SPEC_W, SPEC_MINBLOCKS)
    pfrx_spec_kernel(DevState st, long long ncell, double tran_dt, SpecParams prm, DevSummary *summ) {
  extern __shared__ double smem[];
  double *W = smem + (threadIdx.x & 31);
  switch (threadIdx.x >> 5) {
    case 0: specw_role<0>(st, ncell, tran_dt, prm, summ, W); break;
#if SPEC_W > 1
    case 1: specw_role<1>(st, ncell, tran_dt, prm, summ, W); break;
#endif
#if SPEC_W > 2
    case 2: specw_role<2>(st, ncell, tran_dt, prm, summ, W); break;
    case 3: specw_role<3>(st, ncell, tran_dt, prm, summ, W); break;
#endif
#if SPEC_W > 4
    case 4: specw_role<4>(st, ncell, tran_dt, prm, summ, W); break;
    case 5: specw_role<5>(st, ncell, tran_dt, prm, summ, W); break;
    case 6: specw_role<6>(st, ncell, tran_dt, prm, summ, W); break;
    case 7: specw_role<7>(st, ncell, tran_dt, prm, summ, W); break;
#endif
  }
}
