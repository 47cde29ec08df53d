// pfrx_types.cuh -- kernel argument structs shared by the generic kernels
// (pfrx_device.cuh, pfrx_tpc.cuh) and the network-specialised ones (pfrx_spec.cuh).
#pragma once
#include <stdint.h>

struct DevState {
  int64_t ld;
  double *total, *pri_molal, *immobile, *pri_act_coef, *sec_act_coef, *sec_molal, *ln_act_h2o;
  double *mnrl_volfrac, *mnrl_area, *mnrl_rate, *free_site, *eqsrfcplx_conc, *total_sorb_eq, *kinmr;
  const double *den_kg, *sat, *temp, *porosity, *volume, *soil_particle_density;
  const int *imat;
  int *num_sub_steps, *num_iterations, *num_kinetic_state_updates, *ierror;
  // appended (ABI v2) so that the leading layout the specialised cubins read is unchanged:
  // ELM per-cell scalars and the persisted N:C ratios of the SOMDECOMP sandbox
  const double *elm_w, *elm_o, *elm_t, *elm_zsoil, *elm_kscalar, *elm_bd_dry, *elm_bsw;
  double *somdec_nc;
  const double *elm_plantndemand;
  double *eqionx_ref, *eqionx_conc;
  const double *pres;  // liquid pressure (CNDEGAS sandbox, optional)
  double *sandbox_aux; // rt_auxvar%auxiliary_data of the CALCITE sandbox
  // active gas phase (RTotalGas): global_auxvar%sat(2), rt_auxvar%total(:,2), rt_auxvar%gas_pp
  const double *sat_gas;
  double *total_gas, *gas_pp;
  const double *elm_sucsat, *elm_watfc, *elm_effpor;  // GetMoistureResponse inputs (elm_flow_coupled)
  // refill skeleton: the k-th cell handed out is order[k] (a permutation of the shard: the cells that needed the
  // most Newton iterations in the previous step first); NULL: index order.  Set by the library, not by the caller.
  const int *order;
};

// shard summary accumulated with atomics, one set per warp
struct DevSummary {
  unsigned long long ncell_active, sum_its, num_cut_cells;
  long long first_failed;
  int max_its, max_kin, max_err, max_sub;
  unsigned long long next_cell;  // work counter of the refill skeleton, zeroed before every launch
};


// launch parameters of a network-specialised kernel (pfrx_spec.cuh)
struct SpecParams {
  int max_its, max_cuts;
  double max_dlnC, tol_relchange, tol_res, tol_relres, min_sat;
};
