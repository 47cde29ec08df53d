/* abi_smoke.c -- the boundary exercised from plain C, no Python, no ctypes.
 *
 *   gcc -std=c99 -I include tests/c/abi_smoke.c -L pflotran_elm_interface_b200 -lpfrx_b200 -o abi_smoke
 *   ./abi_smoke <config dump path> [specialised cubin]
 *
 * Builds, by hand, the pfrx_config of the calcite network of BASELINE config C2 (H+ / HCO3- / Ca++,
 * six complexes from database/hanford.dat, kinetic Calcite; shortcourse/1D_Calcite/calcite_tran_only.in),
 * the way a Fortran host would flatten reaction_rt_type (reaction_aux.F90:123-311); steps 64 cells
 * of host-resident state through pfrx_rstep_host; compares with the numbers the CPU oracle wrote
 * into calcite_fixture.h (1e-10, identical Newton iteration counts); writes the configuration with
 * pfrx_config_dump so that `python -m pflotran_elm_interface_b200.specialize <dump>` can generate the
 * specialised kernel, and -- given that cubin -- steps again through the specialised kernel.
 * Exit code 0 = all comparisons passed. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pfrx.h"
#include "calcite_fixture.h"

static double *dup_d(const double *src, size_t n) {
  double *p = (double *)malloc(n * sizeof(double));
  memcpy(p, src, n * sizeof(double));
  return p;
}

static double worst(const double *got, const double *want, size_t n) {
  double w = 0.0;
  for (size_t i = 0; i < n; i++) {
    double s = fmax(fabs(got[i]), fabs(want[i]));
    if (s < 1e-30) continue;
    double e = fabs(got[i] - want[i]) / s;
    if (e > w) w = e;
  }
  return w;
}

int main(int argc, char **argv) {
  /* ---- the reaction network, by hand ------------------------------------------------- */
  static const double pri_Z[3] = {1.0, -1.0, 2.0}, pri_a0[3] = {9.0, 4.0, 6.0};
  /* OH-, CO3--, CO2(aq), CaOH+, CaHCO3+, CaCO3(aq): species ids, stoichiometry, H2O, logK(25 C) */
  static const int32_t cx_ptr[7] = {0, 1, 3, 5, 7, 9, 12};
  static const int32_t cx_id[12] = {0, 0, 1, 0, 1, 0, 2, 1, 2, 0, 1, 2};
  static const double cx_nu[12] = {-1.0, -1.0, 1.0, 1.0, 1.0, -1.0, 1.0, 1.0, 1.0, -1.0, 1.0, 1.0};
  static const double cx_h2o[6] = {1.0, 0.0, -1.0, 1.0, 0.0, 0.0};
  static const double cx_logK[6] = {13.9951, 10.3288, -6.3447, 12.850000000000023, -1.0467, 7.0017};
  static const double cx_Z[6] = {-1.0, -2.0, 0.0, 1.0, 1.0, 0.0}, cx_a0[6] = {3.5, 4.5, 3.0, 4.0, 4.0, 3.0};
  /* Calcite + H+ = Ca++ + HCO3- */
  static const int32_t mn_ptr[2] = {0, 3}, mn_id[3] = {0, 1, 2}, mn_irr[1] = {0};
  static const double mn_nu[3] = {-1.0, 1.0, 1.0}, mn_h2o[1] = {0.0}, mn_logK[1] = {1.8487};
  static const double mn_vol[1] = {3.6933999999999996e-05}, mn_rate[1] = {1e-06}, mn_zero[1] = {0.0};

  pfrx_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.abi_version = PFRX_ABI_VERSION;
  cfg.naqcomp = 3;
  cfg.use_full_geochemistry = 1;
  cfg.use_log_formulation = 1;
  cfg.use_isothermal = 1;
  cfg.act_coef_update_frequency = PFRX_ACT_COEF_FREQUENCY_TIMESTEP;
  cfg.act_coef_update_algorithm = PFRX_ACT_COEF_ALGORITHM_LAG;
  cfg.h2o_aq_id = -1;
  cfg.maximum_reaction_iterations = 20;
  cfg.maximum_reaction_cuts = 10;
  cfg.max_dlnC_rreact = 5.0;
  cfg.max_relative_change_tolerance = 1e-06;
  cfg.max_residual_tolerance = 1e-12;
  cfg.max_rel_residual_tolerance = 1e-08;
  cfg.rt_min_saturation = 1e-40;
  cfg.debyeA = 0.5114;
  cfg.debyeB = 0.3288;
  cfg.debyeBdot = 0.041;
  cfg.primary_spec_Z = pri_Z;
  cfg.primary_spec_a0 = pri_a0;
  cfg.neqcplx = 6;
  cfg.eqcplx_ptr = cx_ptr;
  cfg.eqcplx_specid = cx_id;
  cfg.eqcplx_stoich = cx_nu;
  cfg.eqcplx_h2ostoich = cx_h2o;
  cfg.eqcplx_logK = cx_logK;
  cfg.eqcplx_Z = cx_Z;
  cfg.eqcplx_a0 = cx_a0;
  cfg.nkinmnrl = 1;
  cfg.kinmnrl_ptr = mn_ptr;
  cfg.kinmnrl_specid = mn_id;
  cfg.kinmnrl_stoich = mn_nu;
  cfg.kinmnrl_h2ostoich = mn_h2o;
  cfg.kinmnrl_logK = mn_logK;
  cfg.kinmnrl_molar_vol = mn_vol;
  cfg.kinmnrl_rate_constant = mn_rate;
  cfg.kinmnrl_activation_energy = mn_zero;
  cfg.kinmnrl_affinity_threshold = mn_zero;
  cfg.kinmnrl_rate_limiter = mn_zero;
  cfg.kinmnrl_irreversible = mn_irr;

  if (argc < 2) {
    fprintf(stderr, "usage: abi_smoke <config dump path> [specialised cubin]\n");
    return 2;
  }
  /* set-up side of the boundary: needs no device */
  if (pfrx_config_write(&cfg, argv[1]) != PFRX_OK) {
    fprintf(stderr, "pfrx_config_write: %s\n", pfrx_last_error());
    return 3;
  }
  printf("signature %016llx\n", (unsigned long long)pfrx_config_signature_of(&cfg));
  if (getenv("PFRX_SMOKE_SETUP_ONLY")) return 0;

  pfrx_handle *h = NULL;
  int rc = pfrx_create(&cfg, 0, &h);
  if (rc != PFRX_OK) {
    fprintf(stderr, "pfrx_create failed (%d): %s\n", rc, pfrx_last_error());
    return 4;
  }
  if (pfrx_config_signature(h) != pfrx_config_signature_of(&cfg)) return 5;
  if (argc > 2) {
    rc = pfrx_load_specialized(h, argv[2]);
    if (rc != PFRX_OK) {
      fprintf(stderr, "pfrx_load_specialized failed (%d): %s\n", rc, pfrx_last_error());
      return 6;
    }
  }
  int info[5];
  pfrx_kernel_info(h, info);
  printf("kernel: N %d lanes %d threads %d\n", info[0], info[1], info[2]);

  /* ---- the state: host SoA, field[k * ld + cell] ---------------------------------------- */
  const int64_t n = FX_NCELL;
  pfrx_state st;
  memset(&st, 0, sizeof(st));
  st.ld = n;
  st.total = dup_d(fx_in_total, 3 * n);
  st.pri_molal = dup_d(fx_in_pri_molal, 3 * n);
  st.pri_act_coef = dup_d(fx_in_pri_act_coef, 3 * n);
  st.sec_act_coef = dup_d(fx_in_sec_act_coef, 6 * n);
  st.sec_molal = dup_d(fx_in_sec_molal, 6 * n);
  st.ln_act_h2o = dup_d(fx_in_ln_act_h2o, n);
  st.mnrl_volfrac = dup_d(fx_in_mnrl_volfrac, n);
  st.mnrl_area = dup_d(fx_in_mnrl_area, n);
  st.mnrl_rate = dup_d(fx_in_mnrl_rate, n);
  st.den_kg = fx_in_den_kg;
  st.sat = fx_in_sat;
  st.temp = fx_in_temp;
  st.porosity = fx_in_porosity;
  st.volume = fx_in_volume;
  st.soil_particle_density = fx_in_soil_particle_density;
  st.imat = fx_in_imat;
  st.num_sub_steps = (int32_t *)calloc(n, sizeof(int32_t));
  st.num_iterations = (int32_t *)calloc(n, sizeof(int32_t));
  st.num_kinetic_state_updates = (int32_t *)calloc(n, sizeof(int32_t));
  st.ierror = (int32_t *)calloc(n, sizeof(int32_t));

  pfrx_step_result res;
  rc = pfrx_rstep_host(h, n, &st, FX_TRAN_DT, &res);
  if (rc != PFRX_OK) {
    fprintf(stderr, "pfrx_rstep_host failed (%d): %s\n", rc, pfrx_last_error());
    return 7;
  }
  int bad = 0;
  const double e_tot = worst(st.total, fx_want_total, 3 * n), e_pri = worst(st.pri_molal, fx_want_pri_molal, 3 * n);
  const double e_vf = worst(st.mnrl_volfrac, fx_want_mnrl_volfrac, n), e_sec = worst(st.sec_molal, fx_want_sec_molal, 6 * n);
  printf("max relative error: total %.2e pri_molal %.2e mnrl_volfrac %.2e sec_molal %.2e\n", e_tot, e_pri, e_vf, e_sec);
  if (e_tot > 1e-10 || e_pri > 1e-10 || e_vf > 1e-10 || e_sec > 1e-10) bad |= 1;
  for (int64_t c = 0; c < n; c++)
    if (st.num_iterations[c] != fx_want_num_iterations[c]) bad |= 2;
  if (res.sum_newton_iterations != FX_SUM_NEWTON_ITERATIONS || res.ncell_active != n || res.rstep_error != 0) bad |= 4;
  printf("Newton iterations %lld (oracle %d), launches %lld\n", (long long)res.sum_newton_iterations,
         FX_SUM_NEWTON_ITERATIONS, (long long)pfrx_launch_count(h));
  if (pfrx_config_dump(h, argv[1]) != PFRX_OK) bad |= 8;
  pfrx_destroy(h);
  printf(bad ? "FAILED (%d)\n" : "OK\n", bad);
  return bad ? 10 + bad : 0;
}
