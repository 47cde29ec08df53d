/*
 * pfrx_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * A plain-C restatement of the reference's operator-split chemistry step, in
 * the reference's operation order, used only by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs as the checker and the
 * CPU baseline.  Nothing under pflotran_elm_interface_b200/ may call it.
 *
 * Parity status: the residual/Jacobian functions (RTotalAqueous,
 * RActivityCoefficients, RKineticMineral, surface complexation, CLM_CN_React)
 * are pinned against the reference's GIRT batch golds (tests/golden/, see
 * tests/test_oracle_golden.py).  RStep/RReact themselves (sub-stepping, OS
 * convergence tests, iteration counts) are "parity unpinned by reference
 * tests; pinned by source restatement" -- SURVEY.md section 0.2: the fork's
 * RSolve skips back-substitution on this call path, so its *_os golds encode
 * a no-op.  This oracle back-substitutes (the evident intended algorithm);
 * PFRX_ORACLE_REF_BUG_COMPAT=1 reproduces the no-op for harness self-checks.
 *
 * The reference cannot be compiled in this image (no Fortran compiler, PETSc
 * or MPI), so there is no oracle/_ref; DESIGN.md records that.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (see oracle/Makefile).
 * Every function cites the reference file:line it follows; paths are relative
 * to src/pflotran/ of the reference.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/pfrx.h"

/* tables of pfrx_config read through a pointer: the identity here; in the op-counting build
 * (pfrx_oracle_count.cpp compiles this file as C++ with `double` replaced by a counting scalar of the
 * same layout) the cast that makes the pointer types agree */
#ifdef PFRX_ORACLE_COUNTING
#define CFGP(p) ((const double *)(const void *)(p))
#else
#define CFGP(p) (p)
#endif

/* pflotran_constants.F90:84-92 (truncated on purpose, as in the reference) */
#define LOG_TO_LN 2.30258509299
#define IDEAL_GAS_CONSTANT 8.31446
#define MAX_DOUBLE 1.e20

static int g_ref_bug_compat = 0;

/* ------------------------------------------------------------------------ */
/* one cell, gathered from the SoA views (the role of rt_auxvar +             */
/* global_auxvar + material_auxvar)                                           */
typedef struct {
  int naq, nim, n, ncplx, nkin, nsrfrxn, nsrfcplx, nmrrows;
  double *total, *pri_molal, *immobile, *pri_act_coef, *sec_act_coef;
  double *sec_molal, *mnrl_volfrac, *mnrl_area, *mnrl_rate;
  double *free_site, *eqsrfcplx_conc, *total_sorb_eq, *dtotal_sorb_eq;
  double *kinmr_total_sorb;
  double *dtotal;      /* dtotal(i,j) at [i + j*naq] */
  double ln_act_h2o;
  double den_kg, sat, temp, porosity, volume, soil_particle_density;
  double pres; /* liquid pressure, CNDEGAS only */
  double sandbox_aux; /* rt_auxvar%auxiliary_data of the CALCITE sandbox */
  /* active gas phase */
  int ngas;
  double sat_gas;
  double *total_gas, *dtotal_gas, *gas_pp, *acteq_logK;
  /* ELM per-cell scalars (elm_pflotran builds) */
  double elm_w, elm_o, elm_t, elm_zsoil, elm_kscalar, elm_bd_dry, elm_bsw, elm_plantndemand;
  double elm_sucsat, elm_watfc, elm_effpor; /* GetMoistureResponse inputs (elm_flow_coupled) */
  double *somdec_nc; /* persisted N:C ratios, see pfrx_state.somdec_nc */
  int nsomdec_nc;
  double *eqionx_ref, *eqionx_conc; /* ion exchange: reference-cation sorbed conc, cation concs */
  int nionx, nionxcat;
  /* per-cell copies of the temperature dependent tables
   * (the reference overwrites the shared ones, reaction.F90:6003-6031) */
  double *eqcplx_logK, *kinmnrl_logK, *srfcplx_logK;
  /* scratch */
  double *buf;
  int option_ierror;
} cell_t;

static int mr_rows(const pfrx_config *cfg) {
  int nmr = cfg->nkinmrsrfcplxrxn;
  if (nmr <= 0) return 0;
  return cfg->naqcomp * (cfg->kinmr_rate_ptr[nmr] + nmr);
}

static size_t cell_doubles(const pfrx_config *cfg) {
  size_t naq = cfg->naqcomp, nim = cfg->nimcomp, nc = cfg->neqcplx;
  size_t nk = cfg->nkinmnrl, nr = cfg->nsrfcplxrxn, ns = cfg->nsrfcplx;
  size_t nsd = cfg->somdec ? (size_t)(cfg->somdec->nrxn + cfg->somdec->downstream_ptr[cfg->somdec->nrxn]) : 0;
  return 4 * naq + nim + 3 * nc + 4 * nk + nr + 2 * ns + naq + 2 * naq * naq + naq + naq * naq +
         2 * (size_t)(cfg->nactive_gas > 0 ? cfg->nactive_gas : 0) +
         mr_rows(cfg) + nsd + (cfg->neqionxrxn > 0 ? cfg->neqionxrxn + cfg->eqionx_ptr[cfg->neqionxrxn] : 0) + 64;
}

static void cell_init(cell_t *c, const pfrx_config *cfg) {
  memset(c, 0, sizeof(*c));
  c->naq = cfg->naqcomp;
  c->nim = cfg->nimcomp;
  c->n = c->naq + c->nim;
  c->ncplx = cfg->neqcplx;
  c->nkin = cfg->nkinmnrl;
  c->nsrfrxn = cfg->nsrfcplxrxn;
  c->nsrfcplx = cfg->nsrfcplx;
  c->nmrrows = mr_rows(cfg);
  c->buf = (double *)calloc(cell_doubles(cfg), sizeof(double));
  double *p = c->buf;
#define TAKE(field, cnt) \
  c->field = p;          \
  p += (cnt)
  TAKE(total, c->naq);
  TAKE(pri_molal, c->naq);
  TAKE(immobile, c->nim);
  TAKE(pri_act_coef, c->naq);
  TAKE(sec_act_coef, c->ncplx);
  TAKE(sec_molal, c->ncplx);
  TAKE(mnrl_volfrac, c->nkin);
  TAKE(mnrl_area, c->nkin);
  TAKE(mnrl_rate, c->nkin);
  TAKE(free_site, c->nsrfrxn);
  TAKE(eqsrfcplx_conc, c->nsrfcplx);
  TAKE(total_sorb_eq, c->naq);
  TAKE(dtotal_sorb_eq, c->naq * c->naq);
  TAKE(kinmr_total_sorb, c->nmrrows);
  TAKE(dtotal, c->naq * c->naq);
  TAKE(eqcplx_logK, c->ncplx);
  TAKE(kinmnrl_logK, c->nkin);
  TAKE(srfcplx_logK, c->nsrfcplx);
  c->nsomdec_nc = cfg->somdec ? cfg->somdec->nrxn + cfg->somdec->downstream_ptr[cfg->somdec->nrxn] : 0;
  TAKE(somdec_nc, c->nsomdec_nc);
  c->nionx = cfg->neqionxrxn;
  c->nionxcat = cfg->neqionxrxn > 0 ? cfg->eqionx_ptr[cfg->neqionxrxn] : 0;
  TAKE(eqionx_ref, c->nionx);
  TAKE(eqionx_conc, c->nionxcat);
  c->ngas = cfg->nactive_gas > 0 ? cfg->nactive_gas : 0;
  TAKE(total_gas, c->naq);
  TAKE(dtotal_gas, c->naq * c->naq);
  TAKE(gas_pp, c->ngas);
  TAKE(acteq_logK, c->ngas);
#undef TAKE
  if (c->ngas) memcpy(c->acteq_logK, cfg->acteq_logK, sizeof(double) * c->ngas);
  if (c->ncplx) memcpy(c->eqcplx_logK, cfg->eqcplx_logK, sizeof(double) * c->ncplx);
  if (c->nkin) memcpy(c->kinmnrl_logK, cfg->kinmnrl_logK, sizeof(double) * c->nkin);
  if (c->nsrfcplx) memcpy(c->srfcplx_logK, cfg->srfcplx_logK, sizeof(double) * c->nsrfcplx);
}

static void cell_free(cell_t *c) { free(c->buf); }

#define LD(arr, k) ((arr)[(int64_t)(k) * st->ld + ic])

static void cell_gather(cell_t *c, const pfrx_config *cfg, const pfrx_state *st,
                        int64_t ic) {
  int k;
  for (k = 0; k < c->nsomdec_nc; k++) {
    const pfrx_somdec *sd = cfg->somdec;
    if (st->somdec_nc)
      c->somdec_nc[k] = LD(st->somdec_nc, k);
    else
      c->somdec_nc[k] = k < sd->nrxn ? sd->upstream_nc[k] : sd->downstream_nc[k - sd->nrxn];
  }
  for (k = 0; k < c->nionx; k++)
    c->eqionx_ref[k] = st->eqionx_ref_cation_sorbed_conc ? LD(st->eqionx_ref_cation_sorbed_conc, k) : 1.e-9;
  for (k = 0; k < c->nionxcat; k++) c->eqionx_conc[k] = st->eqionx_conc ? LD(st->eqionx_conc, k) : 0.0;
  for (k = 0; k < c->naq; k++) {
    c->total[k] = LD(st->total, k);
    c->pri_molal[k] = LD(st->pri_molal, k);
    c->pri_act_coef[k] = LD(st->pri_act_coef, k);
    if (st->total_sorb_eq) c->total_sorb_eq[k] = LD(st->total_sorb_eq, k);
  }
  for (k = 0; k < c->nim; k++) c->immobile[k] = LD(st->immobile, k);
  for (k = 0; k < c->ncplx; k++) {
    c->sec_act_coef[k] = LD(st->sec_act_coef, k);
    c->sec_molal[k] = LD(st->sec_molal, k);
  }
  c->ln_act_h2o = st->ln_act_h2o ? LD(st->ln_act_h2o, 0) : 0.0;
  for (k = 0; k < c->nkin; k++) {
    c->mnrl_volfrac[k] = LD(st->mnrl_volfrac, k);
    c->mnrl_area[k] = LD(st->mnrl_area, k);
    c->mnrl_rate[k] = LD(st->mnrl_rate, k);
  }
  for (k = 0; k < c->nsrfrxn; k++) c->free_site[k] = LD(st->srfcplxrxn_free_site_conc, k);
  for (k = 0; k < c->nsrfcplx; k++)
    c->eqsrfcplx_conc[k] = st->eqsrfcplx_conc ? LD(st->eqsrfcplx_conc, k) : 0.0;
  for (k = 0; k < c->nmrrows; k++) c->kinmr_total_sorb[k] = LD(st->kinmr_total_sorb, k);
  c->den_kg = LD(st->den_kg, 0);
  c->sat = LD(st->sat, 0);
  c->temp = LD(st->temp, 0);
  c->porosity = LD(st->porosity, 0);
  c->volume = LD(st->volume, 0);
  c->soil_particle_density = st->soil_particle_density ? LD(st->soil_particle_density, 0) : 0.0;
  c->pres = st->pres ? LD(st->pres, 0) : 101325.0;
  c->sandbox_aux = st->sandbox_aux ? LD(st->sandbox_aux, 0) : 0.0;
  c->sat_gas = st->sat_gas ? LD(st->sat_gas, 0) : 0.0;
  if (c->ngas) {
    for (k = 0; k < c->naq; k++) c->total_gas[k] = st->total_gas ? LD(st->total_gas, k) : 0.0;
    for (k = 0; k < c->ngas; k++) c->gas_pp[k] = st->gas_pp ? LD(st->gas_pp, k) : 0.0;
  }
  c->elm_w = st->elm_w_scalar ? LD(st->elm_w_scalar, 0) : 1.0;
  c->elm_o = st->elm_o_scalar ? LD(st->elm_o_scalar, 0) : 1.0;
  c->elm_t = st->elm_t_scalar ? LD(st->elm_t_scalar, 0) : 1.0;
  c->elm_zsoil = st->elm_zsoil ? LD(st->elm_zsoil, 0) : 0.0;
  c->elm_kscalar = st->elm_kscalar_decomp_c ? LD(st->elm_kscalar_decomp_c, 0) : 1.0;
  c->elm_bd_dry = st->elm_bulkdensity_dry ? LD(st->elm_bulkdensity_dry, 0) : 1.25e3;
  c->elm_bsw = st->elm_bsw ? LD(st->elm_bsw, 0) : 1.0;
  c->elm_sucsat = st->elm_sucsat ? LD(st->elm_sucsat, 0) : 200.0;
  c->elm_watfc = st->elm_watfc ? LD(st->elm_watfc, 0) : 0.1;
  c->elm_effpor = st->elm_effporosity ? LD(st->elm_effporosity, 0) : 0.4;
  c->elm_plantndemand = st->elm_rate_plantndemand ? LD(st->elm_rate_plantndemand, 0) : 0.0;
  c->option_ierror = 0;
}

static void cell_scatter(const cell_t *c, const pfrx_state *st, int64_t ic) {
  int k;
  for (k = 0; k < c->naq; k++) {
    LD(st->total, k) = c->total[k];
    LD(st->pri_molal, k) = c->pri_molal[k];
    LD(st->pri_act_coef, k) = c->pri_act_coef[k];
    if (st->total_sorb_eq) LD(st->total_sorb_eq, k) = c->total_sorb_eq[k];
  }
  for (k = 0; k < c->nim; k++) LD(st->immobile, k) = c->immobile[k];
  for (k = 0; k < c->ncplx; k++) {
    LD(st->sec_act_coef, k) = c->sec_act_coef[k];
    LD(st->sec_molal, k) = c->sec_molal[k];
  }
  if (st->ln_act_h2o) LD(st->ln_act_h2o, 0) = c->ln_act_h2o;
  for (k = 0; k < c->nkin; k++) {
    LD(st->mnrl_volfrac, k) = c->mnrl_volfrac[k];
    LD(st->mnrl_rate, k) = c->mnrl_rate[k];
  }
  for (k = 0; k < c->nsrfrxn; k++) LD(st->srfcplxrxn_free_site_conc, k) = c->free_site[k];
  if (st->eqsrfcplx_conc)
    for (k = 0; k < c->nsrfcplx; k++) LD(st->eqsrfcplx_conc, k) = c->eqsrfcplx_conc[k];
  for (k = 0; k < c->nmrrows; k++) LD(st->kinmr_total_sorb, k) = c->kinmr_total_sorb[k];
  if (st->somdec_nc)
    for (k = 0; k < c->nsomdec_nc; k++) LD(st->somdec_nc, k) = c->somdec_nc[k];
  if (st->eqionx_ref_cation_sorbed_conc)
    for (k = 0; k < c->nionx; k++) LD(st->eqionx_ref_cation_sorbed_conc, k) = c->eqionx_ref[k];
  if (st->eqionx_conc)
    for (k = 0; k < c->nionxcat; k++) LD(st->eqionx_conc, k) = c->eqionx_conc[k];
  if (st->sandbox_aux) LD(st->sandbox_aux, 0) = c->sandbox_aux;
  if (c->ngas) {
    if (st->total_gas)
      for (k = 0; k < c->naq; k++) LD(st->total_gas, k) = c->total_gas[k];
    if (st->gas_pp)
      for (k = 0; k < c->ngas; k++) LD(st->gas_pp, k) = c->gas_pp[k];
  }
}

/* ------------------------------------------------------------------------ */
/* utility.F90:597-688  LUDecomposition1 (NR ludcmp, Crout, implicit scaling) */
static int lu_decomposition(double *A, int N, int *indx) {
  const double tiny = 1.0e-20;
  double VV[PFRX_MAX_NCOMP * 4];
  double *vv = VV, *heap = NULL;
  int i, j, k, imax = 0;
  double aamax, sum, dum;
  if (N > PFRX_MAX_NCOMP * 4) vv = heap = (double *)malloc(sizeof(double) * N);
#define a(i, j) A[(i) + (size_t)(j) * N]
  for (i = 0; i < N; i++) {
    aamax = 0.0;
    for (j = 0; j < N; j++)
      if (fabs(a(i, j)) > aamax) aamax = fabs(a(i, j));
    if (aamax <= 0.0) {
      free(heap);
      return 1; /* singular row, stop_on_error = false */
    }
    vv[i] = 1. / aamax;
  }
  for (j = 0; j < N; j++) {
    for (i = 0; i < j; i++) {
      sum = a(i, j);
      for (k = 0; k < i; k++) sum = sum - a(i, k) * a(k, j);
      a(i, j) = sum;
    }
    aamax = 0;
    imax = j;
    for (i = j; i < N; i++) {
      sum = a(i, j);
      for (k = 0; k < j; k++) sum = sum - a(i, k) * a(k, j);
      a(i, j) = sum;
      dum = vv[i] * fabs(sum);
      if (dum >= aamax) {
        imax = i;
        aamax = dum;
      }
    }
    if (j != imax) {
      for (k = 0; k < N; k++) {
        dum = a(imax, k);
        a(imax, k) = a(j, k);
        a(j, k) = dum;
      }
      vv[imax] = vv[j];
    }
    indx[j] = imax;
    if (a(j, j) == 0.0) a(j, j) = tiny;
    if (j != N - 1) {
      dum = 1.0 / a(j, j);
      for (i = j + 1; i < N; i++) a(i, j) = a(i, j) * dum;
    }
  }
  free(heap);
  return 0;
}

/* utility.F90:692-735  LUBackSubstitution (NR lubksb) */
static void lu_back_substitution(const double *A, int N, const int *indx, double *B) {
  int i, j, ii = -1, ll;
  double sum;
  for (i = 0; i < N; i++) {
    ll = indx[i];
    sum = B[ll];
    B[ll] = B[i];
    if (ii != -1) {
      for (j = ii; j <= i - 1; j++) sum = sum - a(i, j) * B[j];
    } else if (sum != 0.0) {
      ii = i;
    }
    B[i] = sum;
  }
  for (i = N - 1; i >= 0; i--) {
    sum = B[i];
    if (i < N - 1)
      for (j = i + 1; j < N; j++) sum = sum - a(i, j) * B[j];
    B[i] = sum / a(i, i);
  }
}
#undef a

/* reaction_aux.F90:1285-1312  ReactionInterpolateLogK */
static void interpolate_logK(const double *coefs, double *logKs, double temp, int n) {
  double temp_kelvin = temp + 273.15;
  int i;
  for (i = 0; i < n; i++) {
    const double *c = coefs + 5 * i;
    logKs[i] = c[0] * log(temp_kelvin) + c[1] + c[2] * temp_kelvin + c[3] / temp_kelvin +
               c[4] / (temp_kelvin * temp_kelvin);
  }
}

/* reaction.F90:5976-6067  RUpdateTempDependentCoefs (non-hpt branch) */
static void update_temp_dependent_coefs(cell_t *c, const pfrx_config *cfg) {
  if (cfg->eqcplx_logKcoef) interpolate_logK(CFGP(cfg->eqcplx_logKcoef), c->eqcplx_logK, c->temp, c->ncplx);
  if (cfg->kinmnrl_logKcoef) interpolate_logK(CFGP(cfg->kinmnrl_logKcoef), c->kinmnrl_logK, c->temp, c->nkin);
  if (cfg->srfcplx_logKcoef) interpolate_logK(CFGP(cfg->srfcplx_logKcoef), c->srfcplx_logK, c->temp, c->nsrfcplx);
  if (c->ngas && cfg->acteq_logK_coef) interpolate_logK(CFGP(cfg->acteq_logK_coef), c->acteq_logK, c->temp, c->ngas);
}

/* ------------------------------------------------------------------------ */
/* reaction.F90:4368-4614  RActivityCoefficients */
static void r_activity_coefficients(cell_t *c, const pfrx_config *cfg) {
  int icplx, icomp, it, j, jcomp, i;
  double I, sqrt_I, II, f, fpri, didi, dcdi = 0, den, dgamdi, lnQK, sum;
  double sum_pri_molal = 0.0, sum_sec_molal;
  const double *Z = CFGP(cfg->primary_spec_Z), *a0 = CFGP(cfg->primary_spec_a0);
  const double *cZ = CFGP(cfg->eqcplx_Z), *ca0 = CFGP(cfg->eqcplx_a0);

  if (cfg->use_activity_h2o) {
    sum_pri_molal = 0.0;
    for (j = 0; j < c->naq; j++)
      if (j != cfg->h2o_aq_id) sum_pri_molal = sum_pri_molal + c->pri_molal[j];
  }

  if (cfg->act_coef_update_algorithm == PFRX_ACT_COEF_ALGORITHM_NEWTON) {
    double ln_conc[PFRX_MAX_NCOMP * 4], ln_act[PFRX_MAX_NCOMP * 4];
    for (j = 0; j < c->naq; j++) {
      ln_conc[j] = log(c->pri_molal[j]);
      ln_act[j] = ln_conc[j] + log(c->pri_act_coef[j]);
    }
    fpri = 0.0;
    for (j = 0; j < c->naq; j++) fpri = fpri + c->pri_molal[j] * Z[j] * Z[j];
    it = 0;
    II = 0;
    for (;;) {
      it = it + 1;
      if (it > 50) {
        /* reference poisons with NaN and never leaves the loop
         * (reaction.F90:4421-4431); we flag the error and leave */
        for (j = 0; j < c->naq; j++) {
          c->pri_molal[j] = NAN;
          c->pri_act_coef[j] = NAN;
        }
        for (j = 0; j < c->ncplx; j++) c->sec_act_coef[j] = NAN;
        c->option_ierror = 1;
        return;
      }
      I = fpri;
      for (icplx = 0; icplx < c->ncplx; icplx++)
        I = I + c->sec_molal[icplx] * cZ[icplx] * cZ[icplx];
      I = 0.5 * I;
      f = I;
      if (fabs(I - II) < 1.e-6 * I) break;

      if (c->ncplx > 0) {
        didi = 0.0;
        sqrt_I = sqrt(I);
        for (icplx = 0; icplx < c->ncplx; icplx++) {
          if (fabs(cZ[icplx]) > 0.0) {
            double t = 1.0 + cfg->debyeB * ca0[icplx] * sqrt_I;
            sum = 0.5 * cfg->debyeA * cZ[icplx] * cZ[icplx] / (sqrt_I * (t * t)) - cfg->debyeBdot;
            for (i = cfg->eqcplx_ptr[icplx]; i < cfg->eqcplx_ptr[icplx + 1]; i++) {
              j = cfg->eqcplx_specid[i];
              if (fabs(Z[j]) > 0.0) {
                double tj = 1.0 + cfg->debyeB * a0[j] * sqrt_I;
                dgamdi = -0.5 * cfg->debyeA * (Z[j] * Z[j]) / (sqrt_I * (tj * tj)) + cfg->debyeBdot;
                sum = sum + cfg->eqcplx_stoich[i] * dgamdi;
              }
            }
            dcdi = c->sec_molal[icplx] * LOG_TO_LN * sum;
            didi = didi + 0.5 * cZ[icplx] * cZ[icplx] * dcdi;
          }
        }
        den = 1.0 - didi;
        if (fabs(den) > 0.0)
          II = (f - I * didi) / den;
        else
          II = f;
      } else {
        II = f;
      }
      if (II < 0.0) {
        for (j = 0; j < c->naq; j++) {
          c->pri_molal[j] = NAN;
          c->pri_act_coef[j] = NAN;
        }
        for (j = 0; j < c->ncplx; j++) c->sec_act_coef[j] = NAN;
        c->option_ierror = 1;
        return;
      }
      I = II;
      sqrt_I = sqrt(I);
      for (icomp = 0; icomp < c->naq; icomp++) {
        if (fabs(Z[icomp]) > 0.0)
          c->pri_act_coef[icomp] =
              exp((-Z[icomp] * Z[icomp] * sqrt_I * cfg->debyeA / (1.0 + a0[icomp] * cfg->debyeB * sqrt_I) +
                   cfg->debyeBdot * I) *
                  LOG_TO_LN);
        else
          c->pri_act_coef[icomp] = 1.0;
      }
      sum_sec_molal = 0.0;
      for (icplx = 0; icplx < c->ncplx; icplx++) {
        if (fabs(cZ[icplx]) > 0.0)
          c->sec_act_coef[icplx] =
              exp((-cZ[icplx] * cZ[icplx] * sqrt_I * cfg->debyeA / (1.0 + ca0[icplx] * cfg->debyeB * sqrt_I) +
                   cfg->debyeBdot * I) *
                  LOG_TO_LN);
        else
          c->sec_act_coef[icplx] = 1.0;
        lnQK = -c->eqcplx_logK[icplx] * LOG_TO_LN;
        if (cfg->eqcplx_h2ostoich[icplx] != 0.0) lnQK = lnQK + cfg->eqcplx_h2ostoich[icplx] * c->ln_act_h2o;
        for (i = cfg->eqcplx_ptr[icplx]; i < cfg->eqcplx_ptr[icplx + 1]; i++) {
          jcomp = cfg->eqcplx_specid[i];
          lnQK = lnQK + cfg->eqcplx_stoich[i] * ln_act[jcomp];
        }
        c->sec_molal[icplx] = exp(lnQK) / c->sec_act_coef[icplx];
        sum_sec_molal = sum_sec_molal + c->sec_molal[icplx];
      }
      if (cfg->use_activity_h2o) {
        c->ln_act_h2o = 1.0 - 0.017 * (sum_pri_molal + sum_sec_molal);
        if (c->ln_act_h2o > 0.0)
          c->ln_act_h2o = log(c->ln_act_h2o);
        else
          c->ln_act_h2o = 0.0;
      }
    }
  } else {
    /* LAG algorithm, reaction.F90:4553-4612 */
    I = 0.0;
    for (icomp = 0; icomp < c->naq; icomp++) I = I + c->pri_molal[icomp] * Z[icomp] * Z[icomp];
    for (icplx = 0; icplx < c->ncplx; icplx++) I = I + c->sec_molal[icplx] * cZ[icplx] * cZ[icplx];
    I = 0.5 * I;
    sqrt_I = sqrt(I);
    for (icomp = 0; icomp < c->naq; icomp++) {
      if (fabs(Z[icomp]) > 1.e-10)
        c->pri_act_coef[icomp] =
            exp((-Z[icomp] * Z[icomp] * sqrt_I * cfg->debyeA / (1.0 + a0[icomp] * cfg->debyeB * sqrt_I) +
                 cfg->debyeBdot * I) *
                LOG_TO_LN);
      else
        c->pri_act_coef[icomp] = 1.0;
    }
    sum_sec_molal = 0.0;
    for (icplx = 0; icplx < c->ncplx; icplx++) {
      if (fabs(cZ[icplx]) > 1.e-10)
        c->sec_act_coef[icplx] =
            exp((-cZ[icplx] * cZ[icplx] * sqrt_I * cfg->debyeA / (1.0 + ca0[icplx] * cfg->debyeB * sqrt_I) +
                 cfg->debyeBdot * I) *
                LOG_TO_LN);
      else
        c->sec_act_coef[icplx] = 1.0;
      sum_sec_molal = sum_sec_molal + c->sec_molal[icplx];
    }
    if (cfg->use_activity_h2o) {
      c->ln_act_h2o = 1.0 - 0.017 * (sum_pri_molal + sum_sec_molal);
      if (c->ln_act_h2o > 0.0)
        c->ln_act_h2o = log(c->ln_act_h2o);
      else
        c->ln_act_h2o = 0.0;
    }
  }
}

/* ------------------------------------------------------------------------ */
/* reaction.F90:4665-4759  RTotalAqueous */
static void r_total_aqueous(cell_t *c, const pfrx_config *cfg) {
  int i, j, icplx, icomp, jcomp, naq = c->naq;
  double ln_conc[PFRX_MAX_NCOMP * 4], ln_act[PFRX_MAX_NCOMP * 4];
  double lnQK, tempreal, den_kg_per_L;

  den_kg_per_L = c->den_kg * 1.0 * 1.e-3; /* xmass = 1 */
  for (i = 0; i < naq; i++) {
    ln_conc[i] = log(c->pri_molal[i]);
    ln_act[i] = ln_conc[i] + log(c->pri_act_coef[i]);
    c->total[i] = c->pri_molal[i];
  }
  for (i = 0; i < naq * naq; i++) c->dtotal[i] = 0.0;
  for (i = 0; i < naq; i++) c->dtotal[i + i * naq] = 1.0;

  for (icplx = 0; icplx < c->ncplx; icplx++) {
    int p0 = cfg->eqcplx_ptr[icplx], p1 = cfg->eqcplx_ptr[icplx + 1];
    lnQK = -c->eqcplx_logK[icplx] * LOG_TO_LN;
    if (cfg->eqcplx_h2ostoich[icplx] != 0.0) lnQK = lnQK + cfg->eqcplx_h2ostoich[icplx] * c->ln_act_h2o;
    for (i = p0; i < p1; i++) {
      icomp = cfg->eqcplx_specid[i];
      lnQK = lnQK + cfg->eqcplx_stoich[i] * ln_act[icomp];
    }
    c->sec_molal[icplx] = exp(lnQK) / c->sec_act_coef[icplx];
    for (i = p0; i < p1; i++) {
      icomp = cfg->eqcplx_specid[i];
      c->total[icomp] = c->total[icomp] + cfg->eqcplx_stoich[i] * c->sec_molal[icplx];
    }
    for (j = p0; j < p1; j++) {
      jcomp = cfg->eqcplx_specid[j];
      tempreal = cfg->eqcplx_stoich[j] * exp(lnQK - ln_conc[jcomp]) / c->sec_act_coef[icplx];
      for (i = p0; i < p1; i++) {
        icomp = cfg->eqcplx_specid[i];
        c->dtotal[icomp + jcomp * naq] = c->dtotal[icomp + jcomp * naq] + cfg->eqcplx_stoich[i] * tempreal;
      }
    }
  }
  for (i = 0; i < naq; i++) c->total[i] = c->total[i] * den_kg_per_L;
  for (i = 0; i < naq * naq; i++) c->dtotal[i] = c->dtotal[i] * den_kg_per_L;
}

/* reaction_surf_complex.F90:641-900  RTotalSorbEqSurfCplx1
 * srfcplx_out may be NULL (the multirate caller passes a null pointer). */
static void r_total_sorb_eq_surf_cplx1(cell_t *c, const pfrx_config *cfg, int irxn,
                                       double *external_free_site_conc, double *srfcplx_out,
                                       double *external_total_sorb, double *external_dtotal_sorb) {
  int i, j, k, icplx, icomp, jcomp, naq = c->naq;
  double srfcplx_conc[256];
  double dSx_dmi[PFRX_MAX_NCOMP * 4];
  double ln_conc[PFRX_MAX_NCOMP * 4], ln_act[PFRX_MAX_NCOMP * 4];
  double nui_Si_over_Sx, ln_free_site, lnQK, tempreal, total;
  const double tol = 1.e-12;
  int one_more, num_iterations;
  double res, dres_dfree_site, dfree_site_conc, free_site_conc, rel_change;
  double site_density = 0.0, damping_factor;
  int r0 = cfg->srfcplxrxn_ptr[irxn], r1 = cfg->srfcplxrxn_ptr[irxn + 1];

  for (i = 0; i < naq; i++) {
    ln_conc[i] = log(c->pri_molal[i]);
    ln_act[i] = ln_conc[i] + log(c->pri_act_coef[i]);
  }
  free_site_conc = fmax(*external_free_site_conc, 1.e-40);
  for (i = 0; i < c->nsrfcplx; i++) srfcplx_conc[i] = 0.0;

  switch (cfg->srfcplxrxn_surf_type[irxn]) {
    case PFRX_MINERAL_SURFACE:
      site_density = cfg->srfcplxrxn_site_density[irxn] * c->mnrl_volfrac[cfg->srfcplxrxn_to_surf[irxn]];
      break;
    case PFRX_ROCK_SURFACE:
      site_density = cfg->srfcplxrxn_site_density[irxn] * c->soil_particle_density * (1.0 - c->porosity);
      break;
    default:
      site_density = cfg->srfcplxrxn_site_density[irxn];
  }

  if (site_density < 1.e-40) {
    *external_free_site_conc = 0.0;
    if (srfcplx_out)
      for (j = r0; j < r1; j++) srfcplx_out[cfg->srfcplxrxn_to_complex[j]] = 0.0;
    return;
  }

  one_more = 0;
  num_iterations = 0;
  damping_factor = 1.0;
  for (;;) {
    num_iterations = num_iterations + 1;
    total = free_site_conc;
    ln_free_site = log(free_site_conc);
    for (j = r0; j < r1; j++) {
      icplx = cfg->srfcplxrxn_to_complex[j];
      lnQK = -c->srfcplx_logK[icplx] * LOG_TO_LN;
      if (cfg->srfcplx_h2ostoich[icplx] != 0.0) lnQK = lnQK + cfg->srfcplx_h2ostoich[icplx] * c->ln_act_h2o;
      lnQK = lnQK + cfg->srfcplx_free_site_stoich[icplx] * ln_free_site;
      for (i = cfg->srfcplx_ptr[icplx]; i < cfg->srfcplx_ptr[icplx + 1]; i++) {
        icomp = cfg->srfcplx_specid[i];
        lnQK = lnQK + cfg->srfcplx_stoich[i] * ln_act[icomp];
      }
      srfcplx_conc[icplx] = exp(lnQK);
      total = total + cfg->srfcplx_free_site_stoich[icplx] * srfcplx_conc[icplx];
    }
    if (one_more) break;
    if (cfg->srfcplxrxn_stoich_flag[irxn]) {
      res = site_density - total;
      dres_dfree_site = 1.0;
      for (j = r0; j < r1; j++) {
        icplx = cfg->srfcplxrxn_to_complex[j];
        dres_dfree_site =
            dres_dfree_site + cfg->srfcplx_free_site_stoich[icplx] * srfcplx_conc[icplx] / free_site_conc;
      }
      dfree_site_conc = res / dres_dfree_site;
      if (num_iterations > 1000) damping_factor = 0.5;
      free_site_conc = free_site_conc + damping_factor * dfree_site_conc;
      rel_change = fabs(dfree_site_conc / free_site_conc);
      if (rel_change < tol) one_more = 1;
      if (num_iterations > 100000) { /* reference has no cap; avoid a hang */
        c->option_ierror = 1;
        one_more = 1;
      }
    } else {
      total = total / free_site_conc;
      free_site_conc = site_density / total;
      one_more = 1;
    }
  }
  *external_free_site_conc = free_site_conc;

  for (i = 0; i < naq; i++) dSx_dmi[i] = 0.0;
  tempreal = 0.0;
  for (j = r0; j < r1; j++) {
    icplx = cfg->srfcplxrxn_to_complex[j];
    for (i = cfg->srfcplx_ptr[icplx]; i < cfg->srfcplx_ptr[icplx + 1]; i++) {
      icomp = cfg->srfcplx_specid[i];
      dSx_dmi[icomp] =
          dSx_dmi[icomp] + cfg->srfcplx_stoich[i] * cfg->srfcplx_free_site_stoich[icplx] * srfcplx_conc[icplx];
    }
    tempreal = tempreal + cfg->srfcplx_free_site_stoich[icplx] * cfg->srfcplx_free_site_stoich[icplx] *
                              srfcplx_conc[icplx];
  }
  tempreal = tempreal / free_site_conc;
  tempreal = tempreal + 1.0;
  for (i = 0; i < naq; i++) {
    dSx_dmi[i] = -dSx_dmi[i] / tempreal;
    dSx_dmi[i] = dSx_dmi[i] / c->pri_molal[i];
  }
  if (srfcplx_out)
    for (i = 0; i < c->nsrfcplx; i++) srfcplx_out[i] = srfcplx_out[i] + srfcplx_conc[i];

  for (k = r0; k < r1; k++) {
    int p0, p1;
    icplx = cfg->srfcplxrxn_to_complex[k];
    p0 = cfg->srfcplx_ptr[icplx];
    p1 = cfg->srfcplx_ptr[icplx + 1];
    for (i = p0; i < p1; i++) {
      icomp = cfg->srfcplx_specid[i];
      external_total_sorb[icomp] = external_total_sorb[icomp] + cfg->srfcplx_stoich[i] * srfcplx_conc[icplx];
    }
    nui_Si_over_Sx = cfg->srfcplx_free_site_stoich[icplx] * srfcplx_conc[icplx] / free_site_conc;
    for (j = p0; j < p1; j++) {
      jcomp = cfg->srfcplx_specid[j];
      tempreal = cfg->srfcplx_stoich[j] * srfcplx_conc[icplx] / c->pri_molal[jcomp] + nui_Si_over_Sx * dSx_dmi[jcomp];
      for (i = p0; i < p1; i++) {
        icomp = cfg->srfcplx_specid[i];
        external_dtotal_sorb[icomp + jcomp * naq] =
            external_dtotal_sorb[icomp + jcomp * naq] + cfg->srfcplx_stoich[i] * tempreal;
      }
    }
  }
}

/* reaction.F90:881: neqsorb = neqionxrxn + neqkdrxn + neqsrfcplxrxn (+ neqdynamickdrxn) */
static int neqsorb(const pfrx_config *cfg) {
  return cfg->neqsrfcplxrxn + cfg->neqionxrxn + cfg->neqkdrxn + cfg->neqdynamickdrxn;
}

/* reaction.F90:4906-5140  RTotalSorbEqIonx; returns 1 where the reference sets ierror
 * (more than 20000 iterations of the inner Newton) */
static int r_total_sorb_eq_ionx(cell_t *c, const pfrx_config *cfg) {
  const double tol = 1.e-12;
  int naq = c->naq, irxn, i, j;
  double cation_X[PFRX_MAX_NCOMP];
  for (i = 0; i < c->nionxcat; i++) c->eqionx_conc[i] = 0.0;
  for (irxn = 0; irxn < cfg->neqionxrxn; irxn++) {
    int p0 = cfg->eqionx_ptr[irxn], ncomp = cfg->eqionx_ptr[irxn + 1] - p0;
    const int *cat = cfg->eqionx_cationid + p0;
    const double *kk = CFGP(cfg->eqionx_k) + p0;
    double omega, sumZX;
    if (cfg->eqionx_to_surf[irxn] >= 0)
      omega = fmax(cfg->eqionx_CEC[irxn] * c->mnrl_volfrac[cfg->eqionx_to_surf[irxn]], 1.e-40);
    else
      omega = cfg->eqionx_CEC[irxn];
    if (cfg->eqionx_Z_flag[irxn]) {
      int icomp = cat[0], one_more = 0, it = 0;
      double ref_cation_conc = c->pri_molal[icomp] * c->pri_act_coef[icomp];
      double ref_cation_Z = cfg->primary_spec_Z[icomp];
      double ref_cation_k = kk[0];
      double ref_cation_X = ref_cation_Z * c->eqionx_ref[irxn] / omega;
      double KDj, total, dres_dKDj, res, delta_KDj;
      for (j = 0; j < ncomp; j++) cation_X[j] = 0.0;
      KDj = ref_cation_X / (ref_cation_k * ref_cation_conc);
      for (;;) {
        it++;
        if (it > 20000) return 1;
        ref_cation_X = KDj * (ref_cation_k * ref_cation_conc);
        cation_X[0] = ref_cation_X;
        total = ref_cation_X;
        dres_dKDj = 0.0;
        for (j = 1; j < ncomp; j++) {
          icomp = cat[j];
          cation_X[j] = kk[j] * c->pri_molal[icomp] * c->pri_act_coef[icomp] *
                        pow(KDj, cfg->primary_spec_Z[icomp] / ref_cation_Z);
          total = total + cation_X[j];
          dres_dKDj = dres_dKDj + cation_X[j] / KDj * cfg->primary_spec_Z[icomp];
        }
        dres_dKDj = dres_dKDj / ref_cation_Z + (ref_cation_k * ref_cation_conc);
        res = 1.0 - total;
        if (one_more) break;
        delta_KDj = res / dres_dKDj;
        KDj = KDj + delta_KDj;
        KDj = fmax(KDj, 1.e-40);
        if (fabs(delta_KDj / KDj) < tol) one_more = 1;
      }
      c->eqionx_ref[irxn] = ref_cation_X * omega / ref_cation_Z;
    } else {
      double sumkm = 0.0;
      for (j = 0; j < ncomp; j++) {
        int icomp = cat[j];
        cation_X[j] = c->pri_molal[icomp] * c->pri_act_coef[icomp] * kk[j];
        sumkm = sumkm + cation_X[j];
      }
      for (j = 0; j < ncomp; j++) cation_X[j] = cation_X[j] / sumkm;
    }
    sumZX = 0.0;
    for (i = 0; i < ncomp; i++) sumZX = sumZX + cfg->primary_spec_Z[cat[i]] * cation_X[i];
    for (i = 0; i < ncomp; i++) {
      int icomp = cat[i];
      double tempreal1 = cation_X[i] * omega / cfg->primary_spec_Z[icomp];
      double tempreal2;
      c->eqionx_conc[p0 + i] = c->eqionx_conc[p0 + i] + tempreal1;
      c->total_sorb_eq[icomp] = c->total_sorb_eq[icomp] + tempreal1;
      tempreal2 = cfg->primary_spec_Z[icomp] / sumZX;
      for (j = 0; j < ncomp; j++) {
        int jcomp = cat[j];
        if (i == j)
          c->dtotal_sorb_eq[icomp + jcomp * naq] =
              c->dtotal_sorb_eq[icomp + jcomp * naq] + tempreal1 * (1.0 - (tempreal2 * cation_X[j])) / c->pri_molal[jcomp];
        else
          c->dtotal_sorb_eq[icomp + jcomp * naq] =
              c->dtotal_sorb_eq[icomp + jcomp * naq] + (-tempreal1) * tempreal2 * cation_X[j] / c->pri_molal[jcomp];
      }
    }
  }
  return 0;
}

/* reaction.F90:4836-4902  RTotalSorbDynamicKD */
static void r_total_sorb_dynamic_kd(cell_t *c, const pfrx_config *cfg) {
  const double Lwater_m3bulk = 250.0;
  int naq = c->naq, irxn;
  for (irxn = 0; irxn < cfg->neqdynamickdrxn; irxn++) {
    int ikd = cfg->eqdynamickd_specid[irxn], iref = cfg->eqdynamickd_refspecid[irxn];
    double kd_species_molality = c->pri_molal[ikd];
    double ref_high = cfg->eqdynamickd_refspechigh[irxn];
    double ref_species_molality = c->pri_molal[iref];
    double KD_power = cfg->eqdynamickd_power[irxn];
    double KD_low = cfg->eqdynamickd_low[irxn];
    double KD_high_minus_low = cfg->eqdynamickd_high[irxn] - KD_low;
    double tempreal = pow(ref_species_molality / ref_high, KD_power);
    double KD = KD_low + tempreal * KD_high_minus_low;
    double dKD_dref = KD_power * tempreal / ref_species_molality * KD_high_minus_low;
    double total_sorb = KD * kd_species_molality * Lwater_m3bulk;
    double dtotal_sorb_dckd = KD * Lwater_m3bulk;
    double dtotal_sorb_dcref = dKD_dref * kd_species_molality * Lwater_m3bulk;
    c->total_sorb_eq[ikd] = c->total_sorb_eq[ikd] + total_sorb;
    c->dtotal_sorb_eq[ikd + ikd * naq] = c->dtotal_sorb_eq[ikd + ikd * naq] + dtotal_sorb_dckd;
    c->dtotal_sorb_eq[ikd + iref * naq] = c->dtotal_sorb_eq[ikd + iref * naq] + dtotal_sorb_dcref;
  }
}

/* reaction_isotherm.F90:273-359  RTotalSorbKD */
static void r_total_sorb_kd(cell_t *c, const pfrx_config *cfg) {
  int naq = c->naq, irxn;
  for (irxn = 0; irxn < cfg->neqkdrxn; irxn++) {
    int icomp = cfg->eqkd_specid[irxn];
    double molality = c->pri_molal[icomp];
    double kd_kgw_m3b, res, dres_dc, tempreal, one_over_n;
    if (cfg->ikd_units == 1)
      kd_kgw_m3b = cfg->eqkd_coeff[irxn] * c->den_kg * (1.0 - c->porosity) * c->soil_particle_density * 1.e-3;
    else
      kd_kgw_m3b = cfg->eqkd_coeff[irxn];
    if (cfg->eqkd_mineral[irxn] >= 0) kd_kgw_m3b = kd_kgw_m3b * (c->mnrl_volfrac[cfg->eqkd_mineral[irxn]]);
    switch (cfg->eqkd_type[irxn]) {
      case PFRX_SORPTION_LINEAR:
        res = kd_kgw_m3b * molality;
        dres_dc = kd_kgw_m3b;
        break;
      case PFRX_SORPTION_LANGMUIR:
        tempreal = kd_kgw_m3b * molality;
        res = tempreal * cfg->eqkd_langmuir_b[irxn] / (1.0 + tempreal);
        dres_dc = res / molality - res / (1.0 + tempreal) * tempreal / molality;
        break;
      case PFRX_SORPTION_FREUNDLICH:
        one_over_n = 1.0 / cfg->eqkd_freundlich_n[irxn];
        res = kd_kgw_m3b * pow(molality, one_over_n);
        dres_dc = res / molality * one_over_n;
        break;
      default:
        res = 0.0;
        dres_dc = 0.0;
    }
    c->total_sorb_eq[icomp] = c->total_sorb_eq[icomp] + res;
    c->dtotal_sorb_eq[icomp + icomp * naq] = c->dtotal_sorb_eq[icomp + icomp * naq] + dres_dc;
  }
}

/* reaction.F90:4783-4832 RTotalSorb (+RZeroSorb :4765); surface complexation
 * branch reaction_surf_complex.F90:446-487 */
static void r_total_sorb(cell_t *c, const pfrx_config *cfg) {
  int i, ieq;
  for (i = 0; i < c->naq; i++) c->total_sorb_eq[i] = 0.0;
  for (i = 0; i < c->naq * c->naq; i++) c->dtotal_sorb_eq[i] = 0.0;
  for (i = 0; i < c->nsrfcplx; i++) c->eqsrfcplx_conc[i] = 0.0;
  for (ieq = 0; ieq < cfg->neqsrfcplxrxn; ieq++) {
    int irxn = cfg->eqsrfcplxrxn_to_srfcplxrxn[ieq];
    r_total_sorb_eq_surf_cplx1(c, cfg, irxn, &c->free_site[irxn], c->eqsrfcplx_conc, c->total_sorb_eq,
                               c->dtotal_sorb_eq);
  }
  /* RTotalSorb order (reaction.F90:4783-4835): surface complexation, ion exchange, dynamic KD, KD */
  if (cfg->neqionxrxn > 0 && r_total_sorb_eq_ionx(c, cfg)) c->option_ierror = 1;
  if (cfg->neqdynamickdrxn > 0) r_total_sorb_dynamic_kd(c, cfg);
  if (cfg->neqkdrxn > 0) r_total_sorb_kd(c, cfg);
}

/* reaction_gas.F90:304-323 RGasConcentration [mol/m^3] */
static double r_gas_concentration(double gas_pp, double temperature) {
  return gas_pp * 1.e5 / (IDEAL_GAS_CONSTANT * (temperature + 273.15));
}

/* reaction_gas.F90:87-174 RTotalGas */
static void r_total_gas(cell_t *c, const pfrx_config *cfg) {
  int i, j, igas, icomp, jcomp, naq = c->naq;
  double ln_conc[PFRX_MAX_NCOMP * 4], ln_act[PFRX_MAX_NCOMP * 4];
  double lnQK, tempreal, gas_concentration;
  for (i = 0; i < naq; i++) c->total_gas[i] = 0.0;
  for (i = 0; i < naq; i++) {
    ln_conc[i] = log(c->pri_molal[i]);
    ln_act[i] = ln_conc[i] + log(c->pri_act_coef[i]);
  }
  for (i = 0; i < naq * naq; i++) c->dtotal_gas[i] = 0.0;
  for (igas = 0; igas < c->ngas; igas++) {
    int p0 = cfg->acteq_ptr[igas], p1 = cfg->acteq_ptr[igas + 1];
    lnQK = -c->acteq_logK[igas] * LOG_TO_LN;
    if (cfg->acteq_h2ostoich[igas] != 0.0) lnQK = lnQK + cfg->acteq_h2ostoich[igas] * c->ln_act_h2o;
    for (i = p0; i < p1; i++) {
      icomp = cfg->acteq_specid[i];
      lnQK = lnQK + cfg->acteq_stoich[i] * ln_act[icomp];
    }
    c->gas_pp[igas] = exp(lnQK);
    gas_concentration = r_gas_concentration(c->gas_pp[igas], c->temp) * 1.e-3;
    for (i = p0; i < p1; i++) {
      icomp = cfg->acteq_specid[i];
      c->total_gas[icomp] = c->total_gas[icomp] + cfg->acteq_stoich[i] * gas_concentration;
    }
    for (j = p0; j < p1; j++) {
      jcomp = cfg->acteq_specid[j];
      tempreal = cfg->acteq_stoich[j] * r_gas_concentration(exp(lnQK - ln_conc[jcomp]), c->temp) * 1.e-3;
      for (i = p0; i < p1; i++) {
        icomp = cfg->acteq_specid[i];
        c->dtotal_gas[icomp + jcomp * naq] = c->dtotal_gas[icomp + jcomp * naq] + cfg->acteq_stoich[i] * tempreal;
      }
    }
  }
}

/* reaction.F90:4618-4661 RTotal == reaction.F90:5606 RTAuxVarCompute */
static void rt_auxvar_compute(cell_t *c, const pfrx_config *cfg) {
  int i;
  for (i = 0; i < c->naq; i++) c->total[i] = 0.0;
  if (c->naq > 0) r_total_aqueous(c, cfg);
  if (neqsorb(cfg) > 0) r_total_sorb(c, cfg);
  if (c->ngas > 0) r_total_gas(c, cfg);
}

/* reaction.F90:5710-5771 RTAccumulation */
static void rt_accumulation(const cell_t *c, const pfrx_config *cfg, double *Res) {
  int i;
  double psv_t;
  for (i = 0; i < c->n; i++) Res[i] = 0.0;
  if (c->sat < cfg->rt_min_saturation) return;
  psv_t = c->porosity * c->sat * 1000.0 * c->volume;
  for (i = 0; i < c->naq; i++) Res[i] = psv_t * c->total[i];
  for (i = 0; i < c->nim; i++) Res[c->naq + i] = Res[c->naq + i] + c->immobile[i] * c->volume;
  if (c->ngas > 0) { /* :5761-5769 */
    psv_t = c->porosity * c->sat_gas * 1000.0 * c->volume;
    for (i = 0; i < c->naq; i++) Res[i] = Res[i] + psv_t * c->total_gas[i];
  }
}

/* reaction.F90:5775-5848 RTAccumulationDerivative; J(i,j) at [i + j*n] */
static void rt_accumulation_derivative(const cell_t *c, const pfrx_config *cfg, double tran_dt, double *J) {
  int i, j, n = c->n, naq = c->naq;
  double psvd_t;
  for (i = 0; i < n * n; i++) J[i] = 0.0;
  if (c->sat < cfg->rt_min_saturation) {
    for (i = 0; i < n; i++) J[i + i * n] = 1.0;
    return;
  }
  psvd_t = c->porosity * c->sat * 1000.0 * c->volume / tran_dt;
  for (j = 0; j < naq; j++)
    for (i = 0; i < naq; i++) J[i + j * n] = c->dtotal[i + j * naq] * psvd_t;
  for (i = 0; i < c->nim; i++) J[(naq + i) + (naq + i) * n] = c->volume / tran_dt;
  if (c->ngas > 0) { /* :5838-5846 */
    psvd_t = c->porosity * c->sat_gas * 1000.0 * c->volume / tran_dt;
    for (j = 0; j < naq; j++)
      for (i = 0; i < naq; i++) J[i + j * n] = J[i + j * n] + c->dtotal_gas[i + j * naq] * psvd_t;
  }
}

/* reaction.F90:5144-5172 / :5177-5207 */
static void r_accumulation_sorb(const cell_t *c, double *Res) {
  int i;
  for (i = 0; i < c->naq; i++) Res[i] = Res[i] + c->total_sorb_eq[i] * c->volume;
}
static void r_accumulation_sorb_derivative(const cell_t *c, double tran_dt, double *J) {
  int i, j, n = c->n, naq = c->naq;
  double v_t = c->volume / tran_dt;
  for (j = 0; j < naq; j++)
    for (i = 0; i < naq; i++) J[i + j * n] = J[i + j * n] + c->dtotal_sorb_eq[i + j * naq] * v_t;
}

/* ------------------------------------------------------------------------ */
/* reaction_mineral.F90:647-1078  RKineticMineral (non SOLID_SOLUTION build) */
static void r_kinetic_mineral(cell_t *c, const pfrx_config *cfg, double *Res, double *Jac, int compute_derivative) {
  int i, j, imnrl, icomp, jcomp, n = c->n, naq = c->naq;
  int ipref, ipref_species;
  double tempreal, affinity_factor, sign_, Im, Im_const, dIm_dQK;
  double ln_conc[PFRX_MAX_NCOMP * 4], ln_act[PFRX_MAX_NCOMP * 4];
  double *ln_sec_act = NULL;
  double QK, lnQK, dQK_dCj, dQK_dmj, den;
  double ln_spec_act, spec_act_coef, ln_prefactor, ln_numerator, ln_denominator;
  double prefactor[PFRX_MAX_PREFACTORS];
  double ln_prefactor_spec[PFRX_MAX_PREFACTORS][PFRX_MAX_PREFACTOR_SPECIES];
  double sum_prefactor_rate = 0.0, dIm_dsum_prefactor_rate, dIm_dspec;
  double dprefactor_dprefactor_spec, dprefactor_spec_dspec;
  double dprefactor_spec_dspec_numerator, dprefactor_spec_dspec_denominator;
  double denominator, ln_gam_m_beta, arrhenius_factor;
  const int MAXP = PFRX_MAX_PREFACTORS, MAXS = PFRX_MAX_PREFACTOR_SPECIES;

  for (i = 0; i < naq; i++) {
    ln_conc[i] = log(c->pri_molal[i]);
    ln_act[i] = ln_conc[i] + log(c->pri_act_coef[i]);
  }
  if (c->ncplx > 0 && cfg->kinmnrl_num_prefactors) {
    ln_sec_act = (double *)malloc(sizeof(double) * c->ncplx);
    for (i = 0; i < c->ncplx; i++) ln_sec_act[i] = log(c->sec_molal[i]) + log(c->sec_act_coef[i]);
  }
  for (imnrl = 0; imnrl < c->nkin; imnrl++) c->mnrl_rate[imnrl] = 0.0;

  for (imnrl = 0; imnrl < c->nkin; imnrl++) {
    int p0 = cfg->kinmnrl_ptr[imnrl], p1 = cfg->kinmnrl_ptr[imnrl + 1];
    int nprefactors = cfg->kinmnrl_num_prefactors ? cfg->kinmnrl_num_prefactors[imnrl] : 0;
    lnQK = -c->kinmnrl_logK[imnrl] * LOG_TO_LN;
    if (cfg->kinmnrl_h2ostoich[imnrl] != 0.0) lnQK = lnQK + cfg->kinmnrl_h2ostoich[imnrl] * c->ln_act_h2o;
    for (i = p0; i < p1; i++) {
      icomp = cfg->kinmnrl_specid[i];
      lnQK = lnQK + cfg->kinmnrl_stoich[i] * ln_act[icomp];
    }
    QK = exp(lnQK);

    if (cfg->kinmnrl_Temkin_const) {
      if (cfg->kinmnrl_min_scale_factor)
        affinity_factor =
            1.0 - pow(QK, 1.0 / (cfg->kinmnrl_min_scale_factor[imnrl] * cfg->kinmnrl_Temkin_const[imnrl]));
      else
        affinity_factor = 1.0 - pow(QK, 1.0 / cfg->kinmnrl_Temkin_const[imnrl]);
    } else if (cfg->kinmnrl_min_scale_factor) {
      affinity_factor = 1.0 - pow(QK, 1.0 / cfg->kinmnrl_min_scale_factor[imnrl]);
    } else {
      affinity_factor = 1.0 - QK;
    }
    sign_ = copysign(1.0, affinity_factor);

    if (c->mnrl_volfrac[imnrl] > 0 || sign_ < 0.0) {
      if (cfg->kinmnrl_irreversible[imnrl] == 1 && sign_ < 0.0) continue;
      if (cfg->kinmnrl_affinity_threshold[imnrl] > 0.0) {
        if (sign_ < 0.0 && QK < cfg->kinmnrl_affinity_threshold[imnrl]) continue;
      }
      if (cfg->kinmnrl_rate_limiter[imnrl] > 0.0) {
        affinity_factor = affinity_factor / (1.0 + (1.0 - affinity_factor) / cfg->kinmnrl_rate_limiter[imnrl]);
      }
      if (nprefactors > 0) {
        sum_prefactor_rate = 0.0;
        memset(prefactor, 0, sizeof(prefactor));
        memset(ln_prefactor_spec, 0, sizeof(ln_prefactor_spec));
        for (ipref = 0; ipref < nprefactors; ipref++) {
          int nps = cfg->kinmnrl_pref_nspec[imnrl * MAXP + ipref];
          double eact = cfg->kinmnrl_pref_activation_energy[imnrl * MAXP + ipref];
          ln_prefactor = 0.0;
          for (ipref_species = 0; ipref_species < nps; ipref_species++) {
            int q = (imnrl * MAXP + ipref) * MAXS + ipref_species;
            icomp = cfg->kinmnrl_prefactor_id[q];
            if (icomp >= 0)
              ln_spec_act = ln_act[icomp];
            else
              ln_spec_act = ln_sec_act[-icomp - 1];
            ln_numerator = cfg->kinmnrl_pref_alpha[q] * ln_spec_act;
            ln_denominator =
                log(1.0 + exp(log(cfg->kinmnrl_pref_atten_coef[q]) + cfg->kinmnrl_pref_beta[q] * ln_spec_act));
            ln_prefactor = ln_prefactor + ln_numerator;
            ln_prefactor = ln_prefactor - ln_denominator;
            ln_prefactor_spec[ipref][ipref_species] = ln_numerator - ln_denominator;
          }
          prefactor[ipref] = exp(ln_prefactor);
          arrhenius_factor = 1.0;
          if (eact > 0.0)
            arrhenius_factor = exp(eact / IDEAL_GAS_CONSTANT * (1.0 / (25.0 + 273.15) - 1.0 / (c->temp + 273.15)));
          sum_prefactor_rate =
              sum_prefactor_rate + prefactor[ipref] * cfg->kinmnrl_pref_rate[imnrl * MAXP + ipref] * arrhenius_factor;
        }
      } else {
        arrhenius_factor = 1.0;
        if (cfg->kinmnrl_activation_energy[imnrl] > 0.0)
          arrhenius_factor = exp(cfg->kinmnrl_activation_energy[imnrl] / IDEAL_GAS_CONSTANT *
                                 (1.0 / (25.0 + 273.15) - 1.0 / (c->temp + 273.15)));
        sum_prefactor_rate = cfg->kinmnrl_rate_constant[imnrl] * arrhenius_factor;
      }
      Im_const = -c->mnrl_area[imnrl];
      if (cfg->kinmnrl_min_scale_factor) Im_const = Im_const / cfg->kinmnrl_min_scale_factor[imnrl];
      if (cfg->kinmnrl_affinity_power)
        Im = Im_const * sign_ * pow(fabs(affinity_factor), cfg->kinmnrl_affinity_power[imnrl]) * sum_prefactor_rate;
      else
        Im = Im_const * sign_ * fabs(affinity_factor) * sum_prefactor_rate;
      c->mnrl_rate[imnrl] = Im;
    } else {
      continue;
    }

    Im_const = Im_const * c->volume;
    Im = Im * c->volume;
    for (i = p0; i < p1; i++) {
      icomp = cfg->kinmnrl_specid[i];
      Res[icomp] = Res[icomp] + cfg->kinmnrl_stoich[i] * Im;
    }
    if (!compute_derivative) continue;

    if (cfg->kinmnrl_affinity_power)
      dIm_dQK = -Im * cfg->kinmnrl_affinity_power[imnrl] / fabs(affinity_factor);
    else
      dIm_dQK = -Im_const * sum_prefactor_rate;

    if (cfg->kinmnrl_Temkin_const) {
      if (cfg->kinmnrl_min_scale_factor)
        dIm_dQK = dIm_dQK * (1.0 / (cfg->kinmnrl_min_scale_factor[imnrl] * cfg->kinmnrl_Temkin_const[imnrl])) / QK *
                  (1.0 - affinity_factor);
      else
        dIm_dQK = dIm_dQK * (1.0 / cfg->kinmnrl_Temkin_const[imnrl]) / QK * (1.0 - affinity_factor);
    } else if (cfg->kinmnrl_min_scale_factor) {
      dIm_dQK = dIm_dQK * (1.0 / cfg->kinmnrl_min_scale_factor[imnrl]) / QK * (1.0 - affinity_factor);
    }

    if (cfg->kinmnrl_rate_limiter[imnrl] <= 0.0) {
      for (j = p0; j < p1; j++) {
        jcomp = cfg->kinmnrl_specid[j];
        dQK_dCj = cfg->kinmnrl_stoich[j] * QK * exp(-ln_conc[jcomp]);
        dQK_dmj = dQK_dCj * c->den_kg * 1.e-3;
        for (i = p0; i < p1; i++) {
          icomp = cfg->kinmnrl_specid[i];
          Jac[icomp + jcomp * n] = Jac[icomp + jcomp * n] + cfg->kinmnrl_stoich[i] * dIm_dQK * dQK_dmj;
        }
      }
    } else {
      den = 1.0 + (1.0 - affinity_factor) / cfg->kinmnrl_rate_limiter[imnrl];
      for (j = p0; j < p1; j++) {
        jcomp = cfg->kinmnrl_specid[j];
        dQK_dCj = cfg->kinmnrl_stoich[j] * QK * exp(-ln_conc[jcomp]);
        dQK_dmj = dQK_dCj * c->den_kg * 1.e-3;
        for (i = p0; i < p1; i++) {
          icomp = cfg->kinmnrl_specid[i];
          Jac[icomp + jcomp * n] =
              Jac[icomp + jcomp * n] + cfg->kinmnrl_stoich[i] * dIm_dQK *
                                           (1.0 + QK / cfg->kinmnrl_rate_limiter[imnrl] / den) * dQK_dmj / den;
        }
      }
    }

    if (nprefactors > 0) {
      dIm_dsum_prefactor_rate = Im / sum_prefactor_rate;
      for (ipref = 0; ipref < nprefactors; ipref++) {
        int nps = cfg->kinmnrl_pref_nspec[imnrl * MAXP + ipref];
        double eact = cfg->kinmnrl_pref_activation_energy[imnrl * MAXP + ipref];
        arrhenius_factor = 1.0;
        if (eact > 0.0)
          arrhenius_factor = exp(eact / IDEAL_GAS_CONSTANT * (1.0 / (25.0 + 273.15) - 1.0 / (c->temp + 273.15)));
        ln_prefactor = log(prefactor[ipref]);
        for (ipref_species = 0; ipref_species < nps; ipref_species++) {
          int q = (imnrl * MAXP + ipref) * MAXS + ipref_species;
          dprefactor_dprefactor_spec = exp(ln_prefactor - ln_prefactor_spec[ipref][ipref_species]);
          icomp = cfg->kinmnrl_prefactor_id[q];
          if (icomp >= 0) {
            ln_spec_act = ln_act[icomp];
            spec_act_coef = c->pri_act_coef[icomp];
          } else {
            ln_spec_act = ln_sec_act[-icomp - 1];
            spec_act_coef = c->sec_act_coef[-icomp - 1];
          }
          dprefactor_spec_dspec_numerator =
              cfg->kinmnrl_pref_alpha[q] * exp(ln_prefactor_spec[ipref][ipref_species] - ln_spec_act);
          ln_gam_m_beta = cfg->kinmnrl_pref_beta[q] * ln_spec_act;
          denominator = 1.0 + exp(log(cfg->kinmnrl_pref_atten_coef[q]) + ln_gam_m_beta);
          dprefactor_spec_dspec_denominator = -1.0 * exp(ln_prefactor_spec[ipref][ipref_species]) / denominator *
                                              cfg->kinmnrl_pref_atten_coef[q] * cfg->kinmnrl_pref_beta[q] *
                                              exp(ln_gam_m_beta - ln_spec_act);
          dprefactor_spec_dspec = dprefactor_spec_dspec_numerator + dprefactor_spec_dspec_denominator;
          dprefactor_spec_dspec = dprefactor_spec_dspec * spec_act_coef;
          dIm_dspec = dIm_dsum_prefactor_rate * dprefactor_dprefactor_spec * dprefactor_spec_dspec *
                      cfg->kinmnrl_pref_rate[imnrl * MAXP + ipref] * arrhenius_factor;
          if (icomp >= 0) {
            for (i = p0; i < p1; i++) {
              jcomp = cfg->kinmnrl_specid[i];
              Jac[jcomp + icomp * n] = Jac[jcomp + icomp * n] + cfg->kinmnrl_stoich[i] * dIm_dspec;
            }
          } else {
            /* secondary species: reference recomputes lnQK of the complex and
             * (reaction_mineral.F90:1055) clobbers `ncomp`, so its i-loop runs
             * over the COMPLEX's species, not the mineral's.  Restated as is. */
            int icplx = -icomp - 1;
            int q0 = cfg->eqcplx_ptr[icplx], q1 = cfg->eqcplx_ptr[icplx + 1];
            int ii, jj;
            lnQK = -c->eqcplx_logK[icplx] * LOG_TO_LN;
            if (cfg->eqcplx_h2ostoich[icplx] != 0.0) lnQK = lnQK + cfg->eqcplx_h2ostoich[icplx] * c->ln_act_h2o;
            for (ii = q0; ii < q1; ii++) lnQK = lnQK + cfg->eqcplx_stoich[ii] * ln_act[cfg->eqcplx_specid[ii]];
            for (jj = q0; jj < q1; jj++) {
              jcomp = cfg->eqcplx_specid[jj];
              tempreal = cfg->eqcplx_stoich[jj] * exp(lnQK - ln_conc[jcomp]) / c->sec_act_coef[icplx];
              for (ii = q0; ii < q1; ii++) {
                int ic2 = cfg->eqcplx_specid[ii];
                Jac[ic2 + jcomp * n] = Jac[ic2 + jcomp * n] + cfg->eqcplx_stoich[ii] * tempreal * dIm_dspec;
              }
            }
          }
        }
      }
    }
  }
  free(ln_sec_act);
}

/* reaction_surf_complex.F90:552-637  RMultiRateSorption */
static void r_multirate_sorption(cell_t *c, const pfrx_config *cfg, double tran_dt, double *Res, double *Jac,
                                 int compute_derivative) {
  int naq = c->naq, n = c->n, i, j, ikinmrrxn, irate;
  double total_sorb_eq[PFRX_MAX_NCOMP * 4];
  double *dtotal_sorb_eq = (double *)malloc(sizeof(double) * naq * naq);
  for (ikinmrrxn = 0; ikinmrrxn < cfg->nkinmrsrfcplxrxn; ikinmrrxn++) {
    int base = naq * (cfg->kinmr_rate_ptr[ikinmrrxn] + ikinmrrxn);
    for (i = 0; i < naq; i++) c->kinmr_total_sorb[base + i] = 0.0;
  }
  for (ikinmrrxn = 0; ikinmrrxn < cfg->nkinmrsrfcplxrxn; ikinmrrxn++) {
    int irxn = cfg->kinmrsrfcplxrxn_to_srfcplxrxn[ikinmrrxn];
    int r0 = cfg->kinmr_rate_ptr[ikinmrrxn], r1 = cfg->kinmr_rate_ptr[ikinmrrxn + 1];
    int base = naq * (r0 + ikinmrrxn);
    for (i = 0; i < naq; i++) total_sorb_eq[i] = 0.0;
    for (i = 0; i < naq * naq; i++) dtotal_sorb_eq[i] = 0.0;
    r_total_sorb_eq_surf_cplx1(c, cfg, irxn, &c->free_site[irxn], NULL, total_sorb_eq, dtotal_sorb_eq);
    for (irate = r0; irate < r1; irate++) {
      double kdt = cfg->kinmr_rate[irate] * tran_dt;
      double one_plus_kdt = 1.0 + kdt;
      double k_over_one_plus_kdt = cfg->kinmr_rate[irate] / one_plus_kdt;
      const double *S = c->kinmr_total_sorb + base + naq * (irate - r0 + 1);
      for (i = 0; i < naq; i++)
        Res[i] = Res[i] + c->volume * k_over_one_plus_kdt * (cfg->kinmr_frac[irate] * total_sorb_eq[i] - S[i]);
      if (compute_derivative) {
        for (j = 0; j < naq; j++)
          for (i = 0; i < naq; i++)
            Jac[i + j * n] =
                Jac[i + j * n] + c->volume * k_over_one_plus_kdt * cfg->kinmr_frac[irate] * dtotal_sorb_eq[i + j * naq];
      }
    }
    for (i = 0; i < naq; i++) c->kinmr_total_sorb[base + i] = total_sorb_eq[i];
  }
  free(dtotal_sorb_eq);
}

/* reaction_sandbox_clm_cn.F90:468-787  CLM_CN_React */
static void clm_cn_react(cell_t *c, const pfrx_config *cfg, double *Residual, double *Jacobian,
                         int compute_derivative) {
  int n = c->n, off = c->naq;
  int ipool_up, ipool_down, ispec_pool_down, ispecC_pool_up, ispecN_pool_up = -1;
  int ires_pool_down = -1, ires_C, ires_N, iresC_pool_up, iresN_pool_up = -1, ispec_N, irxn;
  double drate, scaled_rate_const, rate, F_t, F_theta, constant_inhibition, temp_K;
  const double one_over_71_02 = 1.408054069e-2;
  const double theta_min = 0.01;
  const double one_over_log_theta_min = -2.17147241e-1;
  double CN_ratio_up, CN_ratio_down, resp_frac, stoich_N, stoich_C;
  double stoich_downstreamC_pool, stoich_upstreamC_pool, stoich_upstreamN_pool;
  double N_inhibition, d_N_inhibition, drate_dN_inhibition = 0.0, temp_real;
  int constant_CN_ratio_up, use_N_inhibition;
#define JAC(i, j) Jacobian[(i) + (size_t)(j) * n]

  temp_K = c->temp + 273.15;
  if (temp_K > 227.15)
    F_t = exp(308.56 * (one_over_71_02 - 1.0 / (temp_K - 227.13)));
  else
    return;
  F_theta = log(theta_min / fmax(theta_min, c->sat)) * one_over_log_theta_min;
  constant_inhibition = F_t * F_theta;

  ires_C = off + cfg->clmcn_C_species_id;
  ispec_N = cfg->clmcn_N_species_id;
  ires_N = off + ispec_N;

  for (irxn = 0; irxn < cfg->clmcn_nrxn; irxn++) {
    scaled_rate_const = cfg->clmcn_rate_constant[irxn] * c->volume * constant_inhibition;
    resp_frac = cfg->clmcn_respiration_fraction[irxn];
    ipool_up = cfg->clmcn_upstream_pool_id[irxn];
    constant_CN_ratio_up = (cfg->clmcn_pool_nspec[ipool_up] == 1);
    if (!constant_CN_ratio_up) {
      ispecC_pool_up = cfg->clmcn_pool_C_id[ipool_up];
      ispecN_pool_up = cfg->clmcn_pool_N_id[ipool_up];
      CN_ratio_up = c->immobile[ispecC_pool_up] / c->immobile[ispecN_pool_up];
    } else {
      ispecC_pool_up = cfg->clmcn_pool_C_id[ipool_up];
      CN_ratio_up = cfg->clmcn_CN_ratio[ipool_up];
    }
    stoich_upstreamC_pool = 1.0;
    stoich_upstreamN_pool = stoich_upstreamC_pool / CN_ratio_up;

    ipool_down = cfg->clmcn_downstream_pool_id[irxn];
    if (ipool_down >= 0) {
      ispec_pool_down = cfg->clmcn_pool_C_id[ipool_down];
      CN_ratio_down = cfg->clmcn_CN_ratio[ipool_down];
      stoich_downstreamC_pool = (1.0 - resp_frac) * stoich_upstreamC_pool;
    } else {
      ispec_pool_down = -1;
      stoich_downstreamC_pool = 0.0;
      CN_ratio_down = 1.0;
    }
    stoich_C = resp_frac * stoich_upstreamC_pool;
    stoich_N = stoich_upstreamN_pool - stoich_downstreamC_pool / CN_ratio_down;

    if (cfg->clmcn_inhibition_constant[irxn] > 1.e-40 && stoich_N < 0.0) {
      use_N_inhibition = 1;
      temp_real = c->immobile[ispec_N] + cfg->clmcn_inhibition_constant[irxn];
      N_inhibition = c->immobile[ispec_N] / temp_real;
      d_N_inhibition = cfg->clmcn_inhibition_constant[irxn] / (temp_real * temp_real);
    } else {
      use_N_inhibition = 0;
      N_inhibition = 1.0;
      d_N_inhibition = 0.0;
    }

    rate = c->immobile[ispecC_pool_up] * scaled_rate_const * N_inhibition;

    Residual[ires_C] = Residual[ires_C] - stoich_C * rate;
    Residual[ires_N] = Residual[ires_N] - stoich_N * rate;
    iresC_pool_up = off + ispecC_pool_up;
    Residual[iresC_pool_up] = Residual[iresC_pool_up] - (-1.0) * stoich_upstreamC_pool * rate;
    if (!constant_CN_ratio_up) {
      iresN_pool_up = off + ispecN_pool_up;
      Residual[iresN_pool_up] = Residual[iresN_pool_up] - (-1.0) * stoich_upstreamN_pool * rate;
    }
    if (ispec_pool_down >= 0) {
      ires_pool_down = off + ispec_pool_down;
      Residual[ires_pool_down] = Residual[ires_pool_down] - stoich_downstreamC_pool * rate;
    }

    if (compute_derivative) {
      drate = scaled_rate_const * N_inhibition;
      JAC(iresC_pool_up, iresC_pool_up) = JAC(iresC_pool_up, iresC_pool_up) - (-1.0) * stoich_upstreamC_pool * drate;
      if (use_N_inhibition) {
        drate_dN_inhibition = c->immobile[ispecC_pool_up] * scaled_rate_const * d_N_inhibition;
        JAC(iresC_pool_up, ires_N) =
            JAC(iresC_pool_up, ires_N) - (-1.0) * stoich_upstreamC_pool * drate_dN_inhibition;
      }
      if (ispec_pool_down >= 0) {
        JAC(ires_pool_down, iresC_pool_up) = JAC(ires_pool_down, iresC_pool_up) - stoich_downstreamC_pool * drate;
        if (use_N_inhibition)
          JAC(ires_pool_down, ires_N) = JAC(ires_pool_down, ires_N) - stoich_downstreamC_pool * drate_dN_inhibition;
      }
      if (!constant_CN_ratio_up) {
        JAC(iresN_pool_up, iresC_pool_up) =
            JAC(iresN_pool_up, iresC_pool_up) - (-1.0) * stoich_upstreamN_pool * drate;
        if (use_N_inhibition)
          JAC(iresN_pool_up, ires_N) =
              JAC(iresN_pool_up, ires_N) - (-1.0) * stoich_upstreamN_pool * drate_dN_inhibition;
        JAC(iresN_pool_up, iresC_pool_up) =
            JAC(iresN_pool_up, iresC_pool_up) - (-1.0) * (-1.0) * c->immobile[ispecN_pool_up] /
                                                    c->immobile[ispecC_pool_up] * scaled_rate_const * N_inhibition;
        JAC(iresN_pool_up, iresN_pool_up) =
            JAC(iresN_pool_up, iresN_pool_up) - (-1.0) * scaled_rate_const * N_inhibition;
        JAC(ires_N, iresC_pool_up) = JAC(ires_N, iresC_pool_up) - (-1.0) * c->immobile[ispecN_pool_up] /
                                                                      c->immobile[ispecC_pool_up] *
                                                                      scaled_rate_const * N_inhibition;
        JAC(ires_N, iresN_pool_up) = JAC(ires_N, iresN_pool_up) - scaled_rate_const * N_inhibition;
      }
      JAC(ires_C, iresC_pool_up) = JAC(ires_C, iresC_pool_up) - stoich_C * drate;
      JAC(ires_N, iresC_pool_up) = JAC(ires_N, iresC_pool_up) - stoich_N * drate;
      if (use_N_inhibition) {
        JAC(ires_C, ires_N) = JAC(ires_C, ires_N) - stoich_C * drate_dN_inhibition;
        JAC(ires_N, ires_N) = JAC(ires_N, ires_N) - stoich_N * drate_dN_inhibition;
      }
    }
  }
#undef JAC
}


/* ------------------------------------------------------------------------ */
/* ELM-CN sandboxes                                                           */

/* utility.F90:2542-2597  HfunctionSmooth */
static void hfunction_smooth(double x, double x_1, double x_0, double *H, double *dH) {
  if (fabs(x_1 - x_0) < 1.e-50) {
    *H = copysign(0.5, (x - x_1)) + 0.5;
    *dH = 0.0;
    return;
  }
  if (((x - x_0) / (x_1 - x_0)) < 0.0) {
    *H = 0.0;
    *dH = 0.0;
  } else if (((x - x_0) / (x_1 - x_0)) > 1.0) {
    *H = 1.0;
    *dH = 0.0;
  } else {
    double x_star = 1.0 - (x - x_0) * (x - x_0) / (x_1 - x_0) / (x_1 - x_0);
    *H = 1.0 - x_star * x_star;
    *dH = 4.0 * x_star * (x - x_0) / (x_1 - x_0) / (x_1 - x_0);
  }
}

/* elm_rspfuncs.F90:321-338  FuncMonod */
static double func_monod(double conc, double monod_k, int compute_derivative) {
  if (!compute_derivative) return conc / (conc + monod_k);
  return monod_k / (conc + monod_k) / (conc + monod_k);
}

/* elm_rspfuncs.F90:342-375  FuncInhibition (PI = 3.14159265358979323846, pflotran_constants.F90) */
static double func_inhibition(int compute_derivative, double conc, double inhibition_C, int has_C2,
                              double inhibition_C2) {
  const double PI = 3.14159265358979323846;
  if (!has_C2) {
    if (compute_derivative) return -inhibition_C / (conc + inhibition_C) / (conc + inhibition_C);
    return inhibition_C / (conc + inhibition_C);
  }
  if (compute_derivative) {
    double tempreal = (conc - inhibition_C) * inhibition_C2;
    return (inhibition_C2 / (1.0 + tempreal * tempreal)) / PI;
  }
  return 0.5 + atan((conc - inhibition_C) * inhibition_C2) / PI;
}

/* elm_rspfuncs.F90:61-123  GetTemperatureResponse */
static double get_temperature_response(double tc, int itype, double Q10orEA) {
  const double one_over_71_02 = 1.408054069e-2;
  const double Frz_Q10 = 2.0;
  double Ft, tk;
  switch (itype) {
    case PFRX_TEMPERATURE_RESPONSE_Q10:
      if (tc > 0.0)
        Ft = pow(Q10orEA, (tc - 25.0) / 10.0);
      else
        Ft = pow(Q10orEA, -25.0 / 10.0) * pow(Frz_Q10, tc / 10.0);
      break;
    case PFRX_TEMPERATURE_RESPONSE_CLMCN:
      tk = tc + 273.15;
      if (tk > 227.15)
        Ft = exp(308.56 * (one_over_71_02 - 1.0 / (tk - 227.13)));
      else
        Ft = 0.0;
      break;
    case PFRX_TEMPERATURE_RESPONSE_DLEM:
      if (tc < -5.0)
        Ft = 0.0;
      else if (tc >= 30.0)
        Ft = 1.0;
      else
        Ft = pow(Q10orEA, (tc - 30.0) / 10.0);
      break;
    case PFRX_TEMPERATURE_RESPONSE_ARRHENIUS:
      Ft = exp(Q10orEA / IDEAL_GAS_CONSTANT * (1.0 / 298.15 - 1.0 / (tc + 273.15)));
      break;
    default:
      Ft = 1.0;
  }
  return Ft;
}

/* elm_rspfuncs.F90:124-237  GetMoistureResponse, ELM_PFLOTRAN build: theta is the volumetric water content */
static double elm_moisture_response(const cell_t *c, double theta, int itype) {
  const double minpsi = -10.0e6; /* Pa */
  const double EARTH_GRAVITY = 9.8068; /* pflotran_constants.F90:94 */
  double F_theta, maxpsi, psi, lsat, thetar, thetas, se;
  switch (itype) {
    case PFRX_MOISTURE_RESPONSE_CLMCN:
      maxpsi = c->elm_sucsat * (-EARTH_GRAVITY);
      lsat = theta / fmin(1.0, 1.0 - fmin(0.9999, c->elm_bd_dry / 2.70e3));
      psi = c->elm_sucsat * (-EARTH_GRAVITY) * pow(lsat, -c->elm_bsw);
      psi = fmin(psi, maxpsi);
      if (psi > minpsi) {
        F_theta = log(minpsi / psi) / log(minpsi / maxpsi);
        if (psi > (maxpsi - 1.0e02)) F_theta = F_theta * 0.10;
      } else {
        F_theta = 0.0;
      }
      break;
    case PFRX_MOISTURE_RESPONSE_DLEM:
      thetas = c->elm_effpor;
      thetar = c->elm_watfc;
      if (theta >= thetas) {
        F_theta = 1.0;
      } else if (theta <= thetar) {
        F_theta = 0.0;
      } else {
        se = (theta - thetar) / (thetas - thetar);
        /* "1.0 - se * se * 0.368 * exp(se)": default-kind literals, single precision */
        F_theta = (double)1.0f - se * se * (double)0.368f * exp(se);
        if (F_theta < 0.0) F_theta = 0.0;
        if (F_theta > 1.0) F_theta = 1.0;
      }
      break;
    default:
      F_theta = 1.0;
  }
  return F_theta;
}

/* elm_rspfuncs.F90:287-317  GetAerobicCondition */
static double get_aerobic_condition(double OXorWFPS, double K_Ox, int itype, int compute_derivative) {
  double F_Ox;
  switch (itype) {
    case PFRX_OX_RESPONSE_MONOD:
      F_Ox = func_monod(OXorWFPS, K_Ox, compute_derivative);
      break;
    case PFRX_OX_RESPONSE_WFPS:
      F_Ox = pow((1.27 - OXorWFPS) / 0.67, 3.1777) * pow((OXorWFPS - 0.0012) / 0.5988, 2.84);
      if (compute_derivative) F_Ox = 0.0;
      break;
    default:
      F_Ox = 1.0;
      if (compute_derivative) F_Ox = 0.0;
  }
  return F_Ox;
}

/* reaction_sandbox_somdec.F90:3809-3870  SomDec_SpeciesConc (total / immobile) */
static double somdec_species_conc(const cell_t *c, int id, int itype) {
  if (itype == PFRX_SPEC_AQUEOUS) return c->total[id];
  return c->immobile[id];
}

/* per-evaluation copy of the reference's `this%` scratch (see pfrx_somdec) */
typedef struct {
  double upstream_nc, mineral_c_stoich, mineral_n_stoich;
  double downstream_nc[16];
} somdec_scratch_t;

#define SD_RES(itype, id) ((itype) == PFRX_SPEC_AQUEOUS ? (id) : off + (id))
#define JAC(i, j) Jacobian[(i) + (size_t)(j) * n]
#define DTOT(i, j) c->dtotal[(i) + (size_t)(j) * c->naq]

/* the factors shared by SomDecReact1/2: MONOD list (other than the NH4/NO3
 * ones when `react2`), INHIBITION list, Ox Monod term.
 * reaction_sandbox_somdec.F90:2062-2150 and :2700-2800 */
static void somdec_rate_modifiers(const cell_t *c, const pfrx_somdec *sd, int rxn, int ispec_uc, int react2,
                                  double theta, double c_nh4, double c_no3, double *crate_uc,
                                  double *dcrate_uc_duc, double *fnh4, double *dfnh4_dnh4, double *fno3,
                                  double *dfno3_dno3) {
  double fmb = 1.0, dfmb = 0.0, fx, dfx, tempreal, f_ox, df_ox;
  int k;
  for (k = sd->monod_ptr[rxn]; k < sd->monod_ptr[rxn + 1]; k++) {
    double monod_k = sd->monod_half_saturation[k];
    double monod_threshold = sd->monod_threshold[k];
    if (react2 && sd->nh4_id >= 0 && sd->nh4_id == sd->monod_specid[k]) {
      tempreal = fmax(0.0, c_nh4 - monod_threshold);
      if (sd->monod_pool_normalized[k]) tempreal = tempreal / c->immobile[ispec_uc];
      *fnh4 = func_monod(tempreal, monod_k, 0);
      *dfnh4_dnh4 = func_monod(tempreal, monod_k, 1);
    } else if (react2 && sd->no3_id >= 0 && sd->no3_id == sd->monod_specid[k]) {
      tempreal = fmax(0.0, c_no3 - monod_threshold);
      if (sd->monod_pool_normalized[k]) tempreal = tempreal / c->immobile[ispec_uc];
      *fno3 = func_monod(tempreal, monod_k, 0);
      *dfno3_dno3 = func_monod(tempreal, monod_k, 1);
    } else {
      tempreal = somdec_species_conc(c, sd->monod_specid[k], sd->monod_specitype[k]);
      tempreal = fmax(0.0, tempreal - monod_threshold);
      if (sd->monod_pool_normalized[k]) {
        tempreal = tempreal / c->immobile[ispec_uc];
        if (sd->monod_specitype[k] == PFRX_SPEC_AQUEOUS) {
          if (react2)
            tempreal = tempreal * theta * 1000.0;
          else
            tempreal = tempreal * c->porosity * c->sat * 1000.0;
        }
      }
      fx = func_monod(tempreal, monod_k, 0);
      dfx = func_monod(tempreal, monod_k, 1);
      if (ispec_uc != sd->monod_specid[k]) dfx = 0.0;
      dfmb = dfmb * fx + fmb * dfx;
      fmb = fmb * fx;
    }
  }
  for (k = sd->inhib_ptr[rxn]; k < sd->inhib_ptr[rxn + 1]; k++) {
    double inhibition_k = sd->inhib_constant[k], inhibition_k2 = sd->inhib_constant2[k];
    tempreal = somdec_species_conc(c, sd->inhib_specid[k], sd->inhib_specitype[k]);
    fx = 1.0;
    dfx = 0.0;
    if (inhibition_k2 == -999.0 || sd->inhib_itype[k] != PFRX_INHIBITION_THRESHOLD) {
      if (sd->inhib_itype[k] == PFRX_INHIBITION_MONOD) {
        fx = func_inhibition(0, tempreal, inhibition_k, 0, 0.0);
        dfx = func_inhibition(1, tempreal, inhibition_k, 0, 0.0);
      } else if (sd->inhib_itype[k] == PFRX_INHIBITION_INVERSE_MONOD) {
        fx = func_monod(tempreal, inhibition_k, 0);
        dfx = func_monod(tempreal, inhibition_k, 1);
      }
    } else {
      fx = func_inhibition(0, tempreal, inhibition_k, 1, inhibition_k2);
      dfx = func_inhibition(1, tempreal, inhibition_k, 1, inhibition_k2);
    }
    if (ispec_uc != sd->inhib_specid[k]) dfx = 0.0;
    dfmb = dfmb * fx + fmb * dfx;
    fmb = fmb * fx;
  }
  *dcrate_uc_duc = *dcrate_uc_duc * fmb + *crate_uc * dfmb;
  *crate_uc = *crate_uc * fmb;

  if (sd->ox_response_function[rxn] == PFRX_OX_RESPONSE_MONOD && sd->ox_specid[rxn] >= 0) {
    double Ox = somdec_species_conc(c, sd->ox_specid[rxn], sd->ox_specitype[rxn]);
    f_ox = get_aerobic_condition(Ox, sd->ox_half_saturation[rxn], sd->ox_response_function[rxn], 0);
    df_ox = get_aerobic_condition(Ox, sd->ox_half_saturation[rxn], sd->ox_response_function[rxn], 1);
  } else {
    f_ox = 1.0;
    df_ox = 0.0;
  }
  *dcrate_uc_duc = *dcrate_uc_duc * f_ox + *crate_uc * df_ox;
  *crate_uc = *crate_uc * f_ox;
}

/* residual entries common to SomDecReact1/2: upstream C, CO2 (+ trackers), O2,
 * downstream C, upstream N.  reaction_sandbox_somdec.F90:2160-2215, :2880-2935 */
static void somdec_common_residual(const cell_t *c, const pfrx_somdec *sd, int irxn, const somdec_scratch_t *w,
                                   double crate, double *Residual, int *ires_ox_out) {
  int off = c->naq, j;
  int ires_uc = SD_RES(sd->upstream_is_aqueous[irxn] ? PFRX_SPEC_AQUEOUS : PFRX_SPEC_IMMOBILE, sd->upstream_c_id[irxn]);
  int ires_co2 = SD_RES(sd->co2_itype, sd->co2_id);
  Residual[ires_uc] = Residual[ires_uc] + crate;
  Residual[ires_co2] = Residual[ires_co2] - w->mineral_c_stoich * crate;
  if (sd->upstream_hr_id[irxn] >= 0)
    Residual[off + sd->upstream_hr_id[irxn]] = Residual[off + sd->upstream_hr_id[irxn]] - w->mineral_c_stoich * crate;
  if (sd->hr_id >= 0) Residual[off + sd->hr_id] = Residual[off + sd->hr_id] - w->mineral_c_stoich * crate;
  *ires_ox_out = -1;
  if (sd->o2_id >= 0) {
    int ires_ox = SD_RES(sd->o2_itype, sd->o2_id);
    Residual[ires_ox] = Residual[ires_ox] + w->mineral_c_stoich * crate;
    *ires_ox_out = ires_ox;
  }
  for (j = sd->downstream_ptr[irxn]; j < sd->downstream_ptr[irxn + 1]; j++) {
    int ispec_dc = sd->downstream_c_id[j];
    int ires_dc = sd->downstream_is_aqueous[j] ? ispec_dc : off + ispec_dc;
    if (ispec_dc >= 0) Residual[ires_dc] = Residual[ires_dc] - sd->downstream_stoich[j] * crate;
  }
  if (sd->upstream_n_id[irxn] >= 0) {
    int ires_un = sd->upstream_is_aqueous[irxn] ? sd->upstream_n_id[irxn] : off + sd->upstream_n_id[irxn];
    Residual[ires_un] = Residual[ires_un] + w->upstream_nc * crate;
  }
}

/* residual of variable-C:N downstream N.  :2238-2252, :2975-2990 */
static void somdec_downstream_n_residual(const cell_t *c, const pfrx_somdec *sd, int irxn, const somdec_scratch_t *w,
                                         double crate, double *Residual) {
  int off = c->naq, j;
  for (j = sd->downstream_ptr[irxn]; j < sd->downstream_ptr[irxn + 1]; j++) {
    int ispec_dn = sd->downstream_n_id[j];
    if (ispec_dn >= 0) {
      int ires_dn = sd->downstream_is_aqueous[j] ? ispec_dn : off + ispec_dn;
      Residual[ires_dn] =
          Residual[ires_dn] - sd->downstream_stoich[j] * crate * w->downstream_nc[j - sd->downstream_ptr[irxn]];
    }
  }
}

/* Jacobian column `jcol` (derivative w.r.t. species of primary id jaq, or an
 * immobile upstream pool when jaq < 0) of the entries every branch shares:
 * CO2 (+O2, trackers), upstream C, downstream C, upstream N, downstream N.
 * d*_dx follow the reference's names with x = uc / nh4 / no3. */
static void somdec_common_jacobian(const cell_t *c, const pfrx_somdec *sd, int irxn, const somdec_scratch_t *w,
                                   int jcol, int jaq, int ires_ox, double dco2_dx, double duc_dx, double dun_dx,
                                   int wrt_uc, double *Jacobian) {
  int off = c->naq, n = c->n, j;
  int up_aq = sd->upstream_is_aqueous[irxn];
  int ispec_uc = sd->upstream_c_id[irxn];
  int ires_uc = up_aq ? ispec_uc : off + ispec_uc;
  int ires_co2 = SD_RES(sd->co2_itype, sd->co2_id);
  /* the reference multiplies by dtotal when the column species is aqueous:
   * for x = uc only when the upstream pool is aqueous; for x = nh4/no3 on the
   * CO2 row always, on pool rows when the pool is aqueous */
  if (wrt_uc) {
    if (up_aq)
      JAC(ires_co2, jcol) = JAC(ires_co2, jcol) - dco2_dx * DTOT(sd->co2_id, ispec_uc);
    else
      JAC(ires_co2, jcol) = JAC(ires_co2, jcol) - dco2_dx;
    if (sd->o2_id >= 0) {
      if (up_aq)
        JAC(ires_ox, jcol) = JAC(ires_ox, jcol) + dco2_dx * DTOT(sd->co2_id, ispec_uc);
      else
        JAC(ires_ox, jcol) = JAC(ires_ox, jcol) + dco2_dx;
    }
  } else {
    JAC(ires_co2, jcol) = JAC(ires_co2, jcol) - dco2_dx * DTOT(sd->co2_id, jaq);
  }
  if (sd->upstream_hr_id[irxn] >= 0)
    JAC(off + sd->upstream_hr_id[irxn], jcol) = JAC(off + sd->upstream_hr_id[irxn], jcol) - dco2_dx;
  if (sd->hr_id >= 0) JAC(off + sd->hr_id, jcol) = JAC(off + sd->hr_id, jcol) - dco2_dx;

  if (up_aq)
    JAC(ires_uc, jcol) = JAC(ires_uc, jcol) - duc_dx * DTOT(ispec_uc, wrt_uc ? ispec_uc : jaq);
  else
    JAC(ires_uc, jcol) = JAC(ires_uc, jcol) - duc_dx;

  for (j = sd->downstream_ptr[irxn]; j < sd->downstream_ptr[irxn + 1]; j++) {
    int ispec_dc = sd->downstream_c_id[j];
    double ddc_dx = sd->downstream_stoich[j] * (-1.0 * duc_dx);
    if (wrt_uc) {
      int ires_dc = sd->downstream_is_aqueous[j] ? ispec_dc : off + ispec_dc;
      if (up_aq && sd->downstream_is_aqueous[j])
        JAC(ires_dc, jcol) = JAC(ires_dc, jcol) - ddc_dx * DTOT(ispec_dc, ispec_uc);
      else
        JAC(ires_dc, jcol) = JAC(ires_dc, jcol) - ddc_dx;
    } else {
      if (sd->downstream_is_aqueous[j])
        JAC(ispec_dc, jcol) = JAC(ispec_dc, jcol) - ddc_dx * DTOT(ispec_dc, jaq);
      else
        JAC(off + ispec_dc, jcol) = JAC(off + ispec_dc, jcol) - ddc_dx;
    }
  }
  (void)dun_dx;
}

/* upstream-N and downstream-N rows of column jcol (written after the NH4/NO3
 * rows in the reference, so kept separate to preserve the order of updates) */
static void somdec_n_rows_jacobian(const cell_t *c, const pfrx_somdec *sd, int irxn, const somdec_scratch_t *w,
                                   int jcol, int jaq, double duc_dx, double dun_dx, int wrt_uc, double *Jacobian) {
  int off = c->naq, n = c->n, j;
  int up_aq = sd->upstream_is_aqueous[irxn];
  int ispec_uc = sd->upstream_c_id[irxn];
  if (sd->upstream_n_id[irxn] >= 0) {
    int ispec_un = sd->upstream_n_id[irxn];
    int ires_un = up_aq ? ispec_un : off + ispec_un;
    if (up_aq)
      JAC(ires_un, jcol) = JAC(ires_un, jcol) - dun_dx * DTOT(ispec_un, wrt_uc ? ispec_uc : jaq);
    else
      JAC(ires_un, jcol) = JAC(ires_un, jcol) - dun_dx;
  }
  for (j = sd->downstream_ptr[irxn]; j < sd->downstream_ptr[irxn + 1]; j++) {
    int ispec_dn = sd->downstream_n_id[j];
    if (ispec_dn >= 0) {
      int ires_dn = sd->downstream_is_aqueous[j] ? ispec_dn : off + ispec_dn;
      double ddn_dx = sd->downstream_stoich[j] * (-1.0 * duc_dx) * w->downstream_nc[j - sd->downstream_ptr[irxn]];
      if (up_aq && sd->downstream_is_aqueous[j])
        JAC(ires_dn, jcol) = JAC(ires_dn, jcol) - ddn_dx * DTOT(ispec_dn, wrt_uc ? ispec_uc : jaq);
      else
        JAC(ires_dn, jcol) = JAC(ires_dn, jcol) - ddn_dx;
    }
  }
}

/* reaction_sandbox_somdec.F90:1914-2418  SomDecReact1 (N mineralisation type) */
static void somdec_react1(const cell_t *c, const pfrx_somdec *sd, int irxn, int rxn, const somdec_scratch_t *w,
                          double crate_uc, double dcrate_uc_duc, double *nmin, double *Residual, double *Jacobian,
                          int compute_derivative) {
  int off = c->naq, n = c->n;
  int ispec_uc = sd->upstream_c_id[irxn];
  int up_aq = sd->upstream_is_aqueous[irxn];
  int ires_uc = up_aq ? ispec_uc : off + ispec_uc;
  int ires_nh4 = sd->nh4_id, ires_ox;
  double crate, dummy1 = 1.0, dummy2 = 0.0, dummy3 = 1.0, dummy4 = 0.0;
  *nmin = 0.0;
  if (w->mineral_n_stoich < 0.0) return;
  somdec_rate_modifiers(c, sd, rxn, ispec_uc, 0, 0.0, 0.0, 0.0, &crate_uc, &dcrate_uc_duc, &dummy1, &dummy2, &dummy3,
                        &dummy4);
  crate = crate_uc;
  somdec_common_residual(c, sd, irxn, w, crate, Residual, &ires_ox);
  if (w->mineral_n_stoich >= 0.0) {
    Residual[ires_nh4] = Residual[ires_nh4] - w->mineral_n_stoich * crate;
    *nmin = w->mineral_n_stoich * crate;
    if (sd->upstream_nmin_id[irxn] >= 0)
      Residual[off + sd->upstream_nmin_id[irxn]] =
          Residual[off + sd->upstream_nmin_id[irxn]] - w->mineral_n_stoich * crate;
    if (sd->nmin_id >= 0) Residual[off + sd->nmin_id] = Residual[off + sd->nmin_id] - w->mineral_n_stoich * crate;
  }
  somdec_downstream_n_residual(c, sd, irxn, w, crate, Residual);

  if (compute_derivative) {
    double dcrate_dx = dcrate_uc_duc;
    double dco2_duc = dcrate_dx * w->mineral_c_stoich;
    double duc_duc = -1.0 * dcrate_dx;
    double dnh4_duc = dco2_duc * w->mineral_n_stoich; /* sic: reaction_sandbox_somdec.F90:2283 */
    double dun_duc = w->upstream_nc * duc_duc;
    somdec_common_jacobian(c, sd, irxn, w, ires_uc, -1, ires_ox, dco2_duc, duc_duc, dun_duc, 1, Jacobian);
    if (up_aq)
      JAC(ires_nh4, ires_uc) = JAC(ires_nh4, ires_uc) - dnh4_duc * DTOT(sd->nh4_id, ispec_uc);
    else
      JAC(ires_nh4, ires_uc) = JAC(ires_nh4, ires_uc) - dnh4_duc;
    if (w->mineral_n_stoich >= 0.0) {
      if (sd->upstream_nmin_id[irxn] >= 0)
        JAC(off + sd->upstream_nmin_id[irxn], ires_uc) = JAC(off + sd->upstream_nmin_id[irxn], ires_uc) - dnh4_duc;
      if (sd->nmin_id >= 0) JAC(off + sd->nmin_id, ires_uc) = JAC(off + sd->nmin_id, ires_uc) - dnh4_duc;
    }
    somdec_n_rows_jacobian(c, sd, irxn, w, ires_uc, -1, duc_duc, dun_duc, 1, Jacobian);
  }
}

/* reaction_sandbox_somdec.F90:2423-3472  SomDecReact2 (N immobilisation type) */
static void somdec_react2(const cell_t *c, const pfrx_somdec *sd, int irxn, int rxn, const somdec_scratch_t *w,
                          double tran_dt, double crate_uc, double dcrate_uc_duc, double *nimm, double *Residual,
                          double *Jacobian, int compute_derivative) {
  int off = c->naq, n = c->n;
  int ispec_uc = sd->upstream_c_id[irxn];
  int up_aq = sd->upstream_is_aqueous[irxn];
  int ires_uc = up_aq ? ispec_uc : off + ispec_uc;
  int ires_nh4 = sd->nh4_id, ires_no3 = sd->no3_id, ires_ox;
  double theta, volume, c_nh4 = 0.0, c_no3 = 0.0;
  double fnh4_inhibit_no3 = 1.0, dfnh4_inhibit_no3_dnh4 = 0.0, dfnh4_inhibit_no3_dno3 = 0.0;
  double fnh4 = 1.0, dfnh4_dnh4 = 0.0, fno3 = 1.0, dfno3_dno3 = 0.0;
  double feps0, dfeps0_dx, dtmin, nratecap, fnratecap, dfnratecap_dnh4, dfnratecap_dno3;
  double crate_nh4, crate_no3, crate, temp_real;
  double ns = w->mineral_n_stoich;
  *nimm = 0.0;
  if (w->mineral_n_stoich >= 0.0) return;
  theta = c->sat * c->porosity;
  volume = c->volume;
  if (sd->nh4_id >= 0) c_nh4 = c->total[sd->nh4_id] * theta * 1000.0;
  if (sd->no3_id >= 0) c_no3 = c->total[sd->no3_id] * theta * 1000.0;

  if (sd->inhibition_nh4_no3 > 0.0) {
    if (c_nh4 > sd->x0eps && c_no3 > sd->x0eps) {
      temp_real = c_nh4 / c_no3;
      fnh4_inhibit_no3 = func_monod(temp_real, 1.0 / sd->inhibition_nh4_no3, 0);
    } else {
      if (c_nh4 > sd->x0eps && c_no3 <= sd->x0eps)
        fnh4_inhibit_no3 = 1.0;
      else if (c_nh4 <= sd->x0eps && c_no3 > sd->x0eps)
        fnh4_inhibit_no3 = 0.0;
      else
        return;
    }
  }

  somdec_rate_modifiers(c, sd, rxn, ispec_uc, 1, theta, c_nh4, c_no3, &crate_uc, &dcrate_uc_duc, &fnh4, &dfnh4_dnh4,
                        &fno3, &dfno3_dno3);

  if (sd->nh4_id >= 0) {
    if (sd->x0eps > 0.0) {
      hfunction_smooth(c_nh4, sd->x0eps * 10.0, sd->x0eps, &feps0, &dfeps0_dx);
    } else {
      feps0 = 1.0;
      dfeps0_dx = 0.0;
    }
    dfnh4_dnh4 = dfnh4_dnh4 * feps0 + fnh4 * dfeps0_dx;
    fnh4 = fnh4 * feps0;
  }
  if (sd->no3_id >= 0) {
    if (sd->x0eps > 0.0) {
      hfunction_smooth(c_no3, sd->x0eps * 10.0, sd->x0eps, &feps0, &dfeps0_dx);
    } else {
      feps0 = 1.0;
      dfeps0_dx = 0.0;
    }
    dfno3_dno3 = dfno3_dno3 * feps0 + fno3 * dfeps0_dx;
    fno3 = fno3 * feps0;
  }

  dtmin = tran_dt;
  nratecap = -crate_uc * ns * dtmin / 0.45;
  if (sd->nh4_id >= 0) {
    if (nratecap * fnh4_inhibit_no3 > c_nh4 * volume) {
      fnratecap = func_monod(c_nh4 * volume, nratecap * fnh4_inhibit_no3 - c_nh4 * volume, 0);
      dfnratecap_dnh4 = func_monod(c_nh4 * volume, nratecap * fnh4_inhibit_no3 - c_nh4 * volume, 1);
    } else {
      fnratecap = 1.0;
      dfnratecap_dnh4 = 0.0;
    }
    dfnh4_dnh4 = dfnh4_dnh4 * fnratecap + fnh4 * dfnratecap_dnh4;
    fnh4 = fnh4 * fnratecap;
  }
  if (sd->no3_id >= 0) {
    if (nratecap * (1.0 - fnh4_inhibit_no3) > c_no3 * volume) {
      fnratecap = func_monod(c_no3 * volume, nratecap * (1.0 - fnh4_inhibit_no3) - c_no3 * volume, 0);
      dfnratecap_dno3 = func_monod(c_no3 * volume, nratecap * (1.0 - fnh4_inhibit_no3) - c_no3 * volume, 1);
    } else {
      fnratecap = 1.0;
      dfnratecap_dno3 = 0.0;
    }
    dfno3_dno3 = dfno3_dno3 * fnratecap + fno3 * dfnratecap_dno3;
    fno3 = fno3 * fnratecap;
  }

  crate_nh4 = crate_uc * fnh4 * fnh4_inhibit_no3;
  crate_no3 = crate_uc * fno3 * (1.0 - fnh4_inhibit_no3);
  crate = crate_nh4 + crate_no3;

  somdec_common_residual(c, sd, irxn, w, crate, Residual, &ires_ox);
  *nimm = 0.0;
  if (sd->nh4_id >= 0) {
    Residual[ires_nh4] = Residual[ires_nh4] - ns * crate_nh4;
    *nimm = *nimm + ns * crate_nh4;
  }
  if (sd->no3_id >= 0) {
    Residual[ires_no3] = Residual[ires_no3] - ns * crate_no3;
    *nimm = *nimm + ns * crate_no3;
  }
  if (sd->upstream_nimm_id[irxn] >= 0)
    Residual[off + sd->upstream_nimm_id[irxn]] = Residual[off + sd->upstream_nimm_id[irxn]] + ns * crate;
  if (sd->upstream_nimp_id[irxn] >= 0)
    Residual[off + sd->upstream_nimp_id[irxn]] = Residual[off + sd->upstream_nimp_id[irxn]] + ns * crate_uc;
  if (sd->nimm_id >= 0) Residual[off + sd->nimm_id] = Residual[off + sd->nimm_id] + ns * crate;
  if (sd->nimp_id >= 0) Residual[off + sd->nimp_id] = Residual[off + sd->nimp_id] + ns * crate_uc;
  somdec_downstream_n_residual(c, sd, irxn, w, crate, Residual);

  if (compute_derivative) {
    double dcrate_dx, dco2_duc, duc_duc, dnh4_duc, dno3_duc, dun_duc;
    double dco2_dnh4, duc_dnh4, dnh4_dnh4, dno3_dnh4, dun_dnh4;
    double dco2_dno3, duc_dno3, dnh4_dno3, dno3_dno3, dun_dno3;
    int unimm = sd->upstream_nimm_id[irxn];
    /* -- d/d(uc) :3000-3010 */
    dcrate_dx = dcrate_uc_duc * (fnh4 * fnh4_inhibit_no3 + fno3 - fno3 * fnh4_inhibit_no3);
    dco2_duc = dcrate_dx * w->mineral_c_stoich;
    duc_duc = -1.0 * dcrate_dx;
    dnh4_duc = dcrate_uc_duc * ns * fnh4 * fnh4_inhibit_no3;
    dno3_duc = dcrate_uc_duc * ns * fno3 * (1.0 - fnh4_inhibit_no3);
    dun_duc = w->upstream_nc * duc_duc;
    /* -- d/d(nh4) */
    dcrate_dx = (dfnh4_dnh4 * fnh4_inhibit_no3 + (fnh4 - fno3) * dfnh4_inhibit_no3_dnh4);
    dcrate_dx = dcrate_dx * crate_uc;
    dco2_dnh4 = dcrate_dx * w->mineral_c_stoich;
    duc_dnh4 = -1.0 * dcrate_dx;
    dnh4_dnh4 = fnh4 * dfnh4_inhibit_no3_dnh4 + dfnh4_dnh4 * fnh4_inhibit_no3;
    dnh4_dnh4 = dnh4_dnh4 * crate_uc * ns;
    dno3_dnh4 = -1.0 * fno3 * dfnh4_inhibit_no3_dnh4;
    dno3_dnh4 = dno3_dnh4 * crate_uc * ns;
    dun_dnh4 = w->upstream_nc * duc_dnh4;
    /* -- d/d(no3) */
    dcrate_dx = (fnh4 - fno3) * dfnh4_inhibit_no3_dno3 + dfno3_dno3 * (1.0 - fnh4 * fnh4_inhibit_no3);
    dcrate_dx = dcrate_dx * crate_uc;
    dco2_dno3 = dcrate_dx * w->mineral_c_stoich;
    duc_dno3 = -1.0 * dcrate_dx;
    dnh4_dno3 = fnh4 * dfnh4_inhibit_no3_dno3 * crate_uc * ns;
    dno3_dno3 = -1.0 * fno3 * dfnh4_inhibit_no3_dno3 + dfno3_dno3 * (1.0 - fnh4_inhibit_no3);
    dno3_dno3 = dno3_dno3 * crate_uc * ns;
    dun_dno3 = w->upstream_nc * duc_dno3;

    /* column uc */
    somdec_common_jacobian(c, sd, irxn, w, ires_uc, -1, ires_ox, dco2_duc, duc_duc, dun_duc, 1, Jacobian);
    if (sd->nh4_id >= 0) {
      if (up_aq)
        JAC(ires_nh4, ires_uc) = JAC(ires_nh4, ires_uc) - dnh4_duc * DTOT(sd->nh4_id, ispec_uc);
      else
        JAC(ires_nh4, ires_uc) = JAC(ires_nh4, ires_uc) - dnh4_duc;
      if (unimm >= 0) JAC(off + unimm, ires_uc) = JAC(off + unimm, ires_uc) + dnh4_duc;
      if (sd->nimm_id >= 0) JAC(off + sd->nimm_id, ires_uc) = JAC(off + sd->nimm_id, ires_uc) + dnh4_duc;
    }
    if (sd->no3_id >= 0) {
      if (up_aq)
        JAC(ires_no3, ires_uc) = JAC(ires_no3, ires_uc) - dno3_duc * DTOT(sd->no3_id, ispec_uc);
      else
        JAC(ires_no3, ires_uc) = JAC(ires_no3, ires_uc) - dno3_duc;
      if (unimm >= 0) JAC(off + unimm, ires_uc) = JAC(off + unimm, ires_uc) + dno3_duc;
      if (sd->nimm_id >= 0) JAC(off + sd->nimm_id, ires_uc) = JAC(off + sd->nimm_id, ires_uc) + dno3_duc;
    }
    somdec_n_rows_jacobian(c, sd, irxn, w, ires_uc, -1, duc_duc, dun_duc, 1, Jacobian);

    /* column nh4 */
    if (sd->nh4_id >= 0) {
      somdec_common_jacobian(c, sd, irxn, w, ires_nh4, sd->nh4_id, ires_ox, dco2_dnh4, duc_dnh4, dun_dnh4, 0,
                             Jacobian);
      JAC(ires_nh4, ires_nh4) = JAC(ires_nh4, ires_nh4) - dnh4_dnh4 * DTOT(sd->nh4_id, sd->nh4_id);
      if (unimm >= 0) JAC(off + unimm, ires_nh4) = JAC(off + unimm, ires_nh4) + dnh4_dnh4;
      if (sd->nimm_id >= 0) JAC(off + sd->nimm_id, ires_nh4) = JAC(off + sd->nimm_id, ires_nh4) + dnh4_dnh4;
      if (sd->no3_id >= 0) {
        JAC(ires_no3, ires_nh4) = JAC(ires_no3, ires_nh4) - dno3_dnh4 * DTOT(sd->no3_id, sd->nh4_id);
        if (unimm >= 0) JAC(off + unimm, ires_nh4) = JAC(off + unimm, ires_nh4) + dno3_dnh4;
        if (sd->nimm_id >= 0) JAC(off + sd->nimm_id, ires_nh4) = JAC(off + sd->nimm_id, ires_nh4) + dno3_dnh4;
      }
      somdec_n_rows_jacobian(c, sd, irxn, w, ires_nh4, sd->nh4_id, duc_dnh4, dun_dnh4, 0, Jacobian);
    }
    /* column no3 */
    if (sd->no3_id >= 0) {
      somdec_common_jacobian(c, sd, irxn, w, ires_no3, sd->no3_id, ires_ox, dco2_dno3, duc_dno3, dun_dno3, 0,
                             Jacobian);
      if (sd->nh4_id >= 0) {
        JAC(ires_nh4, ires_no3) = JAC(ires_nh4, ires_no3) - dnh4_dno3 * DTOT(sd->nh4_id, sd->no3_id);
        if (unimm >= 0) JAC(off + unimm, ires_no3) = JAC(off + unimm, ires_no3) + dnh4_dno3;
        if (sd->nimm_id >= 0) JAC(off + sd->nimm_id, ires_no3) = JAC(off + sd->nimm_id, ires_no3) + dnh4_dno3;
      }
      JAC(ires_no3, ires_no3) = JAC(ires_no3, ires_no3) - dno3_dno3 * DTOT(sd->no3_id, sd->no3_id);
      if (unimm >= 0) JAC(off + unimm, ires_no3) = JAC(off + unimm, ires_no3) + dno3_dno3;
      if (sd->nimm_id >= 0) JAC(off + sd->nimm_id, ires_no3) = JAC(off + sd->nimm_id, ires_no3) + dno3_dno3;
      somdec_n_rows_jacobian(c, sd, irxn, w, ires_no3, sd->no3_id, duc_dno3, dun_dno3, 0, Jacobian);
    }
  }
}

/* reaction_sandbox_somdec.F90:3477-3640  SomDecNemission */
static void somdec_nemission(const cell_t *c, const pfrx_somdec *sd, double tran_dt, double net_nmin_rate,
                             double *Residual, double *Jacobian, int compute_derivative) {
  const double rpi = 3.14159265358979323846;
  int off = c->naq, n = c->n;
  double porosity = c->porosity, volume = c->volume, saturation = c->sat;
  double theta = saturation * porosity, tc = c->temp;
  int ires_nh4 = sd->nh4_id, ires_n2o = sd->n2o_id;
  double c_nh4 = c->total[ires_nh4] * theta * 1000.0;
  double f_t, f_w, ph, f_ph, temp_real, feps0, dfeps0_dx, dtmin, nratecap, fnratecap, dfnratecap_dnh4, rate_n2o;
  if (sd->n2o_id >= 0 && net_nmin_rate > sd->x0eps) {
    f_t = -0.06 + 0.13 * exp(0.07 * tc);
    f_w = pow((1.27 - saturation) / 0.67, 3.1777) * pow((saturation - 0.0012) / 0.5988, 2.84);
    ph = 6.5;
    if (sd->proton_id >= 0) ph = -log10(c->pri_molal[sd->proton_id] * c->pri_act_coef[sd->proton_id]);
    f_ph = 0.56 + atan(rpi * 0.45 * (-5.0 + ph)) / rpi;
    if (f_t > sd->x0eps && f_w > sd->x0eps && f_ph > sd->x0eps) {
      f_t = fmin(f_t, 1.0);
      f_w = fmin(f_w, 1.0);
      f_ph = fmin(f_ph, 1.0);
      temp_real = f_t * f_w * f_ph;
      if (sd->x0eps > 0.0) {
        hfunction_smooth(c_nh4, sd->x0eps * 10.0, sd->x0eps, &feps0, &dfeps0_dx);
      } else {
        feps0 = 1.0;
        dfeps0_dx = 0.0;
      }
      dtmin = tran_dt;
      nratecap = temp_real * sd->n2o_frac_mineralization * net_nmin_rate * dtmin;
      if (nratecap > c_nh4 * volume) {
        fnratecap = func_monod(c_nh4 * volume, nratecap - c_nh4 * volume, 0);
        dfnratecap_dnh4 = func_monod(c_nh4 * volume, nratecap - c_nh4 * volume, 1);
      } else {
        fnratecap = 1.0;
        dfnratecap_dnh4 = 0.0;
      }
      dfeps0_dx = dfeps0_dx * fnratecap + feps0 * dfnratecap_dnh4;
      feps0 = feps0 * fnratecap;
      rate_n2o = temp_real * sd->n2o_frac_mineralization * net_nmin_rate * feps0;
      Residual[ires_nh4] = Residual[ires_nh4] + rate_n2o;
      Residual[ires_n2o] = Residual[ires_n2o] - 0.5 * rate_n2o;
      if (sd->ngasmin_id >= 0) Residual[off + sd->ngasmin_id] = Residual[off + sd->ngasmin_id] - rate_n2o;
      if (compute_derivative) {
        double drate_n2o_dx = temp_real * sd->n2o_frac_mineralization * net_nmin_rate * dfeps0_dx;
        JAC(ires_nh4, ires_nh4) = JAC(ires_nh4, ires_nh4) + drate_n2o_dx * DTOT(sd->nh4_id, sd->nh4_id);
        JAC(ires_n2o, ires_nh4) = JAC(ires_n2o, ires_nh4) - 0.5 * drate_n2o_dx * DTOT(sd->n2o_id, sd->nh4_id);
        if (sd->ngasmin_id >= 0) JAC(off + sd->ngasmin_id, ires_nh4) = JAC(off + sd->ngasmin_id, ires_nh4) - drate_n2o_dx;
      }
    }
  }
}

/* reaction_sandbox_somdec.F90:1504-1910  SomDecReact */
static void somdec_react(cell_t *c, const pfrx_config *cfg, double tran_dt, double *Residual, double *Jacobian,
                         int compute_derivative) {
  const pfrx_somdec *sd = cfg->somdec;
  double porosity = c->porosity, volume = c->volume, saturation = c->sat;
  double theta = saturation * porosity, tc = c->temp;
  double net_nmin_rate = 0.0, nmin = 0.0, nimm = 0.0;
  int irxn, cur = 0, j;
  for (irxn = 0; irxn < sd->nrxn; irxn++) {
    /* `cur` is the reference's cur_rxn, which is NOT advanced by the `cycle`
     * statements below (reaction_sandbox_somdec.F90:1744,1783 vs :1869) */
    double f_w, f_t, f_depth, kd_scalar, k_decomp = 0.0, scaled_crate_const, c_uc, feps0, dfeps0_dx;
    double crate_uc, dcrate_uc_duc;
    int ispec_uc, nd = sd->downstream_ptr[irxn + 1] - sd->downstream_ptr[irxn];
    somdec_scratch_t w;
    if (cfg->elm_pflotran && cfg->elm_flow_coupled &&
        sd->moisture_response_function[cur] != PFRX_MOISTURE_RESPONSE_OFF) {
      /* a flow mode is active (option%nflowspec > 0), :1640-1643 */
      f_w = elm_moisture_response(c, theta, sd->moisture_response_function[cur]);
    } else if (cfg->elm_pflotran) {
      /* BGC-only coupling (option%nflowspec == 0): factors from ELM, :1645-1650 */
      f_w = c->elm_w;
    } else {
      if (sd->moisture_response_function[cur] == PFRX_MOISTURE_RESPONSE_LOGTHETA) {
        /* the reference writes these literals without a kind suffix, so they
         * are single precision (reaction_sandbox_somdec.F90:1645-1649) */
        if (theta <= (double)0.08f)
          f_w = (double)0.01f;
        else
          f_w = log(theta / (double)0.08f) / (double)logf(1.0f / 0.08f);
      } else {
        f_w = 1.0;
      }
    }
    if (sd->ox_response_function[cur] == PFRX_OX_RESPONSE_WFPS) {
      f_w = f_w * get_aerobic_condition(saturation, sd->ox_half_saturation[cur], sd->ox_response_function[cur], 0);
    } else {
      if (cfg->elm_pflotran) f_w = f_w * c->elm_o;
    }
    switch (sd->temperature_response_function[cur]) {
      case PFRX_TEMPERATURE_RESPONSE_ARRHENIUS:
        f_t = get_temperature_response(tc, sd->temperature_response_function[cur], sd->ea[cur]);
        break;
      case PFRX_TEMPERATURE_RESPONSE_CLMCN:
        f_t = get_temperature_response(tc, sd->temperature_response_function[cur], 0.0);
        break;
      case PFRX_TEMPERATURE_RESPONSE_Q10:
      case PFRX_TEMPERATURE_RESPONSE_DLEM:
        f_t = get_temperature_response(tc, sd->temperature_response_function[cur], sd->q10[cur]);
        break;
      default:
        f_t = cfg->elm_pflotran ? c->elm_t : (double)1.0;
    }
    if (cfg->elm_pflotran) {
      if (sd->decomp_depth_efolding[cur] > 0.0) {
        f_depth = exp(-c->elm_zsoil / sd->decomp_depth_efolding[cur]);
        f_depth = fmin(1.0, fmax(1.e-20, f_depth));
      } else {
        f_depth = 1.0;
      }
      kd_scalar = c->elm_kscalar;
    } else {
      f_depth = 1.0;
      kd_scalar = 1.0;
    }
    if (f_t < 1.0e-20 || f_w < 1.0e-20 || f_depth < 1.0e-20) continue;

    if (sd->rate_constant[irxn] >= 0.0) {
      k_decomp = sd->rate_constant[irxn];
    } else if (sd->rate_decomposition[irxn] >= 0.0) {
      k_decomp = 1.0 - exp(-sd->rate_decomposition[irxn] * tran_dt);
      k_decomp = k_decomp / tran_dt;
    }
    k_decomp = sd->rate_ad_factor[irxn] * k_decomp;
    if (kd_scalar > 0.0 && sd->rate_ad_factor[irxn] > 1.0) k_decomp = k_decomp / kd_scalar;
    k_decomp = fmin(k_decomp, 1.0 / tran_dt);
    scaled_crate_const = k_decomp * volume * f_t * f_w * f_depth;

    ispec_uc = sd->upstream_c_id[irxn];
    if (sd->upstream_is_aqueous[irxn]) {
      c_uc = c->total[ispec_uc];
      c_uc = theta * 1000.0 * c_uc;
    } else {
      c_uc = c->immobile[ispec_uc];
    }
    if (sd->x0eps > 0.0) {
      hfunction_smooth(c_uc, sd->x0eps * 10.0, sd->x0eps, &feps0, &dfeps0_dx);
    } else {
      feps0 = 1.0;
      dfeps0_dx = 0.0;
      if (c_uc <= sd->x0eps) continue;
    }
    crate_uc = scaled_crate_const * c_uc * feps0;
    dcrate_uc_duc = scaled_crate_const * (feps0 + c_uc * dfeps0_dx);

    /* scratch: the persisted ratios, then the on-the-fly N:C ratios */
    w.upstream_nc = c->somdec_nc[irxn];
    w.mineral_c_stoich = sd->mineral_c_stoich[irxn];
    w.mineral_n_stoich = sd->mineral_n_stoich[irxn];
    for (j = 0; j < nd; j++) {
      int jj = sd->downstream_ptr[irxn] + j;
      w.downstream_nc[j] = c->somdec_nc[sd->nrxn + jj];
      if (sd->downstream_n_id[jj] >= 0 && sd->downstream_c_id[jj] >= 0) {
        double c_dc, c_dn;
        if (sd->downstream_is_aqueous[jj]) {
          c_dc = c->total[sd->downstream_c_id[jj]];
          c_dc = theta * 1000.0 * c_dc;
          c_dn = c->total[sd->downstream_n_id[jj]];
          c_dn = theta * 1000.0 * c_dn;
        } else {
          c_dc = c->immobile[sd->downstream_c_id[jj]];
          c_dn = c->immobile[sd->downstream_n_id[jj]];
        }
        if (c_dn >= sd->x0eps && c_dc >= sd->x0eps) w.downstream_nc[j] = c_dn / c_dc;
        c->somdec_nc[sd->nrxn + jj] = w.downstream_nc[j];
      }
    }
    if (sd->upstream_n_id[irxn] >= 0) {
      double c_un, stoich_c, stoich_n;
      if (sd->upstream_is_aqueous[irxn]) {
        c_un = c->total[sd->upstream_n_id[irxn]];
        c_un = theta * 1000.0 * c_un;
      } else {
        c_un = c->immobile[sd->upstream_n_id[irxn]];
      }
      if (c_un >= sd->x0eps && c_uc >= sd->x0eps) w.upstream_nc = c_un / c_uc;
      c->somdec_nc[irxn] = w.upstream_nc;
      stoich_c = 1.0;
      for (j = 0; j < nd; j++) stoich_c = stoich_c - sd->downstream_stoich[sd->downstream_ptr[irxn] + j];
      w.mineral_c_stoich = stoich_c;
      stoich_n = w.upstream_nc;
      for (j = 0; j < nd; j++)
        stoich_n = stoich_n - sd->downstream_stoich[sd->downstream_ptr[irxn] + j] * w.downstream_nc[j];
      w.mineral_n_stoich = stoich_n;
    }

    if (w.mineral_n_stoich >= 0.0) {
      somdec_react1(c, sd, irxn, cur, &w, crate_uc, dcrate_uc_duc, &nmin, Residual, Jacobian, compute_derivative);
      net_nmin_rate = net_nmin_rate + nmin;
    } else {
      somdec_react2(c, sd, irxn, cur, &w, tran_dt, crate_uc, dcrate_uc_duc, &nimm, Residual, Jacobian,
                    compute_derivative);
      net_nmin_rate = net_nmin_rate + nimm;
    }
    cur++;
  }
  if (net_nmin_rate > sd->x0eps)
    somdec_nemission(c, sd, tran_dt, net_nmin_rate, Residual, Jacobian, compute_derivative);
}

/* reaction_sandbox_nitrif.F90:234-502  NitrifReact */
static void nitrif_react(cell_t *c, const pfrx_config *cfg, double *Residual, double *Jacobian,
                         int compute_derivative) {
  const pfrx_nitrif *nt = cfg->nitrif;
  const double rpi = 3.14159265358979323846;
  const double N_molecular_weight = 14.0067;
  int off = c->naq, n = c->n;
  double porosity = c->porosity, volume = c->volume, saturation = c->sat;
  double theta = saturation * porosity, L_water = theta * 1.0e3, tc = c->temp;
  int ires_nh4 = nt->nh4_id, ires_no3 = nt->no3_id, ires_n2o = nt->n2o_id;
  double c_nh4, feps0, dfeps0_dx, rate_nitri, drate_nitri_dnh4, f_t, f_w, f_ph, temp_real, rho_b, M_2_ug_per_g;
  double c_nh4_ugg, rate_n2o, drate_n2o_dnh4, ph;
  c_nh4 = c->total[nt->nh4_id] * L_water;
  if (nt->x0eps > 0.0) {
    hfunction_smooth(c_nh4, nt->x0eps * 10.0, nt->x0eps, &feps0, &dfeps0_dx);
  } else {
    feps0 = 1.0;
    dfeps0_dx = 0.0;
    if (c_nh4 < nt->x0eps) return;
  }
  if (nt->nh4_id >= 0 && nt->no3_id >= 0) {
    f_t = exp(0.08 * (tc - 25.0));
    saturation = fmax(0.0, fmin(saturation, 1.0));
    f_w = saturation * (1.0 - saturation) / 0.25;
    temp_real = fmin(nt->k_nitr_max * f_t * f_w * volume, 1.0);
    rate_nitri = temp_real * (c_nh4 * feps0) * (c_nh4 / (c_nh4 + 4.0));
    Residual[ires_nh4] = Residual[ires_nh4] + rate_nitri;
    Residual[ires_no3] = Residual[ires_no3] - rate_nitri;
    if (compute_derivative) {
      temp_real = c_nh4 * c_nh4 / (c_nh4 + 4.0) * dfeps0_dx +
                  c_nh4 * (c_nh4 + 8.0) / (c_nh4 + 4.0) / (c_nh4 + 4.0) * feps0;
      drate_nitri_dnh4 = nt->k_nitr_max * f_t * f_w * volume * temp_real;
      JAC(ires_nh4, ires_nh4) = JAC(ires_nh4, ires_nh4) + drate_nitri_dnh4 * DTOT(nt->nh4_id, nt->nh4_id);
      JAC(ires_no3, ires_nh4) = JAC(ires_no3, ires_nh4) - drate_nitri_dnh4 * DTOT(nt->no3_id, nt->nh4_id);
    }
  }
  rho_b = cfg->elm_pflotran ? c->elm_bd_dry : (double)1.25e3;
  temp_real = N_molecular_weight * 1.0e6;
  M_2_ug_per_g = temp_real / (volume * rho_b * 1.e3);
  c_nh4_ugg = c_nh4 * volume * M_2_ug_per_g;
  if (nt->n2o_id >= 0 && c_nh4_ugg > 3.0) {
    f_t = -0.06 + 0.13 * exp(0.07 * tc);
    f_w = pow((1.27 - saturation) / 0.67, 3.1777) * pow((saturation - 0.0012) / 0.5988, 2.84);
    ph = 6.5;
    if (nt->proton_id >= 0) ph = -log10(c->pri_molal[nt->proton_id] * c->pri_act_coef[nt->proton_id]);
    f_ph = 0.56 + atan(rpi * 0.45 * (-5.0 + ph)) / rpi;
    if (f_t > 0.0 && f_w > 0.0 && f_ph > 0.0) {
      f_t = fmin(f_t, 1.0);
      f_w = fmin(f_w, 1.0);
      f_ph = fmin(f_ph, 1.0);
      temp_real = (1.0 - exp(-0.0105 * c_nh4_ugg)) * f_t * f_w * f_ph * nt->k_nitr_n2o;
      rate_n2o = temp_real * (c_nh4 * feps0) * volume;
      Residual[ires_nh4] = Residual[ires_nh4] + rate_n2o;
      Residual[ires_n2o] = Residual[ires_n2o] - 0.5 * rate_n2o;
      if (nt->ngasnit_id >= 0) Residual[off + nt->ngasnit_id] = Residual[off + nt->ngasnit_id] - rate_n2o;
      if (compute_derivative) {
        temp_real = (c_nh4 * dfeps0_dx + feps0) * (1.0 - exp(-0.0105 * c_nh4_ugg));
        temp_real = temp_real + (c_nh4 * feps0) * 0.0105 * M_2_ug_per_g * exp(-0.0105 * c_nh4_ugg);
        drate_n2o_dnh4 = temp_real * nt->k_nitr_n2o * f_t * f_w * f_ph * volume;
        JAC(ires_nh4, ires_nh4) = JAC(ires_nh4, ires_nh4) + drate_n2o_dnh4 * DTOT(nt->nh4_id, nt->nh4_id);
        JAC(ires_n2o, ires_nh4) = JAC(ires_n2o, ires_nh4) - 0.5 * drate_n2o_dnh4 * DTOT(nt->n2o_id, nt->nh4_id);
        if (nt->ngasnit_id >= 0)
          JAC(off + nt->ngasnit_id, ires_nh4) = JAC(off + nt->ngasnit_id, ires_nh4) - drate_n2o_dnh4;
      }
    }
  }
}

/* reaction_sandbox_denitr.F90:212-404  DenitrReact */
static void denitr_react(cell_t *c, const pfrx_config *cfg, double *Residual, double *Jacobian,
                         int compute_derivative) {
  const pfrx_denitr *dn = cfg->denitr;
  int off = c->naq, n = c->n;
  double porosity = c->porosity, volume = c->volume, saturation = c->sat, tc = c->temp;
  double L_water = porosity * saturation * 1.e3;
  int ires_no3 = dn->no3_id, ires_n2 = dn->n2_id;
  double temp_real, f_t, s_min, f_w, c_no3, feps0, dfeps0_dx, fno3, dfno3_dno3, rate_deni, drate_deni_dno3;
  if (dn->n2_id < 0) return;
  temp_real = cfg->elm_pflotran ? c->elm_bsw : (double)1.0;
  f_t = exp(0.08 * (tc - 25.0));
  s_min = 0.6;
  f_w = 0.0;
  if (saturation > s_min) {
    f_w = (saturation - s_min) / (1.0 - s_min);
    f_w = pow(f_w, temp_real);
  }
  c_no3 = c->total[ires_no3] * L_water;
  if (dn->x0eps > 0.0) {
    hfunction_smooth(c_no3, dn->x0eps * 10.0, dn->x0eps, &feps0, &dfeps0_dx);
  } else {
    feps0 = 1.0;
    dfeps0_dx = 0.0;
    if (c_no3 <= dn->x0eps) return;
  }
  if (dn->half_saturation > 0.0) {
    fno3 = func_monod(c_no3, dn->half_saturation, 0);
    dfno3_dno3 = func_monod(c_no3, dn->half_saturation, 1);
  } else {
    fno3 = 1.0;
    dfno3_dno3 = 0.0;
  }
  if (f_t > 0.0 && f_w > 0.0) {
    rate_deni = dn->k_deni_max * f_t * f_w * fno3 * (c_no3 * volume * feps0);
    Residual[ires_no3] = Residual[ires_no3] + rate_deni;
    Residual[ires_n2] = Residual[ires_n2] - 0.5 * rate_deni;
    if (dn->ngasdeni_id >= 0) Residual[off + dn->ngasdeni_id] = Residual[off + dn->ngasdeni_id] - rate_deni;
    if (compute_derivative) {
      temp_real = dfno3_dno3 * (c_no3 * volume * feps0) + fno3 * (c_no3 * volume * dfeps0_dx + feps0);
      drate_deni_dno3 = dn->k_deni_max * f_t * f_w * temp_real;
      JAC(ires_no3, ires_no3) = JAC(ires_no3, ires_no3) + drate_deni_dno3 * DTOT(dn->no3_id, dn->no3_id);
      JAC(ires_n2, ires_no3) = JAC(ires_n2, ires_no3) - 0.5 * drate_deni_dno3 * DTOT(dn->n2_id, dn->no3_id);
      if (dn->ngasdeni_id >= 0)
        JAC(off + dn->ngasdeni_id, ires_no3) = JAC(off + dn->ngasdeni_id, ires_no3) - drate_deni_dno3;
    }
  }
}
#undef SD_RES
#undef JAC
#undef DTOT

/* reaction_sandbox_plantn.F90:222-640  PlantNReact */
static void plantn_react(cell_t *c, const pfrx_config *cfg, double tran_dt, double *Residual, double *Jacobian,
                         int compute_derivative) {
  const pfrx_plantn *pn = cfg->plantn;
  int off = c->naq, n = c->n;
  double volume = c->volume, porosity = c->porosity, saturation = c->sat, tc = c->temp;
  double theta, L_water, c_nh4 = 0.0, c_no3 = 0.0;
  double fnh4 = 1.0, dfnh4_dnh4 = 0.0, fno3 = 1.0, dfno3_dno3 = 0.0;
  double fnh4_inhibit_no3 = 1.0, dfnh4_inhibit_no3_dnh4 = 0.0, dfnh4_inhibit_no3_dno3 = 0.0;
  double temp_real, feps0, dfeps0_dx, rate_plantndemand, dtmin, nratecap, fnratecap, dfnratecap_dnh4,
      dfnratecap_dno3;
  int ires_plantn = off + pn->plantn_id, ires_nh4 = pn->nh4_id, ires_no3 = pn->no3_id;
  if (saturation < 0.01) return;
  theta = saturation * porosity;
  L_water = theta * 1.0e3;
  if (tc < -0.1) return;
  if (pn->nh4_id >= 0 && pn->no3_id >= 0) {
    c_nh4 = c->total[pn->nh4_id] * L_water;
    c_no3 = c->total[pn->no3_id] * L_water;
    if ((c_nh4 > pn->x0eps_nh4 && c_no3 > pn->x0eps_no3) && pn->inhibition_nh4_no3 > 0.0) {
      temp_real = c_nh4 / c_no3;
      fnh4_inhibit_no3 = func_monod(temp_real, 1.0 / pn->inhibition_nh4_no3, 0);
    } else {
      if (c_nh4 > pn->x0eps_nh4 && c_no3 <= pn->x0eps_no3)
        fnh4_inhibit_no3 = 1.0;
      else if (c_nh4 <= pn->x0eps_nh4 && c_no3 > pn->x0eps_no3)
        fnh4_inhibit_no3 = 0.0;
      else
        return;
    }
  }
  if (pn->nh4_id >= 0) {
    c_nh4 = c->total[pn->nh4_id] * L_water;
    fnh4 = func_monod(c_nh4, pn->half_saturation_nh4, 0);
    dfnh4_dnh4 = func_monod(c_nh4, pn->half_saturation_nh4, 1);
    if (pn->x0eps_nh4 > 0.0) {
      hfunction_smooth(c_nh4, pn->x0eps_nh4 * 10.0, pn->x0eps_nh4, &feps0, &dfeps0_dx);
    } else {
      feps0 = 1.0;
      dfeps0_dx = 0.0;
    }
    dfnh4_dnh4 = dfnh4_dnh4 * feps0 + fnh4 * dfeps0_dx;
    fnh4 = fnh4 * feps0;
  }
  if (pn->no3_id >= 0) {
    c_no3 = c->total[pn->no3_id] * L_water;
    fno3 = func_monod(c_no3, pn->half_saturation_no3, 0);
    dfno3_dno3 = func_monod(c_no3, pn->half_saturation_no3, 1);
    if (pn->x0eps_no3 > 0.0) {
      hfunction_smooth(c_no3, pn->x0eps_no3 * 10.0, pn->x0eps_no3, &feps0, &dfeps0_dx);
    } else {
      feps0 = 1.0;
      dfeps0_dx = 0.0;
    }
    dfno3_dno3 = dfno3_dno3 * feps0 + fno3 * dfeps0_dx;
    fno3 = fno3 * feps0;
  }
  if (cfg->elm_pflotran) {
    rate_plantndemand = fmax(0.0, c->elm_plantndemand * volume);
    if (rate_plantndemand <= 0.0) return;
  } else {
    rate_plantndemand = 1.e-2 * volume;
  }
  if (pn->plantndemand_id >= 0)
    Residual[off + pn->plantndemand_id] = Residual[off + pn->plantndemand_id] - rate_plantndemand;
  if (rate_plantndemand > 0.0) {
    dtmin = tran_dt;
    if (pn->nh4_id >= 0) {
      nratecap = rate_plantndemand * dtmin;
      if (pn->no3_id >= 0) nratecap = rate_plantndemand * fnh4_inhibit_no3 * dtmin;
      if (nratecap > c_nh4 * volume) {
        fnratecap = func_monod(c_nh4 * volume, nratecap - c_nh4 * volume, 0);
        dfnratecap_dnh4 = func_monod(c_nh4 * volume, nratecap - c_nh4 * volume, 1);
      } else {
        fnratecap = 1.0;
        dfnratecap_dnh4 = 0.0;
      }
      dfnh4_dnh4 = dfnh4_dnh4 * fnratecap + fnh4 * dfnratecap_dnh4;
      fnh4 = fnh4 * fnratecap;
    }
    if (pn->no3_id >= 0) {
      nratecap = rate_plantndemand * dtmin;
      if (pn->nh4_id >= 0) nratecap = rate_plantndemand * (1.0 - fnh4_inhibit_no3) * dtmin;
      if (nratecap > c_no3 * volume) {
        fnratecap = func_monod(c_no3 * volume, nratecap - c_no3 * volume, 0);
        dfnratecap_dno3 = func_monod(c_no3 * volume, nratecap - c_no3 * volume, 1);
      } else {
        fnratecap = 1.0;
        dfnratecap_dno3 = 0.0;
      }
      dfno3_dno3 = dfno3_dno3 * fnratecap + fno3 * dfnratecap_dno3;
      fno3 = fno3 * fnratecap;
    }
  }
#define JAC(i, j) Jacobian[(i) + (size_t)(j) * n]
#define DTOT(i, j) c->dtotal[(i) + (size_t)(j) * c->naq]
  if (pn->nh4_id >= 0) {
    double nrate_nh4 = rate_plantndemand * fnh4;
    if (pn->no3_id >= 0) nrate_nh4 = rate_plantndemand * fnh4 * fnh4_inhibit_no3;
    Residual[ires_nh4] = Residual[ires_nh4] + nrate_nh4;
    Residual[ires_plantn] = Residual[ires_plantn] - nrate_nh4;
    if (pn->plantnh4uptake_id >= 0)
      Residual[off + pn->plantnh4uptake_id] = Residual[off + pn->plantnh4uptake_id] - nrate_nh4;
    if (compute_derivative) {
      double dnrate_nh4_dnh4 = rate_plantndemand * dfnh4_dnh4;
      if (pn->no3_id >= 0) {
        temp_real = fnh4 * dfnh4_inhibit_no3_dnh4 + fnh4_inhibit_no3 * dfnh4_dnh4;
        dnrate_nh4_dnh4 = rate_plantndemand * temp_real;
      }
      JAC(ires_nh4, ires_nh4) = JAC(ires_nh4, ires_nh4) + dnrate_nh4_dnh4 * DTOT(pn->nh4_id, pn->nh4_id);
      JAC(ires_plantn, ires_nh4) = JAC(ires_plantn, ires_nh4) - dnrate_nh4_dnh4;
      if (pn->plantnh4uptake_id >= 0)
        JAC(off + pn->plantnh4uptake_id, ires_nh4) = JAC(off + pn->plantnh4uptake_id, ires_nh4) - dnrate_nh4_dnh4;
    }
  }
  if (pn->no3_id >= 0) {
    double nrate_no3 = rate_plantndemand * fno3;
    if (pn->nh4_id >= 0) nrate_no3 = rate_plantndemand * fno3 * (1.0 - fnh4_inhibit_no3);
    Residual[ires_no3] = Residual[ires_no3] + nrate_no3;
    Residual[ires_plantn] = Residual[ires_plantn] - nrate_no3;
    if (pn->plantno3uptake_id >= 0)
      Residual[off + pn->plantno3uptake_id] = Residual[off + pn->plantno3uptake_id] - nrate_no3;
    if (compute_derivative) {
      double dnrate_no3_dno3 = rate_plantndemand * dfno3_dno3;
      if (pn->nh4_id >= 0) {
        temp_real = dfno3_dno3 * (1.0 - fnh4_inhibit_no3) + fno3 * (-1.0 * dfnh4_inhibit_no3_dno3);
        dnrate_no3_dno3 = rate_plantndemand * temp_real;
      }
      JAC(ires_no3, ires_no3) = JAC(ires_no3, ires_no3) + dnrate_no3_dno3 * DTOT(pn->no3_id, pn->no3_id);
      JAC(ires_plantn, ires_no3) = JAC(ires_plantn, ires_no3) - dnrate_no3_dno3;
      if (pn->plantno3uptake_id >= 0)
        JAC(off + pn->plantno3uptake_id, ires_no3) = JAC(off + pn->plantno3uptake_id, ires_no3) - dnrate_no3_dno3;
    }
  }
}

/* reaction_sandbox_langmu.F90:183-330  LangmuirReact */
static void langmuir_react(cell_t *c, const pfrx_config *cfg, double tran_dt, double *Residual, double *Jacobian,
                           int compute_derivative) {
  const pfrx_langmuir *lg = cfg->langmuir;
  int off = c->naq, n = c->n;
  double porosity = c->porosity, volume = c->volume;
  double Lwater = volume * 1000.0 * porosity * c->sat;
  int ires_aq = lg->aq_id, ires_sorb = off + lg->sorb_id;
  double c_aq = c->total[lg->aq_id], c_sorb = c->immobile[lg->sorb_id];
  double rate, drate_daq, drate_dsorb, dtmin, c_aq_eq, temp_real, ratecap, fratecap, dfratecap_dx;
  if (lg->s_max < c_sorb) {
    dtmin = tran_dt;
    rate = (lg->s_max - c_sorb) * volume / dtmin;
    drate_dsorb = -1.0 / dtmin;
    drate_daq = 0.0;
  } else {
    c_aq_eq = 0.999 * c_sorb / (lg->s_max - 0.999 * c_sorb) / lg->k_equilibrium;
    rate = lg->k_kinetic * (c_aq - c_aq_eq) * Lwater;
    temp_real = -lg->k_kinetic / lg->k_equilibrium * Lwater / volume;
    drate_dsorb = temp_real * lg->s_max / (lg->s_max - 0.999 * c_sorb) / (lg->s_max - 0.999 * c_sorb);
    drate_daq = lg->k_kinetic;
    fratecap = 1.0;
    dfratecap_dx = 0.0;
    if (rate > 0.0) {
      dtmin = tran_dt;
      ratecap = 0.999 * (lg->s_max - c_sorb) * volume / dtmin;
      if (ratecap < rate) {
        fratecap = ratecap / rate;
        if (compute_derivative) {
          temp_real = -0.999 / dtmin;
          dfratecap_dx = (ratecap * drate_dsorb - rate * temp_real) / rate / rate;
          drate_dsorb = fratecap * drate_dsorb + rate * dfratecap_dx;
        }
        rate = rate * fratecap;
      }
      ratecap = 0.999 * (c_aq - c_aq_eq) * Lwater / dtmin;
      if (ratecap < rate) {
        fratecap = ratecap / rate;
        if (compute_derivative) {
          temp_real = -0.999 / lg->k_equilibrium * Lwater / volume / dtmin;
          temp_real = temp_real * lg->s_max / (lg->s_max - 0.999 * c_sorb) / (lg->s_max - 0.999 * c_sorb);
          dfratecap_dx = (ratecap * drate_dsorb - rate * temp_real) / rate / rate;
          drate_dsorb = fratecap * drate_dsorb + rate * dfratecap_dx;
          temp_real = 0.999 / dtmin;
          dfratecap_dx = (ratecap * drate_daq - rate * temp_real) / rate / rate;
          drate_daq = fratecap * drate_daq + rate * dfratecap_dx;
        }
        rate = rate * fratecap;
      }
    }
  }
  Residual[ires_aq] = Residual[ires_aq] + rate;
  Residual[ires_sorb] = Residual[ires_sorb] - rate;
  if (compute_derivative) {
    JAC(ires_aq, ires_aq) = JAC(ires_aq, ires_aq) + drate_daq * DTOT(lg->aq_id, lg->aq_id);
    JAC(ires_sorb, ires_aq) = JAC(ires_sorb, ires_aq) - drate_daq;
    JAC(ires_aq, ires_sorb) = JAC(ires_aq, ires_sorb) + drate_dsorb;
    JAC(ires_sorb, ires_sorb) = JAC(ires_sorb, ires_sorb) - drate_dsorb;
  }
#undef JAC
#undef DTOT
}

/* reaction_sandbox_cndegas.F90:548-787: Weiss (1974) CO2, Weiss & Price (1980) N2O, Weiss (1970) N2
 * solubilities.  rgas is a default-real literal in the reference (0.08205601 without d0): its value
 * is the single-precision one. */
#define CND_RGAS ((double)0.08205601f)
static double weiss_co2_xmole(double tt, double tp, double ts, double pco2) {
  const double atm = 1.01325e5, xmwh2o = 18.01534e-3;
  const double a1 = -58.0931, a2 = 90.5069, a3 = 22.2940, b1 = 0.027766, b2 = -0.025888, b3 = 0.0050578;
  double tk = tt + 273.15, tk2 = tk * tk, tk3 = tk2 * tk, tk_100k = tk / 100.0;
  double p_rt = (tp / atm) / CND_RGAS / tk, x1, x2, epsilon, bt, fg, k0, vbar, cco2;
  p_rt = p_rt / 1000.0;
  x1 = pco2 / tp;
  x2 = 1.0 - x1;
  epsilon = 57.7 - 0.118 * tk;
  bt = -1636.75 + 12.0408 * tk - 3.27957e-2 * tk2 + 3.16528e-5 * tk3;
  fg = pco2 * exp((bt + 2.0 * x2 * x2 * epsilon) * p_rt);
  k0 = a1 + a2 / tk_100k + a3 * log(tk_100k) + ts * (b1 + b2 * tk_100k + b3 * tk_100k * tk_100k);
  k0 = exp(k0);
  k0 = k0 / atm;
  vbar = exp((1.0 - tp / atm) * 30.0e-3 / CND_RGAS / tk);
  cco2 = k0 * fg * vbar;
  return cco2 / (1.0 / xmwh2o);
}
static double weiss_price_n2o_xmole(double tt, double tp, double ts, double pn2o) {
  const double atm = 1.01325e5, xmwh2o = 18.01534e-3;
  const double a1 = -62.7076, a2 = 97.3066, a3 = 24.1406, b1 = -0.058420, b2 = 0.033193, b3 = -0.0051313;
  double tk = tt + 273.15, tk2 = tk * tk, tk_100k = tk / 100.0;
  double p_rt = (tp / atm) / CND_RGAS / tk, x1, x2, epsilon, bt, fg, k0, vbar, cn2o;
  p_rt = p_rt / 1000.0;
  x1 = pn2o / tp;
  x2 = 1.0 - x1;
  epsilon = 65.0 - 0.1338 * tk;
  bt = -905.95 + 4.1685 * tk - 0.0052734 * tk2;
  fg = pn2o * exp((bt + 2.0 * x2 * x2 * epsilon) * p_rt);
  k0 = a1 + a2 / tk_100k + a3 * log(tk_100k) + ts * (b1 + b2 * tk_100k + b3 * tk_100k * tk_100k);
  k0 = exp(k0);
  k0 = k0 / atm;
  vbar = exp((1.0 - tp / atm) * 32.3e-3 / CND_RGAS / tk);
  cn2o = k0 * fg * vbar;
  return cn2o / (1.0 / xmwh2o);
}
static double weiss_n2_xmole(double tt, double ts, double pn2) {
  const double atmn2 = 0.78084, atm = 1.01325e5, xmwh2o = 18.01534e-3;
  const double a1 = -172.4965, a2 = 248.4262, a3 = 143.3483, a4 = -21.7120, b1 = -0.049781, b2 = -0.025018,
               b3 = -0.0034861;
  double tk = tt + 273.15, tk_100k = tk / 100.0, k0, kh, cn2;
  k0 = a1 + a2 / tk_100k + a3 * log(tk_100k) + a4 * tk_100k + ts * (b1 + b2 * tk_100k + b3 * tk_100k * tk_100k);
  k0 = exp(k0);
  k0 = (k0 * 1.e-3) / CND_RGAS / 298.15;
  kh = (atmn2 * atm) / k0;
  kh = kh * (1.0 / xmwh2o); /* the reference divides by this kh although it is per mole fraction by now */
  cn2 = (pn2 * 1.0) / kh;
  return cn2 / (1.0 / xmwh2o);
}

/* reaction_sandbox_cndegas.F90:216-546  CNdegasReact */
static void cndegas_react(cell_t *c, const pfrx_config *cfg, double *Residual, double *Jacobian,
                          int compute_derivative) {
  const pfrx_cndegas *cd = cfg->cndegas;
  const double H2O_kg_mol = 18.01534e-3, rgas = 8.3144621;
  int off = c->naq, n = c->n, elm = cfg->elm_pflotran ? 1 : 0;
  double convert_molal_to_molar = cd->initialize_with_molality ? c->den_kg * 1.0 / 1000.0 : (double)1.0;
  double tc = cd->reference_temperature, air_press = cd->reference_pressure, lsat = 0.50;
  double porosity, volume, air_vol, air_molar, temp_real, total_sal, rate, drate;
#define JAC(i, j) Jacobian[(i) + (size_t)(j) * n]
#define DTOT(i, j) c->dtotal[(i) + (size_t)(j) * c->naq]
  if (cd->cell_state_mode >= 1) {
    air_press = fmax(air_press, c->pres);
    lsat = c->sat;
    if (cd->cell_state_mode >= 2) tc = c->temp;
  }
  porosity = c->porosity;
  volume = c->volume;
  air_vol = 1.0;
  air_molar = air_press / rgas / (tc + 273.15);
  if (cd->co2a_id >= 0 && cd->co2g_id >= 0) {
    int ia = cd->co2a_id, ig = cd->co2g_id + off;
    double c_aq = c->total[cd->co2a_id], c_eq;
    double co2_p = 350.0e-6 * cd->reference_pressure;
    if (elm) {
      double co2_molar = c->immobile[cd->co2g_id] / air_vol;
      co2_p = co2_molar / air_molar * air_press;
    }
    temp_real = fmax(fmin(tc, 40.0), -1.0);
    total_sal = 1.e-20;
    c_eq = weiss_co2_xmole(temp_real, air_press, total_sal, co2_p) / H2O_kg_mol;
    temp_real = volume * 1000.0 * porosity * lsat;
    rate = cd->k_kinetic_co2 * (c_aq - c_eq) * temp_real;
    if (fabs(rate) > 1.0e-20) {
      Residual[ia] = Residual[ia] + rate;
      Residual[ig] = Residual[ig] - rate;
      if (compute_derivative) {
        drate = cd->k_kinetic_co2 * temp_real;
        JAC(ia, ia) = JAC(ia, ia) + drate * DTOT(cd->co2a_id, cd->co2a_id);
        JAC(ig, ia) = JAC(ig, ia) - drate;
      }
    }
  }
  if (cd->n2oa_id >= 0 && cd->n2og_id >= 0) {
    int ia = cd->n2oa_id, ig = cd->n2og_id + off;
    double c_aq = c->total[cd->n2oa_id], c_eq;
    double n2o_p = 310.0e-9 * cd->reference_pressure;
    if (elm) {
      double n2o_molar = c->immobile[cd->n2og_id] / air_vol;
      n2o_p = n2o_molar / air_molar * air_press;
    }
    temp_real = fmax(fmin(tc, 40.0), 1.e-20);
    total_sal = 1.0e-20;
    c_eq = weiss_price_n2o_xmole(temp_real, air_press, total_sal, n2o_p) / H2O_kg_mol;
    temp_real = volume * 1000.0 * porosity * lsat;
    rate = cd->k_kinetic_n2o * (c_aq - c_eq) * temp_real;
    if (fabs(rate) > 1.0e-20) {
      Residual[ia] = Residual[ia] + rate;
      Residual[ig] = Residual[ig] - rate;
      if (compute_derivative) {
        drate = cd->k_kinetic_n2o * temp_real;
        JAC(ia, ia) = JAC(ia, ia) + drate * DTOT(cd->n2oa_id, cd->n2oa_id);
        JAC(ig, ia) = JAC(ig, ia) - drate;
      }
    }
  }
  if (cd->n2a_id >= 0 && cd->n2g_id >= 0) {
    int ia = cd->n2a_id, ig = cd->n2g_id + off;
    double c_aq = c->total[cd->n2a_id], c_eq;
    double n2_p = 0.78084 * cd->reference_pressure;
    if (elm) {
      double n2_molar = c->immobile[cd->n2g_id] / air_vol;
      n2_p = n2_molar / air_molar * air_press;
    }
    temp_real = fmax(fmin(tc, 40.0), -2.0);
    total_sal = 1.0e-20;
    c_eq = weiss_n2_xmole(temp_real, total_sal, n2_p) / H2O_kg_mol;
    temp_real = volume * porosity * lsat * 1.e3;
    rate = cd->k_kinetic_n2 * (c_aq - c_eq) * temp_real;
    if (fabs(rate) > 1.0e-20) {
      Residual[ia] = Residual[ia] + rate;
      Residual[ig] = Residual[ig] - rate;
      if (compute_derivative) {
        drate = cd->k_kinetic_n2 * temp_real;
        JAC(ia, ia) = JAC(ia, ia) + drate * DTOT(cd->n2a_id, cd->n2a_id);
        JAC(ig, ia) = JAC(ig, ia) - drate;
      }
    }
  }
  if (cd->fixph_on) {
    int ip = cd->proton_id, ih = cd->himm_id + off;
    double c_h = c->pri_molal[cd->proton_id] * convert_molal_to_molar;
    double c_h_fix = pow(10.0, -1.0 * cd->fixph) / c->pri_act_coef[cd->proton_id];
    temp_real = volume * 1000.0 * porosity * lsat;
    rate = cd->k_kinetic_h * (c_h - c_h_fix) * temp_real;
    if (fabs(rate) > 1.0e-20) {
      Residual[ip] = Residual[ip] + rate;
      Residual[ih] = Residual[ih] - rate;
      if (compute_derivative) {
        drate = cd->k_kinetic_h * convert_molal_to_molar * temp_real;
        JAC(ip, ip) = JAC(ip, ip) + drate;
        /* as written (:501): the Himm row takes the TRANSPOSED entry as its starting value */
        JAC(ih, ip) = JAC(ip, ih) - drate;
      }
    }
  }
#undef JAC
#undef DTOT
}

/* reaction_sandbox_calcite.F90:177-365  CalciteEvaluate */
static void calcite_evaluate(cell_t *c, const pfrx_config *cfg, double *Residual, double *Jacobian,
                             int compute_derivative) {
  const pfrx_calcite_sandbox *this_ = cfg->calcite;
  int n = c->n, i, j, icomp, jcomp, imnrl = this_->mineral_id;
  int p0 = cfg->kinmnrl_ptr[imnrl], p1 = cfg->kinmnrl_ptr[imnrl + 1];
  double ln_conc[PFRX_MAX_NCOMP * 4], ln_act[PFRX_MAX_NCOMP * 4];
  double rate, drate_dQK, lnQK, QK, affinity_factor, dQK_dmj, sign_, molality_to_molarity;
  int calculate_rate;
  molality_to_molarity = c->den_kg * 1.e-3;
  /* reaction path #1: stoichiometry and logK of the mineral from the database */
  for (i = 0; i < c->naq; i++) {
    ln_conc[i] = log(c->pri_molal[i]);
    ln_act[i] = ln_conc[i] + log(c->pri_act_coef[i]);
  }
  lnQK = -c->kinmnrl_logK[imnrl] * LOG_TO_LN;
  if (cfg->kinmnrl_h2ostoich[imnrl] != 0.0) lnQK = lnQK + cfg->kinmnrl_h2ostoich[imnrl] * c->ln_act_h2o;
  for (i = p0; i < p1; i++) {
    icomp = cfg->kinmnrl_specid[i];
    lnQK = lnQK + cfg->kinmnrl_stoich[i] * ln_act[icomp];
  }
  QK = exp(lnQK);
  affinity_factor = 1.0 - QK;
  sign_ = copysign(1.0, affinity_factor);
  rate = 0.0;
  calculate_rate = c->mnrl_volfrac[imnrl] > 0 || sign_ < 0.0;
  if (calculate_rate) rate = -c->mnrl_area[imnrl] * sign_ * fabs(affinity_factor) * this_->rate_constant1;
  c->sandbox_aux = rate;
  rate = rate * c->volume;
  for (i = p0; i < p1; i++) {
    icomp = cfg->kinmnrl_specid[i];
    Residual[icomp] = Residual[icomp] + cfg->kinmnrl_stoich[i] * rate;
  }
  if (compute_derivative && calculate_rate) {
    drate_dQK = c->mnrl_area[imnrl] * this_->rate_constant1 * c->volume;
    for (j = p0; j < p1; j++) {
      jcomp = cfg->kinmnrl_specid[j];
      dQK_dmj = cfg->kinmnrl_stoich[j] * QK * exp(-ln_conc[jcomp]) * molality_to_molarity;
      for (i = p0; i < p1; i++) {
        icomp = cfg->kinmnrl_specid[i];
        Jacobian[icomp + jcomp * n] = Jacobian[icomp + jcomp * n] + cfg->kinmnrl_stoich[i] * drate_dQK * dQK_dmj;
      }
    }
  }
  /* reaction path #2: QK = {Ca++}{HCO3-}/(Keq {H+}), pKeq 1.8487 */
  {
    int ih = this_->h_ion_id, ica = this_->calcium_id, ib = this_->bicarbonate_id;
    lnQK = -1.8487 * LOG_TO_LN - ln_act[ih] + ln_act[ica] + ln_act[ib];
    affinity_factor = 1.0 - exp(lnQK);
    sign_ = copysign(1.0, affinity_factor);
    rate = 0.0;
    calculate_rate = c->mnrl_volfrac[imnrl] > 0 || sign_ < 0.0;
    if (calculate_rate) rate = -c->mnrl_area[imnrl] * sign_ * fabs(affinity_factor) * this_->rate_constant2;
    c->sandbox_aux = c->sandbox_aux + rate;
    rate = rate * c->volume;
    Residual[ih] = Residual[ih] - rate;
    Residual[ica] = Residual[ica] + rate;
    Residual[ib] = Residual[ib] + rate;
    if (compute_derivative && calculate_rate) {
      drate_dQK = c->mnrl_area[imnrl] * this_->rate_constant2 * c->volume;
      jcomp = ih;
      dQK_dmj = -1.0 * exp(lnQK - ln_conc[jcomp]) * molality_to_molarity;
      Jacobian[ih + jcomp * n] = Jacobian[ih + jcomp * n] - drate_dQK * dQK_dmj;
      Jacobian[ica + jcomp * n] = Jacobian[ica + jcomp * n] + drate_dQK * dQK_dmj;
      Jacobian[ib + jcomp * n] = Jacobian[ib + jcomp * n] + drate_dQK * dQK_dmj;
      jcomp = ica;
      dQK_dmj = 1.0 * exp(lnQK - ln_conc[jcomp]) * molality_to_molarity;
      Jacobian[ih + jcomp * n] = Jacobian[ih + jcomp * n] - drate_dQK * dQK_dmj;
      Jacobian[ica + jcomp * n] = Jacobian[ica + jcomp * n] + drate_dQK * dQK_dmj;
      Jacobian[ib + jcomp * n] = Jacobian[ib + jcomp * n] + drate_dQK * dQK_dmj;
      jcomp = ib;
      dQK_dmj = 1.0 * exp(lnQK - ln_conc[jcomp]) * molality_to_molarity;
      Jacobian[ih + jcomp * n] = Jacobian[ih + jcomp * n] - drate_dQK * dQK_dmj;
      Jacobian[ica + jcomp * n] = Jacobian[ica + jcomp * n] + drate_dQK * dQK_dmj;
      Jacobian[ib + jcomp * n] = Jacobian[ib + jcomp * n] + drate_dQK * dQK_dmj;
    }
  }
}

static int n_sandboxes(const pfrx_config *cfg) {
  return (cfg->clmcn_nrxn > 0) + (cfg->somdec != NULL) + (cfg->nitrif != NULL) + (cfg->denitr != NULL) +
         (cfg->plantn != NULL) + (cfg->langmuir != NULL) + (cfg->cndegas != NULL) + (cfg->calcite != NULL) +
         (cfg->radon != NULL);
}

/* reaction_sandbox.F90:294-330  RSandboxEvaluate: walk the list in deck order */
static void r_sandbox_evaluate(cell_t *c, const pfrx_config *cfg, double tran_dt, double *Res, double *Jac,
                               int derivative) {
  static const int32_t default_order[9] = {PFRX_SANDBOX_CLM_CN, PFRX_SANDBOX_SOMDEC,  PFRX_SANDBOX_NITRIF,
                                           PFRX_SANDBOX_DENITR, PFRX_SANDBOX_PLANTN, PFRX_SANDBOX_LANGMUIR,
                                           PFRX_SANDBOX_CNDEGAS, PFRX_SANDBOX_CALCITE, PFRX_SANDBOX_RADON};
  const int32_t *order = cfg->sandbox_list ? cfg->sandbox_list : default_order;
  int ns = cfg->sandbox_list ? cfg->nsandbox : 9, k;
  for (k = 0; k < ns; k++) {
    switch (order[k]) {
      case PFRX_SANDBOX_CLM_CN:
        if (cfg->clmcn_nrxn > 0) clm_cn_react(c, cfg, Res, Jac, derivative);
        break;
      case PFRX_SANDBOX_SOMDEC:
        if (cfg->somdec) somdec_react(c, cfg, tran_dt, Res, Jac, derivative);
        break;
      case PFRX_SANDBOX_NITRIF:
        if (cfg->nitrif) nitrif_react(c, cfg, Res, Jac, derivative);
        break;
      case PFRX_SANDBOX_DENITR:
        if (cfg->denitr) denitr_react(c, cfg, Res, Jac, derivative);
        break;
      case PFRX_SANDBOX_PLANTN:
        if (cfg->plantn) plantn_react(c, cfg, tran_dt, Res, Jac, derivative);
        break;
      case PFRX_SANDBOX_LANGMUIR:
        if (cfg->langmuir) langmuir_react(c, cfg, tran_dt, Res, Jac, derivative);
        break;
      case PFRX_SANDBOX_CNDEGAS:
        if (cfg->cndegas) cndegas_react(c, cfg, Res, Jac, derivative);
        break;
      case PFRX_SANDBOX_CALCITE:
        if (cfg->calcite) calcite_evaluate(c, cfg, Res, Jac, derivative);
        break;
      case PFRX_SANDBOX_RADON: /* reaction_sandbox_radon.F90:150-188 RadonEvaluate: no derivative */
        if (cfg->radon)
          Res[cfg->radon->species_id] = Res[cfg->radon->species_id] -
                                        (1.0) * cfg->radon->radon_generation_rate *
                                            c->mnrl_volfrac[cfg->radon->mineral_id] * c->volume;
        break;
      default:
        break;
    }
  }
}

/* reaction.F90:5211-5311  RRadioactiveDecay: one parent, aqueous + sorbed inventory, any
 * number of daughters */
static void r_radioactive_decay(cell_t *c, const pfrx_config *cfg, double *Res, double *Jac, int compute_derivative) {
  int naq = c->naq, n = c->n, irxn, i, j;
  double L_pore = c->porosity * c->volume * 1.e3;
  double L_water = L_pore * c->sat;
  double L_gas = L_pore * c->sat_gas;
  int have_sorb = neqsorb(cfg) > 0;
  for (irxn = 0; irxn < cfg->nradiodecay_rxn; irxn++) {
    int p0 = cfg->radiodecay_ptr[irxn], p1 = cfg->radiodecay_ptr[irxn + 1];
    int icomp = cfg->radiodecay_forward_specid[irxn], jcomp;
    double sum = c->total[icomp] * L_water, rate, tempreal;
    if (c->ngas > 0) sum = sum + c->total_gas[icomp] * L_gas;
    if (have_sorb) sum = sum + c->total_sorb_eq[icomp] * c->volume;
    rate = sum * cfg->radiodecay_kf[irxn];
    for (i = p0; i < p1; i++) {
      icomp = cfg->radiodecay_specid[i];
      Res[icomp] = Res[icomp] - cfg->radiodecay_stoich[i] * rate;
    }
    if (!compute_derivative) continue;
    tempreal = -1.0 * cfg->radiodecay_kf[irxn];
    jcomp = cfg->radiodecay_forward_specid[irxn];
    for (i = p0; i < p1; i++) {
      icomp = cfg->radiodecay_specid[i];
      for (j = 0; j < naq; j++)
        Jac[icomp + j * n] =
            Jac[icomp + j * n] + tempreal * cfg->radiodecay_stoich[i] * c->dtotal[jcomp + j * naq] * L_water;
    }
    if (c->ngas > 0) {
      for (i = p0; i < p1; i++) {
        icomp = cfg->radiodecay_specid[i];
        for (j = 0; j < naq; j++)
          Jac[icomp + j * n] =
              Jac[icomp + j * n] + tempreal * cfg->radiodecay_stoich[i] * c->dtotal_gas[jcomp + j * naq] * L_gas;
      }
    }
    if (have_sorb) {
      for (i = p0; i < p1; i++) {
        icomp = cfg->radiodecay_specid[i];
        for (j = 0; j < naq; j++)
          Jac[icomp + j * n] = Jac[icomp + j * n] +
                               tempreal * cfg->radiodecay_stoich[i] * c->dtotal_sorb_eq[jcomp + j * naq] * c->volume;
      }
    }
  }
}

/* reaction.F90:5316-5460  RGeneral: forward / backward mass-action rates in activities */
static void r_general(cell_t *c, const pfrx_config *cfg, double *Res, double *Jac, int compute_derivative) {
  int naq = c->naq, n = c->n, irxn, i, j;
  double ln_conc[PFRX_MAX_NCOMP * 4], ln_act[PFRX_MAX_NCOMP * 4];
  for (i = 0; i < naq; i++) {
    ln_conc[i] = log(c->pri_molal[i]);
    ln_act[i] = ln_conc[i] + log(c->pri_act_coef[i]);
  }
  for (irxn = 0; irxn < cfg->ngeneral_rxn; irxn++) {
    double kf = cfg->general_kf[irxn], kr = cfg->general_kr[irxn];
    double lnQkf = 0.0, lnQkr = 0.0, Qkf, Qkr, por_den_sat_vol, tempreal;
    int f0 = cfg->general_fwd_ptr[irxn], f1 = cfg->general_fwd_ptr[irxn + 1];
    int b0 = cfg->general_bwd_ptr[irxn], b1 = cfg->general_bwd_ptr[irxn + 1];
    int p0 = cfg->general_ptr[irxn], p1 = cfg->general_ptr[irxn + 1];
    if (kf > 0.0) {
      lnQkf = log(kf);
      for (i = f0; i < f1; i++) lnQkf = lnQkf + cfg->general_fwd_stoich[i] * ln_act[cfg->general_fwd_specid[i]];
      Qkf = exp(lnQkf);
    } else {
      Qkf = 0.0;
    }
    if (kr > 0.0) {
      lnQkr = log(kr);
      for (i = b0; i < b1; i++) lnQkr = lnQkr + cfg->general_bwd_stoich[i] * ln_act[cfg->general_bwd_specid[i]];
      Qkr = exp(lnQkr);
    } else {
      Qkr = 0.0;
    }
    por_den_sat_vol = c->porosity * c->den_kg * c->sat * c->volume;
    for (i = p0; i < p1; i++) {
      int icomp = cfg->general_specid[i];
      Res[icomp] = Res[icomp] - cfg->general_stoich[i] * (Qkf - Qkr) * por_den_sat_vol;
    }
    if (!compute_derivative) continue;
    if (kf > 0.0) {
      for (j = f0; j < f1; j++) {
        int jcomp = cfg->general_fwd_specid[j];
        tempreal = -1.0 * cfg->general_fwd_stoich[j] * exp(lnQkf - ln_conc[jcomp]) * por_den_sat_vol;
        for (i = p0; i < p1; i++) {
          int icomp = cfg->general_specid[i];
          Jac[icomp + jcomp * n] = Jac[icomp + jcomp * n] + cfg->general_stoich[i] * tempreal;
        }
      }
    }
    if (kr > 0.0) {
      for (j = b0; j < b1; j++) {
        int jcomp = cfg->general_bwd_specid[j];
        tempreal = cfg->general_bwd_stoich[j] * exp(lnQkr - ln_conc[jcomp]) * por_den_sat_vol;
        for (i = p0; i < p1; i++) {
          int icomp = cfg->general_specid[i];
          Jac[icomp + jcomp * n] = Jac[icomp + jcomp * n] + cfg->general_stoich[i] * tempreal;
        }
      }
    }
  }
}

/* reaction_microbial.F90:287-602  RMicrobial: rate = k * prod(Monod) * prod(inhibition) * biomass.
 * The biomass derivative is kept as written (:578-583: without the L_water / volume factor). */
static void r_microbial(cell_t *c, const pfrx_config *cfg, double *Res, double *Jac, int compute_derivative) {
  const double PI = 3.14159265359; /* pflotran_constants.F90:92, truncated as there */
  int naq = c->naq, n = c->n, irxn, i, ii, jj;
  double concentration[PFRX_MAX_NCOMP * 4], dconcentration_dmolal[PFRX_MAX_NCOMP * 4];
  double monod[PFRX_MAX_MONOD_TERMS], inhibition[PFRX_MAX_MONOD_TERMS];
  double L_water;
  for (i = 0; i < naq; i++) {
    switch (cfg->microbial_concentration_units) {
      case PFRX_MICROBIAL_MOLALITY:
        dconcentration_dmolal[i] = 1.0;
        break;
      case PFRX_MICROBIAL_ACTIVITY:
        dconcentration_dmolal[i] = c->pri_act_coef[i];
        break;
      default:
        dconcentration_dmolal[i] = c->den_kg * 1.e-3;
        break;
    }
    concentration[i] = c->pri_molal[i] * dconcentration_dmolal[i];
  }
  L_water = c->porosity * c->sat * c->volume * 1.e3;
  for (irxn = 0; irxn < cfg->nmicrobial_rxn; irxn++) {
    int p0 = cfg->microbial_ptr[irxn], p1 = cfg->microbial_ptr[irxn + 1];
    int m0 = cfg->microbial_monod_ptr[irxn], nmonod = cfg->microbial_monod_ptr[irxn + 1] - m0;
    int h0 = cfg->microbial_inhibition_ptr[irxn], ninh = cfg->microbial_inhibition_ptr[irxn + 1] - h0;
    double effective_rate_constant = cfg->microbial_rate_constant[irxn];
    double yield = 0.0, biomass_conc = 0.0, dbiomass_conc_dconc = 0.0;
    double monod_terms = 1.0, inhibition_terms = 1.0, biomass_term = 1.0, rate;
    int ibiomass = cfg->microbial_biomassid[irxn]; /* 1-based like the reference, 0 none */
    int ibio_row = -1;                              /* 0-based row of the biomass unknown */
    if (cfg->microbial_activation_energy) {
      effective_rate_constant = effective_rate_constant * exp(cfg->microbial_activation_energy[irxn] / IDEAL_GAS_CONSTANT *
                                                              (1.0 / 298.15 - 1.0 / (c->temp + 273.15)));
    }
    for (ii = 0; ii < nmonod; ii++) {
      int imonod = m0 + ii;
      double conc = concentration[cfg->microbial_monod_specid[imonod]];
      monod[ii] = (conc - cfg->microbial_monod_Cth[imonod]) /
                  (cfg->microbial_monod_K[imonod] + conc - cfg->microbial_monod_Cth[imonod]);
      monod_terms = monod_terms * monod[ii];
    }
    for (ii = 0; ii < ninh; ii++) {
      int iinh = h0 + ii;
      double conc = concentration[cfg->microbial_inhibition_specid[iinh]];
      double C1 = cfg->microbial_inhibition_C[iinh], C2 = cfg->microbial_inhibition_C2[iinh];
      switch (cfg->microbial_inhibition_type[iinh]) {
        case PFRX_INHIBITION_MONOD:
          inhibition[ii] = C1 / (C1 + conc);
          break;
        case PFRX_INHIBITION_INVERSE_MONOD:
          inhibition[ii] = conc / (C1 + conc);
          break;
        case PFRX_INHIBITION_THRESHOLD:
          inhibition[ii] = 0.5 + copysign(1.0, C1) * atan((conc - fabs(C1)) * C2) / PI;
          break;
        default: { /* SMOOTHSTEP */
          double log10_conc = log10(conc), log10_C = log10(C1), log10_interval = C2;
          double lower = log10_C - 0.5 * log10_interval;
          double z = (log10_conc - lower) / log10_interval, v;
          if (z < 0.0)
            v = 0.0;
          else if (z > 1.0)
            v = 1.0;
          else
            v = 3.0 * (z * z) - 2.0 * (z * z * z);
          inhibition[ii] = v;
        } break;
      }
      inhibition_terms = inhibition_terms * inhibition[ii];
    }
    if (ibiomass != 0) {
      if (ibiomass > 0) {
        biomass_conc = concentration[ibiomass - 1];
        biomass_term = biomass_term * biomass_conc * L_water;
        dbiomass_conc_dconc = dconcentration_dmolal[ibiomass - 1];
        ibio_row = ibiomass - 1;
      } else {
        int k = -ibiomass - 1;
        biomass_conc = c->immobile[k];
        ibio_row = naq + k;
        biomass_term = biomass_term * biomass_conc * c->volume;
        dbiomass_conc_dconc = 1.0;
      }
      yield = cfg->microbial_biomass_yield[irxn];
    } else {
      biomass_term = biomass_term * L_water;
    }
    rate = effective_rate_constant * monod_terms * inhibition_terms * biomass_term;
    for (i = p0; i < p1; i++) {
      int icomp = cfg->microbial_specid[i];
      Res[icomp] = Res[icomp] - cfg->microbial_stoich[i] * rate;
    }
    if (ibio_row >= 0) Res[ibio_row] = Res[ibio_row] - yield * rate;
    if (!compute_derivative) continue;
    for (ii = 0; ii < nmonod; ii++) {
      int imonod = m0 + ii;
      int jcomp = cfg->microbial_monod_specid[imonod];
      double conc = concentration[jcomp], dconc_dmolal = dconcentration_dmolal[jcomp];
      double dR_dX = effective_rate_constant * inhibition_terms * biomass_term, denominator, dX_dc, dR_dc;
      for (jj = 0; jj < ii; jj++) dR_dX = dR_dX * monod[jj];
      for (jj = ii + 1; jj < nmonod; jj++) dR_dX = dR_dX * monod[jj];
      denominator = cfg->microbial_monod_K[imonod] + conc - cfg->microbial_monod_Cth[imonod];
      dX_dc = dconc_dmolal / denominator -
              dconc_dmolal * (conc - cfg->microbial_monod_Cth[imonod]) / (denominator * denominator);
      dR_dc = -1.0 * dR_dX * dX_dc;
      for (i = p0; i < p1; i++) {
        int icomp = cfg->microbial_specid[i];
        Jac[icomp + jcomp * n] = Jac[icomp + jcomp * n] + cfg->microbial_stoich[i] * dR_dc;
      }
      if (ibio_row >= 0) Jac[ibio_row + jcomp * n] = Jac[ibio_row + jcomp * n] + yield * dR_dc;
    }
    for (ii = 0; ii < ninh; ii++) {
      int iinh = h0 + ii;
      int jcomp = cfg->microbial_inhibition_specid[iinh];
      double conc = concentration[jcomp], dconc_dmolal = dconcentration_dmolal[jcomp];
      double C1 = cfg->microbial_inhibition_C[iinh], C2 = cfg->microbial_inhibition_C2[iinh];
      double dR_dX = effective_rate_constant * monod_terms * biomass_term, denominator, dX_dc, dR_dc, tempreal;
      for (jj = 0; jj < ii; jj++) dR_dX = dR_dX * inhibition[jj];
      for (jj = ii + 1; jj < ninh; jj++) dR_dX = dR_dX * inhibition[jj];
      switch (cfg->microbial_inhibition_type[iinh]) {
        case PFRX_INHIBITION_MONOD:
          denominator = C1 + conc;
          dX_dc = -1.0 * dconc_dmolal * C1 / (denominator * denominator);
          break;
        case PFRX_INHIBITION_INVERSE_MONOD:
          denominator = C1 + conc;
          dX_dc = dconc_dmolal / denominator - dconc_dmolal * conc / (denominator * denominator);
          break;
        case PFRX_INHIBITION_THRESHOLD:
          tempreal = (conc - fabs(C1)) * C2;
          dX_dc = copysign(1.0, C1) * (C2 * dconc_dmolal / (1.0 + tempreal * tempreal)) / PI;
          break;
        default: {
          double log10_conc = log10(conc), log10_C = log10(C1), log10_interval = C2;
          double lower = log10_C - 0.5 * log10_interval;
          double z = (log10_conc - lower) / log10_interval;
          if (z < 0.0 || z > 1.0)
            dX_dc = 0.0;
          else
            dX_dc = (6.0 * z - 6.0 * (z * z)) / (log10_interval * conc * LOG_TO_LN) * dconc_dmolal;
        } break;
      }
      dR_dc = -1.0 * dR_dX * dX_dc;
      for (i = p0; i < p1; i++) {
        int icomp = cfg->microbial_specid[i];
        Jac[icomp + jcomp * n] = Jac[icomp + jcomp * n] + cfg->microbial_stoich[i] * dR_dc;
      }
      if (ibio_row >= 0) Jac[ibio_row + jcomp * n] = Jac[ibio_row + jcomp * n] + yield * dR_dc;
    }
    if (ibio_row >= 0) {
      double dR_dbiomass = effective_rate_constant * monod_terms * inhibition_terms;
      dR_dbiomass = -1.0 * dR_dbiomass * dbiomass_conc_dconc;
      for (i = p0; i < p1; i++) {
        int icomp = cfg->microbial_specid[i];
        Jac[icomp + ibio_row * n] = Jac[icomp + ibio_row * n] + cfg->microbial_stoich[i] * dR_dbiomass;
      }
      Jac[ibio_row + ibio_row * n] = Jac[ibio_row + ibio_row * n] + yield * dR_dbiomass;
    }
  }
}

/* reaction_immobile.F90:244-296  RImmobileDecay */
static void r_immobile_decay(cell_t *c, const pfrx_config *cfg, double *Res, double *Jac, int compute_derivative) {
  int n = c->n, irxn;
  double volume = c->volume;
  for (irxn = 0; irxn < cfg->nimmobile_decay_rxn; irxn++) {
    int icomp = cfg->immobile_decay_specid[irxn];
    double rate_constant = cfg->immobile_decay_constant[irxn] * volume;
    double rate = rate_constant * c->immobile[icomp];
    int immobile_id = c->naq + icomp;
    Res[immobile_id] = Res[immobile_id] + rate;
    if (!compute_derivative) continue;
    Jac[immobile_id + immobile_id * n] = Jac[immobile_id + immobile_id * n] + rate_constant;
  }
}

/* reaction.F90:4059-4130  RReaction (dispatch order preserved: mineral, multirate sorption,
 * [kinetic surface complexation], radioactive decay, general, microbial, immobile decay,
 * sandboxes) */
static void r_reaction(cell_t *c, const pfrx_config *cfg, double tran_dt, double *Res, double *Jac, int derivative) {
  if (c->sat < cfg->rt_min_saturation) return;
  if (c->nkin > 0) r_kinetic_mineral(c, cfg, Res, Jac, derivative);
  if (cfg->nkinmrsrfcplxrxn > 0) r_multirate_sorption(c, cfg, tran_dt, Res, Jac, derivative);
  if (cfg->nradiodecay_rxn > 0) r_radioactive_decay(c, cfg, Res, Jac, derivative);
  if (cfg->ngeneral_rxn > 0) r_general(c, cfg, Res, Jac, derivative);
  if (cfg->nmicrobial_rxn > 0) r_microbial(c, cfg, Res, Jac, derivative);
  if (cfg->nimmobile_decay_rxn > 0) r_immobile_decay(c, cfg, Res, Jac, derivative);
  if (n_sandboxes(cfg) > 0) r_sandbox_evaluate(c, cfg, tran_dt, Res, Jac, derivative);
}

/* reaction.F90:5457-5516  RSolve -- WITH back-substitution */
static int r_solve(double *Res, double *Jac, const double *conc, double *update, int ncomp, int use_log_formulation) {
  int indices[PFRX_MAX_NCOMP * 4];
  double rhs[PFRX_MAX_NCOMP * 4];
  int icomp, j, ierror;
  double norm;
  for (icomp = 0; icomp < ncomp; icomp++) {
    double m = 0.0;
    for (j = 0; j < ncomp; j++) m = fmax(m, fabs(Jac[icomp + j * ncomp]));
    norm = fmax(1.0, m);
    norm = 1.0 / norm;
    rhs[icomp] = Res[icomp] * norm;
    for (j = 0; j < ncomp; j++) Jac[icomp + j * ncomp] = Jac[icomp + j * ncomp] * norm;
  }
  if (use_log_formulation) {
    for (icomp = 0; icomp < ncomp; icomp++)
      for (j = 0; j < ncomp; j++) Jac[j + icomp * ncomp] = Jac[j + icomp * ncomp] * conc[icomp];
  }
  ierror = lu_decomposition(Jac, ncomp, indices);
  if (ierror != 0) return ierror;
  if (g_ref_bug_compat) {
    /* the fork returns here without touching `update` (reaction.F90:5501-5504);
     * the golds show it reads as zeros */
    for (icomp = 0; icomp < ncomp; icomp++) update[icomp] = 0.0;
    return 0;
  }
  lu_back_substitution(Jac, ncomp, indices, rhs);
  for (icomp = 0; icomp < ncomp; icomp++) update[icomp] = rhs[icomp];
  return 0;
}

/* reaction.F90:3742-4055  RReact */
static int r_react(cell_t *c, const pfrx_config *cfg, const double *guess, double tran_dt, int *num_iterations_out) {
  int n = c->n, naq = c->naq, nim = c->nim, i, icomp;
  double residual[PFRX_MAX_NCOMP * 4], fixed_accum[PFRX_MAX_NCOMP * 4];
  double initial_total[PFRX_MAX_NCOMP * 4], prev_solution[PFRX_MAX_NCOMP * 4];
  double latest_solution[PFRX_MAX_NCOMP * 4], update[PFRX_MAX_NCOMP * 4], conc[PFRX_MAX_NCOMP * 4];
  double *J = (double *)malloc(sizeof(double) * n * n);
  double maximum_relative_change, two_norm_r, two_norm_r0 = 0.0, rel_residual, ratio, min_ratio;
  int num_iterations = 0, ierror = 0;

  c->option_ierror = 0;
  if (!cfg->use_isothermal) update_temp_dependent_coefs(c, cfg);

  rt_accumulation(c, cfg, fixed_accum);
  if (neqsorb(cfg) > 0) r_accumulation_sorb(c, fixed_accum);

  for (i = 0; i < naq; i++) initial_total[i] = c->total[i];
  for (i = 0; i < nim; i++) initial_total[naq + i] = c->immobile[i];
  for (i = 0; i < naq; i++) c->pri_molal[i] = guess[i];
  for (i = 0; i < nim; i++) c->immobile[i] = guess[naq + i];

  for (;;) {
    num_iterations = num_iterations + 1;
    if (cfg->act_coef_update_frequency == PFRX_ACT_COEF_FREQUENCY_NEWTON_ITER) r_activity_coefficients(c, cfg);
    rt_auxvar_compute(c, cfg);

    if (num_iterations > cfg->maximum_reaction_iterations) {
      ierror = 1;
      for (i = 0; i < naq; i++) c->total[i] = initial_total[i];
      for (i = 0; i < nim; i++) c->immobile[i] = initial_total[naq + i];
      goto done;
    }

    rt_accumulation(c, cfg, residual);
    rt_accumulation_derivative(c, cfg, tran_dt, J);
    if (neqsorb(cfg) > 0) {
      r_accumulation_sorb(c, residual);
      r_accumulation_sorb_derivative(c, tran_dt, J);
    }
    for (i = 0; i < n; i++) residual[i] = (residual[i] - fixed_accum[i]) / tran_dt;

    r_reaction(c, cfg, tran_dt, residual, J, 1);

    if (c->option_ierror != 0) {
      ierror = c->option_ierror;
      goto done;
    }

    two_norm_r = 0.0;
    for (i = 0; i < n; i++) two_norm_r += residual[i] * residual[i];
    two_norm_r = sqrt(two_norm_r);
    if (num_iterations == 1) two_norm_r0 = two_norm_r;
    rel_residual = two_norm_r / two_norm_r0;
    {
      double m = 0.0;
      for (i = 0; i < n; i++) m = fmax(m, fabs(residual[i]));
      /* maxval(abs(residual)) < tol; NaN-safe like the Fortran comparison */
      if (m < cfg->max_residual_tolerance) break;
    }
    if (rel_residual < cfg->max_rel_residual_tolerance) break;

    for (i = 0; i < naq; i++) conc[i] = c->pri_molal[i];
    for (i = 0; i < nim; i++) conc[naq + i] = c->immobile[i];

    if (r_solve(residual, J, conc, update, n, cfg->use_log_formulation) != 0) {
      ierror = 1; /* solve_error branch, reaction.F90:3964-3967: no restore */
      goto done;
    }

    for (i = 0; i < naq; i++) prev_solution[i] = c->pri_molal[i];
    for (i = 0; i < nim; i++) prev_solution[naq + i] = c->immobile[i];

    if (cfg->use_log_formulation) {
      for (i = 0; i < n; i++) {
        update[i] = copysign(1.0, update[i]) * fmin(fabs(update[i]), cfg->max_dlnC_rreact);
        latest_solution[i] = prev_solution[i] * exp(-update[i]);
      }
    } else {
      min_ratio = MAX_DOUBLE;
      for (icomp = 0; icomp < n; icomp++) {
        if (prev_solution[icomp] <= update[icomp]) {
          ratio = fabs(prev_solution[icomp] / update[icomp]);
          if (ratio < min_ratio) min_ratio = ratio;
        }
      }
      if (min_ratio < 1.0)
        for (i = 0; i < n; i++) update[i] = update[i] * min_ratio * 0.99;
      for (i = 0; i < n; i++) latest_solution[i] = prev_solution[i] - update[i];
    }

    maximum_relative_change = 0.0;
    {
      /* maxval(abs((latest-prev)/prev)): gfortran's MAXVAL skips NaNs unless
       * all are NaN; fmax reproduces that for the mixed case */
      double m = -INFINITY;
      int any = 0;
      for (i = 0; i < n; i++) {
        double v = fabs((latest_solution[i] - prev_solution[i]) / prev_solution[i]);
        if (!isnan(v)) {
          m = any ? fmax(m, v) : v;
          any = 1;
        }
      }
      maximum_relative_change = any ? m : (double)NAN;
    }
    if (maximum_relative_change < cfg->max_relative_change_tolerance) break;

    for (i = 0; i < naq; i++) c->pri_molal[i] = latest_solution[i];
    for (i = 0; i < nim; i++) c->immobile[i] = latest_solution[naq + i];
  }

  /* one last update, reaction.F90:4052 */
  rt_auxvar_compute(c, cfg);

done:
  free(J);
  *num_iterations_out = num_iterations;
  return ierror;
}

/* reaction_mineral.F90:1456-1525 MineralUpdateKineticState,
 * reaction_surf_complex.F90:1107-1145 RSrfCplxMRUpdateKinState,
 * reaction.F90:5935-5972 RUpdateKineticState */
static int r_update_kinetic_state(cell_t *c, const pfrx_config *cfg, double tran_dt) {
  int kinetic_state_updated = 0, imnrl, irxn, irate, i;
  if (c->nkin > 0) {
    double res[PFRX_MAX_NCOMP * 4];
    double jac[1];
    kinetic_state_updated = 1;
    for (i = 0; i < c->n; i++) res[i] = 0.0;
    r_kinetic_mineral(c, cfg, res, jac, 0);
    for (imnrl = 0; imnrl < c->nkin; imnrl++) {
      double delta_volfrac = c->mnrl_rate[imnrl] * cfg->kinmnrl_molar_vol[imnrl] * tran_dt;
      c->mnrl_volfrac[imnrl] = c->mnrl_volfrac[imnrl] + delta_volfrac;
      if (c->mnrl_volfrac[imnrl] < 0.0) c->mnrl_volfrac[imnrl] = 0.0;
    }
  }
  for (irxn = 0; irxn < cfg->nkinmrsrfcplxrxn; irxn++) {
    int r0 = cfg->kinmr_rate_ptr[irxn], r1 = cfg->kinmr_rate_ptr[irxn + 1];
    int base = c->naq * (r0 + irxn);
    kinetic_state_updated = 1;
    for (irate = r0; irate < r1; irate++) {
      double kdt = cfg->kinmr_rate[irate] * tran_dt;
      double one_plus_kdt = 1.0 + kdt;
      double *S = c->kinmr_total_sorb + base + c->naq * (irate - r0 + 1);
      for (i = 0; i < c->naq; i++)
        S[i] = (S[i] + kdt * cfg->kinmr_frac[irate] * c->kinmr_total_sorb[base + i]) / one_plus_kdt;
    }
  }
  if (cfg->calcite) { /* CalciteUpdateKineticState, reaction_sandbox_calcite.F90:369-410 */
    int imnrl = cfg->calcite->mineral_id;
    double delta_volfrac = c->sandbox_aux * cfg->kinmnrl_molar_vol[imnrl] * tran_dt;
    c->mnrl_volfrac[imnrl] = c->mnrl_volfrac[imnrl] + delta_volfrac;
    if (c->mnrl_volfrac[imnrl] < 0.0) c->mnrl_volfrac[imnrl] = 0.0;
  }
  if (n_sandboxes(cfg) > 0) kinetic_state_updated = 1; /* any sandbox => true, reaction.F90:5965 */
  return kinetic_state_updated;
}

/* reaction.F90:3564-3738  RStep */
static int r_step(cell_t *c, const pfrx_config *cfg, double *guess, double target_time, int *num_sub_steps_out,
                  int *num_iterations_out, int *num_kinetic_state_updates_out, int *had_cut) {
  int n = c->n, naq = c->naq, nim = c->nim, i;
  int value_is_initially_small[PFRX_MAX_NCOMP * 4];
  double initial_small_value[PFRX_MAX_NCOMP * 4];
  int num_inner_iterations, num_constant_timesteps_after_cut = 0, num_cuts = 0;
  int num_kinetic_state_updates = 0, num_sub_steps = 0, num_iterations = 0, ierror = 0;
  double cumulative_time = 0.0, tran_dt = target_time;

  *had_cut = 0;
  if (!cfg->use_full_geochemistry) {
    for (i = 0; i < naq; i++) c->pri_molal[i] = c->total[i] / c->den_kg * 1.e3;
    *num_sub_steps_out = 0;
    *num_iterations_out = 0;
    *num_kinetic_state_updates_out = 0;
    return 0;
  }
  for (i = 0; i < n; i++) value_is_initially_small[i] = 0;
  for (i = 0; i < naq; i++) {
    if (c->total[i] <= 1.e-40) {
      value_is_initially_small[i] = 1;
      initial_small_value[i] = c->total[i];
      c->total[i] = 1.e-40;
    }
  }
  for (i = 0; i < nim; i++) {
    if (c->immobile[i] <= 1.e-40) {
      value_is_initially_small[naq + i] = 1;
      initial_small_value[naq + i] = c->immobile[i];
      c->immobile[i] = 1.e-40;
    }
  }
  if (cfg->use_total_as_guess)
    for (i = 0; i < naq; i++) guess[i] = c->total[i]; /* guess(:) = total(:,1) */

  for (;;) {
    if (cumulative_time >= target_time) break;
    ierror = r_react(c, cfg, guess, tran_dt, &num_inner_iterations);
    num_iterations = num_iterations + num_inner_iterations;
    if (ierror != 0) {
      num_cuts = num_cuts + 1;
      *had_cut = 1;
      if (num_cuts > cfg->maximum_reaction_cuts) {
        ierror = 1;
        /* reference returns here WITHOUT restoring the small values */
        *num_sub_steps_out = num_sub_steps;
        *num_iterations_out = num_iterations;
        *num_kinetic_state_updates_out = num_kinetic_state_updates;
        return ierror;
      }
      tran_dt = 0.5 * tran_dt;
      num_constant_timesteps_after_cut = 0;
    } else {
      int updated = r_update_kinetic_state(c, cfg, tran_dt);
      cumulative_time = cumulative_time + tran_dt;
      num_sub_steps = num_sub_steps + 1;
      num_constant_timesteps_after_cut = num_constant_timesteps_after_cut + 1;
      if (updated) num_kinetic_state_updates = num_kinetic_state_updates + 1;
      for (i = 0; i < naq; i++) guess[i] = c->pri_molal[i];
      for (i = 0; i < nim; i++) guess[naq + i] = c->immobile[i];
      if (num_constant_timesteps_after_cut >= 4) {
        num_cuts = num_cuts - 1;
        tran_dt = fmin(2.0 * tran_dt, target_time - cumulative_time);
      }
    }
  }
  for (i = 0; i < naq; i++)
    if (value_is_initially_small[i]) c->total[i] = initial_small_value[i];
  for (i = 0; i < nim; i++)
    if (value_is_initially_small[naq + i]) c->immobile[i] = initial_small_value[naq + i];

  *num_sub_steps_out = num_sub_steps;
  *num_iterations_out = num_iterations;
  *num_kinetic_state_updates_out = num_kinetic_state_updates;
  return ierror;
}

/* ------------------------------------------------------------------------ */
/* The OS cell loop pmc_subsurface_osrt.F90:349-378 over a SoA shard.         */
typedef struct {
  const pfrx_config *cfg;
  const pfrx_state *st;
  int64_t c0, c1;
  double tran_dt;
  pfrx_step_result res;
} job_t;

static void *job_run(void *arg) {
  job_t *jb = (job_t *)arg;
  const pfrx_config *cfg = jb->cfg;
  const pfrx_state *st = jb->st;
  cell_t c;
  double guess[PFRX_MAX_NCOMP * 4];
  int64_t ic;
  int i;
  pfrx_step_result *r = &jb->res;
  memset(r, 0, sizeof(*r));
  r->first_failed_cell = -1;
  cell_init(&c, cfg);
  for (ic = jb->c0; ic < jb->c1; ic++) {
    int nss = 0, nit = 0, nku = 0, had_cut = 0, ierr;
    if (st->imat && st->imat[ic] <= 0) {
      if (st->num_sub_steps) st->num_sub_steps[ic] = 0;
      if (st->num_iterations) st->num_iterations[ic] = 0;
      if (st->num_kinetic_state_updates) st->num_kinetic_state_updates[ic] = 0;
      if (st->ierror) st->ierror[ic] = 0;
      continue;
    }
    cell_gather(&c, cfg, st, ic);
    /* guess has to be free ion concentration, :356-362 */
    for (i = 0; i < c.naq; i++) guess[i] = c.pri_molal[i];
    for (i = 0; i < c.nim; i++) guess[c.naq + i] = c.immobile[i];
    ierr = r_step(&c, cfg, guess, jb->tran_dt, &nss, &nit, &nku, &had_cut);
    cell_scatter(&c, st, ic);
    if (st->num_sub_steps) st->num_sub_steps[ic] = nss;
    if (st->num_iterations) st->num_iterations[ic] = nit;
    if (st->num_kinetic_state_updates) st->num_kinetic_state_updates[ic] = nku;
    if (st->ierror) st->ierror[ic] = ierr;
    r->ncell_active++;
    r->sum_newton_iterations += nit;
    if (nit > r->max_newton_iterations) r->max_newton_iterations = nit;
    if (nku > r->max_num_kinetic_state_updates) r->max_num_kinetic_state_updates = nku;
    if (nss > r->max_sub_steps) r->max_sub_steps = nss;
    if (ierr > r->rstep_error) r->rstep_error = ierr;
    if (had_cut) r->num_cut_cells++;
    if (ierr != 0 && r->first_failed_cell < 0) r->first_failed_cell = ic;
  }
  cell_free(&c);
  #ifdef PFRX_ORACLE_COUNTING
  pfrx_oracle_ops_flush(); /* this thread's operation count into the total */
#endif
  return NULL;
}

static void merge_result(pfrx_step_result *a, const pfrx_step_result *b) {
  a->ncell_active += b->ncell_active;
  a->sum_newton_iterations += b->sum_newton_iterations;
  if (b->max_newton_iterations > a->max_newton_iterations) a->max_newton_iterations = b->max_newton_iterations;
  if (b->max_num_kinetic_state_updates > a->max_num_kinetic_state_updates)
    a->max_num_kinetic_state_updates = b->max_num_kinetic_state_updates;
  if (b->rstep_error > a->rstep_error) a->rstep_error = b->rstep_error;
  if (b->max_sub_steps > a->max_sub_steps) a->max_sub_steps = b->max_sub_steps;
  a->num_cut_cells += b->num_cut_cells;
  if (b->first_failed_cell >= 0 && (a->first_failed_cell < 0 || b->first_failed_cell < a->first_failed_cell))
    a->first_failed_cell = b->first_failed_cell;
}

/* All cells are processed (the reference stops its rank at the first failing
 * cell, :370; first_failed_cell lets a caller reproduce that).  nthreads
 * workers take static contiguous ranges, mimicking `mpirun -n P` ownership. */
int pfrx_oracle_rstep(const pfrx_config *cfg, int64_t ncell, const pfrx_state *st, double tran_dt,
                      pfrx_step_result *out, int nthreads) {
  int t;
  if (!cfg || !st || cfg->naqcomp + cfg->nimcomp > PFRX_MAX_NCOMP * 4) return PFRX_E_INVALID;
  if (nthreads < 1) nthreads = 1;
  if ((int64_t)nthreads > ncell) nthreads = ncell > 0 ? (int)ncell : 1;
  job_t *jobs = (job_t *)calloc(nthreads, sizeof(job_t));
  pthread_t *th = (pthread_t *)calloc(nthreads, sizeof(pthread_t));
  for (t = 0; t < nthreads; t++) {
    jobs[t].cfg = cfg;
    jobs[t].st = st;
    jobs[t].tran_dt = tran_dt;
    jobs[t].c0 = ncell * t / nthreads;
    jobs[t].c1 = ncell * (t + 1) / nthreads;
  }
  if (nthreads == 1) {
    job_run(&jobs[0]);
  } else {
    for (t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, job_run, &jobs[t]);
    for (t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  }
  memset(out, 0, sizeof(*out));
  out->first_failed_cell = -1;
  for (t = 0; t < nthreads; t++) merge_result(out, &jobs[t].res);
  free(jobs);
  free(th);
  return PFRX_OK;
}

void pfrx_oracle_set_ref_bug_compat(int on) { g_ref_bug_compat = on; }

/* ------------------------------------------------------------------------ */
/* Single-cell entry points used by tests/ to replay the reference's GIRT     */
/* batch golds and to check Jacobians.               */

/* RActivityCoefficients on cell ic */
int pfrx_oracle_activity(const pfrx_config *cfg, const pfrx_state *st, int64_t ic) {
  cell_t c;
  cell_init(&c, cfg);
  cell_gather(&c, cfg, st, ic);
  if (!cfg->use_isothermal) update_temp_dependent_coefs(&c, cfg);
  r_activity_coefficients(&c, cfg);
  cell_scatter(&c, st, ic);
  int e = c.option_ierror;
  cell_free(&c);
  return e;
}

/* reaction.F90:1328-2117  ReactionEquilibrateConstraint for the constraint values conc[i*ldc + ic] of
 * cell ic (types / reaction tables in pfrx_constraint, include/pfrx.h).  The state of the cell supplies
 * den_kg, temp, porosity ... and receives the speciation.  Returns 0, or 1 singular Jacobian, 2 non-positive
 * concentration, 3 iteration limit (the reference stops the run in each of the three). */
int pfrx_oracle_equilibrate_constraint(const pfrx_config *cfg, const pfrx_constraint *k, const pfrx_state *st,
                                       int64_t ic, const double *conc_in, int64_t ldc, int *num_iterations_out) {
  cell_t c;
  int naq = cfg->naqcomp, icomp, jcomp, kcomp, p, ierror = 0;
  int num_iterations = 0, num_it_act_coef_turned_on = 0, compute_activity_coefs = 0, use_log_formulation;
  int max_it = k->max_iterations > 0 ? k->max_iterations : 10000;
  double conc[PFRX_MAX_NCOMP * 4], free_conc[PFRX_MAX_NCOMP * 4], total_conc[PFRX_MAX_NCOMP * 4];
  double Res[PFRX_MAX_NCOMP * 4], update[PFRX_MAX_NCOMP * 4], prev_molal[PFRX_MAX_NCOMP * 4];
  double *Jac;
  double convert_molal_to_molar, convert_molar_to_molal, maximum_residual, maximum_relative_change, lnQK;
  cell_init(&c, cfg);
  cell_gather(&c, cfg, st, ic);
  Jac = (double *)malloc(sizeof(double) * (naq * naq + 1));
  if (k->initialize_with_molality) {
    convert_molal_to_molar = c.den_kg / 1000.0;
    convert_molar_to_molal = 1.0;
  } else {
    convert_molal_to_molar = 1.0;
    convert_molar_to_molal = 1000.0 / c.den_kg;
  }
  for (icomp = 0; icomp < naq; icomp++) conc[icomp] = conc_in[icomp * ldc + ic];
  if (num_iterations_out) *num_iterations_out = 0;
  if (!cfg->use_full_geochemistry) { /* :1472-1480 */
    for (icomp = 0; icomp < naq; icomp++) {
      c.pri_molal[icomp] = conc[icomp] * convert_molar_to_molal;
      c.total[icomp] = conc[icomp] * convert_molal_to_molar;
    }
    cell_scatter(&c, st, ic);
    free(Jac);
    cell_free(&c);
    return 0;
  }
  if (!cfg->use_isothermal) update_temp_dependent_coefs(&c, cfg);
  /* a fresh rt_auxvar (RTAuxVarInit): unit activity coefficients, no complexes */
  c.ln_act_h2o = 0.0;
  for (icomp = 0; icomp < naq; icomp++) c.pri_act_coef[icomp] = 1.0;
  for (p = 0; p < c.ncplx; p++) {
    c.sec_act_coef[p] = 1.0;
    c.sec_molal[p] = 0.0;
  }
  for (icomp = 0; icomp < naq; icomp++) {
    free_conc[icomp] = 1.e-9;
    total_conc[icomp] = 0.0;
    switch (k->type[icomp]) {
      case PFRX_CONSTRAINT_NULL:
      case PFRX_CONSTRAINT_TOTAL: total_conc[icomp] = conc[icomp] * convert_molal_to_molar; break;
      case PFRX_CONSTRAINT_FREE: free_conc[icomp] = conc[icomp] * convert_molar_to_molal; break;
      case PFRX_CONSTRAINT_LOG: free_conc[icomp] = pow(10.0, conc[icomp]) * convert_molar_to_molal; break;
      case PFRX_CONSTRAINT_CHARGE_BAL: free_conc[icomp] = conc[icomp] * convert_molar_to_molal; break;
      case PFRX_CONSTRAINT_PH: free_conc[icomp] = pow(10.0, -conc[icomp]); break;
      case PFRX_CONSTRAINT_MINERAL: free_conc[icomp] = conc[icomp] * convert_molar_to_molal; break;
      case PFRX_CONSTRAINT_GAS:
        if (conc[icomp] <= 0.0) conc[icomp] = pow(10.0, conc[icomp]);
        break;
      default: break;
    }
  }
  for (icomp = 0; icomp < naq; icomp++) c.pri_molal[icomp] = free_conc[icomp];
  for (;;) {
    for (icomp = 0; icomp < naq; icomp++)
      if (k->type[icomp] == PFRX_CONSTRAINT_FREE || k->type[icomp] == PFRX_CONSTRAINT_LOG)
        c.pri_molal[icomp] = free_conc[icomp];
    if (cfg->act_coef_update_frequency != PFRX_ACT_COEF_FREQUENCY_OFF && compute_activity_coefs)
      r_activity_coefficients(&c, cfg);
    rt_auxvar_compute(&c, cfg); /* RTotal */
    for (icomp = 0; icomp < naq * naq; icomp++) Jac[icomp] = 0.0;
    for (icomp = 0; icomp < naq; icomp++) {
      switch (k->type[icomp]) {
        case PFRX_CONSTRAINT_NULL:
        case PFRX_CONSTRAINT_TOTAL:
          Res[icomp] = c.total[icomp] - total_conc[icomp];
          for (jcomp = 0; jcomp < naq; jcomp++) Jac[icomp + jcomp * naq] = c.dtotal[icomp + jcomp * naq];
          break;
        case PFRX_CONSTRAINT_FREE:
        case PFRX_CONSTRAINT_LOG:
          Res[icomp] = 0.0;
          Jac[icomp + icomp * naq] = 1.0;
          break;
        case PFRX_CONSTRAINT_CHARGE_BAL:
          Res[icomp] = 0.0;
          for (jcomp = 0; jcomp < naq; jcomp++) {
            Res[icomp] = Res[icomp] + cfg->primary_spec_Z[jcomp] * c.total[jcomp];
            for (kcomp = 0; kcomp < naq; kcomp++)
              Jac[icomp + jcomp * naq] =
                  Jac[icomp + jcomp * naq] + cfg->primary_spec_Z[kcomp] * c.dtotal[kcomp + jcomp * naq];
          }
          break;
        case PFRX_CONSTRAINT_PH:
          Res[icomp] = 0.0;
          c.pri_molal[icomp] = pow(10.0, -conc[icomp]) / c.pri_act_coef[icomp];
          Jac[icomp + icomp * naq] = 1.0;
          break;
        case PFRX_CONSTRAINT_MINERAL:
        case PFRX_CONSTRAINT_GAS: {
          double logK = k->eq_logK[icomp];
          if (!cfg->use_isothermal && k->eq_logK_coef) interpolate_logK(CFGP(k->eq_logK_coef) + 5 * icomp, &logK, c.temp, 1);
          lnQK = -logK * LOG_TO_LN;
          if (k->eq_h2o_stoich[icomp] != 0.0) lnQK = lnQK + k->eq_h2o_stoich[icomp] * c.ln_act_h2o;
          for (p = k->eq_ptr[icomp]; p < k->eq_ptr[icomp + 1]; p++) {
            int comp_id = k->eq_spec[p];
            lnQK = lnQK + k->eq_stoich[p] * log(c.pri_molal[comp_id] * c.pri_act_coef[comp_id]);
          }
          Res[icomp] = k->type[icomp] == PFRX_CONSTRAINT_GAS ? lnQK - log(conc[icomp]) : lnQK;
          for (p = k->eq_ptr[icomp]; p < k->eq_ptr[icomp + 1]; p++) {
            int comp_id = k->eq_spec[p];
            Jac[icomp + comp_id * naq] = k->eq_stoich[p] / c.pri_molal[comp_id];
          }
          break;
        }
        default: break;
      }
    }
    maximum_residual = 0.0;
    for (icomp = 0; icomp < naq; icomp++) maximum_residual = fmax(maximum_residual, fabs(Res[icomp]));
    if (cfg->use_log_formulation) {
      if (num_iterations > 3 && num_iterations < 9)
        use_log_formulation = (num_iterations % 2 == 0);
      else
        use_log_formulation = 1;
    } else {
      use_log_formulation = 0;
    }
    if (r_solve(Res, Jac, c.pri_molal, update, naq, use_log_formulation) != 0) {
      ierror = 1;
      break;
    }
    for (icomp = 0; icomp < naq; icomp++) prev_molal[icomp] = c.pri_molal[icomp];
    if (use_log_formulation) {
      for (icomp = 0; icomp < naq; icomp++) {
        update[icomp] = copysign(1.0, update[icomp]) * fmin(fabs(update[icomp]), cfg->max_dlnC_rreact);
        c.pri_molal[icomp] = c.pri_molal[icomp] * exp(-update[icomp]);
      }
    } else {
      double min_ratio = 1.7976931348623157e308, ratio;
      for (icomp = 0; icomp < naq; icomp++) {
        if (prev_molal[icomp] <= update[icomp]) {
          ratio = fabs(prev_molal[icomp] / update[icomp]);
          if (ratio < min_ratio) min_ratio = ratio;
        }
      }
      if (min_ratio <= 1.0)
        for (icomp = 0; icomp < naq; icomp++) update[icomp] = update[icomp] * min_ratio * 0.99;
      for (icomp = 0; icomp < naq; icomp++) c.pri_molal[icomp] = prev_molal[icomp] - update[icomp];
    }
    num_iterations = num_iterations + 1;
    {
      int bad = 0;
      for (icomp = 0; icomp < naq; icomp++)
        if (!(c.pri_molal[icomp] > 0.0)) bad = 1;
      if (bad) {
        ierror = 2;
        break;
      }
    }
    maximum_relative_change = 0.0;
    for (icomp = 0; icomp < naq; icomp++)
      maximum_relative_change =
          fmax(maximum_relative_change, fabs((c.pri_molal[icomp] - prev_molal[icomp]) / prev_molal[icomp]));
    if (num_iterations >= max_it) {
      ierror = 3;
      break;
    }
    if (maximum_residual < cfg->max_residual_tolerance &&
        maximum_relative_change < cfg->max_relative_change_tolerance) {
      if (compute_activity_coefs && num_iterations - num_it_act_coef_turned_on > 1) break;
      if (!compute_activity_coefs) num_it_act_coef_turned_on = num_iterations;
      compute_activity_coefs = 1;
    }
  }
  if (num_iterations_out) *num_iterations_out = num_iterations;
  if (ierror == 0) {
    /* once equilibrated, the sorbed concentrations (:2036-2071) */
    if (neqsorb(cfg) > 0) r_total_sorb(&c, cfg);
    if (cfg->nkinmrsrfcplxrxn > 0) {
      double total_sorb_eq[PFRX_MAX_NCOMP * 4];
      double *dts = (double *)malloc(sizeof(double) * (naq * naq + 1));
      int q, irate;
      for (q = 0; q < cfg->nkinmrsrfcplxrxn; q++) { /* RTotalSorbMultiRateAsEQ */
        int irxn = cfg->kinmrsrfcplxrxn_to_srfcplxrxn[q];
        int r0 = cfg->kinmr_rate_ptr[q], r1 = cfg->kinmr_rate_ptr[q + 1];
        int base = naq * (r0 + q);
        for (icomp = 0; icomp < naq; icomp++) total_sorb_eq[icomp] = 0.0;
        for (icomp = 0; icomp < naq * naq; icomp++) dts[icomp] = 0.0;
        r_total_sorb_eq_surf_cplx1(&c, cfg, irxn, &c.free_site[irxn], NULL, total_sorb_eq, dts);
        for (icomp = 0; icomp < naq; icomp++) {
          c.kinmr_total_sorb[base + icomp] = total_sorb_eq[icomp];
          for (irate = r0; irate < r1; irate++)
            c.kinmr_total_sorb[base + naq * (irate - r0 + 1) + icomp] = cfg->kinmr_frac[irate] * total_sorb_eq[icomp];
        }
      }
      free(dts);
    }
    cell_scatter(&c, st, ic);
  }
  free(Jac);
  cell_free(&c);
  return ierror;
}

/* RTAuxVarCompute on cell ic (totals, sec_molal, sorbed totals) */
int pfrx_oracle_auxvar_compute(const pfrx_config *cfg, const pfrx_state *st, int64_t ic) {
  cell_t c;
  cell_init(&c, cfg);
  cell_gather(&c, cfg, st, ic);
  if (!cfg->use_isothermal) update_temp_dependent_coefs(&c, cfg);
  rt_auxvar_compute(&c, cfg);
  cell_scatter(&c, st, ic);
  cell_free(&c);
  return 0;
}

/* GIRT single-cell residual and Jacobian (reactive_transport.F90:2398-2438,
 * :2599-2642, :3088-3303):  Res = A(c)/dt + R(c)   (caller subtracts
 * fixed_accum/dt), Jac = dA/dc/dt + dR/dc, column-major Jac[i + j*n].
 * Also returns A(c) itself in accum[] when non-NULL.  State is updated the
 * way RTAuxVarCompute does (total, sec_molal, sorbed, mnrl_rate). */
int pfrx_oracle_girt_residual(const pfrx_config *cfg, const pfrx_state *st, int64_t ic, double tran_dt, double *Res,
                              double *Jac, double *accum) {
  cell_t c;
  int n, i;
  cell_init(&c, cfg);
  cell_gather(&c, cfg, st, ic);
  n = c.n;
  if (!cfg->use_isothermal) update_temp_dependent_coefs(&c, cfg);
  rt_auxvar_compute(&c, cfg);
  rt_accumulation(&c, cfg, Res);
  rt_accumulation_derivative(&c, cfg, tran_dt, Jac);
  if (neqsorb(cfg) > 0) {
    r_accumulation_sorb(&c, Res);
    r_accumulation_sorb_derivative(&c, tran_dt, Jac);
  }
  if (accum)
    for (i = 0; i < n; i++) accum[i] = Res[i];
  for (i = 0; i < n; i++) Res[i] = Res[i] / tran_dt;
  r_reaction(&c, cfg, tran_dt, Res, Jac, 1);
  cell_scatter(&c, st, ic);
  int e = c.option_ierror;
  cell_free(&c);
  return e;
}

/* RReaction + RReactionDerivative alone (reaction.F90:4059-4208) on cell ic, the way the
 * GIRT / ELM caller uses them (reactive_transport.F90:2627, 3288): rt_auxvar as it stands,
 * Res / Jac start from zero, Jac column-major Jac[i + j*n]; mnrl_rate is written back. */
int pfrx_oracle_reaction(const pfrx_config *cfg, const pfrx_state *st, int64_t ic, double tran_dt, double *Res,
                         double *Jac) {
  cell_t c;
  int n, i;
  cell_init(&c, cfg);
  cell_gather(&c, cfg, st, ic);
  n = c.n;
  for (i = 0; i < n; i++) Res[i] = 0.0;
  for (i = 0; i < n * n; i++) Jac[i] = 0.0;
  /* the GIRT caller has just run RTAuxVarCompute (reactive_transport.F90:2599-2642):
   * the sandboxes that read rt_auxvar%aqueous%dtotal need it here too, radioactive decay reads
   * total and dtotal, mineral prefactors on secondary species read sec_molal */
  if (cfg->somdec || cfg->nitrif || cfg->denitr || cfg->plantn || cfg->langmuir || cfg->cndegas ||
      cfg->nradiodecay_rxn > 0 || cfg->kinmnrl_num_prefactors)
    rt_auxvar_compute(&c, cfg);
  r_reaction(&c, cfg, tran_dt, Res, Jac, 1);
  for (i = 0; i < c.nkin; i++) st->mnrl_rate[i * st->ld + ic] = c.mnrl_rate[i];
  if (st->somdec_nc)
    for (i = 0; i < c.nsomdec_nc; i++) st->somdec_nc[i * st->ld + ic] = c.somdec_nc[i];
  if (st->sandbox_aux) st->sandbox_aux[ic] = c.sandbox_aux; /* CalciteEvaluate leaves its rate in auxiliary_data */
  cell_free(&c);
  return 0;
}

/* RUpdateKineticState on cell ic over tran_dt */
int pfrx_oracle_update_kinetic_state(const pfrx_config *cfg, const pfrx_state *st, int64_t ic, double tran_dt) {
  cell_t c;
  int u;
  cell_init(&c, cfg);
  cell_gather(&c, cfg, st, ic);
  if (!cfg->use_isothermal) update_temp_dependent_coefs(&c, cfg);
  u = r_update_kinetic_state(&c, cfg, tran_dt);
  cell_scatter(&c, st, ic);
  cell_free(&c);
  return u;
}

/* RSolve (row scaling, optional log scaling, LU, back-substitution) for KATs */
int pfrx_oracle_rsolve(double *Res, double *Jac, const double *conc, double *update, int ncomp,
                       int use_log_formulation) {
  if (ncomp > PFRX_MAX_NCOMP * 4) return -1;
  return r_solve(Res, Jac, conc, update, ncomp, use_log_formulation);
}

/* bare LU + back-substitution for KATs (utility.F90:597-735) */
int pfrx_oracle_lu_solve(double *A, int N, double *B) {
  int indx[PFRX_MAX_NCOMP * 4];
  int e;
  if (N > PFRX_MAX_NCOMP * 4) return -1;
  e = lu_decomposition(A, N, indx);
  if (e) return e;
  lu_back_substitution(A, N, indx, B);
  return 0;
}
