// Host driver for the GENERATED routines of csrc/pfrx_spec2.cuh (test infrastructure).
//
// Built by tests/test_spec2_host.py as
//   g++ -O1 -ffp-contract=off -shared -fPIC -DS2_HOST -x c++ -include <generated .cu> spec2_host_driver.cpp
// so that spec2_eval / spec2_solve / spec2_update -- the arithmetic the GPU kernel executes --
// run on the CPU for a few cells and are compared with the oracle without a GPU.  The RStep /
// RReact control flow below restates the state machine of pfrx_spec_kernel as nested loops
// (reaction.F90:3564-4055).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

extern "C" int s2_host_rstep(const pfrx_state *ps, long long ncell, double tran_dt, const SpecParams *prm_) {
  const SpecParams prm = *prm_;
  DevState st;
  st.ld = ps->ld;
  st.total = ps->total; st.pri_molal = ps->pri_molal; st.immobile = ps->immobile;
  st.pri_act_coef = ps->pri_act_coef; st.sec_act_coef = ps->sec_act_coef; st.sec_molal = ps->sec_molal;
  st.ln_act_h2o = ps->ln_act_h2o; st.mnrl_volfrac = ps->mnrl_volfrac; st.mnrl_area = ps->mnrl_area;
  st.mnrl_rate = ps->mnrl_rate; st.free_site = ps->srfcplxrxn_free_site_conc; st.eqsrfcplx_conc = ps->eqsrfcplx_conc;
  st.total_sorb_eq = ps->total_sorb_eq; st.kinmr = ps->kinmr_total_sorb; st.den_kg = ps->den_kg; st.sat = ps->sat;
  st.temp = ps->temp; st.porosity = ps->porosity; st.volume = ps->volume;
  st.soil_particle_density = ps->soil_particle_density; st.imat = ps->imat; st.num_sub_steps = ps->num_sub_steps;
  st.num_iterations = ps->num_iterations; st.num_kinetic_state_updates = ps->num_kinetic_state_updates;
  st.ierror = ps->ierror;
  constexpr int N = SPEC_N, NAQ = SPEC_NAQ, NIM = N - NAQ;
  const long long ld = st.ld;
  std::vector<double> slice((size_t)S2_SLOTS * 32, 0.0);
  double *W = slice.data();
  for (long long cell = 0; cell < ncell; cell++) {
    int nss = 0, nit = 0, nku = 0;
    bool aborted = false;
    if (!(st.imat && st.imat[cell] <= 0)) {
      Spec2Cell s;
      const double den_kg = st.den_kg[cell], sat = st.sat[cell], por = st.porosity[cell];
      s.vol = st.volume[cell];
      s.temp = st.temp[cell];
      const double spd = st.soil_particle_density ? st.soil_particle_density[cell] : 0.0;
      const double law = st.ln_act_h2o ? st.ln_act_h2o[cell] : 0.0;
      s.aw = (law == 0.0) ? 1.0 : exp(law);
      s.denL = den_kg * 1.e-3;
      s.dry = sat < prm.min_sat;
      s.psv = por * sat * 1000.0 * s.vol;
      s.rock = spd * (1.0 - por);
      s.Iact = 0.0;
      s.store = true;
      spec2_isec(s, st, cell);
      if (!SPEC_ACT_UPD) spec2_load_frozen(W, st, cell);
      unsigned small_mask = 0u;
      double small_val[N], gimm[NIM > 0 ? NIM : 1];
      for (int i = 0; i < N; i++) {
        double *p = (i < NAQ) ? st.total + i * ld + cell : st.immobile + (i - NAQ) * ld + cell;
        if (i >= NAQ) gimm[i - NAQ] = *p;
        if (*p <= 1.e-40) {
          small_mask |= 1u << i;
          small_val[i] = *p;
          *p = 1.e-40;
        }
      }
      double cumulative = 0.0, dt = tran_dt;
      int ncuts = 0, nconst = 0;
      while (cumulative < tran_dt) {
        for (int i = 0; i < N; i++) SW(S2_OFF_C + i) = (i < NAQ) ? st.pri_molal[i * ld + cell] : gimm[i >= NAQ ? i - NAQ : 0];
        spec2_begin(W, s, st, cell);
        s.rdt = 1.0 / dt;
        int its = 0;
        bool conv = false, fail = false, solve_error = false;
        double norm0 = 0.0, res[N], tv[S2_NTV], ev[S2_NEV];
        for (;;) {
          its++;
          const bool over = its > prm.max_its;
          s.rates = !over;
          spec2_eval(res, tv, ev, s, W, st, cell);
          double mabs = 0.0, ss = 0.0;
          for (int i = 0; i < N; i++) {
            mabs = fmax(mabs, fabs(res[i]));
            ss += res[i] * res[i];
          }
          const double nrm = sqrt(ss);
          if (its == 1) norm0 = nrm;
          const double rel = nrm / norm0;
          conv = (mabs < prm.tol_res) || (rel < prm.tol_relres);
          if (getenv("S2_TRACE") && cell == atoll(getenv("S2_TRACE")))
            printf("cell %lld its %d dt %g mabs %.6e rel %.6e c4 %.17e res4 %.6e\n", cell, its, dt, mabs, rel, SW(S2_OFF_C + 4), res[4]);
#if !SPEC_SYM
          if (getenv("S2_TRACE") && cell == atoll(getenv("S2_TRACE"))) {
            printf("   J4: ");
            for (int j = 0; j < SPEC_NC; j++) printf("%.10e ", W[JX(spec_cmap(4), j)]);
            printf("\n   res: ");
            for (int j = 0; j < N; j++) printf("%.10e ", res[j]);
            printf("\n");
          }
#endif
          if (over) { fail = true; break; }
          if (conv) break;
          if (!spec2_solve(W, res, ev, s)) { fail = true; solve_error = true; break; }
          double cn[N];
          const double maxrel = spec2_update(W, res, prm.max_dlnC, cn);
          if (getenv("S2_TRACE") && cell == atoll(getenv("S2_TRACE"))) printf("   update u4 %.6e maxrel %.6e\n", res[4], maxrel);
          if (maxrel < prm.tol_relchange) { conv = true; break; }
          for (int i = 0; i < N; i++) SW(S2_OFF_C + i) = cn[i];
        }
        nit += its;
        if (fail) {
          spec2_store_totals(tv, W, s, st, cell, solve_error, true);
          if (solve_error)
            for (int i = NAQ; i < N; i++) st.immobile[(i - NAQ) * ld + cell] = SW(S2_OFF_C + i);
          ncuts++;
          if (ncuts > prm.max_cuts) { aborted = true; break; }
          dt = 0.5 * dt;
          nconst = 0;
        } else {
          spec2_store_totals(tv, W, s, st, cell, true, true);
          for (int i = 0; i < N; i++) {
            const double ci = SW(S2_OFF_C + i);
            if (i < NAQ) st.pri_molal[i * ld + cell] = ci;
            else { st.immobile[(i - NAQ) * ld + cell] = ci; gimm[i >= NAQ ? i - NAQ : 0] = ci; }
          }
          bool upd = false;
          if (SPEC_NKIN > 0) {
            upd = true;
            for (int m = 0; m < SPEC_NKIN; m++) {
              double vf = st.mnrl_volfrac[m * ld + cell] + st.mnrl_rate[m * ld + cell] * spec_mn_vol(m) * dt;
              if (vf < 0.0) vf = 0.0;
              st.mnrl_volfrac[m * ld + cell] = vf;
            }
          }
          cumulative += dt;
          nss++;
          nconst++;
          if (upd) nku++;
          if (nconst >= 4) {
            ncuts--;
            dt = fmin(2.0 * dt, tran_dt - cumulative);
          }
        }
      }
      if (aborted) {
        for (int i = 0; i < NAQ; i++) st.pri_molal[i * ld + cell] = SW(S2_OFF_C + i);
      } else {
        for (int i = 0; i < N; i++)
          if ((small_mask >> i) & 1u) {
            double *p = (i < NAQ) ? st.total + i * ld + cell : st.immobile + (i - NAQ) * ld + cell;
            *p = small_val[i];
          }
      }
      spec2_store_sec(W, s, st, cell);
      if (SPEC_ACT_UPD) spec2_store_act(s, st, cell);
      if (st.ln_act_h2o && SPEC_USE_ACT_H2O) st.ln_act_h2o[cell] = (s.aw == 1.0) ? 0.0 : log(s.aw);
    }
    st.num_sub_steps[cell] = nss;
    st.num_iterations[cell] = nit;
    st.num_kinetic_state_updates[cell] = nku;
    st.ierror[cell] = aborted ? 1 : 0;
  }
  return 0;
}

// one evaluation at the state's own iterate (c = rt_auxvar%pri_molal / immobile): residual without the
// fixed accumulation subtracted is not available here, so the caller compares Jt only.  SPEC_SYM 0
// builds only (the full matrix in the slice): Jt[i + j * N] for the coupled species, 0 elsewhere.
extern "C" int s2_host_eval(const pfrx_state *ps, long long cell, double dt, double min_sat, double *res_out,
                            double *jt_out) {
#if SPEC_SYM
  (void)ps; (void)cell; (void)dt; (void)min_sat; (void)res_out; (void)jt_out;
  return 1;
#else
  DevState st;
  st.ld = ps->ld;
  st.total = ps->total; st.pri_molal = ps->pri_molal; st.immobile = ps->immobile;
  st.pri_act_coef = ps->pri_act_coef; st.sec_act_coef = ps->sec_act_coef; st.sec_molal = ps->sec_molal;
  st.ln_act_h2o = ps->ln_act_h2o; st.mnrl_volfrac = ps->mnrl_volfrac; st.mnrl_area = ps->mnrl_area;
  st.mnrl_rate = ps->mnrl_rate; st.free_site = ps->srfcplxrxn_free_site_conc; st.eqsrfcplx_conc = ps->eqsrfcplx_conc;
  st.total_sorb_eq = ps->total_sorb_eq; st.kinmr = ps->kinmr_total_sorb; st.den_kg = ps->den_kg; st.sat = ps->sat;
  st.temp = ps->temp; st.porosity = ps->porosity; st.volume = ps->volume;
  st.soil_particle_density = ps->soil_particle_density; st.imat = ps->imat;
  constexpr int N = SPEC_N, NAQ = SPEC_NAQ;
  const long long ld = st.ld;
  std::vector<double> slice((size_t)S2_SLOTS * 32, 0.0);
  double *W = slice.data();
  Spec2Cell s;
  const double den_kg = st.den_kg[cell], sat = st.sat[cell], por = st.porosity[cell];
  s.vol = st.volume[cell];
  s.temp = st.temp[cell];
  const double spd = st.soil_particle_density ? st.soil_particle_density[cell] : 0.0;
  s.aw = 1.0;
  s.denL = den_kg * 1.e-3;
  s.dry = sat < min_sat;
  s.psv = por * sat * 1000.0 * s.vol;
  s.rock = spd * (1.0 - por);
  s.Iact = 0.0;
  s.store = false;
  s.rates = false;
  s.rdt = 1.0 / dt;
  spec2_isec(s, st, cell);
  if (!SPEC_ACT_UPD) spec2_load_frozen(W, st, cell);
  for (int i = 0; i < N; i++) SW(S2_OFF_C + i) = (i < NAQ) ? st.pri_molal[i * ld + cell] : st.immobile[(i - NAQ) * ld + cell];
  spec2_begin(W, s, st, cell);
  double res[N], tv[S2_NTV], ev[S2_NEV];
  spec2_eval(res, tv, ev, s, W, st, cell);
  for (int i = 0; i < N; i++) res_out[i] = res[i];
  for (int i = 0; i < N * N; i++) jt_out[i] = 0.0;
  for (int ci = 0; ci < SPEC_NC; ci++)
    for (int cj = 0; cj < SPEC_NC; cj++) jt_out[spec_sp_of(ci) + spec_sp_of(cj) * N] = W[JX(ci, cj)];
  return 0;
#endif
}
