// pfrx_tpc.cuh -- THREAD-PER-CELL kernel of the operator-split chemistry step.
//
// The cell-group kernel (pfrx_device.cuh) spends ~4 300 warp-instructions per
// cell-iteration on the 15/88 Hanford network: every table-driven FMA drags
// ~10 integer/address/sync instructions along and a warp serves only two
// cells.  Here one THREAD owns one cell, a warp serves 32 cells per
// instruction, and nothing has to be communicated between lanes:
//  * all per-cell arrays whose index comes from a stoichiometry table (c,
//    ln a, 1/c, totals, residual, the n x n Jacobian) live in a per-thread
//    slice of shared memory (odd stride => conflict free); arrays that are only
//    walked with compile-time indices (fixed accumulation, ln gamma, the
//    small-value bookkeeping of RStep) stay in registers;
//  * the Jacobian is assembled the way RTotalAqueous does it -- complex by
//    complex, nu_i * (nu_j m_k / c_j) scattered into d(total)/d(free) -- so the
//    operation order is the reference's; secondary molalities are not kept
//    on chip: they stream to HBM as they are computed and only their ionic
//    strength / molality sums are carried to the next activity update;
//  * LU (Crout order, implicit-scaled partial pivoting, utility.F90:597-735)
//    runs in shared memory on logical row offsets (no physical row swaps);
//  * a warp loads 32 consecutive cells (coalesced SoA), every lane runs RStep on
//    its own cell -- SIMT keeps lanes that need more Newton iterations or
//    sub-steps in the same loop -- and the warp stores and moves on.
// Occupancy is deliberately low (the Jacobians of 64-96 cells fill the shared
// memory of an SM); throughput comes from instruction efficiency and ILP.
#pragma once
#include "pfrx_device.cuh"
#include "pfrx_sandbox.cuh"

template <int N>
struct CellT {
  const DevCfg &cfg;
  const DevState &st;
  double *ws;
  int64_t cell;
  // per-cell scalars
  double den_kg, sat, temp, por, vol, spd, ln_act_h2o;
  double elm_w, elm_o, elm_t, elm_zsoil, elm_kscalar, elm_bd_dry, elm_bsw, elm_plantndemand;  // ELM scalars
  double elm_sucsat, elm_watfc, elm_effpor;  // GetMoistureResponse (flow-coupled ELM build)
  double Isum, msum;  // sum z^2 m and sum m over the secondary species of the latest RTotal
  bool dry;
  // register-resident per-component arrays (compile-time indices only)
  double fixed[N], lngam[N], small_val[N], guess[N];
  unsigned small_mask;

  __device__ CellT(const DevCfg &c, const DevState &s, double *w) : cfg(c), st(s), ws(w) {}

  __device__ __forceinline__ double &C(int i) { return ws[cfg.off_c + i]; }
  __device__ __forceinline__ double &LNA(int i) { return ws[cfg.off_lnact + i]; }
  __device__ __forceinline__ double &INVC(int i) { return ws[cfg.off_invc + i]; }
  __device__ __forceinline__ double &RES(int i) { return ws[cfg.off_res + i]; }
  __device__ __forceinline__ double &TOT(int i) { return ws[cfg.off_acc + i]; }
  __device__ __forceinline__ double &TMP(int i) { return ws[cfg.off_tmp + i]; }
  __device__ __forceinline__ double &TS(int i) { return ws[cfg.off_ts + i]; }
  __device__ __forceinline__ double &J(int i, int j) { return ws[cfg.off_J + i * cfg.js + j]; }
  // views used by the ELM-CN sandboxes (pfrx_sandbox.cuh)
  __device__ __forceinline__ double Cc(int i) const { return ws[cfg.off_c + i]; }
  __device__ __forceinline__ double TOTc(int i) const { return ws[cfg.off_acc + i]; }
  __device__ __forceinline__ double LNAc(int i) const { return ws[cfg.off_lnact + i]; }
  __device__ __forceinline__ double &DT(int i, int j) { return ws[cfg.off_dt + i * cfg.naq + j]; }
  __device__ __forceinline__ double &DS(int i, int j) { return ws[cfg.off_ds + i * cfg.naq + j]; }
  __device__ __forceinline__ double &NC(int k) { return ws[cfg.off_nc + k]; }
  __device__ __forceinline__ double &TG(int i) { return ws[cfg.off_tg + i]; }
  __device__ __forceinline__ double &DG(int i, int j) { return ws[cfg.off_dg + i * cfg.naq + j]; }
  __device__ __forceinline__ double sat_gas() const { return st.sat_gas ? st.sat_gas[cell] : 0.0; }
  __device__ __forceinline__ double gs_logK(int g) const {
    return (cfg.use_isothermal || !cfg.gs_logKcoef) ? cfg.gs_logK[g] : interp_logK(cfg.gs_logKcoef + 5 * g, temp);
  }

  __device__ __forceinline__ double cx_logK(int k) const {
    return cfg.use_isothermal ? cfg.cx_logK[k] : interp_logK(cfg.cx_logKcoef + 5 * k, temp);
  }
  __device__ __forceinline__ double mn_logK(int m) const {
    return (cfg.use_isothermal || !cfg.mn_logKcoef) ? cfg.mn_logK[m] : interp_logK(cfg.mn_logKcoef + 5 * m, temp);
  }
  __device__ __forceinline__ double sc_logK(int k) const {
    return (cfg.use_isothermal || !cfg.sc_logKcoef) ? cfg.sc_logK[k] : interp_logK(cfg.sc_logKcoef + 5 * k, temp);
  }

  // ---- RActivityCoefficients, LAG branch (reaction.F90:4553-4612) -----------
  // ionic strength from the current c and the secondary sums of the last RTotal
  __device__ __forceinline__ void activity() {
    const int naq = cfg.naq;
    double part = 0.0, mp = 0.0;
#pragma unroll 1
    for (int i = 0; i < naq; i++) {
      double c = C(i);
      part += c * cfg.pri_Z2[i];
      if (i != cfg.h2o_aq_id) mp += c;
    }
    double I = 0.5 * (part + Isum);
    double sq = sqrt(I);
#pragma unroll 1
    for (int q = 0; q < cfg.ncls; q++)
      ws[cfg.off_cls + q] = (cfg.cls_negz2[q] * sq * cfg.debyeA / (1.0 + cfg.cls_a0[q] * cfg.debyeB * sq) +
                             cfg.debyeBdot * I) * PFRX_LOG_TO_LN;
    if (cfg.use_act_h2o) {
      double t = 1.0 - 0.017 * (mp + msum);
      ln_act_h2o = t > 0.0 ? log(t) : 0.0;
    }
#pragma unroll
    for (int i = 0; i < N; i++)
      if (i < naq) {
        int q = cfg.pri_cls[i];
        lngam[i] = q < 0 ? 0.0 : ws[cfg.off_cls + q];
      }
  }

  // ---- RActivityCoefficients, NEWTON branch (reaction.F90:4403-4551): the ionic strength is
  // iterated (<= 50 times) together with the complex concentrations it depends on.  As in
  // the reference, ln a_j of the primary species keeps the coefficients the routine was
  // entered with.  Complex concentrations live in rt_auxvar%sec_molal (HBM / L2).
  // Returns false where the reference poisons the state with NaN (no convergence, I < 0).
  __device__ __forceinline__ bool activity_newton() {
    const int naq = cfg.naq, ncx = cfg.ncplx;
    const int64_t ld = st.ld;
    double fpri = 0.0, mp = 0.0;
#pragma unroll
    for (int i = 0; i < N; i++)
      if (i < naq) {
        const double c = C(i);
        LNA(i) = log(c) + lngam[i];
        fpri = fpri + c * cfg.pri_Z2[i];
        if (i != cfg.h2o_aq_id) mp += c;
      }
    int it = 0;
    double II = 0.0;
    for (;;) {
      it++;
      if (it > 50) return false;
      double I = fpri;
#pragma unroll 1
      for (int k = 0; k < ncx; k++) I = I + st.sec_molal[k * ld + cell] * cfg.cx_Z2[k];
      I = 0.5 * I;
      const double f = I;
      if (fabs(I - II) < 1.e-6 * I) break;
      if (ncx > 0) {
        double didi = 0.0;
        const double sq = sqrt(I);
#pragma unroll 1
        for (int k = 0; k < ncx; k++) {
          const int qk = cfg.cx_cls[k];
          if (qk < 0) continue;
          const double t = 1.0 + cfg.debyeB * cfg.cls_a0[qk] * sq;
          double sum = 0.5 * cfg.debyeA * cfg.cx_Z2[k] / (sq * (t * t)) - cfg.debyeBdot;
#pragma unroll 1
          for (int p = cfg.cx_ptr[k]; p < cfg.cx_ptr[k + 1]; p++) {
            const int qj = cfg.pri_cls[cfg.cx_id[p]];
            if (qj >= 0) {
              const double tj = 1.0 + cfg.debyeB * cfg.cls_a0[qj] * sq;
              const double dgamdi = -0.5 * cfg.debyeA * (-cfg.cls_negz2[qj]) / (sq * (tj * tj)) + cfg.debyeBdot;
              sum = sum + cfg.cx_st[p] * dgamdi;
            }
          }
          const double dcdi = st.sec_molal[k * ld + cell] * PFRX_LOG_TO_LN * sum;
          didi = didi + 0.5 * cfg.cx_Z2[k] * dcdi;
        }
        const double den = 1.0 - didi;
        II = fabs(den) > 0.0 ? (f - I * didi) / den : f;
      } else {
        II = f;
      }
      if (II < 0.0) return false;
      I = II;
      const double sq = sqrt(I);
#pragma unroll 1
      for (int q = 0; q < cfg.ncls; q++)
        ws[cfg.off_cls + q] = (cfg.cls_negz2[q] * sq * cfg.debyeA / (1.0 + cfg.cls_a0[q] * cfg.debyeB * sq) +
                               cfg.debyeBdot * I) * PFRX_LOG_TO_LN;
#pragma unroll
      for (int i = 0; i < N; i++)
        if (i < naq) {
          const int q = cfg.pri_cls[i];
          lngam[i] = q < 0 ? 0.0 : ws[cfg.off_cls + q];
        }
      double ms = 0.0, Is = 0.0;
#pragma unroll 1
      for (int k = 0; k < ncx; k++) {
        const int q = cfg.cx_cls[k];
        const double g = q < 0 ? 1.0 : exp(ws[cfg.off_cls + q]);
        const double sk = exp(cx_lnQK(k)) / g;
        st.sec_molal[k * ld + cell] = sk;
        ms += sk;
        Is += sk * cfg.cx_Z2[k];
      }
      Isum = Is;
      msum = ms;
      if (cfg.use_act_h2o) {
        const double t = 1.0 - 0.017 * (mp + ms);
        ln_act_h2o = t > 0.0 ? log(t) : 0.0;
      }
    }
    return true;
  }

  // ---- RTotalSorbEqSurfCplx1 (reaction_surf_complex.F90:641-900) -------------
  // adds nu*S to ws.ts and (add_J) jscale * dtotal_sorb to the Jacobian
  __device__ __forceinline__ void surf_cplx1(int irxn, double *tsacc, bool add_J, double jscale, bool store_conc,
                                             bool add_ds = false) {
    const int naq = cfg.naq;
    const int r0 = cfg.sr_ptr[irxn], r1 = cfg.sr_ptr[irxn + 1];
    double fs = fmax(ws[cfg.off_fs + irxn], 1.e-40);
    double dens;
    int ty = cfg.sr_type[irxn];
    if (ty == PFRX_MINERAL_SURFACE)
      dens = cfg.sr_dens[irxn] * st.mnrl_volfrac[cfg.sr_surf[irxn] * st.ld + cell];
    else if (ty == PFRX_ROCK_SURFACE)
      dens = cfg.sr_dens[irxn] * spd * (1.0 - por);
    else
      dens = cfg.sr_dens[irxn];
    if (dens < 1.e-40) {
      ws[cfg.off_fs + irxn] = 0.0;
      if (store_conc)
        for (int q = r0; q < r1; q++) ws[cfg.off_sc + cfg.sr_cx[q]] = 0.0;
      return;
    }
    double *base = ws + cfg.off_tmp + N;  // lnQK without the free-site term, then S_q
#pragma unroll 1
    for (int q = r0; q < r1; q++) {
      int k = cfg.sr_cx[q];
      double lnQK = -sc_logK(k) * PFRX_LOG_TO_LN;
      if (cfg.sc_h2o[k] != 0.0) lnQK += cfg.sc_h2o[k] * ln_act_h2o;
#pragma unroll 1
      for (int p = cfg.sc_ptr[k]; p < cfg.sc_ptr[k + 1]; p++) lnQK += cfg.sc_st[p] * LNA(cfg.sc_id[p]);
      base[q - r0] = lnQK;
    }
    if (!cfg.sr_flag[irxn]) {
      // unit free-site stoichiometry: Sx = rho / (1 + sum_q exp(lnQK_q)), S_q = exp(lnQK_q) Sx
      double e = 0.0;
#pragma unroll 1
      for (int q = r0; q < r1; q++) {
        double v = exp(base[q - r0]);
        base[q - r0] = v;
        e += v;
      }
      fs = dens / (1.0 + e);
#pragma unroll 1
      for (int q = r0; q < r1; q++) base[q - r0] *= fs;
    } else {
      double *sconc = base + (r1 - r0);
      bool one_more = false;
      int it = 0;
      double damping = 1.0;
      for (;;) {
        it++;
        double total = fs, lnfs = log(fs);
#pragma unroll 1
        for (int q = r0; q < r1; q++) {
          int k = cfg.sr_cx[q];
          double sck = exp(base[q - r0] + cfg.sc_fs[k] * lnfs);
          sconc[q - r0] = sck;
          total += cfg.sc_fs[k] * sck;
        }
        if (one_more) break;
        double res = dens - total, d = 1.0;
#pragma unroll 1
        for (int q = r0; q < r1; q++) d += cfg.sc_fs[cfg.sr_cx[q]] * sconc[q - r0] / fs;
        double dfs = res / d;
        if (it > 1000) damping = 0.5;
        fs = fs + damping * dfs;
        if (fabs(dfs / fs) < 1.e-12 || it > 100000) one_more = true;
      }
#pragma unroll 1
      for (int q = r0; q < r1; q++) base[q - r0] = sconc[q - r0];
    }
    ws[cfg.off_fs + irxn] = fs;
    double denom = 0.0;
#pragma unroll 1
    for (int q = r0; q < r1; q++) {
      int k = cfg.sr_cx[q];
      denom += cfg.sc_fs[k] * cfg.sc_fs[k] * base[q - r0];
    }
    denom = denom / fs + 1.0;
#pragma unroll 1
    for (int i = 0; i < naq; i++) TMP(i) = 0.0;
#pragma unroll 1
    for (int q = r0; q < r1; q++) {
      int k = cfg.sr_cx[q];
      double Sk = base[q - r0], fk = cfg.sc_fs[k];
      if (store_conc) ws[cfg.off_sc + k] += Sk;
#pragma unroll 1
      for (int p = cfg.sc_ptr[k]; p < cfg.sc_ptr[k + 1]; p++) {
        int i = cfg.sc_id[p];
        double nu = cfg.sc_st[p];
        TMP(i) += nu * fk * Sk;
        tsacc[i] += nu * Sk;
      }
    }
    if (!add_J) return;
#pragma unroll 1
    for (int i = 0; i < naq; i++) TMP(i) = (-TMP(i) / denom) * INVC(i);
#pragma unroll 1
    for (int q = r0; q < r1; q++) {
      int k = cfg.sr_cx[q];
      double Sk = base[q - r0];
      double nuiSx = cfg.sc_fs[k] * Sk / fs;
      const int p0 = cfg.sc_ptr[k], p1 = cfg.sc_ptr[k + 1];
#pragma unroll 1
      for (int p2 = p0; p2 < p1; p2++) {
        int j = cfg.sc_id[p2];
        const double t0 = cfg.sc_st[p2] * Sk * INVC(j) + nuiSx * TMP(j);
        double t = jscale * t0;
#pragma unroll 1
        for (int p = p0; p < p1; p++) J(cfg.sc_id[p], j) += cfg.sc_st[p] * t;
        if (add_ds) {
#pragma unroll 1
          for (int p = p0; p < p1; p++) DS(cfg.sc_id[p], j) += cfg.sc_st[p] * t0;
        }
      }
    }
  }

  // ---- RTAuxVarCompute = RTotal (reaction.F90:4618-4759) + accumulation terms --
  // want_J: also d(accumulation)/dc/dt into the Jacobian (reaction.F90:5775-5848)
  // aq_only (ReactionEquilibrateConstraint): the Jacobian is rt_auxvar%aqueous%dtotal itself, the
  // complexes take the class coefficients of the latest activity(), no sorbed totals
  // cls: the complexes take the class coefficients of the latest activity() whatever the update frequency
  __device__ __forceinline__ void auxvar_compute(bool want_J, double dt, bool aq_only = false, bool cls = false) {
    const int naq = cfg.naq, n = cfg.n, ncx = cfg.ncplx;
    const double denL = den_kg * 1.e-3;
    const double f = aq_only ? denL : denL * (por * sat * 1000.0 * vol / dt);  // dtotal -> Jacobian
    const bool act_upd = aq_only || cls || cfg.act_freq == PFRX_ACT_COEF_FREQUENCY_NEWTON_ITER;
#pragma unroll
    for (int i = 0; i < N; i++)
      if (i < naq) {
        double c = C(i);
        LNA(i) = log(c) + lngam[i];
        INVC(i) = 1.0 / c;
        TOT(i) = c;
      }
    if (want_J) {
#pragma unroll 1
      for (int e = 0; e < n * cfg.js; e++) ws[cfg.off_J + e] = 0.0;
    }
    double Is = 0.0, ms = 0.0;
    const int64_t ld = st.ld;
#pragma unroll 2
    for (int k = 0; k < ncx; k++) {
      const int p0 = cfg.cx_ptr[k], p1 = cfg.cx_ptr[k + 1];
      double lnQK = -cx_logK(k) * PFRX_LOG_TO_LN;
      double h = cfg.cx_h2o[k];
      if (h != 0.0) lnQK += h * ln_act_h2o;
#pragma unroll 1
      for (int p = p0; p < p1; p++) lnQK += cfg.cx_st[p] * LNA(cfg.cx_id[p]);
      double lg;
      if (act_upd) {
        int q = cfg.cx_cls[k];
        lg = q < 0 ? 0.0 : ws[cfg.off_cls + q];
      } else {
        lg = ws[cfg.off_lng + k];
      }
      double sk = exp(lnQK - lg);
      st.sec_molal[k * ld + cell] = sk;  // streams to HBM; never re-read in this step
      Is += sk * cfg.cx_Z2[k];
      ms += sk;
#pragma unroll 1
      for (int p = p0; p < p1; p++) {
        int i = cfg.cx_id[p];
        double nu_i = cfg.cx_st[p];
        TOT(i) += nu_i * sk;
      }
      if (want_J && !dry) {
#pragma unroll 1
        for (int p2 = p0; p2 < p1; p2++) {
          int j = cfg.cx_id[p2];
          double t = (cfg.cx_st[p2] * sk) * INVC(j);  // nu_j exp(lnQK - ln c_j)/gamma
#pragma unroll 1
          for (int p = p0; p < p1; p++) J(cfg.cx_id[p], j) += cfg.cx_st[p] * t;
        }
      }
    }
    Isum = Is;
    msum = ms;
#pragma unroll 1
    for (int i = 0; i < naq; i++) TOT(i) *= denL;
    if (want_J) {
      if (dry) {
#pragma unroll 1
        for (int i = 0; i < n; i++) J(i, i) = 1.0;
      } else {
#pragma unroll 1
        for (int i = 0; i < naq; i++) {
          J(i, i) += 1.0;
          if (cfg.need_dt && !aq_only) {  // rt_auxvar%aqueous%dtotal for the sandboxes (reaction.F90:4757)
#pragma unroll 1
            for (int j = 0; j < naq; j++) DT(i, j) = J(i, j) * denL;
          }
#pragma unroll 1
          for (int j = 0; j < naq; j++) J(i, j) *= f;
        }
#pragma unroll 1
        for (int i = naq; i < n; i++) J(i, i) = vol / dt;
      }
    }
#pragma unroll 1
    for (int i = naq; i < n; i++) TOT(i) = C(i);
    if (cfg.nsorb > 0 && !aq_only) total_sorb(want_J, vol / dt);
    if (cfg.ngas > 0 && !aq_only) total_gas(want_J, want_J, dt);
  }

  // ---- RTotalGas (reaction_gas.F90:87-174): partial pressures of the active gas species in
  // equilibrium with the water, their ideal-gas concentrations as rt_auxvar%total(:,2) and its
  // derivative; the gas share of RTAccumulationDerivative (reaction.F90:5838-5846) goes into the Jacobian
  __device__ __forceinline__ double gas_concentration(double pp) const {  // reaction_gas.F90:304-323, mol/m^3
    return pp * 1.e5 / (PFRX_IDEAL_GAS_CONSTANT * (temp + 273.15));
  }
  __device__ __forceinline__ void total_gas(bool want_J, bool add_J, double dt) {
    const int naq = cfg.naq;
#pragma unroll 1
    for (int i = 0; i < naq; i++) TG(i) = 0.0;
    if (want_J) {
#pragma unroll 1
      for (int e = 0; e < naq * naq; e++) ws[cfg.off_dg + e] = 0.0;
    }
#pragma unroll 1
    for (int g = 0; g < cfg.ngas; g++) {
      const int p0 = cfg.gs_ptr[g], p1 = cfg.gs_ptr[g + 1];
      double lnQK = -gs_logK(g) * PFRX_LOG_TO_LN;
      const double h = cfg.gs_h2o[g];
      if (h != 0.0) lnQK = lnQK + h * ln_act_h2o;
#pragma unroll 1
      for (int p = p0; p < p1; p++) lnQK = lnQK + cfg.gs_st[p] * LNA(cfg.gs_id[p]);
      const double pp = exp(lnQK);
      if (st.gas_pp) st.gas_pp[g * st.ld + cell] = pp;
      const double gc = gas_concentration(pp) * 1.e-3;
#pragma unroll 1
      for (int p = p0; p < p1; p++) TG(cfg.gs_id[p]) = TG(cfg.gs_id[p]) + cfg.gs_st[p] * gc;
      if (want_J) {
#pragma unroll 1
        for (int q = p0; q < p1; q++) {
          const int jc = cfg.gs_id[q];
          const double t = cfg.gs_st[q] * gas_concentration(exp(lnQK - log(C(jc)))) * 1.e-3;
#pragma unroll 1
          for (int p = p0; p < p1; p++) DG(cfg.gs_id[p], jc) = DG(cfg.gs_id[p], jc) + cfg.gs_st[p] * t;
        }
      }
    }
    if (add_J && !dry) {
      const double fg = por * sat_gas() * 1000.0 * vol / dt;
#pragma unroll 1
      for (int i = 0; i < naq; i++)
#pragma unroll 1
        for (int j = 0; j < naq; j++) J(i, j) = J(i, j) + DG(i, j) * fg;
    }
  }

  // ---- RTotalSorb (reaction.F90:4783-4835): surface complexation, ion exchange, dynamic KD, KD;
  // jscale * d(total_sorb)/d(free) goes into the Jacobian (RAccumulationSorbDerivative), and into DS
  // unscaled when the decay of a sorbing species needs it
  __device__ __forceinline__ void total_sorb(bool want_J, double jscale) {
    const int naq = cfg.naq;
#pragma unroll 1
    for (int k = 0; k < cfg.nsrfcplx; k++) ws[cfg.off_sc + k] = 0.0;
#pragma unroll 1
    for (int i = 0; i < naq; i++) TS(i) = 0.0;
    if (want_J && cfg.need_ds) {
#pragma unroll 1
      for (int e = 0; e < naq * naq; e++) ws[cfg.off_ds + e] = 0.0;
    }
#pragma unroll 1
    for (int e = 0; e < cfg.neqsr; e++)
      surf_cplx1(cfg.eqsr[e], ws + cfg.off_ts, want_J, jscale, true, cfg.need_ds != 0);
    if (cfg.nionx > 0) ion_exchange(want_J, jscale);
    if (cfg.ndynkd > 0) dynamic_kd(want_J, jscale);
    if (cfg.nkd > 0) isotherm_kd(want_J, jscale);
  }

  // ---- RTotalSorbEqIonx (reaction.F90:4906-5140): Gaines-Thomas ion exchange; with mixed
  // valences the equivalent fraction of the reference cation comes from a scalar Newton on KDj
  // (tol 1e-12) that starts from the cell's previous answer.  d(total_sorb)/d(free) goes
  // straight into the Jacobian with the V/dt of RAccumulationSorbDerivative.
  __device__ __forceinline__ void ion_exchange(bool want_J, double jscale) {
    const int nix = cfg.nionx;
    double *ref = ws + cfg.off_ix;
    double *conc = ws + cfg.off_ix + nix;
#pragma unroll 1
    for (int r = 0; r < nix; r++) {
      const int p0 = cfg.ix_ptr[r], nc = cfg.ix_ptr[r + 1] - p0;
      double omega;
      if (cfg.ix_surf[r] >= 0)
        omega = fmax(cfg.ix_cec[r] * st.mnrl_volfrac[cfg.ix_surf[r] * st.ld + cell], 1.e-40);
      else
        omega = cfg.ix_cec[r];
      double *X = ws + cfg.off_tmp;  // equivalent fractions of this reaction's cations
      if (cfg.ix_zflag[r]) {
        int ic = cfg.ix_cat[p0];
        const double rconc = exp(LNA(ic));  // molality * activity coefficient
        const double rZ = cfg.pri_Z[ic], rk = cfg.ix_k[p0];
        double rX = rZ * ref[r] / omega;
        double KDj = rX / (rk * rconc);
        bool one_more = false;
        int it = 0;
        for (;;) {
          it++;
          if (it > 20000) break;  // the reference flags an error here; never reached in practice
          rX = KDj * (rk * rconc);
          X[0] = rX;
          double total = rX, dres = 0.0;
#pragma unroll 1
          for (int j = 1; j < nc; j++) {
            ic = cfg.ix_cat[p0 + j];
            const double xj = cfg.ix_k[p0 + j] * exp(LNA(ic)) * pow(KDj, cfg.pri_Z[ic] / rZ);
            X[j] = xj;
            total = total + xj;
            dres = dres + xj / KDj * cfg.pri_Z[ic];
          }
          dres = dres / rZ + (rk * rconc);
          const double res = 1.0 - total;
          if (one_more) break;
          const double dK = res / dres;
          KDj = KDj + dK;
          KDj = fmax(KDj, 1.e-40);
          if (fabs(dK / KDj) < 1.e-12) one_more = true;
        }
        ref[r] = rX * omega / rZ;
      } else {
        double sumkm = 0.0;
#pragma unroll 1
        for (int j = 0; j < nc; j++) {
          const int ic = cfg.ix_cat[p0 + j];
          const double xj = exp(LNA(ic)) * cfg.ix_k[p0 + j];
          X[j] = xj;
          sumkm = sumkm + xj;
        }
#pragma unroll 1
        for (int j = 0; j < nc; j++) X[j] = X[j] / sumkm;
      }
      double sumZX = 0.0;
#pragma unroll 1
      for (int i = 0; i < nc; i++) sumZX = sumZX + cfg.pri_Z[cfg.ix_cat[p0 + i]] * X[i];
#pragma unroll 1
      for (int i = 0; i < nc; i++) {
        const int ic = cfg.ix_cat[p0 + i];
        const double t1 = X[i] * omega / cfg.pri_Z[ic];
        conc[p0 + i] = t1;
        TS(ic) = TS(ic) + t1;
        if (want_J) {
          const double t2 = cfg.pri_Z[ic] / sumZX;
#pragma unroll 1
          for (int j = 0; j < nc; j++) {
            const int jc = cfg.ix_cat[p0 + j];
            double d;
            if (i == j)
              d = t1 * (1.0 - (t2 * X[j])) * INVC(jc);
            else
              d = (-t1) * t2 * X[j] * INVC(jc);
            J(ic, jc) += d * jscale;
            if (cfg.need_ds) DS(ic, jc) += d;
          }
        }
      }
    }
  }

  // ---- RRadioactiveDecay (reaction.F90:5211-5311): aqueous inventory of one parent decays into
  // its daughters; d(total)/d(free) is the DT copy of this iteration.  With equilibrium sorption the
  // sorbed inventory decays too, through total_sorb_eq and the DS copy of d(total_sorb)/d(free).
  __device__ __forceinline__ void radioactive_decay() {
    const int naq = cfg.naq;
    const double L_pore = por * vol * 1.e3;
    const double L_water = L_pore * sat;
    const double L_gas = cfg.ngas > 0 ? L_pore * sat_gas() : 0.0;
#pragma unroll 1
    for (int r = 0; r < cfg.nrd; r++) {
      const int p0 = cfg.rd_ptr[r], p1 = cfg.rd_ptr[r + 1], jc = cfg.rd_fwd[r];
      const double kf = cfg.rd_kf[r];
      double sum = TOT(jc) * L_water;
      if (cfg.ngas > 0) sum = sum + TG(jc) * L_gas;
      if (cfg.nsorb > 0) sum = sum + TS(jc) * vol;
      const double rate = sum * kf;
      const double t = -1.0 * kf;
#pragma unroll 1
      for (int p = p0; p < p1; p++) {
        const int ic = cfg.rd_id[p];
        const double nu = cfg.rd_st[p];
        RES(ic) = RES(ic) - nu * rate;
#pragma unroll 1
        for (int j = 0; j < naq; j++) J(ic, j) = J(ic, j) + t * nu * DT(jc, j) * L_water;
      }
      if (cfg.ngas > 0) {
#pragma unroll 1
        for (int p = p0; p < p1; p++) {
          const int ic = cfg.rd_id[p];
          const double nu = cfg.rd_st[p];
#pragma unroll 1
          for (int j = 0; j < naq; j++) J(ic, j) = J(ic, j) + t * nu * DG(jc, j) * L_gas;
        }
      }
      if (cfg.need_ds) {
#pragma unroll 1
        for (int p = p0; p < p1; p++) {
          const int ic = cfg.rd_id[p];
          const double nu = cfg.rd_st[p];
#pragma unroll 1
          for (int j = 0; j < naq; j++) J(ic, j) = J(ic, j) + t * nu * DS(jc, j) * vol;
        }
      }
    }
  }

  // ---- RGeneral (reaction.F90:5316-5460): forward / backward mass-action rates in activities ----
  __device__ __forceinline__ void general_reactions() {
    const double pdsv = por * den_kg * sat * vol;
#pragma unroll 1
    for (int r = 0; r < cfg.ngen; r++) {
      const double kf = cfg.gn_kf[r], kr = cfg.gn_kr[r];
      const int f0 = cfg.gn_fptr[r], f1 = cfg.gn_fptr[r + 1], b0 = cfg.gn_bptr[r], b1 = cfg.gn_bptr[r + 1];
      const int p0 = cfg.gn_ptr[r], p1 = cfg.gn_ptr[r + 1];
      double lnQkf = 0.0, lnQkr = 0.0, Qkf = 0.0, Qkr = 0.0;
      if (kf > 0.0) {
        lnQkf = log(kf);
#pragma unroll 1
        for (int p = f0; p < f1; p++) lnQkf = lnQkf + cfg.gn_fst[p] * LNA(cfg.gn_fid[p]);
        Qkf = exp(lnQkf);
      }
      if (kr > 0.0) {
        lnQkr = log(kr);
#pragma unroll 1
        for (int p = b0; p < b1; p++) lnQkr = lnQkr + cfg.gn_bst[p] * LNA(cfg.gn_bid[p]);
        Qkr = exp(lnQkr);
      }
#pragma unroll 1
      for (int p = p0; p < p1; p++) RES(cfg.gn_id[p]) = RES(cfg.gn_id[p]) - cfg.gn_st[p] * (Qkf - Qkr) * pdsv;
      if (kf > 0.0) {
#pragma unroll 1
        for (int q = f0; q < f1; q++) {
          const int jc = cfg.gn_fid[q];
          const double t = -1.0 * cfg.gn_fst[q] * exp(lnQkf - log(C(jc))) * pdsv;
#pragma unroll 1
          for (int p = p0; p < p1; p++) J(cfg.gn_id[p], jc) = J(cfg.gn_id[p], jc) + cfg.gn_st[p] * t;
        }
      }
      if (kr > 0.0) {
#pragma unroll 1
        for (int q = b0; q < b1; q++) {
          const int jc = cfg.gn_bid[q];
          const double t = cfg.gn_bst[q] * exp(lnQkr - log(C(jc))) * pdsv;
#pragma unroll 1
          for (int p = p0; p < p1; p++) J(cfg.gn_id[p], jc) = J(cfg.gn_id[p], jc) + cfg.gn_st[p] * t;
        }
      }
    }
  }

  // ---- RMicrobial (reaction_microbial.F90:287-602): rate = k * prod(Monod) * prod(inhibition) *
  // biomass; the biomass column of the Jacobian as written there (:578-583, no L_water / volume) ----
  __device__ __forceinline__ double mb_conc(int i, double &dcdm) {
    dcdm = cfg.mb_units == PFRX_MICROBIAL_MOLALITY ? 1.0
           : cfg.mb_units == PFRX_MICROBIAL_ACTIVITY ? pref_act_coef(i)
                                                     : den_kg * 1.e-3;
    return C(i) * dcdm;
  }
  __device__ __forceinline__ double mb_inhibition(int h, double conc, double dcdm, double &dX) {
    const double PI = 3.14159265359;  // pflotran_constants.F90:92, truncated as there
    const double C1 = cfg.mb_hC[h], C2 = cfg.mb_hC2[h];
    const int type = cfg.mb_htype[h];
    if (type == PFRX_INHIBITION_MONOD) {
      const double den = C1 + conc;
      dX = -1.0 * dcdm * C1 / (den * den);
      return C1 / (C1 + conc);
    }
    if (type == PFRX_INHIBITION_INVERSE_MONOD) {
      const double den = C1 + conc;
      dX = dcdm / den - dcdm * conc / (den * den);
      return conc / (C1 + conc);
    }
    if (type == PFRX_INHIBITION_THRESHOLD) {
      const double t = (conc - fabs(C1)) * C2;
      dX = copysign(1.0, C1) * (C2 * dcdm / (1.0 + t * t)) / PI;
      return 0.5 + copysign(1.0, C1) * atan(t) / PI;
    }
    const double lower = log10(C1) - 0.5 * C2;
    const double z = (log10(conc) - lower) / C2;
    if (z < 0.0) {
      dX = 0.0;
      return 0.0;
    }
    if (z > 1.0) {
      dX = 0.0;
      return 1.0;
    }
    dX = (6.0 * z - 6.0 * (z * z)) / (C2 * conc * 2.30258509299) * dcdm;
    return 3.0 * (z * z) - 2.0 * (z * z * z);
  }
  __device__ __forceinline__ void microbial() {
    const int naq = cfg.naq;
    const double L_water = por * sat * vol * 1.e3;
#pragma unroll 1
    for (int r = 0; r < cfg.nmb; r++) {
      const int p0 = cfg.mb_ptr[r], p1 = cfg.mb_ptr[r + 1];
      const int m0 = cfg.mb_mptr[r], nm = cfg.mb_mptr[r + 1] - m0;
      const int h0 = cfg.mb_hptr[r], nh = cfg.mb_hptr[r + 1] - h0;
      double k_eff = cfg.mb_k[r];
      if (cfg.mb_ea) k_eff = k_eff * exp(cfg.mb_ea[r] / 8.31446 * (1.0 / 298.15 - 1.0 / (temp + 273.15)));
      double monod[PFRX_MAX_MONOD_TERMS], inhib[PFRX_MAX_MONOD_TERMS];
      double monod_terms = 1.0, inhib_terms = 1.0, d;
#pragma unroll 1
      for (int ii = 0; ii < nm; ii++) {
        const double conc = mb_conc(cfg.mb_mid[m0 + ii], d);
        const double cth = cfg.mb_mC[m0 + ii];
        monod[ii] = (conc - cth) / (cfg.mb_mK[m0 + ii] + conc - cth);
        monod_terms = monod_terms * monod[ii];
      }
#pragma unroll 1
      for (int ii = 0; ii < nh; ii++) {
        double dcdm, dX;
        const double conc = mb_conc(cfg.mb_hid[h0 + ii], dcdm);
        inhib[ii] = mb_inhibition(h0 + ii, conc, dcdm, dX);
        inhib_terms = inhib_terms * inhib[ii];
      }
      const int ib = cfg.mb_bio[r];
      int brow = -1;
      double biomass_term = 1.0, yield = 0.0, dbio = 0.0;
      if (ib > 0) {
        const double bc = mb_conc(ib - 1, dbio);
        biomass_term = biomass_term * bc * L_water;
        brow = ib - 1;
        yield = cfg.mb_yield[r];
      } else if (ib < 0) {
        brow = naq + (-ib - 1);
        biomass_term = biomass_term * C(brow) * vol;
        dbio = 1.0;
        yield = cfg.mb_yield[r];
      } else {
        biomass_term = biomass_term * L_water;
      }
      const double rate = k_eff * monod_terms * inhib_terms * biomass_term;
#pragma unroll 1
      for (int p = p0; p < p1; p++) RES(cfg.mb_id[p]) = RES(cfg.mb_id[p]) - cfg.mb_st[p] * rate;
      if (brow >= 0) RES(brow) = RES(brow) - yield * rate;
#pragma unroll 1
      for (int ii = 0; ii < nm; ii++) {
        const int jc = cfg.mb_mid[m0 + ii];
        double dcdm;
        const double conc = mb_conc(jc, dcdm);
        double dR_dX = k_eff * inhib_terms * biomass_term;
        for (int jj = 0; jj < ii; jj++) dR_dX = dR_dX * monod[jj];
        for (int jj = ii + 1; jj < nm; jj++) dR_dX = dR_dX * monod[jj];
        const double cth = cfg.mb_mC[m0 + ii];
        const double den = cfg.mb_mK[m0 + ii] + conc - cth;
        const double dX = dcdm / den - dcdm * (conc - cth) / (den * den);
        const double dR_dc = -1.0 * dR_dX * dX;
#pragma unroll 1
        for (int p = p0; p < p1; p++) J(cfg.mb_id[p], jc) = J(cfg.mb_id[p], jc) + cfg.mb_st[p] * dR_dc;
        if (brow >= 0) J(brow, jc) = J(brow, jc) + yield * dR_dc;
      }
#pragma unroll 1
      for (int ii = 0; ii < nh; ii++) {
        const int jc = cfg.mb_hid[h0 + ii];
        double dcdm, dX;
        const double conc = mb_conc(jc, dcdm);
        double dR_dX = k_eff * monod_terms * biomass_term;
        for (int jj = 0; jj < ii; jj++) dR_dX = dR_dX * inhib[jj];
        for (int jj = ii + 1; jj < nh; jj++) dR_dX = dR_dX * inhib[jj];
        (void)mb_inhibition(h0 + ii, conc, dcdm, dX);
        const double dR_dc = -1.0 * dR_dX * dX;
#pragma unroll 1
        for (int p = p0; p < p1; p++) J(cfg.mb_id[p], jc) = J(cfg.mb_id[p], jc) + cfg.mb_st[p] * dR_dc;
        if (brow >= 0) J(brow, jc) = J(brow, jc) + yield * dR_dc;
      }
      if (brow >= 0) {
        double dRb = k_eff * monod_terms * inhib_terms;
        dRb = -1.0 * dRb * dbio;
#pragma unroll 1
        for (int p = p0; p < p1; p++) J(cfg.mb_id[p], brow) = J(cfg.mb_id[p], brow) + cfg.mb_st[p] * dRb;
        J(brow, brow) = J(brow, brow) + yield * dRb;
      }
    }
  }

  // ---- RImmobileDecay (reaction_immobile.F90:244-296) ------------------------------------------
  __device__ __forceinline__ void immobile_decay() {
    const int naq = cfg.naq;
#pragma unroll 1
    for (int r = 0; r < cfg.nidc; r++) {
      const int id = naq + cfg.idc_id[r];
      const double rc = cfg.idc_k[r] * vol;
      RES(id) = RES(id) + rc * C(id);
      J(id, id) = J(id, id) + rc;
    }
  }

  // ---- RTotalSorbDynamicKD (reaction.F90:4836-4902) ---------------------------------------
  __device__ __forceinline__ void dynamic_kd(bool want_J, double jscale) {
    const double Lw = 250.0;
#pragma unroll 1
    for (int r = 0; r < cfg.ndynkd; r++) {
      const int ikd = cfg.dk_spec[r], iref = cfg.dk_ref[r];
      const double mk = C(ikd), mr = C(iref);
      const double pw = cfg.dk_power[r], lo = cfg.dk_low[r], hml = cfg.dk_high[r] - lo;
      const double t = pow(mr / cfg.dk_refhigh[r], pw);
      const double KD = lo + t * hml;
      const double dKD = pw * t / mr * hml;
      TS(ikd) = TS(ikd) + KD * mk * Lw;
      if (want_J) {
        J(ikd, ikd) += (KD * Lw) * jscale;
        J(ikd, iref) += (dKD * mk * Lw) * jscale;
        if (cfg.need_ds) {
          DS(ikd, ikd) += KD * Lw;
          DS(ikd, iref) += dKD * mk * Lw;
        }
      }
    }
  }

  // ---- RTotalSorbKD (reaction_isotherm.F90:273-359): linear / Langmuir / Freundlich -----------
  __device__ __forceinline__ void isotherm_kd(bool want_J, double jscale) {
#pragma unroll 1
    for (int r = 0; r < cfg.nkd; r++) {
      const int ic = cfg.kd_spec[r];
      const double m = C(ic);
      double kd;
      if (cfg.ikd_units == 1)
        kd = cfg.kd_coeff[r] * den_kg * (1.0 - por) * spd * 1.e-3;
      else
        kd = cfg.kd_coeff[r];
      if (cfg.kd_mnrl[r] >= 0) kd = kd * (st.mnrl_volfrac[cfg.kd_mnrl[r] * st.ld + cell]);
      double res = 0.0, dres = 0.0;
      const int ty = cfg.kd_type[r];
      if (ty == PFRX_SORPTION_LINEAR) {
        res = kd * m;
        dres = kd;
      } else if (ty == PFRX_SORPTION_LANGMUIR) {
        const double t = kd * m;
        res = t * cfg.kd_lb[r] / (1.0 + t);
        dres = res / m - res / (1.0 + t) * t / m;
      } else if (ty == PFRX_SORPTION_FREUNDLICH) {
        const double on = 1.0 / cfg.kd_fn[r];
        res = kd * pow(m, on);
        dres = res / m * on;
      }
      TS(ic) = TS(ic) + res;
      if (want_J) {
        J(ic, ic) += dres * jscale;
        if (cfg.need_ds) DS(ic, ic) += dres;
      }
    }
  }

  // ---- RKineticMineral (reaction_mineral.F90:647-1078), no prefactors ---------
  __device__ __forceinline__ void kinetic_mineral(bool apply) {
#pragma unroll 1
    for (int m = 0; m < cfg.nkin; m++) {
      const int p0 = cfg.mn_ptr[m], p1 = cfg.mn_ptr[m + 1];
      double rate_vol = 0.0, Im = 0.0, dfac = 0.0, sprm = 0.0;
      const int npf = cfg.mn_npref ? cfg.mn_npref[m] : 0;
      double pref[PFRX_MAX_PREFACTORS], lnps[PFRX_MAX_PREFACTORS * PFRX_MAX_PREFACTOR_SPECIES];
      double lnQK = -mn_logK(m) * PFRX_LOG_TO_LN;
      if (cfg.mn_h2o[m] != 0.0) lnQK += cfg.mn_h2o[m] * ln_act_h2o;
#pragma unroll 1
      for (int p = p0; p < p1; p++) lnQK += cfg.mn_st[p] * LNA(cfg.mn_id[p]);
      double QK = exp(lnQK);
      double aff;
      if (cfg.mn_temkin) {
        if (cfg.mn_scale)
          aff = 1.0 - pfrx_pow(QK, 1.0 / (cfg.mn_scale[m] * cfg.mn_temkin[m]));
        else
          aff = 1.0 - pfrx_pow(QK, 1.0 / cfg.mn_temkin[m]);
      } else if (cfg.mn_scale) {
        aff = 1.0 - pfrx_pow(QK, 1.0 / cfg.mn_scale[m]);
      } else {
        aff = 1.0 - QK;
      }
      double sgn = copysign(1.0, aff);
      double volfrac = st.mnrl_volfrac[m * st.ld + cell];
      bool active = (volfrac > 0.0 || sgn < 0.0);
      if (active && cfg.mn_irrev[m] == 1 && sgn < 0.0) active = false;
      if (active && cfg.mn_thresh[m] > 0.0 && sgn < 0.0 && QK < cfg.mn_thresh[m]) active = false;
      if (active) {
        double lim = cfg.mn_limit[m];
        if (lim > 0.0) aff = aff / (1.0 + (1.0 - aff) / lim);
        double spr;  // sum_prefactor_rate
        if (npf > 0) {
          // rate = sum over parallel mechanisms of k_p * prod_s a^alpha / (1 + K a^beta) * Arrhenius
          // (reaction_mineral.F90:838-890)
          spr = 0.0;
#pragma unroll 1
          for (int ip = 0; ip < npf; ip++) {
            const int pp = m * PFRX_MAX_PREFACTORS + ip;
            double lnp = 0.0;
#pragma unroll 1
            for (int is = 0; is < cfg.mn_pref_nspec[pp]; is++) {
              const int q = pp * PFRX_MAX_PREFACTOR_SPECIES + is;
              const double lsa = pref_ln_act(cfg.mn_pref_id[q]);
              const double lnum = cfg.mn_pref_alpha[q] * lsa;
              const double lden = log(1.0 + exp(log(cfg.mn_pref_atten[q]) + cfg.mn_pref_beta[q] * lsa));
              lnp = lnp + lnum;
              lnp = lnp - lden;
              lnps[ip * PFRX_MAX_PREFACTOR_SPECIES + is] = lnum - lden;
            }
            pref[ip] = exp(lnp);
            double arr = 1.0;
            if (cfg.mn_pref_eact[pp] > 0.0)
              arr = exp(cfg.mn_pref_eact[pp] / PFRX_IDEAL_GAS_CONSTANT * (1.0 / (25.0 + 273.15) - 1.0 / (temp + 273.15)));
            spr = spr + pref[ip] * cfg.mn_pref_rate[pp] * arr;
          }
        } else {
          double arr = 1.0;
          if (cfg.mn_eact[m] > 0.0)
            arr = exp(cfg.mn_eact[m] / PFRX_IDEAL_GAS_CONSTANT * (1.0 / (25.0 + 273.15) - 1.0 / (temp + 273.15)));
          spr = cfg.mn_rate[m] * arr;
        }
        sprm = spr;
        double Im_const = -st.mnrl_area[m * st.ld + cell];
        if (cfg.mn_scale) Im_const = Im_const / cfg.mn_scale[m];
        if (cfg.mn_power)
          Im = Im_const * sgn * pfrx_pow(fabs(aff), cfg.mn_power[m]) * spr;
        else
          Im = Im_const * sgn * fabs(aff) * spr;
        rate_vol = Im;
        Im_const = Im_const * vol;
        Im = Im * vol;
        double dIm_dQK;
        if (cfg.mn_power)
          dIm_dQK = -Im * cfg.mn_power[m] / fabs(aff);
        else
          dIm_dQK = -Im_const * spr;
        if (cfg.mn_temkin) {
          if (cfg.mn_scale)
            dIm_dQK = dIm_dQK * (1.0 / (cfg.mn_scale[m] * cfg.mn_temkin[m])) / QK * (1.0 - aff);
          else
            dIm_dQK = dIm_dQK * (1.0 / cfg.mn_temkin[m]) / QK * (1.0 - aff);
        } else if (cfg.mn_scale) {
          dIm_dQK = dIm_dQK * (1.0 / cfg.mn_scale[m]) / QK * (1.0 - aff);
        }
        dfac = dIm_dQK * QK * (den_kg * 1.e-3);
        if (lim > 0.0) {
          double den = 1.0 + (1.0 - aff) / lim;
          dfac = dIm_dQK * (1.0 + QK / lim / den) * QK * (den_kg * 1.e-3) / den;
        }
      }
      ws[cfg.off_mn + m] = rate_vol;
      if (!apply || (Im == 0.0 && dfac == 0.0)) continue;
#pragma unroll 1
      for (int p2 = p0; p2 < p1; p2++) {
        int j = cfg.mn_id[p2];
        double t = dfac * (cfg.mn_st[p2] * INVC(j));
#pragma unroll 1
        for (int p = p0; p < p1; p++) J(cfg.mn_id[p], j) += cfg.mn_st[p] * t;
      }
#pragma unroll 1
      for (int p = p0; p < p1; p++) RES(cfg.mn_id[p]) += cfg.mn_st[p] * Im;
      if (npf > 0 && Im != 0.0) {
        // d Im / d (prefactor species) (reaction_mineral.F90:985-1075)
        const double dIm_dspr = Im / sprm;
#pragma unroll 1
        for (int ip = 0; ip < npf; ip++) {
          const int pp = m * PFRX_MAX_PREFACTORS + ip;
          double arr = 1.0;
          if (cfg.mn_pref_eact[pp] > 0.0)
            arr = exp(cfg.mn_pref_eact[pp] / PFRX_IDEAL_GAS_CONSTANT * (1.0 / (25.0 + 273.15) - 1.0 / (temp + 273.15)));
          const double lnp = log(pref[ip]);
#pragma unroll 1
          for (int is = 0; is < cfg.mn_pref_nspec[pp]; is++) {
            const int q = pp * PFRX_MAX_PREFACTOR_SPECIES + is;
            const double lps = lnps[ip * PFRX_MAX_PREFACTOR_SPECIES + is];
            const double dp_dps = exp(lnp - lps);
            const int id = cfg.mn_pref_id[q];
            const double lsa = pref_ln_act(id);
            const double gam = pref_act_coef(id);
            const double dnum = cfg.mn_pref_alpha[q] * exp(lps - lsa);
            const double lgb = cfg.mn_pref_beta[q] * lsa;
            const double den = 1.0 + exp(log(cfg.mn_pref_atten[q]) + lgb);
            const double dden = -1.0 * exp(lps) / den * cfg.mn_pref_atten[q] * cfg.mn_pref_beta[q] * exp(lgb - lsa);
            double dps = dnum + dden;
            dps = dps * gam;
            const double dIm_dspec = dIm_dspr * dp_dps * dps * cfg.mn_pref_rate[pp] * arr;
            if (id >= 0) {
#pragma unroll 1
              for (int p = p0; p < p1; p++) J(cfg.mn_id[p], id) += cfg.mn_st[p] * dIm_dspec;
            } else {
              // a secondary species: the reference's loop runs over the COMPLEX's species here
              // (reaction_mineral.F90:1055 overwrites ncomp); restated as it executes
              const int k = -id - 1;
              const int q0 = cfg.cx_ptr[k], q1 = cfg.cx_ptr[k + 1];
              const double lq = cx_lnQK(k);
              const double gs = pref_act_coef(id);
#pragma unroll 1
              for (int jj = q0; jj < q1; jj++) {
                const int jc = cfg.cx_id[jj];
                const double tr = cfg.cx_st[jj] * exp(lq - log(C(jc))) / gs;
#pragma unroll 1
                for (int ii = q0; ii < q1; ii++) J(cfg.cx_id[ii], jc) += cfg.cx_st[ii] * tr * dIm_dspec;
              }
            }
          }
        }
      }
    }
  }

  // ln activity / activity coefficient of a prefactor species: id >= 0 primary, else complex -(id+1)
  __device__ __forceinline__ double cx_lnQK(int k) const {
    double lq = -cx_logK(k) * PFRX_LOG_TO_LN;
    if (cfg.cx_h2o[k] != 0.0) lq = lq + cfg.cx_h2o[k] * ln_act_h2o;
#pragma unroll 1
    for (int p = cfg.cx_ptr[k]; p < cfg.cx_ptr[k + 1]; p++) lq = lq + cfg.cx_st[p] * ws[cfg.off_lnact + cfg.cx_id[p]];
    return lq;
  }
  __device__ __forceinline__ double sec_ln_gamma(int k) const {
    if (cfg.act_freq == PFRX_ACT_COEF_FREQUENCY_NEWTON_ITER) {
      const int q = cfg.cx_cls[k];
      return q < 0 ? 0.0 : ws[cfg.off_cls + q];
    }
    return ws[cfg.off_lng + k];
  }
  __device__ __forceinline__ double pref_ln_act(int id) const {
    // ln(m_k gamma_k) of a complex is its lnQK (m_k = exp(lnQK) / gamma_k)
    return id >= 0 ? ws[cfg.off_lnact + id] : cx_lnQK(-id - 1);
  }
  __device__ __forceinline__ double pref_act_coef(int id) const {
    if (id < 0) return exp(sec_ln_gamma(-id - 1));
    double lg = 0.0;
#pragma unroll
    for (int i = 0; i < N; i++)
      if (i == id) lg = lngam[i];
    return exp(lg);
  }

  // ---- RMultiRateSorption (reaction_surf_complex.F90:552-637) -----------------
  __device__ __forceinline__ void multirate(double dt) {
    const int naq = cfg.naq;
#pragma unroll 1
    for (int q = 0; q < cfg.nmr; q++) {
      int irxn = cfg.mr_rxn[q];
      int r0 = cfg.mr_ptr[q], r1 = cfg.mr_ptr[q + 1];
      double A = 0.0;
#pragma unroll 1
      for (int k = r0; k < r1; k++) A += cfg.mr_rate[k] / (1.0 + cfg.mr_rate[k] * dt) * cfg.mr_frac[k];
      double *seq = ws + cfg.off_mr + (2 * q) * N;  // kinmr_total_sorb(:,0,q): the equilibrium target
#pragma unroll 1
      for (int i = 0; i < naq; i++) seq[i] = 0.0;
      surf_cplx1(irxn, seq, true, vol * A, false);
#pragma unroll 1
      for (int i = 0; i < naq; i++) RES(i) += vol * (A * seq[i] - ws[cfg.off_mr + (2 * q + 1) * N + i]);
    }
  }

  __device__ __forceinline__ void multirate_begin(double dt) {
    const int naq = cfg.naq;
#pragma unroll 1
    for (int q = 0; q < cfg.nmr; q++) {
      int r0 = cfg.mr_ptr[q], r1 = cfg.mr_ptr[q + 1];
      int64_t base = (int64_t)naq * (r0 + q);
#pragma unroll 1
      for (int i = 0; i < naq; i++) {
        double B = 0.0;
#pragma unroll 1
        for (int k = r0; k < r1; k++) {
          double kk = cfg.mr_rate[k] / (1.0 + cfg.mr_rate[k] * dt);
          B += kk * st.kinmr[(base + (int64_t)naq * (k - r0 + 1) + i) * st.ld + cell];
        }
        ws[cfg.off_mr + (2 * q + 1) * N + i] = B;
      }
    }
  }

  // ---- CLM_CN_React (reaction_sandbox_clm_cn.F90:468-787) ----------------------
  __device__ __forceinline__ void clm_cn() {
    const int off = cfg.naq;
    double temp_K = temp + 273.15;
    if (!(temp_K > 227.15)) return;
    const double one_over_71_02 = 1.408054069e-2, theta_min = 0.01, one_over_log_theta_min = -2.17147241e-1;
    double F_t = exp(308.56 * (one_over_71_02 - 1.0 / (temp_K - 227.13)));
    double F_theta = log(theta_min / fmax(theta_min, sat)) * one_over_log_theta_min;
    double cinh = F_t * F_theta;
    const int ires_C = off + cfg.cn_C, ispec_N = cfg.cn_N, ires_N = off + ispec_N;
    const double *imm = ws + cfg.off_c + off;
#pragma unroll 1
    for (int x = 0; x < cfg.cn_nrxn; x++) {
      double src = cfg.cn_k[x] * vol * cinh;
      double resp = cfg.cn_resp[x];
      int pu = cfg.cn_up[x];
      bool constCN = (cfg.cn_nspec[pu] == 1);
      int iC = cfg.cn_cid[pu], iN = -1;
      double CNu;
      if (!constCN) {
        iN = cfg.cn_nid[pu];
        CNu = imm[iC] / imm[iN];
      } else {
        CNu = cfg.cn_CN[pu];
      }
      double sUC = 1.0;
      double sUN = sUC / CNu;
      int pd = cfg.cn_down[x];
      int id = -1;
      double sDC = 0.0, CNd = 1.0;
      if (pd >= 0) {
        id = cfg.cn_cid[pd];
        CNd = cfg.cn_CN[pd];
        sDC = (1.0 - resp) * sUC;
      }
      double sC = resp * sUC;
      double sN = sUN - sDC / CNd;
      bool useInh;
      double Ninh, dNinh;
      if (cfg.cn_inhib[x] > 1.e-40 && sN < 0.0) {
        useInh = true;
        double t = imm[ispec_N] + cfg.cn_inhib[x];
        Ninh = imm[ispec_N] / t;
        dNinh = cfg.cn_inhib[x] / (t * t);
      } else {
        useInh = false;
        Ninh = 1.0;
        dNinh = 0.0;
      }
      double rate = imm[iC] * src * Ninh;
      int rUC = off + iC, rUN = off + iN, rD = off + id;
      RES(ires_C) -= sC * rate;
      RES(ires_N) -= sN * rate;
      RES(rUC) -= (-1.0) * sUC * rate;
      if (!constCN) RES(rUN) -= (-1.0) * sUN * rate;
      if (id >= 0) RES(rD) -= sDC * rate;
      double drate = src * Ninh;
      double dInh = 0.0;
      J(rUC, rUC) -= (-1.0) * sUC * drate;
      if (useInh) {
        dInh = imm[iC] * src * dNinh;
        J(rUC, ires_N) -= (-1.0) * sUC * dInh;
      }
      if (id >= 0) {
        J(rD, rUC) -= sDC * drate;
        if (useInh) J(rD, ires_N) -= sDC * dInh;
      }
      if (!constCN) {
        J(rUN, rUC) -= (-1.0) * sUN * drate;
        if (useInh) J(rUN, ires_N) -= (-1.0) * sUN * dInh;
        double nc = imm[iN] / imm[iC] * src * Ninh;
        J(rUN, rUC) -= (-1.0) * (-1.0) * nc;
        J(rUN, rUN) -= (-1.0) * src * Ninh;
        J(ires_N, rUC) -= (-1.0) * nc;
        J(ires_N, rUN) -= src * Ninh;
      }
      J(ires_C, rUC) -= sC * drate;
      J(ires_N, rUC) -= sN * drate;
      if (useInh) {
        J(ires_C, ires_N) -= sC * dInh;
        J(ires_N, ires_N) -= sN * dInh;
      }
    }
  }

  // ---- RSandboxEvaluate (reaction_sandbox.F90:294-330): the deck's order ----------
  __device__ __forceinline__ void sandboxes(double dt) {
#pragma unroll 1
    for (int k = 0; k < cfg.nsbx; k++) {
      const int kind = cfg.sbx[k];
      if (kind == PFRX_SANDBOX_CLM_CN) {
        if (cfg.cn_nrxn > 0) clm_cn();
      } else if (kind == PFRX_SANDBOX_SOMDEC) {
        if (cfg.has_sd) {
          pfrx_sbx::SomDec<CellT<N>> sdx(*this, dt);
          sdx.react();
        }
      } else if (kind == PFRX_SANDBOX_NITRIF) {
        if (cfg.has_nt) pfrx_sbx::nitrif_react(*this);
      } else if (kind == PFRX_SANDBOX_DENITR) {
        if (cfg.has_dn) pfrx_sbx::denitr_react(*this);
      } else if (kind == PFRX_SANDBOX_PLANTN) {
        if (cfg.has_pn) pfrx_sbx::plantn_react(*this, dt);
      } else if (kind == PFRX_SANDBOX_LANGMUIR) {
        if (cfg.has_lg) pfrx_sbx::langmuir_react(*this, dt);
      } else if (kind == PFRX_SANDBOX_CNDEGAS) {
        if (cfg.has_cd) {
          double lgp = 0.0;  // ln gamma of H+ (register array: literal indices only)
#pragma unroll
          for (int i = 0; i < N; i++)
            if (i == cfg.cd.proton_id) lgp = lngam[i];
          pfrx_sbx::cndegas_react(*this, lgp);
        }
      } else if (kind == PFRX_SANDBOX_CALCITE) {
        if (cfg.has_cs) {
          const double aux = pfrx_sbx::calcite_react(*this);
          if (st.sandbox_aux) st.sandbox_aux[cell] = aux;
        }
      } else if (kind == PFRX_SANDBOX_RADON) {
        // RadonEvaluate (reaction_sandbox_radon.F90:150-188): zero-order generation, no derivative
        if (cfg.has_rn)
          RES(cfg.rn.species_id) = RES(cfg.rn.species_id) - (1.0) * cfg.rn.radon_generation_rate *
                                                                st.mnrl_volfrac[cfg.rn.mineral_id * st.ld + cell] * vol;
      }
    }
  }

  // per-cell inputs of the sandboxes; NC() <- persisted ratios or the set-up values
  __device__ __forceinline__ void sandbox_load(int64_t c) {
    if (cfg.elm) {
      elm_w = st.elm_w ? st.elm_w[c] : 1.0;
      elm_o = st.elm_o ? st.elm_o[c] : 1.0;
      elm_t = st.elm_t ? st.elm_t[c] : 1.0;
      elm_zsoil = st.elm_zsoil ? st.elm_zsoil[c] : 0.0;
      elm_kscalar = st.elm_kscalar ? st.elm_kscalar[c] : 1.0;
      elm_bd_dry = st.elm_bd_dry ? st.elm_bd_dry[c] : 1.25e3;
      elm_bsw = st.elm_bsw ? st.elm_bsw[c] : 1.0;
      elm_plantndemand = st.elm_plantndemand ? st.elm_plantndemand[c] : 0.0;
      elm_sucsat = st.elm_sucsat ? st.elm_sucsat[c] : 200.0;
      elm_watfc = st.elm_watfc ? st.elm_watfc[c] : 0.1;
      elm_effpor = st.elm_effpor ? st.elm_effpor[c] : 0.4;
    }
#pragma unroll 1
    for (int k = 0; k < cfg.n_nc; k++) {
      double v;
      if (st.somdec_nc)
        v = st.somdec_nc[k * st.ld + c];
      else
        v = k < cfg.sd.nrxn ? cfg.sd.upstream_nc[k] : cfg.sd.downstream_nc[k - cfg.sd.nrxn];
      NC(k) = v;
    }
  }
  __device__ __forceinline__ void sandbox_store(int64_t c) {
    if (st.somdec_nc) {
#pragma unroll 1
      for (int k = 0; k < cfg.n_nc; k++) st.somdec_nc[k * st.ld + c] = NC(k);
    }
  }

  // rt_auxvar%aqueous%dtotal from the state as it stands, for pfrx_reaction (the
  // GIRT caller has it from its own RTAuxVarCompute, reaction.F90:4665-4759)
  __device__ __forceinline__ void dtotal_from_state() {
    const int naq = cfg.naq, ncx = cfg.ncplx;
    const double denL = den_kg * 1.e-3;
#pragma unroll 1
    for (int e = 0; e < naq * naq; e++) ws[cfg.off_dt + e] = 0.0;
    // the totals too, from the free-ion concentrations, like RTAuxVarCompute before RReaction in
    // the GIRT residual (reactive_transport.F90:2590-2626)
#pragma unroll 1
    for (int i = 0; i < naq; i++) TOT(i) = C(i);
#pragma unroll 1
    for (int k = 0; k < ncx; k++) {
      const int p0 = cfg.cx_ptr[k], p1 = cfg.cx_ptr[k + 1];
      double lnQK = -cx_logK(k) * PFRX_LOG_TO_LN;
      double h = cfg.cx_h2o[k];
      if (h != 0.0) lnQK += h * ln_act_h2o;
#pragma unroll 1
      for (int p = p0; p < p1; p++) lnQK += cfg.cx_st[p] * LNA(cfg.cx_id[p]);
      const double sk = exp(lnQK) / st.sec_act_coef[k * st.ld + cell];
#pragma unroll 1
      for (int p = p0; p < p1; p++) TOT(cfg.cx_id[p]) += cfg.cx_st[p] * sk;
#pragma unroll 1
      for (int p2 = p0; p2 < p1; p2++) {
        const int j = cfg.cx_id[p2];
        const double t = (cfg.cx_st[p2] * sk) * INVC(j);
#pragma unroll 1
        for (int p = p0; p < p1; p++) DT(cfg.cx_id[p], j) += cfg.cx_st[p] * t;
      }
    }
#pragma unroll 1
    for (int i = 0; i < naq; i++) {
      TOT(i) *= denL;
      DT(i, i) += 1.0;
#pragma unroll 1
      for (int j = 0; j < naq; j++) DT(i, j) *= denL;
    }
  }

  // ---- RSolve (reaction.F90:5457-5516) + LU (utility.F90:597-735) in shared ---
  // Jacobian rows are addressed through logical row offsets ro[] (ints kept in
  // ws.x), so a row interchange swaps two offsets.  The update lands in RES.
  __device__ __forceinline__ bool solve() { return solve(cfg.n, cfg.use_log != 0); }
  __device__ __forceinline__ bool solve(const int n, const bool use_log) {
    const int js = cfg.js;
    double *A = ws + cfg.off_J;
    int *ro = reinterpret_cast<int *>(ws + cfg.off_x);  // n row offsets
    double *vv = ws + cfg.off_xs;                        // implicit scaling
    // row scaling, optional d/dlnc scaling, vv (reaction.F90:5485-5498, utility.F90:620-640)
    bool bad = false;
#pragma unroll 1
    for (int i = 0; i < n; i++) {
      double *row = A + i * js;
      double m = 0.0;
#pragma unroll 1
      for (int j = 0; j < n; j++) m = fmax(m, fabs(row[j]));
      double nm = 1.0 / fmax(1.0, m);
      RES(i) *= nm;
      double m2 = 0.0;
#pragma unroll 1
      for (int j = 0; j < n; j++) {
        double v = row[j] * nm;
        if (use_log) v *= C(j);
        row[j] = v;
        m2 = fmax(m2, fabs(v));
      }
      if (!(m2 > 0.0)) bad = true;
      vv[i] = 1. / m2;
      ro[i] = i * js;
    }
    if (bad) return false;
    // Crout's method in the reference's loop order (column by column)
#pragma unroll 1
    for (int j = 0; j < n; j++) {
#pragma unroll 1
      for (int i = 0; i < j; i++) {
        const double *ri = A + ro[i];
        double sum = ri[j];
#pragma unroll 1
        for (int k = 0; k < i; k++) sum -= ri[k] * A[ro[k] + j];
        A[ro[i] + j] = sum;
      }
      double aamax = 0.0;
      int imax = j;
#pragma unroll 1
      for (int i = j; i < n; i++) {
        const double *ri = A + ro[i];
        double sum = ri[j];
#pragma unroll 1
        for (int k = 0; k < j; k++) sum -= ri[k] * A[ro[k] + j];
        A[ro[i] + j] = sum;
        double dum = vv[i] * fabs(sum);
        if (dum >= aamax) {
          imax = i;
          aamax = dum;
        }
      }
      if (j != imax) {
        int t = ro[imax];
        ro[imax] = ro[j];
        ro[j] = t;
        double b = RES(imax);  // the permutation is applied to the rhs right away
        RES(imax) = RES(j);
        RES(j) = b;
        vv[imax] = vv[j];
      }
      double *rj = A + ro[j];
      if (rj[j] == 0.0) rj[j] = 1.0e-20;
      if (j != n - 1) {
        double dum = 1.0 / rj[j];
#pragma unroll 1
        for (int i = j + 1; i < n; i++) A[ro[i] + j] *= dum;
      }
    }
    // forward and back substitution (rhs already permuted)
#pragma unroll 1
    for (int i = 0; i < n; i++) {
      const double *ri = A + ro[i];
      double sum = RES(i);
#pragma unroll 1
      for (int k = 0; k < i; k++) sum -= ri[k] * RES(k);
      RES(i) = sum;
    }
#pragma unroll 1
    for (int i = n - 1; i >= 0; i--) {
      const double *ri = A + ro[i];
      double sum = RES(i);
#pragma unroll 1
      for (int k = i + 1; k < n; k++) sum -= ri[k] * RES(k);
      RES(i) = sum / ri[i];
    }
    return true;
  }

  // rt_auxvar%total(:,2) of the latest RTotal: every way out of RReact leaves it in the state
  __device__ __forceinline__ void store_total_gas() {
    if (cfg.ngas > 0 && st.total_gas) {
#pragma unroll 1
      for (int i = 0; i < cfg.naq; i++) st.total_gas[i * st.ld + cell] = TG(i);
    }
  }

  // ---- RReact (reaction.F90:3742-4055) -------------------------------------------
  // state in: st.total / st.immobile / st.total_sorb_eq hold total*, the guess is
  // in st.pri_molal / st.immobile.  Returns ierror; on success the converged state
  // has been written back (total, pri_molal, immobile, total_sorb_eq).
  __device__ __forceinline__ int react(double dt, int &its_out) {
    const int naq = cfg.naq, n = cfg.n;
    const int64_t ld = st.ld, c = cell;
    const double psv = por * sat * 1000.0 * vol;
    const double psv_g = cfg.ngas > 0 ? por * sat_gas() * 1000.0 * vol : 0.0;
    dry = sat < cfg.min_sat;
#pragma unroll
    for (int i = 0; i < N; i++) {
      double f = 0.0;
      if (i < naq) {
        if (!dry) f = psv * st.total[i * ld + c];
        if (!dry && cfg.ngas > 0 && st.total_gas) f = f + psv_g * st.total_gas[i * ld + c];
        if (cfg.nsorb > 0) f = f + st.total_sorb_eq[i * ld + c] * vol;
        C(i) = guess[i];
      } else if (i < n) {
        if (!dry) f = 0.0 + st.immobile[(i - naq) * ld + c] * vol;
        C(i) = guess[i];
      }
      fixed[i] = f;
    }
    if (cfg.nmr > 0) multirate_begin(dt);
    int its = 0;
    double norm0 = 0.0;
    for (;;) {
      its++;
      bool act_ok = true;
      if (cfg.act_freq == PFRX_ACT_COEF_FREQUENCY_NEWTON_ITER) {
        if (cfg.act_alg == PFRX_ACT_COEF_ALGORITHM_NEWTON)
          act_ok = activity_newton();
        else
          activity();
      }
      auxvar_compute(true, dt);
      if (its > cfg.max_its) {
        // total and immobile keep their initial values (never overwritten in
        // HBM); total_sorb_eq is not restored (reaction.F90:3891-3894)
        if (cfg.nsorb > 0)
          for (int i = 0; i < naq; i++) st.total_sorb_eq[i * ld + c] = TS(i);
        store_total_gas();
        its_out = its;
        return 1;
      }
#pragma unroll
      for (int i = 0; i < N; i++) {
        if (i < n) {
          double a = 0.0;
          if (!dry) a = (i < naq) ? psv * TOT(i) : 0.0 + C(i) * vol;
          if (!dry && cfg.ngas > 0 && i < naq) a = a + psv_g * TG(i);
          if (cfg.nsorb > 0 && i < naq) a = a + TS(i) * vol;
          RES(i) = (a - fixed[i]) / dt;
        }
      }
      if (cfg.nkin > 0) kinetic_mineral(!dry);
      if (!dry) {
        // RReaction's order (reaction.F90:4095-4127)
        if (cfg.nmr > 0) multirate(dt);
        if (cfg.nrd > 0) radioactive_decay();
        if (cfg.ngen > 0) general_reactions();
        if (cfg.nmb > 0) microbial();
        if (cfg.nidc > 0) immobile_decay();
        if (cfg.nsbx > 0) sandboxes(dt);
      }
      if (!act_ok) {
        // the reference has filled the state with NaN by now and leaves RReact with
        // option%ierror set after RReaction (reaction.F90:3921); no restore
        store_total_gas();
        its_out = its;
        return 1;
      }
      double mabs = 0.0, ss = 0.0;
#pragma unroll 1
      for (int i = 0; i < n; i++) {
        double r = RES(i);
        mabs = fmax(mabs, fabs(r));
        ss += r * r;
      }
      double nrm = sqrt(ss);
      if (its == 1) norm0 = nrm;
      double rel = nrm / norm0;
      if (mabs < cfg.tol_res) break;
      if (rel < cfg.tol_relres) break;
      if (!solve()) {
        // solve_error branch: no restore (reaction.F90:3964-3967)
        for (int i = 0; i < naq; i++) {
          st.total[i * ld + c] = TOT(i);
          if (cfg.nsorb > 0) st.total_sorb_eq[i * ld + c] = TS(i);
        }
        for (int i = naq; i < n; i++) st.immobile[(i - naq) * ld + c] = C(i);
        store_total_gas();
        its_out = its;
        return 1;
      }
      // update (reaction.F90:3993-4041); the candidate goes to TMP
      double maxrel = -1.0;
      if (cfg.use_log) {
#pragma unroll 1
        for (int i = 0; i < n; i++) {
          double u = RES(i);
          u = copysign(1.0, u) * fmin(fabs(u), cfg.max_dlnC);
          double cc = C(i), cn = cc * exp(-u);
          TMP(i) = cn;
          double v = fabs((cn - cc) / cc);
          if (!isnan(v)) maxrel = fmax(maxrel, v);
        }
      } else {
        double minr = 1.e20;
#pragma unroll 1
        for (int i = 0; i < n; i++) {
          double u = RES(i), cc = C(i);
          if (cc <= u) minr = fmin(minr, fabs(cc / u));
        }
#pragma unroll 1
        for (int i = 0; i < n; i++) {
          double u = RES(i), cc = C(i);
          if (minr < 1.0) u = u * minr * 0.99;
          double cn = cc - u;
          TMP(i) = cn;
          double v = fabs((cn - cc) / cc);
          if (!isnan(v)) maxrel = fmax(maxrel, v);
        }
      }
      if (maxrel >= 0.0 && maxrel < cfg.tol_relchange) break;
#pragma unroll 1
      for (int i = 0; i < n; i++) C(i) = TMP(i);
    }
    // converged; the reference's "one last update" (reaction.F90:4052) recomputes
    // RTotal at the same c: TOT / TS / sec_molal already hold those values
    for (int i = 0; i < naq; i++) {
      st.total[i * ld + c] = TOT(i);
      if (cfg.nsorb > 0) st.total_sorb_eq[i * ld + c] = TS(i);
    }
    for (int i = naq; i < n; i++) st.immobile[(i - naq) * ld + c] = C(i);
    store_total_gas();
#pragma unroll
    for (int i = 0; i < N; i++)
      if (i < n) guess[i] = C(i);
    its_out = its;
    return 0;
  }

  // ---- RUpdateKineticState (reaction.F90:5935-5972) ------------------------------
  __device__ __forceinline__ bool update_kinetic_state(double dt) {
    const int naq = cfg.naq;
    bool upd = false;
    if (cfg.nkin > 0) {
      upd = true;
      // rates of the converged iterate are in ws.mn (same inputs as the
      // RKineticMineral call of MineralUpdateKineticState)
#pragma unroll 1
      for (int m = 0; m < cfg.nkin; m++) {
        double vf = st.mnrl_volfrac[m * st.ld + cell] + ws[cfg.off_mn + m] * cfg.mn_vol[m] * dt;
        if (vf < 0.0) vf = 0.0;
        st.mnrl_volfrac[m * st.ld + cell] = vf;
      }
    }
#pragma unroll 1
    for (int q = 0; q < cfg.nmr; q++) {
      upd = true;
      int r0 = cfg.mr_ptr[q], r1 = cfg.mr_ptr[q + 1];
      int64_t base = (int64_t)naq * (r0 + q);
#pragma unroll 1
      for (int i = 0; i < naq; i++) {
        double seq = ws[cfg.off_mr + (2 * q) * N + i];
#pragma unroll 1
        for (int k = r0; k < r1; k++) {
          double kdt = cfg.mr_rate[k] * dt;
          int64_t ix = (base + (int64_t)naq * (k - r0 + 1) + i) * st.ld + cell;
          st.kinmr[ix] = (st.kinmr[ix] + kdt * cfg.mr_frac[k] * seq) / (1.0 + kdt);
        }
      }
    }
    if (cfg.has_cs && st.sandbox_aux) {  // CalciteUpdateKineticState (reaction_sandbox_calcite.F90:369-410)
      const int m = cfg.cs.mineral_id;
      double vf = st.mnrl_volfrac[m * st.ld + cell] + st.sandbox_aux[cell] * cfg.mn_vol[m] * dt;
      if (vf < 0.0) vf = 0.0;
      st.mnrl_volfrac[m * st.ld + cell] = vf;
    }
    if (cfg.nsbx > 0) upd = true;  // any sandbox => true (reaction.F90:5965)
    return upd;
  }

  // ---- RReaction + RReactionDerivative (reaction.F90:4059-4130, 4134-4208) for one cell ----
  // The GIRT / ELM caller (reactive_transport.F90:2627, 3288) adds these kinetic terms to
  // its own residual and Jacobian; rt_auxvar is taken as it stands (no RTAuxVarCompute).
  // res[i*ld + c] (mol/s), jac[(i*n + j)*ld + c] = d res_i / d c_j.
  __device__ __forceinline__ void reaction(int64_t c, bool want_jac, double *res, double *jac, double tran_dt) {
    cell = c;
    const int naq = cfg.naq, n = cfg.n;
    const int64_t ld = st.ld;
    den_kg = st.den_kg[c];
    sat = st.sat[c];
    temp = st.temp[c];
    por = st.porosity[c];
    vol = st.volume[c];
    spd = st.soil_particle_density ? st.soil_particle_density[c] : 0.0;
    ln_act_h2o = st.ln_act_h2o ? st.ln_act_h2o[c] : 0.0;
    dry = sat < cfg.min_sat;
#pragma unroll
    for (int i = 0; i < N; i++) {
      if (i < naq) {
        double cc = st.pri_molal[i * ld + c];
        C(i) = cc;
        lngam[i] = log(st.pri_act_coef[i * ld + c]);
        LNA(i) = log(cc) + lngam[i];
        INVC(i) = 1.0 / cc;
      } else if (i < n) {
        C(i) = st.immobile[(i - naq) * ld + c];
      }
      if (i < n) RES(i) = 0.0;
    }
#pragma unroll 1
    for (int e = 0; e < n * cfg.js; e++) ws[cfg.off_J + e] = 0.0;
    if (cfg.mn_npref) {
      // prefactors on secondary species read ln(gamma) of the complexes the way RReact keeps them
      if (cfg.act_freq != PFRX_ACT_COEF_FREQUENCY_NEWTON_ITER) {
#pragma unroll 1
        for (int k = 0; k < cfg.ncplx; k++) ws[cfg.off_lng + k] = log(st.sec_act_coef[k * ld + c]);
      } else {
#pragma unroll 1
        for (int k = 0; k < cfg.ncplx; k++)
          if (cfg.cx_cls[k] >= 0) ws[cfg.off_cls + cfg.cx_cls[k]] = log(st.sec_act_coef[k * ld + c]);
#pragma unroll 1
        for (int i = 0; i < naq; i++)
          if (cfg.pri_cls[i] >= 0) ws[cfg.off_cls + cfg.pri_cls[i]] = log(st.pri_act_coef[i * ld + c]);
      }
    }
    if (!dry) {  // RReaction returns at once in a dry cell (reaction.F90:4085)
      if (cfg.nkin > 0) {
        kinetic_mineral(true);
#pragma unroll 1
        for (int k = 0; k < cfg.nkin; k++) st.mnrl_rate[k * ld + c] = ws[cfg.off_mn + k];
      }
      if (cfg.nmr > 0) {
        // RMultiRateSorption: free-site guesses and the sorbed totals of every rate from the state
#pragma unroll 1
        for (int k = 0; k < cfg.nsrfrxn; k++) ws[cfg.off_fs + k] = st.free_site[k * ld + c];
        multirate_begin(tran_dt);
        multirate(tran_dt);
      }
      if (cfg.need_dt) dtotal_from_state();
      if (cfg.need_ds) {
        // the sorbed inventory and d(total_sorb)/d(free) of RTAuxVarCompute; jscale 0 keeps the
        // accumulation derivative out of RReaction's Jacobian
        if (cfg.nmr == 0) {
#pragma unroll 1
          for (int k = 0; k < cfg.nsrfrxn; k++) ws[cfg.off_fs + k] = st.free_site[k * ld + c];
        }
#pragma unroll 1
        for (int r = 0; r < cfg.nionx; r++) ws[cfg.off_ix + r] = st.eqionx_ref ? st.eqionx_ref[r * ld + c] : 1.e-9;
        total_sorb(true, 0.0);
      }
      if (cfg.ngas > 0 && cfg.nrd > 0) total_gas(true, false, tran_dt);  // the decaying inventory in the gas phase
      if (cfg.nrd > 0) radioactive_decay();
      if (cfg.ngen > 0) general_reactions();
      if (cfg.nmb > 0) microbial();
      if (cfg.nidc > 0) immobile_decay();
      if (cfg.nsbx > 0) {
        sandbox_load(c);
        sandboxes(tran_dt);
        sandbox_store(c);
      }
    }
#pragma unroll 1
    for (int i = 0; i < n; i++) {
      res[i * ld + c] = RES(i);
      if (want_jac) {
#pragma unroll 1
        for (int j = 0; j < n; j++) jac[((int64_t)i * n + j) * ld + c] = J(i, j);
      }
    }
  }

  // ---- RStep (reaction.F90:3564-3738) ----------------------------------------------
  __device__ __forceinline__ void run(int64_t c, double target, int &nss, int &nit, int &nku, int &ierr, bool &had_cut) {
    cell = c;
    const int naq = cfg.naq, n = cfg.n, ncx = cfg.ncplx;
    const int64_t ld = st.ld;
    den_kg = st.den_kg[c];
    sat = st.sat[c];
    temp = st.temp[c];
    por = st.porosity[c];
    vol = st.volume[c];
    spd = st.soil_particle_density ? st.soil_particle_density[c] : 0.0;
    ln_act_h2o = st.ln_act_h2o ? st.ln_act_h2o[c] : 0.0;
    nss = nit = nku = ierr = 0;
    had_cut = false;
    if (!cfg.use_full_geochemistry) {
      for (int i = 0; i < naq; i++) st.pri_molal[i * ld + c] = st.total[i * ld + c] / den_kg * 1.e3;
      return;
    }
    const bool act_upd = cfg.act_freq == PFRX_ACT_COEF_FREQUENCY_NEWTON_ITER;
    // secondary sums of the incoming state (ionic strength of the first update)
    double Is = 0.0, ms = 0.0;
#pragma unroll 1
    for (int k = 0; k < ncx; k++) {
      double s = st.sec_molal[k * ld + c];
      Is += s * cfg.cx_Z2[k];
      ms += s;
      if (!act_upd) ws[cfg.off_lng + k] = log(st.sec_act_coef[k * ld + c]);
    }
    Isum = Is;
    msum = ms;
    if (act_upd && cfg.act_alg == PFRX_ACT_COEF_ALGORITHM_NEWTON) {
      // this branch may leave the coefficients it was entered with (|dI| < 1e-6 I at once)
#pragma unroll 1
      for (int k = 0; k < ncx; k++)
        if (cfg.cx_cls[k] >= 0) ws[cfg.off_cls + cfg.cx_cls[k]] = log(st.sec_act_coef[k * ld + c]);
#pragma unroll 1
      for (int i = 0; i < naq; i++)
        if (cfg.pri_cls[i] >= 0) ws[cfg.off_cls + cfg.pri_cls[i]] = log(st.pri_act_coef[i * ld + c]);
    }
#pragma unroll 1
    for (int k = 0; k < cfg.nsrfrxn; k++) ws[cfg.off_fs + k] = st.free_site[k * ld + c];
#pragma unroll 1
    for (int k = 0; k < cfg.nsrfcplx; k++) ws[cfg.off_sc + k] = 0.0;
#pragma unroll 1
    for (int k = 0; k < cfg.nkin; k++) ws[cfg.off_mn + k] = st.mnrl_rate[k * ld + c];
#pragma unroll 1
    for (int q = 0; q < cfg.nmr; q++) {
      int64_t base = (int64_t)naq * (cfg.mr_ptr[q] + q);
      for (int i = 0; i < naq; i++) ws[cfg.off_mr + (2 * q) * N + i] = st.kinmr[(base + i) * ld + c];
    }
    if (cfg.nsbx > 0) sandbox_load(c);
#pragma unroll 1
    for (int r = 0; r < cfg.nionx; r++)
      ws[cfg.off_ix + r] = st.eqionx_ref ? st.eqionx_ref[r * ld + c] : 1.e-9;
    small_mask = 0u;
#pragma unroll
    for (int i = 0; i < N; i++) {
      small_val[i] = 0.0;
      lngam[i] = 0.0;
      guess[i] = 1.0;
      if (i < naq) {
        lngam[i] = log(st.pri_act_coef[i * ld + c]);
        guess[i] = st.pri_molal[i * ld + c];
        double t = st.total[i * ld + c];
        if (t <= 1.e-40) {
          small_mask |= 1u << i;
          small_val[i] = t;
          t = 1.e-40;
          st.total[i * ld + c] = t;
        }
        if (cfg.use_total_as_guess) guess[i] = t;
      } else if (i < n) {
        double t = st.immobile[(i - naq) * ld + c];
        guess[i] = t;  // the guess keeps the unclamped value (pmc_subsurface_osrt.F90:356-362)
        if (t <= 1.e-40) {
          small_mask |= 1u << i;
          small_val[i] = t;
          st.immobile[(i - naq) * ld + c] = 1.e-40;
        }
      }
    }
    double cumulative = 0.0, dt = target;
    int ncuts = 0, nconst = 0;
    bool aborted = false;
    for (;;) {
      if (cumulative >= target) break;
      int its = 0;
      int e = react(dt, its);
      nit += its;
      if (!step_bookkeeping(e, dt, cumulative, target, ncuts, nconst, nss, nku, had_cut, aborted)) break;
    }
    if (aborted) ierr = 1;
#pragma unroll
    for (int i = 0; i < N; i++)
      if (i < naq) st.pri_molal[i * ld + c] = aborted ? C(i) : guess[i];
    if (!aborted) {
#pragma unroll
      for (int i = 0; i < N; i++) {
        if ((small_mask >> i) & 1u) {
          if (i < naq)
            st.total[i * ld + c] = small_val[i];
          else if (i < n)
            st.immobile[(i - naq) * ld + c] = small_val[i];
        }
      }
    }
    if (act_upd) {
#pragma unroll
      for (int i = 0; i < N; i++)
        if (i < naq) st.pri_act_coef[i * ld + c] = exp(lngam[i]);
#pragma unroll 1
      for (int k = 0; k < ncx; k++) {
        int q = cfg.cx_cls[k];
        st.sec_act_coef[k * ld + c] = q < 0 ? 1.0 : exp(ws[cfg.off_cls + q]);
      }
    }
#pragma unroll 1
    for (int k = 0; k < cfg.nsrfrxn; k++) st.free_site[k * ld + c] = ws[cfg.off_fs + k];
    if (cfg.neqsr > 0 && st.eqsrfcplx_conc)
      for (int k = 0; k < cfg.nsrfcplx; k++) st.eqsrfcplx_conc[k * ld + c] = ws[cfg.off_sc + k];
#pragma unroll 1
    for (int k = 0; k < cfg.nkin; k++) st.mnrl_rate[k * ld + c] = ws[cfg.off_mn + k];
#pragma unroll 1
    for (int q = 0; q < cfg.nmr; q++) {
      int64_t base = (int64_t)naq * (cfg.mr_ptr[q] + q);
      for (int i = 0; i < naq; i++) st.kinmr[(base + i) * ld + c] = ws[cfg.off_mr + (2 * q) * N + i];
    }
    if (st.ln_act_h2o && cfg.use_act_h2o) st.ln_act_h2o[c] = ln_act_h2o;
    if (cfg.n_nc > 0) sandbox_store(c);
    if (cfg.nionx > 0) {
      const int ncat = cfg.ix_ptr[cfg.nionx];
      if (st.eqionx_ref)
        for (int r = 0; r < cfg.nionx; r++) st.eqionx_ref[r * ld + c] = ws[cfg.off_ix + r];
      if (st.eqionx_conc)
        for (int k = 0; k < ncat; k++) st.eqionx_conc[k * ld + c] = ws[cfg.off_ix + cfg.nionx + k];
    }
  }

  // ---- ReactionEquilibrateConstraint (reaction.F90:1328-2117) for the cell's own constraint values.
  // Returns ierror (0 converged, 1 singular Jacobian, 2 non-positive concentration, 3 iteration limit).
  __device__ __forceinline__ int equilibrate(int64_t c, const DevCons &k, int &its_out) {
    cell = c;
    const int naq = cfg.naq, ncx = cfg.ncplx;
    const int64_t ld = st.ld;
    den_kg = st.den_kg[c];
    sat = st.sat[c];
    temp = st.temp[c];
    por = st.porosity[c];
    vol = st.volume[c];
    spd = st.soil_particle_density ? st.soil_particle_density[c] : 0.0;
    ln_act_h2o = 0.0;
    dry = false;
    its_out = 0;
    const double to_molar = k.init_molality ? den_kg / 1000.0 : 1.0;
    const double to_molal = k.init_molality ? 1.0 : 1000.0 / den_kg;
    if (!cfg.use_full_geochemistry) {  // reaction.F90:1472-1480
#pragma unroll 1
      for (int i = 0; i < naq; i++) {
        const double v = k.conc[i * ld + c];
        st.pri_molal[i * ld + c] = v * to_molal;
        st.total[i * ld + c] = v * to_molar;
      }
      return 0;
    }
    // guesses (reaction.F90:1489-1576)
#pragma unroll
    for (int i = 0; i < N; i++) {
      lngam[i] = 0.0;
      if (i < naq) {
        const double v = k.conc[i * ld + c];
        double g = 1.e-9;
        switch (k.type[i]) {
          case PFRX_CONSTRAINT_FREE:
          case PFRX_CONSTRAINT_CHARGE_BAL:
          case PFRX_CONSTRAINT_MINERAL: g = v * to_molal; break;
          case PFRX_CONSTRAINT_LOG: g = pow(10.0, v) * to_molal; break;
          case PFRX_CONSTRAINT_PH: g = pow(10.0, -v); break;
          default: break;
        }
        C(i) = g;
      }
    }
#pragma unroll 1
    for (int q = 0; q < cfg.ncls; q++) ws[cfg.off_cls + q] = 0.0;
    Isum = 0.0;
    msum = 0.0;
    int it = 0, it_act_on = 0;
    bool compute_act = false;
    const int max_it = k.max_iterations > 0 ? k.max_iterations : 10000;
    for (;;) {
#pragma unroll 1
      for (int i = 0; i < naq; i++) {  // reaction.F90:1592-1598
        const int t = k.type[i];
        if (t == PFRX_CONSTRAINT_FREE) C(i) = k.conc[i * ld + c] * to_molal;
        if (t == PFRX_CONSTRAINT_LOG) C(i) = pow(10.0, k.conc[i * ld + c]) * to_molal;
      }
      if (cfg.act_freq != PFRX_ACT_COEF_FREQUENCY_OFF && compute_act) activity();
      auxvar_compute(true, 1.0, true);  // RTotal: TOT = total, J = dtotal
      // charge-balance row before any row is replaced
      double zres = 0.0;
      bool any_z = false;
#pragma unroll 1
      for (int i = 0; i < naq; i++) any_z = any_z || k.type[i] == PFRX_CONSTRAINT_CHARGE_BAL;
      if (any_z) {
#pragma unroll 1
        for (int j = 0; j < naq; j++) {
          zres += k.Z[j] * TOT(j);
          double a = 0.0;
#pragma unroll 1
          for (int kk = 0; kk < naq; kk++) a += k.Z[kk] * J(kk, j);
          TMP(j) = a;
        }
      }
#pragma unroll 1
      for (int i = 0; i < naq; i++) {
        const int t = k.type[i];
        const double v = k.conc[i * ld + c];
        if (t == PFRX_CONSTRAINT_NULL || t == PFRX_CONSTRAINT_TOTAL) {
          RES(i) = TOT(i) - v * to_molar;
        } else if (t == PFRX_CONSTRAINT_CHARGE_BAL) {
          RES(i) = zres;
#pragma unroll 1
          for (int j = 0; j < naq; j++) J(i, j) = TMP(j);
        } else {
#pragma unroll 1
          for (int j = 0; j < naq; j++) J(i, j) = 0.0;
          if (t == PFRX_CONSTRAINT_FREE || t == PFRX_CONSTRAINT_LOG) {
            RES(i) = 0.0;
            J(i, i) = 1.0;
          } else if (t == PFRX_CONSTRAINT_PH) {
            RES(i) = 0.0;
            C(i) = pow(10.0, -v) / exp(lngam_at(i));
            J(i, i) = 1.0;
          } else {  // MINERAL, GAS
            double logK = (cfg.use_isothermal || !k.eq_logKcoef) ? k.eq_logK[i] : interp_logK(k.eq_logKcoef + 5 * i, temp);
            double lnQK = -logK * PFRX_LOG_TO_LN;
            if (k.eq_h2o[i] != 0.0) lnQK += k.eq_h2o[i] * ln_act_h2o;
#pragma unroll 1
            for (int p = k.eq_ptr[i]; p < k.eq_ptr[i + 1]; p++) {
              const int j = k.eq_spec[p];
              lnQK += k.eq_st[p] * log(C(j) * exp(lngam_at(j)));
              J(i, j) = k.eq_st[p] / C(j);
            }
            if (t == PFRX_CONSTRAINT_GAS) {
              const double pp = v <= 0.0 ? pow(10.0, v) : v;
              lnQK -= log(pp);
            }
            RES(i) = lnQK;
          }
        }
      }
      double max_res = 0.0;
#pragma unroll 1
      for (int i = 0; i < naq; i++) max_res = fmax(max_res, fabs(RES(i)));
      const bool use_log = cfg.use_log ? ((it > 3 && it < 9) ? (it % 2 == 0) : true) : false;
      if (!solve(naq, use_log)) {
        its_out = it;
        return 1;
      }
      double max_rel = 0.0, minc = 1.e300;
      bool nonpos = false;
      if (use_log) {
#pragma unroll 1
        for (int i = 0; i < naq; i++) {
          double u = RES(i);
          u = copysign(1.0, u) * fmin(fabs(u), cfg.max_dlnC);
          const double cc = C(i), cn = cc * exp(-u);
          C(i) = cn;
          max_rel = fmax(max_rel, fabs((cn - cc) / cc));
          if (!(cn > 0.0)) nonpos = true;
        }
      } else {
        double minr = 1.7976931348623157e308;
#pragma unroll 1
        for (int i = 0; i < naq; i++) {
          const double u = RES(i), cc = C(i);
          if (cc <= u) minr = fmin(minr, fabs(cc / u));
        }
#pragma unroll 1
        for (int i = 0; i < naq; i++) {
          double u = RES(i);
          const double cc = C(i);
          if (minr <= 1.0) u = u * minr * 0.99;
          const double cn = cc - u;
          C(i) = cn;
          max_rel = fmax(max_rel, fabs((cn - cc) / cc));
          if (!(cn > 0.0)) nonpos = true;
        }
      }
      (void)minc;
      it++;
      if (nonpos) {
        its_out = it;
        return 2;
      }
      if (it >= max_it) {
        its_out = it;
        return 3;
      }
      if (max_res < cfg.tol_res && max_rel < cfg.tol_relchange) {
        if (compute_act && it - it_act_on > 1) break;
        if (!compute_act) it_act_on = it;
        compute_act = true;
      }
    }
    its_out = it;
    // ---- the speciated state; total / sec_molal are the last RTotal's (before the final update)
#pragma unroll
    for (int i = 0; i < N; i++)
      if (i < naq) {
        st.pri_molal[i * ld + c] = C(i);
        st.total[i * ld + c] = TOT(i);
        st.pri_act_coef[i * ld + c] = exp(lngam[i]);
        LNA(i) = log(C(i)) + lngam[i];
        INVC(i) = 1.0 / C(i);
      }
#pragma unroll 1
    for (int kx = 0; kx < ncx; kx++) {
      const int q = cfg.cx_cls[kx];
      st.sec_act_coef[kx * ld + c] = q < 0 ? 1.0 : exp(ws[cfg.off_cls + q]);
    }
    if (st.ln_act_h2o) st.ln_act_h2o[c] = ln_act_h2o;
    // ---- once equilibrated, the sorbed concentrations (reaction.F90:2036-2060)
    if (cfg.nsorb > 0 || cfg.nmr > 0) {
#pragma unroll 1
      for (int r = 0; r < cfg.nsrfrxn; r++) ws[cfg.off_fs + r] = st.free_site[r * ld + c];
#pragma unroll 1
      for (int r = 0; r < cfg.nionx; r++) ws[cfg.off_ix + r] = st.eqionx_ref ? st.eqionx_ref[r * ld + c] : 1.e-9;
      if (cfg.nsorb > 0) {
        total_sorb(false, 0.0);
#pragma unroll 1
        for (int i = 0; i < naq; i++) st.total_sorb_eq[i * ld + c] = TS(i);
      }
      if (cfg.neqsr > 0 && st.eqsrfcplx_conc)
        for (int r = 0; r < cfg.nsrfcplx; r++) st.eqsrfcplx_conc[r * ld + c] = ws[cfg.off_sc + r];
      // RTotalSorbMultiRateAsEQ + the site fractions (reaction.F90:2062-2071)
#pragma unroll 1
      for (int q = 0; q < cfg.nmr; q++) {
        const int r0 = cfg.mr_ptr[q], r1 = cfg.mr_ptr[q + 1];
        double *seq = ws + cfg.off_mr + (2 * q) * N;
#pragma unroll 1
        for (int i = 0; i < naq; i++) seq[i] = 0.0;
        surf_cplx1(cfg.mr_rxn[q], seq, false, 0.0, false);
        const int64_t base = (int64_t)naq * (r0 + q);
#pragma unroll 1
        for (int i = 0; i < naq; i++) {
          st.kinmr[(base + i) * ld + c] = seq[i];
#pragma unroll 1
          for (int r = r0; r < r1; r++)
            st.kinmr[(base + (int64_t)naq * (r - r0 + 1) + i) * ld + c] = cfg.mr_frac[r] * seq[i];
        }
      }
#pragma unroll 1
      for (int r = 0; r < cfg.nsrfrxn; r++) st.free_site[r * ld + c] = ws[cfg.off_fs + r];
      if (cfg.nionx > 0) {
        const int ncat = cfg.ix_ptr[cfg.nionx];
        if (st.eqionx_ref)
          for (int r = 0; r < cfg.nionx; r++) st.eqionx_ref[r * ld + c] = ws[cfg.off_ix + r];
        if (st.eqionx_conc)
          for (int r = 0; r < ncat; r++) st.eqionx_conc[r * ld + c] = ws[cfg.off_ix + cfg.nionx + r];
      }
    }
    return 0;
  }

  // ---- one cell of RTUpdateAuxVars (reactive_transport.F90:3525-3660): [RActivityCoefficients,] RTAuxVarCompute
  // at the free-ion concentrations of the state (or of the block vector xx when given)
  __device__ __forceinline__ bool update_auxvars(int64_t c, const double *xx, bool update_act) {
    cell = c;
    const int naq = cfg.naq, n = cfg.n, ncx = cfg.ncplx;
    const int64_t ld = st.ld;
    den_kg = st.den_kg[c];
    sat = st.sat[c];
    temp = st.temp[c];
    por = st.porosity[c];
    vol = st.volume[c];
    spd = st.soil_particle_density ? st.soil_particle_density[c] : 0.0;
    ln_act_h2o = st.ln_act_h2o ? st.ln_act_h2o[c] : 0.0;
    dry = false;
    const bool per_iter = cfg.act_freq == PFRX_ACT_COEF_FREQUENCY_NEWTON_ITER;
    if (xx) {
#pragma unroll 1
      for (int i = 0; i < n; i++) {
        const double v = xx[c * n + i];
        if (i < naq)
          st.pri_molal[i * ld + c] = v;
        else
          st.immobile[(i - naq) * ld + c] = v;
      }
    }
    double Is = 0.0, ms = 0.0;
#pragma unroll 1
    for (int k = 0; k < ncx; k++) {
      const double sk = st.sec_molal[k * ld + c];
      Is += sk * cfg.cx_Z2[k];
      ms += sk;
      if (!per_iter) ws[cfg.off_lng + k] = log(st.sec_act_coef[k * ld + c]);
    }
    Isum = Is;
    msum = ms;
    const bool newton = cfg.act_alg == PFRX_ACT_COEF_ALGORITHM_NEWTON;
    if (per_iter || (update_act && newton)) {
      // coefficients per (Z, a0) class: the state's own when they are not refreshed here (with the
      // per-iteration update frequency the workspace has no per-complex slots)
#pragma unroll 1
      for (int k = 0; k < ncx; k++)
        if (cfg.cx_cls[k] >= 0) ws[cfg.off_cls + cfg.cx_cls[k]] = log(st.sec_act_coef[k * ld + c]);
#pragma unroll 1
      for (int i = 0; i < naq; i++)
        if (cfg.pri_cls[i] >= 0) ws[cfg.off_cls + cfg.pri_cls[i]] = log(st.pri_act_coef[i * ld + c]);
    }
#pragma unroll 1
    for (int k = 0; k < cfg.nsrfrxn; k++) ws[cfg.off_fs + k] = st.free_site[k * ld + c];
#pragma unroll 1
    for (int r = 0; r < cfg.nionx; r++) ws[cfg.off_ix + r] = st.eqionx_ref ? st.eqionx_ref[r * ld + c] : 1.e-9;
#pragma unroll
    for (int i = 0; i < N; i++) {
      lngam[i] = 0.0;
      if (i < naq) {
        lngam[i] = log(st.pri_act_coef[i * ld + c]);
        C(i) = st.pri_molal[i * ld + c];
      } else if (i < n) {
        C(i) = st.immobile[(i - naq) * ld + c];
      }
    }
    bool ok = true;
    const bool act = update_act && cfg.act_freq != PFRX_ACT_COEF_FREQUENCY_OFF;
    if (act) {
      if (newton)
        ok = activity_newton();
      else
        activity();
    }
    auxvar_compute(false, 1.0, false, act);
    store_total_gas();
#pragma unroll
    for (int i = 0; i < N; i++)
      if (i < naq) {
        st.total[i * ld + c] = TOT(i);
        if (act) st.pri_act_coef[i * ld + c] = exp(lngam[i]);
        if (cfg.nsorb > 0) st.total_sorb_eq[i * ld + c] = TS(i);
      }
    if (act) {
#pragma unroll 1
      for (int k = 0; k < ncx; k++) {
        const int q = cfg.cx_cls[k];
        st.sec_act_coef[k * ld + c] = q < 0 ? 1.0 : exp(ws[cfg.off_cls + q]);
      }
      if (st.ln_act_h2o && cfg.use_act_h2o) st.ln_act_h2o[c] = ln_act_h2o;
    }
#pragma unroll 1
    for (int k = 0; k < cfg.nsrfrxn; k++) st.free_site[k * ld + c] = ws[cfg.off_fs + k];
    if (cfg.neqsr > 0 && st.eqsrfcplx_conc)
      for (int k = 0; k < cfg.nsrfcplx; k++) st.eqsrfcplx_conc[k * ld + c] = ws[cfg.off_sc + k];
    if (cfg.nionx > 0) {
      const int ncat = cfg.ix_ptr[cfg.nionx];
      if (st.eqionx_ref)
        for (int r = 0; r < cfg.nionx; r++) st.eqionx_ref[r * ld + c] = ws[cfg.off_ix + r];
      if (st.eqionx_conc)
        for (int k = 0; k < ncat; k++) st.eqionx_conc[k * ld + c] = ws[cfg.off_ix + cfg.nionx + k];
    }
    return ok;
  }

  // ln gamma of primary species i with a run-time index (lngam[] lives in registers)
  __device__ __forceinline__ double lngam_at(int i) const {
    double g = 0.0;
#pragma unroll
    for (int x = 0; x < N; x++)
      if (x == i) g = lngam[x];
    return g;
  }

  // RStep's reaction to the outcome of one RReact (reaction.F90:3660-3716);
  // returns false when the cell is finished
  __device__ __forceinline__ bool step_bookkeeping(int e, double &dt, double &cumulative, double target, int &ncuts,
                                                   int &nconst, int &nss, int &nku, bool &had_cut, bool &aborted) {
    if (e != 0) {
      ncuts++;
      had_cut = true;
      if (ncuts > cfg.max_cuts) {
        aborted = true;
        return false;
      }
      dt = 0.5 * dt;
      nconst = 0;
    } else {
      bool upd = update_kinetic_state(dt);
      cumulative += dt;
      nss++;
      nconst++;
      if (upd) nku++;
      if (nconst >= 4) {
        ncuts--;
        dt = fmin(2.0 * dt, target - cumulative);
      }
    }
    return true;
  }
};

template <int N>
__global__ void __launch_bounds__(128, (N <= 4 ? 4 : (N <= 8 ? 2 : 1))) pfrx_rstep_tpc_kernel(DevCfg cfg, DevState st, int64_t ncell, double tran_dt,
                                                                 DevSummary *summ) {
  extern __shared__ double smem[];
  const int lane32 = threadIdx.x & 31;
  double *ws = smem + (size_t)threadIdx.x * cfg.ws_stride;
  CellT<N> sol(cfg, st, ws);

  unsigned long long l_active = 0, l_its = 0, l_cut = 0;
  long long l_first = -1;
  int l_maxits = 0, l_maxkin = 0, l_maxerr = 0, l_maxsub = 0;

  const int64_t gthread = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t c = gthread; c < ncell; c += nthreads) {
    int nss = 0, nit = 0, nku = 0, ierr = 0;
    bool cut = false;
    bool active = !(st.imat && st.imat[c] <= 0);
    if (active) sol.run(c, tran_dt, nss, nit, nku, ierr, cut);
    st.num_sub_steps[c] = nss;
    st.num_iterations[c] = nit;
    st.num_kinetic_state_updates[c] = nku;
    st.ierror[c] = ierr;
    if (active) {
      l_active++;
      l_its += (unsigned long long)nit;
      if (cut) l_cut++;
      if (ierr != 0 && (l_first < 0 || c < l_first)) l_first = c;
      l_maxits = max(l_maxits, nit);
      l_maxkin = max(l_maxkin, nku);
      l_maxerr = max(l_maxerr, ierr);
      l_maxsub = max(l_maxsub, nss);
    }
  }
  __syncwarp();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    l_active += __shfl_xor_sync(0xffffffffu, l_active, o);
    l_its += __shfl_xor_sync(0xffffffffu, l_its, o);
    l_cut += __shfl_xor_sync(0xffffffffu, l_cut, o);
    long long f = __shfl_xor_sync(0xffffffffu, l_first, o);
    if (f >= 0 && (l_first < 0 || f < l_first)) l_first = f;
    l_maxits = max(l_maxits, __shfl_xor_sync(0xffffffffu, l_maxits, o));
    l_maxkin = max(l_maxkin, __shfl_xor_sync(0xffffffffu, l_maxkin, o));
    l_maxerr = max(l_maxerr, __shfl_xor_sync(0xffffffffu, l_maxerr, o));
    l_maxsub = max(l_maxsub, __shfl_xor_sync(0xffffffffu, l_maxsub, o));
  }
  if (lane32 == 0) {
    atomicAdd(&summ->ncell_active, l_active);
    atomicAdd(&summ->sum_its, l_its);
    atomicAdd(&summ->num_cut_cells, l_cut);
    if (l_first >= 0) atomicMin(&summ->first_failed, l_first);
    atomicMax(&summ->max_its, l_maxits);
    atomicMax(&summ->max_kin, l_maxkin);
    atomicMax(&summ->max_err, l_maxerr);
    atomicMax(&summ->max_sub, l_maxsub);
  }
}

// RTUpdateAuxVars over the active cells: one thread per cell
template <int N>
__global__ void __launch_bounds__(128, (N <= 4 ? 4 : (N <= 8 ? 2 : 1)))
    pfrx_auxvars_tpc_kernel(DevCfg cfg, DevState st, int64_t ncell, const double *xx, int update_act) {
  extern __shared__ double smem[];
  double *ws = smem + (size_t)threadIdx.x * cfg.ws_stride;
  CellT<N> sol(cfg, st, ws);
  const int64_t gthread = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t c = gthread; c < ncell; c += nthreads)
    if (!(st.imat && st.imat[c] <= 0)) sol.update_auxvars(c, xx, update_act != 0);
}

// batched ReactionEquilibrateConstraint: one thread per cell
template <int N>
__global__ void __launch_bounds__(128, (N <= 4 ? 4 : (N <= 8 ? 2 : 1)))
    pfrx_constraint_tpc_kernel(DevCfg cfg, DevState st, int64_t ncell, DevCons k, int *num_its, int *ierror) {
  extern __shared__ double smem[];
  double *ws = smem + (size_t)threadIdx.x * cfg.ws_stride;
  CellT<N> sol(cfg, st, ws);
  const int64_t gthread = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t c = gthread; c < ncell; c += nthreads) {
    int its = 0, e = 0;
    if (!(st.imat && st.imat[c] <= 0)) e = sol.equilibrate(c, k, its);
    if (num_its) num_its[c] = its;
    if (ierror) ierror[c] = e;
  }
}

// batched RReaction(+Derivative): one thread per cell, inactive cells get zeros
template <int N>
__global__ void __launch_bounds__(128, (N <= 4 ? 4 : (N <= 8 ? 2 : 1)))
    pfrx_reaction_tpc_kernel(DevCfg cfg, DevState st, int64_t ncell, int want_jac, double *res, double *jac,
                             double tran_dt) {
  extern __shared__ double smem[];
  double *ws = smem + (size_t)threadIdx.x * cfg.ws_stride;
  CellT<N> sol(cfg, st, ws);
  const int64_t gthread = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t c = gthread; c < ncell; c += nthreads) {
    if (st.imat && st.imat[c] <= 0) {
      for (int i = 0; i < cfg.n; i++) {
        res[i * st.ld + c] = 0.0;
        if (want_jac)
          for (int j = 0; j < cfg.n; j++) jac[((int64_t)i * cfg.n + j) * st.ld + c] = 0.0;
      }
      continue;
    }
    sol.reaction(c, want_jac != 0, res, jac, tran_dt);
  }
}
