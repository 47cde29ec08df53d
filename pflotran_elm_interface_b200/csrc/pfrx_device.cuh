// pfrx_device.cuh -- device-side tables and the cell-group kernel of the
// operator-split chemistry step (RStep/RReact, reaction.F90:3564-4055).
//
// Design (DESIGN.md has the long form):
//  * a GROUP of L lanes (L = 1,2,4,8,16,32; a power of two inside one warp)
//    owns one cell at a time.  Row i of the Newton system belongs to lane
//    i % L (R = ceil(N/L) rows per lane) and lives in REGISTERS with
//    compile-time indices; everything a lane needs from its neighbours goes
//    through a small per-group shared-memory workspace guarded by
//    __syncwarp(group mask).  L = 1 is the thread-per-cell kernel.
//  * LU is right-looking with implicit-scaled partial pivoting; the pivot
//    row is published through shared memory, the argmax over the group is
//    three REDUX (hi word, lo word, logical row) -- same pivot order and the
//    same per-element operation order as the Crout loops of
//    utility.F90:597-688, forward substitution fused into the elimination.
//  * state is cell-major SoA in HBM, read once at cell entry and written once
//    at exit; stoichiometry tables are read-only and L1/L2 resident.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pfrx.h"

#define PFRX_LOG_TO_LN 2.30258509299  // pflotran_constants.F90:84 (truncated there)
#define PFRX_IDEAL_GAS_CONSTANT 8.31446

struct DevCfg {
  int naq, nim, n;
  int use_full_geochemistry, use_log, use_total_as_guess, use_isothermal;
  int act_freq, act_alg, use_act_h2o, h2o_aq_id;
  int max_its, max_cuts;
  double max_dlnC, tol_relchange, tol_res, tol_relres, min_sat;
  double debyeA, debyeB, debyeBdot;
  const double *pri_Z, *pri_a0;
  // complexes (CSR) + transposed lists (species -> complexes, CSC)
  int ncplx;
  const int *cx_ptr, *cx_id;
  const double *cx_st, *cx_h2o, *cx_logK, *cx_logKcoef, *cx_Z, *cx_a0;
  const int *sp_ptr, *sp_cx;  // species -> complex ids (ascending)
  const double *sp_st;        // matching stoichiometry nu_ki
  // kinetic minerals
  int nkin;
  const int *mn_ptr, *mn_id;
  const double *mn_st, *mn_h2o, *mn_logK, *mn_logKcoef, *mn_vol, *mn_rate, *mn_eact, *mn_thresh, *mn_limit;
  const int *mn_irrev;
  const double *mn_temkin, *mn_scale, *mn_power;
  // surface complexation
  int nsrfrxn, nsrfcplx;
  const int *sr_ptr, *sr_cx, *sr_type, *sr_surf, *sr_flag;
  const double *sr_dens;
  const int *sc_ptr, *sc_id;
  const double *sc_st, *sc_h2o, *sc_fs, *sc_logK, *sc_logKcoef;
  int neqsr;
  const int *eqsr;
  int nmr;
  const int *mr_rxn, *mr_ptr;
  const double *mr_rate, *mr_frac;
  // CLM-CN
  int cn_nrxn, cn_C, cn_N;
  const double *cn_CN, *cn_k, *cn_resp, *cn_inhib;
  const int *cn_nspec, *cn_cid, *cn_nid, *cn_up, *cn_down;
  // workspace layout (doubles, per group)
  int ws_stride, off_c, off_lnact, off_invc, off_x, off_xs, off_sec, off_secg, off_J, off_tmp, off_sc, js;
};

struct DevState {
  int64_t ld;
  double *total, *pri_molal, *immobile, *pri_act_coef, *sec_act_coef, *sec_molal, *ln_act_h2o;
  double *mnrl_volfrac, *mnrl_area, *mnrl_rate, *free_site, *eqsrfcplx_conc, *total_sorb_eq, *kinmr;
  const double *den_kg, *sat, *temp, *porosity, *volume, *soil_particle_density;
  const int *imat;
  int *num_sub_steps, *num_iterations, *num_kinetic_state_updates, *ierror;
};

// shard summary accumulated with atomics, one per block
struct DevSummary {
  unsigned long long ncell_active, sum_its, num_cut_cells;
  long long first_failed;
  int max_its, max_kin, max_err, max_sub;
};

// ---------------------------------------------------------------------------
template <int L>
struct Grp {
  unsigned mask;
  int lane;  // lane within group
  __device__ __forceinline__ void sync() const {
    if (L > 1) __syncwarp(mask);
  }
  __device__ __forceinline__ double maxd(double v) const {
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(mask, v, o));
    return v;
  }
  __device__ __forceinline__ double mind(double v) const {
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(mask, v, o));
    return v;
  }
  __device__ __forceinline__ double sumd(double v) const {
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
    return v;
  }
  __device__ __forceinline__ unsigned maxu(unsigned v) const {
    if (L > 1) return __reduce_max_sync(mask, v);
    return v;
  }
  __device__ __forceinline__ bool any(bool p) const {
    if (L > 1) return __ballot_sync(mask, p) != 0u;
    return p;
  }
};

__device__ __forceinline__ double interp_logK(const double *c, double temp) {
  // reaction_aux.F90:1285-1312
  double tk = temp + 273.15;
  return c[0] * log(tk) + c[1] + c[2] * tk + c[3] / tk + c[4] / (tk * tk);
}

// per-cell scalars every lane of the group carries
struct CellScalars {
  double den_kg, sat, temp, por, vol, spd, ln_act_h2o;
};

// ---------------------------------------------------------------------------
// The kernel.  N = padded system size (>= ncomp), L = lanes per cell.
template <int N, int L>
struct CellSolver {
  static constexpr int R = (N + L - 1) / L;
  const DevCfg &cfg;
  const DevState &st;
  Grp<L> g;
  double *ws;
  int64_t cell;
  CellScalars cs;
  // per-row registers (row i = lane + r*L)
  double cval[R];    // current iterate (pri_molal | immobile)
  double guess[R];   // RStep guess
  double totcur[R];  // rt_auxvar%total (aq rows) / rt_auxvar%immobile (imm rows)
  double sorbcur[R]; // rt_auxvar%total_sorb_eq (aq rows)
  double lngam[R];   // log(pri_act_coef) of aq rows
  double gam[R];
  double totnew[R];  // total(c) of the latest RTotal
  double sorbnew[R];
  bool dry;

  __device__ CellSolver(const DevCfg &c, const DevState &s, Grp<L> gg, double *w) : cfg(c), st(s), g(gg), ws(w) {}

  __device__ __forceinline__ int row(int r) const { return g.lane + r * L; }
  __device__ __forceinline__ double &W(int off, int i) { return ws[off + i]; }
  __device__ __forceinline__ double &Jm(int i, int j) { return ws[cfg.off_J + i * cfg.js + j]; }

  // logK at the cell temperature (RUpdateTempDependentCoefs, reaction.F90:5976)
  __device__ __forceinline__ double cx_logK(int k) const {
    return cfg.use_isothermal ? cfg.cx_logK[k] : interp_logK(cfg.cx_logKcoef + 5 * k, cs.temp);
  }
  __device__ __forceinline__ double mn_logK(int m) const {
    return (cfg.use_isothermal || !cfg.mn_logKcoef) ? cfg.mn_logK[m] : interp_logK(cfg.mn_logKcoef + 5 * m, cs.temp);
  }
  __device__ __forceinline__ double sc_logK(int k) const {
    return (cfg.use_isothermal || !cfg.sc_logKcoef) ? cfg.sc_logK[k] : interp_logK(cfg.sc_logKcoef + 5 * k, cs.temp);
  }

  // ---- RActivityCoefficients, LAG branch (reaction.F90:4553-4612) ----------
  // needs ws.c and ws.sec current; writes gam/lngam registers and ws.secg.
  __device__ void activity() {
    const int naq = cfg.naq, ncx = cfg.ncplx;
    double part = 0.0;
    for (int i = g.lane; i < naq; i += L) {
      double z = cfg.pri_Z[i];
      part += W(cfg.off_c, i) * z * z;
    }
    for (int k = g.lane; k < ncx; k += L) {
      double z = cfg.cx_Z[k];
      part += W(cfg.off_sec, k) * z * z;
    }
    double I = 0.5 * g.sumd(part);
    double sq = sqrt(I);
    double A = cfg.debyeA, B = cfg.debyeB, Bd = cfg.debyeBdot;
#pragma unroll
    for (int r = 0; r < R; r++) {
      int i = row(r);
      if (i < naq) {
        double z = cfg.pri_Z[i];
        if (fabs(z) > 1.e-10) {
          lngam[r] = (-z * z * sq * A / (1.0 + cfg.pri_a0[i] * B * sq) + Bd * I) * PFRX_LOG_TO_LN;
          gam[r] = exp(lngam[r]);
        } else {
          lngam[r] = 0.0;
          gam[r] = 1.0;
        }
      }
    }
    double sum_sec = 0.0;
    for (int k = g.lane; k < ncx; k += L) {
      double z = cfg.cx_Z[k];
      double gk = 1.0;
      if (fabs(z) > 1.e-10) gk = exp((-z * z * sq * A / (1.0 + cfg.cx_a0[k] * B * sq) + Bd * I) * PFRX_LOG_TO_LN);
      W(cfg.off_secg, k) = gk;
      sum_sec += W(cfg.off_sec, k);
    }
    if (cfg.use_act_h2o) {
      double sp = 0.0;
      for (int i = g.lane; i < naq; i += L)
        if (i != cfg.h2o_aq_id) sp += W(cfg.off_c, i);
      double t = 1.0 - 0.017 * (g.sumd(sp) + g.sumd(sum_sec));
      cs.ln_act_h2o = t > 0.0 ? log(t) : 0.0;
    }
    g.sync();
  }

  // ---- RTotalSorbEqSurfCplx1 (reaction_surf_complex.F90:641-900) -----------
  // Free-site solve is replicated on every lane (a handful of exps); each lane
  // then adds the rows it owns.  tot_sorb[r] accumulates nu*S; when add_J the
  // derivative rows are added into the shared Jacobian scaled by jscale.
  __device__ void surf_cplx1(int irxn, double *tot_sorb, bool add_J, double jscale, bool store_conc) {
    const int naq = cfg.naq;
    const int r0 = cfg.sr_ptr[irxn], r1 = cfg.sr_ptr[irxn + 1];
    double fs = fmax(st.free_site[irxn * st.ld + cell], 1.e-40);
    double dens;
    int ty = cfg.sr_type[irxn];
    if (ty == PFRX_MINERAL_SURFACE)
      dens = cfg.sr_dens[irxn] * st.mnrl_volfrac[cfg.sr_surf[irxn] * st.ld + cell];
    else if (ty == PFRX_ROCK_SURFACE)
      dens = cfg.sr_dens[irxn] * cs.spd * (1.0 - cs.por);
    else
      dens = cfg.sr_dens[irxn];
    if (dens < 1.e-40) {
      g.sync();
      if (g.lane == 0) {
        st.free_site[irxn * st.ld + cell] = 0.0;
        if (store_conc)
          for (int q = r0; q < r1; q++) W(cfg.off_sc, cfg.sr_cx[q]) = 0.0;
      }
      g.sync();
      return;
    }
    bool one_more = false;
    int it = 0;
    double damping = 1.0;
    // per-complex concentrations live in ws.tmp[N + q - r0] (replicated value)
    double *sconc = ws + cfg.off_tmp + N;
    for (;;) {
      it++;
      double total = fs;
      double lnfs = log(fs);
      for (int q = r0; q < r1; q++) {
        int k = cfg.sr_cx[q];
        double lnQK = -sc_logK(k) * PFRX_LOG_TO_LN;
        if (cfg.sc_h2o[k] != 0.0) lnQK += cfg.sc_h2o[k] * cs.ln_act_h2o;
        lnQK += cfg.sc_fs[k] * lnfs;
        for (int p = cfg.sc_ptr[k]; p < cfg.sc_ptr[k + 1]; p++) lnQK += cfg.sc_st[p] * W(cfg.off_lnact, cfg.sc_id[p]);
        double sck = exp(lnQK);
        if (g.lane == 0) sconc[q - r0] = sck;
        total += cfg.sc_fs[k] * sck;
      }
      g.sync();
      if (one_more) break;
      if (cfg.sr_flag[irxn]) {
        double res = dens - total;
        double d = 1.0;
        for (int q = r0; q < r1; q++) d += cfg.sc_fs[cfg.sr_cx[q]] * sconc[q - r0] / fs;
        double dfs = res / d;
        if (it > 1000) damping = 0.5;
        fs = fs + damping * dfs;
        if (fabs(dfs / fs) < 1.e-12 || it > 100000) one_more = true;
      } else {
        total = total / fs;
        fs = dens / total;
        one_more = true;
      }
      g.sync();
    }
    if (g.lane == 0) st.free_site[irxn * st.ld + cell] = fs;
    // dSx_dmi (eq. 2.3-46): the lane owning species j publishes entry j
    double denom = 0.0;
    for (int q = r0; q < r1; q++) {
      int k = cfg.sr_cx[q];
      denom += cfg.sc_fs[k] * cfg.sc_fs[k] * sconc[q - r0];
    }
    denom = denom / fs;
    denom = denom + 1.0;
#pragma unroll
    for (int r = 0; r < R; r++) {
      int i = row(r);
      if (i < naq) {
        double s = 0.0;
        for (int q = r0; q < r1; q++) {
          int k = cfg.sr_cx[q];
          for (int p = cfg.sc_ptr[k]; p < cfg.sc_ptr[k + 1]; p++)
            if (cfg.sc_id[p] == i) s += cfg.sc_st[p] * cfg.sc_fs[k] * sconc[q - r0];
        }
        s = -s / denom;
        W(cfg.off_tmp, i) = s / W(cfg.off_c, i);
      }
    }
    g.sync();
    if (store_conc && g.lane == 0)
      for (int q = r0; q < r1; q++) W(cfg.off_sc, cfg.sr_cx[q]) += sconc[q - r0];
#pragma unroll
    for (int r = 0; r < R; r++) {
      int i = row(r);
      if (i < naq) {
        for (int q = r0; q < r1; q++) {
          int k = cfg.sr_cx[q];
          double Sk = sconc[q - r0];
          int p0 = cfg.sc_ptr[k], p1 = cfg.sc_ptr[k + 1];
          for (int p = p0; p < p1; p++) {
            if (cfg.sc_id[p] != i) continue;
            double nui = cfg.sc_st[p];
            tot_sorb[r] += nui * Sk;
            if (add_J) {
              double nuiSx = cfg.sc_fs[k] * Sk / fs;
              for (int p2 = p0; p2 < p1; p2++) {
                int j = cfg.sc_id[p2];
                double t = cfg.sc_st[p2] * Sk / W(cfg.off_c, j) + nuiSx * W(cfg.off_tmp, j);
                Jm(i, j) += jscale * (nui * t);
              }
            }
          }
        }
      }
    }
    g.sync();
  }

  // ---- RTAuxVarCompute = RTotal (reaction.F90:4618-4759) --------------------
  // Leaves: ws.lnact, ws.invc, ws.sec; totnew/sorbnew registers; when
  // want_J the shared Jacobian holds d(accumulation)/dc / dt
  // (RTAccumulationDerivative + RAccumulationSorbDerivative).
  __device__ void auxvar_compute(bool want_J, double dt) {
    const int naq = cfg.naq, n = cfg.n, ncx = cfg.ncplx;
#pragma unroll
    for (int r = 0; r < R; r++) {
      int i = row(r);
      if (i < naq) {
        double lnc = log(cval[r]);
        W(cfg.off_lnact, i) = lnc + lngam[r];
        W(cfg.off_invc, i) = 1.0 / cval[r];
      }
      if (i < n) W(cfg.off_c, i) = cval[r];
    }
    g.sync();
    // secondary species: lanes stride over complexes
    for (int k = g.lane; k < ncx; k += L) {
      double lnQK = -cx_logK(k) * PFRX_LOG_TO_LN;
      double h = cfg.cx_h2o[k];
      if (h != 0.0) lnQK += h * cs.ln_act_h2o;
      for (int p = cfg.cx_ptr[k]; p < cfg.cx_ptr[k + 1]; p++) lnQK += cfg.cx_st[p] * W(cfg.off_lnact, cfg.cx_id[p]);
      W(cfg.off_sec, k) = exp(lnQK) / W(cfg.off_secg, k);
    }
    if (want_J) {
#pragma unroll
      for (int r = 0; r < R; r++) {
        int i = row(r);
        if (i < n)
          for (int j = 0; j < n; j++) Jm(i, j) = 0.0;
      }
    }
    g.sync();
    const double denL = cs.den_kg * 1.e-3;
    const double psvd = cs.por * cs.sat * 1000.0 * cs.vol / dt;
#pragma unroll
    for (int r = 0; r < R; r++) {
      int i = row(r);
      if (i < naq) {
        double tot = cval[r];
        for (int q = cfg.sp_ptr[i]; q < cfg.sp_ptr[i + 1]; q++) {
          int k = cfg.sp_cx[q];
          double nu_i = cfg.sp_st[q];
          double sk = W(cfg.off_sec, k);
          tot += nu_i * sk;
          if (want_J && !dry) {
            double t = nu_i * sk;
            for (int p = cfg.cx_ptr[k]; p < cfg.cx_ptr[k + 1]; p++) Jm(i, cfg.cx_id[p]) += cfg.cx_st[p] * t;
          }
        }
        totnew[r] = tot * denL;
        if (want_J) {
          if (dry) {
            Jm(i, i) = 1.0;
          } else {
            for (int j = 0; j < naq; j++) {
              double d = Jm(i, j) * W(cfg.off_invc, j) + (i == j ? 1.0 : 0.0);
              Jm(i, j) = (d * denL) * psvd;
            }
          }
        }
      } else if (i < n) {
        totnew[r] = cval[r];
        if (want_J) Jm(i, i) = dry ? 1.0 : cs.vol / dt;
      }
      sorbnew[r] = 0.0;
    }
    // equilibrium sorption (RTotalSorb, reaction.F90:4783)
    if (cfg.neqsr > 0) {
      if (g.lane == 0)
        for (int k = 0; k < cfg.nsrfcplx; k++) W(cfg.off_sc, k) = 0.0;
      g.sync();
      for (int e = 0; e < cfg.neqsr; e++) surf_cplx1(cfg.eqsr[e], sorbnew, want_J, cs.vol / dt, true);
    }
  }

  // ---- RKineticMineral (reaction_mineral.F90:647-1078), no prefactors ------
  // Every lane evaluates the rate scalars; row owners add their entries.
  // ws.lnact must be current.  rate_out (may be null) receives mnrl_rate.
  __device__ void kinetic_mineral(double *res, bool derivative, bool store_rate) {
    const int naq = cfg.naq;
    for (int m = 0; m < cfg.nkin; m++) {
      const int p0 = cfg.mn_ptr[m], p1 = cfg.mn_ptr[m + 1];
      double rate_vol = 0.0;  // mnrl_rate default
      double lnQK = -mn_logK(m) * PFRX_LOG_TO_LN;
      if (cfg.mn_h2o[m] != 0.0) lnQK += cfg.mn_h2o[m] * cs.ln_act_h2o;
      for (int p = p0; p < p1; p++) lnQK += cfg.mn_st[p] * W(cfg.off_lnact, cfg.mn_id[p]);
      double QK = exp(lnQK);
      double aff;
      if (cfg.mn_temkin) {
        if (cfg.mn_scale)
          aff = 1.0 - pow(QK, 1.0 / (cfg.mn_scale[m] * cfg.mn_temkin[m]));
        else
          aff = 1.0 - pow(QK, 1.0 / cfg.mn_temkin[m]);
      } else if (cfg.mn_scale) {
        aff = 1.0 - pow(QK, 1.0 / cfg.mn_scale[m]);
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
        double arr = 1.0;
        if (cfg.mn_eact[m] > 0.0)
          arr = exp(cfg.mn_eact[m] / PFRX_IDEAL_GAS_CONSTANT * (1.0 / (25.0 + 273.15) - 1.0 / (cs.temp + 273.15)));
        double spr = cfg.mn_rate[m] * arr;
        double Im_const = -st.mnrl_area[m * st.ld + cell];
        if (cfg.mn_scale) Im_const = Im_const / cfg.mn_scale[m];
        double Im;
        if (cfg.mn_power)
          Im = Im_const * sgn * pow(fabs(aff), cfg.mn_power[m]) * spr;
        else
          Im = Im_const * sgn * fabs(aff) * spr;
        rate_vol = Im;
        Im_const = Im_const * cs.vol;
        Im = Im * cs.vol;
        double dIm_dQK = 0.0;
        if (derivative) {
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
        }
        double den = 1.0, limfac = 1.0;
        if (lim > 0.0) {
          den = 1.0 + (1.0 - aff) / lim;
          limfac = (1.0 + QK / lim / den);
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
          int i = row(r);
          if (i >= naq) continue;
          for (int p = p0; p < p1; p++) {
            if (cfg.mn_id[p] != i) continue;
            double nui = cfg.mn_st[p];
            if (res) res[r] += nui * Im;
            if (derivative) {
              for (int p2 = p0; p2 < p1; p2++) {
                int j = cfg.mn_id[p2];
                // exp(-ln c_j) = 1/c_j
                double dQK_dmj = (cfg.mn_st[p2] * QK * W(cfg.off_invc, j)) * cs.den_kg * 1.e-3;
                if (lim > 0.0)
                  Jm(i, j) += nui * dIm_dQK * limfac * dQK_dmj / den;
                else
                  Jm(i, j) += nui * dIm_dQK * dQK_dmj;
              }
            }
          }
        }
      }
      if (store_rate && g.lane == 0) st.mnrl_rate[m * st.ld + cell] = rate_vol;
    }
  }

  // ---- RMultiRateSorption (reaction_surf_complex.F90:552-637) --------------
  __device__ void multirate(double *res, double dt) {
    const int naq = cfg.naq;
    for (int q = 0; q < cfg.nmr; q++) {
      int irxn = cfg.mr_rxn[q];
      int r0 = cfg.mr_ptr[q], r1 = cfg.mr_ptr[q + 1];
      int64_t base = (int64_t)naq * (r0 + q);
      double seq[R];
#pragma unroll
      for (int r = 0; r < R; r++) seq[r] = 0.0;
      // sum_k V k/(1+k dt) f : the Jacobian gets dtotal_sorb_eq times this
      double jsum = 0.0;
      for (int k = r0; k < r1; k++) {
        double kk = cfg.mr_rate[k] / (1.0 + cfg.mr_rate[k] * dt);
        jsum += cs.vol * kk * cfg.mr_frac[k];
      }
      surf_cplx1(irxn, seq, true, jsum, false);
#pragma unroll
      for (int r = 0; r < R; r++) {
        int i = row(r);
        if (i < naq) {
          for (int k = r0; k < r1; k++) {
            double kdt = cfg.mr_rate[k] * dt;
            double kk = cfg.mr_rate[k] / (1.0 + kdt);
            double S = st.kinmr[(base + (int64_t)naq * (k - r0 + 1) + i) * st.ld + cell];
            res[r] += cs.vol * kk * (cfg.mr_frac[k] * seq[r] - S);
          }
          st.kinmr[(base + i) * st.ld + cell] = seq[r];
        }
      }
    }
  }

  // ---- CLM_CN_React (reaction_sandbox_clm_cn.F90:468-787) ------------------
  __device__ void clm_cn(double *res, bool derivative) {
    const int off = cfg.naq;
    double temp_K = cs.temp + 273.15;
    if (!(temp_K > 227.15)) return;
    const double one_over_71_02 = 1.408054069e-2, theta_min = 0.01, one_over_log_theta_min = -2.17147241e-1;
    double F_t = exp(308.56 * (one_over_71_02 - 1.0 / (temp_K - 227.13)));
    double F_theta = log(theta_min / fmax(theta_min, cs.sat)) * one_over_log_theta_min;
    double cinh = F_t * F_theta;
    const int ires_C = off + cfg.cn_C, ispec_N = cfg.cn_N, ires_N = off + ispec_N;
    const double *imm = ws + cfg.off_c + off;  // immobile(:) of the current iterate
    auto addR = [&](int irow, double v) {
#pragma unroll
      for (int r = 0; r < R; r++)
        if (row(r) == irow) res[r] += v;
    };
    auto addJ = [&](int irow, int jcol, double v) {
      if ((irow % L) == g.lane) Jm(irow, jcol) += v;
    };
    for (int x = 0; x < cfg.cn_nrxn; x++) {
      double src = cfg.cn_k[x] * cs.vol * cinh;
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
      addR(ires_C, -(sC * rate));
      addR(ires_N, -(sN * rate));
      addR(rUC, -((-1.0) * sUC * rate));
      if (!constCN) addR(rUN, -((-1.0) * sUN * rate));
      if (id >= 0) addR(rD, -(sDC * rate));
      if (derivative) {
        double drate = src * Ninh;
        double dInh = 0.0;
        addJ(rUC, rUC, -((-1.0) * sUC * drate));
        if (useInh) {
          dInh = imm[iC] * src * dNinh;
          addJ(rUC, ires_N, -((-1.0) * sUC * dInh));
        }
        if (id >= 0) {
          addJ(rD, rUC, -(sDC * drate));
          if (useInh) addJ(rD, ires_N, -(sDC * dInh));
        }
        if (!constCN) {
          addJ(rUN, rUC, -((-1.0) * sUN * drate));
          if (useInh) addJ(rUN, ires_N, -((-1.0) * sUN * dInh));
          double nc = imm[iN] / imm[iC] * src * Ninh;
          addJ(rUN, rUC, -((-1.0) * (-1.0) * nc));
          addJ(rUN, rUN, -((-1.0) * src * Ninh));
          addJ(ires_N, rUC, -((-1.0) * nc));
          addJ(ires_N, rUN, -(src * Ninh));
        }
        addJ(ires_C, rUC, -(sC * drate));
        addJ(ires_N, rUC, -(sN * drate));
        if (useInh) {
          addJ(ires_C, ires_N, -(sC * dInh));
          addJ(ires_N, ires_N, -(sN * dInh));
        }
      }
    }
  }

  // ---- RSolve + LUDecomposition + LUBackSubstitution -----------------------
  // (reaction.F90:5457-5516, utility.F90:597-735).  a[][] rows in registers.
  // Returns false when a row is all zero (singular) -- the group agrees.
  __device__ bool solve(double (&a)[R][N], double (&b)[R], double *xout /* ws, n entries */) {
    const int n = cfg.n;
    double vv[R];
    int pos[R], step[R];
    bool done[R];
    bool bad = false;
#pragma unroll
    for (int r = 0; r < R; r++) {
      int i = row(r);
      pos[r] = i;
      step[r] = -1;
      done[r] = !(i < n);
      double m = 0.0;
#pragma unroll
      for (int j = 0; j < N; j++) m = fmax(m, fabs(a[r][j]));
      if (i < n && !(m > 0.0)) bad = true;
      vv[r] = 1. / m;
    }
    if (g.any(bad)) return false;
    double *xb = ws + cfg.off_x;  // two pivot-row buffers of N+2
#pragma unroll
    for (int k = 0; k < N; k++) {
      if (k < n) {
        double bestv = 0.0;
        int bestpos = -1, bestr = 0;
#pragma unroll
        for (int r = 0; r < R; r++) {
          if (!done[r]) {
            double dum = vv[r] * fabs(a[r][k]);
            if (dum > bestv || (dum == bestv && pos[r] > bestpos)) {
              bestv = dum;
              bestpos = pos[r];
              bestr = r;
            }
          }
        }
        int wpos = bestpos;
        bool iam = bestpos >= 0;
        if (L > 1) {
          unsigned hi = iam ? (unsigned)__double2hiint(bestv) : 0u;
          unsigned mhi = g.maxu(hi);
          unsigned lo = (iam && hi == mhi) ? (unsigned)__double2loint(bestv) : 0u;
          unsigned mlo = g.maxu(lo);
          unsigned pk = (iam && hi == mhi && lo == mlo) ? (unsigned)(bestpos + 1) : 0u;
          unsigned mpk = g.maxu(pk);
          wpos = (int)mpk - 1;
          iam = iam && pk == mpk;
        }
        if (wpos < 0) {
          // every candidate was NaN: fall back to the row sitting at k
          wpos = k;
          iam = false;
#pragma unroll
          for (int r = 0; r < R; r++)
            if (!done[r] && pos[r] == k) {
              iam = true;
              bestr = r;
            }
        }
        double *buf = xb + (k & 1) * (N + 2);
#pragma unroll
        for (int r = 0; r < R; r++) {
          if (iam && r == bestr) {
            double p = a[r][k];
            if (p == 0.0) {
              p = 1.0e-20;
              a[r][k] = p;
            }
#pragma unroll
            for (int j = k + 1; j < N; j++) buf[j] = a[r][j];
            buf[k] = 1.0 / p;
            buf[N] = b[r];
            done[r] = true;
            step[r] = k;
          } else if (!done[r] && pos[r] == k) {
            pos[r] = wpos;  // rows j and imax trade places (utility.F90:660-668)
          }
        }
        g.sync();
        double pinv = buf[k], yk = buf[N];
#pragma unroll
        for (int r = 0; r < R; r++) {
          if (!done[r]) {
            double m = a[r][k] * pinv;
            a[r][k] = m;
#pragma unroll
            for (int j = k + 1; j < N; j++) a[r][j] -= m * buf[j];
            b[r] -= m * yk;
          }
        }
      }
    }
    // back substitution, column oriented
#pragma unroll
    for (int s = N - 1; s >= 0; s--) {
      if (s < n) {
#pragma unroll
        for (int r = 0; r < R; r++)
          if (step[r] == s) xout[s] = b[r] / a[r][s];
        g.sync();
        double xs = xout[s];
#pragma unroll
        for (int r = 0; r < R; r++)
          if (step[r] >= 0 && step[r] < s) b[r] -= a[r][s] * xs;
      }
    }
    g.sync();
    return true;
  }

  // ---- RReact (reaction.F90:3742-4055) -------------------------------------
  // returns ierror; its = Newton iterations used
  __device__ int react(double dt, int &its_out) {
    const int naq = cfg.naq, n = cfg.n;
    double fixed[R], init_tot[R], res[R];
    const double psv = cs.por * cs.sat * 1000.0 * cs.vol;
    dry = cs.sat < cfg.min_sat;
#pragma unroll
    for (int r = 0; r < R; r++) {
      int i = row(r);
      // RTAccumulation (+RAccumulationSorb) of total^*, immobile^k
      double f = 0.0;
      if (!dry) {
        if (i < naq) {
          f = psv * totcur[r];
        } else if (i < n) {
          f = 0.0 + totcur[r] * cs.vol;
        }
      }
      if (cfg.neqsr > 0 && i < naq) f = f + sorbcur[r] * cs.vol;
      fixed[r] = f;
      init_tot[r] = totcur[r];
      cval[r] = guess[r];
    }
    int its = 0;
    double norm0 = 0.0;
    int ierr = 0;
    for (;;) {
      its++;
      if (cfg.act_freq == PFRX_ACT_COEF_FREQUENCY_NEWTON_ITER) {
        // needs ws.c of the current iterate
#pragma unroll
        for (int r = 0; r < R; r++)
          if (row(r) < n) W(cfg.off_c, row(r)) = cval[r];
        g.sync();
        activity();
      }
      auxvar_compute(true, dt);
      if (its > cfg.max_its) {
        ierr = 1;
#pragma unroll
        for (int r = 0; r < R; r++) {
          totcur[r] = init_tot[r];  // total and immobile restored (reaction.F90:3891-3894)
          sorbcur[r] = sorbnew[r];  // total_sorb_eq is not
        }
        // immobile rows: rt_auxvar%immobile = initial_total -> cval mirrors it
#pragma unroll
        for (int r = 0; r < R; r++)
          if (row(r) >= naq) cval[r] = init_tot[r];
        its_out = its;
        return ierr;
      }
#pragma unroll
      for (int r = 0; r < R; r++) {
        int i = row(r);
        double a = 0.0;
        if (!dry) {
          if (i < naq)
            a = psv * totnew[r];
          else if (i < n)
            a = 0.0 + cval[r] * cs.vol;
        }
        if (cfg.neqsr > 0 && i < naq) a = a + sorbnew[r] * cs.vol;
        res[r] = (a - fixed[r]) / dt;
      }
      // RReaction (reaction.F90:4059-4130), same order
      if (!dry) {
        if (cfg.nkin > 0) kinetic_mineral(res, true, true);
        if (cfg.nmr > 0) multirate(res, dt);
        if (cfg.cn_nrxn > 0) clm_cn(res, true);
      }
      g.sync();
      double mabs = 0.0, ss = 0.0;
#pragma unroll
      for (int r = 0; r < R; r++) {
        if (row(r) < n) {
          mabs = fmax(mabs, fabs(res[r]));
          ss += res[r] * res[r];
        }
      }
      mabs = g.maxd(mabs);
      double nrm = sqrt(g.sumd(ss));
      if (its == 1) norm0 = nrm;
      double rel = nrm / norm0;
      if (mabs < cfg.tol_res) break;
      if (rel < cfg.tol_relres) break;

      // RSolve: row scaling, optional d/dlnc scaling, LU, back-substitution
      double a[R][N], b[R];
#pragma unroll
      for (int r = 0; r < R; r++) {
        int i = row(r);
        double m = 0.0;
#pragma unroll
        for (int j = 0; j < N; j++) {
          double v = (i < n && j < n) ? Jm(i, j) : 0.0;
          a[r][j] = v;
          m = fmax(m, fabs(v));
        }
        double nm = 1.0 / fmax(1.0, m);
        b[r] = res[r] * nm;
#pragma unroll
        for (int j = 0; j < N; j++) {
          double v = a[r][j] * nm;
          if (cfg.use_log && j < n) v = v * W(cfg.off_c, j);
          a[r][j] = v;
        }
      }
      double *xs = ws + cfg.off_xs;
      if (!solve(a, b, xs)) {
        ierr = 1;  // solve_error branch: no restore (reaction.F90:3964-3967)
#pragma unroll
        for (int r = 0; r < R; r++) {
          totcur[r] = totnew[r];
          sorbcur[r] = sorbnew[r];
        }
        its_out = its;
        return ierr;
      }
      double cnew[R], maxrel = 0.0;
      bool anyv = false;
      if (cfg.use_log) {
#pragma unroll
        for (int r = 0; r < R; r++) {
          int i = row(r);
          if (i < n) {
            double u = xs[i];
            u = copysign(1.0, u) * fmin(fabs(u), cfg.max_dlnC);
            cnew[r] = cval[r] * exp(-u);
          }
        }
      } else {
        double minr = 1.e20;
#pragma unroll
        for (int r = 0; r < R; r++) {
          int i = row(r);
          if (i < n) {
            double u = xs[i];
            if (cval[r] <= u) minr = fmin(minr, fabs(cval[r] / u));
          }
        }
        minr = g.mind(minr);
#pragma unroll
        for (int r = 0; r < R; r++) {
          int i = row(r);
          if (i < n) {
            double u = xs[i];
            if (minr < 1.0) u = u * minr * 0.99;
            cnew[r] = cval[r] - u;
          }
        }
      }
#pragma unroll
      for (int r = 0; r < R; r++) {
        if (row(r) < n) {
          double v = fabs((cnew[r] - cval[r]) / cval[r]);
          if (!isnan(v)) {
            maxrel = anyv ? fmax(maxrel, v) : v;
            anyv = true;
          }
        }
      }
      // NaN-skipping maxval, like gfortran's MAXVAL
      double mr = g.maxd(anyv ? maxrel : -1.0);
      bool conv = (mr >= 0.0) && (mr < cfg.tol_relchange);
      if (conv) break;
#pragma unroll
      for (int r = 0; r < R; r++)
        if (row(r) < n) cval[r] = cnew[r];
      g.sync();
    }
    // one last update (reaction.F90:4052)
    auxvar_compute(false, dt);
#pragma unroll
    for (int r = 0; r < R; r++) {
      totcur[r] = totnew[r];
      sorbcur[r] = sorbnew[r];
    }
    its_out = its;
    return 0;
  }

  // ---- RUpdateKineticState (reaction.F90:5935-5972) ------------------------
  __device__ bool update_kinetic_state(double dt) {
    bool updated = false;
    if (cfg.nkin > 0) {
      updated = true;
      // ws.lnact is that of the converged iterate (last auxvar_compute)
      kinetic_mineral(nullptr, false, true);
      g.sync();
      if (g.lane == 0) {
        for (int m = 0; m < cfg.nkin; m++) {
          double rate = st.mnrl_rate[m * st.ld + cell];
          double vf = st.mnrl_volfrac[m * st.ld + cell] + rate * cfg.mn_vol[m] * dt;
          if (vf < 0.0) vf = 0.0;
          st.mnrl_volfrac[m * st.ld + cell] = vf;
        }
      }
      g.sync();
    }
    for (int q = 0; q < cfg.nmr; q++) {
      updated = true;
      int r0 = cfg.mr_ptr[q], r1 = cfg.mr_ptr[q + 1];
      int64_t base = (int64_t)cfg.naq * (r0 + q);
#pragma unroll
      for (int r = 0; r < R; r++) {
        int i = row(r);
        if (i < cfg.naq) {
          double seq = st.kinmr[(base + i) * st.ld + cell];
          for (int k = r0; k < r1; k++) {
            double kdt = cfg.mr_rate[k] * dt;
            int64_t ix = (base + (int64_t)cfg.naq * (k - r0 + 1) + i) * st.ld + cell;
            st.kinmr[ix] = (st.kinmr[ix] + kdt * cfg.mr_frac[k] * seq) / (1.0 + kdt);
          }
        }
      }
    }
    if (cfg.cn_nrxn > 0) updated = true;  // any sandbox => true
    return updated;
  }

  // ---- RStep (reaction.F90:3564-3738) on cell `c` ---------------------------
  __device__ void run(int64_t c, double target, int &nss, int &nit, int &nku, int &ierr, bool &had_cut) {
    cell = c;
    const int naq = cfg.naq, n = cfg.n, ncx = cfg.ncplx;
    const int64_t ld = st.ld;
    cs.den_kg = st.den_kg[c];
    cs.sat = st.sat[c];
    cs.temp = st.temp[c];
    cs.por = st.porosity[c];
    cs.vol = st.volume[c];
    cs.spd = st.soil_particle_density ? st.soil_particle_density[c] : 0.0;
    cs.ln_act_h2o = st.ln_act_h2o ? st.ln_act_h2o[c] : 0.0;
    nss = nit = nku = ierr = 0;
    had_cut = false;
    bool small[R];
    double small_val[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
      int i = row(r);
      small[r] = false;
      small_val[r] = 0.0;
      gam[r] = 1.0;
      lngam[r] = 0.0;
      sorbcur[r] = 0.0;
      totcur[r] = 0.0;
      guess[r] = 1.0;
      cval[r] = 1.0;
      if (i < naq) {
        totcur[r] = st.total[i * ld + c];
        guess[r] = st.pri_molal[i * ld + c];
        gam[r] = st.pri_act_coef[i * ld + c];
        lngam[r] = log(gam[r]);
        if (cfg.neqsr > 0) sorbcur[r] = st.total_sorb_eq[i * ld + c];
      } else if (i < n) {
        totcur[r] = st.immobile[(i - naq) * ld + c];
        guess[r] = totcur[r];
      }
    }
    if (!cfg.use_full_geochemistry) {
#pragma unroll
      for (int r = 0; r < R; r++)
        if (row(r) < naq) st.pri_molal[row(r) * ld + c] = totcur[r] / cs.den_kg * 1.e3;
      return;
    }
    for (int k = g.lane; k < ncx; k += L) {
      W(cfg.off_sec, k) = st.sec_molal[k * ld + c];
      W(cfg.off_secg, k) = st.sec_act_coef[k * ld + c];
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
      int i = row(r);
      if (i < n && totcur[r] <= 1.e-40) {
        small[r] = true;
        small_val[r] = totcur[r];
        totcur[r] = 1.e-40;
        if (i >= naq) guess[r] = guess[r];  // guess was read before the clamp (pmc_subsurface_osrt.F90:356-362)
      }
      if (cfg.use_total_as_guess && i < naq) guess[r] = totcur[r];
    }
    g.sync();
    double cumulative = 0.0, dt = target;
    int ncuts = 0, nconst = 0;
    bool aborted = false;
    for (;;) {
      if (cumulative >= target) break;
      int its = 0;
      int e = react(dt, its);
      nit += its;
      if (e != 0) {
        ncuts++;
        had_cut = true;
        if (ncuts > cfg.max_cuts) {
          ierr = 1;
          aborted = true;
          break;
        }
        dt = 0.5 * dt;
        nconst = 0;
      } else {
        bool upd = update_kinetic_state(dt);
        cumulative += dt;
        nss++;
        nconst++;
        if (upd) nku++;
#pragma unroll
        for (int r = 0; r < R; r++) guess[r] = cval[r];
        if (nconst >= 4) {
          ncuts--;
          dt = fmin(2.0 * dt, target - cumulative);
        }
      }
    }
    if (!aborted) {
#pragma unroll
      for (int r = 0; r < R; r++)
        if (small[r]) totcur[r] = small_val[r];
    }
    // write back (the rt_auxvar of the cell)
#pragma unroll
    for (int r = 0; r < R; r++) {
      int i = row(r);
      if (i < naq) {
        st.total[i * ld + c] = totcur[r];
        st.pri_molal[i * ld + c] = cval[r];
        st.pri_act_coef[i * ld + c] = gam[r];
        if (cfg.neqsr > 0) st.total_sorb_eq[i * ld + c] = sorbcur[r];
      } else if (i < n) {
        st.immobile[(i - naq) * ld + c] = small[r] && !aborted ? small_val[r] : (aborted ? totcur[r] : cval[r]);
      }
    }
    g.sync();
    for (int k = g.lane; k < ncx; k += L) {
      st.sec_molal[k * ld + c] = W(cfg.off_sec, k);
      st.sec_act_coef[k * ld + c] = W(cfg.off_secg, k);
    }
    if (cfg.neqsr > 0 && st.eqsrfcplx_conc)
      for (int k = g.lane; k < cfg.nsrfcplx; k += L) st.eqsrfcplx_conc[k * ld + c] = W(cfg.off_sc, k);
    if (g.lane == 0 && st.ln_act_h2o) st.ln_act_h2o[c] = cs.ln_act_h2o;
    g.sync();
  }
};

template <int N, int L>
__global__ void __launch_bounds__(128) pfrx_rstep_kernel(DevCfg cfg, DevState st, int64_t ncell, double tran_dt,
                                                          DevSummary *summ) {
  extern __shared__ double smem[];
  constexpr int CPW = 32 / L;  // cells per warp pass
  const int lane32 = threadIdx.x & 31;
  const int warp_in_block = threadIdx.x >> 5;
  const int warps_per_block = blockDim.x >> 5;
  Grp<L> g;
  g.lane = lane32 % L;
  g.mask = (L == 32) ? 0xffffffffu : (((1u << L) - 1u) << (lane32 - g.lane));
  const int grp_in_block = threadIdx.x / L;
  double *ws = smem + (size_t)grp_in_block * cfg.ws_stride;
  CellSolver<N, L> sol(cfg, st, g, ws);

  unsigned long long l_active = 0, l_its = 0, l_cut = 0;
  long long l_first = -1;
  int l_maxits = 0, l_maxkin = 0, l_maxerr = 0, l_maxsub = 0;

  const int64_t gwarp = (int64_t)blockIdx.x * warps_per_block + warp_in_block;
  const int64_t nwarps = (int64_t)gridDim.x * warps_per_block;
  for (int64_t base = gwarp * CPW; base < ncell; base += nwarps * CPW) {
    int64_t c = base + lane32 / L;
    if (c >= ncell) continue;
    if (st.imat && st.imat[c] <= 0) {
      if (g.lane == 0) {
        st.num_sub_steps[c] = 0;
        st.num_iterations[c] = 0;
        st.num_kinetic_state_updates[c] = 0;
        st.ierror[c] = 0;
      }
      continue;
    }
    int nss, nit, nku, ierr;
    bool cut;
    sol.run(c, tran_dt, nss, nit, nku, ierr, cut);
    if (g.lane == 0) {
      st.num_sub_steps[c] = nss;
      st.num_iterations[c] = nit;
      st.num_kinetic_state_updates[c] = nku;
      st.ierror[c] = ierr;
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
  // shard summary: warp reduce, then one set of atomics per warp
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
