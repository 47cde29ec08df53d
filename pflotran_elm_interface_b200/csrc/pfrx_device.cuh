// pfrx_device.cuh -- device-side tables and the cell-group kernel of the
// operator-split chemistry step (RStep/RReact, reaction.F90:3564-4055).
//
// Design (DESIGN.md has the long form):
//  * a GROUP of L lanes (L = 1,2,4,8,16,32; a power of two inside one warp)
//    owns one cell at a time.  Row i of the Newton system belongs to lane
//    i % L (R = ceil(N/L) rows per lane) and lives in REGISTERS with
//    compile-time indices; what a lane needs from its neighbours goes through
//    a small per-group shared-memory workspace guarded by __syncwarp(group
//    mask).  L = 1 is the thread-per-cell kernel.
//  * the kernel is a persistent STATE MACHINE: every pass of one loop is one
//    Newton iteration of whatever cell a group currently holds, so all groups
//    of a warp execute the same instructions (no divergence between cells
//    that need different iteration counts) and every routine exists exactly
//    once in the instruction stream (the first version inlined RTotal /
//    RKineticMineral several times, was 340 KB of SASS and fetch-bound).
//  * totals and d(total)/d(free) are sparse weighted sums of the secondary
//    molalities; the (entry, complex, weight) tasks are sorted by entry and
//    cut into L equal contiguous ranges on the host, so the lanes are balanced
//    even though H+ sits in 70 of the 88 Hanford complexes.  Entries cut by a
//    range boundary are finished by a fixed-order fix-up (deterministic).
//  * activity coefficients are evaluated once per (Z, a0) CLASS, and the
//    secondary molality is exp(lnQK - ln gamma): one exp per complex.
//  * LU is right-looking with implicit-scaled partial pivoting; the pivot row
//    is published through shared memory, the argmax over the group is three
//    REDUX (hi word, lo word, logical row) -- same pivot order and the same
//    per-element operation order as the Crout loops of utility.F90:597-688,
//    forward substitution fused into the elimination.
//  * state is cell-major SoA in HBM, read once at cell entry and written once
//    at exit; stoichiometry tables are read-only and L1/L2 resident.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pfrx.h"

#define PFRX_LOG_TO_LN 2.30258509299  // pflotran_constants.F90:84 (truncated there)
#define PFRX_IDEAL_GAS_CONSTANT 8.31446

// constraint of pfrx_equilibrate_constraint (include/pfrx.h): device copies of the tables
struct DevCons {
  int init_molality, max_iterations;
  const int *type;
  const double *Z;  // primary_spec_Z
  const double *eq_logK, *eq_logKcoef, *eq_h2o;
  const int *eq_ptr, *eq_spec;
  const double *eq_st;
  const double *conc;  // [naq][ld]
};

struct DevCfg {
  int naq, nim, n;
  int use_full_geochemistry, use_log, use_total_as_guess, use_isothermal;
  int act_freq, act_alg, use_act_h2o, h2o_aq_id;
  int max_its, max_cuts;
  double max_dlnC, tol_relchange, tol_res, tol_relres, min_sat;
  double debyeA, debyeB, debyeBdot;
  const double *pri_Z2;  // z_i^2 (ionic strength weights)
  const int *pri_cls;    // activity class of a primary species, -1 = neutral
  // complexes (CSR)
  int ncplx;
  const int *cx_ptr, *cx_id;
  const double *cx_st, *cx_h2o, *cx_logK, *cx_logKcoef, *cx_Z2;
  const int *cx_cls;
  // activity classes
  int ncls;
  const double *cls_negz2, *cls_a0;
  // balanced task lists for totals / dtotal (see header comment)
  int tk_per_lane;       // tasks per lane (padded to the same count)
  const int2 *tk_kd;     // [tk_per_lane*L] lane-interleaved: x = complex | (dst1+1)<<16, y = dst2+1
  const double *tk_w;    // weight of the task
  const unsigned *row_mask;  // [naq] bit j set when dtotal(i,j) has at least one task
  int nfix;
  const int *fx_ptr, *fx_dst, *fx_dst2, *fx_src;
  int nacc;              // accumulator slots: naq totals + naq(naq+1)/2 S + partials
  // kinetic minerals
  int nkin;
  const int *mn_ptr, *mn_id;
  const double *mn_st, *mn_h2o, *mn_logK, *mn_logKcoef, *mn_vol, *mn_rate, *mn_eact, *mn_thresh, *mn_limit;
  const int *mn_irrev;
  const int *me_ptr, *me_ij;   // per mineral: all (i,j) species pairs, i | j<<8
  const double *me_coef;       // nu_i * nu_j
  const double *mn_temkin, *mn_scale, *mn_power;
  // prefactors (reaction_mineral.F90:838-890), dense [(m*MAXP + p)*MAXS + s]; thread-per-cell kernel only
  const int *mn_npref, *mn_pref_nspec, *mn_pref_id;
  const double *mn_pref_alpha, *mn_pref_beta, *mn_pref_atten, *mn_pref_rate, *mn_pref_eact;
  // surface complexation
  int nsrfrxn, nsrfcplx;
  const int *sr_ptr, *sr_cx, *sr_type, *sr_surf, *sr_flag;
  const double *sr_dens;
  const int *sc_ptr, *sc_id;
  const double *sc_st, *sc_h2o, *sc_fs, *sc_logK, *sc_logKcoef;
  const int *se_ptr, *se_pp;   // per surface complex: all (p,p2) entry pairs, p | p2<<16
  int neqsr;
  // ion exchange, KD isotherms, dynamic KD (thread-per-cell kernel); nsorb = all equilibrium sorption
  int nsorb, nionx, nkd, ndynkd, ikd_units;
  const int *ix_ptr, *ix_cat, *ix_surf, *ix_zflag;
  const double *ix_k, *ix_cec, *pri_Z;
  const int *kd_spec, *kd_type, *kd_mnrl;
  const double *kd_coeff, *kd_lb, *kd_fn;
  const int *dk_spec, *dk_ref;
  const double *dk_refhigh, *dk_low, *dk_high, *dk_power;
  // general kinetic reactions, radioactive decay, immobile decay (thread-per-cell kernel)
  int ngen, nrd, nidc;
  int need_ds, off_ds;  // rt_auxvar%dtotal_sorb_eq as a matrix of its own (decay of a sorbing parent)
  const int *gn_ptr, *gn_id, *gn_fptr, *gn_fid, *gn_bptr, *gn_bid;
  const double *gn_st, *gn_fst, *gn_bst, *gn_kf, *gn_kr;
  const int *rd_ptr, *rd_id, *rd_fwd;
  const double *rd_st, *rd_kf;
  const int *idc_id;
  const double *idc_k;
  // Monod-type microbial reactions (thread-per-cell kernel)
  int nmb, mb_units;
  const int *mb_ptr, *mb_id, *mb_mptr, *mb_mid, *mb_hptr, *mb_hid, *mb_htype, *mb_bio;
  const double *mb_st, *mb_k, *mb_ea, *mb_mK, *mb_mC, *mb_hC, *mb_hC2, *mb_yield;
  int n_ixcat;  // cations of all ion-exchange reactions
  int off_ix;   // workspace: reference-cation sorbed concentrations, then cation concentrations
  const int *eqsr;
  int nmr;
  const int *mr_rxn, *mr_ptr;
  const double *mr_rate, *mr_frac;
  // CLM-CN
  // ELM-CN sandboxes (pfrx_sandbox.cuh): tables as in include/pfrx.h with device pointers
  int nsbx, sbx[PFRX_MAX_SANDBOXES];  // evaluation order, PFRX_SANDBOX_*
  int has_sd, has_nt, has_dn, has_pn, has_lg, has_cd, has_cs, elm;
  int need_dt;           // a sandbox reads d(total)/d(free): keep a copy next to the Jacobian
  int off_dt, off_nc, n_nc;
  pfrx_somdec sd;
  pfrx_nitrif nt;
  pfrx_denitr dn;
  pfrx_plantn pn;
  pfrx_langmuir lg;
  pfrx_cndegas cd;
  pfrx_calcite_sandbox cs;
  int has_rn;
  int elm_flow;  // elm_flow_coupled: SOMDECOMP's f_w from GetMoistureResponse
  pfrx_radon rn;
  // active gas species (reaction_gas.F90:87-174), CSR; thread-per-cell kernel only
  int ngas, off_tg, off_dg;
  const int *gs_ptr, *gs_id;
  const double *gs_st, *gs_h2o, *gs_logK, *gs_logKcoef;
  int cn_nrxn, cn_C, cn_N;
  const double *cn_CN, *cn_k, *cn_resp, *cn_inhib;
  const int *cn_nspec, *cn_cid, *cn_nid, *cn_up, *cn_down;
  // workspace layout (doubles, per group)
  int ws_stride, off_c, off_lnact, off_invc, off_x, off_xs, off_sec, off_lng, off_J, off_acc, off_tmp, off_sc,
      off_cls, off_mn, off_fs, off_mr, off_res, off_ts, js;
};

#include "pfrx_types.cuh"

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

// only the Temkin / scale-factor / affinity-power options of RKineticMineral
// use pow(); the code is cold for every deck that leaves them out.
static __device__ __forceinline__ double pfrx_pow(double x, double y) { return pow(x, y); }

enum { MODE_LOAD = 0, MODE_ITER = 1, MODE_DONE = 2 };

// ---------------------------------------------------------------------------
// N = padded system size (>= ncomp), L = lanes per cell.
template <int N, int L>
struct CellSolver {
  static constexpr int R = (N + L - 1) / L;
  const DevCfg &cfg;
  const DevState &st;
  Grp<L> g;
  double *ws;
  // ---- per-cell scalars (replicated on the lanes of the group) --------------
  int64_t cell;
  double den_kg, sat, temp, por, vol, spd, ln_act_h2o;
  bool dry;
  // RStep (reaction.F90:3564) bookkeeping
  double target, cumulative, dt;
  int ncuts, nconst, nss, nit, nku, ierr;
  bool had_cut, aborted;
  // RReact (reaction.F90:3742) bookkeeping
  int its;
  double norm0;
  // ---- per-row registers (row i = lane + r*L) --------------------------------
  double cval[R];     // current iterate (pri_molal | immobile)
  double guess[R];    // RStep guess
  double totcur[R];   // rt_auxvar%total (aq rows) / rt_auxvar%immobile (imm rows)
  double sorbcur[R];  // rt_auxvar%total_sorb_eq (aq rows)
  double lngam[R];    // ln(pri_act_coef) of aq rows
  double totnew[R];   // total(c) of the latest RTotal
  double sorbnew[R];
  double fixed[R], init_tot[R];
  double small_val[R];
  bool small[R];

  __device__ CellSolver(const DevCfg &c, const DevState &s, Grp<L> gg, double *w) : cfg(c), st(s), g(gg), ws(w) {}

  __device__ __forceinline__ int row(int r) const { return g.lane + r * L; }
  __device__ __forceinline__ double &W(int off, int i) { return ws[off + i]; }
  __device__ __forceinline__ double &Jm(int i, int j) { return ws[cfg.off_J + i * cfg.js + j]; }
  // logK at the cell temperature (RUpdateTempDependentCoefs, reaction.F90:5976)
  __device__ __forceinline__ double cx_logK(int k) const {
    return cfg.use_isothermal ? cfg.cx_logK[k] : interp_logK(cfg.cx_logKcoef + 5 * k, temp);
  }
  __device__ __forceinline__ double mn_logK(int m) const {
    return (cfg.use_isothermal || !cfg.mn_logKcoef) ? cfg.mn_logK[m] : interp_logK(cfg.mn_logKcoef + 5 * m, temp);
  }
  __device__ __forceinline__ double sc_logK(int k) const {
    return (cfg.use_isothermal || !cfg.sc_logKcoef) ? cfg.sc_logK[k] : interp_logK(cfg.sc_logKcoef + 5 * k, temp);
  }

  // ---- RActivityCoefficients, LAG branch (reaction.F90:4553-4612) ----------
  // one evaluation per (Z, a0) class; ws.c and ws.sec must be current.
  __device__ __forceinline__ void activity() {
    const int naq = cfg.naq, ncx = cfg.ncplx;
    double part = 0.0, msum = 0.0;
#pragma unroll 1
    for (int i = g.lane; i < naq; i += L) {
      double c = W(cfg.off_c, i);
      part += c * cfg.pri_Z2[i];
      if (i != cfg.h2o_aq_id) msum += c;
    }
#pragma unroll 1
    for (int k = g.lane; k < ncx; k += L) {
      double s = W(cfg.off_sec, k);
      part += s * cfg.cx_Z2[k];
      msum += s;
    }
    double I = 0.5 * g.sumd(part);
    double sq = sqrt(I);
#pragma unroll 1
    for (int q = g.lane; q < cfg.ncls; q += L)
      W(cfg.off_cls, q) = (cfg.cls_negz2[q] * sq * cfg.debyeA / (1.0 + cfg.cls_a0[q] * cfg.debyeB * sq) +
                           cfg.debyeBdot * I) * PFRX_LOG_TO_LN;
    if (cfg.use_act_h2o) {
      double t = 1.0 - 0.017 * g.sumd(msum);
      ln_act_h2o = t > 0.0 ? log(t) : 0.0;
    }
    g.sync();
#pragma unroll
    for (int r = 0; r < R; r++) {
      int i = row(r);
      if (i < naq) {
        int q = cfg.pri_cls[i];
        lngam[r] = q < 0 ? 0.0 : W(cfg.off_cls, q);
      }
    }
#pragma unroll 1
    for (int k = g.lane; k < ncx; k += L) {
      int q = cfg.cx_cls[k];
      W(cfg.off_lng, k) = q < 0 ? 0.0 : W(cfg.off_cls, q);
    }
  }

  // ---- RTotalSorbEqSurfCplx1 (reaction_surf_complex.F90:641-900) -----------
  // tot_sorb[r] += nu*S on the rows this lane owns; the derivative rows go into
  // the shared Jacobian scaled by jscale.  ws.fs[irxn] carries the free-site
  // concentration of the cell (rt_auxvar%srfcplxrxn_free_site_conc).
  __device__ __forceinline__ void surf_cplx1(int irxn, bool add_J, double jscale, bool store_conc) {
    const int naq = cfg.naq;
    const int r0 = cfg.sr_ptr[irxn], r1 = cfg.sr_ptr[irxn + 1];
    double fs = fmax(W(cfg.off_fs, irxn), 1.e-40);
    double dens;
    int ty = cfg.sr_type[irxn];
    if (ty == PFRX_MINERAL_SURFACE)
      dens = cfg.sr_dens[irxn] * st.mnrl_volfrac[cfg.sr_surf[irxn] * st.ld + cell];
    else if (ty == PFRX_ROCK_SURFACE)
      dens = cfg.sr_dens[irxn] * spd * (1.0 - por);
    else
      dens = cfg.sr_dens[irxn];
    g.sync();
    if (dens < 1.e-40) {
      if (g.lane == 0) {
        W(cfg.off_fs, irxn) = 0.0;
        if (store_conc)
#pragma unroll 1
          for (int q = r0; q < r1; q++) W(cfg.off_sc, cfg.sr_cx[q]) = 0.0;
      }
      g.sync();
      return;
    }
    // complex q of the reaction: ws.tmp[N+q] = lnQK without the free-site term
    double *base = ws + cfg.off_tmp + N;
#pragma unroll 1
    for (int q = r0 + g.lane; q < r1; q += L) {
      int k = cfg.sr_cx[q];
      double lnQK = -sc_logK(k) * PFRX_LOG_TO_LN;
      if (cfg.sc_h2o[k] != 0.0) lnQK += cfg.sc_h2o[k] * ln_act_h2o;
#pragma unroll 1
      for (int p = cfg.sc_ptr[k]; p < cfg.sc_ptr[k + 1]; p++) lnQK += cfg.sc_st[p] * W(cfg.off_lnact, cfg.sc_id[p]);
      base[q - r0] = lnQK;
    }
    g.sync();
    double *sconc = base + (r1 - r0);
    if (!cfg.sr_flag[irxn]) {
      // every free-site stoichiometry is 1: S_q = exp(lnQK_q) * Sx and
      // Sx = site density / (1 + sum_q exp(lnQK_q))   (closed form, :760-766)
      double e = 0.0;
#pragma unroll 1
      for (int q = r0 + g.lane; q < r1; q += L) {
        double v = exp(base[q - r0]);
        sconc[q - r0] = v;
        e += v;
      }
      e = g.sumd(e);
      fs = dens / (1.0 + e);
      g.sync();
#pragma unroll 1
      for (int q = r0 + g.lane; q < r1; q += L) sconc[q - r0] *= fs;
      g.sync();
    } else {
      // general stoichiometry: Newton on the free-site concentration, replicated
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
          if (g.lane == 0) sconc[q - r0] = sck;
          total += cfg.sc_fs[k] * sck;
        }
        g.sync();
        if (one_more) break;
        double res = dens - total, d = 1.0;
#pragma unroll 1
        for (int q = r0; q < r1; q++) d += cfg.sc_fs[cfg.sr_cx[q]] * sconc[q - r0] / fs;
        double dfs = res / d;
        if (it > 1000) damping = 0.5;
        fs = fs + damping * dfs;
        if (fabs(dfs / fs) < 1.e-12 || it > 100000) one_more = true;
        g.sync();
      }
    }
    if (g.lane == 0) W(cfg.off_fs, irxn) = fs;
    // dSx_dmi (eq. 2.3-46) and the sorbed totals.  Within one complex every
    // species appears once, so lanes stride over its species without conflicts;
    // complexes are processed one after the other.
    double denom = 0.0;
#pragma unroll 1
    for (int q = r0; q < r1; q++) {
      int k = cfg.sr_cx[q];
      denom += cfg.sc_fs[k] * cfg.sc_fs[k] * sconc[q - r0];
    }
    denom = denom / fs + 1.0;
#pragma unroll 1
    for (int i = g.lane; i < naq; i += L) W(cfg.off_tmp, i) = 0.0;
    g.sync();
#pragma unroll 1
    for (int q = r0; q < r1; q++) {
      int k = cfg.sr_cx[q];
      double Sk = sconc[q - r0], fk = cfg.sc_fs[k];
#pragma unroll 1
      for (int p = cfg.sc_ptr[k] + g.lane; p < cfg.sc_ptr[k + 1]; p += L) {
        int i = cfg.sc_id[p];
        double nu = cfg.sc_st[p];
        W(cfg.off_tmp, i) += nu * fk * Sk;
        W(cfg.off_ts, i) += nu * Sk;
      }
      g.sync();
    }
#pragma unroll 1
    for (int i = g.lane; i < naq; i += L) W(cfg.off_tmp, i) = (-W(cfg.off_tmp, i) / denom) * W(cfg.off_invc, i);
    if (store_conc && g.lane == 0)
      for (int q = r0; q < r1; q++) W(cfg.off_sc, cfg.sr_cx[q]) += sconc[q - r0];
    g.sync();
    if (add_J) {
#pragma unroll 1
      for (int q = r0; q < r1; q++) {
        int k = cfg.sr_cx[q];
        double Sk = sconc[q - r0];
        double nuiSx = cfg.sc_fs[k] * Sk / fs;
#pragma unroll 1
        for (int e = cfg.se_ptr[k] + g.lane; e < cfg.se_ptr[k + 1]; e += L) {
          int pp = cfg.se_pp[e];
          int p = pp & 0xffff, p2 = pp >> 16;
          int i = cfg.sc_id[p], j = cfg.sc_id[p2];
          double t = cfg.sc_st[p2] * Sk * W(cfg.off_invc, j) + nuiSx * W(cfg.off_tmp, j);
          Jm(i, j) += jscale * (cfg.sc_st[p] * t);
        }
        g.sync();
      }
    }
  }

  // ---- RTAuxVarCompute = RTotal (reaction.F90:4618-4759) + accumulation -----
  // Leaves ws.lnact/invc/sec, totnew/sorbnew, and the shared Jacobian
  // d(accumulation)/dc/dt (RTAccumulationDerivative, reaction.F90:5775).
  __device__ __forceinline__ void auxvar_compute() {
    const int naq = cfg.naq, n = cfg.n, ncx = cfg.ncplx;
#pragma unroll
    for (int r = 0; r < R; r++) {
      int i = row(r);
      if (i < naq) {
        W(cfg.off_lnact, i) = log(cval[r]) + lngam[r];
        W(cfg.off_invc, i) = 1.0 / cval[r];
      }
    }
    g.sync();
    // secondary species: lanes stride over complexes, one exp each
#pragma unroll 2
    for (int k = g.lane; k < ncx; k += L) {
      double lnQK = -cx_logK(k) * PFRX_LOG_TO_LN;
      double h = cfg.cx_h2o[k];
      if (h != 0.0) lnQK += h * ln_act_h2o;
#pragma unroll 1
      for (int p = cfg.cx_ptr[k]; p < cfg.cx_ptr[k + 1]; p++) lnQK += cfg.cx_st[p] * W(cfg.off_lnact, cfg.cx_id[p]);
      W(cfg.off_sec, k) = exp(lnQK - W(cfg.off_lng, k));
    }
    g.sync();
    // balanced weighted sums of sec: totals (-> ws.acc) and
    // S_ij = sum_k nu_ki nu_kj sec_k (-> both triangles of the shared Jacobian).
    // Every lane runs the same number of tasks (padded), tables are
    // lane-interleaved, a task with dst1 >= 0 closes its segment.  Offsets are
    // relative to ws.J; ws.acc follows it.
    double *jb = ws + cfg.off_J;
    {
      double acc = 0.0;
      const int tpl = cfg.tk_per_lane;
#pragma unroll 4
      for (int t = 0; t < tpl; t++) {
        const int ix = t * L + g.lane;
        const int2 kd = cfg.tk_kd[ix];
        acc = fma(cfg.tk_w[ix], W(cfg.off_sec, kd.x & 0xffff), acc);
        const int d1 = (kd.x >> 16) - 1;
        if (d1 >= 0) {
          jb[d1] = acc;
          if (kd.y > 0) jb[kd.y - 1] = acc;
          acc = 0.0;
        }
      }
    }
    g.sync();
#pragma unroll 1
    for (int f = g.lane; f < cfg.nfix; f += L) {
      double acc = 0.0;
#pragma unroll 1
      for (int p = cfg.fx_ptr[f]; p < cfg.fx_ptr[f + 1]; p++) acc += jb[cfg.fx_src[p]];
      jb[cfg.fx_dst[f]] = acc;
      if (cfg.fx_dst2[f] >= 0) jb[cfg.fx_dst2[f]] = acc;
    }
    g.sync();
    const double denL = den_kg * 1.e-3;
    const double psvd = por * sat * 1000.0 * vol / dt;
    const double *tot_acc = jb + N * cfg.js;
#pragma unroll
    for (int r = 0; r < R; r++) {
      int i = row(r);
      if (i < naq) {
        totnew[r] = (cval[r] + tot_acc[i]) * denL;
        double *jr = jb + i * cfg.js;
        if (dry) {
#pragma unroll 1
          for (int j = 0; j < n; j++) jr[j] = (i == j) ? 1.0 : 0.0;
        } else {
          const unsigned mask = cfg.row_mask[i];
          const double f = denL * psvd;
#pragma unroll 1
          for (int j = 0; j < naq; j++) {
            double sij = ((mask >> j) & 1u) ? jr[j] : 0.0;
            jr[j] = fma(sij, W(cfg.off_invc, j), (i == j ? 1.0 : 0.0)) * f;
          }
#pragma unroll 1
          for (int j = naq; j < n; j++) jr[j] = 0.0;
        }
      } else if (i < n) {
        totnew[r] = cval[r];
        double *jr = jb + i * cfg.js;
#pragma unroll 1
        for (int j = 0; j < n; j++) jr[j] = 0.0;
        jr[i] = dry ? 1.0 : vol / dt;
      }
      sorbnew[r] = 0.0;
    }
    // equilibrium sorption (RTotalSorb, reaction.F90:4783)
    if (cfg.neqsr > 0) {
      if (g.lane == 0)
#pragma unroll 1
        for (int k = 0; k < cfg.nsrfcplx; k++) W(cfg.off_sc, k) = 0.0;
#pragma unroll 1
      for (int i = g.lane; i < naq; i += L) W(cfg.off_ts, i) = 0.0;
#pragma unroll 1
      for (int e = 0; e < cfg.neqsr; e++) surf_cplx1(cfg.eqsr[e], true, vol / dt, true);
#pragma unroll
      for (int r = 0; r < R; r++)
        if (row(r) < naq) sorbnew[r] = W(cfg.off_ts, row(r));
    }
  }

  // ---- RKineticMineral (reaction_mineral.F90:647-1078), no prefactors ------
  // lane m evaluates mineral m and publishes {rate, Im, dIm/dQK * extras, QK};
  // row owners then add their residual / Jacobian entries.
  __device__ __forceinline__ void kinetic_mineral(bool apply) {
#pragma unroll 1
    for (int m = g.lane; m < cfg.nkin; m += L) {
      const int p0 = cfg.mn_ptr[m], p1 = cfg.mn_ptr[m + 1];
      double rate_vol = 0.0, Im = 0.0, dfac = 0.0;
      double lnQK = -mn_logK(m) * PFRX_LOG_TO_LN;
      if (cfg.mn_h2o[m] != 0.0) lnQK += cfg.mn_h2o[m] * ln_act_h2o;
#pragma unroll 1
      for (int p = p0; p < p1; p++) lnQK += cfg.mn_st[p] * W(cfg.off_lnact, cfg.mn_id[p]);
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
        double arr = 1.0;
        if (cfg.mn_eact[m] > 0.0)
          arr = exp(cfg.mn_eact[m] / PFRX_IDEAL_GAS_CONSTANT * (1.0 / (25.0 + 273.15) - 1.0 / (temp + 273.15)));
        double spr = cfg.mn_rate[m] * arr;
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
        // Jac(i,j) += nu_i * dfac * nu_j / c_j   with everything else folded in
        dfac = dIm_dQK * QK * (den_kg * 1.e-3);
        if (lim > 0.0) {
          double den = 1.0 + (1.0 - aff) / lim;
          dfac = dIm_dQK * (1.0 + QK / lim / den) * QK * (den_kg * 1.e-3) / den;
        }
      }
      W(cfg.off_mn, 3 * m + 0) = rate_vol;
      W(cfg.off_mn, 3 * m + 1) = Im;
      W(cfg.off_mn, 3 * m + 2) = dfac;
    }
    g.sync();
    if (!apply) return;
    // within one mineral every (i,j) pair is distinct: lanes stride over the
    // pairs; minerals are applied one after the other (reference order)
#pragma unroll 1
    for (int m = 0; m < cfg.nkin; m++) {
      const double Im = W(cfg.off_mn, 3 * m + 1);
      const double df = W(cfg.off_mn, 3 * m + 2);
      if (Im == 0.0 && df == 0.0) continue;
#pragma unroll 1
      for (int p = cfg.mn_ptr[m] + g.lane; p < cfg.mn_ptr[m + 1]; p += L)
        W(cfg.off_res, cfg.mn_id[p]) += cfg.mn_st[p] * Im;
#pragma unroll 1
      for (int e = cfg.me_ptr[m] + g.lane; e < cfg.me_ptr[m + 1]; e += L) {
        int ij = cfg.me_ij[e];
        int i = ij & 0xff, j = ij >> 8;
        Jm(i, j) += (cfg.me_coef[e] * df) * W(cfg.off_invc, j);
      }
      g.sync();
    }
  }

  // ---- RMultiRateSorption (reaction_surf_complex.F90:552-637) --------------
  // Res_i += V * sum_k kk_k (f_k Seq_i - S_ki) = V * (A Seq_i - B_i) with
  // A = sum_k kk_k f_k and B_i = sum_k kk_k S_ki (B is fixed within a sub-step).
  __device__ __forceinline__ void multirate() {
    const int naq = cfg.naq;
#pragma unroll 1
    for (int q = 0; q < cfg.nmr; q++) {
      int irxn = cfg.mr_rxn[q];
      int r0 = cfg.mr_ptr[q], r1 = cfg.mr_ptr[q + 1];
      double A = 0.0;
#pragma unroll 1
      for (int k = r0; k < r1; k++) A += cfg.mr_rate[k] / (1.0 + cfg.mr_rate[k] * dt) * cfg.mr_frac[k];
#pragma unroll 1
      for (int i = g.lane; i < naq; i += L) W(cfg.off_ts, i) = 0.0;
      surf_cplx1(irxn, true, vol * A, false);
#pragma unroll 1
      for (int i = g.lane; i < naq; i += L) {
        double seq = W(cfg.off_ts, i);
        W(cfg.off_res, i) += vol * (A * seq - W(cfg.off_mr, (2 * q + 1) * N + i));
        W(cfg.off_mr, (2 * q) * N + i) = seq;  // kinmr_total_sorb(:,0,q): the equilibrium target
      }
    }
  }

  // B_i of every multirate reaction for the sub-step that starts now
  __device__ __forceinline__ void multirate_begin() {
    const int naq = cfg.naq;
#pragma unroll 1
    for (int q = 0; q < cfg.nmr; q++) {
      int r0 = cfg.mr_ptr[q], r1 = cfg.mr_ptr[q + 1];
      int64_t base = (int64_t)naq * (r0 + q);
#pragma unroll 1
      for (int i = g.lane; i < naq; i += L) {
        double B = 0.0;
#pragma unroll 1
        for (int k = r0; k < r1; k++) {
          double kk = cfg.mr_rate[k] / (1.0 + cfg.mr_rate[k] * dt);
          B += kk * st.kinmr[(base + (int64_t)naq * (k - r0 + 1) + i) * st.ld + cell];
        }
        W(cfg.off_mr, (2 * q + 1) * N + i) = B;
      }
    }
  }

  // ---- CLM_CN_React (reaction_sandbox_clm_cn.F90:468-787) ------------------
  __device__ __forceinline__ void clm_cn() {
    const int off = cfg.naq;
    double temp_K = temp + 273.15;
    if (!(temp_K > 227.15)) return;
    const double one_over_71_02 = 1.408054069e-2, theta_min = 0.01, one_over_log_theta_min = -2.17147241e-1;
    double F_t = exp(308.56 * (one_over_71_02 - 1.0 / (temp_K - 227.13)));
    double F_theta = log(theta_min / fmax(theta_min, sat)) * one_over_log_theta_min;
    double cinh = F_t * F_theta;
    const int ires_C = off + cfg.cn_C, ispec_N = cfg.cn_N, ires_N = off + ispec_N;
    const double *imm = ws + cfg.off_c + off;  // immobile(:) of the current iterate
    auto addR = [&](int irow, double v) {
      if ((irow % L) == g.lane) W(cfg.off_res, irow) += v;
    };
    auto addJ = [&](int irow, int jcol, double v) {
      if ((irow % L) == g.lane) Jm(irow, jcol) += v;
    };
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
      addR(ires_C, -(sC * rate));
      addR(ires_N, -(sN * rate));
      addR(rUC, -((-1.0) * sUC * rate));
      if (!constCN) addR(rUN, -((-1.0) * sUN * rate));
      if (id >= 0) addR(rD, -(sDC * rate));
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

  // ---- RSolve + LUDecomposition + LUBackSubstitution -----------------------
  // (reaction.F90:5457-5516, utility.F90:597-735).  a[][] rows in registers.
  // Returns false when a row is all zero (singular) -- the group agrees.
  __device__ __forceinline__ bool solve(double (&a)[R][N], double (&b)[R], double *xout /* ws, n entries */) {
    const int n = cfg.n;
    double vv[R], dinv[R];
    int pos[R], step[R];
    bool done[R];
    bool bad = false;
#pragma unroll
    for (int r = 0; r < R; r++) {
      int i = row(r);
      pos[r] = i;
      step[r] = -1;
      dinv[r] = 0.0;
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
            dinv[r] = 1.0 / p;
            buf[k] = dinv[r];
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
          if (step[r] == s) xout[s] = b[r] * dinv[r];  // B(i)=sum/A(i,i), reciprocal kept from the elimination
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

  // ---- start of RReact for a sub-step of length dt (reaction.F90:3826-3850) --
  __device__ __forceinline__ void react_begin() {
    const int naq = cfg.naq, n = cfg.n;
    const double psv = por * sat * 1000.0 * vol;
    dry = sat < cfg.min_sat;
#pragma unroll
    for (int r = 0; r < R; r++) {
      int i = row(r);
      double f = 0.0;
      if (!dry) {
        if (i < naq)
          f = psv * totcur[r];
        else if (i < n)
          f = 0.0 + totcur[r] * vol;
      }
      if (cfg.neqsr > 0 && i < naq) f = f + sorbcur[r] * vol;
      fixed[r] = f;
      init_tot[r] = totcur[r];
      cval[r] = guess[r];
    }
    its = 0;
    norm0 = 0.0;
    if (cfg.nmr > 0) multirate_begin();
  }

  // ---- load a cell and start RStep (reaction.F90:3564-3650) -----------------
  __device__ __forceinline__ void load_cell(int64_t c) {
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
    had_cut = aborted = false;
    cumulative = 0.0;
    dt = target;
    ncuts = nconst = 0;
#pragma unroll
    for (int r = 0; r < R; r++) {
      int i = row(r);
      small[r] = false;
      small_val[r] = 0.0;
      lngam[r] = 0.0;
      sorbcur[r] = 0.0;
      totcur[r] = 0.0;
      guess[r] = 1.0;
      cval[r] = 1.0;
            if (i < naq) {
        totcur[r] = st.total[i * ld + c];
        guess[r] = st.pri_molal[i * ld + c];
        lngam[r] = log(st.pri_act_coef[i * ld + c]);
        if (cfg.neqsr > 0) sorbcur[r] = st.total_sorb_eq[i * ld + c];
      } else if (i < n) {
        totcur[r] = st.immobile[(i - naq) * ld + c];
        guess[r] = totcur[r];  // read before the 1e-40 clamp (pmc_subsurface_osrt.F90:356-362)
      }
    }
    g.sync();  // previous cell's workspace readers are done
#pragma unroll 1
    for (int k = g.lane; k < ncx; k += L) {
      W(cfg.off_sec, k) = st.sec_molal[k * ld + c];
      if (cfg.act_freq != PFRX_ACT_COEF_FREQUENCY_NEWTON_ITER) W(cfg.off_lng, k) = log(st.sec_act_coef[k * ld + c]);
    }
#pragma unroll 1
    for (int k = g.lane; k < cfg.nsrfrxn; k += L) W(cfg.off_fs, k) = st.free_site[k * ld + c];
#pragma unroll 1
    for (int k = g.lane; k < cfg.nsrfcplx; k += L) W(cfg.off_sc, k) = 0.0;
#pragma unroll 1
    for (int k = g.lane; k < cfg.nkin; k += L) W(cfg.off_mn, 3 * k) = st.mnrl_rate[k * ld + c];
#pragma unroll 1
    for (int q = 0; q < cfg.nmr; q++) {
      int64_t base = (int64_t)naq * (cfg.mr_ptr[q] + q);
#pragma unroll 1
      for (int i = g.lane; i < naq; i += L) W(cfg.off_mr, (2 * q) * N + i) = st.kinmr[(base + i) * ld + c];
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
      int i = row(r);
      if (i < n && totcur[r] <= 1.e-40) {
        small[r] = true;
        small_val[r] = totcur[r];
        totcur[r] = 1.e-40;
      }
      if (cfg.use_total_as_guess && i < naq) guess[r] = totcur[r];
    }
    g.sync();
  }

  // ---- write the cell back (rt_auxvar of the cell) ----------------------------
  __device__ __forceinline__ void store_cell() {
    const int naq = cfg.naq, n = cfg.n, ncx = cfg.ncplx;
    const int64_t ld = st.ld, c = cell;
    const bool act_upd = cfg.act_freq == PFRX_ACT_COEF_FREQUENCY_NEWTON_ITER;
#pragma unroll
    for (int r = 0; r < R; r++) {
      int i = row(r);
      if (i < naq) {
        st.total[i * ld + c] = (small[r] && !aborted) ? small_val[r] : totcur[r];
        st.pri_molal[i * ld + c] = cval[r];
        if (act_upd) st.pri_act_coef[i * ld + c] = exp(lngam[r]);
        if (cfg.neqsr > 0) st.total_sorb_eq[i * ld + c] = sorbcur[r];
      } else if (i < n) {
        st.immobile[(i - naq) * ld + c] = (small[r] && !aborted) ? small_val[r] : totcur[r];
      }
    }
    g.sync();
#pragma unroll 1
    for (int k = g.lane; k < ncx; k += L) {
      st.sec_molal[k * ld + c] = W(cfg.off_sec, k);
      if (act_upd) st.sec_act_coef[k * ld + c] = exp(W(cfg.off_lng, k));
    }
#pragma unroll 1
    for (int k = g.lane; k < cfg.nsrfrxn; k += L) st.free_site[k * ld + c] = W(cfg.off_fs, k);
    if (cfg.neqsr > 0 && st.eqsrfcplx_conc)
#pragma unroll 1
      for (int k = g.lane; k < cfg.nsrfcplx; k += L) st.eqsrfcplx_conc[k * ld + c] = W(cfg.off_sc, k);
#pragma unroll 1
    for (int k = g.lane; k < cfg.nkin; k += L) st.mnrl_rate[k * ld + c] = W(cfg.off_mn, 3 * k);
#pragma unroll 1
    for (int q = 0; q < cfg.nmr; q++) {
      int64_t base = (int64_t)naq * (cfg.mr_ptr[q] + q);
#pragma unroll 1
      for (int i = g.lane; i < naq; i += L) st.kinmr[(base + i) * ld + c] = W(cfg.off_mr, (2 * q) * N + i);
    }
    if (g.lane == 0) {
      if (st.ln_act_h2o && cfg.use_act_h2o) st.ln_act_h2o[c] = ln_act_h2o;
      st.num_sub_steps[c] = nss;
      st.num_iterations[c] = nit;
      st.num_kinetic_state_updates[c] = nku;
      st.ierror[c] = ierr;
    }
  }

  // ---- RReact failed (its > max or singular): RStep cuts dt (reaction.F90:3660) -
  // returns true when the cell is finished (too many cuts)
  __device__ __forceinline__ bool cut_or_abort() {
    nit += its;
    ncuts++;
    had_cut = true;
    if (ncuts > cfg.max_cuts) {
      ierr = 1;
      aborted = true;
      return true;
    }
    dt = 0.5 * dt;
    nconst = 0;
    react_begin();
    return false;
  }

  // ---- converged sub-step: RUpdateKineticState + RStep accounting ------------
  // (reaction.F90:3681-3716, :5935-5972, reaction_mineral.F90:1456,
  //  reaction_surf_complex.F90:1107).  Returns true when the cell is finished.
  __device__ __forceinline__ bool finish_substep() {
    const int naq = cfg.naq;
    nit += its;
#pragma unroll
    for (int r = 0; r < R; r++) {
      totcur[r] = totnew[r];
      sorbcur[r] = sorbnew[r];
    }
    bool upd = false;
    if (cfg.nkin > 0) {
      upd = true;
      // the rates of the converged iterate are in ws.mn (same inputs as the
      // RKineticMineral call MineralUpdateKineticState makes)
#pragma unroll 1
      for (int m = g.lane; m < cfg.nkin; m += L) {
        double vf = st.mnrl_volfrac[m * st.ld + cell] + W(cfg.off_mn, 3 * m) * cfg.mn_vol[m] * dt;
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
      for (int i = g.lane; i < naq; i += L) {
        double seq = W(cfg.off_mr, (2 * q) * N + i);
#pragma unroll 1
        for (int k = r0; k < r1; k++) {
          double kdt = cfg.mr_rate[k] * dt;
          int64_t ix = (base + (int64_t)naq * (k - r0 + 1) + i) * st.ld + cell;
          st.kinmr[ix] = (st.kinmr[ix] + kdt * cfg.mr_frac[k] * seq) / (1.0 + kdt);
        }
      }
    }
    if (cfg.cn_nrxn > 0) upd = true;  // any sandbox => true (reaction.F90:5965)
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
    if (cumulative >= target) return true;
    g.sync();
    react_begin();
    return false;
  }

  // ---- one Newton iteration of RReact (reaction.F90:3853-4048) ---------------
  // returns true when the cell is finished (stored by the caller)
  __device__ __forceinline__ bool newton_pass() {
    const int naq = cfg.naq, n = cfg.n;
    its++;
#pragma unroll
    for (int r = 0; r < R; r++)
      if (row(r) < n) W(cfg.off_c, row(r)) = cval[r];
    g.sync();
    if (cfg.act_freq == PFRX_ACT_COEF_FREQUENCY_NEWTON_ITER) activity();
    auxvar_compute();
    if (its > cfg.max_its) {
#pragma unroll
      for (int r = 0; r < R; r++) {
        totcur[r] = init_tot[r];  // total and immobile restored (reaction.F90:3891-3894)
        sorbcur[r] = sorbnew[r];  // total_sorb_eq is not
        if (row(r) >= naq) cval[r] = init_tot[r];
      }
      return cut_or_abort();
    }
    const double psv = por * sat * 1000.0 * vol;
#pragma unroll
    for (int r = 0; r < R; r++) {
      int i = row(r);
      double a = 0.0;
      if (!dry) {
        if (i < naq)
          a = psv * totnew[r];
        else if (i < n)
          a = 0.0 + cval[r] * vol;
      }
      if (cfg.neqsr > 0 && i < naq) a = a + sorbnew[r] * vol;
      if (i < n) W(cfg.off_res, i) = (a - fixed[r]) / dt;
    }
    // RReaction (reaction.F90:4059-4130), same order
    if (cfg.nkin > 0) kinetic_mineral(!dry);
    if (!dry) {
      if (cfg.nmr > 0) multirate();
      if (cfg.cn_nrxn > 0) clm_cn();
    }
    g.sync();
    double res[R];
#pragma unroll
    for (int r = 0; r < R; r++) res[r] = row(r) < n ? W(cfg.off_res, row(r)) : 0.0;
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
    bool conv = (mabs < cfg.tol_res) || (rel < cfg.tol_relres);
    if (!conv) {
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
        // solve_error branch: no restore (reaction.F90:3964-3967)
#pragma unroll
        for (int r = 0; r < R; r++) {
          totcur[r] = totnew[r];
          sorbcur[r] = sorbnew[r];
        }
        return cut_or_abort();
      }
      double cnew[R], maxrel = 0.0;
      bool anyv = false;
      double minr = 1.e20;
      if (!cfg.use_log) {
#pragma unroll
        for (int r = 0; r < R; r++) {
          int i = row(r);
          if (i < n) {
            double u = xs[i];
            if (cval[r] <= u) minr = fmin(minr, fabs(cval[r] / u));
          }
        }
        minr = g.mind(minr);
      }
#pragma unroll
      for (int r = 0; r < R; r++) {
        int i = row(r);
        if (i < n) {
          double u = xs[i];
          if (cfg.use_log) {
            u = copysign(1.0, u) * fmin(fabs(u), cfg.max_dlnC);
            cnew[r] = cval[r] * exp(-u);
          } else {
            if (minr < 1.0) u = u * minr * 0.99;
            cnew[r] = cval[r] - u;
          }
          double v = fabs((cnew[r] - cval[r]) / cval[r]);
          if (!isnan(v)) {  // NaN-skipping maxval, like gfortran's MAXVAL
            maxrel = anyv ? fmax(maxrel, v) : v;
            anyv = true;
          }
        }
      }
      double mr = g.maxd(anyv ? maxrel : -1.0);
      conv = (mr >= 0.0) && (mr < cfg.tol_relchange);
      if (!conv) {
#pragma unroll
        for (int r = 0; r < R; r++)
          if (row(r) < n) cval[r] = cnew[r];
        return false;
      }
    }
    // converged: the reference's "one last update" (reaction.F90:4052) would
    // recompute RTotal at the same c -- identical values, already in hand.
    return finish_substep();
  }
};

template <int N, int L>
__global__ void __launch_bounds__(128, (L >= 16 ? 4 : (L >= 8 ? 3 : 2))) pfrx_rstep_kernel(DevCfg cfg, DevState st, int64_t ncell, double tran_dt,
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
#pragma unroll 1
  for (int i = g.lane; i < cfg.ws_stride; i += L) ws[i] = 0.0;
  g.sync();
  CellSolver<N, L> sol(cfg, st, g, ws);
  sol.target = tran_dt;

  unsigned long long l_active = 0, l_its = 0, l_cut = 0;
  long long l_first = -1;
  int l_maxits = 0, l_maxkin = 0, l_maxerr = 0, l_maxsub = 0;

  const int64_t gwarp = (int64_t)blockIdx.x * warps_per_block + warp_in_block;
  const int64_t nwarps = (int64_t)gridDim.x * warps_per_block;
  int64_t next = gwarp * CPW + lane32 / L;  // this group's next cell
  const int64_t stride = nwarps * CPW;
  int mode = MODE_LOAD;

  for (;;) {
    if (!__any_sync(0xffffffffu, mode != MODE_DONE)) break;
    if (mode == MODE_LOAD) {
      // skip inactive cells (pmc_subsurface_osrt.F90:351)
      while (next < ncell && st.imat && st.imat[next] <= 0) {
        if (g.lane == 0) {
          st.num_sub_steps[next] = 0;
          st.num_iterations[next] = 0;
          st.num_kinetic_state_updates[next] = 0;
          st.ierror[next] = 0;
        }
        next += stride;
      }
      if (next >= ncell) {
        mode = MODE_DONE;
      } else {
        sol.load_cell(next);
        next += stride;
        if (!cfg.use_full_geochemistry) {
          // RStep early-out (reaction.F90:3604-3608)
#pragma unroll
          for (int r = 0; r < sol.R; r++)
            if (sol.row(r) < cfg.naq)
              st.pri_molal[sol.row(r) * st.ld + sol.cell] = sol.totcur[r] / sol.den_kg * 1.e3;
          if (g.lane == 0) {
            st.num_sub_steps[sol.cell] = 0;
            st.num_iterations[sol.cell] = 0;
            st.num_kinetic_state_updates[sol.cell] = 0;
            st.ierror[sol.cell] = 0;
            l_active++;
          }
        } else {
          sol.react_begin();
          mode = MODE_ITER;
        }
      }
    }
    if (mode == MODE_ITER) {
      if (sol.newton_pass()) {
        sol.store_cell();
        if (g.lane == 0) {
          l_active++;
          l_its += (unsigned long long)sol.nit;
          if (sol.had_cut) l_cut++;
          if (sol.ierr != 0 && (l_first < 0 || sol.cell < l_first)) l_first = sol.cell;
          l_maxits = max(l_maxits, sol.nit);
          l_maxkin = max(l_maxkin, sol.nku);
          l_maxerr = max(l_maxerr, sol.ierr);
          l_maxsub = max(l_maxsub, sol.nss);
        }
        mode = MODE_LOAD;
      }
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
