// pfrx_spec.cuh -- network-SPECIALISED thread-per-cell kernel (template part).
//
// specialize.py generates, for one reaction network, a .cu file that defines
//   SPEC_N, SPEC_NAQ, SPEC_NC, SPEC_NCX, SPEC_NCLS, SPEC_NKIN, SPEC_NSRFRXN,
//   SPEC_NSRFCPLX, SPEC_NEQSR, SPEC_USE_LOG, SPEC_ACT_UPD, SPEC_USE_ACT_H2O, SPEC_SIG,
//   SPEC_THREADS, SPEC_MINBLOCKS, the index maps spec_cmap() / spec_sp_of() and the
//   straight-line device functions
//   spec_activity(), spec_rtotal(), spec_minerals(), spec_sorption()
// with every stoichiometric coefficient, logK, charge and index as an immediate,
// and includes this file, which holds what does not depend on the network:
// RStep / RReact control flow (reaction.F90:3564-4055), the unrolled LU
// (utility.F90:597-735) and the launch skeleton.
//
// One thread owns one cell.  c, ln a, 1/c, totals, residual, ln gamma are
// register arrays (all indices are literals after unrolling).  Shared memory
// holds, per thread, the Jacobian of the COUPLED species (those that occur in
// some reaction; NC of them) as NC rows of NC+1 doubles -- the extra column is
// the row's implicit-scaling factor during the decomposition and the right-hand
// side afterwards -- and the fixed accumulation / guess vectors (+ ln gamma of
// the complexes when activity coefficients are frozen).  Species that occur in
// no reaction (tracers, immobile species without a sandbox) have a diagonal
// Jacobian row and column: their update is res/J, bit-identical to what the
// reference's full LU returns for them, and they stay out of the matrix.
//
// Element e of a thread's slice is at slice[e * 32 + lane]: any per-lane row
// permutation of the LU stays bank-conflict free.
#pragma once
#include <cuda_runtime.h>

#include "../../include/pfrx.h"
#include "pfrx_fastmath.cuh"
#include "pfrx_types.cuh"
#ifndef PFRX_IDEAL_GAS_CONSTANT
#define PFRX_IDEAL_GAS_CONSTANT 8.31446
#endif
#include "pfrx_sandbox.cuh"  // response functions shared with the generic kernels

#ifndef SPEC_NKIN3
#define SPEC_NKIN3 0  // general / radioactive-decay / immobile-decay / microbial reactions (spec_kinetic)
#endif
#ifndef SPEC_NDTP
#define SPEC_NDTP 0  // entries of d(total)/d(free) that RRadioactiveDecay reads
#endif
#ifndef SPEC_NKC
#define SPEC_NKC 0  // per-cell constants of the generated code (spec_cell_constants)
#endif
#ifndef SPEC_NDSP
#define SPEC_NDSP 0  // entries of d(total_sorb_eq)/d(free) that RRadioactiveDecay reads
#endif
#ifndef SPEC_NIONX
#define SPEC_NIONX 0  // ion-exchange reactions
#define SPEC_NIXCAT 0
#endif
#ifndef SPEC_NSORB
#define SPEC_NSORB SPEC_NEQSR  // equilibrium sorption of any kind (surface complexation, ion exchange, KD)
#endif
#ifndef SPEC_NSBX
#define SPEC_NSBX 0  // reaction sandboxes of any kind
#endif
#ifndef SPEC_NMR
#define SPEC_NMR 0     // multirate kinetic surface-complexation reactions (one-warp skeleton "s" only)
#endif
#ifndef SPEC_REFILL
#define SPEC_REFILL 0  // lock-step skeleton: finished lanes fetch the next cell (variant "q")
#endif
#ifndef SPEC_NNC
#define SPEC_NNC 0   // persisted N:C ratios of the SOMDECOMP sandbox (pfrx_state.somdec_nc)
#endif
#ifndef SPEC_ELM
#define SPEC_ELM 0   // ELM_PFLOTRAN build: per-cell ELM scalars
#endif

// SPEC_FASTMATH 1: branch-free exp / log / division (pfrx_fastmath.cuh) in the hot
// places whose arguments are known to be normal numbers; 0: CUDA's own everywhere
// SPEC_LOOP_LU 1: the dense solve as rolled loops over shared memory (~400 instructions
// that stay in the instruction cache); 0: fully unrolled Crout with the current column in
// registers (~5 500 straight-line instructions).  Measured: unrolled wins for C3 / C4 / C2,
// rolled wins for C5 (profiles/r01_spec_variants.md).
#ifndef SPEC_LOOP_LU
#define SPEC_LOOP_LU 0
#endif
#ifndef SPEC_FASTMATH
#define SPEC_FASTMATH 1
#endif
__device__ __forceinline__ double sx_exp(double x) { return SPEC_FASTMATH ? pfrx_exp(x) : exp(x); }
__device__ __forceinline__ double sx_log(double x) { return SPEC_FASTMATH ? pfrx_log(x) : log(x); }
__device__ __forceinline__ double sx_rcp(double x) { return SPEC_FASTMATH ? pfrx_rcp(x) : 1.0 / x; }
__device__ __forceinline__ double sx_div(double a, double b) { return SPEC_FASTMATH ? pfrx_div(a, b) : a / b; }
// FuncMonod and its derivative (elm_rspfuncs.F90:321-338) for the generated sandbox code
__device__ __forceinline__ double sx_monod(double c, double k) { return sx_div(c, c + k); }
__device__ __forceinline__ double sx_dmonod(double c, double k) { return sx_div(sx_div(k, c + k), c + k); }

// Linear formulation, the update's scaling factor (reaction.F90:4010-4023): min over the components with
// c_i <= u_i of |c_i / u_i|, starting from 1e20.  The reference divides inside a branch per component; here the
// smallest ratio is found by cross-multiplication (|c_i| |u_b| < |c_b| |u_i|, all factors non-negative) and ONE
// division of the selected pair follows -- the same quotient the reference forms for that pair.  u_i = 0 never
// wins (the reference's Inf / NaN does not either).
template <int NN>
__device__ __forceinline__ double spec_min_ratio(const double (&c)[NN], const double (&u)[NN]) {
  double bc = 1.e20, bu = 1.0;
#pragma unroll
  for (int i = 0; i < NN; i++) {
    const double ac = fabs(c[i]), au = fabs(u[i]);
    const bool cand = (c[i] <= u[i]) && (ac * bu < bc * au);
    bc = cand ? ac : bc;
    bu = cand ? au : bu;
  }
  return bc / bu;
}

#define SPEC_LN 2.30258509299  // pflotran_constants.F90:84 (truncated there)

// shared-memory slots of a thread (doubles)
#ifndef SPEC_LOOP_LU
#define SPEC_LOOP_LU 0
#endif
#if SPEC_LOOP_LU
#define SPEC_JS (SPEC_NC + 2)  // + right-hand side column + scaling-factor / solution column
#else
#define SPEC_JS (SPEC_NC + 1)  // + one column: scaling factor, then right-hand side
#endif
#if SPEC_LOOP_LU
#define SPEC_OFF_C (SPEC_NC * SPEC_JS)
#define SPEC_OFF_LNGSEC (SPEC_OFF_C + SPEC_NC)
#define SPEC_FIXED(i) fixed[i]  // no slots left: registers / local memory
#define SPEC_SMALL(i) small_val[i]
#else
#define SPEC_OFF_FIXED (SPEC_NC * SPEC_JS)
#define SPEC_OFF_SMALL (SPEC_OFF_FIXED + SPEC_N)
#define SPEC_OFF_LNGSEC (SPEC_OFF_SMALL + SPEC_N)
#define SPEC_FIXED(i) SW(SPEC_OFF_FIXED + (i))  // read once per iteration, constant over a sub-step
#define SPEC_SMALL(i) SW(SPEC_OFF_SMALL + (i))  // totals <= 1e-40 to put back at the end (rare)
#endif
#ifndef SPEC_NRO
#define SPEC_NRO 0      // Jacobian entries of the row-only species (kept outside the dense matrix)
#define SPEC_NROSPEC 0
#endif
#define SPEC_OFF_RO (SPEC_OFF_LNGSEC + (SPEC_ACT_UPD ? 0 : SPEC_NCX))
#define SPEC_SLOTS (SPEC_OFF_RO + SPEC_NRO)
// slot -> index relative to the thread's base pointer
#define SW(e) W[(e) * 32]
#define JX(ci, cj) (((ci) * SPEC_JS + (cj)) * 32)

extern "C" {
__device__ const unsigned long long pfrx_spec_sig = SPEC_SIG;
// {N, shared doubles per block, threads per block, min blocks per SM, cells per block}
__device__ const int pfrx_spec_info[5] = {SPEC_N, SPEC_SLOTS * SPEC_THREADS, SPEC_THREADS, SPEC_MINBLOCKS, SPEC_THREADS};
// bit 0: the refill skeleton, which hands out cells through DevState.order
__device__ const int pfrx_spec_flags = SPEC_REFILL ? 1 : 0;
}

struct SpecCell {
  // per-cell scalars
  double den_kg, sat, temp, por, vol, spd, ln_act_h2o;
  double Isec, msec;  // sum z^2 m, sum m over secondary species of the latest RTotal
  double lgcls[SPEC_NCLS > 0 ? SPEC_NCLS : 1];
  double lngam[SPEC_ACT_UPD ? 1 : SPEC_N];  // frozen coefficients only; otherwise ln gamma_i is lgcls[class of i]
  double fsite[SPEC_NSRFRXN > 0 ? SPEC_NSRFRXN : 1];
  double scconc[SPEC_NSRFCPLX > 0 ? SPEC_NSRFCPLX : 1];
  double mrate[SPEC_NKIN > 0 ? SPEC_NKIN : 1];
  // multirate sorption (RMultiRateSorption): per reaction the equilibrium target S_eq (what
  // kinmr_total_sorb(:,0,irxn) holds), B = sum_k k_k/(1+k_k dt) S_k of the sub-step, A = sum_k k_k f_k/(1+k_k dt)
  double mr_seq[SPEC_NMR > 0 ? SPEC_NMR * SPEC_NAQ : 1], mr_B[SPEC_NMR > 0 ? SPEC_NMR * SPEC_NAQ : 1];
  double mr_A[SPEC_NMR > 0 ? SPEC_NMR : 1];
  double nc[SPEC_NNC > 0 ? SPEC_NNC : 1];  // N:C ratios that persist between evaluations
  double dtp[SPEC_NDTP > 0 ? SPEC_NDTP : 1];  // rt_auxvar%aqueous%dtotal(parent, j) of the latest RTotal
  double dsp[SPEC_NDSP > 0 ? SPEC_NDSP : 1];  // rt_auxvar%dtotal_sorb_eq(parent, j)
  double kc[SPEC_NKC > 0 ? SPEC_NKC : 1];     // sub-expressions of per-cell scalars, evaluated once per cell
  double ixref[SPEC_NIONX > 0 ? SPEC_NIONX : 1];    // eqionx_ref_cation_sorbed_conc (guess of the next evaluation)
  double ixconc[SPEC_NIXCAT > 0 ? SPEC_NIXCAT : 1];  // eqionx_conc
  double elm_w, elm_o, elm_t, elm_zsoil, elm_kscalar, elm_bd_dry, elm_bsw, elm_plantndemand;
  double elm_sucsat, elm_watfc, elm_effpor;  // GetMoistureResponse (flow-coupled ELM build)
  bool dry;
  bool store;  // false: the lane has finished its cell, rt_auxvar%sec_molal must not be touched
};

// ln gamma of primary species i (literal i)
#define SPEC_LNGAM(s, i) \
  (SPEC_ACT_UPD ? (spec_pri_cls(i) < 0 ? 0.0 : (s).lgcls[spec_pri_cls(i) < 0 ? 0 : spec_pri_cls(i)]) : (s).lngam[SPEC_ACT_UPD ? 0 : (i)])

// ---- generated for the network (declared here, defined by the generator) ------
__device__ __forceinline__ void spec_activity(const double (&c)[SPEC_N], SpecCell &s);
__device__ __forceinline__ void spec_rtotal(const double (&c)[SPEC_N], double (&lna)[SPEC_N], double (&ic)[SPEC_N],
                                            double (&tot)[SPEC_N], SpecCell &s, double *W, double *sec_out, long long ld,
                                            double dt);
__device__ __forceinline__ void spec_sorption(const double (&c)[SPEC_N], const double (&lna)[SPEC_N],
                                              const double (&ic)[SPEC_N], double (&ts)[SPEC_N], SpecCell &s, double *W,
                                              const DevState &st, long long cell, double jscale);
__device__ __forceinline__ void spec_minerals(const double (&lna)[SPEC_N], const double (&ic)[SPEC_N],
                                              double (&res)[SPEC_N], SpecCell &s, double *W, const DevState &st,
                                              long long cell, bool apply);

#ifndef SPEC_NCLM
#define SPEC_NCLM 0
#endif
__device__ __forceinline__ void spec_sandbox(const double (&c)[SPEC_N], const double (&lna)[SPEC_N],
                                             const double (&tot)[SPEC_N], double (&res)[SPEC_N], SpecCell &s, double *W,
                                             double dt);
// generated: RRadioactiveDecay, RGeneral, RMicrobial, RImmobileDecay (RReaction's order, reaction.F90:4095-4127)
__device__ __forceinline__ void spec_kinetic(const double (&c)[SPEC_N], const double (&lna)[SPEC_N],
                                             const double (&ic)[SPEC_N], const double (&tot)[SPEC_N],
                                             const double (&ts)[SPEC_N], double (&res)[SPEC_N], SpecCell &s, double *W,
                                             double dt);

#if SPEC_NMR > 0
// generated: S_eq of every multirate reaction into s.mr_seq and V * A * dS_eq into the Jacobian
__device__ __forceinline__ void spec_mr_sorption(const double (&lna)[SPEC_N], const double (&ic)[SPEC_N], SpecCell &s,
                                                 double *W, const DevState &st, long long cell);
// RMultiRateSorption, start of a sub-step (reaction_surf_complex.F90:552-637): A and B for this dt.
// The rate loop is rolled (tables), the species loop unrolled; the sum over the rates runs in the
// reference's order for every species.
__device__ __forceinline__ void spec_mr_begin(SpecCell &s, const DevState &st, long long cell, double dt) {
#pragma unroll
  for (int q = 0; q < SPEC_NMR; q++) {
    const int r0 = spec_mr_ptr(q), r1 = spec_mr_ptr(q + 1);
    const long long base = (long long)SPEC_NAQ * (r0 + q);
    double A = 0.0, B[SPEC_NAQ];
#pragma unroll
    for (int i = 0; i < SPEC_NAQ; i++) B[i] = 0.0;
    const long long row = (long long)SPEC_NAQ * st.ld;
    const double *S = st.kinmr + (base + (long long)SPEC_NAQ) * st.ld + cell;
#pragma unroll 2
    for (int k = r0; k < r1; k++, S += row) {
      const double rk = spec_mr_rate_tab[k];
      const double kk = rk / (1.0 + rk * dt);
      A += kk * spec_mr_frac_tab[k];
#pragma unroll
      for (int i = 0; i < SPEC_NAQ; i++) B[i] += kk * S[i * st.ld];
    }
    s.mr_A[q] = A;
#pragma unroll
    for (int i = 0; i < SPEC_NAQ; i++) s.mr_B[q * SPEC_NAQ + i] = B[i];
  }
}
// RSrfCplxMRUpdateKinState (reaction_surf_complex.F90:1107-1145) after an accepted sub-step
__device__ __forceinline__ void spec_mr_update(const SpecCell &s, const DevState &st, long long cell, double dt) {
#pragma unroll
  for (int q = 0; q < SPEC_NMR; q++) {
    const int r0 = spec_mr_ptr(q), r1 = spec_mr_ptr(q + 1);
    const long long base = (long long)SPEC_NAQ * (r0 + q);
    // the equilibrium targets in registers for the whole loop, the row pointer stepped instead of recomputed, two
    // rates per trip so that 30 loads are in flight (the loop is latency-bound: ncu, r02_ncu_c3mr_s1_s3.txt)
    double seq[SPEC_NAQ];
#pragma unroll
    for (int i = 0; i < SPEC_NAQ; i++) seq[i] = s.mr_seq[q * SPEC_NAQ + i];
    const long long row = (long long)SPEC_NAQ * st.ld;
    double *S = st.kinmr + (base + (long long)SPEC_NAQ) * st.ld + cell;
#pragma unroll 2
    for (int k = r0; k < r1; k++, S += row) {
      const double kdt = spec_mr_rate_tab[k] * dt, fk = spec_mr_frac_tab[k];
      // one reciprocal per rate instead of one division per species (the reference divides
      // 15 times by the same 1 + k dt): at most one ulp apart, 700 divisions per cell fewer
      const double inv = 1.0 / (1.0 + kdt);
      const double w = kdt * fk;
      double v[SPEC_NAQ];
#pragma unroll
      for (int i = 0; i < SPEC_NAQ; i++) v[i] = S[i * st.ld];
#pragma unroll
      for (int i = 0; i < SPEC_NAQ; i++) S[i * st.ld] = (v[i] + w * seq[i]) * inv;
    }
  }
}
#endif

__device__ __forceinline__ void spec_cell_constants(SpecCell &s);  // generated

// per-cell inputs of the ELM-CN sandboxes
__device__ __forceinline__ void spec_sandbox_load(SpecCell &s, const DevState &st, long long cell) {
#if SPEC_NIONX > 0
#pragma unroll
  for (int r = 0; r < SPEC_NIONX; r++) s.ixref[r] = st.eqionx_ref ? st.eqionx_ref[r * st.ld + cell] : 1.e-9;
#pragma unroll
  for (int k = 0; k < SPEC_NIXCAT; k++) s.ixconc[k] = 0.0;
#endif
#if SPEC_ELM
  s.elm_w = st.elm_w ? st.elm_w[cell] : 1.0;
  s.elm_o = st.elm_o ? st.elm_o[cell] : 1.0;
  s.elm_t = st.elm_t ? st.elm_t[cell] : 1.0;
  s.elm_zsoil = st.elm_zsoil ? st.elm_zsoil[cell] : 0.0;
  s.elm_kscalar = st.elm_kscalar ? st.elm_kscalar[cell] : 1.0;
  s.elm_bd_dry = st.elm_bd_dry ? st.elm_bd_dry[cell] : 1.25e3;
  s.elm_bsw = st.elm_bsw ? st.elm_bsw[cell] : 1.0;
  s.elm_plantndemand = st.elm_plantndemand ? st.elm_plantndemand[cell] : 0.0;
  s.elm_sucsat = st.elm_sucsat ? st.elm_sucsat[cell] : 200.0;
  s.elm_watfc = st.elm_watfc ? st.elm_watfc[cell] : 0.1;
  s.elm_effpor = st.elm_effpor ? st.elm_effpor[cell] : 0.4;
#endif
#if SPEC_NNC > 0
#pragma unroll
  for (int k = 0; k < SPEC_NNC; k++) s.nc[k] = st.somdec_nc ? st.somdec_nc[k * st.ld + cell] : spec_nc0_tab[k];
#endif
#if SPEC_NKC > 0
  spec_cell_constants(s);
#endif
}
__device__ __forceinline__ void spec_sandbox_store(const SpecCell &s, const DevState &st, long long cell) {
#if SPEC_NIONX > 0
  if (st.eqionx_ref) {
#pragma unroll
    for (int r = 0; r < SPEC_NIONX; r++) st.eqionx_ref[r * st.ld + cell] = s.ixref[r];
  }
  if (st.eqionx_conc) {
#pragma unroll
    for (int k = 0; k < SPEC_NIXCAT; k++) st.eqionx_conc[k * st.ld + cell] = s.ixconc[k];
  }
#endif
#if SPEC_NNC > 0
  if (st.somdec_nc) {
#pragma unroll
    for (int k = 0; k < SPEC_NNC; k++) st.somdec_nc[k * st.ld + cell] = s.nc[k];
  }
#endif
}

// ---- RSolve + LU (reaction.F90:5457-5516, utility.F90:597-735) -------------------
// W = thread's slice (Jacobian of the coupled species), res = residual in
// registers (overwritten by the update).  Crout's method in the reference's
// column order with implicit-scaled partial pivoting.  Rows are never moved:
// ro[i] is the slice offset of the row at logical position i, an interchange
// swaps two offsets, and the scaling factor / right-hand side travel with the
// row in its extra column.  Every register array is indexed by literals.
__device__ __forceinline__ bool spec_rowonly(double *W, double (&res)[SPEC_N], const double (&c)[SPEC_N],
                                             const SpecCell &s, double dt);

#if !SPEC_LOOP_LU
__device__ __forceinline__ bool spec_solve_sparse(double *W, double (&res)[SPEC_N], const double (&c)[SPEC_N]);
#endif

__device__ __forceinline__ bool spec_solve_core(double *W, double (&res)[SPEC_N], const double (&c)[SPEC_N],
                                                const SpecCell &s, double dt) {
  constexpr int N = SPEC_N, NC = SPEC_NC, NCA = NC > 0 ? NC : 1;
  bool bad = false;
  // species outside the matrix: J is diagonal (RTAccumulationDerivative only)
#pragma unroll
  for (int i = 0; i < N; i++) {
    if (spec_cmap(i) < 0 && !spec_is_rowonly(i)) {
      double Jd = (i < SPEC_NAQ) ? (1.0 * (s.den_kg * 1.e-3)) * (s.por * s.sat * 1000.0 * s.vol / dt) : s.vol / dt;
      if (s.dry) Jd = 1.0;
      double nm = sx_rcp(fmax(1.0, fabs(Jd)));
      double a = Jd * nm;
      if (SPEC_USE_LOG) a *= c[i];
      if (!(fabs(a) > 0.0)) bad = true;
      res[i] = sx_div(res[i] * nm, a);
    }
  }
#if SPEC_LOOP_LU
  if (NC == 0) return !bad;
  // The dense part runs as ROLLED loops over the thread's shared-memory slice: the
  // straight-line kernels are bound by instruction fetch (every 128-byte line comes
  // from L2 once per Newton iteration), and an unrolled 13 x 13 LU is 5 500 of those
  // instructions; as loops it is ~300 that stay in the instruction cache.
  constexpr int JS = SPEC_JS, RB = NC, RV = NC + 1;  // columns: right-hand side, scaling factor / solution
  // stage the right-hand side and the iterate (column scaling) -- literal indices
#pragma unroll
  for (int i = 0; i < NC; i++) {
    W[JX(i, RB)] = res[spec_sp_of(i)];
    SW(SPEC_OFF_C + i) = c[spec_sp_of(i)];
  }
  // RSolve: row scaling 1/max(1, max|J_ij|), log formulation column scaling by c_j; the
  // implicit scaling factor of the decomposition in the same pass (utility.F90:611-622)
#pragma unroll 1
  for (int i = 0; i < NC; i++) {
    double *r = W + i * (JS * 32);
    double m = 0.0;
#pragma unroll
    for (int j = 0; j < NC; j++) {
      const double av = fabs(r[j * 32]);
      m = av > m ? av : m;
    }
    const double nm = sx_rcp(fmax(1.0, m));
    r[RB * 32] = r[RB * 32] * nm;
    double m2 = 0.0;
#pragma unroll
    for (int j = 0; j < NC; j++) {
      double v = r[j * 32] * nm;
      if (SPEC_USE_LOG) v *= SW(SPEC_OFF_C + j);
      r[j * 32] = v;
      const double av = fabs(v);
      m2 = av > m2 ? av : m2;
    }
    if (!(m2 > 0.0)) bad = true;
    r[RV * 32] = sx_rcp(m2);
  }
  if (bad) return false;
  // LU with implicit-scaled partial pivoting in right-looking order (same operation
  // sequence per element as the reference's Crout loops), forward substitution fused
  // (the right-hand side is column NC of the augmented rows).  Rows never move:
  // nibble i of perm is the row at logical position i.
  unsigned long long perm = 0xFEDCBA9876543210ull;
#pragma unroll 1
  for (int k = 0; k < NC; k++) {
    double aamax = 0.0;
    int imax = k;
#pragma unroll 1
    for (int i = k; i < NC; i++) {
      const double *r = W + (int)((perm >> (4 * i)) & 15ull) * (JS * 32);
      const double dum = r[RV * 32] * fabs(r[k * 32]);
      if (dum >= aamax) {  // the last maximum wins, like dum.ge.aamax
        imax = i;
        aamax = dum;
      }
    }
    {
      const unsigned long long x = ((perm >> (4 * k)) ^ (perm >> (4 * imax))) & 15ull;
      perm ^= (x << (4 * k)) | (x << (4 * imax));
    }
    double *pr = W + (int)((perm >> (4 * k)) & 15ull) * (JS * 32);
    double pv = pr[k * 32];
    if (pv == 0.0) {
      pv = 1.0e-20;
      pr[k * 32] = pv;
    }
    if (k == NC - 1) break;
    const double rpv = sx_rcp(pv);
    const double *pk = pr + (k + 1) * 32;
    const int cnt = NC - k;  // columns k+1 .. NC-1 and the right-hand side
    // (a version with the pivot row in registers and the tail of a literal-range loop
    // predicated off was measured 60 % slower: 194 vs 122 ms on C3)
#pragma unroll 1
    for (int i = k + 1; i < NC; i++) {
      double *r = W + (int)((perm >> (4 * i)) & 15ull) * (JS * 32) + k * 32;
      const double l = r[0] * rpv;
      r[0] = l;
      r += 32;
#pragma unroll 4
      for (int t = 0; t < cnt; t++) r[t * 32] -= l * pk[t * 32];
    }
  }
  // back substitution (utility.F90:716-735); the solution goes to column RV by logical index
#pragma unroll 1
  for (int i = NC - 1; i >= 0; i--) {
    const double *r = W + (int)((perm >> (4 * i)) & 15ull) * (JS * 32);
    double sum = r[RB * 32];
#pragma unroll 1
    for (int j = i + 1; j < NC; j++) sum -= r[j * 32] * W[JX(j, RV)];
    W[JX(i, RV)] = sx_div(sum, r[i * 32]);
  }
#pragma unroll
  for (int k = 0; k < NC; k++) res[spec_sp_of(k)] = W[JX(k, RV)];
  return true;
}
#else
  if (NC == 0) return !bad;
#ifdef SPEC_SPARSE_LU
  // sparse elimination in the static order the generator chose; false: a multiplier failed the
  // threshold test, the Jacobian is as assembled again and the reference's algorithm below takes over
  if (spec_solve_sparse(W, res, c)) return !bad;
#endif
  double b[NCA];
#pragma unroll
  for (int i = 0; i < NC; i++) {
    double row[NCA];
    double m = 0.0;
    // entries outside the network's structure (spec_jnz) hold exact zeros here: they take no part
    // in the maxima and scaling them is a no-op
#pragma unroll
    for (int j = 0; j < NC; j++) {
      if (spec_jnz(i, j)) {
        row[j] = W[JX(i, j)];
        const double av = fabs(row[j]);
        m = av > m ? av : m;  // maxval(abs()) without fmax's NaN plumbing
      }
    }
    double nm = sx_rcp(fmax(1.0, m));
    b[i] = res[spec_sp_of(i)] * nm;
    double m2 = 0.0;
#pragma unroll
    for (int j = 0; j < NC; j++) {
      if (spec_jnz(i, j)) {
        double v = row[j] * nm;
        if (SPEC_USE_LOG) v *= c[spec_sp_of(j)];
        W[JX(i, j)] = v;
        const double av = fabs(v);
        m2 = av > m2 ? av : m2;
      }
    }
    if (!(m2 > 0.0)) bad = true;
    W[JX(i, NC)] = sx_rcp(m2);
  }
  if (bad) return false;
  int ro[NCA];
#pragma unroll
  for (int i = 0; i < NC; i++) ro[i] = JX(i, 0);
#pragma unroll
  for (int j = 0; j < NC; j++) {
    double u[NCA], sv[NCA];
#pragma unroll
    for (int k = 0; k < NC; k++) {
      if (k < j) {
        const double *r = W + ro[k];
        double sum = r[j * 32];
#pragma unroll
        for (int m = 0; m < NC; m++)
          if (m < k) sum -= r[m * 32] * u[m];
        u[k] = sum;
        if (k > 0) W[ro[k] + j * 32] = sum;
      }
    }
    double aamax = 0.0;
    int imax = j;
#pragma unroll
    for (int i = 0; i < NC; i++) {
      if (i >= j) {
        const double *r = W + ro[i];
        double sum = r[j * 32];
#pragma unroll
        for (int m = 0; m < NC; m++)
          if (m < j) sum -= r[m * 32] * u[m];
        sv[i] = sum;
        double dum = r[NC * 32] * fabs(sum);
        bool ge = dum >= aamax;
        imax = ge ? i : imax;
        aamax = ge ? dum : aamax;
      }
    }
    const int rj = ro[j];
    int rmax = rj;
    double pv = sv[j];
#pragma unroll
    for (int i = 0; i < NC; i++) {
      if (i > j) {
        bool p = (i == imax);
        rmax = p ? ro[i] : rmax;
        pv = p ? sv[i] : pv;
        ro[i] = p ? rj : ro[i];
      }
    }
    ro[j] = rmax;
    if (pv == 0.0) pv = 1.0e-20;
    if (j != NC - 1) {
      double dum = sx_rcp(pv);
      // the logical order has changed already: position imax holds old row j
#pragma unroll
      for (int i = 0; i < NC; i++) {
        if (i > j) {
          bool p = (i == imax);
          double v = p ? sv[j] : sv[i];
          W[ro[i] + j * 32] = v * dum;
        }
      }
    }
    W[rmax + j * 32] = pv;
  }
  // right-hand side into the rows' extra column, then forward / back substitution
#pragma unroll
  for (int i = 0; i < NC; i++) W[JX(i, NC)] = b[i];
#pragma unroll
  for (int k = 0; k < NC; k++) {
    const double *r = W + ro[k];
    double sum = r[NC * 32];
#pragma unroll
    for (int m = 0; m < NC; m++)
      if (m < k) sum -= r[m * 32] * b[m];
    b[k] = sum;
  }
#pragma unroll
  for (int k = NC - 1; k >= 0; k--) {
    const double *r = W + ro[k];
    double sum = b[k];
#pragma unroll
    for (int m = 0; m < NC; m++)
      if (m > k) sum -= r[m * 32] * b[m];
    b[k] = sx_div(sum, r[k * 32]);
  }
#pragma unroll
  for (int k = 0; k < NC; k++) res[spec_sp_of(k)] = b[k];
  return true;
}
#endif

// RSolve for the whole system: the dense core, then the row-only species whose update follows
// from the core's (their columns are diagonal, so nothing in the core depends on them)
__device__ __forceinline__ bool spec_solve(double *W, double (&res)[SPEC_N], const double (&c)[SPEC_N],
                                           const SpecCell &s, double dt) {
  bool ok = spec_solve_core(W, res, c, s, dt);
#if SPEC_NROSPEC > 0
  ok = spec_rowonly(W, res, c, s, dt) && ok;
#endif
  return ok;
}

#ifndef SPEC_LOCKSTEP
#define SPEC_LOCKSTEP 0
#endif
#if !SPEC_LOCKSTEP
// ---- RReact (reaction.F90:3742-4055) ------------------------------------------------
// rt_auxvar%total / %immobile / %total_sorb_eq stay in HBM (st.*), the guess is in
// the thread's shared slice.  Returns ierror; the last iterate is left in c.
__device__ __forceinline__ int spec_react(const DevState &st, const SpecParams &prm, SpecCell &s, double *W,
                                          long long cell, double dt, double (&c)[SPEC_N], double (&guess)[SPEC_N],
                                          int &its_out) {
  constexpr int N = SPEC_N, NAQ = SPEC_NAQ;
  const long long ld = st.ld;
  const double psv = s.por * s.sat * 1000.0 * s.vol;
  s.dry = s.sat < prm.min_sat;
  double fixed[N];  // read once per iteration: registers or local memory, ptxas decides
#pragma unroll
  for (int i = 0; i < N; i++) {
    double f = 0.0;
    if (i < NAQ) {
      if (!s.dry) f = psv * st.total[i * ld + cell];
      if (SPEC_NSORB > 0) f = f + st.total_sorb_eq[i * ld + cell] * s.vol;
    } else {
      if (!s.dry) f = 0.0 + st.immobile[(i - NAQ) * ld + cell] * s.vol;
    }
    SPEC_FIXED(i) = f;
    c[i] = guess[i];
  }
#if SPEC_NMR > 0
  spec_mr_begin(s, st, cell, dt);
#endif
  int its = 0;
  double norm0 = 0.0;
  double lna[N], ic[N], tot[N], res[N], ts[N];
  for (;;) {
    its++;
    if (SPEC_ACT_UPD) spec_activity(c, s);
    spec_rtotal(c, lna, ic, tot, s, W, st.sec_molal + cell, ld, dt);
#pragma unroll
    for (int i = 0; i < N; i++) ts[i] = 0.0;
    if (SPEC_NSORB > 0) spec_sorption(c, lna, ic, ts, s, W, st, cell, s.vol / dt);
    if (its > prm.max_its) {
      // total / immobile keep their initial values in HBM; total_sorb_eq does not
      if (SPEC_NSORB > 0) {
#pragma unroll
        for (int i = 0; i < NAQ; i++) st.total_sorb_eq[i * ld + cell] = ts[i];
      }
      its_out = its;
      return 1;
    }
#pragma unroll
    for (int i = 0; i < N; i++) {
      double a = 0.0;
      if (!s.dry) a = (i < NAQ) ? psv * tot[i] : 0.0 + c[i] * s.vol;
      if (SPEC_NSORB > 0 && i < NAQ) a = a + ts[i] * s.vol;
      res[i] = sx_div(a - SPEC_FIXED(i), dt);
    }
    if (SPEC_NKIN > 0) spec_minerals(lna, ic, res, s, W, st, cell, !s.dry);
#if SPEC_NMR > 0
    if (!s.dry) {  // RMultiRateSorption: Res += V (A S_eq - B), Jac += V A dS_eq
      spec_mr_sorption(lna, ic, s, W, st, cell);
#pragma unroll
      for (int q = 0; q < SPEC_NMR; q++) {
#pragma unroll
        for (int i = 0; i < NAQ; i++)
          res[i] += s.vol * (s.mr_A[q] * s.mr_seq[q * NAQ + i] - s.mr_B[q * NAQ + i]);
      }
    }
#endif
#if SPEC_NKIN3 > 0
    if (!s.dry) spec_kinetic(c, lna, ic, tot, ts, res, s, W, dt);
#endif
#if SPEC_NSBX > 0
    if (!s.dry) spec_sandbox(c, lna, tot, res, s, W, dt);  // RReaction returns before the sandboxes in a dry cell
#endif
    double mabs = 0.0, ss = 0.0;
#pragma unroll
    for (int i = 0; i < N; i++) {
      {
          const double av = fabs(res[i]);  // maxval(abs(residual)): a NaN never wins, as with fmax
          mabs = av > mabs ? av : mabs;
        }
      ss += res[i] * res[i];
    }
    double nrm = sqrt(ss);
    if (its == 1) norm0 = nrm;
    double rel = nrm / norm0;
    bool conv = (mabs < prm.tol_res) || (rel < prm.tol_relres);
    if (!conv) {
      if (!spec_solve(W, res, c, s, dt)) {
        // solve_error branch: no restore (reaction.F90:3964-3967)
#pragma unroll
        for (int i = 0; i < N; i++) {
          if (i < NAQ) {
            st.total[i * ld + cell] = tot[i];
            if (SPEC_NSORB > 0) st.total_sorb_eq[i * ld + cell] = ts[i];
          } else {
            st.immobile[(i - NAQ) * ld + cell] = c[i];
          }
        }
        its_out = its;
        return 1;
      }
      double cn[N], maxrel = -1.0;
      double minr = 1.e20;
      if (!SPEC_USE_LOG) {
        minr = spec_min_ratio(c, res);
      }
#pragma unroll
      for (int i = 0; i < N; i++) {
        double u = res[i];
        if (SPEC_USE_LOG) {
          u = copysign(1.0, u) * fmin(fabs(u), prm.max_dlnC);
          cn[i] = c[i] * sx_exp(-u);
        } else {
          if (minr < 1.0) u = u * minr * 0.99;
          cn[i] = c[i] - u;
        }
        double v = fabs(sx_div(cn[i] - c[i], c[i]));
        // IEEE: x / 0 = Inf (no convergence from an exactly-zero iterate), 0 / 0 = NaN (skipped)
        if (fabs(c[i]) < 2.2250738585072014e-308) v = (cn[i] == c[i]) ? v : (double)INFINITY;
        maxrel = v > maxrel ? v : maxrel;  // a NaN is skipped (false), as by MAXVAL
      }
      conv = (maxrel >= 0.0) && (maxrel < prm.tol_relchange);
      if (!conv) {
#pragma unroll
        for (int i = 0; i < N; i++) c[i] = cn[i];
        continue;
      }
    }
    break;
  }
  // converged: the reference's last RTAuxVarCompute (reaction.F90:4052) recomputes
  // RTotal at the same c -- the values are in tot / ts / st.sec_molal already
#pragma unroll
  for (int i = 0; i < N; i++) {
    if (i < NAQ) {
      st.total[i * ld + cell] = tot[i];
      if (SPEC_NSORB > 0) st.total_sorb_eq[i * ld + cell] = ts[i];
    } else {
      st.immobile[(i - NAQ) * ld + cell] = c[i];
    }
    guess[i] = c[i];
  }
  its_out = its;
  return 0;
}

// ---- RStep (reaction.F90:3564-3738) for one cell ---------------------------------------
__device__ __forceinline__ void spec_run(const DevState &st, const SpecParams &prm, double *W, long long cell,
                                         double target, int &nss, int &nit, int &nku, int &ierr, bool &had_cut) {
  constexpr int N = SPEC_N, NAQ = SPEC_NAQ;
  const long long ld = st.ld;
  SpecCell s;
  s.store = true;
  s.den_kg = st.den_kg[cell];
  s.sat = st.sat[cell];
  s.temp = st.temp[cell];
  s.por = st.porosity[cell];
  s.vol = st.volume[cell];
  s.spd = st.soil_particle_density ? st.soil_particle_density[cell] : 0.0;
  s.ln_act_h2o = st.ln_act_h2o ? st.ln_act_h2o[cell] : 0.0;
  spec_sandbox_load(s, st, cell);
#if SPEC_NMR > 0
#pragma unroll
  for (int q = 0; q < SPEC_NMR; q++) {
    const long long base = (long long)NAQ * (spec_mr_ptr(q) + q);
#pragma unroll
    for (int i = 0; i < NAQ; i++) s.mr_seq[q * NAQ + i] = st.kinmr[(base + i) * ld + cell];
  }
#endif
  nss = nit = nku = ierr = 0;
  had_cut = false;
  double Is = 0.0, ms = 0.0;
#pragma unroll 8
  for (int k = 0; k < SPEC_NCX; k++) {
    double m = st.sec_molal[k * ld + cell];
    Is += m * spec_cx_z2(k);
    ms += m;
    if (!SPEC_ACT_UPD) SW(SPEC_OFF_LNGSEC + k) = log(st.sec_act_coef[k * ld + cell]);
  }
  s.Isec = Is;
  s.msec = ms;
#pragma unroll
  for (int k = 0; k < (SPEC_NCLS > 0 ? SPEC_NCLS : 1); k++) s.lgcls[k] = 0.0;
#pragma unroll
  for (int k = 0; k < SPEC_NSRFRXN; k++) s.fsite[k] = st.free_site[k * ld + cell];
#pragma unroll
  for (int k = 0; k < SPEC_NSRFCPLX; k++) s.scconc[k] = 0.0;
#pragma unroll
  for (int k = 0; k < SPEC_NKIN; k++) s.mrate[k] = st.mnrl_rate[k * ld + cell];
  unsigned small_mask = 0u;
  double small_val[N], guess[N];
  // all loads first: the clamping stores below would otherwise order them
  double in_t[N], in_g[N], in_a[N];
#pragma unroll
  for (int i = 0; i < N; i++) {
    in_a[i] = (i < NAQ) ? st.pri_act_coef[i * ld + cell] : 1.0;
    in_g[i] = (i < NAQ) ? st.pri_molal[i * ld + cell] : 0.0;
    in_t[i] = (i < NAQ) ? st.total[i * ld + cell] : st.immobile[(i - NAQ) * ld + cell];
  }
#pragma unroll
  for (int i = 0; i < N; i++) {
    SPEC_SMALL(i) = 0.0;
    if (!SPEC_ACT_UPD) s.lngam[SPEC_ACT_UPD ? 0 : i] = 0.0;
    if (i < NAQ) {
      if (!SPEC_ACT_UPD) s.lngam[SPEC_ACT_UPD ? 0 : i] = log(in_a[i]);
      double g = in_g[i];
      double t = in_t[i];
      if (t <= 1.e-40) {
        small_mask |= 1u << i;
        SPEC_SMALL(i) = t;
        t = 1.e-40;
        st.total[i * ld + cell] = t;
      }
      guess[i] = g;
    } else {
      double t = in_t[i];
      guess[i] = t;  // the guess keeps the unclamped value
      if (t <= 1.e-40) {
        small_mask |= 1u << i;
        SPEC_SMALL(i) = t;
        st.immobile[(i - NAQ) * ld + cell] = 1.e-40;
      }
    }
  }
  double cumulative = 0.0, dt = target;
  int ncuts = 0, nconst = 0;
  bool aborted = false;
  double c[N];
  for (;;) {
    if (cumulative >= target) break;
    int its = 0;
    int e = spec_react(st, prm, s, W, cell, dt, c, guess, its);
    nit += its;
    if (e != 0) {
      ncuts++;
      had_cut = true;
      if (ncuts > prm.max_cuts) {
        aborted = true;
        break;
      }
      dt = 0.5 * dt;
      nconst = 0;
    } else {
      // RUpdateKineticState: the rates of the converged iterate are in s.mrate
      bool upd = SPEC_NSBX > 0;  // a sandbox forces the kinetic-state update (reaction.F90:5935-5972)
      if (SPEC_NKIN > 0) {
        upd = true;
#pragma unroll
        for (int m = 0; m < SPEC_NKIN; m++) {
          double vf = st.mnrl_volfrac[m * ld + cell] + s.mrate[m] * spec_mn_vol(m) * dt;
          if (vf < 0.0) vf = 0.0;
          st.mnrl_volfrac[m * ld + cell] = vf;
        }
      }
#if SPEC_NMR > 0
      upd = true;
      spec_mr_update(s, st, cell, dt);
#endif
      cumulative += dt;
      nss++;
      nconst++;
      if (upd) nku++;
      if (nconst >= 4) {
        ncuts--;
        dt = fmin(2.0 * dt, target - cumulative);
      }
    }
  }
  if (aborted) ierr = 1;
#pragma unroll
  for (int i = 0; i < N; i++) {
    if (i < NAQ) st.pri_molal[i * ld + cell] = aborted ? c[i] : guess[i];
    if (!aborted && ((small_mask >> i) & 1u)) {
      if (i < NAQ)
        st.total[i * ld + cell] = SPEC_SMALL(i);
      else
        st.immobile[(i - NAQ) * ld + cell] = SPEC_SMALL(i);
    }
  }
  if (SPEC_ACT_UPD) {
#pragma unroll
    for (int i = 0; i < NAQ; i++) st.pri_act_coef[i * ld + cell] = sx_exp(SPEC_LNGAM(s, i));
#pragma unroll 4
    for (int k = 0; k < SPEC_NCX; k++) {
      int q = spec_cx_cls(k);
      double lg = 0.0;
#pragma unroll
      for (int z = 0; z < SPEC_NCLS; z++)
        if (z == q) lg = s.lgcls[z];
      st.sec_act_coef[k * ld + cell] = q < 0 ? 1.0 : sx_exp(lg);
    }
  }
#pragma unroll
  for (int k = 0; k < SPEC_NSRFRXN; k++) st.free_site[k * ld + cell] = s.fsite[k];
  if (SPEC_NEQSR > 0 && st.eqsrfcplx_conc) {
#pragma unroll
    for (int k = 0; k < SPEC_NSRFCPLX; k++) st.eqsrfcplx_conc[k * ld + cell] = s.scconc[k];
  }
#pragma unroll
  for (int k = 0; k < SPEC_NKIN; k++) st.mnrl_rate[k * ld + cell] = s.mrate[k];
  if (st.ln_act_h2o && SPEC_USE_ACT_H2O) st.ln_act_h2o[cell] = s.ln_act_h2o;
  spec_sandbox_store(s, st, cell);
#if SPEC_NMR > 0
#pragma unroll
  for (int q = 0; q < SPEC_NMR; q++) {
    const long long base = (long long)NAQ * (spec_mr_ptr(q) + q);
#pragma unroll
    for (int i = 0; i < NAQ; i++) st.kinmr[(base + i) * ld + cell] = s.mr_seq[q * NAQ + i];
  }
#endif
}

extern "C" __global__ void __launch_bounds__(SPEC_THREADS, SPEC_MINBLOCKS)
    pfrx_spec_kernel(DevState st, long long ncell, double tran_dt, SpecParams prm, DevSummary *summ) {
  extern __shared__ double smem[];
  const int lane32 = threadIdx.x & 31;
  double *W = smem + (size_t)(threadIdx.x >> 5) * (SPEC_SLOTS * 32) + lane32;

  unsigned long long l_active = 0, l_its = 0, l_cut = 0;
  long long l_first = -1;
  int l_maxits = 0, l_maxkin = 0, l_maxerr = 0, l_maxsub = 0;

  const long long gthread = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  for (long long cell = gthread; cell < ncell; cell += nthreads) {
    int nss = 0, nit = 0, nku = 0, ierr = 0;
    bool cut = false;
    bool active = !(st.imat && st.imat[cell] <= 0);
    if (active) spec_run(st, prm, W, cell, tran_dt, nss, nit, nku, ierr, cut);
    st.num_sub_steps[cell] = nss;
    st.num_iterations[cell] = nit;
    st.num_kinetic_state_updates[cell] = nku;
    st.ierror[cell] = ierr;
    if (active) {
      l_active++;
      l_its += (unsigned long long)nit;
      if (cut) l_cut++;
      if (ierr != 0 && (l_first < 0 || cell < l_first)) l_first = cell;
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
#else

#if SPEC_NMR > 0
#error "multirate sorption is generated for the one-warp skeleton (style s) only"
#endif
// ======================================================================================
// Lock-step skeleton (SPEC_LOCKSTEP 1): the four warps of a 128-thread block execute the
// SAME Newton iteration at the same time.
//
// Why (profiles/r01_ncu_c5_spec_s1.txt vs r01_ncu_c3_spec_s1f.txt): the generated code is
// 280 KB per Newton iteration and an SM can pull only ~6 B/clk of DISTINCT instruction
// bytes out of L2.  Four warps that drift apart are four streams (C5: iteration counts
// differ from cell to cell, 6.7 fetch-stall cycles per instruction); four warps at the
// same place are one stream delivered four times.  So RStep / RReact are restated as a
// state machine per lane: one pass of the block loop = one Newton iteration of every
// unfinished cell; a finished lane keeps executing with its stores disabled; the two
// block votes per pass are the barriers that keep the warps together.
// ======================================================================================
extern "C" __global__ void __launch_bounds__(SPEC_THREADS, SPEC_MINBLOCKS)
    pfrx_spec_kernel(DevState st, long long ncell, double tran_dt, SpecParams prm, DevSummary *summ) {
  constexpr int N = SPEC_N, NAQ = SPEC_NAQ;
  extern __shared__ double smem[];
  const int lane32 = threadIdx.x & 31;
  double *W = smem + (size_t)(threadIdx.x >> 5) * (SPEC_SLOTS * 32) + lane32;
  const long long ld = st.ld;
  const double target = tran_dt;

  unsigned long long l_active = 0, l_its = 0, l_cut = 0;
  long long l_first = -1;
  int l_maxits = 0, l_maxkin = 0, l_maxerr = 0, l_maxsub = 0;

  // per-lane state of the cell a lane holds (declared here: it survives passes of the loop)
  SpecCell s;
  s.store = false;
  s.dry = false;
  unsigned small_mask = 0u;
  // the guess of the next sub-step is kept where it ends up anyway: rt_auxvar%pri_molal
  // (aqueous species); immobile species keep theirs in registers
  double small_val[N], gimm[N > NAQ ? N - NAQ : 1], fixed[N], c[N];
  double cumulative = 0.0, dt = target, norm0 = 0.0, psv = 0.0;
  int ncuts = 0, nconst = 0, nss = 0, nit = 0, nku = 0, its = 0;
  bool done = true, aborted = false, had_cut = false, need_begin = false;
  long long cell = ncell - 1;
  bool inrange = false, live = false;
#if SPEC_REFILL
  // Refill (variant "q"): a lane whose cell is finished takes the next unprocessed cell
  // instead of riding along until the slowest cell of its block is done.  Cells are handed
  // out by a warp-aggregated atomic counter; which lane gets which cell does not matter,
  // every cell is computed from its own state only.
  bool have = false, loaded_once = false;
#else
  long long base = (long long)blockIdx.x * blockDim.x;
  bool need_new = true;
#endif

  for (;;) {
    bool fresh = false;
#if SPEC_REFILL
    {
      const unsigned want = __ballot_sync(0xffffffffu, !have);
      if (!have) {
        const int leader = __ffs(want) - 1;
        unsigned long long first = 0;
        if (lane32 == leader) first = atomicAdd(&summ->next_cell, (unsigned long long)__popc(want));
        first = __shfl_sync(want, first, leader);
        const long long mine = (long long)first + __popc(want & ((1u << lane32) - 1u));
        if (mine < ncell) {
          cell = st.order ? (long long)st.order[mine] : mine;
          inrange = true;
          live = !(st.imat && st.imat[cell] <= 0);
          fresh = true;
          have = true;
        } else if (!loaded_once) {
          cell = ncell - 1;  // never got a cell: hold a valid state to ride along with
          inrange = false;
          live = false;
          fresh = true;
        }
      }
      loaded_once = true;
    }
#else
    if (need_new) {
      if (base >= ncell) break;
      inrange = base + threadIdx.x < ncell;
      cell = inrange ? base + threadIdx.x : ncell - 1;
      live = inrange && !(st.imat && st.imat[cell] <= 0);
      base += (long long)gridDim.x * blockDim.x;
      need_new = false;
      fresh = true;
    }
#endif
    if (fresh) {
    // ---- RStep entry (reaction.F90:3600-3650)
    s.den_kg = st.den_kg[cell];
    s.sat = st.sat[cell];
    s.temp = st.temp[cell];
    s.por = st.porosity[cell];
    s.vol = st.volume[cell];
    s.spd = st.soil_particle_density ? st.soil_particle_density[cell] : 0.0;
    s.ln_act_h2o = st.ln_act_h2o ? st.ln_act_h2o[cell] : 0.0;
    spec_sandbox_load(s, st, cell);
    s.dry = s.sat < prm.min_sat;
    psv = s.por * s.sat * 1000.0 * s.vol;
    {
      double Is = 0.0, ms = 0.0;
#pragma unroll 8
      for (int k = 0; k < SPEC_NCX; k++) {
        double m = st.sec_molal[k * ld + cell];
        Is += m * spec_cx_z2(k);
        ms += m;
        if (!SPEC_ACT_UPD) SW(SPEC_OFF_LNGSEC + k) = log(st.sec_act_coef[k * ld + cell]);
      }
      s.Isec = Is;
      s.msec = ms;
    }
#pragma unroll
    for (int k = 0; k < (SPEC_NCLS > 0 ? SPEC_NCLS : 1); k++) s.lgcls[k] = 0.0;
#pragma unroll
    for (int k = 0; k < SPEC_NSRFRXN; k++) s.fsite[k] = st.free_site[k * ld + cell];
#pragma unroll
    for (int k = 0; k < SPEC_NSRFCPLX; k++) s.scconc[k] = 0.0;
#pragma unroll
    for (int k = 0; k < SPEC_NKIN; k++) s.mrate[k] = st.mnrl_rate[k * ld + cell];
    small_mask = 0u;
    {
      double in_t[N], in_g[N], in_a[N];
#pragma unroll
      for (int i = 0; i < N; i++) {
        in_a[i] = (i < NAQ) ? st.pri_act_coef[i * ld + cell] : 1.0;
        in_g[i] = (i < NAQ) ? st.pri_molal[i * ld + cell] : 0.0;
        in_t[i] = (i < NAQ) ? st.total[i * ld + cell] : st.immobile[(i - NAQ) * ld + cell];
      }
#pragma unroll
      for (int i = 0; i < N; i++) {
        SPEC_SMALL(i) = 0.0;
        SPEC_FIXED(i) = 0.0;
        if (!SPEC_ACT_UPD) s.lngam[SPEC_ACT_UPD ? 0 : i] = 0.0;
        if (i < NAQ) {
          if (!SPEC_ACT_UPD) s.lngam[SPEC_ACT_UPD ? 0 : i] = log(in_a[i]);
          c[i] = in_g[i];
          if (in_t[i] <= 1.e-40) {
            small_mask |= 1u << i;
            SPEC_SMALL(i) = in_t[i];
            if (live) st.total[i * ld + cell] = 1.e-40;
          }
        } else {
          gimm[i - NAQ] = in_t[i];  // the guess keeps the unclamped value
          c[i] = in_t[i];
          if (in_t[i] <= 1.e-40) {
            small_mask |= 1u << i;
            SPEC_SMALL(i) = in_t[i];
            if (live) st.immobile[(i - NAQ) * ld + cell] = 1.e-40;
          }
        }
      }
    }
    cumulative = 0.0;
    dt = target;
    norm0 = 0.0;
    ncuts = nconst = nss = nit = nku = its = 0;
    done = !live;
    aborted = false;
    had_cut = false;
    need_begin = true;
    }  // fresh
#if SPEC_REFILL
    if (__syncthreads_and(have ? 0 : 1)) break;  // every cell has been handed out and finished
#endif

    {
      // ---- RReact entry (reaction.F90:3829-3850) for lanes that start a sub-step
      if (!done && need_begin) {
#pragma unroll
        for (int i = 0; i < N; i++) {
          double f = 0.0;
          if (i < NAQ) {
            if (!s.dry) f = psv * st.total[i * ld + cell];
            if (SPEC_NSORB > 0) f = f + st.total_sorb_eq[i * ld + cell] * s.vol;
          } else {
            if (!s.dry) f = 0.0 + st.immobile[(i - NAQ) * ld + cell] * s.vol;
          }
          SPEC_FIXED(i) = f;
          c[i] = (i < NAQ) ? st.pri_molal[i * ld + cell] : gimm[i - NAQ];
        }
        its = 0;
        need_begin = false;
      }
      if (!done) its++;
      s.store = !done;

      // ---- one Newton iteration (reaction.F90:3860-4041), every lane of the block -- except
      // that a warp whose 32 cells are all finished skips the arithmetic (it still votes): at
      // the ragged end of a launch the live warps then have the instruction supply to themselves
      double lna[N], ic[N], tot[N], res[N], ts[N];
      bool conv = false, need_solve = false, fail = false;
      const bool warp_live = __any_sync(0xffffffffu, !done);
      if (warp_live) {
      if (SPEC_ACT_UPD) spec_activity(c, s);
      spec_rtotal(c, lna, ic, tot, s, W, st.sec_molal + cell, ld, dt);
#pragma unroll
      for (int i = 0; i < N; i++) ts[i] = 0.0;
      if (SPEC_NSORB > 0) spec_sorption(c, lna, ic, ts, s, W, st, cell, s.vol / dt);
      const bool over = its > prm.max_its;
#pragma unroll
      for (int i = 0; i < N; i++) {
        double a = 0.0;
        if (!s.dry) a = (i < NAQ) ? psv * tot[i] : 0.0 + c[i] * s.vol;
        if (SPEC_NSORB > 0 && i < NAQ) a = a + ts[i] * s.vol;
        res[i] = sx_div(a - SPEC_FIXED(i), dt);
      }
      if (SPEC_NKIN > 0) spec_minerals(lna, ic, res, s, W, st, cell, !s.dry);
#if SPEC_NKIN3 > 0
      if (!s.dry) spec_kinetic(c, lna, ic, tot, ts, res, s, W, dt);
#endif
#if SPEC_NSBX > 0
      if (!s.dry) spec_sandbox(c, lna, tot, res, s, W, dt);
#endif
      double mabs = 0.0, ss = 0.0;
#pragma unroll
      for (int i = 0; i < N; i++) {
        {
          const double av = fabs(res[i]);  // maxval(abs(residual)): a NaN never wins, as with fmax
          mabs = av > mabs ? av : mabs;
        }
        ss += res[i] * res[i];
      }
      const double nrm = sqrt(ss);
      if (its == 1) norm0 = nrm;
      const double rel = nrm / norm0;
      conv = (mabs < prm.tol_res) || (rel < prm.tol_relres);
      need_solve = !done && !over && !conv;
      fail = !done && over;
      }  // warp_live
      bool solve_error = false;

      if (__syncthreads_or(need_solve ? 1 : 0) && warp_live) {
        const bool ok = spec_solve(W, res, c, s, dt);
        if (need_solve) {
          if (!ok) {
            fail = true;
            solve_error = true;
          } else {
            double cn[N], maxrel = -1.0, minr = 1.e20;
            if (!SPEC_USE_LOG) {
              minr = spec_min_ratio(c, res);
            }
#pragma unroll
            for (int i = 0; i < N; i++) {
              double u = res[i];
              if (SPEC_USE_LOG) {
                u = copysign(1.0, u) * fmin(fabs(u), prm.max_dlnC);
                cn[i] = c[i] * sx_exp(-u);
              } else {
                if (minr < 1.0) u = u * minr * 0.99;
                cn[i] = c[i] - u;
              }
              double v = fabs(sx_div(cn[i] - c[i], c[i]));
              if (fabs(c[i]) < 2.2250738585072014e-308) v = (cn[i] == c[i]) ? v : (double)INFINITY;
              maxrel = v > maxrel ? v : maxrel;  // a NaN is skipped (false), as by MAXVAL
            }
            if ((maxrel >= 0.0) && (maxrel < prm.tol_relchange)) {
              conv = true;
            } else {
#pragma unroll
              for (int i = 0; i < N; i++) c[i] = cn[i];
            }
          }
        }
      }

      // ---- what this pass decided for the lane (RReact exit + RStep, reaction.F90:3655-3738)
      if (!done) {
        if (fail) {
          nit += its;
          // its > max: total / immobile keep their values in HBM, total_sorb_eq does not;
          // solve error: no restore (reaction.F90:3964-3967)
#pragma unroll
          for (int i = 0; i < N; i++) {
            if (i < NAQ) {
              if (SPEC_NSORB > 0) st.total_sorb_eq[i * ld + cell] = ts[i];
              if (solve_error) st.total[i * ld + cell] = tot[i];
            } else if (solve_error) {
              st.immobile[(i - NAQ) * ld + cell] = c[i];
            }
          }
          ncuts++;
          had_cut = true;
          if (ncuts > prm.max_cuts) {
            aborted = true;
            done = true;
          } else {
            dt = 0.5 * dt;
            nconst = 0;
            need_begin = true;
          }
        } else if (conv) {
          nit += its;
#pragma unroll
          for (int i = 0; i < N; i++) {
            if (i < NAQ) {
              st.total[i * ld + cell] = tot[i];
              if (SPEC_NSORB > 0) st.total_sorb_eq[i * ld + cell] = ts[i];
              st.pri_molal[i * ld + cell] = c[i];
            } else {
              st.immobile[(i - NAQ) * ld + cell] = c[i];
              gimm[i - NAQ] = c[i];
            }
          }
          bool upd = SPEC_NSBX > 0;  // a sandbox forces the kinetic-state update (reaction.F90:5935-5972)
          if (SPEC_NKIN > 0) {
            upd = true;
#pragma unroll
            for (int m = 0; m < SPEC_NKIN; m++) {
              double vf = st.mnrl_volfrac[m * ld + cell] + s.mrate[m] * spec_mn_vol(m) * dt;
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
            dt = fmin(2.0 * dt, target - cumulative);
          }
          if (cumulative >= target)
            done = true;
          else
            need_begin = true;
        }
      }
    }
#if SPEC_REFILL
    const bool publish = have && done;
#else
    const bool publish = __syncthreads_and(done ? 1 : 0) != 0;
    if (publish) need_new = true;
#endif
    if (publish) {
    // ---- publish the cell (reaction.F90:3700-3738); the generated routines stop updating a
    // lane's activity / sorption / rate state once s.store is false, so this is the state of
    // the lane's last own pass
    if (live) {
#pragma unroll
      for (int i = 0; i < N; i++) {
        if (i < NAQ && aborted) st.pri_molal[i * ld + cell] = c[i];
        if (!aborted && ((small_mask >> i) & 1u)) {
          if (i < NAQ)
            st.total[i * ld + cell] = SPEC_SMALL(i);
          else
            st.immobile[(i - NAQ) * ld + cell] = SPEC_SMALL(i);
        }
      }
      if (SPEC_ACT_UPD) {
#pragma unroll
        for (int i = 0; i < NAQ; i++) st.pri_act_coef[i * ld + cell] = sx_exp(SPEC_LNGAM(s, i));
#pragma unroll 4
        for (int k = 0; k < SPEC_NCX; k++) {
          int q = spec_cx_cls(k);
          double lg = 0.0;
#pragma unroll
          for (int z = 0; z < SPEC_NCLS; z++)
            if (z == q) lg = s.lgcls[z];
          st.sec_act_coef[k * ld + cell] = q < 0 ? 1.0 : sx_exp(lg);
        }
      }
#pragma unroll
      for (int k = 0; k < SPEC_NSRFRXN; k++) st.free_site[k * ld + cell] = s.fsite[k];
      if (SPEC_NEQSR > 0 && st.eqsrfcplx_conc) {
#pragma unroll
        for (int k = 0; k < SPEC_NSRFCPLX; k++) st.eqsrfcplx_conc[k * ld + cell] = s.scconc[k];
      }
#pragma unroll
      for (int k = 0; k < SPEC_NKIN; k++) st.mnrl_rate[k * ld + cell] = s.mrate[k];
      if (st.ln_act_h2o && SPEC_USE_ACT_H2O) st.ln_act_h2o[cell] = s.ln_act_h2o;
      spec_sandbox_store(s, st, cell);
    }

    if (inrange) {
      st.num_sub_steps[cell] = nss;
      st.num_iterations[cell] = nit;
      st.num_kinetic_state_updates[cell] = nku;
      st.ierror[cell] = aborted ? 1 : 0;
      if (live) {
        l_active++;
        l_its += (unsigned long long)nit;
        if (had_cut) l_cut++;
        if (aborted && (l_first < 0 || cell < l_first)) l_first = cell;
        l_maxits = max(l_maxits, nit);
        l_maxkin = max(l_maxkin, nku);
        l_maxerr = max(l_maxerr, aborted ? 1 : 0);
        l_maxsub = max(l_maxsub, nss);
      }
    }
#if SPEC_REFILL
    have = false;
#endif
    }  // publish
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
#endif
