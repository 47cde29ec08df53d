// pfrx_specr.cuh -- "rolled" network-specialised kernel: SPEC_W warps per group of 32
// cells, ONE instruction stream for all of them.
//
// What the two straight-line variants taught (profiles/r01_ncu_c3_spec*.txt): with
// every coefficient an immediate the Hanford network costs ~19 000 instructions per
// Newton iteration, 300 KB of code that each warp streams through once per
// iteration -- the kernels stall on instruction fetch, and splitting the work over
// more warps (pfrx_specw.cuh) multiplies the streams.  Here the network is data
// again, but data in __constant__ memory (uniform loads), the sizes are
// compile-time, and the loops are short enough to stay in the instruction cache
// while 16 warps per SM share them.
//
// Work split (lane l of every warp of a block = the same cell):
//   species i        -> warp i % W, register slot i / W   (c, guess, fixed, total, residual, J row)
//   complex k        -> warp k % W                        (the exp)
//   LU logical row i -> warp i % W
// Exchange: ln a_j, 1/c_j, c_j, partial norms, pivot candidates in the block's
// shared slice; complex concentrations through rt_auxvar%sec_molal (L2).
// Shared memory per cell: NC x (NC+1) Jacobian (extra column: ln a_i, then the
// row's scaling factor, then the right-hand side) + 1/c + c + 3W exchange slots;
// element e of a cell is at slice[e*32 + lane] -- conflict-free for any per-lane
// row permutation.
//
// The LU is the reference's (utility.F90:597-735) in right-looking order: the same
// fused multiply-add sequence per element, the same `>=' pivot rule.  The row
// permutation is a packed register (4 bits per logical position).
//
// Control flow is lock-step (one pass = one Newton iteration of the whole group);
// every decision that steers a barrier is computed by the same instructions from
// exchanged values in all warps.
#pragma once
#include <cuda_runtime.h>

#include "pfrx_types.cuh"

#define SPEC_LN 2.30258509299  // pflotran_constants.F90:84 (truncated there)

#define SPEC_JS (SPEC_NC + 1)
#define SW_OFF_IC (SPEC_NC * SPEC_JS)
#define SW_OFF_C (SW_OFF_IC + SPEC_NC)
#define SW_OFF_RED (SW_OFF_C + SPEC_NC)
#define SW_RED_SLOTS (3 * SPEC_W)
#define SPEC_SLOTS (SW_OFF_RED + SW_RED_SLOTS)
#define SW(e) W[(e) * 32]
#define JX(ci, cj) (((ci) * SPEC_JS + (cj)) * 32)
#define EXS(ci) ((ci) * SPEC_JS + SPEC_NC)
#define SPEC_QN ((SPEC_N + SPEC_W - 1) / SPEC_W)

static_assert(SPEC_NC >= 2 * SPEC_W && SPEC_NC <= 16, "exchange buffers alias NC-slot regions; 4-bit row indices");
static_assert(SPEC_NCLS <= 16 && SPEC_NC * SPEC_JS >= 16 * SPEC_W, "activity classes use the idle Jacobian as scratch");

extern "C" {
__device__ const unsigned long long pfrx_spec_sig = SPEC_SIG;
// {N, shared doubles per block, threads per block, min blocks per SM, cells per block}
__device__ const int pfrx_spec_info[5] = {SPEC_N, SPEC_SLOTS * 32, 32 * SPEC_W, SPEC_MINBLOCKS, 32};
}

__device__ __forceinline__ void wt_barrier() { asm volatile("bar.sync 0;" ::: "memory"); }

template <int Q>
__device__ __forceinline__ void wt_add_slot(double (&a)[Q], int slot, double v) {
#pragma unroll
  for (int q = 0; q < Q; q++)
    if (q == slot) a[q] += v;
}

// Debye-Hueckel ln gamma of an activity class (reaction.F90:4575-4600, LAG algorithm)
__device__ __forceinline__ double wt_lngamma(int q, double I, double sq) {
  return (T_cls_negz2[q] * sq * SPEC_DEBYE_A / (1.0 + T_cls_a0[q] * SPEC_DEBYE_B * sq) + SPEC_DEBYE_BDOT * I) * SPEC_LN;
}

__device__ __forceinline__ void wt_run(const DevState &st, const long long ncell, const double target,
                                       const SpecParams &prm, DevSummary *summ, double *W, const int wid) {
  constexpr int N = SPEC_N, NAQ = SPEC_NAQ, NC = SPEC_NC, NW = SPEC_W, QN = SPEC_QN;
  constexpr int MAXQ = SPEC_MAXQ, NEQ = SPEC_NEQSR > 0 ? SPEC_NEQSR : 1, NKA = SPEC_NKIN > 0 ? SPEC_NKIN : 1;
  const int lane = threadIdx.x & 31;
  const long long ld = st.ld;
  const unsigned FULL = 0xffffffffu;

  unsigned long long l_active = 0, l_its = 0, l_cut = 0;
  long long l_first = -1;
  int l_maxits = 0, l_maxkin = 0, l_maxerr = 0, l_maxsub = 0;

  for (long long base = (long long)blockIdx.x * 32; base < ncell; base += (long long)gridDim.x * 32) {
    const bool inrange = base + lane < ncell;
    const long long cell = inrange ? base + lane : ncell - 1;
    const bool live = inrange && !(st.imat && st.imat[cell] <= 0);

    const double den_kg = st.den_kg[cell], sat = st.sat[cell], temp = st.temp[cell], por = st.porosity[cell],
                 vol = st.volume[cell];
    const double spd = st.soil_particle_density ? st.soil_particle_density[cell] : 0.0;
    const double ln_act_h2o = st.ln_act_h2o ? st.ln_act_h2o[cell] : 0.0;
    const bool dry = sat < prm.min_sat;
    const double psv = por * sat * 1000.0 * vol;
    const double denL = den_kg * 1.e-3;
    double Is_part = 0.0;
#pragma unroll 2
    for (int k = wid; k < SPEC_NCX; k += NW) Is_part += st.sec_molal[k * ld + cell] * T_cx_z2[k];
    double I_last = 0.0;

    // own species: guess, clamped totals (RStep, reaction.F90:3633-3650)
    double guess[QN], fixed[QN], cdec[QN], small_val[QN], tot[QN], ts[QN], res[QN], lngam[QN];
    unsigned small_mask = 0u;
#pragma unroll
    for (int q = 0; q < QN; q++) {
      guess[q] = fixed[q] = cdec[q] = small_val[q] = tot[q] = ts[q] = res[q] = lngam[q] = 0.0;
      const int i = wid + q * NW;
      if (i >= N) continue;
      if (i < NAQ) {
        lngam[q] = log(st.pri_act_coef[i * ld + cell]);
        guess[q] = st.pri_molal[i * ld + cell];
        double t = st.total[i * ld + cell];
        if (t <= 1.e-40) {
          small_mask |= 1u << q;
          small_val[q] = t;
          if (live) st.total[i * ld + cell] = 1.e-40;
        }
      } else {
        double t = st.immobile[(i - NAQ) * ld + cell];
        guess[q] = t;
        if (t <= 1.e-40) {
          small_mask |= 1u << q;
          small_val[q] = t;
          if (live) st.immobile[(i - NAQ) * ld + cell] = 1.e-40;
        }
      }
    }
    double fsite[NEQ], Ssc[NEQ][MAXQ], mrate[NKA];
#pragma unroll
    for (int e = 0; e < NEQ; e++) {
      fsite[e] = (SPEC_NEQSR > 0) ? st.free_site[T_eq[e] * ld + cell] : 0.0;
#pragma unroll
      for (int qq = 0; qq < MAXQ; qq++) Ssc[e][qq] = 0.0;
    }
#pragma unroll
    for (int m = 0; m < NKA; m++) mrate[m] = (SPEC_NKIN > 0) ? st.mnrl_rate[m * ld + cell] : 0.0;

    // RStep state, identical in every warp of the block
    double cumulative = 0.0, dt = target, norm0 = 0.0;
    int ncuts = 0, nconst = 0, nss = 0, nit = 0, nku = 0, its = 0;
    bool done = !live, aborted = false, had_cut = false, need_begin = true, first = true;

    for (;;) {
      // ---- RReact entry (reaction.F90:3829-3850) for cells that start a sub-step
      if (need_begin && (!done || first)) {
#pragma unroll
        for (int q = 0; q < QN; q++) {
          const int i = wid + q * NW;
          if (i >= N) continue;
          double f = 0.0;
          if (i < NAQ) {
            if (!dry) f = psv * st.total[i * ld + cell];
            if (SPEC_NEQSR > 0) f = f + st.total_sorb_eq[i * ld + cell] * vol;
          } else {
            if (!dry) f = 0.0 + st.immobile[(i - NAQ) * ld + cell] * vol;
          }
          fixed[q] = f;
          const int ci = T_cmap[i];
          if (ci >= 0)
            SW(SW_OFF_C + ci) = guess[q];
          else
            cdec[q] = guess[q];
        }
        its = 0;
        need_begin = false;
      }
      first = false;
      if (!done) its++;

      // ---- A1: ionic strength, activity classes (reaction.F90:4553-4612)
      if (SPEC_ACT_UPD) {
        double Ip = 0.0;
#pragma unroll
        for (int q = 0; q < QN; q++) {
          const int i = wid + q * NW;
          if (i >= NAQ) continue;
          const int ci = T_cmap[i];
          const double c = ci >= 0 ? SW(SW_OFF_C + ci) : cdec[q];
          Ip += c * T_z2[i];
        }
        SW(SW_OFF_RED + 2 * wid) = Ip;
        SW(SW_OFF_RED + 2 * wid + 1) = Is_part;
        wt_barrier();
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int w = 0; w < NW; w++) {
          a += SW(SW_OFF_RED + 2 * w);
          b += SW(SW_OFF_RED + 2 * w + 1);
        }
        const double I = 0.5 * (a + b);
        const double sq = sqrt(I);
        I_last = I;
        // every warp keeps its own copy of the class table in the idle Jacobian
#pragma unroll
        for (int qc = 0; qc < SPEC_NCLS; qc++) SW(16 * wid + qc) = wt_lngamma(qc, I, sq);
#pragma unroll
        for (int q = 0; q < QN; q++) {
          const int i = wid + q * NW;
          if (i >= NAQ) continue;
          const int pc = T_pcls[i];
          lngam[q] = pc < 0 ? 0.0 : SW(16 * wid + pc);
        }
      }
      // ---- A2: ln a_i, 1/c_i of the own species
#pragma unroll
      for (int q = 0; q < QN; q++) {
        const int i = wid + q * NW;
        if (i >= NAQ) continue;
        const int ci = T_cmap[i];
        if (ci < 0) continue;
        const double c = SW(SW_OFF_C + ci);
        SW(EXS(ci)) = log(c) + lngam[q];
        SW(SW_OFF_IC + ci) = 1.0 / c;
      }
      wt_barrier();

      // ---- P1: own complexes (RTotalAqueous, reaction.F90:4708-4727)
      {
        double Is = 0.0;
#pragma unroll 2
        for (int k = wid; k < SPEC_NCX; k += NW) {
          double lq = T_cx_lnk[k];
          if (T_cx_h2o[k] != 0.0) lq = lq + T_cx_h2o[k] * ln_act_h2o;
          for (int p = T_cx_ptr[k]; p < T_cx_ptr[k + 1]; p++) lq = lq + T_cx_nu[p] * SW(EXS(T_cx_id[p]));
          const int qc = T_cx_cls[k];
          double lg = 0.0;
          if (SPEC_ACT_UPD) {
            if (qc >= 0) lg = SW(16 * wid + qc);
          } else {
            lg = log(st.sec_act_coef[k * ld + cell]);
          }
          const double sk = exp(lq - lg);
          if (!done) st.sec_molal[k * ld + cell] = sk;
          Is += sk * T_cx_z2[k];
        }
        Is_part = Is;
      }
      wt_barrier();

      // ---- P2: own rows of d(total)/d(free) (reaction.F90:4728-4757, 5775), totals
#pragma unroll
      for (int q = 0; q < QN; q++) {
        ts[q] = 0.0;
        const int i = wid + q * NW;
        if (i >= N) continue;
        const int ci = T_cmap[i];
        if (ci < 0) {
          tot[q] = (i < NAQ) ? cdec[q] * denL : cdec[q];
          continue;
        }
        double *row = W + JX(ci, 0);
#pragma unroll
        for (int j = 0; j < NC; j++) row[j * 32] = 0.0;
        row[ci * 32] = 1.0;
        double t = SW(SW_OFF_C + ci);
        for (int e = T_sp_ptr[ci]; e < T_sp_ptr[ci + 1]; e++) {
          const int k = T_sp_cx[e];
          const double nui = T_sp_nu[e];
          const double sk = __ldcg(st.sec_molal + k * ld + cell);
          t = t + nui * sk;
          for (int p = T_cx_ptr[k]; p < T_cx_ptr[k + 1]; p++) {
            const int cj = T_cx_id[p];
            const double tj = (T_cx_nu[p] * sk) * SW(SW_OFF_IC + cj);
            row[cj * 32] = row[cj * 32] + nui * tj;
          }
        }
        tot[q] = t * denL;
        const double psvd = por * sat * 1000.0 * vol / dt;
#pragma unroll
        for (int j = 0; j < NC; j++) row[j * 32] = (row[j * 32] * denL) * psvd;
        if (dry) {
#pragma unroll
          for (int j = 0; j < NC; j++) row[j * 32] = (j == ci) ? 1.0 : 0.0;
        }
      }
      // ---- equilibrium surface complexation, unit free-site stoichiometry
      // (RTotalSorbEqSurfCplx1, reaction_surf_complex.F90:641-900): evaluated by every
      // warp with the same instructions, applied to the own species
      if (SPEC_NEQSR > 0) {
        const double jscale = vol / dt;
#pragma unroll
        for (int e = 0; e < SPEC_NEQSR; e++) {
          const int r = T_eq[e];
          double dens = T_sr_dens[r];
          if (T_sr_type[r] == 1)
            dens = dens * __ldcg(st.mnrl_volfrac + T_sr_surf[r] * ld + cell);
          else if (T_sr_type[r] == 2)
            dens = dens * spd * (1.0 - por);
          const int c0 = T_sr_ptr[r], nq = T_sr_ptr[r + 1] - c0;
          if (dens < 1.e-40) {
            fsite[e] = 0.0;
#pragma unroll
            for (int qq = 0; qq < MAXQ; qq++) Ssc[e][qq] = 0.0;
          } else {
            double ex[MAXQ], esum = 0.0;
#pragma unroll
            for (int qq = 0; qq < MAXQ; qq++) {
              ex[qq] = 0.0;
              if (qq < nq) {
                const int k = T_sr_cx[c0 + qq];
                double lq = T_sc_lnk[k];
                if (T_sc_h2o[k] != 0.0) lq = lq + T_sc_h2o[k] * ln_act_h2o;
                for (int p = T_sc_ptr[k]; p < T_sc_ptr[k + 1]; p++) lq = lq + T_sc_nu[p] * SW(EXS(T_sc_id[p]));
                ex[qq] = exp(lq);
                esum += ex[qq];
              }
            }
            const double fs = dens / (1.0 + esum);
            fsite[e] = fs;
            double den = 0.0;
#pragma unroll
            for (int qq = 0; qq < MAXQ; qq++) {
              Ssc[e][qq] = ex[qq] * fs;
              if (qq < nq) den += Ssc[e][qq];
            }
            den = den / fs + 1.0;
#pragma unroll
            for (int qq = 0; qq < MAXQ; qq++) {
              if (qq >= nq) continue;
              const int k = T_sr_cx[c0 + qq];
              const double S = Ssc[e][qq];
              const double nuiSx = S / fs;
              for (int p = T_sc_ptr[k]; p < T_sc_ptr[k + 1]; p++) {
                const int i = T_sc_sp[p];
                if (i % NW == wid) wt_add_slot<QN>(ts, i / NW, T_sc_nu[p] * S);
              }
              for (int p2 = T_sc_ptr[k]; p2 < T_sc_ptr[k + 1]; p2++) {
                const int cj = T_sc_id[p2];
                double tmp = 0.0;
#pragma unroll
                for (int q2 = 0; q2 < MAXQ; q2++)
                  if (q2 < nq) tmp += T_sr_dnu[(r * MAXQ + q2) * NC + cj] * Ssc[e][q2];
                const double icj = SW(SW_OFF_IC + cj);
                const double dsx = (-tmp / den) * icj;
                const double t = T_sc_nu[p2] * S * icj + nuiSx * dsx;
                for (int p = T_sc_ptr[k]; p < T_sc_ptr[k + 1]; p++) {
                  const int i = T_sc_sp[p];
                  if (i % NW == wid) {
                    double *a = W + JX(T_sc_id[p], cj);
                    *a = *a + jscale * (T_sc_nu[p] * t);
                  }
                }
              }
            }
          }
        }
      }
      const bool over = its > prm.max_its;
#pragma unroll
      for (int q = 0; q < QN; q++) {
        const int i = wid + q * NW;
        if (i >= N) continue;
        double a = 0.0;
        if (!dry) a = (i < NAQ) ? psv * tot[q] : 0.0 + cdec[q] * vol;
        if (SPEC_NEQSR > 0 && i < NAQ) a = a + ts[q] * vol;
        res[q] = (a - fixed[q]) / dt;
      }
      // ---- kinetic minerals, TST without prefactors (RKineticMineral, reaction_mineral.F90:647-1078)
      if (SPEC_NKIN > 0) {
#pragma unroll
        for (int m = 0; m < SPEC_NKIN; m++) {
          double lq = T_mn_lnk[m];
          if (T_mn_h2o[m] != 0.0) lq = lq + T_mn_h2o[m] * ln_act_h2o;
          for (int p = T_mn_ptr[m]; p < T_mn_ptr[m + 1]; p++) lq = lq + T_mn_nu[p] * SW(EXS(T_mn_id[p]));
          const double QK = exp(lq);
          double aff = 1.0 - QK;
          const double sgn = copysign(1.0, aff);
          bool active = (__ldcg(st.mnrl_volfrac + m * ld + cell) > 0.0 || sgn < 0.0);
          if (T_mn_irr[m] == 1 && sgn < 0.0) active = false;
          if (T_mn_thr[m] > 0.0 && sgn < 0.0 && QK < T_mn_thr[m]) active = false;
          double rate_vol = 0.0;
          if (active) {
            const double lim = T_mn_lim[m];
            if (lim > 0.0) aff = aff / (1.0 + (1.0 - aff) / lim);
            double spr = T_mn_rate[m] * 1.0;
            if (T_mn_eact[m] > 0.0)
              spr = T_mn_rate[m] * exp(T_mn_eact[m] / 8.31446 * (1.0 / (25.0 + 273.15) - 1.0 / (temp + 273.15)));
            double Im_const = -st.mnrl_area[m * ld + cell];
            double Im = Im_const * sgn * fabs(aff) * spr;
            rate_vol = Im;
            if (!dry) {
              Im_const = Im_const * vol;
              Im = Im * vol;
              const double dIm_dQK = -Im_const * spr;
              double dfac;
              if (lim > 0.0) {
                const double den = 1.0 + (1.0 - aff) / lim;
                dfac = dIm_dQK * (1.0 + QK / lim / den) * QK * denL / den;
              } else {
                dfac = dIm_dQK * QK * denL;
              }
              for (int p = T_mn_ptr[m]; p < T_mn_ptr[m + 1]; p++) {
                const int i = T_mn_sp[p];
                if (i % NW == wid) wt_add_slot<QN>(res, i / NW, T_mn_nu[p] * Im);
              }
              for (int p2 = T_mn_ptr[m]; p2 < T_mn_ptr[m + 1]; p2++) {
                const int cj = T_mn_id[p2];
                const double t = dfac * (T_mn_nu[p2] * SW(SW_OFF_IC + cj));
                for (int p = T_mn_ptr[m]; p < T_mn_ptr[m + 1]; p++) {
                  const int i = T_mn_sp[p];
                  if (i % NW == wid) {
                    double *a = W + JX(T_mn_id[p], cj);
                    *a = *a + T_mn_nu[p] * t;
                  }
                }
              }
            }
          }
          mrate[m] = rate_vol;
        }
      }
      {
        double mabs = 0.0, ss = 0.0;
#pragma unroll
        for (int q = 0; q < QN; q++)
          if (wid + q * NW < N) {
            mabs = fmax(mabs, fabs(res[q]));
            ss += res[q] * res[q];
          }
        SW(SW_OFF_RED + 2 * wid) = mabs;
        SW(SW_OFF_RED + 2 * wid + 1) = ss;
      }
      wt_barrier();

      // ---- P3: convergence on the residual (reaction.F90:3925-3950)
      bool conv;
      {
        double mabs = 0.0, ss = 0.0;
#pragma unroll
        for (int w = 0; w < NW; w++) {
          mabs = fmax(mabs, SW(SW_OFF_RED + 2 * w));
          ss += SW(SW_OFF_RED + 2 * w + 1);
        }
        const double nrm = sqrt(ss);
        if (its == 1) norm0 = nrm;
        const double rel = nrm / norm0;
        conv = (mabs < prm.tol_res) || (rel < prm.tol_relres);
      }
      const bool need_solve = !done && !over && !conv;
      bool fail = !done && over;
      bool solve_error = false;

      if (__any_sync(FULL, need_solve)) {
        // ---- P4: RSolve scaling of the own rows (reaction.F90:5457-5516); species outside
        // the matrix have a diagonal row: their update is res / J
        bool bad = false;
        double b[QN], xdec[QN];
#pragma unroll
        for (int q = 0; q < QN; q++) {
          b[q] = xdec[q] = 0.0;
          const int i = wid + q * NW;
          if (i >= N) continue;
          const int ci = T_cmap[i];
          if (ci < 0) {
            double Jd = (i < NAQ) ? (1.0 * denL) * (por * sat * 1000.0 * vol / dt) : vol / dt;
            if (dry) Jd = 1.0;
            const double nm = 1.0 / fmax(1.0, fabs(Jd));
            double a = Jd * nm;
            if (SPEC_USE_LOG) a *= cdec[q];
            if (!(fabs(a) > 0.0)) bad = true;
            xdec[q] = (res[q] * nm) / a;
          } else {
            double *rowp = W + JX(ci, 0);
            double row[NC];
            double m = 0.0;
#pragma unroll
            for (int j = 0; j < NC; j++) {
              row[j] = rowp[j * 32];
              const double av = fabs(row[j]);
              m = av > m ? av : m;
            }
            const double nm = 1.0 / fmax(1.0, m);
            b[q] = res[q] * nm;
            double m2 = 0.0;
#pragma unroll
            for (int j = 0; j < NC; j++) {
              double v = row[j] * nm;
              if (SPEC_USE_LOG) v *= SW(SW_OFF_C + j);
              rowp[j * 32] = v;
              const double av = fabs(v);
              m2 = av > m2 ? av : m2;
            }
            if (!(m2 > 0.0)) bad = true;
            rowp[NC * 32] = 1. / m2;
          }
        }
        SW(SW_OFF_RED + 2 * NW + wid) = bad ? 1.0 : 0.0;
        wt_barrier();  // rows were scaled by their species' owners, the LU owns them by position
#pragma unroll
        for (int w = 0; w < NW; w++) solve_error = solve_error || (SW(SW_OFF_RED + 2 * NW + w) != 0.0);
        // pivot candidates of column 0 -> buffer 0 (the 1/c region, idle until the next pass)
        {
          double best = -1.0;
          int bi = -1;
          for (int i = wid; i < NC; i += NW) {
            const double dum = SW(EXS(i)) * fabs(W[JX(i, 0)]);
            if (dum >= best) {
              best = dum;
              bi = i;
            }
          }
          SW(SW_OFF_IC + 2 * wid) = best;
          SW(SW_OFF_IC + 2 * wid + 1) = (double)bi;
        }
        unsigned long long perm = 0xFEDCBA9876543210ull;  // logical position -> row

        // ---- P5: right-looking LU, one barrier per column
#pragma unroll
        for (int k = 0; k < NC; k++) {
          const int bufo = (k & 1) ? SW_OFF_RED : SW_OFF_IC;
          const int bufn = (k & 1) ? SW_OFF_IC : SW_OFF_RED;
          wt_barrier();
          // the reference scans i = k..n-1 with `>=': the largest value wins, the last one among equals
          double aamax = -1.0;
          int imax = k;
#pragma unroll
          for (int w = 0; w < NW; w++) {
            const double dum = SW(bufo + 2 * w);
            const int bi = (int)SW(bufo + 2 * w + 1);
            if (bi >= 0 && (dum > aamax || (dum == aamax && bi > imax))) {
              aamax = dum;
              imax = bi;
            }
          }
          {
            const unsigned long long x = ((perm >> (4 * k)) ^ (perm >> (4 * imax))) & 15ull;
            perm ^= (x << (4 * k)) | (x << (4 * imax));
          }
          double *pr = W + (int)((perm >> (4 * k)) & 15ull) * (SPEC_JS * 32);
          double pv = pr[k * 32];
          if (pv == 0.0) {
            pv = 1.0e-20;
            if (k % NW == wid) pr[k * 32] = pv;
          }
          if (k != NC - 1) {
            const double dum = 1.0 / pv;
            double prow[NC];
#pragma unroll
            for (int j = 0; j < NC; j++)
              if (j > k) prow[j] = pr[j * 32];
            double best = -1.0;
            int bi = -1;
            // own logical rows below the pivot: i = first, first + W, ...
            const int first = k + 1 + ((wid - (k + 1)) % NW + NW) % NW;
#pragma unroll
            for (int t = 0; t < (NC - k - 1 + NW - 1) / NW; t++) {
              const int i = first + t * NW;
              if (i < NC) {
                double *r = W + (int)((perm >> (4 * i)) & 15ull) * (SPEC_JS * 32);
                const double l = r[k * 32] * dum;
                r[k * 32] = l;
                double nxt = 0.0;
#pragma unroll
                for (int j = 0; j < NC; j++)
                  if (j > k) {
                    double v = r[j * 32];
                    v -= l * prow[j];
                    r[j * 32] = v;
                    if (j == k + 1) nxt = v;
                  }
                const double dd = r[NC * 32] * fabs(nxt);
                if (dd >= best) {
                  best = dd;
                  bi = i;
                }
              }
            }
            SW(bufn + 2 * wid) = best;
            SW(bufn + 2 * wid + 1) = (double)bi;
          }
        }
        wt_barrier();
        // ---- P6: right-hand side into the rows' extra column, substitution (every warp)
#pragma unroll
        for (int q = 0; q < QN; q++) {
          const int i = wid + q * NW;
          if (i >= N) continue;
          const int ci = T_cmap[i];
          if (ci >= 0) SW(EXS(ci)) = b[q];
        }
        wt_barrier();
        {
          double y[NC];
#pragma unroll
          for (int k = 0; k < NC; k++) {
            const double *r = W + (int)((perm >> (4 * k)) & 15ull) * (SPEC_JS * 32);
            double sum = r[NC * 32];
#pragma unroll
            for (int m = 0; m < NC; m++)
              if (m < k) sum -= r[m * 32] * y[m];
            y[k] = sum;
          }
#pragma unroll
          for (int k = NC - 1; k >= 0; k--) {
            const double *r = W + (int)((perm >> (4 * k)) & 15ull) * (SPEC_JS * 32);
            double sum = y[k];
#pragma unroll
            for (int m = 0; m < NC; m++)
              if (m > k) sum -= r[m * 32] * y[m];
            y[k] = sum / r[k * 32];
          }
          // the same values from every warp; each warp reads back only what it wrote
#pragma unroll
          for (int k = 0; k < NC; k++) SW(SW_OFF_IC + k) = y[k];
        }
        // ---- P7: update of the own species (reaction.F90:3985-4032)
        double cn[QN], maxrel = -1.0;
#pragma unroll
        for (int q = 0; q < QN; q++) {
          cn[q] = 0.0;
          const int i = wid + q * NW;
          if (i >= N) continue;
          const int ci = T_cmap[i];
          const double c = ci >= 0 ? SW(SW_OFF_C + ci) : cdec[q];
          double u = ci >= 0 ? SW(SW_OFF_IC + ci) : xdec[q];
          u = copysign(1.0, u) * fmin(fabs(u), prm.max_dlnC);
          cn[q] = c * exp(-u);
          const double v = fabs((cn[q] - c) / c);
          if (!isnan(v)) maxrel = fmax(maxrel, v);
        }
        SW(SW_OFF_RED + 2 * NW + wid) = maxrel;
        wt_barrier();
        maxrel = -1.0;
#pragma unroll
        for (int w = 0; w < NW; w++) maxrel = fmax(maxrel, SW(SW_OFF_RED + 2 * NW + w));
        const bool conv2 = (maxrel >= 0.0) && (maxrel < prm.tol_relchange);
        if (need_solve) {
          if (solve_error) {
            fail = true;
          } else if (conv2) {
            conv = true;
          } else {
#pragma unroll
            for (int q = 0; q < QN; q++) {
              const int i = wid + q * NW;
              if (i >= N) continue;
              const int ci = T_cmap[i];
              if (ci >= 0)
                SW(SW_OFF_C + ci) = cn[q];
              else
                cdec[q] = cn[q];
            }
          }
        }
      }
      // every exchange slot is quiet before the next pass writes it
      wt_barrier();

      // ---- outcome of this pass for the cell (RStep, reaction.F90:3655-3700)
      if (!done) {
        if (fail) {
          nit += its;
          // its > max: total / immobile keep their values in HBM, total_sorb_eq does not;
          // solve error: no restore (reaction.F90:3964-3967)
#pragma unroll
          for (int q = 0; q < QN; q++) {
            const int i = wid + q * NW;
            if (i >= N) continue;
            if (i < NAQ) {
              if (SPEC_NEQSR > 0) st.total_sorb_eq[i * ld + cell] = ts[q];
              if (solve_error && !over) st.total[i * ld + cell] = tot[q];
            } else if (solve_error && !over) {
              st.immobile[(i - NAQ) * ld + cell] = cdec[q];
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
          for (int q = 0; q < QN; q++) {
            const int i = wid + q * NW;
            if (i >= N) continue;
            const int ci = T_cmap[i];
            const double c = ci >= 0 ? SW(SW_OFF_C + ci) : cdec[q];
            if (i < NAQ) {
              st.total[i * ld + cell] = tot[q];
              if (SPEC_NEQSR > 0) st.total_sorb_eq[i * ld + cell] = ts[q];
            } else {
              st.immobile[(i - NAQ) * ld + cell] = c;
            }
            guess[q] = c;
          }
          // RUpdateKineticState with the rates of the converged iterate
          if (SPEC_NKIN > 0) {
#pragma unroll
            for (int m = 0; m < SPEC_NKIN; m++)
              if (m % NW == wid) {
                double vf = st.mnrl_volfrac[m * ld + cell] + mrate[m] * T_mn_vol[m] * dt;
                if (vf < 0.0) vf = 0.0;
                st.mnrl_volfrac[m * ld + cell] = vf;
              }
            nku++;
          }
          cumulative += dt;
          nss++;
          nconst++;
          if (nconst >= 4) {
            ncuts--;
            dt = fmin(2.0 * dt, target - cumulative);
          }
          if (cumulative >= target)
            done = true;
          else
            need_begin = true;
        }
        if (done) {
          // ---- publish the cell (reaction.F90:3700-3738) now: later passes of the group
          // keep recomputing this lane's activity / sorption state
#pragma unroll
          for (int q = 0; q < QN; q++) {
            const int i = wid + q * NW;
            if (i >= N) continue;
            if (i < NAQ) {
              const int ci = T_cmap[i];
              const double c = ci >= 0 ? SW(SW_OFF_C + ci) : cdec[q];
              st.pri_molal[i * ld + cell] = aborted ? c : guess[q];
              if (SPEC_ACT_UPD) st.pri_act_coef[i * ld + cell] = exp(lngam[q]);
            }
            if (!aborted && ((small_mask >> q) & 1u)) {
              if (i < NAQ)
                st.total[i * ld + cell] = small_val[q];
              else
                st.immobile[(i - NAQ) * ld + cell] = small_val[q];
            }
          }
          if (SPEC_ACT_UPD) {
            const double sq = sqrt(I_last);
            for (int k = wid; k < SPEC_NCX; k += NW) {
              const int qc = T_cx_cls[k];
              st.sec_act_coef[k * ld + cell] = qc < 0 ? 1.0 : exp(wt_lngamma(qc < 0 ? 0 : qc, I_last, sq));
            }
          }
          if (wid == 0) {
#pragma unroll
            for (int e = 0; e < SPEC_NEQSR; e++) {
              const int r = T_eq[e];
              st.free_site[r * ld + cell] = fsite[e];
              if (st.eqsrfcplx_conc) {
#pragma unroll
                for (int qq = 0; qq < MAXQ; qq++)
                  if (qq < T_sr_ptr[r + 1] - T_sr_ptr[r])
                    st.eqsrfcplx_conc[T_sr_cx[T_sr_ptr[r] + qq] * ld + cell] = Ssc[e][qq];
              }
            }
#pragma unroll
            for (int m = 0; m < SPEC_NKIN; m++) st.mnrl_rate[m * ld + cell] = mrate[m];
          }
        }
      }
      if (__all_sync(FULL, done)) break;
    }

    if (wid == 0 && inrange) {
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
    // the next group's first pass writes the exchange slots: everyone is past its reads
    wt_barrier();
  }

  if (wid == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      l_active += __shfl_xor_sync(FULL, l_active, o);
      l_its += __shfl_xor_sync(FULL, l_its, o);
      l_cut += __shfl_xor_sync(FULL, l_cut, o);
      long long f = __shfl_xor_sync(FULL, l_first, o);
      if (f >= 0 && (l_first < 0 || f < l_first)) l_first = f;
      l_maxits = max(l_maxits, __shfl_xor_sync(FULL, l_maxits, o));
      l_maxkin = max(l_maxkin, __shfl_xor_sync(FULL, l_maxkin, o));
      l_maxerr = max(l_maxerr, __shfl_xor_sync(FULL, l_maxerr, o));
      l_maxsub = max(l_maxsub, __shfl_xor_sync(FULL, l_maxsub, o));
    }
    if (lane == 0) {
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
}

extern "C" __global__ void __launch_bounds__(32 * SPEC_W, SPEC_MINBLOCKS)
    pfrx_spec_kernel(DevState st, long long ncell, double tran_dt, SpecParams prm, DevSummary *summ) {
  extern __shared__ double smem[];
  wt_run(st, ncell, tran_dt, prm, summ, smem + (threadIdx.x & 31), (int)(threadIdx.x >> 5));
}
