// pfrx_specw.cuh -- network-specialised kernel, SPEC_W warps per group of 32 cells.
//
// Same arithmetic as pfrx_spec.cuh (one thread per cell), but the per-cell work of
// a Newton iteration is divided between the SPEC_W warps of a block so that an
// SM, whose shared memory holds the Jacobians of only ~130 cells of a 13-unknown
// network, still has SPEC_W times as many warps to hide latency with.  Lane l of
// every warp of the block works on the same cell; warp w ("role" w) owns
//   * the species i with spec_owner(i) == w: their concentration, fixed
//     accumulation, guess, total, residual, Jacobian ROW and LU rows at the
//     logical positions i = w (mod SPEC_W);
//   * the secondary complexes k with spec_cx_owner(k) == w: their exp().
// Warps exchange through the block's shared slice (ln a_j, 1/c_j, c_j, partial
// norms, pivot candidates) and, for the complex concentrations, through
// rt_auxvar%sec_molal in HBM/L2, which the kernel has to write anyway.
//
// The LU is the reference's (utility.F90:597-735) in right-looking order: the
// element (i,j) receives the same  a_ij - l_i0 u_0j - l_i1 u_1j ...  sequence of
// fused multiply-adds, and the pivot of column k is chosen from the same values
// with the same `>=' rule, so the factors are bit-identical to Crout's.
//
// Control flow is lock-step: every pass of the main loop is one Newton iteration
// for all 32 cells of the group; a cell that has finished keeps executing with
// its stores disabled until the whole group is done.  Every decision that steers
// barriers is computed from exchanged values by role-independent code, so the
// warps of a block cannot disagree.
//
// Generated per network (specialize.py): the SPEC_* macros, spec_owner(),
// spec_cx_owner(), spec_cmap(), spec_sp_of(), and per role the straight-line
// functions specw_activity<w>, specw_complexes<w>, specw_rows<w>,
// specw_apply<w>, plus the role-independent __noinline__ specw_eval().
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
#define EXS(ci) ((ci) * SPEC_JS + SPEC_NC)  // a row's extra column: ln a_i / scaling factor / rhs

static_assert(SPEC_NC >= 2 * SPEC_W, "the exchange buffers alias regions of NC slots");

extern "C" {
__device__ const unsigned long long pfrx_spec_sig = SPEC_SIG;
// {N, shared doubles per block, threads per block, min blocks per SM, cells per block}
__device__ const int pfrx_spec_info[5] = {SPEC_N, SPEC_SLOTS * 32, 32 * SPEC_W, SPEC_MINBLOCKS, 32};
}

struct CellW {
  double den_kg, sat, temp, por, vol, spd, ln_act_h2o;
  double Is_part;  // sum z^2 m over this role's complexes (latest RTotal)
  double lgcls[SPEC_NCLS > 0 ? SPEC_NCLS : 1];
  double lngam[SPEC_N];  // own species only
  bool dry;
};

// what specw_eval() hands to every role: sorption and mineral terms of the
// current iterate, evaluated by ONE piece of code (bit-identical in all warps)
struct EvalOut {
  double fsite[SPEC_NSRFRXN > 0 ? SPEC_NSRFRXN : 1];
  double S[SPEC_NSRFCPLX > 0 ? SPEC_NSRFCPLX : 1];  // surface complex concentrations
  double nuis[SPEC_NSRFCPLX > 0 ? SPEC_NSRFCPLX : 1];  // S / free sites (0 when the site is absent)
  double dsx[SPEC_NSRFRXN > 0 ? SPEC_NSRFRXN * SPEC_NC : 1];
  double Im[SPEC_NKIN > 0 ? SPEC_NKIN : 1];     // rate * volume (0 when inactive)
  double dfac[SPEC_NKIN > 0 ? SPEC_NKIN : 1];   // dIm/dQK * QK * den (0 when inactive)
  double mrate[SPEC_NKIN > 0 ? SPEC_NKIN : 1];  // rate per bulk volume
};

template <int WID>
__device__ __forceinline__ void specw_activity(double I, CellW &s);
template <int WID>
__device__ __forceinline__ void specw_complexes(CellW &s, const double *W, double *sec_out, long long ld, bool store);
template <int WID>
__device__ __forceinline__ void specw_rows(const CellW &s, double *W, const double *sec_in, long long ld, double dt,
                                           double (&tot)[SPEC_N]);
template <int WID>
__device__ __forceinline__ void specw_apply(const CellW &s, const EvalOut &e, double *W, double jscale,
                                            double (&ts)[SPEC_N], double (&res)[SPEC_N], bool minerals);
__device__ __noinline__ void specw_eval(const double *W, const DevState &st, long long cell, double den_kg, double por,
                                        double vol, double spd, double temp, double ln_act_h2o, bool apply,
                                        EvalOut *out);

// all warps of the block, whatever role code they are in
__device__ __forceinline__ void specw_barrier() { asm volatile("bar.sync 0;" ::: "memory"); }

// ---- the role ---------------------------------------------------------------------------
template <int WID>
__device__ __forceinline__ void specw_role(const DevState &st, const long long ncell, const double target,
                                           const SpecParams &prm, DevSummary *summ, double *W) {
  constexpr int N = SPEC_N, NAQ = SPEC_NAQ, NC = SPEC_NC, NW = SPEC_W, NCA = NC > 0 ? NC : 1;
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

    CellW s;
    s.den_kg = st.den_kg[cell];
    s.sat = st.sat[cell];
    s.temp = st.temp[cell];
    s.por = st.porosity[cell];
    s.vol = st.volume[cell];
    s.spd = st.soil_particle_density ? st.soil_particle_density[cell] : 0.0;
    s.ln_act_h2o = st.ln_act_h2o ? st.ln_act_h2o[cell] : 0.0;
    s.dry = s.sat < prm.min_sat;
    const double psv = s.por * s.sat * 1000.0 * s.vol;
    {
      double Is = 0.0;
#pragma unroll 4
      for (int k = 0; k < SPEC_NCX; k++)
        if (spec_cx_owner(k) == WID) Is += st.sec_molal[k * ld + cell] * spec_cx_z2(k);
      s.Is_part = Is;
    }
#pragma unroll
    for (int k = 0; k < (SPEC_NCLS > 0 ? SPEC_NCLS : 1); k++) s.lgcls[k] = 0.0;

    // own species: guess, clamped totals (RStep, reaction.F90:3633-3650)
    double guess[N], fixed[N], cdec[N], small_val[N], tot[N], ts[N], res[N];
    unsigned small_mask = 0u;
#pragma unroll
    for (int i = 0; i < N; i++) {
      guess[i] = fixed[i] = cdec[i] = small_val[i] = tot[i] = ts[i] = res[i] = 0.0;
      s.lngam[i] = 0.0;
      if (spec_owner(i) != WID) continue;
      if (i < NAQ) {
        s.lngam[i] = log(st.pri_act_coef[i * ld + cell]);
        guess[i] = st.pri_molal[i * ld + cell];
        double t = st.total[i * ld + cell];
        if (t <= 1.e-40) {
          small_mask |= 1u << i;
          small_val[i] = t;
          if (live) st.total[i * ld + cell] = 1.e-40;
        }
      } else {
        double t = st.immobile[(i - NAQ) * ld + cell];
        guess[i] = t;
        if (t <= 1.e-40) {
          small_mask |= 1u << i;
          small_val[i] = t;
          if (live) st.immobile[(i - NAQ) * ld + cell] = 1.e-40;
        }
      }
    }
    EvalOut ev;
#pragma unroll
    for (int k = 0; k < SPEC_NSRFRXN; k++) ev.fsite[k] = st.free_site[k * ld + cell];
#pragma unroll
    for (int k = 0; k < SPEC_NSRFCPLX; k++) ev.S[k] = 0.0;
#pragma unroll
    for (int k = 0; k < SPEC_NKIN; k++) {
      ev.mrate[k] = st.mnrl_rate[k * ld + cell];
      ev.Im[k] = ev.dfac[k] = 0.0;
    }

    // RStep state, identical in every warp of the block
    double cumulative = 0.0, dt = target, norm0 = 0.0;
    int ncuts = 0, nconst = 0, nss = 0, nit = 0, nku = 0, its = 0;
    bool done = !live, aborted = false, had_cut = false, need_begin = true;

    for (;;) {
      // ---- RReact entry (reaction.F90:3829-3850) for cells that start a sub-step
      if (need_begin && (!done || nss + ncuts + its == 0)) {
#pragma unroll
        for (int i = 0; i < N; i++) {
          if (spec_owner(i) != WID) continue;
          double f = 0.0;
          if (i < NAQ) {
            if (!s.dry) f = psv * st.total[i * ld + cell];
            if (SPEC_NEQSR > 0) f = f + st.total_sorb_eq[i * ld + cell] * s.vol;
          } else {
            if (!s.dry) f = 0.0 + st.immobile[(i - NAQ) * ld + cell] * s.vol;
          }
          fixed[i] = f;
          if (spec_cmap(i) >= 0)
            SW(SW_OFF_C + spec_cmap(i)) = guess[i];
          else
            cdec[i] = guess[i];
        }
        its = 0;
        need_begin = false;
      }
      if (!done) its++;

      // ---- A1: ionic strength partials
      if (SPEC_ACT_UPD) {
        double Ip = 0.0;
#pragma unroll
        for (int i = 0; i < NAQ; i++)
          if (spec_owner(i) == WID && spec_z2(i) != 0.0)
            Ip += (spec_cmap(i) >= 0 ? SW(SW_OFF_C + (spec_cmap(i) >= 0 ? spec_cmap(i) : 0)) : cdec[i]) * spec_z2(i);
        SW(SW_OFF_RED + 2 * WID) = Ip;
        SW(SW_OFF_RED + 2 * WID + 1) = s.Is_part;
        specw_barrier();
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int w = 0; w < NW; w++) {
          a += SW(SW_OFF_RED + 2 * w);
          b += SW(SW_OFF_RED + 2 * w + 1);
        }
        specw_activity<WID>(0.5 * (a + b), s);
      }
      // ---- A2: ln a_i, 1/c_i of the own species
#pragma unroll
      for (int i = 0; i < NAQ; i++) {
        if (spec_owner(i) != WID || spec_cmap(i) < 0) continue;
        const int ci = spec_cmap(i) >= 0 ? spec_cmap(i) : 0;
        double c = SW(SW_OFF_C + ci);
        SW(EXS(ci)) = log(c) + s.lngam[i];
        SW(SW_OFF_IC + ci) = 1.0 / c;
      }
      specw_barrier();

      // ---- P1: own complexes
      specw_complexes<WID>(s, W, st.sec_molal + cell, ld, !done);
      specw_barrier();

      // ---- P2: own rows of d(total)/d(free), totals, sorption, minerals, residual
      specw_rows<WID>(s, W, st.sec_molal + cell, ld, dt, tot);
#pragma unroll
      for (int i = 0; i < N; i++) {
        ts[i] = 0.0;
        if (spec_owner(i) == WID && spec_cmap(i) < 0) tot[i] = (i < NAQ) ? cdec[i] * (s.den_kg * 1.e-3) : cdec[i];
      }
      if (SPEC_NEQSR > 0 || SPEC_NKIN > 0)
        specw_eval(W, st, cell, s.den_kg, s.por, s.vol, s.spd, s.temp, s.ln_act_h2o, !s.dry, &ev);
      if (SPEC_NEQSR > 0) specw_apply<WID>(s, ev, W, s.vol / dt, ts, res, false);
      const bool over = its > prm.max_its;
#pragma unroll
      for (int i = 0; i < N; i++) {
        if (spec_owner(i) != WID) continue;
        double a = 0.0;
        if (!s.dry) a = (i < NAQ) ? psv * tot[i] : 0.0 + cdec[i] * s.vol;
        if (SPEC_NEQSR > 0 && i < NAQ) a = a + ts[i] * s.vol;
        res[i] = (a - fixed[i]) / dt;
      }
      if (SPEC_NKIN > 0) specw_apply<WID>(s, ev, W, 0.0, ts, res, true);
      {
        double mabs = 0.0, ss = 0.0;
#pragma unroll
        for (int i = 0; i < N; i++)
          if (spec_owner(i) == WID) {
            mabs = fmax(mabs, fabs(res[i]));
            ss += res[i] * res[i];
          }
        SW(SW_OFF_RED + 2 * WID) = mabs;
        SW(SW_OFF_RED + 2 * WID + 1) = ss;
      }
      specw_barrier();

      // ---- P3: convergence on the residual (role-independent arithmetic)
      bool conv;
      {
        double mabs = 0.0, ss = 0.0;
#pragma unroll
        for (int w = 0; w < NW; w++) {
          mabs = fmax(mabs, SW(SW_OFF_RED + 2 * w));
          ss += SW(SW_OFF_RED + 2 * w + 1);
        }
        double nrm = sqrt(ss);
        if (its == 1) norm0 = nrm;
        double rel = nrm / norm0;
        conv = (mabs < prm.tol_res) || (rel < prm.tol_relres);
      }
      const bool need_solve = !done && !over && !conv;
      bool fail = !done && over;
      bool solve_error = false;

      if (__any_sync(FULL, need_solve)) {
        // ---- P4: RSolve scaling of the own rows (reaction.F90:5457-5516), decoupled species
        bool bad = false;
        double b[NCA], x[N];
#pragma unroll
        for (int i = 0; i < N; i++) {
          x[i] = 0.0;
          if (spec_owner(i) != WID) continue;
          if (spec_cmap(i) < 0) {
            double Jd = (i < NAQ) ? (1.0 * (s.den_kg * 1.e-3)) * (s.por * s.sat * 1000.0 * s.vol / dt) : s.vol / dt;
            if (s.dry) Jd = 1.0;
            double nm = 1.0 / fmax(1.0, fabs(Jd));
            double a = Jd * nm;
            if (SPEC_USE_LOG) a *= cdec[i];
            if (!(fabs(a) > 0.0)) bad = true;
            x[i] = (res[i] * nm) / a;
          } else {
            const int ci = spec_cmap(i) >= 0 ? spec_cmap(i) : 0;
            double row[NCA];
            double m = 0.0;
#pragma unroll
            for (int j = 0; j < NC; j++) {
              row[j] = W[JX(ci, j)];
              double av = fabs(row[j]);
              m = av > m ? av : m;
            }
            double nm = 1.0 / fmax(1.0, m);
            b[ci] = res[i] * nm;
            double m2 = 0.0;
#pragma unroll
            for (int j = 0; j < NC; j++) {
              double v = row[j] * nm;
              if (SPEC_USE_LOG) v *= SW(SW_OFF_C + j);
              W[JX(ci, j)] = v;
              double av = fabs(v);
              m2 = av > m2 ? av : m2;
            }
            if (!(m2 > 0.0)) bad = true;
            SW(EXS(ci)) = 1. / m2;
          }
        }
        SW(SW_OFF_RED + 2 * NW + WID) = bad ? 1.0 : 0.0;
        // pivot candidates of column 0 from the own rows -> buffer 0 (the 1/c region)
        int ro[NCA];
#pragma unroll
        for (int i = 0; i < NC; i++) ro[i] = JX(i, 0);
        {
          double best = -1.0;
          int bi = -1;
#pragma unroll
          for (int i = 0; i < NC; i++)
            if (i % NW == WID) {
              double dum = SW(EXS(i)) * fabs(W[JX(i, 0)]);
              if (dum >= best) {
                best = dum;
                bi = i;
              }
            }
          SW(SW_OFF_IC + 2 * WID) = best;
          SW(SW_OFF_IC + 2 * WID + 1) = (double)bi;
        }
        specw_barrier();
#pragma unroll
        for (int w = 0; w < NW; w++) solve_error = solve_error || (SW(SW_OFF_RED + 2 * NW + w) != 0.0);

        // ---- P5: right-looking LU, one barrier per column
#pragma unroll
        for (int k = 0; k < NC; k++) {
          const int bufo = (k & 1) ? SW_OFF_RED : SW_OFF_IC;
          const int bufn = (k & 1) ? SW_OFF_IC : SW_OFF_RED;
          if (k > 0) specw_barrier();
          // the reference scans i = k..n-1 with `>=': the last maximum wins
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
          const int rk_old = ro[k];
          int rmax = rk_old;
#pragma unroll
          for (int i = 0; i < NC; i++)
            if (i > k) {
              bool p = (i == imax);
              rmax = p ? ro[i] : rmax;
              ro[i] = p ? rk_old : ro[i];
            }
          ro[k] = rmax;
          const double *pr = W + rmax;
          double pv = pr[k * 32];
          if (pv == 0.0) {
            pv = 1.0e-20;
            if (k % NW == WID) W[rmax + k * 32] = pv;
          }
          if (k != NC - 1) {
            const double dum = 1.0 / pv;
            double prow[NCA];
#pragma unroll
            for (int j = 0; j < NC; j++)
              if (j > k) prow[j] = pr[j * 32];
            double best = -1.0;
            int bi = -1;
#pragma unroll
            for (int i = 0; i < NC; i++)
              if (i > k && i % NW == WID) {
                double *r = W + ro[i];
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
                double dd = r[NC * 32] * fabs(nxt);
                if (dd >= best) {
                  best = dd;
                  bi = i;
                }
              }
            SW(bufn + 2 * WID) = best;
            SW(bufn + 2 * WID + 1) = (double)bi;
          }
        }
        specw_barrier();
        // ---- P6: right-hand side into the rows' extra column, substitution (replicated)
#pragma unroll
        for (int i = 0; i < NC; i++)
          if (spec_owner(spec_sp_of(i)) == WID) SW(EXS(i)) = b[i];
        specw_barrier();
        {
          double y[NCA];
#pragma unroll
          for (int k = 0; k < NC; k++) {
            const double *r = W + ro[k];
            double sum = r[NC * 32];
#pragma unroll
            for (int m = 0; m < NC; m++)
              if (m < k) sum -= r[m * 32] * y[m];
            y[k] = sum;
          }
#pragma unroll
          for (int k = NC - 1; k >= 0; k--) {
            const double *r = W + ro[k];
            double sum = y[k];
#pragma unroll
            for (int m = 0; m < NC; m++)
              if (m > k) sum -= r[m * 32] * y[m];
            y[k] = sum / r[k * 32];
          }
#pragma unroll
          for (int k = 0; k < NC; k++)
            if (spec_owner(spec_sp_of(k)) == WID) x[spec_sp_of(k)] = y[k];
        }
        // ---- P7: update of the own species (reaction.F90:3985-4032)
        double cn[N], maxrel = -1.0;
#pragma unroll
        for (int i = 0; i < N; i++) {
          cn[i] = 0.0;
          if (spec_owner(i) != WID) continue;
          const double c = spec_cmap(i) >= 0 ? SW(SW_OFF_C + (spec_cmap(i) >= 0 ? spec_cmap(i) : 0)) : cdec[i];
          double u = x[i];
          u = copysign(1.0, u) * fmin(fabs(u), prm.max_dlnC);
          cn[i] = c * exp(-u);
          double v = fabs((cn[i] - c) / c);
          if (!isnan(v)) maxrel = fmax(maxrel, v);
        }
        SW(SW_OFF_RED + 2 * NW + WID) = maxrel;
        specw_barrier();
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
            for (int i = 0; i < N; i++) {
              if (spec_owner(i) != WID) continue;
              if (spec_cmap(i) >= 0)
                SW(SW_OFF_C + (spec_cmap(i) >= 0 ? spec_cmap(i) : 0)) = cn[i];
              else
                cdec[i] = cn[i];
            }
          }
        }
      }
      // every exchange slot is quiet before the next pass writes it
      specw_barrier();

      // ---- outcome of this pass for the cell
      if (!done) {
        if (fail) {
          nit += its;
          // its > max: total / immobile keep their values in HBM, total_sorb_eq does not;
          // solve error: no restore (reaction.F90:3964-3967)
#pragma unroll
          for (int i = 0; i < N; i++) {
            if (spec_owner(i) != WID) continue;
            if (i < NAQ) {
              if (SPEC_NEQSR > 0) st.total_sorb_eq[i * ld + cell] = ts[i];
              if (solve_error && !over) st.total[i * ld + cell] = tot[i];
            } else if (solve_error && !over) {
              st.immobile[(i - NAQ) * ld + cell] = cdec[i];
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
            if (spec_owner(i) != WID) continue;
            const double c = spec_cmap(i) >= 0 ? SW(SW_OFF_C + (spec_cmap(i) >= 0 ? spec_cmap(i) : 0)) : cdec[i];
            if (i < NAQ) {
              st.total[i * ld + cell] = tot[i];
              if (SPEC_NEQSR > 0) st.total_sorb_eq[i * ld + cell] = ts[i];
            } else {
              st.immobile[(i - NAQ) * ld + cell] = c;
            }
            guess[i] = c;
          }
          // RUpdateKineticState with the rates of the converged iterate
          if (SPEC_NKIN > 0) {
#pragma unroll
            for (int m = 0; m < SPEC_NKIN; m++)
              if (m % NW == WID) {
                double vf = st.mnrl_volfrac[m * ld + cell] + ev.mrate[m] * spec_mn_vol(m) * dt;
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
          // ---- publish the cell (reaction.F90:3700-3738) now: later passes of the
          // group keep recomputing this lane's activity / sorption state
#pragma unroll
          for (int i = 0; i < N; i++) {
            if (spec_owner(i) != WID) continue;
            if (i < NAQ) {
              const double c = spec_cmap(i) >= 0 ? SW(SW_OFF_C + (spec_cmap(i) >= 0 ? spec_cmap(i) : 0)) : cdec[i];
              st.pri_molal[i * ld + cell] = aborted ? c : guess[i];
              if (SPEC_ACT_UPD) st.pri_act_coef[i * ld + cell] = exp(s.lngam[i]);
            }
            if (!aborted && ((small_mask >> i) & 1u)) {
              if (i < NAQ)
                st.total[i * ld + cell] = small_val[i];
              else
                st.immobile[(i - NAQ) * ld + cell] = small_val[i];
            }
          }
          if (SPEC_ACT_UPD) {
#pragma unroll 4
            for (int k = 0; k < SPEC_NCX; k++) {
              if (spec_cx_owner(k) != WID) continue;
              int q = spec_cx_cls(k);
              double lg = 0.0;
#pragma unroll
              for (int z = 0; z < SPEC_NCLS; z++)
                if (z == q) lg = s.lgcls[z];
              st.sec_act_coef[k * ld + cell] = q < 0 ? 1.0 : exp(lg);
            }
          }
          if (WID == 0) {
#pragma unroll
            for (int k = 0; k < SPEC_NSRFRXN; k++) st.free_site[k * ld + cell] = ev.fsite[k];
            if (SPEC_NEQSR > 0 && st.eqsrfcplx_conc) {
#pragma unroll
              for (int k = 0; k < SPEC_NSRFCPLX; k++) st.eqsrfcplx_conc[k * ld + cell] = ev.S[k];
            }
#pragma unroll
            for (int k = 0; k < SPEC_NKIN; k++) st.mnrl_rate[k * ld + cell] = ev.mrate[k];
          }
        }
      }
      if (__all_sync(FULL, done)) break;
    }

    if (WID == 0 && inrange) {
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
    specw_barrier();
  }

  if (WID == 0) {
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
