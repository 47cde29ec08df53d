// pfrx_spec2.cuh -- network-SPECIALISED thread-per-cell kernel, second form.
//
// Same contract as pfrx_spec.cuh (one thread = one cell, RStep / RReact of
// reaction.F90:3564-4055 around code that specialize2.py generates for one reaction
// network; kernel `pfrx_spec_kernel`, globals `pfrx_spec_sig` / `pfrx_spec_info`), for
// networks in the LOG formulation whose complexes have integer stoichiometry.  What
// changes is the arithmetic formulation of one Newton iteration -- the instruction stream
// of form 1 is 17 700 straight-line instructions per iteration (283 KB) and the kernel is
// bound by instruction supply (profiles/r02_ubench.txt: one warp per scheduler streams
// straight-line code from L2 at <= 0.2 instructions per clock), so the lever is the number
// of instructions per iteration:
//
//  * secondary molalities as PRODUCTS: a_i = c_i gamma_i is split into mantissa f_i in
//    [1, 2) and exponent e_i once per iteration; m_k = K_k^-1 prod a_i^nu / gamma_k is a
//    product of a few mantissa powers and an integer sum of exponents (reaction.F90:4665-4759
//    computes exp(sum nu ln a_i): 88 exp + 15 log per iteration for Hanford; here 11 exp for
//    the activity-coefficient classes).  The exponent sum saturates to 0 / +inf like exp.
//  * the Jacobian is assembled in ln-space: Jt_ij = J_ij c_j is SYMMETRIC for aqueous
//    complexation (sum_k nu_ki nu_kj m_k), equilibrium surface complexation and TST mineral
//    kinetics, so only the upper triangle is accumulated (345 instead of 488 + 202 products
//    for Hanford) and RSolve's column scaling J(:,j) *= c_j (reaction.F90:5493-5497) is
//    already done; its row scaling 1/max(1, max_j |J_ij|) takes |Jt_ij| / c_j.
//  * the linear solve.  For this class of networks Jt is symmetric POSITIVE DEFINITE
//    (K1 (diag(c) + N^T diag(m) N) + a covariance-type sorption term + k A QK nu nu^T per
//    mineral), so the generator emits a sparse L D L^T factorisation without pivoting in a
//    minimum-fill elimination order (SPEC_SYM 1: Hanford 69 stored entries and 185 multiply-adds
//    instead of 182 and 650, no row scaling, no pivot search, every address a literal).  RSolve's
//    row scaling and LUDecomposition's implicit-scaled partial pivoting (reaction.F90:5457-5516,
//    utility.F90:597-688) are what makes Gaussian elimination safe for a GENERAL matrix; for a
//    positive definite one symmetric elimination is backward stable in any order and invariant
//    under diagonal scaling, and the update it returns differs from the reference's by rounding.
//    SPEC_SYM 0 keeps the reference's algorithm on the full matrix (Crout, the reference's column
//    order, pivot search as a tournament with the reference's tie rule) for comparison.
//  * the iterate c lives in the thread's shared-memory slice next to the matrix, the fixed
//    accumulation is re-read from rt_auxvar%total in HBM/L2 (it does not change during a
//    sub-step): 84 doubles per cell instead of 212, so TWO warps per scheduler are resident and
//    execute the same instruction stream (256-thread lock-step blocks).
//
// Results differ from form 1 / the oracle by rounding only (products instead of exp of
// sums); the parity tests hold the same 1e-10 and identical Newton / sub-step / cut counts.
#pragma once

#include "../../include/pfrx.h"
#include "pfrx_types.cuh"

#ifdef S2_HOST
// ---- host build of the generated routines (tests/test_spec2_host.py): same source, libm ----
#include <cmath>
#include <cstring>
#define S2_FN static inline
static inline int s2_hiint(double x) { long long b; std::memcpy(&b, &x, 8); return (int)(b >> 32); }
static inline int s2_loint(double x) { long long b; std::memcpy(&b, &x, 8); return (int)(b & 0xffffffffll); }
static inline double s2_hilo(int hi, int lo) {
  long long b = ((long long)hi << 32) | (unsigned int)lo;
  double x; std::memcpy(&x, &b, 8); return x;
}
static inline double sx_exp(double x) { return std::exp(x); }
static inline double sx_log(double x) { return std::log(x); }
static inline double sx_rcp(double x) { return 1.0 / x; }
static inline double sx_div(double a, double b) { return a / b; }
#define S2_INF HUGE_VAL
#else
#include <cuda_runtime.h>
#include "pfrx_fastmath.cuh"
#define S2_FN __device__ __forceinline__
S2_FN int s2_hiint(double x) { return __double2hiint(x); }
S2_FN int s2_loint(double x) { return __double2loint(x); }
S2_FN double s2_hilo(int hi, int lo) { return __hiloint2double(hi, lo); }
#ifndef SPEC_FASTMATH
#define SPEC_FASTMATH 1
#endif
S2_FN double sx_exp(double x) { return SPEC_FASTMATH ? pfrx_exp(x) : exp(x); }
S2_FN double sx_log(double x) { return SPEC_FASTMATH ? pfrx_log(x) : log(x); }
S2_FN double sx_rcp(double x) { return SPEC_FASTMATH ? pfrx_rcp(x) : 1.0 / x; }
S2_FN double sx_div(double a, double b) { return SPEC_FASTMATH ? pfrx_div(a, b) : a / b; }
#define S2_INF __longlong_as_double(0x7ff0000000000000ll)
#endif

#define SPEC_LN 2.30258509299  // pflotran_constants.F90:84 (truncated there)

#ifndef SPEC_REFILL
#define SPEC_REFILL 0  // finished lanes fetch the next cell (ragged workloads)
#endif
#ifndef SPEC_SYNC
#define SPEC_SYNC 1    // the warps of a block vote once per Newton iteration and share one instruction stream
#endif

// ---- shared-memory slots of a thread (doubles); element e of a thread is at W[e * 32] ----
#ifndef SPEC_SYM
#define SPEC_SYM 0
#endif
#define S2_JS (SPEC_NC + 1)                       // full matrix: + one column (scaling factor, then right-hand side)
#define S2_NJ (SPEC_SYM ? SPEC_NL : SPEC_NC * S2_JS)  // matrix slots: the entries of L, or the full matrix
#define S2_OFF_C S2_NJ                            // the iterate c (N)
#define S2_OFF_FIX (S2_OFF_C + SPEC_N)            // fixed accumulation of the sub-step (N)
#define S2_OFF_MN (S2_OFF_FIX + SPEC_N)           // kinetic minerals: volume fraction, specific area (2 NKIN)
#define S2_OFF_FRZ (S2_OFF_MN + 2 * SPEC_NKIN)    // frozen coefficients: gamma_i (NAQ), then 1 / gamma_k (NCX)
#define S2_SLOTS (S2_OFF_FRZ + (SPEC_ACT_UPD ? 0 : SPEC_NAQ + SPEC_NCX))
#define S2_NTV (SPEC_NSTASH > 0 ? SPEC_NSTASH : 1)
#ifndef SPEC_NEV
#define SPEC_NEV 0
#endif
#define S2_NEV (SPEC_NEV > 0 ? SPEC_NEV : 1)  // entries of the columns the reference's sorption Jacobian leaves incomplete
#define SW(e) W[(e) * 32]
#define JX(ci, cj) (((ci) * S2_JS + (cj)) * 32)

struct Spec2Cell {
  double denL;  // den_kg * 1e-3
  double psv;   // porosity * saturation * 1000 * volume (0 in a dry cell)
  double vol, rock, temp;  // rock = soil particle density * (1 - porosity), the ROCK_SURFACE site basis
  double aw;    // activity of water
  double rdt;   // 1 / dt of the current sub-step
  double Isec, msec;  // sum z^2 m, sum m over the secondary species of the latest RTotal
  double Iact;        // ionic strength of the latest activity-coefficient evaluation
  bool dry;
  bool store;   // false: the lane has finished its cell, rt_auxvar must not be touched
  bool rates;   // false in the pass that only finds its > max (RReaction is not reached, reaction.F90:3880)
};

// p * 2^e for p in [2^-60, 2^60].  Underflow: the exponent is clamped (the value stays below 1e-270
// where exp() of the summed logarithms gives 0 or a denormal); overflow: the largest exponent of an
// evaluation is tracked in emax and spec2_eval poisons the residual with NaN when it exceeds 960,
// which is how the reference's +Inf ends too (NaN iterates until its > max, then a cut)
S2_FN double s2_scale(double p, int e, int &emax) {
  emax = e > emax ? e : emax;
  const int ec = e < -960 ? -960 : e;
  return s2_hilo(s2_hiint(p) + (ec << 20), s2_loint(p));
}
// mantissa in [1, 2) and exponent of a positive normal number
S2_FN double s2_split(double a, int &e) {
  const int hi = s2_hiint(a);
  e = (hi >> 20) - 1023;
  return s2_hilo((hi & 0x000fffff) | 0x3ff00000, s2_loint(a));
}

// ---- generated for the network -------------------------------------------------------
// one evaluation at the iterate in the slice: activity coefficients, RTotal, sorption, residual,
// kinetic minerals; Jt into the slice (SPEC_SYM: the lower triangle in elimination order;
// otherwise both triangles and the structural zeros); the totals of this evaluation into tv;
// free-site / surface-complex concentrations and mineral rates stored through
S2_FN void spec2_eval(double (&res)[SPEC_N], double (&tv)[S2_NTV], double (&ev)[S2_NEV], Spec2Cell &s, double *W,
                      const DevState &st, long long cell);
// rt_auxvar%total (and %total_sorb_eq) <- tv
S2_FN void spec2_store_totals(const double (&tv)[S2_NTV], const double *W, const Spec2Cell &s, const DevState &st,
                              long long cell, bool tot, bool sorb);
// activity coefficients of the cell, recomputed from the ionic strength of the latest evaluation
S2_FN void spec2_store_act(const Spec2Cell &s, const DevState &st, long long cell);
// rt_auxvar%sec_molal of the cell's last evaluation (iterate and activity of water in the slice / s, ionic strength s.Iact)
S2_FN void spec2_store_sec(const double *W, const Spec2Cell &s, const DevState &st, long long cell);
// frozen coefficients into the slice at cell entry
S2_FN void spec2_load_frozen(double *W, const DevState &st, long long cell);
// sub-step entry: fixed accumulation and the mineral inputs into the slice
S2_FN void spec2_begin(double *W, const Spec2Cell &s, const DevState &st, long long cell);
// RStep entry: sum z^2 m and sum m over the state's secondary species
S2_FN void spec2_isec(Spec2Cell &s, const DevState &st, long long cell);
#if SPEC_SYM
// sparse L D L^T of Jt in the slice and the solve for the coupled species: res <- update; false
// when a pivot is not positive (Jt numerically not positive definite)
S2_FN bool spec2_solve_sym(double *W, double (&res)[SPEC_N], const double (&ev)[S2_NEV]);
#endif

// ---- RSolve + LU (reaction.F90:5457-5516, utility.F90:597-735) on Jt ---------------------
S2_FN bool spec2_solve(double *W, double (&res)[SPEC_N], const double (&ev)[S2_NEV], const Spec2Cell &s) {
  constexpr int N = SPEC_N, NC = SPEC_NC, NCA = NC > 0 ? NC : 1;
  bool bad = false;
  // species outside the matrix: J is diagonal (RTAccumulationDerivative only)
#pragma unroll
  for (int i = 0; i < N; i++) {
    if (spec_cmap(i) < 0) {
      double Jd = (i < SPEC_NAQ) ? (1.0 * s.denL) * ((s.dry ? 0.0 : s.psv) * s.rdt) : s.vol * s.rdt;
      if (s.dry) Jd = 1.0;
      const double nm = sx_rcp(fmax(1.0, fabs(Jd)));
      const double a = (Jd * nm) * SW(S2_OFF_C + i);
      if (!(fabs(a) > 0.0)) bad = true;
      res[i] = sx_div(res[i] * nm, a);
    }
  }
  if (NC == 0) return !bad;
#if SPEC_SYM
  (void)NCA;
  return spec2_solve_sym(W, res, ev) && !bad;
#else
  double b[NCA];
  {
    double ic[NCA];
#pragma unroll
    for (int j = 0; j < NC; j++) ic[j] = sx_rcp(SW(S2_OFF_C + spec_sp_of(j)));
    // row scaling: entries outside the network's structure (spec_jnz) hold exact zeros
#pragma unroll
    for (int i = 0; i < NC; i++) {
      double row[NCA];
      double m = 0.0;
#pragma unroll
      for (int j = 0; j < NC; j++) {
        if (spec_jnz(i, j)) {
          row[j] = W[JX(i, j)];
          const double av = fabs(row[j]) * ic[j];  // |J_ij| = |Jt_ij| / c_j
          m = av > m ? av : m;
        }
      }
      const double nm = sx_rcp(fmax(1.0, m));
      b[i] = res[spec_sp_of(i)] * nm;
      double m2 = 0.0;
#pragma unroll
      for (int j = 0; j < NC; j++) {
        if (spec_jnz(i, j)) {
          const double v = row[j] * nm;
          W[JX(i, j)] = v;
          const double av = fabs(v);
          m2 = av > m2 ? av : m2;
        }
      }
      if (!(m2 > 0.0)) bad = true;
      W[JX(i, NC)] = sx_rcp(m2);  // implicit-scaling factor of the row (utility.F90:611-622)
    }
  }
  (void)ev;
  if (bad) return false;
  int ro[NCA];
#pragma unroll
  for (int i = 0; i < NC; i++) ro[i] = JX(i, 0);
#pragma unroll
  for (int j = 0; j < NC; j++) {
    double x[NCA];  // column j: the finished upper part, then the candidates of the lower part
#pragma unroll
    for (int k = 0; k < NC; k++) {
      if (k < j) {
        const double *r = W + ro[k];
        double sum = r[j * 32];
#pragma unroll
        for (int m = 0; m < NC; m++)
          if (m < k) sum -= r[m * 32] * x[m];
        x[k] = sum;
        if (k > 0) W[ro[k] + j * 32] = sum;
      }
    }
    double td[NCA];
    int ti[NCA];
#pragma unroll
    for (int i = 0; i < NC; i++) {
      if (i >= j) {
        const double *r = W + ro[i];
        double sum = r[j * 32];
#pragma unroll
        for (int m = 0; m < NC; m++)
          if (m < j) sum -= r[m * 32] * x[m];
        x[i] = sum;
        td[i - j] = r[NC * 32] * fabs(sum);
        ti[i - j] = i;
      }
    }
    // arg-max, the LAST maximum wins (dum .ge. aamax): tournament, the right operand holds the later rows
#pragma unroll
    for (int sd = 1; sd < NC; sd <<= 1) {
#pragma unroll
      for (int a = 0; a < NC; a += 2 * sd) {
        if (a + sd < NC - j) {
          const bool ge = td[a + sd] >= td[a];
          td[a] = ge ? td[a + sd] : td[a];
          ti[a] = ge ? ti[a + sd] : ti[a];
        }
      }
    }
    const int imax = ti[0];
    const int rj = ro[j];
    int rmax = rj;
    double pv = x[j];
#pragma unroll
    for (int i = 0; i < NC; i++) {
      if (i > j) {
        const bool p = (i == imax);
        rmax = p ? ro[i] : rmax;
        pv = p ? x[i] : pv;
      }
    }
    if (pv == 0.0) pv = 1.0e-20;
    if (j != NC - 1) {
      const double rp = sx_rcp(pv);
      // through the offsets BEFORE the interchange: old row j lands at position imax with x[j] * rp,
      // the pivot row's entry is overwritten below
#pragma unroll
      for (int i = 0; i < NC; i++)
        if (i >= j) W[ro[i] + j * 32] = x[i] * rp;
    }
    W[rmax + j * 32] = pv;
#pragma unroll
    for (int i = 0; i < NC; i++)
      if (i > j) ro[i] = (i == imax) ? rj : ro[i];
    ro[j] = rmax;
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
#endif  // !SPEC_SYM
}

// log-formulation update (reaction.F90:3985-4010): c <- c exp(-clamped update); returns the maximum
// relative change |c_new - c| / c = |exp(-u) - 1|.  `commit` false leaves the slice untouched.
S2_FN double spec2_update(double *W, const double (&res)[SPEC_N], double max_dlnC, double (&cn)[SPEC_N]) {
  double maxrel = 0.0;
#pragma unroll
  for (int i = 0; i < SPEC_N; i++) {
    const double u = copysign(1.0, res[i]) * fmin(fabs(res[i]), max_dlnC);
    const double e = sx_exp(-u);
    cn[i] = SW(S2_OFF_C + i) * e;
    const double v = fabs(e - 1.0);
    maxrel = v > maxrel ? v : maxrel;
    if (v != v) maxrel = v;  // a NaN update never converges
  }
  return maxrel;
}

#ifndef S2_HOST
extern "C" {
__device__ const unsigned long long pfrx_spec_sig = SPEC_SIG;
// {N, shared doubles per block, threads per block, min blocks per SM, cells per block}
__device__ const int pfrx_spec_info[5] = {SPEC_N, S2_SLOTS * SPEC_THREADS, SPEC_THREADS, SPEC_MINBLOCKS, SPEC_THREADS};
// bit 0 would announce a kernel that hands out cells through DevState.order (pfrx_spec.cuh's refill skeleton);
// form 2 serves uniform workloads and takes cells in index order
__device__ const int pfrx_spec_flags = 0;
}

__device__ __forceinline__ int s2_vote_and(int p) { return SPEC_SYNC ? __syncthreads_and(p) : __all_sync(0xffffffffu, p); }

// ======================================================================================
// RStep / RReact as a state machine per lane: one pass of the loop = one Newton iteration of
// every unfinished cell of the group (block with SPEC_SYNC, warp without).  A finished lane
// rides along with its stores disabled, or (SPEC_REFILL) takes the next unprocessed cell.
// ======================================================================================
extern "C" __global__ void __launch_bounds__(SPEC_THREADS, SPEC_MINBLOCKS)
    pfrx_spec_kernel(DevState st, long long ncell, double tran_dt, SpecParams prm, DevSummary *summ) {
  constexpr int N = SPEC_N, NAQ = SPEC_NAQ, NIM = N - NAQ;
  extern __shared__ double smem[];
  const int lane32 = threadIdx.x & 31;
  double *W = smem + (size_t)(threadIdx.x >> 5) * (S2_SLOTS * 32) + lane32;
  const long long ld = st.ld;
  const double target = tran_dt;

  unsigned long long l_active = 0, l_its = 0, l_cut = 0;
  long long l_first = -1;
  int l_maxits = 0, l_maxkin = 0, l_maxerr = 0, l_maxsub = 0;

  Spec2Cell s;
  s.store = false;
  s.rates = false;
  s.dry = false;
  s.denL = s.psv = s.vol = s.rock = s.temp = s.Isec = s.msec = s.Iact = 0.0;
  s.aw = 1.0;
  s.rdt = 1.0;
  unsigned small_mask = 0u;
  double small_val[N];  // totals <= 1e-40 to put back at the end: indexed dynamically, lives in local memory
  double gimm[NIM > 0 ? NIM : 1];
  double cumulative = 0.0, dt = target, norm0 = 0.0;
  int ncuts = 0, nconst = 0, nss = 0, nit = 0, nku = 0, its = 0;
  bool done = true, aborted = false, had_cut = false, need_begin = false;
  long long cell = ncell - 1;
  bool inrange = false, live = false;
#if SPEC_REFILL
  bool have = false, loaded_once = false;
#else
  long long base = SPEC_SYNC ? (long long)blockIdx.x * blockDim.x
                             : ((long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31));
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
          cell = mine;
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
      const int off = SPEC_SYNC ? threadIdx.x : lane32;
      inrange = base + off < ncell;
      cell = inrange ? base + off : ncell - 1;
      live = inrange && !(st.imat && st.imat[cell] <= 0);
      base += (long long)gridDim.x * blockDim.x;
      need_new = false;
      fresh = true;
    }
#endif
    if (fresh) {
      // ---- RStep entry (reaction.F90:3600-3650)
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
      spec2_isec(s, st, cell);
      if (!SPEC_ACT_UPD) spec2_load_frozen(W, st, cell);
      small_mask = 0u;
#pragma unroll
      for (int i = 0; i < N; i++) {
        const double t = (i < NAQ) ? st.total[i * ld + cell] : st.immobile[(i - NAQ) * ld + cell];
        if (t <= 1.e-40) small_mask |= 1u << i;
        if (i >= NAQ) gimm[i - NAQ] = t;  // the guess keeps the unclamped value
      }
      if (small_mask) {  // rare: clamp to 1e-40 for the solve (reaction.F90:3615-3630)
#pragma unroll 1
        for (int i = 0; i < N; i++) {
          if ((small_mask >> i) & 1u) {
            double *p = (i < NAQ) ? st.total + i * ld + cell : st.immobile + (i - NAQ) * ld + cell;
            small_val[i] = *p;
            if (live) *p = 1.e-40;
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
    if (s2_vote_and(have ? 0 : 1)) break;  // every cell has been handed out and finished
#endif

    // ---- RReact entry (reaction.F90:3829-3850) for lanes that start a sub-step: the guess is
    // rt_auxvar%pri_molal (where the last accepted sub-step left it) / the immobile guess
    if (need_begin && (!done || fresh)) {
#pragma unroll
      for (int i = 0; i < N; i++) SW(S2_OFF_C + i) = (i < NAQ) ? st.pri_molal[i * ld + cell] : gimm[i >= NAQ ? i - NAQ : 0];
      spec2_begin(W, s, st, cell);
      s.rdt = 1.0 / dt;
      its = 0;
      need_begin = false;
    }
    if (!done) its++;
    s.store = !done;
    const bool over = its > prm.max_its;
    s.rates = !done && !over;

    // ---- one Newton iteration (reaction.F90:3860-4041); a warp whose 32 cells are all finished
    // skips the arithmetic (it still votes)
    double res[N], tv[S2_NTV], ev[S2_NEV];
    bool conv = false, need_solve = false, fail = false, solve_error = false;
    const bool warp_live = __any_sync(0xffffffffu, !done);
    if (warp_live) {
      spec2_eval(res, tv, ev, s, W, st, cell);
      double mabs = 0.0, ss = 0.0;
#pragma unroll
      for (int i = 0; i < N; i++) {
        mabs = fmax(mabs, fabs(res[i]));
        ss += res[i] * res[i];
      }
      const double nrm = sqrt(ss);
      if (its == 1) norm0 = nrm;
      const double rel = nrm / norm0;
      conv = (mabs < prm.tol_res) || (rel < prm.tol_relres);
      need_solve = !done && !over && !conv;
      fail = !done && over;
    }
    // a warp none of whose cells needs the solve goes straight to the vote that ends the pass
    if (__any_sync(0xffffffffu, need_solve)) {
      const bool ok = spec2_solve(W, res, ev, s);
      if (need_solve) {
        if (!ok) {
          fail = true;
          solve_error = true;
        } else {
          double cn[N];
          const double maxrel = spec2_update(W, res, prm.max_dlnC, cn);
          if (maxrel < prm.tol_relchange) {
            conv = true;
          } else {
#pragma unroll
            for (int i = 0; i < N; i++) SW(S2_OFF_C + i) = cn[i];
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
        spec2_store_totals(tv, W, s, st, cell, solve_error, true);
        if (solve_error) {
#pragma unroll
          for (int i = NAQ; i < N; i++) st.immobile[(i - NAQ) * ld + cell] = SW(S2_OFF_C + i);
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
        spec2_store_totals(tv, W, s, st, cell, true, true);
#pragma unroll
        for (int i = 0; i < N; i++) {
          const double ci = SW(S2_OFF_C + i);
          if (i < NAQ) {
            st.pri_molal[i * ld + cell] = ci;
          } else {
            st.immobile[(i - NAQ) * ld + cell] = ci;
            gimm[i >= NAQ ? i - NAQ : 0] = ci;
          }
        }
        bool upd = false;
        if (SPEC_NKIN > 0) {  // RUpdateKineticState: the rates of the converged iterate are in rt_auxvar%mnrl_rate
          upd = true;
#pragma unroll
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
          dt = fmin(2.0 * dt, target - cumulative);
        }
        if (cumulative >= target)
          done = true;
        else
          need_begin = true;
      }
    }
#if SPEC_REFILL
    const bool publish = have && done;
#else
    const bool publish = s2_vote_and(done ? 1 : 0) != 0;
    if (publish) need_new = true;
#endif
    if (publish) {
      // ---- publish the cell (reaction.F90:3700-3738)
      if (live) {
        if (aborted) {
#pragma unroll
          for (int i = 0; i < NAQ; i++) st.pri_molal[i * ld + cell] = SW(S2_OFF_C + i);
        } else if (small_mask) {
#pragma unroll 1
          for (int i = 0; i < N; i++) {
            if ((small_mask >> i) & 1u) {
              double *p = (i < NAQ) ? st.total + i * ld + cell : st.immobile + (i - NAQ) * ld + cell;
              *p = small_val[i];
            }
          }
        }
        spec2_store_sec(W, s, st, cell);
        if (SPEC_ACT_UPD) spec2_store_act(s, st, cell);
        if (st.ln_act_h2o && SPEC_USE_ACT_H2O) st.ln_act_h2o[cell] = (s.aw == 1.0) ? 0.0 : log(s.aw);
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
#endif  // !S2_HOST
