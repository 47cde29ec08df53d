// pfrx_sandbox.cuh -- the ELM-CN reaction sandboxes on the device:
//   SomDecReact / SomDecReact1 / SomDecReact2 / SomDecNemission
//                         reaction_sandbox_somdec.F90:1504-3640
//   NitrifReact           reaction_sandbox_nitrif.F90:234-502
//   DenitrReact           reaction_sandbox_denitr.F90:212-404
//   response functions    elm_rspfuncs.F90:61-375,  HfunctionSmooth utility.F90:2542-2597
//
// The reference dispatches these through a polymorphic linked list
// (reaction_sandbox_base_type, RSandboxEvaluate reaction_sandbox.F90:294-330)
// and walks per-reaction linked lists of Monod / inhibition terms; here the
// sandbox list is an enum array and the terms are CSR tables (pfrx_somdec).
// One thread evaluates one cell.  `Cell` provides
//   C(i)      free concentration of unknown i (immobile species at naq + k)
//   TOT(i)    rt_auxvar%total(i) of aqueous species i, mol/L
//   LNA(i)    ln activity of aqueous species i
//   DT(i,j)   rt_auxvar%aqueous%dtotal(i,j)
//   RES(i), J(i,j)   residual (mol/s) and Jacobian being assembled
//   NC(k)     the N:C ratios that persist between evaluations (pfrx_state.somdec_nc)
//   sat, por, vol, temp, elm_* per-cell scalars
// The operation order inside every expression is the reference's.
#pragma once

namespace pfrx_sbx {

// GetMoistureResponse of the ELM_PFLOTRAN build (elm_rspfuncs.F90:124-237): f_w of a SOMDECOMP reaction when a flow
// mode is active.  CLMCN: Clapp-Hornberger matric potential on a log scale between -10 MPa and the air-entry
// suction; DLEM: Tian et al. 2010 between field capacity and the effective porosity.
__device__ __forceinline__ double elm_moisture_response(double theta, int itype, double sucsat, double bd_dry, double bsw,
                                                        double watfc, double effpor) {
  const double minpsi = -10.0e6, g = 9.8068;
  if (itype == PFRX_MOISTURE_RESPONSE_CLMCN) {
    const double maxpsi = sucsat * (-g);
    const double lsat = theta / fmin(1.0, 1.0 - fmin(0.9999, bd_dry / 2.70e3));
    double psi = sucsat * (-g) * pow(lsat, -bsw);
    psi = fmin(psi, maxpsi);
    if (!(psi > minpsi)) return 0.0;
    double F = log(minpsi / psi) / log(minpsi / maxpsi);
    if (psi > (maxpsi - 1.0e02)) F = F * 0.10;
    return F;
  }
  if (itype == PFRX_MOISTURE_RESPONSE_DLEM) {
    if (theta >= effpor) return 1.0;
    if (theta <= watfc) return 0.0;
    const double se = (theta - watfc) / (effpor - watfc);
    double F = (double)1.0f - se * se * (double)0.368f * exp(se);
    if (F < 0.0) F = 0.0;
    if (F > 1.0) F = 1.0;
    return F;
  }
  return 1.0;
}


__device__ __forceinline__ void hsmooth(double x, double x_1, double x_0, double &H, double &dH) {
  if (fabs(x_1 - x_0) < 1.e-50) {
    H = copysign(0.5, (x - x_1)) + 0.5;
    dH = 0.0;
    return;
  }
  {
    // Far from the transition the quotient below is decided by comparisons: beyond one
    // transition width on either side it is certainly > 1 or < 0.  Same results, and the
    // five divisions are skipped for every concentration that is not within a decade of
    // the cut-off (all but exhausted pools).
    const double d = x_1 - x_0;
    if (d > 0.0) {
      if (x > x_1 + d) {
        H = 1.0;
        dH = 0.0;
        return;
      }
      if (x < x_0 - d) {
        H = 0.0;
        dH = 0.0;
        return;
      }
    }
  }
  const double r = (x - x_0) / (x_1 - x_0);
  if (r < 0.0) {
    H = 0.0;
    dH = 0.0;
  } else if (r > 1.0) {
    H = 1.0;
    dH = 0.0;
  } else {
    double x_star = 1.0 - (x - x_0) * (x - x_0) / (x_1 - x_0) / (x_1 - x_0);
    H = 1.0 - x_star * x_star;
    dH = 4.0 * x_star * (x - x_0) / (x_1 - x_0) / (x_1 - x_0);
  }
}

__device__ __forceinline__ double monod(double conc, double k) { return conc / (conc + k); }
__device__ __forceinline__ double dmonod(double conc, double k) { return k / (conc + k) / (conc + k); }

__device__ __forceinline__ double wfps(double s) {
  return pow((1.27 - s) / 0.67, 3.1777) * pow((s - 0.0012) / 0.5988, 2.84);
}

__device__ __forceinline__ double temperature_response(double tc, int itype, double q) {
  const double one_over_71_02 = 1.408054069e-2;
  switch (itype) {
    case PFRX_TEMPERATURE_RESPONSE_Q10:
      if (tc > 0.0) return pow(q, (tc - 25.0) / 10.0);
      return pow(q, -25.0 / 10.0) * pow(2.0, tc / 10.0);
    case PFRX_TEMPERATURE_RESPONSE_CLMCN: {
      double tk = tc + 273.15;
      if (tk > 227.15) return exp(308.56 * (one_over_71_02 - 1.0 / (tk - 227.13)));
      return 0.0;
    }
    case PFRX_TEMPERATURE_RESPONSE_DLEM:
      if (tc < -5.0) return 0.0;
      if (tc >= 30.0) return 1.0;
      return pow(q, (tc - 30.0) / 10.0);
    case PFRX_TEMPERATURE_RESPONSE_ARRHENIUS:
      return exp(q / PFRX_IDEAL_GAS_CONSTANT * (1.0 / 298.15 - 1.0 / (tc + 273.15)));
    default:
      return 1.0;
  }
}

// pH factor of the N2O terms (Parton et al. 1996): 0.56 + atan(pi 0.45 (pH - 5))/pi
template <class Cell>
__device__ __forceinline__ double f_ph(const Cell &s, int proton_id) {
  const double rpi = 3.14159265358979323846;
  double ph = 6.5;
  if (proton_id >= 0) ph = -s.LNAc(proton_id) * 0.43429448190325182765;  // -log10(m gamma)
  return 0.56 + atan(rpi * 0.45 * (-5.0 + ph)) / rpi;
}

// ---- SOMDECOMP ---------------------------------------------------------------------
template <class Cell>
struct SomDec {
  Cell &s;
  const pfrx_somdec &sd;
  const int off;  // offset_immobile
  double tran_dt;
  __device__ SomDec(Cell &c, double dt) : s(c), sd(c.cfg.sd), off(c.cfg.naq), tran_dt(dt) {}

  __device__ __forceinline__ double conc(int id, int itype) const {
    return itype == PFRX_SPEC_AQUEOUS ? s.TOTc(id) : s.Cc(off + id);
  }
  __device__ __forceinline__ int row(int id, int itype) const { return itype == PFRX_SPEC_AQUEOUS ? id : off + id; }

  // MONOD / INHIBITION lists and the Ox Monod term (:2062-2150, :2700-2800)
  __device__ __forceinline__ void modifiers(int rxn, int ispec_uc, bool react2, double theta, double c_nh4,
                                            double c_no3, double &crate_uc, double &dcrate_uc_duc, double &fnh4,
                                            double &dfnh4, double &fno3, double &dfno3) {
    double fmb = 1.0, dfmb = 0.0;
#pragma unroll 1
    for (int k = sd.monod_ptr[rxn]; k < sd.monod_ptr[rxn + 1]; k++) {
      const double mk = sd.monod_half_saturation[k], thr = sd.monod_threshold[k];
      const int sid = sd.monod_specid[k];
      double t;
      if (react2 && sd.nh4_id >= 0 && sd.nh4_id == sid) {
        t = fmax(0.0, c_nh4 - thr);
        if (sd.monod_pool_normalized[k]) t = t / s.Cc(off + ispec_uc);
        fnh4 = monod(t, mk);
        dfnh4 = dmonod(t, mk);
      } else if (react2 && sd.no3_id >= 0 && sd.no3_id == sid) {
        t = fmax(0.0, c_no3 - thr);
        if (sd.monod_pool_normalized[k]) t = t / s.Cc(off + ispec_uc);
        fno3 = monod(t, mk);
        dfno3 = dmonod(t, mk);
      } else {
        t = conc(sid, sd.monod_specitype[k]);
        t = fmax(0.0, t - thr);
        if (sd.monod_pool_normalized[k]) {
          t = t / s.Cc(off + ispec_uc);
          if (sd.monod_specitype[k] == PFRX_SPEC_AQUEOUS) {
            if (react2)
              t = t * theta * 1000.0;
            else
              t = t * s.por * s.sat * 1000.0;
          }
        }
        double fx = monod(t, mk), dfx = dmonod(t, mk);
        if (ispec_uc != sid) dfx = 0.0;
        dfmb = dfmb * fx + fmb * dfx;
        fmb = fmb * fx;
      }
    }
#pragma unroll 1
    for (int k = sd.inhib_ptr[rxn]; k < sd.inhib_ptr[rxn + 1]; k++) {
      const double PI = 3.14159265358979323846;
      const double ik = sd.inhib_constant[k], ik2 = sd.inhib_constant2[k];
      const double t = conc(sd.inhib_specid[k], sd.inhib_specitype[k]);
      double fx = 1.0, dfx = 0.0;
      if (ik2 == -999.0 || sd.inhib_itype[k] != PFRX_INHIBITION_THRESHOLD) {
        if (sd.inhib_itype[k] == PFRX_INHIBITION_MONOD) {
          fx = ik / (t + ik);
          dfx = -ik / (t + ik) / (t + ik);
        } else if (sd.inhib_itype[k] == PFRX_INHIBITION_INVERSE_MONOD) {
          fx = monod(t, ik);
          dfx = dmonod(t, ik);
        }
      } else {
        fx = 0.5 + atan((t - ik) * ik2) / PI;
        double u = (t - ik) * ik2;
        dfx = (ik2 / (1.0 + u * u)) / PI;
      }
      if (ispec_uc != sd.inhib_specid[k]) dfx = 0.0;
      dfmb = dfmb * fx + fmb * dfx;
      fmb = fmb * fx;
    }
    dcrate_uc_duc = dcrate_uc_duc * fmb + crate_uc * dfmb;
    crate_uc = crate_uc * fmb;
    double f_ox = 1.0, df_ox = 0.0;
    if (sd.ox_response_function[rxn] == PFRX_OX_RESPONSE_MONOD && sd.ox_specid[rxn] >= 0) {
      double Ox = conc(sd.ox_specid[rxn], sd.ox_specitype[rxn]);
      f_ox = monod(Ox, sd.ox_half_saturation[rxn]);
      df_ox = dmonod(Ox, sd.ox_half_saturation[rxn]);
    }
    dcrate_uc_duc = dcrate_uc_duc * f_ox + crate_uc * df_ox;
    crate_uc = crate_uc * f_ox;
  }

  // residual entries shared by React1/React2 (:2160-2215, :2880-2935)
  __device__ __forceinline__ int common_residual(int irxn, double crate, double cst, double unc) {
    const bool up_aq = sd.upstream_is_aqueous[irxn] != 0;
    const int ires_uc = up_aq ? sd.upstream_c_id[irxn] : off + sd.upstream_c_id[irxn];
    const int ires_co2 = row(sd.co2_id, sd.co2_itype);
    s.RES(ires_uc) = s.RES(ires_uc) + crate;
    s.RES(ires_co2) = s.RES(ires_co2) - cst * crate;
    if (sd.upstream_hr_id[irxn] >= 0) s.RES(off + sd.upstream_hr_id[irxn]) -= cst * crate;
    if (sd.hr_id >= 0) s.RES(off + sd.hr_id) -= cst * crate;
    int ires_ox = -1;
    if (sd.o2_id >= 0) {
      ires_ox = row(sd.o2_id, sd.o2_itype);
      s.RES(ires_ox) = s.RES(ires_ox) + cst * crate;
    }
#pragma unroll 1
    for (int j = sd.downstream_ptr[irxn]; j < sd.downstream_ptr[irxn + 1]; j++) {
      const int dc = sd.downstream_c_id[j];
      if (dc >= 0) {
        const int r = sd.downstream_is_aqueous[j] ? dc : off + dc;
        s.RES(r) = s.RES(r) - sd.downstream_stoich[j] * crate;
      }
    }
    if (sd.upstream_n_id[irxn] >= 0) {
      const int r = up_aq ? sd.upstream_n_id[irxn] : off + sd.upstream_n_id[irxn];
      s.RES(r) = s.RES(r) + unc * crate;
    }
    return ires_ox;
  }

  __device__ __forceinline__ void downstream_n_residual(int irxn, double crate) {
#pragma unroll 1
    for (int j = sd.downstream_ptr[irxn]; j < sd.downstream_ptr[irxn + 1]; j++) {
      const int dn = sd.downstream_n_id[j];
      if (dn >= 0) {
        const int r = sd.downstream_is_aqueous[j] ? dn : off + dn;
        s.RES(r) = s.RES(r) - sd.downstream_stoich[j] * crate * s.NC(sd.nrxn + j);
      }
    }
  }

  // Jacobian column jcol: CO2 (+O2 and trackers), upstream C, downstream C rows
  __device__ __forceinline__ void common_jacobian(int irxn, int jcol, int jaq, int ires_ox, double dco2_dx,
                                                  double duc_dx, bool wrt_uc) {
    const bool up_aq = sd.upstream_is_aqueous[irxn] != 0;
    const int uc = sd.upstream_c_id[irxn];
    const int ires_uc = up_aq ? uc : off + uc;
    const int ires_co2 = row(sd.co2_id, sd.co2_itype);
    if (wrt_uc) {
      const double f = up_aq ? s.DT(sd.co2_id, uc) : 1.0;
      if (up_aq)
        s.J(ires_co2, jcol) = s.J(ires_co2, jcol) - dco2_dx * f;
      else
        s.J(ires_co2, jcol) = s.J(ires_co2, jcol) - dco2_dx;
      if (sd.o2_id >= 0) {
        if (up_aq)
          s.J(ires_ox, jcol) = s.J(ires_ox, jcol) + dco2_dx * f;
        else
          s.J(ires_ox, jcol) = s.J(ires_ox, jcol) + dco2_dx;
      }
    } else {
      s.J(ires_co2, jcol) = s.J(ires_co2, jcol) - dco2_dx * s.DT(sd.co2_id, jaq);
    }
    if (sd.upstream_hr_id[irxn] >= 0) s.J(off + sd.upstream_hr_id[irxn], jcol) -= dco2_dx;
    if (sd.hr_id >= 0) s.J(off + sd.hr_id, jcol) -= dco2_dx;
    if (up_aq)
      s.J(ires_uc, jcol) = s.J(ires_uc, jcol) - duc_dx * s.DT(uc, wrt_uc ? uc : jaq);
    else
      s.J(ires_uc, jcol) = s.J(ires_uc, jcol) - duc_dx;
#pragma unroll 1
    for (int j = sd.downstream_ptr[irxn]; j < sd.downstream_ptr[irxn + 1]; j++) {
      const int dc = sd.downstream_c_id[j];
      const bool d_aq = sd.downstream_is_aqueous[j] != 0;
      const double ddc = sd.downstream_stoich[j] * (-1.0 * duc_dx);
      if (wrt_uc) {
        const int r = d_aq ? dc : off + dc;
        if (up_aq && d_aq)
          s.J(r, jcol) = s.J(r, jcol) - ddc * s.DT(dc, uc);
        else
          s.J(r, jcol) = s.J(r, jcol) - ddc;
      } else {
        if (d_aq)
          s.J(dc, jcol) = s.J(dc, jcol) - ddc * s.DT(dc, jaq);
        else
          s.J(off + dc, jcol) = s.J(off + dc, jcol) - ddc;
      }
    }
  }

  // upstream-N and downstream-N rows of column jcol
  __device__ __forceinline__ void n_rows_jacobian(int irxn, int jcol, int jaq, double duc_dx, double dun_dx,
                                                  bool wrt_uc) {
    const bool up_aq = sd.upstream_is_aqueous[irxn] != 0;
    const int uc = sd.upstream_c_id[irxn];
    if (sd.upstream_n_id[irxn] >= 0) {
      const int un = sd.upstream_n_id[irxn];
      if (up_aq)
        s.J(un, jcol) = s.J(un, jcol) - dun_dx * s.DT(un, wrt_uc ? uc : jaq);
      else
        s.J(off + un, jcol) = s.J(off + un, jcol) - dun_dx;
    }
#pragma unroll 1
    for (int j = sd.downstream_ptr[irxn]; j < sd.downstream_ptr[irxn + 1]; j++) {
      const int dn = sd.downstream_n_id[j];
      if (dn >= 0) {
        const bool d_aq = sd.downstream_is_aqueous[j] != 0;
        const int r = d_aq ? dn : off + dn;
        const double ddn = sd.downstream_stoich[j] * (-1.0 * duc_dx) * s.NC(sd.nrxn + j);
        if (up_aq && d_aq)
          s.J(r, jcol) = s.J(r, jcol) - ddn * s.DT(dn, wrt_uc ? uc : jaq);
        else
          s.J(r, jcol) = s.J(r, jcol) - ddn;
      }
    }
  }

  // SomDecReact1 (:1914-2418): N mineralisation
  __device__ __forceinline__ double react1(int irxn, int rxn, double cst, double nst, double unc, double crate_uc,
                                           double dcrate_uc_duc) {
    const int uc = sd.upstream_c_id[irxn];
    const bool up_aq = sd.upstream_is_aqueous[irxn] != 0;
    const int ires_uc = up_aq ? uc : off + uc;
    const int ires_nh4 = sd.nh4_id;
    double d1 = 1.0, d2 = 0.0, d3 = 1.0, d4 = 0.0;
    modifiers(rxn, uc, false, 0.0, 0.0, 0.0, crate_uc, dcrate_uc_duc, d1, d2, d3, d4);
    const double crate = crate_uc;
    const int ires_ox = common_residual(irxn, crate, cst, unc);
    s.RES(ires_nh4) = s.RES(ires_nh4) - nst * crate;
    const double nmin = nst * crate;
    if (sd.upstream_nmin_id[irxn] >= 0) s.RES(off + sd.upstream_nmin_id[irxn]) -= nst * crate;
    if (sd.nmin_id >= 0) s.RES(off + sd.nmin_id) -= nst * crate;
    downstream_n_residual(irxn, crate);
    // Jacobian
    const double dco2_duc = dcrate_uc_duc * cst;
    const double duc_duc = -1.0 * dcrate_uc_duc;
    const double dnh4_duc = dco2_duc * nst;  // sic, reaction_sandbox_somdec.F90:2283
    const double dun_duc = unc * duc_duc;
    common_jacobian(irxn, ires_uc, -1, ires_ox, dco2_duc, duc_duc, true);
    if (up_aq)
      s.J(ires_nh4, ires_uc) = s.J(ires_nh4, ires_uc) - dnh4_duc * s.DT(sd.nh4_id, uc);
    else
      s.J(ires_nh4, ires_uc) = s.J(ires_nh4, ires_uc) - dnh4_duc;
    if (sd.upstream_nmin_id[irxn] >= 0) s.J(off + sd.upstream_nmin_id[irxn], ires_uc) -= dnh4_duc;
    if (sd.nmin_id >= 0) s.J(off + sd.nmin_id, ires_uc) -= dnh4_duc;
    n_rows_jacobian(irxn, ires_uc, -1, duc_duc, dun_duc, true);
    return nmin;
  }

  // SomDecReact2 (:2423-3472): N immobilisation from NH4+ / NO3-
  __device__ __forceinline__ double react2(int irxn, int rxn, double cst, double ns, double unc, double crate_uc,
                                           double dcrate_uc_duc) {
    const int uc = sd.upstream_c_id[irxn];
    const bool up_aq = sd.upstream_is_aqueous[irxn] != 0;
    const int ires_uc = up_aq ? uc : off + uc;
    const int ires_nh4 = sd.nh4_id, ires_no3 = sd.no3_id;
    const double theta = s.sat * s.por, volume = s.vol;
    double c_nh4 = 0.0, c_no3 = 0.0;
    if (sd.nh4_id >= 0) c_nh4 = s.TOTc(sd.nh4_id) * theta * 1000.0;
    if (sd.no3_id >= 0) c_no3 = s.TOTc(sd.no3_id) * theta * 1000.0;
    double finh = 1.0;  // fnh4_inhibit_no3; its derivatives are zero in the reference
    if (sd.inhibition_nh4_no3 > 0.0) {
      if (c_nh4 > sd.x0eps && c_no3 > sd.x0eps) {
        finh = monod(c_nh4 / c_no3, 1.0 / sd.inhibition_nh4_no3);
      } else {
        if (c_nh4 > sd.x0eps && c_no3 <= sd.x0eps)
          finh = 1.0;
        else if (c_nh4 <= sd.x0eps && c_no3 > sd.x0eps)
          finh = 0.0;
        else
          return 0.0;
      }
    }
    double fnh4 = 1.0, dfnh4 = 0.0, fno3 = 1.0, dfno3 = 0.0;
    modifiers(rxn, uc, true, theta, c_nh4, c_no3, crate_uc, dcrate_uc_duc, fnh4, dfnh4, fno3, dfno3);
    double feps0, dfeps0;
    if (sd.nh4_id >= 0) {
      if (sd.x0eps > 0.0) {
        hsmooth(c_nh4, sd.x0eps * 10.0, sd.x0eps, feps0, dfeps0);
      } else {
        feps0 = 1.0;
        dfeps0 = 0.0;
      }
      dfnh4 = dfnh4 * feps0 + fnh4 * dfeps0;
      fnh4 = fnh4 * feps0;
    }
    if (sd.no3_id >= 0) {
      if (sd.x0eps > 0.0) {
        hsmooth(c_no3, sd.x0eps * 10.0, sd.x0eps, feps0, dfeps0);
      } else {
        feps0 = 1.0;
        dfeps0 = 0.0;
      }
      dfno3 = dfno3 * feps0 + fno3 * dfeps0;
      fno3 = fno3 * feps0;
    }
    const double nratecap = -crate_uc * ns * tran_dt / 0.45;
    if (sd.nh4_id >= 0) {
      double fcap = 1.0, dfcap = 0.0;
      if (nratecap * finh > c_nh4 * volume) {
        fcap = monod(c_nh4 * volume, nratecap * finh - c_nh4 * volume);
        dfcap = dmonod(c_nh4 * volume, nratecap * finh - c_nh4 * volume);
      }
      dfnh4 = dfnh4 * fcap + fnh4 * dfcap;
      fnh4 = fnh4 * fcap;
    }
    if (sd.no3_id >= 0) {
      double fcap = 1.0, dfcap = 0.0;
      if (nratecap * (1.0 - finh) > c_no3 * volume) {
        fcap = monod(c_no3 * volume, nratecap * (1.0 - finh) - c_no3 * volume);
        dfcap = dmonod(c_no3 * volume, nratecap * (1.0 - finh) - c_no3 * volume);
      }
      dfno3 = dfno3 * fcap + fno3 * dfcap;
      fno3 = fno3 * fcap;
    }
    const double crate_nh4 = crate_uc * fnh4 * finh;
    const double crate_no3 = crate_uc * fno3 * (1.0 - finh);
    const double crate = crate_nh4 + crate_no3;
    const int ires_ox = common_residual(irxn, crate, cst, unc);
    double nimm = 0.0;
    if (sd.nh4_id >= 0) {
      s.RES(ires_nh4) = s.RES(ires_nh4) - ns * crate_nh4;
      nimm = nimm + ns * crate_nh4;
    }
    if (sd.no3_id >= 0) {
      s.RES(ires_no3) = s.RES(ires_no3) - ns * crate_no3;
      nimm = nimm + ns * crate_no3;
    }
    const int unimm = sd.upstream_nimm_id[irxn];
    if (unimm >= 0) s.RES(off + unimm) += ns * crate;
    if (sd.upstream_nimp_id[irxn] >= 0) s.RES(off + sd.upstream_nimp_id[irxn]) += ns * crate_uc;
    if (sd.nimm_id >= 0) s.RES(off + sd.nimm_id) += ns * crate;
    if (sd.nimp_id >= 0) s.RES(off + sd.nimp_id) += ns * crate_uc;
    downstream_n_residual(irxn, crate);

    // ---- Jacobian (:3000-3470); d(fnh4_inhibit_no3)/d(nh4|no3) are 0 in the reference
    const double dz = 0.0;
    double dcrate_dx = dcrate_uc_duc * (fnh4 * finh + fno3 - fno3 * finh);
    const double dco2_duc = dcrate_dx * cst;
    const double duc_duc = -1.0 * dcrate_dx;
    const double dnh4_duc = dcrate_uc_duc * ns * fnh4 * finh;
    const double dno3_duc = dcrate_uc_duc * ns * fno3 * (1.0 - finh);
    const double dun_duc = unc * duc_duc;
    dcrate_dx = (dfnh4 * finh + (fnh4 - fno3) * dz);
    dcrate_dx = dcrate_dx * crate_uc;
    const double dco2_dnh4 = dcrate_dx * cst;
    const double duc_dnh4 = -1.0 * dcrate_dx;
    double dnh4_dnh4 = fnh4 * dz + dfnh4 * finh;
    dnh4_dnh4 = dnh4_dnh4 * crate_uc * ns;
    double dno3_dnh4 = -1.0 * fno3 * dz;
    dno3_dnh4 = dno3_dnh4 * crate_uc * ns;
    const double dun_dnh4 = unc * duc_dnh4;
    dcrate_dx = (fnh4 - fno3) * dz + dfno3 * (1.0 - fnh4 * finh);
    dcrate_dx = dcrate_dx * crate_uc;
    const double dco2_dno3 = dcrate_dx * cst;
    const double duc_dno3 = -1.0 * dcrate_dx;
    const double dnh4_dno3 = fnh4 * dz * crate_uc * ns;
    double dno3_dno3 = -1.0 * fno3 * dz + dfno3 * (1.0 - finh);
    dno3_dno3 = dno3_dno3 * crate_uc * ns;
    const double dun_dno3 = unc * duc_dno3;

    // column uc
    common_jacobian(irxn, ires_uc, -1, ires_ox, dco2_duc, duc_duc, true);
    if (sd.nh4_id >= 0) {
      if (up_aq)
        s.J(ires_nh4, ires_uc) = s.J(ires_nh4, ires_uc) - dnh4_duc * s.DT(sd.nh4_id, uc);
      else
        s.J(ires_nh4, ires_uc) = s.J(ires_nh4, ires_uc) - dnh4_duc;
      if (unimm >= 0) s.J(off + unimm, ires_uc) += dnh4_duc;
      if (sd.nimm_id >= 0) s.J(off + sd.nimm_id, ires_uc) += dnh4_duc;
    }
    if (sd.no3_id >= 0) {
      if (up_aq)
        s.J(ires_no3, ires_uc) = s.J(ires_no3, ires_uc) - dno3_duc * s.DT(sd.no3_id, uc);
      else
        s.J(ires_no3, ires_uc) = s.J(ires_no3, ires_uc) - dno3_duc;
      if (unimm >= 0) s.J(off + unimm, ires_uc) += dno3_duc;
      if (sd.nimm_id >= 0) s.J(off + sd.nimm_id, ires_uc) += dno3_duc;
    }
    n_rows_jacobian(irxn, ires_uc, -1, duc_duc, dun_duc, true);
    // column nh4
    if (sd.nh4_id >= 0) {
      common_jacobian(irxn, ires_nh4, sd.nh4_id, ires_ox, dco2_dnh4, duc_dnh4, false);
      s.J(ires_nh4, ires_nh4) = s.J(ires_nh4, ires_nh4) - dnh4_dnh4 * s.DT(sd.nh4_id, sd.nh4_id);
      if (unimm >= 0) s.J(off + unimm, ires_nh4) += dnh4_dnh4;
      if (sd.nimm_id >= 0) s.J(off + sd.nimm_id, ires_nh4) += dnh4_dnh4;
      if (sd.no3_id >= 0) {
        s.J(ires_no3, ires_nh4) = s.J(ires_no3, ires_nh4) - dno3_dnh4 * s.DT(sd.no3_id, sd.nh4_id);
        if (unimm >= 0) s.J(off + unimm, ires_nh4) += dno3_dnh4;
        if (sd.nimm_id >= 0) s.J(off + sd.nimm_id, ires_nh4) += dno3_dnh4;
      }
      n_rows_jacobian(irxn, ires_nh4, sd.nh4_id, duc_dnh4, dun_dnh4, false);
    }
    // column no3
    if (sd.no3_id >= 0) {
      common_jacobian(irxn, ires_no3, sd.no3_id, ires_ox, dco2_dno3, duc_dno3, false);
      if (sd.nh4_id >= 0) {
        s.J(ires_nh4, ires_no3) = s.J(ires_nh4, ires_no3) - dnh4_dno3 * s.DT(sd.nh4_id, sd.no3_id);
        if (unimm >= 0) s.J(off + unimm, ires_no3) += dnh4_dno3;
        if (sd.nimm_id >= 0) s.J(off + sd.nimm_id, ires_no3) += dnh4_dno3;
      }
      s.J(ires_no3, ires_no3) = s.J(ires_no3, ires_no3) - dno3_dno3 * s.DT(sd.no3_id, sd.no3_id);
      if (unimm >= 0) s.J(off + unimm, ires_no3) += dno3_dno3;
      if (sd.nimm_id >= 0) s.J(off + sd.nimm_id, ires_no3) += dno3_dno3;
      n_rows_jacobian(irxn, ires_no3, sd.no3_id, duc_dno3, dun_dno3, false);
    }
    return nimm;
  }

  // SomDecNemission (:3477-3640)
  __device__ __forceinline__ void nemission(double net_nmin_rate) {
    const double saturation = s.sat, theta = s.sat * s.por, volume = s.vol, tc = s.temp;
    const int ires_nh4 = sd.nh4_id, ires_n2o = sd.n2o_id;
    if (!(sd.n2o_id >= 0 && net_nmin_rate > sd.x0eps)) return;
    const double c_nh4 = s.TOTc(ires_nh4) * theta * 1000.0;
    double f_t = -0.06 + 0.13 * exp(0.07 * tc);
    double f_w = wfps(saturation);
    double fph = f_ph(s, sd.proton_id);
    if (f_t > sd.x0eps && f_w > sd.x0eps && fph > sd.x0eps) {
      f_t = fmin(f_t, 1.0);
      f_w = fmin(f_w, 1.0);
      fph = fmin(fph, 1.0);
      const double temp_real = f_t * f_w * fph;
      double feps0, dfeps0;
      if (sd.x0eps > 0.0) {
        hsmooth(c_nh4, sd.x0eps * 10.0, sd.x0eps, feps0, dfeps0);
      } else {
        feps0 = 1.0;
        dfeps0 = 0.0;
      }
      const double nratecap = temp_real * sd.n2o_frac_mineralization * net_nmin_rate * tran_dt;
      double fcap = 1.0, dfcap = 0.0;
      if (nratecap > c_nh4 * volume) {
        fcap = monod(c_nh4 * volume, nratecap - c_nh4 * volume);
        dfcap = dmonod(c_nh4 * volume, nratecap - c_nh4 * volume);
      }
      dfeps0 = dfeps0 * fcap + feps0 * dfcap;
      feps0 = feps0 * fcap;
      const double rate_n2o = temp_real * sd.n2o_frac_mineralization * net_nmin_rate * feps0;
      s.RES(ires_nh4) = s.RES(ires_nh4) + rate_n2o;
      s.RES(ires_n2o) = s.RES(ires_n2o) - 0.5 * rate_n2o;
      if (sd.ngasmin_id >= 0) s.RES(off + sd.ngasmin_id) -= rate_n2o;
      const double drate = temp_real * sd.n2o_frac_mineralization * net_nmin_rate * dfeps0;
      s.J(ires_nh4, ires_nh4) = s.J(ires_nh4, ires_nh4) + drate * s.DT(sd.nh4_id, sd.nh4_id);
      s.J(ires_n2o, ires_nh4) = s.J(ires_n2o, ires_nh4) - 0.5 * drate * s.DT(sd.n2o_id, sd.nh4_id);
      if (sd.ngasmin_id >= 0) s.J(off + sd.ngasmin_id, ires_nh4) -= drate;
    }
  }

  // SomDecReact (:1504-1910)
  __device__ __forceinline__ void react() {
    const double saturation = s.sat, theta = s.sat * s.por, volume = s.vol, tc = s.temp;
    const bool elm = s.cfg.elm != 0;
    double net_nmin_rate = 0.0;
    int cur = 0;  // the reference's cur_rxn: not advanced by `cycle` (:1744,1783 vs :1869)
#pragma unroll 1
    for (int irxn = 0; irxn < sd.nrxn; irxn++) {
      double f_w, f_t, f_depth = 1.0, kd_scalar = 1.0;
      if (elm && s.cfg.elm_flow && sd.moisture_response_function[cur] != PFRX_MOISTURE_RESPONSE_OFF) {
        f_w = elm_moisture_response(theta, sd.moisture_response_function[cur], s.elm_sucsat, s.elm_bd_dry, s.elm_bsw,
                                    s.elm_watfc, s.elm_effpor);
      } else if (elm) {
        f_w = s.elm_w;
      } else if (sd.moisture_response_function[cur] == PFRX_MOISTURE_RESPONSE_LOGTHETA) {
        // single-precision literals of the reference (:1645-1649)
        if (theta <= (double)0.08f)
          f_w = (double)0.01f;
        else
          f_w = log(theta / (double)0.08f) / 2.525728702545166015625;  // (double)logf(1.0f/0.08f)
      } else {
        f_w = 1.0;
      }
      if (sd.ox_response_function[cur] == PFRX_OX_RESPONSE_WFPS)
        f_w = f_w * wfps(saturation);
      else if (elm)
        f_w = f_w * s.elm_o;
      const int tf = sd.temperature_response_function[cur];
      if (tf == PFRX_TEMPERATURE_RESPONSE_ARRHENIUS)
        f_t = temperature_response(tc, tf, sd.ea[cur]);
      else if (tf == PFRX_TEMPERATURE_RESPONSE_CLMCN)
        f_t = temperature_response(tc, tf, 0.0);
      else if (tf == PFRX_TEMPERATURE_RESPONSE_Q10 || tf == PFRX_TEMPERATURE_RESPONSE_DLEM)
        f_t = temperature_response(tc, tf, sd.q10[cur]);
      else
        f_t = elm ? s.elm_t : 1.0;
      if (elm) {
        if (sd.decomp_depth_efolding[cur] > 0.0) {
          f_depth = exp(-s.elm_zsoil / sd.decomp_depth_efolding[cur]);
          f_depth = fmin(1.0, fmax(1.e-20, f_depth));
        }
        kd_scalar = s.elm_kscalar;
      }
      if (f_t < 1.0e-20 || f_w < 1.0e-20 || f_depth < 1.0e-20) continue;
      double k_decomp = 0.0;
      if (sd.rate_constant[irxn] >= 0.0) {
        k_decomp = sd.rate_constant[irxn];
      } else if (sd.rate_decomposition[irxn] >= 0.0) {
        k_decomp = 1.0 - exp(-sd.rate_decomposition[irxn] * tran_dt);
        k_decomp = k_decomp / tran_dt;
      }
      k_decomp = sd.rate_ad_factor[irxn] * k_decomp;
      if (kd_scalar > 0.0 && sd.rate_ad_factor[irxn] > 1.0) k_decomp = k_decomp / kd_scalar;
      k_decomp = fmin(k_decomp, 1.0 / tran_dt);
      const double scaled = k_decomp * volume * f_t * f_w * f_depth;
      const int uc = sd.upstream_c_id[irxn];
      const bool up_aq = sd.upstream_is_aqueous[irxn] != 0;
      double c_uc;
      if (up_aq) {
        c_uc = s.TOTc(uc);
        c_uc = theta * 1000.0 * c_uc;
      } else {
        c_uc = s.Cc(off + uc);
      }
      double feps0, dfeps0;
      if (sd.x0eps > 0.0) {
        hsmooth(c_uc, sd.x0eps * 10.0, sd.x0eps, feps0, dfeps0);
      } else {
        feps0 = 1.0;
        dfeps0 = 0.0;
        if (c_uc <= sd.x0eps) continue;
      }
      const double crate_uc = scaled * c_uc * feps0;
      const double dcrate_uc_duc = scaled * (feps0 + c_uc * dfeps0);
      // N:C ratios on the fly; NC() keeps the last valid ones (:1795-1822)
      const int d0 = sd.downstream_ptr[irxn], d1 = sd.downstream_ptr[irxn + 1];
#pragma unroll 1
      for (int j = d0; j < d1; j++) {
        const int dc = sd.downstream_c_id[j], dn = sd.downstream_n_id[j];
        if (dn >= 0 && dc >= 0) {
          double c_dc, c_dn;
          if (sd.downstream_is_aqueous[j]) {
            c_dc = theta * 1000.0 * s.TOTc(dc);
            c_dn = theta * 1000.0 * s.TOTc(dn);
          } else {
            c_dc = s.Cc(off + dc);
            c_dn = s.Cc(off + dn);
          }
          if (c_dn >= sd.x0eps && c_dc >= sd.x0eps) s.NC(sd.nrxn + j) = c_dn / c_dc;
        }
      }
      double unc = s.NC(irxn), cst = sd.mineral_c_stoich[irxn], nst = sd.mineral_n_stoich[irxn];
      if (sd.upstream_n_id[irxn] >= 0) {
        const int un = sd.upstream_n_id[irxn];
        double c_un;
        if (up_aq)
          c_un = theta * 1000.0 * s.TOTc(un);
        else
          c_un = s.Cc(off + un);
        if (c_un >= sd.x0eps && c_uc >= sd.x0eps) {
          unc = c_un / c_uc;
          s.NC(irxn) = unc;
        }
        cst = 1.0;
#pragma unroll 1
        for (int j = d0; j < d1; j++) cst = cst - sd.downstream_stoich[j];
        nst = unc;
#pragma unroll 1
        for (int j = d0; j < d1; j++) nst = nst - sd.downstream_stoich[j] * s.NC(sd.nrxn + j);
      }
      if (nst >= 0.0)
        net_nmin_rate = net_nmin_rate + react1(irxn, cur, cst, nst, unc, crate_uc, dcrate_uc_duc);
      else
        net_nmin_rate = net_nmin_rate + react2(irxn, cur, cst, nst, unc, crate_uc, dcrate_uc_duc);
      cur++;
    }
    if (net_nmin_rate > sd.x0eps) nemission(net_nmin_rate);
  }
};

// ---- NITRIFICATION -------------------------------------------------------------------
template <class Cell>
__device__ __forceinline__ void nitrif_react(Cell &s) {
  const pfrx_nitrif &nt = s.cfg.nt;
  const int off = s.cfg.naq;
  const double volume = s.vol, tc = s.temp;
  double saturation = s.sat;
  const double L_water = saturation * s.por * 1.0e3;
  const int ires_nh4 = nt.nh4_id, ires_no3 = nt.no3_id, ires_n2o = nt.n2o_id;
  const double c_nh4 = s.TOTc(nt.nh4_id) * L_water;
  double feps0, dfeps0;
  if (nt.x0eps > 0.0) {
    hsmooth(c_nh4, nt.x0eps * 10.0, nt.x0eps, feps0, dfeps0);
  } else {
    feps0 = 1.0;
    dfeps0 = 0.0;
    if (c_nh4 < nt.x0eps) return;
  }
  if (nt.nh4_id >= 0 && nt.no3_id >= 0) {
    const double f_t = exp(0.08 * (tc - 25.0));
    saturation = fmax(0.0, fmin(saturation, 1.0));
    const double f_w = saturation * (1.0 - saturation) / 0.25;
    double t = fmin(nt.k_nitr_max * f_t * f_w * volume, 1.0);
    const double rate = t * (c_nh4 * feps0) * (c_nh4 / (c_nh4 + 4.0));
    s.RES(ires_nh4) = s.RES(ires_nh4) + rate;
    s.RES(ires_no3) = s.RES(ires_no3) - rate;
    t = c_nh4 * c_nh4 / (c_nh4 + 4.0) * dfeps0 + c_nh4 * (c_nh4 + 8.0) / (c_nh4 + 4.0) / (c_nh4 + 4.0) * feps0;
    const double drate = nt.k_nitr_max * f_t * f_w * volume * t;
    s.J(ires_nh4, ires_nh4) = s.J(ires_nh4, ires_nh4) + drate * s.DT(nt.nh4_id, nt.nh4_id);
    s.J(ires_no3, ires_nh4) = s.J(ires_no3, ires_nh4) - drate * s.DT(nt.no3_id, nt.nh4_id);
  }
  const double rho_b = s.cfg.elm ? s.elm_bd_dry : 1.25e3;
  const double M_2_ug_per_g = (14.0067 * 1.0e6) / (volume * rho_b * 1.e3);
  const double c_nh4_ugg = c_nh4 * volume * M_2_ug_per_g;
  if (nt.n2o_id >= 0 && c_nh4_ugg > 3.0) {
    double f_t = -0.06 + 0.13 * exp(0.07 * tc);
    double f_w = wfps(saturation);
    double fph = f_ph(s, nt.proton_id);
    if (f_t > 0.0 && f_w > 0.0 && fph > 0.0) {
      f_t = fmin(f_t, 1.0);
      f_w = fmin(f_w, 1.0);
      fph = fmin(fph, 1.0);
      const double ex = exp(-0.0105 * c_nh4_ugg);
      double t = (1.0 - ex) * f_t * f_w * fph * nt.k_nitr_n2o;
      const double rate_n2o = t * (c_nh4 * feps0) * volume;
      s.RES(ires_nh4) = s.RES(ires_nh4) + rate_n2o;
      s.RES(ires_n2o) = s.RES(ires_n2o) - 0.5 * rate_n2o;
      if (nt.ngasnit_id >= 0) s.RES(off + nt.ngasnit_id) -= rate_n2o;
      t = (c_nh4 * dfeps0 + feps0) * (1.0 - ex);
      t = t + (c_nh4 * feps0) * 0.0105 * M_2_ug_per_g * ex;
      const double drate = t * nt.k_nitr_n2o * f_t * f_w * fph * volume;
      s.J(ires_nh4, ires_nh4) = s.J(ires_nh4, ires_nh4) + drate * s.DT(nt.nh4_id, nt.nh4_id);
      s.J(ires_n2o, ires_nh4) = s.J(ires_n2o, ires_nh4) - 0.5 * drate * s.DT(nt.n2o_id, nt.nh4_id);
      if (nt.ngasnit_id >= 0) s.J(off + nt.ngasnit_id, ires_nh4) -= drate;
    }
  }
}

// ---- DENITRIFICATION -----------------------------------------------------------------
template <class Cell>
__device__ __forceinline__ void denitr_react(Cell &s) {
  const pfrx_denitr &dn = s.cfg.dn;
  const int off = s.cfg.naq;
  if (dn.n2_id < 0) return;
  const double volume = s.vol, saturation = s.sat, tc = s.temp;
  const double L_water = s.por * saturation * 1.e3;
  const int ires_no3 = dn.no3_id, ires_n2 = dn.n2_id;
  const double bsw = s.cfg.elm ? s.elm_bsw : 1.0;
  const double f_t = exp(0.08 * (tc - 25.0));
  const double s_min = 0.6;
  double f_w = 0.0;
  if (saturation > s_min) {
    f_w = (saturation - s_min) / (1.0 - s_min);
    f_w = pow(f_w, bsw);
  }
  const double c_no3 = s.TOTc(ires_no3) * L_water;
  double feps0, dfeps0;
  if (dn.x0eps > 0.0) {
    hsmooth(c_no3, dn.x0eps * 10.0, dn.x0eps, feps0, dfeps0);
  } else {
    feps0 = 1.0;
    dfeps0 = 0.0;
    if (c_no3 <= dn.x0eps) return;
  }
  double fno3 = 1.0, dfno3 = 0.0;
  if (dn.half_saturation > 0.0) {
    fno3 = monod(c_no3, dn.half_saturation);
    dfno3 = dmonod(c_no3, dn.half_saturation);
  }
  if (f_t > 0.0 && f_w > 0.0) {
    const double rate = dn.k_deni_max * f_t * f_w * fno3 * (c_no3 * volume * feps0);
    s.RES(ires_no3) = s.RES(ires_no3) + rate;
    s.RES(ires_n2) = s.RES(ires_n2) - 0.5 * rate;
    if (dn.ngasdeni_id >= 0) s.RES(off + dn.ngasdeni_id) -= rate;
    const double t = dfno3 * (c_no3 * volume * feps0) + fno3 * (c_no3 * volume * dfeps0 + feps0);
    const double drate = dn.k_deni_max * f_t * f_w * t;
    s.J(ires_no3, ires_no3) = s.J(ires_no3, ires_no3) + drate * s.DT(dn.no3_id, dn.no3_id);
    s.J(ires_n2, ires_no3) = s.J(ires_n2, ires_no3) - 0.5 * drate * s.DT(dn.n2_id, dn.no3_id);
    if (dn.ngasdeni_id >= 0) s.J(off + dn.ngasdeni_id, ires_no3) -= drate;
  }
}

// ---- PLANTN (reaction_sandbox_plantn.F90:222-640) -------------------------------------
template <class Cell>
__device__ __forceinline__ void plantn_react(Cell &s, double tran_dt) {
  const pfrx_plantn &pn = s.cfg.pn;
  const int off = s.cfg.naq;
  const double volume = s.vol, saturation = s.sat, tc = s.temp;
  if (saturation < 0.01) return;
  const double L_water = saturation * s.por * 1.0e3;
  if (tc < -0.1) return;
  const int ires_plantn = off + pn.plantn_id, ires_nh4 = pn.nh4_id, ires_no3 = pn.no3_id;
  double c_nh4 = 0.0, c_no3 = 0.0, fnh4 = 1.0, dfnh4 = 0.0, fno3 = 1.0, dfno3 = 0.0, finh = 1.0;
  if (pn.nh4_id >= 0 && pn.no3_id >= 0) {
    c_nh4 = s.TOTc(pn.nh4_id) * L_water;
    c_no3 = s.TOTc(pn.no3_id) * L_water;
    if ((c_nh4 > pn.x0eps_nh4 && c_no3 > pn.x0eps_no3) && pn.inhibition_nh4_no3 > 0.0) {
      finh = monod(c_nh4 / c_no3, 1.0 / pn.inhibition_nh4_no3);
    } else {
      if (c_nh4 > pn.x0eps_nh4 && c_no3 <= pn.x0eps_no3)
        finh = 1.0;
      else if (c_nh4 <= pn.x0eps_nh4 && c_no3 > pn.x0eps_no3)
        finh = 0.0;
      else
        return;
    }
  }
  double feps0, dfeps0;
  if (pn.nh4_id >= 0) {
    c_nh4 = s.TOTc(pn.nh4_id) * L_water;
    fnh4 = monod(c_nh4, pn.half_saturation_nh4);
    dfnh4 = dmonod(c_nh4, pn.half_saturation_nh4);
    if (pn.x0eps_nh4 > 0.0) {
      hsmooth(c_nh4, pn.x0eps_nh4 * 10.0, pn.x0eps_nh4, feps0, dfeps0);
    } else {
      feps0 = 1.0;
      dfeps0 = 0.0;
    }
    dfnh4 = dfnh4 * feps0 + fnh4 * dfeps0;
    fnh4 = fnh4 * feps0;
  }
  if (pn.no3_id >= 0) {
    c_no3 = s.TOTc(pn.no3_id) * L_water;
    fno3 = monod(c_no3, pn.half_saturation_no3);
    dfno3 = dmonod(c_no3, pn.half_saturation_no3);
    if (pn.x0eps_no3 > 0.0) {
      hsmooth(c_no3, pn.x0eps_no3 * 10.0, pn.x0eps_no3, feps0, dfeps0);
    } else {
      feps0 = 1.0;
      dfeps0 = 0.0;
    }
    dfno3 = dfno3 * feps0 + fno3 * dfeps0;
    fno3 = fno3 * feps0;
  }
  double demand;
  if (s.cfg.elm) {
    demand = fmax(0.0, s.elm_plantndemand * volume);
    if (demand <= 0.0) return;
  } else {
    demand = 1.e-2 * volume;
  }
  if (pn.plantndemand_id >= 0) s.RES(off + pn.plantndemand_id) -= demand;
  if (demand > 0.0) {
    if (pn.nh4_id >= 0) {
      double cap = demand * tran_dt;
      if (pn.no3_id >= 0) cap = demand * finh * tran_dt;
      double fcap = 1.0, dfcap = 0.0;
      if (cap > c_nh4 * volume) {
        fcap = monod(c_nh4 * volume, cap - c_nh4 * volume);
        dfcap = dmonod(c_nh4 * volume, cap - c_nh4 * volume);
      }
      dfnh4 = dfnh4 * fcap + fnh4 * dfcap;
      fnh4 = fnh4 * fcap;
    }
    if (pn.no3_id >= 0) {
      double cap = demand * tran_dt;
      if (pn.nh4_id >= 0) cap = demand * (1.0 - finh) * tran_dt;
      double fcap = 1.0, dfcap = 0.0;
      if (cap > c_no3 * volume) {
        fcap = monod(c_no3 * volume, cap - c_no3 * volume);
        dfcap = dmonod(c_no3 * volume, cap - c_no3 * volume);
      }
      dfno3 = dfno3 * fcap + fno3 * dfcap;
      fno3 = fno3 * fcap;
    }
  }
  if (pn.nh4_id >= 0) {
    double nrate = demand * fnh4;
    if (pn.no3_id >= 0) nrate = demand * fnh4 * finh;
    s.RES(ires_nh4) = s.RES(ires_nh4) + nrate;
    s.RES(ires_plantn) = s.RES(ires_plantn) - nrate;
    if (pn.plantnh4uptake_id >= 0) s.RES(off + pn.plantnh4uptake_id) -= nrate;
    double dn = demand * dfnh4;
    if (pn.no3_id >= 0) dn = demand * (fnh4 * 0.0 + finh * dfnh4);
    s.J(ires_nh4, ires_nh4) = s.J(ires_nh4, ires_nh4) + dn * s.DT(pn.nh4_id, pn.nh4_id);
    s.J(ires_plantn, ires_nh4) = s.J(ires_plantn, ires_nh4) - dn;
    if (pn.plantnh4uptake_id >= 0) s.J(off + pn.plantnh4uptake_id, ires_nh4) -= dn;
  }
  if (pn.no3_id >= 0) {
    double nrate = demand * fno3;
    if (pn.nh4_id >= 0) nrate = demand * fno3 * (1.0 - finh);
    s.RES(ires_no3) = s.RES(ires_no3) + nrate;
    s.RES(ires_plantn) = s.RES(ires_plantn) - nrate;
    if (pn.plantno3uptake_id >= 0) s.RES(off + pn.plantno3uptake_id) -= nrate;
    double dn = demand * dfno3;
    if (pn.nh4_id >= 0) dn = demand * (dfno3 * (1.0 - finh) + fno3 * (-1.0 * 0.0));
    s.J(ires_no3, ires_no3) = s.J(ires_no3, ires_no3) + dn * s.DT(pn.no3_id, pn.no3_id);
    s.J(ires_plantn, ires_no3) = s.J(ires_plantn, ires_no3) - dn;
    if (pn.plantno3uptake_id >= 0) s.J(off + pn.plantno3uptake_id, ires_no3) -= dn;
  }
}

// ---- LANGMUIR (reaction_sandbox_langmu.F90:183-330) ------------------------------------
template <class Cell>
__device__ __forceinline__ void langmuir_react(Cell &s, double tran_dt) {
  const pfrx_langmuir &lg = s.cfg.lg;
  const int off = s.cfg.naq;
  const double volume = s.vol;
  const double Lwater = volume * 1000.0 * s.por * s.sat;
  const int ires_aq = lg.aq_id, ires_sorb = off + lg.sorb_id;
  const double c_aq = s.TOTc(lg.aq_id), c_sorb = s.Cc(off + lg.sorb_id);
  double rate, drate_daq, drate_dsorb;
  if (lg.s_max < c_sorb) {
    rate = (lg.s_max - c_sorb) * volume / tran_dt;
    drate_dsorb = -1.0 / tran_dt;
    drate_daq = 0.0;
  } else {
    const double c_aq_eq = 0.999 * c_sorb / (lg.s_max - 0.999 * c_sorb) / lg.k_equilibrium;
    rate = lg.k_kinetic * (c_aq - c_aq_eq) * Lwater;
    double t = -lg.k_kinetic / lg.k_equilibrium * Lwater / volume;
    drate_dsorb = t * lg.s_max / (lg.s_max - 0.999 * c_sorb) / (lg.s_max - 0.999 * c_sorb);
    drate_daq = lg.k_kinetic;
    if (rate > 0.0) {
      double ratecap = 0.999 * (lg.s_max - c_sorb) * volume / tran_dt;
      if (ratecap < rate) {
        const double fcap = ratecap / rate;
        t = -0.999 / tran_dt;
        const double dfcap = (ratecap * drate_dsorb - rate * t) / rate / rate;
        drate_dsorb = fcap * drate_dsorb + rate * dfcap;
        rate = rate * fcap;
      }
      ratecap = 0.999 * (c_aq - c_aq_eq) * Lwater / tran_dt;
      if (ratecap < rate) {
        const double fcap = ratecap / rate;
        t = -0.999 / lg.k_equilibrium * Lwater / volume / tran_dt;
        t = t * lg.s_max / (lg.s_max - 0.999 * c_sorb) / (lg.s_max - 0.999 * c_sorb);
        double dfcap = (ratecap * drate_dsorb - rate * t) / rate / rate;
        drate_dsorb = fcap * drate_dsorb + rate * dfcap;
        t = 0.999 / tran_dt;
        dfcap = (ratecap * drate_daq - rate * t) / rate / rate;
        drate_daq = fcap * drate_daq + rate * dfcap;
        rate = rate * fcap;
      }
    }
  }
  s.RES(ires_aq) = s.RES(ires_aq) + rate;
  s.RES(ires_sorb) = s.RES(ires_sorb) - rate;
  s.J(ires_aq, ires_aq) = s.J(ires_aq, ires_aq) + drate_daq * s.DT(lg.aq_id, lg.aq_id);
  s.J(ires_sorb, ires_aq) = s.J(ires_sorb, ires_aq) - drate_daq;
  s.J(ires_aq, ires_sorb) = s.J(ires_aq, ires_sorb) + drate_dsorb;
  s.J(ires_sorb, ires_sorb) = s.J(ires_sorb, ires_sorb) - drate_dsorb;
}

// ---- CALCITE (reaction_sandbox_calcite.F90:177-365) --------------------------------------
// two parallel TST pathways: #1 with the mineral's stoichiometry and logK from the kinmnrl tables,
// #2 written out for Ca++ + HCO3- - H+ with pKeq 1.8487; returns the sum of the two rates
// [mol/m^3 bulk/s] (rt_auxvar%auxiliary_data)
template <class Cell>
__device__ __forceinline__ double calcite_react(Cell &s) {
  const pfrx_calcite_sandbox &cs = s.cfg.cs;
  const int m = cs.mineral_id;
  const double volume = s.vol, molality_to_molarity = s.den_kg * 1.e-3;
  const double area = s.st.mnrl_area[m * s.st.ld + s.cell], volfrac = s.st.mnrl_volfrac[m * s.st.ld + s.cell];
  const int p0 = s.cfg.mn_ptr[m], p1 = s.cfg.mn_ptr[m + 1];
  // reaction path #1
  double lnQK = -s.mn_logK(m) * PFRX_LOG_TO_LN;
  if (s.cfg.mn_h2o[m] != 0.0) lnQK = lnQK + s.cfg.mn_h2o[m] * s.ln_act_h2o;
#pragma unroll 1
  for (int p = p0; p < p1; p++) lnQK = lnQK + s.cfg.mn_st[p] * s.LNAc(s.cfg.mn_id[p]);
  double QK = exp(lnQK);
  double affinity_factor = 1.0 - QK;
  double sign_ = copysign(1.0, affinity_factor);
  double rate = 0.0;
  bool calculate_rate = volfrac > 0.0 || sign_ < 0.0;
  if (calculate_rate) rate = -area * sign_ * fabs(affinity_factor) * cs.rate_constant1;
  double aux = rate;
  rate = rate * volume;
#pragma unroll 1
  for (int p = p0; p < p1; p++) s.RES(s.cfg.mn_id[p]) = s.RES(s.cfg.mn_id[p]) + s.cfg.mn_st[p] * rate;
  if (calculate_rate) {
    const double drate_dQK = area * cs.rate_constant1 * volume;
#pragma unroll 1
    for (int q = p0; q < p1; q++) {
      const int jcomp = s.cfg.mn_id[q];
      const double dQK_dmj = s.cfg.mn_st[q] * QK * exp(-log(s.Cc(jcomp))) * molality_to_molarity;
#pragma unroll 1
      for (int p = p0; p < p1; p++)
        s.J(s.cfg.mn_id[p], jcomp) = s.J(s.cfg.mn_id[p], jcomp) + s.cfg.mn_st[p] * drate_dQK * dQK_dmj;
    }
  }
  // reaction path #2
  const int ih = cs.h_ion_id, ica = cs.calcium_id, ib = cs.bicarbonate_id;
  lnQK = -1.8487 * PFRX_LOG_TO_LN - s.LNAc(ih) + s.LNAc(ica) + s.LNAc(ib);
  affinity_factor = 1.0 - exp(lnQK);
  sign_ = copysign(1.0, affinity_factor);
  rate = 0.0;
  calculate_rate = volfrac > 0.0 || sign_ < 0.0;
  if (calculate_rate) rate = -area * sign_ * fabs(affinity_factor) * cs.rate_constant2;
  aux = aux + rate;
  rate = rate * volume;
  s.RES(ih) = s.RES(ih) - rate;
  s.RES(ica) = s.RES(ica) + rate;
  s.RES(ib) = s.RES(ib) + rate;
  if (calculate_rate) {
    const double drate_dQK = area * cs.rate_constant2 * volume;
    const int jc[3] = {ih, ica, ib};
    const double sj[3] = {-1.0, 1.0, 1.0};
#pragma unroll
    for (int q = 0; q < 3; q++) {
      const int jcomp = jc[q];
      const double dQK_dmj = sj[q] * exp(lnQK - log(s.Cc(jcomp))) * molality_to_molarity;
      s.J(ih, jcomp) = s.J(ih, jcomp) - drate_dQK * dQK_dmj;
      s.J(ica, jcomp) = s.J(ica, jcomp) + drate_dQK * dQK_dmj;
      s.J(ib, jcomp) = s.J(ib, jcomp) + drate_dQK * dQK_dmj;
    }
  }
  return aux;
}

// ---- CNDEGAS (reaction_sandbox_cndegas.F90:216-546; solubilities :548-787) -------------
// rgas of the solubility routines is a default-real literal in the reference (0.08205601
// without d0): the single-precision value
#define PFRX_CND_RGAS ((double)0.08205601f)
__device__ __forceinline__ double weiss_co2_xmole(double tt, double tp, double ts, double pco2) {
  const double atm = 1.01325e5, xmwh2o = 18.01534e-3;
  const double tk = tt + 273.15, tk2 = tk * tk, tk3 = tk2 * tk, tk_100k = tk / 100.0;
  double p_rt = (tp / atm) / PFRX_CND_RGAS / tk;
  p_rt = p_rt / 1000.0;
  const double x1 = pco2 / tp, x2 = 1.0 - x1;
  const double epsilon = 57.7 - 0.118 * tk;
  const double bt = -1636.75 + 12.0408 * tk - 3.27957e-2 * tk2 + 3.16528e-5 * tk3;
  const double fg = pco2 * exp((bt + 2.0 * x2 * x2 * epsilon) * p_rt);
  double k0 = -58.0931 + 90.5069 / tk_100k + 22.2940 * log(tk_100k) +
              ts * (0.027766 + -0.025888 * tk_100k + 0.0050578 * tk_100k * tk_100k);
  k0 = exp(k0);
  k0 = k0 / atm;
  const double vbar = exp((1.0 - tp / atm) * 30.0e-3 / PFRX_CND_RGAS / tk);
  const double cco2 = k0 * fg * vbar;
  return cco2 / (1.0 / xmwh2o);
}
__device__ __forceinline__ double weiss_price_n2o_xmole(double tt, double tp, double ts, double pn2o) {
  const double atm = 1.01325e5, xmwh2o = 18.01534e-3;
  const double tk = tt + 273.15, tk2 = tk * tk, tk_100k = tk / 100.0;
  double p_rt = (tp / atm) / PFRX_CND_RGAS / tk;
  p_rt = p_rt / 1000.0;
  const double x1 = pn2o / tp, x2 = 1.0 - x1;
  const double epsilon = 65.0 - 0.1338 * tk;
  const double bt = -905.95 + 4.1685 * tk - 0.0052734 * tk2;
  const double fg = pn2o * exp((bt + 2.0 * x2 * x2 * epsilon) * p_rt);
  double k0 = -62.7076 + 97.3066 / tk_100k + 24.1406 * log(tk_100k) +
              ts * (-0.058420 + 0.033193 * tk_100k + -0.0051313 * tk_100k * tk_100k);
  k0 = exp(k0);
  k0 = k0 / atm;
  const double vbar = exp((1.0 - tp / atm) * 32.3e-3 / PFRX_CND_RGAS / tk);
  const double cn2o = k0 * fg * vbar;
  return cn2o / (1.0 / xmwh2o);
}
__device__ __forceinline__ double weiss_n2_xmole(double tt, double ts, double pn2) {
  const double atmn2 = 0.78084, atm = 1.01325e5, xmwh2o = 18.01534e-3;
  const double tk = tt + 273.15, tk_100k = tk / 100.0;
  double k0 = -172.4965 + 248.4262 / tk_100k + 143.3483 * log(tk_100k) + -21.7120 * tk_100k +
              ts * (-0.049781 + -0.025018 * tk_100k + -0.0034861 * tk_100k * tk_100k);
  k0 = exp(k0);
  k0 = (k0 * 1.e-3) / PFRX_CND_RGAS / 298.15;
  double kh = (atmn2 * atm) / k0;
  kh = kh * (1.0 / xmwh2o);
  const double cn2 = (pn2 * 1.0) / kh;
  return cn2 / (1.0 / xmwh2o);
}

// lngam_proton: ln of rt_auxvar%pri_act_coef(H+) (only read by the pH-stat)
template <class Cell>
__device__ __forceinline__ void cndegas_react(Cell &s, double lngam_proton) {
  const pfrx_cndegas &cd = s.cfg.cd;
  const double H2O_kg_mol = 18.01534e-3, rgas = 8.3144621;
  const int off = s.cfg.naq;
  const bool elm = s.cfg.elm != 0;
  const double convert_molal_to_molar = cd.initialize_with_molality ? s.den_kg * 1.0 / 1000.0 : 1.0;
  double tc = cd.reference_temperature, air_press = cd.reference_pressure, lsat = 0.50;
  if (cd.cell_state_mode >= 1) {
    air_press = fmax(air_press, s.st.pres ? s.st.pres[s.cell] : 101325.0);
    lsat = s.sat;
    if (cd.cell_state_mode >= 2) tc = s.temp;
  }
  const double porosity = s.por, volume = s.vol, air_vol = 1.0;
  const double air_molar = air_press / rgas / (tc + 273.15);
#pragma unroll 1
  for (int g = 0; g < 3; g++) {
    const int aq = g == 0 ? cd.co2a_id : g == 1 ? cd.n2oa_id : cd.n2a_id;
    const int gs = g == 0 ? cd.co2g_id : g == 1 ? cd.n2og_id : cd.n2g_id;
    if (aq < 0 || gs < 0) continue;
    const int ia = aq, ig = gs + off;
    const double c_aq = s.TOTc(aq);
    double p = (g == 0 ? 350.0e-6 : g == 1 ? 310.0e-9 : 0.78084) * cd.reference_pressure;
    if (elm) {
      const double molar = s.Cc(ig) / air_vol;
      p = molar / air_molar * air_press;
    }
    const double total_sal = 1.e-20;
    double c_eq, temp_real, kk;
    if (g == 0) {
      temp_real = fmax(fmin(tc, 40.0), -1.0);
      c_eq = weiss_co2_xmole(temp_real, air_press, total_sal, p) / H2O_kg_mol;
      temp_real = volume * 1000.0 * porosity * lsat;
      kk = cd.k_kinetic_co2;
    } else if (g == 1) {
      temp_real = fmax(fmin(tc, 40.0), 1.e-20);
      c_eq = weiss_price_n2o_xmole(temp_real, air_press, total_sal, p) / H2O_kg_mol;
      temp_real = volume * 1000.0 * porosity * lsat;
      kk = cd.k_kinetic_n2o;
    } else {
      temp_real = fmax(fmin(tc, 40.0), -2.0);
      c_eq = weiss_n2_xmole(temp_real, total_sal, p) / H2O_kg_mol;
      temp_real = volume * porosity * lsat * 1.e3;
      kk = cd.k_kinetic_n2;
    }
    const double rate = kk * (c_aq - c_eq) * temp_real;
    if (fabs(rate) > 1.0e-20) {
      s.RES(ia) = s.RES(ia) + rate;
      s.RES(ig) = s.RES(ig) - rate;
      const double drate = kk * temp_real;
      s.J(ia, ia) = s.J(ia, ia) + drate * s.DT(aq, aq);
      s.J(ig, ia) = s.J(ig, ia) - drate;
    }
  }
  if (cd.fixph_on) {
    const int ip = cd.proton_id, ih = cd.himm_id + off;
    const double c_h = s.Cc(ip) * convert_molal_to_molar;
    const double c_h_fix = pow(10.0, -1.0 * cd.fixph) / exp(lngam_proton);
    const double temp_real = volume * 1000.0 * porosity * lsat;
    const double rate = cd.k_kinetic_h * (c_h - c_h_fix) * temp_real;
    if (fabs(rate) > 1.0e-20) {
      s.RES(ip) = s.RES(ip) + rate;
      s.RES(ih) = s.RES(ih) - rate;
      const double drate = cd.k_kinetic_h * convert_molal_to_molar * temp_real;
      s.J(ip, ip) = s.J(ip, ip) + drate;
      // as written (:501): the Himm row takes the TRANSPOSED entry as its starting value
      s.J(ih, ip) = s.J(ip, ih) - drate;
    }
  }
}

}  // namespace pfrx_sbx
