/*
 * pfrx.h -- C ABI of the B200-native operator-split chemistry step.
 *
 * This is the drop-in boundary for ONE path of PFLOTRAN (as carried by
 * fmyuan/pflotran-elm-interface): the per-cell reaction step behind the OS
 * cell loop
 *     PMCSubsurfaceOSRTStepDT   src/pflotran/pmc_subsurface_osrt.F90:346-383
 *       -> RStep                src/pflotran/reaction.F90:3564
 *         -> RReact             src/pflotran/reaction.F90:3742
 * Host Fortran reaches it through ISO_C_BINDING; every
 * signature below uses only plain pointers, sizes and POD structs.
 *
 * Conventions
 *   - all reals are IEEE binary64, all indices int32, species ids are 0-BASED
 *     (the reference is 1-based; the binding subtracts 1 when it flattens
 *     reaction_rt_type);
 *   - ragged stoichiometry tables are CSR: ptr[n+1], id[nnz], stoich[nnz]
 *     (the reference stores id(0:max,n)/stoich(max,n); reaction_aux.F90:171-196);
 *   - per-cell state is "cell-major SoA": field f, component k, cell c lives at
 *     f[k*ld + c] with ld = pfrx_state.ld >= ncell.  A thread-per-cell kernel
 *     then reads consecutive doubles from consecutive lanes;
 *   - return value 0 = success, >0 = error class (PFRX_E_*); nothing aborts
 *     the process (the reference's `stop` in RStep, reaction.F90:3669, becomes
 *     a per-cell ierror).
 */
#ifndef PFRX_H
#define PFRX_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PFRX_ABI_VERSION 9

/* error classes */
#define PFRX_OK 0
#define PFRX_E_INVALID 1   /* bad argument / unsupported configuration      */
#define PFRX_E_CUDA 2      /* CUDA runtime error (see pfrx_last_error())     */
#define PFRX_E_NOTBOUND 3  /* rstep before bind_state                        */
#define PFRX_E_NCCL 4      /* NCCL missing or failed                         */
#define PFRX_E_LIMIT 5     /* problem exceeds a compiled-in size limit       */

/* reaction_aux.F90:33-38 */
#define PFRX_ACT_COEF_FREQUENCY_OFF 0
#define PFRX_ACT_COEF_FREQUENCY_TIMESTEP 1
#define PFRX_ACT_COEF_FREQUENCY_NEWTON_ITER 2
#define PFRX_ACT_COEF_ALGORITHM_LAG 3
#define PFRX_ACT_COEF_ALGORITHM_NEWTON 4

/* reaction_surf_complex_aux.F90: surface types */
#define PFRX_NULL_SURFACE 0
#define PFRX_ROCK_SURFACE 1
#define PFRX_MINERAL_SURFACE 2

/* reaction_isotherm_aux.F90: isotherm types */
#define PFRX_SORPTION_LINEAR 1
#define PFRX_SORPTION_LANGMUIR 2
#define PFRX_SORPTION_FREUNDLICH 3

/* compiled-in limits of the CUDA path (oracle has none beyond memory) */
#define PFRX_MAX_NCOMP 32
#define PFRX_MAX_PREFACTORS 10      /* reaction_mineral.F90:688 */
#define PFRX_MAX_PREFACTOR_SPECIES 5

/* species kinds used by the ELM-CN sandboxes (reaction_sandbox_somdec.F90:160-162;
 * ITYPE_GAS is not supported by the CUDA path) */
#define PFRX_SPEC_AQUEOUS 0
#define PFRX_SPEC_IMMOBILE 2

/* elm_rspfuncs.F90:17-47 */
#define PFRX_TEMPERATURE_RESPONSE_OFF 0
#define PFRX_TEMPERATURE_RESPONSE_CLMCN 1
#define PFRX_TEMPERATURE_RESPONSE_Q10 2
#define PFRX_TEMPERATURE_RESPONSE_DLEM 3
#define PFRX_TEMPERATURE_RESPONSE_ARRHENIUS 4
#define PFRX_MOISTURE_RESPONSE_OFF 0
#define PFRX_MOISTURE_RESPONSE_CLMCN 1
#define PFRX_MOISTURE_RESPONSE_DLEM 2
#define PFRX_MOISTURE_RESPONSE_LOGTHETA 3
#define PFRX_OX_RESPONSE_OFF 0
#define PFRX_OX_RESPONSE_MONOD 1
#define PFRX_OX_RESPONSE_WFPS 2
#define PFRX_INHIBITION_THRESHOLD 1
#define PFRX_INHIBITION_MONOD 3
#define PFRX_INHIBITION_INVERSE_MONOD 4

/* reaction sandboxes in the order of the deck's REACTION_SANDBOX block
 * (RSandboxEvaluate walks the list, reaction_sandbox.F90:294-330) */
#define PFRX_SANDBOX_CLM_CN 1
#define PFRX_SANDBOX_SOMDEC 2
#define PFRX_SANDBOX_NITRIF 3
#define PFRX_SANDBOX_DENITR 4
#define PFRX_SANDBOX_PLANTN 5
#define PFRX_SANDBOX_LANGMUIR 6
#define PFRX_SANDBOX_CNDEGAS 7
#define PFRX_SANDBOX_CALCITE 8
#define PFRX_SANDBOX_RADON 9
#define PFRX_MAX_SANDBOXES 10
/* reaction_microbial_aux.F90:14-22 */
#define PFRX_MICROBIAL_MOLALITY 1
#define PFRX_MICROBIAL_ACTIVITY 2
#define PFRX_MICROBIAL_MOLARITY 3
#define PFRX_INHIBITION_THRESHOLD 1
#define PFRX_INHIBITION_MONOD 3
#define PFRX_INHIBITION_INVERSE_MONOD 4
#define PFRX_INHIBITION_SMOOTHSTEP 5
#define PFRX_MAX_MONOD_TERMS 8

/*
 * SOMDECOMP sandbox: reaction_sandbox_somdec_type after SomDecSetup
 * (reaction_sandbox_somdec.F90:21-113, 987-1501), flattened.  Species ids are
 * 0-based, -1 = absent (the reference tests `> 0` on 1-based ids).  The
 * reference keeps per-reaction scratch (upstream_nc, downstream_nc,
 * mineral_c/n_stoich of variable-C:N pools) in the shared object; here the
 * N:C ratios that survive between evaluations are per-cell state
 * (pfrx_state.somdec_nc) initialised from the set-up values below.
 */
typedef struct pfrx_somdec {
  int32_t nrxn;
  int32_t co2_id, co2_itype;      /* species_id_co2 / species_itype_co2 */
  int32_t o2_id, o2_itype;        /* -1: no O2 species named */
  int32_t nh4_id, no3_id, n2o_id, proton_id;            /* primary ids */
  int32_t hr_id, nmin_id, nimm_id, nimp_id, ngasmin_id; /* immobile ids */
  double x0eps;                   /* 1e-20 */
  double n2o_frac_mineralization; /* 0.02 */
  double inhibition_nh4_no3;      /* 1.0 */
  /* per reaction [nrxn] */
  const double *rate_constant;      /* 1/s, <0 => use rate_decomposition */
  const double *rate_decomposition; /* 1/s */
  const double *rate_ad_factor;
  const int32_t *upstream_c_id;
  const int32_t *upstream_n_id;     /* >=0 => variable C:N upstream pool */
  const int32_t *upstream_is_aqueous;
  const int32_t *upstream_hr_id, *upstream_nmin_id, *upstream_nimp_id, *upstream_nimm_id;
  const double *upstream_nc;        /* mol N / mol C; -999 for variable pools */
  const double *mineral_c_stoich;   /* SomDecSetup values (fixed C:N)       */
  const double *mineral_n_stoich;
  /* downstream pools, CSR over reactions */
  const int32_t *downstream_ptr;    /* [nrxn+1] */
  const int32_t *downstream_c_id, *downstream_n_id, *downstream_is_aqueous;
  const double *downstream_stoich, *downstream_nc;
  /* abiotic factors of each reaction (abiotic_factors_type) */
  const int32_t *temperature_response_function;
  const int32_t *moisture_response_function;
  const int32_t *ox_response_function;
  const double *q10, *ea, *ox_half_saturation, *decomp_depth_efolding;
  const int32_t *ox_specid, *ox_specitype;   /* rxn%Ox_specid, -1 none */
  /* MONOD terms, CSR over reactions (monod2_type) */
  const int32_t *monod_ptr;         /* [nrxn+1] */
  const int32_t *monod_specid, *monod_specitype, *monod_pool_normalized;
  const double *monod_half_saturation, *monod_threshold;
  /* INHIBITION terms, CSR over reactions (inhibition2_type) */
  const int32_t *inhib_ptr;         /* [nrxn+1] */
  const int32_t *inhib_itype, *inhib_specid, *inhib_specitype;
  const double *inhib_constant, *inhib_constant2;
} pfrx_somdec;

/* NITRIFICATION sandbox (reaction_sandbox_nitrif.F90:20-36, 234-502) */
typedef struct pfrx_nitrif {
  int32_t proton_id, nh4_id, no3_id, n2o_id; /* primary ids, -1 absent */
  int32_t ngasnit_id;                        /* immobile id, -1 absent */
  double k_nitr_max;                         /* 1e-6 1/s  */
  double k_nitr_n2o;                         /* 3.5e-8 1/s */
  double x0eps;                              /* 1e-20 */
} pfrx_nitrif;

/* DENITRIFICATION sandbox (reaction_sandbox_denitr.F90:18-33, 212-404) */
typedef struct pfrx_denitr {
  int32_t no3_id, n2_id, n2o_id;  /* primary ids, -1 absent */
  int32_t ngasdeni_id;            /* immobile id, -1 absent */
  double half_saturation;         /* 1e-15 */
  double k_deni_max;              /* 2.5e-6 1/s */
  double x0eps;                   /* 1e-20 */
} pfrx_denitr;

/* PLANTN sandbox: plant N uptake from NH4+ / NO3- (reaction_sandbox_plantn.F90:18-56, 222-640).
 * In the ELM build the demand comes per cell from ELM (pfrx_state.elm_rate_plantndemand,
 * mol/m^3/s), otherwise it is 1e-2 * volume as in the reference's stand-alone build. */
typedef struct pfrx_plantn {
  int32_t nh4_id, no3_id;                 /* primary ids, -1 absent */
  int32_t plantn_id;                      /* immobile id of PlantN (required) */
  int32_t plantndemand_id, plantnh4uptake_id, plantno3uptake_id; /* immobile trackers, -1 absent */
  double half_saturation_nh4;             /* 1e-15 */
  double half_saturation_no3;             /* 1e-15 */
  double inhibition_nh4_no3;              /* 1.0 */
  double x0eps_nh4, x0eps_no3;            /* 1e-20 */
} pfrx_plantn;

/* LANGMUIR sandbox: kinetic Langmuir sorption of one aqueous species onto one immobile
 * species (reaction_sandbox_langmu.F90:20-58, 183-330) */
typedef struct pfrx_langmuir {
  int32_t aq_id;                          /* primary id */
  int32_t sorb_id;                        /* immobile id */
  double k_kinetic;                       /* 1e-5 1/s */
  double k_equilibrium;                   /* 2.5e3 */
  double s_max;                           /* 1e-3 mol/m^3 */
} pfrx_langmuir;

/* CNDEGAS sandbox: first-order exchange of dissolved CO2 / N2O / N2 with a gas reservoir at the
 * Weiss (1974) / Weiss & Price (1980) solubilities, and an optional pH-stat
 * (reaction_sandbox_cndegas.F90:14-29, 216-546; solubilities :548-787).
 * A pair is active when both ids are >= 0.  *_g_id is the index the reference adds to
 * reaction%offset_immobile for the residual row of the gas reservoir (a primary-species id of
 * 'CO2(g)*' or a gas id of 'CO2(g)', CNdegasSetup :160-190 -- used as an immobile index either way). */
typedef struct pfrx_cndegas {
  int32_t co2a_id, n2oa_id, n2a_id;       /* primary ids of CO2(aq), N2O(aq), N2(aq); -1 absent */
  int32_t co2g_id, n2og_id, n2g_id;       /* reservoir rows: naqcomp + id; -1 absent */
  int32_t proton_id, himm_id;             /* pH-stat: H+ (primary), Himm (immobile) */
  int32_t fixph_on;                       /* FIXPH given */
  int32_t initialize_with_molality;       /* reaction%initialize_with_molality: c_h in mol/L */
  /* 0: stand-alone transport (tc, air pressure = the reference values, liquid saturation 0.5);
   * 1: RICHARDS flow (pressure and saturation of the cell); 2: TH flow or an ELM_PFLOTRAN build
   * (also the cell's temperature).  In an ELM_PFLOTRAN build (pfrx_config.elm_pflotran) the partial
   * pressures come from the reservoir concentrations instead of the atmospheric defaults. */
  int32_t cell_state_mode;
  int32_t pad_;
  double k_kinetic_co2, k_kinetic_n2o, k_kinetic_n2, k_kinetic_h;   /* 1e-5 1/s each */
  double fixph;                           /* 6.5 */
  double reference_temperature;           /* option%flow%reference_temperature, C */
  double reference_pressure;              /* option%flow%reference_pressure, Pa */
} pfrx_cndegas;

/* CALCITE sandbox (reaction_sandbox_calcite.F90:177-365): two parallel TST pathways for one kinetic
 * mineral whose RATE_CONSTANT in MINERAL_KINETICS is zero -- the first takes stoichiometry and logK of
 * the mineral from the kinmnrl tables, the second is written out for Calcite = Ca++ + HCO3- - H+ with
 * pKeq 1.8487.  The sum of the two rates [mol/m^3 bulk/s] is kept per cell
 * (rt_auxvar%auxiliary_data -> pfrx_state.sandbox_aux) and moves the mineral's volume fraction when
 * the step is accepted (CalciteUpdateKineticState :369-410). */
typedef struct pfrx_calcite_sandbox {
  int32_t mineral_id;                     /* kinetic-mineral index */
  int32_t h_ion_id, calcium_id, bicarbonate_id; /* primary ids */
  double rate_constant1, rate_constant2;  /* mol/m^2/s */
} pfrx_calcite_sandbox;

/* RADON sandbox (reaction_sandbox_radon.F90:150-188): zero-order generation of a species in proportion
 * to the volume fraction of a mineral */
typedef struct pfrx_radon {
  int32_t species_id;                     /* primary id (Rn(aq)) */
  int32_t mineral_id;                     /* kinetic-mineral index */
  double radon_generation_rate;           /* mol/m^3 mineral/s */
} pfrx_radon;

/*
 * Flattened, read-only reaction description: the subset of
 *   reaction_rt_type            reaction_aux.F90:123-311
 *   mineral_type                reaction_mineral_aux.F90:82-138
 *   surface_complexation_type   reaction_surf_complex_aux.F90:68-128
 *   reaction_sandbox_clm_cn_type reaction_sandbox_clm_cn.F90:22-42
 *   reaction_sandbox_somdec / nitrif / denitr types (pfrx_somdec etc. above)
 * that RStep and its callees read.  pfrx_create() copies everything; the
 * caller keeps ownership of the arrays.
 */
typedef struct pfrx_config {
  int32_t abi_version;           /* = PFRX_ABI_VERSION */

  /* ---- sizes: ncomp = naqcomp + nimcomp, offset_immobile = naqcomp -------- */
  int32_t naqcomp;
  int32_t nimcomp;               /* reaction%immobile%nimmobile */

  /* ---- flags / tolerances (defaults: reaction_aux.F90:371-556) ------------ */
  int32_t use_full_geochemistry; /* 0 => RStep early-outs, reaction.F90:3604  */
  int32_t use_log_formulation;
  int32_t use_total_as_guess;
  int32_t use_isothermal;        /* 0 => logK(T) per cell, reaction.F90:5976  */
  int32_t act_coef_update_frequency;
  int32_t act_coef_update_algorithm;
  int32_t use_activity_h2o;
  int32_t h2o_aq_id;             /* species_idx%h2o_aq_id, -1 if none         */
  int32_t maximum_reaction_iterations;   /* 20 */
  int32_t maximum_reaction_cuts;         /* 10 */
  double max_dlnC_rreact;                /* 5   */
  double max_relative_change_tolerance;  /* 1e-6 */
  double max_residual_tolerance;         /* 1e-12 */
  double max_rel_residual_tolerance;     /* 1e-8 */
  double rt_min_saturation;              /* 1e-40, reactive_transport_aux.F90:19 */
  double debyeA, debyeB, debyeBdot;      /* reaction_database.F90:931-1023 */

  /* ---- primary aqueous species ------------------------------------------- */
  const double *primary_spec_Z;          /* [naqcomp] */
  const double *primary_spec_a0;         /* [naqcomp] */

  /* ---- secondary aqueous complexes (RTotalAqueous, reaction.F90:4665) ---- */
  int32_t neqcplx;
  const int32_t *eqcplx_ptr;             /* [neqcplx+1] */
  const int32_t *eqcplx_specid;          /* [nnz] primary ids */
  const double *eqcplx_stoich;           /* [nnz] */
  const double *eqcplx_h2ostoich;        /* [neqcplx]; 0 when eqcplxh2oid==0 */
  const double *eqcplx_logK;             /* [neqcplx] at the reference T     */
  const double *eqcplx_logKcoef;         /* [5*neqcplx] or NULL (isothermal) */
  const double *eqcplx_Z;                /* [neqcplx] */
  const double *eqcplx_a0;               /* [neqcplx] */

  /* ---- kinetic minerals (RKineticMineral, reaction_mineral.F90:647) ------ */
  int32_t nkinmnrl;
  const int32_t *kinmnrl_ptr;            /* [nkinmnrl+1] */
  const int32_t *kinmnrl_specid;         /* [nnz] primary ids */
  const double *kinmnrl_stoich;          /* [nnz] */
  const double *kinmnrl_h2ostoich;       /* [nkinmnrl] */
  const double *kinmnrl_logK;            /* [nkinmnrl] */
  const double *kinmnrl_logKcoef;        /* [5*nkinmnrl] or NULL */
  const double *kinmnrl_molar_vol;       /* [nkinmnrl] m^3/mol */
  const double *kinmnrl_rate_constant;   /* [nkinmnrl] mol/m^2/s */
  const double *kinmnrl_activation_energy; /* [nkinmnrl] J/mol, 0 = none */
  const double *kinmnrl_affinity_threshold; /* [nkinmnrl] */
  const double *kinmnrl_rate_limiter;    /* [nkinmnrl] */
  const int32_t *kinmnrl_irreversible;   /* [nkinmnrl] */
  /* optional arrays: NULL reproduces the reference's `.not.associated(...)` */
  const double *kinmnrl_Temkin_const;    /* [nkinmnrl] or NULL */
  const double *kinmnrl_min_scale_factor;/* [nkinmnrl] or NULL */
  const double *kinmnrl_affinity_power;  /* [nkinmnrl] or NULL */
  /* prefactors (reaction_mineral.F90:838-890); all NULL when unused.
   * Dense like the reference: pref index p < PFRX_MAX_PREFACTORS, species
   * slot s < PFRX_MAX_PREFACTOR_SPECIES, layout [(m*MAXP + p)*MAXS + s].     */
  const int32_t *kinmnrl_num_prefactors; /* [nkinmnrl] or NULL */
  const int32_t *kinmnrl_pref_nspec;     /* [nkinmnrl*MAXP] */
  const int32_t *kinmnrl_prefactor_id;   /* >=0 primary id; <0 => -(icplx+1) */
  const double *kinmnrl_pref_alpha;
  const double *kinmnrl_pref_beta;
  const double *kinmnrl_pref_atten_coef;
  const double *kinmnrl_pref_rate;       /* [nkinmnrl*MAXP] */
  const double *kinmnrl_pref_activation_energy; /* [nkinmnrl*MAXP] */

  /* ---- surface complexation (reaction_surf_complex.F90:446-900) ---------- */
  int32_t nsrfcplxrxn;
  int32_t nsrfcplx;
  const int32_t *srfcplxrxn_ptr;         /* [nsrfcplxrxn+1] -> complex ids   */
  const int32_t *srfcplxrxn_to_complex;  /* [.] */
  const int32_t *srfcplxrxn_surf_type;   /* [nsrfcplxrxn] PFRX_*_SURFACE     */
  const int32_t *srfcplxrxn_to_surf;     /* [nsrfcplxrxn] kinetic-mineral id */
  const double *srfcplxrxn_site_density; /* [nsrfcplxrxn] */
  const int32_t *srfcplxrxn_stoich_flag; /* [nsrfcplxrxn] 1 => inner Newton  */
  const int32_t *srfcplx_ptr;            /* [nsrfcplx+1] */
  const int32_t *srfcplx_specid;
  const double *srfcplx_stoich;
  const double *srfcplx_h2ostoich;       /* [nsrfcplx] */
  const double *srfcplx_free_site_stoich;/* [nsrfcplx] */
  const double *srfcplx_logK;            /* [nsrfcplx] */
  const double *srfcplx_logKcoef;        /* [5*nsrfcplx] or NULL */
  int32_t neqsrfcplxrxn;
  const int32_t *eqsrfcplxrxn_to_srfcplxrxn;   /* [neqsrfcplxrxn] */
  int32_t nkinmrsrfcplxrxn;
  const int32_t *kinmrsrfcplxrxn_to_srfcplxrxn;/* [nkinmrsrfcplxrxn] */
  const int32_t *kinmr_rate_ptr;         /* [nkinmrsrfcplxrxn+1] */
  const double *kinmr_rate;              /* [.] 1/s */
  const double *kinmr_frac;              /* [.] */

  /* ---- ion exchange (RTotalSorbEqIonx, reaction.F90:4906-5140) ------------ */
  int32_t neqionxrxn;
  const int32_t *eqionx_ptr;             /* [neqionxrxn+1] -> cations; the first
                                            cation of a reaction is its REFERENCE */
  const int32_t *eqionx_cationid;        /* [.] primary ids */
  const double *eqionx_k;                /* [.] selectivity coefficients (1 for the reference) */
  const double *eqionx_CEC;              /* [neqionxrxn] eq/m^3 bulk (or per mineral volume) */
  const int32_t *eqionx_to_surf;         /* [neqionxrxn] kinetic-mineral id, -1: CEC is absolute */
  const int32_t *eqionx_Z_flag;          /* [neqionxrxn] 1: valences differ => inner Newton */
  /* ---- KD isotherms (RTotalSorbKD, reaction_isotherm.F90:273-359) ---------- */
  int32_t neqkdrxn;
  int32_t ikd_units;                     /* 0 kg water/m^3 bulk, 1 mL water/g soil */
  const int32_t *eqkd_specid;            /* [neqkdrxn] primary ids */
  const int32_t *eqkd_type;              /* [neqkdrxn] PFRX_SORPTION_* */
  const int32_t *eqkd_mineral;           /* [neqkdrxn] kinetic-mineral id scaling the KD, -1 none */
  const double *eqkd_coeff;              /* [neqkdrxn] KD */
  const double *eqkd_langmuir_b;         /* [neqkdrxn] */
  const double *eqkd_freundlich_n;       /* [neqkdrxn] */
  /* ---- dynamic KD (RTotalSorbDynamicKD, reaction.F90:4836-4902) ------------ */
  int32_t neqdynamickdrxn;
  const int32_t *eqdynamickd_specid;     /* [.] sorbing species */
  const int32_t *eqdynamickd_refspecid;  /* [.] species the KD depends on */
  const double *eqdynamickd_refspechigh; /* [.] */
  const double *eqdynamickd_low, *eqdynamickd_high, *eqdynamickd_power;

  /* ---- CLM-CN reaction sandbox (reaction_sandbox_clm_cn.F90:468-787) ----- */
  int32_t clmcn_nrxn;                    /* 0 => sandbox absent */
  int32_t clmcn_npool;
  int32_t clmcn_C_species_id;            /* immobile ids, 0-based */
  int32_t clmcn_N_species_id;
  const double *clmcn_CN_ratio;          /* [npool] mol C/mol N; <0 => litter */
  const int32_t *clmcn_pool_nspec;       /* [npool] 1 (SOM) or 2 (litter C,N) */
  const int32_t *clmcn_pool_C_id;        /* [npool] immobile id of C or SOM   */
  const int32_t *clmcn_pool_N_id;        /* [npool] immobile id of N, -1 SOM  */
  const int32_t *clmcn_upstream_pool_id; /* [nrxn] */
  const int32_t *clmcn_downstream_pool_id; /* [nrxn], -1 => none */
  const double *clmcn_rate_constant;     /* [nrxn] 1/s */
  const double *clmcn_respiration_fraction; /* [nrxn] */
  const double *clmcn_inhibition_constant;  /* [nrxn] */

  /* ---- ELM-CN sandboxes (NULL => absent) ---------------------------------- */
  const pfrx_somdec *somdec;     /* reaction_sandbox_somdec.F90:1504-3640 */
  const pfrx_nitrif *nitrif;     /* reaction_sandbox_nitrif.F90:234-502   */
  const pfrx_denitr *denitr;     /* reaction_sandbox_denitr.F90:212-404   */
  const pfrx_plantn *plantn;     /* reaction_sandbox_plantn.F90:222-640   */
  const pfrx_langmuir *langmuir; /* reaction_sandbox_langmu.F90:183-330   */
  const pfrx_cndegas *cndegas;   /* reaction_sandbox_cndegas.F90:216-546  */
  const pfrx_calcite_sandbox *calcite; /* reaction_sandbox_calcite.F90:177-410 */
  const pfrx_radon *radon;       /* reaction_sandbox_radon.F90:150-188    */
  /* evaluation order of the sandboxes (PFRX_SANDBOX_*); NULL => the order
   * CLM-CN, SOMDEC, NITRIF, DENITR, PLANTN, LANGMUIR, CNDEGAS, CALCITE, RADON */
  int32_t nsandbox;
  const int32_t *sandbox_list;
  /* 1 => the behaviour of a reference built with -DELM_PFLOTRAN in BGC-only
   * coupling (option%nflowspec == 0): moisture / oxygen / temperature scalars,
   * soil depth, decomposition scalar, dry bulk density and Clapp-Hornberger b
   * come per cell from ELM (pfrx_state.elm_*) instead of the response
   * functions / constants of the stand-alone build */
  int32_t elm_pflotran;

  /* ---- other kinetic terms of RReaction (ABI v5; counts 0 => absent) -------- */
  /* general kinetic reactions, RGeneral (reaction.F90:5316-5460), tables as built at
   * reaction_database.F90:3060-3145: every species of reaction k with its signed
   * stoichiometry (reactants negative), the reactants again with |stoich| for the
   * forward rate law, the products for the backward one.  Aqueous species only.   */
  int32_t ngeneral_rxn;
  const int32_t *general_ptr;        /* [n+1] CSR into general_specid / general_stoich */
  const int32_t *general_specid;
  const double *general_stoich;
  const int32_t *general_fwd_ptr;    /* [n+1] */
  const int32_t *general_fwd_specid;
  const double *general_fwd_stoich;  /* > 0 */
  const int32_t *general_bwd_ptr;    /* [n+1] */
  const int32_t *general_bwd_specid;
  const double *general_bwd_stoich;
  const double *general_kf;          /* [n] kg^(m-1)/mol^(m-1)-sec; <= 0 => no forward term */
  const double *general_kr;          /* [n] */
  /* radioactive decay of one parent (aqueous + sorbed inventory) into any number
   * of daughters, RRadioactiveDecay (reaction.F90:5211-5311; tables :2940-3017) */
  int32_t nradiodecay_rxn;
  const int32_t *radiodecay_ptr;     /* [n+1] CSR into radiodecay_specid / _stoich */
  const int32_t *radiodecay_specid;
  const double *radiodecay_stoich;   /* parent negative */
  const int32_t *radiodecay_forward_specid; /* [n] the parent */
  const double *radiodecay_kf;       /* [n] 1/s */
  /* first-order decay of immobile species, RImmobileDecay (reaction_immobile.F90:244-296) */
  int32_t nimmobile_decay_rxn;
  const int32_t *immobile_decay_specid;   /* [n] immobile index */
  const double *immobile_decay_constant;  /* [n] 1/s */
  /* Monod-type microbial reactions, RMicrobial (reaction_microbial.F90:287-602; tables
   * reaction_database.F90:3150-3400).  Monod and inhibition terms are CSR lists per
   * reaction (at most PFRX_MAX_MONOD_TERMS each).                                    */
  int32_t nmicrobial_rxn;
  int32_t microbial_concentration_units;     /* PFRX_MICROBIAL_* */
  const int32_t *microbial_ptr;              /* [n+1] CSR into microbial_specid / _stoich */
  const int32_t *microbial_specid;           /* aqueous species */
  const double *microbial_stoich;            /* reactants negative */
  const double *microbial_rate_constant;     /* [n] */
  const double *microbial_activation_energy; /* [n] J/mol, or NULL when no reaction has one */
  const int32_t *microbial_monod_ptr;        /* [n+1] */
  const int32_t *microbial_monod_specid;
  const double *microbial_monod_K;
  const double *microbial_monod_Cth;
  const int32_t *microbial_inhibition_ptr;   /* [n+1] */
  const int32_t *microbial_inhibition_specid;
  const int32_t *microbial_inhibition_type;  /* PFRX_INHIBITION_* */
  const double *microbial_inhibition_C;
  const double *microbial_inhibition_C2;
  const int32_t *microbial_biomassid;        /* [n] 0 none, k+1 aqueous species k, -(k+1) immobile species k */
  const double *microbial_biomass_yield;     /* [n] */
  /* active gas species, RTotalGas (reaction_gas.F90:87-174): a gas phase that holds the components in
   * equilibrium with the water -- partial pressure [bar] = exp(-logK ln10 + h2o ln a_w + sum nu ln a_i),
   * ideal-gas concentration added to rt_auxvar%total(:,2); its accumulation (reaction.F90:5761-5769,
   * 5838-5846) and its share of a decaying inventory (:5243-5296) follow.  Thread-per-cell kernel only. */
  int32_t nactive_gas;
  int32_t pad_gas_;
  const int32_t *acteq_ptr;          /* [nactive_gas+1] CSR into acteq_specid / _stoich          */
  const int32_t *acteq_specid;       /* primary ids                                             */
  const double *acteq_stoich;
  const double *acteq_h2ostoich;     /* [nactive_gas]                                           */
  const double *acteq_logK;          /* [nactive_gas] at the reference temperature              */
  const double *acteq_logK_coef;     /* [nactive_gas][5] or NULL                                */
  /* ELM build WITH a flow mode (option%nflowspec > 0, reaction_sandbox_somdec.F90:1640-1644): SOMDECOMP
   * reactions whose MOISTURE_RESPONSE_FUNCTION is CLMCN or DLEM take f_w from GetMoistureResponse
   * (elm_rspfuncs.F90:124-237: Clapp-Hornberger matric potential from sucsat / bsw / dry bulk density, or
   * the DLEM curve between field capacity and effective porosity) instead of ELM's w_scalar.  0: BGC-only
   * coupling, the function is ignored as in the reference.  Only read when elm_pflotran != 0. */
  int32_t elm_flow_coupled;
  int32_t pad_elm_;
} pfrx_config;

/*
 * Per-cell state views (cell-major SoA).  Device pointers for
 * pfrx_bind_state(), host pointers for pfrx_rstep_host().  A pointer may be
 * NULL when its count is zero.  "in" = read, "io" = read and updated in place.
 *   reactive_transport_auxvar_type  reactive_transport_aux.F90:21-74
 *   global_auxvar_type              global_aux.F90:11-35
 *   material_auxvar_type            material_aux.F90:53-77
 */
typedef struct pfrx_state {
  int64_t ld;                  /* leading dimension (>= ncell) of every field */
  /* rt_auxvar */
  double *total;               /* io [naqcomp]  mol/L; in: transported total* */
  double *pri_molal;           /* io [naqcomp]  mol/kg; in: Newton guess      */
  double *immobile;            /* io [nimcomp]  mol/m^3                       */
  double *pri_act_coef;        /* io [naqcomp]                                */
  double *sec_act_coef;        /* io [neqcplx]                                */
  double *sec_molal;           /* io [neqcplx]                                */
  double *ln_act_h2o;          /* io [1]                                      */
  double *mnrl_volfrac;        /* io [nkinmnrl]                               */
  double *mnrl_area;           /* in [nkinmnrl] m^2/m^3                       */
  double *mnrl_rate;           /* io [nkinmnrl] mol/m^3/s                     */
  double *srfcplxrxn_free_site_conc; /* io [nsrfcplxrxn]                      */
  double *eqsrfcplx_conc;      /* io [nsrfcplx]                               */
  double *total_sorb_eq;       /* io [naqcomp] when any equilibrium sorption reaction
                                  (surface complexation, ion exchange, KD) exists  */
  /* (the two ion-exchange fields are declared after the ELM scalars below) */
  double *kinmr_total_sorb;    /* io [sum_r naqcomp*(nrate_r+1)]: rxn r, rate
                                  slot q (0 = equilibrium target), comp i at
                                  row  naqcomp*(kinmr_rate_ptr[r]+r+q) + i    */
  /* global_auxvar / material_auxvar */
  const double *den_kg;        /* in [1] kg/m^3 */
  const double *sat;           /* in [1] liquid saturation */
  const double *temp;          /* in [1] deg C */
  const double *porosity;      /* in [1] */
  const double *volume;        /* in [1] m^3 */
  const double *soil_particle_density; /* in [1] or NULL */
  const int32_t *imat;         /* in [1] or NULL; <=0 => inactive, skipped
                                  (pmc_subsurface_osrt.F90:351)               */
  /* ELM per-cell scalars (elm_pflotran_interface_data: w_scalar_pfs,
   * o_scalar_pfs, t_scalar_pfs, zsoil_pfs, kscalar_decomp_c_pfs,
   * bulkdensity_dry_pfs, bsw_pfs), read through option%iflag in the reference
   * (reaction_sandbox_somdec.F90:1593,1647-1735).  in [1] each; needed only
   * when pfrx_config.elm_pflotran != 0, else NULL */
  const double *elm_w_scalar;
  const double *elm_o_scalar;
  const double *elm_t_scalar;
  const double *elm_zsoil;
  const double *elm_kscalar_decomp_c;
  const double *elm_bulkdensity_dry;
  const double *elm_bsw;
  const double *elm_rate_plantndemand;   /* rate_plantndemand_pfs, mol/m^3/s (PLANTN) */
  /* SOMDECOMP: last N:C ratios of the variable-C:N pools, io
   * [somdec.nrxn + somdec.downstream_ptr[nrxn]]: upstream_nc(irxn) then
   * downstream_nc(j).  The reference keeps them in the sandbox object and only
   * refreshes a ratio while both pool concentrations are >= x0eps
   * (reaction_sandbox_somdec.F90:1795-1822), so a pool that has decayed below
   * x0eps goes on with its last ratio; here that memory is per cell.  Initial
   * value: pfrx_somdec.upstream_nc / downstream_nc.  NULL => every evaluation
   * starts from those set-up values. */
  double *somdec_nc;
  /* ion exchange: sorbed concentration of every reaction's reference cation, the starting
   * point of its inner Newton (rt_auxvar%eqionx_ref_cation_sorbed_conc, initial value 1e-9,
   * reactive_transport_aux.F90:281-285), io [neqionxrxn]; and the sorbed concentration of
   * every cation (rt_auxvar%eqionx_conc), io [eqionx_ptr[neqionxrxn]], may be NULL */
  double *eqionx_ref_cation_sorbed_conc;
  double *eqionx_conc;
  /* in, optional: liquid pressure [Pa] (global_auxvar%pres(1)); read by the CNDEGAS sandbox when
   * its cell_state_mode is 1 or 2 */
  const double *pres;
  /* io [1] when the CALCITE sandbox is configured: rt_auxvar%auxiliary_data, the sandbox's rate of the
   * latest evaluation [mol/m^3 bulk/s] (read by its kinetic-state update) */
  double *sandbox_aux;
  /* active gas phase (pfrx_config.nactive_gas > 0, else NULL): gas saturation global_auxvar%sat(2), in [1];
   * rt_auxvar%total(:,2) [mol/L gas], io [naqcomp] -- the fixed accumulation of a step reads the value the
   * latest RTotal left, as in the reference ("still need code to overwrite other phases",
   * reaction.F90:3832); rt_auxvar%gas_pp [bar], out [nactive_gas] */
  const double *sat_gas;
  double *total_gas;
  double *gas_pp;
  /* ELM soil hydraulic properties for GetMoistureResponse (elm_pf_idata%sucsat_pfs [mm H2O], watfc_pfs,
   * effporosity_pfs), in [1] each; needed only when pfrx_config.elm_flow_coupled != 0, else NULL */
  const double *elm_sucsat;
  const double *elm_watfc;
  const double *elm_effporosity;
  /* per-cell results of RStep (reaction.F90:3564-3566) */
  int32_t *num_sub_steps;
  int32_t *num_iterations;
  int32_t *num_kinetic_state_updates;
  int32_t *ierror;
} pfrx_state;

/* Shard-level result: what the OS coupler accumulates after the cell loop
 * (pmc_subsurface_osrt.F90:364-388) plus the flags the north star asks for. */
typedef struct pfrx_step_result {
  int64_t ncell_active;        /* cells with imat > 0                        */
  int64_t sum_newton_iterations;
  int32_t max_newton_iterations;
  int32_t max_num_kinetic_state_updates;
  int32_t rstep_error;         /* max over cells of ierror                   */
  int32_t max_sub_steps;
  int64_t num_cut_cells;       /* cells that needed >= 1 reaction-dt cut     */
  int64_t first_failed_cell;   /* lowest failing local cell index, or -1     */
} pfrx_step_result;

typedef struct pfrx_handle pfrx_handle;

/* library identification; safe to call without a GPU */
int pfrx_abi_version(void);
const char *pfrx_last_error(void);
/* sizeof(pfrx_config | pfrx_state | pfrx_step_result) for which = 0 | 1 | 2:
 * lets a foreign-language binding assert its struct layout */
int64_t pfrx_sizeof(int which);

/* Replaces nothing in the reference: flattening of reaction_rt_type is done by
 * the binding.  `device` is the CUDA ordinal this handle is tied to. */
int pfrx_create(const pfrx_config *cfg, int device, pfrx_handle **out);
void pfrx_destroy(pfrx_handle *h);

/* Bind device-resident SoA state of `ncell` cells (zero-copy: the kernels
 * update these arrays in place).  Replaces the rt_auxvars(ghosted_id) /
 * global_auxvars / material_auxvars lookups at pmc_subsurface_osrt.F90:349-363. */
int pfrx_bind_state(pfrx_handle *h, int64_t ncell, const pfrx_state *dev);

/* The cell loop pmc_subsurface_osrt.F90:349-378: RStep on every bound cell over
 * tran_dt.  Asynchronous variant enqueues on the handle's stream;
 * pfrx_rstep_finish() synchronises and returns the shard summary.           */
int pfrx_rstep_async(pfrx_handle *h, double tran_dt);
int pfrx_rstep_finish(pfrx_handle *h, pfrx_step_result *out);
int pfrx_rstep(pfrx_handle *h, double tran_dt, pfrx_step_result *out);

/* Same step for HOST-resident SoA state (what a Fortran caller has after its
 * AoS->SoA pack): H2D of every `in`/`io` field, kernel, D2H of every `io`
 * field and the per-cell results, through pinned staging owned by the handle. */
int pfrx_rstep_host(pfrx_handle *h, int64_t ncell, const pfrx_state *host,
                    double tran_dt, pfrx_step_result *out);

/* Fields that pfrx_rstep_host keeps RESIDENT in its device mirror (bit f of
 * field_mask = the f-th `double *` member of pfrx_state in declaration order,
 * PFRX_FIELD_*): uploaded when the mirror is (re)allocated, not downloaded.
 * For what the step derives and only the next step reads: rt_auxvar%sec_molal
 * and the activity coefficients are 1.4 of the 1.97 KB per cell a Hanford step
 * sends back (reactive_transport_aux.F90:21-74 keeps them per cell for output
 * only).  pfrx_rstep_host_fetch copies the resident fields to the host arrays
 * on request (output / checkpoint times).  Default mask: 0 (everything moves). */
#define PFRX_FIELD_TOTAL 0
#define PFRX_FIELD_PRI_MOLAL 1
#define PFRX_FIELD_IMMOBILE 2
#define PFRX_FIELD_PRI_ACT_COEF 3
#define PFRX_FIELD_SEC_ACT_COEF 4
#define PFRX_FIELD_SEC_MOLAL 5
#define PFRX_FIELD_LN_ACT_H2O 6
#define PFRX_FIELD_MNRL_VOLFRAC 7
#define PFRX_FIELD_MNRL_AREA 8
#define PFRX_FIELD_MNRL_RATE 9
#define PFRX_FIELD_FREE_SITE 10
#define PFRX_FIELD_EQSRFCPLX_CONC 11
#define PFRX_FIELD_TOTAL_SORB_EQ 12
#define PFRX_FIELD_KINMR_TOTAL_SORB 13
int pfrx_rstep_host_resident(pfrx_handle *h, uint64_t field_mask);
int pfrx_rstep_host_fetch(pfrx_handle *h, int64_t ncell, const pfrx_state *host);

/* Multi-GPU: one rank per GPU, cells sharded by ownership range, no halo.
 * pfrx_allreduce() replaces MPI_Allreduce(rstep_error,MAX)+MPI_Barrier at
 * pmc_subsurface_osrt.F90:381-383 (MAX of error / iteration / update counts,
 * SUM of iterations and cell counts) with one NCCL call over NVLink.  Once a
 * communicator is attached, pfrx_rstep_async enqueues that reduction on the
 * kernel stream straight from the device summary (no host staging), and
 * pfrx_allreduce of the result pfrx_rstep_finish returned only reads it back.
 * The 128-byte id comes from rank 0 (pfrx_comm_unique_id) and is distributed
 * by the host (MPI_Bcast in the Fortran caller).                            */
int pfrx_comm_unique_id(void *id128);
int pfrx_comm_init(pfrx_handle *h, int nranks, int rank, const void *id128);
int pfrx_allreduce(pfrx_handle *h, pfrx_step_result *inout);

/* stream the handle launches on (cudaStream_t), for callers that time with
 * CUDA events or chain their own work behind the step */
void *pfrx_stream(pfrx_handle *h);
/* number of kernel launches issued by this handle so far */
int64_t pfrx_launch_count(pfrx_handle *h);
/* Refill kernels (the generated variants for ragged workloads) hand out cells longest-first: after a launch over
 * the whole shard the library sorts the cells by the Newton iterations they just needed, and the next launch on
 * that shard starts with the slowest ones -- chemistry hot spots persist from one transport step to the next,
 * so the cells that cut their step dozens of times no longer trail the launch.  Results do not depend on the
 * order (a cell is computed from its own state).  mode 0 switches it off, 1 (default; PFRX_CELL_ORDER) on. */
int pfrx_cell_order(pfrx_handle *h, int mode);
/* bytes of state read+written per cell-solve by the bound configuration
 * (the algorithmic HBM traffic of SURVEY.md section 8(d))                    */
int64_t pfrx_bytes_per_cell(pfrx_handle *h);

/* bytes the latest pfrx_rstep_host moved over the host link in each direction.
 * Fields that every active cell overwrites before reading (activity
 * coefficients when they are updated per Newton iteration, mineral rates) are
 * not uploaded when the shard has no inactive cells (imat NULL or all > 0).   */
int pfrx_last_transfer_bytes(pfrx_handle *h, int64_t *h2d, int64_t *d2h);

/* kernel configuration chosen for this handle:
 * info5 = {padded system size N, lanes per cell, threads per block,
 *          resident blocks per SM, dynamic shared memory bytes per block}.
 * PFRX_LANES / PFRX_THREADS in the environment override the defaults. */
int pfrx_kernel_info(pfrx_handle *h, int *info5);

/* ---- batched RReaction / RReactionDerivative --------------------------------
 * What the fully implicit (GIRT) and ELM callers take from the chemistry
 * module (reactive_transport.F90:2627 RReaction, :3288 RReactionDerivative;
 * reaction.F90:4059-4208): the kinetic terms -- mineral precipitation /
 * dissolution and the reaction sandboxes -- of every active cell of the bound
 * state, evaluated from rt_auxvar as it stands (pri_molal, pri_act_coef,
 * immobile, mnrl_volfrac/area; no RTAuxVarCompute).  Device pointers, ld of the
 * bound state:  res[i*ld + cell]  (mol/s, the sign convention of Residual),
 * jac[(i*ncomp + j)*ld + cell] = d res_i / d c_j  (written when want_jacobian).
 * Side effect as in the reference: mnrl_rate of the bound state is updated.
 * Inactive cells (imat <= 0) get zeros; dry cells get zeros (reaction.F90:4085).
 * tran_dt is option%tran_dt, which the SOMDECOMP sandbox reads (rate caps,
 * reaction_sandbox_somdec.F90:1762,2853) and RMultiRateSorption uses; the
 * aqueous totals and d(total)/d(free) that the sandboxes and radioactive decay
 * read are recomputed from the free-ion concentrations of the state, and so
 * are the sorbed totals when a sorbing species decays.                         */
int pfrx_reaction(pfrx_handle *h, double tran_dt, int want_jacobian, double *res, double *jac);

/* ---- checkpoint layout of the kinetically sorbed concentrations ---------------
 * RTCheckpointKineticSorptionBinary / HDF5 (reactive_transport.F90:3968-4182) write the explicitly
 * stored sorbed concentrations as a sequence of per-cell vectors: components that take part in a
 * multirate reaction outermost, then the multirate reactions, then their rates
 * (kinmr_total_sorb(icomp, irate, irxn), irate >= 1; the equilibrium target of slot 0 is not
 * checkpointed).  Every such vector is one row of pfrx_state.kinmr_total_sorb -- cell-major, i.e.
 * already the array VecView writes -- and rows[k] is the row of the k-th vector of the checkpoint;
 * *nrows receives their number (rows may be NULL to ask for it).  Host-only, no device needed. */
int pfrx_kinmr_checkpoint_rows(const pfrx_config *cfg, int32_t *rows, int32_t *nrows);

/* ---- RTUpdateAuxVars over the bound state ------------------------------------
 * The refresh the reference runs after a restart, after the initial condition and after every
 * accepted GIRT step (reactive_transport.F90:3525-3660, cells branch): for every active cell,
 * pri_molal / immobile <- tran_xx (device block vector [ncell][ncomp], or NULL: keep the state's),
 * RActivityCoefficients when update_activity_coefs != 0 (and the update frequency is not OFF),
 * then RTAuxVarCompute: total, sec_molal, and the equilibrium-sorbed state (total_sorb_eq,
 * free-site / surface-complex concentrations, ion-exchange state). */
int pfrx_update_auxvars(pfrx_handle *h, const double *tran_xx, int update_activity_coefs);

/* ---- batched ReactionEquilibrateConstraint -----------------------------------
 * The set-up step that turns a CONSTRAINT block into a speciated rt_auxvar
 * (reaction.F90:1328-2117, called per condition by
 * condition_control.F90 / transport_constraint_rt.F90): Newton on the free-ion
 * molalities with one equation per primary species chosen by its constraint
 * type; activity coefficients are switched on after a first convergence and
 * the loop ends at the second.  Here every cell of the bound state is
 * equilibrated at once with its own constraint values (an initial condition
 * that varies from cell to cell) and the same constraint types.
 * Constraint types carry the reference's values (transport_constraint_rt.F90:22-34). */
#define PFRX_CONSTRAINT_NULL 0       /* treated as TOTAL (reaction.F90:1619)                  */
#define PFRX_CONSTRAINT_FREE 1       /* free-ion concentration                                */
#define PFRX_CONSTRAINT_TOTAL 2      /* total aqueous component concentration                 */
#define PFRX_CONSTRAINT_LOG 3        /* log10 of the free-ion concentration                   */
#define PFRX_CONSTRAINT_PH 4         /* pH, on the primary species H+                         */
#define PFRX_CONSTRAINT_MINERAL 7    /* equilibrium with a mineral (tables below)             */
#define PFRX_CONSTRAINT_GAS 8        /* equilibrium with a gas at a partial pressure [bar];
                                        a value <= 0 is log10 of the pressure                 */
#define PFRX_CONSTRAINT_CHARGE_BAL 9 /* charge balance                                        */
/* not built: PE (5), EH (6), TOTAL_SORB (10), SUPERCRIT_CO2 (11), TOTAL_AQ_PLUS_SORB (12):
 * pfrx_equilibrate_constraint returns PFRX_E_INVALID                                         */

typedef struct pfrx_constraint {
  int32_t naqcomp;                   /* = the configuration's                                  */
  int32_t initialize_with_molality;  /* reaction%initialize_with_molality: values are molalities */
  int32_t max_iterations;            /* the reference gives up at 10000 (0 = that)             */
  int32_t reserved;
  const int32_t *type;               /* [naqcomp] PFRX_CONSTRAINT_*                            */
  /* the mineral / gas reaction of a MINERAL / GAS constraint on species i (mnrl_logK, mnrlspecid,
   * mnrlstoich, mnrlh2ostoich / paseq* of the reference): ln Q/K = -logK ln10 + h2o ln a_w +
   * sum_p stoich[p] ln(m gamma)[spec[p]], p in [eq_ptr[i], eq_ptr[i+1]); NULL when no species
   * has such a constraint */
  const double *eq_logK;             /* [naqcomp]                                              */
  const double *eq_logK_coef;        /* [naqcomp][5] or NULL: used when !use_isothermal        */
  const double *eq_h2o_stoich;       /* [naqcomp]                                              */
  const int32_t *eq_ptr;             /* [naqcomp + 1]                                          */
  const int32_t *eq_spec;            /* 0-based primary species                                */
  const double *eq_stoich;
} pfrx_constraint;

/* conc: device pointer, conc[i*ld + cell] = constraint value of species i in the cell (ld of the
 * bound state; the units of the CONSTRAINT block: mol/L or mol/kg, pH, bar, ...).
 * In: den_kg, temp, porosity, saturation, volume, mnrl_volfrac, soil particle density of the
 * bound state.  Out (bound state): pri_molal, total, sec_molal, pri_act_coef, sec_act_coef,
 * ln_act_h2o and, once equilibrated, the sorbed state (total_sorb_eq, free-site and
 * surface-complex concentrations, kinmr_total_sorb filled as at equilibrium, ion-exchange
 * state).  total and sec_molal are those of the last RTotal of the loop, as in the reference.
 * num_iterations / ierror: device int32 [ncell] or NULL.  ierror: 0 converged, 1 singular
 * Jacobian, 2 a non-positive (or NaN) free-ion concentration, 3 max_iterations reached -- the
 * three conditions under which the reference stops the run.  Inactive cells are skipped.
 * With use_full_geochemistry = 0: pri_molal and total from the values (reaction.F90:1472-1480). */
int pfrx_equilibrate_constraint(pfrx_handle *h, const pfrx_constraint *cons, const double *conc,
                                int32_t *num_iterations, int32_t *ierror);

/* ---- the steps either side of the cell loop ---------------------
 * PETSc keeps the transport unknowns and right-hand sides as BLOCK vectors, ncomp
 * values per cell: v[cell*ncomp + i].  The chemistry state here is cell-major SoA,
 * so the reference's per-component VecStrideGather / scatter loops around KSPSolve
 * (pmc_subsurface_osrt.F90:303-333) disappear -- component i of a SoA field IS the
 * contiguous work vector -- and what remains are three transposes of the bound
 * device state against block vectors in device memory (cells with imat <= 0 are
 * skipped like in the reference; entries the reference does not write stay
 * untouched):
 *   pfrx_os_fixed_accum   fixed_accum(cell, i) = porosity*sat*1000*volume*total(i), i < naqcomp
 *                         (:260-274; immobile entries untouched)
 *   pfrx_os_load          total(i) <- solved(cell, i) for i < naqcomp (:322-327, all
 *                         components at once) and immobile(k) <- tran_xx(cell, naqcomp+k)
 *                         (:356-359); either vector may be NULL
 *   pfrx_os_store         tran_xx(cell, i) <- pri_molal(i), tran_xx(cell, naqcomp+k) <-
 *                         immobile(k) after the step (:371-376)
 * All three are HBM-bound; they run on the handle's stream and return when done. */
int pfrx_os_fixed_accum(pfrx_handle *h, double *fixed_accum);
int pfrx_os_load(pfrx_handle *h, const double *solved_total, const double *tran_xx);
int pfrx_os_store(pfrx_handle *h, double *tran_xx);

/* The whole operator-split chemistry step of PMCSubsurfaceOSRTStepDT
 * (pmc_subsurface_osrt.F90:303-378) for a caller whose block vectors live in HOST
 * memory (PETSc Vecs) while the chemistry state stays bound in device memory from
 * step to step, the way rt_auxvars persist in the reference:
 *   upload solved_total (may be NULL: totals already in the state) and -- when the
 *   network has immobile species or the shard has inactive cells (imat <= 0) -- tran_xx;
 *   pfrx_os_load, RStep over tran_dt on every cell, pfrx_os_store;
 *   download tran_xx.
 * ncomp doubles per cell cross the link in each direction instead of the whole
 * state (pfrx_rstep_host).  Chunks of cells are pipelined over three streams (after
 * a warm-up call the library times one chunk against eight, keeps the faster and
 * repeats the trial every 64 calls; PFRX_OS_CHUNKS pins the count); pass page-locked
 * vectors (cudaHostRegister) for the copies to overlap the kernel.
 * Per-cell counts and flags stay in the bound state.                              */
int pfrx_os_step_host(pfrx_handle *h, const double *solved_total, double *tran_xx, double tran_dt,
                      pfrx_step_result *out);

/* ---- network-specialised kernels -------------------------------------------
 * The generic kernels read the reaction network from tables, the way the
 * reference's RTotalAqueous / RKineticMineral loops read reaction%eqcplxspecid
 * etc. (reaction.F90:4708-4757).  For a fixed network
 * pflotran_elm_interface_b200/specialize.py writes the same arithmetic with
 * every stoichiometric coefficient, logK and index as an immediate and nvcc
 * turns it into a cubin; pfrx_load_specialized attaches that cubin to a handle
 * and all later pfrx_rstep* calls launch it.  The cubin embeds
 * pfrx_config_signature() of the configuration it was generated from; a cubin
 * whose signature differs from the handle's is refused with PFRX_E_INVALID.
 * cubin_path == NULL detaches (back to the generic kernel).                  */
int pfrx_load_specialized(pfrx_handle *h, const char *cubin_path);
uint64_t pfrx_config_signature(pfrx_handle *h);
/* The configuration of the handle as a text file with exact bit patterns, for
 *   python -m pflotran_elm_interface_b200.specialize <file>
 * which generates and compiles the specialised cubins for it (set-up time; nvcc
 * needed there, not at run time): a host that flattened reaction_rt_type into
 * pfrx_config reaches the fast kernels without the Python deck reader.  The
 * signature computed from the file equals pfrx_config_signature(h).          */
int pfrx_config_dump(pfrx_handle *h, const char *path);
/* the same file and signature straight from a pfrx_config: no device needed
 * (set-up on a build or login node) */
int pfrx_config_write(const pfrx_config *cfg, const char *path);
uint64_t pfrx_config_signature_of(const pfrx_config *cfg);

/* diagnostics: measured FP64 FMA peak of `device` in TFLOP/s (the FP64
 * roofline denominator; MEASURED_PEAKS.json only carries HBM and bf16), and
 * the SM clock it implies at 64 FMA/clk/SM. */
int pfrx_diag_fp64_peak(int device, double *tflops, double *sm_mhz_est);

#ifdef __cplusplus
}
#endif
#endif /* PFRX_H */
