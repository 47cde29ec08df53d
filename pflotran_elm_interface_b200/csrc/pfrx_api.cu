// pfrx_api.cu -- C ABI of include/pfrx.h on top of the sm_100a kernels.
// No torch, no CPU fallback: every compute entry point needs a CUDA device and
// returns PFRX_E_CUDA otherwise.
#include <dlfcn.h>
#include <limits.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <chrono>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include "pfrx_device.cuh"

static thread_local char g_err[512] = "";

static int set_err(int code, const char *fmt, const char *a = "", const char *b = "") {
  snprintf(g_err, sizeof(g_err), fmt, a, b);
  return code;
}

#define CUDA_OK(call)                                                        \
  do {                                                                       \
    cudaError_t e_ = (call);                                                 \
    if (e_ != cudaSuccess) return set_err(PFRX_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

// ---- NCCL through dlopen (works with the system or the torch-bundled copy) --
typedef struct {
  char internal[128];
} NcclUid;
typedef void *NcclComm;
struct NcclApi {
  void *lib = nullptr;
  int (*GetUniqueId)(NcclUid *) = nullptr;
  int (*CommInitRank)(NcclComm *, int, NcclUid, int) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static int load_nccl() {
  if (g_nccl.lib) return PFRX_OK;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  void *h = nullptr;
  for (const char *n : names) {
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) return set_err(PFRX_E_NCCL, "dlopen(libnccl.so.2) failed: %s", dlerror());
#define SYM(field, name)                                                      \
  *(void **)(&g_nccl.field) = dlsym(h, name);                                 \
  if (!g_nccl.field) return set_err(PFRX_E_NCCL, "NCCL symbol %s missing", name);
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(AllReduce, "ncclAllReduce")
  SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  g_nccl.lib = h;
  return PFRX_OK;
}


// ---- CUDA driver API through dlopen (cubin loader for specialised kernels) ------
struct DrvApi {
  void *lib = nullptr;
  int (*ModuleLoad)(void **, const char *) = nullptr;
  int (*ModuleUnload)(void *) = nullptr;
  int (*ModuleGetFunction)(void **, void *, const char *) = nullptr;
  int (*ModuleGetGlobal)(unsigned long long *, size_t *, void *, const char *) = nullptr;
  int (*MemcpyDtoH)(void *, unsigned long long, size_t) = nullptr;
  int (*FuncSetAttribute)(void *, int, int) = nullptr;
  int (*OccupancyMaxActiveBlocksPerMultiprocessor)(int *, void *, int, size_t) = nullptr;
  int (*LaunchKernel)(void *, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, cudaStream_t,
                      void **, void **) = nullptr;
  int (*GetErrorString)(int, const char **) = nullptr;
};
static DrvApi g_drv;

static int load_driver() {
  if (g_drv.lib) return PFRX_OK;
  void *h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libcuda.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return set_err(PFRX_E_CUDA, "dlopen(libcuda.so.1) failed: %s", dlerror());
#define SYM(field, name)                                                     \
  *(void **)(&g_drv.field) = dlsym(h, name);                                 \
  if (!g_drv.field) return set_err(PFRX_E_CUDA, "driver symbol %s missing", name);
  SYM(ModuleLoad, "cuModuleLoad")
  SYM(ModuleUnload, "cuModuleUnload")
  SYM(ModuleGetFunction, "cuModuleGetFunction")
  SYM(ModuleGetGlobal, "cuModuleGetGlobal_v2")
  SYM(MemcpyDtoH, "cuMemcpyDtoH_v2")
  SYM(FuncSetAttribute, "cuFuncSetAttribute")
  SYM(OccupancyMaxActiveBlocksPerMultiprocessor, "cuOccupancyMaxActiveBlocksPerMultiprocessor")
  SYM(LaunchKernel, "cuLaunchKernel")
  SYM(GetErrorString, "cuGetErrorString")
#undef SYM
  g_drv.lib = h;
  return PFRX_OK;
}
static const char *drv_err(int rc) {
  const char *m = nullptr;
  if (g_drv.GetErrorString) g_drv.GetErrorString(rc, &m);
  return m ? m : "unknown driver error";
}
#define DRV_OK(call)                                                                  \
  do {                                                                                \
    int rc_ = (call);                                                                 \
    if (rc_ != 0) return set_err(PFRX_E_CUDA, "%s: %s", #call, drv_err(rc_));        \
  } while (0)

// ---- signature of the tables a specialised kernel bakes in ------------------------
// FNV-1a over the same byte sequence as specialize.py:signature()
static uint64_t fnv1a(uint64_t h, const void *p, size_t n) {
  const unsigned char *b = (const unsigned char *)p;
  for (size_t i = 0; i < n; i++) {
    h ^= b[i];
    h *= 0x100000001B3ull;
  }
  return h;
}
// One traversal serves the signature and pfrx_config_dump: `sink` sees every scalar group and
// every table the generator bakes in, in a fixed order.  name == NULL: anonymous scalar bytes.
struct CfgSink {
  uint64_t h = 0xCBF29CE484222325ull;
  std::string *dump = nullptr;  // when set: "T <name> <elem size> <count> <hex bytes>" per named table
  void bytes(const char *name, const void *p, size_t elem, size_t count) {
    h = fnv1a(h, p, elem * count);
    if (dump && name) {
      char head[160];
      snprintf(head, sizeof(head), "T %s %zu %zu ", name, elem, count);
      dump->append(head);
      static const char hx[] = "0123456789abcdef";
      const unsigned char *b = (const unsigned char *)p;
      for (size_t i = 0; i < elem * count; i++) {
        dump->push_back(hx[b[i] >> 4]);
        dump->push_back(hx[b[i] & 15]);
      }
      dump->push_back('\n');
    }
  }
};

static uint64_t config_walk(const pfrx_config *c, CfgSink &sink) {
  uint64_t &h = sink.h;

  int32_t head[12] = {c->naqcomp,          c->nimcomp,
                      c->neqcplx,          c->nkinmnrl,
                      c->nsrfcplxrxn,      c->nsrfcplx,
                      c->neqsrfcplxrxn,    c->nkinmrsrfcplxrxn,
                      c->clmcn_nrxn,       c->use_log_formulation,
                      c->act_coef_update_frequency, c->use_activity_h2o};
  double dh[3] = {c->debyeA, c->debyeB, c->debyeBdot};
  h = fnv1a(h, head, sizeof(head));
  h = fnv1a(h, dh, sizeof(dh));
#define ADD(ptr, count)                                                  \
  if ((ptr) && (count) > 0) sink.bytes(#ptr, (ptr), sizeof(*(ptr)), (size_t)(count));
  ADD(c->primary_spec_Z, c->naqcomp)
  ADD(c->primary_spec_a0, c->naqcomp)
  if (c->neqcplx > 0 && c->eqcplx_ptr) {
    int nnz = c->eqcplx_ptr[c->neqcplx];
    ADD(c->eqcplx_ptr, c->neqcplx + 1)
    ADD(c->eqcplx_specid, nnz)
    ADD(c->eqcplx_stoich, nnz)
    ADD(c->eqcplx_h2ostoich, c->neqcplx)
    ADD(c->eqcplx_logK, c->neqcplx)
    ADD(c->eqcplx_Z, c->neqcplx)
    ADD(c->eqcplx_a0, c->neqcplx)
  }
  if (c->nkinmnrl > 0 && c->kinmnrl_ptr) {
    int nnz = c->kinmnrl_ptr[c->nkinmnrl];
    ADD(c->kinmnrl_ptr, c->nkinmnrl + 1)
    ADD(c->kinmnrl_specid, nnz)
    ADD(c->kinmnrl_stoich, nnz)
    ADD(c->kinmnrl_h2ostoich, c->nkinmnrl)
    ADD(c->kinmnrl_logK, c->nkinmnrl)
    ADD(c->kinmnrl_molar_vol, c->nkinmnrl)
    ADD(c->kinmnrl_rate_constant, c->nkinmnrl)
    ADD(c->kinmnrl_activation_energy, c->nkinmnrl)
    ADD(c->kinmnrl_affinity_threshold, c->nkinmnrl)
    ADD(c->kinmnrl_rate_limiter, c->nkinmnrl)
    ADD(c->kinmnrl_irreversible, c->nkinmnrl)
  }
  if (c->nsrfcplxrxn > 0 && c->srfcplxrxn_ptr && c->srfcplx_ptr) {
    int nnz = c->srfcplx_ptr[c->nsrfcplx];
    ADD(c->srfcplxrxn_ptr, c->nsrfcplxrxn + 1)
    ADD(c->srfcplxrxn_to_complex, c->srfcplxrxn_ptr[c->nsrfcplxrxn])
    ADD(c->srfcplxrxn_surf_type, c->nsrfcplxrxn)
    ADD(c->srfcplxrxn_to_surf, c->nsrfcplxrxn)
    ADD(c->srfcplxrxn_site_density, c->nsrfcplxrxn)
    ADD(c->srfcplx_ptr, c->nsrfcplx + 1)
    ADD(c->srfcplx_specid, nnz)
    ADD(c->srfcplx_stoich, nnz)
    ADD(c->srfcplx_h2ostoich, c->nsrfcplx)
    ADD(c->srfcplx_free_site_stoich, c->nsrfcplx)
    ADD(c->srfcplx_logK, c->nsrfcplx)
    ADD(c->eqsrfcplxrxn_to_srfcplxrxn, c->neqsrfcplxrxn)
    if (c->nkinmrsrfcplxrxn > 0 && c->kinmr_rate_ptr) {
      const int nm = c->nkinmrsrfcplxrxn;
      ADD(c->kinmrsrfcplxrxn_to_srfcplxrxn, nm)
      ADD(c->kinmr_rate_ptr, nm + 1)
      ADD(c->kinmr_rate, c->kinmr_rate_ptr[nm])
      ADD(c->kinmr_frac, c->kinmr_rate_ptr[nm])
    }
  }
  if (c->clmcn_nrxn > 0) {
    int32_t ch[3] = {c->clmcn_npool, c->clmcn_C_species_id, c->clmcn_N_species_id};
    h = fnv1a(h, ch, sizeof(ch));
    ADD(c->clmcn_CN_ratio, c->clmcn_npool)
    ADD(c->clmcn_pool_nspec, c->clmcn_npool)
    ADD(c->clmcn_pool_C_id, c->clmcn_npool)
    ADD(c->clmcn_pool_N_id, c->clmcn_npool)
    ADD(c->clmcn_upstream_pool_id, c->clmcn_nrxn)
    ADD(c->clmcn_downstream_pool_id, c->clmcn_nrxn)
    ADD(c->clmcn_rate_constant, c->clmcn_nrxn)
    ADD(c->clmcn_respiration_fraction, c->clmcn_nrxn)
    ADD(c->clmcn_inhibition_constant, c->clmcn_nrxn)
  }
  // ELM-CN sandboxes: every parameter the generated code bakes in
  if (c->somdec) {
    const pfrx_somdec *sd = c->somdec;
    int32_t hi[14] = {sd->nrxn,   sd->co2_id,  sd->co2_itype, sd->o2_id,   sd->o2_itype, sd->nh4_id,  sd->no3_id,
                      sd->n2o_id, sd->proton_id, sd->hr_id,   sd->nmin_id, sd->nimm_id,  sd->nimp_id, sd->ngasmin_id};
    double hd[3] = {sd->x0eps, sd->n2o_frac_mineralization, sd->inhibition_nh4_no3};
    h = fnv1a(h, hi, sizeof(hi));
    h = fnv1a(h, hd, sizeof(hd));
    const int nx = sd->nrxn, nd = sd->downstream_ptr[nx], nm = sd->monod_ptr[nx], ni = sd->inhib_ptr[nx];
    ADD(sd->rate_constant, nx)
    ADD(sd->rate_decomposition, nx)
    ADD(sd->rate_ad_factor, nx)
    ADD(sd->upstream_c_id, nx)
    ADD(sd->upstream_n_id, nx)
    ADD(sd->upstream_is_aqueous, nx)
    ADD(sd->upstream_hr_id, nx)
    ADD(sd->upstream_nmin_id, nx)
    ADD(sd->upstream_nimp_id, nx)
    ADD(sd->upstream_nimm_id, nx)
    ADD(sd->upstream_nc, nx)
    ADD(sd->mineral_c_stoich, nx)
    ADD(sd->mineral_n_stoich, nx)
    ADD(sd->downstream_ptr, nx + 1)
    ADD(sd->downstream_c_id, nd)
    ADD(sd->downstream_n_id, nd)
    ADD(sd->downstream_is_aqueous, nd)
    ADD(sd->downstream_stoich, nd)
    ADD(sd->downstream_nc, nd)
    ADD(sd->temperature_response_function, nx)
    ADD(sd->moisture_response_function, nx)
    ADD(sd->ox_response_function, nx)
    ADD(sd->q10, nx)
    ADD(sd->ea, nx)
    ADD(sd->ox_half_saturation, nx)
    ADD(sd->decomp_depth_efolding, nx)
    ADD(sd->ox_specid, nx)
    ADD(sd->ox_specitype, nx)
    ADD(sd->monod_ptr, nx + 1)
    ADD(sd->monod_specid, nm)
    ADD(sd->monod_specitype, nm)
    ADD(sd->monod_pool_normalized, nm)
    ADD(sd->monod_half_saturation, nm)
    ADD(sd->monod_threshold, nm)
    ADD(sd->inhib_ptr, nx + 1)
    ADD(sd->inhib_itype, ni)
    ADD(sd->inhib_specid, ni)
    ADD(sd->inhib_specitype, ni)
    ADD(sd->inhib_constant, ni)
    ADD(sd->inhib_constant2, ni)
  }
  if (c->nitrif) {
    const pfrx_nitrif *nt = c->nitrif;
    int32_t hi[5] = {nt->proton_id, nt->nh4_id, nt->no3_id, nt->n2o_id, nt->ngasnit_id};
    double hd[3] = {nt->k_nitr_max, nt->k_nitr_n2o, nt->x0eps};
    h = fnv1a(h, hi, sizeof(hi));
    h = fnv1a(h, hd, sizeof(hd));
  }
  if (c->denitr) {
    const pfrx_denitr *dn = c->denitr;
    int32_t hi[4] = {dn->no3_id, dn->n2_id, dn->n2o_id, dn->ngasdeni_id};
    double hd[3] = {dn->half_saturation, dn->k_deni_max, dn->x0eps};
    h = fnv1a(h, hi, sizeof(hi));
    h = fnv1a(h, hd, sizeof(hd));
  }
  if (c->plantn) {
    const pfrx_plantn *pn = c->plantn;
    int32_t hi[6] = {pn->nh4_id, pn->no3_id, pn->plantn_id, pn->plantndemand_id, pn->plantnh4uptake_id,
                     pn->plantno3uptake_id};
    double hd[5] = {pn->half_saturation_nh4, pn->half_saturation_no3, pn->inhibition_nh4_no3, pn->x0eps_nh4,
                    pn->x0eps_no3};
    h = fnv1a(h, hi, sizeof(hi));
    h = fnv1a(h, hd, sizeof(hd));
  }
  if (c->langmuir) {
    const pfrx_langmuir *lg = c->langmuir;
    int32_t hi[2] = {lg->aq_id, lg->sorb_id};
    double hd[3] = {lg->k_kinetic, lg->k_equilibrium, lg->s_max};
    h = fnv1a(h, hi, sizeof(hi));
    h = fnv1a(h, hd, sizeof(hd));
  }
  if (c->cndegas) {  // not covered by the generator: any cubin signature must differ
    h = fnv1a(h, c->cndegas, sizeof(*c->cndegas));
  }
  if (c->calcite) h = fnv1a(h, c->calcite, sizeof(*c->calcite));
  if (c->radon) h = fnv1a(h, c->radon, sizeof(*c->radon));
  if (c->nactive_gas > 0 && c->acteq_ptr) {  // not covered by the generator either
    const int ng = c->nactive_gas, nnz = c->acteq_ptr[ng];
    ADD(c->acteq_ptr, ng + 1)
    ADD(c->acteq_specid, nnz)
    ADD(c->acteq_stoich, nnz)
    ADD(c->acteq_h2ostoich, ng)
    ADD(c->acteq_logK, ng)
  }
  if (c->somdec || c->nitrif || c->denitr || c->plantn || c->langmuir) {
    int32_t e = c->elm_pflotran ? (c->elm_flow_coupled ? 3 : 1) : 0;
    h = fnv1a(h, &e, sizeof(e));
    if (c->sandbox_list) ADD(c->sandbox_list, c->nsandbox)
  }
  // what the generator bakes in or refuses beyond the tables above: a cubin cached by signature
  // must not survive a change of any of these (presence and contents)
  {
    int32_t feat[5] = {c->h2o_aq_id, c->use_isothermal, c->act_coef_update_algorithm, c->use_total_as_guess,
                       c->use_full_geochemistry};
    h = fnv1a(h, feat, sizeof(feat));
    ADD(c->kinmnrl_Temkin_const, c->nkinmnrl)
    ADD(c->kinmnrl_min_scale_factor, c->nkinmnrl)
    ADD(c->kinmnrl_affinity_power, c->nkinmnrl)
    ADD(c->kinmnrl_num_prefactors, c->nkinmnrl)
    ADD(c->srfcplxrxn_stoich_flag, c->nsrfcplxrxn)
  }
  // ion exchange, KD isotherms, dynamic KD (generated by specialize.gen_sorption)
  {
    int32_t sb[4] = {c->neqionxrxn, c->neqkdrxn, c->neqdynamickdrxn, c->neqkdrxn > 0 ? c->ikd_units : 0};
    h = fnv1a(h, sb, sizeof(sb));
    if (c->neqionxrxn > 0 && c->eqionx_ptr) {
      const int n = c->neqionxrxn, nnz = c->eqionx_ptr[n];
      ADD(c->eqionx_ptr, n + 1)
      ADD(c->eqionx_cationid, nnz)
      ADD(c->eqionx_k, nnz)
      ADD(c->eqionx_CEC, n)
      ADD(c->eqionx_to_surf, n)
      ADD(c->eqionx_Z_flag, n)
    }
    if (c->neqkdrxn > 0) {
      const int n = c->neqkdrxn;
      ADD(c->eqkd_specid, n)
      ADD(c->eqkd_type, n)
      ADD(c->eqkd_mineral, n)
      ADD(c->eqkd_coeff, n)
      ADD(c->eqkd_langmuir_b, n)
      ADD(c->eqkd_freundlich_n, n)
    }
    if (c->neqdynamickdrxn > 0) {
      const int n = c->neqdynamickdrxn;
      ADD(c->eqdynamickd_specid, n)
      ADD(c->eqdynamickd_refspecid, n)
      ADD(c->eqdynamickd_refspechigh, n)
      ADD(c->eqdynamickd_low, n)
      ADD(c->eqdynamickd_high, n)
      ADD(c->eqdynamickd_power, n)
    }
  }
  // general / radioactive-decay / immobile-decay / microbial reactions (generated by specialize.gen_kinetic)
  {
    int32_t k3[5] = {c->ngeneral_rxn, c->nradiodecay_rxn, c->nimmobile_decay_rxn, c->nmicrobial_rxn,
                     c->nmicrobial_rxn > 0 ? c->microbial_concentration_units : 0};
    h = fnv1a(h, k3, sizeof(k3));
    if (c->ngeneral_rxn > 0 && c->general_ptr && c->general_fwd_ptr && c->general_bwd_ptr) {
      const int n = c->ngeneral_rxn;
      ADD(c->general_ptr, n + 1)
      ADD(c->general_specid, c->general_ptr[n])
      ADD(c->general_stoich, c->general_ptr[n])
      ADD(c->general_fwd_ptr, n + 1)
      ADD(c->general_fwd_specid, c->general_fwd_ptr[n])
      ADD(c->general_fwd_stoich, c->general_fwd_ptr[n])
      ADD(c->general_bwd_ptr, n + 1)
      ADD(c->general_bwd_specid, c->general_bwd_ptr[n])
      ADD(c->general_bwd_stoich, c->general_bwd_ptr[n])
      ADD(c->general_kf, n)
      ADD(c->general_kr, n)
    }
    if (c->nradiodecay_rxn > 0 && c->radiodecay_ptr) {
      const int n = c->nradiodecay_rxn;
      ADD(c->radiodecay_ptr, n + 1)
      ADD(c->radiodecay_specid, c->radiodecay_ptr[n])
      ADD(c->radiodecay_stoich, c->radiodecay_ptr[n])
      ADD(c->radiodecay_forward_specid, n)
      ADD(c->radiodecay_kf, n)
    }
    if (c->nimmobile_decay_rxn > 0) {
      ADD(c->immobile_decay_specid, c->nimmobile_decay_rxn)
      ADD(c->immobile_decay_constant, c->nimmobile_decay_rxn)
    }
    if (c->nmicrobial_rxn > 0 && c->microbial_ptr && c->microbial_monod_ptr && c->microbial_inhibition_ptr) {
      const int n = c->nmicrobial_rxn;
      ADD(c->microbial_ptr, n + 1)
      ADD(c->microbial_specid, c->microbial_ptr[n])
      ADD(c->microbial_stoich, c->microbial_ptr[n])
      ADD(c->microbial_rate_constant, n)
      ADD(c->microbial_activation_energy, n)
      ADD(c->microbial_monod_ptr, n + 1)
      ADD(c->microbial_monod_specid, c->microbial_monod_ptr[n])
      ADD(c->microbial_monod_K, c->microbial_monod_ptr[n])
      ADD(c->microbial_monod_Cth, c->microbial_monod_ptr[n])
      ADD(c->microbial_inhibition_ptr, n + 1)
      ADD(c->microbial_inhibition_specid, c->microbial_inhibition_ptr[n])
      ADD(c->microbial_inhibition_type, c->microbial_inhibition_ptr[n])
      ADD(c->microbial_inhibition_C, c->microbial_inhibition_ptr[n])
      ADD(c->microbial_inhibition_C2, c->microbial_inhibition_ptr[n])
      ADD(c->microbial_biomassid, n)
      ADD(c->microbial_biomass_yield, n)
    }
  }
#undef ADD
  return h;
}

static uint64_t config_signature(const pfrx_config *c) {
  CfgSink sink;
  return config_walk(c, sink);
}

static void hex_struct(std::string &out, const char *tag, const void *p, size_t n) {
  char head[64];
  snprintf(head, sizeof(head), "S %s %zu ", tag, n);
  out.append(head);
  static const char hx[] = "0123456789abcdef";
  const unsigned char *b = (const unsigned char *)p;
  for (size_t i = 0; i < n; i++) {
    out.push_back(hx[b[i] >> 4]);
    out.push_back(hx[b[i] & 15]);
  }
  out.push_back('\n');
}

// text form of a configuration for `python -m pflotran_elm_interface_b200.specialize <file>`:
// the structs as raw bytes (scalars; pointer members are ignored by the reader) and every table the
// generator reads, with exact bit patterns -- the signature computed from the file is the handle's
static std::string config_dump_text(const pfrx_config *c) {
  std::string out;
  char head[96];
  snprintf(head, sizeof(head), "pfrx_config_dump 1 abi %d\n", PFRX_ABI_VERSION);
  out.append(head);
  hex_struct(out, "config", c, sizeof(*c));
  if (c->somdec) hex_struct(out, "somdec", c->somdec, sizeof(*c->somdec));
  if (c->nitrif) hex_struct(out, "nitrif", c->nitrif, sizeof(*c->nitrif));
  if (c->denitr) hex_struct(out, "denitr", c->denitr, sizeof(*c->denitr));
  if (c->plantn) hex_struct(out, "plantn", c->plantn, sizeof(*c->plantn));
  if (c->langmuir) hex_struct(out, "langmuir", c->langmuir, sizeof(*c->langmuir));
  if (c->cndegas) hex_struct(out, "cndegas", c->cndegas, sizeof(*c->cndegas));
  if (c->calcite) hex_struct(out, "calcite", c->calcite, sizeof(*c->calcite));
  if (c->radon) hex_struct(out, "radon", c->radon, sizeof(*c->radon));
  CfgSink sink;
  sink.dump = &out;
  const uint64_t sig = config_walk(c, sink);
  if (c->sandbox_list && c->nsandbox > 0 && !(c->somdec || c->nitrif || c->denitr || c->plantn || c->langmuir)) {
    CfgSink extra;  // the order of a CLM-CN-only list is read by the generator but is not part of the signature
    extra.dump = &out;
    extra.bytes("c->sandbox_list", c->sandbox_list, sizeof(*c->sandbox_list), (size_t)c->nsandbox);
  }
  snprintf(head, sizeof(head), "signature %016llx\n", (unsigned long long)sig);
  out.append(head);
  return out;
}

#define PFRX_MAX_CHUNKS 32

// ---- handle -----------------------------------------------------------------
struct pfrx_handle {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;
  cudaStream_t stream2 = nullptr;  // second kernel stream of pfrx_os_step_host (odd chunks), created on first use
  cudaEvent_t ev_reset = nullptr;
  cudaStream_t out_stream = nullptr;
  cudaEvent_t ev_in[PFRX_MAX_CHUNKS], ev_k[PFRX_MAX_CHUNKS];
  bool ev_ready = false;
  DevCfg cfg;
  void *arena = nullptr;  // device copy of all tables
  int n = 0, npad = 0, lanes = 1;
  bool tpc = false;  // thread-per-cell kernel (pfrx_tpc.cuh)
  size_t smem_bytes = 0;
  int threads = 128;
  int blocks_per_sm = 1, sm_count = 1;
  void (*kernel)(DevCfg, DevState, int64_t, double, DevSummary *) = nullptr;
  // bound state
  bool bound = false;
  int64_t ncell = 0;
  DevState st;
  DevSummary *d_summ = nullptr;
  DevSummary *h_summ = nullptr;  // pinned
  bool pending = false;
  int64_t launches = 0;
  // row counts of every field, in pfrx_state order
  std::vector<int> rows_d;  // 20 double fields
  // device-side reduction of the shard summary, enqueued by pfrx_rstep_async behind the kernel
  long long *d_red_step = nullptr, *h_red_step = nullptr;
  bool red_inflight = false, red_valid = false;
  pfrx_step_result red_local;
  std::string dump_text;          // pfrx_config_dump
  std::vector<double> pri_Z_host;  // primary_spec_Z (charge-balance constraints)
  std::vector<int> sr_flag_host;  // srfcplxrxn_stoich_flag (host copy: pfrx_load_specialized refuses inner-Newton sites)
  // owned device state for pfrx_rstep_host
  void *own = nullptr;
  int64_t own_ncell = 0;
  DevState own_st;
  // network-specialised kernel (pfrx_load_specialized)
  uint64_t sig = 0;
  SpecParams spec_prm;
  void *spec_module = nullptr, *spec_func = nullptr;
  int spec_threads = 0, spec_blocks_per_sm = 0, spec_cells = 0;
  // longest-first hand-out of the refill skeleton: after a whole-shard launch the cells are sorted by the
  // Newton iterations they just needed (descending), the next launch on the same shard starts with the slowest
  int spec_refill = 0;
  int order_mode = 1;  // PFRX_CELL_ORDER=0 / pfrx_cell_order(h, 0) switch it off
  int *d_order = nullptr, *d_iota = nullptr, *d_keys_out = nullptr;
  void *d_sort_tmp = nullptr;
  size_t sort_tmp_bytes = 0;
  int64_t order_cap = 0, order_ncell = 0;
  const int *order_key = nullptr;  // the num_iterations array the order was built from
  bool order_valid = false;
  size_t spec_smem = 0;
  int64_t last_h2d = 0, last_d2h = 0;  // bytes moved by the latest pfrx_rstep_host / pfrx_os_step_host
  // device staging of the two block vectors of pfrx_os_step_host
  double *os_a = nullptr, *os_b = nullptr;
  int64_t os_cap = 0;
  // pfrx_os_step_host tries one chunk on its first call and eight on its second, then keeps the
  // faster: a ragged workload pays the slowest cell of every chunk, a uniform one gains the overlap
  double os_trial_s[2] = {0.0, 0.0};
  int64_t os_trial_ncell = -1;
  int os_calls = 0;       // position in the trial cycle of pfrx_os_step_host (see there)
  int os_choice = 1;      // chunk count in use
  int os_env_chunks = -1; // PFRX_OS_CHUNKS, read once (0: not set)
  int os_inactive = -1;   // inactive cells of the bound shard (-1: not counted yet)
  uint64_t host_resident_mask = 0;  // pfrx_rstep_host_resident
  // kernel seconds per cell and link seconds per cell seen by the latest pfrx_rstep_host: a step
  // whose kernel outweighs its transfers gains nothing from many chunks and, with the refill
  // skeleton, pays the slowest cell of every chunk
  double host_kernel_s_per_cell = 0.0, host_link_s_per_cell = 0.0;
  cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
  // batched RReaction (pfrx_reaction): always the thread-per-cell layout
  DevCfg rx_cfg;
  void (*rx_kernel)(DevCfg, DevState, int64_t, int, double *, double *, double) = nullptr;
  int rx_threads = 0, rx_blocks_per_sm = 0;
  size_t rx_smem = 0;
  // nccl
  NcclComm comm = nullptr;
  long long *d_red = nullptr;
  long long *h_red = nullptr;
  int nranks = 1;
};

// table arena builder
struct Arena {
  std::vector<unsigned char> bytes;
  std::vector<std::pair<size_t, void **>> fix;  // (offset, address of device pointer field)
  template <typename T>
  void add(const T *src, size_t count, const T **field) {
    if (!src || count == 0) {
      *field = nullptr;
      return;
    }
    size_t off = (bytes.size() + 15) & ~size_t(15);
    bytes.resize(off + count * sizeof(T));
    memcpy(bytes.data() + off, src, count * sizeof(T));
    fix.push_back({off, (void **)field});
  }
};

typedef void (*pfrx_kernel_fn)(DevCfg, DevState, int64_t, double, DevSummary *);
// one getter per padded size N, defined in pfrx_kern.cu (see build.py)
#define PFRX_DECL(N) extern "C" pfrx_kernel_fn pfrx_kernel_##N(int lanes);
PFRX_DECL(3) PFRX_DECL(4) PFRX_DECL(8) PFRX_DECL(13) PFRX_DECL(15) PFRX_DECL(16) PFRX_DECL(32)
typedef void (*pfrx_reaction_fn)(DevCfg, DevState, int64_t, int, double *, double *, double);
#define PFRX_DECLR(N) extern "C" pfrx_reaction_fn pfrx_reaction_kernel_##N(void);
PFRX_DECLR(3) PFRX_DECLR(4) PFRX_DECLR(8) PFRX_DECLR(13) PFRX_DECLR(15) PFRX_DECLR(16) PFRX_DECLR(32)
typedef void (*pfrx_constraint_fn)(DevCfg, DevState, int64_t, DevCons, int *, int *);
#define PFRX_DECLC(N) extern "C" pfrx_constraint_fn pfrx_constraint_kernel_##N(void);
PFRX_DECLC(3) PFRX_DECLC(4) PFRX_DECLC(8) PFRX_DECLC(13) PFRX_DECLC(15) PFRX_DECLC(16) PFRX_DECLC(32)
typedef void (*pfrx_auxvars_fn)(DevCfg, DevState, int64_t, const double *, int);
#define PFRX_DECLA(N) extern "C" pfrx_auxvars_fn pfrx_auxvars_kernel_##N(void);
PFRX_DECLA(3) PFRX_DECLA(4) PFRX_DECLA(8) PFRX_DECLA(13) PFRX_DECLA(15) PFRX_DECLA(16) PFRX_DECLA(32)
struct KernelGetter {
  int n;
  pfrx_kernel_fn (*get)(int);
  pfrx_reaction_fn (*get_rx)(void);
  pfrx_constraint_fn (*get_cons)(void);
  pfrx_auxvars_fn (*get_aux)(void);
};
static const KernelGetter g_getters[] = {
    {3, pfrx_kernel_3, pfrx_reaction_kernel_3, pfrx_constraint_kernel_3, pfrx_auxvars_kernel_3},    {4, pfrx_kernel_4, pfrx_reaction_kernel_4, pfrx_constraint_kernel_4, pfrx_auxvars_kernel_4},
    {8, pfrx_kernel_8, pfrx_reaction_kernel_8, pfrx_constraint_kernel_8, pfrx_auxvars_kernel_8},    {13, pfrx_kernel_13, pfrx_reaction_kernel_13, pfrx_constraint_kernel_13, pfrx_auxvars_kernel_13},
    {15, pfrx_kernel_15, pfrx_reaction_kernel_15, pfrx_constraint_kernel_15, pfrx_auxvars_kernel_15}, {16, pfrx_kernel_16, pfrx_reaction_kernel_16, pfrx_constraint_kernel_16, pfrx_auxvars_kernel_16},
    {32, pfrx_kernel_32, pfrx_reaction_kernel_32, pfrx_constraint_kernel_32, pfrx_auxvars_kernel_32}};

static bool default_tpc(int npad) {
  // measured on B200: the thread-per-cell kernel wins for small networks (C2: 2.6x);
  // at 13-15 unknowns its Jacobians leave room for only two warps per SM
  return npad <= 4;
}

static int default_lanes(int npad) {
  // measured on B200 (profiles/): thread-per-cell up to 4 unknowns, 8 lanes for
  // the 13-dof CLM-CN network, 16 lanes for the 15-dof Hanford network
  if (npad <= 4) return 1;
  if (npad <= 8) return 4;
  if (npad <= 13) return 8;
  if (npad <= 16) return 16;
  return 32;
}

static int pick_kernel(pfrx_handle *h, int want_lanes) {
  const KernelGetter *gt = nullptr;
  for (const auto &k : g_getters)
    if (k.n >= h->n && (!gt || k.n < gt->n)) gt = &k;
  if (!gt) return set_err(PFRX_E_LIMIT, "ncomp exceeds PFRX_MAX_NCOMP%s", "");
  int lanes = want_lanes > 0 ? want_lanes : default_lanes(gt->n);
  if (want_lanes == 0) {
    if (const char *ev = getenv("PFRX_TPC")) {
      if (atoi(ev) != 0) lanes = 0;
    } else if (default_tpc(gt->n)) {
      lanes = 0;
    }
  }
  if (lanes == 0) {
    h->npad = gt->n;
    h->lanes = 1;
    h->tpc = true;
    h->kernel = gt->get(0);
    return PFRX_OK;
  }
  pfrx_kernel_fn fn = gt->get(lanes);
  if (!fn) {
    // nearest instantiated lane count
    const int cand[] = {1, 2, 4, 8, 16, 32};
    int best = -1;
    for (int c : cand)
      if (gt->get(c) && (best < 0 || abs(c - lanes) < abs(best - lanes))) best = c;
    if (best < 0) return set_err(PFRX_E_LIMIT, "no kernel variant for this size%s", "");
    lanes = best;
    fn = gt->get(lanes);
  }
  h->npad = gt->n;
  h->lanes = lanes;
  h->kernel = fn;
  return PFRX_OK;
}

// double fields of pfrx_state in header order: 20 of ABI v1, then the seven ELM
// scalars and the SOMDECOMP N:C memory
#define PFRX_NUM_D 39
static int field_rows(const pfrx_config *c, int *rows /*PFRX_NUM_D*/) {
  int mr = 0;
  if (c->nkinmrsrfcplxrxn > 0) mr = c->naqcomp * (c->kinmr_rate_ptr[c->nkinmrsrfcplxrxn] + c->nkinmrsrfcplxrxn);
  const int e = c->elm_pflotran ? 1 : 0;
  const int nc = c->somdec ? c->somdec->nrxn + c->somdec->downstream_ptr[c->somdec->nrxn] : 0;
  const int nix = c->neqionxrxn, nixc = c->neqionxrxn > 0 ? c->eqionx_ptr[c->neqionxrxn] : 0;
  const int nsorb = c->neqsrfcplxrxn + c->neqionxrxn + c->neqkdrxn + c->neqdynamickdrxn;
  int r[PFRX_NUM_D] = {c->naqcomp, c->naqcomp, c->nimcomp, c->naqcomp, c->neqcplx, c->neqcplx, 1, c->nkinmnrl,
                       c->nkinmnrl, c->nkinmnrl, c->nsrfcplxrxn, c->nsrfcplx, nsorb > 0 ? c->naqcomp : 0,
                       mr, 1, 1, 1, 1, 1, 1, e, e, e, e, e, e, e, nc, e, nix, nixc,
                       (c->cndegas && c->cndegas->cell_state_mode >= 1) ? 1 : 0, c->calcite ? 1 : 0,
                       c->nactive_gas > 0 ? 1 : 0, c->nactive_gas > 0 ? c->naqcomp : 0, c->nactive_gas > 0 ? c->nactive_gas : 0,
                       (c->elm_pflotran && c->elm_flow_coupled) ? 1 : 0, (c->elm_pflotran && c->elm_flow_coupled) ? 1 : 0,
                       (c->elm_pflotran && c->elm_flow_coupled) ? 1 : 0};
  memcpy(rows, r, sizeof(r));
  return 0;
}

extern "C" int pfrx_abi_version(void) { return PFRX_ABI_VERSION; }
extern "C" const char *pfrx_last_error(void) { return g_err; }
extern "C" int64_t pfrx_sizeof(int which) {
  switch (which) {
    case 0: return (int64_t)sizeof(pfrx_config);
    case 1: return (int64_t)sizeof(pfrx_state);
    case 2: return (int64_t)sizeof(pfrx_step_result);
  }
  return -1;
}

// thread-per-cell workspace: one slice per THREAD, exact-size Jacobian
static void tpc_layout(DevCfg &d, int N) {
  int off = 0;
  auto take = [&](int cnt) {
    int o = off;
    off += cnt;
    return o;
  };
  const bool act_upd = d.act_freq == PFRX_ACT_COEF_FREQUENCY_NEWTON_ITER;
  d.off_c = take(N);
  d.off_lnact = take(N);
  d.off_invc = take(N);
  d.off_res = take(N);
  d.off_acc = take(N);
  d.off_tmp = take(N + 2 * d.nsrfcplx + 2);
  d.off_ts = take(N);
  d.js = d.n;
  d.off_J = take(d.n * d.n);
  d.off_cls = take(d.ncls + 1);
  d.off_x = take((N + 1) / 2 + 1);
  d.off_xs = take(N);
  d.off_sc = take(d.nsrfcplx + 1);
  d.off_mn = take(d.nkin + 1);
  d.off_fs = take(d.nsrfrxn + 1);
  d.off_mr = take(2 * d.nmr * N + 1);
  d.off_lng = act_upd ? 0 : take(d.ncplx);
  d.off_sec = 0;
  d.off_dt = d.need_dt ? take(d.naq * d.naq) : 0;
  d.off_ds = d.need_ds ? take(d.naq * d.naq) : 0;
  d.off_nc = d.n_nc > 0 ? take(d.n_nc) : 0;
  d.off_ix = d.nionx > 0 ? take(d.nionx + d.n_ixcat) : 0;
  d.off_tg = d.ngas > 0 ? take(d.naq) : 0;
  d.off_dg = d.ngas > 0 ? take(d.naq * d.naq) : 0;
  d.ws_stride = off | 1;
}

static int layout_and_launch_params(pfrx_handle *h) {
  DevCfg &d = h->cfg;
  const int N = h->npad;
  int off = 0;
  auto take = [&](int cnt) {
    int o = off;
    off += cnt;
    return o;
  };
  if (h->tpc) {
    tpc_layout(d, N);
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, h->device));
    h->sm_count = prop.multiProcessorCount;
    size_t per_thread = (size_t)d.ws_stride * sizeof(double);
    // the largest block (<= 128 threads) whose workspaces fit, preferring a
    // size that lets several blocks share an SM
    int t = h->threads;
    while (t > 32 && per_thread * t > (size_t)prop.sharedMemPerBlockOptin) t /= 2;
    if (per_thread * t > (size_t)prop.sharedMemPerBlockOptin)
      return set_err(PFRX_E_LIMIT, "per-cell workspace does not fit shared memory%s", "");
    CUDA_OK(cudaFuncSetAttribute((const void *)h->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(per_thread * t)));
    if (!getenv("PFRX_THREADS")) {
      // shared memory decides how many cells an SM holds: take the block size with the most resident
      // threads (C8: one 128-thread block per SM, but three of 64 -> 36.3 ms becomes 26.6 ms per 1 M cells)
      int best_t = t, best_res = 0;
      for (int tt = t; tt >= 32; tt -= 32) {
        int nbt = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nbt, (const void *)h->kernel, tt, per_thread * tt) !=
            cudaSuccess)
          continue;
        if (nbt * tt > best_res) {
          best_res = nbt * tt;
          best_t = tt;
        }
      }
      t = best_t;
    }
    h->threads = t;
    h->smem_bytes = per_thread * t;
    CUDA_OK(cudaFuncSetAttribute((const void *)h->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)h->smem_bytes));
    int nb = 0;
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void *)h->kernel, h->threads, h->smem_bytes));
    if (nb < 1) return set_err(PFRX_E_LIMIT, "kernel does not fit on an SM%s", "");
    h->blocks_per_sm = nb;
    return PFRX_OK;
  }
  d.off_c = take(N);
  d.off_lnact = take(N);
  d.off_invc = take(N);
  d.off_x = take(2 * (N + 2));
  d.off_xs = take(N);
  d.off_sec = take(d.ncplx);
  d.off_lng = take(d.ncplx);
  d.js = N | 1;
  d.off_J = take(N * d.js);
  d.off_acc = take(d.nacc + 1);  // must follow ws.J: task destinations are offsets from ws.J
  d.off_tmp = take(N + 2 * d.nsrfcplx + 2);
  d.off_sc = take(d.nsrfcplx + 1);
  d.off_cls = take(d.ncls + 1);
  d.off_mn = take(3 * d.nkin + 1);
  d.off_fs = take(d.nsrfrxn + 1);
  d.off_mr = take(2 * d.nmr * N + 1);
  d.off_res = take(N);
  d.off_ts = take(N);
  d.ws_stride = off | 1;
  int groups = h->threads / h->lanes;
  h->smem_bytes = (size_t)groups * d.ws_stride * sizeof(double);
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, h->device));
  h->sm_count = prop.multiProcessorCount;
  if (h->smem_bytes > (size_t)prop.sharedMemPerBlockOptin) {
    // shrink the block until the workspace fits
    while (h->threads > 32 && h->smem_bytes > (size_t)prop.sharedMemPerBlockOptin) {
      h->threads /= 2;
      groups = h->threads / h->lanes;
      h->smem_bytes = (size_t)groups * d.ws_stride * sizeof(double);
    }
    if (h->smem_bytes > (size_t)prop.sharedMemPerBlockOptin)
      return set_err(PFRX_E_LIMIT, "per-cell workspace does not fit shared memory%s", "");
  }
  CUDA_OK(cudaFuncSetAttribute((const void *)h->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)h->smem_bytes));
  int nb = 0;
  CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void *)h->kernel, h->threads, h->smem_bytes));
  if (nb < 1) return set_err(PFRX_E_LIMIT, "kernel does not fit on an SM%s", "");
  h->blocks_per_sm = nb;
  return PFRX_OK;
}

extern "C" void pfrx_destroy(pfrx_handle *h);
extern "C" int pfrx_create(const pfrx_config *c, int device, pfrx_handle **out) {
  if (!c || !out) return set_err(PFRX_E_INVALID, "null argument%s", "");
  if (c->abi_version != PFRX_ABI_VERSION) return set_err(PFRX_E_INVALID, "abi_version mismatch%s", "");
  int n = c->naqcomp + c->nimcomp;
  if (n < 1 || n > PFRX_MAX_NCOMP) return set_err(PFRX_E_LIMIT, "ncomp out of range%s", "");
  if (c->neqcplx > 4095) return set_err(PFRX_E_LIMIT, "more than 4095 secondary complexes%s", "");
  const bool act_newton = c->act_coef_update_algorithm == PFRX_ACT_COEF_ALGORITHM_NEWTON &&
                          c->act_coef_update_frequency == PFRX_ACT_COEF_FREQUENCY_NEWTON_ITER;
  bool has_pref = false;
  if (c->kinmnrl_num_prefactors) {
    for (int m = 0; m < c->nkinmnrl; m++) {
      const int np = c->kinmnrl_num_prefactors[m];
      if (np < 0 || np > PFRX_MAX_PREFACTORS) return set_err(PFRX_E_LIMIT, "more than 10 prefactors%s", "");
      has_pref = has_pref || np > 0;
    }
    if (has_pref && (!c->kinmnrl_pref_nspec || !c->kinmnrl_prefactor_id || !c->kinmnrl_pref_alpha ||
                     !c->kinmnrl_pref_beta || !c->kinmnrl_pref_atten_coef || !c->kinmnrl_pref_rate ||
                     !c->kinmnrl_pref_activation_energy))
      return set_err(PFRX_E_INVALID, "prefactor arrays missing%s", "");
  }
  const bool has_sorb2 = c->neqionxrxn > 0 || c->neqkdrxn > 0 || c->neqdynamickdrxn > 0;
  if (c->neqionxrxn > 0) {
    if (!c->eqionx_ptr || !c->eqionx_cationid || !c->eqionx_k || !c->eqionx_CEC || !c->eqionx_to_surf ||
        !c->eqionx_Z_flag)
      return set_err(PFRX_E_INVALID, "ion exchange tables missing%s", "");
    for (int r = 0; r < c->neqionxrxn; r++) {
      if (c->eqionx_ptr[r + 1] - c->eqionx_ptr[r] > PFRX_MAX_NCOMP || c->eqionx_ptr[r + 1] - c->eqionx_ptr[r] < 1)
        return set_err(PFRX_E_LIMIT, "ion exchange reaction with too many / no cations%s", "");
      if (c->eqionx_to_surf[r] >= c->nkinmnrl) return set_err(PFRX_E_INVALID, "ion exchange mineral id%s", "");
    }
  }
  if (c->neqkdrxn > 0 && (!c->eqkd_specid || !c->eqkd_type || !c->eqkd_mineral || !c->eqkd_coeff ||
                          !c->eqkd_langmuir_b || !c->eqkd_freundlich_n))
    return set_err(PFRX_E_INVALID, "KD isotherm tables missing%s", "");
  if (c->neqdynamickdrxn > 0 && (!c->eqdynamickd_specid || !c->eqdynamickd_refspecid || !c->eqdynamickd_refspechigh ||
                                 !c->eqdynamickd_low || !c->eqdynamickd_high || !c->eqdynamickd_power))
    return set_err(PFRX_E_INVALID, "dynamic KD tables missing%s", "");
  // general kinetic reactions, radioactive decay, immobile decay
  const bool has_kin3 =
      c->ngeneral_rxn > 0 || c->nradiodecay_rxn > 0 || c->nimmobile_decay_rxn > 0 || c->nmicrobial_rxn > 0;
  if (c->ngeneral_rxn < 0 || c->nradiodecay_rxn < 0 || c->nimmobile_decay_rxn < 0 || c->nmicrobial_rxn < 0)
    return set_err(PFRX_E_INVALID, "negative reaction count%s", "");
  if (c->nmicrobial_rxn > 0) {
    const int nr = c->nmicrobial_rxn;
    if (!c->microbial_ptr || !c->microbial_specid || !c->microbial_stoich || !c->microbial_rate_constant ||
        !c->microbial_monod_ptr || !c->microbial_inhibition_ptr || !c->microbial_biomassid ||
        !c->microbial_biomass_yield)
      return set_err(PFRX_E_INVALID, "microbial reaction tables missing%s", "");
    if (c->microbial_monod_ptr[nr] > 0 && (!c->microbial_monod_specid || !c->microbial_monod_K || !c->microbial_monod_Cth))
      return set_err(PFRX_E_INVALID, "microbial reaction Monod tables missing%s", "");
    if (c->microbial_inhibition_ptr[nr] > 0 && (!c->microbial_inhibition_specid || !c->microbial_inhibition_type ||
                                                !c->microbial_inhibition_C || !c->microbial_inhibition_C2))
      return set_err(PFRX_E_INVALID, "microbial reaction inhibition tables missing%s", "");
    if (c->microbial_concentration_units < PFRX_MICROBIAL_MOLALITY ||
        c->microbial_concentration_units > PFRX_MICROBIAL_MOLARITY)
      return set_err(PFRX_E_INVALID, "microbial_concentration_units%s", "");
    for (int k = 0; k < c->microbial_ptr[nr]; k++)
      if (c->microbial_specid[k] < 0 || c->microbial_specid[k] >= c->naqcomp)
        return set_err(PFRX_E_INVALID, "microbial reaction species id out of range%s", "");
    for (int k = 0; k < c->microbial_monod_ptr[nr]; k++)
      if (c->microbial_monod_specid[k] < 0 || c->microbial_monod_specid[k] >= c->naqcomp)
        return set_err(PFRX_E_INVALID, "microbial reaction species id out of range%s", "");
    for (int k = 0; k < c->microbial_inhibition_ptr[nr]; k++) {
      if (c->microbial_inhibition_specid[k] < 0 || c->microbial_inhibition_specid[k] >= c->naqcomp)
        return set_err(PFRX_E_INVALID, "microbial reaction species id out of range%s", "");
      const int t = c->microbial_inhibition_type[k];
      if (t != PFRX_INHIBITION_THRESHOLD && t != PFRX_INHIBITION_MONOD && t != PFRX_INHIBITION_INVERSE_MONOD &&
          t != PFRX_INHIBITION_SMOOTHSTEP)
        return set_err(PFRX_E_INVALID, "microbial reaction inhibition type%s", "");
    }
    for (int r = 0; r < nr; r++) {
      if (c->microbial_monod_ptr[r + 1] - c->microbial_monod_ptr[r] > PFRX_MAX_MONOD_TERMS ||
          c->microbial_inhibition_ptr[r + 1] - c->microbial_inhibition_ptr[r] > PFRX_MAX_MONOD_TERMS)
        return set_err(PFRX_E_LIMIT, "more than 8 Monod / inhibition terms in a microbial reaction%s", "");
      const int b = c->microbial_biomassid[r];
      if (b > c->naqcomp || -b > c->nimcomp) return set_err(PFRX_E_INVALID, "microbial biomass species id%s", "");
    }
  }
  if (c->ngeneral_rxn > 0) {
    if (!c->general_ptr || !c->general_specid || !c->general_stoich || !c->general_fwd_ptr || !c->general_bwd_ptr ||
        !c->general_kf || !c->general_kr)
      return set_err(PFRX_E_INVALID, "general reaction tables missing%s", "");
    // a side of a reaction may be empty ("D(aq) <->")
    if ((c->general_fwd_ptr[c->ngeneral_rxn] > 0 && (!c->general_fwd_specid || !c->general_fwd_stoich)) ||
        (c->general_bwd_ptr[c->ngeneral_rxn] > 0 && (!c->general_bwd_specid || !c->general_bwd_stoich)))
      return set_err(PFRX_E_INVALID, "general reaction tables missing%s", "");
    for (int k = 0; k < c->general_ptr[c->ngeneral_rxn]; k++)
      if (c->general_specid[k] < 0 || c->general_specid[k] >= c->naqcomp)
        return set_err(PFRX_E_INVALID, "general reaction species id out of range%s", "");
    for (int k = 0; k < c->general_fwd_ptr[c->ngeneral_rxn]; k++)
      if (c->general_fwd_specid[k] < 0 || c->general_fwd_specid[k] >= c->naqcomp)
        return set_err(PFRX_E_INVALID, "general reaction species id out of range%s", "");
    for (int k = 0; k < c->general_bwd_ptr[c->ngeneral_rxn]; k++)
      if (c->general_bwd_specid[k] < 0 || c->general_bwd_specid[k] >= c->naqcomp)
        return set_err(PFRX_E_INVALID, "general reaction species id out of range%s", "");
  }
  if (c->nradiodecay_rxn > 0) {
    if (!c->radiodecay_ptr || !c->radiodecay_specid || !c->radiodecay_stoich || !c->radiodecay_forward_specid ||
        !c->radiodecay_kf)
      return set_err(PFRX_E_INVALID, "radioactive decay tables missing%s", "");
    for (int k = 0; k < c->radiodecay_ptr[c->nradiodecay_rxn]; k++)
      if (c->radiodecay_specid[k] < 0 || c->radiodecay_specid[k] >= c->naqcomp)
        return set_err(PFRX_E_INVALID, "radioactive decay species id out of range%s", "");
    for (int r = 0; r < c->nradiodecay_rxn; r++)
      if (c->radiodecay_forward_specid[r] < 0 || c->radiodecay_forward_specid[r] >= c->naqcomp)
        return set_err(PFRX_E_INVALID, "radioactive decay parent id out of range%s", "");
    // the Jacobian of the sorbed inventory needs rt_auxvar%dtotal_sorb_eq as a matrix of its own
    // (DevCfg.need_ds); multirate sorption keeps its sorbed totals elsewhere (kinmr_total_sorb), like
    // in the reference, and does not take part
  }
  if (c->nimmobile_decay_rxn > 0) {
    if (!c->immobile_decay_specid || !c->immobile_decay_constant)
      return set_err(PFRX_E_INVALID, "immobile decay tables missing%s", "");
    for (int r = 0; r < c->nimmobile_decay_rxn; r++)
      if (c->immobile_decay_specid[r] < 0 || c->immobile_decay_specid[r] >= c->nimcomp)
        return set_err(PFRX_E_INVALID, "immobile decay species id out of range%s", "");
  }
  // ELM-CN sandboxes: what the CUDA path covers (everything else is refused, not approximated)
  const bool has_sbx3 = c->somdec || c->nitrif || c->denitr || c->plantn || c->langmuir || c->cndegas || c->calcite || c->radon;
  if (c->radon) {
    if (c->radon->species_id < 0 || c->radon->species_id >= c->naqcomp)
      return set_err(PFRX_E_INVALID, "RADON sandbox species id out of range%s", "");
    if (c->radon->mineral_id < 0 || c->radon->mineral_id >= c->nkinmnrl)
      return set_err(PFRX_E_INVALID, "RADON sandbox needs its mineral among the kinetic minerals%s", "");
  }
  if (c->nactive_gas < 0) return set_err(PFRX_E_INVALID, "negative nactive_gas%s", "");
  if (c->nactive_gas > 0) {
    if (!c->acteq_ptr || !c->acteq_specid || !c->acteq_stoich || !c->acteq_h2ostoich || !c->acteq_logK)
      return set_err(PFRX_E_INVALID, "active gas tables missing%s", "");
    for (int k = 0; k < c->acteq_ptr[c->nactive_gas]; k++)
      if (c->acteq_specid[k] < 0 || c->acteq_specid[k] >= c->naqcomp)
        return set_err(PFRX_E_INVALID, "active gas species id out of range%s", "");
  }
  if (c->calcite) {
    const pfrx_calcite_sandbox *cs = c->calcite;
    const int ids[3] = {cs->h_ion_id, cs->calcium_id, cs->bicarbonate_id};
    for (int k = 0; k < 3; k++)
      if (ids[k] < 0 || ids[k] >= c->naqcomp) return set_err(PFRX_E_INVALID, "CALCITE sandbox species id out of range%s", "");
    if (cs->mineral_id < 0 || cs->mineral_id >= c->nkinmnrl)
      return set_err(PFRX_E_INVALID, "CALCITE sandbox needs its mineral among the kinetic minerals%s", "");
    // "the RATE_CONSTANT in the default MINERAL_KINETICS block must be set to zero" (reaction_sandbox_calcite.F90:238-243)
    if (c->kinmnrl_rate_constant[cs->mineral_id] != 0.0)
      return set_err(PFRX_E_INVALID, "CALCITE sandbox: the mineral's RATE_CONSTANT must be zero%s", "");
  }
  if (c->cndegas) {
    const pfrx_cndegas *cd = c->cndegas;
    const int gid[3] = {cd->co2g_id, cd->n2og_id, cd->n2g_id}, aid[3] = {cd->co2a_id, cd->n2oa_id, cd->n2a_id};
    for (int k = 0; k < 3; k++)
      if ((aid[k] >= c->naqcomp) || (aid[k] >= 0 && gid[k] >= c->nimcomp))
        return set_err(PFRX_E_INVALID, "CNDEGAS species id out of range%s", "");
    if (cd->fixph_on && (cd->proton_id < 0 || cd->proton_id >= c->naqcomp || cd->himm_id < 0 || cd->himm_id >= c->nimcomp))
      return set_err(PFRX_E_INVALID, "CNDEGAS FIXPH needs H+ and the immobile species Himm%s", "");
    if (cd->cell_state_mode < 0 || cd->cell_state_mode > 2) return set_err(PFRX_E_INVALID, "CNDEGAS cell_state_mode%s", "");
  }
  if (c->plantn && (c->plantn->plantn_id < 0 || (c->plantn->nh4_id < 0 && c->plantn->no3_id < 0)))
    return set_err(PFRX_E_INVALID, "PLANTN needs PlantN and NH4+ or NO3-%s", "");
  if (c->langmuir && (c->langmuir->aq_id < 0 || c->langmuir->sorb_id < 0))
    return set_err(PFRX_E_INVALID, "LANGMUIR needs its aqueous and sorbed species%s", "");
  if (c->somdec) {
    const pfrx_somdec *sd = c->somdec;
    if (sd->nrxn < 1 || !sd->downstream_ptr || !sd->monod_ptr || !sd->inhib_ptr || !sd->upstream_c_id)
      return set_err(PFRX_E_INVALID, "pfrx_somdec tables missing%s", "");
    auto kind_ok = [](int t) { return t == PFRX_SPEC_AQUEOUS || t == PFRX_SPEC_IMMOBILE; };
    if (sd->co2_id < 0 || !kind_ok(sd->co2_itype) || (sd->o2_id >= 0 && !kind_ok(sd->o2_itype)))
      return set_err(PFRX_E_INVALID, "SOMDECOMP: CO2 / O2 must be primary or immobile species (gas not supported)%s", "");
    if (sd->nh4_id < 0) return set_err(PFRX_E_INVALID, "SOMDECOMP needs NH4+ as a primary species%s", "");
    for (int k = 0; k < sd->monod_ptr[sd->nrxn]; k++)
      if (!kind_ok(sd->monod_specitype[k])) return set_err(PFRX_E_INVALID, "SOMDECOMP MONOD on a gas species%s", "");
    for (int k = 0; k < sd->inhib_ptr[sd->nrxn]; k++)
      if (!kind_ok(sd->inhib_specitype[k])) return set_err(PFRX_E_INVALID, "SOMDECOMP INHIBITION on a gas species%s", "");
    for (int r = 0; r < sd->nrxn; r++)
      if (sd->ox_specid[r] >= 0 && !kind_ok(sd->ox_specitype[r]))
        return set_err(PFRX_E_INVALID, "SOMDECOMP Ox species must be primary or immobile%s", "");
  }
  if (c->nitrif && c->nitrif->nh4_id < 0) return set_err(PFRX_E_INVALID, "NITRIFICATION needs NH4+%s", "");
  if (c->denitr && c->denitr->no3_id < 0) return set_err(PFRX_E_INVALID, "DENITRIFICATION needs NO3-%s", "");
  if (c->sandbox_list)
    for (int k = 0; k < c->nsandbox; k++)
      if (c->sandbox_list[k] < PFRX_SANDBOX_CLM_CN || c->sandbox_list[k] > PFRX_SANDBOX_RADON ||
          c->nsandbox > PFRX_MAX_SANDBOXES)
        return set_err(PFRX_E_INVALID, "bad sandbox_list%s", "");
  int ndev = 0;
  CUDA_OK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return set_err(PFRX_E_CUDA, "no such CUDA device%s", "");
  CUDA_OK(cudaSetDevice(device));
  pfrx_handle *h = new pfrx_handle();
  h->device = device;
  h->n = n;
  DevCfg &d = h->cfg;
  memset(&d, 0, sizeof(d));
  d.naq = c->naqcomp;
  d.nim = c->nimcomp;
  d.n = n;
  d.use_full_geochemistry = c->use_full_geochemistry;
  d.use_log = c->use_log_formulation;
  d.use_total_as_guess = c->use_total_as_guess;
  d.use_isothermal = c->use_isothermal || !c->eqcplx_logKcoef;
  if (!c->use_isothermal && c->neqcplx > 0 && !c->eqcplx_logKcoef) {
    delete h;
    return set_err(PFRX_E_INVALID, "anisothermal run needs eqcplx_logKcoef%s", "");
  }
  if (c->neqcplx == 0) d.use_isothermal = c->use_isothermal;
  d.act_freq = c->act_coef_update_frequency;
  d.act_alg = c->act_coef_update_algorithm;
  d.use_act_h2o = c->use_activity_h2o;
  d.h2o_aq_id = c->h2o_aq_id;
  d.max_its = c->maximum_reaction_iterations;
  d.max_cuts = c->maximum_reaction_cuts;
  d.max_dlnC = c->max_dlnC_rreact;
  d.tol_relchange = c->max_relative_change_tolerance;
  d.tol_res = c->max_residual_tolerance;
  d.tol_relres = c->max_rel_residual_tolerance;
  d.min_sat = c->rt_min_saturation;
  h->sig = config_signature(c);
  h->dump_text = config_dump_text(c);
  h->pri_Z_host.assign(c->primary_spec_Z, c->primary_spec_Z + c->naqcomp);
  h->sr_flag_host.clear();
  if (c->nsrfcplxrxn > 0 && c->srfcplxrxn_stoich_flag)
    h->sr_flag_host.assign(c->srfcplxrxn_stoich_flag, c->srfcplxrxn_stoich_flag + c->nsrfcplxrxn);
  h->spec_prm.max_its = d.max_its;
  h->spec_prm.max_cuts = d.max_cuts;
  h->spec_prm.max_dlnC = d.max_dlnC;
  h->spec_prm.tol_relchange = d.tol_relchange;
  h->spec_prm.tol_res = d.tol_res;
  h->spec_prm.tol_relres = d.tol_relres;
  h->spec_prm.min_sat = d.min_sat;
  d.debyeA = c->debyeA;
  d.debyeB = c->debyeB;
  d.debyeBdot = c->debyeBdot;
  d.ncplx = c->neqcplx;
  d.nkin = c->nkinmnrl;
  d.nsrfrxn = c->nsrfcplxrxn;
  d.nsrfcplx = c->nsrfcplx;
  d.neqsr = c->neqsrfcplxrxn;
  d.nionx = c->neqionxrxn;
  d.nkd = c->neqkdrxn;
  d.ndynkd = c->neqdynamickdrxn;
  d.ikd_units = c->ikd_units;
  d.nsorb = d.neqsr + d.nionx + d.nkd + d.ndynkd;
  d.n_ixcat = c->neqionxrxn > 0 ? c->eqionx_ptr[c->neqionxrxn] : 0;
  d.nmr = c->nkinmrsrfcplxrxn;
  d.cn_nrxn = c->clmcn_nrxn;
  d.cn_C = c->clmcn_C_species_id;
  d.cn_N = c->clmcn_N_species_id;
  d.has_sd = c->somdec ? 1 : 0;
  d.has_nt = c->nitrif ? 1 : 0;
  d.has_dn = c->denitr ? 1 : 0;
  d.has_pn = c->plantn ? 1 : 0;
  d.has_lg = c->langmuir ? 1 : 0;
  d.has_cd = c->cndegas ? 1 : 0;
  d.has_cs = c->calcite ? 1 : 0;
  d.has_rn = c->radon ? 1 : 0;
  d.elm_flow = (c->elm_pflotran && c->elm_flow_coupled) ? 1 : 0;
  d.ngas = c->nactive_gas;
  d.elm = c->elm_pflotran ? 1 : 0;
  d.need_dt = (has_sbx3 || c->nradiodecay_rxn > 0) ? 1 : 0;
  d.need_ds = (c->nradiodecay_rxn > 0 && c->neqsrfcplxrxn + c->neqionxrxn + c->neqkdrxn + c->neqdynamickdrxn > 0) ? 1 : 0;
  d.ngen = c->ngeneral_rxn;
  d.nrd = c->nradiodecay_rxn;
  d.nidc = c->nimmobile_decay_rxn;
  d.nmb = c->nmicrobial_rxn;
  d.mb_units = c->microbial_concentration_units;
  d.n_nc = c->somdec ? c->somdec->nrxn + c->somdec->downstream_ptr[c->somdec->nrxn] : 0;
  {
    static const int def_order[9] = {PFRX_SANDBOX_CLM_CN, PFRX_SANDBOX_SOMDEC, PFRX_SANDBOX_NITRIF,
                                     PFRX_SANDBOX_DENITR, PFRX_SANDBOX_PLANTN, PFRX_SANDBOX_LANGMUIR,
                                     PFRX_SANDBOX_CNDEGAS, PFRX_SANDBOX_CALCITE, PFRX_SANDBOX_RADON};
    const int32_t *ord = c->sandbox_list ? c->sandbox_list : def_order;
    const int no = c->sandbox_list ? c->nsandbox : 9;
    d.nsbx = 0;
    for (int k = 0; k < no; k++) {
      const int kind = ord[k];
      const bool present = (kind == PFRX_SANDBOX_CLM_CN && c->clmcn_nrxn > 0) ||
                           (kind == PFRX_SANDBOX_SOMDEC && c->somdec) || (kind == PFRX_SANDBOX_NITRIF && c->nitrif) ||
                           (kind == PFRX_SANDBOX_DENITR && c->denitr) || (kind == PFRX_SANDBOX_PLANTN && c->plantn) ||
                           (kind == PFRX_SANDBOX_LANGMUIR && c->langmuir) || (kind == PFRX_SANDBOX_CNDEGAS && c->cndegas) ||
                           (kind == PFRX_SANDBOX_CALCITE && c->calcite) || (kind == PFRX_SANDBOX_RADON && c->radon);
      if (present) d.sbx[d.nsbx++] = kind;
    }
  }
  if (c->nitrif) d.nt = *c->nitrif;
  if (c->denitr) d.dn = *c->denitr;
  if (c->plantn) d.pn = *c->plantn;
  if (c->langmuir) d.lg = *c->langmuir;
  if (c->cndegas) d.cd = *c->cndegas;
  if (c->calcite) d.cs = *c->calcite;
  if (c->radon) d.rn = *c->radon;

  // kernel variant first: the task partition depends on the lane count
  {
    int want = 0;
    if (const char *ev = getenv("PFRX_LANES")) want = atoi(ev);
    int rc0 = pick_kernel(h, want);  // PFRX_TPC=1 selects the thread-per-cell kernel
    if (!rc0 && (has_pref || act_newton || has_sbx3 || has_sorb2 || has_kin3 || c->nactive_gas > 0) && !h->tpc) {
      // mineral prefactors, the iterated ionic strength and the SOMDECOMP / NITRIFICATION /
      // DENITRIFICATION sandboxes live in the thread-per-cell kernel only
      const KernelGetter *gt = nullptr;
      for (const auto &k : g_getters)
        if (k.n == h->npad) gt = &k;
      h->lanes = 1;
      h->tpc = true;
      h->kernel = gt->get(0);
    }
    if (rc0) {
      delete h;
      return rc0;
    }
  }
  const int L = h->lanes;
  Arena A;
  const int naq = c->naqcomp;
  // ---- activity classes: one Debye-Hueckel evaluation per distinct (Z, a0) ----
  std::vector<double> pri_Z2(naq), cls_negz2, cls_a0, cx_Z2(std::max(c->neqcplx, 0));
  std::vector<int> pri_cls(naq, -1), cx_cls(std::max(c->neqcplx, 0), -1);
  auto class_of = [&](double z, double a0) {
    if (!(fabs(z) > 1.e-10)) return -1;  // neutral: gamma = 1 (reaction.F90:4570)
    for (size_t q = 0; q < cls_a0.size(); q++)
      if (cls_negz2[q] == -z * z && cls_a0[q] == a0) return (int)q;
    cls_negz2.push_back(-z * z);
    cls_a0.push_back(a0);
    return (int)cls_a0.size() - 1;
  };
  for (int i = 0; i < naq; i++) {
    pri_Z2[i] = c->primary_spec_Z[i] * c->primary_spec_Z[i];
    pri_cls[i] = class_of(c->primary_spec_Z[i], c->primary_spec_a0[i]);
  }
  for (int k = 0; k < c->neqcplx; k++) {
    cx_Z2[k] = c->eqcplx_Z[k] * c->eqcplx_Z[k];
    cx_cls[k] = class_of(c->eqcplx_Z[k], c->eqcplx_a0[k]);
  }
  d.ncls = (int)cls_a0.size();
  A.add(pri_Z2.data(), pri_Z2.size(), &d.pri_Z2);
  A.add(pri_cls.data(), pri_cls.size(), &d.pri_cls);
  A.add(cls_negz2.data(), cls_negz2.size(), &d.cls_negz2);
  A.add(cls_a0.data(), cls_a0.size(), &d.cls_a0);
  // ---- balanced task lists: totals_i = sum_k nu_ki sec_k and
  //      S_ij = sum_k nu_ki nu_kj sec_k (i <= j, stored to both triangles).
  //      Destinations are offsets from ws.J: J(i,j) = i*js+j, totals = N*js+i,
  //      partial sums of entries cut by a lane boundary = N*js+naq+p. ------------
  const int Npad = h->npad, js = Npad | 1;
  std::vector<int2> tk_kd;
  std::vector<int> fx_ptr(1, 0), fx_dst, fx_dst2, fx_src;
  std::vector<double> tk_w;
  std::vector<unsigned> row_mask(naq, 0u);
  int npartial = 0;
  if (c->neqcplx > 0) {
    int nnz = c->eqcplx_ptr[c->neqcplx];
    A.add(c->eqcplx_ptr, c->neqcplx + 1, &d.cx_ptr);
    A.add(c->eqcplx_specid, nnz, &d.cx_id);
    A.add(c->eqcplx_stoich, nnz, &d.cx_st);
    A.add(c->eqcplx_h2ostoich, c->neqcplx, &d.cx_h2o);
    A.add(c->eqcplx_logK, c->neqcplx, &d.cx_logK);
    A.add(c->eqcplx_logKcoef, c->eqcplx_logKcoef ? 5 * c->neqcplx : 0, &d.cx_logKcoef);
    A.add(cx_Z2.data(), cx_Z2.size(), &d.cx_Z2);
    A.add(cx_cls.data(), cx_cls.size(), &d.cx_cls);
    struct Task {
      int key, d1, d2, k;  // key orders entries; d1/d2 destinations (d2 = -1: none)
      double w;
    };
    std::vector<Task> tasks;
    for (int k = 0; k < c->neqcplx; k++) {
      for (int p = c->eqcplx_ptr[k]; p < c->eqcplx_ptr[k + 1]; p++) {
        int i = c->eqcplx_specid[p];
        tasks.push_back({i, Npad * js + i, -1, k, c->eqcplx_stoich[p]});
        for (int p2 = c->eqcplx_ptr[k]; p2 < c->eqcplx_ptr[k + 1]; p2++) {
          int j = c->eqcplx_specid[p2];
          if (j < i) continue;
          row_mask[i] |= 1u << j;
          row_mask[j] |= 1u << i;
          tasks.push_back({naq + i * naq + j, i * js + j, (i == j) ? -1 : j * js + i, k,
                           c->eqcplx_stoich[p] * c->eqcplx_stoich[p2]});
        }
      }
    }
    std::stable_sort(tasks.begin(), tasks.end(), [](const Task &a, const Task &b) { return a.key < b.key; });
    const int T = (int)tasks.size();
    const int tpl = (T + L - 1) / L;
    d.tk_per_lane = tpl;
    tk_kd.assign((size_t)tpl * L, make_int2(0, 0));
    tk_w.assign((size_t)tpl * L, 0.0);
    int last_key = -1;
    for (int l = 0; l < L; l++) {
      int t0 = std::min(T, l * tpl), t1 = std::min(T, (l + 1) * tpl);
      for (int t = t0; t < t1; t++) {
        int e = tasks[t].key;
        bool seg_end = (t + 1 == t1) || tasks[t + 1].key != e;
        int d1 = -1, d2 = -1;
        if (seg_end) {
          int u = t;  // first task of this segment within the lane
          while (u > t0 && tasks[u - 1].key == e) u--;
          bool starts = (u == 0) || tasks[u - 1].key != e;
          bool ends = (t + 1 == T) || tasks[t + 1].key != e;
          if (starts && ends) {
            d1 = tasks[t].d1;
            d2 = tasks[t].d2;
          } else {
            d1 = Npad * js + naq + npartial;  // partial slot
            if (e != last_key) {
              fx_dst.push_back(tasks[t].d1);
              fx_dst2.push_back(tasks[t].d2);
              fx_ptr.push_back(fx_ptr.back());
              last_key = e;
            }
            fx_src.push_back(d1);
            fx_ptr.back() = (int)fx_src.size();
            npartial++;
          }
        }
        size_t ix = (size_t)(t - t0) * L + l;
        tk_kd[ix] = make_int2(tasks[t].k | ((d1 + 1) << 16), d2 + 1);
        tk_w[ix] = tasks[t].w;
      }
    }
  }
  d.nacc = naq + npartial;
  d.nfix = (int)fx_dst.size();
  A.add(tk_kd.data(), tk_kd.size(), &d.tk_kd);
  A.add(tk_w.data(), tk_w.size(), &d.tk_w);
  A.add(row_mask.data(), row_mask.size(), &d.row_mask);
  A.add(fx_ptr.data(), fx_ptr.size(), &d.fx_ptr);
  A.add(fx_dst.data(), fx_dst.size(), &d.fx_dst);
  A.add(fx_dst2.data(), fx_dst2.size(), &d.fx_dst2);
  A.add(fx_src.data(), fx_src.size(), &d.fx_src);
  if (c->nkinmnrl > 0) {
    int nk = c->nkinmnrl, nnz = c->kinmnrl_ptr[nk];
    A.add(c->kinmnrl_ptr, nk + 1, &d.mn_ptr);
    A.add(c->kinmnrl_specid, nnz, &d.mn_id);
    A.add(c->kinmnrl_stoich, nnz, &d.mn_st);
    A.add(c->kinmnrl_h2ostoich, nk, &d.mn_h2o);
    A.add(c->kinmnrl_logK, nk, &d.mn_logK);
    A.add(c->kinmnrl_logKcoef, c->kinmnrl_logKcoef ? 5 * nk : 0, &d.mn_logKcoef);
    A.add(c->kinmnrl_molar_vol, nk, &d.mn_vol);
    A.add(c->kinmnrl_rate_constant, nk, &d.mn_rate);
    A.add(c->kinmnrl_activation_energy, nk, &d.mn_eact);
    A.add(c->kinmnrl_affinity_threshold, nk, &d.mn_thresh);
    A.add(c->kinmnrl_rate_limiter, nk, &d.mn_limit);
    A.add(c->kinmnrl_irreversible, nk, &d.mn_irrev);
    A.add(c->kinmnrl_Temkin_const, c->kinmnrl_Temkin_const ? nk : 0, &d.mn_temkin);
    A.add(c->kinmnrl_min_scale_factor, c->kinmnrl_min_scale_factor ? nk : 0, &d.mn_scale);
    A.add(c->kinmnrl_affinity_power, c->kinmnrl_affinity_power ? nk : 0, &d.mn_power);
    {
      const int np = has_pref ? nk * PFRX_MAX_PREFACTORS : 0, ns = np * PFRX_MAX_PREFACTOR_SPECIES;
      A.add(c->kinmnrl_num_prefactors, has_pref ? nk : 0, &d.mn_npref);
      A.add(c->kinmnrl_pref_nspec, np, &d.mn_pref_nspec);
      A.add(c->kinmnrl_prefactor_id, ns, &d.mn_pref_id);
      A.add(c->kinmnrl_pref_alpha, ns, &d.mn_pref_alpha);
      A.add(c->kinmnrl_pref_beta, ns, &d.mn_pref_beta);
      A.add(c->kinmnrl_pref_atten_coef, ns, &d.mn_pref_atten);
      A.add(c->kinmnrl_pref_rate, np, &d.mn_pref_rate);
      A.add(c->kinmnrl_pref_activation_energy, np, &d.mn_pref_eact);
    }
    std::vector<int> me_ptr(nk + 1, 0), me_ij;
    std::vector<double> me_coef;
    for (int m = 0; m < nk; m++) {
      for (int p = c->kinmnrl_ptr[m]; p < c->kinmnrl_ptr[m + 1]; p++)
        for (int p2 = c->kinmnrl_ptr[m]; p2 < c->kinmnrl_ptr[m + 1]; p2++) {
          me_ij.push_back(c->kinmnrl_specid[p] | (c->kinmnrl_specid[p2] << 8));
          me_coef.push_back(c->kinmnrl_stoich[p] * c->kinmnrl_stoich[p2]);
        }
      me_ptr[m + 1] = (int)me_ij.size();
    }
    A.add(me_ptr.data(), me_ptr.size(), &d.me_ptr);
    A.add(me_ij.data(), me_ij.size(), &d.me_ij);
    A.add(me_coef.data(), me_coef.size(), &d.me_coef);
  }
  if (c->nsrfcplxrxn > 0) {
    int nr = c->nsrfcplxrxn, ns = c->nsrfcplx;
    A.add(c->srfcplxrxn_ptr, nr + 1, &d.sr_ptr);
    A.add(c->srfcplxrxn_to_complex, c->srfcplxrxn_ptr[nr], &d.sr_cx);
    A.add(c->srfcplxrxn_surf_type, nr, &d.sr_type);
    A.add(c->srfcplxrxn_to_surf, nr, &d.sr_surf);
    A.add(c->srfcplxrxn_stoich_flag, nr, &d.sr_flag);
    A.add(c->srfcplxrxn_site_density, nr, &d.sr_dens);
    A.add(c->srfcplx_ptr, ns + 1, &d.sc_ptr);
    A.add(c->srfcplx_specid, c->srfcplx_ptr[ns], &d.sc_id);
    A.add(c->srfcplx_stoich, c->srfcplx_ptr[ns], &d.sc_st);
    A.add(c->srfcplx_h2ostoich, ns, &d.sc_h2o);
    A.add(c->srfcplx_free_site_stoich, ns, &d.sc_fs);
    A.add(c->srfcplx_logK, ns, &d.sc_logK);
    A.add(c->srfcplx_logKcoef, c->srfcplx_logKcoef ? 5 * ns : 0, &d.sc_logKcoef);
    A.add(c->eqsrfcplxrxn_to_srfcplxrxn, c->neqsrfcplxrxn, &d.eqsr);
    std::vector<int> se_ptr(ns + 1, 0), se_pp;
    for (int k = 0; k < ns; k++) {
      for (int p = c->srfcplx_ptr[k]; p < c->srfcplx_ptr[k + 1]; p++)
        for (int p2 = c->srfcplx_ptr[k]; p2 < c->srfcplx_ptr[k + 1]; p2++) se_pp.push_back(p | (p2 << 16));
      se_ptr[k + 1] = (int)se_pp.size();
    }
    A.add(se_ptr.data(), se_ptr.size(), &d.se_ptr);
    A.add(se_pp.data(), se_pp.size(), &d.se_pp);
    if (c->nkinmrsrfcplxrxn > 0) {
      int nm = c->nkinmrsrfcplxrxn;
      A.add(c->kinmrsrfcplxrxn_to_srfcplxrxn, nm, &d.mr_rxn);
      A.add(c->kinmr_rate_ptr, nm + 1, &d.mr_ptr);
      A.add(c->kinmr_rate, c->kinmr_rate_ptr[nm], &d.mr_rate);
      A.add(c->kinmr_frac, c->kinmr_rate_ptr[nm], &d.mr_frac);
    }
  }
  if (c->clmcn_nrxn > 0) {
    int nx = c->clmcn_nrxn, np = c->clmcn_npool;
    A.add(c->clmcn_CN_ratio, np, &d.cn_CN);
    A.add(c->clmcn_pool_nspec, np, &d.cn_nspec);
    A.add(c->clmcn_pool_C_id, np, &d.cn_cid);
    A.add(c->clmcn_pool_N_id, np, &d.cn_nid);
    A.add(c->clmcn_upstream_pool_id, nx, &d.cn_up);
    A.add(c->clmcn_downstream_pool_id, nx, &d.cn_down);
    A.add(c->clmcn_rate_constant, nx, &d.cn_k);
    A.add(c->clmcn_respiration_fraction, nx, &d.cn_resp);
    A.add(c->clmcn_inhibition_constant, nx, &d.cn_inhib);
  }
  if (has_sorb2) {
    A.add(c->primary_spec_Z, naq, &d.pri_Z);
    if (c->neqionxrxn > 0) {
      const int nr = c->neqionxrxn, nn = c->eqionx_ptr[nr];
      A.add(c->eqionx_ptr, nr + 1, &d.ix_ptr);
      A.add(c->eqionx_cationid, nn, &d.ix_cat);
      A.add(c->eqionx_k, nn, &d.ix_k);
      A.add(c->eqionx_CEC, nr, &d.ix_cec);
      A.add(c->eqionx_to_surf, nr, &d.ix_surf);
      A.add(c->eqionx_Z_flag, nr, &d.ix_zflag);
    }
    if (c->neqkdrxn > 0) {
      const int nr = c->neqkdrxn;
      A.add(c->eqkd_specid, nr, &d.kd_spec);
      A.add(c->eqkd_type, nr, &d.kd_type);
      A.add(c->eqkd_mineral, nr, &d.kd_mnrl);
      A.add(c->eqkd_coeff, nr, &d.kd_coeff);
      A.add(c->eqkd_langmuir_b, nr, &d.kd_lb);
      A.add(c->eqkd_freundlich_n, nr, &d.kd_fn);
    }
    if (c->neqdynamickdrxn > 0) {
      const int nr = c->neqdynamickdrxn;
      A.add(c->eqdynamickd_specid, nr, &d.dk_spec);
      A.add(c->eqdynamickd_refspecid, nr, &d.dk_ref);
      A.add(c->eqdynamickd_refspechigh, nr, &d.dk_refhigh);
      A.add(c->eqdynamickd_low, nr, &d.dk_low);
      A.add(c->eqdynamickd_high, nr, &d.dk_high);
      A.add(c->eqdynamickd_power, nr, &d.dk_power);
    }
  }
  if (c->ngeneral_rxn > 0) {
    const int nr = c->ngeneral_rxn;
    A.add(c->general_ptr, nr + 1, &d.gn_ptr);
    A.add(c->general_specid, c->general_ptr[nr], &d.gn_id);
    A.add(c->general_stoich, c->general_ptr[nr], &d.gn_st);
    A.add(c->general_fwd_ptr, nr + 1, &d.gn_fptr);
    A.add(c->general_fwd_specid, c->general_fwd_ptr[nr], &d.gn_fid);
    A.add(c->general_fwd_stoich, c->general_fwd_ptr[nr], &d.gn_fst);
    A.add(c->general_bwd_ptr, nr + 1, &d.gn_bptr);
    A.add(c->general_bwd_specid, c->general_bwd_ptr[nr], &d.gn_bid);
    A.add(c->general_bwd_stoich, c->general_bwd_ptr[nr], &d.gn_bst);
    A.add(c->general_kf, nr, &d.gn_kf);
    A.add(c->general_kr, nr, &d.gn_kr);
  }
  if (c->nradiodecay_rxn > 0) {
    const int nr = c->nradiodecay_rxn;
    A.add(c->radiodecay_ptr, nr + 1, &d.rd_ptr);
    A.add(c->radiodecay_specid, c->radiodecay_ptr[nr], &d.rd_id);
    A.add(c->radiodecay_stoich, c->radiodecay_ptr[nr], &d.rd_st);
    A.add(c->radiodecay_forward_specid, nr, &d.rd_fwd);
    A.add(c->radiodecay_kf, nr, &d.rd_kf);
  }
  if (c->nmicrobial_rxn > 0) {
    const int nr = c->nmicrobial_rxn, ns = c->microbial_ptr[nr], nm = c->microbial_monod_ptr[nr],
              nh = c->microbial_inhibition_ptr[nr];
    A.add(c->microbial_ptr, nr + 1, &d.mb_ptr);
    A.add(c->microbial_specid, ns, &d.mb_id);
    A.add(c->microbial_stoich, ns, &d.mb_st);
    A.add(c->microbial_rate_constant, nr, &d.mb_k);
    A.add(c->microbial_activation_energy, c->microbial_activation_energy ? nr : 0, &d.mb_ea);
    A.add(c->microbial_monod_ptr, nr + 1, &d.mb_mptr);
    A.add(c->microbial_monod_specid, nm, &d.mb_mid);
    A.add(c->microbial_monod_K, nm, &d.mb_mK);
    A.add(c->microbial_monod_Cth, nm, &d.mb_mC);
    A.add(c->microbial_inhibition_ptr, nr + 1, &d.mb_hptr);
    A.add(c->microbial_inhibition_specid, nh, &d.mb_hid);
    A.add(c->microbial_inhibition_type, nh, &d.mb_htype);
    A.add(c->microbial_inhibition_C, nh, &d.mb_hC);
    A.add(c->microbial_inhibition_C2, nh, &d.mb_hC2);
    A.add(c->microbial_biomassid, nr, &d.mb_bio);
    A.add(c->microbial_biomass_yield, nr, &d.mb_yield);
  }
  if (c->nactive_gas > 0) {
    const int ng = c->nactive_gas, nnz = c->acteq_ptr[ng];
    A.add(c->acteq_ptr, ng + 1, &d.gs_ptr);
    A.add(c->acteq_specid, nnz, &d.gs_id);
    A.add(c->acteq_stoich, nnz, &d.gs_st);
    A.add(c->acteq_h2ostoich, ng, &d.gs_h2o);
    A.add(c->acteq_logK, ng, &d.gs_logK);
    A.add(c->acteq_logK_coef, (c->acteq_logK_coef && !c->use_isothermal) ? 5 * ng : 0, &d.gs_logKcoef);
  }
  if (c->nimmobile_decay_rxn > 0) {
    A.add(c->immobile_decay_specid, c->nimmobile_decay_rxn, &d.idc_id);
    A.add(c->immobile_decay_constant, c->nimmobile_decay_rxn, &d.idc_k);
  }
  if (c->somdec) {
    const pfrx_somdec *sd = c->somdec;
    d.sd = *sd;  // scalars; every pointer below is re-homed into the arena
    const int nx = sd->nrxn, nd = sd->downstream_ptr[nx], nm = sd->monod_ptr[nx], ni = sd->inhib_ptr[nx];
#define SD_ADD(field, count) A.add(sd->field, (size_t)(count), &d.sd.field)
    SD_ADD(rate_constant, nx);
    SD_ADD(rate_decomposition, nx);
    SD_ADD(rate_ad_factor, nx);
    SD_ADD(upstream_c_id, nx);
    SD_ADD(upstream_n_id, nx);
    SD_ADD(upstream_is_aqueous, nx);
    SD_ADD(upstream_hr_id, nx);
    SD_ADD(upstream_nmin_id, nx);
    SD_ADD(upstream_nimp_id, nx);
    SD_ADD(upstream_nimm_id, nx);
    SD_ADD(upstream_nc, nx);
    SD_ADD(mineral_c_stoich, nx);
    SD_ADD(mineral_n_stoich, nx);
    SD_ADD(downstream_ptr, nx + 1);
    SD_ADD(downstream_c_id, nd);
    SD_ADD(downstream_n_id, nd);
    SD_ADD(downstream_is_aqueous, nd);
    SD_ADD(downstream_stoich, nd);
    SD_ADD(downstream_nc, nd);
    SD_ADD(temperature_response_function, nx);
    SD_ADD(moisture_response_function, nx);
    SD_ADD(ox_response_function, nx);
    SD_ADD(q10, nx);
    SD_ADD(ea, nx);
    SD_ADD(ox_half_saturation, nx);
    SD_ADD(decomp_depth_efolding, nx);
    SD_ADD(ox_specid, nx);
    SD_ADD(ox_specitype, nx);
    SD_ADD(monod_ptr, nx + 1);
    SD_ADD(monod_specid, nm);
    SD_ADD(monod_specitype, nm);
    SD_ADD(monod_pool_normalized, nm);
    SD_ADD(monod_half_saturation, nm);
    SD_ADD(monod_threshold, nm);
    SD_ADD(inhib_ptr, nx + 1);
    SD_ADD(inhib_itype, ni);
    SD_ADD(inhib_specid, ni);
    SD_ADD(inhib_specitype, ni);
    SD_ADD(inhib_constant, ni);
    SD_ADD(inhib_constant2, ni);
#undef SD_ADD
  }
  size_t asz = std::max<size_t>(A.bytes.size(), 16);
  cudaError_t e = cudaMalloc(&h->arena, asz);
  if (e != cudaSuccess) {
    delete h;
    return set_err(PFRX_E_CUDA, "cudaMalloc(tables): %s", cudaGetErrorString(e));
  }
  if (!A.bytes.empty()) {
    e = cudaMemcpy(h->arena, A.bytes.data(), A.bytes.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      cudaFree(h->arena);
      delete h;
      return set_err(PFRX_E_CUDA, "cudaMemcpy(tables): %s", cudaGetErrorString(e));
    }
  }
  for (auto &f : A.fix) *f.second = (unsigned char *)h->arena + f.first;

  h->rows_d.resize(PFRX_NUM_D);
  field_rows(c, h->rows_d.data());

  int rc = 0;
  if (const char *ev = getenv("PFRX_THREADS")) {
    int t = atoi(ev);
    if (t >= 32 && t <= 128 && (t % 32) == 0) h->threads = t;
  }
  rc = layout_and_launch_params(h);
  if (rc) {
    cudaFree(h->arena);
    delete h;
    return rc;
  }
  // nothing half-built is handed out: a handle whose streams or summary buffers are missing would
  // fault at its first step
  e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaMalloc(&h->d_summ, sizeof(DevSummary));
  if (e == cudaSuccess) e = cudaMallocHost(&h->h_summ, sizeof(DevSummary));
  if (e != cudaSuccess) {
    pfrx_destroy(h);
    return set_err(PFRX_E_CUDA, "pfrx_create (streams / summary buffers): %s", cudaGetErrorString(e));
  }
  *out = h;
  return PFRX_OK;
}

extern "C" void pfrx_destroy(pfrx_handle *h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->spec_module && g_drv.ModuleUnload) g_drv.ModuleUnload(h->spec_module);
  if (h->own) cudaFree(h->own);
  if (h->arena) cudaFree(h->arena);
  if (h->d_summ) cudaFree(h->d_summ);
  if (h->h_summ) cudaFreeHost(h->h_summ);
  if (h->ev_t0) cudaEventDestroy(h->ev_t0);
  if (h->ev_t1) cudaEventDestroy(h->ev_t1);
  if (h->d_red) cudaFree(h->d_red);
  if (h->h_red) cudaFreeHost(h->h_red);
  if (h->d_red_step) cudaFree(h->d_red_step);
  if (h->h_red_step) cudaFreeHost(h->h_red_step);
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->d_order) cudaFree(h->d_order);
  if (h->d_iota) cudaFree(h->d_iota);
  if (h->d_keys_out) cudaFree(h->d_keys_out);
  if (h->d_sort_tmp) cudaFree(h->d_sort_tmp);
  if (h->os_a) cudaFree(h->os_a);
  if (h->os_b) cudaFree(h->os_b);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->stream2) cudaStreamDestroy(h->stream2);
  if (h->ev_reset) cudaEventDestroy(h->ev_reset);
  if (h->out_stream) cudaStreamDestroy(h->out_stream);
  if (h->ev_ready)
    for (int i = 0; i < PFRX_MAX_CHUNKS; i++) {
      cudaEventDestroy(h->ev_in[i]);
      cudaEventDestroy(h->ev_k[i]);
    }
  delete h;
}

static int to_dev_state(const pfrx_handle *h, const pfrx_state *s, DevState *d) {
  const std::vector<int> &r = h->rows_d;
  d->ld = s->ld;
  d->total = s->total;
  d->pri_molal = s->pri_molal;
  d->immobile = s->immobile;
  d->pri_act_coef = s->pri_act_coef;
  d->sec_act_coef = s->sec_act_coef;
  d->sec_molal = s->sec_molal;
  d->ln_act_h2o = s->ln_act_h2o;
  d->mnrl_volfrac = s->mnrl_volfrac;
  d->mnrl_area = s->mnrl_area;
  d->mnrl_rate = s->mnrl_rate;
  d->free_site = s->srfcplxrxn_free_site_conc;
  d->eqsrfcplx_conc = s->eqsrfcplx_conc;
  d->total_sorb_eq = s->total_sorb_eq;
  d->kinmr = s->kinmr_total_sorb;
  d->den_kg = s->den_kg;
  d->sat = s->sat;
  d->temp = s->temp;
  d->porosity = s->porosity;
  d->volume = s->volume;
  d->soil_particle_density = s->soil_particle_density;
  d->imat = s->imat;
  d->num_sub_steps = s->num_sub_steps;
  d->num_iterations = s->num_iterations;
  d->num_kinetic_state_updates = s->num_kinetic_state_updates;
  d->ierror = s->ierror;
  d->elm_w = s->elm_w_scalar;
  d->elm_o = s->elm_o_scalar;
  d->elm_t = s->elm_t_scalar;
  d->elm_zsoil = s->elm_zsoil;
  d->elm_kscalar = s->elm_kscalar_decomp_c;
  d->elm_bd_dry = s->elm_bulkdensity_dry;
  d->elm_bsw = s->elm_bsw;
  d->somdec_nc = s->somdec_nc;
  d->elm_plantndemand = s->elm_rate_plantndemand;
  d->eqionx_ref = s->eqionx_ref_cation_sorbed_conc;
  d->eqionx_conc = s->eqionx_conc;
  d->pres = s->pres;
  d->sandbox_aux = s->sandbox_aux;
  d->sat_gas = s->sat_gas;
  d->total_gas = s->total_gas;
  d->gas_pp = s->gas_pp;
  d->elm_sucsat = s->elm_sucsat;
  d->elm_watfc = s->elm_watfc;
  d->elm_effpor = s->elm_effporosity;
  // required pointers
  const void *req[] = {r[0] ? s->total : (void *)1,
                       r[1] ? s->pri_molal : (void *)1,
                       r[2] ? s->immobile : (void *)1,
                       r[3] ? s->pri_act_coef : (void *)1,
                       r[4] ? s->sec_act_coef : (void *)1,
                       r[5] ? s->sec_molal : (void *)1,
                       r[7] ? s->mnrl_volfrac : (void *)1,
                       r[8] ? s->mnrl_area : (void *)1,
                       r[9] ? s->mnrl_rate : (void *)1,
                       r[10] ? s->srfcplxrxn_free_site_conc : (void *)1,
                       r[12] ? s->total_sorb_eq : (void *)1,
                       r[13] ? s->kinmr_total_sorb : (void *)1,
                       s->den_kg,
                       s->sat,
                       s->temp,
                       s->porosity,
                       s->volume,
                       s->num_sub_steps,
                       s->num_iterations,
                       s->num_kinetic_state_updates,
                       s->ierror};
  for (const void *p : req)
    if (!p) return set_err(PFRX_E_INVALID, "a required pfrx_state pointer is NULL%s", "");
  return PFRX_OK;
}

extern "C" int pfrx_bind_state(pfrx_handle *h, int64_t ncell, const pfrx_state *dev) {
  if (!h || !dev || ncell < 0 || dev->ld < ncell) return set_err(PFRX_E_INVALID, "bad bind_state arguments%s", "");
  if (ncell == 0) {
    // a rank that owns no cells: nothing to point at, every call on the shard is a no-op
    h->st = DevState();
    h->ncell = 0;
    h->bound = true;
    return PFRX_OK;
  }
  int rc = to_dev_state(h, dev, &h->st);
  if (rc) return rc;
  h->ncell = ncell;
  h->bound = true;
  h->os_inactive = -1;
  return PFRX_OK;
}

static int summary_reset(pfrx_handle *h, cudaStream_t s) {
  DevSummary z;
  memset(&z, 0, sizeof(z));
  z.first_failed = LLONG_MAX;
  *h->h_summ = z;
  CUDA_OK(cudaMemcpyAsync(h->d_summ, h->h_summ, sizeof(DevSummary), cudaMemcpyHostToDevice, s));
  return PFRX_OK;
}

// one kernel launch over cells [0, ncell) of `st`.  Chunked callers pass a view that starts at
// their first cell; the chunk-local first-failed-cell index is turned into a shard-local one on the
// host when the chunks are merged
__global__ void pfrx_iota_kernel(int *v, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) v[i] = (int)i;
}

// order of the NEXT launch on this shard: cells by the Newton iterations of the launch just enqueued, descending
// (stable: equal counts keep index order).  A ragged workload ends with the few cells that cut their step dozens
// of times; handed out first, their serial chains overlap the bulk instead of trailing it.
static int build_cell_order(pfrx_handle *h, const DevState &st, int64_t ncell, cudaStream_t s) {
  if (ncell > h->order_cap) {
    if (h->d_order) cudaFree(h->d_order);
    if (h->d_iota) cudaFree(h->d_iota);
    if (h->d_keys_out) cudaFree(h->d_keys_out);
    if (h->d_sort_tmp) cudaFree(h->d_sort_tmp);
    h->d_order = h->d_iota = h->d_keys_out = nullptr;
    h->d_sort_tmp = nullptr;
    h->order_cap = 0;
    h->order_valid = false;
    CUDA_OK(cudaMalloc(&h->d_order, ncell * sizeof(int)));
    CUDA_OK(cudaMalloc(&h->d_iota, ncell * sizeof(int)));
    CUDA_OK(cudaMalloc(&h->d_keys_out, ncell * sizeof(int)));
    size_t bytes = 0;
    CUDA_OK(cub::DeviceRadixSort::SortPairsDescending(nullptr, bytes, (const int *)nullptr, (int *)nullptr,
                                                      (const int *)nullptr, (int *)nullptr, (int)ncell, 0, 31, s));
    CUDA_OK(cudaMalloc(&h->d_sort_tmp, std::max<size_t>(bytes, 16)));
    h->sort_tmp_bytes = bytes;
    h->order_cap = ncell;
    pfrx_iota_kernel<<<256, 256, 0, s>>>(h->d_iota, ncell);
    CUDA_OK(cudaGetLastError());
  }
  size_t bytes = h->sort_tmp_bytes;
  CUDA_OK(cub::DeviceRadixSort::SortPairsDescending(h->d_sort_tmp, bytes, (const int *)st.num_iterations, h->d_keys_out,
                                                    (const int *)h->d_iota, h->d_order, (int)ncell, 0, 31, s));
  h->order_valid = true;
  h->order_key = st.num_iterations;
  h->order_ncell = ncell;
  return PFRX_OK;
}

static int launch_kernel(pfrx_handle *h, const DevState &st, int64_t ncell, double tran_dt, cudaStream_t s) {
  if (ncell <= 0) return PFRX_OK;
  int cpw = 32 / h->lanes;
  int wpb = h->threads / 32;
  int64_t need = (ncell + (int64_t)cpw * wpb - 1) / ((int64_t)cpw * wpb);
  if (h->tpc) need = (ncell + h->threads - 1) / h->threads;
  int64_t cap = (int64_t)h->sm_count * h->blocks_per_sm;
  if (h->spec_func) {
    need = (ncell + h->spec_cells - 1) / h->spec_cells;
    cap = (int64_t)h->sm_count * h->spec_blocks_per_sm;
    int grid = (int)std::max<int64_t>(1, std::min<int64_t>(need, cap));
    DevState st_arg = st;
    const bool ordered = h->spec_refill && h->order_mode && ncell >= 65536 && ncell < (int64_t)INT_MAX &&
                         (ncell == h->ncell || ncell == h->own_ncell);
    st_arg.order = (ordered && h->order_valid && h->order_key == st.num_iterations && h->order_ncell == ncell)
                       ? h->d_order
                       : nullptr;
    long long n_arg = ncell;
    double dt_arg = tran_dt;
    SpecParams prm = h->spec_prm;
    DevSummary *summ = h->d_summ;
    void *args[] = {&st_arg, &n_arg, &dt_arg, &prm, &summ};
    // work counter of the refill skeleton (pfrx_spec.cuh, SPEC_REFILL): per launch, not per step
    CUDA_OK(cudaMemsetAsync(&h->d_summ->next_cell, 0, sizeof(unsigned long long), s));
    DRV_OK(g_drv.LaunchKernel(h->spec_func, (unsigned)grid, 1, 1, (unsigned)h->spec_threads, 1, 1,
                              (unsigned)h->spec_smem, s, args, nullptr));
    h->launches++;
    if (ordered) return build_cell_order(h, st, ncell, s);
    return PFRX_OK;
  }
  int grid = (int)std::min<int64_t>(need, cap);
  if (grid < 1) grid = 1;
  h->kernel<<<grid, h->threads, h->smem_bytes, s>>>(h->cfg, st, ncell, tran_dt, h->d_summ);
  CUDA_OK(cudaGetLastError());
  h->launches++;
  return PFRX_OK;
}

static int launch(pfrx_handle *h, const DevState &st, int64_t ncell, double tran_dt, cudaStream_t s) {
  int rc = summary_reset(h, s);
  if (rc) return rc;
  rc = launch_kernel(h, st, ncell, tran_dt, s);
  if (rc) return rc;
  CUDA_OK(cudaMemcpyAsync(h->h_summ, h->d_summ, sizeof(DevSummary), cudaMemcpyDeviceToHost, s));
  h->pending = true;
  return PFRX_OK;
}

static void summary_out(const pfrx_handle *h, pfrx_step_result *out) {
  const DevSummary &s = *h->h_summ;
  out->ncell_active = (int64_t)s.ncell_active;
  out->sum_newton_iterations = (int64_t)s.sum_its;
  out->max_newton_iterations = s.max_its;
  out->max_num_kinetic_state_updates = s.max_kin;
  out->rstep_error = s.max_err;
  out->max_sub_steps = s.max_sub;
  out->num_cut_cells = (int64_t)s.num_cut_cells;
  out->first_failed_cell = s.first_failed == LLONG_MAX ? -1 : s.first_failed;
}

// shard summary -> the eight int64 of the step's collective (SUM of counts, MAX of maxima)
__global__ void pfrx_pack_summary_kernel(const DevSummary *s, long long *red) {
  red[0] = (long long)s->ncell_active;
  red[1] = (long long)s->sum_its;
  red[2] = (long long)s->num_cut_cells;
  red[3] = 0;
  red[4] = s->max_its;
  red[5] = s->max_kin;
  red[6] = s->max_err;
  red[7] = s->max_sub;
}

extern "C" int pfrx_rstep_async(pfrx_handle *h, double tran_dt) {
  if (!h) return set_err(PFRX_E_INVALID, "null handle%s", "");
  if (!h->bound) return set_err(PFRX_E_NOTBOUND, "pfrx_bind_state has not been called%s", "");
  CUDA_OK(cudaSetDevice(h->device));
  int rc = launch(h, h->st, h->ncell, tran_dt, h->stream);
  if (rc) return rc;
  h->red_inflight = false;
  h->red_valid = false;
  if (h->comm && h->d_red_step) {
    // the step's one exchange (MPI_Allreduce(rstep_error, MAX), pmc_subsurface_osrt.F90:381) on the
    // kernel stream, straight from the device summary: no host staging, no extra synchronisation
    pfrx_pack_summary_kernel<<<1, 1, 0, h->stream>>>(h->d_summ, h->d_red_step);
    CUDA_OK(cudaGetLastError());
    const int ncclInt64 = 4, ncclSum = 0, ncclMax = 2;
    g_nccl.GroupStart();
    int e1 = g_nccl.AllReduce(h->d_red_step, h->d_red_step, 4, ncclInt64, ncclSum, h->comm, h->stream);
    int e2 = g_nccl.AllReduce(h->d_red_step + 4, h->d_red_step + 4, 4, ncclInt64, ncclMax, h->comm, h->stream);
    int e3 = g_nccl.GroupEnd();
    if (e1 || e2 || e3)
      return set_err(PFRX_E_NCCL, "ncclAllReduce: %s", g_nccl.GetErrorString(e1 ? e1 : (e2 ? e2 : e3)));
    CUDA_OK(cudaMemcpyAsync(h->h_red_step, h->d_red_step, 8 * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
    h->red_inflight = true;
  }
  return PFRX_OK;
}

extern "C" int pfrx_rstep_finish(pfrx_handle *h, pfrx_step_result *out) {
  if (!h || !out) return set_err(PFRX_E_INVALID, "null argument%s", "");
  if (!h->pending) return set_err(PFRX_E_INVALID, "no step in flight%s", "");
  CUDA_OK(cudaStreamSynchronize(h->stream));
  h->pending = false;
  summary_out(h, out);
  if (h->red_inflight) {
    h->red_inflight = false;
    h->red_valid = true;
    h->red_local = *out;
  }
  return PFRX_OK;
}

extern "C" int pfrx_rstep(pfrx_handle *h, double tran_dt, pfrx_step_result *out) {
  int rc = pfrx_rstep_async(h, tran_dt);
  if (rc) return rc;
  return pfrx_rstep_finish(h, out);
}

// ---- batched RReaction / RReactionDerivative for the GIRT / ELM caller ----------
extern "C" int pfrx_reaction(pfrx_handle *h, double tran_dt, int want_jacobian, double *res, double *jac) {
  if (!h || !res || (want_jacobian && !jac)) return set_err(PFRX_E_INVALID, "null argument%s", "");
  if (!(tran_dt > 0.0)) return set_err(PFRX_E_INVALID, "tran_dt must be positive%s", "");
  if (!h->bound) return set_err(PFRX_E_NOTBOUND, "pfrx_bind_state has not been called%s", "");
  CUDA_OK(cudaSetDevice(h->device));
  if (!h->rx_kernel) {
    const KernelGetter *gt = nullptr;
    for (const auto &k : g_getters)
      if (k.n == h->npad) gt = &k;
    if (!gt) return set_err(PFRX_E_LIMIT, "no kernel variant for this size%s", "");
    h->rx_cfg = h->cfg;
    tpc_layout(h->rx_cfg, h->npad);
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, h->device));
    size_t per_thread = (size_t)h->rx_cfg.ws_stride * sizeof(double);
    int t = 128;
    while (t > 32 && per_thread * t > (size_t)prop.sharedMemPerBlockOptin) t /= 2;
    if (per_thread * t > (size_t)prop.sharedMemPerBlockOptin)
      return set_err(PFRX_E_LIMIT, "per-cell workspace does not fit shared memory%s", "");
    pfrx_reaction_fn fn = gt->get_rx();
    CUDA_OK(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(per_thread * t)));
    int nb = 0;
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void *)fn, t, per_thread * t));
    if (nb < 1) return set_err(PFRX_E_LIMIT, "kernel does not fit on an SM%s", "");
    h->rx_threads = t;
    h->rx_smem = per_thread * t;
    h->rx_blocks_per_sm = nb;
    h->rx_kernel = fn;
  }
  if (h->ncell <= 0) return PFRX_OK;
  int64_t need = (h->ncell + h->rx_threads - 1) / h->rx_threads;
  int grid = (int)std::max<int64_t>(1, std::min<int64_t>(need, (int64_t)h->sm_count * h->rx_blocks_per_sm));
  h->rx_kernel<<<grid, h->rx_threads, h->rx_smem, h->stream>>>(h->rx_cfg, h->st, h->ncell, want_jacobian, res, jac,
                                                               tran_dt);
  CUDA_OK(cudaGetLastError());
  h->launches++;
  CUDA_OK(cudaStreamSynchronize(h->stream));
  return PFRX_OK;
}

// ---- order of the kinetic-sorption checkpoint vectors (reactive_transport.F90:3968-4060) ----
extern "C" int pfrx_kinmr_checkpoint_rows(const pfrx_config *c, int32_t *rows, int32_t *nrows) {
  if (!c || !nrows) return set_err(PFRX_E_INVALID, "null argument%s", "");
  const int naq = c->naqcomp, nmr = c->nkinmrsrfcplxrxn;
  if (naq < 0 || naq > PFRX_MAX_NCOMP) return set_err(PFRX_E_LIMIT, "ncomp out of range%s", "");
  int n = 0;
  if (nmr > 0) {
    if (!c->kinmrsrfcplxrxn_to_srfcplxrxn || !c->kinmr_rate_ptr || !c->srfcplxrxn_ptr || !c->srfcplxrxn_to_complex ||
        !c->srfcplx_ptr || !c->srfcplx_specid)
      return set_err(PFRX_E_INVALID, "multirate surface complexation tables missing%s", "");
    std::vector<char> flag(naq, 0);
    for (int q = 0; q < nmr; q++) {
      const int irxn = c->kinmrsrfcplxrxn_to_srfcplxrxn[q];
      for (int j = c->srfcplxrxn_ptr[irxn]; j < c->srfcplxrxn_ptr[irxn + 1]; j++) {
        const int icplx = c->srfcplxrxn_to_complex[j];
        for (int p = c->srfcplx_ptr[icplx]; p < c->srfcplx_ptr[icplx + 1]; p++) {
          const int icomp = c->srfcplx_specid[p];
          if (icomp < 0 || icomp >= naq) return set_err(PFRX_E_INVALID, "srfcplx_specid out of range%s", "");
          flag[icomp] = 1;
        }
      }
    }
    for (int icomp = 0; icomp < naq; icomp++) {
      if (!flag[icomp]) continue;
      for (int q = 0; q < nmr; q++) {
        const int r0 = c->kinmr_rate_ptr[q], r1 = c->kinmr_rate_ptr[q + 1];
        for (int irate = 1; irate <= r1 - r0; irate++) {
          if (rows) rows[n] = naq * (r0 + q + irate) + icomp;
          n++;
        }
      }
    }
  }
  *nrows = n;
  return PFRX_OK;
}

// launch shape of a thread-per-cell set-up kernel (workspace in shared memory)
static int tpc_launch_shape(pfrx_handle *h, const void *fn, DevCfg *cfg, int *threads, size_t *smem, int *blocks_per_sm) {
  *cfg = h->cfg;
  tpc_layout(*cfg, h->npad);
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, h->device));
  const size_t per_thread = (size_t)cfg->ws_stride * sizeof(double);
  int t = 128;
  while (t > 32 && per_thread * t > (size_t)prop.sharedMemPerBlockOptin) t /= 2;
  if (per_thread * t > (size_t)prop.sharedMemPerBlockOptin)
    return set_err(PFRX_E_LIMIT, "per-cell workspace does not fit shared memory%s", "");
  CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(per_thread * t)));
  int nb = 0;
  CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, t, per_thread * t));
  if (nb < 1) return set_err(PFRX_E_LIMIT, "kernel does not fit on an SM%s", "");
  *threads = t;
  *smem = per_thread * t;
  *blocks_per_sm = nb;
  return PFRX_OK;
}

// ---- RTUpdateAuxVars (reactive_transport.F90:3525-3660) over the bound state ----
extern "C" int pfrx_update_auxvars(pfrx_handle *h, const double *tran_xx, int update_activity_coefs) {
  if (!h) return set_err(PFRX_E_INVALID, "null argument%s", "");
  if (!h->bound) return set_err(PFRX_E_NOTBOUND, "pfrx_bind_state has not been called%s", "");
  CUDA_OK(cudaSetDevice(h->device));
  const KernelGetter *gt = nullptr;
  for (const auto &g : g_getters)
    if (g.n == h->npad) gt = &g;
  if (!gt) return set_err(PFRX_E_LIMIT, "no kernel variant for this size%s", "");
  pfrx_auxvars_fn fn = gt->get_aux();
  DevCfg cfg;
  int t = 0, nb = 0;
  size_t smem = 0;
  int rc = tpc_launch_shape(h, (const void *)fn, &cfg, &t, &smem, &nb);
  if (rc) return rc;
  if (h->ncell <= 0) return PFRX_OK;
  int64_t need = (h->ncell + t - 1) / t;
  int grid = (int)std::max<int64_t>(1, std::min<int64_t>(need, (int64_t)h->sm_count * nb));
  fn<<<grid, t, smem, h->stream>>>(cfg, h->st, h->ncell, tran_xx, update_activity_coefs);
  CUDA_OK(cudaGetLastError());
  h->launches++;
  CUDA_OK(cudaStreamSynchronize(h->stream));
  return PFRX_OK;
}

// ---- batched ReactionEquilibrateConstraint (reaction.F90:1328-2117) ----
extern "C" int pfrx_equilibrate_constraint(pfrx_handle *h, const pfrx_constraint *k, const double *conc,
                                           int32_t *num_iterations, int32_t *ierror) {
  if (!h || !k || !conc || !k->type) return set_err(PFRX_E_INVALID, "null argument%s", "");
  if (!h->bound) return set_err(PFRX_E_NOTBOUND, "pfrx_bind_state has not been called%s", "");
  const DevCfg &d = h->cfg;
  const int naq = d.naq;
  if (k->naqcomp != naq) return set_err(PFRX_E_INVALID, "pfrx_constraint.naqcomp differs from the configuration%s", "");
  if (naq < 1) return set_err(PFRX_E_INVALID, "no aqueous species to equilibrate%s", "");
  if (d.act_freq != PFRX_ACT_COEF_FREQUENCY_OFF && d.act_alg == PFRX_ACT_COEF_ALGORITHM_NEWTON)
    return set_err(PFRX_E_INVALID, "constraint equilibration with the NEWTON activity algorithm is not built%s", "");
  bool any_eq = false;
  for (int i = 0; i < naq; i++) {
    switch (k->type[i]) {
      case PFRX_CONSTRAINT_NULL:
      case PFRX_CONSTRAINT_FREE:
      case PFRX_CONSTRAINT_TOTAL:
      case PFRX_CONSTRAINT_LOG:
      case PFRX_CONSTRAINT_CHARGE_BAL: break;
      case PFRX_CONSTRAINT_PH:
        // the reference allows pH on H+ (or on OH-, or with H+ as a complex): only the first is built
        if (h->pri_Z_host.size() != (size_t)naq || h->pri_Z_host[i] != 1.0)
          return set_err(PFRX_E_INVALID, "a pH constraint must be on the primary species H+%s", "");
        break;
      case PFRX_CONSTRAINT_MINERAL:
      case PFRX_CONSTRAINT_GAS: any_eq = true; break;
      default: return set_err(PFRX_E_INVALID, "constraint type not built (PE, EH, TOTAL_SORB, SUPERCRIT_CO2, TOTAL_AQ_PLUS_SORB)%s", "");
    }
  }
  if (any_eq) {
    if (!k->eq_logK || !k->eq_h2o_stoich || !k->eq_ptr || !k->eq_spec || !k->eq_stoich)
      return set_err(PFRX_E_INVALID, "MINERAL / GAS constraint without its reaction tables%s", "");
    for (int i = 0; i < naq; i++) {
      if (k->eq_ptr[i] > k->eq_ptr[i + 1] || k->eq_ptr[i] < 0) return set_err(PFRX_E_INVALID, "pfrx_constraint.eq_ptr is not a CSR row pointer%s", "");
      for (int p = k->eq_ptr[i]; p < k->eq_ptr[i + 1]; p++)
        if (k->eq_spec[p] < 0 || k->eq_spec[p] >= naq) return set_err(PFRX_E_INVALID, "pfrx_constraint.eq_spec out of range%s", "");
    }
  }
  CUDA_OK(cudaSetDevice(h->device));
  const KernelGetter *gt = nullptr;
  for (const auto &g : g_getters)
    if (g.n == h->npad) gt = &g;
  if (!gt) return set_err(PFRX_E_LIMIT, "no kernel variant for this size%s", "");
  pfrx_constraint_fn fn = gt->get_cons();
  DevCfg cfg;
  int t = 0, nb = 0;
  size_t smem_bytes = 0;
  {
    int rc0 = tpc_launch_shape(h, (const void *)fn, &cfg, &t, &smem_bytes, &nb);
    if (rc0) return rc0;
  }
  if (h->ncell <= 0) return PFRX_OK;
  // the constraint's tables: one small device block per call (set-up path, not the time loop)
  const int nnz = any_eq ? k->eq_ptr[naq] : 0;
  std::vector<double> zero(naq, 0.0);
  std::vector<int> ptr0(naq + 1, 0);
  Arena A;
  DevCons dc;
  memset(&dc, 0, sizeof(dc));
  dc.init_molality = k->initialize_with_molality != 0;
  dc.max_iterations = k->max_iterations;
  A.add(k->type, naq, &dc.type);
  A.add(h->pri_Z_host.data(), naq, &dc.Z);
  A.add(any_eq ? k->eq_logK : zero.data(), naq, &dc.eq_logK);
  if (any_eq && k->eq_logK_coef) A.add(k->eq_logK_coef, 5 * naq, &dc.eq_logKcoef);
  A.add(any_eq ? k->eq_h2o_stoich : zero.data(), naq, &dc.eq_h2o);
  A.add(any_eq ? k->eq_ptr : ptr0.data(), naq + 1, &dc.eq_ptr);
  if (nnz > 0) {
    A.add(k->eq_spec, nnz, &dc.eq_spec);
    A.add(k->eq_stoich, nnz, &dc.eq_st);
  }
  void *blk = nullptr;
  CUDA_OK(cudaMalloc(&blk, std::max<size_t>(A.bytes.size(), 16)));
  int rc = PFRX_OK;
  if (cudaMemcpy(blk, A.bytes.data(), A.bytes.size(), cudaMemcpyHostToDevice) != cudaSuccess) rc = PFRX_E_CUDA;
  if (rc == PFRX_OK) {
    for (auto &f : A.fix) *f.second = (unsigned char *)blk + f.first;
    dc.conc = conc;
    int64_t need = (h->ncell + t - 1) / t;
    int grid = (int)std::max<int64_t>(1, std::min<int64_t>(need, (int64_t)h->sm_count * nb));
    fn<<<grid, t, smem_bytes, h->stream>>>(cfg, h->st, h->ncell, dc, num_iterations, ierror);
    if (cudaGetLastError() != cudaSuccess) rc = PFRX_E_CUDA;
    h->launches++;
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) rc = PFRX_E_CUDA;
  }
  cudaFree(blk);
  if (rc != PFRX_OK) return set_err(rc, "CUDA error in pfrx_equilibrate_constraint%s", "");
  return PFRX_OK;
}

// ---- block vectors <-> SoA around the cell loop (pmc_subsurface_osrt.F90:260-274, 303-333, 356-376) ----
// One block moves a tile of OS_CELLS cells (128 by default) x ncomp components through shared memory, so that both
// sides are coalesced: the block vector is read / written as one contiguous run of OS_CELLS*ncomp
// doubles, the SoA fields as runs of OS_CELLS consecutive cells per component.  The tile is padded
// (ncomp | 1 doubles per cell) so that the cell-wise accesses are bank-conflict free.
#define PFRX_OS_CELLS_DEFAULT 128
enum { OS_FIXED_ACCUM = 0, OS_LOAD = 1, OS_STORE = 2 };

template <int MODE, int OS_CELLS>
__global__ void __launch_bounds__(OS_CELLS) pfrx_os_kernel(DevState st, int64_t ncell, int naq, int nim,
                                                            const double *in_a, const double *in_b, double *out) {
  extern __shared__ double tile[];
  const int n = naq + nim, ldt = n | 1;
  unsigned char *act_s = reinterpret_cast<unsigned char *>(tile + OS_CELLS * ldt);  // activity of the tile's cells
  const int64_t ntile = (ncell + OS_CELLS - 1) / OS_CELLS;
  // element e = threadIdx.x + k * OS_CELLS of the tile's block-vector run is (cell e / n, component
  // e % n): one division per thread, then increments
  const int cc0 = threadIdx.x / n, i0 = threadIdx.x - cc0 * n;
  const int qs = OS_CELLS / n, rs = OS_CELLS - qs * n;
#define PFRX_OS_FOR_ELEMENTS(body)                                        \
  {                                                                       \
    int cc = cc0, i = i0;                                                 \
    _Pragma("unroll 4") for (int e = threadIdx.x; e < nc * n; e += OS_CELLS) { \
      body;                                                               \
      cc += qs;                                                           \
      i += rs;                                                            \
      if (i >= n) {                                                       \
        i -= n;                                                           \
        cc++;                                                             \
      }                                                                   \
    }                                                                     \
  }
  for (int64_t t = blockIdx.x; t < ntile; t += gridDim.x) {
    const int64_t c0 = t * OS_CELLS;
    const int nc = (int)min((int64_t)OS_CELLS, ncell - c0);
    const int64_t c = c0 + threadIdx.x;
    const bool mine = threadIdx.x < nc;
    const bool active = mine && !(st.imat && st.imat[c] <= 0);
    act_s[threadIdx.x] = active ? 1 : 0;
    double *row = tile + threadIdx.x * ldt;
    double *vout = out ? out + c0 * n : nullptr;
    if (MODE == OS_FIXED_ACCUM) {
      // SoA -> tile (per cell), tile -> block vector (contiguous); only aqueous entries of active
      // cells are written, the rest of the vector is left as it is
      if (active) {
        const double f = st.porosity[c] * st.sat[c] * 1000.0 * st.volume[c];
#pragma unroll 8
        for (int k = 0; k < naq; k++) row[k] = f * __ldcs(st.total + k * st.ld + c);
      }
      __syncthreads();
      PFRX_OS_FOR_ELEMENTS(if (i < naq && act_s[cc]) vout[e] = tile[cc * ldt + i]);
      __syncthreads();
    } else if (MODE == OS_LOAD) {
      // block vectors -> tile (contiguous), tile -> SoA (per cell)
      if (in_a) {
        const double *va = in_a + c0 * n;
        PFRX_OS_FOR_ELEMENTS(if (i < naq) tile[cc * ldt + i] = __ldcs(va + e));
      }
      if (in_b && nim > 0) {
        const double *vb = in_b + c0 * n;
        PFRX_OS_FOR_ELEMENTS(if (i >= naq) tile[cc * ldt + i] = __ldcs(vb + e));
      }
      __syncthreads();
      if (active) {
        if (in_a)
#pragma unroll 8
          for (int k = 0; k < naq; k++) st.total[k * st.ld + c] = row[k];
        if (in_b)
#pragma unroll 8
          for (int k = 0; k < nim; k++) st.immobile[k * st.ld + c] = row[naq + k];
      }
      __syncthreads();
    } else {
      if (active) {
#pragma unroll 8
        for (int k = 0; k < naq; k++) row[k] = __ldcs(st.pri_molal + k * st.ld + c);
#pragma unroll 8
        for (int k = 0; k < nim; k++) row[naq + k] = __ldcs(st.immobile + k * st.ld + c);
      }
      __syncthreads();
      PFRX_OS_FOR_ELEMENTS(if (act_s[cc]) vout[e] = tile[cc * ldt + i]);
      __syncthreads();
    }
  }
#undef PFRX_OS_FOR_ELEMENTS
}

// enqueue one transpose of `ncell` cells of `st` (already offset to the first cell) on `s`
template <int MODE, int OS_CELLS>
static int os_enqueue_t(pfrx_handle *h, const DevState &st, int64_t ncell, const double *a, const double *b,
                        double *out, cudaStream_t s) {
  const int naq = h->cfg.naq, nim = h->cfg.nim;
  const size_t smem = (size_t)OS_CELLS * ((naq + nim) | 1) * sizeof(double) + OS_CELLS;  // tile + activity flags
  const int64_t ntile = (ncell + OS_CELLS - 1) / OS_CELLS;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ntile, (int64_t)h->sm_count * (2048 / OS_CELLS)));
  CUDA_OK(cudaFuncSetAttribute((const void *)pfrx_os_kernel<MODE, OS_CELLS>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pfrx_os_kernel<MODE, OS_CELLS><<<grid, OS_CELLS, smem, s>>>(st, ncell, naq, nim, a, b, out);
  CUDA_OK(cudaGetLastError());
  h->launches++;
  return PFRX_OK;
}

template <int MODE>
static int os_enqueue(pfrx_handle *h, const DevState &st, int64_t ncell, const double *a, const double *b,
                      double *out, cudaStream_t s) {
  if (ncell <= 0) return PFRX_OK;
  // tile height: smaller tiles keep more blocks per SM in different phases (load / store) of the
  // load -> barrier -> store cycle; PFRX_OS_CELLS overrides for measurements
  static const int cells = [] {
    const char *ev = getenv("PFRX_OS_CELLS");
    return ev ? atoi(ev) : PFRX_OS_CELLS_DEFAULT;
  }();
  if (cells == 64) return os_enqueue_t<MODE, 64>(h, st, ncell, a, b, out, s);
  if (cells == 128) return os_enqueue_t<MODE, 128>(h, st, ncell, a, b, out, s);
  return os_enqueue_t<MODE, 256>(h, st, ncell, a, b, out, s);
}

template <int MODE>
static int os_launch(pfrx_handle *h, const double *a, const double *b, double *out) {
  if (!h) return set_err(PFRX_E_INVALID, "null argument%s", "");
  if (!h->bound) return set_err(PFRX_E_NOTBOUND, "pfrx_bind_state has not been called%s", "");
  CUDA_OK(cudaSetDevice(h->device));
  if (h->ncell <= 0) return PFRX_OK;
  int rc = os_enqueue<MODE>(h, h->st, h->ncell, a, b, out, h->stream);
  if (rc) return rc;
  CUDA_OK(cudaStreamSynchronize(h->stream));
  return PFRX_OK;
}

extern "C" int pfrx_os_fixed_accum(pfrx_handle *h, double *fixed_accum) {
  if (!fixed_accum) return set_err(PFRX_E_INVALID, "null argument%s", "");
  return os_launch<OS_FIXED_ACCUM>(h, nullptr, nullptr, fixed_accum);
}
extern "C" int pfrx_os_load(pfrx_handle *h, const double *solved_total, const double *tran_xx) {
  if (!solved_total && !tran_xx) return set_err(PFRX_E_INVALID, "null argument%s", "");
  return os_launch<OS_LOAD>(h, solved_total, tran_xx, nullptr);
}
extern "C" int pfrx_os_store(pfrx_handle *h, double *tran_xx) {
  if (!tran_xx) return set_err(PFRX_E_INVALID, "null argument%s", "");
  return os_launch<OS_STORE>(h, nullptr, nullptr, tran_xx);
}

// the views of `d` starting at cell c0 (same leading dimension)
static DevState dev_state_at(const DevState &d, int64_t c0) {
  DevState dc = d;
#define PFRX_OFF(field) \
  if (dc.field) dc.field = dc.field + c0
  PFRX_OFF(total);
  PFRX_OFF(pri_molal);
  PFRX_OFF(immobile);
  PFRX_OFF(pri_act_coef);
  PFRX_OFF(sec_act_coef);
  PFRX_OFF(sec_molal);
  PFRX_OFF(ln_act_h2o);
  PFRX_OFF(mnrl_volfrac);
  PFRX_OFF(mnrl_area);
  PFRX_OFF(mnrl_rate);
  PFRX_OFF(free_site);
  PFRX_OFF(eqsrfcplx_conc);
  PFRX_OFF(total_sorb_eq);
  PFRX_OFF(kinmr);
  PFRX_OFF(den_kg);
  PFRX_OFF(sat);
  PFRX_OFF(temp);
  PFRX_OFF(porosity);
  PFRX_OFF(volume);
  PFRX_OFF(soil_particle_density);
  PFRX_OFF(elm_w);
  PFRX_OFF(elm_o);
  PFRX_OFF(elm_t);
  PFRX_OFF(elm_zsoil);
  PFRX_OFF(elm_kscalar);
  PFRX_OFF(elm_bd_dry);
  PFRX_OFF(elm_bsw);
  PFRX_OFF(somdec_nc);
  PFRX_OFF(elm_plantndemand);
  PFRX_OFF(eqionx_ref);
  PFRX_OFF(eqionx_conc);
  PFRX_OFF(pres);
  PFRX_OFF(sandbox_aux);
  PFRX_OFF(sat_gas);
  PFRX_OFF(total_gas);
  PFRX_OFF(gas_pp);
  PFRX_OFF(elm_sucsat);
  PFRX_OFF(elm_watfc);
  PFRX_OFF(elm_effpor);
  PFRX_OFF(imat);
  PFRX_OFF(num_sub_steps);
  PFRX_OFF(num_iterations);
  PFRX_OFF(num_kinetic_state_updates);
  PFRX_OFF(ierror);
#undef PFRX_OFF
  return dc;
}

static int pipeline_events(pfrx_handle *h) {
  if (!h->ev_ready) {
    CUDA_OK(cudaStreamCreateWithFlags(&h->out_stream, cudaStreamNonBlocking));
    for (int i = 0; i < PFRX_MAX_CHUNKS; i++) {
      CUDA_OK(cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
      CUDA_OK(cudaEventCreateWithFlags(&h->ev_k[i], cudaEventDisableTiming));
    }
    h->ev_ready = true;
  }
  return PFRX_OK;
}

// ---- the operator-split step with the chemistry state resident in HBM --------------
// pmc_subsurface_osrt.F90:303-378 as one call: what crosses the host link per step is what PETSc
// holds -- the solved totals and tran_xx, ncomp doubles per cell each -- while rt_auxvars live in the
// bound device state from step to step.  Chunks of cells are pipelined over three streams: upload of
// chunk i+1, {load transpose, RStep kernel, store transpose} of chunk i, download of chunk i-1.
__global__ void pfrx_count_inactive_kernel(const int *imat, long long ncell, int *count) {
  int local = 0;
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < ncell; c += (long long)gridDim.x * blockDim.x)
    local += imat[c] <= 0 ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(count, local);
}

extern "C" int pfrx_os_step_host(pfrx_handle *h, const double *solved_total, double *tran_xx, double tran_dt,
                                 pfrx_step_result *out) {
  if (!h || !out) return set_err(PFRX_E_INVALID, "null argument%s", "");
  if (!h->bound) return set_err(PFRX_E_NOTBOUND, "pfrx_bind_state has not been called%s", "");
  CUDA_OK(cudaSetDevice(h->device));
  const int64_t ncell = h->ncell;
  if (ncell == 0) {  // a rank without cells: its vectors are empty (and may be NULL)
    memset(out, 0, sizeof(*out));
    out->first_failed_cell = -1;
    return PFRX_OK;
  }
  if (!tran_xx) return set_err(PFRX_E_INVALID, "null argument%s", "");
  const int n = h->cfg.naq + h->cfg.nim;
  if (h->os_cap < ncell) {
    if (h->os_a) CUDA_OK(cudaFree(h->os_a));
    if (h->os_b) CUDA_OK(cudaFree(h->os_b));
    h->os_a = h->os_b = nullptr;
    h->os_cap = 0;
    CUDA_OK(cudaMalloc(&h->os_a, (size_t)ncell * n * sizeof(double)));
    CUDA_OK(cudaMalloc(&h->os_b, (size_t)ncell * n * sizeof(double)));
    h->os_cap = ncell;
  }
  int rc = pipeline_events(h);
  if (rc) return rc;
  // Chunks: a handful hides all but the first upload and the last download behind the kernel, but
  // every chunk of a ragged workload ends with its own slowest cell.  Which wins is measured: calls 0
  // and 1 on a shard run in one chunk and in `many` and are not timed (module load, first touch of
  // the staging buffers), calls 2 and 3 time the two shapes, the faster is kept; the trial is repeated
  // every 64 calls because raggedness changes over a run.  PFRX_OS_CHUNKS (read once) pins the count.
  const int many = (int)std::min<int64_t>(16, std::max<int64_t>(1, ncell / 262144));
  if (h->os_env_chunks < 0) {
    const char *ev = getenv("PFRX_OS_CHUNKS");
    h->os_env_chunks = ev ? std::max(1, std::min(PFRX_MAX_CHUNKS, atoi(ev))) : 0;
  }
  if (h->os_trial_ncell != ncell) {
    h->os_trial_ncell = ncell;
    h->os_calls = 0;
    h->os_choice = 1;
  }
  const bool forced = h->os_env_chunks > 0;
  // 0, 1 warm-up (one chunk, many: first touch of the staging buffers and events of either shape),
  // 2 trial (one chunk), 3 trial (many), >= 4 steady
  const int phase = h->os_calls;
  int nchunk = forced ? h->os_env_chunks : (phase == 0 || phase == 2) ? 1 : (phase == 1 || phase == 3) ? many : h->os_choice;
  if (h->os_inactive < 0 && h->st.imat) {
    int *cnt = nullptr;
    CUDA_OK(cudaMalloc(&cnt, sizeof(int)));
    CUDA_OK(cudaMemsetAsync(cnt, 0, sizeof(int), h->stream));
    pfrx_count_inactive_kernel<<<296, 256, 0, h->stream>>>(h->st.imat, (long long)ncell, cnt);
    int host_cnt = 0;
    CUDA_OK(cudaMemcpyAsync(&host_cnt, cnt, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    CUDA_OK(cudaFree(cnt));
    h->os_inactive = host_cnt;
  }
  const auto wall0 = std::chrono::steady_clock::now();
  cudaStream_t s_in = h->copy_stream, s_k = h->stream, s_out = h->out_stream;
  rc = summary_reset(h, s_k);
  if (rc) return rc;
  // Even and odd chunks run on two kernel streams: a chunk's kernel is a persistent grid of one block per SM,
  // and the blocks of the next chunk start on the SMs the current one has already left instead of waiting for
  // its slowest block.  Not for the refill kernels (one work counter per launch; ragged shards run in one chunk).
  const bool dual = nchunk > 1 && !(h->spec_func && h->spec_refill) && !getenv("PFRX_OS_ONE_KERNEL_STREAM");
  if (dual) {
    if (!h->stream2) {
      CUDA_OK(cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking));
      CUDA_OK(cudaEventCreateWithFlags(&h->ev_reset, cudaEventDisableTiming));
    }
    CUDA_OK(cudaEventRecord(h->ev_reset, s_k));  // the summary is zeroed before either stream adds to it
    CUDA_OK(cudaStreamWaitEvent(h->stream2, h->ev_reset, 0));
  }
  int last_odd = -1;
  // tran_xx goes up when the step reads it (immobile entries) or when the shard has inactive cells,
  // which must keep their entries through the download
  const bool up_xx = h->cfg.nim > 0 || (h->st.imat != nullptr && h->os_inactive != 0);
  h->last_h2d = h->last_d2h = 0;
  const size_t w8 = sizeof(double);
  // Chunk boundaries.  The first upload and the last download are the only transfers the kernel
  // cannot hide, so the chunks grow from the ends towards the middle (1 : 2 : 4 : ... : 4 : 2 : 1),
  // and every boundary sits on a whole number of kernel waves (resident blocks x cells per block)
  // so that no chunk ends in a partly filled wave.
  int64_t bnd[PFRX_MAX_CHUNKS + 1];
  {
    const int64_t wave = h->spec_func ? (int64_t)h->sm_count * h->spec_blocks_per_sm * h->spec_cells
                                      : (int64_t)h->sm_count * h->blocks_per_sm * h->threads;
    double wsum = 0.0, wgt[PFRX_MAX_CHUNKS];
    for (int ch = 0; ch < nchunk; ch++) {
      const int e = std::min(std::min(ch, nchunk - 1 - ch), 3);
      wgt[ch] = (nchunk >= 4 && !getenv("PFRX_OS_EQUAL_CHUNKS")) ? (double)(1 << e) : 1.0;
      wsum += wgt[ch];
    }
    double acc = 0.0;
    bnd[0] = 0;
    for (int ch = 0; ch < nchunk; ch++) {
      acc += wgt[ch];
      int64_t b = (int64_t)((double)ncell * acc / wsum);
      if (wave > 0 && ncell >= 4 * wave * nchunk) b = (b + wave / 2) / wave * wave;
      bnd[ch + 1] = std::max(bnd[ch], std::min(ncell, b));
    }
    bnd[nchunk] = ncell;
  }
  for (int ch = 0; ch < nchunk; ch++) {
    const int64_t c0 = bnd[ch], c1 = bnd[ch + 1], nc = c1 - c0;
    if (nc <= 0) continue;
    const size_t off = (size_t)c0 * n, cnt = (size_t)nc * n;
    if (solved_total) {
      CUDA_OK(cudaMemcpyAsync(h->os_a + off, solved_total + off, cnt * w8, cudaMemcpyHostToDevice, s_in));
      h->last_h2d += (int64_t)(cnt * w8);
    }
    if (up_xx) {
      CUDA_OK(cudaMemcpyAsync(h->os_b + off, tran_xx + off, cnt * w8, cudaMemcpyHostToDevice, s_in));
      h->last_h2d += (int64_t)(cnt * w8);
    }
    CUDA_OK(cudaEventRecord(h->ev_in[ch], s_in));
    cudaStream_t s_c = (dual && (ch & 1)) ? h->stream2 : s_k;  // this chunk's kernel stream
    if (dual && (ch & 1)) last_odd = ch;
    CUDA_OK(cudaStreamWaitEvent(s_c, h->ev_in[ch], 0));
    const DevState dc = dev_state_at(h->st, c0);
    if (solved_total || h->cfg.nim > 0) {
      rc = os_enqueue<OS_LOAD>(h, dc, nc, solved_total ? h->os_a + off : nullptr,
                               h->cfg.nim > 0 ? h->os_b + off : nullptr, nullptr, s_c);
      if (rc) return rc;
    }
    rc = launch_kernel(h, dc, nc, tran_dt, s_c);
    if (rc) return rc;
    rc = os_enqueue<OS_STORE>(h, dc, nc, nullptr, nullptr, h->os_b + off, s_c);
    if (rc) return rc;
    CUDA_OK(cudaEventRecord(h->ev_k[ch], s_c));
    CUDA_OK(cudaStreamWaitEvent(s_out, h->ev_k[ch], 0));
    CUDA_OK(cudaMemcpyAsync(tran_xx + off, h->os_b + off, cnt * w8, cudaMemcpyDeviceToHost, s_out));
    h->last_d2h += (int64_t)(cnt * w8);
  }
  if (last_odd >= 0) CUDA_OK(cudaStreamWaitEvent(s_k, h->ev_k[last_odd], 0));  // the summary of both streams
  CUDA_OK(cudaMemcpyAsync(h->h_summ, h->d_summ, sizeof(DevSummary), cudaMemcpyDeviceToHost, s_k));
  CUDA_OK(cudaStreamSynchronize(s_k));
  CUDA_OK(cudaStreamSynchronize(s_out));
  h->pending = false;
  summary_out(h, out);
  if (!forced) {
    const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count();
    if (phase == 2) h->os_trial_s[0] = el;
    if (phase == 3) {
      h->os_trial_s[1] = el;
      h->os_choice = h->os_trial_s[1] < h->os_trial_s[0] ? many : 1;
    }
    h->os_calls = phase >= 67 ? 2 : phase + 1;  // steady for 64 calls, then the two trials again
  }
  if (out->first_failed_cell >= 0 && nchunk > 1) {
    // chunk-local index: recover the shard index from the per-cell error flags
    std::vector<int> ie((size_t)ncell);
    CUDA_OK(cudaMemcpy(ie.data(), h->st.ierror, (size_t)ncell * sizeof(int), cudaMemcpyDeviceToHost));
    out->first_failed_cell = -1;
    for (int64_t c = 0; c < ncell; c++)
      if (ie[c] != 0) {
        out->first_failed_cell = c;
        break;
      }
  }
  return PFRX_OK;
}

// ---- host-resident state: H2D, kernel, D2H ------------------------------------
static const int kNumD = PFRX_NUM_D;
static size_t field_off(const int *rows, int64_t ld, int f) {
  size_t o = 0;
  for (int i = 0; i < f; i++) o += (size_t)rows[i] * ld;
  return o;
}

extern "C" int pfrx_rstep_host(pfrx_handle *h, int64_t ncell, const pfrx_state *host, double tran_dt,
                               pfrx_step_result *out) {
  if (!h || !host || !out || ncell < 0 || host->ld < ncell) return set_err(PFRX_E_INVALID, "bad arguments%s", "");
  CUDA_OK(cudaSetDevice(h->device));
  if (ncell == 0) {  // empty shard: nothing to copy or launch
    memset(out, 0, sizeof(*out));
    out->first_failed_cell = -1;
    return PFRX_OK;
  }
  const int *rows = h->rows_d.data();
  size_t ndbl = 0;
  for (int i = 0; i < kNumD; i++) ndbl += (size_t)rows[i] * ncell;
  size_t bytes = ndbl * sizeof(double) + (size_t)5 * ncell * sizeof(int);
  bool fresh_alloc = false;
  if (h->own_ncell != ncell) {
    fresh_alloc = true;
    if (h->own) CUDA_OK(cudaFree(h->own));
    h->own = nullptr;
    CUDA_OK(cudaMalloc(&h->own, std::max<size_t>(bytes, 16)));
    h->own_ncell = ncell;
    double *base = (double *)h->own;
    double **dst[kNumD] = {&h->own_st.total,        &h->own_st.pri_molal,    &h->own_st.immobile,
                           &h->own_st.pri_act_coef, &h->own_st.sec_act_coef, &h->own_st.sec_molal,
                           &h->own_st.ln_act_h2o,   &h->own_st.mnrl_volfrac, &h->own_st.mnrl_area,
                           &h->own_st.mnrl_rate,    &h->own_st.free_site,    &h->own_st.eqsrfcplx_conc,
                           &h->own_st.total_sorb_eq, &h->own_st.kinmr,       (double **)&h->own_st.den_kg,
                           (double **)&h->own_st.sat, (double **)&h->own_st.temp, (double **)&h->own_st.porosity,
                           (double **)&h->own_st.volume, (double **)&h->own_st.soil_particle_density,
                           (double **)&h->own_st.elm_w, (double **)&h->own_st.elm_o, (double **)&h->own_st.elm_t,
                           (double **)&h->own_st.elm_zsoil, (double **)&h->own_st.elm_kscalar,
                           (double **)&h->own_st.elm_bd_dry, (double **)&h->own_st.elm_bsw, &h->own_st.somdec_nc,
                           (double **)&h->own_st.elm_plantndemand, &h->own_st.eqionx_ref, &h->own_st.eqionx_conc,
                           (double **)&h->own_st.pres, &h->own_st.sandbox_aux,
                           (double **)&h->own_st.sat_gas, &h->own_st.total_gas, &h->own_st.gas_pp,
                           (double **)&h->own_st.elm_sucsat, (double **)&h->own_st.elm_watfc,
                           (double **)&h->own_st.elm_effpor};
    for (int f = 0; f < kNumD; f++) *dst[f] = rows[f] ? base + field_off(rows, ncell, f) : nullptr;
    int *ib = (int *)(base + ndbl);
    h->own_st.imat = ib;
    h->own_st.num_sub_steps = ib + ncell;
    h->own_st.num_iterations = ib + 2 * ncell;
    h->own_st.num_kinetic_state_updates = ib + 3 * ncell;
    h->own_st.ierror = ib + 4 * ncell;
    h->own_st.ld = ncell;
  }
  DevState d = h->own_st;
  const double *src[kNumD] = {host->total,        host->pri_molal,    host->immobile,  host->pri_act_coef,
                              host->sec_act_coef, host->sec_molal,    host->ln_act_h2o, host->mnrl_volfrac,
                              host->mnrl_area,    host->mnrl_rate,    host->srfcplxrxn_free_site_conc,
                              host->eqsrfcplx_conc, host->total_sorb_eq, host->kinmr_total_sorb, host->den_kg,
                              host->sat,          host->temp,         host->porosity,  host->volume,
                              host->soil_particle_density, host->elm_w_scalar, host->elm_o_scalar, host->elm_t_scalar,
                              host->elm_zsoil,    host->elm_kscalar_decomp_c, host->elm_bulkdensity_dry, host->elm_bsw,
                              host->somdec_nc,    host->elm_rate_plantndemand, host->eqionx_ref_cation_sorbed_conc,
                              host->eqionx_conc,  host->pres,         host->sandbox_aux,
                              host->sat_gas,      host->total_gas,    host->gas_pp,
                              host->elm_sucsat,   host->elm_watfc,    host->elm_effporosity};
  double *dptr[kNumD] = {d.total,        d.pri_molal,    d.immobile,  d.pri_act_coef, d.sec_act_coef,
                         d.sec_molal,    d.ln_act_h2o,   d.mnrl_volfrac, d.mnrl_area, d.mnrl_rate,
                         d.free_site,    d.eqsrfcplx_conc, d.total_sorb_eq, d.kinmr,  (double *)d.den_kg,
                         (double *)d.sat, (double *)d.temp, (double *)d.porosity, (double *)d.volume,
                         (double *)d.soil_particle_density, (double *)d.elm_w, (double *)d.elm_o, (double *)d.elm_t,
                         (double *)d.elm_zsoil, (double *)d.elm_kscalar, (double *)d.elm_bd_dry, (double *)d.elm_bsw,
                         d.somdec_nc,    (double *)d.elm_plantndemand, d.eqionx_ref, d.eqionx_conc, (double *)d.pres,
                         d.sandbox_aux,  (double *)d.sat_gas, d.total_gas, d.gas_pp,
                         (double *)d.elm_sucsat, (double *)d.elm_watfc, (double *)d.elm_effpor};
  bool have_spd = host->soil_particle_density != nullptr;
  if (!have_spd) d.soil_particle_density = nullptr;
  bool have_lnw = host->ln_act_h2o != nullptr;
  if (!have_lnw) d.ln_act_h2o = nullptr;
  bool have_sc = host->eqsrfcplx_conc != nullptr;
  if (!have_sc) d.eqsrfcplx_conc = nullptr;
  if (!host->imat) d.imat = nullptr;
  if (!host->elm_w_scalar) d.elm_w = nullptr;
  if (!host->elm_o_scalar) d.elm_o = nullptr;
  if (!host->elm_t_scalar) d.elm_t = nullptr;
  if (!host->elm_zsoil) d.elm_zsoil = nullptr;
  if (!host->elm_kscalar_decomp_c) d.elm_kscalar = nullptr;
  if (!host->elm_bulkdensity_dry) d.elm_bd_dry = nullptr;
  if (!host->elm_bsw) d.elm_bsw = nullptr;
  if (!host->somdec_nc) d.somdec_nc = nullptr;
  if (!host->elm_rate_plantndemand) d.elm_plantndemand = nullptr;
  if (!host->eqionx_ref_cation_sorbed_conc) d.eqionx_ref = nullptr;
  if (!host->eqionx_conc) d.eqionx_conc = nullptr;
  if (!host->sandbox_aux) d.sandbox_aux = nullptr;
  if (!host->sat_gas) d.sat_gas = nullptr;
  if (!host->total_gas) d.total_gas = nullptr;
  if (!host->gas_pp) d.gas_pp = nullptr;
  if (!host->elm_sucsat) d.elm_sucsat = nullptr;
  if (!host->elm_watfc) d.elm_watfc = nullptr;
  if (!host->elm_effporosity) d.elm_effpor = nullptr;
  {
    DevState chk;
    pfrx_state probe = *host;
    int rc0 = to_dev_state(h, &probe, &chk);
    if (rc0) return rc0;
  }
  // Pipeline: the shard is cut into chunks; chunk i+1 uploads (copy-in stream)
  // while chunk i computes (kernel stream) and chunk i-1 downloads (copy-out
  // stream).  PCIe is full duplex, so the step costs about
  // max(H2D, kernel, D2H) instead of their sum.
  {
    int rce = pipeline_events(h);
    if (rce) return rce;
  }
  int nchunk = 1;
  if (const char *ev = getenv("PFRX_CHUNKS")) nchunk = atoi(ev);
  else {
    nchunk = (int)std::min<int64_t>(PFRX_MAX_CHUNKS, std::max<int64_t>(1, ncell / 131072));
    if (h->host_kernel_s_per_cell > 0.0 && h->host_link_s_per_cell > 0.0 &&
        h->host_kernel_s_per_cell > 2.0 * h->host_link_s_per_cell)
      nchunk = (int)std::max(1.0, floor(nchunk * 2.0 * h->host_link_s_per_cell / h->host_kernel_s_per_cell));
  }
  if (nchunk < 1) nchunk = 1;
  if (nchunk > PFRX_MAX_CHUNKS) nchunk = PFRX_MAX_CHUNKS;
  cudaStream_t s_in = h->copy_stream, s_k = h->stream, s_out = h->out_stream;
  int rc = summary_reset(h, s_k);
  if (rc) return rc;
  const int io[] = {0, 1, 2, 3, 4, 5, 6, 7, 9, 10, 11, 12, 13, 27, 29, 30, 32, 34, 35};
  double *hdst[kNumD] = {host->total,        host->pri_molal,    host->immobile,  host->pri_act_coef,
                         host->sec_act_coef, host->sec_molal,    host->ln_act_h2o, host->mnrl_volfrac,
                         host->mnrl_area,    host->mnrl_rate,    host->srfcplxrxn_free_site_conc,
                         host->eqsrfcplx_conc, host->total_sorb_eq, host->kinmr_total_sorb, nullptr,
                         nullptr,            nullptr,            nullptr,         nullptr,
                         nullptr,            nullptr,            nullptr,         nullptr,
                         nullptr,            nullptr,            nullptr,         nullptr,
                         host->somdec_nc,    nullptr,            host->eqionx_ref_cation_sorbed_conc,
                         host->eqionx_conc,  nullptr,            host->sandbox_aux,
                         nullptr,            host->total_gas,    host->gas_pp,
                         nullptr,            nullptr,            nullptr};
  const size_t w8 = sizeof(double);
  // Fields every active cell overwrites before it reads them need no upload -- as
  // long as every cell is active (imat absent or all positive), otherwise the download would hand the
  // inactive cells garbage.  With activity coefficients updated in every Newton
  // iteration (reaction.F90:3868) both coefficient arrays are outputs only; with
  // frozen coefficients the complex concentrations are (RTotal rewrites them before
  // anything reads them); mineral rates always are.
  bool skip_in[kNumD] = {false};
  bool all_active = true;
  if (host->imat)
    for (int64_t c = 0; c < ncell && all_active; c++) all_active = host->imat[c] > 0;
  if (all_active && h->cfg.use_full_geochemistry && !getenv("PFRX_UPLOAD_ALL")) {
    const bool act_upd = h->cfg.act_freq == PFRX_ACT_COEF_FREQUENCY_NEWTON_ITER;
    if (act_upd && h->cfg.act_alg != PFRX_ACT_COEF_ALGORITHM_NEWTON) skip_in[3] = skip_in[4] = true;
    if (!act_upd && !h->cfg.use_act_h2o) skip_in[5] = true;
    if (h->cfg.nkin > 0) skip_in[9] = true;
  }
  if (fresh_alloc) {
    // poison what is never uploaded: a kernel that reads it anyway shows up as NaN
    for (int f = 0; f < kNumD; f++)
      if (skip_in[f] && rows[f]) CUDA_OK(cudaMemsetAsync(dptr[f], 0xff, (size_t)rows[f] * ncell * w8, s_in));
  }
  h->last_h2d = h->last_d2h = 0;
  if (!h->ev_t0) {
    CUDA_OK(cudaEventCreate(&h->ev_t0));
    CUDA_OK(cudaEventCreate(&h->ev_t1));
  }
  float kernel_ms_sum = 0.f;
  (void)kernel_ms_sum;
  for (int ch = 0; ch < nchunk; ch++) {
    int64_t c0 = ncell * ch / nchunk, c1 = ncell * (ch + 1) / nchunk, nc = c1 - c0;
    if (nc <= 0) continue;
    for (int f = 0; f < kNumD; f++) {
      if (!rows[f] || !src[f] || f == 11) continue;  // eqsrfcplx_conc is output only
      if (skip_in[f]) continue;
      if (!fresh_alloc && ((h->host_resident_mask >> f) & 1ull)) continue;  // the device copy is the current one
      CUDA_OK(cudaMemcpy2DAsync(dptr[f] + c0, ncell * w8, src[f] + c0, host->ld * w8, nc * w8, rows[f],
                                cudaMemcpyHostToDevice, s_in));
      h->last_h2d += (int64_t)(nc * w8) * rows[f];
    }
    if (host->imat) {
      CUDA_OK(cudaMemcpyAsync((void *)(d.imat + c0), host->imat + c0, nc * sizeof(int), cudaMemcpyHostToDevice, s_in));
      h->last_h2d += nc * (int64_t)sizeof(int);
    }
    CUDA_OK(cudaEventRecord(h->ev_in[ch], s_in));
    CUDA_OK(cudaStreamWaitEvent(s_k, h->ev_in[ch], 0));
    DevState dc = dev_state_at(d, c0);
    if (ch == 0) CUDA_OK(cudaEventRecord(h->ev_t0, s_k));
    rc = launch_kernel(h, dc, nc, tran_dt, s_k);
    if (rc) return rc;
    if (ch == nchunk - 1) CUDA_OK(cudaEventRecord(h->ev_t1, s_k));
    CUDA_OK(cudaEventRecord(h->ev_k[ch], s_k));
    CUDA_OK(cudaStreamWaitEvent(s_out, h->ev_k[ch], 0));
    for (int f : io) {
      if (!rows[f] || !hdst[f]) continue;
      if ((h->host_resident_mask >> f) & 1ull) continue;
      CUDA_OK(cudaMemcpy2DAsync(hdst[f] + c0, host->ld * w8, dptr[f] + c0, ncell * w8, nc * w8, rows[f],
                                cudaMemcpyDeviceToHost, s_out));
      h->last_d2h += (int64_t)(nc * w8) * rows[f];
    }
    h->last_d2h += 4 * nc * (int64_t)sizeof(int);
    CUDA_OK(cudaMemcpyAsync(host->num_sub_steps + c0, d.num_sub_steps + c0, nc * sizeof(int), cudaMemcpyDeviceToHost, s_out));
    CUDA_OK(cudaMemcpyAsync(host->num_iterations + c0, d.num_iterations + c0, nc * sizeof(int), cudaMemcpyDeviceToHost, s_out));
    CUDA_OK(cudaMemcpyAsync(host->num_kinetic_state_updates + c0, d.num_kinetic_state_updates + c0, nc * sizeof(int),
                            cudaMemcpyDeviceToHost, s_out));
    CUDA_OK(cudaMemcpyAsync(host->ierror + c0, d.ierror + c0, nc * sizeof(int), cudaMemcpyDeviceToHost, s_out));
  }
  CUDA_OK(cudaMemcpyAsync(h->h_summ, h->d_summ, sizeof(DevSummary), cudaMemcpyDeviceToHost, s_k));
  CUDA_OK(cudaStreamSynchronize(s_k));
  CUDA_OK(cudaStreamSynchronize(s_out));
  h->pending = false;
  summary_out(h, out);
  if (nchunk == 1 || h->host_kernel_s_per_cell == 0.0) {
    // with one chunk the span between the events is the kernel alone; with several it also
    // holds the waits for uploads, so it is only taken as the first estimate
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev_t0, h->ev_t1) == cudaSuccess && ms > 0.f)
      h->host_kernel_s_per_cell = 1.e-3 * ms / (double)ncell;
  }
  // host link at ~45 GB/s per direction (PCIe gen5 x16, measured through this path), full duplex
  h->host_link_s_per_cell = (double)std::max(h->last_h2d, h->last_d2h) / 45.e9 / (double)ncell;
  // first_failed_cell of chunked launches is chunk-local; recover the shard index
  if (out->first_failed_cell >= 0 && nchunk > 1) {
    out->first_failed_cell = -1;
    for (int64_t c = 0; c < ncell; c++)
      if (host->ierror[c] != 0) {
        out->first_failed_cell = c;
        break;
      }
  }
  return PFRX_OK;
}

// Fields of pfrx_state (bit f = the f-th double field in declaration order) that pfrx_rstep_host
// keeps RESIDENT in its device mirror: uploaded when the mirror is (re)allocated, never downloaded.
// Meant for what the step derives and only the next step reads -- rt_auxvar%sec_molal and the
// activity coefficients are 1.4 of the 1.97 KB per cell that a Hanford step sends back.  The host
// copies of resident fields go stale; pfrx_rstep_host_fetch brings them up to date on request
// (output, checkpoint).
extern "C" int pfrx_rstep_host_resident(pfrx_handle *h, uint64_t field_mask) {
  if (!h) return set_err(PFRX_E_INVALID, "null handle%s", "");
  if (field_mask >> 14) return set_err(PFRX_E_INVALID, "field mask has bits beyond the 14 state fields%s", "");
  h->host_resident_mask = field_mask;
  return PFRX_OK;
}

extern "C" int pfrx_rstep_host_fetch(pfrx_handle *h, int64_t ncell, const pfrx_state *host) {
  if (!h || !host) return set_err(PFRX_E_INVALID, "null argument%s", "");
  if (!h->own || h->own_ncell != ncell) return set_err(PFRX_E_NOTBOUND, "no device mirror of that many cells%s", "");
  CUDA_OK(cudaSetDevice(h->device));
  const int *rows = h->rows_d.data();
  double *hdst[PFRX_NUM_D] = {host->total,        host->pri_molal,    host->immobile,  host->pri_act_coef,
                              host->sec_act_coef, host->sec_molal,    host->ln_act_h2o, host->mnrl_volfrac,
                              host->mnrl_area,    host->mnrl_rate,    host->srfcplxrxn_free_site_conc,
                              host->eqsrfcplx_conc, host->total_sorb_eq, host->kinmr_total_sorb};
  double *dptr[14] = {h->own_st.total,        h->own_st.pri_molal,    h->own_st.immobile,  h->own_st.pri_act_coef,
                      h->own_st.sec_act_coef, h->own_st.sec_molal,    h->own_st.ln_act_h2o, h->own_st.mnrl_volfrac,
                      h->own_st.mnrl_area,    h->own_st.mnrl_rate,    h->own_st.free_site,
                      h->own_st.eqsrfcplx_conc, h->own_st.total_sorb_eq, h->own_st.kinmr};
  for (int f = 0; f < 14; f++) {
    if (!((h->host_resident_mask >> f) & 1ull) || !rows[f] || !hdst[f] || !dptr[f]) continue;
    CUDA_OK(cudaMemcpy2DAsync(hdst[f], host->ld * sizeof(double), dptr[f], ncell * sizeof(double), ncell * sizeof(double),
                              rows[f], cudaMemcpyDeviceToHost, h->stream));
  }
  CUDA_OK(cudaStreamSynchronize(h->stream));
  return PFRX_OK;
}

// ---- multi-GPU -----------------------------------------------------------------
extern "C" int pfrx_comm_unique_id(void *id128) {
  int rc = load_nccl();
  if (rc) return rc;
  NcclUid id;
  int e = g_nccl.GetUniqueId(&id);
  if (e) return set_err(PFRX_E_NCCL, "ncclGetUniqueId: %s", g_nccl.GetErrorString(e));
  memcpy(id128, &id, sizeof(id));
  return PFRX_OK;
}

extern "C" int pfrx_comm_init(pfrx_handle *h, int nranks, int rank, const void *id128) {
  if (!h || !id128) return set_err(PFRX_E_INVALID, "null argument%s", "");
  int rc = load_nccl();
  if (rc) return rc;
  CUDA_OK(cudaSetDevice(h->device));
  NcclUid id;
  memcpy(&id, id128, sizeof(id));
  int e = g_nccl.CommInitRank(&h->comm, nranks, id, rank);
  if (e) return set_err(PFRX_E_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString(e));
  h->nranks = nranks;
  CUDA_OK(cudaMalloc(&h->d_red, 8 * sizeof(long long)));
  CUDA_OK(cudaMallocHost(&h->h_red, 8 * sizeof(long long)));
  CUDA_OK(cudaMalloc(&h->d_red_step, 8 * sizeof(long long)));
  CUDA_OK(cudaMallocHost(&h->h_red_step, 8 * sizeof(long long)));
  return PFRX_OK;
}

extern "C" int pfrx_allreduce(pfrx_handle *h, pfrx_step_result *r) {
  if (!h || !r) return set_err(PFRX_E_INVALID, "null argument%s", "");
  if (!h->comm) return PFRX_OK;  // single rank: nothing to reduce
  CUDA_OK(cudaSetDevice(h->device));
  if (h->red_valid && r->ncell_active == h->red_local.ncell_active &&
      r->sum_newton_iterations == h->red_local.sum_newton_iterations && r->num_cut_cells == h->red_local.num_cut_cells &&
      r->max_newton_iterations == h->red_local.max_newton_iterations &&
      r->max_num_kinetic_state_updates == h->red_local.max_num_kinetic_state_updates &&
      r->rstep_error == h->red_local.rstep_error && r->max_sub_steps == h->red_local.max_sub_steps) {
    // the result of the latest pfrx_rstep[_async/_finish]: its reduction ran on the device behind the kernel
    const long long *b = h->h_red_step;
    h->red_valid = false;
    r->ncell_active = b[0];
    r->sum_newton_iterations = b[1];
    r->num_cut_cells = b[2];
    r->max_newton_iterations = (int)b[4];
    r->max_num_kinetic_state_updates = (int)b[5];
    r->rstep_error = (int)b[6];
    r->max_sub_steps = (int)b[7];
    return PFRX_OK;
  }
  h->red_valid = false;
  long long *b = h->h_red;
  b[0] = r->ncell_active;
  b[1] = r->sum_newton_iterations;
  b[2] = r->num_cut_cells;
  b[3] = 0;
  b[4] = r->max_newton_iterations;
  b[5] = r->max_num_kinetic_state_updates;
  b[6] = r->rstep_error;
  b[7] = r->max_sub_steps;
  CUDA_OK(cudaMemcpyAsync(h->d_red, b, 8 * sizeof(long long), cudaMemcpyHostToDevice, h->stream));
  const int ncclInt64 = 4, ncclSum = 0, ncclMax = 2;
  g_nccl.GroupStart();
  int e1 = g_nccl.AllReduce(h->d_red, h->d_red, 4, ncclInt64, ncclSum, h->comm, h->stream);
  int e2 = g_nccl.AllReduce(h->d_red + 4, h->d_red + 4, 4, ncclInt64, ncclMax, h->comm, h->stream);
  int e3 = g_nccl.GroupEnd();
  if (e1 || e2 || e3) return set_err(PFRX_E_NCCL, "ncclAllReduce: %s", g_nccl.GetErrorString(e1 ? e1 : (e2 ? e2 : e3)));
  CUDA_OK(cudaMemcpyAsync(b, h->d_red, 8 * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  r->ncell_active = b[0];
  r->sum_newton_iterations = b[1];
  r->num_cut_cells = b[2];
  r->max_newton_iterations = (int)b[4];
  r->max_num_kinetic_state_updates = (int)b[5];
  r->rstep_error = (int)b[6];
  r->max_sub_steps = (int)b[7];
  return PFRX_OK;
}

extern "C" int pfrx_last_transfer_bytes(pfrx_handle *h, int64_t *h2d, int64_t *d2h) {
  if (!h || !h2d || !d2h) return PFRX_E_INVALID;
  *h2d = h->last_h2d;
  *d2h = h->last_d2h;
  return PFRX_OK;
}

extern "C" void *pfrx_stream(pfrx_handle *h) { return h ? (void *)h->stream : nullptr; }
extern "C" int64_t pfrx_launch_count(pfrx_handle *h) { return h ? h->launches : 0; }

extern "C" int pfrx_cell_order(pfrx_handle *h, int mode) {
  if (!h) return set_err(PFRX_E_INVALID, "null handle%s", "");
  h->order_mode = mode != 0;
  h->order_valid = false;
  return PFRX_OK;
}

extern "C" int64_t pfrx_bytes_per_cell(pfrx_handle *h) {
  // every in/io field read once, every io field
  // written once, four int32 results
  if (!h) return 0;
  const int *r = h->rows_d.data();
  const int io[] = {0, 1, 2, 3, 4, 5, 6, 7, 9, 10, 11, 12, 13, 27, 29, 30, 32, 34, 35};
  int64_t in = 0, outn = 0;
  for (int f = 0; f < kNumD; f++)
    if (f != 11) in += r[f];
  for (int f : io) outn += r[f];
  return 8 * (in + outn) + 16;
}

// kernel configuration report: N, lanes, threads, blocks/SM, smem bytes
extern "C" int pfrx_kernel_info(pfrx_handle *h, int *info5) {
  if (!h || !info5) return PFRX_E_INVALID;
  info5[0] = h->npad;
  info5[1] = h->tpc ? 0 : h->lanes;
  info5[2] = h->threads;
  info5[3] = h->blocks_per_sm;
  info5[4] = (int)h->smem_bytes;
  if (h->spec_func) {  // specialised kernel: lanes reported as -1
    info5[0] = h->n;
    info5[1] = -1;
    info5[2] = h->spec_threads;
    info5[3] = h->spec_blocks_per_sm;
    info5[4] = (int)h->spec_smem;
  }
  return PFRX_OK;
}

extern "C" uint64_t pfrx_config_signature(pfrx_handle *h) { return h ? h->sig : 0; }

// Writes the configuration the handle was created from in the text form the code generator reads
// (python -m pflotran_elm_interface_b200.specialize <path> builds the specialised cubins for it):
// the route from a Fortran / C host that flattened reaction_rt_type into pfrx_config to the fast
// kernels, without the Python deck reader.
static int write_text(const std::string &text, const char *path) {
  FILE *f = fopen(path, "w");
  if (!f) return set_err(PFRX_E_INVALID, "cannot open %s for writing", path);
  const size_t n = fwrite(text.data(), 1, text.size(), f);
  const int rc = fclose(f);
  if (n != text.size() || rc != 0) return set_err(PFRX_E_INVALID, "short write to %s", path);
  return PFRX_OK;
}
extern "C" int pfrx_config_dump(pfrx_handle *h, const char *path) {
  if (!h || !path) return set_err(PFRX_E_INVALID, "null argument%s", "");
  return write_text(h->dump_text, path);
}
// the same file straight from a pfrx_config: needs no device (set-up on a build or login node)
extern "C" int pfrx_config_write(const pfrx_config *cfg, const char *path) {
  if (!cfg || !path) return set_err(PFRX_E_INVALID, "null argument%s", "");
  if (cfg->abi_version != PFRX_ABI_VERSION) return set_err(PFRX_E_INVALID, "pfrx_config.abi_version mismatch%s", "");
  return write_text(config_dump_text(cfg), path);
}
extern "C" uint64_t pfrx_config_signature_of(const pfrx_config *cfg) { return cfg ? config_signature(cfg) : 0; }

// Attach a network-specialised kernel (cubin written by specialize.py / nvcc).
// The cubin carries the signature of the tables it was generated from; a cubin
// for another network is refused.  path == NULL detaches.
extern "C" int pfrx_load_specialized(pfrx_handle *h, const char *cubin_path) {
  if (!h) return set_err(PFRX_E_INVALID, "null handle%s", "");
  CUDA_OK(cudaSetDevice(h->device));
  if (h->pending) CUDA_OK(cudaStreamSynchronize(h->stream));
  if (h->spec_module) {
    g_drv.ModuleUnload(h->spec_module);
    h->spec_module = h->spec_func = nullptr;
  }
  if (!cubin_path) return PFRX_OK;
  // features outside what specialize.py generates (its supported() is the twin of this test)
  const DevCfg &d = h->cfg;
  bool inner_newton = false;
  for (int r = 0; r < d.nsrfrxn && h->sr_flag_host.size() == (size_t)d.nsrfrxn; r++)
    inner_newton = inner_newton || h->sr_flag_host[r] != 0;
  if (!d.use_full_geochemistry || (!d.use_isothermal && (d.ncplx > 0 || d.nkin > 0 || d.nsrfcplx > 0)) ||
      d.use_total_as_guess ||
      d.act_alg != PFRX_ACT_COEF_ALGORITHM_LAG || d.mn_npref ||
      ((d.nionx > 0 || d.nkd > 0 || d.ndynkd > 0) && (d.nmr > 0 || d.nsbx > 0)) || (d.nrd > 0 && d.nionx > 0) ||
      (d.nrd > 0 && d.nsrfrxn > 0) ||
      ((d.ngen > 0 || d.nrd > 0 || d.nidc > 0 || d.nmb > 0) && (d.nmr > 0 || d.nsbx > 0)) || d.mn_temkin || d.mn_scale || d.mn_power || inner_newton || d.has_cd || d.has_cs || d.has_rn || d.ngas > 0 ||
      d.nsrfrxn != d.neqsr + d.nmr)
    return set_err(PFRX_E_INVALID, "configuration uses features the specialised kernels do not cover%s", "");
  int rc = load_driver();
  if (rc) return rc;
  CUDA_OK(cudaFree(0));  // make the primary context current for the driver API
  void *mod = nullptr, *fn = nullptr;
  DRV_OK(g_drv.ModuleLoad(&mod, cubin_path));
  unsigned long long dptr = 0;
  size_t bytes = 0;
  unsigned long long sig = 0;
  int info[5] = {0, 0, 0, 0, 0};
  int e = g_drv.ModuleGetGlobal(&dptr, &bytes, mod, "pfrx_spec_sig");
  if (!e && bytes == sizeof(sig)) e = g_drv.MemcpyDtoH(&sig, dptr, sizeof(sig));
  if (!e) e = g_drv.ModuleGetGlobal(&dptr, &bytes, mod, "pfrx_spec_info");
  if (!e && bytes == sizeof(info)) e = g_drv.MemcpyDtoH(info, dptr, sizeof(info));
  if (!e) e = g_drv.ModuleGetFunction(&fn, mod, "pfrx_spec_kernel");
  if (e) {
    g_drv.ModuleUnload(mod);
    return set_err(PFRX_E_CUDA, "not a pfrx specialised cubin (%s): %s", cubin_path, drv_err(e));
  }
  if (sig != h->sig || info[0] != h->n) {
    g_drv.ModuleUnload(mod);
    char a[24], b[24];
    snprintf(a, sizeof(a), "%016llx", sig);
    snprintf(b, sizeof(b), "%016llx", (unsigned long long)h->sig);
    return set_err(PFRX_E_INVALID, "specialised kernel was generated for another network (signature %s, need %s)", a, b);
  }
  int threads = info[2];
  size_t smem = (size_t)info[1] * sizeof(double);
  int nb = 0;
  if (const char *ev = getenv("PFRX_SPEC_MAXBLOCKS")) {
    // diagnostics: limit residency by asking for more shared memory than needed
    int mb = atoi(ev);
    if (mb >= 1) smem = std::max(smem, (size_t)(232448 / mb - 1024) & ~(size_t)1023);
  }
  e = g_drv.FuncSetAttribute(fn, 8 /* CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES */, (int)smem);
  if (!e) e = g_drv.OccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, threads, smem);
  if (!e && info[3] >= 1 && nb > info[3]) {
    // the kernel asks for at most info[3] resident blocks per SM: ask for just enough
    // shared memory that one more does not fit (the rest stays L1 for its spills)
    size_t pad = ((size_t)232448 / (info[3] + 1) - 1024 + 1024) & ~(size_t)255;
    while (pad > smem) {
      int nb2 = 0;
      if (g_drv.FuncSetAttribute(fn, 8, (int)pad)) break;
      if (g_drv.OccupancyMaxActiveBlocksPerMultiprocessor(&nb2, fn, threads, pad)) break;
      if (nb2 <= info[3] && nb2 >= 1) {
        smem = pad;
        nb = nb2;
        break;
      }
      pad += 1024;
      if (pad > 200 * 1024) break;
    }
  }
  if (e || nb < 1) {
    g_drv.ModuleUnload(mod);
    return set_err(PFRX_E_LIMIT, "specialised kernel does not fit on an SM: %s", e ? drv_err(e) : "0 blocks");
  }
  h->spec_module = mod;
  h->spec_func = fn;
  {
    int flags = 0;
    unsigned long long fptr = 0;
    size_t fbytes = 0;
    if (!g_drv.ModuleGetGlobal(&fptr, &fbytes, mod, "pfrx_spec_flags") && fbytes == sizeof(flags))
      g_drv.MemcpyDtoH(&flags, fptr, sizeof(flags));
    h->spec_refill = flags & 1;
    h->order_valid = false;
    if (const char *ev = getenv("PFRX_CELL_ORDER")) h->order_mode = atoi(ev) != 0;
  }
  h->spec_threads = threads;
  h->spec_cells = info[4];
  h->spec_smem = smem;
  h->spec_blocks_per_sm = nb;
  return PFRX_OK;
}

// ---- diagnostics: FP64 FMA peak of the device (roofline denominator) -----------
// 8 independent DFMA chains per thread, enough threads to fill every SM.
__global__ void pfrx_dfma_peak_kernel(double *out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6,
         x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      x0 = fma(x0, a, b);
      x1 = fma(x1, a, b);
      x2 = fma(x2, a, b);
      x3 = fma(x3, a, b);
      x4 = fma(x4, a, b);
      x5 = fma(x5, a, b);
      x6 = fma(x6, a, b);
      x7 = fma(x7, a, b);
    }
  }
  double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 12345.678) out[0] = s;  // keep the chains alive
}

extern "C" int pfrx_diag_fp64_peak(int device, double *tflops, double *sm_mhz_est) {
  CUDA_OK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, device));
  double *d = nullptr;
  CUDA_OK(cudaMalloc(&d, 8));
  const int threads = 256, blocks = prop.multiProcessorCount * 8, iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 6; rep++) {
    cudaEventRecord(e0);
    pfrx_dfma_peak_kernel<<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 64.0 * (double)iters * threads * blocks;
    double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *tflops = best;
  if (sm_mhz_est) *sm_mhz_est = best * 1e12 / (2.0 * 64.0 * prop.multiProcessorCount) / 1e6;
  return PFRX_OK;
}
