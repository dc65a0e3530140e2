// shell_gpu_compat.cuh -- the reference's OWN C ABI for its `-gpu` shell path, exported by liborgpu.so.
//
// Every entry point of engine/source/elements/shell/coque/shell_gpu_driver.h:44-206 (bound from Fortran by
// shell_gpu_mod.F90:284-715 and called by shell_internal_forces.F90: FORINTC_PREPARE_GPU :370, gpu_shell_launch_async
// :62, gpu_shell_sync_scatter :177) with the same name, argument list and meaning, so that an Engine built WITH_CUDA
// links this library in place of its four .cu files without touching the Fortran side.  Behind it runs the path of
// this library: the super-units (one ShellGPUData each) become shell groups of ONE orgpu engine on the global handle,
// built lazily at the first per-cycle call; a cycle is one forces phase + one deterministic /PARITH/ON assembly, packed
// into the caller's [Fx|Fy|Fz|Mx|My|Mz|STIFN|STIFR] x NUMNOD buffer.
//
// Differences a caller can see:
//   * the through-thickness rule, the strain-rate filter and the element time step are the CPU Engine's (CMAIN3/MULAWC
//     Gauss-Lobatto tables, min(1, PM(9)*DT1), CDT3 with VISCMX/ALPE) -- the reference's kernels use a mid-point rule, take
//     ASRATE as the filter coefficient itself and leave VISCMX out; on flat membrane states they coincide
//     (tests/test_shell_gpu_abi.py);
//   * nodal sums are deterministic (no atomics);
//   * unsupported requests (FISOKIN>0, Ishell not 1/3/4, NPT=1, Ismstr not 1/2/4, HVISC/HELAS/HVLIN other than the
//     Engine's 0.5/0.5/0, per-element material arrays that vary inside a super-unit) print the reason and exit(1), as the
//     reference's CUDA_CHECK does (shell_gpu_driver.cu:47-55);
//   * shell_gpu_min_dt returns the minimum over ALL super-units of the handle (the caller takes that minimum anyway);
//   * shell_gpu_download_aldt_sq[_async] (superseded by shell_gpu_min_dt in the reference itself and not called by
//     shell_internal_forces.F90) is not provided: it exits with a message.
#pragma once
#include "../../include/shell_gpu_abi.h"

struct ShellGPUData;
struct ShellGPUGlobal {
  int NUMNOD = 0;
  orgpu_engine* e = nullptr;
  std::vector<ShellGPUData*> sus;
  bool ran = false;               // the cycle's forces are in d_soa
  double* d_soa = nullptr;        // [8][NUMNOD]
  int numelc = 0;
  std::vector<double> xfer;       // state transfers
};
struct ShellGPUData {
  ShellGPUGlobal* gh = nullptr; bool own_global = false;
  int NUMELC = 0, NUMNOD = 0, NPT = 0, ISMSTR = 0, ITHK = 0, compute_sti = 2, ihbe = 1, nft = 0;
  bool have_mat = false, have_hg = false, have_const = false;
  orgpu_law2 mat{}; orgpu_prop_shell prop{};
  std::vector<int> n[4];
  std::vector<double> thk0, off, ip[7], temp;   // ip: SIGxx, yy, xy, yz, zx, PLA, EPSD_ip
};

static void sg_die(const char* what)
{
  fprintf(stderr, "liborgpu shell_gpu ABI: %s%s%s\n", what, orgpu_last_error()[0] ? ": " : "", orgpu_last_error());
  exit(EXIT_FAILURE);
}
#define SG_OK(call) do { if ((call) < 0) sg_die(#call); } while (0)   /* orgpu_add_*_group return the group index */
#define SG_CUDA(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { fprintf(stderr, "liborgpu shell_gpu ABI: %s: %s\n", #call, cudaGetErrorString(_e)); exit(EXIT_FAILURE); } } while (0)

__global__ void sg_pack_soa_kernel(const double* __restrict__ A, const double* __restrict__ AR, const double* __restrict__ STIFN,
                                   const double* __restrict__ STIFR, double* __restrict__ out, int n, int with_sti)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
  out[i] = A[3 * i]; out[n + i] = A[3 * i + 1]; out[2 * n + i] = A[3 * i + 2];
  out[3 * n + i] = AR ? AR[3 * i] : 0.0; out[4 * n + i] = AR ? AR[3 * i + 1] : 0.0; out[5 * n + i] = AR ? AR[3 * i + 2] : 0.0;
  out[6 * n + i] = with_sti ? STIFN[i] : 0.0; out[7 * n + i] = (with_sti && STIFR) ? STIFR[i] : 0.0;
}

// FORINTC_PREPARE_GPU has registered every super-unit by now: one engine, one skyline, one group list
static void sg_build(ShellGPUGlobal* gh)
{
  if (gh->e) return;
  if (gh->sus.empty()) sg_die("no super-unit was attached to the global handle (shell_gpu_set_global)");
  int dev = 0; SG_CUDA(cudaGetDevice(&dev));
  orgpu_control ctl{};
  ctl.dtfac_brick = ctl.dtfac_shell = ctl.dtfac_node = ctl.dtfac_sh3n = 1.0;   // shell_gpu_min_dt applies the caller's DTFAC1(3)
  ctl.dtmx = K_EP20; ctl.iroddl = 1;
  ctl.nodadt = (gh->sus[0]->compute_sti == 1) ? 1 : 0;                         // shell_internal_forces.F90:637-648
  int ne = 0;
  for (ShellGPUData* g : gh->sus) {
    if (!g->have_mat || !g->have_hg || !g->have_const) sg_die("super-unit used before shell_gpu_set_mat_params / set_hg_params / upload_constant");
    if (((g->compute_sti == 1) ? 1 : 0) != ctl.nodadt) sg_die("super-units with different compute_sti modes on one handle");
    if (g->NUMNOD != gh->NUMNOD) sg_die("super-unit NUMNOD differs from the global handle's");
    g->nft = ne; ne += g->NUMELC;
  }
  gh->numelc = ne;
  SG_OK(orgpu_create(&gh->e, dev, gh->NUMNOD, &ctl));
  std::vector<int> ixc((size_t)7 * ne), iadc((size_t)4 * ne), adsky;
  int su = 0;
  for (ShellGPUData* g : gh->sus) {
    su++;
    for (int i = 0; i < g->NUMELC; i++) {
      int* r = &ixc[(size_t)7 * (g->nft + i)];
      r[0] = su; r[5] = su; r[6] = g->nft + i + 1;
      for (int k = 0; k < 4; k++) {
        const int nd = g->n[k][i];
        if (nd < 0 || nd >= gh->NUMNOD) sg_die("connectivity out of range (0-based node numbers expected, shell_internal_forces.F90:900-905)");
        r[1 + k] = nd + 1;
      }
    }
  }
  // node -> corner rows (ADSKY, 1-based; the slots of a node in element order, as FILLCNE orders them by user id)
  {
    std::vector<int> cnt((size_t)gh->NUMNOD, 0), start((size_t)gh->NUMNOD + 1, 0);
    for (int e = 0; e < ne; e++) for (int k = 0; k < 4; k++) cnt[ixc[(size_t)7 * e + 1 + k] - 1]++;
    start[0] = 1; for (int n = 0; n < gh->NUMNOD; n++) start[n + 1] = start[n] + cnt[n];
    adsky.assign(start.begin(), start.end());
    std::vector<int> next(start.begin(), start.end() - 1);
    for (int e = 0; e < ne; e++) for (int k = 0; k < 4; k++) iadc[(size_t)4 * e + k] = next[ixc[(size_t)7 * e + 1 + k] - 1]++;
  }
  SG_OK(orgpu_set_shells(gh->e, ne, ixc.data(), iadc.data()));
  SG_OK(orgpu_set_pon(gh->e, adsky.data(), 4 * ne));
  for (ShellGPUData* g : gh->sus) {
    int i = 0;
    while (i < g->NUMELC) {                    // IPARG groups: <= 128 elements; ITHK = 0 keeps the initial thickness per group
      int j = i + 1;
      while (j < g->NUMELC && j - i < 128 && (g->ITHK > 0 || g->thk0[j] == g->thk0[i])) j++;
      orgpu_prop_shell p = g->prop; p.thick = g->ITHK > 0 ? g->thk0[0] : g->thk0[i];   // ITHK > 0: GBUF%THK is state (uploaded below), one property so the groups fuse
      SG_OK(orgpu_add_shell_group(gh->e, j - i, g->nft + i, 2, &g->mat, &p));
      i = j;
    }
  }
  SG_OK(orgpu_finalize(gh->e));
  // initial state: OFF, THK, per-point stresses / plastic strain / strain rate / temperature
  int nptmax = 1; for (ShellGPUData* g : gh->sus) nptmax = g->NPT > nptmax ? g->NPT : nptmax;
  std::vector<double>& b = gh->xfer; b.assign((size_t)5 * nptmax * ne, 0.0);
  auto put = [&](int field) { SG_OK(orgpu_upload_shell_state(gh->e, field, b.data())); };
  for (ShellGPUData* g : gh->sus) for (int i = 0; i < g->NUMELC; i++) b[g->nft + i] = g->off[i];
  put(4);
  for (ShellGPUData* g : gh->sus) for (int i = 0; i < g->NUMELC; i++) b[g->nft + i] = g->thk0[i];
  put(3);
  bool have_ip = false; for (ShellGPUData* g : gh->sus) have_ip = have_ip || !g->ip[0].empty();
  if (have_ip) {
    std::fill(b.begin(), b.end(), 0.0);
    for (ShellGPUData* g : gh->sus) if (!g->ip[0].empty())
      for (int t = 0; t < g->NPT; t++) for (int c = 0; c < 5; c++) for (int i = 0; i < g->NUMELC; i++)
        b[(size_t)(5 * t + c) * ne + g->nft + i] = g->ip[c][(size_t)t * g->NUMELC + i];
    put(9);
    for (int f = 0; f < 2; f++) {
      for (ShellGPUData* g : gh->sus) if (!g->ip[0].empty())
        for (int t = 0; t < g->NPT; t++) for (int i = 0; i < g->NUMELC; i++)
          b[(size_t)t * ne + g->nft + i] = g->ip[5 + f][(size_t)t * g->NUMELC + i];
      put(10 + f);
    }
    bool any_temp = false;
    for (ShellGPUData* g : gh->sus) if (g->mat.has_temp && !g->temp.empty()) {
      any_temp = true;
      for (int t = 0; t < g->NPT; t++) for (int i = 0; i < g->NUMELC; i++) b[(size_t)t * ne + g->nft + i] = g->temp[(size_t)t * g->NUMELC + i];
    }
    if (any_temp) put(12);
  }
  SG_CUDA(cudaMalloc((void**)&gh->d_soa, sizeof(double) * 8 * (size_t)gh->NUMNOD));
  SG_CUDA(cudaMemset(gh->d_soa, 0, sizeof(double) * 8 * (size_t)gh->NUMNOD));
  for (ShellGPUData* g : gh->sus) { for (auto& v : g->n) std::vector<int>().swap(v); for (auto& v : g->ip) std::vector<double>().swap(v); }
}

static void sg_run(ShellGPUGlobal* gh, double dt)
{
  sg_build(gh);
  if (gh->ran) return;                           // the first super-unit of the cycle computes all of them
  orgpu_engine* e = gh->e;
  SG_OK(orgpu_forces_phase(e, dt));
  SG_OK(orgpu_assemble(e));
  const int n = gh->NUMNOD;
  sg_pack_soa_kernel<<<(n + 255) / 256, 256, 0, e->st>>>(e->nd.A, e->nd.AR, e->nd.STIFN, e->nd.STIFR, gh->d_soa, n, gh->sus[0]->compute_sti != 0);
  e->launches++;
  SG_CUDA(cudaGetLastError());
  gh->ran = true;
}

static ShellGPUGlobal* sg_global_of(ShellGPUData* g)
{
  if (!g->gh) {                                  // legacy per-SU use without a global handle: a private one
    g->gh = new ShellGPUGlobal(); g->gh->NUMNOD = g->NUMNOD; g->gh->sus.push_back(g); g->own_global = true;
  }
  return g->gh;
}

static void sg_fetch(ShellGPUGlobal* gh, int field, int ncomp)
{
  sg_build(gh);
  if (gh->xfer.size() < (size_t)ncomp * gh->numelc) gh->xfer.resize((size_t)ncomp * gh->numelc);
  SG_OK(orgpu_download_shell_state(gh->e, field, gh->xfer.data()));
}

extern "C" {

ShellGPUGlobal* shell_gpu_global_create(int NUMNOD)
{
  ShellGPUGlobal* gh = new ShellGPUGlobal(); gh->NUMNOD = NUMNOD; return gh;
}
void shell_gpu_global_destroy(ShellGPUGlobal* gh)
{
  if (!gh) return;
  if (gh->d_soa) cudaFree(gh->d_soa);
  if (gh->e) orgpu_destroy(gh->e);
  for (ShellGPUData* g : gh->sus) if (g->gh == gh) g->gh = nullptr;
  delete gh;
}
void shell_gpu_global_upload_nodes(ShellGPUGlobal* gh, const Real* X, const Real* V, const Real* VR)
{
  sg_build(gh);
  SG_OK(orgpu_upload_nodes(gh->e, X, V, VR, nullptr, nullptr, nullptr));
  gh->ran = false;                               // the reference zeroes its accumulators here
}
void shell_gpu_global_download_forces(ShellGPUGlobal* gh, Real* raw_gpu_to_cpu)
{
  sg_build(gh);
  SG_CUDA(cudaMemcpyAsync(raw_gpu_to_cpu, gh->d_soa, sizeof(double) * 8 * (size_t)gh->NUMNOD, cudaMemcpyDeviceToHost, gh->e->st));
}
void shell_gpu_global_synchronize(ShellGPUGlobal* gh) { if (gh && gh->e) SG_CUDA(cudaStreamSynchronize(gh->e->st)); }
void shell_gpu_global_wait_upload(ShellGPUGlobal*, ShellGPUData*) {}      // one stream: already ordered
void shell_gpu_global_wait_su(ShellGPUGlobal*, ShellGPUData*) {}
void shell_gpu_global_pin_host(const Real* X, const Real* V, const Real* VR, Real* raw_gpu_to_cpu, int NUMNOD)
{
  const size_t s3 = sizeof(Real) * 3 * (size_t)NUMNOD, s8 = sizeof(Real) * 8 * (size_t)NUMNOD;
  // the reference ignores "already registered" the same way (shell_gpu_driver.cu:219-232)
  cudaHostRegister((void*)X, s3, cudaHostRegisterDefault); cudaHostRegister((void*)V, s3, cudaHostRegisterDefault);
  cudaHostRegister((void*)VR, s3, cudaHostRegisterDefault); cudaHostRegister((void*)raw_gpu_to_cpu, s8, cudaHostRegisterDefault);
  cudaGetLastError();
}
void shell_gpu_set_global(ShellGPUData* g, ShellGPUGlobal* gh)
{
  if (gh->e) sg_die("shell_gpu_set_global after the first cycle: all super-units must be attached before");
  g->gh = gh; gh->sus.push_back(g);
}

ShellGPUData* shell_gpu_data_create(void) { return new ShellGPUData(); }
void shell_gpu_data_destroy(ShellGPUData* g)
{
  if (!g) return;
  if (g->own_global && g->gh) { ShellGPUGlobal* gh = g->gh; g->gh = nullptr; gh->sus.clear(); shell_gpu_global_destroy(gh); }
  else if (g->gh) { auto& v = g->gh->sus; for (size_t k = 0; k < v.size(); k++) if (v[k] == g) { v.erase(v.begin() + k); break; } }
  delete g;
}
void shell_gpu_allocate(ShellGPUData* g, int NUMELC, int NUMNOD, int NPT, int ISMSTR, int ITHK)
{
  if (NUMELC <= 0 || NUMNOD <= 0) sg_die("shell_gpu_allocate: empty super-unit");
  if (NPT < 2 || NPT > 10) sg_die("shell_gpu_allocate: NPT outside the built path (2..10; Belytschko-Tsay with NPT=1 uses MHVIS3)");
  if (!(ISMSTR == 1 || ISMSTR == 2 || ISMSTR == 4)) sg_die("shell_gpu_allocate: Ismstr outside the built path (1, 2, 4)");
  g->NUMELC = NUMELC; g->NUMNOD = NUMNOD; g->NPT = NPT; g->ISMSTR = ISMSTR; g->ITHK = ITHK;
  g->prop.npt = NPT; g->prop.ismstr = ISMSTR; g->prop.ithk = ITHK; g->prop.istrain = 1; g->prop.ihbe = g->ihbe;
}
void shell_gpu_deallocate(ShellGPUData* g)
{
  if (!g) return;
  for (auto& v : g->n) std::vector<int>().swap(v);
  for (auto& v : g->ip) std::vector<double>().swap(v);
  std::vector<double>().swap(g->thk0); std::vector<double>().swap(g->off); std::vector<double>().swap(g->temp);
  if (g->own_global && g->gh) { ShellGPUGlobal* gh = g->gh; g->gh = nullptr; g->own_global = false; gh->sus.clear(); shell_gpu_global_destroy(gh); }
}

void shell_gpu_set_mat_params(ShellGPUData* g, Real E, Real nu, Real G, Real A11, Real A12, Real CA, Real CB, Real CN, Real CC, Real EPDR,
                              Real EPMX, Real YMAX, Real M_EXP, Real FISOKIN, Real RHOCP, Real TREF, Real TMELT, Real ASRATE,
                              Real RHO, Real SSP, Real SHF_COEF, int IPLA, int VP, int IFORM, int ICC, Real Z3, Real Z4)
{
  if (FISOKIN != 0.0) sg_die("LAW2 kinematic hardening (FISOKIN>0) is outside the built path");
  if (IPLA < 0 || IPLA > 2) sg_die("Iplas outside 0..2");
  orgpu_law2& m = g->mat;
  m.rho0 = RHO; m.young = E; m.nu = nu; m.shear = G; m.bulk = E / (3.0 * (1.0 - 2.0 * nu));
  m.ca = CA; m.cb = CB; m.cn = CN; m.epmx = EPMX; m.sigmx = YMAX; m.cc = CC; m.epdr = EPDR; m.fisokin = 0.0;
  m.asrate = ASRATE;                                  // PM(9) = 2 pi Fcut as shell_internal_forces.F90:715 passes it
  m.israte = ASRATE > 0.0 ? 1 : 0;
  m.z3 = IFORM == 1 ? Z3 : M_EXP; m.z4 = Z4;          // uparam(10), uparam(11) (shell_internal_forces.F90:693-706)
  m.tref = TREF; m.tmelt = TMELT; m.rhocp = RHOCP; m.pshift = 0.0;
  m.a11 = A11; m.a12 = A12; m.ssp = SSP;
  m.gsr = sqrt(fmax(0.0, G)); m.a11sr = sqrt(fmax(0.0, A11)); m.a12sr = sqrt(fmax(0.0, A12)); m.nusr = sqrt(fmax(0.0, nu));
  m.iform = IFORM; m.icc = ICC; m.vp = VP; m.has_temp = RHOCP > 0.0 ? 1 : 0;
  if (m.tini == 0.0) m.tini = TREF;                   // until shell_gpu_upload_ip_state brings TEMPEL
  g->prop.shf = SHF_COEF; g->prop.shfsr = sqrt(fmax(0.0, SHF_COEF)); g->prop.ipla = IPLA;
  g->have_mat = true;
}
void shell_gpu_set_hg_params(ShellGPUData* g, Real H1, Real H2, Real H3, Real SRH1, Real SRH2, Real SRH3, Real HVISC, Real HELAS, Real HVLIN)
{
  if (HVISC != K_HALF || HELAS != K_HALF || HVLIN != K_ZERO) sg_die("HVISC / HELAS / HVLIN other than the Engine's 0.5 / 0.5 / 0 (radioss2.F:641-643)");
  g->prop.h1 = H1; g->prop.h2 = H2; g->prop.h3 = H3; g->prop.srh1 = SRH1; g->prop.srh2 = SRH2; g->prop.srh3 = SRH3;
  g->prop.cvis = 0.0; g->prop.dm = 0.0;
  g->have_hg = true;
}
void shell_gpu_set_compute_sti(ShellGPUData* g, int flag) { g->compute_sti = flag; }
void shell_gpu_set_ihbe(ShellGPUData* g, int ihbe)
{
  if (ihbe == 0) ihbe = 1;
  if (!(ihbe == 1 || ihbe == 3 || ihbe == 4)) sg_die("Ishell outside the built Belytschko-Tsay path (1, 3, 4)");
  g->ihbe = ihbe; g->prop.ihbe = ihbe;
}

void shell_gpu_upload_constant(ShellGPUData* g, const int* h_N1, const int* h_N2, const int* h_N3, const int* h_N4, const Real* h_THK0,
                               const Real* h_OFF, const Real* h_SSP, const Real* h_RHO, const Real* h_YM, const Real* h_NU, const Real* h_A11,
                               const Real* h_G, const Real* h_SHF)
{
  if (g->NUMELC <= 0) sg_die("shell_gpu_upload_constant before shell_gpu_allocate");
  if (g->gh && g->gh->e) sg_die("shell_gpu_upload_constant after the first cycle");
  const int ne = g->NUMELC;
  const int* nn[4] = {h_N1, h_N2, h_N3, h_N4};
  for (int k = 0; k < 4; k++) g->n[k].assign(nn[k], nn[k] + ne);
  g->thk0.assign(h_THK0, h_THK0 + ne); g->off.assign(h_OFF, h_OFF + ne);
  if (g->have_mat) {                                  // one material per super-unit (FORINTC_PREPARE_GPU fills these from it)
    const orgpu_law2& m = g->mat;
    for (int i = 0; i < ne; i++)
      if (h_SSP[i] != m.ssp || h_RHO[i] != m.rho0 || h_YM[i] != m.young || h_NU[i] != m.nu || h_A11[i] != m.a11 || h_G[i] != m.shear || h_SHF[i] != g->prop.shf)
        sg_die("shell_gpu_upload_constant: per-element material arrays differ from shell_gpu_set_mat_params (one material per super-unit)");
  }
  g->have_const = true;
}
void shell_gpu_upload_ip_state(ShellGPUData* g, const Real* h_SIGxx, const Real* h_SIGyy, const Real* h_SIGxy, const Real* h_SIGyz, const Real* h_SIGzx,
                               const Real* h_PLA, const Real* h_EPSD_ip, const Real* h_SIGBAKxx, const Real* h_SIGBAKyy, const Real* h_SIGBAKxy,
                               const Real* h_TEMPEL)
{
  (void)h_SIGBAKxx; (void)h_SIGBAKyy; (void)h_SIGBAKxy;             // back stresses: FISOKIN = 0 only
  if (g->gh && g->gh->e) sg_die("shell_gpu_upload_ip_state after the first cycle (use a restart of the whole handle)");
  const size_t nip = (size_t)g->NPT * g->NUMELC;
  const Real* src[7] = {h_SIGxx, h_SIGyy, h_SIGxy, h_SIGyz, h_SIGzx, h_PLA, h_EPSD_ip};
  for (int k = 0; k < 7; k++) g->ip[k].assign(src[k], src[k] + nip);
  g->temp.assign(h_TEMPEL, h_TEMPEL + nip);
  if (nip) g->mat.tini = h_TEMPEL[0];
}

void shell_gpu_upload_nodes(ShellGPUData* g, const Real* X, const Real* V, const Real* VR)
{
  shell_gpu_global_upload_nodes(sg_global_of(g), X, V, VR);
}
void shell_gpu_download_nodal_forces(const ShellGPUData* g, Real* raw_gpu_to_cpu)
{
  ShellGPUGlobal* gh = sg_global_of(const_cast<ShellGPUData*>(g));
  shell_gpu_global_download_forces(gh, raw_gpu_to_cpu); shell_gpu_global_synchronize(gh);
}
void shell_gpu_download_energy(const ShellGPUData* g, Real* h_EINT)
{
  ShellGPUGlobal* gh = sg_global_of(const_cast<ShellGPUData*>(g));
  sg_fetch(gh, 2, 2);
  for (int k = 0; k < 2; k++) for (int i = 0; i < g->NUMELC; i++) h_EINT[(size_t)k * g->NUMELC + i] = gh->xfer[(size_t)k * gh->numelc + g->nft + i];
}
void shell_gpu_download_state(const ShellGPUData* g, Real* h_OFF, Real* h_THK, Real* h_GSTR, Real* h_EPSD_elem, Real* h_SIGxx, Real* h_SIGyy,
                              Real* h_SIGxy, Real* h_SIGyz, Real* h_SIGzx, Real* h_PLA, Real* h_EPSD_ip, Real* h_SIGBAKxx, Real* h_SIGBAKyy,
                              Real* h_SIGBAKxy, Real* h_TEMPEL)
{
  ShellGPUGlobal* gh = sg_global_of(const_cast<ShellGPUData*>(g));
  const int ne = g->NUMELC, npt = g->NPT;
  auto slice = [&](int field, int ncomp, Real* out) {
    sg_fetch(gh, field, ncomp);
    for (int k = 0; k < ncomp; k++) for (int i = 0; i < ne; i++) out[(size_t)k * ne + i] = gh->xfer[(size_t)k * gh->numelc + g->nft + i];
  };
  slice(4, 1, h_OFF); slice(3, 1, h_THK); slice(5, 8, h_GSTR); slice(6, 1, h_EPSD_elem);
  sg_fetch(gh, 9, 5 * npt);
  Real* sig[5] = {h_SIGxx, h_SIGyy, h_SIGxy, h_SIGyz, h_SIGzx};
  for (int t = 0; t < npt; t++) for (int c = 0; c < 5; c++) for (int i = 0; i < ne; i++)
    sig[c][(size_t)t * ne + i] = gh->xfer[(size_t)(5 * t + c) * gh->numelc + g->nft + i];
  slice(10, npt, h_PLA); slice(11, npt, h_EPSD_ip);
  const size_t nip = (size_t)npt * ne;
  for (size_t k = 0; k < nip; k++) { h_SIGBAKxx[k] = 0.0; h_SIGBAKyy[k] = 0.0; h_SIGBAKxy[k] = 0.0; }
  if (g->mat.has_temp) slice(12, npt, h_TEMPEL);
  else for (size_t k = 0; k < nip; k++) h_TEMPEL[k] = g->mat.tini;
}

void shell_gpu_zero_nodal_arrays(ShellGPUData* g) { sg_global_of(g)->ran = false; }
void shell_gpu_run_kernels(ShellGPUData* g, Real dt) { sg_run(sg_global_of(g), dt); }
void shell_gpu_synchronize(ShellGPUData* g) { if (g && g->gh) shell_gpu_global_synchronize(g->gh); }
void shell_gpu_pin_host_memory(const ShellGPUData* g, Real* raw_cpu_to_gpu, Real* raw_gpu_to_cpu)
{
  cudaHostRegister(raw_cpu_to_gpu, sizeof(Real) * 9 * (size_t)g->NUMNOD, cudaHostRegisterDefault);
  cudaHostRegister(raw_gpu_to_cpu, sizeof(Real) * 8 * (size_t)g->NUMNOD, cudaHostRegisterDefault);
  cudaGetLastError();
}
void shell_gpu_unpin_host_memory(Real* raw_cpu_to_gpu, Real* raw_gpu_to_cpu)
{
  cudaHostUnregister(raw_cpu_to_gpu); cudaHostUnregister(raw_gpu_to_cpu); cudaGetLastError();
}
void shell_gpu_full_step(ShellGPUData* g, Real dt, const Real* raw_cpu_to_gpu, Real* raw_gpu_to_cpu)
{
  const size_t n3 = 3 * (size_t)g->NUMNOD;                          // [X | V | VR], each (3, NUMNOD) (shell_gpu_driver.cu:1095-1112)
  shell_gpu_upload_nodes(g, raw_cpu_to_gpu, raw_cpu_to_gpu + n3, raw_cpu_to_gpu + 2 * n3);
  shell_gpu_run_kernels(g, dt);
  shell_gpu_download_nodal_forces(g, raw_gpu_to_cpu);
}
void shell_gpu_full_step_async(ShellGPUData* g, Real dt, const Real* X, const Real* V, const Real* VR, Real* raw_gpu_to_cpu)
{
  (void)raw_gpu_to_cpu;                                             // the global handle downloads (shell_gpu_driver.cu:1138-1165)
  if (g->own_global || !g->gh) shell_gpu_upload_nodes(g, X, V, VR);
  shell_gpu_run_kernels(g, dt);
}
void shell_gpu_download_aldt_sq_async(const ShellGPUData*, Real*) { sg_die("shell_gpu_download_aldt_sq_async is not provided: use shell_gpu_min_dt"); }
void shell_gpu_download_aldt_sq(const ShellGPUData*, Real*) { sg_die("shell_gpu_download_aldt_sq is not provided: use shell_gpu_min_dt"); }
void shell_gpu_min_dt(ShellGPUData* g, Real dtfac, Real* h_dt_min)
{
  ShellGPUGlobal* gh = sg_global_of(g);
  if (!gh->e || g->compute_sti != 2) { *h_dt_min = K_EP30; return; }   // no element time step in the other modes
  double t[5]; int it[3];
  SG_OK(orgpu_get_time(gh->e, t, it));
  *h_dt_min = dtfac * t[4];
}

} // extern "C"
