// shell_kernel.cuh -- host side of the 4-node shell path: group registration, the
// FORINTC_PREPARE_GPU-style fusion of consecutive compatible groups into super-groups
// (shell_internal_forces.F90:547-558, 829-886), ELBUF -> device SoA, kernel dispatch
// (forintc.F:383-385 QEPH -> CZFORC3, :452 BT -> CFORC3) and state read-back
// (shell_gpu_download_state, shell_gpu_driver.cu:743).
#pragma once
#include <vector>
#include <cstring>
#include <cstdlib>
#include "common.cuh"
#include "shell_common.cuh"
#include "qeph_kernel.cuh"
#include "bt_kernel.cuh"
#include "c3_kernel.cuh"

struct HostShellGroup { int nel, nft, law, sh3n; orgpu_prop_shell prop; orgpu_law2 m2; orgpu_law36 m36; int part; orgpu_fail fail; };
struct ShellSGHost { ShellSG d; int first_elem = 0; bool sh3n = false; int part = 0; std::vector<void*> owned; };

static inline bool shell_is_qeph(const orgpu_prop_shell& p) { return p.ihbe >= 21 && p.ihbe <= 29; }

static int shell_add_group(std::vector<HostShellGroup>& groups, int nel, int nft, int law, const void* mat,
                           const orgpu_prop_shell* prop, bool sh3n = false)
{
  if (law != 2 && law != 36) { orgpu_set_error("shell law %d is outside the built path (2, 36)", law); return -5; }
  if (prop->npt < 1 || prop->npt > 10) { orgpu_set_error("NPT=%d is outside the built path (1..10)", prop->npt); return -5; }
  if (prop->ipla < 0 || prop->ipla > 2) { orgpu_set_error("Iplas=%d is outside the built path (0,1,2)", prop->ipla); return -5; }
  if (!(prop->ismstr == 1 || prop->ismstr == 2 || prop->ismstr == 4)) { orgpu_set_error("shell Ismstr=%d is outside the built path (1,2,4)", prop->ismstr); return -5; }
  const bool qeph = !sh3n && shell_is_qeph(*prop);
  if (sh3n) {
    if (!(prop->ihbe == 1 || prop->ihbe == 2)) { orgpu_set_error("Ish3n=%d is outside the built path (1, 2: C3FORC3)", prop->ihbe); return -5; }
  } else {
    if (!qeph && !(prop->ihbe == 1 || prop->ihbe == 3 || prop->ihbe == 4)) { orgpu_set_error("Ishell=%d is outside the built path (BT 1, 3, 4; QEPH 24)", prop->ihbe); return -5; }
    if (!qeph && prop->npt == 1) { orgpu_set_error("Belytschko-Tsay with NPT=1 (MHVIS3 hourglass) is outside the built path"); return -5; }
  }
  HostShellGroup g; memset(&g, 0, sizeof g);
  g.nel = nel; g.nft = nft; g.law = law; g.sh3n = sh3n ? 1 : 0; g.prop = *prop;
  if (law == 36) {
    g.m36 = *(const orgpu_law36*)mat;
    if (g.m36.fisokin < 0.0 || g.m36.fisokin > 1.0 || g.m36.vp < 0 || g.m36.vp > 1 || g.m36.ifail < 0 || g.m36.ifail > 2) { orgpu_set_error("LAW36 VP outside {0,1} / FISOKIN outside [0,1] are outside the built path"); return -5; }
    if (g.m36.vp == 1 && g.m36.nrate < 2) { orgpu_set_error("LAW36 VP=1 needs more than one curve (the Starter resets VP to 0 for a single curve, hm_read_mat36.F:199)"); return -5; }
    if (g.m36.ifail == 2 && prop->istrain == 0) { orgpu_set_error("LAW36 tensile-strain failure (IFAIL=2) needs the total strains (Istrain=1)"); return -5; }
    if (g.m36.nrate < 1 || g.m36.nrate > ORGPU_MAXFUNC36) { orgpu_set_error("LAW36 NRATE=%d out of range", g.m36.nrate); return -5; }
  } else {
    g.m2 = *(const orgpu_law2*)mat;
    if (g.m2.fisokin < 0.0 || g.m2.fisokin > 1.0) { orgpu_set_error("LAW2 FISOKIN outside [0,1]"); return -5; }
  }
  groups.push_back(g);
  return (int)groups.size() - 1;
}

template <class T> static int sh_alloc(std::vector<void*>& owned, T** p, size_t n) {
  *p = nullptr; if (n == 0) return 0;
  if (cudaMalloc((void**)p, n * sizeof(T)) != cudaSuccess) { orgpu_set_error("cudaMalloc of %zu bytes failed", n * sizeof(T)); return -100; }
  cudaMemset(*p, 0, n * sizeof(T)); owned.push_back(*p); return 0;
}
template <class T> static int sh_upload(std::vector<void*>& owned, T** p, const std::vector<T>& h) {
  if (sh_alloc(owned, p, h.size())) return -100;
  if (h.size() && cudaMemcpy(*p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) { orgpu_set_error("upload failed"); return -100; }
  return 0;
}

// nnode = 4: IXC(7,*) / IADC(4,*); nnode = 3: IXTG(6,*) / IADTG(3,*) (3-node shells, C3FORC3)
static int shell_build_supergroups(std::vector<HostShellGroup>& groups, std::vector<ShellSGHost>& out,
                                   const std::vector<int>& ixc, const std::vector<int>& iadc, const int nnode,
                                   const std::vector<int>& npf, const std::vector<double>& tf,
                                   const orgpu_control& ctl, int numnod, int lsky, int& order, int& blk, std::vector<SGRange>& sgr)
{
  if (groups.empty()) return 0;
  if (!ctl.iroddl) { orgpu_set_error("shell groups need rotational dofs (control.iroddl=1)"); return -4; }
  cudaMemcpyToSymbol(c_Z0, OR_Z0, sizeof OR_Z0); cudaMemcpyToSymbol(c_WF, OR_WF, sizeof OR_WF); cudaMemcpyToSymbol(c_WM, OR_WM, sizeof OR_WM);
  // one device copy of the function table, shared by every LAW36 super-group
  const double* d_tf = nullptr; const int* d_npf = nullptr;
  size_t gi = 0;
  while (gi < groups.size()) {
    size_t gj = gi + 1;
    while (gj < groups.size() && groups[gj].nft == groups[gj - 1].nft + groups[gj - 1].nel && groups[gj].law == groups[gi].law &&
           !memcmp(&groups[gj].prop, &groups[gi].prop, sizeof(orgpu_prop_shell)) &&
           !memcmp(&groups[gj].m2, &groups[gi].m2, sizeof(orgpu_law2)) && !memcmp(&groups[gj].m36, &groups[gi].m36, sizeof(orgpu_law36)) &&
           !memcmp(&groups[gj].fail, &groups[gi].fail, sizeof(orgpu_fail)) && groups[gj].part == groups[gi].part) gj++;
    int ne = 0; for (size_t k = gi; k < gj; k++) ne += groups[k].nel;
    const HostShellGroup& G = groups[gi];
    const int nft = G.nft;
    const int np = ((ne + ORGPU_BLOCK - 1) / ORGPU_BLOCK) * ORGPU_BLOCK;
    out.emplace_back(); ShellSGHost& S = out.back(); S.first_elem = nft; S.sh3n = (nnode == 3); S.part = G.part;
    ShellSG& d = S.d; memset(&d, 0, sizeof d);
    d.ne = ne; d.ne_pad = np; d.order0 = order; d.blk0 = blk; d.law = G.law; d.npt = G.prop.npt;
    d.nvartmp = (G.law == 36) ? 2 + G.m36.nrate : 0;
    d.nhourg = (nnode == 3) ? 0 : shell_is_qeph(G.prop) ? 12 : 5;
    d.m2 = G.m2; d.m36 = G.m36; d.prop = G.prop; d.dtfac = (nnode == 3) ? ctl.dtfac_sh3n : ctl.dtfac_shell; d.nodadt = ctl.nodadt;
    if (G.law == 36) {
      if (npf.empty()) { orgpu_set_error("LAW36 group without a function table (orgpu_set_functions)"); return -4; }
      for (int j = 0; j < G.m36.nrate; j++) {
        const int f = G.m36.ifunc[j];
        if (f < 0 || f + 1 >= (int)npf.size() || npf[f + 1] - npf[f] < 2) { orgpu_set_error("LAW36 curve %d missing or shorter than 2 points", f); return -4; }
      }
      if (!d_tf) {
        double* t; int* n; std::vector<void*>& own = S.owned;
        if (sh_upload(own, &t, tf) || sh_upload(own, &n, npf)) return -100;
        d_tf = t; d_npf = n;
      }
      d.tf = d_tf; d.npf = d_npf;
      curve_tab_fill(d.ct, G.m36, npf, tf);
    }
    // ---- slab word map (shell_common.cuh)
    const bool has_temp = (G.law == 2) && G.m2.has_temp;
    d.w_ip0 = SW_HOURG + d.nhourg; d.nwip = has_temp ? 8 : 7;
    d.iw_sigb = -1;
    if ((G.law == 36 && G.m36.fisokin > 0.0) || (G.law == 2 && G.m2.fisokin > 0.0)) { d.iw_sigb = d.nwip; d.nwip += 3; }      // LBUF%SIGB: back stress of the kinematic hardening
    d.iw_plap = -1;
    if (G.law == 36 && G.m36.vp == 1) d.iw_plap = d.nwip++;     // UVAR(2): filtered plastic strain rate of the point
    d.fail = G.fail; d.iw_dfmax = d.iw_foff = -1; d.fail_pthkf = 0.0;
    if (G.fail.irupt == 1) {                                    // /FAIL/JOHNSON: damage and point flag of every integration point
      d.iw_dfmax = d.nwip++; d.iw_foff = d.nwip++;
      double p = G.fail.pthk; const double pg = G.fail.pthickg;  // fail_setoff_c.F:139-151
      if (p > 0.0)      { p = std::min(p, std::fabs(pg)); p = std::max(std::min(p, 1.0 - 1e-6), 1e-6); }
      else if (p < 0.0) { p = std::max(p, -std::fabs(pg)); p = std::min(std::max(p, -1.0 + 1e-6), -1e-6); }
      else p = pg;
      d.fail_pthkf = p;
    }
    d.w_vt = d.w_ip0 + d.npt * d.nwip;
    // The VARTMP cursors of a rate-dependent LAW36 (2 + NRATE ints per point) would push the NPT = 5 tile past what stages
    // three-per-SM; they are search hints (the segment found does not depend on where the search starts), so the groups the
    // FAST = 2 kernel takes keep one BYTE per rate curve and point when every curve has at most 255 points.
    bool vt_bytes = !getenv("ORGPU_NO_FAST") && G.law == 36 && G.m36.nrate > 1 && G.prop.ipla == 1 && G.m36.ifail == 0 && G.m36.fisokin == 0.0 && G.m36.vp == 0 &&
                    G.fail.irupt == 0 && G.prop.npt <= 5;
    if (vt_bytes) for (int j = 0; j < G.m36.nrate; j++) if (npf[G.m36.ifunc[j] + 1] - npf[G.m36.ifunc[j]] > 255) vt_bytes = false;
    for (;;) {
      d.vt_bytes = vt_bytes;
      d.nvt = (G.law == 36) ? (G.m36.nrate == 1 ? 1 : vt_bytes ? G.m36.nrate : d.nvartmp) : 0;
      d.nw_rw = d.w_vt + (vt_bytes ? (d.npt * d.nvt + 7) / 8 : (d.npt * d.nvt + 1) / 2);
      int w = d.nw_rw;
      d.w_thke = (G.prop.ithk > 0) ? -1 : w; if (d.w_thke >= 0) w++;
      d.w_slot = w; w += 2;
      d.nw = w;
      if (!vt_bytes || (size_t)d.nw * ORGPU_TILE * 8 <= ORGPU_STAGE_MAX_BYTES) break;
      vt_bytes = false;                                          // does not stage either way: the generic kernel and its int rows
    }
    HostSlab H; H.init(d.nw, np);
    const int ixs_ = (nnode == 3) ? 6 : 7, iuid = (nnode == 3) ? 5 : 6;      // row length of IXTG / IXC, column of the user id
    std::vector<int> conn((size_t)nnode * np, 0), ngl(np, 0), conn_t;
    for (int i = 0; i < np; i++) {
      H.at(SW_THK, i) = G.prop.thick; if (d.w_thke >= 0) H.at(d.w_thke, i) = G.prop.thick;
      if (has_temp) for (int ip = 0; ip < d.npt; ip++) H.at(d.w_ip0 + ip * d.nwip + IW_TEMP, i) = G.m2.tini;
      if (d.iw_foff >= 0) for (int ip = 0; ip < d.npt; ip++) H.at(d.w_ip0 + ip * d.nwip + d.iw_foff, i) = 1.0;      // FOFF = 1: point alive
    }
    for (int i = 0; i < ne; i++) {
      const int* ix = &ixc[(size_t)ixs_ * (nft + i)];
      for (int k = 0; k < nnode; k++) {
        const int node = ix[1 + k];
        if (node < 1 || node > numnod) { orgpu_set_error("IXC / IXTG node %d out of range (element %d)", node, nft + i + 1); return -4; }
        const int sl = iadc[(size_t)nnode * (nft + i) + k];
        if (sl < 1 || sl > lsky) { orgpu_set_error("IADC / IADTG slot %d out of range (element %d)", sl, nft + i + 1); return -4; }
        conn[(size_t)k * np + i] = node - 1; H.iat(d.w_slot, k, i) = sl - 1;
      }
      ngl[i] = ix[iuid]; H.at(SW_OFF, i) = 1.0;
    }
    tile_major_ints(conn_t, conn, nnode, np);
    int *dconn, *dngl;
    if (sh_upload(S.owned, &dconn, conn_t) || sh_upload(S.owned, &dngl, ngl) || sh_upload(S.owned, &d.slab, H.h)) return -100;
    d.conn = dconn; d.ngl = dngl;
    if (sh_alloc(S.owned, &d.smstr, (size_t)(nnode == 3 ? 3 : 6) * np)) return -100;
    const int nblk = np / ORGPU_TILE;                    // dt candidate slots: one per CTA
    if ((int)sgr.size() >= ORGPU_MAX_SG) { orgpu_set_error("too many super-groups (%d)", ORGPU_MAX_SG); return -6; }
    sgr.push_back(SGRange{blk, nblk, (nnode == 3) ? ORGPU_FAM_SH3N : shell_is_qeph(G.prop) ? ORGPU_FAM_SHELL_QEPH : ORGPU_FAM_SHELL_BT, d.order0, d.ngl});
    order += ne; blk += nblk; gi = gj;
  }
  return 0;
}

template <class K>
static void shell_launch_one(K kern_staged, K kern_direct, const ShellParams& P, int nblk, cudaStream_t st)
{
  const size_t bytes = (size_t)P.sg.nw * ORGPU_TILE * 8;
#ifndef ORGPU_NO_STAGING
  if (bytes <= ORGPU_STAGE_MAX_BYTES) {
    stage_attr((const void*)kern_staged, bytes, ORGPU_SHELL_MINB);
    kern_staged<<<nblk, ORGPU_SHELL_CTA, bytes, st>>>(P);
    return;
  }
#endif
  kern_direct<<<nblk, ORGPU_SHELL_CTA, 0, st>>>(P);
}

// the compile-time specialisation of the LAW36 kernels: Iplas = 1, no failure inside the law, one static curve small enough for
// the kernel parameters (ORGPU_NO_FAST=1: generic path)
static inline bool shell_fast(const ShellSG& d) {
  static const bool off = getenv("ORGPU_NO_FAST") != nullptr;
  return !off && d.law == 36 && d.prop.ipla == 1 && d.m36.ifail == 0 && d.m36.nrate == 1 && d.ct.n > 0 && d.m36.fisokin == 0.0 && d.fail.irupt == 0
         && d.prop.npt <= 5 && (size_t)d.nw * ORGPU_TILE * 8 <= ORGPU_STAGE_MAX_BYTES;     // the three-pass loop: staged tile, NPT <= 5 (shell_common.cuh)
}

// ... and its wider sibling (FAST = 2): LAW36, Iplas = 1, no failure, isotropic hardening with ANY number of rate curves (in the
// kernel parameters or in global memory) -- rate-dependent /MAT/PLAS_TAB, the usual crash material -- through the three-pass loop
static inline bool shell_fast2(const ShellSG& d) {
  static const bool off = getenv("ORGPU_NO_FAST") != nullptr;
  return !off && !shell_fast(d) && d.law == 36 && d.prop.ipla == 1 && d.m36.ifail == 0 && d.m36.fisokin == 0.0 && d.m36.vp == 0 && d.fail.irupt == 0
         && d.prop.npt <= 5 && (d.m36.nrate == 1 || d.vt_bytes) && (size_t)d.nw * ORGPU_TILE * 8 <= ORGPU_STAGE_MAX_BYTES;
}

static void launch_shell_forces(ShellSGHost& S, const DevNodes& nd, double* fsky, int roww, CycleState* cs,
                                const DtBlocks& db, const FinalizeArgs& fa, cudaStream_t st)
{
  (void)roww; (void)fa;                        // shell models always use 8-wide rows
  ShellParams P{S.d, nd, fsky, cs, db, nullptr, nullptr};
  const int nblk = S.d.ne_pad / ORGPU_TILE;
  if (S.sh3n) {
    if (S.d.law == 36 && S.d.m36.ifail == 2) shell_launch_one(c3_forces_kernel<37, true>, c3_forces_kernel<37, false>, P, nblk, st);
    else if (S.d.law == 36 && shell_fast(S.d)) shell_launch_one(c3_forces_kernel<36, true, 1>, c3_forces_kernel<36, false>, P, nblk, st);
    else if (S.d.law == 36 && shell_fast2(S.d)) shell_launch_one(c3_forces_kernel<36, true, 2>, c3_forces_kernel<36, false>, P, nblk, st);
    else if (S.d.law == 36) shell_launch_one(c3_forces_kernel<36, true>, c3_forces_kernel<36, false>, P, nblk, st);
    else               shell_launch_one(c3_forces_kernel<2, true>, c3_forces_kernel<2, false>, P, nblk, st);
  } else if (shell_is_qeph(S.d.prop)) {
    if (S.d.law == 36 && S.d.m36.ifail == 2) shell_launch_one(qeph_forces_kernel<37, true>, qeph_forces_kernel<37, false>, P, nblk, st);
    else if (S.d.law == 36 && shell_fast(S.d)) shell_launch_one(qeph_forces_kernel<36, true, 1>, qeph_forces_kernel<36, false>, P, nblk, st);
    else if (S.d.law == 36 && shell_fast2(S.d)) shell_launch_one(qeph_forces_kernel<36, true, 2>, qeph_forces_kernel<36, false>, P, nblk, st);
    else if (S.d.law == 36) shell_launch_one(qeph_forces_kernel<36, true>, qeph_forces_kernel<36, false>, P, nblk, st);
    else               shell_launch_one(qeph_forces_kernel<2, true>, qeph_forces_kernel<2, false>, P, nblk, st);
  } else {
    if (S.d.law == 36 && S.d.m36.ifail == 2) shell_launch_one(bt_forces_kernel<37, true>, bt_forces_kernel<37, false>, P, nblk, st);
    else if (S.d.law == 36 && shell_fast(S.d)) shell_launch_one(bt_forces_kernel<36, true, 1>, bt_forces_kernel<36, false>, P, nblk, st);
    else if (S.d.law == 36 && shell_fast2(S.d)) shell_launch_one(bt_forces_kernel<36, true, 2>, bt_forces_kernel<36, false>, P, nblk, st);
    else if (S.d.law == 36) shell_launch_one(bt_forces_kernel<36, true>, bt_forces_kernel<36, false>, P, nblk, st);
    else               shell_launch_one(bt_forces_kernel<2, true>, bt_forces_kernel<2, false>, P, nblk, st);
  }
}

// fields: 0 for(5) 1 mom(3) 2 eint(2) 3 thk 4 off 5 stra(8) 6 epsd 7 hourg(nhourg) 8 smstr(6)
//         9 sig(5*npt) 10 pla(npt) 11 epsd_ip(npt) 12 temp(npt) 13 sigb(3*npt) ; out[k*numelc + e]
static int shell_state_xfer(std::vector<ShellSGHost>& sgs, int numelc, int field, double* out, bool up, bool sh3n = false)
{
  const size_t NE = numelc;
  for (auto& S : sgs) {
    if (S.sh3n != sh3n) continue;
    const ShellSG& d = S.d; double* base = d.slab; int nw = d.nw, nc = 1, w0 = 0, ipw = -1;
    switch (field) {
      case 0: w0 = SW_FOR; nc = 5; break; case 1: w0 = SW_MOM; nc = 3; break; case 2: w0 = SW_EINT; nc = 2; break;
      case 3: w0 = SW_THK; break; case 4: w0 = SW_OFF; break; case 5: w0 = SW_STRA; nc = 8; break; case 6: w0 = SW_EPSD; break;
      case 7: w0 = SW_HOURG; nc = d.nhourg; if (nc == 0) continue; break; case 8: base = d.smstr; nw = S.sh3n ? 3 : 6; w0 = 0; nc = nw; break;
      case 9: ipw = IW_SIG; nc = 5 * d.npt; break; case 10: ipw = IW_PLA; nc = d.npt; break; case 11: ipw = IW_EPSD; nc = d.npt; break;
      case 12: if (d.law != 2 || !d.m2.has_temp) continue; ipw = IW_TEMP; nc = d.npt; break;
      case 13: if (d.iw_sigb < 0) continue; nc = 3 * d.npt; break;
      case 14: if (d.iw_dfmax < 0) continue; ipw = d.iw_dfmax; nc = d.npt; break; case 15: if (d.iw_foff < 0) continue; ipw = d.iw_foff; nc = d.npt; break;
      case 16: if (d.iw_plap < 0) continue; ipw = d.iw_plap; nc = d.npt; break;
      default: orgpu_set_error("unknown shell field %d", field); return -1;
    }
    for (int k = 0; k < nc; k++) {
      int w = w0 + k;
      if (ipw >= 0) w = (field == 9) ? d.w_ip0 + (k / 5) * d.nwip + IW_SIG + (k % 5) : d.w_ip0 + k * d.nwip + ipw;
      if (field == 13) w = d.w_ip0 + (k / 3) * d.nwip + d.iw_sigb + (k % 3);
      if ((up ? slab_upload_word(base, nw, w, d.ne, out + k * NE + S.first_elem)
              : slab_download_word(base, nw, w, d.ne, out + k * NE + S.first_elem)) != cudaSuccess) {
        orgpu_set_error("shell state transfer failed"); return -100; }
    }
  }
  return 0;
}

// ---- batched launches: all super-groups of one kernel variant in ONE launch (table-driven CTA -> (super-group, tile)) --------
// variant of a shell super-group whose table-driven kernel is compiled (-1: launched on its own)
enum { SHV_QEPH36F = 0, SHV_QEPH36R, SHV_QEPH36, SHV_QEPH2, SHV_BT36F, SHV_BT36R, SHV_BT36, SHV_BT2, SHV_COUNT };   // F / R: three-pass loop, one static curve / rate curves
static inline int shell_tab_variant(const ShellSGHost& S)
{
  const ShellSG& d = S.d;
  if (S.sh3n || (size_t)d.nw * ORGPU_TILE * 8 > ORGPU_STAGE_MAX_BYTES) return -1;
  if (d.law == 36 && d.m36.ifail == 2) return -1;
  if (shell_is_qeph(d.prop)) return d.law == 36 ? (shell_fast(d) ? SHV_QEPH36F : shell_fast2(d) ? SHV_QEPH36R : SHV_QEPH36) : SHV_QEPH2;
  return d.law == 36 ? (shell_fast(d) ? SHV_BT36F : shell_fast2(d) ? SHV_BT36R : SHV_BT36) : SHV_BT2;
}
template <class K>
static void shell_launch_tab_k(K kern, const ShellParams& P, int nblk, size_t bytes, cudaStream_t st)
{ stage_attr((const void*)kern, bytes, ORGPU_SHELL_MINB); kern<<<nblk, ORGPU_SHELL_CTA, bytes, st>>>(P); }
static void launch_shell_forces_tab(int variant, const ShellSG* d_tab, const int2* d_map, int nblk, size_t bytes, const DevNodes& nd,
                                    double* fsky, CycleState* cs, const DtBlocks& db, cudaStream_t st)
{
  ShellParams P; memset(&P.sg, 0, sizeof P.sg); P.nd = nd; P.fsky = fsky; P.cs = cs; P.db = db; P.sgtab = d_tab; P.cta_map = d_map;
  switch (variant) {
    case SHV_QEPH36F: shell_launch_tab_k(qeph_forces_kernel<36, true, 1, true>, P, nblk, bytes, st); break;
    case SHV_QEPH36R: shell_launch_tab_k(qeph_forces_kernel<36, true, 2, true>, P, nblk, bytes, st); break;
    case SHV_QEPH36:  shell_launch_tab_k(qeph_forces_kernel<36, true, 0, true>, P, nblk, bytes, st); break;
    case SHV_QEPH2:   shell_launch_tab_k(qeph_forces_kernel<2, true, 0, true>, P, nblk, bytes, st); break;
    case SHV_BT36F:   shell_launch_tab_k(bt_forces_kernel<36, true, 1, true>, P, nblk, bytes, st); break;
    case SHV_BT36R:   shell_launch_tab_k(bt_forces_kernel<36, true, 2, true>, P, nblk, bytes, st); break;
    case SHV_BT36:    shell_launch_tab_k(bt_forces_kernel<36, true, 0, true>, P, nblk, bytes, st); break;
    default:          shell_launch_tab_k(bt_forces_kernel<2, true, 0, true>, P, nblk, bytes, st); break;
  }
}
