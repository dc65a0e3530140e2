// shell_common.cuh -- 4-node shell super-groups: device layout and the through-thickness material
// loop shared by the QEPH (CZFORC3) and Belytschko-Tsay (CFORC3) kernels.
//
// Device layout of ELBUF for one super-group (consecutive groups with identical family, law,
// property): one tile-major slab [tile][word][128] (common.cuh).  Words of one element, in order:
//   G_BUFEL_ (elbufdef_mod.F90:739-1013): FOR 0-4, MOM 5-7, EINT 8-9, THK 10, OFF 11, STRA 12-19,
//   EPSD 20, HOURG 21..21+nhourg-1;
//   then per integration point (L_BUFEL_, :1184-1300; IP-major like shell_internal_forces.F90:829-886)
//   SIG 0-4, PLA 5, EPSD 6 (, TEMP 7 for LAW2 with a temperature buffer);
//   then the LAW36 VARTMP table cursors as int half-rows; then read-only words: the initial thickness
//   (only when ITHK=0 keeps using it) and the four FSKY slot indices (int half-rows).
//
// The material loop restates, one element per thread, all in registers:
//   CMAIN3 -> LAYINI (layini.F:246-254) -> MULAWC (mulawc.F90:542-604, 718-1114, 2630-2662,
//   2818-2845, 2934-3091) with SIGEPS36C (sigeps36c.F:171-661, VP=0) + VINTER (vinter.F:100-130)
//   or SIGEPS02C (sigeps02c.F:91-230) + M2CPLR (m2cplr.F:108-507).
#pragma once
#include "common.cuh"
#include "../../include/or_quadrature.h"

struct ShellSG {
  int ne, ne_pad, order0, blk0;
  int law, npt, nvartmp, nhourg, vt_bytes;
  const int* conn;            // tile-major [tile][4][128], 0-based node
  const int* ngl;             // user ids [ne_pad]
  double* slab;               // [tile][nw][128]
  int nw, nw_rw;              // words per tile / written back
  int w_ip0, nwip;            // first word of integration point 0, words per point
  int iw_sigb;                // LBUF%SIGB (3 words: kinematic hardening, FISOKIN > 0) inside a point's words, -1: none
  int iw_plap;                // UVAR(2) of LAW36 with VP = 1 (filtered plastic strain rate of the point), -1: none
  int w_vt, nvt;              // first word of the VARTMP int rows, int rows per point (1 when NRATE=1: only cursor 3 is live);
                              // FAST = 2 with NRATE > 1: BYTE rows, NRATE per point (the cursors are search hints, not state)
  int w_thke, w_slot;         // initial-thickness word (-1 when ITHK>0), first word of the 4 slot int rows
  double* smstr;              // tile-major [tile][6][128]
  const double* tf; const int* npf;    // LAW36 function table (pairs), 0-based curve starts
  CurveTab ct;                // ... and its parameter-space copy when small (ct.n > 0)
  orgpu_law2 m2; orgpu_law36 m36; orgpu_prop_shell prop;
  double dtfac;               // DTFAC1(3)
  int nodadt;                 // /DT/NODA: nodal stiffnesses of cndt3.F:194-221, no element time step
  double* bal; int bal_ld;    // print cycles: the elements' PARTSAV(1:6) terms, bal[k * bal_ld + e] (null until orgpu_set_print)
  const double* gvol;         // GBUF%VOL of the group's elements (initial area x thickness): the mass of CBILAN
  const unsigned char* xs_ftile;    // several domains: 1 for the tiles that hold an element with a corner row to send (null: one domain)
  orgpu_fail fail;            // /FAIL/JOHNSON (irupt = 1) or none
  int iw_dfmax, iw_foff;      // damage / point flag inside a point's words (-1: no failure model)
  double fail_pthkf;          // P_thickfail as fail_setoff_c.F:139-151 resolves it (card value against the property's)
};

enum { SW_FOR = 0, SW_MOM = 5, SW_EINT = 8, SW_THK = 10, SW_OFF = 11, SW_STRA = 12, SW_EPSD = 20, SW_HOURG = 21 };
enum { IW_SIG = 0, IW_PLA = 5, IW_EPSD = 6, IW_TEMP = 7 };

struct ShellParams {
  ShellSG sg; DevNodes nd; double* fsky; CycleState* cs; DtBlocks db;
  const ShellSG* sgtab; const int2* cta_map;     // batched launch over several super-groups (common.cuh cta_work); null otherwise
};

__constant__ double c_Z0[121];
__constant__ double c_WF[121];
__constant__ double c_WM[121];

// strains handed to the material loop and what comes back (the MULAWC argument list, reduced)
struct MatIO {
  double exx, eyy, exy, exz, eyz, kxx, kyy, kxy;   // in : increments
  double area, thk0, gs, rho, epsd_pg;             // in
  double off;                                      // in/out
  double ssp, viscmx, sigy, zcfac1, zcfac2, vol0;  // out (ssp in: PM(27))
  double fo[5], mo[3];                             // out: new GBUF%FOR / GBUF%MOM
};

struct IpState { double sxx, syy, sxy, syz, szx, pla, epsd, temp; int ipos; double sbx, sby, sbxy, plap; };   // sb*: back stress (FISOKIN > 0); plap: UVAR(2) (LAW36 VP = 1)

// state of one integration point out of / into the CTA's tile
template <int LAW, bool STAGED, int FAST = 0>
__device__ __forceinline__ IpState ip_load(const ShellSG& g, const TileAcc<STAGED>& T, int ipt)
{
  const int w = g.w_ip0 + ipt * g.nwip;
  IpState s;
  s.sxx = T.ld(w + IW_SIG); s.syy = T.ld(w + IW_SIG + 1); s.sxy = T.ld(w + IW_SIG + 2); s.syz = T.ld(w + IW_SIG + 3); s.szx = T.ld(w + IW_SIG + 4);
  s.pla = T.ld(w + IW_PLA);
  s.epsd = T.ld(w + IW_EPSD);
  s.temp = K_ZERO; s.ipos = 0; s.sbx = K_ZERO; s.sby = K_ZERO; s.sbxy = K_ZERO; s.plap = K_ZERO;
  if (FAST == 0 && LAW != 2 && g.iw_plap >= 0) s.plap = T.ld(w + g.iw_plap);
  if (FAST != 1 && g.iw_sigb >= 0) { s.sbx = T.ld(w + g.iw_sigb); s.sby = T.ld(w + g.iw_sigb + 1); s.sbxy = T.ld(w + g.iw_sigb + 2); }
  if (LAW == 2) { if (g.m2.has_temp) s.temp = T.ld(w + IW_TEMP); }
  else if (g.m36.nrate == 1) s.ipos = T.ldi(g.w_vt, ipt);
  return s;
}
template <int LAW, bool STAGED, int FAST = 0>
__device__ __forceinline__ void ip_store(const ShellSG& g, const TileAcc<STAGED>& T, int ipt, const IpState& s, int ipos_old, double temp_old)
{
  const int w = g.w_ip0 + ipt * g.nwip;
  T.st(w + IW_SIG, s.sxx); T.st(w + IW_SIG + 1, s.syy); T.st(w + IW_SIG + 2, s.sxy); T.st(w + IW_SIG + 3, s.syz); T.st(w + IW_SIG + 4, s.szx);
  T.st(w + IW_PLA, s.pla);
  T.st(w + IW_EPSD, s.epsd);
  if (FAST != 1 && g.iw_sigb >= 0) { T.st(w + g.iw_sigb, s.sbx); T.st(w + g.iw_sigb + 1, s.sby); T.st(w + g.iw_sigb + 2, s.sbxy); }
  if (FAST == 0 && LAW != 2 && g.iw_plap >= 0) T.st(w + g.iw_plap, s.plap);
  if (LAW == 2) { if (g.m2.has_temp && s.temp != temp_old) T.st(w + IW_TEMP, s.temp); }
  else if (g.m36.nrate == 1 && s.ipos != ipos_old) T.sti(g.w_vt, ipt, s.ipos);
}

#define ORGPU_SHELL_CTA ORGPU_TILE   // one CTA = one state tile

// CBILAN (cbilan.F:183-275, NN = 4) / C3BILAN (c3bilan.F:150-167, NN = 3) on a print cycle: the element's terms of
// PARTSAV(1:6, part) -- internal energy EINT(1)+EINT(2), kinetic energy from the nodal velocities the force routine
// sees, momenta, mass -- go to the scratch rows; the nodal velocities are gathered again (print cycles only)
template <int NN, bool STAGED>
__device__ __forceinline__ void shell_bilan(const ShellParams& P, const TileAcc<STAGED>& T, int tile, int e, double rho, double off)
{
  const ShellSG& g = P.sg;
  const int* cn = g.conn + (size_t)tile * NN * ORGPU_TILE + threadIdx.x;
  double vx[NN], vy[NN], vz[NN];
  #pragma unroll
  for (int k = 0; k < NN; k++) { const double4 v = ld256_nc(P.nd.vel + __ldg(cn + k * ORGPU_TILE)); vx[k] = v.x; vy[k] = v.y; vz[k] = v.z; }
  double vxa = K_ZERO, vya = K_ZERO, vza = K_ZERO, va2 = K_ZERO;
  #pragma unroll
  for (int k = 0; k < NN; k++) { vxa = vxa + vx[k]; vya = vya + vy[k]; vza = vza + vz[k]; }
  #pragma unroll
  for (int k = 0; k < NN; k++) va2 = va2 + vx[k] * vx[k];
  #pragma unroll
  for (int k = 0; k < NN; k++) va2 = va2 + vy[k] * vy[k];
  #pragma unroll
  for (int k = 0; k < NN; k++) va2 = va2 + vz[k] * vz[k];
  const double xmas = rho * __ldg(g.gvol + e);
  const double ei = T.ld(SW_EINT) + T.ld(SW_EINT + 1);
  const double ek = (NN == 4) ? xmas * va2 * K_ONE_OVER_8 : xmas * va2 * K_ONE_OVER_6;
  const double xmas25 = (NN == 4) ? xmas * K_FOURTH : xmas * K_THIRD;
  double* b = g.bal + e; const size_t ld = g.bal_ld;
  b[0] = ei; b[ld] = ek; b[2 * ld] = xmas25 * vxa; b[3 * ld] = xmas25 * vya; b[4 * ld] = xmas25 * vza;
  b[5 * ld] = (off != K_ZERO) ? xmas : K_ZERO;
}

// ---- SIGEPS36C, VP = 0 -------------------------------------------------------------------
// FAIL2: IFAIL = 2 (tensile-strain damage / failure), compiled as its own kernel variant (template LAW = 37) so that
// the common LAW36 kernels carry none of it.
// The law is written in three stages so that two integration points can be advanced side by side (law36_pair_ip1):
//   law36_trial : elastic predictor :277-281, strain rate :288-296, yield stress and hardening modulus from the curves :317-461
//   the return  : Iplas 0 :472-501, 1 :503-593 (three Newton steps), 2 :595-661
//   the tail    : failure :928-950
template <bool STAGED, bool FAIL2, int FAST = 0>
__device__ __forceinline__ void law36_trial(const ShellSG& g, const TileAcc<STAGED>& T, int ipt, double asrate,
                                            double dexx, double deyy, double dexy, double deyz, double dezx, double dtinv,
                                            double gs, double epsd_pg, double zt, IpState& s, double& YLD, double& H, double& EPST)
{
  const orgpu_law36& m = g.m36;
  // IFAIL = 2: damage factor on the largest in-plane principal total strain at the point (mulawc.F90:856-862,
  // sigeps36c.F:256-264); GBUF%STRA was accumulated by the strain routine earlier in this cycle
  double FAIL = K_ONE; EPST = K_ZERO;
  if (FAIL2) {
    const double epsxx = T.ld(SW_STRA) + zt * T.ld(SW_STRA + 5);
    const double epsyy = T.ld(SW_STRA + 1) + zt * T.ld(SW_STRA + 6);
    const double epsxy = T.ld(SW_STRA + 2) + zt * T.ld(SW_STRA + 7);
    EPST = K_HALF * (epsxx + epsyy + or_sqrt((epsxx - epsyy) * (epsxx - epsyy) + epsxy * epsxy));
    FAIL = fmax(K_EM20, fmin(K_ONE, or_div(m.epsr2 - EPST, m.epsr2 - m.epsr1)));
  }
  const double A1 = m.a1u, A2 = m.a2u, G = m.shear;
  const double pla = s.pla;
  // elastic predictor, from the stress shifted by the back stress when the hardening has a kinematic part (sigeps36c.F:272-274)
  const double FISOKIN = (FAST == 1) ? K_ZERO : m.fisokin;
  double sox = s.sxx, soy = s.syy, soxy = s.sxy;
  if (FISOKIN > K_ZERO) { sox = sox - s.sbx; soy = soy - s.sby; soxy = soxy - s.sbxy; }
  s.sxx = sox + A1 * dexx + A2 * deyy;
  s.syy = soy + A2 * dexx + A1 * deyy;
  s.sxy = soxy + G * dexy;
  s.syz = s.syz + gs * deyz;
  s.szx = s.szx + gs * dezx;
  // strain rate
  // VP = 1 (sigeps36c.F:665-923): the curves are interpolated on the point's filtered PLASTIC strain rate UVAR(2); LBUF%EPSD is
  // left alone (the Starter keeps VP = 0 for a single curve, hm_read_mat36.F:199)
  const bool vp1 = (FAST == 0) && m.vp == 1;
  double epsd;
  if (vp1) epsd = s.plap;
  else {
    if (m.israte == 0) {
      const double exx = dexx * dtinv, eyy = deyy * dtinv, exy = dexy * dtinv;
      epsd = K_HALF * (fabs(exx + eyy) + or_sqrt((exx - eyy) * (exx - eyy) + exy * exy));
    } else {
      epsd = asrate * epsd_pg + (K_ONE - asrate) * s.epsd;
    }
    s.epsd = epsd;
  }
  // yield stress and hardening modulus from the tabulated curves
  if (FAST == 1 || m.nrate == 1) {
    int ipos = s.ipos;
    const int f = m.ifunc[0];
    double dydx, y1;
    if (FAST == 1 || g.ct.n > 0) vinter1c(g.ct, 0, ipos, pla, dydx, y1);
    else { const int i0 = __ldg(g.npf + f), i1 = __ldg(g.npf + f + 1); vinter1(g.tf, i0, i1 - i0, ipos, pla, dydx, y1); }
    s.ipos = ipos;
    const double FACT = FAIL * K_ONE * (m.yfac[0] * K_ONE);
    H = dydx * FACT;
    if (FISOKIN == K_ZERO) YLD = y1 * FACT;                            // :329-337
    else {
      const double YLD0 = (g.ct.n > 0) ? g.ct.tf[2 * g.ct.i0[0] + 1] : __ldg(g.tf + 2 * (size_t)__ldg(g.npf + f) + 1);   // first point of the curve
      YLD = (FISOKIN == K_ONE) ? YLD0 * FACT : ((K_ONE - FISOKIN) * y1 + FISOKIN * YLD0) * FACT;
    }
  } else {
    int JJ = 1;
    for (int J = 2; J <= m.nrate - 1; J++) if (epsd >= m.rate[J - 1]) JJ = J;
    double RFAC;
    if (m.ismooth == 2) {
      const double EPSP1 = fmax(m.rate[JJ - 1], K_EM20), EPSP2 = m.rate[JJ];
      RFAC = or_div(log(or_div(fmax(epsd, K_EM20), EPSP1)), log(or_div(EPSP2, EPSP1)));
    } else {
      const double EPSP1 = m.rate[JJ - 1], EPSP2 = m.rate[JJ];
      RFAC = or_div((epsd - EPSP1), (EPSP2 - EPSP1));
    }
    const double YFAC1 = m.yfac[JJ - 1] * K_ONE, YFAC2 = m.yfac[JJ] * K_ONE;
    const int f1 = m.ifunc[JJ - 1], f2 = m.ifunc[JJ];
    int ipos1, ipos2;
    if constexpr (FAST == 2) { ipos1 = T.ldb_lane(g.w_vt, ipt * g.nvt + JJ - 1, threadIdx.x); ipos2 = T.ldb_lane(g.w_vt, ipt * g.nvt + JJ, threadIdx.x); }   // byte rows, one per rate curve
    else { ipos1 = T.ldi(g.w_vt, ipt * g.nvt + 1 + JJ); ipos2 = T.ldi(g.w_vt, ipt * g.nvt + 2 + JJ); }
    double dydx1, y1, dydx2, y2;
    if (g.ct.n > 0) { vinter1c(g.ct, JJ - 1, ipos1, pla, dydx1, y1); vinter1c(g.ct, JJ, ipos2, pla, dydx2, y2); }
    else {
      { const int i0 = __ldg(g.npf + f1), i1 = __ldg(g.npf + f1 + 1); vinter1(g.tf, i0, i1 - i0, ipos1, pla, dydx1, y1); }
      { const int i0 = __ldg(g.npf + f2), i1 = __ldg(g.npf + f2 + 1); vinter1(g.tf, i0, i1 - i0, ipos2, pla, dydx2, y2); }
    }
    if (FISOKIN == K_ZERO) {                                            // :383-404
      y1 = y1 * YFAC1; y2 = y2 * YFAC2;
      YLD = FAIL * (y1 + RFAC * (y2 - y1));
      YLD = fmax(YLD, K_EM20);
      dydx1 = dydx1 * YFAC1; dydx2 = dydx2 * YFAC2;
      H = FAIL * (dydx1 + RFAC * (dydx2 - dydx1));
      YLD = YLD * fmax(K_ZERO, K_ONE);
      H = H * fmax(K_ZERO, K_ONE);
    } else {
      // first points of the two curves: the yield stress of the kinematic part (:405-460)
      double y01, y02;
      if (g.ct.n > 0) { y01 = g.ct.tf[2 * g.ct.i0[JJ - 1] + 1]; y02 = g.ct.tf[2 * g.ct.i0[JJ] + 1]; }
      else { y01 = __ldg(g.tf + 2 * (size_t)__ldg(g.npf + f1) + 1); y02 = __ldg(g.tf + 2 * (size_t)__ldg(g.npf + f2) + 1); }
      if (FISOKIN == K_ONE) {
        dydx1 = dydx1 * YFAC1; dydx2 = dydx2 * YFAC2;
        H = FAIL * (dydx1 + RFAC * (dydx2 - dydx1));
        y1 = y01 * YFAC1; y2 = y02 * YFAC2;
        YLD = FAIL * (y1 + RFAC * (y2 - y1));
        YLD = YLD * fmax(K_ZERO, K_ONE);
        H = H * fmax(K_ZERO, K_ONE);
      } else {
        y1 = y1 * YFAC1; y2 = y2 * YFAC2;
        YLD = FAIL * (y1 + RFAC * (y2 - y1));
        YLD = fmax(YLD, K_EM20);
        dydx1 = dydx1 * YFAC1; dydx2 = dydx2 * YFAC2;
        H = FAIL * (dydx1 + RFAC * (dydx2 - dydx1));
        y1 = y01 * YFAC1; y2 = y02 * YFAC2;
        YLD = (K_ONE - FISOKIN) * YLD + FISOKIN * (FAIL * (y1 + RFAC * (y2 - y1)));
        YLD = YLD * fmax(K_ZERO, K_ONE);
        H = H * fmax(K_ZERO, K_ONE);
      }
    }
    if constexpr (FAST == 2) { T.stb(g.w_vt, ipt * g.nvt + JJ - 1, ipos1); T.stb(g.w_vt, ipt * g.nvt + JJ, ipos2); }
    else { T.sti(g.w_vt, ipt * g.nvt + 1 + JJ, ipos1); T.sti(g.w_vt, ipt * g.nvt + 2 + JJ, ipos2); }
  }
  if (m.yldcheck == 1) YLD = fmax(YLD, K_EM20);
}

// one Newton step of the Iplas = 1 return (sigeps36c.F:534-560): DPLA_J -> (DPLA_I, DR, PP, QQ, next DPLA_J)
struct L36Newton { double AA, BB, YLD, HI, NU11, NU21, DPLA_I, DPLA_J, DR, PP, QQ; };
__device__ __forceinline__ void law36_newton_step(double E, L36Newton& n, bool last)
{
  n.DPLA_I = n.DPLA_J;
  const double YLD_I = n.YLD + n.HI * n.DPLA_I;
  n.DR = or_div(K_HALF * E * n.DPLA_I, YLD_I);
  n.PP = or_div(K_ONE, (K_ONE + n.DR * n.NU11));
  n.QQ = or_div(K_ONE, (K_ONE + n.DR * n.NU21));
  if (last) return;                                   // the third step's residual is never used
  const double P2 = n.PP * n.PP, Q2 = n.QQ * n.QQ;
  const double F = n.AA * P2 + n.BB * Q2 - YLD_I * YLD_I;
  double DF = -or_div((n.AA * n.NU11 * P2 * n.PP + n.BB * n.NU21 * Q2 * n.QQ) * (E - K_TWO * n.DR * n.HI), YLD_I) - K_TWO * n.HI * YLD_I;
  DF = copysign(fmax(fabs(DF), K_EM20), DF);
  n.DPLA_J = (n.DPLA_I > K_ZERO) ? fmax(K_ZERO, n.DPLA_I - or_div(F, DF)) : K_ZERO;
}

template <bool STAGED, bool FAIL2, int FAST = 0>
__device__ __forceinline__ void law36_ip(const ShellSG& g, const TileAcc<STAGED>& T, int ipt, int ipla, double asrate,
                                         double dexx, double deyy, double dexy, double deyz, double dezx, double dtinv,
                                         double thklyl, double gs, double epsd_pg, double zt, double& off,
                                         IpState& s, double& thk, double& ssp, double& etse, double& yld_out, double dt1)
{
  const orgpu_law36& m = g.m36;
  const double E = m.young, G3 = m.g3;
  const bool vp1 = (FAST == 0) && m.vp == 1;
  if (vp1) ipla = 1;                                   // VP = 1: always the three Newton steps (sigeps36c.F:832-921)
  ssp = m.soundsp; etse = K_ONE;
  double pla = s.pla;
  double YLD, H, EPST;
  law36_trial<STAGED, FAIL2, FAST>(g, T, ipt, asrate, dexx, deyy, dexy, deyz, dezx, dtinv, gs, epsd_pg, zt, s, YLD, H, EPST);
  double DPLA_I = K_ZERO;                              // plastic strain increment of the point (drives the back stress)
  // projection on the yield surface
  if (ipla == 0) {
    const double NU3 = K_ONE - m.nu_mnu;
    const double SVM2 = s.sxx * s.sxx + s.syy * s.syy - s.sxx * s.syy + K_THREE * s.sxy * s.sxy;
    if (SVM2 > YLD * YLD) {
      const double SVM = or_sqrt(SVM2);
      const double R = or_div(YLD, SVM);
      s.sxx = s.sxx * R; s.syy = s.syy * R; s.sxy = s.sxy * R;
      const double DPLA = or_div(off * SVM * (K_ONE - R), (G3 + H));
      DPLA_I = DPLA;
      pla = pla + DPLA;
      double DEZZ = (YLD != 0) ? or_div(DPLA * K_HALF * (s.sxx + s.syy), YLD) : K_ZERO;
      DEZZ = -(dexx + deyy) * m.nu_mnu - NU3 * DEZZ;
      thk = thk + DEZZ * thklyl * off;
      etse = or_div(H, (H + E));
    }
  } else if (ipla == 1) {
    H = fmax(K_ZERO, H);
    double S1 = s.sxx + s.syy, S2 = s.sxx - s.syy; const double S3 = s.sxy;
    const double AA = K_FOURTH * S1 * S1;
    const double BB = K_THREE_OVER_4 * S2 * S2 + K_THREE * S3 * S3;
    const double SVM2 = AA + BB;
    { const double DEZZ = -(dexx + deyy) * m.nu_mnu; thk = thk + DEZZ * thklyl * off; }
    if (SVM2 > YLD * YLD && off == K_ONE) {
      const double SVM = or_sqrt(SVM2);
      etse = or_div(H, (H + E));
      const double HK = K_TWO_THIRD * H * m.fisokin;
      const double NU3 = K_ONE - m.nu_mnu;
      const double AAA = or_div(K_THREE * HK, E);
      L36Newton n;
      n.AA = AA; n.BB = BB; n.YLD = YLD; n.HI = H * (K_ONE - m.fisokin); n.NU11 = m.u_mnu + AAA; n.NU21 = m.t_pnu + AAA;
      n.DPLA_J = or_div((SVM - YLD), (G3 + H));
      n.DPLA_I = K_ZERO; n.DR = K_ZERO; n.PP = K_ONE; n.QQ = K_ONE;
      law36_newton_step(E, n, false); law36_newton_step(E, n, false); law36_newton_step(E, n, true);   // NITER = 3 (sigeps36c.F:167)
      pla = pla + n.DPLA_I; DPLA_I = n.DPLA_I;
      S1 = (s.sxx + s.syy) * n.PP;
      S2 = (s.sxx - s.syy) * n.QQ;
      s.sxx = K_HALF * (S1 + S2);
      s.syy = K_HALF * (S1 - S2);
      s.sxy = s.sxy * n.QQ;
      { const double DEZZ = -or_div(NU3 * n.DR * S1, E); thk = thk + DEZZ * thklyl * off; }
      YLD = YLD + n.HI * n.DPLA_I;
    }
  } else {
    H = fmax(K_ZERO, H);
    const double SVM2 = s.sxx * s.sxx + s.syy * s.syy - s.sxx * s.syy + K_THREE * s.sxy * s.sxy;
    { const double DEZZ = -(dexx + deyy) * m.nu_mnu; thk = thk + DEZZ * thklyl * off; }
    const double YLD2 = YLD * YLD;
    if (SVM2 > YLD2 && off == K_ONE) {
      const double NU3 = K_ONE - m.nu_mnu;
      const double A = or_div((SVM2 - YLD2), (K_FIVE * SVM2 + K_THREE * (-s.sxx * s.syy + s.sxy * s.sxy)));
      const double S1 = (K_ONE - K_TWO * A) * s.sxx + A * s.syy;
      const double S2 = A * s.sxx + (K_ONE - K_TWO * A) * s.syy;
      const double S3 = (K_ONE - K_THREE * A) * s.sxy;
      s.sxx = S1; s.syy = S2; s.sxy = S3;
      double SVM = or_sqrt(SVM2);
      const double DPLA = or_div(off * (SVM - YLD), (G3 + H));
      DPLA_I = DPLA;
      YLD = YLD + H * (K_ONE - m.fisokin) * DPLA;
      SVM = or_sqrt(s.sxx * s.sxx + s.syy * s.syy - s.sxx * s.syy + K_THREE * s.sxy * s.sxy);
      const double R = fmin(K_ONE, or_div(YLD, fmax(K_EM20, SVM)));
      s.sxx = s.sxx * R; s.syy = s.syy * R; s.sxy = s.sxy * R;
      pla = pla + DPLA;
      double DEZZ = or_div(DPLA * K_HALF * (s.sxx + s.syy), YLD);
      DEZZ = -NU3 * DEZZ;
      thk = thk + DEZZ * thklyl * off;
      etse = or_div(H, (H + E));
    }
  }
  // VP = 1: filter of the plastic strain rate (sigeps36c.F:976-982)
  if (vp1) { const double DTINV = or_div(K_ONE, fmax(dt1, K_EM20)); s.plap = asrate * DPLA_I * DTINV + (K_ONE - asrate) * s.plap; }
  // kinematic part of the hardening (sigeps36c.F:986-1002): the back stress grows along the new stress, which gets it back
  if (FAST != 1 && m.fisokin > K_ZERO) {
    const double ALPHA = or_div(m.fisokin * H * DPLA_I, YLD);
    s.sbx = s.sbx + ALPHA * s.sxx; s.sby = s.sby + ALPHA * s.syy; s.sbxy = s.sbxy + ALPHA * s.sxy;
    s.sxx = s.sxx + s.sbx; s.syy = s.syy + s.sby; s.sxy = s.sxy + s.sbxy;
  }
  // IFAIL = 1: failure on the maximum plastic strain (sigeps36c.F:928-938); MULAWC completes the deletion in the same cycle
  if (FAST == 1) {}
  else if (!FAIL2 && m.ifail == 1) { if (off == K_ONE && pla > m.epsmax) off = K_FOUR_OVER_5; }
  else if (FAIL2) { if (off == K_ONE && (pla > m.epsmax || EPST > m.epsf)) off = K_FOUR_OVER_5; }   // :940-950
  s.pla = pla;
  yld_out = YLD;
}

// ---- SIGEPS02C + M2CPLR (FISOKIN = 0) --------------------------------------------------------
__device__ __forceinline__ void law2_ip(const ShellSG& g, int ipla, int npttot, double dt1, double asrate,
                                        double dexx, double deyy, double dexy, double deyz, double dezx, double dtinv,
                                        double thklyl, double gs, double epsd_pg, double& off, double off_old, int& ioff_duct,
                                        double& epchk, IpState& s, double& thk, double& etse, double& sigy)
{
  const orgpu_law2& m = g.m2;
  const double SMALL = K_EM7;
  const double young = m.young, gg = m.shear, nu = m.nu;
  const double a11 = or_div(young, (K_ONE - nu * nu));
  const double a12 = a11 * nu;
  const double cn = m.cn;
  const double epdr = fmax(m.epdr * dt1, K_EM20);
  double pla = s.pla;
  double epsd = s.epsd;
  const bool has_temp = m.has_temp != 0;
  const double tempel = has_temp ? s.temp : K_ZERO;
  double z3, z4, m_exp, tstar = K_ZERO;
  if (m.iform == 1) { z3 = m.z3; z4 = m.z4; m_exp = K_ONE; if (has_temp) tstar = fmax(K_ZERO, or_div((tempel - m.tref), fmax(m.tmelt - m.tref, K_EM20))); }
  else { z3 = K_ZERO; z4 = K_ZERO; m_exp = m.z3; tstar = fmax(K_ZERO, or_div((tempel - m.tref), (m.tmelt - m.tref))); }
  double EZZ = K_ZERO, epsdot = K_ZERO;
  if (m.vp == 1) epsdot = epsd * dt1;
  else if (m.vp == 2) { epsd = asrate * epsd_pg + (K_ONE - asrate) * epsd; epsdot = epsd * dt1; }
  else if (m.vp == 3) {
    const double exx = dexx * dtinv, eyy = deyy * dtinv, exy = dexy * dtinv;
    const double DAV = (exx + eyy) * K_THIRD;
    const double D1 = exx - DAV, D2 = eyy - DAV, D3 = -DAV, D4 = K_HALF * exy;
    epsdot = K_HALF * (D1 * D1 + D2 * D2 + D3 * D3) + D4 * D4;
    epsdot = or_div(or_sqrt(K_THREE * epsdot), K_THREE_HALF);
    if (m.israte > 0) epsdot = asrate * epsdot + (K_ONE - asrate) * epsd;
    epsd = epsdot; epsdot = epsdot * dt1;
  }
  // M2CPLR
  double CA = m.ca, CB = m.cb, YMAX = m.sigmx, H = K_ZERO, DPLA = K_ZERO, YLD;
  etse = K_ONE;
  const double FISOKIN = m.fisokin;
  if (FISOKIN > K_ZERO) { s.sxx = s.sxx - s.sbx; s.syy = s.syy - s.sby; s.sxy = s.sxy - s.sbxy; }   // m2cplr.F:115-121
  s.sxx = s.sxx + a11 * dexx + a12 * deyy;
  { const double t = s.syy + a12 * dexx + a11 * deyy; s.syy = t; }
  s.sxy = s.sxy + gg * dexy;
  s.syz = s.syz + gs * deyz;
  s.szx = s.szx + gs * dezx;
  double EPSP = epsdot, Q;
  if (m.cc != K_ZERO) {
    if (m.iform == 0) {
      if (m.israte == 0 && m.vp == 2) EPSP = fmax(fmax(fabs(dexx), fabs(deyy)), K_HALF * fabs(dexy));
      EPSP = fmax(EPSP, epdr);
      const double LOGEP = log(or_div(EPSP, epdr));
      if (tstar == K_ZERO) Q = (K_ONE + m.cc * LOGEP);
      else Q = (K_ONE + m.cc * LOGEP) * (K_ONE - exp(m_exp * log(tstar)));
      Q = fmax(Q, K_EM20);
      CA = CA * Q; CB = CB * Q;
      if (m.icc == 1) YMAX = YMAX * Q;
    } else if (m.iform == 1) {
      if (m.israte == 0 && m.vp == 2) EPSP = fmax(fmax(fabs(dexx), fabs(deyy)), K_HALF * fabs(dexy));
      EPSP = fmax(EPSP, K_EM20);
      Q = log(or_div(EPSP, epdr));
      Q = m.cc * exp((-z3 + z4 * Q) * tempel);
      if (m.icc == 1) YMAX = YMAX + Q;
      CA = CA + Q;
    }
  } else if (m.iform == 0) {
    if (tstar != K_ZERO) { Q = K_ONE - exp(m_exp * log(tstar)); Q = fmax(Q, K_EM20); CA = CA * Q; CB = CB * Q; }
  }
  if (pla == K_ZERO) YLD = CA;
  else { const double BETA = CB * (K_ONE - m.fisokin); YLD = CA + BETA * exp(cn * log(pla)); }
  YLD = fmin(YLD, YMAX);
  if (ipla == 0) {
    const double SVM = or_sqrt(s.sxx * s.sxx + s.syy * s.syy - s.sxx * s.syy + K_THREE * s.sxy * s.sxy);
    const double R = fmin(K_ONE, or_div(YLD, (SVM + K_EM15)));
    if (R < K_ONE) {
      s.sxx = s.sxx * R; s.syy = s.syy * R; s.sxy = s.sxy * R;
      DPLA = off_old * fmax(K_ZERO, or_div((SVM - YLD), young));
      const double S1 = K_HALF * (s.sxx + s.syy);
      EZZ = or_div(DPLA * S1, YLD);
      pla = pla + DPLA;
      epchk = fmax(pla, epchk);
      H = (YLD >= YMAX) ? K_ZERO : cn * CB * exp((cn - K_ONE) * log(pla + SMALL));
      etse = or_div(H, (H + young));
    }
  } else if (ipla == 1) {
    double S1 = s.sxx + s.syy, S2 = s.sxx - s.syy; const double S3 = s.sxy;
    const double A = K_FOURTH * S1 * S1;
    const double B = K_THREE_OVER_4 * S2 * S2 + K_THREE * S3 * S3;
    const double SVM = or_sqrt(A + B);
    if (SVM > YLD && off_old == K_ONE) {
      const double NU1 = or_div(K_ONE, (K_ONE - nu)), NU2 = or_div(K_ONE, (K_ONE + nu));
      H = (YLD >= YMAX) ? K_ZERO : cn * CB * exp((cn - K_ONE) * log(pla + SMALL));
      double DPLA_J = or_div((SVM - YLD), (K_THREE * gg + H));
      etse = or_div(H, (H + young));
      // FISOKIN = 0: m2cplr.F:289-318; kinematic / mixed hardening :319-363 (HI, HK, modified Poisson terms, isotropic share of CB)
      const bool kin = FISOKIN > K_ZERO;
      const double BETAH = H * FISOKIN;
      const double HI = kin ? H - BETAH : H;
      const double AAA = kin ? or_div(K_THREE * (K_TWO_THIRD * BETAH), young) : K_ZERO;
      const double NU11 = NU1 + AAA, NU21 = K_THREE * NU2 + AAA;
      const double ANU1 = kin ? A * NU11 : A * NU1, BNU2 = kin ? B * NU21 : K_THREE * B * NU2, H2 = K_TWO * HI;
      const double CBI = kin ? (K_ONE - FISOKIN) * CB : CB;
      double DPLA_I = K_ZERO, DR = K_ZERO, P = K_ONE, Qq = K_ONE;
      #pragma unroll 1
      for (int N = 0; N < 3; N++) {                       // NMAX = 3 (m2cplr.F:104)
        DPLA_I = DPLA_J;
        const double PLA_I = pla + DPLA_I;
        DPLA = DPLA_J;
        double YLD_I;
        if (PLA_I == K_ZERO) YLD_I = fmin(YMAX, CA);
        else YLD_I = fmin(YMAX, CA + CBI * exp(cn * log(PLA_I)));
        DR = or_div(K_HALF * young * DPLA_I, YLD_I);
        P = or_div(K_ONE, (K_ONE + DR * (kin ? NU11 : NU1)));
        Qq = or_div(K_ONE, (K_ONE + (kin ? DR * NU21 : K_THREE * DR * NU2)));
        const double P2 = P * P, Q2 = Qq * Qq;
        const double F = A * P2 + B * Q2 - YLD_I * YLD_I;
        const double DF = -or_div((ANU1 * P2 * P + BNU2 * Q2 * Qq) * (young - DR * H2), YLD_I) - H2 * YLD_I;
        DPLA_J = (DPLA_I > K_ZERO) ? fmax(K_ZERO, DPLA_I - or_div(F, DF)) : K_ZERO;
      }
      pla = pla + DPLA_I;
      epchk = fmax(pla, epchk);
      S1 = (s.sxx + s.syy) * P;
      S2 = (s.sxx - s.syy) * Qq;
      s.sxx = K_HALF * (S1 + S2);
      s.syy = K_HALF * (S1 - S2);
      s.sxy = s.sxy * Qq;
      EZZ = or_div(DR * S1, young);
      if (kin) YLD = (pla == K_ZERO) ? CA : fmin(YMAX, CA + (K_ONE - FISOKIN) * CB * exp(cn * log(pla)));   // :474-487
    }
  } else {
    const double SVM2 = s.sxx * s.sxx + s.syy * s.syy - s.sxx * s.syy + K_THREE * s.sxy * s.sxy;
    double SVM = or_sqrt(SVM2);
    const double YLD2 = YLD * YLD;
    if (SVM2 > YLD2 && off_old == K_ONE) {
      H = (YLD >= YMAX) ? K_ZERO : cn * CB * exp((cn - K_ONE) * log(pla + SMALL));
      etse = or_div(H, (H + young));
      const double AA = or_div((SVM2 - YLD2), (K_FIVE * SVM2 + K_THREE * (-s.sxx * s.syy + s.sxy * s.sxy)));
      const double S1 = (K_ONE - K_TWO * AA) * s.sxx + AA * s.syy;
      const double S2 = AA * s.sxx + (K_ONE - K_TWO * AA) * s.syy;
      const double S3 = (K_ONE - K_THREE * AA) * s.sxy;
      s.sxx = S1; s.syy = S2; s.sxy = S3;
      DPLA = or_div(off_old * (SVM - YLD), (K_THREE * gg + H));
      pla = pla + DPLA;
      YLD = YLD + H * DPLA;
      SVM = or_sqrt(s.sxx * s.sxx + s.syy * s.syy - s.sxx * s.syy + K_THREE * s.sxy * s.sxy);
      const double R = fmin(K_ONE, or_div(YLD, fmax(K_EM20, SVM)));
      s.sxx = s.sxx * R; s.syy = s.syy * R; s.sxy = s.sxy * R;
      EZZ = or_div(DPLA * K_HALF * (s.sxx + s.syy), YLD);
    }
  }
  if (FISOKIN > K_ZERO) {                                  // m2cplr.F:488-499: back stress along the new shifted stress
    const double ALPHA = or_div(FISOKIN * H * DPLA, YLD);
    s.sbx = s.sbx + ALPHA * s.sxx; s.sby = s.sby + ALPHA * s.syy; s.sbxy = s.sbxy + ALPHA * s.sxy;
    s.sxx = s.sxx + s.sbx; s.syy = s.syy + s.sby; s.sxy = s.sxy + s.sbxy;
  }
  if (m.vp == 1) { epsdot = or_div(DPLA, fmax(K_EM20, dt1)); epsd = asrate * epsdot + (K_ONE - asrate) * epsd; }
  sigy = sigy + or_div(YLD, npttot);
  if (off == off_old && off > K_ZERO) {
    if (off == K_ONE && epchk >= m.epmx) { off = K_FOUR_OVER_5; ioff_duct = 1; }
    else if (off < K_ONE) off = off * K_FOUR_OVER_5;
  }
  EZZ = -(dexx + deyy) * nu - (K_ONE - K_TWO * nu) * EZZ;
  EZZ = or_div(EZZ, (K_ONE - nu));
  thk = thk + EZZ * thklyl * off;
  if (m.rhocp > K_ZERO && has_temp) s.temp = tempel + or_div(sigy * DPLA, m.rhocp);
  s.pla = pla;
  s.epsd = epsd;
}

// ---- CMAIN3 / MULAWC for one element --------------------------------------------------------
// FLAG_ZCFAC: QEPH (JHBE 21..29) keeps SIGY / ZCFAC for the hourglass plasticity (mulawc.F90:521-522).
// NPTC > 0: compile-time point count (loop fully unrolled so independent points overlap in the
// fp64 pipe); NPTC = 0: run-time count.
// FAST = 1: LAW36 with Iplas = 1, IFAIL = 0, one static curve held in the kernel parameters -- all known at compile time (the
// default /PROP/SHELL + /MAT/PLAS_TAB combination): the other return algorithms, the rate interpolation and the global-memory
// curve walk are not compiled in.  The kernels are instruction-fetch sensitive (a 90 KB body walked by 12 warps in different
// phases against a 32 KB L1.5 instruction cache), so less code is measurably faster: -10 % on the plastic C2 plate.
// (Advancing two integration points side by side through the Newton steps was measured too: the extra live state spills,
// 0.589 vs 0.447 ms at 3 CTAs / SM, 0.533 with 248 registers at 2 -- not kept.)
template <int LAW, bool FLAG_ZCFAC, bool STAGED, int NPTC = 0, int FAST = 0>
__device__ __forceinline__ void shell_material_loop(const ShellSG& g, const TileAcc<STAGED>& T, double dt1, MatIO& io)
{
  const int npt = NPTC > 0 ? NPTC : g.prop.npt;
  const double DM = g.prop.dm;
  double* fo = io.fo; double* mo = io.mo;
  #pragma unroll
  for (int k = 0; k < 5; k++) fo[k] = T.ld(SW_FOR + k);
  #pragma unroll
  for (int k = 0; k < 3; k++) mo[k] = T.ld(SW_MOM + k);
  double degmb = fo[0] * io.exx + fo[1] * io.eyy + fo[2] * io.exy + fo[3] * io.eyz + fo[4] * io.exz;
  double degfx = mo[0] * io.kxx + mo[1] * io.kyy + mo[2] * io.kxy;
  const double vol0 = io.area * io.thk0;
  double thkn = T.ld(SW_THK);
  #pragma unroll
  for (int k = 0; k < 5; k++) fo[k] = K_ZERO;
  #pragma unroll
  for (int k = 0; k < 3; k++) mo[k] = K_ZERO;
  double sigy = io.sigy;
  if (LAW == 2 || !FLAG_ZCFAC) sigy = K_ZERO;
  double zcfac1 = K_ZERO, zcfac2 = FLAG_ZCFAC ? K_ONE : K_ZERO, etse = K_ONE;
  double off = io.off; const double off_old = off;
  int ioff_duct = 0;
  double epchk = K_ZERO, viscmx = K_ZERO, ssp = io.ssp;
  const double dtinv = or_div(dt1, fmax(dt1 * dt1, K_EM20));
  const int israte = (LAW != 2) ? g.m36.israte : g.m2.israte;
  const double pm9 = (LAW != 2) ? g.m36.asrate : g.m2.asrate;
  const double asrate = (israte > 0) ? fmin(K_ONE, pm9 * dt1) : K_ONE;
  const int qrow = (npt - 1) * 11;
  const int ipt0 = 0;
  IpState nxt;
  if (!STAGED) nxt = ip_load<LAW, STAGED, FAST>(g, T, ipt0);
#ifdef ORGPU_IP_UNROLL1
  #pragma unroll 1
#else
  #pragma unroll
#endif
  for (int ipt = ipt0; ipt < npt; ipt++) {
    IpState s;
    if (STAGED) s = ip_load<LAW, STAGED, FAST>(g, T, ipt);        // shared memory: no latency to hide
    else { s = nxt; if (ipt + 1 < npt) nxt = ip_load<LAW, STAGED, FAST>(g, T, ipt + 1); }   // software pipeline: next point's state in flight
    const int ipos_old = s.ipos; const double temp_old = s.temp;
    const double pla_old = s.pla;
    const double thkly = c_WF[qrow + ipt];
    const double posly = c_Z0[qrow + ipt] + K_ZERO;
    const double wmc = c_WM[qrow + ipt];
    const double thklyl = thkly * io.thk0;
    const double zt = posly * io.thk0;
    const double dexx = io.exx + zt * io.kxx;
    const double deyy = io.eyy + zt * io.kyy;
    const double dexy = io.exy + zt * io.kxy;
    if (LAW != 2) {
      const int ipla_ = (FAST == 1) ? 1 : g.prop.ipla;
      law36_ip<STAGED, LAW == 37, FAST>(g, T, ipt, ipla_, asrate, dexx, deyy, dexy, io.eyz, io.exz, dtinv, thklyl, io.gs, io.epsd_pg, zt, off,
               s, thkn, ssp, etse, sigy, dt1);
    } else {
      law2_ip(g, g.prop.ipla, npt, dt1, asrate, dexx, deyy, dexy, io.eyz, io.exz, dtinv, thklyl, io.gs, io.epsd_pg,
              off, off_old, ioff_duct, epchk, s, thkn, etse, sigy);
    }
    viscmx = fmax(DM, viscmx);
    if (FAST != 1 && g.fail.irupt == 1) {
      // /FAIL/JOHNSON after the law (mulawc.F90:2064-2069 DPLA = PLA - PLA0, EPSD = LBUF%EPSD; :2118-2127 -> fail_johnson_c.F:111-130;
      // :2608-2616, 2633-2637: the stress kept in LBUF is scaled by SIGOFF, the sums below take this cycle's stress)
      const int w = g.w_ip0 + ipt * g.nwip;
      double dfmax = T.ld(w + g.iw_dfmax), foff = T.ld(w + g.iw_foff);
      const double dpla = s.pla - pla_old;
      if (off == K_ONE && foff == K_ONE && dpla > K_ZERO) {
        const double PR = K_THIRD * (s.sxx + s.syy);
        const double SVM = or_sqrt(s.sxx * s.sxx + s.syy * s.syy - s.sxx * s.syy + K_THREE * s.sxy * s.sxy);
        double EPSF = or_div(g.fail.d3 * PR, fmax(K_EM20, SVM));
        EPSF = (g.fail.d1 + g.fail.d2 * exp(EPSF));
        if (g.fail.d4 != K_ZERO) EPSF = EPSF * (K_ONE + g.fail.d4 * log(fmax(K_ONE, or_div(s.epsd, g.fail.epsp0))));
        EPSF = fmax(EPSF, g.fail.epsf_min);
        if (EPSF > K_ZERO) dfmax = dfmax + or_div(dpla, EPSF);
        if (dfmax >= K_ONE) foff = K_ZERO;
      }
      dfmax = fmin(K_ONE, dfmax);
      T.st(w + g.iw_dfmax, dfmax); T.st(w + g.iw_foff, foff);
      if (foff == K_ZERO) { IpState z = s; z.sxx = s.sxx * K_ZERO; z.syy = s.syy * K_ZERO; z.sxy = s.sxy * K_ZERO; z.syz = s.syz * K_ZERO; z.szx = s.szx * K_ZERO;
                            ip_store<LAW, STAGED, FAST>(g, T, ipt, z, ipos_old, temp_old); }
      else ip_store<LAW, STAGED, FAST>(g, T, ipt, s, ipos_old, temp_old);
    } else
    ip_store<LAW, STAGED, FAST>(g, T, ipt, s, ipos_old, temp_old);
    fo[0] = fo[0] + thkly * s.sxx; fo[1] = fo[1] + thkly * s.syy; fo[2] = fo[2] + thkly * s.sxy;
    fo[3] = fo[3] + thkly * s.syz; fo[4] = fo[4] + thkly * s.szx;
    mo[0] = mo[0] + wmc * s.sxx; mo[1] = mo[1] + wmc * s.syy; mo[2] = mo[2] + wmc * s.sxy;
    if (FLAG_ZCFAC) {
      if (LAW != 2) zcfac1 = zcfac1 + etse * thkly; else zcfac1 = zcfac1 + or_div(etse, npt);
      zcfac2 = fmin(etse, zcfac2);
    }
  }
  if (FAST != 1 && g.fail.irupt == 1 && off == K_ONE) {
    // FAIL_SETOFF_C, one layer (fail_setoff_c.F:155-186): broken share of the thickness / of the points against P_thickfail
    double thfact = K_ZERO, npfail = K_ZERO;
    for (int ipt = 0; ipt < npt; ipt++)
      if (T.ld(g.w_ip0 + ipt * g.nwip + g.iw_foff) < K_ONE) { thfact = thfact + c_WF[qrow + ipt]; npfail = npfail + or_div(K_ONE, npt); }
    const double pthkf = g.fail_pthkf;
    if ((thfact >= pthkf && pthkf > K_ZERO) || (npfail >= fabs(pthkf) && pthkf < K_ZERO)) off = K_FOUR_OVER_5;
  }
  if ((off == K_FOUR_OVER_5 && ioff_duct == 0) || (off > K_ZERO && off_old < K_EM01)) off = K_ZERO;
  T.st(SW_THK, fmax(thkn, K_EM30));
  const double fact = K_ONEP414 * DM;
  const double visc = fact * ssp * or_sqrt(io.area) * dtinv * io.rho;
  fo[0] = fo[0] + visc * (io.exx + K_HALF * io.eyy);
  fo[1] = fo[1] + visc * (io.eyy + K_HALF * io.exx);
  fo[2] = fo[2] + visc * io.exy * K_THIRD;
  #pragma unroll
  for (int k = 0; k < 5; k++) fo[k] = fo[k] * off;
  #pragma unroll
  for (int k = 0; k < 3; k++) mo[k] = mo[k] * off;
  degmb = degmb + fo[0] * io.exx + fo[1] * io.eyy + fo[2] * io.exy + fo[3] * io.eyz + fo[4] * io.exz;
  degfx = degfx + mo[0] * io.kxx + mo[1] * io.kyy + mo[2] * io.kxy;
  const double vol2 = K_HALF * vol0;
  T.st(SW_EINT, T.ld(SW_EINT) + degmb * vol2);
  T.st(SW_EINT + 1, T.ld(SW_EINT + 1) + degfx * io.thk0 * vol2);
  #pragma unroll
  for (int k = 0; k < 5; k++) T.st(SW_FOR + k, fo[k]);
  #pragma unroll
  for (int k = 0; k < 3; k++) T.st(SW_MOM + k, mo[k]);
  io.off = off; io.ssp = ssp; io.viscmx = viscmx; io.sigy = sigy; io.zcfac1 = zcfac1; io.zcfac2 = zcfac2; io.vol0 = vol0;
}

// Yield stress and slope of a LISTED point once more (pass 2 of the three-pass loop), any NRATE, FISOKIN = 0, no damage factor:
// the same statements as law36_trial (sigeps36c.F:317-404) on the point's stored strain rate and plastic strain; the table
// cursors of the point -- another lane's -- were advanced by pass 1 and are only read here (the walks are idempotent).
__device__ __forceinline__ void law36_yield_again(const ShellSG& g, const TileAcc<true>& T, int ip, int tid, double epsd, double pla, double& YLD, double& H)
{
  const orgpu_law36& m = g.m36;
  if (m.nrate == 1) {
    int ipos = T.ldi_lane(g.w_vt, ip, tid);
    const int f = m.ifunc[0];
    double dydx, y1;
    if (g.ct.n > 0) vinter1c(g.ct, 0, ipos, pla, dydx, y1);
    else { const int i0 = __ldg(g.npf + f), i1 = __ldg(g.npf + f + 1); vinter1(g.tf, i0, i1 - i0, ipos, pla, dydx, y1); }
    const double FACT = K_ONE * K_ONE * (m.yfac[0] * K_ONE);
    H = dydx * FACT; YLD = y1 * FACT;
  } else {
    int JJ = 1;
    for (int J = 2; J <= m.nrate - 1; J++) if (epsd >= m.rate[J - 1]) JJ = J;
    double RFAC;
    if (m.ismooth == 2) {
      const double EPSP1 = fmax(m.rate[JJ - 1], K_EM20), EPSP2 = m.rate[JJ];
      RFAC = or_div(log(or_div(fmax(epsd, K_EM20), EPSP1)), log(or_div(EPSP2, EPSP1)));
    } else {
      const double EPSP1 = m.rate[JJ - 1], EPSP2 = m.rate[JJ];
      RFAC = or_div((epsd - EPSP1), (EPSP2 - EPSP1));
    }
    const double YFAC1 = m.yfac[JJ - 1] * K_ONE, YFAC2 = m.yfac[JJ] * K_ONE;
    const int f1 = m.ifunc[JJ - 1], f2 = m.ifunc[JJ];
    int ipos1 = T.ldb_lane(g.w_vt, ip * g.nvt + JJ - 1, tid), ipos2 = T.ldb_lane(g.w_vt, ip * g.nvt + JJ, tid);
    double dydx1, y1, dydx2, y2;
    if (g.ct.n > 0) { vinter1c(g.ct, JJ - 1, ipos1, pla, dydx1, y1); vinter1c(g.ct, JJ, ipos2, pla, dydx2, y2); }
    else {
      { const int i0 = __ldg(g.npf + f1), i1 = __ldg(g.npf + f1 + 1); vinter1(g.tf, i0, i1 - i0, ipos1, pla, dydx1, y1); }
      { const int i0 = __ldg(g.npf + f2), i1 = __ldg(g.npf + f2 + 1); vinter1(g.tf, i0, i1 - i0, ipos2, pla, dydx2, y2); }
    }
    y1 = y1 * YFAC1; y2 = y2 * YFAC2;
    YLD = K_ONE * (y1 + RFAC * (y2 - y1));
    YLD = fmax(YLD, K_EM20);
    dydx1 = dydx1 * YFAC1; dydx2 = dydx2 * YFAC2;
    H = K_ONE * (dydx1 + RFAC * (dydx2 - dydx1));
    YLD = YLD * fmax(K_ZERO, K_ONE);
    H = H * fmax(K_ZERO, K_ONE);
  }
  if (m.yldcheck == 1) YLD = fmax(YLD, K_EM20);
  H = fmax(K_ZERO, H);
}


// ---- the through-thickness loop of the FAST = 1 kernels, in three passes ---------------------------------------------------
// The Newton return of Iplas = 1 (sigeps36c.F:503-593: three steps, 13 divisions in series) costs a warp its full length for
// every point at which ANY of its lanes yields.  In a deck that yields locally -- the usual state of a crash model, and of the
// C2 plate after its first 200 cycles -- 10-30 % of the points yield in a cycle, spread so that a warp pays 2-5 returns per
// element for what fills one or two.  So:
//   pass 1, every lane, point by point: elastic predictor, strain rate, yield stress, the yield test; the trial state goes
//           back into the staged tile and (lane, point) of every yielding point is appended to a byte list of the warp;
//   pass 2, the warp walks that list DENSELY, one yielding point per lane whoever owns it: re-reads the trial stress of
//           (lane, point) from the tile, redoes the curve lookup (same cursor, same arithmetic) and the Newton return, writes
//           the returned stress and plastic strain back and leaves the thickness-strain term for the owner;
//   pass 3, every lane again, point by point: thickness, force and moment sums in the reference's order.
// Every value is produced by the same operations on the same operands as in law36_ip, so the results are bit-identical; the
// list and the hand-back slots are the tile words of GBUF%FOR / GBUF%MOM, which are dead between the loop's first read and its
// last store.  Needs the staged tile (cross-lane reads) and NPT <= 5 (slots): shell_fast() on the host.
// Measured on C2 (B200, kernel ms; profiles/r02_qeph_forces_ncu.md): 10-30 % of the points yielding 0.437 -> 0.392, no point
// yielding 0.321 -> 0.323, 60-90 % yielding (the plate's first 150 cycles) 0.394 -> 0.418.  Variants that did not pay: two /
// three / four listed points per lane side by side (spills: 0.427 / 0.444 / 0.494 in the 10-30 % state), rows most of whose
// lanes yield returning on the spot (a second copy of the return in the code: +3 % even where no point yields -- the kernel is
// instruction-fetch sensitive), the same with one copy of the return in a merged loop of turns (0.405, and 0.331 elastic).
// FAST = 1: one static curve in the kernel parameters (everything about the law known at compile time); FAST = 2: any number of rate
// curves, in the parameters or in global memory, strain-rate filter -- what /MAT/PLAS_TAB decks with rate dependence use.
template <bool FLAG_ZCFAC, int FAST = 1>
__device__ __forceinline__ void shell_material_loop_compact(const ShellSG& g, const TileAcc<true>& T, double dt1, MatIO& io, unsigned wmask, bool full_tile = false)
{
  const orgpu_law36& m = g.m36;
  const int npt = g.prop.npt;
  const double DM = g.prop.dm, E = m.young, G3 = m.g3, NU3 = K_ONE - m.nu_mnu;
  double* fo = io.fo; double* mo = io.mo;
  #pragma unroll
  for (int k = 0; k < 5; k++) fo[k] = T.ld(SW_FOR + k);
  #pragma unroll
  for (int k = 0; k < 3; k++) mo[k] = T.ld(SW_MOM + k);
  double degmb = fo[0] * io.exx + fo[1] * io.eyy + fo[2] * io.exy + fo[3] * io.eyz + fo[4] * io.exz;
  double degfx = mo[0] * io.kxx + mo[1] * io.kyy + mo[2] * io.kxy;
  const double vol0 = io.area * io.thk0;
  double sigy = io.sigy;
  if (!FLAG_ZCFAC) sigy = K_ZERO;
  double zcfac1 = K_ZERO, zcfac2 = FLAG_ZCFAC ? K_ONE : K_ZERO;
  double off = io.off; const double off_old = off;
  double viscmx = K_ZERO;
  const double dtinv = or_div(dt1, fmax(dt1 * dt1, K_EM20));
  const double asrate = (m.israte > 0) ? fmin(K_ONE, m.asrate * dt1) : K_ONE;
  const int qrow = (npt - 1) * 11;
  const int lane = threadIdx.x & 31;
  const int nact = __popc(wmask);                         // the lanes with an element are the low ones
  double* const wbase = T.t - lane;                       // lane 0 of this warp
  unsigned char* const list = reinterpret_cast<unsigned char*>(wbase + SW_FOR * ORGPU_TILE);   // 256 bytes: word FOR(1) of the 32 lanes
  __syncwarp(wmask);                                      // every lane has read its FOR / MOM words
  // ---- pass 1
  unsigned pmask = 0; int cnt = 0;                  // points of this lane that yield; entries of the warp's list
  #pragma unroll 1
  for (int ipt = 0; ipt < npt; ipt++) {
    IpState s = ip_load<36, true, FAST>(g, T, ipt);
    const int ipos_old = s.ipos;
    const double thkly = c_WF[qrow + ipt];
    const double zt = (c_Z0[qrow + ipt] + K_ZERO) * io.thk0;
    const double dexx = io.exx + zt * io.kxx, deyy = io.eyy + zt * io.kyy, dexy = io.exy + zt * io.kxy;
    double YLD, H, EPST;
    law36_trial<true, false, FAST>(g, T, ipt, asrate, dexx, deyy, dexy, io.eyz, io.exz, dtinv, io.gs, io.epsd_pg, zt, s, YLD, H, EPST);
    H = fmax(K_ZERO, H);
    const double S1 = s.sxx + s.syy, S2 = s.sxx - s.syy, S3 = s.sxy;
    const double SVM2 = K_FOURTH * S1 * S1 + (K_THREE_OVER_4 * S2 * S2 + K_THREE * S3 * S3);
    const bool pl = SVM2 > YLD * YLD && off == K_ONE;
    double etse = K_ONE;
    if (pl) etse = or_div(H, (H + E));
    if (FLAG_ZCFAC) { zcfac1 = zcfac1 + etse * thkly; zcfac2 = fmin(etse, zcfac2); }
    viscmx = fmax(DM, viscmx);
    ip_store<36, true, FAST>(g, T, ipt, s, ipos_old, K_ZERO);         // trial state (PLA unchanged)
    const unsigned b = __ballot_sync(wmask, pl);
    if (pl) { list[cnt + __popc(b & ((1u << lane) - 1u))] = (unsigned char)(lane | (ipt << 5)); pmask |= 1u << ipt; }
    cnt += __popc(b);
    sigy = YLD;
  }
  __syncwarp(wmask);
  PHASE_SYNC(7);
  // ---- pass 2: one listed point per lane and turn (sigeps36c.F:503-593; FISOKIN = 0: HK = 0, AAA = 0 exactly, :520-527)
  #pragma unroll 1
  for (int k = lane; k < cnt; k += nact) {
    const int it = list[k], l = it & 31, ip = it >> 5;
    double* const own = wbase + l;                                   // word 0 of the point's owner
    double* const q = own + (size_t)(g.w_ip0 + ip * g.nwip) * ORGPU_TILE;
    const double sxx = q[IW_SIG * ORGPU_TILE], syy = q[(IW_SIG + 1) * ORGPU_TILE], sxy = q[(IW_SIG + 2) * ORGPU_TILE], pla = q[IW_PLA * ORGPU_TILE];
    double H, YLD;
    if (FAST == 1) {
      int ipos = T.ldi_lane(g.w_vt, ip, threadIdx.x - lane + l);
      double dydx, y1;
      vinter1c(g.ct, 0, ipos, pla, dydx, y1);                        // the cursor already stands on the segment: same slope, same value
      const double FACT = K_ONE * K_ONE * (m.yfac[0] * K_ONE);
      H = dydx * FACT; YLD = y1 * FACT;
      if (m.yldcheck == 1) YLD = fmax(YLD, K_EM20);
      H = fmax(K_ZERO, H);
    } else law36_yield_again(g, T, ip, threadIdx.x - lane + l, q[IW_EPSD * ORGPU_TILE], pla, YLD, H);
    double S1 = sxx + syy, S2 = sxx - syy;
    L36Newton n;
    n.AA = K_FOURTH * S1 * S1;
    n.BB = K_THREE_OVER_4 * S2 * S2 + K_THREE * sxy * sxy;
    const double SVM = or_sqrt(n.AA + n.BB);
    n.YLD = YLD; n.HI = H; n.NU11 = m.u_mnu; n.NU21 = m.t_pnu;
    n.DPLA_J = or_div((SVM - YLD), (G3 + H));
    n.DPLA_I = K_ZERO; n.DR = K_ZERO; n.PP = K_ONE; n.QQ = K_ONE;
    law36_newton_step(E, n, false); law36_newton_step(E, n, false); law36_newton_step(E, n, true);
    S1 = S1 * n.PP;
    S2 = S2 * n.QQ;
    q[IW_SIG * ORGPU_TILE] = K_HALF * (S1 + S2);
    q[(IW_SIG + 1) * ORGPU_TILE] = K_HALF * (S1 - S2);
    q[(IW_SIG + 2) * ORGPU_TILE] = sxy * n.QQ;
    q[IW_PLA * ORGPU_TILE] = pla + n.DPLA_I;
    own[(SW_FOR + 1 + ip) * ORGPU_TILE] = -or_div(NU3 * n.DR * S1, E);             // DEZZ of the return, for the owner
    if (ip == npt - 1) own[(SW_MOM + 2) * ORGPU_TILE] = YLD + n.HI * n.DPLA_I;
  }
  __syncwarp(wmask);
  PHASE_SYNC(7);
  // ---- pass 3
  double thkn = T.ld(SW_THK);
  if ((pmask >> (npt - 1)) & 1u) sigy = T.ld(SW_MOM + 2);
  #pragma unroll
  for (int k = 0; k < 5; k++) fo[k] = K_ZERO;
  #pragma unroll
  for (int k = 0; k < 3; k++) mo[k] = K_ZERO;
  #pragma unroll 1
  for (int ipt = 0; ipt < npt; ipt++) {
    const int w = g.w_ip0 + ipt * g.nwip;
    const double thkly = c_WF[qrow + ipt], wmc = c_WM[qrow + ipt];
    const double thklyl = thkly * io.thk0;
    const double zt = (c_Z0[qrow + ipt] + K_ZERO) * io.thk0;
    const double dexx = io.exx + zt * io.kxx, deyy = io.eyy + zt * io.kyy;
    { const double DEZZ = -(dexx + deyy) * m.nu_mnu; thkn = thkn + DEZZ * thklyl * off; }
    if ((pmask >> ipt) & 1u) thkn = thkn + T.ld(SW_FOR + 1 + ipt) * thklyl * off;
    const double sxx = T.ld(w + IW_SIG), syy = T.ld(w + IW_SIG + 1), sxy = T.ld(w + IW_SIG + 2), syz = T.ld(w + IW_SIG + 3), szx = T.ld(w + IW_SIG + 4);
    fo[0] = fo[0] + thkly * sxx; fo[1] = fo[1] + thkly * syy; fo[2] = fo[2] + thkly * sxy;
    fo[3] = fo[3] + thkly * syz; fo[4] = fo[4] + thkly * szx;
    mo[0] = mo[0] + wmc * sxx; mo[1] = mo[1] + wmc * syy; mo[2] = mo[2] + wmc * sxy;
  }
  if (off == K_FOUR_OVER_5 || (off > K_ZERO && off_old < K_EM01)) off = K_ZERO;
  T.st(SW_THK, fmax(thkn, K_EM30));
  const double fact = K_ONEP414 * DM;
  const double visc = fact * m.soundsp * or_sqrt(io.area) * dtinv * io.rho;
  fo[0] = fo[0] + visc * (io.exx + K_HALF * io.eyy);
  fo[1] = fo[1] + visc * (io.eyy + K_HALF * io.exx);
  fo[2] = fo[2] + visc * io.exy * K_THIRD;
  #pragma unroll
  for (int k = 0; k < 5; k++) fo[k] = fo[k] * off;
  #pragma unroll
  for (int k = 0; k < 3; k++) mo[k] = mo[k] * off;
  degmb = degmb + fo[0] * io.exx + fo[1] * io.eyy + fo[2] * io.exy + fo[3] * io.eyz + fo[4] * io.exz;
  degfx = degfx + mo[0] * io.kxx + mo[1] * io.kyy + mo[2] * io.kxy;
  const double vol2 = K_HALF * vol0;
  T.st(SW_EINT, T.ld(SW_EINT) + degmb * vol2);
  T.st(SW_EINT + 1, T.ld(SW_EINT + 1) + degfx * io.thk0 * vol2);
  __syncwarp(wmask);                                      // the list and the slots have been read by everyone: FOR / MOM get their values
  #pragma unroll
  for (int k = 0; k < 5; k++) T.st(SW_FOR + k, fo[k]);
  #pragma unroll
  for (int k = 0; k < 3; k++) T.st(SW_MOM + k, mo[k]);
  io.off = off; io.ssp = m.soundsp; io.viscmx = viscmx; io.sigy = sigy; io.zcfac1 = zcfac1; io.zcfac2 = zcfac2; io.vol0 = vol0;
}
