// bt_kernel.cuh -- fused internal-force kernel for Belytschko-Tsay 4-node shells (Ishell 1, 3, 4;
// NPT > 1), one element per thread.  One launch does for every element of a super-group what CFORC3
// (engine/source/elements/shell/coque/cforc3.F:403-751, ISHFRAM=0) does per group:
//   CCOOR3 (ccoor3.F:60-140) gather X,V,VR  ->  CNVEC3 (cnvec3.F:75-141) convected frame
//   CDERI3 (cderi3.F:85-190) local coords, small-strain reference, PX/PY, AREA, VHX/VHY
//   CCOEF3 (ccoef3.F:70-190), CDLEN3 (cdlen3.F:60-125), CDEFO3 (cdefo3.F:65-175, IHBE branches),
//   CCURV3 (ccurv3.F:60-100), CSTRA3 (cstra3.F:85-215), epsd_pg (cforc3.F:533-552),
//   CMAIN3/MULAWC + law (shell_common.cuh), CHVIS3 (chvis3.F:120-420), CDT3 (cdt3.F:111-232),
//   CFINT3 (cfint3.F:147-236), CUPDT3P (cupdt3.F:1017-1175): 4 corner rows into FSKY(8,IADC)
// then the CTA (dt, user id) arg-min (strict "<", first minimum wins: cdt3.F:205-216).
#pragma once
#include "shell_common.cuh"

template <int LAW, bool STAGED, int FAST = 0, bool TAB = false>
__global__ void __launch_bounds__(ORGPU_SHELL_CTA, 3 * ORGPU_PER128)
bt_forces_kernel(const __grid_constant__ ShellParams P)
{
  if (P.cs->abort) return;                               // sticky: a peer-memory wait timed out (exchange.cuh)
  const CtaWork<ShellSG> W = cta_work<ShellSG, TAB>(P.sg, P.sgtab, P.cta_map);
  const ShellSG& g = *W.g;
  const int tile = W.tile;
  const int e = tile * ORGPU_TILE + threadIdx.x;
  __shared__ __align__(8) unsigned long long s_bar;
  double* const g_tile = g.slab + (size_t)tile * g.nw * ORGPU_TILE;
  const double* const g_pf = W.tile_pf >= 0 ? W.g_pf->slab + (size_t)W.tile_pf * W.g_pf->nw * ORGPU_TILE : nullptr;   // tile of the CTA one wave ahead
  if (STAGED) tile_load_begin(s_tile_dyn, &s_bar, g_tile, (unsigned)g.nw * ORGPU_TILE * 8u, g_pf);
  const TileAcc<STAGED> T{(STAGED ? s_tile_dyn : g_tile) + threadIdx.x};
  if (!STAGED && threadIdx.x == 0 && g_pf)              // in-place tiles: same wave-ahead L2 prefetch
    bulk_prefetch_l2(g_pf, (unsigned)g.nw * ORGPU_TILE * 8u);
  double* const sm = g.smstr + (size_t)tile * 6 * ORGPU_TILE + threadIdx.x;      // SMSTR word k at sm[k*128]
#if ORGPU_PREFETCH_NEXT > 0
  // a CTA about one wave ahead: start its connectivity toward L2 (its first load is then an L2 hit: -3 % kernel time)
  if (W.tile_nx >= 0 && threadIdx.x < (4 * ORGPU_TILE * 4) / 128) prefetch_l2(reinterpret_cast<const char*>(W.g_nx->conn + (size_t)W.tile_nx * 4 * ORGPU_TILE) + 128 * threadIdx.x);
#endif
  double dt_cand = K_EP30; int order = 0x7fffffff;
  const unsigned wmask = (FAST >= 1) ? __ballot_sync(0xffffffffu, e < g.ne) : 0u;     // the warp's lanes that own an element
  if (e < g.ne) {
    const double DT1 = P.cs->dt2;
    const int ISMSTR = g.prop.ismstr, NPT = g.prop.npt, IHBE = g.prop.ihbe;
    int nc[4];
    { const int* cn = g.conn + (size_t)tile * 4 * ORGPU_TILE + threadIdx.x;
      #pragma unroll
      for (int k = 0; k < 4; k++) nc[k] = __ldg(cn + k * ORGPU_TILE); }
    order = g.order0 + e;
    double xg[4], yg[4], zg[4];
    #pragma unroll
    for (int k = 0; k < 4; k++) { const double4 p = ld256_nc(P.nd.pos + nc[k]); xg[k] = p.x; yg[k] = p.y; zg[k] = p.z; }
    #pragma unroll
    for (int k = 0; k < 4; k++) { prefetch_l1(P.nd.rot + nc[k]); prefetch_l1(P.nd.vel + nc[k]); }
    if (STAGED) mbar_wait(&s_bar, 0);                     // the state tile has landed (issued before the gather)
    double OFFG = T.ld(SW_OFF);
    const bool dead_in = OFFG < K_ZERO;
    double OFF = fmin(K_ONE, fabs(OFFG));
    // ---- frame (CNVEC3) from the four corner positions
    double e1[3], e2[3], e3[3];
    double X2, Y2, X3, Y3, X4, Y4, Z2;
    {
      const double X21 = xg[1] - xg[0], X32 = xg[2] - xg[1], X34 = xg[2] - xg[3], X41 = xg[3] - xg[0];
      const double Y21 = yg[1] - yg[0], Y32 = yg[2] - yg[1], Y34 = yg[2] - yg[3], Y41 = yg[3] - yg[0];
      const double Z21 = zg[1] - zg[0], Z32 = zg[2] - zg[1], Z34 = zg[2] - zg[3], Z41 = zg[3] - zg[0];
      e1[0] = (X21 + X34); e1[1] = (Y21 + Y34); e1[2] = (Z21 + Z34);
      e2[0] = (X32 + X41); e2[1] = (Y32 + Y41); e2[2] = (Z32 + Z41);
      e3[0] = e1[1] * e2[2] - e1[2] * e2[1]; e3[1] = e1[2] * e2[0] - e1[0] * e2[2]; e3[2] = e1[0] * e2[1] - e1[1] * e2[0];
      double S = e3[0] * e3[0] + e3[1] * e3[1] + e3[2] * e3[2];
      S = or_div(K_ONE, fmax(or_sqrt(S), K_EM20));
      e3[0] = e3[0] * S; e3[1] = e3[1] * S; e3[2] = e3[2] * S;
      const double S1 = e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2], S2 = e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2];
      S = or_sqrt(or_div(S1, S2));
      e1[0] = e1[0] + (e2[1] * e3[2] - e2[2] * e3[1]) * S;
      e1[1] = e1[1] + (e2[2] * e3[0] - e2[0] * e3[2]) * S;
      e1[2] = e1[2] + (e2[0] * e3[1] - e2[1] * e3[0]) * S;
      S = e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2];
      S = or_div(K_ONE, fmax(or_sqrt(S), K_EM20));
      e1[0] = e1[0] * S; e1[1] = e1[1] * S; e1[2] = e1[2] * S;
      e2[0] = e3[1] * e1[2] - e3[2] * e1[1]; e2[1] = e3[2] * e1[0] - e3[0] * e1[2]; e2[2] = e3[0] * e1[1] - e3[1] * e1[0];
      // CDERI3: local coordinates relative to node 1
      const double X31 = xg[2] - xg[0], Y31 = yg[2] - yg[0], Z31 = zg[2] - zg[0];
      X2 = e1[0] * X21 + e1[1] * Y21 + e1[2] * Z21; Y2 = e2[0] * X21 + e2[1] * Y21 + e2[2] * Z21;
      Y3 = e2[0] * X31 + e2[1] * Y31 + e2[2] * Z31; X3 = e1[0] * X31 + e1[1] * Y31 + e1[2] * Z31;
      X4 = e1[0] * X41 + e1[1] * Y41 + e1[2] * Z41; Y4 = e2[0] * X41 + e2[1] * Y41 + e2[2] * Z41;
      Z2 = e3[0] * X21 + e3[1] * Y21 + e3[2] * Z21;
    }
    if (ISMSTR == 1 || ISMSTR == 2) {
      if (fabs(OFFG) == K_TWO) {
        X2 = sm[0]; Y2 = sm[ORGPU_TILE]; X3 = sm[2 * ORGPU_TILE];
        Y3 = sm[3 * ORGPU_TILE]; X4 = sm[4 * ORGPU_TILE]; Y4 = sm[5 * ORGPU_TILE]; Z2 = K_ZERO;
      } else {
        __stcs(&sm[0], X2); __stcs(&sm[ORGPU_TILE], Y2); __stcs(&sm[2 * ORGPU_TILE], X3);
        __stcs(&sm[3 * ORGPU_TILE], Y3); __stcs(&sm[4 * ORGPU_TILE], X4); __stcs(&sm[5 * ORGPU_TILE], Y4);
      }
      if (ISMSTR == 1 && OFFG == K_ONE) OFFG = K_TWO;
    }
    const double PX1 = K_HALF * (Y2 - Y4), PY1 = K_HALF * (X4 - X2), PX2 = K_HALF * Y3, PY2 = -K_HALF * X3;
    const double AREA = fmax(K_TWO * (PY2 * PX1 - PY1 * PX2), K_EM20);
    const double VHX = or_div((-X2 + X3 - X4), AREA), VHY = or_div((-Y2 + Y3 - Y4), AREA);
    // ---- CCOEF3
    const double THK0 = (g.prop.ithk > 0) ? T.ld(SW_THK) : T.ld(g.w_thke);
    const double THK02 = THK0 * THK0;
    double RHO, YM, NU, G;
    MatIO io;
    if (LAW != 2) { const orgpu_law36& m = g.m36; RHO = m.rho0; YM = m.young; NU = m.nu; G = m.shear; io.ssp = m.ssp; }
    else           { const orgpu_law2& m = g.m2;   RHO = m.rho0; YM = m.young; NU = m.nu; G = m.shear; io.ssp = m.ssp; }
    const double H1 = g.prop.h1, H2 = g.prop.h2, H3 = g.prop.h3;
    double SHF = K_ZERO;
    if (NPT != 1) { const double FAC1TMP = 2. * (1. + NU) * THK02; const int ISH = 0; const double FSH = g.prop.shf;
                    SHF = FSH * (1. - ISH + or_div(ISH * FAC1TMP, (FSH * AREA + FAC1TMP))); }
    // ---- CDLEN3
    double ALDT;
    {
      const double AL1 = X2 * X2 + Y2 * Y2;
      const double AL2 = (X3 - X2) * (X3 - X2) + (Y3 - Y2) * (Y3 - Y2);
      const double AL6 = X3 * X3 + Y3 * Y3;
      const double AL3 = (X4 - X3) * (X4 - X3) + (Y4 - Y3) * (Y4 - Y3);
      const double AL4 = X4 * X4 + Y4 * Y4;
      const double AL5 = (X4 - X2) * (X4 - X2) + (Y4 - Y2) * (Y4 - Y2);
      double ALMIN = fmin(fmin(AL1, AL2), AL4);
      const double ALQUAD = fmin(fmin(AL3, AL5), AL6);
      if (AL3 != K_ZERO) ALMIN = fmin(ALMIN, ALQUAD);
      const double DTDYN = or_div(AREA * AREA, fmax(fmax(AL5, AL6), K_EM20));
      ALDT = fmax(DTDYN, ALMIN);
      const double DTHOUR = or_div(K_HALF * (ALMIN + ALDT), fmax(H1, H2));
      if (IHBE != 0) { if (DTHOUR < ALDT) ALDT = DTHOUR; } else ALDT = fmin(ALDT, DTHOUR);
      ALDT = or_sqrt(ALDT);
    }
    // ---- CDEFO3: nodal velocities in the local frame, membrane + transverse shear rates
    double VX[4], VY[4], VZ[4], EXX, EYY, EXY, EXZ, EYZ;
    {
      #pragma unroll
      for (int k = 0; k < 4; k++) {
        double4 v = ld256_nc(P.nd.vel + nc[k]);
        if (dead_in) { v.x = K_ZERO; v.y = K_ZERO; v.z = K_ZERO; }
        VX[k] = e1[0] * v.x + e1[1] * v.y + e1[2] * v.z;
        VY[k] = e2[0] * v.x + e2[1] * v.y + e2[2] * v.z;
        VZ[k] = e3[0] * v.x + e3[1] * v.y + e3[2] * v.z;
      }
      const double VZ13 = VZ[0] - VZ[2], VZ24 = VZ[1] - VZ[3];
      EYZ = PY1 * VZ13 + PY2 * VZ24;
      EXZ = PX1 * VZ13 + PX2 * VZ24;
      double VX13, VX24, VY13, VY24;
      if (IHBE <= 1) {
        Z2 = K_ZERO;
        const double DT1V4 = K_FOURTH * DT1;
        const double TMP2A = PY2 + PY1;
        const double TMP3A = copysign(fmax(fabs(TMP2A), K_EM20), TMP2A);
        const double TMP1A = or_div(DT1V4 * (VZ13 - VZ24) * (VZ13 - VZ24), TMP3A);
        VX13 = VX[0] - VX[2]; VX24 = VX[1] - VX[3];
        VX13 = VX13 - TMP1A; VX24 = VX24 + TMP1A;
        const double TMP1B = PX2 - PX1;
        const double TMP3B = copysign(fmax(fabs(TMP1B), K_EM20), TMP1B);
        const double TMP2B = or_div(DT1V4 * (VZ13 + VZ24) * (VZ13 + VZ24), TMP3B);
        VY13 = VY[0] - VY[2]; VY24 = VY[1] - VY[3];
        VY13 = VY13 + TMP2B; VY24 = VY24 + TMP2B;
      } else if (IHBE == 2 || IHBE == 3) {
        const double DT1V4 = K_HALF * DT1;
        const double GZX = or_div(EXZ, AREA), EXZZ2 = GZX * Z2, EXZ2 = GZX * GZX * DT1V4;
        VX[2] = VX[2] - EXZ2 * X3 - VX[0];
        VX[1] = VX[1] + EXZZ2 - EXZ2 * X2 - VX[0];
        VX[3] = VX[3] + EXZZ2 - EXZ2 * X4 - VX[0];
        VX[0] = K_ZERO;
        const double GZY = or_div(EYZ, AREA), EYZZ2 = GZY * Z2, EYZ2 = GZY * GZY * DT1V4;
        VY[2] = VY[2] - EYZ2 * Y3 - VY[0];
        VY[1] = VY[1] + EYZZ2 - EYZ2 * Y2 - VY[0];
        VY[3] = VY[3] + EYZZ2 - EYZ2 * Y4 - VY[0];
        VY[0] = K_ZERO;
        const double ZZZ = (EXZ2 + EYZ2) * Z2;
        VZ[2] = VZ[2] - GZY * Y3 - GZX * X3 - VZ[0];
        VZ[1] = VZ[1] - GZY * Y2 - GZX * X2 - ZZZ - VZ[0];
        VZ[3] = VZ[3] - GZY * Y4 - GZX * X4 - ZZZ - VZ[0];
        VZ[0] = K_ZERO;
        VX13 = -VX[2]; VX24 = VX[1] - VX[3];
        VY13 = -VY[2]; VY24 = VY[1] - VY[3];
      } else {
        const double DT1V4 = K_HALF * DT1;
        const double ZZ2 = K_HALF * Z2;
        const double GZX = or_div(EXZ, AREA), EXZZ2 = GZX * ZZ2, EXZ2 = GZX * GZX * DT1V4, EXZ2PY2 = EXZ2 * PY2, EXZ2PY1 = EXZ2 * PY1;
        VX[0] = VX[0] - EXZZ2 - EXZ2PY2; VX[2] = VX[2] - EXZZ2 + EXZ2PY2;
        VX[1] = VX[1] + EXZZ2 + EXZ2PY1; VX[3] = VX[3] + EXZZ2 - EXZ2PY1;
        const double GZY = or_div(EYZ, AREA), EYZZ2 = GZY * ZZ2, EYZ2 = GZY * GZY * DT1V4, EYZ2PX2 = EYZ2 * PX2, EYZ2PX1 = EYZ2 * PX1;
        VY[0] = VY[0] - EYZZ2 + EYZ2PX2; VY[2] = VY[2] - EYZZ2 - EYZ2PX2;
        VY[1] = VY[1] + EYZZ2 - EYZ2PX1; VY[3] = VY[3] + EYZZ2 + EYZ2PX1;
        VX13 = VX[0] - VX[2]; VX24 = VX[1] - VX[3];
        VY13 = VY[0] - VY[2]; VY24 = VY[1] - VY[3];
      }
      EXX = PX1 * VX13 + PX2 * VX24;
      EXY = PY1 * VX13 + PY2 * VX24;
      EXY = EXY + PX1 * VY13 + PX2 * VY24;
      EYY = PY1 * VY13 + PY2 * VY24;
    }
    // ---- CCURV3
    double RX[4], RY[4], KXX, KYY, KXY;
    {
      #pragma unroll
      for (int k = 0; k < 4; k++) {
        double4 w = ld256_nc(P.nd.rot + nc[k]);
        if (dead_in) { w.x = K_ZERO; w.y = K_ZERO; w.z = K_ZERO; }
        RX[k] = e1[0] * w.x + e1[1] * w.y + e1[2] * w.z;
        RY[k] = e2[0] * w.x + e2[1] * w.y + e2[2] * w.z;
      }
      const double RX13 = RX[0] - RX[2], RXAV = RX[0] + RX[1] + RX[2] + RX[3], RX24 = RX[1] - RX[3];
      KYY = -PY1 * RX13 - PY2 * RX24;
      KXY = PX1 * RX13 + PX2 * RX24;
      const double RY13 = RY[0] - RY[2], RYAV = RY[0] + RY[1] + RY[2] + RY[3], RY24 = RY[1] - RY[3];
      KXX = PX1 * RY13 + PX2 * RY24;
      KXY = PY1 * RY13 + PY2 * RY24 - KXY;
      EXZ = EXZ + RYAV * (.25 * AREA);
      EYZ = EYZ - RXAV * (.25 * AREA);
    }
    // ---- CSTRA3 + element strain rate
    {
      const double FAC1 = or_div(DT1, AREA);
      io.exx = EXX * FAC1; io.eyy = EYY * FAC1; io.exy = EXY * FAC1; io.eyz = EYZ * FAC1; io.exz = EXZ * FAC1;
      io.kxx = KXX * FAC1; io.kyy = KYY * FAC1; io.kxy = KXY * FAC1;
      if (g.prop.istrain != 0) {
        const double de[8] = {io.exx, io.eyy, io.exy, io.eyz, io.exz, io.kxx, io.kyy, io.kxy};
        #pragma unroll
        for (int k = 0; k < 8; k++) T.st(SW_STRA + k, T.ld(SW_STRA + k) + de[k]);
      }
      const double dtinv = or_div(DT1, fmax(DT1 * DT1, K_EM20));
      const double thk = T.ld(SW_THK);
      const double eps_k2 = (io.kxx * io.kxx + io.kyy * io.kyy + io.kxx * io.kyy + K_FOURTH * (io.kxy * io.kxy)) * K_ONE_OVER_9 * (thk * thk);
      const double eps_m2 = K_FOUR_OVER_3 * (io.exx * io.exx + io.eyy * io.eyy + io.exx * io.eyy + K_FOURTH * (io.exy * io.exy));
      io.epsd_pg = or_sqrt(eps_k2 + eps_m2) * dtinv;
      T.st(SW_EPSD, K_ONE * io.epsd_pg + (K_ONE - K_ONE) * T.ld(SW_EPSD));
    }
    // ---- CMAIN3
    io.area = AREA; io.thk0 = THK0; io.gs = G * SHF; io.rho = RHO; io.off = OFF; io.sigy = K_EP30;
    if constexpr (FAST >= 1 && STAGED) shell_material_loop_compact<false, FAST>(g, T, DT1, io, wmask);     // three-pass loop (shell_common.cuh)
    else shell_material_loop<LAW, false, STAGED, 0, FAST>(g, T, DT1, io);
    OFF = io.off;
    if (g.bal && P.cs->ipri) shell_bilan<4, STAGED>(P, T, tile, e, RHO, OFF);        // CBILAN (cforc3.F:648; CHVIS3 leaves EINT alone)
    const double SSP = io.ssp;
    const double VISCMX = or_sqrt(K_ONE + io.viscmx * io.viscmx) - io.viscmx;
    // ---- CHVIS3
    double H11, H12, H13, H21, H22, H23, H31, H32, H33, B1r, B2r;
    double STI = K_ZERO, STIR = K_ZERO;
    const double A11N = (LAW != 2) ? g.m36.a11 : g.m2.a11;     // PM(24)
    {
      const double HELAS = K_HALF, HVISC = K_HALF, HVLIN = K_ZERO;     // radioss2.F:641-643
      const double SR2D2 = or_sqrt(K_TWO) * K_HALF;
      double GAMA1, GAMA2, GAMA3, GAMA4;
      const bool plain = (ISMSTR == 1 || ISMSTR == 11 || IHBE < 1);
      if (!plain) {
        const double PX1V = PX1 * VHX, PX2V = PX2 * VHX, PY1V = PY1 * VHY, PY2V = PY2 * VHY;
        GAMA1 = OFF * (K_ONE - PX1V - PY1V); GAMA3 = OFF * (K_ONE + PX1V + PY1V);
        GAMA2 = OFF * (-K_ONE - PX2V - PY2V); GAMA4 = OFF * (-K_ONE + PX2V + PY2V);
      } else { GAMA1 = OFF; GAMA3 = OFF; GAMA2 = -OFF; GAMA4 = -OFF; }
      const double SHFPR3 = or_div(SHF, (K_THREE * (K_ONE + NU)));
      const double HVISH1 = HVISC * H1, HVISH2 = HVISC * H2;
      double R0 = K_FOURTH * RHO; const double R1 = R0 * K_HUNDRED; R0 = R0 * HVLIN;
      const double A1 = R1 * HVISH1;
      const double A2 = R0 * SR2D2 * g.prop.srh1;
      const double SRSHFPR3 = or_sqrt(SHFPR3);
      const double A3 = R1 * HVISH2 * SRSHFPR3;
      const double A4 = R0 * SR2D2 * g.prop.srh2 * SRSHFPR3;
      double HH3 = HELAS * H3;
      const double A5 = HH3 * R1 * K_ZEP072169;
      HH3 = SR2D2 * g.prop.srh3;
      const double A6 = HH3 * R0 * K_ZEP072169;
      R0 = K_FOURTH * YM * HELAS;
      const double A7 = H1 * R0, A8 = H2 * R0 * SHFPR3;
      const double T2A = THK02 * AREA, TSA = or_sqrt(T2A);
      double H1Q = A1 * TSA, H1L = A2 * SSP * TSA, H2Q = A3 * THK02, H2L = A4 * SSP * THK02, H3Q = A5 * T2A, H3L = A6 * SSP * T2A;
      const double TD = THK0 * DT1;
      double HH1 = A7 * TD;
      const double B1 = PX1 * PX1 + PY1 * PY1, B2 = PX2 * PX2 + PY2 * PY2;
      double HH2 = or_div(A8 * THK02 * TD, (B1 + B2));
      if (nc[2] == nc[3]) { H1Q = H1L = H2Q = H2L = H3Q = H3L = HH1 = HH2 = K_ZERO; }
      double HG1, HG2;
      if (plain) { HG1 = (VX[0] - VX[1] + VX[2] - VX[3]) * OFF; HG2 = (VY[0] - VY[1] + VY[2] - VY[3]) * OFF; }
      else { HG1 = VX[0] * GAMA1 + VX[1] * GAMA2 + VX[2] * GAMA3 + VX[3] * GAMA4; HG2 = VY[0] * GAMA1 + VY[1] * GAMA2 + VY[2] * GAMA3 + VY[3] * GAMA4; }
      double hr1 = T.ld(SW_HOURG), hr2 = T.ld(SW_HOURG + 1), hr3 = T.ld(SW_HOURG + 2);
      hr1 = hr1 + HG1 * HH1;
      hr2 = hr2 + HG2 * HH1;
      const double HOUR1A = hr1 + HG1 * (H1L + H1Q * fabs(HG1));
      H11 = HOUR1A * GAMA1; H12 = HOUR1A * GAMA2; H13 = HOUR1A * GAMA3;
      const double HOUR2A = hr2 + HG2 * (H1L + H1Q * fabs(HG2));
      H21 = HOUR2A * GAMA1; H22 = HOUR2A * GAMA2; H23 = HOUR2A * GAMA3;
      if (plain) HG1 = (VZ[0] - VZ[1] + VZ[2] - VZ[3]) * OFF;
      else HG1 = VZ[0] * GAMA1 + VZ[1] * GAMA2 + VZ[2] * GAMA3 + VZ[3] * GAMA4;
      hr3 = hr3 + HG1 * HH2;
      const double HOUR3A = hr3 + HG1 * (H2L + H2Q * fabs(HG1));
      H31 = HOUR3A * GAMA1; H32 = HOUR3A * GAMA2; H33 = HOUR3A * GAMA3;
      HG1 = RX[0] - RX[1] + RX[2] - RX[3];
      HG2 = RY[0] - RY[1] + RY[2] - RY[3];
      const double hr4 = HG1 * (H3L + H3Q * fabs(HG1));
      const double hr5 = HG2 * (H3L + H3Q * fabs(HG2));
      T.st(SW_HOURG, hr1); T.st(SW_HOURG + 1, hr2); T.st(SW_HOURG + 2, hr3);
      T.st(SW_HOURG + 3, hr4); T.st(SW_HOURG + 4, hr5);
      B1r = hr4 * OFF; B2r = hr5 * OFF;               // B11 = B13 = B1r, B12 = B14 = -B1r ; same for B2x
      if (g.nodadt != 0) {                            // nodal stiffnesses of chvis3.F:196-207, 242-253 (STI enters as ZERO, cderi3.F:91)
        const double SCALE = or_div(fmax(fmax(GAMA1 * GAMA1, GAMA2 * GAMA2), fmax(GAMA3 * GAMA3, GAMA4 * GAMA4))
                                    * DT1 * fmax(fmax(HH1 + H1L, HH2 + H2L), H3L), fmax(DT1 * DT1, K_EM20));
        STI = K_ZERO + SCALE;
        if (OFF == K_ZERO) { STI = K_ZERO; STIR = K_ZERO; }
        else {
          const double VV = VISCMX * VISCMX * K_ONE;
          STI = STI + or_div(fmax(B1, B2) * THK0 * A11N, AREA * VV);
          STIR = STI * (THK02 * K_ONE_OVER_12 + AREA * K_ONE_OVER_9);
        }
      }
    }
    // ---- CDT3 (not called with /DT/NODA, cforc3.F:668)
    if (g.nodadt == 0) {
      ALDT = ALDT * VISCMX;                // / sqrt(ALPE), ALPE = 1: exact
      const double DT = or_div(g.dtfac * ALDT, SSP);
      if (OFFG > K_ZERO && OFF != K_ZERO) dt_cand = DT;
      const double DIVM = fmax(ALDT * ALDT, K_EM20);
      STI = or_div(K_HALF * io.vol0 * YM, DIVM);
      STI = K_ZEP81 * STI * OFF;
    }
    // ---- CFINT3
    double Gf[3][4], Gm[2][4];
    {
      const double* FO = io.fo; const double* MO = io.mo;
      const double F1A = FO[0] * THK0, F2A = FO[1] * THK0, F3A = FO[2] * THK0, F4A = FO[3] * THK0, F5A = FO[4] * THK0;
      double M4 = F4A * AREA, M5 = F5A * AREA;
      const double F12 = F1A * PX2 + F3A * PY2, F22 = F2A * PY2 + F3A * PX2, F32 = F5A * PX2 + F4A * PY2;
      const double F11 = F1A * PX1 + F3A * PY1, F21 = F2A * PY1 + F3A * PX1, F31 = F5A * PX1 + F4A * PY1;
      Gf[0][0] = F11 + H11; Gf[0][2] = H13 - F11; Gf[1][0] = F21 + H21; Gf[1][2] = H23 - F21; Gf[2][0] = F31 + H31; Gf[2][2] = H33 - F31;
      Gf[0][1] = F12 + H12; Gf[1][1] = F22 + H22; Gf[2][1] = F32 + H32;
      if (IHBE >= 2 && NPT != 1) { M4 = M4 + (H21 + H23) * Z2; M5 = M5 + (H11 + H13) * Z2; }
      const double M1A = MO[0] * THK02, M2A = MO[1] * THK02, M3A = MO[2] * THK02;
      M4 = M4 * K_FOURTH; M5 = M5 * K_FOURTH;
      const double M11 = -M2A * PY1 - M3A * PX1, M21 = M1A * PX1 + M3A * PY1, M12 = -M2A * PY2 - M3A * PX2, M22 = M1A * PX2 + M3A * PY2;
      Gm[0][0] = M11 - M4 + B1r;  Gm[0][2] = -M11 - M4 + B1r; Gm[0][1] = M12 - M4 + (-B1r); Gm[0][3] = -M12 - M4 + (-B1r);
      Gm[1][0] = M21 + M5 + B2r;  Gm[1][2] = -M21 + M5 + B2r; Gm[1][1] = M22 + M5 + (-B2r); Gm[1][3] = -M22 + M5 + (-B2r);
    }
    // ---- CUPDT3P
    if (OFF < K_ONE) OFFG = OFF;
    T.st(SW_OFF, OFFG);
    const bool dead = OFFG < K_ZERO;
    if (dead) { STI = K_ZERO; STIR = K_ZERO; }
    int sl[4];
    #pragma unroll
    for (int k = 0; k < 4; k++) sl[k] = T.ldi(g.w_slot, k);
    double f4[3] = {K_ZERO, K_ZERO, K_ZERO};
    #pragma unroll
    for (int J = 0; J < 4; J++) {
      double f[3], mm[3];
      #pragma unroll
      for (int I = 0; I < 3; I++) {
        if (J < 3) f[I] = e1[I] * Gf[0][J] + e2[I] * Gf[1][J] + e3[I] * Gf[2][J];
        mm[I] = e1[I] * Gm[0][J] + e2[I] * Gm[1][J];
      }
      if (J == 0) { f4[0] = -f[0]; f4[1] = -f[1]; f4[2] = -f[2]; }
      else if (J < 3) { f4[0] = f4[0] - f[0]; f4[1] = f4[1] - f[1]; f4[2] = f4[2] - f[2]; }
      else { f[0] = f4[0]; f[1] = f4[1]; f[2] = f4[2]; }
      if (dead) { f[0] = f[1] = f[2] = K_ZERO; mm[0] = mm[1] = mm[2] = K_ZERO; }
      double4* row = reinterpret_cast<double4*>(P.fsky + (size_t)8 * sl[J]);
      const double4 r0 = make_double4(-f[0], -f[1], -f[2], -mm[0]), r1 = make_double4(-mm[1], -mm[2], STI, STIR);
      st256(row, r0); st256(row + 1, r1);
    }
    if (g.xs_ftile && g.xs_ftile[tile]) xsend_rows<8, STAGED>(P.nd.xs, T, g.w_slot, 4, P.fsky);   // frontier tile: rows to the neighbours' windows
  }
  cta_epilogue<false, STAGED>(dt_cand, order, P.db, g.blk0 + tile, g_tile, s_tile_dyn, (unsigned)g.nw_rw * ORGPU_TILE * 8u);
}
