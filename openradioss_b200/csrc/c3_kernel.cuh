// c3_kernel.cuh -- fused internal-force kernel for the 3-node C0 shell (ITY=7, Ish3n 1 / 2), one element per
// thread.  One launch does for every element of a super-group what C3FORC3
// (engine/source/elements/sh3n/coque3n/c3forc3.F:296-720; IFRAM_OLD=1, IGTYP=1, no drilling dof) does per group:
//   C3COOR3 (c3coor3.F) gather X,V,VR  ->  C3EVEC3 (c3evec3.F:89-153) frame: e1 along 1-2, AREA = |x31 x x32| / 2
//   C3DERI3 (c3deri3.F) local coordinates, small-strain reference SMSTR(3), PX1 / PY1 / PY2, ALDT
//   C3COEF3 (c3coef3.F), C3DEFO3 (c3defo3.F: rigid-rotation correction), C3CURV3 (c3curv3.F), C3STRA3 (c3stra3.F),
//   epsd_pg (c3forc3.F:538-556), CMAIN3/MULAWC + law (shell_common.cuh: the loop of the 4-node shells),
//   C3DT3 (c3dt3.F: DTFAC1(7), ITYPTST=7; NODADT 0 and 1), C3FINT3 (c3fint3.F), C3FCUM3 / C3MCUM3,
//   C3UPDT3P (c3updt3.F): 3 corner rows into FSKY(8,IADTG)
// then the CTA (dt, user id) arg-min (strict "<", first minimum wins: c3dt3.F).  State tile as for the 4-node shells
// (shell_common.cuh) without hourglass words; SMSTR has 3 words.
#pragma once
#include "shell_common.cuh"

template <int LAW, bool STAGED, int FAST = 0>
__global__ void __launch_bounds__(ORGPU_SHELL_CTA, 3 * ORGPU_PER128)
c3_forces_kernel(const __grid_constant__ ShellParams P)
{
  if (P.cs->abort) return;                               // sticky: a peer-memory wait timed out (exchange.cuh)
  const CtaWork<ShellSG> W = cta_work<ShellSG, false>(P.sg, nullptr, nullptr);
  const ShellSG& g = *W.g;
  const int tile = W.tile;
  const int e = tile * ORGPU_TILE + threadIdx.x;
  __shared__ __align__(8) unsigned long long s_bar;
  double* const g_tile = g.slab + (size_t)tile * g.nw * ORGPU_TILE;
  const double* const g_pf = W.tile_pf >= 0 ? g.slab + (size_t)W.tile_pf * g.nw * ORGPU_TILE : nullptr;   // tile of the CTA one wave ahead
  if (STAGED) tile_load_begin(s_tile_dyn, &s_bar, g_tile, (unsigned)g.nw * ORGPU_TILE * 8u, g_pf);
  const TileAcc<STAGED> T{(STAGED ? s_tile_dyn : g_tile) + threadIdx.x};
  if (!STAGED && threadIdx.x == 0 && g_pf)
    bulk_prefetch_l2(g_pf, (unsigned)g.nw * ORGPU_TILE * 8u);
  double* const sm = g.smstr + (size_t)tile * 3 * ORGPU_TILE + threadIdx.x;      // SMSTR word k at sm[k*128]
#if ORGPU_PREFETCH_NEXT > 0
  if (W.tile_nx >= 0 && threadIdx.x < (3 * ORGPU_TILE * 4) / 128) prefetch_l2(reinterpret_cast<const char*>(g.conn + (size_t)W.tile_nx * 3 * ORGPU_TILE) + 128 * threadIdx.x);
#endif
  double dt_cand = K_EP30; int order = 0x7fffffff;
  const unsigned wmask = (FAST >= 1) ? __ballot_sync(0xffffffffu, e < g.ne) : 0u;     // the warp's lanes that own an element
  if (e < g.ne) {
    const double DT1 = P.cs->dt2;
    const int ISMSTR = g.prop.ismstr, NPT = g.prop.npt, ISH3N = g.prop.ihbe;
    int nc[3];
    { const int* cn = g.conn + (size_t)tile * 3 * ORGPU_TILE + threadIdx.x;
      #pragma unroll
      for (int k = 0; k < 3; k++) nc[k] = __ldg(cn + k * ORGPU_TILE); }
    order = g.order0 + e;
    double xg[3], yg[3], zg[3];
    #pragma unroll
    for (int k = 0; k < 3; k++) { const double4 p = ld256_nc(P.nd.pos + nc[k]); xg[k] = p.x; yg[k] = p.y; zg[k] = p.z; }
    #pragma unroll
    for (int k = 0; k < 3; k++) { prefetch_l1(P.nd.rot + nc[k]); prefetch_l1(P.nd.vel + nc[k]); }
    if (STAGED) mbar_wait(&s_bar, 0);
    double OFFG = T.ld(SW_OFF);
    const bool dead_in = OFFG < K_ZERO;
    double OFF = fmin(K_ONE, fabs(OFFG));
    // ---- C3EVEC3 (IFRAM_OLD = 1) and C3DERI3
    double e1[3], e2[3], e3[3];
    double AREA, X2, X3, Y3;
    {
      const double X21 = xg[1] - xg[0], Y21 = yg[1] - yg[0], Z21 = zg[1] - zg[0];
      const double X31 = xg[2] - xg[0], Y31 = yg[2] - yg[0], Z31 = zg[2] - zg[0];
      const double X32 = xg[2] - xg[1], Y32 = yg[2] - yg[1], Z32 = zg[2] - zg[1];
      e1[0] = X21; e1[1] = Y21; e1[2] = Z21;
      double S = or_sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
      e1[0] = or_div(e1[0], S); e1[1] = or_div(e1[1], S); e1[2] = or_div(e1[2], S);
      e3[0] = Y31 * Z32 - Z31 * Y32; e3[1] = Z31 * X32 - X31 * Z32; e3[2] = X31 * Y32 - Y31 * X32;
      S = or_sqrt(e3[0] * e3[0] + e3[1] * e3[1] + e3[2] * e3[2]);
      e3[0] = or_div(e3[0], S); e3[1] = or_div(e3[1], S); e3[2] = or_div(e3[2], S);
      AREA = K_HALF * S;
      e2[0] = e3[1] * e1[2] - e3[2] * e1[1]; e2[1] = e3[2] * e1[0] - e3[0] * e1[2]; e2[2] = e3[0] * e1[1] - e3[1] * e1[0];
      S = or_sqrt(e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2]);
      e2[0] = or_div(e2[0], S); e2[1] = or_div(e2[1], S); e2[2] = or_div(e2[2], S);
      X2 = e1[0] * X21 + e1[1] * Y21 + e1[2] * Z21;
      X3 = e1[0] * X31 + e1[1] * Y31 + e1[2] * Z31; Y3 = e2[0] * X31 + e2[1] * Y31 + e2[2] * Z31;
    }
    if (ISMSTR == 1 || ISMSTR == 2) {
      if (OFFG == K_TWO) { X2 = sm[0]; X3 = sm[ORGPU_TILE]; Y3 = sm[2 * ORGPU_TILE]; AREA = K_HALF * X2 * Y3; }
      else { __stcs(&sm[0], X2); __stcs(&sm[ORGPU_TILE], X3); __stcs(&sm[2 * ORGPU_TILE], Y3); }
      if (ISMSTR == 1 && OFFG == K_ONE) OFFG = K_TWO;
    }
    Y3 = copysign(fmax(K_EM15, fabs(Y3)), Y3);
    const double PX1 = -K_HALF * Y3, PY1 = K_HALF * (X3 - X2), PY2 = -K_HALF * X3;
    double ALDT;
    {
      const double AL1 = X2 * X2, AL2 = (X3 - X2) * (X3 - X2) + Y3 * Y3, AL3 = X3 * X3 + Y3 * Y3;
      const double ALMAX = fmax(fmax(AL1, AL2), AL3);
      ALDT = or_div(K_TWO * AREA, or_sqrt(ALMAX));
    }
    // ---- C3COEF3
    const double THK0 = (g.prop.ithk > 0) ? T.ld(SW_THK) : T.ld(g.w_thke);
    const double THK02 = THK0 * THK0;
    double RHO, NU, G, A11;
    MatIO io;
    if (LAW != 2) { const orgpu_law36& m = g.m36; RHO = m.rho0; NU = m.nu; G = m.shear; A11 = m.a11; io.ssp = m.ssp; }
    else           { const orgpu_law2& m = g.m2;   RHO = m.rho0; NU = m.nu; G = m.shear; A11 = m.a11; io.ssp = m.ssp; }
    double SHF = K_ZERO;
    if (NPT != 1) { const double FAC1 = K_TWO * (K_ONE + NU) * THK02; const int ISH = 0; const double FSH = g.prop.shf;
                    SHF = FSH * (K_ONE - ISH + or_div(ISH * FAC1, (FSH * AREA + FAC1))); }
    // ---- C3DEFO3
    double EXX, EYY, EXY, EYZ, EZX;
    {
      double vl[3][3];
      #pragma unroll
      for (int k = 0; k < 3; k++) {
        double4 v = ld256_nc(P.nd.vel + nc[k]);
        if (dead_in) { v.x = K_ZERO; v.y = K_ZERO; v.z = K_ZERO; }
        vl[k][0] = v.x; vl[k][1] = v.y; vl[k][2] = v.z;
      }
      double VX1 = e1[0] * vl[0][0] + e1[1] * vl[0][1] + e1[2] * vl[0][2];
      double VX2 = e1[0] * vl[1][0] + e1[1] * vl[1][1] + e1[2] * vl[1][2];
      double VX3 = e1[0] * vl[2][0] + e1[1] * vl[2][1] + e1[2] * vl[2][2];
      double VY3 = e2[0] * vl[2][0] + e2[1] * vl[2][1] + e2[2] * vl[2][2];
      double VY2 = e2[0] * vl[1][0] + e2[1] * vl[1][1] + e2[2] * vl[1][2];
      double VY1 = e2[0] * vl[0][0] + e2[1] * vl[0][1] + e2[2] * vl[0][2];
      const double VZ1 = e3[0] * vl[0][0] + e3[1] * vl[0][1] + e3[2] * vl[0][2];
      const double VZ2 = e3[0] * vl[1][0] + e3[1] * vl[1][1] + e3[2] * vl[1][2];
      const double VZ3 = e3[0] * vl[2][0] + e3[1] * vl[2][1] + e3[2] * vl[2][2];
      const double DT1V4 = K_FOURTH * DT1;
      const double DT1V4B = (ISH3N < 2) ? K_ZERO : DT1V4;
      const double VZ12 = VZ1 - VZ2, VZ13 = VZ1 - VZ3, VZ23 = VZ2 - VZ3;
      const double TMP1 = or_div(DT1V4 * VZ12, (PY1 + PY2));
      double TMP2 = or_div((PY1 * VZ1 + PY2 * VZ2), (PY1 + PY2));
      TMP2 = or_div(DT1V4 * (TMP2 - VZ3), PX1);
      double VY12 = VY1 - VY2;
      const double TMP11 = or_div(DT1V4B * VY12, (PY1 + PY2));
      double TMP22 = or_div((PY1 * VX1 + PY2 * VX2), (PY1 + PY2));
      TMP22 = or_div(DT1V4B * (TMP22 - VX3), PX1);
      const double VX10 = VX1, VX20 = VX2, VX30 = VX3;
      VX1 = VX1 - VZ1 * TMP1 - VY1 * TMP11;
      VX2 = VX2 - VZ2 * TMP1 - VY2 * TMP11;
      VX3 = VX3 - VZ3 * TMP1 - VY3 * TMP11;
      VY1 = VY1 - VZ1 * TMP2 - VX10 * TMP22;
      VY2 = VY2 - VZ2 * TMP2 - VX20 * TMP22;
      VY3 = VY3 - VZ3 * TMP2 - VX30 * TMP22;
      const double VX12 = VX1 - VX2; VY12 = VY1 - VY2;
      const double VX13 = VX1 - VX3, VY13 = VY1 - VY3, VX23 = VX2 - VX3, VY23 = VY2 - VY3;
      EXX = PX1 * VX12;
      EYY = PY1 * VY13 + PY2 * VY23;
      EXY = PY1 * VX13 + PY2 * VX23 + PX1 * VY12;
      EYZ = PY1 * VZ13 + PY2 * VZ23;
      EZX = PX1 * VZ12;
    }
    // ---- C3CURV3
    double KXX, KYY, KXY;
    {
      double RX[3], RY[3];
      #pragma unroll
      for (int k = 0; k < 3; k++) {
        double4 w = ld256_nc(P.nd.rot + nc[k]);
        if (dead_in) { w.x = K_ZERO; w.y = K_ZERO; w.z = K_ZERO; }
        RX[k] = e1[0] * w.x + e1[1] * w.y + e1[2] * w.z;
        RY[k] = e2[0] * w.x + e2[1] * w.y + e2[2] * w.z;
      }
      const double RX12T = RX[0] - RX[1], RX13T = RX[0] - RX[2], RX23T = RX[1] - RX[2];
      KYY = -PY1 * RX13T - PY2 * RX23T;
      KXY = PX1 * RX12T;
      const double RY12T = RY[0] - RY[1], RY13T = RY[0] - RY[2], RY23T = RY[1] - RY[2];
      KXX = PX1 * RY12T;
      KXY = PY1 * RY13T + PY2 * RY23T - KXY;
      const double RYAVT = PX1 * (PX1 * (-RX[0] + RX[1])
                                  + (K_TWO * PY1 + K_THREE * PY2) * RY[0]
                                  + (K_THREE * PY1 + K_TWO * PY2) * RY[1]
                                  + (PY1 + PY2) * RY[2]);
      const double RXAVT = -PX1 * (+(K_TWO * PY1 + PY2) * RX[0]
                                   + (PY1 + K_TWO * PY2) * RX[1]
                                   + K_THREE * (PY1 + PY2) * RX[2])
                         + PY1 * (PY1 + K_TWO * PY2) * RY[0]
                         - PY2 * (K_TWO * PY1 + PY2) * RY[1]
                         + (PY2 * PY2 - PY1 * PY1) * RY[2];
      EZX = EZX + RYAVT * K_THIRD;
      EYZ = EYZ + RXAVT * K_THIRD;
    }
    // ---- C3STRA3 + element strain rate
    {
      const double FAC1 = or_div(DT1, AREA);
      io.exx = EXX * FAC1; io.eyy = EYY * FAC1; io.exy = EXY * FAC1; io.eyz = EYZ * FAC1; io.exz = EZX * FAC1;
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
    if (g.bal && P.cs->ipri) shell_bilan<3, STAGED>(P, T, tile, e, RHO, OFF);        // C3BILAN (c3forc3.F:616)
    const double SSP = io.ssp;
    // ---- C3DT3 (IGTYP=1, ZOFFSET=0, IDTMIN(7)=0)
    double STI, STIR;
    {
      const double VISCMX = or_sqrt(K_ONE + io.viscmx * io.viscmx) - io.viscmx;
      ALDT = ALDT * VISCMX;                // / sqrt(ALPE), ALPE = 1: exact
      const double F_OSET = K_ONE + or_div(K_HALF * fabs(K_ZERO), THK0);
      if (g.nodadt != 0) {
        if (OFF == K_ZERO) { STI = K_ZERO; STIR = K_ZERO; }
        else {
          const double ATHK = AREA * THK0;
          STI = or_div(ATHK * F_OSET * A11, (ALDT * ALDT));
          STIR = STI * (THK0 * THK0 * K_ONE_OVER_12 + or_div(K_HALF * SHF * AREA * G, A11)) + STI * K_ZERO * K_ZERO;
        }
      } else {
        const double F_DTE = or_div(K_ONE, or_sqrt(F_OSET));
        const double DT = or_div(g.dtfac * F_DTE * ALDT, SSP);
        if (OFFG > K_ZERO && OFF != K_ZERO) dt_cand = DT;
        STI = or_div(AREA * THK0 * F_OSET * A11, (ALDT * ALDT));
        STI = K_ZEP81 * K_ZEP81 * STI * OFF;
        STIR = K_ZERO;
      }
    }
    // ---- C3FINT3
    double FX[3], FY[3], FZ[3], MX[3], MY[3];
    {
      const double* FO = io.fo; const double* MO = io.mo;
      const double F1 = FO[0] * THK0, F3 = FO[2] * THK0;
      FX[0] = F1 * PX1 + F3 * PY1;
      FX[1] = -F1 * PX1 + F3 * PY2;
      FX[2] = -FX[0] - FX[1];
      const double F2 = FO[1] * THK0;
      FY[0] = F2 * PY1 + F3 * PX1;
      FY[1] = F2 * PY2 - F3 * PX1;
      FY[2] = -FY[0] - FY[1];
      const double F4 = FO[3] * THK0, F5 = FO[4] * THK0;
      FZ[0] = F5 * PX1 + F4 * PY1;
      FZ[1] = -F5 * PX1 + F4 * PY2;
      FZ[2] = -FZ[0] - FZ[1];
      const double TH2 = THK0 * THK0;
      const double M2 = MO[1] * TH2, M3 = MO[2] * TH2;
      MX[0] = -M2 * PY1 - M3 * PX1;
      MX[1] = -M2 * PY2 + M3 * PX1;
      MX[2] = -MX[0] - MX[1];
      const double M1 = MO[0] * TH2;
      MY[0] = M1 * PX1 + M3 * PY1;
      MY[1] = -M1 * PX1 + M3 * PY2;
      MY[2] = -MY[0] - MY[1];
      double M4 = F4 * K_THIRD, M5 = F5 * K_THIRD;
      M5 = M5 * PX1;
      MY[0] = MY[0] + M5 * (K_TWO * PY1 + K_THREE * PY2) + M4 * PY1 * (PY1 + K_TWO * PY2);
      MY[1] = MY[1] + M5 * (K_THREE * PY1 + K_TWO * PY2) - M4 * PY2 * (K_TWO * PY1 + PY2);
      MY[2] = MY[2] + M5 * (PY1 + PY2) + M4 * (PY2 * PY2 - PY1 * PY1);
      M5 = M5 * PX1;
      M4 = M4 * PX1;
      MX[0] = MX[0] - M5 - M4 * (K_TWO * PY1 + PY2);
      MX[1] = MX[1] + M5 - M4 * (PY1 + K_TWO * PY2);
      MX[2] = MX[2] - M4 * K_THREE * (PY1 + PY2);
    }
    // ---- C3UPDT3P (C3FCUM3 / C3MCUM3 folded into the row loop)
    if (OFF < K_ONE) OFFG = OFF;
    T.st(SW_OFF, OFFG);
    const bool dead = OFFG < K_ZERO;
    if (dead) { STI = K_ZERO; STIR = K_ZERO; }
    int sl[3];
    #pragma unroll
    for (int k = 0; k < 3; k++) sl[k] = T.ldi(g.w_slot, k);
    #pragma unroll
    for (int J = 0; J < 3; J++) {
      double f[3], mm[3];
      #pragma unroll
      for (int I = 0; I < 3; I++) {
        f[I] = e1[I] * FX[J] + e2[I] * FY[J] + e3[I] * FZ[J];
        mm[I] = e1[I] * MX[J] + e2[I] * MY[J];
      }
      if (dead) { f[0] = f[1] = f[2] = K_ZERO; mm[0] = mm[1] = mm[2] = K_ZERO; }
      double4* row = reinterpret_cast<double4*>(P.fsky + (size_t)8 * sl[J]);
      const double4 r0 = make_double4(-f[0], -f[1], -f[2], -mm[0]), r1 = make_double4(-mm[1], -mm[2], STI, STIR);
      st256(row, r0); st256(row + 1, r1);
    }
    if (g.xs_ftile && g.xs_ftile[tile]) xsend_rows<8, STAGED>(P.nd.xs, T, g.w_slot, 3, P.fsky);   // frontier tile: rows to the neighbours' windows
  }
  cta_epilogue<false, STAGED>(dt_cand, order, P.db, g.blk0 + tile, g_tile, s_tile_dyn, (unsigned)g.nw_rw * ORGPU_TILE * 8u);
}
