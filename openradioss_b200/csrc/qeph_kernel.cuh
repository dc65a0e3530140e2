// qeph_kernel.cuh -- fused internal-force kernel for QEPH 4-node shells (Ishell=24), one element
// per thread.  One launch does for every element of a super-group what CZFORC3
// (engine/source/elements/shell/coquez/czforc3.F:380-823; ISROT=0, IORTH=0) does per group:
//   CZCORC1  (czcorc.F:143-611) + CLSKEW3 (sh3n/coquedk/cdkcoor3.F:336-397, IREP=0)
//            gather X,V,VR (12 x 32-byte nodal records), local frame, small-strain reference,
//            characteristic length, 2nd-order rigid-rotation correction
//   CZCORP5  (czcorp5.F:83-346)   warped-element projection of the nodal velocities
//   CNCOEF3B (sh3n/coquedk/cncoef3.F:77-284), CZDEF (czdef.F:115-200), CZSTRA3 (czstra3.F:77-119)
//   epsd_pg  (czforc3.F:582-591), CMAIN3/MULAWC/SIGEPS36C|SIGEPS02C (shell_common.cuh)
//   CNDT3    (sh3n/coquedk/cndt3.F:85-318), CZFINTCE (czfintce.F:63-104),
//   CZFINTN1 (czfintn.F:86-453, physical hourglass), CZPROJN (czproj.F:1206-1467),
//   CUPDTN3P (coque/cupdtn3.F:545-699): 4 corner rows of 8 doubles into the FSKY slots of IADC
// then the CTA (dt, user id) arg-min (first minimum wins: strict "<" in cndt3.F:299-309).
#pragma once
#include "shell_common.cuh"

// The element working set (local frame, centred coordinates, warped-element projection matrices ...) is
// larger than a thread's register budget at three CTAs per SM, and everything the compiler parks in
// local memory competes with the staged state tile for the SM's L1/shared storage.  So only a minimal
// set crosses the through-thickness loop (local corner coordinates, Z1, area, frame, hourglass rates);
// the derived quantities are RE-DERIVED after the loop by the same expressions -- bit-identical, ~350
// fp64 instructions -- behind an optimisation barrier that stops the compiler from keeping them alive.
#define ORGPU_OPAQUE(x) asm volatile("" : "+d"(x))

struct QephGeo {                 // centred corner coordinates and B-matrix ingredients (czcorc.F:336-375)
  double CX[4], CY[4], X13, X24, Y13, Y24, MX13, MX23, MX34, MY13, MY23, MY34, L13, L24;
};
__device__ __forceinline__ void qeph_geo(double XL2, double YL2, double XL3, double YL3, double XL4, double YL4, QephGeo& q)
{
  const double LX = K_FOURTH * (XL2 + XL3 + XL4), LY = K_FOURTH * (YL2 + YL3 + YL4);
  q.CX[0] = -LX; q.CX[1] = XL2 - LX; q.CX[2] = XL3 - LX; q.CX[3] = XL4 - LX;
  q.CY[0] = -LY; q.CY[1] = YL2 - LY; q.CY[2] = YL3 - LY; q.CY[3] = YL4 - LY;
  q.X13 = (q.CX[0] - q.CX[2]) * K_HALF; q.X24 = (q.CX[1] - q.CX[3]) * K_HALF;
  q.Y13 = (q.CY[0] - q.CY[2]) * K_HALF; q.Y24 = (q.CY[1] - q.CY[3]) * K_HALF;
  q.MX13 = (q.CX[0] + q.CX[2]) * K_HALF; q.MX23 = (q.CX[1] + q.CX[2]) * K_HALF; q.MX34 = (q.CX[2] + q.CX[3]) * K_HALF;
  q.MY13 = (q.CY[0] + q.CY[2]) * K_HALF; q.MY23 = (q.CY[1] + q.CY[2]) * K_HALF; q.MY34 = (q.CY[2] + q.CY[3]) * K_HALF;
  q.L13 = q.X13 * q.X13 + q.Y13 * q.Y13; q.L24 = q.X24 * q.X24 + q.Y24 * q.Y24;
}

// CZCORP5 (czcorp5.F:83-250), geometry-only part for a warped element: nodal normals VQN, the inverse DI of
// the rigid-mode Gram matrix and DB = DI * VQN
// (kept inline twice: as ONE out-of-line copy its 30 results travel through the stack -- 0.392 -> 0.547 ms on C2)
__device__ __forceinline__ void qeph_warp_proj(const QephGeo& q, double Z1, double AREA, double VQN[3][4], double DI[6], double DB[3][4])
{
  const double Z2 = Z1 * Z1;
  const double A_4 = AREA * K_FOURTH;
  double SZ1 = q.MX13 * q.Y24 - q.MY13 * q.X24;
  double SZ2 = A_4 + SZ1;
  double SZ = Z2 * q.L24;
  double SL = or_div(K_ONE, or_sqrt(SZ + SZ2 * SZ2));
  VQN[0][0] = -Z1 * q.Y24; VQN[1][0] = Z1 * q.X24; VQN[2][0] = SZ2 * SL;
  VQN[0][2] = -VQN[0][0]; VQN[1][2] = -VQN[1][0];
  VQN[0][0] = VQN[0][0] * SL; VQN[1][0] = VQN[1][0] * SL;
  SZ2 = A_4 - SZ1;
  SL = or_div(K_ONE, or_sqrt(SZ + SZ2 * SZ2));
  VQN[0][2] = VQN[0][2] * SL; VQN[1][2] = VQN[1][2] * SL; VQN[2][2] = SZ2 * SL;
  SZ1 = q.MX13 * q.Y13 - q.MY13 * q.X13;
  SZ2 = A_4 + SZ1;
  SZ = Z2 * q.L13;
  SL = or_div(K_ONE, or_sqrt(SZ + SZ2 * SZ2));
  VQN[0][1] = -Z1 * q.Y13; VQN[1][1] = Z1 * q.X13; VQN[2][1] = SZ2 * SL;
  VQN[0][3] = -VQN[0][1]; VQN[1][3] = -VQN[1][1];
  VQN[0][1] = VQN[0][1] * SL; VQN[1][1] = VQN[1][1] * SL;
  SZ2 = A_4 - SZ1;
  SL = or_div(K_ONE, or_sqrt(SZ + SZ2 * SZ2));
  VQN[0][3] = VQN[0][3] * SL; VQN[1][3] = VQN[1][3] * SL; VQN[2][3] = SZ2 * SL;
  const double* CX = q.CX; const double* CY = q.CY;
  const double XX = CX[0] * CX[0] + CX[1] * CX[1] + CX[2] * CX[2] + CX[3] * CX[3];
  const double YY = CY[0] * CY[0] + CY[1] * CY[1] + CY[2] * CY[2] + CY[3] * CY[3];
  const double XY = CX[0] * CY[0] + CX[1] * CY[1] + CX[2] * CY[2] + CX[3] * CY[3];
  const double XZ = (CX[0] - CX[1] + CX[2] - CX[3]) * Z1;
  const double YZ = (CY[0] - CY[1] + CY[2] - CY[3]) * Z1;
  const double ZZ = K_FOUR * Z2;
  double D[6];
  D[0] = YY + ZZ + K_FOUR - (VQN[0][0] * VQN[0][0] + VQN[0][1] * VQN[0][1] + VQN[0][2] * VQN[0][2] + VQN[0][3] * VQN[0][3]);
  D[1] = XX + ZZ + K_FOUR - (VQN[1][0] * VQN[1][0] + VQN[1][1] * VQN[1][1] + VQN[1][2] * VQN[1][2] + VQN[1][3] * VQN[1][3]);
  D[2] = XX + YY + K_FOUR - (VQN[2][0] * VQN[2][0] + VQN[2][1] * VQN[2][1] + VQN[2][2] * VQN[2][2] + VQN[2][3] * VQN[2][3]);
  D[3] = -XY - (VQN[0][0] * VQN[1][0] + VQN[0][1] * VQN[1][1] + VQN[0][2] * VQN[1][2] + VQN[0][3] * VQN[1][3]);
  D[4] = -XZ - (VQN[0][0] * VQN[2][0] + VQN[0][1] * VQN[2][1] + VQN[0][2] * VQN[2][2] + VQN[0][3] * VQN[2][3]);
  D[5] = -YZ - (VQN[1][0] * VQN[2][0] + VQN[1][1] * VQN[2][1] + VQN[1][2] * VQN[2][2] + VQN[1][3] * VQN[2][3]);
  const double ABC = D[0] * D[1] * D[2];
  const double XXYZ2 = D[0] * D[5] * D[5], YYXZ2 = D[1] * D[4] * D[4], ZZXY2 = D[2] * D[3] * D[3];
  double DETA = fabs(ABC + K_TWO * D[3] * D[4] * D[5] - XXYZ2 - YYXZ2 - ZZXY2);
  DETA = or_div(K_ONE, fmax(DETA, K_EM20));
  DI[0] = or_div((ABC - XXYZ2) * DETA, fmax(D[0], K_EM20));
  DI[1] = or_div((ABC - YYXZ2) * DETA, fmax(D[1], K_EM20));
  DI[2] = or_div((ABC - ZZXY2) * DETA, fmax(D[2], K_EM20));
  DI[3] = (D[4] * D[5] - D[3] * D[2]) * DETA;
  DI[4] = (D[5] * D[3] - D[4] * D[1]) * DETA;
  DI[5] = (D[3] * D[4] - D[5] * D[0]) * DETA;
  #pragma unroll
  for (int J = 0; J < 4; J++) {
    DB[0][J] = DI[0] * VQN[0][J] + DI[3] * VQN[1][J] + DI[4] * VQN[2][J];
    DB[1][J] = DI[3] * VQN[0][J] + DI[1] * VQN[1][J] + DI[5] * VQN[2][J];
    DB[2][J] = DI[4] * VQN[0][J] + DI[5] * VQN[1][J] + DI[2] * VQN[2][J];
  }
}

#ifndef ORGPU_SHELL_MINB
#define ORGPU_SHELL_MINB 3
#endif

template <int LAW, bool STAGED, int FAST = 0, bool TAB = false>
__global__ void __launch_bounds__(ORGPU_SHELL_CTA, ORGPU_SHELL_MINB * ORGPU_PER128)
qeph_forces_kernel(const __grid_constant__ ShellParams P)
{
#ifndef ORGPU_NO_ABORT
  if (P.cs->abort) return;                               // sticky: a peer-memory wait timed out (exchange.cuh)
#endif
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
  const bool full_tile = (tile + 1) * ORGPU_TILE <= g.ne;    // all 128 threads own an element: the phase barriers are safe (common.cuh)
  if (e < g.ne) {
    const double DT1 = P.cs->dt2;
    const int ISMSTR = g.prop.ismstr, NPT = g.prop.npt;
    int nc[4];
    { const int* cn = g.conn + (size_t)tile * 4 * ORGPU_TILE + threadIdx.x;
      #pragma unroll
      for (int k = 0; k < 4; k++) nc[k] = __ldg(cn + k * ORGPU_TILE); }
    order = g.order0 + e;
    double px[4], py[4], pz[4];
    #pragma unroll
    for (int k = 0; k < 4; k++) { const double4 p = ld256_nc(P.nd.pos + nc[k]); px[k] = p.x; py[k] = p.y; pz[k] = p.z; }
    #pragma unroll
    for (int k = 0; k < 4; k++) { prefetch_l1(P.nd.rot + nc[k]); prefetch_l1(P.nd.vel + nc[k]); }
    if (STAGED) mbar_wait(&s_bar, 0);                     // the state tile has landed (issued before the gather)
    double OFFG = T.ld(SW_OFF);
    // ---- local frame (CLSKEW3, IREP=0)
    double VQ[3][3];                                   // VQ[a][b] = R_ab : columns are e1, e2, e3
    double AREA, AREA_I;
    {
      const double RX = px[1] + px[2] - px[0] - px[3], SX = px[2] + px[3] - px[0] - px[1];
      const double RY = py[1] + py[2] - py[0] - py[3], SY = py[2] + py[3] - py[0] - py[1];
      const double RZ = pz[1] + pz[2] - pz[0] - pz[3], SZ = pz[2] + pz[3] - pz[0] - pz[1];
      double E3X = RY * SZ - RZ * SY, E3Y = RZ * SX - RX * SZ, E3Z = RX * SY - RY * SX;
      double DET = or_sqrt(E3X * E3X + E3Y * E3Y + E3Z * E3Z);
      if (DET < K_EM20 && OFFG != K_ZERO) OFFG = K_ZERO;
      const double OFF_LOC = (fabs(OFFG) != K_ZERO) ? K_ONE : K_ZERO;
      DET = fmax(K_EM20, DET);
      const double CC = fmax(or_div(OFF_LOC, DET), K_EM20);
      E3X = E3X * CC; E3Y = E3Y * CC; E3Z = E3Z * CC;
      const double C1C1 = RX * RX + RY * RY + RZ * RZ, C2C2 = SX * SX + SY * SY + SZ * SZ;
      double C2_1 = K_ZERO, C1_1 = K_ZERO;
      if (C1C1 != K_ZERO) { C2_1 = or_sqrt(or_div(C2C2, fmax(K_EM20, C1C1))); C1_1 = K_ONE; }
      else if (C2C2 != K_ZERO) { C2_1 = K_ONE; C1_1 = or_sqrt(or_div(C1C1, fmax(K_EM20, C2C2))); }
      double E1X = RX * C2_1 + (SY * E3Z - SZ * E3Y) * C1_1;
      double E1Y = RY * C2_1 + (SZ * E3X - SX * E3Z) * C1_1;
      double E1Z = RZ * C2_1 + (SX * E3Y - SY * E3X) * C1_1;
      double C1 = or_sqrt(E1X * E1X + E1Y * E1Y + E1Z * E1Z);
      if (C1 != K_ZERO) C1 = or_div(K_ONE, fmax(K_EM20, C1));
      E1X = E1X * C1; E1Y = E1Y * C1; E1Z = E1Z * C1;
      VQ[0][0] = E1X; VQ[1][0] = E1Y; VQ[2][0] = E1Z;
      VQ[0][1] = E3Y * E1Z - E3Z * E1Y; VQ[1][1] = E3Z * E1X - E3X * E1Z; VQ[2][1] = E3X * E1Y - E3Y * E1X;
      VQ[0][2] = E3X; VQ[1][2] = E3Y; VQ[2][2] = E3Z;
      AREA = K_FOURTH * DET;
      AREA_I = fmax(or_div(OFF_LOC, AREA), K_EM20);
    }
    // ---- local coordinates relative to node 1 (czcorc.F:195-229)
    double XL2, YL2, XL3, YL3, XL4, YL4, Z1;
    {
      const double L0x = K_FOURTH * (px[2] + px[3] + px[0] + px[1]);
      const double L0y = K_FOURTH * (py[2] + py[3] + py[0] + py[1]);
      const double L0z = K_FOURTH * (pz[2] + pz[3] + pz[0] + pz[1]);
      double XX = px[1] - px[0], YY = py[1] - py[0], ZZ = pz[1] - pz[0];
      XL2 = VQ[0][0] * XX + VQ[1][0] * YY + VQ[2][0] * ZZ; YL2 = VQ[0][1] * XX + VQ[1][1] * YY + VQ[2][1] * ZZ;
      XX = px[0] - L0x; YY = py[0] - L0y; ZZ = pz[0] - L0z;
      Z1 = VQ[0][2] * XX + VQ[1][2] * YY + VQ[2][2] * ZZ;
      XX = px[2] - px[0]; YY = py[2] - py[0]; ZZ = pz[2] - pz[0];
      XL3 = VQ[0][0] * XX + VQ[1][0] * YY + VQ[2][0] * ZZ; YL3 = VQ[0][1] * XX + VQ[1][1] * YY + VQ[2][1] * ZZ;
      XX = px[3] - px[0]; YY = py[3] - py[0]; ZZ = pz[3] - pz[0];
      XL4 = VQ[0][0] * XX + VQ[1][0] * YY + VQ[2][0] * ZZ; YL4 = VQ[0][1] * XX + VQ[1][1] * YY + VQ[2][1] * ZZ;
    }
    // ---- small-strain reference (czcorc.F:297-324)
    if (ISMSTR == 1 || ISMSTR == 2) {
      if (fabs(OFFG) == K_TWO) {
        XL2 = sm[0]; YL2 = sm[ORGPU_TILE]; XL3 = sm[2 * ORGPU_TILE];
        YL3 = sm[3 * ORGPU_TILE]; XL4 = sm[4 * ORGPU_TILE]; YL4 = sm[5 * ORGPU_TILE];
        Z1 = K_ZERO;
        AREA = K_HALF * ((XL2 - XL4) * YL3 - XL3 * (YL2 - YL4));
        AREA_I = or_div(K_ONE, fmax(K_EM20, AREA));
      } else {
        __stcs(&sm[0], XL2); __stcs(&sm[ORGPU_TILE], YL2); __stcs(&sm[2 * ORGPU_TILE], XL3);
        __stcs(&sm[3 * ORGPU_TILE], YL3); __stcs(&sm[4 * ORGPU_TILE], XL4); __stcs(&sm[5 * ORGPU_TILE], YL4);
      }
    }
    if (ISMSTR == 1 && OFFG == K_ONE) OFFG = K_TWO;
    // ---- centred corner coordinates and the B-matrix ingredients (czcorc.F:336-375)
    MatIO io;
    double VHG[6], OFF, THK0, LL, FACN1, FACN2;
    bool PLAT;
    {   // ======== everything in this scope is re-derived after the through-thickness loop ========
    QephGeo q0; qeph_geo(XL2, YL2, XL3, YL3, XL4, YL4, q0);
    const double* CX = q0.CX; const double* CY = q0.CY;
    const double X13 = q0.X13, X24 = q0.X24, Y13 = q0.Y13, Y24 = q0.Y24;
    const double MX13 = q0.MX13, MX23 = q0.MX23, MX34 = q0.MX34, MY13 = q0.MY13, MY23 = q0.MY23, MY34 = q0.MY34;
    const double L13 = q0.L13, L24 = q0.L24;
    // ---- characteristic length (czcorc.F:380-404)
    double LM;
    {
      const double c1 = CX[1] * CY[3] - CY[1] * CX[3];
      const double c2 = CX[0] * CY[2] - CY[0] * CX[2];
      const double HS = fmax(fabs(c1), fabs(c2)) * AREA_I;
      const double rx = XL2 + XL3 - XL4, ry = YL2 + YL3 - YL4, sx = -XL2 + XL3 + XL4, sy = -YL2 + YL3 + YL4;
      const double C1 = or_sqrt(rx * rx + ry * ry), C2 = or_sqrt(sx * sx + sy * sy);
      double S1 = K_FOURTH * (or_div(fmax(C1, C2), fmin(C1, C2)) - K_ONE);
      const double f1 = fmin(K_HALF, S1) + K_ONE;
      double f2 = or_div(K_FOUR * AREA, (C1 * C2));
      f2 = (double)3.413f * fmax(K_ZERO, f2 - (double)0.7071f);
      f2 = (double)0.78f + (double)0.22f * f2 * f2 * f2;
      const double FACI = K_TWO * f1 * f2;
      LL = fmax(L13, L24);
      LM = K_HALF * (L13 + L24);
      FACN1 = or_sqrt(or_div(L24, LL)); FACN2 = or_sqrt(or_div(L13, LL));
      S1 = or_sqrt(FACI * (K_FIVE_OVER_4 + HS) * LL);
      S1 = fmax(S1, K_EM10);
      LL = or_div(AREA, S1);
    }
    // ---- nodal velocities: translations to V13/V24/VHI, rotations to the local frame
    double RL[3][4];                                     // RL[2][k] = e3 component (used only when warped)
    #pragma unroll
    for (int k = 0; k < 4; k++) {
      const double4 w = ld256_nc(P.nd.rot + nc[k]);
      RL[0][k] = VQ[0][0] * w.x + VQ[1][0] * w.y + VQ[2][0] * w.z;
      RL[1][k] = VQ[0][1] * w.x + VQ[1][1] * w.y + VQ[2][1] * w.z;
      RL[2][k] = VQ[0][2] * w.x + VQ[1][2] * w.y + VQ[2][2] * w.z;
    }
    double V13[3], V24[3], VHI[3];
    {
      double vx[4], vy[4], vz[4];
      #pragma unroll
      for (int k = 0; k < 4; k++) { const double4 v = ld256_nc(P.nd.vel + nc[k]); vx[k] = v.x; vy[k] = v.y; vz[k] = v.z; }
      const double G13x = vx[0] - vx[2], G24x = vx[1] - vx[3], GHx = vx[0] - vx[1] + vx[2] - vx[3];
      const double G13y = vy[0] - vy[2], G24y = vy[1] - vy[3], GHy = vy[0] - vy[1] + vy[2] - vy[3];
      const double G13z = vz[0] - vz[2], G24z = vz[1] - vz[3], GHz = vz[0] - vz[1] + vz[2] - vz[3];
      #pragma unroll
      for (int c = 0; c < 3; c++) {
        V13[c] = (VQ[0][c] * G13x + VQ[1][c] * G13y + VQ[2][c] * G13z);
        V24[c] = (VQ[0][c] * G24x + VQ[1][c] * G24y + VQ[2][c] * G24z);
        VHI[c] = (VQ[0][c] * GHx + VQ[1][c] * GHy + VQ[2][c] * GHz);
      }
    }
    // ---- 2nd-order rigid-rotation correction (czcorc.F:541-569)
    {
      const double DT05 = K_HALF * DT1, DT025 = K_FOURTH * DT1;
      const double EXZ = Y24 * V13[2] - Y13 * V24[2];
      const double EYZ = -X24 * V13[2] + X13 * V24[2];
      const double DDRY = DT05 * EXZ * AREA_I, DDRX = DT05 * EYZ * AREA_I;
      const double V13X = V13[0], V24X = V24[0], VHIX = VHI[0];
      const double DDRZ1 = (fabs(X13 - X24) < K_EM10) ? K_ZERO : or_div(DT025 * (V13[1] - V24[1]), (X13 - X24));
      V13[0] = V13[0] - DDRY * V13[2] - DDRZ1 * V13[1];
      V24[0] = V24[0] - DDRY * V24[2] - DDRZ1 * V24[1];
      VHI[0] = VHI[0] - DDRY * VHI[2] - DDRZ1 * VHI[1];
      const double DDRZ2 = (fabs(Y13 + Y24) < K_EM10) ? K_ZERO : or_div(DT025 * (V13X + V24X), (Y13 + Y24));
      V13[1] = V13[1] - DDRX * V13[2] - DDRZ2 * V13X;
      V24[1] = V24[1] - DDRX * V24[2] - DDRZ2 * V24X;
      VHI[1] = VHI[1] - DDRX * VHI[2] - DDRZ2 * VHIX;
    }
    // ---- CZCORP5: flat test, nodal normals and full projection for warped elements
    {
      const double Z2 = Z1 * Z1;
      if (Z2 < LM * K_EM8 || NPT == 1) { Z1 = K_ZERO; PLAT = true; }
      else {
        PLAT = false;
        double VQN[3][4], DI[6], DB[3][4];
        qeph_warp_proj(q0, Z1, AREA, VQN, DI, DB);
        double AR[3], AD[4];
        AR[0] = -Z1 * VHI[1] + Y13 * V13[2] + Y24 * V24[2] + MY13 * VHI[2] + RL[0][0] + RL[0][1] + RL[0][2] + RL[0][3];
        AR[1] = Z1 * VHI[0] - X13 * V13[2] - X24 * V24[2] - MX13 * VHI[2] + RL[1][0] + RL[1][1] + RL[1][2] + RL[1][3];
        AR[2] = X13 * V13[1] + X24 * V24[1] + MX13 * VHI[1] - Y13 * V13[0] - Y24 * V24[0] - MY13 * VHI[0]
              + RL[2][0] + RL[2][1] + RL[2][2] + RL[2][3];
        #pragma unroll
        for (int k = 0; k < 4; k++) AD[k] = VQN[0][k] * RL[0][k] + VQN[1][k] * RL[1][k] + VQN[2][k] * RL[2][k];
        double DBAD[3], ALR[3];
        #pragma unroll
        for (int c = 0; c < 3; c++) DBAD[c] = DB[c][0] * AD[0] + DB[c][1] * AD[1] + DB[c][2] * AD[2] + DB[c][3] * AD[3];
        ALR[0] = DI[0] * AR[0] + DI[3] * AR[1] + DI[4] * AR[2] - DBAD[0];
        ALR[1] = DI[3] * AR[0] + DI[1] * AR[1] + DI[5] * AR[2] - DBAD[1];
        ALR[2] = DI[4] * AR[0] + DI[5] * AR[1] + DI[2] * AR[2] - DBAD[2];
        const double C1 = K_TWO * ALR[2];
        V13[0] = V13[0] + C1 * Y13; V24[0] = V24[0] + C1 * Y24;
        VHI[0] = VHI[0] + K_FOUR * (ALR[2] * MY13 - Z1 * ALR[1]);
        V13[1] = V13[1] - C1 * X13; V24[1] = V24[1] - C1 * X24;
        VHI[1] = VHI[1] - K_FOUR * (ALR[2] * MX13 - Z1 * ALR[0]);
        V13[2] = V13[2] - K_TWO * (Y13 * ALR[0] - X13 * ALR[1]);
        V24[2] = V24[2] - K_TWO * (Y24 * ALR[0] - X24 * ALR[1]);
        VHI[2] = VHI[2] + K_FOUR * (MX13 * ALR[1] - MY13 * ALR[0]);
        #pragma unroll
        for (int k = 0; k < 4; k++) {
          const double ALD = AD[k] + VQN[0][k] * DBAD[0] + VQN[1][k] * DBAD[1] + VQN[2][k] * DBAD[2]
                           - DB[0][k] * AR[0] - DB[1][k] * AR[1] - DB[2][k] * AR[2];
          RL[0][k] = RL[0][k] - ALR[0] - VQN[0][k] * ALD;
          RL[1][k] = RL[1][k] - ALR[1] - VQN[1][k] * ALD;
        }
      }
    }
    #pragma unroll
    for (int c = 0; c < 3; c++) { V13[c] = V13[c] * AREA_I; V24[c] = V24[c] * AREA_I; VHI[c] = VHI[c] * K_FOURTH; }
    // ---- CNCOEF3B
    THK0 = (g.prop.ithk > 0) ? fmax(K_EM20, T.ld(SW_THK)) : T.ld(g.w_thke);
    double RHO, G;
    if (LAW != 2) { const orgpu_law36& m = g.m36; RHO = m.rho0; G = m.shear; io.ssp = m.ssp; }
    else           { const orgpu_law2& m = g.m2; RHO = m.rho0; G = m.shear; io.ssp = m.ssp; }
    const double SHF = (NPT == 1) ? K_ZERO : g.prop.shf;
    // ---- CZDEF
    double VDEF[8];
    {
      double R13v[2], R24v[2], RSOM[2], RHI[2];
      #pragma unroll
      for (int c = 0; c < 2; c++) {
        R13v[c] = (RL[c][0] - RL[c][2]) * AREA_I;
        R24v[c] = (RL[c][1] - RL[c][3]) * AREA_I;
        RSOM[c] = (RL[c][3] + RL[c][2] + RL[c][0] + RL[c][1]) * AREA_I;
        RHI[c] = (RL[c][0] - RL[c][1] + RL[c][2] - RL[c][3]) * K_FOURTH;
      }
      VDEF[0] = Y24 * V13[0] - Y13 * V24[0];
      VDEF[1] = -X24 * V13[1] + X13 * V24[1];
      const double BXV2 = Y24 * V13[1] - Y13 * V24[1];
      const double BYV1 = -X24 * V13[0] + X13 * V24[0];
      VDEF[2] = BXV2 + BYV1;
      VDEF[5] = Y24 * R13v[1] - Y13 * R24v[1];
      VDEF[6] = X24 * R13v[0] - X13 * R24v[0];
      const double BXR1 = Y13 * R24v[0] - Y24 * R13v[0];
      const double BYR2 = -X24 * R13v[1] + X13 * R24v[1];
      VDEF[7] = BXR1 + BYR2;
      const double BCXY = AREA * K_FOURTH;
      const double BCX = V13[2] - MY13 * R13v[0] + MX13 * R13v[1];
      const double BCY = V24[2] + MY13 * R24v[0] - MX13 * R24v[1];
      VDEF[3] = Y24 * BCX - Y13 * BCY + BCXY * RSOM[1];
      VDEF[4] = X13 * BCY - X24 * BCX - BCXY * RSOM[0];
      VHG[0] = VHI[0] - MX13 * VDEF[0] - MY13 * BYV1;
      VHG[1] = VHI[1] - MX13 * BXV2 - MY13 * VDEF[1];
      VHG[2] = RHI[1] - MX13 * VDEF[5] - MY13 * BYR2;
      VHG[3] = -RHI[0] - MX13 * BXR1 - MY13 * VDEF[6];
      VHG[4] = (VHI[2] * K_FOUR - (MY13 * RSOM[0] - MY23 * (R13v[0] + R24v[0]) + MX23 * (R13v[1] + R24v[1]) - MX13 * RSOM[1]) * AREA) * K_FOUR;
      VHG[5] = (VHI[2] * K_FOUR - (MY13 * RSOM[0] - MY34 * (R13v[0] - R24v[0]) + MX34 * (R13v[1] - R24v[1]) - MX13 * RSOM[1]) * AREA) * K_FOUR;
      VHG[0] = VHG[0] + (Y24 * V13[2] - Y13 * V24[2]) * Z1;
      VHG[1] = VHG[1] + (-X24 * V13[2] + X13 * V24[2]) * Z1;
      const double DETA = Z1 * K_FOUR * AREA_I;
      VDEF[5] = VDEF[5] + (X13 * V13[0] - X24 * V24[0]) * DETA;
      VDEF[6] = VDEF[6] + (Y13 * V13[1] - Y24 * V24[1]) * DETA;
      VDEF[7] = VDEF[7] + (X13 * V13[1] - X24 * V24[1] + Y13 * V13[0] - Y24 * V24[0]) * DETA;
      OFF = fmin(K_ONE, fabs(OFFG));
      if (OFFG < K_ZERO) {
        #pragma unroll
        for (int k = 0; k < 8; k++) VDEF[k] = K_ZERO;
        #pragma unroll
        for (int k = 0; k < 6; k++) VHG[k] = K_ZERO;
      }
    }
    // ---- CZSTRA3 + element strain rate
    io.exx = VDEF[0] * DT1; io.eyy = VDEF[1] * DT1; io.exy = VDEF[2] * DT1;
    io.eyz = VDEF[4] * DT1; io.exz = VDEF[3] * DT1;
    io.kxx = VDEF[5] * DT1; io.kyy = VDEF[6] * DT1; io.kxy = VDEF[7] * DT1;
    if (g.prop.istrain != 0) {
      const double de[8] = {io.exx, io.eyy, io.exy, io.eyz, io.exz, io.kxx, io.kyy, io.kxy};
      double st[8];                          // all eight loads in flight before the first store (which could alias them)
      #pragma unroll
      for (int k = 0; k < 8; k++) st[k] = T.ld(SW_STRA + k);
      #pragma unroll
      for (int k = 0; k < 8; k++) T.st(SW_STRA + k, st[k] + de[k]);
    }
    {
      const double dtinv = or_div(DT1, fmax(DT1 * DT1, K_EM20));
      const double thk = T.ld(SW_THK);
      const double eps_k2 = (io.kxx * io.kxx + io.kyy * io.kyy + io.kxx * io.kyy + K_FOURTH * (io.kxy * io.kxy)) * K_ONE_OVER_9 * (thk * thk);
      const double eps_m2 = K_FOUR_OVER_3 * (io.exx * io.exx + io.eyy * io.eyy + io.exx * io.eyy + K_FOURTH * (io.exy * io.exy));
      io.epsd_pg = or_sqrt(eps_k2 + eps_m2) * dtinv;
      T.st(SW_EPSD, K_ONE * io.epsd_pg + (K_ONE - K_ONE) * T.ld(SW_EPSD));
    }
    io.area = AREA; io.thk0 = THK0; io.gs = G * SHF; io.rho = RHO; io.off = OFF; io.sigy = K_EP30;
    }   // ======== end of the pre-loop scope ========
    ORGPU_OPAQUE(XL2); ORGPU_OPAQUE(YL2); ORGPU_OPAQUE(XL3); ORGPU_OPAQUE(YL3); ORGPU_OPAQUE(XL4); ORGPU_OPAQUE(YL4);
    ORGPU_OPAQUE(Z1); ORGPU_OPAQUE(AREA);
    // ---- CMAIN3
    PHASE_SYNC(0);
#ifndef ORGPU_NO_COMPACT
    if constexpr (FAST >= 1 && STAGED) shell_material_loop_compact<true, FAST>(g, T, DT1, io, wmask, full_tile);
    else
#endif
    shell_material_loop<LAW, true, STAGED, 0, FAST>(g, T, DT1, io);
    OFF = io.off;
    PHASE_SYNC(1);
#ifndef ORGPU_NO_BILAN
    if (g.bal && P.cs->ipri) shell_bilan<4, STAGED>(P, T, tile, e, io.rho, OFF);     // CBILAN (czforc3.F:639)
#endif
    // ---- re-derive the geometry needed by the force assembly (same expressions as before the loop)
    QephGeo q1; qeph_geo(XL2, YL2, XL3, YL3, XL4, YL4, q1);
    const double* CX = q1.CX; const double* CY = q1.CY;
    const double X13 = q1.X13, X24 = q1.X24, Y13 = q1.Y13, Y24 = q1.Y24;
    const double MX13 = q1.MX13, MX23 = q1.MX23, MX34 = q1.MX34, MY13 = q1.MY13, MY23 = q1.MY23, MY34 = q1.MY34;
    const double THK02 = THK0 * THK0;
    double A11, A12, GSR, A11SR, A12SR;
    const double RHO = io.rho;
    double G;
    if (LAW != 2) { const orgpu_law36& m = g.m36; G = m.shear; A11 = m.a11; GSR = m.gsr; A11SR = m.a11sr; A12 = m.nu * A11; A12SR = m.nusr * A11SR; }
    else           { const orgpu_law2& m = g.m2; G = m.shear; A11 = m.a11; A12 = m.a12; GSR = m.gsr; A11SR = m.a11sr; A12SR = m.a12sr; }
    const double SHF = (NPT == 1) ? K_ZERO : g.prop.shf, SHFSR = (NPT == 1) ? K_ZERO : g.prop.shfsr;
    const double AMU = (g.prop.h1 == K_ZERO) ? K_ZEP01 + K_FIVEEM3 : g.prop.h1;
    // ---- CNDT3
    PHASE_SYNC(4);
    double STI, STIR = K_ZERO;
    {
      double VISCMX = fmax(io.viscmx, AMU);
      VISCMX = or_sqrt(K_ONE + VISCMX * VISCMX) - VISCMX;
      const double ALDT = LL * VISCMX;      // / sqrt(ALPE), ALPE = 1: exact
      const double F_OSET = K_ONE + K_ZERO;   // HALF*|Z_OFFSET*THK0|/THK0 with zero offset: exactly +0
      const double F_DTE = or_div(K_ONE, or_sqrt(F_OSET));
      const double DT = or_div(g.dtfac * F_DTE * ALDT, io.ssp);
      if (g.nodadt != 0) {                                  // cndt3.F:194-221 (IGTYP=1), no element time step (:231, :294)
        if (OFF == K_ZERO) { STI = K_ZERO; STIR = K_ZERO; }
        else { STI = or_div(K_HALF * F_OSET * io.vol0 * A11, ALDT * ALDT);
               STIR = STI * (THK0 * THK0 + AREA) * K_ONE_OVER_12 + STI * (K_ZERO * THK0) * (K_ZERO * THK0); }
      } else {
      if (OFFG > K_ZERO && OFF != K_ZERO) dt_cand = DT;
      const double DIVM = fmax(ALDT * ALDT, K_EM20);
      STI = or_div(K_HALF * F_OSET * io.vol0 * A11 * OFF, DIVM);
      }
    }
    // ---- CZFINTCE : constant part of the generalised internal forces
    const double* FO = io.fo; const double* MO = io.mo;
    double VF[3][4], VM[2][4];
    {
      const double X13S8 = X13 * MO[2], X24S8 = X24 * MO[2], Y13S8 = Y13 * MO[2], Y24S8 = Y24 * MO[2];
      const double S1 = (MY34 * MX23 - MY23 * MX34) * THK0;
      const double S42S = S1 * FO[4], S52S = S1 * FO[3];
      VF[0][0] = THK0 * (Y24 * FO[0] - X24 * FO[2]);
      VF[1][0] = THK0 * (-X24 * FO[1] + Y24 * FO[2]);
      VF[2][0] = THK0 * (-X24 * FO[3] + Y24 * FO[4]);
      VM[0][0] = THK02 * (X24 * MO[1] - Y24S8) - MY13 * VF[2][0];
      VM[1][0] = THK02 * (Y24 * MO[0] - X24S8) + MX13 * VF[2][0];
      VM[0][2] = -S52S; VM[1][2] = S42S;
      VF[0][1] = THK0 * (-Y13 * FO[0] + X13 * FO[2]);
      VF[1][1] = THK0 * (X13 * FO[1] - Y13 * FO[2]);
      VF[2][1] = THK0 * (X13 * FO[3] - Y13 * FO[4]);
      VM[0][1] = THK02 * (-X13 * MO[1] + Y13S8) + MY13 * VF[2][1];
      VM[1][1] = THK02 * (-Y13 * MO[0] + X13S8) - MX13 * VF[2][1];
      VM[0][3] = VM[0][2]; VM[1][3] = VM[1][2];
      const double C2 = THK02 * Z1 * K_FOUR * AREA_I;
      VF[0][0] = VF[0][0] + C2 * (X13 * MO[0] + Y13S8);
      VF[1][0] = VF[1][0] + C2 * (Y13 * MO[1] + X13S8);
      VF[0][1] = VF[0][1] - C2 * (X24 * MO[0] + Y24S8);
      VF[1][1] = VF[1][1] - C2 * (Y24 * MO[1] + X24S8);
    }
    // ---- CZFINTN1 : elasto-plastic hourglass stresses + linear damping
    PHASE_SYNC(2);
    {
      const double FAC1 = g.prop.cvis;
      const double C7 = K_FOUR_OVER_3;
      const double FBEND = (NPT == 1) ? K_ZERO : K_ONE_OVER_12, FBEND_V = (NPT == 1) ? K_ZERO : K_THREEP464;
      const double COEF1 = (NPT == 0) ? K_SIXTEEN : K_TWENTY5;
      double VG[12], DG[12], DHG[6];
      #pragma unroll
      for (int k = 0; k < 12; k++) VG[k] = T.ld(SW_HOURG + k);
      #pragma unroll
      for (int k = 0; k < 6; k++) DHG[k] = VHG[k] * DT1;
      const double C3 = K_FOUR * AREA_I;
      const double HXX = C3 * MY34, HYY = C3 * MX34, HXX_K = C3 * MY23, HYY_K = C3 * MX23;
      {
        const double CXX = HXX * DHG[0], CYY = HYY * DHG[1], CXX_K = HXX_K * DHG[0], CYY_K = HYY_K * DHG[1];
        const double BXX = HXX * DHG[2], BYY = HYY * DHG[3], BXX_K = HXX_K * DHG[2], BYY_K = HYY_K * DHG[3];
        const double C1M = A11 * FAC1, C2M = A12 * FAC1;
        DG[0] = C1M * CXX - C2M * CYY;     DG[1] = C1M * CYY - C2M * CXX;
        DG[2] = C1M * BXX - C2M * BYY;     DG[3] = C1M * BYY - C2M * BXX;
        DG[6] = C1M * CXX_K - C2M * CYY_K; DG[7] = C1M * CYY_K - C2M * CXX_K;
        DG[8] = C1M * BXX_K - C2M * BYY_K; DG[9] = C1M * BYY_K - C2M * BXX_K;
        const double C2 = FAC1 * G * SHF * K_ONE_OVER_64;
        DG[4] = C2 * HXX * DHG[4];  DG[5] = C2 * HYY * DHG[4];
        DG[10] = C2 * HXX_K * DHG[5]; DG[11] = C2 * HYY_K * DHG[5];
      }
      const double C6 = THK02 * FBEND;
      double SS1 = MY34 * VG[0] + MY23 * VG[6];
      double SS2 = MX23 * VG[7] + MX34 * VG[1];
      double SF1 = MY34 * VG[2] + MY23 * VG[8];
      double SF2 = -MX23 * VG[9] - MX34 * VG[3];
      double SC5 = MY34 * VG[4] + MX34 * VG[5];
      double SC6 = MY23 * VG[10] + MX23 * VG[11];
      const double C5 = K_HALF * OFF * THK0 * C7;
      const double ESX = SS1 * DHG[0] + SS2 * DHG[1];
      double ein1 = T.ld(SW_EINT), ein2 = T.ld(SW_EINT + 1);
      ein1 = ein1 + C5 * (ESX + K_FOURTH * (SC5 * DHG[4] + SC6 * DHG[5]));
      const double EMX = (SF1 * DHG[2] - SF2 * DHG[3]) * C6;
      ein2 = ein2 + C5 * EMX;
      #pragma unroll
      for (int k = 0; k < 12; k++) VG[k] = VG[k] + DG[k];
      if (io.sigy < K_ZEP9EP30) {
        const double UFAC = fabs(fmin(io.zcfac1, io.zcfac2) - K_ONE);
        const double SIGY2 = io.sigy * io.sigy;
        double SVM = K_ZERO, SXY0 = K_ZERO;
        if (UFAC < K_EM18) {
          SXY0 = FO[0] * FO[0] + FO[1] * FO[1] - FO[0] * FO[1] + K_THREE * FO[2] * FO[2];
          double MXY0 = MO[0] * MO[0] + MO[1] * MO[1] - MO[0] * MO[1] + K_THREE * MO[2] * MO[2];
          const double CNN = K_ZEP85, CMM = K_ZEP85 * THK0 * K_ONE_OVER_16;
          const double CNNX = CNN * VG[0], CNNY = CNN * VG[1], CNNX_K = CNN * VG[6], CNNY_K = CNN * VG[7];
          const double CMMX = CMM * VG[2], CMMY = CMM * VG[3], CMMX_K = CMM * VG[8], CMMY_K = CMM * VG[9];
          SXY0 = SXY0 + CNNX * CNNX + CNNY * CNNY - CNNX * CNNY;
          MXY0 = MXY0 + CMMX * CMMX + CMMY * CMMY - CMMX * CMMY;
          SXY0 = SXY0 + CNNX_K * CNNX_K + CNNY_K * CNNY_K - CNNX_K * CNNY_K;
          MXY0 = MXY0 + CMMX_K * CMMX_K + CMMY_K * CMMY_K - CMMX_K * CMMY_K;
          SXY0 = SXY0 + fabs(CNNX * (K_TWO * CNNX_K - CNNY_K) + CNNY * (K_TWO * CNNY_K - CNNX_K));
          MXY0 = MXY0 + fabs(CMMX * (K_TWO * CMMX_K - CMMY_K) + CMMY * (K_TWO * CMMY_K - CMMX_K));
          SVM = SXY0 + COEF1 * MXY0;
        }
        if (UFAC >= K_EM18 || SVM > SIGY2) {
          double EH1 = fmin(or_div(SXY0, fmax(SIGY2, K_EM18)), K_ONE);
          EH1 = fmax(K_ZEP999 * EH1, (K_ONE - io.zcfac1));
          double EH2 = fmax(K_ZEP999, (K_ONE - io.zcfac2));
          if (ESX < K_ZERO) EH1 = K_ZERO;
          if (EMX < K_ZERO) EH2 = K_ZERO;
          VG[0] = VG[0] - EH1 * DG[0]; VG[1] = VG[1] - EH1 * DG[1]; VG[6] = VG[6] - EH1 * DG[6]; VG[7] = VG[7] - EH1 * DG[7];
          VG[2] = VG[2] - EH2 * DG[2]; VG[3] = VG[3] - EH2 * DG[3]; VG[8] = VG[8] - EH2 * DG[8]; VG[9] = VG[9] - EH2 * DG[9];
        }
      }
      #pragma unroll
      for (int k = 0; k < 12; k++) T.st(SW_HOURG + k, VG[k]);
      const double C8 = C7 * OFF;
      SS1 = (MY34 * VG[0] + MY23 * VG[6]) * C8;
      SS2 = (MX23 * VG[7] + MX34 * VG[1]) * C8;
      SF1 = (MY34 * VG[2] + MY23 * VG[8]) * C8;
      SF2 = -(MX23 * VG[9] + MX34 * VG[3]) * C8;
      const double HSURA = THK0 * AREA_I;
      double C2 = C8 * THK0;
      SC5 = (MY34 * VG[4] + MX34 * VG[5]) * C2;
      SC6 = (MY23 * VG[10] + MX23 * VG[11]) * C2;
      double SS3 = SC5 + SC6;
      const double HVL = AMU * or_sqrt(RHO * AREA * FAC1) * OFF;
      const double SSV0 = MY23 * MY23, SSV1 = MY34 * MY34, SSV2 = MX23 * MX23, SSV3 = MX34 * MX34;
      const double HXX_V = K_FIVEP333 * (SSV1 + SSV0);
      const double HXY_V = -K_FIVEP333 * (MY34 * MX34 + MY23 * MX23);
      const double HYY_V = K_FIVEP333 * (SSV2 + SSV3);
      C2 = HVL * GSR * SHFSR * or_sqrt(K_ONE_OVER_12);
      const double CXZ_V = (SSV1 + SSV3) * C2, CYZ_V = (SSV2 + SSV0) * C2;
      const double AUX = AREA_I * HVL;
      const double C1Mv = A11SR * AUX, C2Mv = A12SR * AUX;
      const double CXX_V = C1Mv * HXX_V, CYY_V = C1Mv * HYY_V, CXY_V = C2Mv * HXY_V;
      const double SS1_V = CXX_V * VHG[0] + CXY_V * VHG[1];
      const double SS2_V = CYY_V * VHG[1] + CXY_V * VHG[0];
      const double SF1_V = (CXX_V * VHG[2] + CXY_V * VHG[3]) * FBEND_V;
      const double SF2_V = (-CYY_V * VHG[3] - CXY_V * VHG[2]) * FBEND_V;
      const double SC5_V = CXZ_V * VHG[4] * HSURA;
      const double SC6_V = CYZ_V * VHG[5] * HSURA;
      const double SS3_V = SC5_V + SC6_V;
      SS1 = SS1 + SS1_V; SS2 = SS2 + SS2_V; SS3 = SS3 + SS3_V; SC5 = SC5 + SC5_V; SC6 = SC6 + SC6_V; SF1 = SF1 + SF1_V; SF2 = SF2 + SF2_V;
      const double Y13S = MY13 * SS3, X13S = MX13 * SS3, Y34S6 = MY34 * SC6, Y23S5 = MY23 * SC5, X23S5 = MX23 * SC5, X34S6 = MX34 * SC6;
      C2 = K_FOURTH * THK0;
      const double B13 = (MY13 * X24 - MX13 * Y24) * HSURA;
      VF[0][0] = VF[0][0] + B13 * SS1; VF[0][2] = C2 * SS1;
      VF[1][0] = VF[1][0] + B13 * SS2; VF[1][2] = C2 * SS2;
      VF[2][2] = SS3;
      const double B24 = (MX13 * Y13 - MY13 * X13) * HSURA;
      VF[0][1] = VF[0][1] + B24 * SS1; VF[0][3] = -VF[0][2];
      VF[1][1] = VF[1][1] + B24 * SS2; VF[1][3] = -VF[1][2];
      VF[2][3] = -VF[2][2];
      double C3b = C6 * B13; const double C4 = C6 * C2;
      VM[0][0] = VM[0][0] + C3b * SF2 + Y23S5 + Y34S6;
      VM[0][2] = VM[0][2] + C4 * SF2 - Y13S;
      VM[1][0] = VM[1][0] + C3b * SF1 - X23S5 - X34S6;
      VM[1][2] = VM[1][2] + C4 * SF1 + X13S;
      C3b = C6 * B24;
      VM[0][1] = VM[0][1] + C3b * SF2 + Y23S5 - Y34S6;
      VM[0][3] = VM[0][3] - C4 * SF2 - Y13S;
      VM[1][1] = VM[1][1] + C3b * SF1 - X23S5 + X34S6;
      VM[1][3] = VM[1][3] - C4 * SF1 + X13S;
      C2 = Z1 * HSURA;
      VF[2][0] = VF[2][0] + C2 * (SS1 * Y24 - SS2 * X24);
      VF[2][1] = VF[2][1] + C2 * (-SS1 * Y13 + SS2 * X13);
      const double ESY = ((SS1 - SS1_V) * DHG[0] + (SS2 - SS2_V) * DHG[1]) * THK0 + K_FOURTH * ((SC5 - SC5_V) * DHG[4] + (SC6 - SC6_V) * DHG[5]);
      ein1 = ein1 + K_HALF * ESY;
      const double EMY = (SF1 - SF1_V) * DHG[2] - (SF2 - SF2_V) * DHG[3];
      ein2 = ein2 + K_HALF * C6 * EMY * THK0;
      T.st(SW_EINT, ein1); T.st(SW_EINT + 1, ein2);
    }
    // ---- CZPROJN (IFINI=0) + CUPDTN3P
    PHASE_SYNC(3);
    if (OFF < K_ONE) OFFG = OFF;
    T.st(SW_OFF, OFFG);
    const bool dead = OFFG < K_ZERO;
    if (dead) { STI = K_ZERO; STIR = K_ZERO; }
    double FL[3][4], MM[3][4];
    #pragma unroll
    for (int c = 0; c < 3; c++) {
      FL[c][0] = VF[c][0] + VF[c][2]; FL[c][1] = VF[c][1] + VF[c][3];
      FL[c][2] = -VF[c][0] + VF[c][2]; FL[c][3] = -VF[c][1] + VF[c][3];
    }
    #pragma unroll
    for (int c = 0; c < 2; c++) {
      MM[c][0] = VM[c][0] + VM[c][2]; MM[c][1] = VM[c][1] + VM[c][3];
      MM[c][2] = -VM[c][0] + VM[c][2]; MM[c][3] = -VM[c][1] + VM[c][3];
    }
    if (!PLAT) {
      double VQN[3][4], DI[6], DB[3][4];
      qeph_warp_proj(q1, Z1, AREA, VQN, DI, DB);
      double AR[3], AD[4], DBAD[3], ALR[3];
      AR[0] = -Z1 * (FL[1][0] - FL[1][1] + FL[1][2] - FL[1][3])
            + CY[0] * FL[2][0] + MM[0][0] + CY[1] * FL[2][1] + MM[0][1] + CY[2] * FL[2][2] + MM[0][2] + CY[3] * FL[2][3] + MM[0][3];
      AR[1] = Z1 * (FL[0][0] - FL[0][1] + FL[0][2] - FL[0][3])
            - CX[0] * FL[2][0] + MM[1][0] - CX[1] * FL[2][1] + MM[1][1] - CX[2] * FL[2][2] + MM[1][2] - CX[3] * FL[2][3] + MM[1][3];
      AR[2] = -CY[0] * FL[0][0] + CX[0] * FL[1][0] - CY[1] * FL[0][1] + CX[1] * FL[1][1]
            - CY[2] * FL[0][2] + CX[2] * FL[1][2] - CY[3] * FL[0][3] + CX[3] * FL[1][3];
      #pragma unroll
      for (int k = 0; k < 4; k++) AD[k] = VQN[0][k] * MM[0][k] + VQN[1][k] * MM[1][k];
      #pragma unroll
      for (int c = 0; c < 3; c++) DBAD[c] = DB[c][0] * AD[0] + DB[c][1] * AD[1] + DB[c][2] * AD[2] + DB[c][3] * AD[3];
      ALR[0] = DI[0] * AR[0] + DI[3] * AR[1] + DI[4] * AR[2] - DBAD[0];
      ALR[1] = DI[3] * AR[0] + DI[1] * AR[1] + DI[5] * AR[2] - DBAD[1];
      ALR[2] = DI[4] * AR[0] + DI[5] * AR[1] + DI[2] * AR[2] - DBAD[2];
      double C1 = Z1 * ALR[1];
      FL[0][0] = FL[0][0] - C1 + CY[0] * ALR[2]; FL[0][1] = FL[0][1] + C1 + CY[1] * ALR[2];
      FL[0][2] = FL[0][2] - C1 + CY[2] * ALR[2]; FL[0][3] = FL[0][3] + C1 + CY[3] * ALR[2];
      C1 = Z1 * ALR[0];
      FL[1][0] = FL[1][0] + C1 - CX[0] * ALR[2]; FL[1][1] = FL[1][1] - C1 - CX[1] * ALR[2];
      FL[1][2] = FL[1][2] + C1 - CX[2] * ALR[2]; FL[1][3] = FL[1][3] - C1 - CX[3] * ALR[2];
      #pragma unroll
      for (int J = 0; J < 4; J++) {
        const double ALD = AD[J] + VQN[0][J] * DBAD[0] + VQN[1][J] * DBAD[1] + VQN[2][J] * DBAD[2]
                         - DB[0][J] * AR[0] - DB[1][J] * AR[1] - DB[2][J] * AR[2];
        FL[2][J] = FL[2][J] - CY[J] * ALR[0] + CX[J] * ALR[1];
        MM[0][J] = MM[0][J] - ALR[0] - VQN[0][J] * ALD;
        MM[1][J] = MM[1][J] - ALR[1] - VQN[1][J] * ALD;
        MM[2][J] = -ALR[2] - VQN[2][J] * ALD;
      }
    }
    int sl[4];
    #pragma unroll
    for (int k = 0; k < 4; k++) sl[k] = T.ldi(g.w_slot, k);
    #pragma unroll
    for (int J = 0; J < 4; J++) {
      double f[3], mm[3];
      #pragma unroll
      for (int I = 0; I < 3; I++) {
        f[I] = VQ[I][0] * FL[0][J] + VQ[I][1] * FL[1][J] + VQ[I][2] * FL[2][J];
        mm[I] = PLAT ? VQ[I][0] * MM[0][J] + VQ[I][1] * MM[1][J]
                     : VQ[I][0] * MM[0][J] + VQ[I][1] * MM[1][J] + VQ[I][2] * MM[2][J];
        if (dead) { f[I] = K_ZERO; mm[I] = K_ZERO; }
      }
      const double fac = (J & 1) ? FACN2 : FACN1;
      double4* row = reinterpret_cast<double4*>(P.fsky + (size_t)8 * sl[J]);
      const double4 r0 = make_double4(-f[0], -f[1], -f[2], -mm[0]), r1 = make_double4(-mm[1], -mm[2], STI * fac, STIR * fac);
      st256(row, r0); st256(row + 1, r1);
    }
#ifndef ORGPU_NO_XSEND
    if (g.xs_ftile && g.xs_ftile[tile]) xsend_rows<8, STAGED>(P.nd.xs, T, g.w_slot, 4, P.fsky);   // frontier tile: rows to the neighbours' windows
#endif
  }
  cta_epilogue<false, STAGED>(dt_cand, order, P.db, g.blk0 + tile, g_tile, s_tile_dyn, (unsigned)g.nw_rw * ORGPU_TILE * 8u);
}
