// brick_kernel.cuh -- fused internal-force kernel for 8-node bricks, one element per thread.
//
// One launch does, for every element of a super-group, what the reference does in SFORC3
// (engine/source/elements/solid/solide/sforc3.F:131, Lagrangian JCVT=0 path) for one group:
//   SCOOR3 (scoor3.F:107-386) gather            -> 16 x 32-byte nodal records (pos, vel)
//   SDERI3 (sderi3.F:105-336) + SCHKJABT3       -> Jacobian, volume, PXi..PZi, hourglass PXiHj
//   SDLEN3 (sdlen3.F:89-168, slen.F:68-88)      -> characteristic length
//   SDEFO3 (sdefo3.F:115-271)                   -> strain rates, spin
//   SRHO3 (srho3.F:110-239), SROTA3 (srota3.F:72-94), SMALLA3, S8SAV3
//   MMAIN -> M2LAW (m2law.F:133-561) + MQVISCB (mqviscb.F:141-631): stress, dt, nodal stiffness
//         or MULAW (mulaw.F90) -> SIGEPS36 (sigeps36.F) for LAW36 (template parameter LAW)
//   SMALLB3, SHVIS3 (shvis3.F:164-412), SFINT3 (sfint3.F:257-323)
//   SCUMU3P (scumu3p.F:104-309): corner rows into the FSKY slots of IADS
// followed by the CTA-level (dt, user id) arg-min.  The CTA's state tile (common.cuh) is read by one
// TMA bulk copy into shared memory and written back by one bulk store: every state word crosses HBM
// exactly once each way, and no thread ever waits on a state load.
#pragma once
#include "common.cuh"

struct BrickParams {            // scalar copies so the kernel reads them from the constant bank
  BrickSG sg; DevNodes nd; double* fsky; int roww; CycleState* cs; DtBlocks db;
  const BrickSG* sgtab; const int2* cta_map;      // batched launch over several super-groups (common.cuh cta_work); null otherwise
};

// SLEN (slen.F:68-88)
__device__ __forceinline__ double slen_face(const double* x, const double* y, const double* z,
                                            int a, int b, int c, int d)
{
  double X13 = x[c] - x[a], X24 = x[d] - x[b];
  double Y13 = y[c] - y[a], Y24 = y[d] - y[b];
  double Z13 = z[c] - z[a], Z24 = z[d] - z[b];
  double FS1 = X13 - X24, FT1 = X13 + X24;
  double FS2 = Y13 - Y24, FT2 = Y13 + Y24;
  double FS3 = Z13 - Z24, FT3 = Z13 + Z24;
  double E = FS1 * FS1 + FS2 * FS2 + FS3 * FS3;
  double F = FS1 * FT1 + FS2 * FT2 + FS3 * FT3;
  double G = FT1 * FT1 + FT2 * FT2 + FT3 * FT3;
  return E * G - F * F;
}


__device__ __forceinline__ double4 ldg4(const double4* p) {     // read-only 32-byte nodal record
  return ld256_nc(p);
}

struct Jac { double J1, J2, J3, J4, J5, J6, J7, J8, J9, c5968, c6749, c4857, vol; };
__device__ __forceinline__ Jac brick_jac(const double* x, const double* y, const double* z)
{
  Jac r;
  double X17 = x[6] - x[0], X28 = x[7] - x[1], X35 = x[4] - x[2], X46 = x[5] - x[3];
  double Y17 = y[6] - y[0], Y28 = y[7] - y[1], Y35 = y[4] - y[2], Y46 = y[5] - y[3];
  double Z17 = z[6] - z[0], Z28 = z[7] - z[1], Z35 = z[4] - z[2], Z46 = z[5] - z[3];
  r.J1 = X17 + X28 - X35 - X46;
  r.J2 = Y17 + Y28 - Y35 - Y46;
  r.J3 = Z17 + Z28 - Z35 - Z46;
  double xa = X17 + X46, xb = X28 + X35, ya = Y17 + Y46, yb = Y28 + Y35, za = Z17 + Z46, zb = Z28 + Z35;
  r.J4 = xa + xb; r.J5 = ya + yb; r.J6 = za + zb;
  r.J7 = xa - xb; r.J8 = ya - yb; r.J9 = za - zb;
  r.c5968 = r.J5 * r.J9 - r.J6 * r.J8;
  r.c6749 = r.J6 * r.J7 - r.J4 * r.J9;
  r.c4857 = r.J4 * r.J8 - r.J5 * r.J7;
  r.vol = K_ONE_OVER_64 * (r.J1 * r.c5968 + r.J2 * r.c6749 + r.J3 * r.c4857);
  return r;
}

// hourglass shape vectors gamma = h - (h.x) B  (shvis3.F:320-355), from PXkHm
#define BRICK_GAMMA(G_, PH) \
        G_[0][0] =  K_ONE - PH[0][0]; G_[0][1] = -K_ONE - PH[0][1]; G_[0][2] =  K_ONE - PH[0][2]; G_[0][3] = -K_ONE - PH[0][3]; \
        G_[0][4] =  K_ONE + PH[0][2]; G_[0][5] = -K_ONE + PH[0][3]; G_[0][6] =  K_ONE + PH[0][0]; G_[0][7] = -K_ONE + PH[0][1]; \
        G_[1][0] =  K_ONE - PH[1][0]; G_[1][1] =  K_ONE - PH[1][1]; G_[1][2] = -K_ONE - PH[1][2]; G_[1][3] = -K_ONE - PH[1][3]; \
        G_[1][4] = -K_ONE + PH[1][2]; G_[1][5] = -K_ONE + PH[1][3]; G_[1][6] =  K_ONE + PH[1][0]; G_[1][7] =  K_ONE + PH[1][1]; \
        G_[2][0] =  K_ONE - PH[2][0]; G_[2][1] = -K_ONE - PH[2][1]; G_[2][2] = -K_ONE - PH[2][2]; G_[2][3] =  K_ONE - PH[2][3]; \
        G_[2][4] = -K_ONE + PH[2][2]; G_[2][5] =  K_ONE + PH[2][3]; G_[2][6] =  K_ONE + PH[2][0]; G_[2][7] = -K_ONE + PH[2][1];


// MQVISCB (mqviscb.F:141-631; IMPL=0, N2D=0, NPG=1, JTHE=0, IDTMINS/=2): bulk viscosity, equivalent sound speed,
// nodal stiffness and the element's time-step candidate.  Shared by the M2LAW and MULAW branches.
template <int ISMSTR>
__device__ __forceinline__ void brick_mqviscb(const BrickSG& g, double DXX, double DYY, double DZZ, double SSP, double OFF, double OFFG,
                                              double VOLN, double VOLO, double RHON, double RHOREF, double CBV, double DELTAX,
                                              double& QNEW, double& SSP_EQ, double& STI, double& dt_cand)
{
  const orgpu_law2& m = g.mat;
  const double DD = -DXX - DYY - DZZ;
  double AD = K_ZERO, AL = K_ZERO;
  const double CX = SSP + K_ZERO;              // VD2 = 0 (Lagrangian)
  if (OFF == K_ONE) {
    AL = (VOLN > K_ZERO) ? CBV : K_ZERO;               // VOL**THIRD (mqviscb.F)
    AD = fmax(K_ZERO, DD);
  }
  const double NRHO = or_sqrt(RHOREF * m.rho0);
  const double QA = K_ONE * g.prop.qa, QB = K_ONE * g.prop.qb;
  const double CNS1_0 = 1.0 * g.prop.cns1, CNS2_0 = 1.0 * g.prop.cns2;
  const double QAA_0 = QA * QA;
  double CNS1 = CNS1_0 * AL * NRHO * SSP * OFF;
  double CNS2 = CNS2_0 * AL * NRHO * SSP * OFF;
  double QAA = QAA_0 * AD;
  double QX = QB * SSP + AL * QAA
            + K_ZERO /* K_ONE*K_TWO*K_ZERO / max(EM20, RHON*DELTAX): the thermal term of mqviscb.F, exactly +0 */
            + or_div((CNS1 + K_ONE * CNS2), fmax(K_EM20, RHOREF * DELTAX));
  QNEW = RHON * AD * AL * (QAA * AL + QB * SSP);
  SSP_EQ = fmax(K_EM20, QX + or_sqrt(QX * QX + CX * CX));
  double DTX = or_div(DELTAX, SSP_EQ);
  STI = K_ZERO;
  if (!(OFF == K_ZERO || OFFG < K_ZERO)) {
    double TIDT = or_div(K_ONE, DTX), TRHO, TVOL;
    if (ISMSTR == 1 && OFFG > K_ONE) { TRHO = m.rho0 * TIDT; TVOL = VOLO * TIDT; }
    else                             { TRHO = RHON * TIDT;   TVOL = VOLN * TIDT; }
    STI = TRHO * TVOL;
  }
  DTX = g.dtfac * DTX;
  if (VOLN > K_ZERO && OFF > K_ZERO && OFFG > K_ZERO && g.nodadt == 0) dt_cand = DTX;
}

#ifndef ORGPU_BRICK_MINB
#define ORGPU_BRICK_MINB 3
#endif

// JHBE_CVT = JHBE + 10 * JCVT: values 10, 11, 12 are Isolid 0, 1, 2 in Belytschko's co-rotational frame (SRCOOR3 instead of SCOOR3)
template <int JHBE_CVT, int ISMSTR, int LAW, bool STAGED, bool TAB = false>
__global__ void __launch_bounds__(ORGPU_BLOCK, ORGPU_BRICK_MINB * ORGPU_PER128)
brick_forces_kernel(const __grid_constant__ BrickParams P)
{
  constexpr int JHBE = JHBE_CVT % 10; constexpr bool CVT = JHBE_CVT >= 10;
  if (P.cs->abort) return;                               // sticky: a peer-memory wait timed out (exchange.cuh)
  const CtaWork<BrickSG> W = cta_work<BrickSG, TAB>(P.sg, P.sgtab, P.cta_map);
  const BrickSG& g = *W.g;
  const int tile = W.tile;
  const int e = tile * ORGPU_TILE + threadIdx.x;
  __shared__ __align__(8) unsigned long long s_bar;
  double* const g_tile = g.slab + (size_t)tile * g.nw * ORGPU_TILE;
  const double* const g_pf = W.tile_pf >= 0 ? W.g_pf->slab + (size_t)W.tile_pf * W.g_pf->nw * ORGPU_TILE : nullptr;   // tile of the CTA one wave ahead
  if (STAGED) tile_load_begin(s_tile_dyn, &s_bar, g_tile, (unsigned)g.nw * ORGPU_TILE * 8u, g_pf);
  const TileAcc<STAGED> T{(STAGED ? s_tile_dyn : g_tile) + threadIdx.x};
  if (!STAGED && threadIdx.x == 0 && g_pf)              // in-place tiles: same wave-ahead L2 prefetch
    bulk_prefetch_l2(g_pf, (unsigned)g.nw * ORGPU_TILE * 8u);
  // SMSTR (21 words, rewritten every cycle by S8SAV3 / SMALLA3) goes straight to HBM with streaming stores:
  // collecting it in shared memory for a bulk store was measured slower (0.472 vs 0.450 ms on C5)
  double* const sm = g.smstr + (size_t)tile * 21 * ORGPU_TILE + threadIdx.x;     // SMSTR word k at sm[k*TILE]
#if ORGPU_PREFETCH_NEXT > 0
  // a CTA about one wave ahead: start its connectivity toward L2 (its first load is then an L2 hit: -3 % kernel time)
  if (W.tile_nx >= 0 && threadIdx.x < (8 * ORGPU_TILE * 4) / 128) prefetch_l2(reinterpret_cast<const char*>(W.g_nx->conn + (size_t)W.tile_nx * 8 * ORGPU_TILE) + 128 * threadIdx.x);
#endif
  double dt_cand = K_EP30; int order = -1;
  if (e < g.ne) {
    const double DT1 = P.cs->dt2;                         // DT1 = DT2 of the previous cycle (resol.F:2721)
    const double TT = P.cs->tt;
    const orgpu_law2& m = g.mat;
    int nc[8];
    { const int* cn = g.conn + (size_t)tile * 8 * ORGPU_TILE + threadIdx.x;
      #pragma unroll
      for (int k = 0; k < 8; k++) nc[k] = __ldg(cn + k * ORGPU_TILE); }
    order = g.order0 + e;
    // ---- SCOOR3
    double x[8], y[8], z[8];
    #pragma unroll
    for (int k = 0; k < 8; k++) { double4 p = ldg4(P.nd.pos + nc[k]); x[k] = p.x; y[k] = p.y; z[k] = p.z; }
    #pragma unroll
    for (int k = 0; k < 8; k++) prefetch_l1(P.nd.vel + nc[k]);
    if (STAGED) mbar_wait(&s_bar, 0);                     // the state tile has landed (issued before the gather)
    double OFFG = T.ld(BW_OFF);
    const bool dying_in = OFFG < K_ZERO;                  // SCOOR3 zeroes the velocities of such an element (also what SBILAN sees)
    double OFF;
    double R11 = K_ONE, R12 = K_ZERO, R13 = K_ZERO, R21 = K_ZERO, R22 = K_ONE, R23 = K_ZERO, R31 = K_ZERO, R32 = K_ZERO, R33 = K_ONE;
    if (CVT) {
      // SRCOOR3 (srcoor3.F:265-303): frame of the iso-parametric axes of the current coordinates (SREPISO3 srepiso3.F:80-111),
      // made orthonormal by SORTHO3 (sortho3.F:75-150: three sweeps, then e1, e3 = e1 x v, e2 = e3 x e1); columns of R
      const double X17 = x[6] - x[0], X28 = x[7] - x[1], X35 = x[4] - x[2], X46 = x[5] - x[3];
      const double Y17 = y[6] - y[0], Y28 = y[7] - y[1], Y35 = y[4] - y[2], Y46 = y[5] - y[3];
      const double Z17 = z[6] - z[0], Z28 = z[7] - z[1], Z35 = z[4] - z[2], Z46 = z[5] - z[3];
      const double A17 = X17 + X46, A28 = X28 + X35, B17 = Y17 + Y46, B28 = Y28 + Y35, C17 = Z17 + Z46, C28 = Z28 + Z35;
      const double RX = X17 + X28 - X35 - X46, RY = Y17 + Y28 - Y35 - Y46, RZ = Z17 + Z28 - Z35 - Z46;
      const double SX = A17 + A28, SY = B17 + B28, SZ = C17 + C28;
      const double TX = A17 - A28, TY = B17 - B28, TZ = C17 - C28;
      double aa = or_sqrt(RX * RX + RY * RY + RZ * RZ); if (aa != K_ZERO) aa = or_div(K_ONE, aa);
      double Ux = RX * aa, Uy = RY * aa, Uz = RZ * aa;
      aa = or_sqrt(SX * SX + SY * SY + SZ * SZ); if (aa != K_ZERO) aa = or_div(K_ONE, aa);
      double Vx = SX * aa, Vy = SY * aa, Vz = SZ * aa;
      aa = or_sqrt(TX * TX + TY * TY + TZ * TZ); if (aa != K_ZERO) aa = or_div(K_ONE, aa);
      double Wx = TX * aa, Wy = TY * aa, Wz = TZ * aa;
      #pragma unroll 1
      for (int N = 0; N < 3; N++) {
        const double e1x = Vy * Wz - Vz * Wy + Ux, e1y = Vz * Wx - Vx * Wz + Uy, e1z = Vx * Wy - Vy * Wx + Uz;
        const double e2x = Wy * Uz - Wz * Uy + Vx, e2y = Wz * Ux - Wx * Uz + Vy, e2z = Wx * Uy - Wy * Ux + Vz;
        const double e3x = Uy * Vz - Uz * Vy + Wx, e3y = Uz * Vx - Ux * Vz + Wy, e3z = Ux * Vy - Uy * Vx + Wz;
        double bb = or_sqrt(e1x * e1x + e1y * e1y + e1z * e1z); if (bb != K_ZERO) bb = or_div(K_ONE, bb);
        Ux = e1x * bb; Uy = e1y * bb; Uz = e1z * bb;
        bb = or_sqrt(e2x * e2x + e2y * e2y + e2z * e2z); if (bb != K_ZERO) bb = or_div(K_ONE, bb);
        Vx = e2x * bb; Vy = e2y * bb; Vz = e2z * bb;
        bb = or_sqrt(e3x * e3x + e3y * e3y + e3z * e3z); if (bb != K_ZERO) bb = or_div(K_ONE, bb);
        Wx = e3x * bb; Wy = e3y * bb; Wz = e3z * bb;
      }
      double e3x = Uy * Vz - Uz * Vy, e3y = Uz * Vx - Ux * Vz, e3z = Ux * Vy - Uy * Vx;
      aa = or_sqrt(e3x * e3x + e3y * e3y + e3z * e3z); if (aa != K_ZERO) aa = or_div(K_ONE, aa);
      e3x = e3x * aa; e3y = e3y * aa; e3z = e3z * aa;
      R11 = Ux; R21 = Uy; R31 = Uz;
      R13 = e3x; R23 = e3y; R33 = e3z;
      R12 = e3y * Uz - e3z * Uy; R22 = e3z * Ux - e3x * Uz; R32 = e3x * Uy - e3y * Ux;
    }
    if (ISMSTR <= 4 && fabs(OFFG) > K_ONE) {
      #pragma unroll
      for (int k = 0; k < 7; k++) { x[k] = sm[(3 * k) * ORGPU_TILE]; y[k] = sm[(3 * k + 1) * ORGPU_TILE]; z[k] = sm[(3 * k + 2) * ORGPU_TILE]; }
      x[7] = K_ZERO; y[7] = K_ZERO; z[7] = K_ZERO;
      OFF = fabs(OFFG) - K_ONE;
    } else {
      OFF = fabs(OFFG);
      if (CVT) {                                          // X' = t(R) X (srcoor3.F:364-418); the saved reference above already lives in the frame
        #pragma unroll
        for (int k = 0; k < 8; k++) {
          const double XDL = R11 * x[k] + R21 * y[k] + R31 * z[k], YDL = R12 * x[k] + R22 * y[k] + R32 * z[k], ZDL = R13 * x[k] + R23 * y[k] + R33 * z[k];
          x[k] = XDL; y[k] = YDL; z[k] = ZDL;
        }
      }
    }
    // SDLEN3 works on the X1..Z8 copies taken here (before a possible negative-volume switch)
    double areamax = K_EM20;
    areamax = fmax(slen_face(x, y, z, 0, 1, 2, 3), areamax);
    areamax = fmax(slen_face(x, y, z, 4, 5, 6, 7), areamax);
    areamax = fmax(slen_face(x, y, z, 0, 1, 5, 4), areamax);
    areamax = fmax(slen_face(x, y, z, 1, 2, 6, 5), areamax);
    areamax = fmax(slen_face(x, y, z, 2, 3, 7, 6), areamax);
    areamax = fmax(slen_face(x, y, z, 3, 0, 4, 7), areamax);
    // ---- SDERI3 + SCHKJABT3
    Jac J = brick_jac(x, y, z);
    double VOLN = J.vol;
    if (OFF == K_ZERO) VOLN = K_ONE;
    else if (OFFG > K_ONE) {}
    else if (VOLN <= K_ZERO) {
      if (OFFG <= K_ONE && OFFG != K_ZERO) {             // switch to small strain: geometry from SAV
        #pragma unroll
        for (int k = 0; k < 7; k++) { x[k] = sm[(3 * k) * ORGPU_TILE]; y[k] = sm[(3 * k + 1) * ORGPU_TILE]; z[k] = sm[(3 * k + 2) * ORGPU_TILE]; }
        x[7] = K_ZERO; y[7] = K_ZERO; z[7] = K_ZERO;
        J = brick_jac(x, y, z); VOLN = J.vol; OFFG = K_TWO;
      }
    }
    double PX[4], PY[4], PZ[4];
    {
      double DETT = or_div(K_ONE_OVER_64, VOLN);
      double JI1 = DETT * J.c5968, JI4 = DETT * J.c6749, JI7 = DETT * J.c4857;
      double JI2 = DETT * (J.J3 * J.J8 - J.J2 * J.J9);
      double JI5 = DETT * (J.J1 * J.J9 - J.J3 * J.J7);
      double JI8 = DETT * (J.J2 * J.J7 - J.J1 * J.J8);
      double JI3 = DETT * (J.J2 * J.J6 - J.J3 * J.J5);
      double JI6 = DETT * (J.J3 * J.J4 - J.J1 * J.J6);
      double JI9 = DETT * (J.J1 * J.J5 - J.J2 * J.J4);
      double J12 = JI1 + JI2, J45 = JI4 + JI5, J78 = JI7 + JI8;
      PX[0] = -J12 - JI3; PY[0] = -J45 - JI6; PZ[0] = -J78 - JI9;
      PX[1] = -J12 + JI3; PY[1] = -J45 + JI6; PZ[1] = -J78 + JI9;
      J12 = JI1 - JI2; J45 = JI4 - JI5; J78 = JI7 - JI8;
      PX[2] = J12 + JI3; PY[2] = J45 + JI6; PZ[2] = J78 + JI9;
      PX[3] = J12 - JI3; PY[3] = J45 - JI6; PZ[3] = J78 - JI9;
    }
    double PH[3][4];                                      // PXkHm -> PH[m][k]
    if (JHBE != 0) {
      double HX, HY, HZ;
      HX = (x[0] - x[1] + x[2] - x[3] + x[4] - x[5] + x[6] - x[7]);
      HY = (y[0] - y[1] + y[2] - y[3] + y[4] - y[5] + y[6] - y[7]);
      HZ = (z[0] - z[1] + z[2] - z[3] + z[4] - z[5] + z[6] - z[7]);
      #pragma unroll
      for (int k = 0; k < 4; k++) PH[0][k] = PX[k] * HX + PY[k] * HY + PZ[k] * HZ;
      HX = (x[0] + x[1] - x[2] - x[3] - x[4] - x[5] + x[6] + x[7]);
      HY = (y[0] + y[1] - y[2] - y[3] - y[4] - y[5] + y[6] + y[7]);
      HZ = (z[0] + z[1] - z[2] - z[3] - z[4] - z[5] + z[6] + z[7]);
      #pragma unroll
      for (int k = 0; k < 4; k++) PH[1][k] = PX[k] * HX + PY[k] * HY + PZ[k] * HZ;
      HX = (x[0] - x[1] - x[2] + x[3] - x[4] + x[5] + x[6] - x[7]);
      HY = (y[0] - y[1] - y[2] + y[3] - y[4] + y[5] + y[6] - y[7]);
      HZ = (z[0] - z[1] - z[2] + z[3] - z[4] + z[5] + z[6] - z[7]);
      #pragma unroll
      for (int k = 0; k < 4; k++) PH[2][k] = PX[k] * HX + PY[k] * HY + PZ[k] * HZ;
    }
    const double DELTAX = or_div(K_FOUR * VOLN * K_ONE, or_sqrt(areamax));
    // ---- S8SAV3 (reference configuration refresh; done here while the coordinates are live)
    const bool sav_refresh = (ISMSTR <= 4) && (fabs(OFFG) <= K_ONE);
    if (sav_refresh) {                       // exclusive with SMALLA3 (OFFG > 1), so the order is free
      #pragma unroll
      for (int k = 0; k < 7; k++) {
        __stcs(&sm[(3 * k) * ORGPU_TILE], x[k] - x[7]);
        __stcs(&sm[(3 * k + 1) * ORGPU_TILE], y[k] - y[7]);
        __stcs(&sm[(3 * k + 2) * ORGPU_TILE], z[k] - z[7]);
      }
    }
    // ---- velocities (SCOOR3) and SDEFO3
    double vx[8], vy[8], vz[8];
    #pragma unroll
    for (int k = 0; k < 8; k++) { double4 p = ldg4(P.nd.vel + nc[k]); vx[k] = p.x; vy[k] = p.y; vz[k] = p.z; }
    if (CVT) {                                            // V' = t(R) V (srcoor3.F:672-681)
      #pragma unroll
      for (int k = 0; k < 8; k++) {
        const double X = R11 * vx[k] + R21 * vy[k] + R31 * vz[k], Y = R12 * vx[k] + R22 * vy[k] + R32 * vz[k], Z = R13 * vx[k] + R23 * vy[k] + R33 * vz[k];
        vx[k] = X; vy[k] = Y; vz[k] = Z;
      }
    }
    if (OFFG < K_ZERO) {
      #pragma unroll
      for (int k = 0; k < 8; k++) { vx[k] = K_ZERO; vy[k] = K_ZERO; vz[k] = K_ZERO; }
    }
    double DXX, DYY, DZZ, DXY, DXZ, DYX, DYZ, DZX, DZY, D4, D5, D6, WXX, WYY, WZZ;
    {
      double VX17 = vx[0] - vx[6], VX28 = vx[1] - vx[7], VX35 = vx[2] - vx[4], VX46 = vx[3] - vx[5];
      double VY17 = vy[0] - vy[6], VY28 = vy[1] - vy[7], VY35 = vy[2] - vy[4], VY46 = vy[3] - vy[5];
      double VZ17 = vz[0] - vz[6], VZ28 = vz[1] - vz[7], VZ35 = vz[2] - vz[4], VZ46 = vz[3] - vz[5];
      DXX = PX[0] * VX17 + PX[1] * VX28 + PX[2] * VX35 + PX[3] * VX46;
      DYY = PY[0] * VY17 + PY[1] * VY28 + PY[2] * VY35 + PY[3] * VY46;
      DZZ = PZ[0] * VZ17 + PZ[1] * VZ28 + PZ[2] * VZ35 + PZ[3] * VZ46;
      DXY = PY[0] * VX17 + PY[1] * VX28 + PY[2] * VX35 + PY[3] * VX46;
      DXZ = PZ[0] * VX17 + PZ[1] * VX28 + PZ[2] * VX35 + PZ[3] * VX46;
      DYX = PX[0] * VY17 + PX[1] * VY28 + PX[2] * VY35 + PX[3] * VY46;
      DYZ = PZ[0] * VY17 + PZ[1] * VY28 + PZ[2] * VY35 + PZ[3] * VY46;
      DZX = PX[0] * VZ17 + PX[1] * VZ28 + PX[2] * VZ35 + PX[3] * VZ46;
      DZY = PY[0] * VZ17 + PY[1] * VZ28 + PY[2] * VZ35 + PY[3] * VZ46;
    }
    // hourglass mode velocities (SHVIS3 first half) while the nodal velocities are live
    double HGX[4], HGY[4], HGZ[4];
    {
      double G_[3][8];
      if (JHBE == 0) {
        #define HG0(V, H) { double V3478 = V[2] - V[3] - V[6] + V[7], V2358 = V[1] - V[2] - V[4] + V[7], \
                                   V1467 = V[0] - V[3] - V[5] + V[6], V1256 = V[0] - V[1] - V[4] + V[5]; \
                            H[0] = V1467 - V2358; H[1] = V1467 + V2358; H[2] = V1256 - V3478; H[3] = V1256 + V3478; }
        HG0(vx, HGX) HG0(vy, HGY) HG0(vz, HGZ)
        #undef HG0
      } else {
        BRICK_GAMMA(G_, PH)
        #pragma unroll
        for (int mm = 0; mm < 3; mm++) {
          HGX[mm] = G_[mm][0] * vx[0] + G_[mm][1] * vx[1] + G_[mm][2] * vx[2] + G_[mm][3] * vx[3] + G_[mm][4] * vx[4] + G_[mm][5] * vx[5] + G_[mm][6] * vx[6] + G_[mm][7] * vx[7];
          HGY[mm] = G_[mm][0] * vy[0] + G_[mm][1] * vy[1] + G_[mm][2] * vy[2] + G_[mm][3] * vy[3] + G_[mm][4] * vy[4] + G_[mm][5] * vy[5] + G_[mm][6] * vy[6] + G_[mm][7] * vy[7];
          HGZ[mm] = G_[mm][0] * vz[0] + G_[mm][1] * vz[1] + G_[mm][2] * vz[2] + G_[mm][3] * vz[3] + G_[mm][4] * vz[4] + G_[mm][5] * vz[5] + G_[mm][6] * vz[6] + G_[mm][7] * vz[7];
        }
        HGX[3] = vx[0] - vx[1] + vx[2] - vx[3] - vx[4] + vx[5] - vx[6] + vx[7];
        HGY[3] = vy[0] - vy[1] + vy[2] - vy[3] - vy[4] + vy[5] - vy[6] + vy[7];
        HGZ[3] = vz[0] - vz[1] + vz[2] - vz[3] - vz[4] + vz[5] - vz[6] + vz[7];
      }
    }
    const double DT1D2 = K_HALF * DT1;
    if (CVT || JHBE >= 2) {                               // sdefo3.F:158-220 (co-rotational: no spin) / :222-256
      double EXX = DXX, EYY = DYY, EZZ = DZZ, EXY = DXY, EYX = DYX, EXZ = DXZ, EZX = DZX, EYZ = DYZ, EZY = DZY;
      DXX = DXX - DT1D2 * (EXX * EXX + EYX * EYX + EZX * EZX);
      DYY = DYY - DT1D2 * (EYY * EYY + EZY * EZY + EXY * EXY);
      DZZ = DZZ - DT1D2 * (EZZ * EZZ + EXZ * EXZ + EYZ * EYZ);
      double AAA = DT1D2 * (EXX * EXY + EYX * EYY + EZX * EZY);
      DXY = DXY - AAA; DYX = DYX - AAA; D4 = DXY + DYX;
      AAA = DT1D2 * (EYY * EYZ + EZY * EZZ + EXY * EXZ);
      DYZ = DYZ - AAA; DZY = DZY - AAA; D5 = DYZ + DZY;
      AAA = DT1D2 * (EZZ * EZX + EXZ * EXX + EYZ * EYX);
      DXZ = DXZ - AAA; DZX = DZX - AAA; D6 = DXZ + DZX;
      if (CVT) { WXX = K_ZERO; WYY = K_ZERO; WZZ = K_ZERO; }
      else {
      double PXX2 = PX[0] * PX[0] + PX[1] * PX[1] + PX[2] * PX[2] + PX[3] * PX[3];
      double PYY2 = PY[0] * PY[0] + PY[1] * PY[1] + PY[2] * PY[2] + PY[3] * PY[3];
      double PZZ2 = PZ[0] * PZ[0] + PZ[1] * PZ[1] + PZ[2] * PZ[2] + PZ[3] * PZ[3];
      WZZ = or_div(DT1 * (PYY2 * DYX - PXX2 * DXY), (PXX2 + PYY2));
      WXX = or_div(DT1 * (PZZ2 * DZY - PYY2 * DYZ), (PYY2 + PZZ2));
      WYY = or_div(DT1 * (PXX2 * DXZ - PZZ2 * DZX), (PZZ2 + PXX2));
      }
    } else {
      D4 = DXY + DYX; D5 = DYZ + DZY; D6 = DXZ + DZX;
      WZZ = DT1D2 * (DYX - DXY);
      WYY = DT1D2 * (DXZ - DZX);
      WXX = DT1D2 * (DZY - DYZ);
    }
    const double DIVDE = DT1 * (DXX + DYY + DZZ);        // sforc3.F:790
    // ---- SRHO3
    double VOLO = T.ld(g.w_vol);
    double RHON = T.ld(BW_RHO);
    double EINT = T.ld(BW_EINT);
    double DVOL;
    {
      const double RHON_OLD = RHON, RHO0 = m.rho0;
      if (ISMSTR == 1 && TT == K_ZERO && OFFG > K_ONE) VOLO = VOLN;     // srho3.F:128-136 (VOLO is const otherwise)
      if (OFFG == K_ZERO && VOLN == K_ONE) VOLN = VOLO;
      DVOL = VOLN - (or_div(RHO0, RHON)) * VOLO;
      RHON = RHO0 * (or_div(VOLO, VOLN));
      EINT = EINT * VOLO;
      if (ISMSTR <= 4 && OFFG > K_ONE) {
        double RHOREF = RHON;
        RHON = RHON_OLD - RHOREF * DIVDE;
        RHON = fmax(RHON, K_EM30);
        DVOL = VOLN * DIVDE;
      }
    }
    // VOL**(1/3) once for MQVISCB (AL) and SHVIS3 (VOL**(2/3) = its square): cbrt is ~40 instructions against ~180 for
    // each of the two pow calls; <= 3 ulp from the library values (the oracle's glibc pow already differs from CUDA's)
    const double CBV = cbrt(fmax(VOLN, K_ZERO));
    // ---- SROTA3
    double S1 = T.ld(BW_SIG), S2 = T.ld(BW_SIG + 1), S3 = T.ld(BW_SIG + 2), S4 = T.ld(BW_SIG + 3), S5 = T.ld(BW_SIG + 4), S6 = T.ld(BW_SIG + 5);
    double SG1, SG2, SG3, SG4, SG5, SG6;
    if (CVT) { SG1 = S1; SG2 = S2; SG3 = S3; SG4 = S4; SG5 = S5; SG6 = S6; }     // SRMALLA3: the stress lives in the co-rotating frame
    else {
      double Q1 = K_TWO * S4 * WZZ, Q2 = K_TWO * S6 * WYY, Q3 = K_TWO * S5 * WXX;
      SG1 = S1 - Q1 + Q2;
      SG2 = S2 + Q1 - Q3;
      SG3 = S3 - Q2 + Q3;
      SG4 = S4 + WZZ * (S1 - S2) + WYY * S5 - WXX * S6;
      SG5 = S5 + WXX * (S2 - S3) + WZZ * S6 - WYY * S4;
      SG6 = S6 + WYY * (S3 - S1) + WXX * S4 - WZZ * S5;
    }
    // ---- SMALLA3 (rotate the frozen reference) then S8SAV3 (refresh it)
    if (!CVT && ISMSTR <= 4 && OFFG > K_ONE) {
      #pragma unroll
      for (int k = 0; k < 7; k++) {
        double X = sm[(3 * k) * ORGPU_TILE], Y = sm[(3 * k + 1) * ORGPU_TILE], Z = sm[(3 * k + 2) * ORGPU_TILE];
        sm[(3 * k) * ORGPU_TILE] = X - Y * WZZ + Z * WYY;
        sm[(3 * k + 1) * ORGPU_TILE] = Y - Z * WXX + X * WZZ;
        sm[(3 * k + 2) * ORGPU_TILE] = Z - X * WYY + Y * WXX;
      }
    }
    // ---- MMAIN pre-law (mmain.F90:597-800)
    const double QOLD = T.ld(BW_QVIS);
    const double VOL_AVG = VOLN - K_HALF * DVOL;
    const double AMU = or_div(RHON, m.rho0) - K_ONE;
    double RHOREF;
    if (ISMSTR == 1) RHOREF = m.rho0;
    else if (ISMSTR == 2) RHOREF = (fabs(OFFG) <= K_ONE) ? RHON : or_div(m.rho0 * VOLO, fmax(K_EM20, VOLN));
    else RHOREF = RHON;
    double TEMP = K_ZERO, TSTAR = K_ZERO;
    if (m.has_temp) { TEMP = T.ld(g.w_temp); TSTAR = fmax(K_ZERO, or_div((TEMP - m.tref), fmax((m.tmelt - m.tref), K_EM20))); }
    // ---- M2LAW
    double EPXE = T.ld(BW_PLA), EPSD = T.ld(BW_EPSD);
    double SSP, QNEW, STI, SSP_EQ;
    double DPLA_F = K_ZERO, EPSP_F = K_ZERO;
    if (LAW == 2) {
      const double asrate = fmin(K_ONE, m.asrate * DT1);
      const double rhocpi = (m.rhocp > K_ZERO) ? or_div(K_ONE, m.rhocp) : K_ZERO;
      const double G = m.shear * OFF;
      double CA = m.ca, SIGMX = m.sigmx;
      const bool KIN = g.w_sigb >= 0;                         // kinematic / mixed hardening (m2law.F:181-190, 300-337, 364-390)
      if (KIN) { SG1 = SG1 - T.ld(g.w_sigb); SG2 = SG2 - T.ld(g.w_sigb + 1); SG3 = SG3 - T.ld(g.w_sigb + 2);
                 SG4 = SG4 - T.ld(g.w_sigb + 3); SG5 = SG5 - T.ld(g.w_sigb + 4); SG6 = SG6 - T.ld(g.w_sigb + 5); }
      double Pm = -K_THIRD * (SG1 + SG2 + SG3);
      double DAV = -K_THIRD * (DXX + DYY + DZZ);
      double G1 = DT1 * G, G2 = K_TWO * G1;
      SSP = or_sqrt(or_div((K_ONEP333 * G + m.bulk), m.rho0));
      SG1 = SG1 + Pm + G2 * (DXX + DAV);
      SG2 = SG2 + Pm + G2 * (DYY + DAV);
      SG3 = SG3 + Pm + G2 * (DZZ + DAV);
      SG4 = SG4 + G1 * D4;
      SG5 = SG5 + G1 * D5;
      SG6 = SG6 + G1 * D6;
      double AJ2 = K_HALF * (SG1 * SG1 + SG2 * SG2 + SG3 * SG3) + SG4 * SG4 + SG5 * SG5 + SG6 * SG6;
      AJ2 = or_sqrt(K_THREE * AJ2);
      // MSTRAIN_RATE (mstrain_rate.F:60-91)
      if (m.israte >= 0) {
        double epsdot;
        if (m.vp - 2 == 0) {
          double E4 = K_HALF * D4, E5 = K_HALF * D5, E6 = K_HALF * D6;
          double epsp = DXX * DXX + DYY * DYY + DZZ * DZZ + K_TWO * (E4 * E4 + E5 * E5 + E6 * E6);
          epsdot = or_sqrt(epsp);
        } else {
          double dav = (DXX + DYY + DZZ) * K_THIRD;
          double E1 = DXX - dav, E2 = DYY - dav, E3 = DZZ - dav, E4 = K_HALF * D4, E5 = K_HALF * D5, E6 = K_HALF * D6;
          double epsp = K_HALF * (E1 * E1 + E2 * E2 + E3 * E3) + E4 * E4 + E5 * E5 + E6 * E6;
          epsdot = or_div(or_sqrt(K_THREE * epsp), K_THREE_HALF);
        }
        if (m.israte == 0) EPSD = epsdot; else EPSD = asrate * epsdot + (K_ONE - asrate) * EPSD;
      }
      double EPD = K_ONE;
      if (m.cc != K_ZERO) {
        if (m.vp == 1) { EPD = fmax(EPSD, m.epdr); EPD = log(or_div(EPD, m.epdr)); }
        else           { EPD = fmax(EPSD, K_EM15); EPD = log(or_div(EPD, m.epdr)); }
        if (m.iform == 0) {
          double MT = fmax(K_EM15, m.z3);
          EPD = fmax(K_ZERO, EPD);
          EPD = (K_ONE + m.cc * EPD) * (K_ONE - ((TSTAR > K_ZERO) ? pow(TSTAR, MT) : K_ZERO));   // 0**MT = 0 exactly
          if (m.icc == 1) SIGMX = m.sigmx * EPD;
        } else if (m.iform == 1) {
          EPD = m.cc * exp((-m.z3 + m.z4 * EPD) * TEMP);
          if (m.icc == 1) SIGMX = m.sigmx + EPD;
          CA = m.ca + EPD;
          EPD = K_ONE;
        }
      } else if (m.iform == 0) {
        double MT = fmax(K_EM15, m.z3);
        EPD = K_ONE - ((TSTAR > K_ZERO) ? pow(TSTAR, MT) : K_ZERO);
        if (m.icc == 1) SIGMX = m.sigmx * EPD;
      }
      double AK, QH, SIGY;
      const double BETA = K_ONE - m.fisokin;                  // isotropic share of the hardening (1 without a kinematic part)
      if (m.cn == K_ONE) { SIGY = CA + m.cb * EPXE; AK = KIN ? CA + BETA * m.cb * EPXE : SIGY; QH = m.cb * EPD; }
      else if (EPXE > K_ZERO) {
        // one pow instead of two: EPXE**(CN-1) = EPXE**CN / EPXE (m2law.F:230-238; a pow is ~180 instructions, the
        // quotient differs from the library value by <= 2 ulp, far inside the 1e-12 force tolerance)
        const double PN = pow(EPXE, m.cn);
        SIGY = CA + m.cb * PN; AK = KIN ? CA + BETA * m.cb * PN : SIGY;
        if (m.cn > K_ONE) QH = (m.cb * m.cn * or_div(PN, EPXE)) * EPD;
        else              QH = (or_div(m.cb * m.cn, or_div(EPXE, PN))) * EPD;
      } else { AK = CA; SIGY = CA; QH = K_ZERO; }
      AK = AK * EPD;
      if (KIN) SIGY = SIGY * EPD;
      if (SIGMX < AK) { AK = SIGMX; QH = K_ZERO; }
      SIGY = KIN ? fmin(SIGY, SIGMX) : AK;
      if (EPXE > m.epmx) { AK = K_ZERO; QH = K_ZERO; }
      const double SE1 = SG1, SE2 = SG2, SE3 = SG3, SE4 = SG4, SE5 = SG5, SE6 = SG6;     // elastic predictors (:213-222)
      double SCALE = fmin(K_ONE, or_div(AK, fmax(AJ2, K_EM15)));
      const double DPLA = or_div((K_ONE - SCALE) * AJ2, fmax(K_THREE * G + QH, K_EM15));
      AK = AK + (K_ONE - m.fisokin) * DPLA * QH;
      SCALE = fmin(K_ONE, or_div(AK, fmax(AJ2, K_EM15)));
      SG1 = SCALE * SG1; SG2 = SCALE * SG2; SG3 = SCALE * SG3; SG4 = SCALE * SG4; SG5 = SCALE * SG5; SG6 = SCALE * SG6;
      EPXE = EPXE + DPLA;
      if (KIN) {                                              // back stress along the plastic corrector, stress gets it back
        const double HKIN = K_TWO_THIRD * m.fisokin * QH;
        const double ALPHA = or_div(HKIN, fmax(K_TWO * G + HKIN, K_EM15));
        const double B1 = T.ld(g.w_sigb) + ALPHA * (SE1 - SG1), B2 = T.ld(g.w_sigb + 1) + ALPHA * (SE2 - SG2), B3 = T.ld(g.w_sigb + 2) + ALPHA * (SE3 - SG3);
        const double B4 = T.ld(g.w_sigb + 3) + ALPHA * (SE4 - SG4), B5 = T.ld(g.w_sigb + 4) + ALPHA * (SE5 - SG5), B6 = T.ld(g.w_sigb + 5) + ALPHA * (SE6 - SG6);
        T.st(g.w_sigb, B1); T.st(g.w_sigb + 1, B2); T.st(g.w_sigb + 2, B3); T.st(g.w_sigb + 3, B4); T.st(g.w_sigb + 4, B5); T.st(g.w_sigb + 5, B6);
        SG1 = SG1 + B1; SG2 = SG2 + B2; SG3 = SG3 + B3; SG4 = SG4 + B4; SG5 = SG5 + B5; SG6 = SG6 + B6;
      }
      // ---- MQVISCB (IMPL=0, N2D=0, NPG=1, JTHE=0, IDTMINS/=2, NODADT=0)
      brick_mqviscb<ISMSTR>(g, DXX, DYY, DZZ, SSP, OFF, OFFG, VOLN, VOLO, RHON, RHOREF, CBV, DELTAX, QNEW, SSP_EQ, STI, dt_cand);
      // ---- pressure + internal energy (m2law.F:433-453)
      const double DTA = K_HALF * DT1;
      const double PNEW = m.bulk * AMU;
      SG1 = (SG1 - PNEW) * OFF; SG2 = (SG2 - PNEW) * OFF; SG3 = (SG3 - PNEW) * OFF;
      SG4 = SG4 * OFF; SG5 = SG5 * OFF; SG6 = SG6 * OFF;
      double E1 = DXX * (S1 + SG1), E2 = DYY * (S2 + SG2), E3 = DZZ * (S3 + SG3);
      double E4 = D4 * (S4 + SG4), E5 = D5 * (S5 + SG5), E6 = D6 * (S6 + SG6);
      double EINC = VOL_AVG * (E1 + E2 + E3 + E4 + E5 + E6) * DTA - K_HALF * DVOL * (QOLD + QNEW);
      EINT = or_div((EINT + EINC * OFF), fmax(K_EM15, VOLO));
      DPLA_F = DPLA; EPSP_F = EPSD;                        // what M2LAW hands the failure models (DPLA; EPSP = EPSD of m2law.F:229)
      if (m.vp == 1) { double PLAP = or_div(DPLA, fmax(K_EM20, DT1)); EPSD = asrate * PLAP + (K_ONE - asrate) * EPSD; }
      if (m.rhocp > K_ZERO) { SIGY = fmax(SIGY, AK); TEMP = TEMP + SIGY * DPLA * rhocpi; }
    } else {
      // ---- MULAW (mulaw.F90:668-700, 846-905, 1049-1052, 1133-1166, 2187-2219, 2876-2915, 3000-3016) -> SIGEPS36
      //      (sigeps36.F:170-207, 266-295, 398-603, 1453-1460, 1507-1510): VP=0, FISOKIN=0, IFAIL=0, isotropic global frame
      const orgpu_law36& m6 = g.m36;
      const double DEFP0 = EPXE;
      const double EP1 = DXX * OFF, EP2 = DYY * OFF, EP3 = DZZ * OFF, EP4 = D4 * OFF, EP5 = D5 * OFF, EP6 = D6 * OFF;
      const double DE1 = EP1 * DT1, DE2 = EP2 * DT1, DE3 = EP3 * DT1, DE4 = EP4 * DT1, DE5 = EP5 * DT1, DE6 = EP6 * DT1;
      const double SO1 = SG1, SO2 = SG2, SO3 = SG3, SO4 = SG4, SO5 = SG5, SO6 = SG6;
      if (g.w_stra >= 0) {                                 // LBUF%STRA: rotated, then incremented (ISTRAIN>0)
        const double WXXF = WXX * OFF, WYYF = WYY * OFF, WZZF = WZZ * OFF;
        const double T1 = T.ld(g.w_stra), T2 = T.ld(g.w_stra + 1), T3 = T.ld(g.w_stra + 2), T4 = T.ld(g.w_stra + 3), T5 = T.ld(g.w_stra + 4), T6 = T.ld(g.w_stra + 5);
        const double Q1 = T4 * WZZF, Q2 = T6 * WYYF, Q3 = T5 * WXXF;
        const double SS1 = T1 - Q1 + Q2, SS2 = T2 + Q1 - Q3, SS3 = T3 - Q2 + Q3;
        const double SS4 = T4 + 2. * WZZF * (T1 - T2) + WYYF * T5 - WXXF * T6;
        const double SS5 = T5 + 2. * WXXF * (T2 - T3) + WZZF * T6 - WYYF * T4;
        const double SS6 = T6 + 2. * WYYF * (T3 - T1) + WXXF * T4 - WZZF * T5;
        T.st(g.w_stra, SS1 + DE1); T.st(g.w_stra + 1, SS2 + DE2); T.st(g.w_stra + 2, SS3 + DE3);
        T.st(g.w_stra + 3, SS4 + DE4); T.st(g.w_stra + 4, SS5 + DE5); T.st(g.w_stra + 5, SS6 + DE6);
      }
      // MSTRAIN_RATE, IDEV = 1 (mstrain_rate.F:60-91)
      if (m6.israte >= 0) {
        const double dav = (EP1 + EP2 + EP3) * K_THIRD;
        const double E1 = EP1 - dav, E2 = EP2 - dav, E3 = EP3 - dav, E4 = K_HALF * EP4, E5 = K_HALF * EP5, E6 = K_HALF * EP6;
        const double epsp = K_HALF * (E1 * E1 + E2 * E2 + E3 * E3) + E4 * E4 + E5 * E5 + E6 * E6;
        const double epsdot = or_div(or_sqrt(K_THREE * epsp), K_THREE_HALF);
        if (m6.israte == 0) EPSD = epsdot;
        else { const double asrate = fmin(K_ONE, m6.asrate * DT1); EPSD = asrate * epsdot + (K_ONE - asrate) * EPSD; }
      }
      // SIGEPS36: deviatoric elastic predictor
      SSP = m6.ssp3d;
      {
        const double DAV = (DE1 + DE2 + DE3) * K_THIRD;
        const double P0 = -(SO1 + SO2 + SO3) * K_THIRD;
        SG1 = SO1 + P0 + m6.g2 * (DE1 - DAV);
        SG2 = SO2 + P0 + m6.g2 * (DE2 - DAV);
        SG3 = SO3 + P0 + m6.g2 * (DE3 - DAV);
        SG4 = SO4 + m6.shear * DE4;
        SG5 = SO5 + m6.shear * DE5;
        SG6 = SO6 + m6.shear * DE6;
      }
      // IFAIL = 2: largest principal total strain by 4 Newton steps on the deviatoric cubic, damage factor on the
      // yield stress (sigeps36.F:331-395)
      double FAIL = K_ONE, EPSTT = K_ZERO;
      if (LAW == 37) {                                     // its own kernel variant: the common LAW36 kernels carry none of it
        const double EXX = T.ld(g.w_stra), EYY = T.ld(g.w_stra + 1), EZZ = T.ld(g.w_stra + 2);
        const double EXY = T.ld(g.w_stra + 3), EYZ = T.ld(g.w_stra + 4), EZX = T.ld(g.w_stra + 5);
        const double DAV = (EXX + EYY + EZZ) * K_THIRD;
        const double E1 = EXX - DAV, E2 = EYY - DAV, E3 = EZZ - DAV, E4 = K_HALF * EXY, E5 = K_HALF * EYZ, E6 = K_HALF * EZX;
        const double E42 = E4 * E4, E52 = E5 * E5, E62 = E6 * E6;
        const double C = -K_HALF * (E1 * E1 + E2 * E2 + E3 * E3) - E42 - E52 - E62;
        double EPST = or_sqrt(-C * K_THIRD);
        const double EPSR1DAV = fmin(m6.epsr1, m6.epsf) - DAV;
        bool done = !(EPST + EPST < EPSR1DAV);
        if (done) {
          const double D = -E1 * E2 * E3 + E1 * E52 + E2 * E62 + E3 * E42 - K_TWO * E4 * E5 * E6;
          double EPST2 = EPST * EPST;
          double Y = (EPST2 + C) * EPST + D;
          if (fabs(Y) > K_EM8) {
            EPST = K_ONEP75 * EPST;
            #pragma unroll 1
            for (int it = 0; it < 4; it++) {
              EPST2 = EPST * EPST; Y = (EPST2 + C) * EPST + D;
              const double YP = K_THREE * EPST2 + C;
              EPST = EPST - or_div(Y, YP);
              if (it < 3 && EPST < EPSR1DAV) { done = false; break; }
            }
          }
          if (done) {
            EPST = EPST + DAV;
            EPSTT = EPST;
            FAIL = fmax(K_EM20, fmin(K_ONE, or_div(m6.epsr2 - EPST, m6.epsr2 - m6.epsr1)));
          }
        }
      }
      // yield stress and hardening modulus from the tabulated curves (VINTER, forward-only cursors in VARTMP)
      double YLD, H;
      if (m6.nrate == 1) {
        int ipos = T.ldi(g.w_vt, 0); const int ipos_old = ipos;
        const int f = m6.ifunc[0];
        double dydx, y1;
        if (g.ct.n > 0) vinter1c(g.ct, 0, ipos, EPXE, dydx, y1);
        else { const int i0 = __ldg(g.npf + f), i1 = __ldg(g.npf + f + 1); vinter1(g.tf, i0, i1 - i0, ipos, EPXE, dydx, y1); }
        if (ipos != ipos_old) T.sti(g.w_vt, 0, ipos);
        const double FACT = FAIL * K_ONE * (m6.yfac[0] * K_ONE);
        H = dydx * FACT;
        YLD = y1 * FACT;
      } else {
        int JJ = 1;
        for (int J = 2; J <= m6.nrate - 1; J++) if (EPSD >= m6.rate[J - 1]) JJ = J;
        double RFAC;
        if (m6.ismooth == 2) {
          const double EPSP1 = fmax(m6.rate[JJ - 1], K_EM20), EPSP2 = m6.rate[JJ];
          RFAC = or_div(log(or_div(fmax(EPSD, K_EM20), EPSP1)), log(or_div(EPSP2, EPSP1)));
        } else {
          const double EPSP1 = m6.rate[JJ - 1], EPSP2 = m6.rate[JJ];
          RFAC = or_div((EPSD - EPSP1), (EPSP2 - EPSP1));
        }
        const double YFAC1 = m6.yfac[JJ - 1] * K_ONE, YFAC2 = m6.yfac[JJ] * K_ONE;
        const int f1 = m6.ifunc[JJ - 1], f2 = m6.ifunc[JJ];
        int ipos1 = T.ldi(g.w_vt, 1 + JJ), ipos2 = T.ldi(g.w_vt, 2 + JJ);
        double dydx1, y1, dydx2, y2;
        if (g.ct.n > 0) { vinter1c(g.ct, JJ - 1, ipos1, EPXE, dydx1, y1); vinter1c(g.ct, JJ, ipos2, EPXE, dydx2, y2); }
        else {
          { const int i0 = __ldg(g.npf + f1), i1 = __ldg(g.npf + f1 + 1); vinter1(g.tf, i0, i1 - i0, ipos1, EPXE, dydx1, y1); }
          { const int i0 = __ldg(g.npf + f2), i1 = __ldg(g.npf + f2 + 1); vinter1(g.tf, i0, i1 - i0, ipos2, EPXE, dydx2, y2); }
        }
        T.sti(g.w_vt, 1 + JJ, ipos1); T.sti(g.w_vt, 2 + JJ, ipos2);
        y1 = y1 * YFAC1; y2 = y2 * YFAC2;
        YLD = (y1 + RFAC * (y2 - y1)) * (FAIL * K_ONE);
        dydx1 = dydx1 * YFAC1; dydx2 = dydx2 * YFAC2;
        H = (dydx1 + RFAC * (dydx2 - dydx1)) * (FAIL * K_ONE);
      }
      if (m6.yldcheck == 1) YLD = fmax(YLD, K_EM20);
      // projection on the yield surface (radial return), IPLA = 0 / 2 / 1
      {
        double VM = K_THREE * (K_HALF * (SG1 * SG1 + SG2 * SG2 + SG3 * SG3) + SG4 * SG4 + SG5 * SG5 + SG6 * SG6);
        if (VM > YLD * YLD) {
          VM = or_sqrt(VM);
          double R = or_div(YLD, fmax(VM, K_EM20));
          if (g.prop.ipla == 0) {
            EPXE = EPXE + or_div((K_ONE - R) * VM, fmax(m6.g3 + H, K_EM20));
          } else if (g.prop.ipla == 2) {
            EPXE = EPXE + or_div((K_ONE - R) * VM, fmax(m6.g3, K_EM20));
          } else {
            const double DPLA = or_div((K_ONE - R) * VM, fmax(m6.g3 + H, K_EM20));
            YLD = fmax(YLD + (K_ONE - m6.fisokin) * DPLA * H, K_ZERO);
            R = fmin(K_ONE, or_div(YLD, fmax(VM, K_EM20)));
            EPXE = EPXE + DPLA;
          }
          SG1 = SG1 * R; SG2 = SG2 * R; SG3 = SG3 * R; SG4 = SG4 * R; SG5 = SG5 * R; SG6 = SG6 * R;
        }
      }
      { const double Pn = m6.bulk * AMU; SG1 = SG1 - Pn; SG2 = SG2 - Pn; SG3 = SG3 - Pn; }   // IEOS = 0
      if (OFF < K_EM01) OFF = K_ZERO;
      if (OFF < K_ONE) OFF = OFF * K_FOUR_OVER_5;
      if (LAW != 37 && m6.ifail == 1) { if (EPXE > m6.epsmax && OFF == K_ONE) OFF = K_FOUR_OVER_5; }    // sigeps36.F:1546-1555
      else if (LAW == 37) { if ((EPXE > m6.epsmax || EPSTT > m6.epsf) && OFF == K_ONE) OFF = K_FOUR_OVER_5; }   // :1524-1533
      // MULAW: plastic work (L_PLA>0, von Mises of the old and new stresses)
      {
        const double DPLA = EPXE - DEFP0;
        const double VM0 = or_sqrt(K_HALF * ((SO1 - SO2) * (SO1 - SO2) + (SO2 - SO3) * (SO2 - SO3) + (SO3 - SO1) * (SO3 - SO1))
                                   + K_THREE * (SO4 * SO4 + SO5 * SO5 + SO6 * SO6));
        const double VMN = or_sqrt(K_HALF * ((SG1 - SG2) * (SG1 - SG2) + (SG2 - SG3) * (SG2 - SG3) + (SG3 - SG1) * (SG3 - SG1))
                                   + K_THREE * (SG4 * SG4 + SG5 * SG5 + SG6 * SG6));
        T.st(g.w_wpla, T.ld(g.w_wpla) + K_HALF * (VM0 + VMN) * DPLA * VOLN);
      }
      SG1 = SG1 * OFF; SG2 = SG2 * OFF; SG3 = SG3 * OFF; SG4 = SG4 * OFF; SG5 = SG5 * OFF; SG6 = SG6 * OFF;
      if (SSP == K_ZERO) SSP = or_sqrt(or_div(m6.bulk, m6.rho0));
      brick_mqviscb<ISMSTR>(g, DXX, DYY, DZZ, SSP, OFF, OFFG, VOLN, VOLO, RHON, RHOREF, CBV, DELTAX, QNEW, SSP_EQ, STI, dt_cand);
      // internal energy (mulaw.F90:3000-3013), then energy -> energy density (mmain.F90:1996-2004)
      {
        const double P2 = -(S1 + SG1 + S2 + SG2 + S3 + SG3) * K_THIRD;
        const double E1 = DXX * (S1 + SG1 + P2 + K_TWO * K_ZERO), E2 = DYY * (S2 + SG2 + P2 + K_TWO * K_ZERO), E3 = DZZ * (S3 + SG3 + P2 + K_TWO * K_ZERO);
        const double E4 = D4 * (S4 + SG4 + K_TWO * K_ZERO), E5 = D5 * (S5 + SG5 + K_TWO * K_ZERO), E6 = D6 * (S6 + SG6 + K_TWO * K_ZERO);
        const double EINC = OFF * (VOL_AVG * DT1 * (E1 + E2 + E3 + E4 + E5 + E6 + K_ZERO) - DVOL * (QNEW + QOLD + P2)) * K_HALF;
        EINT = EINT + EINC;
        EINT = (VOLO > K_ZERO) ? or_div(EINT, fmax(VOLO, K_EM20)) : K_ZERO;
      }
    }
    // mmain.F90 tail: entropy heating of the artificial viscosity when the buffer tracks temperature
    if (m.has_temp) {
      double cv = or_div(m.rhocp, m.rho0);
      if (cv > K_ZERO && OFF == K_ONE) {
        double mcv = RHON * VOLN * cv;
        double qheat = -K_HALF * (QOLD + QNEW) * DVOL;
        TEMP = TEMP + or_div(qheat, mcv);
        TEMP = fmax(K_ZERO, TEMP);
      }
      T.st(g.w_temp, TEMP);
    }
    if (LAW == 2 && g.w_dfmax >= 0) {
      // /FAIL/JOHNSON behind MMAIN (mmain.F90:2250-2262, 2288-2300, 2410-2416 -> fail_johnson.F:95-141, Ifail_so = 1; EPSP is the
      // strain rate M2LAW hands back, m2law.F:229): elements on their way out relax first, then the damage of the living ones
      // grows by DPLA / eps_f; the energy takes the (zero) stress correction of :2788-2802 -- a multiply and a divide by the volume
      if (OFF < (double)0.1f) OFF = K_ZERO;                 // 0.1 and 0.8 are REAL*4 literals (fail_johnson.F:103-104)
      if (OFF < K_ONE) OFF = OFF * (double)0.8f;
      double DFMAX = T.ld(g.w_dfmax);
      if (OFF == K_ONE) {
        if (DPLA_F != K_ZERO) {
          const double PR = K_THIRD * (SG1 + SG2 + SG3);
          const double SXX = SG1 - PR, SYY = SG2 - PR, SZZ = SG3 - PR;
          double SVM = K_HALF * (SXX * SXX + SYY * SYY + SZZ * SZZ) + SG4 * SG4 + SG6 * SG6 + SG5 * SG5;
          SVM = or_sqrt(K_THREE * SVM);
          double EPSF = or_div(g.fail.d3 * PR, fmax(K_EM20, SVM));
          EPSF = g.fail.d1 + g.fail.d2 * exp(EPSF);
          if (g.fail.d4 != K_ZERO) EPSF = EPSF * (K_ONE + g.fail.d4 * log(fmax(K_ONE, or_div(EPSP_F, g.fail.epsp0))));
          EPSF = fmax(EPSF, g.fail.epsf_min);
          if (EPSF > K_ZERO) DFMAX = DFMAX + or_div(DPLA_F, EPSF);
          DFMAX = fmin(K_ONE, DFMAX);
        }
        if (DFMAX >= K_ONE && OFF == K_ONE) OFF = K_FOUR_OVER_5;
      }
      T.st(g.w_dfmax, DFMAX);
      EINT = or_div(EINT * VOLO, fmax(VOLO, K_EM20));
    }
    // ---- state write-back
    T.st(BW_SIG, SG1); T.st(BW_SIG + 1, SG2); T.st(BW_SIG + 2, SG3); T.st(BW_SIG + 3, SG4); T.st(BW_SIG + 4, SG5); T.st(BW_SIG + 5, SG6);
    T.st(BW_EINT, EINT); T.st(BW_RHO, RHON); T.st(BW_QVIS, QNEW); T.st(BW_PLA, EPXE); T.st(BW_EPSD, EPSD);
    // ---- SMALLB3
    if (ISMSTR == 1 || ISMSTR == 3) { if (OFFG > K_ZERO) OFFG = K_TWO; }
    if (OFF < K_ONE) {
      if (OFF == K_ZERO) OFFG = K_ZERO;
      else if (OFFG > K_ONE) OFFG = K_ONE + OFF;
      else OFFG = OFF;
    }
    T.st(BW_OFF, OFFG);
    // ---- SBILAN (sbilan.F:110-157, sforc3.F:1436) on a print cycle: EI = EINT*VOL, mean of the squared nodal velocities
    if (g.bal && P.cs->ipri) {
      double bx[8], by[8], bz[8];
      #pragma unroll
      for (int k = 0; k < 8; k++) { const double4 v = ldg4(P.nd.vel + nc[k]); bx[k] = v.x; by[k] = v.y; bz[k] = v.z; }
      if (dying_in) {
        #pragma unroll
        for (int k = 0; k < 8; k++) { bx[k] = K_ZERO; by[k] = K_ZERO; bz[k] = K_ZERO; }
      }
      double vxa = bx[0] + bx[1] + bx[2] + bx[3] + bx[4] + bx[5] + bx[6] + bx[7];
      double vya = by[0] + by[1] + by[2] + by[3] + by[4] + by[5] + by[6] + by[7];
      double vza = bz[0] + bz[1] + bz[2] + bz[3] + bz[4] + bz[5] + bz[6] + bz[7];
      double va2 = bx[0] * bx[0] + bx[1] * bx[1] + bx[2] * bx[2] + bx[3] * bx[3] + bx[4] * bx[4] + bx[5] * bx[5] + bx[6] * bx[6] + bx[7] * bx[7]
                 + by[0] * by[0] + by[1] * by[1] + by[2] * by[2] + by[3] * by[3] + by[4] * by[4] + by[5] * by[5] + by[6] * by[6] + by[7] * by[7]
                 + bz[0] * bz[0] + bz[1] * bz[1] + bz[2] * bz[2] + bz[3] * bz[3] + bz[4] * bz[4] + bz[5] * bz[5] + bz[6] * bz[6] + bz[7] * bz[7];
      vxa = vxa * K_ONE_OVER_8; vya = vya * K_ONE_OVER_8; vza = vza * K_ONE_OVER_8; va2 = va2 * K_ONE_OVER_8;
      const double xmas = RHON * VOLN;
      double* b = g.bal + e; const size_t ld = g.bal_ld;
      b[0] = EINT * VOLO; b[ld] = xmas * va2 * K_HALF; b[2 * ld] = xmas * vxa; b[3 * ld] = xmas * vya; b[4 * ld] = xmas * vza;
      b[5 * ld] = (OFFG >= K_ONE) ? xmas : K_ZERO;
    }
    // ---- SHVIS3
    double F1[8], F2[8], F3[8];
    {
      const double CAQ = K_FOURTH * OFF * g.prop.hcoef;
      double FCL, FCQ;
      const double V23 = CBV * CBV;                           // VOL**TWO_THIRD (shvis3.F:203-237)
      if (ISMSTR == 1) FCL = CAQ * m.rho0 * V23;
      else if (ISMSTR == 2 && OFFG > K_ONE) { double AA = or_div(m.rho0 * VOLO, fmax(K_EM20, VOLN)); FCL = CAQ * AA * V23; }
      else FCL = CAQ * RHON * V23;
      FCQ = FCL * CAQ * K_HUNDRED;
      FCL = FCL * SSP;
      double G_[3][8];
      if (JHBE != 0) {
        BRICK_GAMMA(G_, PH)
      }
      double HX[4], HY[4], HZ[4];
      #pragma unroll
      for (int mm = 0; mm < 4; mm++) {
        HX[mm] = HGX[mm] * (FCL + fabs(HGX[mm]) * FCQ);
        HY[mm] = HGY[mm] * (FCL + fabs(HGY[mm]) * FCQ);
        HZ[mm] = HGZ[mm] * (FCL + fabs(HGZ[mm]) * FCQ);
      }
      if (JHBE == 0) {
        #define FILL0(F, H) F[0] = -H[0] - H[1] - H[2] - H[3]; F[1] =  H[0] - H[1] + H[2] + H[3]; \
                            F[2] = -H[0] + H[1] + H[2] - H[3]; F[3] =  H[0] + H[1] - H[2] + H[3]; \
                            F[4] = -H[0] + H[1] + H[2] + H[3]; F[5] =  H[0] + H[1] - H[2] - H[3]; \
                            F[6] = -H[0] - H[1] - H[2] + H[3]; F[7] =  H[0] - H[1] + H[2] - H[3];
        FILL0(F1, HX) FILL0(F2, HY) FILL0(F3, HZ)
        #undef FILL0
      } else {
        #define FILL1(F, H) F[0] = -G_[0][0] * H[0] - G_[1][0] * H[1] - G_[2][0] * H[2] - H[3]; \
                            F[1] = -G_[0][1] * H[0] - G_[1][1] * H[1] - G_[2][1] * H[2] + H[3]; \
                            F[2] = -G_[0][2] * H[0] - G_[1][2] * H[1] - G_[2][2] * H[2] - H[3]; \
                            F[3] = -G_[0][3] * H[0] - G_[1][3] * H[1] - G_[2][3] * H[2] + H[3]; \
                            F[4] = -G_[0][4] * H[0] - G_[1][4] * H[1] - G_[2][4] * H[2] + H[3]; \
                            F[5] = -G_[0][5] * H[0] - G_[1][5] * H[1] - G_[2][5] * H[2] - H[3]; \
                            F[6] = -G_[0][6] * H[0] - G_[1][6] * H[1] - G_[2][6] * H[2] + H[3]; \
                            F[7] = -G_[0][7] * H[0] - G_[1][7] * H[1] - G_[2][7] * H[2] - H[3];
        FILL1(F1, HX) FILL1(F2, HY) FILL1(F3, HZ)
        #undef FILL1
      }
    }
    // ---- SFINT3 (pairs 1-7, 2-8, 3-5, 4-6)
    {
      const double s1 = (SG1 + K_ZERO - QNEW) * VOLN, s2 = (SG2 + K_ZERO - QNEW) * VOLN, s3 = (SG3 + K_ZERO - QNEW) * VOLN;
      const double s4 = (SG4 + K_ZERO) * VOLN, s5 = (SG5 + K_ZERO) * VOLN, s6 = (SG6 + K_ZERO) * VOLN;
      const int a[4] = {0, 1, 2, 3}, b[4] = {6, 7, 4, 5};
      #pragma unroll
      for (int k = 0; k < 4; k++) {
        double FINT = s1 * PX[k] + s4 * PY[k] + s6 * PZ[k];
        F1[a[k]] = F1[a[k]] - FINT; F1[b[k]] = F1[b[k]] + FINT;
        FINT = s2 * PY[k] + s4 * PX[k] + s5 * PZ[k];
        F2[a[k]] = F2[a[k]] - FINT; F2[b[k]] = F2[b[k]] + FINT;
        FINT = s3 * PZ[k] + s6 * PX[k] + s5 * PY[k];
        F3[a[k]] = F3[a[k]] - FINT; F3[b[k]] = F3[b[k]] + FINT;
      }
    }
    if (CVT) {                                            // F = R F' (sforc3.F:1634-1645)
      #pragma unroll
      for (int k = 0; k < 8; k++) {
        const double X = R11 * F1[k] + R12 * F2[k] + R13 * F3[k], Y = R21 * F1[k] + R22 * F2[k] + R23 * F3[k], Z = R31 * F1[k] + R32 * F2[k] + R33 * F3[k];
        F1[k] = X; F2[k] = Y; F3[k] = Z;
      }
    }
    // ---- SCUMU3P
    if (OFFG < K_ZERO) {
      #pragma unroll
      for (int k = 0; k < 8; k++) { F1[k] = K_ZERO; F2[k] = K_ZERO; F3[k] = K_ZERO; }
    }
    STI = K_FOURTH * STI;
    // all eight FSKY slots in one batch of read-only loads ahead of the row stores: a load placed
    // between the stores would serialise behind them (it may alias them as far as the compiler knows)
    int sl[8];
    #pragma unroll
    for (int k = 0; k < 8; k++) sl[k] = T.ldi(g.w_slot, k);
    if (P.roww == 4) {
      #pragma unroll
      for (int k = 0; k < 8; k++) {
        const double4 r0 = make_double4(F1[k], F2[k], F3[k], STI);
        st256(reinterpret_cast<double4*>(P.fsky + (size_t)4 * sl[k]), r0);
      }
    } else {
      #pragma unroll
      for (int k = 0; k < 8; k++) {
        // mixed models (8-wide rows): write the whole row -- two full 32-byte sectors instead of four partial stores that
        // the L2 would have to merge; the moment / rotational-stiffness words of a brick corner are zero anyway
        double4* row = reinterpret_cast<double4*>(P.fsky + (size_t)8 * sl[k]);
        const double4 r0 = make_double4(F1[k], F2[k], F3[k], K_ZERO), r1 = make_double4(K_ZERO, K_ZERO, STI, K_ZERO);
        st256(row, r0); st256(row + 1, r1);
      }
    }
    if (g.xs_ftile && g.xs_ftile[tile]) {                 // frontier tile: rows to the neighbours' windows
      if (P.roww == 4) xsend_rows<4, STAGED>(P.nd.xs, T, g.w_slot, 8, P.fsky); else xsend_rows<8, STAGED>(P.nd.xs, T, g.w_slot, 8, P.fsky);
    }
  }
  cta_epilogue<true, STAGED>(dt_cand, order, P.db, g.blk0 + tile, g_tile, s_tile_dyn, (unsigned)g.nw_rw * ORGPU_TILE * 8u);
}

template <int JHBE, int ISMSTR, int LAW>
static void launch_brick_staged(const BrickParams& P, int nblk, cudaStream_t st)
{
  const size_t bytes = (size_t)P.sg.nw * ORGPU_TILE * 8;
#ifndef ORGPU_NO_STAGING
  if (bytes <= ORGPU_STAGE_MAX_BYTES) {
    stage_attr((const void*)brick_forces_kernel<JHBE, ISMSTR, LAW, true>, bytes, ORGPU_BRICK_MINB);
    brick_forces_kernel<JHBE, ISMSTR, LAW, true><<<nblk, ORGPU_BLOCK, bytes, st>>>(P);
    return;
  }
#endif
  brick_forces_kernel<JHBE, ISMSTR, LAW, false><<<nblk, ORGPU_BLOCK, 0, st>>>(P);
}

template <int JHBE, int LAW>
static void launch_brick_ismstr(const BrickParams& P, int ismstr, int nblk, cudaStream_t st)
{
  switch (ismstr) {
    case 1: launch_brick_staged<JHBE, 1, LAW>(P, nblk, st); break;
    case 2: launch_brick_staged<JHBE, 2, LAW>(P, nblk, st); break;
    default: launch_brick_staged<JHBE, 4, LAW>(P, nblk, st); break;
  }
}

template <int LAW>
static void launch_brick_jhbe(const BrickParams& P, int nblk, cudaStream_t st)
{
  if (LAW != 37 && P.sg.prop.jcvt != 0) {                  // Belytschko's co-rotational frame (Iframe = 2): compiled for Isolid 1
    launch_brick_ismstr<11, LAW>(P, P.sg.prop.ismstr, nblk, st);
    return;
  }
  switch (P.sg.prop.jhbe) {
    case 0: launch_brick_ismstr<0, LAW>(P, P.sg.prop.ismstr, nblk, st); break;
    case 2: launch_brick_ismstr<2, LAW>(P, P.sg.prop.ismstr, nblk, st); break;
    default: launch_brick_ismstr<1, LAW>(P, P.sg.prop.ismstr, nblk, st); break;
  }
}

void launch_brick_forces(const BrickSG& sg, const DevNodes& nd, double* fsky, int roww,
                         CycleState* cs, const DtBlocks& db, const FinalizeArgs& fa, cudaStream_t st)
{
  BrickParams P{sg, nd, fsky, roww, cs, db, nullptr, nullptr};
  const int nblk = sg.ne_pad / ORGPU_TILE;
  if (sg.law == 36 && sg.m36.ifail == 2) launch_brick_jhbe<37>(P, nblk, st);
  else if (sg.law == 36) launch_brick_jhbe<36>(P, nblk, st);
  else              launch_brick_jhbe<2>(P, nblk, st);
}

// batched launch (see shell_kernel.cuh): the two common brick variants (Isolid 1, Ismstr 4, staged tile; LAW2 / LAW36)
enum { BRV_LAW2 = 0, BRV_LAW36, BRV_COUNT };
static inline int brick_tab_variant(const BrickSG& d)
{
  if (d.prop.jhbe != 1 || d.prop.ismstr != 4 || d.prop.jcvt != 0 || (size_t)d.nw * ORGPU_TILE * 8 > ORGPU_STAGE_MAX_BYTES) return -1;
  if (d.law == 36) return d.m36.ifail == 2 ? -1 : BRV_LAW36;
  return BRV_LAW2;
}
void launch_brick_forces_tab(int variant, const BrickSG* d_tab, const int2* d_map, int nblk, size_t bytes, const DevNodes& nd,
                             double* fsky, int roww, CycleState* cs, const DtBlocks& db, cudaStream_t st)
{
  BrickParams P; memset(&P.sg, 0, sizeof P.sg); P.nd = nd; P.fsky = fsky; P.roww = roww; P.cs = cs; P.db = db; P.sgtab = d_tab; P.cta_map = d_map;
  if (variant == BRV_LAW36) { stage_attr((const void*)brick_forces_kernel<1, 4, 36, true, true>, bytes, ORGPU_BRICK_MINB);
                              brick_forces_kernel<1, 4, 36, true, true><<<nblk, ORGPU_BLOCK, bytes, st>>>(P); }
  else { stage_attr((const void*)brick_forces_kernel<1, 4, 2, true, true>, bytes, ORGPU_BRICK_MINB);
         brick_forces_kernel<1, 4, 2, true, true><<<nblk, ORGPU_BLOCK, bytes, st>>>(P); }
}
