// node_kernel.cuh -- per-node kernels: deterministic /PARITH/ON gather and central-difference update.
//
//   ASSPAR4  (engine/source/assembly/asspar4.F:164-181): A(:,N) etc. += FSKY(:,K), K = ADSKY(N)..ADSKY(N+1)-1,
//            a LEFT FOLD in ascending K that starts from the value already in A (the external loads).
//   ACCELE   (accele.F:65-131)   A *= 1/MS (0 when MS<=0), AR *= 1/IN
//   BCS10    fixed-dof masks (constraints/general/bcs/bcs10.F, global-frame codes)
//   VELOCITY (velocity.F:57-89)  V += DT12*A ; A = 0
//   DEPLA    (displacement.F:91-103) D += DT2*V ; X += DT2*V
// One thread owns one node; the slot rows of a node are contiguous, so a warp reads a contiguous
// span of FSKY.  No atomics: the summation order is the Starter's slot order on any GPU count.
#pragma once
#include "common.cuh"

struct NodeAcc { double a[3], ar[3], stifn, stifr; };

// Rows of one node are folded strictly in slot order, but their LOADS are independent: a batch of
// NB rows (all of a regular node's rows: 8 brick corners / 4 shell corners) is issued before the first
// add, so one thread keeps 256 bytes in flight instead of one row (the kernel is a pure stream and
// needs ~45 KB in flight per SM to cover HBM latency).
template <int ROWW>
__device__ __forceinline__ NodeAcc node_gather(const DevNodes& nd, const double* __restrict__ fsky, int n, int iroddl, double fscale)
{
  NodeAcc r;
  const int k0 = nd.adsky[n], k1 = nd.adsky[n + 1];
  constexpr int NB = (ROWW == 4) ? 8 : 4;
  constexpr int NV = ROWW / 4;                       // 32-byte vectors per row
  const double4* base = reinterpret_cast<const double4*>(fsky);
  double4 buf[NB][NV];
  #pragma unroll
  for (int j = 0; j < NB; j++) {
    if (k0 + j < k1) {
      #pragma unroll
      for (int c = 0; c < NV; c++) buf[j][c] = ld256_cs(base + (size_t)NV * (k0 + j) + c);
    }
  }
  // the node's external load (FORCE: FCY * FINTER(IFUN, TT*FCX)): with /PARITH/OFF A starts from it (force.F90:182-312: A += AA before
  // the element loop); with /PARITH/ON it is added behind the element rows, where ASSPAR4 finds FORCE's own rows (force.F90:714-1034)
  double la[3] = {K_ZERO, K_ZERO, K_ZERO}, lar[3] = {K_ZERO, K_ZERO, K_ZERO};
  if (nd.FEXT) { la[0] = nd.FEXT[3 * n] * fscale; la[1] = nd.FEXT[3 * n + 1] * fscale; la[2] = nd.FEXT[3 * n + 2] * fscale; }
  if (nd.MEXT) { lar[0] = nd.MEXT[3 * n] * fscale; lar[1] = nd.MEXT[3 * n + 1] * fscale; lar[2] = nd.MEXT[3 * n + 2] * fscale; }
  const bool lf = nd.load_first != 0;
  #pragma unroll
  for (int c = 0; c < 3; c++) { r.a[c] = lf ? la[c] : K_ZERO; r.ar[c] = lf ? lar[c] : K_ZERO; }
  r.stifn = nd.nodadt ? K_EM20 : K_ZERO; r.stifr = r.stifn;
  for (int kb = k0; kb < k1; kb += NB) {
    if (kb != k0) {                                  // irregular node with more than NB corners
      #pragma unroll
      for (int j = 0; j < NB; j++) {
        if (kb + j < k1) {
          #pragma unroll
          for (int c = 0; c < NV; c++) buf[j][c] = ld256_cs(base + (size_t)NV * (kb + j) + c);
        }
      }
    }
    #pragma unroll
    for (int j = 0; j < NB; j++) {
      if (kb + j < k1) {
        r.a[0] = r.a[0] + buf[j][0].x; r.a[1] = r.a[1] + buf[j][0].y; r.a[2] = r.a[2] + buf[j][0].z;
        if (ROWW == 4) { r.stifn = r.stifn + buf[j][0].w; }
        else {
          r.ar[0] = r.ar[0] + buf[j][0].w; r.ar[1] = r.ar[1] + buf[j][NV - 1].x; r.ar[2] = r.ar[2] + buf[j][NV - 1].y;
          r.stifn = r.stifn + buf[j][NV - 1].z; r.stifr = r.stifr + buf[j][NV - 1].w;
        }
      }
    }
  }
  if (!lf) {
    if (nd.FEXT) { r.a[0] = r.a[0] + la[0]; r.a[1] = r.a[1] + la[1]; r.a[2] = r.a[2] + la[2]; }
    if (nd.MEXT) { r.ar[0] = r.ar[0] + lar[0]; r.ar[1] = r.ar[1] + lar[1]; r.ar[2] = r.ar[2] + lar[2]; }
  }
  (void)iroddl;
  return r;
}

// nodal fields of one node, loaded ahead of the row fold so every load of the thread is in flight at once
struct NodeIn { double ms, in; int ct, cr; double4 v, w, p; double d[3]; };

__device__ __forceinline__ NodeIn node_load(const DevNodes& nd, int n, int iroddl)
{
  NodeIn q;
  q.ms = nd.MS[n]; q.in = iroddl ? nd.IN[n] : K_ZERO;
  q.ct = nd.icodt ? nd.icodt[n] : 0; q.cr = (nd.icodt && iroddl) ? nd.icodr[n] : 0;
  q.v = ld256(nd.vel + n); q.p = ld256(nd.pos + n);
  if (iroddl) q.w = ld256(nd.rot + n);
  q.d[0] = nd.D[3 * n]; q.d[1] = nd.D[3 * n + 1]; q.d[2] = nd.D[3 * n + 2];
  return q;
}

__device__ __forceinline__ void node_update(const DevNodes& nd, int n, const NodeIn& q, NodeAcc& r, double dt12, double dt2, double tt0, int iroddl, const double* gv,
                                            int ipri = 0, double dt1 = 0.0)
{
  double dw = K_ZERO;                                // print cycles: work rate of the imposed velocities on this node
  // ACCELE
  const double ms = q.ms;
  if (ms > K_ZERO) { const double rt = or_div(K_ONE, ms); r.a[0] = r.a[0] * rt; r.a[1] = r.a[1] * rt; r.a[2] = r.a[2] * rt; }
  else { r.a[0] = K_ZERO; r.a[1] = K_ZERO; r.a[2] = K_ZERO; }
  if (iroddl) {
    const double in = q.in;
    if (in > K_ZERO) { const double rt = or_div(K_ONE, in); r.ar[0] = r.ar[0] * rt; r.ar[1] = r.ar[1] * rt; r.ar[2] = r.ar[2] * rt; }
    else { r.ar[0] = K_ZERO; r.ar[1] = K_ZERO; r.ar[2] = K_ZERO; }
  }
  // GRAVIT (loads/general/grav/gravit.F:84-160, global frame, no sensor; resol.F:7123, between ACCELE and BCS10):
  // A(N2,N1) += FCY * FINTER(IFUNC, TT*FCX) for the nodes of each load, loads in their order
  if (nd.gmask) {
    unsigned m = nd.gmask[n];
    while (m) {
      const int l = __ffs(m) - 1; m &= m - 1;
      const int d = nd.gdir[l]; const double g = gv[l];
      if (d == 0) r.a[0] = r.a[0] + g; else if (d == 1) r.a[1] = r.a[1] + g; else r.a[2] = r.a[2] + g;
    }
  }
  // BCS
  if (nd.icodt) {
    const int c = q.ct;
    if (c & 4) r.a[0] = K_ZERO; if (c & 2) r.a[1] = K_ZERO; if (c & 1) r.a[2] = K_ZERO;
    if (iroddl) { const int qq = q.cr; if (qq & 4) r.ar[0] = K_ZERO; if (qq & 2) r.ar[1] = K_ZERO; if (qq & 1) r.ar[2] = K_ZERO; }
  }
  // FIXVEL (constraints/general/impvel/fixvel.F:141-147, 362-378; imposed velocity IBFV(7)=1, global frame, no sensor):
  // the acceleration that makes VELOCITY land on FAC * f((TT + DT2/2) * FACX)
  if (nd.fv_idx) {
    const int k = nd.fv_idx[n];
    if (k >= 0) {
      const FixVelNode& f = nd.fv[k];
      const double vj[3] = {q.v.x, q.v.y, q.v.z};
      #pragma unroll
      for (int j = 0; j < 3; j++) {
        if (f.func[j] >= 0 && !(tt0 < f.tstart[j]) && !(tt0 > f.tstop[j])) {
          const double tsc = (tt0 + K_HALF * dt2) * f.facx[j];
          const int i0 = nd.ft.npf[f.func[j]];
          double yc = or_vinterdp(nd.ft.tf, i0, nd.ft.npf[f.func[j] + 1] - i0, tsc);
          yc = yc * f.fac[j];
          const double aold = r.a[j];
          r.a[j] = or_div(yc - vj[j], dt12);
          // DW = 1/4 MS (A DT12 + 2 V)(A - AOLD)   (fixvel.F:391-394)
          if (ipri) dw = dw + K_FOURTH * ms * (r.a[j] * dt12 + K_TWO * vj[j]) * (r.a[j] - aold);
        }
      }
    }
  }
  // ECRIT on a print cycle (ecrit.F:196-240, 322-336): the node's terms of ENCIN, ENROT, the momenta and the mass with
  // V(n) = V(n-1/2) + DT1/2 A, A after every kinematic condition; and the two halves of the imposed-velocity work
  if (ipri && nd.nbal) {
    const double dt05 = K_HALF * dt1;
    const double vx = q.v.x + dt05 * r.a[0], vy = q.v.y + dt05 * r.a[1], vz = q.v.z + dt05 * r.a[2];
    double* b = nd.nbal + n; const size_t ld = nd.nbal_ld;
    b[0] = (vx * vx + vy * vy + vz * vz) * K_HALF * ms;
    double er = K_ZERO;
    if (iroddl) { const double wx = q.w.x + dt05 * r.ar[0], wy = q.w.y + dt05 * r.ar[1], wz = q.w.z + dt05 * r.ar[2]; er = (wx * wx + wy * wy + wz * wz) * K_HALF * q.in; }
    b[ld] = er; b[2 * ld] = vx * ms; b[3 * ld] = vy * ms; b[4 * ld] = vz * ms; b[5 * ld] = ms;
    b[6 * ld] = dt1 * dw; b[7 * ld] = dt2 * dw;
  }
  // VELOCITY
  double4 v = q.v;
  v.x = v.x + dt12 * r.a[0]; v.y = v.y + dt12 * r.a[1]; v.z = v.z + dt12 * r.a[2];
  st256(nd.vel + n, v);
  if (iroddl) {
    double4 w = q.w;
    w.x = w.x + dt12 * r.ar[0]; w.y = w.y + dt12 * r.ar[1]; w.z = w.z + dt12 * r.ar[2];
    st256(nd.rot + n, w);
  }
  // DEPLA
  double4 p = q.p;
  double vdt = dt2 * v.x; nd.D[3 * n] = q.d[0] + vdt; p.x = p.x + vdt;
  vdt = dt2 * v.y; nd.D[3 * n + 1] = q.d[1] + vdt; p.y = p.y + vdt;
  vdt = dt2 * v.z; nd.D[3 * n + 2] = q.d[2] + vdt; p.z = p.z + vdt;
  st256(nd.pos + n, p);
}

// ASSPAR4 of the nodes [n0, n1) as (8, n) rows Fx,Fy,Fz,Mx,My,Mz,STIFN,STIFR for the host (orgpu_forces_host, engine.cu)
template <int ROWW>
__global__ void __launch_bounds__(ORGPU_NODE_BLOCK)
node_forces8_kernel(const __grid_constant__ DevNodes nd, const double* __restrict__ fsky, const CycleState* __restrict__ cs, int iroddl, int n0, int n1, double* __restrict__ f8)
{
  if (cs->abort) return;
  const int n = n0 + blockIdx.x * ORGPU_NODE_BLOCK + threadIdx.x;
  if (n >= n1) return;
  const NodeAcc r = node_gather<ROWW>(nd, fsky, n, iroddl, K_ZERO);
  double4* o = reinterpret_cast<double4*>(f8 + (size_t)8 * n);
  st256(o, make_double4(r.a[0], r.a[1], r.a[2], r.ar[0])); st256(o + 1, make_double4(r.ar[1], r.ar[2], r.stifn, r.stifr));
}

// phased mode, step 2: ASSPAR4 only (A, AR, STIFN, STIFR stored for the caller)
template <int ROWW>
__global__ void __launch_bounds__(ORGPU_NODE_BLOCK)
node_assemble_kernel(const __grid_constant__ DevNodes nd, const double* __restrict__ fsky, const CycleState* __restrict__ cs, int iroddl)
{
  if (cs->abort) return;
  const int n = blockIdx.x * ORGPU_NODE_BLOCK + threadIdx.x;
  double dtt = K_EP30, dtr = K_EP30;
  if (n < nd.n) {
    NodeAcc r = node_gather<ROWW>(nd, fsky, n, iroddl, cs->fscale);
    nd.A[3 * n] = r.a[0]; nd.A[3 * n + 1] = r.a[1]; nd.A[3 * n + 2] = r.a[2];
    nd.AR[3 * n] = r.ar[0]; nd.AR[3 * n + 1] = r.ar[1]; nd.AR[3 * n + 2] = r.ar[2];
    nd.STIFN[n] = r.stifn; nd.STIFR[n] = r.stifr;
    if (nd.nodadt) {                                        // DTNODA: dtnoda.F:221-260 (translations), :445-462 (rotations)
      const double ms = nd.MS[n];
      if (r.stifn > K_ZERO && ms > K_ZERO) dtt = nd.dtfac_node * or_sqrt(or_div(K_TWO * ms, r.stifn));
      if (iroddl) { const double in = nd.IN[n]; if (r.stifr > K_ZERO && in > K_ZERO) dtr = nd.dtfac_node * or_sqrt(or_div(K_TWO * in, r.stifr)); }
    }
  }
  if (nd.nodadt) {                                          // one candidate per CTA and family; first node in node order keeps a tie
    __shared__ double s_d[2][ORGPU_NODE_BLOCK / 32]; __shared__ int s_n[2][ORGPU_NODE_BLOCK / 32];
    double d[2] = {dtt, dtr}; int id[2] = {n, n};
    #pragma unroll
    for (int f = 0; f < 2; f++) {
      #pragma unroll
      for (int s = 16; s > 0; s >>= 1) {
        const double d2 = __shfl_down_sync(0xffffffffu, d[f], s); const int n2 = __shfl_down_sync(0xffffffffu, id[f], s);
        if (dt_better<false>(d2, n2, d[f], id[f])) { d[f] = d2; id[f] = n2; }
      }
      if ((threadIdx.x & 31) == 0) { s_d[f][threadIdx.x >> 5] = d[f]; s_n[f][threadIdx.x >> 5] = id[f]; }
    }
    __syncthreads();
    if (threadIdx.x < 2) {
      const int f = threadIdx.x; double bd = s_d[f][0]; int bn = s_n[f][0];
      for (int w = 1; w < ORGPU_NODE_BLOCK / 32; w++) if (dt_better<false>(s_d[f][w], s_n[f][w], bd, bn)) { bd = s_d[f][w]; bn = s_n[f][w]; }
      nd.nd_dt[f * gridDim.x + blockIdx.x] = bd; nd.nd_node[f * gridDim.x + blockIdx.x] = bn;
    }
  }
}

// /DT/NODA: fold the per-CTA nodal candidates into DT2T (strict "<": translations in node order, then rotations,
// against what the elements left -- nothing, they do not lower DT2T in this mode), NELTST = ITAB(N), ITYPTST = 11;
// fused = 1 also runs the RESOL bookkeeping that element_finalize_kernel skipped
__global__ void __launch_bounds__(1024)
dtnoda_finalize_kernel(CycleState* cs, const __grid_constant__ DevNodes nd, int ncta, int fused)
{
  __shared__ double s_dt[32]; __shared__ int s_ord[32];
  if (cs->abort) return;
  double cur = cs->dt2t; int curn = -1;
  for (int f = 0; f < 2; f++) {
    double dt = K_EP30; int ord = 0x7fffffff;
    for (int b = threadIdx.x; b < ncta; b += 1024) {
      const double d2 = __ldcg(nd.nd_dt + f * ncta + b); const int o2 = __ldcg(nd.nd_node + f * ncta + b);
      if (dt_better<false>(d2, o2, dt, ord)) { dt = d2; ord = o2; }
    }
    finalize_fold<false>(dt, ord, s_dt, s_ord);
    if (threadIdx.x == 0 && dt < cur) { cur = dt; curn = ord; }
  }
  if (threadIdx.x == 0) {
    if (curn >= 0) { cs->dt2t = cur; cs->neltst = nd.itab ? nd.itab[curn] : curn + 1; cs->ityptst = 11;
                     cs->gkey = nd.gnode ? nd.gnode[curn] : curn; cs->wsolid = 0; }
    if (fused) {
      const double dt1 = cs->dt2;
      double dt2 = K_EP06;
      if (cs->dt2t < dt2) dt2 = cs->dt2t;
      const double c11 = (double)1.1f;
      dt2 = fmin(dt2, fmin(c11 * cs->dt2old, cs->dtmx));
      cs->dt2old = dt2;
      cs->dt12 = K_HALF * (dt1 + dt2);
      cs->dt1 = dt1; cs->dt2 = dt2;
      cs->tt = cs->tt + dt2; cs->ncycle += 1;
    }
  }
}

// phased mode, step 3: ACCELE + BCS + VELOCITY + DEPLA from the stored A / AR
__global__ void __launch_bounds__(ORGPU_NODE_BLOCK)
node_advance_kernel(const __grid_constant__ DevNodes nd, const CycleState* __restrict__ cs, int iroddl)
{
  const int n = blockIdx.x * ORGPU_NODE_BLOCK + threadIdx.x;
  if (n >= nd.n || cs->abort) return;
  NodeAcc r;
  r.a[0] = nd.A[3 * n]; r.a[1] = nd.A[3 * n + 1]; r.a[2] = nd.A[3 * n + 2];
  r.ar[0] = nd.AR[3 * n]; r.ar[1] = nd.AR[3 * n + 1]; r.ar[2] = nd.AR[3 * n + 2];
  const NodeIn q = node_load(nd, n, iroddl);
  node_update(nd, n, q, r, cs->dt12, cs->dt2, cs->tt0, iroddl, cs->gv, cs->ipri, cs->dt1);
  nd.A[3 * n] = K_ZERO; nd.A[3 * n + 1] = K_ZERO; nd.A[3 * n + 2] = K_ZERO;        // velocity.F:62-64
  nd.AR[3 * n] = K_ZERO; nd.AR[3 * n + 1] = K_ZERO; nd.AR[3 * n + 2] = K_ZERO;
}

// device-resident loop: gather + update in one pass, A never leaves registers
template <int ROWW>
__global__ void __launch_bounds__(ORGPU_NODE_BLOCK, ORGPU_NODE_MINB)
node_fused_kernel(const __grid_constant__ DevNodes nd, const double* __restrict__ fsky,
                  const CycleState* __restrict__ cs, int iroddl)
{
  const int n = blockIdx.x * ORGPU_NODE_BLOCK + threadIdx.x;
  if (n >= nd.n || cs->abort) return;
  const NodeIn q = node_load(nd, n, iroddl);
  NodeAcc r = node_gather<ROWW>(nd, fsky, n, iroddl, cs->fscale);
  node_update(nd, n, q, r, cs->dt12, cs->dt2, cs->tt0, iroddl, cs->gv, cs->ipri, cs->dt1);
}

__global__ void set_dt_kernel(CycleState* cs, double dt1, double dt12, double dt2, int which)
{
  if (which == 0) { cs->dt2 = dt1; }                       // forces phase: element kernels read DT1 from dt2
  else { cs->dt1 = cs->dt2; cs->dt12 = dt12; cs->dt2 = dt2; cs->tt = cs->tt + dt2; cs->ncycle += 1; cs->dt2old = dt2; }
}

void launch_node_assemble(const DevNodes& nd, const double* fsky, int roww, const CycleState* cs, int iroddl, cudaStream_t st)
{
  const int nb = (nd.n + ORGPU_NODE_BLOCK - 1) / ORGPU_NODE_BLOCK;
  if (roww == 4) node_assemble_kernel<4><<<nb, ORGPU_NODE_BLOCK, 0, st>>>(nd, fsky, cs, iroddl);
  else           node_assemble_kernel<8><<<nb, ORGPU_NODE_BLOCK, 0, st>>>(nd, fsky, cs, iroddl);
}
void launch_dtnoda_finalize(const DevNodes& nd, CycleState* cs, int fused, cudaStream_t st)
{
  const int nb = (nd.n + ORGPU_NODE_BLOCK - 1) / ORGPU_NODE_BLOCK;
  dtnoda_finalize_kernel<<<1, 1024, 0, st>>>(cs, nd, nb, fused);
}
void launch_node_advance(const DevNodes& nd, const CycleState* cs, int iroddl, cudaStream_t st)
{
  const int nb = (nd.n + ORGPU_NODE_BLOCK - 1) / ORGPU_NODE_BLOCK;
  node_advance_kernel<<<nb, ORGPU_NODE_BLOCK, 0, st>>>(nd, cs, iroddl);
}
void launch_node_fused(const DevNodes& nd, const double* fsky, int roww, const CycleState* cs, int iroddl, cudaStream_t st)
{
  const int nb = (nd.n + ORGPU_NODE_BLOCK - 1) / ORGPU_NODE_BLOCK;
  if (roww == 4) node_fused_kernel<4><<<nb, ORGPU_NODE_BLOCK, 0, st>>>(nd, fsky, cs, iroddl);
  else           node_fused_kernel<8><<<nb, ORGPU_NODE_BLOCK, 0, st>>>(nd, fsky, cs, iroddl);
}
void launch_set_dt(CycleState* cs, double dt1, double dt12, double dt2, int which, cudaStream_t st)
{
  set_dt_kernel<<<1, 1, 0, st>>>(cs, dt1, dt12, dt2, which);
}

// ---- print-cycle energy balances ----------------------------------------------------------------
// Internal energy as SBILAN / CBILAN book it into PARTSAV(1,part) (sbilan.F:138-157: EI = EINT*VOL for
// solids; cbilan.F:183: EI = EINT(1)+EINT(2) for shells, deleted elements excluded :261) and the nodal
// kinetic energies.  Deterministic: fixed 256-wide CTA partial sums in a fixed tree, final sum on the host
// in block order.
__global__ void __launch_bounds__(256)
energy_partial_kernel(const double* __restrict__ a, const double* __restrict__ b, const double* __restrict__ off, int nw, int n, int mode,
                      double* __restrict__ partial)
{
  // nw > 0: a, b, off are word rows of a tile-major slab -> element i sits at ((i>>7)*nw)*128 + (i&127)
  // mode 0: sum a[i]*b[i] (off ignored) ; 1: sum (a[i] + b[i]) where off[i] != 0 ; 2: sum 0.5*a[i]*|v_i|^2 with b = double4 records
  __shared__ double s[256];
  const int i = blockIdx.x * 256 + threadIdx.x;
  double v = 0.0;
  if (i < n) {
    const size_t j = nw > 0 ? ((size_t)(i >> ORGPU_TILE_SHIFT) * nw) * ORGPU_TILE + (i & (ORGPU_TILE - 1)) : (size_t)i;
    if (mode == 0) v = a[j] * b[j];
    else if (mode == 1) v = (off[j] != 0.0) ? a[j] + b[j] : 0.0;
    else { const double4 w = reinterpret_cast<const double4*>(b)[i]; v = 0.5 * a[i] * (w.x * w.x + w.y * w.y + w.z * w.z); }
  }
  s[threadIdx.x] = v; __syncthreads();
  for (int k = 128; k > 0; k >>= 1) { if (threadIdx.x < k) s[threadIdx.x] = s[threadIdx.x] + s[threadIdx.x + k]; __syncthreads(); }
  if (threadIdx.x == 0) partial[blockIdx.x] = s[0];
}

// ---- print-cycle balances: fixed-order reduction (CBILAN / SBILAN -> PARTSAV, ECRIT) ----------------------------------
// Level 1: CTA c folds elements [start, start+n) of each of NCOMP scratch rows (<= ORGPU_BAL_CHUNK of them): every thread a
// strided left fold, then a fixed shared-memory tree.  chunks == null: node rows, chunk c = nodes [c*CHUNK, (c+1)*CHUNK).
template <int NCOMP>
__global__ void __launch_bounds__(256)
balance_chunk_kernel(const double* __restrict__ rows, int ld, int ntot, const BalChunk* __restrict__ chunks, double* __restrict__ partial)
{
  __shared__ double s[256];
  int start, n;
  if (chunks) { start = chunks[blockIdx.x].start; n = chunks[blockIdx.x].n; }
  else { start = blockIdx.x * ORGPU_BAL_CHUNK; n = min(ORGPU_BAL_CHUNK, ntot - start); }
  for (int k = 0; k < NCOMP; k++) {
    double v = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) v = v + rows[(size_t)k * ld + start + i];
    s[threadIdx.x] = v; __syncthreads();
    for (int h = 128; h > 0; h >>= 1) { if (threadIdx.x < h) s[threadIdx.x] = s[threadIdx.x] + s[threadIdx.x + h]; __syncthreads(); }
    if (threadIdx.x == 0) partial[(size_t)blockIdx.x * NCOMP + k] = s[0];
    __syncthreads();
  }
}
// Level 2 (one CTA): PARTSAV(k, part) = left fold of the chunk sums of that part in chunk order; ENINT = sum of PARTSAV(1,:);
// the node sums; the imposed-velocity work with its one-cycle-late half; one row of the history ring
__global__ void __launch_bounds__(256)
balance_finalize_kernel(const BalChunk* __restrict__ chunks, int nchunk, const double* __restrict__ epart, int npart,
                        const double* __restrict__ npartial, int nnchunk, double* __restrict__ partsav, BalState* __restrict__ bs,
                        double* __restrict__ hist, const CycleState* __restrict__ cs)
{
  __shared__ double s_node[8];
  for (int t = threadIdx.x; t < 6 * npart; t += 256) {
    const int part = t / 6, k = t % 6;
    double v = 0.0;
    for (int c = 0; c < nchunk; c++) if (chunks[c].part == part) v = v + epart[(size_t)c * 6 + k];
    partsav[t] = v;
  }
  if (threadIdx.x < 8) {
    double v = 0.0;
    for (int c = 0; c < nnchunk; c++) v = v + npartial[(size_t)c * 8 + threadIdx.x];
    s_node[threadIdx.x] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double enint = 0.0;
    for (int m = 0; m < npart; m++) enint = enint + partsav[6 * m];
    bs->wfext = bs->wfext + bs->pending + s_node[6];
    bs->pending = s_node[7];
    bs->glob[0] = s_node[0]; bs->glob[1] = s_node[1]; bs->glob[2] = enint; bs->glob[3] = bs->wfext;
    bs->glob[4] = s_node[2]; bs->glob[5] = s_node[3]; bs->glob[6] = s_node[4]; bs->glob[7] = s_node[5];
    // the cycle this row belongs to: ncycle was already advanced by the time-step bookkeeping
    const long long row = (cs->ncycle - 1) % ORGPU_BAL_HIST;
    if (row >= 0) for (int k = 0; k < 8; k++) hist[8 * row + k] = bs->glob[k];
  }
}
