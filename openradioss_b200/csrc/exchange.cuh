// exchange.cuh -- domain exchange of the /PARITH/ON skyline and the global time-step arg-min.
//
// Reference: SPMD_EXCH2_A_PON (engine/source/mpi/forces/spmd_exch2_a_pon.F:545-557 pack
// FSKY(:,ISENDP(j)), :1156 MPI_ISEND, :1165 MPI_WAITANY, :1190-1201 unpack into FSKY(:,IRECVP(j)))
// and SPMD_GLOB_MIN5 (engine/source/mpi/generic/spmd_glob_min5.F:34-128, custom op GLOB_MIN :122).
// Here: one process per GPU; rows are packed by a kernel straight from the device skyline, all
// neighbour sends/receives plus the 32-byte (dt, type, id) candidates of every rank travel in ONE
// NCCL group (one fused NCCL kernel over NVLink), and one kernel scatters the received rows into
// their reserved slots and folds the candidates in rank order + advances the RESOL time-step
// bookkeeping.  NCCL is loaded with dlopen so single-GPU users (and the CPU symbol test) do not
// need it at all.
#pragma once
#include <dlfcn.h>
#include <nccl.h>
#include <vector>
#include "common.cuh"

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi* nccl_api()
{
  static NcclApi api; static bool tried = false;
  if (tried) return api.lib ? &api : nullptr;
  tried = true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) { api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.lib) break; }
  if (!api.lib) { orgpu_set_error("NCCL not found (dlopen libnccl.so.2): %s", dlerror()); return nullptr; }
#define LOAD(f) api.f = (decltype(api.f))dlsym(api.lib, "nccl" #f); if (!api.f) { orgpu_set_error("NCCL symbol nccl" #f " missing"); api.lib = nullptr; return nullptr; }
  LOAD(GetUniqueId) LOAD(CommInitRank) LOAD(CommDestroy) LOAD(Send) LOAD(Recv) LOAD(GroupStart) LOAD(GroupEnd) LOAD(GetErrorString)
#undef LOAD
  return &api;
}

struct Exchange {
  int nranks = 1, rank = 0;
  ncclComm_t comm = nullptr;
  // neighbour lists (device): rows sendbuf[j] <- fsky[send_slots[j]], fsky[recv_slots[j]] <- recvbuf[j]
  std::vector<int> nb_rank, send_ptr, recv_ptr;       // host copies: [nneigh], [nneigh+1]
  int* d_send_slots = nullptr; int* d_recv_slots = nullptr;
  double* d_sendbuf = nullptr; double* d_recvbuf = nullptr;
  int nsend = 0, nrecv = 0;
  double* d_cand_send = nullptr;                       // (dt2t, ityptst, neltst, 0)
  double* d_cand_recv = nullptr;                       // 4 doubles per rank
  // scratch for the host-staged entry points
  int* d_slots_tmp = nullptr; double* d_rows_tmp = nullptr; size_t tmp_cap = 0;
  // ---- peer-memory exchange (one-sided pushes over NVLink into the neighbours' windows; see below)
  unsigned char* win = nullptr; size_t win_bytes = 0;   // this rank's receive window (cudaMalloc, exported by CUDA IPC)
  std::vector<unsigned char*> peer;                     // every rank's window mapped here (peer[rank] = win)
  bool p2p = false;
  int* d_send_nb = nullptr; int* d_nb_sendptr = nullptr;
  double** d_nb_rows = nullptr;                          // [nneigh][2] destination of neighbour k's rows, per parity
  double** d_peer_cand = nullptr;                        // [nranks] candidate area of every window
  unsigned long long** d_peer_flag = nullptr;            // [nranks] &flags[rank] inside every window
  unsigned long long* d_xcycle = nullptr; unsigned int* d_done = nullptr; int* d_err = nullptr;
  unsigned long long timeout_ns = 30000000000ull;        // peer wait budget (orgpu_set_exchange_timeout / ORGPU_P2P_TIMEOUT_S)
  unsigned char** d_peer_win = nullptr;                  // [nranks] window bases (for the /DT/NODA candidate exchange)
  int2* d_xsend = nullptr;                               // [lsky] inline-send table of the force kernels (common.cuh XSend)
  // /PARITH/OFF (SPMD_EXCH_A): frontier nodes shared with each neighbour (same order on both sides), 8 doubles per node
  bool parith_off = false; std::vector<int> xn_ptr; int* d_xn_nodes = nullptr; double* d_xn_send = nullptr; double* d_xn_recv = nullptr;
};

// ---- peer-memory exchange --------------------------------------------------------------------------
// Replaces pack -> NCCL send/recv -> unpack by two kernels of ours and no library call, so the multi-GPU
// cycle is one CUDA graph like the single-GPU one.  Every rank owns a receive WINDOW in its HBM:
//   int   hdr[256]            [0]=nranks [1]=nrecv [2+q]=first row of the rows rank q sends here (-1: not a neighbour)
//   u64   flags[nranks]       flags[q] = last exchange cycle rank q has completely pushed into this window
//   u64   flags2[nranks]      second flag set (at +512 B): nodal time-step candidates of /DT/NODA, published after the assembly
//   f64   cand[2][nranks][4]  (dt2t, ityptst, neltst, -) of rank q for cycle parity p ; then cand2[2][nranks][4] likewise
//   f64   rows[2][nrecv][roww] received corner rows for cycle parity p
// p2p_push_kernel copies this rank's send rows straight from its skyline into the neighbours' windows with
// 256-bit stores over NVLink (peer pointers from cudaIpcOpenMemHandle), and its last CTA then publishes the dt
// candidate to every rank and releases flags[rank] = cycle with system scope.  p2p_wait_unpack_kernel
// acquires all flags >= cycle, scatters the received rows into their reserved FSKY slots and folds the
// candidates in rank order (GLOB_MIN).  Two parities: a neighbour may run one exchange ahead, never two
// (its next push comes after its own wait on OUR flag of the cycle in between).
__device__ __forceinline__ size_t win_rows_off_dev(int nranks) { return 2048 + (((size_t)4 * nranks * 32 + 255) / 256) * 256; }
__device__ __forceinline__ double4 ld256_cg(const double4* p) {      // L2 only: written by a peer GPU during this kernel
  double4 v; asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p) : "memory"); return v; }
#define ORGPU_WIN_FLAGS 1024
#define ORGPU_WIN_CAND 2048
#define ORGPU_WIN_FLAGS2 (ORGPU_WIN_FLAGS + 512)
static inline size_t win_cand2_off(int nranks) { return ORGPU_WIN_CAND + (size_t)2 * nranks * 32; }
static inline size_t win_rows_off(int nranks) { return ORGPU_WIN_CAND + (((size_t)4 * nranks * 32 + 255) / 256) * 256; }

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v; asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }

__device__ __forceinline__ unsigned long long global_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
// (dt, key) candidates of the ranks fold with the total order the single-domain fold uses (common.cuh cand_better): smaller dt;
// on a tie a solid replaces a shell ("<=" after "<"), among shells the earlier, among solids the later element of the
// undecomposed model's processing order.  key = order + 2^32 for a solid.
#define ORGPU_KEY_SOLID 4294967296.0
__device__ __forceinline__ bool xcand_better(double da, double ka, double db, double kb) {
  if (da < db) return true;
  if (da > db) return false;
  const bool sa = ka >= ORGPU_KEY_SOLID, sb = kb >= ORGPU_KEY_SOLID;
  if (sa != sb) return sa;
  return sa ? (ka > kb) : (ka < kb);
}
__device__ __forceinline__ double cand_key(const CycleState* cs) { return (double)cs->gkey + (cs->wsolid ? ORGPU_KEY_SOLID : 0.0); }
// fold of the ranks' candidates in rank order + the RESOL bookkeeping (resol.F:2721-2722, 6124-6128, 6327, 6352, 6494-6497, 8599-8608)
__device__ __forceinline__ void fold_candidates_and_advance(CycleState* cs, const double* cand, int nranks, bool cg)
{
  double cur = K_EP06, key = 1.0e300; int typ = 0, ngl = 0;
  for (int r = 0; r < nranks; r++) {
    const double d = cg ? __ldcg(cand + 4 * r) : cand[4 * r], k = cg ? __ldcg(cand + 4 * r + 3) : cand[4 * r + 3];
    if (xcand_better(d, k, cur, key) && d < K_EP06) {
      cur = d; key = k;
      const double t1 = cg ? __ldcg(cand + 4 * r + 1) : cand[4 * r + 1], t2 = cg ? __ldcg(cand + 4 * r + 2) : cand[4 * r + 2];
      typ = (int)t1; ngl = (int)t2;
    }
  }
  cs->dt2t = cur; cs->ityptst = typ; cs->neltst = ngl;
  const double dt1 = cs->dt2;
  double dt2 = K_EP06;
  if (cur < dt2) dt2 = cur;
  const double c11 = (double)1.1f;
  dt2 = fmin(dt2, fmin(c11 * cs->dt2old, cs->dtmx));
  cs->dt2old = dt2;
  cs->dt12 = K_HALF * (dt1 + dt2);
  cs->dt1 = dt1; cs->dt2 = dt2;
  cs->tt = cs->tt + dt2; cs->ncycle += 1;
}
// bounded wait for every rank's flag to reach cycle c; a timeout is FATAL for the handle: *err and cs->abort are sticky, every
// later force / node / exchange kernel returns at once, and the host sees -8 at its next query
// (called by the first warp of a CTA: lane q polls rank q's flag, so the ranks are awaited side by side; <= 32 ranks)
__device__ __forceinline__ bool wait_flags(const unsigned long long* flags, int nranks, unsigned long long c, unsigned long long budget_ns,
                                           int* err, CycleState* cs)
{
  const int q = threadIdx.x & 31;
  bool ok = true;
  if (q < nranks) {
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(flags + q) < c) {
      if (global_ns() - t0 > budget_ns) { *err = 1; cs->abort = 1; __threadfence(); ok = false; break; }
      __nanosleep(64);
    }
  }
  return __all_sync(0xffffffffu, ok);
}

template <int ROWW>
__global__ void __launch_bounds__(256)
p2p_push_kernel(const double* __restrict__ fsky, const int* __restrict__ send_slots, const int* __restrict__ send_nb,
                const int* __restrict__ nb_sendptr, double* const* __restrict__ nb_rows, int nsend,
                const CycleState* cs, double* const* __restrict__ peer_cand, unsigned long long* const* __restrict__ peer_flag,
                int nranks, int rank, unsigned long long* xcycle, unsigned int* done)
{
  if (cs->abort) return;
  const unsigned long long c = *reinterpret_cast<volatile unsigned long long*>(xcycle) + 1ull;   // the exchange being pushed
  const int par = (int)(c & 1ull);
  constexpr int V = ROWW / 4;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nsend * V) {
    const int j = i / V, cc = i - j * V, nb = send_nb[j];
    double* dst = nb_rows[2 * nb + par] + (size_t)(j - nb_sendptr[nb]) * ROWW + 4 * cc;
    st256(reinterpret_cast<double4*>(dst), ld256(reinterpret_cast<const double4*>(fsky + (size_t)send_slots[j] * ROWW + 4 * cc)));
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(done, 1u);
    if (prev == gridDim.x - 1) {                      // last CTA: every row of this rank is on its way
      *done = 0u;
      const double d0 = cs->dt2t, d1 = (double)cs->ityptst, d2 = (double)cs->neltst, d3 = cand_key(cs);
      for (int q = 0; q < nranks; q++) {
        double* cd = peer_cand[q] + ((size_t)par * nranks + rank) * 4;
        cd[0] = d0; cd[1] = d1; cd[2] = d2; cd[3] = d3;
      }
      __threadfence_system();
      for (int q = 0; q < nranks; q++) st_release_sys(peer_flag[q], c);
      *reinterpret_cast<volatile unsigned long long*>(xcycle) = c;
    }
  }
}

// Fused variant (default): the rows leave from inside the force kernels (XSend, common.cuh) -- each CTA stores the corner rows
// a neighbour needs into that neighbour's window right next to the local store, so the transfer overlaps the element loop and
// no pack / push kernel is left on the cycle's critical path; p2p_publish_kernel hands over the dt candidate and releases the
// flags once the local arg-min is known.  Same window layout, same waiting side.  (Measured alternatives, 2 x B200, C2 / C5:
// serial push kernel 0.5295 / 0.5591 ms per cycle; frontier tiles in a first launch + push on a side stream while the
// interior tiles run 0.5415 / 0.5751 -- two waves' tails and scattered tiles cost more than the push kernel they hide.)
// p2p_push_rows_kernel is the stand-alone row push for decompositions with several destinations per slot.
template <int ROWW>
__global__ void __launch_bounds__(256)
p2p_push_rows_kernel(const double* __restrict__ fsky, const int* __restrict__ send_slots, const int* __restrict__ send_nb,
                     const int* __restrict__ nb_sendptr, double* const* __restrict__ nb_rows, int nsend,
                     const CycleState* cs, const unsigned long long* xcycle)
{
  if (cs->abort) return;
  const unsigned long long c = *reinterpret_cast<const volatile unsigned long long*>(xcycle) + 1ull;
  const int par = (int)(c & 1ull);
  constexpr int V = ROWW / 4;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nsend * V) {
    const int j = i / V, cc = i - j * V, nb = send_nb[j];
    double* dst = nb_rows[2 * nb + par] + (size_t)(j - nb_sendptr[nb]) * ROWW + 4 * cc;
    st256(reinterpret_cast<double4*>(dst), ld256(reinterpret_cast<const double4*>(fsky + (size_t)send_slots[j] * ROWW + 4 * cc)));
  }
  __threadfence_system();
}
// one warp: lane q hands rank q this rank's (dt, type, id, key) and releases its flag
__global__ void p2p_publish_kernel(const CycleState* cs, double* const* __restrict__ peer_cand, unsigned long long* const* __restrict__ peer_flag,
                                   int nranks, int rank, unsigned long long* xcycle)
{
  if (cs->abort) return;
  const unsigned long long c = *reinterpret_cast<volatile unsigned long long*>(xcycle) + 1ull;
  const int par = (int)(c & 1ull), q = threadIdx.x;
  if (q < nranks) {
    double* cd = peer_cand[q] + ((size_t)par * nranks + rank) * 4;
    cd[0] = cs->dt2t; cd[1] = (double)cs->ityptst; cd[2] = (double)cs->neltst; cd[3] = cand_key(cs);
    __threadfence_system();
    st_release_sys(peer_flag[q], c);
  }
  __syncwarp();
  if (q == 0) *reinterpret_cast<volatile unsigned long long*>(xcycle) = c;
}

template <int ROWW>
__global__ void __launch_bounds__(256)
p2p_wait_unpack_kernel(double* __restrict__ fsky, const int* __restrict__ recv_slots, int nrecv, const unsigned char* win,
                       CycleState* cs, int nranks, const unsigned long long* xcycle, int* err, int advance, unsigned long long budget_ns)
{
  if (cs->abort) return;
  const unsigned long long c = *reinterpret_cast<const volatile unsigned long long*>(xcycle);
  const int par = (int)(c & 1ull);
  __shared__ int s_ok;
  if (threadIdx.x < 32) { const bool ok = wait_flags(reinterpret_cast<const unsigned long long*>(win + ORGPU_WIN_FLAGS), nranks, c, budget_ns, err, cs);
                          if (threadIdx.x == 0) s_ok = ok ? 1 : 0; }
  __syncthreads();
  if (!s_ok) return;                                  // a peer died: nothing is scattered, the clock does not advance
  constexpr int V = ROWW / 4;
  const double* rows = reinterpret_cast<const double*>(win + win_rows_off_dev(nranks)) + (size_t)par * nrecv * ROWW;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nrecv * V) {
    const int j = i / V, cc = i - j * V;
    const double4 v = ld256_cg(reinterpret_cast<const double4*>(rows + (size_t)j * ROWW + 4 * cc));
    st256(reinterpret_cast<double4*>(fsky + (size_t)recv_slots[j] * ROWW + 4 * cc), v);
  }
  if (i == 0 && advance)                              // GLOB_MIN over the ranks + RESOL bookkeeping
    fold_candidates_and_advance(cs, reinterpret_cast<const double*>(win + ORGPU_WIN_CAND) + (size_t)par * nranks * 4, nranks, true);
}

// rows out of the skyline into a contiguous buffer; ROWW doubles per row on both sides
template <int ROWW>
__global__ void rows_pack_kernel(const double* __restrict__ fsky, const int* __restrict__ slots, int n,
                                 double* __restrict__ buf, const CycleState* cs, double* cand)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;           // one thread per double2 of a row
  constexpr int V = ROWW / 2;
  if (i < n * V) {
    const int j = i / V, c = i - j * V;
    reinterpret_cast<double2*>(buf)[(size_t)j * V + c] = reinterpret_cast<const double2*>(fsky)[(size_t)slots[j] * V + c];
  }
  if (cand && i == 0) { cand[0] = cs->dt2t; cand[1] = (double)cs->ityptst; cand[2] = (double)cs->neltst; cand[3] = cand_key(cs); }
}

// received rows into their reserved slots; thread 0 of block 0 folds the dt candidates of all ranks in
// rank order (strict "<": the lowest rank keeps a tie) and advances the time-step bookkeeping
// (resol.F:2721-2722, 6124-6128, 6327, 6352, 6494-6497, 8599-8608)
template <int ROWW>
__global__ void rows_unpack_kernel(double* __restrict__ fsky, const int* __restrict__ slots, int n,
                                   const double* __restrict__ buf, CycleState* cs, const double* __restrict__ cand, int nranks)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  constexpr int V = ROWW / 2;
  if (i < n * V) {
    const int j = i / V, c = i - j * V;
    reinterpret_cast<double2*>(fsky)[(size_t)slots[j] * V + c] = reinterpret_cast<const double2*>(buf)[(size_t)j * V + c];
  }
  if (cand && i == 0) fold_candidates_and_advance(cs, cand, nranks, false);
}

// host-staged variants always speak 8-double rows (the reference FSKY(8,LSKY)); ROWW=4 device rows
// hold (Fx,Fy,Fz,STI) -> components 1,2,3,7
__global__ void rows_gather8_kernel(const double* __restrict__ fsky, int roww, const int* __restrict__ slots, int n, double* __restrict__ out)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x; if (j >= n) return;
  const double* r = fsky + (size_t)roww * slots[j]; double* o = out + 8 * (size_t)j;
  if (roww == 8) { for (int c = 0; c < 8; c++) o[c] = r[c]; }
  else { o[0] = r[0]; o[1] = r[1]; o[2] = r[2]; o[3] = o[4] = o[5] = 0.0; o[6] = r[3]; o[7] = 0.0; }
}
__global__ void rows_scatter8_kernel(double* __restrict__ fsky, int roww, const int* __restrict__ slots, int n, const double* __restrict__ in)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x; if (j >= n) return;
  double* r = fsky + (size_t)roww * slots[j]; const double* o = in + 8 * (size_t)j;
  if (roww == 8) { for (int c = 0; c < 8; c++) r[c] = o[c]; }
  else { r[0] = o[0]; r[1] = o[1]; r[2] = o[2]; r[3] = o[6]; }
}

// ---- /PARITH/OFF: SPMD_EXCH_A (engine/source/mpi/forces/spmd_exch_a.F) -- every domain assembles the corner rows of its own
// elements only, then the PARTIAL SUMS of the frontier nodes are exchanged: pack :153-166 (A(1:3), AR(1:3), STIFN, STIFR of the
// nodes FR_ELEM shared with one neighbour), add :517-528 (neighbours in rank order, nodes in list order).  8 doubles per node.
__global__ void nodes_pack_kernel(const DevNodes nd, const int* __restrict__ nodes, int n, double* __restrict__ buf)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x; if (j >= n) return;
  const int N = nodes[j]; double* b = buf + 8 * (size_t)j;
  b[0] = nd.A[3 * N]; b[1] = nd.A[3 * N + 1]; b[2] = nd.A[3 * N + 2];
  b[3] = nd.AR[3 * N]; b[4] = nd.AR[3 * N + 1]; b[5] = nd.AR[3 * N + 2];
  b[6] = nd.STIFN[N]; b[7] = nd.STIFR[N];
}
__global__ void nodes_add_kernel(const DevNodes nd, const int* __restrict__ nodes, int n, const double* __restrict__ buf)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x; if (j >= n) return;
  const int N = nodes[j]; const double* b = buf + 8 * (size_t)j;
  nd.A[3 * N] = nd.A[3 * N] + b[0]; nd.A[3 * N + 1] = nd.A[3 * N + 1] + b[1]; nd.A[3 * N + 2] = nd.A[3 * N + 2] + b[2];
  nd.AR[3 * N] = nd.AR[3 * N] + b[3]; nd.AR[3 * N + 1] = nd.AR[3 * N + 1] + b[4]; nd.AR[3 * N + 2] = nd.AR[3 * N + 2] + b[5];
  nd.STIFN[N] = nd.STIFN[N] + b[6]; nd.STIFR[N] = nd.STIFR[N] + b[7];
}
// the ranks' dt candidates alone (the rows of /PARITH/ON travel with them; here the node sums do)
__global__ void cand_pack_kernel(const CycleState* cs, double* cand)
{ cand[0] = cs->dt2t; cand[1] = (double)cs->ityptst; cand[2] = (double)cs->neltst; cand[3] = cand_key(cs); }
__global__ void cand_fold_kernel(CycleState* cs, const double* __restrict__ cand, int nranks)
{ if (threadIdx.x == 0 && blockIdx.x == 0) fold_candidates_and_advance(cs, cand, nranks, false); }

// /DT/NODA across domains: the nodal time step is known only after the assembly, so it travels in a second, tiny
// exchange: one thread publishes this rank's (DT2T, ITYPTST, NELTST) to every window and releases flags2[rank];
// one thread waits for all ranks, folds them in rank order (strict "<") and runs the RESOL bookkeeping.
__global__ void p2p_dt_push_kernel(const CycleState* cs, unsigned char* const* __restrict__ peer_win, int nranks, int rank,
                                   const unsigned long long* xcycle)
{
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const unsigned long long c = *reinterpret_cast<const volatile unsigned long long*>(xcycle);
  const int par = (int)(c & 1ull);
  if (cs->abort) return;
  const double d0 = cs->dt2t, d1 = (double)cs->ityptst, d2 = (double)cs->neltst, d3 = cand_key(cs);
  for (int q = 0; q < nranks; q++) {
    double* cd = reinterpret_cast<double*>(peer_win[q] + ORGPU_WIN_CAND + (size_t)2 * nranks * 32) + ((size_t)par * nranks + rank) * 4;
    cd[0] = d0; cd[1] = d1; cd[2] = d2; cd[3] = d3;
  }
  __threadfence_system();
  for (int q = 0; q < nranks; q++) st_release_sys(reinterpret_cast<unsigned long long*>(peer_win[q] + ORGPU_WIN_FLAGS2) + rank, c);
}
__global__ void p2p_dt_wait_kernel(CycleState* cs, const unsigned char* win, int nranks, const unsigned long long* xcycle, int* err,
                                   unsigned long long budget_ns)
{
  if (threadIdx.x >= 32 || blockIdx.x != 0 || cs->abort) return;
  const unsigned long long c = *reinterpret_cast<const volatile unsigned long long*>(xcycle);
  const int par = (int)(c & 1ull);
  if (!wait_flags(reinterpret_cast<const unsigned long long*>(win + ORGPU_WIN_FLAGS2), nranks, c, budget_ns, err, cs)) return;
  if (threadIdx.x == 0) fold_candidates_and_advance(cs, reinterpret_cast<const double*>(win + ORGPU_WIN_CAND + (size_t)2 * nranks * 32) + (size_t)par * nranks * 4, nranks, true);
}
