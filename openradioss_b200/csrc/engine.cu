// engine.cu -- host side of liborgpu.so: the C ABI of include/orgpu.h.
//
// Mirrors what the reference's Fortran glue does around its own GPU path
// (engine/source/elements/shell/coque/shell_internal_forces.F90: FORINTC_PREPARE_GPU :370 builds
// "super-groups" of consecutive compatible groups :547-558 and flattens ELBUF into SoA :829-886;
// shell_gpu_driver.cu:94-263 owns the device memory), generalised to bricks / QEPH / LAW36 and
// with the nodal arrays, the /PARITH/ON gather and the nodal update resident on the device.
// Single translation unit (kernels are included) so no relocatable device code is needed.
#include <cstdio>
#include <cstdlib>
#include <cstdarg>
#include <cstring>
#include <cstddef>
#include <vector>
#include <string>
#include <algorithm>
#include <ctime>
#include "../../include/orgpu.h"
#include "common.cuh"
#include "brick_kernel.cuh"
#include "shell_kernel.cuh"
#include "node_kernel.cuh"
#include "exchange.cuh"

static thread_local char g_err[1024] = "";
void orgpu_set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap); }
extern "C" const char* orgpu_last_error(void) { return g_err; }

#define FAIL(code, ...) do { orgpu_set_error(__VA_ARGS__); return (code); } while (0)
#define NEED(cond, code, ...) do { if (!(cond)) FAIL(code, __VA_ARGS__); } while (0)

template <class T> static int dev_alloc(T** p, size_t n) {
  *p = nullptr; if (n == 0) return 0;
  CUDA_OK(cudaMalloc((void**)p, n * sizeof(T)));
  CUDA_OK(cudaMemset(*p, 0, n * sizeof(T)));
  return 0;
}

struct HostSolidGroup { int nel, nft, law; orgpu_law2 mat; orgpu_law36 m36; orgpu_prop_solid prop; std::vector<double> vol0; int part = 0; orgpu_fail fail{}; };

struct BrickSGHost { BrickSG d; int first_elem; int part = 0; std::vector<void*> owned; };

struct orgpu_engine {
  int device = 0, numnod = 0;
  orgpu_control ctl{};
  cudaStream_t st = nullptr;
#ifndef ORGPU_NSIDE
#define ORGPU_NSIDE 111     // with the main stream 112 concurrent launches (the device runs up to 128 kernels side by side); 15 / 47 / 111 measured
#endif
  cudaStream_t side[ORGPU_NSIDE] = {};     // side streams: super-groups are independent, their kernels may overlap
  cudaEvent_t ev_fork = nullptr, ev_join[ORGPU_NSIDE] = {};
  bool split = false;                                    // corner rows leave from inside the force kernels (XSend; set by orgpu_p2p_connect)
  DevNodes nd{};                      // device pointers
  double *d_stage3a = nullptr, *d_stage3b = nullptr, *d_stage3c = nullptr;   // (3,N) staging for pack/unpack
  double *d_fext = nullptr, *d_mext = nullptr; int *d_icodt = nullptr, *d_icodr = nullptr, *d_adsky = nullptr;
  // connectivity as given by the caller
  std::vector<int> ixs, iads, ixc, iadc, adsky; int numels = 0, numelc = 0, lsky = 0;
  std::vector<int> ixtg, iadtg; int numeltg = 0;            // 3-node shells: IXTG(6,*), IADTG(3,*)
  std::vector<HostShellGroup> tgroups;
  std::vector<int> npf; std::vector<double> tf;
  int lf_func = -1; double lf_fcx = 1.0;            // time function of the nodal loads
  // concentrated loads record by record (force.F90:188-312: every record its node, direction, time function, FCY, FCX)
  std::vector<int> cl_ib; std::vector<double> cl_fac; int n_cl_nodes = 0;
  int *d_cl_ptr = nullptr, *d_cl_nodes = nullptr; int2* d_cl_rec = nullptr; double2* d_cl_fac = nullptr;
  int ngrav = 0; int gdir[ORGPU_MAXGRAV] = {}, gfunc[ORGPU_MAXGRAV] = {}; double gfcy[ORGPU_MAXGRAV] = {}, gfcx[ORGPU_MAXGRAV] = {};   // /GRAV loads
  std::vector<unsigned int> gmask; unsigned int* d_gmask = nullptr;
  std::vector<int> fv_idx; std::vector<FixVelNode> fv;   // imposed velocities, per node
  double* d_btf = nullptr; int* d_bnpf = nullptr;   // LAW36 function table of the brick super-groups
  double* d_ftf = nullptr; int* d_fnpf = nullptr; int* d_fv_idx = nullptr; FixVelNode* d_fv = nullptr;
  std::vector<int> itab; int* d_itab = nullptr; double* d_nd_dt = nullptr; int* d_nd_node = nullptr;   // /DT/NODA
  std::vector<HostSolidGroup> sgroups;
  std::vector<HostShellGroup> cgroups;
  // device model
  std::vector<BrickSGHost> bsg;
  std::vector<ShellSGHost> csg;
  double* d_fsky = nullptr; int roww = 4;
  CycleState* d_cs = nullptr;
  DtBlocks db{}; FinalizeArgs fa{}; std::vector<SGRange> sgr; SGRange* d_sgr = nullptr;
  bool finalized = false;
  // graph of one fused cycle
  cudaGraphExec_t gexec = nullptr;
  // instrumentation
  long long launches = 0; double last_run_ms = 0; int profile = 0;
  double prof_ms[3] = {0, 0, 0}; long long prof_n[3] = {0, 0, 0};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr; bool ev_recorded = false;     // ev0 / ev1 bracket the last orgpu_run_cycles
  std::vector<cudaEvent_t> evpool;
  Exchange xc;                        // domain exchange (one process per GPU)
  std::vector<int> gord_c, gord_t, gord_s, gnode; int* d_gnode = nullptr;   // global processing order / node index of a domain's elements / nodes
  // batched launches (decks with many parts): super-groups of one kernel variant share a launch when there are enough of them
  struct Batch { bool brick; int variant; std::vector<int> sgs; int ntile = 0; size_t bytes = 0; void* d_tab = nullptr; int2* d_map = nullptr; };
  std::vector<Batch> batches; std::vector<char> sh_batched, br_batched; bool tabs_dirty = true;
  // orgpu_forces_host: the cycle of the reference's -gpu ABI as a pipeline over node chunks (built at the first call)
  struct PipeLaunch { bool brick; int variant; const void* d_tab; int2* d_map; int ntile; size_t bytes; };
  struct HostPipe {
    int K = 0, desc_gen = -1; std::vector<int> nb;               // chunk k = nodes [nb[k], nb[k+1])
    std::vector<std::vector<PipeLaunch>> grp;                    // batched launches that become ready with upload chunk j
    std::vector<std::vector<int>> solo_c, solo_b;                // super-groups launched on their own, by the chunk that completes them
    std::vector<std::vector<int>> done;                          // node chunks complete after element group j
    std::vector<void*> owned;
    cudaStream_t up = nullptr, down = nullptr, asmb = nullptr; std::vector<cudaEvent_t> ev_up, ev_el, ev_as, ev_dn; cudaEvent_t ev_start = nullptr, ev_down = nullptr;
    double* d_f8 = nullptr; CycleState* h_cs = nullptr;
  } hp;
  int desc_gen = 0;                   // bumped whenever a super-group descriptor changes (device tables must follow)
  // print-cycle balances (CBILAN / SBILAN / ECRIT): parts, GBUF%VOL of the shells, scratch rows and their fixed-order reduction
  int npart = 1; bool have_parts = false; std::vector<int> ipartc, iparts, iparttg; std::vector<double> gvolc, gvoltg;
  int ipri = 0; int bal_ld = 0, nbal_ld = 0, nchunk = 0, nnchunk = 0;
  double *d_bal = nullptr, *d_nbal = nullptr, *d_epart = nullptr, *d_npartial = nullptr, *d_partsav = nullptr, *d_hist = nullptr;
  BalChunk* d_chunks = nullptr; BalState* d_bs = nullptr;
};

// ---- small layout kernels ---------------------------------------------------------------
__global__ void pack3to4_kernel(const double* __restrict__ a3, double4* __restrict__ a4, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
  a4[i] = make_double4(a3[3 * i], a3[3 * i + 1], a3[3 * i + 2], 0.0);
}
// up to three (3,n) arrays -> 32-byte records in one launch (orgpu_forces_host: one per upload chunk)
__global__ void pack3to4x3_kernel(const double* __restrict__ a, double4* __restrict__ a4, const double* __restrict__ b, double4* __restrict__ b4,
                                  const double* __restrict__ c, double4* __restrict__ c4, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
  a4[i] = make_double4(a[3 * i], a[3 * i + 1], a[3 * i + 2], 0.0);
  b4[i] = make_double4(b[3 * i], b[3 * i + 1], b[3 * i + 2], 0.0);
  if (c) c4[i] = make_double4(c[3 * i], c[3 * i + 1], c[3 * i + 2], 0.0);
}
__global__ void unpack4to3_kernel(const double4* __restrict__ a4, double* __restrict__ a3, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
  double4 v = a4[i]; a3[3 * i] = v.x; a3[3 * i + 1] = v.y; a3[3 * i + 2] = v.z;
}

static int upload3to4(orgpu_engine* e, const double* h, double4* d4) {
  const int n = e->numnod;
  CUDA_OK(cudaMemcpyAsync(e->d_stage3a, h, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, e->st));
  pack3to4_kernel<<<(n + 255) / 256, 256, 0, e->st>>>(e->d_stage3a, d4, n); e->launches++;
  CUDA_OK(cudaStreamSynchronize(e->st));
  return 0;
}
static int download4to3(orgpu_engine* e, const double4* d4, double* h) {
  const int n = e->numnod;
  unpack4to3_kernel<<<(n + 255) / 256, 256, 0, e->st>>>(d4, e->d_stage3a, n); e->launches++;
  CUDA_OK(cudaMemcpyAsync(h, e->d_stage3a, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, e->st));
  CUDA_OK(cudaStreamSynchronize(e->st));
  return 0;
}

template <class T> static int push_dev(std::vector<void*>& owned, T** p, size_t n) { int r = dev_alloc(p, n); if (!r && *p) owned.push_back(*p); return r; }
template <class T> static int upload_vec(std::vector<void*>& owned, T** p, const std::vector<T>& h) {
  if (push_dev(owned, p, h.size())) return -100;
  if (h.size()) CUDA_OK(cudaMemcpy(*p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

extern "C" {
static int create_body(orgpu_engine* e, int numnod, const orgpu_control* ctl);

int orgpu_create(orgpu_engine** out, int device, int numnod, const orgpu_control* ctl)
{
  NEED(out && ctl && numnod > 0, -1, "orgpu_create: bad arguments");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) FAIL(-2, "orgpu_create: no CUDA device (this library has no CPU path)");
  NEED(device >= 0 && device < ndev, -3, "orgpu_create: device %d out of range (%d devices)", device, ndev);
  CUDA_OK(cudaSetDevice(device));
  orgpu_engine* e = new orgpu_engine();
  e->device = device; e->numnod = numnod; e->ctl = *ctl;
  const int rc = create_body(e, numnod, ctl);
  if (rc) { orgpu_destroy(e); return rc; }          // nothing of a half-built engine is leaked
  *out = e;
  return 0;
}

static int create_body(orgpu_engine* e, int numnod, const orgpu_control* ctl)
{
  CUDA_OK(cudaStreamCreateWithFlags(&e->st, cudaStreamNonBlocking));
  for (int k = 0; k < ORGPU_NSIDE; k++) { CUDA_OK(cudaStreamCreateWithFlags(&e->side[k], cudaStreamNonBlocking)); CUDA_OK(cudaEventCreateWithFlags(&e->ev_join[k], cudaEventDisableTiming)); }
  CUDA_OK(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
  const size_t n = numnod;
  e->nd.n = numnod;
  if (dev_alloc(&e->nd.pos, n) || dev_alloc(&e->nd.vel, n) || dev_alloc(&e->nd.D, 3 * n) || dev_alloc(&e->nd.A, 3 * n) ||
      dev_alloc(&e->nd.AR, 3 * n) || dev_alloc(&e->nd.STIFN, n) || dev_alloc(&e->nd.STIFR, n) || dev_alloc(&e->nd.MS, n) ||
      dev_alloc(&e->nd.IN, n) || dev_alloc(&e->d_stage3a, 3 * n) || dev_alloc(&e->d_stage3b, 3 * n)) return -100;
  if (ctl->iroddl) { if (dev_alloc(&e->nd.rot, n)) return -100; }
  if (dev_alloc(&e->d_cs, 1)) return -100;
  CycleState cs{}; cs.tt = ctl->tt_init; cs.dt2 = ctl->dt_init; cs.dt2old = ctl->dt2old_init; cs.dtmx = ctl->dtmx;
  cs.tt0 = ctl->tt_init; cs.fscale = 1.0;
  CUDA_OK(cudaMemcpy(e->d_cs, &cs, sizeof cs, cudaMemcpyHostToDevice));
  CUDA_OK(cudaEventCreate(&e->ev0)); CUDA_OK(cudaEventCreate(&e->ev1));
  return 0;
}

static void pipe_free(orgpu_engine* e);
int orgpu_destroy(orgpu_engine* e)
{
  if (!e) return 0;
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->st);
  if (e->gexec) cudaGraphExecDestroy(e->gexec);
  for (auto& s : e->bsg) for (void* p : s.owned) cudaFree(p);
  for (auto& s : e->csg) for (void* p : s.owned) cudaFree(p);
  void* ptrs[] = {e->nd.pos, e->nd.vel, e->nd.rot, e->nd.D, e->nd.A, e->nd.AR, e->nd.STIFN, e->nd.STIFR, e->nd.MS, e->nd.IN,
                  e->d_stage3a, e->d_stage3b, e->d_stage3c, e->d_fext, e->d_mext, e->d_icodt, e->d_icodr, e->d_adsky, e->d_fsky, e->d_cs,
                  e->db.dt, e->db.order, e->d_sgr, e->d_btf, e->d_bnpf, e->d_ftf, e->d_fnpf, e->d_fv_idx, e->d_fv, e->d_itab, e->d_nd_dt, e->d_nd_node, e->d_gmask,
                  e->d_gnode, e->d_bal, e->d_nbal, e->d_epart, e->d_npartial, e->d_partsav, e->d_hist, e->d_chunks, e->d_bs};
  for (void* p : ptrs) if (p) cudaFree(p);
  { Exchange& x = e->xc;
    void* xp[] = {x.d_send_slots, x.d_recv_slots, x.d_sendbuf, x.d_recvbuf, x.d_cand_send, x.d_cand_recv, x.d_slots_tmp, x.d_rows_tmp};
    for (void* p : xp) if (p) cudaFree(p);
    for (size_t q = 0; q < x.peer.size(); q++) if (x.peer[q] && (int)q != x.rank) cudaIpcCloseMemHandle(x.peer[q]);
    void* pp[] = {x.win, x.d_send_nb, x.d_nb_sendptr, x.d_nb_rows, x.d_peer_cand, x.d_peer_flag, x.d_xcycle, x.d_done, x.d_err, x.d_peer_win, x.d_xsend, x.d_xn_nodes, x.d_xn_send, x.d_xn_recv};
    for (void* p : pp) if (p) cudaFree(p);
    if (x.comm && nccl_api()) nccl_api()->CommDestroy(x.comm); }
  for (auto& b : e->batches) { if (b.d_tab) cudaFree(b.d_tab); if (b.d_map) cudaFree(b.d_map); }
  pipe_free(e);
  { void* pp[] = {e->d_cl_ptr, e->d_cl_nodes, e->d_cl_rec, e->d_cl_fac}; for (void* p : pp) if (p) cudaFree(p); }
  for (auto ev : e->evpool) cudaEventDestroy(ev);
  if (e->ev0) cudaEventDestroy(e->ev0); if (e->ev1) cudaEventDestroy(e->ev1);
  for (int k = 0; k < ORGPU_NSIDE; k++) { if (e->side[k]) cudaStreamDestroy(e->side[k]); if (e->ev_join[k]) cudaEventDestroy(e->ev_join[k]); }
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  if (e->st) cudaStreamDestroy(e->st);
  delete e;
  return 0;
}

int orgpu_upload_nodes(orgpu_engine* e, const double* X, const double* V, const double* VR,
                       const double* D, const double* MS, const double* IN)
{
  NEED(e, -1, "null handle"); CUDA_OK(cudaSetDevice(e->device));
  const size_t n = e->numnod;
  if (X) { if (upload3to4(e, X, e->nd.pos)) return -100; }
  if (V) { if (upload3to4(e, V, e->nd.vel)) return -100; }
  if (VR && e->nd.rot) { if (upload3to4(e, VR, e->nd.rot)) return -100; }
  if (D) CUDA_OK(cudaMemcpy(e->nd.D, D, 24 * n, cudaMemcpyHostToDevice));
  if (MS) CUDA_OK(cudaMemcpy(e->nd.MS, MS, 8 * n, cudaMemcpyHostToDevice));
  if (IN) CUDA_OK(cudaMemcpy(e->nd.IN, IN, 8 * n, cudaMemcpyHostToDevice));
  return 0;
}

int orgpu_set_loads(orgpu_engine* e, const double* FEXT, const double* MEXT)
{
  NEED(e, -1, "null handle"); CUDA_OK(cudaSetDevice(e->device));
  const size_t n = e->numnod;
  if (FEXT) { if (!e->d_fext && dev_alloc(&e->d_fext, 3 * n)) return -100; CUDA_OK(cudaMemcpy(e->d_fext, FEXT, 24 * n, cudaMemcpyHostToDevice)); e->nd.FEXT = e->d_fext; }
  else e->nd.FEXT = nullptr;
  if (MEXT) { if (!e->d_mext && dev_alloc(&e->d_mext, 3 * n)) return -100; CUDA_OK(cudaMemcpy(e->d_mext, MEXT, 24 * n, cudaMemcpyHostToDevice)); e->nd.MEXT = e->d_mext; }
  else e->nd.MEXT = nullptr;
  if (e->gexec) { cudaGraphExecDestroy(e->gexec); e->gexec = nullptr; }
  return 0;
}

int orgpu_set_bcs(orgpu_engine* e, const int* icodt, const int* icodr)
{
  NEED(e && icodt, -1, "orgpu_set_bcs: bad arguments"); CUDA_OK(cudaSetDevice(e->device));
  const size_t n = e->numnod;
  if (!e->d_icodt && dev_alloc(&e->d_icodt, n)) return -100;
  if (!e->d_icodr && dev_alloc(&e->d_icodr, n)) return -100;
  CUDA_OK(cudaMemcpy(e->d_icodt, icodt, 4 * n, cudaMemcpyHostToDevice));
  if (icodr) CUDA_OK(cudaMemcpy(e->d_icodr, icodr, 4 * n, cudaMemcpyHostToDevice));
  e->nd.icodt = e->d_icodt; e->nd.icodr = e->d_icodr;
  if (e->gexec) { cudaGraphExecDestroy(e->gexec); e->gexec = nullptr; }
  return 0;
}

int orgpu_set_solids(orgpu_engine* e, int numels, const int* ixs, const int* iads)
{
  NEED(e && numels >= 0 && !e->finalized, -1, "orgpu_set_solids: bad arguments / already finalized");
  e->numels = numels; e->ixs.assign(ixs, ixs + (size_t)11 * numels); e->iads.assign(iads, iads + (size_t)8 * numels);
  return 0;
}
int orgpu_set_shells(orgpu_engine* e, int numelc, const int* ixc, const int* iadc)
{
  NEED(e && numelc >= 0 && !e->finalized, -1, "orgpu_set_shells: bad arguments / already finalized");
  e->numelc = numelc; e->ixc.assign(ixc, ixc + (size_t)7 * numelc); e->iadc.assign(iadc, iadc + (size_t)4 * numelc);
  return 0;
}
int orgpu_set_sh3n(orgpu_engine* e, int numeltg, const int* ixtg, const int* iadtg)
{
  NEED(e && numeltg >= 0 && !e->finalized, -1, "orgpu_set_sh3n: bad arguments / already finalized");
  e->numeltg = numeltg; e->ixtg.assign(ixtg, ixtg + (size_t)6 * numeltg); e->iadtg.assign(iadtg, iadtg + (size_t)3 * numeltg);
  return 0;
}
int orgpu_set_pon(orgpu_engine* e, const int* adsky, int lsky)
{
  NEED(e && adsky && lsky >= 0 && !e->finalized, -1, "orgpu_set_pon: bad arguments / already finalized");
  NEED(adsky[0] == 1 && adsky[e->numnod] == lsky + 1, -4, "orgpu_set_pon: ADSKY does not span 1..LSKY+1");
  for (int n = 0; n < e->numnod; n++) NEED(adsky[n + 1] >= adsky[n], -4, "orgpu_set_pon: ADSKY decreases at node %d", n + 1);
  e->adsky.assign(adsky, adsky + e->numnod + 1); e->lsky = lsky;
  return 0;
}
int orgpu_set_functions(orgpu_engine* e, int nfunc, const int* npf, const double* tf)
{
  NEED(e && nfunc >= 0 && !e->finalized, -1, "orgpu_set_functions: bad arguments / already finalized");
  e->npf.assign(npf, npf + nfunc + 1); e->tf.assign(tf, tf + (size_t)2 * npf[nfunc]);
  return 0;
}

int orgpu_set_itab(orgpu_engine* e, const int* itab)
{
  NEED(e && itab && !e->finalized, -1, "orgpu_set_itab: bad arguments / already finalized");
  e->itab.assign(itab, itab + e->numnod);
  return 0;
}

// a peer-memory wait that timed out is fatal for the handle (exchange.cuh: the device state stopped advancing)
static int check_abort(orgpu_engine* e)
{
  if (!e->xc.p2p || !e->xc.d_err) return 0;
  int err = 0; CUDA_OK(cudaMemcpy(&err, e->xc.d_err, 4, cudaMemcpyDeviceToHost));
  NEED(err == 0, -8, "orgpu: peer-memory exchange timed out waiting for a neighbour (a rank stopped stepping); the device state of this handle is frozen at the last completed cycle");
  return 0;
}

// Tie-break keys for the time-step arg-min across domains: the index of every local element in the processing order of the
// undecomposed model (4-node shells, then 3-node shells, then solids: what one domain would use), and the global index of
// every local node (nodal time step).  With them N domains elect, on an exact tie, the element a single domain elects.
// Any pointer may be NULL (local order).  Before orgpu_finalize.
int orgpu_set_global_order(orgpu_engine* e, const int* gshell, const int* gsh3n, const int* gsolid, const int* gnode)
{
  NEED(e && !e->finalized, -1, "orgpu_set_global_order: bad handle / already finalized");
  if (gshell) e->gord_c.assign(gshell, gshell + e->numelc);
  if (gsh3n) e->gord_t.assign(gsh3n, gsh3n + e->numeltg);
  if (gsolid) e->gord_s.assign(gsolid, gsolid + e->numels);
  if (gnode) e->gnode.assign(gnode, gnode + e->numnod);
  return 0;
}

int orgpu_set_exchange_timeout(orgpu_engine* e, double seconds)
{
  NEED(e && seconds > 0.0, -1, "orgpu_set_exchange_timeout: bad arguments");
  e->xc.timeout_ns = (unsigned long long)(seconds * 1.0e9);
  if (e->gexec) { cudaGraphExecDestroy(e->gexec); e->gexec = nullptr; }
  return 0;
}

int orgpu_set_parts(orgpu_engine* e, int npart, const int* ipartc, const int* iparts, const int* iparttg, const double* gvolc, const double* gvoltg)
{
  NEED(e && !e->finalized && npart >= 1, -1, "orgpu_set_parts: bad arguments / already finalized (call it after set_solids / set_shells / set_sh3n)");
  NEED((e->numelc == 0 || (ipartc && gvolc)) && (e->numels == 0 || iparts) && (e->numeltg == 0 || (iparttg && gvoltg)), -1, "orgpu_set_parts: a table is missing");
  auto chk = [&](const int* p, int n) { for (int i = 0; i < n; i++) if (p[i] < 0 || p[i] >= npart) return false; return true; };
  NEED(chk(ipartc, e->numelc) && chk(iparts, e->numels) && chk(iparttg, e->numeltg), -4, "orgpu_set_parts: part index out of range");
  e->npart = npart; e->have_parts = true;
  e->ipartc.assign(ipartc, ipartc + e->numelc); e->iparts.assign(iparts, iparts + e->numels); e->iparttg.assign(iparttg, iparttg + e->numeltg);
  e->gvolc.assign(gvolc, gvolc + e->numelc); e->gvoltg.assign(gvoltg, gvoltg + e->numeltg);
  return 0;
}

int orgpu_set_load_function(orgpu_engine* e, int ifunc, double fcx)
{
  NEED(e && !e->finalized, -1, "orgpu_set_load_function: bad handle / already finalized");
  NEED(ifunc >= -1, -1, "orgpu_set_load_function: bad function index %d", ifunc);
  e->lf_func = ifunc; e->lf_fcx = fcx;
  return 0;
}

// FORCE, concentrated loads with a time-dependent abscissa (force.F90:223-245, 301-312): AA = FCY * FINTER(N3, TT * FCX), added to
// A(N2, N1) / AR(N2 - 3, N1) record after record.  One thread per loaded node sums its records in record order (from zero:
// 0 + AA is AA) into the FEXT / MEXT arrays the node kernels start their fold from; runs once per cycle after the dt fold
// (TT of the cycle is cs->tt0).
__global__ void cload_eval_kernel(const CycleState* __restrict__ cs, const FuncTable ft, int nn, const int* __restrict__ ptr, const int* __restrict__ nodes,
                                  const int2* __restrict__ rec, const double2* __restrict__ fac, double* __restrict__ fext, double* __restrict__ mext)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= nn || cs->abort) return;
  const int n = nodes[i]; const double tt = cs->tt0;
  double a[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  for (int r = ptr[i]; r < ptr[i + 1]; r++) {
    const int2 q = rec[r]; const double2 f = fac[r];
    double v = f.x;
    if (q.y >= 0) { const int i0 = ft.npf[q.y]; v = f.x * or_finter(ft.tf, i0, ft.npf[q.y + 1] - i0, tt * f.y); }
    a[q.x] = a[q.x] + v;
  }
  fext[3 * n] = a[0]; fext[3 * n + 1] = a[1]; fext[3 * n + 2] = a[2];
  if (mext) { mext[3 * n] = a[3]; mext[3 * n + 1] = a[4]; mext[3 * n + 2] = a[5]; }
}

int orgpu_set_cloads(orgpu_engine* e, int nload, const int* ib /*(3,nload): node (1-based), direction 1..6, function (0-based, -1: constant)*/,
                     const double* fac /*(2,nload): FCY, FCX*/)
{
  NEED(e && !e->finalized && nload >= 0 && (nload == 0 || (ib && fac)), -1, "orgpu_set_cloads: bad arguments / already finalized");
  for (int l = 0; l < nload; l++) {
    NEED(ib[3 * l] >= 1 && ib[3 * l] <= e->numnod, -4, "orgpu_set_cloads: record %d: node %d out of range", l + 1, ib[3 * l]);
    NEED(ib[3 * l + 1] >= 1 && ib[3 * l + 1] <= 6, -5, "orgpu_set_cloads: record %d: direction %d (skew frames are outside the built path)", l + 1, ib[3 * l + 1]);
    NEED(ib[3 * l + 1] <= 3 || e->ctl.iroddl, -4, "orgpu_set_cloads: record %d loads a rotation of a model without rotational dofs", l + 1);
  }
  e->cl_ib.assign(ib, ib + (size_t)3 * nload); e->cl_fac.assign(fac, fac + (size_t)2 * nload);
  return 0;
}

int orgpu_set_gravity(orgpu_engine* e, int ngrav, const int* igrv /*(3,n): NN, direction 1..3, function*/, const double* agrv /*(2,n): FCY, FCX*/,
                      const int* ib, int lib)
{
  NEED(e && !e->finalized && ngrav >= 0 && (ngrav == 0 || (igrv && agrv && ib)), -1, "orgpu_set_gravity: bad arguments / already finalized");
  NEED(ngrav <= ORGPU_MAXGRAV, -5, "%d gravity loads: more than the %d carried", ngrav, ORGPU_MAXGRAV);
  e->ngrav = ngrav; e->gmask.assign(ngrav ? e->numnod : 0, 0);
  int iad = 0;
  for (int l = 0; l < ngrav; l++) {
    const int nn = igrv[3 * l], dir = igrv[3 * l + 1], f = igrv[3 * l + 2];
    NEED(dir >= 1 && dir <= 3, -5, "gravity load %d: direction %d (a skew / moving frame, IGRV(2) >= 10) is outside the built path", l, dir);
    NEED(nn >= 0 && iad + nn <= lib && f >= -1, -4, "gravity load %d: bad node count / function", l);
    e->gdir[l] = dir - 1; e->gfunc[l] = f; e->gfcy[l] = agrv[2 * l]; e->gfcx[l] = agrv[2 * l + 1];
    for (int j = 0; j < nn; j++) {
      const int node = abs(ib[iad + j]);              // the sign of IB only selects the nodes counted in the external work
      NEED(node >= 1 && node <= e->numnod, -4, "gravity load %d: node %d out of range", l, node);
      e->gmask[node - 1] |= (1u << l);
    }
    iad += nn;
  }
  return 0;
}

int orgpu_set_fixvel(orgpu_engine* e, int nfxvel, const int* ibfv /*(3,n): node, direction 1..3, function*/,
                     const double* vel /*(4,n): FAC, STARTT, STOPT, FACX*/)
{
  NEED(e && !e->finalized && nfxvel >= 0 && (nfxvel == 0 || (ibfv && vel)), -1, "orgpu_set_fixvel: bad arguments / already finalized");
  e->fv_idx.assign(e->numnod, -1); e->fv.clear();
  for (int k = 0; k < nfxvel; k++) {
    const int node = ibfv[3 * k], j = ibfv[3 * k + 1], f = ibfv[3 * k + 2];
    NEED(node >= 1 && node <= e->numnod, -4, "orgpu_set_fixvel: node %d out of range", node);
    NEED(j >= 1 && j <= 3, -5, "imposed velocity on direction %d (rotations, skew / moving frames) is outside the built path", j);
    NEED(f >= 0, -4, "orgpu_set_fixvel: bad function index %d", f);
    int& idx = e->fv_idx[node - 1];
    if (idx < 0) { idx = (int)e->fv.size(); FixVelNode r; memset(&r, 0, sizeof r); r.func[0] = r.func[1] = r.func[2] = -1; e->fv.push_back(r); }
    FixVelNode& r = e->fv[idx];
    NEED(r.func[j - 1] < 0, -4, "orgpu_set_fixvel: node %d direction %d imposed twice", node, j);
    r.func[j - 1] = f; r.fac[j - 1] = vel[4 * k]; r.tstart[j - 1] = vel[4 * k + 1]; r.tstop[j - 1] = vel[4 * k + 2]; r.facx[j - 1] = vel[4 * k + 3];
  }
  return 0;
}

int orgpu_add_solid_group(orgpu_engine* e, int nel, int nft, const orgpu_law2* mat,
                          const orgpu_prop_solid* prop, const double* vol0)
{
  NEED(e && mat && prop && vol0 && nel > 0 && !e->finalized, -1, "orgpu_add_solid_group: bad arguments / already finalized");
  NEED(nft >= 0 && nft + nel <= e->numels, -4, "orgpu_add_solid_group: elements [%d,%d) outside IXS (%d)", nft, nft + nel, e->numels);
  NEED(mat->fisokin >= 0.0 && mat->fisokin <= 1.0, -4, "LAW2 FISOKIN = %g outside [0, 1]", mat->fisokin);
  NEED(prop->jhbe == 0 || prop->jhbe == 1 || prop->jhbe == 2 || prop->jhbe == 101 || prop->jhbe == 102, -5, "Isolid=%d is outside the built path (0,1,2,101,102)", prop->jhbe);
  NEED(prop->ismstr == 1 || prop->ismstr == 2 || prop->ismstr == 4, -5, "Ismstr=%d is outside the built path (1,2,4)", prop->ismstr);
  NEED(prop->jcvt == 0 || (prop->jcvt == 1 && prop->jhbe != 0), -5, "solid Iframe: JCVT=%d with Isolid=%d is outside the built path (0 global; 1 co-rotational with Isolid 1 / 2: SDEFO3 tests JCVT before JHBE, sdefo3.F:158, 222)", prop->jcvt, prop->jhbe);
  HostSolidGroup g; g.nel = nel; g.nft = nft; g.law = 2; g.mat = *mat; memset(&g.m36, 0, sizeof g.m36); g.prop = *prop; g.vol0.assign(vol0, vol0 + nel);
  if (g.prop.jhbe > 100) g.prop.jhbe = 2;   // Isolid 101 / 102 (forint.F:1159): the Engine only tests JHBE /= 0 (sderi3.F:303), >= 1 (shvis3.F:318), >= 2 (sdefo3.F:222)
  e->sgroups.push_back(std::move(g));
  return (int)e->sgroups.size() - 1;
}

int orgpu_add_solid_group_law(orgpu_engine* e, int nel, int nft, int law, const void* mat,
                              const orgpu_prop_solid* prop, const double* vol0)
{
  if (law == 2) return orgpu_add_solid_group(e, nel, nft, (const orgpu_law2*)mat, prop, vol0);
  NEED(law == 36, -5, "solid law %d is outside the built path (2, 36)", law);
  NEED(e && mat && prop && vol0 && nel > 0 && !e->finalized, -1, "orgpu_add_solid_group_law: bad arguments / already finalized");
  NEED(nft >= 0 && nft + nel <= e->numels, -4, "orgpu_add_solid_group_law: elements [%d,%d) outside IXS (%d)", nft, nft + nel, e->numels);
  const orgpu_law36* m = (const orgpu_law36*)mat;
  NEED(m->fisokin == 0.0 && m->vp == 0 && m->ifail >= 0 && m->ifail <= 2, -5, "LAW36 kinematic hardening / VP=1 are outside the built path");
  NEED(m->ifail != 2 || prop->istrain > 0, -5, "LAW36 tensile-strain failure (IFAIL=2) needs the total strains (Istrain=1)");
  NEED(m->nrate >= 1 && m->nrate <= ORGPU_MAXFUNC36, -5, "LAW36 NRATE=%d out of range", m->nrate);
  NEED(prop->jhbe == 0 || prop->jhbe == 1 || prop->jhbe == 2 || prop->jhbe == 101 || prop->jhbe == 102, -5, "Isolid=%d is outside the built path (0,1,2,101,102)", prop->jhbe);
  NEED(prop->ismstr == 1 || prop->ismstr == 2 || prop->ismstr == 4, -5, "Ismstr=%d is outside the built path (1,2,4)", prop->ismstr);
  NEED(prop->ipla >= 0 && prop->ipla <= 2, -5, "solid Iplas=%d is outside the built path (0,1,2)", prop->ipla);
  NEED(prop->jcvt == 0 || (prop->jcvt == 1 && prop->jhbe != 0 && m->ifail != 2), -5, "solid Iframe: JCVT=%d with Isolid=%d / IFAIL=%d is outside the built path", prop->jcvt, prop->jhbe, m->ifail);
  HostSolidGroup g; g.nel = nel; g.nft = nft; g.law = 36; memset(&g.mat, 0, sizeof g.mat); g.mat.rho0 = m->rho0;
  g.m36 = *m; g.prop = *prop; g.vol0.assign(vol0, vol0 + nel);
  if (g.prop.jhbe > 100) g.prop.jhbe = 2;   // Isolid 101 / 102 (forint.F:1159): the Engine only tests JHBE /= 0 (sderi3.F:303), >= 1 (shvis3.F:318), >= 2 (sdefo3.F:222)
  e->sgroups.push_back(std::move(g));
  return (int)e->sgroups.size() - 1;
}

int orgpu_add_shell_group(orgpu_engine* e, int nel, int nft, int law, const void* mat, const orgpu_prop_shell* prop)
{
  NEED(e && mat && prop && nel > 0 && !e->finalized, -1, "orgpu_add_shell_group: bad arguments / already finalized");
  NEED(nft >= 0 && nft + nel <= e->numelc, -4, "orgpu_add_shell_group: elements [%d,%d) outside IXC (%d)", nft, nft + nel, e->numelc);
  return shell_add_group(e->cgroups, nel, nft, law, mat, prop);
}

int orgpu_add_sh3n_group(orgpu_engine* e, int nel, int nft, int law, const void* mat, const orgpu_prop_shell* prop)
{
  NEED(e && mat && prop && nel > 0 && !e->finalized, -1, "orgpu_add_sh3n_group: bad arguments / already finalized");
  NEED(nft >= 0 && nft + nel <= e->numeltg, -4, "orgpu_add_sh3n_group: elements [%d,%d) outside IXTG (%d)", nft, nft + nel, e->numeltg);
  return shell_add_group(e->tgroups, nel, nft, law, mat, prop, true);
}

// ---- FORINTC_PREPARE_GPU analogue -----------------------------------------------------------

int orgpu_set_shell_group_fail(orgpu_engine* e, int sh3n, int group, const orgpu_fail* f)
{
  NEED(e && f && !e->finalized, -1, "orgpu_set_shell_group_fail: bad arguments / already finalized");
  std::vector<HostShellGroup>& gs = sh3n ? e->tgroups : e->cgroups;
  NEED(group >= 0 && group < (int)gs.size(), -4, "orgpu_set_shell_group_fail: group %d does not exist", group);
  NEED(f->irupt == 0 || f->irupt == 1, -5, "failure model %d is outside the built path (1: /FAIL/JOHNSON)", f->irupt);
  NEED(f->irupt == 0 || f->d5 == 0.0, -5, "/FAIL/JOHNSON with D5 (temperature term) is outside the built path");
  NEED(f->irupt == 0 || f->d4 == 0.0 || f->epsp0 > 0.0, -4, "/FAIL/JOHNSON: D4 needs a positive reference strain rate");
  gs[group].fail = *f; gs[group].fail.pad = 0;
  return 0;
}

int orgpu_set_solid_group_fail(orgpu_engine* e, int group, const orgpu_fail* f)
{
  NEED(e && f && !e->finalized, -1, "orgpu_set_solid_group_fail: bad arguments / already finalized");
  NEED(group >= 0 && group < (int)e->sgroups.size(), -4, "orgpu_set_solid_group_fail: group %d does not exist", group);
  NEED(f->irupt == 0 || f->irupt == 1, -5, "failure model %d is outside the built path (1: /FAIL/JOHNSON)", f->irupt);
  NEED(f->irupt == 0 || e->sgroups[group].law == 2, -5, "/FAIL/JOHNSON on solids is built behind MMAIN's own failure section (LAW2), not behind MULAW (LAW36)");
  NEED(f->irupt == 0 || f->d5 == 0.0, -5, "/FAIL/JOHNSON with D5 (temperature term) is outside the built path");
  NEED(f->irupt == 0 || f->d4 == 0.0 || f->epsp0 > 0.0, -4, "/FAIL/JOHNSON: D4 needs a positive reference strain rate");
  e->sgroups[group].fail = *f; e->sgroups[group].fail.pad = 0;
  return 0;
}

int orgpu_finalize(orgpu_engine* e)
{
  NEED(e && !e->finalized, -1, "orgpu_finalize: bad handle / already finalized");
  NEED(!e->adsky.empty(), -4, "orgpu_finalize: /PARITH/ON tables missing (orgpu_set_pon)");
  CUDA_OK(cudaSetDevice(e->device));
  const bool has_shell = !e->cgroups.empty() || !e->tgroups.empty();
  e->roww = (has_shell || e->ctl.iroddl) ? 8 : 4;
  if (dev_alloc(&e->d_fsky, (size_t)e->roww * (e->lsky > 0 ? e->lsky : 1))) return -100;
  { std::vector<int> a0(e->adsky.size()); for (size_t i = 0; i < a0.size(); i++) a0[i] = e->adsky[i] - 1;
    if (dev_alloc(&e->d_adsky, a0.size())) return -100;
    CUDA_OK(cudaMemcpy(e->d_adsky, a0.data(), 4 * a0.size(), cudaMemcpyHostToDevice)); e->nd.adsky = e->d_adsky; }
  // part of a group = part of its elements (one part per group, as in the Engine's group building)
  { auto gpart = [&](const std::vector<int>& ip, int nft, int nel, int& part) {
      part = 0; if (ip.empty()) return true; part = ip[nft];
      for (int i = 1; i < nel; i++) if (ip[nft + i] != part) return false; return true; };
    for (auto& g : e->cgroups) NEED(gpart(e->ipartc, g.nft, g.nel, g.part), -4, "a shell group spans several parts");
    for (auto& g : e->tgroups) NEED(gpart(e->iparttg, g.nft, g.nel, g.part), -4, "a 3-node shell group spans several parts");
    for (auto& g : e->sgroups) NEED(gpart(e->iparts, g.nft, g.nel, g.part), -4, "a solid group spans several parts"); }
  int order = 0, blk = 0; e->fa.nsg = 0; e->sgr.clear();
  // shells are processed first (FORINTC resol.F:4138), solids after (FORINT resol.F:4225)
  { int rc = shell_build_supergroups(e->cgroups, e->csg, e->ixc, e->iadc, 4, e->npf, e->tf, e->ctl, e->numnod, e->lsky, order, blk, e->sgr); if (rc) return rc; }
  // 3-node shell groups (ITY=7) follow the 4-node ones in the Engine's group list, inside the same FORINTC pass
  { int rc = shell_build_supergroups(e->tgroups, e->csg, e->ixtg, e->iadtg, 3, e->npf, e->tf, e->ctl, e->numnod, e->lsky, order, blk, e->sgr); if (rc) return rc; }
  e->fa.brick_blk0 = blk;          // dt slots from here on belong to solids
  // consecutive solid groups with identical material / property fuse into one super-group
  size_t gi = 0;
  while (gi < e->sgroups.size()) {
    size_t gj = gi + 1;
    while (gj < e->sgroups.size() && e->sgroups[gj].nft == e->sgroups[gj - 1].nft + e->sgroups[gj - 1].nel &&
           e->sgroups[gj].law == e->sgroups[gi].law &&
           !memcmp(&e->sgroups[gj].mat, &e->sgroups[gi].mat, sizeof(orgpu_law2)) && !memcmp(&e->sgroups[gj].m36, &e->sgroups[gi].m36, sizeof(orgpu_law36)) &&
           !memcmp(&e->sgroups[gj].prop, &e->sgroups[gi].prop, sizeof(orgpu_prop_solid)) && e->sgroups[gj].part == e->sgroups[gi].part &&
           !memcmp(&e->sgroups[gj].fail, &e->sgroups[gi].fail, sizeof(orgpu_fail))) gj++;
    int ne = 0; for (size_t k = gi; k < gj; k++) ne += e->sgroups[k].nel;
    const int nft = e->sgroups[gi].nft;
    const int np = ((ne + ORGPU_BLOCK - 1) / ORGPU_BLOCK) * ORGPU_BLOCK;
    e->bsg.emplace_back(); BrickSGHost& S = e->bsg.back(); S.first_elem = nft; S.part = e->sgroups[gi].part;
    BrickSG& d = S.d; d.bal = nullptr; d.bal_ld = 0; d.xs_ftile = nullptr; d.ne = ne; d.ne_pad = np; d.order0 = order; d.blk0 = blk;
    d.mat = e->sgroups[gi].mat; d.prop = e->sgroups[gi].prop; d.dtfac = e->ctl.dtfac_brick; d.nodadt = e->ctl.nodadt;
    std::vector<int> conn((size_t)8 * np, 0), ngl(np, 0), conn_t;
    // state slab: read/write words first (SIG 6, EINT, RHO, QVIS, PLA, EPSD, OFF[, TEMP]), then VOL and the slot rows
    d.law = e->sgroups[gi].law; d.m36 = e->sgroups[gi].m36;
    d.w_temp = d.mat.has_temp ? BW_NFIX : -1;
    d.nw_rw = BW_NFIX + (d.mat.has_temp ? 1 : 0);
    d.w_stra = d.w_wpla = d.w_vt = -1; d.nvt = 0; d.tf = nullptr; d.npf = nullptr; memset(&d.ct, 0, sizeof d.ct);
    d.w_sigb = -1;
    if (d.law == 2 && d.mat.fisokin > 0.0) { d.w_sigb = d.nw_rw; d.nw_rw += 6; }     // LBUF%SIGB (m2law.F:181-190, 364-390)
    d.fail = e->sgroups[gi].fail; d.w_dfmax = -1;
    if (d.fail.irupt == 1) d.w_dfmax = d.nw_rw++;                                     // /FAIL/JOHNSON: FBUF%FLOC%DAMMX of the element
    if (d.law == 36) {                                   // LBUF%WPLA, LBUF%STRA (ISTRAIN>0), VARTMP cursors
      d.w_wpla = d.nw_rw++;
      if (d.prop.istrain > 0) { d.w_stra = d.nw_rw; d.nw_rw += 6; }
      d.nvt = (d.m36.nrate == 1) ? 1 : 2 + d.m36.nrate;
      d.w_vt = d.nw_rw; d.nw_rw += (d.nvt + 1) / 2;
      NEED(!e->npf.empty(), -4, "LAW36 group without a function table (orgpu_set_functions)");
      for (int j = 0; j < d.m36.nrate; j++) {
        const int f = d.m36.ifunc[j];
        NEED(f >= 0 && f + 1 < (int)e->npf.size() && e->npf[f + 1] - e->npf[f] >= 2, -4, "LAW36 curve %d missing or shorter than 2 points", f);
      }
      if (!e->d_btf) {
        if (dev_alloc(&e->d_btf, e->tf.size()) || dev_alloc(&e->d_bnpf, e->npf.size())) return -100;
        CUDA_OK(cudaMemcpy(e->d_btf, e->tf.data(), 8 * e->tf.size(), cudaMemcpyHostToDevice));
        CUDA_OK(cudaMemcpy(e->d_bnpf, e->npf.data(), 4 * e->npf.size(), cudaMemcpyHostToDevice));
      }
      d.tf = e->d_btf; d.npf = e->d_bnpf;
      curve_tab_fill(d.ct, d.m36, e->npf, e->tf);
    }
    d.w_vol = d.nw_rw; d.w_slot = d.nw_rw + 1; d.nw = d.nw_rw + 1 + 4;
    HostSlab H; H.init(d.nw, np);
    for (int i = 0; i < np; i++) { H.at(d.w_vol, i) = 1.0; H.at(BW_RHO, i) = d.mat.rho0; if (d.w_temp >= 0) H.at(d.w_temp, i) = d.mat.tini; }
    for (int i = 0; i < ne; i++) {
      const int* ix = &e->ixs[(size_t)11 * (nft + i)];
      for (int k = 0; k < 8; k++) {
        int node = ix[1 + k]; NEED(node >= 1 && node <= e->numnod, -4, "IXS node %d out of range (element %d)", node, nft + i + 1);
        int sl = e->iads[(size_t)8 * (nft + i) + k]; NEED(sl >= 1 && sl <= e->lsky, -4, "IADS slot %d out of range (element %d)", sl, nft + i + 1);
        conn[(size_t)k * np + i] = node - 1; H.iat(d.w_slot, k, i) = sl - 1;
      }
      ngl[i] = ix[10]; H.at(BW_OFF, i) = 1.0;
    }
    { int i = 0; for (size_t k = gi; k < gj; k++) for (int j = 0; j < e->sgroups[k].nel; j++) H.at(d.w_vol, i++) = e->sgroups[k].vol0[j]; }
    tile_major_ints(conn_t, conn, 8, np);
    int *dconn, *dngl;
    if (upload_vec(S.owned, &dconn, conn_t) || upload_vec(S.owned, &dngl, ngl) || upload_vec(S.owned, &d.slab, H.h)) return -100;
    d.conn = dconn; d.ngl = dngl;
    if (push_dev(S.owned, &d.smstr, (size_t)21 * np)) return -100;
    const int nblk = np / ORGPU_TILE;                    // dt candidate slots: one per CTA
    NEED((int)e->sgr.size() < ORGPU_MAX_SG, -6, "too many super-groups (%d)", ORGPU_MAX_SG);
    e->sgr.push_back(SGRange{blk, nblk, ORGPU_FAM_BRICK, d.order0, d.ngl, nullptr});
    order += ne; blk += nblk; gi = gj;
  }
  NEED(blk > 0, -4, "orgpu_finalize: no element groups");
  // GBUF%VOL of the shells (mass of CBILAN / C3BILAN), element order of each super-group
  for (auto& S : e->csg) {
    S.d.bal = nullptr; S.d.bal_ld = 0; S.d.gvol = nullptr;
    const std::vector<double>& gv = S.sh3n ? e->gvoltg : e->gvolc;
    if (gv.empty()) continue;
    std::vector<double> h(gv.begin() + S.first_elem, gv.begin() + S.first_elem + S.d.ne); h.resize(S.d.ne_pad, 0.0);
    double* dg; if (upload_vec(S.owned, &dg, h)) return -100; S.d.gvol = dg;
  }
  { // global processing order of every super-group's elements (tie-break across domains); sgr follows csg then bsg
    size_t k = 0;
    auto up = [&](std::vector<void*>& owned, const std::vector<int>& g, int first, int ne, SGRange& r) -> int {
      r.gord = nullptr; if (g.empty()) return 0;
      std::vector<int> h(g.begin() + first, g.begin() + first + ne); int* d; if (upload_vec(owned, &d, h)) return -100; r.gord = d; return 0; };
    for (auto& S : e->csg) { if (up(S.owned, S.sh3n ? e->gord_t : e->gord_c, S.first_elem, S.d.ne, e->sgr[k])) return -100; k++; }
    for (auto& S : e->bsg) { if (up(S.owned, e->gord_s, S.first_elem, S.d.ne, e->sgr[k])) return -100; k++; }
    e->nd.gnode = nullptr;
    if (!e->gnode.empty()) { if (dev_alloc(&e->d_gnode, e->gnode.size())) return -100;
      CUDA_OK(cudaMemcpy(e->d_gnode, e->gnode.data(), 4 * e->gnode.size(), cudaMemcpyHostToDevice)); e->nd.gnode = e->d_gnode; } }
  if (dev_alloc(&e->d_sgr, e->sgr.size())) return -100;
  CUDA_OK(cudaMemcpy(e->d_sgr, e->sgr.data(), sizeof(SGRange) * e->sgr.size(), cudaMemcpyHostToDevice));
  e->fa.nsg = (int)e->sgr.size(); e->fa.sg = e->d_sgr;
  // /DT/NODA
  NEED(e->ctl.nodadt == 0 || e->ctl.nodadt == 1, -5, "NODADT=%d is outside the built path (0, 1)", e->ctl.nodadt);
  e->nd.nodadt = e->ctl.nodadt; e->nd.dtfac_node = e->ctl.dtfac_node;
  if (e->ctl.nodadt) {
    const size_t ncta = (e->numnod + ORGPU_NODE_BLOCK - 1) / ORGPU_NODE_BLOCK;
    if (dev_alloc(&e->d_nd_dt, 2 * ncta) || dev_alloc(&e->d_nd_node, 2 * ncta)) return -100;
    e->nd.nd_dt = e->d_nd_dt; e->nd.nd_node = e->d_nd_node;
    if (!e->itab.empty()) { if (dev_alloc(&e->d_itab, e->itab.size())) return -100;
      CUDA_OK(cudaMemcpy(e->d_itab, e->itab.data(), 4 * e->itab.size(), cudaMemcpyHostToDevice)); e->nd.itab = e->d_itab; }
  }
  // /GRAV loads: per-node mask, per-load scalars for the finalize kernel
  e->fa.ngrav = e->ngrav; e->nd.gmask = nullptr;
  for (int l = 0; l < ORGPU_MAXGRAV; l++) { e->fa.gfunc[l] = e->gfunc[l]; e->fa.gfcy[l] = e->gfcy[l]; e->fa.gfcx[l] = e->gfcx[l]; e->nd.gdir[l] = e->gdir[l]; }
  if (e->ngrav > 0) {
    if (dev_alloc(&e->d_gmask, e->gmask.size())) return -100;
    CUDA_OK(cudaMemcpy(e->d_gmask, e->gmask.data(), 4 * e->gmask.size(), cudaMemcpyHostToDevice)); e->nd.gmask = e->d_gmask;
  }
  // time functions used at node level (loads, imposed velocities)
  e->fa.lf_func = -1; e->fa.lf_fcx = 1.0; e->fa.ft = FuncTable{nullptr, nullptr};
  bool gfun = false; for (int l = 0; l < e->ngrav; l++) gfun = gfun || e->gfunc[l] >= 0;
  const int ncl = (int)e->cl_ib.size() / 3;
  for (int l = 0; l < ncl; l++) gfun = gfun || e->cl_ib[3 * l + 2] >= 0;
  NEED(!(ncl && (e->lf_func >= 0 || e->nd.FEXT || e->nd.MEXT)), -4, "concentrated loads come either as records (orgpu_set_cloads) or as nodal arrays (orgpu_set_loads / orgpu_set_load_function), not both");
  if (e->lf_func >= 0 || !e->fv.empty() || gfun) {
    NEED(!e->npf.empty(), -4, "a load / imposed-velocity time function needs orgpu_set_functions");
    const int nf = (int)e->npf.size() - 1;
    auto check = [&](int f, int maxpts) { return f >= 0 && f < nf && e->npf[f + 1] - e->npf[f] >= 1 && e->npf[f + 1] - e->npf[f] <= maxpts; };
    if (e->lf_func >= 0) NEED(check(e->lf_func, 1 << 30), -4, "load function %d missing", e->lf_func);
    for (int l = 0; l < e->ngrav; l++) if (e->gfunc[l] >= 0) NEED(check(e->gfunc[l], 1 << 30), -4, "gravity function %d missing", e->gfunc[l]);
    for (auto& r : e->fv) for (int j = 0; j < 3; j++) if (r.func[j] >= 0) NEED(check(r.func[j], 1 << 30) && e->npf[r.func[j] + 1] - e->npf[r.func[j]] >= 2, -4, "imposed-velocity function %d missing or shorter than 2 points", r.func[j]);
    if (dev_alloc(&e->d_ftf, e->tf.size()) || dev_alloc(&e->d_fnpf, e->npf.size())) return -100;
    CUDA_OK(cudaMemcpy(e->d_ftf, e->tf.data(), 8 * e->tf.size(), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(e->d_fnpf, e->npf.data(), 4 * e->npf.size(), cudaMemcpyHostToDevice));
    e->fa.ft = FuncTable{e->d_ftf, e->d_fnpf}; e->nd.ft = e->fa.ft;
    e->fa.lf_func = e->lf_func; e->fa.lf_fcx = e->lf_fcx;
    for (int l = 0; l < ncl; l++) if (e->cl_ib[3 * l + 2] >= 0) NEED(check(e->cl_ib[3 * l + 2], 1 << 30), -4, "load function %d of record %d missing", e->cl_ib[3 * l + 2], l + 1);
    if (!e->fv.empty()) {
      if (dev_alloc(&e->d_fv_idx, e->fv_idx.size()) || dev_alloc(&e->d_fv, e->fv.size())) return -100;
      CUDA_OK(cudaMemcpy(e->d_fv_idx, e->fv_idx.data(), 4 * e->fv_idx.size(), cudaMemcpyHostToDevice));
      CUDA_OK(cudaMemcpy(e->d_fv, e->fv.data(), sizeof(FixVelNode) * e->fv.size(), cudaMemcpyHostToDevice));
      e->nd.fv_idx = e->d_fv_idx; e->nd.fv = e->d_fv;
    }
  }
  if (ncl) {
    // records grouped by node, record order kept inside a node (the order of the additions)
    std::vector<int> idx(ncl); for (int l = 0; l < ncl; l++) idx[l] = l;
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return e->cl_ib[3 * a] < e->cl_ib[3 * b]; });
    std::vector<int> ptr, nodes; std::vector<int2> rec(ncl); std::vector<double2> fac(ncl);
    bool anyrot = false;
    for (int k = 0; k < ncl; k++) {
      const int l = idx[k], n = e->cl_ib[3 * l] - 1;
      if (nodes.empty() || nodes.back() != n) { nodes.push_back(n); ptr.push_back(k); }
      rec[k] = make_int2(e->cl_ib[3 * l + 1] - 1, e->cl_ib[3 * l + 2]); fac[k] = make_double2(e->cl_fac[2 * l], e->cl_fac[2 * l + 1]);
      anyrot = anyrot || e->cl_ib[3 * l + 1] > 3;
    }
    ptr.push_back(ncl); e->n_cl_nodes = (int)nodes.size();
    if (dev_alloc(&e->d_cl_ptr, ptr.size()) || dev_alloc(&e->d_cl_nodes, nodes.size()) || dev_alloc(&e->d_cl_rec, rec.size()) || dev_alloc(&e->d_cl_fac, fac.size())) return -100;
    CUDA_OK(cudaMemcpy(e->d_cl_ptr, ptr.data(), 4 * ptr.size(), cudaMemcpyHostToDevice)); CUDA_OK(cudaMemcpy(e->d_cl_nodes, nodes.data(), 4 * nodes.size(), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(e->d_cl_rec, rec.data(), sizeof(int2) * rec.size(), cudaMemcpyHostToDevice)); CUDA_OK(cudaMemcpy(e->d_cl_fac, fac.data(), sizeof(double2) * fac.size(), cudaMemcpyHostToDevice));
    const size_t n = e->numnod;
    if (!e->d_fext && dev_alloc(&e->d_fext, 3 * n)) return -100;          // dev_alloc zero-fills: nodes without a record stay unloaded
    e->nd.FEXT = e->d_fext;
    if (anyrot) { if (!e->d_mext && dev_alloc(&e->d_mext, 3 * n)) return -100; e->nd.MEXT = e->d_mext; }
  }
  e->db.nblocks_total = blk;
  if (dev_alloc(&e->db.dt, blk) || dev_alloc(&e->db.order, blk)) return -100;
  e->finalized = true;
  return 0;
}

// ---- stepping ---------------------------------------------------------------------------------
static cudaEvent_t get_event(orgpu_engine* e, size_t i) {
  while (e->evpool.size() <= i) { cudaEvent_t ev; cudaEventCreate(&ev); e->evpool.push_back(ev); }
  return e->evpool[i];
}

// (Re)build the device tables of the batched launches: descriptors change when the balances or the inline sends are switched on,
// so this runs before the first launch and whenever a descriptor changed -- never inside a graph capture.
static int refresh_batches(orgpu_engine* e)
{
  if (!e->tabs_dirty) return 0;
  for (auto& b : e->batches) { if (b.d_tab) cudaFree(b.d_tab); if (b.d_map) cudaFree(b.d_map); }
  e->batches.clear(); e->sh_batched.assign(e->csg.size(), 0); e->br_batched.assign(e->bsg.size(), 0);
  const char* mn = getenv("ORGPU_TAB_MIN"); const size_t tab_min = mn ? (size_t)atoi(mn) : 8;      // fewer super-groups of a variant: one launch each
  for (int brick = 0; brick < 2; brick++) {
    const int nv = brick ? (int)BRV_COUNT : (int)SHV_COUNT;
    for (int v = 0; v < nv; v++) {
      orgpu_engine::Batch b; b.brick = brick != 0; b.variant = v;
      if (brick) { for (size_t k = 0; k < e->bsg.size(); k++) if (brick_tab_variant(e->bsg[k].d) == v) b.sgs.push_back((int)k); }
      else       { for (size_t k = 0; k < e->csg.size(); k++) if (shell_tab_variant(e->csg[k]) == v) b.sgs.push_back((int)k); }
      if (b.sgs.size() < tab_min) continue;
      std::vector<int2> map;
      if (brick) {
        std::vector<BrickSG> tab;
        for (size_t j = 0; j < b.sgs.size(); j++) { const BrickSG& d = e->bsg[b.sgs[j]].d; tab.push_back(d);
          b.bytes = std::max(b.bytes, (size_t)d.nw * ORGPU_TILE * 8); for (int t = 0; t < d.ne_pad / ORGPU_TILE; t++) map.push_back(make_int2((int)j, t)); e->br_batched[b.sgs[j]] = 1; }
        CUDA_OK(cudaMalloc(&b.d_tab, sizeof(BrickSG) * tab.size())); CUDA_OK(cudaMemcpy(b.d_tab, tab.data(), sizeof(BrickSG) * tab.size(), cudaMemcpyHostToDevice));
      } else {
        std::vector<ShellSG> tab;
        for (size_t j = 0; j < b.sgs.size(); j++) { const ShellSG& d = e->csg[b.sgs[j]].d; tab.push_back(d);
          b.bytes = std::max(b.bytes, (size_t)d.nw * ORGPU_TILE * 8); for (int t = 0; t < d.ne_pad / ORGPU_TILE; t++) map.push_back(make_int2((int)j, t)); e->sh_batched[b.sgs[j]] = 1; }
        CUDA_OK(cudaMalloc(&b.d_tab, sizeof(ShellSG) * tab.size())); CUDA_OK(cudaMemcpy(b.d_tab, tab.data(), sizeof(ShellSG) * tab.size(), cudaMemcpyHostToDevice));
      }
      b.ntile = (int)map.size();
      CUDA_OK(cudaMalloc((void**)&b.d_map, sizeof(int2) * map.size())); CUDA_OK(cudaMemcpy(b.d_map, map.data(), sizeof(int2) * map.size(), cudaMemcpyHostToDevice));
      e->batches.push_back(std::move(b));
    }
  }
  e->tabs_dirty = false;
  return 0;
}

static void launch_sg_shell(orgpu_engine* e, ShellSGHost& S, cudaStream_t st)
{ launch_shell_forces(S, e->nd, e->d_fsky, e->roww, e->d_cs, e->db, e->fa, st); e->launches++; }
static void launch_sg_brick(orgpu_engine* e, BrickSGHost& S, cudaStream_t st)
{ launch_brick_forces(S.d, e->nd, e->d_fsky, e->roww, e->d_cs, e->db, e->fa, st); e->launches++; }

static void launch_element_phase(orgpu_engine* e, int fused, size_t* evi)
{
  e->fa.fused = fused;
  // Super-groups write disjoint FSKY rows, state tiles and dt slots: with more than one of them (models with many parts)
  // their kernels are spread over the main stream and up to ORGPU_NSIDE side streams (fork / join by events, also inside the graph
  // capture of run_cycles) so that small launches overlap; the profiled run keeps them in sequence to time each one.
  const size_t nsg = e->csg.size() + e->bsg.size();
  const bool fork = (evi == nullptr) && nsg > 1;
  const int nside = fork ? (int)((nsg - 1 < (size_t)ORGPU_NSIDE) ? nsg - 1 : ORGPU_NSIDE) : 0;
  size_t k = 0;
  auto pick = [&]() -> cudaStream_t {
    const size_t j = k++ % (size_t)(nside + 1);
    return (j == 0) ? e->st : e->side[j - 1];
  };
  if (fork) { cudaEventRecord(e->ev_fork, e->st); for (int j = 0; j < nside; j++) cudaStreamWaitEvent(e->side[j], e->ev_fork, 0); }
  const bool tab = !evi && !e->tabs_dirty;                 // the profiled run times every super-group on its own
  if (tab) for (auto& b : e->batches) {
    if (b.brick) launch_brick_forces_tab(b.variant, (const BrickSG*)b.d_tab, b.d_map, b.ntile, b.bytes, e->nd, e->d_fsky, e->roww, e->d_cs, e->db, pick());
    else         launch_shell_forces_tab(b.variant, (const ShellSG*)b.d_tab, b.d_map, b.ntile, b.bytes, e->nd, e->d_fsky, e->d_cs, e->db, pick());
    e->launches++;
  }
  for (size_t k = 0; k < e->csg.size(); k++) {
    if (tab && e->sh_batched[k]) continue;
    if (evi) cudaEventRecord(get_event(e, (*evi)++), e->st);
    launch_sg_shell(e, e->csg[k], pick());
    if (evi) cudaEventRecord(get_event(e, (*evi)++), e->st);
  }
  for (size_t k = 0; k < e->bsg.size(); k++) {
    if (tab && e->br_batched[k]) continue;
    if (evi) cudaEventRecord(get_event(e, (*evi)++), e->st);
    launch_sg_brick(e, e->bsg[k], pick());
    if (evi) cudaEventRecord(get_event(e, (*evi)++), e->st);
  }
  for (int j = 0; j < nside; j++) { cudaEventRecord(e->ev_join[j], e->side[j]); cudaStreamWaitEvent(e->st, e->ev_join[j], 0); }
  element_finalize_kernel<<<1, ORGPU_FINALIZE_BLOCK, 0, e->st>>>(e->d_cs, e->db, e->fa); e->launches++;
  if (e->n_cl_nodes) {            // concentrated loads of this cycle (TT = cs->tt0), before any node kernel reads FEXT / MEXT
    cload_eval_kernel<<<(e->n_cl_nodes + 127) / 128, 128, 0, e->st>>>(e->d_cs, e->fa.ft, e->n_cl_nodes, e->d_cl_ptr, e->d_cl_nodes, e->d_cl_rec, e->d_cl_fac,
                                                                    e->d_fext, e->nd.MEXT ? e->d_mext : nullptr); e->launches++;
  }
}

// print cycle: fold the element / node scratch rows in a fixed order (three small launches)
static void launch_balance(orgpu_engine* e)
{
  if (!e->ipri) return;
  balance_chunk_kernel<6><<<e->nchunk, 256, 0, e->st>>>(e->d_bal, e->bal_ld, e->bal_ld, e->d_chunks, e->d_epart);
  balance_chunk_kernel<8><<<e->nnchunk, 256, 0, e->st>>>(e->d_nbal, e->nbal_ld, e->numnod, nullptr, e->d_npartial);
  balance_finalize_kernel<<<1, 256, 0, e->st>>>(e->d_chunks, e->nchunk, e->d_epart, e->npart, e->d_npartial, e->nnchunk, e->d_partsav, e->d_bs, e->d_hist, e->d_cs);
  e->launches += 3;
}

// gather + update: one fused kernel, or with /DT/NODA assemble (+ nodal dt candidates) -> fold + clock -> advance
static void launch_node_phase(orgpu_engine* e)
{
  if (e->ctl.nodadt) {
    launch_node_assemble(e->nd, e->d_fsky, e->roww, e->d_cs, e->ctl.iroddl, e->st);
    launch_dtnoda_finalize(e->nd, e->d_cs, 1, e->st);
    launch_node_advance(e->nd, e->d_cs, e->ctl.iroddl, e->st);
    e->launches += 3;
  } else { launch_node_fused(e->nd, e->d_fsky, e->roww, e->d_cs, e->ctl.iroddl, e->st); e->launches++; }
  launch_balance(e);
}

int orgpu_forces_phase(orgpu_engine* e, double dt1)
{
  NEED(e && e->finalized, -1, "orgpu_forces_phase: engine not finalized"); CUDA_OK(cudaSetDevice(e->device));
  { int rc = refresh_batches(e); if (rc) return rc; }
  launch_set_dt(e->d_cs, dt1, 0, 0, 0, e->st); e->launches++;
  launch_element_phase(e, 0, nullptr);
  CUDA_OK(cudaGetLastError());
  return 0;
}
int orgpu_assemble(orgpu_engine* e)
{
  NEED(e && e->finalized, -1, "orgpu_assemble: engine not finalized"); CUDA_OK(cudaSetDevice(e->device));
  launch_node_assemble(e->nd, e->d_fsky, e->roww, e->d_cs, e->ctl.iroddl, e->st); e->launches++;
  if (e->ctl.nodadt) { launch_dtnoda_finalize(e->nd, e->d_cs, 0, e->st); e->launches++; }   // DTNODA: DT2T / NELTST / ITYPTST for the caller
  CUDA_OK(cudaGetLastError());
  return 0;
}
int orgpu_advance(orgpu_engine* e, double dt12, double dt2)
{
  NEED(e && e->finalized, -1, "orgpu_advance: engine not finalized"); CUDA_OK(cudaSetDevice(e->device));
  launch_set_dt(e->d_cs, 0, dt12, dt2, 1, e->st); e->launches++;
  launch_node_advance(e->nd, e->d_cs, e->ctl.iroddl, e->st); e->launches++;
  launch_balance(e);
  CUDA_OK(cudaGetLastError());
  return 0;
}

#define NCCL_OK(call) do { ncclResult_t _r = (call); if (_r != ncclSuccess) { \
  orgpu_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, nccl_api()->GetErrorString(_r)); return -101; } } while (0)

// pack -> one NCCL group (neighbour rows + every rank's dt candidate) -> unpack (+ dt fold when with_dt)
static int exchange_on_stream(orgpu_engine* e, bool with_dt)
{
  Exchange& x = e->xc; NcclApi* N = nccl_api();
  NEED(N && x.comm, -7, "orgpu: exchange requested without orgpu_comm_init");
  const int V = e->roww / 2;
  { const int nthr = x.nsend * V > 1 ? x.nsend * V : 1; const int nb = (nthr + 255) / 256;
    if (e->roww == 8) rows_pack_kernel<8><<<nb, 256, 0, e->st>>>(e->d_fsky, x.d_send_slots, x.nsend, x.d_sendbuf, e->d_cs, with_dt ? x.d_cand_send : nullptr);
    else              rows_pack_kernel<4><<<nb, 256, 0, e->st>>>(e->d_fsky, x.d_send_slots, x.nsend, x.d_sendbuf, e->d_cs, with_dt ? x.d_cand_send : nullptr);
    e->launches++; }
  NCCL_OK(N->GroupStart());
  for (size_t k = 0; k < x.nb_rank.size(); k++) {
    const size_t ns = x.send_ptr[k + 1] - x.send_ptr[k], nr = x.recv_ptr[k + 1] - x.recv_ptr[k];
    if (ns) NCCL_OK(N->Send(x.d_sendbuf + (size_t)e->roww * x.send_ptr[k], ns * e->roww, ncclDouble, x.nb_rank[k], x.comm, e->st));
    if (nr) NCCL_OK(N->Recv(x.d_recvbuf + (size_t)e->roww * x.recv_ptr[k], nr * e->roww, ncclDouble, x.nb_rank[k], x.comm, e->st));
  }
  if (with_dt) {
    for (int q = 0; q < x.nranks; q++) {
      if (q == x.rank) continue;
      NCCL_OK(N->Send(x.d_cand_send, 4, ncclDouble, q, x.comm, e->st));
      NCCL_OK(N->Recv(x.d_cand_recv + 4 * q, 4, ncclDouble, q, x.comm, e->st));
    }
  }
  NCCL_OK(N->GroupEnd());
  if (with_dt) CUDA_OK(cudaMemcpyAsync(x.d_cand_recv + 4 * x.rank, x.d_cand_send, 32, cudaMemcpyDeviceToDevice, e->st));
  { const int nthr = x.nrecv * V > 1 ? x.nrecv * V : 1; const int nb = (nthr + 255) / 256;
    if (e->roww == 8) rows_unpack_kernel<8><<<nb, 256, 0, e->st>>>(e->d_fsky, x.d_recv_slots, x.nrecv, x.d_recvbuf, e->d_cs, with_dt ? x.d_cand_recv : nullptr, x.nranks);
    else              rows_unpack_kernel<4><<<nb, 256, 0, e->st>>>(e->d_fsky, x.d_recv_slots, x.nrecv, x.d_recvbuf, e->d_cs, with_dt ? x.d_cand_recv : nullptr, x.nranks);
    e->launches++; }
  return 0;
}

// peer-memory exchange on the stream: push rows + dt candidate into the neighbours' windows, then wait for
// theirs, scatter them and advance the clock (exchange.cuh).  Pure kernels: capturable in a CUDA graph.
static void p2p_exchange_on_stream(orgpu_engine* e)
{
  Exchange& x = e->xc;
  const int V = e->roww / 4;
  if (e->split) {                                          // the rows left from inside the force kernels (XSend): candidate + flags only
    p2p_publish_kernel<<<1, 32, 0, e->st>>>(e->d_cs, x.d_peer_cand, x.d_peer_flag, x.nranks, x.rank, x.d_xcycle); e->launches++;
  } else
  { const int nthr = x.nsend * V > 1 ? x.nsend * V : 1; const int nb = (nthr + 255) / 256;
    if (e->roww == 8) p2p_push_kernel<8><<<nb, 256, 0, e->st>>>(e->d_fsky, x.d_send_slots, x.d_send_nb, x.d_nb_sendptr, x.d_nb_rows, x.nsend, e->d_cs, x.d_peer_cand, x.d_peer_flag, x.nranks, x.rank, x.d_xcycle, x.d_done);
    else              p2p_push_kernel<4><<<nb, 256, 0, e->st>>>(e->d_fsky, x.d_send_slots, x.d_send_nb, x.d_nb_sendptr, x.d_nb_rows, x.nsend, e->d_cs, x.d_peer_cand, x.d_peer_flag, x.nranks, x.rank, x.d_xcycle, x.d_done);
    e->launches++; }
  { const int nthr = x.nrecv * V > 1 ? x.nrecv * V : 1; const int nb = (nthr + 255) / 256;
    const int adv = e->ctl.nodadt ? 0 : 1;                 // /DT/NODA: the clock advances after the nodal dt exchange
    if (e->roww == 8) p2p_wait_unpack_kernel<8><<<nb, 256, 0, e->st>>>(e->d_fsky, x.d_recv_slots, x.nrecv, x.win, e->d_cs, x.nranks, x.d_xcycle, x.d_err, adv, x.timeout_ns);
    else              p2p_wait_unpack_kernel<4><<<nb, 256, 0, e->st>>>(e->d_fsky, x.d_recv_slots, x.nrecv, x.win, e->d_cs, x.nranks, x.d_xcycle, x.d_err, adv, x.timeout_ns);
    e->launches++; }
}
// node phase of a multi-domain cycle over peer memory
static void p2p_node_phase(orgpu_engine* e)
{
  Exchange& x = e->xc;
  if (!e->ctl.nodadt) { launch_node_fused(e->nd, e->d_fsky, e->roww, e->d_cs, e->ctl.iroddl, e->st); e->launches++; return; }
  launch_node_assemble(e->nd, e->d_fsky, e->roww, e->d_cs, e->ctl.iroddl, e->st);
  launch_dtnoda_finalize(e->nd, e->d_cs, 0, e->st);                       // local nodal DT2T
  p2p_dt_push_kernel<<<1, 32, 0, e->st>>>(e->d_cs, x.d_peer_win, x.nranks, x.rank, x.d_xcycle);
  p2p_dt_wait_kernel<<<1, 32, 0, e->st>>>(e->d_cs, x.win, x.nranks, x.d_xcycle, x.d_err, x.timeout_ns);
  launch_node_advance(e->nd, e->d_cs, e->ctl.iroddl, e->st);
  e->launches += 5;
}

int orgpu_run_cycles(orgpu_engine* e, int ncycles)
{
  NEED(e && e->finalized && ncycles >= 0, -1, "orgpu_run_cycles: engine not finalized"); CUDA_OK(cudaSetDevice(e->device));
  { int rc = refresh_batches(e); if (rc) return rc; }
  const int per_cycle = (int)(e->csg.size() + e->bsg.size()) + (e->ctl.nodadt ? 4 : 2) + (e->ipri ? 3 : 0);   // force kernels + dt finalize + node kernel(s) [+ balances]
  NEED(!(e->ipri && e->xc.nranks > 1), -5, "print-cycle balances across domains (frontier-node weights) are outside the built path");
  NEED(!(e->ctl.nodadt && e->xc.nranks > 1 && (!e->xc.p2p || e->profile)), -5, "/DT/NODA across domains needs the peer-memory exchange (orgpu_p2p_connect), unprofiled");
  if (e->xc.nranks > 1 && e->xc.parith_off) {
    Exchange& x = e->xc; NcclApi* N = nccl_api();
    NEED(N && x.comm, -7, "orgpu: /PARITH/OFF exchange requested without orgpu_comm_init");
    CUDA_OK(cudaEventRecord(e->ev0, e->st));
    const int tot = x.xn_ptr.empty() ? 0 : x.xn_ptr.back();
    for (int c = 0; c < ncycles; c++) {
      launch_element_phase(e, 0, nullptr);                                   // forces + the local (dt, type, id)
      launch_node_assemble(e->nd, e->d_fsky, e->roww, e->d_cs, e->ctl.iroddl, e->st); e->launches++;   // ASSPAR4 of the local rows: partial sums
      if (tot) { nodes_pack_kernel<<<(tot + 255) / 256, 256, 0, e->st>>>(e->nd, x.d_xn_nodes, tot, x.d_xn_send); e->launches++; }
      cand_pack_kernel<<<1, 1, 0, e->st>>>(e->d_cs, x.d_cand_send); e->launches++;
      NCCL_OK(N->GroupStart());
      for (size_t k = 0; k < x.nb_rank.size(); k++) {
        const size_t n = x.xn_ptr[k + 1] - x.xn_ptr[k];
        if (n) { NCCL_OK(N->Send(x.d_xn_send + 8 * (size_t)x.xn_ptr[k], 8 * n, ncclDouble, x.nb_rank[k], x.comm, e->st));
                 NCCL_OK(N->Recv(x.d_xn_recv + 8 * (size_t)x.xn_ptr[k], 8 * n, ncclDouble, x.nb_rank[k], x.comm, e->st)); }
      }
      for (int q = 0; q < x.nranks; q++) {
        if (q == x.rank) continue;
        NCCL_OK(N->Send(x.d_cand_send, 4, ncclDouble, q, x.comm, e->st));
        NCCL_OK(N->Recv(x.d_cand_recv + 4 * q, 4, ncclDouble, q, x.comm, e->st));
      }
      NCCL_OK(N->GroupEnd());
      CUDA_OK(cudaMemcpyAsync(x.d_cand_recv + 4 * x.rank, x.d_cand_send, 32, cudaMemcpyDeviceToDevice, e->st));
      for (size_t k = 0; k < x.nb_rank.size(); k++) {                        // SPMD_EXCH_A :517-528: neighbour after neighbour, rank order
        const int n = x.xn_ptr[k + 1] - x.xn_ptr[k];
        if (n) { nodes_add_kernel<<<(n + 255) / 256, 256, 0, e->st>>>(e->nd, x.d_xn_nodes + x.xn_ptr[k], n, x.d_xn_recv + 8 * (size_t)x.xn_ptr[k]); e->launches++; }
      }
      cand_fold_kernel<<<1, 32, 0, e->st>>>(e->d_cs, x.d_cand_recv, x.nranks); e->launches++;
      launch_node_advance(e->nd, e->d_cs, e->ctl.iroddl, e->st); e->launches++;
    }
    CUDA_OK(cudaEventRecord(e->ev1, e->st)); e->ev_recorded = true;
    CUDA_OK(cudaGetLastError());
    return 0;
  }
  if (e->xc.nranks > 1 && e->xc.p2p && !e->profile) {
    // one process per GPU, peer-memory exchange: the whole cycle (forces, dt fold, push, wait+unpack, gather+update)
    // is one CUDA graph, replayed ncycles times with no host involvement and no library call
    const int per = (int)(e->csg.size() + e->bsg.size()) + (e->ctl.nodadt ? 8 : 4);
    if (!e->gexec) {
      cudaGraph_t g;
      CUDA_OK(cudaStreamBeginCapture(e->st, cudaStreamCaptureModeThreadLocal));
      const long long l0 = e->launches;
      launch_element_phase(e, 0, nullptr);
      p2p_exchange_on_stream(e);
      p2p_node_phase(e);
      e->launches = l0;                      // the capture pass did not execute
      CUDA_OK(cudaStreamEndCapture(e->st, &g));
      CUDA_OK(cudaGraphInstantiate(&e->gexec, g, 0));
      CUDA_OK(cudaGraphDestroy(g));
    }
    CUDA_OK(cudaEventRecord(e->ev0, e->st));
    for (int c = 0; c < ncycles; c++) CUDA_OK(cudaGraphLaunch(e->gexec, e->st));
    CUDA_OK(cudaEventRecord(e->ev1, e->st)); e->ev_recorded = true;
    e->launches += (long long)per * ncycles;
    CUDA_OK(cudaGetLastError());
    return 0;
  }
  if (e->xc.nranks > 1) {
    // one process per GPU: element phase writes the local dt candidate only; the exchange folds all
    // ranks' candidates and advances the clock; then the ordered gather + nodal update
    const bool prof = e->profile != 0;
    if (prof) for (int k = 0; k < 3; k++) { e->prof_ms[k] = 0; e->prof_n[k] = 0; }
    CUDA_OK(cudaEventRecord(e->ev0, e->st));
    const int chunk = 64;
    for (int c0 = 0; c0 < ncycles; c0 += chunk) {
      const int nc = (ncycles - c0 < chunk) ? ncycles - c0 : chunk;
      size_t evi = 0;
      for (int c = 0; c < nc; c++) {
        launch_element_phase(e, 0, prof ? &evi : nullptr);
        if (e->xc.p2p) p2p_exchange_on_stream(e);
        else { int rc = exchange_on_stream(e, true); if (rc) return rc; }
        if (prof) cudaEventRecord(get_event(e, evi++), e->st);
        launch_node_fused(e->nd, e->d_fsky, e->roww, e->d_cs, e->ctl.iroddl, e->st); e->launches++;
        if (prof) cudaEventRecord(get_event(e, evi++), e->st);
      }
      if (prof) {
        CUDA_OK(cudaStreamSynchronize(e->st));
        size_t k = 0;
        for (int c = 0; c < nc; c++) {
          for (size_t s = 0; s < e->csg.size(); s++, k += 2) { float ms; cudaEventElapsedTime(&ms, e->evpool[k], e->evpool[k + 1]); e->prof_ms[1] += ms; e->prof_n[1]++; }
          for (size_t s = 0; s < e->bsg.size(); s++, k += 2) { float ms; cudaEventElapsedTime(&ms, e->evpool[k], e->evpool[k + 1]); e->prof_ms[0] += ms; e->prof_n[0]++; }
          { float ms; cudaEventElapsedTime(&ms, e->evpool[k], e->evpool[k + 1]); e->prof_ms[2] += ms; e->prof_n[2]++; k += 2; }
        }
      }
    }
    CUDA_OK(cudaEventRecord(e->ev1, e->st)); e->ev_recorded = true;
    if (prof) { CUDA_OK(cudaStreamSynchronize(e->st)); float ms; CUDA_OK(cudaEventElapsedTime(&ms, e->ev0, e->ev1)); e->last_run_ms = ms; }
    CUDA_OK(cudaGetLastError());
    return 0;
  }
  if (e->profile) {
    // un-graphed, one event pair around every launch: per-class device time on the launching stream
    for (int k = 0; k < 3; k++) { e->prof_ms[k] = 0; e->prof_n[k] = 0; }
    const int chunk = 64;
    CUDA_OK(cudaEventRecord(e->ev0, e->st));
    for (int c0 = 0; c0 < ncycles; c0 += chunk) {
      const int nc = (ncycles - c0 < chunk) ? ncycles - c0 : chunk;
      size_t evi = 0;
      for (int c = 0; c < nc; c++) {
        launch_element_phase(e, e->ctl.nodadt ? 0 : 1, &evi);
        cudaEventRecord(get_event(e, evi++), e->st);
        launch_node_phase(e);
        cudaEventRecord(get_event(e, evi++), e->st);
      }
      CUDA_OK(cudaStreamSynchronize(e->st));
      size_t k = 0;
      for (int c = 0; c < nc; c++) {
        for (size_t s = 0; s < e->csg.size(); s++, k += 2) { float ms; cudaEventElapsedTime(&ms, e->evpool[k], e->evpool[k + 1]); e->prof_ms[1] += ms; e->prof_n[1]++; }
        for (size_t s = 0; s < e->bsg.size(); s++, k += 2) { float ms; cudaEventElapsedTime(&ms, e->evpool[k], e->evpool[k + 1]); e->prof_ms[0] += ms; e->prof_n[0]++; }
        { float ms; cudaEventElapsedTime(&ms, e->evpool[k], e->evpool[k + 1]); e->prof_ms[2] += ms; e->prof_n[2]++; k += 2; }
      }
    }
    CUDA_OK(cudaEventRecord(e->ev1, e->st)); e->ev_recorded = true;
    CUDA_OK(cudaStreamSynchronize(e->st));
    float ms; CUDA_OK(cudaEventElapsedTime(&ms, e->ev0, e->ev1)); e->last_run_ms = ms;
    CUDA_OK(cudaGetLastError());
    return 0;
  }
  if (!e->gexec) {                         // capture one fused cycle once, replay it ncycles times
    cudaGraph_t g;
    CUDA_OK(cudaStreamBeginCapture(e->st, cudaStreamCaptureModeThreadLocal));
    const long long l0 = e->launches;
    launch_element_phase(e, e->ctl.nodadt ? 0 : 1, nullptr);
    launch_node_phase(e);
    e->launches = l0;                      // the capture pass did not execute
    CUDA_OK(cudaStreamEndCapture(e->st, &g));
    CUDA_OK(cudaGraphInstantiate(&e->gexec, g, 0));
    CUDA_OK(cudaGraphDestroy(g));
  }
  CUDA_OK(cudaEventRecord(e->ev0, e->st));
  for (int c = 0; c < ncycles; c++) CUDA_OK(cudaGraphLaunch(e->gexec, e->st));
  CUDA_OK(cudaEventRecord(e->ev1, e->st)); e->ev_recorded = true;
  e->launches += (long long)per_cycle * ncycles;
  CUDA_OK(cudaGetLastError());
  return 0;
}

int orgpu_synchronize(orgpu_engine* e)
{
  NEED(e, -1, "null handle"); CUDA_OK(cudaSetDevice(e->device));
  CUDA_OK(cudaStreamSynchronize(e->st));
  { int rc = check_abort(e); if (rc) return rc; }
  if (!e->profile && e->ev_recorded) { float ms = 0; if (cudaEventElapsedTime(&ms, e->ev0, e->ev1) == cudaSuccess) e->last_run_ms = ms; else cudaGetLastError(); }
  return 0;
}

int orgpu_get_time(orgpu_engine* e, double out[5], int iout[3])
{
  NEED(e, -1, "null handle"); CUDA_OK(cudaSetDevice(e->device));
  CUDA_OK(cudaStreamSynchronize(e->st));
  { int rc = check_abort(e); if (rc) return rc; }
  CycleState cs; CUDA_OK(cudaMemcpy(&cs, e->d_cs, sizeof cs, cudaMemcpyDeviceToHost));
  out[0] = cs.tt; out[1] = cs.dt1; out[2] = cs.dt2; out[3] = cs.dt12; out[4] = cs.dt2t;
  iout[0] = cs.neltst; iout[1] = cs.ityptst; iout[2] = (int)cs.ncycle;
  return 0;
}

int orgpu_download_nodes(orgpu_engine* e, double* X, double* V, double* VR, double* D,
                         double* A, double* AR, double* STIFN, double* STIFR)
{
  NEED(e, -1, "null handle"); CUDA_OK(cudaSetDevice(e->device));
  CUDA_OK(cudaStreamSynchronize(e->st));
  { int rc = check_abort(e); if (rc) return rc; }
  const size_t n = e->numnod;
  if (X) { if (download4to3(e, e->nd.pos, X)) return -100; }
  if (V) { if (download4to3(e, e->nd.vel, V)) return -100; }
  if (VR) { if (e->nd.rot) { if (download4to3(e, e->nd.rot, VR)) return -100; } else memset(VR, 0, 24 * n); }
  if (D) CUDA_OK(cudaMemcpy(D, e->nd.D, 24 * n, cudaMemcpyDeviceToHost));
  if (A) CUDA_OK(cudaMemcpy(A, e->nd.A, 24 * n, cudaMemcpyDeviceToHost));
  if (AR) CUDA_OK(cudaMemcpy(AR, e->nd.AR, 24 * n, cudaMemcpyDeviceToHost));
  if (STIFN) CUDA_OK(cudaMemcpy(STIFN, e->nd.STIFN, 8 * n, cudaMemcpyDeviceToHost));
  if (STIFR) CUDA_OK(cudaMemcpy(STIFR, e->nd.STIFR, 8 * n, cudaMemcpyDeviceToHost));
  return 0;
}

int orgpu_download_fsky(orgpu_engine* e, double* fsky)
{
  NEED(e && e->finalized, -1, "engine not finalized"); CUDA_OK(cudaSetDevice(e->device));
  CUDA_OK(cudaStreamSynchronize(e->st));
  const size_t L = e->lsky;
  if (e->roww == 8) { CUDA_OK(cudaMemcpy(fsky, e->d_fsky, 64 * L, cudaMemcpyDeviceToHost)); return 0; }
  std::vector<double> h(4 * L);
  CUDA_OK(cudaMemcpy(h.data(), e->d_fsky, 32 * L, cudaMemcpyDeviceToHost));
  for (size_t k = 0; k < L; k++) { double* f = fsky + 8 * k; f[0] = h[4 * k]; f[1] = h[4 * k + 1]; f[2] = h[4 * k + 2]; f[3] = f[4] = f[5] = 0; f[6] = h[4 * k + 3]; f[7] = 0; }
  return 0;
}

static int solid_state_xfer(orgpu_engine* e, int field, double* buf, bool up)
{
  NEED(e && e->finalized && buf, -1, "engine not finalized"); CUDA_OK(cudaSetDevice(e->device));
  CUDA_OK(cudaStreamSynchronize(e->st));
  { int rc = check_abort(e); if (rc) return rc; }
  const size_t NE = e->numels;
  for (auto& S : e->bsg) {
    const BrickSG& d = S.d; int w0 = 0, nc = 1; double* base = d.slab; int nw = d.nw;
    switch (field) { case 0: w0 = BW_SIG; nc = 6; break; case 1: w0 = BW_EINT; break; case 2: w0 = BW_RHO; break; case 3: w0 = BW_QVIS; break;
                     case 4: w0 = BW_PLA; break; case 5: w0 = BW_EPSD; break; case 6: w0 = d.w_vol; break; case 7: w0 = BW_OFF; break;
                     case 8: w0 = d.w_temp; break; case 9: base = d.smstr; nw = 21; w0 = 0; nc = 21; break;
                     case 10: w0 = d.w_stra; nc = 6; if (w0 < 0) continue; break; case 11: w0 = d.w_wpla; if (w0 < 0) continue; break;
                     case 12: w0 = d.w_sigb; nc = 6; if (w0 < 0) continue; break; case 13: w0 = d.w_dfmax; if (w0 < 0) continue; break; default: FAIL(-1, "unknown solid field %d", field); }
    if (w0 < 0) { if (!up) for (int i = 0; i < d.ne; i++) buf[S.first_elem + i] = d.mat.tini; continue; }   // no temperature buffer
    for (int k = 0; k < nc; k++)
      CUDA_OK(up ? slab_upload_word(base, nw, w0 + k, d.ne, buf + k * NE + S.first_elem)
                 : slab_download_word(base, nw, w0 + k, d.ne, buf + k * NE + S.first_elem));
  }
  return 0;
}
int orgpu_download_solid_state(orgpu_engine* e, int field, double* out) { return solid_state_xfer(e, field, out, false); }
int orgpu_upload_solid_state(orgpu_engine* e, int field, const double* in) { return solid_state_xfer(e, field, const_cast<double*>(in), true); }

int orgpu_download_shell_state(orgpu_engine* e, int field, double* out)
{
  NEED(e && e->finalized && out, -1, "engine not finalized"); CUDA_OK(cudaSetDevice(e->device));
  CUDA_OK(cudaStreamSynchronize(e->st));
  { int rc = check_abort(e); if (rc) return rc; }
  return shell_state_xfer(e->csg, e->numelc, field, out, false);
}
int orgpu_upload_shell_state(orgpu_engine* e, int field, const double* in)
{
  NEED(e && e->finalized && in, -1, "engine not finalized"); CUDA_OK(cudaSetDevice(e->device));
  CUDA_OK(cudaStreamSynchronize(e->st));
  return shell_state_xfer(e->csg, e->numelc, field, const_cast<double*>(in), true);
}
int orgpu_download_sh3n_state(orgpu_engine* e, int field, double* out)
{
  NEED(e && e->finalized && out, -1, "engine not finalized"); CUDA_OK(cudaSetDevice(e->device));
  CUDA_OK(cudaStreamSynchronize(e->st));
  return shell_state_xfer(e->csg, e->numeltg, field, out, false, true);
}
int orgpu_upload_sh3n_state(orgpu_engine* e, int field, const double* in)
{
  NEED(e && e->finalized && in, -1, "engine not finalized"); CUDA_OK(cudaSetDevice(e->device));
  CUDA_OK(cudaStreamSynchronize(e->st));
  return shell_state_xfer(e->csg, e->numeltg, field, const_cast<double*>(in), true, true);
}
/* (the LAW36 VARTMP cursors are not part of a hand-over: they are forward-only walks on the monotone plastic strain and re-find
 * their segment from zero in the first cycle after an upload) */
int orgpu_set_time(orgpu_engine* e, double tt, double dt2, double dt2old, long long ncycle)
{
  NEED(e, -1, "null handle"); CUDA_OK(cudaSetDevice(e->device));
  CUDA_OK(cudaStreamSynchronize(e->st));
  CycleState cs; CUDA_OK(cudaMemcpy(&cs, e->d_cs, sizeof cs, cudaMemcpyDeviceToHost));
  cs.tt = tt; cs.tt0 = tt; cs.dt2 = dt2; cs.dt2old = dt2old; cs.ncycle = ncycle;
  CUDA_OK(cudaMemcpy(e->d_cs, &cs, sizeof cs, cudaMemcpyHostToDevice));
  return 0;
}

// Host-owned nodal arrays, one call per step: X, V, VR in (pinned host memory), ncycles on the device, X, V, VR out.  The three
// uploads are queued back to back on the stream (one staging buffer each: no host synchronisation between them), each followed by
// its 24 -> 32-byte record kernel; the downloads likewise.  The call is bound by the PCIe link: 24 bytes per node and array each
// way (bench.py reports the measured link rate beside it) -- the cycle needs every node before it starts and the host needs
// the result before its next call, so nothing of the transfer can hide behind the compute.
static int step_host_impl(orgpu_engine* e, const double* X, const double* V, const double* VR, int ncycles, double* Xout, double* Vout, double* VRout)
{
  NEED(e && e->finalized, -1, "engine not finalized"); CUDA_OK(cudaSetDevice(e->device));
  const int n = e->numnod; const int nb = (n + 255) / 256;
  if (e->nd.rot && (VR || VRout) && !e->d_stage3c) { if (dev_alloc(&e->d_stage3c, 3 * (size_t)n)) return -100; }
  if (X) CUDA_OK(cudaMemcpyAsync(e->d_stage3a, X, 24 * (size_t)n, cudaMemcpyHostToDevice, e->st));
  if (V) CUDA_OK(cudaMemcpyAsync(e->d_stage3b, V, 24 * (size_t)n, cudaMemcpyHostToDevice, e->st));
  if (VR && e->nd.rot) CUDA_OK(cudaMemcpyAsync(e->d_stage3c, VR, 24 * (size_t)n, cudaMemcpyHostToDevice, e->st));
  if (X) { pack3to4_kernel<<<nb, 256, 0, e->st>>>(e->d_stage3a, e->nd.pos, n); e->launches++; }
  if (V) { pack3to4_kernel<<<nb, 256, 0, e->st>>>(e->d_stage3b, e->nd.vel, n); e->launches++; }
  if (VR && e->nd.rot) { pack3to4_kernel<<<nb, 256, 0, e->st>>>(e->d_stage3c, e->nd.rot, n); e->launches++; }
  { int rc = orgpu_run_cycles(e, ncycles); if (rc) return rc; }
  if (Xout) { unpack4to3_kernel<<<nb, 256, 0, e->st>>>(e->nd.pos, e->d_stage3a, n); e->launches++; }
  if (Vout) { unpack4to3_kernel<<<nb, 256, 0, e->st>>>(e->nd.vel, e->d_stage3b, n); e->launches++; }
  if (VRout && e->nd.rot) { unpack4to3_kernel<<<nb, 256, 0, e->st>>>(e->nd.rot, e->d_stage3c, n); e->launches++; }
  if (Xout) CUDA_OK(cudaMemcpyAsync(Xout, e->d_stage3a, 24 * (size_t)n, cudaMemcpyDeviceToHost, e->st));
  if (Vout) CUDA_OK(cudaMemcpyAsync(Vout, e->d_stage3b, 24 * (size_t)n, cudaMemcpyDeviceToHost, e->st));
  if (VRout && e->nd.rot) CUDA_OK(cudaMemcpyAsync(VRout, e->d_stage3c, 24 * (size_t)n, cudaMemcpyDeviceToHost, e->st));
  CUDA_OK(cudaStreamSynchronize(e->st));
  return check_abort(e);
}
int orgpu_step_host(orgpu_engine* e, const double* X, const double* V, const double* VR, int ncycles, double* Xout, double* Vout)
{ return step_host_impl(e, X, V, VR, ncycles, Xout, Vout, nullptr); }
int orgpu_step_host_rot(orgpu_engine* e, const double* X, const double* V, const double* VR, int ncycles, double* Xout, double* Vout, double* VRout)
{ return step_host_impl(e, X, V, VR, ncycles, Xout, Vout, VRout); }

// ---- orgpu_forces_host ------------------------------------------------------------------------------------------------
// The cycle of the reference's own -gpu ABI (shell_internal_forces.F90:106-190: X, V, VR up, internal forces, the assembled
// nodal forces down; the host integrates) moves 24 x 3 bytes per node up and 64 down and is bound by the PCIe link when the
// three steps run one after the other.  Here they are a pipeline over K chunks of the node range:
//   stream `up`  : chunk j of X, V, VR host -> staging, 24 -> 32-byte records                                  (H2D engine)
//   main stream  : after upload j, the state tiles all of whose nodes have arrived (batched launches over a per-chunk
//                  CTA map: common.cuh cta_work), then the dt fold
//   stream `down`: after element group j, ASSPAR4 of the node chunks no later tile touches, 8 doubles per node, -> host
//                                                                                                               (D2H engine)
// so the link runs in both directions at once and the kernels hide behind it (C2 on a B200, PCIe 5 x16: 3.02 -> 2.27 ms per
// cycle; what is left is the copy engines' own cost of chunked copies in both directions at once -- 24 + 8 pieces of 72 MB move
// at 35.7 GB/s each way against 49.8 for two single copies -- and the one-chunk lag of the assembly behind the uploads).  Any node / element numbering is handled
// (a tile waits for its highest node, a node chunk for the last tile that touches its range); a mesh numbered with locality --
// what mesh generators and the Starter's domain decomposition produce -- overlaps almost completely, a random numbering
// degrades to upload, compute, download in sequence.  Results are bit-identical to orgpu_forces_phase + orgpu_assemble.
static void pipe_free(orgpu_engine* e)
{
  auto& hp = e->hp;
  for (void* p : hp.owned) cudaFree(p);
  hp.owned.clear(); hp.grp.clear(); hp.solo_c.clear(); hp.solo_b.clear(); hp.done.clear();
  for (auto ev : hp.ev_up) cudaEventDestroy(ev); for (auto ev : hp.ev_el) cudaEventDestroy(ev); for (auto ev : hp.ev_as) cudaEventDestroy(ev); for (auto ev : hp.ev_dn) cudaEventDestroy(ev);
  hp.ev_up.clear(); hp.ev_el.clear(); hp.ev_as.clear(); hp.ev_dn.clear();
  if (hp.ev_start) { cudaEventDestroy(hp.ev_start); hp.ev_start = nullptr; } if (hp.ev_down) { cudaEventDestroy(hp.ev_down); hp.ev_down = nullptr; }
  if (hp.up) { cudaStreamDestroy(hp.up); hp.up = nullptr; } if (hp.down) { cudaStreamDestroy(hp.down); hp.down = nullptr; }
  if (hp.asmb) { cudaStreamDestroy(hp.asmb); hp.asmb = nullptr; }
  if (hp.d_f8) { cudaFree(hp.d_f8); hp.d_f8 = nullptr; } if (hp.h_cs) { cudaFreeHost(hp.h_cs); hp.h_cs = nullptr; }
  hp.K = 0; hp.desc_gen = -1;
}

static int pipe_build(orgpu_engine* e)
{
  auto& hp = e->hp;
  if (hp.K > 0 && hp.desc_gen == e->desc_gen) return 0;
  pipe_free(e);
  const int n = e->numnod;
  const char* kc = getenv("ORGPU_PIPE_CHUNKS"); int K = kc ? atoi(kc) : 8;     // 4 / 8 / 16 / 32 measured on C2: 2.36 / 2.27 / 2.37 / 2.54 ms (one after the other: 3.02)
  if (K < 1) K = 1; if (K > 64) K = 64; if (K > (n + 255) / 256) K = (n + 255) / 256; if (K < 1) K = 1;
  hp.nb.assign(K + 1, 0);
  for (int k = 0; k <= K; k++) { long long b = (long long)n * k / K; b = (b + 127) / 128 * 128; if (b > n || k == K) b = n; hp.nb[k] = (int)b; }
  auto chunk_of = [&](int node) { int c = (int)(std::upper_bound(hp.nb.begin(), hp.nb.end(), node) - hp.nb.begin()) - 1; return c < 0 ? 0 : (c >= K ? K - 1 : c); };
  hp.grp.assign(K, {}); hp.solo_c.assign(K, {}); hp.solo_b.assign(K, {}); hp.done.assign(K, {});
  std::vector<int> cout(K); for (int k = 0; k < K; k++) cout[k] = k;
  auto touch = [&](int cmin, int cmax) { for (int k = cmin; k <= cmax; k++) if (cout[k] < cmax) cout[k] = cmax; };
  // chunk range of every state tile
  struct TileC { int cmin, cmax; };
  auto tile_chunks = [&](const std::vector<int>& ix, int stride, int nn, int first, int ne, std::vector<TileC>& out) {
    const int nt = (ne + ORGPU_TILE - 1) / ORGPU_TILE; out.assign(nt, TileC{K - 1, 0});
    for (int i = 0; i < ne; i++) {
      const int* r = &ix[(size_t)stride * (first + i)]; TileC& t = out[i / ORGPU_TILE];
      for (int c = 1; c <= nn; c++) { const int ch = chunk_of(r[c] - 1); if (ch < t.cmin) t.cmin = ch; if (ch > t.cmax) t.cmax = ch; }
    }
  };
  std::vector<std::vector<TileC>> tc_c(e->csg.size()), tc_b(e->bsg.size());
  for (size_t k = 0; k < e->csg.size(); k++) {
    const ShellSGHost& S = e->csg[k];
    if (S.sh3n) tile_chunks(e->ixtg, 6, 3, S.first_elem, S.d.ne, tc_c[k]); else tile_chunks(e->ixc, 7, 4, S.first_elem, S.d.ne, tc_c[k]);
  }
  for (size_t k = 0; k < e->bsg.size(); k++) tile_chunks(e->ixs, 11, 8, e->bsg[k].first_elem, e->bsg[k].d.ne, tc_b[k]);
  // batched launches: one device table of descriptors per kernel variant, one CTA map per (variant, chunk)
  for (int brick = 0; brick < 2; brick++) {
    const int nv = brick ? (int)BRV_COUNT : (int)SHV_COUNT; const size_t nsg = brick ? e->bsg.size() : e->csg.size();
    std::vector<char> batched(nsg, 0);
    for (int v = 0; v < nv; v++) {
      std::vector<int> sgs;
      for (size_t k = 0; k < nsg; k++) if ((brick ? brick_tab_variant(e->bsg[k].d) : shell_tab_variant(e->csg[k])) == v) sgs.push_back((int)k);
      if (sgs.empty()) continue;
      void* d_tab = nullptr; size_t bytes = 0;
      if (brick) { std::vector<BrickSG> tab; for (int k : sgs) { tab.push_back(e->bsg[k].d); bytes = std::max(bytes, (size_t)e->bsg[k].d.nw * ORGPU_TILE * 8); }
                   CUDA_OK(cudaMalloc(&d_tab, sizeof(BrickSG) * tab.size())); hp.owned.push_back(d_tab);
                   CUDA_OK(cudaMemcpy(d_tab, tab.data(), sizeof(BrickSG) * tab.size(), cudaMemcpyHostToDevice)); }
      else       { std::vector<ShellSG> tab; for (int k : sgs) { tab.push_back(e->csg[k].d); bytes = std::max(bytes, (size_t)e->csg[k].d.nw * ORGPU_TILE * 8); }
                   CUDA_OK(cudaMalloc(&d_tab, sizeof(ShellSG) * tab.size())); hp.owned.push_back(d_tab);
                   CUDA_OK(cudaMemcpy(d_tab, tab.data(), sizeof(ShellSG) * tab.size(), cudaMemcpyHostToDevice)); }
      std::vector<std::vector<int2>> maps(K);
      for (size_t j = 0; j < sgs.size(); j++) {
        const auto& tc = brick ? tc_b[sgs[j]] : tc_c[sgs[j]]; batched[sgs[j]] = 1;
        for (size_t t = 0; t < tc.size(); t++) { maps[tc[t].cmax].push_back(make_int2((int)j, (int)t)); touch(tc[t].cmin, tc[t].cmax); }
      }
      for (int c = 0; c < K; c++) if (!maps[c].empty()) {
        int2* d_map = nullptr; CUDA_OK(cudaMalloc((void**)&d_map, sizeof(int2) * maps[c].size())); hp.owned.push_back(d_map);
        CUDA_OK(cudaMemcpy(d_map, maps[c].data(), sizeof(int2) * maps[c].size(), cudaMemcpyHostToDevice));
        hp.grp[c].push_back(orgpu_engine::PipeLaunch{brick != 0, v, d_tab, d_map, (int)maps[c].size(), bytes});
      }
    }
    for (size_t k = 0; k < nsg; k++) if (!batched[k]) {       // no table-driven variant of this kernel: the whole super-group when its last node is there
      const auto& tc = brick ? tc_b[k] : tc_c[k]; int cmin = K - 1, cmax = 0;
      for (auto& t : tc) { cmin = std::min(cmin, t.cmin); cmax = std::max(cmax, t.cmax); }
      if (tc.empty()) continue;
      touch(cmin, cmax); (brick ? hp.solo_b : hp.solo_c)[cmax].push_back((int)k);
    }
  }
  for (int k = 0; k < K; k++) hp.done[cout[k]].push_back(k);
  { int lo = 0, hi = 0; cudaDeviceGetStreamPriorityRange(&lo, &hi);      // the transfer-side kernels go first when an SM slot frees up
    CUDA_OK(cudaStreamCreateWithPriority(&hp.up, cudaStreamNonBlocking, hi)); CUDA_OK(cudaStreamCreateWithPriority(&hp.down, cudaStreamNonBlocking, hi));
    CUDA_OK(cudaStreamCreateWithPriority(&hp.asmb, cudaStreamNonBlocking, hi)); }
  hp.ev_up.resize(K); hp.ev_el.resize(K); hp.ev_as.resize(K);
  const unsigned evf = getenv("ORGPU_PIPE_DEBUG") ? cudaEventDefault : cudaEventDisableTiming;
  hp.ev_dn.resize(K);
  for (int k = 0; k < K; k++) { CUDA_OK(cudaEventCreateWithFlags(&hp.ev_up[k], evf)); CUDA_OK(cudaEventCreateWithFlags(&hp.ev_el[k], evf));
                                CUDA_OK(cudaEventCreateWithFlags(&hp.ev_as[k], evf)); CUDA_OK(cudaEventCreateWithFlags(&hp.ev_dn[k], evf)); }
  CUDA_OK(cudaEventCreateWithFlags(&hp.ev_start, evf)); CUDA_OK(cudaEventCreateWithFlags(&hp.ev_down, evf));
  CUDA_OK(cudaMalloc((void**)&hp.d_f8, 64 * (size_t)(n + 1))); CUDA_OK(cudaMallocHost((void**)&hp.h_cs, sizeof(CycleState)));
  if (e->nd.rot && !e->d_stage3c) { if (dev_alloc(&e->d_stage3c, 3 * (size_t)n)) return -100; }
  hp.K = K; hp.desc_gen = e->desc_gen;
  return 0;
}

int orgpu_forces_host(orgpu_engine* e, const double* X, const double* V, const double* VR, double dt1, double* F8, double* dt2t, int* neltst, int* ityptst)
{
  NEED(e && e->finalized && X && V && F8, -1, "orgpu_forces_host: engine not finalized / null array"); CUDA_OK(cudaSetDevice(e->device));
  NEED(e->xc.nranks <= 1 && !e->ctl.nodadt && !e->ipri, -5, "orgpu_forces_host: single domain, element time step, no print cycle (use the phased calls otherwise)");
  NEED(!e->nd.rot || VR, -1, "orgpu_forces_host: the model has rotational dofs, VR is required");
  { int rc = pipe_build(e); if (rc) return rc; }
  auto& hp = e->hp; const int K = hp.K; const bool rot = e->nd.rot != nullptr;
  static const bool dbg = getenv("ORGPU_PIPE_DEBUG") != nullptr;
  timespec ts0; clock_gettime(CLOCK_MONOTONIC, &ts0);
  DevNodes ndi = e->nd; ndi.FEXT = nullptr; ndi.MEXT = nullptr;          // internal forces: the caller owns the loads
  CUDA_OK(cudaEventRecord(hp.ev_start, e->st));                         // behind whatever the handle still has in flight
  CUDA_OK(cudaStreamWaitEvent(hp.up, hp.ev_start, 0)); CUDA_OK(cudaStreamWaitEvent(hp.down, hp.ev_start, 0)); CUDA_OK(cudaStreamWaitEvent(hp.asmb, hp.ev_start, 0));
  launch_set_dt(e->d_cs, dt1, 0, 0, 0, e->st); e->launches++;
  e->fa.fused = 0;
  // (Letting the kernels themselves read X, V, VR from / store the rows into mapped host memory instead of using the copy
  // engines was measured: 3.6 ms (stores only) and 4.1 ms (both) per cycle on C2 against 2.3 ms with the engines.)
  // `up` and `down` carry copies only (back to back on their copy engines: a kernel between two copies of one stream costs the
  // engine a dependency round trip each way); the record kernels run on the main stream, the assembly on a stream of its own
  for (int j = 0; j < K; j++) {
    const int n0 = hp.nb[j], cnt = hp.nb[j + 1] - n0;
    if (cnt > 0) {
      CUDA_OK(cudaMemcpyAsync(e->d_stage3a + 3 * (size_t)n0, X + 3 * (size_t)n0, 24 * (size_t)cnt, cudaMemcpyHostToDevice, hp.up));
      CUDA_OK(cudaMemcpyAsync(e->d_stage3b + 3 * (size_t)n0, V + 3 * (size_t)n0, 24 * (size_t)cnt, cudaMemcpyHostToDevice, hp.up));
      if (rot) CUDA_OK(cudaMemcpyAsync(e->d_stage3c + 3 * (size_t)n0, VR + 3 * (size_t)n0, 24 * (size_t)cnt, cudaMemcpyHostToDevice, hp.up));
    }
    CUDA_OK(cudaEventRecord(hp.ev_up[j], hp.up));
  }
  for (int j = 0; j < K; j++) {
    CUDA_OK(cudaStreamWaitEvent(e->st, hp.ev_up[j], 0));
    { const int n0 = hp.nb[j], cnt = hp.nb[j + 1] - n0;
      if (cnt > 0) { pack3to4x3_kernel<<<(cnt + 255) / 256, 256, 0, e->st>>>(e->d_stage3a + 3 * (size_t)n0, e->nd.pos + n0, e->d_stage3b + 3 * (size_t)n0, e->nd.vel + n0,
                                                                           rot ? e->d_stage3c + 3 * (size_t)n0 : nullptr, rot ? e->nd.rot + n0 : nullptr, cnt); e->launches++; } }
    for (const auto& L : hp.grp[j]) {
      if (L.brick) launch_brick_forces_tab(L.variant, (const BrickSG*)L.d_tab, L.d_map, L.ntile, L.bytes, e->nd, e->d_fsky, e->roww, e->d_cs, e->db, e->st);
      else         launch_shell_forces_tab(L.variant, (const ShellSG*)L.d_tab, L.d_map, L.ntile, L.bytes, e->nd, e->d_fsky, e->d_cs, e->db, e->st);
      e->launches++;
    }
    for (int k : hp.solo_c[j]) launch_sg_shell(e, e->csg[k], e->st);
    for (int k : hp.solo_b[j]) launch_sg_brick(e, e->bsg[k], e->st);
    if (hp.done[j].empty()) continue;
    CUDA_OK(cudaEventRecord(hp.ev_el[j], e->st));
    CUDA_OK(cudaStreamWaitEvent(hp.asmb, hp.ev_el[j], 0));
    for (int k : hp.done[j]) {
      const int n0 = hp.nb[k], n1 = hp.nb[k + 1]; if (n1 <= n0) continue;
      const int nblk = (n1 - n0 + ORGPU_NODE_BLOCK - 1) / ORGPU_NODE_BLOCK;
      if (e->roww == 4) node_forces8_kernel<4><<<nblk, ORGPU_NODE_BLOCK, 0, hp.asmb>>>(ndi, e->d_fsky, e->d_cs, e->ctl.iroddl, n0, n1, hp.d_f8);
      else              node_forces8_kernel<8><<<nblk, ORGPU_NODE_BLOCK, 0, hp.asmb>>>(ndi, e->d_fsky, e->d_cs, e->ctl.iroddl, n0, n1, hp.d_f8);
      e->launches++;
      CUDA_OK(cudaEventRecord(hp.ev_as[k], hp.asmb));
      CUDA_OK(cudaStreamWaitEvent(hp.down, hp.ev_as[k], 0));
      CUDA_OK(cudaMemcpyAsync(F8 + 8 * (size_t)n0, hp.d_f8 + 8 * (size_t)n0, 64 * (size_t)(n1 - n0), cudaMemcpyDeviceToHost, hp.down));
      if (dbg) cudaEventRecord(hp.ev_dn[k], hp.down);
    }
  }
  element_finalize_kernel<<<1, ORGPU_FINALIZE_BLOCK, 0, e->st>>>(e->d_cs, e->db, e->fa); e->launches++;
  CUDA_OK(cudaMemcpyAsync(hp.h_cs, e->d_cs, sizeof(CycleState), cudaMemcpyDeviceToHost, e->st));
  CUDA_OK(cudaEventRecord(hp.ev_down, hp.down)); CUDA_OK(cudaStreamWaitEvent(e->st, hp.ev_down, 0));    // the handle's stream is the one callers synchronise
  timespec ts1; if (dbg) clock_gettime(CLOCK_MONOTONIC, &ts1);
  CUDA_OK(cudaStreamSynchronize(e->st));
  if (dbg) { timespec ts2; clock_gettime(CLOCK_MONOTONIC, &ts2);
    fprintf(stderr, "forces_host: submit %.3f ms, wait %.3f ms\n", (ts1.tv_sec - ts0.tv_sec) * 1e3 + (ts1.tv_nsec - ts0.tv_nsec) * 1e-6, (ts2.tv_sec - ts1.tv_sec) * 1e3 + (ts2.tv_nsec - ts1.tv_nsec) * 1e-6);
    for (int k = 0; k < K; k++) { float a = -1, b = -1, c = -1, d = -1; cudaEventElapsedTime(&a, hp.ev_start, hp.ev_up[k]); cudaEventElapsedTime(&b, hp.ev_start, hp.ev_el[k]);
      cudaEventElapsedTime(&c, hp.ev_start, hp.ev_as[k]); cudaEventElapsedTime(&d, hp.ev_start, hp.ev_dn[k]);
      fprintf(stderr, "  chunk %2d: uploaded %.3f  elements %.3f  assembled %.3f  downloaded %.3f ms\n", k, a, b, c, d); }
    cudaGetLastError(); }
  CUDA_OK(cudaGetLastError());
  if (dt2t) *dt2t = hp.h_cs->dt2t; if (neltst) *neltst = hp.h_cs->neltst; if (ityptst) *ityptst = hp.h_cs->ityptst;
  return check_abort(e);
}

// ---- domain exchange ------------------------------------------------------------------------------

static int rows_tmp(orgpu_engine* e, size_t n)
{
  Exchange& x = e->xc;
  if (n <= x.tmp_cap) return 0;
  if (x.d_slots_tmp) cudaFree(x.d_slots_tmp); if (x.d_rows_tmp) cudaFree(x.d_rows_tmp);
  x.tmp_cap = n + n / 2 + 64;
  CUDA_OK(cudaMalloc((void**)&x.d_slots_tmp, 4 * x.tmp_cap)); CUDA_OK(cudaMalloc((void**)&x.d_rows_tmp, 64 * x.tmp_cap));
  return 0;
}

int orgpu_pack_rows(orgpu_engine* e, int n, const int* slots, double* buf)
{
  NEED(e && e->finalized && n >= 0 && (n == 0 || (slots && buf)), -1, "orgpu_pack_rows: bad arguments"); CUDA_OK(cudaSetDevice(e->device));
  if (n == 0) return 0;
  for (int j = 0; j < n; j++) NEED(slots[j] >= 0 && slots[j] < e->lsky, -4, "orgpu_pack_rows: slot %d out of range", slots[j]);
  if (rows_tmp(e, n)) return -100;
  CUDA_OK(cudaMemcpyAsync(e->xc.d_slots_tmp, slots, 4 * (size_t)n, cudaMemcpyHostToDevice, e->st));
  rows_gather8_kernel<<<(n + 255) / 256, 256, 0, e->st>>>(e->d_fsky, e->roww, e->xc.d_slots_tmp, n, e->xc.d_rows_tmp); e->launches++;
  CUDA_OK(cudaMemcpyAsync(buf, e->xc.d_rows_tmp, 64 * (size_t)n, cudaMemcpyDeviceToHost, e->st));
  CUDA_OK(cudaStreamSynchronize(e->st));
  return 0;
}

int orgpu_unpack_rows(orgpu_engine* e, int n, const int* slots, const double* buf)
{
  NEED(e && e->finalized && n >= 0 && (n == 0 || (slots && buf)), -1, "orgpu_unpack_rows: bad arguments"); CUDA_OK(cudaSetDevice(e->device));
  if (n == 0) return 0;
  for (int j = 0; j < n; j++) NEED(slots[j] >= 0 && slots[j] < e->lsky, -4, "orgpu_unpack_rows: slot %d out of range", slots[j]);
  if (rows_tmp(e, n)) return -100;
  CUDA_OK(cudaMemcpyAsync(e->xc.d_slots_tmp, slots, 4 * (size_t)n, cudaMemcpyHostToDevice, e->st));
  CUDA_OK(cudaMemcpyAsync(e->xc.d_rows_tmp, buf, 64 * (size_t)n, cudaMemcpyHostToDevice, e->st));
  rows_scatter8_kernel<<<(n + 255) / 256, 256, 0, e->st>>>(e->d_fsky, e->roww, e->xc.d_slots_tmp, n, e->xc.d_rows_tmp); e->launches++;
  CUDA_OK(cudaStreamSynchronize(e->st));
  return 0;
}

// host-staged /PARITH/OFF exchange (the Engine keeps its MPI): partial sums of n frontier nodes (0-based) out of / added into
// A, AR, STIFN, STIFR as orgpu_assemble left them; buf is (8, n)
int orgpu_pack_nodes(orgpu_engine* e, int n, const int* nodes, double* buf)
{
  NEED(e && e->finalized && n >= 0 && (n == 0 || (nodes && buf)), -1, "orgpu_pack_nodes: bad arguments"); CUDA_OK(cudaSetDevice(e->device));
  if (n == 0) return 0;
  for (int j = 0; j < n; j++) NEED(nodes[j] >= 0 && nodes[j] < e->numnod, -4, "orgpu_pack_nodes: node %d out of range", nodes[j]);
  if (rows_tmp(e, n)) return -100;
  CUDA_OK(cudaMemcpyAsync(e->xc.d_slots_tmp, nodes, 4 * (size_t)n, cudaMemcpyHostToDevice, e->st));
  nodes_pack_kernel<<<(n + 255) / 256, 256, 0, e->st>>>(e->nd, e->xc.d_slots_tmp, n, e->xc.d_rows_tmp); e->launches++;
  CUDA_OK(cudaMemcpyAsync(buf, e->xc.d_rows_tmp, 64 * (size_t)n, cudaMemcpyDeviceToHost, e->st));
  CUDA_OK(cudaStreamSynchronize(e->st));
  return 0;
}
int orgpu_add_nodes(orgpu_engine* e, int n, const int* nodes, const double* buf)
{
  NEED(e && e->finalized && n >= 0 && (n == 0 || (nodes && buf)), -1, "orgpu_add_nodes: bad arguments"); CUDA_OK(cudaSetDevice(e->device));
  if (n == 0) return 0;
  for (int j = 0; j < n; j++) NEED(nodes[j] >= 0 && nodes[j] < e->numnod, -4, "orgpu_add_nodes: node %d out of range", nodes[j]);
  if (rows_tmp(e, n)) return -100;
  CUDA_OK(cudaMemcpyAsync(e->xc.d_slots_tmp, nodes, 4 * (size_t)n, cudaMemcpyHostToDevice, e->st));
  CUDA_OK(cudaMemcpyAsync(e->xc.d_rows_tmp, buf, 64 * (size_t)n, cudaMemcpyHostToDevice, e->st));
  nodes_add_kernel<<<(n + 255) / 256, 256, 0, e->st>>>(e->nd, e->xc.d_slots_tmp, n, e->xc.d_rows_tmp); e->launches++;
  CUDA_OK(cudaStreamSynchronize(e->st));
  return 0;
}

// device-resident /PARITH/OFF loop: after orgpu_comm_init, instead of orgpu_set_exchange.  Neighbour k shares the nodes
// nodes[ptr[k] .. ptr[k+1]) with this rank (0-based local nodes, the same order on both sides: ascending global id).
// orgpu_run_cycles then runs, per cycle: forces -> local dt arg-min -> ASSPAR4 of the local rows -> pack the frontier sums ->
// one NCCL group (sums to / from every neighbour + the ranks' dt candidates) -> add, neighbours in rank order -> nodal update.
int orgpu_set_exchange_nodes(orgpu_engine* e, int nneigh, const int* ranks, const int* ptr, const int* nodes)
{
  NEED(e && e->finalized && nneigh >= 0 && e->xc.comm, -1, "orgpu_set_exchange_nodes: engine not finalized / orgpu_comm_init missing"); CUDA_OK(cudaSetDevice(e->device));
  NEED(!e->ctl.nodadt, -5, "/PARITH/OFF with /DT/NODA is outside the built path");
  Exchange& x = e->xc;
  for (int k = 0; k < nneigh; k++) { NEED(ranks[k] >= 0 && ranks[k] < x.nranks && ranks[k] != x.rank, -4, "orgpu_set_exchange_nodes: neighbour rank %d invalid", ranks[k]);
                                     NEED(k == 0 || ranks[k] > ranks[k - 1], -4, "orgpu_set_exchange_nodes: neighbours must come in ascending rank order (the order of the additions)"); }
  const int tot = nneigh ? ptr[nneigh] : 0;
  for (int j = 0; j < tot; j++) NEED(nodes[j] >= 0 && nodes[j] < e->numnod, -4, "orgpu_set_exchange_nodes: node out of range");
  x.nb_rank.assign(ranks, ranks + nneigh); x.xn_ptr.assign(ptr, ptr + nneigh + 1);
  if (dev_alloc(&x.d_xn_nodes, (size_t)tot + 1) || dev_alloc(&x.d_xn_send, 8 * ((size_t)tot + 1)) || dev_alloc(&x.d_xn_recv, 8 * ((size_t)tot + 1))) return -100;
  if (tot) CUDA_OK(cudaMemcpy(x.d_xn_nodes, nodes, 4 * (size_t)tot, cudaMemcpyHostToDevice));
  x.parith_off = true; e->nd.load_first = 1;            // /PARITH/OFF: FORCE adds to A before the element loop
  return 0;
}

// IPARIT of the Engine: where FORCE's load records enter a node's sum (DevNodes::load_first).  Before the first cycle.
int orgpu_set_parith(orgpu_engine* e, int iparit)
{
  NEED(e && (iparit == 0 || iparit == 1), -1, "orgpu_set_parith: bad arguments");
  if (e->gexec) { cudaGraphExecDestroy(e->gexec); e->gexec = nullptr; }     // the captured cycle carries the flag in its kernel arguments
  e->nd.load_first = iparit ? 0 : 1;
  return 0;
}

int orgpu_comm_unique_id(unsigned char id[128])
{
  NcclApi* N = nccl_api(); if (!N) return -7;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId u; NCCL_OK(N->GetUniqueId(&u)); memcpy(id, &u, 128);
  return 0;
}

int orgpu_comm_init(orgpu_engine* e, int nranks, int rank, const unsigned char id[128])
{
  NEED(e && nranks >= 1 && rank >= 0 && rank < nranks && id, -1, "orgpu_comm_init: bad arguments"); CUDA_OK(cudaSetDevice(e->device));
  NcclApi* N = nccl_api(); if (!N) return -7;
  ncclUniqueId u; memcpy(&u, id, 128);
  NCCL_OK(N->CommInitRank(&e->xc.comm, nranks, u, rank));
  e->xc.nranks = nranks; e->xc.rank = rank;
  if (dev_alloc(&e->xc.d_cand_send, 4) || dev_alloc(&e->xc.d_cand_recv, 4 * (size_t)nranks)) return -100;
  return 0;
}

int orgpu_set_exchange(orgpu_engine* e, int nneigh, const int* ranks, const int* send_ptr, const int* send_slots,
                       const int* recv_ptr, const int* recv_slots)
{
  NEED(e && e->finalized && nneigh >= 0, -1, "orgpu_set_exchange: engine not finalized / bad arguments"); CUDA_OK(cudaSetDevice(e->device));
  Exchange& x = e->xc;
  NEED(x.nranks <= 32, -7, "orgpu_set_exchange: more than 32 ranks on one node");
  x.nb_rank.assign(ranks, ranks + nneigh); x.send_ptr.assign(send_ptr, send_ptr + nneigh + 1); x.recv_ptr.assign(recv_ptr, recv_ptr + nneigh + 1);
  x.nsend = nneigh ? send_ptr[nneigh] : 0; x.nrecv = nneigh ? recv_ptr[nneigh] : 0;
  for (int j = 0; j < x.nsend; j++) NEED(send_slots[j] >= 0 && send_slots[j] < e->lsky, -4, "orgpu_set_exchange: send slot out of range");
  for (int j = 0; j < x.nrecv; j++) NEED(recv_slots[j] >= 0 && recv_slots[j] < e->lsky, -4, "orgpu_set_exchange: recv slot out of range");
  for (int k = 0; k < nneigh; k++) NEED(ranks[k] >= 0 && ranks[k] < x.nranks && ranks[k] != x.rank, -4, "orgpu_set_exchange: neighbour rank %d invalid", ranks[k]);
  if (dev_alloc(&x.d_send_slots, (size_t)x.nsend + 1) || dev_alloc(&x.d_recv_slots, (size_t)x.nrecv + 1) ||
      dev_alloc(&x.d_sendbuf, (size_t)e->roww * (x.nsend + 1)) || dev_alloc(&x.d_recvbuf, (size_t)e->roww * (x.nrecv + 1))) return -100;
  if (x.nsend) CUDA_OK(cudaMemcpy(x.d_send_slots, send_slots, 4 * (size_t)x.nsend, cudaMemcpyHostToDevice));
  if (x.nrecv) CUDA_OK(cudaMemcpy(x.d_recv_slots, recv_slots, 4 * (size_t)x.nrecv, cudaMemcpyHostToDevice));
  return 0;
}

int orgpu_p2p_export(orgpu_engine* e, unsigned char handle[64])
{
  NEED(e && e->finalized && handle, -1, "orgpu_p2p_export: engine not finalized / bad arguments"); CUDA_OK(cudaSetDevice(e->device));
  Exchange& x = e->xc;
  NEED(x.nranks > 1 && x.d_send_slots, -7, "orgpu_p2p_export: call orgpu_comm_init and orgpu_set_exchange first");   // a rank without neighbours still needs its window: dt candidates and flags
  NEED(x.nranks <= 32, -7, "orgpu_p2p_export: more than 32 ranks on one node");
  NEED(!x.win, -7, "orgpu_p2p_export: window already exported");
  x.win_bytes = win_rows_off(x.nranks) + (size_t)2 * (x.nrecv + 1) * e->roww * 8;
  CUDA_OK(cudaMalloc((void**)&x.win, x.win_bytes));
  CUDA_OK(cudaMemset(x.win, 0, x.win_bytes));
  std::vector<int> hdr(256, -1); hdr[0] = x.nranks; hdr[1] = x.nrecv;
  for (size_t k = 0; k < x.nb_rank.size(); k++) hdr[2 + x.nb_rank[k]] = x.recv_ptr[k];
  CUDA_OK(cudaMemcpy(x.win, hdr.data(), 1024, cudaMemcpyHostToDevice));
  CUDA_OK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h; CUDA_OK(cudaIpcGetMemHandle(&h, x.win));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle, &h, 64);
  return 0;
}

int orgpu_p2p_connect(orgpu_engine* e, const unsigned char* handles /*[nranks][64], rank order*/)
{
  NEED(e && e->finalized && handles, -1, "orgpu_p2p_connect: bad arguments"); CUDA_OK(cudaSetDevice(e->device));
  Exchange& x = e->xc;
  NEED(x.win, -7, "orgpu_p2p_connect: call orgpu_p2p_export first");
  const int R = x.nranks, nn = (int)x.nb_rank.size();
  x.peer.assign(R, nullptr); x.peer[x.rank] = x.win;
  for (int q = 0; q < R; q++) {
    if (q == x.rank) continue;
    cudaIpcMemHandle_t h; memcpy(&h, handles + (size_t)64 * q, 64);
    void* p = nullptr; CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    x.peer[q] = (unsigned char*)p;
  }
  // where do my rows land in each neighbour's window, and how long is its row area (for the parity offset)
  std::vector<double*> nb_rows(2 * (nn > 0 ? nn : 1), nullptr);
  for (int k = 0; k < nn; k++) {
    const int q = x.nb_rank[k]; int hdr[64];
    CUDA_OK(cudaMemcpy(hdr, x.peer[q], sizeof hdr, cudaMemcpyDefault));
    NEED(hdr[0] == R && 2 + x.rank < 64 && hdr[2 + x.rank] >= 0, -7, "orgpu_p2p_connect: rank %d does not list rank %d as a neighbour", q, x.rank);
    const int off = hdr[2 + x.rank], nrecv_q = hdr[1];
    double* rows = reinterpret_cast<double*>(x.peer[q] + win_rows_off(R));
    nb_rows[2 * k] = rows + (size_t)off * e->roww;
    nb_rows[2 * k + 1] = rows + ((size_t)nrecv_q + off) * e->roww;
  }
  std::vector<int> send_nb(x.nsend > 0 ? x.nsend : 1, 0), sendptr(nn > 0 ? nn : 1, 0);
  for (int k = 0; k < nn; k++) { sendptr[k] = x.send_ptr[k]; for (int j = x.send_ptr[k]; j < x.send_ptr[k + 1]; j++) send_nb[j] = k; }
  std::vector<double*> pcand(R); std::vector<unsigned long long*> pflag(R);
  for (int q = 0; q < R; q++) { pcand[q] = reinterpret_cast<double*>(x.peer[q] + ORGPU_WIN_CAND);
                                pflag[q] = reinterpret_cast<unsigned long long*>(x.peer[q] + ORGPU_WIN_FLAGS) + x.rank; }
  if (dev_alloc(&x.d_send_nb, send_nb.size()) || dev_alloc(&x.d_nb_sendptr, sendptr.size()) || dev_alloc(&x.d_nb_rows, nb_rows.size()) ||
      dev_alloc(&x.d_peer_cand, (size_t)R) || dev_alloc(&x.d_peer_flag, (size_t)R) || dev_alloc(&x.d_xcycle, 1) || dev_alloc(&x.d_done, 1) ||
      dev_alloc(&x.d_err, 1)) return -100;
  CUDA_OK(cudaMemcpy(x.d_send_nb, send_nb.data(), 4 * send_nb.size(), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(x.d_nb_sendptr, sendptr.data(), 4 * sendptr.size(), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(x.d_nb_rows, nb_rows.data(), 8 * nb_rows.size(), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(x.d_peer_cand, pcand.data(), 8 * (size_t)R, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(x.d_peer_flag, pflag.data(), 8 * (size_t)R, cudaMemcpyHostToDevice));
  if (dev_alloc(&x.d_peer_win, (size_t)R)) return -100;
  CUDA_OK(cudaMemcpy(x.d_peer_win, x.peer.data(), 8 * (size_t)R, cudaMemcpyHostToDevice));
  CUDA_OK(cudaDeviceSynchronize());
  if (e->gexec) { cudaGraphExecDestroy(e->gexec); e->gexec = nullptr; }
  e->tabs_dirty = true; e->desc_gen++;
  x.p2p = true;
  // inline sends: every send slot with exactly one destination leaves from the force kernel that computes it (XSend, common.cuh);
  // a decomposition where some slot has several (a node shared by 3+ domains) keeps the push kernel.  ORGPU_NO_OVERLAP=1: push kernel.
  e->split = false; e->nd.xs = XSend{nullptr, nullptr, nullptr};
  if (!getenv("ORGPU_NO_OVERLAP") && !e->ctl.nodadt) {
    std::vector<int> ss(x.nsend > 0 ? x.nsend : 1);
    if (x.nsend) CUDA_OK(cudaMemcpy(ss.data(), x.d_send_slots, 4 * (size_t)x.nsend, cudaMemcpyDeviceToHost));
    std::vector<int2> ref(e->lsky > 0 ? e->lsky : 1, make_int2(-1, -1));
    bool single = true;
    for (int j = 0; j < x.nsend && single; j++) {
      const int k = send_nb[j];
      if (ref[ss[j]].x >= 0) single = false; else ref[ss[j]] = make_int2(k, j - sendptr[k]);
    }
    if (single) {
      // frontier tiles of every super-group
      auto ftiles = [&](std::vector<void*>& owned, const std::vector<int>& iad, int nn, int first, int ne, int ne_pad, const unsigned char** out) -> int {
        std::vector<unsigned char> f(ne_pad / ORGPU_TILE, 0);
        for (int i = 0; i < ne; i++) for (int k = 0; k < nn; k++) if (ref[iad[(size_t)nn * (first + i) + k] - 1].x >= 0) { f[i / ORGPU_TILE] = 1; break; }
        unsigned char* d; if (upload_vec(owned, &d, f)) return -100; *out = d; return 0; };
      for (auto& S : e->csg) if (ftiles(S.owned, S.sh3n ? e->iadtg : e->iadc, S.sh3n ? 3 : 4, S.first_elem, S.d.ne, S.d.ne_pad, &S.d.xs_ftile)) return -100;
      for (auto& S : e->bsg) if (ftiles(S.owned, e->iads, 8, S.first_elem, S.d.ne, S.d.ne_pad, &S.d.xs_ftile)) return -100;
      // worth it only when few tiles are frontier tiles (slabs of a block numbered layer by layer: 4 % on C5, -4 us per cycle);
      // strips across an x-fastest numbering put two frontier elements in every mesh row, i.e. in a quarter of the tiles (C2),
      // and the read-back loops then cost more than the push kernel (+18 us): those keep the push kernel
      size_t nt = 0, nf = 0;
      { std::vector<unsigned char> h;
        auto cnt = [&](const unsigned char* d, int n) { h.resize(n); cudaMemcpy(h.data(), d, n, cudaMemcpyDeviceToHost); nt += n; for (int i = 0; i < n; i++) nf += h[i]; };
        for (auto& S : e->csg) cnt(S.d.xs_ftile, S.d.ne_pad / ORGPU_TILE);
        for (auto& S : e->bsg) cnt(S.d.xs_ftile, S.d.ne_pad / ORGPU_TILE); }
      const char* fr = getenv("ORGPU_XSEND_MAX_FRAC"); const double maxfrac = fr ? atof(fr) : 0.10;
      if ((double)nf <= maxfrac * (double)nt) {
        if (dev_alloc(&x.d_xsend, ref.size())) return -100;
        CUDA_OK(cudaMemcpy(x.d_xsend, ref.data(), sizeof(int2) * ref.size(), cudaMemcpyHostToDevice));
        e->nd.xs = XSend{x.d_xsend, x.d_nb_rows, x.d_xcycle};
        e->split = true;
      } else {
        for (auto& S : e->csg) S.d.xs_ftile = nullptr;
        for (auto& S : e->bsg) S.d.xs_ftile = nullptr;
      }
    }
  }
  return 0;
}

int orgpu_exchange(orgpu_engine* e)
{
  NEED(e && e->finalized, -1, "orgpu_exchange: engine not finalized"); CUDA_OK(cudaSetDevice(e->device));
  return exchange_on_stream(e, false);
}

// a, b, off: word rows of a tile-major slab (nw words per tile) when nw > 0, plain arrays when nw == 0
static int energy_sum(orgpu_engine* e, const double* a, const double* b, const double* off, int nw, int n, int mode, double* out)
{
  if (n <= 0) return 0;
  const int nb = (n + 255) / 256;
  double* d_part = nullptr; CUDA_OK(cudaMalloc((void**)&d_part, 8 * (size_t)nb));
  energy_partial_kernel<<<nb, 256, 0, e->st>>>(a, b, off, nw, n, mode, d_part); e->launches++;
  std::vector<double> h(nb);
  CUDA_OK(cudaMemcpyAsync(h.data(), d_part, 8 * (size_t)nb, cudaMemcpyDeviceToHost, e->st));
  CUDA_OK(cudaStreamSynchronize(e->st));
  cudaFree(d_part);
  double s = 0.0; for (int k = 0; k < nb; k++) s += h[k];
  *out += s;
  return 0;
}

int orgpu_get_energies(orgpu_engine* e, double out[4])
{
  NEED(e && e->finalized && out, -1, "orgpu_get_energies: engine not finalized"); CUDA_OK(cudaSetDevice(e->device));
  out[0] = out[1] = out[2] = out[3] = 0.0;        // internal solids, internal shells, kinetic translation, kinetic rotation
  for (auto& S : e->bsg) if (energy_sum(e, S.d.slab + BW_EINT * ORGPU_TILE, S.d.slab + S.d.w_vol * ORGPU_TILE, nullptr, S.d.nw, S.d.ne, 0, &out[0])) return -100;
  for (auto& S : e->csg) if (energy_sum(e, S.d.slab + SW_EINT * ORGPU_TILE, S.d.slab + (SW_EINT + 1) * ORGPU_TILE, S.d.slab + SW_OFF * ORGPU_TILE, S.d.nw, S.d.ne, 1, &out[1])) return -100;
  if (energy_sum(e, e->nd.MS, (const double*)e->nd.vel, nullptr, 0, e->numnod, 2, &out[2])) return -100;
  if (e->nd.rot && energy_sum(e, e->nd.IN, (const double*)e->nd.rot, nullptr, 0, e->numnod, 2, &out[3])) return -100;
  return 0;
}

// IPRI: the following cycles also book the balances of the reference's print cycles.  The scratch rows, chunk table and
// result buffers are allocated once, here (nothing is allocated per cycle or per query).
int orgpu_set_print(orgpu_engine* e, int ipri)
{
  NEED(e && e->finalized, -1, "orgpu_set_print: engine not finalized"); CUDA_OK(cudaSetDevice(e->device));
  CUDA_OK(cudaStreamSynchronize(e->st));
  ipri = ipri ? 1 : 0;
  if (ipri && !e->d_bal) {
    NEED(e->have_parts, -4, "orgpu_set_print: parts and shell volumes missing (orgpu_set_parts before orgpu_finalize)");
    std::vector<BalChunk> ch; int ofs = 0;
    auto add = [&](int ne, int ne_pad, int part) {
      for (int s0 = 0; s0 < ne; s0 += ORGPU_BAL_CHUNK) ch.push_back(BalChunk{ofs + s0, (ne - s0 < ORGPU_BAL_CHUNK) ? ne - s0 : ORGPU_BAL_CHUNK, part});
      const int o = ofs; ofs += ne_pad; return o; };
    std::vector<int> so, bo;
    for (auto& S : e->csg) so.push_back(add(S.d.ne, S.d.ne_pad, S.part));
    for (auto& S : e->bsg) bo.push_back(add(S.d.ne, S.d.ne_pad, S.part));
    e->bal_ld = ofs; e->nchunk = (int)ch.size(); e->nbal_ld = e->numnod; e->nnchunk = (e->numnod + ORGPU_BAL_CHUNK - 1) / ORGPU_BAL_CHUNK;
    if (dev_alloc(&e->d_bal, (size_t)6 * e->bal_ld) || dev_alloc(&e->d_nbal, (size_t)8 * e->nbal_ld) || dev_alloc(&e->d_epart, (size_t)6 * e->nchunk) ||
        dev_alloc(&e->d_npartial, (size_t)8 * e->nnchunk) || dev_alloc(&e->d_partsav, (size_t)6 * e->npart) || dev_alloc(&e->d_hist, (size_t)8 * ORGPU_BAL_HIST) ||
        dev_alloc(&e->d_chunks, ch.size()) || dev_alloc(&e->d_bs, 1)) return -100;
    CUDA_OK(cudaMemcpy(e->d_chunks, ch.data(), sizeof(BalChunk) * ch.size(), cudaMemcpyHostToDevice));
    for (size_t k = 0; k < e->csg.size(); k++) { e->csg[k].d.bal = e->d_bal + so[k]; e->csg[k].d.bal_ld = e->bal_ld; }
    for (size_t k = 0; k < e->bsg.size(); k++) { e->bsg[k].d.bal = e->d_bal + bo[k]; e->bsg[k].d.bal_ld = e->bal_ld; }
    e->nd.nbal = e->d_nbal; e->nd.nbal_ld = e->nbal_ld;
    e->tabs_dirty = true; e->desc_gen++;
  }
  if (ipri != e->ipri && e->gexec) { cudaGraphExecDestroy(e->gexec); e->gexec = nullptr; }   // the cycle graph gains / loses the three balance launches
  e->ipri = ipri;
  CUDA_OK(cudaMemcpy(reinterpret_cast<char*>(e->d_cs) + offsetof(CycleState, ipri), &ipri, sizeof(int), cudaMemcpyHostToDevice));
  return 0;
}

// out: ENCIN, ENROT, ENINT, WFEXT, XMOMT, YMOMT, ZMOMT, XMASS of the last print cycle (ecrit.F:322-352); partsav (6, npart) or null
int orgpu_get_balance(orgpu_engine* e, double out[8], double* partsav)
{
  NEED(e && e->finalized && out && e->d_bs, -1, "orgpu_get_balance: no print cycle was run (orgpu_set_print)"); CUDA_OK(cudaSetDevice(e->device));
  CUDA_OK(cudaStreamSynchronize(e->st));
  BalState bs; CUDA_OK(cudaMemcpy(&bs, e->d_bs, sizeof bs, cudaMemcpyDeviceToHost));
  for (int k = 0; k < 8; k++) out[k] = bs.glob[k];
  if (partsav) CUDA_OK(cudaMemcpy(partsav, e->d_partsav, sizeof(double) * 6 * e->npart, cudaMemcpyDeviceToHost));
  return 0;
}

// rows of the last n print cycles (oldest first), n <= ORGPU_BAL_HIST: what the listing shows cycle by cycle
int orgpu_get_balance_history(orgpu_engine* e, int n, double* out /*[n][8]*/)
{
  NEED(e && e->finalized && out && e->d_bs && n >= 0 && n <= ORGPU_BAL_HIST, -1, "orgpu_get_balance_history: bad arguments / no print cycle was run"); CUDA_OK(cudaSetDevice(e->device));
  CUDA_OK(cudaStreamSynchronize(e->st));
  CycleState cs; CUDA_OK(cudaMemcpy(&cs, e->d_cs, sizeof cs, cudaMemcpyDeviceToHost));
  NEED(cs.ncycle >= n, -4, "orgpu_get_balance_history: only %lld cycles were run", cs.ncycle);
  std::vector<double> h((size_t)8 * ORGPU_BAL_HIST);
  CUDA_OK(cudaMemcpy(h.data(), e->d_hist, 8 * h.size(), cudaMemcpyDeviceToHost));
  for (int j = 0; j < n; j++) { const long long row = (cs.ncycle - n + j) % ORGPU_BAL_HIST; memcpy(out + 8 * (size_t)j, &h[8 * (size_t)row], 64); }
  return 0;
}

// Through-thickness rule of the NPT-point /PROP/SHELL (positions Z0, force weights WF, moment weights WM, coqini.F tables by
// default): replaces the row of the device tables shared by every engine of the process until the next orgpu_finalize.  The
// reference's own CUDA kernels integrate with the mid-point rule (shell_strain_material_kernel.cu:696-701); with that rule
// loaded here the bending response is pinned against them (tests/test_ref_gpu_pin.py).
int orgpu_set_quadrature(orgpu_engine* e, int npt, const double* z0, const double* wf, const double* wm)
{
  NEED(e && e->finalized && npt >= 1 && npt <= 10 && z0 && wf && wm, -1, "orgpu_set_quadrature: bad arguments / engine not finalized"); CUDA_OK(cudaSetDevice(e->device));
  CUDA_OK(cudaStreamSynchronize(e->st));
  const size_t off = sizeof(double) * 11 * (npt - 1);
  CUDA_OK(cudaMemcpyToSymbol(c_Z0, z0, sizeof(double) * npt, off)); CUDA_OK(cudaMemcpyToSymbol(c_WF, wf, sizeof(double) * npt, off));
  CUDA_OK(cudaMemcpyToSymbol(c_WM, wm, sizeof(double) * npt, off));
  return 0;
}

long long orgpu_launch_count(orgpu_engine* e) { return e ? e->launches : 0; }
double orgpu_last_run_ms(orgpu_engine* e) { return e ? e->last_run_ms : 0; }
int orgpu_set_profile(orgpu_engine* e, int profile) { NEED(e, -1, "null handle"); e->profile = profile; return 0; }
int orgpu_get_profile(orgpu_engine* e, int cls, double* ms, long long* launches)
{
  NEED(e && cls >= 0 && cls < 3, -1, "bad class"); *ms = e->prof_ms[cls]; *launches = e->prof_n[cls]; return 0;
}

} // extern "C"

#include "shell_gpu_compat.cuh"   // the reference's own shell_gpu_* ABI on top of the entries above
