/* orgpu.h -- C ABI of liborgpu.so: the B200-native explicit element cycle for the OpenRadioss
 * Engine (8-node bricks SFORC3, 4-node shells CFORC3/CZFORC3, LAW2/LAW36, dt argmin, /PARITH/ON
 * assembly ASSPAR4, ACCELE/VELOCITY/DEPLA).  Plain pointers and sizes only; every entry returns
 * 0 on success or a negative status, and orgpu_last_error() gives the text (the reference's own
 * GPU ABI aborts with exit(EXIT_FAILURE), shell_gpu_driver.cu:47-55; a Fortran shim may do the
 * same on a non-zero status).
 *
 * What each entry replaces in the reference (paths under /root/reference):
 *   engine/source/elements/shell/coque/shell_gpu_driver.h:44-206  -- the shipped "-gpu" C ABI
 *   engine/source/elements/shell/coque/shell_gpu_mod.F90:284-715   -- its ISO_C_BINDING side
 *   engine/source/elements/shell/coque/shell_internal_forces.F90   -- FORINTC_PREPARE_GPU (:370),
 *        gpu_shell_launch_async (:62), gpu_shell_sync_scatter (:177), called from
 *        engine/source/engine/resol.F:2657, 3670, 4295
 * This ABI is a superset: bricks, QEPH and LAW36 are eligible, nodal arrays stay resident on the
 * device, assembly is the deterministic FSKY/ADSKY gather of ASSPAR4 (asspar4.F:164-181) and the
 * nodal update (accele.F, velocity.F, displacement.F) also runs on the device.
 *
 * Array conventions follow the Fortran caller: X(3,NUMNOD) column-major => X[3*n+c]; IXS(11,*),
 * IXC(7,*), IADS(8,*), IADC(4,*), ADSKY(NUMNOD+1) are passed as the Engine holds them (1-based
 * node numbers and 1-based FSKY slot addresses).  Host arrays stay owned by the caller; all
 * device memory is owned by the library.  Calls on one handle must come from one thread at a
 * time (the reference calls its GPU ABI from the master thread only, resol.F:3665-3671).
 */
#ifndef ORGPU_H
#define ORGPU_H
#include "orgpu_model.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orgpu_engine orgpu_engine;

/* -- life cycle: shell_gpu_global_create / shell_gpu_data_create (+ cudaSetDevice from the local rank) */
int  orgpu_create(orgpu_engine** out, int device, int numnod, const orgpu_control* ctl);
int  orgpu_destroy(orgpu_engine* e);
const char* orgpu_last_error(void);

/* -- model upload (once): shell_gpu_global_upload_nodes / shell_gpu_upload_constant /
 *    shell_gpu_upload_ip_state / shell_gpu_set_mat_params, generalised.  NULL = keep current. */
int  orgpu_upload_nodes(orgpu_engine* e, const double* X, const double* V, const double* VR,
                        const double* D, const double* MS, const double* IN);
int  orgpu_set_loads(orgpu_engine* e, const double* FEXT, const double* MEXT); /* constant nodal loads (3,N) */
int  orgpu_set_bcs(orgpu_engine* e, const int* icodt, const int* icodr);       /* BCS10 codes 4:x 2:y 1:z  */
/* time variation of the nodal loads above: A += FEXT * FINTER(ifunc, TT*fcx) as the concentrated loads of
 * FORCE (engine/source/loads/general/force.F90:195-196, 235, 301-312; resol.F:2929) sharing one function;
 * ifunc = 0-based curve of orgpu_set_functions, -1 = constant loads.  Before orgpu_finalize. */
int  orgpu_set_load_function(orgpu_engine* e, int ifunc, double fcx);
/* user node ids ITAB(NUMNOD): NELTST of a nodal time step (control.nodadt = 1: DTNODA, engine/source/time_step/
 * dtnoda.F:221-260, 324-338, 445-462; resol.F:6066).  Before orgpu_finalize. */
int  orgpu_set_itab(orgpu_engine* e, const int* itab);
/* imposed velocities (FIXVEL, engine/source/constraints/general/impvel/fixvel.F:141-147, 334-378; resol.F:7610):
 * record k = IBFV(1,k) node (1-based), IBFV(2,k) direction 1..3 in the global frame, IBFV(3,k) curve (0-based),
 * VEL(1,k) FAC, VEL(2,k) start time, VEL(3,k) stop time, VEL(5,k) FACX.  Before orgpu_finalize. */
int  orgpu_set_fixvel(orgpu_engine* e, int nfxvel, const int* ibfv /*(3,n)*/, const double* vel /*(4,n)*/);
/* gravity loads (GRAVIT, engine/source/loads/general/grav/gravit.F:84-160; resol.F:7123, after ACCELE and before BCS10):
 * A(N2,node) += FCY * FINTER(IFUNC, TT*FCX).  Load l = IGRV(1,l) node count, IGRV(2,l) direction 1..3 (global frame: ISK <= 1),
 * IGRV(3,l) curve (0-based, -1 = constant); AGRV(1,l) FCY, AGRV(2,l) FCX; ib = the node lists IB one after the other
 * (1-based, sign ignored: it only selects the nodes counted in the external work).  No sensor.  Before orgpu_finalize. */
int  orgpu_set_gravity(orgpu_engine* e, int ngrav, const int* igrv /*(3,n)*/, const double* agrv /*(2,n)*/, const int* ib, int lib);
int  orgpu_set_solids(orgpu_engine* e, int numels, const int* ixs, const int* iads);
int  orgpu_set_shells(orgpu_engine* e, int numelc, const int* ixc, const int* iadc);
/* 3-node shells (ITY=7, C3FORC3: engine/source/elements/sh3n/coque3n/c3forc3.F:35, called forintc.F:628): IXTG(6,NUMELTG)
 * and their FSKY slots IADTG(3,NUMELTG) (parith_on_mod.F90), 1-based as the Engine holds them */
int  orgpu_set_sh3n(orgpu_engine* e, int numeltg, const int* ixtg, const int* iadtg);
int  orgpu_set_pon(orgpu_engine* e, const int* adsky, int lsky);              /* parith_on_mod.F90:39-74 */
int  orgpu_set_functions(orgpu_engine* e, int nfunc, const int* npf, const double* tf);
/* one element group (<= NVSIZ elements, one material / property): elements [nft, nft+nel) */
int  orgpu_add_solid_group(orgpu_engine* e, int nel, int nft, const orgpu_law2* mat,
                           const orgpu_prop_solid* prop, const double* vol0);
/* same for any built solid law: law = 2 (orgpu_law2, M2LAW) or 36 (orgpu_law36: MMAIN -> MULAW -> SIGEPS36,
 * engine/source/materials/mat_share/mulaw.F90:1133-1166, mat/mat036/sigeps36.F:35); needs orgpu_set_functions */
int  orgpu_add_solid_group_law(orgpu_engine* e, int nel, int nft, int law, const void* mat,
                               const orgpu_prop_solid* prop, const double* vol0);
int  orgpu_add_shell_group(orgpu_engine* e, int nel, int nft, int law, const void* mat,
                           const orgpu_prop_shell* prop);
/* one 3-node shell group: elements [nft, nft+nel) of IXTG; prop->ihbe carries Ish3n = IPARG(23) (1 or 2) */
int  orgpu_add_sh3n_group(orgpu_engine* e, int nel, int nft, int law, const void* mat,
                          const orgpu_prop_shell* prop);
/* IPARIT of the Engine (default 1 = /PARITH/ON): with /PARITH/ON FORCE leaves each load record in an FSKY row of its own, which the
 * Starter places BEHIND the node's element rows (force.F90:714-1034, starter/source/spmd/domdec2.F:2363-2388), so ASSPAR4 adds a node's
 * load after its element forces; with /PARITH/OFF (iparit = 0) FORCE adds to A before the element loop (force.F90:182-312) and the
 * nodal sum starts from the load.  orgpu_set_exchange_nodes (the device-resident /PARITH/OFF exchange) implies 0. */
int  orgpu_set_parith(orgpu_engine* e, int iparit);
/* Concentrated loads record by record, as the Engine holds them (FORCE, engine/source/loads/general/force.F90:188-312; IB / FAC of
 * /CLOAD): record l loads node ib[3l] (1-based) in direction ib[3l+1] (1..3 forces, 4..6 moments, global frame) with
 * FCY * f(TT * FCX), f = time function ib[3l+2] (0-based index into orgpu_set_functions, -1: constant), FCY = fac[2l], FCX = fac[2l+1].
 * Any number of functions; records of one node are added in record order.  Instead of orgpu_set_loads / orgpu_set_load_function
 * (one curve for all loads).  Sensors, skew frames and displacement- / velocity-dependent abscissae are outside the built path. */
int  orgpu_set_cloads(orgpu_engine* e, int nload, const int* ib /*(3,nload)*/, const double* fac /*(2,nload)*/);
/* /FAIL/JOHNSON for one shell group (sh3n = 0: the index orgpu_add_shell_group returned, 1: orgpu_add_sh3n_group): damage
 * DFMAX += DPLA / eps_f per integration point after the law (mulawc.F90:2118-2127 -> fail_johnson_c.F), a failed point keeps
 * contributing this cycle's stress but restarts from zero stress every cycle (:2608-2637), the element is deleted when the
 * broken share of its thickness / of its points reaches P_thickfail (fail_setoff_c.F:123-186).  Two more words per point
 * (damage, point flag): shell state fields 14 dfmax(npt), 15 foff(npt).  Before orgpu_finalize. */
int  orgpu_set_shell_group_fail(orgpu_engine* e, int sh3n, int group, const orgpu_fail* f);
/* ... and for one LAW2 solid group (the index orgpu_add_solid_group returned): FAIL_JOHNSON behind MMAIN's own failure section
 * (mmain.F90:2250-2416 -> fail/johnson_cook/fail_johnson.F:95-141, Ifail_so = 1): elements whose damage reaches 1 get OFF = 4/5
 * and relax by 0.8 per cycle until they are deleted.  One more state word (solid state field 13 dfmax); pthk / pthickg unused. */
int  orgpu_set_solid_group_fail(orgpu_engine* e, int group, const orgpu_fail* f);
/* FORINTC_PREPARE_GPU analogue: fuse consecutive compatible groups into super-groups, re-lay
 * ELBUF out as device SoA, upload tables.  Must be called once before stepping. */
int  orgpu_finalize(orgpu_engine* e);

/* -- one cycle in the three phases RESOL sees (host supplies the time steps) */
int  orgpu_forces_phase(orgpu_engine* e, double dt1);  /* FORINTC+FORINT: corner rows into FSKY, DT2T argmin */
int  orgpu_assemble(orgpu_engine* e);                  /* ASSPAR4: A, AR, STIFN, STIFR                       */
int  orgpu_advance(orgpu_engine* e, double dt12, double dt2); /* ACCELE, BCS, VELOCITY, DEPLA               */
/* -- device-resident time loop: ncycles passes with the RESOL dt bookkeeping on the device
 *    (resol.F:2721, 6124-6128, 6352, 6494-6497); no host synchronisation inside. */
int  orgpu_run_cycles(orgpu_engine* e, int ncycles);
int  orgpu_synchronize(orgpu_engine* e);

/* -- read-back (output / restart cycles): shell_gpu_global_download_forces / shell_gpu_download_state */
int  orgpu_get_time(orgpu_engine* e, double out[5] /*tt,dt1,dt2,dt12,dt2t*/, int iout[3] /*neltst,ityptst,ncycle*/);
int  orgpu_download_nodes(orgpu_engine* e, double* X, double* V, double* VR, double* D,
                          double* A, double* AR, double* STIFN, double* STIFR);
int  orgpu_download_fsky(orgpu_engine* e, double* fsky /*(8,LSKY)*/);
/* fields: 0 sig(6) 1 eint 2 rho 3 qvis 4 pla 5 epsd 6 vol 7 off 8 temp 9 smstr(21) 10 stra(6) 11 wpla (LAW36) 12 sigb(6) (LAW2 with FISOKIN > 0: LBUF%SIGB) 13 dfmax (/FAIL/JOHNSON); out[k*numels+e] */
int  orgpu_download_solid_state(orgpu_engine* e, int field, double* out);
/* shell fields: 0 for(5) 1 mom(3) 2 eint(2) 3 thk 4 off 5 stra(8) 6 epsd 7 hourg(12 | 5) 8 smstr(6) 9 sig(5*npt) 10 pla(npt) 11 epsd_ip(npt)
 * 12 temp(npt) 13 sigb(3*npt) 14 dfmax(npt) 15 foff(npt) 16 plap(npt) (LAW36 VP = 1: UVAR(2), the filtered plastic strain rate,
 * sigeps36c.F:705, 976-982); out[k*numelc+e] */
int  orgpu_download_shell_state(orgpu_engine* e, int field, double* out);
int  orgpu_download_sh3n_state(orgpu_engine* e, int field, double* out);   /* shell fields; smstr has 3 words, no hourg */
/* -- state hand-over in the other direction (restart / a run that starts from an initial state: the role of
 *    shell_gpu_upload_ip_state, shell_gpu_driver.h, and of RDRESB reading ELBUF from the restart file): same
 *    fields and layout as the downloads; orgpu_set_time restores TT, DT2, DT2OLD, NCYCLE (resol.F restart values). */
int  orgpu_upload_solid_state(orgpu_engine* e, int field, const double* in);
int  orgpu_upload_shell_state(orgpu_engine* e, int field, const double* in);
int  orgpu_upload_sh3n_state(orgpu_engine* e, int field, const double* in);
int  orgpu_set_time(orgpu_engine* e, double tt, double dt2, double dt2old, long long ncycle);

/* -- print-cycle energy balances (SBILAN sbilan.F:138-157, CBILAN cbilan.F:183-275 -> PARTSAV(1,.)
 *    -> ECRIT output/ecrit.F:259-373): out = internal energy of solids, of shells, nodal kinetic
 *    energy of translations, of rotations.  Deterministic (fixed-order) reduction. */
int  orgpu_get_energies(orgpu_engine* e, double out[4]);
/* -- the same balances as the Engine books them on a print cycle (IPRI = 1), per part and for the global line of the
 *    listing.  orgpu_set_parts (before orgpu_finalize): part (0-based) of every element = IPARTC / IPARTS / IPARTTG, and
 *    GBUF%VOL of the shells (initial area x thickness, Starter output), the mass CBILAN uses.  orgpu_set_print(1): every
 *    following cycle also runs
 *      CBILAN  (engine/source/elements/shell/coque/cbilan.F:183-275, called czforc3.F:639 / cforc3.F:648),
 *      C3BILAN (sh3n/coque3n/c3bilan.F:150-167, 282-300; c3forc3.F:616), SBILAN (solid/solide/sbilan.F:110-157; sforc3.F:1436)
 *        -> PARTSAV(1:6, part): internal energy, kinetic energy from the nodal velocities the force routine sees, momenta, mass;
 *      ECRIT   (engine/source/output/ecrit.F:178-240, 322-352): ENCIN / ENROT from V(n-1/2) + DT1/2 A with A after every
 *        kinematic condition, ENINT = sum of PARTSAV(1,:), momenta, mass;
 *      the work of the imposed velocities as FIXVEL books it (constraints/general/impvel/fixvel.F:342-344, 391-394, 834-837).
 *    Scratch rows + a fixed-order two-level reduction (no atomics, nothing allocated per cycle); one domain only.
 *    orgpu_get_balance: out = ENCIN, ENROT, ENINT, WFEXT, XMOMT, YMOMT, ZMOMT, XMASS of the last cycle, partsav (6,npart)
 *    or NULL; orgpu_get_balance_history: the rows of the last n cycles, oldest first (n <= 8192). */
int  orgpu_set_parts(orgpu_engine* e, int npart, const int* ipartc, const int* iparts, const int* iparttg,
                     const double* gvolc, const double* gvoltg);
int  orgpu_set_print(orgpu_engine* e, int ipri);
int  orgpu_get_balance(orgpu_engine* e, double out[8], double* partsav);
int  orgpu_get_balance_history(orgpu_engine* e, int n, double* out /*[n][8]*/);
/* -- /PARITH/OFF: SPMD_EXCH_A (engine/source/mpi/forces/spmd_exch_a.F:34; pack :153-177, add :517-528).  Every domain assembles
 *    the corner rows of its own elements only (the reserved remote slots of its skyline stay zero); the partial sums of the
 *    frontier nodes -- A(1:3), AR(1:3), STIFN, STIFR: 8 doubles per node -- are exchanged and added neighbour by neighbour in
 *    rank order.  Unlike /PARITH/ON the sum order then depends on the decomposition (results agree to rounding, not bitwise).
 *    Host-staged (the Engine keeps its MPI): orgpu_forces_phase, orgpu_assemble, orgpu_pack_nodes per neighbour, MPI,
 *    orgpu_add_nodes per neighbour in rank order, orgpu_advance.  Device-resident: orgpu_comm_init + orgpu_set_exchange_nodes
 *    (instead of orgpu_set_exchange), then orgpu_run_cycles exchanges over NCCL.  External nodal loads of a frontier node must
 *    be given to one domain only (the Starter assigns each load record to one domain). */
int  orgpu_pack_nodes(orgpu_engine* e, int n, const int* nodes /*0-based*/, double* buf /*(8,n)*/);
int  orgpu_add_nodes(orgpu_engine* e, int n, const int* nodes /*0-based*/, const double* buf /*(8,n)*/);
int  orgpu_set_exchange_nodes(orgpu_engine* e, int nneigh, const int* ranks /*ascending*/, const int* ptr /*nneigh+1*/, const int* nodes);
/* -- tie-break keys of the time-step arg-min across domains: index of every local element in the processing order of the
 *    undecomposed model (4-node shells, 3-node shells, solids) and global index of every local node.  SPMD_GLOB_MIN5
 *    (engine/source/mpi/generic/spmd_glob_min5.F:122) keeps the first minimum in reduction order; with these keys N domains
 *    elect, on an exact tie, the element one domain elects (NELTST identical on any GPU count).  NULL = local order.
 *    Before orgpu_finalize. */
int  orgpu_set_global_order(orgpu_engine* e, const int* gshell, const int* gsh3n, const int* gsolid, const int* gnode);
/* -- how long a rank waits for its neighbours inside the peer-memory exchange before the handle is declared dead (default
 *    30 s, or ORGPU_P2P_TIMEOUT_S).  A timeout is sticky: every later kernel of the handle is a no-op, the device state stays at
 *    the last completed cycle, and orgpu_synchronize / get_time / download_* / step_host return -8. */
int  orgpu_set_exchange_timeout(orgpu_engine* e, double seconds);
/* -- through-thickness integration rule of the NPT-point /PROP/SHELL: positions Z0, force weights WF, moment weights WM
 *    (engine/source/elements/shell/coqini.F:46-122 by default; layini.F:246-254, mulawc.F90:769-777).  Replaces the row of the
 *    device tables (shared by the engines of a process) until the next orgpu_finalize.  After orgpu_finalize. */
int  orgpu_set_quadrature(orgpu_engine* e, int npt, const double* z0, const double* wf, const double* wm);

/* -- domain decomposition (one process / MPI rank per GPU).
 *    Host-staged: corner rows of the given 0-based local FSKY slots out of / into the device
 *    skyline as (8,n) rows -- exactly what SPMD_EXCH2_A_PON packs from FSKY(:,ISENDP(j))
 *    (engine/source/mpi/forces/spmd_exch2_a_pon.F:545-557) and unpacks into FSKY(:,IRECVP(j))
 *    (:1190-1201), so the Engine can keep its MPI exchange (resol.F:4801) between
 *    orgpu_forces_phase and orgpu_assemble. */
int  orgpu_pack_rows(orgpu_engine* e, int n, const int* slots, double* rows /*(8,n)*/);
int  orgpu_unpack_rows(orgpu_engine* e, int n, const int* slots, const double* rows /*(8,n)*/);
/*    Device-resident: NCCL over NVLink replaces SPMD_EXCH2_A_PON and SPMD_GLOB_MIN5
 *    (engine/source/mpi/generic/spmd_glob_min5.F:34-128) inside orgpu_run_cycles.  Rank 0 creates the
 *    128-byte id, the host broadcasts it (MPI_BCAST / torch.distributed), every rank calls
 *    orgpu_comm_init, then orgpu_set_exchange with its neighbour lists: neighbour k sends rows of
 *    send_slots[send_ptr[k]..send_ptr[k+1]) and fills recv_slots[recv_ptr[k]..recv_ptr[k+1]); both sides
 *    list a pair's rows in ascending global slot order (IADSDP / IADRCP equivalents). */
int  orgpu_comm_unique_id(unsigned char id[128]);
int  orgpu_comm_init(orgpu_engine* e, int nranks, int rank, const unsigned char id[128]);
int  orgpu_set_exchange(orgpu_engine* e, int nneigh, const int* ranks, const int* send_ptr, const int* send_slots,
                        const int* recv_ptr, const int* recv_slots);
int  orgpu_exchange(orgpu_engine* e);   /* phased mode: pack -> NCCL -> unpack on the library stream */
/*    Peer-memory exchange (ranks of one NVLink / NVSwitch node): after orgpu_set_exchange every rank exports
 *    the CUDA IPC handle of its receive window, the host gathers the nranks handles (MPI_ALLGATHER /
 *    torch.distributed) and every rank connects.  From then on orgpu_run_cycles pushes the corner rows and
 *    the dt candidate straight into the neighbours' HBM with its own kernels (no NCCL call in the cycle,
 *    one CUDA graph per cycle); SPMD_EXCH2_A_PON / SPMD_GLOB_MIN5 semantics are unchanged. */
int  orgpu_p2p_export(orgpu_engine* e, unsigned char handle[64]);
int  orgpu_p2p_connect(orgpu_engine* e, const unsigned char* handles /*[nranks][64] in rank order*/);

/* -- host-owned nodal arrays, one call per step (the usage pattern of the reference's own -gpu path, which re-uploads X, V, VR
 *    every cycle: shell_internal_forces.F90:106): X, V, VR in (NULL: keep the device copy), ncycles on the device, X, V(, VR) out.
 *    Pinned host memory; PCIe-bound (24 bytes per node, array and direction). */
int  orgpu_step_host(orgpu_engine* e, const double* X, const double* V, const double* VR,
                     int ncycles, double* Xout, double* Vout);
int  orgpu_step_host_rot(orgpu_engine* e, const double* X, const double* V, const double* VR,
                         int ncycles, double* Xout, double* Vout, double* VRout);
/* -- the cycle of the reference's -gpu path in ONE call (shell_internal_forces.F90:106-190 + shell_gpu_driver.cu:150-190: upload
 *    X, V, VR, run the force kernels, download the assembled nodal forces; the host integrates): X, V, VR(3,NUMNOD) in, the
 *    internal forces F8(8,NUMNOD) = Fx,Fy,Fz,Mx,My,Mz,STIFN,STIFR per node out (what ASSPAR4 leaves in A, AR, STIFN, STIFR, without
 *    the external loads: those are the caller's), and the element time step DT2T with its element (NELTST, ITYPTST).  Element
 *    state advances by one cycle exactly as in orgpu_forces_phase(dt1).  Uploads, kernels and downloads are pipelined over
 *    chunks of the node range (both directions of the link at once, kernels hidden behind them); host arrays must be pinned
 *    (cudaHostRegister / cudaMallocHost) for that.  Single domain, element time step (no /DT/NODA), not on a print cycle. */
int  orgpu_forces_host(orgpu_engine* e, const double* X, const double* V, const double* VR, double dt1,
                       double* F8, double* dt2t, int* neltst, int* ityptst);

/* -- instrumentation: number of kernels launched by this handle so far; device ms of the last
 *    orgpu_run_cycles measured with CUDA events on the library's stream */
long long orgpu_launch_count(orgpu_engine* e);
double    orgpu_last_run_ms(orgpu_engine* e);
/* per-kernel-class accumulated device time (ms) and launches of the last profiled run:
 *   cls 0 = brick forces, 1 = shell forces, 2 = node gather+update; enable with profile=1 */
int  orgpu_set_profile(orgpu_engine* e, int profile);
int  orgpu_get_profile(orgpu_engine* e, int cls, double* ms, long long* launches);

#ifdef __cplusplus
}
#endif
#endif
