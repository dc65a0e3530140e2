/* shell_gpu_abi.h -- the reference's own C ABI for its GPU shell path, as exported by liborgpu.so.
 *
 * Same names, argument lists and meaning as engine/source/elements/shell/coque/shell_gpu_driver.h:44-206 (the 30
 * `extern "C"` entry points) plus shell_gpu_data_create / shell_gpu_data_destroy (shell_gpu_driver.cu:261-275), which
 * the Fortran module shell_gpu_mod.F90:284-715 binds with bind(c, name=...) and shell_internal_forces.F90 calls
 * (FORINTC_PREPARE_GPU :370, gpu_shell_launch_async :62, gpu_shell_sync_scatter :177; RESOL :2657, :3670, :4295).
 * An Engine built WITH_CUDA therefore links liborgpu.so in place of shell_gpu_driver.cu + the three kernel files with
 * no change on the Fortran side.  Real = double (the Engine's my_real in the r8 build, -DMYREAL8).
 * Implementation: openradioss_b200/csrc/shell_gpu_compat.cuh; visible differences: INTEGRATION.md section 4.
 * The wider ABI of this library (bricks, QEPH, LAW36, device-resident cycles, domains) is include/orgpu.h. */
#ifndef SHELL_GPU_ABI_H
#define SHELL_GPU_ABI_H

#ifdef __cplusplus
extern "C" {
#endif

typedef double Real;
typedef struct ShellGPUGlobal ShellGPUGlobal;   /* shell_gpu_data.h:41-70  */
typedef struct ShellGPUData ShellGPUData;       /* shell_gpu_data.h:72-232 */

/* global node / force handle (shell_gpu_driver.h:44-76) */
ShellGPUGlobal* shell_gpu_global_create(int NUMNOD);
void shell_gpu_global_destroy(ShellGPUGlobal* gh);
void shell_gpu_global_upload_nodes(ShellGPUGlobal* gh, const Real* X, const Real* V, const Real* VR);   /* (3,NUMNOD) each */
void shell_gpu_global_download_forces(ShellGPUGlobal* gh, Real* raw_gpu_to_cpu);   /* [Fx|Fy|Fz|Mx|My|Mz|STIFN|STIFR] x NUMNOD */
void shell_gpu_global_synchronize(ShellGPUGlobal* gh);
void shell_gpu_global_wait_upload(ShellGPUGlobal* gh, ShellGPUData* g);
void shell_gpu_global_wait_su(ShellGPUGlobal* gh, ShellGPUData* g);
void shell_gpu_global_pin_host(const Real* X, const Real* V, const Real* VR, Real* raw_gpu_to_cpu, int NUMNOD);
void shell_gpu_set_global(ShellGPUData* g, ShellGPUGlobal* gh);

/* per super-unit life cycle (shell_gpu_driver.h:82-90, shell_gpu_driver.cu:261-275) */
ShellGPUData* shell_gpu_data_create(void);
void shell_gpu_data_destroy(ShellGPUData* g);
void shell_gpu_allocate(ShellGPUData* g, int NUMELC, int NUMNOD, int NPT, int ISMSTR, int ITHK);
void shell_gpu_deallocate(ShellGPUData* g);

/* parameters (shell_gpu_driver.cu:280-350) */
void shell_gpu_set_mat_params(ShellGPUData* g, Real E, Real nu, Real G, Real A11, Real A12, Real CA, Real CB, Real CN, Real CC, Real EPDR,
                              Real EPMX, Real YMAX, Real M_EXP, Real FISOKIN, Real RHOCP, Real TREF, Real TMELT, Real ASRATE,
                              Real RHO, Real SSP, Real SHF_COEF, int IPLA, int VP, int IFORM, int ICC, Real Z3, Real Z4);
void shell_gpu_set_hg_params(ShellGPUData* g, Real H1, Real H2, Real H3, Real SRH1, Real SRH2, Real SRH3, Real HVISC, Real HELAS, Real HVLIN);
void shell_gpu_set_compute_sti(ShellGPUData* g, int flag);
void shell_gpu_set_ihbe(ShellGPUData* g, int ihbe);

/* one-time uploads (shell_gpu_driver.h:96-118): 0-based connectivity, per-point arrays flattened (IT-1)*NUMELC + e */
void shell_gpu_upload_constant(ShellGPUData* g, const int* h_N1, const int* h_N2, const int* h_N3, const int* h_N4, const Real* h_THK0,
                               const Real* h_OFF, const Real* h_SSP, const Real* h_RHO, const Real* h_YM, const Real* h_NU, const Real* h_A11,
                               const Real* h_G, const Real* h_SHF);
void shell_gpu_upload_ip_state(ShellGPUData* g, const Real* h_SIGxx, const Real* h_SIGyy, const Real* h_SIGxy, const Real* h_SIGyz, const Real* h_SIGzx,
                               const Real* h_PLA, const Real* h_EPSD_ip, const Real* h_SIGBAKxx, const Real* h_SIGBAKyy, const Real* h_SIGBAKxy,
                               const Real* h_TEMPEL);

/* per-cycle transfers and read-back (shell_gpu_driver.h:124-150; upload_nodes as compiled, shell_gpu_driver.cu:654) */
void shell_gpu_upload_nodes(ShellGPUData* g, const Real* X, const Real* V, const Real* VR);
void shell_gpu_download_nodal_forces(const ShellGPUData* g, Real* raw_gpu_to_cpu);
void shell_gpu_download_energy(const ShellGPUData* g, Real* h_EINT);                 /* [2][NUMELC] membrane, bending */
void shell_gpu_download_state(const ShellGPUData* g, Real* h_OFF, Real* h_THK, Real* h_GSTR, Real* h_EPSD_elem, Real* h_SIGxx, Real* h_SIGyy,
                              Real* h_SIGxy, Real* h_SIGyz, Real* h_SIGzx, Real* h_PLA, Real* h_EPSD_ip, Real* h_SIGBAKxx, Real* h_SIGBAKyy,
                              Real* h_SIGBAKxy, Real* h_TEMPEL);

/* execution (shell_gpu_driver.h:156-206) */
void shell_gpu_zero_nodal_arrays(ShellGPUData* g);
void shell_gpu_run_kernels(ShellGPUData* g, Real dt);
void shell_gpu_synchronize(ShellGPUData* g);
void shell_gpu_pin_host_memory(const ShellGPUData* g, Real* raw_cpu_to_gpu, Real* raw_gpu_to_cpu);
void shell_gpu_unpin_host_memory(Real* raw_cpu_to_gpu, Real* raw_gpu_to_cpu);
void shell_gpu_full_step(ShellGPUData* g, Real dt, const Real* raw_cpu_to_gpu, Real* raw_gpu_to_cpu);
void shell_gpu_full_step_async(ShellGPUData* g, Real dt, const Real* X, const Real* V, const Real* VR, Real* raw_gpu_to_cpu);
void shell_gpu_download_aldt_sq_async(const ShellGPUData* g, Real* h_aldt_sq);      /* not provided: exits */
void shell_gpu_download_aldt_sq(const ShellGPUData* g, Real* h_aldt_sq);            /* not provided: exits                     */
void shell_gpu_min_dt(ShellGPUData* g, Real dtfac, Real* h_dt_min);

#ifdef __cplusplus
}
#endif
#endif
