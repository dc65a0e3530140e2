"""TEST INFRASTRUCTURE ONLY: ctypes driver of the reference's OWN GPU shell path, compiled unmodified from
/root/reference/engine/source/elements/shell/coque/{shell_gpu_driver,shell_geometry_kernel,shell_strain_material_kernel,
shell_force_assembly_kernel}.cu into oracle/_ref/libshellgpu_ref.so (`make -C oracle refgpu`).

It covers Belytschko-Tsay shells with LAW2 only and differs from the CPU Engine by design (mid-point through-thickness
rule with weights 1/NPT instead of the Z0/WF/WM tables, fma(), atomics, no rupture), so it is NOT an oracle for the whole
path -- but wherever those differences vanish (membrane response of a flat plate: every integration point sees the same
strain) its nodal forces must agree with the restatement and with the CUDA path to rounding.  The call sequence is the
one of shell_internal_forces.F90 (gpu_shell_launch_async :62-176, gpu_shell_sync_scatter :177-292)."""
import ctypes as C
import os
import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_DIR, "_ref", "libshellgpu_ref.so")
R = C.c_double


def available():
    return os.path.exists(LIB)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class RefShellGPU:
    """One super-unit (all shell groups of the model: one BT property, one LAW2 material)."""

    def __init__(self, m, compute_sti=2, lib=None, asrate=1.0):
        """lib: another library exporting the same ABI (liborgpu.so exports it: include/shell_gpu_abi.h), default the
        reference's own; asrate: what goes into the ASRATE slot (the reference's kernels use it as the filter coefficient,
        its Fortran caller passes PM(9) = 2 pi Fcut, which is what liborgpu expects)"""
        assert m.numels == 0 and m.shell_groups and all(g.law == 2 for g in m.shell_groups)
        L = self.lib = C.CDLL(lib or LIB)
        for f in ("shell_gpu_global_create", "shell_gpu_data_create"):
            getattr(L, f).restype = C.c_void_p
        g0 = m.shell_groups[0]
        mat, prop = g0.mat, g0.prop
        self.n, self.ne, self.npt = m.numnod, m.numelc, prop.npt
        self.gh = C.c_void_p(L.shell_gpu_global_create(C.c_int(self.n)))
        self.g = C.c_void_p(L.shell_gpu_data_create())
        L.shell_gpu_allocate(self.g, C.c_int(self.ne), C.c_int(self.n), C.c_int(prop.npt), C.c_int(prop.ismstr), C.c_int(prop.ithk))
        L.shell_gpu_set_global(self.g, self.gh)
        m_exp = mat.z3 if mat.iform == 0 else 1.0
        args = [mat.young, mat.nu, mat.shear, mat.a11, mat.a12, mat.ca, mat.cb, mat.cn, mat.cc, mat.epdr, mat.epmx, mat.sigmx, m_exp,
                mat.fisokin, mat.rhocp, mat.tref, mat.tmelt, asrate, mat.rho0, mat.ssp, prop.shf]   # ASRATE: for the reference's kernels the filter COEFFICIENT itself (1 = unfiltered)
        L.shell_gpu_set_mat_params(self.g, *[R(float(a)) for a in args], C.c_int(prop.ipla), C.c_int(mat.vp), C.c_int(mat.iform), C.c_int(mat.icc),
                                   R(float(mat.z3)), R(float(mat.z4)))
        L.shell_gpu_set_hg_params(self.g, *[R(float(a)) for a in (prop.h1, prop.h2, prop.h3, prop.srh1, prop.srh2, prop.srh3, 0.5, 0.5, 0.0)])     # HVISC, HELAS, HVLIN: radioss2.F:641-643
        L.shell_gpu_set_compute_sti(self.g, C.c_int(compute_sti))
        L.shell_gpu_set_ihbe(self.g, C.c_int(prop.ihbe))
        ne = self.ne
        conn = [np.ascontiguousarray(m.ixc[:, 1 + k] - 1, np.int32) for k in range(4)]
        ones = lambda v: np.full(ne, float(v))
        self._keep = conn
        L.shell_gpu_upload_constant(self.g, *[_p(c) for c in conn], _p(ones(prop.thick)), _p(ones(1.0)), _p(ones(mat.ssp)), _p(ones(mat.rho0)),
                                    _p(ones(mat.young)), _p(ones(mat.nu)), _p(ones(mat.a11)), _p(ones(mat.shear)), _p(ones(prop.shf)))
        z = np.zeros(self.npt * ne); t = np.full(self.npt * ne, float(mat.tini))
        L.shell_gpu_upload_ip_state(self.g, *[_p(z)] * 10, _p(t))
        self.out = np.zeros(8 * self.n)

    def step(self, dt1, X, V, VR):
        """forces of one cycle from host nodal arrays (N,3): returns (N,8) = Fx,Fy,Fz,Mx,My,Mz,STIFN,STIFR"""
        L = self.lib
        X, V, VR = [np.ascontiguousarray(a, np.float64) for a in (X, V, VR)]
        L.shell_gpu_global_upload_nodes(self.gh, _p(X), _p(V), _p(VR))
        L.shell_gpu_global_wait_upload(self.gh, self.g)
        L.shell_gpu_run_kernels(self.g, R(float(dt1)))
        L.shell_gpu_global_wait_su(self.gh, self.g)
        L.shell_gpu_global_download_forces(self.gh, _p(self.out))
        L.shell_gpu_global_synchronize(self.gh)
        L.shell_gpu_synchronize(self.g)
        return self.out.reshape(8, self.n).T.copy()

    def step_raw(self, dt1, X, V, VR):
        """the same calls without any numpy work around them (timing): result stays in self.out (SoA)"""
        L = self.lib
        L.shell_gpu_global_upload_nodes(self.gh, _p(X), _p(V), _p(VR))
        L.shell_gpu_global_wait_upload(self.gh, self.g)
        L.shell_gpu_run_kernels(self.g, R(float(dt1)))
        L.shell_gpu_global_wait_su(self.gh, self.g)
        L.shell_gpu_global_download_forces(self.gh, _p(self.out))
        L.shell_gpu_global_synchronize(self.gh)

    def pin(self, X, V, VR):
        """shell_gpu_global_pin_host: page-lock the caller's nodal arrays and the force buffer (what FORINTC_PREPARE_GPU does)"""
        self._pinned = [np.ascontiguousarray(a, np.float64) for a in (X, V, VR)]
        self.lib.shell_gpu_global_pin_host(_p(self._pinned[0]), _p(self._pinned[1]), _p(self._pinned[2]), _p(self.out), C.c_int(self.n))
        return self._pinned

    def run_kernels_only(self, dt1):
        self.lib.shell_gpu_run_kernels(self.g, R(float(dt1)))

    def synchronize(self):
        self.lib.shell_gpu_synchronize(self.g)

    def min_dt(self, dtfac):
        out = np.zeros(1)
        self.lib.shell_gpu_min_dt(self.g, R(float(dtfac)), _p(out))
        self.lib.shell_gpu_synchronize(self.g)
        return float(out[0])

    def energy(self):
        out = np.zeros(2 * self.ne)
        self.lib.shell_gpu_download_energy(self.g, _p(out))
        return out.reshape(2, self.ne)

    def state(self):
        """shell_gpu_download_state: dict of OFF, THK, GSTR (8, ne), EPSD, SIG (npt, 5, ne), PLA (npt, ne), EPSD_ip, TEMP"""
        ne, nip = self.ne, self.npt * self.ne
        a = {k: np.zeros(n) for k, n in (("off", ne), ("thk", ne), ("gstr", 8 * ne), ("epsd", ne), ("sxx", nip), ("syy", nip), ("sxy", nip),
                                         ("syz", nip), ("szx", nip), ("pla", nip), ("epsd_ip", nip), ("bxx", nip), ("byy", nip), ("bxy", nip), ("temp", nip))}
        self.lib.shell_gpu_download_state(self.g, *[_p(v) for v in a.values()])
        sig = np.stack([a[k].reshape(self.npt, ne) for k in ("sxx", "syy", "sxy", "syz", "szx")], 1)
        return {"off": a["off"], "thk": a["thk"], "gstr": a["gstr"].reshape(8, ne), "epsd": a["epsd"], "sig": sig,
                "pla": a["pla"].reshape(self.npt, ne), "epsd_ip": a["epsd_ip"].reshape(self.npt, ne), "temp": a["temp"].reshape(self.npt, ne)}

    def close(self):
        L = self.lib
        L.shell_gpu_deallocate(self.g); L.shell_gpu_data_destroy(self.g); L.shell_gpu_global_destroy(self.gh)
