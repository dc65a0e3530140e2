"""liborgpu.so as a literal drop-in for the reference's `-gpu` shell path: it exports the reference's own C ABI
(engine/source/elements/shell/coque/shell_gpu_driver.h:44-206, bound by shell_gpu_mod.F90:284-715) -- include/shell_gpu_abi.h,
openradioss_b200/csrc/shell_gpu_compat.cuh.  The SAME ctypes driver (oracle/refgpu.py: the call sequence of
shell_internal_forces.F90) runs the reference's library (built unmodified into oracle/_ref) and this one."""
import ctypes, os, re
import numpy as np
import pytest
import torch
from refgpu_cases import plate

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_HEADER = "/root/reference/engine/source/elements/shell/coque/shell_gpu_driver.h"
ORGPU = os.path.join(ROOT, "openradioss_b200", "csrc", "liborgpu.so")


def _declared(path):
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S); src = re.sub(r"//[^\n]*", "", src)
    return sorted(set(re.findall(r"\b(shell_gpu_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_the_reference_abi():
    names = _declared(os.path.join(ROOT, "include", "shell_gpu_abi.h"))
    assert len(names) == 33
    lib = ctypes.CDLL(ORGPU)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


@pytest.mark.skipif(not os.path.exists(REF_HEADER), reason="the reference tree is not present (GPU box)")
def test_header_is_the_reference_header_name_for_name():
    ref = _declared(REF_HEADER)
    ours = _declared(os.path.join(ROOT, "include", "shell_gpu_abi.h"))
    assert set(ref) <= set(ours), sorted(set(ref) - set(ours))
    extra = set(ours) - set(ref)          # entry points the driver defines and shell_gpu_mod.F90 binds without a line in the header
    assert extra == {"shell_gpu_data_create", "shell_gpu_data_destroy", "shell_gpu_set_mat_params", "shell_gpu_set_hg_params"}
    drv = open(REF_HEADER.replace(".h", ".cu")).read()
    mod = open(REF_HEADER.replace("shell_gpu_driver.h", "shell_gpu_mod.F90")).read()
    for n in extra:
        assert re.search(r"\b%s\s*\(" % n, drv) and ('name="%s"' % n in mod or "name='%s'" % n in mod), n
    bound = set(re.findall(r"name\s*=\s*[\"'](shell_gpu_[a-z_0-9]+)[\"']", mod))    # everything the Fortran side binds is exported
    assert bound <= set(ours), sorted(bound - set(ours))


def test_without_a_gpu_the_drop_in_fails_loudly():
    """no CPU fallback behind the reference's ABI either: handles and parameters are recorded without a device, the first
    per-cycle call prints the reason and exits with status 1 (the reference's own error convention, shell_gpu_driver.cu:47-55)"""
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import subprocess, sys, textwrap
    code = textwrap.dedent(f"""
        import ctypes as C, numpy as np
        L = C.CDLL({ORGPU!r}); R = C.c_double
        L.shell_gpu_global_create.restype = C.c_void_p; L.shell_gpu_data_create.restype = C.c_void_p
        gh = C.c_void_p(L.shell_gpu_global_create(C.c_int(4))); g = C.c_void_p(L.shell_gpu_data_create())
        L.shell_gpu_allocate(g, C.c_int(1), C.c_int(4), C.c_int(3), C.c_int(2), C.c_int(0)); L.shell_gpu_set_global(g, gh)
        L.shell_gpu_set_mat_params(g, *[R(v) for v in (210e3, .3, 80e3, 230e3, 69e3, 250., 400., .4, 0., 1e-3, 1e30, 1e30, 1., 0., 0., 300., 1e30, 0., 7.85e-3, 5400., 5 / 6)],
                                   C.c_int(1), C.c_int(2), C.c_int(0), C.c_int(1), R(0.), R(0.))
        L.shell_gpu_set_hg_params(g, *[R(v) for v in (.01, .01, .01, .1, .1, .1, .5, .5, 0.)])
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        n = [np.array([k], np.int32) for k in range(4)]; one = lambda v: np.array([float(v)])
        arrs = [one(1.2), one(1.), one(5400.), one(7.85e-3), one(210e3), one(.3), one(230e3), one(80e3), one(5 / 6)]
        L.shell_gpu_upload_constant(g, *[p(a) for a in n], *[p(a) for a in arrs])
        print("recorded", flush=True)
        X = np.zeros((4, 3)); L.shell_gpu_global_upload_nodes(gh, p(X), p(X), p(X))
        print("not reached", flush=True)
    """)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "recorded" in r.stdout and "not reached" not in r.stdout
    assert "liborgpu shell_gpu ABI" in r.stderr


gpu = pytest.mark.gpu
if torch.cuda.is_available():
    from openradioss_b200.engine import Engine
    from oracle import refgpu


def drop_in(m, **kw):
    mat = m.shell_groups[0].mat
    return refgpu.RefShellGPU(m, lib=ORGPU, asrate=mat.asrate, **kw)


@gpu
@pytest.mark.parametrize("ipla,npt,rate,shear", [(1, 5, True, True), (0, 3, False, False), (2, 3, True, False)])
def test_reference_call_sequence_gives_the_engine_forces_bitwise(ipla, npt, rate, shear):
    """global_upload_nodes -> run_kernels -> global_download_forces through the reference's ABI = forces_phase + assemble
    of the orgpu ABI on the same nodal arrays: A, AR, STIFN, STIFR bit for bit; shell_gpu_min_dt = DTFAC1(3) x the CDT3 minimum"""
    m = plate(ipla, npt, rate, shear)
    g, d = Engine(m), drop_in(m)
    dt1 = 0.0
    for c in range(8):
        nd = g.download_nodes(("X", "V", "VR"))
        f = d.step(dt1, nd["X"], nd["V"], nd["VR"])
        g.forces_phase(dt1); g.assemble()
        a = g.download_nodes(("A", "AR", "STIFN", "STIFR"))
        assert np.array_equal(f[:, :3], a["A"]) and np.array_equal(f[:, 3:6], a["AR"]), c
        assert np.array_equal(f[:, 6], a["STIFN"]) and np.array_equal(f[:, 7], a["STIFR"]), c
        assert np.abs(a["A"]).max() > 0.0 or c == 0
        dt2 = g.time()["dt2t"]
        assert d.min_dt(m.control.dtfac_shell) == pytest.approx(dt2, rel=1e-15)
        g.advance(0.5 * (dt1 + dt2), dt2); dt1 = dt2
    assert g.shell_state("pla").max() > 1e-3
    # shell_gpu_download_state / _energy: the element buffer in the reference's layout
    st, en = d.state(), d.energy()
    assert np.array_equal(en, g.shell_state("eint"))
    assert np.array_equal(st["off"], g.shell_state("off")[0]) and np.array_equal(st["thk"], g.shell_state("thk")[0])
    assert np.array_equal(st["gstr"], g.shell_state("stra")) and np.array_equal(st["epsd"], g.shell_state("epsd")[0])
    assert np.array_equal(st["sig"].reshape(-1, m.numelc), g.shell_state("sig"))
    assert np.array_equal(st["pla"], g.shell_state("pla")) and np.array_equal(st["epsd_ip"], g.shell_state("epsd_ip"))
    assert np.all(st["temp"] == m.shell_groups[0].mat.tini)
    d.close()


@gpu
@pytest.mark.skipif(not (torch.cuda.is_available() and refgpu.available()) if torch.cuda.is_available() else True, reason="oracle/_ref/libshellgpu_ref.so not built")
@pytest.mark.parametrize("ipla,npt,rate,shear", [(1, 5, True, True), (0, 3, True, False), (2, 3, False, True)])
def test_same_driver_two_libraries(ipla, npt, rate, shear):
    """the reference's library and liborgpu.so behind the same calls, on the states where the reference GPU path and the
    CPU Engine coincide (tests/test_ref_gpu_pin.py): nodal forces and moments agree to 1e-12 of the largest force"""
    m = plate(ipla, npt, rate, shear)
    r, d, g = refgpu.RefShellGPU(m), drop_in(m), Engine(m)
    dt1 = 0.0
    for c in range(8):
        nd = g.download_nodes(("X", "V", "VR"))
        fr = r.step(dt1, nd["X"], nd["V"], nd["VR"])
        fd = d.step(dt1, nd["X"], nd["V"], nd["VR"])
        if c > 0:
            sf = np.abs(fr[:, :3]).max()
            assert sf > 0.0 and np.abs(fr[:, :3] - fd[:, :3]).max() <= 1e-12 * sf, c
            if shear:
                assert np.abs(fr[:, 3:6] - fd[:, 3:6]).max() <= 1e-12 * max(np.abs(fr[:, 3:6]).max(), 1e-300), c
        g.forces_phase(dt1); g.assemble()
        dt2 = g.time()["dt2t"]
        g.advance(0.5 * (dt1 + dt2), dt2); dt1 = dt2
    d.close()


@gpu
@pytest.mark.parametrize("ihbe,compute_sti", [(4, 2), (1, 1), (3, 0)])
def test_thickness_state_ishell_and_sti_modes(ihbe, compute_sti):
    """ITHK = 1 with a per-element initial thickness (shell_gpu_upload_constant h_THK0 -> GBUF%THK), Ishell 3 / 4,
    compute_sti 1 (nodal stiffnesses of the NODADT /= 0 branch, shell_internal_forces.F90:637-648) and 0 (none)"""
    m = plate(1, 3, True, True)
    for sg in m.shell_groups:
        sg.prop.ithk = 1; sg.prop.ihbe = ihbe
        if ihbe == 3:
            sg.prop.h1 = sg.prop.h2 = 0.1; sg.prop.srh1 = sg.prop.srh2 = float(np.sqrt(0.1))
    m.control.nodadt = 1 if compute_sti == 1 else 0
    thk = 1.2 + 0.3 * np.random.default_rng(3).uniform(size=m.numelc)
    g = Engine(m)
    g.upload_shell_state("thk", thk[None, :])
    mat, prop = m.shell_groups[0].mat, m.shell_groups[0].prop

    d = drop_in(m, compute_sti=compute_sti)
    # the driver uploaded a uniform thickness: give the handle the per-element one before the first cycle
    import ctypes as C
    ne = m.numelc
    conn = [np.ascontiguousarray(m.ixc[:, 1 + k] - 1, np.int32) for k in range(4)]
    ones = lambda v: np.full(ne, float(v))
    arrs = [thk.copy(), ones(1.0), ones(mat.ssp), ones(mat.rho0), ones(mat.young), ones(mat.nu), ones(mat.a11), ones(mat.shear), ones(prop.shf)]
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    d.lib.shell_gpu_upload_constant(d.g, *[p(c) for c in conn], *[p(a) for a in arrs])
    dt1 = 0.0
    for c in range(6):
        nd = g.download_nodes(("X", "V", "VR"))
        f = d.step(dt1, nd["X"], nd["V"], nd["VR"])
        g.forces_phase(dt1); g.assemble()
        a = g.download_nodes(("A", "AR", "STIFN", "STIFR"))
        assert np.array_equal(f[:, :3], a["A"]) and np.array_equal(f[:, 3:6], a["AR"]), c
        if compute_sti == 0:
            assert np.all(f[:, 6:] == 0.0)
        else:
            assert np.array_equal(f[:, 6], a["STIFN"]) and np.array_equal(f[:, 7], a["STIFR"]) and a["STIFN"].max() > 0.0, c
        dt2 = g.time()["dt2t"]
        if compute_sti == 2:
            assert d.min_dt(m.control.dtfac_shell) == pytest.approx(dt2, rel=1e-15)
        else:
            assert d.min_dt(m.control.dtfac_shell) == 1e30
        dt2 = min(dt2, 2.0e-4)
        g.advance(0.5 * (dt1 + dt2), dt2); dt1 = dt2
    assert np.abs(a["A"]).max() > 0.0 and not np.array_equal(d.state()["thk"], thk)      # ITHK = 1: the thickness evolves
    assert np.array_equal(d.state()["thk"], g.shell_state("thk")[0])
    d.close()


@gpu
def test_two_super_units_on_one_global_handle():
    """two ShellGPUData (different thickness) attached to one ShellGPUGlobal, launched one after the other as
    gpu_shell_launch_async does: the sum of both lands in the one force buffer, equal to the single-engine model"""
    import ctypes as C
    m = plate(1, 3, True, True)
    half = (m.numelc // 2)
    thk = np.where(np.arange(m.numelc) < half, 1.2, 1.5)
    # engine model with the same two thicknesses: groups split at `half`
    from openradioss_b200.model import ShellGroup
    import copy
    g0 = m.shell_groups[0]
    pa, pb = copy.deepcopy(g0.prop), copy.deepcopy(g0.prop); pb.thick = 1.5
    m.shell_groups = [ShellGroup(nel=half, nft=0, law=2, mat=g0.mat, prop=pa), ShellGroup(nel=m.numelc - half, nft=half, law=2, mat=g0.mat, prop=pb)]
    g = Engine(m)
    L = C.CDLL(ORGPU)
    L.shell_gpu_global_create.restype = C.c_void_p; L.shell_gpu_data_create.restype = C.c_void_p
    R, p = C.c_double, (lambda a: a.ctypes.data_as(C.c_void_p))
    gh = C.c_void_p(L.shell_gpu_global_create(C.c_int(m.numnod)))
    mat, keep, sus = g0.mat, [], []
    for lo, hi, prop in ((0, half, pa), (half, m.numelc, pb)):
        su = C.c_void_p(L.shell_gpu_data_create()); ne = hi - lo
        L.shell_gpu_allocate(su, C.c_int(ne), C.c_int(m.numnod), C.c_int(prop.npt), C.c_int(prop.ismstr), C.c_int(prop.ithk))
        L.shell_gpu_set_global(su, gh)
        args = [mat.young, mat.nu, mat.shear, mat.a11, mat.a12, mat.ca, mat.cb, mat.cn, mat.cc, mat.epdr, mat.epmx, mat.sigmx, mat.z3,
                mat.fisokin, mat.rhocp, mat.tref, mat.tmelt, mat.asrate, mat.rho0, mat.ssp, prop.shf]
        L.shell_gpu_set_mat_params(su, *[R(float(a)) for a in args], C.c_int(prop.ipla), C.c_int(mat.vp), C.c_int(mat.iform), C.c_int(mat.icc), R(0.0), R(0.0))
        L.shell_gpu_set_hg_params(su, *[R(float(a)) for a in (prop.h1, prop.h2, prop.h3, prop.srh1, prop.srh2, prop.srh3, 0.5, 0.5, 0.0)])
        L.shell_gpu_set_compute_sti(su, C.c_int(2)); L.shell_gpu_set_ihbe(su, C.c_int(prop.ihbe))
        conn = [np.ascontiguousarray(m.ixc[lo:hi, 1 + k] - 1, np.int32) for k in range(4)]
        ones = lambda v: np.full(ne, float(v))
        arrs = [ones(prop.thick), ones(1.0), ones(mat.ssp), ones(mat.rho0), ones(mat.young), ones(mat.nu), ones(mat.a11), ones(mat.shear), ones(prop.shf)]
        keep += conn + arrs
        L.shell_gpu_upload_constant(su, *[p(c) for c in conn], *[p(a) for a in arrs])
        z = np.zeros(prop.npt * ne); t = np.full(prop.npt * ne, float(mat.tini)); keep += [z, t]
        L.shell_gpu_upload_ip_state(su, *[p(z)] * 10, p(t))
        sus.append(su)
    out = np.zeros(8 * m.numnod)
    dt1 = 0.0
    for c in range(6):
        nd = g.download_nodes(("X", "V", "VR"))
        X, V, VR = [np.ascontiguousarray(nd[k]) for k in ("X", "V", "VR")]
        L.shell_gpu_global_upload_nodes(gh, p(X), p(V), p(VR))
        for su in sus:
            L.shell_gpu_full_step_async(su, R(dt1), p(X), p(V), p(VR), p(out))
        for su in sus:
            L.shell_gpu_global_wait_su(gh, su)
        L.shell_gpu_global_download_forces(gh, p(out)); L.shell_gpu_global_synchronize(gh)
        f = out.reshape(8, m.numnod).T
        g.forces_phase(dt1); g.assemble()
        a = g.download_nodes(("A", "AR", "STIFN", "STIFR"))
        assert np.array_equal(f[:, :3], a["A"]) and np.array_equal(f[:, 3:6], a["AR"]) and np.array_equal(f[:, 6], a["STIFN"]), c
        dt2 = g.time()["dt2t"]
        g.advance(0.5 * (dt1 + dt2), dt2); dt1 = dt2
    assert np.abs(a["A"]).max() > 0.0
    e2 = np.zeros(2 * (m.numelc - half)); L.shell_gpu_download_energy(sus[1], p(e2))
    assert np.array_equal(e2.reshape(2, -1), g.shell_state("eint")[:, half:])
    for su in sus:
        L.shell_gpu_deallocate(su); L.shell_gpu_data_destroy(su)
    L.shell_gpu_global_destroy(gh)
