"""Parity of the CUDA 3-node shell path (C3FORC3, LAW36 / LAW2, through the C ABI) against the CPU oracle.
Tolerances are the north_star's: per-cycle nodal forces 1e-12 relative (fp64), after 1000 cycles displacements 1e-8
relative and energies 1e-8."""
import numpy as np
import pytest
import torch
from conftest import rel_err
from openradioss_b200 import meshgen

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from openradioss_b200.engine import Engine
    from oracle.orc import Oracle

FORCE_TOL = 1e-12
DISP_TOL = 1e-8
ENERGY_TOL = 1e-8
STATE3 = ("forc", "mom", "eint", "thk", "off", "stra", "epsd", "smstr", "sig", "pla", "epsd_ip")
STATE4 = ("forc", "mom", "eint", "thk", "off", "stra", "epsd", "hourg", "smstr", "sig", "pla", "epsd_ip")


def pair(m):
    return Engine(m), Oracle(m, threads=0)


def check_state(g, o, m, tol=1e-11):
    for f in STATE3:
        a, b = g.sh3n_state(f), o.sh3n_state(f)
        assert rel_err(a, b) <= tol, ("sh3n", f, rel_err(a, b))
    if m.numelc:
        for f in STATE4:
            a, b = g.shell_state(f), o.shell_state(f)
            assert rel_err(a, b) <= tol, ("shell", f, rel_err(a, b))


def cycle_check(m, ncheck=3, state_tol=1e-11, typ=(7,)):
    g, o = pair(m)
    dt1 = 0.0
    for c in range(ncheck):
        for b in (g, o):
            b.forces_phase(dt1)
        fg, fo = g.download_fsky(), o.download_fsky()
        assert rel_err(fg[:, :3], fo[:, :3]) <= FORCE_TOL, ("F", c, rel_err(fg[:, :3], fo[:, :3]))
        assert rel_err(fg[:, 3:6], fo[:, 3:6]) <= FORCE_TOL, ("M", c, rel_err(fg[:, 3:6], fo[:, 3:6]))
        assert rel_err(fg[:, 6:], fo[:, 6:]) <= FORCE_TOL, ("STI", c)
        tg, to = g.time(), o.time()
        assert tg["dt2t"] == pytest.approx(to["dt2t"], rel=1e-13) and tg["neltst"] == to["neltst"]
        assert tg["ityptst"] == to["ityptst"] and to["ityptst"] in typ
        for b in (g, o):
            b.assemble()
        ng, no = g.download_nodes(("A", "AR", "STIFN")), o.download_nodes(("A", "AR", "STIFN"))
        for k in ("A", "AR", "STIFN"):
            assert rel_err(ng[k], no[k]) <= FORCE_TOL, (k, c)
        dt2 = to["dt2t"]
        for b in (g, o):
            b.advance(0.5 * (dt1 + dt2), dt2)
        ng, no = g.download_nodes(("X", "V", "VR", "D")), o.download_nodes(("X", "V", "VR", "D"))
        for k in ("X", "V", "VR", "D"):
            assert rel_err(ng[k], no[k]) <= 1e-13, (k, c)
        check_state(g, o, m, state_tol)
        dt1 = dt2
    return g, o


@pytest.mark.parametrize("shape", [(6, 5), (1, 1), (13, 11), (32, 9)])
def test_c3_law36_phases_match_oracle(shape):
    nx, ny = shape
    m = meshgen.tri_plate(nx, ny, 10.0 * nx, 10.0 * ny, pressure=20.0, vrand=3.0, user_id_perm=True, clamp=nx > 1)
    cycle_check(m)


@pytest.mark.parametrize("ipla", [0, 1, 2])
@pytest.mark.parametrize("npt", [1, 3, 5])
def test_c3_law36_iplas_npt(ipla, npt):
    prop = meshgen.default_prop_shell(thick=1.5, ihbe=2, npt=npt, ipla=ipla)
    m = meshgen.tri_plate(7, 6, 70.0, 60.0, prop=prop, pressure=50.0, vrand=40.0)
    g, o = cycle_check(m, ncheck=4)
    assert o.sh3n_state("pla").max() > 0.0


@pytest.mark.parametrize("ismstr", [1, 2, 4])
@pytest.mark.parametrize("ish3n", [1, 2])
def test_c3_ismstr_and_ish3n(ismstr, ish3n):
    prop = meshgen.default_prop_shell(ihbe=ish3n, ismstr=ismstr)
    m = meshgen.tri_plate(6, 6, 60.0, 60.0, prop=prop, pressure=10.0, vrand=5.0)
    cycle_check(m, ncheck=4)


def test_c3_law2_johnson_cook():
    m = meshgen.tri_plate(7, 7, 70.0, 70.0, law=2, pressure=30.0, vrand=30.0)
    g, o = cycle_check(m, ncheck=5, state_tol=1e-10)
    assert o.sh3n_state("pla").max() > 0.0


def test_c3_rate_dependent_curves_long_table_in_global_memory():
    """three curves of 20 points each: more than the 48 points the kernel parameters carry -> the global-memory VINTER path"""
    x = np.concatenate([[0.0], np.geomspace(1e-3, 0.5, 19)]); y = 250.0 + 230.0 * x ** 0.4
    curves, rates = [(x, y), (x, 1.15 * y), (x, 1.4 * y)], [0.0, 0.5, 50.0]
    m = meshgen.tri_plate(8, 7, 80.0, 70.0, pressure=30.0, vrand=30.0, curves=curves, rates=rates)
    g, o = cycle_check(m, ncheck=5)
    assert o.sh3n_state("pla").max() > 0.0


def test_c3_law36_vp1_plastic_strain_rate():
    x = np.array([0.0, 0.01, 0.03, 0.08, 0.2, 0.5]); y = np.array([250.0, 300.0, 340.0, 390.0, 440.0, 480.0])
    prop = meshgen.default_prop_shell(thick=1.5, ihbe=2, npt=5)
    m = meshgen.tri_plate(8, 7, 80.0, 70.0, prop=prop, pressure=40.0, vrand=40.0, curves=[(x, y), (x, 1.15 * y), (x, 1.4 * y)], rates=[0.0, 0.5, 50.0])
    for grp in m.sh3n_groups: grp.mat.vp = 1
    g, o = cycle_check(m, ncheck=8)
    assert o.sh3n_state("pla").max() > 0.01 and o.sh3n_state("plap").max() > 0.0
    assert rel_err(g.sh3n_state("plap"), o.sh3n_state("plap")) <= 1e-11


@pytest.mark.parametrize("npt", [1, 5])
def test_c3_rate_dependent_three_pass(npt):
    """three short curves (kernel parameters), byte cursors, the three-pass loop of the triangle kernel"""
    x = np.array([0.0, 0.01, 0.03, 0.08, 0.2, 0.5]); y = np.array([250.0, 300.0, 340.0, 390.0, 440.0, 480.0])
    prop = meshgen.default_prop_shell(thick=1.5, ihbe=2, npt=npt)
    m = meshgen.tri_plate(8, 7, 80.0, 70.0, prop=prop, pressure=40.0, vrand=40.0, curves=[(x, y), (x, 1.15 * y), (x, 1.4 * y)], rates=[0.0, 0.5, 50.0])
    g, o = cycle_check(m, ncheck=8)
    assert o.sh3n_state("pla").max() > 0.01


@pytest.mark.parametrize("ihbe", [24, 1])
def test_triangles_and_quads_share_nodes(ihbe):
    """checkerboard of 4-node and 3-node shells: one skyline, the Starter's slot order (quads before triangles at a node),
    the arg-min may come from either family (ITYPTST 3 or 7)."""
    qp = meshgen.default_prop_shell(ihbe=ihbe, npt=5)
    m = meshgen.tri_plate(9, 8, 90.0, 80.0, quads="checker", quad_prop=qp, pressure=25.0, vrand=20.0, user_id_perm=True)
    assert m.numelc > 0 and m.numeltg > 0
    cycle_check(m, ncheck=4, typ=(3, 7))


def energies(b, m):
    d = b.download_nodes(("V", "VR"))
    ke = 0.5 * (m.MS[:, None] * d["V"] ** 2).sum() + 0.5 * (m.IN[:, None] * d["VR"] ** 2).sum()
    ie = b.sh3n_state("eint").sum() + (b.shell_state("eint").sum() if m.numelc else 0.0)
    return ke, ie


@pytest.mark.parametrize("quads", ["none", "checker"])
def test_c3_plate_1000_cycles(quads):
    m = meshgen.tri_plate(20, 20, 200.0, 200.0, pressure=2.0, quads=quads)
    g, o = pair(m)
    g.run_cycles(1000); o.run_cycles(1000)
    ng, no = g.download_nodes(("D", "X", "V")), o.download_nodes(("D", "X", "V"))
    assert rel_err(ng["D"], no["D"]) <= DISP_TOL
    (keg, ieg), (keo, ieo) = energies(g, m), energies(o, m)
    assert abs(ieg - ieo) <= ENERGY_TOL * abs(ieo) and abs(keg - keo) <= ENERGY_TOL * max(abs(keo), abs(ieo))
    tg, to = g.time(), o.time()
    assert tg["ncycle"] == to["ncycle"] == 1000 and tg["tt"] == pytest.approx(to["tt"], rel=1e-10)
    assert o.sh3n_state("pla").max() > 0.0
    wext = (m.fext * ng["D"]).sum()
    assert abs(ieg + keg - wext) <= 0.02 * wext
    e4 = g.energies()
    assert e4[1] == pytest.approx(ieg, rel=1e-12)          # the device energy sum counts the triangles with the shells


def test_c3_nodal_time_step():
    """/DT/NODA: C3DT3's NODADT branch (nodal stiffnesses, no element dt), DTNODA after the assembly"""
    m = meshgen.tri_plate(8, 8, 80.0, 80.0, pressure=20.0, vrand=10.0, quads="checker")
    m.control.nodadt = 1; m.control.dtfac_node = 0.9
    g, o = pair(m)
    dt1 = 0.0
    for c in range(4):
        for b in (g, o):
            b.forces_phase(dt1); b.assemble()
        ng, no = g.download_nodes(("A", "AR", "STIFN", "STIFR")), o.download_nodes(("A", "AR", "STIFN", "STIFR"))
        for k in ("A", "AR", "STIFN", "STIFR"):
            assert rel_err(ng[k], no[k]) <= FORCE_TOL, (k, c)
        tg, to = g.time(), o.time()
        assert tg["dt2t"] == pytest.approx(to["dt2t"], rel=1e-13) and tg["neltst"] == to["neltst"] and tg["ityptst"] == to["ityptst"] == 11
        dt2 = to["dt2t"]
        for b in (g, o):
            b.advance(0.5 * (dt1 + dt2), dt2)
        dt1 = dt2
    g.run_cycles(200); o.run_cycles(200)
    assert rel_err(g.download_nodes(("D",))["D"], o.download_nodes(("D",))["D"]) <= DISP_TOL


def test_c3_law36_tensile_strain_failure_deletes_the_same_triangles():
    """LAW36 IFAIL = 2 through C3FORC3 -> CMAIN3 -> MULAWC -> SIGEPS36C (sigeps36c.F:256-264, 940-950)"""
    mat, npf, tf = meshgen.steel_law36(eps_t=(1.5e-3, 8.0e-3, 4.0e-3))
    m = meshgen.tri_plate(10, 9, 100.0, 90.0, mat=mat, pressure=60.0, vrand=8.0)
    m.npf, m.tf = npf, tf
    g, o = pair(m)
    g.run_cycles(5); o.run_cycles(5)
    check_state(g, o, m, tol=1e-10)
    g.run_cycles(5); o.run_cycles(5)
    dead = []
    for c in range(4):
        assert np.array_equal(g.sh3n_state("off"), o.sh3n_state("off")), c
        dead.append(int((o.sh3n_state("off") == 0).sum()))
        ng, no = g.download_nodes(("X", "V", "VR")), o.download_nodes(("X", "V", "VR"))
        for k in ("X", "V", "VR"):
            assert rel_err(ng[k], no[k]) <= 1e-9, (k, c)
        g.run_cycles(10); o.run_cycles(10)
    assert 0 < dead[-1] < m.numeltg and dead[-1] > dead[0]


def test_c3_large_plate_properties():
    """500 k triangles: size-independent properties (self-equilibrated elements, finite state, reproducible checksum)"""
    import zlib
    m = meshgen.tri_plate(500, 500, 5000.0, 5000.0, pressure=5.0, vrand=1.0)
    sums = []
    for _ in range(2):
        g = Engine(m)
        g.run_cycles(20)
        g.forces_phase(g.time()["dt2"])
        f = g.download_fsky()
        sums.append(zlib.adler32(np.ascontiguousarray(f).tobytes()))
    assert sums[0] == sums[1]
    assert np.isfinite(f).all()
    F = f[m.iadtg - 1][:, :, :3]
    assert np.abs(F.sum(1)).max() <= 1e-9 * np.abs(F).max()


def test_bad_sh3n_inputs_are_rejected():
    m = meshgen.tri_plate(2, 2, 20.0, 20.0)
    m.sh3n_groups[0].prop.ihbe = 30                     # DKT18: CDKFORC3, not built
    with pytest.raises(RuntimeError, match="outside the built path"):
        Engine(m)
