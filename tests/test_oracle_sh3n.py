"""Analytic pins for the CPU oracle's 3-node shell path (C3FORC3 restatement, oracle/shell_c3.cpp).
The reference holds no routine-level vectors for this path: closed-form patch tests."""
import numpy as np
import pytest
from openradioss_b200 import meshgen
from oracle.orc import Oracle


def flat(nx=4, ny=3, **kw):
    kw.setdefault("zjitter", 0.0); kw.setdefault("pressure", 0.0); kw.setdefault("clamp", False)
    return meshgen.tri_plate(nx, ny, 10.0 * nx, 10.0 * ny, **kw)


def test_biaxial_strain_rate_gives_plane_stress_hooke():
    """In-plane V = eps_dot * X on a flat mesh: after one step every integration point holds sxx = syy = E eps / (1 - nu),
    sxy = 0 in any in-plane frame (the triangle's frame follows its edge 1-2)."""
    m = flat()
    rate = 1e-4
    m.V = np.zeros_like(m.X); m.V[:, :2] = rate * m.X[:, :2]
    o = Oracle(m)
    dt1 = 1e-3
    o.forces_phase(dt1)
    mat = m.sh3n_groups[0].mat
    sig = o.sh3n_state("sig")                       # (5*npt, numeltg)
    exp = mat.young * rate * dt1 / (1.0 - mat.nu)
    npt = m.sh3n_groups[0].prop.npt
    for ip in range(npt):
        assert np.allclose(sig[5 * ip], exp, rtol=1e-9) and np.allclose(sig[5 * ip + 1], exp, rtol=1e-9)
        assert np.abs(sig[5 * ip + 2]).max() < 1e-9 * exp
    assert np.all(o.sh3n_state("pla") == 0.0)
    # membrane force per unit thickness = the same stress; no moments
    assert np.allclose(o.sh3n_state("forc")[0], exp, rtol=1e-9)
    assert np.abs(o.sh3n_state("mom")).max() < 1e-9 * exp


def test_rigid_body_motion_gives_no_force():
    m = flat(zjitter=0.08)
    w = np.array([0.3, -0.2, 0.5]) * 1e-3
    m.V = np.cross(np.tile(w, (m.numnod, 1)), m.X) + np.array([1.0, -2.0, 0.5]) * 1e-2
    m.VR = np.tile(w, (m.numnod, 1))
    o = Oracle(m)
    o.forces_phase(1e-3)
    f = o.download_fsky()
    # scale: what the same velocity magnitude would give as a stretching field
    scale = m.sh3n_groups[0].mat.young * 1e-3 * 1e-3 * 2.0 * 10.0
    assert np.abs(f[:, :6]).max() < 1e-6 * scale


def test_element_forces_are_self_equilibrated():
    m = flat(5, 4, zjitter=0.08, vrand=5.0)
    o = Oracle(m)
    o.forces_phase(0.0); o.forces_phase(1e-3)
    f = o.download_fsky()
    rows = f[m.iadtg - 1]                            # (ne, 3, 8)
    F = rows[:, :, :3]; M = rows[:, :, 3:6]
    assert np.abs(F.sum(1)).max() <= 1e-10 * np.abs(F).max()
    Xc = m.X[m.ixtg[:, 1:4] - 1]
    tot = M.sum(1) + np.cross(Xc, F).sum(1)          # moment balance about the origin
    assert np.abs(tot).max() <= 1e-9 * (np.abs(M).max() + np.abs(np.cross(Xc, F)).max())


def test_time_step_of_a_right_triangle():
    """dt = DTFAC1(7) * (2 A / longest edge) / c, c = the sound speed SIGEPS36C returns (UPARAM: sqrt(E / (1 - nu^2) / rho));
    arg-min type 7."""
    m = meshgen.tri_plate(1, 1, 10.0, 10.0, jitter=0.0, zjitter=0.0, pressure=0.0, clamp=False)
    o = Oracle(m)
    o.forces_phase(0.0)
    t = o.time()
    mat = m.sh3n_groups[0].mat
    aldt = 2.0 * 50.0 / np.sqrt(200.0)
    assert t["ityptst"] == 7
    assert t["dt2t"] == pytest.approx(0.9 * aldt / mat.soundsp, rel=1e-12)


def test_uniform_bending_rate_gives_plate_moment():
    """VR = (0, kappa_dot * x, 0) with the matching deflection rate: curvature kxx only.  Through the thickness sxx is
    antisymmetric, the membrane force vanishes and the moment per t^2 is E t kappa / (12 (1 - nu^2)) (frame of the
    x-aligned triangles)."""
    m = flat(4, 3, jitter=0.0)
    kd = 1e-6; dt1 = 1e-3
    m.V = np.zeros_like(m.X); m.V[:, 2] = -0.5 * kd * m.X[:, 0] ** 2
    m.VR = np.zeros_like(m.X); m.VR[:, 1] = kd * m.X[:, 0]
    o = Oracle(m)
    o.forces_phase(dt1)
    sig = o.sh3n_state("sig"); npt = m.sh3n_groups[0].prop.npt
    top, bot, mid = sig[5 * (npt - 1)], sig[0], sig[5 * (npt // 2)]
    assert np.allclose(top, -bot, rtol=1e-4) and np.abs(mid).max() < 1e-4 * np.abs(top).max()
    assert np.abs(o.sh3n_state("forc")[:3]).max() < 1e-4 * np.abs(top).max()
    mat, t = m.sh3n_groups[0].mat, m.sh3n_groups[0].prop.thick
    assert np.allclose(o.sh3n_state("mom")[0, ::2], mat.young * t * kd * dt1 / (12.0 * (1.0 - mat.nu ** 2)), rtol=1e-6)


def test_openmp_groups_give_identical_results():
    m = meshgen.tri_plate(12, 12, 120.0, 120.0, vrand=5.0, quads="checker")
    a, b = Oracle(m, threads=1), Oracle(m, threads=0)
    a.run_cycles(20); b.run_cycles(20)
    for k in ("X", "V", "VR"):
        assert np.array_equal(a.download_nodes((k,))[k], b.download_nodes((k,))[k])
    assert a.time()["neltst"] == b.time()["neltst"]


def test_law36_epsmax_failure_deletes_the_triangle_in_one_cycle():
    """IFAIL = 1 on shells: the first integration point beyond EPSMAX sets OFF = 0.8 and MULAWC finishes the deletion in
    the same cycle (sigeps36c.F:928-938, mulawc.F90:2937-2941): forces vanish, OFFG = 0 from then on."""
    mat, npf, tf = meshgen.steel_law36(epsmax=1.0e-3)
    m = meshgen.tri_plate(1, 1, 10.0, 10.0, mat=mat, jitter=0.0, zjitter=0.0, pressure=0.0, clamp=False)
    m.npf, m.tf = npf, tf
    m.V = np.zeros_like(m.X); m.V[:, 0] = 0.5 * m.X[:, 0]             # uniaxial stretching at 0.5 / ms
    o = Oracle(m)
    offs, plas = [], []
    for _ in range(12):
        o.forces_phase(1e-3)
        offs.append(o.sh3n_state("off")[0].copy()); plas.append(o.sh3n_state("pla").max())
    offs = np.array(offs)
    k = int(np.argmax(offs[:, 0] == 0.0))
    assert k > 0 and np.all(offs[:k] == 1.0) and np.all(offs[k:] == 0.0)
    assert plas[k - 1] <= 1.0e-3 < plas[k]
    assert np.all(o.download_fsky() == 0.0)
