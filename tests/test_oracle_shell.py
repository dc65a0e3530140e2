"""Analytic pins for the CPU oracle's shell paths (CZFORC3 / CFORC3 + CMAIN3/MULAWC + SIGEPS36C /
SIGEPS02C restatements).  The reference Engine cannot be built here and its QA suite holds no
routine-level vectors for this path: the restatement is pinned by closed-form
known answers and physical invariants, for both element formulations independently."""
import numpy as np
import pytest
from openradioss_b200 import meshgen
from oracle.orc import Oracle

FAMILIES = [24, 1, 3, 4]     # QEPH, BT Ishell 1 / 3 / 4


def plate(ihbe, nx=4, ny=3, jitter=0.05, zjitter=0.0, law=36, **kw):
    prop = meshgen.default_prop_shell(ihbe=ihbe, npt=5)
    return meshgen.shell_plate(nx, ny, 10.0 * nx, 10.0 * ny, prop=prop, law=law, jitter=jitter, zjitter=zjitter,
                               pressure=0.0, clamp=False, **kw)


@pytest.mark.parametrize("ihbe", FAMILIES)
def test_uniform_membrane_strain_rate_gives_plane_stress_elastic_stress(ihbe):
    m = plate(ihbe)
    for g in m.shell_groups:
        g.prop.dm = 0.0                                        # no membrane damping: the resultant is the stress alone
    L = 1e-6 * np.array([[1.0, 0.4], [-0.1, -0.6]])            # in-plane velocity gradient
    m.V[:, :2] = m.X[:, :2] @ L.T
    o = Oracle(m)
    dt1 = 1e-3
    o.forces_phase(dt1)
    mat = m.shell_groups[0].mat
    exx, eyy, gxy = L[0, 0] * dt1, L[1, 1] * dt1, (L[0, 1] + L[1, 0]) * dt1
    a11 = mat.young / (1 - mat.nu ** 2)
    exp = np.array([a11 * (exx + mat.nu * eyy), a11 * (eyy + mat.nu * exx), mat.shear * gxy])
    sig = o.shell_state("sig").reshape(5, 5, -1)                  # (ipt, comp, elem)
    # the stress tensor lives in each element's local frame: compare invariants
    tr = sig[:, 0] + sig[:, 1]; det = sig[:, 0] * sig[:, 1] - sig[:, 2] ** 2
    assert np.allclose(tr, exp[0] + exp[1], rtol=2e-6)
    assert np.allclose(det, exp[0] * exp[1] - exp[2] ** 2, rtol=2e-5)
    assert np.abs(sig[:, 3:]).max() < 1e-9 * abs(exp).max()      # no transverse shear
    # membrane force resultant = sum WF * sigma = sigma (weights sum to one), no moment
    forc = o.shell_state("forc"); mom = o.shell_state("mom")
    assert np.allclose(forc[0] + forc[1], exp[0] + exp[1], rtol=2e-6)
    assert np.abs(mom).max() < 1e-7 * abs(exp).max()
    assert np.all(o.shell_state("pla") == 0.0)


@pytest.mark.parametrize("ihbe", FAMILIES)
def test_rigid_body_motion_gives_no_strain_and_no_force(ihbe):
    m = plate(ihbe, jitter=0.0)
    w = np.array([0.2, -0.3, 0.4]) * 1e-4
    v0 = np.array([3.0, -1.0, 2.0])
    m.V = v0 + np.cross(np.tile(w, (m.numnod, 1)), m.X)
    m.VR = np.tile(w, (m.numnod, 1))
    o = Oracle(m)
    o.forces_phase(1e-3)
    mat = m.shell_groups[0].mat
    scale = mat.young * 1e-4 * 1e-3                              # stress a strain of |w| dt would give
    assert np.abs(o.shell_state("sig")).max() < 1e-6 * scale
    f = o.download_fsky()
    assert np.abs(f[:, :6]).max() < 1e-5 * scale * 10.0 * 2.0


@pytest.mark.parametrize("ihbe", FAMILIES)
def test_pure_bending_rate_gives_plate_moment(ihbe):
    """rotation-rate field theta_y = -k x (cylindrical bending about y): kxx = k dt, M = D * kxx with the
    5-point Lobatto-like rule of /PROP/SHELL (sum WM*z = 1/12 only approximately: use the rule itself)."""
    m = plate(ihbe, jitter=0.0)
    k = 1e-6
    m.VR[:, 1] = k * m.X[:, 0]                                   # d(theta_y)/dx = k  ->  kxx
    m.V[:, 2] = -0.5 * k * m.X[:, 0] ** 2                        # w consistent with theta: no transverse shear
    o = Oracle(m)
    dt1 = 1e-3
    o.forces_phase(dt1)
    mat, prop = m.shell_groups[0].mat, m.shell_groups[0].prop
    a11 = mat.young / (1 - mat.nu ** 2)
    z0 = np.array([-.5, -.25, 0, .25, .5]); wm = np.array([-.0520833, -.0625, 0, .0625, .0520833], np.float32).astype(float)
    kap = k * dt1
    mom = o.shell_state("mom")
    exp = (wm * z0).sum() * prop.thick * a11 * kap               # MOM = sum WM * sigma(z), sigma = a11 * z*t*kappa
    assert np.allclose(np.abs(mom[0]), abs(exp), rtol=1e-3)
    assert np.allclose(np.abs(mom[1]), abs(mat.nu * exp), rtol=1e-3)
    assert abs((wm * z0).sum() - 1.0 / 12.0) < 2e-3              # and the rule approximates t^3/12


@pytest.mark.parametrize("ihbe", [24, 1])
@pytest.mark.parametrize("ipla", [0, 1, 2])
def test_law36_return_lands_on_the_tabulated_yield_curve(ihbe, ipla):
    prop = meshgen.default_prop_shell(ihbe=ihbe, npt=3, ipla=ipla)
    m = meshgen.shell_plate(2, 2, 20.0, 20.0, prop=prop, jitter=0.0, zjitter=0.0, pressure=0.0, clamp=False)
    rate = 2e-2                                                  # uniaxial stretch, 2 % per ms
    m.V[:, 0] = rate * m.X[:, 0]
    o = Oracle(m)
    o.forces_phase(0.0)
    for _ in range(100):
        pla_prev = o.shell_state("pla")
        o.forces_phase(0.02)                                     # 0.04 % strain per call, nodes not advanced
    sig = o.shell_state("sig").reshape(3, 5, -1); pla = o.shell_state("pla")
    assert pla_prev.min() > 0.02 and pla.max() < 0.05            # both ends of the last step on one curve segment
    svm = np.sqrt(sig[:, 0] ** 2 + sig[:, 1] ** 2 - sig[:, 0] * sig[:, 1] + 3 * sig[:, 2] ** 2)
    x = m.tf[0::2]; y = m.tf[1::2]
    # Iplas=0 (radial return) lands on the yield stress at the plastic strain the step started from;
    # Iplas=1 (3 Newton iterations) and Iplas=2 (normal projection + hardening update) on the updated one
    target = np.interp(pla_prev if ipla == 0 else pla, x, y)
    # Iplas=1 keeps the iterate *before* the third Newton update (sigeps36c.F:548-573 uses DPLA_I, not
    # DPLA_J), so the consistency residual is second-order small in the step, not round-off
    assert np.allclose(svm, target, rtol=1e-4 if ipla == 1 else 1e-9)


def test_law2_shell_yield_is_johnson_cook():
    prop = meshgen.default_prop_shell(ihbe=24, npt=3, ipla=1)
    m = meshgen.shell_plate(2, 2, 20.0, 20.0, prop=prop, law=2, jitter=0.0, zjitter=0.0, pressure=0.0, clamp=False)
    mat = m.shell_groups[0].mat
    mat.cc = 0.0                                                 # no rate term: sigma_y = A + B eps^n
    m.V[:, 0] = 2e-2 * m.X[:, 0]
    o = Oracle(m)
    for _ in range(100):
        o.forces_phase(0.02)
    sig = o.shell_state("sig").reshape(3, 5, -1); pla = o.shell_state("pla")
    svm = np.sqrt(sig[:, 0] ** 2 + sig[:, 1] ** 2 - sig[:, 0] * sig[:, 1] + 3 * sig[:, 2] ** 2)
    assert pla.min() > 0.02
    assert np.allclose(svm, mat.ca + mat.cb * pla ** mat.cn, rtol=1e-4)


@pytest.mark.parametrize("ihbe", FAMILIES)
def test_element_time_step_scales_with_size_and_wave_speed(ihbe):
    dts = []
    for h in (5.0, 10.0):
        prop = meshgen.default_prop_shell(ihbe=ihbe, npt=5)
        m = meshgen.shell_plate(3, 3, 3 * h, 3 * h, prop=prop, jitter=0.0, zjitter=0.0, pressure=0.0, clamp=False)
        o = Oracle(m); o.forces_phase(0.0)
        dts.append(o.time()["dt2t"])
    assert dts[1] == pytest.approx(2 * dts[0], rel=1e-12)
    mat = m.shell_groups[0].mat
    c = np.sqrt(mat.young / (1 - mat.nu ** 2) / mat.rho0)
    assert 0.5 * 10.0 / c < dts[1] < 10.0 / c                    # 0.9 * (reduced characteristic length) / c


@pytest.mark.parametrize("ihbe", [24, 1])
def test_energy_balance_under_pressure(ihbe):
    prop = meshgen.default_prop_shell(ihbe=ihbe, npt=5)
    m = meshgen.shell_plate(8, 8, 80.0, 80.0, prop=prop, pressure=5.0)
    o = Oracle(m)
    o.run_cycles(400)
    d = o.download_nodes(("D", "V", "VR"))
    wext = (m.fext * d["D"]).sum()
    ke = 0.5 * (m.MS[:, None] * d["V"] ** 2).sum() + 0.5 * (m.IN[:, None] * d["VR"] ** 2).sum()
    ie = o.shell_state("eint").sum()
    assert o.shell_state("pla").max() > 0.01
    assert abs(ie + ke - wext) < 0.01 * wext                     # remainder: hourglass damping / BT hourglass energy


def test_vinter_cursor_matches_numpy_interp():
    prop = meshgen.default_prop_shell(ihbe=24, npt=1, ipla=0)
    x = np.array([0.0, 0.02, 0.05, 0.2, 0.6]); y = np.array([100.0, 180.0, 210.0, 260.0, 300.0])
    m = meshgen.shell_plate(1, 1, 10.0, 10.0, prop=prop, jitter=0.0, zjitter=0.0, pressure=0.0, clamp=False,
                            curves=[(x, y)], rates=[0.0])
    m.V[:, 0] = 0.05 * m.X[:, 0]
    o = Oracle(m)
    for _ in range(12):
        o.forces_phase(1.0)                                      # 5 % per call: walks across all segments
        sig = o.shell_state("sig")[:, 0]; pla = o.shell_state("pla")[0, 0]
        svm = np.sqrt(sig[0] ** 2 + sig[1] ** 2 - sig[0] * sig[1] + 3 * sig[2] ** 2)
        # radial return (Iplas=0): stress sits on the yield value evaluated at the plastic strain of the step start
        assert svm <= np.interp(pla, x, y) * (1 + 1e-12) + 1e-9
    assert pla > 0.3


# ---- kinematic / mixed hardening of LAW36 (FISOKIN > 0, sigeps36c.F:272-274, 329-337, 986-1002) ---------------------------------
def _stretch_run(fisokin, ncyc=400, reverse_at=None):
    """One flat QEPH element strip stretched in x by prescribed nodal velocities (optionally reversed): returns the history of
    (sigma_xx, back stress xx, plastic strain) of the mid-surface point of element 0."""
    m = meshgen.shell_plate(2, 1, 20.0, 10.0, pressure=0.0, clamp=False, jitter=0.0, zjitter=0.0)
    for g in m.shell_groups:
        g.mat.fisokin = fisokin; g.prop.dm = 0.0
    o = Oracle(m)
    rate = 2.0                                   # 1/ms
    hist = []
    for c in range(ncyc):
        sgn = -1.0 if (reverse_at is not None and c >= reverse_at) else 1.0
        V = np.zeros_like(m.X); V[:, 0] = sgn * rate * m.X[:, 0]; V[:, 1] = -0.5 * sgn * rate * m.X[:, 1] * 0.0
        o.upload_nodes(X=m.X, V=V, VR=np.zeros_like(V))
        o.forces_phase(1.0e-5)
        sig = o.shell_state("sig").reshape(5, 5, -1); sb = o.shell_state("sigb").reshape(5, 3, -1)
        hist.append((sig[2, 0, 0], sig[2, 1, 0], sb[2, 0, 0], o.shell_state("pla")[2, 0]))
    return np.array(hist)


def test_kinematic_hardening_matches_isotropic_under_monotonic_proportional_loading():
    """Radial monotonic loading cannot tell the two apart: |sigma - back stress| stays on the initial yield surface while the
    back stress carries H*eps_p, so the total stress follows the same curve."""
    iso, kin, mix = _stretch_run(0.0), _stretch_run(1.0), _stretch_run(0.5)
    assert iso[-1, 3] > 0.005 and np.all(iso[:, 2] == 0.0) and kin[-1, 2] > 0.0
    for h in (kin, mix):
        assert np.allclose(h[-1, :2], iso[-1, :2], rtol=1e-2)
        assert np.isclose(h[-1, 3], iso[-1, 3], rtol=5e-3)


def test_kinematic_hardening_shows_the_bauschinger_effect():
    """Reverse the stretch after hardening: with isotropic hardening reverse yield waits for -sigma_max; with kinematic
    hardening it comes 2*sigma_y0 below the forward stress, i.e. much earlier -- more plastic strain for the same reversal."""
    iso = _stretch_run(0.0, ncyc=700, reverse_at=400); kin = _stretch_run(1.0, ncyc=700, reverse_at=400)
    assert np.isclose(iso[399, 3], kin[399, 3], rtol=5e-3)            # same state when the load turns
    d_iso, d_kin = iso[-1, 3] - iso[399, 3], kin[-1, 3] - kin[399, 3]
    assert d_kin > d_iso + 2e-4            # ~ (sigma_max - sigma_y0) * 2 / E' of extra reverse flow
    assert kin[-1, 2] < kin[399, 2]                                    # the back stress follows the reversed flow


# ---- LAW36 with VP = 1: curves interpolated on the filtered PLASTIC strain rate (sigeps36c.F:665-923, 976-982) ---------------------
def _vp1_run(vp, ncyc=500, rate=2.0, dt=1.0e-5):
    x = np.array([0.0, 0.01, 0.03, 0.08, 0.2, 0.5]); y = np.array([250.0, 300.0, 340.0, 390.0, 440.0, 480.0])
    m = meshgen.shell_plate(2, 1, 20.0, 10.0, pressure=0.0, clamp=False, jitter=0.0, zjitter=0.0,
                            curves=[(x, y), (x, 1.15 * y), (x, 1.4 * y)], rates=[0.0, 0.5, 50.0])
    for g in m.shell_groups:
        g.mat.vp = vp; g.prop.dm = 0.0
    o = Oracle(m)
    hist = []
    for c in range(ncyc):
        V = np.zeros_like(m.X); V[:, 0] = rate * m.X[:, 0]
        o.upload_nodes(X=m.X, V=V, VR=np.zeros_like(V))
        o.forces_phase(dt)
        sig = o.shell_state("sig").reshape(5, 5, -1)
        hist.append((sig[2, 0, 0], sig[2, 1, 0], o.shell_state("pla")[2, 0], o.shell_state("plap")[2, 0], o.shell_state("epsd_ip")[2, 0]))
    return np.array(hist), m.shell_groups[0].mat, (x, y), dt


def test_law36_vp1_filters_the_plastic_strain_rate_and_leaves_the_total_one_alone():
    h, mat, _, dt = _vp1_run(1)
    a = min(1.0, mat.asrate * dt)
    assert h[-1, 2] > 0.003 and np.all(h[:, 4] == 0.0)               # plastic; LBUF%EPSD untouched
    dpla = np.diff(np.concatenate([[0.0], h[:, 2]]))
    plap = 0.0
    for c in range(len(h)):                                            # UVAR(2) <- a * DPLA / dt + (1 - a) * UVAR(2)
        plap = a * dpla[c] / dt + (1.0 - a) * plap
        assert np.isclose(h[c, 3], plap, rtol=1e-9, atol=1e-12), c
    assert 1.5 < h[-1, 3] < 2.0 * 2.0 / np.sqrt(3.0)                   # steady flow with eps_yy = 0: just below 2/sqrt(3) times the stretch rate of 2 / ms


def test_law36_vp1_stress_sits_on_the_curve_interpolated_at_the_plastic_rate():
    h, mat, (x, y), dt = _vp1_run(1)
    sxx, syy, pla, plap = h[-1, 0], h[-1, 1], h[-1, 2], h[-2, 3]       # the yield stress of the last cycle used the rate of the one before
    svm = np.sqrt(sxx * sxx + syy * syy - sxx * syy)
    rfac = (plap - 0.5) / (50.0 - 0.5)
    y0 = np.interp(pla, x, y)
    assert np.isclose(svm, y0 * (1.15 + rfac * (1.4 - 1.15)), rtol=2e-4)
    # the total-rate form (VP = 0) reads a higher rate during the elastic rise and the same one in steady flow
    h0, *_ = _vp1_run(0)
    assert np.isclose(h0[-1, 0], h[-1, 0], rtol=5e-3) and h0[-1, 3] == 0.0 and h0[-1, 4] > 0.0


def test_law36_vp1_needs_rate_curves():
    m = meshgen.shell_plate(2, 1, 20.0, 10.0)
    for g in m.shell_groups:
        g.mat.vp = 1
    with pytest.raises(Exception):
        Oracle(m)
