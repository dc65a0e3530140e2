"""Analytic pins for the CPU oracle's brick path (SFORC3 + M2LAW + MQVISCB restatement).

The reference Engine cannot be built here and its QA suite holds no routine-level vectors for
this path, so the restatement is pinned by closed-form patch tests."""
import numpy as np
import pytest
from openradioss_b200 import meshgen
from oracle.orc import Oracle


def block(n=3, jitter=0.05, **kw):
    m = meshgen.hex_block(n, n, n, 1.0 * n, 1.0 * n, 1.0 * n, jitter=jitter, **kw)
    return m


def test_linear_velocity_field_gives_exact_elastic_stress():
    m = block(3)
    L = 1e-6 * np.array([[1.0, 0.3, -0.2], [0.1, -0.5, 0.4], [0.25, -0.15, 0.7]])
    m.V = m.X @ L.T
    o = Oracle(m)
    dt1 = 1e-3
    o.forces_phase(dt1)
    sig = o.solid_state("sig")
    D = 0.5 * (L + L.T); tr = np.trace(D)
    G = m.solid_groups[0].mat.shear
    exp = np.array([2 * G * dt1 * (D[0, 0] - tr / 3), 2 * G * dt1 * (D[1, 1] - tr / 3), 2 * G * dt1 * (D[2, 2] - tr / 3),
                    G * dt1 * 2 * D[0, 1], G * dt1 * 2 * D[1, 2], G * dt1 * 2 * D[0, 2]])
    assert np.allclose(sig, exp[:, None], rtol=1e-9, atol=1e-16)
    assert np.all(o.solid_state("pla") == 0.0)


def test_rigid_rotation_velocity_gives_no_stress_and_no_force():
    # un-jittered: the 4th hourglass vector of SHVIS3 (shvis3.F:363) is not orthogonalised against
    # linear fields, so only a parallelepiped mesh gives exactly zero hourglass force here
    m = block(3, jitter=0.0)
    w = np.array([0.3, -0.2, 0.5]) * 1e-3
    m.V = np.cross(np.tile(w, (m.numnod, 1)), m.X)
    o = Oracle(m)
    o.forces_phase(1e-3)
    assert np.abs(o.solid_state("sig")).max() < 1e-9
    f = o.download_fsky()
    assert np.abs(f[:, :3]).max() < 1e-9


def test_hydrostatic_state_matches_bulk_modulus():
    m = block(2, jitter=0.0)
    s = 0.999
    X0 = m.X.copy()
    o = Oracle(m)
    o.upload_nodes(X=X0 * s)            # compress: V = s^3 V0
    o.forces_phase(0.0)
    mat = m.solid_groups[0].mat
    amu = 1.0 / s ** 3 - 1.0
    sig = o.solid_state("sig")
    assert np.allclose(sig[:3], -mat.bulk * amu, rtol=1e-9)
    assert np.abs(sig[3:]).max() < 1e-6 * mat.bulk * amu
    assert np.allclose(o.solid_state("rho"), mat.rho0 / s ** 3, rtol=1e-12)


@pytest.mark.parametrize("jitter", [0.05, 0.0])
def test_element_forces_are_self_equilibrated(jitter):
    m = block(3, jitter=jitter)
    rng = np.random.default_rng(1)
    m.V = rng.normal(size=m.X.shape) * 1e-2
    o = Oracle(m)
    o.run_cycles(3)
    o.forces_phase(o.time()["dt2"])
    f = o.download_fsky()
    X = o.download_nodes(("X",))["X"]
    for e in range(m.numels):
        rows = f[m.iads[e] - 1, :3]
        scale = np.abs(rows).max()
        assert np.abs(rows.sum(0)).max() <= 1e-12 * scale           # linear momentum
        if jitter:          # the un-orthogonalised 4th hourglass mode carries a small moment on distorted hexes
            continue
        xe = X[m.ixs[e, 1:9] - 1]
        mom = np.cross(xe - xe.mean(0), rows).sum(0)
        assert np.abs(mom).max() <= 1e-5 * scale * np.abs(xe - xe.mean(0)).max()  # angular momentum


def test_time_step_of_unit_cube_at_rest():
    m = meshgen.hex_block(2, 2, 2, 2.0, 2.0, 2.0, jitter=0.0)
    o = Oracle(m)
    o.forces_phase(0.0)
    mat, prop = m.solid_groups[0].mat, m.solid_groups[0].prop
    ssp = np.sqrt((1.333 * mat.shear + mat.bulk) / mat.rho0)
    qx = prop.qb * ssp
    ssp_eq = qx + np.sqrt(qx * qx + ssp * ssp)
    assert o.time()["dt2t"] == pytest.approx(0.9 * 1.0 / ssp_eq, rel=1e-12)
    assert o.time()["ityptst"] == 1 and o.time()["neltst"] == m.numels   # last element at the minimum wins


def test_hourglass_mode_is_resisted_without_strain():
    m = meshgen.hex_block(1, 1, 1, 1.0, 1.0, 1.0, jitter=0.0)
    h = np.array([1, -1, 1, -1, -1, 1, -1, 1.0])       # Flanagan-Belytschko mode 4 (shvis3.F:363)
    m.V = np.zeros_like(m.X); m.V[m.ixs[0, 1:9] - 1, 0] = 1e-3 * h
    o = Oracle(m)
    o.forces_phase(1e-3)
    assert np.abs(o.solid_state("sig")).max() < 1e-12
    f = o.download_fsky()[m.iads[0] - 1, 0]
    assert np.all(np.sign(f) == -np.sign(h)) and np.abs(f).min() > 0


def test_energy_is_conserved_for_an_elastic_free_block():
    """No hourglass / bulk viscosity: leap-frog energy (KE at half steps) oscillates around KE0."""
    m = block(4, jitter=0.0)
    mat = m.solid_groups[0].mat
    mat.ca = 1e30; mat.cc = 0.0; mat.has_temp = 0        # elastic
    p = m.solid_groups[0].prop; p.hcoef = 0.0; p.qa = 0.0; p.qb = 0.0
    m.V = (m.X - m.X.mean(0)) * np.array([1e-1, -5e-2, 2e-2])      # smooth breathing field (mm/ms)
    o = Oracle(m)
    ke0 = 0.5 * (m.MS[:, None] * m.V ** 2).sum()
    tot = []
    for _ in range(40):
        o.run_cycles(13)
        V = o.download_nodes(("V",))["V"]
        ke = 0.5 * (m.MS[:, None] * V ** 2).sum()
        ie = (o.solid_state("eint")[0] * o.solid_state("vol")[0]).sum()
        tot.append((ke + ie) / ke0)
    assert 0.85 < np.mean(tot) < 1.25 and min(tot) > 0.5 and max(tot) < 1.6


def test_openmp_groups_give_identical_results():
    m = block(6)
    m.V = np.random.default_rng(3).normal(size=m.X.shape) * 1e-2
    a = Oracle(m, threads=1); b = Oracle(m, threads=4)
    a.run_cycles(20); b.run_cycles(20)
    assert np.array_equal(a.download_nodes(("X",))["X"], b.download_nodes(("X",))["X"])
    assert a.time()["dt2"] == b.time()["dt2"]


# ---- LAW36 solids: MMAIN -> MULAW -> SIGEPS36 -------------------------------------------

def _law36_block(n=3, **kw):
    return meshgen.hex_block(n, n, n, 1.0 * n, 1.0 * n, 1.0 * n, jitter=kw.pop("jitter", 0.05), law=36, **kw)


def test_law36_elastic_step_matches_hooke():
    """Below yield SIGEPS36 is the deviatoric predictor 2G(de - dav) plus the pressure K*mu."""
    m = _law36_block(3)
    L = 1e-6 * np.array([[1.0, 0.3, -0.2], [0.1, -0.5, 0.4], [0.25, -0.15, 0.7]])
    m.V = m.X @ L.T
    o = Oracle(m)
    dt1 = 1e-3
    o.forces_phase(dt1)
    sig = o.solid_state("sig")
    D = 0.5 * (L + L.T); tr = np.trace(D)
    G = m.solid_groups[0].mat.shear
    exp = np.array([2 * G * dt1 * (D[0, 0] - tr / 3), 2 * G * dt1 * (D[1, 1] - tr / 3), 2 * G * dt1 * (D[2, 2] - tr / 3),
                    G * dt1 * 2 * D[0, 1], G * dt1 * 2 * D[1, 2], G * dt1 * 2 * D[0, 2]])
    assert np.allclose(sig, exp[:, None], rtol=1e-9, atol=1e-16)      # rho == rho0 on the first cycle: no pressure yet
    assert np.all(o.solid_state("pla") == 0.0) and np.all(o.solid_state("wpla") == 0.0)


@pytest.mark.parametrize("ipla", [0, 1, 2])
def test_law36_radial_return_lands_on_the_tabulated_yield_curve(ipla):
    """One large pure-shear increment: the von Mises stress after the return equals the tabulated yield
    stress -- at the old plastic strain for Iplas 0 / 2, updated by H*dpla for Iplas 1 -- and the plastic
    strain increment is the closed-form radial-return value."""
    m = _law36_block(2, jitter=0.0, prop=meshgen.default_prop_solid(ipla=ipla))
    gam = 40.0                                              # shear rate: trial stress G*gam*dt1 ~ 3200 MPa >> 250
    m.V = np.zeros_like(m.X); m.V[:, 0] = gam * m.X[:, 1]
    o = Oracle(m)
    dt1 = 1e-3
    o.forces_phase(dt1)
    mat = m.solid_groups[0].mat
    sig, pla = o.solid_state("sig"), o.solid_state("pla")[0]
    vm = np.sqrt(0.5 * ((sig[0] - sig[1]) ** 2 + (sig[1] - sig[2]) ** 2 + (sig[2] - sig[0]) ** 2) + 3 * (sig[3] ** 2 + sig[4] ** 2 + sig[5] ** 2))
    x, y = m.tf[0::2], m.tf[1::2]
    y0, h0 = y[0], (y[1] - y[0]) / (x[1] - x[0])            # static curve at pla = 0
    trial = np.sqrt(3.0) * mat.shear * gam * dt1
    if ipla == 2:
        dp = (trial - y0) / mat.g3; target = y0
    else:
        dp = (trial - y0) / (mat.g3 + h0); target = y0 if ipla == 0 else y0 + h0 * dp
    assert np.allclose(pla, dp, rtol=1e-12)
    assert np.allclose(vm, target, rtol=1e-12)
    # plastic work = 0.5 (vm_old + vm_new) dpla V  (mulaw.F90:2187-2219)
    assert np.allclose(o.solid_state("wpla")[0], 0.5 * (0.0 + vm) * dp * m.vol0, rtol=1e-12)


def test_law36_rate_dependent_curves_interpolate_in_strain_rate():
    """NRATE = 2: yield = y1 + (epsd - r1)/(r2 - r1) (y2 - y1), with the filtered equivalent deviatoric strain rate."""
    x = np.array([0.0, 0.1, 0.5]); y = np.array([250.0, 350.0, 450.0])
    curves = [(x, y), (x, 1.5 * y)]; rates = [0.0, 100.0]
    m = _law36_block(2, jitter=0.0, curves=curves, rates=rates, prop=meshgen.default_prop_solid(ipla=0))
    gam = 40.0
    m.V = np.zeros_like(m.X); m.V[:, 0] = gam * m.X[:, 1]
    o = Oracle(m)
    dt1 = 1e-3
    o.forces_phase(dt1)
    mat = m.solid_groups[0].mat
    epsdot = gam / np.sqrt(3.0)                             # sqrt(3 * (gam/2)^2) / 1.5
    asrate = min(1.0, mat.asrate * dt1)
    epsd = asrate * epsdot
    assert np.allclose(o.solid_state("epsd")[0], epsd, rtol=1e-13)
    yld = 250.0 + epsd / 100.0 * (375.0 - 250.0)
    sig = o.solid_state("sig")
    vm = np.sqrt(3.0) * np.abs(sig[3])
    assert np.allclose(vm, yld, rtol=1e-12)


def test_law36_energy_is_conserved_for_an_elastic_free_block():
    """No hourglass / bulk viscosity, below yield: leap-frog energy oscillates around KE0 (MULAW energy integration)."""
    m = _law36_block(4, jitter=0.0)
    p = m.solid_groups[0].prop; p.hcoef = 0.0; p.qa = 0.0; p.qb = 0.0
    m.V = (m.X - m.X.mean(0)) * np.array([1e-1, -5e-2, 2e-2])
    o = Oracle(m)
    ke0 = 0.5 * (m.MS[:, None] * m.V ** 2).sum()
    tot = []
    for _ in range(40):
        o.run_cycles(13)
        V = o.download_nodes(("V",))["V"]
        ke = 0.5 * (m.MS[:, None] * V ** 2).sum()
        ie = (o.solid_state("eint")[0] * o.solid_state("vol")[0]).sum()
        tot.append((ke + ie) / ke0)
    assert o.solid_state("pla").max() == 0.0
    assert 0.85 < np.mean(tot) < 1.25 and min(tot) > 0.5 and max(tot) < 1.6


def test_law36_istrain_accumulates_total_strain():
    m = _law36_block(2, jitter=0.0, prop=meshgen.default_prop_solid(istrain=1))
    L = 1e-4 * np.array([[1.0, 0.0, 0.0], [0.0, -0.5, 0.0], [0.0, 0.0, 0.25]])
    m.V = m.X @ L.T
    o = Oracle(m)
    o.forces_phase(1e-3)
    st = o.solid_state("stra")
    assert np.allclose(st[0], 1e-7, rtol=1e-9) and np.allclose(st[1], -0.5e-7, rtol=1e-9) and np.allclose(st[2], 0.25e-7, rtol=1e-9)
    assert np.abs(st[3:]).max() < 1e-20


def test_law36_epsmax_failure_relaxes_the_brick_off():
    """IFAIL = 1 on solids: once PLA > EPSMAX the element's OFF goes 1 -> 0.8 and then x 0.8 per cycle until it falls
    below 0.1 and becomes 0 (sigeps36.F:1507-1510, 1546-1555); stresses are scaled by OFF and SMALLB3 copies OFF to OFFG."""
    mat, npf, tf = meshgen.steel_law36(epsmax=1.0e-3)
    m = meshgen.hex_block(1, 1, 1, 10.0, 10.0, 10.0, law=36, mat=mat, jitter=0.0)
    m.npf, m.tf = npf, tf
    m.V = np.zeros_like(m.X); m.V[:, 0] = 50.0 * m.X[:, 1]            # simple shear, well beyond yield
    o = Oracle(m)
    seq = []
    for _ in range(16):
        o.forces_phase(1e-3)
        seq.append(float(o.solid_state("off")[0, 0]))
    k = next(i for i, v in enumerate(seq) if v < 1.0)
    assert seq[k] == pytest.approx(0.8)
    assert np.allclose(seq[k:k + 10], 0.8 * 0.8 ** np.arange(10), rtol=1e-14)   # 0.8^11 = 0.0859 < 0.1 -> zero next
    assert seq[k + 11] == 0.0 and seq[-1] == 0.0
    assert np.abs(o.solid_state("sig")).max() == 0.0


@pytest.mark.parametrize("fisokin", [0.5, 1.0])
def test_law2_kinematic_hardening_shifts_the_yield_surface(fisokin):
    """M2LAW with FISOKIN > 0 (m2law.F:181-190, 300-337, 364-390): after a plastic step the stress SHIFTED by the back stress
    sits on a yield surface that has grown by the isotropic share (1 - FISOKIN) of the hardening only, the back stress has
    grown along the plastic corrector by ALPHA = HKIN / (2G + HKIN), HKIN = 2/3 FISOKIN QH, and reversing the load yields
    earlier than the isotropic law does (Bauschinger)."""
    def run(fk, reverse):
        m = block(2, jitter=0.0)
        mat = m.solid_groups[0].mat
        mat.cc = 0.0; mat.has_temp = 0; mat.rhocp = 0.0; mat.cn = 1.0; mat.fisokin = fk          # linear hardening: QH = CB
        gam = 2.0 * mat.ca / (np.sqrt(3.0) * mat.shear * 1e-3)                                     # trial von Mises = 2 CA in one step
        m.V = np.zeros_like(m.X); m.V[:, 0] = gam * m.X[:, 1]
        o = Oracle(m)
        o.forces_phase(1e-3)
        out = [(o.solid_state("sig")[:, 0].copy(), o.solid_state("sigb")[:, 0].copy() if fk > 0 else np.zeros(6), o.solid_state("pla").ravel()[0])]
        if reverse:
            o.upload_nodes(V=-1.5 * m.V)
            o.forces_phase(1e-3)
            out.append((o.solid_state("sig")[:, 0].copy(), o.solid_state("sigb")[:, 0].copy() if fk > 0 else np.zeros(6), o.solid_state("pla").ravel()[0]))
        return m.solid_groups[0].mat, out
    vm = lambda s: np.sqrt(0.5 * ((s[0] - s[1]) ** 2 + (s[1] - s[2]) ** 2 + (s[2] - s[0]) ** 2) + 3 * (s[3] ** 2 + s[4] ** 2 + s[5] ** 2))
    mat, [(sig, sb, pla)] = run(fisokin, False)
    trial = 2.0 * mat.ca
    dp = (trial - mat.ca) / (3.0 * mat.shear + mat.cb)
    assert pla == pytest.approx(dp, rel=1e-12)
    ak = mat.ca + (1.0 - fisokin) * mat.cb * dp
    assert vm(sig - sb) == pytest.approx(ak, rel=1e-12)                  # shifted stress on the (partly) grown surface
    hkin = 2.0 / 3.0 * fisokin * mat.cb
    alpha = hkin / (2.0 * mat.shear + hkin)
    assert vm(sb) == pytest.approx(alpha * (trial - ak), rel=1e-10)       # back stress = ALPHA * (predictor - returned), same direction
    assert sb[3] > 0.0 and abs(sb[0]) + abs(sb[1]) + abs(sb[2]) < 1e-9 * sb[3]
    # Bauschinger: the same reversed increment produces more plastic strain with a kinematic part than without
    _, out_k = run(fisokin, True)
    _, out_i = run(0.0, True)
    assert out_k[1][2] - out_k[0][2] > (out_i[1][2] - out_i[0][2]) * (1.0 + 1e-6)


def _rot(axis, ang):
    axis = np.asarray(axis, float) / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)


def _impact_block(jcvt, Q=None):
    m = meshgen.hex_block(3, 3, 5, 1.0, 1.0, 1.66, jitter=0.05, v0=(0, 0, -180.0), vrand=8.0,
                          prop=meshgen.default_prop_solid(jcvt=jcvt))
    m.V[m.X[:, 2] < 0.3] *= 0.0                     # the lower layers at rest: the block compresses and yields
    if Q is not None:
        m.X = m.X @ Q.T; m.V = m.V @ Q.T
    return m


def test_corotational_frame_is_objective():
    """JCVT = 1 (SRCOOR3: Belytschko's co-rotational frame): everything is computed in a frame that turns with the element,
    so a model and the same model turned by an arbitrary rotation Q give corner forces, and after 150 yielding cycles positions,
    related by Q to rounding -- also with plastic strain and hourglass forces in play.  (The Jaumann path of JCVT = 0 is
    objective to the accuracy of the time integration only.)"""
    Q = _rot([0.3, -0.5, 0.8], 1.1)
    a, b = Oracle(_impact_block(1)), Oracle(_impact_block(1, Q))
    for o in (a, b):
        o.forces_phase(0.0)
    fa, fb = a.download_fsky()[:, :3], b.download_fsky()[:, :3]
    assert np.abs(fb - fa @ Q.T).max() <= 1e-12 * np.abs(fa).max()
    ma = _impact_block(1)
    a, b = Oracle(ma), Oracle(_impact_block(1, Q))
    a.run_cycles(150); b.run_cycles(150)
    xa, xb = a.download_nodes(("X",))["X"], b.download_nodes(("X",))["X"]
    assert np.abs(xb - xa @ Q.T).max() <= 1e-9 * np.abs(xa).max()
    assert a.solid_state("pla").max() > 0.01
    assert np.allclose(a.solid_state("sig"), b.solid_state("sig"), rtol=0, atol=1e-8 * np.abs(a.solid_state("sig")).max())   # the stress lives in the element's frame
    # and the two formulations describe the same physics: displacements within a fraction of a per cent of each other
    c = Oracle(_impact_block(0)); c.run_cycles(150)
    xc = c.download_nodes(("X",))["X"]
    assert np.abs(xc - xa).max() <= 5e-3 * np.abs(xa - ma.X).max()


def test_corotational_axis_aligned_cube_without_spin_is_the_global_frame_to_second_order():
    """A cube aligned with the axes under pure stretching has R = 1: the only difference to JCVT = 0 (JHBE = 1) is the
    second-order strain-rate term (sdefo3.F:171-199), -dt/2 (L^T L) -- checked in closed form on the first elastic step."""
    L = 1e-3 * np.diag([1.0, -0.4, 0.7])
    out = []
    for jcvt in (0, 1):
        m = meshgen.hex_block(2, 2, 2, 2.0, 2.0, 2.0, jitter=0.0, prop=meshgen.default_prop_solid(jcvt=jcvt))
        m.V = m.X @ L.T
        o = Oracle(m); o.forces_phase(1e-3)
        out.append((o.solid_state("sig")[:, 0].copy(), m.solid_groups[0].mat.shear))
    (s0, G), (s1, _) = out
    dt = 1e-3
    D0 = np.diag(L); D1 = D0 - 0.5 * dt * D0 ** 2
    dev = lambda D: 2 * G * dt * (D - D.sum() / 3)
    assert np.allclose(s0[:3], dev(D0), rtol=1e-10)
    assert np.allclose(np.sort(s1[:3]), np.sort(dev(D1)), rtol=1e-10)      # in the element's own axes: a permutation of x, y, z here


def test_isolid_101_and_102_run_the_second_order_strain_rate_of_isolid_2():
    """The Engine only compares JHBE with 0, >= 1 and >= 2 on this path (sderi3.F:303, shvis3.F:240/318, sdefo3.F:222; forint.F:1159
    sends 1, 2, 101 and 102 to SFORC3): the old-format values 101 / 102 are Isolid 2 to it."""
    runs = {}
    for jhbe in (2, 101, 102, 1):
        m = meshgen.hex_block(3, 3, 4, 1.0, 1.0, 1.5, v0=(0, 0, -120.0), fix_bottom_z=True, vrand=3.0, prop=meshgen.default_prop_solid(jhbe=jhbe))
        o = Oracle(m); o.run_cycles(30)
        runs[jhbe] = (o.download_nodes(("X",))["X"], o.solid_state("sig"))
    for jhbe in (101, 102):
        assert np.array_equal(runs[jhbe][0], runs[2][0]) and np.array_equal(runs[jhbe][1], runs[2][1])
    assert not np.array_equal(runs[1][1], runs[2][1])
