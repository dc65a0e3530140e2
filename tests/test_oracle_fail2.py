"""LAW36 IFAIL = 2 (tensile-strain damage and failure: EPS_t1, EPS_t2, EPS_f) in the CPU restatement, pinned by the
closed forms of the reference's formulas: shells sigeps36c.F:256-264, 940-950 with the point strains of mulawc.F90:856-862;
solids sigeps36.F:331-395 (largest principal strain by Newton on the deviatoric cubic), 1524-1533."""
import numpy as np
import pytest

from openradioss_b200 import meshgen
from openradioss_b200.constants import K
from oracle.orc import Oracle

CURVE_X = np.array([0.0, 1.0])                       # linear hardening: sigma_y = 250 + 1000 eps_p, so that the explicit
CURVE_Y = np.array([250.0, 1250.0])                  # yield update sigma_y(pla_old) + H dpla is the curve itself
LINEAR = dict(curves=[(CURVE_X, CURVE_Y)], rates=[0.0])


def one_brick(eps_t, L, epsmax=None):
    """unit-ish brick with the velocity field V = L . X (strain increment sym(L) dt per forces phase)"""
    mat, npf, tf = meshgen.steel_law36(eps_t=eps_t, epsmax=epsmax, **LINEAR)
    m = meshgen.hex_block(1, 1, 1, 10.0, 10.0, 10.0, law=36, mat=mat, jitter=0.0, prop=meshgen.default_prop_solid(istrain=1))
    m.npf, m.tf = npf, tf
    m.V = m.X @ np.asarray(L, float).T
    return m


def test_ifail2_is_what_the_starter_sets_and_needs_the_total_strains():
    mat, _, _ = meshgen.steel_law36(eps_t=(0.1, 0.2, 0.15))
    assert mat.ifail == 2 and mat.epsr1 == 0.1 and mat.epsr2 == 0.2 and mat.epsf == 0.15 and mat.epsmax == K["INFINITY"]
    mat, _, _ = meshgen.steel_law36()
    assert mat.ifail == 0 and (mat.epsr1, mat.epsr2, mat.epsf) == (K["INFINITY"], 2.0 * K["INFINITY"], 3.0 * K["INFINITY"])      # hm_read_mat36.F:246-248
    mat, npf, tf = meshgen.steel_law36(eps_t=(0.1, 0.2, 0.15))
    m = meshgen.hex_block(1, 1, 1, 10.0, 10.0, 10.0, law=36, mat=mat, jitter=0.0)            # ISTRAIN = 0
    m.npf, m.tf = npf, tf
    with pytest.raises(RuntimeError):
        Oracle(m)


def newton_eps1(st):
    """sigeps36.F:331-384 as written: upper bound sqrt(-C/3) of the largest deviatoric principal strain, kept when the
    cubic's residual is below the ABSOLUTE threshold 1e-8 (strains under ~3e-3), else 4 Newton steps from 1.75 x bound"""
    exx, eyy, ezz, exy, eyz, ezx = st
    dav = (exx + eyy + ezz) / 3.0
    e1, e2, e3, e4, e5, e6 = exx - dav, eyy - dav, ezz - dav, 0.5 * exy, 0.5 * eyz, 0.5 * ezx
    c = -0.5 * (e1 * e1 + e2 * e2 + e3 * e3) - e4 * e4 - e5 * e5 - e6 * e6
    x = np.sqrt(-c / 3.0)
    d = -e1 * e2 * e3 + e1 * e5 * e5 + e2 * e6 * e6 + e3 * e4 * e4 - 2.0 * e4 * e5 * e6
    if abs((x * x + c) * x + d) > 1e-8:
        x = 1.75 * x
        for _ in range(4):
            x = x - ((x * x + c) * x + d) / (3.0 * x * x + c)
    return x + dav


@pytest.mark.parametrize("scale", [0.1, 1.0e-3])
def test_brick_largest_principal_strain_triggers_the_deletion(scale):
    """EPSTT of sigeps36.F:331-384 against the same recurrence in numpy (and, where the Newton steps run, against the
    largest eigenvalue of the total strain tensor): with EPS_f just below it the brick starts its deletion (OFF = 0.8),
    just above it stays alive.  scale = 1e-3: the residual test is absolute, the bound itself is used."""
    rng = np.random.default_rng(7)
    n = 0
    for _ in range(8):
        L = rng.uniform(-1.0, 1.0, (3, 3))
        L = L - np.eye(3) * min(0.0, np.trace(L) / 3.0 - 0.2)      # keep some volumetric tension
        dt = 1e-3
        e = 0.5 * (L + L.T) * dt * (scale / 1e-3)
        o = Oracle(one_brick((1.0, 2.0, 3.0), L * (scale / 1e-3)))
        o.forces_phase(dt)
        st = o.solid_state("stra")[:, 0]
        assert np.allclose([st[0], st[1], st[2], 0.5 * st[3], 0.5 * st[4], 0.5 * st[5]],
                           [e[0, 0], e[1, 1], e[2, 2], e[0, 1], e[1, 2], e[0, 2]], rtol=1e-9, atol=1e-15)
        eps1 = newton_eps1(st)
        lam = np.linalg.eigvalsh(strain_tensor(st))[-1]
        if eps1 <= 0:
            continue
        n += 1
        if scale == 0.1:
            assert lam * (1.0 - 1e-9) <= eps1 <= lam * 1.1         # Newton from above: an upper bound, a few % at worst
        for fac, alive in ((1.0 - 1e-9, False), (1.0 + 1e-9, True)):
            o = Oracle(one_brick((1.0, 2.0, eps1 * fac), L * (scale / 1e-3)))
            o.forces_phase(dt)
            assert (o.solid_state("off")[0, 0] == 1.0) == alive, (eps1, lam, fac)
    assert n >= 4


def test_brick_damage_factor_scales_the_yield_stress():
    """between EPS_t1 and EPS_t2 the yield stress is the curve times (EPS_t2 - eps1)/(EPS_t2 - EPS_t1): a brick in simple
    shear that flows plastically carries sqrt(3) tau = FAIL * sigma_y(pla); the deletion starts once eps1 > EPS_f and
    relaxes OFF by 0.8 per cycle"""
    e1, e2, ef = 2.0e-2, 6.0e-2, 4.2e-2
    L = np.zeros((3, 3)); L[0, 1] = 10.0                                 # gamma = 1e-2 per cycle, eps1 ~ gamma / 2
    o = Oracle(one_brick((e1, e2, ef), L))
    dt = 1e-3
    offs, eps, seen = [], [], 0
    for k in range(1, 14):
        o.forces_phase(dt)
        off = float(o.solid_state("off")[0, 0]); offs.append(off)
        st = o.solid_state("stra")[:, 0]
        eps1 = newton_eps1(st); eps.append(eps1)
        assert eps1 == pytest.approx(np.linalg.eigvalsh(strain_tensor(st))[-1], rel=1e-6)
        if k <= 6: assert eps1 == pytest.approx(k * 5e-3, rel=2e-2)
        if off == 1.0 and eps1 > e1:
            sig = o.solid_state("sig")[:, 0]
            vm = np.sqrt(0.5 * ((sig[0] - sig[1]) ** 2 + (sig[1] - sig[2]) ** 2 + (sig[2] - sig[0]) ** 2) + 3.0 * (sig[3] ** 2 + sig[4] ** 2 + sig[5] ** 2))
            pla = float(o.solid_state("pla")[0, 0])
            fail = (e2 - eps1) / (e2 - e1)
            assert 0.0 < fail < 1.0 and pla > 0.0
            assert vm == pytest.approx(fail * np.interp(pla, CURVE_X, CURVE_Y), rel=1e-12)
            seen += 1
    assert seen >= 3
    k = next(i for i, v in enumerate(offs) if v < 1.0)
    assert eps[k - 1] <= ef < eps[k]
    assert np.allclose(offs[k:k + 4], 0.8 * 0.8 ** np.arange(4), rtol=1e-14)


def strain_tensor(st):
    return np.array([[st[0], 0.5 * st[3], 0.5 * st[5]], [0.5 * st[3], st[1], 0.5 * st[4]], [0.5 * st[5], 0.5 * st[4], st[2]]])


def test_brick_epsmax_still_acts_under_ifail2():
    """IFAIL = 2 keeps the plastic-strain criterion (sigeps36.F:1524-1533: PLA > EPSMAX .OR. EPSTT > EPSF)"""
    L = np.zeros((3, 3)); L[0, 1] = 50.0
    o = Oracle(one_brick((10.0, 20.0, 30.0), L, epsmax=1.0e-3))
    seq = []
    for _ in range(4):
        o.forces_phase(1e-3); seq.append(float(o.solid_state("off")[0, 0]))
    assert seq[0] == pytest.approx(0.8) and seq[1] == pytest.approx(0.64)


@pytest.mark.parametrize("kind", ["qeph", "bt", "tri"])
def test_shell_tensile_strain_failure(kind):
    """shells: eps1 = (exx + eyy + sqrt((exx - eyy)^2 + exy^2)) / 2 at each point; yield scaled by FAIL while
    EPS_t1 < eps1 < EPS_f, element deleted in the cycle where eps1 > EPS_f"""
    e1, e2, ef = 2.0e-3, 6.0e-3, 4.2e-3
    mat, npf, tf = meshgen.steel_law36(eps_t=(e1, e2, ef), **LINEAR)
    if kind == "tri":
        m = meshgen.tri_plate(1, 1, 10.0, 10.0, mat=mat, jitter=0.0, zjitter=0.0, pressure=0.0, clamp=False)
        state = lambda o, f: o.sh3n_state(f)
    else:
        m = meshgen.shell_plate(1, 1, 10.0, 10.0, mat=mat, prop=meshgen.default_prop_shell(ihbe=24 if kind == "qeph" else 1),
                                jitter=0.0, zjitter=0.0, pressure=0.0, clamp=False)
        state = lambda o, f: o.shell_state(f)
    m.npf, m.tf = npf, tf
    m.V = np.zeros_like(m.X); m.V[:, 0] = 0.5 * m.X[:, 0]             # uniaxial stretching: d(exx) = 5e-4 per phase
    o = Oracle(m)
    offs, seen = [], 0
    for k in range(1, 13):
        o.forces_phase(1e-3)
        off = float(state(o, "off")[0, 0]); offs.append(off)
        st = state(o, "stra")[:, 0]
        if len(offs) < 2 or offs[-2] == 1.0:                             # a deleted element stops straining
            assert st[0] == pytest.approx(k * 5e-4, rel=1e-9) and abs(st[1]) < 1e-15 and abs(st[2]) < 1e-15
        eps1 = st[0]
        if off == 1.0 and eps1 > e1:
            sig = state(o, "sig").reshape(-1, 5, state(o, "off").shape[-1])[:, :, 0] if state(o, "sig").ndim == 2 else state(o, "sig")[:, :, 0]
            pla = state(o, "pla").reshape(-1, state(o, "off").shape[-1])[:, 0]
            fail = (e2 - eps1) / (e2 - e1)
            for ip in range(sig.shape[0]):
                sx, sy, sxy = sig[ip, 0], sig[ip, 1], sig[ip, 2]
                vm = np.sqrt(sx * sx + sy * sy - sx * sy + 3.0 * sxy * sxy)
                assert vm == pytest.approx(fail * np.interp(pla[ip], CURVE_X, CURVE_Y), rel=2e-3)
            seen += 1
    assert seen >= 3
    k = next(i for i, v in enumerate(offs) if v < 1.0)
    assert k == 8 and all(v == 0.0 for v in offs[k:])                    # same-cycle deletion through MULAWC
    assert np.all(o.download_fsky() == 0.0)
