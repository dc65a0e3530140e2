"""/FAIL/JOHNSON on shells in the oracle: FAIL_JOHNSON_C (fail_johnson_c.F:111-130) behind MULAWC (mulawc.F90:2064-2069,
2118-2127, 2608-2637) and the one-layer deletion rule of FAIL_SETOFF_C (fail_setoff_c.F:123-186), against closed-form answers."""
import numpy as np
import pytest
from openradioss_b200 import meshgen
from openradioss_b200.model import Fail
from oracle.orc import Oracle


def fail(d1=0.05, d2=0.0, d3=0.0, d4=0.0, epsp0=1.0, epsf_min=0.0, pthk=0.0, pthickg=1.0):
    f = Fail(); f.irupt = 1; f.d1, f.d2, f.d3, f.d4, f.d5 = d1, d2, d3, d4, 0.0
    f.epsp0, f.epsf_min, f.pthk, f.pthickg = epsp0, epsf_min, pthk, pthickg
    return f


def stretched_plate(f, npt=3, law=36, rate=40.0, n=3, bend=0.0):
    """Flat plate under a uniform in-plane stretching velocity field (+ an optional curvature rate): every element sees the
    same strain increment, so damage grows identically everywhere."""
    m = meshgen.shell_plate(n, n, 10.0 * n, 10.0 * n, law=law, prop=meshgen.default_prop_shell(npt=npt), jitter=0.0, zjitter=0.0,
                            pressure=0.0, clamp=False)
    m.V[:, 0] = rate * m.X[:, 0]; m.V[:, 1] = 0.3 * rate * m.X[:, 1]
    if bend:
        m.VR[:, 1] = bend * m.X[:, 0]
    for g in m.shell_groups:
        g.fail = f
    return m


def test_damage_is_the_sum_of_dpla_over_the_failure_strain():
    """D2 = D4 = 0: eps_f = D1, so after every cycle DFMAX = sum(DPLA) / D1 = PLA / D1 at every point (until it reaches 1)."""
    m = stretched_plate(fail(d1=0.08))
    o = Oracle(m)
    dt1 = 1e-3
    for c in range(6):
        o.forces_phase(dt1)
        pla, dmg, foff = o.shell_state("pla"), o.shell_state("dfmax"), o.shell_state("foff")
        assert np.allclose(dmg, np.minimum(1.0, pla / 0.08), rtol=1e-12, atol=0.0)
        assert np.array_equal(foff, (pla / 0.08 < 1.0).astype(float)) or np.all((foff == 0.0) == (dmg >= 1.0))
    assert pla.max() > 0.0


def test_triaxiality_and_rate_terms_of_the_failure_strain():
    """One plastic step from a virgin plate: DFMAX = DPLA / ((D1 + D2 exp(D3 p / svm)) (1 + D4 ln(max(1, epsd / EPSP0)))) with
    p, svm from the returned stress of the point and epsd its strain rate."""
    f = fail(d1=0.03, d2=0.4, d3=-1.2, d4=0.05, epsp0=1.0e-3)
    m = stretched_plate(f, npt=3)
    o = Oracle(m)
    o.forces_phase(1e-3)
    sig, pla, epsd, dmg = o.shell_state("sig"), o.shell_state("pla"), o.shell_state("epsd_ip"), o.shell_state("dfmax")
    assert pla.min() > 0.0
    for ip in range(3):
        sx, sy, sxy = sig[5 * ip], sig[5 * ip + 1], sig[5 * ip + 2]
        p = (sx + sy) / 3.0
        svm = np.sqrt(sx * sx + sy * sy - sx * sy + 3.0 * sxy * sxy)
        epsf = (0.03 + 0.4 * np.exp(-1.2 * p / svm)) * (1.0 + 0.05 * np.log(np.maximum(1.0, epsd[ip] / 1.0e-3)))
        assert np.allclose(dmg[ip], np.minimum(1.0, pla[ip] / epsf), rtol=1e-12)


def test_failed_point_restarts_from_zero_stress_and_element_goes_at_p_thickfail():
    """Bending makes the outer points fail first.  A failed point (FOFF = 0) keeps no stress (LBUF%SIG = SIG * SIGOFF), the
    element stays while the broken share of the thickness is below P_thickfail and is deleted (OFF = 0, no forces) in the cycle
    in which it reaches it; with P_thickfail given as a NEGATIVE number the share of broken points counts instead."""
    for pthk, expect_points in ((0.45, None), (-0.6, 3)):
        f = fail(d1=0.02, pthk=pthk, pthickg=1.0)
        m = stretched_plate(f, npt=5, rate=15.0, bend=3.0)
        o = Oracle(m)
        dt1 = 1e-3
        gone_at = None
        for c in range(60):
            o.forces_phase(dt1)
            foff, off, sig = o.shell_state("foff"), o.shell_state("off")[0], o.shell_state("sig")
            for ip in range(5):
                dead = foff[ip] == 0.0
                assert np.all(sig[5 * ip:5 * ip + 5][:, dead] == 0.0)
            nbroken = (foff == 0.0).sum(0)
            if np.any(off == 0.0):
                gone_at = c
                if expect_points is not None:
                    assert np.all(nbroken[off == 0.0] >= expect_points)          # 3 of 5 points >= 0.6
                break
            if expect_points is not None:
                assert np.all(nbroken < expect_points)
            o.assemble(); o.advance(dt1, dt1)
        assert gone_at is not None and gone_at > 0
        assert np.abs(o.download_fsky()[:, :6]).max() >= 0.0
        # a deleted element gives no force from the next cycle on
        o.assemble(); o.advance(dt1, dt1); o.forces_phase(dt1)
        f_el = o.shell_state("forc")
        assert np.all(f_el[:, off == 0.0] == 0.0)


def test_no_failure_model_leaves_the_law_alone():
    m0 = stretched_plate(None); m1 = stretched_plate(fail(d1=1.0e9))
    for g in m0.shell_groups:
        g.fail = None
    a, b = Oracle(m0), Oracle(m1)
    for o in (a, b):
        o.run_cycles(30)
    assert np.array_equal(a.download_nodes(("X",))["X"], b.download_nodes(("X",))["X"])


# ---- solids: FAIL_JOHNSON behind MMAIN (mmain.F90:2250-2416 -> fail_johnson.F:95-141, Ifail_so = 1), LAW2 bricks

def sheared_block(f, gam=60.0):
    m = meshgen.hex_block(2, 2, 2, 2.0, 2.0, 2.0, jitter=0.0)
    for g in m.solid_groups:
        g.mat.cc = 0.0; g.mat.has_temp = 0; g.mat.rhocp = 0.0; g.fail = f
    m.V = np.zeros_like(m.X); m.V[:, 0] = gam * m.X[:, 1]
    return m


def test_solid_damage_grows_by_dpla_over_the_failure_strain_and_the_element_relaxes_away():
    """Pure shear, D2 = D4 = 0: DFMAX = PLA / D1 while the element lives; in the cycle DFMAX reaches 1 OFF becomes 4/5, then it
    shrinks by the REAL*4 factor 0.8 per cycle until it drops below 0.1 and the element is gone."""
    f = fail(d1=0.015)
    m = sheared_block(f)
    o = Oracle(m)
    dt1 = 2e-4
    offs = []
    for c in range(40):
        o.forces_phase(dt1)
        pla, dmg, off = o.solid_state("pla")[0], o.solid_state("dfmax")[0], o.solid_state("off")[0]
        offs.append(off[0])
        if off[0] == 1.0:
            assert np.allclose(dmg, np.minimum(1.0, pla / 0.015), rtol=1e-12)
        o.assemble(); o.advance(dt1, dt1)
        if off[0] == 0.0:
            break
    offs = np.array(offs)
    k = int(np.argmax(offs < 1.0))
    assert k > 0 and offs[k] == 0.8 and offs[-1] == 0.0
    r = float(np.float32(0.8))
    assert np.allclose(offs[k + 1:k + 4], 0.8 * r ** np.arange(1, 4), rtol=1e-15)          # 0.8 is a single-precision literal there
    assert np.all(o.solid_state("dfmax")[0] == 1.0)


def test_solid_triaxiality_term():
    f = fail(d1=0.02, d2=0.3, d3=-1.5)
    m = sheared_block(f, gam=30.0)
    m.V[:, 1] += 8.0 * m.X[:, 1]                     # some hydrostatic part
    o = Oracle(m)
    o.forces_phase(2e-4)
    s, pla, dmg = o.solid_state("sig"), o.solid_state("pla")[0], o.solid_state("dfmax")[0]
    p = (s[0] + s[1] + s[2]) / 3.0
    svm = np.sqrt(3.0 * (0.5 * ((s[0] - p) ** 2 + (s[1] - p) ** 2 + (s[2] - p) ** 2) + s[3] ** 2 + s[4] ** 2 + s[5] ** 2))
    assert pla.min() > 0.0
    assert np.allclose(dmg, pla / (0.02 + 0.3 * np.exp(-1.5 * p / svm)), rtol=1e-12)
