"""The oracle against numbers the REFERENCE ENGINE ITSELF produced: the per-cycle listings the reference's QA suite keeps as the
expected runs of its decks (qa-tests/miniqa/*/reference/*_0001.out -> tests/golden/qa_*.npz by tests/golden/make_golden_qa.py;
the decks restated by hand in tests/qa_decks.py).  The listing prints TIME, TIME-STEP, I-ENERGY, K-ENERGY T / R and EXT-WORK with
four significant digits, so the bound is the print precision (6e-4 relative, cf. qa-tests/scripts/or_QA.constants) on EVERY
printed cycle -- thousands of consecutive cycles through yield, rate dependence and bending.  What each deck pins:
  ELEM_SAMP      QEPH shells (CZFORC3 chain), LAW36 with 5 rate curves + strain-rate filter (SIGEPS36C, Iplas 1), membrane
                 viscosity default of the Starter, thickness update, FIXVEL + its external work, BCS, CNDT3, ASSPAR4 / ACCELE /
                 VELOCITY / DEPLA, CBILAN + ECRIT
  CT3AV4         3-node shells (C3FORC3), LAW2 (SIGEPS02C / M2CPLR, Iplas 0, NPT 5) in plastic bending, rotational dofs and
                 their kinetic energy, C3DT3 and the controlling element id
  inibri_stress  8-node brick (SFORC3 chain), LAW36 solid (MULAW -> SIGEPS36), initial stress, nodal time step (DTNODA)
"""
import os
import numpy as np
import pytest
import qa_decks
from oracle.orc import Oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PRINT_TOL = 6e-4          # half a unit of the 4th significant digit is 5e-4 at worst


def listing(name):
    return np.load(os.path.join(GOLD, f"qa_{name}.npz"))


def run_listing(b, ncycles):
    """Rows (TIME, DT, ENINT, ENCIN, ENROT, WFEXT, NELTST) as the Engine prints them for cycles 0 .. ncycles-1."""
    b.set_print(True)
    rows = []
    for _ in range(ncycles):
        b.run_cycles(1)
        t, e = b.time(), b.balance()
        rows.append((t["tt"] - t["dt2"], t["dt2"], e["enint"], e["encin"], e["enrot"], e["wfext"], t["neltst"]))
    return np.array(rows)


def close(a, b, floor):
    """|a - b| <= PRINT_TOL * max(|b|, floor): `floor` keeps the first cycles, printed as 0.xxxxE-0y, on the same footing."""
    return np.abs(a - b) <= PRINT_TOL * np.maximum(np.abs(b), floor)


def check_elem_samp(b, ncycles=2045):
    g = listing("elem_samp")
    r = run_listing(b, ncycles)
    n = ncycles
    assert close(r[:, 0], g["time"][:n], 1e-2).all()
    assert close(r[:, 1], g["dt"][:n], 0.0).all()
    assert close(r[:, 2], g["ienergy"][:n], 1e-4).all()
    assert close(r[:, 3], g["kenergy_t"][:n], 1e-5).all()
    assert close(r[:, 5], g["extwork"][:n], 1e-4).all()
    assert (r[:, 6] == g["elid"][:n]).all()                      # element 1 controls the time step until it fails
    assert g["ienergy"][n - 1] > 18.0                            # deep in the plastic range when /FAIL/TAB acts
    return r


def check_ct3a(b, ncycles=3351):
    g = listing("ct3a")
    r = run_listing(b, ncycles)
    k = g["cycle"][g["cycle"] < ncycles]
    gi = np.arange(len(k))
    assert close(r[k, 0], g["time"][gi], 1.0).all()
    assert close(r[k, 1], g["dt"][gi], 0.0).all()
    assert close(r[k[1:], 2], g["ienergy"][gi[1:]], 1e-6).all()
    assert close(r[k, 3], g["kenergy_t"][gi], 1e-6).all()
    assert close(r[k[1:], 4], g["kenergy_r"][gi[1:]], 1e-7).all()
    assert (r[k[1:], 6] == g["elid"][gi[1:]]).all()              # the controlling element, cycle 0 aside (a tie the 4 domains of the reference run break their way)
    return r


def check_inibri(b):
    g = listing("inibri_stress")
    b.upload_solid_state("sig", b.model.initial_solid_sig)
    r = run_listing(b, 3)
    assert close(r[:, 1], g["dt"][:3], 0.0).all()                # nodal time step of cycles 0, 1, 2
    assert close(r[1:2, 2], g["ienergy"][1:2], 0.0).all()        # -37.84: the stress relaxing from the initial 900
    assert close(r[1:3, 3], g["kenergy_t"][1:3], 0.0).all()      # 24.05, 11.52
    # from cycle 2 on the run is a blow-up (a free brick at yield, then the deck's TYPE11 self-contact): differences grow 3x per
    # cycle there, so only the order of magnitude is asked of cycle 2's internal energy
    assert abs(r[2, 2] - g["ienergy"][2]) <= 0.01 * abs(g["ienergy"][2])
    return r


def test_elem_samp_oracle_reproduces_the_reference_listing():
    check_elem_samp(Oracle(qa_decks.elem_samp()))


def test_ct3a_oracle_reproduces_the_reference_listing():
    check_ct3a(Oracle(qa_decks.ct3a()))


def test_inibri_stress_oracle_reproduces_the_reference_listing():
    check_inibri(Oracle(qa_decks.inibri_stress()))


def test_loi13_solide_first_time_step_matches_the_reference_listing():
    """464 LAW2 bricks in the co-rotational frame (Iframe = 2): the element time step of cycle 0 -- SDLEN3's characteristic length,
    the sound speed of M2LAW, DTFAC -- to the listing's four digits.  The later cycles are not comparable: the deck's LAW13
    inclusions are rigid bodies in the reference (qa_decks.loi13_solide), and the listing's energies show it: they run 16-19 %
    above a model with voids in their place, growing with the compression."""
    g = listing("loi13_solide")
    o = Oracle(qa_decks.loi13_solide())
    r = run_listing(o, 101)
    assert close(r[0:1, 1], g["dt"][0:1], 0.0).all() and r[0, 2] == 0.0
    k = int(np.nonzero(g["cycle"] == 100)[0][0])
    assert 0.75 * g["ienergy"][k] < r[100, 2] < g["ienergy"][k]            # softer than the reference, same order
    assert abs(r[100, 0] - g["time"][k]) <= 1e-3 * g["time"][k]


def test_membrane_damping_default_is_the_starters():
    """Without the Starter's 1.5 % membrane damping for QEPH (set_elgroup_param.F:83-108) the internal energy of ELEM_SAMP is
    4 % short at cycle 1 and 0.1 % short throughout: the listing discriminates that term."""
    m = qa_decks.elem_samp()
    m.shell_groups[0].prop.dm = 0.0
    o = Oracle(m); r = run_listing(o, 3)
    g = listing("elem_samp")
    assert not close(r[1:3, 2], g["ienergy"][1:3], 1e-4).any()
