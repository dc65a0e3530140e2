"""Known answers for the oracle's print-cycle balances (CBILAN / C3BILAN / SBILAN -> PARTSAV, ECRIT; oracle/bilan.cpp)."""
import numpy as np
from openradioss_b200 import meshgen
from oracle.orc import Oracle


def two_part_model():
    m = meshgen.shell_on_block(4, 3, 2)
    for g in m.shell_groups:
        g.part = 1                                  # bricks: part 0, shell skin: part 1
    return m


def test_partsav_books_mass_energy_and_momentum_per_part():
    m = two_part_model()
    o = Oracle(m); o.set_print(True)
    o.run_cycles(4)
    b = o.balance(); ps = b["partsav"]
    assert ps.shape == (2, 6)
    # mass: rho * volume of each family (bricks: rho * VNEW = rho0 * V0 exactly for a Lagrangian brick)
    mb = sum(g.mat.rho0 * m.vol0[g.nft:g.nft + g.nel].sum() for g in m.solid_groups)
    area = meshgen.shell_areas(m.X, m.ixc)
    ms = sum(g.mat.rho0 * g.prop.thick * area[g.nft:g.nft + g.nel].sum() for g in m.shell_groups)
    assert np.isclose(ps[0, 5], mb, rtol=1e-12) and np.isclose(ps[1, 5], ms, rtol=1e-12)
    # internal energy: EINT * VOL of the bricks, EINT(1) + EINT(2) of the shells -- before the hourglass energy QEPH adds later in
    # the same cycle (CBILAN sits ahead of CZFINTN1), hence the loose bound on the shells
    eb = (o.solid_state("eint")[0] * o.solid_state("vol")[0]).sum()
    assert np.isclose(ps[0, 0], eb, rtol=1e-12)
    es = o.shell_state("eint").sum()
    assert abs(ps[1, 0] - es) <= 0.05 * abs(es) and ps[1, 0] != 0.0
    assert np.isclose(b["enint"], ps[:, 0].sum(), rtol=1e-14)
    assert np.isclose(b["xmass"], m.MS.sum(), rtol=1e-13)


def test_ecrit_kinetic_energy_is_taken_at_the_full_step():
    """ENCIN = sum 1/2 m |V(n-1/2) + DT1/2 A|^2 = the mean of the two half-step velocities when DT1 = DT2."""
    m = meshgen.hex_block(3, 3, 4, 3.0, 3.0, 4.0, vrand=20.0)
    o = Oracle(m); o.set_print(True)
    o.run_cycles(3)
    v0 = o.download_nodes(("V",))["V"]
    o.run_cycles(1)
    t = o.time(); v1 = o.download_nodes(("V",))["V"]
    vn = v0 + 0.5 * t["dt1"] / t["dt12"] * (v1 - v0)
    assert np.isclose(o.balance()["encin"], 0.5 * (m.MS[:, None] * vn ** 2).sum(), rtol=1e-12)
    mom = (m.MS[:, None] * vn).sum(0)
    b = o.balance()
    assert np.allclose([b["xmomt"], b["ymomt"], b["zmomt"]], mom, rtol=1e-10, atol=1e-12 * np.abs(m.MS[:, None] * vn).sum())


def test_energy_balance_closes_with_the_imposed_velocity_work():
    """Two shells pulled by FIXVEL (the ELEM_SAMP deck): internal + kinetic energy = the work FIXVEL books, cycle after cycle."""
    import qa_decks
    o = Oracle(qa_decks.elem_samp()); o.set_print(True)
    for _ in range(5):
        o.run_cycles(100)
        b = o.balance()
        assert b["wfext"] > 0.0
        assert abs(b["enint"] + b["encin"] + b["enrot"] - b["wfext"]) <= 2e-3 * b["wfext"]
