"""Print-cycle balances of the CUDA path (PARTSAV(1:6, part), the global line of ECRIT, the imposed-velocity work) against the
oracle: same quantities at the same point of the cycle, fixed-order device reduction vs the oracle's sequential sums -> 1e-12."""
import numpy as np
import pytest
import torch
from openradioss_b200 import meshgen
from test_oracle_balance import two_part_model

pytestmark = pytest.mark.gpu
if torch.cuda.is_available():
    from openradioss_b200.engine import Engine
    from oracle.orc import Oracle

TOL = 1e-12


def compare(m, ncycles, phased=False):
    g, o = Engine(m), Oracle(m)
    g.set_print(True); o.set_print(True)
    for c in range(ncycles):
        if phased:
            dt1 = o.time()["dt2"]
            for b in (g, o):
                b.forces_phase(dt1); b.assemble()
            dt2 = o.time()["dt2t"]
            for b in (g, o):
                b.advance(0.5 * (dt1 + dt2), dt2)
        else:
            g.run_cycles(1); o.run_cycles(1)
        bg, bo = g.balance(), o.balance()
        scale = max(abs(bo["enint"]), abs(bo["encin"]), 1e-300)
        for k in ("encin", "enrot", "enint", "wfext"):
            assert abs(bg[k] - bo[k]) <= TOL * max(scale, abs(bo["wfext"])), (c, k, bg[k], bo[k])
        assert abs(bg["xmass"] - bo["xmass"]) <= TOL * bo["xmass"]
        pm = np.abs(m.MS).sum() * max(np.abs(o.download_nodes(("V",))["V"]).max(), 1e-300)
        for k in ("xmomt", "ymomt", "zmomt"):
            assert abs(bg[k] - bo[k]) <= 1e-11 * pm, (c, k)
        pg, po = bg["partsav"], bo["partsav"]
        assert pg.shape == po.shape
        assert np.abs(pg[:, :2] - po[:, :2]).max() <= TOL * max(np.abs(po[:, :2]).max(), 1e-300), (c, pg, po)
        assert np.abs(pg[:, 2:5] - po[:, 2:5]).max() <= 1e-11 * pm
        assert np.allclose(pg[:, 5], po[:, 5], rtol=1e-13)
    return bg


def test_brick_balances():
    compare(meshgen.taylor_bar(scale=8), 6)


def test_qeph_plate_balances():
    compare(meshgen.shell_plate(9, 7, 90.0, 70.0, pressure=3.0, vrand=4.0), 6)


def test_bt_law2_plate_balances():
    prop = meshgen.default_prop_shell(thick=1.2, ihbe=1, npt=3, ipla=1)
    compare(meshgen.shell_plate(8, 8, 80.0, 80.0, law=2, prop=prop, pressure=2.0, vrand=4.0), 5)


def test_mixed_quads_and_triangles_balances():
    compare(meshgen.tri_plate(8, 6, 80.0, 60.0, quads="checker", pressure=10.0, vrand=5.0), 5)


def test_two_parts_bricks_and_shells():
    b = compare(two_part_model(), 5)
    assert b["partsav"].shape == (2, 6) and (b["partsav"][:, 5] > 0).all()


def test_phased_cycle_books_the_same_balances():
    compare(meshgen.shell_plate(6, 6, 60.0, 60.0, pressure=3.0, vrand=4.0), 4, phased=True)


def test_imposed_velocity_work():
    b = compare(meshgen.crush_tube(4, 6, 1, ramp=0.01), 30)
    assert b["wfext"] > 0.0


def test_print_toggle_keeps_the_run_bitwise():
    """Booking the balances must not perturb the cycle: X after 20 cycles is bit-identical with and without IPRI."""
    m = meshgen.shell_plate(9, 7, 90.0, 70.0, pressure=3.0, vrand=4.0)
    a, b = Engine(m), Engine(m)
    b.set_print(True)
    a.run_cycles(20); b.run_cycles(10); b.set_print(False); b.run_cycles(10)
    assert np.array_equal(a.download_nodes(("X",))["X"], b.download_nodes(("X",))["X"])
