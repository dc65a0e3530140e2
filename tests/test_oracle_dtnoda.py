"""/DT/NODA restatement of the oracle (CPU): DT2T = DTFAC1(11) * min over nodes of sqrt(2 M / K), from the assembled
nodal stiffnesses (dtnoda.F:221-260, 445-462), ITYPTST = 11, NELTST = user node id."""
import numpy as np
from openradioss_b200 import meshgen
from oracle.orc import Oracle


def _check(m):
    m.control.nodadt = 1
    m.itab = (np.arange(m.numnod, dtype=np.int32) * 3 + 7).astype(np.int32)
    o = Oracle(m)
    dt1 = 0.0
    for c in range(4):
        o.forces_phase(dt1); o.assemble()
        nd = o.download_nodes(("STIFN", "STIFR"))
        t = o.time()
        cand = np.full(m.numnod, np.inf)
        ok = (nd["STIFN"] > 0) & (m.MS > 0)
        cand[ok] = m.control.dtfac_node * np.sqrt(2.0 * m.MS[ok] / nd["STIFN"][ok])
        best, who = cand.min(), int(cand.argmin())
        if m.control.iroddl:
            okr = (nd["STIFR"] > 0) & (m.IN > 0)
            cr = np.full(m.numnod, np.inf); cr[okr] = m.control.dtfac_node * np.sqrt(2.0 * m.IN[okr] / nd["STIFR"][okr])
            if cr.min() < best:
                best, who = cr.min(), int(cr.argmin())
        assert t["ityptst"] == 11 and t["neltst"] == m.itab[who]
        assert abs(t["dt2t"] - best) <= 1e-15 * best
        o.advance(0.5 * (dt1 + t["dt2t"]), t["dt2t"]); dt1 = t["dt2t"]


def test_nodal_time_step_bricks():
    _check(meshgen.hex_block(5, 4, 6, 5.0, 4.0, 6.0, vrand=1.0))


def test_nodal_time_step_qeph_shells():
    _check(meshgen.shell_plate(7, 6, 70.0, 60.0, pressure=20.0, vrand=5.0))


def test_nodal_time_step_bt_shells():
    _check(meshgen.shell_plate(7, 6, 70.0, 60.0, prop=meshgen.default_prop_shell(ihbe=1, npt=3), pressure=20.0, vrand=5.0))


def test_nodal_and_element_time_steps_are_of_the_same_order():
    for mk in (lambda: meshgen.hex_block(5, 5, 5, 5.0, 5.0, 5.0, vrand=1.0), lambda: meshgen.shell_plate(8, 8, 80.0, 80.0, vrand=5.0)):
        dts = []
        for nd in (0, 1):
            m = mk(); m.control.nodadt = nd
            o = Oracle(m); o.run_cycles(10); dts.append(o.time()["dt2"])
        assert 0.5 < dts[1] / dts[0] < 2.0
