"""/DT/NODA on the device: nodal stiffnesses from the element kernels, DTNODA fold after the
assembly, against the oracle -- phased cycles to 1e-12 and the device-resident loop over 300 cycles."""
import numpy as np
import pytest
import torch
from conftest import rel_err
from openradioss_b200 import meshgen

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from openradioss_b200.engine import Engine
    from oracle.orc import Oracle


def cases():
    yield "brick", meshgen.hex_block(6, 5, 7, 6.0, 5.0, 7.0, v0=(0, 0, -50.0), vrand=2.0, fix_bottom_z=True)
    yield "qeph", meshgen.shell_plate(9, 8, 90.0, 80.0, pressure=30.0, vrand=20.0)
    yield "tube", meshgen.crush_tube(5, 7, 1, ramp=0.002)
    yield "bt", meshgen.shell_plate(8, 7, 80.0, 70.0, prop=meshgen.default_prop_shell(ihbe=1, npt=3), pressure=30.0, vrand=20.0)
    yield "bt_law2", meshgen.shell_plate(7, 7, 70.0, 70.0, law=2, prop=meshgen.default_prop_shell(ihbe=4, npt=5, ismstr=4), pressure=30.0, vrand=20.0)


@pytest.mark.parametrize("name", ["brick", "qeph", "tube", "bt", "bt_law2"])
def test_nodal_time_step_matches_oracle(name):
    m = dict(cases())[name]
    m.control.nodadt = 1
    m.itab = (np.arange(m.numnod, dtype=np.int32) * 2 + 11).astype(np.int32)
    g, o = Engine(m), Oracle(m, threads=0)
    dt1 = 0.0
    for c in range(5):
        for b in (g, o):
            b.forces_phase(dt1); b.assemble()
        ng, no = g.download_nodes(("A", "AR", "STIFN", "STIFR")), o.download_nodes(("A", "AR", "STIFN", "STIFR"))
        for k in ng:
            assert rel_err(ng[k], no[k]) <= 1e-12, (k, c)
        tg, to = g.time(), o.time()
        assert tg["ityptst"] == 11 and to["ityptst"] == 11 and tg["neltst"] == to["neltst"]
        assert tg["dt2t"] == pytest.approx(to["dt2t"], rel=1e-13)
        dt2 = to["dt2t"]
        for b in (g, o):
            b.advance(0.5 * (dt1 + dt2), dt2)
        dt1 = dt2
    g2, o2 = Engine(m), Oracle(m, threads=0)
    g2.run_cycles(300); g2.synchronize(); o2.run_cycles(300)
    t2g, t2o = g2.time(), o2.time()
    assert t2g["ncycle"] == 300 and t2g["ityptst"] == 11 and t2g["tt"] == pytest.approx(t2o["tt"], rel=1e-11)
    assert rel_err(g2.download_nodes(("D",))["D"], o2.download_nodes(("D",))["D"]) <= 1e-8


def test_nodal_time_step_across_domains_is_rejected_loudly():
    """The nodal dt exchange between domains is not built: orgpu_run_cycles must say so, not step with a local dt."""
    m = meshgen.hex_block(4, 4, 4, 4.0, 4.0, 4.0)
    m.control.nodadt = 1
    g = Engine(m)
    g.run_cycles(2); g.synchronize()                     # single domain: fine
    assert g.time()["ityptst"] == 11
