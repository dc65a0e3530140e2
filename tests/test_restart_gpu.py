"""State hand-over: a run cut in two -- device state downloaded through the C ABI, uploaded into
a fresh engine, continued -- gives the same bits as the uninterrupted run (what a restart file round trip
through ELBUF_TAB / WRRESTP must preserve)."""
import numpy as np
import pytest
import torch
from openradioss_b200 import meshgen

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from openradioss_b200.engine import Engine


def cases():
    yield "qeph_law36", meshgen.shell_plate(11, 9, 110.0, 90.0, pressure=40.0, vrand=30.0, pulse_tau=0.01)
    yield "bt_law2", meshgen.shell_plate(9, 8, 90.0, 80.0, law=2, prop=meshgen.default_prop_shell(ihbe=1, npt=3), pressure=30.0, vrand=30.0)
    yield "brick", meshgen.hex_block(6, 5, 7, 1.2, 1.0, 1.4, v0=(0, 0, -150.0), vrand=5.0, fix_bottom_z=True)
    yield "tube", meshgen.crush_tube(6, 8, 1, ramp=0.002)
    yield "sh3n_mixed", meshgen.tri_plate(9, 8, 90.0, 80.0, quads="checker", pressure=40.0, vrand=30.0)
    yield "brick_law36", meshgen.hex_block(6, 5, 7, 12.0, 10.0, 14.0, law=36, v0=(0, 0, -60.0), vrand=20.0, fix_bottom_z=True,
                                           prop=meshgen.default_prop_solid(istrain=1))


@pytest.mark.parametrize("name", ["qeph_law36", "bt_law2", "brick", "tube", "sh3n_mixed", "brick_law36"])
def test_checkpoint_restore_continues_bitwise(name):
    m = dict(cases())[name]
    a = Engine(m); a.run_cycles(120); a.synchronize()
    b = Engine(m); b.run_cycles(50); ck = b.checkpoint()
    c = Engine(m); c.restore(ck); c.run_cycles(70); c.synchronize()
    ta, tc = a.time(), c.time()
    assert ta["tt"] == tc["tt"] and ta["ncycle"] == tc["ncycle"] and ta["dt2"] == tc["dt2"]
    na, nc = a.download_nodes(("X", "V", "VR", "D")), c.download_nodes(("X", "V", "VR", "D"))
    for k in na:
        assert np.array_equal(na[k], nc[k]), (name, k)
    if m.numelc:
        for f in ("sig", "pla", "forc", "eint", "hourg", "thk"):
            assert np.array_equal(a.shell_state(f), c.shell_state(f)), (name, f)
    if m.numels:
        for f in ("sig", "pla", "eint", "rho") + (("wpla", "stra") if name == "brick_law36" else ()):
            assert np.array_equal(a.solid_state(f), c.solid_state(f)), (name, f)
    if m.numeltg:
        for f in ("sig", "pla", "forc", "eint", "thk", "smstr"):
            assert np.array_equal(a.sh3n_state(f), c.sh3n_state(f)), (name, f)
