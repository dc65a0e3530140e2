"""/FAIL/JOHNSON on shells on the device (generic shell kernels, two more words per integration point) against the oracle:
damage, point flags, stresses and deletions agree cycle by cycle -- for QEPH, Belytschko-Tsay and 3-node shells, LAW36 and LAW2."""
import numpy as np
import pytest
import torch
from conftest import rel_err
from openradioss_b200 import meshgen
from openradioss_b200.model import Fail

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from openradioss_b200.engine import Engine
    from oracle.orc import Oracle


def fail(d1=0.01, d2=0.02, d3=-1.0, d4=0.02, epsp0=1.0e-2, pthk=0.5, pthickg=1.0):
    f = Fail(); f.irupt = 1; f.d1, f.d2, f.d3, f.d4, f.d5 = d1, d2, d3, d4, 0.0
    f.epsp0, f.epsf_min, f.pthk, f.pthickg = epsp0, 0.0, pthk, pthickg
    return f


def with_fail(m, f):
    for g in list(m.shell_groups) + list(m.sh3n_groups):
        g.fail = f
    return m


def run_and_compare(m, state, blocks=6, per=10):
    g, o = Engine(m), Oracle(m, threads=0)
    dead = []
    for c in range(blocks):
        g.run_cycles(per); g.synchronize(); o.run_cycles(per)
        sg, so = getattr(g, state), getattr(o, state)
        assert np.array_equal(sg("off"), so("off")), c
        assert np.array_equal(sg("foff"), so("foff")), c
        assert rel_err(sg("dfmax"), so("dfmax")) <= 1e-9, c
        dead.append(int((so("off") == 0).sum()))
        ng, no = g.download_nodes(("X", "V", "VR")), o.download_nodes(("X", "V", "VR"))
        for k in ("X", "V", "VR"):
            assert rel_err(ng[k], no[k]) <= 1e-9, (k, c)
    return g, o, dead


@pytest.mark.parametrize("ihbe", [24, 1])
@pytest.mark.parametrize("law", [36, 2])
def test_quads_fail_in_the_same_cycles(ihbe, law):
    m = meshgen.shell_plate(10, 9, 100.0, 90.0, law=law, prop=meshgen.default_prop_shell(ihbe=ihbe, npt=5), pressure=60.0, vrand=8.0)
    with_fail(m, fail())
    g, o, dead = run_and_compare(m, "shell_state")
    ne = m.numelc
    assert 0 < dead[-1] < ne and dead[-1] > dead[0]
    foff = o.shell_state("foff")
    assert (foff == 0).any() and ((foff == 0).sum(0)[o.shell_state("off")[0] == 1] < 5).all()       # broken points in living elements too
    sig = g.shell_state("sig")
    for ip in range(5):
        assert np.all(sig[5 * ip:5 * ip + 5][:, foff[ip] == 0.0] == 0.0)                               # a failed point keeps no stress


def test_phased_cycles_with_failure_hold_the_force_tolerance():
    m = meshgen.shell_plate(12, 8, 120.0, 80.0, prop=meshgen.default_prop_shell(npt=3), pressure=40.0, vrand=10.0)
    with_fail(m, fail(d1=0.004, d2=0.0, d4=0.0, pthk=-0.6))
    g, o = Engine(m), Oracle(m, threads=0)
    dt1 = 0.0
    for c in range(25):
        for b in (g, o):
            b.forces_phase(dt1)
        fg, fo = g.download_fsky(), o.download_fsky()
        assert rel_err(fg[:, :3], fo[:, :3]) <= 1e-12 and rel_err(fg[:, 3:6], fo[:, 3:6]) <= 1e-12, c
        dt2 = o.time()["dt2t"]
        assert g.time()["dt2t"] == pytest.approx(dt2, rel=1e-13)
        for b in (g, o):
            b.assemble(); b.advance(0.5 * (dt1 + dt2), dt2)
        dt1 = dt2
    assert np.array_equal(g.shell_state("foff"), o.shell_state("foff")) and (o.shell_state("foff") == 0).any()
    assert rel_err(g.shell_state("dfmax"), o.shell_state("dfmax")) <= 1e-11


def test_triangles_and_mixed_plate():
    m = meshgen.tri_plate(10, 8, 100.0, 80.0, quads="checker", pressure=60.0, vrand=8.0)
    with_fail(m, fail())
    g, o = Engine(m), Oracle(m, threads=0)
    for c in range(6):
        g.run_cycles(10); g.synchronize(); o.run_cycles(10)
        assert np.array_equal(g.shell_state("off"), o.shell_state("off")) and np.array_equal(g.sh3n_state("off"), o.sh3n_state("off"))
        assert np.array_equal(g.sh3n_state("foff"), o.sh3n_state("foff"))
        assert rel_err(g.download_nodes(("X",))["X"], o.download_nodes(("X",))["X"]) <= 1e-9
    assert (o.sh3n_state("off") == 0).any() and (o.shell_state("off") == 0).any()


def test_restart_carries_damage_and_point_flags():
    m = meshgen.shell_plate(10, 9, 100.0, 90.0, prop=meshgen.default_prop_shell(npt=5), pressure=60.0, vrand=8.0)
    with_fail(m, fail())
    a = Engine(m); a.run_cycles(35); a.synchronize()
    ck = a.checkpoint()
    assert (ck["shell"]["foff"] == 0).any()
    b = Engine(m); b.restore(ck)
    a.run_cycles(30); b.run_cycles(30); a.synchronize(); b.synchronize()
    assert np.array_equal(a.download_nodes(("X",))["X"], b.download_nodes(("X",))["X"])
    assert np.array_equal(a.shell_state("dfmax"), b.shell_state("dfmax")) and np.array_equal(a.shell_state("off"), b.shell_state("off"))


def test_two_domains_delete_the_same_elements():
    """host-staged exchange, two handles on one GPU: the failure state lives with the element, the decomposition carries the card"""
    from openradioss_b200 import domdec, spmd
    m = meshgen.shell_plate(12, 9, 120.0, 90.0, prop=meshgen.default_prop_shell(npt=5), pressure=60.0, vrand=8.0)
    with_fail(m, fail())
    ref = Engine(m); ref.run_cycles(50); ref.synchronize()
    doms = [domdec.decompose_strips(m, 2, r) for r in range(2)]
    backs = [Engine(d.model) for d in doms]
    spmd.run_local(backs, doms, 50)
    xr = ref.download_nodes(("X",))["X"]
    for b, d in zip(backs, doms):
        assert np.array_equal(b.download_nodes(("X",))["X"], xr[d.node_gid])
    assert (ref.shell_state("off") == 0).any()


def test_law2_bricks_fail_and_relax_in_the_same_cycles():
    """FAIL_JOHNSON behind MMAIN on LAW2 bricks: an impacting bar whose front elements fail, relax by 0.8 per cycle and are
    deleted -- OFF bit-identical to the oracle in every block of cycles, damage 1e-9, restart bitwise."""
    m = meshgen.hex_block(5, 5, 12, 1.0, 1.0, 3.96, v0=(0, 0, -227.0), fix_bottom_z=True, vrand=5.0, user_id_perm=True)
    f = fail(d1=0.04, d2=0.05, d3=-1.0, d4=0.01, epsp0=1.0e-2)
    for grp in m.solid_groups:
        grp.fail = f
    g, o = Engine(m), Oracle(m, threads=0)
    dead = []
    for c in range(8):
        g.run_cycles(40); g.synchronize(); o.run_cycles(40)
        assert np.array_equal(g.solid_state("off"), o.solid_state("off")), c
        assert rel_err(g.solid_state("dfmax"), o.solid_state("dfmax")) <= 1e-9, c
        assert rel_err(g.download_nodes(("X",))["X"], o.download_nodes(("X",))["X"]) <= 1e-9, c
        dead.append(int((o.solid_state("off") == 0).sum()))
    assert 0 < dead[-1] < m.numels
    off = o.solid_state("off")[0]
    a = Engine(m); a.run_cycles(100); a.synchronize()
    ck = a.checkpoint(); assert "dfmax" in ck["solid"]
    b = Engine(m); b.restore(ck)
    a.run_cycles(60); b.run_cycles(60); a.synchronize(); b.synchronize()
    assert np.array_equal(a.download_nodes(("X",))["X"], b.download_nodes(("X",))["X"]) and np.array_equal(a.solid_state("off"), b.solid_state("off"))
    m36 = meshgen.hex_block(2, 2, 2, 1.0, 1.0, 1.0, law=36)
    m36.solid_groups[0].fail = f
    with pytest.raises(RuntimeError):
        Engine(m36)


def test_rejected_outside_its_envelope():
    m = meshgen.shell_plate(4, 4, 40.0, 40.0)
    f = fail(); f.d5 = 0.3
    with_fail(m, f)
    with pytest.raises(RuntimeError, match="D5"):
        Engine(m)
    m = meshgen.shell_plate(4, 4, 40.0, 40.0)
    f = fail(); f.irupt = 7
    with_fail(m, f)
    with pytest.raises(RuntimeError, match="outside the built path"):
        Engine(m)
