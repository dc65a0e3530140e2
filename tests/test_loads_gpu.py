"""Parity of the device-side kinematic conditions and loads against the oracle: imposed velocities
(FIXVEL), time functions on the nodal loads (FORCE / FINTER) and the C4 mixed tube (QEPH shells on a
brick end block, imposed-velocity crush).  Tolerances as north_star: forces 1e-12, displacements 1e-8."""
import numpy as np
import pytest
import torch
from conftest import rel_err
from openradioss_b200 import meshgen, domdec, spmd

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from openradioss_b200.engine import Engine
    from oracle.orc import Oracle


def phased(m, ncyc):
    g, o = Engine(m), Oracle(m, threads=0)
    dt1 = 0.0
    for c in range(ncyc):
        for b in (g, o):
            b.forces_phase(dt1); b.assemble()
        ng, no = g.download_nodes(("A", "AR")), o.download_nodes(("A", "AR"))
        assert rel_err(ng["A"], no["A"]) <= 1e-12 and rel_err(ng["AR"], no["AR"]) <= 1e-12, c
        dt2 = o.time()["dt2t"]
        assert g.time()["dt2t"] == pytest.approx(dt2, rel=1e-13)
        for b in (g, o):
            b.advance(0.5 * (dt1 + dt2), dt2)
        ng, no = g.download_nodes(("X", "V", "VR", "D")), o.download_nodes(("X", "V", "VR", "D"))
        for k in ("X", "V", "VR", "D"):
            assert rel_err(ng[k], no[k]) <= 1e-12, (k, c)
        dt1 = dt2
    return g, o


def test_tube_phased_cycles_match_oracle():
    m = meshgen.crush_tube(6, 8, 1, ramp=0.002)
    g, o = phased(m, 12)
    top = m.ibfv[:, 0] - 1
    vg, vo = g.download_nodes(("V",))["V"], o.download_nodes(("V",))["V"]
    assert np.array_equal(vg[top, 2], vo[top, 2])          # the prescribed dof is bit-identical (no libm on its path)


def test_tube_device_loop_400_cycles():
    m = meshgen.crush_tube(8, 10, 2, ramp=0.01)
    g, o = Engine(m), Oracle(m, threads=0)
    g.run_cycles(400); g.synchronize(); o.run_cycles(400)
    tg, to = g.time(), o.time()
    assert tg["ncycle"] == to["ncycle"] and tg["tt"] == pytest.approx(to["tt"], rel=1e-12)
    ng, no = g.download_nodes(("D", "V")), o.download_nodes(("D", "V"))
    assert rel_err(ng["D"], no["D"]) <= 1e-8 and rel_err(ng["V"], no["V"]) <= 1e-8
    top = m.ibfv[:, 0] - 1
    assert np.allclose(ng["V"][top, 2], -10.0 * min(1.0, (to["tt"] - 0.5 * to["dt2"]) / 0.01), rtol=1e-12)
    assert np.abs(ng["D"][top, 2]).max() > 1e-3            # the ring really moved


def test_plate_pressure_pulse_matches_oracle():
    m = meshgen.shell_plate(10, 9, 100.0, 90.0, pulse_tau=5.0e-3, vrand=2.0)
    phased(m, 6)
    g, o = Engine(m), Oracle(m, threads=0)
    g.run_cycles(300); g.synchronize(); o.run_cycles(300)
    assert rel_err(g.download_nodes(("D",))["D"], o.download_nodes(("D",))["D"]) <= 1e-8


@pytest.mark.parametrize("iparit", [1, 0])
def test_load_enters_the_nodal_sum_where_the_reference_adds_it(iparit):
    """/PARITH/ON (default): behind the element rows, where ASSPAR4 finds FORCE's own FSKY rows; /PARITH/OFF: before them.  The
    device fold is restated in numpy from the device's own rows and must match bit for bit, in both modes."""
    m = meshgen.shell_plate(12, 9, 120.0, 90.0, pressure=30.0, vrand=5.0)
    g = Engine(m); g.set_parith(iparit)
    g.forces_phase(0.0)
    fsky = g.download_fsky()
    g.assemble()
    A = g.download_nodes(("A",))["A"]
    ref = np.zeros_like(A)
    for n in range(m.numnod):
        acc = m.fext[n].copy() if iparit == 0 else np.zeros(3)
        for k in range(m.adsky[n] - 1, m.adsky[n + 1] - 1):
            acc = acc + fsky[k, :3]
        ref[n] = acc + m.fext[n] if iparit == 1 else acc
    assert np.array_equal(A, ref)
    o = Oracle(m, threads=0); o.set_parith(iparit); o.forces_phase(0.0); o.assemble()
    assert rel_err(A, o.download_nodes(("A",))["A"]) <= 1e-12
    # the fused device loop (same gather) against the oracle's loop in the same mode
    g2 = Engine(m); g2.set_parith(iparit); g2.run_cycles(30); g2.synchronize()
    o2 = Oracle(m, threads=0); o2.set_parith(iparit); o2.run_cycles(30)
    assert rel_err(g2.download_nodes(("V",))["V"], o2.download_nodes(("V",))["V"]) <= 1e-10


def test_gravity_loads_match_oracle():
    """GRAVIT on the device (node kernel, between ACCELE and BCS): constant g on all nodes + a ramped lateral load on a
    node subset, on a plate under pressure and on a block; phased 1e-12, then 300 cycles of the device loop"""
    m = meshgen.shell_plate(10, 9, 100.0, 90.0, pulse_tau=5.0e-3, vrand=2.0)
    meshgen.add_gravity(m, 3, -9.81e-3)
    meshgen.add_gravity(m, 1, 5.0e-2, nodes=np.nonzero(m.X[:, 0] > 50.0)[0], curve=([0.0, 1.0e-3, 1.0], [0.0, 1.0, 1.0]), fcx=1.5)
    phased(m, 6)
    g, o = Engine(m), Oracle(m, threads=0)
    g.run_cycles(300); g.synchronize(); o.run_cycles(300)
    assert rel_err(g.download_nodes(("D",))["D"], o.download_nodes(("D",))["D"]) <= 1e-8
    # free fall of an unloaded block: the same bits as the oracle (no libm on the path)
    m = meshgen.hex_block(3, 3, 3, 6.0, 6.0, 6.0)
    meshgen.add_gravity(m, 3, -9.81e-3)
    g, o = Engine(m), Oracle(m, threads=0)
    g.run_cycles(40); g.synchronize(); o.run_cycles(40)
    vg, vo = g.download_nodes(("V",))["V"], o.download_nodes(("V",))["V"]
    assert np.array_equal(vg[:, 2], vo[:, 2]) and vo[:, 2].max() < 0.0


def test_long_time_function_on_the_device_is_bitwise_the_oracle():
    """64-point load curve: the dichotomy branch of FINTER (finter.F:231-356) in element_finalize_kernel, same bits as the oracle"""
    x = np.concatenate([[0.0], np.cumsum(np.linspace(0.5e-4, 2.0e-4, 63))])
    y = np.sin(400.0 * x) + 2.0 * x
    m = meshgen.hex_block(3, 3, 3, 6.0, 6.0, 6.0)
    meshgen.add_gravity(m, 3, 1.0e-3, curve=(x, y), fcx=3.0)
    g, o = Engine(m), Oracle(m, threads=0)
    for n in (5, 60, 200):
        g.run_cycles(n); g.synchronize(); o.run_cycles(n)
        vg, vo = g.download_nodes(("V",))["V"], o.download_nodes(("V",))["V"]
        assert np.array_equal(vg[:, 2], vo[:, 2])
    assert o.time()["tt"] * 3.0 > x[-1]                      # the run went past the end of the curve


def test_load_records_match_oracle_and_follow_their_nodes():
    """orgpu_set_cloads: /CLOAD records with several time functions (force.F90:188-312) -- phased 1e-12 and 200 cycles of the
    device loop against the oracle; with one curve bitwise the nodal-array path; two domains bitwise the single domain."""
    from test_oracle_loads import _plate_with_records
    m, nodes, fz, ib, fac, _ = _plate_with_records(False)
    m.fext = None; m.mext = None; m.load_func = None; m.cload_ib = ib; m.cload_fac = fac
    phased(m, 6)
    g, o = Engine(m), Oracle(m, threads=0)
    g.run_cycles(200); g.synchronize(); o.run_cycles(200)
    assert rel_err(g.download_nodes(("D",))["D"], o.download_nodes(("D",))["D"]) <= 1e-8
    ma, _, _, ib1, fac1, _ = _plate_with_records(True)
    mb, _, _, _, _, _ = _plate_with_records(True)
    mb.fext = None; mb.mext = None; mb.load_func = None; mb.cload_ib = ib1; mb.cload_fac = fac1
    a, b = Engine(ma), Engine(mb)
    a.run_cycles(60); b.run_cycles(60); a.synchronize(); b.synchronize()
    assert np.array_equal(a.download_nodes(("X",))["X"], b.download_nodes(("X",))["X"])
    ref = Engine(m); ref.run_cycles(40); ref.synchronize()
    doms = [domdec.decompose_strips(m, 2, r) for r in range(2)]
    backs = [Engine(d.model) for d in doms]
    spmd.run_local(backs, doms, 40)
    xr = ref.download_nodes(("X",))["X"]
    for bk, d in zip(backs, doms):
        assert np.array_equal(bk.download_nodes(("X",))["X"], xr[d.node_gid])
    mc, _, _, ib2, fac2, _ = _plate_with_records(True)
    mc.cload_ib = ib2; mc.cload_fac = fac2                       # records AND nodal arrays: rejected
    with pytest.raises(RuntimeError, match="not both"):
        Engine(mc)


def test_gravity_rejects_what_is_not_built():
    m = meshgen.hex_block(2, 2, 2, 2.0, 2.0, 2.0)
    meshgen.add_gravity(m, 3, -1.0)
    m.igrv[0, 1] = 13                                   # IGRV(2) = 10*ISK + N2: a skew frame
    with pytest.raises(RuntimeError, match="skew"):
        Engine(m)


@pytest.mark.parametrize("nproc", [2, 3])
def test_tube_domains_bitwise(nproc):
    """z-slab decomposition of the mixed tube (host-staged exchange): same bits as one domain, with the
    imposed-velocity records and the load function following their nodes into the domains."""
    m = meshgen.crush_tube(5, 9, 1, ramp=0.002)
    ref = Engine(m)
    doms = [domdec.decompose_strips(m, nproc, r, axis=2) for r in range(nproc)]
    assert sum(len(d.model.ibfv) if d.model.ibfv is not None else 0 for d in doms) >= len(m.ibfv)
    backs = [Engine(d.model) for d in doms]
    st = spmd.initial_state(m.control)
    ncyc = 25
    for c in range(ncyc):
        dt1 = st["dt2"]
        ref.forces_phase(dt1); ref.assemble()
        dt2 = min(spmd.EP06, ref.time()["dt2t"], float(np.float32(1.1)) * st["dt2old"], st["dtmx"])
        ref.advance(0.5 * (dt1 + dt2), dt2); st["dt2"] = dt2; st["dt2old"] = dt2
    spmd.run_local(backs, doms, ncyc)
    xr = ref.download_nodes(("X", "V"))
    for b, d in zip(backs, doms):
        x = b.download_nodes(("X", "V"))
        assert np.array_equal(x["X"], xr["X"][d.node_gid]) and np.array_equal(x["V"], xr["V"][d.node_gid])
