"""Known answers for the load / kinematic-condition restatements of the oracle (CPU):
FINTER time functions on the nodal loads (force.F90) and imposed velocities (fixvel.F)."""
import numpy as np
from openradioss_b200 import meshgen
from oracle.orc import Oracle


def test_fixvel_drives_the_nodes_to_the_prescribed_velocity():
    m = meshgen.crush_tube(4, 5, 1, ramp=0.004)
    o = Oracle(m)
    top = m.ibfv[:, 0] - 1
    for c in range(40):
        t0 = o.time()
        o.run_cycles(1)
        t1 = o.time()
        # TSC = (TT + DT2/2) * FACX at the start of the cycle; curve = ramp to 1 at `ramp`, then flat
        tsc = t0["tt"] + 0.5 * t1["dt2"]
        want = -10.0 * min(tsc / 0.004, 1.0)
        vz = o.download_nodes(("V",))["V"][top, 2]
        assert np.allclose(vz, want, rtol=1e-13, atol=0.0), (c, vz[:2], want)
    assert o.time()["tt"] > 0.004                      # the flat part of the curve was reached
    d = o.download_nodes(("D", "V"))
    other = np.setdiff1d(np.arange(m.numnod), top)
    assert np.abs(d["V"][other]).max() > 0.0           # the wave has left the driven ring
    assert np.abs(d["V"][m.icodt == 1, 2]).max() == 0.0  # anvil nodes stay z-fixed


def test_load_time_function_scales_the_nodal_loads():
    """A += FEXT * FINTER(f, TT): during the ramp the first-cycle acceleration is proportional to the curve."""
    acc = []
    for tau in (0.0, 1.0):                            # constant load vs ramp that is still ~0 at the first cycles
        m = meshgen.shell_plate(6, 6, 60.0, 60.0, pulse_tau=tau, vrand=0.0, jitter=0.0, zjitter=0.0)
        o = Oracle(m)
        o.forces_phase(0.0); o.assemble()
        acc.append(o.download_nodes(("A",))["A"])
    assert np.abs(acc[0][:, 2]).max() > 0.0
    assert np.abs(acc[1]).max() == 0.0                 # f(0) = 0
    m = meshgen.shell_plate(6, 6, 60.0, 60.0, pulse_tau=0.05, vrand=0.0, jitter=0.0, zjitter=0.0)
    o = Oracle(m); o.run_cycles(5)
    tt = o.time()["tt"]
    o.forces_phase(o.time()["dt2"])
    # A after ASSPAR4 = pre-loaded external load + the node's corner rows
    scale = min(tt / 0.05, 1.0)
    assert 0.0 < scale < 1.0
    o.assemble()
    a = o.download_nodes(("A",))["A"]
    fs = o.download_fsky()
    internal = np.zeros_like(a)
    for n in range(m.numnod):
        internal[n] = fs[m.adsky[n] - 1:m.adsky[n + 1] - 1, :3].sum(0)
    assert np.allclose(a - internal, m.fext * scale, rtol=1e-9, atol=1e-12 * max(np.abs(internal).max(), np.abs(m.fext).max()))


def test_add_function_appends_a_curve():
    m = meshgen.shell_plate(2, 2, 10.0, 10.0)
    n0 = len(m.npf) - 1
    k = meshgen.add_function(m, [0.0, 1.0, 2.0], [0.0, 5.0, 5.0])
    assert k == n0 and m.npf[-1] - m.npf[-2] == 3 and list(m.tf[-6:]) == [0.0, 0.0, 1.0, 5.0, 2.0, 5.0]


def test_gravity_is_a_free_fall_for_an_unloaded_block():
    """GRAVIT (gravit.F:84-160): every node gets A(3) += FCY after ACCELE, so a block without other loads falls rigidly:
    V = sum of DT12 * g, no strain, no internal force; a second load with a time function adds FCY2 * f(TT * FCX)."""
    import pytest
    m = meshgen.hex_block(3, 3, 3, 6.0, 6.0, 6.0)
    meshgen.add_gravity(m, 3, -9.81e-3)
    meshgen.add_gravity(m, 1, 2.0e-3, curve=([0.0, 1.0e-3, 1.0], [0.0, 1.0, 1.0]), fcx=2.0)
    o = Oracle(m)
    vz = vx = 0.0
    for c in range(30):
        t0 = o.time()["tt"]
        o.run_cycles(1)
        t = o.time()
        vz += t["dt12"] * -9.81e-3
        vx += t["dt12"] * 2.0e-3 * min(t0 * 2.0 / 1.0e-3, 1.0)
        v = o.download_nodes(("V",))["V"]
        tol = 1e-13 if c == 0 else 1e-5        # later cycles: K * (rounding of rho / rho0 - 1) is a real, 1e-11 N force on these tiny velocities
        assert np.allclose(v[:, 2], vz, rtol=tol, atol=0.0) and np.allclose(v[:, 0], vx, rtol=tol, atol=1e-300), c
        assert np.abs(v[:, 1]).max() <= 1e-5 * abs(vz)          # unloaded direction: the same rounding-level forces only
    assert o.time()["tt"] * 2.0 > 1.0e-3                               # the flat part of the curve was reached
    assert np.abs(o.solid_state("sig")).max() < 1e-8                   # rigid motion: no stress beyond rounding
    assert vx > 0.0


def test_gravity_acts_before_the_boundary_conditions_and_only_on_its_nodes():
    """resol.F order ACCELE (:6921) -> GRAVIT (:7123) -> BCS10 (:7322): fixed dofs stay fixed; nodes outside IB feel it only
    through the elements"""
    m = meshgen.hex_block(2, 2, 4, 4.0, 4.0, 8.0, fix_bottom_z=True)
    top = np.nonzero(m.X[:, 2] > 7.0)[0]
    meshgen.add_gravity(m, 3, -5.0, nodes=top)
    o = Oracle(m)
    o.run_cycles(1)
    v = o.download_nodes(("V",))["V"]
    dt12 = o.time()["dt12"]
    rest = np.setdiff1d(np.arange(m.numnod), top)
    assert np.allclose(v[top, 2], -5.0 * dt12, rtol=1e-13) and np.abs(v[rest]).max() == 0.0
    o.run_cycles(60)
    v = o.download_nodes(("V",))["V"]
    assert np.abs(v[m.icodt == 1, 2]).max() == 0.0 and np.abs(v[rest, 2]).max() > 0.0


def test_gravity_follows_its_nodes_into_the_domains():
    from openradioss_b200 import domdec, spmd
    m = meshgen.hex_block(4, 3, 6, 8.0, 6.0, 12.0, fix_bottom_z=True, vrand=1.0)
    meshgen.add_gravity(m, 3, -9.81e-3)
    meshgen.add_gravity(m, 2, 4.0e-3, nodes=np.nonzero(m.X[:, 2] > 7.0)[0], curve=([0.0, 1.0e-4, 1.0], [0.0, 1.0, 1.0]))
    ref = Oracle(m)
    ref.run_cycles(25)
    doms = [domdec.decompose_strips(m, 3, r, axis=2) for r in range(3)]
    assert sum(int(d.model.igrv[1, 0]) for d in doms) >= int(m.igrv[1, 0]) and all(len(d.model.igrv) == 2 for d in doms)
    backs = [Oracle(d.model) for d in doms]
    spmd.run_local(backs, doms, 25)
    xr = ref.download_nodes(("X", "V"))
    for b, d in zip(backs, doms):
        x = b.download_nodes(("X", "V"))
        assert np.array_equal(x["X"], xr["X"][d.node_gid]) and np.array_equal(x["V"], xr["V"][d.node_gid])



def test_long_time_function_takes_the_dichotomy_branch_of_finter():
    """FINTER with 20 segments or more (finter.F:231-356: end segments first, dichotomy, walk over the reduced interval) must
    give what the classical walk gives: the linear interpolant, taken from the nearer end point, end segments extrapolating.
    A free-falling block under a load with a 64-point curve sees exactly FCY * f(TT * FCX) as its acceleration."""
    x = np.concatenate([[0.0], np.cumsum(np.linspace(0.5e-4, 2.0e-4, 63))])          # uneven spacing, last point ~ 8e-3
    y = np.sin(400.0 * x) + 2.0 * x
    m = meshgen.hex_block(2, 2, 2, 4.0, 4.0, 4.0)
    meshgen.add_gravity(m, 3, 10.0, curve=(x, y), fcx=3.0)            # large enough that rounding-level internal forces do not show
    o = Oracle(m)
    seen_mid = seen_end = False
    for c in range(400):
        t0 = o.time()["tt"]
        v0 = o.download_nodes(("V",))["V"][0, 2]
        o.run_cycles(1)
        t = o.time()
        xx = t0 * 3.0
        if xx <= x[-1]:
            f = np.interp(xx, x, y); seen_mid = True
        else:                                                    # beyond the last point: the last segment extrapolates
            f = y[-1] + (xx - x[-1]) * (y[-1] - y[-2]) / (x[-1] - x[-2]); seen_end = True
        dv = o.download_nodes(("V",))["V"][0, 2] - v0
        assert dv == __import__("pytest").approx(t["dt12"] * 10.0 * f, rel=1e-7, abs=1e-12), (c, xx)      # rounding-level internal forces act too (see the free-fall test)
    assert seen_mid and seen_end


def _plate_with_records(single_curve):
    """6 x 6 plate; its pressure load rewritten as /CLOAD records (one per node and direction)."""
    m = meshgen.shell_plate(6, 6, 60.0, 60.0, pulse_tau=0.05, vrand=0.0, jitter=0.0, zjitter=0.0)
    fz = m.fext[:, 2].copy()
    nodes = np.nonzero(fz)[0]
    k0 = int(m.load_func[0])
    k1 = meshgen.add_function(m, [0.0, 0.01, 0.2, 1.0], [0.0, 0.3, 1.0, 1.0])
    ib, fac = [], []
    for n in nodes:
        ib.append((n + 1, 3, k0)); fac.append((fz[n], float(m.load_func[1])))
    if not single_curve:                                   # a second record on every other node, another curve, a lateral direction too
        for n in nodes[::2]:
            ib.append((n + 1, 3, k1)); fac.append((0.5 * fz[n], 2.0))
            ib.append((n + 1, 1, k1)); fac.append((0.25 * fz[n], 1.0))
            ib.append((n + 1, 5, -1)); fac.append((1.0e-3 * fz[n], 1.0))          # a constant moment about y
    return m, nodes, fz, np.array(ib, np.int32), np.array(fac, np.float64), (k0, k1)


def test_load_records_with_one_curve_equal_the_nodal_array_path_bitwise():
    m, nodes, fz, ib, fac, _ = _plate_with_records(True)
    a = Oracle(m); a.run_cycles(40)
    m2, _, _, ib, fac, _ = _plate_with_records(True)
    m2.fext = None; m2.mext = None; m2.load_func = None; m2.cload_ib = ib; m2.cload_fac = fac
    b = Oracle(m2); b.run_cycles(40)
    for k in ("X", "V", "VR"):
        assert np.array_equal(a.download_nodes((k,))[k], b.download_nodes((k,))[k])


def test_load_records_with_several_curves_add_up_in_record_order():
    """A(dir, N) starts from the sum of the node's records, each FCY * f(TT * FCX) with its own curve (force.F90:301-312)."""
    m, nodes, fz, ib, fac, (k0, k1) = _plate_with_records(False)
    fcx0 = float(m.load_func[1])
    m.fext = None; m.mext = None; m.load_func = None; m.cload_ib = ib; m.cload_fac = fac
    o = Oracle(m); o.run_cycles(7)
    tt = o.time()["tt"]
    o.forces_phase(o.time()["dt2"]); o.assemble()
    acc = o.download_nodes(("A", "AR"))
    fs = o.download_fsky()
    internal = np.zeros((m.numnod, 6))
    for n in range(m.numnod):
        internal[n] = fs[m.adsky[n] - 1:m.adsky[n + 1] - 1, :6].sum(0)
    curve = lambda k, x: np.interp(x, m.tf[2 * m.npf[k]:2 * m.npf[k + 1]:2], m.tf[2 * m.npf[k] + 1:2 * m.npf[k + 1]:2])
    want = np.zeros((m.numnod, 6))
    for (n, d, k), (fcy, fcx) in zip(ib, fac):
        want[n - 1, d - 1] += fcy * (curve(k, tt * fcx) if k >= 0 else 1.0)
    got = np.hstack([acc["A"], acc["AR"]]) - internal
    scale = max(np.abs(want).max(), np.abs(internal).max())
    assert np.allclose(got, want, rtol=1e-9, atol=1e-12 * scale)
    assert np.abs(want[:, 0]).max() > 0 and np.abs(want[:, 4]).max() > 0 and 0.0 < curve(k1, tt * 2.0) < 1.0


# ---- where FORCE's records enter the nodal sum: /PARITH/ON behind the element rows, /PARITH/OFF before them ---------------------
def test_parith_on_adds_the_load_behind_the_rows_and_parith_off_before_them():
    """/PARITH/ON: FORCE leaves each record in an FSKY row of its own (force.F90:714-1034) and the Starter appends those
    pseudo-elements after all elements of a node (starter/source/spmd/domdec2.F:2363-2388): ASSPAR4 adds the load LAST.
    /PARITH/OFF: A += AA before the element loop (force.F90:182-312).  Both folds are restated here in numpy and must match bit for bit."""
    m = meshgen.shell_plate(6, 5, 60.0, 50.0, pressure=30.0, vrand=5.0)
    res = {}
    for iparit in (1, 0):
        o = Oracle(m); o.set_parith(iparit)
        o.forces_phase(0.0)
        fsky = o.download_fsky()
        o.assemble()
        A = o.download_nodes(("A",))["A"]
        ref = np.zeros_like(A)
        for n in range(m.numnod):
            acc = m.fext[n].copy() if iparit == 0 else np.zeros(3)
            for k in range(m.adsky[n] - 1, m.adsky[n + 1] - 1):
                acc = acc + fsky[k, :3]
            ref[n] = acc + m.fext[n] if iparit == 1 else acc
        assert np.array_equal(A, ref), iparit
        res[iparit] = A
    assert not np.array_equal(res[0], res[1])                      # the order shows in the last bits of some loaded nodes
    assert np.abs(res[0] - res[1]).max() <= 4e-16 * np.abs(res[1]).max()
