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
