"""BASELINE.json's full sizes: C1 = 100 352 bricks x 1000 cycles, C2 = 1 000 000 QEPH shells (LAW36, NPT=5), C5 slab = 2 000 000
bricks (LAW2), C4 = 2.0 M shells + 501 k bricks.
* Parity against the oracle AT these sizes (the oracle on all host cores needs ~50 ms per cycle for 1 M shells): corner rows,
  assembled forces, time step + controlling element, nodal update and every state field over phased cycles of the yielding C2
  plate (7813 tiles: tile boundaries, the wave-ahead prefetches and the 7813-candidate dt fold are all in play) and of the 2 M
  brick slab; the full Taylor bar over 1000 cycles of the device loop (displacements / energies 1e-8).
* Size-independent properties: checksum identity run to run, symmetry, conservation laws, kinematic conditions honoured."""
import zlib
import numpy as np
import pytest
import torch
from conftest import rel_err, rel_err_rows
from openradioss_b200 import meshgen

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from openradioss_b200.engine import Engine
    from oracle.orc import Oracle


def test_c2_plate_1m_phased_cycles_match_the_oracle():
    """C2 at full size, driven into yield by the bench's initial velocity field (x3): 4 phased cycles against the oracle."""
    m = meshgen.shell_plate(1000, 1000, 1000.0, 1000.0, pulse_tau=0.05, vwave=(180.0, 100.0))
    assert m.numelc == 1_000_000
    g, o = Engine(m), Oracle(m, threads=0)
    dt1 = 0.0
    for c in range(4):
        for b in (g, o):
            b.forces_phase(dt1)
        fg, fo = g.download_fsky(), o.download_fsky()
        assert rel_err(fg[:, :3], fo[:, :3]) <= 1e-12 and rel_err(fg[:, 3:6], fo[:, 3:6]) <= 1e-12 and rel_err(fg[:, 6:], fo[:, 6:]) <= 1e-12
        if c > 0:                                                   # row by row, each corner row against its own size
            assert rel_err_rows(fg[:, :3], fo[:, :3]) <= 1e-9 and rel_err_rows(fg[:, 3:6], fo[:, 3:6]) <= 1e-9
        del fg, fo
        tg, to = g.time(), o.time()
        assert tg["dt2t"] == pytest.approx(to["dt2t"], rel=1e-13) and tg["neltst"] == to["neltst"] and tg["ityptst"] == 3
        for b in (g, o):
            b.assemble()
        ng, no = g.download_nodes(("A", "AR", "STIFN")), o.download_nodes(("A", "AR", "STIFN"))
        for k in ("A", "AR", "STIFN"):
            assert rel_err(ng[k], no[k]) <= 1e-12, (k, c)
        dt2 = to["dt2t"]
        for b in (g, o):
            b.advance(0.5 * (dt1 + dt2), dt2)
        ng, no = g.download_nodes(("X", "V", "VR", "D")), o.download_nodes(("X", "V", "VR", "D"))
        for k in ("X", "V", "VR", "D"):
            assert rel_err(ng[k], no[k]) <= 1e-13, (k, c)
        dt1 = dt2
    for f in ("forc", "mom", "eint", "thk", "off", "stra", "epsd", "hourg", "smstr", "sig", "pla", "epsd_ip"):
        a, b = g.shell_state(f), o.shell_state(f)
        assert rel_err(a, b) <= 1e-11, (f, rel_err(a, b))
    pla = o.shell_state("pla")
    assert (pla > 0).mean() > 0.5                                   # most integration points took the plastic return


def test_c5_slab_2m_bricks_phased_cycles_match_the_oracle():
    m = meshgen.hex_block(200, 200, 50, 200.0, 200.0, 50.0, vrand=1.0, vseed=12345)
    g, o = Engine(m), Oracle(m, threads=0)
    dt1 = 0.0
    for c in range(3):
        for b in (g, o):
            b.forces_phase(dt1)
        fg, fo = g.download_fsky(), o.download_fsky()
        assert rel_err(fg, fo) <= 1e-12
        if c > 0:
            assert rel_err_rows(fg[:, :3], fo[:, :3]) <= 1e-9
        # size of the terms ASSPAR4 adds up.  After the first cycle the rows themselves agree to ~1e-10 of it only: the pressure
        # K (rho / rho0 - 1) of this field sits at volumetric strains of 1e-6, so the 1-2 ulp by which CUDA's cbrt / pow differ
        # from glibc's in MQVISCB / the density update come back six digits larger (rel_err_rows above bounds them row by row)
        fscale = np.abs(fo[:, :3]).max()
        del fg, fo
        tg, to = g.time(), o.time()
        assert tg["dt2t"] == pytest.approx(to["dt2t"], rel=1e-14) and tg["neltst"] == to["neltst"] and tg["ityptst"] == 1
        for b in (g, o):
            b.assemble()
        ng, no = g.download_nodes(("A", "STIFN")), o.download_nodes(("A", "STIFN"))
        assert np.abs(ng["A"] - no["A"]).max() <= (1e-11 if c == 0 else 1e-9) * fscale and rel_err(ng["STIFN"], no["STIFN"]) <= 1e-12
        dt2 = to["dt2t"]
        for b in (g, o):
            b.advance(0.5 * (dt1 + dt2), dt2)
        ng, no = g.download_nodes(("X", "V", "D")), o.download_nodes(("X", "V", "D"))
        for k in ("X", "V", "D"):
            assert rel_err(ng[k], no[k]) <= (1e-12 if c == 0 else 1e-9), k     # V += DT12 * A, |DT12 A| ~ |V| for this field
        dt1 = dt2
    for f in ("sig", "eint", "rho", "qvis", "pla", "epsd", "off", "temp", "smstr"):
        assert rel_err(g.solid_state(f), o.solid_state(f)) <= 1e-9, f          # same conditioning (bit-exact fields stay bit-exact: test_brick_gpu.py)


def test_c1_taylor_bar_full_size_1000_cycles_match_the_oracle():
    """C1 as BASELINE.json names it: 100 352 bricks, LAW2, anvil BC, 1000 cycles of the device loop vs the oracle."""
    m = meshgen.taylor_bar(1)
    assert m.numels == 100_352
    g, o = Engine(m), Oracle(m, threads=0)
    g.set_print(True); o.set_print(True)
    g.run_cycles(1000); o.run_cycles(1000)
    ng, no = g.download_nodes(("D", "V")), o.download_nodes(("D", "V"))
    assert rel_err(ng["D"], no["D"]) <= 1e-8 and rel_err(ng["V"], no["V"]) <= 1e-8
    bg, bo = g.balance(), o.balance()
    etot = bo["enint"] + bo["encin"]
    for k in ("enint", "encin"):
        assert abs(bg[k] - bo[k]) <= 1e-8 * etot, (k, bg[k], bo[k])
    assert bo["enint"] > 0.1 * etot and o.solid_state("pla").max() > 0.1       # the bar has mushroomed
    tg, to = g.time(), o.time()
    assert tg["ncycle"] == to["ncycle"] == 1000 and tg["tt"] == pytest.approx(to["tt"], rel=1e-10)


def _checksum(g, names=("X", "V")):
    d = g.download_nodes(names)
    return [zlib.adler32(np.ascontiguousarray(d[k]).tobytes()) for k in names], d     # /DEBUG/CHKSM uses Adler-32 too


def test_c2_plate_1m_reproducible_symmetric_clamped():
    m = meshgen.shell_plate(1000, 1000, 1000.0, 1000.0, pressure=0.5, jitter=0.0, zjitter=0.0, pulse_tau=0.02)
    assert m.numelc == 1_000_000
    sums = []
    for rep in range(2):
        g = Engine(m); g.run_cycles(60); g.synchronize()
        c, d = _checksum(g, ("X", "V", "VR"))
        sums.append(c)
        if rep == 0:
            assert np.isfinite(d["X"]).all() and np.isfinite(d["V"]).all()
            assert np.abs(d["V"][m.icodt == 7]).max() == 0.0                    # clamped edge
            w = d["V"][:, 2].reshape(1001, 1001)
            assert np.abs(w).max() > 0.0
            assert rel_err(w, w[::-1, :]) <= 1e-9 and rel_err(w, w[:, ::-1]) <= 1e-9 and rel_err(w, w.T) <= 1e-9
            t = g.time(); assert t["ncycle"] == 60 and t["ityptst"] == 3
        del g
    assert sums[0] == sums[1]                                                    # bitwise run-to-run


def test_c5_slab_2m_bricks_momentum_and_energy():
    m = meshgen.hex_block(200, 200, 50, 200.0, 200.0, 50.0, vrand=1.0, vseed=12345)
    assert m.numels == 2_000_000
    p0 = (m.MS[:, None] * m.V).sum(0); ke0 = 0.5 * (m.MS[:, None] * m.V ** 2).sum()
    g = Engine(m); g.run_cycles(100); g.synchronize()
    c1, d = _checksum(g)
    assert np.isfinite(d["V"]).all()
    p1 = (m.MS[:, None] * d["V"]).sum(0)
    scale = (m.MS[:, None] * np.abs(m.V)).sum()
    assert np.abs(p1 - p0).max() <= 1e-11 * scale                              # free body: internal forces sum to zero
    ei, _, ek, _ = g.energies()
    assert ek == pytest.approx(0.5 * (m.MS[:, None] * d["V"] ** 2).sum(), rel=1e-12)
    # a node-to-node random velocity field is mostly hourglass modes: the viscous hourglass forces dissipate it (that
    # work is not booked in EINT), so the balance can only lose energy, never gain
    assert 0.0 < ei and ei + ek <= ke0 * (1.0 + 1e-9)
    g2 = Engine(m); g2.run_cycles(100); g2.synchronize()
    c2, _ = _checksum(g2)
    assert c1 == c2


def test_c4_tube_full_size_imposed_velocity():
    m = meshgen.crush_tube(708, 706, 1)
    assert m.numelc == 4 * 708 * 706 and m.numels == 708 * 708
    g = Engine(m); g.run_cycles(40); g.synchronize()
    d = g.download_nodes(("V", "D"))
    t = g.time()
    top = m.ibfv[:, 0] - 1
    want = -10.0 * min(1.0, (t["tt"] - 0.5 * t["dt2"]) / 0.05)
    assert np.allclose(d["V"][top, 2], want, rtol=1e-12)
    assert np.isfinite(d["V"]).all() and np.abs(d["V"][m.icodt == 1, 2]).max() == 0.0
