"""orgpu_forces_host: the reference -gpu path's cycle (X, V, VR up -> internal forces -> assembled nodal forces down,
shell_internal_forces.F90:106-190) as ONE pipelined call.  Its result must be bit-identical to the phased calls
(orgpu_forces_phase + orgpu_assemble) on a twin engine, for any chunk count and any node numbering, and the element state
must advance identically; against the oracle the forces hold the 1e-12 of the phased path."""
import os
import numpy as np
import pytest
import torch
from conftest import rel_err
from openradioss_b200 import meshgen

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from openradioss_b200.engine import Engine
    from oracle.orc import Oracle


def _pinned(a):
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return t, t.numpy()


def _drive(m, cycles=4, chunks=None, fields=("sig", "pla", "eint")):
    if chunks is not None:
        os.environ["ORGPU_PIPE_CHUNKS"] = str(chunks)
    try:
        p, q = Engine(m), Engine(m)            # pipelined call / phased calls
        n = m.numnod
        keep = []
        F8t, F8 = _pinned(np.zeros((n, 8)))
        dt1 = 0.0
        for c in range(cycles):
            nd = q.download_nodes(("X", "V", "VR"))
            hx = [_pinned(nd[k]) for k in ("X", "V", "VR")]; keep.append(hx)
            F8[:] = -7.0
            dt2t, nel, ityp = p.forces_host(hx[0][1], hx[1][1], hx[2][1] if m.control.iroddl else None, dt1, F8)
            q.forces_phase(dt1); q.assemble()
            a = q.download_nodes(("A", "AR", "STIFN", "STIFR")); tq = q.time()
            # external loads are the caller's in forces_host: compare on models without them, or subtract nothing
            assert np.array_equal(F8[:, 0:3], a["A"]), c
            if m.control.iroddl:
                assert np.array_equal(F8[:, 3:6], a["AR"]) and np.array_equal(F8[:, 7], a["STIFR"]), c
            assert np.array_equal(F8[:, 6], a["STIFN"]), c
            assert dt2t == tq["dt2t"] and nel == tq["neltst"] and ityp == tq["ityptst"]
            dt2 = tq["dt2t"]
            q.advance(0.5 * (dt1 + dt2), dt2)
            dt1 = dt2
        return p, q
    finally:
        os.environ.pop("ORGPU_PIPE_CHUNKS", None)


@pytest.mark.parametrize("chunks", [1, 3, 12, 64])
def test_plate_pipelined_forces_are_bitwise_the_phased_ones(chunks):
    m = meshgen.shell_plate(40, 33, 400.0, 330.0, pulse_tau=0.0, vwave=(150.0, 100.0))
    m.fext = None; m.mext = None
    p, q = _drive(m, 4, chunks)
    for f in ("sig", "pla", "eint", "thk", "hourg"):
        assert np.array_equal(p.shell_state(f), q.shell_state(f)), f


def test_random_node_numbering_still_exact():
    """No locality at all: every tile waits for the last upload chunk, every node chunk for the last element group."""
    m = meshgen.shell_plate(24, 20, 240.0, 200.0, vwave=(150.0, 100.0), user_id_perm=True)
    m.fext = None; m.mext = None
    rng = np.random.default_rng(5)
    new = rng.permutation(m.numnod)                       # new index of old node i
    old = np.argsort(new)
    for k in ("X", "V", "VR", "MS", "IN", "icodt", "icodr", "itab"):
        a = getattr(m, k)
        if a is not None:
            setattr(m, k, np.ascontiguousarray(a[old]))
    m.ixc = m.ixc.copy(); m.ixc[:, 1:5] = new[m.ixc[:, 1:5] - 1] + 1
    from openradioss_b200.pon import build_pon
    m.adsky, m.iads, m.iadc, m.lsky = build_pon(m.numnod, m.ixs, m.ixc)
    _drive(m, 3, 8)


def test_bricks_and_mixed_models():
    m = meshgen.hex_block(12, 10, 9, 12.0, 10.0, 9.0, vrand=1.0, vseed=3)
    p, q = _drive(m, 3, 5)
    for f in ("sig", "eint", "rho"):
        assert np.array_equal(p.solid_state(f), q.solid_state(f)), f
    m = meshgen.shell_on_block(12, 8, 3)
    m.fext = None; m.mext = None
    _drive(m, 3, 7)
    m = meshgen.tri_plate(12, 10, 120.0, 100.0, quads="checker", vrand=5.0)      # 3-node shells: no table-driven variant, launched whole
    m.fext = None; m.mext = None
    _drive(m, 3, 6)


def test_forces_match_the_oracle():
    m = meshgen.shell_plate(30, 30, 300.0, 300.0, vwave=(150.0, 100.0))
    m.fext = None; m.mext = None
    g, o = Engine(m), Oracle(m)
    n = m.numnod
    F8t, F8 = _pinned(np.zeros((n, 8)))
    dt1 = 0.0
    for c in range(4):
        nd = o.download_nodes(("X", "V", "VR"))
        hx = [_pinned(nd[k]) for k in ("X", "V", "VR")]
        dt2t, nel, ityp = g.forces_host(hx[0][1], hx[1][1], hx[2][1], dt1, F8)
        o.forces_phase(dt1); o.assemble()
        a = o.download_nodes(("A", "AR", "STIFN")); to = o.time()
        assert rel_err(F8[:, 0:3], a["A"]) <= 1e-12 and rel_err(F8[:, 3:6], a["AR"]) <= 1e-12 and rel_err(F8[:, 6], a["STIFN"]) <= 1e-12
        assert dt2t == pytest.approx(to["dt2t"], rel=1e-13) and nel == to["neltst"]
        dt2 = to["dt2t"]
        o.advance(0.5 * (dt1 + dt2), dt2)
        dt1 = dt2


def test_rejected_outside_its_envelope():
    m = meshgen.shell_plate(8, 8, 80.0, 80.0)
    m.control.nodadt = 1
    g = Engine(m)
    n = m.numnod
    with pytest.raises(RuntimeError):
        g.forces_host(np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3)), 0.0, np.zeros((n, 8)))
