"""Domain decomposition (SPMD_EXCH2_A_PON semantics): the reference's own /PARITH/ON criterion is
that nodal sums are *bitwise identical* on any domain count (qa-tests `-type=pon`).
CPU-only: the oracle plays every domain; the N>1 path also runs as two gloo processes."""
import os
import sys
import numpy as np
import pytest
from openradioss_b200 import meshgen, domdec, spmd
from oracle.orc import Oracle


def models():
    yield "shell", meshgen.shell_plate(9, 7, 90.0, 70.0, pressure=20.0, vrand=5.0, user_id_perm=True)
    yield "brick", meshgen.hex_block(5, 4, 6, 1.0, 0.8, 1.2, v0=(0, 0, -100.0), vrand=5.0, fix_bottom_z=True, user_id_perm=True)
    yield "sh3n_mixed", meshgen.tri_plate(9, 7, 90.0, 70.0, quads="checker", pressure=20.0, vrand=5.0, user_id_perm=True)
    yield "brick_law36", meshgen.hex_block(5, 4, 6, 10.0, 8.0, 12.0, law=36, v0=(0, 0, -60.0), vrand=20.0, fix_bottom_z=True)
    m = meshgen.shell_plate(9, 7, 90.0, 70.0, pressure=20.0, vrand=5.0)      # gravity on all nodes + a ramped load on a node subset
    meshgen.add_gravity(m, 3, -9.81e-3)
    meshgen.add_gravity(m, 1, 0.05, nodes=np.nonzero(m.X[:, 0] > 40.0)[0], curve=([0.0, 1.0e-3, 1.0], [0.0, 1.0, 1.0]))
    yield "shell_gravity", m


@pytest.mark.parametrize("nproc", [2, 3, 4])
def test_decomposition_invariants(nproc):
    for name, m in models():
        doms = [domdec.decompose_strips(m, nproc, r) for r in range(nproc)]
        assert sum(d.model.numelc + d.model.numels + d.model.numeltg for d in doms) == m.numelc + m.numels + m.numeltg
        owned = np.zeros(m.numnod, int)
        for d in doms:
            owned[d.node_gid[d.owner]] += 1
            # local slot lists are copies of the global ones
            assert d.model.lsky == int((np.diff(m.adsky.astype(np.int64))[d.node_gid]).sum())
            for nb in d.neighbors:
                other = next(x for x in doms[nb.rank].neighbors if x.rank == d.rank)
                assert len(nb.send) == len(other.recv) and len(nb.recv) == len(other.send)
            # every local slot is either filled by a local corner or received exactly once
            filled = np.zeros(d.model.lsky, int)
            for a in (d.model.iads, d.model.iadc, d.model.iadtg):
                if a is not None and a.size:
                    filled[a.reshape(-1) - 1] += 1
            for nb in d.neighbors:
                filled[nb.recv] += 1
            assert (filled == 1).all(), name
        assert (owned == 1).all()


@pytest.mark.parametrize("nproc", [2, 4])
def test_domains_reproduce_single_domain_bitwise(nproc):
    for name, m in models():
        ref = Oracle(m)
        doms = [domdec.decompose_strips(m, nproc, r, axis=0) for r in range(nproc)]
        backs = [Oracle(d.model) for d in doms]
        ncyc = 25
        ref_acc = []
        # single-domain reference, phased the same way
        st = spmd.initial_state(m.control)

        class Solo:
            rank, neighbors = 0, []
        for c in range(ncyc):
            dt1 = st["dt2"]
            ref.forces_phase(dt1); ref.assemble()
            ref_acc.append(ref.download_nodes(("A", "AR", "STIFN")))
            dt2 = min(spmd.EP06, ref.time()["dt2t"], float(np.float32(1.1)) * st["dt2old"], st["dtmx"])
            ref.advance(0.5 * (dt1 + dt2), dt2); st["dt2"] = dt2; st["dt2old"] = dt2

        def check(c):
            for b, d in zip(backs, doms):
                a = b.download_nodes(("A", "AR", "STIFN"))
                for k in ("A", "AR", "STIFN"):
                    assert np.array_equal(a[k], ref_acc[c][k][d.node_gid]), (name, c, k, d.rank)
        spmd.run_local(backs, doms, ncyc, on_cycle=check)
        xr = ref.download_nodes(("X", "V"))
        for b, d in zip(backs, doms):
            x = b.download_nodes(("X", "V"))
            assert np.array_equal(x["X"], xr["X"][d.node_gid]) and np.array_equal(x["V"], xr["V"][d.node_gid])


def _gloo_model():
    m = meshgen.shell_plate(8, 6, 80.0, 60.0, pressure=20.0, vrand=5.0)
    meshgen.add_gravity(m, 3, -9.81e-3)                                      # loads follow their nodes into the ranks
    return m


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = _gloo_model()
    d = domdec.decompose_strips(m, world, rank)
    b = Oracle(d.model)
    comm = spmd.TorchComm(dist)
    st = spmd.initial_state(m.control)
    for _ in range(20):
        st = spmd.cycle(b, d, comm, st)
    q.put((rank, d.node_gid, b.download_nodes(("X",))["X"], st["tt"]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gloo_processes_match_single_domain():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    m = _gloo_model()
    ref = Oracle(m); ref.run_cycles(20)
    xr = ref.download_nodes(("X",))["X"]
    for rank, gid, x, tt in res:
        assert np.array_equal(x, xr[gid])
        assert tt == ref.time()["tt"]


# ---- /PARITH/OFF: SPMD_EXCH_A, partial sums of the frontier nodes -------------------------------------------------------------
@pytest.mark.parametrize("nproc", [2, 3])
def test_parith_off_exchange_reproduces_the_single_domain_run_to_rounding(nproc):
    """Every domain assembles its own elements, the frontier sums travel and are added in rank order (spmd_exch_a.F:153-166,
    517-528): the nodal forces equal the single-domain ones up to the order of the additions -- 1e-12, not bitwise."""
    some_differ = False
    for name, m in models():
        ref = Oracle(m)
        doms = [domdec.parith_off(domdec.decompose_strips(m, nproc, r)) for r in range(nproc)]
        backs = [Oracle(d.model) for d in doms]
        st = spmd.initial_state(m.control)
        ref_acc = []
        for c in range(15):
            dt1 = st["dt2"]
            ref.forces_phase(dt1); ref.assemble()
            ref_acc.append(ref.download_nodes(("A", "AR", "STIFN")))
            dt2 = min(spmd.EP06, ref.time()["dt2t"], float(np.float32(1.1)) * st["dt2old"], st["dtmx"])
            ref.advance(0.5 * (dt1 + dt2), dt2); st["dt2"] = dt2; st["dt2old"] = dt2

        def check(c):
            nonlocal some_differ
            for b, d in zip(backs, doms):
                a = b.download_nodes(("A", "AR", "STIFN"))
                for k in ("A", "AR", "STIFN"):
                    want = ref_acc[c][k][d.node_gid]
                    scale = max(np.abs(ref_acc[c][k]).max(), 1e-300)
                    assert np.abs(a[k] - want).max() <= 1e-11 * scale, (name, c, k, d.rank)
                    some_differ = some_differ or not np.array_equal(a[k], want)
        spmd.run_local_off(backs, doms, 15, on_cycle=check)
        xr = ref.download_nodes(("X",))["X"]
        for b, d in zip(backs, doms):
            x = b.download_nodes(("X",))["X"]
            assert np.abs(x - xr[d.node_gid]).max() <= 1e-11 * np.abs(xr).max(), (name, d.rank)
    assert some_differ          # the sum order really is another one: this is not the /PARITH/ON path under another name


def test_parith_off_partial_sums_add_up_at_a_frontier_node():
    m = meshgen.shell_plate(6, 4, 60.0, 40.0, pressure=10.0, vrand=5.0)
    doms = [domdec.parith_off(domdec.decompose_strips(m, 2, r)) for r in range(2)]
    backs = [Oracle(d.model) for d in doms]
    for b in backs:
        b.forces_phase(0.0); b.assemble()
    part = [b.download_nodes(("A",))["A"] for b in backs]
    nb0 = doms[0].neighbors[0]; nb1 = doms[1].neighbors[0]
    assert np.array_equal(doms[0].node_gid[nb0.nodes], doms[1].node_gid[nb1.nodes])      # same nodes, same order on both sides
    buf01 = backs[0].pack_nodes(nb0.nodes); buf10 = backs[1].pack_nodes(nb1.nodes)
    backs[0].add_nodes(nb0.nodes, buf10); backs[1].add_nodes(nb1.nodes, buf01)
    a0 = backs[0].download_nodes(("A",))["A"]; a1 = backs[1].download_nodes(("A",))["A"]
    assert np.array_equal(a0[nb0.nodes], part[0][nb0.nodes] + part[1][nb1.nodes])
    assert np.array_equal(a1[nb1.nodes], part[1][nb1.nodes] + part[0][nb0.nodes])
    # the external load of a frontier node sits on one replica only
    f0, f1 = doms[0].model.fext[nb0.nodes], doms[1].model.fext[nb1.nodes]
    assert np.abs(f0).max() > 0.0 and np.abs(f1).max() == 0.0


def _gloo_worker_off(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = _gloo_model()
    d = domdec.parith_off(domdec.decompose_strips(m, world, rank))
    b = Oracle(d.model)
    comm = spmd.TorchComm(dist)
    st = spmd.initial_state(m.control)
    for _ in range(20):
        st = spmd.cycle_off(b, d, comm, st)
    q.put((rank, d.node_gid, b.download_nodes(("X",))["X"], st["tt"]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gloo_processes_parith_off():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker_off, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    m = _gloo_model()
    ref = Oracle(m); ref.run_cycles(20)
    xr = ref.download_nodes(("X",))["X"]
    for rank, gid, x, tt in res:
        assert np.abs(x - xr[gid]).max() <= 1e-11 * np.abs(xr).max()
        assert abs(tt - ref.time()["tt"]) <= 1e-12 * ref.time()["tt"]
