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
