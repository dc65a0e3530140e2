"""Domain decomposition on the GPU path.

1 GPU: two Engine handles play two domains, rows travel through the host-staged C-ABI entry points
(orgpu_pack_rows / orgpu_unpack_rows) -- nodal sums must be bitwise those of the single-domain run.
>= 2 GPUs: one process per GPU, exchange + dt fold inside orgpu_run_cycles, once through the library's own
peer-memory kernels (CUDA IPC windows over NVLink, one CUDA graph per cycle) and once through NCCL (skipped on 1 GPU)."""
import os
import numpy as np
import pytest
import torch
from openradioss_b200 import meshgen, domdec, spmd

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from openradioss_b200.engine import Engine


def models():
    yield "shell", meshgen.shell_plate(12, 9, 120.0, 90.0, pressure=20.0, vrand=5.0, user_id_perm=True)
    yield "brick", meshgen.hex_block(6, 5, 7, 1.2, 1.0, 1.4, v0=(0, 0, -100.0), vrand=5.0, fix_bottom_z=True, user_id_perm=True)
    yield "tube", meshgen.crush_tube(5, 8, 1, ramp=0.002)      # shells + bricks, imposed velocity, load records follow their nodes
    t = meshgen.crush_tube(5, 8, 1, ramp=0.002); t.control.nodadt = 1
    yield "sh3n_mixed", meshgen.tri_plate(12, 9, 120.0, 90.0, quads="checker", pressure=20.0, vrand=5.0, user_id_perm=True)
    yield "brick_law36", meshgen.hex_block(6, 5, 7, 12.0, 10.0, 14.0, law=36, v0=(0, 0, -60.0), vrand=20.0, fix_bottom_z=True, user_id_perm=True)
    yield "tube_dtnoda", t                                     # /DT/NODA: the nodal dt crosses the domains after the assembly
    # perfectly regular meshes at rest: every element has the same time step, so the arg-min is decided by the tie rules alone
    # (shells: first in processing order; solids: last) -- N domains must elect the element one domain elects
    yield "shell_tie", meshgen.shell_plate(12, 9, 120.0, 90.0, pressure=20.0, jitter=0.0, zjitter=0.0)
    yield "brick_tie", meshgen.hex_block(6, 5, 7, 1.2, 1.0, 1.4, jitter=0.0, v0=(0, 0, -1.0))


@pytest.mark.parametrize("nproc", [2, 3])
def test_host_staged_domains_bitwise(nproc):
    for name, m in models():
        if m.control.nodadt:
            continue                                             # phased mode leaves the global min to the caller
        ref = Engine(m)
        doms = [domdec.decompose_strips(m, nproc, r) for r in range(nproc)]
        backs = [Engine(d.model) for d in doms]
        ncyc = 20
        st = spmd.initial_state(m.control)
        ref_acc = []
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


def _nccl_worker(rank, world, port, q, kind, p2p=True):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    m = dict(models())[kind]
    d = domdec.decompose_strips(m, world, rank)
    g = Engine(d.model, device=rank)
    g.comm_init(dist, d, p2p=p2p)
    g.run_cycles(25); g.synchronize()
    g.run_cycles(15); g.synchronize()          # a second call replays the captured graph / re-enters the exchange
    out = g.download_nodes(("X", "V"))
    q.put((rank, d.node_gid, out["X"], out["V"], g.time()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("kind,p2p", [("shell", True), ("shell", False), ("brick", True), ("brick", False), ("tube", True), ("tube", False),
                                      ("tube_dtnoda", True), ("sh3n_mixed", True), ("brick_law36", True),
                                      ("shell_tie", True), ("shell_tie", False), ("brick_tie", True), ("brick_tie", False)])
def test_multi_gpu_domains_match_single_gpu_bitwise(kind, p2p):
    import torch.multiprocessing as mp
    world = min(4, torch.cuda.device_count())
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, q, kind, p2p)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    m = dict(models())[kind]
    ref = Engine(m); ref.run_cycles(40); ref.synchronize()
    xr = ref.download_nodes(("X", "V")); tr = ref.time()
    for rank, gid, x, v, t in res:
        assert np.array_equal(x, xr["X"][gid]) and np.array_equal(v, xr["V"][gid]), (kind, rank)
        assert t["tt"] == tr["tt"] and t["ncycle"] == tr["ncycle"] and t["dt2"] == tr["dt2"]
        assert t["neltst"] == tr["neltst"] and t["ityptst"] == tr["ityptst"], (kind, rank, t, tr)     # also on exact ties


# ---- /PARITH/OFF (SPMD_EXCH_A) ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nproc", [2, 3])
def test_host_staged_parith_off_matches_the_oracle(nproc):
    """Two / three Engine handles on one GPU play the domains; the partial sums of the frontier nodes go through
    orgpu_pack_nodes / orgpu_add_nodes.  Against the oracle stepping the same domains the same way: 1e-12 on the assembled
    forces every cycle; against the single-domain run: rounding of another sum order."""
    from oracle.orc import Oracle
    for name, m in models():
        if m.control.nodadt:
            continue
        doms = [domdec.parith_off(domdec.decompose_strips(m, nproc, r)) for r in range(nproc)]
        gb = [Engine(d.model) for d in doms]; ob = [Oracle(d.model) for d in doms]
        acc = []
        so = spmd.run_local_off(ob, doms, 12, on_cycle=lambda c: acc.append([o.download_nodes(("A", "AR", "STIFN")) for o in ob]))

        def check(c):
            for g, ao in zip(gb, acc[c]):
                ag = g.download_nodes(("A", "AR", "STIFN"))
                for k in ("A", "AR", "STIFN"):
                    assert np.abs(ag[k] - ao[k]).max() <= 1e-12 * max(np.abs(ao[k]).max(), 1e-300), (name, c, k)
        sg = spmd.run_local_off(gb, doms, 12, on_cycle=check)
        ref = Engine(m); ref.run_cycles(12)
        xr = ref.download_nodes(("X",))["X"]
        for g, o, d in zip(gb, ob, doms):
            xg, xo = g.download_nodes(("X",))["X"], o.download_nodes(("X",))["X"]
            assert np.abs(xg - xo).max() <= 1e-12 * np.abs(xo).max(), (name, d.rank)
            assert np.abs(xg - xr[d.node_gid]).max() <= 1e-11 * np.abs(xr).max(), (name, d.rank)
        assert sg[0]["dt2"] == pytest.approx(so[0]["dt2"], rel=1e-12)


def _nccl_worker_off(rank, world, port, q, kind):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    m = dict(models())[kind]
    d = domdec.parith_off(domdec.decompose_strips(m, world, rank))
    g = Engine(d.model, device=rank)
    g.comm_init(dist, d, parith_off=True)
    g.run_cycles(25); g.synchronize(); g.run_cycles(15); g.synchronize()
    out = g.download_nodes(("X", "V"))
    q.put((rank, d.node_gid, out["X"], out["V"], g.time()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("kind", ["shell", "brick", "tube"])
def test_multi_gpu_parith_off_over_nccl(kind):
    import torch.multiprocessing as mp
    world = min(4, torch.cuda.device_count())
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31700 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_nccl_worker_off, args=(r, world, port, q, kind)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    m = dict(models())[kind]
    ref = Engine(m); ref.run_cycles(40); ref.synchronize()
    xr = ref.download_nodes(("X", "V")); tr = ref.time()
    for rank, gid, x, v, t in res:
        assert np.abs(x - xr["X"][gid]).max() <= 1e-10 * np.abs(xr["X"]).max(), (kind, rank)
        assert np.abs(v - xr["V"][gid]).max() <= 1e-9 * np.abs(xr["V"]).max(), (kind, rank)
        assert t["ncycle"] == tr["ncycle"] and t["tt"] == pytest.approx(tr["tt"], rel=1e-11)
