"""Host-staged SPMD cycle over a set of domains: the RESOL call sequence around the hot path when
the exchange stays on the host (the reference's own arrangement: MPI in SPMD_EXCH2_A_PON,
resol.F:4801, and SPMD_GLOB_MIN5, resol.F:6327).

    forces_phase(dt1) -> pack rows -> exchange -> unpack rows -> assemble -> global dt min -> advance

`Comm` backends: `LocalComm` runs all domains in one process (tests), `TorchComm` is one process
per domain over torch.distributed (gloo on CPU, NCCL on GPUs).  The device-resident multi-GPU
loop (orgpu_run_cycles with orgpu_comm_init) does the same on the GPU without host staging.
"""
from __future__ import annotations
import numpy as np

EP06 = 1.0e6


class TorchComm:
    """torch.distributed point-to-point exchange of corner rows + min-allreduce of dt."""

    def __init__(self, dist, device="cpu"):
        import torch
        self.dist, self.torch, self.device = dist, torch, device
        self.rank, self.size = dist.get_rank(), dist.get_world_size()

    def exchange(self, sends):
        """sends: {rank: ndarray rows}; returns {rank: ndarray rows} (sizes agreed through the domain lists)."""
        t = self.torch
        reqs, recvs = [], {}
        for q, (buf, nrecv) in sends.items():
            out = t.from_numpy(np.ascontiguousarray(buf)).to(self.device)
            inp = t.empty((nrecv, 8), dtype=t.float64, device=self.device)
            recvs[q] = inp
            if len(buf):
                reqs.append(self.dist.isend(out, q))
            if nrecv:
                reqs.append(self.dist.irecv(inp, q))
        for r in reqs:
            r.wait()
        return {q: v.cpu().numpy() for q, v in recvs.items()}

    def min_dt(self, dt, ityp, ngl):
        t = self.torch
        v = t.tensor([dt], dtype=t.float64, device=self.device)
        self.dist.all_reduce(v, op=self.dist.ReduceOp.MIN)
        return float(v.item())


def cycle(backend, dom, comm, state):
    """One explicit cycle of one domain (reference order: resol.F:2721-2722, 4138-4225, 4801, 4858,
    6124-6128, 6327, 6352, 6494-6497, 6921-9042)."""
    dt1 = state["dt2"]
    backend.forces_phase(dt1)
    sends = {nb.rank: (backend.pack_rows(nb.send), len(nb.recv)) for nb in dom.neighbors}
    got = comm.exchange(sends)
    for nb in dom.neighbors:
        backend.unpack_rows(nb.recv, got[nb.rank])
    backend.assemble()
    t = backend.time()
    dt2 = min(EP06, t["dt2t"])                       # DT2 = EP06 ; IF (DT2T < DT2) DT2 = DT2T
    dt2 = comm.min_dt(dt2, t["ityptst"], t["neltst"])
    dt2 = min(dt2, float(np.float32(1.1)) * state["dt2old"], state["dtmx"])
    state["dt2old"] = dt2
    dt12 = 0.5 * (dt1 + dt2)
    backend.advance(dt12, dt2)
    state["dt2"] = dt2
    state["tt"] = state.get("tt", 0.0) + dt2
    return state


def initial_state(control):
    return {"dt2": control.dt_init, "dt2old": control.dt2old_init, "dtmx": control.dtmx, "tt": control.tt_init}


class LocalComm:
    """All domains in one process: `run_local` steps them in lock-step and moves rows between them."""
    pass


def run_local(backends, doms, ncycles, on_cycle=None):
    states = [initial_state(d.model.control) for d in doms]
    for c in range(ncycles):
        dt1 = states[0]["dt2"]
        for b in backends:
            b.forces_phase(dt1)
        packed = {(d.rank, nb.rank): b.pack_rows(nb.send) for b, d in zip(backends, doms) for nb in d.neighbors}
        for b, d in zip(backends, doms):
            for nb in d.neighbors:
                buf = packed[(nb.rank, d.rank)]
                assert len(buf) == len(nb.recv), "send/recv lists of a domain pair disagree"
                b.unpack_rows(nb.recv, buf)
        for b in backends:
            b.assemble()
        if on_cycle:
            on_cycle(c)
        dt2 = min([EP06] + [b.time()["dt2t"] for b in backends])
        dt2 = min(dt2, float(np.float32(1.1)) * states[0]["dt2old"], states[0]["dtmx"])
        dt12 = 0.5 * (dt1 + dt2)
        for b, s in zip(backends, states):
            b.advance(dt12, dt2)
            s["dt2old"] = dt2; s["dt2"] = dt2
    return states


def run_local_off(backends, doms, ncycles, on_cycle=None):
    """/PARITH/OFF in lock-step: every domain assembles the corner rows of its OWN elements (the reserved remote slots of its
    skyline stay zero), then the partial sums of the frontier nodes travel -- SPMD_EXCH_A (spmd_exch_a.F:153-166 pack,
    :517-528 add): 8 values per node, added neighbour by neighbour in rank order.  `doms` must come from domdec.parith_off."""
    states = [initial_state(d.model.control) for d in doms]
    for b in backends:
        b.set_parith(0)                       # /PARITH/OFF: FORCE adds to A before the element loop (force.F90:182-312)
    for c in range(ncycles):
        dt1 = states[0]["dt2"]
        for b in backends:
            b.forces_phase(dt1); b.assemble()
        packed = {(d.rank, nb.rank): b.pack_nodes(nb.nodes) for b, d in zip(backends, doms) for nb in d.neighbors}
        for b, d in zip(backends, doms):
            for nb in sorted(d.neighbors, key=lambda n: n.rank):
                b.add_nodes(nb.nodes, packed[(nb.rank, d.rank)])
        if on_cycle:
            on_cycle(c)
        dt2 = min([EP06] + [b.time()["dt2t"] for b in backends])
        dt2 = min(dt2, float(np.float32(1.1)) * states[0]["dt2old"], states[0]["dtmx"])
        dt12 = 0.5 * (dt1 + dt2)
        for b, s in zip(backends, states):
            b.advance(dt12, dt2)
            s["dt2old"] = dt2; s["dt2"] = dt2
    return states


def cycle_off(backend, dom, comm, state):
    """One /PARITH/OFF cycle of one domain in its own process (SPMD_EXCH_A + SPMD_GLOB_MIN5 over `comm`)."""
    dt1 = state["dt2"]
    if not state.get("iparit0"):
        backend.set_parith(0); state["iparit0"] = True      # /PARITH/OFF: the nodal sum starts from the load
    backend.forces_phase(dt1); backend.assemble()
    sends = {nb.rank: (backend.pack_nodes(nb.nodes), len(nb.nodes)) for nb in dom.neighbors}
    got = comm.exchange(sends)
    for nb in sorted(dom.neighbors, key=lambda n: n.rank):
        backend.add_nodes(nb.nodes, got[nb.rank])
    t = backend.time()
    dt2 = comm.min_dt(min(EP06, t["dt2t"]), t["ityptst"], t["neltst"])
    dt2 = min(dt2, float(np.float32(1.1)) * state["dt2old"], state["dtmx"])
    state["dt2old"] = dt2
    backend.advance(0.5 * (dt1 + dt2), dt2)
    state["dt2"] = dt2; state["tt"] = state.get("tt", 0.0) + dt2
    return state
