"""Domain decomposition of a model for one-process-per-GPU runs, the way the Starter does it.

Reference behaviour: elements are owned by exactly one domain
(starter/source/spmd/domdec2.F); frontier nodes are duplicated on every domain that touches
them **with identical full slot lists** (the local ADDCNE is copied from the global one,
starter/source/restart/ddsplit/w_pon.F:617-631) and PROCNE records the owner of each slot.  At run
time SPMD_EXCH2_A_PON (engine/source/mpi/forces/spmd_exch2_a_pon.F:545-557 pack from
FSKY(:,ISENDP), :1190-1201 unpack into FSKY(:,IRECVP)) ships the corner rows of remote elements
into their reserved slots, then the same ordered ASSPAR4 gather runs on every replica, so nodal
sums are bitwise identical on any domain count.

`decompose` builds, for one rank, the local Model (local node / element / slot numbering) plus the
per-neighbour send and receive slot lists (ISENDP / IRECVP equivalents, 0-based local slots, both
sides ordered by ascending global slot so the message layouts agree without negotiation).
"""
from __future__ import annotations
from dataclasses import dataclass, field
from typing import List, Optional
import numpy as np
from .model import Model, SolidGroup, ShellGroup, NVSIZ


@dataclass
class Neighbor:
    rank: int
    send: np.ndarray          # local 0-based FSKY slots whose rows go to `rank`   (ISENDP)
    recv: np.ndarray          # local 0-based FSKY slots filled by rows from `rank` (IRECVP)
    nodes: np.ndarray = None  # local 0-based nodes shared with `rank`, ascending global id on both sides (FR_ELEM / IAD_ELEM of
                              # SPMD_EXCH_A: the /PARITH/OFF exchange of nodal partial sums)


@dataclass
class Domain:
    rank: int
    nproc: int
    model: Model
    node_gid: np.ndarray      # local node -> global node (0-based)
    shell_gid: np.ndarray     # local shell -> global shell (0-based)
    solid_gid: np.ndarray
    sh3n_gid: np.ndarray = None   # local 3-node shell -> global (0-based)
    owner: np.ndarray = None        # bool per local node: this rank is the lowest rank holding it (WEIGHT)
    neighbors: List[Neighbor] = field(default_factory=list)


def strips(centroids: np.ndarray, nproc: int, axis: int = 0) -> np.ndarray:
    """Balanced contiguous strips along one axis (C3: x strips, C4/C5: z slabs): element -> domain."""
    order = np.argsort(centroids[:, axis], kind="stable")
    dom = np.empty(len(order), np.int32)
    dom[order] = (np.arange(len(order), dtype=np.int64) * nproc // max(1, len(order))).astype(np.int32)
    return dom


def element_centroids(m: Model):
    cs = m.X[m.ixs[:, 1:9] - 1].mean(axis=1) if m.numels else np.zeros((0, 3))
    cc = m.X[m.ixc[:, 1:5] - 1].mean(axis=1) if m.numelc else np.zeros((0, 3))
    return cs, cc


def sh3n_centroids(m: Model):
    return m.X[m.ixtg[:, 1:4] - 1].mean(axis=1) if m.numeltg else np.zeros((0, 3))


def _regroup(groups, gid_of_local, cls):
    """local element groups: runs of consecutive local elements that came from one global group."""
    out = []
    if len(gid_of_local) == 0:
        return out
    start = 0
    for i in range(1, len(gid_of_local) + 1):
        if i == len(gid_of_local) or gid_of_local[i] != gid_of_local[start] or i - start == NVSIZ:
            g = groups[gid_of_local[start]]
            if cls is SolidGroup:
                out.append(SolidGroup(nft=start, nel=i - start, mat=g.mat, prop=g.prop, law=getattr(g, "law", 2), fail=getattr(g, "fail", None)))
            else:
                out.append(ShellGroup(nft=start, nel=i - start, law=g.law, mat=g.mat, prop=g.prop, fail=getattr(g, "fail", None)))
            start = i
    return out


def decompose(m: Model, dom_s: Optional[np.ndarray], dom_c: Optional[np.ndarray], nproc: int, rank: int,
              dom_t: Optional[np.ndarray] = None) -> Domain:
    """dom_s / dom_c / dom_t: domain of every brick / 4-node shell / 3-node shell."""
    numnod = m.numnod
    dom_s = np.zeros(0, np.int32) if dom_s is None else np.asarray(dom_s, np.int32)
    dom_c = np.zeros(0, np.int32) if dom_c is None else np.asarray(dom_c, np.int32)
    dom_t = np.zeros(0, np.int32) if dom_t is None else np.asarray(dom_t, np.int32)
    assert len(dom_t) == m.numeltg, "dom_t must give the domain of every 3-node shell"
    adsky0 = m.adsky.astype(np.int64) - 1                       # 0-based global slot offsets per node
    # owner domain of every global slot (PROCNE)
    slot_dom = np.full(m.lsky, -1, np.int32)
    if m.numels:
        slot_dom[m.iads.reshape(-1).astype(np.int64) - 1] = np.repeat(dom_s, 8)
    if m.numelc:
        slot_dom[m.iadc.reshape(-1).astype(np.int64) - 1] = np.repeat(dom_c, 4)
    if m.numeltg:
        slot_dom[m.iadtg.reshape(-1).astype(np.int64) - 1] = np.repeat(dom_t, 3)
    assert (slot_dom >= 0).all(), "every FSKY slot must belong to an element corner"
    slot_node = np.repeat(np.arange(numnod, dtype=np.int64), np.diff(adsky0))
    # node sets per domain
    masks = np.zeros((nproc, numnod), bool)
    for q in range(nproc):
        if m.numels:
            masks[q, (m.ixs[dom_s == q, 1:9] - 1).reshape(-1)] = True
        if m.numelc:
            masks[q, (m.ixc[dom_c == q, 1:5] - 1).reshape(-1)] = True
        if m.numeltg:
            masks[q, (m.ixtg[dom_t == q, 1:4] - 1).reshape(-1)] = True
    mine = masks[rank]
    node_gid = np.nonzero(mine)[0]
    g2l = np.full(numnod, -1, np.int64); g2l[node_gid] = np.arange(len(node_gid))
    counts = (adsky0[1:] - adsky0[:-1])[node_gid]
    ladsky0 = np.zeros(len(node_gid) + 1, np.int64); ladsky0[1:] = np.cumsum(counts)
    lsky = int(ladsky0[-1])
    # global slot -> local slot for slots of local nodes
    gslots = np.nonzero(mine[slot_node])[0]                    # ascending global slot == local slot order
    gs2l = np.full(m.lsky, -1, np.int64); gs2l[gslots] = np.arange(lsky)
    solid_gid = np.nonzero(dom_s == rank)[0]; shell_gid = np.nonzero(dom_c == rank)[0]
    ixs = m.ixs[solid_gid].copy(); ixc = m.ixc[shell_gid].copy()
    if len(solid_gid):
        ixs[:, 1:9] = (g2l[ixs[:, 1:9] - 1] + 1).astype(np.int32)
    if len(shell_gid):
        ixc[:, 1:5] = (g2l[ixc[:, 1:5] - 1] + 1).astype(np.int32)
    iads = (gs2l[m.iads[solid_gid].astype(np.int64) - 1] + 1).astype(np.int32) if len(solid_gid) else np.zeros((0, 8), np.int32)
    iadc = (gs2l[m.iadc[shell_gid].astype(np.int64) - 1] + 1).astype(np.int32) if len(shell_gid) else np.zeros((0, 4), np.int32)
    sh3n_gid = np.nonzero(dom_t == rank)[0]
    ixtg = m.ixtg[sh3n_gid].copy()
    if len(sh3n_gid):
        ixtg[:, 1:4] = (g2l[ixtg[:, 1:4] - 1] + 1).astype(np.int32)
    iadtg = (gs2l[m.iadtg[sh3n_gid].astype(np.int64) - 1] + 1).astype(np.int32) if len(sh3n_gid) else np.zeros((0, 3), np.int32)
    sub = lambda a: None if a is None else np.ascontiguousarray(a[node_gid])
    lm = Model(X=sub(m.X), V=sub(m.V), VR=sub(m.VR), MS=sub(m.MS), IN=sub(m.IN), control=m.control, ixs=ixs, ixc=ixc, ixtg=ixtg,
               vol0=m.vol0[solid_gid] if len(m.vol0) else m.vol0, icodt=sub(m.icodt), icodr=sub(m.icodr),
               fext=sub(m.fext), mext=sub(m.mext), itab=sub(m.itab), npf=m.npf, tf=m.tf, load_func=m.load_func)
    if m.cload_ib is not None and len(m.cload_ib):               # load records follow their node into every domain that holds it (record order kept)
        keep = mine[m.cload_ib[:, 0] - 1]
        lm.cload_ib = m.cload_ib[keep].copy(); lm.cload_fac = m.cload_fac[keep].copy()
        lm.cload_ib[:, 0] = (g2l[lm.cload_ib[:, 0] - 1] + 1).astype(np.int32)
    if m.igrv is not None and len(m.igrv):                       # gravity: every domain keeps the loads, with its own nodes of each list
        lm.igrv = m.igrv.copy(); lm.agrv = m.agrv.copy(); parts = []; iad = 0
        for l in range(len(m.igrv)):
            ib = np.abs(m.ibgrv[iad:iad + m.igrv[l, 0]]); iad += m.igrv[l, 0]
            loc = (g2l[ib[mine[ib - 1]] - 1] + 1).astype(np.int32)
            lm.igrv[l, 0] = len(loc); parts.append(loc)
        lm.ibgrv = np.concatenate(parts) if parts else np.zeros(0, np.int32)
    if m.ibfv is not None and len(m.ibfv):
        keep = mine[m.ibfv[:, 0] - 1]
        lm.ibfv = m.ibfv[keep].copy(); lm.vel = m.vel[keep].copy()
        lm.ibfv[:, 0] = (g2l[lm.ibfv[:, 0] - 1] + 1).astype(np.int32)
    # tie-break keys of the time-step arg-min: where each local element sits in the undecomposed model's processing order
    # (4-node shells, then 3-node shells, then solids), and the global node index
    lm.gorder = dict(shell=shell_gid.astype(np.int32), sh3n=(m.numelc + sh3n_gid).astype(np.int32),
                     solid=(m.numelc + m.numeltg + solid_gid).astype(np.int32), node=node_gid.astype(np.int32))
    # parts follow their elements
    if m.ipartc is not None: lm.ipartc = m.ipartc[shell_gid]
    if m.iparts is not None: lm.iparts = m.iparts[solid_gid]
    if m.iparttg is not None: lm.iparttg = m.iparttg[sh3n_gid]
    lm.adsky = (ladsky0 + 1).astype(np.int32); lm.iads = iads; lm.iadc = iadc; lm.iadtg = iadtg; lm.lsky = lsky
    # groups
    if m.solid_groups:
        gid = np.zeros(m.numels, np.int64)
        for k, g in enumerate(m.solid_groups):
            gid[g.nft:g.nft + g.nel] = k
        lm.solid_groups = _regroup(m.solid_groups, gid[solid_gid], SolidGroup)
    if m.shell_groups:
        gid = np.zeros(m.numelc, np.int64)
        for k, g in enumerate(m.shell_groups):
            gid[g.nft:g.nft + g.nel] = k
        lm.shell_groups = _regroup(m.shell_groups, gid[shell_gid], ShellGroup)
    if m.sh3n_groups:
        gid = np.zeros(m.numeltg, np.int64)
        for k, g in enumerate(m.sh3n_groups):
            gid[g.nft:g.nft + g.nel] = k
        lm.sh3n_groups = _regroup(m.sh3n_groups, gid[sh3n_gid], ShellGroup)
    # ownership (lowest rank holding the node) for global reductions
    first = np.argmax(masks, axis=0)
    owner = first[node_gid] == rank
    d = Domain(rank=rank, nproc=nproc, model=lm, node_gid=node_gid, shell_gid=shell_gid, solid_gid=solid_gid, sh3n_gid=sh3n_gid, owner=owner)
    # exchange lists
    ldom = slot_dom[gslots]                                     # owner of each local slot
    for q in range(nproc):
        if q == rank:
            continue
        recv = np.nonzero(ldom == q)[0]                         # rows rank q computes for my nodes
        theirs = np.nonzero((slot_dom == rank) & masks[q][slot_node])[0]   # my rows at nodes q also holds
        send = gs2l[theirs]
        shared = g2l[np.nonzero(mine & masks[q])[0]]            # ascending global node id
        if len(recv) or len(send) or len(shared):
            d.neighbors.append(Neighbor(rank=q, send=send.astype(np.int32), recv=recv.astype(np.int32), nodes=shared.astype(np.int32)))
    return d


def parith_off(d: Domain) -> Domain:
    """The same domain prepared for the /PARITH/OFF exchange (SPMD_EXCH_A: every domain assembles its own elements, the
    partial sums of the frontier nodes are exchanged and added): an external nodal load must then sit on ONE replica of a
    frontier node -- the Starter gives each load record to one domain -- here the lowest rank holding the node."""
    m = d.model
    for name in ("fext", "mext"):
        a = getattr(m, name)
        if a is not None:
            a = a.copy(); a[~d.owner] = 0.0; setattr(m, name, a)
    if m.cload_ib is not None and len(m.cload_ib):
        keep = d.owner[m.cload_ib[:, 0] - 1]
        m.cload_ib = m.cload_ib[keep].copy(); m.cload_fac = m.cload_fac[keep].copy()
    return d


def decompose_strips(m: Model, nproc: int, rank: int, axis: int = 0) -> Domain:
    cs, cc = element_centroids(m)
    ct = sh3n_centroids(m)
    both = np.concatenate([cs, cc, ct])
    dom = strips(both, nproc, axis)
    return decompose(m, dom[:len(cs)], dom[len(cs):len(cs) + len(cc)], nproc, rank, dom_t=dom[len(cs) + len(cc):])
