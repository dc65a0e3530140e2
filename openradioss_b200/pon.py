"""/PARITH/ON skyline tables, built the way the Starter does.

Slot order at a node (starter/source/spmd/domdec2.F FILLCNE :2138-2240): element types in the
order solids, quads, 4-node shells, ...; inside a type ascending USER element id
(IXS(11,.) / IXC(7,.)); inside an element the corner order K=1..8 (1..4).  IADS(K,I) /
IADC(K,I) hold the 1-based FSKY slot of each corner (starter/source/restart/ddsplit/w_pon.F
:617-631, get_fsky_address); ADSKY(N)..ADSKY(N+1)-1 are the slots of node N.
"""
import numpy as np


def build_pon(numnod: int, ixs: np.ndarray, ixc: np.ndarray, ixtg: np.ndarray = None):
    """Return (adsky[numnod+1], iads[numels,8], iadc[numelc,4], lsky), all 1-based int32; with `ixtg` (3-node shells,
    which FILLCNE places after the 4-node shells, sorted by IXTG(6,.): domdec2.F:2295-2310) also iadtg[numeltg,3]
    as a fifth value."""
    numels, numelc = ixs.shape[0], ixc.shape[0]
    numeltg = 0 if ixtg is None else ixtg.shape[0]
    nodes, owner = [], []
    if numels:
        order = np.argsort(ixs[:, 10], kind="stable")
        nodes.append(ixs[order, 1:9].reshape(-1))
        owner.append((order[:, None] * 8 + np.arange(8)[None, :]).reshape(-1))
    if numelc:
        order = np.argsort(ixc[:, 6], kind="stable")
        nodes.append(ixc[order, 1:5].reshape(-1))
        owner.append((8 * numels + order[:, None] * 4 + np.arange(4)[None, :]).reshape(-1))
    if numeltg:
        order = np.argsort(ixtg[:, 5], kind="stable")
        nodes.append(ixtg[order, 1:4].reshape(-1))
        owner.append((8 * numels + 4 * numelc + order[:, None] * 3 + np.arange(3)[None, :]).reshape(-1))
    nodes = np.concatenate(nodes).astype(np.int64)
    owner = np.concatenate(owner).astype(np.int64)
    lsky = nodes.size
    perm = np.argsort(nodes, kind="stable")          # processing order kept inside a node
    slot_of_corner = np.empty(lsky, np.int64)
    slot_of_corner[owner[perm]] = np.arange(1, lsky + 1)
    counts = np.bincount(nodes - 1, minlength=numnod)
    adsky = np.ones(numnod + 1, np.int64)
    adsky[1:] = 1 + np.cumsum(counts)
    iads = slot_of_corner[:8 * numels].reshape(numels, 8).astype(np.int32)
    iadc = slot_of_corner[8 * numels:8 * numels + 4 * numelc].reshape(numelc, 4).astype(np.int32)
    if ixtg is not None:
        iadtg = slot_of_corner[8 * numels + 4 * numelc:].reshape(numeltg, 3).astype(np.int32)
        return adsky.astype(np.int32), iads, iadc, int(lsky), iadtg
    return adsky.astype(np.int32), iads, iadc, int(lsky)
