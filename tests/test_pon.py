"""/PARITH/ON tables: slot order = element type order then ascending user id (domdec2.F FILLCNE)."""
import numpy as np
from openradioss_b200 import meshgen
from openradioss_b200.pon import build_pon


def _ref_pon(numnod, ixs, ixc):
    """Literal loop restatement of FILLCNE (starter/source/spmd/domdec2.F:2138-2240)."""
    cnt = np.zeros(numnod + 2, int)
    for e in ixs: 
        for n in e[1:9]: cnt[n] += 1
    for e in ixc:
        for n in e[1:5]: cnt[n] += 1
    adsky = np.ones(numnod + 2, int)
    for n in range(1, numnod + 1): adsky[n + 1] = adsky[n] + cnt[n]
    cur = adsky.copy()
    iads = np.zeros((len(ixs), 8), int); iadc = np.zeros((len(ixc), 4), int)
    for i in np.argsort(ixs[:, 10], kind="stable") if len(ixs) else []:
        for k in range(8):
            n = ixs[i, 1 + k]; iads[i, k] = cur[n]; cur[n] += 1
    for i in np.argsort(ixc[:, 6], kind="stable") if len(ixc) else []:
        for k in range(4):
            n = ixc[i, 1 + k]; iadc[i, k] = cur[n]; cur[n] += 1
    return adsky[1:numnod + 2], iads, iadc


def test_pon_matches_loop_restatement_with_permuted_user_ids():
    m = meshgen.hex_block(3, 4, 2, 1.0, 1.0, 1.0, user_id_perm=True)
    adsky, iads, iadc = _ref_pon(m.numnod, m.ixs, m.ixc)
    assert np.array_equal(adsky, m.adsky) and np.array_equal(iads, m.iads)
    assert m.lsky == 8 * m.numels
    # every slot used exactly once
    assert np.array_equal(np.sort(m.iads.reshape(-1)), np.arange(1, m.lsky + 1))


def test_slots_ascend_with_user_id_at_each_node():
    m = meshgen.hex_block(3, 3, 3, 1.0, 1.0, 1.0, user_id_perm=True)
    for n in range(1, m.numnod + 1):
        slots = np.arange(m.adsky[n - 1], m.adsky[n])
        owners = []
        for s in slots:
            e, k = np.argwhere(m.iads == s)[0]
            assert m.ixs[e, 1 + k] == n
            owners.append(m.ixs[e, 10])
        assert owners == sorted(owners)
