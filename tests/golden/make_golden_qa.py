"""Golden vectors held by the reference's own QA suite: the Engine listings (`*_0001.out`) that
`qa-tests/miniqa` keeps beside each deck as the expected run.  This script reads them where they lie under
/root/reference (it is run in the build container; the GPU box only sees the committed .npz files) and writes, per
deck, the per-cycle table the Engine printed:

    cycle, time, dt, ienergy, kenergy_t, kenergy_r, extwork, and the (type, id) that controls the time step.

    python tests/golden/make_golden_qa.py            # rewrites tests/golden/qa_*.npz
"""
import os
import re
import numpy as np

REF = "/root/reference/qa-tests/miniqa"
HERE = os.path.dirname(os.path.abspath(__file__))

DECKS = {
    # QEPH shells (Ishell=24), LAW36 with five rate curves + strain-rate filter, IMPVEL, BCS; two elements
    "elem_samp": "RUPTURE/FAIL_TAB/ELEM_SAMP/reference/1ELEM_SAMP_0001.out",
    # 3-node shells (C3FORC3) + LAW2, initial velocities
    "ct3a": "COQUES3N/ct3a/reference/CT3AV4_0001.out",
    # one brick, LAW36, /INIBRI/STRESS, nodal time step (only the cycles before the self-contact acts are used)
    "inibri_stress": "SOLIDES/inibri_stress/reference/TEST_002_0001.out",
    # twisted beam: shells in bending / warping, elastic, concentrated loads
    "twisbeam": "SMOKE_TEST/reference/TWISBEAM_0001.out",
    # 464 bricks LAW2 crushed by an imposed velocity (+ 48 LAW13 bricks whose nodes are all prescribed)
    "loi13_solide": "LOIS/LOI13/solide/reference/MODELE_0001.out",
}

_num = r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[EeDd][-+]?\d+)?"
_line = re.compile(r"^\s*(\d+)\s+(" + _num + r")\s+(" + _num + r")\s+(\w+)\s+(\d+)\s+(-?\d+\.\d+)%\s+(" + _num + r")\s+(" + _num +
                   r")\s+(" + _num + r")\s+(" + _num + r")\s+(" + _num + r")\s+(" + _num + r")\s+(" + _num + r")")


def parse_listing(path):
    rows, typ = [], []
    with open(path, errors="ignore") as f:
        for ln in f:
            m = _line.match(ln)
            if not m:
                continue
            g = m.groups()
            rows.append([float(g[0]), float(g[1]), float(g[2]), float(g[4]), float(g[5]), float(g[6]), float(g[7]),
                         float(g[8]), float(g[9]), float(g[11])])
            typ.append(g[3])
    a = np.array(rows)
    return dict(cycle=a[:, 0].astype(np.int64), time=a[:, 1], dt=a[:, 2], elid=a[:, 3].astype(np.int64), err_pct=a[:, 4],
                ienergy=a[:, 5], kenergy_t=a[:, 6], kenergy_r=a[:, 7], extwork=a[:, 8], mass=a[:, 9], eltype=np.array(typ))


def export_loi13_deck():
    """Nodes, bricks (with their part) and node groups of qa-tests/miniqa/LOIS/LOI13/solide/data/MODELE_0000.rad: what
    tests/qa_decks.loi13_solide() builds its model from (732 nodes, 464 + 48 bricks)."""
    p = os.path.join(REF, "LOIS/LOI13/solide/data/MODELE_0000.rad")
    if not os.path.exists(p):
        print("missing", p); return
    lines = [l.rstrip("\n") for l in open(p, errors="ignore")]

    def block(key):
        out, on = [], False
        for l in lines:
            if l.startswith("/"):
                on = (l.strip() == key); continue
            if on and not l.startswith("#") and l.strip():
                out.append(l)
        return out
    nodes = [(int(l[:10]), float(l[10:30]), float(l[30:50]), float(l[50:70])) for l in block("/NODE")]
    bricks, part = [], []
    for k in (1, 2, 4, 5):
        for l in block(f"/BRICK/{k}"):
            bricks.append([int(l[i * 10:(i + 1) * 10]) for i in range(9)]); part.append(k)
    grnod = lambda k: np.array([int(x) for l in block(f"/GRNOD/NODE/{k}")[1:] for x in l.split()], np.int32)
    np.savez_compressed(os.path.join(HERE, "qa_loi13_deck.npz"), node_id=np.array([n[0] for n in nodes], np.int32),
                        X=np.array([n[1:] for n in nodes]), brick=np.array(bricks, np.int32), part=np.array(part, np.int32),
                        grnod2=grnod(2), grnod3=grnod(3))
    print(f"loi13 deck: {len(nodes)} nodes, {len(bricks)} bricks")


if __name__ == "__main__":
    export_loi13_deck()
    for name, rel in DECKS.items():
        p = os.path.join(REF, rel)
        if not os.path.exists(p):
            print("missing", p); continue
        d = parse_listing(p)
        np.savez_compressed(os.path.join(HERE, f"qa_{name}.npz"), source=np.array(rel), **d)
        print(f"{name}: {len(d['cycle'])} listing lines, cycles {d['cycle'][0]}..{d['cycle'][-1]}")
