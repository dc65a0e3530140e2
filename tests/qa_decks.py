"""TEST INFRASTRUCTURE: hand-built Model equivalents of decks of the reference's QA suite (qa-tests/miniqa), i.e. what
the Starter would hand the Engine for them.  The numbers (nodes, connectivity, cards) are the decks'; each builder
cites the deck it restates.  The reference's expected listings for the same decks are committed as
tests/golden/qa_*.npz (tests/golden/make_golden_qa.py), and tests/test_qa_decks*.py hold the oracle and the CUDA path
to them -- the pin of the restatement on numbers the reference Engine itself produced.
"""
import numpy as np
from openradioss_b200 import meshgen
from openradioss_b200.model import Model, Law36, PropShell, ShellGroup
from openradioss_b200.pon import build_pon
from openradioss_b200.constants import K


def law36(rho0, young, nu, curves, rates, yfac, fcut=0.0, ismooth=1):
    """/MAT/PLAS_TAB as hm_read_mat36.F:198-342 stores it: `curves` = one (x, y) pair per rate AFTER the Starter's
    insertion of the zero rate (:207-216).  Returns (Law36, npf, tf)."""
    m = Law36()
    m.rho0 = rho0; m.young = young; m.nu = nu
    m.shear = 0.5 * young / (1.0 + nu)
    m.bulk = young / 3.0 / (1.0 - 2.0 * nu)
    m.a1u = young / (1.0 - nu * nu); m.a2u = nu * m.a1u
    m.g3 = 3.0 * m.shear; m.g2 = 2.0 * m.shear
    m.ssp3d = np.sqrt((m.bulk + K["FOUR_OVER_3"] * m.shear) / rho0)
    m.soundsp = np.sqrt(young / (1.0 - nu * nu) / rho0)
    m.nu_mnu = nu / (1.0 - nu); m.t_pnu = 3.0 / (1.0 + nu); m.u_mnu = 1.0 / (1.0 - nu)
    m.epsmax = K["INFINITY"]; m.fisokin = 0.0
    m.epsr1 = K["INFINITY"]; m.epsr2 = 2.0 * K["INFINITY"]; m.epsf = 3.0 * K["INFINITY"]
    m.a11 = young / (1.0 - nu ** 2); m.a12 = 0.0
    m.ssp = np.sqrt(young / rho0)
    meshgen._sqrt_constants(m)
    m.nrate = len(curves)
    npf = [0]; tf = []
    for c, (x, y) in enumerate(curves):
        m.rate[c] = rates[c]; m.yfac[c] = yfac[c]; m.ifunc[c] = c
        tf.append(np.stack([np.asarray(x, float), np.asarray(y, float)], 1).reshape(-1)); npf.append(npf[-1] + len(x))
    m.israte = 0 if m.nrate == 1 else 1
    m.asrate = 0.0
    if m.nrate > 1:
        m.asrate = 2.0 * np.pi * (fcut if fcut > 0.0 else 10000.0)
    m.vp = 0; m.ifail = 0; m.yldcheck = 0; m.ismooth = 0 if m.nrate == 1 else ismooth
    return m, np.asarray(npf, np.int32), np.concatenate(tf)


def _shell_masses(X, ixc, rho0, thick, ihbe):
    """Nodal masses / inertias of 4-node shells as the Starter lumps them (cinmas.F:851, 921-927, 1383)."""
    area = meshgen.shell_areas(X, ixc)
    ems = rho0 * thick * area * 0.25
    fac = 12.0 if ihbe >= 11 else 9.0
    xi = ems * (area / fac + thick * thick * (1.0 / 12.0))
    n = X.shape[0]
    MS = np.zeros(n); IN = np.zeros(n)
    np.add.at(MS, (ixc[:, 1:5] - 1).reshape(-1), np.repeat(ems, 4))
    np.add.at(IN, (ixc[:, 1:5] - 1).reshape(-1), np.repeat(xi, 4))
    return MS, IN


# /FUNCT/14 "Steel" and /FUNCT/1 "curve +1" of qa-tests/miniqa/RUPTURE/FAIL_TAB/ELEM_SAMP/data/1ELEM_SAMP_0000.rad
_SAMP_STEEL = np.array([
    0, .306, .00112, .415, .00218, .445, .003, .461, .00404, .474, .00517, .489, .00613, .498, .0071, .505, .00806, .512,
    .00901, .522, .0102, .53, .0121, .543, .013, .55, .014, .555, .015, .561, .0159, .567, .0171, .572, .0181, .577,
    .0204, .592, .0303, .632, .0405, .663, .0502, .687, .06, .706, .0702, .722, .0807, .737, .09, .749, .0997, .758,
    .101, .759, .11, .768, .15000001, .805, .2, .84, .30000001, .9, .5, 1, 1, 1.21]).reshape(-1, 2)
_SAMP_F1_Y = [1, 1.0513, 1.1052, 1.1618, 1.2214, 1.284, 1.3499, 1.4191, 1.4918, 1.5683, 1.6487, 1.7333, 1.8221, 1.9155, 2.0138,
              2.117, 2.2255, 2.3396, 2.4596, 2.5857, 2.7183, 2.8577, 3.0042, 3.1582, 3.3201, 3.4903, 3.6693, 3.8574, 4.0552,
              4.2631, 4.4817, 4.7115, 4.953, 5.207, 5.4739, 5.7546, 6.0496, 6.3598, 6.6859, 7.0287, 7.3891]


def elem_samp() -> Model:
    """qa-tests/miniqa/RUPTURE/FAIL_TAB/ELEM_SAMP: two QEPH shells (Ishell 24, N=5, Ithick=1, Iplas=1, Ismstr -> 2, thickness
    1), /MAT/PLAS_TAB with curve 14 at rates 1e-6, 1e-5, 0.1, 1 (scale 1, 1, 1.2, 1.2; the Starter prepends rate 0), Fsmooth = 1,
    Fcut = 10; element 1 pulled along x, element 9 along y by /IMPVEL with the curve exp(t/10); /FAIL/TAB deletes element 1
    at cycle 2046 (t = 1.479) -- until then it only accumulates damage and leaves the stresses alone.  Units kg, mm, ms."""
    X = np.array([[0, 0, 0], [5, 0, 0], [5, 5, 0], [0, 5, 0], [80, -20, 0], [90, -20, 0], [90, -10, 0], [80, -10, 0]], float)
    itab = np.array([1, 2, 3, 4, 81, 82, 83, 84], np.int32)
    ixc = np.zeros((2, 7), np.int32)
    ixc[0] = [1, 1, 2, 3, 4, 1, 1]
    ixc[1] = [1, 5, 6, 7, 8, 1, 9]
    crv = (_SAMP_STEEL[:, 0], _SAMP_STEEL[:, 1])
    mat, npf, tf = law36(7.8e-6, 210.0, 0.3, [crv] * 5, [0.0, 1e-6, 1e-5, 0.1, 1.0], [1.0, 1.0, 1.0, 1.2, 1.2], fcut=10.0)
    prop = meshgen.default_prop_shell(thick=1.0, ihbe=24, npt=5, ismstr=2, ithk=1, ipla=1)
    MS, IN = _shell_masses(X, ixc, mat.rho0, 1.0, 24)
    # /BCS: 111, 011, 001, 101 on the translations of nodes 1..4 and 81..84 (4: x, 2: y, 1: z); rotations free
    icodt = np.array([7, 3, 1, 5, 7, 3, 1, 5], np.int32)
    m = Model(X=X, V=np.zeros_like(X), VR=np.zeros_like(X), MS=MS, IN=IN, control=meshgen.default_control(1), ixc=ixc,
              icodt=icodt, icodr=np.zeros(8, np.int32), itab=itab, npf=npf, tf=tf)
    m.shell_groups = [ShellGroup(nft=0, nel=2, law=36, mat=mat, prop=prop)]
    m.adsky, m.iads, m.iadc, m.lsky = build_pon(8, m.ixs, m.ixc)
    f1 = meshgen.add_function(m, np.arange(41) * 0.5, _SAMP_F1_Y)
    # /IMPVEL/1: x on nodes 2, 3; /IMPVEL/12: y on nodes 83, 84; scales 1, no start / stop time
    m.ibfv = np.array([[2, 1, f1], [3, 1, f1], [7, 2, f1], [8, 2, f1]], np.int32)
    m.vel = np.tile(np.array([1.0, 0.0, 1.0e30, 1.0]), (4, 1))
    return m


def law2(rho0, young, nu, ca, cb, cn, epsm=0.0, sigm=0.0, cc=0.0, eps0=0.0, icc=0, fcut=0.0, m_exp=0.0, tmelt=0.0, rhocp=0.0,
         tref=0.0, vp=0):
    """/MAT/PLAS_JOHNS as hm_read_mat02_jc.F90:170-200 completes the card (defaults, ISRATE, ASRATE)."""
    from openradioss_b200.model import Law2, elastic_constants
    g, k, a11, a12 = elastic_constants(young, nu)
    m = Law2()
    m.rho0 = rho0; m.young = young; m.nu = nu; m.shear = g; m.bulk = k
    m.ca, m.cb, m.cn = ca, cb, cn
    m.epmx = epsm if epsm else 1e20; m.sigmx = sigm if sigm else 1e20
    m.cc = cc; m.epdr = eps0 if cc else 1.0
    m.fisokin = 0.0
    m.israte = 1 if cc else 0
    m.vp = vp if vp else 2
    if m.vp == 1:
        fcut = 10000.0
    m.asrate = 1e20 if fcut == 0.0 else 2.0 * np.pi * fcut
    m.z3 = m_exp if m_exp else 1.0; m.z4 = 0.0
    m.tref = tref if tref > 0 else 300.0; m.tmelt = tmelt if tmelt else 1e20; m.rhocp = rhocp; m.tini = m.tref
    m.pshift = 0.0; m.a11 = a11; m.a12 = a12
    m.ssp = np.sqrt(young / rho0)
    m.iform = 0; m.icc = icc if icc else 1; m.has_temp = 0
    meshgen._sqrt_constants(m)
    return m


def _sh3n_masses(X, ixtg, rho0, thick):
    """3-node shells: element mass and inertia distributed by the corner angles (c3inmas.F:598, 1113, 1126-1140)."""
    tri = ixtg[:, 1:4]
    P = X[tri - 1]
    a = np.linalg.norm(P[:, 1] - P[:, 0], axis=1); b = np.linalg.norm(P[:, 2] - P[:, 1], axis=1); c = np.linalg.norm(P[:, 2] - P[:, 0], axis=1)
    area = 0.5 * np.linalg.norm(np.cross(P[:, 1] - P[:, 0], P[:, 2] - P[:, 0]), axis=1)
    ang = np.stack([np.arccos((a * a + c * c - b * b) / (2 * a * c)), np.arccos((a * a + b * b - c * c) / (2 * a * b)),
                    np.arccos((b * b + c * c - a * a) / (2 * b * c))], 1) / np.pi
    em = rho0 * thick * area
    xi = em * (area / 4.5 + thick * thick / 12.0)
    n = X.shape[0]
    MS = np.zeros(n); IN = np.zeros(n)
    np.add.at(MS, (tri - 1).reshape(-1), (em[:, None] * ang).reshape(-1))
    np.add.at(IN, (tri - 1).reshape(-1), (xi[:, None] * ang).reshape(-1))
    return MS, IN


def ct3a() -> Model:
    """qa-tests/miniqa/COQUES3N/ct3a (CT3AV4): a beam of 40 3-node shells (Ish3n 1, N=5, Ithick=0, Iplas=0, Ismstr -> 2,
    thickness .3175) on supports, /MAT/PLAS_JOHNS (A=.00295, B=.00543, n=1, SIG-MAX=.00345, no rate term), initial velocity
    -0.0117208 along z on the eight nodes of the far end, /DT scale 0.5.  Units kg, m, s as written in the deck."""
    from openradioss_b200.model import Control
    X = np.zeros((32, 3))
    for k in range(11):
        X[k] = [0.0, 1.225 * k, 0.0]; X[11 + k] = [1.525, 1.225 * k, 0.0]
    for c in range(10):
        X[22 + c] = [0.7625, 0.6125 + 1.225 * c, 0.0]
    ixtg = np.zeros((40, 6), np.int32)
    for c in range(10):
        a, b, e, d_, ctr = c + 1, c + 2, c + 12, c + 13, c + 23
        for j, (p, q) in enumerate([(b, a), (d_, b), (e, d_), (a, e)]):
            ixtg[4 * c + j] = [1, p, ctr, q, 1, 4 * c + j + 1]
    mat = law2(2.7, 0.7173, 0.3, 0.00295, 0.00543, 1.0, sigm=0.00345)
    prop = meshgen.default_prop_shell(thick=0.3175, ihbe=1, npt=5, ismstr=2, ithk=0, ipla=0)
    MS, IN = _sh3n_masses(X, ixtg, mat.rho0, 0.3175)
    # /BCS: translations 4: x, 2: y, 1: z; rotations alike
    icodt = np.zeros(32, np.int32); icodr = np.zeros(32, np.int32)
    icodt[0] = 7; icodr[0] = 7                       # node 1: 111 111
    icodt[1:10] = 4; icodr[1:10] = 3                 # nodes 2..10: 100 011
    icodt[10] = 6; icodr[10] = 7                     # node 11: 110 111
    icodt[11] = 3; icodr[11] = 7                     # node 12: 011 111
    icodt[12:21] = 0; icodr[12:21] = 3               # nodes 13..21: 000 011
    icodt[21] = 2; icodr[21] = 7                     # node 22: 010 111
    V = np.zeros_like(X)
    for n in (9, 10, 11, 20, 21, 22, 31, 32):
        V[n - 1, 2] = -0.0117208
    ctl = meshgen.default_control(1)
    ctl.dtfac_brick = ctl.dtfac_shell = ctl.dtfac_sh3n = 0.5
    m = Model(X=X, V=V, VR=np.zeros_like(X), MS=MS, IN=IN, control=ctl, ixtg=ixtg, icodt=icodt, icodr=icodr,
              itab=np.arange(1, 33, dtype=np.int32))
    m.sh3n_groups = [ShellGroup(nft=0, nel=40, law=2, mat=mat, prop=prop)]
    m.adsky, m.iads, m.iadc, m.lsky, m.iadtg = build_pon(32, m.ixs, m.ixc, m.ixtg)
    return m


def inibri_stress() -> Model:
    """qa-tests/miniqa/SOLIDES/inibri_stress (TEST_002): ONE 8-node brick (Isolid 1, Ismstr 4, Iframe 1 = not co-rotational,
    qa 1.1, qb 0.05, h 0.1), /MAT/PLAS_TAB with the single static curve /FUNCT/2, /INIBRI/STRESS sigma_xx = 900 (= the first
    point of the curve), nodal time step /DT/NODA/CST 0.9 1e-7 (no mass is added at dt = 3.6e-7).  The deck also holds a
    TYPE11 self-contact that takes the run over from cycle 7; only the cycles before are comparable (contact is outside the
    path)."""
    from openradioss_b200.model import SolidGroup
    z1 = 4.2666698
    X = np.array([[-2.5, 2.5, 0], [-2.5, 0, 0], [0, 0, 0], [0, 2.5, 0], [-2.5, 2.5, z1], [-2.5, 0, z1], [0, 0, z1], [0, 2.5, z1]], float)
    itab = np.array([33093, 33098, 33094, 33095, 33158, 33163, 33159, 33160], np.int32)
    ixs = np.zeros((1, 11), np.int32); ixs[0] = [1, 1, 2, 3, 4, 5, 6, 7, 8, 1, 26512]
    fx = [0, .00499999989, .00999999978, .0199999996, .0299999993, .0599999987, .0900000036, .100000001, .300000012, 1]
    fy = [900, 940, 960, 970, 980, 990, 1000, 1003, 1005, 1007]
    mat, npf, tf = law36(7.85e-9, 210000.0, 0.3, [(fx, fy)], [0.0], [1.0])
    prop = meshgen.default_prop_solid(jhbe=1, ismstr=4, ipla=2, istrain=1)
    vol0 = meshgen.brick_volumes(X, ixs)
    MS = np.full(8, mat.rho0 * vol0[0] / 8.0)
    ctl = meshgen.default_control(0); ctl.nodadt = 1; ctl.dtfac_node = 0.9
    m = Model(X=X, V=np.zeros_like(X), VR=np.zeros_like(X), MS=MS, IN=np.zeros(8), control=ctl, ixs=ixs, vol0=vol0, itab=itab,
              npf=npf, tf=tf)
    m.solid_groups = [SolidGroup(nft=0, nel=1, mat=mat, prop=prop, law=36)]
    m.adsky, m.iads, m.iadc, m.lsky = build_pon(8, m.ixs, m.ixc)
    m.initial_solid_sig = np.array([[900.0], [0.0], [0.0], [0.0], [0.0], [0.0]])
    return m


def loi13_solide() -> Model:
    """qa-tests/miniqa/LOIS/LOI13/solide (MODELE): a 30 mm cube of 464 8-node bricks (Isolid 1, Ismstr default -> 4, Iframe 2 =
    Belytschko's co-rotational frame, qa 1.1, qb 0.05, h 0.1) of /MAT/PLAS_JOHNS (rho 0.78, E 210000, nu 0.3, a 206, b 450,
    n 0.5, no rate term) crushed along z: the top face (/GRNOD 2) clamped, the bottom face (/GRNOD 3) held in x, y and driven in z
    by /IMPVEL with the ramp v = 0.05 t.  48 more bricks of /MAT/RIGID (LAW13) sit inside: the Engine's force loop skips LAW13
    groups (forint.F:355) -- but the Starter turns every LAW13 part into a rigid body of its own (lectur.F:8670-8690 RIGID_MAT; the
    deck's nodes 730-732 are their masters), and rigid bodies are outside the built path.  Here the 48 bricks only add their mass
    to the nodes they share, i.e. the inclusions are voids instead of rigid: the model is softer than the reference's, so only
    what does not depend on them is comparable -- the time step of cycle 0 (test_qa_decks.py); from there the energies of the
    listing run 16-19 % higher.  Units Mg, mm, s; /DT 0.9.  Geometry from tests/golden/qa_loi13_deck.npz
    (make_golden_qa.export_loi13_deck)."""
    import os
    from openradioss_b200.model import SolidGroup
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "qa_loi13_deck.npz"))
    nid, X = d["node_id"], d["X"].astype(float)
    loc = {int(n): i + 1 for i, n in enumerate(nid)}                # user id -> 1-based index
    br, part = d["brick"], d["part"]
    conn = np.vectorize(loc.get)(br[:, 1:9]).astype(np.int32)
    ixs_all = np.zeros((len(br), 11), np.int32); ixs_all[:, 0] = 1; ixs_all[:, 1:9] = conn; ixs_all[:, 9] = 1; ixs_all[:, 10] = br[:, 0]
    vol_all = meshgen.brick_volumes(X, ixs_all)
    rho = 0.78
    MS = np.zeros(len(nid)); np.add.at(MS, (conn - 1).reshape(-1), np.repeat(rho * vol_all / 8.0, 8))     # LAW13 bricks included: same density
    keep = part == 1
    ixs = np.ascontiguousarray(ixs_all[keep]); vol0 = vol_all[keep]
    mat = law2(rho, 210000.0, 0.3, 206.0, 450.0, 0.5)
    prop = meshgen.default_prop_solid(jhbe=1, ismstr=4, jcvt=1)
    n = len(nid)
    icodt = np.zeros(n, np.int32)
    icodt[[loc[int(k)] - 1 for k in d["grnod2"]]] = 7               # /BCS 111
    g3 = np.array([loc[int(k)] for k in d["grnod3"]], np.int32)
    icodt[g3 - 1] = 6                                               # /BCS 110: x, y
    ctl = meshgen.default_control(0); ctl.dtfac_brick = 0.9
    m = Model(X=X, V=np.zeros_like(X), VR=np.zeros_like(X), MS=MS, IN=np.zeros(n), control=ctl, ixs=ixs, vol0=vol0,
              icodt=icodt, itab=nid.astype(np.int32))
    ne = len(ixs)
    m.solid_groups = [SolidGroup(nft=s0, nel=min(128, ne - s0), mat=mat, prop=prop, law=2) for s0 in range(0, ne, 128)]
    m.adsky, m.iads, m.iadc, m.lsky = build_pon(n, m.ixs, m.ixc)
    f1 = meshgen.add_function(m, [0.0, 100.0], [0.0, 5.0])
    m.ibfv = np.stack([g3, np.full(len(g3), 3, np.int32), np.full(len(g3), f1, np.int32)], 1).astype(np.int32)
    m.vel = np.tile(np.array([1.0, 0.0, 1.0e30, 1.0]), (len(g3), 1))
    return m
