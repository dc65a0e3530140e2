"""Synthetic decks for the benchmark configurations of BASELINE.json.

This plays the Starter's role for structured meshes: nodes, IXS/IXC with user ids, element
groups of <= NVSIZ elements, lumped masses / inertias, BCS codes, nodal loads and the
/PARITH/ON tables.  Everything is seeded and deterministic.
"""
from __future__ import annotations
import numpy as np
from .model import (Model, Control, Law2, Law36, PropSolid, PropShell, SolidGroup, ShellGroup,
                    elastic_constants, NVSIZ, MAXFUNC36)
from .pon import build_pon
from .constants import K

EP20 = 1e20


def brick_volumes(X: np.ndarray, ixs: np.ndarray) -> np.ndarray:
    """Initial brick volumes with the Jacobian formula of sderi3.F:105-149 (det/64)."""
    c = X[ixs[:, 1:9] - 1]                      # (ne,8,3)
    x, y, z = c[:, :, 0], c[:, :, 1], c[:, :, 2]
    def d(a, p, q): return a[:, p] - a[:, q]
    x17, x28, x35, x46 = d(x, 6, 0), d(x, 7, 1), d(x, 4, 2), d(x, 5, 3)
    y17, y28, y35, y46 = d(y, 6, 0), d(y, 7, 1), d(y, 4, 2), d(y, 5, 3)
    z17, z28, z35, z46 = d(z, 6, 0), d(z, 7, 1), d(z, 4, 2), d(z, 5, 3)
    j1 = x17 + x28 - x35 - x46; j2 = y17 + y28 - y35 - y46; j3 = z17 + z28 - z35 - z46
    xa, xb = x17 + x46, x28 + x35
    ya, yb = y17 + y46, y28 + y35
    za, zb = z17 + z46, z28 + z35
    j4, j5, j6 = xa + xb, ya + yb, za + zb
    j7, j8, j9 = xa - xb, ya - yb, za - zb
    return (1.0 / 64.0) * (j1 * (j5 * j9 - j6 * j8) + j2 * (j6 * j7 - j4 * j9) + j3 * (j4 * j8 - j5 * j7))


def copper_law2(tini: float = 300.0) -> Law2:
    """OFHC copper Johnson-Cook set of SURVEY.md 8d (units mm, ms, g -> MPa)."""
    young, nu = 117.0e3, 0.35
    g, k, a11, a12 = elastic_constants(young, nu)
    m = Law2()
    m.rho0 = 8.93e-3; m.young = young; m.nu = nu; m.shear = g; m.bulk = k
    m.ca, m.cb, m.cn = 90.0, 292.0, 0.31
    m.epmx = 1e30; m.sigmx = 1e30
    m.cc = 0.025; m.epdr = 1.0e-3            # 1/s -> 1e-3 /ms
    m.fisokin = 0.0; m.asrate = 0.0
    m.z3 = 1.09; m.z4 = 0.0
    m.tref = 300.0; m.tmelt = 1356.0; m.rhocp = 3.44; m.tini = tini
    m.pshift = 0.0; m.a11 = a11; m.a12 = a12
    m.ssp = np.sqrt(a11 / m.rho0)
    m.iform = 0; m.icc = 1; m.vp = 2; m.israte = 0; m.has_temp = 1
    _sqrt_constants(m)
    return m


def _sqrt_constants(m):
    """PM(12), PM(13), PM(14), PM(190) (starter/source/materials/mat/hm_read_mat.F90:1638-1641)."""
    m.gsr = np.sqrt(max(0.0, m.shear)); m.a11sr = np.sqrt(max(0.0, m.a11))
    m.a12sr = np.sqrt(max(0.0, m.a12)); m.nusr = np.sqrt(max(0.0, m.nu))


def steel_law2_shell() -> Law2:
    """Mild steel Johnson-Cook set for shells (mm, ms, g): used by the BT / LAW2 decks."""
    young, nu = 210.0e3, 0.3
    g, k, a11, a12 = elastic_constants(young, nu)
    m = Law2()
    m.rho0 = 7.85e-3; m.young = young; m.nu = nu; m.shear = g; m.bulk = k
    m.ca, m.cb, m.cn = 250.0, 400.0, 0.4
    m.epmx = 1e30; m.sigmx = 1e30
    m.cc = 0.02; m.epdr = 1.0e-3
    m.fisokin = 0.0; m.asrate = 0.0
    m.z3 = 1.0; m.z4 = 0.0
    m.tref = 300.0; m.tmelt = 1.0e30; m.rhocp = 0.0; m.tini = 300.0
    m.pshift = 0.0; m.a11 = a11; m.a12 = a12
    m.ssp = np.sqrt(a11 / m.rho0)
    m.iform = 0; m.icc = 1; m.vp = 2; m.israte = 0; m.has_temp = 0
    _sqrt_constants(m)
    return m


def steel_law36(curves=None, rates=None, epsmax=None, eps_t=None):
    """/MAT/LAW36 steel of SURVEY.md 8d, filled as the Starter does
    (starter/source/materials/mat/mat036/hm_read_mat36.F:268-320; generic PM slots
    hm_read_mat.F90:1466-1479).  Returns (Law36, npf, tf): one static curve of 8 points,
    sigma_y 250 -> 450 MPa over eps_p 0 -> 0.3, unless `curves` (list of (x,y) arrays) and `rates`
    are given."""
    young, nu, rho0 = 210.0e3, 0.3, 7.85e-3
    m = Law36()
    m.rho0 = rho0; m.young = young; m.nu = nu
    m.shear = 0.5 * young / (1.0 + nu)
    m.bulk = young / 3.0 / (1.0 - 2.0 * nu)
    m.a1u = young / (1.0 - nu * nu); m.a2u = nu * m.a1u
    m.g3 = 3.0 * m.shear
    m.g2 = 2.0 * m.shear
    m.ssp3d = np.sqrt((m.bulk + K["FOUR_OVER_3"] * m.shear) / rho0)   # solids (hm_read_mat36.F:271)
    m.soundsp = np.sqrt(young / (1.0 - nu * nu) / rho0)
    m.nu_mnu = nu / (1.0 - nu); m.t_pnu = 3.0 / (1.0 + nu); m.u_mnu = 1.0 / (1.0 - nu)
    m.epsmax = K["INFINITY"]; m.fisokin = 0.0; m.asrate = 0.0
    m.epsr1 = K["INFINITY"]; m.epsr2 = 2.0 * K["INFINITY"]; m.epsf = 3.0 * K["INFINITY"]
    m.a11 = young / (1.0 - nu ** 2); m.a12 = 0.0          # PM(25) is not set for user-type laws
    m.ssp = np.sqrt(young / rho0)                          # PM(27) (hm_read_mat36.F:325)
    _sqrt_constants(m)
    if curves is None:
        x = np.array([0.0, 0.01, 0.02, 0.05, 0.10, 0.15, 0.20, 0.30])
        y = np.array([250.0, 290.0, 315.0, 355.0, 395.0, 420.0, 435.0, 450.0])
        curves, rates = [(x, y)], [0.0]
    m.nrate = len(curves)
    npf = [0]; tf = []
    for c, (x, y) in enumerate(curves):
        m.rate[c] = rates[c]; m.yfac[c] = 1.0; m.ifunc[c] = c
        tf.append(np.stack([x, y], 1).reshape(-1)); npf.append(npf[-1] + len(x))
    m.israte = 0 if m.nrate == 1 else 1
    if m.nrate > 1:
        m.asrate = 2.0 * np.pi * 10000.0                   # FCUT default (hm_read_mat36.F:216-218)
    m.vp = 0; m.ifail = 0; m.yldcheck = 0; m.ismooth = 0 if m.nrate == 1 else 1
    if epsmax is not None:                                 # failure plastic strain: IFAIL = 1 (hm_read_mat36.F:228-233)
        m.epsmax = epsmax; m.ifail = 1
    if eps_t is not None:                                  # (EPSR1, EPSR2, EPSF) tensile failure: IFAIL = 2 (hm_read_mat36.F:234-236)
        m.epsr1, m.epsr2, m.epsf = eps_t; m.ifail = 2
    return m, np.asarray(npf, np.int32), np.concatenate(tf)


def default_prop_shell(thick=2.0, ihbe=24, npt=5, ismstr=2, ithk=1, ipla=1) -> PropShell:
    """/PROP/SHELL defaults as the Starter fills GEO (hm_read_prop01.F:196-262)."""
    p = PropShell()
    p.thick = thick
    if 11 < ihbe < 29:                       # QEPH: GEO(13) <- Dn (default 0.015), GEO(17) <- CVIS = 1
        p.h1 = K["ZEP015"]; p.h2 = K["EM02"]; p.h3 = K["EM02"]; p.cvis = 1.0
    elif ihbe == 3:
        p.h1 = K["EM01"]; p.h2 = K["EM01"]; p.h3 = K["EM02"]; p.cvis = 0.0
    else:
        p.h1 = K["EM02"]; p.h2 = K["EM02"]; p.h3 = K["EM02"]; p.cvis = 0.0
    p.srh1 = np.sqrt(p.h1); p.srh2 = np.sqrt(p.h2); p.srh3 = np.sqrt(p.h3)
    p.shf = 0.0 if npt == 1 else K["FIVE_OVER_6"]
    p.shfsr = np.sqrt(p.shf)
    # membrane damping: GEO(16) = 0 in the deck -> the Starter's group default (set_elgroup_param.F:83-108): 1.5 % for QEPH
    # (internal IHBE 23 = Ishell 24, hm_read_prop01.F:314) with /PROP/SHELL and a law outside its special list, 0 for BT
    p.dm = K["ZEP015"] if ihbe == 24 else 0.0
    p.npt = npt; p.ismstr = ismstr; p.ithk = ithk; p.ipla = ipla; p.ihbe = ihbe; p.istrain = 1
    return p


def default_prop_solid(jhbe=1, ismstr=4, ipla=1, istrain=0, jcvt=0) -> PropSolid:
    p = PropSolid()
    p.qa, p.qb = 1.1, 0.05
    p.cns1 = p.cns2 = 0.0
    p.hcoef = 0.1
    p.dtmin = 0.0
    p.jhbe = jhbe; p.ismstr = ismstr
    p.ipla = ipla; p.istrain = istrain; p.jcvt = jcvt; p.pad = 0
    return p


def default_control(iroddl=0) -> Control:
    c = Control()
    c.dtfac_brick = 0.9; c.dtfac_shell = 0.9; c.dtfac_sh3n = 0.9
    c.dtmx = EP20
    c.dt_init = 0.0            # DT1 of cycle 0
    c.dt2old_init = EP20
    c.tt_init = 0.0
    c.iroddl = iroddl; c.nodadt = 0; c.dtfac_node = 0.9
    return c


def _groups(ne: int):
    return [(s, min(NVSIZ, ne - s)) for s in range(0, ne, NVSIZ)]


def hex_block(nx: int, ny: int, nz: int, lx: float, ly: float, lz: float, *, mat: Law2 = None,
              prop: PropSolid = None, jitter: float = 0.05, seed: int = 2024, v0=(0.0, 0.0, 0.0),
              vrand: float = 0.0, vseed: int = 12345, fix_bottom_z: bool = False,
              user_id_perm: bool = False, law: int = 2, curves=None, rates=None) -> Model:
    """Structured block of nx*ny*nz 8-node bricks (Taylor bar C1 / weak-scaling block C5).
    law=36: the steel of steel_law36 through MMAIN -> MULAW -> SIGEPS36."""
    npf = tf = None
    if law == 36 and mat is None:
        mat, npf, tf = steel_law36(curves, rates)
    mat = mat or copper_law2(); prop = prop or default_prop_solid()
    nnx, nny, nnz = nx + 1, ny + 1, nz + 1
    gx, gy, gz = np.meshgrid(np.arange(nnx), np.arange(nny), np.arange(nnz), indexing="ij")
    nid = lambda i, j, k: i + nnx * (j + nny * k)        # 0-based node number, x fastest
    numnod = nnx * nny * nnz
    X = np.empty((numnod, 3))
    idx = nid(gx, gy, gz).reshape(-1)
    X[idx, 0] = (gx * (lx / nx)).reshape(-1); X[idx, 1] = (gy * (ly / ny)).reshape(-1); X[idx, 2] = (gz * (lz / nz)).reshape(-1)
    if jitter:
        rng = np.random.default_rng(seed)
        h = min(lx / nx, ly / ny, lz / nz)
        X += rng.uniform(-jitter, jitter, X.shape) * h
    ex, ey, ez = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    ex, ey, ez = [a.transpose(2, 1, 0).reshape(-1) for a in (ex, ey, ez)]   # element order: x fastest
    ne = nx * ny * nz
    ixs = np.zeros((ne, 11), np.int32)
    ixs[:, 0] = 1
    corners = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
    for c, (a, b, cc) in enumerate(corners):
        ixs[:, 1 + c] = nid(ex + a, ey + b, ez + cc) + 1
    ixs[:, 9] = 1
    uid = np.arange(1, ne + 1, dtype=np.int32)
    if user_id_perm:                                    # user ids not in storage order
        uid = np.random.default_rng(seed + 1).permutation(ne).astype(np.int32) + 1
    ixs[:, 10] = uid
    vol0 = brick_volumes(X, ixs)
    MS = np.zeros(numnod)
    np.add.at(MS, (ixs[:, 1:9] - 1).reshape(-1), np.repeat(mat.rho0 * vol0 / 8.0, 8))
    V = np.tile(np.asarray(v0, float), (numnod, 1))
    if vrand:
        V += np.random.Generator(np.random.PCG64(vseed)).uniform(-vrand, vrand, V.shape)
    icodt = None
    if fix_bottom_z:
        icodt = np.zeros(numnod, np.int32)
        icodt[nid(gx[:, :, 0], gy[:, :, 0], 0).reshape(-1)] = 1      # z fixed (anvil)
        V[icodt == 1, 2] = 0.0
    m = Model(X=X, V=V, VR=np.zeros_like(X), MS=MS, IN=np.zeros(numnod), control=default_control(0),
              ixs=ixs, vol0=vol0, icodt=icodt, itab=np.arange(1, numnod + 1, dtype=np.int32))
    m.solid_groups = [SolidGroup(nft=s, nel=n, mat=mat, prop=prop, law=law) for s, n in _groups(ne)]
    if npf is not None:
        m.npf, m.tf = npf, tf
    m.adsky, m.iads, m.iadc, m.lsky = build_pon(numnod, m.ixs, m.ixc)
    return m


def taylor_bar(scale: int = 1) -> Model:
    """C1: 32x32x98 = 100 352 bricks, 6.4 x 6.4 x 32.4 mm copper bar, V0z = -227 m/s, anvil z-BC."""
    nx = ny = max(2, 32 // scale); nz = max(2, 98 // scale)
    return hex_block(nx, ny, nz, 6.4, 6.4, 32.4, v0=(0.0, 0.0, -227.0), fix_bottom_z=True)


def shell_areas(X: np.ndarray, ixc: np.ndarray) -> np.ndarray:
    """Quad areas as the Starter measures them: half the norm of the diagonal cross product."""
    c = X[ixc[:, 1:5] - 1]
    d1 = c[:, 2] - c[:, 0]; d2 = c[:, 3] - c[:, 1]
    return 0.5 * np.linalg.norm(np.cross(d1, d2), axis=1)


def shell_plate(nx: int, ny: int, lx: float = 1000.0, ly: float = 1000.0, *, thick: float = 2.0, law: int = 36,
                mat=None, prop: PropShell = None, jitter: float = 0.05, zjitter: float = 0.05, seed: int = 2024,
                pressure: float = 1.0, clamp: bool = True, vrand: float = 0.0, vseed: int = 12345,
                user_id_perm: bool = False, curves=None, rates=None, pulse_tau: float = 0.0, vwave=None) -> Model:
    """Square plate of nx*ny 4-node shells in the xy plane (C2: 1000 x 1000 QEPH / LAW36, clamped
    edges, uniform pressure as nodal forces; pulse_tau > 0 ramps them as p0*min(t/tau, 1) through a time
    function, the /CLOAD path of force.F90)."""
    prop = prop or default_prop_shell(thick=thick)
    npf = tf = None
    if mat is None:
        if law == 36:
            mat, npf, tf = steel_law36(curves, rates)
        else:
            mat = steel_law2_shell()
    nnx, nny = nx + 1, ny + 1
    gx, gy = np.meshgrid(np.arange(nnx), np.arange(nny), indexing="ij")
    nid = lambda i, j: i + nnx * j
    numnod = nnx * nny
    X = np.zeros((numnod, 3))
    idx = nid(gx, gy).reshape(-1)
    X[idx, 0] = (gx * (lx / nx)).reshape(-1); X[idx, 1] = (gy * (ly / ny)).reshape(-1)
    h = min(lx / nx, ly / ny)
    rng = np.random.default_rng(seed)
    if jitter:
        X[:, :2] += rng.uniform(-jitter, jitter, (numnod, 2)) * h
    if zjitter:
        X[:, 2] += rng.uniform(-zjitter, zjitter, numnod) * h
    ex, ey = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    ex, ey = ex.T.reshape(-1), ey.T.reshape(-1)
    ne = nx * ny
    ixc = np.zeros((ne, 7), np.int32)
    ixc[:, 0] = 1; ixc[:, 5] = 1
    for c, (a, b) in enumerate([(0, 0), (1, 0), (1, 1), (0, 1)]):
        ixc[:, 1 + c] = nid(ex + a, ey + b) + 1
    uid = np.arange(1, ne + 1, dtype=np.int32)
    if user_id_perm:
        uid = np.random.default_rng(seed + 1).permutation(ne).astype(np.int32) + 1
    ixc[:, 6] = uid
    area = shell_areas(X, ixc)
    ems = mat.rho0 * prop.thick * area * 0.25                       # cinmas.F:851
    fac = 12.0 if prop.ihbe >= 11 else 9.0                          # cinmas.F:921-927
    xi = ems * (area / fac + prop.thick * prop.thick * (1.0 / 12.0))  # cinmas.F:1383
    MS = np.zeros(numnod); IN = np.zeros(numnod)
    np.add.at(MS, (ixc[:, 1:5] - 1).reshape(-1), np.repeat(ems, 4))
    np.add.at(IN, (ixc[:, 1:5] - 1).reshape(-1), np.repeat(xi, 4))
    V = np.zeros((numnod, 3)); VR = np.zeros((numnod, 3))
    if vrand:
        g = np.random.Generator(np.random.PCG64(vseed))
        V += g.uniform(-vrand, vrand, V.shape); VR += g.uniform(-vrand, vrand, VR.shape) / h
    if vwave is not None:
        # smooth initial velocity field (amplitude A, wavelength lam): in-plane stretching / compression plus an out-of-plane
        # bulge, strong enough that the plate yields within the first cycles (the plastic return of the law is then live)
        A, lam = vwave
        kx = 2.0 * np.pi / lam
        V[:, 0] += A * np.sin(kx * X[:, 0]); V[:, 1] += A * np.sin(kx * X[:, 1])
        V[:, 2] += 0.5 * A * np.sin(kx * X[:, 0]) * np.sin(kx * X[:, 1])
    fext = None
    if pressure:
        fext = np.zeros((numnod, 3))
        np.add.at(fext[:, 2], (ixc[:, 1:5] - 1).reshape(-1), np.repeat(pressure * area * 0.25, 4))
    icodt = icodr = None
    if clamp:
        icodt = np.zeros(numnod, np.int32); icodr = np.zeros(numnod, np.int32)
        edge = (gx == 0) | (gx == nx) | (gy == 0) | (gy == ny)
        icodt[nid(gx, gy)[edge]] = 7; icodr[nid(gx, gy)[edge]] = 7
        V[icodt == 7] = 0.0; VR[icodr == 7] = 0.0
    m = Model(X=X, V=V, VR=VR, MS=MS, IN=IN, control=default_control(1), ixc=ixc, icodt=icodt, icodr=icodr,
              fext=fext, itab=np.arange(1, numnod + 1, dtype=np.int32), npf=npf, tf=tf)
    m.shell_groups = [ShellGroup(nft=s, nel=n, law=law, mat=mat, prop=prop) for s, n in _groups(ne)]
    m.adsky, m.iads, m.iadc, m.lsky = build_pon(numnod, m.ixs, m.ixc)
    if pulse_tau > 0.0 and fext is not None:
        m.load_func = (add_function(m, [0.0, pulse_tau, 1.0e30], [0.0, 1.0, 1.0]), 1.0)
    return m


def tri_plate(nx: int, ny: int, lx: float = 1000.0, ly: float = 1000.0, *, thick: float = 2.0, law: int = 36, mat=None,
              prop: PropShell = None, quads: str = "none", jitter: float = 0.05, zjitter: float = 0.05, seed: int = 2024,
              pressure: float = 1.0, clamp: bool = True, vrand: float = 0.0, vseed: int = 12345, user_id_perm: bool = False,
              curves=None, rates=None, quad_prop: PropShell = None, vwave=None) -> Model:
    """Plate of 3-node shells (C3FORC3, Ish3n = prop.ihbe in {1, 2}): every cell of an nx*ny grid split into two triangles;
    quads="checker" keeps every other cell as a 4-node shell (quad_prop, default QEPH) so that both families share nodes.
    Nodal masses / inertias of the triangles as the Starter distributes them (c3inmas.F:598, 1113, 1126-1140: by the
    corner angles)."""
    prop = prop or default_prop_shell(thick=thick, ihbe=2)
    quad_prop = quad_prop or default_prop_shell(thick=thick)
    npf = tf = None
    if mat is None:
        if law == 36:
            mat, npf, tf = steel_law36(curves, rates)
        else:
            mat = steel_law2_shell()
    nnx, nny = nx + 1, ny + 1
    gx, gy = np.meshgrid(np.arange(nnx), np.arange(nny), indexing="ij")
    nid = lambda i, j: i + nnx * j
    numnod = nnx * nny
    X = np.zeros((numnod, 3))
    idx = nid(gx, gy).reshape(-1)
    X[idx, 0] = (gx * (lx / nx)).reshape(-1); X[idx, 1] = (gy * (ly / ny)).reshape(-1)
    h = min(lx / nx, ly / ny)
    rng = np.random.default_rng(seed)
    if jitter:
        X[:, :2] += rng.uniform(-jitter, jitter, (numnod, 2)) * h
    if zjitter:
        X[:, 2] += rng.uniform(-zjitter, zjitter, numnod) * h
    ex, ey = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    ex, ey = ex.T.reshape(-1), ey.T.reshape(-1)
    isq = ((ex + ey) % 2 == 0) if quads == "checker" else np.zeros(ex.size, bool)
    c00, c10, c11, c01 = nid(ex, ey) + 1, nid(ex + 1, ey) + 1, nid(ex + 1, ey + 1) + 1, nid(ex, ey + 1) + 1
    tq = ~isq
    tri = np.concatenate([np.stack([c00[tq], c10[tq], c11[tq]], 1), np.stack([c00[tq], c11[tq], c01[tq]], 1)]).astype(np.int32)
    order = np.argsort(np.concatenate([np.flatnonzero(tq) * 2, np.flatnonzero(tq) * 2 + 1]), kind="stable")
    tri = tri[order]
    ntg, nq = tri.shape[0], int(isq.sum())
    ixtg = np.zeros((ntg, 6), np.int32); ixtg[:, 0] = 1; ixtg[:, 1:4] = tri; ixtg[:, 4] = 2
    ixc = np.zeros((nq, 7), np.int32); ixc[:, 0] = 1; ixc[:, 5] = 1
    ixc[:, 1] = c00[isq]; ixc[:, 2] = c10[isq]; ixc[:, 3] = c11[isq]; ixc[:, 4] = c01[isq]
    uidq = np.arange(1, nq + 1, dtype=np.int32); uidt = np.arange(1, ntg + 1, dtype=np.int32) + 10 * (nq + 1)
    if user_id_perm:
        r = np.random.default_rng(seed + 1)
        uidq = r.permutation(nq).astype(np.int32) + 1; uidt = r.permutation(ntg).astype(np.int32) + 1 + 10 * (nq + 1)
    ixc[:, 6] = uidq; ixtg[:, 5] = uidt
    MS = np.zeros(numnod); IN = np.zeros(numnod)
    fext = np.zeros((numnod, 3)) if pressure else None
    # triangles
    P = X[tri - 1]
    a = np.linalg.norm(P[:, 1] - P[:, 0], axis=1); b = np.linalg.norm(P[:, 2] - P[:, 1], axis=1); c = np.linalg.norm(P[:, 2] - P[:, 0], axis=1)
    area_t = 0.5 * np.linalg.norm(np.cross(P[:, 1] - P[:, 0], P[:, 2] - P[:, 0]), axis=1)
    ang = np.stack([np.arccos((a * a + c * c - b * b) / (2 * a * c)), np.arccos((a * a + b * b - c * c) / (2 * a * b)),
                    np.arccos((b * b + c * c - a * a) / (2 * b * c))], 1) / np.pi
    em = mat.rho0 * prop.thick * area_t
    xi = em * (area_t / 4.5 + prop.thick * prop.thick / 12.0)
    np.add.at(MS, (tri - 1).reshape(-1), (em[:, None] * ang).reshape(-1))
    np.add.at(IN, (tri - 1).reshape(-1), (xi[:, None] * ang).reshape(-1))
    if pressure:
        np.add.at(fext[:, 2], (tri - 1).reshape(-1), np.repeat(pressure * area_t / 3.0, 3))
    if nq:
        area = shell_areas(X, ixc)
        ems = mat.rho0 * quad_prop.thick * area * 0.25
        fac = 12.0 if quad_prop.ihbe >= 11 else 9.0
        xq = ems * (area / fac + quad_prop.thick * quad_prop.thick * (1.0 / 12.0))
        np.add.at(MS, (ixc[:, 1:5] - 1).reshape(-1), np.repeat(ems, 4))
        np.add.at(IN, (ixc[:, 1:5] - 1).reshape(-1), np.repeat(xq, 4))
        if pressure:
            np.add.at(fext[:, 2], (ixc[:, 1:5] - 1).reshape(-1), np.repeat(pressure * area * 0.25, 4))
    V = np.zeros((numnod, 3)); VR = np.zeros((numnod, 3))
    if vrand:
        g = np.random.Generator(np.random.PCG64(vseed))
        V += g.uniform(-vrand, vrand, V.shape); VR += g.uniform(-vrand, vrand, VR.shape) / h
    if vwave is not None:                                   # the smooth yielding field of shell_plate
        A, lam = vwave
        kx = 2.0 * np.pi / lam
        V[:, 0] += A * np.sin(kx * X[:, 0]); V[:, 1] += A * np.sin(kx * X[:, 1])
        V[:, 2] += 0.5 * A * np.sin(kx * X[:, 0]) * np.sin(kx * X[:, 1])
    icodt = icodr = None
    if clamp:
        icodt = np.zeros(numnod, np.int32); icodr = np.zeros(numnod, np.int32)
        edge = (gx == 0) | (gx == nx) | (gy == 0) | (gy == ny)
        icodt[nid(gx, gy)[edge]] = 7; icodr[nid(gx, gy)[edge]] = 7
        V[icodt == 7] = 0.0; VR[icodr == 7] = 0.0
    m = Model(X=X, V=V, VR=VR, MS=MS, IN=IN, control=default_control(1), ixc=ixc, ixtg=ixtg, icodt=icodt, icodr=icodr,
              fext=fext, itab=np.arange(1, numnod + 1, dtype=np.int32), npf=npf, tf=tf)
    m.shell_groups = [ShellGroup(nft=s, nel=n, law=law, mat=mat, prop=quad_prop) for s, n in _groups(nq)] if nq else []
    m.sh3n_groups = [ShellGroup(nft=s, nel=n, law=law, mat=mat, prop=prop) for s, n in _groups(ntg)]
    m.adsky, m.iads, m.iadc, m.lsky, m.iadtg = build_pon(numnod, m.ixs, m.ixc, m.ixtg)
    return m


def add_function(m: Model, x, y) -> int:
    """Append one (x, y) curve to the model's NPC / TF table; returns its 0-based index."""
    x = np.asarray(x, float); y = np.asarray(y, float)
    npf = np.zeros(1, np.int32) if m.npf is None else np.asarray(m.npf, np.int32)
    tf = np.zeros(0) if m.tf is None else np.asarray(m.tf, float)
    m.tf = np.concatenate([tf, np.stack([x, y], 1).reshape(-1)])
    m.npf = np.concatenate([npf, [npf[-1] + len(x)]]).astype(np.int32)
    return len(m.npf) - 2


def add_gravity(m: Model, direction: int, fcy: float, nodes=None, curve=None, fcx: float = 1.0) -> None:
    """One /GRAV load (gravit.F): acceleration fcy * f(TT * fcx) along `direction` (1..3) on `nodes` (0-based; all when None);
    curve = (x, y) time function or None for a constant."""
    nodes = np.arange(m.numnod) if nodes is None else np.asarray(nodes)
    f = -1 if curve is None else add_function(m, *curve)
    row = np.array([[len(nodes), direction, f]], np.int32); a = np.array([[fcy, fcx]], float)
    m.igrv = row if m.igrv is None else np.concatenate([m.igrv, row])
    m.agrv = a if m.agrv is None else np.concatenate([m.agrv, a])
    ib = (nodes + 1).astype(np.int32)
    m.ibgrv = ib if m.ibgrv is None else np.concatenate([m.ibgrv, ib])


def plate_c2(scale: int = 1) -> Model:
    """C2: 1000 x 1000 QEPH shells, LAW36, NPT=5, clamped, pressure pulse (1 MPa step)."""
    n = max(4, 1000 // scale)
    return shell_plate(n, n)


def shell_on_block(nx: int, ny: int, nz: int, h: float = 5.0, *, thick: float = 1.0, ihbe: int = 24, vrand: float = 5.0,
                   seed: int = 2024) -> Model:
    """Mixed model (C4 in miniature): an nx*ny*nz brick block (LAW2) with an nx*ny shell skin (LAW36)
    glued on its top face -- shells and bricks share nodes, one skyline, solids' slots first
    (FILLCNE type order, starter/source/spmd/domdec2.F:2138-2240)."""
    b = hex_block(nx, ny, nz, h * nx, h * ny, h * nz, jitter=0.05, seed=seed, vrand=vrand, fix_bottom_z=True)
    nnx, nny = nx + 1, ny + 1
    nid = lambda i, j, k: i + nnx * (j + nny * k)
    ex, ey = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    ex, ey = ex.T.reshape(-1), ey.T.reshape(-1)
    ne = nx * ny
    ixc = np.zeros((ne, 7), np.int32); ixc[:, 0] = 2; ixc[:, 5] = 2
    for c, (a, bb) in enumerate([(0, 0), (1, 0), (1, 1), (0, 1)]):
        ixc[:, 1 + c] = nid(ex + a, ey + bb, nz) + 1
    ixc[:, 6] = np.arange(1, ne + 1) + 10_000_000            # user ids of another part
    prop = default_prop_shell(thick=thick, ihbe=ihbe)
    mat, npf, tf = steel_law36()
    area = shell_areas(b.X, ixc)
    ems = mat.rho0 * thick * area * 0.25
    xi = ems * (area / (12.0 if ihbe >= 11 else 9.0) + thick * thick / 12.0)
    MS = b.MS.copy(); IN = np.zeros(b.numnod)
    np.add.at(MS, (ixc[:, 1:5] - 1).reshape(-1), np.repeat(ems, 4))
    np.add.at(IN, (ixc[:, 1:5] - 1).reshape(-1), np.repeat(xi, 4))
    rng = np.random.default_rng(seed + 7)
    VR = np.zeros_like(b.X)
    top = np.unique(ixc[:, 1:5] - 1)
    VR[top] = rng.uniform(-vrand, vrand, (len(top), 3)) / h
    ctl = default_control(1)
    m = Model(X=b.X, V=b.V, VR=VR, MS=MS, IN=IN, control=ctl, ixs=b.ixs, ixc=ixc, vol0=b.vol0, icodt=b.icodt,
              icodr=np.zeros(b.numnod, np.int32), itab=b.itab, npf=npf, tf=tf)
    m.solid_groups = b.solid_groups
    m.shell_groups = [ShellGroup(nft=s, nel=n, law=36, mat=mat, prop=prop) for s, n in _groups(ne)]
    m.adsky, m.iads, m.iadc, m.lsky = build_pon(m.numnod, m.ixs, m.ixc)
    return m


def crush_tube(nw: int, nz: int, nbz: int = 1, h: float = 2.5, *, thick: float = 1.5, v_imp: float = -10.0,
               ramp: float = 0.05, jitter: float = 0.05, seed: int = 2024) -> Model:
    """C4: thin-walled square tube (4 walls x nw x nz QEPH shells, LAW36) standing on an nw x nw x nbz brick
    end block (LAW2) whose top-face perimeter nodes are the tube's bottom ring; the block's bottom face is
    z-fixed (BCS), the top ring is driven at v_imp (mm/ms = m/s) along z through FIXVEL with a ramp of
    `ramp` ms.  BASELINE size: crush_tube(708, 706) = 2.0 M shells + 501 k bricks."""
    b = hex_block(nw, nw, nbz, h * nw, h * nw, h * nbz, jitter=jitter, seed=seed, fix_bottom_z=True)
    b.X[:, 2] -= h * nbz                                   # block top face at z = 0
    nn1 = nw + 1
    nid = lambda i, j, k: i + nn1 * (j + nn1 * k)
    P = 4 * nw
    p = np.arange(P)
    side, q = p // nw, p % nw
    ri = np.select([side == 0, side == 1, side == 2, side == 3], [q, nw, nw - q, 0])
    rj = np.select([side == 0, side == 1, side == 2, side == 3], [0, q, nw, nw - q])
    ring0 = nid(ri, rj, nbz)                               # tube ring 0 = perimeter of the block's top face
    nb_nodes = b.numnod
    numnod = nb_nodes + nz * P
    X = np.zeros((numnod, 3)); X[:nb_nodes] = b.X
    rng = np.random.default_rng(seed + 11)
    for k in range(1, nz + 1):
        sl = slice(nb_nodes + (k - 1) * P, nb_nodes + k * P)
        X[sl, 0] = ri * h; X[sl, 1] = rj * h; X[sl, 2] = k * h
        X[sl] += rng.uniform(-jitter, jitter, (P, 3)) * h
    ring = lambda pp, k: np.where(k == 0, ring0[pp % P], nb_nodes + (k - 1) * P + (pp % P))
    ek, ep = np.meshgrid(np.arange(nz), p, indexing="ij")
    ek, ep = ek.reshape(-1), ep.reshape(-1)
    ne = nz * P
    ixc = np.zeros((ne, 7), np.int32); ixc[:, 0] = 2; ixc[:, 5] = 2
    ixc[:, 1] = ring(ep, ek) + 1; ixc[:, 2] = ring(ep + 1, ek) + 1
    ixc[:, 3] = ring(ep + 1, ek + 1) + 1; ixc[:, 4] = ring(ep, ek + 1) + 1
    ixc[:, 6] = np.arange(1, ne + 1) + 10_000_000
    prop = default_prop_shell(thick=thick)
    mat, npf, tf = steel_law36()
    area = shell_areas(X, ixc)
    ems = mat.rho0 * thick * area * 0.25
    xi = ems * (area / 12.0 + thick * thick / 12.0)
    MS = np.zeros(numnod); MS[:nb_nodes] = b.MS; IN = np.zeros(numnod)
    np.add.at(MS, (ixc[:, 1:5] - 1).reshape(-1), np.repeat(ems, 4))
    np.add.at(IN, (ixc[:, 1:5] - 1).reshape(-1), np.repeat(xi, 4))
    icodt = np.zeros(numnod, np.int32); icodt[:nb_nodes] = b.icodt
    m = Model(X=X, V=np.zeros((numnod, 3)), VR=np.zeros((numnod, 3)), MS=MS, IN=IN, control=default_control(1), ixs=b.ixs, ixc=ixc,
              vol0=b.vol0, icodt=icodt, icodr=np.zeros(numnod, np.int32), itab=np.arange(1, numnod + 1, dtype=np.int32), npf=npf, tf=tf)
    m.solid_groups = b.solid_groups
    m.shell_groups = [ShellGroup(nft=s, nel=n, law=36, mat=mat, prop=prop) for s, n in _groups(ne)]
    m.adsky, m.iads, m.iadc, m.lsky = build_pon(numnod, m.ixs, m.ixc)
    f = add_function(m, [0.0, ramp, 1.0e30], [0.0, 1.0, 1.0])
    top = nb_nodes + (nz - 1) * P + p if nz >= 1 else ring0
    m.ibfv = np.stack([top + 1, np.full(P, 3), np.full(P, f)], 1).astype(np.int32)
    m.vel = np.tile(np.array([v_imp, 0.0, 1.0e30, 1.0]), (P, 1))
    return m
