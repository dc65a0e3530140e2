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
    return m


def default_prop_solid(jhbe=1, ismstr=4) -> PropSolid:
    p = PropSolid()
    p.qa, p.qb = 1.1, 0.05
    p.cns1 = p.cns2 = 0.0
    p.hcoef = 0.1
    p.dtmin = 0.0
    p.jhbe = jhbe; p.ismstr = ismstr
    return p


def default_control(iroddl=0) -> Control:
    c = Control()
    c.dtfac_brick = 0.9; c.dtfac_shell = 0.9
    c.dtmx = EP20
    c.dt_init = 0.0            # DT1 of cycle 0
    c.dt2old_init = EP20
    c.tt_init = 0.0
    c.iroddl = iroddl; c.nodadt = 0
    return c


def _groups(ne: int):
    return [(s, min(NVSIZ, ne - s)) for s in range(0, ne, NVSIZ)]


def hex_block(nx: int, ny: int, nz: int, lx: float, ly: float, lz: float, *, mat: Law2 = None,
              prop: PropSolid = None, jitter: float = 0.05, seed: int = 2024, v0=(0.0, 0.0, 0.0),
              vrand: float = 0.0, vseed: int = 12345, fix_bottom_z: bool = False,
              user_id_perm: bool = False) -> Model:
    """Structured block of nx*ny*nz 8-node bricks (Taylor bar C1 / weak-scaling block C5)."""
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
    m.solid_groups = [SolidGroup(nft=s, nel=n, mat=mat, prop=prop) for s, n in _groups(ne)]
    m.adsky, m.iads, m.iadc, m.lsky = build_pon(numnod, m.ixs, m.ixc)
    return m


def taylor_bar(scale: int = 1) -> Model:
    """C1: 32x32x98 = 100 352 bricks, 6.4 x 6.4 x 32.4 mm copper bar, V0z = -227 m/s, anvil z-BC."""
    nx = ny = max(2, 32 // scale); nz = max(2, 98 // scale)
    return hex_block(nx, ny, nz, 6.4, 6.4, 32.4, v0=(0.0, 0.0, -227.0), fix_bottom_z=True)
