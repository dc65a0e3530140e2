"""Model description: ctypes mirrors of include/orgpu_model.h plus the in-memory "restart"
(what the Starter hands the Engine: nodes, connectivity, groups, /PARITH/ON tables).

Reference data model: common_source/modules/nodal_arrays.F90:125-176 (nodal arrays),
common_source/modules/parith_on_mod.F90:39-74 (ADSKY/IADS/IADC/FSKY),
engine/source/elements/forintc.F:254-300 (group descriptors).
"""
from __future__ import annotations
import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional
import numpy as np

MAXFUNC36 = 10
NVSIZ = 128          # group size (engine/share/spe_inc/mvsiz_p.inc:51-55)

d, i = C.c_double, C.c_int


class Law2(C.Structure):
    _fields_ = [(n, d) for n in ("rho0 young nu shear bulk ca cb cn epmx sigmx cc epdr fisokin asrate "
                                 "z3 z4 tref tmelt rhocp tini pshift a11 a12 ssp gsr a11sr a12sr nusr").split()] + \
               [(n, i) for n in "iform icc vp israte has_temp".split()]


class Law36(C.Structure):
    _fields_ = [(n, d) for n in ("rho0 young nu shear bulk a11 a12 ssp gsr a11sr a12sr nusr a1u a2u g3 g2 ssp3d soundsp nu_mnu t_pnu u_mnu "
                                 "epsmax epsr1 epsr2 epsf fisokin asrate").split()] + \
               [("rate", d * MAXFUNC36), ("yfac", d * MAXFUNC36), ("ifunc", i * MAXFUNC36)] + \
               [(n, i) for n in "nrate israte vp ifail yldcheck ismooth".split()]


class PropSolid(C.Structure):
    _fields_ = [(n, d) for n in "qa qb cns1 cns2 hcoef dtmin".split()] + [("jhbe", i), ("ismstr", i), ("ipla", i), ("istrain", i), ("jcvt", i), ("pad", i)]


class PropShell(C.Structure):
    _fields_ = [(n, d) for n in "thick h1 h2 h3 srh1 srh2 srh3 shf shfsr cvis dm".split()] + \
               [(n, i) for n in "npt ismstr ithk ipla ihbe istrain".split()]


class Fail(C.Structure):
    """orgpu_fail: /FAIL/JOHNSON of a shell group's material (irupt = 1) + P_thickfail (fail%pthk, GEO(42))."""
    _fields_ = [("irupt", i), ("pad", i)] + [(n, d) for n in "d1 d2 d3 d4 d5 epsp0 epsf_min pthk pthickg".split()]


class Control(C.Structure):
    _fields_ = [(n, d) for n in "dtfac_brick dtfac_shell dtmx dt_init dt2old_init tt_init".split()] + \
               [("iroddl", i), ("nodadt", i), ("dtfac_node", d), ("dtfac_sh3n", d)]


def elastic_constants(young: float, nu: float):
    """PM(22)=G, PM(32)=K, PM(24)=A11, PM(25)=A12 as the Starter derives them."""
    g = young / (2.0 * (1.0 + nu))
    k = young / (3.0 * (1.0 - 2.0 * nu))
    a11 = young / (1.0 - nu * nu)
    a12 = nu * a11
    return g, k, a11, a12


@dataclass
class SolidGroup:
    nft: int
    nel: int
    mat: object              # Law2 | Law36
    prop: PropSolid
    law: int = 2             # 2 (M2LAW) or 36 (MULAW -> SIGEPS36)
    part: int = 0            # 0-based part of the group's elements
    fail: Optional["Fail"] = None   # /FAIL/JOHNSON of the group's material (LAW2)


@dataclass
class ShellGroup:
    nft: int
    nel: int
    law: int                 # 2 or 36
    mat: object              # Law2 | Law36
    prop: PropShell
    part: int = 0            # 0-based part of the group's elements
    fail: Optional[Fail] = None   # /FAIL/JOHNSON of the group's material


@dataclass
class Model:
    """Everything a domain's restart file gives the Engine for this path."""
    X: np.ndarray                       # (numnod,3) float64  == Fortran X(3,NUMNOD)
    V: np.ndarray
    VR: np.ndarray
    MS: np.ndarray
    IN: np.ndarray
    control: Control
    ixs: np.ndarray = field(default_factory=lambda: np.zeros((0, 11), np.int32))   # IXS(11,NUMELS)^T
    ixc: np.ndarray = field(default_factory=lambda: np.zeros((0, 7), np.int32))    # IXC(7,NUMELC)^T
    ixtg: np.ndarray = field(default_factory=lambda: np.zeros((0, 6), np.int32))   # IXTG(6,NUMELTG)^T: mat, n1..n3, pid, user id
    vol0: np.ndarray = field(default_factory=lambda: np.zeros(0))                  # brick initial volumes
    solid_groups: List[SolidGroup] = field(default_factory=list)
    shell_groups: List[ShellGroup] = field(default_factory=list)
    sh3n_groups: List[ShellGroup] = field(default_factory=list)                    # 3-node shells (ITY=7); prop.ihbe = Ish3n
    icodt: Optional[np.ndarray] = None  # BCS translation codes (4:x 2:y 1:z)
    icodr: Optional[np.ndarray] = None
    fext: Optional[np.ndarray] = None   # constant nodal loads (numnod,3)
    mext: Optional[np.ndarray] = None
    itab: Optional[np.ndarray] = None   # user node ids
    # /PARITH/ON tables (1-based), filled by pon.build_pon
    adsky: Optional[np.ndarray] = None
    iads: Optional[np.ndarray] = None
    iadc: Optional[np.ndarray] = None
    iadtg: Optional[np.ndarray] = None
    lsky: int = 0
    # LAW36 function table
    npf: Optional[np.ndarray] = None
    tf: Optional[np.ndarray] = None
    load_func: Optional[tuple] = None   # (curve index, FCX): time function shared by the nodal loads fext / mext
    cload_ib: Optional[np.ndarray] = None    # concentrated loads record by record (n,3) int32: node (1-based), direction 1..6, curve (0-based, -1 constant)
    cload_fac: Optional[np.ndarray] = None   # (n,2): FCY, FCX   (instead of fext / mext / load_func)
    ibfv: Optional[np.ndarray] = None   # imposed velocities (n,3) int32: node (1-based), direction 1..3, curve index
    vel: Optional[np.ndarray] = None    # (n,4): FAC, STARTT, STOPT, FACX
    igrv: Optional[np.ndarray] = None   # gravity loads (n,3) int32: node count, direction 1..3, curve index (-1: constant)   (gravit.F)
    agrv: Optional[np.ndarray] = None   # (n,2): FCY, FCX
    ibgrv: Optional[np.ndarray] = None  # node lists of the loads, one after the other (1-based)
    ipartc: Optional[np.ndarray] = None  # part (0-based) of every 4-node shell / brick / 3-node shell: IPARTC, IPARTS, IPARTTG
    iparts: Optional[np.ndarray] = None  #   (default: the `part` attribute of the element's group, else 0)
    iparttg: Optional[np.ndarray] = None
    gorder: Optional[dict] = None       # a domain of a decomposed model: global processing order of its elements / global node index (domdec)

    @property
    def numnod(self): return int(self.X.shape[0])
    @property
    def numels(self): return int(self.ixs.shape[0])
    @property
    def numelc(self): return int(self.ixc.shape[0])
    @property
    def numeltg(self): return int(self.ixtg.shape[0])
