"""ctypes plumbing shared by the product binding (engine.py -> liborgpu.so) and the test-only
oracle binding (oracle/orc.py -> liborc.so).  Both libraries export the same call surface
(prefix ``orgpu_`` / ``orc_``), so one driver loads a :class:`Model` into either."""
from __future__ import annotations
import ctypes as C
import numpy as np
from .model import Model, Law2, Law36, PropSolid, PropShell, Control

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def _opt(a, dtype):
    """numpy array -> pointer (or NULL)."""
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=dtype)
    return a.ctypes.data_as(C.c_void_p)


class Binding:
    """Thin object wrapper over one handle of either library."""
    SOLID_FIELDS = dict(sig=(0, 6), eint=(1, 1), rho=(2, 1), qvis=(3, 1), pla=(4, 1), epsd=(5, 1),
                        vol=(6, 1), off=(7, 1), temp=(8, 1), smstr=(9, 21), stra=(10, 6), wpla=(11, 1), sigb=(12, 6), dfmax=(13, 1))
    SHELL_FIELDS = dict(forc=(0, 5), mom=(1, 3), eint=(2, 2), thk=(3, 1), off=(4, 1), stra=(5, 8),
                        epsd=(6, 1), hourg=(7, 12), smstr=(8, 6), sig=(9, 5), pla=(10, 1),
                        epsd_ip=(11, 1), temp=(12, 1), sigb=(13, 3), dfmax=(14, 1), foff=(15, 1), plap=(16, 1))

    def __init__(self, lib: C.CDLL, prefix: str, returns_status: bool):
        self.lib, self.p, self.status = lib, prefix, returns_status
        self.h = C.c_void_p()
        self.model = None
        self._keep = []

    # -- call helper: raise on a negative status for the product library
    def _f(self, name):
        return getattr(self.lib, self.p + name)

    def _call(self, name, *args):
        fn = self._f(name)
        fn.restype = C.c_int if self.status else None
        r = fn(*args)
        if self.status and r is not None and r < 0:
            err = self._f("last_error"); err.restype = C.c_char_p
            raise RuntimeError(f"{self.p}{name} failed ({r}): {err().decode()}")
        return r

    # -- model load --------------------------------------------------------------------
    def load(self, m: Model, device: int = 0):
        self.model = m
        ctl = m.control
        if self.status:
            self._call("create", C.byref(self.h), C.c_int(device), C.c_int(m.numnod), C.byref(ctl))
        else:
            fn = self._f("create"); fn.restype = C.c_void_p
            self.h = C.c_void_p(fn(C.c_int(m.numnod), C.byref(ctl)))
        self.upload_nodes(X=m.X, V=m.V, VR=m.VR, MS=m.MS, IN=m.IN)
        if m.fext is not None or m.mext is not None:
            self._call("set_loads", self.h, _opt(m.fext, np.float64), _opt(m.mext, np.float64))
        if m.icodt is not None:
            self._call("set_bcs", self.h, _opt(m.icodt, np.int32), _opt(m.icodr, np.int32))
        if m.numels:
            self._call("set_solids", self.h, C.c_int(m.numels), _opt(m.ixs, np.int32), _opt(m.iads, np.int32))
        if m.numelc:
            self._call("set_shells", self.h, C.c_int(m.numelc), _opt(m.ixc, np.int32), _opt(m.iadc, np.int32))
        if m.numeltg:
            self._call("set_sh3n", self.h, C.c_int(m.numeltg), _opt(m.ixtg, np.int32), _opt(m.iadtg, np.int32))
        self._call("set_pon", self.h, _opt(m.adsky, np.int32), C.c_int(m.lsky))
        if m.npf is not None:
            self._call("set_functions", self.h, C.c_int(len(m.npf) - 1), _opt(m.npf, np.int32), _opt(m.tf, np.float64))
        if m.itab is not None:
            self._call("set_itab", self.h, _opt(m.itab, np.int32))
        if m.cload_ib is not None and len(m.cload_ib):
            self._call("set_cloads", self.h, C.c_int(len(m.cload_ib)), _opt(m.cload_ib, np.int32), _opt(m.cload_fac, np.float64))
        if m.load_func is not None:
            self._call("set_load_function", self.h, C.c_int(int(m.load_func[0])), C.c_double(float(m.load_func[1])))
        if m.ibfv is not None and len(m.ibfv):
            self._call("set_fixvel", self.h, C.c_int(len(m.ibfv)), _opt(m.ibfv, np.int32), _opt(m.vel, np.float64))
        if m.igrv is not None and len(m.igrv):
            self._call("set_gravity", self.h, C.c_int(len(m.igrv)), _opt(m.igrv, np.int32), _opt(m.agrv, np.float64), _opt(m.ibgrv, np.int32), C.c_int(len(m.ibgrv)))
        for g in m.shell_groups:
            r = self._call_group("add_shell_group", self.h, C.c_int(g.nel), C.c_int(g.nft), C.c_int(g.law),
                                 C.byref(g.mat), C.byref(g.prop))
            if getattr(g, "fail", None) is not None:
                self._call_group("set_shell_group_fail", self.h, C.c_int(0), C.c_int(r), C.byref(g.fail))
        for g in m.sh3n_groups:
            r = self._call_group("add_sh3n_group", self.h, C.c_int(g.nel), C.c_int(g.nft), C.c_int(g.law),
                                 C.byref(g.mat), C.byref(g.prop))
            if getattr(g, "fail", None) is not None:
                self._call_group("set_shell_group_fail", self.h, C.c_int(1), C.c_int(r), C.byref(g.fail))
        for g in m.solid_groups:
            v0 = np.ascontiguousarray(m.vol0[g.nft:g.nft + g.nel])
            if getattr(g, "law", 2) == 2:
                r = self._call_group("add_solid_group", self.h, C.c_int(g.nel), C.c_int(g.nft), C.byref(g.mat),
                                     C.byref(g.prop), v0.ctypes.data_as(C.c_void_p))
            else:
                r = self._call_group("add_solid_group_law", self.h, C.c_int(g.nel), C.c_int(g.nft), C.c_int(g.law),
                                     C.byref(g.mat), C.byref(g.prop), v0.ctypes.data_as(C.c_void_p))
            if getattr(g, "fail", None) is not None:
                self._call_group("set_solid_group_fail", self.h, C.c_int(r), C.byref(g.fail))
        self._set_parts(m)
        if self.status and getattr(m, "gorder", None) is not None:      # a domain of a decomposed model: tie-break keys of the dt arg-min
            go = m.gorder
            self._call("set_global_order", self.h, _opt(go.get("shell"), np.int32), _opt(go.get("sh3n"), np.int32),
                       _opt(go.get("solid"), np.int32), _opt(go.get("node"), np.int32))
        self._call("finalize", self.h)
        return self

    def _set_parts(self, m: Model):
        """Part of every element (IPARTC / IPARTS / IPARTTG, 0-based here) and GBUF%VOL of the shells (initial area x
        thickness, what the Starter stores): the inputs of the print-cycle balances (CBILAN / SBILAN / ECRIT)."""
        from . import meshgen
        def parts(groups, n, given):
            if given is not None:
                return np.ascontiguousarray(given, np.int32)
            out = np.zeros(n, np.int32)
            for g in groups:
                out[g.nft:g.nft + g.nel] = getattr(g, "part", 0)
            return out
        ipc = parts(m.shell_groups, m.numelc, m.ipartc); ips = parts(m.solid_groups, m.numels, m.iparts)
        ipt = parts(m.sh3n_groups, m.numeltg, m.iparttg)
        npart = int(max([0] + [a.max() + 1 for a in (ipc, ips, ipt) if a.size]))
        gvc = np.zeros(m.numelc); gvt = np.zeros(m.numeltg)
        if m.numelc:
            area = meshgen.shell_areas(m.X, m.ixc)
            for g in m.shell_groups:
                gvc[g.nft:g.nft + g.nel] = area[g.nft:g.nft + g.nel] * g.prop.thick
        if m.numeltg:
            P = m.X[m.ixtg[:, 1:4] - 1]
            area = 0.5 * np.linalg.norm(np.cross(P[:, 1] - P[:, 0], P[:, 2] - P[:, 0]), axis=1)
            for g in m.sh3n_groups:
                gvt[g.nft:g.nft + g.nel] = area[g.nft:g.nft + g.nel] * g.prop.thick
        self.npart = max(npart, 1)
        self._call("set_parts", self.h, C.c_int(self.npart), _opt(ipc, np.int32), _opt(ips, np.int32), _opt(ipt, np.int32),
                   _opt(gvc, np.float64), _opt(gvt, np.float64))

    def set_quadrature(self, npt, z0, wf, wm):
        """Through-thickness rule of the NPT-point shells (positions, force weights, moment weights); process-wide."""
        a = [np.ascontiguousarray(x, np.float64) for x in (z0, wf, wm)]
        self._call("set_quadrature", self.h, C.c_int(npt), *[x.ctypes.data_as(C.c_void_p) for x in a])

    def balance_history(self, n):
        out = np.zeros((n, 8))
        self._call("get_balance_history", self.h, C.c_int(n), out.ctypes.data_as(C.c_void_p))
        return out

    def set_print(self, on: bool = True):
        """IPRI = 1: every following cycle also books the balances of the reference's print cycles."""
        self._call("set_print", self.h, C.c_int(1 if on else 0))

    def balance(self):
        """The global line ECRIT prints for the last cycle (ecrit.F:178-352) and PARTSAV(1:6, part)."""
        out = np.zeros(8); ps = np.zeros((self.npart, 6))
        self._call("get_balance", self.h, out.ctypes.data_as(C.c_void_p), ps.ctypes.data_as(C.c_void_p))
        d = dict(zip(("encin", "enrot", "enint", "wfext", "xmomt", "ymomt", "zmomt", "xmass"), out))
        d["partsav"] = ps
        return d

    def _call_group(self, name, *args):
        fn = self._f(name); fn.restype = C.c_int
        r = fn(*args)
        if r < 0:
            msg = ""
            if self.status:
                err = self._f("last_error"); err.restype = C.c_char_p; msg = err().decode()
            raise RuntimeError(f"{self.p}{name} failed ({r}) {msg}")
        return r

    def close(self):
        if self.h:
            self._call("destroy", self.h); self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- nodal arrays ------------------------------------------------------------------
    def upload_nodes(self, X=None, V=None, VR=None, D=None, MS=None, IN=None):
        self._call("upload_nodes", self.h, *[_opt(a, np.float64) for a in (X, V, VR, D, MS, IN)])

    def download_nodes(self, names=("X", "V", "D")):
        n = self.model.numnod
        order = ("X", "V", "VR", "D", "A", "AR", "STIFN", "STIFR")
        out = {k: (np.zeros((n, 3)) if k not in ("STIFN", "STIFR") else np.zeros(n)) for k in names}
        args = [out[k].ctypes.data_as(C.c_void_p) if k in out else None for k in order]
        self._call("download_nodes", self.h, *args)
        return out

    def download_fsky(self):
        f = np.zeros((self.model.lsky, 8))
        self._call("download_fsky", self.h, f.ctypes.data_as(C.c_void_p))
        return f

    def solid_state(self, name):
        fid, nc = self.SOLID_FIELDS[name]
        out = np.zeros((nc, self.model.numels))
        self._call("download_solid_state", self.h, C.c_int(fid), out.ctypes.data_as(C.c_void_p))
        return out

    def shell_state(self, name):
        fid, nc = self.SHELL_FIELDS[name]
        npt = 1
        if name in ("sig", "pla", "epsd_ip", "temp", "sigb", "dfmax", "foff", "plap"):
            npt = max(g.prop.npt for g in self.model.shell_groups)
        if name == "hourg" and not all(21 <= g.prop.ihbe <= 29 for g in self.model.shell_groups):
            nc = 12 if any(21 <= g.prop.ihbe <= 29 for g in self.model.shell_groups) else 5
        out = np.zeros((nc * npt, self.model.numelc))
        self._call("download_shell_state", self.h, C.c_int(fid), out.ctypes.data_as(C.c_void_p))
        return out

    def sh3n_state(self, name):
        """State of the 3-node shells, same field names as shell_state (no hourglass words; smstr has 3 components)."""
        fid, nc = self.SHELL_FIELDS[name]
        npt = 1
        if name in ("sig", "pla", "epsd_ip", "temp", "sigb", "dfmax", "foff", "plap"):
            npt = max(g.prop.npt for g in self.model.sh3n_groups)
        if name == "smstr":
            nc = 3
        out = np.zeros((nc * npt, self.model.numeltg))
        self._call("download_sh3n_state", self.h, C.c_int(fid), out.ctypes.data_as(C.c_void_p))
        return out

    # -- restart / state hand-over ------------------------------------------------------------------
    def upload_solid_state(self, name, arr):
        fid, nc = self.SOLID_FIELDS[name]
        a = np.ascontiguousarray(arr, np.float64); assert a.shape == (nc, self.model.numels)
        self._call("upload_solid_state", self.h, C.c_int(fid), a.ctypes.data_as(C.c_void_p))

    def upload_shell_state(self, name, arr):
        fid, _ = self.SHELL_FIELDS[name]
        a = np.ascontiguousarray(arr, np.float64); assert a.shape[1] == self.model.numelc
        self._call("upload_shell_state", self.h, C.c_int(fid), a.ctypes.data_as(C.c_void_p))

    def upload_sh3n_state(self, name, arr):
        fid, _ = self.SHELL_FIELDS[name]
        a = np.ascontiguousarray(arr, np.float64); assert a.shape[1] == self.model.numeltg
        self._call("upload_sh3n_state", self.h, C.c_int(fid), a.ctypes.data_as(C.c_void_p))

    def set_time(self, tt, dt2, dt2old, ncycle):
        self._call("set_time", self.h, C.c_double(tt), C.c_double(dt2), C.c_double(dt2old), C.c_longlong(ncycle))

    # -- corner rows of the skyline (domain exchange) ------------------------------------
    def pack_rows(self, slots):
        slots = np.ascontiguousarray(slots, np.int32)
        buf = np.zeros((len(slots), 8))
        if len(slots):
            self._call("pack_rows", self.h, C.c_int(len(slots)), slots.ctypes.data_as(C.c_void_p), buf.ctypes.data_as(C.c_void_p))
        return buf

    def unpack_rows(self, slots, buf):
        slots = np.ascontiguousarray(slots, np.int32); buf = np.ascontiguousarray(buf, np.float64)
        if len(slots):
            self._call("unpack_rows", self.h, C.c_int(len(slots)), slots.ctypes.data_as(C.c_void_p), buf.ctypes.data_as(C.c_void_p))

    # -- nodal partial sums of frontier nodes (/PARITH/OFF exchange, SPMD_EXCH_A) -----------------------
    def set_parith(self, iparit):
        """IPARIT: 1 (default) the load of a node is added behind its element rows (FORCE's own FSKY rows, /PARITH/ON), 0 the nodal
        sum starts from it (/PARITH/OFF: FORCE adds to A before the element loop)."""
        self._call("set_parith", self.h, C.c_int(iparit))

    def pack_nodes(self, nodes):
        nodes = np.ascontiguousarray(nodes, np.int32)
        buf = np.zeros((len(nodes), 8))
        if len(nodes):
            self._call("pack_nodes", self.h, C.c_int(len(nodes)), nodes.ctypes.data_as(C.c_void_p), buf.ctypes.data_as(C.c_void_p))
        return buf

    def add_nodes(self, nodes, buf):
        nodes = np.ascontiguousarray(nodes, np.int32); buf = np.ascontiguousarray(buf, np.float64)
        if len(nodes):
            self._call("add_nodes", self.h, C.c_int(len(nodes)), nodes.ctypes.data_as(C.c_void_p), buf.ctypes.data_as(C.c_void_p))

    # -- stepping ----------------------------------------------------------------------
    def forces_phase(self, dt1): self._call("forces_phase", self.h, C.c_double(dt1))
    def assemble(self): self._call("assemble", self.h)
    def advance(self, dt12, dt2): self._call("advance", self.h, C.c_double(dt12), C.c_double(dt2))
    def run_cycles(self, n): self._call("run_cycles", self.h, C.c_int(n))

    def synchronize(self):
        if self.status:
            self._call("synchronize", self.h)

    def time(self):
        out = (C.c_double * 5)(); iout = (C.c_int * 3)()
        self._call("get_time", self.h, out, iout)
        return dict(tt=out[0], dt1=out[1], dt2=out[2], dt12=out[3], dt2t=out[4],
                    neltst=iout[0], ityptst=iout[1], ncycle=iout[2])
