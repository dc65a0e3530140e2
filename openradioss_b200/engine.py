"""Product binding: the CUDA library ``csrc/liborgpu.so`` behind the C ABI of include/orgpu.h.

There is deliberately no fallback: if the library is not built (``__graft_entry__.build()`` or
``make -C openradioss_b200/csrc``) or no GPU is visible, creating an engine raises.
"""
from __future__ import annotations
import ctypes as C
import os
import numpy as np
from ._binding import Binding, _opt
from .model import Model

_LIB_PATH = os.environ.get("ORGPU_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "liborgpu.so")
_lib = None

EXPORTS = """create destroy last_error upload_nodes set_loads set_bcs set_solids set_shells set_pon
set_functions add_solid_group add_solid_group_law add_shell_group set_sh3n add_sh3n_group download_sh3n_state upload_sh3n_state finalize forces_phase assemble advance run_cycles
synchronize get_time download_nodes download_fsky download_solid_state download_shell_state
step_host launch_count last_run_ms set_profile get_profile pack_rows unpack_rows comm_unique_id comm_init
set_exchange exchange get_energies p2p_export p2p_connect set_load_function set_fixvel set_gravity upload_solid_state upload_shell_state set_time set_itab
set_parts set_print get_balance get_balance_history set_quadrature set_global_order set_exchange_timeout pack_nodes add_nodes set_exchange_nodes step_host_rot forces_host set_shell_group_fail set_cloads set_solid_group_fail set_parith""".split()


def load_library() -> C.CDLL:
    """dlopen liborgpu.so (no compute call, works without a GPU).  Raises if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(f"{_LIB_PATH} is not built: run __graft_entry__.build() "
                               "(nvcc, sm_100a).  There is no CPU fallback for this path.")
        _lib = C.CDLL(_LIB_PATH)
    return _lib


class Engine(Binding):
    """One device-resident model (one domain / one GPU)."""

    def __init__(self, model: Model = None, device: int = 0):
        super().__init__(load_library(), "orgpu_", True)
        if model is not None:
            self.load(model, device)

    def step_host(self, X, V, VR, ncycles, Xout, Vout, VRout=None):
        """End-to-end entry: host nodal arrays in, ncycles on the device, host arrays out."""
        if VRout is None:
            self._call("step_host", self.h, _opt(X, np.float64), _opt(V, np.float64), _opt(VR, np.float64),
                       C.c_int(ncycles), Xout.ctypes.data_as(C.c_void_p), Vout.ctypes.data_as(C.c_void_p))
        else:
            self._call("step_host_rot", self.h, _opt(X, np.float64), _opt(V, np.float64), _opt(VR, np.float64),
                       C.c_int(ncycles), Xout.ctypes.data_as(C.c_void_p), Vout.ctypes.data_as(C.c_void_p), VRout.ctypes.data_as(C.c_void_p))

    def forces_host(self, X, V, VR, dt1, F8):
        """The reference -gpu path's cycle in one pipelined call: host X, V, VR (n,3) in, internal nodal forces F8 (n,8) =
        Fx,Fy,Fz,Mx,My,Mz,STIFN,STIFR out; returns (dt2t, neltst, ityptst).  Pinned host arrays overlap the two directions."""
        dt2t = C.c_double(0.0); nel = C.c_int(0); ityp = C.c_int(0)
        self._call("forces_host", self.h, _opt(X, np.float64), _opt(V, np.float64), _opt(VR, np.float64), C.c_double(dt1),
                   F8.ctypes.data_as(C.c_void_p), C.byref(dt2t), C.byref(nel), C.byref(ityp))
        return dt2t.value, nel.value, ityp.value

    # -- one process per GPU: NCCL exchange inside run_cycles ------------------------------------
    def comm_init(self, dist, domain, p2p=True, parith_off=False):
        """Create the NCCL communicator (id from rank 0, broadcast through torch.distributed) and
        register the neighbour send / receive slot lists of this rank's Domain.  With p2p (default, ranks of
        one NVLink node) the receive windows are then exchanged through CUDA IPC and run_cycles uses the
        library's own peer-memory exchange kernels instead of NCCL."""
        import torch
        rank, world = dist.get_rank(), dist.get_world_size()
        uid = (C.c_ubyte * 128)()
        if rank == 0:
            self._call("comm_unique_id", uid)
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.tensor(list(uid), dtype=torch.uint8, device=dev)
        dist.broadcast(t, 0)
        uid = (C.c_ubyte * 128)(*t.cpu().tolist())
        self._call("comm_init", self.h, C.c_int(world), C.c_int(rank), uid)
        if os.environ.get("ORGPU_P2P_TIMEOUT_S"):
            self._call("set_exchange_timeout", self.h, C.c_double(float(os.environ["ORGPU_P2P_TIMEOUT_S"])))
        if parith_off:
            # /PARITH/OFF (SPMD_EXCH_A): partial sums of the frontier nodes over NCCL, neighbours in ascending rank order
            nbs = sorted([nb for nb in domain.neighbors if nb.nodes is not None and len(nb.nodes)], key=lambda n: n.rank)
            ranks = np.array([nb.rank for nb in nbs], np.int32)
            ptr = np.zeros(len(nbs) + 1, np.int32)
            for k, nb in enumerate(nbs):
                ptr[k + 1] = ptr[k] + len(nb.nodes)
            nodes = np.concatenate([nb.nodes for nb in nbs]).astype(np.int32) if nbs else np.zeros(0, np.int32)
            self._call("set_exchange_nodes", self.h, C.c_int(len(nbs)), _opt(ranks, np.int32), _opt(ptr, np.int32), _opt(nodes, np.int32))
            return
        nbs = domain.neighbors
        ranks = np.array([nb.rank for nb in nbs], np.int32)
        sp = np.zeros(len(nbs) + 1, np.int32); rp = np.zeros(len(nbs) + 1, np.int32)
        for k, nb in enumerate(nbs):
            sp[k + 1] = sp[k] + len(nb.send); rp[k + 1] = rp[k] + len(nb.recv)
        ss = np.concatenate([nb.send for nb in nbs]).astype(np.int32) if nbs else np.zeros(0, np.int32)
        rs = np.concatenate([nb.recv for nb in nbs]).astype(np.int32) if nbs else np.zeros(0, np.int32)
        self._call("set_exchange", self.h, C.c_int(len(nbs)), _opt(ranks, np.int32), _opt(sp, np.int32), _opt(ss, np.int32),
                   _opt(rp, np.int32), _opt(rs, np.int32))
        if p2p and os.environ.get("ORGPU_NO_P2P", "0") != "1":
            h = (C.c_ubyte * 64)()
            self._call("p2p_export", self.h, h)
            mine = torch.tensor(list(h), dtype=torch.uint8, device=dev)
            allh = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allh, mine)                       # also orders "window initialised" before "peers read it"
            flat = np.concatenate([t_.cpu().numpy() for t_ in allh]).astype(np.uint8)
            self._call("p2p_connect", self.h, flat.ctypes.data_as(C.c_void_p))
            dist.barrier()

    def exchange(self): self._call("exchange", self.h)

    def checkpoint(self):
        """Everything a restart needs: nodal arrays, clock, element state of both families."""
        self.synchronize()
        ck = dict(nodes=self.download_nodes(("X", "V", "VR", "D")), time=self.time())
        if self.model.numels:
            ck["solid"] = {f: self.solid_state(f) for f in ("sig", "eint", "rho", "qvis", "pla", "epsd", "off", "temp", "smstr")}
        therm = lambda groups: any(g.law == 2 and getattr(g.mat, "has_temp", 0) for g in groups)
        if self.model.numelc:
            ck["shell"] = {f: self.shell_state(f) for f in ("forc", "mom", "eint", "thk", "off", "stra", "epsd", "hourg", "smstr", "sig", "pla", "epsd_ip")}
            if therm(self.model.shell_groups):
                ck["shell"]["temp"] = self.shell_state("temp")          # per-point temperature of thermal Johnson-Cook shells
            if any(getattr(g.mat, "fisokin", 0.0) > 0.0 for g in self.model.shell_groups):
                ck["shell"]["sigb"] = self.shell_state("sigb")          # back stress of the kinematic hardening
            if any(g.law == 36 and getattr(g.mat, "vp", 0) == 1 for g in self.model.shell_groups):
                ck["shell"]["plap"] = self.shell_state("plap")          # LAW36 VP = 1: filtered plastic strain rate of the points (UVAR(2))
            if any(getattr(g, "fail", None) is not None for g in self.model.shell_groups):
                ck["shell"].update(dfmax=self.shell_state("dfmax"), foff=self.shell_state("foff"))     # /FAIL/JOHNSON damage and point flags
        if self.model.numeltg:
            ck["sh3n"] = {f: self.sh3n_state(f) for f in ("forc", "mom", "eint", "thk", "off", "stra", "epsd", "smstr", "sig", "pla", "epsd_ip")}
            if therm(self.model.sh3n_groups):
                ck["sh3n"]["temp"] = self.sh3n_state("temp")
            if any(getattr(g.mat, "fisokin", 0.0) > 0.0 for g in self.model.sh3n_groups):
                ck["sh3n"]["sigb"] = self.sh3n_state("sigb")
            if any(g.law == 36 and getattr(g.mat, "vp", 0) == 1 for g in self.model.sh3n_groups):
                ck["sh3n"]["plap"] = self.sh3n_state("plap")
            if any(getattr(g, "fail", None) is not None for g in self.model.sh3n_groups):
                ck["sh3n"].update(dfmax=self.sh3n_state("dfmax"), foff=self.sh3n_state("foff"))
        if self.model.numels and any(getattr(g, "law", 2) == 2 and getattr(g.mat, "fisokin", 0.0) > 0.0 for g in self.model.solid_groups):
            ck["solid"]["sigb"] = self.solid_state("sigb")             # back stress of the kinematic hardening (LBUF%SIGB)
        if self.model.numels and any(getattr(g, "fail", None) is not None for g in self.model.solid_groups):
            ck["solid"]["dfmax"] = self.solid_state("dfmax")           # /FAIL/JOHNSON damage
        if self.model.numels and any(getattr(g, "law", 2) == 36 for g in self.model.solid_groups):
            ck["solid"].update({f: self.solid_state(f) for f in ("wpla", "stra")})
        return ck

    def restore(self, ck):
        n = ck["nodes"]
        self.upload_nodes(X=n["X"], V=n["V"], VR=n["VR"], D=n["D"])
        t = ck["time"]
        self.set_time(t["tt"], t["dt2"], t["dt2"], t["ncycle"])       # DT2OLD = DT2 after a completed cycle (resol.F:6494)
        for f, a in ck.get("solid", {}).items():
            if f == "temp" and not any(getattr(g.mat, "has_temp", 0) for g in self.model.solid_groups):
                continue
            self.upload_solid_state(f, a)
        for f, a in ck.get("shell", {}).items():
            self.upload_shell_state(f, a)
        for f, a in ck.get("sh3n", {}).items():
            self.upload_sh3n_state(f, a)

    def energies(self):
        """(internal solids, internal shells, kinetic translation, kinetic rotation), summed on the device."""
        out = (C.c_double * 4)()
        self._call("get_energies", self.h, out)
        return tuple(out)

    def launch_count(self) -> int:
        fn = self.lib.orgpu_launch_count; fn.restype = C.c_longlong
        return int(fn(self.h))

    def last_run_ms(self) -> float:
        fn = self.lib.orgpu_last_run_ms; fn.restype = C.c_double
        return float(fn(self.h))

    def set_profile(self, on: bool): self._call("set_profile", self.h, C.c_int(1 if on else 0))

    def profile(self, cls: int):
        ms = C.c_double(); n = C.c_longlong()
        self._call("get_profile", self.h, C.c_int(cls), C.byref(ms), C.byref(n))
        return ms.value, n.value
