#!/usr/bin/env python3
"""Generates tests/golden/refgpu_bt_law2_*.npz ON A GPU BOX (gpurun): nodal forces computed by the reference's own CUDA
shell path (oracle/_ref/libshellgpu_ref.so, built unmodified from /root/reference by `make -C oracle refgpu`) for seeded
flat-plate cases, together with the nodal arrays it was fed.  The CPU suite (tests/test_golden_refgpu.py) replays the same
cases through the oracle.  Usage: gpurun -- 'python tests/golden/make_golden_refgpu.py'  then copy gpurun_out/golden/* to tests/golden/."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle.orc import Oracle
from oracle import refgpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
from refgpu_cases import plate          # the seeded cases of the GPU pin test

out = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "gpurun_out", "golden")
os.makedirs(out, exist_ok=True)
for ipla, npt, rate, shear in [(0, 3, True, False), (1, 5, True, True), (2, 3, False, True), (1, 3, False, False)]:
    m = plate(ipla, npt, rate, shear)
    o, r = Oracle(m), refgpu.RefShellGPU(m)
    dt1 = 0.0; rec = dict(X=[], V=[], VR=[], dt1=[], F=[], dt_ref=[])
    for c in range(8):
        nd = o.download_nodes(("X", "V", "VR"))
        fr = r.step(dt1, nd["X"], nd["V"], nd["VR"])
        rec["X"].append(nd["X"]); rec["V"].append(nd["V"]); rec["VR"].append(nd["VR"]); rec["dt1"].append(dt1); rec["F"].append(fr[:, :6])
        rec["dt_ref"].append(r.min_dt(m.control.dtfac_shell))
        o.forces_phase(dt1); o.assemble()
        dt2 = o.time()["dt2t"]; o.advance(0.5 * (dt1 + dt2), dt2); dt1 = dt2
    name = f"refgpu_bt_law2_ipla{ipla}_npt{npt}_rate{int(rate)}_shear{int(shear)}.npz"
    np.savez_compressed(os.path.join(out, name), **{k: np.asarray(v) for k, v in rec.items()})
    print("wrote", name, "max |F|", np.abs(rec["F"][-1]).max())
