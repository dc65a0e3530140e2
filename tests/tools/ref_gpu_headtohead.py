#!/usr/bin/env python3
"""Head to head on one B200: the reference's own CUDA shell path (oracle/_ref/libshellgpu_ref.so, built unmodified from
/root/reference: three kernels per cycle, atomics, nodal arrays re-uploaded and forces downloaded every cycle) against
liborgpu on the same Belytschko-Tsay / LAW2 / NPT=5 plate.  Usage (gpurun): python tests/tools/ref_gpu_headtohead.py [nx ny]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from openradioss_b200 import meshgen
from openradioss_b200.engine import Engine
from oracle import refgpu

nx, ny = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1000, 1000)
prop = meshgen.default_prop_shell(thick=2.0, ihbe=1, npt=5, ipla=1, ismstr=2, ithk=1)
m = meshgen.shell_plate(nx, ny, float(nx), float(ny), law=2, prop=prop, pressure=1.0, vrand=2.0)
ne = m.numelc
res = {"elements": ne, "nodes": m.numnod, "case": "Belytschko-Tsay Ishell=1, LAW2 Johnson-Cook, NPT=5, Iplas=1, Ismstr=2"}

g = Engine(m)
g.run_cycles(20); g.synchronize()
g.run_cycles(100); g.synchronize()
res["orgpu_ms_per_cycle"] = g.last_run_ms() / 100
g.set_profile(True); g.run_cycles(50); g.synchronize()
res["orgpu_kernel_ms"] = {k: (g.profile(i)[0] / max(g.profile(i)[1], 1)) for i, k in enumerate(("brick", "shell_forces", "node"))}
g.set_profile(False)
import torch
n = m.numnod
hX = torch.empty((n, 3), dtype=torch.float64).pin_memory(); hV = torch.empty((n, 3), dtype=torch.float64).pin_memory()
oX = torch.empty((n, 3), dtype=torch.float64).pin_memory(); oV = torch.empty((n, 3), dtype=torch.float64).pin_memory()
nd = g.download_nodes(("X", "V")); hX.numpy()[:] = nd["X"]; hV.numpy()[:] = nd["V"]
for _ in range(3):
    g.step_host(hX.numpy(), hV.numpy(), None, 1, oX.numpy(), oV.numpy())
t0 = time.perf_counter()
for _ in range(30):
    g.step_host(hX.numpy(), hV.numpy(), None, 1, oX.numpy(), oV.numpy())
res["orgpu_e2e_ms_per_cycle"] = (time.perf_counter() - t0) / 30 * 1e3

r = refgpu.RefShellGPU(m)
nd = g.download_nodes(("X", "V", "VR"))
dt = g.time()["dt2"]
pX, pV, pVR = r.pin(nd["X"], nd["V"], nd["VR"])                               # page-locked in place, as the Engine does
for _ in range(3):
    r.step(dt, pX, pV, pVR)
t0 = time.perf_counter()
for _ in range(30):
    r.step(dt, pX, pV, pVR)
res["ref_e2e_with_numpy_ms_per_cycle"] = (time.perf_counter() - t0) / 30 * 1e3   # includes the harness's numpy transposes: NOT a library time (see same_abi_*)
r.synchronize()
t0 = time.perf_counter()
for _ in range(100):
    r.run_kernels_only(dt)
r.synchronize()
res["ref_kernels_ms_per_cycle"] = (time.perf_counter() - t0) / 100 * 1e3    # the three kernels only (forces + atomics; no nodal update)
# the SAME ABI and the SAME driver calls, two libraries: liborgpu.so exports the reference's shell_gpu_* entry points too
ORGPU = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "openradioss_b200", "csrc", "liborgpu.so")
d = refgpu.RefShellGPU(m, lib=ORGPU, asrate=m.shell_groups[0].mat.asrate)
dX, dV, dVR = d.pin(nd["X"], nd["V"], nd["VR"])
for name, lib, arrs in (("ref", r, (pX, pV, pVR)), ("orgpu", d, (dX, dV, dVR))):
    for _ in range(3):
        lib.step_raw(dt, *arrs)
    t0 = time.perf_counter()
    for _ in range(30):
        lib.step_raw(dt, *arrs)
    res["same_abi_%s_ms_per_cycle" % name] = (time.perf_counter() - t0) / 30 * 1e3   # upload X,V,VR + forces + download 8N, no numpy around
res["same_abi_speedup"] = res["same_abi_ref_ms_per_cycle"] / res["same_abi_orgpu_ms_per_cycle"]
res["speedup_device_forces_only"] = res["ref_kernels_ms_per_cycle"] / res["orgpu_kernel_ms"]["shell_forces"]
res["speedup_device_cycle_vs_ref_forces_only"] = res["ref_kernels_ms_per_cycle"] / res["orgpu_ms_per_cycle"]
res["speedup_e2e_host_arrays"] = res["same_abi_ref_ms_per_cycle"] / res["orgpu_e2e_ms_per_cycle"]          # reference ABI (X,V,VR up, 8N down) vs orgpu_step_host (X,V up and down)
res["speedup_resident_cycle_vs_ref_e2e"] = res["same_abi_ref_ms_per_cycle"] / res["orgpu_ms_per_cycle"]   # what keeping the nodal arrays on the device buys
print(json.dumps(res))
