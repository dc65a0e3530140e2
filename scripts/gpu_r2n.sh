mkdir -p gpurun_out
python -m pytest tests/test_forces_host_gpu.py -m gpu -q -x 2>&1 | tail -25 > gpurun_out/pytest_r2n.log
grep -E "^E  |FAILED|passed|failed|Error" gpurun_out/pytest_r2n.log | cut -c1-300 | head -20
python - <<'PY'
import os, time, numpy as np, torch
from openradioss_b200 import meshgen
from openradioss_b200.engine import Engine
m = meshgen.shell_plate(1000, 1000, 1000.0, 1000.0, pulse_tau=0.05, vwave=(60.0, 100.0))
m.fext = None; m.mext = None
n = m.numnod
for K in (1, 4, 8, 12, 16, 24, 32, 48):
    os.environ["ORGPU_PIPE_CHUNKS"] = str(K)
    g = Engine(m)
    g.run_cycles(200); g.synchronize()
    nd = g.download_nodes(("X", "V", "VR"))
    h = [torch.from_numpy(np.ascontiguousarray(nd[k])).pin_memory() for k in ("X", "V", "VR")]
    F = torch.empty((n, 8), dtype=torch.float64).pin_memory()
    dt1 = g.time()["dt2"]
    for _ in range(3):
        g.forces_host(h[0].numpy(), h[1].numpy(), h[2].numpy(), dt1, F.numpy())
    t0 = time.perf_counter()
    for _ in range(30):
        g.forces_host(h[0].numpy(), h[1].numpy(), h[2].numpy(), dt1, F.numpy())
    dt = (time.perf_counter() - t0) / 30
    print(f"K={K} forces_host {dt*1e3:.3f} ms/step  {m.numelc/dt/1e9:.3f} G el-cyc/s", flush=True)
    del g
PY
