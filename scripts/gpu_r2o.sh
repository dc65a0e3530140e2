python - <<'PY'
import os, time, numpy as np, torch
# the link: H2D alone, D2H alone, both at once (72 MB each)
n = 9_000_000
h1 = torch.empty(n, dtype=torch.float64).pin_memory(); h2 = torch.empty(n, dtype=torch.float64).pin_memory()
d1 = torch.empty(n, dtype=torch.float64, device="cuda"); d2 = torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, rep=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(rep): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / rep
def up():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
def down():
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
def both(): up(); down()
tu, td, tb = t(up), t(down), t(both)
print(f"H2D {72e-3/tu:.1f} GB/s  D2H {72e-3/td:.1f} GB/s  both at once: {tb*1e3:.3f} ms for 72 MB each way = {72e-3/tb:.1f} GB/s per direction", flush=True)
from openradioss_b200 import meshgen
from openradioss_b200.engine import Engine
os.environ["ORGPU_PIPE_DEBUG"] = "1"
m = meshgen.shell_plate(1000, 1000, 1000.0, 1000.0, pulse_tau=0.05, vwave=(60.0, 100.0))
m.fext = None; m.mext = None
nn = m.numnod
for K in (8,):
    os.environ["ORGPU_PIPE_CHUNKS"] = str(K)
    g = Engine(m); g.run_cycles(200); g.synchronize()
    nd = g.download_nodes(("X", "V", "VR"))
    h = [torch.from_numpy(np.ascontiguousarray(nd[k])).pin_memory() for k in ("X", "V", "VR")]
    F = torch.empty((nn, 8), dtype=torch.float64).pin_memory()
    dt1 = g.time()["dt2"]
    print("K", K, flush=True)
    for _ in range(3):
        g.forces_host(h[0].numpy(), h[1].numpy(), h[2].numpy(), dt1, F.numpy())
    del g
PY
