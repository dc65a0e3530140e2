python - <<'PY'
import time, torch
n = 9_000_000
h1 = torch.empty(n, dtype=torch.float64).pin_memory(); h2 = torch.empty(n, dtype=torch.float64).pin_memory()
d1 = torch.empty(n, dtype=torch.float64, device="cuda"); d2 = torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, rep=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(rep): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / rep
for pieces_up, pieces_dn in ((1, 1), (8, 8), (24, 8), (48, 16), (96, 32)):
    cu = n // pieces_up; cd = n // pieces_dn
    def up():
        with torch.cuda.stream(s1):
            for i in range(pieces_up): d1[i*cu:(i+1)*cu].copy_(h1[i*cu:(i+1)*cu], non_blocking=True)
    def down():
        with torch.cuda.stream(s2):
            for i in range(pieces_dn): h2[i*cd:(i+1)*cd].copy_(d2[i*cd:(i+1)*cd], non_blocking=True)
    def both(): up(); down()
    tu, td, tb = t(up), t(down), t(both)
    print(f"pieces up {pieces_up} down {pieces_dn}: H2D alone {72e-3/tu:.1f} GB/s, D2H alone {72e-3/td:.1f} GB/s, together {tb*1e3:.3f} ms = {72e-3/tb:.1f} GB/s each way", flush=True)
PY
