import sys; sys.path.insert(0, '.')
import numpy as np
from openradioss_b200 import meshgen
from openradioss_b200.engine import Engine
from oracle.orc import Oracle
from oracle import refgpu
def run(tag, **kw):
    prop = meshgen.default_prop_shell(thick=1.2, ihbe=1, npt=3, ipla=kw.get('ipla', 0), ismstr=2, ithk=0)
    for k in ('dm','h1','h2','h3'):
        if k in kw: setattr(prop, k, kw[k])
    m = meshgen.shell_plate(9, 7, 90.0, 70.0, law=2, prop=prop, pressure=0.0, clamp=False, zjitter=0.0, vrand=0.0, jitter=kw.get('jitter', 0.05))
    rng = np.random.default_rng(5)
    m.V[:, :2] = rng.uniform(-60.0, 60.0, (m.numnod, 2)) * kw.get('vs', 1.0); m.VR[:] = 0
    if 'israte' in kw:
        for g_ in m.shell_groups: g_.mat.israte = 1; g_.mat.asrate = 1.0e30
    if 'cc' in kw:
        for g_ in m.shell_groups: g_.mat.cc = kw['cc']
    o, r = Oracle(m), refgpu.RefShellGPU(m)
    dt1 = 0.0; errs = []
    for c in range(4):
        nd = o.download_nodes(("X","V","VR"))
        fr = r.step(dt1, nd["X"], nd["V"], nd["VR"])
        o.forces_phase(dt1); o.assemble()
        fo = o.download_nodes(("A",))["A"]
        s = np.abs(fo).max()
        errs.append(np.abs(fr[:, :3]-fo).max()/s if s>0 else 0.0)
        dt2 = o.time()["dt2t"]; o.advance(0.5*(dt1+dt2), dt2); dt1 = dt2
    print(tag, ['%.2e'%e for e in errs], 'pla', o.shell_state('pla').max(), flush=True)
for ipla in (0,1,2):
    for vs in (0.2, 1.0):
        run(f'ipla{ipla} cc!=0 israte=1 asrate=inf vs={vs}', ipla=ipla, vs=vs, israte=1)
