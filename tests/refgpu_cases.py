"""Seeded flat-plate cases shared by the reference-GPU pin test, the golden-vector generator and the CPU golden test."""
import numpy as np
from openradioss_b200 import meshgen


def plate(ipla, npt, rate, shear, seed=5, vs=1.0):
    prop = meshgen.default_prop_shell(thick=1.2, ihbe=1, npt=npt, ipla=ipla, ismstr=2, ithk=0)   # the reference GPU path: Ismstr 1, 2, 11; Ishell <= 1
    m = meshgen.shell_plate(9, 7, 90.0, 70.0, law=2, prop=prop, pressure=0.0, clamp=False, zjitter=0.0, vrand=0.0)
    rng = np.random.default_rng(seed)
    m.V[:, :2] = rng.uniform(-60.0, 60.0, (m.numnod, 2)) * vs    # strong enough to yield within a few cycles
    m.VR[:] = 0.0
    if shear:
        m.V[:, 2] = rng.uniform(-20.0, 20.0, m.numnod) * vs
        m.IN[:] = 1.0e30                                         # rotations stay zero: no curvature ever develops
    for g in m.shell_groups:
        if rate:
            g.mat.israte = 1; g.mat.asrate = 1.0e30              # filter coefficient min(1, ASRATE*DT1) = 1 on both sides
        else:
            g.mat.cc = 0.0
    return m
