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


def midpoint_rule(npt):
    """The through-thickness rule of the reference's CUDA kernels (shell_strain_material_kernel.cu:696-701, 812-819): points at the
    layer centres zeta = -1/2 + (i + 1/2)/NPT, force weight 1/NPT, moment weight zeta/NPT."""
    z = np.array([-0.5 + (i + 0.5) / npt for i in range(npt)])
    return z, np.full(npt, 1.0 / npt), z / npt


def bent_plate(ipla, npt, seed=7, vs=1.0, warp=0.0):
    """Free BT / LAW2 plate with random translational AND rotational velocities: curvature, moments, rotational hourglass."""
    prop = meshgen.default_prop_shell(thick=1.2, ihbe=1, npt=npt, ipla=ipla, ismstr=2, ithk=0)
    m = meshgen.shell_plate(9, 7, 90.0, 70.0, law=2, prop=prop, pressure=0.0, clamp=False, zjitter=warp, vrand=0.0)
    rng = np.random.default_rng(seed)
    m.V[:, :2] = rng.uniform(-30.0, 30.0, (m.numnod, 2)) * vs
    m.V[:, 2] = rng.uniform(-20.0, 20.0, m.numnod) * vs
    m.VR[:] = rng.uniform(-4.0, 4.0, (m.numnod, 3)) * vs
    for g in m.shell_groups:
        g.mat.cc = 0.0
    return m
