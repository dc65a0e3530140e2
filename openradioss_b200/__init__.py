"""openradioss_b200 -- B200-native explicit element cycle for the OpenRadioss Engine.

Host-side mirror of the reference's GPU glue for ONE hot path:
internal forces of 8-node bricks (SFORC3) and 4-node shells (CFORC3 / CZFORC3) with LAW2 /
LAW36, the element time-step argmin, /PARITH/ON assembly (ASSPAR4) and the central-difference
nodal update (ACCELE / VELOCITY / DEPLA).  All compute lives in the C-ABI CUDA library
``openradioss_b200/csrc/liborgpu.so`` (include/orgpu.h); this package only builds models
(the Starter's job in the reference) and binds the library through ctypes.  There is no CPU
fallback: importing :mod:`openradioss_b200.engine` without the built library raises.
"""
from .model import (Law2, Law36, PropSolid, PropShell, Control, Model, SolidGroup, ShellGroup)  # noqa: F401
