"""Named fp64 constants of the reference (common_source/modules/constant_mod.F), read from the
generated header include/or_constants.h so the host-side model builder (which plays the
Starter's role) uses the same bits as the kernels."""
import os
import re

_HDR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "or_constants.h")
K = {}
for _m in re.finditer(r"#define K_(\w+) \(([^)]+)\)", open(_HDR).read()):
    K[_m.group(1)] = float.fromhex(_m.group(2))
