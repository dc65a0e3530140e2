"""CPU suite: the oracle against GOLDEN VECTORS computed by the reference's own code.

tests/golden/refgpu_bt_law2_*.npz hold, per cycle of a seeded flat-plate case, the nodal arrays fed to -- and the nodal
forces / moments and time step returned by -- the reference's CUDA shell path (Belytschko-Tsay + LAW2), compiled unmodified
from /root/reference and run on a B200 (tests/golden/make_golden_refgpu.py; the cases are tests/test_ref_gpu_pin.py::plate).
The oracle replays the same cycles from the same initial model and must reproduce those forces to rounding (the reference
kernels contract fma and sum with atomics: 1e-12 of the largest nodal force)."""
import glob
import os
import re
import numpy as np
import pytest
from oracle.orc import Oracle
from refgpu_cases import plate

FILES = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refgpu_bt_law2_*.npz")))


def test_golden_files_are_committed():
    assert len(FILES) >= 4


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_oracle_reproduces_reference_gpu_forces(path):
    ipla, npt, rate, shear = map(int, re.search(r"ipla(\d)_npt(\d)_rate(\d)_shear(\d)", path).groups())
    gold = np.load(path)
    m = plate(ipla, npt, bool(rate), bool(shear))
    o = Oracle(m)
    dt1 = 0.0
    for c in range(len(gold["dt1"])):
        nd = o.download_nodes(("X", "V", "VR"))
        # the oracle's own trajectory is the one the reference path was fed (bitwise: same build, same seeds)
        assert np.array_equal(nd["X"], gold["X"][c]) and np.array_equal(nd["V"], gold["V"][c]) and dt1 == gold["dt1"][c]
        o.forces_phase(dt1); o.assemble()
        f = o.download_nodes(("A", "AR"))
        F = gold["F"][c]
        if c > 0:
            sf = np.abs(F[:, :3]).max()
            assert np.abs(f["A"] - F[:, :3]).max() <= 1e-12 * sf, (c, np.abs(f["A"] - F[:, :3]).max() / sf)
            if shear:
                sm = np.abs(F[:, 3:6]).max()
                assert np.abs(f["AR"] - F[:, 3:6]).max() <= 1e-12 * sm
        dt2 = o.time()["dt2t"]
        assert dt2 == pytest.approx(float(gold["dt_ref"][c]), rel=1e-12)
        o.advance(0.5 * (dt1 + dt2), dt2); dt1 = dt2
    assert o.shell_state("pla").max() > 1e-3
