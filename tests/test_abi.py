"""The C-ABI library loads and exports every symbol include/orgpu.h declares (no compute call)."""
import ctypes, os, re
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "orgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(orgpu_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_entry_points():
    names = _declared()
    assert "orgpu_create" in names and "orgpu_run_cycles" in names and len(names) >= 25


def test_library_exports_every_declared_symbol():
    from openradioss_b200 import engine
    lib = engine.load_library()
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted("orgpu_" + n for n in engine.EXPORTS) == _declared()


def test_create_without_gpu_fails_loudly():
    """No CPU fallback: without a device orgpu_create returns an error, it does not compute."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from openradioss_b200 import engine, meshgen
    m = meshgen.hex_block(2, 2, 2, 1.0, 1.0, 1.0)
    with pytest.raises(RuntimeError, match="no CUDA device|failed"):
        engine.Engine(m)
