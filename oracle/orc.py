"""TEST INFRASTRUCTURE ONLY: ctypes binding of the CPU oracle (oracle/liborc.so).

Imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs -- never by the
product package."""
import ctypes as C
import os
import subprocess
from openradioss_b200._binding import Binding

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_DIR, "liborc.so")
_lib = None


def build():
    subprocess.run(["make", "-C", _DIR], check=True, capture_output=True)


def load_library():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = C.CDLL(_LIB)
    return _lib


class Oracle(Binding):
    def __init__(self, model=None, threads: int = 1):
        super().__init__(load_library(), "orc_", False)
        if model is not None:
            self.load(model)
            self.set_threads(threads)

    def set_threads(self, n):
        self.lib.orc_set_threads(self.h, C.c_int(n))
