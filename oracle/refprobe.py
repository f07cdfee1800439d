"""ctypes front-end of oracle/_ref/libclpp_probe.so -- TEST INFRASTRUCTURE, not product code.

The probe (oracle/probe.cpp) wraps the UNMODIFIED reference CLASS++ built from
/root/reference by oracle/Makefile.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference arm may import this module.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")
_LIB = None

LEVELS = {"background": 0, "thermodynamics": 1, "perturb": 2, "primordial": 3, "nonlinear": 4,
          "transfer": 5, "spectra": 6, "lensing": 7}


def available():
    return os.path.exists(os.path.join(_REF, "libclpp_probe.so"))


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(os.path.join(_REF, "libclpp_probe.so"))
        lib.rp_create.restype = ctypes.c_void_p
        lib.rp_create.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
        lib.rp_destroy.argtypes = [ctypes.c_void_p]
        lib.rp_compute.restype = ctypes.c_int
        lib.rp_compute.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p]
        lib.rp_get.restype = ctypes.c_long
        lib.rp_get.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_long]
        _LIB = lib
    return _LIB


class RefCosmology:
    """One reference `Cosmology` object (source/cosmology.h) built from an .ini-style dict."""

    def __init__(self, params, threads=None):
        params = dict(params)
        # run-time data files (BBN table) are copied to oracle/_ref/ by the Makefile
        params.setdefault("class_dir", _REF)
        params.setdefault("sBBN file", os.path.join(_REF, "bbn", "sBBN_2017.dat"))
        if threads is not None:
            params["threads"] = int(threads)
        text = "".join("%s = %s\n" % (k, v) for k, v in params.items())
        self._err = ctypes.create_string_buffer(2048)
        self._h = _lib().rp_create(text.encode(), self._err)
        if not self._h:
            raise RuntimeError("reference input error: " + self._err.value.decode(errors="replace"))

    def compute(self, level="lensing"):
        rc = _lib().rp_compute(self._h, LEVELS[level], self._err)
        if rc != 0:
            raise RuntimeError("reference compute error: " + self._err.value.decode(errors="replace"))
        return self

    def get(self, name):
        n = _lib().rp_get(self._h, name.encode(), None, 0)
        if n < 0:
            raise KeyError(name)
        out = np.empty(n, dtype=np.float64)
        if n:
            _lib().rp_get(self._h, name.encode(), out.ctypes.data_as(ctypes.c_void_p), n)
        return out

    def scalar(self, name):
        return float(self.get(name)[0])

    def iscalar(self, name):
        return int(round(self.scalar(name)))

    def close(self):
        if self._h:
            _lib().rp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
