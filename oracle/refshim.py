"""ctypes front-end of oracle/_ref/libippl_refshim.so: pieces of the REAL reference headers compiled
in place from /root/reference (see ref_shim/refshim.cpp).  TEST INFRASTRUCTURE ONLY; available only
where the library has been built (`make -C oracle ref`, needs /root/reference)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libippl_refshim.so")
_lib = None


def available(try_build=True):
    if os.path.exists(_LIB_PATH):
        return True
    if try_build and os.path.isdir("/root/reference/src"):
        try:
            subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
        except Exception:
            return False
        return os.path.exists(_LIB_PATH)
    return False


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("reference shim not built (needs /root/reference)")
        _lib = C.CDLL(_LIB_PATH)
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _i3(v):
    return (C.c_int * 3)(*[int(t) for t in v])


def _d3(v):
    return (C.c_double * 3)(*[float(t) for t in v])


def scatter(mesh, x, y, z, q, rho):
    ex, ey, _ = mesh.ext
    lib().ref_scatter(C.c_long(len(x)), _p(x), _p(y), _p(z), _p(q), _d3(mesh.origin), _d3(mesh.h),
                      _i3(mesh.first), mesh.nghost, _p(rho), C.c_long(ex), C.c_long(ey))
    return rho


def gather(mesh, x, y, z, efield):
    ex, ey, _ = mesh.ext
    out = [np.zeros(len(x)) for _ in range(3)]
    lib().ref_gather(C.c_long(len(x)), _p(x), _p(y), _p(z), _d3(mesh.origin), _d3(mesh.h),
                     _i3(mesh.first), mesh.nghost, _p(efield), C.c_long(ex), C.c_long(ey),
                     _p(out[0]), _p(out[1]), _p(out[2]))
    return out


def periodic_bc(x, y, z, lo, hi):
    lib().ref_periodic_bc(C.c_long(len(x)), _p(x), _p(y), _p(z), _d3(lo), _d3(hi))


def partition(ng, nranks, parallel=(1, 1, 1)):
    boxes = np.zeros((nranks, 6), dtype=np.int32)
    rc = lib().ref_partition(_i3(ng), _i3(parallel), nranks, _p(boxes))
    if rc != 0:
        raise RuntimeError("reference Partitioner::split failed")
    return boxes


def neighbors(ng, nranks, my, nghost=1, periodic=True, parallel=(1, 1, 1)):
    boxes = np.zeros((nranks, 6), dtype=np.int32)
    out = np.zeros((512, 14), dtype=np.int32)
    n = lib().ref_neighbors(_i3(ng), _i3(parallel), nranks, int(periodic), nghost, my, _p(boxes), _p(out), 512)
    assert 0 <= n <= 512
    return boxes, out[:n].copy()


def matching_index(i):
    return lib().ref_matching_index(int(i))
