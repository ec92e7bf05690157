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


def neighbors_boxes(ng, boxes, my, nghost=1, periodic=True):
    """neighbour tables of rank `my` after FieldLayout::updateLayout(boxes) (an ORB layout)"""
    b = np.ascontiguousarray(boxes, dtype=np.int32)
    out = np.zeros((512, 14), dtype=np.int32)
    n = lib().ref_neighbors_boxes(_i3(ng), b.shape[0], int(periodic), nghost, my, _p(b), _p(out), 512)
    assert 0 <= n <= 512
    return out[:n].copy()


def matching_index(i):
    return lib().ref_matching_index(int(i))


# ---- sampling path of the reference (oracle/_ref/libippl_refshim_random.so, ref_shim/refshim_random.cpp) ----------------
_RLIB_PATH = os.path.join(_HERE, "_ref", "libippl_refshim_random.so")
_rlib = None


def random_available(try_build=True):
    if os.path.exists(_RLIB_PATH):
        return True
    if try_build and os.path.isdir("/root/reference/src"):
        try:
            subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
        except Exception:
            return False
        return os.path.exists(_RLIB_PATH)
    return False


def rlib():
    global _rlib
    if _rlib is None:
        if not random_available():
            raise RuntimeError("reference sampling shim not built (needs /root/reference)")
        _rlib = C.CDLL(_RLIB_PATH)
        _rlib.refrand_eval.restype = C.c_double
        _rlib.refrand_full_pdf.restype = C.c_double
        _rlib.refrand_newton.restype = C.c_double
    return _rlib


def _par(par):
    return (C.c_double * 6)(*[float(p) for p in par])


def rand_eval(kind, par, which, d, x, aux=0.0):
    """kind 1 cosine functors / 2 NormalDistribution; which 0 cdf, 1 pdf, 2 estimate, 3 objective(x, u=aux), 4 its derivative"""
    return rlib().refrand_eval(kind, _par(par), which, d, C.c_double(x), C.c_double(aux))


def rand_full_pdf(kind, par, x):
    return rlib().refrand_full_pdf(kind, _par(par), _d3(x))


def rand_newton(kind, par, d, x0, u):
    return rlib().refrand_newton(kind, _par(par), d, C.c_double(x0), C.c_double(u))


def rand_sampling(kind, par, rmin, rmax, regions, ntotal, gen_rank=-1, u01=None):
    """The reference's InverseTransformSampling constructor for every rank -> (nlocal[nranks], ubounds[nranks][6]) and,
    for gen_rank >= 0, generate() with the uniforms u01[3][nlocal] replayed -> x[3][nlocal]."""
    reg = np.ascontiguousarray(regions, dtype=np.float64)
    nr = reg.shape[0]
    nloc = (C.c_long * nr)()
    ub = np.zeros((nr, 6))
    out = None
    if gen_rank >= 0:
        u01 = np.ascontiguousarray(u01, dtype=np.float64)
        out = np.zeros_like(u01)
    rlib().refrand_sampling(kind, _par(par), _d3(rmin), _d3(rmax), _p(reg), nr, C.c_long(ntotal), nloc, _p(ub), gen_rank,
                            _p(u01) if out is not None else None, _p(out) if out is not None else None)
    return list(nloc), ub, out


def rand_randn(mu, sd, g):
    """randn::operator() with the standard normals g[n][3] replayed -> v[n][3]"""
    g = np.ascontiguousarray(g, dtype=np.float64)
    out = np.zeros_like(g)
    rlib().refrand_randn(_d3(mu), _d3(sd), C.c_long(g.shape[0]), _p(g), _p(out))
    return out


# ---- the reference's OrthogonalRecursiveBisection (oracle/_ref/libippl_refshim_orb.so, ref_shim/refshim_orb.cpp) --------
_OLIB_PATH = os.path.join(_HERE, "_ref", "libippl_refshim_orb.so")
_olib = None


def orb_available(try_build=True):
    if os.path.exists(_OLIB_PATH):
        return True
    if try_build and os.path.isdir("/root/reference/src"):
        try:
            subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
        except Exception:
            return False
        return os.path.exists(_OLIB_PATH)
    return False


def olib():
    global _olib
    if _olib is None:
        if not orb_available():
            raise RuntimeError("reference ORB shim not built (needs /root/reference)")
        _olib = C.CDLL(_OLIB_PATH)
    return _olib


def orb_find_median(w):
    a = np.ascontiguousarray(w, dtype=np.float64)
    return int(olib().reforb_find_median(_p(a), len(a)))


def orb_repartition(ng, nranks, weight):
    """binaryRepartition of the reference on the global interior weights weight[z][y][x] -> (boxes[nranks][6], ok)"""
    w = np.ascontiguousarray(weight, dtype=np.float64)
    boxes = np.zeros((nranks, 6), dtype=np.int32)
    ok = olib().reforb_repartition(_i3(ng), nranks, _p(w), _p(boxes))
    return boxes, bool(ok)


# ---- the reference's in-rank periodic halo (oracle/_ref/libippl_refshim_halo.so, ref_shim/refshim_halo.cpp) -------------
_HLIB_PATH = os.path.join(_HERE, "_ref", "libippl_refshim_halo.so")
_hlib = None


def halo_available(try_build=True):
    if os.path.exists(_HLIB_PATH):
        return True
    if try_build and os.path.isdir("/root/reference/src"):
        try:
            subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
        except Exception:
            return False
        return os.path.exists(_HLIB_PATH)
    return False


def halo_exchange(ng, boxes, fields, ncomp, mode, nghost=1, use_boxes=False):
    """BareField::fillHalo / accumulateHalo of the reference for EVERY rank at once: HaloCells::exchangeBoundaries (pack,
    isend / recv over an in-process mailbox, unpack) on one thread per rank, then applyPeriodicSerialDim.  fields[r]:
    rank r's ghosted array (ncomp doubles per cell), modified in place.  use_boxes: apply `boxes` with
    FieldLayout::updateLayout (an ORB layout) instead of the default partition."""
    global _hlib
    if _hlib is None:
        if not halo_available():
            raise RuntimeError("reference halo shim not built (needs /root/reference)")
        _hlib = C.CDLL(_HLIB_PATH)
    nr = len(fields)
    arr = (C.c_void_p * nr)(*[f.ctypes.data for f in fields])
    b = np.ascontiguousarray(boxes, dtype=np.int32) if use_boxes else None
    _hlib.refhalo_exchange(_i3(ng), nr, _p(b) if b is not None else None, nghost, ncomp, 0 if mode == "fill" else 1, arr)
    return fields


def halo_periodic(field, ng, mode, nghost=1):
    """HaloCells::applyPeriodicSerialDim<assign | rhs_plus_assign> of the reference, in place, on a ghosted scalar field
    (x fastest) of a single-rank all-periodic layout; mode "fill" or "accumulate"."""
    global _hlib
    if _hlib is None:
        if not halo_available():
            raise RuntimeError("reference halo shim not built (needs /root/reference)")
        _hlib = C.CDLL(_HLIB_PATH)
    _hlib.refhalo_periodic(_i3(ng), nghost, 0 if mode == "fill" else 1, _p(field))
    return field


# ---- the reference's RegionLayout (oracle/_ref/libippl_refshim_region.so, ref_shim/refshim_region.cpp) -------------------
_GLIB_PATH = os.path.join(_HERE, "_ref", "libippl_refshim_region.so")
_glib = None


def region_available(try_build=True):
    if os.path.exists(_GLIB_PATH):
        return True
    if try_build and os.path.isdir("/root/reference/src"):
        try:
            subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
        except Exception:
            return False
        return os.path.exists(_GLIB_PATH)
    return False


def regions(ng, nranks, origin, h, boxes=None):
    """detail::RegionLayout(FieldLayout, UniformCartesian) of the reference -> regions[nranks][6] = min[3], max[3]; with
    `boxes` ([nranks][6] lo, hi inclusive) after FieldLayout::updateLayout(boxes)"""
    global _glib
    if _glib is None:
        if not region_available():
            raise RuntimeError("reference region shim not built (needs /root/reference)")
        _glib = C.CDLL(_GLIB_PATH)
    out = np.zeros((nranks, 6))
    b = None if boxes is None else np.ascontiguousarray(boxes, dtype=np.int32)
    _glib.refregion_regions(_i3(ng), nranks, _p(b) if b is not None else None, _d3(origin), _d3(h), _p(out))
    return out


def orb_scatter_r(ng, origin, h, x, y, z):
    """OrthogonalRecursiveBisection::scatterR of the reference (the same particle loop as ParticleAttrib::scatter, weight
    1) on a single-rank layout -> ghosted (ng + 2)^3 field, x fastest, before any halo accumulation"""
    x, y, z = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, z))
    out = np.zeros((ng[0] + 2) * (ng[1] + 2) * (ng[2] + 2))
    olib().reforb_scatterR(_i3(ng), _d3(origin), _d3(h), C.c_long(len(x)), _p(x), _p(y), _p(z), _p(out))
    return out


# ---- the reference's PenningTrap kicks (oracle/_ref/libippl_refshim_penning.so, ref_shim/refshim_penning.cpp) ------------
_PLIB_PATH = os.path.join(_HERE, "_ref", "libippl_refshim_penning.so")
_plib = None


def penning_available(try_build=True):
    if os.path.exists(_PLIB_PATH):
        return True
    if try_build and os.path.isdir("/root/reference/src"):
        try:
            subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
        except Exception:
            return False
        return os.path.exists(_PLIB_PATH)
    return False


def penning_kick(which, R, P, E, origin, length, V0, alpha, Bext, DrInv):
    """the "Kick1" / "Kick2" lambda bodies of demos/alpine/PenningTrapManager.h over n particles; R, P, E lists of three
    arrays; returns the new P as three arrays"""
    global _plib
    if _plib is None:
        if not penning_available():
            raise RuntimeError("reference Penning shim not built (needs /root/reference)")
        _plib = C.CDLL(_PLIB_PATH)
    Ra, Pa, Ea = (np.ascontiguousarray(np.stack(a, axis=1), dtype=np.float64) for a in (R, P, E))
    _plib.refpenning_kick(int(which), C.c_long(Ra.shape[0]), _p(Ra), _p(Pa), _p(Ea), _d3(origin), _d3(length), C.c_double(V0),
                          C.c_double(alpha), C.c_double(Bext), C.c_double(DrInv))
    return [np.ascontiguousarray(Pa[:, d]) for d in range(3)]


# ---- the reference's k-space gradient step (oracle/_ref/libippl_refshim_poisson.so, ref_shim/refshim_poisson.cpp) --------
_KLIB_PATH = os.path.join(_HERE, "_ref", "libippl_refshim_poisson.so")
_klib = None


def poisson_available(try_build=True):
    if os.path.exists(_KLIB_PATH):
        return True
    if try_build and os.path.isdir("/root/reference/src"):
        try:
            subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
        except Exception:
            return False
        return os.path.exists(_KLIB_PATH)
    return False


def poisson_grad_kspace(spec, origin, h, gd):
    """the lambda "Gradient FFTPeriodicPoissonSolver" applied to every index of the complex array spec[nz][ny][nx]"""
    global _klib
    if _klib is None:
        if not poisson_available():
            raise RuntimeError("reference Poisson shim not built (needs /root/reference)")
        _klib = C.CDLL(_KLIB_PATH)
    spec = np.ascontiguousarray(spec, dtype=np.complex128)
    nz, ny, nx = spec.shape
    out = np.zeros_like(spec)
    _klib.refpoisson_grad_kspace(_i3((nx, ny, nz)), _d3(origin), _d3(h), int(gd), _p(spec), _p(out))
    return out


# ---- ParticleAttrib::scatter / ::gather lambda bodies (oracle/_ref/libippl_refshim_attrib.so, ref_shim/refshim_attrib.cpp)
_ALIB_PATH = os.path.join(_HERE, "_ref", "libippl_refshim_attrib.so")
_alib = None


def attrib_available(try_build=True):
    if os.path.exists(_ALIB_PATH):
        return True
    if try_build and os.path.isdir("/root/reference/src"):
        try:
            subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
        except Exception:
            return False
        return os.path.exists(_ALIB_PATH)
    return False


def _alib_get():
    global _alib
    if _alib is None:
        if not attrib_available():
            raise RuntimeError("reference attribute shim not built (needs /root/reference)")
        _alib = C.CDLL(_ALIB_PATH)
    return _alib


def attrib_scatter(mesh, x, y, z, q, rho, begin=0, end=None, hash=None):
    """the body of the "ParticleAttrib::scatter" lambda over particles [begin, end), optional hash remap; rho += in place"""
    R = np.ascontiguousarray(np.stack([x, y, z], axis=1), dtype=np.float64)
    q = np.ascontiguousarray(q, dtype=np.float64)
    end = len(x) if end is None else end
    hv = None if hash is None else np.ascontiguousarray(hash, dtype=np.int32)
    _alib_get().refattrib_scatter(C.c_long(begin), C.c_long(end), _p(R), _p(q), _p(hv) if hv is not None else None,
                                  C.c_long(0 if hv is None else len(hv)), _d3(mesh.origin), _d3(mesh.h), _i3(mesh.first),
                                  mesh.nghost, _i3(mesh.ext), _p(rho))
    return rho


def attrib_gather(mesh, x, y, z, efield, E, add=False):
    """the body of the "ParticleAttrib::gather" lambda for a Vector<double,3> field; E = list of three arrays (in / out)"""
    R = np.ascontiguousarray(np.stack([x, y, z], axis=1), dtype=np.float64)
    Ea = np.ascontiguousarray(np.stack(E, axis=1), dtype=np.float64)
    _alib_get().refattrib_gather(C.c_long(len(x)), _p(R), _d3(mesh.origin), _d3(mesh.h), _i3(mesh.first), mesh.nghost,
                                 _i3(mesh.ext), _p(np.ascontiguousarray(efield, dtype=np.float64)), int(add), _p(Ea))
    return [np.ascontiguousarray(Ea[:, d]) for d in range(3)]


# ---- particle ownership (oracle/_ref/libippl_refshim_locate.so, ref_shim/refshim_locate.cpp) ------------------------------
_LLIB_PATH = os.path.join(_HERE, "_ref", "libippl_refshim_locate.so")
_llib = None


def locate_available(try_build=True):
    if os.path.exists(_LLIB_PATH):
        return True
    if try_build and os.path.isdir("/root/reference/src"):
        try:
            subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
        except Exception:
            return False
        return os.path.exists(_LLIB_PATH)
    return False


def dest_rank(regions, my, x, y, z, neighbours=None):
    """the destRankOf lambda of ParticleSpatialLayout::locateParticlesPacked for rank `my`: regions[nranks][6]; neighbours =
    the cached neighbour ranks searched before the global scan (default: every other rank)"""
    global _llib
    if _llib is None:
        if not locate_available():
            raise RuntimeError("reference ownership shim not built (needs /root/reference)")
        _llib = C.CDLL(_LLIB_PATH)
    reg = np.ascontiguousarray(regions, dtype=np.float64)
    nr = reg.shape[0]
    nb = np.ascontiguousarray([r for r in range(nr) if r != my] if neighbours is None else neighbours, dtype=np.int32)
    R = np.ascontiguousarray(np.stack([x, y, z], axis=1), dtype=np.float64)
    out = np.zeros(R.shape[0], dtype=np.int32)
    _llib.reflocate_dest_rank(nr, _p(reg), int(my), _p(nb), len(nb), C.c_long(R.shape[0]), _p(R), _p(out))
    return out
