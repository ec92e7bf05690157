"""ctypes bindings of include/ippl_b200.h + small torch-backed helpers."""
import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class IpplbError(RuntimeError):
    pass


def lib_path():
    # IPPLB_LIB_VARIANT=<name> loads libippl_b200_<name>.so (A/B builds made by scripts/build_variants.sh; experiments only)
    v = os.environ.get("IPPLB_LIB_VARIANT")
    return os.path.join(_HERE, f"libippl_b200_{v}.so" if v else "libippl_b200.so")


def lib():
    """Loads the product library.  Raises (never falls back) when it has not been built."""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise IpplbError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(the CUDA extension is required; there is no CPU fallback)")
        # torch bundles its own libnccl.so.2 (2.28); load it first so the library binds to that copy
        # instead of pulling the older system NCCL in ahead of torch (same SONAME, one copy per process)
        import torch  # noqa: F401
        _LIB = C.CDLL(path)
        _LIB.ipplb_last_error.restype = C.c_char_p
        _LIB.ipplb_version.restype = C.c_char_p
        _LIB.ipplb_ctx_stream.restype = C.c_void_p
        _LIB.ipplb_launch_count.restype = C.c_long
        _LIB.ipplb_sort_ncells.restype = C.c_long
    return _LIB


def exported_symbols():
    """Every function name declared in include/ippl_b200.h."""
    hdr = open(os.path.join(_HERE, "..", "include", "ippl_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(ipplb_[a-z0-9_]+)\s*\(", hdr)))


def _check(rc):
    if rc != 0:
        raise IpplbError(f"ippl_b200 error {rc}: {lib().ipplb_last_error().decode()}")


class Mesh(C.Structure):
    _fields_ = [("ng", C.c_int * 3), ("first", C.c_int * 3), ("nl", C.c_int * 3), ("nghost", C.c_int),
                ("origin", C.c_double * 3), ("h", C.c_double * 3)]

    @staticmethod
    def make(ng, origin, h, first=(0, 0, 0), nl=None, nghost=1):
        m = Mesh()
        nl = ng if nl is None else nl
        for d in range(3):
            m.ng[d], m.first[d], m.nl[d] = int(ng[d]), int(first[d]), int(nl[d])
            m.origin[d], m.h[d] = float(origin[d]), float(h[d])
        m.nghost = nghost
        return m

    @property
    def ext(self):
        return tuple(self.nl[d] + 2 * self.nghost for d in range(3))

    @property
    def cells(self):
        e = self.ext
        return e[0] * e[1] * e[2]

    @property
    def sort_ncells(self):
        """size of the (tile-major, padded) cell-key space: cell_offsets needs sort_ncells + 1 ints"""
        return int(lib().ipplb_sort_ncells(C.byref(self)))

    @property
    def serial_mask(self):
        return sum(1 << d for d in range(3) if self.nl[d] == self.ng[d])


class _Particles(C.Structure):
    _fields_ = [("x", C.c_void_p), ("y", C.c_void_p), ("z", C.c_void_p), ("px", C.c_void_p),
                ("py", C.c_void_p), ("pz", C.c_void_p), ("q", C.c_void_p), ("q_scalar", C.c_double),
                ("n", C.c_long), ("capacity", C.c_long)]


class Push(C.Structure):
    _fields_ = [("kind", C.c_int), ("dt", C.c_double), ("do_kick2", C.c_int), ("do_kick1", C.c_int),
                ("do_drift", C.c_int), ("do_bc", C.c_int), ("origin", C.c_double * 3),
                ("length", C.c_double * 3), ("V0", C.c_double), ("alpha", C.c_double),
                ("Bext", C.c_double), ("DrInv", C.c_double)]


def leapfrog_push(dt, kick2=1, kick1=1, drift=1, bc=1):
    p = Push()
    p.kind, p.dt, p.do_kick2, p.do_kick1, p.do_drift, p.do_bc = 0, dt, kick2, kick1, drift, bc
    return p


def penning_push(dt, origin, length, Bext=5.0, kick2=1, kick1=1, drift=1, bc=1):
    p = Push()
    p.kind, p.dt, p.do_kick2, p.do_kick1, p.do_drift, p.do_bc = 1, dt, kick2, kick1, drift, bc
    for d in range(3):
        p.origin[d], p.length[d] = origin[d], length[d]
    p.V0 = 30 * length[2]
    p.alpha = -0.5 * dt
    p.Bext = Bext
    p.DrInv = 1.0 / (1 + (p.alpha * Bext) ** 2)
    return p


class Dist(C.Structure):
    """ipplb_dist: per-dimension distribution kind (0 uniform, 1 cosine, 2 normal) + parameters par[2d], par[2d+1]"""
    _fields_ = [("kind", C.c_int * 3), ("par", C.c_double * 6)]

    @staticmethod
    def make(kind, par):
        d = Dist()
        for k in range(3):
            d.kind[k] = int(kind[k])
        for k in range(6):
            d.par[k] = float(par[k])
        return d


def sample_counts(dist, rmin, rmax, regions, ntotal):
    """InverseTransformSampling::updateBounds for every rank (host only): (nlocal[nranks], ubounds[nranks][6])"""
    import numpy as np
    reg = np.ascontiguousarray(regions, dtype=np.float64)
    nr = reg.shape[0]
    nloc = (C.c_long * nr)()
    ub = np.zeros((nr, 6), dtype=np.float64)
    _check(lib().ipplb_sample_counts(C.byref(dist), (C.c_double * 3)(*rmin), (C.c_double * 3)(*rmax),
                                     reg.ctypes.data_as(C.c_void_p), nr, C.c_long(ntotal), nloc,
                                     ub.ctypes.data_as(C.c_void_p)))
    return list(nloc), ub


class Orb:
    """Host state machine of OrthogonalRecursiveBisection::binaryRepartition (works without a GPU)."""

    def __init__(self, ng, nranks):
        self._h = C.c_void_p()
        self.nranks = nranks
        _check(lib().ipplb_orb_begin(C.byref(self._h), (C.c_int * 3)(*ng), nranks))

    def next(self):
        """(lo[3], hi[3], axis) of the pending cut, or None"""
        lo, hi, ax, pend = (C.c_int * 3)(), (C.c_int * 3)(), C.c_int(), C.c_int()
        _check(lib().ipplb_orb_next(self._h, lo, hi, C.byref(ax), C.byref(pend)))
        return (list(lo), list(hi), ax.value) if pend.value else None

    def cut(self, reduced):
        import numpy as np
        w = np.ascontiguousarray(reduced, dtype=np.float64)
        _check(lib().ipplb_orb_cut(self._h, w.ctypes.data_as(C.c_void_p), len(w)))

    def finish(self):
        import numpy as np
        boxes = np.zeros((self.nranks, 6), dtype=np.int32)
        ok = C.c_int()
        _check(lib().ipplb_orb_finish(self._h, boxes.ctypes.data_as(C.c_void_p), C.byref(ok)))
        self._h = C.c_void_p()
        return boxes, bool(ok.value)


class Particles:
    """SoA fp64 particle bundle in device memory (torch tensors own the storage)."""

    NAMES = ("x", "y", "z", "px", "py", "pz")

    def __init__(self, capacity, device, q=None, with_q_array=False):
        import torch
        self.capacity = int(capacity)
        self.device = device
        self.arr = {k: torch.empty(self.capacity, dtype=torch.float64, device=device) for k in self.NAMES}
        self.qarr = torch.empty(self.capacity, dtype=torch.float64, device=device) if with_q_array else None
        self.q_scalar = 0.0 if q is None else float(q)
        self.n = 0

    @staticmethod
    def from_host(R, P, device, q=0.0, capacity=None):
        import numpy as np
        import torch
        n = len(R[0])
        qa = isinstance(q, np.ndarray)
        p = Particles(capacity or n, device, None if qa else q, with_q_array=qa)
        for k, a in zip(Particles.NAMES, list(R) + list(P)):
            p.arr[k][:n].copy_(torch.from_numpy(np.ascontiguousarray(a)))
        if qa:
            p.qarr[:n].copy_(torch.from_numpy(np.ascontiguousarray(q)))
        p.n = n
        return p

    def struct(self):
        s = _Particles()
        for k in self.NAMES:
            setattr(s, k, self.arr[k].data_ptr())
        s.q = self.qarr.data_ptr() if self.qarr is not None else None
        s.q_scalar = self.q_scalar
        s.n, s.capacity = self.n, self.capacity
        return s

    def adopt(self, s):
        """Take over the pointers/count of a struct the library has updated (sort swaps buffers)."""
        self.n = int(s.n)

    def host(self, names=None):
        names = names or self.NAMES
        return [self.arr[k][: self.n].cpu().numpy() for k in names]


def lib_particles_array(structs):
    """ctypes array of ipplb_particles from a list of structs (for the *_batches entry points)"""
    arr = (_Particles * len(structs))()
    for i, st in enumerate(structs):
        arr[i] = st
    return arr


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class Context:
    """One ipplb_ctx on one GPU, enqueuing on torch's current stream of that device."""

    def __init__(self, device=0, use_torch_stream=True):
        import torch
        if not torch.cuda.is_available():
            raise IpplbError("no CUDA device: ippl_b200 has no CPU fallback")
        self.torch = torch
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        self._h = C.c_void_p()
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream) if use_torch_stream else None
        _check(lib().ipplb_ctx_create(C.byref(self._h), device, stream, 0 if use_torch_stream else 1))
        self.rank, self.nranks = 0, 1

    def close(self):
        if self._h:
            lib().ipplb_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers -----------------------------------------------------------------------------
    def zeros(self, n):
        return self.torch.zeros(int(n), dtype=self.torch.float64, device=self.device)

    def field(self, mesh, ncomp=1):
        return self.zeros(mesh.cells * ncomp)

    def sync(self):
        _check(lib().ipplb_sync(self._h))

    def set_gather_variant(self, variant):
        """1: 8-byte field loads (default); 2: 16-byte loads per x-pair of stencil nodes (ipplb_ctx_set_gather_variant)"""
        _check(lib().ipplb_ctx_set_gather_variant(self._h, int(variant)))

    @property
    def launches(self):
        return lib().ipplb_launch_count(self._h)

    # -- kernels -----------------------------------------------------------------------------
    def scatter(self, mesh, x, y, z, q, rho, begin=0, end=None, hash=None):
        end = len(x) if end is None else end
        qa = q if hasattr(q, "data_ptr") else None
        qs = 0.0 if qa is not None else float(q)
        _check(lib().ipplb_scatter_cic(self._h, C.byref(mesh), C.c_long(begin), C.c_long(end), _ptr(x),
                                       _ptr(y), _ptr(z), _ptr(qa), C.c_double(qs), _ptr(hash), _ptr(rho)))

    def scatter_sorted(self, mesh, n, x, y, z, q, offsets, rho):
        qa = q if hasattr(q, "data_ptr") else None
        qs = 0.0 if qa is not None else float(q)
        _check(lib().ipplb_scatter_cic_sorted(self._h, C.byref(mesh), C.c_long(n), _ptr(x), _ptr(y), _ptr(z),
                                              _ptr(qa), C.c_double(qs), _ptr(offsets), _ptr(rho)))

    def gather(self, mesh, x, y, z, field, out, add=False):
        ncomp = len(out)
        arr = (C.c_void_p * 3)(*([o.data_ptr() for o in out] + [None] * (3 - ncomp)))
        _check(lib().ipplb_gather_cic(self._h, C.byref(mesh), C.c_long(len(x)), _ptr(x), _ptr(y), _ptr(z),
                                      _ptr(field), ncomp, arr, int(add)))

    def gather_push(self, mesh, push, parts, efield):
        s = parts.struct()
        _check(lib().ipplb_gather_push(self._h, C.byref(mesh), C.byref(push), C.byref(s), _ptr(efield)))

    def axpy(self, a, x, y, n=None):
        n = len(x) if n is None else n
        _check(lib().ipplb_axpy(self._h, C.c_long(n), C.c_double(a), _ptr(x), _ptr(y)))

    def apply_periodic_bc(self, x, y, z, lo, hi, mask=7, n=None):
        n = len(x) if n is None else n
        _check(lib().ipplb_apply_periodic_bc(self._h, C.c_long(n), _ptr(x), _ptr(y), _ptr(z),
                                             (C.c_double * 3)(*lo), (C.c_double * 3)(*hi), mask))

    def penning_kick(self, which, push, R, P, E, n=None):
        n = len(R[0]) if n is None else n
        _check(lib().ipplb_penning_kick(self._h, which, C.byref(push), C.c_long(n), _ptr(R[0]), _ptr(R[1]),
                                        _ptr(R[2]), _ptr(P[0]), _ptr(P[1]), _ptr(P[2]), _ptr(E[0]),
                                        _ptr(E[1]), _ptr(E[2])))

    def sort_by_cell(self, mesh, src, dst, offsets):
        s, d = src.struct(), dst.struct()
        _check(lib().ipplb_sort_by_cell(self._h, C.byref(mesh), C.byref(s), C.byref(d), _ptr(offsets)))
        dst.n, dst.q_scalar = src.n, src.q_scalar

    def offsets_buffer(self, mesh):
        return self.torch.zeros(mesh.sort_ncells + 1, dtype=self.torch.int32, device=self.device)

    def field_fill(self, f, value=0.0):
        _check(lib().ipplb_field_fill(self._h, _ptr(f), C.c_long(f.numel()), C.c_double(value)))

    def field_sum(self, mesh, f):
        out = C.c_double()
        _check(lib().ipplb_field_sum(self._h, C.byref(mesh), _ptr(f), C.byref(out)))
        return out.value

    def field_ex_stats(self, mesh, ef):
        out = (C.c_double * 2)()
        _check(lib().ipplb_field_ex_stats(self._h, C.byref(mesh), _ptr(ef), out))
        return out[0], out[1]

    def field_density(self, mesh, f, cell_volume, shift):
        _check(lib().ipplb_field_density(self._h, C.byref(mesh), _ptr(f), C.c_double(cell_volume),
                                         C.c_double(shift)))

    def halo_accumulate_periodic(self, mesh, f, ncomp=1, mask=None):
        mask = mesh.serial_mask if mask is None else mask
        _check(lib().ipplb_halo_accumulate_periodic(self._h, C.byref(mesh), _ptr(f), ncomp, mask))

    def halo_fill_periodic(self, mesh, f, ncomp=1, mask=None):
        mask = mesh.serial_mask if mask is None else mask
        _check(lib().ipplb_halo_fill_periodic(self._h, C.byref(mesh), _ptr(f), ncomp, mask))

    def pic_step(self, mesh, push, parts, scratch, offsets, efield, rho, do_sort=True, bins=None):
        """One metric step.  do_sort 1: counting sort + sorted scatter, 2: fused single-pass step on bucketed
        particles (`bins`); in both `parts` and `scratch` swap storage."""
        s = parts.struct()
        sc = scratch.struct() if scratch is not None else None
        _check(lib().ipplb_pic_step(self._h, C.byref(mesh), C.byref(push), C.byref(s),
                                    C.byref(sc) if sc is not None else None, _ptr(offsets),
                                    bins._h if bins is not None else None, _ptr(efield), _ptr(rho), int(do_sort)))
        if do_sort:
            parts.arr, scratch.arr = scratch.arr, parts.arr
            parts.qarr, scratch.qarr = scratch.qarr, parts.qarr
        parts.n = int(s.n)

    def pic_step_host_fields(self, mesh, push, parts, scratch, bins, efield_host, rho_host, efield_dev, rho_dev):
        """ipplb_pic_step_host_fields: E from (pinned) host memory, fused step on the resident bucketed particles, rho back
        to (pinned) host memory; synchronises.  `parts` and `scratch` swap storage."""
        s, sc = parts.struct(), scratch.struct()
        _check(lib().ipplb_pic_step_host_fields(self._h, C.byref(mesh), C.byref(push), C.byref(s), C.byref(sc), bins._h,
                                                C.c_void_p(efield_host.data_ptr()), C.c_void_p(rho_host.data_ptr()),
                                                _ptr(efield_dev), _ptr(rho_dev)))
        parts.arr, scratch.arr = scratch.arr, parts.arr
        parts.qarr, scratch.qarr = scratch.qarr, parts.qarr
        parts.n = int(s.n)

    # -- diagnostics, sampling, ORB -----------------------------------------------
    def field_energy_stats(self, mesh, ef):
        """(sum E_d^2 [3], max|E_d| [3], sum dot(E,E)) over the interior"""
        out = (C.c_double * 7)()
        _check(lib().ipplb_field_energy_stats(self._h, C.byref(mesh), _ptr(ef), out))
        return list(out[0:3]), list(out[3:6]), out[6]

    def field_norm_stats(self, mesh, f):
        out = (C.c_double * 2)()
        _check(lib().ipplb_field_norm_stats(self._h, C.byref(mesh), _ptr(f), out))
        return out[0], out[1]

    def particles_kinetic(self, parts, n=None):
        out = C.c_double()
        n = parts.n if n is None else n
        _check(lib().ipplb_particles_kinetic(self._h, C.c_long(n), _ptr(parts.arr["px"]), _ptr(parts.arr["py"]),
                                             _ptr(parts.arr["pz"]), C.byref(out)))
        return out.value

    def sample_positions(self, dist, umin, umax, seed, first_id, n, parts):
        _check(lib().ipplb_sample_positions(self._h, C.byref(dist), (C.c_double * 3)(*umin), (C.c_double * 3)(*umax),
                                            C.c_uint64(seed), C.c_long(first_id), C.c_long(n), _ptr(parts.arr["x"]),
                                            _ptr(parts.arr["y"]), _ptr(parts.arr["z"])))

    def sample_normal(self, mu, sd, seed, first_id, n, parts):
        _check(lib().ipplb_sample_normal(self._h, (C.c_double * 3)(*mu), (C.c_double * 3)(*sd), C.c_uint64(seed),
                                         C.c_long(first_id), C.c_long(n), _ptr(parts.arr["px"]), _ptr(parts.arr["py"]),
                                         _ptr(parts.arr["pz"])))

    def field_fill_pdf(self, mesh, dist, f):
        _check(lib().ipplb_field_fill_pdf(self._h, C.byref(mesh), C.byref(dist), _ptr(f)))

    def orb_plane_sums(self, mesh, f, axis, lo, hi):
        import numpy as np
        out = np.zeros(hi[axis] - lo[axis] + 1, dtype=np.float64)
        _check(lib().ipplb_orb_plane_sums(self._h, C.byref(mesh), _ptr(f), axis, (C.c_int * 3)(*lo), (C.c_int * 3)(*hi),
                                          out.ctypes.data_as(C.c_void_p)))
        return out

    def orb_repartition(self, mesh, nranks, weight):
        import numpy as np
        boxes = np.zeros((nranks, 6), dtype=np.int32)
        ok = C.c_int()
        _check(lib().ipplb_orb_repartition(self._h, C.byref(mesh), nranks, _ptr(weight), boxes.ctypes.data_as(C.c_void_p),
                                           C.byref(ok)))
        return boxes, bool(ok.value)

    # -- multi-GPU ---------------------------------------------------------------------------
    def comm_init(self, rank, nranks, id_bytes=None):
        self.rank, self.nranks = rank, nranks
        buf = C.create_string_buffer(bytes(id_bytes), 128) if id_bytes is not None else None
        _check(lib().ipplb_comm_init(self._h, rank, nranks, buf))

    def set_layout(self, layout, origin, h):
        _check(lib().ipplb_ctx_set_layout(self._h, layout._h, (C.c_double * 3)(*origin), (C.c_double * 3)(*h)))

    def halo_exchange(self, f, ncomp, mode):
        _check(lib().ipplb_halo_exchange(self._h, _ptr(f), ncomp, 0 if mode == "fill" else 1))

    def update(self, parts):
        s = parts.struct()
        sent = (C.c_long * self.nranks)()
        recv = (C.c_long * self.nranks)()
        _check(lib().ipplb_update(self._h, C.byref(s), sent, recv))
        parts.n = int(s.n)
        return list(sent), list(recv)

    def update_plan(self, parts):
        """ipplb_update_plan: (n_after, sent, recv) -- collective; IpplbError (on every rank) when some rank does not fit"""
        s = parts.struct()
        n_after = C.c_long()
        sent, recv = (C.c_long * self.nranks)(), (C.c_long * self.nranks)()
        rc = lib().ipplb_update_plan(self._h, C.byref(s), C.byref(n_after), sent, recv)
        return rc, n_after.value, list(sent), list(recv)

    def update_commit(self, parts):
        s = parts.struct()
        _check(lib().ipplb_update_commit(self._h, C.byref(s)))
        parts.n = int(s.n)

    def migrate_connect(self, seg_cap):
        _check(lib().ipplb_migrate_connect(self._h, C.c_long(int(seg_cap))))

    def migrate_counts(self):
        sent, recv = (C.c_long * self.nranks)(), (C.c_long * self.nranks)()
        _check(lib().ipplb_migrate_counts(self._h, sent, recv))
        return list(sent), list(recv)

    def allreduce_sum(self, v):
        if isinstance(v, int):
            c = C.c_long(v)
            _check(lib().ipplb_allreduce_sum_i64(self._h, C.byref(c)))
        else:
            c = C.c_double(v)
            _check(lib().ipplb_allreduce_sum_f64(self._h, C.byref(c)))
        return c.value


def nccl_unique_id():
    buf = C.create_string_buffer(128)
    _check(lib().ipplb_nccl_unique_id(buf))
    return buf.raw


class Bins:
    """ipplb_bins: per-tile bucketed particle store behind the fused step.  `cur` holds the particles,
    `nxt` is the spare bundle; both must have `capacity` elements per array."""

    def __init__(self, ctx, mesh, capacity):
        self.ctx, self.mesh, self.capacity = ctx, mesh, int(capacity)
        self._h = C.c_void_p()
        _check(lib().ipplb_bins_create(ctx._h, C.byref(mesh), C.c_long(self.capacity), C.byref(self._h)))
        self.ntiles = lib().ipplb_bins_ntiles(self._h)

    def close(self):
        if self._h:
            lib().ipplb_bins_destroy(self._h)
            self._h = C.c_void_p()

    def build(self, src, dst):
        """counting sort of the contiguous `src` into the buckets of `dst`"""
        s, d = src.struct(), dst.struct()
        _check(lib().ipplb_bins_build(self.ctx._h, self._h, C.byref(s), C.byref(d)))
        dst.n, dst.q_scalar = src.n, src.q_scalar

    def step(self, push, cur, nxt, efield, rho, exit_buf=None, region=None):
        """ipplb_bins_step; the caller's `cur` / `nxt` swap storage (like the sort).  Asynchronous."""
        s, d = cur.struct(), nxt.struct()
        cap = 0 if exit_buf is None else exit_buf.numel() // 6
        rmin = (C.c_double * 3)(*region[:3]) if region is not None else None
        rmax = (C.c_double * 3)(*region[3:]) if region is not None else None
        _check(lib().ipplb_bins_step(self.ctx._h, self._h, C.byref(push), C.byref(s), C.byref(d), _ptr(efield),
                                     _ptr(rho), _ptr(exit_buf), cap, rmin, rmax))
        cur.arr, nxt.arr = nxt.arr, cur.arr
        cur.qarr, nxt.qarr = nxt.qarr, cur.qarr

    def set_build_variant(self, variant):
        """1: per-cell positions (default); 2: arrival order, warp-aggregated tile cursors (ipplb_bins_set_build_variant)"""
        _check(lib().ipplb_bins_set_build_variant(self._h, int(variant)))

    def set_timing(self, on=True):
        _check(lib().ipplb_bins_set_timing(self._h, 1 if on else 0))

    def kernel_ms(self):
        """per-launch durations (ms) of the fused kernel since set_timing(True) (synchronises)"""
        out, n = (C.c_double * 256)(), C.c_int()
        _check(lib().ipplb_bins_kernel_ms(self.ctx._h, self._h, out, 256, C.byref(n)))
        return [out[i] for i in range(n.value)]

    def status(self):
        """(n_local, n_tail, n_exit, flags) after the last build / step / append (synchronises)"""
        n, t, e, f = C.c_long(), C.c_long(), C.c_long(), C.c_int()
        _check(lib().ipplb_bins_status(self.ctx._h, self._h, C.byref(n), C.byref(t), C.byref(e), C.byref(f)))
        return n.value, t.value, e.value, f.value

    def append(self, cur, src, count):
        """append `count` particles (six device tensors) to the tail of `cur`"""
        s = cur.struct()
        arr = (C.c_void_p * 6)(*[a.data_ptr() for a in src])
        _check(lib().ipplb_bins_append(self.ctx._h, self._h, C.byref(s), arr, C.c_long(count)))
        cur.n += count

    def migrate(self, cur, exit_buf, rho=None):
        """ipplb_bins_migrate: exchange the leavers of the last step, append + deposit the arrivals"""
        s = cur.struct()
        nr = self.ctx.nranks
        sent, recv = (C.c_long * nr)(), (C.c_long * nr)()
        cap = 0 if exit_buf is None else exit_buf.numel() // 6
        _check(lib().ipplb_bins_migrate(self.ctx._h, self._h, C.byref(s), _ptr(exit_buf), cap, _ptr(rho), sent, recv))
        cur.n = int(s.n)
        return list(sent), list(recv)

    def migrate_async(self, cur, rho=None):
        """ipplb_bins_migrate_async: peer-memory migration of the leavers of the last step; no host synchronisation"""
        s = cur.struct()
        _check(lib().ipplb_bins_migrate_async(self.ctx._h, self._h, C.byref(s), _ptr(rho)))

    def compact(self, cur, out):
        s, d = cur.struct(), out.struct()
        _check(lib().ipplb_bins_compact(self.ctx._h, self._h, C.byref(s), C.byref(d)))
        out.n, out.q_scalar = int(d.n), cur.q_scalar
        return out.n

    def kinetic(self, cur):
        """sum_i dot(P_i, P_i) over the bucketed store (no compaction)"""
        out = C.c_double()
        s = cur.struct()
        _check(lib().ipplb_bins_kinetic(self.ctx._h, self._h, C.byref(s), C.byref(out)))
        return out.value

    def tables(self):
        import numpy as np
        st, cp, ct = (np.zeros(self.ntiles, dtype=np.int32) for _ in range(3))
        _check(lib().ipplb_bins_tables(self.ctx._h, self._h, st.ctypes.data_as(C.c_void_p),
                                       cp.ctypes.data_as(C.c_void_p), ct.ctypes.data_as(C.c_void_p)))
        return st, cp, ct


class Poisson:
    """cuFFT periodic Poisson solve (non-owned stage).  Single rank: Poisson(ctx, mesh) with the whole domain;
    multi rank: Poisson(ctx, None, layout=layout, origin=..., h=...) = replicated solve over NCCL."""

    def __init__(self, ctx, mesh, layout=None, origin=None, h=None, slab=False):
        self.ctx = ctx
        self._h = C.c_void_p()
        if layout is not None:
            # slab=True: the slab-decomposed solve (ipplb_poisson_create_slab) instead of the replicated one
            create = lib().ipplb_poisson_create_slab if slab else lib().ipplb_poisson_create_dist
            _check(create(ctx._h, layout._h, (C.c_double * 3)(*origin), (C.c_double * 3)(*h), C.byref(self._h)))
        else:
            _check(lib().ipplb_poisson_create(ctx._h, C.byref(mesh), C.byref(self._h)))

    def solve(self, rho, efield):
        _check(lib().ipplb_poisson_solve(self._h, _ptr(rho), _ptr(efield)))

    def close(self):
        if self._h:
            lib().ipplb_poisson_destroy(self._h)
            self._h = C.c_void_p()


class Layout:
    """Host-only FieldLayout mirror (works without a GPU)."""

    def __init__(self, ng, nranks, parallel=(1, 1, 1), periodic=True, nghost=1):
        self._h = C.c_void_p()
        self.ng, self.nranks, self.nghost, self.periodic = tuple(ng), nranks, nghost, periodic
        _check(lib().ipplb_layout_create(C.byref(self._h), (C.c_int * 3)(*ng), (C.c_int * 3)(*parallel), nranks,
                                         int(periodic), nghost))

    def close(self):
        if self._h:
            lib().ipplb_layout_destroy(self._h)
            self._h = C.c_void_p()

    def boxes(self):
        import numpy as np
        out = np.zeros((self.nranks, 6), dtype=np.int32)
        _check(lib().ipplb_layout_boxes(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def set_boxes(self, boxes):
        import numpy as np
        b = np.ascontiguousarray(boxes, dtype=np.int32)
        _check(lib().ipplb_layout_set_boxes(self._h, b.ctypes.data_as(C.c_void_p)))

    def neighbors(self, rank):
        import numpy as np
        out = np.zeros((512, 14), dtype=np.int32)
        n = lib().ipplb_layout_neighbors(self._h, rank, out.ctypes.data_as(C.c_void_p), 512)
        if n < 0 or n > 512:
            raise IpplbError("layout_neighbors failed")
        return out[:n].copy()

    def regions(self, origin, h):
        import numpy as np
        out = np.zeros((self.nranks, 6), dtype=np.float64)
        _check(lib().ipplb_layout_regions(self._h, (C.c_double * 3)(*origin), (C.c_double * 3)(*h),
                                          out.ctypes.data_as(C.c_void_p)))
        return out

    def mesh(self, rank, origin, h):
        m = Mesh()
        _check(lib().ipplb_layout_mesh(self._h, rank, (C.c_double * 3)(*origin), (C.c_double * 3)(*h), C.byref(m)))
        return m


class SlabPlan:
    """Host-only plan of the slab-decomposed FFT solve (ipplb_slabplan_*; works without a GPU): the sub-box copies and
    messages of the four phases for one rank of a layout."""
    BUFS = ("rho", "ef", "real", "spec2d", "specz", "send", "recv")

    def __init__(self, layout, rank):
        self._h = C.c_void_p()
        _check(lib().ipplb_slabplan_create(layout._h, rank, C.byref(self._h)))
        info = (C.c_long * 16)()
        _check(lib().ipplb_slabplan_info(self._h, info))
        v = list(info)
        self.nranks, self.rank, self.ng, self.nxh = v[0], v[1], tuple(v[2:5]), v[5]
        self.zs, self.ze, self.ys, self.ye = v[6:10]
        self.size = dict(zip(self.BUFS[2:], v[10:15]))
        self.nghost = v[15]

    def rows(self, phase, which):
        """which: 0 copies before the exchange, 1 messages, 2 copies after it.  Copies come back as dicts."""
        import numpy as np
        n = C.c_int(0)
        _check(lib().ipplb_slabplan_rows(self._h, phase, which, None, 0, C.byref(n)))
        out = np.zeros((max(n.value, 1), 16), dtype=np.int64)
        _check(lib().ipplb_slabplan_rows(self._h, phase, which, out.ctypes.data_as(C.c_void_p), n.value, C.byref(n)))
        out = out[:n.value]
        if which == 1:
            return [dict(peer=int(r[0]), soff=int(r[1]), scount=int(r[2]), roff=int(r[3]), rcount=int(r[4])) for r in out]
        return [dict(src=self.BUFS[r[0]], dst=self.BUFS[r[1]], src_off=int(r[2]), dst_off=int(r[3]), ss=tuple(int(x) for x in r[4:7]),
                     ds=tuple(int(x) for x in r[7:10]), n=tuple(int(x) for x in r[10:13]), elem=int(r[13])) for r in out]

    def close(self):
        if self._h:
            lib().ipplb_slabplan_destroy(self._h)
            self._h = C.c_void_p()


class Loop:
    """ipplb_loop: all ranks of a small job in ONE process on ONE device (no NCCL); the same kernels and tables as the
    NCCL path with device-to-device copies as the transport.  ctxs[r] becomes rank r."""

    def __init__(self, ctxs):
        self.ctxs = list(ctxs)
        self.n = len(self.ctxs)
        self._h = C.c_void_p()
        arr = (C.c_void_p * self.n)(*[c._h.value for c in self.ctxs])
        _check(lib().ipplb_loop_create(C.byref(self._h), arr, self.n))
        for r, c in enumerate(self.ctxs):
            c.rank, c.nranks = r, self.n

    def close(self):
        if self._h:
            lib().ipplb_loop_destroy(self._h)
            self._h = C.c_void_p()

    def halo_exchange(self, fields, ncomp, mode):
        arr = (C.c_void_p * self.n)(*[f.data_ptr() for f in fields])
        _check(lib().ipplb_loop_halo_exchange(self._h, arr, ncomp, 0 if mode == "fill" else 1))

    def update(self, parts):
        """parts[r]: Particles of rank r (contiguous).  Returns (sent[r][t], recv[r][t])."""
        structs = lib_particles_array([p.struct() for p in parts])
        sent, recv = (C.c_long * (self.n * self.n))(), (C.c_long * (self.n * self.n))()
        _check(lib().ipplb_loop_update(self._h, structs, sent, recv))
        for r, p in enumerate(parts):
            p.n = int(structs[r].n)
        n = self.n
        return ([[sent[r * n + t] for t in range(n)] for r in range(n)], [[recv[r * n + t] for t in range(n)] for r in range(n)])

    def migrate_connect(self, seg_cap):
        _check(lib().ipplb_loop_migrate_connect(self._h, C.c_long(int(seg_cap))))

    def poisson_solve(self, solvers, rho, efield):
        """slab-decomposed solve of every in-process rank (solvers[r] = Poisson(ctxs[r], None, layout=..., slab=True))"""
        sarr = (C.c_void_p * self.n)(*[q._h.value for q in solvers])
        rarr = (C.c_void_p * self.n)(*[r.data_ptr() for r in rho])
        earr = (C.c_void_p * self.n)(*[e.data_ptr() for e in efield])
        _check(lib().ipplb_loop_poisson_solve(self._h, sarr, rarr, earr))

    def bins_migrate(self, bins, cur, rho=None):
        structs = lib_particles_array([p.struct() for p in cur])
        barr = (C.c_void_p * self.n)(*[b._h.value for b in bins])
        rarr = (C.c_void_p * self.n)(*[r.data_ptr() for r in rho]) if rho is not None else None
        _check(lib().ipplb_loop_bins_migrate(self._h, barr, structs, rarr))
