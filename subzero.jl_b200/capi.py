"""ctypes binding of include/subzero_b200.h.

`Library(path, prefix)` binds one shared object.  The product library is
`csrc/libsubzero_b200.so` (prefix `sz_`, CUDA sm_100a); the CPU oracle under `oracle/`
exports the same ABI with prefix `szo_` and is bound with the same class — but ONLY by
tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs (see oracle/szo.py).

There is no CPU fallback: `product()` raises if the CUDA library is missing or fails to
load, and `Handle` calls raise `SubzeroError` on any non-zero status.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# SZ_B200_LIB: another build of the SAME CUDA library (A/B kernel experiments); never a fallback
PRODUCT_LIB = os.environ.get("SZ_B200_LIB") or os.path.join(HERE, "csrc", "libsubzero_b200.so")

c_double_p = C.POINTER(C.c_double)
c_i64_p = C.POINTER(C.c_int64)
c_i32_p = C.POINTER(C.c_int32)
c_u32_p = C.POINTER(C.c_uint32)
c_u8_p = C.POINTER(C.c_uint8)

STATUS_ACTIVE, STATUS_REMOVE, STATUS_FUSE = 1, 2, 3
# outputs of calc_eulerian_data! (output.jl:859-905) in SZ_GRID_* order
GRID_OUTPUTS = ("u_grid", "v_grid", "dudt_grid", "dvdt_grid", "si_frac_grid", "overarea_grid", "mass_grid", "area_grid",
                "height_grid", "stress_xx_grid", "stress_yx_grid", "stress_xy_grid", "stress_yy_grid", "stress_eig_grid",
                "strain_ux_grid", "strain_vx_grid", "strain_uy_grid", "strain_vy_grid")
BOUNDARY_OPEN, BOUNDARY_PERIODIC, BOUNDARY_COLLISION, BOUNDARY_MOVING = 0, 1, 2, 3
WARN_HEIGHT_CAPPED, WARN_FORCE_SCALED, WARN_VELOCITY_LIMITED, WARN_XI_CLAMPED = 1, 2, 4, 8


class SubzeroError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("subzero_b200 status %d: %s" % (code, msg))
        self.code = code


class Config(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "rho_o", "rho_a", "Cd_io", "Cd_ia", "Cd_ao", "f", "turn_theta", "L", "k", "nu", "mu", "E",
        "floe_floe_max_overlap", "floe_domain_max_overlap",
        "rho_i", "max_floe_height", "maximum_xi", "stress_lambda")] + [
        (n, C.c_int32) for n in (
            "coupling_dd", "two_way_coupling_on", "dt", "device", "max_regions_per_pair",
            "max_pairs_per_floe", "threads", "reserved0")] + [("floe_capacity", C.c_int64)]


DOUBLE_FIELDS = (
    "centroid_x", "centroid_y", "height", "area", "mass", "rmax", "moment", "alpha", "u", "v",
    "xi", "fxOA", "fyOA", "trqOA", "hflx_factor", "overarea", "collision_force", "collision_trq",
    "stress_accum", "stress_instant", "strain", "p_dxdt", "p_dydt", "p_dudt", "p_dvdt", "p_dxidt",
    "p_dalphadt")
FIELD_WIDTH = {"collision_force": 2, "stress_accum": 4, "stress_instant": 4, "strain": 4}


class FloeSoA(C.Structure):
    _fields_ = ([("n", C.c_int64), ("n_init", C.c_int64)] +
                [(n, c_double_p) for n in DOUBLE_FIELDS] +
                [("status_tag", c_i32_p), ("id", c_i64_p), ("ghost_id", c_i64_p),
                 ("ghost_offsets", c_i64_p), ("ghost_index", c_i64_p),
                 ("vert_offsets", c_i64_p), ("vert_xy", c_double_p),
                 ("mc_offsets", c_i64_p), ("mc_x", c_double_p), ("mc_y", c_double_p)])


class PointsGenerator(C.Structure):
    _fields_ = [("kind", C.c_int32), ("npoints", C.c_int32), ("err", C.c_double), ("delta_g", C.c_double), ("seed", C.c_uint64)]


POINTS_MONTE_CARLO, POINTS_SUB_GRID = 0, 1


class Counts(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "n_init", "n_total", "n_vertices", "n_mc", "n_ghost_links", "n_candidates", "n_pairs",
        "n_overlap", "n_fuse", "n_rows", "n_domain_pairs", "n_clip_fail")]

    def asdict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def _dp(a):
    return a.ctypes.data_as(c_double_p) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(c_i64_p) if a is not None else None


class Library:
    """One loaded shared object exporting the subzero_b200 C ABI under `prefix`."""

    SIGS = {
        "default_config": (None, [C.POINTER(Config)]),
        "create": (C.c_int32, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
        "destroy": (None, [C.c_void_p]),
        "last_error": (C.c_char_p, [C.c_void_p]),
        "version": (C.c_char_p, []),
        "set_grid": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32] + [C.c_double] * 4),
        "set_fields": (C.c_int32, [C.c_void_p] + [c_double_p] * 5),
        "set_temperatures": (C.c_int32, [C.c_void_p, c_double_p, c_double_p]),
        "get_ocean_fields": (C.c_int32, [C.c_void_p] + [c_double_p] * 4),
        "get_cell_floes": (C.c_int32, [C.c_void_p, c_i64_p, c_i64_p, c_i64_p, c_double_p]),
        "set_domain": (C.c_int32, [C.c_void_p, c_i32_p, c_double_p, c_double_p, c_double_p,
                                   C.c_int32, c_i64_p, c_double_p, c_double_p, c_double_p]),
        "get_domain": (C.c_int32, [C.c_void_p, c_double_p, c_double_p]),
        "upload_floes": (C.c_int32, [C.c_void_p, C.POINTER(FloeSoA)]),
        "upload_state": (C.c_int32, [C.c_void_p, C.POINTER(FloeSoA)]),
        "get_counts": (C.c_int32, [C.c_void_p, C.POINTER(Counts)]),
        "download_floes": (C.c_int32, [C.c_void_p, C.POINTER(FloeSoA)]),
        "add_ghosts": (C.c_int32, [C.c_void_p, c_i64_p]),
        "step_collisions": (C.c_int32, [C.c_void_p]),
        "remove_ghosts": (C.c_int32, [C.c_void_p]),
        "step_coupling": (C.c_int32, [C.c_void_p]),
        "step_floe_properties": (C.c_int32, [C.c_void_p, C.c_int64]),
        "step": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int32]),
        "step_host": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int32, C.POINTER(FloeSoA), C.POINTER(FloeSoA)]),
        "step_host_partial": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int32, C.POINTER(FloeSoA), C.POINTER(FloeSoA)]),
        "upload_state_begin": (C.c_int32, [C.c_void_p, C.c_int32, C.POINTER(FloeSoA)]),
        "coupling_begin": (C.c_int32, [C.c_void_p]),
        "get_interactions": (C.c_int32, [C.c_void_p, c_i64_p, c_double_p]),
        "set_interactions": (C.c_int32, [C.c_void_p, c_i64_p, c_double_p]),
        "get_pairs": (C.c_int32, [C.c_void_p, C.c_int32, c_i64_p]),
        "get_warnings": (C.c_int32, [C.c_void_p, c_u32_p]),
        "get_timings": (C.c_int32, [C.c_void_p, c_double_p]),
        "halo_configure": (C.c_int32, [C.c_void_p, C.c_int32, c_i64_p, c_i64_p]),
        "halo_bytes": (C.c_int32, [C.c_void_p, C.c_int32, c_i64_p]),
        "halo_pack": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64]),
        "halo_unpack": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64]),
        "halo_pack_on": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]),
        "halo_unpack_on": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]),
        "pair_overlap_areas": (C.c_int32, [C.c_void_p, C.c_int64, c_i64_p, c_double_p, c_u8_p]),
        "eulerian_data": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, c_double_p, c_double_p, C.c_int32, c_i32_p, c_double_p]),
        "clip_polygons": (C.c_int32, [C.c_void_p, c_double_p, C.c_int32, c_double_p, C.c_int32,
                                      C.c_int32, C.c_int32, c_i32_p, c_double_p, c_double_p]),
        "generate_subfloe_points": (C.c_int32, [C.c_void_p, C.POINTER(PointsGenerator), C.c_int64, c_i64_p, c_i64_p, c_double_p,
                                                c_double_p, C.c_int64, c_i32_p, C.c_int32]),
        # slab decomposition inside the library
        "slab_create": (C.c_int32, [C.POINTER(Config), C.c_int32, C.c_int32, C.c_int32, c_i32_p, C.c_double, C.c_void_p,
                                    C.c_void_p, C.POINTER(C.c_void_p)]),
        "slab_destroy": (None, [C.c_void_p]),
        "slab_last_error": (C.c_char_p, [C.c_void_p]),
        "slab_handle": (C.c_int32, [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]),
        "slab_set_grid": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32] + [C.c_double] * 4),
        "slab_set_fields": (C.c_int32, [C.c_void_p] + [c_double_p] * 5),
        "slab_set_domain": (C.c_int32, [C.c_void_p, c_i32_p, c_double_p, c_double_p, c_double_p,
                                        C.c_int32, c_i64_p, c_double_p, c_double_p, c_double_p]),
        "slab_set_edges": (C.c_int32, [C.c_void_p, c_double_p]),
        "slab_build": (C.c_int32, [C.c_void_p, C.POINTER(C.POINTER(FloeSoA)), C.POINTER(c_i64_p)]),
        "slab_local_count": (C.c_int32, [C.c_void_p, C.c_int32, c_i64_p, c_i64_p]),
        "slab_local_index": (C.c_int32, [C.c_void_p, C.c_int32, c_i64_p, c_i32_p]),
        "slab_step": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int32]),
        "slab_step_host": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int32, C.POINTER(C.POINTER(FloeSoA)),
                                       C.POINTER(C.POINTER(FloeSoA))]),
        "slab_step_host_partial": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int32, C.POINTER(C.POINTER(FloeSoA)),
                                               C.POINTER(C.POINTER(FloeSoA))]),
        "slab_max_displacement": (C.c_int32, [C.c_void_p, c_double_p]),
        "slab_rebuild": (C.c_int32, [C.c_void_p]),
        "slab_refresh_halo": (C.c_int32, [C.c_void_p]),
        "slab_stats": (C.c_int32, [C.c_void_p, C.c_int32, c_i64_p, c_i64_p, c_i64_p]),
    }

    def __init__(self, path, prefix="sz_"):
        if not os.path.exists(path):
            raise FileNotFoundError(
                "%s not found — build it first (python -c 'import __graft_entry__ as g; g.build()')" % path)
        self.path, self.prefix = path, prefix
        self.dll = C.CDLL(path)
        for name, (res, args) in self.SIGS.items():
            fn = getattr(self.dll, prefix + name)  # AttributeError if a symbol is missing
            fn.restype, fn.argtypes = res, args
            setattr(self, name, fn)

    def exported(self):
        return [self.prefix + n for n in self.SIGS]

    def default_config_struct(self):
        cfg = Config()
        self.default_config(C.byref(cfg))
        return cfg


_product = None


def product():
    """The CUDA product library.  Fails loudly when it is missing: no CPU fallback exists."""
    global _product
    if _product is None:
        _product = Library(PRODUCT_LIB, "sz_")
    return _product


class FloeArrays:
    """Host-side SoA container matching sz_floe_soa (numpy arrays, caller-owned)."""

    def __init__(self, n, n_init=None):
        self.n = int(n)
        self.n_init = self.n if n_init is None else int(n_init)
        for name in DOUBLE_FIELDS:
            w = FIELD_WIDTH.get(name, 1)
            setattr(self, name, np.zeros((self.n, w) if w > 1 else self.n, dtype=np.float64))
        self.status_tag = np.full(self.n, STATUS_ACTIVE, dtype=np.int32)
        self.id = np.arange(1, self.n + 1, dtype=np.int64)
        self.ghost_id = np.zeros(self.n, dtype=np.int64)
        self.ghost_offsets = np.zeros(self.n + 1, dtype=np.int64)
        self.ghost_index = np.zeros(0, dtype=np.int64)
        self.vert_offsets = np.zeros(self.n + 1, dtype=np.int64)
        self.vert_xy = np.zeros((0, 2), dtype=np.float64)
        self.mc_offsets = np.zeros(self.n + 1, dtype=np.int64)
        self.mc_x = np.zeros(0, dtype=np.float64)
        self.mc_y = np.zeros(0, dtype=np.float64)

    def as_struct(self):
        """ctypes view of the arrays.  Cached: rebuilt only when an array object was replaced."""
        key = tuple(id(getattr(self, n)) for n in DOUBLE_FIELDS + ("vert_xy", "mc_x", "mc_y", "status_tag", "id", "ghost_id",
                                                                   "ghost_offsets", "ghost_index", "vert_offsets", "mc_offsets"))
        cached = getattr(self, "_struct_cache", None)
        if cached is not None and cached[0] == key and cached[1].n == self.n and cached[1].n_init == self.n_init:
            return cached[1]
        s = self._build_struct()
        key = tuple(id(getattr(self, n)) for n in DOUBLE_FIELDS + ("vert_xy", "mc_x", "mc_y", "status_tag", "id", "ghost_id",
                                                                   "ghost_offsets", "ghost_index", "vert_offsets", "mc_offsets"))
        self._struct_cache = (key, s)
        return s

    def _build_struct(self):
        s = FloeSoA()
        s.n, s.n_init = self.n, self.n_init
        keep = []
        for name in DOUBLE_FIELDS + ("vert_xy", "mc_x", "mc_y"):
            a = np.ascontiguousarray(getattr(self, name), dtype=np.float64)
            setattr(self, name, a)
            keep.append(a)
            setattr(s, name, _dp(a))
        a = np.ascontiguousarray(self.status_tag, dtype=np.int32)
        self.status_tag = a
        s.status_tag = a.ctypes.data_as(c_i32_p)
        for name in ("id", "ghost_id", "ghost_offsets", "ghost_index", "vert_offsets", "mc_offsets"):
            a = np.ascontiguousarray(getattr(self, name), dtype=np.int64)
            setattr(self, name, a)
            setattr(s, name, _ip(a))
        s._keep = keep
        return s

    def adopt(self, fa):
        """Take over the arrays of a downloaded FloeArrays (host-only attributes are kept)."""
        keep = {k: getattr(self, k) for k in ("interactions", "num_inters", "fuse_idx", "warnings") if hasattr(self, k)}
        self.__dict__.update(fa.__dict__)
        self.__dict__.update(keep)

    def ring(self, i):
        return self.vert_xy[self.vert_offsets[i]:self.vert_offsets[i + 1]]

    def centroid(self, i):
        return np.array([self.centroid_x[i], self.centroid_y[i]])

    def coords(self, i):
        return self.ring(i)

    def ghosts(self, i):
        return list(self.ghost_index[self.ghost_offsets[i]:self.ghost_offsets[i + 1]])


def marshal_fields(Nx, Ny, arrays):
    """Five (Nx+1, Ny+1) arrays indexed [x, y] -> flat arrays with element [ix + (Nx+1) iy]."""
    out = []
    for a in arrays:
        a = np.asarray(a, dtype=np.float64)
        assert a.shape == (Nx + 1, Ny + 1), (a.shape, Nx, Ny)
        out.append(np.asfortranarray(a).ravel(order="F"))
    return out


def marshal_domain(kinds, vals, uv, rect, topo_rings=(), topo_centroid=None, topo_rmax=None):
    """Arguments of sz_set_domain / sz_slab_set_domain behind the handle pointer (arrays are kept alive by the tuple)."""
    kinds = np.ascontiguousarray(kinds, dtype=np.int32)
    vals = np.ascontiguousarray(vals, dtype=np.float64)
    uv = np.ascontiguousarray(uv, dtype=np.float64).reshape(8)
    rect = np.ascontiguousarray(rect, dtype=np.float64).reshape(16)
    nt = len(topo_rings)
    offs = np.zeros(nt + 1, dtype=np.int64)
    for k, r in enumerate(topo_rings):
        offs[k + 1] = offs[k] + len(r)
    xy = (np.concatenate([np.asarray(r, dtype=np.float64) for r in topo_rings])
          if nt else np.zeros((0, 2)))
    xy = np.ascontiguousarray(xy, dtype=np.float64)
    cen = np.ascontiguousarray(topo_centroid if nt else np.zeros((0, 2)), dtype=np.float64)
    rm = np.ascontiguousarray(topo_rmax if nt else np.zeros(0), dtype=np.float64)
    keep = (kinds, vals, uv, rect, offs, xy, cen, rm)
    return (kinds.ctypes.data_as(c_i32_p), _dp(vals), _dp(uv), _dp(rect), nt, _ip(offs), _dp(xy), _dp(cen), _dp(rm)), keep


class Handle:
    """RAII wrapper of sz_handle for one Library."""

    @classmethod
    def borrow(cls, lib, ptr, Nx=None, Ny=None):
        """A non-owning wrapper of a handle that belongs to a sz_slab."""
        self = cls.__new__(cls)
        self.lib, self.cfg, self.h, self.borrowed = lib, None, C.c_void_p(ptr), True
        if Nx is not None:
            self.Nx, self.Ny = Nx, Ny
        return self

    def __init__(self, lib, cfg=None, **overrides):
        self.lib = lib
        self.cfg = cfg if cfg is not None else lib.default_config_struct()
        for k, v in overrides.items():
            if not hasattr(self.cfg, k):
                raise AttributeError("sz_config has no field %r" % k)
            setattr(self.cfg, k, v)
        self.h = C.c_void_p()
        rc = lib.create(C.byref(self.cfg), C.byref(self.h))
        if rc != 0:
            raise SubzeroError(rc, "sz_create failed (library %s)" % lib.path)

    def close(self):
        if self.h and not getattr(self, "borrowed", False):
            self.lib.destroy(self.h)
        self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            msg = self.lib.last_error(self.h)
            raise SubzeroError(rc, msg.decode() if msg else "")

    # model description ------------------------------------------------------------------
    def set_grid(self, Nx, Ny, x0, xf, y0, yf):
        self._ck(self.lib.set_grid(self.h, Nx, Ny, x0, xf, y0, yf))
        self.Nx, self.Ny = Nx, Ny

    def set_fields(self, ocean_u, ocean_v, ocean_hflx, atmos_u, atmos_v):
        arrs = marshal_fields(self.Nx, self.Ny, (ocean_u, ocean_v, ocean_hflx, atmos_u, atmos_v))
        self._ck(self.lib.set_fields(self.h, *[_dp(a) for a in arrs]))

    def set_temperatures(self, ocean_temp, atmos_temp):
        arrs = []
        for a in (ocean_temp, atmos_temp):
            a = np.asarray(a, dtype=np.float64)
            assert a.shape == (self.Nx + 1, self.Ny + 1), (a.shape, self.Nx, self.Ny)
            arrs.append(np.asfortranarray(a).ravel(order="F"))
        self._ck(self.lib.set_temperatures(self.h, *[_dp(a) for a in arrs]))

    def ocean_fields(self):
        """(tau_x, tau_y, si_frac, hflx_factor), each (Nx+1, Ny+1) indexed [x, y]."""
        out = [np.zeros((self.Nx + 1) * (self.Ny + 1)) for _ in range(4)]
        self._ck(self.lib.get_ocean_fields(self.h, *[_dp(a) for a in out]))
        return [a.reshape((self.Nx + 1, self.Ny + 1), order="F") for a in out]

    def cell_floes(self):
        """The floe -> cell registry: (cell_xy [n,2] 1-based, floe [n] 1-based, vals [n,5])."""
        n = C.c_int64(0)
        self._ck(self.lib.get_cell_floes(self.h, C.byref(n), None, None, None))
        cell = np.zeros((max(n.value, 1), 2), dtype=np.int64)
        floe = np.zeros(max(n.value, 1), dtype=np.int64)
        vals = np.zeros((max(n.value, 1), 5))
        self._ck(self.lib.get_cell_floes(self.h, C.byref(n), _ip(cell), _ip(floe), _dp(vals)))
        return cell[:n.value], floe[:n.value], vals[:n.value]

    def set_domain(self, kinds, vals, uv, rect, topo_rings=(), topo_centroid=None, topo_rmax=None):
        args, keep = marshal_domain(kinds, vals, uv, rect, topo_rings, topo_centroid, topo_rmax)
        self._ck(self.lib.set_domain(self.h, *args))
        del keep

    def get_domain(self):
        vals, rect = np.zeros(4), np.zeros(16)
        self._ck(self.lib.get_domain(self.h, _dp(vals), _dp(rect)))
        return vals, rect.reshape(4, 4)

    # floe state ---------------------------------------------------------------------------
    def upload_floes(self, fa):
        s = fa.as_struct()
        self._ck(self.lib.upload_floes(self.h, C.byref(s)))

    def upload_state(self, fa):
        s = fa.as_struct()
        self._ck(self.lib.upload_state(self.h, C.byref(s)))

    def counts(self):
        c = Counts()
        self._ck(self.lib.get_counts(self.h, C.byref(c)))
        return c.asdict()

    def download_floes(self, into=None, mc=True):
        """Download the floe list.  `into`: a FloeArrays of matching sizes to reuse (e.g. backed
        by pinned memory); mc=False skips the (static) Monte-Carlo points."""
        if into is None:
            c = self.counts()
            fa = FloeArrays(c["n_total"], c["n_init"])
            fa.vert_xy = np.zeros((c["n_vertices"], 2))
            fa.mc_x = np.zeros(c["n_mc"] if mc else 0)
            fa.mc_y = np.zeros(c["n_mc"] if mc else 0)
            fa.ghost_index = np.zeros(c["n_ghost_links"], dtype=np.int64)
        else:
            fa = into
        s = fa.as_struct()
        if not mc:  # a private copy of the (cached) struct with the Monte-Carlo pointers cleared
            s2 = FloeSoA()
            C.memmove(C.byref(s2), C.byref(s), C.sizeof(FloeSoA))
            s2.mc_x, s2.mc_y = None, None
            s = s2
        self._ck(self.lib.download_floes(self.h, C.byref(s)))
        return fa

    # hot path -------------------------------------------------------------------------------
    def add_ghosts(self):
        n = C.c_int64(0)
        self._ck(self.lib.add_ghosts(self.h, C.byref(n)))
        return n.value

    def step_collisions(self):
        self._ck(self.lib.step_collisions(self.h))

    def remove_ghosts(self):
        self._ck(self.lib.remove_ghosts(self.h))

    def step_coupling(self):
        self._ck(self.lib.step_coupling(self.h))

    def step_floe_properties(self, tstep=0):
        self._ck(self.lib.step_floe_properties(self.h, tstep))

    def step(self, tstep=0, do_coupling=True):
        self._ck(self.lib.step(self.h, tstep, 1 if do_coupling else 0))

    def step_host(self, fa, tstep=0, do_coupling=True, out=None):
        """One timestep on host arrays (upload_state + step + download in one call, copies overlapped with the
        kernels).  `out` defaults to `fa` (in-place)."""
        if fa is None:  # the uploads were enqueued by upload_state_begin
            o = out.as_struct()
            self._ck(self.lib.step_host(self.h, tstep, 1 if do_coupling else 0, None, C.byref(o)))
            return out
        s = fa.as_struct()
        o = s if out is None or out is fa else out.as_struct()
        self._ck(self.lib.step_host(self.h, tstep, 1 if do_coupling else 0, C.byref(s), C.byref(o)))
        return fa if out is None else out

    def step_host_partial(self, fa, tstep=0, do_coupling=True, upload=(), download=()):
        """One timestep exchanging only the named fields with the host arrays `fa` (FloeSoA field names): the others keep
        their device-resident values / are not downloaded."""
        def masked(names):
            full = fa.as_struct()
            m = FloeSoA()
            m.n, m.n_init = full.n, full.n_init
            for name in names:
                setattr(m, name, getattr(full, name))
            m._keep = full
            return m
        i = masked(upload) if upload else None
        o = masked(download)
        self._ck(self.lib.step_host_partial(self.h, tstep, 1 if do_coupling else 0, C.byref(i) if i is not None else None, C.byref(o)))
        return fa

    def coupling_begin(self):
        """Slab ranks: start the coming step's one-way coupling before the halo exchange (no-op where the order matters)."""
        self._ck(self.lib.coupling_begin(self.h))

    def upload_state_begin(self, fa, do_coupling=True):
        """The upload half of step_host (returns at once); follow with the halo exchange and step_host(None, ..., out=fa)."""
        s = fa.as_struct()
        self._ck(self.lib.upload_state_begin(self.h, 1 if do_coupling else 0, C.byref(s)))

    # results ----------------------------------------------------------------------------------
    def interactions(self):
        c = self.counts()
        offs = np.zeros(c["n_total"] + 1, dtype=np.int64)
        rows = np.zeros((max(c["n_rows"], 1), 7))
        self._ck(self.lib.get_interactions(self.h, _ip(offs), _dp(rows)))
        return offs, rows[:c["n_rows"]]

    def set_interactions(self, rows_per_floe):
        n = len(rows_per_floe)
        offs = np.zeros(n + 1, dtype=np.int64)
        for i, r in enumerate(rows_per_floe):
            offs[i + 1] = offs[i] + len(r)
        rows = (np.concatenate([np.asarray(r, dtype=np.float64).reshape(-1, 7) for r in rows_per_floe])
                if offs[-1] else np.zeros((1, 7)))
        rows = np.ascontiguousarray(rows)
        self._ck(self.lib.set_interactions(self.h, _ip(offs), _dp(rows)))

    def floe_interactions(self, i):
        """Rows of floe i (0-based), shape (k, 7)."""
        offs, rows = self.interactions()
        return rows[offs[i]:offs[i + 1]]

    def pairs(self, which):
        key = ("n_candidates", "n_pairs", "n_overlap", "n_fuse")[which]
        m = self.counts()[key]
        out = np.zeros((max(m, 1), 2), dtype=np.int64)
        self._ck(self.lib.get_pairs(self.h, which, _ip(out)))
        return out[:m]

    def warnings(self):
        n = self.counts()["n_init"]
        out = np.zeros(max(n, 1), dtype=np.uint32)
        self._ck(self.lib.get_warnings(self.h, out.ctypes.data_as(c_u32_p)))
        return out[:n]

    def timings_raw(self):
        ms = np.zeros(8)
        self._ck(self.lib.get_timings(self.h, _dp(ms)))
        return ms

    def timings(self):
        ms = np.zeros(8)
        self._ck(self.lib.get_timings(self.h, _dp(ms)))
        return dict(zip(("ghosts", "broad", "narrow", "reduce", "coupling", "update", "total"), ms[:7]))

    # slab decomposition ---------------------------------------------------------------------------
    def halo_configure(self, lists):
        """lists: sequences of 0-based local floe indices (one per exchange partner and direction)."""
        offs = np.zeros(len(lists) + 1, dtype=np.int64)
        for k, l in enumerate(lists):
            offs[k + 1] = offs[k] + len(l)
        idx = (np.concatenate([np.asarray(l, dtype=np.int64) for l in lists]) + 1 if offs[-1] else np.zeros(1, dtype=np.int64))
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        self._ck(self.lib.halo_configure(self.h, len(lists), _ip(offs), _ip(idx)))

    def halo_bytes(self, k):
        b = C.c_int64(0)
        self._ck(self.lib.halo_bytes(self.h, k, C.byref(b)))
        return b.value

    def halo_pack(self, k, ptr, capacity):
        self._ck(self.lib.halo_pack(self.h, k, C.c_void_p(ptr), capacity))

    def halo_unpack(self, k, ptr, nbytes):
        self._ck(self.lib.halo_unpack(self.h, k, C.c_void_p(ptr), nbytes))

    def halo_pack_on(self, k, ptr, capacity, stream):
        """Stream-ordered pack (no host synchronisation); stream: cudaStream_t as an integer."""
        self._ck(self.lib.halo_pack_on(self.h, k, C.c_void_p(ptr), capacity, C.c_void_p(stream)))

    def halo_unpack_on(self, k, ptr, nbytes, stream):
        self._ck(self.lib.halo_unpack_on(self.h, k, C.c_void_p(ptr), nbytes, C.c_void_p(stream)))

    # services for the host-side processes -------------------------------------------------------------
    def pair_overlap_areas(self, pairs):
        """pairs: [n, 2] 1-based ordered floe pairs -> (areas [n], interacts [n] bool):
        potential_interaction and sum(GO.area, intersect_polys(poly_i, poly_j))."""
        pairs = np.ascontiguousarray(pairs, dtype=np.int64).reshape(-1, 2)
        n = len(pairs)
        areas = np.zeros(n)
        inter = np.zeros(n, dtype=np.uint8)
        self._ck(self.lib.pair_overlap_areas(self.h, n, _ip(pairs), _dp(areas), inter.ctypes.data_as(c_u8_p)))
        return areas, inter.astype(bool)

    def eulerian_data(self, xg, yg, kinds):
        """calc_eulerian_data! on the grid lines xg, yg -> data [nx, ny, n_out] (writer.data layout)."""
        xg = np.ascontiguousarray(xg, dtype=np.float64)
        yg = np.ascontiguousarray(yg, dtype=np.float64)
        kinds = np.ascontiguousarray(kinds, dtype=np.int32)
        nx, ny = len(xg) - 1, len(yg) - 1
        data = np.zeros(nx * ny * len(kinds))
        self._ck(self.lib.eulerian_data(self.h, nx, ny, _dp(xg), _dp(yg), len(kinds), kinds.ctypes.data_as(c_i32_p), _dp(data)))
        return data.reshape((nx, ny, len(kinds)), order="F")

    def generate_subfloe_points(self, kind, npoints=1000, err=0.1, delta_g=0.0, seed=0, floes=None, install=False):
        """generate_subfloe_points (coupling.jl:172-321) for the resident floes (`floes`: 1-based indices, None = all)
        -> (offsets [n+1], x, y, status [n]) in the body frame."""
        g = PointsGenerator(kind, int(npoints), float(err), float(delta_g), int(seed))
        if floes is None:
            n, fp = self.counts()["n_init"], None
        else:
            fl = np.ascontiguousarray(floes, dtype=np.int64)
            n, fp = len(fl), _ip(fl)
        offs = np.zeros(n + 1, dtype=np.int64)
        status = np.zeros(max(n, 1), dtype=np.int32)
        sp = status.ctypes.data_as(c_i32_p)
        self._ck(self.lib.generate_subfloe_points(self.h, C.byref(g), n, fp, _ip(offs), None, None, 0, sp, 0))  # sizes
        m = int(offs[-1])
        x, y = np.zeros(max(m, 1)), np.zeros(max(m, 1))
        self._ck(self.lib.generate_subfloe_points(self.h, C.byref(g), n, fp, _ip(offs), _dp(x), _dp(y), m, sp, 1 if install else 0))
        return offs, x[:m], y[:m], status[:n]

    def clip_polygons(self, p, q, cap_regions=64, cap_points=8192):
        p = np.ascontiguousarray(p, dtype=np.float64)
        q = np.ascontiguousarray(q, dtype=np.float64)
        offs = np.zeros(cap_regions + 1, dtype=np.int32)
        xy = np.zeros((cap_points, 2))
        areas = np.zeros(cap_regions)
        n = self.lib.clip_polygons(self.h, _dp(p), len(p), _dp(q), len(q), cap_regions, cap_points,
                                   offs.ctypes.data_as(c_i32_p), _dp(xy), _dp(areas))
        if n < 0:
            raise SubzeroError(n, "clip_polygons")
        return [xy[offs[r]:offs[r + 1]].copy() for r in range(n)], areas[:n].copy()
