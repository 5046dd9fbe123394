"""Slab decomposition: Python view of the library's sz_slab_* API.

The partition, the halo lists, the per-step halo update, the staleness test and the rebuild with
migration of ownership all live in the C library (csrc/sz_slab.cpp + the push / unpack kernels of
csrc/sz_kernels_fp.cu); this module only marshals arrays and, when every rank is its own process
(torchrun), lends the library a byte transport for its set-up / rebuild messages
(`torch.distributed` point-to-point on a gloo group).  A Julia host binds the same entry points
with `ccall` and needs none of this file.

    Slab(lib, field, world)                       one process drives every rank (emulated ranks / one host, k GPUs)
    Slab(lib, field, world, rank=r, group=g)      one rank per process

The floe-list helpers at the bottom (extract / concat / shift_x) are used by tests and bench.py.
"""
import ctypes as C

import numpy as np

from . import capi

ALLTOALLV = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_void_p, C.POINTER(C.c_int64))


def _make_transport(group, rank, world):
    """MPI_Alltoallv on bytes over torch.distributed point-to-point (gloo: host tensors).  Returns the ctypes callback."""
    import torch
    import torch.distributed as dist

    def cb(ctx, send, send_off, recv, recv_off):
        try:
            so = [send_off[i] for i in range(world + 1)]
            ro = [recv_off[i] for i in range(world + 1)]
            sbuf = np.ctypeslib.as_array((C.c_uint8 * max(so[-1], 1)).from_address(send))
            rbuf = np.ctypeslib.as_array((C.c_uint8 * max(ro[-1], 1)).from_address(recv))
            ops, keep = [], []
            for d in range(world):
                ns, nr = so[d + 1] - so[d], ro[d + 1] - ro[d]
                if d == rank:
                    rbuf[ro[d]:ro[d + 1]] = sbuf[so[d]:so[d + 1]]
                    continue
                if ns:
                    t = torch.from_numpy(sbuf[so[d]:so[d + 1]].copy())
                    keep.append(t)
                    ops.append(dist.P2POp(dist.isend, t, dist.get_global_rank(group, d) if group is not None else d, group))
                if nr:
                    t = torch.empty(nr, dtype=torch.uint8)
                    keep.append((t, ro[d], ro[d + 1]))
                    ops.append(dist.P2POp(dist.irecv, t, dist.get_global_rank(group, d) if group is not None else d, group))
            if ops:
                for r in dist.batch_isend_irecv(ops):
                    r.wait()
            for k in keep:
                if isinstance(k, tuple):
                    rbuf[k[1]:k[2]] = k[0].numpy()
            return 0
        except Exception:  # pragma: no cover - reported through the library's status
            import traceback
            traceback.print_exc()
            return 1

    return ALLTOALLV(cb)


class Slab:
    """RAII wrapper of sz_slab.  `field` supplies grid, ocean / atmosphere, domain and constants (synth.Field);
    floes are handed over with build()."""

    def __init__(self, lib, field, world, rank=None, group=None, devices=None, skin=3000.0, dt=10, **overrides):
        from . import host
        self.lib, self.world = lib, int(world)
        self.rank_first = 0 if rank is None else int(rank)
        self.n_local = self.world if rank is None else 1
        proto = host._make_handle(lib, field.consts, dt, None, None, None, **overrides)
        cfg = proto.cfg
        proto.close()
        devices = list(devices) if devices is not None else [int(getattr(cfg, "device", 0))] * self.n_local
        dev = (C.c_int32 * self.n_local)(*devices)
        self._cb = _make_transport(group, self.rank_first, self.world) if rank is not None else None
        self.s = C.c_void_p()
        rc = lib.slab_create(C.byref(cfg), self.world, self.rank_first, self.n_local, dev, float(skin),
                             C.cast(self._cb, C.c_void_p) if self._cb else None, None, C.byref(self.s))
        if rc != 0:
            raise capi.SubzeroError(rc, "sz_slab_create failed (library %s)" % lib.path)
        self.skin = float(skin)
        g = field.grid
        self.Nx, self.Ny = g.Nx, g.Ny
        self._ck(lib.slab_set_grid(self.s, g.Nx, g.Ny, g.x0, g.xf, g.y0, g.yf))
        arrs = capi.marshal_fields(g.Nx, g.Ny, (field.ocean.u, field.ocean.v, field.ocean.hflx_factor, field.atmos.u, field.atmos.v))
        self._ck(lib.slab_set_fields(self.s, *[capi._dp(a) for a in arrs]))
        field.domain.push(self)  # calls set_domain below
        self.handles = []
        for k in range(self.n_local):
            p = C.c_void_p()
            self._ck(lib.slab_handle(self.s, k, C.byref(p)))
            self.handles.append(capi.Handle.borrow(lib, p.value, g.Nx, g.Ny))

    def _ck(self, rc):
        if rc != 0:
            msg = self.lib.slab_last_error(self.s)
            raise capi.SubzeroError(rc, msg.decode() if msg else "")

    def close(self):
        if self.s:
            self.lib.slab_destroy(self.s)
            self.s = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_domain(self, kinds, vals, uv, rect, topo_rings=(), topo_centroid=None, topo_rmax=None):
        args, keep = capi.marshal_domain(kinds, vals, uv, rect, topo_rings, topo_centroid, topo_rmax)
        self._ck(self.lib.slab_set_domain(self.s, *args))
        del keep

    def set_edges(self, edges):
        e = np.ascontiguousarray(edges, dtype=np.float64)
        assert len(e) == self.world + 1
        self._ck(self.lib.slab_set_edges(self.s, capi._dp(e)))

    def build(self, floes, gidx=None):
        """floes: one FloeArrays per local rank (None = nothing), gidx: their 0-based global indices.  A single FloeArrays
        is the whole list handed to local rank 0 (every floe migrates to its slab inside the library)."""
        if isinstance(floes, capi.FloeArrays):
            floes = [floes] + [None] * (self.n_local - 1)
            gidx = [np.arange(floes[0].n, dtype=np.int64) if gidx is None else gidx] + [None] * (self.n_local - 1)
        soa = (C.POINTER(capi.FloeSoA) * self.n_local)()
        gp = (capi.c_i64_p * self.n_local)()
        keep = []
        for k in range(self.n_local):
            if floes[k] is None:
                continue
            s = floes[k].as_struct()
            g = np.ascontiguousarray(gidx[k], dtype=np.int64)
            keep.append((s, g))
            soa[k] = C.pointer(s)
            gp[k] = capi._ip(g)
        self._ck(self.lib.slab_build(self.s, soa, gp))

    def step(self, tstep=0, do_coupling=True):
        self._ck(self.lib.slab_step(self.s, tstep, 1 if do_coupling else 0))

    def step_host(self, arrays, tstep=0, do_coupling=True):
        """arrays: one FloeArrays per local rank in the layout of local_index(k); updated in place."""
        ptr = (C.POINTER(capi.FloeSoA) * self.n_local)()
        keep = []
        for k in range(self.n_local):
            s = arrays[k].as_struct()
            keep.append(s)
            ptr[k] = C.pointer(s)
        self._ck(self.lib.slab_step_host(self.s, tstep, 1 if do_coupling else 0, ptr, ptr))

    def step_host_partial(self, arrays, tstep=0, do_coupling=True, upload=(), download=()):
        """step_host exchanging only the named fields (FloeSoA field names) of every local rank's arrays: the others keep
        their device-resident values / are not downloaded (sz_slab_step_host_partial)."""
        def masked(fa, names):
            full = fa.as_struct()
            m = capi.FloeSoA()
            m.n, m.n_init = full.n, full.n_init
            for name in names:
                setattr(m, name, getattr(full, name))
            m._keep = full
            return m
        pin = (C.POINTER(capi.FloeSoA) * self.n_local)()
        pout = (C.POINTER(capi.FloeSoA) * self.n_local)()
        keep = []
        for k in range(self.n_local):
            o = masked(arrays[k], download)
            keep.append(o)
            pout[k] = C.pointer(o)
            if upload:
                i = masked(arrays[k], upload)
                keep.append(i)
                pin[k] = C.pointer(i)
        self._ck(self.lib.slab_step_host_partial(self.s, tstep, 1 if do_coupling else 0, pin if upload else None, pout))

    def max_displacement(self):
        d = C.c_double(0.0)
        self._ck(self.lib.slab_max_displacement(self.s, C.byref(d)))
        return d.value

    def stale(self):
        return self.max_displacement() > 0.5 * self.skin

    def rebuild(self):
        self._ck(self.lib.slab_rebuild(self.s))

    def refresh_halo(self):
        """Collective: halo copies := their owners' current state (no step)."""
        self._ck(self.lib.slab_refresh_halo(self.s))

    def local_index(self, k=0):
        """(global indices, owning ranks) of rank k's local list."""
        n, no = C.c_int64(0), C.c_int64(0)
        self._ck(self.lib.slab_local_count(self.s, k, C.byref(n), C.byref(no)))
        g = np.zeros(max(n.value, 1), dtype=np.int64)
        o = np.zeros(max(n.value, 1), dtype=np.int32)
        self._ck(self.lib.slab_local_index(self.s, k, capi._ip(g), o.ctypes.data_as(capi.c_i32_p)))
        return g[:n.value], o[:n.value]

    def stats(self, k=0):
        a, b, c = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        self._ck(self.lib.slab_stats(self.s, k, C.byref(a), C.byref(b), C.byref(c)))
        return {"send_bytes_per_step": a.value, "halo_floes": b.value, "rebuilds": c.value}

    def owned_state(self, k=0):
        """(global indices, FloeArrays) of the floes local rank k owns, downloaded from its handle."""
        g, o = self.local_index(k)
        fa = self.handles[k].download_floes(mc=False)
        idx = np.nonzero(o == self.rank_first + k)[0]
        return g[idx], extract(fa, idx)


# ---- floe-list helpers (tests, bench.py) ------------------------------------------------------------------------
def extract(fa, idx):
    """Sub-list of a FloeArrays (CSR rings and Monte-Carlo points included), in the order of idx."""
    idx = np.asarray(idx, dtype=np.int64)
    out = capi.FloeArrays(len(idx))
    for name in capi.DOUBLE_FIELDS:
        setattr(out, name, np.ascontiguousarray(getattr(fa, name)[idx]))
    out.status_tag = np.ascontiguousarray(fa.status_tag[idx])
    out.id = np.ascontiguousarray(fa.id[idx])
    out.ghost_id = np.ascontiguousarray(fa.ghost_id[idx])

    def csr(offs, *arrs):
        cnt = (offs[1:] - offs[:-1])[idx]
        no = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
        if no[-1] == 0 or len(arrs[0]) < offs[-1]:  # e.g. a download without Monte-Carlo points
            return np.zeros(len(idx) + 1, dtype=np.int64), [a[:0] for a in arrs]
        src = np.repeat(offs[:-1][idx] - no[:-1], cnt) + np.arange(no[-1])
        return no, [np.ascontiguousarray(a[src]) for a in arrs]

    out.vert_offsets, (out.vert_xy,) = csr(fa.vert_offsets, fa.vert_xy)
    out.mc_offsets, (out.mc_x, out.mc_y) = csr(fa.mc_offsets, fa.mc_x, fa.mc_y)
    return out


def concat(parts):
    n = sum(p.n for p in parts)
    out = capi.FloeArrays(n)
    for name in capi.DOUBLE_FIELDS + ("status_tag", "id", "ghost_id", "vert_xy", "mc_x", "mc_y"):
        setattr(out, name, np.concatenate([getattr(p, name) for p in parts]))
    for oname in ("vert_offsets", "mc_offsets"):
        offs, base = [np.zeros(1, dtype=np.int64)], 0
        for p in parts:
            o = getattr(p, oname)
            offs.append(o[1:] + base)
            base += o[-1]
        setattr(out, oname, np.concatenate(offs))
    return out


def shift_x(fa, dx):
    """Translate a floe list in x (tiles of a weak-scaling field)."""
    fa.centroid_x = fa.centroid_x + dx
    fa.vert_xy = fa.vert_xy.copy()
    fa.vert_xy[:, 0] += dx
    return fa


def tile_edges(world, tile_L, periodic):
    """Slab boundaries of `world` tiles of width tile_L side by side in x (bench.py's weak-scaling field)."""
    lo, hi = (0.0, world * tile_L) if periodic else (-np.inf, np.inf)
    return np.concatenate([[lo], tile_L * np.arange(1, world), [hi]])
