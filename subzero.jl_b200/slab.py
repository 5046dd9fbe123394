"""1-D spatial slab decomposition with halo floes.

The reference is single-process; this is new.  One rank (one process, one GPU, one C-ABI handle)
owns the floes whose centroid lies in its x-slab and additionally holds copies ("halo floes") of
every other rank's floes that can touch one of its own: |x - slab| < rmax + rmax_max + skin, with
the east/west period taken into account so that the reference's periodic ghost floes
(collisions.jl:925-952) appear on the rank that needs them simply because `add_ghosts!` runs on
the local list.  The local list is sorted by GLOBAL floe index, so pair orientation (i < j is
polygon 1), candidate order, row order and the canonical image pair are those of the single-rank
run: results of owned floes are bit-identical to 1 GPU.

Every step the owner sends the dynamic state of the floes another rank holds copies of
(`sz_halo_pack` -> NCCL/gloo send/recv through torch.distributed -> `sz_halo_unpack`); halo
results computed locally are discarded (cross-slab pairs are evaluated redundantly on both sides
with the same orientation, which removes the return exchange of the north_star sketch).
The floe LIST of a rank is fixed between rebuilds (Verlet-list style): `skin` is the distance floes
may travel before `stale()` asks for a rebuild from the host (download, repartition, upload).
"""
import numpy as np

from . import capi

DOUBLE1 = [n for n in capi.DOUBLE_FIELDS if n not in capi.FIELD_WIDTH]


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


def strip_mc(fa):
    """Halo copies need no Monte-Carlo points (coupling results of halo floes are discarded)."""
    fa.mc_offsets = np.zeros(fa.n + 1, dtype=np.int64)
    fa.mc_x = np.zeros(0)
    fa.mc_y = np.zeros(0)
    return fa


def x_distance(cx, xa, xb, period):
    """Distance in x from points to the interval [xa, xb), minimised over the periodic images."""
    def d(x):
        return np.maximum(np.maximum(xa - x, x - xb), 0.0)
    out = d(cx)
    if period:
        out = np.minimum(out, np.minimum(d(cx + period), d(cx - period)))
    return out


class SlabRank:
    """The local floe list, halo lists and C-ABI handle of one rank."""

    def __init__(self, rank, world, edges, period_x, skin, rmax_max, period_y=None):
        self.rank, self.world, self.period_y = rank, world, period_y
        self.edges = np.asarray(edges, dtype=np.float64)  # world + 1 slab boundaries; outermost are +-inf
        self.period_x, self.skin, self.rmax_max = period_x, float(skin), float(rmax_max)
        self.h = None

    def interval(self, r=None):
        r = self.rank if r is None else r
        return self.edges[r], self.edges[r + 1]

    def owner_of(self, cx):
        return np.clip(np.searchsorted(self.edges, cx, side="right") - 1, 0, self.world - 1)

    def needs(self, cx, rmax, r=None):
        xa, xb = self.interval(r)
        return x_distance(cx, xa, xb, self.period_x) < rmax + self.rmax_max + self.skin

    def build(self, known, gidx, owner):
        """known: FloeArrays of every floe this rank may need (its own and candidates of others),
        gidx: their global indices, owner: their owning ranks."""
        mine = owner == self.rank
        take = mine | self.needs(known.centroid_x, known.rmax)
        sel = np.nonzero(take)[0]
        sel = sel[np.argsort(gidx[sel], kind="stable")]
        self.gidx = gidx[sel]
        self.owner = owner[sel]
        self.owned = self.owner == self.rank
        own_part = extract(known, sel)
        # halo copies carry no Monte-Carlo points
        cnt = np.diff(own_part.mc_offsets)
        cnt[~self.owned] = 0
        keep = np.repeat(self.owned, np.diff(own_part.mc_offsets))
        own_part.mc_x, own_part.mc_y = own_part.mc_x[keep], own_part.mc_y[keep]
        own_part.mc_offsets = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
        self.local = own_part
        self.x0 = own_part.centroid_x.copy()
        self.y0 = own_part.centroid_y.copy()
        # receive lists: my halo floes grouped by owner, ascending global index
        self.recv = {int(s): np.nonzero(self.owner == s)[0] for s in np.unique(self.owner) if s != self.rank}
        return self

    def set_send_lists(self, wanted):
        """wanted[s] = global indices rank s holds copies of and this rank owns (ascending)."""
        lookup = {int(g): k for k, g in enumerate(self.gidx)}
        self.send = {int(s): np.array([lookup[int(g)] for g in gl], dtype=np.int64) for s, gl in wanted.items() if len(gl)}
        return self

    def attach(self, handle):
        """Upload the local list and register the halo lists: list 2k = send to partner k, 2k+1 = receive."""
        self.h = handle
        handle.upload_floes(self.local)
        self.partners = sorted(set(self.send) | set(self.recv))
        lists = []
        for s in self.partners:
            lists.append(self.send.get(s, np.zeros(0, dtype=np.int64)))
            lists.append(self.recv.get(s, np.zeros(0, dtype=np.int64)))
        handle.halo_configure(lists)
        self.nbytes = [handle.halo_bytes(k) for k in range(len(lists))]
        return self

    # ---- per-step exchange ----------------------------------------------------------------------
    def make_buffers(self, device):
        import torch
        self._plan = None
        self.sbuf = [torch.empty(max(self.nbytes[2 * k], 8), dtype=torch.uint8, device=device) for k in range(len(self.partners))]
        self.rbuf = [torch.empty(max(self.nbytes[2 * k + 1], 8), dtype=torch.uint8, device=device) for k in range(len(self.partners))]

    def exchange(self):
        """pack -> send/recv (torch.distributed: NCCL for CUDA buffers, gloo for host buffers) -> unpack.
        CUDA buffers: everything is ordered on torch's current stream (sz_halo_*_on): the pack kernels, the NCCL
        send/recv pairs, the unpack kernels and the following sz_step run back to back on the device, the host
        does not wait anywhere."""
        import torch.distributed as dist
        cuda = bool(self.sbuf) and self.sbuf[0].is_cuda
        plan = getattr(self, "_plan", None)
        if plan is None:  # the lists are fixed between rebuilds: slices, pointers and P2P descriptors are built once
            sends = [(2 * k, self.sbuf[k].data_ptr(), self.nbytes[2 * k]) for k in range(len(self.partners)) if self.nbytes[2 * k]]
            recvs = [(2 * k + 1, self.rbuf[k].data_ptr(), self.nbytes[2 * k + 1]) for k in range(len(self.partners)) if self.nbytes[2 * k + 1]]
            ops = []
            for k, s in enumerate(self.partners):
                if self.nbytes[2 * k]:
                    ops.append(dist.P2POp(dist.isend, self.sbuf[k][:self.nbytes[2 * k]], s))
                if self.nbytes[2 * k + 1]:
                    ops.append(dist.P2POp(dist.irecv, self.rbuf[k][:self.nbytes[2 * k + 1]], s))
            plan = self._plan = (sends, recvs, ops)
        sends, recvs, ops = plan
        if cuda:
            import torch
            stream = torch.cuda.current_stream().cuda_stream
            for lst, ptr, nb in sends:
                self.h.halo_pack_on(lst, ptr, nb, stream)
        else:
            for lst, ptr, nb in sends:
                self.h.halo_pack(lst, ptr, nb)
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()  # NCCL: the current stream waits for the transfer, the host does not
        if cuda:
            for lst, ptr, nb in recvs:
                self.h.halo_unpack_on(lst, ptr, nb, stream)
        else:
            for lst, ptr, nb in recvs:
                self.h.halo_unpack(lst, ptr, nb)

    def owned_state(self):
        """(global indices, FloeArrays) of the owned floes, downloaded from the handle."""
        fa = self.h.download_floes(mc=False)
        idx = np.nonzero(self.owned)[0]
        return self.gidx[idx], extract(fa, idx)

    def stale(self):
        """True when an owned floe travelled more than skin / 2 since the lists were built."""
        return self.max_displacement() > 0.5 * self.skin

    def max_displacement(self):
        fa = self.h.download_floes(mc=False)
        dx = np.abs(fa.centroid_x[:self.local.n] - self.x0)
        dy = np.abs(fa.centroid_y[:self.local.n] - self.y0)
        if self.period_x:  # add_ghosts! wraps a parent that left the domain (collisions.jl:943-949)
            dx = np.minimum(dx, np.abs(dx - self.period_x))
        if self.period_y:
            dy = np.minimum(dy, np.abs(dy - self.period_y))
        d = np.hypot(dx, dy)[self.owned]
        return float(d.max()) if len(d) else 0.0


# ---- rebuild: migration of ownership + fresh halo lists ------------------------------------------------
# Two collective rounds (all_gather of python objects; set-up path, not per step):
#   1. every rank downloads its owned floes, re-assigns them by centroid and hands migrants (full
#      record incl. Monte-Carlo points) to their new owner;
#   2. with the post-migration owned set, every rank sends the floes each other rank needs as halo.
def _rebuild_round1(me):
    fa = me.h.download_floes(mc=False)
    idx = np.nonzero(me.owned)[0]
    own = extract(fa, idx)
    src = extract(me.local, idx)  # the static Monte-Carlo points live in the uploaded list
    own.mc_offsets, own.mc_x, own.mc_y = src.mc_offsets, src.mc_x, src.mc_y
    g = me.gidx[idx]
    new_owner = me.owner_of(_wrap(own.centroid_x, me))
    out = {}
    for s_ in range(me.world):
        if s_ != me.rank:
            sel = np.nonzero(new_owner == s_)[0]
            if len(sel):
                out[s_] = (g[sel], extract(own, sel))
    keep = np.nonzero(new_owner == me.rank)[0]
    me._own, me._own_g = extract(own, keep), g[keep]
    return out


def _wrap(cx, me):
    """Centroids of floes that left a periodic domain are owned by the slab of their wrapped image."""
    if not me.period_x:
        return cx
    lo = me.edges[0]
    return lo + np.mod(cx - lo, me.period_x)


def _rebuild_round2(me, all1):
    parts, gl = [me._own], [me._own_g]
    for s_ in range(me.world):
        if s_ != me.rank and me.rank in all1[s_]:
            g, fa = all1[s_][me.rank]
            parts.append(fa)
            gl.append(g)
    own, g = concat(parts), np.concatenate(gl)
    order = np.argsort(g, kind="stable")
    me._own, me._own_g = extract(own, order), g[order]
    out, me._send_sel = {}, {}
    for s_ in range(me.world):
        if s_ == me.rank:
            continue
        sel = np.nonzero(me.needs(me._own.centroid_x, me._own.rmax, r=s_))[0]
        if len(sel):
            me._send_sel[s_] = sel
            out[s_] = (me._own_g[sel], strip_mc(extract(me._own, sel)))
    return out


def _rebuild_finish(me, all2):
    parts, gl, ow = [me._own], [me._own_g], [np.full(me._own.n, me.rank, dtype=np.int64)]
    for s_ in range(me.world):
        if s_ != me.rank and me.rank in all2[s_]:
            g, fa = all2[s_][me.rank]
            parts.append(fa)
            gl.append(g)
            ow.append(np.full(len(g), s_, dtype=np.int64))
    me.build(concat(parts), np.concatenate(gl), np.concatenate(ow))
    me.set_send_lists({s_: me._own_g[sel] for s_, sel in me._send_sel.items()})
    device = me.sbuf[0].device if getattr(me, "sbuf", None) else "cpu"
    me.attach(me.h)
    me.make_buffers(device)
    del me._own, me._own_g, me._send_sel


def rebuild(me):
    """Collective over torch.distributed: call on every rank (e.g. when any rank reports stale())."""
    import torch.distributed as dist
    all1 = [None] * me.world
    dist.all_gather_object(all1, _rebuild_round1(me))
    all2 = [None] * me.world
    dist.all_gather_object(all2, _rebuild_round2(me, all1))
    _rebuild_finish(me, all2)


def rebuild_local(ranks):
    """The same for ranks emulated in one process."""
    all1 = [_rebuild_round1(r) for r in ranks]
    all2 = [_rebuild_round2(r, all1) for r in ranks]
    for r in ranks:
        _rebuild_finish(r, all2)


def exchange_local(ranks):
    """Single-process stand-in for the send/recv (tests: several ranks emulated on one device or
    on the CPU oracle): pack on the owner, hand the buffer over, unpack on the copy holder.  On a CUDA
    device it takes the stream-ordered path of `exchange` (sz_halo_*_on on torch's current stream)."""
    import torch
    cuda = bool(ranks) and bool(ranks[0].sbuf) and ranks[0].sbuf[0].is_cuda
    stream = torch.cuda.current_stream().cuda_stream if cuda else None
    for a in ranks:
        for k, s in enumerate(a.partners):
            nb = a.nbytes[2 * k]
            if not nb:
                continue
            if cuda:
                a.h.halo_pack_on(2 * k, a.sbuf[k].data_ptr(), nb, stream)
            else:
                a.h.halo_pack(2 * k, a.sbuf[k].data_ptr(), nb)
            b = ranks[s]
            kb = b.partners.index(a.rank)
            assert b.nbytes[2 * kb + 1] == nb, (a.rank, s, nb, b.nbytes[2 * kb + 1])
            b.rbuf[kb][:nb].copy_(a.sbuf[k][:nb])
    for b in ranks:
        for k, s in enumerate(b.partners):
            nb = b.nbytes[2 * k + 1]
            if nb:
                if cuda:
                    b.h.halo_unpack_on(2 * k + 1, b.rbuf[k].data_ptr(), nb, stream)
                else:
                    b.h.halo_unpack(2 * k + 1, b.rbuf[k].data_ptr(), nb)


def equal_count_edges(cx, world, period_x=None, x_west=0.0):
    """Slab boundaries with (nearly) equal floe counts.  The outer slabs are unbounded unless the
    domain is periodic in x: then they end at the walls, so that the periodic images of a far floe
    are measured against the true slab."""
    q = np.quantile(cx, np.arange(1, world) / world) if world > 1 else np.zeros(0)
    if period_x:
        return np.concatenate([[x_west], q, [x_west + period_x]])
    return np.concatenate([[-np.inf], q, [np.inf]])


def partition_global(fa, world, period_x, skin, edges=None, period_y=None, x_west=0.0):
    """Every rank of a decomposition built from ONE global list (tests, small fields).  Returns
    the SlabRank objects (not yet attached to handles)."""
    edges = equal_count_edges(fa.centroid_x, world, period_x, x_west) if edges is None else edges
    gidx = np.arange(fa.n, dtype=np.int64)
    rmax_max = float(fa.rmax.max()) if fa.n else 0.0
    probe = SlabRank(0, world, edges, period_x, skin, rmax_max)
    owner = probe.owner_of(fa.centroid_x)
    ranks = [SlabRank(r, world, edges, period_x, skin, rmax_max, period_y).build(fa, gidx, owner) for r in range(world)]
    for r in ranks:
        wanted = {s.rank: s.gidx[s.owner == r.rank] for s in ranks if s.rank != r.rank}
        r.set_send_lists(wanted)
    return ranks


def shift_x(fa, dx):
    """Translate a floe list in x (tiles of a weak-scaling field)."""
    fa.centroid_x = fa.centroid_x + dx
    fa.vert_xy = fa.vert_xy.copy()
    fa.vert_xy[:, 0] += dx
    return fa


def partition_tiles(own, rank, world, tile_L, period_x, skin, period_y=None):
    """Weak-scaling construction: every rank brings its OWN tile `own` (floes with centroids in
    [rank tile_L, (rank + 1) tile_L), global index = rank * n + local index, equal n on all ranks) and
    learns about the neighbours' boundary floes through torch.distributed object collectives (set-up
    time only).  The sender applies the receiver's halo criterion, so no second round is needed."""
    import torch.distributed as dist
    n = own.n
    lo, hi = (0.0, world * tile_L) if period_x else (-np.inf, np.inf)
    edges = np.concatenate([[lo], tile_L * np.arange(1, world), [hi]])
    rm = [None] * world
    dist.all_gather_object(rm, float(own.rmax.max()))
    me = SlabRank(rank, world, edges, period_x, skin, max(rm), period_y)
    gidx_own = rank * n + np.arange(n, dtype=np.int64)
    own.id = gidx_own + 1  # floe ids must be unique over all tiles (collisions.jl:751-758 compares ids)
    out, send_sel = {}, {}
    for s_ in range(world):
        if s_ == rank:
            continue
        sel = np.nonzero(me.needs(own.centroid_x, own.rmax, r=s_))[0]
        if len(sel):
            send_sel[s_] = sel
            out[s_] = (gidx_own[sel], strip_mc(extract(own, sel)))
    allout = [None] * world
    dist.all_gather_object(allout, out)
    parts, gl, ow = [own], [gidx_own], [np.full(n, rank, dtype=np.int64)]
    for s_ in range(world):
        if s_ != rank and rank in allout[s_]:
            g, fa = allout[s_][rank]
            parts.append(fa)
            gl.append(g)
            ow.append(np.full(len(g), s_, dtype=np.int64))
    known = concat(parts)
    me.build(known, np.concatenate(gl), np.concatenate(ow))
    me.set_send_lists({s_: gidx_own[sel] for s_, sel in send_sel.items()})
    return me
