// sz_slab.cpp — slab decomposition inside the library (include/subzero_b200.h, sz_slab_*).
//
// Host logic only: which rank owns which floe, which floes a neighbour needs copies of, the local floe list of every
// rank (sorted by GLOBAL index, so pair orientation, candidate order, row order and the canonical image pair of
// collisions.jl:745-775 are those of the single-list run and owned results are bit-identical to one GPU), migration
// of ownership and renewal of the halo lists.  The reference is single-process; nothing here restates reference code.
//
// Compiled twice from this one source:
//   * into libsubzero_b200.so (prefix sz_): the per-step halo update is the peer-memory push / unpack kernel pair of
//     sz_kernels_fp.cu, reached through sz_slab_backend.h; Monte-Carlo points never leave the device on a rebuild
//     (only those of migrating floes do);
//   * with -DSZ_ORACLE_BUILD into the oracle's test library (prefix szo_): the same lists, the records move through
//     szo_halo_pack / host memory / the alltoallv callback (gloo in the CPU tests).
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include <chrono>
#include <cstdlib>

#include "sz_slab_backend.h"

#define FN(name) SZ_FN(name)

#ifdef SZ_ORACLE_BUILD
extern "C" int32_t szo_upload_partial(sz_handle *h, int32_t do_coupling, const sz_floe_soa *in);  // oracle/szo.c
#endif

namespace {

constexpr int W = 40;  // doubles per floe record
constexpr int C_CX = 0, C_CY = 1, C_RMAX = 5, C_STATUS = 37, C_ID = 38, C_GID = 39;

struct DField {
    size_t off;
    int width, col;
};
#define DF(name, w, c) {offsetof(sz_floe_soa, name), w, c}
const DField DFIELDS[] = {DF(centroid_x, 1, 0), DF(centroid_y, 1, 1), DF(height, 1, 2), DF(area, 1, 3), DF(mass, 1, 4),
                          DF(rmax, 1, 5), DF(moment, 1, 6), DF(alpha, 1, 7), DF(u, 1, 8), DF(v, 1, 9), DF(xi, 1, 10),
                          DF(fxOA, 1, 11), DF(fyOA, 1, 12), DF(trqOA, 1, 13), DF(hflx_factor, 1, 14), DF(overarea, 1, 15),
                          DF(collision_force, 2, 16), DF(collision_trq, 1, 18), DF(stress_accum, 4, 19),
                          DF(stress_instant, 4, 23), DF(strain, 4, 27), DF(p_dxdt, 1, 31), DF(p_dydt, 1, 32),
                          DF(p_dudt, 1, 33), DF(p_dvdt, 1, 34), DF(p_dxidt, 1, 35), DF(p_dalphadt, 1, 36)};
#undef DF
constexpr int NDF = sizeof(DFIELDS) / sizeof(DFIELDS[0]);

inline double *&dptr(sz_floe_soa &s, const DField &f) { return *(double **)((char *)&s + f.off); }
inline const double *cdptr(const sz_floe_soa &s, const DField &f) { return *(double *const *)((const char *)&s + f.off); }
inline double i64_as_double(int64_t v) { double d; memcpy(&d, &v, 8); return d; }
inline int64_t double_as_i64(double d) { int64_t v; memcpy(&v, &d, 8); return v; }

// Full floe records on the host.  Monte-Carlo points of a floe are either held here (msrc = -1, at [mhoff, mhoff + mcnt)
// of mx / my) or resident in the owning rank's device array at offset msrc (CUDA build).
struct FloeList {
    int64_t n = 0;
    std::vector<double> rec;  // [n][W]
    std::vector<int64_t> gidx, vcnt, voff, mcnt, msrc, mhoff;
    std::vector<double> vxy, mx, my;

    double cx(int64_t i) const { return rec[i * W + C_CX]; }
    double rmax(int64_t i) const { return rec[i * W + C_RMAX]; }
};

// A floe of some FloeList that stays alive: lists are merged, filtered and sorted as references and copied only twice
// (device -> records, records -> upload arrays).  mc = false: a halo copy, its Monte-Carlo points are not needed.
struct Ref {
    const FloeList *L;
    int64_t i;
    bool mc;
    int owner;
    int64_t g() const { return L->gidx[i]; }
};
typedef std::vector<Ref> RefList;

void sort_by_gidx(RefList &r) {
    std::stable_sort(r.begin(), r.end(), [](const Ref &a, const Ref &b) { return a.g() < b.g(); });
}

// grow-only host staging of a rank (page-locked in the CUDA build: the copies of a rebuild run at PCIe speed)
struct HostArena {
    char *base = nullptr;
    size_t cap = 0, used = 0;
    bool ensure(size_t bytes) {
        used = 0;
        if (bytes <= cap) return true;
        release();
        const size_t want = bytes + bytes / 4 + 4096;
#ifndef SZ_ORACLE_BUILD
        void *p = nullptr;
        if (szb_host_alloc(want, &p) != SZ_OK) return false;
        base = (char *)p;
#else
        base = (char *)malloc(want);
        if (!base) return false;
#endif
        cap = want;
        return true;
    }
    void release() {
        if (!base) return;
#ifndef SZ_ORACLE_BUILD
        szb_host_free(base);
#else
        free(base);
#endif
        base = nullptr;
        cap = 0;
    }
    template <typename T>
    T *take(size_t count) {
        used = (used + 63) / 64 * 64;
        T *p = (T *)(base + used);
        used += sizeof(T) * std::max<size_t>(count, 1);
        return p;
    }
};

// arrays behind a sz_floe_soa (download target / upload source), carved from a rank's HostArena
struct SoaBuf {
    sz_floe_soa s;
    double *d[32];
    int32_t *status;
    int64_t *id, *gid, *goff, *gindex, *voff, *moff;
    double *vxy, *mx, *my;
    bool alloc(HostArena &A, int64_t n, int64_t V, int64_t M) {
        size_t bytes = 64 * 48;
        for (int f = 0; f < NDF; ++f) bytes += 8 * (size_t)std::max<int64_t>(n * DFIELDS[f].width, 1);
        bytes += (size_t)std::max<int64_t>(n, 1) * (4 + 8 + 8) + ((size_t)n + 2) * 8 * 3 + 8;
        bytes += 8 * (size_t)std::max<int64_t>(2 * V, 2) + 16 * (size_t)std::max<int64_t>(M, 1);
        if (!A.ensure(bytes)) return false;
        memset(&s, 0, sizeof(s));
        s.n = s.n_init = n;
        for (int f = 0; f < NDF; ++f) {
            d[f] = A.take<double>((size_t)(n * DFIELDS[f].width));
            dptr(s, DFIELDS[f]) = d[f];
        }
        status = A.take<int32_t>((size_t)n);
        id = A.take<int64_t>((size_t)n);
        gid = A.take<int64_t>((size_t)n);
        goff = A.take<int64_t>((size_t)n + 1);
        gindex = A.take<int64_t>(1);
        voff = A.take<int64_t>((size_t)n + 1);
        moff = A.take<int64_t>((size_t)n + 1);
        vxy = A.take<double>((size_t)(2 * V));
        mx = A.take<double>((size_t)M);
        my = A.take<double>((size_t)M);
        memset(goff, 0, sizeof(int64_t) * ((size_t)n + 1));
        gindex[0] = 0;
        voff[0] = moff[0] = 0;
        if (n > 0) { status[0] = SZ_STATUS_ACTIVE; }
        s.status_tag = status; s.id = id; s.ghost_id = gid;
        s.ghost_offsets = goff; s.ghost_index = gindex;
        s.vert_offsets = voff; s.vert_xy = vxy;
        s.mc_offsets = moff; s.mc_x = mx; s.mc_y = my;
        return true;
    }
};

// record i of a caller's SoA
void record_from_soa(const sz_floe_soa &s, int64_t i, int64_t g, double *r) {
    for (int f = 0; f < NDF; ++f) {
        const double *p = cdptr(s, DFIELDS[f]);
        for (int w = 0; w < DFIELDS[f].width; ++w) r[DFIELDS[f].col + w] = p ? p[i * DFIELDS[f].width + w] : 0.0;
    }
    r[C_STATUS] = (double)(s.status_tag ? s.status_tag[i] : SZ_STATUS_ACTIVE);
    r[C_ID] = i64_as_double(s.id ? s.id[i] : g + 1);
    r[C_GID] = i64_as_double(s.ghost_id ? s.ghost_id[i] : 0);
}

typedef std::vector<char> Blob;
template <typename T>
void put(Blob &b, const T *p, size_t count) {
    const char *c = (const char *)p;
    b.insert(b.end(), c, c + sizeof(T) * count);
}
template <typename T>
void get(const Blob &b, size_t &pos, T *p, size_t count) {
    memcpy(p, b.data() + pos, sizeof(T) * count);
    pos += sizeof(T) * count;
}

struct Rank {
    int rank = 0, device = 0;
    sz_handle *h = nullptr;
    std::vector<int64_t> gidx;
    std::vector<int32_t> owner;
    int64_t n_owned = 0;
    std::vector<int32_t> partners;
    std::vector<int64_t> send_off, send_idx, recv_off, recv_idx;
    int64_t send_bytes = 0;
    bool built = false;
    HostArena stage;
#ifdef SZ_ORACLE_BUILD
    std::vector<double> refx, refy;
    std::vector<int64_t> sbytes, rbytes;  // per partner
    double disp = 0.0;
#endif
};

}  // namespace

struct sz_slab {
    sz_config cfg;
    int world = 1, rank_first = 0, n_local = 1;
    double skin = 0.0;
    sz_alltoallv_fn cb = nullptr;
    void *ctx = nullptr;
    std::vector<double> edges;
    bool have_edges = false, have_domain = false;
    int32_t kinds[4] = {0, 0, 0, 0};
    double vals[4] = {0, 0, 0, 0};
    double period_x = 0.0, period_y = 0.0, x_west = 0.0;
    double rmax_max = 0.0;
    std::vector<Rank> ranks;
    int64_t rebuilds = 0;
    char err[512] = {0};
};

namespace {

// SZ_SLAB_DEBUG=1: wall-clock of the phases of a (re)build on stderr
struct PhaseTimer {
    bool on;
    std::chrono::steady_clock::time_point t0;
    PhaseTimer() : on(getenv("SZ_SLAB_DEBUG") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void lap(int rank, const char *what) {
        if (!on) return;
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[slab rank %d] %-28s %8.2f ms\n", rank, what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

int32_t sfail(sz_slab *s, int32_t code, const char *msg) {
    if (s) snprintf(s->err, sizeof(s->err), "%s", msg);
    return code;
}
int32_t hfail(sz_slab *s, Rank &r, int32_t code, const char *what) {
    const char *m = FN(last_error)(r.h);
    snprintf(s->err, sizeof(s->err), "rank %d: %s: %s", r.rank, what, m ? m : "");
    return code;
}
#define HCK(call, what)                                   \
    do {                                                  \
        int32_t _rc = (call);                             \
        if (_rc != SZ_OK) return hfail(S, R, _rc, what); \
    } while (0)

// out[k][d]: from local rank k to global rank d  ->  in[k][s]: for local rank k from global rank s
int32_t exchange(sz_slab *S, std::vector<std::vector<Blob>> &out, std::vector<std::vector<Blob>> &in) {
    const int Wd = S->world;
    in.assign(S->n_local, std::vector<Blob>(Wd));
    if (S->n_local == Wd) {
        for (int k = 0; k < Wd; ++k)
            for (int d = 0; d < Wd; ++d) in[d][k] = std::move(out[k][d]);
        return SZ_OK;
    }
    // one rank per process: sizes first, then the payload (MPI_Alltoallv semantics)
    std::vector<int64_t> ssz(Wd), rsz(Wd, 0), o8(Wd + 1), so(Wd + 1, 0), ro(Wd + 1, 0);
    for (int d = 0; d <= Wd; ++d) o8[d] = 8 * (int64_t)d;
    for (int d = 0; d < Wd; ++d) ssz[d] = (int64_t)out[0][d].size();
    if (S->cb(S->ctx, ssz.data(), o8.data(), rsz.data(), o8.data()) != 0) return sfail(S, SZ_ERR_INVALID, "slab: alltoallv callback failed (sizes)");
    for (int d = 0; d < Wd; ++d) {
        so[d + 1] = so[d] + ssz[d];
        ro[d + 1] = ro[d] + rsz[d];
    }
    Blob sb((size_t)std::max<int64_t>(so[Wd], 1)), rb((size_t)std::max<int64_t>(ro[Wd], 1));
    for (int d = 0; d < Wd; ++d)
        if (ssz[d]) memcpy(sb.data() + so[d], out[0][d].data(), (size_t)ssz[d]);
    if (S->cb(S->ctx, sb.data(), so.data(), rb.data(), ro.data()) != 0) return sfail(S, SZ_ERR_INVALID, "slab: alltoallv callback failed (payload)");
    for (int s = 0; s < Wd; ++s) in[0][s].assign(rb.begin() + ro[s], rb.begin() + ro[s + 1]);
    return SZ_OK;
}

int owner_of(const sz_slab *S, double cx) {
    if (S->period_x > 0.0) {  // a floe that left a periodic domain belongs to the slab of its wrapped image
        double t = fmod(cx - S->x_west, S->period_x);
        if (t < 0) t += S->period_x;
        cx = S->x_west + t;
    }
    // number of edges <= cx, minus one (searchsorted side = right)
    int r = (int)(std::upper_bound(S->edges.begin(), S->edges.end(), cx) - S->edges.begin()) - 1;
    return std::min(std::max(r, 0), S->world - 1);
}

// does rank r need a copy of a floe at cx with bounding radius rmax?  |x - slab_r| < rmax + rmax_max + skin, minimised
// over the east/west period: the reference's periodic ghosts (collisions.jl:925-952) then appear on the rank that
// needs them simply because add_ghosts! runs on its local list
bool needs(const sz_slab *S, int r, double cx, double rmax) {
    const double xa = S->edges[r], xb = S->edges[r + 1];
    auto dist = [&](double x) { return std::max(std::max(xa - x, x - xb), 0.0); };
    double d = dist(cx);
    if (S->period_x > 0.0) d = std::min(d, std::min(dist(cx + S->period_x), dist(cx - S->period_x)));
    return d < rmax + S->rmax_max + S->skin;
}

// Monte-Carlo points of floe i of a list owned by rank R, wherever they are
int32_t fetch_mc(sz_slab *S, Rank &R, const FloeList &L, int64_t i, double *x, double *y) {
    if (L.mcnt[i] == 0) return SZ_OK;
    if (L.msrc[i] < 0) {
        memcpy(x, L.mx.data() + L.mhoff[i], sizeof(double) * L.mcnt[i]);
        memcpy(y, L.my.data() + L.mhoff[i], sizeof(double) * L.mcnt[i]);
        return SZ_OK;
    }
#ifndef SZ_ORACLE_BUILD
    HCK(szb_fetch_mc(R.h, L.msrc[i], L.mcnt[i], x, y), "fetch Monte-Carlo points");
    return SZ_OK;
#else
    (void)S; (void)R;
    return SZ_ERR_INVALID;
#endif
}

int32_t serialize(sz_slab *S, Rank &R, const RefList &refs, const std::vector<int64_t> &sel, bool with_mc, Blob &b) {
    int64_t cnt = (int64_t)sel.size();
    size_t bytes = 8;
    for (int64_t q : sel) {
        const FloeList &L = *refs[q].L;
        const int64_t i = refs[q].i;
        bytes += 24 + 8 * W + 16 * (size_t)L.vcnt[i] + (with_mc ? 16 * (size_t)L.mcnt[i] : 0);
    }
    b.reserve(b.size() + bytes);
    put(b, &cnt, 1);
    std::vector<double> tx, ty;
    for (int64_t q : sel) {
        const FloeList &L = *refs[q].L;
        const int64_t i = refs[q].i;
        int64_t hdr[3] = {L.gidx[i], L.vcnt[i], with_mc ? L.mcnt[i] : 0};
        put(b, hdr, 3);
        put(b, L.rec.data() + i * W, W);
        put(b, L.vxy.data() + 2 * L.voff[i], (size_t)(2 * L.vcnt[i]));
        if (with_mc && L.mcnt[i] > 0) {
            tx.resize((size_t)L.mcnt[i]); ty.resize((size_t)L.mcnt[i]);
            int32_t rc = fetch_mc(S, R, L, i, tx.data(), ty.data());
            if (rc) return rc;
            put(b, tx.data(), tx.size());
            put(b, ty.data(), ty.size());
        }
    }
    return SZ_OK;
}

void deserialize(const Blob &b, FloeList &L) {
    if (b.size() < 8) return;
    size_t pos = 0;
    int64_t cnt = 0;
    get(b, pos, &cnt, 1);
    for (int64_t k = 0; k < cnt; ++k) {
        int64_t hdr[3];
        get(b, pos, hdr, 3);
        L.gidx.push_back(hdr[0]);
        L.vcnt.push_back(hdr[1]);
        L.mcnt.push_back(hdr[2]);
        L.msrc.push_back(-1);
        L.rec.resize(L.rec.size() + W);
        get(b, pos, L.rec.data() + L.n * W, W);
        L.voff.push_back((int64_t)L.vxy.size() / 2);
        L.vxy.resize(L.vxy.size() + 2 * hdr[1]);
        get(b, pos, L.vxy.data() + 2 * L.voff.back(), (size_t)(2 * hdr[1]));
        L.mhoff.push_back((int64_t)L.mx.size());
        if (hdr[2] > 0) {
            L.mx.resize(L.mx.size() + hdr[2]); L.my.resize(L.my.size() + hdr[2]);
            get(b, pos, L.mx.data() + L.mhoff.back(), (size_t)hdr[2]);
            get(b, pos, L.my.data() + L.mhoff.back(), (size_t)hdr[2]);
        }
        L.n++;
    }
}

// every floe rank R holds, as full records (owned ones only); Monte-Carlo points stay resident in the CUDA build
int32_t download_owned(sz_slab *S, Rank &R, FloeList &L) {
    sz_counts c;
    HCK(FN(get_counts)(R.h, &c), "get_counts");
    const int64_t n = c.n_total;
    SoaBuf B;
#ifdef SZ_ORACLE_BUILD
    if (!B.alloc(R.stage, n, c.n_vertices, c.n_mc)) return sfail(S, SZ_ERR_NOMEM, "slab: host staging");
#else
    if (!B.alloc(R.stage, n, c.n_vertices, 0)) return sfail(S, SZ_ERR_NOMEM, "slab: host staging");
    B.s.mc_x = B.s.mc_y = nullptr;
#endif
    HCK(FN(download_floes)(R.h, &B.s), "download_floes");
#ifndef SZ_ORACLE_BUILD
    HCK(szb_mc_offsets(R.h, B.moff), "mc offsets");
#endif
    int64_t no = 0, Vo = 0, Mo = 0;
    for (int64_t i = 0; i < n; ++i)
        if (R.owner[i] == R.rank) { no++; Vo += B.voff[i + 1] - B.voff[i]; Mo += B.moff[i + 1] - B.moff[i]; }
    L.n = no;
    L.rec.resize((size_t)no * W);
    L.gidx.resize(no); L.vcnt.resize(no); L.voff.resize(no); L.mcnt.resize(no); L.msrc.resize(no); L.mhoff.resize(no);
    L.vxy.resize((size_t)(2 * Vo));
#ifdef SZ_ORACLE_BUILD
    L.mx.resize((size_t)Mo); L.my.resize((size_t)Mo);
#else
    (void)Mo;
#endif
    int64_t q = 0, vo = 0, mo = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (R.owner[i] != R.rank) continue;
        record_from_soa(B.s, i, R.gidx[i], L.rec.data() + q * W);
        L.gidx[q] = R.gidx[i];
        L.vcnt[q] = B.voff[i + 1] - B.voff[i];
        L.voff[q] = vo;
        memcpy(L.vxy.data() + 2 * vo, B.vxy + 2 * B.voff[i], sizeof(double) * 2 * L.vcnt[q]);
        vo += L.vcnt[q];
        L.mcnt[q] = B.moff[i + 1] - B.moff[i];
#ifdef SZ_ORACLE_BUILD
        L.msrc[q] = -1;
        L.mhoff[q] = mo;
        if (L.mcnt[q] > 0) {
            memcpy(L.mx.data() + mo, B.mx + B.moff[i], sizeof(double) * L.mcnt[q]);
            memcpy(L.my.data() + mo, B.my + B.moff[i], sizeof(double) * L.mcnt[q]);
        }
        mo += L.mcnt[q];
#else
        L.msrc[q] = B.moff[i];  // resident: offset in the rank's device array
        L.mhoff[q] = 0;
#endif
        q++;
    }
    (void)mo;
    return SZ_OK;
}

// upload the local list of rank R (owned + halo, ascending global index)
int32_t upload_local(sz_slab *S, Rank &R, const RefList &refs) {
    const int64_t n = (int64_t)refs.size();
    int64_t V = 0, M = 0, n_extra = 0;
    bool any_resident = false;
    for (const Ref &r : refs) {
        const FloeList &L = *r.L;
        V += L.vcnt[r.i];
        if (!r.mc) continue;
        M += L.mcnt[r.i];
        if (L.mcnt[r.i] > 0 && L.msrc[r.i] >= 0) any_resident = true;
        else n_extra += L.mcnt[r.i];
    }
    SoaBuf B;
    if (!B.alloc(R.stage, n, V, any_resident ? n_extra : M)) return sfail(S, SZ_ERR_NOMEM, "slab: host staging");
    std::vector<int64_t> mc_src((size_t)std::max<int64_t>(n, 1), 0);
    int64_t vo = 0, mo = 0, eo = 0;
    for (int64_t i = 0; i < n; ++i) {
        const FloeList &L = *refs[i].L;
        const int64_t j = refs[i].i;
        const double *r = L.rec.data() + j * W;
        for (int f = 0; f < NDF; ++f)
            for (int w = 0; w < DFIELDS[f].width; ++w) B.d[f][(size_t)(i * DFIELDS[f].width + w)] = r[DFIELDS[f].col + w];
        B.status[i] = (int32_t)r[C_STATUS];
        B.id[i] = double_as_i64(r[C_ID]);
        B.gid[i] = double_as_i64(r[C_GID]);
        B.voff[i] = vo;
        memcpy(B.vxy + 2 * vo, L.vxy.data() + 2 * L.voff[j], sizeof(double) * 2 * L.vcnt[j]);
        vo += L.vcnt[j];
        B.moff[i] = mo;
        const int64_t mc = refs[i].mc ? L.mcnt[j] : 0;
        if (mc > 0) {
            if (L.msrc[j] >= 0) {
                mc_src[i] = L.msrc[j];
            } else {
                const int64_t at = any_resident ? eo : mo;
                memcpy(B.mx + at, L.mx.data() + L.mhoff[j], sizeof(double) * mc);
                memcpy(B.my + at, L.my.data() + L.mhoff[j], sizeof(double) * mc);
                mc_src[i] = -1 - eo;
                eo += mc;
            }
        }
        mo += mc;
    }
    B.voff[n] = vo;
    B.moff[n] = mo;
#ifndef SZ_ORACLE_BUILD
    if (any_resident) {
        HCK(szb_upload_floes_resident_mc(R.h, &B.s, mc_src.data(), n_extra), "upload_floes (resident Monte-Carlo points)");
        return SZ_OK;
    }
#endif
    HCK(FN(upload_floes)(R.h, &B.s), "upload_floes");
    return SZ_OK;
}

// Collective.  own[k]: the floes local rank k currently holds as owner (any distribution).  Migrates every floe to
// the slab of its centroid, builds the halo lists, uploads the local lists and wires the per-step exchange.
int32_t repartition(sz_slab *S, std::vector<FloeList> &own) {
    const int Wd = S->world, NL = S->n_local;
    if (!S->have_domain) return sfail(S, SZ_ERR_INVALID, "slab: build before sz_slab_set_domain");
    PhaseTimer pt;
    const int r0 = S->ranks[0].rank;
    // ---- global bounding radius (and equal-count edges when none were given) ----------------------------------------
    {
        std::vector<std::vector<Blob>> out(NL, std::vector<Blob>(Wd)), in;
        for (int k = 0; k < NL; ++k) {
            double m = 0.0;
            for (int64_t i = 0; i < own[k].n; ++i) m = std::max(m, own[k].rmax(i));
            for (int d = 0; d < Wd; ++d) put(out[k][d], &m, 1);
        }
        int32_t rc = exchange(S, out, in);
        if (rc) return rc;
        S->rmax_max = 0.0;
        for (int s = 0; s < Wd; ++s) {
            double m = 0.0;
            size_t pos = 0;
            if (in[0][s].size() >= 8) get(in[0][s], pos, &m, 1);
            S->rmax_max = std::max(S->rmax_max, m);
        }
    }
    if (!S->have_edges) {
        if (NL != Wd) return sfail(S, SZ_ERR_INVALID, "slab: one rank per process needs sz_slab_set_edges");
        std::vector<double> cx;
        for (int k = 0; k < NL; ++k)
            for (int64_t i = 0; i < own[k].n; ++i) cx.push_back(own[k].cx(i));
        std::sort(cx.begin(), cx.end());
        S->edges.assign(Wd + 1, 0.0);
        const double inf = std::numeric_limits<double>::infinity();
        S->edges[0] = S->period_x > 0.0 ? S->x_west : -inf;
        S->edges[Wd] = S->period_x > 0.0 ? S->x_west + S->period_x : inf;
        for (int r = 1; r < Wd; ++r) {  // linear-interpolated quantile r / world
            if (cx.empty()) { S->edges[r] = 0.0; continue; }
            double pos = (double)(cx.size() - 1) * r / Wd;
            size_t lo = (size_t)floor(pos), hi = std::min(lo + 1, cx.size() - 1);
            S->edges[r] = cx[lo] + (pos - (double)lo) * (cx[hi] - cx[lo]);
        }
        S->have_edges = true;
    }
    pt.lap(r0, "rmax / edges");
    // ---- 1. migration of ownership (full records incl. Monte-Carlo points) -----------------------------------------------
    std::vector<RefList> mine(NL);
    std::vector<FloeList> got(NL);
    {
        std::vector<std::vector<Blob>> out(NL, std::vector<Blob>(Wd)), in;
        for (int k = 0; k < NL; ++k) {
            Rank &R = S->ranks[k];
            RefList all((size_t)own[k].n);
            std::vector<std::vector<int64_t>> sel(Wd);
            for (int64_t i = 0; i < own[k].n; ++i) {
                all[i] = {&own[k], i, true, 0};
                const int d = owner_of(S, own[k].cx(i));
                if (d == R.rank) mine[k].push_back({&own[k], i, true, R.rank});
                else sel[d].push_back(i);
            }
            for (int d = 0; d < Wd; ++d)
                if (d != R.rank && !sel[d].empty()) {
                    int32_t rc = serialize(S, R, all, sel[d], true, out[k][d]);
                    if (rc) return rc;
                }
        }
        int32_t rc = exchange(S, out, in);
        if (rc) return rc;
        for (int k = 0; k < NL; ++k) {
            for (int s = 0; s < Wd; ++s) deserialize(in[k][s], got[k]);
            for (int64_t i = 0; i < got[k].n; ++i) mine[k].push_back({&got[k], i, true, S->ranks[k].rank});
            sort_by_gidx(mine[k]);
        }
    }
    pt.lap(r0, "migration");
    // ---- 2. halo copies: the OWNER applies the receiver's criterion (no second round needed) --------------------------------
    std::vector<std::vector<std::vector<int64_t>>> send_sel(NL, std::vector<std::vector<int64_t>>(Wd));
    std::vector<RefList> local(NL);
    std::vector<std::vector<FloeList>> halo(NL, std::vector<FloeList>(Wd));
    {
        std::vector<std::vector<Blob>> out(NL, std::vector<Blob>(Wd)), in;
        for (int k = 0; k < NL; ++k) {
            Rank &R = S->ranks[k];
            for (int d = 0; d < Wd; ++d) {
                if (d == R.rank) continue;
                for (int64_t q = 0; q < (int64_t)mine[k].size(); ++q) {
                    const Ref &r = mine[k][q];
                    if (needs(S, d, r.L->cx(r.i), r.L->rmax(r.i))) send_sel[k][d].push_back(q);
                }
                if (!send_sel[k][d].empty()) {
                    int32_t rc = serialize(S, R, mine[k], send_sel[k][d], false, out[k][d]);
                    if (rc) return rc;
                }
            }
        }
        int32_t rc = exchange(S, out, in);
        if (rc) return rc;
        for (int k = 0; k < NL; ++k) {
            local[k] = mine[k];
            for (int s = 0; s < Wd; ++s) {
                deserialize(in[k][s], halo[k][s]);
                for (int64_t i = 0; i < halo[k][s].n; ++i) local[k].push_back({&halo[k][s], i, false, s});
            }
            sort_by_gidx(local[k]);
        }
    }
    pt.lap(r0, "halo selection + exchange");
    // ---- 3. local lists, exchange lists, upload, wiring ----------------------------------------------------------------------------
    std::vector<std::vector<SlabWire>> wires(NL);
    for (int k = 0; k < NL; ++k) {
        Rank &R = S->ranks[k];
        const RefList &L = local[k];
        const int64_t Ln = (int64_t)L.size();
        R.gidx.resize((size_t)Ln);
        R.owner.resize((size_t)Ln);
        for (int64_t i = 0; i < Ln; ++i) {
            R.gidx[i] = L[i].g();
            R.owner[i] = L[i].owner;
            if (i > 0 && R.gidx[i] == R.gidx[i - 1]) return sfail(S, SZ_ERR_INVALID, "slab: duplicate global floe index");
        }
        R.n_owned = (int64_t)mine[k].size();
        // position of every owned floe in the local list (both ascending in the global index)
        std::vector<int64_t> pos_owned((size_t)mine[k].size());
        std::vector<std::vector<int64_t>> recv_by(Wd);
        {
            int64_t q = 0;
            for (int64_t i = 0; i < Ln; ++i) {
                if (R.owner[i] == R.rank) pos_owned[q++] = i;
                else recv_by[R.owner[i]].push_back(i);
            }
        }
        R.partners.clear();
        for (int d = 0; d < Wd; ++d)
            if (d != R.rank && (!send_sel[k][d].empty() || !recv_by[d].empty())) R.partners.push_back(d);
        const int np = (int)R.partners.size();
        if (np > SZ_SLAB_MAX_PARTNERS) return sfail(S, SZ_ERR_UNSUPPORTED, "slab: a rank has more than 16 exchange partners (slabs thinner than the interaction range)");
        R.send_off.assign(np + 1, 0); R.recv_off.assign(np + 1, 0);
        R.send_idx.clear(); R.recv_idx.clear();
        for (int p = 0; p < np; ++p) {
            const int d = R.partners[p];
            for (int64_t i : send_sel[k][d]) R.send_idx.push_back(pos_owned[(size_t)i]);
            R.recv_idx.insert(R.recv_idx.end(), recv_by[d].begin(), recv_by[d].end());
            R.send_off[p + 1] = (int64_t)R.send_idx.size();
            R.recv_off[p + 1] = (int64_t)R.recv_idx.size();
        }
        pt.lap(R.rank, "exchange lists");
        int32_t rc = upload_local(S, R, L);
        if (rc) return rc;
        pt.lap(R.rank, "upload local list");
        std::vector<uint8_t> owned((size_t)std::max<int64_t>(Ln, 1), 0);
        for (int64_t i = 0; i < Ln; ++i) owned[i] = R.owner[i] == R.rank;
        R.send_bytes = 0;
        for (int p = 0; p < np; ++p)
            for (int64_t q = R.send_off[p]; q < R.send_off[p + 1]; ++q) {
                const Ref &r = L[(size_t)R.send_idx[q]];
                R.send_bytes += 64 + 16 * r.L->vcnt[r.i];
            }
#ifndef SZ_ORACLE_BUILD
        SlabLists ls;
        ls.rank = R.rank; ls.n_partners = np; ls.partner_rank = R.partners.data();
        ls.send_off = R.send_off.data(); ls.send_idx = R.send_idx.data(); ls.recv_off = R.recv_off.data(); ls.recv_idx = R.recv_idx.data();
        ls.owned = owned.data(); ls.period_x = S->period_x; ls.period_y = S->period_y;
        wires[k].resize((size_t)std::max(np, 1));
        HCK(szb_configure(R.h, &ls, wires[k].data()), "configure halo lists");
#else
        // oracle build: the records move through szo_halo_pack / szo_halo_unpack; list 2p = send to partner p, 2p+1 = receive
        std::vector<int64_t> off(2 * np + 1, 0), idx;
        for (int p = 0; p < np; ++p) {
            for (int64_t q = R.send_off[p]; q < R.send_off[p + 1]; ++q) idx.push_back(R.send_idx[q] + 1);
            off[2 * p + 1] = (int64_t)idx.size();
            for (int64_t q = R.recv_off[p]; q < R.recv_off[p + 1]; ++q) idx.push_back(R.recv_idx[q] + 1);
            off[2 * p + 2] = (int64_t)idx.size();
        }
        if (idx.empty()) idx.push_back(1);
        HCK(FN(halo_configure)(R.h, 2 * np, off.data(), idx.data()), "halo_configure");
        R.sbytes.assign(np, 0); R.rbytes.assign(np, 0);
        for (int p = 0; p < np; ++p) {
            HCK(FN(halo_bytes)(R.h, 2 * p, &R.sbytes[p]), "halo_bytes");
            HCK(FN(halo_bytes)(R.h, 2 * p + 1, &R.rbytes[p]), "halo_bytes");
        }
        R.refx.resize((size_t)Ln); R.refy.resize((size_t)Ln);
        for (int64_t i = 0; i < Ln; ++i) {
            R.refx[i] = L[i].L->rec[L[i].i * W + C_CX];
            R.refy[i] = L[i].L->rec[L[i].i * W + C_CY];
        }
        R.disp = 0.0;
#endif
        R.built = true;
    }
    pt.lap(r0, "configure");
#ifndef SZ_ORACLE_BUILD
    {   // tell every partner where to write (wire p of rank k is for partner p), then map and publish epoch 1
        std::vector<std::vector<Blob>> out(NL, std::vector<Blob>(Wd)), in;
        for (int k = 0; k < NL; ++k) {
            Rank &R = S->ranks[k];
            for (size_t p = 0; p < R.partners.size(); ++p) put(out[k][R.partners[p]], &wires[k][p], 1);
        }
        int32_t rc = exchange(S, out, in);
        if (rc) return rc;
        for (int k = 0; k < NL; ++k) {
            Rank &R = S->ranks[k];
            std::vector<SlabWire> peer((size_t)std::max<size_t>(R.partners.size(), 1));
            for (size_t p = 0; p < R.partners.size(); ++p) {
                const Blob &b = in[k][R.partners[p]];
                if (b.size() != sizeof(SlabWire)) return sfail(S, SZ_ERR_INVALID, "slab: partner lists are not symmetric");
                memcpy(&peer[p], b.data(), sizeof(SlabWire));
            }
            HCK(szb_connect(R.h, peer.data()), "connect partners");
        }
    }
#endif
    pt.lap(r0, "wiring + first publication");
    return SZ_OK;
}

#ifdef SZ_ORACLE_BUILD
// per-step exchange of the oracle build: pack -> host memory / alltoallv -> unpack
int32_t host_exchange(sz_slab *S) {
    const int Wd = S->world, NL = S->n_local;
    std::vector<std::vector<Blob>> out(NL, std::vector<Blob>(Wd)), in;
    for (int k = 0; k < NL; ++k) {
        Rank &R = S->ranks[k];
        for (size_t p = 0; p < R.partners.size(); ++p) {
            Blob &b = out[k][R.partners[p]];
            b.resize((size_t)R.sbytes[p]);
            if (R.sbytes[p]) HCK(FN(halo_pack)(R.h, 2 * (int)p, b.data(), R.sbytes[p]), "halo_pack");
        }
    }
    int32_t rc = exchange(S, out, in);
    if (rc) return rc;
    for (int k = 0; k < NL; ++k) {
        Rank &R = S->ranks[k];
        for (size_t p = 0; p < R.partners.size(); ++p) {
            const Blob &b = in[k][R.partners[p]];
            if ((int64_t)b.size() != R.rbytes[p]) return sfail(S, SZ_ERR_INVALID, "slab: halo message size differs from the list");
            if (R.rbytes[p]) HCK(FN(halo_unpack)(R.h, 2 * (int)p + 1, b.data(), R.rbytes[p]), "halo_unpack");
        }
    }
    return SZ_OK;
}

int32_t host_displacement(sz_slab *S, Rank &R) {
    const int64_t n = (int64_t)R.gidx.size();
    if (n == 0) return SZ_OK;
    std::vector<double> cx((size_t)n), cy((size_t)n);
    sz_floe_soa q;
    memset(&q, 0, sizeof(q));
    q.centroid_x = cx.data(); q.centroid_y = cy.data();
    HCK(FN(download_floes)(R.h, &q), "download centroids");
    for (int64_t i = 0; i < n; ++i) {
        if (R.owner[i] != R.rank) continue;
        double dx = fabs(cx[i] - R.refx[i]), dy = fabs(cy[i] - R.refy[i]);
        if (S->period_x > 0.0) dx = std::min(dx, fabs(dx - S->period_x));
        if (S->period_y > 0.0) dy = std::min(dy, fabs(dy - S->period_y));
        R.disp = std::max(R.disp, sqrt(dx * dx + dy * dy));
    }
    return SZ_OK;
}
#endif

double local_max_disp(sz_slab *S) {
    double m = 0.0;
    for (Rank &R : S->ranks) {
#ifdef SZ_ORACLE_BUILD
        m = std::max(m, R.disp);
#else
        m = std::max(m, szb_max_displacement(R.h));
#endif
    }
    return m;
}

int32_t do_rebuild(sz_slab *S) {
    std::vector<FloeList> own(S->n_local);
    PhaseTimer pt;
#ifndef SZ_ORACLE_BUILD
    for (Rank &R : S->ranks) HCK(szb_release_peers(R.h, 0), "release partners");
#endif
    for (int k = 0; k < S->n_local; ++k) {
        int32_t rc = download_owned(S, S->ranks[k], own[k]);
        if (rc) return rc;
    }
    pt.lap(S->ranks[0].rank, "download owned floes");
    int32_t rc = repartition(S, own);
    if (rc == SZ_OK) S->rebuilds++;
    return rc;
}

}  // namespace

// ---- C ABI --------------------------------------------------------------------------------------------------------------
extern "C" {

int32_t FN(slab_create)(const sz_config *cfg, int32_t world, int32_t rank_first, int32_t n_local, const int32_t *devices,
                        double skin, sz_alltoallv_fn alltoallv, void *ctx, sz_slab **out) {
    if (!cfg || !out || world < 1 || n_local < 1 || rank_first < 0 || rank_first + n_local > world || !(skin >= 0.0)) return SZ_ERR_INVALID;
    *out = nullptr;
    if (n_local != world && (n_local != 1 || !alltoallv)) return SZ_ERR_INVALID;  // all ranks here, or one rank per process + a transport
    sz_slab *S = new (std::nothrow) sz_slab();
    if (!S) return SZ_ERR_NOMEM;
    S->cfg = *cfg;
    S->world = world; S->rank_first = rank_first; S->n_local = n_local; S->skin = skin;
    S->cb = alltoallv; S->ctx = ctx;
    S->ranks.resize(n_local);
    for (int k = 0; k < n_local; ++k) {
        Rank &R = S->ranks[k];
        R.rank = rank_first + k;
        R.device = devices ? devices[k] : 0;
        sz_config c = *cfg;
        c.device = R.device;
        int32_t rc = FN(create)(&c, &R.h);
        if (rc != SZ_OK) {
            for (int j = 0; j < k; ++j) FN(destroy)(S->ranks[j].h);
            delete S;
            return rc;
        }
    }
    *out = S;
    return SZ_OK;
}

void FN(slab_destroy)(sz_slab *S) {
    if (!S) return;
#ifndef SZ_ORACLE_BUILD
    for (Rank &R : S->ranks) if (R.h) szb_release_peers(R.h, 1);
#endif
    for (Rank &R : S->ranks) {
        R.stage.release();
        if (R.h) FN(destroy)(R.h);
    }
    delete S;
}

const char *FN(slab_last_error)(sz_slab *S) { return S ? S->err : "null slab"; }

int32_t FN(slab_handle)(sz_slab *S, int32_t k, sz_handle **out) {
    if (!S || !out || k < 0 || k >= S->n_local) return SZ_ERR_INVALID;
    *out = S->ranks[k].h;
    return SZ_OK;
}

int32_t FN(slab_set_grid)(sz_slab *S, int32_t Nx, int32_t Ny, double x0, double xf, double y0, double yf) {
    if (!S) return SZ_ERR_INVALID;
    for (Rank &R : S->ranks) HCK(FN(set_grid)(R.h, Nx, Ny, x0, xf, y0, yf), "set_grid");
    return SZ_OK;
}

int32_t FN(slab_set_fields)(sz_slab *S, const double *ou, const double *ov, const double *oh, const double *au, const double *av) {
    if (!S) return SZ_ERR_INVALID;
    for (Rank &R : S->ranks) HCK(FN(set_fields)(R.h, ou, ov, oh, au, av), "set_fields");
    return SZ_OK;
}

int32_t FN(slab_set_domain)(sz_slab *S, const int32_t kinds[4], const double vals[4], const double uv[8], const double rect[16],
                            int32_t n_topo, const int64_t *toff, const double *txy, const double *tcent, const double *trmax) {
    if (!S || !kinds || !vals) return SZ_ERR_INVALID;
    for (Rank &R : S->ranks) HCK(FN(set_domain)(R.h, kinds, vals, uv, rect, n_topo, toff, txy, tcent, trmax), "set_domain");
    for (int w = 0; w < 4; ++w) { S->kinds[w] = kinds[w]; S->vals[w] = vals[w]; }
    S->period_x = kinds[2] == SZ_BOUNDARY_PERIODIC ? vals[2] - vals[3] : 0.0;
    S->period_y = kinds[0] == SZ_BOUNDARY_PERIODIC ? vals[0] - vals[1] : 0.0;
    S->x_west = vals[3];
    S->have_domain = true;
    return SZ_OK;
}

int32_t FN(slab_set_edges)(sz_slab *S, const double *edges) {
    if (!S || !edges) return SZ_ERR_INVALID;
    for (int r = 0; r < S->world; ++r)
        if (!(edges[r] <= edges[r + 1])) return sfail(S, SZ_ERR_INVALID, "slab: edges must ascend");
    S->edges.assign(edges, edges + S->world + 1);
    S->have_edges = true;
    return SZ_OK;
}

int32_t FN(slab_build)(sz_slab *S, const sz_floe_soa *const *floes, const int64_t *const *gidx) {
    if (!S || !floes || !gidx) return SZ_ERR_INVALID;
    std::vector<FloeList> own(S->n_local);
    for (int k = 0; k < S->n_local; ++k) {
        const sz_floe_soa *s = floes[k];
        if (!s || s->n == 0) continue;
        if (s->n != s->n_init) return sfail(S, SZ_ERR_INVALID, "slab_build: floe lists must not contain ghosts");
        if (!gidx[k] || !s->centroid_x || !s->centroid_y || !s->rmax || !s->vert_offsets || !s->vert_xy)
            return sfail(S, SZ_ERR_INVALID, "slab_build: global indices and geometry arrays are required");
        FloeList &L = own[k];
        L.n = s->n;
        L.rec.assign((size_t)s->n * W, 0.0);
        L.gidx.assign(gidx[k], gidx[k] + s->n);
        L.vcnt.resize(s->n); L.voff.resize(s->n); L.mcnt.resize(s->n); L.msrc.assign(s->n, -1); L.mhoff.resize(s->n);
        for (int64_t i = 0; i < s->n; ++i) {
            record_from_soa(*s, i, gidx[k][i], L.rec.data() + i * W);
            L.voff[i] = s->vert_offsets[i];
            L.vcnt[i] = s->vert_offsets[i + 1] - s->vert_offsets[i];
            L.mhoff[i] = s->mc_offsets ? s->mc_offsets[i] : 0;
            L.mcnt[i] = s->mc_offsets ? s->mc_offsets[i + 1] - s->mc_offsets[i] : 0;
        }
        L.vxy.assign(s->vert_xy, s->vert_xy + 2 * s->vert_offsets[s->n]);
        if (s->mc_offsets && s->mc_offsets[s->n] > 0) {
            if (!s->mc_x || !s->mc_y) return sfail(S, SZ_ERR_INVALID, "slab_build: Monte-Carlo offsets without points");
            L.mx.assign(s->mc_x, s->mc_x + s->mc_offsets[s->n]);
            L.my.assign(s->mc_y, s->mc_y + s->mc_offsets[s->n]);
        }
    }
#ifndef SZ_ORACLE_BUILD
    for (Rank &R : S->ranks) if (R.built) HCK(szb_release_peers(R.h, 0), "release partners");
#endif
    int32_t rc = repartition(S, own);
#ifndef SZ_ORACLE_BUILD
    // A rebuild right away (nothing moved: same lists) makes every first-time allocation of the rebuild path — the second
    // Monte-Carlo array of the on-device regather, the page-locked staging, the gather scratch — part of the set-up
    // instead of the first rebuild of the run (0.35 s there against ~45 ms for the later ones).
    if (rc == SZ_OK && S->world > 1 && !getenv("SZ_SLAB_NO_PREWARM")) {
        rc = do_rebuild(S);
        S->rebuilds = 0;
    }
#endif
    return rc;
}

int32_t FN(slab_local_count)(sz_slab *S, int32_t k, int64_t *n, int64_t *n_owned) {
    if (!S || k < 0 || k >= S->n_local) return SZ_ERR_INVALID;
    if (n) *n = (int64_t)S->ranks[k].gidx.size();
    if (n_owned) *n_owned = S->ranks[k].n_owned;
    return SZ_OK;
}

int32_t FN(slab_local_index)(sz_slab *S, int32_t k, int64_t *gidx, int32_t *owner) {
    if (!S || k < 0 || k >= S->n_local) return SZ_ERR_INVALID;
    const Rank &R = S->ranks[k];
    if (gidx) std::copy(R.gidx.begin(), R.gidx.end(), gidx);
    if (owner) std::copy(R.owner.begin(), R.owner.end(), owner);
    return SZ_OK;
}

static int32_t slab_step_common(sz_slab *S, int64_t tstep, int32_t do_coupling, const sz_floe_soa *const *in, sz_floe_soa *const *out,
                                bool partial = false) {
    if (!S) return SZ_ERR_INVALID;
    for (Rank &R : S->ranks) if (!R.built) return sfail(S, SZ_ERR_INVALID, "slab: step before sz_slab_build");
    const bool host_mode = out != nullptr;
#ifndef SZ_ORACLE_BUILD
    // one host thread drives every local device: uploads + publication everywhere, kernels everywhere, wait everywhere
    for (int k = 0; k < S->n_local; ++k) {
        Rank &R = S->ranks[k];
        HCK(szb_step_publish(R.h, tstep, do_coupling, host_mode && in ? in[k] : nullptr, host_mode ? out[k] : nullptr,
                             host_mode ? (partial ? 2 : 1) : 0), "step");
    }
    for (int k = 0; k < S->n_local; ++k) {
        Rank &R = S->ranks[k];
        HCK(szb_step_begin(R.h), "step");
    }
    for (int k = 0; k < S->n_local; ++k) {
        Rank &R = S->ranks[k];
        HCK(szb_step_end(R.h, host_mode), "step");
    }
#else
    if (host_mode && !partial)
        for (int k = 0; k < S->n_local; ++k) {
            Rank &R = S->ranks[k];
            HCK(FN(upload_state)(R.h, in[k]), "upload_state");
        }
    if (partial) {  // the CPU restatement: sz_step_host_partial's upload half (kernels none), then the exchange
        for (int k = 0; k < S->n_local; ++k) {
            Rank &R = S->ranks[k];
            HCK(szo_upload_partial(R.h, do_coupling, in ? in[k] : nullptr), "upload_partial");
        }
    }
    {
        int32_t rc = host_exchange(S);
        if (rc) return rc;
    }
    for (int k = 0; k < S->n_local; ++k) {
        Rank &R = S->ranks[k];
        HCK(FN(step)(R.h, tstep, do_coupling), "step");
        if (host_mode) {  // like sz_step_host: Monte-Carlo points and ghost lists are not transferred
            sz_floe_soa o = *out[k];
            o.mc_x = o.mc_y = nullptr;
            o.ghost_index = nullptr;
            HCK(FN(download_floes)(R.h, &o), "download_floes");
            out[k]->n = o.n;
            out[k]->n_init = o.n_init;
        }
        int32_t rc = host_displacement(S, R);
        if (rc) return rc;
    }
#endif
    // every rank lives here: the library renews the lists by itself
    if (S->n_local == S->world && S->world > 1 && local_max_disp(S) > 0.5 * S->skin) return do_rebuild(S);
    return SZ_OK;
}

int32_t FN(slab_step)(sz_slab *S, int64_t tstep, int32_t do_coupling) { return slab_step_common(S, tstep, do_coupling, nullptr, nullptr); }

int32_t FN(slab_step_host)(sz_slab *S, int64_t tstep, int32_t do_coupling, const sz_floe_soa *const *in, sz_floe_soa *const *out) {
    if (!S || !in || !out) return SZ_ERR_INVALID;
    for (int k = 0; k < S->n_local; ++k)
        if (!in[k] || !out[k]) return sfail(S, SZ_ERR_INVALID, "slab_step_host: arrays of every local rank are required");
    return slab_step_common(S, tstep, do_coupling, in, out);
}

int32_t FN(slab_step_host_partial)(sz_slab *S, int64_t tstep, int32_t do_coupling, const sz_floe_soa *const *in, sz_floe_soa *const *out) {
    if (!S || !out) return SZ_ERR_INVALID;
    for (int k = 0; k < S->n_local; ++k)
        if (!out[k]) return sfail(S, SZ_ERR_INVALID, "slab_step_host_partial: output arrays of every local rank are required");
    return slab_step_common(S, tstep, do_coupling, in, out, true);
}

int32_t FN(slab_max_displacement)(sz_slab *S, double *metres) {
    if (!S || !metres) return SZ_ERR_INVALID;
    *metres = local_max_disp(S);
    return SZ_OK;
}

int32_t FN(slab_rebuild)(sz_slab *S) {
    if (!S) return SZ_ERR_INVALID;
    for (Rank &R : S->ranks) if (!R.built) return sfail(S, SZ_ERR_INVALID, "slab: rebuild before sz_slab_build");
    return do_rebuild(S);
}

int32_t FN(slab_refresh_halo)(sz_slab *S) {
    if (!S) return SZ_ERR_INVALID;
    for (Rank &R : S->ranks) if (!R.built) return sfail(S, SZ_ERR_INVALID, "slab: refresh before sz_slab_build");
#ifndef SZ_ORACLE_BUILD
    for (Rank &R : S->ranks) HCK(szb_refresh_publish(R.h), "publish");
    for (Rank &R : S->ranks) HCK(szb_refresh_consume(R.h), "consume");
    return SZ_OK;
#else
    return host_exchange(S);
#endif
}

int32_t FN(slab_stats)(sz_slab *S, int32_t k, int64_t *send_bytes, int64_t *halo_floes, int64_t *rebuilds) {
    if (!S || k < 0 || k >= S->n_local) return SZ_ERR_INVALID;
    const Rank &R = S->ranks[k];
    if (send_bytes) *send_bytes = R.send_bytes;
    if (halo_floes) *halo_floes = (int64_t)R.gidx.size() - R.n_owned;
    if (rebuilds) *rebuilds = S->rebuilds;
    return SZ_OK;
}

}  // extern "C"
