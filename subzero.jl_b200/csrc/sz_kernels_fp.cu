// sz_kernels_fp.cu — K6 (ocean/atmosphere coupling) and K7 (state update).
//
// These two kernels carry no discrete geometric decision that must match the reference bit for
// bit: their parity bar is 1e-9 relative (BASELINE.json north_star).  They are therefore compiled
// with FMA contraction ON (unlike sz_kernels.cu) and use algebraic identities the tolerance covers.
#include <cuda_pipeline.h>
#include <stdlib.h>

#include "sz_common.cuh"

#define FULLMASK 0xffffffffu
void szk_count_launches(int n);

// ---- field packing ---------------------------------------------------------------------------------
// The five coupling fields are interleaved per grid node: (atm_u, atm_v, ocn_u, ocn_v, hflx, 0, 0, 0)
// = 64 B, so one bilinear corner is three vector loads from one cache line instead of five scattered
// 8-byte loads from five arrays.
__global__ void k_pack_fields(const double *__restrict__ au, const double *__restrict__ av,
                              const double *__restrict__ ou, const double *__restrict__ ov,
                              const double *__restrict__ oh, double *__restrict__ out, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double *o = out + (size_t)i * 8;
        o[0] = au[i];
        o[1] = av[i];
        o[2] = ou[i];
        o[3] = ov[i];
        o[4] = oh[i];
        o[5] = o[6] = o[7] = 0.0;
    }
}
void szk_pack_fields(const Launch &L, const Store &S, int n) {
    k_pack_fields<<<(n + 255) / 256, 256, 0, L.stream>>>(S.atm_u, S.atm_v, S.ocn_u, S.ocn_v, S.ocn_hflx, S.fields8, n);
}

// ---- K6: one-way ocean/atmosphere coupling (coupling.jl:1486-1589) ---------------------------------------
// One warp per floe.  Lanes read the floe's body-frame Monte-Carlo points as consecutive double2
// (512 B per warp load, streaming: they are read once per step and must not evict the fields),
// rotate/translate them (calc_subfloe_values!, :627-657), drop points outside a non-periodic grid
// extent (in_bounds, :494-597), gather the five fields bilinearly (mc_interpolation :845-902 ==
// bilinear on the lattice; periodic axes wrap on lines 1..N) and reduce stress and torque with
// shuffles in a fixed order.  With r = (xc, yc), theta = atan(yc, xc): rad sin(theta) = yc and
// rad cos(theta) = xc, so the reference's u - xi rad sin(theta) (:1534-1537) and
// (-tx sin + ty cos) rad (:1562) need no transcendental call.
struct CpConst {
    double x0, y0, xf, yf, inv_dx, inv_dy, dx, dy, ct, sn, ka, ko, f, koct, kosn;
    int Nx, Ny, per_x, per_y;
};

struct CpAcc {
    double tx, ty, trq, hf;
    int n;
};

// what one Monte-Carlo point contributes to the floe -> cell registry (floe_to_grid_info!, coupling.jl:1417-1454)
struct CpReg {
    int cell;        // shifted cell index ix + (Nx+1) iy, -1 = point not in bounds
    int sdx, sdy;    // shifted - unshifted grid-line index (periodic wrap), in cells
    double tox, toy; // ocean stress on the ice at the point
};

// |v| for the quadratic drag laws: MUFU.RSQ64H seed (rsqrt.approx.ftz.f64, relative error ~2^-22) + ONE Newton step
// on the reciprocal root (-> ~1e-13) instead of the IEEE-rounded sqrt() expansion with its slow-path call: ncu r1k had
// the two sqrt() of a point at 23 % of the kernel's instructions, and the kernel is bound by the FP64 pipe.  The
// parity bar of this kernel is 1e-9 relative (sz_kernels_fp.cu header); the tests see ~1e-13.
__device__ __forceinline__ double cp_norm(double s) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(s));
    // one Newton step written on the root itself: y = s r, sqrt(s) ~ y + (r / 2)(s - y y) — four dependent FP64
    // instructions instead of five
    const double y = s * r;
    const double root = fma(0.5 * r, fma(-y, y, s), y);
    return s > 2.2250738585072014e-308 ? root : 0.0;  // ftz: a denormal argument gives +inf, its norm is 0 to 1e-154
}

// largest distance of a floe's sub-floe points from its centroid (body frame): warp per floe
__global__ void k_mc_radius(Store S) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < S.n_init; i += gridDim.x * wpb) {
        double m = 0.0;
        for (long long k = S.mc_off[i] + lane, e = S.mc_off[i + 1]; k < e; k += 32) {
            const double2 p = S.mc[k];
            m = fmax(m, p.x * p.x + p.y * p.y);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(FULLMASK, m, o));
        if (lane == 0) S.mc_r[i] = sqrt(m);
    }
}
void szk_mc_radius(const Launch &L, const Store &S) {
    if (S.n_init <= 0) return;
    long long blocks = ((long long)S.n_init + 7) / 8, cap = (long long)L.sms * 16;
    k_mc_radius<<<(int)(blocks < cap ? blocks : cap), 256, 0, L.stream>>>(S);
}

// ATM / HFLX: false when the atmosphere / heat-flux fields are identically zero (sz_set_fields checks): the
// loads and multiply-adds of a zero field are skipped, the result is the same
// INTERIOR: the floe's bounding circle lies strictly inside the grid extent (warp-uniform, decided per floe): every
// point is in bounds and 0 <= ci < Nx, 0 <= cj < Ny, so the in-bounds test, the clamps and the periodic modulo are
// no-ops and are not executed (15 % of the instructions of a point in ncu r1k); the arithmetic is unchanged
template <bool REG, bool ATM, bool HFLX, bool INTERIOR = false>
__device__ __forceinline__ void cp_point(const CpConst &c, const double *__restrict__ F, double2 b, double ca,
                                         double sa, double cx, double cy, double u, double v, double xi, double mf,
                                         CpAcc &acc, CpReg *reg) {
    double xc = ca * b.x - sa * b.y, yc = sa * b.x + ca * b.y;  // body frame -> world, about the centroid
    double xr, yr, gx, gy;
    if (INTERIOR && !REG) {
        // The kernel is bound by the FP64 issue rate (75 of its ~155 instructions per point): on the interior one-way
        // path the world position itself is never needed.  Lever arm = (xc, yc) instead of ((xc + cx) - cx, ...) — the
        // reference's round trip through a ~1e5 m coordinate differs by ~1e-11 m on a ~1e3 m arm —, grid coordinate
        // = xc / dx + (cx - x0) / dx with the second term constant per floe (hoisted by the compiler): 6 FP64
        // instructions per point instead of 12.  A point within one rounding of a grid line may land in the
        // neighbouring cell with weight 1 - eps instead of eps: the bilinear value is continuous there.  The two-way
        // path (REG), whose cell index is a discrete result, keeps the reference's expressions.
        xr = xc;
        yr = yc;
        gx = fma(xc, c.inv_dx, (cx - c.x0) * c.inv_dx);
        gy = fma(yc, c.inv_dy, (cy - c.y0) * c.inv_dy);
    } else {
        double x = xc + cx, y = yc + cy;
        if (!INTERIOR) {
            bool inb = (c.per_x || (c.x0 <= x && x <= c.xf)) && (c.per_y || (c.y0 <= y && y <= c.yf));
            if (REG) reg->cell = -1;
            if (!inb) return;
        }
        // the reference recomputes (x - cx, y - cy) from the translated point; the difference to (xc, yc) is
        // one rounding of a ~1e5 coordinate, i.e. ~1e-11 m on a ~1e3 m lever arm (1e-14 relative)
        xr = x - cx;
        yr = y - cy;
        gx = (x - c.x0) * c.inv_dx;
        gy = (y - c.y0) * c.inv_dy;
    }
    acc.n++;
    double up = u - xi * yr, vp = v + xi * xr;
    double fx = floor(gx), fy = floor(gy);
    int ci = (int)fx, cj = (int)fy;
    double wx = gx - fx, wy = gy - fy;
    int i0, i1, j0, j1;
    if (INTERIOR) {
        i0 = ci;
        i1 = (c.per_x && ci + 1 == c.Nx) ? 0 : ci + 1;  // periodic axes: line N+1 is line 1 (coupling.jl:722-745)
    } else if (c.per_x) {
        i0 = ci % c.Nx;
        if (i0 < 0) i0 += c.Nx;
        i1 = i0 + 1 == c.Nx ? 0 : i0 + 1;
    } else {
        if (ci >= c.Nx) { ci = c.Nx - 1; wx = 1.0; }
        if (ci < 0) { ci = 0; wx = 0.0; }
        i0 = ci;
        i1 = ci + 1;
    }
    if (INTERIOR) {
        j0 = cj;
        j1 = (c.per_y && cj + 1 == c.Ny) ? 0 : cj + 1;
    } else if (c.per_y) {
        j0 = cj % c.Ny;
        if (j0 < 0) j0 += c.Ny;
        j1 = j0 + 1 == c.Ny ? 0 : j0 + 1;
    } else {
        if (cj >= c.Ny) { cj = c.Ny - 1; wy = 1.0; }
        if (cj < 0) { cj = 0; wy = 0.0; }
        j0 = cj;
        j1 = cj + 1;
    }
    const int s = c.Nx + 1;
    const double2 *n00 = (const double2 *)(F + (size_t)(i0 + s * j0) * 8);
    const double2 *n10 = (const double2 *)(F + (size_t)(i1 + s * j0) * 8);
    const double2 *n01 = (const double2 *)(F + (size_t)(i0 + s * j1) * 8);
    const double2 *n11 = (const double2 *)(F + (size_t)(i1 + s * j1) * 8);
    // bilinear as three lerps per field (6 FP64 instructions instead of 4 + the four shared weights)
#define CP_LERP2(f00, f10, f01, f11) \
    ((f00 + wx * (f10 - f00)) + wy * ((f01 + wx * (f11 - f01)) - (f00 + wx * (f10 - f00))))
    double2 o00 = __ldg(n00 + 1), o10 = __ldg(n10 + 1), o01 = __ldg(n01 + 1), o11 = __ldg(n11 + 1);  // ocn u, v
    double uatm = 0.0, vatm = 0.0, hfl = 0.0;
    if (ATM) {
        double2 a00 = __ldg(n00), a10 = __ldg(n10), a01 = __ldg(n01), a11 = __ldg(n11);  // atm u, v
        uatm = CP_LERP2(a00.x, a10.x, a01.x, a11.x);
        vatm = CP_LERP2(a00.y, a10.y, a01.y, a11.y);
    }
    if (HFLX) {
        double h00 = __ldg((const double *)(n00 + 2)), h10 = __ldg((const double *)(n10 + 2)),
               h01 = __ldg((const double *)(n01 + 2)), h11 = __ldg((const double *)(n11 + 2));
        hfl = CP_LERP2(h00, h10, h01, h11);
    }
    double uocn = CP_LERP2(o00.x, o10.x, o01.x, o11.x);
    double vocn = CP_LERP2(o00.y, o10.y, o01.y, o11.y);
#undef CP_LERP2
    double dua = uatm - up, dva = vatm - vp;  // calc_atmosphere_forcing, coupling.jl:1212-1232
    double na = cp_norm(dua * dua + dva * dva);
    double duo = uocn - up, dvo = vocn - vp;  // calc_ocean_forcing!, coupling.jl:1277-1299
    double no = cp_norm(duo * duo + dvo * dvo);
    double tox = no * (c.koct * duo - c.kosn * dvo), toy = no * (c.kosn * duo + c.koct * dvo);  // ko (cos, sin) folded on the host
    double kna = c.ka * na;
    double tx = fma(kna, dua, fma(-mf, vocn, tox));
    double ty = fma(kna, dva, fma(mf, uocn, toy));
    if (REG) {
        // find_center_cell_index (coupling.jl:466-470) and shift_cell_idx (:1155-1182), 0-based here
        int xi0 = (int)floor(gx + 0.5), yi0 = (int)floor(gy + 0.5);
        int sx = xi0, sy = yi0;
        if (c.per_x) sx = xi0 < 0 ? xi0 + c.Nx : (xi0 >= c.Nx ? xi0 - c.Nx : xi0);
        if (c.per_y) sy = yi0 < 0 ? yi0 + c.Ny : (yi0 >= c.Ny ? yi0 - c.Ny : yi0);
        reg->cell = sx + (c.Nx + 1) * sy;
        reg->sdx = sx - xi0;
        reg->sdy = sy - yi0;
        reg->tox = tox;
        reg->toy = toy;
    }
    acc.tx += tx;
    acc.ty += ty;
    acc.trq = fma(ty, xr, fma(-tx, yr, acc.trq));
    acc.hf += hfl;
}

template <bool ATM, bool HFLX>
__global__ void __launch_bounds__(128, 6) k_coupling(Store S, CpConst c) {
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5, wib = threadIdx.x >> 5;
    const double *__restrict__ F = S.fields8;
    const int n = S.n_init;
    // The Monte-Carlo points stream through a two-stage shared-memory ring filled with cp.async (LDGSTS, 16 B per lane):
    // the loads of the next 128 points — of this floe or, behind its last chunk, of the warp's NEXT floe — are in
    // flight while the current 128 are integrated.  ncu r1m: with four plain loads per lane issued and then consumed,
    // 46 % of the stall samples sat on the first use of a loaded point (2.1 TB/s).  Every lane reads back what it
    // copied itself, so no barrier is needed.
    __shared__ double2 sbuf[2][4][128];  // [stage][load][thread]: 16 KB per block
    const int tid = threadIdx.x, stride = gridDim.x * wpb;
    int stage = 0;
    bool have = false;  // chunk 0 of the current floe is already in flight in `stage`
    int i = blockIdx.x * wpb + wib;
    long long m0 = 0, m1 = 0;
    if (i < n) {
        m0 = S.mc_off[i];
        m1 = S.mc_off[i + 1];
    }
    for (; i < n; i += stride) {
        const double a = S.alpha[i], cx = S.cx[i], cy = S.cy[i], u = S.u[i], v = S.v[i], xi = S.xi[i];
        const double ar = S.area[i], mass = S.mass[i];
        const int inext = i + stride;
        long long m0n = 0, m1n = 0;
        if (inext < n) {
            m0n = S.mc_off[inext];
            m1n = S.mc_off[inext + 1];
        }
        double sa, ca;
        sincos(a, &sa, &ca);
        const double mf = mass / ar * c.f;
        CpAcc acc = {0.0, 0.0, 0.0, 0.0, 0};
        long long k = m0 + lane;
        // sub-floe points lie within max(rmax, mc_r) of the centroid (a small margin covers rounding)
        const double rm = fmax(S.rmax[i], S.mc_r[i]) * (1.0 + 1e-9) + 1e-6;
        const bool interior = cx - rm > c.x0 && cx + rm < c.xf && cy - rm > c.y0 && cy + rm < c.yf;
        const int nchunk = (int)((m1 - m0 + 127) >> 7);
        if (!have && nchunk > 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (k + 32 * j < m1) __pipeline_memcpy_async(&sbuf[stage][j][tid], S.mc + k + 32 * j, sizeof(double2));
            __pipeline_commit();
        }
        have = false;
        for (int cch = 0; cch < nchunk; ++cch, k += 128) {
            const int nxt = stage ^ 1;
            if (cch + 1 < nchunk) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (k + 128 + 32 * j < m1) __pipeline_memcpy_async(&sbuf[nxt][j][tid], S.mc + k + 128 + 32 * j, sizeof(double2));
            } else if (m1n > m0n) {  // behind the last chunk: the first chunk of the warp's next floe
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (m0n + lane + 32 * j < m1n) __pipeline_memcpy_async(&sbuf[nxt][j][tid], S.mc + m0n + lane + 32 * j, sizeof(double2));
                have = true;
            }
            __pipeline_commit();
            __pipeline_wait_prior(1);
            if (interior && k + 96 < m1) {
                cp_point<false, ATM, HFLX, true>(c, F, sbuf[stage][0][tid], ca, sa, cx, cy, u, v, xi, mf, acc, nullptr);
                cp_point<false, ATM, HFLX, true>(c, F, sbuf[stage][1][tid], ca, sa, cx, cy, u, v, xi, mf, acc, nullptr);
                cp_point<false, ATM, HFLX, true>(c, F, sbuf[stage][2][tid], ca, sa, cx, cy, u, v, xi, mf, acc, nullptr);
                cp_point<false, ATM, HFLX, true>(c, F, sbuf[stage][3][tid], ca, sa, cx, cy, u, v, xi, mf, acc, nullptr);
            } else if (interior) {
#pragma unroll 1
                for (int j = 0; j < 4; ++j)
                    if (k + 32 * j < m1)
                        cp_point<false, ATM, HFLX, true>(c, F, sbuf[stage][j][tid], ca, sa, cx, cy, u, v, xi, mf, acc, nullptr);
            } else {  // floes at the edge of the grid (about 1 % of a large field): the general path, not unrolled
#pragma unroll 1
                for (int j = 0; j < 4; ++j)
                    if (k + 32 * j < m1)
                        cp_point<false, ATM, HFLX, false>(c, F, sbuf[stage][j][tid], ca, sa, cx, cy, u, v, xi, mf, acc, nullptr);
            }
            stage = nxt;
        }
        m0 = m0n;
        m1 = m1n;
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            acc.tx += __shfl_xor_sync(FULLMASK, acc.tx, o);
            acc.ty += __shfl_xor_sync(FULLMASK, acc.ty, o);
            acc.trq += __shfl_xor_sync(FULLMASK, acc.trq, o);
            acc.hf += __shfl_xor_sync(FULLMASK, acc.hf, o);
            acc.n += __shfl_xor_sync(FULLMASK, acc.n, o);
        }
        if (lane == 0) {
            if (acc.n == 0) {
                S.cpl_remove[i] = 1;  // coupling.jl:1507-1508; applied to status.tag by k_apply_coupling_tags
            } else {
                S.cpl_remove[i] = 0;
                double np_ = (double)acc.n;
                double tot_x = np_ * (mf * v) + acc.tx, tot_y = -np_ * (mf * u) + acc.ty;  // Coriolis, :1522-1525
                S.fxOA[i] = tot_x / np_ * ar;  // :1583-1586
                S.fyOA[i] = tot_y / np_ * ar;
                S.trqOA[i] = acc.trq / np_ * ar;
                S.hflx[i] = acc.hf / np_;
            }
        }
    }
    __pipeline_wait_prior(0);
}

// ---- the same kernel with the Monte-Carlo points brought in by the bulk-copy engine (cp.async.bulk, SASS UBLKCP) ----
// A floe's points are ONE contiguous block of the CSR array — the one perfectly contiguous stream of the step — so a
// chunk of 128 points (2 KB) is a single 1-D bulk copy global -> shared issued by one lane and completed on an
// mbarrier (complete_tx::bytes); the other 31 lanes issue nothing for the load at all (the LDGSTS version spends four
// predicated 16-byte copies + their address arithmetic per lane and chunk in a kernel that is bound by its issue slots).
// Two stages per warp, each with its own barrier; the copy of the next chunk — of this floe or, behind its last
// chunk, of the warp's next floe — is in flight while the current one is integrated.
__device__ __forceinline__ unsigned cp_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_bulk_load(double2 *dst, const double2 *src, unsigned bytes, unsigned long long *bar) {
    const unsigned d = cp_smem_u32(dst), b = cp_smem_u32(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(src), "r"(bytes), "r"(b)
                 : "memory");
}
__device__ __forceinline__ void cp_bar_wait(unsigned long long *bar, unsigned parity) {
    const unsigned b = cp_smem_u32(bar);
    unsigned ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok)
                     : "r"(b), "r"(parity)
                     : "memory");
    } while (!ok);
}

#ifndef CP_MINB
#define CP_MINB 6  // blocks per SM the register allocation aims at (A/B: tools/ab_coupling_minb.sh)
#endif
template <bool ATM, bool HFLX>
__global__ void __launch_bounds__(128, CP_MINB) k_coupling_bulk(Store S, CpConst c) {
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5, wib = threadIdx.x >> 5;
    const double *__restrict__ F = S.fields8;
    const int n = S.n_init;
    __shared__ __align__(128) double2 sbuf[2][4][128];  // [stage][warp][point]: 16 KB per block
    __shared__ __align__(8) unsigned long long mbar[2][4];
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(cp_smem_u32(&mbar[0][wib])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(cp_smem_u32(&mbar[1][wib])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    const int stride = gridDim.x * wpb;
    int stage = 0;
    unsigned ph0 = 0, ph1 = 0;  // phase parity of the two barriers
    bool have = false;          // chunk 0 of the current floe is already in flight in `stage`
    int i = blockIdx.x * wpb + wib;
    long long m0 = 0, m1 = 0;
    if (i < n) {
        m0 = S.mc_off[i];
        m1 = S.mc_off[i + 1];
    }
    for (; i < n; i += stride) {
        const double a = S.alpha[i], cx = S.cx[i], cy = S.cy[i], u = S.u[i], v = S.v[i], xi = S.xi[i];
        const double ar = S.area[i], mass = S.mass[i];
        const int inext = i + stride;
        long long m0n = 0, m1n = 0;
        if (inext < n) {
            m0n = S.mc_off[inext];
            m1n = S.mc_off[inext + 1];
        }
        double sa, ca;
        sincos(a, &sa, &ca);
        const double mf = mass / ar * c.f;
        CpAcc acc = {0.0, 0.0, 0.0, 0.0, 0};
        long long k = m0;
        const double rm = fmax(S.rmax[i], S.mc_r[i]) * (1.0 + 1e-9) + 1e-6;
        const bool interior = cx - rm > c.x0 && cx + rm < c.xf && cy - rm > c.y0 && cy + rm < c.yf;
        const int nchunk = (int)((m1 - m0 + 127) >> 7);
        if (!have && nchunk > 0 && lane == 0) {
            const long long cntp = m1 - k < 128 ? m1 - k : 128;
            cp_bulk_load(&sbuf[stage][wib][0], S.mc + k, (unsigned)(cntp * 16), &mbar[stage][wib]);
        }
        have = false;
        for (int cch = 0; cch < nchunk; ++cch, k += 128) {
            const int nxt = stage ^ 1;
            __syncwarp();  // every lane has consumed what the previous iteration read from stage `nxt`
            if (cch + 1 < nchunk) {
                if (lane == 0) {
                    const long long rest = m1 - (k + 128), cntp = rest < 128 ? rest : 128;
                    cp_bulk_load(&sbuf[nxt][wib][0], S.mc + k + 128, (unsigned)(cntp * 16), &mbar[nxt][wib]);
                }
            } else if (m1n > m0n) {  // behind the last chunk: the first chunk of the warp's next floe
                if (lane == 0) {
                    const long long rest = m1n - m0n, cntp = rest < 128 ? rest : 128;
                    cp_bulk_load(&sbuf[nxt][wib][0], S.mc + m0n, (unsigned)(cntp * 16), &mbar[nxt][wib]);
                }
                have = true;
            }
            if (stage == 0) { cp_bar_wait(&mbar[0][wib], ph0); ph0 ^= 1u; }
            else { cp_bar_wait(&mbar[1][wib], ph1); ph1 ^= 1u; }
            const double2 *sb = &sbuf[stage][wib][lane];
            if (interior && k + 128 <= m1) {
                cp_point<false, ATM, HFLX, true>(c, F, sb[0], ca, sa, cx, cy, u, v, xi, mf, acc, nullptr);
                cp_point<false, ATM, HFLX, true>(c, F, sb[32], ca, sa, cx, cy, u, v, xi, mf, acc, nullptr);
                cp_point<false, ATM, HFLX, true>(c, F, sb[64], ca, sa, cx, cy, u, v, xi, mf, acc, nullptr);
                cp_point<false, ATM, HFLX, true>(c, F, sb[96], ca, sa, cx, cy, u, v, xi, mf, acc, nullptr);
            } else if (interior) {
#pragma unroll 1
                for (int j = 0; j < 4; ++j)
                    if (k + lane + 32 * j < m1)
                        cp_point<false, ATM, HFLX, true>(c, F, sb[32 * j], ca, sa, cx, cy, u, v, xi, mf, acc, nullptr);
            } else {  // floes at the edge of the grid (about 1 % of a large field): the general path, not unrolled
#pragma unroll 1
                for (int j = 0; j < 4; ++j)
                    if (k + lane + 32 * j < m1)
                        cp_point<false, ATM, HFLX, false>(c, F, sb[32 * j], ca, sa, cx, cy, u, v, xi, mf, acc, nullptr);
            }
            stage = nxt;
        }
        m0 = m0n;
        m1 = m1n;
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            acc.tx += __shfl_xor_sync(FULLMASK, acc.tx, o);
            acc.ty += __shfl_xor_sync(FULLMASK, acc.ty, o);
            acc.trq += __shfl_xor_sync(FULLMASK, acc.trq, o);
            acc.hf += __shfl_xor_sync(FULLMASK, acc.hf, o);
            acc.n += __shfl_xor_sync(FULLMASK, acc.n, o);
        }
        if (lane == 0) {
            if (acc.n == 0) {
                S.cpl_remove[i] = 1;  // coupling.jl:1507-1508; applied to status.tag by k_apply_coupling_tags
            } else {
                S.cpl_remove[i] = 0;
                double np_ = (double)acc.n;
                double tot_x = np_ * (mf * v) + acc.tx, tot_y = -np_ * (mf * u) + acc.ty;  // Coriolis, :1522-1525
                S.fxOA[i] = tot_x / np_ * ar;  // :1583-1586
                S.fyOA[i] = tot_y / np_ * ar;
                S.trqOA[i] = acc.trq / np_ * ar;
                S.hflx[i] = acc.hf / np_;
            }
        }
    }
}

// The same integration plus the floe -> cell registry (grid.floe_locations / ocean.scells): the points of a
// warp iteration are grouped by cell with ballots, each group is reduced and added to a small per-floe table in
// shared memory; the table is appended to the global record list (sorted by (cell, floe) afterwards).
#define CP_TABLE 32     // distinct cells of one floe held in shared memory ...
#define CP_SPILL 2048   // ... and in a block of global memory claimed on demand (floes wider than ~5 grid cells)
__global__ void __launch_bounds__(128, 6) k_coupling_reg(Store S, CouplingBuf CB, CpConst c) {
    __shared__ int t_cell[4][CP_TABLE], t_n[4][CP_TABLE], t_sd[4][CP_TABLE][2];
    __shared__ double t_t[4][CP_TABLE][2];
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5, wib = threadIdx.x >> 5;
    const double *__restrict__ F = S.fields8;
    const int n = S.n_init;
    for (int i = blockIdx.x * wpb + wib; i < n; i += gridDim.x * wpb) {
        const double a = S.alpha[i], cx = S.cx[i], cy = S.cy[i], u = S.u[i], v = S.v[i], xi = S.xi[i];
        double sa, ca;
        sincos(a, &sa, &ca);
        const double ar = S.area[i];
        const double mf = S.mass[i] / ar * c.f;
        CpAcc acc = {0.0, 0.0, 0.0, 0.0, 0};
        const long long m0 = S.mc_off[i], m1 = S.mc_off[i + 1];
        int ntab = 0, nsp = 0;   // entries in the shared table / in this floe's spill block (warp-uniform)
        long long sp = -1;       // first entry of the spill block, -1 = none claimed yet, -2 = no room (step is repeated)
        bool overflow = false;
        for (long long base = m0; base < m1; base += 32) {
            const long long k = base + lane;
            CpReg r;
            r.cell = -1;
            r.sdx = r.sdy = 0;
            r.tox = r.toy = 0.0;
            if (k < m1) cp_point<true, true, true>(c, F, __ldcs(S.mc + k), ca, sa, cx, cy, u, v, xi, mf, acc, &r);
            unsigned pending = __ballot_sync(FULLMASK, r.cell >= 0);
            while (pending) {
                const int leader = __ffs(pending) - 1;
                const int cc = __shfl_sync(FULLMASK, r.cell, leader);
                const int sdx = __shfl_sync(FULLMASK, r.sdx, leader), sdy = __shfl_sync(FULLMASK, r.sdy, leader);
                const bool mine = r.cell == cc;
                const unsigned m = __ballot_sync(FULLMASK, mine);
                double sx_ = mine ? -r.tox : 0.0, sy_ = mine ? -r.toy : 0.0;  // add_point! stores the stress ON the ocean
#pragma unroll
                for (int o = 16; o; o >>= 1) {
                    sx_ += __shfl_xor_sync(FULLMASK, sx_, o);
                    sy_ += __shfl_xor_sync(FULLMASK, sy_, o);
                }
                int e = -1;
                if (lane == 0) {
                    int q = 0;
                    while (q < ntab && t_cell[wib][q] != cc) ++q;
                    if (q < ntab) e = q;
                    else if (ntab < CP_TABLE) {
                        t_cell[wib][q] = cc;
                        t_n[wib][q] = 0;
                        t_sd[wib][q][0] = sdx;
                        t_sd[wib][q][1] = sdy;
                        t_t[wib][q][0] = t_t[wib][q][1] = 0.0;
                        e = q;
                        ntab++;
                    }
                    if (e >= 0) {
                        t_t[wib][e][0] += sx_;
                        t_t[wib][e][1] += sy_;
                        t_n[wib][e] += __popc(m);
                    }
                }
                e = __shfl_sync(FULLMASK, e, 0);
                ntab = __shfl_sync(FULLMASK, ntab, 0);
                if (e < 0) {  // the shared table is full: this floe's block of the global spill table
                    if (sp == -1) {
                        if (lane == 0) {
                            sp = (long long)atomicAdd(&cnt->n_spill, CP_SPILL);
                            if (sp + CP_SPILL > CB.cap_spill) {
                                atomicOr(&cnt->error, ERR_SPILL_CAP);
                                sp = -2;
                            }
                        }
                        sp = __shfl_sync(FULLMASK, sp, 0);
                    }
                    if (sp >= 0) {
                        int f = -1;
                        for (int b = 0; b < nsp && f < 0; b += 32) {  // the search is spread over the lanes
                            const int idx = b + lane;
                            const unsigned hit = __ballot_sync(FULLMASK, idx < nsp && CB.sp_cell[sp + idx] == cc);
                            if (hit) f = b + __ffs(hit) - 1;
                        }
                        if (f < 0) {
                            if (nsp < CP_SPILL) {
                                f = nsp++;
                                if (lane == 0) {
                                    CB.sp_cell[sp + f] = cc;
                                    CB.sp_n[sp + f] = 0;
                                    CB.sp_sd[sp + f] = make_int2(sdx, sdy);
                                    CB.sp_t[sp + f] = make_double2(0.0, 0.0);
                                }
                            } else {
                                overflow = true;
                            }
                        }
                        if (f >= 0 && lane == 0) {
                            double2 t = CB.sp_t[sp + f];
                            CB.sp_t[sp + f] = make_double2(t.x + sx_, t.y + sy_);
                            CB.sp_n[sp + f] += __popc(m);
                        }
                        __syncwarp();
                    }
                }
                pending &= ~m;
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            acc.tx += __shfl_xor_sync(FULLMASK, acc.tx, o);
            acc.ty += __shfl_xor_sync(FULLMASK, acc.ty, o);
            acc.trq += __shfl_xor_sync(FULLMASK, acc.trq, o);
            acc.hf += __shfl_xor_sync(FULLMASK, acc.hf, o);
            acc.n += __shfl_xor_sync(FULLMASK, acc.n, o);
        }
        __syncwarp();
        if (lane == 0) {
            if (overflow) atomicOr(&cnt->error, ERR_CELL_TABLE);
            if (acc.n == 0) {
                S.cpl_remove[i] = 1;
            } else {
                S.cpl_remove[i] = 0;
                double np_ = (double)acc.n;
                double tot_x = np_ * (mf * v) + acc.tx, tot_y = -np_ * (mf * u) + acc.ty;
                S.fxOA[i] = tot_x / np_ * ar;
                S.fyOA[i] = tot_y / np_ * ar;
                S.trqOA[i] = acc.trq / np_ * ar;
                S.hflx[i] = acc.hf / np_;
            }
        }
        int slot = 0;
        const int nrec = ntab + nsp;
        if (lane == 0 && nrec > 0) {
            slot = atomicAdd(&cnt->n_crec, nrec);
            if (slot + nrec > CB.cap_crec) {
                atomicOr(&cnt->error, ERR_CREC_CAP);
                slot = -1;
            }
        }
        slot = __shfl_sync(FULLMASK, slot, 0);
        if (slot >= 0 && lane < ntab) {
            const int e = lane, q = slot + e;
            CB.rec_cell[q] = t_cell[wib][e];
            CB.rec_floe[q] = i;
            CB.rec_npts[q] = t_n[wib][e];
            CB.rec_t[q] = make_double2(t_t[wib][e][0], t_t[wib][e][1]);
            CB.rec_d[q] = make_double2(t_sd[wib][e][0] * c.dx, t_sd[wib][e][1] * c.dy);
        }
        if (slot >= 0 && sp >= 0)
            for (int e = lane; e < nsp; e += 32) {
                const int q = slot + ntab + e;
                const int2 sd = CB.sp_sd[sp + e];
                CB.rec_cell[q] = CB.sp_cell[sp + e];
                CB.rec_floe[q] = i;
                CB.rec_npts[q] = CB.sp_n[sp + e];
                CB.rec_t[q] = CB.sp_t[sp + e];
                CB.rec_d[q] = make_double2(sd.x * c.dx, sd.y * c.dy);
            }
        __syncwarp();
    }
}

__global__ void k_crec_reset(Counters *cnt) {
    if (!cnt->error) {
        cnt->n_crec = 0;
        cnt->n_spill = 0;
    }
}

void szk_coupling_reg(const Launch &L, const Store &S, const CouplingBuf &CB, const Params &P) {
    k_crec_reset<<<1, 1, 0, L.stream>>>(S.cnt);
    szk_count_launches(1);
    if (S.n_init <= 0) return;
    CpConst c;
    c.x0 = P.x0; c.y0 = P.y0; c.xf = P.xf; c.yf = P.yf;
    c.inv_dx = 1.0 / P.dx; c.inv_dy = 1.0 / P.dy;
    c.dx = P.dx; c.dy = P.dy;
    c.ct = cos(P.cfg.turn_theta); c.sn = sin(P.cfg.turn_theta);
    c.ka = P.cfg.rho_a * P.cfg.Cd_ia; c.ko = P.cfg.rho_o * P.cfg.Cd_io; c.f = P.cfg.f;
    c.koct = c.ko * c.ct; c.kosn = c.ko * c.sn;
    c.Nx = P.Nx; c.Ny = P.Ny;
    c.per_x = P.per_x; c.per_y = P.per_y;
    long long blocks = ((long long)S.n_init + 3) / 4, cap = (long long)L.sms * 48;
    k_coupling_reg<<<(int)(blocks < cap ? blocks : cap), 128, 0, L.stream>>>(S, CB, c);
    szk_count_launches(1);
}

// two-way coupling, last pass (calc_two_way_coupling!, coupling.jl:1617-1680): one thread per grid cell walks its
// records in ascending floe order (the reference's order), then adds the atmosphere's drag on the open water
// and refreshes ocean.hflx_factor.
__global__ void k_cells_final(Store S, CouplingBuf CB, Params P) {
    if (S.cnt->error) return;
    const int ncell = (P.Nx + 1) * (P.Ny + 1);
    const double cell_area = P.dx * P.dy;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < ncell; q += gridDim.x * blockDim.x) {
        double tx = 0, ty = 0, si = 0;
        for (int k = CB.cell_start[q], ke = CB.cell_start[q + 1]; k < ke; ++k) {
            const int r = CB.perm[k];
            const double a = CB.rec_area[r];
            if (a > 0) {
                const double2 t = CB.rec_t[r];
                const double np_ = (double)CB.rec_npts[r];
                tx += (t.x / np_) * a;
                ty += (t.y / np_) * a;
                si += a;
            }
        }
        if (si > 0) {
            tx /= si;
            ty /= si;
            si /= cell_area;
        }
        double du = S.atm_u[q] - S.ocn_u[q], dv = S.atm_v[q] - S.ocn_v[q];
        double ocn_frac = 1 - si, norm = sqrt(du * du + dv * dv);
        tx += P.cfg.rho_a * P.cfg.Cd_ao * ocn_frac * norm * du;
        ty += P.cfg.rho_a * P.cfg.Cd_ao * ocn_frac * norm * dv;
        S.taux[q] = tx;
        S.tauy[q] = ty;
        S.sifrac[q] = si;
        double hf = P.cfg.dt * P.cfg.k / (P.cfg.rho_i * P.cfg.L) * (S.ocn_temp[q] - S.atm_temp[q]);
        S.ocn_hflx[q] = hf;
        S.fields8[(size_t)q * 8 + 4] = hf;  // the packed copy the next coupling step interpolates
    }
}
void szk_cells_final(const Launch &L, const Store &S, const CouplingBuf &CB, const Params &P) {
    int ncell = (P.Nx + 1) * (P.Ny + 1);
    k_cells_final<<<(ncell + 127) / 128, 128, 0, L.stream>>>(S, CB, P);
    szk_count_launches(1);
}

void szk_coupling(const Launch &L, const Store &S, const Params &P) {
    if (S.n_init <= 0) return;
    CpConst c;
    c.x0 = P.x0; c.y0 = P.y0; c.xf = P.xf; c.yf = P.yf;
    c.inv_dx = 1.0 / P.dx; c.inv_dy = 1.0 / P.dy;
    c.dx = P.dx; c.dy = P.dy;
    c.ct = cos(P.cfg.turn_theta); c.sn = sin(P.cfg.turn_theta);
    c.ka = P.cfg.rho_a * P.cfg.Cd_ia; c.ko = P.cfg.rho_o * P.cfg.Cd_io; c.f = P.cfg.f;
    c.koct = c.ko * c.ct; c.kosn = c.ko * c.sn;
    c.Nx = P.Nx; c.Ny = P.Ny;
    c.per_x = P.per_x; c.per_y = P.per_y;
    // 128-thread blocks: small enough to share an SM with the narrow-phase blocks when the two run on
    // different streams (sz_step overlaps coupling with the collision kernels)
    // L.coupling_blocks_per_sm > 0 (sz_step: coupling shares the SMs with the collision kernels): a persistent
    // grid of that many blocks per SM leaves registers for the high-priority stream's blocks
    long long blocks = ((long long)S.n_init + 3) / 4;
    long long cap = (long long)L.sms * (L.coupling_blocks_per_sm > 0 ? L.coupling_blocks_per_sm : 48);
    const int g = (int)(blocks < cap ? blocks : cap);
    const bool atm = P.atm_nonzero, hf = P.hflx_nonzero;
    static const bool ldgsts = getenv("SZ_COUPLING_LDGSTS") != nullptr;  // A/B: the cp.async (LDGSTS) ring of round 1
    if (ldgsts) {
        if (atm && hf) k_coupling<true, true><<<g, 128, 0, L.stream>>>(S, c);
        else if (atm) k_coupling<true, false><<<g, 128, 0, L.stream>>>(S, c);
        else if (hf) k_coupling<false, true><<<g, 128, 0, L.stream>>>(S, c);
        else k_coupling<false, false><<<g, 128, 0, L.stream>>>(S, c);
    } else {
        if (atm && hf) k_coupling_bulk<true, true><<<g, 128, 0, L.stream>>>(S, c);
        else if (atm) k_coupling_bulk<true, false><<<g, 128, 0, L.stream>>>(S, c);
        else if (hf) k_coupling_bulk<false, true><<<g, 128, 0, L.stream>>>(S, c);
        else k_coupling_bulk<false, false><<<g, 128, 0, L.stream>>>(S, c);
    }
    szk_count_launches(1);
}

// status.tag = remove for floes without an in-bounds Monte-Carlo point (coupling.jl:1507-1508).  A separate
// pass because coupling may run concurrently with the collision kernels, which also write status.tag.
__global__ void k_apply_coupling_tags(Store S) {
    if (S.cnt->error) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < S.n_init; i += gridDim.x * blockDim.x)
        if (S.cpl_remove[i]) S.status[i] = SZ_STATUS_REMOVE;
}
void szk_apply_coupling_tags(const Launch &L, const Store &S) {
    if (S.n_init <= 0) return;
    k_apply_coupling_tags<<<(S.n_init + 255) / 256, 256, 0, L.stream>>>(S);
    szk_count_launches(1);
}

__global__ void k_apply_remove_flags(Store S, const int *__restrict__ flags, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        if (flags[i]) S.status[i] = SZ_STATUS_REMOVE;
}
void szk_apply_remove_flags(const Launch &L, const Store &S, const int *flags, int n) {
    if (n <= 0) return;
    k_apply_remove_flags<<<(n + 255) / 256, 256, 0, L.stream>>>(S, flags, n);
}

// ---- K7: state update (update_floe.jl:392-551) -------------------------------------------------------------
// calc_strain! (:425-453) evaluates u - xi r sin(theta), u + xi r cos(theta)
// at every vertex; with r sin(theta) = y and r cos(theta) = x those are u - xi y and u + xi x
// (the v terms use floe.u exactly as the reference does, :441-442).
// One THREAD per floe.  The update is ~25 scalars and a handful of divisions per floe plus a loop over ~7 ring points
// and ~6 rows: as a warp per floe (r1e-r1k) every one of the ~840 scalar instructions was issued once per FLOE
// (83.7 M warp instructions per step at 100 k floes, 42 % issue slots); per thread they are issued once per 32
// floes and the SoA loads are coalesced across the warp.  Rings of consecutive floes are contiguous, so the ring
// loop of a warp still walks one compact region.
__global__ void __launch_bounds__(128) k_update(Store S, StepBuf B, Params P) {
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    const double dt = (double)P.cfg.dt;
    const int n = S.n_init;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t warn = 0;
        double cfx = S.cfx[i], cfy = S.cfy[i], ctrq = S.ctrq[i];
        const double cx = S.cx[i], cy = S.cy[i], area = S.area[i];
        double height = S.height[i], mass = S.mass[i], moment = S.moment[i];
        const double hflx = S.hflx[i], u0 = S.u[i], v0 = S.v[i], xi0 = S.xi[i], alpha0 = S.alpha[i];
        const double pdx = S.p_dxdt[i], pdy = S.p_dydt[i], pda = S.p_dalphadt[i];
        const double pdu = S.p_dudt[i], pdv = S.p_dvdt[i], pdxi = S.p_dxidt[i];
        const double fxOA = S.fxOA[i], fyOA = S.fyOA[i], trqOA = S.trqOA[i];
        const double2 acc01 = ((const double2 *)S.stress_accum)[2 * (size_t)i], acc23 = ((const double2 *)S.stress_accum)[2 * (size_t)i + 1];
        const int r0 = B.row_off[i], r1 = B.row_off[i + 1];
        const int vs = S.vstart[i], nv = S.vcount[i];
        // calc_stress!, :392-414 (pre-move centroid)
        double s11 = 0, s12 = 0, s22 = 0;
        if (r1 > r0) {
            for (int k = r0; k < r1; ++k) {
                const double *r = B.rows + (size_t)k * NCOL;
                double fx = r[COL_FX], fy = r[COL_FY], px = r[COL_PX], py = r[COL_PY];
                s11 += (px - cx) * fx;
                s12 += (py - cy) * fx + (px - cx) * fy;
                s22 += (py - cy) * fy;
            }
            s12 *= 0.5;
            double inv = 1 / (area * height);
            s11 *= inv;
            s12 *= inv;
            s22 *= inv;
        }
        const double lam = P.cfg.stress_lambda;  // stress_calculators.jl:118-122
        ((double2 *)S.stress_accum)[2 * (size_t)i] = make_double2((1 - lam) * acc01.x + lam * s11, (1 - lam) * acc01.y + lam * s12);
        ((double2 *)S.stress_accum)[2 * (size_t)i + 1] = make_double2((1 - lam) * acc23.x + lam * s12, (1 - lam) * acc23.y + lam * s22);
        ((double2 *)S.stress_instant)[2 * (size_t)i] = make_double2(s11, s12);
        ((double2 *)S.stress_instant)[2 * (size_t)i + 1] = make_double2(s12, s22);
        if (height > P.cfg.max_floe_height) {  // :482-485
            height = P.cfg.max_floe_height;
            warn |= SZ_WARN_HEIGHT_CAPPED;
        }
        while (fmax(fabs(cfx), fabs(cfy)) > mass / (5 * dt)) {  // :487-491
            cfx = cfx / 10;
            cfy = cfy / 10;
            ctrq = ctrq / 10;
            warn |= SZ_WARN_FORCE_SCALED;
        }
        double hh = height;  // :494-500
        double dh = hflx / hh;
        double hfrac = (hh + dh) / hh;
        mass *= hfrac;
        moment *= hfrac;
        height -= dh;
        hh = height;
        double Dx = 1.5 * dt * u0 - 0.5 * dt * pdx;  // :503-506
        double Dy = 1.5 * dt * v0 - 0.5 * dt * pdy;
        double Da = 1.5 * dt * xi0 - 0.5 * dt * pda;
        // _move_floe! / _move_poly, floe_utils.jl:74-93: p -> R p + ((R(-c) + c) + D)
        double sn, cs;
        sincos(Da, &sn, &cs);
        double tx = ((cs * (-cx) - sn * (-cy)) + cx) + Dx;
        double ty = ((sn * (-cx) + cs * (-cy)) + cy) + Dy;
        const double ncx = cx + Dx, ncy = cy + Dy;
        double dudt = (fxOA + cfx) / mass;  // :514-531
        double dvdt = (fyOA + cfy) / mass;
        double frac = 1;
        double au = fabs(dt * dudt), av = fabs(dt * dvdt), lim = hh / 2;
        double sgu = (double)((dudt > 0) - (dudt < 0)), sgv = (double)((dvdt > 0) - (dvdt < 0));
        if (au > lim && av > lim) {
            double f1 = (sgu * hh / (2 * dt)) / dudt, f2 = (sgv * hh / (2 * dt)) / dvdt;
            frac = f1 < f2 ? f1 : f2;
        } else if (au > lim && av < lim) frac = (sgu * hh / (2 * dt)) / dudt;
        else if (au < lim && av > lim) frac = (sgv * hh / (2 * dt)) / dvdt;
        if (frac != 1) {
            dudt = frac * dudt;
            dvdt = frac * dvdt;
            warn |= SZ_WARN_VELOCITY_LIMITED;
        }
        const double un = u0 + (1.5 * dt * dudt - 0.5 * dt * pdu);  // :532-535
        const double vn = v0 + (1.5 * dt * dvdt - 0.5 * dt * pdv);
        double dxidt = (trqOA + ctrq) / moment;  // :537-545
        dxidt = frac * dxidt;
        double xin = xi0 + 1.5 * dt * dxidt - 0.5 * dt * pdxi;
        if (fabs(xin) > P.cfg.maximum_xi) {
            xin = (double)((xin > 0) - (xin < 0)) * P.cfg.maximum_xi;
            warn |= SZ_WARN_XI_CLAMPED;
        }
        // rigid move of the ring + calc_strain! on the moved ring with the updated u, xi
        double e11 = 0, e12 = 0, e22 = 0;
        if (nv > 0) {
            double2 p = S.verts[vs];
            double2 q = make_double2((cs * p.x - sn * p.y) + tx, (sn * p.x + cs * p.y) + ty);
            for (int k = 0; k + 1 < nv; ++k) {
                const double2 p2 = S.verts[vs + k + 1];
                const double2 q2 = make_double2((cs * p2.x - sn * p2.y) + tx, (sn * p2.x + cs * p2.y) + ty);
                double x1 = q.x - ncx, y1 = q.y - ncy, x2 = q2.x - ncx, y2 = q2.y - ncy;
                double xd = x2 - x1, yd = y2 - y1;
                double ud = (un - xin * y2) - (un - xin * y1), vd = (un + xin * x2) - (un + xin * x1);
                e11 += ud * yd;
                e12 += ud * xd + vd * yd;
                e22 += vd * xd;
                S.verts[vs + k] = q;
                q = q2;
            }
            S.verts[vs + nv - 1] = q;
        }
        e12 *= 0.5;
        const double iden = 1.0 / (2 * area);
        ((double2 *)S.strain)[2 * (size_t)i] = make_double2(e11 * iden, e12 * iden);
        ((double2 *)S.strain)[2 * (size_t)i + 1] = make_double2(e12 * iden, e22 * iden);
        S.height[i] = height;
        S.mass[i] = mass;
        S.moment[i] = moment;
        S.alpha[i] = alpha0 + Da;
        S.cx[i] = ncx;
        S.cy[i] = ncy;
        S.p_dxdt[i] = u0;  // :509-511
        S.p_dydt[i] = v0;
        S.p_dalphadt[i] = xi0;
        S.u[i] = un;
        S.v[i] = vn;
        S.p_dudt[i] = dudt;
        S.p_dvdt[i] = dvdt;
        S.xi[i] = xin;
        S.p_dxidt[i] = dxidt;
        S.warn[i] = warn;
    }
}

void szk_update(const Launch &L, const Store &S, const StepBuf &B, const Params &P) {
    if (S.n_init <= 0) return;
    long long blocks = ((long long)S.n_init + 127) / 128, cap = (long long)L.sms * 32;
    k_update<<<(int)(blocks < cap ? blocks : cap), 128, 0, L.stream>>>(S, B, P);
    szk_count_launches(1);
}

// ---- halo exchange (slab decomposition) ------------------------------------------------------
// record k of a list: 8 doubles (cx, cy, u, v, xi, height, status, alpha); ring points follow the records
template <bool PACK>
__global__ void k_halo(Store S, const int *__restrict__ idx, const long long *__restrict__ voff, int n, double *buf) {
    double2 *vx = (double2 *)(buf + 8 * (size_t)n);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const int i = idx[k];
        double *r = buf + 8 * (size_t)k;
        const int vs = S.vstart[i], nv = S.vcount[i];
        double2 *v = vx + voff[k];
        if (PACK) {
            r[0] = S.cx[i]; r[1] = S.cy[i]; r[2] = S.u[i]; r[3] = S.v[i]; r[4] = S.xi[i];
            r[5] = S.height[i]; r[6] = (double)S.status[i]; r[7] = S.alpha[i];
            for (int q = 0; q < nv; ++q) v[q] = S.verts[vs + q];
        } else {
            S.cx[i] = r[0]; S.cy[i] = r[1]; S.u[i] = r[2]; S.v[i] = r[3]; S.xi[i] = r[4];
            S.height[i] = r[5]; S.status[i] = (int)r[6]; S.alpha[i] = r[7];
            for (int q = 0; q < nv; ++q) S.verts[vs + q] = v[q];
        }
    }
}
void szk_halo(const Launch &L, const Store &S, const int *idx, const long long *voff, int n, double *buf, bool pack) {
    if (n <= 0) return;
    int blocks = (n + 127) / 128;
    if (pack) k_halo<true><<<blocks, 128, 0, L.stream>>>(S, idx, voff, n, buf);
    else k_halo<false><<<blocks, 128, 0, L.stream>>>(S, idx, voff, n, buf);
    szk_count_launches(1);
}

// ---- slab data plane: one push kernel over peer memory instead of pack -> ncclSend/ncclRecv -> unpack ----------------
// The reference is single-process (collisions.jl:734-864 sees one floe list); this is the per-step halo update of
// SURVEY §8(e).  k_slab_push runs right behind k_update: blockIdx.y = partner; every thread copies one boundary floe
// (8 doubles + its ring) with plain stores into the partner's receive arena — peer-mapped device memory, NVLink —
// fences, and the last block of a partner raises `ready = epoch` there.  The extra y-slice measures how far the owned
// floes travelled since the lists were built.  k_slab_unpack opens the partner's next step: spin on `ready`, scatter
// the records into the store, last block writes `ack = epoch` back.  Arenas alternate with the epoch's parity, so a
// sender only has to see the ack of epoch - 2 (one whole step old: never waited for in practice).
__device__ __forceinline__ int ld_acquire_sys(const int *p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(int *p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Spin until *flag >= want.  A neighbour that never publishes (its step failed, its process died) must not hang this GPU:
// after timeout_ns the step is abandoned with ERR_SLAB_TIMEOUT (every later kernel of the step returns at once).
__device__ __forceinline__ bool slab_wait(const int *flag, int want, unsigned long long timeout_ns, Counters *cnt) {
    if (ld_acquire_sys(flag) >= want) return true;
    const unsigned long long t0 = global_ns();
    for (;;) {
        for (int k = 0; k < 64; ++k) {
            if (ld_acquire_sys(flag) >= want) return true;
            __nanosleep(100);
        }
        if (global_ns() - t0 > timeout_ns) {
            atomicOr(&cnt->error, ERR_SLAB_TIMEOUT);
            return false;
        }
    }
}
__device__ __forceinline__ unsigned long long enc_disp(double x) {  // x >= 0: the bit pattern is order-preserving
    return (unsigned long long)__double_as_longlong(x);
}

__global__ void __launch_bounds__(128) k_slab_push(Store S, SlabDev D, int epoch) {
    Counters *cnt = S.cnt;
    __shared__ int ok;
    if (cnt->error) return;  // an overflowing step is repeated: nothing is published
    const int y = blockIdx.y;
    if (y == D.n_partners) {  // displacement of the owned floes (periodic wrap taken out, collisions.jl:943-949)
        double m = 0.0;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < S.n_init; i += gridDim.x * blockDim.x) {
            if (!D.owned[i]) continue;
            double dx = fabs(S.cx[i] - D.refx[i]), dy = fabs(S.cy[i] - D.refy[i]);
            if (D.period_x > 0.0) dx = fmin(dx, fabs(dx - D.period_x));
            if (D.period_y > 0.0) dy = fmin(dy, fabs(dy - D.period_y));
            m = fmax(m, sqrt(dx * dx + dy * dy));
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(FULLMASK, m, o));
        if ((threadIdx.x & 31) == 0 && m > 0.0) atomicMax(&cnt->slab_disp, enc_disp(m));
        return;
    }
    const SlabPartnerDev &p = D.p[y];
    const int nb = (p.send_n + blockDim.x - 1) / blockDim.x;
    if ((int)blockIdx.x >= nb) return;
    if (threadIdx.x == 0) ok = slab_wait(p.l_ack, epoch - 2, D.timeout_ns, cnt);  // the partner is done with this half of its arena
    __syncthreads();
    if (!ok) return;
    double *buf = p.r_stage[epoch & 1];
    double2 *vx = (double2 *)(buf + 8 * (size_t)p.send_n);
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < p.send_n) {
        const int i = D.send_idx[p.send_off + k];
        double2 *r = (double2 *)(buf + 8 * (size_t)k);
        r[0] = make_double2(S.cx[i], S.cy[i]);
        r[1] = make_double2(S.u[i], S.v[i]);
        r[2] = make_double2(S.xi[i], S.height[i]);
        r[3] = make_double2((double)S.status[i], S.alpha[i]);
        const int vs = S.vstart[i], nv = S.vcount[i];
        double2 *v = vx + D.send_voff[p.send_off + k];
        for (int q = 0; q < nv; ++q) v[q] = S.verts[vs + q];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int done = atomicAdd(&D.push_count[y], 1);
        if (done == nb - 1) {
            D.push_count[y] = 0;
            __threadfence_system();
            st_release_sys(p.r_ready, epoch);
        }
    }
}

__global__ void __launch_bounds__(128) k_slab_unpack(Store S, SlabDev D, int epoch) {
    __shared__ int ok;
    const int y = blockIdx.y;
    const SlabPartnerDev &p = D.p[y];
    const int nb = (p.recv_n + blockDim.x - 1) / blockDim.x;
    if ((int)blockIdx.x >= nb) return;
    if (threadIdx.x == 0) ok = slab_wait(p.l_ready, epoch, D.timeout_ns, S.cnt);
    __syncthreads();
    if (!ok) return;
    const double *buf = p.l_stage[epoch & 1];
    const double2 *vx = (const double2 *)(buf + 8 * (size_t)p.recv_n);
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < p.recv_n) {
        const int i = D.recv_idx[p.recv_off + k];
        const double2 *r = (const double2 *)(buf + 8 * (size_t)k);
        // written by another GPU: read through L2 (ld.cg), never from a stale L1 line
        const double2 a = __ldcg(r), b = __ldcg(r + 1), c = __ldcg(r + 2), d = __ldcg(r + 3);
        S.cx[i] = a.x; S.cy[i] = a.y; S.u[i] = b.x; S.v[i] = b.y; S.xi[i] = c.x; S.height[i] = c.y;
        S.status[i] = (int)d.x; S.alpha[i] = d.y;
        const int vs = S.vstart[i], nv = S.vcount[i];
        const double2 *v = vx + D.recv_voff[p.recv_off + k];
        for (int q = 0; q < nv; ++q) S.verts[vs + q] = __ldcg(v + q);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int done = atomicAdd(&D.unpack_count[y], 1);
        if (done == nb - 1) {
            D.unpack_count[y] = 0;
            st_release_sys(p.r_ack, epoch);
        }
    }
}

void szk_slab_push(const Launch &L, const Store &S, const SlabDev &D, int epoch, int max_send) {
    int bx = (max_send + 127) / 128, bd = (S.n_init + 127) / 128;
    if (bd > 4 * L.sms) bd = 4 * L.sms;
    if (bx < bd) bx = bd;
    if (bx < 1) bx = 1;
    k_slab_push<<<dim3(bx, D.n_partners + 1), 128, 0, L.stream>>>(S, D, epoch);
    szk_count_launches(1);
}
void szk_slab_unpack(const Launch &L, const Store &S, const SlabDev &D, int epoch, int max_recv) {
    if (D.n_partners <= 0 || max_recv <= 0) return;
    k_slab_unpack<<<dim3((max_recv + 127) / 128, D.n_partners), 128, 0, L.stream>>>(S, D, epoch);
    szk_count_launches(1);
}
__global__ void k_slab_reset_disp(Counters *cnt) { cnt->slab_disp = 0ull; }
void szk_slab_reset_disp(const Launch &L, const Store &S) { k_slab_reset_disp<<<1, 1, 0, L.stream>>>(S.cnt); }

// Monte-Carlo points of a re-built local list (sz_slab_rebuild): warp per floe, segment copy
__global__ void k_mc_regather(double2 *__restrict__ dst, const long long *__restrict__ dst_off, const double2 *__restrict__ old_mc,
                              const double2 *__restrict__ extra, const long long *__restrict__ src, int n) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
        const long long a = dst_off[i], m = dst_off[i + 1] - a, s = src[i];
        const double2 *from = s >= 0 ? old_mc + s : extra + (-1 - s);
        for (long long k = lane; k < m; k += 32) dst[a + k] = from[k];
    }
}
void szk_mc_regather(const Launch &L, double2 *dst, const long long *dst_off, const double2 *old_mc, const double2 *extra,
                     const long long *src, int n) {
    if (n <= 0) return;
    long long blocks = ((long long)n + 7) / 8, cap = (long long)L.sms * 16;
    k_mc_regather<<<(int)(blocks < cap ? blocks : cap), 256, 0, L.stream>>>(dst, dst_off, old_mc, extra, src, n);
}
