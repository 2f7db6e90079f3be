/*
 * oracle.c -- CPU restatement of the reference's AW-gridding hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (libskagrid.so) never links, loads or falls back to anything in this directory.
 *
 * Parity status: the reference (Haskell/Accelerate) cannot be built offline (no GHC,
 * no LLVM 7, no libhdf5), so this restatement is pinned only by
 *   - old/BrokenNumbers.hs:85-91  (permute (+) scatter-add golden, tests/golden/)
 *   - test/SmallTest.hs:51-76     (known inputs; three formulations must agree)
 * Rounding ties, FFT normalisation and shift conventions are "parity unpinned"
 * (SURVEY.md section 8c, quirks Q1-Q6 are the working spec).
 *
 * All citations are file:line in /root/reference (sakehl/SKA-SDP-Accelerate-gridding).
 * Build: see oracle/Makefile (gcc -O3 -march=x86-64-v3 -ffp-contract=off -fopenmp; not -march=native, the .so is built in the dev container and shipped to the GPU box).
 *
 * Conventions at this boundary (src/Types.hs:7-28): F = double, complex = interleaved
 * (re,im) doubles, indices = int64, arrays row-major outermost-first, uvw as SoA.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef int64_t i64;

/* Haskell `div` on Int is floor division (differs from C for negatives). */
static inline i64 fdiv(i64 a, i64 b) {
    i64 q = a / b, r = a % b;
    return (r != 0 && ((r < 0) != (b < 0))) ? q - 1 : q;
}

/* ---------------------------------------------------------------------------
 * frac_coord  (src/Gridding.hs:126-140)
 *   x     = halfn + a*n            (mul then add, no FMA)
 *   flx   = floor(x + 0.5/qpx)
 *   fracx = round((x - flx) * qpx) (libm round = half away from zero, Q3)
 * normalise != 0 applies SURVEY Q3's repair of the -1 / qpx tie results:
 *   frac<0  -> frac+=qpx, fl-=1 ;  frac>=qpx -> frac-=qpx, fl+=1
 * ------------------------------------------------------------------------- */
void orc_frac_coord(i64 n, i64 qpx, i64 cnt, const double *p, i64 *fl, i64 *frac,
                    int normalise) {
    const double halfnf = (double)fdiv(n, 2);
    const double nf = (double)n;
    const double qpxf = (double)qpx;
    const double qpxfrac = 0.5 / qpxf;
    for (i64 k = 0; k < cnt; ++k) {
        volatile double prod = p[k] * nf; /* volatile: forbid contraction */
        double x = halfnf + prod;
        double f = floor(x + qpxfrac);
        i64 flx = (i64)f;
        volatile double d = x - (double)flx;
        i64 fr = (i64)round(d * qpxf);
        if (normalise) {
            if (fr < 0) { fr += qpx; flx -= 1; }
            else if (fr >= qpx) { fr -= qpx; flx += 1; }
        }
        fl[k] = flx;
        frac[k] = fr;
    }
}

/* frac_coords (src/Gridding.hs:142-151): x,xf from u with WIDTH; y,yf from v with HEIGHT. */
void orc_frac_coords(i64 h, i64 w, i64 qpx, i64 cnt, const double *u, const double *v,
                     i64 *x, i64 *xf, i64 *y, i64 *yf, int normalise) {
    orc_frac_coord(w, qpx, cnt, u, x, xf, normalise);
    orc_frac_coord(h, qpx, cnt, v, y, yf, normalise);
}

/* ---------------------------------------------------------------------------
 * findClosest (src/Gridding.hs:895-907): binary search with (min,max)=(0,len),
 * loop while (max-min) div 2 >= 1; id=(max+min) div 2; w > ws[id] ? (id,max):(min,id);
 * result r1 iff |w-ws[r1]| < |w-ws[r2]| else r2.  Q4: r2 can equal len (the reference
 * then reads out of bounds); we clamp that case to r1 (= len-1).
 * ------------------------------------------------------------------------- */
i64 orc_find_closest1(i64 len, const double *ws, double w) {
    i64 mn = 0, mx = len;
    while (fdiv(mx - mn, 2) >= 1) {
        i64 id = fdiv(mx + mn, 2);
        if (w > ws[id]) mn = id; else mx = id;
    }
    if (mx >= len) return mn; /* Q4 clamp */
    return (fabs(w - ws[mn]) < fabs(w - ws[mx])) ? mn : mx;
}
void orc_find_closest(i64 len, const double *ws, i64 cnt, const double *w, i64 *out) {
    for (i64 k = 0; k < cnt; ++k) out[k] = orc_find_closest1(len, ws, w[k]);
}

/* uvw_lambda (src/ImageDataset.hs:181-187): a = f/299792458.0 on the host, then a*u. */
void orc_uvw_lambda(double f, i64 cnt, double *u, double *v, double *w) {
    const double a = f / 299792458.0;
    for (i64 k = 0; k < cnt; ++k) { u[k] = a * u[k]; v[k] = a * v[k]; w[k] = a * w[k]; }
}

/* div3 (src/Gridding.hs:838-839): true division by lam. */
void orc_div3(double lam, i64 cnt, double *u, double *v, double *w) {
    for (i64 k = 0; k < cnt; ++k) { u[k] = u[k] / lam; v[k] = v[k] / lam; w[k] = w[k] / lam; }
}

/* mirror_uvw (src/Gridding.hs:551-562): v<0 -> (-u,-v,-w), conj(vis). */
void orc_mirror_uvw(i64 cnt, double *u, double *v, double *w, double *vis) {
    for (i64 k = 0; k < cnt; ++k)
        if (v[k] < 0) { u[k] = -u[k]; v[k] = -v[k]; w[k] = -w[k]; vis[2 * k + 1] = -vis[2 * k + 1]; }
}

/* ---------------------------------------------------------------------------
 * doweight (src/Gridding.hs:564-583): cell = frac_coords (n,n) 1 (uvw/lam);
 * weights = histogram of cells (permute (+) of ones); vis /= weights[cell].
 * u,v here are in wavelengths (NOT yet divided by lam), as the reference passes them.
 * The reference has no bounds check (out-of-grid cells are UB there); we return -1.
 * ------------------------------------------------------------------------- */
int orc_doweight(double theta, i64 lam, i64 cnt, const double *u, const double *v, double *vis) {
    const double lamf = (double)lam;
    const i64 n = (i64)llround(theta * lamf); /* P.round: ties never occur for sane theta*lam */
    double *gw = (double *)calloc((size_t)(n * n), sizeof(double));
    i64 *xy = (i64 *)malloc((size_t)cnt * sizeof(i64));
    if (!gw || !xy) { free(gw); free(xy); return -2; }
    int rc = 0;
    for (i64 k = 0; k < cnt; ++k) {
        double pu = u[k] / lamf, pv = v[k] / lamf;
        i64 x, xf, y, yf;
        orc_frac_coord(n, 1, 1, &pu, &x, &xf, 0);
        orc_frac_coord(n, 1, 1, &pv, &y, &yf, 0);
        if (x < 0 || y < 0 || x >= n || y >= n) { rc = -1; xy[k] = -1; continue; }
        xy[k] = y * n + x;
        gw[xy[k]] += 1.0;
    }
    for (i64 k = 0; k < cnt; ++k) {
        if (xy[k] < 0) continue;
        double wgt = gw[xy[k]];
        vis[2 * k] = vis[2 * k] / wgt; /* complex / (wgt :+ 0) */
        vis[2 * k + 1] = vis[2 * k + 1] / wgt;
    }
    free(gw); free(xy);
    return rc;
}

/* ---------------------------------------------------------------------------
 * permute (+) scatter-add of (x,y,val) triples into an h x w grid, index = (row=y, col=x)
 * (the primitive behind src/Gridding.hs:99,197,244,317,377; golden old/BrokenNumbers.hs:85-91)
 * ------------------------------------------------------------------------- */
void orc_scatter_add(i64 h, i64 w, double *grid, i64 cnt, const i64 *x, const i64 *y,
                     const double *val) {
    (void)h;
    for (i64 k = 0; k < cnt; ++k) {
        double *g = grid + 2 * (y[k] * w + x[k]);
        g[0] += val[2 * k]; g[1] += val[2 * k + 1];
    }
}

/* grid (src/Gridding.hs:95-112): cell = n/2 + floor(0.5 + n*p); a[y,x] += v.  n = #rows
 * for BOTH coordinates (the reference takes n from the first dim only).  Out-of-range is UB
 * in the reference; we skip and count. */
i64 orc_grid_simple(i64 h, i64 w, double *grid, i64 cnt, const double *u, const double *v,
                    const double *vis) {
    const i64 halfn = fdiv(h, 2);
    const double nf = (double)h;
    i64 dropped = 0;
    for (i64 k = 0; k < cnt; ++k) {
        volatile double pu = nf * u[k], pv = nf * v[k];
        i64 x = halfn + (i64)floor(0.5 + pu);
        i64 y = halfn + (i64)floor(0.5 + pv);
        if (x < 0 || y < 0 || x >= w || y >= h) { ++dropped; continue; }
        double *g = grid + 2 * (y * w + x);
        g[0] += vis[2 * k]; g[1] += vis[2 * k + 1];
    }
    return dropped;
}

/* ---------------------------------------------------------------------------
 * convgrid / convgrid2 (src/Gridding.hs:153-197 / :199-244):
 *   grid[y - gh/2 + i, x - gw/2 + j] += vis * gcf[(wbin,) yf, xf, i, j]
 * out-of-range taps dropped (fixoutofbounds, :883-891: they become (0,0,+0) adds).
 * wbin == NULL -> convgrid (4-D table).  (x,xf,y,yf) from frac_coords (h,w) qpx p.
 * Complex multiply is the textbook (a+bi)(c+di) with separate mul/add (no FMA).
 * ------------------------------------------------------------------------- */
static inline void cmul_acc(double *g, double vr, double vi, double kr, double ki) {
    /* built with -ffp-contract=off: separate mul / add, never fused */
    const double rr = vr * kr, ii = vi * ki, ri = vr * ki, ir = vi * kr;
    g[0] += rr - ii;
    g[1] += ri + ir;
}

void orc_convgrid2(i64 nw, i64 qpx, i64 gh, i64 gw, const double *gcf, i64 h, i64 w,
                   double *grid, i64 cnt, const double *u, const double *v, const i64 *wbin,
                   const double *vis, int normalise) {
    (void)nw;
    const i64 halfgh = fdiv(gh, 2), halfgw = fdiv(gw, 2);
    for (i64 k = 0; k < cnt; ++k) {
        i64 x, xf, y, yf;
        orc_frac_coord(w, qpx, 1, u + k, &x, &xf, normalise);
        orc_frac_coord(h, qpx, 1, v + k, &y, &yf, normalise);
        const i64 wb = wbin ? wbin[k] : 0;
        const double *kern = gcf + 2 * ((((wb * qpx) + yf) * qpx + xf) * gh * gw);
        const double vr = vis[2 * k], vi = vis[2 * k + 1];
        for (i64 i = 0; i < gh; ++i) {
            i64 gy = y - halfgh + i;
            if (gy < 0 || gy >= h) continue;
            for (i64 j = 0; j < gw; ++j) {
                i64 gx = x - halfgw + j;
                if (gx < 0 || gx >= w) continue;
                const double *kk = kern + 2 * (i * gw + j);
                cmul_acc(grid + 2 * (gy * w + gx), vr, vi, kk[0], kk[1]);
            }
        }
    }
}

/* Adjoint of convgrid/convgrid2 (NOT in the reference -> parity unpinned):
 *   vis'[k] = sum_{i,j} conj(gcf[(wbin,)yf,xf,i,j]) * grid[y-gh/2+i, x-gw/2+j]      */
void orc_convdegrid2(i64 nw, i64 qpx, i64 gh, i64 gw, const double *gcf, i64 h, i64 w,
                     const double *grid, i64 cnt, const double *u, const double *v,
                     const i64 *wbin, double *vis_out, int normalise) {
    (void)nw;
    const i64 halfgh = fdiv(gh, 2), halfgw = fdiv(gw, 2);
    for (i64 k = 0; k < cnt; ++k) {
        i64 x, xf, y, yf;
        orc_frac_coord(w, qpx, 1, u + k, &x, &xf, normalise);
        orc_frac_coord(h, qpx, 1, v + k, &y, &yf, normalise);
        const i64 wb = wbin ? wbin[k] : 0;
        const double *kern = gcf + 2 * ((((wb * qpx) + yf) * qpx + xf) * gh * gw);
        double acc[2] = {0.0, 0.0};
        for (i64 i = 0; i < gh; ++i) {
            i64 gy = y - halfgh + i;
            if (gy < 0 || gy >= h) continue;
            for (i64 j = 0; j < gw; ++j) {
                i64 gx = x - halfgw + j;
                if (gx < 0 || gx >= w) continue;
                const double *kk = kern + 2 * (i * gw + j);
                const double *g = grid + 2 * (gy * w + gx);
                cmul_acc(acc, g[0], g[1], kk[0], -kk[1]);
            }
        }
        vis_out[2 * k] = acc[0]; vis_out[2 * k + 1] = acc[1];
    }
}

/* ---------------------------------------------------------------------------
 * convolve2d, direct form of the reference's FFT route (src/Gridding.hs:795-811 with
 * pad_mid :682-691, padder :863-877 (transposes: reads array[oldx,oldy], Q1), extract_mid
 * :694-707).  Net effect (re-verified in tests against the literal FFT route
 * restated in oracle/oracle.py):
 *   out[ty,tx] = sum_{ky,kx} a1T[ky,kx] * a2T[ty + n/2 - ky, tx + n/2 - kx],  aT = transpose
 * ------------------------------------------------------------------------- */
void orc_convolve2d(i64 n, const double *a1, const double *a2, double *out) {
    const i64 c = fdiv(n, 2);
    for (i64 ty = 0; ty < n; ++ty)
        for (i64 tx = 0; tx < n; ++tx) {
            double sr = 0.0, si = 0.0;
            for (i64 ky = 0; ky < n; ++ky) {
                i64 qy = ty + c - ky;
                if (qy < 0 || qy >= n) continue;
                for (i64 kx = 0; kx < n; ++kx) {
                    i64 qx = tx + c - kx;
                    if (qx < 0 || qx >= n) continue;
                    /* transposed reads: a1T[ky,kx] = a1[kx,ky] */
                    const double *p = a1 + 2 * (kx * n + ky);
                    const double *q = a2 + 2 * (qx * n + qy);
                    const double rr = p[0] * q[0], ii = p[1] * q[1], ri = p[0] * q[1], ir = p[1] * q[0];
                    sr += rr - ii; si += ri + ir;
                }
            }
            out[2 * (ty * n + tx)] = sr; out[2 * (ty * n + tx) + 1] = si;
        }
}

/* aw_kernel_fn2 (src/Gridding.hs:761-775): convolve2d (convolve2d a1 a2) (w[yf,xf]). */
void orc_aw_kernel(i64 qpx, i64 s, const double *wkern_plane /* [qpx,qpx,s,s] */, i64 yf, i64 xf,
                   const double *a1, const double *a2, double *out, double *scratch /* s*s*2 */) {
    orc_convolve2d(s, a1, a2, scratch);
    orc_convolve2d(s, scratch, wkern_plane + 2 * ((yf * qpx + xf) * s * s), out);
}

/* ---------------------------------------------------------------------------
 * convgrid3 / convgrid4 (src/Gridding.hs:246-317 / :318-396, processOne2 :379-396):
 *   AW_k = aw_kernel_fn2 yf xf wkerns[wbin] akerns[a1] akerns[a2]
 *   grid[y + i - gh/2, x + j - gw/2] += vis_k * conj(AW_k[i,j]), out-of-range dropped.
 * degrid != 0 computes the adjoint instead: vis'[k] = sum AW_k[i,j] * grid[...]
 * (conj of the conj'd kernel).  vis is input for gridding, output for degridding.
 * ------------------------------------------------------------------------- */
void orc_convgrid_aw(i64 nw, i64 qpx, i64 s, const double *wkerns, i64 nant, const double *akerns,
                     i64 h, i64 w, double *grid, i64 cnt, const double *u, const double *v,
                     const i64 *wbin, const i64 *a1, const i64 *a2, double *vis, int normalise,
                     int degrid) {
    (void)nw; (void)nant;
    const i64 half = fdiv(s, 2);
    double *aw = (double *)malloc((size_t)(4 * s * s) * sizeof(double));
    double *scratch = aw + 2 * s * s;
    for (i64 k = 0; k < cnt; ++k) {
        i64 x, xf, y, yf;
        orc_frac_coord(w, qpx, 1, u + k, &x, &xf, normalise);
        orc_frac_coord(h, qpx, 1, v + k, &y, &yf, normalise);
        orc_aw_kernel(qpx, s, wkerns + 2 * (wbin[k] * qpx * qpx * s * s), yf, xf,
                      akerns + 2 * (a1[k] * s * s), akerns + 2 * (a2[k] * s * s), aw, scratch);
        double acc[2] = {0.0, 0.0};
        const double vr = vis[2 * k], vi = vis[2 * k + 1];
        for (i64 i = 0; i < s; ++i) {
            i64 gy = y + i - half;
            if (gy < 0 || gy >= h) continue;
            for (i64 j = 0; j < s; ++j) {
                i64 gx = x + j - half;
                if (gx < 0 || gx >= w) continue;
                const double *kk = aw + 2 * (i * s + j);
                double *g = grid + 2 * (gy * w + gx);
                if (degrid) cmul_acc(acc, g[0], g[1], kk[0], kk[1]);
                else cmul_acc(g, vr, vi, kk[0], -kk[1]);
            }
        }
        if (degrid) { vis[2 * k] = acc[0]; vis[2 * k + 1] = acc[1]; }
    }
    free(aw);
}

/* ---------------------------------------------------------------------------
 * make_grid_hermitian (src/Gridding.hs:585-605):
 *  even n: out[y,x] = g[y,x] + (x==0||y==0 ? 0 : conj g[n-y,n-x])
 *  odd  n: out = g + conj(flipud(fliplr g))
 * ------------------------------------------------------------------------- */
void orc_make_grid_hermitian(i64 n, const double *g, double *out) {
    const int even = (n % 2 == 0);
    for (i64 y = 0; y < n; ++y)
        for (i64 x = 0; x < n; ++x) {
            double mr = 0.0, mi = 0.0;
            if (even) {
                if (!(x == 0 || y == 0)) {
                    const double *m = g + 2 * ((n - y) * n + (n - x));
                    mr = m[0]; mi = -m[1];
                }
            } else {
                const double *m = g + 2 * ((n - 1 - y) * n + (n - 1 - x));
                mr = m[0]; mi = -m[1];
            }
            out[2 * (y * n + x)] = g[2 * (y * n + x)] + mr;
            out[2 * (y * n + x) + 1] = g[2 * (y * n + x) + 1] + mi;
        }
}

/* ===========================================================================
 * CPU BASELINE (bench.py cpu_baseline / --impl reference): the same convgrid2 semantics,
 * parallelised over disjoint grid row bands so no atomics are needed (BASELINE.md section 4,
 * variant 1).  Binning is done once up front; the visibilities are listed per row band (in
 * visibility order) and the bands are claimed dynamically by the threads.  Per-cell summation
 * order is visibility order, as in orc_convgrid2, so the result is bit-identical to it.
 * ========================================================================= */
/* torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline wants every host core it may use */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_convgrid2_omp(i64 nw, i64 qpx, i64 gh, i64 gw, const double *gcf, i64 h, i64 w,
                       double *grid, i64 cnt, const double *u, const double *v, const i64 *wbin,
                       const double *vis, int normalise) {
    (void)nw;
    const i64 halfgh = fdiv(gh, 2), halfgw = fdiv(gw, 2);
    i64 *xs = (i64 *)malloc((size_t)(cnt > 0 ? cnt : 1) * 4 * sizeof(i64));
    i64 *xfs = xs + cnt, *ys = xfs + cnt, *yfs = ys + cnt;
    orc_frac_coord(w, qpx, cnt, u, xs, xfs, normalise);
    orc_frac_coord(h, qpx, cnt, v, ys, yfs, normalise);
    /* row bands (many more than threads, claimed dynamically: the uv coverage is core-dominated); every band
     * gets the list of visibilities whose footprint rows intersect it, in visibility order */
    const i64 nb = (i64)orc_num_threads() * 16 < h ? (i64)orc_num_threads() * 16 : (h > 0 ? h : 1);
    i64 *start = (i64 *)calloc((size_t)nb + 1, sizeof(i64));
#define BAND_OF(r) ((((r) + 1) * nb - 1) / h) /* largest b with h*b/nb <= r */
    for (i64 k = 0; k < cnt; ++k) {
        i64 y0 = ys[k] - halfgh, y1 = y0 + gh - 1;
        if (y1 < 0 || y0 >= h) continue;
        if (y0 < 0) y0 = 0;
        if (y1 >= h) y1 = h - 1;
        for (i64 b = BAND_OF(y0); b <= BAND_OF(y1); ++b) start[b + 1]++;
    }
    for (i64 b = 0; b < nb; ++b) start[b + 1] += start[b];
    i64 *list = (i64 *)malloc((size_t)(start[nb] > 0 ? start[nb] : 1) * sizeof(i64));
    i64 *fill = (i64 *)malloc((size_t)nb * sizeof(i64));
    memcpy(fill, start, (size_t)nb * sizeof(i64));
    for (i64 k = 0; k < cnt; ++k) {
        i64 y0 = ys[k] - halfgh, y1 = y0 + gh - 1;
        if (y1 < 0 || y0 >= h) continue;
        if (y0 < 0) y0 = 0;
        if (y1 >= h) y1 = h - 1;
        for (i64 b = BAND_OF(y0); b <= BAND_OF(y1); ++b) list[fill[b]++] = k;
    }
#pragma omp parallel for schedule(dynamic, 1)
    for (i64 b = 0; b < nb; ++b) {
        const i64 r0 = h * b / nb, r1 = h * (b + 1) / nb;
        for (i64 q = start[b]; q < start[b + 1]; ++q) {
            const i64 k = list[q];
            const i64 y0 = ys[k] - halfgh;
            const i64 wb = wbin ? wbin[k] : 0;
            const double *kern = gcf + 2 * ((((wb * qpx) + yfs[k]) * qpx + xfs[k]) * gh * gw);
            const double vr = vis[2 * k], vi = vis[2 * k + 1];
            const i64 ilo = (r0 - y0 > 0) ? r0 - y0 : 0;
            const i64 ihi = (r1 - y0 < gh) ? r1 - y0 : gh;
            for (i64 i = ilo; i < ihi; ++i) {
                const i64 gy = y0 + i;
                for (i64 j = 0; j < gw; ++j) {
                    const i64 gx = xs[k] - halfgw + j;
                    if (gx < 0 || gx >= w) continue;
                    const double *kk = kern + 2 * (i * gw + j);
                    cmul_acc(grid + 2 * (gy * w + gx), vr, vi, kk[0], kk[1]);
                }
            }
        }
    }
#undef BAND_OF
    free(fill); free(list); free(start); free(xs);
}

void orc_convdegrid2_omp(i64 nw, i64 qpx, i64 gh, i64 gw, const double *gcf, i64 h, i64 w,
                         const double *grid, i64 cnt, const double *u, const double *v,
                         const i64 *wbin, double *vis_out, int normalise) {
#pragma omp parallel
    {
#ifdef _OPENMP
        const int nt = omp_get_num_threads(), tid = omp_get_thread_num();
#else
        const int nt = 1, tid = 0;
#endif
        const i64 k0 = cnt * tid / nt, k1 = cnt * (tid + 1) / nt;
        orc_convdegrid2(nw, qpx, gh, gw, gcf, h, w, grid, k1 - k0, u + k0, v + k0,
                        wbin ? wbin + k0 : NULL, vis_out + 2 * k0, normalise);
    }
}
