// api.cu -- the host-pointer C ABI (include/skagrid.h): what the reference's Haskell layer binds with
// `foreign import ccall`.  Every function copies its inputs to the device, runs the sm_100a kernels and
// copies the results back; there is no CPU implementation of any of them in this library.
//
// The table gridders / degridders stream the visibilities in chunks: the H2D copy of chunk c+1 runs on the
// copy stream while chunk c is binned, bucketed and gridded on the compute stream (pinned caller buffers
// make the copy truly asynchronous; pageable ones still work).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

// visibilities per pipelined chunk of the table gridders (SKAGRID_VIS_CHUNK_LOG2 overrides the exponent: tuning experiments)
// Gridding from host coordinates is bound by the upload (40 B per visibility): small chunks shorten the pipeline fill and
// the per-chunk kernels hide behind the copies anyway.  Degridding, and gridding at resident coordinates, are bound by
// the kernels, which lose efficiency on small batches (every chunk zeroes / stages every non-empty uv tile again).
// B200, 1e8 visibilities, e2e step of bench.py: 2^22 / 2^23 / 2^24 / 2^25 -> gridding call 75.5 / 77.0 / 78.7 / 82.8 ms,
// degridding call 60.8 / 45.3 / 40.1 / 42.0 ms.
static i64 vis_chunk(bool kernel_bound) {
    static const int forced = getenv("SKAGRID_VIS_CHUNK_LOG2") ? atoi(getenv("SKAGRID_VIS_CHUNK_LOG2")) : 0;
    if (forced >= 10 && forced <= 30) return (i64)1 << forced;
    return (i64)1 << (kernel_bound ? 24 : 22);
}
static const i64 RES_MAX = (i64)1 << 28;   // at most this many coordinates are kept resident between calls (24 B each)
static const i64 AW_CHUNK = (i64)1 << 20;   // most visibilities per chunk of the AW path (aw_core_dev)

static inline int up(skagrid_ctx *ctx, const char *name, const void *host, size_t bytes, void **dev) { return sk_api_up(ctx, name, host, bytes, dev); }
static inline int check_flags(skagrid_ctx *ctx, const char *what) { return sk_api_check_flags(ctx, what); }
static inline int plan_acquire(skagrid_ctx *ctx, const skagrid_geom *geom, i64 capacity, int slice_override, skagrid_plan **out) {
    return sk_api_plan_acquire(ctx, geom, capacity, slice_override, out);
}

struct Timer {
    skagrid_ctx *ctx;
    explicit Timer(skagrid_ctx *c) : ctx(c) { cudaEventRecord(ctx->ev0, ctx->stream); }
    int finish() {
        SK_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        SK_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
        float ms = 0.f;
        SK_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        ctx->last_ms = ms;
        return SKAGRID_OK;
    }
};

int sk_api_enter(skagrid_ctx *ctx) {
    if (!ctx) return SKAGRID_EINVAL;
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->err.clear();
    return SKAGRID_OK;
}

// scratch + H2D on the compute stream
int sk_api_up(skagrid_ctx *ctx, const char *name, const void *host, size_t bytes, void **dev) {
    SK_TRY(sk_scratch(ctx, name, bytes ? bytes : 16, dev));
    if (bytes) SK_CUDA(ctx, cudaMemcpyAsync(*dev, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return SKAGRID_OK;
}
static int down(skagrid_ctx *ctx, void *host, const void *dev, size_t bytes) {
    if (bytes) SK_CUDA(ctx, cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return SKAGRID_OK;
}
int sk_api_check_flags(skagrid_ctx *ctx, const char *what) {
    uint32_t f = 0;
    SK_TRY(sk_take_flags(ctx, ctx->stream, &f));
    if (f & 1u) return sk_fail(ctx, SKAGRID_ERANGE, "%s: a w-plane, oversampling or antenna index is out of range", what);
    if (f & 2u) return sk_fail(ctx, SKAGRID_ERANGE, "%s: a visibility falls outside the weight grid", what);
    return SKAGRID_OK;
}

// The host-pointer functions keep ONE plan alive in the context (cudaMalloc/cudaFree of its ~0.5 GB of buckets and
// records per call would cost more than the gridding of a small batch); it is reused when the geometry matches.
int sk_api_plan_acquire(skagrid_ctx *ctx, const skagrid_geom *geom, i64 capacity, int slice_override, skagrid_plan **out) {
    skagrid_plan *p = ctx->cached_plan;
    if (p) {
        const Geom &g = p->g;
        const bool same = g.height == geom->height && g.width == geom->width && g.row0 == geom->row0 && g.row1 == geom->row1 &&
                          g.nw == geom->nw && g.qpx == geom->qpx && g.gh == geom->gh && g.gw == geom->gw &&
                          p->slice_override == slice_override && p->capacity >= capacity;
        if (same) { *out = p; return SKAGRID_OK; }
        sk_plan_free(p);
        ctx->cached_plan = nullptr;
    }
    SK_TRY(sk_plan_alloc(ctx, geom, capacity, slice_override, &p));
    ctx->cached_plan = p;
    *out = p;
    return SKAGRID_OK;
}
static void plan_release(skagrid_ctx *, skagrid_plan *) {}  // stays cached; freed by skagrid_destroy

// Context-resident grid: the host-pointer gridders / degridders / grid_to_image accept grid == NULL, meaning "the grid
// the previous call on this context left on the device".  A chain conv_imaging2 -> grid_to_image -> convdegrid2 then
// moves only visibilities over PCIe, like the single fused Accelerate program of the reference does.
static int grid_in(skagrid_ctx *ctx, const double *host, i64 h, i64 w, void **dgrid, const char *what) {
    const size_t bytes = (size_t)(h * w) * 16;
    if (host) {
        SK_TRY(up(ctx, "grid", host, bytes, dgrid));
    } else {
        if (ctx->resident_h != h || ctx->resident_w != w)
            return sk_fail(ctx, SKAGRID_EINVAL, "%s: grid is NULL but the context holds no resident %lld x %lld grid", what, h, w);
        SK_TRY(sk_scratch(ctx, "grid", bytes, dgrid));
    }
    ctx->resident_h = h; ctx->resident_w = w;
    return SKAGRID_OK;
}
static int grid_fresh(skagrid_ctx *ctx, i64 h, i64 w, void **dgrid) {
    const size_t bytes = (size_t)(h * w) * 16;
    ctx->resident_h = ctx->resident_w = 0;
    SK_TRY(sk_scratch(ctx, "grid", bytes, dgrid));
    SK_CUDA(ctx, cudaMemsetAsync(*dgrid, 0, bytes, ctx->stream));
    ctx->resident_h = h; ctx->resident_w = w;
    return SKAGRID_OK;
}

#define NEED(ctx, cond, what) \
    do { if (!(cond)) return sk_fail((ctx), SKAGRID_EINVAL, "%s", (what)); } while (0)

// ------------------------------------------------------------------------------------------ binning
extern "C" int skagrid_frac_coord(skagrid_ctx *ctx, int64_t n, int64_t qpx, int64_t count, const double *p, int64_t *fl, int64_t *frac,
                                  int flags) {
    SK_TRY(sk_api_enter(ctx));
    NEED(ctx, n > 0 && qpx > 0 && count >= 0, "frac_coord: n, qpx must be positive");
    if (count == 0) return SKAGRID_OK;
    NEED(ctx, p && fl && frac, "frac_coord: NULL pointer");
    Timer t(ctx);
    void *dp, *dfl, *dfr;
    SK_TRY(up(ctx, "fc_p", p, (size_t)count * 8, &dp));
    SK_TRY(sk_scratch(ctx, "fc_fl", (size_t)count * 8, &dfl));
    SK_TRY(sk_scratch(ctx, "fc_fr", (size_t)count * 8, &dfr));
    SK_TRY(sk_frac_coord_dev(ctx, n, qpx, count, (double *)dp, (i64 *)dfl, (i64 *)dfr, flags & SKAGRID_FRAC_NORMALISE, ctx->stream));
    SK_TRY(down(ctx, fl, dfl, (size_t)count * 8));
    SK_TRY(down(ctx, frac, dfr, (size_t)count * 8));
    return t.finish();
}

extern "C" int skagrid_frac_coords(skagrid_ctx *ctx, int64_t height, int64_t width, int64_t qpx, int64_t count, const double *u,
                                   const double *v, int64_t *x, int64_t *xf, int64_t *y, int64_t *yf, int flags) {
    // src/Gridding.hs:142-151: x,xf from u with WIDTH; y,yf from v with HEIGHT
    SK_TRY(skagrid_frac_coord(ctx, width, qpx, count, u, x, xf, flags));
    return skagrid_frac_coord(ctx, height, qpx, count, v, y, yf, flags);
}

extern "C" int skagrid_find_closest(skagrid_ctx *ctx, int64_t nw, const double *wbins, int64_t count, const double *w, int64_t *out) {
    SK_TRY(sk_api_enter(ctx));
    NEED(ctx, nw > 0 && count >= 0, "find_closest: empty wbins");
    if (count == 0) return SKAGRID_OK;
    NEED(ctx, wbins && w && out, "find_closest: NULL pointer");
    Timer t(ctx);
    void *dws, *dw, *dout;
    SK_TRY(up(ctx, "fcl_ws", wbins, (size_t)nw * 8, &dws));
    SK_TRY(up(ctx, "fcl_w", w, (size_t)count * 8, &dw));
    SK_TRY(sk_scratch(ctx, "fcl_out", (size_t)count * 8, &dout));
    SK_TRY(sk_find_closest_dev(ctx, nw, (double *)dws, count, (double *)dw, (i64 *)dout, ctx->stream));
    SK_TRY(down(ctx, out, dout, (size_t)count * 8));
    return t.finish();
}

// ------------------------------------------------------------------------------------------ pre-steps
static int up_uvw(skagrid_ctx *ctx, i64 count, const double *u, const double *v, const double *w, double **du, double **dv, double **dw) {
    SK_TRY(up(ctx, "pre_u", u, (size_t)count * 8, (void **)du));
    SK_TRY(up(ctx, "pre_v", v, (size_t)count * 8, (void **)dv));
    if (w) SK_TRY(up(ctx, "pre_w", w, (size_t)count * 8, (void **)dw));
    return SKAGRID_OK;
}

extern "C" int skagrid_uvw_lambda(skagrid_ctx *ctx, double freq, int64_t count, double *u, double *v, double *w) {
    SK_TRY(sk_api_enter(ctx));
    if (count <= 0) return SKAGRID_OK;
    NEED(ctx, u && v && w, "uvw_lambda: NULL pointer");
    Timer t(ctx);
    double *du, *dv, *dw;
    SK_TRY(up_uvw(ctx, count, u, v, w, &du, &dv, &dw));
    const double a = freq / 299792458.0;  // formed on the host in double, src/ImageDataset.hs:184
    SK_TRY(sk_scale3_dev(ctx, count, du, dv, dw, a, 0, ctx->stream));
    SK_TRY(down(ctx, u, du, (size_t)count * 8));
    SK_TRY(down(ctx, v, dv, (size_t)count * 8));
    SK_TRY(down(ctx, w, dw, (size_t)count * 8));
    return t.finish();
}

extern "C" int skagrid_mirror_uvw(skagrid_ctx *ctx, int64_t count, double *u, double *v, double *w, double *vis) {
    SK_TRY(sk_api_enter(ctx));
    if (count <= 0) return SKAGRID_OK;
    NEED(ctx, u && v && w && vis, "mirror_uvw: NULL pointer");
    Timer t(ctx);
    double *du, *dv, *dw;
    void *dvis;
    SK_TRY(up_uvw(ctx, count, u, v, w, &du, &dv, &dw));
    SK_TRY(up(ctx, "pre_vis", vis, (size_t)count * 16, &dvis));
    SK_TRY(sk_mirror_dev(ctx, count, du, dv, dw, (double *)dvis, ctx->stream));
    SK_TRY(down(ctx, u, du, (size_t)count * 8));
    SK_TRY(down(ctx, v, dv, (size_t)count * 8));
    SK_TRY(down(ctx, w, dw, (size_t)count * 8));
    SK_TRY(down(ctx, vis, dvis, (size_t)count * 16));
    return t.finish();
}

// N = P.round (theta * lam) (src/Gridding.hs:87, :118, :416, :466, :571): Haskell's round is half-to-even, which is what
// nearbyint does in the default rounding mode.  The one definition every layer (C++ mirror, Python, Haskell shim) uses.
extern "C" int64_t skagrid_grid_side(double theta, int64_t lam) { return (int64_t)nearbyint(theta * (double)lam); }
static i64 grid_side(double theta, i64 lam) { return (i64)skagrid_grid_side(theta, lam); }

extern "C" int skagrid_doweight(skagrid_ctx *ctx, double theta, int64_t lam, int64_t count, const double *u, const double *v, double *vis) {
    SK_TRY(sk_api_enter(ctx));
    if (count <= 0) return SKAGRID_OK;
    NEED(ctx, u && v && vis, "doweight: NULL pointer");
    const i64 n = grid_side(theta, lam);
    NEED(ctx, n > 0, "doweight: round(theta*lam) must be positive");
    Timer t(ctx);
    double *du, *dv, *dw = nullptr;
    void *dvis;
    SK_TRY(up_uvw(ctx, count, u, v, nullptr, &du, &dv, &dw));
    SK_TRY(up(ctx, "pre_vis", vis, (size_t)count * 16, &dvis));
    SK_TRY(sk_doweight_dev(ctx, n, (double)lam, count, du, dv, (double *)dvis, ctx->d_flags, ctx->stream));
    SK_TRY(down(ctx, vis, dvis, (size_t)count * 16));
    SK_TRY(t.finish());
    return check_flags(ctx, "doweight");
}

// ------------------------------------------------------------------------------------------ table gridders
// Device-side core: streams `count` host visibilities through a plan in double-buffered chunks.
//   degrid == 0: d_grid[row0:row1] += sum vis_k * d_table[slice_k]      (vis is the input)
//   degrid != 0: vis_out[k] = sum conj(d_table[slice_k]) * d_grid[...]  (vis_out is the output, host)
//
// sk_api_stream_enqueue only enqueues (events order the three streams; no host synchronisation), so one host thread can
// keep several devices busy (mgpu.cu); sk_api_stream_wait drains the context's streams.
int sk_api_stream_enqueue(skagrid_ctx *ctx, const skagrid_geom *geom, const double *d_table, double *d_grid, i64 count, const double *u,
                          const double *v, const int64_t *wbin, const double *vis, double *vis_out, int degrid, double lam, int want_wbin) {
    if (count <= 0) return SKAGRID_OK;
    const i64 chunk = std::min<i64>(count, vis_chunk(degrid || (!u && !v)));
    // Resident coordinates: the uploaded (u, v, wbin) stay on the device in per-context arrays, so the next call may pass
    // u == v == wbin == NULL ("the coordinates of the previous call", include/skagrid.h) and skip 24 of its 40 bytes per
    // visibility of PCIe traffic -- an imaging major cycle grids and degrids the same uvw.
    const bool reuse = !u && !v;
    bool resident = false;
    double *ru = nullptr, *rv = nullptr;
    i64 *rwb = nullptr;
    if (reuse) {
        if (ctx->res_count != count)
            return sk_fail(ctx, SKAGRID_EINVAL, "coordinates are NULL but the context holds %lld resident ones, not %lld", (long long)ctx->res_count,
                           (long long)count);
        if (want_wbin && !ctx->res_has_wbin) return sk_fail(ctx, SKAGRID_EINVAL, "coordinates are NULL but the resident ones carry no w-plane indices");
        if (lam > 0.0) return sk_fail(ctx, SKAGRID_EINVAL, "resident coordinates are already divided by lam");
        resident = true;
    } else {
        ctx->res_count = 0;
        resident = count <= RES_MAX;
    }
    if (resident) {
        int ra = sk_scratch(ctx, "res_u", (size_t)count * 8, (void **)&ru);
        if (!ra) ra = sk_scratch(ctx, "res_v", (size_t)count * 8, (void **)&rv);
        if (!ra && want_wbin) ra = sk_scratch(ctx, "res_wb", (size_t)count * 8, (void **)&rwb);
        if (ra) {  // no room for the resident copy: plain double-buffered chunks
            if (reuse) return ra;
            resident = false;
            ctx->err.clear();
        }
    }
    skagrid_plan *plan = nullptr;
    SK_TRY(plan_acquire(ctx, geom, chunk, 0, &plan));
    double *du[2] = {nullptr, nullptr}, *dv[2] = {nullptr, nullptr}, *dvis[2];
    i64 *dwb[2] = {nullptr, nullptr};
    int rc = SKAGRID_OK;
    for (int b = 0; b < 2 && !rc; ++b) {
        const char *nu = b ? "st_u1" : "st_u0", *nv = b ? "st_v1" : "st_v0", *nwb = b ? "st_w1" : "st_w0", *nvis = b ? "st_vis1" : "st_vis0";
        if (!resident) {
            rc = sk_scratch(ctx, nu, (size_t)chunk * 8, (void **)&du[b]);
            if (!rc) rc = sk_scratch(ctx, nv, (size_t)chunk * 8, (void **)&dv[b]);
            if (!rc && want_wbin) rc = sk_scratch(ctx, nwb, (size_t)chunk * 8, (void **)&dwb[b]);
        }
        if (!rc) rc = sk_scratch(ctx, nvis, (size_t)chunk * 16, (void **)&dvis[b]);
    }
    if (rc) { plan_release(ctx, plan); return rc; }
    cudaError_t e = cudaSuccess;
    // the copy stream must not overwrite scratch that earlier work on the compute stream still uses
    e = cudaEventRecord(ctx->ev_done[0], ctx->stream);
    if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_done[1], ctx->stream);
    if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_d2h[0], ctx->d2h_stream);
    if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_d2h[1], ctx->d2h_stream);
    i64 ci = 0;
    for (i64 off = 0; off < count && e == cudaSuccess && !rc; off += chunk, ++ci) {
        const int b = (int)(ci & 1);
        const i64 n = std::min<i64>(chunk, count - off);
        double *cu = resident ? ru + off : du[b], *cv = resident ? rv + off : dv[b];
        i64 *cwb = !want_wbin ? nullptr : (resident ? rwb + off : dwb[b]);
        e = cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[b], 0);
        if (!reuse) {
            if (e == cudaSuccess) e = cudaMemcpyAsync(cu, u + off, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->copy_stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(cv, v + off, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->copy_stream);
            if (e == cudaSuccess && cwb) e = cudaMemcpyAsync(cwb, wbin + off, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->copy_stream);
        }
        if (e == cudaSuccess && !degrid) e = cudaMemcpyAsync(dvis[b], vis + 2 * off, (size_t)n * 16, cudaMemcpyHostToDevice, ctx->copy_stream);
        if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_copy[b], ctx->copy_stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[b], 0);
        if (e != cudaSuccess) break;
        if (lam > 0.0) {  // div3 (src/Gridding.hs:838-839) on u and v; w does not enter the table gridders
            void *dscr;
            rc = sk_scratch(ctx, "st_wscr", (size_t)chunk * 8, &dscr);
            if (!rc) { cudaMemsetAsync(dscr, 0, (size_t)n * 8, ctx->stream); rc = sk_scale3_dev(ctx, n, cu, cv, (double *)dscr, lam, 1, ctx->stream); }
            if (rc) break;
        }
        rc = sk_plan_fill(ctx, plan, n, cu, cv, cwb, degrid ? nullptr : dvis[b], ctx->stream);
        if (!rc) {
            if (degrid) {
                // dvis[b] still holds the results of chunk c-2 until their D2H copy (on the third stream) is done
                e = cudaStreamWaitEvent(ctx->stream, ctx->ev_d2h[b], 0);
                if (e == cudaSuccess) rc = skagrid_dev_degrid(ctx, plan, d_table, d_grid, dvis[b], ctx->stream);
                if (!rc && e == cudaSuccess) e = cudaEventRecord(ctx->ev_k[b], ctx->stream);
                if (!rc && e == cudaSuccess) e = cudaStreamWaitEvent(ctx->d2h_stream, ctx->ev_k[b], 0);
                if (!rc && e == cudaSuccess) e = cudaMemcpyAsync(vis_out + 2 * off, dvis[b], (size_t)n * 16, cudaMemcpyDeviceToHost, ctx->d2h_stream);
                if (!rc && e == cudaSuccess) e = cudaEventRecord(ctx->ev_d2h[b], ctx->d2h_stream);
            } else {
                rc = skagrid_dev_grid(ctx, plan, d_table, d_grid, 0, ctx->stream);
            }
        }
        if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_done[b], ctx->stream);
    }
    // later work on the compute stream (a reduction, a D2H of the grid) must also see the degridder's result copies
    if (e == cudaSuccess && degrid) {
        e = cudaStreamWaitEvent(ctx->stream, ctx->ev_d2h[0], 0);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ctx->ev_d2h[1], 0);
    }
    plan_release(ctx, plan);
    if (rc) return rc;
    if (e != cudaSuccess) return sk_fail(ctx, SKAGRID_ECUDA, "table gridder: %s", cudaGetErrorString(e));
    if (resident && !reuse) { ctx->res_count = count; ctx->res_has_wbin = want_wbin ? 1 : 0; }
    return SKAGRID_OK;
}

int sk_api_stream_wait(skagrid_ctx *ctx) {
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    const cudaError_t e1 = cudaStreamSynchronize(ctx->copy_stream);
    const cudaError_t e2 = cudaStreamSynchronize(ctx->d2h_stream);
    if (e == cudaSuccess) e = e1;
    if (e == cudaSuccess) e = e2;
    if (e != cudaSuccess) return sk_fail(ctx, SKAGRID_ECUDA, "table gridder: %s", cudaGetErrorString(e));
    return SKAGRID_OK;
}

static int stream_table(skagrid_ctx *ctx, const skagrid_geom *geom, const double *d_table, double *d_grid, i64 count, const double *u,
                        const double *v, const int64_t *wbin, const double *vis, double *vis_out, int degrid, int want_wbin, double lam = 0.0) {
    const int rc = sk_api_stream_enqueue(ctx, geom, d_table, d_grid, count, u, v, wbin, vis, vis_out, degrid, lam, want_wbin);
    const int rw = sk_api_stream_wait(ctx);  // always drain, also after a failed enqueue
    return rc ? rc : rw;
}

static int check_table_args(skagrid_ctx *ctx, i64 nw, i64 qpx, i64 gh, i64 gw, i64 height, i64 width, i64 count) {
    NEED(ctx, nw > 0 && qpx > 0 && gh > 0 && gw > 0, "kernel table: non-positive dimension");
    NEED(ctx, height > 0 && width > 0 && height <= 65536 && width <= 65536, "grid: size outside [1,65536]");
    NEED(ctx, count >= 0, "negative visibility count");
    return SKAGRID_OK;
}

static int table_host(skagrid_ctx *ctx, i64 nw, i64 qpx, i64 gh, i64 gw, const double *gcf, i64 height, i64 width, double *grid, i64 count,
                      const double *u, const double *v, const int64_t *wbin, const double *vis, double *vis_out, int degrid, int want_wbin,
                      const char *what) {
    SK_TRY(sk_api_enter(ctx));
    SK_TRY(check_table_args(ctx, nw, qpx, gh, gw, height, width, count));
    NEED(ctx, gcf, "NULL kernel table");
    if (count > 0) {
        // u == v == wbin == NULL: the coordinates the previous call left on the device (sk_api_stream_enqueue)
        const bool given = u && v && (wbin || !want_wbin), resident = !u && !v && !wbin;
        NEED(ctx, (given || resident) && (degrid ? vis_out != nullptr : vis != nullptr), "NULL visibility array");
    }
    Timer t(ctx);
    void *dtab, *dgrid;
    const size_t tab_bytes = (size_t)(nw * qpx * qpx * gh * gw) * 16, grid_bytes = (size_t)(height * width) * 16;
    SK_TRY(up(ctx, "tab", gcf, tab_bytes, &dtab));
    SK_TRY(grid_in(ctx, grid, height, width, &dgrid, what));
    skagrid_geom geom = {height, width, 0, height, nw, qpx, gh, gw};
    SK_TRY(stream_table(ctx, &geom, (double *)dtab, (double *)dgrid, count, u, v, wbin, vis, vis_out, degrid, want_wbin));
    if (!degrid && grid) SK_TRY(down(ctx, grid, dgrid, grid_bytes));
    SK_TRY(t.finish());
    return check_flags(ctx, what);
}

extern "C" int skagrid_convgrid(skagrid_ctx *ctx, int64_t qpx, int64_t gh, int64_t gw, const double *gcf, int64_t height, int64_t width,
                                double *grid, int64_t count, const double *u, const double *v, const double *vis) {
    return table_host(ctx, 1, qpx, gh, gw, gcf, height, width, grid, count, u, v, nullptr, vis, nullptr, 0, 0, "convgrid");
}

extern "C" int skagrid_convgrid2(skagrid_ctx *ctx, int64_t nw, int64_t qpx, int64_t gh, int64_t gw, const double *gcf, int64_t height,
                                 int64_t width, double *grid, int64_t count, const double *u, const double *v, const int64_t *wbin,
                                 const double *vis) {
    return table_host(ctx, nw, qpx, gh, gw, gcf, height, width, grid, count, u, v, wbin, vis, nullptr, 0, 1, "convgrid2");
}

extern "C" int skagrid_convdegrid(skagrid_ctx *ctx, int64_t qpx, int64_t gh, int64_t gw, const double *gcf, int64_t height, int64_t width,
                                  const double *grid, int64_t count, const double *u, const double *v, double *vis_out) {
    return table_host(ctx, 1, qpx, gh, gw, gcf, height, width, const_cast<double *>(grid), count, u, v, nullptr, nullptr, vis_out, 1, 0, "convdegrid");
}

extern "C" int skagrid_convdegrid2(skagrid_ctx *ctx, int64_t nw, int64_t qpx, int64_t gh, int64_t gw, const double *gcf, int64_t height,
                                   int64_t width, const double *grid, int64_t count, const double *u, const double *v, const int64_t *wbin,
                                   double *vis_out) {
    return table_host(ctx, nw, qpx, gh, gw, gcf, height, width, const_cast<double *>(grid), count, u, v, wbin, nullptr, vis_out, 1, 1, "convdegrid2");
}

extern "C" int skagrid_grid(skagrid_ctx *ctx, int64_t height, int64_t width, double *grid, int64_t count, const double *u, const double *v,
                            const double *vis) {
    SK_TRY(sk_api_enter(ctx));
    NEED(ctx, height > 0 && width > 0 && grid && count >= 0, "grid: bad size or NULL grid");
    if (count == 0) return SKAGRID_OK;
    NEED(ctx, u && v && vis, "grid: NULL visibility array");
    Timer t(ctx);
    double *du, *dv, *dw = nullptr;
    void *dvis, *dgrid;
    SK_TRY(up_uvw(ctx, count, u, v, nullptr, &du, &dv, &dw));
    SK_TRY(up(ctx, "pre_vis", vis, (size_t)count * 16, &dvis));
    SK_TRY((ctx->resident_h = ctx->resident_w = 0, up(ctx, "grid", grid, (size_t)(height * width) * 16, &dgrid)));  // overwrites the resident grid
    SK_TRY(sk_grid_simple_dev(ctx, height, width, (double *)dgrid, count, du, dv, (double *)dvis, ctx->stream));
    SK_TRY(down(ctx, grid, dgrid, (size_t)(height * width) * 16));
    return t.finish();
}

// ------------------------------------------------------------------------------------------ AW path
// Device core of convgrid3/convgrid4 (src/Gridding.hs:246-396) and its adjoint.  All pointers are device
// pointers; u,v are p coordinates.  Per chunk: bin -> per-visibility AW kernels (conjugated, :391) -> plan
// with slice_k = k -> tiled gridder / degridder.
static int aw_core_dev(skagrid_ctx *ctx, i64 nw, i64 qpx, i64 s, const double *d_wk, i64 nant, const double *d_ak, i64 height, i64 width,
                       double *d_grid, i64 count, const double *du, const double *dv, const i64 *dwb, const i64 *da1, const i64 *da2,
                       double *dvis, int degrid) {
    if (count <= 0) return SKAGRID_OK;
    // one S x S kernel per visibility: at most 4 GB of them per chunk (2^20 visibilities at S = 15, 67 000 at S = 63).  Round 1
    // used 0.5 GB; the gridder's cost per chunk is mostly per touched tile (zero + flush of the subgrid), not per visibility,
    // at AW densities -- 1.0 ms per 131072 visibilities (profiles/r02_launches_aw_1e6_summary.txt) -- so fewer, larger chunks.
    i64 chunk = std::min<i64>(count, std::max<i64>(4096, std::min<i64>(AW_CHUNK, ((i64)4096 << 20) / (s * s * 16))));
    if (const char *e = getenv("SKAGRID_AW_CHUNK")) {  // tests: force several chunks on a small input
        const i64 forced = atoll(e);
        if (forced > 0) chunk = std::min<i64>(count, forced);
    }
    skagrid_geom geom = {height, width, 0, height, 1, qpx, s, s};  // slice_override: the table has one slice per visibility
    skagrid_plan *plan = nullptr;
    SK_TRY(plan_acquire(ctx, &geom, chunk, 1, &plan));
    void *dx, *dxf, *dy, *dyf, *dk;
    int rc = sk_scratch(ctx, "aw_x", (size_t)chunk * 8, &dx);
    if (!rc) rc = sk_scratch(ctx, "aw_xf", (size_t)chunk * 8, &dxf);
    if (!rc) rc = sk_scratch(ctx, "aw_y", (size_t)chunk * 8, &dy);
    if (!rc) rc = sk_scratch(ctx, "aw_yf", (size_t)chunk * 8, &dyf);
    if (!rc) rc = sk_scratch(ctx, "aw_kern", (size_t)(chunk * s * s) * 16, &dk);
    for (i64 off = 0; off < count && !rc; off += chunk) {
        const i64 n = std::min<i64>(chunk, count - off);
        rc = sk_frac_coord_dev(ctx, width, qpx, n, du + off, (i64 *)dx, (i64 *)dxf, 1, ctx->stream);
        if (!rc) rc = sk_frac_coord_dev(ctx, height, qpx, n, dv + off, (i64 *)dy, (i64 *)dyf, 1, ctx->stream);
        if (!rc) rc = sk_aw_kernels_dev(ctx, nw, qpx, s, d_wk, nant, d_ak, n, dwb + off, (i64 *)dyf, (i64 *)dxf, da1 + off, da2 + off,
                                        (double *)dk, 1, ctx->d_flags, ctx->stream);
        if (!rc) rc = sk_plan_fill(ctx, plan, n, du + off, dv + off, nullptr, degrid ? nullptr : dvis + 2 * off, ctx->stream);
        if (!rc) {
            if (degrid) rc = skagrid_dev_degrid(ctx, plan, (double *)dk, d_grid, dvis + 2 * off, ctx->stream);
            else rc = skagrid_dev_grid(ctx, plan, (double *)dk, d_grid, 0, ctx->stream);
        }
    }
    cudaStreamSynchronize(ctx->stream);
    plan_release(ctx, plan);
    return rc;
}

static int aw_host(skagrid_ctx *ctx, i64 nw, i64 qpx, i64 s, const double *wkerns, i64 nant, const double *akerns, i64 height, i64 width,
                   double *grid, i64 count, const double *u, const double *v, const int64_t *wbin, const int64_t *a1, const int64_t *a2,
                   const double *vis, double *vis_out, int degrid, const char *what) {
    SK_TRY(sk_api_enter(ctx));
    SK_TRY(check_table_args(ctx, nw, qpx, s, s, height, width, count));
    NEED(ctx, nant > 0 && wkerns && akerns && grid, "NULL kernels / grid or nant <= 0");
    NEED(ctx, s <= 63, "AW path: support above 63 is not supported");
    if (count > 0) NEED(ctx, u && v && wbin && a1 && a2 && (degrid ? vis_out != nullptr : vis != nullptr), "NULL visibility array");
    Timer t(ctx);
    void *dwk, *dak, *dgrid, *du, *dv, *dwb, *da1, *da2, *dvis;
    SK_TRY(up(ctx, "aw_wk", wkerns, (size_t)(nw * qpx * qpx * s * s) * 16, &dwk));
    SK_TRY(up(ctx, "aw_ak", akerns, (size_t)(nant * s * s) * 16, &dak));
    SK_TRY((ctx->resident_h = ctx->resident_w = 0, up(ctx, "grid", grid, (size_t)(height * width) * 16, &dgrid)));  // overwrites the resident grid
    SK_TRY(up(ctx, "pre_u", u, (size_t)count * 8, &du));
    SK_TRY(up(ctx, "pre_v", v, (size_t)count * 8, &dv));
    SK_TRY(up(ctx, "aw_wb", wbin, (size_t)count * 8, &dwb));
    SK_TRY(up(ctx, "aw_a1", a1, (size_t)count * 8, &da1));
    SK_TRY(up(ctx, "aw_a2", a2, (size_t)count * 8, &da2));
    if (degrid) SK_TRY(sk_scratch(ctx, "pre_vis", (size_t)std::max<i64>(count, 1) * 16, &dvis));
    else SK_TRY(up(ctx, "pre_vis", vis, (size_t)count * 16, &dvis));
    SK_TRY(aw_core_dev(ctx, nw, qpx, s, (double *)dwk, nant, (double *)dak, height, width, (double *)dgrid, count, (double *)du, (double *)dv,
                       (i64 *)dwb, (i64 *)da1, (i64 *)da2, (double *)dvis, degrid));
    if (degrid) SK_TRY(down(ctx, vis_out, dvis, (size_t)count * 16));
    else SK_TRY(down(ctx, grid, dgrid, (size_t)(height * width) * 16));
    SK_TRY(t.finish());
    return check_flags(ctx, what);
}

extern "C" int skagrid_convgrid_aw(skagrid_ctx *ctx, int64_t nw, int64_t qpx, int64_t s, const double *wkerns, int64_t nant,
                                   const double *akerns, int64_t height, int64_t width, double *grid, int64_t count, const double *u,
                                   const double *v, const int64_t *wbin, const int64_t *a1, const int64_t *a2, const double *vis) {
    return aw_host(ctx, nw, qpx, s, wkerns, nant, akerns, height, width, grid, count, u, v, wbin, a1, a2, vis, nullptr, 0, "convgrid_aw");
}

extern "C" int skagrid_convdegrid_aw(skagrid_ctx *ctx, int64_t nw, int64_t qpx, int64_t s, const double *wkerns, int64_t nant,
                                     const double *akerns, int64_t height, int64_t width, const double *grid, int64_t count,
                                     const double *u, const double *v, const int64_t *wbin, const int64_t *a1, const int64_t *a2,
                                     double *vis_out) {
    return aw_host(ctx, nw, qpx, s, wkerns, nant, akerns, height, width, const_cast<double *>(grid), count, u, v, wbin, a1, a2, nullptr,
                   vis_out, 1, "convdegrid_aw");
}

extern "C" int skagrid_convolve2d(skagrid_ctx *ctx, int64_t n, const double *a1, const double *a2, double *out) {
    SK_TRY(sk_api_enter(ctx));
    NEED(ctx, n > 0 && n <= 64 && a1 && a2 && out, "convolve2d: size outside [1,64] or NULL pointer");
    Timer t(ctx);
    void *d1, *d2, *dout;
    SK_TRY(up(ctx, "cv_a", a1, (size_t)(n * n) * 16, &d1));
    SK_TRY(up(ctx, "cv_b", a2, (size_t)(n * n) * 16, &d2));
    SK_TRY(sk_scratch(ctx, "cv_out", (size_t)(n * n) * 16, &dout));
    SK_TRY(sk_convolve2d_dev(ctx, n, 1, (double *)d1, nullptr, (double *)d2, nullptr, (double *)dout, 0, ctx->stream));
    SK_TRY(down(ctx, out, dout, (size_t)(n * n) * 16));
    return t.finish();
}

extern "C" int skagrid_aw_kernel(skagrid_ctx *ctx, int64_t nw, int64_t qpx, int64_t s, const double *wkerns, int64_t nant,
                                 const double *akerns, int64_t count, const int64_t *wbin, const int64_t *yf, const int64_t *xf,
                                 const int64_t *a1, const int64_t *a2, double *out) {
    SK_TRY(sk_api_enter(ctx));
    NEED(ctx, nw > 0 && qpx > 0 && s > 0 && s <= 64 && nant > 0 && count >= 0, "aw_kernel: bad dimension");
    if (count == 0) return SKAGRID_OK;
    NEED(ctx, wkerns && akerns && wbin && yf && xf && a1 && a2 && out, "aw_kernel: NULL pointer");
    Timer t(ctx);
    void *dwk, *dak, *dwb, *dyf, *dxf, *da1, *da2, *dout;
    SK_TRY(up(ctx, "aw_wk", wkerns, (size_t)(nw * qpx * qpx * s * s) * 16, &dwk));
    SK_TRY(up(ctx, "aw_ak", akerns, (size_t)(nant * s * s) * 16, &dak));
    SK_TRY(up(ctx, "aw_wb", wbin, (size_t)count * 8, &dwb));
    SK_TRY(up(ctx, "aw_yf", yf, (size_t)count * 8, &dyf));
    SK_TRY(up(ctx, "aw_xf", xf, (size_t)count * 8, &dxf));
    SK_TRY(up(ctx, "aw_a1", a1, (size_t)count * 8, &da1));
    SK_TRY(up(ctx, "aw_a2", a2, (size_t)count * 8, &da2));
    SK_TRY(sk_scratch(ctx, "aw_kern", (size_t)(count * s * s) * 16, &dout));
    SK_TRY(sk_aw_kernels_dev(ctx, nw, qpx, s, (double *)dwk, nant, (double *)dak, count, (i64 *)dwb, (i64 *)dyf, (i64 *)dxf, (i64 *)da1,
                             (i64 *)da2, (double *)dout, 0, ctx->d_flags, ctx->stream));
    SK_TRY(down(ctx, out, dout, (size_t)(count * s * s) * 16));
    SK_TRY(t.finish());
    return check_flags(ctx, "aw_kernel");
}

// ------------------------------------------------------------------------------------------ grid -> image
extern "C" int skagrid_make_grid_hermitian(skagrid_ctx *ctx, int64_t n, const double *grid, double *out) {
    SK_TRY(sk_api_enter(ctx));
    NEED(ctx, n > 0 && grid && out, "make_grid_hermitian: bad size or NULL pointer");
    Timer t(ctx);
    void *dg;
    SK_TRY((ctx->resident_h = ctx->resident_w = 0, up(ctx, "grid", grid, (size_t)(n * n) * 16, &dg)));  // overwrites the resident grid
    SK_TRY(sk_hermitian_dev(ctx, n, (double *)dg, (double *)dg, ctx->stream));
    SK_TRY(down(ctx, out, dg, (size_t)(n * n) * 16));
    return t.finish();
}

extern "C" int skagrid_ifft(skagrid_ctx *ctx, int64_t n, const double *grid, double *out) {
    SK_TRY(sk_api_enter(ctx));
    NEED(ctx, n > 0 && grid && out, "ifft: bad size or NULL pointer");
    Timer t(ctx);
    void *dg;
    SK_TRY((ctx->resident_h = ctx->resident_w = 0, up(ctx, "grid", grid, (size_t)(n * n) * 16, &dg)));  // overwrites the resident grid
    SK_TRY(sk_fft2c_dev(ctx, n, (double *)dg, (double *)dg, 1, ctx->stream));
    SK_TRY(down(ctx, out, dg, (size_t)(n * n) * 16));
    return t.finish();
}

extern "C" int skagrid_fft(skagrid_ctx *ctx, int64_t n, const double *grid, double *out) {
    // src/Gridding.hs:821-826: pad to the next power of two (transposing padder), centred forward FFT, crop
    SK_TRY(sk_api_enter(ctx));
    NEED(ctx, n > 0 && grid && out, "fft: bad size or NULL pointer");
    Timer t(ctx);
    i64 big = 1;
    while (big < n) big <<= 1;
    void *dg;
    SK_TRY((ctx->resident_h = ctx->resident_w = 0, up(ctx, "grid", grid, (size_t)(n * n) * 16, &dg)));  // overwrites the resident grid
    if (big == n) {
        SK_TRY(sk_fft2c_dev(ctx, n, (double *)dg, (double *)dg, 0, ctx->stream));
        SK_TRY(down(ctx, out, dg, (size_t)(n * n) * 16));
    } else {
        void *db;
        SK_TRY(sk_scratch(ctx, "fft_big", (size_t)(big * big) * 16, &db));
        SK_TRY(sk_pad_crop_dev(ctx, n, (double *)dg, big, (double *)db, ctx->stream));
        SK_TRY(sk_fft2c_dev(ctx, big, (double *)db, (double *)db, 0, ctx->stream));
        SK_TRY(sk_pad_crop_dev(ctx, big, (double *)db, n, (double *)dg, ctx->stream));
        SK_TRY(down(ctx, out, dg, (size_t)(n * n) * 16));
    }
    return t.finish();
}

extern "C" int skagrid_grid_to_image(skagrid_ctx *ctx, int64_t n, const double *grid, double *image, double *max_out) {
    SK_TRY(sk_api_enter(ctx));
    NEED(ctx, n > 0, "grid_to_image: bad size");
    Timer t(ctx);
    void *dg, *dimg = nullptr, *dmax;
    SK_TRY(grid_in(ctx, grid, n, n, &dg, "grid_to_image"));
    if (image) SK_TRY(sk_scratch(ctx, "image", (size_t)(n * n) * 8, &dimg));
    SK_TRY(sk_scratch(ctx, "image_max", 16, &dmax));
    SK_TRY(sk_grid_to_image_dev(ctx, n, (double *)dg, (double *)dimg, (double *)dmax, ctx->stream));
    if (image) SK_TRY(down(ctx, image, dimg, (size_t)(n * n) * 8));
    if (max_out) SK_TRY(down(ctx, max_out, dmax, 8));
    return t.finish();
}

extern "C" int skagrid_dev_grid_to_image(skagrid_ctx *ctx, int64_t n, double *grid, double *image, double *max_out, void *stream) {
    SK_TRY(sk_api_enter(ctx));
    NEED(ctx, n > 0 && grid, "dev_grid_to_image: bad size or NULL grid");
    return sk_grid_to_image_dev(ctx, n, grid, image, max_out, sk_stream(ctx, stream));
}

// ------------------------------------------------------------------------------------------ imaging drivers
extern "C" int skagrid_simple_imaging(skagrid_ctx *ctx, double theta, int64_t lam, int64_t count, const double *u, const double *v,
                                      const double *w, const double *vis, double *grid_out) {
    SK_TRY(sk_api_enter(ctx));
    const i64 n = grid_side(theta, lam);
    NEED(ctx, n > 0 && grid_out && count >= 0, "simple_imaging: bad size or NULL grid");
    (void)w;
    Timer t(ctx);
    void *dgrid;
    SK_TRY((ctx->resident_h = ctx->resident_w = 0, sk_scratch(ctx, "grid", (size_t)(n * n) * 16, &dgrid)));  // overwrites the resident grid
    SK_CUDA(ctx, cudaMemsetAsync(dgrid, 0, (size_t)(n * n) * 16, ctx->stream));
    if (count > 0) {
        NEED(ctx, u && v && vis, "simple_imaging: NULL visibility array");
        double *du, *dv, *dw = nullptr;
        void *dvis;
        SK_TRY(up_uvw(ctx, count, u, v, nullptr, &du, &dv, &dw));
        SK_TRY(up(ctx, "pre_vis", vis, (size_t)count * 16, &dvis));
        void *dscr;  // div3 divides all three coordinates; w is not used by `grid`
        SK_TRY(sk_scratch(ctx, "pre_w", (size_t)count * 8, &dscr));
        SK_CUDA(ctx, cudaMemsetAsync(dscr, 0, (size_t)count * 8, ctx->stream));
        SK_TRY(sk_scale3_dev(ctx, count, du, dv, (double *)dscr, (double)lam, 1, ctx->stream));
        SK_TRY(sk_grid_simple_dev(ctx, n, n, (double *)dgrid, count, du, dv, (double *)dvis, ctx->stream));
    }
    SK_TRY(down(ctx, grid_out, dgrid, (size_t)(n * n) * 16));
    return t.finish();
}

extern "C" int skagrid_conv_imaging(skagrid_ctx *ctx, int64_t qpx, int64_t gh, int64_t gw, const double *gcf, double theta, int64_t lam,
                                    int64_t count, const double *u, const double *v, const double *w, const double *vis, double *grid_out) {
    SK_TRY(sk_api_enter(ctx));
    const i64 n = grid_side(theta, lam);
    SK_TRY(check_table_args(ctx, 1, qpx, gh, gw, n, n, count));
    NEED(ctx, gcf && grid_out, "conv_imaging: NULL kernel or grid");
    (void)w;
    Timer t(ctx);
    void *dgrid, *dtab;
    SK_TRY(up(ctx, "tab", gcf, (size_t)(qpx * qpx * gh * gw) * 16, &dtab));
    SK_TRY((ctx->resident_h = ctx->resident_w = 0, sk_scratch(ctx, "grid", (size_t)(n * n) * 16, &dgrid)));  // overwrites the resident grid
    SK_CUDA(ctx, cudaMemsetAsync(dgrid, 0, (size_t)(n * n) * 16, ctx->stream));
    if (count > 0) {
        NEED(ctx, u && v && vis, "conv_imaging: NULL visibility array");
        double *du, *dv, *dw = nullptr;
        void *dvis;
        SK_TRY(up_uvw(ctx, count, u, v, nullptr, &du, &dv, &dw));
        SK_TRY(up(ctx, "pre_vis", vis, (size_t)count * 16, &dvis));
        void *dscr;
        SK_TRY(sk_scratch(ctx, "pre_w", (size_t)count * 8, &dscr));
        SK_CUDA(ctx, cudaMemsetAsync(dscr, 0, (size_t)count * 8, ctx->stream));
        SK_TRY(sk_scale3_dev(ctx, count, du, dv, (double *)dscr, (double)lam, 1, ctx->stream));
        skagrid_geom geom = {n, n, 0, n, 1, qpx, gh, gw};
        skagrid_plan *plan = nullptr;
        SK_TRY(plan_acquire(ctx, &geom, count, 0, &plan));
        int rc = sk_plan_fill(ctx, plan, count, du, dv, nullptr, (double *)dvis, ctx->stream);
        if (!rc) rc = skagrid_dev_grid(ctx, plan, (double *)dtab, (double *)dgrid, 0, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
        plan_release(ctx, plan);
        SK_TRY(rc);
    }
    SK_TRY(down(ctx, grid_out, dgrid, (size_t)(n * n) * 16));
    return t.finish();
}

// conv_imaging with a w-indexed table: the 5-D analogue of conv_imaging (src/Gridding.hs:115-124) and the last step of
// w_cache_imaging (:421-449): zero grid of side round(theta*lam), p = uvw/lam, convgrid2.  Nothing but the
// visibilities and the table goes up, only the grid comes back.
extern "C" int skagrid_conv_imaging2(skagrid_ctx *ctx, int64_t nw, int64_t qpx, int64_t gh, int64_t gw, const double *gcf, double theta,
                                     int64_t lam, int64_t count, const double *u, const double *v, const double *w, const int64_t *wbin,
                                     const double *vis, double *grid_out) {
    SK_TRY(sk_api_enter(ctx));
    const i64 n = grid_side(theta, lam);
    SK_TRY(check_table_args(ctx, nw, qpx, gh, gw, n, n, count));
    NEED(ctx, gcf, "conv_imaging2: NULL kernel table");
    if (count > 0) NEED(ctx, u && v && wbin && vis, "conv_imaging2: NULL visibility array");
    (void)w;
    Timer t(ctx);
    void *dgrid, *dtab;
    SK_TRY(up(ctx, "tab", gcf, (size_t)(nw * qpx * qpx * gh * gw) * 16, &dtab));
    SK_TRY(grid_fresh(ctx, n, n, &dgrid));
    skagrid_geom geom = {n, n, 0, n, nw, qpx, gh, gw};
    SK_TRY(stream_table(ctx, &geom, (double *)dtab, (double *)dgrid, count, u, v, wbin, vis, nullptr, 0, 1, (double)lam));
    if (grid_out) SK_TRY(down(ctx, grid_out, dgrid, (size_t)(n * n) * 16));
    SK_TRY(t.finish());
    return check_flags(ctx, "conv_imaging2");
}

// aw_imaging on device-resident inputs: p = uvw/lam (in place), wbin = findClosest wbins w (w in wavelengths,
// NOT divided: src/Gridding.hs:473 uses the original w), convgrid4.  d_grid must be zeroed by the caller.
static int aw_imaging_dev(skagrid_ctx *ctx, i64 n, i64 lam, i64 nw, i64 qpx, i64 s, const double *d_wk, const double *d_wbins, i64 nant,
                          const double *d_ak, i64 count, double *du, double *dv, double *dw, const i64 *da1, const i64 *da2, double *dvis,
                          double *d_grid) {
    if (count <= 0) return SKAGRID_OK;
    void *dwb;
    SK_TRY(sk_scratch(ctx, "aw_wb", (size_t)count * 8, &dwb));
    SK_TRY(sk_find_closest_dev(ctx, nw, d_wbins, count, dw, (i64 *)dwb, ctx->stream));
    SK_TRY(sk_scale3_dev(ctx, count, du, dv, dw, (double)lam, 1, ctx->stream));
    return aw_core_dev(ctx, nw, qpx, s, d_wk, nant, d_ak, n, n, d_grid, count, du, dv, (i64 *)dwb, da1, da2, dvis, 0);
}

extern "C" int skagrid_aw_imaging(skagrid_ctx *ctx, double theta, int64_t lam, int64_t nw, int64_t qpx, int64_t s, const double *wkerns,
                                  const double *wbins, int64_t nant, const double *akerns, int64_t count, const double *u, const double *v,
                                  const double *w, const int64_t *a1, const int64_t *a2, const double *vis, double *grid_out) {
    return skagrid_aw_gridding(ctx, theta, lam, nw, qpx, s, wkerns, wbins, nant, akerns, count, u, v, w, a1, a2, -1.0, vis, nullptr, nullptr,
                               grid_out);
}

// freq < 0 selects plain aw_imaging (no uvw_lambda / doweight / mirror / image stage).
extern "C" int skagrid_aw_gridding(skagrid_ctx *ctx, double theta, int64_t lam, int64_t nw, int64_t qpx, int64_t s, const double *wkerns,
                                   const double *wbins, int64_t nant, const double *akerns, int64_t count, const double *u_m,
                                   const double *v_m, const double *w_m, const int64_t *a1, const int64_t *a2, double freq,
                                   const double *vis, double *image, double *max_out, double *grid_out) {
    SK_TRY(sk_api_enter(ctx));
    const i64 n = grid_side(theta, lam);
    SK_TRY(check_table_args(ctx, nw, qpx, s, s, n, n, count));
    NEED(ctx, nant > 0 && wkerns && wbins && akerns, "aw_gridding: NULL kernels or nant <= 0");
    NEED(ctx, s <= 63, "AW path: support above 63 is not supported");
    if (count > 0) NEED(ctx, u_m && v_m && w_m && a1 && a2 && vis, "aw_gridding: NULL visibility array");
    const bool full = freq >= 0.0;
    Timer t(ctx);
    void *dwk, *dak, *dwbins, *dgrid, *du, *dv, *dw, *da1, *da2, *dvis;
    SK_TRY(up(ctx, "aw_wk", wkerns, (size_t)(nw * qpx * qpx * s * s) * 16, &dwk));
    SK_TRY(up(ctx, "aw_ak", akerns, (size_t)(nant * s * s) * 16, &dak));
    SK_TRY(up(ctx, "aw_wbins", wbins, (size_t)nw * 8, &dwbins));
    SK_TRY((ctx->resident_h = ctx->resident_w = 0, sk_scratch(ctx, "grid", (size_t)(n * n) * 16, &dgrid)));  // overwrites the resident grid
    SK_CUDA(ctx, cudaMemsetAsync(dgrid, 0, (size_t)(n * n) * 16, ctx->stream));
    SK_TRY(up(ctx, "pre_u", u_m, (size_t)count * 8, &du));
    SK_TRY(up(ctx, "pre_v", v_m, (size_t)count * 8, &dv));
    SK_TRY(up(ctx, "pre_w", w_m, (size_t)count * 8, &dw));
    SK_TRY(up(ctx, "aw_a1", a1, (size_t)count * 8, &da1));
    SK_TRY(up(ctx, "aw_a2", a2, (size_t)count * 8, &da2));
    SK_TRY(up(ctx, "pre_vis", vis, (size_t)count * 16, &dvis));
    if (full && count > 0) {
        // src/ImageDataset.hs:55-60,72: uvw_lambda; wt = doweight(ones) on the UN-mirrored uvw; mirror; vis*wt
        SK_TRY(sk_scale3_dev(ctx, count, (double *)du, (double *)dv, (double *)dw, freq / 299792458.0, 0, ctx->stream));
        void *dwt;
        SK_TRY(sk_scratch(ctx, "aw_wt", (size_t)count * 16, &dwt));
        SK_TRY(sk_fill_complex_dev(ctx, count, (double *)dwt, 1.0, 0.0, ctx->stream));  // `ones` of src/ImageDataset.hs:58
        SK_TRY(sk_doweight_dev(ctx, n, (double)lam, count, (double *)du, (double *)dv, (double *)dwt, ctx->d_flags + 0, ctx->stream));
        SK_TRY(sk_mirror_dev(ctx, count, (double *)du, (double *)dv, (double *)dw, (double *)dvis, ctx->stream));
        SK_TRY(sk_cmul_dev(ctx, count, (double *)dvis, (double *)dwt, ctx->stream));
    }
    SK_TRY(aw_imaging_dev(ctx, n, lam, nw, qpx, s, (double *)dwk, (double *)dwbins, nant, (double *)dak, count, (double *)du, (double *)dv,
                          (double *)dw, (i64 *)da1, (i64 *)da2, (double *)dvis, (double *)dgrid));
    if (grid_out) SK_TRY(down(ctx, grid_out, dgrid, (size_t)(n * n) * 16));
    if (full && (image || max_out)) {
        void *dimg = nullptr, *dmax;
        if (image) SK_TRY(sk_scratch(ctx, "image", (size_t)(n * n) * 8, &dimg));
        SK_TRY(sk_scratch(ctx, "image_max", 16, &dmax));
        SK_TRY(sk_grid_to_image_dev(ctx, n, (double *)dgrid, (double *)dimg, (double *)dmax, ctx->stream));
        if (image) SK_TRY(down(ctx, image, dimg, (size_t)(n * n) * 8));
        if (max_out) SK_TRY(down(ctx, max_out, dmax, 8));
    }
    SK_TRY(t.finish());
    return check_flags(ctx, "aw_gridding");
}

// ------------------------------------------------------------------------------------------ w-kernels
extern "C" int skagrid_w_kernels(skagrid_ctx *ctx, double theta, int64_t nw, const double *w, int64_t npixff, int64_t npixkern, int64_t qpx,
                                 int conjugate, double *out) {
    SK_TRY(sk_api_enter(ctx));
    NEED(ctx, nw > 0 && w && out, "w_kernels: NULL pointer or nw <= 0");
    Timer t(ctx);
    const size_t bytes = (size_t)(nw * qpx * qpx * npixkern * npixkern) * 16;
    void *dout;
    SK_TRY(sk_scratch(ctx, "wkern_out", bytes, &dout));
    SK_TRY(sk_w_kernels_dev(ctx, theta, nw, w, npixff, npixkern, qpx, conjugate, (double *)dout, ctx->stream));
    SK_TRY(down(ctx, out, dout, bytes));
    return t.finish();
}

// Device-output variant used by bench.py to build the kernel table without a host round trip.
// w_kernel with the KernelOptions of src/Gridding.hs:30-38 that move the far-field coordinates (kernel_coordinates, :620-635):
// transmat = patTransMat as 4 doubles, row-major t[r][c] (NULL: identity), dl / dm = patHorShift / patVerShift.
extern "C" int skagrid_w_kernels_ex(skagrid_ctx *ctx, double theta, int64_t nw, const double *w, int64_t npixff, int64_t npixkern, int64_t qpx,
                                    int conjugate, const double *transmat, double dl, double dm, double *out) {
    SK_TRY(sk_api_enter(ctx));
    NEED(ctx, nw > 0 && w && out, "w_kernels_ex: NULL pointer or nw <= 0");
    Timer t(ctx);
    const size_t bytes = (size_t)(nw * qpx * qpx * npixkern * npixkern) * 16;
    void *dout;
    SK_TRY(sk_scratch(ctx, "wkern_out", bytes, &dout));
    SK_TRY(sk_w_kernels_ex_dev(ctx, theta, nw, w, npixff, npixkern, qpx, conjugate, transmat, dl, dm, (double *)dout, ctx->stream));
    SK_TRY(down(ctx, out, dout, bytes));
    return t.finish();
}

extern "C" int skagrid_dev_w_kernels(skagrid_ctx *ctx, double theta, int64_t nw, const double *w_host, int64_t npixff, int64_t npixkern,
                                     int64_t qpx, int conjugate, double *d_out, void *stream) {
    SK_TRY(sk_api_enter(ctx));
    NEED(ctx, nw > 0 && w_host && d_out, "dev_w_kernels: NULL pointer or nw <= 0");
    return sk_w_kernels_dev(ctx, theta, nw, w_host, npixff, npixkern, qpx, conjugate, d_out, sk_stream(ctx, stream));
}

// Device-resident binning (used by the uv-tile-sharded router: the owner of a visibility is an integer function
// of its bit-exact y cell).
extern "C" int skagrid_dev_frac_coord(skagrid_ctx *ctx, int64_t n, int64_t qpx, int64_t count, const double *d_p, int64_t *d_fl,
                                      int64_t *d_frac, int flags, void *stream) {
    SK_TRY(sk_api_enter(ctx));
    NEED(ctx, n > 0 && qpx > 0 && count >= 0, "dev_frac_coord: n, qpx must be positive");
    if (count == 0) return SKAGRID_OK;
    NEED(ctx, d_p && d_fl && d_frac, "dev_frac_coord: NULL pointer");
    return sk_frac_coord_dev(ctx, n, qpx, count, d_p, (i64 *)d_fl, (i64 *)d_frac, flags & SKAGRID_FRAC_NORMALISE, sk_stream(ctx, stream));
}

// ------------------------------------------------------------------------------------------ device-resident pre-steps
// The element-wise pre-steps of ImageDataset.aw_gridding on device arrays, for callers that keep the visibilities
// resident (same kernels as the host-pointer functions above).
extern "C" int skagrid_dev_uvw_scale(skagrid_ctx *ctx, int64_t count, double *d_u, double *d_v, double *d_w, double a, int divide,
                                     void *stream) {
    SK_TRY(sk_api_enter(ctx));
    if (count <= 0) return SKAGRID_OK;
    NEED(ctx, d_u && d_v && d_w, "dev_uvw_scale: NULL pointer");
    return sk_scale3_dev(ctx, count, d_u, d_v, d_w, a, divide, sk_stream(ctx, stream));
}

extern "C" int skagrid_dev_mirror_uvw(skagrid_ctx *ctx, int64_t count, double *d_u, double *d_v, double *d_w, double *d_vis, void *stream) {
    SK_TRY(sk_api_enter(ctx));
    if (count <= 0) return SKAGRID_OK;
    NEED(ctx, d_u && d_v && d_w, "dev_mirror_uvw: NULL pointer");
    return sk_mirror_dev(ctx, count, d_u, d_v, d_w, d_vis, sk_stream(ctx, stream));
}

extern "C" int skagrid_dev_find_closest(skagrid_ctx *ctx, int64_t nw, const double *d_wbins, int64_t count, const double *d_w,
                                        int64_t *d_out, void *stream) {
    SK_TRY(sk_api_enter(ctx));
    NEED(ctx, nw > 0, "dev_find_closest: empty wbins");
    if (count <= 0) return SKAGRID_OK;
    NEED(ctx, d_wbins && d_w && d_out, "dev_find_closest: NULL pointer");
    return sk_find_closest_dev(ctx, nw, d_wbins, count, d_w, (i64 *)d_out, sk_stream(ctx, stream));
}

// Out-of-grid visibilities set bit 1 of the context's error word; skagrid_dev_take_error reads and clears it.
extern "C" int skagrid_dev_doweight(skagrid_ctx *ctx, double theta, int64_t lam, int64_t count, const double *d_u, const double *d_v,
                                    double *d_vis, void *stream) {
    SK_TRY(sk_api_enter(ctx));
    if (count <= 0) return SKAGRID_OK;
    NEED(ctx, d_u && d_v && d_vis, "dev_doweight: NULL pointer");
    const i64 n = grid_side(theta, lam);
    NEED(ctx, n > 0, "dev_doweight: round(theta*lam) must be positive");
    return sk_doweight_dev(ctx, n, (double)lam, count, d_u, d_v, d_vis, ctx->d_flags, sk_stream(ctx, stream));
}

extern "C" int skagrid_dev_slab_fft_rows(skagrid_ctx *ctx, int64_t n, int64_t row0, int64_t nrows, double *d_slab, void *stream) {
    SK_TRY(sk_api_enter(ctx));
    NEED(ctx, d_slab || nrows == 0, "dev_slab_fft_rows: NULL slab");
    return sk_slab_fft_rows_dev(ctx, n, row0, nrows, d_slab, sk_stream(ctx, stream));
}

extern "C" int skagrid_dev_slab_fft_cols(skagrid_ctx *ctx, int64_t n, int64_t col0, int64_t ncols, double *d_cols, double *d_image,
                                         double *d_max, void *stream) {
    SK_TRY(sk_api_enter(ctx));
    NEED(ctx, d_cols || ncols == 0, "dev_slab_fft_cols: NULL column slab");
    return sk_slab_fft_cols_dev(ctx, n, col0, ncols, d_cols, d_image, d_max, sk_stream(ctx, stream));
}

extern "C" int skagrid_dev_weight_count(skagrid_ctx *ctx, double theta, int64_t lam, int64_t count, const double *d_u, const double *d_v,
                                        int32_t *d_hist, void *stream) {
    SK_TRY(sk_api_enter(ctx));
    if (count <= 0) return SKAGRID_OK;
    NEED(ctx, d_u && d_v && d_hist, "dev_weight_count: NULL pointer");
    const i64 n = grid_side(theta, lam);
    NEED(ctx, n > 0, "dev_weight_count: round(theta*lam) must be positive");
    return sk_weight_count_dev(ctx, n, (double)lam, count, d_u, d_v, reinterpret_cast<uint32_t *>(d_hist), ctx->d_flags, sk_stream(ctx, stream));
}

extern "C" int skagrid_dev_weight_apply(skagrid_ctx *ctx, double theta, int64_t lam, int64_t count, const double *d_u, const double *d_v,
                                        const int32_t *d_hist, double *d_vis, void *stream) {
    SK_TRY(sk_api_enter(ctx));
    if (count <= 0) return SKAGRID_OK;
    NEED(ctx, d_u && d_v && d_hist && d_vis, "dev_weight_apply: NULL pointer");
    const i64 n = grid_side(theta, lam);
    NEED(ctx, n > 0, "dev_weight_apply: round(theta*lam) must be positive");
    return sk_weight_apply_dev(ctx, n, (double)lam, count, d_u, d_v, reinterpret_cast<const uint32_t *>(d_hist), d_vis, sk_stream(ctx, stream));
}

// Device pointer and shape of the grid the last host-pointer call left resident (NULL grid pointers, include/skagrid.h).
// For callers that run one process per GPU and reduce the per-process grids themselves (NCCL all-reduce on this buffer
// between skagrid_conv_imaging2 and skagrid_grid_to_image / skagrid_convdegrid2): the library does no inter-process exchange.
extern "C" int skagrid_resident_grid(skagrid_ctx *ctx, double **d_grid, int64_t *height, int64_t *width) {
    SK_TRY(sk_api_enter(ctx));
    NEED(ctx, d_grid && height && width, "resident_grid: NULL output pointer");
    *d_grid = nullptr; *height = *width = 0;
    if (ctx->resident_h <= 0 || ctx->resident_w <= 0) return sk_fail(ctx, SKAGRID_EINVAL, "resident_grid: the context holds no resident grid");
    void *p = nullptr;
    SK_TRY(sk_scratch(ctx, "grid", (size_t)(ctx->resident_h * ctx->resident_w) * 16, &p));  // existing buffer: never grows here
    *d_grid = (double *)p; *height = ctx->resident_h; *width = ctx->resident_w;
    return SKAGRID_OK;
}

extern "C" int skagrid_dev_take_error(skagrid_ctx *ctx, void *stream, int *flags_out) {
    SK_TRY(sk_api_enter(ctx));
    uint32_t f = 0;
    SK_TRY(sk_take_flags(ctx, sk_stream(ctx, stream), &f));
    if (flags_out) *flags_out = (int)f;
    return SKAGRID_OK;
}
