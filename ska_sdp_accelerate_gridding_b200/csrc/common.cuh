// common.cuh -- internal declarations shared by the libskagrid translation units.
// Not part of the public ABI (that is include/skagrid.h).
#pragma once

#include <cuda_runtime.h>
#include <cufft.h>
#include <stdint.h>

#include <map>
#include <set>
#include <string>
#include <vector>

#include "skagrid.h"

typedef long long i64;
typedef unsigned long long u64;

// ---------------------------------------------------------------------------------------------
// Sorted visibility record: what the gridder / degridder stream. 32 bytes so that a record is two
// 128-bit loads and a bucket is a legal cp.async.bulk (TMA) source (16-byte aligned, 16-byte multiple).
// ---------------------------------------------------------------------------------------------
struct __align__(16) VisRec {
    double re, im;       // visibility (0 for degrid-only plans)
    uint32_t kbase;      // (slice * gh*kpitch - (dy*kpitch + dx)) mod 2^32, slice = (wbin*qpx + yf)*qpx + xf or the visibility index (AW):
                         // padded-table element of tap (i,j) of this visibility = kbase + (dy+i)*kpitch + (dx+j)
    uint32_t loc;        // one-hot(dy*MT + dx) << 16 | ly << 8 | lx: footprint origin inside the tile (0..tile-1) and inside its micro-tile
    uint32_t index;      // position of the visibility in the caller's arrays (degrid output slot)
    uint32_t tile;       // uv tile ty * ntx + tx (lets a record be placed on the grid without its work item)
};
static_assert(sizeof(VisRec) == 32, "VisRec must be 32 bytes");

// One unit of gridder work: a run of records that all belong to one uv tile.
struct WorkItem {
    uint32_t tile;       // ty * ntx + tx
    uint32_t begin, end; // record range
    uint32_t pad;
};

// The uv tile edge (in footprint-origin cells) is a per-plan choice, 16 or 32 (Geom::tile): small tiles keep the
// shared-memory subgrid small (more resident blocks per SM: best for dense uv coverage), large tiles amortise the
// per-tile zero/flush over more visibilities (best for sparse coverage and for wide kernels).
constexpr int CHUNK = 4096;                    // max records per work item (load balance)

// Geometry of one plan.  A uv tile is TILE x TILE footprint origins; inside a tile the origins are
// bucketed by MT x MT micro-tiles.  All footprints of one micro-tile lie inside an R x R cell region
// (MT - 1 + S <= R), which is what one pass of the tiled gridder holds in registers.
struct Geom {
    i64 height, width, row0, row1;
    i64 nw, qpx, gh, gw;
    int ntx, nty;        // tiles per dimension
    int R;               // register region edge: 16, 32, 48 or 64 (0: shape not supported by the tiled kernels)
    int MT;              // micro-tile edge: 2 or 4
    int tile, tshift;    // uv tile edge (16 or 32) and its log2
    int MTR;             // micro-tiles per tile row = tile / MT
    int SG;              // shared-memory subgrid edge = tile - MT + R
    int kpitch;          // row pitch (taps) of the padded copy of the kernel table the kernels read: gw rounded up to 16
                         // (rows start on 256-byte boundaries: a 15-tap row is 2 L1 lines instead of up to 3); gw if R == 0
    int krows;           // rows per slice of the padded copy: gh, or R in the dense layout
    int dense;           // dense layout: every slice is R x R taps (zero beyond gh x gw) behind one leading all-zero slice, so a
                         // thread of the tiled gridder can load the tap of ANY of its residues without a validity test -- taps
                         // outside the footprint read zeros of this or the previous slice (grid_dense_kernel).  Chosen when the
                         // padding costs <= 15 % more L2 traffic (S = 14, 15, 30, 31; the headline configurations)
    int kpt;             // bucket keys per uv tile: MTR*MTR (micro-tile buckets) or tile*tile (cell buckets, `cellsort`)
    int cellsort;        // records are additionally grouped by exact footprint origin inside each micro-tile (lets the
                         // degridder keep its grid column in registers across a run of same-cell visibilities)
    i64 nkeys;           // ntx * nty * kpt
    int normalise;
};

struct skagrid_plan {
    Geom g;
    i64 capacity;        // max visibilities
    i64 count;           // visibilities of the current batch
    int slice_override;
    int has_vis;         // the records carry visibilities (0: degrid-only plan)
    uint32_t *d_offs;    // [nkeys + 1] bucket ends after the scatter (starts after the scan)
    VisRec *d_rec;       // [capacity]
    WorkItem *d_items;   // [max_items]
    i64 max_items;
    // d_counters: [0] n_items, [1] gridder queue head, [2] kept, [3] dropped, [4] non-empty tiles,
    //             [5] degridder queue head
    uint32_t *d_counters;
    uint32_t *d_blocksums;  // scan scratch
    i64 nblocksums;
    double2 *d_table;       // padded copy of the kernel table (refreshed by every grid / degrid call), lazily allocated
    size_t table_bytes;
};

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
};

struct skagrid_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;      // compute stream
    cudaStream_t copy_stream = nullptr; // H2D prefetch stream of the chunked host API
    cudaStream_t d2h_stream = nullptr;  // D2H stream of the chunked degridder (results leave while the next chunk computes)
    cudaEvent_t ev_k[2] = {nullptr, nullptr}, ev_d2h[2] = {nullptr, nullptr};
    cudaEvent_t ev_mg[2] = {nullptr, nullptr};  // multi-device entry points: [0] local phase done, [1] exchange phase done
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_copy[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    std::string err;
    uint32_t *d_flags = nullptr;        // device error word: bit0 index out of range, bit1 cell outside the weight grid
    double last_ms = 0.0;
    i64 launches = 0;
    std::map<std::string, DevBuf> pool;   // named scratch buffers, grown on demand, freed at destroy
    std::map<i64, cufftHandle> fft_plans; // n -> Z2Z n x n plan
    std::map<i64, DevBuf> fft_work;
    i64 resident_h = 0, resident_w = 0;   // shape of the grid the last host-pointer call left in the "grid" scratch (0: none)
    std::set<const void *> smem_configured;  // kernels whose dynamic shared-memory limit was raised on this device
    i64 res_count = 0;                    // visibilities whose (u, v, wbin) the last table call left in the res_* scratch (0: none)
    int res_has_wbin = 0;
    void *h_pinned = nullptr;             // pinned host staging of the multi-device calls (sk_host_scratch)
    size_t h_pinned_bytes = 0;
    skagrid_plan *cached_plan = nullptr;  // plan kept between host-pointer calls (api.cu plan_acquire)
};

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
int sk_fail(skagrid_ctx *ctx, int code, const char *fmt, ...);

#define SK_CUDA(ctx, call)                                                                      \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess)                                                                 \
            return sk_fail((ctx), SKAGRID_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call,     \
                           cudaGetErrorString(e__));                                            \
    } while (0)

#define SK_TRY(call)                 \
    do {                             \
        int rc__ = (call);           \
        if (rc__ != SKAGRID_OK) return rc__; \
    } while (0)

#define SK_LAUNCH_CHECK(ctx)                                                                    \
    do {                                                                                        \
        (ctx)->launches++;                                                                      \
        cudaError_t e__ = cudaGetLastError();                                                   \
        if (e__ != cudaSuccess)                                                                 \
            return sk_fail((ctx), SKAGRID_ECUDA, "%s:%d kernel launch: %s", __FILE__, __LINE__, \
                           cudaGetErrorString(e__));                                            \
    } while (0)

// named, growable device scratch (never shrinks; contents undefined after a grow)
int sk_scratch(skagrid_ctx *ctx, const char *name, size_t bytes, void **out);
int sk_host_scratch(skagrid_ctx *ctx, size_t bytes, void **out);  // one pinned host buffer per context, grown on demand
// device-resident API: the caller's stream as is; NULL is the CUDA (legacy) default stream, which is what
// torch uses unless told otherwise -- NOT the context's private stream, or launches would race the caller's work
static inline cudaStream_t sk_stream(skagrid_ctx *ctx, void *s) { (void)ctx; return (cudaStream_t)s; }

// ---------------------------------------------------------------------------------------------
// internal device-pointer entry points implemented across the TUs (all asynchronous on `st`)
// ---------------------------------------------------------------------------------------------
int sk_geom_init(skagrid_ctx *ctx, const skagrid_geom *in, i64 capacity, Geom *g);
int sk_plan_alloc(skagrid_ctx *ctx, const skagrid_geom *geom, i64 capacity, int slice_override, skagrid_plan **out);
int sk_plan_fill(skagrid_ctx *ctx, skagrid_plan *p, i64 count, const double *u, const double *v, const i64 *wbin, const double *vis,
                 cudaStream_t st);
void sk_plan_free(skagrid_plan *p);
int sk_take_flags(skagrid_ctx *ctx, cudaStream_t st, uint32_t *flags);

int sk_frac_coord_dev(skagrid_ctx *ctx, i64 n, i64 qpx, i64 count, const double *p, i64 *fl, i64 *frac,
                      int normalise, cudaStream_t st);
int sk_find_closest_dev(skagrid_ctx *ctx, i64 nw, const double *wbins, i64 count, const double *w,
                        i64 *out, cudaStream_t st);
int sk_scale3_dev(skagrid_ctx *ctx, i64 count, double *u, double *v, double *w, double a, int divide,
                  cudaStream_t st);
int sk_mirror_dev(skagrid_ctx *ctx, i64 count, double *u, double *v, double *w, double *vis, cudaStream_t st);
int sk_doweight_dev(skagrid_ctx *ctx, i64 n, double lam, i64 count, const double *u, const double *v,
                    double *vis, uint32_t *err_flag, cudaStream_t st);
int sk_weight_count_dev(skagrid_ctx *ctx, i64 n, double lam, i64 count, const double *u, const double *v, uint32_t *hist, uint32_t *err_flag,
                        cudaStream_t st);
int sk_weight_apply_dev(skagrid_ctx *ctx, i64 n, double lam, i64 count, const double *u, const double *v, const uint32_t *hist, double *vis,
                        cudaStream_t st);
int sk_grid_simple_dev(skagrid_ctx *ctx, i64 h, i64 w, double *grid, i64 count, const double *u,
                       const double *v, const double *vis, cudaStream_t st);
int sk_cmul_dev(skagrid_ctx *ctx, i64 count, double *a, const double *b, cudaStream_t st);
int sk_fill_complex_dev(skagrid_ctx *ctx, i64 count, double *a, double re, double im, cudaStream_t st);

int sk_convolve2d_dev(skagrid_ctx *ctx, i64 n, i64 count, const double *a, const i64 *ai, const double *b,
                      const i64 *bi, double *out, int conj_out, cudaStream_t st);
int sk_aw_kernels_dev(skagrid_ctx *ctx, i64 nw, i64 qpx, i64 s, const double *wkerns, i64 nant,
                      const double *akerns, i64 count, const i64 *wbin, const i64 *yf, const i64 *xf,
                      const i64 *a1, const i64 *a2, double *out, int conj_out, uint32_t *err_flag,
                      cudaStream_t st);

int sk_hermitian_dev(skagrid_ctx *ctx, i64 n, const double *g, double *out, cudaStream_t st);
int sk_fft2c_dev(skagrid_ctx *ctx, i64 n, const double *in, double *out, int inverse, cudaStream_t st);
int sk_grid_to_image_dev(skagrid_ctx *ctx, i64 n, double *grid, double *image, double *max_out,
                         cudaStream_t st);
int sk_slab_fft_rows_dev(skagrid_ctx *ctx, i64 n, i64 row0, i64 nrows, double *slab, cudaStream_t st);
int sk_slab_fft_cols_dev(skagrid_ctx *ctx, i64 n, i64 col0, i64 ncols, double *cols, double *image, double *max_out, cudaStream_t st);
int sk_pad_crop_dev(skagrid_ctx *ctx, i64 n_in, const double *in, i64 n_out, double *out, cudaStream_t st);
int sk_w_kernels_dev(skagrid_ctx *ctx, double theta, i64 nw, const double *w, i64 npixff, i64 npixkern,
                     i64 qpx, int conjugate, double *out, cudaStream_t st);
int sk_w_kernels_ex_dev(skagrid_ctx *ctx, double theta, i64 nw, const double *w, i64 npixff, i64 npixkern, i64 qpx, int conjugate,
                        const double *transmat, double dl, double dm, double *out, cudaStream_t st);

// host-pointer plumbing of api.cu shared with the multi-device entry points (mgpu.cu)
int sk_api_enter(skagrid_ctx *ctx);  // cudaSetDevice + clear the error string
int sk_api_up(skagrid_ctx *ctx, const char *name, const void *host, size_t bytes, void **dev);  // named scratch + H2D on ctx->stream
int sk_api_check_flags(skagrid_ctx *ctx, const char *what);
int sk_api_plan_acquire(skagrid_ctx *ctx, const skagrid_geom *geom, i64 capacity, int slice_override, skagrid_plan **out);
int sk_api_stream_enqueue(skagrid_ctx *ctx, const skagrid_geom *geom, const double *d_table, double *d_grid, i64 count, const double *u,
                          const double *v, const int64_t *wbin, const double *vis, double *vis_out, int degrid, double lam, int want_wbin);
int sk_api_stream_wait(skagrid_ctx *ctx);

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
// frac_coord, src/Gridding.hs:126-140.  Every operation is an explicitly rounded IEEE op so nvcc can
// not contract mul+add into an FMA; the CPU oracle does the same sequence with -ffp-contract=off.
__device__ __forceinline__ void frac_coord_one(double p, double halfnf, double nf, double qpxf,
                                               double qpxfrac, i64 qpx, int normalise, i64 &fl, i64 &fr) {
    const double x = __dadd_rn(halfnf, __dmul_rn(p, nf));
    const double f = floor(__dadd_rn(x, qpxfrac));
    i64 flx = (i64)f;
    const double d = __dsub_rn(x, (double)flx);
    i64 r = (i64)round(__dmul_rn(d, qpxf));  // half away from zero, as libm round
    if (normalise) {
        if (r < 0) { r += qpx; flx -= 1; }
        else if (r >= qpx) { r -= qpx; flx += 1; }
    }
    fl = flx;
    fr = r;
}
#endif
