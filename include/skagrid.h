/*
 * skagrid.h -- C ABI of libskagrid.so, the B200-native (sm_100a) AW-projection gridding hot path.
 *
 * This is the drop-in boundary for the Haskell reference sakehl/SKA-SDP-Accelerate-gridding: every
 * entry point replaces one top-level of src/Gridding.hs / src/ImageDataset.hs that today builds an
 * Accelerate AST and hands it to a backend `run` (src/Gridding.hs:28 `Runners`).  The calling
 * convention mirrors the reference's own FFI idiom (src/Hdf5.hs:30-67 `foreign import ccall`,
 * hdf5/hdf5.cc:59-186 `extern "C"`): flat functions, POD arguments only.
 *
 * Data conventions (src/Types.hs:7-28, src/Hdf5.hs:136,165-167, hdf5/hdf5.cc:14-17):
 *   F            double
 *   complex      interleaved (re, im) doubles; an array of n complex numbers is `double[2n]`
 *   indices      int64_t (Haskell Int / Int64)
 *   arrays       row-major, outermost dimension first, exactly the HDF5 / Accelerate order
 *   uvw          structure of arrays: three `const double*` (what `toForeignPtrs` yields)
 *   kernels      gcf  [qpx, qpx, gh, gw]  (Types.Kernel,   DIM4)   index (yf, xf, i->y, j->x)
 *                wkern[nw, qpx, qpx, s, s] (Types.WKernels, DIM5)
 *                akern[nant, s, s]         (Types.AKernels, DIM3)
 *   grids        [height, width] complex, grid[y, x]; y from v, x from u (src/Gridding.hs:106-109)
 *
 * Ownership: the caller allocates every input and output; the library never frees or keeps caller
 * memory after the call returns.  Device memory lives in the opaque context.
 *
 * Errors: every function returns 0 on success or a negative SKAGRID_E* code; the message is at
 * skagrid_last_error(ctx).  No function aborts the process.  There is NO CPU fallback: without a
 * CUDA device skagrid_create fails.
 *
 * Threading: one context per host thread; calls on a context are serialised by the caller.
 * Haskell should bind with `foreign import ccall safe` (calls block for ms..s).
 *
 * Host-pointer functions ("skagrid_<name>") copy inputs H2D, compute, copy results D2H.
 * Context-resident grid: skagrid_convgrid, skagrid_convgrid2, skagrid_convdegrid, skagrid_convdegrid2,
 * skagrid_conv_imaging2 (grid_out) and skagrid_grid_to_image accept a NULL grid pointer, meaning "the grid the
 * previous of these calls on this context left on the device" (no upload, no download; SKAGRID_EINVAL if there
 * is none of that shape).  A chain conv_imaging2 -> grid_to_image -> convdegrid2 then moves only visibilities
 * over PCIe, as the reference's single fused Accelerate program does.  skagrid_grid_to_image leaves the resident grid
 * as it is when n is even (complex-to-real route); for odd n it transforms in place and the resident buffer then holds the
 * (complex) image plane.
 * Resident coordinates: the same four table functions (and the _mgpu_vis forms) keep the (u, v, wbin) they uploaded on
 * the device (up to 2^28 visibilities; for skagrid_conv_imaging2 the coordinates after the division by lam).  The next
 * call may pass u == v == wbin == NULL with the same count, meaning "at the coordinates of the previous call": an imaging
 * major cycle grids and degrids the same uvw, and then moves 16 instead of 40 bytes per visibility over PCIe
 * (SKAGRID_EINVAL if the count differs or nothing is resident).
 * Device-pointer functions ("skagrid_dev_<name>") work on device-resident buffers on a caller
 * stream and never synchronise the host unless documented (used by bench.py, the multi-GPU host
 * layer, and callers that keep kernels/grids resident across calls).
 */
#ifndef SKAGRID_H
#define SKAGRID_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct skagrid_ctx skagrid_ctx;
typedef struct skagrid_plan skagrid_plan;

enum {
    SKAGRID_OK = 0,
    SKAGRID_EINVAL = -1,   /* bad argument */
    SKAGRID_ECUDA = -2,    /* CUDA / cuFFT runtime error */
    SKAGRID_ENOMEM = -3,   /* device or host allocation failed */
    SKAGRID_ENODEV = -4,   /* no usable CUDA device */
    SKAGRID_ERANGE = -5    /* an index (wbin, antenna, cell) is out of range */
};

/* flags for the binning functions */
enum {
    SKAGRID_FRAC_RAW = 0,       /* literal src/Gridding.hs:126-140 (frac may be -1 or qpx on ties) */
    SKAGRID_FRAC_NORMALISE = 1  /* fold frac into [0,qpx) adjusting the cell (default of the gridders) */
};

/* ------------------------------------------------------------------ context */
int skagrid_create(int device, skagrid_ctx **out);
void skagrid_destroy(skagrid_ctx *ctx);
const char *skagrid_last_error(const skagrid_ctx *ctx);
const char *skagrid_version(void);
/* Grid side N = P.round (theta * lam) (src/Gridding.hs:87, :118, :416, :466, :571): Haskell's round, half to even.  Every
 * imaging entry point below sizes its grid with this; callers use it to allocate. */
int64_t skagrid_grid_side(double theta, int64_t lam);
/* CUDA-event milliseconds of the device work of the last host-pointer call on ctx. */
double skagrid_last_device_ms(const skagrid_ctx *ctx);
/* Device pointer and shape of the grid the last host-pointer call left resident in the context (see "Context-resident
 * grid" above; SKAGRID_EINVAL if there is none).  For one-process-per-GPU callers that sum the per-process grids
 * themselves (an NCCL all-reduce on this buffer between skagrid_conv_imaging2 and skagrid_grid_to_image /
 * skagrid_convdegrid2 with NULL grids); the pointer stays valid until the next call that replaces the resident grid. */
int skagrid_resident_grid(skagrid_ctx *ctx, double **d_grid, int64_t *height, int64_t *width);
/* Number of kernels launched by this context since creation (bench.py "gpu_launches"). */
int64_t skagrid_launch_count(const skagrid_ctx *ctx);
/* Measures the FP64 FMA peak of the device with a register-resident DFMA loop (TFLOP/s). */
int skagrid_measure_fp64_tflops(skagrid_ctx *ctx, double *tflops);
/* Measures the L2 -> SM read bandwidth (TB/s) with 128-bit L2-only loads over a `bytes`-sized, L2-resident buffer:
 * the ceiling of the tiled gridder / degridder, whose kernel taps stream from an L2-resident table. */
int skagrid_measure_l2_read_tbs(skagrid_ctx *ctx, int64_t bytes, double *tbs);
/* The same with an access pattern: 0 = fully coalesced stream, 1 = what the S=15 gridder / degridder do (every
 * half-warp reads the 15 taps of 15 consecutive 256-byte rows of a pseudo-random slice); useful bytes only. */
int skagrid_measure_l2_pattern_tbs(skagrid_ctx *ctx, int64_t bytes, int pattern, double *tbs);

/* ------------------------------------------------------------------ binning (bit-exact)
 * frac_coord  src/Gridding.hs:126-140, frac_coords :142-151 (x,xf from u with width; y,yf from v with height) */
int skagrid_frac_coord(skagrid_ctx *ctx, int64_t n, int64_t qpx, int64_t count, const double *p,
                       int64_t *fl, int64_t *frac, int flags);
int skagrid_frac_coords(skagrid_ctx *ctx, int64_t height, int64_t width, int64_t qpx, int64_t count,
                        const double *u, const double *v, int64_t *x, int64_t *xf, int64_t *y,
                        int64_t *yf, int flags);
/* findClosest src/Gridding.hs:895-907 (w above the last plane clamps to nw-1, Q4). */
int skagrid_find_closest(skagrid_ctx *ctx, int64_t nw, const double *wbins, int64_t count,
                         const double *w, int64_t *out);

/* ------------------------------------------------------------------ pre-steps (in place)
 * uvw_lambda src/ImageDataset.hs:181-187; mirror_uvw src/Gridding.hs:551-562; doweight :564-583 */
int skagrid_uvw_lambda(skagrid_ctx *ctx, double freq, int64_t count, double *u, double *v, double *w);
int skagrid_mirror_uvw(skagrid_ctx *ctx, int64_t count, double *u, double *v, double *w, double *vis);
/* u,v in wavelengths (the function divides by lam itself, as the reference does). vis is divided
 * by the number of visibilities sharing its cell.  Out-of-grid cells -> SKAGRID_ERANGE. */
int skagrid_doweight(skagrid_ctx *ctx, double theta, int64_t lam, int64_t count, const double *u,
                     const double *v, double *vis);

/* ------------------------------------------------------------------ gridding (grid += ...)
 * grid          src/Gridding.hs:95-112   nearest cell
 * convgrid      src/Gridding.hs:153-197  gcf[qpx,qpx,gh,gw]
 * convgrid2     src/Gridding.hs:199-244  gcf[nw,qpx,qpx,gh,gw] + wbin per visibility
 * convgrid_aw   src/Gridding.hs:246-317 (convgrid3) and :318-396 (convgrid4): same result
 * u,v are "p" coordinates (uvw/lam, in (-.5,.5)).  `grid` is read, accumulated into and written back.
 * Out-of-grid taps are dropped (fixoutofbounds, src/Gridding.hs:883-891). */
int skagrid_grid(skagrid_ctx *ctx, int64_t height, int64_t width, double *grid, int64_t count,
                 const double *u, const double *v, const double *vis);
int skagrid_convgrid(skagrid_ctx *ctx, int64_t qpx, int64_t gh, int64_t gw, const double *gcf,
                     int64_t height, int64_t width, double *grid, int64_t count, const double *u,
                     const double *v, const double *vis);
int skagrid_convgrid2(skagrid_ctx *ctx, int64_t nw, int64_t qpx, int64_t gh, int64_t gw,
                      const double *gcf, int64_t height, int64_t width, double *grid, int64_t count,
                      const double *u, const double *v, const int64_t *wbin, const double *vis);
int skagrid_convgrid_aw(skagrid_ctx *ctx, int64_t nw, int64_t qpx, int64_t s, const double *wkerns,
                        int64_t nant, const double *akerns, int64_t height, int64_t width,
                        double *grid, int64_t count, const double *u, const double *v,
                        const int64_t *wbin, const int64_t *a1, const int64_t *a2, const double *vis);

/* ------------------------------------------------------------------ degridding (new; exact adjoints)
 * vis_out[k] = sum_{i,j} conj(c_k[i,j]) * grid[y_k - gh/2 + i, x_k - gw/2 + j], c_k being the factor the
 * matching gridder multiplies vis_k with (gcf slice, or conj(AW_k)).  Not in the reference. */
int skagrid_convdegrid(skagrid_ctx *ctx, int64_t qpx, int64_t gh, int64_t gw, const double *gcf,
                       int64_t height, int64_t width, const double *grid, int64_t count,
                       const double *u, const double *v, double *vis_out);
int skagrid_convdegrid2(skagrid_ctx *ctx, int64_t nw, int64_t qpx, int64_t gh, int64_t gw,
                        const double *gcf, int64_t height, int64_t width, const double *grid,
                        int64_t count, const double *u, const double *v, const int64_t *wbin,
                        double *vis_out);
int skagrid_convdegrid_aw(skagrid_ctx *ctx, int64_t nw, int64_t qpx, int64_t s, const double *wkerns,
                          int64_t nant, const double *akerns, int64_t height, int64_t width,
                          const double *grid, int64_t count, const double *u, const double *v,
                          const int64_t *wbin, const int64_t *a1, const int64_t *a2, double *vis_out);

/* ------------------------------------------------------------------ AW kernel formation
 * convolve2d src/Gridding.hs:795-811 (direct evaluation of its FFT route, incl. the padder transpose);
 * aw_kernel = aw_kernel_fn2 src/Gridding.hs:761-775 for `count` (wbin,yf,xf,a1,a2) tuples -> out[count,s,s]
 * (NOT conjugated: conjugation happens in processOne2, src/Gridding.hs:391). */
int skagrid_convolve2d(skagrid_ctx *ctx, int64_t n, const double *a1, const double *a2, double *out);
int skagrid_aw_kernel(skagrid_ctx *ctx, int64_t nw, int64_t qpx, int64_t s, const double *wkerns,
                      int64_t nant, const double *akerns, int64_t count, const int64_t *wbin,
                      const int64_t *yf, const int64_t *xf, const int64_t *a1, const int64_t *a2,
                      double *out);

/* ------------------------------------------------------------------ grid -> image
 * make_grid_hermitian src/Gridding.hs:585-605; ifft :828-829 (centred, 1/N^2); fft :821-826 (centred,
 * zero-padded to the next power of two and cropped back, as the reference does). n x n complex. */
int skagrid_make_grid_hermitian(skagrid_ctx *ctx, int64_t n, const double *grid, double *out);
int skagrid_ifft(skagrid_ctx *ctx, int64_t n, const double *grid, double *out);
int skagrid_fft(skagrid_ctx *ctx, int64_t n, const double *grid, double *out);
/* Fused make_grid_hermitian -> ifft -> real -> maximum (src/ImageDataset.hs:74-77).
 * image may be NULL when only the maximum is wanted. */
int skagrid_grid_to_image(skagrid_ctx *ctx, int64_t n, const double *grid, double *image, double *max_out);

/* ------------------------------------------------------------------ imaging drivers
 * simple_imaging src/Gridding.hs:84-93; conv_imaging :115-124; aw_imaging :452-478 (aw_imagingOld
 * :480-506 gives the same grid).  u,v,w in wavelengths; grid_out is n x n with n = round(theta*lam). */
int skagrid_simple_imaging(skagrid_ctx *ctx, double theta, int64_t lam, int64_t count, const double *u,
                           const double *v, const double *w, const double *vis, double *grid_out);
int skagrid_conv_imaging(skagrid_ctx *ctx, int64_t qpx, int64_t gh, int64_t gw, const double *gcf,
                         double theta, int64_t lam, int64_t count, const double *u, const double *v,
                         const double *w, const double *vis, double *grid_out);
/* conv_imaging with a w-indexed table [nw,qpx,qpx,gh,gw] + wbin: the last step of w_cache_imaging
 * (src/Gridding.hs:421-449: zero grid, p = uvw/lam, convgrid2). */
int skagrid_conv_imaging2(skagrid_ctx *ctx, int64_t nw, int64_t qpx, int64_t gh, int64_t gw, const double *gcf,
                          double theta, int64_t lam, int64_t count, const double *u, const double *v,
                          const double *w, const int64_t *wbin, const double *vis, double *grid_out);
int skagrid_aw_imaging(skagrid_ctx *ctx, double theta, int64_t lam, int64_t nw, int64_t qpx, int64_t s,
                       const double *wkerns, const double *wbins, int64_t nant, const double *akerns,
                       int64_t count, const double *u, const double *v, const double *w,
                       const int64_t *a1, const int64_t *a2, const double *vis, double *grid_out);
/* ImageDataset.aw_gridding from the loaded arrays on (src/ImageDataset.hs:47-77): uvw in METRES,
 * uvw_lambda(freq), doweight on the un-mirrored uvw, mirror_uvw, aw_imaging, make_grid_hermitian,
 * real(ifft), maximum.  image (n x n doubles) and grid_out (n x n complex) may each be NULL. */
int skagrid_aw_gridding(skagrid_ctx *ctx, double theta, int64_t lam, int64_t nw, int64_t qpx, int64_t s,
                        const double *wkerns, const double *wbins, int64_t nant, const double *akerns,
                        int64_t count, const double *u_m, const double *v_m, const double *w_m,
                        const int64_t *a1, const int64_t *a2, double freq, const double *vis,
                        double *image, double *max_out, double *grid_out);

/* ------------------------------------------------------------------ w-kernel generation
 * w_kernel src/Gridding.hs:610-728 for `nw` w values -> out[nw,qpx,qpx,npixkern,npixkern];
 * conjugate != 0 applies the `map conjugate` of w_cache_imaging (src/Gridding.hs:441). */
int skagrid_w_kernels(skagrid_ctx *ctx, double theta, int64_t nw, const double *w, int64_t npixff,
                      int64_t npixkern, int64_t qpx, int conjugate, double *out);
/* The same with the KernelOptions that move the far-field coordinates (kernel_coordinates, src/Gridding.hs:620-635):
 * (l, m) = theta * coordinates2 -> (t00 l + t10 m, t01 l + t11 m) -> + (dl, dm).  transmat = patTransMat as 4 doubles,
 * row-major t[r][c] (NULL: identity); dl, dm = patHorShift, patVerShift. */
int skagrid_w_kernels_ex(skagrid_ctx *ctx, double theta, int64_t nw, const double *w, int64_t npixff,
                         int64_t npixkern, int64_t qpx, int conjugate, const double *transmat, double dl, double dm,
                         double *out);

/* ================================================================== multi-GPU, single process
 * ONE host thread drives `nctx` (1..16) contexts, one per device (ctxs[i] from skagrid_create(device_i); contexts on the
 * same device are allowed).  Everything is enqueued on the contexts' streams, so the devices work concurrently; the
 * call returns when all of them are done.  Arguments as skagrid_convgrid2 / skagrid_convdegrid2 (nw = 1 and
 * wbin = NULL give convgrid / convdegrid).  Errors of any context are reported through ctxs[0];
 * skagrid_last_device_ms(ctxs[0]) covers the whole call.
 *
 * _vis  visibility-sharded (BASELINE config 4): context d takes the d-th contiguous share of the visibilities.
 *       Gridding: full local grids, then a reduce-scatter over NVLink peer memory (context d sums row slab d of every
 *       peer), the slabs return to `grid` (in/out, accumulated into) over nctx PCIe links in parallel, and an
 *       all-gather leaves the sum RESIDENT on every context.  grid == NULL: start from zero, no download.
 *       Degridding: every context uploads one row slab of `grid`, the rest arrives by all-gather; grid == NULL uses
 *       the resident grids.  The result differs from the single-device one only by summation order.
 * _tile uv-tile-sharded (BASELINE config 5): context d owns grid rows [bounds[d], bounds[d+1]), chosen at the d/nctx
 *       quantiles of this batch's footprint rows; visibilities are routed device-to-device to every owner their
 *       footprint intersects, owners clip taps to their rows (fixoutofbounds, src/Gridding.hs:883-891); no grid
 *       reduction; degridding returns the owners' partial sums to the source device and adds them.  No context ever
 *       holds more than its slab of the grid.  bounds_out (nctx+1 values, may be NULL) receives the bounds used. */
int skagrid_convgrid2_mgpu_vis(skagrid_ctx *const *ctxs, int nctx, int64_t nw, int64_t qpx, int64_t gh, int64_t gw,
                               const double *gcf, int64_t height, int64_t width, double *grid, int64_t count,
                               const double *u, const double *v, const int64_t *wbin, const double *vis);
int skagrid_convdegrid2_mgpu_vis(skagrid_ctx *const *ctxs, int nctx, int64_t nw, int64_t qpx, int64_t gh, int64_t gw,
                                 const double *gcf, int64_t height, int64_t width, const double *grid, int64_t count,
                                 const double *u, const double *v, const int64_t *wbin, double *vis_out);
int skagrid_convgrid2_mgpu_tile(skagrid_ctx *const *ctxs, int nctx, int64_t nw, int64_t qpx, int64_t gh, int64_t gw,
                                const double *gcf, int64_t height, int64_t width, double *grid, int64_t count,
                                const double *u, const double *v, const int64_t *wbin, const double *vis,
                                int64_t *bounds_out);
int skagrid_convdegrid2_mgpu_tile(skagrid_ctx *const *ctxs, int nctx, int64_t nw, int64_t qpx, int64_t gh, int64_t gw,
                                  const double *gcf, int64_t height, int64_t width, const double *grid, int64_t count,
                                  const double *u, const double *v, const int64_t *wbin, double *vis_out,
                                  int64_t *bounds_out);

/* ================================================================== device-resident API
 * All pointers are DEVICE pointers on ctx's device; `stream` is a cudaStream_t passed as void*
 * (NULL = the CUDA legacy default stream).  Work is ordered on that stream only; nothing here
 * synchronises the host except where stated. */

typedef struct skagrid_geom {
    int64_t height, width;   /* full grid size (binning uses these) */
    int64_t row0, row1;      /* rows [row0,row1) are owned: `grid` points at row0 and has row1-row0
                                rows; taps outside are dropped (uv-tile-sharded mode). 0,height = all */
    int64_t nw, qpx, gh, gw; /* kernel table shape [nw,qpx,qpx,gh,gw] (nw = 1 for convgrid) */
} skagrid_geom;

/* A plan holds `count` visibilities binned (bit-exact) and bucketed by (uv tile, 2x2 micro-tile).
 * wbin may be NULL (all 0); vis may be NULL for a degrid-only plan.  slice_override != 0: the kernel slice of visibility k is k itself
 * (per-visibility kernels, the AW path) instead of (wbin,yf,xf).  Synchronises `stream` once. */
int skagrid_dev_plan_create(skagrid_ctx *ctx, const skagrid_geom *geom, int64_t count, const double *u,
                            const double *v, const int64_t *wbin, const double *vis, int slice_override,
                            void *stream, skagrid_plan **out);
void skagrid_dev_plan_destroy(skagrid_ctx *ctx, skagrid_plan *plan);
/* Re-bin and re-bucket a new batch of at most the plan's capacity into an existing plan (no
 * allocation, no host synchronisation). */
int skagrid_dev_plan_update(skagrid_ctx *ctx, skagrid_plan *plan, int64_t count, const double *u,
                            const double *v, const int64_t *wbin, const double *vis, void *stream);
/* An empty plan with room for `capacity` visibilities, and the update from records of `width` doubles
 * {u, v, wbin (int64 bits), [re, im]} (width 5; 3 for a degrid-only plan) -- the layout skagrid_dev_route_pack produces,
 * so routed visibilities are binned where the exchange left them. */
int skagrid_dev_plan_alloc(skagrid_ctx *ctx, const skagrid_geom *geom, int64_t capacity, int slice_override,
                           skagrid_plan **out);
int skagrid_dev_plan_update_packed(skagrid_ctx *ctx, skagrid_plan *plan, int64_t count, const double *d_rec,
                                   int width, void *stream);
/* New visibility values at the coordinates the plan was built from: the records keep their place, only their visibilities
 * are refreshed.  For major cycles over the same uvw: the binning and the bucket sort are done once per data set.
 * in_plan_order == 0: d_vis holds `count` complex numbers in the caller's order (one random 16-byte read per record);
 * != 0: d_vis[r] belongs to record r -- the caller permuted its data once with plan_order (d_index[r] = position of record
 * r's visibility in the caller's arrays, r < plan_stats[0]) and the refresh is a sequential pass. */
int skagrid_dev_plan_set_vis(skagrid_ctx *ctx, skagrid_plan *plan, const double *d_vis, int in_plan_order, void *stream);
int skagrid_dev_plan_order(skagrid_ctx *ctx, skagrid_plan *plan, uint32_t *d_index, void *stream);
/* Plan statistics: [0] visibilities kept, [1] dropped (no tap on the owned rows), [2] work items,
 * [3] uv tiles, [4] non-empty tiles.  Synchronises `stream`. */
int skagrid_dev_plan_stats(skagrid_ctx *ctx, skagrid_plan *plan, void *stream, int64_t stats[5]);

/* grid[row0:row1, :] += sum over the plan's visibilities of vis_k * table[slice_k].
 * variant: 0 = tiled shared-memory gridder (default), 1 = one-thread-per-tap global-atomic gridder
 * (the literal `permute (+)`; baseline and cross-check), 2..4 = earlier thread layouts / pipeline depths of
 * the tiled gridder kept for A/B measurements (see gridder.cu). */
int skagrid_dev_grid(skagrid_ctx *ctx, skagrid_plan *plan, const double *table, double *grid,
                     int variant, void *stream);
/* vis_out[k] = sum conj(table[slice_k][i,j]) * grid[...] for the plan's visibilities (others = 0). */
int skagrid_dev_degrid(skagrid_ctx *ctx, skagrid_plan *plan, const double *table, const double *grid,
                       double *vis_out, void *stream);
/* The same with the results in the plan's own order: vis_out[r] belongs to record r (r < plan_stats[0], see plan_order) --
 * sequential full-sector writes instead of one 16-byte store at a random index per visibility. */
int skagrid_dev_degrid_plan_order(skagrid_ctx *ctx, skagrid_plan *plan, const double *table, const double *grid,
                                  double *vis_out, void *stream);
/* Grid -> image stage on an n x n device grid: hermitian + centred inverse FFT; writes real(image) into image (n*n
 * doubles, may be NULL) and the maximum into max_out (1 double).  Even n: only the real part is wanted, so the hermitian
 * half of the spectrum goes through a complex-to-REAL transform (half the passes of the complex one) and `grid` is not
 * modified; odd n: `grid` is transformed in place. */
int skagrid_dev_grid_to_image(skagrid_ctx *ctx, int64_t n, double *grid, double *image,
                              double *max_out, void *stream);
/* Synthetic SKA1-Low-shaped visibilities generated on the device: counter-based
 * splitmix64(seed, index).  Fills u,v (p coordinates, v>=0 after mirroring), w-plane index, vis. */
int skagrid_dev_synth_vis(skagrid_ctx *ctx, uint64_t seed, int64_t first, int64_t count, int64_t n,
                          int64_t support, int64_t nw, int uniform, double *u, double *v,
                          int64_t *wbin, double *vis, void *stream);

/* Pre-steps on device arrays (same kernels as the host-pointer functions): (u,v,w) *= a or /= a (uvw_lambda with
 * a = f/299792458.0, src/ImageDataset.hs:181-187; div3 with a = lam, src/Gridding.hs:838-839), mirror_uvw (:551-562;
 * d_vis may be NULL), findClosest (:895-907), doweight (:564-583).  Range problems (a visibility outside the weight
 * grid, an index out of range in a plan) are recorded in the context's device error word: skagrid_dev_take_error
 * reads and clears it (bit 0: index out of range, bit 1: outside the weight grid) and synchronises `stream`. */
int skagrid_dev_uvw_scale(skagrid_ctx *ctx, int64_t count, double *d_u, double *d_v, double *d_w, double a,
                          int divide, void *stream);
int skagrid_dev_mirror_uvw(skagrid_ctx *ctx, int64_t count, double *d_u, double *d_v, double *d_w, double *d_vis,
                           void *stream);
int skagrid_dev_find_closest(skagrid_ctx *ctx, int64_t nw, const double *d_wbins, int64_t count, const double *d_w,
                             int64_t *d_out, void *stream);
int skagrid_dev_doweight(skagrid_ctx *ctx, double theta, int64_t lam, int64_t count, const double *d_u,
                         const double *d_v, double *d_vis, void *stream);
/* Grid -> image when the n x n grid (n even) is spread over devices as row slabs (uv-tile-sharded gridding), without
 * gathering it.  Stage 1, in place on this device's rows [row0, row0+nrows): the hermitian weighting (real(ifft(
 * make_grid_hermitian g)) == real(ifft(g scaled by 2 except on row 0 and column 0)), so no mirrored rows are exchanged),
 * the centring factor and the inverse FFT along x.  The caller then transposes across devices (all-to-all: device h
 * receives columns [col0, col0+ncols) of every row) into an [n, ncols] row-major array.  Stage 2, in place on that array:
 * inverse FFT along y; d_image[n, ncols] (may be NULL) = real part, centred, 1/n^2; d_max (1 double, may be NULL) = the
 * maximum pixel of this column slab.  src/Gridding.hs:585-605, :828-829; src/ImageDataset.hs:74-77. */
int skagrid_dev_slab_fft_rows(skagrid_ctx *ctx, int64_t n, int64_t row0, int64_t nrows, double *d_slab, void *stream);
int skagrid_dev_slab_fft_cols(skagrid_ctx *ctx, int64_t n, int64_t col0, int64_t ncols, double *d_cols,
                              double *d_image, double *d_max, void *stream);
/* doweight in two phases, for visibilities sharded over devices (the weight grid is a sum, src/Gridding.hs:580):
 * every device adds the cell counts of its share to d_hist (n x n int32, n = round(theta*lam), zeroed by the caller), the
 * caller sums the histograms across devices (one all-reduce), then every device divides its share by the summed counts. */
int skagrid_dev_weight_count(skagrid_ctx *ctx, double theta, int64_t lam, int64_t count, const double *d_u,
                             const double *d_v, int32_t *d_hist, void *stream);
int skagrid_dev_weight_apply(skagrid_ctx *ctx, double theta, int64_t lam, int64_t count, const double *d_u,
                             const double *d_v, const int32_t *d_hist, double *d_vis, void *stream);
int skagrid_dev_take_error(skagrid_ctx *ctx, void *stream, int *flags_out);
/* uv-tile-sharded mode with one process per GPU (the exchange itself is the caller's: NCCL all-to-all).
 * A visibility covers grid rows [y - gh/2, y - gh/2 + gh), y = frac_coord(height, qpx, v) bit-exact (src/Gridding.hs:126-140),
 * and goes to every rank g whose slab [bounds[g], bounds[g+1]) (host array, nranks + 1 entries from 0 to height)
 * intersects them; the owner clips taps to its slab (fixoutofbounds, src/Gridding.hs:883-891).
 *   row_hist     d_hist[row of the footprint centre] += 1 (uint32, height entries): the slab balance
 *   route_count  d_counts[g] (uint32, nranks entries, zeroed by the call) = records this device sends to rank g
 *   route_pack   appends the records, destination-major, to d_send: rank g's segment starts at record seg[g] (host array);
 *                a record is 5 doubles {u, v, wbin, re, im}, or 3 when d_vis == NULL; d_sidx (may be NULL) receives
 *                the source index of every record.  Must follow route_count of the same (count, d_v, bounds) on this
 *                context: it places the records with the per-block offsets that call left behind (no global atomics)
 *   scatter_add  d_out[d_sidx[i]] += d_back[i] (complex): returned degridding partial sums, one per routed record */
int skagrid_dev_row_hist(skagrid_ctx *ctx, int64_t height, int64_t qpx, int64_t gh, int64_t count, const double *d_v,
                         uint32_t *d_hist, void *stream);
int skagrid_dev_route_count(skagrid_ctx *ctx, int64_t height, int64_t qpx, int64_t gh, int nranks, const int64_t *bounds,
                            int64_t count, const double *d_v, uint32_t *d_counts, void *stream);
int skagrid_dev_route_pack(skagrid_ctx *ctx, int64_t height, int64_t qpx, int64_t gh, int nranks, const int64_t *bounds,
                           int64_t count, const double *d_u, const double *d_v, const int64_t *d_wbin, const double *d_vis,
                           const int64_t *seg, double *d_send, uint32_t *d_sidx, void *stream);
int skagrid_dev_scatter_add(skagrid_ctx *ctx, int64_t n, const uint32_t *d_sidx, const double *d_back, double *d_out,
                            void *stream);
/* Peer-memory exchange between one-process-per-GPU ranks over NVLink (csrc/ipc.cu), no collective library in the data path.
 * ipc_alloc: a device buffer (zero-filled) plus its 64-byte CUDA IPC handle; the ranks exchange the handles by any means
 * and ipc_open each other's (peer access between the devices is required); ipc_close / ipc_free release them.
 *   peer_sum      d_own[i] += sum_k d_peers[k][i] over `ncomplex` complex values; d_peers is a HOST array of opened peer
 *                 pointers: the reduce-scatter of the visibility-sharded mode (permute (+) is a sum, src/Gridding.hs:377).
 *                 broadcast != 0: the sum is also stored to every d_peers[k][i] by the same kernel (reduce-scatter and
 *                 all-gather fused; every rank must call it on ITS slab between two barriers)
 *   peer_barrier  stream-ordered barrier of `nranks` ranks: d_flags is a HOST array with every rank's flag buffer (at least
 *                 64 uint32, ipc_alloc'ed; entry `rank` is the local one), epoch increases by one per barrier on every
 *                 rank.  A peer that does not arrive within ~30 s sets bit 2 of the device error word (dev_take_error)
 *   peer_copy     cudaMemcpyAsync between local / opened peer memory (copy engines); peer_copy2d the strided form
 *                 (width_bytes x rows, pitches in bytes)
 *   peer_gather   the same for up to 64 (destination, source, bytes) segments in ONE kernel: the SMs pull from peer memory
 *                 (8-byte aligned segments; HOST arrays of pointers / sizes); peer_gather2d the strided form for 16-byte
 *                 elements, all segments with the same width and destination pitch (the transpose of the slab image).
 *                 max_blocks > 0 caps the kernel's grid (one block per SM, say) so that kernels on another stream run
 *                 beside the exchange; 0 = fill the device */
int skagrid_ipc_alloc(skagrid_ctx *ctx, int64_t bytes, void **d_ptr, unsigned char handle[64]);
int skagrid_ipc_free(skagrid_ctx *ctx, void *d_ptr);
int skagrid_ipc_open(skagrid_ctx *ctx, const unsigned char handle[64], void **d_ptr);
int skagrid_ipc_close(skagrid_ctx *ctx, void *d_ptr);
int skagrid_dev_peer_sum(skagrid_ctx *ctx, int npeers, double *const *d_peers, double *d_own, int64_t ncomplex,
                         int broadcast, void *stream);
int skagrid_dev_peer_barrier(skagrid_ctx *ctx, int nranks, int rank, uint32_t *const *d_flags, uint32_t epoch,
                             void *stream);
int skagrid_dev_peer_copy(skagrid_ctx *ctx, void *d_dst, const void *d_src, int64_t bytes, void *stream);
int skagrid_dev_peer_gather(skagrid_ctx *ctx, int nseg, void *const *d_dst, const void *const *d_src, const int64_t *bytes,
                            int max_blocks, void *stream);
int skagrid_dev_peer_gather2d(skagrid_ctx *ctx, int nseg, void *const *d_dst, int64_t dpitch, const void *const *d_src,
                              const int64_t *spitch, int64_t width_bytes, const int64_t *rows, int max_blocks, void *stream);
int skagrid_dev_peer_copy2d(skagrid_ctx *ctx, void *d_dst, int64_t dpitch, const void *d_src, int64_t spitch,
                            int64_t width_bytes, int64_t rows, void *stream);
/* frac_coord (src/Gridding.hs:126-140) on device arrays. */
int skagrid_dev_frac_coord(skagrid_ctx *ctx, int64_t n, int64_t qpx, int64_t count, const double *d_p,
                           int64_t *d_fl, int64_t *d_frac, int flags, void *stream);
/* w_kernel table built directly into device memory (w values on the host). */
int skagrid_dev_w_kernels(skagrid_ctx *ctx, double theta, int64_t nw, const double *w_host, int64_t npixff,
                          int64_t npixkern, int64_t qpx, int conjugate, double *d_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SKAGRID_H */
