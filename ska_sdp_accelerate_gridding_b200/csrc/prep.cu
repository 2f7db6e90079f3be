// prep.cu -- the element-wise pre-steps of ImageDataset.aw_gridding and the nearest-cell gridder.
//
//   uvw_lambda  src/ImageDataset.hs:181-187   (u,v,w) * (f / 299792458.0), scalar formed on the host
//   div3        src/Gridding.hs:838-839       true division by lam
//   mirror_uvw  src/Gridding.hs:551-562       v < 0 -> (-u,-v,-w), conj(vis)
//   doweight    src/Gridding.hs:564-583       per-cell visibility count (permute (+) of ones), vis /= count
//   grid        src/Gridding.hs:95-112        nearest cell scatter-add
//
// Every floating-point step that feeds an integer (cell index) uses explicitly rounded IEEE operations
// (__dmul_rn / __dadd_rn / __ddiv_rn) so that nvcc cannot contract them into FMAs: the indices must be
// bit-identical to the reference semantics.
#include "common.cuh"

__global__ void __launch_bounds__(256) scale3_kernel(i64 count, double *__restrict__ u, double *__restrict__ v, double *__restrict__ w,
                                                     double a, int divide) {
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) {
        if (divide) {
            u[k] = __ddiv_rn(u[k], a); v[k] = __ddiv_rn(v[k], a); w[k] = __ddiv_rn(w[k], a);
        } else {
            u[k] = __dmul_rn(a, u[k]); v[k] = __dmul_rn(a, v[k]); w[k] = __dmul_rn(a, w[k]);
        }
    }
}

static int blocks_for(skagrid_ctx *ctx, i64 count) {
    i64 b = (count + 255) / 256;
    const i64 cap = (i64)ctx->sm_count * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

int sk_scale3_dev(skagrid_ctx *ctx, i64 count, double *u, double *v, double *w, double a, int divide, cudaStream_t st) {
    if (count <= 0) return SKAGRID_OK;
    scale3_kernel<<<blocks_for(ctx, count), 256, 0, st>>>(count, u, v, w, a, divide);
    SK_LAUNCH_CHECK(ctx);
    return SKAGRID_OK;
}

__global__ void __launch_bounds__(256) mirror_kernel(i64 count, double *__restrict__ u, double *__restrict__ v, double *__restrict__ w,
                                                     double *__restrict__ vis) {
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) {
        if (v[k] < 0.0) {
            u[k] = -u[k]; v[k] = -v[k]; w[k] = -w[k];
            if (vis) vis[2 * k + 1] = -vis[2 * k + 1];
        }
    }
}

int sk_mirror_dev(skagrid_ctx *ctx, i64 count, double *u, double *v, double *w, double *vis, cudaStream_t st) {
    if (count <= 0) return SKAGRID_OK;
    mirror_kernel<<<blocks_for(ctx, count), 256, 0, st>>>(count, u, v, w, vis);
    SK_LAUNCH_CHECK(ctx);
    return SKAGRID_OK;
}

// doweight: cell of visibility k = frac_coords (n,n) 1 (u/lam, v/lam); only the integer part is used.
__device__ __forceinline__ bool weight_cell(i64 n, double halfnf, double nf, double lam, double u, double v, i64 &cell) {
    i64 x, xf, y, yf;
    frac_coord_one(__ddiv_rn(u, lam), halfnf, nf, 1.0, 0.5, 1, 0, x, xf);
    frac_coord_one(__ddiv_rn(v, lam), halfnf, nf, 1.0, 0.5, 1, 0, y, yf);
    if (x < 0 || y < 0 || x >= n || y >= n) return false;
    cell = y * n + x;
    return true;
}

__global__ void __launch_bounds__(256) weight_hist_kernel(i64 n, double lam, i64 count, const double *__restrict__ u,
                                                          const double *__restrict__ v, uint32_t *__restrict__ hist,
                                                          uint32_t *__restrict__ err_flag) {
    const double halfnf = (double)(n / 2), nf = (double)n;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) {
        i64 cell;
        if (!(fabs(u[k]) < 1e300) || !(fabs(v[k]) < 1e300) || !weight_cell(n, halfnf, nf, lam, u[k], v[k], cell)) { atomicOr(err_flag, 2u); continue; }
        atomicAdd(&hist[cell], 1u);
    }
}

__global__ void __launch_bounds__(256) weight_apply_kernel(i64 n, double lam, i64 count, const double *__restrict__ u,
                                                           const double *__restrict__ v, const uint32_t *__restrict__ hist,
                                                           double *__restrict__ vis) {
    const double halfnf = (double)(n / 2), nf = (double)n;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) {
        i64 cell;
        if (!(fabs(u[k]) < 1e300) || !(fabs(v[k]) < 1e300) || !weight_cell(n, halfnf, nf, lam, u[k], v[k], cell)) continue;
        const double wgt = (double)hist[cell];  // the reference accumulates 1.0 in a Double grid: exact
        vis[2 * k] = __ddiv_rn(vis[2 * k], wgt);
        vis[2 * k + 1] = __ddiv_rn(vis[2 * k + 1], wgt);
    }
}

int sk_doweight_dev(skagrid_ctx *ctx, i64 n, double lam, i64 count, const double *u, const double *v, double *vis, uint32_t *err_flag,
                    cudaStream_t st) {
    if (count <= 0) return SKAGRID_OK;
    if (n <= 0 || n > 65536) return sk_fail(ctx, SKAGRID_EINVAL, "doweight: grid side %lld out of range", n);
    void *hist;
    SK_TRY(sk_scratch(ctx, "weight_hist", (size_t)(n * n) * sizeof(uint32_t), &hist));
    SK_CUDA(ctx, cudaMemsetAsync(hist, 0, (size_t)(n * n) * sizeof(uint32_t), st));
    weight_hist_kernel<<<blocks_for(ctx, count), 256, 0, st>>>(n, lam, count, u, v, (uint32_t *)hist, err_flag);
    SK_LAUNCH_CHECK(ctx);
    weight_apply_kernel<<<blocks_for(ctx, count), 256, 0, st>>>(n, lam, count, u, v, (const uint32_t *)hist, vis);
    SK_LAUNCH_CHECK(ctx);
    return SKAGRID_OK;
}

// The two phases of doweight on their own, for visibilities sharded over devices: every device counts its share into
// hist (n x n uint32, accumulated), the counts are summed across devices (doweight's `permute (+)` is a sum,
// src/Gridding.hs:580), then every device divides its share.
int sk_weight_count_dev(skagrid_ctx *ctx, i64 n, double lam, i64 count, const double *u, const double *v, uint32_t *hist, uint32_t *err_flag,
                        cudaStream_t st) {
    if (count <= 0) return SKAGRID_OK;
    if (n <= 0 || n > 65536) return sk_fail(ctx, SKAGRID_EINVAL, "doweight: grid side %lld out of range", n);
    weight_hist_kernel<<<blocks_for(ctx, count), 256, 0, st>>>(n, lam, count, u, v, hist, err_flag);
    SK_LAUNCH_CHECK(ctx);
    return SKAGRID_OK;
}
int sk_weight_apply_dev(skagrid_ctx *ctx, i64 n, double lam, i64 count, const double *u, const double *v, const uint32_t *hist, double *vis,
                        cudaStream_t st) {
    if (count <= 0) return SKAGRID_OK;
    if (n <= 0 || n > 65536) return sk_fail(ctx, SKAGRID_EINVAL, "doweight: grid side %lld out of range", n);
    weight_apply_kernel<<<blocks_for(ctx, count), 256, 0, st>>>(n, lam, count, u, v, hist, vis);
    SK_LAUNCH_CHECK(ctx);
    return SKAGRID_OK;
}

// grid (src/Gridding.hs:95-112): n = number of rows for BOTH coordinates; cell = n/2 + floor(0.5 + n*p).
__global__ void __launch_bounds__(256) grid_simple_kernel(i64 h, i64 w, double *__restrict__ grid, i64 count, const double *__restrict__ u,
                                                          const double *__restrict__ v, const double *__restrict__ vis) {
    const i64 halfn = h / 2;
    const double nf = (double)h;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) {
        const double pu = __dmul_rn(nf, u[k]), pv = __dmul_rn(nf, v[k]);
        if (!(fabs(pu) < 1e15) || !(fabs(pv) < 1e15)) continue;
        const i64 x = halfn + (i64)floor(__dadd_rn(0.5, pu));
        const i64 y = halfn + (i64)floor(__dadd_rn(0.5, pv));
        if (x < 0 || y < 0 || x >= w || y >= h) continue;
        double *g = grid + 2 * (y * w + x);
        atomicAdd(g, vis[2 * k]);
        atomicAdd(g + 1, vis[2 * k + 1]);
    }
}

int sk_grid_simple_dev(skagrid_ctx *ctx, i64 h, i64 w, double *grid, i64 count, const double *u, const double *v, const double *vis,
                       cudaStream_t st) {
    if (count <= 0) return SKAGRID_OK;
    grid_simple_kernel<<<blocks_for(ctx, count), 256, 0, st>>>(h, w, grid, count, u, v, vis);
    SK_LAUNCH_CHECK(ctx);
    return SKAGRID_OK;
}

// a[k] *= b[k] (complex), used for vis * weight (src/ImageDataset.hs:72).
__global__ void __launch_bounds__(256) cmul_kernel(i64 count, double2 *__restrict__ a, const double2 *__restrict__ b) {
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) {
        const double2 x = a[k], y = b[k];
        a[k] = make_double2(__dsub_rn(__dmul_rn(x.x, y.x), __dmul_rn(x.y, y.y)), __dadd_rn(__dmul_rn(x.x, y.y), __dmul_rn(x.y, y.x)));
    }
}

int sk_cmul_dev(skagrid_ctx *ctx, i64 count, double *a, const double *b, cudaStream_t st) {
    if (count <= 0) return SKAGRID_OK;
    cmul_kernel<<<blocks_for(ctx, count), 256, 0, st>>>(count, (double2 *)a, (const double2 *)b);
    SK_LAUNCH_CHECK(ctx);
    return SKAGRID_OK;
}

// a[k] = re + i im (the `ones` vector of src/ImageDataset.hs:58 is built on the device)
__global__ void __launch_bounds__(256) fill_complex_kernel(i64 count, double2 *__restrict__ a, double re, double im) {
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) a[k] = make_double2(re, im);
}

int sk_fill_complex_dev(skagrid_ctx *ctx, i64 count, double *a, double re, double im, cudaStream_t st) {
    if (count <= 0) return SKAGRID_OK;
    fill_complex_kernel<<<blocks_for(ctx, count), 256, 0, st>>>(count, (double2 *)a, re, im);
    SK_LAUNCH_CHECK(ctx);
    return SKAGRID_OK;
}
