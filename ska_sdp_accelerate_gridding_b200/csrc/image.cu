// image.cu -- grid -> image stage and w-kernel generation.
//
//   make_grid_hermitian  src/Gridding.hs:585-605
//   ifft / fft           src/Gridding.hs:828-829 / :821-826   shift2D . fft2D . ishift2D (accelerate-fft)
//   map real, maximum    src/ImageDataset.hs:76-77
//   w_kernel             src/Gridding.hs:610-728
//
// The reference delegates the 2-D transform to a library (FFTW / cuFFT through accelerate-fft); so does
// this file (cuFFT Z2Z).  What is hand-written is everything around it, fused to the minimum number of
// passes over the N x N complex grid (HBM-bound, 16 B per cell per pass):
//   pass 1  hermitian symmetrisation + (-1)^(x+y) modulation, in place, pairwise (one thread updates a
//           cell and its mirror, so no second buffer is needed);
//   cuFFT   in place;
//   pass 2  (-1)^(x+y) / N^2, real part, block maximum -> one atomic per block.
// For even N, shift2D . ifft2 . ishift2D (G) == M .* ifft2(M .* G), M = (-1)^(x+y); odd N uses
// explicit rotations.
#include <cstdlib>

#include "common.cuh"

static const char *cufft_str(cufftResult r) {
    switch (r) {
        case CUFFT_SUCCESS: return "CUFFT_SUCCESS";
        case CUFFT_INVALID_PLAN: return "CUFFT_INVALID_PLAN";
        case CUFFT_ALLOC_FAILED: return "CUFFT_ALLOC_FAILED";
        case CUFFT_INVALID_VALUE: return "CUFFT_INVALID_VALUE";
        case CUFFT_INTERNAL_ERROR: return "CUFFT_INTERNAL_ERROR";
        case CUFFT_EXEC_FAILED: return "CUFFT_EXEC_FAILED";
        case CUFFT_SETUP_FAILED: return "CUFFT_SETUP_FAILED";
        case CUFFT_INVALID_SIZE: return "CUFFT_INVALID_SIZE";
        default: return "CUFFT error";
    }
}

#define SK_CUFFT(ctx, call)                                                                             \
    do {                                                                                                \
        cufftResult r__ = (call);                                                                       \
        if (r__ != CUFFT_SUCCESS)                                                                       \
            return sk_fail((ctx), SKAGRID_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cufft_str(r__)); \
    } while (0)

static int get_fft_plan(skagrid_ctx *ctx, i64 n, cufftHandle *out) {
    auto it = ctx->fft_plans.find(n);
    if (it != ctx->fft_plans.end()) { *out = it->second; return SKAGRID_OK; }
    if (n <= 0 || n > (1 << 17)) return sk_fail(ctx, SKAGRID_EINVAL, "fft: size %lld out of range", n);
    cufftHandle h;
    SK_CUFFT(ctx, cufftCreate(&h));
    size_t ws = 0;
    cufftResult r = cufftSetAutoAllocation(h, 0);
    if (r == CUFFT_SUCCESS) r = cufftMakePlan2d(h, (int)n, (int)n, CUFFT_Z2Z, &ws);
    if (r != CUFFT_SUCCESS) { cufftDestroy(h); return sk_fail(ctx, SKAGRID_ECUDA, "cufftMakePlan2d(%lld): %s", n, cufft_str(r)); }
    DevBuf wb;
    if (ws > 0) {
        if (cudaMalloc(&wb.p, ws) != cudaSuccess) {
            cudaGetLastError();
            cufftDestroy(h);
            return sk_fail(ctx, SKAGRID_ENOMEM, "fft: work area of %zu bytes for n=%lld", ws, n);
        }
        wb.bytes = ws;
        r = cufftSetWorkArea(h, wb.p);
        if (r != CUFFT_SUCCESS) { cudaFree(wb.p); cufftDestroy(h); return sk_fail(ctx, SKAGRID_ECUDA, "cufftSetWorkArea: %s", cufft_str(r)); }
    }
    ctx->fft_plans[n] = h;
    ctx->fft_work[n] = wb;
    *out = h;
    return SKAGRID_OK;
}

// n x n complex-to-real plan (hermitian half-spectrum [n][n/2+1] -> real [n][n]); key 3 << 40 | n in the plan cache
static int get_fft_plan_z2d(skagrid_ctx *ctx, i64 n, cufftHandle *out) {
    const i64 key = ((i64)3 << 40) | n;
    auto it = ctx->fft_plans.find(key);
    if (it != ctx->fft_plans.end()) { *out = it->second; return SKAGRID_OK; }
    if (n <= 0 || n > (1 << 17)) return sk_fail(ctx, SKAGRID_EINVAL, "fft: size %lld out of range", n);
    cufftHandle h;
    SK_CUFFT(ctx, cufftCreate(&h));
    size_t ws = 0;
    cufftResult r = cufftSetAutoAllocation(h, 0);
    if (r == CUFFT_SUCCESS) r = cufftMakePlan2d(h, (int)n, (int)n, CUFFT_Z2D, &ws);
    if (r != CUFFT_SUCCESS) { cufftDestroy(h); return sk_fail(ctx, SKAGRID_ECUDA, "cufftMakePlan2d(%lld, Z2D): %s", n, cufft_str(r)); }
    DevBuf wb;
    if (ws > 0) {
        if (cudaMalloc(&wb.p, ws) != cudaSuccess) {
            cudaGetLastError();
            cufftDestroy(h);
            return sk_fail(ctx, SKAGRID_ENOMEM, "fft: work area of %zu bytes for n=%lld (Z2D)", ws, n);
        }
        wb.bytes = ws;
        r = cufftSetWorkArea(h, wb.p);
        if (r != CUFFT_SUCCESS) { cudaFree(wb.p); cufftDestroy(h); return sk_fail(ctx, SKAGRID_ECUDA, "cufftSetWorkArea: %s", cufft_str(r)); }
    }
    ctx->fft_plans[key] = h;
    ctx->fft_work[key] = wb;
    *out = h;
    return SKAGRID_OK;
}

// ---------------------------------------------------------------------------------------- hermitian
// even n: out[y,x] = g[y,x] + (x==0||y==0 ? 0 : conj g[n-y,n-x]);  odd n: g + conj(g[n-1-y,n-1-x]).
// One thread per unordered pair {cell, mirror}; safe in place (out == g).  modulate != 0 additionally
// multiplies by (-1)^(x+y) (first half of the centred transform for even n).
__global__ void __launch_bounds__(256) hermitian_kernel(i64 n, const double2 *g, double2 *out, int modulate)  /* g == out allowed: no __restrict__ */ {
    const i64 total = n * n;
    const int even = (n % 2 == 0);
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += stride) {
        const i64 y = c / n, x = c - y * n;
        const double sgn = (modulate && ((x + y) & 1)) ? -1.0 : 1.0;
        if (even && (x == 0 || y == 0)) {
            const double2 a = g[c];
            out[c] = make_double2(sgn * a.x, sgn * a.y);
            continue;
        }
        const i64 my = even ? n - y : n - 1 - y, mx = even ? n - x : n - 1 - x;
        const i64 m = my * n + mx;
        if (m < c) continue;  // the pair is handled by the thread of the smaller index
        const double2 a = g[c], b = g[m];
        const double2 ra = make_double2(a.x + b.x, a.y - b.y);
        const double2 rb = make_double2(b.x + a.x, b.y - a.y);
        // x+y and mx+my have the same parity for even n (their sum is 2n), so one sign serves both
        out[c] = make_double2(sgn * ra.x, sgn * ra.y);
        if (m != c) out[m] = make_double2(sgn * rb.x, sgn * rb.y);
    }
}

static int nblocks(skagrid_ctx *ctx, i64 total) {
    i64 b = (total + 255) / 256;
    const i64 cap = (i64)ctx->sm_count * 16;
    return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

int sk_hermitian_dev(skagrid_ctx *ctx, i64 n, const double *g, double *out, cudaStream_t st) {
    if (n <= 0) return sk_fail(ctx, SKAGRID_EINVAL, "make_grid_hermitian: n = %lld", n);
    hermitian_kernel<<<nblocks(ctx, n * n), 256, 0, st>>>(n, (const double2 *)g, (double2 *)out, 0);
    SK_LAUNCH_CHECK(ctx);
    return SKAGRID_OK;
}

// ---------------------------------------------------------------------------------------- centred FFT
__global__ void __launch_bounds__(256) modulate_kernel(i64 n, const double2 *in, double2 *out, double scale)  /* in == out allowed */ {
    const i64 total = n * n;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += stride) {
        const i64 y = c / n, x = c - y * n;
        const double s = ((x + y) & 1) ? -scale : scale;
        const double2 a = in[c];
        out[c] = make_double2(s * a.x, s * a.y);
    }
}

// out[y,x] = scale * in[(y+sh) mod n, (x+sh) mod n]
__global__ void __launch_bounds__(256) rotate_kernel(i64 n, i64 sh, const double2 *__restrict__ in, double2 *__restrict__ out, double scale) {
    const i64 total = n * n;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += stride) {
        const i64 y = c / n, x = c - y * n;
        const double2 a = in[((y + sh) % n) * n + (x + sh) % n];
        out[c] = make_double2(scale * a.x, scale * a.y);
    }
}

// Centred transform: out = shift2D(fft2D(ishift2D(in))), inverse is 1/n^2-normalised.  in may equal out.
int sk_fft2c_dev(skagrid_ctx *ctx, i64 n, const double *in, double *out, int inverse, cudaStream_t st) {
    cufftHandle plan;
    SK_TRY(get_fft_plan(ctx, n, &plan));
    SK_CUFFT(ctx, cufftSetStream(plan, st));
    const double scale = inverse ? 1.0 / ((double)n * (double)n) : 1.0;
    const int dir = inverse ? CUFFT_INVERSE : CUFFT_FORWARD;
    const int nb = nblocks(ctx, n * n);
    if (n % 2 == 0) {
        modulate_kernel<<<nb, 256, 0, st>>>(n, (const double2 *)in, (double2 *)out, 1.0);
        SK_LAUNCH_CHECK(ctx);
        SK_CUFFT(ctx, cufftExecZ2Z(plan, (cufftDoubleComplex *)out, (cufftDoubleComplex *)out, dir));
        ctx->launches++;
        modulate_kernel<<<nb, 256, 0, st>>>(n, (const double2 *)out, (double2 *)out, scale);
        SK_LAUNCH_CHECK(ctx);
    } else {
        void *tmp;
        SK_TRY(sk_scratch(ctx, "fft_tmp", (size_t)(n * n) * sizeof(double2), &tmp));
        // ishift2D: out[j] = in[(j + n/2) mod n]; shift2D: out[j] = in[(j + ceil(n/2)) mod n]
        rotate_kernel<<<nb, 256, 0, st>>>(n, n / 2, (const double2 *)in, (double2 *)tmp, 1.0);
        SK_LAUNCH_CHECK(ctx);
        SK_CUFFT(ctx, cufftExecZ2Z(plan, (cufftDoubleComplex *)tmp, (cufftDoubleComplex *)tmp, dir));
        ctx->launches++;
        rotate_kernel<<<nb, 256, 0, st>>>(n, (n + 1) / 2, (const double2 *)tmp, (double2 *)out, scale);
        SK_LAUNCH_CHECK(ctx);
    }
    return SKAGRID_OK;
}

// ---------------------------------------------------------------------------------------- real + max
__device__ __forceinline__ void atomic_max_double(double *addr, double v) {
    // monotone map double -> signed 64-bit: non-negative doubles compare as their bits, negative ones reversed
    long long *a = reinterpret_cast<long long *>(addr);
    long long old = *a, assumed;
    do {
        assumed = old;
        if (__longlong_as_double(assumed) >= v) break;
        old = atomicCAS(reinterpret_cast<unsigned long long *>(a), (unsigned long long)assumed, (unsigned long long)__double_as_longlong(v));
    } while (old != assumed);
}

__global__ void __launch_bounds__(256) real_max_kernel(i64 n, const double2 *__restrict__ g, double *__restrict__ image, double scale,
                                                       int modulate, double *__restrict__ max_out) {
    __shared__ double wmax[8];
    const i64 total = n * n;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    double m = -INFINITY;
    for (i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += stride) {
        const i64 y = c / n, x = c - y * n;
        const double s = (modulate && ((x + y) & 1)) ? -scale : scale;
        const double r = s * g[c].x;
        if (image) image[c] = r;
        m = fmax(m, r);
    }
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) m = fmax(m, wmax[i]);
        if (max_out) atomic_max_double(max_out, m);
    }
}

__global__ void set_double_kernel(double *p, double v) { *p = v; }

// Even n, real output only: real(ifft(A)) = ifft(Herm(A)) with Herm(A)[k] = (A[k] + conj(A[-k])) / 2, and an exactly hermitian
// spectrum needs only its half [n][n/2+1] and a complex-to-REAL transform -- half the passes of the Z2Z route and no N^2 complex
// intermediate.  With A = M (.) make_grid_hermitian(g), M = (-1)^(x+y) (src/Gridding.hs:585-605, :828-829):
//   x != 0 and y != 0:  make_grid_hermitian already pairs the cells: Herm(A) = M (g[y,x] + conj g[n-y,n-x])
//   row 0 / column 0:    left unsymmetrised by the reference (its ifft has an imaginary part there, which `map real` drops):
//                        Herm(A) = M (g[y,x] + conj g[(n-y)%n,(n-x)%n]) / 2
// half[y][x] for x <= n/2.  The grid itself is NOT modified.
__global__ void __launch_bounds__(256) herm_half_kernel(i64 n, const double2 *__restrict__ g, double2 *__restrict__ half) {
    const i64 hw = n / 2 + 1;
    const i64 total = n * hw;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += stride) {
        const i64 y = c / hw, x = c - y * hw;
        const i64 ym = y == 0 ? 0 : n - y, xm = x == 0 ? 0 : n - x;
        const double2 a = g[y * n + x], b = g[ym * n + xm];
        double f = (x == 0 || y == 0) ? 0.5 : 1.0;
        if ((x + y) & 1) f = -f;
        half[c] = make_double2(f * (a.x + b.x), f * (a.y - b.y));
    }
}

// image[y,x] = (-1)^(x+y) / n^2 * re[y,x] (in place when image == re); block maximum into max_out
__global__ void __launch_bounds__(256) real_finish_kernel(i64 n, const double *re, double *image, double scale, double *__restrict__ max_out) {
    __shared__ double wmax[8];
    const i64 total = n * n;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    double m = -INFINITY;
    for (i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += stride) {
        const i64 y = c / n, x = c - y * n;
        const double r = (((x + y) & 1) ? -scale : scale) * re[c];
        if (image) image[c] = r;
        m = fmax(m, r);
    }
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) m = fmax(m, wmax[i]);
        if (max_out) atomic_max_double(max_out, m);
    }
}

static int grid_to_image_c2r(skagrid_ctx *ctx, i64 n, const double *grid, double *image, double *max_out, cudaStream_t st) {
    cufftHandle plan;
    SK_TRY(get_fft_plan_z2d(ctx, n, &plan));
    SK_CUFFT(ctx, cufftSetStream(plan, st));
    void *half, *re = image;
    SK_TRY(sk_scratch(ctx, "g2i_half", (size_t)(n * (n / 2 + 1)) * sizeof(double2), &half));
    if (!re) SK_TRY(sk_scratch(ctx, "g2i_real", (size_t)(n * n) * sizeof(double), &re));
    herm_half_kernel<<<nblocks(ctx, n * (n / 2 + 1)), 256, 0, st>>>(n, (const double2 *)grid, (double2 *)half);
    SK_LAUNCH_CHECK(ctx);
    SK_CUFFT(ctx, cufftExecZ2D(plan, (cufftDoubleComplex *)half, (cufftDoubleReal *)re));
    ctx->launches++;
    real_finish_kernel<<<nblocks(ctx, n * n), 256, 0, st>>>(n, (const double *)re, image, 1.0 / ((double)n * (double)n), max_out);
    SK_LAUNCH_CHECK(ctx);
    return SKAGRID_OK;
}

// hermitian -> centred inverse FFT; image = real part; max_out = maximum pixel.  Even n: complex-to-real route above, `grid`
// is left as it is.  Odd n (and SKAGRID_G2I_Z2Z=1): in place on `grid`, which then holds the complex image plane.
int sk_grid_to_image_dev(skagrid_ctx *ctx, i64 n, double *grid, double *image, double *max_out, cudaStream_t st) {
    if (n <= 0) return sk_fail(ctx, SKAGRID_EINVAL, "grid_to_image: n = %lld", n);
    static const int force_z2z = getenv("SKAGRID_G2I_Z2Z") ? atoi(getenv("SKAGRID_G2I_Z2Z")) : 0;  // A/B measurements
    if (n % 2 == 0 && !force_z2z) {
        if (max_out) { set_double_kernel<<<1, 1, 0, st>>>(max_out, -INFINITY); SK_LAUNCH_CHECK(ctx); }
        return grid_to_image_c2r(ctx, n, grid, image, max_out, st);
    }
    cufftHandle plan;
    SK_TRY(get_fft_plan(ctx, n, &plan));
    SK_CUFFT(ctx, cufftSetStream(plan, st));
    const int nb = nblocks(ctx, n * n);
    const double scale = 1.0 / ((double)n * (double)n);
    if (max_out) { set_double_kernel<<<1, 1, 0, st>>>(max_out, -INFINITY); SK_LAUNCH_CHECK(ctx); }
    if (n % 2 == 0) {
        hermitian_kernel<<<nb, 256, 0, st>>>(n, (const double2 *)grid, (double2 *)grid, 1);
        SK_LAUNCH_CHECK(ctx);
        SK_CUFFT(ctx, cufftExecZ2Z(plan, (cufftDoubleComplex *)grid, (cufftDoubleComplex *)grid, CUFFT_INVERSE));
        ctx->launches++;
        real_max_kernel<<<nb, 256, 0, st>>>(n, (const double2 *)grid, image, scale, 1, max_out);
        SK_LAUNCH_CHECK(ctx);
    } else {
        hermitian_kernel<<<nb, 256, 0, st>>>(n, (const double2 *)grid, (double2 *)grid, 0);
        SK_LAUNCH_CHECK(ctx);
        SK_TRY(sk_fft2c_dev(ctx, n, grid, grid, 1, st));
        real_max_kernel<<<nb, 256, 0, st>>>(n, (const double2 *)grid, image, 1.0, 0, max_out);
        SK_LAUNCH_CHECK(ctx);
    }
    return SKAGRID_OK;
}

// ---------------------------------------------------------------------------------------- slab-distributed grid -> image
// When the grid is spread over devices as row slabs (uv-tile-sharded gridding, or a reduce-scatter of visibility-sharded
// grids), the image is formed without gathering it: per device 1-D inverse FFTs along x of its rows, an all-to-all
// transpose (done by the caller: torch.distributed / peer copies), 1-D inverse FFTs along y of its column slab.
//
// make_grid_hermitian needs no exchange of mirrored rows when only real(ifft) is wanted (src/ImageDataset.hs:74-76):
// for even n, H = g + conj(R g') with g' = g with row 0 and column 0 zeroed and (R a)[y,x] = a[(n-y) mod n, (n-x) mod n];
// ifft(conj(R a)) = conj(ifft(a)), so real(ifft(H)) = real(ifft(g + g')) -- the grid scaled by 2 except on row 0 and
// column 0.  The centring factor M = (-1)^(x+y) commutes with R, so the same holds for the centred transform.
static int get_fft1d_plan(skagrid_ctx *ctx, i64 n, i64 batch, int strided, cufftHandle *out) {
    if (n <= 0 || n > (1 << 17) || batch <= 0 || batch > (1 << 17)) return sk_fail(ctx, SKAGRID_EINVAL, "fft1d: size %lld x %lld out of range", n, batch);
    const i64 key = ((i64)(strided ? 2 : 1) << 40) | (n << 20) | batch;  // 2-D plans use the bare n as key
    auto it = ctx->fft_plans.find(key);
    if (it != ctx->fft_plans.end()) { *out = it->second; return SKAGRID_OK; }
    cufftHandle h;
    SK_CUFFT(ctx, cufftCreate(&h));
    size_t ws = 0;
    int dims[1] = {(int)n};
    // contiguous: `batch` rows of n; strided: `batch` columns of an [n, batch] row-major array
    int inembed[1] = {(int)n};
    cufftResult r = cufftSetAutoAllocation(h, 0);
    if (r == CUFFT_SUCCESS)
        r = strided ? cufftMakePlanMany(h, 1, dims, inembed, (int)batch, 1, inembed, (int)batch, 1, CUFFT_Z2Z, (int)batch, &ws)
                    : cufftMakePlanMany(h, 1, dims, inembed, 1, (int)n, inembed, 1, (int)n, CUFFT_Z2Z, (int)batch, &ws);
    if (r != CUFFT_SUCCESS) { cufftDestroy(h); return sk_fail(ctx, SKAGRID_ECUDA, "cufftMakePlanMany(%lld x %lld): %s", n, batch, cufft_str(r)); }
    DevBuf wb;
    if (ws > 0) {
        if (cudaMalloc(&wb.p, ws) != cudaSuccess) {
            cudaGetLastError();
            cufftDestroy(h);
            return sk_fail(ctx, SKAGRID_ENOMEM, "fft1d: work area of %zu bytes", ws);
        }
        wb.bytes = ws;
        r = cufftSetWorkArea(h, wb.p);
        if (r != CUFFT_SUCCESS) { cudaFree(wb.p); cufftDestroy(h); return sk_fail(ctx, SKAGRID_ECUDA, "cufftSetWorkArea: %s", cufft_str(r)); }
    }
    ctx->fft_plans[key] = h;
    ctx->fft_work[key] = wb;
    *out = h;
    return SKAGRID_OK;
}

// slab[y - row0, x] *= (x == 0 || y == 0 ? 1 : 2) * (-1)^(x+y)
__global__ void __launch_bounds__(256) slab_prescale_kernel(i64 n, i64 row0, i64 nrows, double2 *__restrict__ slab) {
    const i64 total = nrows * n;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += stride) {
        const i64 y = row0 + c / n, x = c % n;
        double f = (x == 0 || y == 0) ? 1.0 : 2.0;
        if ((x + y) & 1) f = -f;
        double2 v = slab[c];
        v.x *= f; v.y *= f;
        slab[c] = v;
    }
}

// image[y, x - col0] = real(cols[y, x - col0]) * (-1)^(x+y) / n^2; block maximum into max_out
__global__ void __launch_bounds__(256) slab_finish_kernel(i64 n, i64 col0, i64 ncols, const double2 *__restrict__ cols, double *__restrict__ image,
                                                          double scale, double *__restrict__ max_out) {
    __shared__ double wmax[8];
    const i64 total = n * ncols;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    double m = -INFINITY;
    for (i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += stride) {
        const i64 y = c / ncols, x = col0 + c % ncols;
        const double r = (((x + y) & 1) ? -scale : scale) * cols[c].x;
        if (image) image[c] = r;
        m = fmax(m, r);
    }
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) m = fmax(m, wmax[i]);
        if (max_out) atomic_max_double(max_out, m);
    }
}

// Stage 1, in place on rows [row0, row0 + nrows) of an n x n grid (n even): hermitian weighting + centring + inverse FFT along x.
int sk_slab_fft_rows_dev(skagrid_ctx *ctx, i64 n, i64 row0, i64 nrows, double *slab, cudaStream_t st) {
    if (n <= 0 || (n & 1)) return sk_fail(ctx, SKAGRID_EINVAL, "slab grid_to_image: the grid side must be even (n = %lld)", n);
    if (row0 < 0 || nrows < 0 || row0 + nrows > n) return sk_fail(ctx, SKAGRID_EINVAL, "slab grid_to_image: rows [%lld, %lld) outside the grid", row0, row0 + nrows);
    if (nrows == 0) return SKAGRID_OK;
    slab_prescale_kernel<<<nblocks(ctx, nrows * n), 256, 0, st>>>(n, row0, nrows, (double2 *)slab);
    SK_LAUNCH_CHECK(ctx);
    cufftHandle plan;
    SK_TRY(get_fft1d_plan(ctx, n, nrows, 0, &plan));
    SK_CUFFT(ctx, cufftSetStream(plan, st));
    SK_CUFFT(ctx, cufftExecZ2Z(plan, (cufftDoubleComplex *)slab, (cufftDoubleComplex *)slab, CUFFT_INVERSE));
    ctx->launches++;
    return SKAGRID_OK;
}

// Stage 2, on columns [col0, col0 + ncols) held as an [n, ncols] row-major array (transformed in place): inverse FFT along
// y, then image = real part, centred and normalised; max_out (may be NULL) receives the maximum pixel of this column slab.
int sk_slab_fft_cols_dev(skagrid_ctx *ctx, i64 n, i64 col0, i64 ncols, double *cols, double *image, double *max_out, cudaStream_t st) {
    if (n <= 0 || (n & 1)) return sk_fail(ctx, SKAGRID_EINVAL, "slab grid_to_image: the grid side must be even (n = %lld)", n);
    if (col0 < 0 || ncols < 0 || col0 + ncols > n) return sk_fail(ctx, SKAGRID_EINVAL, "slab grid_to_image: columns [%lld, %lld) outside the grid", col0, col0 + ncols);
    if (max_out) { set_double_kernel<<<1, 1, 0, st>>>(max_out, -INFINITY); SK_LAUNCH_CHECK(ctx); }
    if (ncols == 0) return SKAGRID_OK;
    cufftHandle plan;
    SK_TRY(get_fft1d_plan(ctx, n, ncols, 1, &plan));
    SK_CUFFT(ctx, cufftSetStream(plan, st));
    SK_CUFFT(ctx, cufftExecZ2Z(plan, (cufftDoubleComplex *)cols, (cufftDoubleComplex *)cols, CUFFT_INVERSE));
    ctx->launches++;
    slab_finish_kernel<<<nblocks(ctx, n * ncols), 256, 0, st>>>(n, col0, ncols, (const double2 *)cols, image, 1.0 / ((double)n * (double)n), max_out);
    SK_LAUNCH_CHECK(ctx);
    return SKAGRID_OK;
}

// ---------------------------------------------------------------------------------------- pad / crop
// Centre-pad (n_out > n_in; pad_mid src/Gridding.hs:682-691 incl. the padder transpose when transpose != 0)
// or centre-crop (n_out < n_in; extract_mid :694-707).  n_out == n_in copies.
__global__ void __launch_bounds__(256) pad_crop_kernel(i64 n_in, const double2 *__restrict__ in, i64 n_out, double2 *__restrict__ out, int transpose) {
    const i64 total = n_out * n_out;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    // pad: before = n_out/2 - n_in/2 ; crop: start = n_in/2 - n_out/2
    const i64 off = n_out >= n_in ? -(n_out / 2 - n_in / 2) : (n_in / 2 - n_out / 2);
    for (i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += stride) {
        const i64 y = c / n_out, x = c - y * n_out;
        const i64 sy = y + off, sx = x + off;
        double2 v = make_double2(0.0, 0.0);
        if (sy >= 0 && sy < n_in && sx >= 0 && sx < n_in) v = transpose ? in[sx * n_in + sy] : in[sy * n_in + sx];
        out[c] = v;
    }
}

int sk_pad_crop_dev(skagrid_ctx *ctx, i64 n_in, const double *in, i64 n_out, double *out, cudaStream_t st) {
    pad_crop_kernel<<<nblocks(ctx, n_out * n_out), 256, 0, st>>>(n_in, (const double2 *)in, n_out, (double2 *)out, n_out > n_in);
    SK_LAUNCH_CHECK(ctx);
    return SKAGRID_OK;
}

// ---------------------------------------------------------------------------------------- w-kernels
// w_kernel (src/Gridding.hs:610-619): far field exp(2 pi i w (1 - sqrt(1 - l^2 - m^2))) on the npixff grid
// of coordinates2 (:637-648) scaled by theta, centre-padded to npixff*qpx (the transposing padder is
// harmless: the far field is symmetric in l,m -- we still honour it), centred inverse FFT,
// extract_oversampled (:709-728): out[yf,xf,y,x] = qpx^2 * af[c0 - yf + qpx*y, c0 - xf + qpx*x].
// kernel_coordinates (:620-635): (l, m) = theta * coordinates2, optionally through the 2 x 2 matrix t
// (l, m) -> (t00 l + t10 m, t01 l + t11 m), then shifted by (dl, dm) = (patHorShift, patVerShift).
struct WCoord {
    double t00, t10, t01, t11, dl, dm;
};
__global__ void __launch_bounds__(256) wfarfield_kernel(i64 npixff, i64 big, double theta, double w, WCoord K, double2 *__restrict__ out) {
    const i64 total = big * big;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    const i64 before = big / 2 - npixff / 2;
    const double step = 1.0 / (double)npixff;
    const i64 n2 = npixff / 2;
    for (i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += stride) {
        const i64 y = c / big, x = c - y * big;
        const i64 oy = y - before, ox = x - before;
        double2 v = make_double2(0.0, 0.0);
        if (oy >= 0 && oy < npixff && ox >= 0 && ox < npixff) {
            // padder reads ff[ox, oy] (transpose); ff[r, c] uses l = base[c], m = base[r]
            // pad_mid returns the far field untouched -- NOT transposed -- when no padding is needed (qpx = 1, :688)
            const i64 ci = big == npixff ? ox : oy, ri = big == npixff ? oy : ox;
            const double l1 = ((double)(-n2) * step + (double)ci * step) * theta;
            const double m1 = ((double)(-n2) * step + (double)ri * step) * theta;
            const double l = K.t00 * l1 + K.t10 * m1 + K.dl;
            const double m = K.t01 * l1 + K.t11 * m1 + K.dm;
            const double r2 = l * l + m * m;
            const double ph = 1.0 - sqrt(1.0 - r2);
            double sn, cs;
            sincos(2.0 * 3.14159265358979323846 * w * ph, &sn, &cs);
            v = make_double2(cs, sn);
        }
        out[c] = v;
    }
}

__global__ void __launch_bounds__(256) extract_oversampled_kernel(i64 big, const double2 *__restrict__ af, i64 qpx, i64 n, int conjugate,
                                                                  double2 *__restrict__ out) {
    const i64 total = qpx * qpx * n * n;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    const i64 cons = big / 2 - qpx * (n / 2);
    const double sc = (double)(qpx * qpx);
    for (i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += stride) {
        i64 t = c;
        const i64 x = t % n; t /= n;
        const i64 y = t % n; t /= n;
        const i64 xf = t % qpx; const i64 yf = t / qpx;
        const i64 sy = cons - yf + qpx * y, sx = cons - xf + qpx * x;
        double2 v = make_double2(0.0, 0.0);
        if (sy >= 0 && sy < big && sx >= 0 && sx < big) v = af[sy * big + sx];
        out[c] = make_double2(sc * v.x, conjugate ? -sc * v.y : sc * v.y);
    }
}

int sk_w_kernels_dev(skagrid_ctx *ctx, double theta, i64 nw, const double *w_host, i64 npixff, i64 npixkern, i64 qpx, int conjugate,
                     double *out, cudaStream_t st) {
    return sk_w_kernels_ex_dev(ctx, theta, nw, w_host, npixff, npixkern, qpx, conjugate, nullptr, 0.0, 0.0, out, st);
}

int sk_w_kernels_ex_dev(skagrid_ctx *ctx, double theta, i64 nw, const double *w_host, i64 npixff, i64 npixkern, i64 qpx, int conjugate,
                        const double *transmat, double dl, double dm, double *out, cudaStream_t st) {
    WCoord K = {1.0, 0.0, 0.0, 1.0, dl, dm};
    if (transmat) { K.t00 = transmat[0]; K.t01 = transmat[1]; K.t10 = transmat[2]; K.t11 = transmat[3]; }  // row-major t[r][c]
    if (nw <= 0 || npixff <= 0 || npixkern <= 0 || qpx <= 0) return sk_fail(ctx, SKAGRID_EINVAL, "w_kernels: non-positive size");
    const i64 big = npixff * qpx;
    if (qpx * (npixkern / 2) > big / 2 || big > (1 << 15)) return sk_fail(ctx, SKAGRID_EINVAL, "w_kernels: kernel %lld x oversampling %lld does not fit the %lld far field", npixkern, qpx, npixff);
    void *buf;
    SK_TRY(sk_scratch(ctx, "wkern_ff", (size_t)(big * big) * sizeof(double2), &buf));
    const i64 per = qpx * qpx * npixkern * npixkern;
    for (i64 i = 0; i < nw; ++i) {
        wfarfield_kernel<<<nblocks(ctx, big * big), 256, 0, st>>>(npixff, big, theta, w_host[i], K, (double2 *)buf);
        SK_LAUNCH_CHECK(ctx);
        SK_TRY(sk_fft2c_dev(ctx, big, (double *)buf, (double *)buf, 1, st));
        extract_oversampled_kernel<<<nblocks(ctx, per), 256, 0, st>>>(big, (const double2 *)buf, qpx, npixkern, conjugate, (double2 *)out + i * per);
        SK_LAUNCH_CHECK(ctx);
    }
    return SKAGRID_OK;
}
