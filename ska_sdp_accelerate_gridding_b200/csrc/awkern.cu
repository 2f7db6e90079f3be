// awkern.cu -- AW-kernel formation.
//
//   convolve2d     src/Gridding.hs:795-811  (pad_mid :682-691, padder :863-877, extract_mid :694-707)
//   aw_kernel_fn2  src/Gridding.hs:761-775  convolve2d (convolve2d a1 a2) (w[yf,xf])
//
// The reference evaluates convolve2d with three zero-padded m x m FFTs (m = 2^ceil(log2(2n-1))).  Its
// padder reads array ! (x, y), i.e. it transposes while padding, so the net effect is
//     out[ty,tx] = sum_{ky,kx} a1[kx,ky] * a2[tx + c - kx, ty + c - ky],   c = n div 2
// (the centre-"same" convolution of the transposed inputs).  Here it is evaluated directly in fp64: one
// thread block per output kernel, operands staged in shared memory, one output tap per thread.
// For n = 15 the direct form is n^4 = 50 625 complex MACs per convolution, all on-chip.
#include "common.cuh"

__device__ __forceinline__ void conv_same_t(int n, const double2 *__restrict__ a, const double2 *__restrict__ b, int ty, int tx, double2 &out) {
    const int c = n / 2;
    double sr = 0.0, si = 0.0;
    for (int ky = 0; ky < n; ++ky) {
        const int qy = ty + c - ky;
        if (qy < 0 || qy >= n) continue;
        for (int kx = 0; kx < n; ++kx) {
            const int qx = tx + c - kx;
            if (qx < 0 || qx >= n) continue;
            const double2 p = a[kx * n + ky];
            const double2 q = b[qx * n + qy];
            sr = fma(p.x, q.x, sr); sr = fma(-p.y, q.y, sr);
            si = fma(p.x, q.y, si); si = fma(p.y, q.x, si);
        }
    }
    out = make_double2(sr, si);
}

// out[k] = convolve2d(a[ai[k]], b[bi[k]]) (ai/bi NULL: k itself)
__global__ void __launch_bounds__(256) convolve2d_kernel(int n, const double2 *__restrict__ a, const i64 *__restrict__ ai,
                                                         const double2 *__restrict__ b, const i64 *__restrict__ bi,
                                                         double2 *__restrict__ out, int conj_out) {
    extern __shared__ double2 sm[];
    const int n2 = n * n;
    double2 *sa = sm, *sb = sm + n2;
    const i64 k = blockIdx.x;
    const double2 *pa = a + (ai ? ai[k] : k) * n2;
    const double2 *pb = b + (bi ? bi[k] : k) * n2;
    for (int t = threadIdx.x; t < n2; t += blockDim.x) { sa[t] = pa[t]; sb[t] = pb[t]; }
    __syncthreads();
    for (int t = threadIdx.x; t < n2; t += blockDim.x) {
        double2 r;
        conv_same_t(n, sa, sb, t / n, t % n, r);
        if (conj_out) r.y = -r.y;
        out[k * n2 + t] = r;
    }
}

int sk_convolve2d_dev(skagrid_ctx *ctx, i64 n, i64 count, const double *a, const i64 *ai, const double *b, const i64 *bi, double *out,
                      int conj_out, cudaStream_t st) {
    if (count <= 0) return SKAGRID_OK;
    if (n <= 0 || n > 64) return sk_fail(ctx, SKAGRID_EINVAL, "convolve2d: size %lld outside [1,64]", n);
    const size_t smem = (size_t)(2 * n * n) * sizeof(double2);
    if (smem > 48 * 1024) SK_CUDA(ctx, cudaFuncSetAttribute(convolve2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  // per call: cheap, and correct on any device
    convolve2d_kernel<<<(unsigned)count, 256, smem, st>>>((int)n, (const double2 *)a, ai, (const double2 *)b, bi, (double2 *)out, conj_out);
    SK_LAUNCH_CHECK(ctx);
    return SKAGRID_OK;
}

// out[k] = aw_kernel_fn2 yf[k] xf[k] wkerns[wbin[k]] akerns[a1[k]] akerns[a2[k]]   (optionally conjugated:
// processOne2, src/Gridding.hs:391, multiplies the visibility with conj(AW))
__global__ void __launch_bounds__(256) aw_kernel_kernel(int s, i64 nw, i64 qpx, const double2 *__restrict__ wkerns, i64 nant,
                                                        const double2 *__restrict__ akerns, const i64 *__restrict__ wbin,
                                                        const i64 *__restrict__ yf, const i64 *__restrict__ xf,
                                                        const i64 *__restrict__ a1, const i64 *__restrict__ a2,
                                                        double2 *__restrict__ out, int conj_out, uint32_t *__restrict__ err_flag) {
    extern __shared__ double2 sm[];
    const int s2 = s * s;
    double2 *sa = sm, *sb = sm + s2, *sc = sm + 2 * s2;
    const i64 k = blockIdx.x;
    const i64 wb = wbin[k], iy = yf[k], ix = xf[k], i1 = a1[k], i2 = a2[k];
    if (wb < 0 || wb >= nw || iy < 0 || iy >= qpx || ix < 0 || ix >= qpx || i1 < 0 || i1 >= nant || i2 < 0 || i2 >= nant) {
        if (threadIdx.x == 0) atomicOr(err_flag, 1u);
        for (int t = threadIdx.x; t < s2; t += blockDim.x) out[k * s2 + t] = make_double2(0.0, 0.0);
        return;
    }
    const double2 *p1 = akerns + i1 * s2, *p2 = akerns + i2 * s2;
    const double2 *pw = wkerns + ((wb * qpx + iy) * qpx + ix) * s2;
    for (int t = threadIdx.x; t < s2; t += blockDim.x) { sa[t] = p1[t]; sb[t] = p2[t]; }
    __syncthreads();
    for (int t = threadIdx.x; t < s2; t += blockDim.x) {
        double2 r;
        conv_same_t(s, sa, sb, t / s, t % s, r);
        sc[t] = r;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < s2; t += blockDim.x) sa[t] = pw[t];
    __syncthreads();
    for (int t = threadIdx.x; t < s2; t += blockDim.x) {
        double2 r;
        conv_same_t(s, sc, sa, t / s, t % s, r);
        if (conj_out) r.y = -r.y;
        out[k * s2 + t] = r;
    }
}

int sk_aw_kernels_dev(skagrid_ctx *ctx, i64 nw, i64 qpx, i64 s, const double *wkerns, i64 nant, const double *akerns, i64 count,
                      const i64 *wbin, const i64 *yf, const i64 *xf, const i64 *a1, const i64 *a2, double *out, int conj_out,
                      uint32_t *err_flag, cudaStream_t st) {
    if (count <= 0) return SKAGRID_OK;
    if (s <= 0 || s > 64) return sk_fail(ctx, SKAGRID_EINVAL, "aw_kernel: support %lld outside [1,64]", s);
    const size_t smem = (size_t)(3 * s * s) * sizeof(double2);
    if (smem > 48 * 1024) SK_CUDA(ctx, cudaFuncSetAttribute(aw_kernel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    aw_kernel_kernel<<<(unsigned)count, 256, smem, st>>>((int)s, nw, qpx, (const double2 *)wkerns, nant, (const double2 *)akerns, wbin, yf, xf,
                                                         a1, a2, (double2 *)out, conj_out, err_flag);
    SK_LAUNCH_CHECK(ctx);
    return SKAGRID_OK;
}
