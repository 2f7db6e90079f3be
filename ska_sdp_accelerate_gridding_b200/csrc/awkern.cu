// awkern.cu -- AW-kernel formation.
//
//   convolve2d     src/Gridding.hs:795-811  (pad_mid :682-691, padder :863-877, extract_mid :694-707)
//   aw_kernel_fn2  src/Gridding.hs:761-775  convolve2d (convolve2d a1 a2) (w[yf,xf])
//
// The reference evaluates convolve2d with three zero-padded m x m FFTs (m = 2^ceil(log2(2n-1))).  Its
// padder reads array ! (x, y), i.e. it transposes while padding, so the net effect is
//     out[ty,tx] = sum_{ky,kx} a1[kx,ky] * a2[tx + c - kx, ty + c - ky],   c = n div 2
// (the centre-"same" convolution of the transposed inputs).  Here it is evaluated directly in fp64, on-chip:
// n^4 = 50 625 complex MACs per convolution for n = 15 (28 561 of them on non-zero operands).
//
// aw_kernel_fn2 for a batch of visibilities runs in two stages, because the inner convolution depends only on the
// antenna pair: (1) the distinct (a1, a2) pairs of the batch are found (one compare-and-swap per visibility into an
// nant x nant slot table) and convolved once each; (2) every visibility convolves its pair's result with its
// w-kernel slice.  That halves the work at worst and leaves stage 1 negligible for real arrays (nant^2 << batch).
#include <algorithm>
#include <cstdlib>

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

// Generic batch kernel (any n <= 64): out[k] = convolve2d(a[ai[k]], b[bi[k]]) (ai/bi NULL: k itself), one block per
// output kernel, one output tap per thread.  count_dev != NULL: only the first *count_dev outputs exist (a count known
// on the device only); ai[k] < 0: the output is zero (a visibility whose indices were out of range).
__global__ void __launch_bounds__(256) convolve2d_kernel(int n, const uint32_t *__restrict__ count_dev, const double2 *__restrict__ a,
                                                         const i64 *__restrict__ ai, const double2 *__restrict__ b, const i64 *__restrict__ bi,
                                                         double2 *__restrict__ out, int conj_out) {
    extern __shared__ double2 sm[];
    const int n2 = n * n;
    double2 *sa = sm, *sb = sm + n2;
    const i64 k = blockIdx.x;
    if (count_dev && k >= (i64)*count_dev) return;
    const i64 ia = ai ? ai[k] : k;
    if (ia < 0) {
        for (int t = threadIdx.x; t < n2; t += blockDim.x) out[k * n2 + t] = make_double2(0.0, 0.0);
        return;
    }
    const double2 *pa = a + ia * n2;
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

// Register-tiled batch kernel for small odd supports (N <= 17; the reference's kernels are 15 x 15).
//   * both operands are staged transposed, so the sum becomes a plain row-major 2-D convolution
//         out[ty,tx] = sum_{ky,kx} aT[ky][kx] * bT[ty+C-ky][tx+C-kx];
//     bT is zero-padded in x (and has one all-zero row for qy outside the kernel): no bounds tests in the loops;
//   * a thread owns TX = 4 consecutive outputs of one row.  Per ky it loads the TX+N-1 taps of the bT row it needs
//     once (a sliding window in registers) and then does N x TX complex FMAs against broadcast loads of the aT row:
//     0.55 shared-memory loads per complex MAC instead of 2, which moves the kernel from the shared-memory limit to
//     the FP64 pipe;
//   * lanes run down the rows (ty = j mod N) and the padded row pitch is 1 mod 8 taps, so the window loads of a
//     quarter-warp hit 8 distinct 16-byte bank groups; the aT loads are warp-uniform (one wavefront);
//   * ceil(N*ceil(N/4)/32) warps per output kernel, several output kernels per 256-thread block.
template <int N>
struct ConvTile {
    static constexpr int TX = 4;
    static constexpr int C = N / 2;
    static constexpr int GX = (N + TX - 1) / TX;                                // column groups per row
    static constexpr int PW = ((GX * TX + N - 1) + 7) / 8 * 8 + 1;             // padded row pitch of bT, 1 mod 8
    static constexpr int PADLO = N - 1 - C;                                     // bT column qx lives at qx + PADLO
    static constexpr int THREADS_PER = (N * GX + 31) / 32 * 32;                 // threads per output kernel
    static constexpr int PER_BLOCK = 256 / THREADS_PER > 0 ? 256 / THREADS_PER : 1;
    static constexpr int THREADS = PER_BLOCK * THREADS_PER;
    static constexpr int SM_PER = N * N + (N + 1) * PW;                         // double2 per output kernel: aT, bT rows + zero row
};

template <int N>
__global__ void __launch_bounds__(ConvTile<N>::THREADS) conv_tiled_kernel(i64 count, const uint32_t *__restrict__ count_dev,
                                                                          const double2 *__restrict__ a, const i64 *__restrict__ ai,
                                                                          const double2 *__restrict__ b, const i64 *__restrict__ bi,
                                                                          double2 *__restrict__ out, int conj_out) {
    using T = ConvTile<N>;
    constexpr int TX = T::TX, C = T::C, GX = T::GX, PW = T::PW, N2 = N * N;
    extern __shared__ double2 sm[];
    const int grp = threadIdx.x / T::THREADS_PER, j = threadIdx.x % T::THREADS_PER;
    double2 *saT = sm + grp * T::SM_PER, *sbT = saT + N2;
    const i64 k = (i64)blockIdx.x * T::PER_BLOCK + grp;
    const i64 limit = count_dev ? min(count, (i64)*count_dev) : count;
    const bool live = k < limit;
    const i64 ia = live ? (ai ? ai[k] : k) : -1;
    if (ia >= 0) {
        const double2 *pa = a + ia * N2;
        for (int t = j; t < (N + 1) * PW; t += T::THREADS_PER) sbT[t] = make_double2(0.0, 0.0);
        for (int t = j; t < N2; t += T::THREADS_PER) saT[(t % N) * N + t / N] = pa[t];  // aT[ky][kx] = a[kx,ky]
    }
    __syncthreads();
    if (ia >= 0) {
        const double2 *pb = b + (bi ? bi[k] : k) * N2;
        for (int t = j; t < N2; t += T::THREADS_PER) sbT[(t % N) * PW + t / N + T::PADLO] = pb[t];  // bT[qy][qx] = b[qx,qy]
    }
    __syncthreads();
    const int ty = j % N, tx0 = (j / N) * TX;
    if (!live || j >= N * GX) return;
    double2 *po = out + k * N2 + ty * N;
    if (ia < 0) {
#pragma unroll
        for (int t = 0; t < TX; ++t)
            if (tx0 + t < N) po[tx0 + t] = make_double2(0.0, 0.0);
        return;
    }
    double accr[TX], acci[TX];
#pragma unroll
    for (int t = 0; t < TX; ++t) accr[t] = acci[t] = 0.0;
#pragma unroll 1
    for (int ky = 0; ky < N; ++ky) {
        const int qy = ty + C - ky;
        const double2 *brow = sbT + ((qy >= 0 && qy < N) ? qy : N) * PW + tx0;  // row N is all zero
        const double2 *arow = saT + ky * N;
        double2 w[TX + N - 1];
#pragma unroll
        for (int i = 0; i < TX + N - 1; ++i) w[i] = brow[i];
#pragma unroll
        for (int kx = 0; kx < N; ++kx) {
            const double2 p = arow[kx];
#pragma unroll
            for (int t = 0; t < TX; ++t) {
                const double2 q = w[t + N - 1 - kx];
                accr[t] = fma(p.x, q.x, accr[t]); accr[t] = fma(-p.y, q.y, accr[t]);
                acci[t] = fma(p.x, q.y, acci[t]); acci[t] = fma(p.y, q.x, acci[t]);
            }
        }
    }
#pragma unroll
    for (int t = 0; t < TX; ++t)
        if (tx0 + t < N) po[tx0 + t] = make_double2(accr[t], conj_out ? -acci[t] : acci[t]);
}

template <int N>
static int launch_tiled(skagrid_ctx *ctx, i64 count, const uint32_t *count_dev, const double2 *a, const i64 *ai, const double2 *b, const i64 *bi,
                        double2 *out, int conj_out, cudaStream_t st) {
    using T = ConvTile<N>;
    const size_t smem = (size_t)T::PER_BLOCK * T::SM_PER * sizeof(double2);
    const void *fn = (const void *)conv_tiled_kernel<N>;
    if (smem > 48 * 1024 && !ctx->smem_configured.count(fn)) {
        SK_CUDA(ctx, cudaFuncSetAttribute(conv_tiled_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->smem_configured.insert(fn);
    }
    const i64 blocks = (count + T::PER_BLOCK - 1) / T::PER_BLOCK;
    conv_tiled_kernel<N><<<(unsigned)blocks, T::THREADS, smem, st>>>(count, count_dev, a, ai, b, bi, out, conj_out);
    SK_LAUNCH_CHECK(ctx);
    return SKAGRID_OK;
}

// out[k] = convolve2d(a[ai[k]], b[bi[k]]) for k < min(count, *count_dev)
static int conv_batch(skagrid_ctx *ctx, i64 n, i64 count, const uint32_t *count_dev, const double *a, const i64 *ai, const double *b, const i64 *bi,
                      double *out, int conj_out, cudaStream_t st) {
    if (count <= 0) return SKAGRID_OK;
    if (n <= 0 || n > 64) return sk_fail(ctx, SKAGRID_EINVAL, "convolve2d: size %lld outside [1,64]", n);
    const double2 *pa = (const double2 *)a, *pb = (const double2 *)b;
    double2 *po = (double2 *)out;
    static const bool generic_only = getenv("SKAGRID_CONV_GENERIC") && atoi(getenv("SKAGRID_CONV_GENERIC")) != 0;  // A/B measurements
    if (!generic_only) {
        switch (n) {
            case 5: return launch_tiled<5>(ctx, count, count_dev, pa, ai, pb, bi, po, conj_out, st);
            case 7: return launch_tiled<7>(ctx, count, count_dev, pa, ai, pb, bi, po, conj_out, st);
            case 9: return launch_tiled<9>(ctx, count, count_dev, pa, ai, pb, bi, po, conj_out, st);
            case 11: return launch_tiled<11>(ctx, count, count_dev, pa, ai, pb, bi, po, conj_out, st);
            case 13: return launch_tiled<13>(ctx, count, count_dev, pa, ai, pb, bi, po, conj_out, st);
            case 15: return launch_tiled<15>(ctx, count, count_dev, pa, ai, pb, bi, po, conj_out, st);
            case 17: return launch_tiled<17>(ctx, count, count_dev, pa, ai, pb, bi, po, conj_out, st);
            default: break;
        }
    }
    const size_t smem = (size_t)(2 * n * n) * sizeof(double2);
    if (smem > 48 * 1024) SK_CUDA(ctx, cudaFuncSetAttribute(convolve2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  // per call: cheap, and correct on any device
    convolve2d_kernel<<<(unsigned)count, 256, smem, st>>>((int)n, count_dev, pa, ai, pb, bi, po, conj_out);
    SK_LAUNCH_CHECK(ctx);
    return SKAGRID_OK;
}

int sk_convolve2d_dev(skagrid_ctx *ctx, i64 n, i64 count, const double *a, const i64 *ai, const double *b, const i64 *bi, double *out,
                      int conj_out, cudaStream_t st) {
    return conv_batch(ctx, n, count, nullptr, a, ai, b, bi, out, conj_out, st);
}

// ---------------------------------------------------------------------------------------------------------------
// aw_kernel_fn2 for a batch
// ---------------------------------------------------------------------------------------------------------------
// Pass 1: validate the indices of every visibility; the first visibility of each antenna pair claims a pair id.
// slot[a1*nant + a2]: 0 = unseen, 1 = being claimed, id + 2 = claimed.  slot == NULL: no de-duplication (id = k).
__global__ void __launch_bounds__(256) aw_mark_kernel(i64 count, i64 nw, i64 qpx, i64 nant, const i64 *__restrict__ wbin, const i64 *__restrict__ yf,
                                                      const i64 *__restrict__ xf, const i64 *__restrict__ a1, const i64 *__restrict__ a2,
                                                      uint32_t *__restrict__ slot, uint32_t *__restrict__ npairs, i64 *__restrict__ pa1,
                                                      i64 *__restrict__ pa2, i64 *__restrict__ ai, uint32_t *__restrict__ err_flag) {
    const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const i64 wb = wbin[k], iy = yf[k], ix = xf[k], i1 = a1[k], i2 = a2[k];
    if (wb < 0 || wb >= nw || iy < 0 || iy >= qpx || ix < 0 || ix >= qpx || i1 < 0 || i1 >= nant || i2 < 0 || i2 >= nant) {
        atomicOr(err_flag, 1u);
        ai[k] = -1;  // its kernel is zero
        return;
    }
    ai[k] = 0;
    if (!slot) {
        pa1[k] = i1;
        pa2[k] = i2;
    } else if (atomicCAS(&slot[i1 * nant + i2], 0u, 1u) == 0u) {
        const uint32_t id = atomicAdd(npairs, 1u);
        pa1[id] = i1;
        pa2[id] = i2;
        slot[i1 * nant + i2] = id + 2u;  // read by the next kernel only
    }
}

// Pass 2: per visibility, the pair id and the w-kernel slice
__global__ void __launch_bounds__(256) aw_index_kernel(i64 count, i64 qpx, i64 nant, const i64 *__restrict__ wbin, const i64 *__restrict__ yf,
                                                       const i64 *__restrict__ xf, const i64 *__restrict__ a1, const i64 *__restrict__ a2,
                                                       const uint32_t *__restrict__ slot, i64 *__restrict__ ai, i64 *__restrict__ bi) {
    const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    if (ai[k] < 0) { bi[k] = 0; return; }
    ai[k] = slot ? (i64)slot[a1[k] * nant + a2[k]] - 2 : k;
    bi[k] = (wbin[k] * qpx + yf[k]) * qpx + xf[k];
}

// out[k] = aw_kernel_fn2 yf[k] xf[k] wkerns[wbin[k]] akerns[a1[k]] akerns[a2[k]]   (optionally conjugated:
// processOne2, src/Gridding.hs:391, multiplies the visibility with conj(AW)).  Out-of-range indices set bit 0 of
// err_flag and give a zero kernel.
int sk_aw_kernels_dev(skagrid_ctx *ctx, i64 nw, i64 qpx, i64 s, const double *wkerns, i64 nant, const double *akerns, i64 count,
                      const i64 *wbin, const i64 *yf, const i64 *xf, const i64 *a1, const i64 *a2, double *out, int conj_out,
                      uint32_t *err_flag, cudaStream_t st) {
    if (count <= 0) return SKAGRID_OK;
    if (s <= 0 || s > 64) return sk_fail(ctx, SKAGRID_EINVAL, "aw_kernel: support %lld outside [1,64]", s);
    if (count >= (i64)0x7FFFFFFFll) return sk_fail(ctx, SKAGRID_EINVAL, "aw_kernel: batch too large");
    const bool dedupe = nant <= 2048;  // slot table of at most 16 MB
    const i64 max_pairs = dedupe ? std::min<i64>(count, nant * nant) : count;
    uint32_t *slot = nullptr, *npairs = nullptr;
    i64 *pa1, *pa2, *ai, *bi;
    double *pairs;
    if (dedupe) {
        SK_TRY(sk_scratch(ctx, "aw_slot", (size_t)(nant * nant) * 4, (void **)&slot));
        SK_TRY(sk_scratch(ctx, "aw_np", 16, (void **)&npairs));
        SK_CUDA(ctx, cudaMemsetAsync(slot, 0, (size_t)(nant * nant) * 4, st));
        SK_CUDA(ctx, cudaMemsetAsync(npairs, 0, 16, st));
    }
    SK_TRY(sk_scratch(ctx, "aw_pa1", (size_t)max_pairs * 8, (void **)&pa1));
    SK_TRY(sk_scratch(ctx, "aw_pa2", (size_t)max_pairs * 8, (void **)&pa2));
    SK_TRY(sk_scratch(ctx, "aw_ai", (size_t)count * 8, (void **)&ai));
    SK_TRY(sk_scratch(ctx, "aw_bi", (size_t)count * 8, (void **)&bi));
    SK_TRY(sk_scratch(ctx, "aw_pair", (size_t)(max_pairs * s * s) * 16, (void **)&pairs));
    const unsigned blocks = (unsigned)((count + 255) / 256);
    if (!dedupe) {  // ids are the visibility numbers: entries of invalid visibilities must still be valid antenna numbers
        SK_CUDA(ctx, cudaMemsetAsync(pa1, 0, (size_t)max_pairs * 8, st));
        SK_CUDA(ctx, cudaMemsetAsync(pa2, 0, (size_t)max_pairs * 8, st));
    }
    aw_mark_kernel<<<blocks, 256, 0, st>>>(count, nw, qpx, nant, wbin, yf, xf, a1, a2, slot, npairs, pa1, pa2, ai, err_flag);
    SK_LAUNCH_CHECK(ctx);
    aw_index_kernel<<<blocks, 256, 0, st>>>(count, qpx, nant, wbin, yf, xf, a1, a2, slot, ai, bi);
    SK_LAUNCH_CHECK(ctx);
    SK_TRY(conv_batch(ctx, s, max_pairs, npairs, akerns, pa1, akerns, pa2, pairs, 0, st));
    return conv_batch(ctx, s, count, nullptr, pairs, ai, wkerns, bi, out, conj_out, st);
}
