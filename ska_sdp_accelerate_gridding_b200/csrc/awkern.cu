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

// Round 2: row-pair kernel.  Two changes against conv_tiled_kernel, same sums in the same order (bit-identical
// results for finite operands):
//   * a warp owns ONE column group, so tx0 is a compile-time constant of the code path it runs and the taps whose
//     operand column tx+C-kx lies outside the kernel are not issued at all: 169 complex MACs per (ky, output row)
//     instead of 240 for N = 15 (group by group 38 / 54 / 50 / 27 of 60) -- the x half of the 1.77x zero padding;
//   * a thread owns TWO adjacent output rows.  Output row ty at step ky reads operand row ty+C-ky, so the window the
//     lower output row loads at step ky is the one the upper row needs at step ky+1: one window load per step feeds
//     2 x TX outputs, and the aT loads are shared by 8 outputs instead of 4 (0.2 shared-memory loads per complex MAC).
// Lanes: (N+1)/2 row pairs of one output kernel, 32 / ((N+1)/2) output kernels per warp (N = 15: a quarter-warp per
// kernel, so the aT loads are quarter-uniform); operand rows are stored even rows first, then odd rows, with a pitch
// of 1 mod 8 taps: the rows a quarter-warp loads in one step have the same parity, hence consecutive slots and 8
// distinct 16-byte bank groups.  Rows outside the kernel map to one all-zero row (the y half of the padding stays).
template <int N>
struct ConvPair {
    static_assert(N % 2 == 1, "odd supports only");
    static constexpr int TX = 4;
    static constexpr int C = N / 2;
    static constexpr int GX = (N + TX - 1) / TX;            // column groups = warps per block
    static constexpr int RP = (N + 1) / 2;                  // row pairs per output kernel
    static constexpr int HR = (N + 1) / 2;                  // even operand rows (slots 0..HR-1), odd rows follow
    static constexpr int KPW = 32 / RP;                     // output kernels per warp = per block
    static constexpr int PW = (N + 6) / 8 * 8 + 1;          // operand row pitch: >= N, 1 mod 8
    static constexpr int THREADS = GX * 32;
    static constexpr int SM_PER = N * N + (N + 1) * PW;     // double2 per output kernel: aT, bT rows + the zero row
};

template <int N, int G>
struct ConvPairGroup {
    using T = ConvPair<N>;
    static constexpr int TX0 = G * T::TX;
    static constexpr int TXN = (TX0 + T::TX <= N) ? T::TX : N - TX0;                        // outputs of this group that exist
    static constexpr int QLO = (TX0 + T::C - (N - 1) > 0) ? TX0 + T::C - (N - 1) : 0;       // operand columns the group touches
    static constexpr int QHI = (TX0 + TXN - 1 + T::C < N - 1) ? TX0 + TXN - 1 + T::C : N - 1;
    static constexpr int WN = QHI - QLO + 1;
};

// acc[r][t] += sum_kx aT[ky][kx] * row_r[tx0 + t + C - kx]  for r = 0 (window lo) and 1 (window hi)
template <int N, int G>
__device__ __forceinline__ void conv_pair_mac(const double2 *__restrict__ arow, const double2 (&lo)[ConvPairGroup<N, G>::WN],
                                              const double2 (&hi)[ConvPairGroup<N, G>::WN], double (&accr)[2][4], double (&acci)[2][4]) {
    using GRP = ConvPairGroup<N, G>;
    constexpr int C = N / 2;
#pragma unroll
    for (int kx = 0; kx < N; ++kx) {
        if (GRP::TX0 + GRP::TXN - 1 + C - kx < 0 || GRP::TX0 + C - kx >= N) continue;      // no output of the group uses this tap
        const double2 p = arow[kx];
#pragma unroll
        for (int t = 0; t < GRP::TXN; ++t) {
            const int qx = GRP::TX0 + t + C - kx;
            if (qx < 0 || qx >= N) continue;
            const double2 q0 = lo[qx - GRP::QLO], q1 = hi[qx - GRP::QLO];
            accr[0][t] = fma(p.x, q0.x, accr[0][t]); accr[0][t] = fma(-p.y, q0.y, accr[0][t]);
            acci[0][t] = fma(p.x, q0.y, acci[0][t]); acci[0][t] = fma(p.y, q0.x, acci[0][t]);
            accr[1][t] = fma(p.x, q1.x, accr[1][t]); accr[1][t] = fma(-p.y, q1.y, accr[1][t]);
            acci[1][t] = fma(p.x, q1.y, acci[1][t]); acci[1][t] = fma(p.y, q1.x, acci[1][t]);
        }
    }
}

template <int N, int G>
__device__ __forceinline__ void conv_pair_group(const double2 *__restrict__ saT, const double2 *__restrict__ sbT, int rp, double2 *__restrict__ po,
                                                int conj_out) {
    using T = ConvPair<N>;
    using GRP = ConvPairGroup<N, G>;
    constexpr int C = T::C, WN = GRP::WN;
    auto row = [&](int q) { return sbT + ((q >= 0 && q < N) ? ((q >> 1) + (q & 1) * T::HR) : N) * T::PW + GRP::QLO; };
    double accr[2][4], acci[2][4];
#pragma unroll
    for (int t = 0; t < 4; ++t) accr[0][t] = accr[1][t] = acci[0][t] = acci[1][t] = 0.0;
    double2 lo[WN], hi[WN];
    const int q00 = 2 * rp + C;  // operand row of the lower output row at ky = 0
    {
        const double2 *r = row(q00 + 1);
#pragma unroll
        for (int i = 0; i < WN; ++i) hi[i] = r[i];
    }
    // One copy of the multiply-accumulate code per column group: the step's window becomes the upper row's window of
    // the next step by register moves (free next to 8 DFMA per tap on an FP64-bound loop).  Unrolling two steps to
    // rotate the windows by renaming instead doubled the loop body to 46 KB over the four groups -- beyond the 32 KB
    // instruction cache of the SM, and the kernel ran slower than round 1's (warps stalled on instruction fetch).
#pragma unroll 1
    for (int ky = 0; ky < N; ++ky) {
        const double2 *r = row(q00 - ky);
#pragma unroll
        for (int i = 0; i < WN; ++i) lo[i] = r[i];
        conv_pair_mac<N, G>(saT + ky * N, lo, hi, accr, acci);
#pragma unroll
        for (int i = 0; i < WN; ++i) hi[i] = lo[i];
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        if (2 * rp + r >= N) break;
#pragma unroll
        for (int t = 0; t < GRP::TXN; ++t) po[(2 * rp + r) * N + GRP::TX0 + t] = make_double2(accr[r][t], conj_out ? -acci[r][t] : acci[r][t]);
    }
}

template <int N, int G>
__device__ __forceinline__ void conv_pair_dispatch(int g, const double2 *__restrict__ saT, const double2 *__restrict__ sbT, int rp,
                                                   double2 *__restrict__ po, int conj_out) {
    if (g == G) conv_pair_group<N, G>(saT, sbT, rp, po, conj_out);
    else if constexpr (G + 1 < ConvPair<N>::GX) conv_pair_dispatch<N, G + 1>(g, saT, sbT, rp, po, conj_out);
}

template <int N>
__global__ void __launch_bounds__(ConvPair<N>::THREADS) conv_pair_kernel(i64 count, const uint32_t *__restrict__ count_dev,
                                                                         const double2 *__restrict__ a, const i64 *__restrict__ ai,
                                                                         const double2 *__restrict__ b, const i64 *__restrict__ bi,
                                                                         double2 *__restrict__ out, int conj_out) {
    using T = ConvPair<N>;
    constexpr int N2 = N * N;
    extern __shared__ double2 sm[];
    __shared__ i64 s_ia[T::KPW], s_ib[T::KPW];
    const i64 k0 = (i64)blockIdx.x * T::KPW;
    const i64 limit = count_dev ? min(count, (i64)*count_dev) : count;
    if (threadIdx.x < T::KPW) {
        const i64 k = k0 + threadIdx.x;
        const i64 ia = k < limit ? (ai ? ai[k] : k) : -1;
        s_ia[threadIdx.x] = ia;
        s_ib[threadIdx.x] = ia >= 0 ? (bi ? bi[k] : k) : -1;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < T::KPW * N2; idx += T::THREADS) {
        const int s = idx / N2, t = idx % N2;
        const i64 ia = s_ia[s];
        if (ia < 0) continue;
        double2 *saT = sm + s * T::SM_PER, *sbT = saT + N2;
        const int r = t % N, c = t / N;                                     // element [c, r] of the operands
        saT[r * N + c] = a[ia * N2 + t];                                    // aT[ky][kx] = a[kx,ky]
        sbT[((r >> 1) + (r & 1) * T::HR) * T::PW + c] = b[s_ib[s] * N2 + t];  // bT[qy][qx] = b[qx,qy], even rows first
    }
    for (int idx = threadIdx.x; idx < T::KPW * T::PW; idx += T::THREADS)
        sm[(idx / T::PW) * T::SM_PER + N2 + N * T::PW + idx % T::PW] = make_double2(0.0, 0.0);  // the zero row
    __syncthreads();
    const int g = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = lane / T::RP, rp = lane % T::RP;
    if (s >= T::KPW || k0 + s >= limit) return;
    double2 *po = out + (k0 + s) * N2;
    if (s_ia[s] < 0) {  // a visibility whose indices were out of range: zero kernel
        constexpr int TX = T::TX;
        for (int r = 0; r < 2; ++r)
            for (int t = 0; t < TX; ++t)
                if (2 * rp + r < N && g * TX + t < N) po[(2 * rp + r) * N + g * TX + t] = make_double2(0.0, 0.0);
        return;
    }
    const double2 *saT = sm + s * T::SM_PER;
    conv_pair_dispatch<N, 0>(g, saT, saT + N2, rp, po, conj_out);
}

template <int N>
static int launch_pair(skagrid_ctx *ctx, i64 count, const uint32_t *count_dev, const double2 *a, const i64 *ai, const double2 *b, const i64 *bi,
                       double2 *out, int conj_out, cudaStream_t st) {
    using T = ConvPair<N>;
    const size_t smem = (size_t)T::KPW * T::SM_PER * sizeof(double2);
    static_assert((size_t)T::KPW * T::SM_PER * sizeof(double2) <= 47 * 1024, "conv_pair_kernel: dynamic shared memory above the default limit");
    const i64 blocks = (count + T::KPW - 1) / T::KPW;
    conv_pair_kernel<N><<<(unsigned)blocks, T::THREADS, smem, st>>>(count, count_dev, a, ai, b, bi, out, conj_out);
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
    static const bool round1_tiled = getenv("SKAGRID_CONV_TILED") && atoi(getenv("SKAGRID_CONV_TILED")) != 0;     // the round-1 kernel
    if (!generic_only && !round1_tiled) {
        switch (n) {
            case 5: return launch_pair<5>(ctx, count, count_dev, pa, ai, pb, bi, po, conj_out, st);
            case 7: return launch_pair<7>(ctx, count, count_dev, pa, ai, pb, bi, po, conj_out, st);
            case 9: return launch_pair<9>(ctx, count, count_dev, pa, ai, pb, bi, po, conj_out, st);
            case 11: return launch_pair<11>(ctx, count, count_dev, pa, ai, pb, bi, po, conj_out, st);
            case 13: return launch_pair<13>(ctx, count, count_dev, pa, ai, pb, bi, po, conj_out, st);
            case 15: return launch_pair<15>(ctx, count, count_dev, pa, ai, pb, bi, po, conj_out, st);
            case 17: return launch_pair<17>(ctx, count, count_dev, pa, ai, pb, bi, po, conj_out, st);
            default: break;
        }
    }
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
