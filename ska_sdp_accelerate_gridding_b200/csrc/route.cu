// route.cu -- device side of the uv-tile-sharded mode when every GPU has its own process (torch.distributed / NCCL does the
// exchange): row histogram for the slab balance, per-destination counts, packing of the records into destination-major
// send buffers, and the accumulation of returned degridding partial sums.
//
// Ownership rule (bit-exact with the reference's binning): a visibility's footprint covers grid rows
// [y - gh/2, y - gh/2 + gh) with (y, yf) = frac_coord(height, qpx, v) (src/Gridding.hs:126-140); it is sent to every rank g
// whose slab [bounds[g], bounds[g+1]) intersects those rows (clipped to the grid: fixoutofbounds, :883-891).  The owner
// then clips taps to its slab, so a footprint straddling two slabs is gridded once, half by each owner.
//
// mgpu.cu holds the single-process form of the same steps (one host thread, peer copies instead of NCCL).
#include <algorithm>

#include "common.cuh"

constexpr int RT_MAX = 64;  // ranks

struct RtBounds {
    i64 b[RT_MAX + 1];
    int n;
};
struct RtSeg {
    uint32_t s[RT_MAX];
};

__device__ __forceinline__ bool rt_rows(double pv, double halfhf, double hf, double qpxf, double qpxfrac, i64 qpx, i64 height, i64 gh, i64 &y,
                                        i64 &oy) {
    if (!(fabs(pv) < 1.0e9)) return false;  // NaN / inf / absurd: no tap on the grid (as bin_vis)
    i64 yf;
    frac_coord_one(pv, halfhf, hf, qpxf, qpxfrac, qpx, 1, y, yf);
    oy = y - gh / 2;
    return oy + gh > 0 && oy < height;
}
__device__ __forceinline__ int rt_owner(const RtBounds &B, i64 row) {
    int g = 0;
    for (int k = 1; k < B.n; ++k) g += (B.b[k] <= row) ? 1 : 0;
    return g;
}

// hist[row of the footprint centre, clamped to the grid] += 1
__global__ void __launch_bounds__(256) rt_row_hist_kernel(i64 count, const double *__restrict__ v, i64 height, i64 qpx, i64 gh,
                                                          uint32_t *__restrict__ hist) {
    const double halfhf = (double)(height / 2), hf = (double)height, qpxf = (double)qpx, qpxfrac = 0.5 / (double)qpx;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) {
        i64 y, oy;
        if (!rt_rows(v[k], halfhf, hf, qpxf, qpxfrac, qpx, height, gh, y, oy)) continue;
        atomicAdd(&hist[min(max(y, (i64)0), height - 1)], 1u);
    }
}

// Both passes walk the visibilities with the SAME block / iteration mapping, so the records a block sends to every rank are
// known after pass 1 and pass 2 needs no global atomics at all (a same-address atomic per warp and destination serialises:
// 5.8 + 8.3 ms per 1.25e8 visibilities in the first version, profiles/r02_config5_substages.md):
//   pass 1  every block counts its records per destination in shared memory -> blockcounts[block][g]
//   scan    bases[block][g] = records of the blocks before it for destination g; totals[g]
//   pass 2  the block's shared-memory cursors start at seg[g] + bases[block][g]; a warp reserves its run with one
//           shared-memory atomic per destination and appends `W` doubles {u, v, wbin, [re, im]} per record; sidx (optional)
//           keeps the source index of every record.
template <int W, bool PACK>
__global__ void __launch_bounds__(256) rt_route_kernel(i64 count, const double *__restrict__ u, const double *__restrict__ v,
                                                       const i64 *__restrict__ wbin, const double2 *__restrict__ vis, i64 height, i64 qpx, i64 gh,
                                                       RtBounds B, uint32_t *__restrict__ blockcounts, const uint32_t *__restrict__ bases, RtSeg seg,
                                                       double *__restrict__ send, uint32_t *__restrict__ sidx) {
    __shared__ uint32_t s_cur[RT_MAX];
    const double halfhf = (double)(height / 2), hf = (double)height, qpxf = (double)qpx, qpxfrac = 0.5 / (double)qpx;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    if ((int)threadIdx.x < B.n) s_cur[threadIdx.x] = PACK ? seg.s[threadIdx.x] + bases[(size_t)blockIdx.x * B.n + threadIdx.x] : 0u;
    __syncthreads();
    for (i64 base = (i64)blockIdx.x * blockDim.x; base < count; base += stride) {  // warp-uniform trip count
        const i64 k = base + threadIdx.x;
        int lo = 1, hi = 0;
        if (k < count) {
            i64 y, oy;
            if (rt_rows(v[k], halfhf, hf, qpxf, qpxfrac, qpx, height, gh, y, oy)) {
                lo = rt_owner(B, max(oy, (i64)0));
                hi = rt_owner(B, min(oy + gh - 1, height - 1));
            }
        }
        // destinations any lane of this warp sends to: [wlo, whi]
        int wlo = lo <= hi ? lo : RT_MAX, whi = lo <= hi ? hi : -1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            wlo = min(wlo, __shfl_xor_sync(0xffffffffu, wlo, o));
            whi = max(whi, __shfl_xor_sync(0xffffffffu, whi, o));
        }
        for (int g = wlo; g <= whi; ++g) {
            const bool mine = lo <= g && g <= hi;
            const unsigned m = __ballot_sync(0xffffffffu, mine);
            if (!m) continue;
            const int leader = __ffs(m) - 1;
            uint32_t pos = 0;
            if (lane == leader) pos = atomicAdd(&s_cur[g], (uint32_t)__popc(m));
            if (!PACK) continue;
            pos = __shfl_sync(0xffffffffu, pos, leader) + (uint32_t)__popc(m & ((1u << lane) - 1u));
            if (mine) {
                double *r = send + (size_t)pos * W;
                r[0] = u[k];
                r[1] = v[k];
                r[2] = __longlong_as_double(wbin ? wbin[k] : 0);
                if constexpr (W == 5) {
                    const double2 x = vis[k];
                    r[3] = x.x;
                    r[4] = x.y;
                }
                if (sidx) sidx[pos] = (uint32_t)k;
            }
        }
    }
    if (!PACK) {
        __syncthreads();
        if ((int)threadIdx.x < B.n) blockcounts[(size_t)blockIdx.x * B.n + threadIdx.x] = s_cur[threadIdx.x];
    }
}

// bases[b][g] = sum of blockcounts[b'][g] over b' < b; totals[g] = the column sum.  One thread per destination.
__global__ void rt_scan_kernel(const uint32_t *__restrict__ blockcounts, uint32_t *__restrict__ bases, uint32_t *__restrict__ totals, int nblocks, int n) {
    const int g = threadIdx.x;
    if (g >= n) return;
    uint32_t run = 0;
    for (int b = 0; b < nblocks; ++b) {
        const uint32_t c = blockcounts[(size_t)b * n + g];
        bases[(size_t)b * n + g] = run;
        run += c;
    }
    totals[g] = run;
}

// out[sidx[i]] += back[i]
__global__ void __launch_bounds__(256) rt_scatter_add_kernel(i64 n, const uint32_t *__restrict__ sidx, const double2 *__restrict__ back,
                                                             double *__restrict__ out) {
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double2 b = back[i];
        const i64 k = sidx[i];
        atomicAdd(&out[2 * k], b.x);
        atomicAdd(&out[2 * k + 1], b.y);
    }
}

static unsigned rt_blocks(skagrid_ctx *ctx, i64 n) { return (unsigned)std::max<i64>(1, std::min<i64>((n + 255) / 256, (i64)ctx->sm_count * 8)); }

static int rt_bounds(skagrid_ctx *ctx, const char *what, int64_t height, int nranks, const int64_t *bounds, RtBounds *B) {
    if (nranks < 1 || nranks > RT_MAX) return sk_fail(ctx, SKAGRID_EINVAL, "%s: 1..%d ranks", what, RT_MAX);
    if (!bounds) return sk_fail(ctx, SKAGRID_EINVAL, "%s: NULL bounds", what);
    for (int g = 0; g <= nranks; ++g) B->b[g] = bounds[g];
    for (int g = 0; g < nranks; ++g)
        if (bounds[g] > bounds[g + 1]) return sk_fail(ctx, SKAGRID_EINVAL, "%s: bounds must not decrease", what);
    if (bounds[0] != 0 || bounds[nranks] != height) return sk_fail(ctx, SKAGRID_EINVAL, "%s: bounds must run from 0 to the grid height", what);
    B->n = nranks;
    return SKAGRID_OK;
}

extern "C" int skagrid_dev_row_hist(skagrid_ctx *ctx, int64_t height, int64_t qpx, int64_t gh, int64_t count, const double *d_v,
                                    uint32_t *d_hist, void *stream) {
    SK_TRY(sk_api_enter(ctx));
    if (height <= 0 || qpx <= 0 || gh <= 0) return sk_fail(ctx, SKAGRID_EINVAL, "dev_row_hist: non-positive dimension");
    if (count <= 0) return SKAGRID_OK;
    if (!d_v || !d_hist) return sk_fail(ctx, SKAGRID_EINVAL, "dev_row_hist: NULL pointer");
    rt_row_hist_kernel<<<rt_blocks(ctx, count), 256, 0, sk_stream(ctx, stream)>>>(count, d_v, height, qpx, gh, d_hist);
    SK_LAUNCH_CHECK(ctx);
    return SKAGRID_OK;
}

// per-block counts and bases of the last route_count on this context (consumed by route_pack)
static int rt_state(skagrid_ctx *ctx, unsigned blocks, int nranks, uint32_t **blockcounts, uint32_t **bases) {
    const size_t bytes = (size_t)blocks * nranks * sizeof(uint32_t);
    SK_TRY(sk_scratch(ctx, "rt_blockcounts", bytes, (void **)blockcounts));
    SK_TRY(sk_scratch(ctx, "rt_bases", bytes, (void **)bases));
    return SKAGRID_OK;
}

extern "C" int skagrid_dev_route_count(skagrid_ctx *ctx, int64_t height, int64_t qpx, int64_t gh, int nranks, const int64_t *bounds,
                                       int64_t count, const double *d_v, uint32_t *d_counts, void *stream) {
    SK_TRY(sk_api_enter(ctx));
    RtBounds B;
    SK_TRY(rt_bounds(ctx, "dev_route_count", height, nranks, bounds, &B));
    if (height <= 0 || qpx <= 0 || gh <= 0) return sk_fail(ctx, SKAGRID_EINVAL, "dev_route_count: non-positive dimension");
    if (!d_counts) return sk_fail(ctx, SKAGRID_EINVAL, "dev_route_count: NULL counts");
    cudaStream_t st = sk_stream(ctx, stream);
    SK_CUDA(ctx, cudaMemsetAsync(d_counts, 0, (size_t)nranks * sizeof(uint32_t), st));
    if (count <= 0) return SKAGRID_OK;
    if (!d_v) return sk_fail(ctx, SKAGRID_EINVAL, "dev_route_count: NULL v");
    const unsigned blocks = rt_blocks(ctx, count);
    uint32_t *blockcounts, *bases;
    SK_TRY(rt_state(ctx, blocks, nranks, &blockcounts, &bases));
    RtSeg S;
    for (int g = 0; g < RT_MAX; ++g) S.s[g] = 0;
    rt_route_kernel<3, false><<<blocks, 256, 0, st>>>(count, nullptr, d_v, nullptr, nullptr, height, qpx, gh, B, blockcounts, nullptr, S, nullptr, nullptr);
    SK_LAUNCH_CHECK(ctx);
    rt_scan_kernel<<<1, RT_MAX, 0, st>>>(blockcounts, bases, d_counts, (int)blocks, nranks);
    SK_LAUNCH_CHECK(ctx);
    return SKAGRID_OK;
}

// Must follow skagrid_dev_route_count of the same (count, d_v, bounds) on this context: it consumes the per-block bases that
// call left in the context (no global atomics in the packing pass).
extern "C" int skagrid_dev_route_pack(skagrid_ctx *ctx, int64_t height, int64_t qpx, int64_t gh, int nranks, const int64_t *bounds,
                                      int64_t count, const double *d_u, const double *d_v, const int64_t *d_wbin, const double *d_vis,
                                      const int64_t *seg, double *d_send, uint32_t *d_sidx, void *stream) {
    SK_TRY(sk_api_enter(ctx));
    RtBounds B;
    SK_TRY(rt_bounds(ctx, "dev_route_pack", height, nranks, bounds, &B));
    if (height <= 0 || qpx <= 0 || gh <= 0) return sk_fail(ctx, SKAGRID_EINVAL, "dev_route_pack: non-positive dimension");
    if (!seg) return sk_fail(ctx, SKAGRID_EINVAL, "dev_route_pack: NULL segment starts");
    if (count <= 0) return SKAGRID_OK;
    if (!d_u || !d_v || !d_send) return sk_fail(ctx, SKAGRID_EINVAL, "dev_route_pack: NULL pointer");
    RtSeg S;
    for (int g = 0; g < RT_MAX; ++g) S.s[g] = 0;
    for (int g = 0; g < nranks; ++g) {
        if (seg[g] < 0 || seg[g] >= (int64_t)0xFFFFFFF0ll) return sk_fail(ctx, SKAGRID_EINVAL, "dev_route_pack: segment start out of range");
        S.s[g] = (uint32_t)seg[g];
    }
    cudaStream_t st = sk_stream(ctx, stream);
    const unsigned blocks = rt_blocks(ctx, count);
    uint32_t *blockcounts, *bases;
    SK_TRY(rt_state(ctx, blocks, nranks, &blockcounts, &bases));
    if (d_vis)
        rt_route_kernel<5, true><<<blocks, 256, 0, st>>>(count, d_u, d_v, (const i64 *)d_wbin, (const double2 *)d_vis, height, qpx, gh, B, nullptr, bases, S,
                                                         d_send, d_sidx);
    else
        rt_route_kernel<3, true><<<blocks, 256, 0, st>>>(count, d_u, d_v, (const i64 *)d_wbin, nullptr, height, qpx, gh, B, nullptr, bases, S, d_send, d_sidx);
    SK_LAUNCH_CHECK(ctx);
    return SKAGRID_OK;
}

extern "C" int skagrid_dev_scatter_add(skagrid_ctx *ctx, int64_t n, const uint32_t *d_sidx, const double *d_back, double *d_out, void *stream) {
    SK_TRY(sk_api_enter(ctx));
    if (n <= 0) return SKAGRID_OK;
    if (!d_sidx || !d_back || !d_out) return sk_fail(ctx, SKAGRID_EINVAL, "dev_scatter_add: NULL pointer");
    rt_scatter_add_kernel<<<rt_blocks(ctx, n), 256, 0, sk_stream(ctx, stream)>>>(n, d_sidx, (const double2 *)d_back, d_out);
    SK_LAUNCH_CHECK(ctx);
    return SKAGRID_OK;
}
