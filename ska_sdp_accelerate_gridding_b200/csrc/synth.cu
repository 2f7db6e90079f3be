// synth.cu -- synthetic SKA1-Low-shaped visibilities, generated on the device.
//
// Counter-based: every value is a pure function of (seed, visibility index, stream), so any slice of
// the data set can be produced on any GPU without communication (visibility-sharded runs generate
// [first, first+count) per rank).  The uv distribution is the core-dominated mixture of SURVEY 8d:
// 60 % sigma = 0.02, 30 % sigma = 0.08, 10 % sigma = 0.20 of the grid extent (per axis, approximately
// normal: sum of four uniforms), clipped so that the whole footprint stays on the grid, then mirrored to
// v >= 0 (mirror_uvw, src/Gridding.hs:551-562).  uniform != 0 draws uv uniformly instead.
#include "common.cuh"

__device__ __forceinline__ u64 splitmix64(u64 x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__device__ __forceinline__ double u01(u64 seed, u64 idx, u64 stream) {
    const u64 r = splitmix64(seed ^ splitmix64(idx * 64ull + stream));
    return (double)(r >> 11) * (1.0 / 9007199254740992.0);  // [0,1)
}

__device__ __forceinline__ double gauss4(u64 seed, u64 idx, u64 stream) {
    const double s = u01(seed, idx, stream) + u01(seed, idx, stream + 1) + u01(seed, idx, stream + 2) + u01(seed, idx, stream + 3);
    return (s - 2.0) * 1.7320508075688772;  // unit variance
}

__global__ void __launch_bounds__(256) synth_vis_kernel(u64 seed, i64 first, i64 count, double lim, i64 nw, int uniform,
                                                        double *__restrict__ u, double *__restrict__ v, i64 *__restrict__ wbin,
                                                        double *__restrict__ vis) {
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) {
        const u64 idx = (u64)(first + k);
        double px, py;
        if (uniform) {
            px = (2.0 * u01(seed, idx, 1) - 1.0) * lim;
            py = (2.0 * u01(seed, idx, 5) - 1.0) * lim;
        } else {
            const double c = u01(seed, idx, 0);
            const double sigma = c < 0.6 ? 0.02 : (c < 0.9 ? 0.08 : 0.20);
            px = sigma * gauss4(seed, idx, 1);
            py = sigma * gauss4(seed, idx, 5);
            for (int t = 1; t < 4 && (fabs(px) >= lim || fabs(py) >= lim); ++t) {
                px = sigma * gauss4(seed, idx, 1 + 16 * t);
                py = sigma * gauss4(seed, idx, 5 + 16 * t);
            }
            if (fabs(px) >= lim) px *= 0.25;
            if (fabs(py) >= lim) py *= 0.25;
        }
        if (py < 0.0) { px = -px; py = -py; }
        u[k] = px; v[k] = py;
        if (wbin) {
            i64 wb = (i64)(u01(seed, idx, 9) * (double)nw);
            wbin[k] = wb >= nw ? nw - 1 : wb;
        }
        if (vis) {
            vis[2 * k] = gauss4(seed, idx, 10);
            vis[2 * k + 1] = gauss4(seed, idx, 14 + 32);
        }
    }
}

extern "C" int skagrid_dev_synth_vis(skagrid_ctx *ctx, uint64_t seed, int64_t first, int64_t count, int64_t n, int64_t support,
                                     int64_t nw, int uniform, double *u, double *v, int64_t *wbin, double *vis, void *stream) {
    if (!ctx || !u || !v) return SKAGRID_EINVAL;
    if (count <= 0) return SKAGRID_OK;
    if (n <= 0 || support <= 0 || nw <= 0) return sk_fail(ctx, SKAGRID_EINVAL, "synth_vis: non-positive n/support/nw");
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    const double lim = 0.5 - ((double)(support / 2) + 1.0) / (double)n;
    if (!(lim > 0.0)) return sk_fail(ctx, SKAGRID_EINVAL, "synth_vis: support %lld does not fit a grid of %lld", (i64)support, (i64)n);
    i64 b = (count + 255) / 256;
    if (b > (i64)ctx->sm_count * 16) b = (i64)ctx->sm_count * 16;
    synth_vis_kernel<<<(unsigned)b, 256, 0, sk_stream(ctx, stream)>>>(seed, first, count, lim, nw, uniform, u, v, (i64 *)wbin, vis);
    SK_LAUNCH_CHECK(ctx);
    return SKAGRID_OK;
}
