// ctx.cu -- context lifetime, error strings, scratch pool, FP64 peak micro-benchmark.
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

static thread_local std::string g_create_err;

int sk_fail(skagrid_ctx *ctx, int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_create_err = buf;
    return code;
}

int sk_scratch(skagrid_ctx *ctx, const char *name, size_t bytes, void **out) {
    DevBuf &b = ctx->pool[name];
    if (b.bytes < bytes) {
        if (b.p) { cudaFree(b.p); b.p = nullptr; b.bytes = 0; }
        size_t want = bytes + bytes / 8 + 256;  // slack so slowly growing batches do not realloc
        cudaError_t e = cudaMalloc(&b.p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            e = cudaMalloc(&b.p, bytes);
            want = bytes;
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            b.p = nullptr;
            return sk_fail(ctx, SKAGRID_ENOMEM, "cudaMalloc(%zu bytes) for scratch '%s' failed: %s", bytes, name,
                           cudaGetErrorString(e));
        }
        b.bytes = want;
    }
    *out = b.p;
    return SKAGRID_OK;
}

// pinned host staging (small control data of the multi-device calls: histograms, counters), grown on demand
int sk_host_scratch(skagrid_ctx *ctx, size_t bytes, void **out) {
    if (ctx->h_pinned_bytes < bytes) {
        if (ctx->h_pinned) { cudaFreeHost(ctx->h_pinned); ctx->h_pinned = nullptr; ctx->h_pinned_bytes = 0; }
        const cudaError_t e = cudaMallocHost(&ctx->h_pinned, bytes);
        if (e != cudaSuccess) {
            cudaGetLastError();
            ctx->h_pinned = nullptr;
            return sk_fail(ctx, SKAGRID_ENOMEM, "cudaMallocHost(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        }
        ctx->h_pinned_bytes = bytes;
    }
    *out = ctx->h_pinned;
    return SKAGRID_OK;
}

extern "C" int skagrid_create(int device, skagrid_ctx **out) {
    if (!out) return SKAGRID_EINVAL;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return sk_fail(nullptr, SKAGRID_ENODEV, "no CUDA device (%s); libskagrid has no CPU fallback",
                       e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    }
    if (device < 0 || device >= ndev) return sk_fail(nullptr, SKAGRID_EINVAL, "device %d out of range [0,%d)", device, ndev);
    skagrid_ctx *ctx = new skagrid_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return sk_fail(nullptr, SKAGRID_ENODEV, "cudaSetDevice(%d) failed", device); }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return sk_fail(nullptr, SKAGRID_ECUDA, "cudaGetDeviceProperties failed"); }
    if (prop.major < 10) {
        delete ctx;
        return sk_fail(nullptr, SKAGRID_ENODEV, "device %d is sm_%d%d; libskagrid is built for sm_100a (B200) only", device, prop.major, prop.minor);
    }
    ctx->sm_count = prop.multiProcessorCount;
    bool ok = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreate(&ctx->ev0) == cudaSuccess && cudaEventCreate(&ctx->ev1) == cudaSuccess;
    for (int i = 0; i < 2 && ok; ++i)
        ok = cudaEventCreateWithFlags(&ctx->ev_copy[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->ev_k[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->ev_d2h[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->ev_mg[i], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaMalloc(&ctx->d_flags, 16 * sizeof(uint32_t)) == cudaSuccess && cudaMemset(ctx->d_flags, 0, 16 * sizeof(uint32_t)) == cudaSuccess;
    if (!ok) { skagrid_destroy(ctx); return sk_fail(nullptr, SKAGRID_ECUDA, "stream/event creation failed"); }
    *out = ctx;
    return SKAGRID_OK;
}

extern "C" void skagrid_destroy(skagrid_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->cached_plan) { sk_plan_free(ctx->cached_plan); ctx->cached_plan = nullptr; }
    for (auto &kv : ctx->fft_plans) cufftDestroy(kv.second);
    for (auto &kv : ctx->fft_work) if (kv.second.p) cudaFree(kv.second.p);
    for (auto &kv : ctx->pool) if (kv.second.p) cudaFree(kv.second.p);
    for (int i = 0; i < 2; ++i) {
        if (ctx->ev_copy[i]) cudaEventDestroy(ctx->ev_copy[i]);
        if (ctx->ev_done[i]) cudaEventDestroy(ctx->ev_done[i]);
        if (ctx->ev_k[i]) cudaEventDestroy(ctx->ev_k[i]);
        if (ctx->ev_d2h[i]) cudaEventDestroy(ctx->ev_d2h[i]);
        if (ctx->ev_mg[i]) cudaEventDestroy(ctx->ev_mg[i]);
    }
    if (ctx->d_flags) cudaFree(ctx->d_flags);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char *skagrid_last_error(const skagrid_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }
extern "C" const char *skagrid_version(void) { return "skagrid-b200 0.1 (sm_100a)"; }
extern "C" double skagrid_last_device_ms(const skagrid_ctx *ctx) { return ctx ? ctx->last_ms : 0.0; }
extern "C" int64_t skagrid_launch_count(const skagrid_ctx *ctx) { return ctx ? ctx->launches : 0; }

// ---------------------------------------------------------------------------------------------
// FP64 FMA peak: 8 independent dependent-chains of DFMA per thread, all in registers.
// MEASURED_PEAKS.json has no FP64 figure; the roofline of the gridder is FP64-vector bound
// (SURVEY.md 8d), so the denominator is measured here, on the same device, in the same run.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 12345.678) out[0] = s;  // never true; keeps the chain alive
}

extern "C" int skagrid_measure_fp64_tflops(skagrid_ctx *ctx, double *tflops) {
    if (!ctx || !tflops) return SKAGRID_EINVAL;
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    void *out;
    SK_TRY(sk_scratch(ctx, "dfma_out", 64, &out));
    const int blocks = ctx->sm_count * 8, iters = 4096;
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        SK_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        dfma_peak_kernel<<<blocks, 256, 0, ctx->stream>>>((double *)out, iters, 1.0 + rep);
        SK_LAUNCH_CHECK(ctx);
        SK_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        SK_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
        float ms = 0;
        SK_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        const double flop = 2.0 * 64.0 * iters * 256.0 * blocks;
        const double tf = flop / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    *tflops = best;
    return SKAGRID_OK;
}

// ---------------------------------------------------------------------------------------------
// L2 -> SM read bandwidth: every warp streams 128-bit loads (ld.global.cg: L2 only, no L1 allocation) over a
// buffer that fits the L2, eight independent loads in flight per thread.  This is the ceiling the tiled gridder
// and the degridder run against (their kernel taps come from an L2-resident table with ~2 % L1 hits).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) l2_read_kernel(const double2 *__restrict__ buf, size_t n, int iters, double *out) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    double sx = 0.0, sy = 0.0;
    size_t idx = tid % n;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
        double2 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            v[u] = __ldcg(buf + idx);
            idx += nthreads;
            if (idx >= n) idx -= n;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) { sx += v[u].x; sy += v[u].y; }
    }
    if (sx == 1.2345e300 && sy == 1.0) out[0] = sx;  // never true; keeps the loads alive
}

// The same measurement with the gridder's / degridder's access pattern: every half-warp reads the 15 (of 16) taps of
// 15 consecutive 256-byte rows of a pseudo-randomly chosen 3840-byte slice, slice after slice -- no arithmetic, no
// other traffic.  This is the ceiling of "kernel taps streamed from an L2-resident table" for S = 15.
__global__ void __launch_bounds__(256) l2_slice_read_kernel(const double2 *__restrict__ buf, unsigned nslices, int iters, double *out) {
    const unsigned hw = (blockIdx.x * blockDim.x + threadIdx.x) >> 4, lane = threadIdx.x & 15;
    unsigned state = hw * 2654435761u + 12345u;
    double sx = 0.0, sy = 0.0;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
        state = state * 1664525u + 1013904223u;
        const double2 *p = buf + (size_t)((state >> 8) % nslices) * 240 + lane;
        double2 v[15];
#pragma unroll
        for (int r = 0; r < 15; ++r) v[r] = lane < 15 ? __ldcg(p + r * 16) : make_double2(0.0, 0.0);
#pragma unroll
        for (int r = 0; r < 15; ++r) { sx += v[r].x; sy += v[r].y; }
    }
    if (sx == 1.2345e300 && sy == 1.0) out[0] = sx;
}

// pattern: 0 = fully coalesced stream (skagrid_measure_l2_read_tbs), 1 = random 15x15-tap slices
extern "C" int skagrid_measure_l2_pattern_tbs(skagrid_ctx *ctx, int64_t bytes, int pattern, double *tbs) {
    if (pattern == 0) return skagrid_measure_l2_read_tbs(ctx, bytes, tbs);
    if (!ctx || !tbs || bytes < (1 << 20) || bytes > ((int64_t)1 << 30)) return SKAGRID_EINVAL;
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    void *buf, *out;
    SK_TRY(sk_scratch(ctx, "l2_buf", (size_t)bytes, &buf));
    SK_TRY(sk_scratch(ctx, "dfma_out", 64, &out));
    SK_CUDA(ctx, cudaMemsetAsync(buf, 0, (size_t)bytes, ctx->stream));
    const unsigned nslices = (unsigned)((size_t)bytes / 3840);
    const int blocks = ctx->sm_count * 8, iters = 512;
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        SK_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        l2_slice_read_kernel<<<blocks, 256, 0, ctx->stream>>>((const double2 *)buf, nslices, iters, (double *)out);
        SK_LAUNCH_CHECK(ctx);
        SK_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        SK_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
        float ms = 0;
        SK_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        const double moved = 3600.0 * iters * 16.0 * blocks;  // useful bytes: 225 taps x 16 B per half-warp and iteration
        const double t = moved / (ms * 1e-3) / 1e12;
        if (rep > 0 && t > best) best = t;
    }
    *tbs = best;
    return SKAGRID_OK;
}

extern "C" int skagrid_measure_l2_read_tbs(skagrid_ctx *ctx, int64_t bytes, double *tbs) {
    if (!ctx || !tbs || bytes < (1 << 20) || bytes > ((int64_t)1 << 30)) return SKAGRID_EINVAL;
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    void *buf, *out;
    SK_TRY(sk_scratch(ctx, "l2_buf", (size_t)bytes, &buf));
    SK_TRY(sk_scratch(ctx, "dfma_out", 64, &out));
    SK_CUDA(ctx, cudaMemsetAsync(buf, 0, (size_t)bytes, ctx->stream));
    const size_t n = (size_t)bytes / sizeof(double2);
    const int blocks = ctx->sm_count * 8, iters = 256;
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        SK_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        l2_read_kernel<<<blocks, 256, 0, ctx->stream>>>((const double2 *)buf, n, iters, (double *)out);
        SK_LAUNCH_CHECK(ctx);
        SK_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        SK_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
        float ms = 0;
        SK_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        const double moved = 16.0 * 8.0 * iters * 256.0 * blocks;
        const double t = moved / (ms * 1e-3) / 1e12;
        if (rep > 0 && t > best) best = t;
    }
    *tbs = best;
    return SKAGRID_OK;
}
