// mgpu.cu -- multi-GPU entry points of the host-pointer ABI (include/skagrid.h, "multi-GPU, single process").
//
// A Haskell (or C) caller has no torch.distributed: ONE host thread drives `nctx` contexts, one per device, and every
// step is only enqueued (streams + events), so the devices run concurrently.  Two partitionings:
//
//   visibility-sharded  device d grids the d-th contiguous share of the visibilities into a full local grid
//                       (gridding is linear in the visibilities: permute (+), src/Gridding.hs:377), then owns row slab d
//                       of the sum: it pulls that slab from every peer over NVLink peer memory (one kernel, all peers
//                       in flight), and the slabs leave for the host over P PCIe links in parallel.  An all-gather of
//                       the reduced slabs (copy engines) leaves the full grid resident on every context.
//   uv-tile-sharded     device d owns grid rows [bounds[d], bounds[d+1]) (bounds = quantiles of the footprint-row
//                       histogram: SKA1-Low coverage is core-dominated, equal-height slabs would idle most devices).
//                       Every visibility is routed to each owner its footprint rows intersect; owners clip taps to
//                       their slab (the rule of fixoutofbounds, src/Gridding.hs:883-891).  No grid reduction.
//
// Contexts on the same device are allowed (the exchange then degenerates to device-local copies), which is how the
// single-GPU test-suite exercises this file.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"

constexpr int MG_MAX = 16;

struct MgPeers {
    const double2 *p[MG_MAX];
    int n;
};
struct MgBounds {
    i64 b[MG_MAX + 1];
    int n;
};

// own[i] += sum over peers of peer[i]; the peer pointers are peer-device memory read over NVLink
__global__ void __launch_bounds__(256) mg_slab_sum_kernel(double2 *__restrict__ own, MgPeers peers, i64 n) {
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double2 a = own[i];
        for (int k = 0; k < peers.n; ++k) {
            const double2 b = peers.p[k][i];
            a.x += b.x;
            a.y += b.y;
        }
        own[i] = a;
    }
}

// footprint rows of one visibility: [oy, oy + gh).  false: non-finite coordinate or no row on the grid.
__device__ __forceinline__ bool mg_rows(double pv, double halfhf, double hf, double qpxf, double qpxfrac, i64 qpx, i64 height, i64 gh,
                                        i64 &y, i64 &oy) {
    if (!(fabs(pv) < 1.0e9)) return false;
    i64 yf;
    frac_coord_one(pv, halfhf, hf, qpxf, qpxfrac, qpx, 1, y, yf);
    oy = y - gh / 2;
    return oy + gh > 0 && oy < height;
}
__device__ __forceinline__ int mg_owner(const MgBounds &B, i64 row) {
    int g = 0;
    for (int k = 1; k < B.n; ++k) g += (B.b[k] <= row) ? 1 : 0;
    return g;
}

// Histogram of footprint-centre rows.  SKA1-Low coverage piles most visibilities onto a few hundred rows, so the counts
// are privatised per block in shared memory (rows <= MG_HIST_SMEM_ROWS) and merged once; taller grids count in global memory.
constexpr i64 MG_HIST_SMEM_ROWS = 49152;  // 192 KB of counters
__global__ void __launch_bounds__(1024) mg_row_hist_kernel(i64 count, const double *__restrict__ v, i64 height, i64 qpx, i64 gh,
                                                           uint32_t *__restrict__ hist, int privatise) {
    extern __shared__ uint32_t sh_hist[];
    const double halfhf = (double)(height / 2), hf = (double)height, qpxf = (double)qpx, qpxfrac = 0.5 / (double)qpx;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    if (privatise) {
        for (i64 r = threadIdx.x; r < height; r += blockDim.x) sh_hist[r] = 0;
        __syncthreads();
    }
    for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) {
        i64 y, oy;
        if (!mg_rows(v[k], halfhf, hf, qpxf, qpxfrac, qpx, height, gh, y, oy)) continue;
        const i64 row = min(max(y, (i64)0), height - 1);
        atomicAdd(privatise ? &sh_hist[row] : &hist[row], 1u);
    }
    if (privatise) {
        __syncthreads();
        for (i64 r = threadIdx.x; r < height; r += blockDim.x)
            if (sh_hist[r]) atomicAdd(&hist[r], sh_hist[r]);
    }
}

// One kernel for both passes of the routing: cursor == NULL counts records per destination (counts[g]); otherwise the
// records are appended to the destination-major send buffers at cursor[g]++ (warp-aggregated).
__global__ void __launch_bounds__(256) mg_route_kernel(i64 count, const double *__restrict__ u, const double *__restrict__ v,
                                                       const i64 *__restrict__ wbin, const double2 *__restrict__ vis, i64 height, i64 qpx,
                                                       i64 gh, MgBounds B, uint32_t *__restrict__ counts, uint32_t *__restrict__ cursor,
                                                       double *__restrict__ su, double *__restrict__ sv, i64 *__restrict__ swb,
                                                       double2 *__restrict__ svis, uint32_t *__restrict__ sidx) {
    const double halfhf = (double)(height / 2), hf = (double)height, qpxf = (double)qpx, qpxfrac = 0.5 / (double)qpx;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    for (i64 base = (i64)blockIdx.x * blockDim.x; base < count; base += stride) {  // warp-uniform trip count
        const i64 k = base + threadIdx.x;
        int lo = 1, hi = 0;
        if (k < count) {
            i64 y, oy;
            if (mg_rows(v[k], halfhf, hf, qpxf, qpxfrac, qpx, height, gh, y, oy)) {
                lo = mg_owner(B, max(oy, (i64)0));
                hi = mg_owner(B, min(oy + gh - 1, height - 1));
            }
        }
        for (int g = 0; g < B.n; ++g) {
            const bool mine = lo <= g && g <= hi;
            const unsigned m = __ballot_sync(0xffffffffu, mine);
            if (!m) continue;
            const int leader = __ffs(m) - 1;
            uint32_t pos = 0;
            if (lane == leader) pos = atomicAdd(cursor ? &cursor[g] : &counts[g], (uint32_t)__popc(m));
            if (!cursor) continue;
            pos = __shfl_sync(0xffffffffu, pos, leader) + (uint32_t)__popc(m & ((1u << lane) - 1u));
            if (mine) {
                su[pos] = u[k];
                sv[pos] = v[k];
                swb[pos] = wbin ? wbin[k] : 0;
                if (svis) svis[pos] = vis[k];
                if (sidx) sidx[pos] = (uint32_t)k;
            }
        }
    }
}

// out[sidx[i]] += back[i]: partial degridding sums returned by the slab owners (a footprint straddling slabs has one per owner)
__global__ void __launch_bounds__(256) mg_accumulate_kernel(i64 n, const uint32_t *__restrict__ sidx, const double2 *__restrict__ back,
                                                            double *__restrict__ out) {
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double2 b = back[i];
        const i64 k = sidx[i];
        atomicAdd(&out[2 * k], b.x);
        atomicAdd(&out[2 * k + 1], b.y);
    }
}

// ---------------------------------------------------------------------------------------------------------------
namespace {

struct Mg {
    skagrid_ctx *const *c;
    int n;
    i64 nw, qpx, gh, gw, height, width, count;
    const double *gcf;
    std::vector<double *> dtab;

    skagrid_ctx *ctx(int d) const { return c[d]; }
    int fail(int d, int rc) const {  // the caller reads messages from the first context
        if (d != 0) c[0]->err = "context " + std::to_string(d) + " (device " + std::to_string(c[d]->device) + "): " + c[d]->err;
        return rc;
    }
    // waits for every stream of every context; returns the first context whose work failed (-1: none)
    int drain(cudaError_t *err) const {
        int bad = -1;
        for (int d = 0; d < n; ++d) {
            cudaSetDevice(c[d]->device);
            cudaStream_t ss[3] = {c[d]->stream, c[d]->copy_stream, c[d]->d2h_stream};
            for (cudaStream_t st : ss) {
                const cudaError_t e = cudaStreamSynchronize(st);
                if (e != cudaSuccess && bad < 0) { bad = d; *err = e; }
            }
        }
        cudaGetLastError();
        return bad;
    }
    void shard(int d, i64 &first, i64 &cnt) const {
        const i64 base = count / n, rem = count % n;
        first = d * base + std::min<i64>(d, rem);
        cnt = base + (d < rem ? 1 : 0);
    }
    i64 even_row(int d) const { return height * d / n; }
};

// SKAGRID_MGPU_TRACE=1: drain every device at each phase boundary and print the phase's wall time to stderr (this
// serialises the phases, so the sum exceeds the untraced call; for finding out where the time goes)
void mg_trace(const Mg &m, const char *label) {
    static const bool on = getenv("SKAGRID_MGPU_TRACE") && atoi(getenv("SKAGRID_MGPU_TRACE")) != 0;
    static std::chrono::steady_clock::time_point last;
    if (!on) return;
    cudaError_t e;
    m.drain(&e);
    const auto now = std::chrono::steady_clock::now();
    if (label) fprintf(stderr, "[skagrid mgpu] %-28s %8.3f ms\n", label, std::chrono::duration<double, std::milli>(now - last).count());
    last = now;
}

int mg_init(Mg &m, skagrid_ctx *const *ctxs, int nctx, i64 nw, i64 qpx, i64 gh, i64 gw, const double *gcf, i64 height, i64 width, i64 count,
            const char *what) {
    if (!ctxs || nctx < 1 || !ctxs[0]) return SKAGRID_EINVAL;
    skagrid_ctx *c0 = ctxs[0];
    c0->err.clear();
    if (nctx > MG_MAX) return sk_fail(c0, SKAGRID_EINVAL, "%s: at most %d contexts", what, MG_MAX);
    for (int d = 0; d < nctx; ++d) {
        if (!ctxs[d]) return sk_fail(c0, SKAGRID_EINVAL, "%s: context %d is NULL", what, d);
        for (int e = 0; e < d; ++e)
            if (ctxs[e] == ctxs[d]) return sk_fail(c0, SKAGRID_EINVAL, "%s: context %d is listed twice", what, d);
    }
    if (!(nw > 0 && qpx > 0 && gh > 0 && gw > 0)) return sk_fail(c0, SKAGRID_EINVAL, "%s: non-positive kernel table dimension", what);
    if (!(height > 0 && width > 0 && height <= 65536 && width <= 65536)) return sk_fail(c0, SKAGRID_EINVAL, "%s: grid size outside [1,65536]", what);
    if (height < nctx) return sk_fail(c0, SKAGRID_EINVAL, "%s: fewer grid rows than contexts", what);
    if (count < 0) return sk_fail(c0, SKAGRID_EINVAL, "%s: negative visibility count", what);
    if (!gcf) return sk_fail(c0, SKAGRID_EINVAL, "%s: NULL kernel table", what);
    m.c = ctxs; m.n = nctx; m.nw = nw; m.qpx = qpx; m.gh = gh; m.gw = gw; m.height = height; m.width = width; m.count = count; m.gcf = gcf;
    m.dtab.assign(nctx, nullptr);
    return SKAGRID_OK;
}

// can a kernel on context a's device dereference context b's memory?
bool mg_peer(skagrid_ctx *a, skagrid_ctx *b) {
    if (a->device == b->device) return true;
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, a->device, b->device) != cudaSuccess || !can) { cudaGetLastError(); return false; }
    cudaSetDevice(a->device);
    const cudaError_t e = cudaDeviceEnablePeerAccess(b->device, 0);
    if (e == cudaSuccess) return true;
    cudaGetLastError();
    return e == cudaErrorPeerAccessAlreadyEnabled;
}

// enter every context and upload the kernel table to it
int mg_tables(Mg &m) {
    const size_t tab_bytes = (size_t)(m.nw * m.qpx * m.qpx * m.gh * m.gw) * 16;
    for (int d = 0; d < m.n; ++d) {
        int rc = sk_api_enter(m.ctx(d));
        if (!rc) rc = sk_api_up(m.ctx(d), "tab", m.gcf, tab_bytes, (void **)&m.dtab[d]);
        if (rc) return m.fail(d, rc);
    }
    // map every peer once (best effort): without it cudaMemcpyPeerAsync is staged through host memory instead of NVLink
    for (int a = 0; a < m.n; ++a)
        for (int b = 0; b < m.n; ++b)
            if (a != b) mg_peer(m.ctx(a), m.ctx(b));
    return SKAGRID_OK;
}

int mg_finish(Mg &m, int rc, int failed, const char *what, cudaEvent_t t0) {
    // the first context's stream joins every device's last event so that last_device_ms covers the whole call
    skagrid_ctx *c0 = m.ctx(0);
    if (!rc) {
        cudaSetDevice(c0->device);
        for (int d = 1; d < m.n; ++d) cudaStreamWaitEvent(c0->stream, m.ctx(d)->ev_mg[1], 0);
        cudaEventRecord(c0->ev1, c0->stream);
    }
    cudaError_t e = cudaSuccess;
    const int bad = m.drain(&e);  // always: no context may return while a peer still reads its buffers
    if (rc) return m.fail(failed, rc);
    if (bad >= 0) return m.fail(bad, sk_fail(m.ctx(bad), SKAGRID_ECUDA, "%s: %s", what, cudaGetErrorString(e)));
    for (int d = 0; d < m.n; ++d) {
        cudaSetDevice(m.ctx(d)->device);
        const int rf = sk_api_check_flags(m.ctx(d), what);
        if (rf) return m.fail(d, rf);
    }
    float ms = 0.f;
    cudaSetDevice(c0->device);
    if (cudaEventElapsedTime(&ms, t0, c0->ev1) == cudaSuccess) c0->last_ms = ms; else cudaGetLastError();
    return SKAGRID_OK;
}

#define MG_CUDA(d, call)                                                                                          \
    do {                                                                                                          \
        const cudaError_t e__ = (call);                                                                           \
        if (e__ != cudaSuccess) {                                                                                 \
            rc = sk_fail(m.ctx(d), SKAGRID_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            failed = (d);                                                                                         \
            goto done;                                                                                            \
        }                                                                                                         \
    } while (0)
#define MG_TRY(d, call)                   \
    do {                                  \
        rc = (call);                      \
        if (rc) { failed = (d); goto done; } \
    } while (0)

unsigned mg_blocks(skagrid_ctx *ctx, i64 n) { return (unsigned)std::max<i64>(1, std::min<i64>((n + 255) / 256, (i64)ctx->sm_count * 8)); }

// All-gather of the row slabs [row(d), row(d+1)) of the "grid" scratch: context d pulls every peer's slab (copy engines).
// Waits for ev_mg[wait_ev] of each peer first.
int mg_allgather(Mg &m, const std::vector<double *> &dgrid, int wait_ev, int &failed) {
    int rc = SKAGRID_OK;
    failed = 0;
    for (int d = 0; d < m.n; ++d) {
        skagrid_ctx *ctx = m.ctx(d);
        MG_CUDA(d, cudaSetDevice(ctx->device));
        for (int p = 0; p < m.n; ++p) {
            if (p == d) continue;
            const i64 r0 = m.even_row(p), r1 = m.even_row(p + 1);
            if (r1 <= r0) continue;
            MG_CUDA(d, cudaStreamWaitEvent(ctx->stream, m.ctx(p)->ev_mg[wait_ev], 0));
            MG_CUDA(d, cudaMemcpyPeerAsync(dgrid[d] + 2 * r0 * m.width, ctx->device, dgrid[p] + 2 * r0 * m.width, m.ctx(p)->device,
                                           (size_t)((r1 - r0) * m.width) * 16, ctx->stream));
        }
    }
done:
    return rc;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// visibility-sharded
// ---------------------------------------------------------------------------------------------------------------
extern "C" int skagrid_convgrid2_mgpu_vis(skagrid_ctx *const *ctxs, int nctx, int64_t nw, int64_t qpx, int64_t gh, int64_t gw, const double *gcf,
                                          int64_t height, int64_t width, double *grid, int64_t count, const double *u, const double *v,
                                          const int64_t *wbin, const double *vis) {
    Mg m;
    SK_TRY(mg_init(m, ctxs, nctx, nw, qpx, gh, gw, gcf, height, width, count, "convgrid2_mgpu_vis"));
    // u == v == wbin == NULL: every context reuses the coordinates of its share from the previous _mgpu_vis call (same count, same contexts)
    const bool coords_resident = !u && !v && !wbin;
    const int want_wbin = coords_resident ? (nw > 1 ? 1 : ctxs[0]->res_has_wbin) : (wbin != nullptr);
    if (count > 0 && !(vis && (coords_resident || (u && v && (wbin || nw == 1)))))
        return sk_fail(ctxs[0], SKAGRID_EINVAL, "convgrid2_mgpu_vis: NULL visibility array");
    SK_TRY(mg_tables(m));
    int rc = SKAGRID_OK, failed = 0;
    std::vector<double *> dgrid(nctx, nullptr);
    const size_t grid_bytes = (size_t)(height * width) * 16;
    skagrid_geom geom = {height, width, 0, height, nw, qpx, gh, gw};
    cudaSetDevice(m.ctx(0)->device);
    cudaEventRecord(m.ctx(0)->ev0, m.ctx(0)->stream);
    // phase 1: every device grids its share into a full local grid; its own slab starts from the caller's grid
    for (int d = 0; d < nctx; ++d) {
        skagrid_ctx *ctx = m.ctx(d);
        MG_CUDA(d, cudaSetDevice(ctx->device));
        ctx->resident_h = ctx->resident_w = 0;
        MG_TRY(d, sk_scratch(ctx, "grid", grid_bytes, (void **)&dgrid[d]));
        MG_CUDA(d, cudaMemsetAsync(dgrid[d], 0, grid_bytes, ctx->stream));
        const i64 r0 = m.even_row(d), r1 = m.even_row(d + 1);
        if (grid && r1 > r0)
            MG_CUDA(d, cudaMemcpyAsync(dgrid[d] + 2 * r0 * width, grid + 2 * r0 * width, (size_t)((r1 - r0) * width) * 16, cudaMemcpyHostToDevice,
                                       ctx->stream));
        i64 first, n;
        m.shard(d, first, n);
        MG_TRY(d, sk_api_stream_enqueue(ctx, &geom, m.dtab[d], dgrid[d], n, u ? u + first : nullptr, v ? v + first : nullptr,
                                        wbin ? wbin + first : nullptr, vis + 2 * first, nullptr, 0, 0.0, want_wbin));
        MG_CUDA(d, cudaEventRecord(ctx->ev_mg[0], ctx->stream));
    }
    // phase 2: reduce-scatter over peer memory -- device d sums row slab d of every peer into its own, then ships it home
    for (int d = 0; d < nctx; ++d) {
        skagrid_ctx *ctx = m.ctx(d);
        MG_CUDA(d, cudaSetDevice(ctx->device));
        const i64 r0 = m.even_row(d), r1 = m.even_row(d + 1), cells = (r1 - r0) * width;
        double2 *own = (double2 *)dgrid[d] + r0 * width;
        MgPeers direct;
        direct.n = 0;
        for (int p = 0; p < nctx && cells > 0; ++p) {
            if (p == d) continue;
            MG_CUDA(d, cudaStreamWaitEvent(ctx->stream, m.ctx(p)->ev_mg[0], 0));
            const double2 *theirs = (const double2 *)dgrid[p] + r0 * width;
            if (mg_peer(ctx, m.ctx(p))) {
                direct.p[direct.n++] = theirs;
            } else {  // no peer mapping between the two devices: stage the slab through the copy engines
                MG_CUDA(d, cudaSetDevice(ctx->device));
                double2 *stage;
                MG_TRY(d, sk_scratch(ctx, "mg_stage", (size_t)cells * 16, (void **)&stage));
                MG_CUDA(d, cudaMemcpyPeerAsync(stage, ctx->device, theirs, m.ctx(p)->device, (size_t)cells * 16, ctx->stream));
                MgPeers one;
                one.n = 1;
                one.p[0] = stage;
                mg_slab_sum_kernel<<<mg_blocks(ctx, cells), 256, 0, ctx->stream>>>(own, one, cells);
                ctx->launches++;
            }
        }
        MG_CUDA(d, cudaSetDevice(ctx->device));
        if (direct.n > 0) {
            mg_slab_sum_kernel<<<mg_blocks(ctx, cells), 256, 0, ctx->stream>>>(own, direct, cells);
            ctx->launches++;
        }
        MG_CUDA(d, cudaGetLastError());
        MG_CUDA(d, cudaEventRecord(ctx->ev_mg[1], ctx->stream));
        if (grid && cells > 0) {  // on the third stream: the slab leaves over PCIe while the all-gather below runs over NVLink
            MG_CUDA(d, cudaStreamWaitEvent(ctx->d2h_stream, ctx->ev_mg[1], 0));
            MG_CUDA(d, cudaMemcpyAsync(grid + 2 * r0 * width, own, (size_t)cells * 16, cudaMemcpyDeviceToHost, ctx->d2h_stream));
            MG_CUDA(d, cudaEventRecord(ctx->ev_d2h[0], ctx->d2h_stream));
        }
    }
    // phase 3: all-gather of the reduced slabs, so the sum is resident on every context (a following
    // skagrid_convdegrid2_mgpu_vis / skagrid_grid_to_image with grid == NULL needs no PCIe traffic)
    MG_TRY(failed, mg_allgather(m, dgrid, 1, failed));
    for (int d = 0; d < nctx; ++d) {
        MG_CUDA(d, cudaSetDevice(m.ctx(d)->device));
        if (grid) MG_CUDA(d, cudaStreamWaitEvent(m.ctx(d)->stream, m.ctx(d)->ev_d2h[0], 0));
        MG_CUDA(d, cudaEventRecord(m.ctx(d)->ev_mg[1], m.ctx(d)->stream));
    }
done:
    rc = mg_finish(m, rc, failed, "convgrid2_mgpu_vis", m.ctx(0)->ev0);
    if (!rc) for (int d = 0; d < nctx; ++d) { m.ctx(d)->resident_h = height; m.ctx(d)->resident_w = width; }
    return rc;
}

extern "C" int skagrid_convdegrid2_mgpu_vis(skagrid_ctx *const *ctxs, int nctx, int64_t nw, int64_t qpx, int64_t gh, int64_t gw,
                                            const double *gcf, int64_t height, int64_t width, const double *grid, int64_t count, const double *u,
                                            const double *v, const int64_t *wbin, double *vis_out) {
    Mg m;
    SK_TRY(mg_init(m, ctxs, nctx, nw, qpx, gh, gw, gcf, height, width, count, "convdegrid2_mgpu_vis"));
    const bool coords_resident = !u && !v && !wbin;
    const int want_wbin = coords_resident ? (nw > 1 ? 1 : ctxs[0]->res_has_wbin) : (wbin != nullptr);
    if (count > 0 && !(vis_out && (coords_resident || (u && v && (wbin || nw == 1)))))
        return sk_fail(ctxs[0], SKAGRID_EINVAL, "convdegrid2_mgpu_vis: NULL visibility array");
    if (!grid)
        for (int d = 0; d < nctx; ++d)
            if (ctxs[d]->resident_h != height || ctxs[d]->resident_w != width)
                return sk_fail(ctxs[0], SKAGRID_EINVAL, "convdegrid2_mgpu_vis: grid is NULL but context %d holds no resident %lld x %lld grid", d,
                               (long long)height, (long long)width);
    SK_TRY(mg_tables(m));
    int rc = SKAGRID_OK, failed = 0;
    std::vector<double *> dgrid(nctx, nullptr);
    const size_t grid_bytes = (size_t)(height * width) * 16;
    skagrid_geom geom = {height, width, 0, height, nw, qpx, gh, gw};
    cudaSetDevice(m.ctx(0)->device);
    cudaEventRecord(m.ctx(0)->ev0, m.ctx(0)->stream);
    // the model grid: every device uploads one row slab (P PCIe links in parallel), the rest arrives over NVLink
    for (int d = 0; d < nctx; ++d) {
        skagrid_ctx *ctx = m.ctx(d);
        MG_CUDA(d, cudaSetDevice(ctx->device));
        MG_TRY(d, sk_scratch(ctx, "grid", grid_bytes, (void **)&dgrid[d]));
        if (grid) {
            ctx->resident_h = ctx->resident_w = 0;
            const i64 r0 = m.even_row(d), r1 = m.even_row(d + 1);
            if (r1 > r0)
                MG_CUDA(d, cudaMemcpyAsync(dgrid[d] + 2 * r0 * width, grid + 2 * r0 * width, (size_t)((r1 - r0) * width) * 16,
                                           cudaMemcpyHostToDevice, ctx->stream));
        }
        MG_CUDA(d, cudaEventRecord(ctx->ev_mg[0], ctx->stream));
    }
    if (grid) MG_TRY(failed, mg_allgather(m, dgrid, 0, failed));
    for (int d = 0; d < nctx; ++d) {
        skagrid_ctx *ctx = m.ctx(d);
        MG_CUDA(d, cudaSetDevice(ctx->device));
        i64 first, n;
        m.shard(d, first, n);
        MG_TRY(d, sk_api_stream_enqueue(ctx, &geom, m.dtab[d], dgrid[d], n, u ? u + first : nullptr, v ? v + first : nullptr,
                                        wbin ? wbin + first : nullptr, nullptr, vis_out + 2 * first, 1, 0.0, want_wbin));
        MG_CUDA(d, cudaEventRecord(ctx->ev_mg[1], ctx->stream));
    }
done:
    rc = mg_finish(m, rc, failed, "convdegrid2_mgpu_vis", m.ctx(0)->ev0);
    if (!rc) for (int d = 0; d < nctx; ++d) { m.ctx(d)->resident_h = height; m.ctx(d)->resident_w = width; }
    return rc;
}

// ---------------------------------------------------------------------------------------------------------------
// uv-tile-sharded
// ---------------------------------------------------------------------------------------------------------------
namespace {

struct Route {
    MgBounds B;
    std::vector<std::vector<uint32_t>> counts;  // [src][dst]
    std::vector<std::vector<i64>> seg;          // [src][dst] start of the dst segment in src's send buffers
    std::vector<std::vector<i64>> offs;         // [src][dst] start of src's records in dst's receive buffers
    std::vector<i64> sent, recv;                // totals per src / per dst
    std::vector<i64> first, cnt;                // shard of the caller's arrays per src
    std::vector<double *> du, dv, dvis, su, sv, svis, ru, rv, rvis;
    std::vector<i64 *> dwb, swb, rwb;
    std::vector<uint32_t *> sidx;
    std::vector<double *> slab;                 // [dst] this owner's rows of the caller's grid (uploaded on the copy stream)
};

// Uploads the shards, balances the slabs, routes every record to the owners of its footprint rows.  On return the
// receive buffers ru/rv/rwb(/rvis) of every context are complete once its stream has passed ev_mg[0] of every peer
// (the function already makes every stream wait for them).
int mg_route(Mg &m, Route &R, const double *u, const double *v, const int64_t *wbin, const double *vis, bool keep_index, const double *grid,
             int &failed) {
    const int P = m.n;
    int rc = SKAGRID_OK;
    failed = 0;
    R.counts.assign(P, std::vector<uint32_t>(MG_MAX, 0));
    R.seg.assign(P, std::vector<i64>(P + 1, 0));
    R.offs.assign(P, std::vector<i64>(P, 0));
    R.sent.assign(P, 0); R.recv.assign(P, 0); R.first.assign(P, 0); R.cnt.assign(P, 0);
    R.du.assign(P, nullptr); R.dv.assign(P, nullptr); R.dvis.assign(P, nullptr); R.dwb.assign(P, nullptr);
    R.su.assign(P, nullptr); R.sv.assign(P, nullptr); R.svis.assign(P, nullptr); R.swb.assign(P, nullptr);
    R.ru.assign(P, nullptr); R.rv.assign(P, nullptr); R.rvis.assign(P, nullptr); R.rwb.assign(P, nullptr);
    R.sidx.assign(P, nullptr);
    R.slab.assign(P, nullptr);
    std::vector<uint32_t *> hist(P, nullptr);  // pinned host staging per context: [height] histogram, then [MG_MAX] counts
    std::vector<uint32_t *> dhist(P, nullptr), dcnt(P, nullptr);
    mg_trace(m, nullptr);
    // (a) shards to the devices + histogram of footprint-centre rows
    for (int d = 0; d < P; ++d) {
        skagrid_ctx *ctx = m.ctx(d);
        MG_CUDA(d, cudaSetDevice(ctx->device));
        m.shard(d, R.first[d], R.cnt[d]);
        const i64 n = R.cnt[d], f = R.first[d];
        if (n >= (i64)0xFFFFFFF0ll) { rc = sk_fail(ctx, SKAGRID_EINVAL, "mgpu_tile: more than 2^32 visibilities per device"); failed = d; goto done; }
        MG_TRY(d, sk_api_up(ctx, "mg_u", u + f, (size_t)n * 8, (void **)&R.du[d]));
        MG_TRY(d, sk_api_up(ctx, "mg_v", v + f, (size_t)n * 8, (void **)&R.dv[d]));
        if (wbin) MG_TRY(d, sk_api_up(ctx, "mg_wb", wbin + f, (size_t)n * 8, (void **)&R.dwb[d]));
        if (vis) MG_TRY(d, sk_api_up(ctx, "mg_vis", vis + 2 * f, (size_t)n * 16, (void **)&R.dvis[d]));
        MG_TRY(d, sk_scratch(ctx, "mg_hist", (size_t)m.height * 4, (void **)&dhist[d]));
        MG_TRY(d, sk_scratch(ctx, "mg_cnt", 2 * MG_MAX * 4, (void **)&dcnt[d]));
        MG_CUDA(d, cudaMemsetAsync(dhist[d], 0, (size_t)m.height * 4, ctx->stream));
        MG_CUDA(d, cudaMemsetAsync(dcnt[d], 0, 2 * MG_MAX * 4, ctx->stream));
        MG_TRY(d, sk_host_scratch(ctx, (size_t)std::max<i64>(m.height, 2 * MG_MAX) * 4, (void **)&hist[d]));
        if (n > 0) {
            const int privatise = m.height <= MG_HIST_SMEM_ROWS ? 1 : 0;
            const size_t smem = privatise ? (size_t)m.height * 4 : 0;
            if (smem > 48 * 1024) MG_CUDA(d, cudaFuncSetAttribute(mg_row_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const unsigned blocks = (unsigned)std::max<i64>(1, std::min<i64>((n + 1023) / 1024, (i64)ctx->sm_count * (privatise ? 1 : 2)));
            mg_row_hist_kernel<<<blocks, 1024, smem, ctx->stream>>>(n, R.dv[d], m.height, m.qpx, m.gh, dhist[d], privatise);
            ctx->launches++;
            MG_CUDA(d, cudaGetLastError());
        }
        MG_CUDA(d, cudaMemcpyAsync(hist[d], dhist[d], (size_t)m.height * 4, cudaMemcpyDeviceToHost, ctx->stream));  // pinned: no host stall
    }
    for (int d = 0; d < P; ++d) { MG_CUDA(d, cudaSetDevice(m.ctx(d)->device)); MG_CUDA(d, cudaStreamSynchronize(m.ctx(d)->stream)); }
    mg_trace(m, "shard H2D + row histogram");
    {   // (b) slab bounds at the k/P quantiles of the summed histogram; every slab keeps at least one row
        std::vector<i64> cum((size_t)m.height);
        i64 run = 0;
        for (i64 r = 0; r < m.height; ++r) { for (int d = 0; d < P; ++d) run += hist[d][(size_t)r]; cum[(size_t)r] = run; }
        R.B.n = P;
        R.B.b[0] = 0;
        for (int g = 1; g < P; ++g) {
            const i64 target = run * g / P;
            i64 b = (i64)(std::upper_bound(cum.begin(), cum.end(), target) - cum.begin());
            b = std::min<i64>(std::max<i64>(b, R.B.b[g - 1] + 1), m.height - (P - g));
            R.B.b[g] = b;
        }
        R.B.b[P] = m.height;
    }
    // the owners' slabs of the caller's grid travel on the copy streams while the routing below runs
    for (int g = 0; g < P; ++g) {
        skagrid_ctx *ctx = m.ctx(g);
        MG_CUDA(g, cudaSetDevice(ctx->device));
        const i64 r0 = R.B.b[g], r1 = R.B.b[g + 1];
        const size_t slab_bytes = (size_t)((r1 - r0) * m.width) * 16;
        ctx->resident_h = ctx->resident_w = 0;  // the "grid" scratch now holds a slab, not a resident full grid
        MG_TRY(g, sk_scratch(ctx, "grid", slab_bytes, (void **)&R.slab[g]));
        MG_CUDA(g, cudaMemcpyAsync(R.slab[g], grid + 2 * r0 * m.width, slab_bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
        MG_CUDA(g, cudaEventRecord(ctx->ev_copy[0], ctx->copy_stream));
    }
    // (c) records per destination
    for (int d = 0; d < P; ++d) {
        skagrid_ctx *ctx = m.ctx(d);
        MG_CUDA(d, cudaSetDevice(ctx->device));
        if (R.cnt[d] > 0) {
            mg_route_kernel<<<mg_blocks(ctx, R.cnt[d]), 256, 0, ctx->stream>>>(R.cnt[d], R.du[d], R.dv[d], R.dwb[d], nullptr, m.height, m.qpx, m.gh,
                                                                               R.B, dcnt[d], nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
            ctx->launches++;
            MG_CUDA(d, cudaGetLastError());
        }
        MG_CUDA(d, cudaMemcpyAsync(hist[d], dcnt[d], MG_MAX * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    for (int d = 0; d < P; ++d) {
        MG_CUDA(d, cudaSetDevice(m.ctx(d)->device));
        MG_CUDA(d, cudaStreamSynchronize(m.ctx(d)->stream));
        for (int g = 0; g < MG_MAX; ++g) R.counts[d][g] = hist[d][g];
    }
    mg_trace(m, "count (+ slab H2D)");
    for (int s = 0; s < P; ++s) {
        for (int g = 0; g < P; ++g) {
            R.seg[s][g + 1] = R.seg[s][g] + R.counts[s][g];
            R.offs[s][g] = R.recv[g];
            R.recv[g] += R.counts[s][g];
        }
        R.sent[s] = R.seg[s][P];
    }
    // (d) receive buffers on the owners
    for (int g = 0; g < P; ++g) {
        skagrid_ctx *ctx = m.ctx(g);
        MG_CUDA(g, cudaSetDevice(ctx->device));
        if (R.recv[g] >= (i64)0xFFFFFFF0ll) { rc = sk_fail(ctx, SKAGRID_EINVAL, "mgpu_tile: more than 2^32 routed visibilities per device"); failed = g; goto done; }
        const size_t n = (size_t)std::max<i64>(R.recv[g], 1);
        MG_TRY(g, sk_scratch(ctx, "mg_ru", n * 8, (void **)&R.ru[g]));
        MG_TRY(g, sk_scratch(ctx, "mg_rv", n * 8, (void **)&R.rv[g]));
        MG_TRY(g, sk_scratch(ctx, "mg_rwb", n * 8, (void **)&R.rwb[g]));
        if (vis) MG_TRY(g, sk_scratch(ctx, "mg_rvis", n * 16, (void **)&R.rvis[g]));
    }
    // (e) destination-major send buffers, then one peer copy per (array, destination) over NVLink
    for (int s = 0; s < P; ++s) {
        skagrid_ctx *ctx = m.ctx(s);
        MG_CUDA(s, cudaSetDevice(ctx->device));
        const size_t n = (size_t)std::max<i64>(R.sent[s], 1);
        MG_TRY(s, sk_scratch(ctx, "mg_su", n * 8, (void **)&R.su[s]));
        MG_TRY(s, sk_scratch(ctx, "mg_sv", n * 8, (void **)&R.sv[s]));
        MG_TRY(s, sk_scratch(ctx, "mg_swb", n * 8, (void **)&R.swb[s]));
        if (vis) MG_TRY(s, sk_scratch(ctx, "mg_svis", n * 16, (void **)&R.svis[s]));
        if (keep_index) MG_TRY(s, sk_scratch(ctx, "mg_sidx", n * 4, (void **)&R.sidx[s]));
        uint32_t *start = hist[s] + MG_MAX;  // pinned; the counts in [0, MG_MAX) were consumed above
        for (int g = 0; g < MG_MAX; ++g) start[g] = g < P ? (uint32_t)R.seg[s][g] : 0u;
        uint32_t *cursor = dcnt[s] + MG_MAX;
        MG_CUDA(s, cudaMemcpyAsync(cursor, start, MG_MAX * 4, cudaMemcpyHostToDevice, ctx->stream));
        if (R.cnt[s] > 0) {
            mg_route_kernel<<<mg_blocks(ctx, R.cnt[s]), 256, 0, ctx->stream>>>(R.cnt[s], R.du[s], R.dv[s], R.dwb[s], (const double2 *)R.dvis[s],
                                                                               m.height, m.qpx, m.gh, R.B, nullptr, cursor, R.su[s], R.sv[s], R.swb[s],
                                                                               (double2 *)R.svis[s], R.sidx[s]);
            ctx->launches++;
            MG_CUDA(s, cudaGetLastError());
        }
        for (int g = 0; g < P; ++g) {
            const i64 c = R.counts[s][g];
            if (c == 0) continue;
            const int dd = m.ctx(g)->device, sd = ctx->device;
            MG_CUDA(s, cudaMemcpyPeerAsync(R.ru[g] + R.offs[s][g], dd, R.su[s] + R.seg[s][g], sd, (size_t)c * 8, ctx->stream));
            MG_CUDA(s, cudaMemcpyPeerAsync(R.rv[g] + R.offs[s][g], dd, R.sv[s] + R.seg[s][g], sd, (size_t)c * 8, ctx->stream));
            MG_CUDA(s, cudaMemcpyPeerAsync(R.rwb[g] + R.offs[s][g], dd, R.swb[s] + R.seg[s][g], sd, (size_t)c * 8, ctx->stream));
            if (vis) MG_CUDA(s, cudaMemcpyPeerAsync(R.rvis[g] + 2 * R.offs[s][g], dd, R.svis[s] + 2 * R.seg[s][g], sd, (size_t)c * 16, ctx->stream));
        }
        MG_CUDA(s, cudaEventRecord(ctx->ev_mg[0], ctx->stream));
    }
    for (int g = 0; g < P; ++g) {
        MG_CUDA(g, cudaSetDevice(m.ctx(g)->device));
        for (int s = 0; s < P; ++s)
            if (s != g) MG_CUDA(g, cudaStreamWaitEvent(m.ctx(g)->stream, m.ctx(s)->ev_mg[0], 0));
        MG_CUDA(g, cudaStreamWaitEvent(m.ctx(g)->stream, m.ctx(g)->ev_copy[0], 0));
    }
    mg_trace(m, "scatter + peer copies");
done:
    return rc;
}

}  // namespace

extern "C" int skagrid_convgrid2_mgpu_tile(skagrid_ctx *const *ctxs, int nctx, int64_t nw, int64_t qpx, int64_t gh, int64_t gw, const double *gcf,
                                           int64_t height, int64_t width, double *grid, int64_t count, const double *u, const double *v,
                                           const int64_t *wbin, const double *vis, int64_t *bounds_out) {
    Mg m;
    SK_TRY(mg_init(m, ctxs, nctx, nw, qpx, gh, gw, gcf, height, width, count, "convgrid2_mgpu_tile"));
    if (!grid) return sk_fail(ctxs[0], SKAGRID_EINVAL, "convgrid2_mgpu_tile: NULL grid");
    if (count > 0 && !(u && v && vis && (wbin || nw == 1))) return sk_fail(ctxs[0], SKAGRID_EINVAL, "convgrid2_mgpu_tile: NULL visibility array");
    SK_TRY(mg_tables(m));
    int rc = SKAGRID_OK, failed = 0;
    Route R;
    cudaSetDevice(m.ctx(0)->device);
    cudaEventRecord(m.ctx(0)->ev0, m.ctx(0)->stream);
    MG_TRY(failed, mg_route(m, R, u, v, wbin, vis, false, grid, failed));
    for (int g = 0; g < nctx; ++g) {
        skagrid_ctx *ctx = m.ctx(g);
        MG_CUDA(g, cudaSetDevice(ctx->device));
        const i64 r0 = R.B.b[g], r1 = R.B.b[g + 1];
        const size_t slab_bytes = (size_t)((r1 - r0) * width) * 16;
        double *slab = R.slab[g];
        if (R.recv[g] > 0) {
            skagrid_geom geom = {height, width, r0, r1, nw, qpx, gh, gw};
            skagrid_plan *plan = nullptr;
            MG_TRY(g, sk_api_plan_acquire(ctx, &geom, R.recv[g], 0, &plan));
            MG_TRY(g, sk_plan_fill(ctx, plan, R.recv[g], R.ru[g], R.rv[g], wbin ? R.rwb[g] : nullptr, R.rvis[g], ctx->stream));
            MG_TRY(g, skagrid_dev_grid(ctx, plan, m.dtab[g], slab, 0, ctx->stream));
        }
        MG_CUDA(g, cudaMemcpyAsync(grid + 2 * r0 * width, slab, slab_bytes, cudaMemcpyDeviceToHost, ctx->stream));
        MG_CUDA(g, cudaEventRecord(ctx->ev_mg[1], ctx->stream));
    }
    mg_trace(m, "plan + grid + slab D2H");
    if (bounds_out) for (int g = 0; g <= nctx; ++g) bounds_out[g] = R.B.b[g];
done:
    return mg_finish(m, rc, failed, "convgrid2_mgpu_tile", m.ctx(0)->ev0);
}

extern "C" int skagrid_convdegrid2_mgpu_tile(skagrid_ctx *const *ctxs, int nctx, int64_t nw, int64_t qpx, int64_t gh, int64_t gw,
                                             const double *gcf, int64_t height, int64_t width, const double *grid, int64_t count, const double *u,
                                             const double *v, const int64_t *wbin, double *vis_out, int64_t *bounds_out) {
    Mg m;
    SK_TRY(mg_init(m, ctxs, nctx, nw, qpx, gh, gw, gcf, height, width, count, "convdegrid2_mgpu_tile"));
    if (!grid) return sk_fail(ctxs[0], SKAGRID_EINVAL, "convdegrid2_mgpu_tile: NULL grid");
    if (count > 0 && !(u && v && vis_out && (wbin || nw == 1))) return sk_fail(ctxs[0], SKAGRID_EINVAL, "convdegrid2_mgpu_tile: NULL visibility array");
    SK_TRY(mg_tables(m));
    int rc = SKAGRID_OK, failed = 0;
    Route R;
    std::vector<double *> partial(nctx, nullptr);
    cudaSetDevice(m.ctx(0)->device);
    cudaEventRecord(m.ctx(0)->ev0, m.ctx(0)->stream);
    MG_TRY(failed, mg_route(m, R, u, v, wbin, nullptr, true, grid, failed));
    // owners: degrid the taps on their rows of the model grid
    for (int g = 0; g < nctx; ++g) {
        skagrid_ctx *ctx = m.ctx(g);
        MG_CUDA(g, cudaSetDevice(ctx->device));
        const i64 r0 = R.B.b[g], r1 = R.B.b[g + 1];
        double *slab = R.slab[g];
        MG_TRY(g, sk_scratch(ctx, "mg_part", (size_t)std::max<i64>(R.recv[g], 1) * 16, (void **)&partial[g]));
        if (R.recv[g] > 0) {
            skagrid_geom geom = {height, width, r0, r1, nw, qpx, gh, gw};
            skagrid_plan *plan = nullptr;
            MG_TRY(g, sk_api_plan_acquire(ctx, &geom, R.recv[g], 0, &plan));
            MG_TRY(g, sk_plan_fill(ctx, plan, R.recv[g], R.ru[g], R.rv[g], wbin ? R.rwb[g] : nullptr, nullptr, ctx->stream));
            MG_TRY(g, skagrid_dev_degrid(ctx, plan, m.dtab[g], slab, partial[g], ctx->stream));
        }
        MG_CUDA(g, cudaEventRecord(ctx->ev_mg[1], ctx->stream));
    }
    mg_trace(m, "plan + degrid");
    // sources: pull the partial sums back (send order), add them per visibility, ship the share home
    for (int s = 0; s < nctx; ++s) {
        skagrid_ctx *ctx = m.ctx(s);
        MG_CUDA(s, cudaSetDevice(ctx->device));
        double *back, *out;
        MG_TRY(s, sk_scratch(ctx, "mg_back", (size_t)std::max<i64>(R.sent[s], 1) * 16, (void **)&back));
        MG_TRY(s, sk_scratch(ctx, "mg_out", (size_t)std::max<i64>(R.cnt[s], 1) * 16, (void **)&out));
        MG_CUDA(s, cudaMemsetAsync(out, 0, (size_t)std::max<i64>(R.cnt[s], 1) * 16, ctx->stream));
        for (int g = 0; g < nctx; ++g) {
            const i64 c = R.counts[s][g];
            if (c == 0) continue;
            if (g != s) MG_CUDA(s, cudaStreamWaitEvent(ctx->stream, m.ctx(g)->ev_mg[1], 0));
            MG_CUDA(s, cudaMemcpyPeerAsync(back + 2 * R.seg[s][g], ctx->device, partial[g] + 2 * R.offs[s][g], m.ctx(g)->device, (size_t)c * 16,
                                           ctx->stream));
        }
        if (R.sent[s] > 0) {
            mg_accumulate_kernel<<<mg_blocks(ctx, R.sent[s]), 256, 0, ctx->stream>>>(R.sent[s], R.sidx[s], (const double2 *)back, out);
            ctx->launches++;
            MG_CUDA(s, cudaGetLastError());
        }
        if (R.cnt[s] > 0) MG_CUDA(s, cudaMemcpyAsync(vis_out + 2 * R.first[s], out, (size_t)R.cnt[s] * 16, cudaMemcpyDeviceToHost, ctx->stream));
    }
    // ev_mg[1] of the owners was consumed above; record the final events only after every source has enqueued its waits
    for (int s = 0; s < nctx; ++s) {
        MG_CUDA(s, cudaSetDevice(m.ctx(s)->device));
        MG_CUDA(s, cudaEventRecord(m.ctx(s)->ev_mg[1], m.ctx(s)->stream));
    }
    mg_trace(m, "return + accumulate + D2H");
    if (bounds_out) for (int g = 0; g <= nctx; ++g) bounds_out[g] = R.B.b[g];
done:
    return mg_finish(m, rc, failed, "convdegrid2_mgpu_tile", m.ctx(0)->ev0);
}
