// ipc.cu -- peer-memory exchange for the one-process-per-GPU multi-GPU modes: the data path over NVLink 5 / NVSwitch without
// a collective library.  Every rank allocates its exchange buffers here (cudaMalloc), exports a CUDA IPC handle, the ranks
// swap the 64-byte handles once (any transport: torch.distributed in the Python layer) and open each other's buffers; from
// then on the kernels below dereference peer memory directly and the copy engines move slabs between devices.
//
//   peer_sum       own[i] += sum over peers of peer[i]   -- the reduce-scatter of the visibility-sharded mode (gridding is
//                  linear in the visibilities, permute (+), src/Gridding.hs:377): rank r sums row slab r of every peer's grid;
//                  with `broadcast` the same kernel writes the sum back into every peer's grid (all-reduce in one pass)
//   peer_barrier   stream-ordered barrier between the ranks: a flag per (rank, peer) in peer memory, written with
//                  st.release.sys, polled with ld.acquire.sys (bounded), so no host thread and no collective is involved
//   peer_copy(2d)  cudaMemcpyAsync / cudaMemcpy2DAsync between an opened peer buffer and a local one (copy engines over
//                  NVLink): all-gather of reduced slabs, routed records, returned partial sums, the image transpose
#include <algorithm>
#include <cstring>

#include "common.cuh"

constexpr int IPC_MAX = 64;

struct IpcPeers {
    double2 *p[IPC_MAX];
    int n;
};
struct IpcFlags {
    uint32_t *p[IPC_MAX];  // p[k]: rank k's flag array (IPC_MAX words), p[me] the local one
    int n, me;
};

// BCAST: the sum is also written back into every peer's copy -- reduce-scatter and all-gather in ONE kernel: an element is
// read from and then written to each peer by the same thread, and no other rank touches this rank's slab of anybody's grid,
// so no ordering beyond the barriers around the kernel is needed; loads and stores travel in opposite NVLink directions.
template <bool BCAST>
__global__ void __launch_bounds__(256) ipc_sum_kernel(double2 *own, IpcPeers peers, i64 n) {
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double2 v[IPC_MAX / 8];
        double2 a = own[i];
        // peer loads first (independent, all in flight over NVLink), then the sum
        for (int k0 = 0; k0 < peers.n; k0 += IPC_MAX / 8) {
#pragma unroll
            for (int k = 0; k < IPC_MAX / 8; ++k)
                if (k0 + k < peers.n) v[k] = __ldcv(peers.p[k0 + k] + i);
#pragma unroll
            for (int k = 0; k < IPC_MAX / 8; ++k)
                if (k0 + k < peers.n) { a.x += v[k].x; a.y += v[k].y; }
        }
        own[i] = a;
        if (BCAST)
            for (int k = 0; k < peers.n; ++k) peers.p[k][i] = a;
    }
}

// One block, one thread per peer.  Thread k tells rank k "rank `me` has reached epoch e" (release: everything this rank's
// stream did before is visible system-wide first) and waits until rank k has said the same to us (acquire).
__global__ void ipc_barrier_kernel(IpcFlags F, uint32_t epoch, uint32_t *err_flag) {
    const int k = threadIdx.x;
    if (k >= F.n || k == F.me) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(F.p[k] + F.me), "r"(epoch) : "memory");
    const uint32_t *mine = F.p[F.me] + k;
    uint32_t seen = 0;
    for (long long spin = 0; spin < (1ll << 26); ++spin) {  // ~30 s
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
        if ((int32_t)(seen - epoch) >= 0) return;
        __nanosleep(64);
    }
    atomicOr(err_flag, 4u);  // a peer never arrived (bit 2 of the context's error word): do not hang the device forever
}

// dst[s][i] = src[s][i] for every segment s: the SMs pull from peer memory (8-byte elements: routed records are 24 or 40 bytes
// long, so segments are only 8-byte aligned).  Alternative to the copy-engine pulls (peer_copy); which one is faster is a
// measurement (profiles/r02_config5_substages.md).
struct IpcSegs {
    const double *src[IPC_MAX];
    double *dst[IPC_MAX];
    long long n[IPC_MAX];
    int count;
};
__global__ void __launch_bounds__(256) ipc_gather_kernel(IpcSegs S, long long nmax) {
    // element i of EVERY segment per trip: one load per peer in flight per thread (a kernel that walks the segments one after
    // the other reads from one peer at a time: 217 GB/s per rank on 8 GPUs against 635 GB/s for the all-peers-at-once
    // pattern of ipc_sum_kernel, profiles/r02_peer_primitives_n8.json)
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < nmax; i += stride) {
        for (int k0 = 0; k0 < S.count; k0 += 8) {
            double v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k0 + k < S.count && i < S.n[k0 + k]) v[k] = __ldcv(S.src[k0 + k] + i);
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k0 + k < S.count && i < S.n[k0 + k]) S.dst[k0 + k][i] = v[k];
        }
    }
}

// Strided form, 16-byte elements: dst[s][r * dpitch + c] = src[s][r * spitch + c] for r < rows, c < width (all in complex128
// units).  The transpose of the slab-distributed grid -> image: every rank pulls its column block of every peer's rows.
struct IpcSegs2d {
    const double2 *src[IPC_MAX];
    double2 *dst[IPC_MAX];
    long long rows[IPC_MAX], spitch[IPC_MAX];
    long long width, dpitch;
    int count;
};
__global__ void __launch_bounds__(256) ipc_gather2d_kernel(IpcSegs2d S, long long rmax) {
    const i64 stride = (i64)gridDim.x * blockDim.x;
    const i64 nmax = rmax * S.width;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < nmax; i += stride) {  // element (r, c) of every segment per trip
        const i64 r = i / S.width, c = i - r * S.width;
        for (int k0 = 0; k0 < S.count; k0 += 8) {
            double2 v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k0 + k < S.count && r < S.rows[k0 + k]) v[k] = __ldcv(S.src[k0 + k] + r * S.spitch[k0 + k] + c);
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k0 + k < S.count && r < S.rows[k0 + k]) S.dst[k0 + k][r * S.dpitch + c] = v[k];
        }
    }
}

extern "C" int skagrid_ipc_alloc(skagrid_ctx *ctx, int64_t bytes, void **d_ptr, unsigned char handle[64]) {
    SK_TRY(sk_api_enter(ctx));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    if (!d_ptr || !handle || bytes <= 0) return sk_fail(ctx, SKAGRID_EINVAL, "ipc_alloc: bad argument");
    *d_ptr = nullptr;
    void *p = nullptr;
    if (cudaMalloc(&p, (size_t)bytes) != cudaSuccess) {
        cudaGetLastError();
        return sk_fail(ctx, SKAGRID_ENOMEM, "ipc_alloc: %lld bytes", (long long)bytes);
    }
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return sk_fail(ctx, SKAGRID_ECUDA, "ipc_alloc: cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    }
    SK_CUDA(ctx, cudaMemset(p, 0, (size_t)bytes));
    memcpy(handle, &h, 64);
    *d_ptr = p;
    return SKAGRID_OK;
}

extern "C" int skagrid_ipc_free(skagrid_ctx *ctx, void *d_ptr) {
    SK_TRY(sk_api_enter(ctx));
    if (d_ptr) SK_CUDA(ctx, cudaFree(d_ptr));
    return SKAGRID_OK;
}

extern "C" int skagrid_ipc_open(skagrid_ctx *ctx, const unsigned char handle[64], void **d_ptr) {
    SK_TRY(sk_api_enter(ctx));
    if (!d_ptr || !handle) return sk_fail(ctx, SKAGRID_EINVAL, "ipc_open: NULL argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    *d_ptr = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return sk_fail(ctx, SKAGRID_ECUDA, "ipc_open: cudaIpcOpenMemHandle: %s (peer access between the two devices is required)", cudaGetErrorString(e));
    }
    return SKAGRID_OK;
}

extern "C" int skagrid_ipc_close(skagrid_ctx *ctx, void *d_ptr) {
    SK_TRY(sk_api_enter(ctx));
    if (d_ptr) SK_CUDA(ctx, cudaIpcCloseMemHandle(d_ptr));
    return SKAGRID_OK;
}

extern "C" int skagrid_dev_peer_sum(skagrid_ctx *ctx, int npeers, double *const *d_peers, double *d_own, int64_t ncomplex, int broadcast, void *stream) {
    SK_TRY(sk_api_enter(ctx));
    if (npeers < 0 || npeers > IPC_MAX) return sk_fail(ctx, SKAGRID_EINVAL, "dev_peer_sum: 0..%d peers", IPC_MAX);
    if (ncomplex <= 0 || npeers == 0) return SKAGRID_OK;
    if (!d_peers || !d_own) return sk_fail(ctx, SKAGRID_EINVAL, "dev_peer_sum: NULL pointer");
    IpcPeers P;
    P.n = npeers;
    for (int k = 0; k < npeers; ++k) {
        if (!d_peers[k]) return sk_fail(ctx, SKAGRID_EINVAL, "dev_peer_sum: peer %d is NULL", k);
        P.p[k] = reinterpret_cast<double2 *>(d_peers[k]);
    }
    const unsigned blocks = (unsigned)std::max<i64>(1, std::min<i64>((ncomplex + 255) / 256, (i64)ctx->sm_count * 8));
    if (broadcast) ipc_sum_kernel<true><<<blocks, 256, 0, sk_stream(ctx, stream)>>>(reinterpret_cast<double2 *>(d_own), P, ncomplex);
    else ipc_sum_kernel<false><<<blocks, 256, 0, sk_stream(ctx, stream)>>>(reinterpret_cast<double2 *>(d_own), P, ncomplex);
    SK_LAUNCH_CHECK(ctx);
    return SKAGRID_OK;
}

extern "C" int skagrid_dev_peer_gather(skagrid_ctx *ctx, int nseg, void *const *d_dst, const void *const *d_src, const int64_t *bytes, int max_blocks,
                                       void *stream) {
    SK_TRY(sk_api_enter(ctx));
    if (nseg < 0 || nseg > IPC_MAX) return sk_fail(ctx, SKAGRID_EINVAL, "dev_peer_gather: 0..%d segments", IPC_MAX);
    if (nseg == 0) return SKAGRID_OK;
    if (!d_dst || !d_src || !bytes) return sk_fail(ctx, SKAGRID_EINVAL, "dev_peer_gather: NULL argument");
    IpcSegs S;
    S.count = 0;
    i64 total = 0, nmax = 0;
    for (int k = 0; k < nseg; ++k) {
        if (bytes[k] <= 0) continue;
        if (!d_dst[k] || !d_src[k]) return sk_fail(ctx, SKAGRID_EINVAL, "dev_peer_gather: segment %d is NULL", k);
        if ((bytes[k] & 7) || ((uintptr_t)d_dst[k] & 7) || ((uintptr_t)d_src[k] & 7)) return sk_fail(ctx, SKAGRID_EINVAL, "dev_peer_gather: segment %d is not 8-byte aligned", k);
        S.src[S.count] = static_cast<const double *>(d_src[k]);
        S.dst[S.count] = static_cast<double *>(d_dst[k]);
        S.n[S.count] = bytes[k] / 8;
        total += bytes[k] / 8;
        nmax = std::max<i64>(nmax, bytes[k] / 8);
        ++S.count;
    }
    if (S.count == 0) return SKAGRID_OK;
    // max_blocks > 0 caps the grid (e.g. one block per SM) so that kernels of another stream can run beside the exchange
    const unsigned blocks = (unsigned)std::max<i64>(1, std::min<i64>((nmax + 255) / 256, max_blocks > 0 ? (i64)max_blocks : (i64)ctx->sm_count * 8));
    (void)total;
    ipc_gather_kernel<<<blocks, 256, 0, sk_stream(ctx, stream)>>>(S, nmax);
    SK_LAUNCH_CHECK(ctx);
    return SKAGRID_OK;
}

// All segments share the destination pitch and the width (one column block); pitches and width in BYTES, multiples of 16.
extern "C" int skagrid_dev_peer_gather2d(skagrid_ctx *ctx, int nseg, void *const *d_dst, int64_t dpitch, const void *const *d_src,
                                         const int64_t *spitch, int64_t width_bytes, const int64_t *rows, int max_blocks, void *stream) {
    SK_TRY(sk_api_enter(ctx));
    if (nseg < 0 || nseg > IPC_MAX) return sk_fail(ctx, SKAGRID_EINVAL, "dev_peer_gather2d: 0..%d segments", IPC_MAX);
    if (nseg == 0 || width_bytes <= 0) return SKAGRID_OK;
    if (!d_dst || !d_src || !spitch || !rows) return sk_fail(ctx, SKAGRID_EINVAL, "dev_peer_gather2d: NULL argument");
    if ((width_bytes & 15) || (dpitch & 15)) return sk_fail(ctx, SKAGRID_EINVAL, "dev_peer_gather2d: width and pitches must be multiples of 16 bytes");
    IpcSegs2d S;
    S.count = 0; S.width = width_bytes / 16; S.dpitch = dpitch / 16;
    i64 total = 0, rmax = 0;
    for (int k = 0; k < nseg; ++k) {
        if (rows[k] <= 0) continue;
        if (!d_dst[k] || !d_src[k] || (spitch[k] & 15) || ((uintptr_t)d_dst[k] & 15) || ((uintptr_t)d_src[k] & 15))
            return sk_fail(ctx, SKAGRID_EINVAL, "dev_peer_gather2d: segment %d is NULL or not 16-byte aligned", k);
        S.src[S.count] = static_cast<const double2 *>(d_src[k]);
        S.dst[S.count] = static_cast<double2 *>(d_dst[k]);
        S.rows[S.count] = rows[k];
        S.spitch[S.count] = spitch[k] / 16;
        total += rows[k] * S.width;
        rmax = std::max<i64>(rmax, rows[k]);
        ++S.count;
    }
    if (S.count == 0) return SKAGRID_OK;
    const unsigned blocks = (unsigned)std::max<i64>(1, std::min<i64>((rmax * S.width + 255) / 256, max_blocks > 0 ? (i64)max_blocks : (i64)ctx->sm_count * 8));
    (void)total;
    ipc_gather2d_kernel<<<blocks, 256, 0, sk_stream(ctx, stream)>>>(S, rmax);
    SK_LAUNCH_CHECK(ctx);
    return SKAGRID_OK;
}

extern "C" int skagrid_dev_peer_barrier(skagrid_ctx *ctx, int nranks, int rank, uint32_t *const *d_flags, uint32_t epoch, void *stream) {
    SK_TRY(sk_api_enter(ctx));
    if (nranks < 1 || nranks > IPC_MAX || rank < 0 || rank >= nranks) return sk_fail(ctx, SKAGRID_EINVAL, "dev_peer_barrier: bad rank / world");
    if (nranks == 1) return SKAGRID_OK;
    if (!d_flags) return sk_fail(ctx, SKAGRID_EINVAL, "dev_peer_barrier: NULL flags");
    IpcFlags F;
    F.n = nranks; F.me = rank;
    for (int k = 0; k < nranks; ++k) {
        if (!d_flags[k]) return sk_fail(ctx, SKAGRID_EINVAL, "dev_peer_barrier: flags of rank %d are NULL", k);
        F.p[k] = d_flags[k];
    }
    ipc_barrier_kernel<<<1, IPC_MAX, 0, sk_stream(ctx, stream)>>>(F, epoch, ctx->d_flags);
    SK_LAUNCH_CHECK(ctx);
    return SKAGRID_OK;
}

extern "C" int skagrid_dev_peer_copy(skagrid_ctx *ctx, void *d_dst, const void *d_src, int64_t bytes, void *stream) {
    SK_TRY(sk_api_enter(ctx));
    if (bytes <= 0) return SKAGRID_OK;
    if (!d_dst || !d_src) return sk_fail(ctx, SKAGRID_EINVAL, "dev_peer_copy: NULL pointer");
    SK_CUDA(ctx, cudaMemcpyAsync(d_dst, d_src, (size_t)bytes, cudaMemcpyDefault, sk_stream(ctx, stream)));
    return SKAGRID_OK;
}

extern "C" int skagrid_dev_peer_copy2d(skagrid_ctx *ctx, void *d_dst, int64_t dpitch, const void *d_src, int64_t spitch, int64_t width_bytes,
                                       int64_t rows, void *stream) {
    SK_TRY(sk_api_enter(ctx));
    if (width_bytes <= 0 || rows <= 0) return SKAGRID_OK;
    if (!d_dst || !d_src) return sk_fail(ctx, SKAGRID_EINVAL, "dev_peer_copy2d: NULL pointer");
    SK_CUDA(ctx, cudaMemcpy2DAsync(d_dst, (size_t)dpitch, d_src, (size_t)spitch, (size_t)width_bytes, (size_t)rows, cudaMemcpyDefault,
                                   sk_stream(ctx, stream)));
    return SKAGRID_OK;
}
