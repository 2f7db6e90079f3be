// gridder.cu -- the convolutional gridder and its adjoint, hand-written for sm_100a.
//
// Reference semantics (src/Gridding.hs:153-197 convgrid, :199-244 convgrid2, :318-377 convgrid4):
//     grid[y - gh/2 + i, x - gw/2 + j] += vis * table[slice][i, j]        for all taps (i, j)
// with (x, xf, y, yf) = frac_coords, out-of-grid taps dropped (fixoutofbounds, :883-891).  The reference
// lowers this to ONE unordered atomic scatter of V*gh*gw complex128 updates (`permute (+)`, :377).
//
// B200 design (not a translation):
//  * the plan (plan.cu) has already grouped the visibilities per uv tile and per MT x MT micro-tile;
//  * a persistent thread block pulls work items (runs of <= CHUNK records of one tile) from a queue and
//    owns a (TILE-MT+R)^2 complex-double subgrid in shared memory for the duration of the item;
//  * inside the block every subgrid cell is owned by exactly one thread for the whole item, by residue:
//    cell (cy, cx) belongs to the thread with (cy mod R, cx mod R) in its residue set.  A footprint of
//    S <= R-MT+1 taps per dimension covers each residue at most once, so a thread has at most one tap per
//    residue per visibility, accumulates it in REGISTERS while the micro-tile stays the same, and folds
//    the registers into its own shared-memory cells when it changes -- no atomics, no barriers, no bank
//    conflicts in the hot loop;
//  * the kernel taps are 128-bit loads; the 16 threads of a row read 16 consecutive taps of the slice;
//  * at the end of the item the subgrid is added to the grid in HBM/L2 with fp64 RED (no return value).
//    Dense tiles are split over many blocks for load balance, which is why the flush is a reduction and
//    not a plain store; its share of the runtime is small (one flush per <= 4096 visibilities).
//
// Variant 1 is the literal `permute (+)`: one thread per (visibility, tap) with a global atomic; it is the
// cross-check of the tiled kernel and the path for kernel shapes the tiled kernel does not cover.
#include "common.cuh"

struct GridArgs {
    const VisRec *rec;
    const WorkItem *items;
    uint32_t *counters;
    const double2 *table;
    double2 *grid;    // points at the first owned row
    double2 *vis_out; // degridder output
    int gh, gw, s2;
    int mt_mask;      // ~(MT - 1)
    int SG;           // subgrid edge (= pitch)
    int ntx;
    int width, nrows; // grid width, owned rows
    int queue;        // index into counters of this launch's queue head
    int prefetch;     // issue L1 prefetches of upcoming kernel slices
};

__device__ __forceinline__ double2 ldg2(const double2 *p) {
    return __ldg(p);
}

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

__device__ __forceinline__ void red_add(double *addr, double v) {
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
}

template <int R>
__global__ void __launch_bounds__(GRID_THREADS, (R == 16 ? 3 : (R == 32 ? 2 : 1))) grid_tiled_kernel(const GridArgs A) {
    constexpr int C = R / 16;  // residues per thread per dimension
    extern __shared__ double2 sg[];
    __shared__ uint32_t s_item;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15, warp = tid >> 5, lane = tid & 31;
    const int ncell = A.SG * A.SG;
    const uint32_t n_items = A.counters[0];
    const int slice_lines = (A.s2 * 16 + 127) / 128;
    constexpr int PD = 12;  // prefetch distance in records

    for (;;) {
        if (tid == 0) s_item = atomicAdd(&A.counters[A.queue], 1u);
        __syncthreads();
        const uint32_t item = s_item;
        if (item >= n_items) break;
        for (int c = tid; c < ncell; c += GRID_THREADS) sg[c] = make_double2(0.0, 0.0);
        const WorkItem it = A.items[item];
        __syncthreads();

        double2 acc[C][C];
        int roty[C], rotx[C];
#pragma unroll
        for (int a = 0; a < C; ++a) {
            roty[a] = 0; rotx[a] = 0;
#pragma unroll
            for (int b = 0; b < C; ++b) acc[a][b] = make_double2(0.0, 0.0);
        }
        int cur_mx = -1, cur_my = -1;

        // software pipeline: the record of iteration r+1 is loaded while r is processed
        const VisRec *rp = A.rec + it.begin;
        double2 vis = ldg2(reinterpret_cast<const double2 *>(rp));
        uint4 meta = __ldg(reinterpret_cast<const uint4 *>(rp) + 1);
        uint32_t pf_slice = 0xFFFFFFFFu;

        for (uint32_t r = it.begin; r < it.end; ++r) {
            const uint32_t rn = (r + 1 < it.end) ? r + 1 : r;
            const double2 vis_n = ldg2(reinterpret_cast<const double2 *>(A.rec + rn));
            const uint4 meta_n = __ldg(reinterpret_cast<const uint4 *>(A.rec + rn) + 1);

            if (A.prefetch && ((r & 7u) == (uint32_t)warp)) {
                // this warp prefetches the kernel slice of a record PD ahead; the slice id it uses was
                // loaded one turn (8 records) earlier so the prefetch never waits on that load
                if (pf_slice != 0xFFFFFFFFu && lane < slice_lines)
                    prefetch_l1(reinterpret_cast<const char *>(A.table + (size_t)pf_slice * A.s2) + lane * 128);
                const uint32_t rpf = r + PD + 8;
                pf_slice = rpf < it.end ? __ldg(&A.rec[rpf].slice) : 0xFFFFFFFFu;
            }

            const int lx = (int)(meta.y & 255u), ly = (int)(meta.y >> 8);
            const int mx = lx & A.mt_mask, my = ly & A.mt_mask;
            if (mx != cur_mx || my != cur_my) {  // warp-uniform: all threads walk the same records
                if (cur_mx >= 0) {
#pragma unroll
                    for (int a = 0; a < C; ++a)
#pragma unroll
                        for (int b = 0; b < C; ++b) {
                            if (acc[a][b].x != 0.0 || acc[a][b].y != 0.0) {
                                double2 *cell = sg + (cur_my + roty[a]) * A.SG + cur_mx + rotx[b];
                                double2 t = *cell;
                                t.x += acc[a][b].x; t.y += acc[a][b].y;
                                *cell = t;
                                acc[a][b] = make_double2(0.0, 0.0);
                            }
                        }
                }
                cur_mx = mx; cur_my = my;
#pragma unroll
                for (int a = 0; a < C; ++a) {
                    roty[a] = (ty + 16 * a - my) & (R - 1);
                    rotx[a] = (tx + 16 * a - mx) & (R - 1);
                }
            }
            const int dy = ly - my, dx = lx - mx;
            const double2 *kp = A.table + (size_t)meta.x * A.s2;
            double2 k[C][C];
            bool ok[C][C];
#pragma unroll
            for (int a = 0; a < C; ++a) {
                const int i = roty[a] - dy;
#pragma unroll
                for (int b = 0; b < C; ++b) {
                    const int j = rotx[b] - dx;
                    ok[a][b] = (unsigned)i < (unsigned)A.gh && (unsigned)j < (unsigned)A.gw;
                    k[a][b] = make_double2(0.0, 0.0);
                    if (ok[a][b]) k[a][b] = ldg2(kp + i * A.gw + j);
                }
            }
#pragma unroll
            for (int a = 0; a < C; ++a)
#pragma unroll
                for (int b = 0; b < C; ++b) {
                    // (vr + i vi)(kr + i ki); a zero tap (invalid cell) adds +0
                    acc[a][b].x = fma(vis.x, k[a][b].x, acc[a][b].x);
                    acc[a][b].x = fma(-vis.y, k[a][b].y, acc[a][b].x);
                    acc[a][b].y = fma(vis.x, k[a][b].y, acc[a][b].y);
                    acc[a][b].y = fma(vis.y, k[a][b].x, acc[a][b].y);
                }
            vis = vis_n; meta = meta_n;
        }
        if (cur_mx >= 0) {
#pragma unroll
            for (int a = 0; a < C; ++a)
#pragma unroll
                for (int b = 0; b < C; ++b) {
                    if (acc[a][b].x != 0.0 || acc[a][b].y != 0.0) {
                        double2 *cell = sg + (cur_my + roty[a]) * A.SG + cur_mx + rotx[b];
                        double2 t = *cell;
                        t.x += acc[a][b].x; t.y += acc[a][b].y;
                        *cell = t;
                    }
                }
        }
        __syncthreads();

        // subgrid -> grid.  Cells outside the owned rows / the grid are dropped (fixoutofbounds).
        const int tyi = (int)(it.tile / (uint32_t)A.ntx), txi = (int)(it.tile % (uint32_t)A.ntx);
        const int gx0 = txi * TILE - (A.gw - 1), gy0 = tyi * TILE - (A.gh - 1);
        for (int c = tid; c < ncell; c += GRID_THREADS) {
            const double2 v = sg[c];
            if (v.x == 0.0 && v.y == 0.0) continue;
            const int cy = c / A.SG, cx = c - cy * A.SG;
            const int gx = gx0 + cx, gy = gy0 + cy;
            if ((unsigned)gx < (unsigned)A.width && (unsigned)gy < (unsigned)A.nrows) {
                double *g = reinterpret_cast<double *>(A.grid + (size_t)gy * A.width + gx);
                red_add(g, v.x);
                red_add(g + 1, v.y);
            }
        }
        // the next iteration's first barrier orders these reads before the subgrid is zeroed again
    }
}

// Variant 1: the literal unordered scatter.  One thread per (record, tap).
__global__ void __launch_bounds__(256) grid_atomic_kernel(const GridArgs A) {
    const i64 total = (i64)A.counters[2] * A.s2;  // counters[2] = records kept by the plan
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const i64 r = idx / A.s2;
        const int t = (int)(idx - r * A.s2);
        const int i = t / A.gw, j = t - i * A.gw;
        const double2 vis = ldg2(reinterpret_cast<const double2 *>(A.rec + r));
        const uint4 meta = __ldg(reinterpret_cast<const uint4 *>(A.rec + r) + 1);
        const int lx = (int)(meta.y & 255u), ly = (int)(meta.y >> 8);
        const int tyi = (int)(meta.w / (uint32_t)A.ntx), txi = (int)(meta.w % (uint32_t)A.ntx);
        const int gx = txi * TILE + lx - (A.gw - 1) + j, gy = tyi * TILE + ly - (A.gh - 1) + i;
        if ((unsigned)gx >= (unsigned)A.width || (unsigned)gy >= (unsigned)A.nrows) continue;
        const double2 k = ldg2(A.table + (size_t)meta.x * A.s2 + t);
        double *g = reinterpret_cast<double *>(A.grid + (size_t)gy * A.width + gx);
        red_add(g, vis.x * k.x - vis.y * k.y);
        red_add(g + 1, vis.x * k.y + vis.y * k.x);
    }
}

// Degridder: the adjoint gather, one warp per visibility, lanes over taps, shuffle reduction.
//   vis_out[index] = sum_{i,j} conj(table[slice][i,j]) * grid[y0 + i, x0 + j]
// Records are in tile order, so the warps of a block read the same few KB of the grid (L1/L2 hits) and
// the slice is one contiguous, coalesced stream.
__global__ void __launch_bounds__(256) degrid_warp_kernel(const GridArgs A) {
    const int lane = threadIdx.x & 31;
    const i64 warp0 = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const i64 nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
    const i64 count = (i64)A.counters[2];
    for (i64 r = warp0; r < count; r += nwarps) {
        const uint4 meta = __ldg(reinterpret_cast<const uint4 *>(A.rec + r) + 1);
        const int lx = (int)(meta.y & 255u), ly = (int)(meta.y >> 8);
        const int tyi = (int)(meta.w / (uint32_t)A.ntx), txi = (int)(meta.w % (uint32_t)A.ntx);
        const int gx0 = txi * TILE + lx - (A.gw - 1), gy0 = tyi * TILE + ly - (A.gh - 1);
        const double2 *kp = A.table + (size_t)meta.x * A.s2;
        double ar = 0.0, ai = 0.0;
        int i = lane / A.gw, j = lane - i * A.gw;
        const int di = 32 / A.gw, dj = 32 - di * A.gw;
        for (int t = lane; t < A.s2; t += 32) {
            const int gx = gx0 + j, gy = gy0 + i;
            if ((unsigned)gx < (unsigned)A.width && (unsigned)gy < (unsigned)A.nrows) {
                const double2 k = ldg2(kp + t);
                const double2 g = ldg2(A.grid + (size_t)gy * A.width + gx);
                // conj(k) * g
                ar = fma(k.x, g.x, ar); ar = fma(k.y, g.y, ar);
                ai = fma(k.x, g.y, ai); ai = fma(-k.y, g.x, ai);
            }
            i += di; j += dj;
            if (j >= A.gw) { j -= A.gw; ++i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ar += __shfl_xor_sync(0xffffffffu, ar, o);
            ai += __shfl_xor_sync(0xffffffffu, ai, o);
        }
        if (lane == 0) A.vis_out[meta.z] = make_double2(ar, ai);
    }
}

static GridArgs make_args(skagrid_plan *plan, const double *table, double *grid) {
    const Geom &g = plan->g;
    GridArgs A;
    A.rec = plan->d_rec; A.items = plan->d_items; A.counters = plan->d_counters;
    A.table = reinterpret_cast<const double2 *>(table);
    A.grid = reinterpret_cast<double2 *>(grid);
    A.vis_out = nullptr;
    A.gh = (int)g.gh; A.gw = (int)g.gw; A.s2 = (int)(g.gh * g.gw);
    A.mt_mask = ~(g.MT - 1);
    A.SG = g.SG; A.ntx = g.ntx;
    A.width = (int)g.width; A.nrows = (int)(g.row1 - g.row0);
    A.queue = 1; A.prefetch = 0;
    return A;
}

template <int R>
static int launch_tiled(skagrid_ctx *ctx, const GridArgs &A, cudaStream_t st) {
    const size_t smem = (size_t)A.SG * A.SG * sizeof(double2);
    static bool configured = false;
    if (!configured) {
        SK_CUDA(ctx, cudaFuncSetAttribute(grid_tiled_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = true;
    }
    int per_sm = 0;
    SK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, grid_tiled_kernel<R>, GRID_THREADS, smem));
    if (per_sm < 1) return sk_fail(ctx, SKAGRID_ECUDA, "tiled gridder does not fit on an SM (smem %zu)", smem);
    grid_tiled_kernel<R><<<ctx->sm_count * per_sm, GRID_THREADS, smem, st>>>(A);
    SK_LAUNCH_CHECK(ctx);
    return SKAGRID_OK;
}

extern "C" int skagrid_dev_grid(skagrid_ctx *ctx, skagrid_plan *plan, const double *table, double *grid, int variant, void *stream) {
    if (!ctx || !plan || !table || !grid) return SKAGRID_EINVAL;
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = sk_stream(ctx, stream);
    if (plan->count == 0) return SKAGRID_OK;
    if (!plan->has_vis) return sk_fail(ctx, SKAGRID_EINVAL, "grid: the plan was built without visibilities (degrid-only)");
    GridArgs A = make_args(plan, table, grid);
    const int R = plan->g.R;
    if (variant == 1 || R == 0) {
        grid_atomic_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(A);
        SK_LAUNCH_CHECK(ctx);
        return SKAGRID_OK;
    }
    A.prefetch = (variant == 2) ? 0 : 1;
    SK_CUDA(ctx, cudaMemsetAsync(plan->d_counters + 1, 0, sizeof(uint32_t), st));
    if (R == 16) return launch_tiled<16>(ctx, A, st);
    if (R == 32) return launch_tiled<32>(ctx, A, st);
    return launch_tiled<64>(ctx, A, st);
}

extern "C" int skagrid_dev_degrid(skagrid_ctx *ctx, skagrid_plan *plan, const double *table, const double *grid, double *vis_out,
                                  void *stream) {
    if (!ctx || !plan || !table || !grid || !vis_out) return SKAGRID_EINVAL;
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = sk_stream(ctx, stream);
    if (plan->count == 0) return SKAGRID_OK;
    // visibilities without a tap on the owned rows are not in the record list: their output is 0
    SK_CUDA(ctx, cudaMemsetAsync(vis_out, 0, (size_t)plan->count * sizeof(double2), st));
    GridArgs A = make_args(plan, table, const_cast<double *>(grid));
    A.vis_out = reinterpret_cast<double2 *>(vis_out);
    degrid_warp_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(A);
    SK_LAUNCH_CHECK(ctx);
    return SKAGRID_OK;
}
