// plan.cu -- bit-exact binning (frac_coord / findClosest) and the uv-tile bucket sort that feeds the
// gridder and degridder.
//
// Reference semantics restated here:
//   frac_coord   src/Gridding.hs:126-140      frac_coords  src/Gridding.hs:142-151
//   findClosest  src/Gridding.hs:895-907      fixoutofbounds (clip, never wrap) src/Gridding.hs:883-891
// The reference has no sort: its `permute (+)` scatters V*S^2 taps unordered.  Here every visibility
// gets an integer key -- uv tile of its footprint origin (16 or 32 cells, chosen per plan), MT x MT micro-tile
// inside the tile, and (when the offset table stays small) the exact origin inside the micro-tile -- and a
// counting sort (histogram -> exclusive scan -> scatter, one 256-bit store per record) groups the 32-byte
// records per key.  One thread per tile then cuts the tile's records into work items of <= CHUNK records.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

struct BinParams {
    double halfwf, wf, halfhf, hf, qpxf, qpxfrac;
    i64 qpx, width, height, row0, row1, gh, gw, halfgh, halfgw, nw, kpitch;
    uint32_t kstride, klead;  // taps per slice of the padded table; leading all-zero slices (dense layout: 1)
    int ntx, nty, normalise, slice_override;
    int mt_shift, mtr;  // micro-tile edge = 1 << mt_shift; micro-tiles per tile row
    int tshift;         // log2 of the uv tile edge
    int kpt, cellsort;  // bucket keys per tile; cell-granular buckets
};

static BinParams make_bin_params(const Geom &g, int slice_override) {
    BinParams p;
    p.halfwf = (double)(g.width / 2);   // n `div` 2, n > 0
    p.wf = (double)g.width;
    p.halfhf = (double)(g.height / 2);
    p.hf = (double)g.height;
    p.qpxf = (double)g.qpx;
    p.qpxfrac = 0.5 / p.qpxf;
    p.qpx = g.qpx; p.width = g.width; p.height = g.height; p.row0 = g.row0; p.row1 = g.row1;
    p.gh = g.gh; p.gw = g.gw; p.halfgh = g.gh / 2; p.halfgw = g.gw / 2; p.nw = g.nw;
    p.ntx = g.ntx; p.nty = g.nty; p.normalise = g.normalise; p.slice_override = slice_override;
    p.mt_shift = g.MT == 4 ? 2 : 1; p.mtr = g.MTR; p.kpitch = g.kpitch; p.kstride = (uint32_t)(g.krows * g.kpitch); p.klead = g.dense ? 1u : 0u; p.tshift = g.tshift; p.kpt = g.kpt; p.cellsort = g.cellsort;
    return p;
}

int sk_geom_init(skagrid_ctx *ctx, const skagrid_geom *in, i64 capacity, Geom *g) {
    if (!in) return sk_fail(ctx, SKAGRID_EINVAL, "geom is NULL");
    if (in->height <= 0 || in->width <= 0 || in->qpx <= 0 || in->gh <= 0 || in->gw <= 0 || in->nw <= 0)
        return sk_fail(ctx, SKAGRID_EINVAL, "geom: non-positive dimension");
    if (in->row0 < 0 || in->row1 > in->height || in->row0 >= in->row1)
        return sk_fail(ctx, SKAGRID_EINVAL, "geom: bad owned row range [%lld,%lld) for height %lld", (i64)in->row0, (i64)in->row1, (i64)in->height);
    if (in->gh > 127 || in->gw > 127) return sk_fail(ctx, SKAGRID_EINVAL, "geom: kernel support %lldx%lld above the supported 127", (i64)in->gh, (i64)in->gw);
    g->height = in->height; g->width = in->width; g->row0 = in->row0; g->row1 = in->row1;
    g->nw = in->nw; g->qpx = in->qpx; g->gh = in->gh; g->gw = in->gw;
    // register region / micro-tile of the tiled kernels: smallest R in {16,32,48,64} with R >= S+1, then the
    // largest power-of-two micro-tile (<= 4, so that (dy,dx) fits a 16-bit one-hot) whose footprints still fit:
    // MT - 1 + S <= R
    const i64 smax = in->gh > in->gw ? in->gh : in->gw;
    g->R = smax <= 15 ? 16 : (smax <= 31 ? 32 : (smax <= 47 ? 48 : (smax <= 63 ? 64 : 0)));
    const int rr = g->R ? g->R : 64;
    g->MT = 2;
    while (g->MT < 4 && g->MT * 2 - 1 + smax <= rr) g->MT *= 2;
    // uv tile edge: 16 when the batch is dense enough (>= 128 visibilities per 16x16 tile on average over the owned area)
    // and the kernel fits the 16-wide region, else 32 (measured on B200: config 4, S=15, 1e8 visibilities on 8192^2:
    // tile 16 -> 22.5 ms, tile 32 -> 25.7 ms; config-5 shape, S=31, 5e7 on 32768^2: tile 16 -> 81.7 ms, tile 32 -> 64.9 ms)
    const double tiles16 = ((double)in->width / 16.0) * ((double)(in->row1 - in->row0) / 16.0);
    g->tile = (g->R == 16 && (double)capacity >= 128.0 * tiles16) ? 16 : 32;
    if (const char *e = getenv("SKAGRID_TILE")) { const int t = atoi(e); if (t == 16 || t == 32) g->tile = t; }  // tuning experiments
    g->tshift = g->tile == 16 ? 4 : 5;
    const i64 ntx = (in->width + in->gw - 1 + g->tile - 1) / g->tile;
    const i64 nty = ((in->row1 - in->row0) + in->gh - 1 + g->tile - 1) / g->tile;
    g->MTR = g->tile / g->MT;
    g->SG = g->tile - g->MT + rr;
    g->kpitch = g->R ? (int)((in->gw + 15) / 16 * 16) : (int)in->gw;
    g->krows = (int)in->gh;
    g->dense = 0;
    if ((g->R == 16 || g->R == 32) && (double)g->R * g->R <= 1.15 * (double)(in->gh * g->kpitch)) g->dense = 1;
    if (const char *e = getenv("SKAGRID_DENSE")) g->dense = (atoi(e) && (g->R == 16 || g->R == 32)) ? 1 : 0;  // tuning experiments
    if (g->dense) { g->kpitch = g->R; g->krows = g->R; }
    // cell-granular buckets when the offset table stays small (<= 2^27 keys, 0.5 GB); else micro-tile buckets
    g->cellsort = (ntx * nty * (i64)g->tile * g->tile <= ((i64)1 << 27)) ? 1 : 0;
    if (const char *e = getenv("SKAGRID_CELLSORT")) g->cellsort = atoi(e) ? 1 : 0;  // tuning experiments
    g->kpt = g->cellsort ? g->tile * g->tile : g->MTR * g->MTR;
    const i64 nkeys = ntx * nty * g->kpt;
    if (nkeys >= (i64)0xFFFFFFF0ll) return sk_fail(ctx, SKAGRID_EINVAL, "geom: grid too large for 32-bit bucket keys");
    if ((in->nw * in->qpx * in->qpx + 1) * g->krows * g->kpitch >= (i64)0xFFFFFFFFll) return sk_fail(ctx, SKAGRID_EINVAL, "geom: kernel table has more than 2^32 taps");
    g->ntx = (int)ntx; g->nty = (int)nty; g->nkeys = nkeys; g->normalise = 1;
    return SKAGRID_OK;
}

// Computes key / slice / loc of one visibility. Returns false when no tap can land on the owned rows.
__device__ __forceinline__ bool bin_vis(const BinParams &P, double pu, double pv, i64 wb, i64 k,
                                        uint32_t &key, uint32_t &slice, uint32_t &loc, bool &range_err) {
    range_err = false;
    if (!(fabs(pu) < 1.0e9) || !(fabs(pv) < 1.0e9)) return false;  // NaN / inf / absurd: no tap on the grid
    i64 x, xf, y, yf;
    frac_coord_one(pu, P.halfwf, P.wf, P.qpxf, P.qpxfrac, P.qpx, P.normalise, x, xf);
    frac_coord_one(pv, P.halfhf, P.hf, P.qpxf, P.qpxfrac, P.qpx, P.normalise, y, yf);
    const i64 ox = x - P.halfgw, oy = y - P.halfgh;
    if (ox + P.gw <= 0 || ox >= P.width || oy + P.gh <= P.row0 || oy >= P.row1) return false;
    if (P.slice_override) {
        slice = (uint32_t)k;
    } else {
        if (wb < 0 || wb >= P.nw || xf < 0 || xf >= P.qpx || yf < 0 || yf >= P.qpx) { range_err = true; return false; }
        slice = (uint32_t)((wb * P.qpx + yf) * P.qpx + xf);
    }
    const i64 oxs = ox + P.gw - 1, oys = oy + P.gh - 1 - P.row0;
    const int tx = (int)(oxs >> P.tshift), ty = (int)(oys >> P.tshift);
    const int lx = (int)(oxs & ((1 << P.tshift) - 1)), ly = (int)(oys & ((1 << P.tshift) - 1));
    const uint32_t mt = (uint32_t)((ly >> P.mt_shift) * P.mtr + (lx >> P.mt_shift));
    const uint32_t mtm = (1u << P.mt_shift) - 1u, dx = (uint32_t)lx & mtm, dy = (uint32_t)ly & mtm;
    loc = (0x10000u << ((dy << P.mt_shift) | dx)) | ((uint32_t)ly << 8) | (uint32_t)lx;
    // element offset of tap (-dy, -dx) of the slice, modulo 2^32: the gridder adds its per-thread tap offset
    slice = (slice + P.klead) * P.kstride - (dy * (uint32_t)P.kpitch + dx);
    // micro-tile major; inside the micro-tile optionally by exact origin (dy, dx)
    key = (uint32_t)(ty * P.ntx + tx) * (uint32_t)P.kpt + (P.cellsort ? ((mt << (2 * P.mt_shift)) | (dy << P.mt_shift) | dx) : mt);
    return true;
}

// ES: distance (in 8-byte elements) between consecutive visibilities in u / v / wbin -- 1 for the reference's structure of
// arrays, W for records of W doubles {u, v, wbin, [re, im]} as the uv-tile-sharded routing delivers them
// (skagrid_dev_plan_update_packed); the visibility itself is then two doubles at stride W instead of an aligned double2.
template <int ES>
__global__ void __launch_bounds__(256) bin_hist_kernel(BinParams P, i64 count, const double *__restrict__ u,
                                                       const double *__restrict__ v, const i64 *__restrict__ wbin,
                                                       uint32_t *__restrict__ hist, uint32_t *__restrict__ counters,
                                                       uint32_t *__restrict__ err_flag) {
    const i64 stride = (i64)gridDim.x * blockDim.x;
    uint32_t kept = 0, dropped = 0, rerr = 0;
    for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) {
        uint32_t key, slice, loc; bool re;
        if (bin_vis(P, u[k * ES], v[k * ES], wbin ? wbin[k * ES] : 0, k, key, slice, loc, re)) { atomicAdd(&hist[key], 1u); ++kept; }
        else { ++dropped; rerr |= re ? 1u : 0u; }
    }
    // block-level reduction of the statistics: one atomic per warp
    for (int o = 16; o > 0; o >>= 1) {
        kept += __shfl_xor_sync(0xffffffffu, kept, o);
        dropped += __shfl_xor_sync(0xffffffffu, dropped, o);
        rerr |= __shfl_xor_sync(0xffffffffu, rerr, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (kept) atomicAdd(&counters[2], kept);
        if (dropped) atomicAdd(&counters[3], dropped);
        if (rerr) atomicOr(err_flag, 1u);
    }
}

// Four records per thread and iteration, phase by phase (all coordinate loads, then all bucket atomics, then all
// stores): the kernel is bound by the latency of random DRAM / L2 accesses at full occupancy (ncu: every warp on the
// long scoreboard, DRAM at 28 %), so the only lever is more independent accesses in flight per thread.
constexpr int SCATTER_U = 4;
// MINB: resident blocks per SM the register allocation must allow.  ncu (profiles/r02_ncu_summary.txt): 62 registers -> 4 blocks
// -> 49 % of the warp slots, every warp on the long scoreboard: the kernel wants more random accesses in flight.
template <int ES, int MINB = 4>
__global__ void __launch_bounds__(256, MINB) bin_scatter_kernel(BinParams P, i64 count, const double *__restrict__ u,
                                                          const double *__restrict__ v, const i64 *__restrict__ wbin,
                                                          const double *__restrict__ vis, uint32_t *__restrict__ offs,
                                                          VisRec *__restrict__ rec) {
    const i64 stride = (i64)gridDim.x * blockDim.x * SCATTER_U;
    for (i64 k0 = (i64)blockIdx.x * blockDim.x * SCATTER_U + threadIdx.x; k0 < count; k0 += stride) {
        uint32_t key[SCATTER_U], slice[SCATTER_U], loc[SCATTER_U], pos[SCATTER_U];
        bool ok[SCATTER_U];
        double pu[SCATTER_U], pv[SCATTER_U];
        i64 wb[SCATTER_U];
#pragma unroll
        for (int i = 0; i < SCATTER_U; ++i) {
            const i64 k = k0 + (i64)i * blockDim.x;
            ok[i] = k < count;
            pu[i] = ok[i] ? u[k * ES] : 0.0;
            pv[i] = ok[i] ? v[k * ES] : 0.0;
            wb[i] = (ok[i] && wbin) ? wbin[k * ES] : 0;
        }
#pragma unroll
        for (int i = 0; i < SCATTER_U; ++i) {
            bool re;
            ok[i] = ok[i] && bin_vis(P, pu[i], pv[i], wb[i], k0 + (i64)i * blockDim.x, key[i], slice[i], loc[i], re);
        }
#pragma unroll
        for (int i = 0; i < SCATTER_U; ++i) pos[i] = ok[i] ? atomicAdd(&offs[key[i]], 1u) : 0u;
#pragma unroll
        for (int i = 0; i < SCATTER_U; ++i) {
            if (!ok[i]) continue;
            const i64 k = k0 + (i64)i * blockDim.x;
            double2 vv = make_double2(0.0, 0.0);
            if (vis) {
                if constexpr (ES == 1) vv = reinterpret_cast<const double2 *>(vis)[k];
                else vv = make_double2(vis[k * ES], vis[k * ES + 1]);
            }
            // one 256-bit store per record (STG.E.ENL2.256 on sm_100a): the destination is a random 32-byte slot, so this
            // halves the store instructions the LSU has to queue compared with two 128-bit stores
            const unsigned long long q0 = (unsigned long long)__double_as_longlong(vv.x), q1 = (unsigned long long)__double_as_longlong(vv.y);
            const unsigned long long q2 = (unsigned long long)slice[i] | ((unsigned long long)loc[i] << 32);
            const unsigned long long q3 = (unsigned long long)(uint32_t)k | ((unsigned long long)(key[i] / (uint32_t)P.kpt) << 32);
            asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(rec + pos[i]), "l"(q0), "l"(q1), "l"(q2), "l"(q3) : "memory");
        }
    }
}

// ------------------------------------------------------------------------------------------ scan
// Exclusive prefix sum over uint32, three passes (tile sums, scan of tile sums, apply). 4096 per block.
constexpr int SCAN_T = 256, SCAN_PER = 16, SCAN_TILE = SCAN_T * SCAN_PER;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t val, uint32_t *warp_sums, uint32_t &total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = val;
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = lane < (SCAN_T / 32) ? warp_sums[lane] : 0;
        uint32_t winc = w;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < (SCAN_T / 32)) warp_sums[lane] = winc - w;  // exclusive warp offsets
        if (lane == 31) warp_sums[32] = winc;                   // block total
    }
    __syncthreads();
    total = warp_sums[32];
    return warp_sums[wid] + inc - val;
}

__global__ void __launch_bounds__(SCAN_T) scan_tile_sums(const uint32_t *__restrict__ data, i64 n, uint32_t *__restrict__ sums) {
    __shared__ uint32_t ws[33];
    const i64 base = (i64)blockIdx.x * SCAN_TILE + (i64)threadIdx.x * SCAN_PER;
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_PER; ++i) if (base + i < n) s += data[base + i];
    uint32_t total;
    block_exclusive_scan(s, ws, total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_T) scan_sums(uint32_t *sums, i64 nb) {
    __shared__ uint32_t ws[33];
    uint32_t carry = 0;
    for (i64 base = 0; base < nb; base += SCAN_T) {
        const i64 i = base + threadIdx.x;
        uint32_t v = i < nb ? sums[i] : 0, total;
        uint32_t ex = block_exclusive_scan(v, ws, total);
        if (i < nb) sums[i] = carry + ex;
        carry += total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SCAN_T) scan_apply(uint32_t *__restrict__ data, i64 n, const uint32_t *__restrict__ sums) {
    __shared__ uint32_t ws[33];
    const i64 base = (i64)blockIdx.x * SCAN_TILE + (i64)threadIdx.x * SCAN_PER;
    uint32_t v[SCAN_PER], s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_PER; ++i) { v[i] = (base + i < n) ? data[base + i] : 0; s += v[i]; }
    uint32_t total;
    uint32_t ex = block_exclusive_scan(s, ws, total) + sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_PER; ++i) { if (base + i < n) data[base + i] = ex; ex += v[i]; }
}

// ------------------------------------------------------------------------------------------ work items
// After the scatter offs[k] is the END of bucket k. One thread per uv tile cuts the tile's records into
// runs of at most CHUNK so that a dense tile is shared by many blocks.
__global__ void __launch_bounds__(256) make_items_kernel(const uint32_t *__restrict__ offs, int ntiles, int mt_per_tile, WorkItem *__restrict__ items,
                                                         uint32_t *__restrict__ counters, uint32_t max_items) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    const uint32_t begin = t == 0 ? 0u : offs[(i64)t * mt_per_tile - 1];
    const uint32_t end = offs[(i64)t * mt_per_tile + mt_per_tile - 1];
    if (end == begin) return;
    atomicAdd(&counters[4], 1u);
    const uint32_t n = (end - begin + CHUNK - 1) / CHUNK;
    const uint32_t base = atomicAdd(&counters[0], n);
    for (uint32_t c = 0; c < n && base + c < max_items; ++c) {
        WorkItem it;
        it.tile = (uint32_t)t;
        it.begin = begin + c * CHUNK;
        it.end = min(end, it.begin + CHUNK);
        it.pad = 0;
        items[base + c] = it;
    }
}

// ------------------------------------------------------------------------------------------ plan API
void sk_plan_free(skagrid_plan *p) {
    if (!p) return;
    if (p->d_offs) cudaFree(p->d_offs);
    if (p->d_rec) cudaFree(p->d_rec);
    if (p->d_items) cudaFree(p->d_items);
    if (p->d_counters) cudaFree(p->d_counters);
    if (p->d_blocksums) cudaFree(p->d_blocksums);
    if (p->d_table) cudaFree(p->d_table);
    delete p;
}

static int plan_fill(skagrid_ctx *ctx, skagrid_plan *p, i64 count, const double *u, const double *v, const i64 *wbin, const double *vis,
                     int es, cudaStream_t st);
int sk_plan_fill(skagrid_ctx *ctx, skagrid_plan *p, i64 count, const double *u, const double *v,
                     const i64 *wbin, const double *vis, cudaStream_t st) {
    return plan_fill(ctx, p, count, u, v, wbin, vis, 1, st);
}

static int plan_fill(skagrid_ctx *ctx, skagrid_plan *p, i64 count, const double *u, const double *v, const i64 *wbin, const double *vis,
                     int es, cudaStream_t st) {
    if (count < 0 || count > p->capacity) return sk_fail(ctx, SKAGRID_EINVAL, "plan: count %lld exceeds capacity %lld", count, p->capacity);
    if (p->slice_override && (count + 1) * p->g.krows * p->g.kpitch >= (i64)0xFFFFFFFFll) return sk_fail(ctx, SKAGRID_EINVAL, "plan: per-visibility kernel table has more than 2^32 taps");
    if (count > 0 && (!u || !v)) return sk_fail(ctx, SKAGRID_EINVAL, "plan: u/v is NULL");
    p->count = count;
    p->has_vis = vis != nullptr;
    const Geom &g = p->g;
    const BinParams P = make_bin_params(g, p->slice_override);
    SK_CUDA(ctx, cudaMemsetAsync(p->d_offs, 0, (size_t)(g.nkeys + 1) * sizeof(uint32_t), st));
    SK_CUDA(ctx, cudaMemsetAsync(p->d_counters, 0, 16 * sizeof(uint32_t), st));
    const int blocks = ctx->sm_count * 8;
    if (count > 0) {
        if (es == 1) bin_hist_kernel<1><<<blocks, 256, 0, st>>>(P, count, u, v, wbin, p->d_offs, p->d_counters, ctx->d_flags);
        else if (es == 3) bin_hist_kernel<3><<<blocks, 256, 0, st>>>(P, count, u, v, wbin, p->d_offs, p->d_counters, ctx->d_flags);
        else bin_hist_kernel<5><<<blocks, 256, 0, st>>>(P, count, u, v, wbin, p->d_offs, p->d_counters, ctx->d_flags);
        SK_LAUNCH_CHECK(ctx);
    }
    const i64 n = g.nkeys + 1;
    const i64 nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    scan_tile_sums<<<(unsigned)nb, SCAN_T, 0, st>>>(p->d_offs, n, p->d_blocksums);
    SK_LAUNCH_CHECK(ctx);
    scan_sums<<<1, SCAN_T, 0, st>>>(p->d_blocksums, nb);
    SK_LAUNCH_CHECK(ctx);
    scan_apply<<<(unsigned)nb, SCAN_T, 0, st>>>(p->d_offs, n, p->d_blocksums);
    SK_LAUNCH_CHECK(ctx);
    if (count > 0) {
        static const int occ = getenv("SKAGRID_SCATTER_OCC") ? atoi(getenv("SKAGRID_SCATTER_OCC")) : 0;  // tuning experiments
        if (es == 1 && occ == 6) bin_scatter_kernel<1, 6><<<ctx->sm_count * 6, 256, 0, st>>>(P, count, u, v, wbin, vis, p->d_offs, p->d_rec);
        else if (es == 1 && occ == 8) bin_scatter_kernel<1, 8><<<ctx->sm_count * 8, 256, 0, st>>>(P, count, u, v, wbin, vis, p->d_offs, p->d_rec);
        else if (es == 1) bin_scatter_kernel<1><<<blocks, 256, 0, st>>>(P, count, u, v, wbin, vis, p->d_offs, p->d_rec);
        else if (es == 3) bin_scatter_kernel<3><<<blocks, 256, 0, st>>>(P, count, u, v, wbin, vis, p->d_offs, p->d_rec);
        else bin_scatter_kernel<5><<<blocks, 256, 0, st>>>(P, count, u, v, wbin, vis, p->d_offs, p->d_rec);
        SK_LAUNCH_CHECK(ctx);
    }
    const int ntiles = g.ntx * g.nty;
    make_items_kernel<<<(ntiles + 255) / 256, 256, 0, st>>>(p->d_offs, ntiles, g.kpt, p->d_items, p->d_counters, (uint32_t)p->max_items);
    SK_LAUNCH_CHECK(ctx);
    return SKAGRID_OK;
}

int sk_plan_alloc(skagrid_ctx *ctx, const skagrid_geom *geom, i64 capacity, int slice_override, skagrid_plan **out) {
    *out = nullptr;
    if (capacity < 0 || capacity >= (i64)0xFFFFFFF0ll) return sk_fail(ctx, SKAGRID_EINVAL, "plan: count %lld out of range", capacity);
    skagrid_plan *p = new skagrid_plan();
    memset(p, 0, sizeof *p);
    int rc = sk_geom_init(ctx, geom, capacity, &p->g);
    if (rc) { delete p; return rc; }
    p->capacity = capacity > 0 ? capacity : 1;
    p->slice_override = slice_override;
    const Geom &g = p->g;
    const i64 ntiles = (i64)g.ntx * g.nty;
    p->max_items = p->capacity / CHUNK + ntiles + 1;
    p->nblocksums = (g.nkeys + 1 + SCAN_TILE - 1) / SCAN_TILE;
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = cudaMalloc(&p->d_offs, (size_t)(g.nkeys + 1) * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&p->d_rec, (size_t)p->capacity * sizeof(VisRec));
    if (e == cudaSuccess) e = cudaMalloc(&p->d_items, (size_t)p->max_items * sizeof(WorkItem));
    if (e == cudaSuccess) e = cudaMalloc(&p->d_counters, 16 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&p->d_blocksums, (size_t)(p->nblocksums + 1) * sizeof(uint32_t));
    if (e != cudaSuccess) {
        cudaGetLastError();
        sk_plan_free(p);
        return sk_fail(ctx, SKAGRID_ENOMEM, "plan: device allocation failed for %lld visibilities, %lld buckets: %s", capacity, g.nkeys, cudaGetErrorString(e));
    }
    *out = p;
    return SKAGRID_OK;
}

// Reads and clears the context's device error word (bit 0: a w-plane / oversampling / antenna index out of
// range, bit 1: a visibility outside the weight grid).  Synchronises `st`.
int sk_take_flags(skagrid_ctx *ctx, cudaStream_t st, uint32_t *flags) {
    SK_CUDA(ctx, cudaMemcpyAsync(flags, ctx->d_flags, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    SK_CUDA(ctx, cudaMemsetAsync(ctx->d_flags, 0, sizeof(uint32_t), st));
    SK_CUDA(ctx, cudaStreamSynchronize(st));
    return SKAGRID_OK;
}

extern "C" int skagrid_dev_plan_create(skagrid_ctx *ctx, const skagrid_geom *geom, int64_t count, const double *u,
                                       const double *v, const int64_t *wbin, const double *vis, int slice_override,
                                       void *stream, skagrid_plan **out) {
    if (!ctx || !out) return SKAGRID_EINVAL;
    *out = nullptr;
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    skagrid_plan *p = nullptr;
    SK_TRY(sk_plan_alloc(ctx, geom, count, slice_override, &p));
    cudaStream_t st = sk_stream(ctx, stream);
    int rc = sk_plan_fill(ctx, p, count, u, v, (const i64 *)wbin, vis, st);
    uint32_t flag = 0;
    if (!rc) rc = sk_take_flags(ctx, st, &flag);
    if (rc) { sk_plan_free(p); return rc; }
    if (flag & 1u) { sk_plan_free(p); return sk_fail(ctx, SKAGRID_ERANGE, "plan: a w-plane index is outside [0,%lld)", (i64)geom->nw); }
    *out = p;
    return SKAGRID_OK;
}

// An empty plan with room for `capacity` visibilities (filled later by skagrid_dev_plan_update / _update_packed).
extern "C" int skagrid_dev_plan_alloc(skagrid_ctx *ctx, const skagrid_geom *geom, int64_t capacity, int slice_override, skagrid_plan **out) {
    if (!ctx || !out) return SKAGRID_EINVAL;
    *out = nullptr;
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    return sk_plan_alloc(ctx, geom, capacity, slice_override, out);
}

extern "C" void skagrid_dev_plan_destroy(skagrid_ctx *ctx, skagrid_plan *plan) {
    if (ctx) cudaSetDevice(ctx->device);
    sk_plan_free(plan);
}

extern "C" int skagrid_dev_plan_update(skagrid_ctx *ctx, skagrid_plan *plan, int64_t count, const double *u,
                                       const double *v, const int64_t *wbin, const double *vis, void *stream) {
    if (!ctx || !plan) return SKAGRID_EINVAL;
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    return sk_plan_fill(ctx, plan, count, u, v, (const i64 *)wbin, vis, sk_stream(ctx, stream));
}

// Records of `width` doubles {u, v, wbin (int64 bits), [re, im]}: width 5 with visibilities, 3 for degrid-only plans.
extern "C" int skagrid_dev_plan_update_packed(skagrid_ctx *ctx, skagrid_plan *plan, int64_t count, const double *d_rec, int width, void *stream) {
    if (!ctx || !plan) return SKAGRID_EINVAL;
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (width != 3 && width != 5) return sk_fail(ctx, SKAGRID_EINVAL, "plan_update_packed: record width must be 3 or 5 doubles");
    if (count > 0 && !d_rec) return sk_fail(ctx, SKAGRID_EINVAL, "plan_update_packed: NULL records");
    return plan_fill(ctx, plan, count, d_rec, d_rec + 1, reinterpret_cast<const i64 *>(d_rec + 2), width == 5 ? d_rec + 3 : nullptr, width,
                     sk_stream(ctx, stream));
}

// New visibility VALUES for the coordinates the plan was built from: rec[r].vis = vis[rec[r].index] (the caller's order).
// An imaging major cycle grids residuals of the SAME uvw again and again; with this the binning and the bucket sort are paid
// once per data set, not once per cycle (sequential pass over the records, one random 16-byte read per visibility).
// SORTED: vis is already in plan order (vis[r] belongs to record r: the caller permuted its data once with plan_order), so the
// pass is purely sequential; otherwise one random 16-byte read per record.
template <bool SORTED>
__global__ void __launch_bounds__(256) plan_set_vis_kernel(const uint32_t *__restrict__ counters, VisRec *rec, const double2 *__restrict__ vis) {
    const i64 kept = (i64)counters[2];
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x; r < kept; r += stride) {
        const i64 idx = SORTED ? r : (i64) reinterpret_cast<const uint4 *>(rec + r)[1].z;
        *reinterpret_cast<double2 *>(rec + r) = vis[idx];
    }
}

// index[r] = position in the caller's arrays of the visibility behind record r (r < kept, plan_stats[0])
__global__ void __launch_bounds__(256) plan_order_kernel(const uint32_t *__restrict__ counters, const VisRec *__restrict__ rec, uint32_t *__restrict__ index) {
    const i64 kept = (i64)counters[2];
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x; r < kept; r += stride) index[r] = reinterpret_cast<const uint4 *>(rec + r)[1].z;
}

extern "C" int skagrid_dev_plan_set_vis(skagrid_ctx *ctx, skagrid_plan *plan, const double *d_vis, int in_plan_order, void *stream) {
    if (!ctx || !plan) return SKAGRID_EINVAL;
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (plan->count == 0) return SKAGRID_OK;
    if (!d_vis) return sk_fail(ctx, SKAGRID_EINVAL, "plan_set_vis: NULL visibilities");
    if (in_plan_order)
        plan_set_vis_kernel<true><<<ctx->sm_count * 8, 256, 0, sk_stream(ctx, stream)>>>(plan->d_counters, plan->d_rec, reinterpret_cast<const double2 *>(d_vis));
    else
        plan_set_vis_kernel<false><<<ctx->sm_count * 8, 256, 0, sk_stream(ctx, stream)>>>(plan->d_counters, plan->d_rec, reinterpret_cast<const double2 *>(d_vis));
    SK_LAUNCH_CHECK(ctx);
    plan->has_vis = 1;
    return SKAGRID_OK;
}

extern "C" int skagrid_dev_plan_order(skagrid_ctx *ctx, skagrid_plan *plan, uint32_t *d_index, void *stream) {
    if (!ctx || !plan) return SKAGRID_EINVAL;
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (plan->count == 0) return SKAGRID_OK;
    if (!d_index) return sk_fail(ctx, SKAGRID_EINVAL, "plan_order: NULL index array");
    plan_order_kernel<<<ctx->sm_count * 8, 256, 0, sk_stream(ctx, stream)>>>(plan->d_counters, plan->d_rec, d_index);
    SK_LAUNCH_CHECK(ctx);
    return SKAGRID_OK;
}

extern "C" int skagrid_dev_plan_stats(skagrid_ctx *ctx, skagrid_plan *plan, void *stream, int64_t stats[5]) {
    if (!ctx || !plan || !stats) return SKAGRID_EINVAL;
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    uint32_t c[16];
    cudaStream_t st = sk_stream(ctx, stream);
    SK_CUDA(ctx, cudaMemcpyAsync(c, plan->d_counters, sizeof c, cudaMemcpyDeviceToHost, st));
    SK_CUDA(ctx, cudaStreamSynchronize(st));
    stats[0] = c[2]; stats[1] = c[3]; stats[2] = c[0]; stats[3] = (i64)plan->g.ntx * plan->g.nty; stats[4] = c[4];
    return SKAGRID_OK;
}

// ------------------------------------------------------------------------------------------ public binning kernels
__global__ void __launch_bounds__(256) frac_coord_kernel(i64 n, i64 qpx, i64 count, const double *__restrict__ p, i64 *__restrict__ fl,
                                                         i64 *__restrict__ fr, int normalise) {
    const double halfnf = (double)(n / 2), nf = (double)n, qpxf = (double)qpx, qpxfrac = 0.5 / (double)qpx;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) {
        i64 a, b;
        frac_coord_one(p[k], halfnf, nf, qpxf, qpxfrac, qpx, normalise, a, b);
        fl[k] = a; fr[k] = b;
    }
}

int sk_frac_coord_dev(skagrid_ctx *ctx, i64 n, i64 qpx, i64 count, const double *p, i64 *fl, i64 *frac, int normalise, cudaStream_t st) {
    if (count <= 0) return SKAGRID_OK;
    const int blocks = (int)std::min<i64>((count + 255) / 256, (i64)ctx->sm_count * 16);
    frac_coord_kernel<<<blocks, 256, 0, st>>>(n, qpx, count, p, fl, frac, normalise);
    SK_LAUNCH_CHECK(ctx);
    return SKAGRID_OK;
}

// findClosest, src/Gridding.hs:895-907, with the Q4 clamp (w above the last plane -> nw-1).
__device__ __forceinline__ i64 find_closest_one(i64 len, const double *__restrict__ ws, double w) {
    i64 mn = 0, mx = len;
    while ((mx - mn) / 2 >= 1) {  // operands non-negative: C division == Haskell div
        const i64 id = (mx + mn) / 2;
        if (w > ws[id]) mn = id; else mx = id;
    }
    if (mx >= len) return mn;
    return (fabs(w - ws[mn]) < fabs(w - ws[mx])) ? mn : mx;
}

__global__ void __launch_bounds__(256) find_closest_kernel(i64 len, const double *__restrict__ ws, i64 count, const double *__restrict__ w,
                                                           i64 *__restrict__ out) {
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) out[k] = find_closest_one(len, ws, w[k]);
}

int sk_find_closest_dev(skagrid_ctx *ctx, i64 nw, const double *wbins, i64 count, const double *w, i64 *out, cudaStream_t st) {
    if (count <= 0) return SKAGRID_OK;
    if (nw <= 0) return sk_fail(ctx, SKAGRID_EINVAL, "find_closest: empty wbins");
    const int blocks = (int)std::min<i64>((count + 255) / 256, (i64)ctx->sm_count * 16);
    find_closest_kernel<<<blocks, 256, 0, st>>>(nw, wbins, count, w, out);
    SK_LAUNCH_CHECK(ctx);
    return SKAGRID_OK;
}
