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
//    owns a (tile-MT+R)^2 complex-double subgrid in shared memory for the duration of the item (tile = 16 or
//    32 footprint origins per dimension, chosen per plan);
//  * inside the block every subgrid cell is owned by exactly one thread for the whole item, by residue:
//    cell (cy, cx) belongs to the thread with (cy mod R, cx mod R) in its residue set.  A footprint of
//    S <= R-MT+1 taps per dimension covers each residue at most once, so a thread has at most one tap per
//    residue per visibility, accumulates it in REGISTERS while the micro-tile stays the same, and folds
//    the registers into its own shared-memory cells when it changes -- no atomics, no barriers, no bank
//    conflicts in the hot loop;
//  * the kernel taps are 128-bit loads from a padded copy of the table (rows on 256-byte boundaries); the 16
//    threads of a row read 16 consecutive taps of the slice; records are staged through shared memory with
//    cp.async and DEPTH tap slots rotate per residue, so DEPTH x residues loads per thread are in flight before
//    the first FMA needs one;
//  * at the end of the item the subgrid is added to the grid in HBM/L2 with fp64 RED (no return value).
//    Dense tiles are split over many blocks for load balance, which is why the flush is a reduction and
//    not a plain store; its share of the runtime is small (one flush per <= 4096 visibilities).
//
// Variant 1 is the literal `permute (+)`: one thread per (visibility, tap) with a global atomic; it is the
// cross-check of the tiled kernel and the path for kernel shapes the tiled kernel does not cover.
#include <cstdlib>

#include "common.cuh"

struct GridArgs {
    const VisRec *rec;
    const WorkItem *items;
    uint32_t *counters;
    const double2 *table;
    double2 *grid;    // points at the first owned row
    double2 *vis_out; // degridder output
    int gh, gw, s2;
    int kpitch;       // row pitch (taps) of the padded kernel table; a slice is gh * kpitch taps
    int mt_mask;      // ~(MT - 1)
    int tile, tshift; // uv tile edge (footprint origins) and its log2
    int SG;           // subgrid edge (= pitch)
    int ntx;
    int width, nrows; // grid width, owned rows
    int queue;        // index into counters of this launch's queue head
    int out_by_record;  // degridder: write result r at vis_out[r] (plan order, sequential full sectors) instead of at the caller's index
};

__device__ __forceinline__ double2 ldg2(const double2 *p) { return __ldg(p); }

// Kernel-tap load.  The taps stream out of the L2-resident table with ~1-3 % L1 hits, and ncu shows the dense gridder at 96 %
// of the L1TEX pipe (profiles/r02_ncu_summary.txt): LD selects how the miss is handled --
//   0  ld.global.nc (LDG.CONSTANT): allocates the line in L1 (r01 behaviour)
//   1  ld.global.cg: cached in L2 only
//   2  ld.global.nc.L1::no_allocate: read-only path, no L1 allocation
template <int LD>
__device__ __forceinline__ double2 ld_tap(const double2 *p) {
    if constexpr (LD == 1) {
        return __ldcg(p);
    } else if constexpr (LD == 2) {
        double2 v;
        asm("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
        return v;
    } else {
        return __ldg(p);
    }
}
static int tap_load_mode() {
    static const int m = getenv("SKAGRID_TAP_LOAD") ? atoi(getenv("SKAGRID_TAP_LOAD")) : 0;
    return (m == 1 || m == 2) ? m : 0;
}

__device__ __forceinline__ void red_add(double *addr, double v) {
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
}

// Per-thread state of one micro-tile: where the thread's residues sit inside the R x R region.
template <int CY, int CX>
struct MtState {
    uint32_t key;                  // loc & mtkey_mask of the micro-tile (0xFFFFFFFF: none)
    int toff[CY][CX];              // tap offset roty*kpitch + rotx of the thread's (a,b) residue
    uint32_t vmask[CY][CX];        // bit 16 + dy*MT+dx set: residue (a,b) has a valid tap for a footprint at (dy,dx)
    int cell[CY][CX];              // shared-memory cell of residue (a,b)
};

// (x mod R) for x in (-R, 2R); R = 48 is the one region edge that is not a power of two
template <int R>
__device__ __forceinline__ int mod_region(int x) {
    if constexpr ((R & (R - 1)) == 0) return x & (R - 1);
    else { x += x < 0 ? R : 0; return x >= R ? x - R : x; }
}

template <int R, int CY, int CX, int MT>
__device__ __forceinline__ void mt_setup(MtState<CY, CX> &S, uint32_t key, int ty, int tx, const GridArgs &A) {
    constexpr int TY = R / CY;  // thread rows of the block
    S.key = key;
    const int mx = (int)(key & 255u), my = (int)((key >> 8) & 255u);
    int roty[CY], rotx[CX];
    uint32_t ym[CY], xm[CX];
#pragma unroll
    for (int a = 0; a < CY; ++a) {
        roty[a] = mod_region<R>(ty + TY * a - my);
        ym[a] = 0;
#pragma unroll
        for (int d = 0; d < MT; ++d)
            if ((unsigned)(roty[a] - d) < (unsigned)A.gh) ym[a] |= 1u << d;
    }
#pragma unroll
    for (int b = 0; b < CX; ++b) {
        rotx[b] = mod_region<R>(tx + 16 * b - mx);
        xm[b] = 0;
#pragma unroll
        for (int d = 0; d < MT; ++d)
            if ((unsigned)(rotx[b] - d) < (unsigned)A.gw) xm[b] |= 1u << d;
    }
#pragma unroll
    for (int a = 0; a < CY; ++a)
#pragma unroll
        for (int b = 0; b < CX; ++b) {
            S.toff[a][b] = roty[a] * A.kpitch + rotx[b];
            S.cell[a][b] = (my + roty[a]) * A.SG + mx + rotx[b];
            uint32_t m = 0;
#pragma unroll
            for (int d = 0; d < MT; ++d)
                if ((ym[a] >> d) & 1u) m |= xm[b] << (d * MT);
            S.vmask[a][b] = m << 16;  // lines up with the one-hot (dy*MT+dx) field of VisRec::loc
        }
}

// records staged in shared memory per batch: 48 for the 16-wide region (keeps 8 blocks of 17.6 KB resident per SM),
// 128 for the wider ones (their subgrid, not the staging buffers, limits residency)
template <int R> struct RecBatch { static constexpr int value = R == 16 ? 48 : 128; };

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// TY thread rows x 16 thread columns; a thread owns CY x CX = (R/TY) x (R/16) residues.  Fewer, fatter threads
// (TY = 8 for R = 16) halve the per-visibility bookkeeping (record decode, broadcast shared-memory reads, loop
// control), which matters because the kernel is bound by the L1/shared-memory data pipe and the issue slots.
template <int R, int MT, int DEPTH, int TY>
__global__ void __launch_bounds__(16 * TY, (R == 16 ? (TY == 8 ? 8 : 6) : (R == 32 ? 2 : 1))) grid_tiled_kernel(const GridArgs A) {
    constexpr int CY = R / TY, CX = R / 16;  // residues per thread
    constexpr int NT = 16 * TY;              // threads per block
    constexpr int REC_BATCH = RecBatch<R>::value;
    static_assert(REC_BATCH <= NT, "one staging thread per record");
    extern __shared__ double2 sg[];
    __shared__ __align__(16) uint4 s_rec[2][REC_BATCH * 2];  // double-buffered record batches (32 B each)
    __shared__ uint32_t s_item;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int ncell = A.SG * A.SG;
    const uint32_t n_items = A.counters[0];
    constexpr uint32_t mtkey_mask = (uint32_t)(~(MT - 1) & 255) * 0x0101u;
    const uint4 *recq = reinterpret_cast<const uint4 *>(A.rec);

    for (;;) {
        if (tid == 0) s_item = atomicAdd(&A.counters[A.queue], 1u);
        __syncthreads();
        const uint32_t item = s_item;
        if (item >= n_items) break;
        const WorkItem it = A.items[item];
        const uint32_t nrec = it.end - it.begin;
        const uint32_t nbatch = (nrec + REC_BATCH - 1) / REC_BATCH;
        // stage batch 0 (asynchronously) while the subgrid is zeroed
        if (tid < REC_BATCH && (uint32_t)tid < nrec) {
            cp_async16(&s_rec[0][2 * tid], recq + 2 * (size_t)(it.begin + tid));
            cp_async16(&s_rec[0][2 * tid + 1], recq + 2 * (size_t)(it.begin + tid) + 1);
        }
        cp_async_commit();
        for (int c = tid; c < ncell; c += NT) sg[c] = make_double2(0.0, 0.0);

        double2 acc[CY][CX];
        int cell_cur[CY][CX];
        MtState<CY, CX> S;  // micro-tile state of the record whose taps were requested last
        S.key = 0xFFFFFFFFu;
#pragma unroll
        for (int a = 0; a < CY; ++a)
#pragma unroll
            for (int b = 0; b < CX; ++b) { acc[a][b] = make_double2(0.0, 0.0); cell_cur[a][b] = 0; S.cell[a][b] = 0; S.toff[a][b] = 0; S.vmask[a][b] = 0; }

        // A tap slot: the taps of one staged record in flight, where its products go and whether it opens a new
        // micro-tile (then the registers are folded before it is consumed).
        struct Slot { double2 k[CY][CX]; int cell[CY][CX]; bool sw; };
        auto issue = [&](const uint4 *buf, uint32_t j, Slot &s) {
            const uint2 m = *reinterpret_cast<const uint2 *>(&buf[2 * j + 1]);  // kbase, loc (broadcast read)
            const uint32_t key = m.y & mtkey_mask;
            s.sw = key != S.key;
            if (s.sw) {  // warp-uniform: all threads walk the same records
                mt_setup<R, CY, CX, MT>(S, key, ty, tx, A);
#pragma unroll
                for (int a = 0; a < CY; ++a)
#pragma unroll
                    for (int b = 0; b < CX; ++b) s.cell[a][b] = S.cell[a][b];  // only read when s.sw
            }
#pragma unroll
            for (int a = 0; a < CY; ++a)
#pragma unroll
                for (int b = 0; b < CX; ++b) {
                    s.k[a][b] = make_double2(0.0, 0.0);  // a residue without a tap adds +0 (cheaper than predicating the FMAs)
                    if (S.vmask[a][b] & m.y) s.k[a][b] = ldg2(A.table + (uint32_t)(m.x + (uint32_t)S.toff[a][b]));
                }
        };
        // acc += vis_j * k
        auto consume = [&](const uint4 *buf, uint32_t j, const Slot &s) {
            const double2 vis = *reinterpret_cast<const double2 *>(&buf[2 * j]);
            if (s.sw) {  // fold the register accumulators into the thread's own subgrid cells and retarget them
#pragma unroll
                for (int a = 0; a < CY; ++a)
#pragma unroll
                    for (int b = 0; b < CX; ++b) {
                        if (acc[a][b].x != 0.0 || acc[a][b].y != 0.0) {
                            double2 t = sg[cell_cur[a][b]];
                            t.x += acc[a][b].x; t.y += acc[a][b].y;
                            sg[cell_cur[a][b]] = t;
                            acc[a][b] = make_double2(0.0, 0.0);
                        }
                        cell_cur[a][b] = s.cell[a][b];
                    }
            }
#pragma unroll
            for (int a = 0; a < CY; ++a)
#pragma unroll
                for (int b = 0; b < CX; ++b) {  // (vr + i vi)(kr + i ki)
                    acc[a][b].x = fma(vis.x, s.k[a][b].x, acc[a][b].x);
                    acc[a][b].x = fma(-vis.y, s.k[a][b].y, acc[a][b].x);
                    acc[a][b].y = fma(vis.x, s.k[a][b].y, acc[a][b].y);
                    acc[a][b].y = fma(vis.y, s.k[a][b].x, acc[a][b].y);
                }
        };

        for (uint32_t bi = 0; bi < nbatch; ++bi) {
            const uint32_t base = bi * REC_BATCH;
            const uint32_t m = min((uint32_t)REC_BATCH, nrec - base);
            // stage batch bi+1 into the other buffer (its previous contents were consumed before the barrier
            // that ended iteration bi-1)
            if (bi + 1 < nbatch) {
                const uint32_t nb = base + REC_BATCH + tid;
                if (tid < REC_BATCH && nb < nrec) {
                    cp_async16(&s_rec[(bi + 1) & 1][2 * tid], recq + 2 * (size_t)(it.begin + nb));
                    cp_async16(&s_rec[(bi + 1) & 1][2 * tid + 1], recq + 2 * (size_t)(it.begin + nb) + 1);
                }
            }
            cp_async_commit();
            cp_async_wait<1>();   // batch bi has landed (for this thread); the barrier publishes it block-wide
            __syncthreads();      // (first iteration: also orders the subgrid zeroing before any fold)
            const uint4 *buf = s_rec[bi & 1];

            // DEPTH tap slots in rotation: the taps of records j+1 .. j+DEPTH-1 are requested before the FMAs of
            // record j issue, so each warp keeps DEPTH 128-bit tap loads per residue in flight.
            if constexpr (DEPTH == 2) {
                Slot s0, s1;
                issue(buf, 0, s0);
                uint32_t j = 0;
                for (; j + 1 < m; j += 2) {
                    issue(buf, j + 1, s1);
                    consume(buf, j, s0);
                    if (j + 2 < m) issue(buf, j + 2, s0);
                    consume(buf, j + 1, s1);
                }
                if (j < m) consume(buf, j, s0);
            } else {
                Slot s0, s1, s2;
                issue(buf, 0, s0);
                if (m > 1) issue(buf, 1, s1);
                uint32_t j = 0;
                for (; j + 2 < m; j += 3) {
                    issue(buf, j + 2, s2);
                    consume(buf, j, s0);
                    if (j + 3 < m) issue(buf, j + 3, s0);
                    consume(buf, j + 1, s1);
                    if (j + 4 < m) issue(buf, j + 4, s1);
                    consume(buf, j + 2, s2);
                }
                if (j < m) consume(buf, j, s0);
                if (j + 1 < m) consume(buf, j + 1, s1);
            }
            __syncthreads();  // every warp is done with buf before it is refilled
        }
#pragma unroll
        for (int a = 0; a < CY; ++a)
#pragma unroll
            for (int b = 0; b < CX; ++b) {
                if (acc[a][b].x != 0.0 || acc[a][b].y != 0.0) {
                    double2 t = sg[cell_cur[a][b]];
                    t.x += acc[a][b].x; t.y += acc[a][b].y;
                    sg[cell_cur[a][b]] = t;
                }
            }
        __syncthreads();

        // subgrid -> grid.  Cells outside the owned rows / the grid are dropped (fixoutofbounds).
        const int tyi = (int)(it.tile / (uint32_t)A.ntx), txi = (int)(it.tile % (uint32_t)A.ntx);
        const int gx0 = (txi << A.tshift) - (A.gw - 1), gy0 = (tyi << A.tshift) - (A.gh - 1);
        for (int c = tid; c < ncell; c += NT) {
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
        cp_async_wait<0>();
        // the next iteration's first barrier orders these reads before the subgrid is zeroed again
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Dense-layout gridder (Geom::dense: slices of R x R taps, zero outside the footprint, one leading zero slice).
// Same ownership scheme as grid_tiled_kernel, with the per-record bookkeeping stripped to what the arithmetic needs:
//   * no validity test and no zeroing per residue: every residue loads a tap, the ones outside the footprint read a
//     zero of this slice's padding (or of the previous slice's last rows / columns) and add +0;
//   * the tap address is ONE instruction per residue: a 64-bit per-residue pointer, set up once per micro-tile, plus the
//     record's 32-bit table offset (IMAD.WIDE.U32);
//   * no trip-count tests inside the three-slot rotation: the last batch of an item is padded to a multiple of three
//     with dummy records (vis = 0, offset 0 = the leading zero slice).
// r01 SASS of grid_tiled_kernel<16,2,3,8>: ~43 instructions per record and thread for 8 DFMA + 2 LDG; this kernel:
// see profiles/r02_sass_grid_dense.txt.
template <int R, int TY> struct DenseCfg {
    static constexpr int rec_batch = R == 16 ? 48 : 126;  // multiples of 6
    static constexpr int min_blocks = R == 16 ? (TY == 8 ? 8 : (TY == 4 ? 10 : 4)) : 2;
};

// PAIR: the staged records are split into a visibility array and an array of 8-byte (table offset, loc) pairs, and the loop
// runs six records per trip so that ONE 128-bit shared-memory load brings the pairs of two records: 1.5 instead of 2
// broadcast loads per record and warp on the LSU data pipe the kernel saturates (96 %, 27 points of it these broadcasts).
template <int R, int MT, int TY, int LD = 0, bool PAIR = false>
__global__ void __launch_bounds__(16 * TY, DenseCfg<R, TY>::min_blocks) grid_dense_kernel(const GridArgs A) {
    constexpr int CY = R / TY, CX = R / 16;  // residues per thread
    constexpr int NT = 16 * TY;
    constexpr int REC_BATCH = DenseCfg<R, TY>::rec_batch;
    static_assert(REC_BATCH % 6 == 0 && REC_BATCH <= NT, "one staging thread per record, batches in threes (sixes when PAIR)");
    constexpr uint32_t GROUP = PAIR ? 6u : 3u;   // records per trip: the last batch of an item is padded to a multiple of it
    extern __shared__ double2 sg[];
    // !PAIR: records as they are, [2*j] = visibility, [2*j+1] = (offset, loc, index, tile).  PAIR: [j] = visibility for
    // j < REC_BATCH, then REC_BATCH/2 uint4 holding the (offset, loc) of two records each
    __shared__ __align__(16) uint4 s_rec[2][REC_BATCH * 2];
    __shared__ uint32_t s_item;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int ncell = A.SG * A.SG;
    const uint32_t n_items = A.counters[0];
    constexpr uint32_t mtkey_mask = (uint32_t)(~(MT - 1) & 255) * 0x0101u;
    const uint4 *recq = reinterpret_cast<const uint4 *>(A.rec);

    for (;;) {
        if (tid == 0) s_item = atomicAdd(&A.counters[A.queue], 1u);
        __syncthreads();
        const uint32_t item = s_item;
        if (item >= n_items) break;
        const WorkItem it = A.items[item];
        const uint32_t nrec = it.end - it.begin;
        const uint32_t nbatch = (nrec + REC_BATCH - 1) / REC_BATCH;

        // stages batch `bi` into buffer bi & 1: cp.async for real records, plain stores of dummies up to the next multiple of 3
        auto stage = [&](uint32_t bi) {
            const uint32_t r = bi * REC_BATCH + (uint32_t)tid;
            if (tid < REC_BATCH) {
                uint4 *vdst = PAIR ? &s_rec[bi & 1][tid] : &s_rec[bi & 1][2 * tid];
                uint2 *mdst = PAIR ? reinterpret_cast<uint2 *>(&s_rec[bi & 1][REC_BATCH]) + tid : reinterpret_cast<uint2 *>(&s_rec[bi & 1][2 * tid + 1]);
                if (r < nrec) {
                    cp_async16(vdst, recq + 2 * (size_t)(it.begin + r));
                    if constexpr (PAIR) cp_async8(mdst, recq + 2 * (size_t)(it.begin + r) + 1);
                    else cp_async16(mdst, recq + 2 * (size_t)(it.begin + r) + 1);
                } else if (r < (nrec + GROUP - 1u) / GROUP * GROUP) {
                    *vdst = make_uint4(0u, 0u, 0u, 0u);        // vis = 0
                    *mdst = make_uint2(0u, 0x00010000u);       // table offset 0 (leading zero slice), micro-tile (0,0)
                }
            }
            cp_async_commit();
        };
        stage(0);
        for (int c = tid; c < ncell; c += NT) sg[c] = make_double2(0.0, 0.0);

        double2 acc[CY][CX];
        int cell_cur[CY][CX];
        const double2 *tp[CY][CX];  // table + (tap offset of residue (a,b) inside the R x R slice) for the micro-tile issued last
        uint32_t key_cur = 0xFFFFFFFFu;
#pragma unroll
        for (int a = 0; a < CY; ++a)
#pragma unroll
            for (int b = 0; b < CX; ++b) { acc[a][b] = make_double2(0.0, 0.0); cell_cur[a][b] = 0; tp[a][b] = A.table; }

        struct Slot { double2 k[CY][CX]; int cell[CY][CX]; bool sw; };
        auto issue_m = [&](const uint2 m, Slot &s) {
            const uint32_t key = m.y & mtkey_mask;
            s.sw = key != key_cur;
            if (s.sw) {  // block-uniform: all threads walk the same records
                key_cur = key;
                const int mx = (int)(key & 255u), my = (int)((key >> 8) & 255u);
#pragma unroll
                for (int a = 0; a < CY; ++a) {
                    const int roty = (ty + TY * a - my) & (R - 1);
#pragma unroll
                    for (int b = 0; b < CX; ++b) {
                        const int rotx = (tx + 16 * b - mx) & (R - 1);
                        tp[a][b] = A.table + (roty * R + rotx);  // kpitch == R in the dense layout
                        s.cell[a][b] = (my + roty) * A.SG + mx + rotx;
                    }
                }
            }
#pragma unroll
            for (int a = 0; a < CY; ++a)
#pragma unroll
                for (int b = 0; b < CX; ++b) s.k[a][b] = ld_tap<LD>(tp[a][b] + m.x);
        };
        auto issue = [&](const uint4 *buf, uint32_t j, Slot &s) {  // !PAIR: one 64-bit broadcast read per record
            issue_m(*reinterpret_cast<const uint2 *>(&buf[2 * j + 1]), s);
        };
        auto consume = [&](const uint4 *buf, uint32_t j, const Slot &s) {
            const double2 vis = *reinterpret_cast<const double2 *>(&buf[PAIR ? j : 2 * j]);
            if (s.sw) {  // fold the register accumulators into the thread's own subgrid cells and retarget them
#pragma unroll
                for (int a = 0; a < CY; ++a)
#pragma unroll
                    for (int b = 0; b < CX; ++b) {
                        if (acc[a][b].x != 0.0 || acc[a][b].y != 0.0) {
                            double2 t = sg[cell_cur[a][b]];
                            t.x += acc[a][b].x; t.y += acc[a][b].y;
                            sg[cell_cur[a][b]] = t;
                            acc[a][b] = make_double2(0.0, 0.0);
                        }
                        cell_cur[a][b] = s.cell[a][b];
                    }
            }
#pragma unroll
            for (int a = 0; a < CY; ++a)
#pragma unroll
                for (int b = 0; b < CX; ++b) {  // (vr + i vi)(kr + i ki)
                    acc[a][b].x = fma(vis.x, s.k[a][b].x, acc[a][b].x);
                    acc[a][b].x = fma(-vis.y, s.k[a][b].y, acc[a][b].x);
                    acc[a][b].y = fma(vis.x, s.k[a][b].y, acc[a][b].y);
                    acc[a][b].y = fma(vis.y, s.k[a][b].x, acc[a][b].y);
                }
        };

        for (uint32_t bi = 0; bi < nbatch; ++bi) {
            if (bi + 1 < nbatch) stage(bi + 1); else cp_async_commit();
            cp_async_wait<1>();   // batch bi has landed (for this thread); the barrier publishes it block-wide
            __syncthreads();      // (first iteration: also orders the subgrid zeroing before any fold)
            const uint4 *buf = s_rec[bi & 1];
            const uint32_t m = min((uint32_t)REC_BATCH, nrec - bi * REC_BATCH);
            const uint32_t m3 = (m + GROUP - 1u) / GROUP * GROUP;  // records incl. dummies
            Slot s0, s1, s2;
            if constexpr (!PAIR) {
                issue(buf, 0, s0);
                issue(buf, 1, s1);
                uint32_t j = 0;
                for (; j + 3 < m3; j += 3) {
                    issue(buf, j + 2, s2);
                    consume(buf, j, s0);
                    issue(buf, j + 3, s0);
                    consume(buf, j + 1, s1);
                    issue(buf, j + 4, s1);
                    consume(buf, j + 2, s2);
                }
                issue(buf, j + 2, s2);
                consume(buf, j, s0);
                consume(buf, j + 1, s1);
                consume(buf, j + 2, s2);
            } else {
                const uint4 *meta = buf + REC_BATCH;  // meta[p] = (offset, loc) of records 2p and 2p+1
                uint4 mm = meta[0];
                issue_m(make_uint2(mm.x, mm.y), s0);
                issue_m(make_uint2(mm.z, mm.w), s1);
                uint32_t j = 0;
                for (; j + 6 < m3; j += 6) {
                    mm = meta[(j >> 1) + 1];
                    issue_m(make_uint2(mm.x, mm.y), s2);
                    consume(buf, j, s0);
                    issue_m(make_uint2(mm.z, mm.w), s0);
                    consume(buf, j + 1, s1);
                    mm = meta[(j >> 1) + 2];
                    issue_m(make_uint2(mm.x, mm.y), s1);
                    consume(buf, j + 2, s2);
                    issue_m(make_uint2(mm.z, mm.w), s2);
                    consume(buf, j + 3, s0);
                    mm = meta[(j >> 1) + 3];
                    issue_m(make_uint2(mm.x, mm.y), s0);
                    consume(buf, j + 4, s1);
                    issue_m(make_uint2(mm.z, mm.w), s1);
                    consume(buf, j + 5, s2);
                }
                mm = meta[(j >> 1) + 1];
                issue_m(make_uint2(mm.x, mm.y), s2);
                consume(buf, j, s0);
                issue_m(make_uint2(mm.z, mm.w), s0);
                consume(buf, j + 1, s1);
                mm = meta[(j >> 1) + 2];
                issue_m(make_uint2(mm.x, mm.y), s1);
                consume(buf, j + 2, s2);
                issue_m(make_uint2(mm.z, mm.w), s2);
                consume(buf, j + 3, s0);
                consume(buf, j + 4, s1);
                consume(buf, j + 5, s2);
            }
            __syncthreads();  // every warp is done with buf before it is refilled
        }
#pragma unroll
        for (int a = 0; a < CY; ++a)
#pragma unroll
            for (int b = 0; b < CX; ++b) {
                if (acc[a][b].x != 0.0 || acc[a][b].y != 0.0) {
                    double2 t = sg[cell_cur[a][b]];
                    t.x += acc[a][b].x; t.y += acc[a][b].y;
                    sg[cell_cur[a][b]] = t;
                }
            }
        __syncthreads();

        // subgrid -> grid.  Cells outside the owned rows / the grid are dropped (fixoutofbounds).
        const int tyi = (int)(it.tile / (uint32_t)A.ntx), txi = (int)(it.tile % (uint32_t)A.ntx);
        const int gx0 = (txi << A.tshift) - (A.gw - 1), gy0 = (tyi << A.tshift) - (A.gh - 1);
        for (int c = tid; c < ncell; c += NT) {
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
        cp_async_wait<0>();
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
        const int lx = (int)(meta.y & 255u), ly = (int)((meta.y >> 8) & 255u);
        const uint32_t dx = (uint32_t)(lx & ~A.mt_mask), dy = (uint32_t)(ly & ~A.mt_mask);
        const int tyi = (int)(meta.w / (uint32_t)A.ntx), txi = (int)(meta.w % (uint32_t)A.ntx);
        const int gx = (txi << A.tshift) + lx - (A.gw - 1) + j, gy = (tyi << A.tshift) + ly - (A.gh - 1) + i;
        if ((unsigned)gx >= (unsigned)A.width || (unsigned)gy >= (unsigned)A.nrows) continue;
        const double2 k = ldg2(A.table + (uint32_t)(meta.x + (dy + (uint32_t)i) * (uint32_t)A.kpitch + dx + (uint32_t)j));
        double *g = reinterpret_cast<double *>(A.grid + (size_t)gy * A.width + gx);
        red_add(g, vis.x * k.x - vis.y * k.y);
        red_add(g + 1, vis.x * k.y + vis.y * k.x);
    }
}

// Degridder: the adjoint gather.  Half a warp per visibility: lane j of the half-warp walks column j (+16, +32..)
// of the footprint row by row, so the kernel taps and the grid cells are both contiguous 16-byte loads; the
// partial sums are combined with four shuffle steps.
//   vis_out[index] = sum_{i,j} conj(table[slice][i,j]) * grid[y0 + i, x0 + j]
// Records are in tile order, so the warps of a block read the same few KB of the grid (L1/L2 hits).
template <int UNROLL>
__global__ void __launch_bounds__(256) degrid_warp_kernel(const GridArgs A) {
    const int hl = threadIdx.x & 15;
    const i64 hw0 = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const i64 nhw = ((i64)gridDim.x * blockDim.x) >> 4;
    const i64 count = (i64)A.counters[2];
    const i64 rounds = (count + nhw - 1) / nhw;
    for (i64 it = 0; it < rounds; ++it) {
        const i64 r = it * nhw + hw0;   // both halves of a warp stay in the loop together (shuffles below)
        const bool live = r < count;
        double ar = 0.0, ai = 0.0;
        uint32_t out_index = 0;
        if (live) {
            const uint4 meta = __ldg(reinterpret_cast<const uint4 *>(A.rec + r) + 1);
            out_index = meta.z;
            const int lx = (int)(meta.y & 255u), ly = (int)((meta.y >> 8) & 255u);
            const uint32_t dx = (uint32_t)(lx & ~A.mt_mask), dy = (uint32_t)(ly & ~A.mt_mask);
            const int tyi = (int)(meta.w / (uint32_t)A.ntx), txi = (int)(meta.w % (uint32_t)A.ntx);
            const int gx0 = (txi << A.tshift) + lx - (A.gw - 1), gy0 = (tyi << A.tshift) + ly - (A.gh - 1);
            const uint32_t kslice = meta.x + dy * (uint32_t)A.kpitch + dx;
            const int i0 = max(0, -gy0), i1 = min(A.gh, A.nrows - gy0);
            for (int j = hl; j < A.gw; j += 16) {
                const int gx = gx0 + j;
                if ((unsigned)gx >= (unsigned)A.width) continue;
                const double2 *kp = A.table + (uint32_t)(kslice + (uint32_t)(i0 * A.kpitch + j));
                const double2 *gp = A.grid + (size_t)(gy0 + i0) * A.width + gx;
#pragma unroll UNROLL
                for (int i = i0; i < i1; ++i) {
                    const double2 k = ldg2(kp);
                    const double2 g = ldg2(gp);
                    // conj(k) * g
                    ar = fma(k.x, g.x, ar); ar = fma(k.y, g.y, ar);
                    ai = fma(k.x, g.y, ai); ai = fma(-k.y, g.x, ai);
                    kp += A.kpitch; gp += A.width;
                }
            }
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            ar += __shfl_xor_sync(0xffffffffu, ar, o);
            ai += __shfl_xor_sync(0xffffffffu, ai, o);
        }
        if (live && hl == 0) A.vis_out[A.out_by_record ? (uint32_t)r : out_index] = make_double2(ar, ai);
    }
}

// Degridder, tiled: a persistent block pulls the same work items as the gridder, stages the item's subgrid
// (tile-1+S)^2 from the grid into shared memory once (zero outside the owned rows / the grid, so the hot loop has no
// bounds checks), then its 16 half-warps take the item's records round-robin.  Grid cells are then conflict-free
// 128-bit shared-memory loads (two wavefronts per 15-lane row instead of three unaligned L1 lines); only the kernel
// taps still come from L2.
template <int UNROLL, int NT>
__global__ void __launch_bounds__(NT, 1024 / NT) degrid_tile_kernel(const GridArgs A) {
    constexpr int NHW = NT / 16;  // half-warps per block
    extern __shared__ double2 sg[];
    __shared__ uint32_t s_item;
    const int tid = threadIdx.x, hl = tid & 15, hw = tid >> 4;
    const int SGW = A.tile - 1 + A.gw, SGH = A.tile - 1 + A.gh;  // staged region (pitch SGW)
    const uint32_t n_items = A.counters[0];
    for (;;) {
        __syncthreads();  // all half-warps are done with the previous subgrid
        if (tid == 0) s_item = atomicAdd(&A.counters[A.queue], 1u);
        __syncthreads();
        const uint32_t item = s_item;
        if (item >= n_items) break;
        const WorkItem it = A.items[item];
        const int tyi = (int)(it.tile / (uint32_t)A.ntx), txi = (int)(it.tile % (uint32_t)A.ntx);
        const int gx0 = (txi << A.tshift) - (A.gw - 1), gy0 = (tyi << A.tshift) - (A.gh - 1);
        for (int c = tid; c < SGW * SGH; c += NT) {
            const int cy = c / SGW, cx = c - cy * SGW;
            const int gx = gx0 + cx, gy = gy0 + cy;
            double2 v = make_double2(0.0, 0.0);
            if ((unsigned)gx < (unsigned)A.width && (unsigned)gy < (unsigned)A.nrows) v = ldg2(A.grid + (size_t)gy * A.width + gx);
            sg[c] = v;
        }
        __syncthreads();
        const uint32_t nrec = it.end - it.begin;
        // contiguous run of records per half-warp, walked in blocks of 16 whose second halves (table offset, loc, output index)
        // arrive with one request per block, prefetched one block ahead, and are broadcast by shuffle (as degrid_reg_kernel)
        const uint32_t chunk = (nrec + (uint32_t)NHW - 1u) / (uint32_t)NHW;
        const uint32_t r0 = (uint32_t)hw * chunk;
        const uint32_t r1 = min(nrec, r0 + chunk);
        const uint4 *recm = reinterpret_cast<const uint4 *>(A.rec + it.begin) + 1;
        auto load_meta = [&](uint32_t q0) {
            const uint32_t r = r0 + q0 + (uint32_t)hl;
            return r < r1 ? __ldg(recm + 2 * (size_t)r) : make_uint4(0u, 0u, 0u, 0u);
        };
        uint4 meta_next = load_meta(0);
        for (uint32_t q0 = 0; q0 < chunk; q0 += 16) {  // both halves of a warp run the same trip count (shuffles below)
            const uint4 meta = meta_next;
            if (q0 + 16 < chunk) meta_next = load_meta(q0 + 16);
            double2 res = make_double2(0.0, 0.0);
#pragma unroll 1
            for (int qq = 0; qq < 16; ++qq) {
                const uint32_t kb = __shfl_sync(0xffffffffu, meta.x, qq, 16);
                const uint32_t loc = __shfl_sync(0xffffffffu, meta.y, qq, 16);
                const bool live = r0 + q0 + (uint32_t)qq < r1;
                double ar = 0.0, ai = 0.0;
                if (live) {
                    const int lx = (int)(loc & 255u), ly = (int)((loc >> 8) & 255u);
                    const uint32_t dx = (uint32_t)(lx & ~A.mt_mask), dy = (uint32_t)(ly & ~A.mt_mask);
                    const uint32_t kslice = kb + dy * (uint32_t)A.kpitch + dx;
                    for (int j = hl; j < A.gw; j += 16) {
                        const double2 *kp = A.table + (uint32_t)(kslice + (uint32_t)j);
                        const double2 *gp = sg + ly * SGW + lx + j;
#pragma unroll UNROLL
                        for (int i = 0; i < A.gh; ++i) {
                            const double2 k = ldg2(kp);
                            const double2 g = *gp;
                            ar = fma(k.x, g.x, ar); ar = fma(k.y, g.y, ar);   // conj(k) * g
                            ai = fma(k.x, g.y, ai); ai = fma(-k.y, g.x, ai);
                            kp += A.kpitch; gp += SGW;
                        }
                    }
                }
                const bool hi = (hl & 8) != 0;
                const double send = hi ? ar : ai, keep = hi ? ai : ar;
                double sum = keep + __shfl_xor_sync(0xffffffffu, send, 8);
#pragma unroll
                for (int o = 4; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                const double other = __shfl_xor_sync(0xffffffffu, sum, 8);
                if (hl == qq) res = hi ? make_double2(other, sum) : make_double2(sum, other);
            }
            if (r0 + q0 + (uint32_t)hl < r1) A.vis_out[A.out_by_record ? it.begin + r0 + q0 + (uint32_t)hl : meta.z] = res;
        }
    }
}

// Degridder, register-resident grid column (gh, gw <= 16; plans with cell-granular buckets): as degrid_tile_kernel,
// but every half-warp takes a CONTIGUOUS run of the item's records, which the plan has grouped by exact footprint
// origin, and keeps its column of the footprint's grid cells (gh complex values per lane) in registers while the origin
// does not change.  In dense uv regions hundreds of consecutive visibilities share the origin, so the per-visibility
// work is just the tap stream: gh 128-bit loads and 4*gh FMAs per lane -- half of the L1/shared-memory wavefronts of
// the tiled kernel, which is bound by exactly that pipe.
template <int GH, int NT, int MINB, bool EXACT, int LD = 0>
__global__ void __launch_bounds__(NT, MINB) degrid_reg_kernel(const GridArgs A) {
    // EXACT: gh == GH, so the row loops carry no predicate; the padded table (pitch 16 >= gw, zero pad columns) lets every
    // lane load its tap unconditionally, lanes >= gw just read the pad
    constexpr int NHW = NT / 16;
    extern __shared__ double2 sg[];
    __shared__ uint32_t s_item;
    const int tid = threadIdx.x, hl = tid & 15, hw = tid >> 4;
    const int SGW = A.tile - 1 + A.gw, SGH = A.tile - 1 + A.gh;
    const uint32_t n_items = A.counters[0];
    const bool col = hl < A.gw;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = atomicAdd(&A.counters[A.queue], 1u);
        __syncthreads();
        const uint32_t item = s_item;
        if (item >= n_items) break;
        const WorkItem it = A.items[item];
        const int tyi = (int)(it.tile / (uint32_t)A.ntx), txi = (int)(it.tile % (uint32_t)A.ntx);
        const int gx0 = (txi << A.tshift) - (A.gw - 1), gy0 = (tyi << A.tshift) - (A.gh - 1);
        for (int c = tid; c < SGW * SGH; c += NT) {
            const int cy = c / SGW, cx = c - cy * SGW;
            const int gx = gx0 + cx, gy = gy0 + cy;
            double2 v = make_double2(0.0, 0.0);
            if ((unsigned)gx < (unsigned)A.width && (unsigned)gy < (unsigned)A.nrows) v = ldg2(A.grid + (size_t)gy * A.width + gx);
            sg[c] = v;
        }
        __syncthreads();
        const uint32_t nrec = it.end - it.begin;
        const uint32_t chunk = (nrec + (uint32_t)NHW - 1u) / (uint32_t)NHW;  // contiguous records per half-warp
        const uint32_t r0 = (uint32_t)hw * chunk;
        const uint32_t r1 = min(nrec, r0 + chunk);  // this half-warp's run is [r0, r1)
        double2 g[GH];
#pragma unroll
        for (int i = 0; i < GH; ++i) g[i] = make_double2(0.0, 0.0);
        uint32_t cur = 0xFFFFFFFFu;
        // The run is walked in blocks of 16 records: lane l of the half-warp loads the second half (table offset, loc, output
        // index, tile) of record l of the block with ONE coalesced request -- the next block's while the current one is
        // processed -- and the 16 records are then broadcast by shuffle.  (r01: one dependent global load per record in
        // front of its tap loads; ncu long_scoreboard 5.0, 25 % occupancy.)  Lane l also keeps the result of record l,
        // so a block ends with one 16-byte store per lane instead of two 8-byte stores per record.
        const uint4 *recm = reinterpret_cast<const uint4 *>(A.rec + it.begin) + 1;
        auto load_meta = [&](uint32_t q0) {
            const uint32_t r = r0 + q0 + (uint32_t)hl;
            return r < r1 ? __ldg(recm + 2 * (size_t)r) : make_uint4(0u, 0u, 0u, 0u);
        };
        uint4 meta_next = load_meta(0);
        for (uint32_t q0 = 0; q0 < chunk; q0 += 16) {  // both halves of a warp run the same trip count (shuffles below)
            const uint4 meta = meta_next;
            if (q0 + 16 < chunk) meta_next = load_meta(q0 + 16);
            double2 res = make_double2(0.0, 0.0);
#pragma unroll 1
            for (int qq = 0; qq < 16; ++qq) {
                const uint32_t kb = __shfl_sync(0xffffffffu, meta.x, qq, 16);
                const uint32_t origin = __shfl_sync(0xffffffffu, meta.y, qq, 16) & 0xFFFFu;
                const bool live = r0 + q0 + (uint32_t)qq < r1;
                double ar = 0.0, ai = 0.0;
                if (live) {
                    if (origin != cur) {  // half-warp-uniform: (re)load this lane's column of the footprint
                        cur = origin;
                        const int lx = (int)(origin & 255u), ly = (int)(origin >> 8);
                        const double2 *gp = sg + ly * SGW + lx + hl;
#pragma unroll
                        for (int i = 0; i < GH; ++i) g[i] = (col && (EXACT || i < A.gh)) ? gp[i * SGW] : make_double2(0.0, 0.0);
                    }
                    const int lx = (int)(origin & 255u), ly = (int)(origin >> 8);
                    const uint32_t dx = (uint32_t)(lx & ~A.mt_mask), dy = (uint32_t)(ly & ~A.mt_mask);
                    const double2 *kp = A.table + (uint32_t)(kb + dy * (uint32_t)A.kpitch + dx + (uint32_t)hl);
                    double2 k[GH];
#pragma unroll
                    for (int i = 0; i < GH; ++i) k[i] = (EXACT || i < A.gh) ? ld_tap<LD>(kp + i * 16) : make_double2(0.0, 0.0);  // kpitch == 16 here: immediate offsets
#pragma unroll
                    for (int i = 0; i < GH; ++i) {  // conj(k) * g
                        ar = fma(k[i].x, g[i].x, ar); ar = fma(k[i].y, g[i].y, ar);
                        ai = fma(k[i].x, g[i].y, ai); ai = fma(-k[i].y, g[i].x, ai);
                    }
                }
                // 16-lane reduction of (ar, ai) with half the shuffles: after the first exchange the lanes with bit 3 clear
                // carry only the real sum and the others only the imaginary one, so the remaining three steps move one double;
                // a last exchange gives every lane both
                const bool hi = (hl & 8) != 0;
                const double send = hi ? ar : ai, keep = hi ? ai : ar;
                double sum = keep + __shfl_xor_sync(0xffffffffu, send, 8);
#pragma unroll
                for (int o = 4; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                const double other = __shfl_xor_sync(0xffffffffu, sum, 8);
                if (hl == qq) res = hi ? make_double2(other, sum) : make_double2(sum, other);
            }
            if (r0 + q0 + (uint32_t)hl < r1) A.vis_out[A.out_by_record ? it.begin + r0 + q0 + (uint32_t)hl : meta.z] = res;
        }
    }
}

// Padded copy of the caller's kernel table: [lead + slice][krows][kpitch], zero outside the gh x gw taps (and in the
// `lead` leading slices of the dense layout).  A few MB for a w-kernel table (microseconds), one S x S kernel per
// visibility on the AW path.
__global__ void __launch_bounds__(256) pad_table_kernel(const double2 *__restrict__ in, double2 *__restrict__ out, i64 nslices, int gh, int gw,
                                                        int krows, int kpitch, int lead) {
    const i64 per = (i64)krows * kpitch;
    const i64 total = (nslices + lead) * per;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += stride) {
        const int j = (int)(c % kpitch);
        const i64 row = c / kpitch;
        const int i = (int)(row % krows);
        const i64 sl = row / krows - lead;
        out[c] = (sl >= 0 && i < gh && j < gw) ? in[(sl * gh + i) * gw + j] : make_double2(0.0, 0.0);
    }
}

static int prepare_table(skagrid_ctx *ctx, skagrid_plan *plan, const double *table, cudaStream_t st, const double **out) {
    const Geom &g = plan->g;
    if (!g.dense && g.kpitch == (int)g.gw) { *out = table; return SKAGRID_OK; }
    const int lead = g.dense ? 1 : 0;
    const i64 nslices = plan->slice_override ? plan->count : g.nw * g.qpx * g.qpx;
    const i64 cap_slices = plan->slice_override ? plan->capacity : nslices;
    const size_t need = (size_t)((cap_slices + lead) * g.krows * g.kpitch) * sizeof(double2);
    if (plan->table_bytes < need) {
        if (plan->d_table) { SK_CUDA(ctx, cudaStreamSynchronize(st)); cudaFree(plan->d_table); plan->d_table = nullptr; plan->table_bytes = 0; }
        if (cudaMalloc(&plan->d_table, need) != cudaSuccess) {
            cudaGetLastError();
            return sk_fail(ctx, SKAGRID_ENOMEM, "padded kernel table: %zu bytes", need);
        }
        plan->table_bytes = need;
    }
    if (nslices > 0) {
        i64 b = ((nslices + lead) * g.krows * g.kpitch + 255) / 256;
        if (b > (i64)ctx->sm_count * 16) b = (i64)ctx->sm_count * 16;
        pad_table_kernel<<<(unsigned)b, 256, 0, st>>>(reinterpret_cast<const double2 *>(table), plan->d_table, nslices, (int)g.gh, (int)g.gw, g.krows,
                                                       g.kpitch, lead);
        SK_LAUNCH_CHECK(ctx);
    }
    *out = reinterpret_cast<const double *>(plan->d_table);
    return SKAGRID_OK;
}

static GridArgs make_args(skagrid_plan *plan, const double *table, double *grid) {
    const Geom &g = plan->g;
    GridArgs A;
    A.rec = plan->d_rec; A.items = plan->d_items; A.counters = plan->d_counters;
    A.table = reinterpret_cast<const double2 *>(table);
    A.grid = reinterpret_cast<double2 *>(grid);
    A.vis_out = nullptr;
    A.gh = (int)g.gh; A.gw = (int)g.gw; A.s2 = (int)(g.gh * g.gw);
    A.kpitch = g.kpitch;
    A.tile = g.tile; A.tshift = g.tshift;
    A.mt_mask = ~(g.MT - 1);
    A.out_by_record = 0;
    A.SG = g.SG; A.ntx = g.ntx;
    A.width = (int)g.width; A.nrows = (int)(g.row1 - g.row0);
    A.queue = 1;
    return A;
}

template <int R, int MT, int DEPTH, int TY>
static int launch_tiled(skagrid_ctx *ctx, const GridArgs &A, cudaStream_t st) {
    constexpr int NT = 16 * TY;
    const size_t smem = (size_t)A.SG * A.SG * sizeof(double2);
    if (ctx->smem_configured.insert((const void *)grid_tiled_kernel<R, MT, DEPTH, TY>).second)  // per device, hence per context
        SK_CUDA(ctx, cudaFuncSetAttribute(grid_tiled_kernel<R, MT, DEPTH, TY>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    int per_sm = 0;
    SK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, grid_tiled_kernel<R, MT, DEPTH, TY>, NT, smem));
    if (per_sm < 1) return sk_fail(ctx, SKAGRID_ECUDA, "tiled gridder does not fit on an SM (smem %zu)", smem);
    grid_tiled_kernel<R, MT, DEPTH, TY><<<ctx->sm_count * per_sm, NT, smem, st>>>(A);
    SK_LAUNCH_CHECK(ctx);
    return SKAGRID_OK;
}

template <int R, int MT, int TY, int LD = 0, bool PAIR = false>
static int launch_dense(skagrid_ctx *ctx, const GridArgs &A, cudaStream_t st) {
    constexpr int NT = 16 * TY;
    const size_t smem = (size_t)A.SG * A.SG * sizeof(double2);
    if (ctx->smem_configured.insert((const void *)grid_dense_kernel<R, MT, TY, LD, PAIR>).second)
        SK_CUDA(ctx, cudaFuncSetAttribute(grid_dense_kernel<R, MT, TY, LD, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    int per_sm = 0;
    SK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, grid_dense_kernel<R, MT, TY, LD, PAIR>, NT, smem));
    if (per_sm < 1) return sk_fail(ctx, SKAGRID_ECUDA, "dense gridder does not fit on an SM (smem %zu)", smem);
    grid_dense_kernel<R, MT, TY, LD, PAIR><<<ctx->sm_count * per_sm, NT, smem, st>>>(A);
    SK_LAUNCH_CHECK(ctx);
    return SKAGRID_OK;
}

extern "C" int skagrid_dev_grid(skagrid_ctx *ctx, skagrid_plan *plan, const double *table, double *grid, int variant, void *stream) {
    if (!ctx || !plan || !table || !grid) return SKAGRID_EINVAL;
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = sk_stream(ctx, stream);
    if (plan->count == 0) return SKAGRID_OK;
    if (!plan->has_vis) return sk_fail(ctx, SKAGRID_EINVAL, "grid: the plan was built without visibilities (degrid-only)");
    const double *ptab;
    SK_TRY(prepare_table(ctx, plan, table, st, &ptab));
    GridArgs A = make_args(plan, ptab, grid);
    const int R = plan->g.R;
    if (variant == 1 || R == 0) {
        grid_atomic_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(A);
        SK_LAUNCH_CHECK(ctx);
        return SKAGRID_OK;
    }
    SK_CUDA(ctx, cudaMemsetAsync(plan->d_counters + 1, 0, sizeof(uint32_t), st));
    const int MT = plan->g.MT;
    if (plan->g.dense && variant != 2 && variant != 3 && variant != 4) {  // (6 is a dense variant)
        // dense layout (S = 14, 15, 29..31).  B200, r02 (profiles/r02_gridder_ab.md): R=16, 1e8 visibilities: 8x16 threads with two
        // residues each 21.3 ms (default), 4x16 threads with four residues 22.5 ms (variant 5), r01 predicated kernel 21.65 ms;
        // R=32 (S=31), 5e7 visibilities: 16x16 threads with four residues 42.8 ms (default), 32x16 threads with two 44.2 ms
        // (variant 5), r01 kernel 44.6 ms
        if (R == 16) {
            if (variant == 5) return MT == 2 ? launch_dense<16, 2, 4>(ctx, A, st) : launch_dense<16, 4, 4>(ctx, A, st);
            if (variant == 6) return MT == 2 ? launch_dense<16, 2, 16>(ctx, A, st) : launch_dense<16, 4, 16>(ctx, A, st);  // 16x16 threads, one residue each
            if (variant == 7 && MT == 2) return launch_dense<16, 2, 8, 0, true>(ctx, A, st);  // paired (offset, loc) broadcasts
            if (MT == 2 && tap_load_mode() == 1) return launch_dense<16, 2, 8, 1>(ctx, A, st);
            if (MT == 2 && tap_load_mode() == 2) return launch_dense<16, 2, 8, 2>(ctx, A, st);
            return MT == 2 ? launch_dense<16, 2, 8>(ctx, A, st) : launch_dense<16, 4, 8>(ctx, A, st);
        }
        if (variant == 5) return MT == 2 ? launch_dense<32, 2, 32>(ctx, A, st) : launch_dense<32, 4, 32>(ctx, A, st);
        if (variant == 7 && MT == 2) return launch_dense<32, 2, 16, 0, true>(ctx, A, st);
        if (MT == 2 && tap_load_mode() == 1) return launch_dense<32, 2, 16, 1>(ctx, A, st);
        if (MT == 2 && tap_load_mode() == 2) return launch_dense<32, 2, 16, 2>(ctx, A, st);
        return MT == 2 ? launch_dense<32, 2, 16>(ctx, A, st) : launch_dense<32, 4, 16>(ctx, A, st);
    }
    // variants (A/B measurements on B200; 1e8 visibilities, config 4, per launch):
    //   0  default.  R=16: 8x16 threads, two residues each, three tap slots (22.5 ms)
    //                R=32: 32x16 threads, two residues each, three tap slots (S=31, 5e7 on 32768^2: 62 ms)
    //   2  16x16 threads, R/16 x R/16 residues each, two tap slots (R=16: 27.6 ms with 32-cell tiles; R=32: 71 ms)
    //   3  R=16: 16x16 threads, three tap slots (27.1 ms with 32-cell tiles)
    //   4  R=16: 8x16 threads, two tap slots (23.5 ms)
    if (R == 16) {
        if (variant == 2) return MT == 2 ? launch_tiled<16, 2, 2, 16>(ctx, A, st) : launch_tiled<16, 4, 2, 16>(ctx, A, st);
        if (variant == 3) return MT == 2 ? launch_tiled<16, 2, 3, 16>(ctx, A, st) : launch_tiled<16, 4, 3, 16>(ctx, A, st);
        if (variant == 4) return MT == 2 ? launch_tiled<16, 2, 2, 8>(ctx, A, st) : launch_tiled<16, 4, 2, 8>(ctx, A, st);
        return MT == 2 ? launch_tiled<16, 2, 3, 8>(ctx, A, st) : launch_tiled<16, 4, 3, 8>(ctx, A, st);
    }
    if (R == 32) {
        if (variant == 2) return MT == 2 ? launch_tiled<32, 2, 2, 16>(ctx, A, st) : launch_tiled<32, 4, 2, 16>(ctx, A, st);
        return MT == 2 ? launch_tiled<32, 2, 3, 32>(ctx, A, st) : launch_tiled<32, 4, 3, 32>(ctx, A, st);
    }
    if (R == 48) return MT == 2 ? launch_tiled<48, 2, 2, 24>(ctx, A, st) : launch_tiled<48, 4, 2, 24>(ctx, A, st);  // 24x16 threads, 2x3 residues each
    return MT == 2 ? launch_tiled<64, 2, 2, 32>(ctx, A, st) : launch_tiled<64, 4, 2, 32>(ctx, A, st);  // 32x16 threads, 2x4 residues each
}

static int degrid_impl(skagrid_ctx *ctx, skagrid_plan *plan, const double *table, const double *grid, double *vis_out, int plan_order, void *stream);

extern "C" int skagrid_dev_degrid(skagrid_ctx *ctx, skagrid_plan *plan, const double *table, const double *grid, double *vis_out,
                                  void *stream) {
    return degrid_impl(ctx, plan, table, grid, vis_out, 0, stream);
}

// The same with the results in the plan's own order: vis_out[r] belongs to record r (r < kept; see skagrid_dev_plan_order), a
// sequential stream of full 32-byte sectors instead of one half-sector store at a random caller index per visibility (ncu r02:
// 3.2 GB of DRAM writes for 1.6 GB of results).  For callers that keep their visibilities in plan order across major cycles.
extern "C" int skagrid_dev_degrid_plan_order(skagrid_ctx *ctx, skagrid_plan *plan, const double *table, const double *grid, double *vis_out,
                                             void *stream) {
    return degrid_impl(ctx, plan, table, grid, vis_out, 1, stream);
}

static int degrid_impl(skagrid_ctx *ctx, skagrid_plan *plan, const double *table, const double *grid, double *vis_out, int plan_order, void *stream) {
    if (!ctx || !plan || !table || !grid || !vis_out) return SKAGRID_EINVAL;
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = sk_stream(ctx, stream);
    if (plan->count == 0) return SKAGRID_OK;
    // visibilities without a tap on the owned rows are not in the record list: their output is 0 (caller order); in plan order every
    // kept record is written, nothing to clear
    if (!plan_order) SK_CUDA(ctx, cudaMemsetAsync(vis_out, 0, (size_t)plan->count * sizeof(double2), st));
    const double *ptab;
    SK_TRY(prepare_table(ctx, plan, table, st, &ptab));
    GridArgs A = make_args(plan, ptab, const_cast<double *>(grid));
    A.vis_out = reinterpret_cast<double2 *>(vis_out);
    A.out_by_record = plan_order;
    // SKAGRID_DEGRID_VARIANT (A/B measurements): 0 register-column kernel when applicable (4 blocks/SM, 128 registers),
    // 4 the same squeezed to 5 blocks/SM, 3 tiled (128 threads), 2 tiled (256 threads), 1 untiled.
    // B200, S=15, 1e8 visibilities: untiled 35.1 ms, tiled 28.6 ms, register-column 23.9 ms (5 blocks: 25.1 ms)
    static const int variant = getenv("SKAGRID_DEGRID_VARIANT") ? atoi(getenv("SKAGRID_DEGRID_VARIANT")) : 0;
    const size_t tile_smem = (size_t)(A.tile - 1 + A.gw) * (A.tile - 1 + A.gh) * sizeof(double2);
    if (variant != 1 && variant != 2 && variant != 3 && plan->g.cellsort && A.gh <= 16 && A.gw <= 16 && A.kpitch == 16 && tile_smem <= 48 * 1024) {
        A.queue = 5;
        SK_CUDA(ctx, cudaMemsetAsync(plan->d_counters + 5, 0, sizeof(uint32_t), st));
        int per_sm = 0;
        if (A.gh == 15 && variant == 4) {
            SK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, degrid_reg_kernel<15, 128, 5, true>, 128, tile_smem));
            degrid_reg_kernel<15, 128, 5, true><<<ctx->sm_count * (per_sm < 1 ? 1 : per_sm), 128, tile_smem, st>>>(A);
        } else if (A.gh == 15 && tap_load_mode() == 1) {
            SK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, degrid_reg_kernel<15, 128, 4, true, 1>, 128, tile_smem));
            degrid_reg_kernel<15, 128, 4, true, 1><<<ctx->sm_count * (per_sm < 1 ? 1 : per_sm), 128, tile_smem, st>>>(A);
        } else if (A.gh == 15 && tap_load_mode() == 2) {
            SK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, degrid_reg_kernel<15, 128, 4, true, 2>, 128, tile_smem));
            degrid_reg_kernel<15, 128, 4, true, 2><<<ctx->sm_count * (per_sm < 1 ? 1 : per_sm), 128, tile_smem, st>>>(A);
        } else if (A.gh == 15) {
            SK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, degrid_reg_kernel<15, 128, 4, true>, 128, tile_smem));
            degrid_reg_kernel<15, 128, 4, true><<<ctx->sm_count * (per_sm < 1 ? 1 : per_sm), 128, tile_smem, st>>>(A);
        } else {
            SK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, degrid_reg_kernel<16, 128, 4, false>, 128, tile_smem));
            degrid_reg_kernel<16, 128, 4, false><<<ctx->sm_count * (per_sm < 1 ? 1 : per_sm), 128, tile_smem, st>>>(A);
        }
    } else if (variant != 1 && tile_smem <= 110 * 1024) {  // at least two blocks per SM; beyond that the untiled kernel wins (S=63: 110 vs 281 ms per 2e7)
        A.queue = 5;
        SK_CUDA(ctx, cudaMemsetAsync(plan->d_counters + 5, 0, sizeof(uint32_t), st));
        if (ctx->smem_configured.insert((const void *)degrid_tile_kernel<15, 256>).second) {
            SK_CUDA(ctx, cudaFuncSetAttribute(degrid_tile_kernel<15, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            SK_CUDA(ctx, cudaFuncSetAttribute(degrid_tile_kernel<15, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        }
        int per_sm = 0;
        if (variant == 2 || (variant == 0 && A.tile == 32)) {  // large tiles: more threads to stage the subgrid (28.3 vs 31.8 ms at S=15)
            SK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, degrid_tile_kernel<15, 256>, 256, tile_smem));
            if (per_sm < 1) per_sm = 1;
            degrid_tile_kernel<15, 256><<<ctx->sm_count * per_sm, 256, tile_smem, st>>>(A);
        } else {
            SK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, degrid_tile_kernel<15, 128>, 128, tile_smem));
            if (per_sm < 1) per_sm = 1;
            degrid_tile_kernel<15, 128><<<ctx->sm_count * per_sm, 128, tile_smem, st>>>(A);
        }
    } else {
        degrid_warp_kernel<15><<<ctx->sm_count * 8, 256, 0, st>>>(A);  // unroll 15 -> 2.9e9 vis/s, unroll 5 -> 2.1e9 (B200, S=15)
    }
    SK_LAUNCH_CHECK(ctx);
    return SKAGRID_OK;
}
