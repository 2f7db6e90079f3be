/* c_abi_smoke.c -- the drop-in boundary exercised from plain C (no Python, no torch): this is what a
 * `foreign import ccall` binding sees.  Replays old/BrokenNumbers.hs:85-91 through skagrid_grid (the permute (+)
 * golden), checks frac_coords / convgrid2 / convdegrid2 on a tiny hand-checkable case and the adjoint identity.
 * Build + run (GPU box):  gcc -std=c99 -Iinclude tests/c_abi_smoke.c -L<pkg> -lskagrid -lm -Wl,-rpath,<pkg> -o /tmp/c_abi_smoke */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "skagrid.h"

#define CHECK(rc, what)                                                                      \
    do {                                                                                     \
        if ((rc) != 0) {                                                                     \
            fprintf(stderr, "%s failed (%d): %s\n", what, (rc), skagrid_last_error(ctx));    \
            return 1;                                                                        \
        }                                                                                    \
    } while (0)

int main(void) {
    skagrid_ctx *ctx = NULL;
    int rc = skagrid_create(0, &ctx);
    if (rc != 0) { fprintf(stderr, "skagrid_create: %s\n", skagrid_last_error(NULL)); return 1; }

    /* 1. BrokenNumbers golden: two applications of the scatter-add of [((2k)%5, (3k+1)%5, (k+5)+1i) | k<10] */
    double grid[5 * 5 * 2];
    memset(grid, 0, sizeof grid);
    double pu[10], pv[10], val[20];
    for (int k = 0; k < 10; ++k) {
        pu[k] = ((2 * k) % 5 - 2) / 5.0;      /* cell = n/2 + floor(0.5 + n p)  (src/Gridding.hs:95-112) */
        pv[k] = ((3 * k + 1) % 5 - 2) / 5.0;
        val[2 * k] = k + 5.0; val[2 * k + 1] = 1.0;
    }
    for (int rep = 0; rep < 2; ++rep) CHECK(skagrid_grid(ctx, 5, 5, grid, 10, pu, pv, val), "skagrid_grid");
    const double exp_re[25] = {0, 42, 0, 0, 0, 30, 0, 0, 0, 0, 0, 0, 0, 0, 38, 0, 0, 0, 46, 0, 0, 0, 34, 0, 0};
    for (int c = 0; c < 25; ++c) {
        const double eim = exp_re[c] != 0 ? 4.0 : 0.0;
        if (grid[2 * c] != exp_re[c] || grid[2 * c + 1] != eim) { fprintf(stderr, "BrokenNumbers golden mismatch at cell %d\n", c); return 1; }
    }

    /* 2. one visibility, 3x3 kernel of ones*(1+2i), qpx 1, 8x8 grid: p=(0.125,-0.25) -> x = 4+1 = 5, y = 4-2 = 2 */
    const double u[1] = {0.125}, v[1] = {-0.25}, vis[2] = {2.0, -1.0};
    const int64_t wb[1] = {0};
    int64_t x, xf, y, yf;
    CHECK(skagrid_frac_coords(ctx, 8, 8, 1, 1, u, v, &x, &xf, &y, &yf, SKAGRID_FRAC_NORMALISE), "skagrid_frac_coords");
    if (x != 5 || y != 2 || xf != 0 || yf != 0) { fprintf(stderr, "frac_coords: got (%lld,%lld,%lld,%lld)\n", (long long)x, (long long)xf, (long long)y, (long long)yf); return 1; }
    double gcf[9 * 2], g8[8 * 8 * 2];
    for (int t = 0; t < 9; ++t) { gcf[2 * t] = 1.0; gcf[2 * t + 1] = 2.0; }
    memset(g8, 0, sizeof g8);
    CHECK(skagrid_convgrid2(ctx, 1, 1, 3, 3, gcf, 8, 8, g8, 1, u, v, wb, vis), "skagrid_convgrid2");
    /* (2 - i)(1 + 2i) = 4 + 3i on rows 1..3, cols 4..6 */
    for (int r = 0; r < 8; ++r)
        for (int c = 0; c < 8; ++c) {
            const int in = r >= 1 && r <= 3 && c >= 4 && c <= 6;
            if (fabs(g8[2 * (r * 8 + c)] - (in ? 4.0 : 0.0)) > 1e-14 || fabs(g8[2 * (r * 8 + c) + 1] - (in ? 3.0 : 0.0)) > 1e-14) {
                fprintf(stderr, "convgrid2 mismatch at (%d,%d)\n", r, c); return 1;
            }
        }
    /* 3. adjoint: vis' = sum conj(k) g = 9 * (1 - 2i)(4 + 3i) = 9 * (10 - 5i) */
    double out[2];
    CHECK(skagrid_convdegrid2(ctx, 1, 1, 3, 3, gcf, 8, 8, g8, 1, u, v, wb, out), "skagrid_convdegrid2");
    if (fabs(out[0] - 90.0) > 1e-12 || fabs(out[1] + 45.0) > 1e-12) { fprintf(stderr, "convdegrid2: got %g%+gi\n", out[0], out[1]); return 1; }
    /* 4. error path: w-plane index out of range is an error code, not UB */
    const int64_t bad[1] = {7};
    rc = skagrid_convgrid2(ctx, 1, 1, 3, 3, gcf, 8, 8, g8, 1, u, v, bad, vis);
    if (rc != SKAGRID_ERANGE) { fprintf(stderr, "expected SKAGRID_ERANGE, got %d\n", rc); return 1; }
    /* 5. two contexts driven by this thread (both on device 0 here): same grid, same visibility */
    {
        skagrid_ctx *two[2] = {ctx, NULL};
        if (skagrid_create(0, &two[1]) != 0) { fprintf(stderr, "second skagrid_create: %s\n", skagrid_last_error(NULL)); return 1; }
        double g2[8 * 8 * 2], out2[2];
        int64_t bounds[3];
        for (int tile = 0; tile < 2; ++tile) {
            memset(g2, 0, sizeof g2);
            if (tile) { CHECK(skagrid_convgrid2_mgpu_tile(two, 2, 1, 1, 3, 3, gcf, 8, 8, g2, 1, u, v, wb, vis, bounds), "skagrid_convgrid2_mgpu_tile"); }
            else { CHECK(skagrid_convgrid2_mgpu_vis(two, 2, 1, 1, 3, 3, gcf, 8, 8, g2, 1, u, v, wb, vis), "skagrid_convgrid2_mgpu_vis"); }
            for (int t = 0; t < 8 * 8 * 2; ++t)
                if (fabs(g2[t] - g8[t]) > 1e-14) { fprintf(stderr, "mgpu grid mismatch (tile=%d) at %d\n", tile, t); return 1; }
            if (tile) { CHECK(skagrid_convdegrid2_mgpu_tile(two, 2, 1, 1, 3, 3, gcf, 8, 8, g8, 1, u, v, wb, out2, bounds), "skagrid_convdegrid2_mgpu_tile"); }
            else { CHECK(skagrid_convdegrid2_mgpu_vis(two, 2, 1, 1, 3, 3, gcf, 8, 8, g8, 1, u, v, wb, out2), "skagrid_convdegrid2_mgpu_vis"); }
            if (fabs(out2[0] - 90.0) > 1e-12 || fabs(out2[1] + 45.0) > 1e-12) { fprintf(stderr, "mgpu degrid (tile=%d): got %g%+gi\n", tile, out2[0], out2[1]); return 1; }
        }
        if (bounds[0] != 0 || bounds[2] != 8 || bounds[1] <= 0 || bounds[1] >= 8) { fprintf(stderr, "mgpu tile bounds %lld %lld %lld\n", (long long)bounds[0], (long long)bounds[1], (long long)bounds[2]); return 1; }
        skagrid_destroy(two[1]);
    }
    double mx = 0.0;
    CHECK(skagrid_grid_to_image(ctx, 8, g8, NULL, &mx), "skagrid_grid_to_image");
    printf("c_abi_smoke ok (%s, %lld kernel launches, image max %.6g)\n", skagrid_version(), (long long)skagrid_launch_count(ctx), mx);
    skagrid_destroy(ctx);
    return 0;
}
