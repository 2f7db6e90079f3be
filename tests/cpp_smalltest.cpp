// cpp_smalltest.cpp -- the reference's own manual tests, replayed through the C++ host mirror (include/skagrid.hpp):
//   test/SmallTest.hs:51-76   wkerns [1,1,1,15,15] = x*(0.01+0.005i)+0.1, akerns = fill [3,15,15] 0.1, 10x10 grid, 2 visibilities,
//                             convTest = convgrid3 wkerns akerns dest uvw index vis   (the reference stores no expected output:
//                             the known answers below are the oracle's, SURVEY.md 8c / tests/golden/smalltest_oracle.npz)
//   old/BrokenNumbers.hs:85-91  double permute (+) of ten triples into a 5x5 grid -- the only expected output the reference stores
// Build (GPU box): g++ -std=c++17 -Iinclude tests/cpp_smalltest.cpp -L<pkg> -lskagrid -Wl,-rpath,<pkg> -o /tmp/cpp_smalltest
#include <cstdio>

#include "skagrid.hpp"

using namespace skagrid;

static int fail(const char *what) {
    std::fprintf(stderr, "FAILED: %s\n", what);
    return 1;
}

int main() {
    try {
        Context ctx(0);

        // ---- old/BrokenNumbers.hs
        Matrix<Visibility> a(5, 5);
        BaseLines p;
        std::vector<Visibility> val;
        for (int k = 0; k < 10; ++k) {
            p.u.push_back(((2 * k) % 5 - 2) / 5.0);      // cell = n/2 + floor(0.5 + n p)
            p.v.push_back(((3 * k + 1) % 5 - 2) / 5.0);
            p.w.push_back(0.0);
            val.emplace_back(k + 5.0, 1.0);
        }
        a = Gridding::grid(ctx, Gridding::grid(ctx, a, p, val), p, val);
        const double expect[25] = {0, 42, 0, 0, 0, 30, 0, 0, 0, 0, 0, 0, 0, 0, 38, 0, 0, 0, 46, 0, 0, 0, 34, 0, 0};
        for (int c = 0; c < 25; ++c)
            if (a.data[c] != Visibility(expect[c], expect[c] != 0 ? 4.0 : 0.0)) return fail("BrokenNumbers golden");

        // ---- test/SmallTest.hs
        NdArray<Visibility> wkerns({1, 1, 1, 15, 15}), akerns({3, 15, 15}, Visibility(0.1, 0.0));
        for (int y = 0; y < 15; ++y)
            for (int x = 0; x < 15; ++x) wkerns.data[y * 15 + x] = Visibility(0.01 * x + 0.1, 0.005 * x);
        Matrix<Visibility> dest(10, 10);
        BaseLines uvw{{0.1, -0.1}, {0.2, 0.4}, {0.3, 0.1}};
        Gridding::AwIndex index{{0, 0}, {0, 0}, {1, 2}};
        std::vector<Visibility> vis{{0.3, 0.5}, {0.4, 0.2}};

        auto [x, xf, y, yf] = Gridding::frac_coords(ctx, {10, 10}, 1, uvw);
        if (x[0] != 6 || y[0] != 7 || x[1] != 4 || y[1] != 9 || xf[0] || yf[0] || xf[1] || yf[1]) return fail("frac_coords of the SmallTest uvw");

        const Matrix<Visibility> conv3 = Gridding::convgrid3(ctx, wkerns, akerns, dest, uvw, index, vis);   // convTest
        const Matrix<Visibility> conv4 = Gridding::convgrid4(ctx, wkerns, akerns, dest, uvw, index, vis);
        Visibility sum = 0;
        double peak = 0;
        Index py = 0, px = 0;
        for (Index yy = 0; yy < 10; ++yy)
            for (Index xx = 0; xx < 10; ++xx) {
                sum += conv3(yy, xx);
                if (std::abs(conv3(yy, xx)) > peak) { peak = std::abs(conv3(yy, xx)); py = yy; px = xx; }
                if (std::abs(conv3(yy, xx) - conv4(yy, xx)) > 1e-12) return fail("convgrid3 == convgrid4");
            }
        if (std::abs(sum - Visibility(2258.959, 1709.566)) > 2e-3) return fail("sum of the SmallTest grid");
        if (py != 9 || px != 6 || std::abs(peak - 46.27454822566) > 1e-9) return fail("peak of the SmallTest grid");
        if (std::abs(conv3(7, 6) - Visibility(35.417885, 25.898745)) > 1e-5) return fail("grid[7,6]");

        // the adjoint (not in the reference): <grid(v), g> == <v, degrid(g)>
        Matrix<Visibility> g(10, 10);
        for (size_t i = 0; i < g.data.size(); ++i) g.data[i] = Visibility(std::sin(0.7 * i), std::cos(1.3 * i));
        const std::vector<Visibility> d = Gridding::convdegrid3(ctx, wkerns, akerns, g, uvw, index);
        Visibility lhs = 0, rhs = 0;
        for (size_t i = 0; i < g.data.size(); ++i) lhs += std::conj(g.data[i]) * conv3.data[i];
        for (size_t k = 0; k < vis.size(); ++k) rhs += std::conj(d[k]) * vis[k];
        if (std::abs(lhs - rhs) > 1e-10 * std::abs(lhs)) return fail("adjoint identity");

        // errors are exceptions, as `error` in the reference: antenna index out of range
        bool threw = false;
        try {
            Gridding::AwIndex bad{{0, 0}, {0, 0}, {1, 9}};
            Gridding::convgrid3(ctx, wkerns, akerns, dest, uvw, bad, vis);
        } catch (const Error &e) {
            threw = e.code == SKAGRID_ERANGE;
        }
        if (!threw) return fail("out-of-range antenna index must raise");

        // several contexts driven by this thread (here: two on device 0) give the single-context grid and visibilities
        {
            MultiContext two({0, 0});
            NdArray<Visibility> gcf({2, 2, 2, 5, 5});
            for (size_t i = 0; i < gcf.data.size(); ++i) gcf.data[i] = Visibility(std::cos(0.37 * i), std::sin(0.11 * i));
            BaseLines q;
            std::vector<Index> wb;
            std::vector<Visibility> vv;
            for (int k = 0; k < 500; ++k) {
                q.u.push_back(0.45 * std::sin(1.7 * k));
                q.v.push_back(0.45 * std::cos(0.9 * k));
                q.w.push_back(0.0);
                wb.push_back(k % 2);
                vv.emplace_back(std::sin(0.3 * k), std::cos(0.2 * k));
            }
            Matrix<Visibility> zero(48, 48);
            const Matrix<Visibility> one = Gridding::convgrid2(ctx, gcf, zero, q, wb, vv);
            const std::vector<Visibility> d1 = Gridding::convdegrid2(ctx, gcf, one, q, wb);
            const std::vector<Visibility> d1r = Gridding::convdegrid2(ctx, gcf, one, (Index)q.size());  // resident coordinates
            for (size_t k = 0; k < d1.size(); ++k)
                if (std::abs(d1[k] - d1r[k]) > 1e-12) return fail("convdegrid2 at the resident coordinates");
            for (Sharding mode : {Sharding::Visibilities, Sharding::UvTiles}) {
                const Matrix<Visibility> many = Gridding::convgrid2(two, mode, gcf, zero, q, wb, vv);
                for (size_t i = 0; i < one.data.size(); ++i)
                    if (std::abs(one.data[i] - many.data[i]) > 1e-11) return fail("multi-context convgrid2");
                const std::vector<Visibility> dm = Gridding::convdegrid2(two, mode, gcf, one, q, wb);
                for (size_t k = 0; k < d1.size(); ++k)
                    if (std::abs(d1[k] - dm[k]) > 1e-10) return fail("multi-context convdegrid2");
            }
        }

        auto [img, mx] = Gridding::grid_to_image(ctx, conv3);
        std::printf("cpp_smalltest ok: sum %.6f%+.6fi, peak %.11f at [%lld,%lld], image max %.9g\n", sum.real(), sum.imag(), peak, (long long)py,
                    (long long)px, mx);
        (void)img;
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "exception: %s\n", e.what());
        return 1;
    }
}
