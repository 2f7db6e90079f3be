// skagrid.hpp -- C++ host-side mirror of the reference's Haskell modules `Gridding` (src/Gridding.hs) and
// `ImageDataset` (src/ImageDataset.hs) on top of the C ABI (include/skagrid.h).
//
// The reference's own host language is Haskell (binding: haskell/SkaGridFFI.hs, uncompiled here -- no GHC); this header
// is the compiled-language equivalent: the SAME top-level names, argument order and meaning as the Haskell functions,
// host arrays instead of `Acc` terms, exceptions where the reference calls `error`.  Header-only; link with -lskagrid.
//
//   Haskell (src/Types.hs:7-28)            here
//   F = Double                             skagrid::F
//   Visibility = Complex Double            skagrid::Visibility  (std::complex<double>, interleaved re,im)
//   Vector BaseLines (SoA in Accelerate)   skagrid::BaseLines {u, v, w}
//   Matrix e                               skagrid::Matrix<e>   row-major [height][width], grid[y][x]
//   Kernel DIM4 / WKernels DIM5 / AKernels DIM3   skagrid::NdArray<Visibility> with explicit shape
#ifndef SKAGRID_HPP
#define SKAGRID_HPP

#include <cmath>
#include <complex>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "skagrid.h"

namespace skagrid {

using F = double;
using Visibility = std::complex<double>;
using Index = std::int64_t;

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error("libskagrid error " + std::to_string(c) + ": " + m), code(c) {}
};

template <class T>
struct NdArray {
    std::vector<Index> shape;
    std::vector<T> data;
    NdArray() = default;
    explicit NdArray(std::vector<Index> s, T fill = T()) : shape(std::move(s)) {
        Index n = 1;
        for (Index d : shape) n *= d;
        data.assign((size_t)n, fill);
    }
    Index dim(size_t i) const { return shape.at(i); }
};

template <class T>
struct Matrix {
    Index height = 0, width = 0;
    std::vector<T> data;
    Matrix() = default;
    Matrix(Index h, Index w, T fill = T()) : height(h), width(w), data((size_t)(h * w), fill) {}
    T &operator()(Index y, Index x) { return data[(size_t)(y * width + x)]; }
    const T &operator()(Index y, Index x) const { return data[(size_t)(y * width + x)]; }
};

struct BaseLines {  // Vector (F, F, F): Accelerate stores the three components as separate arrays
    std::vector<F> u, v, w;
    size_t size() const { return u.size(); }
};

struct SourceInfo {  // Vector (Antenna, Antenna, Time, Frequency)
    std::vector<Index> a1, a2;
    std::vector<F> time, frequency;
};

// One context per host thread / GPU (include/skagrid.h "Threading").  There is no CPU fallback: construction throws
// without a CUDA device.
class Context {
   public:
    explicit Context(int device = 0) {
        const int rc = skagrid_create(device, &h_);
        if (rc != SKAGRID_OK) throw Error(rc, skagrid_last_error(nullptr));
    }
    ~Context() { skagrid_destroy(h_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    skagrid_ctx *get() const { return h_; }
    void check(int rc) const {
        if (rc != SKAGRID_OK) throw Error(rc, skagrid_last_error(h_));
    }

   private:
    skagrid_ctx *h_ = nullptr;
};

// Several devices driven by this one host thread (include/skagrid.h "multi-GPU, single process"): one context per
// entry of `devices`.  Errors of any device are reported through the first context.
class MultiContext {
   public:
    explicit MultiContext(const std::vector<int> &devices) {
        if (devices.empty()) throw Error(SKAGRID_EINVAL, "MultiContext: no devices");
        for (int d : devices) {
            skagrid_ctx *h = nullptr;
            const int rc = skagrid_create(d, &h);
            if (rc != SKAGRID_OK) {
                const std::string msg = skagrid_last_error(nullptr);
                for (skagrid_ctx *c : h_) skagrid_destroy(c);
                throw Error(rc, msg);
            }
            h_.push_back(h);
        }
    }
    ~MultiContext() { for (skagrid_ctx *c : h_) skagrid_destroy(c); }
    MultiContext(const MultiContext &) = delete;
    MultiContext &operator=(const MultiContext &) = delete;
    skagrid_ctx *const *get() const { return h_.data(); }
    int size() const { return (int)h_.size(); }
    void check(int rc) const {
        if (rc != SKAGRID_OK) throw Error(rc, skagrid_last_error(h_[0]));
    }

   private:
    std::vector<skagrid_ctx *> h_;
};

enum class Sharding { Visibilities, UvTiles };  // BASELINE config 4 (shares of the visibilities + reduce) / config 5 (row slabs + routing)

inline const double *cptr(const std::vector<Visibility> &v) { return reinterpret_cast<const double *>(v.data()); }
inline double *cptr(std::vector<Visibility> &v) { return reinterpret_cast<double *>(v.data()); }

namespace Gridding {

// frac_coord (src/Gridding.hs:126-140): n qpx p -> (flx, fracx)
inline std::pair<std::vector<Index>, std::vector<Index>> frac_coord(const Context &ctx, Index n, Index qpx, const std::vector<F> &p) {
    std::vector<Index> fl(p.size()), fr(p.size());
    ctx.check(skagrid_frac_coord(ctx.get(), n, qpx, (Index)p.size(), p.data(), fl.data(), fr.data(), SKAGRID_FRAC_NORMALISE));
    return {std::move(fl), std::move(fr)};
}

// frac_coords (src/Gridding.hs:142-151): (height, width) qpx p -> (x, xf, y, yf)
inline std::tuple<std::vector<Index>, std::vector<Index>, std::vector<Index>, std::vector<Index>> frac_coords(
    const Context &ctx, std::pair<Index, Index> shape, Index qpx, const BaseLines &p) {
    const size_t n = p.size();
    std::vector<Index> x(n), xf(n), y(n), yf(n);
    ctx.check(skagrid_frac_coords(ctx.get(), shape.first, shape.second, qpx, (Index)n, p.u.data(), p.v.data(), x.data(), xf.data(), y.data(),
                                  yf.data(), SKAGRID_FRAC_NORMALISE));
    return {std::move(x), std::move(xf), std::move(y), std::move(yf)};
}

// findClosest (src/Gridding.hs:895-907), mapped over a vector of w as aw_imaging does (:473)
inline std::vector<Index> findClosest(const Context &ctx, const std::vector<F> &ws, const std::vector<F> &w) {
    std::vector<Index> out(w.size());
    ctx.check(skagrid_find_closest(ctx.get(), (Index)ws.size(), ws.data(), (Index)w.size(), w.data(), out.data()));
    return out;
}

// mirror_uvw (src/Gridding.hs:551-562)
inline std::pair<BaseLines, std::vector<Visibility>> mirror_uvw(const Context &ctx, BaseLines uvw, std::vector<Visibility> vis) {
    ctx.check(skagrid_mirror_uvw(ctx.get(), (Index)uvw.size(), uvw.u.data(), uvw.v.data(), uvw.w.data(), cptr(vis)));
    return {std::move(uvw), std::move(vis)};
}

// doweight (src/Gridding.hs:564-583): theta lam p v
inline std::vector<Visibility> doweight(const Context &ctx, F theta, Index lam, const BaseLines &p, std::vector<Visibility> v) {
    ctx.check(skagrid_doweight(ctx.get(), theta, lam, (Index)p.size(), p.u.data(), p.v.data(), cptr(v)));
    return v;
}

// grid (src/Gridding.hs:95-112): a p v
inline Matrix<Visibility> grid(const Context &ctx, Matrix<Visibility> a, const BaseLines &p, const std::vector<Visibility> &v) {
    ctx.check(skagrid_grid(ctx.get(), a.height, a.width, cptr(a.data), (Index)p.size(), p.u.data(), p.v.data(), cptr(v)));
    return a;
}

// convgrid (src/Gridding.hs:153-197): gcf[qpx,qpx,gh,gw] a p v
inline Matrix<Visibility> convgrid(const Context &ctx, const NdArray<Visibility> &gcf, Matrix<Visibility> a, const BaseLines &p,
                                   const std::vector<Visibility> &v) {
    if (gcf.shape.size() != 4 || gcf.dim(0) != gcf.dim(1)) throw Error(SKAGRID_EINVAL, "convgrid: gcf must be [qpx,qpx,gh,gw]");
    ctx.check(skagrid_convgrid(ctx.get(), gcf.dim(0), gcf.dim(2), gcf.dim(3), cptr(gcf.data), a.height, a.width, cptr(a.data), (Index)p.size(),
                               p.u.data(), p.v.data(), cptr(v)));
    return a;
}

// convgrid2 (src/Gridding.hs:199-244): gcf[nw,qpx,qpx,gh,gw] a p wbin v
inline Matrix<Visibility> convgrid2(const Context &ctx, const NdArray<Visibility> &gcf, Matrix<Visibility> a, const BaseLines &p,
                                    const std::vector<Index> &wbin, const std::vector<Visibility> &v) {
    if (gcf.shape.size() != 5 || gcf.dim(1) != gcf.dim(2)) throw Error(SKAGRID_EINVAL, "convgrid2: gcf must be [nw,qpx,qpx,gh,gw]");
    ctx.check(skagrid_convgrid2(ctx.get(), gcf.dim(0), gcf.dim(1), gcf.dim(3), gcf.dim(4), cptr(gcf.data), a.height, a.width, cptr(a.data),
                                (Index)p.size(), p.u.data(), p.v.data(), wbin.data(), cptr(v)));
    return a;
}

struct AwIndex {  // Vector (Int, Int, Int): (wbin, a1, a2)
    std::vector<Index> wbin, a1, a2;
};

// convgrid3 (src/Gridding.hs:246-317) and convgrid4 (:318-377): wkerns akerns a p index v -- same grid
inline Matrix<Visibility> convgrid3(const Context &ctx, const NdArray<Visibility> &wkerns, const NdArray<Visibility> &akerns,
                                    Matrix<Visibility> a, const BaseLines &p, const AwIndex &index, const std::vector<Visibility> &v) {
    if (wkerns.shape.size() != 5 || akerns.shape.size() != 3) throw Error(SKAGRID_EINVAL, "convgrid3: wkerns [nw,qpx,qpx,s,s], akerns [nant,s,s]");
    ctx.check(skagrid_convgrid_aw(ctx.get(), wkerns.dim(0), wkerns.dim(1), wkerns.dim(3), cptr(wkerns.data), akerns.dim(0), cptr(akerns.data),
                                  a.height, a.width, cptr(a.data), (Index)p.size(), p.u.data(), p.v.data(), index.wbin.data(), index.a1.data(),
                                  index.a2.data(), cptr(v)));
    return a;
}
inline Matrix<Visibility> convgrid4(const Context &ctx, const NdArray<Visibility> &wkerns, const NdArray<Visibility> &akerns,
                                    Matrix<Visibility> a, const BaseLines &p, const AwIndex &index, const std::vector<Visibility> &v) {
    return convgrid3(ctx, wkerns, akerns, std::move(a), p, index, v);
}

// degridding: not in the reference; exact adjoints
inline std::vector<Visibility> convdegrid2(const Context &ctx, const NdArray<Visibility> &gcf, const Matrix<Visibility> &a, const BaseLines &p,
                                           const std::vector<Index> &wbin) {
    std::vector<Visibility> out(p.size());
    ctx.check(skagrid_convdegrid2(ctx.get(), gcf.dim(0), gcf.dim(1), gcf.dim(3), gcf.dim(4), cptr(gcf.data), a.height, a.width, cptr(a.data),
                                  (Index)p.size(), p.u.data(), p.v.data(), wbin.data(), cptr(out)));
    return out;
}
// ... at the coordinates (and w-plane indices) the previous table call on ctx uploaded ("resident coordinates", include/skagrid.h)
inline std::vector<Visibility> convdegrid2(const Context &ctx, const NdArray<Visibility> &gcf, const Matrix<Visibility> &a, Index count) {
    std::vector<Visibility> out((size_t)count);
    ctx.check(skagrid_convdegrid2(ctx.get(), gcf.dim(0), gcf.dim(1), gcf.dim(3), gcf.dim(4), cptr(gcf.data), a.height, a.width, cptr(a.data), count,
                                  nullptr, nullptr, nullptr, cptr(out)));
    return out;
}
// convgrid2 / convdegrid2 over several devices
inline Matrix<Visibility> convgrid2(const MultiContext &ctxs, Sharding mode, const NdArray<Visibility> &gcf, Matrix<Visibility> a,
                                    const BaseLines &p, const std::vector<Index> &wbin, const std::vector<Visibility> &v) {
    if (gcf.shape.size() != 5 || gcf.dim(1) != gcf.dim(2)) throw Error(SKAGRID_EINVAL, "convgrid2: gcf must be [nw,qpx,qpx,gh,gw]");
    if (mode == Sharding::Visibilities)
        ctxs.check(skagrid_convgrid2_mgpu_vis(ctxs.get(), ctxs.size(), gcf.dim(0), gcf.dim(1), gcf.dim(3), gcf.dim(4), cptr(gcf.data), a.height,
                                              a.width, cptr(a.data), (Index)p.size(), p.u.data(), p.v.data(), wbin.data(), cptr(v)));
    else
        ctxs.check(skagrid_convgrid2_mgpu_tile(ctxs.get(), ctxs.size(), gcf.dim(0), gcf.dim(1), gcf.dim(3), gcf.dim(4), cptr(gcf.data), a.height,
                                               a.width, cptr(a.data), (Index)p.size(), p.u.data(), p.v.data(), wbin.data(), cptr(v), nullptr));
    return a;
}
inline std::vector<Visibility> convdegrid2(const MultiContext &ctxs, Sharding mode, const NdArray<Visibility> &gcf, const Matrix<Visibility> &a,
                                           const BaseLines &p, const std::vector<Index> &wbin) {
    std::vector<Visibility> out(p.size());
    if (mode == Sharding::Visibilities)
        ctxs.check(skagrid_convdegrid2_mgpu_vis(ctxs.get(), ctxs.size(), gcf.dim(0), gcf.dim(1), gcf.dim(3), gcf.dim(4), cptr(gcf.data), a.height,
                                                a.width, cptr(a.data), (Index)p.size(), p.u.data(), p.v.data(), wbin.data(), cptr(out)));
    else
        ctxs.check(skagrid_convdegrid2_mgpu_tile(ctxs.get(), ctxs.size(), gcf.dim(0), gcf.dim(1), gcf.dim(3), gcf.dim(4), cptr(gcf.data), a.height,
                                                 a.width, cptr(a.data), (Index)p.size(), p.u.data(), p.v.data(), wbin.data(), cptr(out), nullptr));
    return out;
}
inline std::vector<Visibility> convdegrid3(const Context &ctx, const NdArray<Visibility> &wkerns, const NdArray<Visibility> &akerns,
                                           const Matrix<Visibility> &a, const BaseLines &p, const AwIndex &index) {
    std::vector<Visibility> out(p.size());
    ctx.check(skagrid_convdegrid_aw(ctx.get(), wkerns.dim(0), wkerns.dim(1), wkerns.dim(3), cptr(wkerns.data), akerns.dim(0), cptr(akerns.data),
                                    a.height, a.width, cptr(a.data), (Index)p.size(), p.u.data(), p.v.data(), index.wbin.data(), index.a1.data(),
                                    index.a2.data(), cptr(out)));
    return out;
}

// convolve2d (src/Gridding.hs:795-811)
inline Matrix<Visibility> convolve2d(const Context &ctx, const Matrix<Visibility> &a1, const Matrix<Visibility> &a2) {
    if (a1.height != a1.width || a2.height != a1.height || a2.width != a1.width) throw Error(SKAGRID_EINVAL, "convolve2d: equal square matrices");
    Matrix<Visibility> out(a1.height, a1.width);
    ctx.check(skagrid_convolve2d(ctx.get(), a1.height, cptr(a1.data), cptr(a2.data), cptr(out.data)));
    return out;
}

// make_grid_hermitian (src/Gridding.hs:585-605), ifft (:828-829), fft (:821-826)
inline Matrix<Visibility> make_grid_hermitian(const Context &ctx, const Matrix<Visibility> &g) {
    Matrix<Visibility> out(g.height, g.width);
    ctx.check(skagrid_make_grid_hermitian(ctx.get(), g.height, cptr(g.data), cptr(out.data)));
    return out;
}
inline Matrix<Visibility> ifft(const Context &ctx, const Matrix<Visibility> &g) {
    Matrix<Visibility> out(g.height, g.width);
    ctx.check(skagrid_ifft(ctx.get(), g.height, cptr(g.data), cptr(out.data)));
    return out;
}
inline Matrix<Visibility> fft(const Context &ctx, const Matrix<Visibility> &g) {
    Matrix<Visibility> out(g.height, g.width);
    ctx.check(skagrid_fft(ctx.get(), g.height, cptr(g.data), cptr(out.data)));
    return out;
}

inline Index grid_side(F theta, Index lam) { return (Index)skagrid_grid_side(theta, lam); }  // P.round: half to even

// simple_imaging (src/Gridding.hs:84-93), conv_imaging (:115-124), aw_imaging (:452-478): ImagingFunction argument order
inline Matrix<Visibility> simple_imaging(const Context &ctx, F theta, Index lam, const BaseLines &uvw, const SourceInfo &, const std::vector<Visibility> &vis) {
    const Index n = grid_side(theta, lam);
    Matrix<Visibility> out(n, n);
    ctx.check(skagrid_simple_imaging(ctx.get(), theta, lam, (Index)uvw.size(), uvw.u.data(), uvw.v.data(), uvw.w.data(), cptr(vis), cptr(out.data)));
    return out;
}
inline Matrix<Visibility> conv_imaging(const Context &ctx, const NdArray<Visibility> &kv, F theta, Index lam, const BaseLines &uvw,
                                       const SourceInfo &, const std::vector<Visibility> &vis) {
    const Index n = grid_side(theta, lam);
    Matrix<Visibility> out(n, n);
    ctx.check(skagrid_conv_imaging(ctx.get(), kv.dim(0), kv.dim(2), kv.dim(3), cptr(kv.data), theta, lam, (Index)uvw.size(), uvw.u.data(),
                                   uvw.v.data(), uvw.w.data(), cptr(vis), cptr(out.data)));
    return out;
}
inline Matrix<Visibility> aw_imaging(const Context &ctx, F theta, Index lam, const NdArray<Visibility> &wkernels, const std::vector<F> &wbins,
                                     const NdArray<Visibility> &akernels, const BaseLines &uvw, const SourceInfo &src,
                                     const std::vector<Visibility> &vis) {
    const Index n = grid_side(theta, lam);
    Matrix<Visibility> out(n, n);
    ctx.check(skagrid_aw_imaging(ctx.get(), theta, lam, wkernels.dim(0), wkernels.dim(1), wkernels.dim(3), cptr(wkernels.data), wbins.data(),
                                 akernels.dim(0), cptr(akernels.data), (Index)uvw.size(), uvw.u.data(), uvw.v.data(), uvw.w.data(), src.a1.data(),
                                 src.a2.data(), cptr(vis), cptr(out.data)));
    return out;
}

// KernelOptions (src/Gridding.hs:30-38); a negative size means Nothing
struct KernelOptions {
    Index patHorShift = 0, patVerShift = 0;
    std::vector<F> patTransMat;   // empty = Nothing, else 2 x 2 row-major
    Index wstep = -1, qpx = -1, npixFF = -1, npixKern = -1;
};

// w_kernel (src/Gridding.hs:610-619) for a list of w -> [nw, qpx, qpx, npixKern, npixKern]; conjugate as w_cache_imaging does (:441)
inline NdArray<Visibility> w_kernel(const Context &ctx, F theta, const std::vector<F> &w, const KernelOptions &k, bool conjugate = false) {
    if (k.qpx <= 0 || k.npixFF <= 0 || k.npixKern <= 0) throw Error(SKAGRID_EINVAL, "w_kernel: qpx, npixFF and npixKern must be given");
    if (!k.patTransMat.empty() && k.patTransMat.size() != 4) throw Error(SKAGRID_EINVAL, "w_kernel: patTransMat must be 2 x 2");
    NdArray<Visibility> out({(Index)w.size(), k.qpx, k.qpx, k.npixKern, k.npixKern});
    ctx.check(skagrid_w_kernels_ex(ctx.get(), theta, (Index)w.size(), w.data(), k.npixFF, k.npixKern, k.qpx, conjugate ? 1 : 0,
                                   k.patTransMat.empty() ? nullptr : k.patTransMat.data(), (F)k.patHorShift, (F)k.patVerShift, cptr(out.data)));
    return out;
}

// map real . ifft . make_grid_hermitian and its maximum, fused (src/ImageDataset.hs:74-77)
inline std::pair<Matrix<F>, F> grid_to_image(const Context &ctx, const Matrix<Visibility> &g) {
    Matrix<F> img(g.height, g.width);
    F mx = 0;
    ctx.check(skagrid_grid_to_image(ctx.get(), g.height, cptr(g.data), img.data.data(), &mx));
    return {std::move(img), mx};
}

}  // namespace Gridding

namespace ImageDataset {

// uvw_lambda (src/ImageDataset.hs:181-187)
inline BaseLines uvw_lambda(const Context &ctx, F f, BaseLines uvw) {
    ctx.check(skagrid_uvw_lambda(ctx.get(), f, (Index)uvw.size(), uvw.u.data(), uvw.v.data(), uvw.w.data()));
    return uvw;
}

// aw_gridding from the loaded arrays on (src/ImageDataset.hs:47-77): uvw in metres -> (image, maximum)
inline std::pair<Matrix<F>, F> aw_gridding(const Context &ctx, F theta, Index lam, const NdArray<Visibility> &wkernels, const std::vector<F> &wbins,
                                           const NdArray<Visibility> &akernels, const BaseLines &uvw_m, const SourceInfo &src, F freq,
                                           const std::vector<Visibility> &vis) {
    const Index n = Gridding::grid_side(theta, lam);
    Matrix<F> img(n, n);
    F mx = 0;
    ctx.check(skagrid_aw_gridding(ctx.get(), theta, lam, wkernels.dim(0), wkernels.dim(1), wkernels.dim(3), cptr(wkernels.data), wbins.data(),
                                  akernels.dim(0), cptr(akernels.data), (Index)uvw_m.size(), uvw_m.u.data(), uvw_m.v.data(), uvw_m.w.data(),
                                  src.a1.data(), src.a2.data(), freq, cptr(vis), img.data.data(), &mx, nullptr));
    return {std::move(img), mx};
}

}  // namespace ImageDataset
}  // namespace skagrid
#endif  // SKAGRID_HPP
