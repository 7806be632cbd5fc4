/*
 * ref_driver.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Thin extern "C" shim (ours) over the REAL reference classes, compiled by
 * oracle/Makefile from the sources where they lie under /root/reference into
 * oracle/_ref/libfracref_{nofma,fma}.so.  It exports the interface of
 * frac_oracle.h with the prefix `fr_` so the same test harness can drive the
 * restatement (fo_) and the reference (fr_).  No reference source is copied;
 * this file only #includes reference headers at build time.
 *
 * Threading: the reference's own EncodingEngineCore2::encode has a lost-wakeup
 * deadlock, so the per-range work -- the reference's
 * TransformEstimator2::estimate, unmodified -- is fanned out with OpenMP here.
 */
#include <cassert>
#include <cmath>
#include <cstdint>
#include <mutex>
#include <condition_variable>

#include "encode/Classifier2.hpp"
#include "encode/DecodeUtils.hpp"
#include "encode/Encoder2.hpp"
#include "encode/Quantizer.hpp"
#include "encode/TransformEstimator2.hpp"
#include "encode/transformmatcher.h"
#include "image/Image2.hpp"
#include "image/ImageIO.hpp"
#include "image/ImageStatistics.hpp"
#include "image/metrics.h"
#include "image/partition2.hpp"
#include "image/sampler.h"

#include "frac_oracle.h"

#include <cstring>
#include <memory>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace Frac2;

namespace {

ImagePlane make_plane(const fo_plane* p) {
    std::vector<uint8_t> bytes(p->px, p->px + (size_t)p->height * p->stride);
    return ImagePlane(Size32u(p->width, p->height), p->stride, std::move(bytes));
}

UniformGridItem make_item(const fo_grid_item& it) {
    GridItemData d;
    d.bb_classifierBin = it.bin;
    return UniformGridItem(Point2du(it.x, it.y), Size32u(it.w, it.h), std::move(d));
}

UniformGrid make_grid(const fo_grid_item* items, size_t n) {
    UniformGrid g;
    g.reserve(n);
    for (size_t i = 0; i < n; ++i) {
        GridItemData d;
        d.bb_classifierBin = items[i].bin;
        g.add(Point2du(items[i].x, items[i].y), Size32u(items[i].w, items[i].h), std::move(d));
    }
    return g;
}

void store(fo_encode_item* out, const UniformGridItem& r, const item_match_t& m) {
    std::memset(out, 0, sizeof(*out));
    out->x = r.origin.x();
    out->y = r.origin.y();
    out->w = r.size.x();
    out->h = r.size.y();
    out->distance = m.score.distance;
    out->contrast = m.score.contrast;
    out->brightness = m.score.brightness;
    out->transform = static_cast<int32_t>(m.score.transform);
    out->match_x = m.x;
    out->match_y = m.y;
    out->src_w = m.sourceItemSize.x();
    out->src_h = m.sourceItemSize.y();
}

std::unique_ptr<Classifier2> make_classifier(int use, const ImagePlane& s, const ImagePlane& t) {
    if (use) return std::make_unique<BrightnessBlocksClassifier2>(s, t);
    return std::make_unique<DummyClassifier>(s, t);
}

template <TransformType T>
double dist(const ImagePlane& a, const ImagePlane& b, const GridItemBase& sa, const GridItemBase& sb) {
    return Frac::RootMeanSquare<T>().distance(a, b, sa, sb);
}

void encode_level(const ImagePlane& src, const ImagePlane& tgt, const UniformGrid& dom,
                  const fo_grid_item* ranges, size_t n_rng, const fo_params* p, int nthreads,
                  size_t sample_stride, fo_encode_item* out) {
    TransformEstimator2 est(src, tgt, make_classifier(p->use_classifier, src, tgt),
                            std::make_shared<TransformMatcher>(p->rms_threshold, p->s_max), dom);
    if (sample_stride == 0) sample_stride = 1;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_num_procs();
#endif
    (void)nthreads;
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
    for (long long i = 0; i < (long long)n_rng; ++i) {
        if ((size_t)i % sample_stride) {
            std::memset(&out[i], 0, sizeof(out[i]));
            continue;
        }
        const UniformGridItem r = make_item(ranges[i]);
        store(&out[i], r, est.estimate(r));
    }
}

} // namespace

extern "C" {

const char* fr_version(void) {
#ifdef __FMA__
    return "sebsgit/fractencode reference, compiled from /root/reference (FMA contraction on)";
#else
    return "sebsgit/fractencode reference, compiled from /root/reference (no FMA)";
#endif
}

int fr_hardware_threads(void) {
#ifdef _OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}

int fr_sample_sum4(const fo_plane* img, uint32_t px, uint32_t py, uint32_t pw, uint32_t ph,
                   uint32_t lx, uint32_t ly, int transform) {
    const ImagePlane plane = make_plane(img);
    const GridItemBase patch{Point2du(px, py), Size32u(pw, ph)};
    const float v = Frac::SamplerBilinear::sample<float>(plane, patch, lx, ly, static_cast<TransformType>(transform));
    return (int)(v * 4.0f);
}

uint32_t fr_block_sum(const fo_plane* img, uint32_t x, uint32_t y, uint32_t w, uint32_t h) {
    const ImagePlane plane = make_plane(img);
    return (uint32_t)ImageStatistics2::sum<double>(plane, GridItemBase{Point2du(x, y), Size32u(w, h)});
}

double fr_distance(const fo_plane* a, const fo_plane* b, uint32_t ax, uint32_t ay, uint32_t aw,
                   uint32_t ah, uint32_t bx, uint32_t by, uint32_t bw, uint32_t bh, int t) {
    const ImagePlane pa = make_plane(a), pb = make_plane(b);
    const GridItemBase sa{Point2du(ax, ay), Size32u(aw, ah)}, sb{Point2du(bx, by), Size32u(bw, bh)};
    switch (static_cast<TransformType>(t)) {
    case TransformType::Id: return dist<TransformType::Id>(pa, pb, sa, sb);
    case TransformType::Rotate_90: return dist<TransformType::Rotate_90>(pa, pb, sa, sb);
    case TransformType::Rotate_180: return dist<TransformType::Rotate_180>(pa, pb, sa, sb);
    case TransformType::Rotate_270: return dist<TransformType::Rotate_270>(pa, pb, sa, sb);
    case TransformType::Flip: return dist<TransformType::Flip>(pa, pb, sa, sb);
    case TransformType::Flip_Rotate_90: return dist<TransformType::Flip_Rotate_90>(pa, pb, sa, sb);
    case TransformType::Flip_Rotate_180: return dist<TransformType::Flip_Rotate_180>(pa, pb, sa, sb);
    case TransformType::Flip_Rotate_270: return dist<TransformType::Flip_Rotate_270>(pa, pb, sa, sb);
    }
    return -1.0;
}

int fr_category(const fo_plane* img, uint32_t x, uint32_t y, uint32_t w, uint32_t h) {
    const ImagePlane plane = make_plane(img);
    return BrightnessBlocksClassifier2::getCategory(plane, UniformGridItem(Point2du(x, y), Size32u(w, h)));
}

/* Bulk variant (one plane copy): categories of n items. */
void fr_preclassify(const fo_plane* img, fo_grid_item* items, size_t n) {
    const ImagePlane plane = make_plane(img);
    const BrightnessBlocksClassifier2 cls(plane, plane);
    for (size_t i = 0; i < n; ++i) {
        GridItemData d;
        cls.preclassify(Point2du(items[i].x, items[i].y), Size32u(items[i].w, items[i].h), d);
        items[i].bin = d.bb_classifierBin;
    }
}

/* The 4-double overload is private; reach it through a 2x2 image whose
 * quadrants are the single pixels a1..a4 (valid for integers 0..255). */
int fr_category4(double a1, double a2, double a3, double a4) {
    ImagePlane plane(Size32u(2, 2), 2, std::vector<uint8_t>{(uint8_t)a1, (uint8_t)a2, (uint8_t)a3, (uint8_t)a4});
    return BrightnessBlocksClassifier2::getCategory(plane, UniformGridItem(Point2du(0, 0), Size32u(2, 2)));
}

size_t fr_create_uniform_grid(uint32_t W, uint32_t H, uint32_t sx, uint32_t sy, uint32_t ox,
                              uint32_t oy, fo_grid_item* out, size_t cap) {
    if (!sx || !sy || !ox || !oy || W % sx || H % sy || W % ox || H % oy) return 0; /* would FRAC_ASSERT+exit */
    const auto g = createUniformGrid(Size32u(W, H), Size32u(sx, sy), Size32u(ox, oy));
    size_t n = 0;
    for (const auto& it : g.items()) {
        if (n < cap && out) out[n] = fo_grid_item{it.origin.x(), it.origin.y(), it.size.x(), it.size.y(), it.data.bb_classifierBin};
        ++n;
    }
    return n;
}

void fr_match(const fo_plane* src, const fo_grid_item* dom, const fo_plane* tgt,
              const fo_grid_item* rng, const fo_params* p, fo_encode_item* out) {
    const ImagePlane ps = make_plane(src), pt = make_plane(tgt);
    const Frac::TransformMatcher m(p->rms_threshold, p->s_max);
    const auto s = m.match(ps, make_item(*dom), pt, make_item(*rng));
    std::memset(out, 0, sizeof(*out));
    out->distance = s.distance;
    out->contrast = s.contrast;
    out->brightness = s.brightness;
    out->transform = static_cast<int32_t>(s.transform);
}

void fr_encode_level(const fo_plane* src, const fo_plane* tgt, const fo_grid_item* domains,
                     size_t n_dom, const fo_grid_item* ranges, size_t n_rng, const fo_params* p,
                     int nthreads, size_t sample_stride, fo_encode_item* out) {
    const ImagePlane ps = make_plane(src), pt = make_plane(tgt);
    const UniformGrid dom = make_grid(domains, n_dom);
    encode_level(ps, pt, dom, ranges, n_rng, p, nthreads, sample_stride, out);
}

void fr_estimate(const fo_plane* src, const fo_plane* tgt, const fo_grid_item* domains, size_t n_dom,
                 const fo_grid_item* rng, const fo_params* p, fo_encode_item* out) {
    fr_encode_level(src, tgt, domains, n_dom, rng, 1, p, 1, 1, out);
}

/* Quadtree = composition of reference parts: createUniformGrid,
 * preclassify, TransformEstimator2::estimate, checkDistance, topLeft..bottomRight. */
size_t fr_encode_quadtree(const fo_plane* img, uint32_t t_max, uint32_t t_min, const fo_params* p,
                          int nthreads, fo_encode_item* out, size_t cap, size_t* level_counts) {
    const ImagePlane plane = make_plane(img);
    const Size32u isz(img->width, img->height);
    if (!t_max || isz.x() % t_max || isz.y() % t_max) return (size_t)-1;
    const Frac::TransformMatcher matcher(p->rms_threshold, p->s_max);
    auto classifier = make_classifier(p->use_classifier, plane, plane);
    auto cb = [&](const Point2du& o, const Size32u& s) {
        GridItemData d;
        classifier->preclassify(o, s, d);
        return d;
    };
    std::vector<fo_grid_item> pending;
    const UniformGrid top = createUniformGrid(isz, Size32u(t_max, t_max), Size32u(t_max, t_max));
    for (const auto& it : top.items())
        pending.push_back(fo_grid_item{it.origin.x(), it.origin.y(), it.size.x(), it.size.y(), -1});
    size_t n_out = 0;
    int level = 0;
    for (uint32_t T = t_max; T >= t_min && !pending.empty(); T /= 2, ++level) {
        const UniformGrid dom = createUniformGrid(isz, Size32u(2 * T, 2 * T), Size32u(T, T), cb);
        for (auto& r : pending) {
            GridItemData d;
            classifier->preclassify(Point2du(r.x, r.y), Size32u(r.w, r.h), d);
            r.bin = d.bb_classifierBin;
        }
        std::vector<fo_encode_item> res(pending.size());
        encode_level(plane, plane, dom, pending.data(), pending.size(), p, nthreads, 1, res.data());
        std::vector<fo_grid_item> next;
        size_t emitted = 0;
        for (size_t i = 0; i < pending.size(); ++i) {
            if (matcher.checkDistance(res[i].distance) || T / 2 < t_min) {
                if (n_out >= cap) return (size_t)-1;
                out[n_out++] = res[i];
                ++emitted;
            } else {
                const GridItemBase b{Point2du(pending[i].x, pending[i].y), Size32u(T, T)};
                for (const GridItemBase& c : {b.topLeft(), b.topRight(), b.bottomLeft(), b.bottomRight()})
                    next.push_back(fo_grid_item{c.origin.x(), c.origin.y(), c.size.x(), c.size.y(), -1});
            }
        }
        if (level_counts) level_counts[level] = emitted;
        pending.swap(next);
    }
    return n_out;
}

void fr_decode(const fo_encode_item* items, size_t n, uint8_t* target, uint32_t width,
               uint32_t height, uint32_t stride, int max_iters, double rms_eps, int /*fma: fixed at build*/,
               int* iterations_out, double* rms_out) {
    Frac::grid_encode_data_t data;
    data.encoded.reserve(n);
    for (size_t i = 0; i < n; ++i) {
        if (!items[i].src_w || !items[i].src_h) continue; /* would FRAC_ASSERT+exit in the sampler */
        Frac::encode_item_t e;
        e.x = items[i].x;
        e.y = items[i].y;
        e.w = items[i].w;
        e.h = items[i].h;
        e.match.score.distance = items[i].distance;
        e.match.score.contrast = items[i].contrast;
        e.match.score.brightness = items[i].brightness;
        e.match.score.transform = static_cast<TransformType>(items[i].transform);
        e.match.x = items[i].match_x;
        e.match.y = items[i].match_y;
        e.match.sourceItemSize = Size32u(items[i].src_w, items[i].src_h);
        data.encoded.push_back(e);
    }
    std::vector<uint8_t> bytes(target, target + (size_t)height * stride);
    ImagePlane plane(Size32u(width, height), stride, std::move(bytes));
    Decoder2 dec(plane, max_iters, rms_eps, false);
    const auto st = dec.decode(data);
    std::memcpy(target, plane.data(), (size_t)height * stride);
    if (iterations_out) *iterations_out = st.iterations;
    if (rms_out) *rms_out = st.rms;
}

uint64_t fr_quantize(double v, double vmin, double vmax, int bits) {
    return Frac::Quantizerd(vmin, vmax, bits).quantized(v);
}

double fr_dequantize(uint64_t q, double vmin, double vmax, int bits) {
    return Frac::Quantizerd(vmin, vmax, bits).value(q);
}

/* ImageIO::loadImage -> luma plane, tightly packed into out (w*h bytes). */
int fr_load_luma(const char* path, uint8_t* out, size_t cap, uint32_t* w, uint32_t* h) {
    auto planes = ImageIO::loadImage(path);
    const ImagePlane& y = planes[0];
    *w = y.width();
    *h = y.height();
    if ((size_t)y.width() * y.height() > cap) return -1;
    for (uint32_t r = 0; r < y.height(); ++r)
        std::memcpy(out + (size_t)r * y.width(), y.data() + (size_t)r * y.stride(), y.width());
    return 0;
}

/* ImageIO::rgb2yuv / yuv2rgb (image/ImageIO.cpp:40-84) on caller buffers: the colour path of main.cpp:193-200. */
void fr_rgb2yuv(const uint8_t* rgb, uint32_t w, uint32_t h, uint32_t stride, uint8_t* y, uint32_t ys, uint8_t* u, uint32_t us, uint8_t* v,
                uint32_t vs, int fma) {
    (void)fma; /* decided by this library's build flags */
    ImageIO::rgb2yuv({rgb, (std::ptrdiff_t)((size_t)stride * h)}, w, h, stride, {y, (std::ptrdiff_t)((size_t)ys * h)}, ys,
                     {u, (std::ptrdiff_t)((size_t)us * ((h + 1) / 2))}, us, {v, (std::ptrdiff_t)((size_t)vs * ((h + 1) / 2))}, vs);
}
void fr_yuv2rgb(const uint8_t* y, uint32_t w, uint32_t h, uint32_t ys, const uint8_t* u, uint32_t us, const uint8_t* v, uint32_t vs,
                uint8_t* rgb, uint32_t rgb_stride, int fma) {
    (void)fma;
    ImageIO::yuv2rgb({y, (std::ptrdiff_t)((size_t)ys * h)}, w, h, ys, {u, (std::ptrdiff_t)((size_t)us * ((h + 1) / 2))}, us,
                     {v, (std::ptrdiff_t)((size_t)vs * ((h + 1) / 2))}, vs, {rgb, (std::ptrdiff_t)((size_t)rgb_stride * 3 * h)}, rgb_stride);
}

} // extern "C"
