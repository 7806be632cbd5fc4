/*
 * frac_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see frac_oracle.h).
 *
 * Literal CPU restatement of the reference search path.  Deliberately written
 * in the reference's own (slow, per-pixel) structure so that it is an
 * independent check of the "net rules" the CUDA path implements.  Compile
 * with -ffp-contract=off: fused multiply-adds happen only where `fma` asks.
 *
 * Parity: PINNED -- tests/test_oracle_kat.py (the reference's own KATs),
 * tests/test_oracle_vs_ref.py (against oracle/_ref, the compiled reference)
 * and tests/test_oracle_golden.py (committed fixtures made by oracle/_ref).
 */
#include "frac_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* image/transform.h:32-41 (__map_lookup). */
static const int kMap[8][8] = {
    {1, 0, 0, 0, 0, 1, 0, 0},   /* Id */
    {0, 1, 0, 0, -1, 0, 1, 0},  /* Rotate_90 */
    {-1, 0, 1, 0, 0, -1, 0, 1}, /* Rotate_180 */
    {0, -1, 0, 1, 1, 0, 0, 0},  /* Rotate_270 */
    {1, 0, 0, 0, 0, -1, 0, 1},  /* Flip */
    {0, 1, 0, 0, 1, 0, 0, 0},   /* Flip_Rotate_90 */
    {-1, 0, 1, 0, 0, 1, 0, 0},  /* Flip_Rotate_180 */
    {0, -1, 0, 1, -1, 0, 1, 0}, /* Flip_Rotate_270 */
};

const char* fo_version(void) { return "frac_oracle 1 (C restatement of sebsgit/fractencode encode path)"; }

int fo_hardware_threads(void) {
#ifdef _OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}

/* ImagePlane::value, image/Image2.hpp:107-111. */
static inline int px_at(const fo_plane* img, int64_t x, int64_t y) {
    return img->px[y * (int64_t)img->stride + x];
}

/* SamplerBilinear::sample (image/sampler.h:22-38): edge decrement, then
 * Transform::generateSampleOffsets (image/transform.h:96-109) and
 * ImagePlane::sumAt (image/Image2.hpp:126-131). */
int fo_sample_sum4(const fo_plane* img, uint32_t px, uint32_t py, uint32_t pw, uint32_t ph,
                   uint32_t lx, uint32_t ly, int t) {
    const int* m = kMap[t];
    if (lx == pw - 1) --lx; /* sampler.h:32-33 */
    if (ly == ph - 1) --ly; /* sampler.h:34-35 */
    const int64_t s = img->stride;
    const int64_t sx1 = (int64_t)pw - 1, sy1 = (int64_t)ph - 1;
    const int64_t off_x = (int64_t)px + m[0] * (int64_t)lx + m[1] * (int64_t)ly + m[2] * sx1 + m[3] * sy1;
    const int64_t row = s * ((int64_t)py + m[4] * (int64_t)lx + m[5] * (int64_t)ly + m[6] * sx1 + m[7] * sy1);
    const int64_t p0 = row + off_x;                                  /* transform.h:103 */
    const int64_t p1 = m[4] * s + row + m[0] + off_x;                /* :104 */
    const int64_t p2 = m[5] * s + row + m[1] + off_x;                /* :105 */
    const int64_t p3 = (m[4] + m[5]) * s + row + m[0] + m[1] + off_x; /* :106 */
    return img->px[p0] + img->px[p1] + img->px[p2] + img->px[p3];
}

/* ImageStatistics2::sum (image/ImageStatistics.hpp:13-17): sum_u16 for widths
 * <= 16 (max 65280, no wrap), sum_u32 otherwise; the SSE special cases of
 * ImageStatistics.cpp:15-41 are plain sums. */
uint32_t fo_block_sum(const fo_plane* img, uint32_t x, uint32_t y, uint32_t w, uint32_t h) {
    uint32_t r = 0;
    for (uint32_t j = 0; j < h; ++j)
        for (uint32_t i = 0; i < w; ++i) r += (uint32_t)px_at(img, x + i, y + j);
    if (w <= 16) r = (uint16_t)r; /* sum_u16 returns uint16_t */
    return r;
}

/* RootMeanSquare<T>::distance, image/metrics.h:21-51. */
double fo_distance(const fo_plane* a, const fo_plane* b, uint32_t ax, uint32_t ay, uint32_t aw,
                   uint32_t ah, uint32_t bx, uint32_t by, uint32_t bw, uint32_t bh, int t) {
    const int* m = kMap[t];
    if (aw == bw && ah == bh) { /* metrics.h:26-36: int32 accumulator (wraps) */
        uint32_t sum = 0;
        for (uint32_t y = 0; y < bh; ++y)
            for (uint32_t x = 0; x < bw; ++x) {
                const int vb = px_at(b, bx + x, by + y);
                /* Transform::map(x,y,ox,oy,w,h), transform.h:74-87 */
                const int64_t mx = m[0] * (int64_t)x + m[1] * (int64_t)y + m[2] * ((int64_t)aw - 1) + m[3] * ((int64_t)ah - 1);
                const int64_t my = m[4] * (int64_t)x + m[5] * (int64_t)y + m[6] * ((int64_t)aw - 1) + m[7] * ((int64_t)ah - 1);
                const int v = px_at(a, mx + ax, my + ay) - vb;
                sum += (uint32_t)(v * v);
            }
        return (double)(int32_t)sum / (double)(aw * ah);
    }
    float sum = 0.0f; /* metrics.h:38-49 */
    const int16_t wr = (int16_t)(aw / bw), hr = (int16_t)(ah / bh);
    for (uint32_t y = 0; y < bh; ++y)
        for (uint32_t x = 0; x < bw; ++x) {
            const int16_t vb = (int16_t)px_at(b, bx + x, by + y);
            const float smp = (float)fo_sample_sum4(a, ax, ay, aw, ah, x * (uint32_t)wr, y * (uint32_t)hr, t) / 4.0f;
            const float v = (float)vb - smp;
            sum += v * v;
        }
    return (double)sum / (double)(aw * ah);
}

/* BrightnessBlocksClassifier2::getCategory(double x4), encode/Classifier2.cpp:8-53. */
int fo_category4(double a1, double a2, double a3, double a4) {
    const int a1a2 = a1 > a2, a1a3 = a1 > a3, a1a4 = a1 > a4;
    const int a2a1 = a2 > a1, a2a3 = a2 > a3, a2a4 = a2 > a4;
    const int a3a1 = a3 > a1, a3a2 = a3 > a2, a3a4 = a3 > a4;
    const int a4a1 = a4 > a1, a4a2 = a4 > a2, a4a3 = a4 > a3;
    if (a1a2 && a2a3 && a3a4) return 0;
    if (a3a1 && a1a4 && a4a2) return 0;
    if (a4a3 && a3a2 && a2a1) return 0;
    if (a2a4 && a4a1 && a1a3) return 0;
    if (a1a3 && a3a2 && a2a4) return 1;
    if (a2a1 && a1a4 && a4a3) return 1;
    if (a4a2 && a2a3 && a3a1) return 1;
    if (a3a4 && a4a1 && a1a2) return 1;
    if (a1a4 && a4a3 && a3a2) return 2;
    if (a4a1 && a1a2 && a2a3) return 2;
    if (a3a2 && a2a4 && a4a1) return 2;
    if (a2a3 && a3a1 && a1a4) return 2;
    if (a1a2 && a2a4 && a4a3) return 3;
    if (a3a1 && a1a2 && a2a4) return 3;
    if (a4a3 && a3a1 && a1a2) return 3;
    if (a2a4 && a4a3 && a3a1) return 3;
    if (a2a1 && a1a3 && a3a4) return 4;
    if (a1a3 && a3a4 && a4a2) return 4;
    if (a3a4 && a4a2 && a2a1) return 4;
    if (a4a2 && a2a1 && a1a3) return 4;
    if (a1a4 && a4a2 && a2a3) return 5;
    if (a4a1 && a1a3 && a3a4) return 5;
    if (a2a3 && a3a4 && a4a1) return 5;
    if (a3a2 && a2a1 && a1a4) return 5;
    return -1;
}

/* getCategory(image, item), Classifier2.cpp:55-62; quadrants per partition2.hpp:19-30. */
int fo_category(const fo_plane* img, uint32_t x, uint32_t y, uint32_t w, uint32_t h) {
    const uint32_t hw = w / 2, hh = h / 2;
    const double a1 = fo_block_sum(img, x, y, hw, hh);
    const double a2 = fo_block_sum(img, x + hw, y, hw, hh);
    const double a3 = fo_block_sum(img, x, y + hh, hw, hh);
    const double a4 = fo_block_sum(img, x + hw, y + hh, hw, hh);
    return fo_category4(a1, a2, a3, a4);
}

/* createUniformGrid, image/partition2.hpp:110-135. */
size_t fo_create_uniform_grid(uint32_t W, uint32_t H, uint32_t sx, uint32_t sy, uint32_t ox,
                              uint32_t oy, fo_grid_item* out, size_t cap) {
    if (!sx || !sy || !ox || !oy) return 0;
    if (W % sx || H % sy || W % ox || H % oy) return 0; /* FRAC_ASSERTs :119-120 */
    size_t n = 0;
    uint32_t x = 0, y = 0;
    for (;;) {
        if (n < cap && out) {
            fo_grid_item it = {x, y, sx, sy, -1};
            out[n] = it;
        }
        ++n;
        x += ox;
        if (x + sx > W) {
            x = 0;
            y += oy;
            if (y + sy > H) break;
        }
    }
    return n;
}

void fo_preclassify(const fo_plane* img, fo_grid_item* items, size_t n) {
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n; ++i)
        items[i].bin = fo_category(img, items[i].x, items[i].y, items[i].w, items[i].h);
}

typedef struct {
    double distance, contrast, brightness;
    int transform;
} score_t;

static const score_t kDefaultScore = {100000.0, 0.0, 0.0, 0}; /* encode/datatypes.h:8-13 */

static double truncate_smax(double s, double smax) { /* transformmatcher.h:27-31 */
    if (smax > 0.0) return s > smax ? smax : (s < -smax ? -smax : s);
    return s;
}

/* TransformMatcher::match_generic / match_16to4, transformmatcher.h:81-144. */
static score_t match_type(const fo_plane* src, const fo_grid_item* d, const fo_plane* tgt,
                          const fo_grid_item* r, const fo_params* p, int t, score_t prev) {
    const double sumA = (double)fo_block_sum(tgt, r->x, r->y, r->w, r->h); /* :85 */
    score_t c;
    c.distance = fo_distance(src, tgt, d->x, d->y, d->w, d->h, r->x, r->y, r->w, r->h, t);
    c.transform = t;
    c.contrast = 0.0;
    c.brightness = 0.0;
    if (c.distance <= prev.distance) {
        const double N = (double)(r->w * r->h);
        double sumA2 = 0.0, sumB = 0.0, sumAB = 0.0;
        for (uint32_t y = 0; y < r->h; ++y)
            for (uint32_t x = 0; x < r->w; ++x) {
                const uint32_t sy = (y * d->h) / r->h;
                const uint32_t sx = (x * d->w) / r->w;
                const double a = (double)px_at(tgt, r->x + x, r->y + y);
                const double b = (double)fo_sample_sum4(src, d->x, d->y, d->w, d->h, sx, sy, t) / 4.0;
                sumB += b;
                sumA2 += a * a;
                sumAB += a * b;
            }
        const double tmp = N * sumA2 - (sumA - 1) * sumA; /* :103 */
        const double s = truncate_smax(fabs(tmp) < 0.00001 ? 0.0 : (N * sumAB - sumA * sumB) / tmp, p->s_max);
        const double o = p->fma ? fma(-s, sumA, sumB) / N : (sumB - s * sumA) / N; /* :105, SURVEY S10 */
        c.contrast = s;
        c.brightness = o;
        return c;
    }
    return prev;
}

/* TransformMatcher::match / matchTransformTypes, transformmatcher.h:38-69. */
static score_t match_chain(const fo_plane* src, const fo_grid_item* d, const fo_plane* tgt,
                           const fo_grid_item* r, const fo_params* p) {
    score_t prev = kDefaultScore;
    const int n_iso = p->isometries == 8 ? 8 : 4;
    for (int t = 0; t < n_iso; ++t) {
        const score_t res = match_type(src, d, tgt, r, p, t, prev);
        if (res.distance <= p->rms_threshold) return res; /* checkDistance :32-34 */
        prev = res.distance <= prev.distance ? res : prev;
    }
    return prev;
}

void fo_match(const fo_plane* src, const fo_grid_item* dom, const fo_plane* tgt,
              const fo_grid_item* rng, const fo_params* p, fo_encode_item* out) {
    const score_t s = match_chain(src, dom, tgt, rng, p);
    memset(out, 0, sizeof(*out));
    out->distance = s.distance;
    out->contrast = s.contrast;
    out->brightness = s.brightness;
    out->transform = s.transform;
}

/* BrightnessBlocksClassifier2::compare, Classifier2.cpp:70-81. */
static int classifier_compare(const fo_plane* src, const fo_plane* tgt, const fo_grid_item* d,
                              const fo_grid_item* r) {
    int sc = d->bin, tc = r->bin;
    if (sc == -1) sc = fo_category(src, d->x, d->y, d->w, d->h);
    if (tc == -1) tc = fo_category(tgt, r->x, r->y, r->w, r->h);
    return sc == tc;
}

/* TransformEstimator2::estimate (TransformEstimator2.hpp:29-48) wrapped as
 * CpuEncodingEngine2::encode_impl (EncodingEngine2.hpp:100-109). */
void fo_estimate(const fo_plane* src, const fo_plane* tgt, const fo_grid_item* domains, size_t n_dom,
                 const fo_grid_item* rng, const fo_params* p, fo_encode_item* out) {
    score_t best = kDefaultScore;
    uint32_t bx = 0, by = 0, bw = 0, bh = 0;
    fo_grid_item r = *rng;
    if (p->use_classifier && r.bin == -1) /* result-neutral hoist of the per-compare recompute */
        r.bin = fo_category(tgt, r.x, r.y, r.w, r.h);
    for (size_t i = 0; i < n_dom; ++i) {
        const fo_grid_item* d = &domains[i];
        if (!p->use_classifier || classifier_compare(src, tgt, d, &r)) {
            const score_t s = match_chain(src, d, tgt, &r, p);
            if (s.distance < best.distance) {
                best = s;
                bx = d->x;
                by = d->y;
                bw = d->w;
                bh = d->h;
            }
            if (best.distance <= p->rms_threshold) break;
        }
    }
    memset(out, 0, sizeof(*out));
    out->x = r.x;
    out->y = r.y;
    out->w = r.w;
    out->h = r.h;
    out->distance = best.distance;
    out->contrast = best.contrast;
    out->brightness = best.brightness;
    out->transform = best.transform;
    out->match_x = bx;
    out->match_y = by;
    out->src_w = bw;
    out->src_h = bh;
}

void fo_encode_level(const fo_plane* src, const fo_plane* tgt, const fo_grid_item* domains,
                     size_t n_dom, const fo_grid_item* ranges, size_t n_rng, const fo_params* p,
                     int nthreads, size_t sample_stride, fo_encode_item* out) {
    if (sample_stride == 0) sample_stride = 1;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_num_procs();
#endif
    (void)nthreads;
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
    for (long long i = 0; i < (long long)n_rng; ++i) {
        if ((size_t)i % sample_stride) {
            memset(&out[i], 0, sizeof(out[i]));
            continue;
        }
        fo_estimate(src, tgt, domains, n_dom, &ranges[i], p, &out[i]);
    }
}

size_t fo_encode_quadtree(const fo_plane* img, uint32_t t_max, uint32_t t_min, const fo_params* p,
                          int nthreads, fo_encode_item* out, size_t cap, size_t* level_counts) {
    size_t n_pending = fo_create_uniform_grid(img->width, img->height, t_max, t_max, t_max, t_max, NULL, 0);
    if (!n_pending) return (size_t)-1;
    fo_grid_item* pending = (fo_grid_item*)malloc(n_pending * sizeof(fo_grid_item));
    fo_create_uniform_grid(img->width, img->height, t_max, t_max, t_max, t_max, pending, n_pending);
    size_t n_out = 0;
    int level = 0;
    for (uint32_t T = t_max; T >= t_min && n_pending; T /= 2, ++level) {
        const size_t n_dom = fo_create_uniform_grid(img->width, img->height, 2 * T, 2 * T, T, T, NULL, 0);
        fo_grid_item* dom = (fo_grid_item*)malloc((n_dom ? n_dom : 1) * sizeof(fo_grid_item));
        fo_create_uniform_grid(img->width, img->height, 2 * T, 2 * T, T, T, dom, n_dom);
        if (p->use_classifier) {
            fo_preclassify(img, dom, n_dom);
            fo_preclassify(img, pending, n_pending);
        }
        fo_encode_item* res = (fo_encode_item*)malloc(n_pending * sizeof(fo_encode_item));
        fo_encode_level(img, img, dom, n_dom, pending, n_pending, p, nthreads, 1, res);
        fo_grid_item* next = (fo_grid_item*)malloc(4 * n_pending * sizeof(fo_grid_item));
        size_t n_next = 0, emitted = 0;
        for (size_t i = 0; i < n_pending; ++i) {
            const int ok = res[i].distance <= p->rms_threshold; /* checkDistance */
            if (ok || T / 2 < t_min) {
                if (n_out >= cap) {
                    free(next); free(res); free(dom); free(pending);
                    return (size_t)-1;
                }
                out[n_out++] = res[i];
                ++emitted;
            } else {
                const uint32_t h = T / 2, x = pending[i].x, y = pending[i].y;
                const fo_grid_item c0 = {x, y, h, h, -1}, c1 = {x + h, y, h, h, -1};
                const fo_grid_item c2 = {x, y + h, h, h, -1}, c3 = {x + h, y + h, h, h, -1};
                next[n_next++] = c0; /* topLeft */
                next[n_next++] = c1; /* topRight */
                next[n_next++] = c2; /* bottomLeft */
                next[n_next++] = c3; /* bottomRight */
            }
        }
        if (level_counts) level_counts[level] = emitted;
        free(res);
        free(dom);
        free(pending);
        pending = next;
        n_pending = n_next;
    }
    free(pending);
    return n_out;
}

/* Frac::copy, encode/DecodeUtils.hpp:9-25. */
static void copy_item(const fo_plane* src, uint8_t* tgt, uint32_t tstride, const fo_encode_item* e, int use_fma) {
    for (uint32_t y = 0; y < e->h; ++y)
        for (uint32_t x = 0; x < e->w; ++x) {
            const uint32_t sx = (x * e->src_w) / e->w;
            const uint32_t sy = (y * e->src_h) / e->h;
            const double smp = (double)fo_sample_sum4(src, e->match_x, e->match_y, e->src_w, e->src_h, sx, sy, e->transform) / 4.0;
            const double v = use_fma ? fma(e->contrast, smp, e->brightness) : e->contrast * smp + e->brightness;
            tgt[(size_t)(e->x + x) + (size_t)(e->y + y) * tstride] = v < 0.0 ? 0 : v > 255 ? 255 : (uint8_t)v;
        }
}

/* Decoder2::decode / decodeStep, encode/Encoder2.hpp:67-99.  Items whose source
 * size is 0x0 (the default item_match_t of an empty classifier bucket) are
 * skipped: the reference would trip FRAC_ASSERT in the sampler and exit. */
void fo_decode(const fo_encode_item* items, size_t n, uint8_t* target, uint32_t width,
               uint32_t height, uint32_t stride, int max_iters, double rms_eps, int use_fma,
               int* iterations_out, double* rms_out) {
    const int iters = max_iters < 0 ? 300 : max_iters;
    const size_t bytes = (size_t)height * stride;
    uint8_t* srcbuf = (uint8_t*)malloc(bytes);
    memset(srcbuf, 100, bytes); /* Encoder2.hpp:69 */
    fo_plane src = {srcbuf, width, height, stride};
    fo_plane tgt = {target, width, height, stride};
    int i = 0;
    double rms = 0.0;
    for (; i < iters; ++i) {
#pragma omp parallel for schedule(static)
        for (long long k = 0; k < (long long)n; ++k)
            if (items[k].src_w && items[k].src_h) copy_item(&src, target, stride, &items[k], use_fma);
        rms = fo_distance(&src, &tgt, 0, 0, width, height, 0, 0, width, height, 0);
        if (rms < rms_eps) break;
        memcpy(srcbuf, target, bytes); /* source = _target.copy() */
    }
    free(srcbuf);
    if (iterations_out) *iterations_out = i;
    if (rms_out) *rms_out = rms;
}

/* Quantizer<double>, encode/Quantizer.hpp:13-36. */
uint64_t fo_quantize(double v, double vmin, double vmax, int bits) {
    const double step = fabs(vmax - vmin) / (double)(1 << bits);
    const uint64_t maxq = (uint64_t)((1 << bits) - 1);
    const uint64_t q = (uint64_t)floor((v - vmin) / step);
    return q < maxq ? q : maxq;
}

double fo_dequantize(uint64_t q, double vmin, double vmax, int bits) {
    const double step = fabs(vmax - vmin) / (double)(1 << bits);
    return (double)q * step + vmin + step / 2;
}

/* ---- synthetic inputs (ours, the reference has none) ---- */
static uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static uint32_t lattice(uint64_t seed, uint64_t i, uint64_t j, uint64_t o) {
    return (uint32_t)(splitmix64(seed ^ (o * 0xD6E8FEB86659FD93ull) ^ (i * 0x9E3779B97F4A7C15ull) ^ (j * 0xC2B2AE3D27D4EB4Full)) >> 56);
}

/* kind: 0 natural (4-octave integer value noise), 1 noise, 2 pattern (11x+43y+124)%256. */
void fo_synth_image(uint8_t* out, uint32_t w, uint32_t h, uint32_t stride, uint64_t seed, int kind) {
    static const uint32_t cell[4] = {64, 16, 4, 1}, wt[4] = {4, 2, 1, 1};
#pragma omp parallel for schedule(static)
    for (long long yy = 0; yy < (long long)h; ++yy) {
        const uint32_t y = (uint32_t)yy;
        for (uint32_t x = 0; x < w; ++x) {
            uint32_t v;
            if (kind == 2) {
                v = (11u * x + 43u * y + 124u) % 256u;
            } else if (kind == 1) {
                v = (uint32_t)(splitmix64(seed ^ ((uint64_t)x * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)y * 0xC2B2AE3D27D4EB4Full)) >> 56);
            } else {
                uint32_t acc = 0;
                for (int o = 0; o < 4; ++o) {
                    const uint32_t c = cell[o], i = x / c, j = y / c, fx = x % c, fy = y % c;
                    const uint32_t vo = ((c - fx) * (c - fy) * lattice(seed, i, j, o) + fx * (c - fy) * lattice(seed, i + 1, j, o) +
                                         (c - fx) * fy * lattice(seed, i, j + 1, o) + fx * fy * lattice(seed, i + 1, j + 1, o)) / (c * c);
                    acc += wt[o] * vo;
                }
                v = acc / 8;
            }
            out[(size_t)y * stride + x] = (uint8_t)v;
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * colour path: ImageIO::rgb2yuv (image/ImageIO.cpp:40-57) and yuv2rgb (:68-84); main.cpp:193-200 encodes the three planes.
 * Chroma is "subsampled" by overwriting: the last pixel of every 2x2 cell (odd x, odd y) wins.
 * ------------------------------------------------------------------------------------------------ */
static uint8_t clamp_u8(double x) { return x < 0.0 ? 0 : x > 255 ? 255 : (uint8_t)x; }   /* ImageIO.cpp:11-13 */

void fo_rgb2yuv(const uint8_t* rgb, uint32_t w, uint32_t h, uint32_t stride, uint8_t* yb, uint32_t ys, uint8_t* ub, uint32_t us, uint8_t* vb,
                uint32_t vs, int use_fma) {
    for (size_t y = 0; y < h; ++y)
        for (size_t x = 0; x < w; ++x) {
            const double r = rgb[x * 3 + y * stride + 0], g = rgb[x * 3 + y * stride + 1], b = rgb[x * 3 + y * stride + 2];
            double yp, up, vp;
            if (use_fma) {   /* what GCC -O2 -march=x86-64-v3 emits for a*r + b*g + c*b: the middle product stays a product, the first
                                and the last are fused onto it (checked against the compiled reference on every input that differs) */
                yp = fma(0.114, b, fma(0.299, r, 0.587 * g));
                up = fma(0.499, b, fma(-0.169, r, -0.331 * g)) + 128;
                vp = fma(-0.0813, b, fma(0.499, r, -0.418 * g)) + 128;
            } else {
                yp = 0.299 * r + 0.587 * g + 0.114 * b;
                up = -0.169 * r - 0.331 * g + 0.499 * b + 128;
                vp = 0.499 * r - 0.418 * g - 0.0813 * b + 128;
            }
            yb[x + y * ys] = clamp_u8(yp);
            ub[(x / 2) + (y / 2) * us] = clamp_u8(up);
            vb[(x / 2) + (y / 2) * vs] = clamp_u8(vp);
        }
}

void fo_yuv2rgb(const uint8_t* yb, uint32_t w, uint32_t h, uint32_t ys, const uint8_t* ub, uint32_t us, const uint8_t* vb, uint32_t vs,
                uint8_t* rgb, uint32_t rgb_stride, int use_fma) {
    for (size_t y = 0; y < h; ++y)
        for (size_t x = 0; x < w; ++x) {
            uint8_t* p = rgb + (x * 3 + y * rgb_stride * 3);
            const double yp = yb[x + y * ys], up = ub[(x / 2) + (y / 2) * us], vp = vb[(x / 2) + (y / 2) * vs];
            if (use_fma) {
                p[0] = clamp_u8(fma(1.402, vp - 128, yp));
                p[1] = clamp_u8(fma(-0.714, vp - 128, fma(-0.344, up - 128, yp)));
                p[2] = clamp_u8(fma(1.772, up - 128, yp));
            } else {
                p[0] = clamp_u8(yp + 1.402 * (vp - 128));
                p[1] = clamp_u8(yp - 0.344 * (up - 128) - 0.714 * (vp - 128));
                p[2] = clamp_u8(yp + 1.772 * (up - 128));
            }
        }
}
