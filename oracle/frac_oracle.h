/*
 * frac_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the reference's fractal-encoding search path,
 * used ONLY as the parity checker by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs.  The product path
 * (fractencode_b200/csrc + include/fractencode_b200.h) never links, loads or
 * calls anything in oracle/.
 *
 * Every function follows, statement by statement, the reference file:line it
 * cites (paths relative to the reference checkout, sebsgit/fractencode).  The
 * restatement is pinned two ways (see oracle/README.md and tests/):
 *   - against the reference's own known-answer tests (tests/TransformMatcherTest.cpp,
 *     TransformEstimatorTest.cpp, ImageSamplerTest.cpp, ClassifierTest.cpp,
 *     PartitionTests.cpp, ImageStatisticsTest.cpp), and
 *   - against the real reference compiled from /root/reference into
 *     oracle/_ref/ (oracle/ref_driver.cpp), on seeded images, with the outputs
 *     committed as fixtures under tests/golden/.
 *
 * The same C interface is exported twice: by this restatement with the prefix
 * `fo_` and by oracle/ref_driver.cpp (the real reference classes) with the
 * prefix `fr_`.
 */
#ifndef FRAC_ORACLE_H
#define FRAC_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Frac2::UniformGridItem, image/partition2.hpp:13-16,93-99 (20 bytes). */
typedef struct {
    uint32_t x, y, w, h;
    int32_t bin; /* GridItemData::bb_classifierBin, -1 = not classified */
} fo_grid_item;

/* Frac::encode_item_t, encode/datatypes.h:8-23 (64 bytes). */
typedef struct {
    uint32_t x, y, w, h;
    double distance, contrast, brightness;
    int32_t transform;
    int32_t pad_;
    uint32_t match_x, match_y;
    uint32_t src_w, src_h;
} fo_encode_item;

typedef struct {
    const uint8_t* px;
    uint32_t width, height, stride;
} fo_plane;

/* encode/transformmatcher.h:20-34 + the compile-flag ambiguity of SURVEY S10:
 * fma != 0 evaluates `sumB - s*sumA` and `contrast*sample + brightness` as one
 * fused multiply-add (what GCC emits for the reference with -march=native). */
typedef struct {
    double rms_threshold;
    double s_max;
    int use_classifier; /* 0 = DummyClassifier, 1 = BrightnessBlocksClassifier2 */
    int fma;
    int isometries;     /* 0 / 4: the four rotations TransformMatcher::match tries (transformmatcher.h:38-46); 8: the chain goes on
                           through Flip .. Flip_Rotate_270 (image/transform.h:20-24) with the same rules -- OURS, the
                           compiled reference ignores it */
    int reserved_;
} fo_params;

/* image/sampler.h:22-38 + image/transform.h:96-109.  Returns the 2x2 box SUM
 * (= 4 * SamplerBilinear::sample), an exact integer 0..1020. */
int fo_sample_sum4(const fo_plane* img, uint32_t px, uint32_t py, uint32_t pw, uint32_t ph,
                   uint32_t lx, uint32_t ly, int transform);

/* image/ImageStatistics.hpp:13-17, .cpp:4-47 (plain integer block sum). */
uint32_t fo_block_sum(const fo_plane* img, uint32_t x, uint32_t y, uint32_t w, uint32_t h);

/* image/metrics.h:21-51, both branches. */
double fo_distance(const fo_plane* a, const fo_plane* b, uint32_t ax, uint32_t ay, uint32_t aw,
                   uint32_t ah, uint32_t bx, uint32_t by, uint32_t bw, uint32_t bh, int transform);

/* encode/Classifier2.cpp:8-62. */
int fo_category4(double a1, double a2, double a3, double a4);
int fo_category(const fo_plane* img, uint32_t x, uint32_t y, uint32_t w, uint32_t h);

/* image/partition2.hpp:110-135.  Returns the item count; writes at most cap
 * items (bin = -1).  Returns 0 when the image is not aligned to size/step. */
size_t fo_create_uniform_grid(uint32_t img_w, uint32_t img_h, uint32_t size_x, uint32_t size_y,
                              uint32_t step_x, uint32_t step_y, fo_grid_item* out, size_t cap);

/* encode/Classifier2.cpp:64-68 applied to a list (main.cpp:155-162). */
void fo_preclassify(const fo_plane* img, fo_grid_item* items, size_t n);

/* encode/transformmatcher.h:38-69: one domain, rotation chain. Writes
 * distance/contrast/brightness/transform of the returned score. */
void fo_match(const fo_plane* src, const fo_grid_item* dom, const fo_plane* tgt,
              const fo_grid_item* rng, const fo_params* p, fo_encode_item* score_out);

/* encode/TransformEstimator2.hpp:29-48 + encode/EncodingEngine2.hpp:100-109. */
void fo_estimate(const fo_plane* src, const fo_plane* tgt, const fo_grid_item* domains, size_t n_dom,
                 const fo_grid_item* rng, const fo_params* p, fo_encode_item* out);

/* EncodingEngineCore2::encode without the racy queue: every range item
 * independently, out[i] <-> ranges[i].  nthreads <= 0 -> all cores.
 * Ranges i with (i % sample_stride) != 0 are skipped (out[i] zeroed); pass 1. */
void fo_encode_level(const fo_plane* src, const fo_plane* tgt, const fo_grid_item* domains,
                     size_t n_dom, const fo_grid_item* ranges, size_t n_rng, const fo_params* p,
                     int nthreads, size_t sample_stride, fo_encode_item* out);

/* Quadtree driver (OURS; the reference has none): level T block
 * is emitted if checkDistance(best) or T == t_min, else split into
 * topLeft, topRight, bottomLeft, bottomRight (image/partition2.hpp:19-30);
 * domains at level T: size 2T, step T.  Emission order: level by level
 * (t_max first), inside a level in pending-list order.  level_counts[l] gets
 * the number of items emitted at level l (T = t_max >> l). Returns n emitted,
 * or (size_t)-1 if cap is too small. */
size_t fo_encode_quadtree(const fo_plane* img, uint32_t t_max, uint32_t t_min, const fo_params* p,
                          int nthreads, fo_encode_item* out, size_t cap, size_t* level_counts);

/* encode/Encoder2.hpp:54-99 + encode/DecodeUtils.hpp:9-25.  `target` is
 * in/out (height*stride bytes). max_iters < 0 -> 300.  Returns Decoder2's
 * decode_stats_t {iterations, rms}. */
void fo_decode(const fo_encode_item* items, size_t n, uint8_t* target, uint32_t width,
               uint32_t height, uint32_t stride, int max_iters, double rms_eps, int fma,
               int* iterations_out, double* rms_out);

/* encode/Quantizer.hpp:13-36. */
uint64_t fo_quantize(double v, double vmin, double vmax, int bits);
double fo_dequantize(uint64_t q, double vmin, double vmax, int bits);

/* Synthetic images of SURVEY 8d (ours; identical generator in the product's bench). */
void fo_synth_image(uint8_t* out, uint32_t w, uint32_t h, uint32_t stride, uint64_t seed, int kind);
/* image/ImageIO.cpp:40-57 and :68-84.  `stride` of rgb2yuv is in BYTES per row (the reference indexes x*3 + y*stride);
 * `rgb_stride` of yuv2rgb is in PIXELS per row (the reference indexes x*3 + y*rgbStride*3).  fma: the contraction GCC applies
 * under -march=native. */
void fo_rgb2yuv(const uint8_t* rgb, uint32_t w, uint32_t h, uint32_t stride, uint8_t* y, uint32_t ys, uint8_t* u, uint32_t us, uint8_t* v,
                uint32_t vs, int fma);
void fo_yuv2rgb(const uint8_t* y, uint32_t w, uint32_t h, uint32_t ys, const uint8_t* u, uint32_t us, const uint8_t* v, uint32_t vs,
                uint8_t* rgb, uint32_t rgb_stride, int fma);

int fo_hardware_threads(void);
const char* fo_version(void);

#ifdef __cplusplus
}
#endif
#endif
