/*
 * fractencode_b200.h -- C ABI of the B200-native fractal-encoding search.
 *
 * This is the drop-in boundary for the reference's encode/ hot path
 * (sebsgit/fractencode; file:line citations are into that repository).  A
 * reference maintainer binds these entry points from a
 * Frac2::AbstractEncodingEngine2 subclass (encode/EncodingEngine2.hpp:50-85;
 * the slot left open at encode/EncodingEngine2.cpp:21-26) -- see INTEGRATION.md
 * and fractencode_b200/host/ for that subclass.  Plain pointers and sizes only;
 * no C++ or torch types cross this boundary; no exceptions; every function
 * returns FE_OK (0) or a negative fe_status and records a message readable
 * with fe_last_error().
 *
 * Threading: one fe_ctx per device and per host thread (the reference runs one
 * thread per engine, encode/EncodingEngine2.hpp:126-155).  A ctx is not
 * thread-safe.  All work is issued on the ctx's CUDA stream.  Every call makes
 * the ctx's device the calling thread's current CUDA device (cudaSetDevice) and
 * leaves it so: a host that drives several devices from ONE thread restores its
 * own current device after a call.
 *
 * There is NO CPU fallback: fe_create fails when no sm_100 device is present.
 */
#ifndef FRACTENCODE_B200_H
#define FRACTENCODE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FE_ABI_VERSION 3

typedef enum {
    FE_OK = 0,
    FE_ERR_INVALID = -1,     /* bad argument (null pointer, zero size, unaligned image ...) */
    FE_ERR_UNSUPPORTED = -2, /* geometry outside the supported family (see fe_encode_level) */
    FE_ERR_CUDA = -3,        /* CUDA runtime error; message holds cudaGetErrorString */
    FE_ERR_NO_DEVICE = -4,   /* no sm_100 GPU: the product has no CPU path */
    FE_ERR_CAPACITY = -5,    /* caller buffer too small */
    FE_ERR_STATE = -6        /* call order (e.g. encode before fe_set_image) */
} fe_status;

/* Frac2::UniformGridItem = GridItemBase{origin,size} + GridItemData{bb_classifierBin}
 * (image/partition2.hpp:13-16, 88-99): 20 bytes, same layout. */
typedef struct {
    uint32_t x, y, w, h;
    int32_t bin; /* -1 = not classified; the engine classifies it like Classifier2::compare does */
} fe_grid_item;

/* Frac::encode_item_t (encode/datatypes.h:8-23): 64 bytes, same layout
 * (x,y,w,h | transform_score_t{distance,contrast,brightness,transform} | match x,y | sourceItemSize). */
typedef struct {
    uint32_t x, y, w, h;
    double distance, contrast, brightness;
    int32_t transform; /* Frac::TransformType, image/transform.h:16-25 */
    int32_t pad_;
    uint32_t match_x, match_y;
    uint32_t src_w, src_h;
} fe_encode_item;

/* Search engine selection (diagnostics / A-B parity; FE_SEARCH_AUTO in production). */
typedef enum {
    FE_SEARCH_AUTO = 0,  /* tcgen05 path where its exactness proof holds, exact integer path otherwise */
    FE_SEARCH_EXACT = 1, /* CUDA-core integer (dp4a) search, any geometry */
    FE_SEARCH_UMMA = 2   /* force the tcgen05/TMEM path; FE_ERR_UNSUPPORTED when not applicable */
} fe_search_impl;

/* Frac::encode_parameters_t (encode/encode_parameters.h:5-14) + TransformMatcher(rmsThreshold, sMax)
 * (encode/transformmatcher.h:20-34) + the classifier choice of main.cpp:152-154. */
typedef struct {
    double rms_threshold; /* checkDistance: d <= rms_threshold; reference default 0 */
    double s_max;         /* truncateSMax: clamp contrast to +-s_max when > 0; default -1 */
    int32_t use_classifier; /* 0 = DummyClassifier, 1 = BrightnessBlocksClassifier2 */
    int32_t fma;          /* 0: brightness = (sumB - s*sumA)/N with separate roundings (reference built
                             without FMA); 1: fused, as GCC emits for the reference's -march=native
                             build.  Same switch for decode's s*v+o. */
    int32_t search_impl;  /* fe_search_impl */
    int32_t isometries;   /* 0 or 4: the four rotations TransformMatcher::match tries (transformmatcher.h:38-46) -- the reference's
                             results.  8: the chain goes on through Flip, Flip_Rotate_90/180/270 (image/transform.h:20-24, 37-40)
                             with the same rules (first isometry under the threshold, else the minimum, ties to the later one):
                             an extension, results differ from the reference by design; tensor paths only. */
} fe_params;

typedef struct {
    uint64_t matches;          /* (range, domain, rotation) candidates of the workload = admissible pairs x 4: what a scan
                                * without early-out scores (the figure BASELINE.md calls "matches") */
    uint64_t kernel_launches;  /* kernels launched by this ctx since fe_stats_reset */
    uint64_t fp32_regime_items;/* items whose best SSE >= 2^20 (reference fp32 sum rounds; distance emulated) */
    uint64_t umma_levels;      /* levels searched on the tcgen05 path */
    uint64_t exact_levels;     /* levels searched on the exact integer path */
    uint64_t level_items[8];   /* last quadtree: items emitted per level */
    uint64_t level_ranges[8];  /* last quadtree: range blocks searched per level */
    uint64_t level_matches[8]; /* last quadtree: matches per level */
    float level_search_ms[8];  /* last quadtree: search-kernel time per level (CUDA events on the ctx stream) */
    float level_prep_ms[8];    /* last quadtree: pool/range prep + classify time per level */
    float last_decode_ms;
    uint32_t reserved_;
    uint64_t evaluated;        /* candidates the search kernels actually scored: < matches when a threshold lets range
                                * blocks stop at their first hit (the reference's break, TransformEstimator2.hpp:40-41) */
    uint64_t level_evaluated[8]; /* last quadtree: evaluated per level */
    uint64_t level_passes[8];  /* last quadtree: search passes (kernel launches) per level */
    uint64_t prefiltered;      /* candidates only looked at by the lower-bound prefilter of the large-block levels (an 8 x 8 bound,
                                * 64 products each; those it lets through are scored exactly and count as `evaluated`) */
    uint64_t level_prefiltered[8];
} fe_stats;

typedef struct fe_ctx fe_ctx;

/* Replaces: construction of an engine for EncodingEngineCore2's engine list
 * (encode/EncodingEngine2.cpp:12-29).  `stream` is a cudaStream_t to issue on
 * (e.g. the caller's current stream) or NULL for a ctx-owned stream. */
int fe_create(fe_ctx** out, int device, void* stream);
void fe_destroy(fe_ctx* ctx);
const char* fe_last_error(const fe_ctx* ctx); /* ctx may be NULL: error of the last failed fe_create */
int fe_abi_version(void);

/* Replaces: the `const ImagePlane& sourceImage` the engine/estimator constructors
 * capture (encode/EncodingEngine2.hpp:52-59, encode/TransformEstimator2.hpp:14-27).
 * Host pixels are copied to the device; the caller keeps ownership.  `px` has
 * height*stride bytes, stride >= width (image/Image2.hpp:84-111).
 * The copy is enqueued on the context's stream (cudaMemcpyAsync): pageable memory has been
 * read when the call returns; PAGE-LOCKED memory is read asynchronously and must stay
 * unchanged until the next call on this context that returns results to the host (any
 * fe_encode_*, fe_classify, fe_get_image ...), which orders itself behind the copy. */
int fe_set_image(fe_ctx* ctx, const uint8_t* px, uint32_t width, uint32_t height, uint32_t stride);
/* Separate source (domain) and target (range) planes, as TransformEstimator2's
 * (sourceImage, targetImage) pair allows (tests/TransformEstimatorTest.cpp:13-47). */
int fe_set_images(fe_ctx* ctx, const uint8_t* src_px, uint32_t src_w, uint32_t src_h, uint32_t src_stride,
                  const uint8_t* tgt_px, uint32_t tgt_w, uint32_t tgt_h, uint32_t tgt_stride);
/* Same, but `dev_px` already lives in this device's memory (copied device-to-device ON THE
 * CTX's STREAM: pixels written on another stream must be complete -- event or synchronise --
 * before this call; with the caller's own stream passed to fe_create there is nothing to do). */
int fe_set_image_device(fe_ctx* ctx, const void* dev_px, uint32_t width, uint32_t height, uint32_t stride);

/* Replaces: BrightnessBlocksClassifier2::preclassify / getCategory
 * (encode/Classifier2.cpp:55-68) over a list.  which: 0 = source image, 1 = target image.
 * bins_out[i] in {-1, 0..5}. */
int fe_classify(fe_ctx* ctx, int which, const fe_grid_item* items, size_t n, int32_t* bins_out);

/* Replaces: init() + encode(item) x n + finalize() of an AbstractEncodingEngine2
 * (encode/EncodingEngine2.hpp:63-72,100-109), i.e. TransformEstimator2::estimate
 * (encode/TransformEstimator2.hpp:29-48) for every range item against the domain
 * list in list order.  out[i] corresponds to ranges[i].
 * Supported family: all domains one square size S, all ranges one square size T,
 * S a multiple of T, S > T, T <= 64, blocks inside their image.  The tcgen05 path
 * additionally needs S == 2T and even domain origins (true for createUniformGrid
 * lattices with step S/2, main.cpp:145-162); other geometries (e.g. the
 * reference's default 16->4) run on the exact integer path. */
int fe_encode_level(fe_ctx* ctx, const fe_grid_item* domains, size_t n_domains,
                    const fe_grid_item* ranges, size_t n_ranges, const fe_params* params,
                    fe_encode_item* out);

/* Quadtree partition (image/partition2.hpp:18-30 children + transformmatcher.h:32-34
 * checkDistance; the reference parses --quadtree but never implements it).
 * Level T (t_max, t_max/2, ..., t_min): domains = createUniformGrid(size 2T, step T),
 * a block is emitted when checkDistance(best) or T == t_min, else replaced by its
 * topLeft, topRight, bottomLeft, bottomRight children.  Items are written level by
 * level, in pending-list order.  level_counts (may be NULL) gets items per level. */
int fe_encode_quadtree(fe_ctx* ctx, uint32_t t_max, uint32_t t_min, const fe_params* params,
                       fe_encode_item* out, size_t cap, size_t* n_out, size_t* level_counts);
/* Same search, results left in device memory (for device-resident pipelines and
 * kernel-only timing); fetch with fe_fetch_items. */
int fe_encode_quadtree_device(fe_ctx* ctx, uint32_t t_max, uint32_t t_min, const fe_params* params,
                              size_t* n_out);
/* One shard of the same search (range blocks sharded across GPUs, pool replicated): only the top-level range
 * blocks first_block .. first_block + n_blocks - 1 of createUniformGrid(t_max, t_max) order (x fastest) are encoded -- they
 * and their descendants, against the domains of the WHOLE image.  The shards of a partition of the top-level grid together
 * give exactly the transform list of fe_encode_quadtree (range blocks are independent). */
int fe_encode_quadtree_slice_device(fe_ctx* ctx, uint32_t t_max, uint32_t t_min, const fe_params* params,
                                    size_t first_block, size_t n_blocks, size_t* n_out);
/* Batch mode (BASELINE config 5): n_images planes of the same size, each encoded like fe_encode_quadtree does -- what the
 * reference does plane after plane (main.cpp:142-181; the three planes of --color, main.cpp:193-200).  Image i's items
 * go to out[i * cap_per_image ...], its count to n_out[i].  The images are pipelined over two internal streams (the upload
 * and the levels of one overlap the other's), so pinned host buffers pay; results are identical to n single calls. */
int fe_encode_batch(fe_ctx* ctx, const uint8_t* const* images, size_t n_images, uint32_t width, uint32_t height, uint32_t stride,
                    uint32_t t_max, uint32_t t_min, const fe_params* params, fe_encode_item* out, size_t cap_per_image,
                    size_t* n_out);
/* The same for planes of DIFFERENT sizes -- the colour path: main.cpp:193-200 encodes the luma plane and the two half-size chroma
 * planes of a YUV 4:2:0 image with the same parameters.  Plane i's items go to out[out_offsets[i] ...] (capacity caps[i]). */
int fe_encode_planes(fe_ctx* ctx, const uint8_t* const* planes, size_t n_planes, const uint32_t* widths, const uint32_t* heights,
                     const uint32_t* strides, uint32_t t_max, uint32_t t_min, const fe_params* params, fe_encode_item* out,
                     const size_t* out_offsets, const size_t* caps, size_t* n_out);
/* Replaces: ImageIO::rgb2yuv / ImageIO::yuv2rgb (image/ImageIO.cpp:40-57, 68-84): interleaved 8-bit RGB <-> Y (width x height)
 * and U, V (width/2 x height/2; the chroma of a 2x2 cell is that of its last pixel, as the reference's loop leaves it).  Host
 * buffers; width and height even.  rgb_stride_bytes = bytes per RGB row (rgb2yuv indexes x*3 + y*stride); rgb_stride_pixels =
 * pixels per RGB row (yuv2rgb indexes 3*(x + y*stride)) -- the reference's own two conventions.  fma as in fe_params. */
int fe_rgb_to_yuv420(fe_ctx* ctx, const uint8_t* rgb, uint32_t width, uint32_t height, uint32_t rgb_stride_bytes, uint8_t* y,
                     uint32_t y_stride, uint8_t* u, uint32_t u_stride, uint8_t* v, uint32_t v_stride, int fma);
int fe_yuv420_to_rgb(fe_ctx* ctx, const uint8_t* y, uint32_t width, uint32_t height, uint32_t y_stride, const uint8_t* u,
                     uint32_t u_stride, const uint8_t* v, uint32_t v_stride, uint8_t* rgb, uint32_t rgb_stride_pixels, int fma);
int fe_fetch_items(fe_ctx* ctx, fe_encode_item* out, size_t cap, size_t* n_out);
/* Device pointer to the last result list (n items of 64 bytes), valid until the next encode. */
const void* fe_device_items(const fe_ctx* ctx, size_t* n_out);

/* Replaces: Decoder2::decode (encode/Encoder2.hpp:54-99) with Frac::copy
 * (encode/DecodeUtils.hpp:9-25) as a gather kernel.  `target` (height*stride bytes)
 * is in/out like Decoder2's ImagePlane&; the iteration source starts at 100.
 * max_iters < 0 -> 300.  Returns Decoder2::decode_stats_t in *iterations / *rms. */
int fe_decode(fe_ctx* ctx, const fe_encode_item* items, size_t n, uint8_t* target, uint32_t width,
              uint32_t height, uint32_t stride, int max_iters, double rms_eps, int fma,
              int* iterations, double* rms);

/* Replaces: Frac::copy (encode/DecodeUtils.hpp:9-25) / Decoder2::decodeStep (encode/Encoder2.hpp:91-99):
 * one pass target[item] = clamp(trunc(s * sample(source, item) + o)) over n items with an explicit
 * source plane.  `source` and `target` are host planes of the same width/height/stride; target is in/out. */
int fe_copy_items(fe_ctx* ctx, const uint8_t* source, uint8_t* target, uint32_t width, uint32_t height,
                  uint32_t stride, const fe_encode_item* items, size_t n, int fma);

/* Replaces: the Quantizer post-pass of main.cpp:106-140 (encode/Quantizer.hpp:13-36):
 * min/max over the list, then quantized(contrast) with bits_s bits and
 * quantized(brightness) with bits_o bits.  minmax_out = {min_s, max_s, min_o, max_o}. */
int fe_quantize(fe_ctx* ctx, const fe_encode_item* items, size_t n, int bits_s, int bits_o,
                uint32_t* q_s_out, uint32_t* q_o_out, double minmax_out[4]);

/* Packed, quantised transform records (ours: the reference has no serialised format; built from its
 * Quantizer, encode/Quantizer.hpp:13-36, with the bit widths of main.cpp:120-121).  One 64-bit word per item:
 *   bits  0..10  range  x / T        bits 11..21  range  y / T        (T = t_max >> level)
 *   bits 22..32  domain x / T        bits 33..43  domain y / T        (domain origins are multiples of T)
 *   bits 44..45  level               bits 46..48  transform
 *   bits 49..53  quantized(contrast), bits_s <= 5 bits     bits 54..60  quantized(brightness), bits_o <= 7 bits
 *   bit  63      item has no match (default item_match_t)
 * minmax = {min_s, max_s, min_o, max_o} over the list as main.cpp:109-118 computes them; it is the stream header
 * together with (t_max, bits_s, bits_o).  Requires square power-of-two blocks with S = 2T on their lattice. */
int fe_pack_items(fe_ctx* ctx, const fe_encode_item* items, size_t n, uint32_t t_max, int bits_s, int bits_o,
                  uint64_t* packed_out, double minmax_out[4]);
/* The same post-pass on the DEVICE-RESIDENT result list of the last quadtree encode, for pipelines that never bring the 64-byte
 * records to the host (multi-GPU gathers ship the 8-byte records: SURVEY 8e/8f-1).  Nothing here synchronises.
 *   fe_items_minmax_device: minmax_dev[4] (device memory) = {min_s, max_s, min_o, max_o} over the list as main.cpp:109-118
 *                           computes them (max starts at -1, min at DBL_MAX).  When one image is sharded over several GPUs
 *                           the caller reduces the four values across ranks (min / max) before packing.
 *   fe_pack_items_device:   packed_dev[0 .. n) (device memory, capacity `cap` words) = packed records of the list, quantised
 *                           with minmax_dev.  *n_out = n.  Items outside the format are written as 0 and counted:
 *   fe_pack_errors:         synchronises and returns that count (0 = every record is valid). */
int fe_items_minmax_device(fe_ctx* ctx, double* minmax_dev);
int fe_pack_items_device(fe_ctx* ctx, uint32_t t_max, int bits_s, int bits_o, const double* minmax_dev, uint64_t* packed_dev,
                         size_t cap, size_t* n_out);
int fe_pack_errors(fe_ctx* ctx, uint32_t* n_bad);
/* Inverse: contrast = Quantizer::value(q_s), brightness = Quantizer::value(q_o) (fma: q*step+min fused as an
 * FMA-contracting build of the reference would), distance = 0.  The result feeds fe_decode. */
int fe_unpack_items(fe_ctx* ctx, const uint64_t* packed, size_t n, uint32_t t_max, int bits_s, int bits_o,
                    const double minmax[4], int fma, fe_encode_item* items_out);

/* Host-only (no GPU, no ctx): how a threshold search of T x T range blocks against S x S domain blocks is planned.
 * thr16 = the largest integer n16 = 16 * SSE whose reference distance double(float(n16 / 16)) / (S * S)
 * (image/metrics.h:38-49) is <= rms_threshold (use_threshold = 0 when none is).  A candidate with n16 <= thr16 has
 * |sum(4 r) - sum(D)| <= radius (Cauchy-Schwarz over the T*T pixels), so with blocks keyed by sum / bin_width it lies
 * within bin_span bins of the range block's bin; n_bins = 0 when the bins would not prune (then they are not used). */
typedef struct {
    int32_t use_threshold;
    uint32_t thr16;
    uint64_t radius;
    uint32_t bin_width, n_bins, bin_span;
} fe_threshold_plan;
int fe_plan_threshold(double rms_threshold, uint32_t S, uint32_t T, fe_threshold_plan* out);

int fe_get_stats(const fe_ctx* ctx, fe_stats* out);
int fe_stats_reset(fe_ctx* ctx);
int fe_synchronize(fe_ctx* ctx);

/* Synthetic inputs of the benchmark, generated on the device into
 * the ctx's image: kind 0 natural, 1 noise, 2 pattern. */
int fe_set_synthetic_image(fe_ctx* ctx, uint32_t width, uint32_t height, uint64_t seed, int kind);
int fe_get_image(fe_ctx* ctx, uint8_t* out, uint32_t stride); /* copy the ctx's source image back */

#ifdef __cplusplus
}
#endif
#endif /* FRACTENCODE_B200_H */
