// fe_api.cu -- the extern "C" entry points of include/fractencode_b200.h and the host-side
// orchestration of one search level / the quadtree / decode on the ctx stream.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <utility>

#include "fe_kernels.cuh"
#include "fe_plan.cuh"

static thread_local std::string g_create_error;   // error of the last failed fe_create on this thread

void fe_free_job(fe_ctx* ctx);

int fe_fail(fe_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_create_error = buf;
    return code;
}

#define LAUNCH(ctx, kernel, grid, block, ...)                          \
    do {                                                               \
        kernel<<<(grid), (block), 0, (ctx)->stream>>>(__VA_ARGS__);    \
        (ctx)->stats.kernel_launches++;                                \
        FE_CUDA(ctx, cudaGetLastError());                              \
    } while (0)

static inline uint32_t cdiv(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }
// block [x, x + w) x [y, y + h) inside a W x H plane, without 32-bit wrap-around (the C ABI is a trust boundary)
static inline bool inside(uint32_t x, uint32_t y, uint32_t w, uint32_t h, uint32_t W, uint32_t H) {
    return (uint64_t)x + w <= W && (uint64_t)y + h <= H;
}

// Largest n16 (16 * SSE) whose reference distance double(float(n16/16)) / (S*S) is <= thr, looked
// for in the exact regime n16 < 2^24 (SURVEY hard part 4).  Returns false when no n16 qualifies.
static bool threshold_n16(double thr, uint32_t S, uint32_t* out) {
    auto dist = [&](uint32_t n16) { return (double)(float)((double)n16 / 16.0) / (double)(S * S); };
    if (!(thr >= 0.0) || dist(0) > thr) return false;
    uint32_t lo = 0, hi = (1u << 24) - 1; // dist is monotone in n16
    if (dist(hi) <= thr) { *out = hi; return true; }
    while (hi - lo > 1) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if (dist(mid) <= thr) lo = mid; else hi = mid;
    }
    *out = lo;
    return true;
}

// Brightness bins of a threshold search (fe_plan.cu: search_level_device): radius of the Cauchy-Schwarz bound, bin width, bins, span.
void fe_plan_bins(uint32_t N, uint32_t thr16, fe_threshold_plan* pl) {
    const uint64_t nt = (uint64_t)N * thr16;
    uint64_t R = (uint64_t)std::sqrt((double)nt);
    while (R * R > nt) --R;
    while ((R + 1) * (R + 1) <= nt) ++R;
    const uint64_t maxsum = 1020ull * N;
    // bins half as wide as the radius (a range then meets 5 bins = 2.5 radii instead of 3 bins = 3 radii), as long as
    // FE_MAX_BUCKETS bins cover the value range; |sumA - sumB| <= R  =>  |binA - binB| <= floor(R / width) + 1
    const uint64_t width = std::max<uint64_t>((R + 2) / 2, (maxsum + FE_MAX_BUCKETS) / FE_MAX_BUCKETS);
    const int nbins = (int)(maxsum / width) + 1;
    const int span = (int)(R / width) + 1;
    pl->radius = R;
    pl->bin_width = (uint32_t)width;
    pl->bin_span = (uint32_t)span;
    pl->n_bins = (nbins >= 2 * (2 * span + 1) && nbins <= FE_MAX_BUCKETS) ? (uint32_t)nbins : 0u;
}

extern "C" int fe_plan_threshold(double rms_threshold, uint32_t S, uint32_t T, fe_threshold_plan* out) {
    if (!out || T < 2 || S <= T || S % T || T > 64) return FE_ERR_INVALID;
    *out = fe_threshold_plan{};
    if (rms_threshold * (double)(S * S) >= 1048576.0) return FE_ERR_UNSUPPORTED;   // fp32-rounding regime of the reference distance
    uint32_t thr16 = 0;
    if (!threshold_n16(rms_threshold, S, &thr16)) return FE_OK;
    out->use_threshold = 1;
    out->thr16 = thr16;
    fe_plan_bins(T * T, thr16, out);
    return FE_OK;
}

// -------------------------------------------------------------------------------------------------
// ctx
// -------------------------------------------------------------------------------------------------
extern "C" int fe_abi_version(void) { return FE_ABI_VERSION; }

extern "C" const char* fe_last_error(const fe_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int fe_create(fe_ctx** out, int device, void* stream) {
    if (!out) return fe_fail(nullptr, FE_ERR_INVALID, "fe_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fe_fail(nullptr, FE_ERR_NO_DEVICE, "fe_create: no CUDA device (%s); this library has no CPU path",
                       e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
    if (device < 0 || device >= count) return fe_fail(nullptr, FE_ERR_INVALID, "fe_create: device %d out of range [0,%d)", device, count);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return fe_fail(nullptr, FE_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return fe_fail(nullptr, FE_ERR_NO_DEVICE, "fe_create: device %d is sm_%d%d; kernels are built for sm_100a only", device, prop.major, prop.minor);
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fe_fail(nullptr, FE_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    fe_ctx* ctx = new fe_ctx();
    ctx->device = device;
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
    } else {
        if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
            delete ctx;
            return fe_fail(nullptr, FE_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
        }
        ctx->own_stream = true;
    }
    ctx->n_sm = prop.multiProcessorCount;
    e = cudaSuccess;
    for (auto& ev : ctx->ev) if (e == cudaSuccess) e = cudaEventCreate(&ev);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_copy, cudaEventDisableTiming);
    for (auto& ev : ctx->ev_pass) if (e == cudaSuccess) e = cudaEventCreate(&ev);
    if (e == cudaSuccess) e = cudaHostAlloc(&ctx->h_summary, 256, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        fe_fail(nullptr, FE_ERR_CUDA, "fe_create: %s", cudaGetErrorString(e));
        fe_destroy(ctx);
        return FE_ERR_CUDA;
    }
    *out = ctx;
    return FE_OK;
}

extern "C" void fe_destroy(fe_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    DevBuf* bufs[] = {&ctx->b_plan, &ctx->b_ctl, &ctx->b_list[0], &ctx->b_list[1], &ctx->b_itemrec, &ctx->b_posb, &ctx->b_summary,
                      &ctx->b_src, &ctx->b_tgt, &ctx->b_dom, &ctx->b_rng, &ctx->b_dom_cls, &ctx->b_rng_cls, &ctx->b_dom_order,
                      &ctx->b_rng_order, &ctx->b_sort_tmp, &ctx->b_keys_tmp, &ctx->b_vals_tmp, &ctx->b_A, &ctx->b_Blo, &ctx->b_Bhi,
                      &ctx->b_rowc, &ctx->b_coln, &ctx->b_rowbest, &ctx->b_rowhit, &ctx->b_hist, &ctx->b_level_items, &ctx->b_split,
                      &ctx->b_scan, &ctx->b_scan_tmp, &ctx->b_rng_next, &ctx->b_counters, &ctx->b_A16, &ctx->b_B16, &ctx->b_tmaps,
                      &ctx->b_items, &ctx->b_dec_a, &ctx->b_dec_b, &ctx->b_dec_items, &ctx->b_dec_sum, &ctx->b_q, &ctx->b_bound,
                      &ctx->b_flag_idx, &ctx->b_blob_dom, &ctx->b_tileseg, &ctx->b_dom_order2, &ctx->b_rng_order2, &ctx->b_rng2, &ctx->b_pos_of, &ctx->b_lbq, &ctx->b_lbcand, &ctx->b_cells, &ctx->b_dq[0], &ctx->b_dq[1], &ctx->b_dom_lvl[0], &ctx->b_dom_lvl[1], &ctx->b_dom_lvl[2],
                      &ctx->b_dom_lvl[3], &ctx->b_dom_lvl[4], &ctx->b_dom_lvl[5], &ctx->b_dom_lvl[6], &ctx->b_dom_lvl[7]};
    for (DevBuf* b : bufs) b->release();
    for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    if (ctx->ev_copy) cudaEventDestroy(ctx->ev_copy);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (auto& ev : ctx->ev_pass) if (ev) cudaEventDestroy(ev);
    for (auto& c : ctx->sub) { if (c) fe_destroy(c); c = nullptr; }
    fe_free_job(ctx);
    if (ctx->h_summary) cudaFreeHost(ctx->h_summary);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int fe_synchronize(fe_ctx* ctx) {
    if (!ctx) return FE_ERR_INVALID;
    FE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FE_OK;
}

extern "C" int fe_get_stats(const fe_ctx* ctx, fe_stats* out) {
    if (!ctx || !out) return FE_ERR_INVALID;
    *out = ctx->stats;
    return FE_OK;
}

extern "C" int fe_stats_reset(fe_ctx* ctx) {
    if (!ctx) return FE_ERR_INVALID;
    ctx->stats = fe_stats{};
    return FE_OK;
}

// -------------------------------------------------------------------------------------------------
// images
// -------------------------------------------------------------------------------------------------
static int upload_plane(fe_ctx* ctx, DevBuf& buf, Plane& pl, const void* px, uint32_t w, uint32_t h, uint32_t stride, cudaMemcpyKind kind) {
    if (!px || !w || !h || stride < w) return fe_fail(ctx, FE_ERR_INVALID, "image: null pixels, zero size or stride < width");
    FE_CUDA(ctx, cudaSetDevice(ctx->device));
    FE_CUDA(ctx, buf.ensure((size_t)h * stride + 64));
    FE_CUDA(ctx, cudaMemcpyAsync(buf.p, px, (size_t)h * stride, kind, ctx->stream));
    pl.px = buf.as<uint8_t>();
    pl.w = w; pl.h = h; pl.stride = stride;
    return FE_OK;
}

extern "C" int fe_set_image(fe_ctx* ctx, const uint8_t* px, uint32_t w, uint32_t h, uint32_t stride) {
    if (!ctx) return FE_ERR_INVALID;
    FE_TRY(upload_plane(ctx, ctx->b_src, ctx->src, px, w, h, stride, cudaMemcpyHostToDevice));
    ctx->tgt = ctx->src;
    return FE_OK;
}

extern "C" int fe_set_images(fe_ctx* ctx, const uint8_t* spx, uint32_t sw, uint32_t sh, uint32_t ss, const uint8_t* tpx,
                             uint32_t tw, uint32_t th, uint32_t ts) {
    if (!ctx) return FE_ERR_INVALID;
    FE_TRY(upload_plane(ctx, ctx->b_src, ctx->src, spx, sw, sh, ss, cudaMemcpyHostToDevice));
    FE_TRY(upload_plane(ctx, ctx->b_tgt, ctx->tgt, tpx, tw, th, ts, cudaMemcpyHostToDevice));
    return FE_OK;
}

extern "C" int fe_set_image_device(fe_ctx* ctx, const void* dpx, uint32_t w, uint32_t h, uint32_t stride) {
    if (!ctx) return FE_ERR_INVALID;
    FE_TRY(upload_plane(ctx, ctx->b_src, ctx->src, dpx, w, h, stride, cudaMemcpyDeviceToDevice));
    ctx->tgt = ctx->src;
    return FE_OK;
}

extern "C" int fe_set_synthetic_image(fe_ctx* ctx, uint32_t w, uint32_t h, uint64_t seed, int kind) {
    if (!ctx || !w || !h || kind < 0 || kind > 2) return fe_fail(ctx, FE_ERR_INVALID, "fe_set_synthetic_image: bad arguments");
    FE_CUDA(ctx, cudaSetDevice(ctx->device));
    FE_CUDA(ctx, ctx->b_src.ensure((size_t)h * w + 64));
    ctx->src.px = ctx->b_src.as<uint8_t>();
    ctx->src.w = w; ctx->src.h = h; ctx->src.stride = w;
    ctx->tgt = ctx->src;
    dim3 block(32, 8), grid(cdiv(w, 32), cdiv(h, 8));
    LAUNCH(ctx, k_synth, grid, block, ctx->src.px, w, h, w, (unsigned long long)seed, kind);
    return FE_OK;
}

extern "C" int fe_get_image(fe_ctx* ctx, uint8_t* out, uint32_t stride) {
    if (!ctx || !out) return FE_ERR_INVALID;
    if (!ctx->src.px) return fe_fail(ctx, FE_ERR_STATE, "fe_get_image: no image set");
    if (stride < ctx->src.w) return fe_fail(ctx, FE_ERR_INVALID, "fe_get_image: stride < width");
    FE_CUDA(ctx, cudaMemcpy2DAsync(out, stride, ctx->src.px, ctx->src.stride, ctx->src.w, ctx->src.h, cudaMemcpyDeviceToHost, ctx->stream));
    FE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FE_OK;
}

// -------------------------------------------------------------------------------------------------
// classification of a host list
// -------------------------------------------------------------------------------------------------
extern "C" int fe_classify(fe_ctx* ctx, int which, const fe_grid_item* items, size_t n, int32_t* bins_out) {
    if (!ctx || !items || !bins_out) return fe_fail(ctx, FE_ERR_INVALID, "fe_classify: null argument");
    const Plane& pl = which ? ctx->tgt : ctx->src;
    if (!pl.px) return fe_fail(ctx, FE_ERR_STATE, "fe_classify: no image set");
    if (n == 0) return FE_OK;
    for (size_t i = 0; i < n; ++i)
        if (items[i].w < 2 || items[i].h < 2 || !inside(items[i].x, items[i].y, items[i].w, items[i].h, pl.w, pl.h))
            return fe_fail(ctx, FE_ERR_INVALID, "fe_classify: item %zu outside the image or smaller than 2x2", i);
    FE_CUDA(ctx, cudaSetDevice(ctx->device));
    FE_CUDA(ctx, ctx->b_dom.ensure(n * sizeof(fe_grid_item)));
    FE_CUDA(ctx, ctx->b_dom_cls.ensure(n * sizeof(int32_t)));
    FE_CUDA(ctx, cudaMemcpyAsync(ctx->b_dom.p, items, n * sizeof(fe_grid_item), cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(ctx, k_classify, cdiv(n * 32, 256), 256, pl.px, pl.stride, ctx->b_dom.as<fe_grid_item>(), (uint32_t)n, ctx->b_dom_cls.as<int32_t>(), 1);
    FE_CUDA(ctx, cudaMemcpyAsync(bins_out, ctx->b_dom_cls.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    FE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FE_OK;
}

// -------------------------------------------------------------------------------------------------
// one search level on device lists
// -------------------------------------------------------------------------------------------------
struct LevelIO {
    const fe_grid_item* d_dom = nullptr; uint32_t nD = 0;
    const fe_grid_item* d_rng = nullptr; uint32_t nR = 0;
    LevelGeom g;
    fe_encode_item* d_out = nullptr; // [nR], by range index
    uint32_t* d_split = nullptr;     // [nR] or NULL
    int can_split = 0;
    int lattice = 0;                 // the blocks lie on the quadtree's lattices (range origins multiples of T, domain origins of T)
    uint32_t dnx = 0;                // lattice: domains per row, domain d has its origin at (T (d % dnx), T (d / dnx))
    int stat_level = -1;             // quadtree level index for the per-level stats, -1 = none
};

// Stable bucketing of n items by class (7 buckets: -1, 0..5).  order[pos] = item index, off[c]..off[c+1] = bucket c.
static int bucket_by_class(fe_ctx* ctx, const int32_t* d_cls, uint32_t n, DevBuf& order, uint32_t off[8]) {
    FE_CUDA(ctx, order.ensure((size_t)n * sizeof(uint32_t)));
    FE_CUDA(ctx, ctx->b_keys_tmp.ensure((size_t)n * 2 + 64));
    FE_CUDA(ctx, ctx->b_vals_tmp.ensure((size_t)n * sizeof(uint32_t)));
    FE_CUDA(ctx, ctx->b_hist.ensure(8 * sizeof(uint32_t)));
    uint8_t* keys_in = ctx->b_keys_tmp.as<uint8_t>();
    uint8_t* keys_out = keys_in + n;
    FE_CUDA(ctx, cudaMemsetAsync(ctx->b_hist.p, 0, 8 * sizeof(uint32_t), ctx->stream));
    LAUNCH(ctx, k_class_keys, cdiv(n, 256), 256, d_cls, n, keys_in, ctx->b_hist.as<uint32_t>());
    LAUNCH(ctx, k_iota, cdiv(n, 256), 256, ctx->b_vals_tmp.as<uint32_t>(), n);
    size_t tmp_bytes = 0;
    FE_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in, keys_out, ctx->b_vals_tmp.as<uint32_t>(), order.as<uint32_t>(), (int)n, 0, 3, ctx->stream));
    FE_CUDA(ctx, ctx->b_sort_tmp.ensure(tmp_bytes));
    FE_CUDA(ctx, cub::DeviceRadixSort::SortPairs(ctx->b_sort_tmp.p, tmp_bytes, keys_in, keys_out, ctx->b_vals_tmp.as<uint32_t>(), order.as<uint32_t>(), (int)n, 0, 3, ctx->stream));
    ctx->stats.kernel_launches += 3; // cub: histogram + scan + one onesweep pass
    uint32_t hist[8];
    FE_CUDA(ctx, cudaMemcpyAsync(hist, ctx->b_hist.p, sizeof(hist), cudaMemcpyDeviceToHost, ctx->stream));
    FE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    off[0] = 0;
    for (int c = 0; c < 7; ++c) off[c + 1] = off[c] + hist[c];
    return FE_OK;
}

static int launch_finalize(fe_ctx* ctx, const FinalizeArgs& f, uint32_t T) {
    launch_k_finalize(ctx->stream, f, T);
    ctx->stats.kernel_launches++;
    FE_CUDA(ctx, cudaGetLastError());
    return FE_OK;
}

// fp32-regime re-rank (SURVEY hard part 2): ranges whose best SSE is >= 2^20 are re-scored with the reference's
// rounded fp32 running sum over every candidate inside the rounding band of the exact minimum.  Rare (noise-like or
// saturated blocks at T >= 16); runs the exact integer kernel on the flagged ranges only.
static int rerank_fp32_regime(fe_ctx* ctx, const LevelIO& io, const fe_params& p, const uint32_t* dom_order, const uint32_t* rng_order,
                              const uint32_t doff[8], const uint32_t roff[8], int nbuckets) {
    const LevelGeom& g = io.g;
    const uint32_t nR = io.nR, nD = io.nD;
    std::vector<uint32_t> bound(nR), order;
    FE_CUDA(ctx, cudaMemcpyAsync(bound.data(), ctx->b_bound.p, (size_t)nR * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (rng_order) {
        order.resize(nR);
        FE_CUDA(ctx, cudaMemcpyAsync(order.data(), rng_order, (size_t)nR * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    FE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<uint32_t> flag_idx, flag_bound;
    uint32_t foff[8] = {0};
    for (int c = 0; c < nbuckets; ++c) {
        for (uint32_t j = roff[c]; j < roff[c + 1]; ++j)
            if (bound[j]) {
                flag_idx.push_back(rng_order ? order[j] : j);
                flag_bound.push_back(bound[j]);
            }
        foff[c + 1] = (uint32_t)flag_idx.size();
    }
    const uint32_t nF = (uint32_t)flag_idx.size();
    if (!nF) return FE_OK;
    const uint32_t npool = g.fast ? 1u : 4u;
    FE_CUDA(ctx, ctx->b_flag_idx.ensure((size_t)nF * 8 + 16));
    uint32_t* d_idx = ctx->b_flag_idx.as<uint32_t>();
    uint32_t* d_bound = d_idx + nF;
    FE_CUDA(ctx, cudaMemcpyAsync(d_idx, flag_idx.data(), (size_t)nF * 4, cudaMemcpyHostToDevice, ctx->stream));
    FE_CUDA(ctx, cudaMemcpyAsync(d_bound, flag_bound.data(), (size_t)nF * 4, cudaMemcpyHostToDevice, ctx->stream));
    // exact-path operands: rows of the flagged ranges only, the whole (sorted) pool
    FE_CUDA(ctx, ctx->b_A.ensure((size_t)nF * 4 * g.Npad));
    FE_CUDA(ctx, ctx->b_rowc.ensure((size_t)std::max(nF, nR) * 4));
    FE_CUDA(ctx, ctx->b_Blo.ensure((size_t)nD * npool * g.Npad));
    FE_CUDA(ctx, ctx->b_Bhi.ensure((size_t)nD * npool * g.Npad));
    FE_CUDA(ctx, ctx->b_coln.ensure((size_t)nD * npool * 4));
    LAUNCH(ctx, k_build_rows, cdiv((uint64_t)nF * 32, 256), 256, ctx->tgt.px, ctx->tgt.stride, io.d_rng, d_idx, nF, g.T, g.Npad, g.fast ? 1 : 0,
           ctx->b_A.as<uint8_t>(), ctx->b_rowc.as<uint32_t>());
    LAUNCH(ctx, k_build_pool, cdiv((uint64_t)nD * npool * 32, 256), 256, ctx->src.px, ctx->src.stride, io.d_dom, dom_order, nD, npool, g.T, g.rho,
           g.Npad, ctx->b_Blo.as<uint8_t>(), ctx->b_Bhi.as<uint8_t>(), ctx->b_coln.as<uint32_t>());
    LAUNCH(ctx, k_fill_u64, cdiv((uint64_t)nF * 4, 256), 256, ctx->b_rowbest.as<unsigned long long>(), FE_INF64, (size_t)nF * 4);
    for (int c = 0; c < nbuckets; ++c) {
        const uint32_t rc = foff[c + 1] - foff[c], dc = doff[c + 1] - doff[c];
        if (!rc || !dc) continue;
        SearchArgs a{};
        a.A = ctx->b_A.as<uint8_t>();
        a.Blo = ctx->b_Blo.as<uint8_t>();
        a.Bhi = ctx->b_Bhi.as<uint8_t>();
        a.rowc = ctx->b_rowc.as<uint32_t>();
        a.coln = ctx->b_coln.as<uint32_t>();
        a.rowbest = ctx->b_rowbest.as<unsigned long long>();
        a.rowhit = ctx->b_rowhit.as<uint32_t>();
        a.row0 = foff[c] * 4; a.nrows = rc * 4;
        a.col0 = doff[c]; a.ncols = dc;
        a.Npad = g.Npad;
        a.pool_stride_cols = g.fast ? 0 : nD;
        a.rowbound = d_bound;
        a.src = ctx->src.px; a.src_stride = ctx->src.stride;
        a.tgt = ctx->tgt.px; a.tgt_stride = ctx->tgt.stride;
        a.dom = io.d_dom; a.dom_order = dom_order;
        a.rng = io.d_rng; a.row_range = d_idx;
        a.rho = g.rho;
        FE_CUDA(ctx, launch_search_exact(ctx, a, true));
    }
    FinalizeArgs f{};
    f.src = ctx->src.px; f.src_stride = ctx->src.stride;
    f.tgt = ctx->tgt.px; f.tgt_stride = ctx->tgt.stride;
    f.dom = io.d_dom; f.rng = io.d_rng;
    f.dom_order = dom_order; f.rng_order = d_idx;
    f.rowbest = ctx->b_rowbest.as<unsigned long long>();
    f.rowhit = ctx->b_rowhit.as<uint32_t>();
    f.n = nF;
    f.use_thr = 0;
    f.thr = p.rms_threshold; f.s_max = p.s_max; f.fma = p.fma;
    f.can_split = io.can_split;
    f.out = io.d_out; f.split = io.d_split;
    f.mismatch = ctx->b_counters.as<uint32_t>() + 8;
    f.fp32_regime = ctx->b_counters.as<uint32_t>() + 9;
    f.bound_out = nullptr;
    f.rerank = 1;
    FE_TRY(launch_finalize(ctx, f, g.T));
    FE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FE_OK;
}

// One level on the exact integer path (CUDA cores, dp4a): u8 rows, low/high byte pools.  Every geometry the tensor paths do
// not take -- rho != 2, odd domain origins, T = 2 or 64 -- and FE_SEARCH_EXACT.  Runs to completion (synchronises).
static int run_level_exact(fe_ctx* ctx, const LevelIO& io, const fe_params& p) {
    const LevelGeom& g = io.g;
    const uint32_t nR = io.nR, nD = io.nD;
    if (nR == 0) return FE_OK;
    const bool timed = io.stat_level >= 0 && io.stat_level < 8;
    if (timed) cudaEventRecord(ctx->ev[0], ctx->stream);
    FE_CUDA(ctx, ctx->b_counters.ensure(16 * sizeof(uint32_t)));
    FE_CUDA(ctx, cudaMemsetAsync(ctx->b_counters.p, 0, 16 * sizeof(uint32_t), ctx->stream));

    // ---- classes and buckets ----
    uint32_t doff[8] = {0, nD, nD, nD, nD, nD, nD, nD}, roff[8] = {0, nR, nR, nR, nR, nR, nR, nR};
    const uint32_t* dom_order = nullptr;
    const uint32_t* rng_order = nullptr;
    int nbuckets = 1;
    if (p.use_classifier && nD) {
        FE_CUDA(ctx, ctx->b_dom_cls.ensure((size_t)nD * 4));
        FE_CUDA(ctx, ctx->b_rng_cls.ensure((size_t)nR * 4));
        launch_classify(ctx->stream, ctx->src.px, ctx->src.stride, io.d_dom, nD, g.S, ctx->b_dom_cls.as<int32_t>(), 0);
        launch_classify(ctx->stream, ctx->tgt.px, ctx->tgt.stride, io.d_rng, nR, g.T, ctx->b_rng_cls.as<int32_t>(), 0);
        ctx->stats.kernel_launches += 2;
        FE_CUDA(ctx, cudaGetLastError());
        FE_TRY(bucket_by_class(ctx, ctx->b_dom_cls.as<int32_t>(), nD, ctx->b_dom_order, doff));
        FE_TRY(bucket_by_class(ctx, ctx->b_rng_cls.as<int32_t>(), nR, ctx->b_rng_order, roff));
        dom_order = ctx->b_dom_order.as<uint32_t>();
        rng_order = ctx->b_rng_order.as<uint32_t>();
        nbuckets = 7;
    }
    if (p.rms_threshold * (double)(g.S * g.S) >= 1048576.0)
        return fe_fail(ctx, FE_ERR_UNSUPPORTED, "rms_threshold %g reaches the fp32-rounding regime of the reference distance (SSE >= 2^20) at S=%u", p.rms_threshold, g.S);
    uint32_t thr16 = 0;
    const bool use_thr = threshold_n16(p.rms_threshold, g.S, &thr16);
    uint64_t matches = 0;
    for (int c = 0; c < nbuckets; ++c) matches += (uint64_t)(roff[c + 1] - roff[c]) * (doff[c + 1] - doff[c]) * 4;
    FE_CUDA(ctx, ctx->b_rowbest.ensure((size_t)nR * 4 * 8));
    FE_CUDA(ctx, ctx->b_rowhit.ensure((size_t)nR * 4 * 4));
    FE_CUDA(ctx, ctx->b_rowc.ensure((size_t)nR * 4));

    const uint32_t npool = g.fast ? 1u : 4u;
    FE_CUDA(ctx, ctx->b_A.ensure((size_t)nR * 4 * g.Npad));
    LAUNCH(ctx, k_build_rows, cdiv((uint64_t)nR * 32, 256), 256, ctx->tgt.px, ctx->tgt.stride, io.d_rng, rng_order, nR, g.T, g.Npad,
           g.fast ? 1 : 0, ctx->b_A.as<uint8_t>(), ctx->b_rowc.as<uint32_t>());
    if (nD) {
        FE_CUDA(ctx, ctx->b_Blo.ensure((size_t)nD * npool * g.Npad));
        FE_CUDA(ctx, ctx->b_Bhi.ensure((size_t)nD * npool * g.Npad));
        FE_CUDA(ctx, ctx->b_coln.ensure((size_t)nD * npool * 4));
        LAUNCH(ctx, k_build_pool, cdiv((uint64_t)nD * npool * 32, 256), 256, ctx->src.px, ctx->src.stride, io.d_dom, dom_order, nD, npool,
               g.T, g.rho, g.Npad, ctx->b_Blo.as<uint8_t>(), ctx->b_Bhi.as<uint8_t>(), ctx->b_coln.as<uint32_t>());
    }
    LAUNCH(ctx, k_fill_u64, cdiv((uint64_t)nR * 4, 256), 256, ctx->b_rowbest.as<unsigned long long>(), FE_INF64, (size_t)nR * 4);
    LAUNCH(ctx, k_fill_u32, cdiv((uint64_t)nR * 4, 256), 256, ctx->b_rowhit.as<uint32_t>(), FE_NONE32, (size_t)nR * 4);
    if (timed) cudaEventRecord(ctx->ev[1], ctx->stream);
    for (int c = 0; c < nbuckets; ++c) {
        const uint32_t rc = roff[c + 1] - roff[c], dc = doff[c + 1] - doff[c];
        if (!rc || !dc) continue;
        SearchArgs a{};
        a.A = ctx->b_A.as<uint8_t>();
        a.Blo = ctx->b_Blo.as<uint8_t>();
        a.Bhi = ctx->b_Bhi.as<uint8_t>();
        a.rowc = ctx->b_rowc.as<uint32_t>();
        a.coln = ctx->b_coln.as<uint32_t>();
        a.rowbest = ctx->b_rowbest.as<unsigned long long>();
        a.rowhit = ctx->b_rowhit.as<uint32_t>();
        a.row0 = roff[c] * 4; a.nrows = rc * 4;
        a.col0 = doff[c]; a.ncols = dc;
        a.Npad = g.Npad;
        a.pool_stride_cols = g.fast ? 0 : nD;
        a.thr16 = thr16;
        a.use_thr = use_thr ? 1u : 0u;
        FE_CUDA(ctx, launch_search_exact(ctx, a));
    }
    ctx->stats.exact_levels++;
    ctx->stats.matches += matches;
    ctx->stats.evaluated += matches;
    if (timed) cudaEventRecord(ctx->ev[2], ctx->stream);

    // ---- winners ----
    FinalizeArgs f{};
    f.src = ctx->src.px; f.src_stride = ctx->src.stride;
    f.tgt = ctx->tgt.px; f.tgt_stride = ctx->tgt.stride;
    f.dom = io.d_dom; f.rng = io.d_rng;
    f.dom_order = dom_order; f.rng_order = rng_order;
    f.rowbest = ctx->b_rowbest.as<unsigned long long>();
    f.rowhit = ctx->b_rowhit.as<uint32_t>();
    f.n = nR;
    f.use_thr = use_thr ? 1u : 0u;
    f.thr = p.rms_threshold; f.s_max = p.s_max; f.fma = p.fma;
    f.can_split = io.can_split;
    f.thr16 = thr16;
    f.out = io.d_out; f.split = io.d_split;
    f.mismatch = ctx->b_counters.as<uint32_t>();
    f.fp32_regime = ctx->b_counters.as<uint32_t>() + 1;
    FE_CUDA(ctx, ctx->b_bound.ensure((size_t)nR * 4 + 4));
    f.bound_out = ctx->b_bound.as<uint32_t>();
    f.rerank = 0;
    f.hit_is_domain = 0;                 // the exact path reports sorted column positions
    f.no_min = 0;
    FE_TRY(launch_finalize(ctx, f, g.T));
    if (timed) cudaEventRecord(ctx->ev[3], ctx->stream);

    uint32_t counters[4];
    FE_CUDA(ctx, cudaMemcpyAsync(counters, ctx->b_counters.p, sizeof(counters), cudaMemcpyDeviceToHost, ctx->stream));
    FE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (counters[0]) return fe_fail(ctx, FE_ERR_CUDA, "internal: %u winners whose search score disagrees with the direct recomputation (T=%u)", counters[0], g.T);
    ctx->stats.fp32_regime_items += counters[1];
    if (counters[1]) FE_TRY(rerank_fp32_regime(ctx, io, p, dom_order, rng_order, doff, roff, nbuckets));
    if (timed) {
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]); ctx->stats.level_prep_ms[io.stat_level] = ms;
        cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]); ctx->stats.level_search_ms[io.stat_level] = ms;
        ctx->stats.level_ranges[io.stat_level] = nR;
        ctx->stats.level_matches[io.stat_level] = matches;
        ctx->stats.level_evaluated[io.stat_level] = matches;
        ctx->stats.level_passes[io.stat_level] = 1;
    }
    return FE_OK;
}

// ---------------------------------------------------------------------------------------------------
// A level in two halves: everything is enqueued, then ONE synchronisation reads the level's summary back (together with
// the quadtree's split count).  Levels of the device-scheduled tcgen05 path (fe_plan.cu) never synchronise in between;
// the exact integer path (generic geometries) still runs to completion inside the first half.
// ---------------------------------------------------------------------------------------------------
struct LevelPending {
    bool device = false;
    int kind = 0;                         // 0: kind::f16 (T = 4, 8), 1: kind::i8, 2: lower-bound prefilter + exact check
    int skip = 0;
    DeviceLevelState st;
    uint32_t thr16 = 0;
    bool use_thr = false, timed = false;
    FinalizeArgs fin{};
};

static inline int hint_slot(uint32_t T) { int l = 0; while ((1u << l) < T && l < 7) ++l; return l; }

// skip: bit 0 = not the kind::f16 search (a winner sat in its inexact band), bit 1 = not the lower-bound prefilter (its
// candidate list overflowed) -- set when a level is redone
static int run_level_enqueue(fe_ctx* ctx, const LevelIO& io, const fe_params& p, LevelPending* lp, int skip = 0) {
    const LevelGeom& g = io.g;
    const uint32_t nR = io.nR, nD = io.nD;
    *lp = LevelPending{};
    if (nR == 0) return FE_OK;
    const bool f16_ok = nD && f16_level_supported(g) && !(skip & 1), i8_ok = nD && i8_level_supported(g);
    const bool device = (f16_ok || i8_ok) && p.search_impl != FE_SEARCH_EXACT;
    const bool flips = p.isometries == 8;
    if (p.isometries != 0 && p.isometries != 4 && p.isometries != 8) return fe_fail(ctx, FE_ERR_INVALID, "isometries must be 0, 4 or 8 (got %d)", p.isometries);
    if (!device) {
        if (p.search_impl == FE_SEARCH_UMMA && nD)
            return fe_fail(ctx, FE_ERR_UNSUPPORTED, "tcgen05 paths need S == 2T, even domain origins and T <= 32 (got S=%u T=%u)", g.S, g.T);
        if (flips && nD) return fe_fail(ctx, FE_ERR_UNSUPPORTED, "the flip isometries run on the tensor paths only (S == 2T, even domain origins, 4 <= T <= 32)");
        return run_level_exact(ctx, io, p);
    }
    lp->kind = f16_ok ? 0 : 1;
    lp->skip = skip;
    if (p.rms_threshold * (double)(g.S * g.S) >= 1048576.0)
        return fe_fail(ctx, FE_ERR_UNSUPPORTED, "rms_threshold %g reaches the fp32-rounding regime of the reference distance (SSE >= 2^20) at S=%u", p.rms_threshold, g.S);
    lp->device = true;
    lp->timed = io.stat_level >= 0 && io.stat_level < 8;
    lp->use_thr = threshold_n16(p.rms_threshold, g.S, &lp->thr16);
    // Large blocks of a level that only looks for hits (it splits the range blocks that find none): lower-bound prefilter on
    // the 8 x 8 grid of cell sums + exact check of the few candidates it lets through (fe_lb.cu)
    {
        const char* mt = getenv("FE_LB_MIN_T");
        const uint32_t lb_min_t = mt ? (uint32_t)atoi(mt) : 32u;
        if (lp->kind == 1 && !(skip & 2) && g.T >= lb_min_t && g.T >= 16 && lp->use_thr && io.can_split && !flips && ctx->src.px == ctx->tgt.px &&
            ctx->src.w % (g.T / 8) == 0 && ctx->src.h % (g.T / 8) == 0 && io.lattice && !getenv("FE_NO_LB"))
            lp->kind = 2;
    }
    if (lp->timed) cudaEventRecord(ctx->ev[0], ctx->stream);
    ctx->level_host_t0 = std::chrono::steady_clock::now();
    FE_CUDA(ctx, ctx->b_counters.ensure(16 * sizeof(uint32_t)));
    FE_CUDA(ctx, cudaMemsetAsync(ctx->b_counters.p, 0, 16 * sizeof(uint32_t), ctx->stream));
    // flip isometries: every range block is searched twice, the second copy mirrored left-right (its four rotation rows are
    // the four flip isometries of the block); the list the search sees has 2 nR entries
    const uint32_t nS = flips ? 2 * nR : nR;
    const fe_grid_item* d_rng = io.d_rng;
    if (flips) {
        FE_CUDA(ctx, ctx->b_rng2.ensure((size_t)nS * sizeof(fe_grid_item)));
        LAUNCH(ctx, k_dup_items, cdiv(nS, 256), 256, io.d_rng, nR, ctx->b_rng2.as<fe_grid_item>());
        d_rng = ctx->b_rng2.as<fe_grid_item>();
    }
    FE_CUDA(ctx, ctx->b_rowbest.ensure((size_t)nS * 4 * 8));
    FE_CUDA(ctx, ctx->b_rowhit.ensure((size_t)nS * 4 * 4));
    DeviceLevel lv{};
    lv.d_dom = io.d_dom; lv.nD = nD; lv.d_rng = d_rng; lv.nR = nS; lv.g = g; lv.flips = flips;
    // lattice levels: the image's T x T cell sums give every domain its class and its brightness bin (fe_kernels.cu)
    if (io.lattice && io.dnx && g.S == 2 * g.T && (p.use_classifier || lp->use_thr) && cell_grid_supported(ctx->src.px, ctx->src.stride, ctx->src.w, ctx->src.h, g.T) && !getenv("FE_NO_CELLS")) {
        const size_t ncell = (size_t)(ctx->src.w / g.T) * (ctx->src.h / g.T);
        FE_CUDA(ctx, ctx->b_cells.ensure(ncell * 12));
        launch_cell_grid(ctx->stream, ctx->src.px, ctx->src.stride, ctx->src.w, ctx->src.h, g.T, ctx->b_cells.as<uint32_t>(), ctx->b_cells.as<uint32_t>() + ncell,
                         ctx->b_cells.as<uint32_t>() + 2 * ncell);
        lv.cellsD2 = ctx->b_cells.as<uint32_t>() + 2 * ncell;
        ctx->stats.kernel_launches++;
        lv.cells = ctx->b_cells.as<uint32_t>(); lv.cells_w = ctx->src.w / g.T; lv.dnx = io.dnx;
        if (ctx->src.px == ctx->tgt.px && !flips) lv.cells2 = lv.cells + ncell;
    }
    if (p.use_classifier) {
        FE_CUDA(ctx, ctx->b_dom_cls.ensure((size_t)nD * 4));
        FE_CUDA(ctx, ctx->b_rng_cls.ensure((size_t)nS * 4));
        if (lv.cells) launch_dom_from_cells(ctx->stream, lv.cells, lv.cells_w, lv.dnx, nD, ctx->b_dom_cls.as<int32_t>(), 1u, nullptr, nullptr);
        else launch_classify(ctx->stream, ctx->src.px, ctx->src.stride, io.d_dom, nD, g.S, ctx->b_dom_cls.as<int32_t>(), 0);
        launch_classify(ctx->stream, ctx->tgt.px, ctx->tgt.stride, d_rng, nS, g.T, ctx->b_rng_cls.as<int32_t>(), 0);
        ctx->stats.kernel_launches += 2;
        FE_CUDA(ctx, cudaGetLastError());
        lv.dom_cls = ctx->b_dom_cls.as<int32_t>();
        lv.rng_cls = ctx->b_rng_cls.as<int32_t>();
    }
    lv.thr16 = lp->thr16; lv.use_thr = lp->use_thr; lv.need_min = !io.can_split; lv.timed = lp->timed;
    // How many slices the level will need is only known on the device.  The host enqueues as many as the same kind of level
    // needed last time (a launch past the end costs a few microseconds, a missing one a second synchronisation) -- the whole
    // train when nothing is known yet.
    const fe_ctx::SliceHint& hint = ctx->hint[lp->kind][hint_slot(g.T)];
    const bool use_hint = hint.known && !getenv("FE_NO_HINT");
    FE_TRY(search_level_device(ctx, lv, lp->kind, use_hint ? hint.slices : 0u, use_hint ? hint.with_min != 0 : true, &lp->st));
    if (lp->timed) cudaEventRecord(ctx->ev[2], ctx->stream);
    FinalizeArgs& f = lp->fin;
    f.src = ctx->src.px; f.src_stride = ctx->src.stride;
    f.tgt = ctx->tgt.px; f.tgt_stride = ctx->tgt.stride;
    f.dom = io.d_dom; f.rng = d_rng;
    f.dom_order = nullptr; f.rng_order = lp->st.rng_order;        // keys and hits are domain indices
    if (flips) {
        FE_CUDA(ctx, ctx->b_pos_of.ensure((size_t)nS * 4));
        LAUNCH(ctx, k_pos_of, cdiv(nS, 256), 256, lp->st.rng_order, nS, ctx->b_pos_of.as<uint32_t>());
        f.flips = 1;
        f.pos_of = ctx->b_pos_of.as<uint32_t>();
    }
    f.rowbest = ctx->b_rowbest.as<unsigned long long>();
    f.rowhit = ctx->b_rowhit.as<uint32_t>();
    f.n = nR;
    f.use_thr = lp->use_thr ? 1u : 0u;
    f.thr = p.rms_threshold; f.s_max = p.s_max; f.fma = p.fma;
    f.can_split = io.can_split;
    f.thr16 = lp->thr16;
    f.out = io.d_out; f.split = io.d_split;
    f.mismatch = ctx->b_counters.as<uint32_t>();
    f.fp32_regime = ctx->b_counters.as<uint32_t>() + 1;
    FE_CUDA(ctx, ctx->b_bound.ensure((size_t)nR * 4 + 4));
    f.bound_out = ctx->b_bound.as<uint32_t>();
    f.rerank = 0;
    f.hit_is_domain = 1;
    f.no_min = (lp->use_thr && io.can_split) ? 1 : 0;   // a range without a hit splits: its minimum was not even tracked
    FE_TRY(launch_finalize(ctx, f, g.T));
    if (lp->timed) cudaEventRecord(ctx->ev[3], ctx->stream);
    return FE_OK;
}

// Second half: the quadtree's scan of the split flags (do_scan), ONE D2H copy + synchronisation for the level's summary
// and the split count, and the rare follow-ups (more slices than were enqueued, the inexact fp16 band, the fp32-regime
// re-rank), each of which costs another round.
static int run_level_complete(fe_ctx* ctx, const LevelIO& io, const fe_params& p, LevelPending* lp, bool do_scan, size_t* n_split) {
    if (n_split) *n_split = 0;
    if (io.nR == 0) return FE_OK;
    do_scan = do_scan && io.can_split;
    if (!lp->device && !do_scan) return FE_OK;
    FE_CUDA(ctx, ctx->b_summary.ensure(sizeof(LevelSummary)));
    FE_CUDA(ctx, ctx->b_counters.ensure(16 * sizeof(uint32_t)));
    if (do_scan) FE_CUDA(ctx, ctx->b_scan.ensure((size_t)io.nR * 4 + 4));
    LevelSummary* hs = reinterpret_cast<LevelSummary*>(ctx->h_summary);
    const LevelGeom& g = io.g;
    for (int round = 0;; ++round) {
        if (round > 4) return fe_fail(ctx, FE_ERR_CUDA, "internal: level T=%u does not settle", g.T);
        if (do_scan) {
            size_t tmp_bytes = 0;
            FE_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, io.d_split, ctx->b_scan.as<uint32_t>(), (int)io.nR, ctx->stream));
            FE_CUDA(ctx, ctx->b_scan_tmp.ensure(tmp_bytes));
            FE_CUDA(ctx, cub::DeviceScan::ExclusiveSum(ctx->b_scan_tmp.p, tmp_bytes, io.d_split, ctx->b_scan.as<uint32_t>(), (int)io.nR, ctx->stream));
            ctx->stats.kernel_launches += 2;
        }
        LAUNCH(ctx, k_level_summary, 1, 1, lp->device ? ctx->b_ctl.as<SliceCtl>() : nullptr, lp->device ? ctx->b_plan.as<LevelPlan>() : nullptr,
               ctx->b_counters.as<uint32_t>(), do_scan ? ctx->b_scan.as<uint32_t>() + (io.nR - 1) : nullptr, do_scan ? io.d_split + (io.nR - 1) : nullptr,
               lp->device && lp->st.wants_min_pass ? 1u : 0u, ctx->b_summary.as<LevelSummary>());
        FE_CUDA(ctx, cudaMemcpyAsync(hs, ctx->b_summary.p, sizeof(LevelSummary), cudaMemcpyDeviceToHost, ctx->stream));
        FE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (n_split) *n_split = (size_t)hs->last_scan + hs->last_flag;
        if (!lp->device) return FE_OK;
        if (hs->overflow) return fe_fail(ctx, FE_ERR_CUDA, "internal: work-item buffer of the level too small (T=%u)", g.T);
        if (hs->flags & 1u) {
            // a winner of the kind::f16 search sits in the fp32-inexact band: the level is searched again on the integer kind
            FE_TRY(run_level_enqueue(ctx, io, p, lp, lp->skip | 1));
            continue;
        }
        if (hs->flags & 2u) {
            // the candidate list of the lower-bound prefilter overflowed (everything matches everything): exact kind
            FE_TRY(run_level_enqueue(ctx, io, p, lp, lp->skip | 2));
            continue;
        }
        if (!hs->done) {
            // the level needed more slices (or a minimum pass) than were enqueued: the rest of the train, winners again
            FE_TRY(search_level_more(ctx, &lp->st));
            if (lp->timed) cudaEventRecord(ctx->ev[2], ctx->stream);
            FE_CUDA(ctx, cudaMemsetAsync(ctx->b_counters.p, 0, 2 * sizeof(uint32_t), ctx->stream));
            FE_TRY(launch_finalize(ctx, lp->fin, g.T));
            continue;
        }
        break;
    }
    if (hs->mismatch) return fe_fail(ctx, FE_ERR_CUDA, "internal: %u winners whose search score disagrees with the direct recomputation (T=%u)", hs->mismatch, g.T);
    fe_ctx::SliceHint& hint = ctx->hint[lp->kind][hint_slot(g.T)];
    hint.known = 1; hint.slices = (uint8_t)std::max(1u, hs->slices); hint.with_min = hs->min_ran ? 1 : 0;
    // lower-bound prefilter levels: `evaluated` = candidates scored exactly (1024 products each at T = 32); the pairs the
    // 8 x 8 bound looked at (64 products each) are counted apart
    const unsigned long long evaluated = lp->kind == 2 ? hs->lb_candidates : hs->evaluated;
    const unsigned long long prefiltered = lp->kind == 2 ? hs->evaluated : 0ull;
    ctx->stats.matches += hs->matches;
    ctx->stats.evaluated += evaluated;
    ctx->stats.prefiltered += prefiltered;
    ctx->stats.umma_levels++;
    ctx->stats.fp32_regime_items += hs->fp32_regime;
    if (lp->timed) {
        float kernel_ms = 0.f;
        for (uint32_t i = 0; i < lp->st.n_launches; ++i) {
            float ms = 0;
            cudaEventElapsedTime(&ms, ctx->ev_pass[2 * i], ctx->ev_pass[2 * i + 1]);
            kernel_ms += ms;
            if (getenv("FE_PASS_TIMES")) fprintf(stderr, "[level] T=%u search launch %u: %.3f ms\n", g.T, i, ms);
        }
        float ms = 0;
        if (getenv("FE_PASS_TIMES") && lp->st.n_launches) {       // device time from the level's first launch to its first search launch
            cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev_pass[0]);
            fprintf(stderr, "[level] T=%u device time before the first search launch %.1f us (host: %.1f us)\n", g.T, ms * 1e3, lp->st.host_us_first_search);
        }
        cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[2]);
        ctx->stats.level_search_ms[io.stat_level] = kernel_ms;
        ctx->stats.level_prep_ms[io.stat_level] = ms - kernel_ms;
        ctx->stats.level_ranges[io.stat_level] = io.nR;
        ctx->stats.level_matches[io.stat_level] = hs->matches;
        ctx->stats.level_evaluated[io.stat_level] = evaluated;
        ctx->stats.level_prefiltered[io.stat_level] = prefiltered;
        ctx->stats.level_passes[io.stat_level] = hs->passes;
    }
    if (hs->fp32_regime && lp->fin.flips)
        return fe_fail(ctx, FE_ERR_UNSUPPORTED, "flip isometries: %u winners in the fp32-rounding regime of the reference distance (SSE >= 2^20, T=%u); "
                                                "the re-rank pass covers the four rotations only", hs->fp32_regime, g.T);
    if (hs->fp32_regime) {
        // re-rank of the flagged range blocks on the exact kernel: every domain of the block's class in scan order
        const uint32_t nD = io.nD, nR = io.nR;
        if (p.use_classifier) {
            uint32_t doff[8], roff[8] = {0};
            FE_TRY(bucket_by_class(ctx, ctx->b_dom_cls.as<int32_t>(), nD, ctx->b_dom_order, doff));
            std::vector<uint32_t> lr(FE_MAX_TOTAL + 1);
            FE_CUDA(ctx, cudaMemcpy(lr.data(), reinterpret_cast<const uint8_t*>(ctx->b_plan.p) + offsetof(LevelPlan, roff), lr.size() * 4, cudaMemcpyDeviceToHost));
            for (int c = 0; c <= 7; ++c) roff[c] = lr[(size_t)c * lp->st.nbins];
            FE_TRY(rerank_fp32_regime(ctx, io, p, ctx->b_dom_order.as<uint32_t>(), lp->st.rng_order, doff, roff, 7));
        } else {
            const uint32_t d1[8] = {0, nD, nD, nD, nD, nD, nD, nD}, r1[8] = {0, nR, nR, nR, nR, nR, nR, nR};
            FE_TRY(rerank_fp32_regime(ctx, io, p, nullptr, lp->st.rng_order, d1, r1, 1));
        }
        if (do_scan) {   // the re-rank rewrote items and split flags: scan again
            LevelPending none;
            FE_TRY(run_level_complete(ctx, io, p, &none, true, n_split));
        }
    }
    return FE_OK;
}

static int run_level(fe_ctx* ctx, const LevelIO& io, const fe_params& p) {
    LevelPending lp;
    FE_TRY(run_level_enqueue(ctx, io, p, &lp));
    return run_level_complete(ctx, io, p, &lp, false, nullptr);
}

static int make_geom(fe_ctx* ctx, uint32_t S, uint32_t T, bool even_origins, LevelGeom* g) {
    if (T < 2 || S <= T || S % T) return fe_fail(ctx, FE_ERR_UNSUPPORTED, "geometry: need S a multiple of T, S > T >= 2 (got S=%u T=%u)", S, T);
    if (T > 64) return fe_fail(ctx, FE_ERR_UNSUPPORTED, "geometry: range size %u > 64 (32-bit score arithmetic)", T);
    g->S = S; g->T = T; g->rho = S / T;
    g->N = T * T;
    g->Npad = (g->N + 15u) & ~15u;
    g->fast = (g->rho == 2) && even_origins;
    return FE_OK;
}

extern "C" int fe_encode_level(fe_ctx* ctx, const fe_grid_item* domains, size_t n_dom, const fe_grid_item* ranges, size_t n_rng,
                               const fe_params* params, fe_encode_item* out) {
    if (!ctx) return FE_ERR_INVALID;
    if (!params || (!ranges && n_rng) || (!domains && n_dom) || (!out && n_rng)) return fe_fail(ctx, FE_ERR_INVALID, "fe_encode_level: null argument");
    if (!ctx->src.px) return fe_fail(ctx, FE_ERR_STATE, "fe_encode_level: call fe_set_image first");
    if (n_rng == 0) return FE_OK;
    if (n_dom > 0x3FFFFFFFu || n_rng > 0x3FFFFFFFu) return fe_fail(ctx, FE_ERR_UNSUPPORTED, "fe_encode_level: list too long");
    const uint32_t T = ranges[0].w;
    for (size_t i = 0; i < n_rng; ++i) {
        const fe_grid_item& r = ranges[i];
        if (r.w != T || r.h != T) return fe_fail(ctx, FE_ERR_UNSUPPORTED, "fe_encode_level: range %zu is %ux%u, expected square %u", i, r.w, r.h, T);
        if (!inside(r.x, r.y, T, T, ctx->tgt.w, ctx->tgt.h)) return fe_fail(ctx, FE_ERR_INVALID, "fe_encode_level: range %zu outside the target image", i);
        if (params->use_classifier && (r.bin < -1 || r.bin > 5))
            return fe_fail(ctx, FE_ERR_INVALID, "fe_encode_level: range %zu has classifier bin %d; bins are -1 (not classified) or the classes 0..5 of "
                                                "BrightnessBlocksClassifier2::getCategory", i, r.bin);
    }
    FE_CUDA(ctx, cudaSetDevice(ctx->device));
    FE_CUDA(ctx, ctx->b_level_items.ensure(n_rng * sizeof(fe_encode_item)));
    FE_CUDA(ctx, ctx->b_rng.ensure(n_rng * sizeof(fe_grid_item)));
    FE_CUDA(ctx, cudaMemcpyAsync(ctx->b_rng.p, ranges, n_rng * sizeof(fe_grid_item), cudaMemcpyHostToDevice, ctx->stream));
    LevelIO io;
    io.d_rng = ctx->b_rng.as<fe_grid_item>(); io.nR = (uint32_t)n_rng;
    io.d_out = ctx->b_level_items.as<fe_encode_item>();
    if (n_dom) {
        const uint32_t S = domains[0].w;
        bool even = true;
        for (size_t i = 0; i < n_dom; ++i) {
            const fe_grid_item& d = domains[i];
            if (d.w != S || d.h != S) return fe_fail(ctx, FE_ERR_UNSUPPORTED, "fe_encode_level: domain %zu is %ux%u, expected square %u", i, d.w, d.h, S);
            if (!inside(d.x, d.y, S, S, ctx->src.w, ctx->src.h)) return fe_fail(ctx, FE_ERR_INVALID, "fe_encode_level: domain %zu outside the source image", i);
            if (params->use_classifier && (d.bin < -1 || d.bin > 5))
                return fe_fail(ctx, FE_ERR_INVALID, "fe_encode_level: domain %zu has classifier bin %d; bins are -1 (not classified) or the classes 0..5 of "
                                                    "BrightnessBlocksClassifier2::getCategory", i, d.bin);
            even = even && !(d.x & 1) && !(d.y & 1);
        }
        FE_TRY(make_geom(ctx, S, T, even, &io.g));
        FE_CUDA(ctx, ctx->b_dom.ensure(n_dom * sizeof(fe_grid_item)));
        FE_CUDA(ctx, cudaMemcpyAsync(ctx->b_dom.p, domains, n_dom * sizeof(fe_grid_item), cudaMemcpyHostToDevice, ctx->stream));
        io.d_dom = ctx->b_dom.as<fe_grid_item>(); io.nD = (uint32_t)n_dom;
    } else {
        FE_TRY(make_geom(ctx, 2 * T, T, true, &io.g));
    }
    FE_TRY(run_level(ctx, io, *params));
    FE_CUDA(ctx, cudaMemcpyAsync(out, io.d_out, n_rng * sizeof(fe_encode_item), cudaMemcpyDeviceToHost, ctx->stream));
    FE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FE_OK;
}

// -------------------------------------------------------------------------------------------------
// quadtree
// -------------------------------------------------------------------------------------------------
extern "C" int fe_encode_quadtree_device(fe_ctx* ctx, uint32_t t_max, uint32_t t_min, const fe_params* params, size_t* n_out) {
    return fe_encode_quadtree_slice_device(ctx, t_max, t_min, params, 0, (size_t)-1, n_out);
}

// The quadtree encode of one image as a resumable job: begin, then (enqueue level, complete level) until finished.  Between
// the two halves of a level the host is free -- fe_encode_batch uses that to keep a second image's level in flight.
struct QuadJob {
    uint32_t t_max = 0, t_min = 0, T = 0;
    int level = 0;
    size_t n_pending = 0, offset = 0;
    fe_params params{};
    LevelIO io;
    LevelPending lp;
    bool level_open = false;
};

static QuadJob* job_of(fe_ctx* ctx) {
    if (!ctx->job) ctx->job = new QuadJob();
    return reinterpret_cast<QuadJob*>(ctx->job);
}
void fe_free_job(fe_ctx* ctx) {
    delete reinterpret_cast<QuadJob*>(ctx->job);
    ctx->job = nullptr;
}

static int quad_begin(fe_ctx* ctx, uint32_t t_max, uint32_t t_min, const fe_params* params, size_t first_block, size_t n_blocks) {
    if (!params) return fe_fail(ctx, FE_ERR_INVALID, "fe_encode_quadtree: params is NULL");
    if (!ctx->src.px) return fe_fail(ctx, FE_ERR_STATE, "fe_encode_quadtree: call fe_set_image first");
    if (ctx->tgt.px != ctx->src.px) return fe_fail(ctx, FE_ERR_STATE, "fe_encode_quadtree: needs a single image (fe_set_image)");
    const uint32_t W = ctx->src.w, H = ctx->src.h;
    if (t_min < 2 || t_max < t_min || t_max > 64 || (t_max & (t_max - 1)) || (t_min & (t_min - 1)))
        return fe_fail(ctx, FE_ERR_UNSUPPORTED, "fe_encode_quadtree: block sizes must be powers of two, 2 <= t_min <= t_max <= 64");
    if (W % t_max || H % t_max) return fe_fail(ctx, FE_ERR_INVALID, "fe_encode_quadtree: image %ux%u not aligned to %u", W, H, t_max);
    FE_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n_top = (size_t)(W / t_max) * (H / t_max);
    if (n_blocks == (size_t)-1) n_blocks = first_block <= n_top ? n_top - first_block : 0;
    if (first_block > n_top || n_blocks > n_top - first_block)
        return fe_fail(ctx, FE_ERR_INVALID, "fe_encode_quadtree_slice: blocks %zu..+%zu outside the %zu top-level blocks", first_block, n_blocks, n_top);
    const size_t per_top = (size_t)(t_max / t_min) * (t_max / t_min);
    const size_t cap = std::max<size_t>(n_blocks * per_top, 1);
    FE_CUDA(ctx, ctx->b_items.ensure(cap * sizeof(fe_encode_item)));
    FE_CUDA(ctx, ctx->b_rng.ensure(cap * sizeof(fe_grid_item)));
    FE_CUDA(ctx, ctx->b_rng_next.ensure(cap * sizeof(fe_grid_item)));
    if (n_blocks)
        LAUNCH(ctx, k_uniform_grid, cdiv(n_blocks, 256), 256, ctx->b_rng.as<fe_grid_item>(), W / t_max, (uint32_t)n_blocks, t_max, t_max,
               (uint32_t)first_block);
    for (int l = 0; l < 8; ++l) {
        ctx->stats.level_items[l] = ctx->stats.level_ranges[l] = ctx->stats.level_matches[l] = 0;
        ctx->stats.level_evaluated[l] = ctx->stats.level_passes[l] = ctx->stats.level_prefiltered[l] = 0;
        ctx->stats.level_search_ms[l] = ctx->stats.level_prep_ms[l] = 0.f;
    }
    QuadJob* j = job_of(ctx);
    *j = QuadJob{};
    j->t_max = t_max; j->t_min = t_min; j->T = t_max;
    j->n_pending = n_blocks;
    j->params = *params;
    ctx->n_items = 0;
    return FE_OK;
}

static bool quad_finished(const QuadJob* j) { return !j->level_open && (j->T < j->t_min || j->n_pending == 0); }

// first half of the current level: grids, search, winners -- nothing here waits for the device (tensor paths)
static int quad_enqueue(fe_ctx* ctx) {
    QuadJob* j = job_of(ctx);
    if (quad_finished(j)) return FE_OK;
    const uint32_t W = ctx->src.w, H = ctx->src.h, T = j->T, S = 2 * T;
    const uint32_t dnx = W >= S ? W / T - 1 : 0, dny = H >= S ? H / T - 1 : 0;
    const size_t nD = (size_t)dnx * dny, n_pending = j->n_pending;
    LevelIO& io = j->io;
    io = LevelIO{};
    FE_TRY(make_geom(ctx, S, T, true, &io.g));
    // the level's domain grid depends on the image size only: kept per level across encodes
    DevBuf& domb = ctx->b_dom_lvl[j->level & 7];
    unsigned long long& dtag = ctx->dom_lvl_tag[j->level & 7];
    const unsigned long long want = ((unsigned long long)W << 40) ^ ((unsigned long long)H << 16) ^ ((unsigned long long)S << 8) ^ T;
    if (nD && dtag != want) {
        FE_CUDA(ctx, domb.ensure(nD * sizeof(fe_grid_item)));
        LAUNCH(ctx, k_uniform_grid, cdiv(nD, 256), 256, domb.as<fe_grid_item>(), dnx, (uint32_t)nD, S, T, 0u);
        dtag = want;
    }
    FE_CUDA(ctx, ctx->b_level_items.ensure(n_pending * sizeof(fe_encode_item)));
    FE_CUDA(ctx, ctx->b_split.ensure(n_pending * 4 + 4));
    io.d_dom = domb.as<fe_grid_item>(); io.nD = (uint32_t)nD;
    io.d_rng = ctx->b_rng.as<fe_grid_item>(); io.nR = (uint32_t)n_pending;
    io.d_out = ctx->b_level_items.as<fe_encode_item>();
    io.can_split = (T / 2 >= j->t_min) ? 1 : 0;
    io.lattice = 1;
    io.dnx = dnx;
    io.d_split = ctx->b_split.as<uint32_t>();
    io.stat_level = j->level;
    const auto h0 = std::chrono::steady_clock::now();
    FE_TRY(run_level_enqueue(ctx, io, j->params, &j->lp));
    if (getenv("FE_PASS_TIMES"))
        fprintf(stderr, "[level] T=%u host enqueue %.1f us\n", T, std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - h0).count());
    j->level_open = true;
    return FE_OK;
}

// second half: ONE synchronisation (level summary + split count), then the children / the emitted items
static int quad_complete(fe_ctx* ctx) {
    QuadJob* j = job_of(ctx);
    if (!j->level_open) return FE_OK;
    const LevelIO& io = j->io;
    const size_t n_pending = j->n_pending, offset = j->offset;
    size_t n_split = 0;
    FE_TRY(run_level_complete(ctx, io, j->params, &j->lp, true, &n_split));
    if (io.can_split) {
        LAUNCH(ctx, k_quadtree_scatter, cdiv(n_pending, 256), 256, ctx->b_rng.as<fe_grid_item>(), io.d_out, io.d_split, ctx->b_scan.as<uint32_t>(),
               (uint32_t)n_pending, ctx->b_rng_next.as<fe_grid_item>(), ctx->b_items.as<fe_encode_item>() + offset);
    } else {
        FE_CUDA(ctx, cudaMemcpyAsync(ctx->b_items.as<fe_encode_item>() + offset, io.d_out, n_pending * sizeof(fe_encode_item), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    const size_t kept = n_pending - n_split;
    ctx->stats.level_items[j->level] = kept;
    if (ctx->host_out && kept && offset + kept <= ctx->host_cap && ctx->host_copied == offset) {
        // the level's items are final: send them to the caller's buffer behind the next level's work
        FE_CUDA(ctx, cudaEventRecord(ctx->ev_copy, ctx->stream));
        FE_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_copy, 0));
        FE_CUDA(ctx, cudaMemcpyAsync(ctx->host_out + offset, ctx->b_items.as<fe_encode_item>() + offset, kept * sizeof(fe_encode_item),
                                     cudaMemcpyDeviceToHost, ctx->copy_stream));
        ctx->host_copied = offset + kept;
    }
    j->offset = offset + kept;
    std::swap(ctx->b_rng, ctx->b_rng_next);
    j->n_pending = 4 * n_split;
    j->T /= 2;
    j->level += 1;
    j->level_open = false;
    ctx->n_items = j->offset;
    return FE_OK;
}

extern "C" int fe_encode_quadtree_slice_device(fe_ctx* ctx, uint32_t t_max, uint32_t t_min, const fe_params* params, size_t first_block,
                                               size_t n_blocks, size_t* n_out) {
    if (!ctx) return FE_ERR_INVALID;
    FE_TRY(quad_begin(ctx, t_max, t_min, params, first_block, n_blocks));
    QuadJob* j = job_of(ctx);
    while (!quad_finished(j)) {
        FE_TRY(quad_enqueue(ctx));
        FE_TRY(quad_complete(ctx));
    }
    if (n_out) *n_out = ctx->n_items;
    return FE_OK;
}

// Batch mode (BASELINE config 5; the reference encodes one plane after the other, main.cpp:142-181,193-200): the images
// alternate between two child contexts with their own streams and scratch.  A level is enqueued without waiting and
// completed with one synchronisation, so while the host waits for image A's level, image B's level (and the H2D of its
// pixels, the D2H of a finished list) is already queued on the other stream: the GPU never idles on the host.
extern "C" int fe_encode_planes(fe_ctx* ctx, const uint8_t* const* images, size_t n_images, const uint32_t* widths, const uint32_t* heights,
                                const uint32_t* strides, uint32_t t_max, uint32_t t_min, const fe_params* params, fe_encode_item* out,
                                const size_t* out_offsets, const size_t* caps, size_t* n_out) {
    if (!ctx) return FE_ERR_INVALID;
    if (!params || (n_images && (!images || !out || !widths || !heights || !strides || !out_offsets || !caps)) || !n_out)
        return fe_fail(ctx, FE_ERR_INVALID, "fe_encode_planes: null argument");
    for (size_t i = 0; i < n_images; ++i)
        if (!widths[i] || !heights[i] || strides[i] < widths[i]) return fe_fail(ctx, FE_ERR_INVALID, "fe_encode_planes: plane %zu: zero size or stride < width", i);
    for (int k = 0; k < 2; ++k)
        if (!ctx->sub[k]) {
            const int rc = fe_create(&ctx->sub[k], ctx->device, nullptr);
            if (rc != FE_OK) return fe_fail(ctx, rc, "fe_encode_planes: %s", fe_last_error(nullptr));
        }
    size_t next = 0;
    int image_of[2] = {-1, -1};
    auto fail_from = [&](int k, int rc) { return fe_fail(ctx, rc, "fe_encode_planes (plane %d): %s", image_of[k], ctx->sub[k]->err.c_str()); };
    auto start_next = [&](int k) -> int {
        image_of[k] = -1;
        if (next >= n_images) return FE_OK;
        fe_ctx* c = ctx->sub[k];
        image_of[k] = (int)next;
        if (!images[next]) { c->err = "null image"; return FE_ERR_INVALID; }
        int rc = fe_set_image(c, images[next], widths[next], heights[next], strides[next]);
        ++next;
        if (rc == FE_OK) rc = quad_begin(c, t_max, t_min, params, 0, (size_t)-1);
        if (rc == FE_OK) rc = quad_enqueue(c);
        return rc;
    };
    for (int k = 0; k < 2; ++k) {
        const int rc = start_next(k);
        if (rc != FE_OK) return fail_from(k, rc);
    }
    while (image_of[0] >= 0 || image_of[1] >= 0) {
        for (int k = 0; k < 2; ++k) {
            if (image_of[k] < 0) continue;
            fe_ctx* c = ctx->sub[k];
            int rc = quad_complete(c);
            if (rc != FE_OK) return fail_from(k, rc);
            if (!quad_finished(job_of(c))) {
                rc = quad_enqueue(c);
                if (rc != FE_OK) return fail_from(k, rc);
                continue;
            }
            const size_t n = c->n_items, img = (size_t)image_of[k];
            if (n > caps[img]) return fe_fail(ctx, FE_ERR_CAPACITY, "fe_encode_planes: plane %zu has %zu items, capacity %zu", img, n, caps[img]);
            if (n) FE_CUDA(ctx, cudaMemcpyAsync(out + out_offsets[img], c->b_items.p, n * sizeof(fe_encode_item), cudaMemcpyDeviceToHost, c->stream));
            n_out[img] = n;
            rc = start_next(k);        // the child's stream orders the copy before the next image's kernels touch b_items
            if (rc != FE_OK) return fail_from(k, rc);
        }
    }
    for (int k = 0; k < 2; ++k) {
        FE_CUDA(ctx, cudaStreamSynchronize(ctx->sub[k]->stream));
        ctx->stats.kernel_launches += ctx->sub[k]->stats.kernel_launches;
        ctx->stats.matches += ctx->sub[k]->stats.matches;
        ctx->stats.evaluated += ctx->sub[k]->stats.evaluated;
        ctx->stats.umma_levels += ctx->sub[k]->stats.umma_levels;
        ctx->stats.exact_levels += ctx->sub[k]->stats.exact_levels;
        fe_stats_reset(ctx->sub[k]);
    }
    return FE_OK;
}

extern "C" int fe_encode_batch(fe_ctx* ctx, const uint8_t* const* images, size_t n_images, uint32_t width, uint32_t height, uint32_t stride,
                               uint32_t t_max, uint32_t t_min, const fe_params* params, fe_encode_item* out, size_t cap_per_image, size_t* n_out) {
    if (!ctx) return FE_ERR_INVALID;
    std::vector<uint32_t> w(n_images, width), h(n_images, height), st(n_images, stride);
    std::vector<size_t> off(n_images), caps(n_images, cap_per_image);
    for (size_t i = 0; i < n_images; ++i) off[i] = i * cap_per_image;
    return fe_encode_planes(ctx, images, n_images, w.data(), h.data(), st.data(), t_max, t_min, params, out, off.data(), caps.data(), n_out);
}

// ImageIO::rgb2yuv / yuv2rgb (image/ImageIO.cpp:40-84) on host buffers through the device.
extern "C" int fe_rgb_to_yuv420(fe_ctx* ctx, const uint8_t* rgb, uint32_t width, uint32_t height, uint32_t rgb_stride_bytes, uint8_t* y, uint32_t y_stride,
                                uint8_t* u, uint32_t u_stride, uint8_t* v, uint32_t v_stride, int fma) {
    if (!ctx || !rgb || !y || !u || !v) return fe_fail(ctx, FE_ERR_INVALID, "fe_rgb_to_yuv420: null argument");
    if (!width || !height || (width & 1) || (height & 1)) return fe_fail(ctx, FE_ERR_INVALID, "fe_rgb_to_yuv420: width and height must be even and non-zero");
    if (rgb_stride_bytes < 3 * width || y_stride < width || u_stride < width / 2 || v_stride < width / 2) return fe_fail(ctx, FE_ERR_INVALID, "fe_rgb_to_yuv420: stride too small");
    FE_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t nrgb = (size_t)rgb_stride_bytes * height, ny = (size_t)width * height, nc = ny / 4;
    FE_CUDA(ctx, ctx->b_dec_a.ensure(nrgb + 64));
    FE_CUDA(ctx, ctx->b_dec_b.ensure(ny + 2 * nc + 64));
    uint8_t* d_y = ctx->b_dec_b.as<uint8_t>();
    uint8_t* d_u = d_y + ny;
    uint8_t* d_v = d_u + nc;
    FE_CUDA(ctx, cudaMemcpyAsync(ctx->b_dec_a.p, rgb, nrgb, cudaMemcpyHostToDevice, ctx->stream));
    dim3 block(32, 8), grid(cdiv(width / 2, 32), cdiv(height / 2, 8));
    LAUNCH(ctx, k_rgb2yuv420, grid, block, ctx->b_dec_a.as<uint8_t>(), width, height, rgb_stride_bytes, d_y, width, d_u, width / 2, d_v, width / 2, fma);
    FE_CUDA(ctx, cudaMemcpy2DAsync(y, y_stride, d_y, width, width, height, cudaMemcpyDeviceToHost, ctx->stream));
    FE_CUDA(ctx, cudaMemcpy2DAsync(u, u_stride, d_u, width / 2, width / 2, height / 2, cudaMemcpyDeviceToHost, ctx->stream));
    FE_CUDA(ctx, cudaMemcpy2DAsync(v, v_stride, d_v, width / 2, width / 2, height / 2, cudaMemcpyDeviceToHost, ctx->stream));
    FE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FE_OK;
}

extern "C" int fe_yuv420_to_rgb(fe_ctx* ctx, const uint8_t* y, uint32_t width, uint32_t height, uint32_t y_stride, const uint8_t* u, uint32_t u_stride,
                                const uint8_t* v, uint32_t v_stride, uint8_t* rgb, uint32_t rgb_stride_pixels, int fma) {
    if (!ctx || !rgb || !y || !u || !v) return fe_fail(ctx, FE_ERR_INVALID, "fe_yuv420_to_rgb: null argument");
    if (!width || !height || (width & 1) || (height & 1)) return fe_fail(ctx, FE_ERR_INVALID, "fe_yuv420_to_rgb: width and height must be even and non-zero");
    if (rgb_stride_pixels < width || y_stride < width || u_stride < width / 2 || v_stride < width / 2) return fe_fail(ctx, FE_ERR_INVALID, "fe_yuv420_to_rgb: stride too small");
    FE_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t nrgb = (size_t)rgb_stride_pixels * 3 * height, ny = (size_t)width * height, nc = ny / 4;
    FE_CUDA(ctx, ctx->b_dec_a.ensure(nrgb + 64));
    FE_CUDA(ctx, ctx->b_dec_b.ensure(ny + 2 * nc + 64));
    uint8_t* d_y = ctx->b_dec_b.as<uint8_t>();
    uint8_t* d_u = d_y + ny;
    uint8_t* d_v = d_u + nc;
    FE_CUDA(ctx, cudaMemcpy2DAsync(d_y, width, y, y_stride, width, height, cudaMemcpyHostToDevice, ctx->stream));
    FE_CUDA(ctx, cudaMemcpy2DAsync(d_u, width / 2, u, u_stride, width / 2, height / 2, cudaMemcpyHostToDevice, ctx->stream));
    FE_CUDA(ctx, cudaMemcpy2DAsync(d_v, width / 2, v, v_stride, width / 2, height / 2, cudaMemcpyHostToDevice, ctx->stream));
    dim3 block(32, 8), grid(cdiv(width, 32), cdiv(height, 8));
    LAUNCH(ctx, k_yuv420_to_rgb, grid, block, d_y, width, height, width, d_u, width / 2, d_v, width / 2, ctx->b_dec_a.as<uint8_t>(), rgb_stride_pixels, fma);
    // only the pixels of every row go back: the caller's row padding is left alone
    FE_CUDA(ctx, cudaMemcpy2DAsync(rgb, (size_t)rgb_stride_pixels * 3, ctx->b_dec_a.p, (size_t)rgb_stride_pixels * 3, (size_t)width * 3, height, cudaMemcpyDeviceToHost, ctx->stream));
    FE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FE_OK;
}

extern "C" int fe_fetch_items(fe_ctx* ctx, fe_encode_item* out, size_t cap, size_t* n_out) {
    if (!ctx || (!out && ctx->n_items)) return fe_fail(ctx, FE_ERR_INVALID, "fe_fetch_items: null argument");
    if (cap < ctx->n_items) return fe_fail(ctx, FE_ERR_CAPACITY, "fe_fetch_items: %zu items, capacity %zu", ctx->n_items, cap);
    if (ctx->n_items) FE_CUDA(ctx, cudaMemcpyAsync(out, ctx->b_items.p, ctx->n_items * sizeof(fe_encode_item), cudaMemcpyDeviceToHost, ctx->stream));
    FE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (n_out) *n_out = ctx->n_items;
    return FE_OK;
}

extern "C" const void* fe_device_items(const fe_ctx* ctx, size_t* n_out) {
    if (!ctx) return nullptr;
    if (n_out) *n_out = ctx->n_items;
    return ctx->b_items.p;
}

extern "C" int fe_encode_quadtree(fe_ctx* ctx, uint32_t t_max, uint32_t t_min, const fe_params* params, fe_encode_item* out,
                                  size_t cap, size_t* n_out, size_t* level_counts) {
    size_t n = 0;
    if (!ctx) return FE_ERR_INVALID;
    ctx->host_out = out; ctx->host_cap = out ? cap : 0; ctx->host_copied = 0;
    const int rc = fe_encode_quadtree_device(ctx, t_max, t_min, params, &n);
    ctx->host_out = nullptr;
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    if (rc != FE_OK) return rc;
    if (level_counts) {
        int level = 0;
        for (uint32_t T = t_max; T >= t_min; T /= 2, ++level) level_counts[level] = (size_t)ctx->stats.level_items[level];
    }
    if (ctx->host_copied == n && n <= cap) { // everything already went out level by level
        FE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (n_out) *n_out = n;
        return FE_OK;
    }
    FE_TRY(fe_fetch_items(ctx, out, cap, n_out));
    return FE_OK;
}

// -------------------------------------------------------------------------------------------------
// decode
// -------------------------------------------------------------------------------------------------
extern "C" int fe_decode(fe_ctx* ctx, const fe_encode_item* items, size_t n, uint8_t* target, uint32_t width, uint32_t height,
                         uint32_t stride, int max_iters, double rms_eps, int use_fma, int* iterations, double* rms_out) {
    if (!ctx) return FE_ERR_INVALID;
    if ((!items && n) || !target || !width || !height || stride < width) return fe_fail(ctx, FE_ERR_INVALID, "fe_decode: bad arguments");
    if (n > 0x7FFFFFFFu || (uint64_t)width * height > 0xFFFFFFFFull) return fe_fail(ctx, FE_ERR_UNSUPPORTED, "fe_decode: too large");
    // validate; group the items by block size: every group of square blocks with T % 4 == 0 runs on the 4-pixels-per-thread
    // kernel, the rest (odd sizes, non-square) on the per-pixel kernel with prefix offsets
    bool has_default = false;
    uint64_t area = 0;
    std::vector<uint32_t> sizes;   // distinct fast sizes
    for (size_t i = 0; i < n; ++i) {
        const fe_encode_item& e = items[i];
        if (!e.w || !e.h || !inside(e.x, e.y, e.w, e.h, width, height)) return fe_fail(ctx, FE_ERR_INVALID, "fe_decode: item %zu outside the image", i);
        if (e.src_w && (e.src_w != e.src_h || !inside(e.match_x, e.match_y, e.src_w, e.src_h, width, height) || e.transform < 0 || e.transform > 7 || e.src_w < 2))
            return fe_fail(ctx, FE_ERR_INVALID, "fe_decode: item %zu has an invalid source block", i);
        area += (uint64_t)e.w * e.h;
        if (!e.src_w || !e.src_h) has_default = true;
        if (e.w == e.h && e.w % 4 == 0 && std::find(sizes.begin(), sizes.end(), e.w) == sizes.end() && sizes.size() < 8) sizes.push_back(e.w);
    }
    if (area > 0xFFFFFFFFull) return fe_fail(ctx, FE_ERR_UNSUPPORTED, "fe_decode: item areas overflow");
    std::vector<fe_encode_item> grouped;
    grouped.reserve(n);
    std::vector<size_t> goff{0};
    std::vector<char> tiled;       // per size: every item has a 2T source block at a 4-byte aligned origin inside aligned rows
    bool even_rows = true;         // every item of the fast sizes starts at an even row (its 2 x 2 boxes are its own)
    for (uint32_t T : sizes) {
        bool ok = (stride % 4 == 0) && T <= 32;
        for (size_t i = 0; i < n; ++i)
            if (items[i].w == T && items[i].h == T) {
                const fe_encode_item& e = items[i];
                grouped.push_back(e);
                ok = ok && e.src_w == 2 * T && e.src_h == 2 * T && e.match_x % 4 == 0 && e.match_y % 2 == 0 && e.x % 4 == 0;
                even_rows = even_rows && e.y % 2 == 0;
            }
        goff.push_back(grouped.size());
        tiled.push_back(ok ? 1 : 0);
    }
    const size_t n_fast = grouped.size();
    std::vector<uint32_t> pix_off;
    uint64_t slow_area = 0;
    for (size_t i = 0; i < n; ++i) {
        const fe_encode_item& e = items[i];
        if (e.w == e.h && std::find(sizes.begin(), sizes.end(), e.w) != sizes.end()) continue;
        grouped.push_back(e);
        pix_off.push_back((uint32_t)slow_area);
        slow_area += (uint64_t)e.w * e.h;
    }
    pix_off.push_back((uint32_t)slow_area);
    const size_t n_slow = n - n_fast;
    const int iters = max_iters < 0 ? 300 : max_iters;
    const size_t bytes = (size_t)height * stride;
    FE_CUDA(ctx, cudaSetDevice(ctx->device));
    FE_CUDA(ctx, ctx->b_dec_a.ensure(bytes + 64));
    FE_CUDA(ctx, ctx->b_dec_b.ensure(bytes + 64));
    FE_CUDA(ctx, ctx->b_dec_items.ensure(n * sizeof(fe_encode_item) + (n_slow + 1) * 4 + 64));
    FE_CUDA(ctx, ctx->b_dec_sum.ensure(64 * 8 + 64));
    fe_encode_item* d_items = ctx->b_dec_items.as<fe_encode_item>();
    uint32_t* d_off = reinterpret_cast<uint32_t*>(d_items + n);
    if (n) {
        FE_CUDA(ctx, cudaMemcpyAsync(d_items, grouped.data(), n * sizeof(fe_encode_item), cudaMemcpyHostToDevice, ctx->stream));
        FE_CUDA(ctx, cudaMemcpyAsync(d_off, pix_off.data(), (n_slow + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    unsigned long long* d_sum = ctx->b_dec_sum.as<unsigned long long>();
    uint32_t* d_state = reinterpret_cast<uint32_t*>(d_sum + 64);       // {done, iterations, rms bits lo, hi} ... [8..] coverage scratch
    FE_CUDA(ctx, cudaMemsetAsync(d_sum, 0, 64 * 8 + 64, ctx->stream));

    // ---- do the items tile the plane?  Equal area does not prove it: every pixel must be written exactly once (bitmap). ----
    bool covered = false;
    if (n) {
        const bool maybe = !has_default && area == (uint64_t)width * height;
        const uint32_t wpr = (width + 31) / 32;
        std::vector<uint32_t> row_off(n + 1, 0);
        for (size_t i = 0; i < n; ++i) row_off[i + 1] = row_off[i] + grouped[i].h;
        FE_CUDA(ctx, ctx->b_q.ensure((size_t)wpr * height * 4 + (n + 1) * 4 + 64));
        uint32_t* d_bm = ctx->b_q.as<uint32_t>();
        uint32_t* d_rows = d_bm + (size_t)wpr * height;
        FE_CUDA(ctx, cudaMemsetAsync(d_bm, 0, (size_t)wpr * height * 4, ctx->stream));
        FE_CUDA(ctx, cudaMemcpyAsync(d_rows, row_off.data(), (n + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
        LAUNCH(ctx, k_cover_bitmap, cdiv(row_off[n], 256), 256, d_items, (uint32_t)n, d_rows, row_off[n], wpr, d_bm, d_state + 8);
        if (maybe) LAUNCH(ctx, k_popcount, ctx->n_sm * 4, 256, d_bm, (size_t)wpr * height, reinterpret_cast<unsigned long long*>(d_state + 10));
        uint32_t cov[4] = {0, 0, 0, 0};                                    // overlap flag, -, covered pixels (u64)
        FE_CUDA(ctx, cudaMemcpyAsync(cov, d_state + 8, sizeof(cov), cudaMemcpyDeviceToHost, ctx->stream));
        FE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        // gaps are fine (the copy path below); overlaps are not: the reference applies its items in list order
        if (cov[0])
            return fe_fail(ctx, FE_ERR_INVALID, "fe_decode: items overlap (Decoder2 applies its items in list order, the last one wins; "
                                                "overlapping lists are not supported)");
        covered = maybe && (((uint64_t)cov[3] << 32) | cov[2]) == (uint64_t)width * height;
    }
    const bool fused_sq = covered && n_slow == 0;                          // convergence sum inside the gather kernel
    uint8_t* buf[2] = {ctx->b_dec_a.as<uint8_t>(), ctx->b_dec_b.as<uint8_t>()};
    FE_CUDA(ctx, cudaMemsetAsync(buf[0], 100, bytes, ctx->stream)); // Encoder2.hpp:69
    // The encoder's own lists (4- and 8-pixel blocks on the lattice, tiling the plane): the iterate carries its half-resolution
    // plane of 2 x 2 box sums, the gather reads those (k_decode_step_small<.., DQ>)
    bool with_dq = fused_sq && even_rows && width % 4 == 0 && height % 2 == 0 && !getenv("FE_NO_DQ");
    for (size_t gidx = 0; gidx < sizes.size(); ++gidx) with_dq = with_dq && tiled[gidx] && (sizes[gidx] == 4 || sizes[gidx] == 8);
    uint16_t* dq[2] = {nullptr, nullptr};
    const uint32_t dq_stride = width / 2;
    if (with_dq) {
        const size_t qn = (size_t)(width / 2) * (height / 2);
        FE_CUDA(ctx, ctx->b_dq[0].ensure(qn * 2 + 64));
        FE_CUDA(ctx, ctx->b_dq[1].ensure(qn * 2 + 64));
        dq[0] = ctx->b_dq[0].as<uint16_t>(); dq[1] = ctx->b_dq[1].as<uint16_t>();
        launch_boxsum_plane(ctx->stream, buf[0], stride, width, height, dq[0]);
        ctx->stats.kernel_launches++;
        FE_CUDA(ctx, cudaGetLastError());
    }
    FE_CUDA(ctx, cudaMemcpy2DAsync(buf[1], stride, target, stride, width, height, cudaMemcpyHostToDevice, ctx->stream));
    cudaEventRecord(ctx->ev[0], ctx->stream);
    // ---- the iteration train: BATCH iterations per host look; once `done` is up the rest of a batch are no-ops ----
    const int BATCH = 8;
    uint32_t state[4] = {0, 0, 0, 0};
    const uint32_t npix = (uint32_t)(width * height);
    for (int i0 = 0; i0 < iters && !state[0]; i0 += BATCH) {
        for (int i = i0; i < std::min(iters, i0 + BATCH); ++i) {
            // tiled plane: the two buffers ping-pong (every pixel is rewritten); else buffer 0 is the source, refreshed by a copy
            const uint8_t* src = covered ? buf[i & 1] : buf[0];
            uint8_t* dst = covered ? buf[(i + 1) & 1] : buf[1];
            for (size_t gidx = 0; gidx < sizes.size(); ++gidx) {
                const uint32_t T = sizes[gidx];
                const size_t cnt = goff[gidx + 1] - goff[gidx];
                if (!cnt) continue;
                if (tiled[gidx]) {
                    if (!launch_decode_step_small(ctx->stream, src, dst, stride, d_items + goff[gidx], (uint32_t)cnt, T, use_fma, fused_sq ? d_sum : nullptr, d_state,
                                                  with_dq ? dq[i & 1] : nullptr, with_dq ? dq[(i + 1) & 1] : nullptr, dq_stride))
                        k_decode_step_tiled<<<cdiv(cnt, 8), 256, 8 * T * T * sizeof(uint16_t), ctx->stream>>>(src, dst, stride, d_items + goff[gidx], (uint32_t)cnt, T,
                                                                                                            use_fma, fused_sq ? d_sum : nullptr, d_state);
                    ctx->stats.kernel_launches++;
                    FE_CUDA(ctx, cudaGetLastError());
                } else {
                    LAUNCH(ctx, k_decode_step_uniform, cdiv((uint64_t)cnt * T * (T / 4), 256), 256, src, dst, stride, d_items + goff[gidx], (uint32_t)cnt, T,
                           use_fma, fused_sq ? d_sum : nullptr, d_state);
                }
            }
            if (n_slow)
                LAUNCH(ctx, k_decode_step, cdiv(slow_area, 256), 256, src, dst, stride, d_items + n_fast, d_off, (uint32_t)n_slow, (uint32_t)slow_area, use_fma,
                       d_state);
            if (!fused_sq) LAUNCH(ctx, k_sqdiff, ctx->n_sm * 8, 256, src, dst, width, height, stride, d_sum, d_state);
            LAUNCH(ctx, k_decode_check, 1, 32, d_sum, d_state, npix, rms_eps, (uint32_t)i);
            if (!covered)   // source = target.copy() (Encoder2.hpp:86), unless this iteration converged
                LAUNCH(ctx, k_copy_plane_if_running, ctx->n_sm * 4, 256, reinterpret_cast<const uint4*>(buf[1]), reinterpret_cast<uint4*>(buf[0]), (bytes + 15) / 16,
                       d_state);
        }
        FE_CUDA(ctx, cudaMemcpyAsync(state, d_state, sizeof(state), cudaMemcpyDeviceToHost, ctx->stream));
        FE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    const int done_iter = (int)state[1];                                     // index of the converged iteration, or `iters`
    const unsigned long long rbits = ((unsigned long long)state[3] << 32) | state[2];
    double rms = 0.0;
    memcpy(&rms, &rbits, 8);
    // freshest plane: the destination of the last iteration that ran
    const int last = state[0] ? done_iter : iters - 1;
    const uint8_t* result = (covered && last >= 0) ? buf[(last + 1) & 1] : buf[1];
    cudaEventRecord(ctx->ev[1], ctx->stream);
    FE_CUDA(ctx, cudaMemcpy2DAsync(target, stride, result, stride, width, height, cudaMemcpyDeviceToHost, ctx->stream));   // never the caller's padding
    FE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->stats.last_decode_ms, ctx->ev[0], ctx->ev[1]);
    if (iterations) *iterations = done_iter;
    if (rms_out) *rms_out = rms;
    return FE_OK;
}

extern "C" int fe_copy_items(fe_ctx* ctx, const uint8_t* source, uint8_t* target, uint32_t width, uint32_t height, uint32_t stride,
                             const fe_encode_item* items, size_t n, int use_fma) {
    if (!ctx) return FE_ERR_INVALID;
    if (!source || !target || (!items && n) || !width || !height || stride < width) return fe_fail(ctx, FE_ERR_INVALID, "fe_copy_items: bad arguments");
    if (n > 0x7FFFFFFFu) return fe_fail(ctx, FE_ERR_UNSUPPORTED, "fe_copy_items: too many items");
    std::vector<uint32_t> pix_off(n + 1, 0);
    uint64_t area = 0;
    for (size_t i = 0; i < n; ++i) {
        const fe_encode_item& e = items[i];
        if (!e.w || !e.h || !inside(e.x, e.y, e.w, e.h, width, height)) return fe_fail(ctx, FE_ERR_INVALID, "fe_copy_items: item %zu outside the image", i);
        if (e.src_w && (e.src_w != e.src_h || !inside(e.match_x, e.match_y, e.src_w, e.src_h, width, height) || e.transform < 0 || e.transform > 7 || e.src_w < 2))
            return fe_fail(ctx, FE_ERR_INVALID, "fe_copy_items: item %zu has an invalid source block", i);
        pix_off[i] = (uint32_t)area;
        area += (uint64_t)e.w * e.h;
    }
    if (area > 0xFFFFFFFFull) return fe_fail(ctx, FE_ERR_UNSUPPORTED, "fe_copy_items: item areas overflow");
    pix_off[n] = (uint32_t)area;
    const size_t bytes = (size_t)height * stride;
    FE_CUDA(ctx, cudaSetDevice(ctx->device));
    FE_CUDA(ctx, ctx->b_dec_a.ensure(bytes + 64));
    FE_CUDA(ctx, ctx->b_dec_b.ensure(bytes + 64));
    FE_CUDA(ctx, ctx->b_dec_items.ensure(n * sizeof(fe_encode_item) + (n + 1) * 4 + 64));
    fe_encode_item* d_items = ctx->b_dec_items.as<fe_encode_item>();
    uint32_t* d_off = reinterpret_cast<uint32_t*>(d_items + n);
    FE_CUDA(ctx, cudaMemcpyAsync(ctx->b_dec_a.p, source, bytes, cudaMemcpyHostToDevice, ctx->stream));
    FE_CUDA(ctx, cudaMemcpyAsync(ctx->b_dec_b.p, target, bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (n) {
        FE_CUDA(ctx, cudaMemcpyAsync(d_items, items, n * sizeof(fe_encode_item), cudaMemcpyHostToDevice, ctx->stream));
        FE_CUDA(ctx, cudaMemcpyAsync(d_off, pix_off.data(), (n + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
        LAUNCH(ctx, k_decode_step, cdiv(area, 256), 256, ctx->b_dec_a.as<uint8_t>(), ctx->b_dec_b.as<uint8_t>(), stride, d_items, d_off, (uint32_t)n,
               (uint32_t)area, use_fma, nullptr);
    }
    FE_CUDA(ctx, cudaMemcpyAsync(target, ctx->b_dec_b.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    FE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FE_OK;
}

// -------------------------------------------------------------------------------------------------
// quantizer post-pass
// -------------------------------------------------------------------------------------------------
static double key_to_double(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    double d;
    memcpy(&d, &b, 8);
    return d;
}

extern "C" int fe_quantize(fe_ctx* ctx, const fe_encode_item* items, size_t n, int bits_s, int bits_o, uint32_t* qs, uint32_t* qo,
                           double minmax_out[4]) {
    if (!ctx) return FE_ERR_INVALID;
    if (!items || !n || !qs || !qo) return fe_fail(ctx, FE_ERR_INVALID, "fe_quantize: null or empty input");
    if (bits_s < 2 || bits_s > 30 || bits_o < 2 || bits_o > 30) return fe_fail(ctx, FE_ERR_INVALID, "fe_quantize: bits out of range");
    FE_CUDA(ctx, cudaSetDevice(ctx->device));
    FE_CUDA(ctx, ctx->b_dec_items.ensure(n * sizeof(fe_encode_item)));
    FE_CUDA(ctx, ctx->b_q.ensure(64 + n * 8));
    FE_CUDA(ctx, cudaMemcpyAsync(ctx->b_dec_items.p, items, n * sizeof(fe_encode_item), cudaMemcpyHostToDevice, ctx->stream));
    unsigned long long mm[4] = {FE_INF64, 0, FE_INF64, 0};
    unsigned long long* d_mm = reinterpret_cast<unsigned long long*>(ctx->b_q.as<uint8_t>());
    FE_CUDA(ctx, cudaMemcpyAsync(d_mm, mm, sizeof(mm), cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(ctx, k_minmax, 148 * 4, 256, ctx->b_dec_items.as<fe_encode_item>(), (uint32_t)n, d_mm);
    FE_CUDA(ctx, cudaMemcpyAsync(mm, d_mm, sizeof(mm), cudaMemcpyDeviceToHost, ctx->stream));
    FE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // main.cpp:109-118: max starts at -1, min at DBL_MAX
    const double min_s = std::fmin(1.7976931348623157e308, key_to_double(mm[0])), max_s = std::fmax(-1.0, key_to_double(mm[1]));
    const double min_o = std::fmin(1.7976931348623157e308, key_to_double(mm[2])), max_o = std::fmax(-1.0, key_to_double(mm[3]));
    if (minmax_out) { minmax_out[0] = min_s; minmax_out[1] = max_s; minmax_out[2] = min_o; minmax_out[3] = max_o; }
    if (!(max_s > min_s) || !(max_o > min_o)) return fe_fail(ctx, FE_ERR_INVALID, "fe_quantize: degenerate value range (Quantizer asserts max > min)");
    uint32_t* d_qs = reinterpret_cast<uint32_t*>(ctx->b_q.as<uint8_t>() + 64);
    uint32_t* d_qo = d_qs + n;
    LAUNCH(ctx, k_quantize, cdiv(n, 256), 256, ctx->b_dec_items.as<fe_encode_item>(), (uint32_t)n, min_s, max_s, min_o, max_o, bits_s, bits_o, d_qs, d_qo);
    FE_CUDA(ctx, cudaMemcpyAsync(qs, d_qs, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    FE_CUDA(ctx, cudaMemcpyAsync(qo, d_qo, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    FE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FE_OK;
}

// -------------------------------------------------------------------------------------------------
// packed quantised records
// -------------------------------------------------------------------------------------------------
static int list_minmax(fe_ctx* ctx, const fe_encode_item* d_items, size_t n, double mm_out[4]) {
    unsigned long long mm[4] = {FE_INF64, 0, FE_INF64, 0};
    unsigned long long* d_mm = reinterpret_cast<unsigned long long*>(ctx->b_q.as<uint8_t>());
    FE_CUDA(ctx, cudaMemcpyAsync(d_mm, mm, sizeof(mm), cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(ctx, k_minmax, 148 * 4, 256, d_items, (uint32_t)n, d_mm);
    FE_CUDA(ctx, cudaMemcpyAsync(mm, d_mm, sizeof(mm), cudaMemcpyDeviceToHost, ctx->stream));
    FE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    mm_out[0] = std::fmin(1.7976931348623157e308, key_to_double(mm[0]));
    mm_out[1] = std::fmax(-1.0, key_to_double(mm[1]));
    mm_out[2] = std::fmin(1.7976931348623157e308, key_to_double(mm[2]));
    mm_out[3] = std::fmax(-1.0, key_to_double(mm[3]));
    return FE_OK;
}

extern "C" int fe_pack_items(fe_ctx* ctx, const fe_encode_item* items, size_t n, uint32_t t_max, int bits_s, int bits_o,
                             uint64_t* packed_out, double minmax_out[4]) {
    if (!ctx) return FE_ERR_INVALID;
    if (!items || !n || !packed_out || !minmax_out) return fe_fail(ctx, FE_ERR_INVALID, "fe_pack_items: null or empty input");
    if (bits_s < 2 || bits_s > 5 || bits_o < 2 || bits_o > 7) return fe_fail(ctx, FE_ERR_INVALID, "fe_pack_items: bits_s in 2..5, bits_o in 2..7");
    if (!t_max || (t_max & (t_max - 1)) || n > 0x7FFFFFFFu) return fe_fail(ctx, FE_ERR_INVALID, "fe_pack_items: t_max must be a power of two");
    FE_CUDA(ctx, cudaSetDevice(ctx->device));
    FE_CUDA(ctx, ctx->b_dec_items.ensure(n * sizeof(fe_encode_item)));
    FE_CUDA(ctx, ctx->b_q.ensure(128 + n * 8));
    FE_CUDA(ctx, cudaMemcpyAsync(ctx->b_dec_items.p, items, n * sizeof(fe_encode_item), cudaMemcpyHostToDevice, ctx->stream));
    FE_TRY(list_minmax(ctx, ctx->b_dec_items.as<fe_encode_item>(), n, minmax_out));
    if (!(minmax_out[1] > minmax_out[0]) || !(minmax_out[3] > minmax_out[2]))
        return fe_fail(ctx, FE_ERR_INVALID, "fe_pack_items: degenerate value range (Quantizer asserts max > min)");
    uint32_t* d_bad = reinterpret_cast<uint32_t*>(ctx->b_q.as<uint8_t>() + 64);
    double* d_mm = reinterpret_cast<double*>(ctx->b_q.as<uint8_t>() + 96);
    unsigned long long* d_out = reinterpret_cast<unsigned long long*>(ctx->b_q.as<uint8_t>() + 128);
    FE_CUDA(ctx, cudaMemsetAsync(d_bad, 0, 4, ctx->stream));
    FE_CUDA(ctx, cudaMemcpyAsync(d_mm, minmax_out, 4 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(ctx, k_pack, cdiv(n, 256), 256, ctx->b_dec_items.as<fe_encode_item>(), (uint32_t)n, t_max, d_mm, bits_s, bits_o, d_out, d_bad);
    uint32_t bad = 0;
    FE_CUDA(ctx, cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, ctx->stream));
    FE_CUDA(ctx, cudaMemcpyAsync(packed_out, d_out, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    FE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (bad) return fe_fail(ctx, FE_ERR_UNSUPPORTED, "fe_pack_items: %u items are not power-of-two lattice blocks with S = 2T below t_max=%u", bad, t_max);
    return FE_OK;
}

extern "C" int fe_items_minmax_device(fe_ctx* ctx, double* minmax_dev) {
    if (!ctx || !minmax_dev) return fe_fail(ctx, FE_ERR_INVALID, "fe_items_minmax_device: null argument");
    FE_CUDA(ctx, cudaSetDevice(ctx->device));
    FE_CUDA(ctx, ctx->b_q.ensure(256));
    unsigned long long mm[4] = {FE_INF64, 0, FE_INF64, 0};
    unsigned long long* d_keys = reinterpret_cast<unsigned long long*>(ctx->b_q.as<uint8_t>());
    FE_CUDA(ctx, cudaMemcpyAsync(d_keys, mm, sizeof(mm), cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->n_items) LAUNCH(ctx, k_minmax, ctx->n_sm * 4, 256, ctx->b_items.as<fe_encode_item>(), (uint32_t)ctx->n_items, d_keys);
    LAUNCH(ctx, k_minmax_finish, 1, 32, d_keys, minmax_dev);
    return FE_OK;
}

extern "C" int fe_pack_items_device(fe_ctx* ctx, uint32_t t_max, int bits_s, int bits_o, const double* minmax_dev, uint64_t* packed_dev,
                                    size_t cap, size_t* n_out) {
    if (!ctx || !minmax_dev || (!packed_dev && ctx->n_items)) return fe_fail(ctx, FE_ERR_INVALID, "fe_pack_items_device: null argument");
    if (bits_s < 2 || bits_s > 5 || bits_o < 2 || bits_o > 7) return fe_fail(ctx, FE_ERR_INVALID, "fe_pack_items_device: bits_s in 2..5, bits_o in 2..7");
    if (!t_max || (t_max & (t_max - 1))) return fe_fail(ctx, FE_ERR_INVALID, "fe_pack_items_device: t_max must be a power of two");
    if (cap < ctx->n_items) return fe_fail(ctx, FE_ERR_CAPACITY, "fe_pack_items_device: %zu items, capacity %zu", ctx->n_items, cap);
    FE_CUDA(ctx, cudaSetDevice(ctx->device));
    FE_CUDA(ctx, ctx->b_counters.ensure(16 * sizeof(uint32_t)));
    uint32_t* d_bad = ctx->b_counters.as<uint32_t>() + 12;     // read by fe_pack_errors
    FE_CUDA(ctx, cudaMemsetAsync(d_bad, 0, 4, ctx->stream));
    if (ctx->n_items)
        LAUNCH(ctx, k_pack, cdiv(ctx->n_items, 256), 256, ctx->b_items.as<fe_encode_item>(), (uint32_t)ctx->n_items, t_max, minmax_dev, bits_s, bits_o,
               reinterpret_cast<unsigned long long*>(packed_dev), d_bad);
    if (n_out) *n_out = ctx->n_items;
    return FE_OK;
}

extern "C" int fe_pack_errors(fe_ctx* ctx, uint32_t* n_bad) {
    if (!ctx || !n_bad) return FE_ERR_INVALID;
    *n_bad = 0;
    if (!ctx->b_counters.p) return FE_OK;
    FE_CUDA(ctx, cudaMemcpyAsync(n_bad, ctx->b_counters.as<uint32_t>() + 12, 4, cudaMemcpyDeviceToHost, ctx->stream));
    FE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FE_OK;
}

extern "C" int fe_unpack_items(fe_ctx* ctx, const uint64_t* packed, size_t n, uint32_t t_max, int bits_s, int bits_o, const double minmax[4],
                               int use_fma, fe_encode_item* items_out) {
    if (!ctx) return FE_ERR_INVALID;
    if (!packed || !n || !items_out || !minmax) return fe_fail(ctx, FE_ERR_INVALID, "fe_unpack_items: null or empty input");
    if (bits_s < 2 || bits_s > 5 || bits_o < 2 || bits_o > 7 || !t_max || (t_max & (t_max - 1)) || n > 0x7FFFFFFFu)
        return fe_fail(ctx, FE_ERR_INVALID, "fe_unpack_items: bad header");
    FE_CUDA(ctx, cudaSetDevice(ctx->device));
    FE_CUDA(ctx, ctx->b_dec_items.ensure(n * sizeof(fe_encode_item)));
    FE_CUDA(ctx, ctx->b_q.ensure(128 + n * 8));
    unsigned long long* d_in = reinterpret_cast<unsigned long long*>(ctx->b_q.as<uint8_t>() + 128);
    FE_CUDA(ctx, cudaMemcpyAsync(d_in, packed, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(ctx, k_unpack, cdiv(n, 256), 256, d_in, (uint32_t)n, t_max, minmax[0], minmax[1], minmax[2], minmax[3], bits_s, bits_o, use_fma,
           ctx->b_dec_items.as<fe_encode_item>());
    FE_CUDA(ctx, cudaMemcpyAsync(items_out, ctx->b_dec_items.p, n * sizeof(fe_encode_item), cudaMemcpyDeviceToHost, ctx->stream));
    FE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FE_OK;
}
