// fe_internal.cuh -- shared declarations of the B200 fractal-search library (not part of the C ABI).
#pragma once
#include <chrono>

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "fractencode_b200.h"

#define FE_NONE32 0xFFFFFFFFu
#define FE_INF64 0xFFFFFFFFFFFFFFFFull

// Growable device buffer owned by a ctx (no allocation inside steady-state loops).
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct Plane {
    uint8_t* px = nullptr; // device
    uint32_t w = 0, h = 0, stride = 0;
};

// Geometry of one search level (all blocks square, uniform sizes).
struct LevelGeom {
    uint32_t S = 0, T = 0, rho = 0; // domain size, range size, S/T
    uint32_t N = 0;                 // T*T
    uint32_t Npad = 0;              // N rounded up to a multiple of 16 (bytes per operand row)
    bool fast = false;              // rho == 2 and every domain origin even: rotations move to the range side
};

// Device-side description of one search launch (one classifier bucket).
struct SearchArgs {
    const uint8_t* A;       // [rows][Npad] u8 : range rows, row = 4*rangePos + k
    const uint8_t* Blo;     // [npool][cols][Npad] low byte of the domain box sums D (0..1020)
    const uint8_t* Bhi;     // high byte (0..3)
    const uint32_t* rowc;   // [rangePos] 16*sum(r^2)
    const uint32_t* coln;   // [npool][cols] sum(D^2)
    unsigned long long* rowbest; // [rows] (n16 << 32 | col), FE_INF64 when none
    uint32_t* rowhit;       // [rows] first col with n16 <= thr16, FE_NONE32 when none
    uint32_t row0, nrows;   // row range of this bucket (multiples of 4)
    uint32_t col0, ncols;   // column range of this bucket
    uint32_t Npad;
    uint32_t pool_stride_cols; // columns per rotation pool (generic geometry: 4 pools), 0 when fast
    uint32_t thr16;
    uint32_t use_thr;
    // fp32-regime re-rank pass (reference distance = fp32 running sum, image/metrics.h:38-49): candidates with
    // n16 <= rowbound[row >> 2] are scored by the emulated float sum instead (key = float bits << 32 | column)
    const uint32_t* rowbound;
    const uint8_t* src; uint32_t src_stride;
    const uint8_t* tgt; uint32_t tgt_stride;
    const fe_grid_item* dom; const uint32_t* dom_order; // column -> domain item
    const fe_grid_item* rng; const uint32_t* row_range; // row >> 2 -> range item index
    uint32_t rho;
};

constexpr int FE_MAX_LAUNCHES = 40;  // search launches of one level: the slice train, the minimum pass, and once more after a re-run
constexpr int FE_MAX_BUCKETS = 64;   // buckets of one search launch: classifier classes (7) or brightness bins (<= 64)
constexpr int FE_MAX_GROUPS = 8;     // groups of buckets searched by separate launches of a slice: classifier classes when bins are on
constexpr int FE_MAX_TOTAL = FE_MAX_GROUPS * FE_MAX_BUCKETS;

struct fe_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    Plane src, tgt; // tgt.px == src.px when one image serves both roles
    DevBuf b_src, b_tgt;
    // level scratch
    DevBuf b_dom, b_rng, b_dom_cls, b_rng_cls, b_dom_order, b_rng_order, b_sort_tmp, b_keys_tmp, b_vals_tmp;
    DevBuf b_A, b_Blo, b_Bhi, b_rowc, b_coln, b_rowbest, b_rowhit, b_hist, b_level_items, b_split, b_scan, b_scan_tmp;
    DevBuf b_rng_next, b_counters, b_bound, b_flag_idx;
    // classifier classes x brightness bins: the composite orders (the class-only orders stay in b_dom_order / b_rng_order)
    DevBuf b_dom_order2, b_rng_order2;
    // flip isometries: the doubled range list and the position of every copy after bucketing
    DevBuf b_rng2, b_pos_of;
    // lower-bound prefilter: cell-sum plane, candidate list (+ its counter)
    DevBuf b_lbq, b_lbcand, b_cells, b_dq[2];
    DevBuf b_dom_lvl[8];                  // quadtree: the domain grid of level l (a function of the image size only)
    unsigned long long dom_lvl_tag[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    std::chrono::steady_clock::time_point level_host_t0;   // FE_PASS_TIMES: host clock at the level's first launch
    // tcgen05 path operands
    DevBuf b_A16, b_B16, b_tmaps, b_blob_dom, b_tileseg;
    // device-scheduled levels (fe_plan.cuh): plan, slice state, the two lists of open range blocks, work items, bucket of every
    // level position, and the per-level summary the host reads back (device record + pinned host copy)
    DevBuf b_plan, b_ctl, b_list[2], b_itemrec, b_posb, b_summary;
    void* h_summary = nullptr;
    void* job = nullptr;            // QuadJob (fe_api.cu): the quadtree encode in progress
    fe_ctx* sub[2] = {nullptr, nullptr};   // fe_encode_batch: two child contexts (own stream and scratch) the images alternate between
    int n_sm = 148;                 // multiProcessorCount of the device
    // slices the last level of this kind (f16 / i8) and block size needed: how many the next one gets enqueued up front
    struct SliceHint { uint8_t known = 0, slices = 0, with_min = 0; } hint[3][8];
    // results
    DevBuf b_items;
    size_t n_items = 0;
    // decode scratch
    DevBuf b_dec_a, b_dec_b, b_dec_items, b_dec_sum, b_q;
    fe_stats stats{};
    // fe_encode_quadtree with a host buffer: the items of a level go out on a second stream while the next level runs
    fe_encode_item* host_out = nullptr;
    size_t host_cap = 0, host_copied = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copy = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_pass[2 * FE_MAX_LAUNCHES] = {}; // start/stop around each search launch of a level
};

int fe_fail(fe_ctx* ctx, int code, const char* fmt, ...);
#define FE_CUDA(ctx, call)                                                                       \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess)                                                                  \
            return fe_fail(ctx, FE_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)
#define FE_TRY(expr)               \
    do {                           \
        int rc__ = (expr);         \
        if (rc__ != FE_OK) return rc__; \
    } while (0)

// ---- kernels' host launchers (each returns a cudaError_t from cudaGetLastError) ----
cudaError_t launch_search_exact(fe_ctx* ctx, const SearchArgs& a, bool rerank = false);

