// fe_plan.cuh -- device-resident schedule of one search level (fe_plan.cu plans it, fe_search_f16.cu / fe_search_i8.cu run it).
//
// The reference scans the domains of a range block in order and stops at the first one under the threshold
// (encode/TransformEstimator2.hpp:29-47).  Here the scan of a level is cut into FE_NK base intervals of the domain order
// (cumulative fractions 2^k / 128); a SLICE searches one or more consecutive intervals for the range blocks that are still
// open, after which the closed ones are dropped and the rest compacted.  Everything a slice needs -- the survivors, how
// far to scan next, the work items -- is produced by kernels from device memory, so a level runs without the host: the
// host enqueues a fixed train of (plan, expand, search) launches and the launches past the end of the level return at once.
//
// Layout of the level's operand blob: domains are keyed by bucket (classifier class x brightness bin), the blob holds
// them interval-major: chunk (k, b) = the domains of bucket b whose position lies in interval k.  The buckets a range
// bucket meets (bins c-span .. c+span of its class) are adjacent inside an interval, so a work item is one row tile of
// 32 range blocks x one contiguous run of column tiles.
#pragma once
#include "fe_umma.cuh"

constexpr int FE_NK = 8;                  // base intervals of the scan

struct LevelPlan {
    uint32_t nb;                          // buckets of the level = ngroups * nbins
    uint32_t nbins, ngroups;              // buckets per group (classifier class); a neighbourhood never leaves its group
    uint32_t span;                        // range bucket c meets the domain buckets c-span .. c+span of its group
    uint32_t nD, nR;
    uint32_t nt;                          // columns per blob tile (128: f16 kind, 64: i8 kind)
    uint32_t n_tiles;                     // blob tiles
    uint32_t nk;                          // leading intervals that hold anything (1 when the scan is not sliced)
    uint32_t cut[FE_NK];                  // domain-index cutoff of interval k: every admissible domain below it lies in 0..k
    uint32_t doff[FE_MAX_TOTAL + 1];      // domain positions (sorted order) of bucket b
    uint32_t roff[FE_MAX_TOTAL + 1];      // level positions of the range blocks of bucket b
    uint32_t dend[FE_NK][FE_MAX_TOTAL];   // bucket-relative end of interval k (columns)
    uint32_t dendP[FE_NK][FE_MAX_TOTAL + 1];   // dendP[k][b] = sum of dend[k][b'] over b' < b: columns of a bucket neighbourhood by difference
    uint32_t tile0[FE_NK * FE_MAX_TOTAL + 1];   // first blob tile of chunk k * nb + b; [nk * nb] = n_tiles
};

// One open range block of the level: its result slot (level position), its origin and its norm -- everything a search
// item needs about the block in one 16-byte load.
struct __align__(16) ListEntry {
    uint32_t slot;
    uint32_t xy;                          // x | y << 16
    uint32_t a2;                          // f16 kind: sum (4 r - 510)^2; i8 kind: 16 sum r^2
    uint32_t mirror;                      // 1: the block is read mirrored left-right (flip isometries: the second copy of a range block)
};

// One work item of a slice: a row tile (up to 32 range blocks = 128 rows) against a run of blob tiles.
struct ItemRec {
    uint32_t pos0;                        // first list position of the row tile
    uint32_t nrows;                       // valid rows (4 per range block); 0: empty item
    uint32_t t0, t1;                      // blob tiles [t0, t1)
    uint32_t cols_left;                   // runs inside one chunk: valid columns from tile t0 to the end of the chunk
    uint32_t a_tile;                      // i8 kind: row tile in the slice's A blob
    uint32_t pad_[2];
};

struct SliceCtl {
    // ---- state of the level ----
    uint32_t cnt[2][FE_MAX_TOTAL];        // open range blocks per bucket, per list
    uint32_t list;                        // list the next search reads
    uint32_t k_done;                      // intervals 0 .. k_done-1 have been scanned
    uint32_t done;                        // the slicing is over
    uint32_t open;                        // range blocks may still be without their first hit
    uint32_t passes;                      // search launches planned so far (slices + minimum pass)
    uint32_t slices;                      // of those, slices of the scan
    uint32_t min_ran;                     // the minimum pass was planned
    uint32_t ticket;
    uint32_t cutoff;                      // hits at domain indices below this are final
    unsigned long long evaluated;         // (range, domain, rotation) candidates scored
    // ---- the slice about to run ----
    uint32_t active;                      // ordinal of the planned slice: the expand / search launches of that ordinal run it
    uint32_t overflow;                    // the item buffer was too small (host sizing bug): the level is invalid
    uint32_t k0, k1;                      // intervals of the slice
    uint32_t run_len;                     // runs are cut into work items of about this many column tiles
    uint32_t whole_group;                 // runs cover every bucket of the group (minimum pass)
    uint32_t no_min;
    uint32_t n_items;
    uint32_t n_row_tiles;
    uint32_t item_prefix[FE_MAX_TOTAL + 1];   // work items of the buckets before b
    uint32_t tile_prefix[FE_MAX_TOTAL + 1];   // row tiles of the buckets before b (i8 kind: index into the slice's A blob)
};

// Level constants the planner needs (kernel parameter).
struct PlanArgs {
    LevelPlan* plan;
    SliceCtl* ctl;
    ListEntry* list[2];
    ItemRec* items;
    const uint32_t* rowhit;
    const uint16_t* pos_bucket;           // [level position] bucket of the range block
    uint32_t N;                           // T * T
    uint32_t use_thr, need_min, bins, multipass;
    uint32_t min_tiles;                   // column tiles a bucket advances per interval at least
    uint32_t max_items;
    uint32_t n_sm;
    uint32_t min_run, items_per_sm;       // work items: never shorter than min_run column tiles, about items_per_sm per SM
    uint32_t tiles_per_item;              // row tiles of a work item: 1, or 2 (i8 kind, M = 256)
};

enum { FE_PHASE_SLICE = 0, FE_PHASE_MIN = 1 };

__global__ void k_level_plan(PlanArgs a, const uint32_t* dom_hist, const uint32_t* rng_hist, const uint32_t* pre, uint32_t nb, uint32_t nbins,
                             uint32_t ngroups, uint32_t span, uint32_t nD, uint32_t nR, uint32_t nt);
__global__ void k_level_ranges(const uint8_t* img, uint32_t stride, const fe_grid_item* rng, const uint32_t* order, const LevelPlan* plan,
                               uint32_t T, int centred, int flips, ListEntry* list0, uint16_t* pos_bucket);
__global__ void k_slice_plan(PlanArgs a, int phase, uint32_t ordinal);
__global__ void k_expand_items(PlanArgs a, uint32_t ordinal);

// Device-scheduled tcgen05 search of one level (fe_plan.cu): enqueues everything on the ctx stream, never synchronises.
struct DeviceLevel {
    const fe_grid_item* d_dom; uint32_t nD;
    const fe_grid_item* d_rng; uint32_t nR;
    LevelGeom g;
    const int32_t* dom_cls;               // NULL: no classifier (one class)
    const int32_t* rng_cls;
    uint32_t thr16;
    bool use_thr, need_min, timed;
    bool flips;                           // d_rng holds every range block twice (2 i, 2 i + 1): the odd copy is searched mirrored
    const uint32_t* cells;                // lattice levels: sums of the T x T cells of the image ([h / T][cells_w]), else NULL
    const uint32_t* cells2;               // ... and of their squared pixels (src == tgt: a range block on the lattice is a cell)
    const uint32_t* cellsD2;              // ... and of their squared 2 x 2 box sums (sum D^2 of a domain = its four cells)
    uint32_t cells_w, dnx;                // domain d sits at cell (d % dnx, d / dnx)
};
struct DeviceLevelState;


// Everything the host reads back after a level, in one record (one D2H copy, one synchronisation per level).
struct LevelSummary {
    uint32_t mismatch, fp32_regime, flags, passes;
    unsigned long long evaluated;
    uint32_t last_scan, last_flag;
    uint32_t slices, min_ran;
    uint32_t overflow, done;              // done: the slicing is over and no minimum pass is outstanding
    unsigned long long matches;           // nominal candidates of the level: sum over classes of ranges x domains x 4
    unsigned long long lb_candidates;     // lower-bound prefilter: candidates it let through (scored exactly)
};
__global__ void k_level_summary(const SliceCtl* ctl, const LevelPlan* plan, const uint32_t* counters, const uint32_t* scan_last,
                                const uint32_t* split_last, uint32_t wants_min_pass, LevelSummary* out);

// kind::f16 kernel, device-scheduled (fe_search_f16.cu)
struct F16Args {
    const uint8_t* img; uint32_t stride;  // range image
    const void* B16;
    const uint4* colmeta;
    const uint32_t* blob_dom;
    const ListEntry* list[2];
    const ItemRec* items;
    const SliceCtl* ctl;
    unsigned long long* rowbest;
    uint32_t* rowhit;
    uint32_t* flags;
    uint32_t thr16, use_thr;
    uint32_t ordinal;                     // slice ordinal this launch belongs to
    uint32_t pair;                        // work items are two row tiles (PlanArgs::tiles_per_item == 2)
    // lower-bound prefilter (fe_lb.cu): A tiles come from a blob, candidates go to a list
    const void* A16;                      // [row tile][K / 8][128 rows][8 halves]
    uint2* cand;                          // {result row, domain index}
    uint32_t* cand_count;
    uint32_t cand_cap;
};
int f16_level_supported(const LevelGeom& g);   // fast geometry (S = 2T, even domain origins), T = 4 or 8
int f16_build_pool(fe_ctx* ctx, const LevelGeom& g, const fe_grid_item* d_dom, const uint32_t* dom_order, const LevelPlan* plan, uint32_t max_tiles);
int f16_launch_search(fe_ctx* ctx, const LevelGeom& g, const F16Args& a, bool retire, bool meta, cudaEvent_t ev0, cudaEvent_t ev1);
int f16_launch_search_lb(fe_ctx* ctx, const F16Args& a, cudaEvent_t ev0, cudaEvent_t ev1);

// kind::i8 kernel, device-scheduled (fe_search_i8.cu)
struct I8Args {
    const void* A8;                       // A blob of the slice: [row tile][K / 16][128 rows][16 B]
    const void* B8;                       // B blob of the level: [tile][K stage][plane lo, hi][kc / 16][64 columns][16 B]
    const uint32_t* tileseg;              // [tile] chunk of the tile
    const uint32_t* blob_dom;             // [tile][64] domain index of every blob column
    const uint32_t* coln;                 // [tile][64] sum D^2 (INT_MAX: padding column)
    const ListEntry* list[2];
    const ItemRec* items;
    const SliceCtl* ctl;
    unsigned long long* rowbest;
    uint32_t* rowhit;
    uint32_t thr16, use_thr, meta;
    uint32_t Kpad, stages, n_abuf;
    uint32_t ordinal;
    uint32_t pair;                        // work items are pairs of row tiles (rows 0-127 / 128-255 of the A tile): one compute warpgroup each
};
int i8_level_supported(const LevelGeom& g);    // fast geometry, 4 <= T <= 32
struct DeviceLevel;
int i8_build_pool(fe_ctx* ctx, const LevelGeom& g, const DeviceLevel& lv, const uint32_t* dom_order, const LevelPlan* plan, uint32_t nD,
                  uint32_t max_tiles);
int i8_build_rows(fe_ctx* ctx, const LevelGeom& g, const LevelPlan* plan, const SliceCtl* ctl, const ListEntry* const list[2], uint32_t ordinal,
                  uint32_t max_row_tiles);
int i8_launch_search(fe_ctx* ctx, const LevelGeom& g, I8Args a, cudaEvent_t ev0, cudaEvent_t ev1);
uint32_t i8_kpad(const LevelGeom& g);

// lower-bound prefilter (fe_lb.cu)
struct LbState {
    const uint16_t* Q = nullptr;          // c x c cell sums of the image
    uint32_t qw = 0, c = 0, s = 0;
    uint2* cand = nullptr;
    uint32_t* cand_count = nullptr;
    uint32_t cand_cap = 0;
};
uint32_t lb_threshold(uint32_t thr16);
int lb_prepare(fe_ctx* ctx, const LevelGeom& g, const fe_grid_item* d_dom, const uint32_t* dom_order, const LevelPlan* plan, uint32_t nR,
               ListEntry* list0, uint32_t max_tiles, LbState* lb);
int lb_build_rows(fe_ctx* ctx, const LbState& lb, const LevelPlan* plan, const SliceCtl* ctl, const ListEntry* const list[2], uint32_t ordinal,
                  uint32_t max_row_tiles);
int lb_verify(fe_ctx* ctx, const LbState& lb, const fe_grid_item* d_dom, const fe_grid_item* d_rng, const uint32_t* rng_order, uint32_t thr16,
              const SliceCtl* ctl, uint32_t ordinal);

// State of one device-scheduled level between its enqueue calls.
struct DeviceLevelState {
    bool bins = false;                    // brightness bins are on (inside the classes when there are classes)
    uint32_t nbins = 1, span = 0, ngroups = 1;
    uint32_t n_launches = 0;              // search launches enqueued (events ctx->ev_pass[2 i], [2 i + 1] when timed)
    double host_us_first_search = 0;      // FE_PASS_TIMES: host time from the level's first launch to its first search launch
    const uint32_t* dom_order = nullptr;  // sorted position -> domain index (NULL: identity)
    const uint32_t* rng_order = nullptr;  // level position -> range index (NULL: identity)
    // continuation
    int kind = 0;                         // 0: kind::f16, 1: kind::i8, 2: lower-bound prefilter on the f16 kernel + exact verification
    LbState lb;
    const fe_grid_item* d_dom = nullptr;
    const fe_grid_item* d_rng = nullptr;
    uint32_t thr16 = 0;
    bool multipass = false, retire = false, timed = false, wants_min_pass = false;
    uint32_t slices_enqueued = 0, max_row_tiles = 0, nR = 0;
    bool min_enqueued = false;
    LevelGeom g;
    PlanArgs pa{};
    F16Args fa{};
    I8Args ia{};
};
// Prepares the level (buckets, plan, operand blob) and enqueues `n_slices` slices (0: as many as the scan can need) and, when
// `with_min`, the minimum pass.  search_level_more enqueues what a level turned out to need beyond that.
int search_level_device(fe_ctx* ctx, const DeviceLevel& lv, int kind, uint32_t n_slices, bool with_min, DeviceLevelState* st);
int search_level_more(fe_ctx* ctx, DeviceLevelState* st);
