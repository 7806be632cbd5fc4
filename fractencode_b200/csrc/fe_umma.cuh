// fe_umma.cuh -- declarations of the tcgen05 search path (fe_search_umma.cu).
#pragma once
#include <algorithm>

#include "fe_internal.cuh"

constexpr int UM_NT = 128;         // domain columns per tile (UMMA N)
constexpr int UM_ROWS = 128;       // rows per tile = 32 ranges x 4 rotations (UMMA M)
constexpr int UM_WGS = 2;          // compute warpgroups (each owns 2 TMEM accumulators of UM_NT columns)
constexpr int UM_HALF = UM_NT / 2;   // columns a compute thread holds per register set
// MMA-issuer warps.  kind::i8 kernel: one issuer thread per accumulator buffer (4): the issue loop of a tile is latency bound
// (barrier probes ~200 cycles, commit ~170) and four loops overlap.  The kind::f16 kernel's compute threads need 96 registers:
// a 21st warp would put six warps on one SM sub-partition (16384 registers) and cap them at 80, so there the four issuers
// are also their own producers and no producer warp exists (20 warps).
constexpr int UM_ISSUERS_F16 = 2 * UM_WGS;
constexpr int UM_ISSUERS_I8 = 2 * UM_WGS;
constexpr int UM_THREADS_F16 = 32 * UM_ISSUERS_F16 + 256 * UM_WGS;      // issuer warps (each its own producer), 8 compute warps per group
constexpr int UM_THREADS_I8 = 32 + 32 * UM_ISSUERS_I8 + 256 * UM_WGS;
constexpr int UM_MAX_STAGES = 8;
constexpr int UM_STAGES = 4;       // B stages: tile t reuses the stage of tile t-4, freed by that tile's accumulator-full commit
constexpr int UM_MAX_NK = 5;
constexpr int I8_NT = 64;          // i8 kind: domain columns per tile (two s32 accumulators, low/high byte plane, of 64 columns)
constexpr int I8_KC = 256;         // i8 kind: bytes of K per shared-memory stage
constexpr int I8_MAX_STAGES = 4;       // K steps of 16 per tile: T=4 -> 2, T=8 -> 5

struct UmmaBucket {
    uint32_t row_tile0, n_row_tiles; // A blobs of this classifier bucket
    uint32_t col_tile0, n_col_tiles; // B blobs
    uint32_t row0, nrows;            // first global row (= 4 * first range position), valid rows
    uint32_t ncols;                  // valid columns of the bucket's column tiles (meaningful when it meets one domain bucket)
    uint32_t chunks;                 // column chunks per row tile (work items = n_row_tiles * chunks)
};

struct UmmaArgs {
    const void* A16;
    const void* B16;
    const uint4* colmeta;            // f16 kind: [col tile][2 halves] {parity of sum(b^2) of columns 0-31, 32-63, valid columns of
                                     // the half, domain bucket of the tile}
    const uint32_t* tileseg;         // i8 kind: [col tile] domain bucket of the tile
    const uint32_t* blob_dom;        // [col tile][nt] domain index of every blob column (FE_NONE32: padding)
    const uint32_t* rowA2;           // [range position] sum(a^2)
    unsigned long long* rowbest;
    uint32_t* rowhit;
    uint32_t* flags;                 // bit 0: a winner sits in the fp32-inexact band
    UmmaBucket b[FE_MAX_BUCKETS];
    uint32_t item_end[FE_MAX_BUCKETS]; // work items of the buckets 0..i (running total)
    int nb;
    uint32_t Kpad, stages, total_items, thr16, use_thr;
    uint32_t nt;                     // domain columns per tile (UM_NT for the f16 kind, I8_NT for the i8 kind)
    const uint32_t* coln;            // i8 kind: [sorted column] sum(D^2) (0x3FFFFFFF for padding columns)
    uint32_t n_abuf;                 // i8 kind: A buffers in shared memory (2, or 1 when the tile is 128 KB)
    uint32_t dbg;                    // tuning probes (FE_UMMA_DBG): 1 skip TMEM drain, 2 skip MMA issue, 4 skip B copies
    const uint32_t* rowslot;         // [range position of this pass] -> range position of the level (result slot); NULL = identity
    uint32_t meta;                   // work items may cross domain buckets: per-tile bucket ids matter
    uint32_t no_min;                 // only threshold hits are wanted (levels that split: a range without a hit is split and its
                                     // minimum never read) -- skip the running-minimum bookkeeping
};

struct UmmaBuckets {                 // operand-layout view for the blob builders
    uint32_t range_off[FE_MAX_BUCKETS + 1], row_tile0[FE_MAX_BUCKETS + 1], col_tile0[FE_MAX_BUCKETS + 1]; // [nb] entries are the totals
    uint32_t dom_off[FE_MAX_BUCKETS + 1], dom_end[FE_MAX_BUCKETS + 1];   // slice of domain positions of each bucket
    uint32_t n_ranges, n_domains;
    int nb;
};

// One search pass of a level: every range position of the pass (grouped by classifier bucket) against a slice of the
// bucket's domain positions.  A level is one pass over everything, or -- with a threshold -- several passes over growing
// slices of the scan, each over the ranges that have not met the threshold yet (the reference's `break` at the first
// candidate under the threshold, TransformEstimator2.hpp:40-41, at pass granularity).
struct SearchPass {
    const uint32_t* dom_order;       // domain position -> domain item index (NULL = identity)
    const uint32_t* rng_items;       // range position of this pass -> range item index (NULL = identity)
    const uint32_t* rowslot;         // range position of this pass -> range position of the level (NULL = identity)
    uint32_t dbeg[FE_MAX_BUCKETS], dend[FE_MAX_BUCKETS]; // per DOMAIN bucket: domain positions searched in this pass
    int span;                        // range bucket c meets the domain buckets c-span .. c+span (0: its own; 1: brightness bins)
    uint32_t roff[FE_MAX_BUCKETS + 1];                   // per bucket: range positions of this pass (prefix offsets)
    int nbuckets;
    uint32_t n_dom;                  // all domain positions of the level
    bool no_min;                     // only threshold hits are wanted
    bool reuse_rows;                 // the A blob and row norms built by the previous pass are still valid
    bool reuse_dom_norms;            // i8 kind: the per-position domain norms of the level are already built
    cudaEvent_t ev0, ev1;            // recorded around the search launch (NULL: not timed)
};

int umma_i8_level_supported(const LevelGeom& g);
// Both enqueue the operand builders and the search kernel on the ctx stream and return without synchronising; the f16
// kind raises bit 0 of ctx->b_counters[2] when a winner sits in the fp32-inexact band (the caller re-runs on the i8 kind).
int umma_i8_prepare_and_search(fe_ctx* ctx, const LevelGeom& g, const fe_grid_item* d_dom, const fe_grid_item* d_rng, const SearchPass& sp,
                               uint32_t thr16, bool use_thr);
int umma_prepare_and_search(fe_ctx* ctx, const LevelGeom& g, const fe_grid_item* d_dom, const fe_grid_item* d_rng, const SearchPass& sp,
                            uint32_t thr16, bool use_thr);
