// fe_plan.cu -- device-side planning of a search level (see fe_plan.cuh): bucket layout, slice schedule, survivor
// compaction and work-item expansion, all from device memory; and the host code that enqueues one level.
#include <chrono>
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "fe_kernels.cuh"
#include "fe_plan.cuh"
#include "fe_umma.cuh"

namespace {

constexpr uint32_t GR = 128;   // slice granularity in columns: whole tiles of both kinds

__device__ __forceinline__ uint32_t grp_lo(const LevelPlan* p, uint32_t c, bool whole) {
    const uint32_t g0 = (c / p->nbins) * p->nbins, b = c - g0;
    return whole ? g0 : g0 + (b > p->span ? b - p->span : 0u);
}
__device__ __forceinline__ uint32_t grp_hi(const LevelPlan* p, uint32_t c, bool whole) {
    const uint32_t g0 = (c / p->nbins) * p->nbins, b = c - g0;
    return whole ? g0 + p->nbins - 1 : g0 + min(p->nbins - 1, b + p->span);
}

// block-wide sums over 256 threads
__device__ __forceinline__ unsigned long long block_sum_u64(unsigned long long v, unsigned long long* sh) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned long long t = 0;
    for (uint32_t w = 0; w < (blockDim.x >> 5); ++w) t += sh[w];
    return t;
}
// exclusive prefix sums over the block (blockDim <= 1024, sh: 33 words); *total = sum of all; two barriers
template <typename V>
__device__ __forceinline__ V block_excl_scan(V v, V* sh, V* total) {
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    V inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        const V x = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= (uint32_t)o) inc += x;
    }
    __syncthreads();                       // sh may still be read by the previous call
    if (lane == 31) sh[w] = inc;
    __syncthreads();
    V base = 0, tot = 0;
    for (uint32_t i = 0; i < nw; ++i) {
        const V x = sh[i];
        if (i < w) base += x;
        tot += x;
    }
    if (total) *total = tot;
    return base + inc - v;
}
} // namespace

// key = (class + 1) * nbins + bin (class 0 when cls is NULL, bin 0 when bins is NULL); hist[key] counts (per-block
// histogram in shared memory first: a few dozen hot global addresses would serialise the whole grid).
// Both lists in one go (positions 0..nD-1 the domains, nD.. the range blocks; bins likewise): the range keys carry `list_bit`, so
// ONE stable radix sort orders both lists, and vals = the index inside the own list.  hist_d / hist_r as before.
// cells != NULL (lattice levels, one image): the brightness bin is computed here from the level's cell sums -- a domain is the
// 2 x 2 cells at (i % dnx, i / dnx), a range block the cell at its origin (sum of 4 r = 4 x the cell) -- bin_width = the bin width.
__global__ void __launch_bounds__(256) k_bucket_keys(const int32_t* __restrict__ dom_cls, const int32_t* __restrict__ rng_cls, const uint8_t* __restrict__ bins,
                                                     uint32_t nD, uint32_t nR, uint32_t nbins, uint32_t nb, uint32_t list_bit, uint16_t* __restrict__ keys,
                                                     uint32_t* __restrict__ vals, uint32_t* __restrict__ hist_d, uint32_t* __restrict__ hist_r,
                                                     const uint32_t* __restrict__ cells, uint32_t cells_w, uint32_t dnx, const fe_grid_item* __restrict__ rng,
                                                     uint32_t T, uint32_t bin_width) {
    __shared__ uint32_t sh[2 * FE_MAX_TOTAL];
    for (uint32_t b = threadIdx.x; b < 2 * FE_MAX_TOTAL; b += blockDim.x) sh[b] = 0;
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nD + nR; i += gridDim.x * blockDim.x) {
        const bool is_rng = i >= nD;
        const uint32_t j = is_rng ? i - nD : i;
        const int32_t* cls = is_rng ? rng_cls : dom_cls;
        uint32_t bin = bins ? (uint32_t)bins[i] : 0u;
        if (cells) {
            uint32_t s;
            if (is_rng) {
                const fe_grid_item r = rng[j];
                s = 4u * __ldg(cells + (size_t)(r.y / T) * cells_w + r.x / T);
            } else {
                const uint32_t* c = cells + (size_t)(j / dnx) * cells_w + j % dnx;
                s = __ldg(c) + __ldg(c + 1) + __ldg(c + cells_w) + __ldg(c + cells_w + 1);
            }
            bin = min(s / bin_width, (uint32_t)FE_MAX_BUCKETS - 1);
        }
        const uint32_t k = (cls ? (uint32_t)(cls[j] + 1) : 0u) * nbins + bin;
        keys[i] = (uint16_t)(k | (is_rng ? list_bit : 0u));
        vals[i] = j;
        atomicAdd(&sh[k + (is_rng ? FE_MAX_TOTAL : 0)], 1u);
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < nb; b += blockDim.x) {
        if (sh[b]) atomicAdd(&hist_d[b], sh[b]);
        if (sh[FE_MAX_TOTAL + b]) atomicAdd(&hist_r[b], sh[FE_MAX_TOTAL + b]);
    }
}

// One CTA of 512 threads: bucket offsets, the interval ends of every bucket, the tile layout of the operand blob and the
// initial slice state.  pre[b * 8 + k] = positions of bucket b with a domain index below cut[k] (k_bin_prefix).
__global__ void __launch_bounds__(512) k_level_plan(PlanArgs a, const uint32_t* __restrict__ dom_hist, const uint32_t* __restrict__ rng_hist,
                                                    const uint32_t* __restrict__ pre, uint32_t nb, uint32_t nbins, uint32_t ngroups, uint32_t span,
                                                    uint32_t nD, uint32_t nR, uint32_t nt) {
    __shared__ uint32_t s_scan[33];
    LevelPlan* p = a.plan;
    SliceCtl* ctl = a.ctl;
    const uint32_t t = threadIdx.x;
    const uint32_t dc = t < nb ? dom_hist[t] : 0u, rc = t < nb ? rng_hist[t] : 0u;
    // exclusive scans of the two histograms
    for (int pass = 0; pass < 2; ++pass) {
        const uint32_t v = pass ? rc : dc;
        const uint32_t ex = block_excl_scan(v, s_scan, (uint32_t*)nullptr);
        uint32_t* dst = pass ? p->roff : p->doff;
        if (t < nb) dst[t + 1] = ex + v;
        if (t == 0) dst[0] = 0;
    }
    // interval ends of my bucket
    const uint32_t min_step = a.min_tiles * GR;
    uint32_t tiles[FE_NK];
    uint32_t done = 0;
#pragma unroll
    for (int k = 0; k < FE_NK; ++k) {
        uint32_t hi = dc;
        if (a.multipass && k < FE_NK - 1) {
            const uint32_t target = pre[t < nb ? t * 8 + k : 0];
            const uint64_t up = ((uint64_t)max(min_step, target) + GR - 1) / GR * GR;
            hi = (uint32_t)min((uint64_t)dc, max(up, (uint64_t)done + min_step));
            if (dc - hi < min_step / 2) hi = dc;     // no slivers at the end of the scan
        }
        if (t < nb) p->dend[k][t] = hi;
        {   // prefix over the buckets: the slice planner takes neighbourhood sums by difference
            const uint32_t v = t < nb ? hi : 0u;
            const uint32_t ex = block_excl_scan(v, s_scan, (uint32_t*)nullptr);
            if (t < nb) p->dendP[k][t + 1] = ex + v;
            if (t == 0) p->dendP[k][0] = 0;
        }
        tiles[k] = t < nb ? (hi - done + nt - 1) / nt : 0u;
        done = hi;
    }
    // tile layout, interval-major: exclusive scan over (k, b)
    uint32_t base = 0;
    const uint32_t nk = a.multipass ? FE_NK : 1;
#pragma unroll
    for (uint32_t k = 0; k < FE_NK; ++k) {
        if (k >= nk) break;
        uint32_t tot = 0;
        const uint32_t ex = block_excl_scan(tiles[k], s_scan, &tot);
        if (t < nb) p->tile0[k * nb + t] = base + ex;
        base += tot;
    }
    if (t == 0) {
        p->tile0[nk * nb] = base;
        p->n_tiles = base;
        p->nb = nb; p->nbins = nbins; p->ngroups = ngroups; p->span = span;
        p->nD = nD; p->nR = nR; p->nt = nt; p->nk = nk;
        for (int k = 0; k < FE_NK; ++k)
            p->cut[k] = !a.multipass ? FE_NONE32 : (k == FE_NK - 1 ? nD : (uint32_t)((((uint64_t)nD) << k) / 128 + 1));
        ctl->list = 0; ctl->k_done = 0; ctl->done = 0; ctl->open = 1; ctl->passes = 0; ctl->ticket = 0; ctl->slices = 0; ctl->min_ran = 0;
        ctl->cutoff = FE_NONE32; ctl->evaluated = 0ull; ctl->active = FE_NONE32; ctl->n_items = 0; ctl->no_min = 0;
        ctl->overflow = 0;
    }
    for (uint32_t b = t; b < FE_MAX_TOTAL; b += 512) {
        ctl->cnt[0][b] = b < nb ? rng_hist[b] : 0u;
        ctl->cnt[1][b] = 0;
    }
}

// Per level position: the first list (every range block open), sum a^2 of the block (centred: a = 4 r - 510, f16 kind;
// else 16 sum r^2, i8 kind) and the bucket of the position.
// cells / cells2 (lattice levels): the block is cell (x / T, y / T) of the image's T-cell sums -- no pixel is read, one thread per block.
__global__ void k_level_ranges(const uint8_t* __restrict__ img, uint32_t stride, const fe_grid_item* __restrict__ rng,
                               const uint32_t* __restrict__ order, const LevelPlan* __restrict__ plan, uint32_t T, int centred, int flips,
                               const uint32_t* __restrict__ cells, const uint32_t* __restrict__ cells2, uint32_t cells_w,
                               ListEntry* __restrict__ list0, uint16_t* __restrict__ pos_bucket, unsigned long long* __restrict__ rowbest,
                               uint32_t* __restrict__ rowhit) {
    // a warp per block from T = 16 on (coalesced rows), a thread per block below
    const uint32_t lanes = (T >= 16 && !cells2) ? 32u : 1u;
    const uint32_t gt = blockIdx.x * blockDim.x + threadIdx.x, p = gt / lanes, lane = gt % lanes;
    if (p >= plan->nR) return;
    const uint32_t idx = order ? order[p] : p;
    const fe_grid_item r = rng[idx];
    uint32_t s2 = 0;
    if (cells2) {
        const size_t c = (size_t)(r.y / T) * cells_w + r.x / T;
        const uint32_t q = 16u * __ldg(cells2 + c);                     // 16 sum r^2 <= 16 * 1024 * 255^2 < 2^31
        s2 = centred ? q - 4080u * __ldg(cells + c) + T * T * 260100u : q;   // sum (4 r - 510)^2
    } else {
        const uint8_t* base = img + (size_t)r.y * stride + r.x;
        for (uint32_t e = lane; e < T * T; e += lanes) {
            const uint32_t y = e / T, x = e - y * T;
            const int v = centred ? 4 * (int)base[(size_t)y * stride + x] - 510 : 4 * (int)base[(size_t)y * stride + x];
            s2 += (uint32_t)(v * v);
        }
        if (lanes == 32)
            for (int o = 16; o; o >>= 1) s2 += __shfl_xor_sync(0xFFFFFFFFu, s2, o);
    }
    if (lane) return;
    ListEntry e;
    e.slot = p;
    e.xy = r.x | (r.y << 16);
    e.a2 = s2;
    e.mirror = flips ? (idx & 1u) : 0u;
    list0[p] = e;
    // the four result rows of the position start empty
    reinterpret_cast<ulonglong2*>(rowbest)[2 * p] = make_ulonglong2(FE_INF64, FE_INF64);
    reinterpret_cast<ulonglong2*>(rowbest)[2 * p + 1] = make_ulonglong2(FE_INF64, FE_INF64);
    reinterpret_cast<uint4*>(rowhit)[p] = make_uint4(FE_NONE32, FE_NONE32, FE_NONE32, FE_NONE32);
    uint32_t lo = 0, hi = plan->nb - 1;               // bucket b with roff[b] <= p < roff[b + 1]
    while (lo < hi) {
        const uint32_t mid = (lo + hi + 1) >> 1;
        if (plan->roff[mid] <= p) lo = mid; else hi = mid - 1;
    }
    pos_bucket[p] = (uint16_t)lo;
}

// Survivors of the slice that just ran, then the plan of the next one.
//  phase SLICE (ordinal s): no-op once the slicing is over.  First call: every range block is open.
//  phase MIN: after the slicing, the range blocks still without a proven first hit get one pass over every domain of their
//             class (first hit and minimum) -- only threshold searches with brightness bins that want the minimum.
// A range block stays open while none of its four rows has a hit below the cutoff the last slice scanned to.
__global__ void __launch_bounds__(256) k_slice_plan(PlanArgs a, int phase, uint32_t ordinal) {
    LevelPlan* p = a.plan;
    SliceCtl* ctl = a.ctl;
    __shared__ unsigned long long s_red[8];
    __shared__ uint32_t s_last;
    __shared__ uint32_t s_cnt[FE_MAX_TOTAL], s_tiles[FE_MAX_TOTAL];
    __shared__ uint32_t s_k0, s_k1, s_go;
    if (phase == FE_PHASE_SLICE && ctl->done) return;
    if (phase == FE_PHASE_MIN && !(ctl->done && ctl->open && a.use_thr && a.bins && a.need_min)) return;
    const uint32_t cur = ctl->list;
    const bool first = phase == FE_PHASE_SLICE && ctl->slices == 0;
    const uint32_t nxt = first ? cur : cur ^ 1u;
    if (!first) {
        // ---- compaction: the open range blocks of list `cur` that stay open go to list `nxt`, bucket regions kept ----
        const uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
        bool alive = false;
        uint32_t b = 0;
        ListEntry e{};
        if (pos < p->nR) {
            b = a.pos_bucket[pos];
            if (pos - p->roff[b] < ctl->cnt[cur][b]) {
                e = a.list[cur][pos];
                const uint4 h = reinterpret_cast<const uint4*>(a.rowhit)[e.slot];
                alive = min(min(h.x, h.y), min(h.z, h.w)) >= ctl->cutoff;
            }
            // minimum pass: a range block meets every domain of its class, so the survivors of a class share row tiles --
            // they are gathered in the region of the class's first bucket
            if (phase == FE_PHASE_MIN) b = (b / p->nbins) * p->nbins;
        }
        // Append per bucket, aggregated per block: positions are sorted by bucket, so a block holds one or two buckets -- ranks
        // come from shared-memory counters, one global atomic per block and bucket reserves the space.
        __shared__ uint32_t s_bmin, s_n[32], s_base[32];
        if (threadIdx.x == 0) {
            const uint32_t p0 = min(blockIdx.x * blockDim.x, p->nR - 1);
            uint32_t b0 = a.pos_bucket[p0];
            if (phase == FE_PHASE_MIN) b0 = (b0 / p->nbins) * p->nbins;
            s_bmin = b0;
        }
        if (threadIdx.x < 32) s_n[threadIdx.x] = 0;
        __syncthreads();
        const uint32_t d = b - s_bmin;
        uint32_t rank = 0;
        if (alive) rank = d < 32 ? atomicAdd(&s_n[d], 1u) : atomicAdd(&ctl->cnt[nxt][b], 1u);
        __syncthreads();
        if (threadIdx.x < 32 && s_n[threadIdx.x]) s_base[threadIdx.x] = atomicAdd(&ctl->cnt[nxt][s_bmin + threadIdx.x], s_n[threadIdx.x]);
        __syncthreads();
        if (alive) a.list[nxt][p->roff[b] + (d < 32 ? s_base[d] : 0u) + rank] = e;
    }
    // ---- last CTA plans ----
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&ctl->ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const uint32_t t = threadIdx.x, nb = p->nb, nk = p->nk;
    unsigned long long before = 0, S = 0;
    for (uint32_t b = t; b < nb; b += 256) {
        const uint32_t c = __ldcg(&ctl->cnt[nxt][b]);
        s_cnt[b] = c;
        S += c;
        before += __ldcg(&ctl->cnt[cur][b]);
    }
    S = block_sum_u64(S, s_red);
    before = block_sum_u64(before, s_red);
    const uint32_t kd = ctl->k_done;
    const bool whole = phase == FE_PHASE_MIN;
    // work left for the open range blocks if everything that remains were scanned now
    unsigned long long left = 0;
    for (uint32_t c = t; c < nb; c += 256) {
        if (!s_cnt[c]) continue;
        const uint32_t lo = grp_lo(p, c, whole), hi = grp_hi(p, c, whole);     // neighbourhood sums by difference of prefixes
        const unsigned long long cols = (p->doff[hi + 1] - p->doff[lo]) - ((kd && !whole) ? p->dendP[kd - 1][hi + 1] - p->dendP[kd - 1][lo] : 0u);
        left += cols * s_cnt[c];
    }
    left = block_sum_u64(left, s_red);
    if (t == 0) {
        uint32_t go = 1, k0 = 0, k1 = nk - 1;
        if (phase == FE_PHASE_SLICE) {
            if (!first) {
                const double resolved = 1.0 - (double)S / (double)max(before, 1ull);
                if (S == 0) { go = 0; ctl->open = 0; }
                // Bins only pay when ranges end on a hit: first slices that close next to nothing say "no hits on this
                // level" -- the minimum pass will scan everything anyway, go there now.
                else if (a.bins && a.need_min && ctl->passes <= 2 && resolved < 0.05) go = 0;
                else if (kd >= nk) go = 0;
                else {
                    // slices that close little grow faster; one that closes nothing says early-out will not pay
                    const uint32_t step = resolved >= 0.03 ? 1u : (resolved >= 0.002 ? 2u : 8u);
                    k0 = kd;
                    k1 = min(nk - 1, kd + step - 1);
                    // what is left is small: one more slice for all of it costs less than several
                    if ((double)left * 4.0 * 2.0 * (double)a.N <= 1.5e11) k1 = nk - 1;
                }
            } else {
                k0 = 0;
                k1 = a.multipass ? 0u : nk - 1;
                if ((double)left * 4.0 * 2.0 * (double)a.N <= 1.5e11) k1 = nk - 1;
            }
            if (!go) ctl->done = 1;
        } else {
            if (S == 0) go = 0;
            ctl->open = 0;
        }
        s_go = go; s_k0 = k0; s_k1 = k1;
    }
    __syncthreads();
    const uint32_t k0 = s_k0, k1 = s_k1;
    if (s_go) {
        // ---- candidates and tile steps (row tiles x column tiles) of the slice ----
        unsigned long long work = 0, steps = 0;
        for (uint32_t c = t; c < nb; c += 256) {
            unsigned long long cols = 0, run = 0;
            if (s_cnt[c]) {
                const uint32_t lo = grp_lo(p, c, whole), hi = grp_hi(p, c, whole);
                cols = (p->dendP[k1][hi + 1] - p->dendP[k1][lo]) - (k0 ? p->dendP[k0 - 1][hi + 1] - p->dendP[k0 - 1][lo] : 0u);
#pragma unroll
                for (uint32_t k = 0; k < FE_NK; ++k)      // unrolled and predicated: the loads of all intervals are in flight together
                    if (k >= k0 && k <= k1) run += p->tile0[k * nb + hi + 1] - p->tile0[k * nb + lo];
            }
            s_tiles[c] = cols ? (s_cnt[c] + 31) / 32 : 0u;
            work += cols * s_cnt[c] * 4ull;
            steps += run * ((s_tiles[c] + a.tiles_per_item - 1) / a.tiles_per_item);
        }
        work = block_sum_u64(work, s_red);
        steps = block_sum_u64(steps, s_red);
        // Work items of about equal length: long runs are cut so that every SM gets eight items or more, short runs stay
        // whole (an item pays for its A tile and for filling the pipeline: never below 16 column tiles).
        const uint32_t run_len = (uint32_t)min(max(steps / ((unsigned long long)a.items_per_sm * a.n_sm), (unsigned long long)a.min_run), 1ull << 20);
        __syncthreads();
        for (uint32_t c = t; c < nb; c += 256) {
            uint32_t per_tile = 0;
            if (s_tiles[c]) {
                const uint32_t lo = grp_lo(p, c, whole), hi = grp_hi(p, c, whole);
#pragma unroll
                for (uint32_t k = 0; k < FE_NK; ++k)
                    if (k >= k0 && k <= k1) {
                        const uint32_t nrun = p->tile0[k * nb + hi + 1] - p->tile0[k * nb + lo];
                        per_tile += (nrun + run_len - 1) / run_len;
                    }
            }
            s_cnt[c] = ((s_tiles[c] + a.tiles_per_item - 1) / a.tiles_per_item) * per_tile;   // work items of the bucket (s_cnt is not needed any more)
        }
        __syncthreads();
        // prefixes of the work items and the row tiles over the buckets: thread t owns buckets 2 t and 2 t + 1
        unsigned long long acc = 0;
        uint32_t tiles = 0;
        {
            const uint32_t c0 = 2 * t, c1 = 2 * t + 1;
            const unsigned long long i0 = c0 < nb ? s_cnt[c0] : 0u, i1 = c1 < nb ? s_cnt[c1] : 0u;
            const uint32_t t0 = c0 < nb ? s_tiles[c0] : 0u, t1 = c1 < nb ? s_tiles[c1] : 0u;
            const unsigned long long ei = block_excl_scan(i0 + i1, s_red, &acc);
            const uint32_t et = block_excl_scan(t0 + t1, reinterpret_cast<uint32_t*>(s_red), &tiles);
            if (c0 < nb) { ctl->item_prefix[c0] = (uint32_t)min(ei, (unsigned long long)0xFFFFFFFFull); ctl->tile_prefix[c0] = et; }
            if (c1 < nb) { ctl->item_prefix[c1] = (uint32_t)min(ei + i0, (unsigned long long)0xFFFFFFFFull); ctl->tile_prefix[c1] = et + t0; }
        }
        if (t == 0) {
            ctl->item_prefix[nb] = (uint32_t)min(acc, (unsigned long long)0xFFFFFFFFull);
            ctl->tile_prefix[nb] = tiles;
            if (acc > a.max_items) ctl->overflow = 1;
            ctl->n_row_tiles = tiles;
            ctl->k0 = k0; ctl->k1 = k1; ctl->run_len = run_len;
            ctl->whole_group = whole ? 1u : 0u;
            ctl->no_min = (a.use_thr && (!a.need_min || (a.bins && !whole))) ? 1u : 0u;
            ctl->n_items = (uint32_t)min(acc, (unsigned long long)a.max_items);
            ctl->evaluated += work;
            ctl->passes += 1;
            ctl->active = ordinal;
            if (phase == FE_PHASE_MIN) ctl->min_ran = 1;
            if (phase == FE_PHASE_SLICE) {
                ctl->slices += 1;
                ctl->k_done = k1 + 1;
                ctl->cutoff = (k1 == nk - 1) ? FE_NONE32 : p->cut[k1];
                if (k1 == nk - 1) ctl->done = 1;       // nothing left to slice (the launches of later ordinals return at once)
            }
        }
    }
    // the list that was read becomes the target of the next compaction
    if (!first)
        for (uint32_t b = t; b < FE_MAX_TOTAL; b += 256) ctl->cnt[cur][b] = 0;
    if (t == 0) { ctl->list = nxt; ctl->ticket = 0; }
}

// Work items of the planned slice: (bucket c, row tile, interval k, piece of the run), one record each, in that order.
__global__ void k_expand_items(PlanArgs a, uint32_t ordinal) {
    const LevelPlan* p = a.plan;
    const SliceCtl* ctl = a.ctl;
    if (ctl->active != ordinal) return;
    const uint32_t n = ctl->n_items, L = ctl->run_len, k0 = ctl->k0, k1 = ctl->k1, nb = p->nb, nt = p->nt;
    const bool whole = ctl->whole_group != 0;
    const uint32_t cur = ctl->list;
    for (uint32_t id = blockIdx.x * blockDim.x + threadIdx.x; id < n; id += gridDim.x * blockDim.x) {
        uint32_t lo = 0, hi = nb - 1;                 // bucket c with item_prefix[c] <= id < item_prefix[c + 1]
        while (lo < hi) {
            const uint32_t mid = (lo + hi + 1) >> 1;
            if (ctl->item_prefix[mid] <= id) lo = mid; else hi = mid - 1;
        }
        const uint32_t c = lo, blo = grp_lo(p, c, whole), bhi = grp_hi(p, c, whole);
        const uint32_t tpi = a.tiles_per_item, rpi = 32 * tpi;          // row tiles / range blocks per work item
        const uint32_t cnt = ctl->cnt[cur][c], ntile = (cnt + rpi - 1) / rpi;
        const uint32_t per_tile = (ctl->item_prefix[c + 1] - ctl->item_prefix[c]) / ntile;
        uint32_t rem = id - ctl->item_prefix[c];
        const uint32_t rt = rem / per_tile;
        rem -= rt * per_tile;
        ItemRec r{};
        for (uint32_t k = k0; k <= k1; ++k) {
            const uint32_t T0 = p->tile0[k * nb + blo], nrun = p->tile0[k * nb + bhi + 1] - T0, q = (nrun + L - 1) / L;
            if (rem < q) {
                r.t0 = T0 + (uint32_t)(((uint64_t)rem * nrun) / q);
                r.t1 = T0 + (uint32_t)(((uint64_t)(rem + 1) * nrun) / q);
                r.pos0 = p->roff[c] + rpi * rt;
                r.nrows = 4 * min(rpi, cnt - rpi * rt);
                r.cols_left = (p->dend[k][blo] - (k ? p->dend[k - 1][blo] : 0u)) - (r.t0 - T0) * nt;
                r.a_tile = ctl->tile_prefix[c] + tpi * rt;
                break;
            }
            rem -= q;
        }
        a.items[id] = r;
    }
}

__global__ void k_level_summary(const SliceCtl* ctl, const LevelPlan* plan, const uint32_t* counters, const uint32_t* scan_last,
                                const uint32_t* split_last, uint32_t wants_min_pass, LevelSummary* out) {
    LevelSummary s{};
    s.mismatch = counters[0]; s.fp32_regime = counters[1]; s.flags = counters[2];
    s.lb_candidates = *reinterpret_cast<const unsigned long long*>(counters + 4);
    s.done = 1;
    if (ctl) {
        s.passes = ctl->passes; s.evaluated = ctl->evaluated; s.overflow = ctl->overflow;
        s.slices = ctl->slices; s.min_ran = ctl->min_ran;
        s.done = (ctl->done && !(ctl->open && wants_min_pass)) ? 1u : 0u;
    }
    if (plan)
        for (uint32_t g = 0; g < plan->ngroups; ++g) {
            const uint32_t b0 = g * plan->nbins, b1 = b0 + plan->nbins;
            s.matches += (unsigned long long)(plan->roff[b1] - plan->roff[b0]) * (plan->doff[b1] - plan->doff[b0]) * 4ull;
        }
    if (scan_last) { s.last_scan = *scan_last; s.last_flag = *split_last; }
    *out = s;
}

// ---------------------------------------------------------------------------------------------------
// host side: enqueue one level
// ---------------------------------------------------------------------------------------------------
static inline uint32_t cdiv_u(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }

void fe_plan_bins(uint32_t N, uint32_t thr16, fe_threshold_plan* pl);   // fe_api.cu

#define PLAUNCH(ctx, kernel, grid, block, ...)                         \
    do {                                                               \
        kernel<<<(grid), (block), 0, (ctx)->stream>>>(__VA_ARGS__);    \
        (ctx)->stats.kernel_launches++;                                \
        FE_CUDA(ctx, cudaGetLastError());                              \
    } while (0)

static int enqueue_slice(fe_ctx* ctx, DeviceLevelState* st, int phase, uint32_t ordinal, bool meta) {
    const PlanArgs& pa = st->pa;
    const LevelGeom& g = st->g;
    PLAUNCH(ctx, k_slice_plan, cdiv_u(st->nR, 256), 256, pa, phase, ordinal);
    PLAUNCH(ctx, k_expand_items, 2 * ctx->n_sm, 256, pa, ordinal);
    cudaEvent_t e0 = st->timed ? ctx->ev_pass[2 * st->n_launches] : nullptr, e1 = st->timed ? ctx->ev_pass[2 * st->n_launches + 1] : nullptr;
    if (st->n_launches == 0) st->host_us_first_search = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - ctx->level_host_t0).count();
    if (st->kind == 0) {
        st->fa.ordinal = ordinal;
        FE_TRY(f16_launch_search(ctx, g, st->fa, st->retire, meta, e0, e1));
    } else if (st->kind == 2) {
        // lower-bound prefilter: cell-sum rows of the open range blocks, the 8 x 8 contraction emitting candidates, exact check
        const ListEntry* lists[2] = {pa.list[0], pa.list[1]};
        FE_TRY(lb_build_rows(ctx, st->lb, pa.plan, pa.ctl, lists, ordinal, st->max_row_tiles));
        st->fa.ordinal = ordinal;
        FE_TRY(f16_launch_search_lb(ctx, st->fa, e0, e1));
        FE_TRY(lb_verify(ctx, st->lb, st->d_dom, st->d_rng, st->rng_order, st->thr16, pa.ctl, ordinal));
    } else {
        const ListEntry* lists[2] = {pa.list[0], pa.list[1]};
        FE_TRY(i8_build_rows(ctx, g, pa.plan, pa.ctl, lists, ordinal, st->max_row_tiles));
        st->ia.ordinal = ordinal;
        st->ia.meta = meta ? 1u : 0u;
        FE_TRY(i8_launch_search(ctx, g, st->ia, e0, e1));
    }
    ++st->n_launches;
    if (getenv("FE_PASS_DBG")) {                      // debugging aid: synchronises after every launch
        SliceCtl* h = new SliceCtl;
        cudaStreamSynchronize(ctx->stream);
        cudaMemcpy(h, pa.ctl, sizeof(SliceCtl), cudaMemcpyDeviceToHost);
        float ms = 0;
        if (e0) cudaEventElapsedTime(&ms, e0, e1);
        if (st->kind == 2) {
            uint32_t nc = 0;
            cudaMemcpy(&nc, st->lb.cand_count, 4, cudaMemcpyDeviceToHost);
            fprintf(stderr, "[slice] lower-bound prefilter: %u candidates\n", nc);
        }
        fprintf(stderr, "[slice] T=%u kind=%d ordinal=%u ran=%d k=%u..%u run_len=%u items=%u row_tiles=%u passes=%u evaluated=%.3e done=%u open=%u kernel %.3f ms\n",
                g.T, st->kind, ordinal, h->active == ordinal, h->k0, h->k1, h->run_len, h->n_items, h->n_row_tiles, h->passes, (double)h->evaluated, h->done,
                h->open, ms);
        delete h;
    }
    return FE_OK;
}

// What the level still needs after the host has seen that the enqueued slices did not finish it: the rest of the slice
// train and the minimum pass.
int search_level_more(fe_ctx* ctx, DeviceLevelState* st) {
    if ((size_t)st->n_launches + FE_NK + 1 > (size_t)FE_MAX_LAUNCHES) return fe_fail(ctx, FE_ERR_CUDA, "internal: search launch budget exceeded");
    const uint32_t n_slices = st->multipass ? (uint32_t)FE_NK : 1u;
    for (uint32_t s = st->slices_enqueued; s < n_slices; ++s) FE_TRY(enqueue_slice(ctx, st, FE_PHASE_SLICE, s, st->span > 0));
    st->slices_enqueued = n_slices;
    if (st->wants_min_pass) FE_TRY(enqueue_slice(ctx, st, FE_PHASE_MIN, FE_NK + (st->min_enqueued ? 1u : 0u), true));
    st->min_enqueued = true;
    return FE_OK;
}

int search_level_device(fe_ctx* ctx, const DeviceLevel& lv, int kind, uint32_t n_slices_hint, bool with_min, DeviceLevelState* st) {
    const LevelGeom& g = lv.g;
    const uint32_t nD = lv.nD, nR = lv.nR;
    const uint32_t nt = kind == 1 ? (uint32_t)I8_NT : (uint32_t)UM_NT;
    const bool multipass = lv.use_thr && !getenv("FE_SINGLE_PASS");
    *st = DeviceLevelState{};
    // ---- brightness bins (host arithmetic only) ----
    uint32_t width = 0;
    if (multipass && !getenv("FE_NO_BINS")) {
        fe_threshold_plan pl{};
        fe_plan_bins(g.N, lv.thr16, &pl);
        if (pl.n_bins) { st->bins = true; st->nbins = pl.n_bins; st->span = pl.bin_span; width = pl.bin_width; }
    }
    st->ngroups = lv.dom_cls ? 7u : 1u;
    const uint32_t nb = st->ngroups * st->nbins;
    if (nb > (uint32_t)FE_MAX_TOTAL) return fe_fail(ctx, FE_ERR_UNSUPPORTED, "internal: %u buckets", nb);
    const bool sorted = nb > 1;
    const uint32_t max_tiles = cdiv_u(nD, nt) + nb + 1;
    const uint32_t max_items = (nR / 32 + nb) * FE_NK + 8 * (uint32_t)ctx->n_sm + 64;   // row tiles x intervals + pieces of long runs

    // ---- buffers ----
    const size_t n = (size_t)nD + nR;
    FE_CUDA(ctx, ctx->b_keys_tmp.ensure(n * 6 + 512));
    FE_CUDA(ctx, ctx->b_vals_tmp.ensure(n * 4));
    FE_CUDA(ctx, ctx->b_dom_order2.ensure(n * 4 + 16));     // both orders: the domains', then the range blocks' 
    FE_CUDA(ctx, ctx->b_hist.ensure((size_t)FE_MAX_TOTAL * 10 * 4 + 2 * FE_MAX_BUCKETS * 4));
    FE_CUDA(ctx, ctx->b_plan.ensure(sizeof(LevelPlan)));
    FE_CUDA(ctx, ctx->b_ctl.ensure(sizeof(SliceCtl)));
    FE_CUDA(ctx, ctx->b_list[0].ensure((size_t)nR * sizeof(ListEntry) + 16));
    FE_CUDA(ctx, ctx->b_list[1].ensure((size_t)nR * sizeof(ListEntry) + 16));
    FE_CUDA(ctx, ctx->b_itemrec.ensure((size_t)max_items * sizeof(ItemRec)));
    FE_CUDA(ctx, ctx->b_posb.ensure((size_t)nR * 2 + 16));
    uint8_t* bins8 = ctx->b_keys_tmp.as<uint8_t>();
    uint16_t* keys = reinterpret_cast<uint16_t*>(bins8 + ((n + 255) & ~(size_t)255));
    uint16_t* keys_out = keys + n;
    uint32_t* hist_d = ctx->b_hist.as<uint32_t>();
    uint32_t* hist_r = hist_d + FE_MAX_TOTAL;
    uint32_t* pre = hist_d + 2 * FE_MAX_TOTAL;
    uint32_t* scratch = pre + FE_MAX_TOTAL * 8;
    FE_CUDA(ctx, cudaMemsetAsync(hist_d, 0, ((size_t)FE_MAX_TOTAL * 10 + 2 * FE_MAX_BUCKETS) * 4, ctx->stream));

    // ---- bucket keys, orders ----
    const bool bins_from_cells = st->bins && lv.cells && lv.cells2;       // both lists on the lattice of one image: bins inside k_bucket_keys
    if (st->bins && !bins_from_cells) {
        if (lv.cells) launch_dom_from_cells(ctx->stream, lv.cells, lv.cells_w, lv.dnx, nD, nullptr, width, bins8, scratch);
        else launch_brightness_bins(ctx->stream, ctx->src.px, ctx->src.stride, lv.d_dom, nD, g.S, 1u, width, bins8, scratch);
        launch_brightness_bins(ctx->stream, ctx->tgt.px, ctx->tgt.stride, lv.d_rng, nR, g.T, 4u, width, bins8 + nD, scratch + FE_MAX_BUCKETS);
        ctx->stats.kernel_launches += 2;
        FE_CUDA(ctx, cudaGetLastError());
    }
    int bits = 1;
    while ((1u << bits) < nb) ++bits;
    PLAUNCH(ctx, k_bucket_keys, std::min(cdiv_u(n, 1024), 4u * ctx->n_sm), 256, lv.dom_cls, lv.rng_cls, (st->bins && !bins_from_cells) ? bins8 : nullptr, nD, nR, st->nbins,
            nb, 1u << bits, keys, ctx->b_vals_tmp.as<uint32_t>(), hist_d, hist_r, bins_from_cells ? lv.cells : nullptr, lv.cells_w, lv.dnx, lv.d_rng, g.T, width);
    if (sorted) {
        // one stable sort for both lists: the list bit is the top key bit
        size_t tmp = 0;
        FE_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp, keys, keys_out, ctx->b_vals_tmp.as<uint32_t>(), ctx->b_dom_order2.as<uint32_t>(), (int)n, 0, bits + 1, ctx->stream));
        FE_CUDA(ctx, ctx->b_sort_tmp.ensure(tmp));
        FE_CUDA(ctx, cub::DeviceRadixSort::SortPairs(ctx->b_sort_tmp.p, tmp, keys, keys_out, ctx->b_vals_tmp.as<uint32_t>(), ctx->b_dom_order2.as<uint32_t>(), (int)n, 0, bits + 1, ctx->stream));
        ctx->stats.kernel_launches += 3;
        st->dom_order = ctx->b_dom_order2.as<uint32_t>();
        st->rng_order = ctx->b_dom_order2.as<uint32_t>() + nD;
    }
    BucketOff c8{};
    for (int k = 0; k < FE_NK; ++k) c8.v[k] = !multipass ? FE_NONE32 : (k == FE_NK - 1 ? nD : (uint32_t)((((uint64_t)nD) << k) / 128 + 1));
    PLAUNCH(ctx, k_bin_prefix, cdiv_u((uint64_t)nb * 8, 128), 128, st->dom_order, hist_d, (int)nb, c8, pre);

    // ---- the level's plan, first list, operand blob ----
    PlanArgs& pa = st->pa;
    pa = PlanArgs{};
    pa.plan = ctx->b_plan.as<LevelPlan>();
    pa.ctl = ctx->b_ctl.as<SliceCtl>();
    pa.list[0] = ctx->b_list[0].as<ListEntry>();
    pa.list[1] = ctx->b_list[1].as<ListEntry>();
    pa.items = ctx->b_itemrec.as<ItemRec>();
    pa.rowhit = ctx->b_rowhit.as<uint32_t>();
    pa.pos_bucket = ctx->b_posb.as<uint16_t>();
    pa.N = g.N;
    pa.use_thr = lv.use_thr; pa.need_min = lv.need_min; pa.bins = st->bins; pa.multipass = multipass;
    const char* ms_env = getenv("FE_MIN_STEP");      // tuning: column tiles (of 128) a bucket advances per interval at least
    pa.min_tiles = ms_env ? (uint32_t)std::max(1, atoi(ms_env)) : (st->bins ? (uint32_t)std::max(2, 10 / (2 * (int)st->span + 1)) : 16u);
    pa.max_items = max_items;
    pa.n_sm = (uint32_t)ctx->n_sm;
    // work items: the i8 kind reloads its A tile per item (128 KB at T = 32, one buffer): longer and fewer items there
    pa.min_run = kind != 1 ? 16u : (g.T >= 32 ? 48u : 24u);
    pa.items_per_sm = kind != 1 ? 8u : (g.T >= 32 ? 4u : 6u);
    if (const char* e = getenv(kind == 1 ? "FE_I8_ITEMS" : "FE_F16_ITEMS")) pa.items_per_sm = (uint32_t)std::max(1, atoi(e));   // tuning
    if (const char* e = getenv(kind == 1 ? "FE_I8_MIN_RUN" : "FE_F16_MIN_RUN")) pa.min_run = (uint32_t)std::max(1, atoi(e));
    // i8 kind up to T = 16: a work item is TWO row tiles (64 range blocks, M = 256 through two accumulator pairs) sharing every
    // B stage -- halves the L2 -> shared-memory operand stream, which bounds the kernel before the tensor pipe does
    // The kind::f16 search does the same at T = 8 (fe_search_f16.cu, PAIR; measured: T = 8 -2 %, T = 4 +14 % -- its short K makes
    // the epilogue, not the operand stream, the bound, and half as many work items balance worse).
    pa.tiles_per_item = (((kind == 1 && g.T <= 16) || (kind == 0 && g.T == 8)) && !getenv("FE_NO_PAIR")) ? 2u : 1u;
    PLAUNCH(ctx, k_level_plan, 1, 512, pa, hist_d, hist_r, pre, nb, st->nbins, st->ngroups, st->span, nD, nR, nt);
    PLAUNCH(ctx, k_level_ranges, cdiv_u((uint64_t)nR * ((g.T >= 16 && !lv.cells2) ? 32 : 1), 128), 128, ctx->tgt.px, ctx->tgt.stride, lv.d_rng, st->rng_order, pa.plan, g.T,
            kind == 1 ? 0 : 1, lv.flips ? 1 : 0, lv.cells, lv.cells2, lv.cells_w, pa.list[0], ctx->b_posb.as<uint16_t>(),
            ctx->b_rowbest.as<unsigned long long>(), ctx->b_rowhit.as<uint32_t>());
    FE_CUDA(ctx, cudaMemsetAsync(ctx->b_counters.as<uint32_t>() + 2, 0, 2 * sizeof(uint32_t), ctx->stream));
    F16Args& fa = st->fa;
    I8Args& ia = st->ia;
    fa = F16Args{};
    ia = I8Args{};
    const uint32_t max_row_tiles = nR / 32 + nb;
    st->d_dom = lv.d_dom; st->d_rng = lv.d_rng; st->thr16 = lv.thr16;
    if (kind == 0 || kind == 2) {
        if (kind == 0) FE_TRY(f16_build_pool(ctx, g, lv.d_dom, st->dom_order, pa.plan, max_tiles));
        else FE_TRY(lb_prepare(ctx, g, lv.d_dom, st->dom_order, pa.plan, nR, pa.list[0], max_tiles, &st->lb));
        fa.img = ctx->tgt.px; fa.stride = ctx->tgt.stride;
        fa.B16 = ctx->b_B16.p;
        fa.colmeta = ctx->b_tmaps.as<uint4>();
        fa.blob_dom = ctx->b_blob_dom.as<uint32_t>();
        fa.list[0] = pa.list[0]; fa.list[1] = pa.list[1];
        fa.items = pa.items;
        fa.ctl = pa.ctl;
        fa.rowbest = ctx->b_rowbest.as<unsigned long long>();
        fa.rowhit = ctx->b_rowhit.as<uint32_t>();
        fa.flags = ctx->b_counters.as<uint32_t>() + 2;
        fa.thr16 = lv.thr16; fa.use_thr = lv.use_thr ? 1u : 0u;
        fa.pair = (kind == 0 && pa.tiles_per_item == 2) ? 1u : 0u;
        if (kind == 2) {
            FE_CUDA(ctx, ctx->b_A16.ensure((size_t)max_row_tiles * UM_ROWS * 80 * 2 + 256));
            fa.A16 = ctx->b_A16.p;
            fa.cand = st->lb.cand; fa.cand_count = st->lb.cand_count; fa.cand_cap = st->lb.cand_cap;
            fa.thr16 = lb_threshold(lv.thr16);
        }
    } else {
        FE_TRY(i8_build_pool(ctx, g, lv, st->dom_order, pa.plan, nD, max_tiles));
        FE_CUDA(ctx, ctx->b_A16.ensure((size_t)(max_row_tiles + 2) * UM_ROWS * i8_kpad(g) + 256));
        ia.A8 = ctx->b_A16.p;
        ia.B8 = ctx->b_B16.p;
        ia.tileseg = ctx->b_tileseg.as<uint32_t>();
        ia.blob_dom = ctx->b_blob_dom.as<uint32_t>();
        ia.coln = ctx->b_coln.as<uint32_t>();
        ia.list[0] = pa.list[0]; ia.list[1] = pa.list[1];
        ia.items = pa.items;
        ia.ctl = pa.ctl;
        ia.rowbest = ctx->b_rowbest.as<unsigned long long>();
        ia.rowhit = ctx->b_rowhit.as<uint32_t>();
        ia.thr16 = lv.thr16; ia.use_thr = lv.use_thr ? 1u : 0u;
        ia.pair = pa.tiles_per_item == 2 ? 1u : 0u;
    }
    st->kind = kind;
    st->multipass = multipass;
    st->retire = lv.use_thr && g.T == 4;
    st->timed = lv.timed;
    st->wants_min_pass = st->bins && lv.need_min;
    st->max_row_tiles = max_row_tiles;
    st->nR = nR;
    st->g = g;
    const uint32_t n_slices = multipass ? (n_slices_hint ? std::min(n_slices_hint, (uint32_t)FE_NK) : (uint32_t)FE_NK) : 1u;
    for (uint32_t s = 0; s < n_slices; ++s) FE_TRY(enqueue_slice(ctx, st, FE_PHASE_SLICE, s, st->span > 0));
    st->slices_enqueued = n_slices;
    if (st->wants_min_pass && with_min) {
        FE_TRY(enqueue_slice(ctx, st, FE_PHASE_MIN, FE_NK, true));
        st->min_enqueued = true;
    }
    return FE_OK;
}
