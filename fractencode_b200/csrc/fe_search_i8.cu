// fe_search_i8.cu -- tcgen05 kind::i8 search of the large range blocks (T >= 16) and exact fallback of the fp16 kind,
// scheduled from device memory (fe_plan.cuh).
//
// fp32 accumulation stops being integer-exact when K = T^2 grows (sum r*D reaches 2.7e8 at T=32), so the large levels
// use the integer tensor path of sm_100a:  A = range pixels r (u8, inverse-rotated rows), B = the domain box sums D
// split into a low-byte plane and a high-byte plane (D <= 1020 -> high byte 0..3), two s32 accumulators per tile:
//     cross = acc_lo + 256 * acc_hi,      n16 = 16 sum r^2 - 8 cross + sum D^2      (exact, any input)
// Two planes at the 2x int8 rate cost the same tensor time as one fp16 pass.  Tile = 128 rows x 64 domain columns
// (2 x 64 TMEM columns per buffer, two buffers per warpgroup, two warpgroups = all 512 columns); K is streamed in
// stages of 256 bytes (32 KB per stage: both planes), the A tile (128 x K bytes, up to 128 KB at T=32) stays in
// shared memory for the whole work item.
//
// Warp roles (672 threads): warp 0 = producer (lane 0: bulk copies of the A tile and the B stages; lane 1: forwarder of
// the landed-stage count), warps 1-4 = one tcgen05.mma issuer thread per accumulator buffer, warps 5-20 = two compute
// warpgroups (thread = one row x 32 columns: tcgen05.ld of both planes -> IMAD combine -> VIMNMX3 argmin).
// Work items and slice parameters come from device memory (ItemRec / SliceCtl), the B operand is one blob per level,
// the A operand one blob per slice built from the list of open range blocks (k_build_rows_i8).
#include <cstdio>
#include <cstdlib>

#include "fe_kernels.cuh"
#include "fe_plan.cuh"
#include "fe_umma_dev.cuh"

using namespace umma_dev;

namespace {

__global__ void __launch_bounds__(UM_THREADS_I8, 1) k_search_i8(const I8Args a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const SliceCtl* __restrict__ ctl = a.ctl;
    if (ctl->active != a.ordinal) return;                  // this slice was never planned (the level ended earlier)
    const uint32_t n_items = ctl->n_items;
    if (blockIdx.x >= n_items) return;
    const uint32_t no_min = ctl->no_min;
    const ListEntry* __restrict__ list = a.list[ctl->list];
    const uint32_t Kpad = a.Kpad;
    const uint32_t kc = min(Kpad, (uint32_t)I8_KC);          // bytes of K per stage
    const uint32_t nch = Kpad / kc;                          // stages per tile
    const uint32_t pair = a.pair;                            // M = 256: two row tiles per item, warpgroup g owns row tile g
    const uint32_t bytesA1 = UM_ROWS * Kpad, bytesA = (pair ? 2u : 1u) * bytesA1, bytesB = 2 * I8_NT * kc;
    const uint32_t S = a.stages, NA = a.n_abuf;
    uint8_t* sA = smem;
    uint8_t* sB = smem + NA * bytesA;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)S * bytesB);
    const uint32_t bar0 = smem_u32(bars);
    auto A_FULL = [&](uint32_t i) { return bar0 + 8 * (0 + i); };
    auto A_EMPTY = [&](uint32_t i) { return bar0 + 8 * (2 + i); };
    auto ACC_FULL = [&](uint32_t g, uint32_t b) { return bar0 + 8 * (4 + 2 * g + b); };
    auto ACC_EMPTY = [&](uint32_t g, uint32_t b) { return bar0 + 8 * (8 + 2 * g + b); };
    auto B_FULL = [&](uint32_t i) { return bar0 + 8 * (12 + i); };
    auto B_EMPTY = [&](uint32_t i) { return bar0 + 8 * (12 + I8_MAX_STAGES + i); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12 + 2 * I8_MAX_STAGES);
    const uint32_t B_LANDED = smem_u32(tmem_slot + 4); // number of B stages that have landed, published in order by the forwarder lane

    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < 2; ++i) {
            mbar_init(A_FULL(i), 1);
            mbar_init(A_EMPTY(i), UM_ISSUERS_I8);
        }
        for (uint32_t i = 0; i < 4; ++i) {
            mbar_init(bar0 + 8 * (4 + i), 1);
            mbar_init(bar0 + 8 * (8 + i), 8);
        }
        for (uint32_t i = 0; i < S; ++i) {
            mbar_init(B_FULL(i), 1);
            mbar_init(B_EMPTY(i), pair ? 2 : 1);          // pair: the issuers of both row tiles release a stage
        }
        tmem_slot[4] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // record of work item w: {pos0, nrows, t0, t1}
    auto item_rec = [&](uint32_t w) { return __ldg(reinterpret_cast<const uint4*>(a.items + w)); };

    if (warp == 0) {
        // ================= producer =================
        if (lane == 0) {
            uint32_t ic = 0, wi = 0; // running stage counter, item counter
            for (uint32_t w = blockIdx.x; w < n_items; w += gridDim.x, ++wi) {
                const uint4 rec = item_rec(w);
                const uint32_t a_tile = __ldg(&a.items[w].a_tile);
                if (w + gridDim.x < n_items) asm volatile("prefetch.global.L1 [%0];" ::"l"(a.items + w + gridDim.x));
                const uint32_t ab = wi % NA;
                mbar_wait(A_EMPTY(ab), ((wi / NA) & 1) ^ 1);
                mbar_expect_tx(A_FULL(ab), bytesA);
                // the bulk copy engine takes at most ~1 MB per request; 128 KB tiles go as 32 KB pieces
                for (uint32_t off = 0; off < bytesA; off += 32768) {
                    const uint32_t n = min(32768u, bytesA - off);
                    bulk_g2s(smem_u32(sA + ab * bytesA + off), reinterpret_cast<const uint8_t*>(a.A8) + (size_t)a_tile * bytesA1 + off, n, A_FULL(ab));
                }
                for (uint32_t t = rec.z; t < rec.w; ++t)
                    for (uint32_t c = 0; c < nch; ++c, ++ic) {
                        const uint32_t s = ic % S;
                        mbar_wait(B_EMPTY(s), ((ic / S) & 1) ^ 1);
                        mbar_expect_tx(B_FULL(s), bytesB);
                        bulk_g2s(smem_u32(sB + (size_t)s * bytesB), reinterpret_cast<const uint8_t*>(a.B8) + ((size_t)t * nch + c) * bytesB, bytesB, B_FULL(s));
                    }
            }
        } else if (lane == 1) {
            // forwarder: a stage ring shared by two issuers means an issuer may look at a stage barrier that is still
            // TWO phases behind the one it needs, which a parity wait cannot tell from "done".  This lane observes
            // every phase in order (never ambiguous) and publishes the count of landed stages.
            uint32_t ic = 0;
            for (uint32_t w = blockIdx.x; w < n_items; w += gridDim.x) {
                const uint4 rec = item_rec(w);
                const uint32_t total = (rec.w - rec.z) * nch;
                for (uint32_t q = 0; q < total; ++q, ++ic) {
                    mbar_wait(B_FULL(ic % S), (ic / S) & 1);
                    flag_store_release(B_LANDED, ic + 1);
                }
            }
        }
    } else if (warp <= UM_ISSUERS_I8) {
        // ================= MMA issuers: one thread per accumulator buffer (g, ib) =================
        if (lane == 0) {
            const uint32_t g = (warp - 1) >> 1, ib = (warp - 1) & 1;
            // D = S32, A = B = unsigned 8 bit, K-major both, N = 64, M = 128
            const uint32_t idesc = (2u << 4) | ((uint32_t)(I8_NT >> 3) << 17) | ((uint32_t)(UM_ROWS >> 4) << 24);
            const uint32_t nj = kc / 32; // MMAs (K = 32 bytes) per stage and plane
            uint32_t it0 = 0, wi = 0, jb = 0;
            for (uint32_t w = blockIdx.x; w < n_items; w += gridDim.x, ++wi) {
                const uint4 rec = item_rec(w);
                const uint32_t ab = wi % NA, n = rec.w - rec.z;
                // single tiles: the column tiles alternate between the warpgroups; pairs: every column tile goes to both (row tile g)
                const uint32_t first = pair ? 0u : (g + UM_WGS - (it0 % UM_WGS)) % UM_WGS, ustep = pair ? 1u : (uint32_t)UM_WGS;
                mbar_wait(A_FULL(ab), (wi / NA) & 1);
                const uint32_t a_addr = smem_u32(sA + ab * bytesA) + (pair ? g * bytesA1 : 0u);
                bool any = false;
                for (uint32_t u = first; u < n; u += ustep, ++jb) {
                    const uint32_t gi = it0 + u, buf = jb & 1;
                    if (buf != ib) continue;
                    const uint32_t d_lo = tmem_base + (g * 2 + buf) * 2 * I8_NT, d_hi = d_lo + I8_NT;
                    mbar_wait(ACC_EMPTY(g, buf), ((jb >> 1) & 1) ^ 1);
                    for (uint32_t c = 0; c < nch; ++c) {
                        const uint32_t ic = gi * nch + c, s = ic % S;
                        flag_wait_ge(B_LANDED, ic + 1);       // the barrier has reached the phase we need ...
                        mbar_wait(B_FULL(s), (ic / S) & 1);   // ... so this parity probe is unambiguous (and is the formal acquire)
                        tc_fence_after();
                        const uint32_t b_addr = smem_u32(sB + (size_t)s * bytesB);
                        for (uint32_t j = 0; j < nj; ++j) {
                            // chunk-major blobs: a 16-byte K chunk of all rows is contiguous (rows * 16 bytes)
                            const uint64_t adesc = make_desc(a_addr + (c * (kc / 16) + 2 * j) * (UM_ROWS * 16), UM_ROWS * 16, 128);
                            const uint64_t blo = make_desc(b_addr + (2 * j) * (I8_NT * 16), I8_NT * 16, 128);
                            const uint64_t bhi = make_desc(b_addr + I8_NT * kc + (2 * j) * (I8_NT * 16), I8_NT * 16, 128);
                            const uint32_t acc = (c | j) ? 1u : 0u;
                            tc_mma<1>(d_lo, adesc, blo, idesc, acc);
                            tc_mma<1>(d_hi, adesc, bhi, idesc, acc);
                        }
                        tc_commit(B_EMPTY(s));
                    }
                    tc_commit(ACC_FULL(g, buf));
                    any = true;
                }
                if (any) tc_commit(A_EMPTY(ab)); else mbar_arrive(A_EMPTY(ab));
                it0 += n;
            }
        }
    } else {
        // ================= compute warps: thread = one row x 32 columns =================
        const uint32_t cw = warp - 1 - UM_ISSUERS_I8;
        const uint32_t g = cw >> 3;
        const uint32_t h = (cw >> 2) & 1;             // column half (32 columns)
        const uint32_t sp = warp & 3;
        const uint32_t lrow = sp * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((sp * 32u) << 16) + h * 32;
        uint32_t it0 = 0, jb = 0;
        for (uint32_t w = blockIdx.x; w < n_items; w += gridDim.x) {
            const uint4 rec = item_rec(w);
            const uint32_t item_t0 = rec.z, n = rec.w - rec.z;
            const uint32_t irow = pair ? g * UM_ROWS + lrow : lrow;     // row inside the item (pairs: 256 rows, warpgroup g has rows 128 g ..)
            const bool row_ok = irow < rec.y;
            uint32_t slot = 0, rc = 0;                // rc = 16 * sum r^2
            if (row_ok) {
                const uint4 ent = __ldg(reinterpret_cast<const uint4*>(list + rec.x + (irow >> 2)));
                slot = ent.x; rc = ent.z;
            }
            const uint32_t srow = 4u * slot + (lrow & 3u);   // result row of the level
            // w = sum D^2 - 8 cross (signed); n16 = rc + w;  n16 <= thr16  <=>  w <= thr16 - rc
            long long wt = (long long)a.thr16 - (long long)rc;
            wt = max(-2147483647ll, min(2147483646ll, wt));                  // INT_MAX is the padding columns' score
            const int wthr = (a.use_thr && row_ok) ? (int)wt : (int)0x80000000;
            int bestw = 0x7FFFFFFF;
            uint32_t bestcol = FE_NONE32, hit = FE_NONE32;
            uint32_t cur_seg = FE_NONE32;       // chunk of the tiles this thread is scanning
            const uint32_t first = pair ? 0u : (g + UM_WGS - (it0 % UM_WGS)) % UM_WGS, ustep = pair ? 1u : (uint32_t)UM_WGS;
            const uint32_t my_tiles = first < n ? (n - first + ustep - 1) / ustep : 0;
            for (uint32_t j = 0; j < my_tiles; ++j, ++jb) {
                const uint32_t u = first + j * ustep, buf = jb & 1;
                const uint32_t colbase = u * I8_NT + h * 32;             // column inside the item
                const uint32_t taddr = lane_addr + (g * 2 + buf) * 2 * I8_NT;
                // runs that cross chunks of the blob (brightness-bin neighbourhoods, minimum pass): chunk id of the tile, loaded
                // ahead of the accumulator wait and looked at after the accumulator has been read
                const uint32_t seg = a.meta ? __ldg(a.tileseg + item_t0 + u) : 0u;
                uint32_t lo[32], hi[32];
                mbar_wait(ACC_FULL(g, buf), (jb >> 1) & 1);
                tc_fence_after();
                TMEM_LD32(taddr, lo);
                TMEM_LD32(taddr + I8_NT, hi);
                tmem_wait_ld();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(ACC_EMPTY(g, buf));
                if (a.meta && seg != cur_seg) {
                    // another chunk: its columns restart at low domain indices, so bank the first hit and the running minimum of
                    // the chunk behind us (the 64-bit key settles ties by domain index) and go on
                    if (hit != FE_NONE32) {
                        atomicMin(&a.rowhit[srow], a.blob_dom[(size_t)item_t0 * I8_NT + hit]);
                        hit = FE_NONE32;
                    }
                    if (bestcol != FE_NONE32) {
                        atomicMin(&a.rowbest[srow], ((unsigned long long)(rc + (uint32_t)bestw) << 32) |
                                                        (unsigned long long)a.blob_dom[(size_t)item_t0 * I8_NT + bestcol]);
                        bestw = bestw == 0x7FFFFFFF ? bestw : bestw + 1;     // an equal score in the next chunk is still looked at
                        bestcol = FE_NONE32;
                    }
                    cur_seg = seg;
                }
                // sum(D^2) per column, stored by padded (tile, column) position so the 16-byte loads stay aligned
                const uint4* cn4 = reinterpret_cast<const uint4*>(a.coln + (size_t)(item_t0 + u) * I8_NT + h * 32);
                // w = sum D^2 - 8 (lo + 256 hi): two IMADs per column (FMA pipe), written over the low-plane registers
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint4 cn = __ldg(cn4 + q);
                    const uint32_t cnv[4] = {cn.x, cn.y, cn.z, cn.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int i = 4 * q + e;
                        lo[i] = (uint32_t)((int)cnv[e] - 8 * (int)lo[i] - 2048 * (int)hi[i]);
                    }
                }
                // tile minimum through 3-input integer minima (VIMNMX3), column located only when the row improves
                int m0 = 0x7FFFFFFF, m1 = 0x7FFFFFFF;
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    m0 = min(min(m0, (int)lo[i]), (int)lo[i + 1]);
                    m1 = min(min(m1, (int)lo[i + 2]), (int)lo[i + 3]);
                }
                const int tmin = min(m0, m1);
                const bool improve = row_ok && !no_min && tmin < bestw;
                const bool need_hit = row_ok && hit == FE_NONE32 && tmin <= wthr;
                if (improve | need_hit) {
                    uint32_t c_best = FE_NONE32, c_hit = FE_NONE32;
#pragma unroll
                    for (int i = 31; i >= 0; --i) {       // descending: the smallest qualifying column survives
                        c_best = ((int)lo[i] == tmin) ? (uint32_t)i : c_best;
                        c_hit = ((int)lo[i] <= wthr) ? (uint32_t)i : c_hit;
                    }
                    if (improve) { bestw = tmin; bestcol = colbase + c_best; }
                    if (need_hit && c_hit != FE_NONE32) hit = colbase + c_hit;
                }
            }
            it0 += n;
            if (row_ok) {
                if (bestcol != FE_NONE32) {
                    const uint32_t n16 = rc + (uint32_t)bestw;
                    const unsigned long long key = ((unsigned long long)n16 << 32) | (unsigned long long)a.blob_dom[(size_t)item_t0 * I8_NT + bestcol];
                    atomicMin(&a.rowbest[srow], key);
                }
                if (hit != FE_NONE32) atomicMin(&a.rowhit[srow], a.blob_dom[(size_t)item_t0 * I8_NT + hit]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------
// operand blobs (u8).  A: [row tile][K/16][128 rows][16 B].  B: [col tile][K stage][plane lo,hi][kc/16][64 cols][16 B].
// ---------------------------------------------------------------------------------------------------
// A: one CTA per row tile of the slice (32 open range blocks of one bucket, from the list).  The 32 blocks are staged in
// shared memory with coalesced word loads, then every thread assembles 16-byte K chunks of the four rotations from shared
// memory -- rotations 0 and 2 are straight / reversed 16-byte runs, 1 and 3 are strided byte gathers that never leave the
// SM (block stride padded by 16 bytes: the eight blocks a warp gathers from sit in different banks).  Every byte of the
// blob is written exactly once (padding rows as zeros): no memset.
template <uint32_t TT>   // TT = 0: block size only known at run time (T_rt), e.g. T = 6 or 12
__global__ void __launch_bounds__(256) k_build_rows_i8(const uint8_t* __restrict__ img, uint32_t stride, const LevelPlan* __restrict__ plan,
                                                       const SliceCtl* __restrict__ ctl, const ListEntry* __restrict__ list0,
                                                       const ListEntry* __restrict__ list1, uint32_t ordinal, uint32_t T_rt, uint32_t Kpad,
                                                       uint4* __restrict__ A8) {
    const uint32_t T = TT ? TT : T_rt;
    extern __shared__ __align__(16) uint8_t sblk[];            // [32][N + 16]
    __shared__ uint32_t sxy[32], smir[32];
    __shared__ uint32_t s_pos0, s_nvalid;
    if (ctl->active != ordinal) return;
    const uint32_t N = T * T, NS = N + 16;                     // T a template parameter: the index arithmetic below is shifts
    const uint32_t nch = Kpad / 16;
    const ListEntry* list = ctl->list ? list1 : list0;
    for (uint32_t tile = blockIdx.x; tile < ctl->n_row_tiles; tile += gridDim.x) {
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t lo = 0, hi = plan->nb - 1;               // bucket c with tile_prefix[c] <= tile < tile_prefix[c + 1]
        while (lo < hi) {
            const uint32_t mid = (lo + hi + 1) >> 1;
            if (ctl->tile_prefix[mid] <= tile) lo = mid; else hi = mid - 1;
        }
        const uint32_t rt = tile - ctl->tile_prefix[lo];
        s_pos0 = plan->roff[lo] + 32 * rt;
        s_nvalid = min(32u, ctl->cnt[ctl->list][lo] - 32 * rt);
    }
    __syncthreads();
    const uint32_t nvalid = s_nvalid;
    if (threadIdx.x < nvalid) {
        const ListEntry e = list[s_pos0 + threadIdx.x];
        sxy[threadIdx.x] = e.xy;
        smir[threadIdx.x] = e.mirror;
    }
    __syncthreads();
    // ---- stage the blocks ----
    const bool words = (T & 3u) == 0 && (stride & 3u) == 0 && (reinterpret_cast<uintptr_t>(img) & 3u) == 0;
    if (words) {
        const uint32_t wpr = T / 4, wpb = N / 4;              // words per block row / per block
        for (uint32_t idx = threadIdx.x; idx < nvalid * wpb; idx += blockDim.x) {
            const uint32_t lr = idx / wpb, e = idx - lr * wpb, y = e / wpr, xw = e - y * wpr;
            const uint32_t xy = sxy[lr], x = xy & 0xFFFFu;
            const uint8_t* p = img + (size_t)((xy >> 16) + y) * stride + x + 4 * xw;
            uint32_t v;
            if ((x & 3u) == 0) v = __ldg(reinterpret_cast<const uint32_t*>(p));
            else v = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
            if (smir[lr]) reinterpret_cast<uint32_t*>(sblk + lr * NS)[y * wpr + (wpr - 1 - xw)] = __byte_perm(v, 0, 0x0123);   // read right to left
            else reinterpret_cast<uint32_t*>(sblk + lr * NS)[e] = v;
        }
    } else {
        for (uint32_t idx = threadIdx.x; idx < nvalid * N; idx += blockDim.x) {
            const uint32_t lr = idx / N, e = idx - lr * N, y = e / T, x = e - y * T;
            const uint32_t xy = sxy[lr];
            sblk[lr * NS + y * T + (smir[lr] ? T - 1 - x : x)] = img[(size_t)((xy >> 16) + y) * stride + (xy & 0xFFFFu) + x];
        }
    }
    __syncthreads();
    // ---- chunks: [K/16][128 rows][16 B] ----
    uint4* out = A8 + (size_t)tile * nch * UM_ROWS;
    if constexpr (TT >= 16) {
        // A thread turns a 4 x 16 patch of a rotated block into four chunks.  A warp is one rotation of eight blocks x four
        // patches: no divergence, and the word reads of the odd rotations (16 rows of 4 pixels, transposed with PRMT) fall into
        // 32 different banks (block stride 68 words, patches one word apart).
        constexpr uint32_t nY4 = TT / 4, nX16 = TT / 16, units = UM_ROWS * nY4 * nX16;
        for (uint32_t idx = threadIdx.x; idx < units; idx += blockDim.x) {
            const uint32_t lr8 = idx & 7u, ysub = (idx >> 3) & 3u, rest = idx >> 5;
            const uint32_t k = rest & 3u, bg = (rest >> 2) & 3u, r2 = rest >> 4;
            const uint32_t Y = ((r2 % (nY4 / 4)) * 4 + ysub) * 4, X0 = (r2 / (nY4 / 4)) * 16;
            const uint32_t lr = bg * 8 + lr8, row = 4 * lr + k;
            uint32_t o[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int m = 0; m < 4; ++m) o[i][m] = 0;
            if (lr < nvalid) {
                const uint8_t* sb = sblk + lr * NS;
                if (k == 0) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint4 v = *reinterpret_cast<const uint4*>(sb + (Y + i) * TT + X0);
                        o[i][0] = v.x; o[i][1] = v.y; o[i][2] = v.z; o[i][3] = v.w;
                    }
                } else if (k == 2) {                           // element e of the rotated block = element N-1-e of the block
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint4 v = *reinterpret_cast<const uint4*>(sb + (TT - 1 - Y - i) * TT + (TT - 16 - X0));
                        o[i][0] = __byte_perm(v.w, 0, 0x0123); o[i][1] = __byte_perm(v.z, 0, 0x0123);
                        o[i][2] = __byte_perm(v.y, 0, 0x0123); o[i][3] = __byte_perm(v.x, 0, 0x0123);
                    }
                } else {
                    // k = 1: out(Y + i, X0 + q) = blk[X0 + q][T - 1 - Y - i]; k = 3: blk[T - 1 - X0 - q][Y + i]
                    const uint32_t* sw = reinterpret_cast<const uint32_t*>(sb);
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        uint32_t w[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint32_t q = 4 * m + j;
                            w[j] = k == 1 ? sw[((X0 + q) * TT + (TT - 4 - Y)) >> 2] : sw[((TT - 1 - X0 - q) * TT + Y) >> 2];
                        }
                        const uint32_t ab_lo = __byte_perm(w[0], w[1], 0x5140), ab_hi = __byte_perm(w[0], w[1], 0x7362);
                        const uint32_t cd_lo = __byte_perm(w[2], w[3], 0x5140), cd_hi = __byte_perm(w[2], w[3], 0x7362);
                        const uint32_t t0 = __byte_perm(ab_lo, cd_lo, 0x5410), t1 = __byte_perm(ab_lo, cd_lo, 0x7632);
                        const uint32_t t2 = __byte_perm(ab_hi, cd_hi, 0x5410), t3 = __byte_perm(ab_hi, cd_hi, 0x7632);
                        o[0][m] = k == 1 ? t3 : t0; o[1][m] = k == 1 ? t2 : t1;
                        o[2][m] = k == 1 ? t1 : t2; o[3][m] = k == 1 ? t0 : t3;
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) out[(((Y + i) * TT + X0) >> 4) * UM_ROWS + row] = make_uint4(o[i][0], o[i][1], o[i][2], o[i][3]);
        }
        for (uint32_t idx = N / 16 * UM_ROWS + threadIdx.x; idx < nch * UM_ROWS; idx += blockDim.x) out[idx] = make_uint4(0, 0, 0, 0);
    } else {
    for (uint32_t idx = threadIdx.x; idx < nch * UM_ROWS; idx += blockDim.x) {
        const uint32_t row = idx % UM_ROWS, ch = idx / UM_ROWS, lr = row >> 2, k = row & 3u;
        uint32_t w[4] = {0, 0, 0, 0};
        if (lr < nvalid && ch * 16 < N) {
            const uint8_t* sb = sblk + lr * NS;
            if (k == 0 && (N & 15u) == 0) {
                const uint4 v = *reinterpret_cast<const uint4*>(sb + ch * 16);
                w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
            } else if (k == 2 && (N & 15u) == 0) {             // element e of the rotated block = element N-1-e of the block
                const uint4 v = *reinterpret_cast<const uint4*>(sb + N - 16 - ch * 16);
                w[0] = __byte_perm(v.w, 0, 0x0123); w[1] = __byte_perm(v.z, 0, 0x0123);
                w[2] = __byte_perm(v.y, 0, 0x0123); w[3] = __byte_perm(v.x, 0, 0x0123);
            } else {
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const uint32_t e = ch * 16 + q;
                    if (e < N) {
                        const uint32_t Y = e / T, X = e - Y * T;
                        uint32_t py, px;
                        if (k == 0) { py = Y; px = X; }
                        else if (k == 1) { py = X; px = T - 1 - Y; }
                        else if (k == 2) { py = T - 1 - Y; px = T - 1 - X; }
                        else { py = T - 1 - X; px = Y; }
                        w[q >> 2] |= (uint32_t)sb[py * T + px] << (8 * (q & 3));
                    }
                }
            }
        }
        out[idx] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    }
    }
}

// B of the level: [tile][K stage][plane lo,hi][kc/16][64 cols][16 B].  One CTA per (tile, stage); thread -> (16-byte K
// chunk, column): writes the low-plane and the high-plane 16 bytes.  coln (by padded position): sum D^2, INT_MAX for
// padding columns.  colS2[sorted position] = sum D^2 of the domain (k_block_norms).
__global__ void k_build_pool_i8_level(const uint8_t* __restrict__ img, uint32_t stride, const fe_grid_item* __restrict__ dom,
                                      const uint32_t* __restrict__ order, const LevelPlan* __restrict__ plan, uint32_t T, uint32_t Kpad, uint32_t kc,
                                      const uint32_t* __restrict__ colS2, uint4* __restrict__ B8, uint32_t* __restrict__ coln_tiles,
                                      uint32_t* __restrict__ blob_dom, uint32_t* __restrict__ tileseg) {
    __shared__ uint32_t s_chunk, s_col0, s_end, s_doff;
    const uint32_t nst = Kpad / kc, ncs = kc / 16;
    const uint32_t tile = blockIdx.x / nst, st = blockIdx.x % nst;
    if (tile >= plan->n_tiles) return;
    if (threadIdx.x == 0) {
        const uint32_t nb = plan->nb;
        uint32_t lo = 0, hi = plan->nk * nb - 1;          // chunk with tile0[chunk] <= tile < tile0[chunk + 1]
        while (lo < hi) {
            const uint32_t mid = (lo + hi + 1) >> 1;
            if (plan->tile0[mid] <= tile) lo = mid; else hi = mid - 1;
        }
        const uint32_t k = lo / nb, b = lo - k * nb;
        s_chunk = lo;
        s_col0 = (k ? plan->dend[k - 1][b] : 0u) + (tile - plan->tile0[lo]) * I8_NT;
        s_end = plan->dend[k][b];
        s_doff = plan->doff[b];
    }
    __syncthreads();
    const uint32_t l = threadIdx.x % I8_NT, cs = threadIdx.x / I8_NT;      // blockDim = ncs * 64
    const uint32_t col = s_col0 + l;
    const uint32_t N = T * T;
    uint32_t lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0};
    const bool valid = col < s_end;
    const uint32_t c = s_doff + col;                                        // sorted position
    const uint32_t di = valid ? (order ? order[c] : c) : FE_NONE32;
    if (st == 0 && cs == 0) {
        blob_dom[(size_t)tile * I8_NT + l] = di;
        if (l == 0) tileseg[tile] = s_chunk;
    }
    const uint32_t e0 = st * kc + cs * 16;
    bool done = false;
    if (valid && (T & 15u) == 0 && e0 < N) {
        // 16 box sums = 16 consecutive values of one decimated row: two source rows of 32 pixels, as 16-byte loads when aligned
        const fe_grid_item d = dom[di];
        const uint32_t Y = e0 / T, X0 = e0 - Y * T;
        const uint8_t* p0 = img + (size_t)(d.y + 2 * Y) * stride + d.x + 2 * X0;
        if (((reinterpret_cast<uintptr_t>(p0) | stride) & 15u) == 0) {
            const uint4* r0 = reinterpret_cast<const uint4*>(p0);
            const uint4* r1 = reinterpret_cast<const uint4*>(p0 + stride);
            const uint4 ta = __ldg(r0), tb = __ldg(r0 + 1), ba = __ldg(r1), bb = __ldg(r1 + 1);
            const uint32_t top[8] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};
            const uint32_t bot[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
            uint32_t dd[8];                                   // two box sums per word, 16 bits each
#pragma unroll
            for (int i = 0; i < 8; ++i)
                dd[i] = (top[i] & 0x00FF00FFu) + ((top[i] >> 8) & 0x00FF00FFu) + (bot[i] & 0x00FF00FFu) + ((bot[i] >> 8) & 0x00FF00FFu);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                lo[i] = __byte_perm(dd[2 * i], dd[2 * i + 1], 0x6420);
                hi[i] = __byte_perm(dd[2 * i], dd[2 * i + 1], 0x7531);
            }
            done = true;
        }
    }
    if (valid && !done) {
        const fe_grid_item d = dom[di];
        const uint8_t* base = img + (size_t)d.y * stride + d.x;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const uint32_t e = st * kc + cs * 16 + q;
            if (e < N) {
                const uint8_t* p = base + (size_t)(2 * (e / T)) * stride + 2 * (e % T);
                const uint32_t D = (uint32_t)p[0] + p[1] + p[stride] + p[stride + 1];
                lo[q >> 2] |= (D & 255u) << (8 * (q & 3));
                hi[q >> 2] |= (D >> 8) << (8 * (q & 3));
            }
        }
    }
    if (st == 0 && cs == 0) // padding columns can never be a strict minimum nor pass the threshold
        coln_tiles[(size_t)tile * I8_NT + l] = valid ? colS2[c] : 0x7FFFFFFFu;
    // stage blob = [plane][kc/16][64][16 B]: in uint4 units plane stride = ncs * 64
    const size_t stage_base = ((size_t)tile * nst + st) * 2 * ncs * I8_NT;
    B8[stage_base + (size_t)cs * I8_NT + l] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    B8[stage_base + (size_t)ncs * I8_NT + (size_t)cs * I8_NT + l] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
}

} // namespace

// Per-block second moments, one warp per block (sorted position p -> item order[p]).
// mode 0: range, sum (4 r - 510)^2   mode 1: range, 16 sum r^2   mode 2: domain, sum (D - 510)^2   mode 3: domain, sum D^2
__global__ void k_block_norms(const uint8_t* __restrict__ img, uint32_t stride, const fe_grid_item* __restrict__ items,
                              const uint32_t* __restrict__ order, uint32_t n, uint32_t T, int mode, uint32_t* __restrict__ out) {
    const uint32_t p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (p >= n) return;
    const fe_grid_item it = items[order ? order[p] : p];
    const uint8_t* base = img + (size_t)it.y * stride + it.x;
    const uint32_t N = T * T;
    uint32_t s2 = 0;
    for (uint32_t e = lane; e < N; e += 32) {
        int v;
        if (mode < 2) {
            const int r = base[(size_t)(e / T) * stride + (e % T)];
            v = mode == 0 ? 4 * r - 510 : 4 * r;              // (4r)^2 = 16 r^2
        } else {
            const uint8_t* q = base + (size_t)(2 * (e / T)) * stride + 2 * (e % T);
            const int D = (int)q[0] + (int)q[1] + (int)q[stride] + (int)q[stride + 1];
            v = mode == 2 ? D - 510 : D;
        }
        s2 += (uint32_t)(v * v);
    }
    for (int o = 16; o; o >>= 1) s2 += __shfl_xor_sync(0xFFFFFFFFu, s2, o);
    if (lane == 0) out[p] = s2;
}

uint32_t i8_kpad(const LevelGeom& g) { return g.N <= (uint32_t)I8_KC ? ((g.N + 31u) & ~31u) : ((g.N + I8_KC - 1) / I8_KC) * I8_KC; }

int i8_level_supported(const LevelGeom& g) { return g.fast && g.T >= 4 && g.T <= 32; }

int i8_build_pool(fe_ctx* ctx, const LevelGeom& g, const DeviceLevel& lv, const uint32_t* dom_order, const LevelPlan* plan, uint32_t nD,
                  uint32_t max_tiles) {
    const fe_grid_item* d_dom = lv.d_dom;
    const uint32_t Kpad = i8_kpad(g), kc = std::min(Kpad, (uint32_t)I8_KC), nst = Kpad / kc, ncs = kc / 16;
    FE_CUDA(ctx, ctx->b_B16.ensure((size_t)max_tiles * 2 * I8_NT * Kpad + 256));
    FE_CUDA(ctx, ctx->b_coln.ensure((size_t)max_tiles * I8_NT * 4 + 64));
    FE_CUDA(ctx, ctx->b_blob_dom.ensure((size_t)max_tiles * I8_NT * 4 + 64));
    FE_CUDA(ctx, ctx->b_tileseg.ensure((size_t)max_tiles * 4 + 64));
    FE_CUDA(ctx, ctx->b_tmaps.ensure((size_t)nD * 4 + 64));
    if (lv.cellsD2) launch_dom_norms_from_cells(ctx->stream, lv.cellsD2, lv.cells_w, lv.dnx, dom_order, nD, ctx->b_tmaps.as<uint32_t>());
    else
        k_block_norms<<<(unsigned)(((uint64_t)nD * 32 + 255) / 256), 256, 0, ctx->stream>>>(ctx->src.px, ctx->src.stride, d_dom, dom_order, nD, g.T, 3,
                                                                                          ctx->b_tmaps.as<uint32_t>());
    k_build_pool_i8_level<<<max_tiles * nst, ncs * I8_NT, 0, ctx->stream>>>(ctx->src.px, ctx->src.stride, d_dom, dom_order, plan, g.T, Kpad, kc,
                                                                           ctx->b_tmaps.as<uint32_t>(), ctx->b_B16.as<uint4>(), ctx->b_coln.as<uint32_t>(),
                                                                           ctx->b_blob_dom.as<uint32_t>(), ctx->b_tileseg.as<uint32_t>());
    FE_CUDA(ctx, cudaGetLastError());
    ctx->stats.kernel_launches += 2;
    return FE_OK;
}

int i8_build_rows(fe_ctx* ctx, const LevelGeom& g, const LevelPlan* plan, const SliceCtl* ctl, const ListEntry* const list[2], uint32_t ordinal,
                  uint32_t max_row_tiles) {
    const uint32_t Kpad = i8_kpad(g);
    FE_CUDA(ctx, ctx->b_A16.ensure((size_t)max_row_tiles * UM_ROWS * Kpad + 256));
    const size_t smem = (size_t)32 * (g.N + 16);
    const uint32_t grid = std::min(max_row_tiles, 4u * (uint32_t)ctx->n_sm);   // grid-stride over the slice's row tiles
#define ROWS_I8(TT)                                                                                                                       \
    k_build_rows_i8<TT><<<grid, 256, smem, ctx->stream>>>(ctx->tgt.px, ctx->tgt.stride, plan, ctl, list[0], list[1], ordinal, g.T, Kpad, \
                                                          ctx->b_A16.as<uint4>())
    switch (g.T) {
    case 4: ROWS_I8(4); break;
    case 8: ROWS_I8(8); break;
    case 16: ROWS_I8(16); break;
    case 32: ROWS_I8(32); break;
    default: ROWS_I8(0); break;
    }
#undef ROWS_I8
    FE_CUDA(ctx, cudaGetLastError());
    ctx->stats.kernel_launches++;
    return FE_OK;
}

int i8_launch_search(fe_ctx* ctx, const LevelGeom& g, I8Args a, cudaEvent_t ev0, cudaEvent_t ev1) {
    const uint32_t Kpad = i8_kpad(g), kc = std::min(Kpad, (uint32_t)I8_KC);
    a.Kpad = Kpad;
    const uint32_t stage_bytes = 2 * I8_NT * kc, a_bytes = (a.pair ? 2u : 1u) * UM_ROWS * Kpad;
    const uint32_t budget = 226 * 1024 - 512;
    a.n_abuf = (2 * a_bytes + 2 * stage_bytes <= budget) ? 2 : 1;
    uint32_t stages = (budget - a.n_abuf * a_bytes) / stage_bytes;
    stages = std::min((uint32_t)I8_MAX_STAGES, stages);
    if (stages < 2) return fe_fail(ctx, FE_ERR_UNSUPPORTED, "i8 kind: operands do not fit shared memory (T=%u)", g.T);
    a.stages = stages;
    const size_t smem = (size_t)a.n_abuf * a_bytes + (size_t)stages * stage_bytes + (12 + 2 * I8_MAX_STAGES) * 8 + 128;
    FE_CUDA(ctx, cudaFuncSetAttribute(k_search_i8, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    if (ev0) cudaEventRecord(ev0, ctx->stream);
    k_search_i8<<<ctx->n_sm, UM_THREADS_I8, smem, ctx->stream>>>(a);
    FE_CUDA(ctx, cudaGetLastError());
    if (ev1) cudaEventRecord(ev1, ctx->stream);
    ctx->stats.kernel_launches++;
    return FE_OK;
}
