// fe_search_umma.cu -- the (range x rotation) x domain cross-correlation as a grouped dense
// contraction on tcgen05 tensor cores with TMEM accumulators, fused with the argmin epilogue.
// sm_100a only (tcgen05.mma / tcgen05.ld / cp.async.bulk / mbarrier, hand-written PTX).
//
// Formulation (fast geometry S = 2T):
//   a = 4 r - 510   (range pixel, inverse-rotated per row)        |a| <= 510
//   b = D   - 510   (domain 2x2 box sum)                          |b| <= 510
//   n16 = sum (a - b)^2 = sum a^2 + 2 V + p,   V = h - sum a b,   sum b^2 = 2 h + p,  p in {0,1}
// For a fixed row, argmin n16 == argmin (V, p, column).  The MMA computes V directly:
//   A row  = [ -a_0 .. -a_{N-1} | 1 | 2048 | 2048 | 0 .. ]          (fp16, all exact integers <= 2048)
//   B col  = [  b_0 ..  b_{N-1} | h0 | h1 | 2048*h2 | 0 .. ]        h = h0 + 2048 h1 + 2048^2 h2
// so the fp32 accumulator holds the INTEGER V.  Exactness: every product and every partial sum is an
// integer of magnitude < 2^24 whenever the final V is (|sum a b| <= N * 510^2 <= 16,646,400 for
// N <= 64), hence exact in fp32.  A winner with V >= 2^24 - 64 raises `inexact` and the host reruns
// the level on the exact integer kernel (only pathological images: best match worse than ~180 grey
// levels rms).  tests/test_gpu_umma.py holds the max-magnitude known-answer test.
//
// Kernel anatomy (one persistent CTA per SM, 608 threads; work item = (row tile, column chunk)):
//   warp 0      : TMEM allocator + bulk-async (TMA engine, cp.async.bulk) producer: A blob per work item, B blob per tile
//   warps 1-2   : one tcgen05.mma issuer thread per warpgroup; 4 accumulators x 128 TMEM columns; one tcgen05.commit per
//                 tile, which also tells the producer that the tile's B stage can be refilled
//   warps 3-18  : two compute warpgroups of 8 warps; thread = one row x 64 columns: tcgen05.ld -> accumulator released ->
//                 FMNMX3 tree; the column is only located (in registers, through groups of 8) when the row improves
// Operands are pre-laid-out in global memory in the no-swizzle K-major core-matrix order, one
// contiguous blob per tile, so a stage is ONE bulk copy and needs no tensor map.
#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>

#include "fe_umma_epi.cuh"

#ifndef FE_UMMA_PROF
#define FE_UMMA_PROF 0
#endif
#ifndef FE_USE_FMNMX3
#define FE_USE_FMNMX3 0
#endif

namespace {

using namespace umma_dev;

// RETIRE: retire quads of rows after their first threshold hit (extra per-thread state: worth it on the ALU-bound T=4 level)
// META: work items may run over several domain buckets (brightness bins): per-tile metadata (valid columns, bucket id,
// parity words) is read for every tile.  Otherwise a work item lies inside one bucket, only its last tile can be partial and
// nothing is loaded per tile.
template <int KIND, bool RETIRE, bool META>
__global__ void __launch_bounds__(UM_THREADS_F16, 1) k_search_umma(const UmmaArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bytesA = UM_ROWS * a.Kpad * 2, bytesB = UM_NT * a.Kpad * 2;
    const uint32_t S = 2 * UM_ISSUERS_F16;   // two B stages per issuer
    uint8_t* sA = smem;                     // 2 buffers
    uint8_t* sB = smem + 2 * bytesA;        // S stages
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)S * bytesB);
    const uint32_t bar0 = smem_u32(bars);
    auto A_FULL = [&](uint32_t i) { return bar0 + 8 * (0 + i); };
    auto A_EMPTY = [&](uint32_t i) { return bar0 + 8 * (2 + i); };
    auto ACC_FULL = [&](uint32_t g, uint32_t b) { return bar0 + 8 * (4 + 2 * g + b); };
    auto ACC_EMPTY = [&](uint32_t g, uint32_t b) { return bar0 + 8 * (8 + 2 * g + b); };
    auto B_FULL = [&](uint32_t i) { return bar0 + 8 * (12 + i); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12 + 2 * UM_MAX_STAGES);

    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < 2; ++i) {
            mbar_init(A_FULL(i), 1);
            mbar_init(A_EMPTY(i), UM_ISSUERS_F16);
        }
        for (uint32_t i = 0; i < 4; ++i) {
            mbar_init(bar0 + 8 * (4 + i), 1);  // ACC_FULL: one tcgen05.commit
            mbar_init(bar0 + 8 * (8 + i), 8);  // ACC_EMPTY: one arrive per compute warp of the warpgroup
        }
        for (uint32_t i = 0; i < S; ++i) mbar_init(B_FULL(i), 1);
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

    if (warp < UM_ISSUERS_F16) {
        // ================= issuers: one thread per accumulator buffer (g, ib), each also its own producer =================
        // Issuer (g, ib) owns accumulator (g, ib), the two B stages 2*(2g+ib), 2*(2g+ib)+1 and every tile whose global
        // counter is congruent to g + 2*ib modulo 4.  Per tile it (1) waits for the MMAs of its previous tile (its own
        // commit barrier) and refills that tile's stage with the bulk copy of the tile two steps ahead, (2) waits for this
        // tile's stage and for the compute warps to have drained the accumulator, (3) issues the MMAs and one commit.  The
        // four latency-bound loops overlap; there is no separate producer warp (20 warps = 5 per SM sub-partition keeps
        // the 96 registers the compute threads need).  Issuer (0,0) additionally keeps the two A buffers filled.
        if (lane == 0) {
            const uint32_t g = warp >> 1, ib = warp & 1, res = g + 2 * ib;
            const uint32_t sb = 2 * warp;                                  // first of this issuer's two stages
            const uint32_t idesc = (1u << 4) | ((uint32_t)(UM_NT >> 3) << 17) | ((uint32_t)(UM_ROWS >> 4) << 24);
            const uint32_t nk = (a.dbg & 2) ? 0u : a.Kpad / 16;
            struct Cursor {
                uint32_t w, wi, it0, u, n;
                WorkItem item;
                bool valid;
            };
            uint32_t memo_w = FE_NONE32;     // the two cursors visit the same items a tile or two apart: decode each once
            WorkItem memo{};
            auto load_item = [&](Cursor& c) {
                c.valid = c.w < a.total_items;
                if (c.valid) {
                    if (c.w != memo_w) { memo = decode_item(a, c.w); memo_w = c.w; }
                    c.item = memo; c.n = c.item.t1 - c.item.t0;
                }
            };
            Cursor ld, mm;
            ld.w = mm.w = blockIdx.x; ld.wi = mm.wi = 0; ld.it0 = mm.it0 = 0;
            load_item(ld);
            mm = ld;
            // ---- A tiles (issuer 0 only): A(wi) lives in buffer wi & 1 ----
            // The refill of a buffer has to wait until all four issuers' MMAs of the item two back are complete; issuer 0
            // must not sit in that wait (its accumulator would idle once per item -- with short items, e.g. a slice of
            // three brightness bins, that is most of the time), so the refill is posted and polled from the tile loop.
            uint32_t pend_A = FE_NONE32;
            auto load_A = [&](bool block) {
                if (pend_A == FE_NONE32) return;
                const uint32_t wl = blockIdx.x + pend_A * gridDim.x;
                if (wl >= a.total_items) { pend_A = FE_NONE32; return; }
                const uint32_t ab = pend_A & 1;
                if (pend_A >= 2) {                                              // all four issuers are done with A(pend_A - 2)
                    const uint32_t par = ((pend_A >> 1) & 1) ^ 1;
                    if (block) mbar_wait(A_EMPTY(ab), par);
                    else if (!mbar_try_wait(A_EMPTY(ab), par)) return;
                }
                const WorkItem it = decode_item(a, wl);
                mbar_expect_tx(A_FULL(ab), bytesA);
                bulk_g2s(smem_u32(sA + ab * bytesA), reinterpret_cast<const uint8_t*>(a.A16) + (size_t)it.a_blob * bytesA, bytesA, A_FULL(ab));
                pend_A = FE_NONE32;
            };
            if (warp == 0) { pend_A = 0; load_A(true); pend_A = 1; load_A(true); }
            // ---- cursor movement; the MMA cursor signs off every item it leaves (A_EMPTY needs all four issuers) ----
            bool mm_had_tiles = false;
            auto next_item = [&](Cursor& c, bool is_mm) {
                if (is_mm) {
                    // An issuer without a tile in this item (fewer than four column tiles) must not sign it off before the
                    // item's A tile has even been loaded: it could run two items ahead and arrive twice on the same
                    // A_EMPTY phase, releasing a buffer whose MMAs are still in flight.  Waiting for A_FULL of the item
                    // orders it behind the completion of the item two back, like the issuers that do have tiles.
                    if (!mm_had_tiles) mbar_wait(A_FULL(c.wi & 1), (c.wi >> 1) & 1);
                    if (mm_had_tiles) tc_commit(A_EMPTY(c.wi & 1)); else mbar_arrive(A_EMPTY(c.wi & 1));
                    mm_had_tiles = false;
                    if (warp == 0) { load_A(true); pend_A = c.wi + 2; }         // (an older refill still posted: finish it first)
                }
                c.it0 += c.n; c.w += gridDim.x; ++c.wi;
                load_item(c);
            };
            auto seek = [&](Cursor& c, bool is_mm) {
                while (c.valid) {
                    c.u = (res + 4 - (c.it0 & 3)) & 3;
                    if (c.u < c.n) return;
                    next_item(c, is_mm);
                }
            };
            auto advance = [&](Cursor& c, bool is_mm) {
                c.u += 4;
                if (c.u >= c.n) { next_item(c, is_mm); seek(c, is_mm); }
            };
            seek(ld, false);
            seek(mm, true);
            uint32_t m_ld = 0, m = 0;      // tiles loaded / issued by this issuer so far
            auto load_B = [&]() {           // bulk copy of the load cursor's tile into stage sb + (m_ld & 1)
                const uint32_t s = sb + (m_ld & 1);
                if (m_ld >= 2) mbar_wait(ACC_FULL(g, ib), (m_ld - 2) & 1);   // the stage's previous tile: its MMAs have completed
                if (a.dbg & 4) mbar_arrive(B_FULL(s));
                else {
                    mbar_expect_tx(B_FULL(s), bytesB);
                    bulk_g2s(smem_u32(sB + (size_t)s * bytesB), reinterpret_cast<const uint8_t*>(a.B16) + (size_t)(ld.item.t0 + ld.u) * bytesB, bytesB, B_FULL(s));
                }
                ++m_ld;
                advance(ld, false);
            };
            if (ld.valid) load_B();
            if (ld.valid) load_B();
            uint32_t cur_wi = 0xFFFFFFFFu, a_addr = 0;
            while (mm.valid) {
                // refill: tile m+1 goes into the stage of tile m-1, whose MMAs were issued one iteration ago
                if (m >= 1 && ld.valid) load_B();
                if (warp == 0) load_A(false);
                if (mm.wi != cur_wi) {       // first tile of this issuer in a new item: its A tile must have landed
                    cur_wi = mm.wi;
                    if (warp == 0 && pend_A == cur_wi) load_A(true);
                    mbar_wait(A_FULL(cur_wi & 1), (cur_wi >> 1) & 1);
                    a_addr = smem_u32(sA + (cur_wi & 1) * bytesA);
                }
                const uint32_t s = sb + (m & 1);
                mbar_wait(B_FULL(s), (m >> 1) & 1);
                mbar_wait(ACC_EMPTY(g, ib), (m & 1) ^ 1);
                tc_fence_after();
                const uint32_t b_addr = smem_u32(sB + (size_t)s * bytesB);
                const uint32_t d_tmem = tmem_base + (g * 2 + ib) * UM_NT;
                for (uint32_t kk = 0; kk < nk; ++kk) {
                    // chunk-major blobs: K chunk c (8 halves = 16 bytes) of all rows is contiguous
                    const uint64_t adesc = make_desc(a_addr + kk * 2 * (UM_ROWS * 16), UM_ROWS * 16, 128);
                    const uint64_t bdesc = make_desc(b_addr + kk * 2 * (UM_NT * 16), UM_NT * 16, 128);
                    tc_mma<KIND>(d_tmem, adesc, bdesc, idesc, kk > 0 ? 1u : 0u);
                }
                tc_commit(ACC_FULL(g, ib));
                mm_had_tiles = true;
                ++m;
                advance(mm, true);
            }
        }
    } else {
        // ================= compute warps: TMEM -> registers -> row argmin =================
        // Warpgroup g = 8 warps: warp (sp, h) reads TMEM lanes 32*sp..+31 (the sub-partition its warp id maps
        // to) and columns h*64..+63 of the group's accumulators, i.e. every thread owns one row x 64 columns of a
        // tile.  Four compute warps per SM sub-partition hide the TMEM-load and FMNMX3 latencies of each other;
        // an accumulator goes back to the issuer as soon as all eight warps hold their part in registers.
        const uint32_t cw = warp - UM_ISSUERS_F16;
        const uint32_t g = cw >> 3;
        const uint32_t h = (cw >> 2) & 1;             // column half
        const uint32_t sp = warp & 3;                 // TMEM sub-partition this warp may read
        const uint32_t lrow = sp * 32 + lane;         // row inside the tile == TMEM lane
        const uint32_t lane_addr = tmem_base + ((sp * 32u) << 16) + h * UM_HALF;
        uint32_t it0 = 0, jb = 0;
        for (uint32_t w = blockIdx.x; w < a.total_items; w += gridDim.x) {
            const WorkItem item = decode_item(a, w);
            const uint32_t n = item.t1 - item.t0;
            const bool row_ok = lrow < item.nrows;
            const uint32_t grow = item.row0 + lrow;
            const uint32_t a2 = row_ok ? a.rowA2[grow >> 2] : 0u;
            const uint32_t srow = (row_ok && a.rowslot) ? 4u * a.rowslot[grow >> 2] + (grow & 3u) : grow;   // result slot of the level
            RowState st;
            st.bestV = 3.0e38f; st.bestp = 0; st.bestcol = FE_NONE32; st.hit = FE_NONE32;
            st.vthr0 = -3.0e38f; st.vthr1 = -3.0e38f;
            st.par_item = reinterpret_cast<const uint32_t*>(a.colmeta + (size_t)item.t0 * 2);
            uint32_t cur_seg = FE_NONE32;       // domain bucket of the tiles this thread is scanning
            if (a.use_thr && row_ok) { // n16 <= thr16  <=>  V <= floor((thr16 - a2 - p) / 2)
                const long long tt = (long long)a.thr16 - (long long)a2;
                long long f0 = tt >= 0 ? tt / 2 : -((-tt + 1) / 2);
                long long f1 = (tt - 1) >= 0 ? (tt - 1) / 2 : -((-(tt - 1) + 1) / 2);
                f0 = max(-16777216ll, min(16777215ll, f0));
                f1 = max(-16777216ll, min(16777215ll, f1));
                st.vthr0 = (float)f0;
                st.vthr1 = (float)f1;
            }
            if (row_ok && !a.no_min) {
                // Seed the running minimum with what earlier passes / column chunks already found for this row, plus one:
                // anything this item finds at or below the recorded score still gets written (ties are settled by the
                // column index inside the 64-bit key), everything above it is rejected by the cheap tile-minimum test.
                const unsigned long long seen = __ldcg(&a.rowbest[srow]);
                if (seen != FE_INF64) {
                    const long long d = (long long)(uint32_t)(seen >> 32) + 1ll - (long long)a2;   // = 2 V + p
                    const long long V = d >= 0 ? d / 2 : -((-d + 1) / 2);
                    if (V > -16777216ll && V < 16777216ll) { st.bestV = (float)V; st.bestp = (uint32_t)(d - 2 * V); }
                }
            }
            const uint32_t first = (g + UM_WGS - (it0 % UM_WGS)) % UM_WGS;
            const uint32_t my_tiles = first < n ? (n - first + UM_WGS - 1) / UM_WGS : 0;
            // Threshold runs: a range is decided by its first hit in scan order, so once any of its four rotation rows
            // (four adjacent lanes) has crossed the threshold the whole quad is retired for the rest of the item, and a
            // warp whose 32 rows are all retired only keeps the accumulator hand-shake going.
            bool retired = !row_ok;
            bool warp_done = false;
            for (uint32_t j = 0; j < my_tiles; ++j, ++jb) {
                const uint32_t u = first + j * UM_WGS, buf = jb & 1;
                const uint32_t colbase = u * UM_NT + h * UM_HALF;
                // metadata of this half tile (META: one 16-byte broadcast load per warp, issued ahead of the accumulator wait)
                static_assert(UM_HALF == 64, "two parity words per half tile");
                const uint4* meta_p = a.colmeta + (size_t)(item.t0 + u) * 2 + h;
                uint4 meta = make_uint4(0, 0, 0, 0);
                ParitySrc par;
                par.ptr = reinterpret_cast<const uint32_t*>(meta_p);
                uint32_t nvalid;
                if (META) {
                    meta = __ldg(meta_p);
                    par.reg[0] = meta.x; par.reg[1] = meta.y;
                    nvalid = meta.z;
                } else {
                    const uint32_t tile_valid = min((uint32_t)UM_NT, item.cols_left - u * UM_NT);
                    nvalid = tile_valid > h * UM_HALF ? min((uint32_t)UM_HALF, tile_valid - h * UM_HALF) : 0u;
                }
                // The item moves into another domain bucket: columns restart at low domain indices there, so the first hit
                // found so far is only the first of the bucket behind us -- bank it and start over.  (Looked at after the
                // accumulator has been read, when the metadata load has long landed; the retiring kernel needs it earlier.)
                auto bucket_change = [&]() {
                    if (META && meta.w != cur_seg) {
                        if (st.hit != FE_NONE32) {
                            atomicMin(&a.rowhit[srow], a.blob_dom[(size_t)item.t0 * UM_NT + st.hit]);
                            st.hit = FE_NONE32;
                        }
                        cur_seg = meta.w;
                        if (RETIRE) { retired = !row_ok; warp_done = false; }
                    }
                };
                if (RETIRE) bucket_change();
                uint32_t v[UM_HALF];
                mbar_wait(ACC_FULL(g, buf), (jb >> 1) & 1);
                tc_fence_after();
                if (!(a.dbg & 1) && !(RETIRE && warp_done)) {
                    const uint32_t taddr = lane_addr + (g * 2 + buf) * UM_NT;
                    TMEM_LD32(taddr, (v + 0));
                    TMEM_LD32(taddr + 32, (v + 32));
                    tmem_wait_ld();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(ACC_EMPTY(g, buf));
                if (!RETIRE) bucket_change();
                if (!(a.dbg & 1) && !(RETIRE && warp_done)) {
                    process_half<META>(v, st, RETIRE ? !retired : row_ok, colbase, nvalid, par, a.no_min == 0);
                    if (RETIRE && (j & 3) == 3) {   // every 4th tile is enough: retirement only saves work
                        uint32_t hm = __ballot_sync(0xFFFFFFFFu, st.hit != FE_NONE32);
                        hm = (hm | (hm >> 1) | (hm >> 2) | (hm >> 3)) & 0x11111111u;   // one bit per quad of lanes
                        retired = retired || (((hm * 15u) >> lane) & 1u);
                        warp_done = __all_sync(0xFFFFFFFFu, retired);
                    }
                }
            }
            it0 += n;
            if (row_ok) {
                if (st.bestcol != FE_NONE32) {
                    if (st.bestp == 2) st.bestp = row_parity(st, st.bestcol);
                    const long long n16 = (long long)a2 + 2ll * (long long)st.bestV + (long long)st.bestp;
                    const unsigned long long key = ((unsigned long long)(uint32_t)n16 << 32) | (unsigned long long)a.blob_dom[(size_t)item.t0 * UM_NT + st.bestcol];
                    atomicMin(&a.rowbest[srow], key);
                    if (st.bestV >= 16777216.0f - 64.0f) atomicOr(a.flags, 1u);
                }
                if (st.hit != FE_NONE32) atomicMin(&a.rowhit[srow], a.blob_dom[(size_t)item.t0 * UM_NT + st.hit]);
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

} // namespace

// ---------------------------------------------------------------------------------------------------
// operand blobs
// ---------------------------------------------------------------------------------------------------
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

// Both builders: one thread per block (range / domain).  The thread reads its block with T-byte vector loads (byte loads
// when the origin is not T-aligned), forms every K chunk (8 halves = 16 bytes) of its rows / column in registers and stores
// them; consecutive threads own consecutive rows / columns, so each store instruction of a warp covers one contiguous run
// of the blob.  The block norm falls out of the same registers (no separate norm pass) and every byte of every blob
// (padding rows, columns and K included) is written exactly once -- no memset, write traffic = blob size.
// A rows.  Row (4*lr + k) of row tile `tile`: range lr of the tile under the inverse of rotation k, value 510 - 4 r, then
// the constant columns [1, 2048, 2048].  rowA2[j] = sum (4r - 510)^2.
template <int T>
__global__ void __launch_bounds__(128) k_build_rows16(const uint8_t* __restrict__ img, uint32_t stride, const fe_grid_item* __restrict__ rng,
                                                      const uint32_t* __restrict__ order, UmmaBuckets bk, uint4* __restrict__ A16,
                                                      uint32_t* __restrict__ rowA2) {
    constexpr int N = T * T, KPAD = (N + 3 + 15) & ~15, NCH = KPAD / 8, W = T / 4;
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= bk.row_tile0[bk.nb] * 32u) return;
    const uint32_t tile = idx / 32, lr = idx % 32;
    int bi = 0;
    while (bi + 1 < bk.nb && tile >= bk.row_tile0[bi + 1]) ++bi;
    const uint32_t j = bk.range_off[bi] + (tile - bk.row_tile0[bi]) * 32 + lr;
    uint4* out = A16 + (size_t)tile * NCH * UM_ROWS + 4 * lr;
    if (j >= bk.range_off[bi + 1]) {
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
            for (int k = 0; k < 4; ++k) out[ch * UM_ROWS + k] = make_uint4(0, 0, 0, 0);
        return;
    }
    const fe_grid_item r = rng[order ? order[j] : j];
    const uint8_t* base = img + (size_t)r.y * stride + r.x;
    uint32_t w[T][W];
#pragma unroll
    for (int y = 0; y < T; ++y) load_px<T>(base + (size_t)y * stride, w[y]);
    auto px = [&](int py, int pxx) { return (int)((w[py][pxx >> 2] >> (8 * (pxx & 3))) & 255u); };
    uint32_t s2 = 0;
#pragma unroll
    for (int y = 0; y < T; ++y)
#pragma unroll
        for (int x = 0; x < T; ++x) { const int a = 4 * px(y, x) - 510; s2 += (uint32_t)(a * a); }
    rowA2[j] = s2;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int e = ch * 8 + q;
                if (e < N) {
                    const int Y = e / T, X = e % T;
                    const int py = k == 0 ? Y : k == 1 ? X : k == 2 ? T - 1 - Y : T - 1 - X;
                    const int pxx = k == 0 ? X : k == 1 ? T - 1 - Y : k == 2 ? T - 1 - X : Y;
                    v[q] = (float)(510 - 4 * px(py, pxx));
                } else {
                    v[q] = e == N ? 1.0f : (e == N + 1 || e == N + 2) ? 2048.0f : 0.0f;
                }
            }
            out[ch * UM_ROWS + k] = make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]), pack_half2(v[6], v[7]));
        }
    }
}

// B columns.  b = D - 510; limbs of h = floor(sum b^2 / 2) in the three columns after the data; parity bit per column
// (one ballot word per 32 columns).
template <int T>
__global__ void __launch_bounds__(128) k_build_pool16(const uint8_t* __restrict__ img, uint32_t stride, const fe_grid_item* __restrict__ dom,
                                                      const uint32_t* __restrict__ order, UmmaBuckets bk, uint4* __restrict__ B16,
                                                      uint32_t* __restrict__ colmeta, uint32_t* __restrict__ blob_dom) {
    constexpr int N = T * T, KPAD = (N + 3 + 15) & ~15, NCH = KPAD / 8, W = T / 4;
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= bk.col_tile0[bk.nb] * (uint32_t)UM_NT) return;          // whole warps: UM_NT is a multiple of 32
    const uint32_t tile = idx / UM_NT, l = idx % UM_NT;
    int bi = 0;
    while (bi + 1 < bk.nb && tile >= bk.col_tile0[bi + 1]) ++bi;
    const uint32_t c = bk.dom_off[bi] + (tile - bk.col_tile0[bi]) * UM_NT + l;
    uint4* out = B16 + (size_t)tile * NCH * UM_NT + l;
    const bool live = c < bk.dom_end[bi];
    uint32_t s2 = 0;
    const uint32_t di = live ? (order ? order[c] : c) : FE_NONE32;
    blob_dom[idx] = di;
    if (live) {
        const fe_grid_item d = dom[di];
        const uint8_t* base = img + (size_t)d.y * stride + d.x;
        float v[8];
#pragma unroll
        for (int Y = 0; Y < T; ++Y) {
            uint32_t top[2 * W], bot[2 * W];
            {
                uint32_t lo[W], hi[W];
                load_px<T>(base + (size_t)(2 * Y) * stride, lo);
                load_px<T>(base + (size_t)(2 * Y) * stride + T, hi);
#pragma unroll
                for (int i = 0; i < W; ++i) { top[i] = lo[i]; top[W + i] = hi[i]; }
                load_px<T>(base + (size_t)(2 * Y + 1) * stride, lo);
                load_px<T>(base + (size_t)(2 * Y + 1) * stride + T, hi);
#pragma unroll
                for (int i = 0; i < W; ++i) { bot[i] = lo[i]; bot[W + i] = hi[i]; }
            }
#pragma unroll
            for (int i = 0; i < 2 * W; ++i) {
                // two box sums per word pair: horizontal byte pairs in 16-bit lanes, then the two rows
                const uint32_t t = (top[i] & 0x00FF00FFu) + ((top[i] >> 8) & 0x00FF00FFu);
                const uint32_t b = (bot[i] & 0x00FF00FFu) + ((bot[i] >> 8) & 0x00FF00FFu);
                const uint32_t dd = t + b;
                const int d0 = (int)(dd & 0xFFFFu) - 510, d1 = (int)(dd >> 16) - 510;
                s2 += (uint32_t)(d0 * d0) + (uint32_t)(d1 * d1);
                v[(Y * T + 2 * i) & 7] = (float)d0;
                v[(Y * T + 2 * i + 1) & 7] = (float)d1;
            }
            if (((Y + 1) * T) % 8 == 0)
                out[(((Y + 1) * T) / 8 - 1) * UM_NT] = make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]), pack_half2(v[6], v[7]));
        }
        const uint32_t h = s2 >> 1;
        out[(N / 8) * UM_NT] = make_uint4(pack_half2((float)(h & 2047u), (float)((h >> 11) & 2047u)), pack_half2((float)((h >> 22) * 2048u), 0.f), 0, 0);
#pragma unroll
        for (int ch = N / 8 + 1; ch < NCH; ++ch) out[ch * UM_NT] = make_uint4(0, 0, 0, 0);
    } else {
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) out[ch * UM_NT] = make_uint4(0, 0, 0, 0);
    }
    // tile metadata, 4 words per half tile: parity of columns 0-31, 32-63, valid columns of the half, domain bucket
    const uint32_t par = __ballot_sync(0xFFFFFFFFu, live && (s2 & 1u));
    if ((l & 31) == 0) {
        uint32_t* m = colmeta + ((size_t)tile * 2 + (l >> 6)) * 4;
        m[(l >> 5) & 1u] = par;
        if ((l & 63) == 0) {
            const uint32_t end = bk.dom_end[bi];
            m[2] = c >= end ? 0u : min(end - c, (uint32_t)UM_HALF);
            m[3] = (uint32_t)bi;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
int umma_level_supported(const LevelGeom& g) { return g.fast && (g.T == 4 || g.T == 8); }

uint32_t umma_kpad(const LevelGeom& g) { return (g.N + 3 + 15u) & ~15u; }

int umma_prepare_and_search(fe_ctx* ctx, const LevelGeom& g, const fe_grid_item* d_dom, const fe_grid_item* d_rng, const SearchPass& sp,
                            uint32_t thr16, bool use_thr) {
    const uint32_t Kpad = umma_kpad(g);
    const int nbuckets = sp.nbuckets;
    const uint32_t* dom_order = sp.dom_order;
    const uint32_t* rng_order = sp.rng_items;
    UmmaBuckets bk{};
    UmmaArgs a{};
    uint32_t rt = 0, ct = 0, nb = 0;
    uint64_t total_items = 0;
    // B blob: the slices of the domain buckets, one after the other (whole column tiles each)
    uint32_t nct[FE_MAX_BUCKETS];
    for (int b = 0; b < nbuckets; ++b) {
        const uint32_t dc = sp.dend[b] - sp.dbeg[b];
        bk.dom_off[b] = sp.dbeg[b];
        bk.dom_end[b] = sp.dend[b];
        bk.col_tile0[b] = ct;
        nct[b] = (dc + UM_NT - 1) / UM_NT;
        ct += nct[b];
    }
    bk.col_tile0[nbuckets] = ct;
    // A blob: the range buckets; bucket c meets the column tiles of the domain buckets c-span .. c+span (adjacent in the blob)
    for (int c = 0; c < nbuckets; ++c) {
        const uint32_t rc = sp.roff[c + 1] - sp.roff[c];
        bk.range_off[nb] = sp.roff[c];
        bk.row_tile0[nb] = rt;
        UmmaBucket& b = a.b[nb];
        b.row_tile0 = rt; b.n_row_tiles = (rc + 31) / 32;
        const int b0 = std::max(0, c - sp.span), b1 = std::min(nbuckets - 1, c + sp.span);
        b.col_tile0 = bk.col_tile0[b0];
        b.n_col_tiles = bk.col_tile0[b1 + 1] - bk.col_tile0[b0];
        b.row0 = sp.roff[c] * 4; b.nrows = rc * 4;
        b.ncols = sp.dend[b0] - sp.dbeg[b0];
        rt += b.n_row_tiles;
        ++nb;
    }
    bk.nb = (int)nb;
    bk.range_off[nb] = sp.roff[nbuckets];
    bk.row_tile0[nb] = rt;
    bk.n_ranges = sp.roff[nbuckets];
    bk.n_domains = sp.n_dom;
    // column chunking: aim at >= 2 work items per SM when there are few row tiles
    uint32_t live_row_tiles = 0;
    for (uint32_t i = 0; i < nb; ++i)
        if (a.b[i].n_col_tiles) live_row_tiles += a.b[i].n_row_tiles;
    const uint32_t want_chunks = live_row_tiles ? (2 * 148 + live_row_tiles - 1) / live_row_tiles : 1;
    for (uint32_t i = 0; i < nb; ++i) {
        UmmaBucket& b = a.b[i];
        b.chunks = b.n_col_tiles ? std::max(1u, std::min(want_chunks, b.n_col_tiles)) : 0;
        if (!b.n_row_tiles) b.chunks = 0;
        total_items += (uint64_t)b.n_row_tiles * b.chunks;
        a.item_end[i] = (uint32_t)std::min<uint64_t>(total_items, 0xFFFFFFFFull);
    }
    if (total_items == 0) {
        if (sp.ev0) { cudaEventRecord(sp.ev0, ctx->stream); cudaEventRecord(sp.ev1, ctx->stream); }
        return FE_OK;
    }
    if (total_items > 0x7FFFFFFFull) return fe_fail(ctx, FE_ERR_UNSUPPORTED, "umma: too many work items");

    const size_t bytesA = (size_t)rt * UM_ROWS * Kpad * 2, bytesB = (size_t)ct * UM_NT * Kpad * 2;
    FE_CUDA(ctx, ctx->b_A16.ensure(bytesA + 256));
    FE_CUDA(ctx, ctx->b_B16.ensure(bytesB + 256));
    FE_CUDA(ctx, ctx->b_tmaps.ensure((size_t)ct * 32 + 64));
    FE_CUDA(ctx, ctx->b_blob_dom.ensure((size_t)ct * UM_NT * 4 + 64));
    FE_CUDA(ctx, ctx->b_rowc.ensure((size_t)bk.n_ranges * 4 + 4));
    uint32_t* flags = ctx->b_counters.as<uint32_t>() + 2;
    {
        const unsigned gr = (rt * 32 + 127) / 128, gc = (ct * UM_NT + 127) / 128;
        uint4* A16 = ctx->b_A16.as<uint4>();
        uint4* B16 = ctx->b_B16.as<uint4>();
        uint32_t* rowA2 = ctx->b_rowc.as<uint32_t>();
        uint32_t* colmeta = ctx->b_tmaps.as<uint32_t>();
        uint32_t* blob_dom = ctx->b_blob_dom.as<uint32_t>();
        if (g.T == 4) {
            if (!sp.reuse_rows) k_build_rows16<4><<<gr, 128, 0, ctx->stream>>>(ctx->tgt.px, ctx->tgt.stride, d_rng, rng_order, bk, A16, rowA2);
            k_build_pool16<4><<<gc, 128, 0, ctx->stream>>>(ctx->src.px, ctx->src.stride, d_dom, dom_order, bk, B16, colmeta, blob_dom);
        } else {
            if (!sp.reuse_rows) k_build_rows16<8><<<gr, 128, 0, ctx->stream>>>(ctx->tgt.px, ctx->tgt.stride, d_rng, rng_order, bk, A16, rowA2);
            k_build_pool16<8><<<gc, 128, 0, ctx->stream>>>(ctx->src.px, ctx->src.stride, d_dom, dom_order, bk, B16, colmeta, blob_dom);
        }
        FE_CUDA(ctx, cudaGetLastError());
    }
    ctx->stats.kernel_launches += sp.reuse_rows ? 1 : 2;

    a.A16 = ctx->b_A16.p;
    a.B16 = ctx->b_B16.p;
    a.colmeta = ctx->b_tmaps.as<uint4>();
    a.blob_dom = ctx->b_blob_dom.as<uint32_t>();
    a.rowA2 = ctx->b_rowc.as<uint32_t>();
    a.rowbest = ctx->b_rowbest.as<unsigned long long>();
    a.rowhit = ctx->b_rowhit.as<uint32_t>();
    a.flags = flags;
    a.nb = (int)nb;
    a.Kpad = Kpad;
    a.total_items = (uint32_t)total_items;
    a.thr16 = thr16;
    a.use_thr = use_thr ? 1u : 0u;
    a.nt = UM_NT;
    a.rowslot = sp.rowslot;
    a.meta = sp.span > 0 ? 1u : 0u;
    a.no_min = sp.no_min ? 1u : 0u;
    { const char* e = getenv("FE_UMMA_DBG"); a.dbg = e ? (uint32_t)atoi(e) : 0u; }
    const uint32_t stage_bytes = UM_NT * Kpad * 2, a_bytes = 2 * UM_ROWS * Kpad * 2;
    const uint32_t stages = 2 * UM_ISSUERS_F16;
    a.stages = stages;
    if (Kpad / 16 > UM_MAX_NK) return fe_fail(ctx, FE_ERR_UNSUPPORTED, "umma: K too large");
    const size_t smem = (size_t)a_bytes + (size_t)stages * stage_bytes + (12 + 2 * UM_MAX_STAGES) * 8 + 64;
    const bool retire = use_thr && g.T == 4;
    const bool meta = sp.span > 0;
    auto kern = retire ? (meta ? k_search_umma<0, true, true> : k_search_umma<0, true, false>)
                       : (meta ? k_search_umma<0, false, true> : k_search_umma<0, false, false>);
    FE_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    const uint32_t grid = (uint32_t)std::min<uint64_t>(total_items, 148);
    if (sp.ev0) cudaEventRecord(sp.ev0, ctx->stream);
    kern<<<grid, UM_THREADS_F16, smem, ctx->stream>>>(a);
    FE_CUDA(ctx, cudaGetLastError());
    if (sp.ev1) cudaEventRecord(sp.ev1, ctx->stream);
    ctx->stats.kernel_launches++;
    if (a.dbg & 32) {
        unsigned long long c[6];
        cudaStreamSynchronize(ctx->stream);
        cudaMemcpy(c, flags + 2, sizeof(c), cudaMemcpyDeviceToHost);
        const double nt = (double)std::max<unsigned long long>(1, c[5]);
        fprintf(stderr, "[umma prof] T=%u per warp-tile cycles: wait_ldA %.0f  procA %.0f  wait_ldB %.0f  release+wait_full+issue %.0f  procB %.0f  (tiles %.0f)\n",
                g.T, c[0] / nt, c[1] / nt, c[2] / nt, c[3] / nt, c[4] / nt, nt);
    }
    return FE_OK;
}
