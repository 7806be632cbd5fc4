// fe_search_f16.cu -- tcgen05 kind::f16 search of the small range blocks (T = 4, 8), scheduled from device memory.
//
// Contraction: operands centred at 510 (a = 4 r - 510, b = D - 510); with sum b^2 = 2 h + p the score is
// n16 = sum a^2 + 2 V + p, V = h - sum a b.  The MMA computes V itself: A row = [-a | 1 | 2048 | 2048 | 0...], B column =
// [b | h0 | h1 | 2048 h2 | 0...] (limbs of h; every entry an integer <= 2048, exact in fp16).  Every product and partial sum is
// an integer below 2^24 whenever the final V is, so the fp32 TMEM accumulator is exact; a winner with V >= 2^24 - 64 raises a
// flag and the level is redone on the i8 kind.  Epilogue: FMNMX3 row argmin (fe_umma_epi.cuh).
//
//  * work items come from a record list in device memory (fe_plan.cu: k_expand_items) and the slice's parameters from
//    SliceCtl, so the host launches the kernel without knowing what the previous slice left -- a launch whose ordinal was
//    not planned returns at once;
//  * the A tile (32 range blocks x 4 rotations, fp16, K-major core-matrix order) is built INSIDE the kernel from the u8
//    pixels of the range blocks by four builder warps, one lane per (range block, rotation): 64 bytes of pixels in, 640
//    bytes of operand rows out, never through global memory;
//  * the B operand is one blob per LEVEL (k_build_pool16_level), interval-major, so every slice addresses tile runs of it.
//
// Kernel anatomy (one persistent CTA per SM, 768 threads, register budget re-split with setmaxnreg):
//   warps 0-3   issuers (one thread each): B tile bulk copies two steps ahead, tcgen05.mma M128 x N128 x K16, one commit
//               per tile; accumulator t % 4 of the 4 x 128 TMEM columns (PAIR: two issuers, see below)   (40 registers)
//   warps 4-7   A builders: 8 range blocks per warp, lane = (block, rotation)                              (56 registers)
//   warps 8-23  two compute warpgroups of 8 warps: tcgen05.ld 32x32b.x32 x 2 -> release -> FMNMX3 argmin   (96 registers)
#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>

#include "fe_plan.cuh"
#include "fe_umma_epi.cuh"

namespace {

using namespace umma_dev;

constexpr int F16_THREADS = 768;
constexpr int F16_STAGES = 8;              // two B stages per issuer
constexpr int F16_ABUFS = 3;               // A tiles in shared memory (T = 8: 3 x 20 KB + 8 x 20 KB of B stages = 220 KB)
constexpr int F16_ISSUERS = 4;
constexpr int F16_BUILDERS = 4;

// Rows 4 r + k (k = 0..3) of the A tile: range block r of the item under the inverse of rotation k, values 510 - 4 p, then
// the constant columns [1, 2048, 2048].  Lane = (r, k); the four lanes of a block read the same 64 pixels.
template <int T>
__device__ __forceinline__ void build_rows(uint8_t* sAbuf, const uint8_t* __restrict__ img, uint32_t stride, uint32_t xy, bool mirror, bool valid,
                                           uint32_t bw, uint32_t lane) {
    constexpr int N = T * T, KPAD = (N + 3 + 15) & ~15, NCH = KPAD / 8, W = T / 4;
    const uint32_t r = bw * 8 + (lane >> 2), k = lane & 3u;
    uint4* out = reinterpret_cast<uint4*>(sAbuf) + (4 * r + k);          // K chunk ch of this row: out[ch * UM_ROWS]
    if (!valid) {
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) out[ch * UM_ROWS] = make_uint4(0, 0, 0, 0);
        return;
    }
    const uint8_t* base = img + (size_t)(xy >> 16) * stride + (xy & 0xFFFFu);
    uint32_t o[T][W];
#pragma unroll
    for (int y = 0; y < T; ++y) load_px<T>(base + (size_t)y * stride, o[y]);
    if (mirror) {                              // flip isometries: the block read right to left
#pragma unroll
        for (int y = 0; y < T; ++y) {
            uint32_t m[W];
#pragma unroll
            for (int wq = 0; wq < W; ++wq) m[wq] = __byte_perm(o[y][W - 1 - wq], 0, 0x0123);
#pragma unroll
            for (int wq = 0; wq < W; ++wq) o[y][wq] = m[wq];
        }
    }
    // transpose: tr[c] byte y = o[y] byte c
    uint32_t tr[T][W];
#pragma unroll
    for (int c = 0; c < T; ++c)
#pragma unroll
        for (int wq = 0; wq < W; ++wq) {
            uint32_t v = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) v |= ((o[4 * wq + j][c >> 2] >> (8 * (c & 3))) & 255u) << (8 * j);
            tr[c][wq] = v;
        }
    // element (Y, X) of the row: k = 0: o[Y][X]; 1: o[X][T-1-Y] = tr[T-1-Y][X]; 2: o[T-1-Y][T-1-X]; 3: o[T-1-X][Y] = tr[Y][T-1-X]
    const bool odd = (k & 1u) != 0, flip = k == 1 || k == 2, rev = k >= 2;
    const __half2 c1024 = __float2half2_rn(1024.f), cm4 = __float2half2_rn(-4.f), c510 = __float2half2_rn(510.f);
    constexpr int RPC = 8 / T;                 // block rows per K chunk of 8 elements (T = 8: 1, T = 4: 2)
#pragma unroll
    for (int ch = 0; ch < N / 8; ++ch) {
        uint32_t hw[4];                        // the chunk as packed fp16 pairs
#pragma unroll
        for (int yy = 0; yy < RPC; ++yy)
#pragma unroll
            for (int wq = 0; wq < W; ++wq) {
                const int Y = ch * RPC + yy;
                const uint32_t a_o = flip ? o[T - 1 - Y][wq] : o[Y][wq], a_t = flip ? tr[T - 1 - Y][wq] : tr[Y][wq];
                const uint32_t b_o = flip ? o[T - 1 - Y][W - 1 - wq] : o[Y][W - 1 - wq], b_t = flip ? tr[T - 1 - Y][W - 1 - wq] : tr[Y][W - 1 - wq];
                const uint32_t x = rev ? __byte_perm(odd ? b_t : b_o, 0, 0x0123) : (odd ? a_t : a_o);
                // bytes -> halves 1024 + p (exponent trick), minus 1024, times -4 plus 510: all exact in fp16
                uint32_t h01 = __byte_perm(x, 0x64646464u, 0x5140), h23 = __byte_perm(x, 0x64646464u, 0x5342);
                const __half2 v01 = __hfma2(__hsub2(*reinterpret_cast<__half2*>(&h01), c1024), cm4, c510);
                const __half2 v23 = __hfma2(__hsub2(*reinterpret_cast<__half2*>(&h23), c1024), cm4, c510);
                hw[(yy * W + wq) * 2] = *reinterpret_cast<const uint32_t*>(&v01);
                hw[(yy * W + wq) * 2 + 1] = *reinterpret_cast<const uint32_t*>(&v23);
            }
        out[ch * UM_ROWS] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    }
    out[(N / 8) * UM_ROWS] = make_uint4(0x68003C00u, 0x00006800u, 0, 0);   // [1, 2048 | 2048, 0 | 0 ...]
#pragma unroll
    for (int ch = N / 8 + 1; ch < NCH; ++ch) out[ch * UM_ROWS] = make_uint4(0, 0, 0, 0);
}

// RETIRE: retire quads of rows after their first threshold hit (T = 4 threshold levels: ALU bound)
// META: runs cross chunks of the blob (brightness-bin neighbourhoods, minimum pass): per-tile metadata is read
// MODE 1: lower-bound prefilter (fe_lb.cu) -- the A tile is bulk-copied from a blob (operands are 16-bit cell sums, not
// pixels) and the epilogue emits every column under the threshold as a candidate instead of tracking first hit / minimum
// PAIR: a work item is TWO row tiles (64 range blocks, M = 256 as two MMAs per B tile into the accumulators of the two compute
// groups): every B stage is used twice, which halves the operand stream into shared memory -- the stream, not the tensor pipe,
// bounds the single-tile form.  Two issuers (tile parity = accumulator buffer) with three B stages each; group g drains row tile g.
template <int T, bool RETIRE, bool META, int MODE, bool PAIR>
__global__ void __launch_bounds__(F16_THREADS, 1) k_search_f16(const F16Args a) {
    constexpr uint32_t N = T * T, KPAD = (N + 3 + 15) & ~15u;
    constexpr uint32_t bytesA = UM_ROWS * KPAD * 2, bytesB = UM_NT * KPAD * 2;
    constexpr uint32_t AT = PAIR ? 2 : 1;                  // row tiles per item / per A buffer
    constexpr uint32_t NA = PAIR ? 2 : F16_ABUFS;          // A buffers in flight: the builders run up to NA - 1 items ahead
    constexpr uint32_t NST = PAIR ? 6 : F16_STAGES;        // B stages
    constexpr uint32_t NISS = PAIR ? 2 : F16_ISSUERS;      // issuers with work; tiles are dealt to them round-robin
    constexpr uint32_t SPI = NST / NISS;                   // B stages per issuer
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const SliceCtl* __restrict__ ctl = a.ctl;
    if (ctl->active != a.ordinal) return;                  // this slice was never planned (the level ended earlier)
    const uint32_t n_items = ctl->n_items;
    if (blockIdx.x >= n_items) return;
    const uint32_t no_min = ctl->no_min;
    const ListEntry* __restrict__ list = a.list[ctl->list];

    uint8_t* sA = smem;                                    // NA buffers
    uint8_t* sB = smem + NA * AT * bytesA;                 // NST stages
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)NST * bytesB);
    const uint32_t bar0 = smem_u32(bars);
    auto ACC_FULL = [&](uint32_t g, uint32_t b) { return bar0 + 8 * (0 + 2 * g + b); };
    auto ACC_EMPTY = [&](uint32_t g, uint32_t b) { return bar0 + 8 * (4 + 2 * g + b); };
    auto B_FULL = [&](uint32_t i) { return bar0 + 8 * (8 + i); };
    auto A_FULL = [&](uint32_t i) { return bar0 + 8 * (8 + NST + i); };
    auto A_EMPTY = [&](uint32_t i) { return bar0 + 8 * (8 + NST + NA + i); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8 + NST + 2 * NA);

    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < NA; ++i) {
            mbar_init(A_FULL(i), F16_BUILDERS);            // one arrive per builder warp
            mbar_init(A_EMPTY(i), NISS);
        }
        for (uint32_t i = 0; i < 4; ++i) {
            mbar_init(bar0 + 8 * (0 + i), 1);              // ACC_FULL: one tcgen05.commit
            mbar_init(bar0 + 8 * (4 + i), 8);              // ACC_EMPTY: one arrive per compute warp of the warpgroup
        }
        for (uint32_t i = 0; i < NST; ++i) mbar_init(B_FULL(i), 1);
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

    if (warp < F16_ISSUERS) {
        // ================= issuers: one thread per accumulator (g, ib), each also the producer of its B tiles =================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (lane == 0 && warp < NISS) {
            // single: tile t of the CTA's running count goes to group t % 2, buffer (t / 2) % 2; pair: to buffer t % 2 of both groups
            const uint32_t g = PAIR ? 1u : warp >> 1, ib = PAIR ? warp : warp & 1, res = PAIR ? warp : g + 2 * ib;
            const uint32_t sb = SPI * warp;                // first of this issuer's stages
            const uint32_t idesc = (1u << 4) | ((uint32_t)(UM_NT >> 3) << 17) | ((uint32_t)(UM_ROWS >> 4) << 24);
            constexpr uint32_t nk = KPAD / 16;
            struct Cursor {
                uint32_t w, wi, it0, u, n, t0, nrows;
                bool valid;
            };
            auto load_item = [&](Cursor& c) {
                c.valid = c.w < n_items;
                if (c.valid) {
                    const uint4 rec = __ldg(reinterpret_cast<const uint4*>(a.items + c.w));
                    c.t0 = rec.z; c.n = rec.w - rec.z; c.nrows = rec.y;
                    if (c.w + gridDim.x < n_items) asm volatile("prefetch.global.L1 [%0];" ::"l"(a.items + c.w + gridDim.x));
                }
            };
            Cursor ld, mm;
            ld.w = blockIdx.x; ld.wi = 0; ld.it0 = 0; ld.u = 0;
            load_item(ld);
            mm = ld;
            // the MMA cursor signs off every item it leaves (A_EMPTY needs all four issuers)
            bool mm_had_tiles = false;
            auto next_item = [&](Cursor& c, bool is_mm) {
                if (is_mm) {
                    // An issuer without a tile in this item must not sign it off before the item's A tile exists: it could run
                    // NA items ahead and arrive twice on the same A_EMPTY phase.  Waiting for A_FULL orders it behind the
                    // completion of the item NA back (the builders wait for that), like the issuers that do have tiles.
                    if (!mm_had_tiles) mbar_wait(A_FULL(c.wi % NA), (c.wi / NA) & 1);
                    if (mm_had_tiles) tc_commit(A_EMPTY(c.wi % NA)); else mbar_arrive(A_EMPTY(c.wi % NA));
                    mm_had_tiles = false;
                }
                c.it0 += c.n; c.w += gridDim.x; ++c.wi;
                load_item(c);
            };
            auto seek = [&](Cursor& c, bool is_mm) {
                while (c.valid) {
                    c.u = (res + NISS - (c.it0 % NISS)) % NISS;
                    if (c.u < c.n) return;
                    next_item(c, is_mm);
                }
            };
            auto advance = [&](Cursor& c, bool is_mm) {
                c.u += NISS;
                if (c.u >= c.n) { next_item(c, is_mm); seek(c, is_mm); }
            };
            seek(ld, false);
            seek(mm, true);
            uint32_t m_ld = 0, m = 0;                      // tiles loaded / issued by this issuer so far
            auto load_B = [&]() {                           // bulk copy of the load cursor's tile into stage sb + (m_ld & 1)
                const uint32_t s = sb + (m_ld % SPI);
                if (m_ld >= SPI) mbar_wait(ACC_FULL(g, ib), (m_ld - SPI) & 1);   // the stage's previous tile: its MMAs have completed
                mbar_expect_tx(B_FULL(s), bytesB);
                bulk_g2s(smem_u32(sB + (size_t)s * bytesB), reinterpret_cast<const uint8_t*>(a.B16) + (size_t)(ld.t0 + ld.u) * bytesB, bytesB, B_FULL(s));
                ++m_ld;
                advance(ld, false);
            };
#pragma unroll
            for (uint32_t i = 0; i < SPI; ++i)
                if (ld.valid) load_B();
            uint32_t cur_wi = 0xFFFFFFFFu, a_addr = 0;
            while (mm.valid) {
                if (m >= 1 && ld.valid) load_B();          // tile m + SPI - 1 goes into the stage of tile m - 1
                if (mm.wi != cur_wi) {                      // first tile of this issuer in a new item: its A tile must be built
                    cur_wi = mm.wi;
                    mbar_wait(A_FULL(cur_wi % NA), (cur_wi / NA) & 1);
                    a_addr = smem_u32(sA + (cur_wi % NA) * (AT * bytesA));
                }
                const uint32_t s = sb + (m % SPI);
                mbar_wait(B_FULL(s), (m / SPI) & 1);
                mbar_wait(ACC_EMPTY(g, ib), (m & 1) ^ 1);
                if (PAIR) mbar_wait(ACC_EMPTY(0, ib), (m & 1) ^ 1);
                tc_fence_after();
                const uint32_t b_addr = smem_u32(sB + (size_t)s * bytesB);
#pragma unroll
                for (uint32_t rt = 0; rt < AT; ++rt) {
                    if (PAIR && rt == 1 && mm.nrows <= (uint32_t)UM_ROWS) break;     // a lone row tile at the end of a bucket
                    const uint32_t d_tmem = tmem_base + ((PAIR ? rt : g) * 2 + ib) * UM_NT;
#pragma unroll
                    for (uint32_t kk = 0; kk < nk; ++kk) {
                        const uint64_t adesc = make_desc(a_addr + rt * bytesA + kk * 2 * (UM_ROWS * 16), UM_ROWS * 16, 128);
                        const uint64_t bdesc = make_desc(b_addr + kk * 2 * (UM_NT * 16), UM_NT * 16, 128);
                        tc_mma<0>(d_tmem, adesc, bdesc, idesc, kk > 0 ? 1u : 0u);
                    }
                    if (PAIR && rt == 0) tc_commit(ACC_FULL(0, ib));   // group 0 starts while the second row tile multiplies
                }
                tc_commit(ACC_FULL(g, ib));
                mm_had_tiles = true;
                ++m;
                advance(mm, true);
            }
        }
    } else if (warp < F16_ISSUERS + F16_BUILDERS) {
        // ================= A builders =================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        // Lane = (range block lane >> 2 of this warp's eight, rotation lane & 3).  The chain item record -> list entry ->
        // pixels is three dependent loads; it is software-pipelined: the entry of the NEXT item is fetched while this item is
        // built and its pixel rows are prefetched into L1, so an item costs one L1 round trip plus the arithmetic.
        const uint32_t bw = warp - F16_ISSUERS;
        const uint32_t r = bw * 8 + (lane >> 2);
        // entry = (pixel origin, mirror flag) of this lane's range block in each of the item's row tiles
        auto fetch = [&](uint32_t w, uint2 (&ent)[AT], bool (&valid)[AT]) {
#pragma unroll
            for (uint32_t rt = 0; rt < AT; ++rt) valid[rt] = false;
            if (w < n_items) {
                const uint4 rec = __ldg(reinterpret_cast<const uint4*>(a.items + w));
#pragma unroll
                for (uint32_t rt = 0; rt < AT; ++rt) {
                    valid[rt] = 4 * (r + 32 * rt) < rec.y;
                    if (valid[rt]) {
                        const uint4 e = __ldg(reinterpret_cast<const uint4*>(list + rec.x + r + 32 * rt));
                        ent[rt] = make_uint2(e.y, e.w);
                        const uint8_t* base = a.img + (size_t)(e.y >> 16) * a.stride + (e.y & 0xFFFFu);
#pragma unroll
                        for (int y = 0; y < T; ++y) asm volatile("prefetch.global.L1 [%0];" ::"l"(base + (size_t)y * a.stride));
                    }
                }
            }
        };
        uint2 ent[AT], ent_next[AT];
        bool valid[AT], valid_next[AT];
#pragma unroll
        for (uint32_t rt = 0; rt < AT; ++rt) { ent[rt] = ent_next[rt] = make_uint2(0, 0); valid[rt] = valid_next[rt] = false; }
        if (MODE == 0) fetch(blockIdx.x, ent, valid);
        uint32_t wi = 0;
        if (MODE == 1) {
            // the A tile of the item is one bulk copy from the slice's blob; the other three builder warps only sign the barrier
            if (lane == 0)
                for (uint32_t w = blockIdx.x; w < n_items; w += gridDim.x, ++wi) {
                    const uint32_t ab = wi % NA;
                    if (wi >= NA) mbar_wait(A_EMPTY(ab), ((wi / NA) & 1) ^ 1);
                    if (bw == 0) {
                        const uint32_t a_tile = __ldg(&a.items[w].a_tile);
                        mbar_expect_tx(A_FULL(ab), bytesA);
                        bulk_g2s(smem_u32(sA + ab * bytesA), reinterpret_cast<const uint8_t*>(a.A16) + (size_t)a_tile * bytesA, bytesA, A_FULL(ab));
                    } else {
                        mbar_arrive(A_FULL(ab));
                    }
                }
        } else {
        for (uint32_t w = blockIdx.x; w < n_items; w += gridDim.x, ++wi) {
            const uint32_t ab = wi % NA;
            fetch(w + gridDim.x, ent_next, valid_next);
            if (wi >= NA) mbar_wait(A_EMPTY(ab), ((wi / NA) & 1) ^ 1);   // all four issuers are done with item wi - NA
#pragma unroll
            for (uint32_t rt = 0; rt < AT; ++rt) {
                // the second row tile of a lone-tile item is never multiplied: leave it as it is
                if (rt == 0 || __any_sync(0xFFFFFFFFu, valid[rt]))
                    build_rows<T>(sA + (ab * AT + rt) * bytesA, a.img, a.stride, ent[rt].x, ent[rt].y != 0, valid[rt], bw, lane);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to tcgen05.mma
            __syncwarp();
            if (lane == 0) mbar_arrive(A_FULL(ab));
#pragma unroll
            for (uint32_t rt = 0; rt < AT; ++rt) { ent[rt] = ent_next[rt]; valid[rt] = valid_next[rt]; }
        }
        }
    } else {
        // ================= compute warps: TMEM -> registers -> row argmin =================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 96;");
        const uint32_t cw = warp - F16_ISSUERS - F16_BUILDERS;
        const uint32_t g = cw >> 3;
        const uint32_t h = (cw >> 2) & 1;             // column half
        const uint32_t sp = warp & 3;                 // TMEM sub-partition this warp may read
        const uint32_t lrow = sp * 32 + lane;         // row inside the tile == TMEM lane
        const uint32_t lane_addr = tmem_base + ((sp * 32u) << 16) + h * UM_HALF;
        uint32_t it0 = 0, jb = 0;
        for (uint32_t w = blockIdx.x; w < n_items; w += gridDim.x) {
            const uint4 rec = __ldg(reinterpret_cast<const uint4*>(a.items + w));
            const uint32_t cols_left = __ldg(&a.items[w].cols_left);
            const uint32_t item_t0 = rec.z, n = rec.w - rec.z;
            const uint32_t trow = PAIR ? g * UM_ROWS + lrow : lrow;          // row inside the item
            const bool row_ok = trow < rec.y;
            const bool tile_empty = PAIR && g * UM_ROWS >= rec.y;          // lone row tile: this group only keeps the hand-shake going
            uint32_t slot = 0, a2 = 0;
            if (row_ok) {
                const uint4 ent = __ldg(reinterpret_cast<const uint4*>(list + rec.x + (trow >> 2)));
                slot = ent.x; a2 = ent.z;
            }
            const uint32_t srow = 4u * slot + (lrow & 3u);     // result row of the level
            RowState st;
            st.bestV = 3.0e38f; st.bestp = 0; st.bestcol = FE_NONE32; st.hit = FE_NONE32;
            st.vthr0 = -3.0e38f; st.vthr1 = -3.0e38f;
            st.par_item = reinterpret_cast<const uint32_t*>(a.colmeta + (size_t)item_t0 * 2);
            uint32_t cur_seg = FE_NONE32;       // chunk of the tiles this thread is scanning
            if (a.use_thr && row_ok) { // n16 <= thr16  <=>  V <= floor((thr16 - a2 - p) / 2)
                const long long tt = (long long)a.thr16 - (long long)a2;
                long long f0 = tt >= 0 ? tt / 2 : -((-tt + 1) / 2);
                long long f1 = (tt - 1) >= 0 ? (tt - 1) / 2 : -((-(tt - 1) + 1) / 2);
                if (MODE == 1) {
                    // prefilter operands are 12-bit: V runs to +-2^29 and is not exact -- the threshold is rounded UP, the test
                    // stays conservative (fe_lb.cu)
                    st.vthr0 = __ll2float_ru(f0);
                    st.vthr1 = __ll2float_ru(f1);
                } else {
                    f0 = max(-16777216ll, min(16777215ll, f0));
                    f1 = max(-16777216ll, min(16777215ll, f1));
                    st.vthr0 = (float)f0;
                    st.vthr1 = (float)f1;
                }
            }
            if (row_ok && !no_min) {
                // Seed the running minimum with what earlier slices / column chunks already found for this row, plus one:
                // anything this item finds at or below the recorded score still gets written (ties are settled by the
                // domain index inside the 64-bit key), everything above it is rejected by the cheap tile-minimum test.
                const unsigned long long seen = __ldcg(&a.rowbest[srow]);
                if (seen != FE_INF64) {
                    const long long d = (long long)(uint32_t)(seen >> 32) + 1ll - (long long)a2;   // = 2 V + p
                    const long long V = d >= 0 ? d / 2 : -((-d + 1) / 2);
                    if (V > -16777216ll && V < 16777216ll) { st.bestV = (float)V; st.bestp = (uint32_t)(d - 2 * V); }
                }
            }
            // single: the two groups take alternate tiles of the item; pair: every tile, for their own row tile
            const uint32_t tstep = PAIR ? 1u : (uint32_t)UM_WGS;
            const uint32_t first = PAIR ? 0u : (g + UM_WGS - (it0 % UM_WGS)) % UM_WGS;
            const uint32_t my_tiles = first < n ? (n - first + tstep - 1) / tstep : 0;
            // Threshold runs: a range is decided by its first hit in scan order, so once any of its four rotation rows
            // (four adjacent lanes) has crossed the threshold the whole quad is retired for the rest of the chunk, and a
            // warp whose 32 rows are all retired only keeps the accumulator hand-shake going.
            bool retired = !row_ok;
            bool warp_done = tile_empty;
            for (uint32_t j = 0; j < my_tiles; ++j, ++jb) {
                const uint32_t u = first + j * tstep, buf = jb & 1;
                const uint32_t colbase = u * UM_NT + h * UM_HALF;
                static_assert(UM_HALF == 64, "two parity words per half tile");
                const uint4* meta_p = a.colmeta + (size_t)(item_t0 + u) * 2 + h;
                uint4 meta = make_uint4(0, 0, 0, 0);
                ParitySrc par;
                par.ptr = reinterpret_cast<const uint32_t*>(meta_p);
                uint32_t nvalid;
                if (META) {
                    meta = __ldg(meta_p);
                    par.reg[0] = meta.x; par.reg[1] = meta.y;
                    nvalid = meta.z;
                } else {
                    const uint32_t tile_valid = min((uint32_t)UM_NT, cols_left - u * UM_NT);
                    nvalid = tile_valid > h * UM_HALF ? min((uint32_t)UM_HALF, tile_valid - h * UM_HALF) : 0u;
                }
                // The run moves into another chunk: domain indices restart low there, so the first hit found so far is only
                // the first of the chunk behind us -- bank it and start over.
                auto chunk_change = [&]() {
                    if (META && meta.w != cur_seg) {
                        if (st.hit != FE_NONE32) {
                            atomicMin(&a.rowhit[srow], a.blob_dom[(size_t)item_t0 * UM_NT + st.hit]);
                            st.hit = FE_NONE32;
                        }
                        // Same for the running minimum: "first column wins a tie" is only "smallest domain index wins" inside a
                        // chunk.  Bank it (the 64-bit key settles ties by domain index) and go on from its score plus one, so
                        // that an equal score in the next chunk is still looked at.
                        if (st.bestcol != FE_NONE32) {
                            if (st.bestp == 2) st.bestp = row_parity(st, st.bestcol);
                            const long long d = 2ll * (long long)st.bestV + (long long)st.bestp;
                            atomicMin(&a.rowbest[srow], ((unsigned long long)(uint32_t)((long long)a2 + d) << 32) |
                                                            (unsigned long long)a.blob_dom[(size_t)item_t0 * UM_NT + st.bestcol]);
                            if (st.bestV >= 16777216.0f - 64.0f) atomicOr(a.flags, 1u);
                            const long long V = (d + 1) >= 0 ? (d + 1) / 2 : -((-(d + 1) + 1) / 2);
                            st.bestV = (float)V; st.bestp = (uint32_t)(d + 1 - 2 * V); st.bestcol = FE_NONE32;
                        }
                        cur_seg = meta.w;
                        if (RETIRE) { retired = !row_ok; warp_done = tile_empty; }
                    }
                };
                if (RETIRE) chunk_change();
                uint32_t v[UM_HALF];
                mbar_wait(ACC_FULL(g, buf), (jb >> 1) & 1);
                tc_fence_after();
                if (!((RETIRE || PAIR) && warp_done)) {
                    const uint32_t taddr = lane_addr + (g * 2 + buf) * UM_NT;
                    TMEM_LD32(taddr, (v + 0));
                    TMEM_LD32(taddr + 32, (v + 32));
                    tmem_wait_ld();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(ACC_EMPTY(g, buf));
                if (!RETIRE) chunk_change();
                if (MODE == 1) {
                    const size_t col0 = (size_t)item_t0 * UM_NT + colbase;
                    candidates_half<META>(v, st.vthr0, st.vthr1, row_ok, nvalid, par, [&](uint32_t c) {
                        const uint32_t idx = atomicAdd(a.cand_count, 1u);
                        if (idx < a.cand_cap) a.cand[idx] = make_uint2(srow, a.blob_dom[col0 + c]);
                    });
                } else if (!((RETIRE || PAIR) && warp_done)) {
                    process_half<META>(v, st, RETIRE ? !retired : row_ok, colbase, nvalid, par, no_min == 0);
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
                    const unsigned long long key = ((unsigned long long)(uint32_t)n16 << 32) | (unsigned long long)a.blob_dom[(size_t)item_t0 * UM_NT + st.bestcol];
                    atomicMin(&a.rowbest[srow], key);
                    if (st.bestV >= 16777216.0f - 64.0f) atomicOr(a.flags, 1u);
                }
                if (st.hit != FE_NONE32) atomicMin(&a.rowhit[srow], a.blob_dom[(size_t)item_t0 * UM_NT + st.hit]);
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

// B columns of the level blob.  One CTA per blob tile (128 columns); thread = column.  b = D - 510; limbs of
// h = floor(sum b^2 / 2) in the three K columns after the data; per half tile: parity words, valid columns, chunk id.
template <int T>
__global__ void __launch_bounds__(128) k_build_pool16_level(const uint8_t* __restrict__ img, uint32_t stride, const fe_grid_item* __restrict__ dom,
                                                            const uint32_t* __restrict__ order, const LevelPlan* __restrict__ plan,
                                                            uint4* __restrict__ B16, uint32_t* __restrict__ colmeta, uint32_t* __restrict__ blob_dom) {
    constexpr int N = T * T, KPAD = (N + 3 + 15) & ~15, NCH = KPAD / 8, W = T / 4;
    __shared__ uint32_t s_chunk, s_col0, s_end, s_doff;
    const uint32_t tile = blockIdx.x, l = threadIdx.x;
    if (tile >= plan->n_tiles) return;
    if (l == 0) {
        const uint32_t nb = plan->nb;
        uint32_t lo = 0, hi = plan->nk * nb - 1;          // chunk with tile0[chunk] <= tile < tile0[chunk + 1]
        while (lo < hi) {
            const uint32_t mid = (lo + hi + 1) >> 1;
            if (plan->tile0[mid] <= tile) lo = mid; else hi = mid - 1;
        }
        const uint32_t k = lo / nb, b = lo - k * nb;
        s_chunk = lo;
        s_col0 = (k ? plan->dend[k - 1][b] : 0u) + (tile - plan->tile0[lo]) * UM_NT;
        s_end = plan->dend[k][b];
        s_doff = plan->doff[b];
    }
    __syncthreads();
    const uint32_t col = s_col0 + l, end = s_end;
    const bool live = col < end;
    uint4* out = B16 + (size_t)tile * NCH * UM_NT + l;
    uint32_t s2 = 0;
    uint32_t di = FE_NONE32;
    if (live) di = order ? order[s_doff + col] : s_doff + col;
    blob_dom[(size_t)tile * UM_NT + l] = di;
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
    // tile metadata, 4 words per half tile: parity of columns 0-31, 32-63, valid columns of the half, chunk of the tile
    const uint32_t par = __ballot_sync(0xFFFFFFFFu, live && (s2 & 1u));
    if ((l & 31) == 0) {
        uint32_t* m = colmeta + ((size_t)tile * 2 + (l >> 6)) * 4;
        m[(l >> 5) & 1u] = par;
        if ((l & 63) == 0) {
            m[2] = col >= end ? 0u : min(end - col, (uint32_t)UM_HALF);
            m[3] = s_chunk;
        }
    }
}

} // namespace

int f16_level_supported(const LevelGeom& g) { return g.fast && (g.T == 4 || g.T == 8); }

int f16_build_pool(fe_ctx* ctx, const LevelGeom& g, const fe_grid_item* d_dom, const uint32_t* dom_order, const LevelPlan* plan, uint32_t max_tiles) {
    const uint32_t Kpad = (g.N + 3 + 15u) & ~15u;
    FE_CUDA(ctx, ctx->b_B16.ensure((size_t)max_tiles * UM_NT * Kpad * 2 + 256));
    FE_CUDA(ctx, ctx->b_tmaps.ensure((size_t)max_tiles * 32 + 64));
    FE_CUDA(ctx, ctx->b_blob_dom.ensure((size_t)max_tiles * UM_NT * 4 + 64));
    if (g.T == 4)
        k_build_pool16_level<4><<<max_tiles, 128, 0, ctx->stream>>>(ctx->src.px, ctx->src.stride, d_dom, dom_order, plan, ctx->b_B16.as<uint4>(),
                                                                    ctx->b_tmaps.as<uint32_t>(), ctx->b_blob_dom.as<uint32_t>());
    else
        k_build_pool16_level<8><<<max_tiles, 128, 0, ctx->stream>>>(ctx->src.px, ctx->src.stride, d_dom, dom_order, plan, ctx->b_B16.as<uint4>(),
                                                                    ctx->b_tmaps.as<uint32_t>(), ctx->b_blob_dom.as<uint32_t>());
    FE_CUDA(ctx, cudaGetLastError());
    ctx->stats.kernel_launches++;
    return FE_OK;
}

template <int T, bool PAIR>
static int launch_T(fe_ctx* ctx, const F16Args& a, bool retire, bool meta, cudaEvent_t ev0, cudaEvent_t ev1) {
    constexpr uint32_t Kpad = (T * T + 3 + 15u) & ~15u;
    constexpr uint32_t na = PAIR ? 2 : F16_ABUFS, at = PAIR ? 2 : 1, nst = PAIR ? 6 : F16_STAGES;
    const size_t smem = (size_t)na * at * UM_ROWS * Kpad * 2 + (size_t)nst * UM_NT * Kpad * 2 + (8 + nst + 2 * na) * 8 + 64;
    auto kern = retire ? (meta ? k_search_f16<T, true, true, 0, PAIR> : k_search_f16<T, true, false, 0, PAIR>)
                       : (meta ? k_search_f16<T, false, true, 0, PAIR> : k_search_f16<T, false, false, 0, PAIR>);
    FE_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    if (ev0) cudaEventRecord(ev0, ctx->stream);
    kern<<<ctx->n_sm, F16_THREADS, smem, ctx->stream>>>(a);
    FE_CUDA(ctx, cudaGetLastError());
    if (ev1) cudaEventRecord(ev1, ctx->stream);
    ctx->stats.kernel_launches++;
    return FE_OK;
}

int f16_launch_search(fe_ctx* ctx, const LevelGeom& g, const F16Args& a, bool retire, bool meta, cudaEvent_t ev0, cudaEvent_t ev1) {
    if (a.pair) return g.T == 4 ? launch_T<4, true>(ctx, a, retire, meta, ev0, ev1) : launch_T<8, true>(ctx, a, retire, meta, ev0, ev1);
    return g.T == 4 ? launch_T<4, false>(ctx, a, retire, meta, ev0, ev1) : launch_T<8, false>(ctx, a, retire, meta, ev0, ev1);
}

// lower-bound prefilter launch: the T' = 8 contraction on cell sums, candidates out (fe_lb.cu)
int f16_launch_search_lb(fe_ctx* ctx, const F16Args& a, cudaEvent_t ev0, cudaEvent_t ev1) {
    constexpr uint32_t Kpad = (8 * 8 + 3 + 15u) & ~15u;
    const size_t smem = (size_t)F16_ABUFS * UM_ROWS * Kpad * 2 + (size_t)F16_STAGES * UM_NT * Kpad * 2 + (8 + F16_STAGES + 2 * F16_ABUFS) * 8 + 64;
    auto kern = k_search_f16<8, false, true, 1, false>;
    FE_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    if (ev0) cudaEventRecord(ev0, ctx->stream);
    kern<<<ctx->n_sm, F16_THREADS, smem, ctx->stream>>>(a);
    FE_CUDA(ctx, cudaGetLastError());
    if (ev1) cudaEventRecord(ev1, ctx->stream);
    ctx->stats.kernel_launches++;
    return FE_OK;
}
