// fe_search_umma_i8.cu -- tcgen05 search for large blocks (T >= 16): kind::i8 MMAs with exact s32 accumulators.
//
// fp32 accumulation stops being integer-exact when K = T^2 grows (sum r*D reaches 2.7e8 at T=32), so the large levels
// use the integer tensor path of sm_100a:  A = range pixels r (u8, inverse-rotated rows), B = the domain box sums D
// split into a low-byte plane and a high-byte plane (D <= 1020 -> high byte 0..3), two s32 accumulators per tile:
//     cross = acc_lo + 256 * acc_hi,      n16 = 16 sum r^2 - 8 cross + sum D^2      (exact, any input)
// Two planes at the 2x int8 rate cost the same tensor time as one fp16 pass.  Tile = 128 rows x 64 domain columns
// (2 x 64 TMEM columns per buffer, two buffers per warpgroup, two warpgroups = all 512 columns); K is streamed in
// stages of 256 bytes (32 KB per stage: both planes), the A tile (128 x K bytes, up to 128 KB at T=32) stays in
// shared memory for the whole work item.  Same warp roles as the fp16 kernel (fe_search_umma.cu).
#include <cstdio>
#include <cstdlib>

#include "fe_kernels.cuh"
#include "fe_umma_dev.cuh"

using namespace umma_dev;

namespace {

template <int DUMMY>
__global__ void __launch_bounds__(UM_THREADS_I8, 1) k_search_umma_i8(const UmmaArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t Kpad = a.Kpad;
    const uint32_t kc = min(Kpad, (uint32_t)I8_KC);          // bytes of K per stage
    const uint32_t nch = Kpad / kc;                          // stages per tile
    const uint32_t bytesA = UM_ROWS * Kpad, bytesB = 2 * I8_NT * kc;
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
            mbar_init(B_EMPTY(i), 1);
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

    if (warp == 0) {
        // ================= producer =================
        if (lane == 0) {
            uint32_t ic = 0, wi = 0; // running stage counter, item counter
            for (uint32_t w = blockIdx.x; w < a.total_items; w += gridDim.x, ++wi) {
                const WorkItem item = decode_item(a, w);
                const uint32_t ab = wi % NA;
                mbar_wait(A_EMPTY(ab), ((wi / NA) & 1) ^ 1);
                mbar_expect_tx(A_FULL(ab), bytesA);
                // the bulk copy engine takes at most ~1 MB per request; 128 KB tiles go as 32 KB pieces
                for (uint32_t off = 0; off < bytesA; off += 32768) {
                    const uint32_t n = min(32768u, bytesA - off);
                    bulk_g2s(smem_u32(sA + ab * bytesA + off), reinterpret_cast<const uint8_t*>(a.A16) + (size_t)item.a_blob * bytesA + off, n, A_FULL(ab));
                }
                for (uint32_t t = item.t0; t < item.t1; ++t)
                    for (uint32_t c = 0; c < nch; ++c, ++ic) {
                        const uint32_t s = ic % S;
                        mbar_wait(B_EMPTY(s), ((ic / S) & 1) ^ 1);
                        mbar_expect_tx(B_FULL(s), bytesB);
                        bulk_g2s(smem_u32(sB + (size_t)s * bytesB), reinterpret_cast<const uint8_t*>(a.B16) + ((size_t)t * nch + c) * bytesB, bytesB, B_FULL(s));
                    }
            }
        } else if (lane == 1) {
            // forwarder: a stage ring shared by two issuers means an issuer may look at a stage barrier that is still
            // TWO phases behind the one it needs, which a parity wait cannot tell from "done".  This lane observes
            // every phase in order (never ambiguous) and publishes the count of landed stages.
            uint32_t ic = 0;
            for (uint32_t w = blockIdx.x; w < a.total_items; w += gridDim.x) {
                const WorkItem item = decode_item(a, w);
                const uint32_t total = (item.t1 - item.t0) * nch;
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
            for (uint32_t w = blockIdx.x; w < a.total_items; w += gridDim.x, ++wi) {
                const WorkItem item = decode_item(a, w);
                const uint32_t ab = wi % NA, n = item.t1 - item.t0;
                const uint32_t first = (g + UM_WGS - (it0 % UM_WGS)) % UM_WGS;
                mbar_wait(A_FULL(ab), (wi / NA) & 1);
                const uint32_t a_addr = smem_u32(sA + ab * bytesA);
                bool any = false;
                for (uint32_t u = first; u < n; u += UM_WGS, ++jb) {
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
        for (uint32_t w = blockIdx.x; w < a.total_items; w += gridDim.x) {
            const WorkItem item = decode_item(a, w);
            const uint32_t n = item.t1 - item.t0;
            const bool row_ok = lrow < item.nrows;
            const uint32_t grow = item.row0 + lrow;
            const uint32_t rc = row_ok ? a.rowA2[grow >> 2] : 0u;       // 16 * sum r^2
            const uint32_t srow = (row_ok && a.rowslot) ? 4u * a.rowslot[grow >> 2] + (grow & 3u) : grow;   // result slot of the level
            // w = sum D^2 - 8 cross (signed); n16 = rc + w;  n16 <= thr16  <=>  w <= thr16 - rc
            long long wt = (long long)a.thr16 - (long long)rc;
            wt = max(-2147483647ll, min(2147483646ll, wt));                  // INT_MAX is the padding columns' score
            const int wthr = (a.use_thr && row_ok) ? (int)wt : (int)0x80000000;
            int bestw = 0x7FFFFFFF;
            uint32_t bestcol = FE_NONE32, hit = FE_NONE32;
            uint32_t cur_seg = FE_NONE32;       // domain bucket of the tiles this thread is scanning
            const uint32_t first = (g + UM_WGS - (it0 % UM_WGS)) % UM_WGS;
            const uint32_t my_tiles = first < n ? (n - first + UM_WGS - 1) / UM_WGS : 0;
            for (uint32_t j = 0; j < my_tiles; ++j, ++jb) {
                const uint32_t u = first + j * UM_WGS, buf = jb & 1;
                const uint32_t colbase = u * I8_NT + h * 32;             // column inside the item
                const uint32_t taddr = lane_addr + (g * 2 + buf) * 2 * I8_NT;
                // work items that run over several domain buckets (brightness bins): bucket id of the tile, loaded ahead of the
                // accumulator wait and looked at after the accumulator has been read
                const uint32_t seg = a.meta ? __ldg(a.tileseg + item.t0 + u) : 0u;
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
                    // another domain bucket: its columns restart at low domain indices, so bank the first hit of the bucket
                    // behind us and look for this bucket's
                    if (hit != FE_NONE32) {
                        atomicMin(&a.rowhit[srow], a.blob_dom[(size_t)item.t0 * I8_NT + hit]);
                        hit = FE_NONE32;
                    }
                    cur_seg = seg;
                }
                // sum(D^2) per column, stored by padded (tile, column) position so the 16-byte loads stay aligned
                const uint4* cn4 = reinterpret_cast<const uint4*>(a.coln + (size_t)(item.t0 + u) * I8_NT + h * 32);
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
                const bool improve = row_ok && !a.no_min && tmin < bestw;
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
                    const unsigned long long key = ((unsigned long long)n16 << 32) | (unsigned long long)a.blob_dom[(size_t)item.t0 * I8_NT + bestcol];
                    atomicMin(&a.rowbest[srow], key);
                }
                if (hit != FE_NONE32) atomicMin(&a.rowhit[srow], a.blob_dom[(size_t)item.t0 * I8_NT + hit]);
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
// operand blobs (u8).  A: [row tile][K/16][128 rows][16 B].  B: [col tile][K stage][plane lo,hi][kc/16][64 cols][16 B].
// ---------------------------------------------------------------------------------------------------
// A: one CTA per row tile (32 ranges).  The 32 blocks are staged in shared memory with coalesced word loads (their
// 16 * sum r^2 falls out on the way), then every thread assembles 16-byte K chunks of the four rotations from shared memory
// -- rotations 0 and 2 are straight / reversed 16-byte runs, 1 and 3 are strided byte gathers that never leave the SM.
// Every byte of the blob is written exactly once (padding rows as zeros): no memset.
__global__ void __launch_bounds__(256) k_build_rows_i8(const uint8_t* __restrict__ img, uint32_t stride, const fe_grid_item* __restrict__ rng,
                                                       const uint32_t* __restrict__ order, UmmaBuckets bk, uint32_t T, uint32_t Kpad,
                                                       uint4* __restrict__ A8, uint32_t* __restrict__ rowc) {
    extern __shared__ __align__(16) uint8_t sblk[];            // [32][N]
    __shared__ uint32_t s2[32];
    __shared__ fe_grid_item sit[32];
    const uint32_t tile = blockIdx.x, N = T * T, nch = Kpad / 16;
    int bi = 0;
    while (bi + 1 < bk.nb && tile >= bk.row_tile0[bi + 1]) ++bi;
    const uint32_t j0 = bk.range_off[bi] + (tile - bk.row_tile0[bi]) * 32;
    const uint32_t jend = bk.range_off[bi + 1];
    const uint32_t nvalid = j0 < jend ? min(32u, jend - j0) : 0u;
    if (threadIdx.x < 32) {
        s2[threadIdx.x] = 0;
        if (threadIdx.x < nvalid) sit[threadIdx.x] = rng[order ? order[j0 + threadIdx.x] : j0 + threadIdx.x];
    }
    __syncthreads();
    // ---- stage the blocks ----
    const bool words = (T & 3u) == 0 && (stride & 3u) == 0 && (reinterpret_cast<uintptr_t>(img) & 3u) == 0;
    if (words) {
        const uint32_t wpr = T / 4, wpb = N / 4;              // words per block row / per block
        for (uint32_t idx = threadIdx.x; idx < nvalid * wpb; idx += blockDim.x) {
            const uint32_t lr = idx / wpb, e = idx - lr * wpb, y = e / wpr, xw = e - y * wpr;
            const fe_grid_item it = sit[lr];
            const uint8_t* p = img + (size_t)(it.y + y) * stride + it.x + 4 * xw;
            uint32_t v;
            if ((it.x & 3u) == 0) v = __ldg(reinterpret_cast<const uint32_t*>(p));
            else v = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
            reinterpret_cast<uint32_t*>(sblk)[lr * wpb + e] = v;
            uint32_t q = __dp4a(v, v, 0u);
            if ((wpb & 31u) == 0 && (nvalid * wpb) % blockDim.x == 0) {   // a warp lies inside one block: one atomic per warp
                for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xFFFFFFFFu, q, o);
                if ((threadIdx.x & 31u) == 0) atomicAdd(&s2[lr], q);
            } else {
                atomicAdd(&s2[lr], q);
            }
        }
    } else {
        for (uint32_t idx = threadIdx.x; idx < nvalid * N; idx += blockDim.x) {
            const uint32_t lr = idx / N, e = idx - lr * N, y = e / T, x = e - y * T;
            const fe_grid_item it = sit[lr];
            const uint32_t v = img[(size_t)(it.y + y) * stride + it.x + x];
            sblk[lr * N + e] = (uint8_t)v;
            atomicAdd(&s2[lr], v * v);
        }
    }
    __syncthreads();
    if (threadIdx.x < nvalid) rowc[j0 + threadIdx.x] = 16u * s2[threadIdx.x];
    // ---- chunks: [K/16][128 rows][16 B] ----
    uint4* out = A8 + (size_t)tile * nch * UM_ROWS;
    for (uint32_t idx = threadIdx.x; idx < nch * UM_ROWS; idx += blockDim.x) {
        const uint32_t row = idx % UM_ROWS, ch = idx / UM_ROWS, lr = row >> 2, k = row & 3u;
        uint32_t w[4] = {0, 0, 0, 0};
        if (lr < nvalid && ch * 16 < N) {
            const uint8_t* sb = sblk + lr * N;
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

// B: [col tile][K stage][plane lo,hi][kc/16][64 cols][16 B].  Thread -> (tile, stage, 16-byte K chunk, column): writes the
// low-plane and the high-plane 16 bytes.  coln (by padded position): sum D^2, INT_MAX for padding columns.
__global__ void k_build_pool_i8(const uint8_t* __restrict__ img, uint32_t stride, const fe_grid_item* __restrict__ dom,
                                const uint32_t* __restrict__ order, UmmaBuckets bk, uint32_t T, uint32_t Kpad, uint32_t kc,
                                const uint32_t* __restrict__ colS2, uint4* __restrict__ B8, uint32_t* __restrict__ coln_tiles,
                                uint32_t* __restrict__ blob_dom, uint32_t* __restrict__ tileseg) {
    const uint32_t nst = Kpad / kc, ncs = kc / 16;
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)bk.col_tile0[bk.nb] * nst * ncs * I8_NT) return;
    const uint32_t l = (uint32_t)(idx % I8_NT), cs = (uint32_t)((idx / I8_NT) % ncs), st = (uint32_t)((idx / ((uint64_t)I8_NT * ncs)) % nst);
    const uint32_t tile = (uint32_t)(idx / ((uint64_t)I8_NT * ncs * nst));
    int bi = 0;
    while (bi + 1 < bk.nb && tile >= bk.col_tile0[bi + 1]) ++bi;
    const uint32_t c = bk.dom_off[bi] + (tile - bk.col_tile0[bi]) * I8_NT + l;
    const uint32_t N = T * T;
    uint32_t lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0};
    const bool valid = c < bk.dom_end[bi];
    const uint32_t di = valid ? (order ? order[c] : c) : FE_NONE32;
    if (st == 0 && cs == 0) {
        blob_dom[(size_t)tile * I8_NT + l] = di;
        if (l == 0) tileseg[tile] = (uint32_t)bi;
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

int umma_i8_level_supported(const LevelGeom& g) { return g.fast && g.T >= 4 && g.T <= 32; }

static uint32_t i8_kpad(const LevelGeom& g) { return g.N <= (uint32_t)I8_KC ? ((g.N + 31u) & ~31u) : ((g.N + I8_KC - 1) / I8_KC) * I8_KC; }

int umma_i8_prepare_and_search(fe_ctx* ctx, const LevelGeom& g, const fe_grid_item* d_dom, const fe_grid_item* d_rng, const SearchPass& sp,
                               uint32_t thr16, bool use_thr) {
    const uint32_t Kpad = i8_kpad(g), kc = std::min(Kpad, (uint32_t)I8_KC);
    const int nbuckets = sp.nbuckets;
    const uint32_t* dom_order = sp.dom_order;
    const uint32_t* rng_order = sp.rng_items;
    UmmaBuckets bk{};
    UmmaArgs a{};
    uint32_t rt = 0, ct = 0, nb = 0;
    uint64_t total_items = 0;
    // B blob: the slices of the domain buckets, one after the other (whole column tiles each)
    for (int b = 0; b < nbuckets; ++b) {
        const uint32_t dc = sp.dend[b] - sp.dbeg[b];
        bk.dom_off[b] = sp.dbeg[b];
        bk.dom_end[b] = sp.dend[b];
        bk.col_tile0[b] = ct;
        ct += (dc + I8_NT - 1) / I8_NT;
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
    uint32_t live_row_tiles = 0;
    for (uint32_t i = 0; i < nb; ++i)
        if (a.b[i].n_col_tiles) live_row_tiles += a.b[i].n_row_tiles;
    const uint32_t want_chunks = live_row_tiles ? (2 * 148 + live_row_tiles - 1) / live_row_tiles : 1;
    for (uint32_t i = 0; i < nb; ++i) {
        UmmaBucket& b = a.b[i];
        b.chunks = (b.n_col_tiles && b.n_row_tiles) ? std::max(1u, std::min(want_chunks, b.n_col_tiles)) : 0;
        total_items += (uint64_t)b.n_row_tiles * b.chunks;
        a.item_end[i] = (uint32_t)std::min<uint64_t>(total_items, 0xFFFFFFFFull);
    }
    if (total_items == 0) {
        if (sp.ev0) { cudaEventRecord(sp.ev0, ctx->stream); cudaEventRecord(sp.ev1, ctx->stream); }
        return FE_OK;
    }
    if (total_items > 0x7FFFFFFFull) return fe_fail(ctx, FE_ERR_UNSUPPORTED, "umma i8: too many work items");

    const size_t bytesA = (size_t)rt * UM_ROWS * Kpad, bytesB = (size_t)ct * 2 * I8_NT * Kpad;
    FE_CUDA(ctx, ctx->b_A16.ensure(bytesA + 256));
    FE_CUDA(ctx, ctx->b_B16.ensure(bytesB + 256));
    FE_CUDA(ctx, ctx->b_coln.ensure((size_t)ct * I8_NT * 4 + 64));
    FE_CUDA(ctx, ctx->b_blob_dom.ensure((size_t)ct * I8_NT * 4 + 64));
    FE_CUDA(ctx, ctx->b_tileseg.ensure((size_t)ct * 4 + 64));
    FE_CUDA(ctx, ctx->b_rowc.ensure((size_t)bk.n_ranges * 4 + 4));
    FE_CUDA(ctx, ctx->b_tmaps.ensure((size_t)bk.n_domains * 4 + 64));
    if (!sp.reuse_rows) {
        k_build_rows_i8<<<rt, 256, 32 * g.N, ctx->stream>>>(ctx->tgt.px, ctx->tgt.stride, d_rng, rng_order, bk, g.T, Kpad, ctx->b_A16.as<uint4>(),
                                                           ctx->b_rowc.as<uint32_t>());
        ctx->stats.kernel_launches++;
    }
    if (!sp.reuse_dom_norms) {
        k_block_norms<<<(unsigned)(((uint64_t)bk.n_domains * 32 + 255) / 256), 256, 0, ctx->stream>>>(ctx->src.px, ctx->src.stride, d_dom, dom_order,
                                                                                                    bk.n_domains, g.T, 3, ctx->b_tmaps.as<uint32_t>());
        ctx->stats.kernel_launches++;
    }
    FE_CUDA(ctx, cudaGetLastError());
    k_build_pool_i8<<<(unsigned)((bytesB / 32 + 255) / 256), 256, 0, ctx->stream>>>(
        ctx->src.px, ctx->src.stride, d_dom, dom_order, bk, g.T, Kpad, kc, ctx->b_tmaps.as<uint32_t>(), ctx->b_B16.as<uint4>(), ctx->b_coln.as<uint32_t>(),
        ctx->b_blob_dom.as<uint32_t>(), ctx->b_tileseg.as<uint32_t>());
    FE_CUDA(ctx, cudaGetLastError());
    ctx->stats.kernel_launches++;

    a.A16 = ctx->b_A16.p;
    a.B16 = ctx->b_B16.p;
    a.colmeta = nullptr;
    a.tileseg = ctx->b_tileseg.as<uint32_t>();
    a.blob_dom = ctx->b_blob_dom.as<uint32_t>();
    a.coln = ctx->b_coln.as<uint32_t>();
    a.rowA2 = ctx->b_rowc.as<uint32_t>();
    a.rowbest = ctx->b_rowbest.as<unsigned long long>();
    a.rowhit = ctx->b_rowhit.as<uint32_t>();
    a.flags = ctx->b_counters.as<uint32_t>() + 2;
    a.nb = (int)nb;
    a.Kpad = Kpad;
    a.total_items = (uint32_t)total_items;
    a.thr16 = thr16;
    a.use_thr = use_thr ? 1u : 0u;
    a.nt = I8_NT;
    a.rowslot = sp.rowslot;
    a.meta = sp.span > 0 ? 1u : 0u;
    a.no_min = sp.no_min ? 1u : 0u;
    const uint32_t stage_bytes = 2 * I8_NT * kc, a_bytes = UM_ROWS * Kpad;
    const uint32_t budget = 226 * 1024 - 512;
    a.n_abuf = (2 * a_bytes + 2 * stage_bytes <= budget) ? 2 : 1;
    uint32_t stages = (budget - a.n_abuf * a_bytes) / stage_bytes;
    stages = std::min((uint32_t)I8_MAX_STAGES, stages);
    if (stages < 2) return fe_fail(ctx, FE_ERR_UNSUPPORTED, "umma i8: operands do not fit shared memory (T=%u)", g.T);
    a.stages = stages;
    const size_t smem = (size_t)a.n_abuf * a_bytes + (size_t)stages * stage_bytes + (12 + 2 * I8_MAX_STAGES) * 8 + 128;
    FE_CUDA(ctx, cudaFuncSetAttribute(k_search_umma_i8<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    const uint32_t grid = (uint32_t)std::min<uint64_t>(total_items, 148);
    if (sp.ev0) cudaEventRecord(sp.ev0, ctx->stream);
    k_search_umma_i8<0><<<grid, UM_THREADS_I8, smem, ctx->stream>>>(a);
    FE_CUDA(ctx, cudaGetLastError());
    if (sp.ev1) cudaEventRecord(sp.ev1, ctx->stream);
    ctx->stats.kernel_launches++;
    return FE_OK;
}
