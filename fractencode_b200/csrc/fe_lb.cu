// fe_lb.cu -- lower-bound prefilter for the large range blocks of a threshold search that only looks for hits.
//
// A level whose range blocks split when they find no candidate under the threshold (the quadtree levels above the last)
// never needs its minimum: all it has to establish per range block is the first domain under the threshold, or that there is
// none.  For T = 32 that proof costs K = 1024 products per candidate on the exact path -- but almost no candidate comes close:
// by Cauchy-Schwarz over the c x c cells of a block (c = T / 8)
//        sum_cells (A_cell - B_cell)^2  <=  c^2 * sum_pixels (4 r - D)^2  =  c^2 * n16,
// A_cell / B_cell the cell sums of 4 r / of the decimated domain D.  With a' = (s * sum_cell r) - 2040 and
// b' = (s / 4) * (sum of the 2c x 2c source pixels of the cell) - 2040  (s = 16 / c^2: 1 for c = 4, 2 for c = 2) this reads
//        sum_cells (a' - b')^2  <=  n16:
// an 8 x 8 problem of exactly the shape the kind::f16 kernel solves (rotating a block rotates its cell grid), on 12-bit
// operands.  b' is rounded to an integer b'' (|b'' - b'| <= 1/2 over 64 cells moves the norm by at most 4), the fp32
// accumulator is no longer exact at these magnitudes (partial sums < 2^29, bounded error), so the test is made conservative:
//        candidate  <=>  a2' + 2 V + p  <=  (sqrt(thr16) + 4)^2 + 2^16,
// which every true hit passes.  The kernel (k_search_f16 MODE 1) emits the candidates; k_lb_verify scores each exactly from
// the pixels and records real hits like the exact kernels do (atomicMin on the domain index).  On the benchmark image
// 5e-5 of the candidates of the T = 32 level survive the filter: the level costs 1/16 of the exact pass.
// If the candidate list overflows (smooth images: everything matches everything) the level is redone on the exact kind.
#include <cmath>

#include "fe_kernels.cuh"
#include "fe_plan.cuh"
#include "fe_umma_dev.cuh"

using namespace umma_dev;

namespace {

// Q[j][i] = sum of the c x c pixels at (c i, c j): the cell sums of every range block on the c-lattice; a domain cell is
// the sum of 2 x 2 of them.
__global__ void k_cellsum(const uint8_t* __restrict__ img, uint32_t stride, uint32_t qw, uint32_t qh, uint32_t c, uint16_t* __restrict__ Q) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= qw || j >= qh) return;
    uint32_t s = 0;
    for (uint32_t y = 0; y < c; ++y)
        for (uint32_t x = 0; x < c; ++x) s += img[(size_t)(c * j + y) * stride + c * i + x];
    Q[(size_t)j * qw + i] = (uint16_t)s;
}

// ListEntry.a2 of the prefilter rows: sum over the 64 cells of (s * Q - 2040)^2
__global__ void k_lb_norms(const uint16_t* __restrict__ Q, uint32_t qw, uint32_t c, uint32_t s, uint32_t n, ListEntry* __restrict__ list0) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t xy = list0[p].xy, qx = (xy & 0xFFFFu) / c, qy = (xy >> 16) / c;
    uint32_t a2 = 0;
    for (uint32_t y = 0; y < 8; ++y)
        for (uint32_t x = 0; x < 8; ++x) {
            const int a = (int)(s * Q[(size_t)(qy + y) * qw + qx + x]) - 2040;
            a2 += (uint32_t)(a * a);
        }
    list0[p].a2 = a2;
}

// B blob of the level: one CTA per blob tile, thread = column (as k_build_pool16_level<8>, on cell sums).
// b'' = round(s * Dq / 4) - 2040, Dq = 2 x 2 box sum of Q.
__global__ void __launch_bounds__(128) k_build_pool_lb(const uint16_t* __restrict__ Q, uint32_t qw, uint32_t c, uint32_t s,
                                                       const fe_grid_item* __restrict__ dom, const uint32_t* __restrict__ order,
                                                       const LevelPlan* __restrict__ plan, uint4* __restrict__ B16, uint32_t* __restrict__ colmeta,
                                                       uint32_t* __restrict__ blob_dom) {
    constexpr int T = 8, N = 64, NCH = 10;
    __shared__ uint32_t s_chunk, s_col0, s_end, s_doff;
    const uint32_t tile = blockIdx.x, l = threadIdx.x;
    if (tile >= plan->n_tiles) return;
    if (l == 0) {
        const uint32_t nb = plan->nb;
        uint32_t lo = 0, hi = plan->nk * nb - 1;
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
        const uint16_t* base = Q + (size_t)(d.y / c) * qw + d.x / c;     // 16 x 16 cell sums of the domain block
        float v[8];
#pragma unroll
        for (int Y = 0; Y < T; ++Y) {
            const uint16_t* r0 = base + (size_t)(2 * Y) * qw;
            const uint16_t* r1 = r0 + qw;
#pragma unroll
            for (int X = 0; X < T; ++X) {
                const uint32_t dq = (uint32_t)r0[2 * X] + r0[2 * X + 1] + r1[2 * X] + r1[2 * X + 1];
                const int b = (int)((s * dq + 2u) >> 2) - 2040;
                s2 += (uint32_t)(b * b);
                v[X] = (float)b;
            }
            out[Y * UM_NT] = make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]), pack_half2(v[6], v[7]));
        }
        const uint32_t h = s2 >> 1;   // < 2^27: limbs h0, h1 < 2048, h2 < 32 (2048 * h2 is exact in fp16)
        out[(N / 8) * UM_NT] = make_uint4(pack_half2((float)(h & 2047u), (float)((h >> 11) & 2047u)), pack_half2((float)((h >> 22) * 2048u), 0.f), 0, 0);
        out[(N / 8 + 1) * UM_NT] = make_uint4(0, 0, 0, 0);
    } else {
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) out[ch * UM_NT] = make_uint4(0, 0, 0, 0);
    }
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

// A blob of the slice: one CTA per row tile, thread = row (range block row >> 2 of the tile under the inverse of rotation
// row & 3): values 2040 - s * Q, then the constant columns [1, 2048, 2048].
__global__ void __launch_bounds__(128) k_build_rows_lb(const uint16_t* __restrict__ Q, uint32_t qw, uint32_t c, uint32_t s,
                                                       const LevelPlan* __restrict__ plan, const SliceCtl* __restrict__ ctl,
                                                       const ListEntry* __restrict__ list0, const ListEntry* __restrict__ list1, uint32_t ordinal,
                                                       uint4* __restrict__ A16) {
    constexpr int T = 8, N = 64, NCH = 10;
    __shared__ uint32_t s_pos0, s_nvalid;
    if (ctl->active != ordinal) return;
    const ListEntry* list = ctl->list ? list1 : list0;
    for (uint32_t tile = blockIdx.x; tile < ctl->n_row_tiles; tile += gridDim.x) {
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t lo = 0, hi = plan->nb - 1;
            while (lo < hi) {
                const uint32_t mid = (lo + hi + 1) >> 1;
                if (ctl->tile_prefix[mid] <= tile) lo = mid; else hi = mid - 1;
            }
            const uint32_t rt = tile - ctl->tile_prefix[lo];
            s_pos0 = plan->roff[lo] + 32 * rt;
            s_nvalid = min(32u, ctl->cnt[ctl->list][lo] - 32 * rt);
        }
        __syncthreads();
        const uint32_t row = threadIdx.x, lr = row >> 2, k = row & 3u;
        uint4* out = A16 + (size_t)tile * NCH * UM_ROWS + row;
        if (lr >= s_nvalid) {
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) out[ch * UM_ROWS] = make_uint4(0, 0, 0, 0);
            continue;
        }
        const ListEntry e = list[s_pos0 + lr];
        const uint16_t* base = Q + (size_t)((e.xy >> 16) / c) * qw + (e.xy & 0xFFFFu) / c;
#pragma unroll
        for (int Y = 0; Y < T; ++Y) {
            float v[8];
#pragma unroll
            for (int X = 0; X < T; ++X) {
                int py = k == 0 ? Y : k == 1 ? X : k == 2 ? T - 1 - Y : T - 1 - X;
                int px = k == 0 ? X : k == 1 ? T - 1 - Y : k == 2 ? T - 1 - X : Y;
                v[X] = (float)(2040 - (int)(s * base[(size_t)py * qw + px]));
            }
            out[Y * UM_ROWS] = make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]), pack_half2(v[6], v[7]));
        }
        out[(N / 8) * UM_ROWS] = make_uint4(0x68003C00u, 0x00006800u, 0, 0);   // [1, 2048 | 2048, 0 | 0 ...]
        out[(N / 8 + 1) * UM_ROWS] = make_uint4(0, 0, 0, 0);
    }
}

// Exact score of every candidate from the pixels (one warp each; lane = row of the decimated domain block, whose two source
// rows are read as 16-byte vectors); real hits are recorded like the exact kernels do.  cand_total += candidates of the slice.
__global__ void k_lb_verify(const uint8_t* __restrict__ src, uint32_t src_stride, const uint8_t* __restrict__ tgt, uint32_t tgt_stride,
                            const fe_grid_item* __restrict__ dom, const fe_grid_item* __restrict__ rng, const uint32_t* __restrict__ rng_order,
                            const uint2* __restrict__ cand, const uint32_t* __restrict__ cand_count, uint32_t cap, uint32_t thr16,
                            const SliceCtl* __restrict__ ctl, uint32_t ordinal, uint32_t* __restrict__ rowhit, uint32_t* __restrict__ flags,
                            unsigned long long* __restrict__ cand_total) {
    if (ctl->active != ordinal) return;
    const uint32_t n = *cand_count;
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(cand_total, (unsigned long long)n);
    if (n > cap) {                           // the list overflowed: the level is redone on the exact kind
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(flags, 2u);
        return;
    }
    const uint32_t lane = threadIdx.x & 31, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += nw) {
        const uint2 cd = cand[i];
        const uint32_t srow = cd.x, slot = srow >> 2, k = srow & 3u, d = cd.y;
        if (__ldcg(&rowhit[srow]) <= d) continue;                 // an earlier domain is already a hit for this row
        const fe_grid_item r = rng[rng_order ? rng_order[slot] : slot];
        const fe_grid_item dm = dom[d];
        const uint32_t T = r.w;
        const bool vec = ((dm.x | src_stride | (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u)) & 15u) == 0 && (T & 7u) == 0;
        uint32_t sA2 = 0, sAB = 0, sB2 = 0;
        for (uint32_t ty = lane; ty < T; ty += 32) {
            const uint8_t* q0 = src + (size_t)(dm.y + 2 * ty) * src_stride + dm.x;
            for (uint32_t t0 = 0; t0 < T; t0 += 8) {              // eight box sums = 16 source pixels of two rows
                uint32_t D[8];
                if (vec) {
                    const uint4 a = __ldg(reinterpret_cast<const uint4*>(q0 + 2 * t0)), b = __ldg(reinterpret_cast<const uint4*>(q0 + src_stride + 2 * t0));
                    const uint32_t ta[4] = {a.x, a.y, a.z, a.w}, tb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        const uint32_t dd = (ta[w] & 0x00FF00FFu) + ((ta[w] >> 8) & 0x00FF00FFu) + (tb[w] & 0x00FF00FFu) + ((tb[w] >> 8) & 0x00FF00FFu);
                        D[2 * w] = dd & 0xFFFFu; D[2 * w + 1] = dd >> 16;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint8_t* q = q0 + 2 * (t0 + j);
                        D[j] = (t0 + j < T) ? (uint32_t)q[0] + q[1] + q[src_stride] + q[src_stride + 1] : 0u;
                    }
                }
                // row k = the range block under the INVERSE of rotation k against the unrotated domain.  k even: the eight pixels
                // lie in ONE row of the range block (the lanes of the warp read 32 different rows: byte loads would be 32
                // sectors per instruction) -- one 8-byte load, read backwards for k = 2; k odd: a column, which the lanes of the
                // warp read as one row segment
                if (vec && (k & 1u) == 0 && ((r.x | tgt_stride | (uint32_t)(reinterpret_cast<uintptr_t>(tgt) & 7u)) & 7u) == 0) {
                    const uint32_t py = k == 0 ? ty : T - 1 - ty, px0 = k == 0 ? t0 : T - 8 - t0;
                    const uint2 w = __ldg(reinterpret_cast<const uint2*>(tgt + (size_t)(r.y + py) * tgt_stride + r.x + px0));
                    const uint32_t lo = k == 0 ? w.x : __byte_perm(w.y, 0, 0x0123), hi = k == 0 ? w.y : __byte_perm(w.x, 0, 0x0123);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint32_t a = ((j < 4 ? lo : hi) >> (8 * (j & 3))) & 255u;
                        sA2 += a * a; sAB += a * D[j]; sB2 += D[j] * D[j];
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint32_t tx = t0 + j;
                        if (tx < T) {
                            const uint32_t py = k == 0 ? ty : k == 1 ? tx : k == 2 ? T - 1 - ty : T - 1 - tx;
                            const uint32_t px = k == 0 ? tx : k == 1 ? T - 1 - ty : k == 2 ? T - 1 - tx : ty;
                            const uint32_t a = tgt[(size_t)(r.y + py) * tgt_stride + r.x + px];
                            sA2 += a * a; sAB += a * D[j]; sB2 += D[j] * D[j];
                        }
                    }
                }
            }
        }
        for (int o = 16; o; o >>= 1) {
            sA2 += __shfl_xor_sync(0xFFFFFFFFu, sA2, o);
            sAB += __shfl_xor_sync(0xFFFFFFFFu, sAB, o);
            sB2 += __shfl_xor_sync(0xFFFFFFFFu, sB2, o);
        }
        const uint32_t n16 = 16u * sA2 - 8u * sAB + sB2;          // exact for T <= 64
        if (lane == 0 && n16 <= thr16) atomicMin(&rowhit[srow], d);
    }
}

} // namespace

uint32_t lb_threshold(uint32_t thr16) {
    const double t = std::sqrt((double)thr16) + 4.0;
    const double v = std::ceil(t * t) + 65536.0;              // rounding of b'' (norm + 4), accumulator error (2^16)
    return v >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)v;
}

int lb_prepare(fe_ctx* ctx, const LevelGeom& g, const fe_grid_item* d_dom, const uint32_t* dom_order, const LevelPlan* plan, uint32_t nR,
               ListEntry* list0, uint32_t max_tiles, LbState* lb) {
    const uint32_t c = g.T / 8, s = 16 / (c * c);
    const uint32_t qw = ctx->tgt.w / c, qh = ctx->tgt.h / c;
    lb->c = c; lb->s = s; lb->qw = qw;
    FE_CUDA(ctx, ctx->b_lbq.ensure((size_t)qw * qh * 2 + 64));
    lb->Q = ctx->b_lbq.as<uint16_t>();
    dim3 block(32, 8), grid((qw + 31) / 32, (qh + 7) / 8);
    k_cellsum<<<grid, block, 0, ctx->stream>>>(ctx->tgt.px, ctx->tgt.stride, qw, qh, c, ctx->b_lbq.as<uint16_t>());
    k_lb_norms<<<(nR + 127) / 128, 128, 0, ctx->stream>>>(lb->Q, qw, c, s, nR, list0);
    const uint32_t Kpad = 80;
    FE_CUDA(ctx, ctx->b_B16.ensure((size_t)max_tiles * UM_NT * Kpad * 2 + 256));
    FE_CUDA(ctx, ctx->b_tmaps.ensure((size_t)max_tiles * 32 + 64));
    FE_CUDA(ctx, ctx->b_blob_dom.ensure((size_t)max_tiles * UM_NT * 4 + 64));
    k_build_pool_lb<<<max_tiles, 128, 0, ctx->stream>>>(lb->Q, qw, c, s, d_dom, dom_order, plan, ctx->b_B16.as<uint4>(), ctx->b_tmaps.as<uint32_t>(),
                                                       ctx->b_blob_dom.as<uint32_t>());
    FE_CUDA(ctx, cudaGetLastError());
    ctx->stats.kernel_launches += 3;
    lb->cand_cap = 4u << 20;
    FE_CUDA(ctx, ctx->b_lbcand.ensure((size_t)lb->cand_cap * sizeof(uint2) + 64));
    lb->cand = ctx->b_lbcand.as<uint2>();
    lb->cand_count = reinterpret_cast<uint32_t*>(ctx->b_lbcand.as<uint8_t>() + (size_t)lb->cand_cap * sizeof(uint2));
    return FE_OK;
}

int lb_build_rows(fe_ctx* ctx, const LbState& lb, const LevelPlan* plan, const SliceCtl* ctl, const ListEntry* const list[2], uint32_t ordinal,
                  uint32_t max_row_tiles) {
    FE_CUDA(ctx, ctx->b_A16.ensure((size_t)max_row_tiles * UM_ROWS * 80 * 2 + 256));
    FE_CUDA(ctx, cudaMemsetAsync(lb.cand_count, 0, 4, ctx->stream));
    k_build_rows_lb<<<std::min(max_row_tiles, 4u * (uint32_t)ctx->n_sm), 128, 0, ctx->stream>>>(lb.Q, lb.qw, lb.c, lb.s, plan, ctl, list[0], list[1], ordinal,
                                                                                               ctx->b_A16.as<uint4>());
    FE_CUDA(ctx, cudaGetLastError());
    ctx->stats.kernel_launches++;
    return FE_OK;
}

int lb_verify(fe_ctx* ctx, const LbState& lb, const fe_grid_item* d_dom, const fe_grid_item* d_rng, const uint32_t* rng_order, uint32_t thr16,
              const SliceCtl* ctl, uint32_t ordinal) {
    k_lb_verify<<<16 * ctx->n_sm, 256, 0, ctx->stream>>>(ctx->src.px, ctx->src.stride, ctx->tgt.px, ctx->tgt.stride, d_dom, d_rng, rng_order, lb.cand,
                                                       lb.cand_count, lb.cand_cap, thr16, ctl, ordinal, ctx->b_rowhit.as<uint32_t>(),
                                                       ctx->b_counters.as<uint32_t>() + 2, reinterpret_cast<unsigned long long*>(ctx->b_counters.as<uint32_t>() + 4));
    FE_CUDA(ctx, cudaGetLastError());
    ctx->stats.kernel_launches++;
    return FE_OK;
}
