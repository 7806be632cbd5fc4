// fe_kernels.cu -- CUDA kernels of the search path except the tcgen05 contraction
// (fe_search_f16.cu, fe_search_i8.cu): grids, classification, cell sums, operand preparation, the exact integer
// search, winner finalisation (least squares s/o), quadtree bookkeeping, decode gather,
// quantizer and the synthetic image generator.  sm_100a only.
//
// Reference semantics restated here (sebsgit/fractencode, file:line):
//   sampling        image/sampler.h:22-38, image/transform.h:32-41,96-109
//   distance        image/metrics.h:37-50  (fp32 running sum of exact 1/16 multiples)
//   classifier      encode/Classifier2.cpp:8-81
//   s / o           encode/transformmatcher.h:98-108
//   decode          encode/DecodeUtils.hpp:9-25, encode/Encoder2.hpp:67-99
//   quantizer       encode/Quantizer.hpp:13-36
#include "fe_kernels.cuh"

__device__ __constant__ int8_t kMapDev[8][8] = {
    {1, 0, 0, 0, 0, 1, 0, 0},   {0, 1, 0, 0, -1, 0, 1, 0}, {-1, 0, 1, 0, 0, -1, 0, 1}, {0, -1, 0, 1, 1, 0, 0, 0},
    {1, 0, 0, 0, 0, -1, 0, 1},  {0, 1, 0, 0, 1, 0, 0, 0},  {-1, 0, 1, 0, 0, 1, 0, 0},  {0, -1, 0, 1, -1, 0, 1, 0},
};

// 2x2 box SUM (= 4 * SamplerBilinear::sample) at local (lx,ly) of patch (px,py,ps x ps) under isometry t.
__device__ __forceinline__ int sample_sum4(const uint8_t* __restrict__ img, uint32_t stride, uint32_t px, uint32_t py,
                                           uint32_t ps, uint32_t lx, uint32_t ly, int t) {
    if (lx == ps - 1) --lx; // sampler.h:32-35
    if (ly == ps - 1) --ly;
    const int m0 = kMapDev[t][0], m1 = kMapDev[t][1], m4 = kMapDev[t][4], m5 = kMapDev[t][5];
    const int e = (int)ps - 1;
    const int gx = (int)px + m0 * (int)lx + m1 * (int)ly + (kMapDev[t][2] + kMapDev[t][3]) * e;
    const int gy = (int)py + m4 * (int)lx + m5 * (int)ly + (kMapDev[t][6] + kMapDev[t][7]) * e;
    const uint8_t* p = img + (size_t)gy * stride + gx;
    const int s = (int)stride;
    return p[0] + p[m4 * s + m0] + p[m5 * s + m1] + p[(m4 + m5) * s + m0 + m1];
}

// BrightnessBlocksClassifier2::getCategory (Classifier2.cpp:8-53): class of the strict ordering of the
// four quadrant sums per the literal table; -1 when any two are equal or for the one ordering the
// table misses (see below).
__device__ __forceinline__ int category4(uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4) {
    if (a1 == a2 || a1 == a3 || a1 == a4 || a2 == a3 || a2 == a4 || a3 == a4) return -1;
    // rank of each quadrant in descending order (0 = largest)
    const int r1 = (a2 > a1) + (a3 > a1) + (a4 > a1);
    const int r2 = (a1 > a2) + (a3 > a2) + (a4 > a2);
    const int r3 = (a1 > a3) + (a2 > a3) + (a4 > a3);
    const int r4 = (a1 > a4) + (a2 > a4) + (a3 > a4);
    // descending order as digits: ord[rank] = quadrant id (1..4)
    int ord[4];
    ord[r1] = 1;
    ord[r2] = 2;
    ord[r3] = 3;
    ord[r4] = 4;
    const int code = ord[0] * 1000 + ord[1] * 100 + ord[2] * 10 + ord[3];
    switch (code) {
    case 1234: case 3142: case 4321: case 2413: return 0;
    case 1324: case 2143: case 4231: case 3412: return 1;
    case 1432: case 4123: case 3241: case 2314: return 2;
    case 1243: case 3124: case 4312: case 2431: return 3;
    case 2134: case 1342: case 3421: case 4213: return 4;
    case 1423: case 2341: case 3214: return 5;
    // 4132 (a4 > a1 > a3 > a2) has no row: the reference's fourth class-5 test reads
    // `a4a1 && a1a3 && a3a4` (Classifier2.cpp:48), which can never hold, so that ordering is -1.
    }
    return -1;
}

__device__ __forceinline__ uint32_t block_sum_dev(const uint8_t* __restrict__ img, uint32_t stride, uint32_t x, uint32_t y,
                                                  uint32_t w, uint32_t h) {
    uint32_t s = 0;
    for (uint32_t j = 0; j < h; ++j) {
        const uint8_t* row = img + (size_t)(y + j) * stride + x;
        for (uint32_t i = 0; i < w; ++i) s += row[i];
    }
    if (w <= 16) s &= 0xFFFFu; // ImageStatistics2::sum -> sum_u16 (uint16_t) for widths <= 16
    return s;
}

// ---------------------------------------------------------------------------------------------
// grids (image/partition2.hpp:110-135) and quadtree children (partition2.hpp:18-30)
// ---------------------------------------------------------------------------------------------
__global__ void k_uniform_grid(fe_grid_item* out, uint32_t nx, uint32_t n, uint32_t size, uint32_t step, uint32_t first) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fe_grid_item it;
    it.x = ((first + i) % nx) * step;
    it.y = ((first + i) / nx) * step;
    it.w = size;
    it.h = size;
    it.bin = -1;
    out[i] = it;
}

// After an exclusive scan of the split flags: children of split block i go to next[4*scan[i] ..+3]
// (topLeft, topRight, bottomLeft, bottomRight); kept blocks are copied to items_out[out_base + (i - scan[i])].
__global__ void k_quadtree_scatter(const fe_grid_item* __restrict__ rng, const fe_encode_item* __restrict__ level_items,
                                   const uint32_t* __restrict__ split, const uint32_t* __restrict__ scan, uint32_t n,
                                   fe_grid_item* __restrict__ next, fe_encode_item* __restrict__ items_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = scan[i];
    if (split[i]) {
        const fe_grid_item r = rng[i];
        const uint32_t h = r.w / 2;
        fe_grid_item c;
        c.w = h;
        c.h = h;
        c.bin = -1;
        c.x = r.x; c.y = r.y; next[4 * s + 0] = c;
        c.x = r.x + h; c.y = r.y; next[4 * s + 1] = c;
        c.x = r.x; c.y = r.y + h; next[4 * s + 2] = c;
        c.x = r.x + h; c.y = r.y + h; next[4 * s + 3] = c;
    } else {
        items_out[i - s] = level_items[i];
    }
}

// ---------------------------------------------------------------------------------------------
// classification
// ---------------------------------------------------------------------------------------------
// One warp per item: quadrant sums by lane-strided rows; class kept as given when bin != -1
// (Classifier2::compare only recomputes for -1, Classifier2.cpp:70-81).
__global__ void k_classify(const uint8_t* __restrict__ img, uint32_t stride, const fe_grid_item* __restrict__ items, uint32_t n,
                           int32_t* __restrict__ cls, int force) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n) return;
    const fe_grid_item it = items[warp];
    if (!force && it.bin != -1) {
        if (lane == 0) cls[warp] = it.bin;
        return;
    }
    const uint32_t hw = it.w / 2, hh = it.h / 2;
    uint32_t a[4] = {0, 0, 0, 0};
    // quadrant q pixel p: lanes stride over the hw*hh pixels of each quadrant
    for (int q = 0; q < 4; ++q) {
        const uint32_t qx = it.x + (q & 1) * hw, qy = it.y + (q >> 1) * hh;
        uint32_t s = 0;
        for (uint32_t p = lane; p < hw * hh; p += 32) s += img[(size_t)(qy + p / hw) * stride + qx + p % hw];
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
        if (hw <= 16) s &= 0xFFFFu;
        a[q] = s;
    }
    if (lane == 0) cls[warp] = category4(a[0], a[1], a[2], a[3]);
}

// The same classes for a list of square blocks of one power-of-two edge (every list of a search level): min(EDGE, 32) lanes per
// block instead of a warp, a lane sums the left and the right half of its rows with 4-byte loads.  A block of another size or at
// an unaligned origin is summed pixel by pixel by the group's first lane (same result, the slow way).
template <int EDGE>
__global__ void __launch_bounds__(256) k_classify_edge(const uint8_t* __restrict__ img, uint32_t stride, const fe_grid_item* __restrict__ items, uint32_t n,
                                                       int32_t* __restrict__ cls, int force) {
    constexpr int L = EDGE < 32 ? EDGE : 32, RPL = EDGE / L, HALF = EDGE / 2;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, item = t / L, lane = t % L;
    const bool live = item < n;
    fe_grid_item it{};
    if (live) it = items[item];
    const bool keep = live && !force && it.bin != -1;
    const bool regular = it.w == EDGE && it.h == EDGE && (((uint32_t)it.x | stride | (uint32_t)(reinterpret_cast<uintptr_t>(img) & 3u)) & 3u) == 0;
    uint32_t a[4] = {0, 0, 0, 0};
    if (live && !keep) {
        if (regular) {
#pragma unroll
            for (int k = 0; k < RPL; ++k) {
                const int r = (int)lane + k * L;
                const uint32_t* row = reinterpret_cast<const uint32_t*>(img + (size_t)(it.y + r) * stride + it.x);
                uint32_t l = 0, rr = 0;
                if (EDGE == 4) {
                    const uint32_t v = __ldg(row);
                    l = __dp4a(v, 0x00000101u, 0u); rr = __dp4a(v, 0x01010000u, 0u);
                } else {
#pragma unroll
                    for (int w = 0; w < EDGE / 8; ++w) {
                        l = __dp4a(__ldg(row + w), 0x01010101u, l);
                        rr = __dp4a(__ldg(row + EDGE / 8 + w), 0x01010101u, rr);
                    }
                }
                if (r < HALF) { a[0] += l; a[1] += rr; } else { a[2] += l; a[3] += rr; }
            }
        } else if (lane == 0) {
            const uint32_t hw = it.w / 2, hh = it.h / 2;
            for (int q = 0; q < 4; ++q) {
                const uint32_t qx = it.x + (q & 1) * hw, qy = it.y + (q >> 1) * hh;
                uint32_t s = 0;
                for (uint32_t p = 0; p < hw * hh; ++p) s += img[(size_t)(qy + p / hw) * stride + qx + p % hw];
                a[q] = hw <= 16 ? (s & 0xFFFFu) : s;
            }
        }
    }
#pragma unroll
    for (int o = L / 2; o; o >>= 1) {
#pragma unroll
        for (int q = 0; q < 4; ++q) a[q] += __shfl_xor_sync(0xFFFFFFFFu, a[q], o);
    }
    if (live && lane == 0) cls[item] = keep ? it.bin : category4(a[0], a[1], a[2], a[3]);
}

// classes of a list whose blocks are (expected to be) edge x edge
void launch_classify(cudaStream_t stream, const uint8_t* img, uint32_t stride, const fe_grid_item* items, uint32_t n, uint32_t edge, int32_t* cls, int force) {
    if (!n) return;
    const uint64_t lanes = (uint64_t)n * (edge < 32 ? edge : 32);
    const unsigned grid = (unsigned)((lanes + 255) / 256);
    switch (edge) {
    case 4: k_classify_edge<4><<<grid, 256, 0, stream>>>(img, stride, items, n, cls, force); break;
    case 8: k_classify_edge<8><<<grid, 256, 0, stream>>>(img, stride, items, n, cls, force); break;
    case 16: k_classify_edge<16><<<grid, 256, 0, stream>>>(img, stride, items, n, cls, force); break;
    case 32: k_classify_edge<32><<<grid, 256, 0, stream>>>(img, stride, items, n, cls, force); break;
    case 64: k_classify_edge<64><<<grid, 256, 0, stream>>>(img, stride, items, n, cls, force); break;
    case 128: k_classify_edge<128><<<grid, 256, 0, stream>>>(img, stride, items, n, cls, force); break;
    default: k_classify<<<(unsigned)(((uint64_t)n * 32 + 255) / 256), 256, 0, stream>>>(img, stride, items, n, cls, force);
    }
}

__global__ void k_fill_u32(uint32_t* p, uint32_t v, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void k_fill_u64(unsigned long long* p, unsigned long long v, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
// Brightness bin of every block of a uniform list: key = (mul * sum of the block's edge x edge pixels) / width.  For a range
// block mul = 4 (sum of 4 r); for a domain block mul = 1 and edge = S (the sum of its 2x2 box sums D is the sum of its pixels).
// hist[key] counts.  WARP = true: one warp per block (large blocks); false: one thread per block (edge <= 16; neighbouring
// threads read neighbouring, for domains overlapping, blocks).
template <bool WARP>
__global__ void k_brightness_bins_t(const uint8_t* __restrict__ img, uint32_t stride, const fe_grid_item* __restrict__ items, uint32_t n,
                                    uint32_t edge, uint32_t mul, uint32_t width, uint8_t* __restrict__ keys, uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[FE_MAX_BUCKETS];
    if (threadIdx.x < FE_MAX_BUCKETS) sh[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t p = WARP ? t >> 5 : t, lane = WARP ? threadIdx.x & 31 : 0, step = WARP ? 32 : 1;
    if (p < n) {
        const fe_grid_item it = items[p];
        const uint8_t* base = img + (size_t)it.y * stride + it.x;
        uint32_t s = 0;
        if ((edge & 3u) == 0 && ((reinterpret_cast<uintptr_t>(base) | stride) & 3u) == 0) {
            const uint32_t wpr = edge / 4, nw = wpr * edge;   // 4-byte words per row / per block
            for (uint32_t e = lane; e < nw; e += step) {
                const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(base + (size_t)(e / wpr) * stride) + (e % wpr));
                s += __dp4a(v, 0x01010101u, 0u);
            }
        } else {
            for (uint32_t e = lane; e < edge * edge; e += step) s += base[(size_t)(e / edge) * stride + (e % edge)];
        }
        if (WARP)
            for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
        if (lane == 0) {
            const uint32_t k = min((mul * s) / width, (uint32_t)FE_MAX_BUCKETS - 1);
            keys[p] = (uint8_t)k;
            atomicAdd(&sh[k], 1u);
        }
    }
    __syncthreads();
    if (threadIdx.x < FE_MAX_BUCKETS && sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}

void launch_brightness_bins(cudaStream_t stream, const uint8_t* img, uint32_t stride, const fe_grid_item* items, uint32_t n, uint32_t edge,
                            uint32_t mul, uint32_t width, uint8_t* keys, uint32_t* hist) {
    if (!n) return;
    if (edge <= 16) k_brightness_bins_t<false><<<(n + 127) / 128, 128, 0, stream>>>(img, stride, items, n, edge, mul, width, keys, hist);
    else k_brightness_bins_t<true><<<(unsigned)(((uint64_t)n * 32 + 255) / 256), 256, 0, stream>>>(img, stride, items, n, edge, mul, width, keys, hist);
}

// ---------------------------------------------------------------------------------------------
// the domain lattice of a quadtree level through cell sums
// ---------------------------------------------------------------------------------------------
// On the quadtree's lattice a domain block (origin (T i, T j), edge 2T) is 2 x 2 cells of T x T pixels: one coalesced pass over
// the image gives every cell sum, and both the Classifier2 class (quadrant sums = the four cells) and the brightness bin (their
// total) of every domain come from four 4-byte reads -- instead of every domain re-reading its own (overlapping) pixels.
// cells[j][i] = sum of the C x C pixels at (C i, C j), C a power of two in 4..64: a thread adds up a 4-pixel wide column of a cell
// row, C / 4 neighbouring lanes combine.  cells2 = the same for the squared pixels (a range block on the lattice IS a cell: its
// operand norm comes from the two sums, k_level_ranges).
__global__ void __launch_bounds__(256) k_cell_grid(const uint8_t* __restrict__ img, uint32_t stride, uint32_t wpr, uint32_t ch, uint32_t C,
                                                   uint32_t* __restrict__ cells, uint32_t* __restrict__ cells2, uint32_t* __restrict__ cellsD2) {
    const uint32_t wx = blockIdx.x * 32 + threadIdx.x, j = blockIdx.y * 8 + threadIdx.y, lanes = C / 4;
    const bool live = wx < wpr && j < ch;
    uint32_t s = 0, s2 = 0, d2 = 0;
    if (live) {
        const uint8_t* p = img + (size_t)(C * j) * stride + 4 * (size_t)wx;
        for (uint32_t y = 0; y < C; y += 2) {
            const uint32_t w0 = __ldg(reinterpret_cast<const uint32_t*>(p + (size_t)y * stride));
            const uint32_t w1 = __ldg(reinterpret_cast<const uint32_t*>(p + (size_t)(y + 1) * stride));
            s = __dp4a(w1, 0x01010101u, __dp4a(w0, 0x01010101u, s));
            s2 = __dp4a(w1, w1, __dp4a(w0, w0, s2));
            const uint32_t Da = __dp4a(w1, 0x00000101u, __dp4a(w0, 0x00000101u, 0u));   // the two 2 x 2 box sums of the row pair
            const uint32_t Db = __dp4a(w1, 0x01010000u, __dp4a(w0, 0x01010000u, 0u));
            d2 += Da * Da + Db * Db;
        }
    }
    for (uint32_t o = lanes / 2; o; o >>= 1) {
        s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
        s2 += __shfl_xor_sync(0xFFFFFFFFu, s2, o);
        d2 += __shfl_xor_sync(0xFFFFFFFFu, d2, o);
    }
    if (live && (wx & (lanes - 1)) == 0) {
        const size_t c = (size_t)j * (wpr / lanes) + wx / lanes;
        cells[c] = s;
        cells2[c] = s2;
        cellsD2[c] = d2;
    }
}
// sum D^2 of the lattice domains in sorted order (what k_block_norms mode 3 computes from the pixels)
__global__ void k_dom_norms_from_cells(const uint32_t* __restrict__ cellsD2, uint32_t cw, uint32_t dnx, const uint32_t* __restrict__ order, uint32_t n,
                                       uint32_t* __restrict__ out) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t d = order ? order[p] : p;
    const uint32_t* c = cellsD2 + (size_t)(d / dnx) * cw + d % dnx;
    out[p] = __ldg(c) + __ldg(c + 1) + __ldg(c + cw) + __ldg(c + cw + 1);
}
void launch_dom_norms_from_cells(cudaStream_t stream, const uint32_t* cellsD2, uint32_t cw, uint32_t dnx, const uint32_t* order, uint32_t n, uint32_t* out) {
    if (n) k_dom_norms_from_cells<<<(n + 255) / 256, 256, 0, stream>>>(cellsD2, cw, dnx, order, n, out);
}
// domain d = (d % dnx, d / dnx) on the lattice (k_uniform_grid order); cls and/or keys + hist, whichever is asked for
__global__ void __launch_bounds__(256) k_dom_from_cells(const uint32_t* __restrict__ cells, uint32_t cw, uint32_t dnx, uint32_t n, int32_t* __restrict__ cls,
                                                        uint32_t width, uint8_t* __restrict__ keys, uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[FE_MAX_BUCKETS];
    if (keys) {
        for (uint32_t b = threadIdx.x; b < FE_MAX_BUCKETS; b += blockDim.x) sh[b] = 0;
        __syncthreads();
    }
    const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d < n) {
        const uint32_t i = d % dnx, j = d / dnx;
        const uint32_t* c = cells + (size_t)j * cw + i;
        const uint32_t a1 = __ldg(c), a2 = __ldg(c + 1), a3 = __ldg(c + cw), a4 = __ldg(c + cw + 1);
        if (cls) cls[d] = category4(a1, a2, a3, a4);
        if (keys) {
            const uint32_t k = min((a1 + a2 + a3 + a4) / width, (uint32_t)FE_MAX_BUCKETS - 1);
            keys[d] = (uint8_t)k;
            atomicAdd(&sh[k], 1u);
        }
    }
    if (keys) {
        __syncthreads();
        for (uint32_t b = threadIdx.x; b < FE_MAX_BUCKETS; b += blockDim.x)
            if (sh[b]) atomicAdd(&hist[b], sh[b]);
    }
}
bool cell_grid_supported(const uint8_t* img, uint32_t stride, uint32_t w, uint32_t h, uint32_t C) {
    return C >= 4 && C <= 64 && (C & (C - 1)) == 0 && w % C == 0 && h % C == 0 && stride % 4 == 0 && (reinterpret_cast<uintptr_t>(img) & 3u) == 0;
}
void launch_cell_grid(cudaStream_t stream, const uint8_t* img, uint32_t stride, uint32_t w, uint32_t h, uint32_t C, uint32_t* cells, uint32_t* cells2,
                      uint32_t* cellsD2) {
    const uint32_t wpr = w / 4, ch = h / C;
    k_cell_grid<<<dim3((wpr + 31) / 32, (ch + 7) / 8), dim3(32, 8), 0, stream>>>(img, stride, wpr, ch, C, cells, cells2, cellsD2);
}
void launch_dom_from_cells(cudaStream_t stream, const uint32_t* cells, uint32_t cw, uint32_t dnx, uint32_t n, int32_t* cls, uint32_t width, uint8_t* keys,
                           uint32_t* hist) {
    if (n) k_dom_from_cells<<<(n + 255) / 256, 256, 0, stream>>>(cells, cw, dnx, n, cls, width, keys, hist);
}

// out[b * 8 + k] = number of positions of domain bucket b whose domain index is below cut.v[k] (positions of a bucket are
// in ascending domain index, so these are prefix lengths).  hist = bucket sizes; one thread per (b, k).
__global__ void k_bin_prefix(const uint32_t* __restrict__ dom_order, const uint32_t* __restrict__ hist, int nb, BucketOff cut,
                             uint32_t* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nb * 8) return;
    const int b = t / 8, k = t % 8;
    uint32_t lo = 0;
    for (int i = 0; i < b; ++i) lo += hist[i];   // (<= 63 adds)
    const uint32_t beg = lo;
    uint32_t hi = lo + hist[b];
    const uint32_t c = cut.v[k];
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if ((dom_order ? dom_order[mid] : mid) < c) lo = mid + 1; else hi = mid;   // NULL order: one bucket, position = index
    }
    out[t] = lo - beg;
}

// flip isometries: every range block twice (the odd copy is searched mirrored), and the position of every copy after sorting
__global__ void k_dup_items(const fe_grid_item* __restrict__ in, uint32_t n, fe_grid_item* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * n) return;
    out[i] = in[i >> 1];
}
__global__ void k_pos_of(const uint32_t* __restrict__ order, uint32_t n, uint32_t* __restrict__ pos_of) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    pos_of[order ? order[p] : p] = p;
}
__global__ void k_iota(uint32_t* p, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}
// sort key = class + 1 (0..6), histogram of keys
__global__ void k_class_keys(const int32_t* __restrict__ cls, uint32_t n, uint8_t* __restrict__ keys, uint32_t* __restrict__ hist) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = cls[i] + 1;
    keys[i] = (uint8_t)c;
    atomicAdd(&hist[c], 1u);
}

// ---------------------------------------------------------------------------------------------
// operand preparation
// ---------------------------------------------------------------------------------------------
// Range rows.  One warp per range position j (range index order[j]).  Writes the four rows
// 4j+k: fast geometry -> the range block under the INVERSE of rotation k (so that one unrotated
// domain pool serves all four isometries); generic geometry -> four copies of the block.
// Also rowc[j] = 16*sum(r^2).  (Exact dp4a path only: the tensor kinds build their own operands.)
__global__ void k_build_rows(const uint8_t* __restrict__ img, uint32_t stride, const fe_grid_item* __restrict__ rng,
                             const uint32_t* __restrict__ order, uint32_t n, uint32_t T, uint32_t Npad, int fast,
                             uint8_t* __restrict__ A, uint32_t* __restrict__ rowc) {
    const uint32_t j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (j >= n) return;
    const fe_grid_item r = rng[order ? order[j] : j];
    const uint32_t N = T * T;
    uint8_t* row = A + (size_t)j * 4 * Npad;
    uint32_t s2 = 0;
    for (uint32_t e = lane; e < Npad; e += 32) {
        uint8_t v0 = 0, v1 = 0, v2 = 0, v3 = 0;
        if (e < N) {
            const uint32_t Y = e / T, X = e % T;
            const uint8_t* base = img + (size_t)r.y * stride + r.x;
            v0 = base[(size_t)Y * stride + X];
            s2 += (uint32_t)v0 * v0;
            if (fast) {
                v1 = base[(size_t)X * stride + (T - 1 - Y)];           // A_1[Y][X] = R[X][T-1-Y]
                v2 = base[(size_t)(T - 1 - Y) * stride + (T - 1 - X)]; // A_2[Y][X] = R[T-1-Y][T-1-X]
                v3 = base[(size_t)(T - 1 - X) * stride + Y];           // A_3[Y][X] = R[T-1-X][Y]
            } else {
                v1 = v2 = v3 = v0;
            }
        }
        row[e] = v0;
        row[Npad + e] = v1;
        row[2 * Npad + e] = v2;
        row[3 * Npad + e] = v3;
    }
    for (int o = 16; o; o >>= 1) s2 += __shfl_xor_sync(0xFFFFFFFFu, s2, o);
    if (lane == 0) rowc[j] = 16u * s2;
}

// Domain pool.  One warp per (pool k, column c): D[ty][tx] = box sum at local (rho*tx, rho*ty) under
// isometry k (k = 0 only in the fast geometry).  Split into low/high bytes for the dp4a search;
// coln = sum(D^2).
__global__ void k_build_pool(const uint8_t* __restrict__ img, uint32_t stride, const fe_grid_item* __restrict__ dom,
                             const uint32_t* __restrict__ order, uint32_t n, uint32_t npool, uint32_t T, uint32_t rho,
                             uint32_t Npad, uint8_t* __restrict__ Blo, uint8_t* __restrict__ Bhi, uint32_t* __restrict__ coln) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n * npool) return;
    const uint32_t k = w / n, c = w % n;
    const fe_grid_item d = dom[order ? order[c] : c];
    const uint32_t N = T * T;
    uint8_t* lo = Blo + (size_t)w * Npad;
    uint8_t* hi = Bhi + (size_t)w * Npad;
    uint32_t s2 = 0;
    for (uint32_t e = lane; e < Npad; e += 32) {
        uint32_t D = 0;
        if (e < N) D = (uint32_t)sample_sum4(img, stride, d.x, d.y, d.w, (e % T) * rho, (e / T) * rho, (int)k);
        lo[e] = (uint8_t)(D & 255u);
        hi[e] = (uint8_t)(D >> 8);
        s2 += D * D;
    }
    for (int o = 16; o; o >>= 1) s2 += __shfl_xor_sync(0xFFFFFFFFu, s2, o);
    if (lane == 0) coln[w] = s2;
}

// ---------------------------------------------------------------------------------------------
// exact integer search (CUDA cores, dp4a): n16[row][col] = rowc - 8*sum(r*D) + coln, exact in u32
// for T <= 64.  Per row: packed-key argmin (n16 << 32 | col) and first col with n16 <= thr16.
// Tile 128 rows x 64 cols per CTA, 256 threads, 8x4 outputs per thread, K in chunks of 16 words.
// ---------------------------------------------------------------------------------------------
// The reference's distance for one (range, domain, rotation): fp32 running sum over the block, row-major
// (image/metrics.h:38-49); v*v is exact in fp32 so only the additions round.
__device__ float float_seq_sse(const uint8_t* __restrict__ tgt, uint32_t tstride, const fe_grid_item& r, const uint8_t* __restrict__ src,
                               uint32_t sstride, const fe_grid_item& dm, uint32_t rho, int k) {
    float sum = 0.0f;
    const uint32_t T = r.w;
    for (uint32_t ty = 0; ty < T; ++ty)
        for (uint32_t tx = 0; tx < T; ++tx) {
            const float a = (float)tgt[(size_t)(r.y + ty) * tstride + r.x + tx];
            const float d = (float)sample_sum4(src, sstride, dm.x, dm.y, dm.w, tx * rho, ty * rho, k) * 0.25f;
            const float v = __fsub_rn(a, d);
            sum = __fadd_rn(sum, __fmul_rn(v, v));
        }
    return sum;
}

#define SX_TM 128
#define SX_TN 64
#define SX_KW 16
#define SX_APAD 132

template <bool RERANK>
__global__ void __launch_bounds__(256) k_search_exact(SearchArgs a) {
    __shared__ __align__(16) uint32_t As[SX_KW][SX_APAD];
    __shared__ __align__(16) uint32_t Bl[SX_KW][SX_TN];
    __shared__ __align__(16) uint32_t Bh[SX_KW][SX_TN];
    const uint32_t tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const uint32_t nw = a.Npad / 4;
    // a tile never mixes rotations' pools in the generic geometry: blockIdx.y = pool
    const uint32_t pool = blockIdx.y;
    const uint32_t row_base = a.row0 + blockIdx.x * SX_TM;
    const uint32_t rows_here = min((uint32_t)SX_TM, a.row0 + a.nrows - row_base);
    const uint8_t* Blo = a.Blo + (size_t)pool * a.pool_stride_cols * a.Npad;
    const uint8_t* Bhi = a.Bhi + (size_t)pool * a.pool_stride_cols * a.Npad;
    const uint32_t* coln = a.coln + (size_t)pool * a.pool_stride_cols;

    unsigned long long best[8];
    uint32_t hit[8], rc[8], bound[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        best[i] = FE_INF64;
        hit[i] = FE_NONE32;
        const uint32_t r = row_base + ty * 8 + i;
        const bool ok = ty * 8 + i < rows_here;
        rc[i] = ok ? a.rowc[r >> 2] : 0u;
        bound[i] = (RERANK && ok) ? a.rowbound[r >> 2] : 0u;
    }

    for (uint32_t ct = 0; ct < a.ncols; ct += SX_TN) {
        uint32_t lo[8][4], hi[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) lo[i][j] = hi[i][j] = 0;
        for (uint32_t w0 = 0; w0 < nw; w0 += SX_KW) {
            const uint32_t kw = min((uint32_t)SX_KW, nw - w0);
            __syncthreads();
            // A chunk: rows_here x kw words, transposed into As[w][row]
            for (uint32_t idx = tid; idx < SX_TM * kw; idx += 256) {
                const uint32_t r = idx / kw, w = idx % kw;
                uint32_t v = 0;
                if (r < rows_here) v = *reinterpret_cast<const uint32_t*>(a.A + (size_t)(row_base + r) * a.Npad + (size_t)(w0 + w) * 4);
                As[w][r] = v;
            }
            for (uint32_t idx = tid; idx < SX_TN * kw; idx += 256) {
                const uint32_t c = idx / kw, w = idx % kw;
                uint32_t vl = 0, vh = 0;
                if (ct + c < a.ncols) {
                    const size_t off = (size_t)(a.col0 + ct + c) * a.Npad + (size_t)(w0 + w) * 4;
                    vl = *reinterpret_cast<const uint32_t*>(Blo + off);
                    vh = *reinterpret_cast<const uint32_t*>(Bhi + off);
                }
                Bl[w][c] = vl;
                Bh[w][c] = vh;
            }
            __syncthreads();
#pragma unroll 4
            for (uint32_t w = 0; w < kw; ++w) {
                const uint4 a0 = *reinterpret_cast<const uint4*>(&As[w][ty * 8]);
                const uint4 a1 = *reinterpret_cast<const uint4*>(&As[w][ty * 8 + 4]);
                const uint4 bl = *reinterpret_cast<const uint4*>(&Bl[w][tx * 4]);
                const uint4 bh = *reinterpret_cast<const uint4*>(&Bh[w][tx * 4]);
                const uint32_t av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const uint32_t blv[4] = {bl.x, bl.y, bl.z, bl.w};
                const uint32_t bhv[4] = {bh.x, bh.y, bh.z, bh.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        lo[i][j] = __dp4a(av[i], blv[j], lo[i][j]);
                        hi[i][j] = __dp4a(av[i], bhv[j], hi[i][j]);
                    }
            }
        }
        // epilogue of this column tile
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t c = ct + tx * 4 + j;
            if (c < a.ncols) {
                const uint32_t gc = a.col0 + c;
                const uint32_t cn = coln[gc];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint32_t cross = lo[i][j] + (hi[i][j] << 8);
                    const uint32_t n16 = rc[i] - 8u * cross + cn;
                    if (RERANK) {
                        if (ty * 8 + i < rows_here && n16 <= bound[i]) { // near-minimal candidate: score it like the reference does
                            const uint32_t r = row_base + ty * 8 + i;
                            const fe_grid_item rg = a.rng[a.row_range[r >> 2]];
                            const fe_grid_item dm = a.dom[a.dom_order ? a.dom_order[gc] : gc];
                            const int k = a.pool_stride_cols ? (int)pool : (int)(r & 3u);
                            const float f = float_seq_sse(a.tgt, a.tgt_stride, rg, a.src, a.src_stride, dm, a.rho, k);
                            const unsigned long long key = ((unsigned long long)__float_as_uint(f) << 32) | gc;
                            best[i] = key < best[i] ? key : best[i];
                        }
                    } else {
                        const unsigned long long key = ((unsigned long long)n16 << 32) | gc;
                        best[i] = key < best[i] ? key : best[i];
                        if (a.use_thr && n16 <= a.thr16) hit[i] = min(hit[i], gc);
                    }
                }
            }
        }
    }
    // reduce across the 16 threads (tx) sharing the same rows: lanes differ in their low 4 bits
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int o = 8; o; o >>= 1) {
            const unsigned long long ob = __shfl_xor_sync(0xFFFFFFFFu, best[i], o);
            const uint32_t oh = __shfl_xor_sync(0xFFFFFFFFu, hit[i], o);
            best[i] = ob < best[i] ? ob : best[i];
            hit[i] = min(hit[i], oh);
        }
        if (tx == 0 && ty * 8 + i < rows_here) {
            // generic geometry: pool k only answers rows of rotation k
            const uint32_t r = row_base + ty * 8 + i;
            if (a.pool_stride_cols == 0 || (r & 3u) == pool) {
                a.rowbest[r] = best[i];
                a.rowhit[r] = hit[i];
            }
        }
    }
}

cudaError_t launch_search_exact(fe_ctx* ctx, const SearchArgs& a, bool rerank) {
    if (a.nrows == 0 || a.ncols == 0) return cudaSuccess;
    dim3 grid((a.nrows + SX_TM - 1) / SX_TM, a.pool_stride_cols ? 4 : 1);
    if (rerank) k_search_exact<true><<<grid, 256, 0, ctx->stream>>>(a);
    else k_search_exact<false><<<grid, 256, 0, ctx->stream>>>(a);
    ctx->stats.kernel_launches++;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// winner finalisation: TransformEstimator2::estimate's net rule over the four
// rotation rows of a range, then s/o for the winner only (transformmatcher.h:98-108).
// LPR lanes per range position j (launch_k_finalize).
// ---------------------------------------------------------------------------------------------
template <int LPR>
__global__ void k_finalize_t(FinalizeArgs f) {
    const uint32_t gt = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t j = gt / LPR, lane = gt % LPR;              // range position, lane inside the block's group of LPR lanes
    const uint32_t gmask = LPR == 32 ? 0xFFFFFFFFu : (((1u << LPR) - 1u) << ((threadIdx.x & 31u) / LPR * LPR));
    if (j >= f.n) return;                                      // whole groups leave together
    // Rows of this range block and the isometry each one stands for.  Rotations only: rows 4 j + k.  With the flip
    // isometries the block was searched twice (f.rng holds it at 2 j and, mirrored, at 2 j + 1): the mirrored copy under the
    // inverse of rotation q is the block under Flip_Rotate_180, Flip_Rotate_90, Flip, Flip_Rotate_270 for q = 0..3
    // (image/transform.h:37-40 worked out on the decimated grid).
    const int nrow = f.flips ? 8 : 4;
    uint32_t rowidx[8];
    int kof[8];
    uint32_t ri;
    if (!f.flips) {
        ri = f.rng_order ? f.rng_order[j] : j;
#pragma unroll
        for (int k = 0; k < 4; ++k) { rowidx[k] = 4 * j + k; kof[k] = k; rowidx[4 + k] = 0; kof[4 + k] = 0; }
    } else {
        ri = j;
        const uint32_t pa = f.pos_of[2 * j], pb = f.pos_of[2 * j + 1];
        const int km[4] = {6, 5, 4, 7};
#pragma unroll
        for (int k = 0; k < 4; ++k) { rowidx[k] = 4 * pa + k; kof[k] = k; rowidx[4 + k] = 4 * pb + k; kof[4 + k] = km[k]; }
    }
    const fe_grid_item r = f.rng[f.flips ? 2 * ri : ri];
    // ---- pick the winner (identical on all lanes) ----
    int wk = -1, wrow = -1;
    uint32_t wd = 0, wn16 = 0;
    bool from_min = false;
    if (f.use_thr) { // first candidate in scan order c = 8*d + k with distance <= threshold
        unsigned long long bestscan = FE_INF64;
        for (int q = 0; q < nrow; ++q) {
            const uint32_t h = f.rowhit[rowidx[q]];
            if (h != FE_NONE32) {
                const uint32_t d = (f.dom_order && !f.hit_is_domain) ? f.dom_order[h] : h;
                const unsigned long long scan = (unsigned long long)d * 8 + (unsigned)kof[q];
                if (scan < bestscan) { bestscan = scan; wrow = q; }
            }
        }
        if (bestscan != FE_INF64) {
            wk = (int)(bestscan & 7);
            wd = (uint32_t)(bestscan >> 3);
        }
    }
    if (wk < 0 && !f.no_min) { // minimum n16; ties -> smallest domain index, then LARGEST k
        uint32_t bn = 0, bd = 0;
        for (int q = 0; q < nrow; ++q) {
            const unsigned long long key = f.rowbest[rowidx[q]];
            if (key == FE_INF64) continue;
            // normal pass: n16; re-rank pass: bits of the emulated fp32 sum (monotone for non-negative floats)
            const uint32_t n16 = (uint32_t)(key >> 32), col = (uint32_t)key;
            const uint32_t d = f.dom_order ? f.dom_order[col] : col;
            if (wk < 0 || n16 < bn || (n16 == bn && (d < bd || (d == bd && kof[q] > wk)))) {
                wk = kof[q];
                wrow = q;
                bn = n16;
                bd = d;
            }
        }
        wd = bd;
        from_min = true;
    }
    fe_encode_item out;
    out.x = r.x; out.y = r.y; out.w = r.w; out.h = r.h;
    out.pad_ = 0;
    if (wk < 0) { // no admissible candidate: default item_match_t (encode/datatypes.h:8-19)
        out.distance = 100000.0; out.contrast = 0.0; out.brightness = 0.0; out.transform = 0;
        out.match_x = 0; out.match_y = 0; out.src_w = 0; out.src_h = 0;
        if (lane == 0) {
            f.out[ri] = out;
            if (f.split) f.split[ri] = (f.can_split && !(100000.0 <= f.thr)) ? 1u : 0u;
            if (f.bound_out && !f.flips) f.bound_out[j] = 0;
        }
        return;
    }
    const fe_grid_item dm = f.dom[wd];
    // ---- exact sums for the winner ----
    const uint32_t T = r.w, N = T * T, rho = dm.w / T;
    uint32_t sA = 0, sA2 = 0, sB = 0, sAB = 0, sB2 = 0;
    if (rho == 2 && wk < 4 && (T & 1u) == 0 && ((dm.x | f.src_stride) & 3u) == 0 && (reinterpret_cast<uintptr_t>(f.src) & 3u) == 0) {
        // The common case (S = 2T, a rotation): a lane takes whole rows of the decimated domain block -- two source rows read as
        // words, two box sums per word pair -- against the range block under the INVERSE rotation (the same pairs as below).
        // even rotations: the range pixels of a decimated row lie in ONE row of the block (read forwards or backwards) -- words
        const bool rowwise = (wk & 1) == 0 && (T & 3u) == 0 && ((r.x | f.tgt_stride) & 3u) == 0 && (reinterpret_cast<uintptr_t>(f.tgt) & 3u) == 0;
        for (uint32_t ty = lane; ty < T; ty += LPR) {
            const uint32_t* q0 = reinterpret_cast<const uint32_t*>(f.src + (size_t)(dm.y + 2 * ty) * f.src_stride + dm.x);
            const uint32_t* q1 = reinterpret_cast<const uint32_t*>(f.src + (size_t)(dm.y + 2 * ty + 1) * f.src_stride + dm.x);
            if (rowwise) {
                const uint8_t* arow = f.tgt + (size_t)(r.y + (wk == 0 ? ty : T - 1 - ty)) * f.tgt_stride + r.x;
                for (uint32_t tx = 0; tx < T; tx += 4) {
                    uint32_t aw = __ldg(reinterpret_cast<const uint32_t*>(arow + (wk == 0 ? tx : T - 4 - tx)));
                    if (wk == 2) aw = __byte_perm(aw, 0, 0x0123);
#pragma unroll
                    for (uint32_t hlf = 0; hlf < 2; ++hlf) {
                        const uint32_t w0 = __ldg(q0 + tx / 2 + hlf), w1 = __ldg(q1 + tx / 2 + hlf);
                        const uint32_t dd = (w0 & 0x00FF00FFu) + ((w0 >> 8) & 0x00FF00FFu) + (w1 & 0x00FF00FFu) + ((w1 >> 8) & 0x00FF00FFu);
#pragma unroll
                        for (uint32_t j = 0; j < 2; ++j) {
                            const uint32_t D = j ? dd >> 16 : dd & 0xFFFFu;
                            const uint32_t a = (aw >> (8 * (2 * hlf + j))) & 255u;
                            sA += a; sA2 += a * a; sB += D; sAB += a * D; sB2 += D * D;
                        }
                    }
                }
                continue;
            }
            for (uint32_t tx = 0; tx < T; tx += 2) {
                const uint32_t w0 = __ldg(q0 + tx / 2), w1 = __ldg(q1 + tx / 2);
                const uint32_t dd = (w0 & 0x00FF00FFu) + ((w0 >> 8) & 0x00FF00FFu) + (w1 & 0x00FF00FFu) + ((w1 >> 8) & 0x00FF00FFu);
#pragma unroll
                for (uint32_t j = 0; j < 2; ++j) {
                    const uint32_t D = j ? dd >> 16 : dd & 0xFFFFu, x = tx + j;
                    const uint32_t py = wk == 0 ? ty : wk == 1 ? x : wk == 2 ? T - 1 - ty : T - 1 - x;
                    const uint32_t px = wk == 0 ? x : wk == 1 ? T - 1 - ty : wk == 2 ? T - 1 - x : ty;
                    const uint32_t a = f.tgt[(size_t)(r.y + py) * f.tgt_stride + r.x + px];
                    sA += a; sA2 += a * a; sB += D; sAB += a * D; sB2 += D * D;
                }
            }
        }
    } else {
        for (uint32_t e = lane; e < N; e += LPR) {
            const uint32_t ty = e / T, tx = e % T;
            const uint32_t a = f.tgt[(size_t)(r.y + ty) * f.tgt_stride + r.x + tx];
            const uint32_t D = (uint32_t)sample_sum4(f.src, f.src_stride, dm.x, dm.y, dm.w, tx * rho, ty * rho, wk);
            sA += a; sA2 += a * a; sB += D; sAB += a * D; sB2 += D * D;
        }
    }
    for (int o = LPR / 2; o; o >>= 1) {
        sA += __shfl_xor_sync(gmask, sA, o);
        sA2 += __shfl_xor_sync(gmask, sA2, o);
        sB += __shfl_xor_sync(gmask, sB, o);
        sAB += __shfl_xor_sync(gmask, sAB, o);
        sB2 += __shfl_xor_sync(gmask, sB2, o);
    }
    wn16 = 16u * sA2 - 8u * sAB + sB2; // exact for T <= 64
    if (lane != 0) return;
    // self-check of the search kernel's arithmetic against the direct recomputation
    if (f.rerank) {
        // keys are float sums here; nothing to cross-check against the integer recomputation
    } else if (from_min) {
        const unsigned long long key = f.rowbest[rowidx[wrow]];
        if ((uint32_t)(key >> 32) != wn16) atomicAdd(f.mismatch, 1u);
    } else if (wn16 > f.thr16) {
        atomicAdd(f.mismatch, 1u);
    }
    if (f.bound_out && !f.flips) {
        // fp32 regime (SSE >= 2^20): the reference ranks candidates by a ROUNDED running sum, so every candidate whose
        // exact score lies within the worst-case rounding band of the exact minimum must be re-scored.
        // |fl_seq(S) - S| <= N * ulp(S) / 2; band = 2 * N * ulp(S_min) in SSE units = 32 * N * ulp in n16 units.
        uint32_t b = 0;
        if (wn16 >= (1u << 24) && from_min) {
            const float smin = (float)(wn16 >> 4);
            const float ulp = __uint_as_float((__float_as_uint(smin) & 0x7F800000u) - (23u << 23));
            const double band = 32.0 * (double)N * (double)ulp;
            const double bb = (double)wn16 + band;
            b = bb >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)bb;
        }
        f.bound_out[j] = b;
    }
    double distance;
    if (wn16 < (1u << 24)) {
        distance = ((double)wn16 / 16.0) / (double)(dm.w * dm.h);
    } else { // reference fp32 running sum rounds (metrics.h:38-49): emulate it, row-major
        float sum = 0.0f;
        for (uint32_t ty = 0; ty < T; ++ty)
            for (uint32_t tx = 0; tx < T; ++tx) {
                const float a = (float)f.tgt[(size_t)(r.y + ty) * f.tgt_stride + r.x + tx];
                const float d = (float)sample_sum4(f.src, f.src_stride, dm.x, dm.y, dm.w, tx * rho, ty * rho, wk) * 0.25f;
                const float v = __fsub_rn(a, d);
                sum = __fadd_rn(sum, __fmul_rn(v, v));
            }
        distance = (double)sum / (double)(dm.w * dm.h);
        atomicAdd(f.fp32_regime, 1u);
    }
    // least squares exactly as transformmatcher.h:98-108 (A = range, B = decimated domain)
    const double Nd = (double)N, sumA = (double)sA, sumA2 = (double)sA2;
    const double sumB = (double)sB * 0.25, sumAB = (double)sAB * 0.25;
    const double tmp = __dsub_rn(__dmul_rn(Nd, sumA2), __dmul_rn(sumA - 1.0, sumA));
    double s = fabs(tmp) < 0.00001 ? 0.0 : __ddiv_rn(__dsub_rn(__dmul_rn(Nd, sumAB), __dmul_rn(sumA, sumB)), tmp);
    if (f.s_max > 0.0) s = s > f.s_max ? f.s_max : (s < -f.s_max ? -f.s_max : s);
    const double o = f.fma ? __ddiv_rn(__fma_rn(-s, sumA, sumB), Nd) : __ddiv_rn(__dsub_rn(sumB, __dmul_rn(s, sumA)), Nd);
    out.distance = distance;
    out.contrast = s;
    out.brightness = o;
    out.transform = wk;
    out.match_x = dm.x; out.match_y = dm.y; out.src_w = dm.w; out.src_h = dm.h;
    f.out[ri] = out;
    if (f.split) f.split[ri] = (f.can_split && !(distance <= f.thr)) ? 1u : 0u;
}

// Lanes per range block: a warp per block leaves most lanes idle on the small blocks (64 pixels at T = 8) and the kernel
// is a chain of dependent loads per block, so small blocks share a warp.
void launch_k_finalize(cudaStream_t stream, const FinalizeArgs& f, uint32_t T) {
    const uint32_t N = T * T;
    if (N <= 16) k_finalize_t<4><<<(unsigned)(((uint64_t)f.n * 4 + 255) / 256), 256, 0, stream>>>(f);
    else if (N <= 64) k_finalize_t<8><<<(unsigned)(((uint64_t)f.n * 8 + 255) / 256), 256, 0, stream>>>(f);
    else k_finalize_t<32><<<(unsigned)(((uint64_t)f.n * 32 + 255) / 256), 256, 0, stream>>>(f);
}

// ---------------------------------------------------------------------------------------------
// decode (Frac::copy as a gather kernel) + convergence sum
// ---------------------------------------------------------------------------------------------
// One thread per target pixel of an item list with prefix offsets (pix_off[i] = first pixel id of item i).
__global__ void k_decode_step(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uint32_t stride,
                              const fe_encode_item* __restrict__ items, const uint32_t* __restrict__ pix_off, uint32_t n_items,
                              uint32_t total_pix, int use_fma, const uint32_t* __restrict__ done) {
    if (done && *done) return;            // the iteration already converged: every later launch is a no-op
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total_pix) return;
    // binary search the owning item
    uint32_t lo = 0, hi = n_items;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (pix_off[mid] <= p) lo = mid; else hi = mid;
    }
    const fe_encode_item e = items[lo];
    if (e.src_w == 0 || e.src_h == 0) return; // default item: nothing to sample (see oracle/frac_oracle.c)
    const uint32_t q = p - pix_off[lo], x = q % e.w, y = q / e.w;
    const uint32_t sx = (x * e.src_w) / e.w, sy = (y * e.src_h) / e.h;
    const double smp = (double)sample_sum4(src, stride, e.match_x, e.match_y, e.src_w, sx, sy, e.transform) * 0.25;
    const double v = use_fma ? __fma_rn(e.contrast, smp, e.brightness) : __dadd_rn(__dmul_rn(e.contrast, smp), e.brightness);
    dst[(size_t)(e.y + y) * stride + e.x + x] = v < 0.0 ? 0 : (v > 255.0 ? 255 : (uint8_t)v);
}

// Square items of one size T (T % 4 == 0): one thread per 4 horizontal pixels, packed 4-byte store.  When `sq_out` is
// given the thread also accumulates (old - new)^2 of its pixels (old = source at the same position): the convergence sum
// of Decoder2 (metrics.h:26-36) without a second pass over both planes -- valid when the items tile the plane.
__global__ void k_decode_step_uniform(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uint32_t stride,
                                      const fe_encode_item* __restrict__ items, uint32_t n_items, uint32_t T, int use_fma,
                                      unsigned long long* __restrict__ sq_out, const uint32_t* __restrict__ done) {
    if (done && *done) return;            // the iteration already converged: every later launch is a no-op
    const uint32_t segs = T / 4, per_item = T * segs;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long sq = 0;
    if (t < n_items * per_item) {
        const uint32_t i = t / per_item, q = t % per_item, y = q / segs, x0 = (q % segs) * 4;
        const fe_encode_item e = items[i];
        if (e.src_w != 0 && e.src_h != 0) {
            uint32_t packed = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t x = x0 + k;
                const uint32_t sx = (x * e.src_w) / T, sy = (y * e.src_h) / T;
                const double smp = (double)sample_sum4(src, stride, e.match_x, e.match_y, e.src_w, sx, sy, e.transform) * 0.25;
                const double v = use_fma ? __fma_rn(e.contrast, smp, e.brightness) : __dadd_rn(__dmul_rn(e.contrast, smp), e.brightness);
                const uint32_t b = v < 0.0 ? 0u : (v > 255.0 ? 255u : (uint32_t)(uint8_t)v);
                packed |= b << (8 * k);
            }
            const size_t off = (size_t)(e.y + y) * stride + e.x + x0;
            uint8_t* o = dst + off;
            if ((reinterpret_cast<uintptr_t>(o) & 3u) == 0) *reinterpret_cast<uint32_t*>(o) = packed;
            else { o[0] = packed & 255; o[1] = (packed >> 8) & 255; o[2] = (packed >> 16) & 255; o[3] = packed >> 24; }
            if (sq_out) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int d = (int)src[off + k] - (int)((packed >> (8 * k)) & 255u);
                    sq += (unsigned long long)(d * d);
                }
            }
        }
    }
    if (sq_out) {
        __shared__ unsigned long long wsum[32];
        for (int o = 16; o; o >>= 1) sq += __shfl_xor_sync(0xFFFFFFFFu, sq, o);
        if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = sq;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long tot = 0;
            for (uint32_t w = 0; w < (blockDim.x >> 5); ++w) tot += wsum[w];
            if (tot) atomicAdd(sq_out + (blockIdx.x & 63u), tot);
        }
    }
}

// Decode gather for the encoder's own geometry (square T x T item, 2T x 2T source block, even origin, T % 4 == 0,
// 4-byte aligned rows): one warp per item.  The source block is read once with coalesced 4-byte loads and reduced to the
// T x T grid of 2x2 box sums in shared memory; every output pixel is then one shared-memory read under the item's isometry,
// s * (D/4) + o in fp64 (the reference's arithmetic), and the row is written back as packed 4-byte stores.  Traffic per
// iteration = one read of every source block (4x the item area, served by L2: the plane is re-read 4x but fits) + one
// write of the plane (+ one read of the old plane for the fused convergence sum).
__global__ void __launch_bounds__(256) k_decode_step_tiled(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uint32_t stride,
                                                           const fe_encode_item* __restrict__ items, uint32_t n_items, uint32_t T, int use_fma,
                                                           unsigned long long* __restrict__ sq_out, const uint32_t* __restrict__ done) {
    if (done && *done) return;            // the iteration already converged: every later launch is a no-op
    extern __shared__ uint16_t sh_box[];                       // [warps per block][T*T]
    const uint32_t warp_in_block = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t i = blockIdx.x * (blockDim.x >> 5) + warp_in_block;
    unsigned long long sq = 0;
    if (i < n_items) {
        const fe_encode_item e = items[i];
        uint16_t* box = sh_box + (size_t)warp_in_block * T * T;
        const uint32_t S = 2 * T, wpr = S / 4;                  // 4-byte words per source row
        // pairs of source rows -> one row of box sums; a lane handles one word (4 pixels -> 2 boxes) of a row pair
        for (uint32_t w = lane; w < T * wpr; w += 32) {
            const uint32_t Y = w / wpr, xw = w % wpr;
            const uint8_t* p = src + (size_t)(e.match_y + 2 * Y) * stride + e.match_x + 4 * xw;
            const uint32_t r0 = *reinterpret_cast<const uint32_t*>(p), r1 = *reinterpret_cast<const uint32_t*>(p + stride);
            const uint32_t b0 = (r0 & 255u) + ((r0 >> 8) & 255u) + (r1 & 255u) + ((r1 >> 8) & 255u);
            const uint32_t b1 = ((r0 >> 16) & 255u) + (r0 >> 24) + ((r1 >> 16) & 255u) + (r1 >> 24);
            *reinterpret_cast<uint32_t*>(box + Y * T + 2 * xw) = b0 | (b1 << 16);
        }
        __syncwarp();
        const int t = e.transform;
        const int m0 = kMapDev[t][0], m1 = kMapDev[t][1], m4 = kMapDev[t][4], m5 = kMapDev[t][5];
        const int cx = (kMapDev[t][2] + kMapDev[t][3]) * ((int)S - 1), cy = (kMapDev[t][6] + kMapDev[t][7]) * ((int)S - 1);
        const int ax = (m0 + m1) < 0 ? -1 : 0, ay = (m4 + m5) < 0 ? -1 : 0;     // min corner of the mapped 2x2 box
        const uint32_t segs = T / 4;
        for (uint32_t q = lane; q < T * segs; q += 32) {
            const uint32_t y = q / segs, x0 = (q % segs) * 4;
            uint32_t packed = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int lx = 2 * (int)(x0 + k), ly = 2 * (int)y;
                const int gx = m0 * lx + m1 * ly + cx + ax, gy = m4 * lx + m5 * ly + cy + ay;
                const double smp = (double)box[(gy >> 1) * (int)T + (gx >> 1)] * 0.25;
                const double v = use_fma ? __fma_rn(e.contrast, smp, e.brightness) : __dadd_rn(__dmul_rn(e.contrast, smp), e.brightness);
                const uint32_t b = v < 0.0 ? 0u : (v > 255.0 ? 255u : (uint32_t)(uint8_t)v);
                packed |= b << (8 * k);
            }
            const size_t off = (size_t)(e.y + y) * stride + e.x + x0;
            *reinterpret_cast<uint32_t*>(dst + off) = packed;
            if (sq_out) {
                const uint32_t old = *reinterpret_cast<const uint32_t*>(src + off);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int d = (int)((old >> (8 * k)) & 255u) - (int)((packed >> (8 * k)) & 255u);
                    sq += (unsigned long long)(d * d);
                }
            }
        }
    }
    if (sq_out) { // one atomic per block, spread over 64 slots: a single hot address would serialise the whole grid
        __shared__ unsigned long long wsum[8];
        for (int o = 16; o; o >>= 1) sq += __shfl_xor_sync(0xFFFFFFFFu, sq, o);
        if (lane == 0) wsum[warp_in_block] = sq;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long tot = 0;
            for (uint32_t w = 0; w < (blockDim.x >> 5); ++w) tot += wsum[w];
            if (tot) atomicAdd(sq_out + (blockIdx.x & 63u), tot);
        }
    }
}

// Same gather for the small blocks (T = 4, T = 8: eight items per warp): a warp per item leaves most lanes idle there and
// chains item load -> source loads -> store once per item; here a warp fetches IPW item records with one coalesced read,
// issues all its source loads before it touches any of them, and every lane has output work.
// DQ: the iterate travels with its half-resolution plane of 2 x 2 box sums (u16).  The gather reads the T x T box sums of an
// item's source block from that plane (T rows of 2 T bytes instead of 2 T rows of 2 T bytes: half the lines), and writes the
// box sums of its own output next to the pixels for the next iteration (items sit at even origins, so a box never straddles two).
template <int T, int IPW, bool DQ, bool FMA>
__global__ void __launch_bounds__(256) k_decode_step_small(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uint32_t stride,
                                                           const fe_encode_item* __restrict__ items, uint32_t n_items,
                                                           unsigned long long* __restrict__ sq_out, const uint32_t* __restrict__ done,
                                                           const uint16_t* __restrict__ dq_src, uint16_t* __restrict__ dq_dst, uint32_t dq_stride) {
    if (done && *done) return;            // the iteration already converged: every later launch is a no-op
    constexpr int N = T * T, S = 2 * T, WPR = S / 4, UNITS = DQ ? T * (T / 2) : T * WPR;   // words (DQ) / row-pair words per item
    constexpr int U = (IPW * UNITS + 31) / 32;                                // load units per lane
    constexpr int SEGS = T / 4, OUT = T * SEGS, Q = (IPW * OUT + 31) / 32;    // 4-pixel output segments per item / per lane
    // item records, 80 bytes apart: the lanes of an output instruction read the same field of eight different records
    __shared__ uint4 sitem_raw[8][IPW * 5];
    auto item_of = [&](uint32_t w, uint32_t k) -> const fe_encode_item& { return *reinterpret_cast<const fe_encode_item*>(&sitem_raw[w][5 * k]); };
    // Box sums of the warp's items.  The lanes of an output instruction are (rows) x (items) x (segments): the items' slots are
    // skewed so that an identity-oriented read of one instruction falls into 32 different banks (T = 8: slot k starts at word
    // 32 k + 8 (k / 2) + k % 2; T = 4: 8 k + k / 4).
    auto box_of = [](uint16_t* base, uint32_t k) { return base + 2 * (T == 8 ? 32 * k + 8 * (k >> 1) + (k & 1u) : (N / 2) * k + (k >> 2)); };
    constexpr int WARP_U16 = 2 * (T == 8 ? 32 * (IPW - 1) + 8 * ((IPW - 1) >> 1) + 1 + 32 : (N / 2) * IPW + (IPW >> 2) + 1);
    __shared__ __align__(16) uint16_t sbox_raw[8][(WARP_U16 + 7) & ~7];
    const uint32_t warp_in_block = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t first = (blockIdx.x * 8 + warp_in_block) * IPW;
    unsigned long long sq = 0;
    if (first < n_items) {
        const uint32_t cnt = min((uint32_t)IPW, n_items - first);
        // item records: 4 x 16 bytes each, one coalesced read
        for (uint32_t q = lane; q < cnt * 4; q += 32)
            sitem_raw[warp_in_block][5 * (q >> 2) + (q & 3u)] = __ldg(reinterpret_cast<const uint4*>(items + first) + q);
        __syncwarp();
        // all source loads first, then the box sums
        uint32_t r0[U], r1[DQ ? 1 : U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t idx = lane + 32 * u, k = idx / UNITS, w = idx % UNITS;
            r0[u] = 0;
            if (!DQ) r1[u] = 0;
            if (k < cnt) {
                const fe_encode_item& e = item_of(warp_in_block, k);
                if (DQ) {
                    r0[u] = __ldg(reinterpret_cast<const uint32_t*>(dq_src + (size_t)(e.match_y / 2 + w / (T / 2)) * dq_stride + e.match_x / 2) + w % (T / 2));
                } else {
                    const uint8_t* p = src + (size_t)(e.match_y + 2 * (w / WPR)) * stride + e.match_x + 4 * (w % WPR);
                    r0[u] = __ldg(reinterpret_cast<const uint32_t*>(p));
                    r1[u] = __ldg(reinterpret_cast<const uint32_t*>(p + stride));
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t idx = lane + 32 * u, k = idx / UNITS, w = idx % UNITS;
            if (k < cnt) {
                if (DQ) {
                    *reinterpret_cast<uint32_t*>(box_of(sbox_raw[warp_in_block], k) + (w / (T / 2)) * T + 2 * (w % (T / 2))) = r0[u];
                } else {
                    const uint32_t dd = (r0[u] & 0x00FF00FFu) + ((r0[u] >> 8) & 0x00FF00FFu) + (r1[u] & 0x00FF00FFu) + ((r1[u] >> 8) & 0x00FF00FFu);
                    *reinterpret_cast<uint32_t*>(box_of(sbox_raw[warp_in_block], k) + (w / WPR) * T + 2 * (w % WPR)) = dd;
                }
            }
        }
        __syncwarp();
        // Lanes run along an output row ACROSS the warp's items: consecutive items of a level list are neighbours in the plane
        // (Z-order), so the segments of a row that lie side by side share a sector for the store and the old-value read.  A lane
        // keeps its (item, segment) for all its rows: the item's fields are read once.
        static_assert(32 % (IPW * SEGS) == 0 && (IPW * OUT) % 32 == 0, "rows of the warp's items per instruction");
        constexpr uint32_t LPRW = IPW * SEGS, RPI = 32 / LPRW;       // lanes per output row of the warp, rows per instruction
        static_assert(RPI % 2 == 0, "a lane and its box partner (the next row) sit in the same instruction");
        const uint32_t k = (lane % LPRW) / SEGS, x0 = (lane % SEGS) * 4;
        const bool live = k < cnt;
        const fe_encode_item& e = item_of(warp_in_block, live ? k : 0);
        const uint16_t* box = box_of(sbox_raw[warp_in_block], live ? k : 0);
        // rows of kMapDev as bit fields (m + 1, two bits each: m0, m1, m4, m5), one byte per isometry: the lanes of a warp hold
        // different isometries, and a constant-bank read with a divergent index is replayed once per distinct value
        const uint32_t code = (uint32_t)(0x4194691661144996ull >> (8 * (e.transform & 7))) & 0xFFu;
        const int m0 = (int)(code & 3u) - 1, m1 = (int)((code >> 2) & 3u) - 1, m4 = (int)((code >> 4) & 3u) - 1, m5 = (int)(code >> 6) - 1;
        // Box of output pixel (x, y): the mapped 2 x 2 source box has its min corner at the even position
        // (2 (m0 x + m1 y) + cx + ax, ...) with cx + ax = S - 2 where the coefficients sum to -1, else 0 (transform.h:32-41):
        // box index = base0 + y * sy + kk * sk, all per lane.
        const int cxh = (m0 + m1) < 0 ? T - 1 : 0, cyh = (m4 + m5) < 0 ? T - 1 : 0;
        const int base0 = (m4 * (int)x0 + cyh) * T + m0 * (int)x0 + cxh, sy = m5 * T + m1, sk = m4 * T + m0;
        // s * (D / 4) + o: the scaling by a power of two commutes with the rounding of the product, so (s / 4) * D is the same double
        const double cs4 = e.contrast * 0.25, br = e.brightness;
        const uint32_t ex = e.x, ey = e.y;
        uint32_t sq32 = 0;                         // <= 16 pixels x 255^2 per lane
#pragma unroll
        for (int qq = 0; qq < Q; ++qq) {
            const uint32_t y = lane / LPRW + qq * RPI;
            uint32_t packed = 0;
            if (live) {
                const uint16_t* bp = box + base0 + (int)y * sy;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const double D = (double)(uint32_t)bp[kk * sk];
                    const double v = FMA ? __fma_rn(cs4, D, br) : __dadd_rn(__dmul_rn(cs4, D), br);
                    // v < 0 -> 0, v > 255 -> 255, else truncation (DecodeUtils.hpp:19-22): the saturating conversion and a minimum
                    packed |= min(__double2uint_rz(v), 255u) << (8 * kk);
                }
            }
            if (DQ) {   // box sums of the output: this row and the next one (lane ^ LPRW), even rows write
                const uint32_t other = __shfl_xor_sync(0xFFFFFFFFu, packed, LPRW);
                if (live && (y & 1u) == 0) {
                    const uint32_t lo = __dp4a(packed, 0x00000101u, __dp4a(other, 0x00000101u, 0u));
                    const uint32_t hi = __dp4a(packed, 0x01010000u, __dp4a(other, 0x01010000u, 0u));
                    *reinterpret_cast<uint32_t*>(dq_dst + (size_t)((ey + y) / 2) * dq_stride + (ex + x0) / 2) = lo | (hi << 16);
                }
            }
            if (live) {
                const size_t off = (size_t)(ey + y) * stride + ex + x0;
                *reinterpret_cast<uint32_t*>(dst + off) = packed;
                if (sq_out) {
                    const uint32_t ad = __vabsdiffu4(__ldg(reinterpret_cast<const uint32_t*>(src + off)), packed);
                    sq32 = __dp4a(ad, ad, sq32);
                }
            }
        }
        sq = sq32;
    }
    if (sq_out) { // one atomic per block, spread over 64 slots
        __shared__ unsigned long long wsum[8];
        for (int o = 16; o; o >>= 1) sq += __shfl_xor_sync(0xFFFFFFFFu, sq, o);
        if (lane == 0) wsum[warp_in_block] = sq;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long tot = 0;
            for (uint32_t w = 0; w < 8; ++w) tot += wsum[w];
            if (tot) atomicAdd(sq_out + (blockIdx.x & 63u), tot);
        }
    }
}

bool launch_decode_step_small(cudaStream_t stream, const uint8_t* src, uint8_t* dst, uint32_t stride, const fe_encode_item* items, uint32_t n,
                              uint32_t T, int use_fma, unsigned long long* sq_out, const uint32_t* done, const uint16_t* dq_src, uint16_t* dq_dst,
                              uint32_t dq_stride) {
    const unsigned grid = (n + 63) / 64;
    if (T != 4 && T != 8) return false;
#define DEC_SMALL(TT, DQ_, FMA_) \
    k_decode_step_small<TT, 8, DQ_, FMA_><<<grid, 256, 0, stream>>>(src, dst, stride, items, n, sq_out, done, dq_src, dq_dst, dq_stride)
    const int sel = (T == 8 ? 4 : 0) | (dq_src ? 2 : 0) | (use_fma ? 1 : 0);
    switch (sel) {
    case 0: DEC_SMALL(4, false, false); break;
    case 1: DEC_SMALL(4, false, true); break;
    case 2: DEC_SMALL(4, true, false); break;
    case 3: DEC_SMALL(4, true, true); break;
    case 4: DEC_SMALL(8, false, false); break;
    case 5: DEC_SMALL(8, false, true); break;
    case 6: DEC_SMALL(8, true, false); break;
    default: DEC_SMALL(8, true, true); break;
    }
#undef DEC_SMALL
    return true;
}
// box-sum plane of a u8 plane (first iterate of a decode that carries one)
__global__ void k_boxsum_plane(const uint8_t* __restrict__ src, uint32_t stride, uint32_t qw, uint32_t qh, uint16_t* __restrict__ dq) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;    // x: pairs of boxes
    if (2 * x >= qw || y >= qh) return;
    const uint32_t w0 = *reinterpret_cast<const uint32_t*>(src + (size_t)(2 * y) * stride + 4 * x);
    const uint32_t w1 = *reinterpret_cast<const uint32_t*>(src + (size_t)(2 * y + 1) * stride + 4 * x);
    const uint32_t lo = __dp4a(w0, 0x00000101u, __dp4a(w1, 0x00000101u, 0u)), hi = __dp4a(w0, 0x01010000u, __dp4a(w1, 0x01010000u, 0u));
    *reinterpret_cast<uint32_t*>(dq + (size_t)y * qw + 2 * x) = lo | (hi << 16);
}
void launch_boxsum_plane(cudaStream_t stream, const uint8_t* src, uint32_t stride, uint32_t w, uint32_t h, uint16_t* dq) {
    const uint32_t qw = w / 2, qh = h / 2;
    k_boxsum_plane<<<dim3((qw / 2 + 31) / 32, (qh + 7) / 8), dim3(32, 8), 0, stream>>>(src, stride, qw, qh, dq);
}

// Coverage proof for the ping-pong decode: every pixel must be written by exactly one item.  One thread per (item, row):
// the row's pixels are OR-ed into a bitmap (wpr words per image row); a bit that was already set means two items overlap.
__global__ void k_cover_bitmap(const fe_encode_item* __restrict__ items, uint32_t n, const uint32_t* __restrict__ row_off, uint32_t total_rows,
                               uint32_t wpr, uint32_t* __restrict__ bitmap, uint32_t* __restrict__ overlap) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total_rows) return;
    uint32_t lo = 0, hi = n;                       // item whose rows contain t
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (row_off[mid] <= t) lo = mid; else hi = mid;
    }
    const fe_encode_item e = items[lo];
    if (e.src_w == 0 || e.src_h == 0) return;      // default item: writes nothing
    const uint32_t y = e.y + (t - row_off[lo]), x0 = e.x, x1 = e.x + e.w;   // [x0, x1)
    for (uint32_t w = x0 / 32; w <= (x1 - 1) / 32; ++w) {
        const uint32_t a = max(x0, w * 32) - w * 32, b = min(x1, w * 32 + 32) - w * 32;   // bits [a, b)
        const uint32_t mask = (b - a == 32 ? 0xFFFFFFFFu : ((1u << (b - a)) - 1u) << a);
        const uint32_t old = atomicOr(&bitmap[(size_t)y * wpr + w], mask);
        if (old & mask) atomicExch(overlap, 1u);
    }
}
__global__ void k_popcount(const uint32_t* __restrict__ words, size_t n, unsigned long long* __restrict__ out) {
    unsigned long long s = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) s += __popc(words[i]);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}
// Decoder2's convergence test (encode/Encoder2.hpp:83-85) on the device, once per iteration: the int32-wrapped sum of
// squared differences over the pixel count; the first iteration below eps raises `done`, which turns every later launch of the
// train into a no-op, so the host only looks every few iterations.  state = {done, iterations, rms bits lo, hi}; one warp.
__global__ void k_decode_check(unsigned long long* __restrict__ sums, uint32_t* __restrict__ state, uint32_t npix, double eps, uint32_t iter) {
    if (state[0]) return;
    unsigned long long v = sums[threadIdx.x] + sums[threadIdx.x + 32];
    sums[threadIdx.x] = 0; sums[threadIdx.x + 32] = 0;
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if (threadIdx.x) return;
    const int32_t wrapped = (int32_t)(uint32_t)(v & 0xFFFFFFFFull);      // the reference's int32 accumulator (metrics.h:27)
    const double rms = __ddiv_rn((double)wrapped, (double)npix);
    const unsigned long long bits = (unsigned long long)__double_as_longlong(rms);
    state[2] = (uint32_t)bits; state[3] = (uint32_t)(bits >> 32);
    if (rms < eps) { state[0] = 1; state[1] = iter; } else state[1] = iter + 1;
}
// source = target.copy() of a decode that does not tile the plane (Encoder2.hpp:86), skipped once converged
__global__ void k_copy_plane_if_running(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n16, const uint32_t* __restrict__ done) {
    if (*done) return;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

// sum over the plane of (a-b)^2 as uint64 (the reference accumulates in int32, metrics.h:27; the
// host wraps the 64-bit sum to int32 to reproduce it).
__global__ void k_sqdiff(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, uint32_t w, uint32_t h, uint32_t stride,
                         unsigned long long* __restrict__ out, const uint32_t* __restrict__ done) {
    if (done && *done) return;            // the iteration already converged: every later launch is a no-op
    unsigned long long s = 0;
    const size_t n = (size_t)w * h;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t off = (i / w) * stride + (i % w);
        const int d = (int)a[off] - (int)b[off];
        s += (unsigned long long)(d * d);
    }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    __shared__ unsigned long long ws[32];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? ws[threadIdx.x] : 0ull;
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
        if (threadIdx.x == 0) atomicAdd(out + (blockIdx.x & 63u), s);
    }
}

// ---------------------------------------------------------------------------------------------
// quantizer (Quantizer.hpp:13-36) : min/max then quantized()
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long dbl_key(double v) { // order-preserving map double -> u64
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__global__ void k_minmax(const fe_encode_item* __restrict__ items, uint32_t n, unsigned long long* __restrict__ mm) {
    // mm[0]=min key s, mm[1]=max key s, mm[2]=min key o, mm[3]=max key o
    unsigned long long mns = FE_INF64, mxs = 0, mno = FE_INF64, mxo = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned long long ks = dbl_key(items[i].contrast), ko = dbl_key(items[i].brightness);
        mns = ks < mns ? ks : mns; mxs = ks > mxs ? ks : mxs;
        mno = ko < mno ? ko : mno; mxo = ko > mxo ? ko : mxo;
    }
    for (int o = 16; o; o >>= 1) {
        unsigned long long t;
        t = __shfl_xor_sync(0xFFFFFFFFu, mns, o); mns = t < mns ? t : mns;
        t = __shfl_xor_sync(0xFFFFFFFFu, mxs, o); mxs = t > mxs ? t : mxs;
        t = __shfl_xor_sync(0xFFFFFFFFu, mno, o); mno = t < mno ? t : mno;
        t = __shfl_xor_sync(0xFFFFFFFFu, mxo, o); mxo = t > mxo ? t : mxo;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&mm[0], mns); atomicMax(&mm[1], mxs); atomicMin(&mm[2], mno); atomicMax(&mm[3], mxo);
    }
}
// order-preserving keys of k_minmax -> the doubles main.cpp:109-118 ends up with (max starts at -1, min at DBL_MAX)
__global__ void k_minmax_finish(const unsigned long long* __restrict__ keys, double* __restrict__ out) {
    const int i = threadIdx.x;
    if (i >= 4) return;
    const unsigned long long k = keys[i];
    const unsigned long long b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    const double d = __longlong_as_double((long long)b);
    out[i] = (i & 1) ? fmax(-1.0, d) : fmin(1.7976931348623157e308, d);
}
__global__ void k_quantize(const fe_encode_item* __restrict__ items, uint32_t n, double min_s, double max_s, double min_o,
                           double max_o, int bits_s, int bits_o, uint32_t* __restrict__ qs, uint32_t* __restrict__ qo) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double step_s = __ddiv_rn(fabs(__dsub_rn(max_s, min_s)), (double)(1 << bits_s));
    const double step_o = __ddiv_rn(fabs(__dsub_rn(max_o, min_o)), (double)(1 << bits_o));
    const unsigned long long mq_s = (1ull << bits_s) - 1, mq_o = (1ull << bits_o) - 1;
    const unsigned long long a = (unsigned long long)floor(__ddiv_rn(__dsub_rn(items[i].contrast, min_s), step_s));
    const unsigned long long b = (unsigned long long)floor(__ddiv_rn(__dsub_rn(items[i].brightness, min_o), step_o));
    qs[i] = (uint32_t)(a < mq_s ? a : mq_s);
    qo[i] = (uint32_t)(b < mq_o ? b : mq_o);
}

// packed 64-bit records (layout in include/fractencode_b200.h)
__global__ void k_pack(const fe_encode_item* __restrict__ items, uint32_t n, uint32_t t_max, const double* __restrict__ mm, int bits_s,
                       int bits_o, unsigned long long* __restrict__ out, uint32_t* __restrict__ bad) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double min_s = mm[0], max_s = mm[1], min_o = mm[2], max_o = mm[3];   // device memory: may come from an all-reduce
    const fe_encode_item e = items[i];
    const uint32_t T = e.w;
    uint32_t level = 0;
    while ((t_max >> level) > T && level < 4) ++level;
    const bool has = e.src_w != 0;
    const bool ok = T && e.w == e.h && (t_max >> level) == T && level < 4 && e.x % T == 0 && e.y % T == 0 && e.x / T < 2048 && e.y / T < 2048 &&
                    (!has || (e.src_w == 2 * T && e.src_h == 2 * T && e.match_x % T == 0 && e.match_y % T == 0 && e.match_x / T < 2048 &&
                              e.match_y / T < 2048 && e.transform >= 0 && e.transform < 8));
    if (!ok) { atomicAdd(bad, 1u); out[i] = 0; return; }
    unsigned long long w = (unsigned long long)(e.x / T) | ((unsigned long long)(e.y / T) << 11) | ((unsigned long long)level << 44);
    if (has) {
        const double step_s = __ddiv_rn(fabs(__dsub_rn(max_s, min_s)), (double)(1 << bits_s));
        const double step_o = __ddiv_rn(fabs(__dsub_rn(max_o, min_o)), (double)(1 << bits_o));
        const unsigned long long mq_s = (1ull << bits_s) - 1, mq_o = (1ull << bits_o) - 1;
        unsigned long long qs = (unsigned long long)floor(__ddiv_rn(__dsub_rn(e.contrast, min_s), step_s));
        unsigned long long qo = (unsigned long long)floor(__ddiv_rn(__dsub_rn(e.brightness, min_o), step_o));
        qs = qs < mq_s ? qs : mq_s;
        qo = qo < mq_o ? qo : mq_o;
        w |= ((unsigned long long)(e.match_x / T) << 22) | ((unsigned long long)(e.match_y / T) << 33) | ((unsigned long long)e.transform << 46) |
             (qs << 49) | (qo << 54);
    } else {
        w |= 1ull << 63;
    }
    out[i] = w;
}

__global__ void k_unpack(const unsigned long long* __restrict__ in, uint32_t n, uint32_t t_max, double min_s, double max_s, double min_o,
                         double max_o, int bits_s, int bits_o, int use_fma, fe_encode_item* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long w = in[i];
    const uint32_t level = (uint32_t)((w >> 44) & 3), T = t_max >> level;
    fe_encode_item e;
    e.x = (uint32_t)(w & 2047) * T; e.y = (uint32_t)((w >> 11) & 2047) * T; e.w = T; e.h = T;
    e.distance = 0.0; e.pad_ = 0;
    if (w >> 63) {
        e.distance = 100000.0; e.contrast = 0.0; e.brightness = 0.0; e.transform = 0; e.match_x = e.match_y = e.src_w = e.src_h = 0;
    } else {
        const double step_s = __ddiv_rn(fabs(__dsub_rn(max_s, min_s)), (double)(1 << bits_s));
        const double step_o = __ddiv_rn(fabs(__dsub_rn(max_o, min_o)), (double)(1 << bits_o));
        const double qs = (double)((w >> 49) & 31), qo = (double)((w >> 54) & 127);
        // Quantizer::value: quant * step + min + step / 2 (Quantizer.hpp:31-35)
        e.contrast = __dadd_rn(use_fma ? __fma_rn(qs, step_s, min_s) : __dadd_rn(__dmul_rn(qs, step_s), min_s), __dmul_rn(step_s, 0.5));
        e.brightness = __dadd_rn(use_fma ? __fma_rn(qo, step_o, min_o) : __dadd_rn(__dmul_rn(qo, step_o), min_o), __dmul_rn(step_o, 0.5));
        e.transform = (int32_t)((w >> 46) & 7);
        e.match_x = (uint32_t)((w >> 22) & 2047) * T; e.match_y = (uint32_t)((w >> 33) & 2047) * T;
        e.src_w = e.src_h = 2 * T;
    }
    out[i] = e;
}

// ---------------------------------------------------------------------------------------------
// colour path: ImageIO::rgb2yuv / yuv2rgb (image/ImageIO.cpp:40-57, 68-84).  fp64 with explicit roundings; fma = the contraction
// GCC applies under -march=native (middle product plain, first and last fused onto it: oracle/frac_oracle.c:fo_rgb2yuv).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint8_t clamp_u8_dev(double x) { return x < 0.0 ? 0 : (x > 255.0 ? 255 : (uint8_t)x); }
__device__ __forceinline__ double mix3(double c0, double r, double c1, double g, double c2, double b, int fma) {
    return fma ? __fma_rn(c2, b, __fma_rn(c0, r, __dmul_rn(c1, g))) : __dadd_rn(__dadd_rn(__dmul_rn(c0, r), __dmul_rn(c1, g)), __dmul_rn(c2, b));
}
// one thread per 2x2 cell: four luma values; the chroma of the cell is that of its LAST pixel (odd x, odd y), as the reference's
// overwriting loop leaves it
__global__ void k_rgb2yuv420(const uint8_t* __restrict__ rgb, uint32_t w, uint32_t h, uint32_t stride, uint8_t* __restrict__ yb, uint32_t ys,
                             uint8_t* __restrict__ ub, uint32_t us, uint8_t* __restrict__ vb, uint32_t vs, int fma) {
    const uint32_t cx = blockIdx.x * blockDim.x + threadIdx.x, cy = blockIdx.y * blockDim.y + threadIdx.y;
    if (cx >= w / 2 || cy >= h / 2) return;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
            const uint32_t x = 2 * cx + dx, y = 2 * cy + dy;
            const uint8_t* p = rgb + (size_t)y * stride + 3 * (size_t)x;
            const double r = p[0], g = p[1], b = p[2];
            yb[(size_t)y * ys + x] = clamp_u8_dev(mix3(0.299, r, 0.587, g, 0.114, b, fma));
            if (dx && dy) {
                ub[(size_t)cy * us + cx] = clamp_u8_dev(__dadd_rn(mix3(-0.169, r, -0.331, g, 0.499, b, fma), 128.0));
                vb[(size_t)cy * vs + cx] = clamp_u8_dev(__dadd_rn(mix3(0.499, r, -0.418, g, -0.0813, b, fma), 128.0));
            }
        }
}
__global__ void k_yuv420_to_rgb(const uint8_t* __restrict__ yb, uint32_t w, uint32_t h, uint32_t ys, const uint8_t* __restrict__ ub, uint32_t us,
                                const uint8_t* __restrict__ vb, uint32_t vs, uint8_t* __restrict__ rgb, uint32_t rgb_stride, int fma) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const double yp = yb[(size_t)y * ys + x], du = __dsub_rn((double)ub[(size_t)(y / 2) * us + x / 2], 128.0),
                 dv = __dsub_rn((double)vb[(size_t)(y / 2) * vs + x / 2], 128.0);
    uint8_t* p = rgb + 3 * ((size_t)y * rgb_stride + x);
    if (fma) {
        p[0] = clamp_u8_dev(__fma_rn(1.402, dv, yp));
        p[1] = clamp_u8_dev(__fma_rn(-0.714, dv, __fma_rn(-0.344, du, yp)));
        p[2] = clamp_u8_dev(__fma_rn(1.772, du, yp));
    } else {
        p[0] = clamp_u8_dev(__dadd_rn(yp, __dmul_rn(1.402, dv)));
        p[1] = clamp_u8_dev(__dsub_rn(__dsub_rn(yp, __dmul_rn(0.344, du)), __dmul_rn(0.714, dv)));
        p[2] = clamp_u8_dev(__dadd_rn(yp, __dmul_rn(1.772, du)));
    }
}

// ---------------------------------------------------------------------------------------------
// synthetic images: same integer formulas as oracle/frac_oracle.c:fo_synth_image
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    unsigned long long z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ uint32_t lattice(unsigned long long seed, unsigned long long i, unsigned long long j, unsigned long long o) {
    return (uint32_t)(splitmix64(seed ^ (o * 0xD6E8FEB86659FD93ull) ^ (i * 0x9E3779B97F4A7C15ull) ^ (j * 0xC2B2AE3D27D4EB4Full)) >> 56);
}
__global__ void k_synth(uint8_t* out, uint32_t w, uint32_t h, uint32_t stride, unsigned long long seed, int kind) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    uint32_t v;
    if (kind == 2) {
        v = (11u * x + 43u * y + 124u) % 256u;
    } else if (kind == 1) {
        v = (uint32_t)(splitmix64(seed ^ ((unsigned long long)x * 0x9E3779B97F4A7C15ull) ^ ((unsigned long long)y * 0xC2B2AE3D27D4EB4Full)) >> 56);
    } else {
        const uint32_t cell[4] = {64, 16, 4, 1}, wt[4] = {4, 2, 1, 1};
        uint32_t acc = 0;
        for (int o = 0; o < 4; ++o) {
            const uint32_t c = cell[o], i = x / c, j = y / c, fx = x % c, fy = y % c;
            const uint32_t vo = ((c - fx) * (c - fy) * lattice(seed, i, j, o) + fx * (c - fy) * lattice(seed, i + 1, j, o) +
                                 (c - fx) * fy * lattice(seed, i, j + 1, o) + fx * fy * lattice(seed, i + 1, j + 1, o)) / (c * c);
            acc += wt[o] * vo;
        }
        v = acc / 8;
    }
    out[(size_t)y * stride + x] = (uint8_t)v;
}
