// fe_kernels.cuh -- kernel declarations shared by fe_kernels.cu and fe_api.cu.
#pragma once
#include "fe_internal.cuh"

struct FinalizeArgs {
    const uint8_t* src; uint32_t src_stride;
    const uint8_t* tgt; uint32_t tgt_stride;
    const fe_grid_item* dom;        // original (unsorted) domain list
    const fe_grid_item* rng;        // original range list
    const uint32_t* dom_order;      // column -> domain index (NULL: identity)
    const uint32_t* rng_order;      // range position -> range index (NULL: identity)
    const unsigned long long* rowbest;
    const uint32_t* rowhit;
    uint32_t n;                     // ranges
    uint32_t use_thr;
    double thr, s_max;
    int fma;
    int can_split;                  // quadtree: T > t_min
    uint32_t thr16;
    fe_encode_item* out;            // [range index]
    uint32_t* split;                // NULL or per-range-index split flag
    uint32_t* mismatch;             // device counter: search n16 != recomputed n16
    uint32_t* fp32_regime;          // device counter: winners with SSE >= 2^20
    uint32_t* bound_out;            // NULL or [range position]: 0, or the n16 bound of the re-rank band for fp32-regime ranges
    int rerank;                     // rowbest keys hold float bits (re-rank pass)
    int no_min;                     // the minimum is not wanted: a range without a hit gets the default item (and splits)
    int hit_is_domain;              // rowhit holds domain indices (tcgen05 paths), not sorted column positions (exact path)
    int flips;                      // eight isometries: rng holds every block twice, n counts blocks, pos_of[copy] = its row position
    const uint32_t* pos_of;
};

__global__ void k_uniform_grid(fe_grid_item* out, uint32_t nx, uint32_t n, uint32_t size, uint32_t step, uint32_t first);
__global__ void k_quadtree_scatter(const fe_grid_item* rng, const fe_encode_item* level_items, const uint32_t* split,
                                   const uint32_t* scan, uint32_t n, fe_grid_item* next, fe_encode_item* items_out);
__global__ void k_classify(const uint8_t* img, uint32_t stride, const fe_grid_item* items, uint32_t n, int32_t* cls, int force);
bool cell_grid_supported(const uint8_t* img, uint32_t stride, uint32_t w, uint32_t h, uint32_t C);
void launch_cell_grid(cudaStream_t stream, const uint8_t* img, uint32_t stride, uint32_t w, uint32_t h, uint32_t C, uint32_t* cells, uint32_t* cells2,
                      uint32_t* cellsD2);
void launch_dom_norms_from_cells(cudaStream_t stream, const uint32_t* cellsD2, uint32_t cw, uint32_t dnx, const uint32_t* order, uint32_t n, uint32_t* out);
void launch_dom_from_cells(cudaStream_t stream, const uint32_t* cells, uint32_t cw, uint32_t dnx, uint32_t n, int32_t* cls, uint32_t width, uint8_t* keys,
                           uint32_t* hist);
void launch_classify(cudaStream_t stream, const uint8_t* img, uint32_t stride, const fe_grid_item* items, uint32_t n, uint32_t edge, int32_t* cls, int force);
__global__ void k_fill_u32(uint32_t* p, uint32_t v, size_t n);
__global__ void k_fill_u64(unsigned long long* p, unsigned long long v, size_t n);
__global__ void k_iota(uint32_t* p, uint32_t n);
__global__ void k_dup_items(const fe_grid_item* in, uint32_t n, fe_grid_item* out);
__global__ void k_pos_of(const uint32_t* order, uint32_t n, uint32_t* pos_of);
struct BucketOff { uint32_t v[FE_MAX_BUCKETS + 1]; };
void launch_brightness_bins(cudaStream_t stream, const uint8_t* img, uint32_t stride, const fe_grid_item* items, uint32_t n, uint32_t edge,
                            uint32_t mul, uint32_t width, uint8_t* keys, uint32_t* hist);
__global__ void k_bin_prefix(const uint32_t* dom_order, const uint32_t* hist, int nb, BucketOff cut, uint32_t* out);
__global__ void k_class_keys(const int32_t* cls, uint32_t n, uint8_t* keys, uint32_t* hist);
__global__ void k_build_rows(const uint8_t* img, uint32_t stride, const fe_grid_item* rng, const uint32_t* order, uint32_t n,
                             uint32_t T, uint32_t Npad, int fast, uint8_t* A, uint32_t* rowc);
__global__ void k_build_pool(const uint8_t* img, uint32_t stride, const fe_grid_item* dom, const uint32_t* order, uint32_t n,
                             uint32_t npool, uint32_t T, uint32_t rho, uint32_t Npad, uint8_t* Blo, uint8_t* Bhi, uint32_t* coln);
void launch_k_finalize(cudaStream_t stream, const FinalizeArgs& f, uint32_t T);
__global__ void k_decode_step(const uint8_t* src, uint8_t* dst, uint32_t stride, const fe_encode_item* items,
                              const uint32_t* pix_off, uint32_t n_items, uint32_t total_pix, int use_fma, const uint32_t* done);
__global__ void k_decode_step_uniform(const uint8_t* src, uint8_t* dst, uint32_t stride, const fe_encode_item* items,
                                      uint32_t n_items, uint32_t T, int use_fma, unsigned long long* sq_out, const uint32_t* done);
bool launch_decode_step_small(cudaStream_t stream, const uint8_t* src, uint8_t* dst, uint32_t stride, const fe_encode_item* items, uint32_t n,
                              uint32_t T, int use_fma, unsigned long long* sq_out, const uint32_t* done, const uint16_t* dq_src, uint16_t* dq_dst,
                              uint32_t dq_stride);
void launch_boxsum_plane(cudaStream_t stream, const uint8_t* src, uint32_t stride, uint32_t w, uint32_t h, uint16_t* dq);
__global__ void k_decode_step_tiled(const uint8_t* src, uint8_t* dst, uint32_t stride, const fe_encode_item* items, uint32_t n_items,
                                    uint32_t T, int use_fma, unsigned long long* sq_out, const uint32_t* done);
__global__ void k_sqdiff(const uint8_t* a, const uint8_t* b, uint32_t w, uint32_t h, uint32_t stride, unsigned long long* out, const uint32_t* done);
__global__ void k_cover_bitmap(const fe_encode_item* items, uint32_t n, const uint32_t* row_off, uint32_t total_rows, uint32_t wpr, uint32_t* bitmap,
                               uint32_t* overlap);
__global__ void k_popcount(const uint32_t* words, size_t n, unsigned long long* out);
__global__ void k_decode_check(unsigned long long* sums, uint32_t* state, uint32_t npix, double eps, uint32_t iter);
__global__ void k_copy_plane_if_running(const uint4* src, uint4* dst, size_t n16, const uint32_t* done);
__global__ void k_minmax(const fe_encode_item* items, uint32_t n, unsigned long long* mm);
__global__ void k_quantize(const fe_encode_item* items, uint32_t n, double min_s, double max_s, double min_o, double max_o,
                           int bits_s, int bits_o, uint32_t* qs, uint32_t* qo);
__global__ void k_rgb2yuv420(const uint8_t* rgb, uint32_t w, uint32_t h, uint32_t stride, uint8_t* yb, uint32_t ys, uint8_t* ub, uint32_t us,
                             uint8_t* vb, uint32_t vs, int fma);
__global__ void k_yuv420_to_rgb(const uint8_t* yb, uint32_t w, uint32_t h, uint32_t ys, const uint8_t* ub, uint32_t us, const uint8_t* vb, uint32_t vs,
                                uint8_t* rgb, uint32_t rgb_stride, int fma);
__global__ void k_synth(uint8_t* out, uint32_t w, uint32_t h, uint32_t stride, unsigned long long seed, int kind);
__global__ void k_pack(const fe_encode_item* items, uint32_t n, uint32_t t_max, const double* mm, int bits_s, int bits_o,
                       unsigned long long* out, uint32_t* bad);
__global__ void k_minmax_finish(const unsigned long long* keys, double* out);
__global__ void k_unpack(const unsigned long long* in, uint32_t n, uint32_t t_max, double min_s, double max_s, double min_o, double max_o,
                         int bits_s, int bits_o, int use_fma, fe_encode_item* out);
