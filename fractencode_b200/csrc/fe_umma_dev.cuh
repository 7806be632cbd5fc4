// fe_umma_dev.cuh -- device-side PTX wrappers shared by the tcgen05 search kernels (fp16 and int8 kinds).
#pragma once
#include <cuda_fp16.h>

#include "fe_umma.cuh"

namespace umma_dev {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must not hang the GPU -- trap after ~2^26 probes (each probe suspends the warp in
// hardware for up to the mbarrier time limit) instead of spinning forever.  No clock reads in the loop: the
// spinning producer/issuer warps share their SM sub-partition's ALU pipe with the compute warps.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    uint32_t probes = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++probes > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
// plain shared-memory flags: ~30-cycle polls instead of ~200-cycle mbarrier probes
__device__ __forceinline__ void flag_add_release(uint32_t addr, uint32_t v) {
    asm volatile("red.release.cta.shared::cta.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void flag_store_release(uint32_t addr, uint32_t v) {
    asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t flag_load_acquire(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void flag_wait_ge(uint32_t addr, uint32_t want) {
    if ((int32_t)(flag_load_acquire(addr) - want) >= 0) return;
    const long long t0 = clock64();
    while ((int32_t)(flag_load_acquire(addr) - want) < 0) {
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
template <int KIND>  // 0: kind::f16, 1: kind::i8
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (KIND == 0) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}
// K-major, no swizzle: core matrices of 8 rows x 16 bytes; LBO = byte stride between the K chunks,
// SBO = byte stride between 8-row groups (cute/atom/mma_traits_sm100.hpp, LayoutType::INTERLEAVE).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
    const uint32_t hi = ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14); // version = 1 (Blackwell), layout_type = 0
    return ((uint64_t)hi << 32) | lo;
}

#define TMEM_LD32(taddr, v)                                                                                                   \
    asm volatile(                                                                                                             \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                             \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                             \
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                             \
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),         \
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), \
          "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),             \
          "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                                        \
        : "r"(taddr)                                                                                                          \
        : "memory")
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// T bytes at p as T/4 little-endian words
template <int T>
__device__ __forceinline__ void load_px(const uint8_t* __restrict__ p, uint32_t (&w)[T / 4]) {
    if ((reinterpret_cast<uintptr_t>(p) & (T - 1)) == 0) {
        if constexpr (T == 4) {
            w[0] = __ldg(reinterpret_cast<const uint32_t*>(p));
        } else {
            const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
            w[0] = v.x; w[1] = v.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < T / 4; ++i)
            w[i] = (uint32_t)p[4 * i] | ((uint32_t)p[4 * i + 1] << 8) | ((uint32_t)p[4 * i + 2] << 16) | ((uint32_t)p[4 * i + 3] << 24);
    }
}

} // namespace umma_dev

// Per-block second moments for the operand builders, one warp per block (sorted position p -> item order[p]).
// mode 0: range, sum (4 r - 510)^2   mode 1: range, 16 sum r^2   mode 2: domain, sum (D - 510)^2   mode 3: domain, sum D^2
__global__ void k_block_norms(const uint8_t* img, uint32_t stride, const fe_grid_item* items, const uint32_t* order, uint32_t n, uint32_t T,
                              int mode, uint32_t* out);

