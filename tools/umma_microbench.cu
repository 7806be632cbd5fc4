// Micro-benchmarks of the tcgen05 / mbarrier primitives the search kernel is built from (cycles, one CTA).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/umb tools/umma_microbench.cu && /tmp/umb
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) { if (clock64() - t0 > 200000000ll) __trap(); }
}
__device__ __forceinline__ void tc_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (((lbo >> 4) & 0x3FFFu) << 16);
    const uint32_t hi = ((sbo >> 4) & 0x3FFFu) | (1u << 14);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ void tc_mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
#define TMEM_LD32(taddr, v)                                                                                                   \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                    \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                             \
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                             \
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),         \
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), \
          "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),             \
          "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                                        \
        : "r"(taddr) : "memory")
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// results[k]: average cycles of experiment k
__global__ void __launch_bounds__(160, 1) k_bench(long long* results, uint32_t* sink, int N, int nk, int reps) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bars[8];
    __shared__ uint32_t tmem_slot;
    const uint32_t warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u; // fp16 1.0
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bars[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t b0 = smem_u32(&bars[0]);
    if (threadIdx.x == 0) {
        // (0) try_wait on an already-complete phase
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) mbar_wait(b0, 1);
        results[0] = (clock64() - t0) / reps;
        // (1) commit with nothing pending -> wait
        uint32_t ph = 0;
        t0 = clock64();
        for (int r = 0; r < reps; ++r) { tc_commit(b0); mbar_wait(b0, ph); ph ^= 1; }
        results[1] = (clock64() - t0) / reps;
        // (2) nk MMAs (M128 x N x 16) + commit -> wait   (latency of one tile)
        const uint64_t ad = make_desc(smem_u32(smem), 2048, 128), bd = make_desc(smem_u32(smem) + 8192, N * 16, 128);
        t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            for (int k = 0; k < nk; ++k) tc_mma(tmem, ad, bd, idesc, k > 0);
            tc_commit(b0); mbar_wait(b0, ph); ph ^= 1;
        }
        results[2] = (clock64() - t0) / reps;
        // (3) throughput: 64 tiles back-to-back on alternating accumulators, one commit at the end
        t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            for (int t = 0; t < 64; ++t)
                for (int k = 0; k < nk; ++k) tc_mma(tmem + (t & 3) * 128, ad, bd, idesc, k > 0);
            tc_commit(b0); mbar_wait(b0, ph); ph ^= 1;
        }
        results[3] = (clock64() - t0) / reps / 64;
        // (4) issue cost alone of one tile (nk MMAs + commit), no wait
        t0 = clock64();
        const uint32_t b1 = smem_u32(&bars[1]);
        for (int r = 0; r < reps; ++r) {
            for (int k = 0; k < nk; ++k) tc_mma(tmem, ad, bd, idesc, k > 0);
            tc_commit(b1);  // bars[1] is never waited on
        }
        results[4] = (clock64() - t0) / reps;
        tc_commit(b0); mbar_wait(b0, ph); ph ^= 1;
        // (5) plain arrive -> wait
        t0 = clock64();
        for (int r = 0; r < reps; ++r) { mbar_arrive(b0); mbar_wait(b0, ph); ph ^= 1; }
        results[5] = (clock64() - t0) / reps;
    }
    __syncthreads();
    tc_fence_after();
    if (warp >= 1) {
        // (6) one warpgroup: 4 x LDTM.x32 + wait (128 columns), all 4 warps concurrently
        const uint32_t sp = warp & 3;
        const uint32_t taddr = tmem + ((sp * 32u) << 16);
        uint32_t v[128];
        uint32_t acc = 0;
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            TMEM_LD32(taddr, (v + 0)); TMEM_LD32(taddr + 32, (v + 32)); TMEM_LD32(taddr + 64, (v + 64)); TMEM_LD32(taddr + 96, (v + 96));
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 128; ++i) acc ^= v[i];
        }
        long long dt = (clock64() - t0) / reps;
        if (threadIdx.x == 32) results[6] = dt;
        // (7) single LDTM.x32 + wait
        t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            TMEM_LD32(taddr, (v + 0));
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) acc ^= v[i];
        }
        dt = (clock64() - t0) / reps;
        if (threadIdx.x == 32) results[7] = dt;
        // (8) bar.sync of 128 threads
        t0 = clock64();
        for (int r = 0; r < reps; ++r) asm volatile("bar.sync 1, 128;" ::: "memory");
        dt = (clock64() - t0) / reps;
        if (threadIdx.x == 32) results[8] = dt;
        sink[threadIdx.x] = acc;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
    long long* d_res; uint32_t* d_sink;
    cudaMalloc(&d_res, 16 * sizeof(long long)); cudaMalloc(&d_sink, 1024 * 4);
    cudaFuncSetAttribute(k_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const char* names[] = {"try_wait (complete)", "commit(empty)->wait", "tile MMAs+commit->wait (latency)", "tile MMAs throughput (per tile)",
                           "tile issue cost (no wait)", "arrive->wait", "WG 4xLDTM.x32+wait (128 cols)", "LDTM.x32+wait", "bar.sync 128"};
    for (int N : {128, 256}) for (int nk : {2, 5}) {
        cudaMemset(d_res, 0, 16 * sizeof(long long));
        k_bench<<<1, 160, 64 * 1024>>>(d_res, d_sink, N, nk, 200);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long h[16]; cudaMemcpy(h, d_res, sizeof(h), cudaMemcpyDeviceToHost);
        printf("N=%d nk=%d:\n", N, nk);
        for (int i = 0; i < 9; ++i) printf("  %-36s %6lld cycles\n", names[i], h[i]);
    }
    return 0;
}
