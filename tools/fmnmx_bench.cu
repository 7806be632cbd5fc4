// FMNMX3 vs FMNMX vs integer min throughput/latency (cycles per warp-instruction), 1 CTA of 256 threads.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float fmin3(float a, float b, float c) { float r; asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float fmin2(float a, float b) { float r; asm volatile("min.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ int imin2(int a, int b) { int r; asm volatile("min.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__global__ void k(long long* res, float* sink, const float* in) {
    float v[64];
    for (int i = 0; i < 64; ++i) v[i] = in[(threadIdx.x + i * 7) & 1023];
    __syncthreads();
    float m0 = 1e30f, m1 = 1e30f, m2 = 1e30f, m3 = 1e30f;
    long long t0 = clock64();
#pragma unroll
    for (int r = 0; r < 16; ++r)
#pragma unroll
        for (int i = 0; i < 64; i += 8) {
            m0 = fmin3(m0, v[i], v[i + 1]); m1 = fmin3(m1, v[i + 2], v[i + 3]); m2 = fmin3(m2, v[i + 4], v[i + 5]); m3 = fmin3(m3, v[i + 6], v[i + 7]);
        }
    long long t1 = clock64();
#pragma unroll
    for (int r = 0; r < 16; ++r)
#pragma unroll
        for (int i = 0; i < 64; i += 4) {
            m0 = fmin2(m0, v[i]); m1 = fmin2(m1, v[i + 1]); m2 = fmin2(m2, v[i + 2]); m3 = fmin2(m3, v[i + 3]);
        }
    long long t2 = clock64();
    int a0 = __float_as_int(m0), a1 = __float_as_int(m1), a2 = __float_as_int(m2), a3 = __float_as_int(m3);
#pragma unroll
    for (int r = 0; r < 16; ++r)
#pragma unroll
        for (int i = 0; i < 64; i += 4) {
            a0 = imin2(a0, __float_as_int(v[i])); a1 = imin2(a1, __float_as_int(v[i + 1])); a2 = imin2(a2, __float_as_int(v[i + 2])); a3 = imin2(a3, __float_as_int(v[i + 3]));
        }
    long long t3 = clock64();
    if (threadIdx.x == 0) { res[0] = t1 - t0; res[1] = t2 - t1; res[2] = t3 - t2; }
    sink[threadIdx.x] = m0 + m1 + m2 + m3 + a0 + a1 + a2 + a3;
}
int main() {
    long long* d; float* s; float* in;
    cudaMalloc(&d, 64); cudaMalloc(&s, 4096); cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096);
    for (int threads : {32, 128, 256, 512}) {
        k<<<1, threads>>>(d, s, in); cudaDeviceSynchronize();
        long long h[3]; cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
        printf("threads=%d: 512 FMNMX3 (1024 elems) %lld cyc | 1024 FMNMX %lld cyc | 1024 IMNMX %lld cyc\n", threads, h[0], h[1], h[2]);
    }
    return 0;
}
