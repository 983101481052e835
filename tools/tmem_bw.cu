// Microbenchmark: TMEM -> register load throughput on sm_100a (tcgen05.ld), per shape and warp count.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_bw tools/tmem_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int N>
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t (&r)[64]);

template <>
__device__ __forceinline__ void ld<16>(uint32_t taddr, uint32_t (&r)[64]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
template <>
__device__ __forceinline__ void ld<32>(uint32_t taddr, uint32_t (&r)[64]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
template <>
__device__ __forceinline__ void ld<8>(uint32_t taddr, uint32_t (&r)[64]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}

// N columns per load, `per_wait` loads between waits, `iters` rounds.  Each warp reads its own lane quarter (warp % 4).
template <int N>
__global__ void bench(int iters, int per_wait, long long *cycles, uint32_t *sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t r[64];
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        for (int j = 0; j < per_wait; ++j) {
            ld<N>(base + ((it * per_wait + j) * N) % (512 - N), r);
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc ^= r[0] ^ r[N - 1];
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) sink[threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}

template <int N>
void run(int warps, int per_wait) {
    long long *d_c, h_c;
    uint32_t *d_s;
    cudaMalloc(&d_c, 8 * 148);
    cudaMalloc(&d_s, 4096);
    const int iters = 2000;
    bench<N><<<1, warps * 32>>>(iters, per_wait, d_c, d_s);
    bench<N><<<1, warps * 32>>>(iters, per_wait, d_c, d_s);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(&h_c, d_c, 8, cudaMemcpyDeviceToHost);
    const double bytes = (double)iters * per_wait * N * 4 * 32 * warps;
    printf("x%-2d warps=%2d per_wait=%d: %lld cycles, %.1f B/clk/SM, %.1f B/clk/warp, %.1f cycles/load  (%s)\n", N, warps, per_wait,
           h_c, bytes / h_c, bytes / h_c / warps, (double)h_c / (iters * per_wait), cudaGetErrorString(e));
    cudaFree(d_c);
    cudaFree(d_s);
}

int main() {
    for (int warps : {1, 2, 4, 8, 16}) {
        run<16>(warps, 1);
        run<16>(warps, 3);
        run<16>(warps, 6);
        run<32>(warps, 3);
        run<8>(warps, 6);
    }
    return 0;
}
