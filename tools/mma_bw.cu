// Microbenchmark: tcgen05.mma (kind::f16, M = 128, K = 16, SMEM operands, no swizzle) issue/execute rate on sm_100a for the
// small-N shapes the conv kernels use, and what a change of shape or accumulator address between MMAs costs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_bw tools/mma_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0, laneid = 0;
    asm volatile("{\n.reg .b32 %%rx;\n.reg .pred %%px;\nelect.sync %%rx|%%px, %2;\n@%%px mov.s32 %1, 1;\nmov.s32 %0, %%rx;\n}\n"
                 : "+r"(laneid), "+r"(pred) : "r"(0xffffffffu));
    return pred != 0;
}
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__host__ __device__ constexpr uint32_t idesc(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}

// pattern p: which (N, D column, A offset) sequence one "round" of 15 MMAs uses
template <int P, int CONT>
__global__ void __launch_bounds__(192) bench(int rounds, long long *cycles, const uint8_t *gsrc) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar, lbar;
    __shared__ volatile int done;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 98304 / 16; i += blockDim.x) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) {
        done = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&lbar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp >= 1 && warp <= 4) {            // contention: TMEM loads from columns the MMAs do not touch
        if (CONT & 1) {
            const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + 448;
            uint32_t acc = 0;
            while (!done) {
                uint32_t r[16];
                for (int j = 0; j < 3; ++j)
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(base + 16 * j) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc ^= r[0];
            }
            if (acc == 0x1234567u) cycles[1] = acc;
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        return;
    }
    if (warp == 5) {                          // contention: bulk copies global -> shared into a region the MMAs do not read
        if (CONT & 2) {
            uint32_t ph = 0;
            while (!done) {
                if (elect_one()) {
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&lbar)), "r"(3 * 6144u) : "memory");
                    for (int j = 0; j < 3; ++j)
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     ::"r"(smem_u32(smem) + 73728 + j * 6144), "l"(gsrc + (size_t)((ph * 3 + j) % 512) * 6144), "r"(6144u), "r"(smem_u32(&lbar)) : "memory");
                }
                __syncwarp();
                uint32_t ok = 0;
                while (!ok) asm volatile("{\n.reg .pred P;\nmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(ok) : "r"(smem_u32(&lbar)), "r"(ph & 1) : "memory");
                ++ph;
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        return;
    }
    const uint32_t tm = slot;
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 65536;
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 2; ++rep) {
        t0 = clock64();
        if (elect_one()) {
            for (int r = 0; r < rounds; ++r) {
#pragma unroll
                for (int i = 0; i < 15; ++i) {
                    int n, dcol, aoff;
                    const int v = i % 5, ky = i / 5;
                    if (P == 0) { n = 144; dcol = 0; aoff = 0; }
                    else if (P == 1) { n = 96; dcol = 0; aoff = 0; }
                    else if (P == 2) { n = 48; dcol = 0; aoff = 0; }
                    else if (P == 3) { n = 256; dcol = 0; aoff = 0; }
                    else if (P == 4) { n = v == 0 ? 144 : (v < 3 ? 96 : 48); dcol = 0; aoff = 0; }                 // shapes alternate, same D
                    else if (P == 5) { n = 48; dcol = 48 * (i % 9); aoff = 0; }                                    // same shape, D moves
                    else if (P == 6) { n = 144; dcol = 144 * (i % 3); aoff = 0; }
                    else if (P == 7) { n = v == 0 ? 144 : (v < 3 ? 96 : 48); dcol = v == 0 ? 0 : (v == 1 ? 0 : v == 2 ? 48 : v == 3 ? 0 : 96); aoff = 6144 * v + 464 * ky; }   // the conv_mid round as issued today
                    else if (P == 8) { n = i < 3 ? 144 : (i < 9 ? 96 : 48); dcol = i < 3 ? 0 : (i < 9 ? 48 * (i & 1) : 96 * (i & 1)); aoff = 6144 * (i % 5) + 464 * (i % 3); }    // same 15 MMAs, sorted by shape
                    else { n = 144; dcol = 0; aoff = 6144 * v + 464 * ky; }                                        // N=144, A address moves
                    mma(tm + dcol, smem_desc(a0 + aoff, 3072, 128), smem_desc(b0 + ky * 6912, n * 16, 128), idesc(n), (r | i) ? 1u : 0u);
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        __syncwarp();
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n.reg .pred P;\nmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\nselp.u32 %0, 1, 0, P;\n}\n"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"((uint32_t)rep) : "memory");
        }
        t1 = clock64();
    }
    if (threadIdx.x == 0) { cycles[0] = t1 - t0; done = 1; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}

template <int P, int CONT = 0>
void run(const char *what, double ideal_per_round) {
    long long *d, h = 0;
    uint8_t *g;
    cudaMalloc(&d, 16);
    cudaMalloc(&g, 512 * 6144);
    cudaMemset(g, 0, 512 * 6144);
    cudaFuncSetAttribute(bench<P, CONT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304);
    const int rounds = 200;
    bench<P, CONT><<<1, 192, 98304>>>(rounds, d, g);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("P%d %-52s %8lld cycles  %6.1f per MMA  (math floor %.1f per MMA)  %s\n", P, what, h, (double)h / (rounds * 15),
           ideal_per_round / 15, cudaGetErrorString(e));
    cudaFree(d);
    cudaFree(g);
}

int main() {
    run<0>("N=144 x15, same D, same A", 15 * 72.0);
    run<1>("N=96 x15", 15 * 48.0);
    run<2>("N=48 x15", 15 * 24.0);
    run<3>("N=256 x15", 15 * 128.0);
    run<4>("shapes 144,96,96,48,48 repeating, same D/A", 3 * (72 + 48 + 48 + 24 + 24.0));
    run<5>("N=48, D column moves every MMA", 15 * 24.0);
    run<6>("N=144, D column moves every MMA", 15 * 72.0);
    run<7>("conv_mid round as issued (shape, D, A all move)", 3 * (72 + 48 + 48 + 24 + 24.0));
    run<8>("same MMAs sorted by shape", 3 * (72 + 48 + 48 + 24 + 24.0));
    run<9>("N=144, A address moves", 15 * 72.0);
    run<7, 1>("conv_mid round + 4 warps of TMEM loads", 3 * (72 + 48 + 48 + 24 + 24.0));
    run<7, 2>("conv_mid round + bulk copies into smem", 3 * (72 + 48 + 48 + 24 + 24.0));
    run<7, 3>("conv_mid round + both", 3 * (72 + 48 + 48 + 24 + 24.0));
    run<0, 3>("N=144 + both", 15 * 72.0);
    return 0;
}
