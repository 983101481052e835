// Experiment: A operand from TMEM.  tcgen05.cp (128x256b) copies a 128 x 16 fp16 K-major A tile from shared memory into 8 TMEM
// columns; tcgen05.mma then takes [a_tmem] instead of an A descriptor, so the tile is read from shared memory ONCE however many
// MMAs use it.  Checks the result against the shared-memory-operand MMA and against the CPU, then times the conv_mid MMA mix
// (45 MMAs over 25 distinct A views per K-step) both ways.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_ts tools/mma_ts.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_fp16.h>
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
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t id, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
                 ::"r"(d), "r"(a_tmem), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void cp_128x256b(uint32_t tmem, uint64_t desc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tmem), "l"(desc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n.reg .pred P;\nmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}

constexpr int N = 144, A_LBO = 2048, B_LBO = N * 16;
constexpr int A_COLS = 432;     // TMEM columns of the A buffers (8 per tile)

// mode 0: correctness (out[0..128*144) = SS result, then TS result).  mode 1/2: timing of the conv_mid mix, SS / TS.
__global__ void __launch_bounds__(128) kern(int mode, int rounds, const __half *a_g, const __half *b_g, float *out, long long *cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    uint8_t *sa = smem, *sb = smem + 65536;
    for (int i = threadIdx.x; i < 98304 / 16; i += blockDim.x) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    for (int i = threadIdx.x; i < 128 * 16; i += blockDim.x) {       // A[m][k] -> interleaved K-major
        const int m = i / 16, k = i % 16;
        *reinterpret_cast<__half *>(sa + (m / 8) * 128 + (k / 8) * A_LBO + (m % 8) * 16 + (k % 8) * 2) = a_g[i];
    }
    for (int i = threadIdx.x; i < N * 16; i += blockDim.x) {
        const int n = i / 16, k = i % 16;
        *reinterpret_cast<__half *>(sb + (n / 8) * 128 + (k / 8) * B_LBO + (n % 8) * 16 + (k % 8) * 2) = b_g[i];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = slot;
    const uint32_t a0 = smem_u32(sa), b0 = smem_u32(sb);
    if (mode == 0) {
        if (warp == 0) {
            if (elect_one()) {
                mma_ss(tm, smem_desc(a0, A_LBO, 128), smem_desc(b0, B_LBO, 128), idesc(N), 0u);
                cp_128x256b(tm + A_COLS, smem_desc(a0, A_LBO, 128));
                mma_ts(tm + N, tm + A_COLS, smem_desc(b0, B_LBO, 128), idesc(N), 0u);
                commit(&bar);
            }
            __syncwarp();
        }
        wait(&bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int c = 0; c < 2 * N + 8; c += 8) {      // both results, then the 8 columns of the A copy
            const int col = c < 2 * N ? c : A_COLS;
            uint32_t r[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                         : "r"(tm + ((uint32_t)(warp * 32) << 16) + col) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 8; ++j) out[(size_t)(warp * 32 + lane) * (2 * N + 8) + c + j] = __uint_as_float(r[j]);
        }
    } else if (warp == 0) {
        long long t0 = 0, t1 = 0;
        for (int rep = 0; rep < 2; ++rep) {
            t0 = clock64();
            if (elect_one()) {
                for (int r = 0; r < rounds; ++r) {
                    // one K-step of conv_mid: oy = -1..3, per oy the 5 ox views, each used by n_dy(oy) MMAs of N = 48 n_dx(ox)
                    int buf = 0;
#pragma unroll
                    for (int oy = -1; oy <= 3; ++oy) {
                        const int n_dy = oy == -1 || oy == 3 ? 1 : (oy == 1 ? 3 : 2);
                        if (mode == 2) {
#pragma unroll
                            for (int v = 0; v < 5; ++v)
                                cp_128x256b(tm + A_COLS + buf * 40 + v * 8, smem_desc(a0 + 6144 * v + 464 * (oy + 1), A_LBO, 128));
                        }
#pragma unroll
                        for (int d = 0; d < 3; ++d) {
                            if (d >= n_dy) continue;
#pragma unroll
                            for (int v = 0; v < 5; ++v) {
                                const int n = v == 0 ? 144 : (v < 3 ? 96 : 48);
                                const int dcol = 144 * d + (v == 2 || v == 4 ? 48 * (v == 2 ? 1 : 2) : 0);
                                const uint64_t db = smem_desc(b0 + (d % 3) * 6912 % 20736, n * 16, 128);
                                if (mode == 2) mma_ts(tm + dcol, tm + A_COLS + buf * 40 + v * 8, db, idesc(n), 1u);
                                else mma_ss(tm + dcol, smem_desc(a0 + 6144 * v + 464 * (oy + 1), A_LBO, 128), db, idesc(n), 1u);
                            }
                        }
                        buf ^= 1;
                    }
                }
                commit(&bar);
            }
            __syncwarp();
            wait(&bar, (uint32_t)rep);
            t1 = clock64();
        }
        if (lane == 0) cycles[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}

int main() {
    __half ha[128 * 16], hb[N * 16];
    srand(1);
    for (auto &v : ha) v = __float2half((float)(rand() % 17 - 8));
    for (auto &v : hb) v = __float2half((float)(rand() % 9 - 4) * 0.5f);
    __half *da, *db;
    float *dout;
    long long *dc, hc = 0;
    cudaMalloc(&da, sizeof(ha)); cudaMalloc(&db, sizeof(hb)); cudaMalloc(&dout, 128 * (2 * N + 8) * 4); cudaMalloc(&dc, 8);
    cudaMemcpy(da, ha, sizeof(ha), cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb, sizeof(hb), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304);
    kern<<<1, 128, 98304>>>(0, 0, da, db, dout, dc);
    cudaError_t e = cudaDeviceSynchronize();
    printf("correctness launch: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    static float out[128 * (2 * N + 8)];
    cudaMemcpy(out, dout, sizeof(out), cudaMemcpyDeviceToHost);
    int bad_ss = 0, bad_ts = 0, bad_cp = 0;
    for (int m = 0; m < 128; ++m) {
        for (int n = 0; n < N; ++n) {
            float ref = 0;
            for (int k = 0; k < 16; ++k) ref += __half2float(ha[m * 16 + k]) * __half2float(hb[n * 16 + k]);
            if (out[m * (2 * N + 8) + n] != ref) ++bad_ss;
            if (out[m * (2 * N + 8) + N + n] != ref) ++bad_ts;
        }
        for (int j = 0; j < 8; ++j) {               // A copy: lane m, column j = (A[m][2j], A[m][2j+1])
            uint32_t w; memcpy(&w, &out[m * (2 * N + 8) + 2 * N + j], 4);
            uint16_t lo, hi; memcpy(&lo, &ha[m * 16 + 2 * j], 2); memcpy(&hi, &ha[m * 16 + 2 * j + 1], 2);
            if (w != ((uint32_t)lo | ((uint32_t)hi << 16))) ++bad_cp;
        }
    }
    printf("mismatches vs CPU: smem-operand MMA %d, TMEM-operand MMA %d of %d; A copy (row m -> lane m, K pairs per column) %d of %d\n",
           bad_ss, bad_ts, 128 * N, bad_cp, 128 * 8);
    for (int mode = 1; mode <= 2; ++mode) {
        const int rounds = 200;
        kern<<<1, 128, 98304>>>(mode, rounds, da, db, dout, dc);
        e = cudaDeviceSynchronize();
        cudaMemcpy(&hc, dc, 8, cudaMemcpyDeviceToHost);
        printf("conv_mid K-step (45 MMAs, 25 views), A from %s: %.0f cycles per K-step  %s\n", mode == 1 ? "shared memory" : "TMEM (25 tcgen05.cp)",
               (double)hc / rounds, cudaGetErrorString(e));
    }
    return 0;
}
