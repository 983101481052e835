// Microbenchmark: HBM read throughput for the access pattern of the 720p -> 256x144 resize (rows 5y+2 of every frame:
// 3,840 contiguous bytes every 19,200) against a contiguous read of the same volume.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/hbm_rows tools/hbm_rows.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

// Each CTA takes (frame, row) items round-robin; a row is `row_bytes` contiguous bytes at frame*frame_stride + (off + y*step)*pitch.
__global__ void __launch_bounds__(256) read_rows(const uint8_t *base, long long frame_stride, int pitch, int off, int step, int rows,
                                                 int row_bytes, int frames, unsigned long long *sink) {
    uint4 acc = make_uint4(0, 0, 0, 0);
    const long long items = (long long)frames * rows;
    for (long long it = blockIdx.x; it < items; it += gridDim.x) {
        const int f = (int)(it / rows), y = (int)(it % rows);
        const uint4 *p = reinterpret_cast<const uint4 *>(base + f * frame_stride + (long long)(off + y * step) * pitch);
        for (int i = threadIdx.x; i < row_bytes / 16; i += blockDim.x) {
            const uint4 v = __ldcs(p + i);
            acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
        }
    }
    if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345678u) sink[0] = 1;
}

// One CTA per frame at a time (frames round-robin over CTAs), rows in order: the fused kernel's schedule.
__global__ void __launch_bounds__(256) read_frames(const uint8_t *base, long long frame_stride, int pitch, int off, int step, int rows,
                                                   int row_bytes, int frames, unsigned long long *sink) {
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (int f = blockIdx.x; f < frames; f += gridDim.x)
        for (int y = 0; y < rows; ++y) {
            const uint4 *p = reinterpret_cast<const uint4 *>(base + f * frame_stride + (long long)(off + y * step) * pitch);
            for (int i = threadIdx.x; i < row_bytes / 16; i += blockDim.x) {
                const uint4 v = __ldcs(p + i);
                acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
            }
        }
    if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345678u) sink[0] = 1;
}

// The fused kernel's row supply in isolation: 8 warps, each owning 2 of 16 shared-memory slots; a warp waits for its row,
// "processes" it for `proc` cycles, then issues the cp.async copies of the row 16 further on into the same slot.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(256) ring_loader(const uint8_t *base, long long frame_stride, int pitch, int off, int step, int rows,
                                                   int frames, int proc, long long *lat_sum, unsigned long long *sink) {
    extern __shared__ __align__(128) uint8_t raw[];            // 16 x 3840
    __shared__ uint64_t full[16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < 16) asm volatile("mbarrier.init.shared::cta.b64 [%0], 32;" ::"r"(smem_u32(&full[threadIdx.x])));
    __syncthreads();
    const int n_frames_cta = (frames - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total = n_frames_cta * rows;
    auto issue = [&](int n) {
        if (n >= total) return;
        const int fi = n / rows, y = n - fi * rows, slot = n & 15;
        const uint8_t *g = base + (long long)(blockIdx.x + (long long)fi * gridDim.x) * frame_stride + (long long)(off + y * step) * pitch;
        for (int c = lane * 16; c < 3840; c += 512)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(raw + slot * 3840 + c)), "l"(g + c) : "memory");
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full[slot])) : "memory");
    };
    long long t_issue[2] = {clock64(), clock64()}, lat = 0;
    uint32_t acc = 0;
    issue(warp); issue(warp + 8);
    for (int n = warp, k = 0; n < total; n += 8, ++k) {
        const int slot = n & 15;
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n.reg .pred P;\nmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(ok) : "r"(smem_u32(&full[slot])), "r"((uint32_t)((n >> 4) & 1)) : "memory");
        lat += clock64() - t_issue[k & 1];
        acc ^= reinterpret_cast<const uint32_t *>(raw + slot * 3840)[lane];
        const long long t0 = clock64();
        while (clock64() - t0 < proc) { }
        __syncwarp();
        t_issue[k & 1] = clock64();
        issue(n + 16);
    }
    if (lane == 0) atomicAdd((unsigned long long *)lat_sum, (unsigned long long)lat);
    if (acc == 0x12345678u) sink[0] = 1;
}

// Same ring, but rows arrive by cp.async.bulk (one elected thread of a dedicated warp issues them; full/empty mbarriers).
__global__ void __launch_bounds__(288) ring_loader_bulk(const uint8_t *base, long long frame_stride, int pitch, int off, int step, int rows,
                                                        int frames, int proc, int n_slots, long long *lat_sum, unsigned long long *sink) {
    extern __shared__ __align__(128) uint8_t raw[];            // n_slots x 3840
    __shared__ uint64_t full[64], empty[64];
    __shared__ long long t_issue[64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < n_slots) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[threadIdx.x])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&empty[threadIdx.x])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const int n_frames_cta = (frames - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total = n_frames_cta * rows;
    auto wait = [&](uint64_t *bar, uint32_t parity) {
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n.reg .pred P;\nmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    };
    if (warp == 8) {
        if (lane == 0)
            for (int n = 0; n < total; ++n) {
                const int slot = n % n_slots, use = n / n_slots;
                wait(&empty[slot], (use & 1) ^ 1);
                const int fi = n / rows, y = n - fi * rows;
                const uint8_t *g = base + (long long)(blockIdx.x + (long long)fi * gridDim.x) * frame_stride + (long long)(off + y * step) * pitch;
                t_issue[slot] = clock64();
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[slot])), "r"(3840) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(raw + slot * 3840)), "l"(g), "r"(3840), "r"(smem_u32(&full[slot])) : "memory");
            }
        return;
    }
    long long lat = 0;
    uint32_t acc = 0;
    for (int n = warp; n < total; n += 8) {
        const int slot = n % n_slots, use = n / n_slots;
        wait(&full[slot], use & 1);
        lat += clock64() - *(volatile long long *)&t_issue[slot];
        acc ^= reinterpret_cast<const uint32_t *>(raw + slot * 3840)[lane];
        const long long t0 = clock64();
        while (clock64() - t0 < proc) { }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[slot])) : "memory");
    }
    if (lane == 0) atomicAdd((unsigned long long *)lat_sum, (unsigned long long)lat);
    if (acc == 0x12345678u) sink[0] = 1;
}

// cp.async.bulk again, with cheap index arithmetic (power-of-two slots, rows == 144 folded) and `prod` issuing warps
// (warp 8 + k issues rows n = k mod prod), to separate the copy engine's throughput from the issue loop's.
__global__ void __launch_bounds__(384) ring_loader_bulk2(const uint8_t *base, long long frame_stride, int pitch, int off, int step,
                                                         int frames, int proc, int log2_slots, int prod, int split,
                                                         unsigned long long *sink) {
    extern __shared__ __align__(128) uint8_t raw[];
    __shared__ uint64_t full[64], empty[64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_slots = 1 << log2_slots, rows = 144;
    if (threadIdx.x < n_slots) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[threadIdx.x])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&empty[threadIdx.x])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const int n_frames_cta = (frames - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total = n_frames_cta * rows;
    auto wait = [&](uint64_t *bar, uint32_t parity) {
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n.reg .pred P;\nmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    };
    if (warp >= 8) {
        const int k = warp - 8;
        if (k < prod && lane == 0) {
            int fi = 0, y = k;
            const int piece = 3840 / split;
            for (int n = k; n < total; n += prod, y += prod) {
                if (y >= rows) { y -= rows; ++fi; }
                const int slot = n & (n_slots - 1);
                wait(&empty[slot], ((n >> log2_slots) & 1) ^ 1);
                const uint8_t *g = base + (long long)(blockIdx.x + (long long)fi * gridDim.x) * frame_stride + (long long)(off + y * step) * pitch;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[slot])), "r"(3840) : "memory");
                for (int q = 0; q < split; ++q)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(raw + slot * 3840 + q * piece)), "l"(g + q * piece), "r"(piece), "r"(smem_u32(&full[slot])) : "memory");
            }
        }
        return;
    }
    uint32_t acc = 0;
    for (int n = warp; n < total; n += 8) {
        const int slot = n & (n_slots - 1);
        wait(&full[slot], (n >> log2_slots) & 1);
        acc ^= reinterpret_cast<const uint32_t *>(raw + slot * 3840)[lane];
        const long long t0 = clock64();
        while (clock64() - t0 < proc) { }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[slot])) : "memory");
    }
    if (acc == 0x12345678u) sink[0] = 1;
}

// cp.async (16 B per lane) from `prod` dedicated loader warps; 8 consumer warps as above.
__global__ void __launch_bounds__(768) ring_loader_ldgsts(const uint8_t *base, long long frame_stride, int pitch, int off, int step,
                                                          int frames, int proc, int log2_slots, int prod, unsigned long long *sink) {
    extern __shared__ __align__(128) uint8_t raw[];
    __shared__ uint64_t full[64], empty[64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_slots = 1 << log2_slots, rows = 144;
    if (threadIdx.x < n_slots) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 32;" ::"r"(smem_u32(&full[threadIdx.x])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&empty[threadIdx.x])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const int n_frames_cta = (frames - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total = n_frames_cta * rows;
    auto wait = [&](uint64_t *bar, uint32_t parity) {
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n.reg .pred P;\nmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    };
    if (warp >= 8) {
        const int k = warp - 8;
        if (k >= prod) return;
        int fi = 0, y = k;
        for (int n = k; n < total; n += prod, y += prod) {
            if (y >= rows) { y -= rows; ++fi; }
            const int slot = n & (n_slots - 1);
            wait(&empty[slot], ((n >> log2_slots) & 1) ^ 1);
            const uint8_t *g = base + (long long)(blockIdx.x + (long long)fi * gridDim.x) * frame_stride + (long long)(off + y * step) * pitch;
            for (int c = lane * 16; c < 3840; c += 512)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(raw + slot * 3840 + c)), "l"(g + c) : "memory");
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full[slot])) : "memory");
        }
        return;
    }
    uint32_t acc = 0;
    for (int n = warp; n < total; n += 8) {
        const int slot = n & (n_slots - 1);
        wait(&full[slot], (n >> log2_slots) & 1);
        acc ^= reinterpret_cast<const uint32_t *>(raw + slot * 3840)[lane];
        const long long t0 = clock64();
        while (clock64() - t0 < proc) { }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[slot])) : "memory");
    }
    if (acc == 0x12345678u) sink[0] = 1;
}

template <typename F>
double time_ms(F launch) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    launch();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    const int frames = 1184, h = 720, pitch = 3840;
    const long long frame_stride = (long long)h * pitch;
    uint8_t *d;
    unsigned long long *sink;
    if (cudaMalloc(&d, frames * frame_stride) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMalloc(&sink, 8);
    cudaMemset(d, 1, frames * frame_stride);
    const double sparse_bytes = (double)frames * 144 * 3840;
    for (int grid : {148, 592, 2368, 9472}) {
        double ms = time_ms([&] { read_rows<<<grid, 256>>>(d, frame_stride, pitch, 2, 5, 144, 3840, frames, sink); });
        printf("rows 5y+2 (3840 B every 19200), items round-robin, grid %5d: %7.3f ms  %7.1f GB/s\n", grid, ms, sparse_bytes / ms / 1e6);
    }
    for (int grid : {148, 592}) {
        double ms = time_ms([&] { read_frames<<<grid, 256>>>(d, frame_stride, pitch, 2, 5, 144, 3840, frames, sink); });
        printf("rows 5y+2, one frame per CTA at a time,        grid %5d: %7.3f ms  %7.1f GB/s\n", grid, ms, sparse_bytes / ms / 1e6);
    }
    for (int grid : {592, 9472}) {     // contiguous: the same number of bytes, rows back to back
        double ms = time_ms([&] { read_rows<<<grid, 256>>>(d, (long long)144 * 3840, pitch, 0, 1, 144, 3840, frames, sink); });
        printf("contiguous, same volume,                         grid %5d: %7.3f ms  %7.1f GB/s\n", grid, ms, sparse_bytes / ms / 1e6);
    }
    {   // every row of every frame (what a non-sparse K1 would read)
        double ms = time_ms([&] { read_rows<<<9472, 256>>>(d, frame_stride, pitch, 0, 1, 720, 3840, frames, sink); });
        printf("all 720 rows,                                    grid  9472: %7.3f ms  %7.1f GB/s\n", ms, (double)frames * frame_stride / ms / 1e6);
    }
    long long *lat;
    cudaMalloc(&lat, 8);
    cudaFuncSetAttribute(ring_loader, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 3840);
    for (int proc : {0, 1700, 3000}) {
        const int fr = 148 * 4;
        cudaMemset(lat, 0, 8);
        double ms = time_ms([&] { cudaMemsetAsync(lat, 0, 8); ring_loader<<<148, 256, 16 * 3840>>>(d, frame_stride, pitch, 2, 5, 144, fr, proc, lat, sink); });
        long long h = 0;
        cudaMemcpy(&h, lat, 8, cudaMemcpyDeviceToHost);
        printf("ring loader alone (16 slots/SM, cp.async), proc %4d cycles: %7.3f ms  %7.1f GB/s  mean issue->arrival %.0f cycles\n", proc, ms,
               (double)fr * 144 * 3840 / ms / 1e6, (double)h / ((double)fr * 144));
    }
    cudaFuncSetAttribute(ring_loader_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 3840);
    for (int n_slots : {8, 16, 24, 32, 48})
        for (int proc : {0, 1000, 1700}) {
            const int fr = 148 * 4;
            double ms = time_ms([&] { cudaMemsetAsync(lat, 0, 8); ring_loader_bulk<<<148, 288, 48 * 3840>>>(d, frame_stride, pitch, 2, 5, 144, fr, proc, n_slots, lat, sink); });
            long long h = 0;
            cudaMemcpy(&h, lat, 8, cudaMemcpyDeviceToHost);
            printf("ring loader, cp.async.bulk, %2d slots/SM, proc %4d cycles: %7.3f ms  %7.1f GB/s  mean issue->seen %.0f cycles\n", n_slots, proc, ms,
                   (double)fr * 144 * 3840 / ms / 1e6, (double)h / ((double)fr * 144));
        }
    cudaFuncSetAttribute(ring_loader_bulk2, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 3840);
    cudaFuncSetAttribute(ring_loader_ldgsts, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 3840);
    for (int log2_slots : {4, 5})
        for (int prod : {1, 2, 4})
            for (int split : {1, 2}) {
                const int fr = 148 * 4, proc = 1000;
                double ms = time_ms([&] { ring_loader_bulk2<<<148, 384, 32 * 3840>>>(d, frame_stride, pitch, 2, 5, fr, proc, log2_slots, prod, split, sink); });
                printf("bulk2: %2d slots/SM, %d issuing warps, %d copies per row, proc %d: %7.3f ms  %7.1f GB/s\n", 1 << log2_slots, prod, split, proc, ms,
                       (double)fr * 144 * 3840 / ms / 1e6);
            }
    for (int log2_slots : {4, 5})
        for (int prod : {2, 4, 8, 16})
            for (int proc : {0, 1000}) {
                const int fr = 148 * 4;
                double ms = time_ms([&] { ring_loader_ldgsts<<<148, 768, 32 * 3840>>>(d, frame_stride, pitch, 2, 5, fr, proc, log2_slots, prod, sink); });
                printf("ldgsts: %2d slots/SM, %2d loader warps, proc %4d: %7.3f ms  %7.1f GB/s\n", 1 << log2_slots, prod, proc, ms,
                       (double)fr * 144 * 3840 / ms / 1e6);
            }
    return 0;
}
