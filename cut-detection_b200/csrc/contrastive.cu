// Forward pass of the reference's ContrastiveLoss (frameID/metrics.py:9-47, the NT-Xent objective of learn_contrasts.py:108):
//   x [2B, D] -> (h_norm) L2-normalise the rows -> h1 = x[:B], h2 = x[B:]
//   logits_ab = h1 h2^T / T, logits_aa = h1 h1^T / T - 1e9 I, logits_bb = h2 h2^T / T - 1e9 I, logits_ba = logits_ab^T
//   loss = mean_i( CE([ab | aa]_i, i) + CE([ba | bb]_i, i) )
// One block per row i: the four rows of logits it needs are 4B dot products of D-vectors (D = 8, B = 32 in the reference),
// two log-sum-exps by warp shuffles; a second one-block kernel sums the B row losses in a fixed order.
#include "common.cuh"

namespace cutdet {

namespace {

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

constexpr int CL_THREADS = 128;

// scale[r] = 1 / max(||x_r||, 1e-12) (F.normalize), or 1 without h_norm
__global__ void row_scale_kernel(const float *__restrict__ x, int rows, int d, int h_norm, float *__restrict__ scale) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    float s = 1.f;
    if (h_norm) {
        float ss = 0.f;
        for (int k = 0; k < d; ++k) ss = fmaf(x[(int64_t)r * d + k], x[(int64_t)r * d + k], ss);
        s = 1.f / fmaxf(sqrtf(ss), 1e-12f);
    }
    scale[r] = s;
}

__global__ void __launch_bounds__(CL_THREADS) contrastive_rows_kernel(const float *__restrict__ x, const float *__restrict__ scale,
                                                                      int B, int d, float inv_t, float *__restrict__ logits_ab,
                                                                      float *__restrict__ row_loss) {
    extern __shared__ float sm[];            // [2][2B]: the concatenated logits of row i for loss_a and loss_b
    float *la = sm, *lb = sm + 2 * B;
    __shared__ float red[2][CL_THREADS / 32];
    const int i = blockIdx.x;
    const float *a_i = x + (int64_t)i * d, *b_i = x + (int64_t)(B + i) * d;
    const float sa = scale[i], sb = scale[B + i];
    for (int j = threadIdx.x; j < B; j += blockDim.x) {
        const float *a_j = x + (int64_t)j * d, *b_j = x + (int64_t)(B + j) * d;
        float ab = 0.f, aa = 0.f, ba = 0.f, bb = 0.f;
        for (int k = 0; k < d; ++k) {
            ab = fmaf(a_i[k], b_j[k], ab); aa = fmaf(a_i[k], a_j[k], aa);
            ba = fmaf(b_i[k], a_j[k], ba); bb = fmaf(b_i[k], b_j[k], bb);
        }
        const float sja = scale[j], sjb = scale[B + j];
        ab *= sa * sjb * inv_t; aa *= sa * sja * inv_t; ba *= sb * sja * inv_t; bb *= sb * sjb * inv_t;
        if (j == i) { aa -= 1e9f; bb -= 1e9f; }
        la[j] = ab; la[B + j] = aa;
        lb[j] = ba; lb[B + j] = bb;
        if (logits_ab) logits_ab[(int64_t)i * B + j] = ab;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float ma = -INFINITY, mb = -INFINITY;
    for (int j = threadIdx.x; j < 2 * B; j += blockDim.x) { ma = fmaxf(ma, la[j]); mb = fmaxf(mb, lb[j]); }
    ma = warp_max(ma); mb = warp_max(mb);
    if (lane == 0) { red[0][warp] = ma; red[1][warp] = mb; }
    __syncthreads();
    ma = red[0][0]; mb = red[1][0];
    for (int w = 1; w < CL_THREADS / 32; ++w) { ma = fmaxf(ma, red[0][w]); mb = fmaxf(mb, red[1][w]); }
    __syncthreads();
    float ea = 0.f, eb = 0.f;
    for (int j = threadIdx.x; j < 2 * B; j += blockDim.x) { ea += expf(la[j] - ma); eb += expf(lb[j] - mb); }
    ea = warp_sum(ea); eb = warp_sum(eb);
    if (lane == 0) { red[0][warp] = ea; red[1][warp] = eb; }
    __syncthreads();
    if (threadIdx.x == 0) {
        ea = eb = 0.f;
        for (int w = 0; w < CL_THREADS / 32; ++w) { ea += red[0][w]; eb += red[1][w]; }
        // cross entropy with target i: logsumexp - logit_i
        row_loss[i] = (ma + logf(ea) - la[i]) + (mb + logf(eb) - lb[i]);
    }
}

__global__ void mean_kernel(const float *__restrict__ v, int n, float *__restrict__ out) {
    __shared__ double sh[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += v[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = (float)(sh[0] / n);
}

}  // namespace

}  // namespace cutdet

using namespace cutdet;

extern "C" size_t cutdet_contrastive_loss_workspace_bytes(int pairs) { return (size_t)(pairs > 0 ? pairs : 0) * 3 * sizeof(float) + 256; }

extern "C" int cutdet_contrastive_loss(const float *x, int pairs, int dim, float temperature, int h_norm, float *loss,
                                       float *logits_ab, void *workspace, size_t workspace_bytes, cutdet_stream_t stream_) {
    CUTDET_REQUIRE(x && loss && workspace, "contrastive_loss: null pointer");
    CUTDET_REQUIRE(pairs > 0 && dim > 0 && temperature > 0.f, "contrastive_loss: bad shape or temperature");
    CUTDET_REQUIRE(pairs <= 6000, "contrastive_loss: %d pairs need more shared memory than a block has", pairs);
    if (workspace_bytes < cutdet_contrastive_loss_workspace_bytes(pairs))
        return fail(CUTDET_ECAPACITY, "contrastive_loss: workspace %zu < %zu bytes", workspace_bytes,
                    cutdet_contrastive_loss_workspace_bytes(pairs));
    cudaStream_t stream = as_stream(stream_);
    float *scale = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    float *row_loss = scale + 2 * pairs;
    const size_t smem = (size_t)4 * pairs * sizeof(float);
    if (smem > 48 * 1024)
        CUTDET_CUDA(cudaFuncSetAttribute(contrastive_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        KernelScope scope("contrastive_row_scale", stream);
        row_scale_kernel<<<(unsigned)ceil_div(2 * pairs, 128), 128, 0, stream>>>(x, 2 * pairs, dim, h_norm, scale);
    }
    CUTDET_LAUNCH_CHECK("row_scale_kernel");
    {
        KernelScope scope("contrastive_rows", stream);
        contrastive_rows_kernel<<<pairs, CL_THREADS, smem, stream>>>(x, scale, pairs, dim, 1.f / temperature, logits_ab, row_loss);
    }
    CUTDET_LAUNCH_CHECK("contrastive_rows_kernel");
    {
        KernelScope scope("contrastive_mean", stream);
        mean_kernel<<<1, 256, 0, stream>>>(row_loss, pairs, loss);
    }
    CUTDET_LAUNCH_CHECK("mean_kernel");
    return CUTDET_OK;
}
