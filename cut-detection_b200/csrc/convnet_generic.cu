// Generic CUDA-core kernels for the classifier: any FrameConvNet / FrameLinearNet the reference's
// constructors can build (frameID/net.py:71-189), float32 NCHW.  The tcgen05 kernels in conv_tc.cu take
// over for the architectures they are specialised for; these cover everything else and every odd input size.
//
// One CNNLayer (net.py:33-40) = conv3x3(pad 1) -> ReLU -> MaxPool(3, stride 3, floor) -> BatchNorm(eval)
// is ONE kernel: the 3x3 block of conv outputs under each pooled pixel stays in registers, the pool is a
// 9-way max of the raw accumulators (ReLU and +bias commute with max), then the BN affine -- applied after
// the pool, in the reference's order.  The full-resolution activation never exists in memory.
#include "common.cuh"
#include "net.cuh"

namespace cutdet {

namespace {

constexpr int PT_X = 32;               // pooled pixels per block along x (one warp: conflict-free stride-3 smem reads)
constexpr int PT_Y = 4;                // pooled rows per block
constexpr int CO_T = 8;                // output channels per thread
constexpr int CI_T = 8;                // input channels staged per iteration
constexpr int IN_W = 3 * PT_X + 2;     // 98
constexpr int IN_H = 3 * PT_Y + 2;     // 14

__global__ void __launch_bounds__(PT_X *PT_Y)
conv3x3_relu_pool3_bn_kernel(const float *__restrict__ in, float *__restrict__ out, const float *__restrict__ w_t,
                             const float *__restrict__ bias, const float *__restrict__ scale,
                             const float *__restrict__ shift, int cin, int cout, int cout_pad, int h, int w, int ph,
                             int pw, int co_groups) {
    __shared__ float s_in[CI_T][IN_H][IN_W];
    __shared__ __align__(16) float s_w[CI_T][9][CO_T];

    const int tx = threadIdx.x % PT_X, ty = threadIdx.x / PT_X;
    const int b = blockIdx.z / co_groups, cg = blockIdx.z % co_groups;
    const int px0 = blockIdx.x * PT_X, py0 = blockIdx.y * PT_Y;
    const int ix0 = 3 * px0 - 1, iy0 = 3 * py0 - 1;      // input coordinate of s_in[.][0][0]
    const float *in_b = in + (int64_t)b * cin * h * w;

    float acc[9][CO_T];
#pragma unroll
    for (int p = 0; p < 9; ++p)
#pragma unroll
        for (int c = 0; c < CO_T; ++c) acc[p][c] = 0.f;

    for (int ci0 = 0; ci0 < cin; ci0 += CI_T) {
        __syncthreads();
        for (int i = threadIdx.x; i < CI_T * IN_H * IN_W; i += PT_X * PT_Y) {
            const int ci = i / (IN_H * IN_W), r = (i / IN_W) % IN_H, c = i % IN_W;
            const int gy = iy0 + r, gx = ix0 + c;
            float v = 0.f;
            if (ci0 + ci < cin && gy >= 0 && gy < h && gx >= 0 && gx < w)
                v = in_b[((int64_t)(ci0 + ci) * h + gy) * w + gx];
            s_in[ci][r][c] = v;
        }
        for (int i = threadIdx.x; i < CI_T * 9 * CO_T; i += PT_X * PT_Y) {
            const int ci = i / (9 * CO_T), t = (i / CO_T) % 9, c = i % CO_T;
            float v = 0.f;
            if (ci0 + ci < cin) v = w_t[((int64_t)(ci0 + ci) * 9 + t) * cout_pad + cg * CO_T + c];
            s_w[ci][t][c] = v;
        }
        __syncthreads();
        const int nci = min(CI_T, cin - ci0);
        for (int ci = 0; ci < nci; ++ci) {
            float patch[5][5];
#pragma unroll
            for (int r = 0; r < 5; ++r)
#pragma unroll
                for (int c = 0; c < 5; ++c) patch[r][c] = s_in[ci][3 * ty + r][3 * tx + c];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float4 wa = *reinterpret_cast<const float4 *>(&s_w[ci][ky * 3 + kx][0]);
                    const float4 wb = *reinterpret_cast<const float4 *>(&s_w[ci][ky * 3 + kx][4]);
                    const float wv[CO_T] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            const float v = patch[dy + ky][dx + kx];
#pragma unroll
                            for (int c = 0; c < CO_T; ++c) acc[dy * 3 + dx][c] = fmaf(v, wv[c], acc[dy * 3 + dx][c]);
                        }
                }
        }
    }

    const int px = px0 + tx, py = py0 + ty;
    if (px >= pw || py >= ph) return;
#pragma unroll
    for (int c = 0; c < CO_T; ++c) {
        const int co = cg * CO_T + c;
        if (co >= cout) break;
        float m = acc[0][c];
#pragma unroll
        for (int p = 1; p < 9; ++p) m = fmaxf(m, acc[p][c]);
        m = fmaxf(m + bias[co], 0.f);
        out[(((int64_t)b * cout + co) * ph + py) * pw + px] = scale ? fmaf(m, scale[co], shift[co]) : m;
    }
}

// AdaptiveAvgPool2d(pool) + flatten in (c, i, j) order (net.py:130-131).
__global__ void avgpool_flatten_kernel(const float *__restrict__ in, float *__restrict__ out, int64_t total, int c, int h,
                                       int w, int pool) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int j = idx % pool, i = (idx / pool) % pool;
    const int64_t bc = idx / (pool * pool);
    const int r0 = (i * h) / pool, r1 = ((i + 1) * h + pool - 1) / pool;
    const int c0 = (j * w) / pool, c1 = ((j + 1) * w + pool - 1) / pool;
    const float *p = in + bc * h * w;
    float s = 0.f;
    for (int r = r0; r < r1; ++r)
        for (int cc = c0; cc < c1; ++cc) s += p[r * w + cc];
    out[idx] = s / (float)((r1 - r0) * (c1 - c0));
}

// FCLayer (net.py:62-68): Linear -> ReLU -> BatchNorm1d(eval); the last layer is Linear only.
// One warp per (frame, output feature).
__global__ void fc_kernel(const float *__restrict__ in, float *__restrict__ out, const float *__restrict__ w,
                          const float *__restrict__ bias, const float *__restrict__ scale,
                          const float *__restrict__ shift, int64_t batch, int n_in, int n_out, int relu) {
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= batch * n_out) return;
    const int64_t b = warp / n_out;
    const int o = warp % n_out;
    const float *x = in + b * n_in;
    const float *wr = w + (int64_t)o * n_in;
    float s = 0.f;
    for (int k = lane; k < n_in; k += 32) s = fmaf(x[k], wr[k], s);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) {
        s += bias[o];
        if (relu) s = fmaxf(s, 0.f);
        if (scale) s = fmaf(s, scale[o], shift[o]);
        out[b * n_out + o] = s;
    }
}

// BatchNorm in training mode (the reference's learn_contrasts.py never calls .eval(): nn.BatchNorm2d/1d normalise with the
// batch mean and the BIASED batch variance).  One block per channel: sums in double, then the affine in place.
__global__ void __launch_bounds__(256) bn_batchstats_kernel(float *__restrict__ data, int outer, int channels, int inner,
                                                            const float *__restrict__ gamma, const float *__restrict__ beta, float eps) {
    const int c = blockIdx.x;
    const int64_t n = (int64_t)outer * inner;
    double s = 0.0, ss = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = data[((i / inner) * channels + c) * inner + i % inner];
        s += v; ss += (double)v * v;
    }
    __shared__ double sh[2][256];
    sh[0][threadIdx.x] = s; sh[1][threadIdx.x] = ss;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { sh[0][threadIdx.x] += sh[0][threadIdx.x + o]; sh[1][threadIdx.x] += sh[1][threadIdx.x + o]; }
        __syncthreads();
    }
    const double mean = sh[0][0] / (double)n;
    const double var = fmax(sh[1][0] / (double)n - mean * mean, 0.0);
    const float scale = gamma[c] * (float)(1.0 / sqrt(var + (double)eps)), shift = beta[c] - (float)mean * scale;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        float *q = data + ((i / inner) * channels + c) * inner + i % inner;
        *q = fmaf(*q, scale, shift);
    }
}

}  // namespace

int launch_bn_batchstats(float *data, int outer, int channels, int inner, const float *gamma, const float *beta, float eps,
                         cudaStream_t stream) {
    if (outer <= 0 || channels <= 0 || inner <= 0) return CUTDET_OK;
    {
        KernelScope scope("bn_batchstats_kernel", stream);
        bn_batchstats_kernel<<<channels, 256, 0, stream>>>(data, outer, channels, inner, gamma, beta, eps);
    }
    CUTDET_LAUNCH_CHECK("bn_batchstats_kernel");
    return CUTDET_OK;
}

int launch_conv_block_generic(const float *in, float *out, const ConvLayer &L, int layer, int batch, int h, int w,
                              cudaStream_t stream, bool apply_bn) {
    static const char *const kNames[] = {"conv_block_generic_L0", "conv_block_generic_L1", "conv_block_generic_L2",
                                         "conv_block_generic_L3+"};
    const char *name = kNames[layer < 3 ? layer : 3];
    const int ph = h / 3, pw = w / 3;
    if (ph <= 0 || pw <= 0) return fail(CUTDET_EINVAL, "conv block: %dx%d input is smaller than the 3x3 pool", h, w);
    const int co_groups = (L.cout + CO_T - 1) / CO_T;
    const int cout_pad = co_groups * CO_T;
    const int max_b = 65535 / co_groups;
    for (int b0 = 0; b0 < batch; b0 += max_b) {
        const int nb = batch - b0 < max_b ? batch - b0 : max_b;
        dim3 grid((unsigned)ceil_div(pw, PT_X), (unsigned)ceil_div(ph, PT_Y), (unsigned)(nb * co_groups));
        {
            KernelScope scope(name, stream);
            conv3x3_relu_pool3_bn_kernel<<<grid, PT_X * PT_Y, 0, stream>>>(
            in + (int64_t)b0 * L.cin * h * w, out + (int64_t)b0 * L.cout * ph * pw, L.d_w_t, L.d_bias, apply_bn ? L.d_scale : nullptr,
            L.d_shift, L.cin, L.cout, cout_pad, h, w, ph, pw, co_groups);
        }
        CUTDET_LAUNCH_CHECK("conv3x3_relu_pool3_bn_kernel");
    }
    return CUTDET_OK;
}

int launch_avgpool_flatten(const float *in, float *out, int batch, int c, int h, int w, int pool,
                           cudaStream_t stream) {
    const int64_t total = (int64_t)batch * c * pool * pool;
    if (total == 0) return CUTDET_OK;
    {
        KernelScope scope("avgpool_flatten_kernel", stream);
        avgpool_flatten_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, stream>>>(in, out, total, c, h, w, pool);
    }
    CUTDET_LAUNCH_CHECK("avgpool_flatten_kernel");
    return CUTDET_OK;
}

int launch_fc(const float *in, float *out, const FcLayer &L, int batch, bool relu, cudaStream_t stream, bool apply_bn) {
    const int64_t warps = (int64_t)batch * L.out;
    if (warps == 0) return CUTDET_OK;
    {
        KernelScope scope("fc_kernel", stream);
        fc_kernel<<<(unsigned)ceil_div(warps * 32, 256), 256, 0, stream>>>(in, out, L.d_w, L.d_bias,
                                                                      L.has_bn && apply_bn ? L.d_scale : nullptr,
                                                                      L.has_bn && apply_bn ? L.d_shift : nullptr, batch, L.in, L.out,
                                                                      relu ? 1 : 0);
    }
    CUTDET_LAUNCH_CHECK("fc_kernel");
    return CUTDET_OK;
}

}  // namespace cutdet
