// Forward pieces of the reference's supervised training / validation loop (training_scripts/supervised_training.py):
//   criterion = torch.nn.CrossEntropyLoss(reduction="sum")       :132   loss = criterion(pred, labels)      :148, 186
//   pc = torch.max(pred, dim=1)[1];  correct[c] += sum(pc[labels == c] == c);  total[c] += sum(labels == c)   :188-193
// One pass over the [N, C] logits: per frame log-sum-exp minus the label's logit (float32, as torch computes it on a float32
// tensor), the first-index argmax, and the per-class counters.  Sums are accumulated in float64 in the caller's workspace and
// rounded once.  Forward only: no gradient is produced (the optimisation step is out of scope).
#include <algorithm>

#include "common.cuh"

namespace cutdet {

namespace {

constexpr int CE_THREADS = 256;
constexpr int CE_MAX_CLASSES = 64;      // counters live in shared memory

__global__ void __launch_bounds__(CE_THREADS) cross_entropy_sum_kernel(const float *__restrict__ logits, const int64_t *__restrict__ labels,
                                                                       int64_t n, int c, double *__restrict__ loss_acc,
                                                                       unsigned long long *__restrict__ counts /* [2][c] or null */,
                                                                       int *__restrict__ bad_label) {
    __shared__ double s_loss[CE_THREADS / 32];
    __shared__ unsigned int s_cnt[2][CE_MAX_CLASSES];
    for (int i = threadIdx.x; i < 2 * CE_MAX_CLASSES; i += blockDim.x) (&s_cnt[0][0])[i] = 0u;
    __syncthreads();
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float *row = logits + i * c;
        const int64_t y = labels[i];
        if (y < 0 || y >= c) { *bad_label = 1; continue; }      // torch raises "Target out of bounds": reported to the host
        float m = row[0];
        int arg = 0;
        for (int k = 1; k < c; ++k) {
            const float v = row[k];
            if (v > m) { m = v; arg = k; }                       // first maximum wins, as torch.max
        }
        float se = 0.f;
        for (int k = 0; k < c; ++k) se += expf(row[k] - m);
        acc += (double)(logf(se) + m - row[y]);                  // -log_softmax(row)[y]
        if (counts) {
            atomicAdd(&s_cnt[1][y], 1u);
            if (arg == (int)y) atomicAdd(&s_cnt[0][y], 1u);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_loss[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < CE_THREADS / 32; ++w) t += s_loss[w];
        atomicAdd(loss_acc, t);
    }
    if (counts)
        for (int k = threadIdx.x; k < c; k += blockDim.x) {
            if (s_cnt[0][k]) atomicAdd(&counts[k], (unsigned long long)s_cnt[0][k]);
            if (s_cnt[1][k]) atomicAdd(&counts[c + k], (unsigned long long)s_cnt[1][k]);
        }
}

__global__ void cross_entropy_finish_kernel(const double *__restrict__ loss_acc, const unsigned long long *__restrict__ counts, int c,
                                            float *__restrict__ loss, int64_t *__restrict__ correct, int64_t *__restrict__ total) {
    if (threadIdx.x == 0 && loss) *loss = (float)*loss_acc;
    for (int k = threadIdx.x; k < c; k += blockDim.x) {
        if (correct) correct[k] = (int64_t)counts[k];
        if (total) total[k] = (int64_t)counts[c + k];
    }
}

}  // namespace
}  // namespace cutdet

using namespace cutdet;

extern "C" size_t cutdet_cross_entropy_workspace_bytes(int n_classes) {
    return n_classes > 0 ? 16 + 16 * (size_t)n_classes : 0;      // loss accumulator, bad-label flag, 2 x n_classes counters
}

extern "C" int cutdet_cross_entropy_sum(const float *logits, const int64_t *labels, int64_t n, int n_classes, float *loss,
                                        int64_t *correct, int64_t *total, void *workspace, size_t workspace_bytes, int *bad_label_host,
                                        cutdet_stream_t stream) {
    CUTDET_REQUIRE(n >= 0 && n_classes >= 1 && n_classes <= CE_MAX_CLASSES, "cross_entropy_sum: bad shape [%lld, %d]", (long long)n, n_classes);
    CUTDET_REQUIRE(loss && workspace && workspace_bytes >= cutdet_cross_entropy_workspace_bytes(n_classes) &&
                       reinterpret_cast<uintptr_t>(workspace) % 8 == 0,
                   "cross_entropy_sum: null output or workspace too small / misaligned");
    CUTDET_REQUIRE(n == 0 || (logits && labels), "cross_entropy_sum: null input");
    char *ws = reinterpret_cast<char *>(workspace);
    double *loss_acc = reinterpret_cast<double *>(ws);
    int *bad = reinterpret_cast<int *>(ws + 8);
    unsigned long long *counts = reinterpret_cast<unsigned long long *>(ws + 16);
    const bool want_counts = correct || total;
    CUTDET_CUDA(cudaMemsetAsync(workspace, 0, cutdet_cross_entropy_workspace_bytes(n_classes), as_stream(stream)));
    if (n > 0) {
        const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(n, CE_THREADS), 4 * (int64_t)sm_count());
        KernelScope scope("cross_entropy_sum_kernel", as_stream(stream));
        cross_entropy_sum_kernel<<<grid, CE_THREADS, 0, as_stream(stream)>>>(logits, labels, n, n_classes, loss_acc, want_counts ? counts : nullptr, bad);
    }
    CUTDET_LAUNCH_CHECK("cross_entropy_sum_kernel");
    {
        KernelScope scope("cross_entropy_finish_kernel", as_stream(stream));
        cross_entropy_finish_kernel<<<1, 64, 0, as_stream(stream)>>>(loss_acc, counts, n_classes, loss, correct, total);
    }
    CUTDET_LAUNCH_CHECK("cross_entropy_finish_kernel");
    if (bad_label_host) {       // optional synchronous check, as torch's "Target out of bounds" assertion
        CUTDET_CUDA(cudaMemcpyAsync(bad_label_host, bad, sizeof(int), cudaMemcpyDeviceToHost, as_stream(stream)));
        CUTDET_CUDA(cudaStreamSynchronize(as_stream(stream)));
        if (*bad_label_host) return fail(CUTDET_EINVAL, "cross_entropy_sum: a label is outside [0, %d)", n_classes);
    }
    return CUTDET_OK;
}
