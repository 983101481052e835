// Internal layout of a cutdet_net and the launchers of its kernels.
#pragma once

#include <vector>

#include "common.cuh"

namespace cutdet {

struct TcState;   // tensor-core path (conv_tc.cu)

struct ConvLayer {
    int cin = 0, cout = 0;
    bool set = false;
    std::vector<float> w;       // [cout][cin][3][3] as given
    std::vector<float> bias;    // [cout]
    std::vector<float> scale;   // gamma / sqrt(var + eps)
    std::vector<float> shift;   // beta - mean * scale
    std::vector<float> gamma, beta;   // BatchNorm weight/bias as given (batch-statistics forward)
    float eps = 1e-5f;
    // generic path (device, float32)
    float *d_w_t = nullptr;     // [cin][9][cout_padded8]
    float *d_bias = nullptr, *d_scale = nullptr, *d_shift = nullptr;
    float *d_gamma = nullptr, *d_beta = nullptr;
};

struct FcLayer {
    int in = 0, out = 0;
    bool set = false, has_bn = false;
    std::vector<float> w, bias, scale, shift, gamma, beta;
    float eps = 1e-5f;
    float *d_w = nullptr, *d_bias = nullptr, *d_scale = nullptr, *d_shift = nullptr, *d_gamma = nullptr, *d_beta = nullptr;
};

struct LayerGeom {
    int cin, cout, h, w, ph, pw;   // conv input h x w, pooled output ph x pw
};

// generic CUDA-core kernels (any architecture the reference's constructors can build)
int launch_conv_block_generic(const float *in, float *out, const ConvLayer &L, int layer, int batch, int h, int w,
                              cudaStream_t stream, bool apply_bn = true);
int launch_avgpool_flatten(const float *in, float *out, int batch, int c, int h, int w, int pool,
                           cudaStream_t stream);
int launch_fc(const float *in, float *out, const FcLayer &L, int batch, bool relu, cudaStream_t stream, bool apply_bn = true);
// BatchNorm with the statistics of THIS batch (training-mode forward, biased variance), in place over [outer][channels][inner]
int launch_bn_batchstats(float *data, int outer, int channels, int inner, const float *gamma, const float *beta, float eps,
                         cudaStream_t stream);

}  // namespace cutdet

// cutdet_net_set_option: experiment and test switches, per net (nothing is read from the environment)
struct cutdet_net_options {
    int conv1_acc32 = 0;      // layer 1 of the fused frames kernel accumulates in fp32 instead of fp16
    int sub_batch = 0;        // frames per conv1/conv2 pass (0 = default, one frame per SM)
    int group_frames = 0;     // frames gathered per conv3 launch (0 = default)
    int no_pdl = 0;           // ordinary launches instead of programmatic dependent launch
    int conv1_grid = 0;       // cap on the fused conv1 grid (test hook: several frames per CTA)
    int ring_cap = 0;         // conv12_frames with two source rows per output row: 1 = keep the full operand ring (see CUTDET_OPT_RING_CAP)
    int src_prefetch = 0;     // experiment: conv12_frames' loaders prefetch source rows into the L2 (see CUTDET_OPT_SRC_PREFETCH)
    int l2_persist = 0;       // experiment: conv12_frames' layer-1 slots as a persisting L2 window
    int conv1_variant = 0;    // 1 = experiment: the fused conv1 kernel with two epilogue sets and an MMA issuer per block row (A/B runs)
    int timeline_kernel = 0;  // cutdet_net_debug_timeline: 1 = conv1_fused_tc, 2 = conv2_tc
    long long *timeline_dev = nullptr;
};

struct cutdet_net {
    cutdet_net_config cfg;
    cutdet_net_options opt;
    std::vector<cutdet::ConvLayer> conv;
    std::vector<cutdet::FcLayer> fc;
    bool finalized = false;
    std::vector<void *> dev_allocs;
    cutdet::TcState *tc = nullptr;  // tensor-core path (conv_tc.cu), null when not applicable
};
