// cutdet_net: parameter packing and forward orchestration (C ABI).
//
// Stands in for load_and_glue_nets (reference frameID/net.py:193-217: two state_dicts -> one callable) and for
// nn.Sequential(FrameConvNet, FrameLinearNet).forward (net.py:122-133, 180-186).
#include <math.h>

#include "common.cuh"
#include "net.cuh"
#include "preprocess.cuh"
#include "conv_tc.cuh"

using namespace cutdet;

namespace {

int upload(cutdet_net *net, const std::vector<float> &h, float **d) {
    void *p = nullptr;
    CUTDET_CUDA(cudaMalloc(&p, h.size() * sizeof(float)));
    net->dev_allocs.push_back(p);
    CUTDET_CUDA(cudaMemcpy(p, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
    CUTDET_CUDA(cudaDeviceSynchronize());        // see upload_bytes (conv_tc.cu): the copy's DMA runs on the legacy stream
    *d = reinterpret_cast<float *>(p);
    return CUTDET_OK;
}

void fold_bn(int n, const float *g, const float *b, const float *m, const float *v, float eps, std::vector<float> &scale,
             std::vector<float> &shift) {
    scale.resize(n);
    shift.resize(n);
    for (int i = 0; i < n; ++i) {
        const double s = (double)g[i] / sqrt((double)v[i] + (double)eps);
        scale[i] = (float)s;
        shift[i] = (float)((double)b[i] - (double)m[i] * s);
    }
}

// Shapes of every conv layer for a given input size.
std::vector<LayerGeom> layer_geometry(const cutdet_net *net, int h, int w) {
    std::vector<LayerGeom> g;
    int cin = net->cfg.input_channels;
    for (int i = 0; i < net->cfg.n_conv_layers; ++i) {
        LayerGeom L{cin, net->cfg.hidden_channels, h, w, h / 3, w / 3};
        g.push_back(L);
        cin = L.cout; h = L.ph; w = L.pw;
    }
    return g;
}

struct GenericWorkspace {
    std::vector<size_t> conv_out;   // byte offset of each conv layer's float32 NCHW output
    size_t pooled = 0;              // [B, fc_input]
    std::vector<size_t> fc_out;     // hidden FC outputs (the last FC writes the caller's logits)
    size_t input_f32 = 0;           // forward_frames: preprocessed float32 input
    size_t total = 0;
};

size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

GenericWorkspace generic_workspace(const cutdet_net *net, int batch, int h, int w, bool with_input) {
    GenericWorkspace ws;
    size_t off = 0;
    if (with_input) {
        ws.input_f32 = off;
        off = align_up(off + (size_t)batch * net->cfg.input_channels * h * w * sizeof(float));
    }
    for (const LayerGeom &L : layer_geometry(net, h, w)) {
        ws.conv_out.push_back(off);
        off = align_up(off + (size_t)batch * L.cout * L.ph * L.pw * sizeof(float));
    }
    ws.pooled = off;
    off = align_up(off + (size_t)batch * net->cfg.fc_input_size * sizeof(float));
    for (int j = 0; j + 1 < net->cfg.n_fc_layers; ++j) {
        ws.fc_out.push_back(off);
        off = align_up(off + (size_t)batch * net->fc[j].out * sizeof(float));
    }
    ws.total = off;
    return ws;
}

int check_input_size(const cutdet_net *net, int h, int w) {
    if (net->cfg.n_conv_layers == 0 && (h != 1 || w != 1))
        return fail(CUTDET_EINVAL, "an FC-only net takes [B, %d, 1, 1] inputs", net->cfg.fc_input_size);
    for (const LayerGeom &L : layer_geometry(net, h, w))
        if (L.ph < 1 || L.pw < 1)
            return fail(CUTDET_EINVAL, "input %dx%d is too small for %d conv+pool layers", h, w, net->cfg.n_conv_layers);
    return CUTDET_OK;
}

// batch_stats: every BatchNorm uses the statistics of this batch (training-mode forward) instead of its running statistics
int forward_generic(cutdet_net *net, const float *x, int batch, int h, int w, float *logits, char *ws_base,
                    const GenericWorkspace &ws, cudaStream_t stream, bool batch_stats = false) {
    const float *cur = x;
    std::vector<LayerGeom> geom = layer_geometry(net, h, w);
    for (size_t i = 0; i < geom.size(); ++i) {
        float *out = reinterpret_cast<float *>(ws_base + ws.conv_out[i]);
        const ConvLayer &L = net->conv[i];
        if (int rc = launch_conv_block_generic(cur, out, L, (int)i, batch, geom[i].h, geom[i].w, stream, !batch_stats)) return rc;
        if (batch_stats)
            if (int rc = launch_bn_batchstats(out, batch, L.cout, geom[i].ph * geom[i].pw, L.d_gamma, L.d_beta, L.eps, stream)) return rc;
        cur = out;
    }
    if (!geom.empty()) {
        const LayerGeom &last = geom.back();
        float *pooled = net->cfg.n_fc_layers == 0 ? logits : reinterpret_cast<float *>(ws_base + ws.pooled);
        if (int rc = launch_avgpool_flatten(cur, pooled, batch, last.cout, last.ph, last.pw, net->cfg.avg_pool_size, stream))
            return rc;
        cur = pooled;
    }
    for (int j = 0; j < net->cfg.n_fc_layers; ++j) {
        const bool is_last = j + 1 == net->cfg.n_fc_layers;
        float *out = is_last ? logits : reinterpret_cast<float *>(ws_base + ws.fc_out[j]);
        if (int rc = launch_fc(cur, out, net->fc[j], batch, !is_last, stream, !batch_stats)) return rc;
        if (batch_stats && net->fc[j].has_bn)
            if (int rc = launch_bn_batchstats(out, batch, net->fc[j].out, 1, net->fc[j].d_gamma, net->fc[j].d_beta, net->fc[j].eps, stream))
                return rc;
        cur = out;
    }
    return CUTDET_OK;
}

}  // namespace

extern "C" int cutdet_net_create(const cutdet_net_config *cfg, cutdet_net **out) {
    CUTDET_REQUIRE(cfg && out, "net_create: null argument");
    CUTDET_REQUIRE(cfg->input_channels > 0 && cfg->hidden_channels > 0 && cfg->n_conv_layers >= 0 && cfg->avg_pool_size > 0 &&
                       cfg->n_fc_layers >= 0 && cfg->fc_hidden_size > 0 && cfg->fc_output_size > 0 &&
                       cfg->n_conv_layers + cfg->n_fc_layers > 0,
                   "net_create: bad size in config");
    // A conv-only net (n_fc_layers == 0, a bare FrameConvNet) returns the flattened pooled features; an FC-only net
    // (n_conv_layers == 0, a bare FrameLinearNet) takes [B, fc_input_size] features as a [B, fc_input_size, 1, 1] input.
    if (cfg->n_conv_layers > 0)
        CUTDET_REQUIRE(cfg->fc_input_size == cfg->hidden_channels * cfg->avg_pool_size * cfg->avg_pool_size,
                       "net_create: fc_input_size %d != hidden_channels * avg_pool_size^2 = %d", cfg->fc_input_size,
                       cfg->hidden_channels * cfg->avg_pool_size * cfg->avg_pool_size);
    else
        CUTDET_REQUIRE(cfg->input_channels == cfg->fc_input_size && cfg->avg_pool_size == 1,
                       "net_create: an FC-only net needs input_channels == fc_input_size and avg_pool_size == 1");
    cutdet_net *net = new cutdet_net();
    net->cfg = *cfg;
    net->conv.resize(cfg->n_conv_layers);
    net->fc.resize(cfg->n_fc_layers);
    int cin = cfg->input_channels;
    for (auto &L : net->conv) { L.cin = cin; L.cout = cfg->hidden_channels; cin = L.cout; }
    for (int j = 0; j < cfg->n_fc_layers; ++j) {
        net->fc[j].in = j == 0 ? cfg->fc_input_size : cfg->fc_hidden_size;
        net->fc[j].out = j + 1 == cfg->n_fc_layers ? cfg->fc_output_size : cfg->fc_hidden_size;
        net->fc[j].has_bn = j + 1 < cfg->n_fc_layers;
    }
    *out = net;
    return CUTDET_OK;
}

extern "C" void cutdet_net_destroy(cutdet_net *net) {
    if (!net) return;
    tc_destroy(net);
    for (void *p : net->dev_allocs) cudaFree(p);
    delete net;
}

extern "C" int cutdet_net_set_conv_layer(cutdet_net *net, int layer, const float *w, const float *b, const float *g,
                                         const float *beta, const float *mean, const float *var, float eps) {
    CUTDET_REQUIRE(net && !net->finalized, "set_conv_layer: null or finalized net");
    CUTDET_REQUIRE(layer >= 0 && layer < (int)net->conv.size(), "set_conv_layer: layer %d out of range", layer);
    CUTDET_REQUIRE(w && b && g && beta && mean && var, "set_conv_layer: null parameter array");
    ConvLayer &L = net->conv[layer];
    L.w.assign(w, w + (size_t)L.cout * L.cin * 9);
    L.bias.assign(b, b + L.cout);
    fold_bn(L.cout, g, beta, mean, var, eps, L.scale, L.shift);
    L.gamma.assign(g, g + L.cout);
    L.beta.assign(beta, beta + L.cout);
    L.eps = eps;
    L.set = true;
    return CUTDET_OK;
}

extern "C" int cutdet_net_set_fc_layer(cutdet_net *net, int layer, const float *w, const float *b, const float *g,
                                       const float *beta, const float *mean, const float *var, float eps) {
    CUTDET_REQUIRE(net && !net->finalized, "set_fc_layer: null or finalized net");
    CUTDET_REQUIRE(layer >= 0 && layer < (int)net->fc.size(), "set_fc_layer: layer %d out of range", layer);
    CUTDET_REQUIRE(w && b, "set_fc_layer: null weight/bias");
    FcLayer &L = net->fc[layer];
    if (L.has_bn) {
        CUTDET_REQUIRE(g && beta && mean && var, "set_fc_layer: layer %d needs BatchNorm parameters", layer);
        fold_bn(L.out, g, beta, mean, var, eps, L.scale, L.shift);
        L.gamma.assign(g, g + L.out);
        L.beta.assign(beta, beta + L.out);
        L.eps = eps;
    } else if (g || beta || mean || var) {
        // FrameLinearNet never puts a BatchNorm on its last layer (net.py:164-167), a lone FCLayer may have one (net.py:43-68):
        // accepted on a one-layer FC-only net, which is what FCLayer.forward builds
        CUTDET_REQUIRE(net->cfg.n_conv_layers == 0 && net->cfg.n_fc_layers == 1 && g && beta && mean && var,
                       "set_fc_layer: the last layer has no BatchNorm (net.py:164-167)");
        L.has_bn = true;
        fold_bn(L.out, g, beta, mean, var, eps, L.scale, L.shift);
        L.gamma.assign(g, g + L.out);
        L.beta.assign(beta, beta + L.out);
        L.eps = eps;
    }
    L.w.assign(w, w + (size_t)L.out * L.in);
    L.bias.assign(b, b + L.out);
    L.set = true;
    return CUTDET_OK;
}

extern "C" int cutdet_net_finalize(cutdet_net *net) {
    CUTDET_REQUIRE(net && !net->finalized, "net_finalize: null or already finalized");
    for (size_t i = 0; i < net->conv.size(); ++i) CUTDET_REQUIRE(net->conv[i].set, "net_finalize: conv layer %zu not set", i);
    for (size_t j = 0; j < net->fc.size(); ++j) CUTDET_REQUIRE(net->fc[j].set, "net_finalize: fc layer %zu not set", j);
    if (int rc = cutdet_device_check(nullptr, nullptr, nullptr)) return rc;
    for (ConvLayer &L : net->conv) {
        const int cout_pad = (L.cout + 7) / 8 * 8;
        std::vector<float> wt((size_t)L.cin * 9 * cout_pad, 0.f);
        for (int co = 0; co < L.cout; ++co)
            for (int ci = 0; ci < L.cin; ++ci)
                for (int t = 0; t < 9; ++t) wt[((size_t)ci * 9 + t) * cout_pad + co] = L.w[((size_t)co * L.cin + ci) * 9 + t];
        if (int rc = upload(net, wt, &L.d_w_t)) return rc;
        if (int rc = upload(net, L.bias, &L.d_bias)) return rc;
        if (int rc = upload(net, L.scale, &L.d_scale)) return rc;
        if (int rc = upload(net, L.shift, &L.d_shift)) return rc;
        if (int rc = upload(net, L.gamma, &L.d_gamma)) return rc;
        if (int rc = upload(net, L.beta, &L.d_beta)) return rc;
    }
    for (FcLayer &L : net->fc) {
        if (int rc = upload(net, L.w, &L.d_w)) return rc;
        if (int rc = upload(net, L.bias, &L.d_bias)) return rc;
        if (L.has_bn) {
            if (int rc = upload(net, L.scale, &L.d_scale)) return rc;
            if (int rc = upload(net, L.shift, &L.d_shift)) return rc;
            if (int rc = upload(net, L.gamma, &L.d_gamma)) return rc;
            if (int rc = upload(net, L.beta, &L.d_beta)) return rc;
        }
    }
    if (int rc = tc_prepare(net)) return rc;
    net->finalized = true;
    return CUTDET_OK;
}

extern "C" int cutdet_net_set_option(cutdet_net *net, int option, int value) {
    CUTDET_REQUIRE(net, "net_set_option: null net");
    CUTDET_REQUIRE(value >= 0, "net_set_option: negative value");
    switch (option) {
        case CUTDET_OPT_CONV1_ACC32: net->opt.conv1_acc32 = value != 0; break;
        case CUTDET_OPT_SUB_BATCH: net->opt.sub_batch = value; break;
        case CUTDET_OPT_GROUP_FRAMES: net->opt.group_frames = value; break;
        case CUTDET_OPT_NO_PDL: net->opt.no_pdl = value != 0; break;
        case CUTDET_OPT_CONV1_GRID: net->opt.conv1_grid = value; break;
        case CUTDET_OPT_CONV1_VARIANT: net->opt.conv1_variant = value; break;
        case CUTDET_OPT_L2_PERSIST: net->opt.l2_persist = value; break;
        case CUTDET_OPT_RING_CAP: net->opt.ring_cap = value; break;
        case CUTDET_OPT_SRC_PREFETCH: net->opt.src_prefetch = value; break;
        default: return fail(CUTDET_EINVAL, "net_set_option: unknown option %d", option);
    }
    return CUTDET_OK;
}

extern "C" int cutdet_net_get_option(const cutdet_net *net, int option, int *value) {
    CUTDET_REQUIRE(net && value, "net_get_option: null argument");
    switch (option) {
        case CUTDET_OPT_CONV1_ACC32: *value = net->opt.conv1_acc32; break;
        case CUTDET_OPT_SUB_BATCH: *value = net->opt.sub_batch; break;
        case CUTDET_OPT_GROUP_FRAMES: *value = net->opt.group_frames; break;
        case CUTDET_OPT_NO_PDL: *value = net->opt.no_pdl; break;
        case CUTDET_OPT_CONV1_GRID: *value = net->opt.conv1_grid; break;
        case CUTDET_OPT_CONV1_VARIANT: *value = net->opt.conv1_variant; break;
        case CUTDET_OPT_L2_PERSIST: *value = net->opt.l2_persist; break;
        case CUTDET_OPT_RING_CAP: *value = net->opt.ring_cap; break;
        case CUTDET_OPT_SRC_PREFETCH: *value = net->opt.src_prefetch; break;
        default: return fail(CUTDET_EINVAL, "net_get_option: unknown option %d", option);
    }
    return CUTDET_OK;
}

extern "C" int cutdet_net_debug_timeline(cutdet_net *net, int kernel, long long *stamps_dev, size_t n_entries) {
    CUTDET_REQUIRE(net, "net_debug_timeline: null net");
    if (kernel == 0 || !stamps_dev) { net->opt.timeline_kernel = 0; net->opt.timeline_dev = nullptr; return CUTDET_OK; }
    CUTDET_REQUIRE(kernel >= 1 && kernel <= 3 && n_entries >= 4096, "net_debug_timeline: kernel 1, 2 or 3 and >= 4096 entries");
    net->opt.timeline_kernel = kernel;
    net->opt.timeline_dev = stamps_dev;
    return CUTDET_OK;
}

extern "C" int cutdet_net_forward_conv_layer(cutdet_net *net, int layer, const float *x, int batch, int height, int width,
                                             float *out, int bn_mode, cutdet_stream_t stream) {
    CUTDET_REQUIRE(net && net->finalized, "net_forward_conv_layer: net not finalized");
    CUTDET_REQUIRE(layer >= 0 && layer < net->cfg.n_conv_layers, "net_forward_conv_layer: layer %d out of range", layer);
    CUTDET_REQUIRE(batch >= 0 && height > 0 && width > 0 && bn_mode >= 0 && bn_mode <= 2, "net_forward_conv_layer: bad argument");
    if (batch == 0) return CUTDET_OK;
    CUTDET_REQUIRE(x && out, "net_forward_conv_layer: null pointer");
    CUTDET_REQUIRE(bn_mode != 2 || batch * (height / 3) * (width / 3) >= 2,
                   "net_forward_conv_layer: batch statistics need more than one value per channel");
    const ConvLayer &L = net->conv[layer];
    if (int rc = launch_conv_block_generic(x, out, L, layer, batch, height, width, as_stream(stream), bn_mode == 1)) return rc;
    if (bn_mode == 2)
        return launch_bn_batchstats(out, batch, L.cout, (height / 3) * (width / 3), L.d_gamma, L.d_beta, L.eps, as_stream(stream));
    return CUTDET_OK;
}

extern "C" int cutdet_net_forward_fc_layer(cutdet_net *net, int layer, const float *x, int batch, float *out, int relu, int bn_mode,
                                           cutdet_stream_t stream) {
    CUTDET_REQUIRE(net && net->finalized, "net_forward_fc_layer: net not finalized");
    CUTDET_REQUIRE(layer >= 0 && layer < net->cfg.n_fc_layers, "net_forward_fc_layer: layer %d out of range", layer);
    CUTDET_REQUIRE(batch >= 0 && bn_mode >= 0 && bn_mode <= 2, "net_forward_fc_layer: bad argument");
    if (batch == 0) return CUTDET_OK;
    CUTDET_REQUIRE(x && out, "net_forward_fc_layer: null pointer");
    const FcLayer &L = net->fc[layer];
    CUTDET_REQUIRE(bn_mode == 0 || L.has_bn, "net_forward_fc_layer: layer %d has no BatchNorm", layer);
    CUTDET_REQUIRE(bn_mode != 2 || batch >= 2, "net_forward_fc_layer: batch statistics need more than one value per channel");
    if (int rc = launch_fc(x, out, L, batch, relu != 0, as_stream(stream), bn_mode == 1)) return rc;
    if (bn_mode == 2) return launch_bn_batchstats(out, batch, L.out, 1, L.d_gamma, L.d_beta, L.eps, as_stream(stream));
    return CUTDET_OK;
}

extern "C" int cutdet_net_uses_tensor_cores(const cutdet_net *net, int height, int width) {
    return net && net->finalized && tc_supported(net, height, width) ? 1 : 0;
}

extern "C" int cutdet_net_workspace_bytes(const cutdet_net *net, int batch, int height, int width, size_t *bytes) {
    CUTDET_REQUIRE(net && bytes && batch >= 0 && height > 0 && width > 0, "net_workspace_bytes: bad argument");
    if (int rc = check_input_size(net, height, width)) return rc;
    size_t need = generic_workspace(net, batch, height, width, true).total;
    if (net->finalized && tc_supported(net, height, width)) {
        const size_t t = tc_workspace_bytes(net, batch, height, width);
        if (t > need) need = t;
    }
    *bytes = need + 256;
    return CUTDET_OK;
}

extern "C" int cutdet_net_forward_f32(cutdet_net *net, const float *x, int batch, int height, int width, float *logits,
                                      void *workspace, size_t workspace_bytes, cutdet_stream_t stream) {
    CUTDET_REQUIRE(net && net->finalized, "net_forward_f32: net not finalized");
    CUTDET_REQUIRE(batch >= 0 && height > 0 && width > 0, "net_forward_f32: bad shape");
    if (batch == 0) return CUTDET_OK;
    CUTDET_REQUIRE(x && logits && workspace, "net_forward_f32: null pointer");
    if (int rc = check_input_size(net, height, width)) return rc;
    size_t need = 0;
    if (int rc = cutdet_net_workspace_bytes(net, batch, height, width, &need)) return rc;
    if (workspace_bytes < need)
        return fail(CUTDET_ECAPACITY, "net_forward_f32: workspace %zu < %zu bytes", workspace_bytes, need);
    char *base = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    if (tc_supported(net, height, width))
        return tc_forward_f32(net, x, batch, height, width, logits, base, as_stream(stream));
    GenericWorkspace ws = generic_workspace(net, batch, height, width, true);
    return forward_generic(net, x, batch, height, width, logits, base, ws, as_stream(stream));
}

extern "C" int cutdet_net_forward_f32_batchstats(cutdet_net *net, const float *x, int batch, int height, int width, float *out,
                                                 void *workspace, size_t workspace_bytes, int use_tensor_cores, cutdet_stream_t stream) {
    CUTDET_REQUIRE(net && net->finalized, "net_forward_f32_batchstats: net not finalized");
    CUTDET_REQUIRE(batch >= 0 && height > 0 && width > 0, "net_forward_f32_batchstats: bad shape");
    if (batch == 0) return CUTDET_OK;
    CUTDET_REQUIRE(batch >= 2, "net_forward_f32_batchstats: batch statistics need more than one value per channel");
    CUTDET_REQUIRE(x && out && workspace, "net_forward_f32_batchstats: null pointer");
    if (int rc = check_input_size(net, height, width)) return rc;
    size_t need = 0;
    if (int rc = cutdet_net_workspace_bytes(net, batch, height, width, &need)) return rc;
    if (workspace_bytes < need)
        return fail(CUTDET_ECAPACITY, "net_forward_f32_batchstats: workspace %zu < %zu bytes", workspace_bytes, need);
    char *base = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    if (use_tensor_cores && tc_batchstats_supported(net, batch, height, width))
        return tc_forward_f32_batchstats(net, x, batch, height, width, out, base, as_stream(stream));
    GenericWorkspace ws = generic_workspace(net, batch, height, width, true);
    return forward_generic(net, x, batch, height, width, out, base, ws, as_stream(stream), true);
}

extern "C" int cutdet_net_forward_frames(cutdet_net *net, const cutdet_resize_plan *plan, const cutdet_frames *src,
                                         float *logits, void *workspace, size_t workspace_bytes, cutdet_stream_t stream) {
    CUTDET_REQUIRE(net && net->finalized, "net_forward_frames: net not finalized");
    if (int rc = check_frames(plan, src)) return rc;
    CUTDET_REQUIRE(net->cfg.input_channels == 3, "net_forward_frames: frames have 3 channels, the net expects %d",
                   net->cfg.input_channels);
    const int batch = src->batch, h = plan->host.dst_h, w = plan->host.dst_w;
    if (batch == 0) return CUTDET_OK;
    CUTDET_REQUIRE(logits && workspace, "net_forward_frames: null pointer");
    if (int rc = check_input_size(net, h, w)) return rc;
    size_t need = 0;
    if (int rc = cutdet_net_workspace_bytes(net, batch, h, w, &need)) return rc;
    if (workspace_bytes < need)
        return fail(CUTDET_ECAPACITY, "net_forward_frames: workspace %zu < %zu bytes", workspace_bytes, need);
    char *base = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    if (tc_supported(net, h, w)) return tc_forward_frames(net, plan, src, logits, base, as_stream(stream));
    GenericWorkspace ws = generic_workspace(net, batch, h, w, true);
    float *x = reinterpret_cast<float *>(base + ws.input_f32);
    if (int rc = cutdet_preprocess_f32(plan, src, x, stream)) return rc;
    return forward_generic(net, x, batch, h, w, logits, base, ws, as_stream(stream));
}

extern "C" int cutdet_net_debug_conv_output(cutdet_net *net, int layer, int batch, int height, int width,
                                            const void *workspace, float *out, cutdet_stream_t stream) {
    CUTDET_REQUIRE(net && net->finalized && workspace && out, "net_debug_conv_output: bad argument");
    CUTDET_REQUIRE(layer >= 0 && layer < net->cfg.n_conv_layers, "net_debug_conv_output: layer out of range");
    const char *base = reinterpret_cast<const char *>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    if (tc_supported(net, height, width))
        return tc_debug_conv_output(net, layer, batch, height, width, base, out, as_stream(stream));
    GenericWorkspace ws = generic_workspace(net, batch, height, width, true);
    std::vector<LayerGeom> geom = layer_geometry(net, height, width);
    const size_t bytes = (size_t)batch * geom[layer].cout * geom[layer].ph * geom[layer].pw * sizeof(float);
    CUTDET_CUDA(cudaMemcpyAsync(out, base + ws.conv_out[layer], bytes, cudaMemcpyDeviceToDevice, as_stream(stream)));
    return CUTDET_OK;
}
