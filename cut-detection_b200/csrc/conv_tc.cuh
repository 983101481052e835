// Tensor-core (tcgen05 / TMEM / TMA) path of the classifier: interface used by net.cu.
#pragma once

#include "common.cuh"
#include "net.cuh"
#include "preprocess.cuh"

namespace cutdet {

// Packs the weights into the UMMA operand layouts if the architecture is one the kernels are specialised for.
int tc_prepare(cutdet_net *net);
void tc_destroy(cutdet_net *net);
bool tc_supported(const cutdet_net *net, int height, int width);
size_t tc_workspace_bytes(const cutdet_net *net, int batch, int height, int width);
int tc_forward_f32(cutdet_net *net, const float *x, int batch, int height, int width, float *logits, char *ws,
                   cudaStream_t stream);
// Training-mode forward (BatchNorm on batch statistics) on the tensor-core path: one sub-batch at most.
bool tc_batchstats_supported(const cutdet_net *net, int batch, int height, int width);
int tc_forward_f32_batchstats(cutdet_net *net, const float *x, int batch, int height, int width, float *out, char *ws,
                              cudaStream_t stream);
int tc_forward_frames(cutdet_net *net, const cutdet_resize_plan *plan, const cutdet_frames *src, float *logits,
                      char *ws, cudaStream_t stream);
int tc_debug_conv_output(cutdet_net *net, int layer, int batch, int height, int width, const char *ws, float *out,
                         cudaStream_t stream);

}  // namespace cutdet
