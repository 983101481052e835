// Tensor-core path: placeholder until the tcgen05 kernels land (everything runs on the generic kernels).
#include "conv_tc.cuh"

namespace cutdet {

int tc_prepare(cutdet_net *) { return CUTDET_OK; }
void tc_destroy(cutdet_net *) {}
bool tc_supported(const cutdet_net *, int, int) { return false; }
size_t tc_workspace_bytes(const cutdet_net *, int, int, int) { return 0; }
int tc_forward_f32(cutdet_net *, const float *, int, int, int, float *, char *, cudaStream_t) {
    return fail(CUTDET_EUNSUPPORTED, "tensor-core path not built");
}
int tc_forward_frames(cutdet_net *, const cutdet_resize_plan *, const cutdet_frames *, float *, char *, cudaStream_t) {
    return fail(CUTDET_EUNSUPPORTED, "tensor-core path not built");
}
int tc_debug_conv_output(cutdet_net *, int, int, int, int, const char *, float *, cudaStream_t) {
    return fail(CUTDET_EUNSUPPORTED, "tensor-core path not built");
}

}  // namespace cutdet
