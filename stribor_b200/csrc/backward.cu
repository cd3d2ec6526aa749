// backward kernels -- placeholder
#include "common.cuh"
namespace stb {
uint64_t layer_backward_workspace_bytes(const stb_layer*, int64_t) { return 0; }
int layer_backward(const stb_layer*, int, const float*, const float*, const float*, const float*,
                   const float*, float*, float*, float*, const stb_layer_grads*, void*, int64_t,
                   cudaStream_t) {
    return set_error(STB_ENOTSUP, "backward not built yet");
}
}  // namespace stb
