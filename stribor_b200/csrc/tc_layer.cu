// tcgen05 / TMA path -- placeholder until the tensor-core kernel lands: reports "unsupported"
// so every layer takes the generic path.
#include "common.cuh"
namespace stb {
bool tc_layer_supported(const stb_layer*) { return false; }
uint64_t tc_packed_bytes(const stb_layer*) { return 0; }
int tc_pack_layer(const stb_layer*, void*, cudaStream_t) { return set_error(STB_ENOTSUP, "tensor path not built"); }
int tc_layer_apply(const stb_layer*, int, const float*, float*, float*, int, int, int64_t, cudaStream_t) {
    return set_error(STB_ENOTSUP, "tensor path not built");
}
}  // namespace stb
