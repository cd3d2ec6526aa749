// extern "C" surface of libstribor_b200.so (see include/stribor_b200.h).
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include "common.cuh"

namespace stb {

static thread_local char g_err[512] = "";
static thread_local uint64_t g_launches = 0;

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
void count_launch() { ++g_launches; }

static int apply_one(const stb_layer* L, int direction, const float* x, const float* latent,
                     const float* t, float* y, float* ldj, int ldj_mode, int base_lp, int64_t rows,
                     cudaStream_t s, float* ldiag = nullptr, int32_t* bins = nullptr) {
    int rc = validate_layer(L);
    if (rc) return rc;
    if (rows < 0) return set_error(STB_EINVAL, "rows < 0");
    if (rows == 0) return STB_OK;
    if (!x || !y) return set_error(STB_EINVAL, "x / y is NULL");
    if (bins) {                                   // -1 everywhere the kernels do not search
        cudaError_t e = cudaMemsetAsync(bins, 0xff, (size_t)rows * L->dim * sizeof(int32_t), s);
        if (e != cudaSuccess) return set_error(STB_ECUDA, "memset: %s", cudaGetErrorString(e));
    }
    if (L->kind >= STB_PERMUTE) {
        if (ldiag) return set_error(STB_ENOTSUP, "no per-dimension log-derivative for this layer kind");
        if (ldj_mode != STB_LDJ_NONE && !ldj) return set_error(STB_EINVAL, "ldj_mode set but ldj is NULL");
        return pointwise_layer_apply(L, direction, x, y, ldj, ldj_mode, base_lp, rows, s);
    }
    if (L->latent_dim > 0 && !latent) return set_error(STB_EINVAL, "layer expects a latent input");
    if ((L->kind == STB_CONT_AFFINE) && !t) return set_error(STB_EINVAL, "layer expects a time input");
    if (ldj_mode != STB_LDJ_NONE && !ldj) return set_error(STB_EINVAL, "ldj_mode set but ldj is NULL");
    if (L->packed && !ldiag && tc_layer_supported(L))
        return tc_layer_apply(L, direction, x, y, ldj, ldj_mode, base_lp, rows, s, bins);
    if (L->packed && !ldiag && tcw_layer_supported(L) && tcw_image_present(L))
        return tcw_layer_apply(L, tcw_image(L), direction, x, y, ldj, ldj_mode, base_lp, rows, s, bins);
    if (L->packed && !ldiag && tcm_layer_supported(L))
        return tcm_layer_apply(L, direction, x, t, y, ldj, ldj_mode, base_lp, rows, s);
    return generic_layer_apply(L, direction, x, latent, t, y, ldj, ldj_mode, base_lp, ldiag, rows, s, bins);
}

// A sequence of layers (already in application order).  Maximal runs of 2..8 packed spline couplings of the
// 256-row tensor-core kernel with the same dim and kind go out as ONE launch each (tc_layer.cu, CHAIN: the
// tile stays in shared memory between the layers); everything else layer by layer.
// STRIBOR_B200_NO_CHAIN=1: one launch per layer, for comparison.
static int apply_sequence(const stb_layer* const* seq, int n, int direction, const float* x, const float* latent,
                          const float* t, float* out, float* ldj, int ldj_mode, int base_lp_last, int64_t rows,
                          cudaStream_t s) {
    static const bool chain_off = [] { const char* e = getenv("STRIBOR_B200_NO_CHAIN"); return e && e[0] == '1'; }();
    const float* cur = x;
    int mode = ldj_mode;
    int i = 0;
    while (i < n) {
        int j = i + 1;
        bool mlp_chain = false;
        if (!chain_off && rows > 0 && x && out && !(ldj_mode != STB_LDJ_NONE && !ldj) && validate_layer(seq[i]) == 0) {
            while (j < n && j - i < 8 && validate_layer(seq[j]) == 0 && tc_chain_supported(seq + i, j - i + 1)) ++j;
            if (j == i + 1) {               // affine / continuous-affine couplings with small conditioners (tc_mlp.cu)
                while (j < n && j - i < 8 && validate_layer(seq[j]) == 0 && tcm_chain_supported(seq + i, j - i + 1)) ++j;
                mlp_chain = j - i >= 2;
            }
        }
        const int last = (j == n);
        int rc;
        if (mlp_chain) {
            rc = tcm_chain_apply(seq + i, j - i, direction, cur, t, out, ldj, ldj ? mode : STB_LDJ_NONE,
                                 last && base_lp_last, rows, s);
        } else if (j - i >= 2) {
            rc = tc_chain_apply(seq + i, j - i, direction, cur, out, ldj, ldj ? mode : STB_LDJ_NONE,
                                last && base_lp_last, rows, s);
        } else {
            if (seq[i]->kind == STB_PERMUTE && cur == out)
                return set_error(STB_ENOTSUP, "a permutation inside a fused chain needs out != x for that hop (use the layer-by-layer path)");
            rc = apply_one(seq[i], direction, cur, latent, t, out, ldj, ldj ? mode : STB_LDJ_NONE,
                           last && base_lp_last, rows, s);
        }
        if (rc) return rc;
        cur = out;
        if (mode == STB_LDJ_SET) mode = STB_LDJ_ADD;       // later layers accumulate
        i = j;
    }
    return STB_OK;
}

}  // namespace stb

using namespace stb;

extern "C" {

int stb_abi_version(void) { return STB_ABI_VERSION; }
const char* stb_last_error(void) { return g_err; }
uint64_t stb_launch_count(void) { return g_launches; }
uint64_t stb_sizeof_layer(void) { return sizeof(stb_layer); }

int stb_layer_apply(const stb_layer* layer, int direction, const float* x, const float* latent,
                    const float* t, float* y, float* ldj, int ldj_mode, int base_log_prob,
                    int64_t rows, void* stream) {
    if (direction != STB_FORWARD && direction != STB_INVERSE) return set_error(STB_EINVAL, "bad direction");
    return apply_one(layer, direction, x, latent, t, y, ldj, ldj_mode, base_log_prob, rows,
                     (cudaStream_t)stream);
}

int stb_layer_apply_bins(const stb_layer* layer, int direction, const float* x, const float* latent,
                         const float* t, float* y, float* ldj, int ldj_mode, int32_t* bins,
                         int64_t rows, void* stream) {
    if (direction != STB_FORWARD && direction != STB_INVERSE) return set_error(STB_EINVAL, "bad direction");
    if (!bins) return set_error(STB_EINVAL, "bins is NULL");
    return apply_one(layer, direction, x, latent, t, y, ldj, ldj_mode, 0, rows, (cudaStream_t)stream, nullptr, bins);
}

int stb_layer_apply_diag(const stb_layer* layer, int direction, const float* x, const float* latent,
                         const float* t, float* y, float* ldiag, int64_t rows, void* stream) {
    if (direction != STB_FORWARD && direction != STB_INVERSE) return set_error(STB_EINVAL, "bad direction");
    if (!ldiag) return set_error(STB_EINVAL, "ldiag is NULL");
    return apply_one(layer, direction, x, latent, t, y, nullptr, STB_LDJ_NONE, 0, rows,
                     (cudaStream_t)stream, ldiag);
}

int stb_flow_apply(const stb_layer* layers, int n_layers, int direction, const float* x,
                   const float* latent, const float* t, float* out, float* ldj, int ldj_mode,
                   int64_t rows, void* stream) {
    if (n_layers < 0 || (n_layers > 0 && !layers)) return set_error(STB_EINVAL, "bad layer list");
    if (direction != STB_FORWARD && direction != STB_INVERSE) return set_error(STB_EINVAL, "bad direction");
    cudaStream_t s = (cudaStream_t)stream;
    if (n_layers == 0) return set_error(STB_EINVAL, "empty flow: nothing to apply");
    if (n_layers > 64) return set_error(STB_EINVAL, "more than 64 layers");
    const stb_layer* order[64];
    for (int i = 0; i < n_layers; ++i) order[i] = &layers[direction == STB_FORWARD ? i : n_layers - 1 - i];
    return apply_sequence(order, n_layers, direction, x, latent, t, out, ldj, ldj_mode, 0, rows, s);
}

int stb_flow_log_prob(const stb_layer* layers, int n_layers, const float* y, const float* latent,
                      const float* t, float* x_out, float* lp, int64_t rows, void* stream) {
    if (n_layers < 1 || !layers) return set_error(STB_EINVAL, "log_prob needs at least one layer");
    if (!x_out || !lp) return set_error(STB_EINVAL, "x_out / lp is NULL");
    cudaStream_t s = (cudaStream_t)stream;
    if (n_layers > 64) return set_error(STB_EINVAL, "more than 64 layers");
    const stb_layer* order[64];
    for (int i = 0; i < n_layers; ++i) order[i] = &layers[n_layers - 1 - i];
    return apply_sequence(order, n_layers, STB_INVERSE, y, latent, t, x_out, lp, STB_LDJ_SET, 1, rows, s);
}

int stb_unit_normal_log_prob(const float* x, float* lp, int accumulate, int32_t dim, int64_t rows,
                             void* stream) {
    if (!x || !lp || dim < 1 || rows < 0) return set_error(STB_EINVAL, "bad argument");
    return unit_normal_apply(x, lp, accumulate, dim, rows, (cudaStream_t)stream);
}

uint64_t stb_layer_backward_workspace_bytes(const stb_layer* layer, int64_t rows) {
    return layer_backward_workspace_bytes(layer, rows);
}

int stb_layer_backward(const stb_layer* layer, int direction, const float* x, const float* latent,
                       const float* t, const float* g_out, const float* g_ldj, float* g_x,
                       float* g_latent, float* g_t, const stb_layer_grads* grads, void* workspace,
                       int64_t rows, void* stream) {
    int rc = validate_layer(layer);
    if (rc) return rc;
    return layer_backward(layer, direction, x, latent, t, g_out, g_ldj, g_x, g_latent, g_t, grads,
                          workspace, rows, (cudaStream_t)stream);
}

int stb_layer_backward_diag(const stb_layer* layer, int direction, const float* x, const float* g_out,
                            const float* g_ldiag, float* g_x, const stb_layer_grads* grads, int64_t rows,
                            void* stream) {
    int rc = validate_layer(layer);
    if (rc) return rc;
    if (layer->net.n_linear > 0 || layer->kind >= STB_CONT_AFFINE)
        return set_error(STB_ENOTSUP, "stb_layer_backward_diag: affine / spline layers with row_out or const_out parameters only");
    return layer_backward(layer, direction, x, nullptr, nullptr, g_out, nullptr, g_x, nullptr, nullptr, grads,
                          nullptr, rows, (cudaStream_t)stream, g_ldiag);
}

uint64_t stb_packed_bytes(const stb_layer* layer) {
    if (validate_layer(layer)) return 0;
    if (tc_layer_supported(layer)) return tc_packed_bytes(layer) + tcw_packed_bytes(layer);   // both images
    if (tcw_layer_supported(layer)) return tcw_packed_bytes(layer);
    return tcm_layer_supported(layer) ? tcm_packed_bytes(layer) : 0;
}

int stb_pack_layer(const stb_layer* layer, void* packed_out, void* stream) {
    int rc = validate_layer(layer);
    if (rc) return rc;
    if (!packed_out) return set_error(STB_EINVAL, "packed_out is NULL");
    if (tc_layer_supported(layer)) {
        // the 256-row inference kernel's image, then the 128-row / backward kernel's (tc_wide.cu)
        rc = tc_pack_layer(layer, packed_out, (cudaStream_t)stream);
        if (rc) return rc;
        return tcw_pack_layer(layer, static_cast<uint8_t*>(packed_out) + tc_packed_bytes(layer), (cudaStream_t)stream);
    }
    if (tcw_layer_supported(layer)) return tcw_pack_layer(layer, packed_out, (cudaStream_t)stream);
    if (tcm_layer_supported(layer)) return tcm_pack_layer(layer, packed_out, (cudaStream_t)stream);
    return set_error(STB_ENOTSUP, "layer has no tensor-core path");
}

int stb_layer_uses_tensor_path(const stb_layer* layer) {
    if (validate_layer(layer)) return 0;
    return (layer->packed && (tc_layer_supported(layer) || tcw_layer_supported(layer) || tcm_layer_supported(layer))) ? 1 : 0;
}

}  // extern "C"
