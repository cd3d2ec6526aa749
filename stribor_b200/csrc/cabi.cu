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
                     cudaStream_t s, float* ldiag = nullptr) {
    int rc = validate_layer(L);
    if (rc) return rc;
    if (rows < 0) return set_error(STB_EINVAL, "rows < 0");
    if (rows == 0) return STB_OK;
    if (!x || !y) return set_error(STB_EINVAL, "x / y is NULL");
    if (L->kind >= STB_PERMUTE) {
        if (ldiag) return set_error(STB_ENOTSUP, "no per-dimension log-derivative for this layer kind");
        if (ldj_mode != STB_LDJ_NONE && !ldj) return set_error(STB_EINVAL, "ldj_mode set but ldj is NULL");
        return pointwise_layer_apply(L, direction, x, y, ldj, ldj_mode, base_lp, rows, s);
    }
    if (L->latent_dim > 0 && !latent) return set_error(STB_EINVAL, "layer expects a latent input");
    if ((L->kind == STB_CONT_AFFINE) && !t) return set_error(STB_EINVAL, "layer expects a time input");
    if (ldj_mode != STB_LDJ_NONE && !ldj) return set_error(STB_EINVAL, "ldj_mode set but ldj is NULL");
    if (L->packed && !ldiag && tc_layer_supported(L))
        return tc_layer_apply(L, direction, x, y, ldj, ldj_mode, base_lp, rows, s);
    if (L->packed && !ldiag && tcw_layer_supported(L) && tcw_image_present(L))
        return tcw_layer_apply(L, tcw_image(L), direction, x, y, ldj, ldj_mode, base_lp, rows, s);
    if (L->packed && !ldiag && tcm_layer_supported(L))
        return tcm_layer_apply(L, direction, x, t, y, ldj, ldj_mode, base_lp, rows, s);
    return generic_layer_apply(L, direction, x, latent, t, y, ldj, ldj_mode, base_lp, ldiag, rows, s);
}

// A whole flow in one launch: every layer a packed spline coupling of the 256-row tensor-core kernel, same
// dim and kind (STRIBOR_B200_NO_CHAIN=1: one launch per layer, for comparison)
static bool chain_fusable(const stb_layer* layers, int n) {
    if (n < 2 || n > 8) return false;
    static const bool off = [] { const char* e = getenv("STRIBOR_B200_NO_CHAIN"); return e && e[0] == '1'; }();
    if (off) return false;
    const stb_layer* p[8];
    for (int i = 0; i < n; ++i) {
        if (validate_layer(&layers[i])) return false;
        p[i] = &layers[i];
    }
    return tc_chain_supported(p, n);
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
    if (rows > 0 && x && out && chain_fusable(layers, n_layers) && !(ldj_mode != STB_LDJ_NONE && !ldj)) {
        const stb_layer* order[8];
        for (int i = 0; i < n_layers; ++i) order[i] = &layers[direction == STB_FORWARD ? i : n_layers - 1 - i];
        return tc_chain_apply(order, n_layers, direction, x, out, ldj, ldj ? ldj_mode : STB_LDJ_NONE, 0, rows, s);
    }
    const float* cur = x;
    int mode = ldj_mode;
    for (int i = 0; i < n_layers; ++i) {
        const stb_layer* L = &layers[direction == STB_FORWARD ? i : n_layers - 1 - i];
        if (L->kind == STB_PERMUTE && cur == out)
            return set_error(STB_ENOTSUP, "a permutation inside a fused chain needs out != x for that hop (use the layer-by-layer path)");
        int rc = apply_one(L, direction, cur, latent, t, out, ldj, ldj ? mode : STB_LDJ_NONE, 0, rows, s);
        if (rc) return rc;
        cur = out;
        if (mode == STB_LDJ_SET) mode = STB_LDJ_ADD;       // later layers accumulate
    }
    return STB_OK;
}

int stb_flow_log_prob(const stb_layer* layers, int n_layers, const float* y, const float* latent,
                      const float* t, float* x_out, float* lp, int64_t rows, void* stream) {
    if (n_layers < 1 || !layers) return set_error(STB_EINVAL, "log_prob needs at least one layer");
    if (!x_out || !lp) return set_error(STB_EINVAL, "x_out / lp is NULL");
    cudaStream_t s = (cudaStream_t)stream;
    if (rows > 0 && y && chain_fusable(layers, n_layers)) {
        const stb_layer* order[8];
        for (int i = 0; i < n_layers; ++i) order[i] = &layers[n_layers - 1 - i];
        return tc_chain_apply(order, n_layers, STB_INVERSE, y, x_out, lp, STB_LDJ_SET, 1, rows, s);
    }
    const float* cur = y;
    for (int i = 0; i < n_layers; ++i) {
        const stb_layer* L = &layers[n_layers - 1 - i];
        const int last = (i == n_layers - 1);
        if (L->kind == STB_PERMUTE && cur == x_out)
            return set_error(STB_ENOTSUP, "a permutation inside a fused chain needs out != x for that hop (use the layer-by-layer path)");
        int rc = apply_one(L, STB_INVERSE, cur, latent, t, x_out, lp, i == 0 ? STB_LDJ_SET : STB_LDJ_ADD,
                           last, rows, s);
        if (rc) return rc;
        cur = x_out;
    }
    return STB_OK;
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
