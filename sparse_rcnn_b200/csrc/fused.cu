// Multi-kernel entry points: one C-ABI call enqueues every kernel of a residual unit's forward or backward.
// The step was 100 % host bound (profiles/r1_e_launches_fused.md); these calls replace ~10 Python->C round trips per
// unit by 2 without changing which kernels run.
#include "common.cuh"

using namespace scn;

#define SCN_TRY(call)             \
    do {                          \
        int rc__ = (call);        \
        if (rc__) return rc__;    \
    } while (0)

extern "C" {

int scn_residual_unit_fwd(const float* x, int n, int C, const int32_t* map, int K, const float* w1, const float* b1,
                          const float* w2, const float* b2, void* img1, void* img2, int repack, float* r, float* h, float* y,
                          int use_tf32, scn_stream_t stream) {
    SCN_REQUIRE(n >= 0 && C > 0 && K > 0, "residual_unit_fwd: bad shape");
    // the caller has already recorded the images as packed for the current weight version: pack BEFORE any early return
    // (an empty crop would otherwise leave an unpacked image marked fresh for the next, non-empty call)
    if (use_tf32 && repack) {
        SCN_TRY(scn_conv_pack_weights(w1, K, C, C, 0, 0, img1, stream));
        SCN_TRY(scn_conv_pack_weights(w2, K, C, C, 0, 0, img2, stream));
    }
    if (n == 0) return SCN_OK;
    const int64_t total = (int64_t)n * C;
    SCN_TRY(scn_relu_fwd(x, r, total, use_tf32 ? 1 : 0, stream));
    if (use_tf32) {
        SCN_TRY(scn_conv_fwd_tf32(r, C, C, n, map, n, K, img1, b1, nullptr, 0, nullptr, 0, h, C, C, SCN_EPI_RELU | SCN_EPI_ROUND,
                                  stream));
        SCN_TRY(scn_conv_fwd_tf32(h, C, C, n, map, n, K, img2, b2, x, C, nullptr, 0, y, C, C, SCN_EPI_ADD, stream));
    } else {
        SCN_TRY(scn_conv_fwd_fp32(r, C, C, map, n, K, w1, 0, 0, b1, nullptr, 0, nullptr, 0, h, C, C, SCN_EPI_RELU, stream));
        SCN_TRY(scn_conv_fwd_fp32(h, C, C, map, n, K, w2, 0, 0, b2, x, C, nullptr, 0, y, C, C, SCN_EPI_ADD, stream));
    }
    return SCN_OK;
}

int scn_residual_unit_bwd(const float* gy, const float* r, const float* h, int n, int C, const int32_t* map, int K,
                          const float* w1, const float* w2, void* img1t, void* img2t, int repack, float* gyr, float* gh,
                          float* gx, float* gw1, float* gb1, float* gw2, float* gb2, int accumulate, int use_tf32,
                          scn_stream_t stream) {
    SCN_REQUIRE(n >= 0 && C > 0 && K > 0, "residual_unit_bwd: bad shape");
    cudaStream_t st = as_stream(stream);
    const size_t wbytes = (size_t)K * C * C * sizeof(float);
    // accumulate != 0: gw* / gb* are the parameters' gradient buffers (already zeroed or holding earlier
    // contributions); the weight-gradient kernel adds atomically anyway, so nothing is cleared and the bias sums add
    if (!accumulate) {
        if (gw1) cudaMemsetAsync(gw1, 0, wbytes, st);
        if (gw2) cudaMemsetAsync(gw2, 0, wbytes, st);
    }
    if (use_tf32 && repack) {      // before any early return: the caller has marked the images packed
        SCN_TRY(scn_conv_pack_weights(w1, K, C, C, 1, 1, img1t, stream));
        SCN_TRY(scn_conv_pack_weights(w2, K, C, C, 1, 1, img2t, stream));
    }
    if (n == 0) {
        if (!accumulate) {
            if (gb1) cudaMemsetAsync(gb1, 0, C * sizeof(float), st);
            if (gb2) cudaMemsetAsync(gb2, 0, C * sizeof(float), st);
        }
        return check_launch("residual_unit_bwd(memset)");
    }
    auto bias_sum = accumulate ? scn_col_sum_add : scn_col_sum;
    const int64_t total = (int64_t)n * C;
    const float* g_op = gy;      // operand of the transposed convolutions / weight gradients
    if (use_tf32) {
        SCN_TRY(scn_round_tf32(gy, gyr, total, stream));
        g_op = gyr;
        // d/dh through conv2, masked by relu'(h), rounded so that it feeds the next MMAs directly
        SCN_TRY(scn_conv_fwd_tf32(g_op, C, C, n, map, n, K, img2t, nullptr, nullptr, 0, h, C, gh, C, C,
                                  SCN_EPI_MASK | SCN_EPI_ROUND, stream));
    } else {
        SCN_TRY(scn_conv_fwd_fp32(g_op, C, C, map, n, K, w2, 1, 1, nullptr, nullptr, 0, h, C, gh, C, C, SCN_EPI_MASK, stream));
    }
    // bias gradients ride in the weight-gradient kernel (column sums of its grad-out tiles, added to gb)
    if (!accumulate) {
        if (gb1 && gw1) cudaMemsetAsync(gb1, 0, C * sizeof(float), st);
        if (gb2 && gw2) cudaMemsetAsync(gb2, 0, C * sizeof(float), st);
    }
    if (gw2) SCN_TRY(scn_conv_bwd_weight(h, C, C, map, n, K, g_op, C, C, gw2, gb2, use_tf32, stream));
    else if (gb2) SCN_TRY(bias_sum(gy, C, n, C, gb2, stream));
    if (gx) {
        // d/dx = gy + relu'(x) * conv1^T(gh)
        if (use_tf32)
            SCN_TRY(scn_conv_fwd_tf32(gh, C, C, n, map, n, K, img1t, nullptr, gy, C, r, C, gx, C, C, SCN_EPI_MASK | SCN_EPI_ADD,
                                      stream));
        else
            SCN_TRY(scn_conv_fwd_fp32(gh, C, C, map, n, K, w1, 1, 1, nullptr, gy, C, r, C, gx, C, C, SCN_EPI_MASK | SCN_EPI_ADD,
                                      stream));
    }
    if (gw1) SCN_TRY(scn_conv_bwd_weight(r, C, C, map, n, K, gh, C, C, gw1, gb1, use_tf32, stream));
    else if (gb1) SCN_TRY(bias_sum(gh, C, n, C, gb1, stream));
    return SCN_OK;
}

// One convolution layer (SubM / strided / transposed / 1x1) per direction as ONE call: operand rounding, weight packing
// when stale, the gather-GEMM(s) and the weight + bias gradient.  Same kernels as the separate entry points.
int scn_conv_layer_fwd(const float* x, int ld_x, int n_in, int Cin, int x_exact, float* x_round, const int32_t* map, int n_out,
                       int K, const float* w, void* image, int repack, const float* bias, float* out, int Cout, int use_tf32,
                       scn_stream_t stream) {
    SCN_REQUIRE(Cin > 0 && Cout > 0 && K > 0 && n_in >= 0 && n_out >= 0, "conv_layer_fwd: bad shape");
    if (use_tf32 && repack) SCN_TRY(scn_conv_pack_weights(w, K, Cin, Cout, 0, 0, image, stream));      // before any early return
    if (n_out == 0) return SCN_OK;
    if (!use_tf32)
        return scn_conv_fwd_fp32(x, ld_x, Cin, map, n_out, K, w, 0, 0, bias, nullptr, 0, nullptr, 0, out, Cout, Cout, 0, stream);
    const float* xin = x;
    int ld = ld_x;
    if (!x_exact) {      // tcgen05 truncates fp32 operands: round the gathered operand to nearest once
        SCN_REQUIRE(x_round && ld_x == Cin, "conv_layer_fwd: rounding needs a scratch buffer and a dense input");
        SCN_TRY(scn_round_tf32(x, x_round, (int64_t)n_in * Cin, stream));
        xin = x_round, ld = Cin;
    }
    return scn_conv_fwd_tf32(xin, ld, Cin, n_in, map, n_out, K, image, bias, nullptr, 0, nullptr, 0, out, Cout, Cout, 0, stream);
}

int scn_conv_layer_bwd(const float* go, int n_out, int Cout, int go_exact, float* go_round, const float* x, int ld_x, int n_in,
                       int Cin, const int32_t* fmap, const int32_t* bmap, int K, const float* w, void* image_t, int repack,
                       int reverse_bwd, float* gx, float* gw, float* gb, int use_tf32, scn_stream_t stream) {
    SCN_REQUIRE(Cin > 0 && Cout > 0 && K > 0 && n_in >= 0 && n_out >= 0, "conv_layer_bwd: bad shape");
    // the caller has marked the transposed image packed: pack it before any early return (empty crops)
    if (use_tf32 && repack && image_t) SCN_TRY(scn_conv_pack_weights(w, K, Cout, Cin, 1, reverse_bwd, image_t, stream));
    if (n_out == 0) {
        if (gx && n_in > 0) cudaMemsetAsync(gx, 0, sizeof(float) * (size_t)n_in * Cin, as_stream(stream));
        return check_launch("conv_layer_bwd(memset)");
    }
    const float* g = go;
    if (use_tf32 && !go_exact && (gx || gw)) {
        SCN_REQUIRE(go_round, "conv_layer_bwd: rounding needs a scratch buffer");
        SCN_TRY(scn_round_tf32(go, go_round, (int64_t)n_out * Cout, stream));
        g = go_round;
    }
    if (gx && n_in > 0) {
        // d/dx: the transposed (and, for submanifold layers, offset-reversed) weights over the backward map
        if (use_tf32) {
            SCN_TRY(scn_conv_fwd_tf32(g, Cout, Cout, n_out, bmap, n_in, K, image_t, nullptr, nullptr, 0, nullptr, 0, gx, Cin, Cin,
                                      0, stream));
        } else {
            SCN_TRY(scn_conv_fwd_fp32(g, Cout, Cout, bmap, n_in, K, w, 1, reverse_bwd, nullptr, nullptr, 0, nullptr, 0, gx, Cin,
                                      Cin, 0, stream));
        }
    }
    if (gw) SCN_TRY(scn_conv_bwd_weight(x, ld_x, Cin, fmap, n_out, K, g, Cout, Cout, gw, gb, use_tf32, stream));
    else if (gb) SCN_TRY(scn_col_sum_add(go, Cout, n_out, Cout, gb, stream));
    return SCN_OK;
}

}  // extern "C"
