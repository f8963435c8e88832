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
    if (n == 0) return SCN_OK;
    const int64_t total = (int64_t)n * C;
    SCN_TRY(scn_relu_fwd(x, r, total, use_tf32 ? 1 : 0, stream));
    if (use_tf32) {
        if (repack) {
            SCN_TRY(scn_conv_pack_weights(w1, K, C, C, 0, 0, img1, stream));
            SCN_TRY(scn_conv_pack_weights(w2, K, C, C, 0, 0, img2, stream));
        }
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
        if (repack) {
            SCN_TRY(scn_conv_pack_weights(w1, K, C, C, 1, 1, img1t, stream));
            SCN_TRY(scn_conv_pack_weights(w2, K, C, C, 1, 1, img2t, stream));
        }
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

}  // extern "C"
