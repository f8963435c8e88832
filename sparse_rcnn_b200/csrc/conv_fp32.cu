// Exact-fp32 implicit gather-GEMM (verification mode, <= 1e-5 vs the oracle) and the weight
// gradient.  Output-stationary: a CTA owns 64 output rows x 64 output channels, loops the kernel
// offsets, gathers the contributing input rows through the neighbour map into shared memory and
// accumulates with FFMA in registers (4x4 per thread).  No atomics on the forward / input-gradient
// path: every output row is written exactly once.
#include "common.cuh"

namespace scn {

constexpr int BM = 64, BN = 64, BK = 16;

// W_eff[o][ci][co] of the (possibly transposed / offset-reversed) raw weight tensor w[K][A][B]
struct WView {
    const float* w;
    int K, A, B, transpose, reverse;
    __device__ __forceinline__ float at(int o, int ci, int co) const {
        int oo = reverse ? K - 1 - o : o;
        const float* p = w + (int64_t)oo * A * B;
        return transpose ? p[(int64_t)co * B + ci] : p[(int64_t)ci * B + co];
    }
};

__global__ void __launch_bounds__(256) k_conv_fp32(const float* __restrict__ in, int ld_in, int Cin,
                                                   const int32_t* __restrict__ map, int n_out, int K, WView wv,
                                                   const float* __restrict__ bias, const float* __restrict__ residual,
                                                   int ld_res, const float* __restrict__ mask, int ld_mask,
                                                   float* __restrict__ out, int ld_out, int Cout, int epi) {
    __shared__ float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN];
    __shared__ int s_row[BM];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int row0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int o = 0; o < K; ++o) {
        int any = 0;
        if (tid < BM) {
            int r = row0 + tid, s = -1;
            if (r < n_out) s = map ? map[(int64_t)o * n_out + r] : r;
            s_row[tid] = s;
            any = s >= 0;
        }
        any = __syncthreads_or(any);
        if (!any) continue;
        for (int k0 = 0; k0 < Cin; k0 += BK) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int e = tid + j * 256;
                int r = e >> 4, k = e & 15;
                int s = s_row[r];
                float v = 0.f;
                if (s >= 0 && k0 + k < Cin) v = __ldg(in + (int64_t)s * ld_in + k0 + k);
                As[k][r] = v;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int e = tid + j * 256;
                int k = e >> 6, n = e & 63;
                float v = 0.f;
                if (k0 + k < Cin && n0 + n < Cout) v = wv.at(o, k0 + k, n0 + n);
                Bs[k][n] = v;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < BK; ++k) {
                float a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
                float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
                b[0] = bv.x, b[1] = bv.y, b[2] = bv.z, b[3] = bv.w;
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int r = row0 + ty * 4 + i;
        if (r >= n_out) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int c = n0 + tx * 4 + j;
            if (c >= Cout) continue;
            float v = acc[i][j];
            if (bias) v += bias[c];
            if ((epi & SCN_EPI_MASK) && !(mask[(int64_t)r * ld_mask + c] > 0.f)) v = 0.f;
            if (epi & SCN_EPI_ADD) v += residual[(int64_t)r * ld_res + c];
            if (epi & SCN_EPI_RELU) v = fmaxf(v, 0.f);
            out[(int64_t)r * ld_out + c] = v;       // SCN_EPI_ROUND is a no-op in the exact-fp32 mode
        }
    }
}

// grad_w[o][ci][co] += sum_{r in chunk} in[map[o][r]][ci] * go[r][co]
// grid (row chunks, K, ci-blocks * co-blocks); 64x64 tile per CTA, rows streamed 16 at a time.
constexpr int WG_ROWS = 2048;
__global__ void __launch_bounds__(256) k_conv_bwd_weight(const float* __restrict__ in, int ld_in, int Cin,
                                                         const int32_t* __restrict__ map, int n_out, int K,
                                                         const float* __restrict__ go, int ld_go, int Cout,
                                                         float* __restrict__ gw) {
    __shared__ __align__(16) float As[BK][BM];
    __shared__ __align__(16) float Gs[BK][BN];
    __shared__ int s_row[BK];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int o = blockIdx.y;
    const int n_cb = (Cout + BN - 1) / BN;
    const int ci0 = (blockIdx.z / n_cb) * BM, co0 = (blockIdx.z % n_cb) * BN;
    const int r_begin = blockIdx.x * WG_ROWS, r_end = min(n_out, r_begin + WG_ROWS);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    bool touched = false;
    for (int r0 = r_begin; r0 < r_end; r0 += BK) {
        int any = 0;
        if (tid < BK) {
            int r = r0 + tid, s = -1;
            if (r < r_end) s = map ? map[(int64_t)o * n_out + r] : r;
            s_row[tid] = s;
            any = s >= 0;
        }
        any = __syncthreads_or(any);
        if (!any) continue;
        touched = true;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int e = tid + j * 256;
            int k = e >> 6, c = e & 63;
            int s = s_row[k];
            float a = 0.f, g = 0.f;
            if (s >= 0) {
                if (ci0 + c < Cin) a = __ldg(in + (int64_t)s * ld_in + ci0 + c);
                if (co0 + c < Cout) g = __ldg(go + (int64_t)(r0 + k) * ld_go + co0 + c);
            }
            As[k][c] = a;
            Gs[k][c] = g;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            float4 gv = *reinterpret_cast<const float4*>(&Gs[k][tx * 4]);
            float a[4] = {av.x, av.y, av.z, av.w}, g[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], g[j], acc[i][j]);
        }
        __syncthreads();
    }
    if (!touched) return;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int ci = ci0 + ty * 4 + i;
        if (ci >= Cin) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int co = co0 + tx * 4 + j;
            if (co < Cout) atomicAdd(gw + ((int64_t)o * Cin + ci) * Cout + co, acc[i][j]);
        }
    }
}

}  // namespace scn

using namespace scn;

extern "C" {

int scn_conv_fwd_fp32(const float* in, int ld_in, int Cin, const int32_t* map, int n_out, int K, const float* w,
                      int transpose, int reverse, const float* bias, const float* residual, int ld_res,
                      const float* mask, int ld_mask, float* out, int ld_out, int Cout, int epi_flags,
                      scn_stream_t stream) {
    SCN_REQUIRE(Cin > 0 && Cout > 0 && K > 0, "conv_fwd_fp32: bad shape Cin=%d Cout=%d K=%d", Cin, Cout, K);
    SCN_REQUIRE(map || K == 1, "conv_fwd_fp32: identity map requires K == 1");
    SCN_REQUIRE(!(epi_flags & SCN_EPI_ADD) || residual, "conv_fwd_fp32: SCN_EPI_ADD needs a residual pointer");
    SCN_REQUIRE(!(epi_flags & SCN_EPI_MASK) || mask, "conv_fwd_fp32: SCN_EPI_MASK needs a mask pointer");
    if (n_out <= 0) return SCN_OK;
    WView wv;
    wv.w = w, wv.K = K, wv.transpose = transpose, wv.reverse = reverse;
    // raw tensor is [K, A, B]; without transpose A=Cin,B=Cout; with transpose A=Cout,B=Cin
    wv.A = transpose ? Cout : Cin;
    wv.B = transpose ? Cin : Cout;
    dim3 grid(cdiv(n_out, BM), cdiv(Cout, BN));
    k_conv_fp32<<<grid, 256, 0, as_stream(stream)>>>(in, ld_in, Cin, map, n_out, K, wv, bias, residual, ld_res, mask, ld_mask,
                                                      out, ld_out, Cout, epi_flags);
    return check_launch("conv_fwd_fp32");
}

int scn_conv_bwd_weight_fp32(const float* in, int ld_in, int Cin, const int32_t* map, int n_out, int K,
                             const float* grad_out, int ld_go, int Cout, float* grad_w, scn_stream_t stream) {
    SCN_REQUIRE(Cin > 0 && Cout > 0 && K > 0, "conv_bwd_weight: bad shape");
    SCN_REQUIRE(map || K == 1, "conv_bwd_weight: identity map requires K == 1");
    SCN_REQUIRE(K <= 65535, "conv_bwd_weight: K too large");
    if (n_out <= 0) return SCN_OK;
    dim3 grid(cdiv(n_out, WG_ROWS), K, cdiv(Cin, BM) * cdiv(Cout, BN));
    k_conv_bwd_weight<<<grid, 256, 0, as_stream(stream)>>>(in, ld_in, Cin, map, n_out, K, grad_out, ld_go, Cout, grad_w);
    return check_launch("conv_bwd_weight");
}

}  // extern "C"
