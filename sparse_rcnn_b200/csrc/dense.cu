// Dense stage behind the sparse backbone: the region-proposal trunk the reference builds with
// get_dilation_network (module_factory.py:581-611) is SparseToDense followed by `num_dilations` x [Conv3d 3^3 'same'
// (dilated) + ReLU] on the dense grid.  Here the dense grid is kept CHANNELS-LAST -- one row of C floats per cell, cells in
// (b, x, y, z) order -- which is exactly the feature layout of the gather-GEMM kernels (conv_tc.cu, conv_wgrad_tc.cu): a
// dense 'same' convolution is the submanifold convolution over the trivial neighbour map  map[o][r] = r + delta(o)  (or -1
// outside the grid), so forward, input gradient and weight gradient of the trunk run on the same tcgen05 kernels, with the
// same epilogues (bias, ReLU, TF32 rounding), as the sparse layers.  This file holds what is new: the dense neighbour map,
// the zero-filling scatter of sparse rows into dense rows (and its gather backward), and the layout change to / from the
// [B, C, X, Y, Z] tensors the reference's anchor heads consume.
#include "common.cuh"

namespace scn {

constexpr int DB = 256;

__global__ void k_dense_map(int B, int X, int Y, int Z, int dil, int32_t* __restrict__ map) {
    const int64_t vol = (int64_t)X * Y * Z, N = vol * B, total = 27 * N;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int o = (int)(p / N);
        const int64_t r = p - (int64_t)o * N, s = r % vol;
        const int z = (int)(s % Z), y = (int)((s / Z) % Y), x = (int)(s / ((int64_t)Z * Y));
        const int qx = x + (o / 9 - 1) * dil, qy = y + ((o / 3) % 3 - 1) * dil, qz = z + (o % 3 - 1) * dil;
        const bool in = qx >= 0 && qy >= 0 && qz >= 0 && qx < X && qy < Y && qz < Z;
        map[p] = in ? (int32_t)(r + (((int64_t)(qx - x) * Y + (qy - y)) * Z + (qz - z))) : -1;
    }
}

// dense-stationary: one thread per (cell, 4-channel chunk); the zero fill is fused with the scatter
template <int VEC>
__global__ void k_s2d_rows_fwd(const float* __restrict__ in, int C, const uint64_t* __restrict__ tk, const int32_t* __restrict__ tv,
                               uint32_t mask, int B, int X, int Y, int Z, float* __restrict__ out) {
    const int CV = C / VEC;
    const int64_t vol = (int64_t)X * Y * Z, total = vol * B * CV;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t cell = i / CV;
        const int cv = (int)(i - cell * CV);
        const int b = (int)(cell / vol);
        const int64_t s = cell % vol;
        const int z = (int)(s % Z), y = (int)((s / Z) % Y), x = (int)(s / ((int64_t)Z * Y));
        const int r = hash_lookup(tk, tv, mask, make_key(x, y, z, b));
        if (VEC == 4) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r >= 0) v = __ldg(reinterpret_cast<const float4*>(in + (int64_t)r * C) + cv);
            reinterpret_cast<float4*>(out + cell * C)[cv] = v;
        } else {
            out[cell * C + cv] = r >= 0 ? __ldg(in + (int64_t)r * C + cv) : 0.f;
        }
    }
}
__global__ void k_s2d_rows_bwd(const float* __restrict__ gd, const uint64_t* __restrict__ row_keys, int N, int C, int X, int Y, int Z,
                               float* __restrict__ gi) {
    const int64_t vol = (int64_t)X * Y * Z, total = (int64_t)N * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / C), c = (int)(i - (int64_t)r * C);
        const uint64_t k = row_keys[r];
        const int64_t cell = (int64_t)key_b(k) * vol + ((int64_t)key_x(k) * Y + key_y(k)) * Z + key_z(k);
        gi[i] = gd[cell * C + c];
    }
}

// [B][vol][C] <-> [B][C][vol] through 32 x 32 shared-memory tiles (both sides coalesced)
__global__ void k_transpose_tiles(const float* __restrict__ in, int rows, int cols, float* __restrict__ out) {
    __shared__ float tile[32][33];
    const int64_t b = blockIdx.z;
    const float* src = in + b * (int64_t)rows * cols;
    float* dst = out + b * (int64_t)rows * cols;
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int r = r0 + j, c = c0 + threadIdx.x;
        tile[j][threadIdx.x] = (r < rows && c < cols) ? src[(int64_t)r * cols + c] : 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int c = c0 + j, r = r0 + threadIdx.x;
        if (r < rows && c < cols) dst[(int64_t)c * rows + r] = tile[threadIdx.x][j];
    }
}

}  // namespace scn

using namespace scn;

extern "C" {

int scn_dense_map(int B, int X, int Y, int Z, int dilation, int32_t* map, scn_stream_t stream) {
    SCN_REQUIRE(B > 0 && X > 0 && Y > 0 && Z > 0 && dilation > 0 && map, "dense_map: bad arguments");
    SCN_REQUIRE((int64_t)B * X * Y * Z < (1ll << 31), "dense_map: more than 2^31 cells");
    k_dense_map<<<grid_for(27ll * B * X * Y * Z, DB, 8), DB, 0, as_stream(stream)>>>(B, X, Y, Z, dilation, map);
    return check_launch("dense_map");
}

int scn_sparse_to_dense_rows_fwd(const float* in, int C, const uint64_t* tk, const int32_t* tv, uint32_t cap, int B, int X, int Y,
                                 int Z, float* out, scn_stream_t stream) {
    SCN_REQUIRE(cap && !(cap & (cap - 1)), "hash: capacity must be a power of two (got %u)", cap);
    SCN_REQUIRE(C > 0, "sparse_to_dense_rows: bad channel count");
    const int64_t cells = (int64_t)B * X * Y * Z;
    if (cells <= 0) return SCN_OK;
    const bool v4 = C % 4 == 0 && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    if (v4) k_s2d_rows_fwd<4><<<grid_for(cells * (C / 4), DB, 8), DB, 0, as_stream(stream)>>>(in, C, tk, tv, cap - 1, B, X, Y, Z, out);
    else k_s2d_rows_fwd<1><<<grid_for(cells * C, DB, 8), DB, 0, as_stream(stream)>>>(in, C, tk, tv, cap - 1, B, X, Y, Z, out);
    return check_launch("s2d_rows_fwd");
}

int scn_sparse_to_dense_rows_bwd(const float* grad_dense, const uint64_t* row_keys, int N, int C, int X, int Y, int Z, float* grad_in,
                                 scn_stream_t stream) {
    if (N <= 0) return SCN_OK;
    k_s2d_rows_bwd<<<grid_for((int64_t)N * C, DB, 8), DB, 0, as_stream(stream)>>>(grad_dense, row_keys, N, C, X, Y, Z, grad_in);
    return check_launch("s2d_rows_bwd");
}

int scn_transpose_batched(const float* in, int batches, int rows, int cols, float* out, scn_stream_t stream) {
    SCN_REQUIRE(batches >= 0 && rows >= 0 && cols >= 0, "transpose: bad shape");
    if (!batches || !rows || !cols) return SCN_OK;
    SCN_REQUIRE(batches <= 65535 && (rows + 31) / 32 <= 65535, "transpose: grid too large");
    dim3 grid((cols + 31) / 32, (rows + 31) / 32, batches);
    k_transpose_tiles<<<grid, dim3(32, 8), 0, as_stream(stream)>>>(in, rows, cols, out);
    return check_launch("transpose");
}

}  // extern "C"
