// Elementwise, IO-layer, pooling, sparse-to-dense and mask-crop kernels.
// All of these are HBM-bound: every kernel streams each tensor once with coalesced (and, where
// the row width allows, 128-bit) accesses; grids are sized as multiples of the SM count.
#include <float.h>
#include <mutex>
#include <unordered_map>
#include "common.cuh"

namespace scn {

constexpr int TB = 256;

// ------------------------------------------------------------------------------ elementwise
__device__ __forceinline__ float rna_tf32(float v) {
    uint32_t t;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v));
    return __uint_as_float(t);
}
// MODE 0: relu, 1: relu + tf32 rounding, 2: tf32 rounding only
template <int MODE>
__device__ __forceinline__ float ew(float v) {
    if (MODE != 2) v = fmaxf(v, 0.f);
    if (MODE != 0) v = rna_tf32(v);
    return v;
}
template <int MODE>
__global__ void k_relu_fwd(const float* __restrict__ in, float* __restrict__ out, int64_t n) {
    pdl_trigger();      // PDL (common.cuh): let the convolution that follows start its prologue
    pdl_wait();
    int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
    bool vec = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    for (; i < n; i += stride) {
        if (vec && i + 3 < n) {
            float4 v = *reinterpret_cast<const float4*>(in + i);
            v.x = ew<MODE>(v.x), v.y = ew<MODE>(v.y), v.z = ew<MODE>(v.z), v.w = ew<MODE>(v.w);
            *reinterpret_cast<float4*>(out + i) = v;
        } else {
            for (int j = 0; j < 4 && i + j < n; ++j) out[i + j] = ew<MODE>(in[i + j]);
        }
    }
}
template <bool ROUND>
__global__ void k_relu_bwd(const float* __restrict__ y, const float* __restrict__ go, float* __restrict__ gi, int64_t n) {
    pdl_trigger();
    pdl_wait();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float g = y[i] > 0.f ? go[i] : 0.f;
        gi[i] = ROUND ? rna_tf32(g) : g;
    }
}
__global__ void k_add(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i < n; i += (int64_t)gridDim.x * blockDim.x) o[i] = a[i] + b[i];
}

// Column reductions (bias gradients, BatchNorm statistics): ONE launch, deterministic.  Blocks (32 channels x 8 row
// lanes) write slab partials; the last block of every channel group to finish (ticket counter, self resetting) adds the
// partials in slab order.
template <bool SQ_DIFF>
__global__ void k_col_reduce(const float* __restrict__ in, int ld, int n, int C, const float* __restrict__ mean, float scale,
                             float* __restrict__ partial, unsigned int* __restrict__ tickets, float* __restrict__ out, int accumulate) {
    __shared__ float sm[8][33];
    __shared__ bool last;
    const int c = blockIdx.y * 32 + threadIdx.x;
    float acc = 0.f;
    const float m = (SQ_DIFF && c < C) ? mean[c] : 0.f;
    if (c < C) {
        // four independent row streams per thread: four loads in flight instead of one (HBM-latency bound otherwise)
        const int step = gridDim.x * 8;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int r = blockIdx.x * 8 + threadIdx.y;
        for (; r + 3 * step < n; r += 4 * step) {
            float v0 = in[(int64_t)r * ld + c], v1 = in[(int64_t)(r + step) * ld + c];
            float v2 = in[(int64_t)(r + 2 * step) * ld + c], v3 = in[(int64_t)(r + 3 * step) * ld + c];
            if (SQ_DIFF) {
                v0 -= m, v1 -= m, v2 -= m, v3 -= m;
                v0 *= v0, v1 *= v1, v2 *= v2, v3 *= v3;
            }
            a0 += v0, a1 += v1, a2 += v2, a3 += v3;
        }
        for (; r < n; r += step) {
            float v = in[(int64_t)r * ld + c];
            if (SQ_DIFF) {
                v -= m;
                v *= v;
            }
            a0 += v;
        }
        acc = (a0 + a1) + (a2 + a3);
    }
    sm[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) s += sm[j][threadIdx.x];
        partial[(int64_t)blockIdx.x * C + c] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        const unsigned int t = atomicAdd(tickets + blockIdx.y, 1u);
        last = (t == gridDim.x - 1);
        if (last) tickets[blockIdx.y] = 0;      // ready for the next launch on this stream
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    float s = 0.f;
    if (c < C)
        for (int j = threadIdx.y; j < (int)gridDim.x; j += 8) s += __ldcg(partial + (int64_t)j * C + c);
    sm[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        float t = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) t += sm[j][threadIdx.x];
        out[c] = accumulate ? out[c] + t * scale : t * scale;
    }
}

__global__ void k_bn_apply(const float* __restrict__ in, int64_t total, int C, const float* __restrict__ mean,
                           const float* __restrict__ var, const float* __restrict__ gamma,
                           const float* __restrict__ beta, float eps, float leak, float* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        float y = (in[i] - mean[c]) * rsqrtf(var[c] + eps);
        if (gamma) y = y * gamma[c] + beta[c];
        out[i] = y > 0.f ? y : y * leak;
    }
}
// dz = dy * (y>0 ? 1 : leak); partial sums of dz and dz*xhat per channel
__global__ void k_bn_bwd_partial(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ dy,
                                 int n, int C, const float* __restrict__ mean, const float* __restrict__ var, float eps,
                                 float leak, float* __restrict__ partial /* [slabs, 2, C] */) {
    __shared__ float s1[8][33], s2[8][33];
    int c = blockIdx.y * 32 + threadIdx.x;
    float a1 = 0.f, a2 = 0.f;
    if (c < C) {
        float m = mean[c], is = rsqrtf(var[c] + eps);
        for (int r = blockIdx.x * 8 + threadIdx.y; r < n; r += gridDim.x * 8) {
            int64_t i = (int64_t)r * C + c;
            float dz = dy[i] * (y[i] > 0.f ? 1.f : leak);
            a1 += dz;
            a2 += dz * (x[i] - m) * is;
        }
    }
    s1[threadIdx.y][threadIdx.x] = a1;
    s2[threadIdx.y][threadIdx.x] = a2;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        float t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) t1 += s1[j][threadIdx.x], t2 += s2[j][threadIdx.x];
        partial[((int64_t)blockIdx.x * 2 + 0) * C + c] = t1;
        partial[((int64_t)blockIdx.x * 2 + 1) * C + c] = t2;
    }
}
__global__ void k_bn_bwd_final(const float* __restrict__ partial, int nslab, int C, float* __restrict__ sums /*[2,C]*/) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float t1 = 0.f, t2 = 0.f;
    for (int j = 0; j < nslab; ++j) {
        t1 += partial[((int64_t)j * 2 + 0) * C + c];
        t2 += partial[((int64_t)j * 2 + 1) * C + c];
    }
    sums[c] = t1;
    sums[C + c] = t2;
}
__global__ void k_bn_bwd_dx(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ dy, int n,
                            int C, const float* __restrict__ mean, const float* __restrict__ var,
                            const float* __restrict__ gamma, float eps, float leak, int training,
                            const float* __restrict__ sums, float* __restrict__ dx) {
    int64_t total = (int64_t)n * C;
    float inv_n = 1.f / (float)n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        float is = rsqrtf(var[c] + eps);
        float g = gamma ? gamma[c] : 1.f;
        float dz = dy[i] * (y[i] > 0.f ? 1.f : leak);
        if (training) {
            float xh = (x[i] - mean[c]) * is;
            dx[i] = g * is * (dz - sums[c] * inv_n - xh * sums[C + c] * inv_n);
        } else {
            dx[i] = g * is * dz;
        }
    }
}

// ------------------------------------------------------------------------------ io layers
// CSR reduce: one warp per output row when C >= 32, else sub-warp groups.  lanes stride channels.
__global__ void k_input_fwd(const float* __restrict__ feats, int ld, int C, const int32_t* __restrict__ row_ptr,
                            const int32_t* __restrict__ row_pts, int N, int mode, float* __restrict__ out) {
    int64_t total = (int64_t)N * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int r = (int)(i / C), c = (int)(i % C);
        int b = row_ptr[r], e = row_ptr[r + 1];
        float v = 0.f;
        if (e > b) {
            if (mode == 1) {
                v = feats[(int64_t)row_pts[e - 1] * ld + c];
            } else if (mode == 2) {
                v = feats[(int64_t)row_pts[b] * ld + c];
            } else {
                for (int j = b; j < e; ++j) v += feats[(int64_t)row_pts[j] * ld + c];
                if (mode == 4) v = v / (float)(e - b);
            }
        }
        out[i] = v;
    }
}
__global__ void k_input_bwd(const float* __restrict__ go, int C, const int32_t* __restrict__ point_row,
                            const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ row_pts, int P, int mode,
                            float* __restrict__ gf) {
    int64_t total = (int64_t)P * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int p = (int)(i / C), c = (int)(i % C);
        int r = point_row[p];
        float g = go[(int64_t)r * C + c];
        int b = row_ptr[r], e = row_ptr[r + 1];
        if (mode == 4) g = g / (float)(e - b);
        else if (mode == 1) g = (row_pts[e - 1] == p) ? g : 0.f;
        else if (mode == 2) g = (row_pts[b] == p) ? g : 0.f;
        gf[i] = g;
    }
}
__global__ void k_gather_rows(const float* __restrict__ in, int ld_in, const int32_t* __restrict__ idx, int n, int C,
                              float* __restrict__ out, int ld_out) {
    int64_t total = (int64_t)n * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int r = (int)(i / C), c = (int)(i % C);
        int s = idx[r];
        out[(int64_t)r * ld_out + c] = s >= 0 ? in[(int64_t)s * ld_in + c] : 0.f;
    }
}
// 128-bit variant: C % 4 == 0, ld % 4 == 0, 16-byte aligned bases
__global__ void k_gather_rows_v4(const float4* __restrict__ in, int ld4_in, const int32_t* __restrict__ idx, int n, int C4,
                                 float4* __restrict__ out, int ld4_out) {
    int64_t total = (int64_t)n * C4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int r = (int)(i / C4), c = (int)(i % C4);
        int s = idx[r];
        out[(int64_t)r * ld4_out + c] = s >= 0 ? __ldg(in + (int64_t)s * ld4_in + c) : make_float4(0, 0, 0, 0);
    }
}
__global__ void k_scatter_add_rows(const float* __restrict__ in, const int32_t* __restrict__ idx, int n, int C,
                                   float* __restrict__ out) {
    int64_t total = (int64_t)n * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int r = (int)(i / C), c = (int)(i % C);
        int d = idx[r];
        if (d >= 0) atomicAdd(out + (int64_t)d * C + c, in[i]);
    }
}

// dense-stationary sparse-to-dense: thread per dense cell (z fastest => coalesced writes per channel);
// fuses the zero fill with the scatter.  Reads of the feature row are 4-byte scattered but L2 resident.
__global__ void k_s2d_fwd(const float* __restrict__ in, int C, const uint64_t* __restrict__ tk,
                          const int32_t* __restrict__ tv, uint32_t mask, int B, int X, int Y, int Z,
                          float* __restrict__ out) {
    int64_t vol = (int64_t)X * Y * Z, cells = vol * B;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += (int64_t)gridDim.x * blockDim.x) {
        int b = (int)(i / vol);
        int64_t s = i % vol;
        int z = (int)(s % Z), y = (int)((s / Z) % Y), x = (int)(s / ((int64_t)Z * Y));
        int r = hash_lookup(tk, tv, mask, make_key(x, y, z, b));
        float* o = out + (int64_t)b * C * vol + s;
        if (r >= 0) {
            const float* f = in + (int64_t)r * C;
            for (int c = 0; c < C; ++c) o[(int64_t)c * vol] = __ldg(f + c);
        } else {
            for (int c = 0; c < C; ++c) o[(int64_t)c * vol] = 0.f;
        }
    }
}
__global__ void k_s2d_bwd(const float* __restrict__ gd, const uint64_t* __restrict__ row_keys, int N, int C, int X, int Y,
                          int Z, float* __restrict__ gi) {
    int64_t vol = (int64_t)X * Y * Z, total = (int64_t)N * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int r = (int)(i / C), c = (int)(i % C);
        uint64_t k = row_keys[r];
        int64_t s = ((int64_t)key_x(k) * Y + key_y(k)) * Z + key_z(k);
        gi[i] = gd[((int64_t)key_b(k) * C + c) * vol + s];
    }
}

// ------------------------------------------------------------------------------ pooling
__global__ void k_pool_fwd(const float* __restrict__ in, int C, const int32_t* __restrict__ cmap, int n_out, int K,
                           int is_max, float inv_vol, float* __restrict__ out) {
    int64_t total = (int64_t)n_out * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int r = (int)(i / C), c = (int)(i % C);
        float acc = is_max ? -INFINITY : 0.f;
        for (int o = 0; o < K; ++o) {
            int s = cmap[(int64_t)o * n_out + r];
            if (s >= 0) {
                float v = in[(int64_t)s * C + c];
                acc = is_max ? fmaxf(acc, v) : acc + v;
            }
        }
        out[i] = is_max ? acc : acc * inv_vol;
    }
}
__global__ void k_pool_bwd(const float* __restrict__ in, const float* __restrict__ out, const float* __restrict__ go,
                           int C, const int32_t* __restrict__ parent, int n_in, int is_max, float inv_vol,
                           float* __restrict__ gi) {
    int64_t total = (int64_t)n_in * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int r = (int)(i / C), c = (int)(i % C);
        int64_t j = (int64_t)parent[r] * C + c;
        gi[i] = is_max ? (in[i] == out[j] ? go[j] : 0.f) : go[j] * inv_vol;
    }
}
// one block per (segment, 32-channel group): threads (32 channels, 8 row lanes)
__global__ void k_segment_mean_fwd(const float* __restrict__ in, int C, const int32_t* __restrict__ seg_ptr,
                                   float* __restrict__ out) {
    __shared__ float sm[8][33];
    int seg = blockIdx.x, c = blockIdx.y * 32 + threadIdx.x;
    int b = seg_ptr[seg], e = seg_ptr[seg + 1];
    float acc = 0.f;
    if (c < C)
        for (int r = b + threadIdx.y; r < e; r += 8) acc += in[(int64_t)r * C + c];
    sm[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) s += sm[j][threadIdx.x];
        out[(int64_t)seg * C + c] = e > b ? s / (float)(e - b) : 0.f;
    }
}
__global__ void k_segment_mean_bwd(const float* __restrict__ go, int C, const int32_t* __restrict__ seg_ptr,
                                   float* __restrict__ gi) {
    int seg = blockIdx.x;
    int b = seg_ptr[seg], e = seg_ptr[seg + 1];
    if (e <= b) return;
    float inv = 1.f / (float)(e - b);
    int64_t total = (int64_t)(e - b) * C;
    for (int64_t i = threadIdx.x; i < total; i += blockDim.x) {
        int c = (int)(i % C);
        gi[(int64_t)b * C + i] = go[(int64_t)seg * C + c] * inv;
    }
}

// ------------------------------------------------------------------------------ mask crop
// Box-vs-point test with half-open integer boxes; only the points of the box's own sample are
// visited (points are grouped by sample), i.e. BB * P_sample tests instead of the reference's
// dense BB x P x 4 broadcast.  Two passes (count, ordered compaction) keep (box, point) order.
__device__ __forceinline__ bool inside_box(uint64_t k, const int32_t* __restrict__ bx) {
    int x = key_x(k), y = key_y(k), z = key_z(k);
    return x >= bx[0] && x < bx[3] && y >= bx[1] && y < bx[4] && z >= bx[2] && z < bx[5];
}
__global__ void __launch_bounds__(256) k_crop_count(const uint64_t* __restrict__ keys, const int32_t* __restrict__ sample_ptr,
                                                    const int32_t* __restrict__ boxes, const int32_t* __restrict__ box_sample,
                                                    int n_chunks, int32_t* __restrict__ counts) {
    int box = blockIdx.y, chunk = blockIdx.x;
    __shared__ int bx[6];
    __shared__ int wsum[8];
    if (threadIdx.x < 6) bx[threadIdx.x] = boxes[box * 6 + threadIdx.x];
    __syncthreads();
    int s = box_sample[box];
    int p0 = sample_ptr[s] + chunk * SCN_CROP_CHUNK, p1 = min(sample_ptr[s + 1], p0 + SCN_CROP_CHUNK);
    int cnt = 0;
    for (int p = p0 + threadIdx.x; p < p1; p += 256) cnt += inside_box(keys[p], bx) ? 1 : 0;
    for (int d = 16; d; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int j = 0; j < 8; ++j) t += wsum[j];
        counts[box * n_chunks + chunk] = t;
    }
}
__global__ void __launch_bounds__(256) k_crop_select(const uint64_t* __restrict__ keys, const int32_t* __restrict__ sample_ptr,
                                                     const int32_t* __restrict__ boxes, const int32_t* __restrict__ box_sample,
                                                     int n_chunks, const int32_t* __restrict__ offsets, int P,
                                                     int32_t* __restrict__ sel_pt, uint64_t* __restrict__ new_keys,
                                                     uint8_t* __restrict__ is_inside) {
    int box = blockIdx.y, chunk = blockIdx.x;
    __shared__ int bx[6];
    __shared__ int wcnt[8];
    __shared__ int base;
    // most (box, chunk) pairs select nothing (a box covers a small part of its scene): the count pass already knows
    if (offsets[box * n_chunks + chunk + 1] == offsets[box * n_chunks + chunk] && !is_inside) return;
    if (threadIdx.x < 6) bx[threadIdx.x] = boxes[box * 6 + threadIdx.x];
    if (threadIdx.x == 0) base = offsets[box * n_chunks + chunk];
    __syncthreads();
    int s = box_sample[box];
    int p0 = sample_ptr[s] + chunk * SCN_CROP_CHUNK, p1 = min(sample_ptr[s + 1], p0 + SCN_CROP_CHUNK);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int it = 0; it < SCN_CROP_CHUNK; it += 256) {
        int p = p0 + it + threadIdx.x;
        uint64_t k = 0;
        bool in = false;
        if (p < p1) {
            k = keys[p];
            in = inside_box(k, bx);
            if (is_inside) is_inside[(int64_t)box * P + p] = in ? 1 : 0;
        }
        unsigned bal = __ballot_sync(0xffffffffu, in);
        if (lane == 0) wcnt[w] = __popc(bal);
        __syncthreads();
        int woff = 0, tot = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int cj = wcnt[j];
            if (j < w) woff += cj;
            tot += cj;
        }
        if (in) {
            int dst = base + woff + __popc(bal & ((1u << lane) - 1));
            sel_pt[dst] = p;
            new_keys[dst] = (k & 0x0000FFFFFFFFFFFFull) | ((uint64_t)box << 48);
        }
        __syncthreads();
        if (threadIdx.x == 0) base += tot;
        __syncthreads();
        if (p0 + it + 256 >= p1) break;
    }
}

}  // namespace scn

using namespace scn;

// ---------------------------------------------------------------- per-box (segment) BCE with logits
// torch's formulation (aten binary_cross_entropy_with_logits): (1 - t) x + m + log(exp(-m) + exp(-x - m)), m = max(-x, 0)
__device__ __forceinline__ float bce_logits(float x, float t) {
    const float m = fmaxf(-x, 0.f);
    return (1.f - t) * x + m + logf(expf(-m) + expf(-x - m));
}
// one block per segment; fixed-shape tree reduction => deterministic
__global__ void k_segment_bce_fwd(const float* __restrict__ x, const uint8_t* __restrict__ t, const int32_t* __restrict__ ptr,
                                  float* __restrict__ out) {
    __shared__ float sm[256];
    const int s = blockIdx.x, lo = ptr[s], hi = ptr[s + 1];
    float acc = 0.f;
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) acc += bce_logits(x[i], t[i] ? 1.f : 0.f);
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int d = 128; d > 0; d >>= 1) {
        if (threadIdx.x < d) sm[threadIdx.x] += sm[threadIdx.x + d];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[s] = sm[0] / (float)(hi - lo);      // 0/0 = NaN for an empty box, like torch.mean
}
__global__ void k_segment_bce_bwd(const float* __restrict__ x, const uint8_t* __restrict__ t, const int32_t* __restrict__ ptr,
                                  const float* __restrict__ g, float* __restrict__ gx) {
    const int s = blockIdx.x, lo = ptr[s], hi = ptr[s + 1];
    if (hi <= lo) return;
    const float scale = g[s] / (float)(hi - lo);
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const float sig = 1.f / (1.f + expf(-x[i]));
        gx[i] = scale * (sig - (t[i] ? 1.f : 0.f));
    }
}

// ---------------------------------------------------------------- point-wise cross entropy
// one thread per row (C is the number of classes, ~20): max, log-sum-exp, weighted negative log-likelihood
__global__ void k_ce_rows(const float* __restrict__ x, int ld, int64_t n, int C, const int64_t* __restrict__ y,
                          const float* __restrict__ w, int64_t ignore, float* __restrict__ lse, float* __restrict__ rows) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float* xi = x + i * ld;
        float m = -INFINITY;
        for (int c = 0; c < C; ++c) m = fmaxf(m, xi[c]);
        float sum = 0.f;
        for (int c = 0; c < C; ++c) sum += __expf(xi[c] - m);
        const float l = m + __logf(sum);
        lse[i] = l;
        const int64_t yi = y[i];
        float wi = 0.f, li = 0.f;
        if (yi != ignore && yi >= 0 && yi < C) {
            wi = w ? w[yi] : 1.f;
            li = wi * (l - xi[yi]);
        }
        rows[2 * i] = li;
        rows[2 * i + 1] = wi;
    }
}
__global__ void k_ce_bwd(const float* __restrict__ x, int ld, int64_t n, int C, const int64_t* __restrict__ y,
                         const float* __restrict__ w, int64_t ignore, const float* __restrict__ lse,
                         const float* __restrict__ stats, const float* __restrict__ g, float* __restrict__ dx) {
    const float scale = g[0] / stats[1];
    const int64_t total = n * C;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / C;
        const int c = (int)(e - i * C);
        const int64_t yi = y[i];
        float v = 0.f;
        if (yi != ignore && yi >= 0 && yi < C) {
            const float wi = w ? w[yi] : 1.f;
            v = scale * wi * (__expf(x[i * ld + c] - lse[i]) - (c == yi ? 1.f : 0.f));
        }
        dx[e] = v;
    }
}

static int col_reduce(const float* in, int ld, int n, int C, const float* mean, float scale, float* out, float* partial,
                      unsigned int* tickets, int nslab, cudaStream_t st, int accumulate = 0) {
    int slabs = cdiv(n, 64);
    if (slabs > nslab) slabs = nslab;
    if (slabs < 1) slabs = 1;
    dim3 grid(slabs, cdiv(C, 32)), block(32, 8);
    if (mean)
        k_col_reduce<true><<<grid, block, 0, st>>>(in, ld, n, C, mean, scale, partial, tickets, out, accumulate);
    else
        k_col_reduce<false><<<grid, block, 0, st>>>(in, ld, n, C, nullptr, scale, partial, tickets, out, accumulate);
    return check_launch("col_reduce");
}

// scratch + ticket counters of the two-stage column reductions: one set PER STREAM (host threads drive one stream each in
// SparseInference.run_many, the executor runs weight / bias gradients on a side stream), created under a mutex and only ever
// grown after the stream has drained (an earlier launch may still read the old buffer).  ADVICE r1.
struct ColScratch {
    float* buf = nullptr;
    size_t bytes = 0;
    unsigned int* tickets = nullptr;      // 64 self-resetting ticket counters (bias sums / BN mean use 0..31, BN variance 32..63)
};
static int col_scratch(cudaStream_t st, size_t bytes, ColScratch* out) {
    static std::mutex mu;
    static std::unordered_map<cudaStream_t, ColScratch> pool;
    std::lock_guard<std::mutex> lock(mu);
    ColScratch& w = pool[st];
    if (!w.tickets) {
        if (cudaMalloc(&w.tickets, 64 * sizeof(unsigned int)) != cudaSuccess ||
            cudaMemsetAsync(w.tickets, 0, 64 * sizeof(unsigned int), st) != cudaSuccess) {
            cudaGetLastError();
            w.tickets = nullptr;
            set_error("ticket alloc failed");
            return SCN_ERR_CUDA;
        }
    }
    if (bytes > w.bytes) {
        if (w.buf) {
            cudaStreamSynchronize(st);
            cudaFree(w.buf);
        }
        w.buf = nullptr, w.bytes = 0;
        cudaError_t e = cudaMalloc(&w.buf, bytes);
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("scratch alloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
            return SCN_ERR_CUDA;
        }
        w.bytes = bytes;
    }
    *out = w;
    return SCN_OK;
}
constexpr int NSLAB = 296;  // 2 x 148 SMs

extern "C" {

// SCN_EW_PDL=0: launch the elementwise kernels without the programmatic-dependent-launch attribute (A/B switch)
static bool ew_pdl() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SCN_EW_PDL");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}
int scn_relu_fwd(const float* in, float* out, int64_t n, int round_tf32, scn_stream_t stream) {
    if (n <= 0) return SCN_OK;
    if (round_tf32)
        {
            PdlMaskScope ew_scope;
            if (!ew_pdl()) ew_scope.exclude(as_stream(stream));
            PdlLaunch L(dim3(grid_for((n + 3) / 4, TB)), dim3(TB), 0, as_stream(stream));
            cudaLaunchKernelEx(&L.cfg, k_relu_fwd<1>, in, out, n);
        }
    else
        {
            PdlMaskScope ew_scope;
            if (!ew_pdl()) ew_scope.exclude(as_stream(stream));
            PdlLaunch L(dim3(grid_for((n + 3) / 4, TB)), dim3(TB), 0, as_stream(stream));
            cudaLaunchKernelEx(&L.cfg, k_relu_fwd<0>, in, out, n);
        }
    return check_launch("relu_fwd");
}
int scn_relu_bwd(const float* y, const float* go, float* gi, int64_t n, int round_tf32, scn_stream_t stream) {
    if (n <= 0) return SCN_OK;
    if (round_tf32)
        {
            PdlMaskScope ew_scope;
            if (!ew_pdl()) ew_scope.exclude(as_stream(stream));
            PdlLaunch L(dim3(grid_for(n, TB)), dim3(TB), 0, as_stream(stream));
            cudaLaunchKernelEx(&L.cfg, k_relu_bwd<true>, y, go, gi, n);
        }
    else
        {
            PdlMaskScope ew_scope;
            if (!ew_pdl()) ew_scope.exclude(as_stream(stream));
            PdlLaunch L(dim3(grid_for(n, TB)), dim3(TB), 0, as_stream(stream));
            cudaLaunchKernelEx(&L.cfg, k_relu_bwd<false>, y, go, gi, n);
        }
    return check_launch("relu_bwd");
}
int scn_round_tf32(const float* in, float* out, int64_t n, scn_stream_t stream) {
    if (n <= 0) return SCN_OK;
    {
        PdlMaskScope ew_scope;
        if (!ew_pdl()) ew_scope.exclude(as_stream(stream));
        PdlLaunch L(dim3(grid_for((n + 3) / 4, TB)), dim3(TB), 0, as_stream(stream));
        cudaLaunchKernelEx(&L.cfg, k_relu_fwd<2>, in, out, n);
    }
    return check_launch("round_tf32");
}
int scn_add(const float* a, const float* b, float* out, int64_t n, scn_stream_t stream) {
    if (n <= 0) return SCN_OK;
    k_add<<<grid_for(n, TB), TB, 0, as_stream(stream)>>>(a, b, out, n);
    return check_launch("add");
}
int scn_col_sum(const float* in, int ld, int n, int C, float* out, scn_stream_t stream) {
    SCN_REQUIRE(C > 0 && n >= 0, "col_sum: bad shape");
    ColScratch ws;
    int rc = col_scratch(as_stream(stream), (size_t)NSLAB * 2 * C * sizeof(float), &ws);
    if (rc) return rc;
    SCN_REQUIRE(C <= 2048, "col_sum: C > 2048 not supported");
    return col_reduce(in, ld, n, C, nullptr, 1.f, out, ws.buf, ws.tickets, NSLAB, as_stream(stream));
}
int scn_col_sum_add(const float* in, int ld, int n, int C, float* out, scn_stream_t stream) {
    SCN_REQUIRE(C > 0 && n >= 0, "col_sum_add: bad shape");
    if (n == 0) return SCN_OK;
    ColScratch ws;
    int rc = col_scratch(as_stream(stream), (size_t)NSLAB * 2 * C * sizeof(float), &ws);
    if (rc) return rc;
    SCN_REQUIRE(C <= 2048, "col_sum_add: C > 2048 not supported");
    return col_reduce(in, ld, n, C, nullptr, 1.f, out, ws.buf, ws.tickets, NSLAB, as_stream(stream), 1);
}
int scn_bn_stats(const float* in, int n, int C, float* mean, float* var, scn_stream_t stream) {
    SCN_REQUIRE(C > 0 && n > 0, "bn_stats: needs at least one active row");
    SCN_REQUIRE(C <= 1024, "bn_stats: C > 1024 not supported (mean and variance share the 64 ticket counters)");
    ColScratch ws;
    int rc = col_scratch(as_stream(stream), (size_t)NSLAB * 2 * C * sizeof(float), &ws);
    if (rc) return rc;
    rc = col_reduce(in, C, n, C, nullptr, 1.f / n, mean, ws.buf, ws.tickets, NSLAB, as_stream(stream));
    if (rc) return rc;
    return col_reduce(in, C, n, C, mean, 1.f / n, var, ws.buf + (size_t)NSLAB * C, ws.tickets + 32, NSLAB, as_stream(stream));
}
int scn_bn_apply(const float* in, int n, int C, const float* mean, const float* var, const float* gamma,
                 const float* beta, float eps, float leak, float* out, scn_stream_t stream) {
    if (n <= 0) return SCN_OK;
    k_bn_apply<<<grid_for((int64_t)n * C, TB), TB, 0, as_stream(stream)>>>(in, (int64_t)n * C, C, mean, var, gamma, beta,
                                                                            eps, leak, out);
    return check_launch("bn_apply");
}
int scn_bn_bwd(const float* x, const float* y, const float* dy, int n, int C, const float* mean, const float* var,
               const float* gamma, float eps, float leak, int training, float* dx, float* dgamma, float* dbeta,
               float* tmp2C, scn_stream_t stream) {
    if (n <= 0) return SCN_OK;
    ColScratch ws;
    int rc = col_scratch(as_stream(stream), (size_t)NSLAB * 2 * C * sizeof(float), &ws);
    if (rc) return rc;
    cudaStream_t st = as_stream(stream);
    dim3 grid(NSLAB, cdiv(C, 32)), block(32, 8);
    k_bn_bwd_partial<<<grid, block, 0, st>>>(x, y, dy, n, C, mean, var, eps, leak, ws.buf);
    rc = check_launch("bn_bwd_partial");
    if (rc) return rc;
    k_bn_bwd_final<<<cdiv(C, 128), 128, 0, st>>>(ws.buf, NSLAB, C, tmp2C);
    rc = check_launch("bn_bwd_final");
    if (rc) return rc;
    k_bn_bwd_dx<<<grid_for((int64_t)n * C, TB), TB, 0, st>>>(x, y, dy, n, C, mean, var, gamma, eps, leak, training, tmp2C, dx);
    rc = check_launch("bn_bwd_dx");
    if (rc) return rc;
    if (dbeta) cudaMemcpyAsync(dbeta, tmp2C, C * sizeof(float), cudaMemcpyDeviceToDevice, st);
    if (dgamma) cudaMemcpyAsync(dgamma, tmp2C + C, C * sizeof(float), cudaMemcpyDeviceToDevice, st);
    return check_launch("bn_bwd_copy");
}

int scn_input_fwd(const float* feats, int ld, int C, const int32_t* row_ptr, const int32_t* row_pts, int N, int mode,
                  float* out, scn_stream_t stream) {
    SCN_REQUIRE(mode >= 1 && mode <= 4, "input_fwd: mode must be 1..4 (mode 0 is a plain copy)");
    if (N <= 0) return SCN_OK;
    k_input_fwd<<<grid_for((int64_t)N * C, TB), TB, 0, as_stream(stream)>>>(feats, ld, C, row_ptr, row_pts, N, mode, out);
    return check_launch("input_fwd");
}
int scn_input_bwd(const float* go, int C, const int32_t* point_row, const int32_t* row_ptr, const int32_t* row_pts, int P,
                  int mode, float* gf, scn_stream_t stream) {
    SCN_REQUIRE(mode >= 1 && mode <= 4, "input_bwd: mode must be 1..4");
    if (P <= 0) return SCN_OK;
    k_input_bwd<<<grid_for((int64_t)P * C, TB), TB, 0, as_stream(stream)>>>(go, C, point_row, row_ptr, row_pts, P, mode, gf);
    return check_launch("input_bwd");
}
int scn_gather_rows(const float* in, int ld_in, const int32_t* idx, int n, int C, float* out, int ld_out,
                    scn_stream_t stream) {
    if (n <= 0) return SCN_OK;
    bool v4 = (C % 4 == 0) && (ld_in % 4 == 0) && (ld_out % 4 == 0) &&
              ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    if (v4)
        k_gather_rows_v4<<<grid_for((int64_t)n * C / 4, TB), TB, 0, as_stream(stream)>>>(
            reinterpret_cast<const float4*>(in), ld_in / 4, idx, n, C / 4, reinterpret_cast<float4*>(out), ld_out / 4);
    else
        k_gather_rows<<<grid_for((int64_t)n * C, TB), TB, 0, as_stream(stream)>>>(in, ld_in, idx, n, C, out, ld_out);
    return check_launch("gather_rows");
}
int scn_scatter_add_rows(const float* in, const int32_t* idx, int n, int C, float* out, scn_stream_t stream) {
    if (n <= 0) return SCN_OK;
    k_scatter_add_rows<<<grid_for((int64_t)n * C, TB), TB, 0, as_stream(stream)>>>(in, idx, n, C, out);
    return check_launch("scatter_add_rows");
}
int scn_sparse_to_dense_fwd(const float* in, int C, const uint64_t* tk, const int32_t* tv, uint32_t cap, int B, int X,
                            int Y, int Z, float* out, scn_stream_t stream) {
    SCN_REQUIRE(cap && !(cap & (cap - 1)), "hash: capacity must be a power of two (got %u)", cap);
    int64_t cells = (int64_t)B * X * Y * Z;
    if (cells <= 0) return SCN_OK;
    k_s2d_fwd<<<grid_for(cells, TB, 16), TB, 0, as_stream(stream)>>>(in, C, tk, tv, cap - 1, B, X, Y, Z, out);
    return check_launch("s2d_fwd");
}
int scn_sparse_to_dense_bwd(const float* gd, const uint64_t* row_keys, int N, int C, int X, int Y, int Z, float* gi,
                            scn_stream_t stream) {
    if (N <= 0) return SCN_OK;
    k_s2d_bwd<<<grid_for((int64_t)N * C, TB), TB, 0, as_stream(stream)>>>(gd, row_keys, N, C, X, Y, Z, gi);
    return check_launch("s2d_bwd");
}
int scn_pool_fwd(const float* in, int C, const int32_t* cmap, int n_out, int K, int is_max, float inv_volume, float* out,
                 scn_stream_t stream) {
    if (n_out <= 0) return SCN_OK;
    k_pool_fwd<<<grid_for((int64_t)n_out * C, TB), TB, 0, as_stream(stream)>>>(in, C, cmap, n_out, K, is_max, inv_volume, out);
    return check_launch("pool_fwd");
}
int scn_pool_bwd(const float* in, const float* out, const float* go, int C, const int32_t* parent_row, int n_in,
                 int is_max, float inv_volume, float* gi, scn_stream_t stream) {
    if (n_in <= 0) return SCN_OK;
    k_pool_bwd<<<grid_for((int64_t)n_in * C, TB), TB, 0, as_stream(stream)>>>(in, out, go, C, parent_row, n_in, is_max,
                                                                              inv_volume, gi);
    return check_launch("pool_bwd");
}
int scn_segment_bce_fwd(const float* logits, const uint8_t* targets, const int32_t* seg_ptr, int n_seg, float* mean_out,
                        scn_stream_t stream) {
    SCN_REQUIRE(n_seg >= 0, "segment_bce_fwd: bad shape");
    if (n_seg == 0) return SCN_OK;
    k_segment_bce_fwd<<<n_seg, 256, 0, as_stream(stream)>>>(logits, targets, seg_ptr, mean_out);
    return check_launch("segment_bce_fwd");
}
int scn_segment_bce_bwd(const float* logits, const uint8_t* targets, const int32_t* seg_ptr, int n_seg,
                        const float* grad_mean, float* grad_logits, scn_stream_t stream) {
    SCN_REQUIRE(n_seg >= 0, "segment_bce_bwd: bad shape");
    if (n_seg == 0) return SCN_OK;
    k_segment_bce_bwd<<<n_seg, 256, 0, as_stream(stream)>>>(logits, targets, seg_ptr, grad_mean, grad_logits);
    return check_launch("segment_bce_bwd");
}
int scn_cross_entropy_fwd(const float* logits, int ld, int64_t n, int C, const int64_t* labels, const float* weight,
                          int64_t ignore_index, float* lse, float* row_scratch, float* stats, scn_stream_t stream) {
    SCN_REQUIRE(n > 0 && C > 0 && ld >= C, "cross_entropy_fwd: bad shape");
    SCN_REQUIRE(n < (int64_t)1 << 31, "cross_entropy_fwd: too many rows");
    k_ce_rows<<<grid_for(n, TB), TB, 0, as_stream(stream)>>>(logits, ld, n, C, labels, weight, ignore_index, lse, row_scratch);
    int rc = check_launch("cross_entropy_rows");
    if (rc) return rc;
    ColScratch ws;
    rc = col_scratch(as_stream(stream), (size_t)NSLAB * 2 * 2 * sizeof(float), &ws);
    if (rc) return rc;
    return col_reduce(row_scratch, 2, (int)n, 2, nullptr, 1.f, stats, ws.buf, ws.tickets, NSLAB, as_stream(stream));
}
int scn_cross_entropy_bwd(const float* logits, int ld, int64_t n, int C, const int64_t* labels, const float* weight,
                          int64_t ignore_index, const float* lse, const float* stats, const float* grad_loss, float* dlogits,
                          scn_stream_t stream) {
    SCN_REQUIRE(n > 0 && C > 0 && ld >= C, "cross_entropy_bwd: bad shape");
    k_ce_bwd<<<grid_for(n * C, TB), TB, 0, as_stream(stream)>>>(logits, ld, n, C, labels, weight, ignore_index, lse, stats,
                                                               grad_loss, dlogits);
    return check_launch("cross_entropy_bwd");
}
int scn_segment_mean_fwd(const float* in, int C, const int32_t* seg_ptr, int n_seg, float* out, scn_stream_t stream) {
    if (n_seg <= 0) return SCN_OK;
    dim3 grid(n_seg, cdiv(C, 32)), block(32, 8);
    k_segment_mean_fwd<<<grid, block, 0, as_stream(stream)>>>(in, C, seg_ptr, out);
    return check_launch("segment_mean_fwd");
}
int scn_segment_mean_bwd(const float* go, int C, const int32_t* seg_ptr, int n_seg, float* gi, scn_stream_t stream) {
    if (n_seg <= 0) return SCN_OK;
    k_segment_mean_bwd<<<n_seg, 256, 0, as_stream(stream)>>>(go, C, seg_ptr, gi);
    return check_launch("segment_mean_bwd");
}
int scn_crop_count(const uint64_t* keys, const int32_t* sample_ptr, const int32_t* boxes, const int32_t* box_sample, int BB,
                   int n_chunks, int32_t* counts, scn_stream_t stream) {
    if (BB <= 0 || n_chunks <= 0) return SCN_OK;
    SCN_REQUIRE(BB <= 65535, "crop: at most 65535 boxes per call (got %d)", BB);
    dim3 grid(n_chunks, BB);
    k_crop_count<<<grid, 256, 0, as_stream(stream)>>>(keys, sample_ptr, boxes, box_sample, n_chunks, counts);
    return check_launch("crop_count");
}
int scn_crop_select(const uint64_t* keys, const int32_t* sample_ptr, const int32_t* boxes, const int32_t* box_sample, int BB,
                    int n_chunks, const int32_t* offsets, int P, int32_t* sel_pt, uint64_t* new_keys, uint8_t* is_inside,
                    scn_stream_t stream) {
    if (BB <= 0 || n_chunks <= 0) return SCN_OK;
    SCN_REQUIRE(BB <= 65535, "crop: at most 65535 boxes per call (got %d)", BB);
    dim3 grid(n_chunks, BB);
    k_crop_select<<<grid, 256, 0, as_stream(stream)>>>(keys, sample_ptr, boxes, box_sample, n_chunks, offsets, P, sel_pt,
                                                        new_keys, is_inside);
    return check_launch("crop_select");
}

}  // extern "C"
