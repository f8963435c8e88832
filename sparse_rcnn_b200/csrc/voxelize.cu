// Voxelisation + collation on the device: the deterministic core of the reference's per-sample conversion
//   ndsis/data/sparse_augmentation.py:81-126 (augment_coords: project, shift to the origin, discretise, cut out),
//   :135-190 (augment_features: rows that stay, normals rotated, common noise), ndsis/data/data.py:88-115 (collate_fn: batch
//   index column, concatenation over the samples of a batch)
// for a whole batch at once.  The random draws of the reference (distortion matrix, sub-pixel offset, cut-out start, noise
// vectors) are INPUTS here: the host mirror (sparse_rcnn_b200/voxelize.py) draws them with the reference's own torch calls.
//
// Arithmetic is the reference's, bit for bit: a [P,3] x [3,3] fp32 product is x0*m0 rounded, then two fused multiply-adds
// (what torch's CPU matmul does for K = 3, pinned by the goldens of oracle/make_golden_voxelize.py); the shift is
// (-min + offset) in fp32, the discrete coordinate the truncation of the fp32 sum.  HBM-bound integer / copy work:
// thread per point, coalesced 12-byte rows, one warp-aggregated atomic per sample for the minimum, the library's own scan
// for the compaction; 44 algorithmic bytes per point for the coordinate path (12 in, 32 out) + 2 x 4 C for the features.
#include "common.cuh"

namespace scn {

constexpr int VOX_TB = 256;

__device__ __forceinline__ unsigned int f32_ordered(float f) {
    const unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float f32_unordered(unsigned int u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

// sample of point p (rows are grouped by sample; B is small): last b with sample_ptr[b] <= p
__device__ __forceinline__ int sample_of(const int* __restrict__ sample_ptr, int B, int p) {
    int lo = 0, hi = B;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(sample_ptr + mid) <= p) lo = mid;
        else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ void project(const float* __restrict__ m, float x, float y, float z, float& a0, float& a1, float& a2) {
    a0 = __fmaf_rn(z, m[6], __fmaf_rn(y, m[3], __fmul_rn(x, m[0])));
    a1 = __fmaf_rn(z, m[7], __fmaf_rn(y, m[4], __fmul_rn(x, m[1])));
    a2 = __fmaf_rn(z, m[8], __fmaf_rn(y, m[5], __fmul_rn(x, m[2])));
}

__global__ void k_vox_init(unsigned int* mins, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) mins[i] = 0xFFFFFFFFu;
}

// pass 1: projected coordinates [P,3] and their per-sample minimum (ordered-uint atomicMin, one per warp and dimension when
// the warp lies inside one sample)
__global__ void __launch_bounds__(VOX_TB) k_vox_project(const float* __restrict__ pts, int P, const int* __restrict__ sample_ptr, int B,
                                                        const float* __restrict__ proj, float* __restrict__ aug,
                                                        unsigned int* __restrict__ mins) {
    const int p = blockIdx.x * VOX_TB + threadIdx.x;
    const bool live = p < P;
    int b = 0;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    if (live) {
        b = sample_of(sample_ptr, B, p);
        const float* m = proj + 9 * b;
        float mm[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) mm[i] = __ldg(m + i);
        project(mm, pts[3 * (int64_t)p], pts[3 * (int64_t)p + 1], pts[3 * (int64_t)p + 2], a0, a1, a2);
        aug[3 * (int64_t)p] = a0, aug[3 * (int64_t)p + 1] = a1, aug[3 * (int64_t)p + 2] = a2;
    }
    const unsigned int act = __ballot_sync(0xffffffffu, live);
    if (!live) return;
    const int b0 = __shfl_sync(act, b, __ffs(act) - 1);
    const bool uniform = __all_sync(act, b == b0);
    unsigned int o0 = f32_ordered(a0), o1 = f32_ordered(a1), o2 = f32_ordered(a2);
    if (uniform) {
        o0 = __reduce_min_sync(act, o0), o1 = __reduce_min_sync(act, o1), o2 = __reduce_min_sync(act, o2);
        if ((threadIdx.x & 31) == __ffs(act) - 1) {
            atomicMin(mins + 3 * b, o0), atomicMin(mins + 3 * b + 1, o1), atomicMin(mins + 3 * b + 2, o2);
        }
    } else {
        atomicMin(mins + 3 * b, o0), atomicMin(mins + 3 * b + 1, o1), atomicMin(mins + 3 * b + 2, o2);
    }
}

// pass 2: discrete coordinate = trunc(aug + (-min + offset)); inside = inside the cut-out window [start, start + size)
// (fix_cut_out tests the UNMOVED coordinate against [0, size) and moves by `shift`: start = 0 for the test, -shift for the
// move; a drawn random cut-out passes start = its start positions and move = -start)
__global__ void __launch_bounds__(VOX_TB) k_vox_discretise(const float* __restrict__ aug, int P, const int* __restrict__ sample_ptr, int B,
                                                           const unsigned int* __restrict__ mins, const float* __restrict__ offset,
                                                           const int* __restrict__ window, int* __restrict__ disc,
                                                           int* __restrict__ inside, float* __restrict__ shift_out) {
    const int p = blockIdx.x * VOX_TB + threadIdx.x;
    if (p >= P) return;
    const int b = sample_of(sample_ptr, B, p);
    const int* w = window + 9 * b;      // start[3], size[3], move[3]
    bool in = true;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float sh = __fadd_rn(-f32_unordered(__ldg(mins + 3 * b + d)), __ldg(offset + 3 * b + d));
        if (p == __ldg(sample_ptr + b)) shift_out[3 * b + d] = sh;      // complete_shift before the cut-out (one writer per sample)
        const int v = (int)__fadd_rn(aug[3 * (int64_t)p + d], sh);      // .long(): truncation toward zero
        const int rel = v - __ldg(w + d);
        in = in && rel >= 0 && rel < __ldg(w + 3 + d);
        disc[3 * (int64_t)p + d] = v + __ldg(w + 6 + d);
    }
    inside[p] = in ? 1 : 0;
}

// pass 3 (after the exclusive scan of `inside`): collated coordinates [P', 4] int64 (x, y, z, sample) and the kept point of
// every output row
__global__ void __launch_bounds__(VOX_TB) k_vox_emit(const int* __restrict__ disc, const int* __restrict__ inside,
                                                     const int* __restrict__ rank, int P, const int* __restrict__ sample_ptr, int B,
                                                     int64_t* __restrict__ coords, int* __restrict__ kept, int* __restrict__ out_ptr) {
    const int p = blockIdx.x * VOX_TB + threadIdx.x;
    if (p <= B && p < B + 1) out_ptr[p] = rank[__ldg(sample_ptr + p)];      // batch_splits as offsets (sample_ptr[B] = P)
    if (p >= P || !inside[p]) return;
    const int b = sample_of(sample_ptr, B, p), r = rank[p];
    longlong2 lo, hi;
    lo.x = disc[3 * (int64_t)p], lo.y = disc[3 * (int64_t)p + 1], hi.x = disc[3 * (int64_t)p + 2], hi.y = b;
    reinterpret_cast<longlong2*>(coords)[2 * (int64_t)r] = lo;
    reinterpret_cast<longlong2*>(coords)[2 * (int64_t)r + 1] = hi;
    kept[r] = p;
}

// features of the kept points: [colours (+ common shift) | ones | normals @ rotation (+ common shift)], any part optional
__global__ void __launch_bounds__(VOX_TB) k_vox_features(const int* __restrict__ kept, int n, const int* __restrict__ out_ptr, int B,
                                                         const float* __restrict__ colors, const float* __restrict__ color_shift,
                                                         int use_ones, const float* __restrict__ normals,
                                                         const float* __restrict__ rotation, const float* __restrict__ normal_shift,
                                                         float* __restrict__ out, int C) {
    const int r = blockIdx.x * VOX_TB + threadIdx.x;
    if (r >= n) return;
    const int p = kept[r], b = sample_of(out_ptr, B, r);
    float* o = out + (int64_t)r * C;
    int c = 0;
    if (colors) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            float v = colors[3 * (int64_t)p + d];
            if (color_shift) v = __fadd_rn(v, __ldg(color_shift + 3 * b + d));
            o[c++] = v;
        }
    }
    if (use_ones) o[c++] = 1.f;
    if (normals) {
        float a0 = normals[3 * (int64_t)p], a1 = normals[3 * (int64_t)p + 1], a2 = normals[3 * (int64_t)p + 2];
        if (rotation) {
            float mm[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) mm[i] = __ldg(rotation + 9 * b + i);
            float r0, r1, r2;
            project(mm, a0, a1, a2, r0, r1, r2);
            a0 = r0, a1 = r1, a2 = r2;
        }
        if (normal_shift) {
            a0 = __fadd_rn(a0, __ldg(normal_shift + 3 * b)), a1 = __fadd_rn(a1, __ldg(normal_shift + 3 * b + 1));
            a2 = __fadd_rn(a2, __ldg(normal_shift + 3 * b + 2));
        }
        o[c] = a0, o[c + 1] = a1, o[c + 2] = a2;
    }
}

}  // namespace scn

using namespace scn;

extern "C" {

int64_t scn_voxelize_ws_bytes(int P, int B) {
    // aug [P,3] f32 | disc [P,3] i32 | inside [P] | rank [P+1] | scan tmp | mins [3B]
    const int64_t p = P > 0 ? P : 1;
    int64_t n = 12 * p + 12 * p + 4 * p + 4 * (p + 1) + 4 * scn_scan_tmp_elems(p) + 12 * (int64_t)(B > 0 ? B : 1);
    return (n + 1023) / 256 * 256 + 6 * 256;
}

int scn_voxelize(const float* points, int P, const int32_t* sample_ptr, int B, const float* proj, const float* offset,
                 const int32_t* window, void* ws, int64_t* coords, int32_t* kept, int32_t* out_ptr, float* shift_out,
                 scn_stream_t stream) {
    SCN_REQUIRE(P >= 0 && B >= 1 && B <= 4096, "voxelize: bad sizes P=%d B=%d", P, B);
    SCN_REQUIRE(sample_ptr && proj && offset && window && ws && out_ptr && shift_out && (P == 0 || (points && coords && kept)),
                "voxelize: null argument");
    SCN_REQUIRE((reinterpret_cast<uintptr_t>(coords) & 15) == 0 && (reinterpret_cast<uintptr_t>(ws) & 255) == 0,
                "voxelize: coords must be 16-byte and the workspace 256-byte aligned");
    cudaStream_t st = as_stream(stream);
    const int64_t p = P > 0 ? P : 1;
    auto up = [](int64_t v) { return (v + 255) / 256 * 256; };
    uint8_t* w = static_cast<uint8_t*>(ws);
    float* aug = reinterpret_cast<float*>(w);
    w += up(12 * p);
    int* disc = reinterpret_cast<int*>(w);
    w += up(12 * p);
    int* inside = reinterpret_cast<int*>(w);
    w += up(4 * p);
    int* rank = reinterpret_cast<int*>(w);
    w += up(4 * (p + 1));
    int* tmp = reinterpret_cast<int*>(w);
    w += up(4 * scn_scan_tmp_elems(p));
    unsigned int* mins = reinterpret_cast<unsigned int*>(w);
    k_vox_init<<<cdiv(3 * B, VOX_TB), VOX_TB, 0, st>>>(mins, 3 * B);
    if (P == 0) {
        cudaMemsetAsync(out_ptr, 0, sizeof(int32_t) * (B + 1), st);
        cudaMemsetAsync(shift_out, 0, sizeof(float) * 3 * B, st);
        return check_launch("voxelize(empty)");
    }
    const int grid = cdiv(P, VOX_TB);
    cudaMemsetAsync(shift_out, 0, sizeof(float) * 3 * B, st);      // samples without points keep a zero shift
    k_vox_project<<<grid, VOX_TB, 0, st>>>(points, P, sample_ptr, B, proj, aug, mins);
    k_vox_discretise<<<grid, VOX_TB, 0, st>>>(aug, P, sample_ptr, B, mins, offset, window, disc, inside, shift_out);
    int rc = scn_exclusive_scan(inside, rank, P, tmp, stream);
    if (rc) return rc;
    k_vox_emit<<<cdiv(P > B ? P : B + 1, VOX_TB), VOX_TB, 0, st>>>(disc, inside, rank, P, sample_ptr, B, coords, kept, out_ptr);
    return check_launch("voxelize");
}

int scn_voxelize_features(const int32_t* kept, int n, const int32_t* out_ptr, int B, const float* colors, const float* color_shift,
                          int use_ones, const float* normals, const float* rotation, const float* normal_shift, float* out, int C,
                          scn_stream_t stream) {
    SCN_REQUIRE(n >= 0 && B >= 1, "voxelize_features: bad sizes");
    SCN_REQUIRE(C == (colors ? 3 : 0) + (use_ones ? 1 : 0) + (normals ? 3 : 0) && C > 0, "voxelize_features: C = %d does not match the parts", C);
    if (n == 0) return SCN_OK;
    SCN_REQUIRE(kept && out_ptr && out, "voxelize_features: null argument");
    k_vox_features<<<cdiv(n, VOX_TB), VOX_TB, 0, as_stream(stream)>>>(kept, n, out_ptr, B, colors, color_shift, use_ones, normals, rotation,
                                                                     normal_shift, out, C);
    return check_launch("voxelize_features");
}

}  // extern "C"
