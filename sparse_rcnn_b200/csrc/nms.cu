// Proposal selection: greedy 3-D non-maximum suppression on the device.
//
// Reference: ProposalSelector.forward (ndsis/modules/proposal_selector.py:60-89) sorts the proposals of every sample by
// score, copies the order to the host, and calls non_maximum_supression (ndsis/utils/bbox.py:713-759): an n x n IoU matrix
// followed by a Python loop of n iterations x 2 tiny kernels.  Here: one kernel builds the strict-upper-triangle
// "IoU > threshold" bit matrix, one warp per sample walks it from shared memory.  IoU arithmetic follows
// bbox_overlap_unsqueezed_area_start_end (utils/bbox.py:235-240) operation by operation in fp32 so that the threshold
// decisions are identical.
#include "common.cuh"

namespace scn {

// word w of row j: bit b set iff box i = 32 w + b comes AFTER j (i > j) and IoU(i, j) > thresh
__global__ void k_nms_mask(const float* __restrict__ boxes, int n, int W, float thresh, uint32_t* __restrict__ mask) {
    const int b = blockIdx.y;
    const float* bx = boxes + (int64_t)b * n * 6;
    uint32_t* mk = mask + (int64_t)b * n * W;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < (int64_t)n * W; t += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(t / W), w = (int)(t % W);
        uint32_t word = 0;
        if (w * 32 + 31 > j) {
            const float js0 = bx[j * 6 + 0], js1 = bx[j * 6 + 1], js2 = bx[j * 6 + 2];
            const float je0 = bx[j * 6 + 3], je1 = bx[j * 6 + 4], je2 = bx[j * 6 + 5];
            const float ja = __fmul_rn(__fmul_rn(je0 - js0, je1 - js1), je2 - js2);      // size.prod(): ((x*y)*z)
            for (int k = 0; k < 32; ++k) {
                const int i = w * 32 + k;
                if (i <= j || i >= n) continue;
                const float is0 = bx[i * 6 + 0], is1 = bx[i * 6 + 1], is2 = bx[i * 6 + 2];
                const float ie0 = bx[i * 6 + 3], ie1 = bx[i * 6 + 4], ie2 = bx[i * 6 + 5];
                const float ia = __fmul_rn(__fmul_rn(ie0 - is0, ie1 - is1), ie2 - is2);
                const float d0 = fmaxf(fminf(ie0, je0) - fmaxf(is0, js0), 0.f);
                const float d1 = fmaxf(fminf(ie1, je1) - fmaxf(is1, js1), 0.f);
                const float d2 = fmaxf(fminf(ie2, je2) - fmaxf(is2, js2), 0.f);
                const float inter = __fmul_rn(__fmul_rn(d0, d1), d2);
                const float uni = __fsub_rn(__fadd_rn(ia, ja), inter);      // area_a + area_b - intersection
                const float iou = __fdiv_rn(inter, uni);                    // NaN for degenerate pairs: never > thresh
                if (iou > thresh) word |= 1u << k;
            }
        }
        mk[(int64_t)j * W + w] = word;
    }
}

// one CTA per sample: stage the bit matrix in shared memory (when it fits), then warp 0 walks it in score order
__global__ void k_nms_scan(const uint32_t* __restrict__ mask, int n, int W, int max_keep, int use_smem,
                           uint8_t* __restrict__ keep, int32_t* __restrict__ keep_idx, int32_t* __restrict__ counts) {
    extern __shared__ uint32_t sm[];
    const int b = blockIdx.x;
    const uint32_t* mk = mask + (int64_t)b * n * W;
    if (use_smem) {
        for (int64_t t = threadIdx.x; t < (int64_t)n * W; t += blockDim.x) sm[t] = mk[t];
        __syncthreads();
        mk = sm;
    }
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    // lane L owns words L, L+32, L+64, L+96 (n <= 4096)
    uint32_t alive[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const int w = lane + 32 * s;
        uint32_t a = 0;
        if (w < W) {
            const int left = n - w * 32;
            a = left >= 32 ? 0xffffffffu : (left > 0 ? (1u << left) - 1u : 0u);
        }
        alive[s] = a;
    }
    for (int j = 0; j < n; ++j) {
        const int w = j >> 5, s = w >> 5;
        uint32_t word = s == 0 ? alive[0] : (s == 1 ? alive[1] : (s == 2 ? alive[2] : alive[3]));
        word = __shfl_sync(0xffffffffu, word, w & 31);
        if ((word >> (j & 31)) & 1u) {      // box j survived everything before it: it suppresses its overlaps
            const uint32_t* row = mk + (int64_t)j * W;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (lane + 32 * q < W) alive[q] &= ~row[lane + 32 * q];
        }
    }
    // keep flags + the first max_keep survivors in score order
    int base = 0;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const int w = lane + 32 * s;
        const uint32_t a = alive[s];
        const int c = __popc(a);
        int incl = c;
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        int pos = base + incl - c;
        if (w < W) {
            for (int k = 0; k < 32; ++k) {
                const int i = w * 32 + k;
                if (i >= n) break;
                const bool on = (a >> k) & 1u;
                keep[(int64_t)b * n + i] = on ? 1 : 0;
                if (on) {
                    if (pos < max_keep) keep_idx[(int64_t)b * max_keep + pos] = i;
                    ++pos;
                }
            }
        }
        base += total;
    }
    if (lane == 0) counts[b] = base < max_keep ? base : max_keep;
}

}  // namespace scn

using namespace scn;

extern "C" {

int64_t scn_nms3d_workspace_bytes(int B, int n) {
    if (B <= 0 || n <= 0) return 0;
    return (int64_t)B * n * ((n + 31) / 32) * 4;
}

int scn_nms3d(const float* boxes, int B, int n, float thresh, int max_keep, void* workspace, uint8_t* keep,
              int32_t* keep_idx, int32_t* counts, scn_stream_t stream) {
    SCN_REQUIRE(B >= 0 && n >= 0 && max_keep >= 0, "nms3d: bad shape");
    SCN_REQUIRE(n <= 4096, "nms3d: at most 4096 proposals per sample (got %d)", n);
    if (B == 0) return SCN_OK;
    cudaStream_t st = as_stream(stream);
    if (n == 0) {
        cudaMemsetAsync(counts, 0, sizeof(int32_t) * B, st);
        return check_launch("nms3d(memset)");
    }
    const int W = (n + 31) / 32;
    uint32_t* mask = reinterpret_cast<uint32_t*>(workspace);
    dim3 grid(grid_for((int64_t)n * W, 256), B);
    k_nms_mask<<<grid, 256, 0, st>>>(boxes, n, W, thresh, mask);
    int rc = check_launch("nms3d_mask");
    if (rc) return rc;
    const size_t bytes = (size_t)n * W * 4;
    const int use_smem = bytes <= 200 * 1024;
    if (use_smem) {
        cudaError_t e = (cudaError_t)scn::ensure_dynamic_smem(reinterpret_cast<const void*>(k_nms_scan), (int)bytes);
        if (e != cudaSuccess) {
            cudaGetLastError();
            scn::set_error("nms3d: cudaFuncSetAttribute(%zu bytes): %s", bytes, cudaGetErrorString(e));
            return SCN_ERR_CUDA;
        }
    }
    k_nms_scan<<<B, 256, use_smem ? bytes : 0, st>>>(mask, n, W, max_keep, use_smem, keep, keep_idx, counts);
    return check_launch("nms3d_scan");
}

}  // extern "C"
