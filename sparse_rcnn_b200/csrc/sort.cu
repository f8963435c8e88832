// Spatially coherent row order: stable LSD radix sort of the points by (b, Morton(x, y, z)).
//
// Why (profiles/r2_a_row_order.md): with rows numbered by first appearance (SparseConvNet's level-0 rule; a scan has no
// spatial order) the 27 neighbours of the 128 rows of an output tile are 1270 different input rows; numbered along a
// Morton curve they are 276, so a tile's whole input fits in shared memory and is fetched ONCE per layer instead of
// once per kernel offset (conv_ts.cu).  SURVEY 8c: any batch-sorted row order is admissible, parity is judged after the
// canonical (b, x, y, z) sort.  The rulebook builder itself is unchanged: it numbers voxels by first appearance over the
// SORTED point list (= Morton order of the voxels), and a coarse level numbered by "first fine row" inherits the order
// because a Morton code's parent is its prefix.
//
// Kernels: k_morton_keys (pack), then per 8-bit digit k_radix_hist -> exclusive scan (grid.cu) -> k_radix_scatter with a
// stable in-block rank (warp-private digit counters + __match_any_sync), finally k_sort_finish (permutation + sorted keys).
// Integer, HBM-bound: 12 bytes read + 12 written per point and pass; only digits that can be non-zero are sorted.
#include "common.cuh"

namespace scn {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 8;
constexpr int RS_CHUNK = RS_THREADS * RS_ITEMS;      // keys per block

__device__ __forceinline__ uint64_t spread3(uint32_t v) {
    uint64_t x = v & 0xFFFFu;
    x = (x | (x << 32)) & 0x1f00000000ffffull;
    x = (x | (x << 16)) & 0x1f0000ff0000ffull;
    x = (x | (x << 8)) & 0x100f00f00f00f00full;
    x = (x | (x << 4)) & 0x10c30c30c30c30c3ull;
    x = (x | (x << 2)) & 0x1249249249249249ull;
    return x;
}

// sort key: b << 48 | interleave(x, y, z) with z in the lowest bit (the in-box offset index of a 2^3 convolution)
__global__ void k_morton_keys(const uint64_t* __restrict__ keys, int P, uint64_t* __restrict__ mkeys) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
        const uint64_t k = keys[i];
        mkeys[i] = ((uint64_t)key_b(k) << 48) | (spread3(key_x(k)) << 2) | (spread3(key_y(k)) << 1) | spread3(key_z(k));
    }
}

__global__ void __launch_bounds__(RS_THREADS) k_radix_hist(const uint64_t* __restrict__ keys, int n, int shift, int nblocks,
                                                           int32_t* __restrict__ hist) {
    __shared__ int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int base = blockIdx.x * RS_CHUNK;
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        const int i = base + j * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(int)((keys[i] >> shift) & 255u)], 1);
    }
    __syncthreads();
    hist[threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];      // digit-major: one scan gives every (digit, block) base
}

// Stable scatter: warp w of a block owns 256 consecutive keys and walks them in 8 rounds of 32 consecutive keys; the rank of a
// key among the keys of its digit inside the block = (count in the earlier warps) + (count in this warp's earlier rounds)
// + (lower lanes of this round with the same digit).
__global__ void __launch_bounds__(RS_THREADS) k_radix_scatter(const uint64_t* __restrict__ keys_in,
                                                              const int32_t* __restrict__ vals_in, int n, int shift,
                                                              int nblocks, const int32_t* __restrict__ base,
                                                              uint64_t* __restrict__ keys_out, int32_t* __restrict__ vals_out) {
    __shared__ int cnt[RS_WARPS][256];
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&cnt[0][0])[i] = 0;
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const int wbase = blockIdx.x * RS_CHUNK + w * (32 * RS_ITEMS);
    uint64_t k[RS_ITEMS];
    int rank[RS_ITEMS];
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        const int i = wbase + j * 32 + lane;
        const bool valid = i < n;
        k[j] = valid ? keys_in[i] : 0ull;
        const int d = valid ? (int)((k[j] >> shift) & 255u) : 256 + lane;      // invalid lanes match nobody
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const int prev = valid ? cnt[w][d] : 0;
        __syncwarp();
        const int r = __popc(peers & lt);
        if (valid && r == 0) cnt[w][d] = prev + __popc(peers);
        __syncwarp();
        rank[j] = prev + r;
    }
    __syncthreads();
    {
        const int t = threadIdx.x;      // digit t: exclusive prefix over the block's warps on top of the global base
        int run = base[t * nblocks + blockIdx.x];
#pragma unroll
        for (int w2 = 0; w2 < RS_WARPS; ++w2) {
            const int c = cnt[w2][t];
            cnt[w2][t] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        const int i = wbase + j * 32 + lane;
        if (i < n) {
            const int d = (int)((k[j] >> shift) & 255u);
            const int pos = cnt[w][d] + rank[j];
            keys_out[pos] = k[j];
            vals_out[pos] = vals_in ? vals_in[i] : i;
        }
    }
}

__global__ void k_sort_finish(const int32_t* __restrict__ vals, const uint64_t* __restrict__ keys, int P,
                              int32_t* __restrict__ perm, uint64_t* __restrict__ sorted_keys) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
        const int v = vals[i];
        perm[i] = v;
        sorted_keys[i] = keys[v];
    }
}

__global__ void k_iota_copy(const uint64_t* __restrict__ keys, int P, int32_t* __restrict__ perm, uint64_t* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) perm[i] = i, out[i] = keys[i];
}

__global__ void k_scatter_i32(const int32_t* __restrict__ src, const int32_t* __restrict__ perm, int n, int32_t* __restrict__ dst) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[perm[i]] = src[i];
}

static inline int64_t align256(int64_t v) { return (v + 255) / 256 * 256; }
struct SortWs {
    uint64_t *ka, *kb;
    int32_t *va, *vb, *hist, *scanned, *scan_tmp;
    int64_t bytes;
    int nblocks;
};
static SortWs sort_ws(void* ws, int P) {
    SortWs s;
    s.nblocks = (P + RS_CHUNK - 1) / RS_CHUNK;
    if (s.nblocks < 1) s.nblocks = 1;
    const int64_t nh = 256ll * s.nblocks;
    char* p = reinterpret_cast<char*>(ws);
    int64_t off = 0;
    auto take = [&](int64_t bytes) {
        char* q = p ? p + off : nullptr;
        off += align256(bytes);
        return q;
    };
    s.ka = reinterpret_cast<uint64_t*>(take(8ll * P));
    s.kb = reinterpret_cast<uint64_t*>(take(8ll * P));
    s.va = reinterpret_cast<int32_t*>(take(4ll * P));
    s.vb = reinterpret_cast<int32_t*>(take(4ll * P));
    s.hist = reinterpret_cast<int32_t*>(take(4 * nh));
    s.scanned = reinterpret_cast<int32_t*>(take(4 * (nh + 1)));
    s.scan_tmp = reinterpret_cast<int32_t*>(take(4 * scn_scan_tmp_elems(nh)));
    s.bytes = off;
    return s;
}

}  // namespace scn

using namespace scn;

extern "C" {

int64_t scn_morton_order_ws_bytes(int P) { return sort_ws(nullptr, P < 0 ? 0 : P).bytes; }

int scn_morton_order(const uint64_t* keys, int P, int coord_bits, int batch_bits, int32_t* perm, uint64_t* sorted_keys,
                     void* ws, scn_stream_t stream) {
    SCN_REQUIRE(P >= 0 && coord_bits >= 0 && coord_bits <= 16 && batch_bits >= 0 && batch_bits <= 16,
                "morton_order: bad arguments (P=%d coord_bits=%d batch_bits=%d)", P, coord_bits, batch_bits);
    if (P == 0) return SCN_OK;
    SCN_REQUIRE(ws && perm && sorted_keys, "morton_order: null buffer");
    cudaStream_t st = as_stream(stream);
    SortWs s = sort_ws(ws, P);
    // digits that can be non-zero: the interleaved coordinate bits [0, 3 * coord_bits) and the batch bits [48, 48 + batch_bits)
    int shifts[8], np = 0;
    for (int b = 0; b < 3 * coord_bits; b += 8) shifts[np++] = b;
    for (int b = 0; b < batch_bits; b += 8) shifts[np++] = 48 + b;
    if (np == 0) {
        k_iota_copy<<<grid_for(P, 256), 256, 0, st>>>(keys, P, perm, sorted_keys);
        return check_launch("morton_order(iota)");
    }
    k_morton_keys<<<grid_for(P, 256), 256, 0, st>>>(keys, P, s.ka);
    int rc = check_launch("morton_keys");
    if (rc) return rc;
    const uint64_t* kin = s.ka;
    uint64_t* kout = s.kb;
    const int32_t* vin = nullptr;
    int32_t* vout = s.va;
    for (int p = 0; p < np; ++p) {
        k_radix_hist<<<s.nblocks, RS_THREADS, 0, st>>>(kin, P, shifts[p], s.nblocks, s.hist);
        if ((rc = check_launch("radix_hist"))) return rc;
        if ((rc = scn_exclusive_scan(s.hist, s.scanned, 256ll * s.nblocks, s.scan_tmp, stream))) return rc;
        k_radix_scatter<<<s.nblocks, RS_THREADS, 0, st>>>(kin, vin, P, shifts[p], s.nblocks, s.scanned, kout, vout);
        if ((rc = check_launch("radix_scatter"))) return rc;
        vin = vout;
        vout = (vout == s.va) ? s.vb : s.va;
        const uint64_t* t = kout;
        kout = const_cast<uint64_t*>(kin);
        kin = t;
    }
    k_sort_finish<<<grid_for(P, 256), 256, 0, st>>>(vin, keys, P, perm, sorted_keys);
    return check_launch("sort_finish");
}

int scn_scatter_i32(const int32_t* src, const int32_t* perm, int n, int32_t* dst, scn_stream_t stream) {
    if (n <= 0) return SCN_OK;
    k_scatter_i32<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(src, perm, n, dst);
    return check_launch("scatter_i32");
}

}  // extern "C"
