// Rulebook builder: GPU voxel hash table, first-appearance row numbering, input-rule CSR,
// submanifold neighbour maps and strided (conv / deconv / pooling) maps.
//
// Replaces SparseConvNet's host-side Metadata<d> (hash maps + std::vector rulebooks that are
// re-uploaded per convolution call).  Everything stays resident in HBM.  All kernels are
// HBM/L2-latency bound integer work: one thread per point / row, coalesced streaming of the
// key arrays, random probes into a table sized at load factor <= 0.5 that lives in the 126 MB L2.
#include <stdarg.h>
#include <atomic>
#include <mutex>
#include <stdlib.h>
#include "common.cuh"

namespace scn {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
static std::atomic<int64_t> g_launches{0};      // launches come from two host threads when geometry is prefetched
int check_launch(const char* what) {
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return SCN_ERR_CUDA;
    }
    return SCN_OK;
}
bool pdl_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("SCN_PDL");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on == 1;
}

PdlMask& pdl_mask() {
    static thread_local PdlMask m;
    return m;
}

int sm_count() {
    static int n = -1;
    if (n < 0) {
        int dev = 0;
        cudaDeviceProp p;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
            cudaGetLastError();
            return 148;
        }
        n = p.multiProcessorCount;
    }
    return n;
}

int ensure_dynamic_smem(const void* kernel, int bytes) {
    static const void* fns[32];
    static int vals[32];
    static int n = 0;
    static std::mutex mu;      // launches may come from several host threads (one stream each)
    std::lock_guard<std::mutex> lock(mu);
    int i = 0;
    for (; i < n; ++i)
        if (fns[i] == kernel) break;
    if (i < n && vals[i] >= bytes) return (int)cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return (int)e;
    if (i == n && n < 32) fns[n++] = kernel;
    if (i < 32) vals[i] = bytes;
    return (int)cudaSuccess;
}

// ------------------------------------------------------------------------------ kernels
__global__ void k_pack_coords(const int64_t* __restrict__ coords, int P, int ncol, uint64_t* __restrict__ keys,
                              int* __restrict__ err) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
        const int64_t* c = coords + (int64_t)i * ncol;
        int64_t x = c[0], y = c[1], z = c[2], b = ncol > 3 ? c[3] : 0;
        if ((x | y | z | b) < 0 || x > 65534 || y > 65534 || z > 65534 || b > 65534) {
            *err = 1;
            x = y = z = b = 0;
        }
        keys[i] = make_key((uint32_t)x, (uint32_t)y, (uint32_t)z, (uint32_t)b);
    }
}

__global__ void k_unpack_keys(const uint64_t* __restrict__ keys, int n, int64_t* __restrict__ coords) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint64_t k = keys[i];
        longlong4 v = make_longlong4(key_x(k), key_y(k), key_z(k), key_b(k));
        reinterpret_cast<longlong4*>(coords)[i] = v;
    }
}

__global__ void k_hash_clear(uint64_t* __restrict__ tk, int32_t* __restrict__ tv, uint32_t cap) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += gridDim.x * blockDim.x) {
        tk[i] = SCN_EMPTY_KEY;
        tv[i] = 0x7FFFFFFF;
    }
}

__global__ void k_hash_insert_first(const uint64_t* __restrict__ keys, int P, uint64_t* tk, int32_t* tv,
                                    uint32_t mask) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
        uint32_t s = hash_insert_slot(tk, mask, keys[i]);
        atomicMin(tv + s, i);
    }
}

__global__ void k_hash_first_flags(const uint64_t* __restrict__ keys, int P, const uint64_t* __restrict__ tk,
                                   const int32_t* __restrict__ tv, uint32_t mask, int32_t* __restrict__ first) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
        first[i] = (hash_lookup(tk, tv, mask, keys[i]) == i) ? 1 : 0;
    }
}

__global__ void k_hash_assign_rows(const uint64_t* __restrict__ keys, int P, const uint64_t* __restrict__ tk,
                                   const int32_t* __restrict__ tv, uint32_t mask, const int32_t* __restrict__ rank,
                                   int32_t* __restrict__ point_row, uint64_t* __restrict__ row_keys) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
        uint64_t k = keys[i];
        int first = hash_lookup(tk, tv, mask, k);
        int row = rank[first];
        point_row[i] = row;
        if (first == i) row_keys[row] = k;
    }
}

__global__ void k_hash_finalize(const uint64_t* __restrict__ tk, int32_t* __restrict__ tv, uint32_t cap,
                                const int32_t* __restrict__ rank) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += gridDim.x * blockDim.x) {
        if (tk[i] != SCN_EMPTY_KEY) tv[i] = rank[tv[i]];
    }
}

__global__ void k_hash_lookup(const uint64_t* __restrict__ keys, int n, const uint64_t* __restrict__ tk,
                              const int32_t* __restrict__ tv, uint32_t mask, int32_t* __restrict__ rows) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        rows[i] = hash_lookup(tk, tv, mask, keys[i]);
    }
}

// ------------------------------------------------------------------------------ scan
// Three-phase exclusive scan (block reduce -> recursive scan of block sums -> block scan).
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = SCN_SCAN_BLOCK / SCAN_THREADS;  // 8

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// exclusive scan of one value per thread across the block; returns exclusive prefix, total via *total
__device__ __forceinline__ int block_excl_scan(int v, int* total) {
    __shared__ int warp_sums[32];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = warp_incl_scan(v, lane);
    if (lane == 31) warp_sums[w] = inc;
    __syncthreads();
    if (w == 0) {
        int nw = blockDim.x >> 5;
        int s = lane < nw ? warp_sums[lane] : 0;
        int si = warp_incl_scan(s, lane);
        warp_sums[lane] = si - s;          // exclusive warp offsets
    }
    __syncthreads();
    int off = warp_sums[w];
    // total = offset of last warp + its sum
    if (total) {
        __shared__ int tot;
        if (threadIdx.x == blockDim.x - 1) tot = off + inc;
        __syncthreads();
        *total = tot;
    }
    return off + inc - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_block_sums(const int32_t* __restrict__ in, int64_t n,
                                                                  int32_t* __restrict__ sums) {
    int64_t base = (int64_t)blockIdx.x * SCN_SCAN_BLOCK;
    int s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        int64_t i = base + (int64_t)j * SCAN_THREADS + threadIdx.x;
        if (i < n) s += in[i];
    }
    int total;
    block_excl_scan(s, &total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_block(const int32_t* __restrict__ in, int64_t n,
                                                             const int32_t* __restrict__ block_off,
                                                             int32_t* __restrict__ out, int write_total) {
    // each thread owns SCAN_ITEMS consecutive items (blocked arrangement)
    int64_t base = (int64_t)blockIdx.x * SCN_SCAN_BLOCK + (int64_t)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        v[j] = (base + j < n) ? in[base + j] : 0;
        s += v[j];
    }
    int total;
    int ex = block_excl_scan(s, &total);
    int off = block_off ? block_off[blockIdx.x] : 0;
    int run = off + ex;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        if (base + j < n) out[base + j] = run;
        run += v[j];
    }
    if (write_total && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out[n] = off + total;
}

static int scan_rec(const int32_t* in, int32_t* out, int64_t n, int32_t* tmp, int write_total, cudaStream_t st) {
    int64_t nb = (n + SCN_SCAN_BLOCK - 1) / SCN_SCAN_BLOCK;
    if (nb <= 1) {
        k_scan_block<<<1, SCAN_THREADS, 0, st>>>(in, n, nullptr, out, write_total);
        return check_launch("scan_block");
    }
    int32_t* sums = tmp;             // nb entries (+1 for the total written by the recursion)
    int32_t* rest = tmp + nb + 1;
    k_scan_block_sums<<<(int)nb, SCAN_THREADS, 0, st>>>(in, n, sums);
    int rc = check_launch("scan_block_sums");
    if (rc) return rc;
    rc = scan_rec(sums, sums, nb, rest, 0, st);  // in-place exclusive scan of block sums
    if (rc) return rc;
    k_scan_block<<<(int)nb, SCAN_THREADS, 0, st>>>(in, n, sums, out, write_total);
    return check_launch("scan_block");
}

// ------------------------------------------------------------------------------ input rule CSR
__global__ void k_rule_count(const int32_t* __restrict__ point_row, int P, int32_t* row_cnt) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x)
        atomicAdd(row_cnt + point_row[i], 1);
}
__global__ void k_rule_fill(const int32_t* __restrict__ point_row, int P, const int32_t* __restrict__ row_ptr,
                            int32_t* cursor, int32_t* __restrict__ row_pts) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
        int r = point_row[i];
        int pos = atomicAdd(cursor + r, 1);
        row_pts[row_ptr[r] + pos] = i;
    }
}
// segments are tiny (points per voxel): one thread insertion-sorts one row => deterministic order
__global__ void k_rule_sort(const int32_t* __restrict__ row_ptr, int N, int32_t* __restrict__ row_pts) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < N; r += gridDim.x * blockDim.x) {
        int b = row_ptr[r], e = row_ptr[r + 1];
        for (int i = b + 1; i < e; ++i) {
            int v = row_pts[i], j = i - 1;
            while (j >= b && row_pts[j] > v) {
                row_pts[j + 1] = row_pts[j];
                --j;
            }
            row_pts[j + 1] = v;
        }
    }
}

// ------------------------------------------------------------------------------ neighbour maps
// Neighbour map of an odd filter box, offsets enumerated with the last dimension fastest (SparseConvNet order).
// The map is symmetric: map[o][r] == q  <=>  map[K-1-o][q] == r, so only the first K/2 offsets are probed -- ONE thread per
// (offset, row) pair, 13 N independent probes for a 3^3 filter instead of 27 dependent ones per thread.  Threads of a warp
// share the offset and hold consecutive rows: the key read and the write of map[o][r] are coalesced, the mirrored write
// map[K-1-o][q] goes to rows near r (Morton order keeps neighbours close).  Entries of the upper half without a neighbour keep
// the -1 that scn_subm_map fills in first.  Round 1 walked all 27 probes of a row serially in one thread: 21-35 us per level
// whatever its size (a chain of L2 misses); measured after the rewrite in profiles/r2_c_kernel_table.md.
__global__ void k_subm_map(const uint64_t* __restrict__ row_keys, int N, const uint64_t* __restrict__ tk,
                           const int32_t* __restrict__ tv, uint32_t mask, int fx, int fy, int fz,
                           int32_t* __restrict__ map) {
    const int K = fx * fy * fz, half = K / 2;
    const int64_t total = (int64_t)half * N;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int o = (int)(p / N), r = (int)(p - (int64_t)o * N);
        const uint64_t k = __ldg(row_keys + r);
        const int dz = o % fz - fz / 2, dy = (o / fz) % fy - fy / 2, dx = o / (fz * fy) - fx / 2;
        const int qx = key_x(k) + dx, qy = key_y(k) + dy, qz = key_z(k) + dz;
        int v = -1;
        if (qx >= 0 && qy >= 0 && qz >= 0 && qx < 65535 && qy < 65535 && qz < 65535)
            v = hash_lookup(tk, tv, mask, make_key(qx, qy, qz, key_b(k)));
        map[(int64_t)o * N + r] = v;
        if (v >= 0) map[(int64_t)(K - 1 - o) * N + v] = r;
        if (o == 0) map[(int64_t)half * N + r] = r;      // the centre offset
    }
    if (half == 0)      // 1^3 filter: the identity
        for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < N; r += gridDim.x * blockDim.x) map[r] = r;
}

__global__ void k_stride_keys(const uint64_t* __restrict__ row_keys, int N, int sx, int sy, int sz,
                              uint64_t* __restrict__ parent_keys, int32_t* __restrict__ offs) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < N; r += gridDim.x * blockDim.x) {
        uint64_t k = row_keys[r];
        int x = key_x(k), y = key_y(k), z = key_z(k), b = key_b(k);
        int px = x / sx, py = y / sy, pz = z / sz;
        parent_keys[r] = make_key(px, py, pz, b);
        offs[r] = ((x - px * sx) * sy + (y - py * sy)) * sz + (z - pz * sz);
    }
}

__global__ void k_strided_maps(const int32_t* __restrict__ parent_row, const int32_t* __restrict__ offs, int n_in,
                               int n_out, int K, int32_t* __restrict__ cmap, int32_t* __restrict__ dmap) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_in; i += gridDim.x * blockDim.x) {
        int p = parent_row[i], o = offs[i];
        cmap[(int64_t)o * n_out + p] = i;
        for (int q = 0; q < K; ++q) dmap[(int64_t)q * n_in + i] = (q == o) ? p : -1;
    }
}

// seg_ptr[b] = first row with batch >= b.  rows are batch-sorted.
__global__ void k_batch_offsets(const uint64_t* __restrict__ row_keys, int N, int n_seg, int32_t* __restrict__ seg_ptr) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r <= N; r += gridDim.x * blockDim.x) {
        int b_prev = r == 0 ? -1 : key_b(row_keys[r - 1]);
        int b_cur = r == N ? n_seg : key_b(row_keys[r]);
        if (b_cur > n_seg) b_cur = n_seg;
        for (int b = b_prev + 1; b <= b_cur; ++b) seg_ptr[b] = r;
    }
}

}  // namespace scn

using namespace scn;
static bool is_pow2(uint32_t v) { return v && !(v & (v - 1)); }
constexpr int TB = 256;

extern "C" {

const char* scn_last_error(void) { return g_err; }
int scn_version(void) { return 100; }
int64_t scn_launch_count(void) { return g_launches.load(); }
int scn_device_sm_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return 0;
    }
    return sm_count();
}
int scn_device_is_sm100(void) {
    int n = 0, dev = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return 0;
    }
    cudaDeviceProp p;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 0;
    return p.major == 10 ? 1 : 0;
}

int scn_pack_coords(const int64_t* coords, int P, int ncol, uint64_t* keys, int* err_flag, scn_stream_t stream) {
    SCN_REQUIRE(ncol == 3 || ncol == 4, "pack_coords: ncol must be 3 or 4 (got %d)", ncol);
    if (P <= 0) return SCN_OK;
    k_pack_coords<<<grid_for(P, TB), TB, 0, as_stream(stream)>>>(coords, P, ncol, keys, err_flag);
    return check_launch("pack_coords");
}
int scn_unpack_keys(const uint64_t* keys, int n, int64_t* coords, scn_stream_t stream) {
    if (n <= 0) return SCN_OK;
    SCN_REQUIRE((reinterpret_cast<uintptr_t>(coords) & 31) == 0, "unpack_keys: coords must be 32-byte aligned");
    k_unpack_keys<<<grid_for(n, TB), TB, 0, as_stream(stream)>>>(keys, n, coords);
    return check_launch("unpack_keys");
}
int scn_hash_clear(uint64_t* tk, int32_t* tv, uint32_t cap, scn_stream_t stream) {
    SCN_REQUIRE(is_pow2(cap), "hash: capacity must be a power of two (got %u)", cap);
    k_hash_clear<<<grid_for(cap, TB), TB, 0, as_stream(stream)>>>(tk, tv, cap);
    return check_launch("hash_clear");
}
int scn_hash_insert_first(const uint64_t* keys, int P, uint64_t* tk, int32_t* tv, uint32_t cap, scn_stream_t stream) {
    SCN_REQUIRE(is_pow2(cap), "hash: capacity must be a power of two (got %u)", cap);
    SCN_REQUIRE((uint64_t)P * 2 <= (uint64_t)cap || P < 16, "hash: load factor > 0.5 (P=%d cap=%u)", P, cap);
    if (P <= 0) return SCN_OK;
    k_hash_insert_first<<<grid_for(P, TB), TB, 0, as_stream(stream)>>>(keys, P, tk, tv, cap - 1);
    return check_launch("hash_insert_first");
}
int scn_hash_first_flags(const uint64_t* keys, int P, const uint64_t* tk, const int32_t* tv, uint32_t cap,
                         int32_t* first, scn_stream_t stream) {
    SCN_REQUIRE(is_pow2(cap), "hash: capacity must be a power of two (got %u)", cap);
    if (P <= 0) return SCN_OK;
    k_hash_first_flags<<<grid_for(P, TB), TB, 0, as_stream(stream)>>>(keys, P, tk, tv, cap - 1, first);
    return check_launch("hash_first_flags");
}
int64_t scn_scan_tmp_elems(int64_t n) {
    int64_t total = 0;
    while (n > SCN_SCAN_BLOCK) {
        n = (n + SCN_SCAN_BLOCK - 1) / SCN_SCAN_BLOCK;
        total += n + 1;
    }
    return total + 2;
}
int scn_exclusive_scan(const int32_t* in, int32_t* out, int64_t n, int32_t* tmp, scn_stream_t stream) {
    SCN_REQUIRE(n >= 0 && n < (int64_t)1 << 31, "scan: n out of range");
    return scan_rec(in, out, n, tmp, 1, as_stream(stream));
}
int scn_hash_assign_rows(const uint64_t* keys, int P, const uint64_t* tk, const int32_t* tv, uint32_t cap,
                         const int32_t* rank, int32_t* point_row, uint64_t* row_keys, scn_stream_t stream) {
    SCN_REQUIRE(is_pow2(cap), "hash: capacity must be a power of two (got %u)", cap);
    if (P <= 0) return SCN_OK;
    k_hash_assign_rows<<<grid_for(P, TB), TB, 0, as_stream(stream)>>>(keys, P, tk, tv, cap - 1, rank, point_row, row_keys);
    return check_launch("hash_assign_rows");
}
// The two halves of building one level around its single host round trip (the active-row count), each as ONE call: the
// training step is host bound and a level otherwise costs six Python -> C round trips.
int scn_level_count(const uint64_t* keys, int P, uint64_t* tk, int32_t* tv, uint32_t cap, int32_t* first, int32_t* rank,
                    int32_t* scan_tmp, scn_stream_t stream) {
    int rc = scn_hash_clear(tk, tv, cap, stream);
    if (rc) return rc;
    if ((rc = scn_hash_insert_first(keys, P, tk, tv, cap, stream))) return rc;
    if ((rc = scn_hash_first_flags(keys, P, tk, tv, cap, first, stream))) return rc;
    return scn_exclusive_scan(first, rank, P, scan_tmp, stream);
}
int scn_level_finish(const uint64_t* keys, int P, uint64_t* tk, int32_t* tv, uint32_t cap, const int32_t* rank,
                     int32_t* point_row, uint64_t* row_keys, scn_stream_t stream) {
    int rc = scn_hash_assign_rows(keys, P, tk, tv, cap, rank, point_row, row_keys, stream);
    if (rc) return rc;
    return scn_hash_finalize(tk, tv, cap, rank, stream);
}
int scn_hash_finalize(const uint64_t* tk, int32_t* tv, uint32_t cap, const int32_t* rank, scn_stream_t stream) {
    k_hash_finalize<<<grid_for(cap, TB), TB, 0, as_stream(stream)>>>(tk, tv, cap, rank);
    return check_launch("hash_finalize");
}
int scn_hash_lookup(const uint64_t* keys, int n, const uint64_t* tk, const int32_t* tv, uint32_t cap, int32_t* rows,
                    scn_stream_t stream) {
    SCN_REQUIRE(is_pow2(cap), "hash: capacity must be a power of two (got %u)", cap);
    if (n <= 0) return SCN_OK;
    k_hash_lookup<<<grid_for(n, TB), TB, 0, as_stream(stream)>>>(keys, n, tk, tv, cap - 1, rows);
    return check_launch("hash_lookup");
}
int scn_rule_count(const int32_t* point_row, int P, int32_t* row_cnt, scn_stream_t stream) {
    if (P <= 0) return SCN_OK;
    k_rule_count<<<grid_for(P, TB), TB, 0, as_stream(stream)>>>(point_row, P, row_cnt);
    return check_launch("rule_count");
}
int scn_rule_fill(const int32_t* point_row, int P, const int32_t* row_ptr, int32_t* cursor, int32_t* row_pts,
                  scn_stream_t stream) {
    if (P <= 0) return SCN_OK;
    k_rule_fill<<<grid_for(P, TB), TB, 0, as_stream(stream)>>>(point_row, P, row_ptr, cursor, row_pts);
    return check_launch("rule_fill");
}
int scn_rule_sort(const int32_t* row_ptr, int N, int32_t* row_pts, scn_stream_t stream) {
    if (N <= 0) return SCN_OK;
    k_rule_sort<<<grid_for(N, TB), TB, 0, as_stream(stream)>>>(row_ptr, N, row_pts);
    return check_launch("rule_sort");
}
// input rule CSR in one call: count -> exclusive scan -> fill -> per-row sort (cursor: N int32 of scratch)
int scn_input_rule(const int32_t* point_row, int P, int N, int32_t* cursor, int32_t* row_ptr, int32_t* row_pts,
                   int32_t* scan_tmp, scn_stream_t stream) {
    SCN_REQUIRE(P >= 0 && N >= 0, "input_rule: bad shape");
    cudaStream_t st = as_stream(stream);
    if (N > 0) cudaMemsetAsync(cursor, 0, sizeof(int32_t) * N, st);
    int rc = scn_rule_count(point_row, P, cursor, stream);
    if (rc) return rc;
    if ((rc = scn_exclusive_scan(cursor, row_ptr, N, scan_tmp, stream))) return rc;
    if (N > 0) cudaMemsetAsync(cursor, 0, sizeof(int32_t) * N, st);
    if ((rc = scn_rule_fill(point_row, P, row_ptr, cursor, row_pts, stream))) return rc;
    return scn_rule_sort(row_ptr, N, row_pts, stream);
}
int scn_subm_map(const uint64_t* row_keys, int N, const uint64_t* tk, const int32_t* tv, uint32_t cap, int fx, int fy,
                 int fz, int32_t* map, scn_stream_t stream) {
    SCN_REQUIRE(is_pow2(cap), "hash: capacity must be a power of two (got %u)", cap);
    SCN_REQUIRE((fx & 1) && (fy & 1) && (fz & 1) && fx > 0 && fy > 0 && fz > 0 && fx * fy * fz <= 343,
                "subm_map: filter must be odd and <= 7^3 (got %dx%dx%d)", fx, fy, fz);
    if (N <= 0) return SCN_OK;
    const int K = fx * fy * fz, half = K / 2;
    if (half > 0)      // upper half: -1 wherever the mirrored write does not land
        cudaMemsetAsync(map + (int64_t)(half + 1) * N, 0xFF, sizeof(int32_t) * (size_t)half * N, as_stream(stream));
    k_subm_map<<<grid_for((int64_t)(half > 0 ? half : 1) * N, 256, 8), 256, 0, as_stream(stream)>>>(row_keys, N, tk, tv, cap - 1, fx, fy,
                                                                                                   fz, map);
    return check_launch("subm_map");
}
int scn_stride_keys(const uint64_t* row_keys, int N, int sx, int sy, int sz, uint64_t* parent_keys, int32_t* offs,
                    scn_stream_t stream) {
    SCN_REQUIRE(sx > 0 && sy > 0 && sz > 0, "stride_keys: stride must be positive");
    if (N <= 0) return SCN_OK;
    k_stride_keys<<<grid_for(N, TB), TB, 0, as_stream(stream)>>>(row_keys, N, sx, sy, sz, parent_keys, offs);
    return check_launch("stride_keys");
}
int scn_strided_maps(const int32_t* parent_row, const int32_t* offs, int n_in, int n_out, int K, int32_t* cmap,
                     int32_t* dmap, scn_stream_t stream) {
    if (n_in <= 0) return SCN_OK;
    k_strided_maps<<<grid_for(n_in, TB), TB, 0, as_stream(stream)>>>(parent_row, offs, n_in, n_out, K, cmap, dmap);
    return check_launch("strided_maps");
}
// A strided level around its host round trip: parent keys + the coarse level's count half, then its finish half + the
// child / parent maps (cmap is cleared to -1 here).
int scn_strided_level_count(const uint64_t* fine_keys, int n_in, int sx, int sy, int sz, uint64_t* parent_keys, int32_t* offs,
                            uint64_t* tk, int32_t* tv, uint32_t cap, int32_t* first, int32_t* rank, int32_t* scan_tmp,
                            scn_stream_t stream) {
    int rc = scn_stride_keys(fine_keys, n_in, sx, sy, sz, parent_keys, offs, stream);
    if (rc) return rc;
    return scn_level_count(parent_keys, n_in, tk, tv, cap, first, rank, scan_tmp, stream);
}
int scn_strided_level_finish(const uint64_t* parent_keys, const int32_t* offs, int n_in, uint64_t* tk, int32_t* tv,
                             uint32_t cap, const int32_t* rank, int32_t* parent_row, uint64_t* row_keys, int n_out, int K,
                             int32_t* cmap, int32_t* dmap, scn_stream_t stream) {
    int rc = scn_level_finish(parent_keys, n_in, tk, tv, cap, rank, parent_row, row_keys, stream);
    if (rc) return rc;
    if (n_out > 0 && K > 0) cudaMemsetAsync(cmap, 0xFF, sizeof(int32_t) * (size_t)K * n_out, as_stream(stream));
    return scn_strided_maps(parent_row, offs, n_in, n_out, K, cmap, dmap, stream);
}
int scn_batch_offsets(const uint64_t* row_keys, int N, int n_seg, int32_t* seg_ptr, scn_stream_t stream) {
    SCN_REQUIRE(n_seg >= 1, "batch_offsets: n_seg must be >= 1");
    k_batch_offsets<<<grid_for(N + 1, TB), TB, 0, as_stream(stream)>>>(row_keys, N, n_seg, seg_ptr);
    return check_launch("batch_offsets");
}

}  // extern "C"
