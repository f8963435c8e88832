// Tile-local TF32 gather-GEMM for submanifold 3^3 convolutions on spatially coherent (Morton-ordered) rows.
//
//   out[r, :] = epi( bias + sum_o  in[map[o][r], :] . W[o] )       r in [0, n_out),  27 offsets
//
// Round 1's kernel (conv_tc.cu) gathers the 128 input rows of every (tile, offset) unit from L2 with cp.async: 27 x 128
// row copies per tile of which 63 % are zero fill, bounded by the LDGSTS rate and by the SS-form MMA reading its A operand
// from shared memory (measured: 40 cycles per M128 N32 K8 MMA from shared memory, 16 from tensor memory;
// profiles/r2_b_tile_local_kernel.md).  With Morton-ordered rows the 27 x 128 references of a tile hit only ~280 distinct
// rows, so here
//   * a per-level TILE BOOK (k_tile_book, built once per rulebook) lists each tile's distinct input rows (its halo set),
//     re-expresses the neighbour map as 16-bit indices into that list and records which of the 27 offsets have any active
//     pair at all (sequence word, centre offset first: the centre map is the identity, so the first MMA of a tile -- which
//     overwrites the accumulator -- initialises every row);
//   * loader warps copy the halo set of the NEXT tile into shared memory once (cp.async, padded row pitch);
//   * gather warps, thread = output row, read the row's neighbour for one offset from shared memory (LDS.128, odd pitch)
//     and store it into TENSOR MEMORY (tcgen05.st.32x32b): the A operand of the MMA lives in TMEM.  A row without a
//     neighbour issues no shared-memory request; its registers hold zeros;
//   * one elected thread issues tcgen05.mma kind::tf32 in the TS form (A from TMEM, B = W[o] from shared memory); fp32
//     accumulation in TMEM across the tile's offsets, double-buffered accumulators, same epilogue as conv_tc.cu.
// Units are handed from the gather warps to the MMA warp in batches of G (one unit per gather group): the issuing thread's
// serial chain is paid per batch.  Weights stay RESIDENT in shared memory when all 27 blocks fit (C <= 32: 108 KB), else
// they are streamed through a ring by cp.async.bulk.  One persistent CTA per SM:
//   warps 0-3 epilogue | 4 .. 4+4G-1 gather (G groups x 4 TMEM lane quarters) | 2 halo loaders | weight streamer | MMA issuer.
// Rows whose local index does not fit the shared-memory halo buffer (rare) are gathered from global memory by the same thread.
// What bounds it (clock-stamp traces and ncu, profiles/r2_b_tile_local_kernel.md): the shared-memory / tensor-memory data
// pipe.  Per unit the gather moves 128 x C x 4 bytes through LDS.128 (4.7 wavefronts per instruction: 71 % of the
// quarter-warps hold an active row, 1.37 wavefronts per active quarter from bank conflicts between non-adjacent halo rows)
// and again through tcgen05.st; every other access of a warp queues behind those, so the per-unit chains carry no dependent
// shared-memory loads (units come from the sequence word by integer instructions, neighbour codes are read two units
// ahead, the MMA warp reads no shared memory).
#include <atomic>
#include <mutex>
#include <type_traits>
#include <unordered_map>
#include "tile_book.cuh"

namespace scn {

// ------------------------------------------------------------------------------------------------ tile book builder
constexpr int TBB_THREADS = 256;
constexpr int TBB_SLOTS = 4096;      // >= 27 * 128 = 3456 distinct rows in the worst case
constexpr int TBB_SORT_MAX = 1024;   // halo sets beyond this are not coherent anyway: keep list order

__global__ void __launch_bounds__(TBB_THREADS) k_tile_book(const int32_t* __restrict__ map, int n_out, uint8_t* __restrict__ blobs,
                                                           int32_t* __restrict__ rows, int32_t* __restrict__ nloc,
                                                           uint32_t* __restrict__ useq) {
    __shared__ int32_t tab[TBB_SLOTS];
    __shared__ uint16_t ids[TBB_SLOTS];
    __shared__ int32_t vals[TS_K * TILE_M];
    __shared__ uint16_t rank[TS_K * TILE_M];
    __shared__ int warp_sums[TBB_THREADS / 32];
    __shared__ uint32_t any_active[TS_K];
    const int tile = blockIdx.x, row0 = tile * TILE_M, tid = threadIdx.x;
    for (int i = tid; i < TBB_SLOTS; i += TBB_THREADS) tab[i] = -1;
    if (tid < TS_K) any_active[tid] = 0u;
    __syncthreads();
    auto slot_of = [](int v) { return (int)(((uint32_t)v * 2654435761u) >> 20); };      // 12 bits
    for (int e = tid; e < TS_K * TILE_M; e += TBB_THREADS) {
        const int o = e >> 7, row = row0 + (e & 127);
        const int v = row < n_out ? __ldg(map + (int64_t)o * n_out + row) : -1;
        if (v >= 0) {
            int s = slot_of(v);
            while (true) {
                const int prev = atomicCAS(&tab[s], -1, v);
                if (prev == -1 || prev == v) break;
                s = (s + 1) & (TBB_SLOTS - 1);
            }
        }
    }
    __syncthreads();
    // distinct rows -> compact list (slot order), then local id = RANK of the row among the tile's rows: ids ascend with the
    // global (Morton) row index, so the 8 consecutive output rows a quarter warp gathers for one offset read mostly
    // consecutive halo rows -- conflict-free under the odd row pitch of k_conv_ts (a hash-order numbering measured 6.5
    // wavefronts per LDS.128 instead of 4)
    constexpr int PER = TBB_SLOTS / TBB_THREADS;      // 16 consecutive slots per thread
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) cnt += tab[tid * PER + j] >= 0 ? 1 : 0;
    const int lane = tid & 31, w = tid >> 5;
    int inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) warp_sums[w] = inc;
    __syncthreads();
    int off = 0, total = 0;
#pragma unroll
    for (int j = 0; j < TBB_THREADS / 32; ++j) {
        if (j < w) off += warp_sums[j];
        total += warp_sums[j];
    }
    int run = off + inc - cnt;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        const int s = tid * PER + j, v = tab[s];
        if (v >= 0) {
            vals[run] = v;
            ids[s] = (uint16_t)run;      // position in the compact list, replaced by the rank below
            ++run;
        }
    }
    __syncthreads();
    if (total <= TBB_SORT_MAX) {
        for (int i = tid; i < total; i += TBB_THREADS) {
            const int v = vals[i];
            int r = 0;
            for (int k = 0; k < total; ++k) r += vals[k] < v ? 1 : 0;      // all threads read the same word: broadcast
            rank[i] = (uint16_t)r;
        }
    } else {
        for (int i = tid; i < total; i += TBB_THREADS) rank[i] = (uint16_t)(i < 65534 ? i : 65534);
    }
    __syncthreads();
    for (int i = tid; i < total; i += TBB_THREADS) {
        const int r = rank[i];
        if (r < TS_ROWS_CAP) rows[(int64_t)tile * TS_ROWS_CAP + r] = vals[i];
    }
    if (tid == 0) nloc[tile] = total < TS_ROWS_CAP ? total : TS_ROWS_CAP;
    uint8_t* blob = blobs + (int64_t)tile * TS_BLOB_BYTES;
    uint16_t* lmap = reinterpret_cast<uint16_t*>(blob + TS_BLOB_LMAP);
    for (int e = tid; e < TS_K * TILE_M; e += TBB_THREADS) {
        const int o = e >> 7, r = e & 127, row = row0 + r;
        const int v = row < n_out ? __ldg(map + (int64_t)o * n_out + row) : -1;
        uint32_t code = TS_INACTIVE;
        if (v >= 0) {
            int s = slot_of(v);
            while (tab[s] != v) s = (s + 1) & (TBB_SLOTS - 1);
            const uint32_t rk = rank[ids[s]];
            code = rk < (uint32_t)TS_ROWS_CAP ? rk : TS_GLOBAL;
            any_active[o] = 1u;      // benign race: every writer stores the same value
        }
        lmap[e] = (uint16_t)code;
    }
    __syncthreads();
    if (tid == 0) {
        // processing order: the centre offset first (its map is the identity, so the first MMA of a tile -- which overwrites
        // the accumulator -- initialises every valid row), then the other offsets with at least one active pair, ascending
        auto active = [&](int o) { return any_active[o] != 0u; };
        int n = 0;
        uint32_t seq = 0;
        if (active(TS_K / 2)) blob[1 + n++] = (uint8_t)(TS_K / 2), seq |= 1u;
        for (int o = 0; o < TS_K; ++o)
            if (o != TS_K / 2 && active(o)) blob[1 + n++] = (uint8_t)o, seq |= 1u << seq_bit((uint32_t)o);
        blob[0] = (uint8_t)n;
        useq[tile] = seq;
        for (int i = 1 + n; i < TS_BLOB_LMAP; ++i) blob[i] = 0xFF;
    }
}

// ------------------------------------------------------------------------------------------------ the convolution
struct ConvTsParams {
    const float* in;
    int ld_in;
    TileBook book;
    const int32_t* map;      // global neighbour map [27][n_out] (rows beyond the shared-memory halo)
    const uint8_t* image;
    const float* bias;
    const float* residual;
    int ld_res;
    const float* mask;
    int ld_mask;
    float* out;
    int ld_out, epi;
    float* out2;             // optional second output, out2 = epi2(out) (common.cuh: epi2_apply)
    int ld_out2, epi2;
    int cap;                 // halo rows per shared-memory buffer
};

// D[tmem] (+)= A[tmem] . B[smem desc]
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(db), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}

// NCH 16-byte chunks of one shared-memory row into v[0 .. 4 NCH): immediate offsets (the padded pitch needs no swizzle)
template <int NCH>
__device__ __forceinline__ void lds_row(uint32_t rb, uint32_t (&v)[32]) {
#define TS_LD(J, OFF)                                                                                                  \
    if (J < NCH)                                                                                                       \
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4+" #OFF "];"                                               \
                     : "=r"(v[4 * J]), "=r"(v[4 * J + 1]), "=r"(v[4 * J + 2]), "=r"(v[4 * J + 3])                     \
                     : "r"(rb));
    TS_LD(0, 0) TS_LD(1, 16) TS_LD(2, 32) TS_LD(3, 48) TS_LD(4, 64) TS_LD(5, 80) TS_LD(6, 96) TS_LD(7, 112)
#undef TS_LD
}

#ifdef SCN_TS_TRACE
// timing experiment: clock64 stamps of CTA 0 (slot -> SM cycles), read back with scn_debug_ts_trace
__device__ long long g_ts_trace[16384];
#define TS_STAMP(slot)                                                                                   \
    do {                                                                                                 \
        if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (slot) < 16384) g_ts_trace[slot] = clock64(); \
    } while (0)
#else
#define TS_STAMP(slot) \
    do {               \
    } while (0)
#endif

struct UnitRing {      // position in a ring of `n` stages + the parity of the current round
    int s;
    uint32_t par;
    __device__ __forceinline__ void step(int n) {
        if (++s == n) s = 0, par ^= 1;
    }
};

// compile-time shape of a layer: C input = output channels, weights resident in shared memory or streamed, G gather groups
template <int C, bool RESIDENT, int GROUPS>
struct TsShape {
    static constexpr int NKB = (C + KB - 1) / KB;
    static constexpr int PITCH = C * 4 + 16;                        // odd number of 16-byte chunks: row r, chunk j sits in bank
    static constexpr int CPR = C / 4;                               // group (r * odd + j) mod 8 -- no swizzle arithmetic needed
    static constexpr int WBLOCK = NKB * C * 128;                    // bytes of one offset's packed weights
    static constexpr int NWB = RESIDENT ? TS_K : (WBLOCK <= 12288 ? 4 : 3);   // >= G: a batch never waits for its own release
    static constexpr int NW = RESIDENT ? 1 : NWB;                   // weight barriers
    static constexpr int NA_MAX = ((512 - 2 * C) / C) > 12 ? 12 : ((512 - 2 * C) / C);
    static constexpr int NA = NA_MAX / GROUPS * GROUPS;             // A stages: NB batches of G
    static constexpr int NK8 = C / 8;
    static constexpr int G = GROUPS;
    static constexpr int THREADS = (8 + 4 * G) * 32;
    static constexpr int NB = NA / GROUPS;
    static constexpr int N_BARS = 8 + 2 * NB + 2 * NW;
    static constexpr int FIXED_SMEM = 1024 + NWB * WBLOCK + 2 * TS_BLOB_BYTES + 512;      // + 2 * cap * PITCH
};

// CL > 1 (streamed weights only): the CTAs of a cluster work on consecutive tiles IN LOCKSTEP -- every tile walks all 27
// offsets in the same order -- so that a weight block is read from L2 ONCE per cluster: CTA r fetches the r-th 1/CL of the
// block and multicasts it into the same ring stage of every CTA of the cluster (cp.async.bulk ... .multicast::cluster, the
// completion bytes land on each CTA's own barrier); a stage is refilled when all CL consumers have released it
// (tcgen05.commit ... .multicast::cluster arrives on every CTA's w_empty).  Why: with streamed weights every CTA pulled
// every 12-16 KB block it used from L2 -- 390 MB per level-0-sized layer at C = 48, 6 TB/s at 65 us: L2 -> SM bandwidth
// on the WEIGHTS was the bound (profiles/r2_b).
template <int C, bool RESIDENT, int GROUPS, int CL = 1>
__global__ void __launch_bounds__(TsShape<C, RESIDENT, GROUPS>::THREADS, 1) k_conv_ts(const ConvTsParams p) {
    using S = TsShape<C, RESIDENT, GROUPS>;
    constexpr bool MC = CL > 1 && !RESIDENT;
    constexpr uint32_t ALL_UNITS = (1u << TS_K) - 1u;
    constexpr int NB = S::NB, NW = S::NW, PITCH = S::PITCH, CPR = S::CPR, G = S::G;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t w_base = smem_base;
    const uint32_t halo_bytes = (uint32_t)p.cap * PITCH;
    const uint32_t halo0 = w_base + (uint32_t)(S::NWB * S::WBLOCK);
    const uint32_t blob0 = halo0 + 2u * halo_bytes;
    const uint32_t bars = blob0 + 2u * TS_BLOB_BYTES;
    auto halo_full = [&](int b) { return bars + 8u * b; };
    auto halo_empty = [&](int b) { return bars + 8u * (2 + b); };
    auto acc_full = [&](int b) { return bars + 8u * (4 + b); };
    auto acc_empty = [&](int b) { return bars + 8u * (6 + b); };
    // A stages are handed over in BATCHES of G units (one per gather group): the MMA warp's serial chain -- an mbarrier
    // wait costs ~90 cycles even when the barrier completed long ago, measured with clock stamps -- is paid once per
    // batch, and the stages of a batch are released by one commit
    auto b_full = [&](uint32_t b) { return bars + 8u * (8 + b); };
    auto b_empty = [&](uint32_t b) { return bars + 8u * (8 + NB + b); };
    auto w_full = [&](int s) { return bars + 8u * (8 + 2 * NB + s); };
    auto w_empty = [&](int s) { return bars + 8u * (8 + 2 * NB + NW + s); };
    const uint32_t tmem_slot = bars + 8u * S::N_BARS;

    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    // warp roles; the issue arbiter of an SM sub-partition prefers the HIGHEST warp id (B300 guide): the MMA issuer sits on top
    constexpr int W_GATHER0 = 4, W_LOAD0 = 4 + 4 * G, W_WEIGHTS = W_LOAD0 + 2, W_MMA = W_LOAD0 + 3;
    if (tid == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(halo_full(b), 64 + 1);            // 2 loader warps x 32 lanes (cp.async arrivals) + the blob bulk copy
            mbar_init(halo_empty(b), 4 * G);            // one arrival per gather warp
            mbar_init(acc_full(b), 1);
            mbar_init(acc_empty(b), 4);                 // one arrival per epilogue warp
        }
        for (int b = 0; b < NB; ++b) {
            mbar_init(b_full(b), 4 * G);                // one arrival per gather warp (each group fills one unit of the batch)
            mbar_init(b_empty(b), 1);
        }
        for (int s = 0; s < NW; ++s) {
            mbar_init(w_full(s), 1);
            mbar_init(w_empty(s), MC ? CL : 1);      // every consumer CTA of the cluster releases the stage
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    if (MC) cluster_sync_all();      // every CTA's barriers exist before a peer multicasts into them
    pdl_trigger();
    pdl_wait();
    if (warp == 0) TS_STAMP(16000);

    const int n_tiles = p.book.n_tiles;
    constexpr uint32_t A_COL0 = 2u * C;      // A stages follow the two accumulators

    if (warp >= W_GATHER0 && warp < W_LOAD0) {
        // ===================== gather warps: shared-memory halo -> registers -> tensor memory =====================
        // Batch j of a tile = its active offsets number j*G .. j*G+G-1 (processing order of the blob header); group g fills
        // unit g of every batch.  The batch ring position (br) advances alike in every role.
        const int g = (warp - W_GATHER0) >> 2, q = warp & 3;
        const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16) + A_COL0 + (uint32_t)(g * C);
        const uint32_t cap = (uint32_t)p.cap;
        UnitRing br{0, 0};
        int it = 0;
        [[maybe_unused]] int tb = 0;      // trace: batch counter
        int nl_next = blockIdx.x < n_tiles ? __ldg(p.book.nloc + blockIdx.x) : 0;
        uint32_t seq_next = blockIdx.x < n_tiles ? __ldg(p.book.useq + blockIdx.x) : 0u;
        uint32_t v[32];
        bool dirty = false;      // v holds a real row (must be zeroed before it stands in for a missing neighbour)
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
#pragma unroll 1
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            const int nl = nl_next;
            uint32_t seq = MC ? ALL_UNITS : seq_next;
            {
                const int tn = tile + gridDim.x;      // the next tile's header, one tile ahead
                if (tn < n_tiles) nl_next = __ldg(p.book.nloc + tn), seq_next = __ldg(p.book.useq + tn);
            }
            const bool slow = nl > (int)cap;          // some references live outside the shared-memory halo
            if (g == 0 && q == 0) TS_STAMP(13000 + 2 * it);
            mbar_wait(halo_full(buf), (uint32_t)(it >> 1) & 1u);
            if (g == 0 && q == 0) TS_STAMP(13000 + 2 * it + 1);
            const uint32_t hb = halo0 + (uint32_t)buf * halo_bytes;
            const uint32_t blob = blob0 + (uint32_t)buf * TS_BLOB_BYTES;
            const uint32_t lm = blob + TS_BLOB_LMAP + (uint32_t)(q * 32 + lane) * 2u;
            const int row = tile * TILE_M + q * 32 + lane;
            const int n_units = __popc(seq);
            // Software pipeline of one gather warp: the shared-memory reads of its NEXT unit are issued right after the
            // tensor-memory stores of the current one (the stores read their source registers when they issue), so they
            // run under the store completion wait, the hand-over and the stage wait.  Measured before (clock stamps,
            // profiles/r2_b): all gather warps read in phase (~600 cycles of saturated LDS per batch), then stored and
            // waited in phase (~550 cycles with an idle load/store unit).
            constexpr int CH0 = (C < 32 ? C : 32) / 4;      // 16-byte chunks of the first register pass
            auto fetch = [&](uint32_t cur, uint32_t o_cur, uint32_t byte0, int nch_sel) {
                // rows without a neighbour issue no shared-memory request and write no zeros
                if (cur < cap) {
                    const uint32_t rb = hb + cur * PITCH + byte0;
                    if (nch_sel == 0) lds_row<CH0>(rb, v);
                    else lds_row<(C > 32 ? C - 32 : 4) / 4>(rb, v);
                    dirty = true;
                } else if (slow && cur != TS_INACTIVE) {
                    dirty = true;
                    const float* gsrc = p.in + (int64_t)__ldg(p.map + (int64_t)o_cur * p.book.n_out + row) * p.ld_in + (byte0 >> 2);
                    const int nc = nch_sel == 0 ? CH0 * 4 : C - 32;
#pragma unroll
                    for (int c0 = 0; c0 < 32; c0 += 4)
                        if (c0 < nc) {
                            const uint4 t = __ldg(reinterpret_cast<const uint4*>(gsrc + c0));
                            v[c0] = t.x, v[c0 + 1] = t.y, v[c0 + 2] = t.z, v[c0 + 3] = t.w;
                        }
                } else if (dirty && nch_sel == 0) {      // a row without a neighbour multiplies zeros
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0u;
                    dirty = false;
                }
            };
            // own units: every G-th set bit of seq starting at bit number g; the neighbour codes are read TWO own units ahead
            // (a dependent load behind the other warps' queued 16-byte reads costs hundreds of cycles, measured)
#pragma unroll
            for (int i = 0; i < G - 1; ++i)
                if (i < g) seq &= seq - 1u;
            auto pop = [&]() -> uint32_t {      // offset of the next own unit (0xFF: none), advances seq by G units
                const uint32_t o = seq ? seq_offset((uint32_t)__ffs((int)seq) - 1u) : 0xFFu;
#pragma unroll
                for (int i = 0; i < G; ++i) seq &= seq - 1u;      // 0 & 0xFFFFFFFF stays 0
                return o;
            };
            uint32_t o_a = pop();
            uint32_t code_a = o_a != 0xFFu ? lds_u16(lm + o_a * (TILE_M * 2)) : TS_INACTIVE;
            uint32_t o_b = pop();
            uint32_t code_b = o_b != 0xFFu ? lds_u16(lm + o_b * (TILE_M * 2)) : TS_INACTIVE;
            fetch(code_a, o_a, 0u, 0);
#pragma unroll 1
            for (int nb = (n_units + G - 1) / G; nb > 0; --nb) {
                const bool have = o_a != 0xFFu;
                if (q == 0 && have) TS_STAMP(4096 + 16 * tb + 4 * g);
                const uint32_t cur = code_a, o_cur = o_a;
                o_a = o_b, code_a = code_b;
                o_b = pop();
                code_b = o_b != 0xFFu ? lds_u16(lm + o_b * (TILE_M * 2)) : TS_INACTIVE;
                const bool more = o_a != 0xFFu;
                mbar_wait(b_empty(br.s), br.par ^ 1u);      // the stage's previous occupant has been multiplied
                if (q == 0 && have) TS_STAMP(4096 + 16 * tb + 4 * g + 1);
                if (have) {
                    tc_fence_after();
                    const uint32_t tcol = tq + (uint32_t)br.s * (uint32_t)(G * C);
                    if (C >= 32) tmem_st32(tcol, v);
                    else tmem_st16(tcol, v);
                    if (C > 32) {      // channels 32 .. C-1: the same registers again (C = 48: 16 more, C = 64: 32 more)
                        fetch(cur, o_cur, 128u, 1);
                        if (C - 32 >= 32) tmem_st32(tcol + 32u, v);
                        else tmem_st16(tcol + 32u, v);
                    }
                }
                if (more) fetch(code_a, o_a, 0u, 0);         // the next unit's rows: in flight across the hand-over
                if (have) {
                    if (q == 0) TS_STAMP(4096 + 16 * tb + 4 * g + 2);
                    tmem_st_wait();
                    tc_fence_before();
                    if (q == 0) TS_STAMP(4096 + 16 * tb + 4 * g + 3);
                }
                ++tb;
                __syncwarp();
                if (elect_one()) mbar_arrive(b_full(br.s));      // also for a partial last batch without a unit for this group
                br.step(NB);
            }
            __syncwarp();
            if (elect_one()) mbar_arrive(halo_empty(buf));
        }
    } else if (warp == W_LOAD0 || warp == W_LOAD0 + 1) {
        // ===================== halo loaders: the next tile's distinct input rows, once =====================
        const int lw = warp - W_LOAD0;
        int it = 0;
        const char* in_c = reinterpret_cast<const char*>(p.in);
        const uint32_t row_bytes = (uint32_t)p.ld_in * 4u;
        int cnt_next = blockIdx.x < n_tiles ? __ldg(p.book.nloc + blockIdx.x) : 0;      // one tile ahead: no dependent load chain
#pragma unroll 1
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            int cnt = cnt_next;
            if (tile + (int)gridDim.x < n_tiles) cnt_next = __ldg(p.book.nloc + tile + gridDim.x);
            if (cnt > p.cap) cnt = p.cap;
            const int32_t* rl = p.book.rows + (int64_t)tile * TS_ROWS_CAP;
            // super-round = 32 consecutive rows of the list: lane L holds the index of row base + L, the copies fetch it by
            // shuffle; the two loader warps alternate super-rounds.  All of a tile's indices are requested up front (one
            // coalesced load per super-round, independent of each other): with a cold L2 a load-per-round chain cost one
            // DRAM latency per 32 rows, more than the tile's compute time (bench.py flushes L2 before every launch)
            constexpr int ROUNDS = TS_ROWS_CAP / 64;
            int idx[ROUNDS];
#pragma unroll
            for (int k = 0; k < ROUNDS; ++k) {
                const int at = lw * 32 + 64 * k + lane;
                idx[k] = at < cnt ? __ldg(rl + at) : -1;
            }
            if (lw == 0) TS_STAMP(12000 + 4 * it);
            mbar_wait(halo_empty(buf), ((uint32_t)(it >> 1) & 1u) ^ 1u);
            if (lw == 0) TS_STAMP(12000 + 4 * it + 1);
            const uint32_t hb = halo0 + (uint32_t)buf * halo_bytes;
            if (lw == 0 && elect_one()) {
                mbar_arrive_expect_tx(halo_full(buf), TS_BLOB_BYTES);
                bulk_g2s(blob0 + (uint32_t)buf * TS_BLOB_BYTES, p.book.blobs + (int64_t)tile * TS_BLOB_BYTES, TS_BLOB_BYTES,
                         halo_full(buf));
            }
#pragma unroll
            for (int k = 0; k < ROUNDS; ++k) {
                const int base = lw * 32 + 64 * k;
                if (base < cnt) {
                    const int mine = idx[k];
#pragma unroll
                    for (int i = 0; i < CPR; ++i) {
                        const int t = lane + 32 * i;              // chunk t of the super-round: row t / CPR, chunk t % CPR
                        const int jr = t / CPR, c = t % CPR;
                        const int ridx = __shfl_sync(0xffffffffu, mine, jr);
                        if (ridx >= 0) {
                            const char* src = in_c + (uint64_t)(uint32_t)ridx * row_bytes + c * 16;
                            const uint32_t dst = hb + (uint32_t)(base + jr) * PITCH + (uint32_t)(c * 16);
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                        }
                    }
                }
            }
            cp_async_mbar_arrive_noinc(halo_full(buf));
            if (lw == 0) TS_STAMP(12000 + 4 * it + 2);
        }
        cp_async_wait_all();
    } else if (warp == W_WEIGHTS) {
        // ===================== weights: resident (one load) or streamed through a ring, one block per unit =====================
        if (RESIDENT) {
            if (elect_one()) {
                mbar_arrive_expect_tx(w_full(0), (uint32_t)(TS_K * S::WBLOCK));
                for (int o = 0; o < TS_K; ++o)
                    bulk_g2s(w_base + (uint32_t)(o * S::WBLOCK), p.image + (size_t)o * S::WBLOCK, (uint32_t)S::WBLOCK, w_full(0));
            }
        } else {
            UnitRing wr{0, 0};
            const uint32_t rank = MC ? cluster_ctarank() : 0u;
            constexpr uint32_t PART = (uint32_t)S::WBLOCK / (MC ? CL : 1);
            // MC: every CTA of the cluster runs the same number of rounds (a CTA without a tile in the last round still
            // fetches its share of every block and releases the stages: its peers depend on both)
            const int rounds = (n_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
#pragma unroll 1
            for (int it = 0; it < rounds; ++it) {
                const int tile = (int)blockIdx.x + it * (int)gridDim.x;
                if (!MC && tile >= n_tiles) break;
                uint32_t seq = MC ? ALL_UNITS : __ldg(p.book.useq + tile);      // the tile's units in processing order
#pragma unroll 1
                for (; seq; seq &= seq - 1u) {
                    const int o = (int)seq_offset((uint32_t)__ffs((int)seq) - 1u);
                    mbar_wait(w_empty(wr.s), wr.par ^ 1u);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(w_full(wr.s), (uint32_t)S::WBLOCK);
                        const uint32_t dst = w_base + (uint32_t)(wr.s * S::WBLOCK) + rank * PART;
                        const uint8_t* src = p.image + (size_t)o * S::WBLOCK + rank * PART;
                        if (MC) bulk_g2s_multicast(dst, src, PART, w_full(wr.s), (uint16_t)((1u << CL) - 1u));
                        else bulk_g2s(dst, src, PART, w_full(wr.s));
                    }
                    wr.step(NW);
                }
            }
        }
    } else if (warp == W_MMA) {
        // ===================== MMA issuer: A from tensor memory, B = W[o] from shared memory =====================
        const uint32_t idesc = make_idesc_tf32(TILE_M, C);
        const uint64_t wdesc0 = make_desc_sw128(w_base);
        constexpr uint32_t WBLOCK_D = (uint32_t)S::WBLOCK >> 4, KB_D = (uint32_t)(C * 128) >> 4;
        const uint32_t ta0 = tmem_base + A_COL0;
        UnitRing br{0, 0}, wr{0, 0};
        int it = 0;
        [[maybe_unused]] int tb = 0;
        if (RESIDENT) mbar_wait(w_full(0), 0);
        const int rounds = (n_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
#pragma unroll 1
        for (int tile = blockIdx.x; it < rounds; tile += gridDim.x, ++it) {
            if (tile >= n_tiles) {
                if constexpr (MC) {
                    // no tile in the last round: consume the weight stages like the peers do, multiply nothing
#pragma unroll 1
                    for (int u = 0; u < TS_K; ++u) {
                        mbar_wait(w_full(wr.s), wr.par);
                        if (elect_one()) mma_commit_multicast(w_empty(wr.s), (uint16_t)((1u << CL) - 1u));
                        __syncwarp();
                        wr.step(NW);
                    }
                    continue;
                } else {
                    break;
                }
            }
            const int b = it & 1;
            uint32_t seq = MC ? ALL_UNITS : __ldg(p.book.useq + tile);      // the tile's units; this warp reads no shared memory
            mbar_wait(acc_empty(b), ((uint32_t)(it >> 1) & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)(b * C);
            uint32_t accum = 0;
#pragma unroll 1
            while (seq) {
                // one batch: up to G units, ONE barrier wait, one elected region, one commit
                TS_STAMP(4 * tb);
                uint32_t o[G];
#pragma unroll
                for (int i = 0; i < G; ++i) {
                    o[i] = seq ? seq_offset((uint32_t)__ffs((int)seq) - 1u) : 0xFFu;
                    seq &= seq - 1u;
                }
                mbar_wait(b_full(br.s), br.par);
                tc_fence_after();
                TS_STAMP(4 * tb + 1);
                const bool leader = elect_one();
                if (leader) {
                    const uint32_t ta = ta0 + (uint32_t)br.s * (uint32_t)(G * C);
#pragma unroll
                    for (int i = 0; i < G; ++i) {
                        if (o[i] != 0xFFu) {
                            if (!RESIDENT) {
                                mbar_wait(w_full(wr.s), wr.par);
                                tc_fence_after();
                            }
                            const uint64_t wd = wdesc0 + (uint64_t)((RESIDENT ? o[i] : (uint32_t)wr.s) * WBLOCK_D);
#pragma unroll
                            for (int kk = 0; kk < S::NK8; ++kk)
#ifdef SCN_TS_NOMMA
                                if (p.cap < 0)
#endif
                                mma_tf32_ts(tmem_d, ta + (uint32_t)(i * C + kk * 8),
                                            wd + (uint64_t)((uint32_t)(kk >> 2) * KB_D + (uint32_t)((kk & 3) * 2)), idesc,
                                            (i == 0 && kk == 0) ? accum : 1u);
                            if (!RESIDENT) {
                                if (MC) mma_commit_multicast(w_empty(wr.s), (uint16_t)((1u << CL) - 1u));
                                else mma_commit(w_empty(wr.s));
                                wr.step(NW);
                            }
                        }
                    }
                    mma_commit(b_empty(br.s));      // the batch's stages may be refilled
                }
                if (!RESIDENT) {      // keep the ring position warp-uniform (only the elected lane advanced it)
                    const int src = __ffs(__ballot_sync(0xffffffffu, leader)) - 1;
                    wr.s = __shfl_sync(0xffffffffu, wr.s, src);
                    wr.par = __shfl_sync(0xffffffffu, wr.par, src);
                }
                TS_STAMP(4 * tb + 2);
                ++tb;
                accum = 1u;
                br.step(NB);
            }
            __syncwarp();
            if (elect_one()) {
                mma_commit(acc_full(b));
            }
        }
    } else if (warp < 4) {
        // ===================== epilogue warps 0..3 (as conv_tc.cu, whole tiles only) =====================
        const int epi = p.epi;
        const bool vec_ok = (p.ld_out % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0) &&
                            (!p.out2 || ((p.ld_out2 % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.out2) & 15) == 0))) &&
                            (!(epi & SCN_EPI_ADD) || ((p.ld_res % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.residual) & 15) == 0))) &&
                            (!(epi & SCN_EPI_MASK) || ((p.ld_mask % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.mask) & 15) == 0)));
        auto finish = [&](float x, float m, float r) {      // order: (bias), MASK, ADD, RELU, ROUND
            if ((epi & SCN_EPI_MASK) && !(m > 0.f)) x = 0.f;
            if (epi & SCN_EPI_ADD) x += r;
            if (epi & SCN_EPI_RELU) x = fmaxf(x, 0.f);
            if (epi & SCN_EPI_ROUND) {
                uint32_t t;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(x));
                x = __uint_as_float(t);
            }
            return x;
        };
        int it = 0;
#pragma unroll 1
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int b = it & 1;
            if (warp == 0) TS_STAMP(14000 + 4 * it);
            mbar_wait<2000>(acc_full(b), (uint32_t)(it >> 1) & 1u);
            tc_fence_after();
            if (warp == 0) TS_STAMP(14000 + 4 * it + 1);
            const int row = tile * TILE_M + warp * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(b * C);
#pragma unroll 1
            for (int c0 = 0; c0 < C; c0 += 16) {
                float v[16];
                tmem_ld16(taddr + c0, v);
                if (row < p.book.n_out) {
                    float* orow = p.out + (int64_t)row * p.ld_out + c0;
                    float* orow2 = p.out2 ? p.out2 + (int64_t)row * p.ld_out2 + c0 : nullptr;
                    const float* rrow = (epi & SCN_EPI_ADD) ? p.residual + (int64_t)row * p.ld_res + c0 : nullptr;
                    const float* mrow = (epi & SCN_EPI_MASK) ? p.mask + (int64_t)row * p.ld_mask + c0 : nullptr;
                    if (p.bias) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] += __ldg(p.bias + c0 + j);
                    }
                    if (vec_ok) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float4 r4 = rrow ? *reinterpret_cast<const float4*>(rrow + j) : make_float4(0, 0, 0, 0);
                            const float4 m4 = mrow ? *reinterpret_cast<const float4*>(mrow + j) : make_float4(1, 1, 1, 1);
                            float4 x;
                            x.x = finish(v[j], m4.x, r4.x), x.y = finish(v[j + 1], m4.y, r4.y);
                            x.z = finish(v[j + 2], m4.z, r4.z), x.w = finish(v[j + 3], m4.w, r4.w);
                            *reinterpret_cast<float4*>(orow + j) = x;
                            if (orow2) *reinterpret_cast<float4*>(orow2 + j) = epi2_apply4(x, p.epi2);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float y = finish(v[j], mrow ? mrow[j] : 1.f, rrow ? rrow[j] : 0.f);
                            orow[j] = y;
                            if (orow2) orow2[j] = epi2_apply(y, p.epi2);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (warp == 0) TS_STAMP(14000 + 4 * it + 2);
            if (elect_one()) mbar_arrive(acc_empty(b));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (MC) cluster_sync_all();      // no CTA leaves while a peer may still write into its shared memory / barriers
    if (warp == W_MMA) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

// ------------------------------------------------------------------------------------------------ registry: map -> book
struct BookEntry {
    const void* book;
    int n_out;
};
static std::unordered_map<const void*, BookEntry> g_books;
static std::mutex g_books_mutex;
static std::atomic<int64_t> g_ts_launches{0};

bool tile_book_lookup(const int32_t* map, int n_out, TileBook* out) {
    std::lock_guard<std::mutex> lock(g_books_mutex);
    auto it = g_books.find(map);
    if (it == g_books.end() || it->second.n_out != n_out) return false;
    *out = book_layout(it->second.book, n_out);
    return true;
}

template <int C, bool RESIDENT, int GROUPS, int CL = 1>
static int launch_ts(ConvTsParams& p, cudaStream_t stream) {
    using S = TsShape<C, RESIDENT, GROUPS>;
    constexpr int MAX_SMEM = 227 * 1024;
    int cap = (MAX_SMEM - S::FIXED_SMEM) / (2 * S::PITCH);
    if (cap > TS_ROWS_CAP) cap = TS_ROWS_CAP;
    cap &= ~7;
    if (cap < 192) return 0;
    p.cap = cap;
    const int smem = S::FIXED_SMEM + 2 * cap * S::PITCH;
    auto kern = k_conv_ts<C, RESIDENT, GROUPS, CL>;
    cudaError_t e = (cudaError_t)ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("conv_ts: cudaFuncSetAttribute(%d bytes): %s", smem, cudaGetErrorString(e));
        return -SCN_ERR_CUDA;
    }
    int grid = p.book.n_tiles < sm_count() ? p.book.n_tiles : sm_count();
    grid = grid / CL * CL;      // whole clusters
    if (grid < CL) return 0;
    PdlLaunch L(dim3(grid), dim3(S::THREADS), smem, stream, CL);
    e = cudaLaunchKernelEx(&L.cfg, kern, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("conv_ts: launch failed: %s", cudaGetErrorString(e));
        return -SCN_ERR_CUDA;
    }
    const int rc = check_launch("conv_ts");
    if (!rc) ++g_ts_launches;
    return rc ? -rc : 1;
}

// Launches the tile-local kernel when a tile book is attached to `map` and the layer qualifies; returns 1 if it did,
// 0 if the caller should use conv_tc.cu, < 0 (negated status) on error.
int conv_ts_try(const float* in, int ld_in, int Cin, const int32_t* map, int n_out, int K, const void* image, const float* bias,
                const float* residual, int ld_res, const float* mask, int ld_mask, float* out, int ld_out, int Cout, int epi,
                float* out2, int ld_out2, int epi2, cudaStream_t stream) {
    if (K != TS_K || !map || Cin != Cout) return 0;
    const char* ev = getenv("SCN_CONV_TS");      // read per call: tests run both kernels in one process
    if (ev && ev[0] == '0') return 0;
    if ((Cin != 16 && Cin != 32 && Cin != 48 && Cin != 64) || ld_in % 4 != 0 || (reinterpret_cast<uintptr_t>(in) & 15)) return 0;
    static int min_tiles = -1;
    if (min_tiles < 0) {
        const char* e = getenv("SCN_CONV_TS_MIN_TILES");
        min_tiles = e ? atoi(e) : 2 * sm_count();
    }
    const int n_tiles = (n_out + TILE_M - 1) / TILE_M;
    if (n_tiles < min_tiles) return 0;
    BookEntry be;
    {
        std::lock_guard<std::mutex> lock(g_books_mutex);
        auto it = g_books.find(map);
        if (it == g_books.end() || it->second.n_out != n_out) return 0;
        be = it->second;
    }
    ConvTsParams p;
    p.in = in, p.ld_in = ld_in, p.book = book_layout(be.book, n_out), p.map = map;
    p.image = reinterpret_cast<const uint8_t*>(image), p.bias = bias, p.residual = residual, p.ld_res = ld_res;
    p.mask = mask, p.ld_mask = ld_mask, p.out = out, p.ld_out = ld_out, p.epi = epi;
    p.out2 = out2, p.ld_out2 = ld_out2, p.epi2 = epi2;
    static int groups = -1;
    if (groups < 0) {
        const char* e = getenv("SCN_CONV_TS_GROUPS");
        groups = e ? atoi(e) : 4;
    }
    // streamed weights: CTAs per cluster sharing each weight block by multicast.  Parity-tested (SCN_CONV_TS_CLUSTER=2) but
    // measured neutral (C = 48: 69.2 vs 68.1 us, C = 64: 94.6 vs 89.8 us): halving the L2 -> SM weight bytes does not help
    // because the ring is bound by the LATENCY of a refill behind its own release, not by bandwidth; off by default
    const char* ce = getenv("SCN_CONV_TS_CLUSTER");
    const int cluster = ce ? atoi(ce) : 1;
    switch (Cin) {
        case 16: return groups == 6 ? launch_ts<16, true, 6>(p, stream) : launch_ts<16, true, 4>(p, stream);
        case 32: return groups == 6 ? launch_ts<32, true, 6>(p, stream) : launch_ts<32, true, 4>(p, stream);
        case 48: return cluster == 2 ? launch_ts<48, false, 4, 2>(p, stream) : launch_ts<48, false, 4>(p, stream);
        default:      // 6 A stages of 64 columns: two batches of three
            return cluster == 2 ? launch_ts<64, false, 3, 2>(p, stream) : launch_ts<64, false, 3>(p, stream);
    }
}

}  // namespace scn

using namespace scn;

extern "C" {

int64_t scn_tile_book_bytes(int n_out) { return n_out > 0 ? book_bytes(n_out) : 256; }

int scn_tile_book_build(const int32_t* map, int n_out, int K, void* book, scn_stream_t stream) {
    SCN_REQUIRE(K == TS_K, "tile_book: 3x3x3 submanifold maps only (K = %d)", K);
    SCN_REQUIRE(n_out >= 0 && (n_out == 0 || (map && book)), "tile_book: bad arguments");
    if (n_out == 0) return SCN_OK;
    SCN_REQUIRE((reinterpret_cast<uintptr_t>(book) & 255) == 0, "tile_book: buffer must be 256-byte aligned");
    TileBook b = book_layout(book, n_out);
    k_tile_book<<<b.n_tiles, TBB_THREADS, 0, as_stream(stream)>>>(map, n_out, const_cast<uint8_t*>(b.blobs), const_cast<int32_t*>(b.rows),
                                                                  const_cast<int32_t*>(b.nloc), const_cast<uint32_t*>(b.useq));
    return check_launch("tile_book");
}

int scn_tile_book_attach(const int32_t* map, const void* book, int n_out) {
    SCN_REQUIRE(map && book && n_out > 0, "tile_book_attach: bad arguments");
    std::lock_guard<std::mutex> lock(g_books_mutex);
    g_books[map] = BookEntry{book, n_out};
    return SCN_OK;
}

#ifdef SCN_TS_TRACE
int scn_debug_ts_trace(long long* out) {
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(out, scn::g_ts_trace, sizeof(long long) * 16384) == cudaSuccess ? 0 : 1;
}
#endif
int64_t scn_conv_ts_launch_count(void) { return g_ts_launches.load(); }

int scn_tile_book_detach_if(const int32_t* map, const void* book) {
    std::lock_guard<std::mutex> lock(g_books_mutex);
    auto it = g_books.find(map);
    if (it != g_books.end() && it->second.book == book) g_books.erase(it);
    return SCN_OK;
}

int scn_tile_book_detach(const int32_t* map) {
    std::lock_guard<std::mutex> lock(g_books_mutex);
    g_books.erase(map);
    return SCN_OK;
}

}  // extern "C"
