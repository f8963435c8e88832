// Tile book of a Morton-ordered level (built by k_tile_book in conv_ts.cu, consumed by the tile-local kernels conv_ts.cu and
// conv_wgrad_ts.cu): per 128-row output tile the distinct input rows its 27 x 128 neighbour references hit (the halo set), the
// neighbour map re-expressed as 16-bit indices into that list, and which offsets have any active pair.
#pragma once
#include "tc_common.cuh"

namespace scn {

constexpr int TS_K = 27;
constexpr int TS_ROWS_CAP = 512;                     // rows stored per tile in the book (local ids beyond: 0xFFFE)
// per-tile blob (one bulk copy): [0] number of active offsets, [1 .. 27] the active offsets in processing order (centre
// first; the kernel reads the same order from the sequence word), [64 ..) the local neighbour map uint16 [27][128]
constexpr int TS_BLOB_LMAP = 64;
constexpr int TS_BLOB_BYTES = TS_BLOB_LMAP + TS_K * TILE_M * 2;      // 6976
constexpr uint32_t TS_INACTIVE = 0xFFFFu, TS_GLOBAL = 0xFFFEu;

struct TileBook {
    const uint8_t* blobs;      // [n_tiles][TS_BLOB_BYTES]
    const int32_t* rows;       // [n_tiles][TS_ROWS_CAP] distinct input rows of the tile, ascending
    const int32_t* nloc;       // [n_tiles] rows stored
    const uint32_t* useq;      // [n_tiles] active offsets as a bit sequence in processing order (seq_offset)
    int n_out, n_tiles;
};
static inline int64_t align256(int64_t v) { return (v + 255) / 256 * 256; }
// Processing order of a tile's active offsets as a bit sequence: bit 0 = the centre offset (13), bits 1..13 = offsets 0..12,
// bits 14..26 = offsets 14..26; units are the set bits in ascending order.  Every role derives its units from this one word
// with integer instructions -- no dependent shared-memory loads on the per-unit chains.
__host__ __device__ __forceinline__ uint32_t seq_offset(uint32_t b) { return b == 0u ? 13u : (b <= 13u ? b - 1u : b); }
__host__ __device__ __forceinline__ uint32_t seq_bit(uint32_t o) { return o == 13u ? 0u : (o < 13u ? o + 1u : o); }
static inline TileBook book_layout(const void* base, int n_out) {
    TileBook b;
    b.n_out = n_out, b.n_tiles = (n_out + TILE_M - 1) / TILE_M;
    const char* p = reinterpret_cast<const char*>(base);
    int64_t off = 0;
    b.blobs = reinterpret_cast<const uint8_t*>(p + off), off += align256((int64_t)b.n_tiles * TS_BLOB_BYTES);
    b.rows = reinterpret_cast<const int32_t*>(p + off), off += align256((int64_t)b.n_tiles * TS_ROWS_CAP * 4);
    b.nloc = reinterpret_cast<const int32_t*>(p + off), off += align256((int64_t)b.n_tiles * 4);
    b.useq = reinterpret_cast<const uint32_t*>(p + off), off += align256((int64_t)b.n_tiles * 4);
    return b;
}
static inline int64_t book_bytes(int n_out) {
    const int64_t nt = (n_out + TILE_M - 1) / TILE_M;
    return align256(nt * TS_BLOB_BYTES) + align256(nt * TS_ROWS_CAP * 4) + 2 * align256(nt * 4);
}

// the book attached to a neighbour map (scn_tile_book_attach); false if there is none for this map / row count
bool tile_book_lookup(const int32_t* map, int n_out, TileBook* out);

// PTX helpers of the tile-local kernels
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
        "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

}  // namespace scn
