// formats.hpp — host descriptions of the device formats (see DESIGN.md "Data layout in HBM").
#pragma once
#include <cstdint>
#include <vector>

namespace spmv {

// ------------------------------------------------------------------------------------------
// WSP: CSR of A^T ("one list per output column", the orientation of the reference's
// CSRMatrix, matrix_csr.cpp:8-22) cut into 128-bit groups of four non-zeros:
//   vals  float4 [groups]            values, list order = ascending row
//   idx   ushort4|uint4 [groups]     row index of each value (16 bit when M < 65536)
//   colptr uint32 [N+1]              first group of each column (sentinel included)
// A column is padded to a multiple of 4 with (value 0, index M); the kernels keep one zero
// at x[M] so a pad contributes an exact 0 regardless of x.
// Row panels (tall matrices whose x does not fit shared memory, e.g. config 5): the rows are
// cut into `panels` panels of `panel_rows` rows, every (panel, column) pair has its own list
// (panel-major: colptr[p*N + c]), row ids are panel-local 16-bit and the pad index is
// `panel_rows`; a CTA works inside one panel with that panel's slice of x in shared memory, and
// the per-panel sums are added in panel order.
// ------------------------------------------------------------------------------------------
struct HostWsp {
    int64_t M = 0, N = 0, nnz = 0, groups = 0;
    int index_bits = 16;
    int panels = 1;
    int64_t panel_rows = 0;        // rows per panel (== M when panels == 1)
    std::vector<uint32_t> colptr;  // panels*N+1
    std::vector<float> vals;       // 4*groups
    std::vector<uint16_t> idx16;   // 4*groups (index_bits == 16)
    std::vector<uint32_t> idx32;   // 4*groups (index_bits == 32)
    int64_t max_col_groups = 0;
};

// ------------------------------------------------------------------------------------------
// Row-panel format shared by AWSP and TCSR.  A is cut into column slabs of `slab_cols`
// outputs (a power of two, 256 .. 4096).  The unit of storage is the *row segment* (one row
// of one slab): its non-zeros in ascending column order, cut into 128-bit groups of four:
//   vals  float4 [groups]                     values
//   idx   uchar4 (slab_cols == 256) | ushort4 column of each value inside the slab
// A segment is padded to a multiple of 4 with (value 0, smallest column absent from the
// segment): a pad adds an exact 0 to an accumulator no real entry of that row touches, so the
// kernels need no per-entry predicate and no two entries of a segment share a column.
// Segments are laid out slab-major, row-minor, so everything one (slab, row-range) CTA reads
// is one contiguous run, and the segment of a row with x[row] == 0 is simply not read.
//   AWSP: off[slab*(M+1) + row]  = first group of the segment (32-bit, row-addressable)
//   TCSR: 32 consecutive rows form a tile (the reference's tcsr.cpp tiles are 32x32):
//         tile_off[slab*(RB+1) + rb] = first group of the tile (32-bit)
//         rel[(slab*RB + rb)*32 + r] = first group of row r inside the tile (16-bit)
//         i.e. a two-level offset: 2 bytes per segment instead of 4.
// ------------------------------------------------------------------------------------------
//
// Lane-owned blocks (chunk_mode 3; very sparse matrices with fairly dense activations, e.g.
// BASELINE config 5): the multi-row schedule has to retire a chunk one row at a time because two
// rows may hit the same column.  Here lane l of a warp only ever handles columns congruent to l
// modulo 32, so no two lanes can meet in an accumulator (and bank = lane: no bank conflicts) and a
// chunk retires in a single pass whatever rows it mixes.  Rows are cut into blocks of `block_rows`;
// inside a (slab, block) every lane has its own stream of entries (row ascending), cut into groups
// of four and padded to the longest lane's length; group (g, lane) is stored at off[block] + 32 g +
// lane, so a warp's 32 groups are one contiguous 512-byte run:
//   vals  float4 [groups]     values (pads: 0)
//   idx   ushort4 [groups]    (row inside the block) << cbits | (column inside the slab) >> 5,
//                              cbits = log2(slab_cols / 32), block_rows = min(1024, 2^(16 - cbits))
//   off[slab*(blocks+1) + block]   first group of the block (32-bit, multiple of 32)
// Every stored non-zero is read whatever x is (x enters as a multiplier), so this form wins when
// more than about a third of the activations are non-zero.
constexpr int kTileRows = 32;
constexpr int kMinSlabCols = 256;
constexpr int kMaxSlabCols = 4096;
constexpr int kLobSlabCols = 2048;       // lane-owned blocks: 2048 columns x 1024 rows per block
constexpr int kLobMaxBlockRows = 1024;   // rows per block: min(this, 2^(16 - cbits))

struct HostPanel {
    int64_t M = 0, N = 0, nnz = 0, groups = 0, nonempty_segments = 0;
    int slab_cols = 256;
    int index_bits = 8;          // 8 when slab_cols == 256, else 16
    int slabs = 0;
    int row_blocks = 0;          // ceil(M/32)
    bool tiled = false;          // false: AWSP (per-row offsets), true: TCSR (two-level offsets)
    int block_rows = 0;          // > 0: lane-owned blocks of this many rows (off is per block)
    int lob_blocks = 0;          //      ceil(M / block_rows)
    std::vector<uint32_t> off;       // AWSP: slabs*(M+1);  TCSR: slabs*(row_blocks+1)
    std::vector<uint16_t> rel;       // TCSR: slabs*row_blocks*32
    std::vector<float> vals;         // 4*groups
    std::vector<uint8_t> idx8;       // 4*groups (index_bits == 8)
    std::vector<uint16_t> idx16;     // 4*groups (index_bits == 16)
    std::vector<int32_t> row_nnz;    // [M]  stored nnz per row (all slabs)   — traffic accounting
    std::vector<int32_t> row_groups; // [M]  groups per row (all slabs)       — traffic accounting
    std::vector<int32_t> row_segs;   // [M]  non-empty segments per row       — traffic accounting
};

// ------------------------------------------------------------------------------------------
// Row strips (chunk_mode 4; very sparse matrices, e.g. BASELINE config 5 at 1 %): the form that
// keeps both promises of the awsp variant when a row segment holds only a few non-zeros — a row
// with x[row] == 0 is never read, and what IS read comes in DRAM-friendly pieces.
//   * the N outputs are cut into strips of `strip_cols` columns (about 20 non-zeros per row and
//     strip), kStripsPerBand = 16 consecutive strips form a band;
//   * storage is band-major, row-minor, strip-minor: the 16 strip segments of one row of one band
//     are contiguous (one 2-3 KB run per active row), so the unit DRAM sees is the band row while
//     the unit a warp works on is its own strip of that row;
//   * an entry is 8 bytes {fp32 value, u32 (column inside the strip) + 1}; inside a (row, strip)
//     the entries are in ascending column order, so they are distinct accumulators; the all-zero
//     entry (what a zero-filled copy produces for an idle lane) addresses accumulator 0, which no
//     column uses;
//   * a segment is padded with all-zero entries to a multiple of kStripPad = 4 entries, so it starts
//     on a 32-byte sector boundary (a warp's load of ~20 entries then touches 6 sectors, not 7);
//   * soff[(band*M + row)*16 + strip] = first entry of the segment (32-bit; one sentinel at the end) —
//     the host / file order.  In HBM the table is strip-major, [band][strip 0..16][row] with strip 16
//     = the end of the row's band segment: a warp fetches its strip's starts lane = row, and 32
//     nearby rows of one strip share a few sectors (row-major records cost one sector per row:
//     2 of the 7 sectors a segment touched).
// The kernel (strips.cu) gives a band's row range to a 16-warp CTA, warp w owns strip w: one
// row segment = one 32-lane window, one entry per lane, no two lanes on one accumulator, so a
// window retires in a single pass whatever its length.
// ------------------------------------------------------------------------------------------
constexpr int kStripsPerBand = 16;
constexpr int kMaxStripCols = 2112;      // 16 x (strip accumulators + ring + row table) must fit 227 KB
constexpr int kStripPad = 4;             // entries per padding unit (32 bytes)
constexpr double kStripTargetNnz = 20.5; // mean non-zeros per (row, strip): P(> 32) stays under 1 %

struct HostStrips {
    int64_t M = 0, N = 0, nnz = 0;
    int strip_cols = 0;
    int bands = 0;
    std::vector<uint64_t> ent;       // value bits | (uint64)(column + 1) << 32
    std::vector<uint32_t> soff;      // bands*M*16 + 1
    std::vector<int32_t> row_nnz;    // [M] non-zeros per row (all bands)              — traffic accounting
    std::vector<int32_t> row_groups; // [M] stored entries per row / kStripPad (pads included)
};

} // namespace spmv
