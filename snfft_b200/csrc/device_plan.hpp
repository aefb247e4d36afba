// Device-resident view of a plan (plain structs passed to kernels by value).
#pragma once
#include <cstdint>

#include "plan.hpp"

namespace snfft {

struct PanelDev {
  int32_t d;         // d_mu
  int32_t rows;      // k * d_mu
  int32_t in_off;    // packed offset of mu inside a level-(k-1) spectrum
  int32_t blk_off;   // first block of stage 1 in LevelDev::blocks (stage s at +(s-1)*d)
  int32_t slot_off;  // first slot in dest / fscale / iscale / rowbase
  int32_t item_off;  // prefix sum of d*d over panels
};

struct LevelDev {
  int32_t k;
  int32_t npanels;
  int32_t kfact;    // k!
  int32_t km1fact;  // (k-1)!
  // distance between consecutive groups on the input side (k sub-spectra) and on the output side
  // (one spectrum); both k! for a whole transform, smaller for a column shard (snfft_shard_*)
  int64_t in_stride;
  int64_t out_stride;
  const PanelDev* panels;
  const Block* blocks;
  const double* coef;
  const uint32_t* dest;      // slot -> offset of the finished row in the k! output
  const uint32_t* rowbase;   // slot -> offset of the row in the k! input
  // only for the level run table-driven inside the fused kernel (k = 7), else null:
  const uint32_t* item_tab;  // [(k-1)!] : (panel << 24) | (block-in-stage << 12) | column
};

// A unit of work of the streaming level kernel: columns [c0, c0+w) of one panel, `cap` groups per
// CTA iteration (smem = cap * rows * w doubles).
struct Tile {
  int32_t panel;
  int32_t c0;
  int32_t w;
  int32_t cap;
};

// Row-group phase of a level (levels >= 9): see plan.hpp.  The groups of a panel are packed into
// "bundles" (as many whole groups as fit the shared-memory tile of a CTA); a CTA loads the rows of
// one bundle for a tile of 32 * cols_per_lane columns, runs the bundle's blocks stage by stage
// (one barrier per stage, warps take blocks round-robin, lanes = columns) and stores the rows.
struct alignas(16) BundleBlock {
  uint16_t loc[kMaxBlock];  // rows of the block, as indices into the bundle's row list
  uint8_t m;
  uint8_t pad;
  uint32_t coef;            // offset (doubles) of the m x m matrix in the level's coefficient pool
};
static_assert(sizeof(BundleBlock) == 16, "BundleBlock must be 16 bytes");
struct Bundle {
  uint32_t row_off;    // first row in BundleDev::row_xoff / row_dest
  uint32_t blk_off;    // first block in BundleDev::blocks_*
  uint32_t stage_off;  // stage_first[stage_off + s * 17 + w] .. [+ w + 1]: blocks of warp w in the s-th stage (relative)
  uint32_t nrows;      // low 16 bits: rows; high 16 bits: rows that are NOT finished in this phase (they come first)
};
struct BundlePanelDev {
  int32_t d;           // columns of the panel
  int32_t tiles;       // ceil(d / (32 * cols_per_lane)); 1 when several lg members share the tile
  int32_t bundle_off;  // first bundle in BundleDev::bundles
  int32_t nbundles;
  int64_t item_off;    // prefix sum of nbundles * tiles
  int32_t pack;        // narrow panels: floor(32 * cols_per_lane / d) members of the lg loop side by side
  int32_t magic;       // ceil(65536 / d): (c * magic) >> 16 = c / d for c < 256
};
struct BundleDev {
  const BundlePanelDev* panels;
  const Bundle* bundles;
  const uint32_t* row_xoff;     // per row: offset of its in-place slot in the x-side buffer
  const uint32_t* row_dest;     // per row: offset of the finished row in the spectrum-side buffer
  const BundleBlock* blocks_f;  // forward coefficients
  const BundleBlock* blocks_i;  // inverse coefficients (same rows)
  const uint32_t* stage_first;
  const double* coef;
  int64_t nitems;
  int64_t in_stride, out_stride;  // group strides of the X-side and spectrum-side buffers
  int32_t npanels;
  int32_t max_rows;       // rows of the largest bundle (shared-memory tile = max_rows x 32 * cols_per_lane)
  int32_t cols_per_lane;  // 1, 2, 4 or 8
  int32_t nstages;
};

// In-place phase 1 of a large level as generated register programs: one entry per (panel, upper
// chain down to W^5); rows (i, t), i < 5, t < d_sigma at base + i * (k-1)! + t * d.
struct P1Entry {
  int32_t sigma;     // program (shape of W^5)
  int32_t d;         // columns and row pitch of the panel
  int64_t base;      // in_off + ubase * d
  int64_t item_off;  // prefix sum of ceil(d / 32) over the entries
};
// A row that is finished after stage 4: forward out[dest] = scale_f * x[xoff]; inverse
// x[xoff] = scale_i * out[dest].
struct P0Row {
  uint32_t xoff, dest;
  int32_t d;
  int32_t pad;
  double scale_f, scale_i;
  int64_t item_off;
};
struct P1Dev {
  const P1Entry* entries;
  const P0Row* rows;
  int64_t nitems, nrow_items;
  int64_t in_stride, out_stride;
  int64_t rI;  // (k-1)!
  int32_t nentries, nrows;
};

constexpr int kFusedMax = 7;  // 2 * 7! doubles = 80,640 B of shared memory

struct FusedDev {
  int32_t k0;
  int32_t k0fact;
  LevelDev level[kFusedMax + 1];  // level[k], k = 2..k0
};

// Arguments of the peer-memory exchange of the sharded transform (kernels.cu: shard_pull_kernel)
constexpr int kMaxWorld = 16;
struct ShardPull {
  const double* peer[kMaxWorld];  // per rank: its column-shard buffer [all blocks][size(r, m)]
  int64_t first[kMaxWorld + 1];   // first block of every rank; first[world] = number of blocks
  int32_t off[kMaxWorld + 1];     // idx[off[r] .. off[r+1]) = positions of rank r's columns in an S_m spectrum
  const int32_t* idx;             // device, m! entries
  int64_t nblocks, mfact;
  int32_t rank, world;
};

// Device-side barrier between ranks (kernels.cu: shard_barrier_kernel).  pads[r] = rank r's signal
// pad in peer-accessible memory: unsigned long long [channels][kMaxWorld].
struct ShardBarrier {
  unsigned long long* pads[kMaxWorld];
  unsigned long long epoch;
  int* status;  // device word of the calling rank, set to 1 on a timeout
  int32_t rank, world, channel;
};

}  // namespace snfft
