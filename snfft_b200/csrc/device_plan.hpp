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

// Row-group phase of a level (levels >= 9): see plan.hpp.  Every group of a panel comes with a
// "program blob" that the host lays out once (capi.cu: build_phase_devs): the row table followed by
// the group's blocks as a byte stream that a warp copies into shared memory and then interprets
// for many column chunks - no pointer chasing and no table decoding in the inner loop.
//   blob = u32 xoff[nrows], u32 dest[nrows] (padded to 16 bytes), then per block
//          a 16-byte header {u16 byte offset of its m rows in the warp's tile [5]; u16 m; u32 pair flag}
//          followed by the m x m matrix (doubles, row-major out x in, padded to 16 bytes),
//          terminated by a header with m = 0.
// Rows that stay on the x side come first, rows finished in this phase last.
struct alignas(16) U128 {
  uint32_t x, y, z, w;
};
struct PhaseGroupDev {
  uint32_t blob_off;   // in 16-byte units, into PhaseDev::blob_f / blob_i
  uint32_t blob_len;   // 16-byte units (row table + stream + terminator)
  uint32_t nrows;      // low 16 bits: rows; high 16 bits: rows that are NOT finished in this phase
  uint32_t reserved;
};
struct PhasePanelDev {
  int32_t d;           // columns of the panel
  int32_t tiles;       // ceil(d / 32); 1 when several lg members share the warp
  int32_t grp_off;     // first group in PhaseDev::groups
  int32_t ngroups;
  int32_t pack;        // narrow panels (d <= 16): floor(32 / d) members of the lg loop side by side
  int32_t magic;       // ceil(65536 / d): (c * magic) >> 16 = c / d for c < 256
};
struct PhaseDev {
  const PhasePanelDev* panels;
  const PhaseGroupDev* groups;
  const U128* blob_f;    // forward programs
  const U128* blob_i;    // inverse programs (blocks in reverse order, transposed weighted matrices)
  int64_t in_stride, out_stride;  // group strides of the X-side and spectrum-side buffers
  int32_t npanels;
  int32_t max_rows;       // rows of the largest group (tile of a warp = max_rows x 32 doubles)
  int32_t max_blob;       // 16-byte units of the largest blob
  int32_t nstages;
};

// In-place phase 1 of a large level as generated register programs: one entry per (panel, upper
// chain down to W^5); rows (i, t), i <= kP1Stages, t < d_sigma at base + i * (k-1)! + t * d.
struct P1Entry {
  int32_t sigma;     // program (shape of W^5)
  int32_t d;         // columns and row pitch of the panel
  int64_t base;      // in_off + ubase * d
  int64_t item_off;  // prefix sum of ceil(d / 32) over the entries
};
// A row that is finished after the generated stages (1..kP1Stages): forward out[dest] = scale_f * x[xoff]; inverse
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
  const uint32_t* item_entry;  // item (entry, 32-column tile) -> entry: one table read instead of a binary search
  const uint32_t* item_row;    // item (finished row, 32-column tile) -> row
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

// Sampled inverse (kernels.cu: inverse_at_kernel): one item = columns [c0, c0 + w) of one irrep.
struct InvAtItem {
  int32_t irrep, c0, w, pad;
};
struct InvAtIrrep {
  int32_t d, pad;
  int64_t off;             // packed offset of f_hat(lambda)
  const int32_t* partner;  // [(n-1)][d] generator tables (as for yor_kernel)
  const double* diag;
  double weight;           // d_lambda / n!
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
