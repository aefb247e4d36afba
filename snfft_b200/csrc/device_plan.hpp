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

// Row-group phase of a level (levels >= 9): see plan.hpp
struct PhasePanelDev {
  int32_t d;         // columns of the panel
  int32_t tiles;     // ceil(d / (32 * cols_per_lane))
  int32_t grp_off;   // first group in PhaseDev::groups
  int32_t ngroups;
  int64_t item_off;  // prefix sum of ngroups * tiles
  int32_t pack;      // narrow panels (d <= 16): floor(32 / d) members of the lg loop share a warp
  int32_t reserved;
};

struct PhaseDev {
  const PhasePanelDev* panels;
  const PhaseGroup* groups;  // row_off / blk_off are absolute here
  const PhaseRow* rows;
  const PhaseBlock* blocks;
  const double* coef;
  int64_t nitems;
  int64_t in_stride, out_stride;  // group strides of the X-side and spectrum-side buffers
  int32_t npanels;
  int32_t max_rows;
  int32_t cols_per_lane;  // 1, 2 or 4: a warp covers 32 * cols_per_lane columns (PhasePanelDev::tiles)
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

}  // namespace snfft
