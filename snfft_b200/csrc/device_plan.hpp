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

constexpr int kFusedMax = 7;  // 2 * 7! doubles = 80,640 B of shared memory

struct FusedDev {
  int32_t k0;
  int32_t k0fact;
  LevelDev level[kFusedMax + 1];  // level[k], k = 2..k0
};

}  // namespace snfft
