// Host-side plan of the S_n transform: Young lattice, tableau dimensions, packed layout and the
// staged block programs of every level.  Pure C++ (no CUDA) so it can be built and inspected on a
// CPU-only box.  Reference anchors: utils.py:52-67 (partitions order), young_tableau.py:62-81
// (branch_down order), :193-206 (last-letter order), yor.py:87-117 (generator coefficients),
// fft.py:96-112 (the level step that the block programs factorise).
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

namespace snfft {

constexpr int kMaxN = 12;
constexpr int kMaxBlock = 5;  // a shape with <= 11 cells has <= 4 corners -> blocks are <= 5x5

struct Branch {
  int idx;      // index of the neighbouring shape in its own level (partitions order)
  int row, col; // the cell that is removed (down) / added (up), 0-based
  int64_t off;  // down only: row/column offset of the mu-block inside the lambda matrix
};

struct ShapeInfo {
  std::vector<int> rows;     // the partition
  int64_t dim = 0;           // number of standard tableaux
  int64_t off = 0;           // offset of the d x d block in the packed spectrum of its level
  std::vector<Branch> down;  // young_tableau.py:62-81 order (rows top -> bottom)
  std::vector<Branch> up;    // rows top -> bottom
};

struct LevelShapes {
  std::vector<ShapeInfo> shapes;  // utils.partitions(k) order
};

// One small dense block of a stage, in place on `slot`:
//   forward  out[a] = sum_b F[a][b] in[b]      (F at coef_f, row-major m x m)
//   inverse  in[b]  = sum_a I[b][a] out[a]     (I at coef_i, row-major m x m)
// I is F transposed with the Schur weights d_lambda/(k d_mu) folded in; all output scalings
// (sqrt(1 - 1/dist^2) chains) are folded into F, so loads and stores never scale.
// final_mask bit a: output a is the last write of its slot in the forward direction (equivalently
// the first read of the inverse), i.e. it is the finished row that goes to `dest[slot]`.
struct alignas(16) Block {
  uint16_t slot[kMaxBlock];
  uint8_t m;
  uint8_t final_mask;
  uint32_t coef_f;  // offsets (in doubles) into the level's coefficient pool
  uint32_t coef_i;
  uint32_t xoff;    // offset in the k! input of the fresh row X_{j+1}[W] (slot[0]), start of its run
  uint32_t xoff1;   // stage 1 only: offset of the X_1[W] row (slot[1]); else 0xffffffff
  uint32_t reserved;
};
static_assert(sizeof(Block) == 32, "Block must be 32 bytes");

// ---- row-group phases (levels >= 9) -------------------------------------------------------------
// The k-1 stages of a level are cut into a few consecutive ranges ("phases", phase_ranges()).
// Inside one phase the rows of a panel fall into groups that are closed under the phase's blocks
// (connected components of "block touches row"), so a CTA can load a bundle of groups for a tile of
// columns into shared memory, run all their blocks stage by stage and store them back: one
// coalesced pass over HBM per phase.  Phase 0 is always stages 1..kP1Stages (generated in-place programs).
struct PhaseRow {
  uint32_t xoff;   // offset of the row in the X-side buffer (its in-place slot), start of the run
  uint32_t dest;   // offset of the finished row in the spectrum-side buffer
  uint32_t final;  // 1: the row is finished in this phase (forward: store to dest; inverse: load from dest)
};
struct PhaseBlock {        // host only; the device form is BundleBlock (device_plan.hpp)
  uint16_t loc[kMaxBlock];  // rows of the block, as indices into the group's row list
  uint8_t m;
  uint8_t stage;            // 1-based stage of the level
  uint32_t coef_f, coef_i;
};
struct PhaseGroup {
  uint32_t row_off, blk_off;  // into the panel's row / block arrays of this phase
  uint32_t nrows, nblk;
};
struct PanelPhase {
  int stage_lo = 0, stage_hi = 0;  // stages [lo, hi], 1-based
  std::vector<PhaseGroup> groups;  // row_off / blk_off relative to this panel's vectors
  std::vector<PhaseRow> rows;
  std::vector<uint32_t> row_slot;  // slot of every entry of `rows` (host only; shards remap offsets)
  std::vector<PhaseBlock> blocks;  // per group: stage-major
  int max_rows = 0;
};

// Program of one column block mu |- k-1 of level k: k*d rows ("slots"), k-1 stages of d blocks.
struct PanelProgram {
  int k = 0, mu = 0;
  int d = 0;                       // d_mu
  int64_t in_off = 0;              // packed offset of mu in a level-(k-1) spectrum
  std::vector<Block> blocks;       // stage-major: stage s (1..k-1) is blocks[(s-1)*d, s*d)
  std::vector<uint32_t> dest;      // slot -> offset in the k! output (start of a run of d)
  std::vector<uint32_t> rowbase;   // slot -> offset in the k! input (start of a run of d)
  std::vector<PanelPhase> phases;  // k >= 9 only
  // ---- host-only metadata for the code generator (gen_programs.cpp) ----
  std::vector<uint32_t> coef_raw;  // per block: forward matrix before the final scales are folded in
  std::vector<int32_t> block_chain;  // per block: index of its chain W in tab(mu)
  std::vector<double> fscale;      // slot -> pending scale of the finished row (folded into coef_f)
  std::vector<double> weight;      // slot -> d_lambda / (k d_mu)
  std::vector<int32_t> lam_dim;    // slot -> d_lambda of the finished row (row pitch of dest)
  // two-phase split at stage 5 (k >= 6): rows are closed under stages 1..4 inside one
  // "upper chain" (mu > .. > W^5) and under stages >= 5 inside one lower tableau of size 4
  std::vector<int32_t> low_rho;    // slot -> id (partitions(4) order) of the shape of its lower tableau
  std::vector<int32_t> low_idx;    // slot -> index of that lower tableau inside tab(rho)
  // identity of the C_5 row held by a slot after stage 4 (-1 for rows X_i, i > 5):
  // C_5[(T^5, rho, t) ; (upper chain .., U^5, U^4)]
  std::vector<int32_t> c5_t5;      // id of T^5 (partitions(5) order)
  std::vector<int32_t> c5_u4;      // id of U^4 (partitions(4) order)
  std::vector<int32_t> c5_ubase;   // first chain index of the upper chain (mu > .. > U^5)
  std::vector<double> c5_scale;    // pending scale of that C_5 row at the boundary (1 for X rows)
  std::vector<int32_t> chain_sigma;  // chain -> id of W^5 (partitions(5) order)
  std::vector<int32_t> chain_base;   // chain -> first index of the chains sharing its upper part down to W^5
};

struct LevelProgram {
  int k = 0;
  std::vector<PanelProgram> panels;  // one per mu in partitions(k-1) order
  std::vector<double> coef;          // deduplicated coefficient matrices
  double macs = 0;                   // multiply-adds per group of k! elements
};

struct HostPlan {
  int n = 0;
  std::vector<LevelShapes> levels;     // levels[0..n]; levels[0] holds the empty shape
  std::vector<LevelProgram> programs;  // programs[k], k = 2..n (0,1 unused)
  int64_t fact[kMaxN + 2];
};

// Throws std::runtime_error on invalid n.
void build_host_plan(int n, HostPlan& plan);

constexpr int kPhaseMinLevel = 9;   // levels >= 9 run as row-group phases
// Stages 1..kP1Stages of a level >= 9 run as generated in-place register programs (one per shape of
// W^5: the rows (i, t), i <= kP1Stages, t < d_sigma of one upper chain are closed under them: <= 36
// rows); the windows of the phase kernel start at stage kP1Stages + 1.
constexpr int kP1Stages = 5;
constexpr int kPhaseMaxRows = 64;   // rows of a group (the phase kernel keeps the row table in two registers per lane)

// Stage ranges of the phases of level k >= 9: [1..kP1Stages] first, then the ranges given by their
// last stages.  Default: windows of two stages, 6..7, 8..9, .. (groups of <= 8 rows: four column
// chunks per warp in the phase kernel); the environment variable SNFFT_WINDOWS overrides it, e.g.
// "12:9,11/11:10" (per level the last stage of every phase after the first).
std::vector<std::pair<int, int>> phase_ranges(int k);

// Dense rho_lambda((j j+1)) for shape `irrep` of level k (row-major d*d), from lattice chains.
void yor_trans_dense(const HostPlan& plan, int k, int irrep, int j, double* out);

// Per-row description of rho_lambda((j j+1)): partner row (or -1) and the diagonal coefficient.
void yor_trans_rows(const HostPlan& plan, int k, int irrep, int j, std::vector<int32_t>& partner,
                    std::vector<double>& diag);

// coset-major flat index -> lexicographic rank (host).
int64_t coset_to_lex(int n, int64_t flat, const int64_t* fact);

}  // namespace snfft
