// Kernel launchers (definitions in kernels.cu).  All launches go on the given stream and bump the
// library's launch counter.
#pragma once
#include <cuda_runtime.h>

#include <atomic>

#include "device_plan.hpp"

namespace snfft {

extern std::atomic<int64_t> g_launch_count;
extern std::atomic<bool> g_force_table_driven;  // tests: route level 8 through the table-driven kernel too
extern std::atomic<bool> g_fused_tma;           // experiment: TMA bulk stage-in of the inverse fused kernel

// Optional per-launch timing (bench.py): when enabled every launch is bracketed by CUDA events on
// its own stream; timing_read() synchronises the recorded events and returns the totals per kind.
enum KernelKind { kKindFusedFwd = 0, kKindFusedInv, kKindLevel8Fwd, kKindLevel8Inv, kKindLevelTable,
                  kKindP1, kKindP0, kKindPhaseFwd, kKindPhaseInv, kKindExchange, kKindBarrier, kKindReorder,
                  kKindYor, kKindDgemm, kKindOther, kNumKinds };
void timing_enable(bool on);
void timing_read(double* ms_by_kind, int64_t* launches_by_kind);  // also resets the totals

// Streaming level kernel: global -> shared tile -> stages -> global.  table_driven = false lets
// level 8 of a whole transform run as generated register programs (they assume the full 8! pitch,
// so column shards always pass true).
cudaError_t launch_level(const LevelDev& L, const Tile* tiles_dev, int ntiles, size_t smem_bytes,
                         const double* in, double* out, int64_t groups, bool inverse, bool table_driven,
                         int num_sms, cudaStream_t stream);

// One row-group phase of a level: xside holds the in-place rows (input of the level, clobbered by
// the forward direction; output of the level in the inverse direction), sside the finished rows.
// panel_groups / panel_tiles / panel_pack: host copies of the per-panel fields of the device tables.
cudaError_t launch_phase(const PhaseDev& P, const int32_t* panel_groups, const int32_t* panel_tiles,
                         const int32_t* panel_pack, double* xside, double* sside, int64_t lgroups,
                         bool inverse, int num_sms, cudaStream_t stream);

// Stages 1..kP1Stages of a large level as in-place generated programs plus the pass that moves the
// rows finished by then to / from the spectrum-side buffer.
cudaError_t launch_p1_inplace(const P1Dev& P, double* xside, double* sside, int64_t lgroups, bool inverse,
                              int num_sms, cudaStream_t stream);

// Fused low levels 2..k0 of `groups` independent S_{k0} blocks, entirely in shared memory.
cudaError_t launch_fused(const FusedDev& F, const double* in, double* out, int64_t groups,
                         bool inverse, int num_sms, cudaStream_t stream);

// Element reorder lex <-> coset-major.  table may be null (index computed on the fly).
cudaError_t launch_reorder(int n, const uint32_t* table, const double* src, double* dst,
                           int64_t batch, bool to_coset, cudaStream_t stream);
cudaError_t launch_build_lex_table(int n, uint32_t* table, cudaStream_t stream);

// rho_lambda(sigma) for a list of permutations: rows tables are int32 partner[(n-1)][d] and
// double diag[(n-1)][d] for the generators s_1..s_{n-1} of one irrep.
cudaError_t launch_yor(int n, int d, const int32_t* partner, const double* diag,
                       const int32_t* perms, int64_t count, double* out, bool seminormal, cudaStream_t stream);

// f(sigma_j) for `count` permutations from one packed spectrum (all irreps in one launch); out is zeroed first.
cudaError_t launch_inverse_at(int n, const InvAtItem* items, int nitems, const InvAtIrrep* irreps, size_t smem_bytes,
                              const double* fhat, const int32_t* perms, int64_t count, double* out,
                              cudaStream_t stream);

// C[M,N] = A[M,K] * B[K,N], fp64 row-major, on the DMMA pipe (mma.sync.m8n8k4.f64).
cudaError_t launch_dgemm(const double* A, const double* B, double* C, int M, int N, int K,
                         cudaStream_t stream);

// DFT over C_3^m of `count` rows: forward real [count][3^m] -> complex at the nfreq listed frequency
// indices; inverse (adjoint): complex [count][nfreq] -> scale * complex [count][3^m].  out_im == nullptr:
// the output is interleaved complex128 in out_re; inverse with in_im == nullptr: so is the input.
cudaError_t launch_c3_dft(int m, const double* in_re, const double* in_im, const int32_t* freq, int nfreq,
                          int64_t count, double* out_re, double* out_im, bool inverse, double scale,
                          cudaStream_t stream);

// tables and layout helpers of the wreath plan (capi.cu: snfft_wreath_*)
cudaError_t launch_wreath_index(int n, int N, int nH, const int8_t* T, const int8_t* Tinv, const int8_t* H,
                                int64_t* idx, int64_t* inv_idx, cudaStream_t stream);
cudaError_t launch_wreath_permute(int N, int b, double* mat, double* planes, double scale, bool to_matrix,
                                  cudaStream_t stream);
// f(o_j, sigma_j) for `count` sampled states: sigma as lexicographic ranks, o as orientation indices
cudaError_t launch_wreath_inverse_at(int m, int N, int b, int nH, int64_t nfact, int n_ori, const int64_t* inv_idx,
                                     const double* R, const double* Mre, const double* Mim, const int32_t* freq,
                                     const int64_t* sigma, const int32_t* ori, int64_t count, double scale,
                                     double* out_re, double* out_im, cudaStream_t stream);
cudaError_t launch_transpose(const double* in, double* out, int64_t rows, int64_t cols, cudaStream_t stream);
// [batch][rows][cols] -> [batch][cols][rows]
cudaError_t launch_transpose_batched(const double* in, double* out, int64_t batch, int64_t rows, int64_t cols,
                                     cudaStream_t stream);
// X[pair][e3][e2][e1] (what the factor transforms leave) <-> P[pair][b*b] in the Kronecker order of R[h]
cudaError_t launch_wreath_kron_permute(int64_t pairs, int d1, int d2, int d3, double* X, double* P, bool to_kron,
                                       cudaStream_t stream);
cudaError_t launch_transpose_i64(const int64_t* in, int32_t* out, int64_t rows, int64_t cols, cudaStream_t stream);  // narrows to 32 bits
// stage 2 fused for small blocks (b*b in {1,4,9,16}, many cosets): gather + block sums + placement in one kernel
bool wreath_stage2_fits(int N, int b, int nH);
// F: interleaved complex128 [n!][N]; forward reads it through idx_t ([h][pair]), the adjoint writes it through inv_idx
cudaError_t launch_wreath_stage2(int N, int b, int nH, int64_t nfact, const int32_t* idx_t, const int64_t* inv_idx,
                                 const double* R, double* F, double* Mre, double* Mim, double* blocks, bool inverse,
                                 cudaStream_t stream);  // blocks: 2 N^2 b^2 doubles of scratch (adjoint)

// dst[e][0..width) = src[idx[e]][0..width)
cudaError_t launch_gather(const double* src, const int64_t* idx, double* dst, int64_t count, int width,
                          cudaStream_t stream);

// the exchange of the sharded transform as P2P loads from peer memory (see ShardPull)
cudaError_t launch_shard_pull(const ShardPull& X, double* local, bool inverse, cudaStream_t stream);

// device-side barrier between the ranks of a sharded transform (see ShardBarrier)
cudaError_t launch_shard_barrier(const ShardBarrier& B, cudaStream_t stream);

cudaError_t prepare_kernels();  // opt in to large dynamic shared memory (once per process)

}  // namespace snfft
