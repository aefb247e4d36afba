/* snfft_b200 -- C ABI of the B200-native S_n Fourier transform hot path.
 *
 * The reference (horacepan/SnFFT) is pure Python and has no FFI of its own; its boundary for
 * this path is the set of Python call signatures in fft.py / yor.py.  This header is the C ABI
 * that the Python mirror of those signatures (snfft_b200/fft.py, yor.py) binds with ctypes, and
 * that a maintainer of the reference would bind the same way (INTEGRATION.md shows the stub).
 * Each entry point cites the reference interface it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary;
 *   - every function returns an int status (0 = OK); the message of the last failure on the
 *     calling thread is available from snfft_last_error(); nothing throws across the ABI;
 *   - device buffers are owned by the caller; a plan owns its tables and is immutable after
 *     creation (two lazily built, mutex-guarded caches excepted: the generator tables of snfft_yor
 *     and the matrix of snfft_dft_direct), so concurrent calls on distinct streams are safe;
 *   - process-wide state is limited to three atomics: the launch counter (snfft_launch_count), the
 *     timing switch (snfft_timing_enable) and the test hook snfft_debug_force_table_driven;
 *   - all device work is asynchronous on the stream that is passed (a cudaStream_t as void*);
 *   - there is no CPU fallback: calls that need a device fail with SNFFT_ERR_NO_DEVICE when the
 *     plan was created host-only (device = -1).
 *
 * Data layout
 *   functions on S_n : double[batch][n!]; order flag SNFFT_ORDER_LEX = rank of the permutation
 *                      tuple in itertools.permutations order (reference perm2.py:214-233 `sn`),
 *                      SNFFT_ORDER_COSET = coset-major order, flat = sum_k (i_k-1)(k-1)! for
 *                      sigma = c^(n)_{i_n} ... c^(2)_{i_2}, c^(k)_i = (i i+1 ... k)
 *                      (the order in which fft.py:96-98 visits the group);
 *   spectrum         : double[batch][n!], "packed": for lambda in utils.partitions(n) order
 *                      (reference utils.py:52-67) the row-major d_lambda x d_lambda matrix
 *                      f_hat(lambda), rows/columns in the reference's tableau order
 *                      (young_tableau.py:193-206).  sum_lambda d_lambda^2 = n!.
 */
#ifndef SNFFT_B200_H
#define SNFFT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNFFT_OK 0
#define SNFFT_ERR_INVALID 1    /* bad argument */
#define SNFFT_ERR_NO_DEVICE 2  /* host-only plan used for device work, or no CUDA device */
#define SNFFT_ERR_CUDA 3       /* a CUDA runtime call failed; see snfft_last_error() */
#define SNFFT_ERR_ALLOC 4

#define SNFFT_ORDER_COSET 0
#define SNFFT_ORDER_LEX 1

#define SNFFT_MAX_N 12

typedef struct snfft_plan snfft_plan;

/* Version of this ABI (bumped when a signature changes). */
int snfft_abi_version(void);

/* Message of the last failure on the calling thread ("" if none). */
const char* snfft_last_error(void);

/* Build the per-n tables: partitions in utils.partitions order, Young-lattice chains, tableau
 * dimensions, packed offsets, the staged block programs of every level 2..n, and upload them to
 * `device`.  device = -1 builds the host tables only (for inspection / CPU tests; such a plan
 * cannot launch).  Replaces the tableau brute force of young_tableau.py:98-125 and the
 * generator construction of yor.py:87-117 (K0 of SURVEY.md 7.1). */
int snfft_plan_create(int n, int device, snfft_plan** out);
int snfft_plan_destroy(snfft_plan* plan);

int snfft_plan_n(const snfft_plan* plan);
int snfft_plan_device(const snfft_plan* plan);
/* number of irreps of S_k, 1 <= k <= n */
int snfft_num_irreps(const snfft_plan* plan, int k);

/* Irreps of S_k in utils.partitions(k) order (utils.py:52-67).  parts: int32[n_irreps][k]
 * zero-padded rows; dims: int32[n_irreps] = number of standard tableaux (young_tableau.py:127);
 * offsets: int64[n_irreps] = start of the row-major d x d block in the packed spectrum.
 * Any output pointer may be NULL. */
int snfft_layout(const snfft_plan* plan, int k, int32_t* parts, int32_t* dims, int64_t* offsets);

/* branch_down of irrep `irrep` of S_k (young_tableau.py:62-81): indices into level k-1, in the
 * reference's order, and the row/column offset of each block inside the d x d matrix.
 * idx, off: int32[<= k]; *count receives the number of branches. */
int snfft_branch_down(const snfft_plan* plan, int k, int irrep, int32_t* idx, int32_t* off,
                      int32_t* count);

/* rho_lambda((j j+1)) for irrep `irrep` of S_k, dense row-major double[d*d] on the HOST, read off
 * the plan's chain-indexed coefficient tables (yor.py:87-117 `yor_trans`).  1 <= j < k. */
int snfft_yor_trans(const snfft_plan* plan, int k, int irrep, int j, double* out_host);

/* Forward transform f_hat(lambda) = sum_sigma f(sigma) rho_lambda(sigma) for every lambda |- n
 * (fft.py:16-37 `ft_full`, :72-113 `fft`, :39-70 `fft2`).  f_dev: double[batch][n!] in `order`;
 * fhat_dev: double[batch][n!] packed; work_dev: double[batch][n!] scratch (may alias neither).
 * f_dev is preserved.  All three buffers must be 16-byte aligned (the low levels stage their
 * input with 16-byte asynchronous copies); a misaligned pointer is rejected with
 * SNFFT_ERR_INVALID. */
int snfft_forward(const snfft_plan* plan, const double* f_dev, double* fhat_dev, double* work_dev,
                  int64_t batch, int order, void* stream);

/* Inverse transform f(sigma) = (1/n!) sum_lambda d_lambda tr(rho_lambda(sigma)^T f_hat(lambda))
 * (tile/par_inv_ft.py:27-45 evaluated for every sigma, by running the levels backwards).
 * fhat_dev is preserved.  Same 16-byte alignment requirement as snfft_forward. */
int snfft_inverse(const snfft_plan* plan, const double* fhat_dev, double* f_dev, double* work_dev,
                  int64_t batch, int order, void* stream);

/* One level of the recursion on its own (used by the sharded driver and by tests):
 * in: double[groups][k][(k-1)!] = k packed S_{k-1} spectra per group; out: double[groups][k!]
 * packed S_k spectra (fft.py:96-112).  inverse != 0 runs the level backwards. */
int snfft_level(const snfft_plan* plan, int k, const double* in_dev, double* out_dev,
                int64_t groups, int inverse, void* stream);

/* Reorder between lexicographic and coset-major element order (the closure chain
 * `f_i = lambda pi: f(cyc * pi)` of fft.py:97-98 as an index map).
 * to_coset != 0: dst[b][t] = src[b][lex(t)]; else dst[b][lex(t)] = src[b][t]. */
int snfft_reorder(const snfft_plan* plan, const double* src_dev, double* dst_dev, int64_t batch,
                  int to_coset, void* stream);

/* Host copy of the coset-major -> lexicographic index map, int64[n!]. */
int snfft_coset_to_lex_host(const snfft_plan* plan, int64_t* out_host);

/* rho_lambda(sigma) for `count` permutations (yor.py:52-85 `yor`): perms_dev int32[count][n]
 * 1-based tuples, out_dev double[count][d*d] row-major, for irrep `irrep` of S_n. */
int snfft_yor(const snfft_plan* plan, int irrep, const int32_t* perms_dev, int64_t count,
              double* out_dev, void* stream);

/* The same in Young's SEMINORMAL form (yor.py:119-182 `ysemi` / `ysemi_t`): generator blocks
 * [[1/d, 1 - 1/d^2], [1, -1/d]] on tableau pairs (t, t'), t before t'; same tables, same word. */
int snfft_ysemi(const snfft_plan* plan, int irrep, const int32_t* perms_dev, int64_t count,
                double* out_dev, void* stream);

/* Sampled inverse: f(sigma_j) for `count` permutations straight from ONE packed spectrum,
 * f(sigma) = (1/n!) sum_lambda d_lambda tr(rho_lambda(sigma)^T f_hat(lambda))
 * (tile/par_inv_ft.py:27-45 summed over the irreps; the per-state evaluation of
 * compute_policy.py:66-88).  All irreps in one launch; rho is never materialised (the generator
 * word of sigma is applied to the rows of f_hat(lambda) as 2-sparse rotations).
 * perms_dev int32[count][n] 1-based tuples, out_dev double[count]. */
int snfft_inverse_at(const snfft_plan* plan, const double* fhat_dev, const int32_t* perms_dev, int64_t count,
                     double* out_dev, void* stream);

/* Direct sum f_hat = sum_sigma f(sigma) rho_lambda(sigma) for every lambda (fft.py:115-151
 * `fourier_transform{,2}`) as one dense fp64 GEMM [batch, n!] x [n!, n!] on the DMMA pipe; the
 * [n!, n!] matrix of all rho_lambda(sigma) is built on the device on first use.  n <= 8 (13 GB at n = 8).
 * f_dev: lexicographic order. */
int snfft_dft_direct(snfft_plan* plan, const double* f_dev, double* fhat_dev, int64_t batch,
                     void* stream);
/* Frees the matrix that snfft_dft_direct cached in the plan (synchronises the device first). */
int snfft_dft_release(snfft_plan* plan);

/* ---- Sharded transform (multi-GPU) ----------------------------------------------
 * The coset recursion of fft.py:96-98 splits S_n into n!/m! blocks of S_m (coset-major order puts
 * them contiguously).  Rank r owns a contiguous range of blocks and runs levels 2..m on them with
 * the ordinary plan of S_m.  Columns of f_hat are independent (the level step multiplies from the
 * left), so ONE all-to-all then gives every rank all blocks but only "its" columns: for every
 * nu |- m a contiguous range of the d_nu columns, and for larger shapes the union of the ranges of
 * their branches.  Levels m+1..n run on these compact column shards without further
 * communication; rank r ends with f_hat(lambda)[:, owned columns] for every lambda.  This replaces
 * the gather-then-sum of cube_ft.py:102-106 (a reduce of full-size partial matrices) by an
 * exchange that moves world-times fewer bytes.
 * Compact level-k spectrum of rank r: for lambda in partitions(k) order, row-major
 * d_lambda x w_r(lambda). */
typedef struct snfft_shard snfft_shard;
int snfft_shard_create(const snfft_plan* plan, int rank, int world, int m, snfft_shard** out);
int snfft_shard_destroy(snfft_shard* shard);
/* S_m blocks [*begin, *end) owned by rank r (of n!/m! blocks in coset-major order). */
int snfft_shard_blocks(const snfft_shard* shard, int r, int64_t* begin, int64_t* end);
/* Number of doubles of one compact level-k spectrum on rank r (m <= k <= n). */
int snfft_shard_size(const snfft_shard* shard, int r, int k, int64_t* size);
/* idx int64[size(r, m)]: positions inside a packed S_m spectrum of the entries that rank r needs
 * (its columns of every nu |- m), in compact order: the send buffer for rank r is block[idx]. */
int snfft_shard_gather_index(const snfft_shard* shard, int r, int64_t* idx);
/* Columns of f_hat(irrep of S_k) owned by rank r, in compact order.  cols may be NULL. */
int snfft_shard_columns(const snfft_shard* shard, int r, int k, int irrep, int32_t* cols, int32_t* count);
/* One level k (m < k <= n) on the calling rank's shard: in double[groups][k][size(k-1)] ->
 * out double[groups][size(k)] (inverse != 0: the other way). */
int snfft_shard_level(const snfft_shard* shard, int k, const double* in_dev, double* out_dev,
                      int64_t groups, int inverse, void* stream);
/* Same level with explicit roles: x_dev holds the k sub-spectra per group (size k * size(k-1)), s_dev
 * the spectra (size(k)).  Forward reads x_dev and MAY OVERWRITE it with intermediate rows (levels
 * >= 9 run as in-place row-group phases), writing s_dev; inverse reads s_dev and writes x_dev. */
int snfft_shard_level_scratch(const snfft_shard* shard, int k, double* x_dev, double* s_dev,
                              int64_t groups, int inverse, void* stream);
/* Host export of the shard tables of one panel of level k (CPU tests): owned width, per-slot
 * destination / input offsets in the compact layouts, and the group strides.  The blocks and
 * coefficients are those of snfft_export_panel. */
int snfft_shard_export_panel(const snfft_shard* shard, int k, int panel, int32_t* width, int64_t* dest,
                             int64_t* rowbase, int64_t* in_stride, int64_t* out_stride);

/* The exchange over peer memory (NVLink P2P stores / loads inside one kernel; no packing pass, no
 * NCCL).  peer_bufs = host array of `world` device pointers that this process can dereference
 * (CUDA IPC / symmetric memory; entry `rank` is the caller's own buffer): rank r's column-shard
 * buffer [all blocks][size(r, m)].  local_dev = this rank's packed S_m spectra [its blocks][m!].
 *   inverse = 0 (push): after levels 2..m, every rank writes the columns owned by rank r into
 *                       rank r's column-shard buffer (contiguous remote runs of size(r, m) doubles);
 *   inverse = 1 (pull): after the inverse levels n..m+1, every rank reads its blocks out of every
 *                       rank's column-shard buffer into its packed spectra.
 * The caller orders the ranks around the call (push: the destination buffers are free before and
 * complete only after every rank has finished; pull: every peer buffer complete before, none
 * reused until every rank is done).  Replaces the Gather + sum of cube_ft.py:102-106 like the
 * NCCL path. */
int snfft_shard_exchange(snfft_shard* shard, const void* const* peer_bufs, void* local_dev, int inverse,
                         void* stream);

/* The whole sharded transform of the calling rank in one call (SURVEY.md 8b `snfft_forward_sharded`):
 * levels 2..m on its blocks, the peer-memory push, levels m+1..n on its column shard - all launched
 * on `stream`, with device-side barriers between the ranks (flags in peer-accessible signal pads),
 * no host synchronisation and no allocation.  Replaces the MPI driver of cube_ft.py:82-106.
 * The workspace is owned by the caller and reused across calls:
 *   peer_cols[r] : rank r's column-shard buffer, peer-accessible (CUDA IPC / symmetric memory),
 *                  sizes[0] doubles on every rank;
 *   peer_pads[r] : rank r's signal pad, peer-accessible, sizes[3] 8-byte words, zeroed once
 *                  (by all ranks, before the first call of any rank);
 *   local_a/b    : sizes[1] doubles each (this rank's S_m spectra and scratch), 16-byte aligned;
 *   shard_a/b    : sizes[2] doubles each (intermediate levels of the column shard; may be NULL
 *                  when n = m + 1);
 *   status       : device int, zeroed once; set to 1 if a barrier gave up waiting for a peer (~2 s).
 * f_local_dev: this rank's slice of the coset-major function (snfft_shard_blocks x m!, preserved);
 * fhat_compact_dev: snfft_shard_size(rank, n) doubles, the compact spectrum of this rank.
 * Every rank must make the same sequence of sharded calls, and one shard object runs one sharded
 * call at a time (the barrier epochs are counted per shard). */
typedef struct snfft_shard_ws {
  void* peer_cols[16];
  void* peer_pads[16];
  void* local_a;
  void* local_b;
  void* shard_a;
  void* shard_b;
  int* status;
} snfft_shard_ws;
int snfft_shard_ws_sizes(const snfft_shard* shard, int64_t* sizes /* [4] */);
int snfft_forward_sharded(snfft_shard* shard, const double* f_local_dev, double* fhat_compact_dev,
                          const snfft_shard_ws* ws, void* stream);
/* The inverse (fhat_compact_dev is preserved): levels n..m+1 backwards on the column shard, the
 * peer-memory pull, levels m..2 backwards on this rank's blocks. */
int snfft_inverse_sharded(snfft_shard* shard, const double* fhat_compact_dev, double* f_local_dev,
                          const snfft_shard_ws* ws, void* stream);

/* ---- Wreath-product transform (C3 wr S_n: the 2x2 cube group for n = 8) as one object -----------
 * f_hat = sum_{(o, sigma)} f(sigma, o) rho(o, sigma) at the irrep (alpha, parts) (multi.py:25-46,
 * cube_ft.py:33-73 with the irrep matrices of wreath.py:281-343): rho = D(o) Ind_{S_alpha}^{S_n}
 * (kron_k rho_{parts_k}).  The plan builds on the device what the reference pickles per irrep: coset
 * representatives of the Young subgroup (coset_utils.py:72-100), the frequencies of the characters
 * (wreath.py:43-58, :123-143), kron_k rho_{parts_k}(h) for h in S_alpha (wreath.py:191-224) and the
 * index of sigma = t_i h t_j^-1 for every coset pair.  alpha: int32[3], a weak partition of n;
 * parts: int32[3][n], the three partitions zero-padded to n entries; all_orientations = 0: the
 * orientation axis holds the 3^(n-1) tuples with zero sum mod 3 in product order (utils.py:206-209:
 * index = base-3 number of the first n-1 letters), != 0: all 3^n tuples.  At most 3^8 orientations
 * (n <= 9 with zero sum, n <= 8 with all).
 * f_dev: double[n!][orientations], sigma in perm2.sn order (lexicographic).  Output: the complex
 * D x D matrix as two row-major planes (D = cosets x block).  Stage 2 (the sums over the Young
 * subgroup per coset pair) runs as one fused gather kernel for small blocks, through the S_a level
 * recursion of the Young-subgroup factors (the kernels of snfft_forward / snfft_inverse) otherwise;
 * gather + DMMA GEMM only when every factor is trivial.  A plan serialises its transforms
 * (one shared workspace).  dims int64[6] = {D, cosets N, block, |S_alpha|, n!, orientations}. */
typedef struct snfft_wreath snfft_wreath;
int snfft_wreath_create(int n, const int32_t* alpha, const int32_t* parts, int all_orientations, int device,
                        snfft_wreath** out);
int snfft_wreath_destroy(snfft_wreath* plan);
int snfft_wreath_dims(const snfft_wreath* plan, int64_t* dims);
int snfft_wreath_forward(snfft_wreath* plan, const double* f_dev, double* out_re_dev, double* out_im_dev,
                         void* stream);
/* D tr(rho(g)^H f_hat) / |G| at every group element (cube_ift.py:56-67, fft.py:158-172 divided by
 * the group order): out planes double[n!][orientations]; with out_im_dev == NULL the result is
 * written to out_re_dev as interleaved complex128 [n!][orientations] (16-byte aligned). */
int snfft_wreath_inverse(snfft_wreath* plan, const double* fhat_re_dev, const double* fhat_im_dev,
                         double* out_re_dev, double* out_im_dev, void* stream);
/* The same at `count` sampled group elements in ONE launch (cube_ift.py:56-67 per state, the
 * consumer-facing form of compute_policy.py:66-88): element j = (orientation with index
 * ori_index_dev[j] on the plan's orientation axis, permutation with lexicographic rank
 * sigma_rank_dev[j] in perm2.sn order); out planes double[count].  Elements out of range give NaN.
 * Reads the plan's tables only (no workspace: safe next to a running transform of the same plan). */
int snfft_wreath_inverse_at(snfft_wreath* plan, const double* fhat_re_dev, const double* fhat_im_dev,
                            const int64_t* sigma_rank_dev, const int32_t* ori_index_dev, int64_t count,
                            double* out_re_dev, double* out_im_dev, void* stream);

/* ---- Building blocks of the wreath-product transform (wreath.py / multi.py / cube_ft.py) -------
 * C[m,n] = A[m,k] B[k,n], fp64 row-major device matrices, on the DMMA tensor pipe
 * (mma.sync.m8n8k4.f64).  The direct sums sum_g f(g) rho(g) of multi.py:25-46 and the character
 * sums over orientations (wreath.py:43-58) are evaluated as such products. */
int snfft_dgemm(const double* a_dev, const double* b_dev, double* c_dev, int64_t m, int64_t n, int64_t k,
                void* stream);
/* Character sums over the orientations (wreath.py:43-58, :123-143) as a DFT over C_3^m: every row
 * of in_re [count][3^m] (orientation index = base-3 number of the tuple, first letter most
 * significant) is transformed with m radix-3 passes in shared memory and only the nfreq frequencies
 * listed in freq_dev (same indexing) are written: out[count][nfreq] = sum_o in[o] omega^{+k.o}.
 * inverse != 0 is the adjoint: in [count][nfreq] complex is scattered onto the frequency grid and
 * out [count][3^m] = sum_k in[k] omega^{-k.o} (no normalisation).  Complex data are two planes, or
 * interleaved complex128 (16-byte aligned) where the second pointer is NULL: out_im_dev == NULL
 * (either direction) writes complex128 to out_re_dev, in_im_dev == NULL with inverse != 0 reads
 * complex128 from in_re_dev (the forward input is real: in_im_dev is ignored, and two rows share
 * one complex transform).  m <= 8. */
int snfft_c3_dft(int m, const double* in_re_dev, const double* in_im_dev, const int32_t* freq_dev, int nfreq,
                 int64_t count, double* out_re_dev, double* out_im_dev, int inverse, void* stream);
/* dst[e][0..width) = src[idx[e]][0..width) for e < count: regroups S_n by cosets of a Young
 * subgroup (the double loop over coset representatives of wreath.py:304-313 as an index map). */
int snfft_gather(const double* src_dev, const int64_t* idx_dev, double* dst_dev, int64_t count, int width,
                 void* stream);

/* Sizes of the plan's tables, for DESIGN.md / tests: bytes on the device, and multiply-adds per
 * group element of level k (0 for k outside 2..n). */
int64_t snfft_plan_device_bytes(const snfft_plan* plan);
double snfft_level_macs_per_element(const snfft_plan* plan, int k);

/* Number of kernels this library has launched on the calling process since load (bench.py's
 * gpu_launches claim is read from here). */
int64_t snfft_launch_count(void);

/* Test hook: level 8 normally runs as generated register programs; on != 0 routes it through the
 * table-driven streaming kernel that serves k >= 9, so that tests can cover that kernel at a size
 * the oracle reaches. */
int snfft_debug_force_table_driven(int on);
/* A/B hook: the inverse fused kernel (levels 7..2) stages its input with ONE cp.async.bulk (1-D
 * TMA) + mbarrier per S_7 spectrum (default, on != 0) or with 2 520 16-byte cp.async (on = 0);
 * results are identical, the timing comparison (tools/ab_tma.py) is in profiles/README.md. */
int snfft_debug_fused_tma(int on);

/* Per-kernel timing for bench.py: when enabled, every launch of this library is bracketed by CUDA
 * events on its stream.  snfft_timing_read synchronises them, fills ms_by_kind[15] and
 * launches_by_kind[15] and resets the totals.  Kinds: 0 fused forward (levels 2..7), 1 fused
 * inverse, 2 level 8 forward (generated), 3 level 8 inverse, 4 table-driven level kernel,
 * 5 in-place stages 1..5 of a level >= 9, 6 finished-row move of those levels, 7 row-group phase
 * forward, 8 row-group phase inverse, 9 peer-memory exchange, 10 device-side barrier between
 * ranks (includes waiting for the slowest rank), 11 reorder, 12 yor, 13 dgemm, 14 other. */
#define SNFFT_NUM_KERNEL_KINDS 15
int snfft_timing_enable(int on);
int snfft_timing_read(double* ms_by_kind, int64_t* launches_by_kind);

/* Host-side export of one panel program of level k (CPU tests interpret it in numpy to check the
 * plan builder without a GPU).  panel = index of mu in partitions(k-1).  Buffers may be NULL to
 * query sizes: *n_blocks = (k-1)*d_mu blocks in stage-major order; block_m[b] in 2..5;
 * block_stage[b] in 1..k-1; block_slots int32[n_blocks][5] (-1 padded); block_final[b] = bit mask
 * of the outputs that are finished rows; coef_fwd / coef_inv double[n_blocks][25] (row-major
 * m x m: forward out x in, inverse in x out, all scalings folded in); dest int64[k*d_mu] = offset
 * of each slot's finished row in the k! output; rowbase int64[k*d_mu] = offset of each slot's
 * input row in the k! input. */
int snfft_export_panel(const snfft_plan* plan, int k, int panel, int32_t* n_blocks,
                       int32_t* block_m, int32_t* block_stage, int32_t* block_slots,
                       int32_t* block_final, double* coef_fwd, double* coef_inv, int64_t* dest,
                       int64_t* rowbase);

/* Host-side export of the device tables of one row-group phase of level k >= 9 (phase >= 1; phase 0
 * is stages 1..5, generated programs), so that CPU tests can interpret exactly what the phase
 * kernel executes.  sizes int32[8] = {panels, groups, 16-byte words of a blob array, rows of the
 * largest group, words of the largest blob, stages, first stage, phases of the level}; the other
 * buffers may be NULL to query sizes: panels int32[..][5] = {columns, tiles, first group, groups,
 * pack}; groups int64[..][4] = {first word of the blob, words, rows, rows that stay on the x side
 * (they come first)}; blob_fwd / blob_inv: the program blobs (16 x words bytes each).  A blob is
 * u32 xoff[rows], u32 dest[rows], padding to 16 bytes, then per block a 16-byte header {u16 byte
 * offset (row index * 256) of its rows [5], u16 m, u32 pair flag (1: the next block is of the same
 * stage and size, the kernel interleaves the two)} followed by the m x m matrix (doubles,
 * row-major, y = C x in place on the rows; padded to 16 bytes), terminated by a header with m = 0;
 * the inverse blob lists the blocks backwards with the inverse matrices. */
int snfft_export_phase(const snfft_plan* plan, int k, int phase, int32_t* sizes, int32_t* panels,
                       int64_t* groups, void* blob_fwd, void* blob_inv);
/* The coefficient pool of level k that the block / phase tables index (doubles). */
int snfft_export_coef(const snfft_plan* plan, int k, double* coef, int64_t* count);

#ifdef __cplusplus
}
#endif
#endif /* SNFFT_B200_H */
