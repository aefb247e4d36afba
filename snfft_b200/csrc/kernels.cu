// sm_100a kernels of the S_n transform.
//
// Every level of the recursion (fft.py:96-112) is executed as the staged block program built by
// plan.cpp: rows of a column block are combined by small dense blocks (2x2 .. 5x5), columns are
// a passive, contiguous dimension.  The drivers:
//   * fused_kernel         : levels 2..k0 (k0 <= 7) of one S_{k0} block entirely in shared memory
//                            (generated register programs, gen_small.cuh), one HBM read and one
//                            HBM write for k0-1 levels; the next block arrives by cp.async;
//   * level8_kernel        : level 8 as two generated register phases per (panel, group chunk);
//   * p1_inplace_kernel, p0_rows_kernel, phase_kernel : a level k >= 9 as in-place generated
//                            stages 1..5 plus interpreted row-group phases for the later stages;
//   * level_kernel         : any level, table-driven column tiles (snfft_level, tests);
//   * reorder_tiled_kernel / reorder_kernel : lexicographic <-> coset-major element order;
//   * shard_exchange_kernel: the one exchange of the multi-GPU transform over peer memory;
//   * yor_kernel, dgemm_kernel (DMMA), gather_kernel : irrep matrices, direct sums, wreath stages.
// fp64 has no tcgen05 kind, so the arithmetic is DFMA on the CUDA cores; the kernels are sized
// against HBM (16 B per element per pass) and shared-memory bandwidth.
#include "kernels.cuh"

#include "gen_small.cuh"

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <vector>

namespace snfft {

std::atomic<int64_t> g_launch_count{0};
std::atomic<bool> g_force_table_driven{false};
std::atomic<bool> g_fused_tma{true};  // measured 1.3 % faster than the LDGSTS stage-in (tools/ab_tma.py)

namespace {
struct TimedLaunch {
  int kind;
  cudaEvent_t start, stop;
};
std::atomic<bool> g_timing{false};
std::vector<TimedLaunch> g_timed;
std::vector<cudaEvent_t> g_event_pool;
std::mutex g_timing_mu;

cudaEvent_t get_event() {
  if (!g_event_pool.empty()) {
    cudaEvent_t e = g_event_pool.back();
    g_event_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

// brackets one launch with events when timing is on
struct LaunchScope {
  cudaStream_t stream;
  bool on;
  TimedLaunch t;
  LaunchScope(int kind, cudaStream_t s) : stream(s), on(g_timing.load(std::memory_order_relaxed)) {
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    if (on) {
      std::lock_guard<std::mutex> lock(g_timing_mu);
      t.kind = kind;
      t.start = get_event();
      t.stop = get_event();
      cudaEventRecord(t.start, stream);
    }
  }
  ~LaunchScope() {
    if (on) {
      cudaEventRecord(t.stop, stream);
      std::lock_guard<std::mutex> lock(g_timing_mu);
      g_timed.push_back(t);
    }
  }
};
}  // namespace

void timing_enable(bool on) { g_timing = on; }

void timing_read(double* ms_by_kind, int64_t* launches_by_kind) {
  std::lock_guard<std::mutex> lock(g_timing_mu);
  for (int i = 0; i < kNumKinds; ++i) {
    ms_by_kind[i] = 0.0;
    launches_by_kind[i] = 0;
  }
  for (TimedLaunch& t : g_timed) {
    cudaEventSynchronize(t.stop);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, t.start, t.stop) == cudaSuccess) {
      ms_by_kind[t.kind] += ms;
      launches_by_kind[t.kind] += 1;
    }
    g_event_pool.push_back(t.start);
    g_event_pool.push_back(t.stop);
  }
  g_timed.clear();
}

namespace {

constexpr int kThreads = 256;
#ifndef SNFFT_FUSED_THREADS
#define SNFFT_FUSED_THREADS 256
#endif
#ifndef SNFFT_FUSED_WAVES
#define SNFFT_FUSED_WAVES 2
#endif
#ifndef SNFFT_L8_SUPER
#define SNFFT_L8_SUPER 30
#endif
#ifndef SNFFT_L8_MINBLOCKS
#define SNFFT_L8_MINBLOCKS 2
#endif
#ifndef SNFFT_L8_THREADS
#define SNFFT_L8_THREADS 256
#endif
constexpr int kL8Threads = SNFFT_L8_THREADS;
constexpr int kFusedThreads = SNFFT_FUSED_THREADS;


// ---------------------------------------------------------------------------------------------
// table-driven block executor
// ---------------------------------------------------------------------------------------------
// floor(e / w) for 0 <= e < 2^20 with inv = 1/w (the +0.5 keeps the quotient off integer edges)
__device__ __forceinline__ int fast_div(int e, float inv) {
  return __float2int_rz(((float)e + 0.5f) * inv);
}

template <int M>
__device__ __forceinline__ void matvec(const double* __restrict__ cm, const double (&x)[kMaxBlock],
                                       double (&y)[kMaxBlock]) {
#pragma unroll
  for (int a = 0; a < M; ++a) {
    double acc = __ldg(cm + a * M) * x[0];
#pragma unroll
    for (int b = 1; b < M; ++b) acc = fma(__ldg(cm + a * M + b), x[b], acc);
    y[a] = acc;
  }
}

__device__ __forceinline__ void matvec_m(int m, const double* __restrict__ cm,
                                         const double (&x)[kMaxBlock], double (&y)[kMaxBlock]) {
  switch (m) {
    case 2: matvec<2>(cm, x, y); break;
    case 3: matvec<3>(cm, x, y); break;
    case 4: matvec<4>(cm, x, y); break;
    default: matvec<5>(cm, x, y); break;
  }
}

// ---------------------------------------------------------------------------------------------
// streaming level kernel (any k): column tiles of one panel staged through shared memory, stages
// in place, finished rows stored straight from registers.
// grid.x = tiles of the level, grid.y = group chunks (grid-stride over groups)
// ---------------------------------------------------------------------------------------------
template <bool INV>
__global__ void __launch_bounds__(kThreads)
level_kernel(LevelDev L, const Tile* __restrict__ tiles, const double* __restrict__ in,
             double* __restrict__ out, int64_t groups) {
  extern __shared__ double buf[];
  const Tile t = tiles[blockIdx.x];
  const PanelDev P = L.panels[t.panel];
  const int d = P.d, rows = P.rows, w = t.w, k = L.k;
  const uint32_t* __restrict__ dest = L.dest + P.slot_off;
  const uint32_t* __restrict__ rowbase = L.rowbase + P.slot_off;
  const uint32_t* __restrict__ first = INV ? dest : rowbase;  // where a slot's row is read from
  const int tile_elems = rows * w;
  const float inv_w = 1.0f / (float)w, inv_tile = 1.0f / (float)tile_elems, inv_d = 1.0f / (float)d;

  for (int64_t g0 = (int64_t)blockIdx.y * t.cap; g0 < groups; g0 += (int64_t)gridDim.y * t.cap) {
    const int ng = (int)min((int64_t)t.cap, groups - g0);
    // ---- stage the tile: buf[(gb * rows + slot) * w + c] -------------------------------------
    for (int e = threadIdx.x; e < ng * tile_elems; e += kThreads) {
      const int gb = fast_div(e, inv_tile);
      const int r = e - gb * tile_elems;
      const int s = fast_div(r, inv_w);
      const int c = r - s * w;
      buf[e] = in[(g0 + gb) * (INV ? L.out_stride : L.in_stride) + first[s] + t.c0 + c];
    }
    __syncthreads();
    // ---- stages ------------------------------------------------------------------------------
    for (int st = 0; st < k - 1; ++st) {
      const int stage = INV ? (k - 2 - st) : st;
      const Block* __restrict__ blocks = L.blocks + P.blk_off + stage * d;
      const int items = ng * d * w;
      for (int e = threadIdx.x; e < items; e += kThreads) {
        const int tq = fast_div(e, inv_w);
        const int c = e - tq * w;
        const int gb = fast_div(tq, inv_d);
        const int q = tq - gb * d;
        const Block b = blocks[q];
        double* __restrict__ base = buf + gb * tile_elems + c;
        double* __restrict__ gout = out + (g0 + gb) * (INV ? L.in_stride : L.out_stride) + t.c0 + c;
        double x[kMaxBlock], y[kMaxBlock];
#pragma unroll
        for (int i = 0; i < kMaxBlock; ++i)
          if (i < b.m) x[i] = base[(int)b.slot[i] * w];
        matvec_m(b.m, L.coef + (INV ? b.coef_i : b.coef_f), x, y);
        if (!INV) {
#pragma unroll
          for (int i = 0; i < kMaxBlock; ++i)
            if (i < b.m) {
              if ((b.final_mask >> i) & 1)
                gout[dest[b.slot[i]]] = y[i];
              else
                base[(int)b.slot[i] * w] = y[i];
            }
        } else {
          gout[b.xoff] = y[0];  // X_{j+1}[W] is finished
          if (b.xoff1 != 0xffffffffu)
            gout[b.xoff1] = y[1];  // stage 1: X_1[W] too
          else
            base[(int)b.slot[1] * w] = y[1];
#pragma unroll
          for (int i = 2; i < kMaxBlock; ++i)
            if (i < b.m) base[(int)b.slot[i] * w] = y[i];
        }
      }
      __syncthreads();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// fused low levels.  K0 = min(n, 7).  Levels 2..4 run in registers per S_4 block (24 values per
// thread), levels 5 and 6 as generated whole-column register programs with one shared-memory
// exchange each, level 7 as the table-driven executor in shared memory with finished rows stored
// straight to global memory.  Shared-memory layouts are padded (gen::kS4/kS5/kP6/kS6) so that
// consecutive sub-groups hit distinct banks.
// ---------------------------------------------------------------------------------------------
// Shared memory of the fused kernel: P holds padded S_4 / S_6 blocks, Q holds padded S_5 blocks,
// then the C_5 rows of level 7 (3600) followed by the staged packed S_7 spectrum (5040).
constexpr int kFusedP = 5256;  // >= 210 * kS4
constexpr int kFusedQ = 8640;  // >= 3600 + 5040 and >= 42 * kS5 + 7
constexpr int kFusedC = 3600;  // C_5 region of level 7 at the start of Q

template <int K0>
struct FusedShape {
  static constexpr int kFact = K0 == 4 ? 24 : K0 == 5 ? 120 : K0 == 6 ? 720 : 5040;
  static constexpr int kS4Blocks = kFact / 24;      // S_4 blocks per group
  static constexpr int kGroups = 210 / kS4Blocks;   // groups per iteration (210 S_4 blocks)
};

// all instances of one generated level (5 or 6): warp-tasks = (panel, 32 instances), instances
// ordered sub-group fastest so that a warp touches consecutive sub-groups (distinct banks)
template <int LEVEL, bool INV, class AddrIn, class AddrOut>
__device__ __forceinline__ void run_generated_level(int nsub, AddrIn in_of, AddrOut out_of) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  constexpr int np = LEVEL == 5 ? gen::kL5Panels : gen::kL6Panels;
  int task = warp;  // round-robin over the flattened (panel, chunk) list
  int seen = 0;
#pragma unroll 1
  for (int p = 0; p < np; ++p) {
    const int dp = LEVEL == 5 ? gen::kL5Dims[p] : gen::kL6Dims[p];
    const int ninst = nsub * dp;
    const int chunks = (ninst + 31) >> 5;
    while (task < seen + chunks) {
      const int inst = (task - seen) * 32 + lane;
      if (inst < ninst) {
        const int c = inst / nsub;
        const int g = inst - c * nsub;
        if (LEVEL == 5)
          gen::l5_run<INV>(p, in_of(g) + c, out_of(g) + c);
        else
          gen::l6_run<INV>(p, in_of(g) + c, out_of(g) + c);
      }
      task += nwarps;
    }
    seen += chunks;
  }
}

// Round-robin distribution of 32-instance chunks of a list of segments over the warps of the CTA:
// body(segment, instance) runs with all lanes of a warp inside the same segment.
template <class Count, class Body>
__device__ __forceinline__ void for_each_chunk(int seg_begin, int seg_end, Count count, Body body) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  int task = warp, seen = 0;
#pragma unroll 1
  for (int sgm = seg_begin; sgm < seg_end; ++sgm) {
    const int n = count(sgm);
    const int chunks = (n + 31) >> 5;
    while (task < seen + chunks) {
      const int inst = (task - seen) * 32 + lane;
      if (inst < n) body(sgm, inst);
      task += nwarps;
    }
    seen += chunks;
  }
}

// Static variant for a full S_7 group (K0 == 7): the (segment, 32-instance chunk) list of a phase
// is a compile-time table ordered heaviest segment first, so a warp reads its tasks instead of
// scanning the segment list, and the second round of a phase holds the lightest chunks.
template <int MAXT>
struct ChunkTable {
  int n;
  unsigned char seg[MAXT];
  unsigned short inst0[MAXT];
  unsigned short count[MAXT];
};
template <int MAXT, class Count, class Cost>
constexpr ChunkTable<MAXT> make_chunk_table(int nseg, Count count, Cost cost) {
  ChunkTable<MAXT> T{};
  int weight[MAXT] = {};
  for (int sgm = 0; sgm < nseg; ++sgm)
    for (int i = 0; i < count(sgm); i += 32) {
      int at = T.n++;
      // insertion keeps the list sorted by cost, heaviest first (stable)
      while (at > 0 && weight[at - 1] < cost(sgm)) {
        T.seg[at] = T.seg[at - 1];
        T.inst0[at] = T.inst0[at - 1];
        T.count[at] = T.count[at - 1];
        weight[at] = weight[at - 1];
        --at;
      }
      T.seg[at] = (unsigned char)sgm;
      T.inst0[at] = (unsigned short)i;
      T.count[at] = (unsigned short)count(sgm);
      weight[at] = cost(sgm);
    }
  return T;
}
// (the tables are evaluated at compile time; reading the __device__ constexpr arrays there is fine)
#pragma nv_diag_suppress 20091
struct L5Count { constexpr int operator()(int p) const { return 42 * gen::kL5Dims[p]; } };
struct L5Cost { constexpr int operator()(int p) const { return gen::kL5Dims[p]; } };
struct L6Count { constexpr int operator()(int p) const { return 7 * gen::kL6Dims[p]; } };
struct L6Cost { constexpr int operator()(int p) const { return gen::kL6Dims[p]; } };
struct P1Count { constexpr int operator()(int s) const { return gen::kL7P1SigmaStart[s + 1] - gen::kL7P1SigmaStart[s]; } };
struct P1Cost { constexpr int operator()(int) const { return 1; } };
struct P2Count { constexpr int operator()(int p) const { return gen::kL7P2MaxInst[p]; } };
struct P2Cost { constexpr int operator()(int p) const { return gen::kL7PanelDim[p]; } };
#pragma nv_diag_default 20091
__device__ constexpr ChunkTable<24> kTabL5 = make_chunk_table<24>(gen::kL5Panels, L5Count{}, L5Cost{});
__device__ constexpr ChunkTable<16> kTabL6 = make_chunk_table<16>(gen::kL6Panels, L6Count{}, L6Cost{});
__device__ constexpr ChunkTable<16> kTabP1 = make_chunk_table<16>(7, P1Count{}, P1Cost{});
__device__ constexpr ChunkTable<16> kTabP2 = make_chunk_table<16>(gen::kL7NumPanels, P2Count{}, P2Cost{});

template <int MAXT, class Body>
__device__ __forceinline__ void for_each_static_chunk(const ChunkTable<MAXT>& T, Body body) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll 1
  for (int t = warp; t < T.n; t += kFusedThreads / 32) {
    const int inst = T.inst0[t] + lane;
    if (inst < T.count[t]) body((int)T.seg[t], inst);
  }
}

template <int LEVEL, bool INV, class AddrIn, class AddrOut>
__device__ __forceinline__ void run_generated_static(AddrIn in_of, AddrOut out_of) {
  constexpr int nsub = LEVEL == 5 ? 42 : 7;
  auto body = [&](int p, int inst) {
    const int c = inst / nsub;
    const int g = inst - c * nsub;
    if (LEVEL == 5)
      gen::l5_run<INV>(p, in_of(g) + c, out_of(g) + c);
    else
      gen::l6_run<INV>(p, in_of(g) + c, out_of(g) + c);
  };
  if (LEVEL == 5)
    for_each_static_chunk(kTabL5, body);
  else
    for_each_static_chunk(kTabL6, body);
}

// level 7 inside the fused kernel (one S_7 group): phase 1 turns the rows X_1..X_5 of every upper
// chain (bufA, padded S_6 blocks) into C_5 rows (bufB, kind layout, panel region at 5*in_off);
// phase 2 combines them with X_6, X_7 and stores the finished rows to global memory.
template <bool INV>
__device__ __forceinline__ void fused_level7_phase1(double* bufA, double* bufB) {
  for_each_static_chunk(
      kTabP1, [&](int sigma, int inst) {
        const int at = gen::kL7P1SigmaStart[sigma] + inst;
        const int e = gen::kL7P1InstEntry[at], c = gen::kL7P1InstCol[at];
        const int dmu = gen::kL7P1DMu[e], in_off = gen::kL7P1InOff[e], ubase = gen::kL7P1Base[e];
        double* xrows = bufA + in_off + ubase * dmu + c;          // rows (i, t): i * kS6 + t * dmu
        double* crows = bufB + 5 * in_off + 5 * ubase * dmu + c;  // rows p: p * dmu
        if (!INV)
          gen::p1_run<false, gen::kS6>(sigma, xrows, dmu, crows, dmu);
        else
          gen::p1_run<true, gen::kS6>(sigma, crows, dmu, xrows, dmu);
      });
}

template <bool INV>
__device__ __forceinline__ void fused_level7_phase2(double* bufA, double* bufB, double* gspec) {
  for_each_static_chunk(
      kTabP2, [&](int panel, int inst) { gen::l7p2_panel<INV>(panel, inst, bufB, bufA, gspec); });
}

#ifdef SNFFT_PHASE_CLOCKS
// debug build only: cycles between the barriers of the fused kernel, summed over CTAs (thread 0)
__device__ unsigned long long g_phase_cycles[32];
#define SNFFT_PH(i)                                                       \
  if (K0 == 7 && threadIdx.x == 0) {                                      \
    const long long now_ = clock64();                                     \
    atomicAdd(&g_phase_cycles[(INV ? 16 : 0) + (i)], (unsigned long long)(now_ - last_)); \
    last_ = now_;                                                         \
  }
#else
#define SNFFT_PH(i)
#endif

__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// Asynchronous stage-in of one S_7 group (K0 == 7) into Q + kFusedC (5040 doubles), issued while
// the previous group is still in its last phases: in the forward direction Q is free behind the
// C_5 rows once level 6 has read it, in the inverse once level 5 has.  Forward: the 16-byte chunks
// of S_4 block b are rotated by b, so that the threads of the S_4 stage (one block each, 12
// LDS.128) do not all hit the same two bank groups.
template <bool INV>
__device__ __forceinline__ void fused7_stage_async(double* Q, const double* gin) {
  constexpr int kPer = (2520 + kFusedThreads - 1) / kFusedThreads;
#pragma unroll
  for (int j = 0; j < kPer; ++j) {
    const int e = threadIdx.x + j * kFusedThreads;
    if (e < 2520) {
      int at = e;
      if (!INV) {
        const int b = e / 12, c = e - b * 12;
        int pc = c + b % 12;
        pc -= pc >= 12 ? 12 : 0;
        at = b * 12 + pc;
      }
      cp_async16(Q + kFusedC + 2 * at, gin + 2 * e);
    }
  }
}

// ---- TMA variant of the inverse stage-in (measured against the LDGSTS one, see profiles/README.md):
// the 40 320 bytes of an S_7 spectrum are one contiguous, 16-byte aligned run, so ONE thread can
// issue ONE bulk copy (cp.async.bulk, the 1-D form of TMA) that completes on an mbarrier, instead
// of 2 520 16-byte cp.async spread over the CTA.  (The forward direction rotates the 16-byte chunks
// of every S_4 block on the way in - see fused7_stage_async - which a bulk copy cannot do.)
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(double* smem_dst, const double* gsrc, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
      "r"(parity)
      : "memory");
}

template <int K0, bool INV, bool TMA = false>
__global__ void __launch_bounds__(kFusedThreads, 2)
fused_kernel(FusedDev F, const double* in, double* out, int64_t groups) {  // in may alias out
  extern __shared__ double sm[];
  using S = FusedShape<K0>;
  double* P = sm;
  double* Q = sm + kFusedP;
  const int tid = threadIdx.x, nthreads = blockDim.x;
  __shared__ unsigned long long stage_bar;
  unsigned stage_parity = 0;
  constexpr bool kBulk = TMA && INV && K0 == 7;
  if (kBulk) {
    if (tid == 0) mbar_init(&stage_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
  }
  auto stage_in = [&](const double* gsrc) {
    if (kBulk) {
      if (tid == 0) {
        // the generic-proxy reads of this region by earlier phases are ordered by the barrier before
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&stage_bar, 5040 * 8);
        bulk_g2s(Q + kFusedC, gsrc, 5040 * 8, &stage_bar);
      }
    } else {
      fused7_stage_async<INV>(Q, gsrc);
    }
  };
  auto stage_wait = [&]() {
    if (kBulk) {
      mbar_wait(&stage_bar, stage_parity);
      stage_parity ^= 1u;
    } else {
      cp_async_wait_all();
    }
  };

  if (K0 == 7 && (int64_t)blockIdx.x < groups) {
    stage_in(in + (int64_t)blockIdx.x * S::kFact);
    stage_wait();
    __syncthreads();
  }
  for (int64_t g0 = (int64_t)blockIdx.x * S::kGroups; g0 < groups; g0 += (int64_t)gridDim.x * S::kGroups) {
    const int ng = (int)min((int64_t)S::kGroups, groups - g0);
    const int64_t gnext = g0 + (int64_t)gridDim.x * S::kGroups;  // this CTA's next chunk
#ifdef SNFFT_PHASE_CLOCKS
    long long last_ = clock64();
#endif
    const double* gin = in + g0 * S::kFact;
    double* gout = out + g0 * S::kFact;
    const int n4 = ng * S::kS4Blocks;  // S_4 blocks in this chunk (<= 210)
    const int nelem = n4 * 24;
#ifndef SNFFT_NO_L2_PREFETCH
    if (K0 == 7 && tid == 0) {
      // the chunk after the next one is pulled into L2 (the next one is already being staged)
      const int64_t gn = gnext + (int64_t)gridDim.x * S::kGroups;
      if (gn < groups)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(in + gn * S::kFact),
                     "r"(5040 * 8)
                     : "memory");
    }
#endif
    if (!INV) {
      // ---- coalesced stage-in: S_4 blocks padded to stride kS4 (K0 == 7: staged asynchronously) --
      if (K0 != 7) {
        for (int e = tid; e < nelem; e += nthreads) {
          const int b = e / 24;
          P[b * gen::kS4 + (e - b * 24)] = gin[e];
        }
        __syncthreads();
      }
      SNFFT_PH(0)
      // ---- levels 2..4 in registers, in place ---------------------------------------------------
      if (tid < n4) {
        double r[24];
        double* blk = P + tid * gen::kS4;
        if (K0 == 7) {
          const double2* st = reinterpret_cast<const double2*>(Q + kFusedC) + tid * 12;
          const int rot = tid % 12;
#pragma unroll
          for (int c = 0; c < 12; ++c) {
            int pc = c + rot;
            pc -= pc >= 12 ? 12 : 0;
            const double2 v = st[pc];
            r[2 * c] = v.x;
            r[2 * c + 1] = v.y;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 24; ++i) r[i] = blk[i];
        }
        gen::s4_fwd(r);
#pragma unroll
        for (int i = 0; i < 24; ++i) blk[i] = r[i];
      }
      __syncthreads();
      SNFFT_PH(1)
      if (K0 == 4) {
        for (int e = tid; e < nelem; e += nthreads) {
          const int b = e / 24;
          gout[e] = P[b * gen::kS4 + (e - b * 24)];
        }
      }
      if (K0 >= 5) {
        const int nsub = n4 / 5;
        if (K0 == 5)
          run_generated_level<5, false>(nsub, [&](int g) { return (const double*)P + g * gen::kG5; },
                                        [&](int g) { return gout + (int64_t)g * 120; });
        else if (K0 == 7)
          run_generated_static<5, false>([&](int g) { return (const double*)P + g * gen::kG5; },
                                         [&](int g) { return Q + g * gen::kS5 + (g / 6) * gen::kP6; });
        else
          run_generated_level<5, false>(nsub, [&](int g) { return (const double*)P + g * gen::kG5; },
                                        [&](int g) { return Q + g * gen::kS5 + (g / 6) * gen::kP6; });
      }
      if (K0 >= 6) {
        __syncthreads();
        SNFFT_PH(2)
        const int nsub = n4 / 30;
        if (K0 == 6)
          run_generated_level<6, false>(nsub, [&](int g) { return (const double*)Q + g * gen::kG6; },
                                        [&](int g) { return gout + (int64_t)g * 720; });
        else
          run_generated_static<6, false>([&](int g) { return (const double*)Q + g * gen::kG6; },
                                         [&](int g) { return P + g * gen::kS6; });
      }
      if (K0 == 7) {
        // ---- level 7 as two register phases ---------------------------------------------------
        __syncthreads();
        SNFFT_PH(3)
        if (gnext < groups) stage_in(in + gnext * S::kFact);
        fused_level7_phase1<false>(P, Q);
        __syncthreads();
        SNFFT_PH(4)
        // finished rows go straight to global memory in runs of d_mu doubles (measured 5 % faster
        // than staging the spectrum in shared memory first; the inverse does stage its input)
        fused_level7_phase2<false>(P, Q, gout);
      }
    } else {
      if (K0 == 7) {
        SNFFT_PH(5)
        fused_level7_phase2<true>(P, Q, Q + kFusedC);
        __syncthreads();
        SNFFT_PH(6)
        fused_level7_phase1<true>(P, Q);
        __syncthreads();
        SNFFT_PH(7)
      }
      if (K0 >= 6) {
        const int nsub = n4 / 30;
        if (K0 == 6)
          run_generated_level<6, true>(nsub, [&](int g) { return gin + (int64_t)g * 720; },
                                       [&](int g) { return Q + g * gen::kG6; });
        else
          run_generated_static<6, true>([&](int g) { return (const double*)P + g * gen::kS6; },
                                        [&](int g) { return Q + g * gen::kG6; });
        __syncthreads();
        SNFFT_PH(8)
      }
      if (K0 >= 5) {
        const int nsub = n4 / 5;
        if (K0 == 5)
          run_generated_level<5, true>(nsub, [&](int g) { return gin + (int64_t)g * 120; },
                                       [&](int g) { return P + g * gen::kG5; });
        else if (K0 == 7)
          run_generated_static<5, true>(
              [&](int g) { return (const double*)Q + g * gen::kS5 + (g / 6) * gen::kP6; },
              [&](int g) { return P + g * gen::kG5; });
        else
          run_generated_level<5, true>(
              nsub, [&](int g) { return (const double*)Q + g * gen::kS5 + (g / 6) * gen::kP6; },
              [&](int g) { return P + g * gen::kG5; });
        __syncthreads();
        SNFFT_PH(9)
        if (K0 == 7 && gnext < groups) stage_in(in + gnext * S::kFact);
      }
      if (K0 == 4) {
        for (int e = tid; e < nelem; e += nthreads) {
          const int b = e / 24;
          P[b * gen::kS4 + (e - b * 24)] = gin[e];
        }
        __syncthreads();
        SNFFT_PH(10)
      }
      if (tid < n4) {
        double r[24];
        double* blk = P + tid * gen::kS4;
#pragma unroll
        for (int i = 0; i < 24; ++i) r[i] = blk[i];
        gen::s4_inv(r);
#pragma unroll
        for (int i = 0; i < 24; ++i) blk[i] = r[i];
      }
      __syncthreads();
      SNFFT_PH(11)
      for (int e = tid; e < nelem; e += nthreads) {
        const int b = e / 24;
        gout[e] = P[b * gen::kS4 + (e - b * 24)];
      }
    }
    if (K0 == 7 && (!kBulk || gnext < groups)) stage_wait();  // the next chunk's input has landed
    __syncthreads();                                          // buffers are reused by the next chunk
    SNFFT_PH(12)
  }
}

// ---------------------------------------------------------------------------------------------
// level 8 as two register phases (generated programs).  One CTA = one panel mu |- 7 of `cap`
// consecutive groups: phase 1 reads the rows X_1..X_5 of every upper chain straight from global
// memory and leaves the C_5 rows in shared memory ([group][5 d rows][d columns]); phase 2 reads
// them together with X_6..X_8 from global memory and stores the finished rows.  The inverse runs
// phase 2 transposed first.
// ---------------------------------------------------------------------------------------------
// CTAs are numbered super-chunk-major (kL8Super consecutive groups, all panels, heaviest panel
// first): the d_mu-double row segments that neighbouring panels write into (read from) the same
// 32-byte sectors then meet in L2 instead of costing a DRAM fill each.
constexpr int kL8Super = SNFFT_L8_SUPER;
struct L8Launch {
  int first_cta[gen::kL8NumPanels + 1];  // within a super-chunk: CTAs [first_cta[p], first_cta[p+1]) -> order[p]
  int order[gen::kL8NumPanels];          // panels, heaviest first
  int cap[gen::kL8NumPanels];            // groups per CTA, indexed like `order`
};

template <bool INV>
__global__ void __launch_bounds__(kL8Threads, SNFFT_L8_MINBLOCKS)
level8_kernel(L8Launch launch, const double* __restrict__ in, double* __restrict__ out, int64_t groups) {
  extern __shared__ double cbuf[];
  const int per_super = launch.first_cta[gen::kL8NumPanels];
  const int super = blockIdx.x / per_super, within = blockIdx.x - super * per_super;
  int slot = 0;
  while (slot + 1 < gen::kL8NumPanels && within >= launch.first_cta[slot + 1]) ++slot;
  const int panel = launch.order[slot], cap = launch.cap[slot];
  const int d = gen::kL8PanelDim[panel], in_off = gen::kL8PanelInOff[panel];
  const int64_t g0 = (int64_t)super * kL8Super + (int64_t)(within - launch.first_cta[slot]) * cap;
  const int64_t gend = min(groups, (int64_t)(super + 1) * kL8Super);
  if (g0 >= gend) return;
  const int ng = (int)min((int64_t)cap, gend - g0);
  const int cgroup = 5 * d * d;  // C_5 rows of one group
  // forward: in = 8 packed S_7 spectra per group, out = packed S_8 spectrum; inverse: swapped
  const double* xin = in + g0 * 40320;
  double* xout = out + g0 * 40320;

  auto phase1 = [&]() {
    for_each_chunk(
        gen::kL8P1RunFirst[panel], gen::kL8P1RunFirst[panel + 1],
        [&](int run) { return ng * (gen::kL8P1RunStart[run + 1] - gen::kL8P1RunStart[run]) * d; },
        [&](int run, int inst) {
          const int e0 = gen::kL8P1RunStart[run], ne = gen::kL8P1RunStart[run + 1] - e0;
          const int t = inst / d, c = inst - t * d;
          const int gb = t / ne, e = e0 + (t - gb * ne);
          const int ubase = gen::kL8P1Base[e];
          double* crows = cbuf + gb * cgroup + 5 * ubase * d + c;
          const int64_t xoff = (int64_t)gb * 40320 + in_off + ubase * d + c;
          if (!INV)
            gen::p1_run<false, 5040>(gen::kL8P1Sigma[e], xin + xoff, d, crows, d);
          else
            gen::p1_run<true, 5040>(gen::kL8P1Sigma[e], crows, d, xout + xoff, d);
        });
  };
  auto phase2 = [&]() {
    for_each_chunk(
        gen::kL8P2First[panel], gen::kL8P2First[panel + 1],
        [&](int pid) { return ng * gen::kL8P2DRho[pid] * d; },
        [&](int pid, int inst) {
          const int t = inst / d, c = inst - t * d;
          const int drho = gen::kL8P2DRho[pid];
          const int gb = t / drho, shift = t - gb * drho;
          double* crows = cbuf + gb * cgroup + shift * d + c;
          const int64_t goff = (int64_t)gb * 40320;
          if (!INV)
            gen::l8p2_run<false>(pid, crows, const_cast<double*>(xin) + goff + in_off + shift * d + c,
                                 xout + goff + c, shift);
          else
            gen::l8p2_run<true>(pid, crows, xout + goff + in_off + shift * d + c,
                                const_cast<double*>(xin) + goff + c, shift);
        });
  };
  if (!INV) {
#ifndef SNFFT_NO_L8_PREFETCH
    // X_6..X_8 (three runs of d*d doubles per group) are only read in phase 2: pull them into L2 now
    if ((int)threadIdx.x < 3 * ng) {
      const int gb = threadIdx.x / 3, i = 5 + (int)threadIdx.x - 3 * gb;
      const uintptr_t a = reinterpret_cast<uintptr_t>(xin + (int64_t)gb * 40320 + i * 5040 + in_off);
      const uintptr_t a0 = a & ~(uintptr_t)15;
      const unsigned bytes = (unsigned)(((a + (uintptr_t)d * d * 8 + 15) & ~(uintptr_t)15) - a0);
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0), "r"(bytes) : "memory");
    }
#endif
    phase1();
    __syncthreads();
    phase2();
  } else {
    phase2();
    __syncthreads();
    phase1();
  }
}

// ---------------------------------------------------------------------------------------------
// row-group phase kernel (levels >= 9).  One warp = one (row group, segment of its column chunks):
// the warp copies the group's program blob (row table + block stream with inline coefficients,
// laid out by the host) into its private shared memory once, then for every chunk of 32 columns
// (lanes = columns; narrow panels put several members of the group loop side by side) stages the
// group's rows with cp.async, interprets the stream on its lane-private column - one 16-byte
// broadcast read per block header, the matrix from shared memory, no global table reads, no
// barriers - and stores the rows back in place or, when finished, to the spectrum.
// ---------------------------------------------------------------------------------------------
constexpr int kPhaseMaxPanels = 64;  // >= number of partitions of 11 (56)
struct PhaseLaunch {
  uint32_t work_off[kPhaseMaxPanels + 1];  // prefix sums of groups * segments over the panels
  uint32_t seg_len[kPhaseMaxPanels];       // chunks per segment of a panel
  uint32_t nchunks[kPhaseMaxPanels];       // tiles * ceil(lgroups / pack) of a panel
};

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async8_s(unsigned smem_addr, const double* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16_s(unsigned smem_addr, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gsrc) : "memory");
}

__device__ __forceinline__ double lds_f64(unsigned addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f64(unsigned addr, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v));
}
__device__ __forceinline__ U128 lds_u128(unsigned addr) {
  U128 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void lds_f64x2(unsigned addr, double& a, double& b) {
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
}

// one block of the stream: y = C x on the m rows named by the header, in place in the lane's A
// columns.  pc = shared-memory address of the header, col = shared-memory address of the lane's
// first column; row r of the tile starts at r * 256 * A bytes, the lane's columns are 256 bytes apart.
template <int M, int A>
__device__ __forceinline__ void stream_block(unsigned pc, const U128 h, unsigned col) {
  double cm[M * M + 1];
#pragma unroll
  for (int i = 0; i < (M * M + 1) / 2; ++i) {
    double a, b;
    lds_f64x2(pc + 16u + 16u * i, a, b);
    cm[2 * i] = a;
    if (2 * i + 1 < M * M + 1) cm[2 * i + 1] = b;
  }
  unsigned row[5];
  row[0] = col + (h.x & 0xffffu) * A;
  row[1] = col + (h.x >> 16) * A;
  row[2] = col + (h.y & 0xffffu) * A;
  row[3] = col + (h.y >> 16) * A;
  row[4] = col + (h.z & 0xffffu) * A;
  double x[A][M], y[A][M];
#pragma unroll
  for (int q = 0; q < A; ++q)
#pragma unroll
    for (int b = 0; b < M; ++b) x[q][b] = lds_f64(row[b] + 256u * q);
#pragma unroll
  for (int q = 0; q < A; ++q)
#pragma unroll
    for (int a = 0; a < M; ++a) {
      double acc = cm[a * M] * x[q][0];
#pragma unroll
      for (int b = 1; b < M; ++b) acc = fma(cm[a * M + b], x[q][b], acc);
      y[q][a] = acc;
    }
#pragma unroll
  for (int q = 0; q < A; ++q)
#pragma unroll
    for (int a = 0; a < M; ++a) sts_f64(row[a] + 256u * q, y[q][a]);
}

// two blocks of one stage (disjoint rows, equal size) interleaved: twice the independent work per
// warp between the shared-memory round trips (the host marks such pairs in the header, M <= 3)
template <int M, int A>
__device__ __forceinline__ void stream_block_pair(unsigned pc0, const U128 h0, unsigned pc1, const U128 h1, unsigned col) {
  double cm[2][M * M + 1];
#pragma unroll
  for (int i = 0; i < (M * M + 1) / 2; ++i) {
    double a, b;
    lds_f64x2(pc0 + 16u + 16u * i, a, b);
    cm[0][2 * i] = a;
    if (2 * i + 1 < M * M + 1) cm[0][2 * i + 1] = b;
    lds_f64x2(pc1 + 16u + 16u * i, a, b);
    cm[1][2 * i] = a;
    if (2 * i + 1 < M * M + 1) cm[1][2 * i + 1] = b;
  }
  unsigned row[2][3];
  row[0][0] = col + (h0.x & 0xffffu) * A;
  row[0][1] = col + (h0.x >> 16) * A;
  row[0][2] = col + (h0.y & 0xffffu) * A;
  row[1][0] = col + (h1.x & 0xffffu) * A;
  row[1][1] = col + (h1.x >> 16) * A;
  row[1][2] = col + (h1.y & 0xffffu) * A;
  double x[2][A][M], y[2][A][M];
#pragma unroll
  for (int p = 0; p < 2; ++p)
#pragma unroll
    for (int q = 0; q < A; ++q)
#pragma unroll
      for (int b = 0; b < M; ++b) x[p][q][b] = lds_f64(row[p][b] + 256u * q);
#pragma unroll
  for (int p = 0; p < 2; ++p)
#pragma unroll
    for (int q = 0; q < A; ++q)
#pragma unroll
      for (int a = 0; a < M; ++a) {
        double acc = cm[p][a * M] * x[p][q][0];
#pragma unroll
        for (int b = 1; b < M; ++b) acc = fma(cm[p][a * M + b], x[p][q][b], acc);
        y[p][q][a] = acc;
      }
#pragma unroll
  for (int p = 0; p < 2; ++p)
#pragma unroll
    for (int q = 0; q < A; ++q)
#pragma unroll
      for (int a = 0; a < M; ++a) sts_f64(row[p][a] + 256u * q, y[p][q][a]);
}

// A = chunks of 32 columns a warp runs side by side (one read of every header and matrix for both)
template <bool INV, int A>
__global__ void __launch_bounds__(512, 1)
phase_kernel(PhaseDev P, PhaseLaunch Lp, double* __restrict__ xside, double* __restrict__ sside,
             uint32_t lgroups, int warp_doubles) {
  extern __shared__ double sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  // the warp's private shared memory: [tile: max_rows x 32 A doubles][blob: max_blob x 16 bytes]
  double* tile = sm + (size_t)warp * warp_doubles;
  const unsigned tile_s = (unsigned)__cvta_generic_to_shared(tile + lane);  // the lane's first column
  const unsigned blob_s = (unsigned)__cvta_generic_to_shared(tile + P.max_rows * 32 * A);
  const U128* __restrict__ blobs = INV ? P.blob_i : P.blob_f;
  const uint32_t nwork = Lp.work_off[P.npanels];

  for (uint32_t w = blockIdx.x * nwarps + warp; w < nwork; w += gridDim.x * nwarps) {
    int lo = 0, hi = P.npanels - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (Lp.work_off[mid] <= w) lo = mid; else hi = mid - 1;
    }
    const PhasePanelDev pan = P.panels[lo];
    const uint32_t seg = Lp.seg_len[lo], nch = Lp.nchunks[lo];
    const uint32_t nseg = (nch + seg - 1) / seg;
    const uint32_t rem = w - Lp.work_off[lo];
    const uint32_t g = rem / nseg, sg = rem - g * nseg;  // consecutive warps share a group (and its blob in L2)
    const PhaseGroupDev G = P.groups[pan.grp_off + g];
    const int nrows = (int)(G.nrows & 0xffffu), nlow = (int)(G.nrows >> 16);
    // ---- the group's program: global -> the warp's shared memory, once per unit of work ---------
    __syncwarp();  // the previous unit's stream is no longer read
    const U128* __restrict__ gblob = blobs + G.blob_off;
    for (uint32_t i = lane; i < G.blob_len; i += 32) cp_async16_s(blob_s + i * 16u, gblob + i);
    cp_async_commit();
    // the row table stays in registers, two rows per lane (groups have <= 64 rows), and is handed
    // out with shuffles: no shared-memory read sits in front of the asynchronous copies
    const uint32_t* __restrict__ gtab = reinterpret_cast<const uint32_t*>(gblob);
    const uint32_t xoff0 = lane < nrows ? __ldg(gtab + lane) : 0u;
    const uint32_t xoff1 = lane + 32 < nrows ? __ldg(gtab + lane + 32) : 0u;
    const uint32_t dst0 = lane < nrows ? __ldg(gtab + nrows + lane) : 0u;
    const uint32_t dst1 = lane + 32 < nrows ? __ldg(gtab + nrows + lane + 32) : 0u;
    const unsigned stream = blob_s + 16u * (unsigned)((2 * nrows + 3) >> 2);
    // lane -> (member of the lg loop, column)
    int sub = 0, cg = lane;
    if (pan.pack > 1) {
      sub = (lane * pan.magic) >> 16;
      cg = lane - sub * pan.d;
    }
    const uint32_t ntl = (uint32_t)pan.tiles;
    const uint32_t c_end = min(nch, (sg + 1) * seg);
    for (uint32_t ch0 = sg * seg; ch0 < c_end; ch0 += A) {
      uintptr_t xb[A], sb[A];
      bool ok[A];
#pragma unroll
      for (int q = 0; q < A; ++q) {
        // the tile index is the fastest one: neighbouring tiles of a row share 32-byte sectors, and
        // they meet in L2 when one warp runs them back to back
        const uint32_t ch = ch0 + q;
        const uint32_t lgi = ch / ntl, tl = ch - lgi * ntl;
        int64_t lg;
        int c;
        if (pan.pack > 1) {
          lg = (int64_t)lgi * pan.pack + sub;
          ok[q] = ch < c_end && sub < pan.pack && lg < (int64_t)lgroups;
          c = cg;
        } else {
          lg = lgi;
          c = (int)tl * 32 + lane;
          ok[q] = ch < c_end && c < pan.d;
        }
        xb[q] = reinterpret_cast<uintptr_t>(xside + lg * P.in_stride + c);
        sb[q] = reinterpret_cast<uintptr_t>(sside + lg * P.out_stride + c);
      }
      // ---- rows: global -> lane-private shared columns, all in flight at once ---------------------
      auto load_rows = [&](int r0, int r1, uint32_t tab, const uintptr_t (&base)[A]) {
#pragma unroll 4
        for (int r = r0; r < r1; ++r) {
          const uint32_t off = __shfl_sync(0xffffffffu, tab, r & 31);
#pragma unroll
          for (int q = 0; q < A; ++q)
            if (ok[q])
              cp_async8_s(tile_s + (unsigned)r * (256u * A) + 256u * q,
                          reinterpret_cast<const double*>(base[q] + (uint64_t)off * 8u));
        }
      };
      const int nx = INV ? nlow : nrows;
      load_rows(0, min(nx, 32), xoff0, xb);
      load_rows(32, nx, xoff1, xb);
      if (INV) {
        load_rows(nlow, min(nrows, 32), dst0, sb);
        load_rows(max(nlow, 32), nrows, dst1, sb);
      }
      cp_async_commit();
      cp_async_wait<0>();  // a lane only reads the rows it copied itself ...
      __syncwarp();        // ... but the blob (first chunk) was copied by all lanes
      // ---- the stream: the next header is fetched before the current block runs -----------------
      unsigned pc = stream;
      U128 h = lds_u128(pc);
      while (true) {
        const uint32_t m = h.z >> 16;
        if (m == 0) break;
        const unsigned len = 16u + 16u * ((m * m + 1) >> 1);
        const unsigned next = pc + len;
        const U128 hn = lds_u128(next);
        if (h.w != 0 && A <= 2) {  // this block and the next one: same stage, same size (2 or 3), disjoint rows
          const unsigned next2 = next + len;
          const U128 hn2 = lds_u128(next2);
          if (m == 2)
            stream_block_pair<2, A <= 2 ? A : 1>(pc, h, next, hn, tile_s);
          else
            stream_block_pair<3, A <= 2 ? A : 1>(pc, h, next, hn, tile_s);
          pc = next2;
          h = hn2;
          continue;
        }
        switch (m) {
          case 2: stream_block<2, A>(pc, h, tile_s); break;
          case 3: stream_block<3, A>(pc, h, tile_s); break;
          case 4: stream_block<4, A>(pc, h, tile_s); break;
          default: stream_block<5, A>(pc, h, tile_s); break;
        }
        pc = next;
        h = hn;
      }
      // ---- rows back: in place, or to the spectrum when finished ----------------------------------
      auto store_rows = [&](int r0, int r1, uint32_t tab, const uintptr_t (&base)[A]) {
#pragma unroll 4
        for (int r = r0; r < r1; ++r) {
          const uint32_t off = __shfl_sync(0xffffffffu, tab, r & 31);
#pragma unroll
          for (int q = 0; q < A; ++q)
            if (ok[q])
              *reinterpret_cast<double*>(base[q] + (uint64_t)off * 8u) =
                  lds_f64(tile_s + (unsigned)r * (256u * A) + 256u * q);
        }
      };
      const int nxs = INV ? nrows : nlow;
      store_rows(0, min(nxs, 32), xoff0, xb);
      store_rows(32, nxs, xoff1, xb);
      if (!INV) {
        store_rows(nlow, min(nrows, 32), dst0, sb);
        store_rows(max(nlow, 32), nrows, dst1, sb);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// stages 1..5 of a large level (k >= 9) as generated in-place register programs: one warp = one
// (upper chain, 32-column tile), lanes = columns; plus the scaled row move of finished rows.
// ---------------------------------------------------------------------------------------------
template <bool INV>
__global__ void __launch_bounds__(256)
p1_inplace_kernel(P1Dev P, double* __restrict__ xside, int64_t lgroups) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int64_t item = (int64_t)blockIdx.x * nwarps + warp; item < P.nitems; item += (int64_t)gridDim.x * nwarps) {
    const P1Entry e = P.entries[P.item_entry[item]];  // entries are sorted by sigma: warps of a CTA share code
    const int c = (int)(item - e.item_off) * 32 + lane;
    if (c >= e.d) continue;
    for (int64_t lg = blockIdx.y; lg < lgroups; lg += gridDim.y)
      gen::p1ip_run<INV>(e.sigma, xside + lg * P.in_stride + e.base + c, P.rI, e.d);
  }
}

template <bool INV>
__global__ void __launch_bounds__(256)
p0_rows_kernel(P1Dev P, double* __restrict__ xside, double* __restrict__ sside, int64_t lgroups) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t item = (int64_t)blockIdx.x * 8 + warp; item < P.nrow_items; item += (int64_t)gridDim.x * 8) {
    const P0Row r = P.rows[P.item_row[item]];
    const int c = (int)(item - r.item_off) * 32 + lane;
    if (c >= r.d) continue;
    for (int64_t lg = blockIdx.y; lg < lgroups; lg += gridDim.y) {
      double* xs = xside + lg * P.in_stride + r.xoff + c;
      double* ss = sside + lg * P.out_stride + r.dest + c;
      if (!INV) *ss = r.scale_f * *xs; else *xs = r.scale_i * *ss;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// element order
// ---------------------------------------------------------------------------------------------
struct Facts {
  int64_t f[kMaxN + 1];
};

__device__ __forceinline__ int64_t coset_to_lex_dev(int n, int64_t flat, const Facts& F) {
  int p[kMaxN + 1];
#pragma unroll
  for (int x = 1; x <= kMaxN; ++x) p[x] = x;
  for (int k = n; k >= 2; --k) {
    const int i = (int)((flat / F.f[k - 1]) % k) + 1;
    const int first = p[i];
    for (int x = i; x < k; ++x) p[x] = p[x + 1];
    p[k] = first;
  }
  int64_t rank = 0;
  for (int a = 1; a <= n; ++a) {
    int smaller = 0;
    for (int b = a + 1; b <= n; ++b) smaller += p[b] < p[a];
    rank += smaller * F.f[n - a];
  }
  return rank;
}

__global__ void build_lex_table_kernel(int n, Facts F, uint32_t* __restrict__ table) {
  const int64_t total = F.f[n];
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x)
    table[t] = (uint32_t)coset_to_lex_dev(n, t, F);
}

template <bool TO_COSET>
__global__ void reorder_kernel(int n, Facts F, const uint32_t* __restrict__ table,
                               const double* __restrict__ src, double* __restrict__ dst,
                               int64_t batch) {
  const int64_t total = F.f[n];
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t lex = table ? (int64_t)table[t] : coset_to_lex_dev(n, t, F);
    for (int64_t b = blockIdx.y; b < batch; b += gridDim.y) {
      if (TO_COSET)
        dst[b * total + t] = src[b * total + lex];
      else
        dst[b * total + lex] = src[b * total + t];
    }
  }
}

// Tiled reorder for n >= kTileMinN.  With R = 5 top coset digits and S = 4 low ones the index map
// factors: the top digits fix the last R letters of sigma (an arrangement A of a value set V) and
// leave the first n-R letters in increasing order, so for fixed V and fixed middle digits
//   coset[top(A) + mid * S! + low] = lex[base(V, mid, low) + A],   A < R!, low < S!
// i.e. every tile is a (S! x R!) transpose between 960-byte lexicographic runs and 192-byte
// coset-major runs, both sector aligned (the element-wise gather above moves 16x the bytes).
constexpr int kTileR = 5, kTileRF = 120, kTileSF = 24, kTileMinN = 9;  // R = 5, S = 4
constexpr int kTilePitch = kTileSF + 1;
constexpr int kTileThreads = 256;

template <bool TO_COSET>
__global__ void __launch_bounds__(kTileThreads)
reorder_tiled_kernel(int n, Facts F, int nsubsets, int64_t nmid, int mids_per_cta,
                     const double* __restrict__ src, double* __restrict__ dst, int64_t batch) {
  __shared__ double tile[kTileRF * kTilePitch];
  __shared__ int64_t top_off[kTileRF];
  __shared__ int64_t lex_base[kTileSF];
  __shared__ int iota[kMaxN + 1];  // iota[x] = x-th smallest value outside V (1-based)
  const int tid = threadIdx.x;
  const int m = n - kTileR;  // letters permuted by the middle and low digits
  const int64_t chunks = (nmid + mids_per_cta - 1) / mids_per_cta;
  const int64_t total = F.f[n];
  for (int64_t work = blockIdx.x; work < (int64_t)nsubsets * chunks; work += gridDim.x) {
    const int vi = (int)(work / chunks);
    const int64_t mid0 = (work - (int64_t)vi * chunks) * mids_per_cta;
    __syncthreads();
    if (tid < kTileRF) {
      // unrank the subset (lexicographic combinadic), every thread for itself: R values
      int V[kTileR];
      {
        int rem = vi, x = 1;
        for (int j = 0; j < kTileR; ++j) {
          for (;; ++x) {
            // subsets that continue with x at position j: C(n - x, R - 1 - j)
            int c = 1;
            const int top = n - x, k = kTileR - 1 - j;
            for (int t = 1; t <= k; ++t) c = c * (top - k + t) / t;
            if (top < k) c = 0;
            if (rem < c) break;
            rem -= c;
          }
          V[j] = x++;
        }
      }
      if (tid == 0) {
        int j = 0, cnt = 0;
        for (int x = 1; x <= n; ++x) {
          if (j < kTileR && V[j] == x) { ++j; continue; }
          iota[++cnt] = x;
        }
      }
      // arrangement tid of V: suffix letters sigma(m+1..n) by Lehmer decoding of tid
      int suf[kTileR], pool[kTileR];
#pragma unroll
      for (int j = 0; j < kTileR; ++j) pool[j] = V[j];
      int a = tid;
      int fact = kTileRF;
#pragma unroll
      for (int j = 0; j < kTileR; ++j) {
        fact /= (kTileR - j);
        const int q = a / fact;
        a -= q * fact;
        suf[j] = pool[q];
        for (int t = q; t + 1 < kTileR - j; ++t) pool[t] = pool[t + 1];
      }
      // top digits: at level k the letter suf[k-m-1] sits at position value - #(extracted smaller)
      int64_t off = 0;
#pragma unroll
      for (int k = kTileR - 1; k >= 0; --k) {
        int pos = suf[k];
#pragma unroll
        for (int t = k + 1; t < kTileR; ++t) pos -= suf[t] < suf[k];
        off += (int64_t)(pos - 1) * F.f[m + k];
      }
      top_off[tid] = off;
    }
    for (int mi = 0; mi < mids_per_cta && mid0 + mi < nmid; ++mi) {
      const int64_t mid = mid0 + mi;
      __syncthreads();
      if (tid >= 128 && tid < 128 + kTileSF) {
        const int low = tid - 128;
        int q[kMaxN + 1];
#pragma unroll
        for (int x = 1; x <= kMaxN; ++x) q[x] = x;
        const int64_t flat = mid * kTileSF + low;
        for (int k = m; k >= 2; --k) {
          const int i = (int)((flat / F.f[k - 1]) % k) + 1;
          const int first = q[i];
          for (int x = i; x < k; ++x) q[x] = q[x + 1];
          q[k] = first;
        }
        int64_t base = 0;
        for (int a = 1; a <= m; ++a) {
          int smaller = iota[q[a]] - q[a];
          for (int b = a + 1; b <= m; ++b) smaller += q[b] < q[a];
          base += smaller * F.f[n - a];
        }
        lex_base[low] = base;
      }
      __syncthreads();
      const int64_t coset0 = mid * kTileSF;
      for (int64_t b = blockIdx.y; b < batch; b += gridDim.y) {
        const double* __restrict__ sb = src + b * total;
        double* __restrict__ db = dst + b * total;
        if (TO_COSET) {
          for (int e = tid; e < kTileRF * kTileSF; e += kTileThreads) {
            const int low = e / kTileRF, a = e - low * kTileRF;
            tile[a * kTilePitch + low] = sb[lex_base[low] + a];
          }
          __syncthreads();
          for (int e = tid; e < kTileRF * kTileSF; e += kTileThreads) {
            const int a = e / kTileSF, low = e - a * kTileSF;
            db[top_off[a] + coset0 + low] = tile[a * kTilePitch + low];
          }
        } else {
          for (int e = tid; e < kTileRF * kTileSF; e += kTileThreads) {
            const int a = e / kTileSF, low = e - a * kTileSF;
            tile[a * kTilePitch + low] = sb[top_off[a] + coset0 + low];
          }
          __syncthreads();
          for (int e = tid; e < kTileRF * kTileSF; e += kTileThreads) {
            const int low = e / kTileRF, a = e - low * kTileRF;
            db[lex_base[low] + a] = tile[a * kTilePitch + low];
          }
        }
        __syncthreads();
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// rho_lambda(sigma): apply the generator word to the identity, one CTA per permutation
// ---------------------------------------------------------------------------------------------
// SEMI: Young's seminormal form (yor.py:119-182): for a tableau pair (r, q), r < q, the generator
// block is [[a, 1 - a^2], [1, -a]] instead of the orthogonal [[a, b], [b, -a]], b = sqrt(1 - a^2).
template <bool SEMI>
__global__ void __launch_bounds__(kThreads)
yor_kernel(int n, int d, const int32_t* __restrict__ partner, const double* __restrict__ diag,
           const int32_t* __restrict__ perms, double* __restrict__ out) {
  __shared__ int word[kMaxN * kMaxN];
  __shared__ int wlen;
  const int64_t pid = blockIdx.x;
  double* __restrict__ M = out + pid * (int64_t)d * d;
  if (threadIdx.x == 0) {
    // bubble-sort word of the permutation map, as yor.py:19-38 does for a cycle
    int p[kMaxN];
    for (int i = 0; i < n; ++i) p[i] = perms[pid * n + i];
    int len = 0;
    for (int i = 0; i < n; ++i)
      for (int j = n - 1; j > i; --j)
        if (p[j] < p[j - 1]) {
          const int tmp = p[j];
          p[j] = p[j - 1];
          p[j - 1] = tmp;
          word[len++] = j;  // the transposition (j, j+1), 1-based
        }
    wlen = len;
  }
  for (int64_t e = threadIdx.x; e < (int64_t)d * d; e += kThreads) M[e] = (e / d == e % d) ? 1.0 : 0.0;
  __syncthreads();
  // rho(sigma) = rho(t_L) ... rho(t_1) read in the order the swaps were recorded reversed, i.e.
  // the product "reversed(factors)" left to right; applied to I from the rightmost factor, which
  // is the first recorded swap.
  for (int t = 0; t < wlen; ++t) {
    const int j = word[t];
    const int32_t* __restrict__ pr = partner + (int64_t)(j - 1) * d;
    const double* __restrict__ dg = diag + (int64_t)(j - 1) * d;
    for (int64_t e = threadIdx.x; e < (int64_t)d * d; e += kThreads) {
      const int r = (int)(e / d);
      const int c = (int)(e - (int64_t)r * d);
      const int q = pr[r];
      const double a = dg[r];
      if (q < 0) {
        M[e] = a * M[e];
      } else if (q > r) {
        const double b = SEMI ? 1 - a * a : sqrt(1 - a * a);
        const double x = M[(int64_t)r * d + c], y = M[(int64_t)q * d + c];
        M[(int64_t)r * d + c] = a * x + b * y;
        M[(int64_t)q * d + c] = (SEMI ? x : b * x) - a * y;
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// Stage 1 of the wreath-product transform (wreath.py:43-58, :123-143): the
// character sums over the orientations, F_i(sigma) = sum_o f(sigma, o) chi_i(o), are values of the
// DFT of f(sigma, .) over C_3^m at the frequencies of the alpha-orbit.  One CTA = one sigma: the
// 3^m orientation values (17.5 KB for the 2x2 cube) sit in shared memory as re / im planes, m
// radix-3 passes run in place, and only the N needed frequencies are written.  INV: the adjoint -
// the N values are scattered onto the frequency grid and the conjugate DFT gives all 3^m values.
// ---------------------------------------------------------------------------------------------
// D consecutive base-3 digits (strides s, 3s, .. of the grid) transformed in registers: a 3^D-point
// DFT per thread, D radix-3 stages on compile-time indices, one shared-memory round trip.
// OUT = 0: back to shared memory; 1: straight to the global output planes (the last pass of the
// adjoint); 2: the same as interleaved complex128 (one 16-byte store per point).  FROM_GLOBAL: the
// points come from two real input rows (ga -> real part, gb -> imaginary part or null: zero) - the
// first pass of the forward transform.
template <int D, bool INV, int OUT, bool FROM_GLOBAL = false>
__device__ __forceinline__ void c3_digits(double2* __restrict__ g, int base, int s,
                                          double* __restrict__ gre, double* __restrict__ gim,
                                          const double* __restrict__ ga = nullptr,
                                          const double* __restrict__ gb = nullptr) {
  constexpr int P = D == 1 ? 3 : (D == 2 ? 9 : 27);
  const double h = INV ? -0.8660254037844386467637231707529 : 0.8660254037844386467637231707529;  // +-sqrt(3)/2
  double xr[P], xi[P];
#pragma unroll
  for (int j = 0; j < P; ++j) {
    if (FROM_GLOBAL) {
      xr[j] = ga[base + j * s];
      xi[j] = gb ? gb[base + j * s] : 0.0;
    } else {
      const double2 v = g[base + j * s];
      xr[j] = v.x;
      xi[j] = v.y;
    }
  }
#pragma unroll
  for (int dg = 0; dg < D; ++dg) {
    const int st = dg == 0 ? 1 : (dg == 1 ? 3 : 9);
#pragma unroll
    for (int j = 0; j < P; ++j) {
      if ((j / st) % 3 != 0) continue;  // j = index with digit dg equal to 0
      const double x0r = xr[j], x0i = xi[j];
      const double x1r = xr[j + st], x1i = xi[j + st];
      const double x2r = xr[j + 2 * st], x2i = xi[j + 2 * st];
      const double sr = x1r + x2r, si = x1i + x2i;
      const double dr = (x1r - x2r) * h, di = (x1i - x2i) * h;
      xr[j] = x0r + sr;
      xi[j] = x0i + si;
      // X_1 = x0 - (x1 + x2)/2 + i (x1 - x2) sqrt(3)/2 ; X_2 = its mirror (signs flipped for the adjoint)
      xr[j + st] = x0r - 0.5 * sr - di;
      xi[j + st] = x0i - 0.5 * si + dr;
      xr[j + 2 * st] = x0r - 0.5 * sr + di;
      xi[j + 2 * st] = x0i - 0.5 * si - dr;
    }
  }
#pragma unroll
  for (int j = 0; j < P; ++j) {
    if (OUT == 2) {
      reinterpret_cast<double2*>(gre)[base + j * s] = make_double2(xr[j], xi[j]);
    } else if (OUT == 1) {
      gre[base + j * s] = xr[j];
      gim[base + j * s] = xi[j];
    } else {
      g[base + j * s] = make_double2(xr[j], xi[j]);
    }
  }
}

// threads per CTA, measured for the cube (3^7 grid, 243 radix-9 tasks per pass): the forward kernel is
// fastest with all tasks of a pass in one round (0.300 -> 0.267 ms), the adjoint with two rounds of 128
// (adjoint: 0.582 / 0.560 / 0.487 / 0.550 / 0.560 / 0.582 ms with 64 / 96 / 128 / 160 / 192 / 256 threads)
#ifndef SNFFT_C3_FWD_THREADS
#define SNFFT_C3_FWD_THREADS 256
#endif
#ifndef SNFFT_C3_ADJ_THREADS
#define SNFFT_C3_ADJ_THREADS 128
#endif
constexpr int kDftThreadsFwd = SNFFT_C3_FWD_THREADS, kDftThreadsAdj = SNFFT_C3_ADJ_THREADS;

// Forward: the input rows are REAL, so one complex transform serves TWO rows: z = f(s1,.) + i f(s2,.),
// Z = DFT(z), and F_1(k) = (Z(k) + conj Z(-k)) / 2, F_2(k) = (Z(k) - conj Z(-k)) / 2i.  One CTA =
// one pair of rows; half the arithmetic and half the shared-memory traffic per row.  The adjoint
// (complex in, complex out) has no such symmetry: one CTA = one row, the factor `scale` rides on
// the nfreq scattered values and the last pass stores to global memory from registers (lanes run
// along the low digits: coalesced).  INTERLEAVED: the output is complex128 in out_re instead of
// two planes; in_im == nullptr: the adjoint's input is complex128 in in_re.
//
// The grid lives in shared memory as interleaved complex (double2): every point is ONE 16-byte
// access, and the task starts of all passes (9 t, 81 hi + lo, ...) are distinct mod 8 within a
// quarter warp, so the accesses are conflict-free.  (Two planes of doubles: the stride-9 pass hit 9
// of 16 bank pairs per half warp; ncu counted 36 % / 29 % of the wavefronts of the forward / adjoint
// kernel as bank conflicts.)
template <bool INV, bool INTERLEAVED>
__global__ void __launch_bounds__(INV ? kDftThreadsAdj : kDftThreadsFwd)
c3_dft_kernel(int m, int size, int64_t count, const double* __restrict__ in_re, const double* __restrict__ in_im,
              const int32_t* __restrict__ freq, int nfreq, double* __restrict__ out_re,
              double* __restrict__ out_im, double scale) {
  extern __shared__ double2 grid[];
  constexpr int kDftThreads = INV ? kDftThreadsAdj : kDftThreadsFwd;
  const int64_t sig = INV ? (int64_t)blockIdx.x : 2 * (int64_t)blockIdx.x;
  const bool second = !INV && sig + 1 < count;
  int top = 0;  // digits already transformed on the way in (forward)
  if (!INV) {
    // the top one or two digits straight from global memory: lanes run along the low digits, so
    // the strided points of a thread are coalesced across the warp, and the grid is written once
    // instead of written and read back (0.353 -> 0.331 ms for the cube)
    top = m >= 2 ? 2 : 1;
    const int pts = top == 2 ? 9 : 3, stride = size / pts;
    const double* ga = in_re + sig * size;
    const double* gb = second ? in_re + (sig + 1) * size : nullptr;
    for (int t = threadIdx.x; t < stride; t += kDftThreads) {
      if (top == 2)
        c3_digits<2, INV, 0, true>(grid, t, stride, nullptr, nullptr, ga, gb);
      else
        c3_digits<1, INV, 0, true>(grid, t, stride, nullptr, nullptr, ga, gb);
    }
  } else {
    for (int e = threadIdx.x; e < size; e += kDftThreads) grid[e] = make_double2(0.0, 0.0);
    __syncthreads();
    for (int i = threadIdx.x; i < nfreq; i += kDftThreads) {
      double vr, vi;
      if (in_im) {
        vr = in_re[sig * nfreq + i];
        vi = in_im[sig * nfreq + i];
      } else {
        const double2 v = reinterpret_cast<const double2*>(in_re)[sig * nfreq + i];
        vr = v.x;
        vi = v.y;
      }
      grid[freq[i]] = make_double2(scale * vr, scale * vi);
    }
  }
  __syncthreads();
  // forward: omega^{+k o}; adjoint: omega^{-k o}.  Two digits per pass (9 points per thread in
  // registers: 243 independent sub-transforms per pass for the 2x2 cube, ~80 registers), then the
  // odd digit: 4 barriers instead of 7.  (Three digits per pass - 27 points, 168..190 registers,
  // 3 passes - measured slower: 0.63 vs 0.56 ms for the cube, too few warps per SM.)
  double* gre = INTERLEAVED ? out_re + 2 * sig * size : out_re + sig * size;
  double* gim = INTERLEAVED ? nullptr : out_im + sig * size;
  int done = 0, stride = 1;
  const int low = m - top;  // digits left for the passes over shared memory
  while (done < low) {
    const int dg = low - done >= 2 ? 2 : 1;
    const int pts = dg == 2 ? 9 : 3;
    const bool to_global = INV && done + dg == m;
    for (int t = threadIdx.x; t < size / pts; t += kDftThreads) {
      const int hi = t / stride, lo = t - hi * stride;
      const int base = hi * pts * stride + lo;
      if (dg == 2) {
        if (to_global)
          c3_digits<2, INV, INTERLEAVED ? 2 : 1>(grid, base, stride, gre, gim);
        else
          c3_digits<2, INV, 0>(grid, base, stride, nullptr, nullptr);
      } else {
        if (to_global)
          c3_digits<1, INV, INTERLEAVED ? 2 : 1>(grid, base, stride, gre, gim);
        else
          c3_digits<1, INV, 0>(grid, base, stride, nullptr, nullptr);
      }
    }
    __syncthreads();
    done += dg;
    stride *= pts;
  }
  if (!INV) {
    for (int i = threadIdx.x; i < nfreq; i += kDftThreads) {
      const int k = freq[i];
      int kn = 0;  // the index of -k: every base-3 digit negated
      for (int a = 0, r = k, p = 1; a < m; ++a, p *= 3) {
        const int d = r % 3;
        r /= 3;
        kn += (d == 0 ? 0 : 3 - d) * p;
      }
      const double2 z = grid[k], w = grid[kn];
      if (INTERLEAVED) {
        double2* o = reinterpret_cast<double2*>(out_re);
        o[sig * nfreq + i] = make_double2(0.5 * (z.x + w.x), 0.5 * (z.y - w.y));
        if (second) o[(sig + 1) * nfreq + i] = make_double2(0.5 * (z.y + w.y), 0.5 * (w.x - z.x));
      } else {
        out_re[sig * nfreq + i] = 0.5 * (z.x + w.x);
        out_im[sig * nfreq + i] = 0.5 * (z.y - w.y);
        if (second) {
          out_re[(sig + 1) * nfreq + i] = 0.5 * (z.y + w.y);
          out_im[(sig + 1) * nfreq + i] = 0.5 * (w.x - z.x);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Sampled inverse (tile/par_inv_ft.py:27-45, compute_policy.py:66-88): f(sigma_j) for a batch of
// permutations straight from the packed spectrum,
//   f(sigma) = (1/n!) sum_lambda d_lambda tr(rho_lambda(sigma)^T f_hat(lambda)),
// without materialising rho: rho(sigma)^T = rho(t_1) ... rho(t_L) applied to f_hat(lambda) from the
// left as 2-sparse row rotations (columns are independent), then the trace.  One CTA = one
// (sigma_j, irrep, tile of columns): the d x w tile sits in shared memory, the word is applied
// generator by generator, the partial trace is added to out[j].  All irreps in ONE launch.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
inverse_at_kernel(int n, const InvAtItem* __restrict__ items, const InvAtIrrep* __restrict__ irreps,
                  const double* __restrict__ fhat, const int32_t* __restrict__ perms,
                  double* __restrict__ out) {
  extern __shared__ double tile[];
  __shared__ int word[kMaxN * kMaxN];
  __shared__ int wlen;
  __shared__ double partial[kThreads / 32];
  const InvAtItem it = items[blockIdx.x];
  const InvAtIrrep R = irreps[it.irrep];
  const int64_t pid = blockIdx.y;
  const int d = R.d, w = it.w;
  if (threadIdx.x == 0) {
    // the same bubble-sort word as yor_kernel (yor.py:19-38)
    int p[kMaxN];
    for (int i = 0; i < n; ++i) p[i] = perms[pid * n + i];
    int len = 0;
    for (int i = 0; i < n; ++i)
      for (int j = n - 1; j > i; --j)
        if (p[j] < p[j - 1]) {
          const int tmp = p[j];
          p[j] = p[j - 1];
          p[j - 1] = tmp;
          word[len++] = j;
        }
    wlen = len;
  }
  const double* __restrict__ F = fhat + R.off + it.c0;
  for (int e = threadIdx.x; e < d * w; e += kThreads) {
    const int r = e / w, c = e - r * w;
    tile[e] = F[(int64_t)r * d + c];
  }
  __syncthreads();
  // yor_kernel builds rho(sigma) = G_L ... G_1 (G_t = generator of the t-th recorded swap, all
  // symmetric), so rho(sigma)^T f_hat = G_1 (G_2 ... (G_L f_hat)): the last recorded swap first
  for (int t = wlen - 1; t >= 0; --t) {
    const int j = word[t];
    const int32_t* __restrict__ pr = R.partner + (int64_t)(j - 1) * d;
    const double* __restrict__ dg = R.diag + (int64_t)(j - 1) * d;
    for (int e = threadIdx.x; e < d * w; e += kThreads) {
      const int r = e / w, c = e - r * w;
      const int q = pr[r];
      const double a = dg[r];
      if (q < 0) {
        tile[e] = a * tile[e];
      } else if (q > r) {
        const double b = sqrt(1 - a * a);
        const double x = tile[r * w + c], y = tile[q * w + c];
        tile[r * w + c] = a * x + b * y;
        tile[q * w + c] = b * x - a * y;
      }
    }
    __syncthreads();
  }
  double acc = 0.0;
  for (int c = threadIdx.x; c < w; c += kThreads) acc += tile[(it.c0 + c) * w + c];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) partial[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double sum = 0.0;
    for (int i = 0; i < kThreads / 32; ++i) sum += partial[i];
    atomicAdd(out + pid, R.weight * sum);
  }
}

// ---------------------------------------------------------------------------------------------
// fp64 GEMM on the DMMA pipe: C[M,N] = A[M,K] B[K,N], row-major.  Each warp owns a 16x16 tile
// of C as 2x2 mma.m8n8k4 fragments; a CTA of 8 warps covers 64x32... (simple, oracle-sized).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

constexpr int kGemmBM = 64, kGemmBN = 64, kGemmBK = 16;

// generic shapes (any M, N, K; edges handled per element): synchronous shared-memory tiles
__global__ void __launch_bounds__(256)
dgemm_kernel(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C,
             int M, int N, int K) {
  __shared__ double sA[kGemmBM][kGemmBK + 1];
  __shared__ double sB[kGemmBK][kGemmBN + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * kGemmBM, n0 = blockIdx.x * kGemmBN;
  // 8 warps: 4 along M (16 rows each) x 2 along N (32 cols each); per warp 2 x 4 fragments
  const int wm = (warp & 3) * 16, wn = (warp >> 2) * 32;
  double acc[2][4][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  const int grp = lane >> 2, tig = lane & 3;  // fragment coordinates of mma.m8n8k4
  for (int k0 = 0; k0 < K; k0 += kGemmBK) {
    for (int e = threadIdx.x; e < kGemmBM * kGemmBK; e += 256) {
      const int r = e / kGemmBK, c = e % kGemmBK;
      sA[r][c] = (m0 + r < M && k0 + c < K) ? A[(int64_t)(m0 + r) * K + k0 + c] : 0.0;
    }
    for (int e = threadIdx.x; e < kGemmBK * kGemmBN; e += 256) {
      const int r = e / kGemmBN, c = e % kGemmBN;
      sB[r][c] = (k0 + r < K && n0 + c < N) ? B[(int64_t)(k0 + r) * N + n0 + c] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kGemmBK; kk += 4) {
      double a[2], b[4];
#pragma unroll
      for (int i = 0; i < 2; ++i) a[i] = sA[wm + i * 8 + grp][kk + tig];  // A frag: row grp, col tig
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sB[kk + tig][wn + j * 8 + grp];  // B frag: row tig, col grp
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    __syncthreads();
  }
  // C frag: row grp, cols 2*tig, 2*tig+1
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = m0 + wm + i * 8 + grp;
      const int c = n0 + wn + j * 8 + 2 * tig;
      if (r < M) {
        if (c < N) C[(int64_t)r * N + c] = acc[i][j][0];
        if (c + 1 < N) C[(int64_t)r * N + c + 1] = acc[i][j][1];
      }
    }
}

// Large aligned shapes (K a multiple of 16, N even, 16-byte aligned operands): 128 x 128 x 16 tiles,
// three stages of 16-byte cp.async (zero fill past the edges), 8 warps of 64 x 32 (8 x 4 fragments
// of mma.m8n8k4: 32 DMMAs per 12 shared-memory loads).  Tile strides 20 / 132 doubles keep both
// fragment loads conflict-free per half warp.
constexpr int kG2BM = 128, kG2BN = 128, kG2BK = 16, kG2Stages = 3;
constexpr int kG2SA = kG2BK + 4, kG2SB = kG2BN + 4;  // padded strides (doubles)
constexpr int kG2StageDoubles = kG2BM * kG2SA + kG2BK * kG2SB;

__device__ __forceinline__ void cp_async16_zfill(unsigned smem_addr, const void* gsrc, bool valid) {
  const int bytes = valid ? 16 : 0;  // src-size 0: the 16 bytes are zero-filled, the source is not read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr), "l"(gsrc), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(256)
dgemm_pipe_kernel(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C,
                  int M, int N, int K) {
  extern __shared__ double gsm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * kG2BM, n0 = blockIdx.x * kG2BN;
  const int wm = (warp & 1) * 64, wn = (warp >> 1) * 32;  // 2 warps along M x 4 along N
  const int grp = lane >> 2, tig = lane & 3;
  const unsigned smem0 = (unsigned)__cvta_generic_to_shared(gsm);
  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  auto stage_in = [&](int stage, int k0) {
    const unsigned sa = smem0 + (unsigned)(stage * kG2StageDoubles) * 8u;
    const unsigned sb = sa + (unsigned)(kG2BM * kG2SA) * 8u;
    // A tile: 128 rows x 16 doubles = 8 chunks of 16 bytes per row
#pragma unroll
    for (int q = 0; q < (kG2BM * kG2BK / 2) / 256; ++q) {
      const int e = threadIdx.x + q * 256;
      const int r = e >> 3, c = (e & 7) * 2;
      const bool ok = m0 + r < M;  // K is a multiple of 16: no edge along k
      const double* src = A + (int64_t)(ok ? m0 + r : 0) * K + k0 + c;
      cp_async16_zfill(sa + (unsigned)(r * kG2SA + c) * 8u, src, ok);
    }
    // B tile: 16 rows x 128 doubles = 64 chunks per row
#pragma unroll
    for (int q = 0; q < (kG2BK * kG2BN / 2) / 256; ++q) {
      const int e = threadIdx.x + q * 256;
      const int r = e >> 6, c = (e & 63) * 2;
      const bool ok = n0 + c < N;  // N is even: a chunk is inside or outside as a whole
      const double* src = B + (int64_t)(k0 + r) * N + (ok ? n0 + c : 0);
      cp_async16_zfill(sb + (unsigned)(r * kG2SB + c) * 8u, src, ok);
    }
    cp_async_commit();
  };

  const int nk = K / kG2BK;
#pragma unroll
  for (int s = 0; s < kG2Stages - 1; ++s) {
    if (s < nk) stage_in(s, s * kG2BK); else cp_async_commit();
  }
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<kG2Stages - 2>();  // tile kt has landed (for this thread's copies) ...
    __syncthreads();                 // ... and for everybody's; the stage read last iteration is free
    if (kt + kG2Stages - 1 < nk) stage_in((kt + kG2Stages - 1) % kG2Stages, (kt + kG2Stages - 1) * kG2BK);
    else cp_async_commit();
    const double* sA = gsm + (kt % kG2Stages) * kG2StageDoubles;
    const double* sB = sA + kG2BM * kG2SA;
#pragma unroll
    for (int kk = 0; kk < kG2BK; kk += 4) {
      double a[8], b[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = sA[(wm + i * 8 + grp) * kG2SA + kk + tig];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sB[(kk + tig) * kG2SB + wn + j * 8 + grp];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  cp_async_wait<0>();
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = m0 + wm + i * 8 + grp;
      const int c = n0 + wn + j * 8 + 2 * tig;
      if (r < M && c < N)  // N even, c even: the pair is inside or outside as a whole
        *reinterpret_cast<double2*>(C + (int64_t)r * N + c) = make_double2(acc[i][j][0], acc[i][j][1]);
    }
}

// ---------------------------------------------------------------------------------------------
// wreath plan tables and layout kernels (snfft_wreath_*): the gather index sigma = t_i h t_j^-1 of
// every (coset pair, Young-subgroup element), its inverse, and the block <-> matrix permutes.
// ---------------------------------------------------------------------------------------------
__global__ void wreath_index_kernel(int n, int N, int nH, const int8_t* __restrict__ T,
                                    const int8_t* __restrict__ Tinv, const int8_t* __restrict__ H,
                                    int64_t* __restrict__ idx, int64_t* __restrict__ inv_idx) {
  const int64_t total = (int64_t)N * N * nH;
  int64_t fact[kMaxN + 1];
  fact[0] = 1;
  for (int i = 1; i <= kMaxN; ++i) fact[i] = fact[i - 1] * i;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int h = (int)(e % nH);
    const int j = (int)((e / nH) % N);
    const int i = (int)(e / ((int64_t)nH * N));
    int sig[kMaxN];
    for (int p = 0; p < n; ++p) sig[p] = T[i * n + H[h * n + Tinv[j * n + p] - 1] - 1];  // t_i(h(t_j^-1(p)))
    int64_t rank = 0;
    for (int a = 0; a < n; ++a) {
      int smaller = 0;
      for (int b = a + 1; b < n; ++b) smaller += sig[b] < sig[a];
      rank += smaller * fact[n - 1 - a];
    }
    const int64_t at = rank * N + i;   // F[sigma][i]
    idx[e] = at;
    inv_idx[at] = e;
  }
}

// planes [N*N][b*b] (coset pair major) <-> matrix [N*b][N*b]
template <bool TO_MATRIX>
__global__ void wreath_permute_kernel(int N, int b, double* __restrict__ mat, double* __restrict__ planes, double scale) {
  const int64_t D = (int64_t)N * b, total = D * D;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = e / D, col = e - row * D;
    const int64_t i = row / b, a = row - i * b, j = col / b, c = col - j * b;
    const int64_t q = (i * N + j) * b * b + a * b + c;
    if (TO_MATRIX) mat[e] = planes[q] * scale; else planes[q] = mat[e] * scale;
  }
}

// Stage 2 of the wreath transform fused (small blocks): per coset pair (i, j) the b x b block
//   sum_h F_i(t_i h t_j^-1) R[h]
// with the gather index transposed to [h][pair] (coalesced), R[h] in shared memory, F as
// interleaved complex128 (one 16-byte gather per term instead of two 8-byte ones in different
// sectors), and the block written straight into the D x D matrices: one kernel instead of gather +
// GEMM + permute per plane.  One thread = one coset pair.
template <int BB>
__global__ void __launch_bounds__(256)
wreath_stage2_kernel(int N, int b, int nH, const int32_t* __restrict__ idx_t, const double* __restrict__ R,
                     const double2* __restrict__ F, double* __restrict__ Mre, double* __restrict__ Mim) {
  extern __shared__ double Rs[];  // [nH][BB]
  for (int e = threadIdx.x; e < nH * BB; e += 256) Rs[e] = R[e];
  __syncthreads();
  const int64_t pairs = (int64_t)N * N, D = (int64_t)N * b;
  for (int64_t p = blockIdx.x * 256LL + threadIdx.x; p < pairs; p += (int64_t)gridDim.x * 256) {
    const int64_t i = p / N, j = p - i * N;
    double ar[BB], ai[BB];
#pragma unroll
    for (int q = 0; q < BB; ++q) ar[q] = ai[q] = 0.0;
    // four independent gathers in flight per thread (the loop is latency-bound: ncu long scoreboard)
    int h = 0;
    for (; h + 4 <= nH; h += 4) {
      double2 fv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) fv[u] = F[idx_t[(int64_t)(h + u) * pairs + p]];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int q = 0; q < BB; ++q) {
          const double r = Rs[(h + u) * BB + q];
          ar[q] = fma(fv[u].x, r, ar[q]);
          ai[q] = fma(fv[u].y, r, ai[q]);
        }
    }
    for (; h < nH; ++h) {
      const double2 fv = F[idx_t[(int64_t)h * pairs + p]];
#pragma unroll
      for (int q = 0; q < BB; ++q) {
        const double r = Rs[h * BB + q];
        ar[q] = fma(fv.x, r, ar[q]);
        ai[q] = fma(fv.y, r, ai[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < BB; ++q) {
      const int a = q / b, c = q - a * b;  // b * b == BB
      Mre[(i * b + a) * D + j * b + c] = ar[q];
      Mim[(i * b + a) * D + j * b + c] = ai[q];
    }
  }
}

// The adjoint of stage 2, F-major: every (sigma, i) belongs to exactly one (coset pair, h)
// (sigma = t_i h t_j^-1: inv_idx), so one thread = one element of F:
//   F_i(sigma) = <block(i, j), R[h]>
// reads the blocks (L2-resident gathers) and WRITES F coalesced - the pair-major form scattered
// 8-byte stores over F (ncu: 0.8 GB read + 1.0 GB written for 0.36 GB of results).  The blocks are
// first regrouped from the two D x D planes into [pair][BB] complex128 (wreath_blocks_kernel): one
// contiguous 16 BB-byte gather per element instead of 2 BB 8-byte ones (L2 traffic 6.4 -> 1.5 GB).
__global__ void wreath_blocks_kernel(int N, int b, const double* __restrict__ Mre, const double* __restrict__ Mim,
                                     double2* __restrict__ blocks) {
  const int64_t D = (int64_t)N * b, total = D * D;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = e / D, col = e - row * D;
    const int64_t i = row / b, a = row - i * b, j = col / b, c = col - j * b;
    blocks[(i * N + j) * b * b + a * b + c] = make_double2(Mre[e], Mim[e]);
  }
}

template <int BB>
__global__ void __launch_bounds__(256)
wreath_stage2_adjoint_kernel(int nH, int64_t total, const int64_t* __restrict__ inv_idx,
                             const double* __restrict__ R, double2* __restrict__ F,
                             const double2* __restrict__ blocks) {
  extern __shared__ double Rs[];  // [nH][BB]
  for (int e = threadIdx.x; e < nH * BB; e += 256) Rs[e] = R[e];
  __syncthreads();
  for (int64_t at = blockIdx.x * 256LL + threadIdx.x; at < total; at += (int64_t)gridDim.x * 256) {
    const uint64_t e = (uint64_t)inv_idx[at];  // (i N + j) nH + h
    const uint64_t pair = e / (uint64_t)nH;
    const int h = (int)(e - pair * (uint64_t)nH);
    const double2* __restrict__ blk = blocks + pair * BB;
    double er = 0.0, ei = 0.0;
#pragma unroll
    for (int q = 0; q < BB; ++q) {
      const double2 v = blk[q];
      const double r = Rs[h * BB + q];
      er = fma(v.x, r, er);
      ei = fma(v.y, r, ei);
    }
    F[at] = make_double2(er, ei);
  }
}

// Sampled wreath inverse (cube_ift.py:56-67 per state, the consumer of compute_policy.py:66-88):
//   f(o, sigma) = D / |G| * sum_i conj(chi_i(o)) <block(i, j_i), R[h_i]>,   sigma = t_i h_i t_{j_i}^-1,
// i.e. the F-major adjoint of stage 2 at the row sigma followed by ONE term of the adjoint DFT.
// One CTA = one state; threads run over the cosets i.  States out of range give NaN.
__global__ void __launch_bounds__(256)
wreath_inverse_at_kernel(int m, int N, int b, int nH, int64_t nfact, int n_ori, const int64_t* __restrict__ inv_idx,
                         const double* __restrict__ R, const double* __restrict__ Mre,
                         const double* __restrict__ Mim, const int32_t* __restrict__ freq,
                         const int64_t* __restrict__ sigma, const int32_t* __restrict__ ori, double scale,
                         double* __restrict__ out_re, double* __restrict__ out_im) {
  __shared__ double part[2][8];
  const int64_t s = sigma[blockIdx.x];
  const int o = ori[blockIdx.x];
  const bool ok = s >= 0 && s < nfact && o >= 0 && o < n_ori;
  const int64_t D = (int64_t)N * b;
  const int bb = b * b;
  double ar = 0.0, ai = 0.0;
  for (int i = threadIdx.x; ok && i < N; i += 256) {
    const uint64_t e = (uint64_t)inv_idx[s * N + i];  // (i N + j) nH + h
    const uint64_t pair = e / (uint64_t)nH;
    const int h = (int)(e - pair * (uint64_t)nH);
    const int64_t j = (int64_t)(pair - (uint64_t)i * (uint64_t)N);
    double er = 0.0, ei = 0.0;
    for (int q = 0; q < bb; ++q) {
      const int a = q / b, c = q - a * b;
      const double r = R[h * bb + q];
      er = fma(Mre[((int64_t)i * b + a) * D + j * b + c], r, er);
      ei = fma(Mim[((int64_t)i * b + a) * D + j * b + c], r, ei);
    }
    int ex = 0;  // k_i . o over the base-3 digits
    for (int d = 0, k = freq[i], oo = o; d < m; ++d, k /= 3, oo /= 3) ex += (k % 3) * (oo % 3);
    ex %= 3;
    // conj(chi) = omega^{-ex}
    const double cr = ex == 0 ? 1.0 : -0.5;
    const double ci = ex == 0 ? 0.0 : (ex == 1 ? -0.8660254037844386467637231707529 : 0.8660254037844386467637231707529);
    ar += er * cr - ei * ci;
    ai += er * ci + ei * cr;
  }
  for (int w = 16; w > 0; w >>= 1) {
    ar += __shfl_down_sync(0xffffffffu, ar, w);
    ai += __shfl_down_sync(0xffffffffu, ai, w);
  }
  if ((threadIdx.x & 31) == 0) {
    part[0][threadIdx.x >> 5] = ar;
    part[1][threadIdx.x >> 5] = ai;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double sr = 0.0, si = 0.0;
    for (int w = 0; w < 8; ++w) {
      sr += part[0][w];
      si += part[1][w];
    }
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    out_re[blockIdx.x] = ok ? scale * sr : nan;
    out_im[blockIdx.x] = ok ? scale * si : nan;
  }
}

// the gather index transposed to [h][pair] and narrowed to 32 bits (n! N < 2^31 for every plan)
__global__ void transpose_i64_kernel(const int64_t* __restrict__ in, int32_t* __restrict__ out, int64_t rows, int64_t cols) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < rows * cols; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / cols, c = e - r * cols;
    out[c * rows + r] = (int32_t)in[e];
  }
}

// [batch][rows][cols] -> [batch][cols][rows]
__global__ void transpose_batched_kernel(const double* __restrict__ in, double* __restrict__ out, int64_t batch,
                                         int64_t rows, int64_t cols) {
  const int64_t per = rows * cols;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < batch * per; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t bq = e / per, r = (e - bq * per) / cols, c = e - bq * per - r * cols;
    out[bq * per + c * rows + r] = in[e];
  }
}

// Blocks of a Young-subgroup irrep between the nesting the factor transforms leave,
//   X[pair][e3][e2][e1], e_k = r_k d_k + c_k (entry of the factor-k matrix),
// and the Kronecker order of R[h] = rho_1 (x) rho_2 (x) rho_3 (wreath.py:191-224),
//   P[pair][((r1 d2 + r2) d3 + r3) b + (c1 d2 + c2) d3 + c3],  b = d1 d2 d3.
template <bool TO_KRON>
__global__ void wreath_kron_permute_kernel(int64_t pairs, int d1, int d2, int d3, double* __restrict__ X,
                                           double* __restrict__ P) {
  const int64_t b = (int64_t)d1 * d2 * d3, bb = b * b;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < pairs * bb; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = e / bb, q = e - p * bb;  // q indexes P
    const int64_t row = q / b, col = q - row * b;
    const int r3 = (int)(row % d3), r2 = (int)((row / d3) % d2), r1 = (int)(row / ((int64_t)d3 * d2));
    const int c3 = (int)(col % d3), c2 = (int)((col / d3) % d2), c1 = (int)(col / ((int64_t)d3 * d2));
    const int64_t e1 = r1 * d1 + c1, e2 = r2 * d2 + c2, e3 = r3 * d3 + c3;
    const int64_t x = p * bb + (e3 * d2 * d2 + e2) * d1 * d1 + e1;
    if (TO_KRON) P[e] = X[x]; else X[x] = P[e];
  }
}

__global__ void transpose_kernel(const double* __restrict__ in, double* __restrict__ out, int64_t rows, int64_t cols) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < rows * cols; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / cols, c = e - r * cols;
    out[c * rows + r] = in[e];
  }
}

__global__ void gather_kernel(const double* __restrict__ src, const int64_t* __restrict__ idx,
                              double* __restrict__ dst, int64_t count, int width) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < count * width;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / width;
    dst[e] = src[idx[r] * width + (e - r * width)];
  }
}

// ---------------------------------------------------------------------------------------------
// The one exchange of the sharded transform over peer memory (NVLink P2P inside one kernel): no
// packing pass on the sender, no unpacking pass on the receiver, no NCCL.  Both directions walk
// (own block b, position p of the concatenated per-rank column lists) and keep the REMOTE side
// contiguous (runs of size(r, m) doubles per block), the scattered side local:
//   forward (push): peer[r].cols[first_q + b][p - off[r]] = spectra[b][idx[p]]
//   inverse (pull): spectra[b][idx[p]] = peer[r].cols[first_q + b][p - off[r]]
// idx[off[r] .. off[r+1]) = positions of rank r's columns inside a packed S_m spectrum.
// ---------------------------------------------------------------------------------------------
template <bool INV, int UNROLL>
__global__ void __launch_bounds__(256)
shard_exchange_kernel(ShardPull X, double* __restrict__ local) {
  const int q = X.rank;
  const int64_t nb = X.first[q + 1] - X.first[q];
  const int stride = gridDim.x * blockDim.x;
  for (int64_t b = blockIdx.y; b < nb; b += gridDim.y) {
    double* __restrict__ spec = local + b * X.mfact;
    const int64_t B = X.first[q] + b;
    for (int p0 = blockIdx.x * blockDim.x + threadIdx.x; p0 < X.mfact; p0 += stride * UNROLL) {
      double v[UNROLL];
      double* remote[UNROLL];
      int at[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int p = p0 + u * stride;
        at[u] = -1;
        if (p < X.mfact) {
          int r = 0;
          while (r + 1 < X.world && p >= X.off[r + 1]) ++r;
          const int wr = X.off[r + 1] - X.off[r];
          remote[u] = const_cast<double*>(X.peer[r]) + B * wr + (p - X.off[r]);
          at[u] = X.idx[p];
          v[u] = INV ? *remote[u] : spec[at[u]];
        }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
        if (at[u] >= 0) {
          if (INV)
            spec[at[u]] = v[u];
          else
            *remote[u] = v[u];
        }
    }
  }
}

// Barrier between the ranks of a sharded transform, on the device: every rank raises its flag in
// every peer's signal pad (release, system scope) and waits until all peers have raised theirs in
// its own pad (acquire).  Flags are epochs that only grow, so the pads are never reset.  A peer that
// never arrives makes the kernel give up after ~2 s and set the status word instead of hanging.
__global__ void shard_barrier_kernel(ShardBarrier Bq) {
  const int t = threadIdx.x;
  if (t >= Bq.world) return;
  __threadfence_system();
  unsigned long long* theirs = Bq.pads[t] + (size_t)Bq.channel * kMaxWorld + Bq.rank;
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(theirs), "l"(Bq.epoch) : "memory");
  const unsigned long long* mine = Bq.pads[Bq.rank] + (size_t)Bq.channel * kMaxWorld + t;
  const long long t0 = clock64();
  while (true) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
    if (v >= Bq.epoch) break;
    if (clock64() - t0 > 4000000000LL) {
      *Bq.status = 1;  // a peer did not arrive
      break;
    }
    __nanosleep(200);
  }
  __threadfence_system();
}

Facts make_facts() {
  Facts F;
  F.f[0] = 1;
  for (int i = 1; i <= kMaxN; ++i) F.f[i] = F.f[i - 1] * i;
  return F;
}

}  // namespace

namespace {
template <class K>
cudaError_t allow_big_smem(K kernel, int kb = 227) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kb * 1024);
}
}  // namespace

cudaError_t prepare_kernels() {
  cudaError_t e;
  if ((e = allow_big_smem(level_kernel<false>)) != cudaSuccess) return e;
  if ((e = allow_big_smem(level_kernel<true>)) != cudaSuccess) return e;
  if ((e = allow_big_smem(dgemm_pipe_kernel, 200)) != cudaSuccess) return e;
  if ((e = allow_big_smem(wreath_stage2_kernel<1>, 100)) != cudaSuccess) return e;
  if ((e = allow_big_smem(wreath_stage2_adjoint_kernel<1>, 100)) != cudaSuccess) return e;
  if ((e = allow_big_smem(wreath_stage2_kernel<4>, 100)) != cudaSuccess) return e;
  if ((e = allow_big_smem(wreath_stage2_adjoint_kernel<4>, 100)) != cudaSuccess) return e;
  if ((e = allow_big_smem(wreath_stage2_kernel<9>, 100)) != cudaSuccess) return e;
  if ((e = allow_big_smem(wreath_stage2_adjoint_kernel<9>, 100)) != cudaSuccess) return e;
  if ((e = allow_big_smem(wreath_stage2_kernel<16>, 100)) != cudaSuccess) return e;
  if ((e = allow_big_smem(wreath_stage2_adjoint_kernel<16>, 100)) != cudaSuccess) return e;
  // (kernels with static shared memory of their own: dynamic + static must stay <= 227 KB)
  if ((e = allow_big_smem(inverse_at_kernel, 200)) != cudaSuccess) return e;
  if ((e = allow_big_smem(c3_dft_kernel<false, false>, 200)) != cudaSuccess) return e;
  if ((e = allow_big_smem(c3_dft_kernel<false, true>, 200)) != cudaSuccess) return e;
  if ((e = allow_big_smem(c3_dft_kernel<true, false>, 200)) != cudaSuccess) return e;
  if ((e = allow_big_smem(c3_dft_kernel<true, true>, 200)) != cudaSuccess) return e;
  if ((e = allow_big_smem(phase_kernel<false, 1>)) != cudaSuccess) return e;
  if ((e = allow_big_smem(phase_kernel<true, 1>)) != cudaSuccess) return e;
  if ((e = allow_big_smem(phase_kernel<false, 2>)) != cudaSuccess) return e;
  if ((e = allow_big_smem(phase_kernel<true, 2>)) != cudaSuccess) return e;
  if ((e = allow_big_smem(phase_kernel<false, 4>)) != cudaSuccess) return e;
  if ((e = allow_big_smem(phase_kernel<true, 4>)) != cudaSuccess) return e;
  if ((e = allow_big_smem(level8_kernel<false>)) != cudaSuccess) return e;
  if ((e = allow_big_smem(level8_kernel<true>)) != cudaSuccess) return e;
  if ((e = allow_big_smem(fused_kernel<4, false>)) != cudaSuccess) return e;
  if ((e = allow_big_smem(fused_kernel<4, true>)) != cudaSuccess) return e;
  if ((e = allow_big_smem(fused_kernel<5, false>)) != cudaSuccess) return e;
  if ((e = allow_big_smem(fused_kernel<5, true>)) != cudaSuccess) return e;
  if ((e = allow_big_smem(fused_kernel<6, false>)) != cudaSuccess) return e;
  if ((e = allow_big_smem(fused_kernel<6, true>)) != cudaSuccess) return e;
  if ((e = allow_big_smem(fused_kernel<7, false>)) != cudaSuccess) return e;
  if ((e = allow_big_smem(fused_kernel<7, true>)) != cudaSuccess) return e;
  if ((e = allow_big_smem(fused_kernel<7, true, true>, 200)) != cudaSuccess) return e;
  return cudaSuccess;
}

namespace {
constexpr int kL8SmemDoubles = 6144;  // >= 5 * 35 * 35: C_5 rows of the widest panel (49 KB)

cudaError_t launch_level8(const double* in, double* out, int64_t groups, bool inverse, cudaStream_t stream) {
  static const int dims[gen::kL8NumPanels] = {1, 6, 15, 20, 15, 6, 1, 14, 35, 35, 14, 21, 14, 21, 14};
  L8Launch launch;
  int order[gen::kL8NumPanels];
  for (int p = 0; p < gen::kL8NumPanels; ++p) order[p] = p;
  std::stable_sort(order, order + gen::kL8NumPanels, [&](int a, int b) { return dims[a] > dims[b]; });
  int64_t total = 0;
  for (int i = 0; i < gen::kL8NumPanels; ++i) {
    const int d = dims[order[i]];
    int cap = kL8SmemDoubles / (5 * d * d);
    if (cap > 64) cap = 64;
    if (cap > kL8Super) cap = kL8Super;
    if (cap < 1) cap = 1;
    launch.order[i] = order[i];
    launch.cap[i] = cap;
    launch.first_cta[i] = (int)total;
    total += (kL8Super + cap - 1) / cap;
  }
  launch.first_cta[gen::kL8NumPanels] = (int)total;  // CTAs per super-chunk
  total *= (groups + kL8Super - 1) / kL8Super;
  if (total > 0x7fffffffLL) return cudaErrorInvalidValue;
  LaunchScope scope(inverse ? kKindLevel8Inv : kKindLevel8Fwd, stream);
  const size_t smem = kL8SmemDoubles * sizeof(double);
  if (inverse)
    level8_kernel<true><<<(unsigned)total, kL8Threads, smem, stream>>>(launch, in, out, groups);
  else
    level8_kernel<false><<<(unsigned)total, kL8Threads, smem, stream>>>(launch, in, out, groups);
  return cudaGetLastError();
}
}  // namespace

cudaError_t launch_level(const LevelDev& L, const Tile* tiles_dev, int ntiles, size_t smem_bytes,
                         const double* in, double* out, int64_t groups, bool inverse, bool table_driven,
                         int num_sms, cudaStream_t stream) {
  if (groups <= 0 || ntiles <= 0) return cudaSuccess;
  if (L.k == 8 && !table_driven) return launch_level8(in, out, groups, inverse, stream);
  // enough CTAs to fill the machine a few times over; each CTA strides over group chunks
  int64_t y = (8LL * num_sms + ntiles - 1) / ntiles;
  if (y > groups) y = groups;
  if (y < 1) y = 1;
  if (y > 65535) y = 65535;
  dim3 grid((unsigned)ntiles, (unsigned)y);
  LaunchScope scope(kKindLevelTable, stream);
  if (inverse)
    level_kernel<true><<<grid, kThreads, smem_bytes, stream>>>(L, tiles_dev, in, out, groups);
  else
    level_kernel<false><<<grid, kThreads, smem_bytes, stream>>>(L, tiles_dev, in, out, groups);
  return cudaGetLastError();
}

template <int K0>
static cudaError_t launch_fused_k0(const FusedDev& F, const double* in, double* out, int64_t groups,
                                   bool inverse, int num_sms, cudaStream_t stream) {
  const size_t smem = (size_t)(kFusedP + kFusedQ) * sizeof(double);
  const int64_t per_iter = FusedShape<K0>::kGroups;
  int64_t grid = (groups + per_iter - 1) / per_iter;
  if (grid > 2LL * num_sms * SNFFT_FUSED_WAVES) grid = 2LL * num_sms * SNFFT_FUSED_WAVES;  // persistent beyond that
  LaunchScope scope(inverse ? kKindFusedInv : kKindFusedFwd, stream);
  if (inverse && K0 == 7 && g_fused_tma.load(std::memory_order_relaxed) &&
      (reinterpret_cast<uintptr_t>(in) & 15) == 0)
    fused_kernel<7, true, true><<<(unsigned)grid, kFusedThreads, smem, stream>>>(F, in, out, groups);
  else if (inverse)
    fused_kernel<K0, true><<<(unsigned)grid, kFusedThreads, smem, stream>>>(F, in, out, groups);
  else
    fused_kernel<K0, false><<<(unsigned)grid, kFusedThreads, smem, stream>>>(F, in, out, groups);
  return cudaGetLastError();
}

cudaError_t launch_phase(const PhaseDev& P, const int32_t* panel_groups, const int32_t* panel_tiles,
                         const int32_t* panel_pack, double* xside, double* sside, int64_t lgroups,
                         bool inverse, int num_sms, cudaStream_t stream) {
  if (lgroups <= 0 || P.npanels <= 0) return cudaSuccess;
  if (P.npanels > kPhaseMaxPanels || lgroups > 0x7fffffffLL) return cudaErrorInvalidValue;
  // shared memory of a warp: its tile (A chunks of 32 columns side by side) plus the largest program
  // blob; as many warps per CTA (one CTA per SM) as fit, at most 16.  Two chunks per warp halve the
  // per-block overhead and double the independent work of a warp: measured faster even when only
  // 7 warps fit (48-row groups; S_12 forward + inverse 45.0 -> 41.7 ms).  SNFFT_PHASE_A overrides.
  int A = 1;
  {
    const size_t two = ((size_t)P.max_rows * 64 + P.max_blob * 2) * sizeof(double);
    const size_t four = ((size_t)P.max_rows * 128 + P.max_blob * 2) * sizeof(double);
    if ((226 * 1024) / two >= 6) A = 2;
    if ((226 * 1024) / four >= 8) A = 4;
    if (const char* env = std::getenv("SNFFT_PHASE_A")) {
      const int want = std::atoi(env);
      if (want == 1 || want == 2 || (want == 4 && (226 * 1024) / four >= 1)) A = want;
    }
  }
  const int warp_doubles = P.max_rows * 32 * A + P.max_blob * 2;
  int nwarps = (int)((226 * 1024) / ((size_t)warp_doubles * sizeof(double)));
  if (nwarps > 16) nwarps = 16;
  if (nwarps < 1) return cudaErrorInvalidValue;
  // units of work: (group, segment of its chunks); segments long enough to amortise the blob copy,
  // short enough to give every warp of the machine several units
  PhaseLaunch Lp;
  int64_t total_chunks = 0;
  for (int p = 0; p < P.npanels; ++p) {
    const int64_t nlg = (lgroups + panel_pack[p] - 1) / panel_pack[p];
    Lp.nchunks[p] = (uint32_t)(panel_tiles[p] * nlg);  // chunks of 32 columns
    total_chunks += (int64_t)panel_groups[p] * Lp.nchunks[p];
  }
  const int64_t warps_total = (int64_t)num_sms * nwarps;
  int64_t seg = total_chunks / (warps_total * 6);
  if (seg < 4) seg = 4;
  if (seg > 64) seg = 64;
  seg = (seg + A - 1) / A * A;  // a warp runs A chunks at a time
  int64_t at = 0;
  for (int p = 0; p < P.npanels; ++p) {
    Lp.work_off[p] = (uint32_t)at;
    Lp.seg_len[p] = (uint32_t)seg;
    at += (int64_t)panel_groups[p] * ((Lp.nchunks[p] + seg - 1) / seg);
    if (at > 0x7fffffffLL) return cudaErrorInvalidValue;
  }
  for (int p = P.npanels; p <= kPhaseMaxPanels; ++p) Lp.work_off[p] = (uint32_t)at;
  for (int p = P.npanels; p < kPhaseMaxPanels; ++p) Lp.seg_len[p] = 1, Lp.nchunks[p] = 0;
  if (at == 0) return cudaSuccess;
  int64_t grid = (at + nwarps - 1) / nwarps;
  if (grid > num_sms) grid = num_sms;  // persistent: one CTA per SM, units of work grid-strided
  const size_t smem = (size_t)nwarps * warp_doubles * sizeof(double);
  LaunchScope scope(inverse ? kKindPhaseInv : kKindPhaseFwd, stream);
  if (inverse) {
    if (A == 4)
      phase_kernel<true, 4><<<(unsigned)grid, nwarps * 32, smem, stream>>>(P, Lp, xside, sside, (uint32_t)lgroups, warp_doubles);
    else if (A == 2)
      phase_kernel<true, 2><<<(unsigned)grid, nwarps * 32, smem, stream>>>(P, Lp, xside, sside, (uint32_t)lgroups, warp_doubles);
    else
      phase_kernel<true, 1><<<(unsigned)grid, nwarps * 32, smem, stream>>>(P, Lp, xside, sside, (uint32_t)lgroups, warp_doubles);
  } else {
    if (A == 4)
      phase_kernel<false, 4><<<(unsigned)grid, nwarps * 32, smem, stream>>>(P, Lp, xside, sside, (uint32_t)lgroups, warp_doubles);
    else if (A == 2)
      phase_kernel<false, 2><<<(unsigned)grid, nwarps * 32, smem, stream>>>(P, Lp, xside, sside, (uint32_t)lgroups, warp_doubles);
    else
      phase_kernel<false, 1><<<(unsigned)grid, nwarps * 32, smem, stream>>>(P, Lp, xside, sside, (uint32_t)lgroups, warp_doubles);
  }
  return cudaGetLastError();
}

cudaError_t launch_p1_inplace(const P1Dev& P, double* xside, double* sside, int64_t lgroups, bool inverse,
                              int num_sms, cudaStream_t stream) {
  if (lgroups <= 0) return cudaSuccess;
  // the in-place programs hold up to 36 rows in registers (172..184 registers per thread): small CTAs
  // pack more warps onto an SM (register allocation is per CTA): S_12 p1 time 7.29 ms with 256
  // threads, 7.16 with 128, 7.13 with 64 or 32.  SNFFT_P1_THREADS overrides.
  int p1_threads = 64;
  if (const char* env = std::getenv("SNFFT_P1_THREADS")) {
    const int want = std::atoi(env);
    if (want == 32 || want == 64 || want == 128 || want == 256) p1_threads = want;
  }
  auto grid_for_warps = [&](int64_t items, int warps) {
    int64_t gy = lgroups < 64 ? lgroups : 64;
    int64_t gx = (items + warps - 1) / warps;
    const int64_t want = 32LL * num_sms * 8 / warps;
    if (gx * gy > want) gx = (want + gy - 1) / gy;
    if (gx < 1) gx = 1;
    return dim3((unsigned)gx, (unsigned)gy);
  };
  auto grid_for = [&](int64_t items) { return grid_for_warps(items, 8); };
  // forward: programs, then move the finished rows out; inverse: bring them in (scaled) first
  for (int step = 0; step < 2; ++step) {
    const bool rows_pass = inverse ? step == 0 : step == 1;
    if (rows_pass) {
      if (P.nrow_items <= 0) continue;
      LaunchScope scope(kKindP0, stream);
      if (inverse)
        p0_rows_kernel<true><<<grid_for(P.nrow_items), 256, 0, stream>>>(P, xside, sside, lgroups);
      else
        p0_rows_kernel<false><<<grid_for(P.nrow_items), 256, 0, stream>>>(P, xside, sside, lgroups);
    } else {
      if (P.nitems <= 0) continue;
      LaunchScope scope(kKindP1, stream);
      if (inverse)
        p1_inplace_kernel<true><<<grid_for_warps(P.nitems, p1_threads / 32), p1_threads, 0, stream>>>(P, xside, lgroups);
      else
        p1_inplace_kernel<false><<<grid_for_warps(P.nitems, p1_threads / 32), p1_threads, 0, stream>>>(P, xside, lgroups);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

cudaError_t launch_fused(const FusedDev& F, const double* in, double* out, int64_t groups,
                         bool inverse, int num_sms, cudaStream_t stream) {
  if (groups <= 0) return cudaSuccess;
  switch (F.k0) {
    case 4: return launch_fused_k0<4>(F, in, out, groups, inverse, num_sms, stream);
    case 5: return launch_fused_k0<5>(F, in, out, groups, inverse, num_sms, stream);
    case 6: return launch_fused_k0<6>(F, in, out, groups, inverse, num_sms, stream);
    case 7: return launch_fused_k0<7>(F, in, out, groups, inverse, num_sms, stream);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_build_lex_table(int n, uint32_t* table, cudaStream_t stream) {
  const Facts F = make_facts();
  const int64_t total = F.f[n];
  const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  LaunchScope scope(kKindOther, stream);
  build_lex_table_kernel<<<blocks, 256, 0, stream>>>(n, F, table);
  return cudaGetLastError();
}

cudaError_t launch_reorder(int n, const uint32_t* table, const double* src, double* dst,
                           int64_t batch, bool to_coset, cudaStream_t stream) {
  if (batch <= 0) return cudaSuccess;
  const Facts F = make_facts();
  const int64_t total = F.f[n];
  LaunchScope scope(kKindReorder, stream);
  if (n >= kTileMinN && !g_force_table_driven) {
    int nsubsets = 1;  // C(n, R)
    for (int t = 1; t <= kTileR; ++t) nsubsets = nsubsets * (n - kTileR + t) / t;
    const int64_t nmid = F.f[n - kTileR] / kTileSF;
    int64_t by = batch < 16 ? batch : 16;
    // enough CTAs for ~16 waves of the machine, at most 8 tiles of one subset per CTA
    int mids = 1;
    while (mids < 8 && (int64_t)nsubsets * ((nmid + mids) / (mids + 1)) * by >= 148 * 8 * 16) ++mids;
    int64_t bx = (int64_t)nsubsets * ((nmid + mids - 1) / mids);
    if (bx > 0x7fffffffLL) return cudaErrorInvalidValue;
    dim3 grid((unsigned)bx, (unsigned)by);
    if (to_coset)
      reorder_tiled_kernel<true><<<grid, kTileThreads, 0, stream>>>(n, F, nsubsets, nmid, mids, src, dst, batch);
    else
      reorder_tiled_kernel<false><<<grid, kTileThreads, 0, stream>>>(n, F, nsubsets, nmid, mids, src, dst, batch);
    return cudaGetLastError();
  }
  int64_t bx = (total + 255) / 256;
  if (bx > 2048) bx = 2048;
  int64_t by = batch < 64 ? batch : 64;
  dim3 grid((unsigned)bx, (unsigned)by);
  if (to_coset)
    reorder_kernel<true><<<grid, 256, 0, stream>>>(n, F, table, src, dst, batch);
  else
    reorder_kernel<false><<<grid, 256, 0, stream>>>(n, F, table, src, dst, batch);
  return cudaGetLastError();
}

cudaError_t launch_yor(int n, int d, const int32_t* partner, const double* diag,
                       const int32_t* perms, int64_t count, double* out, bool seminormal, cudaStream_t stream) {
  if (count <= 0) return cudaSuccess;
  LaunchScope scope(kKindYor, stream);
  if (seminormal)
    yor_kernel<true><<<(unsigned)count, kThreads, 0, stream>>>(n, d, partner, diag, perms, out);
  else
    yor_kernel<false><<<(unsigned)count, kThreads, 0, stream>>>(n, d, partner, diag, perms, out);
  return cudaGetLastError();
}

cudaError_t launch_inverse_at(int n, const InvAtItem* items, int nitems, const InvAtIrrep* irreps, size_t smem_bytes,
                              const double* fhat, const int32_t* perms, int64_t count, double* out,
                              cudaStream_t stream) {
  if (count <= 0 || nitems <= 0) return cudaSuccess;
  if (count > 65535) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(out, 0, (size_t)count * sizeof(double), stream);
  if (e != cudaSuccess) return e;
  dim3 grid((unsigned)nitems, (unsigned)count);
  LaunchScope scope(kKindYor, stream);
  inverse_at_kernel<<<grid, kThreads, smem_bytes, stream>>>(n, items, irreps, fhat, perms, out);
  return cudaGetLastError();
}

static unsigned grid_for(int64_t total) {
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

cudaError_t launch_wreath_index(int n, int N, int nH, const int8_t* T, const int8_t* Tinv, const int8_t* H,
                                int64_t* idx, int64_t* inv_idx, cudaStream_t stream) {
  LaunchScope scope(kKindOther, stream);
  wreath_index_kernel<<<grid_for((int64_t)N * N * nH), 256, 0, stream>>>(n, N, nH, T, Tinv, H, idx, inv_idx);
  return cudaGetLastError();
}

cudaError_t launch_wreath_permute(int N, int b, double* mat, double* planes, double scale, bool to_matrix,
                                  cudaStream_t stream) {
  LaunchScope scope(kKindOther, stream);
  const int64_t D = (int64_t)N * b;
  if (to_matrix)
    wreath_permute_kernel<true><<<grid_for(D * D), 256, 0, stream>>>(N, b, mat, planes, scale);
  else
    wreath_permute_kernel<false><<<grid_for(D * D), 256, 0, stream>>>(N, b, mat, planes, scale);
  return cudaGetLastError();
}

cudaError_t launch_wreath_inverse_at(int m, int N, int b, int nH, int64_t nfact, int n_ori, const int64_t* inv_idx,
                                     const double* R, const double* Mre, const double* Mim, const int32_t* freq,
                                     const int64_t* sigma, const int32_t* ori, int64_t count, double scale,
                                     double* out_re, double* out_im, cudaStream_t stream) {
  if (count <= 0) return cudaSuccess;
  if (count > 0x7fffffffLL) return cudaErrorInvalidValue;
  LaunchScope scope(kKindOther, stream);
  wreath_inverse_at_kernel<<<(unsigned)count, 256, 0, stream>>>(m, N, b, nH, nfact, n_ori, inv_idx, R, Mre, Mim, freq,
                                                               sigma, ori, scale, out_re, out_im);
  return cudaGetLastError();
}

cudaError_t launch_transpose_batched(const double* in, double* out, int64_t batch, int64_t rows, int64_t cols,
                                     cudaStream_t stream) {
  if (batch * rows * cols <= 0) return cudaSuccess;
  LaunchScope scope(kKindOther, stream);
  transpose_batched_kernel<<<grid_for(batch * rows * cols), 256, 0, stream>>>(in, out, batch, rows, cols);
  return cudaGetLastError();
}

cudaError_t launch_wreath_kron_permute(int64_t pairs, int d1, int d2, int d3, double* X, double* P, bool to_kron,
                                       cudaStream_t stream) {
  LaunchScope scope(kKindOther, stream);
  const int64_t b = (int64_t)d1 * d2 * d3;
  if (to_kron)
    wreath_kron_permute_kernel<true><<<grid_for(pairs * b * b), 256, 0, stream>>>(pairs, d1, d2, d3, X, P);
  else
    wreath_kron_permute_kernel<false><<<grid_for(pairs * b * b), 256, 0, stream>>>(pairs, d1, d2, d3, X, P);
  return cudaGetLastError();
}

cudaError_t launch_transpose_i64(const int64_t* in, int32_t* out, int64_t rows, int64_t cols, cudaStream_t stream) {
  LaunchScope scope(kKindOther, stream);
  transpose_i64_kernel<<<grid_for(rows * cols), 256, 0, stream>>>(in, out, rows, cols);
  return cudaGetLastError();
}

bool wreath_stage2_fits(int N, int b, int nH) {
  const int bb = b * b;
  return (bb == 1 || bb == 4 || bb == 9 || bb == 16) && N >= 32 && (size_t)nH * bb * sizeof(double) <= 96 * 1024;
}

cudaError_t launch_wreath_stage2(int N, int b, int nH, int64_t nfact, const int32_t* idx_t, const int64_t* inv_idx,
                                 const double* R, double* F, double* Mre, double* Mim, double* blocks, bool inverse,
                                 cudaStream_t stream) {
  const int bb = b * b;
  const size_t smem = (size_t)nH * bb * sizeof(double);
  const int64_t total = nfact * N;  // elements of F = (coset pairs) x |S_alpha|
  const unsigned grid = inverse ? grid_for(total) : grid_for((int64_t)N * N);
  double2* F2 = reinterpret_cast<double2*>(F);
  double2* B2 = reinterpret_cast<double2*>(blocks);
  if (inverse) {
    LaunchScope scope(kKindOther, stream);
    wreath_blocks_kernel<<<grid_for((int64_t)N * b * N * b), 256, 0, stream>>>(N, b, Mre, Mim, B2);
  }
  LaunchScope scope(kKindDgemm, stream);
#define SNFFT_S2(BBV)                                                                                          \
  if (inverse)                                                                                                 \
    wreath_stage2_adjoint_kernel<BBV><<<grid, 256, smem, stream>>>(nH, total, inv_idx, R, F2, B2);              \
  else                                                                                                         \
    wreath_stage2_kernel<BBV><<<grid, 256, smem, stream>>>(N, b, nH, idx_t, R, F2, Mre, Mim)
  switch (bb) {
    case 1: SNFFT_S2(1); break;
    case 4: SNFFT_S2(4); break;
    case 9: SNFFT_S2(9); break;
    case 16: SNFFT_S2(16); break;
    default: return cudaErrorInvalidValue;
  }
#undef SNFFT_S2
  return cudaGetLastError();
}

cudaError_t launch_transpose(const double* in, double* out, int64_t rows, int64_t cols, cudaStream_t stream) {
  LaunchScope scope(kKindOther, stream);
  transpose_kernel<<<grid_for(rows * cols), 256, 0, stream>>>(in, out, rows, cols);
  return cudaGetLastError();
}

cudaError_t launch_c3_dft(int m, const double* in_re, const double* in_im, const int32_t* freq, int nfreq,
                          int64_t count, double* out_re, double* out_im, bool inverse, double scale,
                          cudaStream_t stream) {
  if (count <= 0) return cudaSuccess;
  if (m < 1 || m > 8 || count > 0x7fffffffLL) return cudaErrorInvalidValue;
  int size = 1;
  for (int a = 0; a < m; ++a) size *= 3;
  const size_t smem = 2 * (size_t)size * sizeof(double);
  LaunchScope scope(kKindOther, stream);
  if (!inverse && out_im)
    c3_dft_kernel<false, false><<<(unsigned)((count + 1) / 2), kDftThreadsFwd, smem, stream>>>(
        m, size, count, in_re, in_im, freq, nfreq, out_re, out_im, scale);
  else if (!inverse)
    c3_dft_kernel<false, true><<<(unsigned)((count + 1) / 2), kDftThreadsFwd, smem, stream>>>(
        m, size, count, in_re, in_im, freq, nfreq, out_re, out_im, scale);
  else if (out_im)
    c3_dft_kernel<true, false><<<(unsigned)count, kDftThreadsAdj, smem, stream>>>(m, size, count, in_re, in_im, freq,
                                                                               nfreq, out_re, out_im, scale);
  else
    c3_dft_kernel<true, true><<<(unsigned)count, kDftThreadsAdj, smem, stream>>>(m, size, count, in_re, in_im, freq,
                                                                              nfreq, out_re, out_im, scale);
  return cudaGetLastError();
}

cudaError_t launch_gather(const double* src, const int64_t* idx, double* dst, int64_t count, int width,
                          cudaStream_t stream) {
  if (count <= 0) return cudaSuccess;
  int64_t blocks = (count * width + 255) / 256;
  if (blocks > 65535 * 16) blocks = 65535 * 16;
  LaunchScope scope(kKindOther, stream);
  gather_kernel<<<(unsigned)blocks, 256, 0, stream>>>(src, idx, dst, count, width);
  return cudaGetLastError();
}

cudaError_t launch_shard_pull(const ShardPull& X, double* local, bool inverse, cudaStream_t stream) {
  const int q = X.rank;
  const int64_t rows = X.first[q + 1] - X.first[q];
  if (rows <= 0 || X.mfact <= 0) return cudaSuccess;
  // tuning knobs for the NVLink sweep (tools/exchange_sweep.sh): loads in flight per thread, CTAs per block row
  static const int unroll = [] { const char* e = std::getenv("SNFFT_XCHG_UNROLL"); return e ? std::atoi(e) : 4; }();
  // measured on 2 GPUs (S_12, 958 MB per rank per direction): 256 CTAs per block row with a grid-stride
  // loop 456 GB/s, 1024: 532, 8192: 564, one CTA per 1024 positions (no loop): 579 GB/s
  static const int gx_cap = [] { const char* e = std::getenv("SNFFT_XCHG_GX"); return e ? std::atoi(e) : 0x7fffffff; }();
  const int un = unroll == 8 ? 8 : (unroll == 2 ? 2 : 4);
  int64_t gx = (X.mfact + 256 * un - 1) / (256 * un), gy = rows;
  if (gx > gx_cap) gx = gx_cap;
  if (gx < 1) gx = 1;
  if (gy > 65535) gy = 65535;
  dim3 grid((unsigned)gx, (unsigned)gy);
  LaunchScope scope(kKindExchange, stream);
#define SNFFT_XCHG(U)                                                             \
  if (inverse)                                                                    \
    shard_exchange_kernel<true, U><<<grid, 256, 0, stream>>>(X, local);           \
  else                                                                            \
    shard_exchange_kernel<false, U><<<grid, 256, 0, stream>>>(X, local)
  if (un == 8) { SNFFT_XCHG(8); } else if (un == 2) { SNFFT_XCHG(2); } else { SNFFT_XCHG(4); }
#undef SNFFT_XCHG
  return cudaGetLastError();
}

cudaError_t launch_shard_barrier(const ShardBarrier& Bq, cudaStream_t stream) {
  LaunchScope scope(kKindBarrier, stream);
  shard_barrier_kernel<<<1, 32, 0, stream>>>(Bq);
  return cudaGetLastError();
}

cudaError_t launch_dgemm(const double* A, const double* B, double* C, int M, int N, int K,
                         cudaStream_t stream) {
  if (M <= 0 || N <= 0) return cudaSuccess;
  LaunchScope scope(kKindDgemm, stream);
  const bool aligned = ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) |
                         reinterpret_cast<uintptr_t>(C)) & 15) == 0;
  if (aligned && K > 0 && K % kG2BK == 0 && N % 2 == 0 && N >= 64 && (int64_t)M * N >= 128 * 128) {
    dim3 grid((N + kG2BN - 1) / kG2BN, (M + kG2BM - 1) / kG2BM);
    const size_t smem = (size_t)kG2Stages * kG2StageDoubles * sizeof(double);
    dgemm_pipe_kernel<<<grid, 256, smem, stream>>>(A, B, C, M, N, K);
    return cudaGetLastError();
  }
  dim3 grid((N + kGemmBN - 1) / kGemmBN, (M + kGemmBM - 1) / kGemmBM);
  dgemm_kernel<<<grid, 256, 0, stream>>>(A, B, C, M, N, K);
  return cudaGetLastError();
}

}  // namespace snfft

#ifdef SNFFT_PHASE_CLOCKS
extern "C" int snfft_debug_phase_cycles(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, snfft::g_phase_cycles, sizeof(unsigned long long) * 32);
}
#endif
